// Pairwise text x mel log-likelihood for sm_100a: one batched GEMM with the whole
// ConvAttention epilogue fused behind it.
//
// Reference semantics (tts/models/acoustic/modules/alignment.py in the reference):
//   :189 matmul  :190 scale  :192 clamp  :18-37,:195 batch_diagonal_prior
//   :196 log_softmax over ALL T2max columns + log(prior + 1e-6)   -> attn_logits (:198)
//   :201-206 masked softmax over the valid text columns, * mask   -> attn_soft
//
// One CTA computes one tile of 128 mel frames x all text tokens of one utterance:
//   warp 0  : TMA loads of the Q tile and the utterance's K operand (3-D tensor maps,
//             128 B swizzle, one mbarrier per 128 B-wide K-slab), then tcgen05.mma
//             (kind::f16 for bf16 operands, kind::tf32 for fp32 operands) issued by one
//             elected lane, accumulating S = Q.K^T in TMEM (128 lanes x <=512 columns).
//   warps 1-8: epilogue, two threads per frame: warps w and w+4 share a TMEM lane quadrant and take
//             the even / odd 16-column chunks; row statistics meet in shared memory.  The row never
//             leaves the SM: pass 1 reads S from TMEM and keeps an online max / sum-of-exp
//             (log_softmax denominator, padded columns added in closed form) plus the prior's row
//             sum; pass 2 re-reads S, forms attn_logits = S - lse + log(p + 1e-6) with the Gaussian
//             prior evaluated in registers, and stashes w = exp(S-m)*(p+1e-6) back into TMEM;
//             pass 3 turns w into attn_soft = w / sum(w).  Chunks are transposed through shared
//             memory so that global stores are 16 B, coalesced.
// Two CTAs are resident per SM (256 TMEM columns each) when T2max <= 256, so one CTA's
// loads and MMAs hide under the other's epilogue.  The kernel is bound by the 8 B/cell of
// fp32 output it must write, not by the tensor pipe.  Padding never
// reaches the arithmetic: a tile without a valid frame is a straight constant fill by the
// whole CTA, a chunk of padded tokens one constant per row.
//
// Contract on the operands: Q rows >= mel_len[b] and K rows >= text_len[b] are zero (the
// reference guarantees it, alignment.py:75-76).  The kernel uses it to treat padded text
// columns (S == 0) in closed form and to skip the GEMM for tiles that are all padding.

#include <cuda.h>
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "common.cuh"
#include "isp_internal.h"
#include "isp_prior.cuh"

namespace isp {

constexpr int kTileM = 128;
constexpr int kCW = 16;                         // columns per epilogue chunk (one tcgen05.ld.x16)
constexpr int kStagePitch = 20;                 // words per staged row: 16 B aligned, 5 x 16 B: conflict-free 128-bit stores
constexpr int kEpiWarps = 8;                    // two per TMEM lane quadrant
constexpr int kThreads = 32 * (1 + kEpiWarps);
constexpr int kMaxSlabs = 8;                    // 128 B-wide K-slabs (D * elem <= 1024 B)

struct LoglikParams {
    const int64_t* text_len;
    const int64_t* mel_len;
    float* logits;
    float* soft;
    int B, T1max, T2max, D;
    int npad;            // T2max rounded up to 16 (MMA N granularity)
    int nt;              // 256-column accumulator chunks (1 or 2)
    int boxrows_b;       // rows per TMA box of the K operand
    int kslabs;          // 128 B-wide slabs along D
    int ksteps;          // 32 B-wide MMA k-steps along D
    int elem;            // bytes per operand element (2 or 4)
    int tmem_cols;       // 256 or 512
    float scale;
    int prior;
    int vec4;            // T2max % 4 == 0 and outputs 16 B aligned
    uint32_t idesc_base; // instruction descriptor without N
    int debug_scores;    // 1: write scale*S into `logits`, zeros into `soft`
    float* psum_out;     // (B, T1max) row sums of the raw prior for the backward kernel (valid frames only are written), or nullptr
    unsigned long long* tstamp;   // debug (align.trace): [0] first CTA start, [1] last CTA end (%globaltimer), or nullptr
    int* ready;          // per utterance: tiles whose outputs are complete (isp_align_forward: the MAS kernel waits on it), or nullptr
};

// isp_align_forward: the MAS kernel takes an utterance's attn_logits as soon as every frame tile of it is in global memory.
// A release costs the round trip of the stores before it, and a CTA that waited for it at its end would hold its SM slot that
// much longer (measured: +10 % on the kernel).  So the publishing is warp 0's job -- it has nothing to do once the MMAs are
// issued: the epilogue warps arrive on a named barrier when their attn_logits stores are ISSUED (they go on to attn_soft), warp 0
// waits on that barrier and counts the tile with a release.
constexpr uint32_t kPublishBar = 6;
ISP_DEVINL void publish_arrive() { asm volatile("bar.arrive %0, %1;" ::"r"(kPublishBar), "r"(kThreads) : "memory"); }
ISP_DEVINL void publish_tile(int* ready, int b, int lane) {        // warp 0, converged
    asm volatile("bar.sync %0, %1;" ::"r"(kPublishBar), "r"(kThreads) : "memory");
    // (release at device scope, cumulative over the stores the barrier ordered before this thread)
    if (lane == 0) asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(ready + b) : "memory");
    __syncwarp();
}

// ---- tcgen05 / TMA wrappers ---------------------------------------------------------
ISP_DEVINL void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
        : "memory");
}
ISP_DEVINL void tmem_alloc(uint32_t* slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
ISP_DEVINL void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
ISP_DEVINL void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
ISP_DEVINL void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <bool TF32>
ISP_DEVINL void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if (TF32) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
    }
}
ISP_DEVINL void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// K-major operand, 128 B swizzle: rows of 128 B, 8-row groups 1024 B apart (SBO), version 1.
ISP_DEVINL uint64_t smem_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= uint64_t((saddr & 0x3ffff) >> 4);       // start address, 16 B units
    d |= uint64_t(1) << 16;                      // leading byte offset (unused for swizzled K-major)
    d |= uint64_t(1024 >> 4) << 32;              // stride byte offset
    d |= uint64_t(1) << 46;                      // descriptor version (Blackwell)
    d |= uint64_t(2) << 61;                      // SWIZZLE_128B
    return d;
}
ISP_DEVINL void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int k = 0; k < 32; ++k) v[k] = __uint_as_float(r[k]);
}
ISP_DEVINL void tmem_st32(uint32_t taddr, const float (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr),
          "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
          "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
          "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
          "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
          "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
          "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
          "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])),
          "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
ISP_DEVINL void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] = __uint_as_float(r[k]);
}
ISP_DEVINL void tmem_st16(uint32_t taddr, const float (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr),
          "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
          "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
          "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
          "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
ISP_DEVINL float fast_lg2(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// ---- staged, coalesced store of one 32-row x 16-column chunk -------------------------
// `stage` holds the warp's 32 rows (row = lane that produced it), kStagePitch words apart.
ISP_DEVINL void store_chunk(const float* stage, float* gbase, int lane, int row0, int rows_valid,
                            int j0, int T2max, bool vec4) {
    // gbase points at element (b, 0, 0); row0 is the global frame index of staged row 0
    if (vec4) {
        const int c4 = (lane & 3) * 4;
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int rr = (lane >> 2) + 8 * it;
            if (rr < rows_valid && j0 + c4 < T2max) {
                const float4 t = *reinterpret_cast<const float4*>(stage + rr * kStagePitch + c4);
                *reinterpret_cast<float4*>(gbase + size_t(row0 + rr) * T2max + j0 + c4) = t;
            }
        }
    } else {
        const int cc = lane & 15;
        const bool colok = j0 + cc < T2max;
        for (int rr = lane >> 4; rr < rows_valid; rr += 2) {
            if (colok) gbase[size_t(row0 + rr) * T2max + j0 + cc] = stage[rr * kStagePitch + cc];
        }
    }
}

// DBG: the debug options (loglik.debug_scores, align.trace) live in an instantiation of their own, so that the production kernel
// does not carry their branches (in the MAS kernels, instrumentation that is compiled in but never executed cost 3-20 %)
template <bool TF32, bool DBG>
__global__ void __launch_bounds__(kThreads, 2)
loglik_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
              const LoglikParams pp) {
    struct P : LoglikParams {
        __device__ P(const LoglikParams& o) : LoglikParams(o) { if (!DBG) { debug_scores = 0; tstamp = nullptr; } }
    };
    const P p(pp);
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    // A kernel launched behind this one with programmatic stream serialisation (isp_mas_forward does, see isp_mas2.cu) may become
    // resident on the SMs this grid's last wave leaves free and run its set-up there; it waits for this grid's completion
    // (griddepcontrol.wait) before it touches global memory.  Without such a dependent this is a no-op.
    if (p.tstamp != nullptr && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        atomicMin(p.tstamp, t);
    }
    if (p.ready != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
        // the linked MAS launch's time origin (the word behind the ready counts): unset until its first CTA starts
        *reinterpret_cast<unsigned long long*>(p.ready + ((p.B + 1) & ~1)) = 0ull;
        __threadfence();
    }
    asm volatile("griddepcontrol.launch_dependents;");
    // utterance-major launch order on purpose: it interleaves arithmetic tiles with the straight fills of padded tiles,
    // so the two kinds share an SM and HBM writes overlap the epilogue (tile-major order measured 4 % slower on cfg3)
    const int mt = blockIdx.x;       // frame tile
    const int b = blockIdx.y;        // utterance

    // ---- shared memory carve-up --------------------------------------------------------
    const uint32_t a_slab_bytes = kTileM * 128;
    const uint32_t b_slab_bytes = uint32_t(p.nt) * p.boxrows_b * 128;
    // the 128 B swizzle atoms need 1024 B alignment; the launch adds 1 KB of slack for this
    unsigned char* smem_a = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* smem_b = smem_a + size_t(p.kslabs) * a_slab_bytes;
    float* stage_all = reinterpret_cast<float*>(smem_b + size_t(p.kslabs) * b_slab_bytes);   // [8 warps][32][kStagePitch]
    float* gt = stage_all + kEpiWarps * 32 * kStagePitch;               // [npad] j / T2_b
    float4* xch = reinterpret_cast<float4*>(gt + ((p.npad + 31) & ~31)); // [2 halves][128 rows]: partial row statistics
    uint64_t* bars = reinterpret_cast<uint64_t*>(xch + 2 * kTileM);
    uint64_t* slab_full = bars;                                         // [kMaxSlabs]
    uint64_t* mma_done = bars + kMaxSlabs;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kMaxSlabs + 1);

    long long n64 = p.mel_len[b], m64 = p.text_len[b];
    const int T1b = int(n64 < 1 ? 1 : (n64 > p.T1max ? p.T1max : n64));
    const int T2b = int(m64 < 1 ? 1 : (m64 > p.T2max ? p.T2max : m64));
    const int row_tile0 = mt * kTileM;
    const bool all_padding = row_tile0 >= T1b && !p.debug_scores;       // no valid frame in this tile
    const int nb = min(p.npad, (T2b + 15) & ~15);                       // MMA / epilogue column extent

    if (threadIdx.x == 0) {
        for (int s = 0; s < kMaxSlabs; ++s) mbar_init(&slab_full[s], 1);
        mbar_init(mma_done, 1);
        fence_mbar_init();
    }
    if (warp == 0 && !all_padding) tmem_alloc(tmem_slot, p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = all_padding ? 0u : *tmem_slot;

    if (all_padding) {
        // S == 0 on the whole tile: lse = log(T2max); attn_logits = -lse + log(1e-6); attn_soft = 0.  The tile's rows are
        // one contiguous block of each output: a straight coalesced fill by every thread of the CTA.
        const float cst = p.prior ? kLogPriorFloor - logf(float(p.T2max)) : 0.0f;
        const int rows = min(kTileM, p.T1max - row_tile0);
        const size_t off = (size_t(b) * p.T1max + row_tile0) * p.T2max;
        const size_t nelem = size_t(rows) * p.T2max;
        float* gl = p.logits + off;
        float* gs = p.soft + off;
        if (p.ready == nullptr) {
            if (p.vec4) {
                const float4 c4 = make_float4(cst, cst, cst, cst), z4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                const size_t n4 = nelem >> 2;
                for (size_t idx = threadIdx.x; idx < n4; idx += kThreads) {
                    __stcs(reinterpret_cast<float4*>(gl) + idx, c4);
                    __stcs(reinterpret_cast<float4*>(gs) + idx, z4);
                }
            } else {
                for (size_t idx = threadIdx.x; idx < nelem; idx += kThreads) { gl[idx] = cst; gs[idx] = 0.0f; }
            }
            return;
        }
        // linked to the MAS kernel: attn_logits first, published by warp 0 while the others write attn_soft
        if (p.vec4) {
            const float4 c4 = make_float4(cst, cst, cst, cst);
            for (size_t idx = threadIdx.x; idx < (nelem >> 2); idx += kThreads) __stcs(reinterpret_cast<float4*>(gl) + idx, c4);
        } else {
            for (size_t idx = threadIdx.x; idx < nelem; idx += kThreads) gl[idx] = cst;
        }
        if (warp == 0) publish_tile(p.ready, b, lane);
        else publish_arrive();
        if (p.vec4) {
            const float4 z4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            for (size_t idx = threadIdx.x; idx < (nelem >> 2); idx += kThreads) __stcs(reinterpret_cast<float4*>(gs) + idx, z4);
        } else {
            for (size_t idx = threadIdx.x; idx < nelem; idx += kThreads) gs[idx] = 0.0f;
        }
        return;
    }

    if (warp == 0) {
        // ============================ TMA + MMA issue ===================================
        if (!all_padding && lane == 0) {
            const int ke = 128 / p.elem;   // elements per slab row
            for (int kb = 0; kb < p.kslabs; ++kb) {
                mbar_arrive_expect_tx(&slab_full[kb], a_slab_bytes + b_slab_bytes);
                tma_load_3d(smem_a + size_t(kb) * a_slab_bytes, &tmap_q, kb * ke, row_tile0, b, &slab_full[kb]);
                for (int t = 0; t < p.nt; ++t)
                    tma_load_3d(smem_b + size_t(kb) * b_slab_bytes + size_t(t) * p.boxrows_b * 128, &tmap_k,
                                kb * ke, t * p.boxrows_b, b, &slab_full[kb]);
            }
            int ks_done = 0;
            for (int kb = 0; kb < p.kslabs; ++kb) {
                mbar_wait(&slab_full[kb], 0);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(smem_a + size_t(kb) * a_slab_bytes);
                const uint32_t b_addr = smem_u32(smem_b + size_t(kb) * b_slab_bytes);
                const int ks_here = min(4, p.ksteps - kb * 4);     // 4 k-steps of 32 B per 128 B slab
                for (int t = 0; t < p.nt; ++t) {
                    const int ncols = min(256, nb - t * 256);
                    if (ncols <= 0) break;
                    const uint32_t idesc = p.idesc_base | (uint32_t(ncols >> 3) << 17);
                    for (int k = 0; k < ks_here; ++k) {
                        const uint64_t ad = smem_desc_sw128(a_addr + k * 32);
                        const uint64_t bd = smem_desc_sw128(b_addr + uint32_t(t) * p.boxrows_b * 128 + k * 32);
                        umma<TF32>(tmem_base + t * 256, ad, bd, idesc, (ks_done + k) > 0 ? 1u : 0u);
                    }
                }
                ks_done += ks_here;
            }
            umma_commit(mma_done);
        }
        __syncwarp();
        if (p.ready != nullptr) publish_tile(p.ready, b, lane);
    } else {
        // ============================ epilogue: two threads per frame ===================
        // Warps w and w + 4 share a TMEM lane quadrant (a warp may only touch lanes 32 (warp % 4) ..) and split the
        // row's 16-column chunks between them: even chunks to the first "half", odd ones to the second.  Row statistics
        // are combined through shared memory, twice per tile.  Eight epilogue warps per CTA, sixteen per SM: the epilogue
        // is a chain of MUFU and TMEM latencies, and only more warps hide it.
        const int quad = warp & 3;                       // TMEM lane quadrant this warp may touch
        const int half = (warp - 1) >> 2;                // 0: chunks 0, 2, 4, ...   1: chunks 1, 3, 5, ...
        const int r_in_tile = quad * 32 + lane;
        const int i = row_tile0 + r_in_tile;             // frame index
        const int warp_row0 = row_tile0 + quad * 32;
        const int rows_valid = max(0, min(32, p.T1max - warp_row0));
        float* stage = stage_all + (warp - 1) * 32 * kStagePitch;
        float* my_stage = stage + lane * kStagePitch;
        float* g_logits = p.logits + size_t(b) * p.T1max * p.T2max;
        float* g_soft = p.soft + size_t(b) * p.T1max * p.T2max;
        const bool vec4 = p.vec4 != 0;
        const int nchunks_all = (p.T2max + kCW - 1) / kCW;

        {
            // j / T2_b exactly as the reference divides (alignment.py:22), once per CTA, stored pre-scaled by
            // sqrt(50 log2 e): the Gaussian prior exp(-(g - u)^2 / (2 * 0.1^2)) is then ex2(-d * d) with d = gs[j] - us
            const float t2f = float(T2b);
            for (int j = threadIdx.x - 32; j < p.npad; j += 32 * kEpiWarps) gt[j] = __fdiv_rn(float(j), t2f) * kPriorScale;
            asm volatile("bar.sync 1, 256;" ::: "memory");   // epilogue warps only

            const bool row_valid = i < T1b;
            const float us = __fdiv_rn(float(i), float(T1b)) * kPriorScale;   // alignment.py:25
            const float c = p.scale * kLog2e;
            const uint32_t tlane = tmem_base + (uint32_t(quad * 32) << 16);
            const int nchunks = nb / kCW;                      // chunks that hold MMA output (nb is a multiple of 16)
            const uint32_t pair_bar = 2u + uint32_t(quad);     // the two warps of a TMEM quadrant meet on their own barrier

            // per-thread pieces of the staged, coalesced stores (see store_chunk_fast)
            const int c4 = (lane & 3) * 4;
            const int rb = lane >> 2;
            const float* st_rd = stage + rb * kStagePitch + c4;
            const size_t g_off = size_t(warp_row0 + rb) * p.T2max + c4;
            const size_t g_rs = size_t(8) * p.T2max;
            uint32_t rowok = 0;
#pragma unroll
            for (int it = 0; it < 4; ++it) rowok |= (rb + 8 * it < rows_valid) ? (1u << it) : 0u;
            auto store_chunk_fast = [&](float* gout, int j0) __attribute__((always_inline)) {
                if (vec4) {
                    if (j0 + c4 < p.T2max) {
                        float* gp = gout + g_off + j0;
#pragma unroll
                        for (int it = 0; it < 4; ++it) {
                            if (rowok & (1u << it)) {
                                const float4 t = *reinterpret_cast<const float4*>(st_rd + it * 8 * kStagePitch);
                                *reinterpret_cast<float4*>(gp + it * g_rs) = t;
                            }
                        }
                    }
                } else {
                    store_chunk(stage, gout, lane, warp_row0, rows_valid, j0, p.T2max, false);
                }
            };

            float v[kCW];
            if (warp_row0 >= T1b && !p.debug_scores) {
                // the 32 frames of this quadrant are all padding (Q rows == 0): S == 0, so the rows are the constant of an
                // all-padding tile.  The quadrant's two warps fill one output each and skip the passes (they only ever
                // meet each other on the pair barrier).
                const float cst = (half == 0 && p.prior) ? kLogPriorFloor - logf(float(p.T2max)) : 0.0f;
                float* gout = (half == 0 ? g_logits : g_soft) + size_t(warp_row0) * p.T2max;
                const int nelem = rows_valid * p.T2max;
                if (p.ready != nullptr && half == 1) publish_arrive();         // (this warp writes attn_soft only)
                if (vec4) {
                    const float4 c4v = make_float4(cst, cst, cst, cst);
                    for (int idx = lane; idx < (nelem >> 2); idx += 32) __stcs(reinterpret_cast<float4*>(gout) + idx, c4v);
                } else {
                    for (int idx = lane; idx < nelem; idx += 32) gout[idx] = cst;
                }
                if (p.ready != nullptr && half == 0) publish_arrive();
            } else if (p.debug_scores) {
                mbar_wait(mma_done, 0);
                tc_fence_after();
                for (int ch = half; ch < nchunks_all; ch += 2) {
                    if (ch < nchunks) tmem_ld16(tlane + ch * kCW, v);
#pragma unroll
                    for (int k = 0; k < kCW; ++k) my_stage[k] = (ch < nchunks && ch * kCW + k < nb) ? v[k] * p.scale : 0.0f;
                    __syncwarp();
                    store_chunk(stage, g_logits, lane, warp_row0, rows_valid, ch * kCW, p.T2max, vec4);
                    __syncwarp();
#pragma unroll
                    for (int k = 0; k < kCW; ++k) my_stage[k] = 0.0f;
                    __syncwarp();
                    store_chunk(stage, g_soft, lane, warp_row0, rows_valid, ch * kCW, p.T2max, vec4);
                    __syncwarp();
                }
                if (p.ready != nullptr) publish_arrive();
            } else {
                // ---- pass 0: the prior's row sum over this half's valid columns.  It does not depend on the scores, so it
                // runs while the operands are still in flight (the wait for the MMA comes after it) -------------------
                float psum = 0.0f;
                if (p.prior) {
                    float ps0 = 0.0f, ps1 = 0.0f;
                    for (int ch = half; ch < nchunks; ch += 2) {
                        const int j0 = ch * kCW;
                        const int kmax = min(kCW, T2b - j0);
                        if (kmax <= 0) break;
                        // terms below ~1e-11 of the peak cannot change an fp32 sum: skip far chunks (|g - u| > 0.71)
                        const float dlo = gt[j0] - us, dhi = gt[j0 + kmax - 1] - us;
                        const bool near = row_valid && dlo <= 0.71f * kPriorScale && dhi >= -0.71f * kPriorScale;
                        if (!__any_sync(0xffffffffu, near)) continue;
                        if (kmax == kCW) {
#pragma unroll
                            for (int k4 = 0; k4 < kCW / 4; ++k4) {
                                const float4 g4 = *reinterpret_cast<const float4*>(gt + j0 + 4 * k4);
                                const float d0 = g4.x - us, d1 = g4.y - us, d2 = g4.z - us, d3 = g4.w - us;
                                ps0 += fast_ex2(-d0 * d0); ps1 += fast_ex2(-d1 * d1);
                                ps0 += fast_ex2(-d2 * d2); ps1 += fast_ex2(-d3 * d3);
                            }
                        } else {
                            for (int k = 0; k < kmax; ++k) { const float d = gt[j0 + k] - us; ps0 += fast_ex2(-d * d); }
                        }
                    }
                    psum = row_valid ? ps0 + ps1 : 0.0f;
                }

                mbar_wait(mma_done, 0);
                tc_fence_after();

                // ---- pass 1: online max / sum of exp over this half's valid columns -------------------------------
                float m = -CUDART_INF_F, sum_e = 0.0f;
                for (int ch = half; ch < nchunks; ch += 2) {
                    const int j0 = ch * kCW;
                    const int kmax = min(kCW, T2b - j0);
                    if (kmax <= 0) break;
                    tmem_ld16(tlane + j0, v);
                    float cm = -CUDART_INF_F, acc0 = 0.0f, acc1 = 0.0f;
                    if (kmax == kCW) {                          // warp-uniform: no column masks in the common case
#pragma unroll
                        for (int k = 0; k < kCW; ++k) cm = fmaxf(cm, v[k]);
                        const float mn = fmaxf(m, cm);
                        const float mnc = mn * c;
#pragma unroll
                        for (int k = 0; k < kCW; k += 2) {
                            acc0 += fast_ex2(fmaf(v[k], c, -mnc));
                            acc1 += fast_ex2(fmaf(v[k + 1], c, -mnc));
                        }
                        sum_e = fmaf(sum_e, fast_ex2((m - mn) * c), acc0 + acc1);
                        m = mn;
                    } else {
#pragma unroll
                        for (int k = 0; k < kCW; ++k) cm = fmaxf(cm, k < kmax ? v[k] : -CUDART_INF_F);
                        const float mn = fmaxf(m, cm);
                        const float mnc = mn * c;
#pragma unroll
                        for (int k = 0; k < kCW; ++k) {
                            const float e = fast_ex2(fmaf(v[k], c, -mnc));
                            acc0 += k < kmax ? e : 0.0f;
                        }
                        sum_e = fmaf(sum_e, fast_ex2((m - mn) * c), acc0);
                        m = mn;
                    }
                }
                // combine the two halves' statistics
                xch[half * kTileM + r_in_tile] = make_float4(m, sum_e, psum, 0.0f);
                asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
                {
                    const float4 o = xch[(half ^ 1) * kTileM + r_in_tile];
                    const float mn = fmaxf(m, o.x);
                    // a half with no valid chunk reports m = -inf, sum 0: its factor must not become NaN
                    const float f0 = m == -CUDART_INF_F ? 0.0f : fast_ex2((m - mn) * c);
                    const float f1 = o.x == -CUDART_INF_F ? 0.0f : fast_ex2((o.x - mn) * c);
                    sum_e = sum_e * f0 + o.y * f1;
                    psum = half == 0 ? psum + o.z : o.z + psum;       // the same order of addition in both halves
                    m = mn;
                }
                // the backward pass re-derives the prior's cells from this sum (isp_loglik_backward_from_logits)
                if (p.psum_out != nullptr && half == 0 && row_valid) p.psum_out[size_t(b) * p.T1max + i] = psum;
                if (T2b < p.T2max) {
                    // padded text columns have S == 0 exactly: add them in closed form
                    const float mn = fmaxf(m, 0.0f);
                    sum_e = sum_e * fast_ex2((m - mn) * c) + float(p.T2max - T2b) * fast_ex2(-mn * c);
                    m = mn;
                }
                const float lse = m * p.scale + logf(sum_e);
                const float mc = m * c;
                const float inv_psum = 1.0f / (psum + 1e-5f);                  // alignment.py:34
                // cells with (g - u)^2 above this cannot pass the 1e-4 threshold (alignment.py:35)
                const float thr_arg = kPriorThreshold * (psum + 1e-5f);
                const float dband = (row_valid && p.prior && thr_arg < 1.0f)
                                        ? (sqrtf(-logf(thr_arg) * (1.0f / 50.0f)) * 1.001f + 1e-6f) * kPriorScale
                                        : -1.0f;
                const float cst_pad = p.prior ? kLogPriorFloor - lse : 0.0f;   // attn_logits of a padded text column (S == 0)
                // chunks of padded text columns are written straight from registers (the row constants come by shuffle)
                auto fill_chunk = [&](float* gout, int j0, bool logits) __attribute__((always_inline)) {
                    const bool colok = j0 + c4 < p.T2max;
                    {
                        float* gp = gout + g_off + j0;
#pragma unroll
                        for (int it = 0; it < 4; ++it) {
                            const float cv = logits ? __shfl_sync(0xffffffffu, cst_pad, rb + 8 * it) : 0.0f;
                            if (colok && (rowok & (1u << it))) __stcs(reinterpret_cast<float4*>(gp + it * g_rs), make_float4(cv, cv, cv, cv));
                        }
                    }
                };

                // ---- pass 2: attn_logits, and w = exp(S - m) * (p + 1e-6) stashed in TMEM -------
                float sw0 = 0.0f, sw1 = 0.0f;
                for (int ch = half; ch < nchunks_all; ch += 2) {
                    const int j0 = ch * kCW;
                    const int kmax = min(kCW, T2b - j0);                       // valid text columns here
                    if (kmax <= 0) {
                        // padded text columns only: S == 0, so attn_logits is one constant per row and w is 0
                        if (vec4) { fill_chunk(g_logits, j0, true); continue; }
                        const float4 c4v = make_float4(cst_pad, cst_pad, cst_pad, cst_pad);
#pragma unroll
                        for (int k4 = 0; k4 < kCW / 4; ++k4) *reinterpret_cast<float4*>(my_stage + 4 * k4) = c4v;
                        __syncwarp();
                        store_chunk_fast(g_logits, j0);
                        __syncwarp();
                        continue;
                    }
                    if (ch < nchunks) {
                        tmem_ld16(tlane + j0, v);
                    } else {
#pragma unroll
                        for (int k = 0; k < kCW; ++k) v[k] = 0.0f;             // beyond the MMA extent: S == 0
                    }
                    if (!p.prior) {
#pragma unroll
                        for (int k = 0; k < kCW; ++k) {
                            my_stage[k] = v[k] * p.scale;
                            const float e = fast_ex2(fmaf(v[k], c, -mc));
                            v[k] = (k < kmax && row_valid) ? e : 0.0f;
                            sw0 += v[k];
                        }
                    } else {
                        const float dlo = gt[j0] - us, dhi = gt[j0 + kmax - 1] - us;
                        const bool near = dlo <= dband && dhi >= -dband;
                        // warp-uniform: every column of the chunk is a valid token and every row of the warp a valid frame
                        const bool full = kmax == kCW && __all_sync(0xffffffffu, row_valid);
                        if (__any_sync(0xffffffffu, near)) {
                            if (full) {
#pragma unroll
                                for (int k4 = 0; k4 < kCW / 4; ++k4) {
                                    const float4 g4 = *reinterpret_cast<const float4*>(gt + j0 + 4 * k4);
                                    const float dd[4] = {g4.x - us, g4.y - us, g4.z - us, g4.w - us};
                                    float lg[4];
#pragma unroll
                                    for (int q = 0; q < 4; ++q) {
                                        const int k = 4 * k4 + q;
                                        float pr = fast_ex2(-dd[q] * dd[q]) * inv_psum;
                                        pr = pr >= kPriorThreshold ? pr : 0.0f;
                                        const float P = pr + kPriorEps;
                                        lg[q] = fmaf(fast_lg2(P), kLn2, fmaf(v[k], p.scale, -lse));
                                        v[k] = fast_ex2(fmaf(v[k], c, -mc)) * P;
                                    }
                                    *reinterpret_cast<float4*>(my_stage + 4 * k4) = make_float4(lg[0], lg[1], lg[2], lg[3]);
                                    sw0 += v[4 * k4]; sw1 += v[4 * k4 + 1];
                                    sw0 += v[4 * k4 + 2]; sw1 += v[4 * k4 + 3];
                                }
                            } else {
#pragma unroll
                                for (int k4 = 0; k4 < kCW / 4; ++k4) {
                                    const float4 g4 = *reinterpret_cast<const float4*>(gt + j0 + 4 * k4);
                                    const float dd[4] = {g4.x - us, g4.y - us, g4.z - us, g4.w - us};
#pragma unroll
                                    for (int q = 0; q < 4; ++q) {
                                        const int k = 4 * k4 + q;
                                        const bool ok = k < kmax && row_valid;
                                        float pr = fast_ex2(-dd[q] * dd[q]) * inv_psum;
                                        pr = (ok && pr >= kPriorThreshold) ? pr : 0.0f;
                                        const float P = pr + kPriorEps;
                                        my_stage[k] = fmaf(fast_lg2(P), kLn2, fmaf(v[k], p.scale, -lse));
                                        const float e = fast_ex2(fmaf(v[k], c, -mc));
                                        v[k] = ok ? e * P : 0.0f;
                                        sw0 += v[k];
                                    }
                                }
                            }
                        } else if (full) {
                            const float lse_f = lse - kLogPriorFloor;
#pragma unroll
                            for (int k4 = 0; k4 < kCW / 4; ++k4) {
                                float lg[4];
#pragma unroll
                                for (int q = 0; q < 4; ++q) {
                                    const int k = 4 * k4 + q;
                                    lg[q] = fmaf(v[k], p.scale, -lse_f);
                                    v[k] = fast_ex2(fmaf(v[k], c, -mc)) * kPriorEps;
                                }
                                *reinterpret_cast<float4*>(my_stage + 4 * k4) = make_float4(lg[0], lg[1], lg[2], lg[3]);
                                sw0 += v[4 * k4]; sw1 += v[4 * k4 + 1];
                                sw0 += v[4 * k4 + 2]; sw1 += v[4 * k4 + 3];
                            }
                        } else {
#pragma unroll
                            for (int k = 0; k < kCW; ++k) {
                                my_stage[k] = fmaf(v[k], p.scale, -lse) + kLogPriorFloor;
                                const float e = fast_ex2(fmaf(v[k], c, -mc));
                                v[k] = (k < kmax && row_valid) ? e * kPriorEps : 0.0f;
                                sw0 += v[k];
                            }
                        }
                    }
                    if (ch < nchunks) tmem_st16(tlane + j0, v);
                    __syncwarp();
                    store_chunk_fast(g_logits, j0);
                    __syncwarp();
                }
                if (p.ready != nullptr) publish_arrive();          // this warp's attn_logits stores are issued

                // ---- pass 3: attn_soft = w / sum(w) on valid cells, 0 elsewhere -----------------
                // (.w alone is written: the partner may still be reading x, y, z of the pass-1 exchange)
                float sum_w = sw0 + sw1;
                xch[half * kTileM + r_in_tile].w = sum_w;
                asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
                {
                    const float o = xch[(half ^ 1) * kTileM + r_in_tile].w;
                    sum_w = half == 0 ? sum_w + o : o + sum_w;
                }
                const float inv_w = row_valid ? 1.0f / sum_w : 0.0f;
                tc_fence_before();
                __syncwarp();
                tc_fence_after();
                for (int ch = half; ch < nchunks_all; ch += 2) {
                    const int j0 = ch * kCW;
                    const int kmax = min(kCW, T2b - j0);
                    if (ch < nchunks && kmax > 0) {
                        tmem_ld16(tlane + j0, v);
                        if (kmax == kCW) {
#pragma unroll
                            for (int k4 = 0; k4 < kCW / 4; ++k4)
                                *reinterpret_cast<float4*>(my_stage + 4 * k4) =
                                    make_float4(v[4 * k4] * inv_w, v[4 * k4 + 1] * inv_w, v[4 * k4 + 2] * inv_w, v[4 * k4 + 3] * inv_w);
                        } else {
#pragma unroll
                            for (int k = 0; k < kCW; ++k) my_stage[k] = k < kmax ? v[k] * inv_w : 0.0f;
                        }
                    } else {
                        if (vec4) { fill_chunk(g_soft, j0, false); continue; }
                        const float4 z4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll
                        for (int k4 = 0; k4 < kCW / 4; ++k4) *reinterpret_cast<float4*>(my_stage + 4 * k4) = z4;
                    }
                    __syncwarp();
                    store_chunk_fast(g_soft, j0);
                    __syncwarp();
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0 && !all_padding) {
        tc_fence_after();
        tmem_dealloc(tmem_base, p.tmem_cols);
    }
    if (p.tstamp != nullptr && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        atomicMax(p.tstamp + 1, t);
    }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    if (fn) return fn;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess || !ptr) return nullptr;
    fn = reinterpret_cast<PFN_encodeTiled>(ptr);
    return fn;
}

static int g_opt_debug_scores = 0;
static int g_opt_trace = 0;

int loglik_set_option(const char* key, int value, int* prev) {
    if (!strcmp(key, "align.trace")) { *prev = g_opt_trace; g_opt_trace = value; return 0; }
    if (!strcmp(key, "loglik.debug_scores")) { *prev = g_opt_debug_scores; g_opt_debug_scores = value; return 0; }
    return -1;
}

// the workspace receives the prior's row sums, one float per frame (what isp_loglik_backward_from_logits takes)
size_t loglik_workspace_bytes(int B, int T1max, int, int, int) { return B > 0 && T1max > 0 ? size_t(B) * T1max * sizeof(float) : 0; }
// shared memory of one CTA: operand slabs, the epilogue's staging, the prior's table, the row-statistics exchange, barriers
static size_t loglik_smem_bytes(int T2max, int D, int elem) {
    const int npad = (T2max + 15) & ~15, nt = (npad + 255) / 256, boxrows_b = nt == 1 ? npad : 256, kslabs = (D * elem + 127) / 128;
    return size_t(kslabs) * (kTileM * 128 + size_t(nt) * boxrows_b * 128)
         + sizeof(float) * (kEpiWarps * 32 * kStagePitch + ((npad + 31) & ~31) + 4 * 2 * kTileM)
         + sizeof(uint64_t) * (kMaxSlabs + 2) + 1024 /* base alignment slack */;
}

// does the fused kernel cover this shape?  (isp_loglik_supported: callers route the rest through isp_gemm_batched + isp_loglik_rows)
bool loglik_supported(int T2max, int D, int dtype) {
    if (dtype != ISP_DTYPE_F32 && dtype != ISP_DTYPE_BF16) return false;
    const int elem = dtype == ISP_DTYPE_BF16 ? 2 : 4;
    if (T2max <= 0 || D <= 0 || D % 8 != 0 || D > ISP_LOGLIK_MAX_D || T2max > ISP_LOGLIK_MAX_T2) return false;
    if ((D * elem + 127) / 128 > kMaxSlabs) return false;
    return loglik_smem_bytes(T2max, D, elem) <= 227 * 1024;
}

int loglik_tiles_per_utterance(int T1max) { return (T1max + kTileM - 1) / kTileM; }

// operand tensor map: (D, T, B) elements, box (128 B worth of D, rows, 1), 128 B swizzle
static int make_map(CUtensorMap* map, const void* base, int dtype, int D, int T, int B, int boxrows) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled not available from the driver"); return ISP_ERR_DEVICE; }
    const int elem = dtype == ISP_DTYPE_BF16 ? 2 : 4;
    cuuint64_t dims[3] = {cuuint64_t(D), cuuint64_t(T), cuuint64_t(B)};
    cuuint64_t strides[2] = {cuuint64_t(D) * elem, cuuint64_t(D) * elem * cuuint64_t(T)};
    cuuint32_t box[3] = {cuuint32_t(128 / elem), cuuint32_t(boxrows), 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, dtype == ISP_DTYPE_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                     3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d (D=%d T=%d B=%d)", int(r), D, T, B); return ISP_ERR_INVALID; }
    return 0;
}

int loglik_forward(const void* Q, const void* K, int dtype, const int64_t* text_len, const int64_t* mel_len,
                   int B, int T1max, int T2max, int D, float scale, int attention_prior,
                   float* attn_logits, float* attn_soft, void* ws, size_t ws_bytes, cudaStream_t stream, int* ready) {
    if (!Q || !K || !text_len || !mel_len || !attn_logits || !attn_soft) { set_error("isp_loglik_forward: null pointer"); return ISP_ERR_INVALID; }
    if (B <= 0 || T1max <= 0 || T2max <= 0 || D <= 0) { set_error("isp_loglik_forward: sizes must be positive"); return ISP_ERR_INVALID; }
    if (dtype != ISP_DTYPE_F32 && dtype != ISP_DTYPE_BF16) { set_error("isp_loglik_forward: dtype must be ISP_DTYPE_F32 or ISP_DTYPE_BF16"); return ISP_ERR_INVALID; }
    const int elem = dtype == ISP_DTYPE_BF16 ? 2 : 4;
    if (D % 8 != 0 || D > ISP_LOGLIK_MAX_D) { set_error("isp_loglik_forward: attention_dim D=%d must be a multiple of 8 and <= %d", D, ISP_LOGLIK_MAX_D); return ISP_ERR_UNSUPPORTED; }
    if (T2max > ISP_LOGLIK_MAX_T2) { set_error("isp_loglik_forward: T2max=%d > %d text tokens is not covered", T2max, ISP_LOGLIK_MAX_T2); return ISP_ERR_UNSUPPORTED; }
    if (B > 65535) { set_error("isp_loglik_forward: B=%d > 65535", B); return ISP_ERR_UNSUPPORTED; }
    if ((reinterpret_cast<uintptr_t>(Q) & 15) || (reinterpret_cast<uintptr_t>(K) & 15)) { set_error("isp_loglik_forward: Q and K must be 16 B aligned"); return ISP_ERR_INVALID; }
    if (!(scale > 0.0f)) { set_error("isp_loglik_forward: scale must be positive"); return ISP_ERR_INVALID; }

    LoglikParams p;
    p.text_len = text_len; p.mel_len = mel_len; p.logits = attn_logits; p.soft = attn_soft;
    p.B = B; p.T1max = T1max; p.T2max = T2max; p.D = D;
    p.npad = (T2max + 15) & ~15;
    p.nt = (p.npad + 255) / 256;
    p.boxrows_b = p.nt == 1 ? p.npad : 256;
    p.elem = elem;
    p.kslabs = (D * elem + 127) / 128;
    p.ksteps = (D * elem + 31) / 32;
    p.tmem_cols = p.nt == 1 ? 256 : 512;
    p.scale = scale;
    p.prior = attention_prior ? 1 : 0;
    p.vec4 = (T2max % 4 == 0 && (reinterpret_cast<uintptr_t>(attn_logits) & 15) == 0 && (reinterpret_cast<uintptr_t>(attn_soft) & 15) == 0) ? 1 : 0;
    p.debug_scores = g_opt_debug_scores;
    p.ready = ready;
    p.psum_out = (ws != nullptr && ws_bytes >= size_t(B) * T1max * sizeof(float) && (reinterpret_cast<uintptr_t>(ws) & 3) == 0) ? static_cast<float*>(ws) : nullptr;
    p.tstamp = (ready != nullptr && g_opt_trace) ? reinterpret_cast<unsigned long long*>(ready + ((B + 1) & ~1)) + 1 : nullptr;
    const uint32_t fmt = dtype == ISP_DTYPE_BF16 ? 1u : 2u;          // UMMA F16F32Format: BF16 = 1, TF32 = 2
    p.idesc_base = (1u << 4) | (fmt << 7) | (fmt << 10) | (uint32_t(kTileM >> 4) << 24);   // D=f32, A/B K-major
    if (p.kslabs > kMaxSlabs) { set_error("isp_loglik_forward: D * elem = %d B exceeds %d B", D * elem, kMaxSlabs * 128); return ISP_ERR_UNSUPPORTED; }

    CUtensorMap mq, mk;
    int rc = make_map(&mq, Q, dtype, D, T1max, B, kTileM);
    if (rc) return rc;
    rc = make_map(&mk, K, dtype, D, T2max, B, p.boxrows_b);
    if (rc) return rc;

    const size_t smem = loglik_smem_bytes(T2max, D, elem);
    if (smem > 227 * 1024) { set_error("isp_loglik_forward: needs %zu B of shared memory (T2max=%d, D=%d)", smem, T2max, D); return ISP_ERR_UNSUPPORTED; }

    const dim3 grid((T1max + kTileM - 1) / kTileM, B);
    const bool dbg = p.debug_scores != 0 || p.tstamp != nullptr;
    cudaError_t e;
    if (dtype == ISP_DTYPE_F32) {
        auto kern = dbg ? loglik_kernel<true, true> : loglik_kernel<true, false>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(loglik_kernel<tf32>)");
        kern<<<grid, kThreads, smem, stream>>>(mq, mk, p);
    } else {
        auto kern = dbg ? loglik_kernel<false, true> : loglik_kernel<false, false>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(loglik_kernel<bf16>)");
        kern<<<grid, kThreads, smem, stream>>>(mq, mk, p);
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "loglik_kernel launch");
    return 0;
}

}  // namespace isp
