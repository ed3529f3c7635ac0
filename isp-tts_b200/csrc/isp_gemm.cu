// Batched, ragged GEMM for sm_100a: C[b] = act(alpha * A[b] . B[b]) with TMA-fed tcgen05.mma and the accumulator in TMEM.
//
// One kernel behind every dense contraction next to the Aligner's hot path:
//   f-1  scores S = Q.K^T, dQ = dS.K, dK = dS^T.Q      (tts/models/acoustic/modules/alignment.py:189 differentiated)
//   f-3  LengthRegulator on the soft route, out = attn_soft @ x and its two gradients
//        (tts/models/acoustic/modules/temporal_adaptor.py:417-419)
//   f-2  the Conv1d layers of the projection stacks as implicit GEMMs over `taps` shifted row windows of a
//        channels-last activation, GELU / ReLU and the masked-instance-norm column sums in the epilogue
//        (alignment.py:40-83,118-154; tts/modules/normalization.py:160-208)
//
// Layout of one CTA (192 threads, one 128 x BN tile of one batch entry):
//   warp 0      producer: one lane issues the TMA loads of a stage (A: 128 rows x 128 B, B: BN rows x 128 B, both with
//               the 128 B swizzle) into a ring of `stages` buffers guarded by full / empty mbarriers.
//   warp 1      owns TMEM (BN fp32 columns x 128 lanes); one lane issues 4 tcgen05.mma per stage (kind::f16 for bf16
//               operands, kind::tf32 for fp32 operands) and releases the stage with tcgen05.commit.
//   warps 2-9   epilogue, two warps per TMEM lane quadrant taking alternate column chunks: tcgen05.ld of 32 accumulator
//               columns per step, alpha / activation / ragged masks in registers, the 32 x 128 B chunk transposed through
//               shared memory (swizzled, conflict-free) so that every global store instruction writes four whole 128 B rows.
//               (Measured and dropped: one TMA store per chunk -- the async-proxy fence and the single issuing lane made a
//               chunk cost ~1 us of latency per warp; and a persistent CTA per SM with two TMEM accumulators -- four
//               epilogue warps per SM cannot keep up with the stores: scores 87 -> 158 us, stacks 0.93 -> 1.28 ms.)
// Operands may be K-major (the contraction index contiguous) or MN-major (the row / column index contiguous): the
// second is a different TMA box and shared-memory descriptor (leading byte offset between 128 B-wide chunks), so
// transposed uses of a tensor (dS^T, x as (T2, C)) need no copy.  Ragged batches: tiles past m_len[b] / n_len[b] are
// filled with zeros without touching TMEM, rows and columns past the lengths inside a tile are masked, and the
// contraction stops at k_len[b].
// With BN <= 128 two CTAs share an SM (<= 256 TMEM columns and <= 96 KB each), so one tile's epilogue runs under the
// other's loads and MMAs; BN = 256 uses four stages and one CTA per SM (the compute-bound convolution).

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>

#include "common.cuh"
#include "isp_internal.h"
#include "isp_tc05.cuh"

namespace isp {

namespace {

constexpr int kGM = 128;                 // tile rows = UMMA M
constexpr int kEpiWarps = 8;              // two per TMEM lane quadrant, alternating column chunks
constexpr int kGThreads = 32 * (2 + kEpiWarps);
constexpr int kStageA = kGM * 128;       // bytes of A per stage
constexpr int kMaxStages = 6;

struct GemmParams {
    const int64_t* m_len;
    const int64_t* n_len;
    const int64_t* k_len;
    unsigned char* c;
    float* col_stats;
    long long* trace;                    // debug: 8 SM-clock stamps per CTA (see isp_gemm_desc.trace), or nullptr
    long long ldc, c_batch;              // elements
    int batch, M, N, K;
    int BN, tmem_cols, stages, kb, kblocks;   // BN: tile width, a multiple of 16 (64 / 32 for an MN-major B), <= 256
    int taps, tap_shift;
    int a_mn, b_mn, a_shared, b_mode;    // b_mode: third TMA coordinate of B = 0: batch, 1: zero, 2: tap
    int c_bf16, c_f16, act, fill_padding;   // c_bf16: C has 2-byte elements (bf16, or fp16 when c_f16)
    int parts;                           // row partitions of the column statistics: m_tiles * 4
    float alpha;
    uint32_t idesc;
};

// GELU with the exact (erf) form of nn.GELU(), erf by Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, i.e. below fp32 rounding of
// the product): one MUFU.RCP, one MUFU.EX2 and a degree-5 Horner chain instead of erff's branches -- the epilogue of a 128 x 256
// tile evaluates 256 of these per thread with a single warp per scheduler.
ISP_DEVINL float gelu_erf(float x) {
    const float z = fabsf(x) * 0.70710678118654752f;
    float t, e;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
    float poly = fmaf(1.061405429f, t, -1.453152027f);
    poly = fmaf(poly, t, 1.421413741f);
    poly = fmaf(poly, t, -0.284496736f);
    poly = fmaf(poly, t, 0.254829592f);
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-z * z * 1.4426950408889634f));
    const float erf_abs = fmaf(-poly * t, e, 1.0f);            // erf(|x| / sqrt 2)
    return 0.5f * x * (1.0f + copysignf(erf_abs, x));
}

ISP_DEVINL uint32_t pack_bf16(float a, float b) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}
ISP_DEVINL uint32_t pack_f16(float a, float b) {
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}
ISP_DEVINL float f16_lo(uint32_t w) { float f; asm("{.reg .f16 lo, hi; mov.b32 {lo, hi}, %1; cvt.f32.f16 %0, lo;}" : "=f"(f) : "r"(w)); return f; }
ISP_DEVINL float f16_hi(uint32_t w) { float f; asm("{.reg .f16 lo, hi; mov.b32 {lo, hi}, %1; cvt.f32.f16 %0, hi;}" : "=f"(f) : "r"(w)); return f; }

// ACT: 0 none, 1 ReLU, 2 GELU.  CT: element type of C, 0 fp32, 1 bf16, 2 fp16.  Compile-time so that the epilogue's inner loops
// carry no per-element decisions (with run-time switches a 32 x 32 chunk cost ~800 warp instructions and the sixteen epilogue
// warps of an SM were issue-bound: 16 us per tile pair on the f-1 shapes).
template <bool TF32, int ACT, int CT>
__global__ void __launch_bounds__(kGThreads)
gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const GemmParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * p.BN, m0 = blockIdx.y * kGM, b = blockIdx.z;

    unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t stage_bytes = kStageA + uint32_t(p.BN) * 128u;
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + size_t(p.stages) * stage_bytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + kMaxStages;
    uint64_t* acc_full = bars + 2 * kMaxStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 1);

    auto clampi = [](long long v, int hi) { return int(v < 0 ? 0 : (v > hi ? hi : v)); };
    const int m_valid = p.m_len ? clampi(p.m_len[b], p.M) : p.M;
    const int n_valid = p.n_len ? clampi(p.n_len[b], p.N) : p.N;
    const int k_valid = p.k_len ? clampi(p.k_len[b], p.K) : p.K;
    const int kblocks = (k_valid + p.kb - 1) / p.kb;
    const int kiters = p.taps * kblocks;

    if (m0 >= m_valid || n0 >= n_valid || kiters == 0) {
        // nothing of this tile is inside the batch entry's own extent: zeros, straight from registers
        if (!p.fill_padding) return;
        const int esz = p.c_bf16 ? 2 : 4;
        const int rows = min(kGM, p.M - m0);
        const int row_bytes = min(p.BN, p.N - n0) * esz;
        const int vecs = row_bytes >> 4;
        unsigned char* cb = p.c + (size_t(b) * p.c_batch + size_t(m0) * p.ldc + n0) * esz;
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
        for (int idx = threadIdx.x; idx < rows * vecs; idx += kGThreads) {
            const int r = idx / vecs, v = idx - r * vecs;
            st_cs_v4(cb + size_t(r) * p.ldc * esz + size_t(v) * 16, z);
        }
        const int tail = row_bytes - (vecs << 4);          // N * esz not a multiple of 16 B
        if (tail) {
            for (int idx = threadIdx.x; idx < rows * (tail >> 1); idx += kGThreads) {
                const int r = idx / (tail >> 1), v = idx - r * (tail >> 1);
                *reinterpret_cast<uint16_t*>(cb + size_t(r) * p.ldc * esz + (vecs << 4) + v * 2) = 0;
            }
        }
        return;
    }

    const int cta_lin = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    long long* tr = p.trace ? p.trace + size_t(cta_lin) * 8 : nullptr;
    if (tr && threadIdx.x == 0) tr[0] = clock64();
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(acc_full, 1);
        fence_mbar_init();
        tc::prefetch_tmap(&tmap_a);
        tc::prefetch_tmap(&tmap_b);
    }
    if (warp == 1) tc::tmem_alloc(tmem_slot, uint32_t(p.tmem_cols));
    tc::fence_before();
    __syncthreads();
    tc::fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (tr && threadIdx.x == 0) tr[1] = clock64();             // barriers, tensor-map prefetch, TMEM allocation, CTA sync

    const int me = TF32 ? 32 : 64;                 // M/N indices per 128 B row of an MN-major operand
    const uint32_t lbo = uint32_t(p.kb) * 128u;    // bytes between two such chunks (one TMA box each)

    if (warp == 0) {
        if (lane == 0) {
            for (int it = 0; it < kiters; ++it) {
                const int s = it % p.stages;
                const uint32_t ph = uint32_t(it / p.stages) & 1u;
                mbar_wait(&empty[s], ph ^ 1u);
                const int tap = it / kblocks, kbi = it - tap * kblocks;
                unsigned char* sa = base + size_t(s) * stage_bytes;
                unsigned char* sb = sa + kStageA;
                mbar_arrive_expect_tx(&full[s], stage_bytes);
                const int ab = p.a_shared ? 0 : b;
                if (!p.a_mn) {
                    tc::tma_load_3d(sa, &tmap_a, kbi * p.kb, m0 + tap + p.tap_shift, ab, &full[s]);
                } else {
                    for (int c = 0; c < kGM / me; ++c) tc::tma_load_3d(sa + c * lbo, &tmap_a, m0 + c * me, kbi * p.kb, ab, &full[s]);
                }
                const int third = p.b_mode == 0 ? b : (p.b_mode == 1 ? 0 : tap);
                if (!p.b_mn) {
                    tc::tma_load_3d(sb, &tmap_b, kbi * p.kb, n0, third, &full[s]);
                } else {
                    for (int c = 0; c < p.BN / me; ++c) tc::tma_load_3d(sb + c * lbo, &tmap_b, n0 + c * me, kbi * p.kb, third, &full[s]);
                }
                if (tr && it == 0) tr[2] = clock64();            // first stage issued
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t kstep_mn = (TF32 ? 8u : 16u) * 128u;      // bytes per MMA k-step of an MN-major operand
            for (int it = 0; it < kiters; ++it) {
                const int s = it % p.stages;
                const uint32_t ph = uint32_t(it / p.stages) & 1u;
                mbar_wait(&full[s], ph);
                tc::fence_after();
                if (tr && it == 0) tr[3] = clock64();            // first stage landed
                const uint32_t sa = smem_u32(base + size_t(s) * stage_bytes);
                const uint32_t sb = sa + kStageA;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    // MN-major: 16-bit operands use the plain 128 B swizzle (8 k-rows per atom); TF32 the 32 B-atom form (4 k-rows)
                    const uint64_t ad = p.a_mn ? tc::smem_desc_sw128(sa + k * kstep_mn, lbo, TF32 ? 512 : 1024, TF32 ? 1 : 2)
                                               : tc::smem_desc_sw128(sa + k * 32, 16, 1024);
                    const uint64_t bd = p.b_mn ? tc::smem_desc_sw128(sb + k * kstep_mn, lbo, TF32 ? 512 : 1024, TF32 ? 1 : 2)
                                               : tc::smem_desc_sw128(sb + k * 32, 16, 1024);
                    tc::umma<TF32>(tmem_base, ad, bd, p.idesc, (it > 0 || k > 0) ? 1u : 0u);
                }
                tc::umma_commit(&empty[s]);
            }
            tc::umma_commit(acc_full);
            if (tr) tr[4] = clock64();                           // every MMA issued
        }
    } else {
        // ================================ epilogue ================================
        constexpr int ce = CT == 0 ? 32 : 64;                    // output columns per 128 B staged row
        constexpr int esz = CT == 0 ? 4 : 2;
        const int quad = warp & 3;                              // the TMEM lane quadrant this warp may read
        const int half = (warp - 2) >> 2;                        // 0: chunks 0, 2, 4, ...   1: chunks 1, 3, 5, ...
        const int row = m0 + quad * 32 + lane;
        const bool row_ok = row < m_valid;
        mbar_wait(acc_full, 0);                                  // all MMAs done: the stage ring is free, staging aliases it
        tc::fence_after();
        if (tr && threadIdx.x == 64) tr[5] = clock64();          // accumulator complete
        unsigned char* buf = base + size_t(warp - 2) * 4096;     // 32 rows x 128 B, 16 B chunks XOR-swizzled by (row & 7)
        const uint32_t tlane = tmem_base + (uint32_t(quad * 32) << 16);
        const int ncols = min(p.BN, p.N - n0);
        const int sw = lane & 7;
        // store side: lane -> (row rb + 4 * it, 16 B chunk q_st) of the staged block; four whole rows per instruction
        const int rb = lane >> 3, q_st = lane & 7;
        const int rows_here = min(32, p.M - (m0 + quad * 32));   // rows of this warp's slab inside the tensor
        const size_t row_pitch = size_t(p.ldc) * esz;
        unsigned char* cst = p.c + (size_t(b) * p.c_batch + size_t(m0 + quad * 32 + rb) * p.ldc + n0) * esz + (q_st << 4);
        const unsigned char* lds0 = buf + rb * 128;              // + it * 512; chunk (q_st ^ (r & 7)) with r & 7 = (rb + 4 it) & 7
        unsigned char* myrow = buf + lane * 128;
        const int n_b = ncols * esz;                             // bytes of a tile row inside the tensor
        for (int c0 = half * ce; c0 < ncols; c0 += 2 * ce) {
            const bool cols_in = n0 + c0 + ce <= n_valid;        // no column of this chunk is past the batch entry's n_len
#pragma unroll
            for (int h = 0; h < (CT == 0 ? 1 : 2); ++h) {
                float v[32];
                tc::tmem_ld32(tlane + uint32_t(c0 + 32 * h), v);
#pragma unroll
                for (int k = 0; k < 32; ++k) {
                    float x = v[k] * p.alpha;
                    if (ACT == 1) x = fmaxf(x, 0.0f);
                    if (ACT == 2) x = gelu_erf(x);
                    v[k] = x;
                }
                if (!(row_ok && cols_in)) {                      // ragged edge: the rare path carries the masks
                    const int jbase = n0 + c0 + 32 * h;
#pragma unroll
                    for (int k = 0; k < 32; ++k) v[k] = (row_ok && jbase + k < n_valid) ? v[k] : 0.0f;
                }
                if (CT == 0) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const uint4 w = make_uint4(__float_as_uint(v[4 * q]), __float_as_uint(v[4 * q + 1]),
                                                   __float_as_uint(v[4 * q + 2]), __float_as_uint(v[4 * q + 3]));
                        *reinterpret_cast<uint4*>(myrow + ((q ^ sw) << 4)) = w;
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint4 w = CT == 2
                            ? make_uint4(pack_f16(v[8 * q], v[8 * q + 1]), pack_f16(v[8 * q + 2], v[8 * q + 3]),
                                         pack_f16(v[8 * q + 4], v[8 * q + 5]), pack_f16(v[8 * q + 6], v[8 * q + 7]))
                            : make_uint4(pack_bf16(v[8 * q], v[8 * q + 1]), pack_bf16(v[8 * q + 2], v[8 * q + 3]),
                                         pack_bf16(v[8 * q + 4], v[8 * q + 5]), pack_bf16(v[8 * q + 6], v[8 * q + 7]));
                        *reinterpret_cast<uint4*>(myrow + (((4 * h + q) ^ sw) << 4)) = w;
                    }
                }
            }
            __syncwarp();
            if (p.col_stats) {
                // column sums over this warp's 32 rows (masked rows hold zeros): lane L takes the 32-bit word L of every
                // staged row -- one fp32 column, or two 16-bit columns (the values the next layer will read)
                float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
                const int chunk = lane >> 2, within = (lane & 3) << 2;
#pragma unroll 8
                for (int r = 0; r < 32; ++r) {
                    const uint32_t w = *reinterpret_cast<const uint32_t*>(buf + r * 128 + ((chunk ^ (r & 7)) << 4) + within);
                    if (CT != 0) {
                        const float a = CT == 2 ? f16_lo(w) : __uint_as_float(w << 16);
                        const float c = CT == 2 ? f16_hi(w) : __uint_as_float(w & 0xffff0000u);
                        s0 += a; q0 += a * a; s1 += c; q1 += c * c;
                    } else {
                        const float a = __uint_as_float(w);
                        s0 += a; q0 += a * a;
                    }
                }
                float* st = p.col_stats + (size_t(b) * p.parts + size_t(blockIdx.y) * 4 + quad) * size_t(p.N) * 2;
                if (CT != 0) {
                    const int col = n0 + c0 + 2 * lane;
                    if (col < p.N) *reinterpret_cast<float2*>(st + size_t(col) * 2) = make_float2(s0, q0);
                    if (col + 1 < p.N) *reinterpret_cast<float2*>(st + size_t(col + 1) * 2) = make_float2(s1, q1);
                } else {
                    const int col = n0 + c0 + lane;
                    if (col < p.N) *reinterpret_cast<float2*>(st + size_t(col) * 2) = make_float2(s0, q0);
                }
            }
            // coalesced stores: 8 lanes cover one staged row (128 B), a warp instruction four rows
            unsigned char* dst0 = cst + size_t(c0) * esz;
            const int col_b = c0 * esz + (q_st << 4);                          // byte offset of this lane's 16 B inside the tile row
            if (rows_here == 32 && c0 * esz + 128 <= n_b) {                    // the whole chunk is inside the tensor
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const uint4 w = *reinterpret_cast<const uint4*>(lds0 + it * 512 + ((q_st ^ ((rb + 4 * it) & 7)) << 4));
                    *reinterpret_cast<uint4*>(dst0 + size_t(4 * it) * row_pitch) = w;
                }
            } else {
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int r = rb + 4 * it;
                    if (r < rows_here && col_b < n_b) {
                        const uint4 w = *reinterpret_cast<const uint4*>(lds0 + it * 512 + ((q_st ^ (r & 7)) << 4));
                        unsigned char* dst = dst0 + size_t(4 * it) * row_pitch;
                        if (col_b + 16 <= n_b) {
                            *reinterpret_cast<uint4*>(dst) = w;
                        } else {                                               // N * esz is not a multiple of 16 B: the last few elements
                            const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
                            for (int e2 = 0; e2 < (n_b - col_b) >> 1; ++e2)
                                reinterpret_cast<uint16_t*>(dst)[e2] = uint16_t(ws[e2 >> 1] >> ((e2 & 1) * 16));
                        }
                    }
                }
            }
            __syncwarp();
        }
        if (tr && threadIdx.x == 64) tr[6] = clock64();          // first epilogue warp done
    }

    tc::fence_before();
    __syncthreads();
    if (warp == 1) {
        tc::fence_after();
        tc::tmem_dealloc(tmem_base, uint32_t(p.tmem_cols));
    }
    if (tr && threadIdx.x == 32) tr[7] = clock64();
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled encoder() {
    static PFN_encodeTiled fn = nullptr;
    if (fn) return fn;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess || !ptr) return nullptr;
    fn = reinterpret_cast<PFN_encodeTiled>(ptr);
    return fn;
}

// (inner, rows, third) tensor with a (box_inner, box_rows, 1) box and the 128 B swizzle
int make_map3(CUtensorMap* map, const void* ptr, int dtype, long long inner, long long rows, long long third,
              long long row_stride_elems, long long third_stride_elems, int box_inner, int box_rows, const char* what,
              bool atom32 = false) {
    PFN_encodeTiled enc = encoder();
    if (!enc) { set_error("cuTensorMapEncodeTiled not available from the driver"); return ISP_ERR_DEVICE; }
    const int elem = dtype == ISP_DTYPE_F32 ? 4 : 2;
    if (third < 1) third = 1;
    if (third_stride_elems <= 0) third_stride_elems = row_stride_elems * rows;     // a single slice: any legal stride
    cuuint64_t dims[3] = {cuuint64_t(inner), cuuint64_t(rows), cuuint64_t(third)};
    cuuint64_t strides[2] = {cuuint64_t(row_stride_elems) * elem, cuuint64_t(third_stride_elems) * elem};
    cuuint32_t box[3] = {cuuint32_t(box_inner), cuuint32_t(box_rows), 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, dtype == ISP_DTYPE_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : (dtype == ISP_DTYPE_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32), 3,
                     const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("isp_gemm_batched: cuTensorMapEncodeTiled(%s) failed with CUresult %d (inner=%lld rows=%lld third=%lld ld=%lld)",
                  what, int(r), inner, rows, third, row_stride_elems);
        return ISP_ERR_INVALID;
    }
    return 0;
}

template <bool TF32, int ACT, int CT>
cudaError_t launch_one(dim3 grid, size_t smem, cudaStream_t stream, const CUtensorMap& ma, const CUtensorMap& mb, const GemmParams& p) {
    cudaError_t e = cudaFuncSetAttribute(gemm_kernel<TF32, ACT, CT>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    gemm_kernel<TF32, ACT, CT><<<grid, kGThreads, smem, stream>>>(ma, mb, p);
    return cudaSuccess;
}

template <bool TF32, int ACT>
cudaError_t launch_ct(int ct, dim3 grid, size_t smem, cudaStream_t stream, const CUtensorMap& ma, const CUtensorMap& mb, const GemmParams& p) {
    if (ct == 0) return launch_one<TF32, ACT, 0>(grid, smem, stream, ma, mb, p);
    if (ct == 1) return launch_one<TF32, ACT, 1>(grid, smem, stream, ma, mb, p);
    return launch_one<TF32, ACT, 2>(grid, smem, stream, ma, mb, p);
}

cudaError_t launch_gemm(bool tf32, int act, int ct, dim3 grid, size_t smem, cudaStream_t stream, const CUtensorMap& ma, const CUtensorMap& mb,
                        const GemmParams& p) {
    if (tf32) {
        if (act == 0) return launch_ct<true, 0>(ct, grid, smem, stream, ma, mb, p);
        if (act == 1) return launch_ct<true, 1>(ct, grid, smem, stream, ma, mb, p);
        return launch_ct<true, 2>(ct, grid, smem, stream, ma, mb, p);
    }
    if (act == 0) return launch_ct<false, 0>(ct, grid, smem, stream, ma, mb, p);
    if (act == 1) return launch_ct<false, 1>(ct, grid, smem, stream, ma, mb, p);
    return launch_ct<false, 2>(ct, grid, smem, stream, ma, mb, p);
}

}  // namespace

int gemm_batched(const isp_gemm_desc* d, cudaStream_t stream) {
    if (!d || !d->a || !d->b || !d->c) { set_error("isp_gemm_batched: null pointer"); return ISP_ERR_INVALID; }
    if (d->batch <= 0 || d->M <= 0 || d->N <= 0 || d->K <= 0) { set_error("isp_gemm_batched: sizes must be positive"); return ISP_ERR_INVALID; }
    if (d->batch > 65535) { set_error("isp_gemm_batched: batch=%d > 65535", d->batch); return ISP_ERR_UNSUPPORTED; }
    if (d->dtype_ab < ISP_DTYPE_F32 || d->dtype_ab > ISP_DTYPE_F16 || d->dtype_c < ISP_DTYPE_F32 || d->dtype_c > ISP_DTYPE_F16) {
        set_error("isp_gemm_batched: dtypes must be ISP_DTYPE_F32, ISP_DTYPE_BF16 or ISP_DTYPE_F16"); return ISP_ERR_INVALID;
    }
    const int taps = d->taps > 0 ? d->taps : 1;
    if (taps > 1 && d->a_mn_major) { set_error("isp_gemm_batched: taps > 1 needs a K-major A"); return ISP_ERR_UNSUPPORTED; }
    const int elem = d->dtype_ab == ISP_DTYPE_F32 ? 4 : 2, esz_c = d->dtype_c == ISP_DTYPE_F32 ? 4 : 2;
    auto mis = [](const void* ptr, long long s1, long long s2, int e) {
        return (reinterpret_cast<uintptr_t>(ptr) & 15) || ((s1 * e) & 15) || ((s2 * e) & 15);
    };
    if (mis(d->a, d->lda, d->a_batch, elem) || mis(d->b, d->ldb, d->b_batch, elem) || mis(d->b, d->b_tap_stride, 0, elem) ||
        mis(d->c, d->ldc, d->c_batch, esz_c)) {
        set_error("isp_gemm_batched: every base pointer and every leading / batch stride must be a multiple of 16 B"); return ISP_ERR_INVALID;
    }
    if (d->act < 0 || d->act > 2) { set_error("isp_gemm_batched: act must be 0 (none), 1 (relu) or 2 (gelu)"); return ISP_ERR_INVALID; }

    GemmParams p;
    p.m_len = d->m_len; p.n_len = d->n_len; p.k_len = d->k_len;
    p.c = static_cast<unsigned char*>(d->c);
    p.col_stats = d->col_stats;
    p.trace = static_cast<long long*>(d->trace);
    p.ldc = d->ldc; p.c_batch = d->c_batch;
    p.batch = d->batch; p.M = d->M; p.N = d->N; p.K = d->K;
    // Tile width: as few column tiles as 256 TMEM columns allow, each as narrow as N permits (UMMA N is any multiple of 16),
    // so that N = 160, 200 or 384 waste nothing; an MN-major B arrives in 128 B-wide boxes, hence multiples of 64 / 32.
    const int n_unit = d->b_mn_major ? 128 / elem : 16;
    int bn = d->bn;
    if (bn == 0) {
        const int nt = (d->N + 255) / 256;
        bn = ((d->N + nt - 1) / nt + n_unit - 1) / n_unit * n_unit;
    }
    if (bn < 16 || bn > 256 || bn % n_unit) { set_error("isp_gemm_batched: bn must be 0 or a multiple of %d up to 256", n_unit); return ISP_ERR_INVALID; }
    p.BN = bn;
    p.tmem_cols = 32;
    while (p.tmem_cols < bn) p.tmem_cols <<= 1;
    // Ring depth: two stages.  What pays on every shape measured (tools/conv_bench.py, tools/gemm_bench.py) is more CTAs per SM,
    // not a deeper ring per CTA: each CTA runs load -> MMA -> epilogue in sequence, so the overlap of one tile's epilogue with
    // another's loads and MMAs comes from co-resident CTAs (key convolution, 128 x 256 tiles, K = 1920: 178 us with four stages and
    // one CTA per SM, 142 us with two stages and two CTAs).
    p.stages = 2;
    if (d->stages >= 2 && d->stages <= kMaxStages) p.stages = d->stages;      // tuning override
    p.kb = 128 / elem;
    p.kblocks = (d->K + p.kb - 1) / p.kb;
    p.taps = taps; p.tap_shift = d->tap_shift;
    p.a_mn = d->a_mn_major ? 1 : 0; p.b_mn = d->b_mn_major ? 1 : 0;
    p.a_shared = d->a_batch == 0 ? 1 : 0;
    p.b_mode = taps > 1 ? 2 : (d->b_batch == 0 ? 1 : 0);
    if (taps > 1 && d->b_batch != 0) { set_error("isp_gemm_batched: with taps > 1 B is the shared (taps, N, K) filter: b_batch must be 0"); return ISP_ERR_UNSUPPORTED; }
    p.c_bf16 = d->dtype_c == ISP_DTYPE_F32 ? 0 : 1;
    p.c_f16 = d->dtype_c == ISP_DTYPE_F16 ? 1 : 0;
    p.act = d->act;
    p.fill_padding = d->skip_padding ? 0 : 1;
    p.alpha = d->alpha;
    const int m_tiles = (d->M + kGM - 1) / kGM;
    p.parts = m_tiles * 4;
    const uint32_t fmt = d->dtype_ab == ISP_DTYPE_BF16 ? 1u : (d->dtype_ab == ISP_DTYPE_F16 ? 0u : 2u);   // UMMA F16 = 0, BF16 = 1, TF32 = 2
    p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (uint32_t(p.a_mn) << 15) | (uint32_t(p.b_mn) << 16) |
              (uint32_t(bn >> 3) << 17) | (uint32_t(kGM >> 4) << 24);

    const int me = 128 / elem;          // 64 bf16 / 32 fp32 indices per 128 B
    CUtensorMap ma, mb;
    int rc;
    if (!p.a_mn) rc = make_map3(&ma, d->a, d->dtype_ab, d->K, d->M, p.a_shared ? 1 : d->batch, d->lda, d->a_batch, p.kb, kGM, "A");
    else         rc = make_map3(&ma, d->a, d->dtype_ab, d->M, d->K, p.a_shared ? 1 : d->batch, d->lda, d->a_batch, me, p.kb, "A", elem == 4);
    if (rc) return rc;
    const long long b_third = p.b_mode == 0 ? d->batch : (p.b_mode == 1 ? 1 : taps);
    const long long b_third_stride = p.b_mode == 0 ? d->b_batch : (p.b_mode == 1 ? 0 : d->b_tap_stride);
    if (!p.b_mn) rc = make_map3(&mb, d->b, d->dtype_ab, d->K, d->N, b_third, d->ldb, b_third_stride, p.kb, bn, "B");
    else         rc = make_map3(&mb, d->b, d->dtype_ab, d->N, d->K, b_third, d->ldb, b_third_stride, me, p.kb, "B", elem == 4);
    if (rc) return rc;

    const size_t smem = size_t(p.stages) * (kStageA + size_t(bn) * 128) + sizeof(uint64_t) * (2 * kMaxStages + 2) + 1024;
    const dim3 grid((d->N + bn - 1) / bn, m_tiles, d->batch);
    cudaError_t e = launch_gemm(d->dtype_ab == ISP_DTYPE_F32, d->act, d->dtype_c == ISP_DTYPE_F32 ? 0 : (d->dtype_c == ISP_DTYPE_BF16 ? 1 : 2),
                                grid, smem, stream, ma, mb, p);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(gemm_kernel)");
    e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "gemm_kernel launch");
    return 0;
}

}  // namespace isp
