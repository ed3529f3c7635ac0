// Monotonic Alignment Search for sm_100a: forward max-plus DP, packed backpointer
// bits, warp-parallel backtrack, dense int16 path and int64 durations -- one launch.
//
// Reference semantics (paths relative to the reference root):
//   tts/modules/aligner/mas.py:8-26      mas_width1 (DP, tie rule, backtrack)
//   tts/modules/aligner/cuda_mas.py:11-46 cuda_b_mas (the GPU kernel this replaces)
//   tts/models/acoustic/modules/alignment.py:275  durations = attn_hard.sum(dim=1)
//
// Shape of the computation.  Q[i][j] = x[i][j] + max(Q[i-1][j-1], Q[i-1][j]) depends on
// row i-1 only: the text axis is parallel, the frame axis is a serial chain of T1 steps, and
// the time of a batch is the time of its longest chain unless HBM saturates first.  The
// kernel is therefore built around the latency of ONE row step, and everything that is not
// arithmetic of that step is moved off the warp that runs the chain (measured costs behind
// these choices: tools/ubench/chain.cu, tools/ubench/sync.cu):
//
//   * One "slot" = one utterance in flight.  A CTA holds 1 or 2 slots; a slot is NS strip
//     warps (128 text columns each, 4 per lane), NS loader warps and one filler warp.  Strip
//     warps get the lowest warp ids, so the strips of co-resident utterances sit on different
//     SM sub-partitions.
//   * Inside a strip warp the wavefront is skewed across LANES: lane l owns 4 consecutive
//     columns and, at step t, works on row t - l.  The only cross-lane value a row needs
//     (Q[i-1][first column - 1]) was produced by the left neighbour two steps earlier, so
//     its shuffle is issued one step ahead and its latency never sits on the chain.  What
//     is left on the chain per step is FMNMX -> FADD; per cell the step costs FSET + FMNMX
//     (ALU pipe) and FFMA + FADD (FMA pipe): the backpointer bit is accumulated as a float
//     (FSET gives 1.0/0.0; an FFMA tree packs a row's 4 bits, one more FFMA files them under
//     the step's nibble), so a lane emits one 32-bit word of bits per 8 steps.
//   * Steps come in chunks of 16.  Everything that synchronises happens once per chunk: an
//     mbarrier wait costs the warp >= 80 cycles even when the phase is long complete, an
//     acquire load ~70, so neither may appear per step.
//   * The logits reach a strip through a ring of 16-row stages in shared memory, filled by the
//     strip's own loader warp (one lane) with tiled TMA copies (cp.async.bulk.tensor.3d, box =
//     16 rows x the utterance's columns of the strip rounded up to 8; one op costs the issuing
//     lane ~150 cycles whatever its size, hence >= 4 KB boxes and one loader per strip), "full"
//     and "empty" mbarriers.  Because of the lane skew a stage stays live for 4 chunks; the
//     rest of the ring is prefetch depth.  The row pitch is the box width, a multiple of 8
//     floats, which makes the skewed 16 B reads bank-conflict-free.  Rows past T1_b and columns
//     past T2_b (rounded up) are never requested.  Bases or strides that are not 16 B aligned
//     take 4 B cp.async copies issued by the whole loader warp instead.
//   * Strips of an utterance wider than 128 tokens form a second, coarser wavefront: strip s
//     runs 3 chunks behind strip s-1 and takes its boundary values, 16 per chunk, from a small
//     shared-memory ring; progress counters in both directions are plain shared words, read a
//     chunk ahead of their use.
//   * Backpointers cost one bit per cell: word (c, L) holds, for global lane L = column / 4,
//     the nibbles of steps 8c .. 8c+7 (rows 8c - l + k).  In shared memory when the slot's
//     bits fit next to a deep enough ring, else in the caller's workspace (coalesced 128 B
//     stores, L2-resident).
//   * The filler warp zero-fills the utterance's dense int16 block: bulk shared->global
//     copies of a zero page (one lane), paced over the duration of the DP.
//   * backtrack (strip warp 0): 32 rows per block.  Every lane assembles the 32-column window
//     its row can touch (the path moves at most one column per row) from the skewed words;
//     the windows are broadcast through shared memory and the dependent chain runs on a
//     one-hot position register, two ALU levels per row:  R' = (R & ~A) | ((R >> 1) & (A >> 1)).
//     Then all 32 lanes write their row's 1 and the durations of the tokens that start in
//     the block (ballot + clz), so nothing is re-read to build durations.
//
// Bit-exactness: each cell does exactly the reference's one fp32 add on top of an exact
// max; the comparison is the reference's `>=` (ties and -inf >= -inf take the diagonal).

#include <algorithm>
#include <cuda.h>
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include <type_traits>

#include "common.cuh"
#include "isp_internal.h"
#include "isp_mas_ptx.cuh"

namespace isp {

constexpr int kR = 16;                // steps per chunk = rows per ring stage = rows per TMA box
constexpr int kC = 4;                 // columns per lane
constexpr int kW = 32 * kC;           // columns per strip warp
constexpr int kMaxStages = 16;
constexpr int kMinStages = 5;         // 4 live + 1 in flight
constexpr int kMaxStrips = 5;         // ISP_MAS_MAX_T2 / kW
constexpr int kMaxSlots = 3;
constexpr int kMaxThreads = 480;      // slots * (2 * strips + 1) warps <= 15 (3 slots x 2 strips; 1 slot x 5 strips = 11)
constexpr int kBnd = 128;             // rows in a strip-boundary ring
constexpr int kBndChunks = kBnd / kR; // = 8, power of two
constexpr int kZeroPage = 4096;       // bytes of zeros behind the bulk zero-fill
constexpr int kSlotHdr = 2560;        // barriers (2 x 1 KB, 512 B used), progress counters (256 B), backtrack windows (256 B)
constexpr int kNumBox = kW / 8;       // TMA box widths: 8, 16, ..., 128 columns

struct MasParams {
    const float* logp;
    int64_t sB, sT1;
    const int64_t* text_len;
    const int64_t* mel_len;
    int B, T1max, T2max;
    int16_t* hard;
    int64_t* dur;
    uint32_t* bits_ws;     // global backpointer bits (BITS_SMEM == false)
    int64_t bits_stride;   // words per utterance in bits_ws
    int16_t* path_ws;      // (B, T1max) int16: the path's column per row, until the zero-fill has landed (the caller's `path` if given)
    int path_fill;         // 1: path_ws is the caller's output: frames past mel_len get -1
    int* status;           // count of utterances with out-of-contract lengths
    long long* probe;      // clock64 stamps of utterance 0 (tools/mas_probe.py)
    int ns;                // strip warps per utterance
    int slots;             // utterances per CTA
    int nstg;              // ring stages
    int wlast;             // ring row floats allocated for the last strip (multiple of 8)
    int slot_bytes;        // shared memory per slot
    int tma;               // 1: the tensor maps are valid -> tiled TMA copies, else 4 B async copies
    int dbg;               // debug/profiling switches (mas.dbg)
};

struct MasMaps { CUtensorMap m[kNumBox]; };

ISP_DEVINL void lds_row(float (&x)[kC], uint32_t saddr) {
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x[0]), "=f"(x[1]), "=f"(x[2]), "=f"(x[3]) : "r"(saddr));
}

// ---- one DP row for one lane --------------------------------------------------------
// q[] holds row i-1 on entry and row i on exit.  Returns the lane's 4 backpointer bits as a
// float 0..15 (bit c set <=> the predecessor of column base+c is the diagonal).
ISP_DEVINL float dp_row(float (&q)[kC], const float (&x)[kC], float left) {
    float s[kC];
#pragma unroll
    for (int c = kC - 1; c >= 1; --c) {
        s[c] = set_ge(q[c - 1], q[c]);                // mas.py:17 -- ties take j-1
        q[c] = x[c] + fmaxf(q[c - 1], q[c]);          // mas.py:14 -- one fp32 add per cell
    }
    s[0] = set_ge(left, q[0]);                        // left == NaN at global column 0: false, keeps q[0]
    q[0] = x[0] + fmaxf(left, q[0]);
    return fmaf(fmaf(s[3], 2.0f, s[2]), 4.0f, fmaf(s[1], 2.0f, s[0]));
}

template <bool BITS_SMEM> ISP_DEVINL uint4 load_bits4(const uint32_t* gp, uint32_t sa) {
    if (BITS_SMEM) return lds_v4(sa);
    return __ldcg(reinterpret_cast<const uint4*>(gp));
}

template <bool BITS_SMEM, bool MULTI>
__global__ void __launch_bounds__(kMaxThreads, 1)
mas_kernel(const __grid_constant__ MasMaps maps, const MasParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];

    // warps of a CTA: [slots x ns strips][slots x ns loaders][slots fillers]
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int ns = p.ns;
    const int nstrip_total = p.slots * ns;
    const int role = warp < nstrip_total ? 0 : (warp < 2 * nstrip_total ? 1 : 2);    // strip, loader, filler
    const int wrel = warp - role * nstrip_total;
    const int slot = role == 2 ? wrel : wrel / ns;
    const int s = role == 2 ? 0 : wrel - slot * ns;
    const bool is_strip = role == 0;
    const int b = blockIdx.x * p.slots + slot;
    if (b >= p.B) return;                               // the whole slot leaves together
    const uint32_t slot_threads = 32u * (2 * ns + 1);
    const int nstg = p.nstg;

    // ---- carve the slot's shared memory -------------------------------------------------
    unsigned char* sm = smem_raw + size_t(slot) * p.slot_bytes;
    const uint32_t sm_sa = smem_u32(sm);
    const uint32_t full_sa = sm_sa;                                             // [kMaxStrips][kMaxStages] u64: stage has landed
    const uint32_t empty_sa = sm_sa + 1024;                                     // [kMaxStrips][kMaxStages] u64: stage may be refilled
    const uint32_t prog_sa = sm_sa + 2048;                                      // [kMaxStrips] chunks whose last column strip s has published
    const uint32_t cons_sa = prog_sa + 4 * kMaxStrips;                          // [kMaxStrips] chunks whose boundary values strip s has taken
    const uint32_t landed_sa = cons_sa + 4 * kMaxStrips;                        // [kMaxStrips] chunks of logits that have landed in strip s's ring
    uint32_t* winbuf = reinterpret_cast<uint32_t*>(sm + 2304);                  // [32][2] backtrack windows (A, A >> 1)
    uint32_t off = kSlotHdr;
    const uint32_t bnd_sa = sm_sa + off;                                        // [ns-1][kBnd] floats
    off += MULTI ? 4u * uint32_t(ns - 1) * kBnd : 0u;
    const uint32_t zero_sa = sm_sa + off;
    off += kZeroPage;
    const uint32_t ring_sa = sm_sa + off;                                       // strip s: ring_sa + s * nstg * kR * kW * 4
    off += uint32_t(nstg) * kR * uint32_t((ns - 1) * kW + p.wlast) * 4u;
    const int wpt = ns * 32;                                                    // words of bits per 8 steps
    const uint32_t bits_sa = sm_sa + off;
    uint32_t* bits_g = BITS_SMEM ? nullptr : p.bits_ws + size_t(b) * p.bits_stride;

    // ---- lengths (read on device; clamped for memory safety, reported via status) ------
    const long long n64 = p.mel_len[b], m64 = p.text_len[b];
    const bool bad = n64 < 1 || n64 > p.T1max || m64 < 1 || m64 > p.T2max;
    const int n = int(n64 < 1 ? 1 : (n64 > p.T1max ? p.T1max : n64));   // frames
    const int m = int(m64 < 1 ? 1 : (m64 > p.T2max ? p.T2max : m64));   // tokens
    const int ns_active = (m + kW - 1) / kW;
    const int nch = (n + 31 + kR - 1) / kR;                             // chunks of a strip's n + 31 steps
    const bool probe_w = p.probe != nullptr && b == 0 && s == 0;        // warp-uniform
    const bool probe = probe_w && is_strip && lane == 0;
    long long pc_full = 0, pc_flag = 0, pc_loop = 0, pc_nloop = 0;

    // per strip: the utterance's columns, the TMA box / ring row width and the ring
    const int mcols = max(0, min(kW, m - s * kW));
    const int wb = (mcols + 7) & ~7;                                    // floats per ring row
    const uint32_t pitchB = uint32_t(wb) * 4u;
    const uint32_t stageB = uint32_t(kR) * pitchB;
    const uint32_t ringB = uint32_t(nstg) * stageB;
    const uint32_t ring_s = ring_sa + uint32_t(s) * uint32_t(nstg) * kR * kW * 4u;
    const uint32_t full_s = full_sa + uint32_t(s) * kMaxStages * 8u;
    const uint32_t empty_s = empty_sa + uint32_t(s) * kMaxStages * 8u;

    if (role == 0) {
        if (lane == 0) {
            if (s == 0 && bad) atomicAdd(p.status, 1);
            for (int st = 0; st < nstg; ++st) {
                mbar_init(reinterpret_cast<uint64_t*>(sm) + s * kMaxStages + st, p.tma ? 1 : 32);
                mbar_init(reinterpret_cast<uint64_t*>(sm + 1024) + s * kMaxStages + st, 1);
            }
            reinterpret_cast<int*>(sm + 2048)[s] = 0;
            reinterpret_cast<int*>(sm + 2048)[kMaxStrips + s] = 0;
            reinterpret_cast<int*>(sm + 2048)[2 * kMaxStrips + s] = 0;
            if (s == 0) {
                for (int i = 0; i < 32; ++i) reinterpret_cast<int*>(sm + 2048 + 128)[i] = 0;   // converter / chain counters
                for (int i = 0; i < 2 * (2 * kMaxStrips - 1); ++i) mbar_init(reinterpret_cast<uint64_t*>(sm + 640) + i, 1);   // converter staging
            }
            fence_mbar_init();
        }
    } else if (role == 2) {
        for (int i = lane; i < kZeroPage / 16; i += 32) reinterpret_cast<uint4*>(sm + (zero_sa - sm_sa))[i] = make_uint4(0, 0, 0, 0);
        fence_proxy_async();                            // the zero page is read by bulk copies (async proxy)
    }
    asm volatile("bar.sync %0, %1;" ::"r"(slot + 1), "r"(slot_threads) : "memory");
    if (probe) p.probe[0] = clock64();

    if (role == 0) {
        if (s < ns_active) {
            // =========================== strip warp: forward DP ===========================
            const uint32_t lane_ring = ring_s + min(uint32_t(lane) * 16u, pitchB - 16u);   // lanes past the box read (and discard) its last columns
            const bool has_prev = MULTI && s > 0;
            const bool has_next = MULTI && s + 1 < ns_active;
            const bool pub = has_next && lane == 31;                // this lane publishes the strip's last column
            // boundary rings: row r of strip s's last column lives in slot (r + 31) mod kBnd of ring s, i.e. at the
            // producer's step index -- chunk-aligned for the writer; the reader's 16 rows straddle two chunks (15 | 0..14)
            const uint32_t bnd_mine = bnd_sa + uint32_t(s) * kBnd * 4u;
            const uint32_t bnd_prev = bnd_sa + uint32_t(s > 0 ? s - 1 : 0) * kBnd * 4u;
            const uint32_t landed_s = landed_sa + 4u * s;
            const float qnan = __int_as_float(0x7fffffff);
            const int gcol0 = s * kW + lane * kC;
            const bool lane0 = lane == 0;

            float q[kC], xc[kC];
#pragma unroll
            for (int c = 0; c < kC; ++c) q[c] = -CUDART_INF_F;
            float left_cur = qnan;
            float bvals[kR];                                        // left neighbours of lane 0's next 16 rows (NaN: there is none)
#pragma unroll
            for (int k = 0; k < kR; ++k) bvals[k] = qnan;

            // ring row (byte offset inside the ring) this lane reads next: row (t + 1 - lane) mod ring at step t
            uint32_t rd = (uint32_t(nstg * kR) - uint32_t(lane)) * pitchB;
            if (rd >= ringB) rd -= ringB;
            int landed_seen = 0, l_early = 0;
            {
                uint32_t spins = 0;
                while ((landed_seen = ld_acquire_sa(landed_s)) < 1) { if (++spins > (1u << 26)) __trap(); }   // chunk 0
            }
            lds_row(xc, lane_ring + rd);                               // the row of step 0
            rd += pitchB; if (rd >= ringB) rd -= ringB;
            int st_free = 0;                                           // stage of chunk ch - 3
            uint32_t bits_off = uint32_t(s * 32 + lane) * 4u;          // byte offset of this lane's first word of chunk ch
            int prog_seen = 0, cons_seen = 0, p_early = 0, c_early = 0;

            // All waits below are on plain shared counters, read one chunk (or half a chunk) before they are looked at:
            // an mbarrier test costs the warp ~100 cycles even when the phase is long complete, an acquire load ~70.
            // Branches on them are made warp-uniform with a vote, which keeps the common path free of divergence handling.
            auto run_chunk = [&](const int ch, auto head_tag) __attribute__((always_inline)) {
                constexpr bool HEAD = decltype(head_tag)::value;   // some lane starts its row 0 in this chunk (chunks 0 and 1)
                const int t0 = ch * kR;
                // ---- chunk top: free the stage whose last reader has moved on (chunk ch - 3) ----
                __syncwarp();
                mbar_arrive_if_sa(empty_s + uint32_t(st_free) * 8u, lane0 && ch >= 3);
                st_free = ch >= 3 ? (st_free + 1 == nstg ? 0 : st_free + 1) : 0;
                // lane 0's read-ahead crosses into the next stage on the last step: chunk ch + 1 must have landed by then
                const int need_landed = min(ch + 2, nch);
                landed_seen = max(landed_seen, l_early);
                l_early = ld_volatile_sa(landed_s);
                uint32_t pub_sa = 0;
                if (MULTI) {
                    if (has_prev) {
                        // the boundary values of rows t0 .. t0+15: the producer must have published its chunks <= ch + 2
                        const int need = min(ch + 3, nch);
                        prog_seen = max(prog_seen, p_early);
                        if (__any_sync(0xffffffffu, prog_seen < need)) {
                            long long c0 = 0;
                            if (probe_w) c0 = clock64();
                            uint32_t spins = 0;
                            do {
                                prog_seen = ld_acquire_sa(prog_sa + 4u * (s - 1));
                                if (++spins > (1u << 26)) __trap();
                            } while (prog_seen < need);
                            if (probe_w) pc_flag += clock64() - c0;
                        }
                        const uint32_t b0 = bnd_prev + uint32_t(((ch + 1) & (kBndChunks - 1)) * kR) * 4u;
                        const uint32_t b1 = bnd_prev + uint32_t(((ch + 2) & (kBndChunks - 1)) * kR) * 4u;
                        bvals[0] = lds_f32(b0 + 60u);
                        const uint4 u0 = lds_v4(b1), u1 = lds_v4(b1 + 16u), u2 = lds_v4(b1 + 32u), u3 = lds_v4(b1 + 48u);
                        bvals[1] = __uint_as_float(u0.x); bvals[2] = __uint_as_float(u0.y); bvals[3] = __uint_as_float(u0.z); bvals[4] = __uint_as_float(u0.w);
                        bvals[5] = __uint_as_float(u1.x); bvals[6] = __uint_as_float(u1.y); bvals[7] = __uint_as_float(u1.z); bvals[8] = __uint_as_float(u1.w);
                        bvals[9] = __uint_as_float(u2.x); bvals[10] = __uint_as_float(u2.y); bvals[11] = __uint_as_float(u2.z); bvals[12] = __uint_as_float(u2.w);
                        bvals[13] = __uint_as_float(u3.x); bvals[14] = __uint_as_float(u3.y); bvals[15] = __uint_as_float(u3.z);
                        p_early = ld_volatile_sa(prog_sa + 4u * (s - 1));   // looked at one chunk from now
                        st_volatile_if_sa(cons_sa + 4u * s, ch, lane0);      // the values of chunks < ch are in registers
                    }
                    if (has_next) {
                        // this chunk overwrites the slots written 8 chunks ago, which the reader takes in its chunks <= ch - 9
                        const int need = ch - 8;
                        cons_seen = max(cons_seen, c_early);
                        if (__any_sync(0xffffffffu, cons_seen < need)) {
                            long long c0 = 0;
                            if (probe_w) c0 = clock64();
                            uint32_t spins = 0;
                            do {
                                cons_seen = ld_acquire_sa(cons_sa + 4u * (s + 1));
                                if (++spins > (1u << 26)) __trap();
                            } while (cons_seen < need);
                            if (probe_w) pc_flag += clock64() - c0;
                        }
                        c_early = ld_volatile_sa(cons_sa + 4u * (s + 1));
                        pub_sa = bnd_mine + uint32_t((ch & (kBndChunks - 1)) * kR) * 4u;
                    }
                }
                uint32_t word0, word1;
                {
                    // one step; WRAP: the ring may wrap inside the chunk for some lanes; HEAD: some lane starts its row 0 here.
                    // Rows past n - 1 need no special case: what the lanes compute there is never stored or consumed.
#define ISP_MAS_STEP(WRAP, HEAD)                                                                                   \
                    {                                                                                              \
                        if (k == 8) {                                                                              \
                            landed_seen = max(landed_seen, l_early);                                               \
                            if (__any_sync(0xffffffffu, landed_seen < need_landed)) {                              \
                                long long c0 = 0;                                                                  \
                                if (probe_w) c0 = clock64();                                                       \
                                uint32_t spins = 0;                                                                \
                                do {                                                                               \
                                    landed_seen = ld_acquire_sa(landed_s);                                         \
                                    if (++spins > (1u << 26)) __trap();                                            \
                                } while (landed_seen < need_landed);                                               \
                                if (probe_w) pc_full += clock64() - c0;                                            \
                            }                                                                                      \
                        }                                                                                          \
                        const float nxt = __shfl_up_sync(0xffffffffu, q[kC - 1], 1);                               \
                        float xn[kC];                                                                              \
                        lds_row(xn, lane_ring + rd);                                                               \
                        rd += pitchB;                                                                              \
                        if (WRAP) { if (rd >= ringB) rd -= ringB; }                                                \
                        const float v = dp_row(q, xc, left_cur);                                                   \
                        if (HEAD) {                                                                                \
                            if (t0 + k == lane) {                                                                  \
                                /* row 0: Q[0][0] = x[0][0], Q[0][j>0] = -inf   (mas.py:11) */                     \
                                _Pragma("unroll") for (int c = 0; c < kC; ++c) q[c] = (gcol0 + c == 0) ? xc[c] : -CUDART_INF_F; \
                            }                                                                                      \
                        }                                                                                          \
                        acc[k >> 2] = fmaf(v, float(1 << (4 * (k & 3))), acc[k >> 2]);                             \
                        if (MULTI) sts_f32_if(pub_sa + uint32_t(k) * 4u, q[kC - 1], pub);                          \
                        left_cur = lane0 ? bvals[k] : nxt;                                                         \
                        _Pragma("unroll") for (int c = 0; c < kC; ++c) xc[c] = xn[c];                              \
                    }
                    float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
                    for (int k = 0; k < kR; ++k) ISP_MAS_STEP(true, HEAD)
#undef ISP_MAS_STEP
                    word0 = __byte_perm(__float_as_uint(acc[0] + 8388608.0f), __float_as_uint(acc[1] + 8388608.0f), 0x5410);
                    word1 = __byte_perm(__float_as_uint(acc[2] + 8388608.0f), __float_as_uint(acc[3] + 8388608.0f), 0x5410);
                }
                if (BITS_SMEM) {
                    sts_u32(bits_sa + bits_off, word0);
                    sts_u32(bits_sa + bits_off + uint32_t(wpt) * 4u, word1);
                } else {
                    bits_g[bits_off >> 2] = word0;
                    bits_g[(bits_off >> 2) + wpt] = word1;
                }
                bits_off += uint32_t(wpt) * 8u;
                // the publishing lane's 16 stores precede this one in program order; shared memory keeps a thread's stores in order
                if (MULTI) st_volatile_if_sa(prog_sa + 4u * s, ch + 1, pub);
            };
            // One straight-line body per chunk: jumping between specialised variants of the 16 steps cost more in instruction
            // fetch (a few far branches per chunk) than the two instructions per step that the ring-wrap check takes.
            for (int ch = 0; ch < 2 && ch < nch; ++ch) run_chunk(ch, std::true_type{});
            for (int ch = 2; ch < nch; ++ch) run_chunk(ch, std::false_type{});
        }
    } else if (role == 1) {
        // =========================== loader warp of strip s ================================
        if (s < ns_active) {
            const bool lprobe = probe_w && lane == 0;
            long long lp_wait = 0, lp_t0 = 0;
            if (lprobe) lp_t0 = clock64();
            const uint64_t pol = policy_evict_first();
            const CUtensorMap* map = &maps.m[wb / 8 - 1];
            const float* src_b = p.logp + int64_t(b) * p.sB + s * kW;
            // Two cursors over the chunks: `issued` (gated by the strip's "empty" arrivals) and `landed` (the "full"
            // barriers, turned into the plain counter the strip polls).  mbarrier tests cost ~100 cycles of this warp only.
            const bool solo = p.tma != 0;                              // TMA: one lane does everything
            if (!solo || lane == 0) {
                int issued = 0, landed = 0, st_i = 0, st_l = 0;
                uint32_t ph_i = 1u, ph_l = 0u;                         // empty[st_i]: phase (issued / nstg - 1);  full[st_l]: phase (landed / nstg)
                uint32_t idle = 0;
                while (landed < nch) {
                    bool did = false;
                    if (issued < nch && (issued < nstg || mbar_test_sa(empty_s + uint32_t(st_i) * 8u, ph_i))) {
                        const uint32_t full_b = full_s + uint32_t(st_i) * 8u;
                        const int r0 = kR * issued;
                        const uint32_t dst = ring_s + uint32_t(st_i) * stageB;
                        const bool fetch = r0 < n && !(p.dbg & 8);     // past the last row the strip's tail runs on stale rows
                        if (solo) {
                            if (fetch) {
                                mbar_expect_tx_sa(full_b, stageB);
                                tma_load_box(dst, map, s * kW, r0, b, full_b, pol);
                            } else {
                                mbar_arrive_sa(full_b);
                            }
                        } else {
                            // unaligned base or strides: 4 B async copies, one column per lane and pass
                            if (fetch) {
                                const int rows = min(kR, n - r0);
                                for (int r = 0; r < rows; ++r)
                                    for (int col = lane; col < mcols; col += 32)
                                        cp_async4(dst + uint32_t(r) * pitchB + uint32_t(col) * 4u, src_b + int64_t(r0 + r) * p.sT1 + col);
                            }
                            cp_async_arrive_noinc_sa(full_b);
                        }
                        ++issued;
                        if (++st_i == nstg) { st_i = 0; ph_i ^= 1u; }
                        did = true;
                    }
                    if (landed < issued && mbar_test_sa(full_s + uint32_t(st_l) * 8u, ph_l)) {
                        ++landed;
                        if (++st_l == nstg) { st_l = 0; ph_l ^= 1u; }
                        if (solo || lane == 0) st_release_sa(landed_sa + 4u * s, landed);
                        did = true;
                    }
                    if (!did) {
                        __nanosleep(100);
                        if (++idle > (1u << 24)) __trap();
                        if (lprobe) lp_wait += 1;
                    }
                }
            }
            if (lprobe) { p.probe[6] = lp_wait; p.probe[7] = clock64() - lp_t0; }
        }
        __syncwarp();
    } else {
        // =========================== filler warp: zero the outputs =========================
        const size_t cells = size_t(p.T1max) * p.T2max;
        char* zbeg = reinterpret_cast<char*>(p.hard + size_t(b) * cells);
        char* zend = zbeg + cells * sizeof(int16_t);
        char* zb = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(zbeg) + 15) & ~uintptr_t(15));
        char* ze = reinterpret_cast<char*>(reinterpret_cast<uintptr_t>(zend) & ~uintptr_t(15));
        if (ze < zb) ze = zb;
        if (!(p.dbg & 4)) {
            // unaligned head / tail of the dense block (at most 7 int16 each)
            for (char* q2 = zbeg + 2 * lane; q2 < zb && q2 < zend; q2 += 64) *reinterpret_cast<int16_t*>(q2) = 0;
            for (char* q2 = ze + 2 * lane; q2 < zend; q2 += 64) *reinterpret_cast<int16_t*>(q2) = 0;
        } else {
            ze = zb;
        }
        if (p.dur) {
            int64_t* d = p.dur + size_t(b) * p.T2max;
            for (int j = lane; j < p.T2max; j += 32) d[j] = 0;
        }
        if (p.path_fill) {
            int16_t* pp = p.path_ws + size_t(b) * p.T1max;
            for (int r = n + lane; r < p.T1max; r += 32) pp[r] = -1;
        }
        // The filler takes no part in the backtrack: it announces itself at the slot's barrier now (its plain stores above are
        // ordered before the barrier completes) and keeps filling while the others go on.
        __syncwarp();
        asm volatile("bar.arrive %0, %1;" ::"r"(slot + 1), "r"(slot_threads) : "memory");
        if (lane == 0) {
            // The copies follow the slot's progress (chunks of logits landed for strip 0, then blocks of the backtrack done),
            // so that they share HBM with the logit loads without running ahead of them, spill into the backtrack, when HBM
            // is otherwise idle, and are complete shortly before the path's ones are due (fill_done).
            const int pieces = int((size_t(ze - zb) + kZeroPage - 1) / kZeroPage);
            const int nblk_f = (n + 31) >> 5;
            const float wf = 0.65f * float(pieces) / float(nch), wb = 0.50f * float(pieces) / float(nblk_f);
            int issued = 0;
            uint32_t idle = 0;
            char* zp = zb;
            while (issued < pieces) {
                const int fwd = ld_volatile_sa(landed_sa), bt = ld_volatile_sa(sm_sa + 2048 + 192);
                int target = (p.dbg & 16) ? pieces : int(wf * float(fwd) + wb * float(bt)) + 2;
                if (target > pieces) target = pieces;
                if (issued < target) {
                    for (; issued < target; ++issued) {
                        const uint32_t bytes = uint32_t(min(size_t(kZeroPage), size_t(ze - zp)));
                        bulk_s2g(zp, zero_sa, bytes);
                        zp += bytes;
                    }
                    bulk_commit();
                } else {
                    __nanosleep(200);
                    if (++idle > (1u << 24)) __trap();
                }
            }
            bulk_wait_all();
            fence_proxy_async();
            __threadfence_block();
            st_release_sa(sm_sa + 2048 + 196, 1);         // fill_done
        }
        return;
    }

    // bits (shared or global) and the zero-filled outputs become visible to the whole slot
    if (probe) { p.probe[1] = clock64(); p.probe[4] = pc_full; p.probe[5] = pc_flag; p.probe[8] = pc_loop; p.probe[9] = pc_nloop; }
    if (!BITS_SMEM) __threadfence_block();
    asm volatile("bar.sync %0, %1;" ::"r"(slot + 1), "r"(slot_threads) : "memory");
    if (probe) p.probe[2] = clock64();

    // =============================== backtrack ========================================
    // mas.py:20-24.  Blocks of 32 rows, block k = rows i0 .. i0-31 with i0 = n-1-32k; lane t owns row i0-t.
    // The slot's other warps ("converters") turn the skewed words into row-major ones -- rm[row][q] = the backpointer
    // bits of columns 32q .. 32q+31 -- a block at a time, round-robin, into a ring of kRmBlocks blocks laid over the
    // idle logits ring; strip warp 0 runs the dependent chain on them.  Flags are plain shared counters again.
    const int kRmBlocks = (p.dbg & 32) ? 64 : 16;
    const unsigned conv_sleep = (p.dbg & 64) ? 1000u : 40u;
    const int nconv = 2 * ns - 1;                               // strips 1.. and the loaders (the filler is still filling)
    const int nw = (m + 31) >> 5;                               // words per row
    const int nwp = nw | 1;                                     // odd pitch: the 32 rows of a block hit 32 banks
    const int nblk = (n + 31) >> 5;
    const uint32_t rm_sa = ring_sa;
    const uint32_t conv_sa = sm_sa + 2048 + 128;                // [nconv] blocks converted by converter h (it owns blocks k = h mod nconv)
    const uint32_t btdone_sa = sm_sa + 2048 + 192;              // blocks the chain has consumed
    const uint32_t filldone_sa = sm_sa + 2048 + 196;            // set by the filler once every zero has landed

    if (!(role == 0 && s == 0)) {
        // ------------------------------- converter warp -----------------------------------
        const int h = role == 0 ? s - 1 : ns - 1 + s;
        // Bits in the workspace come in through shared memory: the <= 9 word-rows a block of 32 rows touches are one bulk
        // copy, double-buffered per converter behind the row-major ring (an L2 round trip per word would starve the chain).
        constexpr int kStgRows = 9;
        const uint32_t stgB = uint32_t(kStgRows) * uint32_t(wpt) * 4u;
        const uint32_t cstg_sa = rm_sa + ((uint32_t(kRmBlocks * 32 * nwp) * 4u + 127u) & ~127u) + uint32_t(2 * h) * stgB;
        const uint32_t cbar_sa = sm_sa + 640 + uint32_t(2 * h) * 8u;
        const int nct = 2 * ((p.T1max + 31 + kR - 1) / kR) + 1;                 // word-rows per utterance in the workspace
        auto stage_block = [&](int kb, int u) {
            if (lane == 0) {
                const int rlo = max(n - 1 - 32 * kb - 31, 0);
                const int cb = rlo >> 3;
                const uint32_t bytes = uint32_t(min(kStgRows, nct - cb)) * uint32_t(wpt) * 4u;
                mbar_expect_tx_sa(cbar_sa + uint32_t(u) * 8u, bytes);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(cstg_sa + uint32_t(u) * stgB), "l"(bits_g + size_t(cb) * wpt), "r"(bytes), "r"(cbar_sa + uint32_t(u) * 8u)
                             : "memory");
            }
        };
        // The 32 backpointer bits of columns [32 qq, 32 qq + 32) of row `row`: global lane L = 8 qq + i holds row `row` in
        // the word of 8-step chunk (row + l) >> 3 (l = L mod 32) at nibble (row + l) & 7  ->  for i = 0..7 the nibble index
        // runs cyclically from a = row & 7 and the chunk steps once, where a + i reaches 8.  `base`: address of word-row 0.
        auto bits_word = [&](uint32_t base, int row, int qq) -> uint32_t {
            const int a = row & 7;
            const int c0 = (row >> 3) + (qq & 3);
            const uint32_t wa = base + (uint32_t(c0) * uint32_t(wpt) + uint32_t(qq) * 8u) * 4u;
            const uint4 A0 = lds_v4(wa), A1 = lds_v4(wa + 16u);
            const uint4 B0 = lds_v4(wa + uint32_t(wpt) * 4u), B1 = lds_v4(wa + uint32_t(wpt) * 4u + 16u);
            const uint32_t A[8] = {A0.x, A0.y, A0.z, A0.w, A1.x, A1.y, A1.z, A1.w};
            const uint32_t Bv[8] = {B0.x, B0.y, B0.z, B0.w, B1.x, B1.y, B1.z, B1.w};
            const int sh = 4 * a;
            uint32_t out = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint32_t X = (a + i < 8) ? A[i] : Bv[i];
                const uint32_t Y = __funnelshift_r(X, X, sh);           // nibble (a + i) & 7 -> nibble i
                out |= Y & (0xfu << (4 * i));
            }
            return out;
        };
        int done = 0, freed = 0, u = 0;
        uint32_t ph0 = 0u, ph1 = 0u;
        if (!BITS_SMEM) {
            fence_proxy_async();                                  // the ring was read and written through the generic proxy
            if (h < nblk) stage_block(h, 0);
        }
        for (int k = h; k < nblk; k += nconv) {
            uint32_t base = bits_sa;
            if (!BITS_SMEM) {
                if (k + nconv < nblk) stage_block(k + nconv, u ^ 1);
                mbar_wait_sa(cbar_sa + uint32_t(u) * 8u, u ? ph1 : ph0);
                if (u) ph1 ^= 1u; else ph0 ^= 1u;
                base = cstg_sa + uint32_t(u) * stgB - uint32_t(max(n - 1 - 32 * k - 31, 0) >> 3) * uint32_t(wpt) * 4u;
                u ^= 1;
            }
            const int need = k - kRmBlocks + 1;                   // the ring slot's previous block must have been consumed
            uint32_t spins = 0;
            while (freed < need) {
                freed = ld_acquire_sa(btdone_sa);
                if (freed < need) { __nanosleep(conv_sleep); if (++spins > (1u << 24)) __trap(); }
            }
            const int row = n - 1 - 32 * k - lane;
            const uint32_t dst = rm_sa + uint32_t(((k & (kRmBlocks - 1)) * 32 + lane) * nwp) * 4u;
            for (int qq = 0; qq < nw; ++qq) sts_u32(dst + uint32_t(qq) * 4u, row >= 1 ? bits_word(base, row, qq) : 0u);   // row 0 has no predecessor
            __syncwarp();
            ++done;
            if (lane == 0) st_release_sa(conv_sa + 4u * h, done);
        }
        return;
    }

    // ----------------------------------- chain warp (strip warp 0) -------------------------
    int16_t* hard_b = p.hard + size_t(b) * p.T1max * p.T2max;
    int64_t* dur_b = p.dur ? p.dur + size_t(b) * p.T2max : nullptr;
    int16_t* path_b = p.path_ws + size_t(b) * p.T1max;
    long long pc_conv = 0;
    auto wait_conv = [&](int h, int need, int seen) {            // warp-uniform: converter h has finished `need` blocks
        if (__any_sync(0xffffffffu, seen < need)) {
            const uint32_t fa = conv_sa + 4u * uint32_t(h);
            uint32_t spins = 0;
            long long c0 = 0;
            if (probe_w) c0 = clock64();
            // plain loads: the converter's words and its counter are stored in program order by one lane group and shared
            // memory keeps that order; an acquire load would put a MEMBAR on the chain for every block
            while (ld_volatile_sa(fa) < need) { if (++spins > (1u << 26)) __trap(); }
            if (probe_w) pc_conv += clock64() - c0;
        }
    };
    int j = m - 1;           // token index on row i0
    int last_start = n;      // first row of token j+1 (exclusive end of token j)
    const uint32_t lt_mask = (1u << lane) - 1u;
    const bool lane0b = lane == 0;
    const bool want_hard = !(p.dbg & 1), want_dur = dur_b != nullptr && !(p.dbg & 2);
    // Outputs of a finished block: every lane writes its row's 1 and, if a token starts on its row, that token's
    // duration.  Branch-free, and called one iteration late so that it fills the issue slots of the next chain.
    auto emit = [&](int bi0, int bj, uint32_t bmyR, uint32_t bwin, int bk) {
        const int row = bi0 - lane;
        const int col = bj - __clz(bmyR);
        const uint32_t dec = __ballot_sync(0xffffffffu, (bmyR & bwin) != 0u);   // bit t: path leaves row bi0-t diagonally
        const bool valid = row >= 0;
        stg_u16_if(path_b + (valid ? row : 0), col, valid);                     // the 1 itself waits for the zero-fill
        // a token starts on this row if the path leaves it diagonally, or on row 0
        const bool starts = valid && (((dec >> lane) & 1u) || row == 0);
        const uint32_t smask = __ballot_sync(0xffffffffu, starts);
        const uint32_t lower = smask & lt_mask;                                  // starts of later tokens in this block
        const int next_start = lower ? bi0 - (31 - __clz(lower)) : last_start;
        stg_s64_if(dur_b + col, int64_t(next_start - row), starts && want_dur);
        last_start = smask ? bi0 - (31 - __clz(smask)) : last_start;
        st_volatile_if_sa(btdone_sa, bk + 1, lane0b);                            // the block's words were read long ago
    };
    // candidate words of the NEXT block's window are fetched before the chain runs (j moves <= 32)
    uint32_t w0 = 0, w1 = 0, w2 = 0;
    int qw_pref = j >> 5;
    wait_conv(0, 1, 0);
    {
        const uint32_t src = rm_sa + uint32_t(lane * nwp) * 4u;
        w0 = uint32_t(ld_volatile_sa(src + uint32_t(qw_pref) * 4u));
        if (qw_pref > 0) w1 = uint32_t(ld_volatile_sa(src + uint32_t(qw_pref - 1) * 4u));
    }
    // converter and count for block k + 1, and for block k + 2 (whose counter is read a block ahead of its use)
    int h1 = 1 % nconv, need1 = 1 / nconv + 1;
    int h2 = 2 % nconv, need2 = 2 / nconv + 1;
    int f_early = ld_volatile_sa(conv_sa + 4u * uint32_t(h1));
    const uint32_t win_sa = smem_u32(winbuf);
    int k = 0;
    int pi0 = -1, pj = 0;                                     // the previous block, whose outputs are still to be written
    uint32_t pmyR = 0x80000000u, pwin = 0;
    long long pc_pre = 0, pc_chain = 0, pc_epi = 0;
    for (int i0 = n - 1; i0 >= 0; i0 -= 32, ++k) {
        // window of this lane's row: bit k <-> column j-31+k  (the path stays within it)
        long long c_a = 0, c_b = 0, c_c = 0;
        if (probe_w) c_a = clock64();
        const int qw = j >> 5;
        const uint32_t hi = qw == qw_pref ? w0 : w1;
        const uint32_t lo = qw == qw_pref ? w1 : w2;
        const uint32_t win = __funnelshift_rc(lo, hi, (j & 31) + 1);
        sts_u64(win_sa + uint32_t(lane) * 8u, win, win >> 1);
        __syncwarp();
        // all 32 windows into registers first: a shared-memory load inside the dependent chain doubles its time
        // (tools/ubench/chain.cu, B vs D), and the chain is ALU-bound at ~11 cycles per row
        uint4 wv[16];                                           // A_t, A_t >> 1, A_t+1, A_t+1 >> 1
#pragma unroll
        for (int t = 0; t < 16; ++t) wv[t] = lds_v4(win_sa + uint32_t(t) * 16u);
        // prefetch for the block below: its j is in [j-32, j]  ->  word index in {qw, qw-1, qw-2}
        w0 = w1 = w2 = 0u;
        if (k + 1 < nblk) {
            wait_conv(h1, need1, f_early);
            const uint32_t src = rm_sa + uint32_t((((k + 1) & (kRmBlocks - 1)) * 32 + lane) * nwp) * 4u;
            w0 = uint32_t(ld_volatile_sa(src + uint32_t(qw) * 4u));
            w1 = uint32_t(ld_volatile_sa(src + uint32_t(max(qw - 1, 0)) * 4u));
            w2 = uint32_t(ld_volatile_sa(src + uint32_t(max(qw - 2, 0)) * 4u));
            w1 = qw > 0 ? w1 : 0u;
            w2 = qw > 1 ? w2 : 0u;
            f_early = ld_volatile_sa(conv_sa + 4u * uint32_t(h2));
            h1 = h2; need1 = need2;
            if (++h2 == nconv) { h2 = 0; ++need2; }
        }
        qw_pref = qw;
        if (probe_w) c_b = clock64();
        emit(pi0, pj, pmyR, pwin, k - 1);                       // (no branch: the first call has no valid row and writes nothing)
        if (probe_w) c_c = clock64();
        uint32_t myR = 0x80000000u;
        uint32_t R = 0x80000000u;    // one-hot position inside the window; bit 31 <-> column j
#pragma unroll
        for (int t = 0; t < 32; t += 2) {
            const uint4 a = wv[t >> 1];
            myR = lane == t ? R : myR;
            R = bt_step(R, a.x, a.y);
            myR = lane == t + 1 ? R : myR;
            R = bt_step(R, a.z, a.w);
        }
        __syncwarp();                                           // winbuf is rewritten by the next block
        pi0 = i0; pj = j; pmyR = myR; pwin = win;
        if (probe_w) { const long long c_d = clock64(); pc_pre += c_b - c_a; pc_epi += c_c - c_b; pc_chain += c_d - c_c; }
        j -= __clz(R);                                          // R: position after the block's 32 rows (rows < 1 do not move it)
    }
    emit(pi0, pj, pmyR, pwin, k - 1);
    if (probe) p.probe[14] = clock64();
    // the ones: after the last zero has landed
    {
        uint32_t spins = 0;
        while (ld_acquire_sa(filldone_sa) == 0) { __nanosleep(100); if (++spins > (1u << 24)) __trap(); }
    }
    __syncwarp();
    if (want_hard) {
        for (int r0 = 0; r0 < n; r0 += 256) {
            int cols[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) { const int r = r0 + 32 * i + lane; cols[i] = r < n ? int(__ldcg(path_b + r)) : 0; }
#pragma unroll
            for (int i = 0; i < 8; ++i) { const int r = r0 + 32 * i + lane; if (r < n) hard_b[size_t(r) * p.T2max + cols[i]] = 1; }
        }
    }
    if (probe) { p.probe[3] = clock64(); p.probe[10] = pc_conv; p.probe[11] = pc_pre; p.probe[12] = pc_chain; p.probe[13] = pc_epi; }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------

static int g_opt_ring_rows = 0;
static int g_opt_slots = 0;
static int g_opt_dbg = 0;
static int g_opt_bits_global = 0;
static int g_opt_no_tma = 0;
static int g_opt_cols = 0;     // accepted for compatibility with older tools; the kernel has one strip width
static int g_opt_impl = 0;     // 0: isp_mas2.cu where it covers the shape, else isp_mas_cluster.cu, else this file's kernel, else isp_mas_wide.cu;
                               // 1: always this file's kernel, 2: isp_mas2.cu or fail, 3: isp_mas_wide.cu, 4: isp_mas_cluster.cu or fail

int mas_set_option(const char* key, int value, int* prev) {
    if (!strcmp(key, "mas.ring_rows")) { *prev = g_opt_ring_rows; g_opt_ring_rows = value; return 0; }
    if (!strcmp(key, "mas.slots")) { *prev = g_opt_slots; g_opt_slots = value; return 0; }
    if (!strcmp(key, "mas.dbg")) { *prev = g_opt_dbg; g_opt_dbg = value; return 0; }
    if (!strcmp(key, "mas.bits_global")) { *prev = g_opt_bits_global; g_opt_bits_global = value; return 0; }
    if (!strcmp(key, "mas.no_tma")) { *prev = g_opt_no_tma; g_opt_no_tma = value; return 0; }
    if (!strcmp(key, "mas.cols_per_lane")) { *prev = g_opt_cols; g_opt_cols = value; return 0; }
    if (!strcmp(key, "mas.impl")) { *prev = g_opt_impl; g_opt_impl = value; return 0; }
    if (mas_cluster_set_option(key, value, prev) == 0) return 0;
    return mas2_set_option(key, value, prev);
}

struct MasPlan {
    int ns, slots, nstg, wlast;
    bool bits_smem;
    size_t slot_bytes;
    size_t bits_words;     // per utterance
};

static size_t bits_words_for(int T1max, int ns) {
    const size_t nct = 2 * (size_t(T1max + 31 + kR - 1) / kR) + 1;   // 8-step words; + 1: the backtrack reads one past the last
    return nct * size_t(ns) * 32;
}

static int mas_plan(int B, int T1max, int T2max, MasPlan* pl) {
    int sm_count = 148;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
    const size_t smem_limit = 227 * 1024;
    const int ns = (T2max + kW - 1) / kW;
    if (ns > kMaxStrips) return ISP_ERR_UNSUPPORTED;
    pl->ns = ns;
    pl->wlast = (T2max - (ns - 1) * kW + 7) & ~7;
    pl->bits_words = bits_words_for(T1max, ns);
    const size_t bits_bytes = pl->bits_words * 4;
    const size_t stage_bytes = size_t(kR) * size_t((ns - 1) * kW + pl->wlast) * 4;
    const size_t fixed = kSlotHdr + (ns > 1 ? 4 * size_t(ns - 1) * kBnd : 0) + kZeroPage;
    // one utterance per CTA while every utterance still gets its own SM; two beyond that, so
    // that co-resident chains sit on different sub-partitions of one SM
    // (A third utterance per CTA is supported -- "mas.slots" = 3 -- but measured slower than two even at 28 utterances
    // per SM: its ring leaves 5 stages, one more than the lane skew keeps live.)
    int slots = g_opt_slots > 0 ? g_opt_slots : (B > sm_count ? 2 : 1);
    if (slots > kMaxSlots) slots = kMaxSlots;
    while (slots > 1 && 32 * slots * (2 * ns + 1) > kMaxThreads) --slots;
    if (slots == 2 && 2 * ns > 4) slots = 1;       // more strip warps than sub-partitions: co-residency buys a chain nothing
    if (slots == 3 && ns > 2) slots = 1;
    for (;; --slots) {
        const size_t budget = (smem_limit / slots) & ~size_t(127);
        int stg[2] = {0, 0};                      // [0]: bits in the workspace, [1]: bits in shared memory
        for (int in_smem = 0; in_smem < 2; ++in_smem) {
            const size_t need = fixed + (in_smem ? bits_bytes : 0);
            if (need + kMinStages * stage_bytes > budget) continue;
            size_t k = (budget - need) / stage_bytes;
            stg[in_smem] = int(k > kMaxStages ? kMaxStages : k);
        }
        // resident bits save the backtrack an L2 round trip per block; a ring shallower than 8 stages (4 of them
        // live) starves the chain
        int in_smem = stg[1] >= 8 || (stg[1] > 0 && stg[1] >= stg[0]) ? 1 : 0;
        if (g_opt_bits_global && stg[0] > 0) in_smem = 0;
        int nstg = stg[in_smem];
        if (nstg > 0) {
            if (g_opt_ring_rows > 0) {
                int want = (g_opt_ring_rows + kR - 1) / kR;
                if (want < kMinStages) want = kMinStages;
                if (want < nstg) nstg = want;
            }
            // the backtrack lays its row-major ring (16 blocks) and the converters' staging buffers over the logits ring
            const size_t nwp = size_t((T2max + 31) / 32) | 1;
            const size_t bt_bytes = ((16 * 32 * nwp * 4 + 127) & ~size_t(127)) + (in_smem ? 0 : size_t(2 * ns - 1) * 2 * 9 * ns * 128);
            if (bt_bytes > size_t(nstg) * stage_bytes) { if (slots == 1) break; continue; }
            pl->slots = slots;
            pl->nstg = nstg;
            pl->bits_smem = in_smem != 0;
            pl->slot_bytes = (fixed + size_t(nstg) * stage_bytes + (in_smem ? bits_bytes : 0) + 127) & ~size_t(127);
            return 0;
        }
        if (slots == 1) break;
    }
    return ISP_ERR_UNSUPPORTED;
}

// tensor maps over the logits: (T2max, T1max, B) fp32, box = (8 i columns, kR rows, 1), no swizzle
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static bool make_maps(MasMaps* maps, const float* logp, int64_t sB, int64_t sT1, int B, int T1max, int T2max) {
    static PFN_encodeTiled enc = nullptr;
    if (!enc) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess || !ptr) return false;
        enc = reinterpret_cast<PFN_encodeTiled>(ptr);
    }
    if ((reinterpret_cast<uintptr_t>(logp) & 15) || (sT1 & 3) || (sB & 3)) return false;
    // Encoding 16 maps costs the host ~20 us; a training loop calls with the same buffer geometry over and over (the
    // caching allocator hands the same block back), so the last set is kept per thread.  A map holds no device state.
    struct Key { const float* p; int64_t sB, sT1; int B, T1, T2; };
    static thread_local Key last = {nullptr, 0, 0, 0, 0, 0};
    static thread_local MasMaps last_maps;
    if (last.p == logp && last.sB == sB && last.sT1 == sT1 && last.B == B && last.T1 == T1max && last.T2 == T2max) {
        *maps = last_maps;
        return true;
    }
    cuuint64_t dims[3] = {cuuint64_t(T2max), cuuint64_t(T1max), cuuint64_t(B)};
    cuuint64_t strides[2] = {cuuint64_t(sT1) * 4, cuuint64_t(B > 1 ? sB : sT1 * T1max) * 4};
    cuuint32_t estr[3] = {1, 1, 1};
    for (int i = 0; i < kNumBox; ++i) {
        cuuint32_t box[3] = {cuuint32_t(8 * (i + 1)), cuuint32_t(kR), 1};
        CUresult r = enc(&maps->m[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(logp), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { last.p = nullptr; return false; }
    }
    last = Key{logp, sB, sT1, B, T1max, T2max};
    last_maps = *maps;
    return true;
}

size_t mas_workspace_bytes(int B, int T1max, int T2max) {
    if (B <= 0 || T1max <= 0 || T2max <= 0) return 0;
    if (T2max > ISP_MAS_MAX_T2) return std::max(mas_wide_workspace_bytes(B, T1max, T2max), mas_cluster_workspace_bytes(B, T1max, T2max));   // isp_mas_cluster.cu / isp_mas_wide.cu
    const int ns = (T2max + kW - 1) / kW;
    const size_t v1 = 256 + size_t(B) * bits_words_for(T1max, ns) * 4 + ((size_t(B) * T1max * 2 + 15) & ~size_t(15));
    const size_t v2 = mas2_workspace_bytes(B);
    const size_t v3 = mas_wide_workspace_bytes(B, T1max, T2max);     // "mas.impl" = 3 runs the general kernel on any shape (tests)
    return std::max(std::max(v1, mas_cluster_workspace_bytes(B, T1max, T2max)), std::max(v2, v3));
}

template <bool BS, bool MULTI>
static int launch_one(const MasMaps& maps, const MasParams& p, const MasPlan& pl, cudaStream_t stream) {
    auto kern = mas_kernel<BS, MULTI>;
    const size_t smem = pl.slot_bytes * pl.slots;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(mas_kernel)");
    const int grid = (p.B + pl.slots - 1) / pl.slots;
    kern<<<grid, 32 * pl.slots * (2 * pl.ns + 1), smem, stream>>>(maps, p);
    e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "mas_kernel launch");
    return 0;
}

// would mas_forward hand this shape to the kernel that takes isp_align_forward's ready counts?
bool mas_linkable(int B, int T1max, int T2max) {
    if (T2max > ISP_MAS_MAX_T2 || g_opt_impl == 3 || g_opt_impl == 1 || g_opt_bits_global || g_opt_slots == 3) return false;
    return mas2_linkable(B, T1max, T2max, g_opt_dbg);
}

int mas_forward(const float* logp, int64_t sB, int64_t sT1, int64_t sT2,
                const int64_t* text_len, const int64_t* mel_len, int B, int T1max, int T2max,
                int16_t* attn_hard, int64_t* durations, int16_t* path, void* ws, size_t ws_bytes, cudaStream_t stream,
                const int* ready, int ready_need) {
    // (ready: isp_align_forward's per-utterance tile counts; only isp_mas2.cu starts before the kernel in front has completed,
    // the other kernels are launched plainly behind it and find every count final)
    if (!logp || !text_len || !mel_len || !ws) { set_error("isp_mas_forward: null pointer"); return ISP_ERR_INVALID; }
    if (!attn_hard && !path) { set_error("isp_mas_forward: attn_hard may be NULL only when the path is returned (isp_mas_forward_path)"); return ISP_ERR_INVALID; }
    if (B <= 0 || T1max <= 0 || T2max <= 0) { set_error("isp_mas_forward: B, T1max, T2max must be positive"); return ISP_ERR_INVALID; }
    if (sT2 != 1) { set_error("isp_mas_forward: sT2 must be 1 (token axis contiguous), got %lld", (long long)sT2); return ISP_ERR_INVALID; }
    if (sT1 < T2max || (B > 1 && sB < int64_t(T1max - 1) * sT1 + T2max)) { set_error("isp_mas_forward: overlapping strides"); return ISP_ERR_INVALID; }
    if (T1max >= (1 << 24) || (T2max > ISP_MAS_MAX_T2 && !mas_wide_supported(T2max))) {
        set_error("isp_mas_forward: T2max=%d > %d or T1max=%d >= 2^24 is not covered", T2max, ISP_MAS_WIDE_MAX_T2, T1max);
        return ISP_ERR_UNSUPPORTED;
    }
    if (ws_bytes < mas_workspace_bytes(B, T1max, T2max) || (reinterpret_cast<uintptr_t>(ws) & 15)) {
        set_error("isp_mas_forward: workspace too small or not 16 B aligned (%zu < %zu)", ws_bytes, mas_workspace_bytes(B, T1max, T2max));
        return ISP_ERR_WORKSPACE;
    }
    if (g_opt_impl == 4) {                               // forced (tests): the cluster kernel on any shape it covers
        if (!mas_cluster_supported(B, T1max, T2max)) { set_error("isp_mas_forward: mas.impl=4 but T1max=%d T2max=%d is outside isp_mas_cluster.cu's range", T1max, T2max); return ISP_ERR_UNSUPPORTED; }
        return mas_cluster_forward(logp, sB, sT1, text_len, mel_len, B, T1max, T2max, attn_hard, durations, path, ws, stream);
    }
    const bool want2 = g_opt_impl == 0 || g_opt_impl == 2;
    const bool fits2 = want2 && !g_opt_bits_global && g_opt_slots != 3 && T2max <= ISP_MAS_MAX_T2 && mas2_supported(B, T1max, T2max);
    // One thread-block cluster per utterance (isp_mas_cluster.cu): everything from 641 to 1024 tokens (~10x the general kernel), and
    // below that every shape outside isp_mas2.cu's range whose clusters all run at once (tools/masc_compare.py: BASELINE configs[3],
    // 16 x 4096 x 512, 225 us against 280 us for this file's single-CTA kernel; 16 x 1000 x 300 90 against 94).  With a second wave
    // of clusters the single-CTA kernel wins (37 x 2500 x 512: 276 against 189 us) and keeps the shape.
    if (g_opt_impl == 0 && !fits2 && mas_cluster_supported(B, T1max, T2max) && (T2max > ISP_MAS_MAX_T2 || mas_cluster_one_wave(B, T1max, T2max)))
        return mas_cluster_forward(logp, sB, sT1, text_len, mel_len, B, T1max, T2max, attn_hard, durations, path, ws, stream);
    if (T2max > ISP_MAS_MAX_T2 || g_opt_impl == 3)       // wider than the strip kernels: the general kernel (or forced, for tests)
        return mas_wide_forward(logp, sB, sT1, text_len, mel_len, B, T1max, T2max, attn_hard, durations, path, ws, stream);
    if (fits2)
        return mas2_forward(logp, sB, sT1, text_len, mel_len, B, T1max, T2max, attn_hard, durations, path, ws,
                            g_opt_no_tma, g_opt_ring_rows, g_opt_slots, g_opt_dbg, stream, ready, ready_need);
    if (g_opt_impl == 2) { set_error("isp_mas_forward: mas.impl=2 but T1max=%d T2max=%d is outside isp_mas2.cu's range", T1max, T2max); return ISP_ERR_UNSUPPORTED; }
    MasPlan pl;
    int rc = mas_plan(B, T1max, T2max, &pl);
    if (rc) { set_error("isp_mas_forward: no kernel configuration for T1max=%d T2max=%d", T1max, T2max); return rc; }

    MasParams p;
    p.logp = logp; p.sB = sB; p.sT1 = sT1;
    p.text_len = text_len; p.mel_len = mel_len;
    p.B = B; p.T1max = T1max; p.T2max = T2max;
    p.hard = attn_hard; p.dur = durations;
    p.status = reinterpret_cast<int*>(ws);
    p.probe = reinterpret_cast<long long*>(reinterpret_cast<char*>(ws) + 64);
    p.bits_ws = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(ws) + 256);
    p.bits_stride = int64_t(pl.bits_words);
    p.path_ws = path ? path : reinterpret_cast<int16_t*>(reinterpret_cast<char*>(ws) + 256 + size_t(B) * pl.bits_words * 4);
    p.path_fill = path ? 1 : 0;
    p.ns = pl.ns; p.slots = pl.slots; p.nstg = pl.nstg; p.wlast = pl.wlast;
    p.slot_bytes = int(pl.slot_bytes); p.dbg = g_opt_dbg;
    if (!attn_hard) p.dbg |= 1 | 4;      // no dense output: neither its zero fill nor the path's ones (the path itself is returned)
    MasMaps maps;
    memset(&maps, 0, sizeof(maps));
    p.tma = (!g_opt_no_tma && make_maps(&maps, logp, sB, sT1, B, T1max, T2max)) ? 1 : 0;

    cudaError_t e = cudaMemsetAsync(ws, 0, 256, stream);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(status)");

    if (pl.bits_smem) return pl.ns > 1 ? launch_one<true, true>(maps, p, pl, stream) : launch_one<true, false>(maps, p, pl, stream);
    return pl.ns > 1 ? launch_one<false, true>(maps, p, pl, stream) : launch_one<false, false>(maps, p, pl, stream);
}

}  // namespace isp
