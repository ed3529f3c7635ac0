// Monotonic Alignment Search for sm_100a, second kernel: two columns per lane, a backtrack that is parallel
// over groups of 32 frames, utterances paired longest-with-shortest on an SM.
//
// Reference semantics (paths relative to the reference root):
//   tts/modules/aligner/mas.py:8-26        mas_width1 (DP, tie rule, backtrack)
//   tts/modules/aligner/cuda_mas.py:11-46  cuda_b_mas (the GPU kernel this replaces)
//   tts/models/acoustic/modules/alignment.py:275  durations = attn_hard.sum(dim=1)
// tools/mas2_model.py is an executable model of the layouts below (checked against the oracle on the CPU).
//
// What bounds the path is one dependent chain of T1 row steps per utterance plus whatever follows it, so:
//
//   * Forward.  A strip warp owns 64 text columns, lane l the columns 2l, 2l+1, and at local step u it is on row
//     u - l (the wavefront is skewed across lanes, which takes the exchange of the lane-boundary value off the
//     dependent chain).  Per step a lane issues 11 instructions: one 8 B shared load of its logits (tiled TMA box
//     ring, rows read diagonally; the row pitch is an odd multiple of 8 B, so the skewed reads are conflict-free),
//     one 4 B store of its last column and one 4 B load of its left neighbour's (a 16-slot array per lane whose
//     addresses are immediates -- no select for lane 0, no register file of boundary values: lane 0 of a strip
//     points at a ring the previous strip's lane 31 writes, lane 0 of strip 0 at a constant -inf page), 2 x
//     (FSET, FMNMX, FADD, FFMA): the backpointer bit is accumulated as a float, one 32-bit word per lane and
//     16 steps.  Nothing in a step is conditional: rows before the first are zeros in the ring and -inf in the
//     accumulators, with Q[-1][-1] = 0 standing in for the reference's special first row (mas.py:11-12).
//   * Everything that synchronises happens once per 16 steps and only through plain shared counters read a
//     chunk ahead (an mbarrier test costs its warp >= 80 cycles even on a completed phase): a loader warp per
//     strip turns the TMA "full" barriers into a counter, strips publish their progress for their neighbours.
//   * Backtrack.  j <- j - bit[i][j] is a serial chain of T1 dependent lookups; instead, while the forward sweep
//     is still running, helper warps (a) transpose the strips' words into row-major ones, 32 rows x 64 columns
//     per pass (five butterfly stages through shuffles), and (b) compose, per group of 32 rows, the map "column
//     at the group's last row -> column at the row above its first", bit-sliced over the columns (8 planes of
//     the column index; a row costs one funnel shift and one LOP3 per plane and 32 columns).  What is left
//     after the sweep is one lookup per group (T1/32 hops), then every group walks its own 32 rows from its
//     entry column, one lane per group, and the durations fall out of the path's change points.
//   * Placement.  A plan kernel sorts the utterances by length; CTA c takes the c-th longest and, when there are
//     more utterances than SMs, the c-th shortest next to it.  The two share the SM's shared memory in
//     proportion to what they need (rings of equal depth); a pair that would leave either ring shallow runs one
//     after the other instead.
//   * The dense int16 output is zero-filled by bulk shared->global copies of a zero page, paced by the sweep;
//     the path's ones follow once the last copy has landed.
//
// Bit-exactness: each cell does exactly the reference's one fp32 add on top of an exact max; the comparison is
// the reference's `>=` (ties and -inf >= -inf take the diagonal).  Everything after the bits is integer.

#include <cuda.h>
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "common.cuh"
#include "isp_internal.h"
#include "isp_mas_ptx.cuh"

#ifndef ISP_MAS2_PROBE
#define ISP_MAS2_PROBE 0
#endif
#ifndef ISP_MAS2_SHFL
#define ISP_MAS2_SHFL 0
#endif
#ifndef ISP_MAS2_ABL
#define ISP_MAS2_ABL 0          // micro-benchmark only (tools/ubench/step2.cu): 1 drops the neighbour exchange, 2 the logit loads
#endif

namespace isp {
namespace mas2 {

constexpr int kR = 16;                  // steps per chunk = rows per TMA box = rows per ring stage
constexpr int kStrip = 64;              // text columns per strip warp
constexpr int kMaxNS = 4;               // strips per utterance -> T2max <= 256
constexpr int kPlanes = 8;              // bits of a column index
constexpr int kMapWords = 8;            // row-major words per row (kMaxNS * 2)
constexpr int kMaxStages = 24;
constexpr int kRm = 256;                // rows in the row-major ring between transposer and mapper (8 blocks)
constexpr int kBnd = 128;               // slots in a strip-boundary ring
constexpr int kZeroPage = 4096;
constexpr int kHdr = 2048;              // per slot: barriers and counters
constexpr int kVLane = 68;              // bytes per lane in a strip's exchange array: 16 slots + 1 word of padding
constexpr int kVBytes = 33 * kVLane + kBnd * 4 + 12;   // 33 lanes (the 33rd is lane 31's dump) + the boundary ring = 2768
constexpr int kWarpsPerSlot = 11;       // 4 strips, 4 loaders, filler, transposer, mapper
constexpr int kSlotThreads = 32 * kWarpsPerSlot;
constexpr int kThreads = 2 * kSlotThreads;
constexpr int kNumBox = kStrip / 8;
constexpr uint32_t kSmemTotal = 227 * 1024;
constexpr int kRankMax = 512;            // up to here a CTA ranks the batch itself; beyond, a plan kernel sorts it

// header offsets (bytes from the slot's header)
constexpr uint32_t kOffFull = 0;                                  // [kMaxNS][kMaxStages] u64
constexpr uint32_t kOffEmpty = kMaxNS * kMaxStages * 8;           // 768
constexpr uint32_t kOffProg = 2 * kMaxNS * kMaxStages * 8;        // 1536: [kMaxNS] chunks completed by strip s
constexpr uint32_t kOffLanded = kOffProg + 16;                    // [kMaxNS] chunks of logits landed for strip s
constexpr uint32_t kOffTdone = kOffLanded + 16;                   // blocks of 32 rows transposed
constexpr uint32_t kOffMdone = kOffTdone + 4;                     // blocks of 32 rows the mapper has consumed
constexpr uint32_t kOffFillDone = kOffMdone + 4;
constexpr uint32_t kOffSlotDone = kOffFillDone + 4;
constexpr uint32_t kOffPathDone = kOffSlotDone + 4;               // the path is in shared memory (mapper warp -> transposer warp)
constexpr uint32_t kOffOnesDone = kOffPathDone + 4;               // the transposer warp has written its half of the ones
constexpr uint32_t kOffNegInf = 1600;                             // 16 floats of -inf

struct Params {
    const float* logp;
    int64_t sB, sT1;
    const int64_t* text_len;
    const int64_t* mel_len;
    int B, T1max, T2max;
    int16_t* hard;
    int64_t* dur;
    int16_t* path;          // caller's (B, T1max) int16 or nullptr
    int* status;
    long long* probe;
    const int* order;       // utterances, longest first (plan kernel); nullptr: every CTA ranks the batch itself (B <= kRankMax)
    unsigned char* bad;     // (B) per utterance: 1 if a length was outside [1, Tmax]
    long long* trace;       // (B, 4) globaltimer at the start / end of every utterance, CTA / slot, ring stages (debug), or nullptr
    int nsingle;            // CTAs [0, nsingle) hold one utterance, the rest two
    int tma;
    int dbg;
    int max_stages;         // cap on ring stages (tests: few rows in flight)
    int min_pair_stages;    // a pair whose rings would be shallower than this runs one after the other
    float fill_cycles;      // SM cycles over which an utterance's zero fill is spread
    float pace_cycles_per_step;   // what a step of the longest chain is expected to take (0: no pacing of the loaders)
    int together;           // 1: the short member of a pair starts with the long one (tests)
    int linger;             // 1: every CTA is resident from the start, so a slot may outlive its sweep to keep the fill slow
    int early_lengths;      // 1: the lengths are not written by the grid in front of this one: they may be read before it completes
    const int* ready;       // isp_align_forward: ready[b] reaches ready_need when utterance b's logits are in global memory (nullptr:
    int ready_need;         // the logits are complete when the grid in front of this one is)
    unsigned long long* origin;   // linked launches: %globaltimer of the first CTA to start (0 until then); the windows of the
    float cycles_per_ns;          // paced loads and of the zero fill are counted from it, not from each CTA's own (staggered) start
};

struct Maps { CUtensorMap m[kNumBox]; };

// Geometry of one utterance in shared memory: everything but the logits ring.
struct Geo {
    int n, m, bad;          // frames, tokens (clamped), out-of-contract flag
    int nl, nlp, ns;        // lanes, lanes rounded to 4 (pitch of W), strips
    int nch;                // chunks of a strip's n + 31 steps
    int g1;                 // groups of 32 rows
    uint32_t cols;          // sum over strips of the box widths (floats per ring row set)
    uint32_t fixed;         // bytes besides header and ring
    uint32_t off_v, off_w, off_maps, off_rm, off_entry, off_path, off_start;   // from the slot's body
};

__host__ __device__ inline uint32_t align16(uint32_t x) { return (x + 15u) & ~15u; }

__host__ __device__ inline Geo make_geo(long long n64, long long m64, int T1max, int T2max) {
    Geo g;
    g.bad = n64 < 1 || n64 > T1max || m64 < 1 || m64 > T2max;
    g.n = int(n64 < 1 ? 1 : (n64 > T1max ? T1max : n64));
    g.m = int(m64 < 1 ? 1 : (m64 > T2max ? T2max : m64));
    g.nl = (g.m + 1) >> 1;
    g.nlp = (g.nl + 3) & ~3;
    g.ns = (g.nl + 31) >> 5;
    g.nch = (g.n + 31 + kR - 1) / kR;
    g.g1 = ((g.n - 1) >> 5) + 1;
    g.cols = uint32_t((g.m + 7) & ~7);                 // only the last strip's box is narrower than 64
    uint32_t off = 0;
    g.off_v = off;      off += uint32_t(g.ns) * kVBytes;
    g.off_w = off;      off += align16(uint32_t(g.nch + 2) * uint32_t(g.nlp) * 4u);
    g.off_maps = off;   off += uint32_t(g.g1) * kPlanes * kMapWords * 4u;
    g.off_rm = off;     off += kRm * kMapWords * 4u;
    g.off_entry = off;  off += align16(uint32_t(g.g1 + 1) * 4u);
    g.off_path = off;   off += align16(uint32_t(g.n) * 2u);
    g.off_start = off;  off += align16(uint32_t(g.m + 2) * 2u);
    g.fixed = (off + 127u) & ~127u;                    // the ring behind it takes TMA boxes: 128 B aligned
    return g;
}

__host__ __device__ inline uint32_t stage_bytes(const Geo& g) { return uint32_t(kR) * g.cols * 4u; }

// ---- small helpers ---------------------------------------------------------------------------------------
ISP_DEVINL float2 lds_f32x2(uint32_t saddr) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(saddr));
    return v;
}
ISP_DEVINL uint32_t lds_u32(uint32_t saddr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}
ISP_DEVINL void sts_u16(uint32_t saddr, int v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(saddr), "h"(short(v)) : "memory"); }
ISP_DEVINL int lds_s16(uint32_t saddr) {
    short v;
    asm volatile("ld.shared.s16 %0, [%1];" : "=h"(v) : "r"(saddr));
    return int(v);
}
// SM cycles since the launch's first CTA started (linked launches: the CTAs of this grid become resident one by one)
ISP_DEVINL float cycles_since_origin(unsigned long long* origin, float cycles_per_ns) {
    unsigned long long now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
    const unsigned long long old = atomicCAS(origin, 0ull, now);
    return old == 0ull || now <= old ? 0.0f : float(now - old) * cycles_per_ns;
}
ISP_DEVINL long long gtimer() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
ISP_DEVINL void named_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }
ISP_DEVINL void wait_counter_idle(uint32_t sa, int need, unsigned ns) {       // helper warps: sleep between polls
    uint32_t spins = 0;
    while (ld_acquire_sa(sa) < need) { __nanosleep(ns); if (++spins > (1u << 26)) __trap(); }
}

// =================================== the strip warp: forward DP ===========================================
struct StripCtx {
    uint32_t lane_ring;     // ring base + this lane's column offset
    uint32_t pitchB, ringB;
    uint32_t full_s, empty_s, landed_s, prog_sa;
    uint32_t w_sa;          // this lane's word of chunk 0
    uint32_t w_step;        // bytes between chunks in W
    int nstg, nch, s, lane;
    bool has_prev, has_next, w_ok;
    uint32_t v_rd, v_wr;    // exchange array addresses of lanes 1..31 (read) / 0..30 (write); lane 0 / 31 see below
    uint32_t bnd_prev, bnd_mine;
};

// hard[r][path[r]] = 1 for the rows of one parity of 256-row blocks; eight rows per lane and pass so that the loads overlap
ISP_DEVINL void write_ones(int16_t* hard_b, uint32_t path_sa, int n, int T2max, int lane, int parity) {
    for (int r0 = 256 * parity; r0 < n; r0 += 512) {
        int pj[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { const int r = r0 + 32 * i + lane; pj[i] = r < n ? lds_s16(path_sa + 2u * r) : 0; }
#pragma unroll
        for (int i = 0; i < 8; ++i) { const int r = r0 + 32 * i + lane; if (r < n) hard_b[size_t(r) * T2max + pj[i]] = 1; }
    }
}

// Out of line on purpose: the sweep's registers stay in registers, and the rare wait costs a call.
__device__ __noinline__ int3 poll_flags(uint32_t landed_sa, uint32_t prev_sa, uint32_t next_sa, int need_l, int need_p, int need_n) {
    int3 r = make_int3(0, 0x7fffffff, 0x7fffffff);
    uint32_t spins = 0;
    for (;;) {
        r.x = ld_acquire_sa(landed_sa);
        if (prev_sa) r.y = ld_acquire_sa(prev_sa);
        if (next_sa) r.z = ld_acquire_sa(next_sa);
        if (r.x >= need_l && r.y >= need_p && r.z >= need_n) break;
        __nanosleep(20);                                    // a waiting strip must not take issue slots from the warps it waits for
        if (++spins > (1u << 26)) __trap();
    }
    return r;
}

// FULL: the strip is 64 columns wide (every strip but an utterance's last), so the ring pitch is a constant and the
// logits' addresses are immediates.
template <bool FULL>
ISP_DEVINL void strip_forward(const StripCtx& c, bool probe_w, long long* pc_wait) {
    const int lane = c.lane, s = c.s, nch = c.nch, nstg = c.nstg;
    const bool lane0 = lane == 0, lane31 = lane == 31;
    const uint32_t pitchB = FULL ? uint32_t(kStrip * 4) : c.pitchB;
    float q0 = -CUDART_INF_F, q1 = -CUDART_INF_F;
    // Q[-1][-1] = 0 makes row 0 come out as mas.py:11-12 wants it: Q[0][0] = x[0][0] + max(0, -inf), Q[0][j>0] = -inf
    float left = (s == 0 && lane0) ? 0.0f : -CUDART_INF_F;
    float lnext = -CUDART_INF_F;                            // `left` of the step after this one
    // ring row (byte offset) of the row of step 2: real row r lives in ring row r + 32 (two stages of zeros in front)
    uint32_t rd = uint32_t(34 - lane) * pitchB;
    // Flags (chunks of logits landed, chunks completed by the two neighbour strips) are plain shared counters, read in the
    // middle of a chunk and looked at on top of the next one: an mbarrier test costs the warp ~100 cycles even when the phase
    // is long complete, an acquire load ~70.  Only a flag that was not far enough half a chunk early sends the warp to poll.
    int l_seen = 0, pp_seen = 0, pn_seen = 0;
    const uint32_t prev_sa = c.has_prev ? c.prog_sa + 4u * uint32_t(s - 1) : 0u, next_sa = c.has_next ? c.prog_sa + 4u * uint32_t(s + 1) : 0u;
    auto poll = [&](int need_l, int need_p, int need_n) __attribute__((always_inline)) {
        long long c0 = 0;
        if (probe_w) c0 = clock64();
        const int3 r = poll_flags(c.landed_s, prev_sa, next_sa, need_l, need_p, need_n);
        if (probe_w) pc_wait[r.x >= need_l && l_seen >= need_l ? 1 : 0] += clock64() - c0;
        l_seen = r.x; pp_seen = r.y; pn_seen = r.z;
    };
    poll(1, min(3, nch), -8);                               // chunk 0's own needs
    // Loads are issued two steps ahead of their use: a step is ~30 cycles of issue, a shared-memory load ~30 cycles of latency
    // (more behind a store to the same word), and the warp issues in order.
    float2 xc = lds_f32x2(c.lane_ring + uint32_t(32 - lane) * pitchB);        // the row of step 0
    float2 xn = lds_f32x2(c.lane_ring + uint32_t(33 - lane) * pitchB);        // the row of step 1
    if (c.has_prev && lane0) lnext = lds_f32(c.bnd_prev + 31u * 4u);          // step 1: the previous strip's row 0
    int st_free = 0;                                                          // stage of (shifted) chunk ch - 1
    uint32_t w_sa = c.w_sa;

    for (int ch = 0; ch < nch; ++ch) {
        // ---- chunk top: everything that synchronises ----
        mbar_arrive_if_sa(c.empty_s + uint32_t(st_free) * 8u, lane0 && ch >= 1);
        st_free = ch >= 1 ? (st_free + 1 == nstg ? 0 : st_free + 1) : 0;
        {
            // lane 0's read-ahead crosses into the next stage on the last step: chunk ch + 1 must have landed by then;
            // lane 0's left neighbours of steps 16ch+1 .. 16ch+16 were written by the previous strip's lane 31 at its steps
            // 16ch+31 .. 16ch+46: it must have completed its chunk ch + 2; this chunk overwrites the boundary slots of chunk
            // ch - 8, which the next strip reads in its chunks <= ch - 9
            const int need_l = min(ch + 2, nch), need_p = min(ch + 3, nch), need_n = ch - 8;
            const bool bad = l_seen < need_l || (c.has_prev && pp_seen < need_p) || (c.has_next && pn_seen < need_n);
            if (__any_sync(0xffffffffu, bad)) poll(need_l, need_p, need_n);
        }
        // a lane's `left` of step u + 2 is its left neighbour's last column after step u: lanes 1..31 read it right behind the
        // neighbour's store of this step; lane 0 of a later strip reads what the previous strip's lane 31 stored at its step u + 32
        uint32_t rB = c.v_rd, wV = c.v_wr;
        if (c.has_prev && lane0) rB = c.bnd_prev + uint32_t(((ch + 2) & (kBnd / kR - 1)) * kR) * 4u;
        if (c.has_next && lane31) wV = c.bnd_mine + uint32_t((ch & (kBnd / kR - 1)) * kR) * 4u;
        const uint32_t xa = c.lane_ring + rd;
        float acc0 = 0.0f, acc1 = 0.0f;
#pragma unroll
        for (int k = 0; k < kR; ++k) {
            if (k == 8) {
                l_seen = ld_volatile_sa(c.landed_s);
                if (c.has_prev) pp_seen = ld_volatile_sa(c.prog_sa + 4u * uint32_t(s - 1));
                if (c.has_next) pn_seen = ld_volatile_sa(c.prog_sa + 4u * uint32_t(s + 1));
            }
#if ISP_MAS2_ABL & 2
            const float2 xn2 = xn;
#else
            const float2 xn2 = lds_f32x2(xa + uint32_t(k) * pitchB);                              // the row of step k + 2
#endif
            const float s1 = set_ge(q0, q1);                    // mas.py:17 -- ties take j-1
            const float m1 = fmaxf(q0, q1);
            const float s0 = set_ge(left, q0);
            const float m0 = fmaxf(left, q0);
            q1 = xc.y + m1;                                     // mas.py:14 -- one fp32 add per cell
#if ISP_MAS2_ABL & 1
            const float l2 = lnext;
#elif ISP_MAS2_ABL & 4
            sts_f32(wV + uint32_t(4 * k), q1);
            const float l2 = lnext;
#elif ISP_MAS2_ABL & 8
            const float l2 = lds_f32(rB + uint32_t(4 * k));
#elif ISP_MAS2_SHFL
            // the exchange through a shuffle: lane 0 takes the previous strip's boundary value (or -inf), lane 31 publishes its own
            float l2 = __shfl_up_sync(0xffffffffu, q1, 1);
            if (c.has_prev) { const float bv = lds_f32(rB + uint32_t(4 * k)); l2 = lane0 ? bv : l2; }
            else l2 = lane0 ? -CUDART_INF_F : l2;
            if (c.has_next) sts_f32_if(wV + uint32_t(4 * k), q1, lane31);
#else
            sts_f32(wV + uint32_t(4 * k), q1);
            const float l2 = lds_f32(rB + uint32_t(4 * k));     // `left` of step k + 2
#endif
            q0 = xc.x + m0;
            if (k < 8) {
                acc0 = fmaf(s0, float(1 << (2 * k)), acc0);
                acc0 = fmaf(s1, float(2 << (2 * k)), acc0);
            } else {
                acc1 = fmaf(s0, float(1 << (2 * (k - 8))), acc1);
                acc1 = fmaf(s1, float(2 << (2 * (k - 8))), acc1);
            }
            left = lnext; lnext = l2;
            xc = xn; xn = xn2;
        }
        rd += uint32_t(kR) * pitchB;
        if (rd >= c.ringB) rd -= c.ringB;
        const uint32_t word = __byte_perm(__float_as_uint(acc0 + 8388608.0f), __float_as_uint(acc1 + 8388608.0f), 0x5410);
        st_volatile_if_sa(w_sa, int(word), c.w_ok);
        w_sa += c.w_step;
        // lane 31's stores of the chunk precede this one in program order; shared memory keeps a thread's stores in order
        st_volatile_if_sa(c.prog_sa + 4u * uint32_t(s), ch + 1, lane31);
    }
}

// =================================== the kernel =============================================================
__global__ void __launch_bounds__(kThreads, 1)
mas2_kernel(const __grid_constant__ Maps maps, const Params p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t smem_sa = smem_u32(smem_raw);
    // Programmatic dependent launch: this grid may have become resident while the grid in front of it in the stream (the
    // log-likelihood kernel) is still draining.  Nothing in global memory is read or written before griddepcontrol.wait, except --
    // when the caller vouches that the previous grid does not write them -- the lengths, so that the ranking, the geometry and the
    // barrier set-up run under the previous kernel's tail.  (A no-op for a launch without the attribute.)
    // With per-utterance ready counts (isp_align_forward) there is no wait at all: the logits are the only thing the grid in front
    // produces, and a loader takes its utterance's rows as soon as their count is final -- the chains of this kernel start under
    // the last wave of the log-likelihood kernel.
    const bool linked = p.ready != nullptr && p.order == nullptr;
    const bool early = linked || (p.early_lengths != 0 && p.order == nullptr);
    if (!early) asm volatile("griddepcontrol.wait;" ::: "memory");

    // warps: [fillers x2][loaders 2 x 4][helpers 2 x 2][strips of slot 1][strips of slot 0] -- the chains get the highest ids
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int slot, role, s = 0;          // role 0 strip, 1 loader, 2 filler, 3 transposer, 4 mapper + backtrack
    if (warp < 2) { role = 2; slot = warp; }
    else if (warp < 10) { role = 1; slot = (warp - 2) >> 2; s = (warp - 2) & 3; }
    else if (warp < 14) { role = 3 + ((warp - 10) & 1); slot = (warp - 10) >> 1; }
    else { role = 0; slot = warp < 18 ? 1 : 0; s = (warp - 14) & 3; }

    // ---- which utterances, and how the two share the SM ----
    const int c = blockIdx.x;
    int rank0 = 0, rank1 = -1;
    auto members = [&](int cc, int& r0, int& r1) __attribute__((always_inline)) {
        r0 = cc; r1 = -1;
        if (cc >= p.nsingle) { r1 = p.B - 1 - (cc - p.nsingle); if (r1 <= r0) r1 = -1; }
    };
    int b0, b1 = -1, blong;
    uint32_t key0 = 0, key1 = 0, keyl = 0;              // self-ranking: (frames << 17 | tokens - 1 << 9 | index'), clamped lengths
    if (p.order != nullptr) {
        members(c, rank0, rank1);
        if (rank0 >= p.B) return;
        if (slot == 1 && rank1 < 0) return;
        b0 = p.order[rank0];
        if (rank1 >= 0) b1 = p.order[rank1];
        blong = p.order[0];
    } else {
        // Longest first (frames, then tokens, then index) without a sort: every utterance's rank is the number of keys above
        // its own -- B compares for each of B threads, ~1 us at B = 256, where a separate sorting kernel in front of this one
        // costs 6 us of launch and dependency.  Keys are unique (the index is part of them).
        uint32_t* keys = reinterpret_cast<uint32_t*>(smem_raw + kZeroPage + 2 * kHdr);
        const int bp = (p.B + 3) & ~3;
        int* byrank = reinterpret_cast<int*>(keys + bp);                            // utterance of every rank
        int* sel = reinterpret_cast<int*>(smem_raw + kZeroPage + kOffNegInf);      // (the -inf page is written after this)
        for (int t = threadIdx.x; t < bp; t += kThreads) {
            uint32_t key = 0;
            if (t < p.B) {
                const long long n64 = p.mel_len[t], m64 = p.text_len[t];
                const uint32_t nn = uint32_t(n64 < 1 ? 1 : (n64 > p.T1max ? p.T1max : n64)), mm = uint32_t(m64 < 1 ? 1 : (m64 > p.T2max ? p.T2max : m64));
                key = (nn << 17) | ((mm - 1u) << 9) | uint32_t(kRankMax - 1 - t);
            }
            keys[t] = key;
        }
        __syncthreads();
        for (int t = threadIdx.x; t < p.B; t += kThreads) {
            const uint32_t kt = keys[t];
            int r = 0;
            for (int u = 0; u < bp; u += 4) {
                const uint4 ku = *reinterpret_cast<const uint4*>(keys + u);
                r += (ku.x > kt) + (ku.y > kt) + (ku.z > kt) + (ku.w > kt);
            }
            byrank[r] = t;
        }
        __syncthreads();
        members(c, rank0, rank1);
        b0 = byrank[rank0];
        if (rank1 >= 0) b1 = byrank[rank1];
        blong = byrank[0];
        key0 = keys[b0]; key1 = keys[b1 >= 0 ? b1 : b0]; keyl = keys[blong];
        __syncthreads();                                // the key scratch becomes the slots' bodies
        if (slot == 1 && rank1 < 0) return;
    }
    // (the keys hold the clamped lengths; the out-of-contract flag needs the originals, which only strip 0's lane 0 re-reads)
    const bool from_keys = p.order == nullptr;
    const Geo g0 = from_keys ? make_geo(key0 >> 17, ((key0 >> 9) & 0xffu) + 1, p.T1max, p.T2max) : make_geo(p.mel_len[b0], p.text_len[b0], p.T1max, p.T2max);
    Geo g1 = g0;
    if (b1 >= 0) g1 = from_keys ? make_geo(key1 >> 17, ((key1 >> 9) & 0xffu) + 1, p.T1max, p.T2max) : make_geo(p.mel_len[b1], p.text_len[b1], p.T1max, p.T2max);
    const uint32_t body0 = kZeroPage + 2 * kHdr;
    const uint32_t avail = kSmemTotal - body0;
    int nstg;
    uint32_t body_off;                  // this slot's body (fixed part, then ring)
    bool deferred = false;
    {
        const uint32_t sb0 = stage_bytes(g0), sb1 = stage_bytes(g1);
        int both = -1;
        if (b1 >= 0 && g0.fixed + g1.fixed < avail) both = int((avail - g0.fixed - g1.fixed) / (sb0 + sb1)) - 1;
        if (both > kMaxStages) both = kMaxStages;
        if (p.max_stages > 0 && both > p.max_stages) both = p.max_stages;
        if (b1 >= 0 && both >= p.min_pair_stages) {
            nstg = both;
            body_off = slot == 0 ? body0 : body0 + g0.fixed + uint32_t(nstg + 1) * sb0;
        } else {
            deferred = b1 >= 0;
            const Geo& gx = slot == 0 ? g0 : g1;
            nstg = int((avail - gx.fixed) / stage_bytes(gx)) - 1;
            if (nstg > kMaxStages) nstg = kMaxStages;
            if (p.max_stages > 0 && nstg > p.max_stages) nstg = p.max_stages;
            body_off = body0;
        }
    }
    const Geo& g = slot == 0 ? g0 : g1;
    const int b = slot == 0 ? b0 : b1;
    // When every CTA is resident from the start, the launch ends with its longest chain, and whatever the others take from HBM
    // early on, they take from it (bandwidth goes to whoever has the most bytes in flight; queues of requests are latency for
    // everybody).  So the others are paced: an utterance's loaders spread its chunks over the longest chain's expected sweep.
    float pace = 0.0f;
    if (p.linger && p.pace_cycles_per_step > 0.0f) {
        const Geo gl = from_keys ? make_geo(keyl >> 17, ((keyl >> 9) & 0xffu) + 1, p.T1max, p.T2max) : make_geo(p.mel_len[blong], p.text_len[blong], p.T1max, p.T2max);
        float t_long = float(gl.n + 31 + 51 * (gl.ns - 1)) * p.pace_cycles_per_step;
        // (the chains within a fifth of the longest are the ones being protected: they run free)
        if (5 * g.n < 4 * gl.n) {
            if (linked && role == 1) {
                // a CTA that became resident late gets what is left of the longest chain's sweep, but no less than its own
                const float own = float(g.n + 31 + 51 * (g.ns - 1)) * p.pace_cycles_per_step;
                float el = lane == 0 ? cycles_since_origin(p.origin, p.cycles_per_ns) : 0.0f;
                el = __shfl_sync(0xffffffffu, el, 0);
                t_long = fmaxf(t_long - el, own);
            }
            pace = t_long / float(g.nch + 3 * (g.ns - 1));
        }
    }
    const int n = g.n, m = g.m, nch = g.nch;
    const uint32_t hdr_sa = smem_sa + kZeroPage + uint32_t(slot) * kHdr;
    const uint32_t zero_sa = smem_sa;
    const uint32_t body_sa = smem_sa + body_off;
    const uint32_t ring_sa = body_sa + g.fixed;
    const uint32_t prog_sa = hdr_sa + kOffProg, landed_sa = hdr_sa + kOffLanded;
    const uint32_t tdone_sa = hdr_sa + kOffTdone, mdone_sa = hdr_sa + kOffMdone;
    const uint32_t filldone_sa = hdr_sa + kOffFillDone;
    const uint32_t slot0_done_sa = smem_sa + kZeroPage + kOffSlotDone;
#if ISP_MAS2_PROBE
    const bool probe_w = p.probe != nullptr && rank0 == 0 && slot == 0 && s == 0;
#else
    const bool probe_w = false;       // (the phase probe of tools/mas2_probe.py is a build option: tools/ab_build.sh probe isp_mas2.cu -DISP_MAS2_PROBE=1)
#endif

    // the zero page and slot 0's "done" flag belong to the CTA: set up by slot 0's filler before anybody needs them
    if (role == 2 && slot == 0) {
        for (int i = lane; i < kZeroPage / 16; i += 32) reinterpret_cast<uint4*>(smem_raw)[i] = make_uint4(0, 0, 0, 0);
        if (lane == 0) st_volatile_sa(slot0_done_sa, 0);
        fence_proxy_async();                            // the zero page is read by bulk copies (async proxy)
    }
    asm volatile("bar.sync 3, %0;" ::"r"(b1 >= 0 ? kThreads : kSlotThreads) : "memory");      // everybody who is still here
    if (deferred) {
        // the pair does not fit side by side: slot 1 takes the whole SM once slot 0 is through
        if (slot == 1) {
            uint32_t spins = 0;
            while (ld_acquire_sa(slot0_done_sa) == 0) { __nanosleep(500); if (++spins > (1u << 24)) __trap(); }
        }
    }

    // per strip: the utterance's columns, the TMA box / ring row width and the ring
    const int mcols = max(0, min(kStrip, m - s * kStrip));
    const int wb = (mcols + 7) & ~7;
    const uint32_t pitchB = uint32_t(wb) * 4u;
    const uint32_t stageB = uint32_t(kR) * pitchB;
    const uint32_t ringB = uint32_t(nstg) * stageB;
    const uint32_t ring_s = ring_sa + uint32_t(nstg + 1) * uint32_t(kR) * uint32_t(s * kStrip) * 4u;   // strips before this one are full width
    const uint32_t full_s = hdr_sa + kOffFull + uint32_t(s) * kMaxStages * 8u;
    const uint32_t empty_s = hdr_sa + kOffEmpty + uint32_t(s) * kMaxStages * 8u;
    const uint32_t v_sa = body_sa + g.off_v + uint32_t(s) * kVBytes;
    const bool active = s < g.ns;

    // ---- set-up: barriers, counters, exchange arrays, the zero rows in front of row 0 ----
    if (role == 0) {
        if (lane == 0) {
            for (int st = 0; st < nstg; ++st) {
                mbar_init(reinterpret_cast<uint64_t*>(smem_raw + (full_s - smem_sa)) + st, p.tma ? 1 : 32);
                mbar_init(reinterpret_cast<uint64_t*>(smem_raw + (empty_s - smem_sa)) + st, 1);
            }
            st_volatile_sa(prog_sa + 4u * s, 0);
            st_volatile_sa(landed_sa + 4u * s, 0);
            if (s == 0) {
                st_volatile_sa(tdone_sa, 0);
                st_volatile_sa(mdone_sa, 0);
                st_volatile_sa(hdr_sa + kOffPathDone, 0);
                st_volatile_sa(hdr_sa + kOffOnesDone, 0);
            }
            fence_mbar_init();
        }
        if (s == 0 && lane < 16) sts_f32(hdr_sa + kOffNegInf + 4u * lane, -CUDART_INF_F);
        if (active) for (int i = lane; i < kVBytes / 4; i += 32) sts_f32(v_sa + 4u * i, -CUDART_INF_F);
    } else if (role == 2) {
        if (lane == 0) st_volatile_sa(filldone_sa, 0);      // the filler's own flag: it does not wait at the slot's barrier
        __syncwarp();
    } else if (role == 1 && active) {
        // ring rows 0..31 stand for rows -32..-1: zeros (any finite value keeps the accumulators at -inf)
        for (uint32_t i = lane; i < 2u * stageB / 16u; i += 32) *reinterpret_cast<uint4*>(smem_raw + (ring_s - smem_sa) + 16u * i) = make_uint4(0, 0, 0, 0);
        fence_proxy_async();                            // the TMA boxes overwrite them later
    }
    if (early && !linked) asm volatile("griddepcontrol.wait;" ::: "memory");
    // the filler only announces itself: it needs nothing the others set up, and must not wait for a slot that starts late
    if (role == 2) asm volatile("bar.arrive %0, %1;" ::"r"(slot + 1), "r"(kSlotThreads) : "memory");
    else named_sync(slot + 1, kSlotThreads);
    if (probe_w && role == 0 && lane == 0) { p.probe[0] = clock64(); p.probe[15] = gtimer(); }
    if (p.trace != nullptr && role == 0 && s == 0 && lane == 0) { p.trace[4 * b] = gtimer(); p.trace[4 * b + 2] = (long long)blockIdx.x * 256 + slot * 16 + (deferred ? 1 : 0); p.trace[4 * b + 3] = nstg; }

    if (role == 0) {
        if (!active) return;
        StripCtx cx;
        cx.lane = lane; cx.s = s; cx.nstg = nstg; cx.nch = nch;
        cx.pitchB = pitchB; cx.ringB = ringB;
        cx.lane_ring = ring_s + min(uint32_t(lane) * 8u, pitchB - 8u);     // lanes past the box read (and discard) its last columns
        cx.full_s = full_s; cx.empty_s = empty_s; cx.landed_s = landed_sa + 4u * s; cx.prog_sa = prog_sa;
        cx.has_prev = s > 0; cx.has_next = s + 1 < g.ns;
        const int L = s * 32 + lane;
        cx.w_ok = L < g.nlp;
        cx.w_sa = body_sa + g.off_w + uint32_t(L) * 4u;
        cx.w_step = uint32_t(g.nlp) * 4u;
        cx.v_rd = lane == 0 ? hdr_sa + kOffNegInf : v_sa + uint32_t(lane) * kVLane;
        cx.v_wr = v_sa + uint32_t(lane + 1) * kVLane;
        cx.bnd_mine = v_sa + 33 * kVLane;
        cx.bnd_prev = v_sa - kVBytes + 33 * kVLane;
        long long pc_wait[2] = {0, 0};
        if (wb == kStrip) strip_forward<true>(cx, probe_w, pc_wait);
        else strip_forward<false>(cx, probe_w, pc_wait);
        if (probe_w && lane == 0) { p.probe[1] = clock64(); p.probe[4] = pc_wait[0]; p.probe[5] = pc_wait[1]; }
        if (p.probe != nullptr && rank0 == 0 && slot == 0 && s == g.ns - 1 && lane == 0) p.probe[10] = clock64();
        return;
    }

    if (role == 1) {
        // =========================== loader warp of strip s ================================
        if (!active) return;
        const uint64_t pol = policy_evict_first();
        const CUtensorMap* map = &maps.m[wb / 8 - 1];
        const float* src_b = p.logp + int64_t(b) * p.sB + s * kStrip;
        const bool solo = p.tma != 0;
        // Shifted chunk index c' = chunk + 2: stage c' mod nstg; the two virtual chunks in front (the zero rows) take part in the
        // "empty" protocol like real ones, so that stages 0 and 1 are first overwritten when the strip has left them.
        if (solo) {
            if (lane == 0) { mbar_arrive_sa(full_s); mbar_arrive_sa(full_s + 8u); }      // virtual chunks: phase 0 of stages 0, 1
        } else {
            mbar_arrive_sa(full_s); mbar_arrive_sa(full_s + 8u);                         // (these barriers count the 32 lanes)
        }
        __syncwarp();
        if (linked) {
            // the utterance's logits: every frame tile published by the log-likelihood kernel (release / acquire at device scope);
            // the TMA reads them through the async proxy, hence the proxy fence behind the acquire
            uint32_t spins = 0;
            int seen;
            do {
                asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(p.ready + b) : "memory");
                if (seen >= p.ready_need) break;
                __nanosleep(200);
                if (++spins > (1u << 24)) __trap();
            } while (true);
            asm volatile("fence.proxy.async;" ::: "memory");
            __syncwarp();
        }
        if (!solo || lane == 0) {
            int issued = 0, landed = 0;
            int st_i = 2 % nstg, st_l = 2 % nstg;
            uint32_t ph_e = (2 / nstg) & 1 ? 0u : 1u;          // parity of completion (c'/nstg - 1) of empty[st_i]
            uint32_t ph_f = (2 / nstg) & 1;                    // parity of completion (c'/nstg) of full[st_l]
            uint32_t idle = 0;
            const long long t0 = clock64();
            while (landed < nch) {
                bool did = false;
                const bool due = pace == 0.0f || float(clock64() - t0) >= pace * float(issued - 2);      // (the first chunks at once)
                if (issued < nch && due && (issued + 2 < nstg || mbar_test_sa(empty_s + uint32_t(st_i) * 8u, ph_e))) {
                    const uint32_t full_b = full_s + uint32_t(st_i) * 8u;
                    const int r0 = kR * issued;
                    const uint32_t dst = ring_s + uint32_t(st_i) * stageB;
                    const bool fetch = r0 < n;                 // past the last row the strip's tail runs on stale rows
                    const bool twice = st_i == 0;              // stage 0 is mirrored behind the last stage: reads never wrap inside a chunk
                    if (solo) {
                        if (fetch) {
                            mbar_expect_tx_sa(full_b, twice ? 2u * stageB : stageB);
                            tma_load_box(dst, map, s * kStrip, r0, b, full_b, pol);
                            if (twice) tma_load_box(ring_s + ringB, map, s * kStrip, r0, b, full_b, pol);
                        } else {
                            mbar_arrive_sa(full_b);
                        }
                    } else {
                        if (fetch) {
                            const int rows = min(kR, n - r0);
                            for (int r = 0; r < rows; ++r)
                                for (int col = lane; col < mcols; col += 32) {
                                    const float* src = src_b + int64_t(r0 + r) * p.sT1 + col;
                                    cp_async4(dst + uint32_t(r) * pitchB + uint32_t(col) * 4u, src);
                                    if (twice) cp_async4(ring_s + ringB + uint32_t(r) * pitchB + uint32_t(col) * 4u, src);
                                }
                        }
                        cp_async_arrive_noinc_sa(full_b);
                    }
                    ++issued;
                    if (++st_i == nstg) { st_i = 0; ph_e ^= 1u; }
                    did = true;
                }
                if (landed < issued && mbar_test_sa(full_s + uint32_t(st_l) * 8u, ph_f)) {
                    ++landed;
                    if (++st_l == nstg) { st_l = 0; ph_f ^= 1u; }
                    if (lane == 0) st_release_sa(landed_sa + 4u * s, landed);
                    did = true;
                }
                if (!did) {
                    __nanosleep(100);
                    if (++idle > (1u << 24)) __trap();
                }
            }
        }
        return;
    }

    if (role == 2) {
        // =========================== filler warp: zero the dense output ===================
        const size_t cells = size_t(p.T1max) * p.T2max;
        if (p.hard != nullptr && lane == 0 && !(p.dbg & 2048)) {
            char* zbeg = reinterpret_cast<char*>(p.hard + size_t(b) * cells);
            char* zend = zbeg + cells * sizeof(int16_t);
            char* zb = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(zbeg) + 15) & ~uintptr_t(15));
            char* ze = reinterpret_cast<char*>(reinterpret_cast<uintptr_t>(zend) & ~uintptr_t(15));
            if (ze < zb) ze = zb;
            for (char* q2 = zbeg; q2 < zb && q2 < zend; q2 += 2) *reinterpret_cast<int16_t*>(q2) = 0;
            for (char* q2 = ze; q2 < zend; q2 += 2) *reinterpret_cast<int16_t*>(q2) = 0;
            // The copies are spread over a window of time chosen by the host (about the launch's expected duration): a short
            // utterance that filled its 400 KB block during its own short sweep would, together with the other short ones,
            // saturate HBM for the first third of the launch and starve the long chains that decide when it ends.
            const int pieces = int((size_t(ze - zb) + kZeroPage - 1) / kZeroPage);
            const float own = 0.8f * 45.0f * float(n + 31 + 51 * (g.ns - 1));           // most of this utterance's own sweep
            const float window = linked ? p.fill_cycles - cycles_since_origin(p.origin, p.cycles_per_ns) : p.fill_cycles;
            const float rate = float(pieces) / fmaxf(window, own);
            const long long t0 = clock64();
            int issued = 0;
            uint32_t idle = 0;
            char* zp = zb;
            while (issued < pieces) {
                int target = int(rate * float(clock64() - t0)) + 2;
                if (!p.linger && ld_volatile_sa(prog_sa + 4u * uint32_t(g.ns - 1)) >= nch) target = pieces;    // other CTAs are waiting for this SM
                if (p.dbg & 16) target = pieces;
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
        }
        if (lane == 0) st_release_sa(filldone_sa, 1);
        return;
    }

    const uint32_t w_sa = body_sa + g.off_w, rm_sa = body_sa + g.off_rm, maps_sa = body_sa + g.off_maps;
    const int nblk = (n + 31) >> 5;

    if ((p.dbg & 1024) && role >= 3) wait_counter_idle(prog_sa + 4u * uint32_t(g.ns - 1), nch, 1000);     // experiment: helpers after the sweep
    if (role == 3) {
        // =========================== transposer: strip words -> row-major words ===========
        // Lane i of strip s holds, as a stream of 2-bit elements over steps, the bits of columns 2L, 2L+1 (L = 32s + i); the 32
        // elements of rows 32g .. 32g+31 start at bit 64g + 2i of the stream.  A 32 x 32 transpose of 2-bit elements (five
        // butterfly stages) turns them into one 64-bit word per row: bit 2i+c = column 64s + 2i + c.
        uint32_t K[4], rot[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int sh = 8 >> t;                                   // elements: 8, 4, 2, 1
            uint32_t M = 0;                                          // bits of elements e with (e & sh) == 0, period 4 sh bits
            for (int e = 0; e < 16; ++e) if (!(e & sh)) M |= 3u << (2 * e);
            const bool up = (lane & sh) != 0;
            K[t] = up ? ~M : M;
            rot[t] = up ? uint32_t(32 - 2 * sh) : uint32_t(2 * sh);
        }
        int mdone_seen = 0;
        long long tp_m = 0, tp_p = 0;
        for (int gb = 0; gb < nblk; ++gb) {
            // the ring slot's previous block (gb - 8) must have been consumed by the mapper
            if (gb >= kRm / 32 && mdone_seen < gb - kRm / 32 + 1) {
                uint32_t spins = 0;
                long long c0 = 0;
                if (probe_w) c0 = clock64();
                while ((mdone_seen = ld_acquire_sa(mdone_sa)) < gb - kRm / 32 + 1) { __nanosleep(20); if (++spins > (1u << 26)) __trap(); }
                if (probe_w) tp_m += clock64() - c0;
            }
            const int row = 32 * gb + lane;
            const bool row_ok = row >= 1 && row < n;
            const uint32_t dst = rm_sa + uint32_t((row & (kRm - 1)) * kMapWords) * 4u;
            {
                // a strip completes chunk ch only after its left neighbour has completed ch + 2: the last one is the one to wait for
                long long c0 = 0;
                if (probe_w) c0 = clock64();
                wait_counter_idle(prog_sa + 4u * uint32_t(g.ns - 1), min(2 * gb + 4, nch), 20);
                if (probe_w) tp_p += clock64() - c0;
            }
            // the four strips side by side: their butterflies are independent, which hides the shuffle latencies
            uint32_t lo[kMaxNS], hi[kMaxNS];
            const uint32_t shf = uint32_t(2 * lane) & 31u;
            const int wi = 2 * gb + (lane >> 4);
#pragma unroll
            for (int ss = 0; ss < kMaxNS; ++ss) {
                const int L = 32 * ss + lane;
                uint32_t w0 = 0, w1 = 0, w2 = 0;
                if (L < g.nlp) {
                    const uint32_t a = w_sa + (uint32_t(wi) * uint32_t(g.nlp) + uint32_t(L)) * 4u;
                    w0 = lds_u32(a); w1 = lds_u32(a + uint32_t(g.nlp) * 4u); w2 = lds_u32(a + uint32_t(g.nlp) * 8u);
                }
                lo[ss] = __funnelshift_r(w0, w1, shf);
                hi[ss] = __funnelshift_r(w1, w2, shf);
            }
            // stage 16 elements = one word: the lower lane's high word and the upper lane's low word change places
            {
                const bool up = (lane & 16) != 0;
#pragma unroll
                for (int ss = 0; ss < kMaxNS; ++ss) {
                    const uint32_t got = __shfl_xor_sync(0xffffffffu, up ? lo[ss] : hi[ss], 16);
                    if (up) lo[ss] = got; else hi[ss] = got;
                }
            }
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int sh = 8 >> t;
#pragma unroll
                for (int ss = 0; ss < kMaxNS; ++ss) {
                    const uint32_t plo = __shfl_xor_sync(0xffffffffu, lo[ss], sh), phi = __shfl_xor_sync(0xffffffffu, hi[ss], sh);
                    const uint32_t rlo = __funnelshift_l(plo, plo, rot[t]), rhi = __funnelshift_l(phi, phi, rot[t]);
                    lo[ss] = (lo[ss] & K[t]) | (rlo & ~K[t]);
                    hi[ss] = (hi[ss] & K[t]) | (rhi & ~K[t]);
                }
            }
#pragma unroll
            for (int ss = 0; ss < kMaxNS; ++ss) {
                // row 0 has no predecessor; rows past the last do not exist; strips past the last hold no columns
                const bool ok = row_ok && ss < g.ns;
                sts_u64(dst + uint32_t(ss) * 8u, ok ? lo[ss] : 0u, ok ? hi[ss] : 0u);
            }
            __syncwarp();
            if (lane == 0) st_release_sa(tdone_sa, gb + 1);
        }
        if (probe_w && lane == 0) { p.probe[6] = clock64(); p.probe[7] = tp_p; p.probe[8] = tp_m; }
        // ... and then the odd half of the path's ones (scattered 2 B stores: one warp's cost ~2 cycles per row in the load /
        // store unit, two warps' interleave)
        if (p.hard) {
            wait_counter_idle(hdr_sa + kOffPathDone, 1, 100);
            write_ones(p.hard + size_t(b) * p.T1max * p.T2max, body_sa + g.off_path, n, p.T2max, lane, 1);
            __syncwarp();
            if (lane == 0) st_release_sa(hdr_sa + kOffOnesDone, 1);
        }
        return;
    }

    // =========================== mapper, then the backtrack proper =========================
    {
        if (lane == 0) {
            // (the keys hold the clamped lengths; the out-of-contract flag needs the originals -- read here, off the chains' warps)
            const long long n64 = p.mel_len[b], m64 = p.text_len[b];
            p.bad[b] = (unsigned char)(n64 < 1 || n64 > p.T1max || m64 < 1 || m64 > p.T2max);
            if (blockIdx.x == 0 && slot == 0) {
                p.status[0] = -2;                       // isp_mas_status: the per-utterance flags follow, there is no counter
                p.status[1] = p.B;
            }
        }
        const int k = lane & 7, gq = lane >> 3;
        const int nq = (g.g1 + 3) >> 2;
        // plane k of the identity map, word w: bit b = bit k of (32 w + b).  k < 5: a pattern in b; k >= 5: all of bit (k - 5) of w.
        // Branch-free (a switch on k diverges eight ways): 0xffffffff / (2^(2^k) + 1) is the low half of each 2^(k+1)-bit period.
        const uint32_t patk = k < 5 ? (0xffffffffu / ((1u << (1u << (k & 7))) + 1u)) << (1u << (k & 7)) : 0u;
        uint32_t ident = 0;                                     // bit w: word w of plane k starts as ~patk (k >= 5: all ones)
#pragma unroll
        for (int w = 0; w < kMapWords; ++w) ident |= (k >= 5 && ((w >> ((k - 5) & 3)) & 1)) ? (1u << w) : 0u;
        long long mp_t = 0, mp_c = 0;
        for (int qd = 0; qd < nq; ++qd) {
            {
                long long c0 = 0;
                if (probe_w) c0 = clock64();
                wait_counter_idle(tdone_sa, min(4 * qd + 4, nblk), 20);
                if (probe_w) mp_t += clock64() - c0;
            }
            long long c1 = 0;
            if (probe_w) c1 = clock64();
            const int gg = 4 * qd + gq;
            uint32_t P[kMapWords];
#pragma unroll
            for (int w = 0; w < kMapWords; ++w) P[w] = (ident >> w) & 1u ? ~patk : patk;     // the identity map: plane k of column 32w + b
            const uint32_t rows = rm_sa + uint32_t(((32 * gg) & (kRm - 1)) * kMapWords) * 4u;
#pragma unroll 4
            for (int t = 0; t < 32; ++t) {
                const uint4 a0 = lds_v4(rows + uint32_t(t) * 32u), a1 = lds_v4(rows + uint32_t(t) * 32u + 16u);
                const uint32_t A[kMapWords] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
                for (int w = kMapWords - 1; w >= 0; --w) {
                    const uint32_t sft = w > 0 ? __funnelshift_l(P[w - 1], P[w], 1) : (P[0] << 1);
                    P[w] = (A[w] & sft) | (~A[w] & P[w]);
                }
            }
            if (gg < g.g1) {
                const uint32_t dst = maps_sa + uint32_t((gg * kPlanes + k) * kMapWords) * 4u;
#pragma unroll
                for (int w = 0; w < kMapWords; ++w) sts_u32(dst + 4u * w, P[w]);
            }
            __syncwarp();
            if (probe_w) mp_c += clock64() - c1;
            if (lane == 0) st_release_sa(mdone_sa, 4 * qd + 4);
        }
        __syncwarp();
        if (probe_w && lane == 0) { p.probe[2] = clock64(); p.probe[9] = mp_t; p.probe[19] = mp_c; }

        // ---- hops: the path's column at the last row of every group ----
        const uint32_t entry_sa = body_sa + g.off_entry, path_sa = body_sa + g.off_path, start_sa = body_sa + g.off_start;
        int j = m - 1;
        for (int gg = g.g1 - 1; gg >= 0; --gg) {
            if (lane == 0) sts_u32(entry_sa + 4u * gg, uint32_t(j));
            const uint32_t wv = lds_u32(maps_sa + uint32_t((gg * kPlanes + k) * kMapWords + (j >> 5)) * 4u);
            j = int(__ballot_sync(0xffffffffu, (wv >> (j & 31)) & 1u) & 0xffu);
        }
        const int path0 = j;
        __syncwarp();
        if (probe_w && lane == 0) p.probe[11] = clock64();
        // ---- expansion: one lane per group walks its 32 rows (mas.py:22-24).  What the walk passes is final, so it leaves as
        //      it goes: the path's column (to the caller, and to shared memory for the ones below) and the first row of every
        //      token -- the row the path enters the token's column diagonally.
        int16_t* path_g = p.path ? p.path + size_t(b) * p.T1max : nullptr;
        int16_t* hard_b = p.hard ? p.hard + size_t(b) * p.T1max * p.T2max : nullptr;
        for (int gbase = 0; gbase < g.g1; gbase += 32) {
            const int gg = gbase + lane;
            if (gg < g.g1) {
                int jj = int(lds_u32(entry_sa + 4u * gg));
                const int hi_row = min(32 * gg + 31, n - 1), lo_row = max(32 * gg, 1);
                for (int i = hi_row; i >= lo_row; --i) {
                    sts_u16(path_sa + 2u * i, jj);
                    if (path_g) path_g[i] = int16_t(jj);
                    const int L = jj >> 1, u = i + (L & 31);
                    const uint32_t wv = lds_u32(w_sa + (uint32_t(u >> 4) * uint32_t(g.nlp) + uint32_t(L)) * 4u);
                    const uint32_t bit = (wv >> (2 * (u & 15) + (jj & 1))) & 1u;
                    if (bit) sts_u16(start_sa + 2u * jj, i);
                    jj -= int(bit);
                }
                if (gg == 0) {
                    sts_u16(path_sa, jj);
                    if (path_g) path_g[0] = int16_t(jj);
                    sts_u16(start_sa + 2u * jj, 0);
                }
            }
        }
        if (path_g) for (int r = n + lane; r < p.T1max; r += 32) path_g[r] = -1;
        if (lane == 0) sts_u16(start_sa + 2u * m, n);
        __syncwarp();
        if (probe_w && lane == 0) p.probe[12] = clock64();
        // ---- durations (alignment.py:275): the distance between the first rows of neighbouring tokens ----
        if (p.dur) {
            int64_t* d = p.dur + size_t(b) * p.T2max;
            for (int jj = lane; jj < p.T2max; jj += 32) {
                int64_t v = 0;
                if (jj < m && jj >= path0) v = int64_t(lds_s16(start_sa + 2u * (jj + 1)) - lds_s16(start_sa + 2u * jj));
                d[jj] = v;
            }
        }
        // ---- the path's ones in the dense tensor, after the last zero has landed; eight rows per lane and pass ----
        wait_counter_idle(filldone_sa, 1, 20);
        if (probe_w && lane == 0) p.probe[13] = clock64();
        if (hard_b) {
            __syncwarp();
            if (lane == 0) st_release_sa(hdr_sa + kOffPathDone, 1);       // (the fill has landed: both warps may write)
            write_ones(hard_b, path_sa, n, p.T2max, lane, 0);
            wait_counter_idle(hdr_sa + kOffOnesDone, 1, 20);
        }
        if (probe_w && lane == 0) { p.probe[3] = clock64(); p.probe[16] = gtimer(); }
        if (p.trace != nullptr && lane == 0) p.trace[4 * b + 1] = gtimer();
        __syncwarp();
        // the utterance's ready count goes back to zero for the next isp_align_forward on this workspace (every loader of the
        // utterance has long seen it final; the next call's increments come after this grid in stream order)
        if (linked && lane == 0) const_cast<int*>(p.ready)[b] = 0;
        if (slot == 0 && lane == 0) st_release_sa(slot0_done_sa, 1);
    }
}

// ---- plan kernel: utterances by decreasing length (frames, then tokens), the status words cleared -----------------
__global__ void __launch_bounds__(1024, 1)
mas2_plan_kernel(const int64_t* text_len, const int64_t* mel_len, int B, int npow2, int* order, int* status) {
    extern __shared__ unsigned long long keys[];
    const int t = threadIdx.x;
    if (t == 0) { long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g)); reinterpret_cast<long long*>(status)[8 + 18] = g; }
    for (int i = t; i < npow2; i += blockDim.x) {
        unsigned long long key = ~0ull;                 // padding sorts last
        if (i < B) {
            long long n = mel_len[i], m = text_len[i];
            n = n < 0 ? 0 : (n > 0xfffff ? 0xfffff : n);
            m = m < 0 ? 0 : (m > 0xfffff ? 0xfffff : m);
            key = ((unsigned long long)(0xfffff - n) << 44) | ((unsigned long long)(0xfffff - m) << 24) | (unsigned long long)i;
        }
        keys[i] = key;
    }
    __syncthreads();
    for (int k = 2; k <= npow2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = t; i < npow2; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long a = keys[i], bb = keys[ixj];
                    const bool up = (i & k) == 0;
                    if ((a > bb) == up) { keys[i] = bb; keys[ixj] = a; }
                }
            }
            __syncthreads();
        }
    for (int i = t; i < B; i += blockDim.x) order[i] = int(keys[i] & 0xffffffu);
}

__global__ void mas2_identity_order_kernel(int B, int* order, int* status) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B) order[i] = i;
}

}  // namespace mas2

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
using namespace mas2;

static int g2_min_pair_stages = 0;
static int g2_single = 0;
static int g2_fill_us = 0;
static int g2_together = 0;
static int g2_pace = -1;
static int g2_pdl = 0;              // 0: plain launch; 1: programmatic stream serialisation; 2: ... and the lengths may be read early

int mas2_set_option(const char* key, int value, int* prev) {
    if (!strcmp(key, "mas2.min_pair_stages")) { *prev = g2_min_pair_stages; g2_min_pair_stages = value; return 0; }
    if (!strcmp(key, "mas2.single")) { *prev = g2_single; g2_single = value; return 0; }
    if (!strcmp(key, "mas2.fill_us")) { *prev = g2_fill_us; g2_fill_us = value; return 0; }
    if (!strcmp(key, "mas2.together")) { *prev = g2_together; g2_together = value; return 0; }
    if (!strcmp(key, "mas2.pace")) { *prev = g2_pace; g2_pace = value; return 0; }
    if (!strcmp(key, "mas.pdl")) { *prev = g2_pdl; g2_pdl = value; return 0; }
    return -1;
}

// isp_align_forward: does mas2_forward take the ready counts for this shape (the CTAs rank the batch themselves)?
bool mas2_linkable(int B, int T1max, int T2max, int dbg) {
    return mas2_supported(B, T1max, T2max) && B <= kRankMax && T1max < (1 << 14) && !(dbg & 128);
}

bool mas2_supported(int B, int T1max, int T2max) {
    if (T2max > kMaxNS * kStrip || B >= (1 << 24) || T1max > 32000) return false;
    const Geo g = make_geo(T1max, T2max, T1max, T2max);
    const uint32_t body0 = kZeroPage + 2 * kHdr;
    return g.fixed + 8u * stage_bytes(g) <= kSmemTotal - body0;        // at least 7 stages + the mirror for the largest utterance alone
}

// workspace: [0, 256) header (word 0: -2 = "flags follow", word 1: B; probe stamps from byte 64), the order (B ints), the
// per-utterance out-of-contract flags (B bytes), the per-utterance trace (B x 4 int64, debug)
static size_t ws2_off_bad(int B) { return 256 + ((size_t(B) * 4 + 15) & ~size_t(15)); }
static size_t ws2_off_trace(int B) { return ws2_off_bad(B) + ((size_t(B) + 15) & ~size_t(15)); }
size_t mas2_workspace_bytes(int B) { return ws2_off_trace(B) + size_t(B) * 32; }

typedef CUresult (*PFN_encodeTiled2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                     const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                     CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static bool make_maps2(Maps* maps, const float* logp, int64_t sB, int64_t sT1, int B, int T1max, int T2max) {
    static PFN_encodeTiled2 enc = nullptr;
    if (!enc) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess || !ptr) return false;
        enc = reinterpret_cast<PFN_encodeTiled2>(ptr);
    }
    if ((reinterpret_cast<uintptr_t>(logp) & 15) || (sT1 & 3) || (sB & 3)) return false;
    struct Key { const float* p; int64_t sB, sT1; int B, T1, T2; };
    static thread_local Key last = {nullptr, 0, 0, 0, 0, 0};
    static thread_local Maps last_maps;
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

int mas2_forward(const float* logp, int64_t sB, int64_t sT1, const int64_t* text_len, const int64_t* mel_len,
                 int B, int T1max, int T2max, int16_t* attn_hard, int64_t* durations, int16_t* path, void* ws,
                 int no_tma, int ring_rows, int slots, int dbg, cudaStream_t stream, const int* ready, int ready_need) {
    int sm_count = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);

    Params p;
    p.logp = logp; p.sB = sB; p.sT1 = sT1;
    p.text_len = text_len; p.mel_len = mel_len;
    p.B = B; p.T1max = T1max; p.T2max = T2max;
    p.hard = attn_hard; p.dur = durations; p.path = path;
    p.status = reinterpret_cast<int*>(ws);
    p.probe = reinterpret_cast<long long*>(reinterpret_cast<char*>(ws) + 64);
    int* order = reinterpret_cast<int*>(reinterpret_cast<char*>(ws) + 256);
    const bool self_rank = B <= kRankMax && T1max < (1 << 14) && !(dbg & 128);
    p.order = self_rank ? nullptr : order;
    p.bad = reinterpret_cast<unsigned char*>(ws) + ws2_off_bad(B);
    p.trace = (dbg & 64) ? reinterpret_cast<long long*>(reinterpret_cast<char*>(ws) + ws2_off_trace(B)) : nullptr;
    p.dbg = dbg;
    p.max_stages = ring_rows > 0 ? (ring_rows + kR - 1) / kR : 0;
    if (p.max_stages > 0 && p.max_stages < 6) p.max_stages = 6;
    p.min_pair_stages = g2_min_pair_stages > 0 ? g2_min_pair_stages : 8;
    int khz_launch = 1965000;
    {
        // the zero fill's window: what the launch is expected to take -- the longest chain, or its share of the bytes at 80 % of
        // HBM bandwidth when the batch needs several waves of CTAs
        static int khz_of[64] = {0};                     // the attribute query costs milliseconds
        if (dev >= 0 && dev < 64 && khz_of[dev] == 0) { int v = 0; cudaDeviceGetAttribute(&v, cudaDevAttrClockRate, dev); khz_of[dev] = v > 0 ? v : 1965000; }
        const int khz = (dev >= 0 && dev < 64) ? khz_of[dev] : 1965000;
        khz_launch = khz;
        const double chain_us = (double(T1max) + 31.0 + 50.0 * ((T2max + kStrip - 1) / kStrip - 1)) * 40.0 / (khz * 1e-3) + 5.0;
        const double waves = B > 2 * sm_count ? double(B) / (2.0 * sm_count) : 1.0;
        const double bytes_us = double(B) * T1max * T2max * 4.0 / (0.8 * 6.5e6) / waves;
        double win_us = bytes_us;                    // (the kernel stretches it to 80 % of the utterance's own sweep when that is longer)
        (void)chain_us;
        if (g2_fill_us > 0) win_us = g2_fill_us;
        p.fill_cycles = float(win_us * khz * 1e-3);
    }
    if (p.max_stages > 0 && p.min_pair_stages > p.max_stages) p.min_pair_stages = p.max_stages;
    // one utterance per CTA while every utterance still gets its own SM; beyond that the longest keep an SM to themselves
    // for as long as the CTAs fit in one wave, and the rest pair up longest with shortest
    const bool single = slots == 1 || g2_single;
    int npair = single ? 0 : (B <= sm_count ? 0 : (B <= 2 * sm_count ? B - sm_count : B / 2));
    if (slots >= 2 && !g2_single) npair = B / 2;
    p.nsingle = B - 2 * npair;
    const int grid = p.nsingle + npair;
    p.linger = grid <= sm_count ? 1 : 0;
    p.together = g2_together;
    p.pace_cycles_per_step = g2_pace >= 0 ? float(g2_pace) : 56.0f;      // measured at cfg3: 52 .. 60 is flat, 44 and 70 lose 5 %
    Maps maps;
    memset(&maps, 0, sizeof(maps));
    p.tma = (!no_tma && make_maps2(&maps, logp, sB, sT1, B, T1max, T2max)) ? 1 : 0;

    cudaError_t e = cudaFuncSetAttribute(mas2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmemTotal));
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(mas2_kernel)");
    e = cudaFuncSetAttribute(mas2_plan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 8);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(mas2_plan_kernel)");
    if (!self_rank) {
        int npow2 = 1;
        while (npow2 < B) npow2 <<= 1;
        if (npow2 <= 8192) {
            mas2_plan_kernel<<<1, 1024, size_t(npow2) * 8, stream>>>(text_len, mel_len, B, npow2, order, p.status);
        } else {
            mas2_identity_order_kernel<<<(B + 255) / 256, 256, 0, stream>>>(B, order, p.status);
        }
        e = cudaGetLastError();
        if (e != cudaSuccess) return cuda_fail(e, "mas2 plan kernel launch");
    }
    p.early_lengths = g2_pdl >= 2 ? 1 : 0;
    p.ready = ready; p.ready_need = ready_need;
    // (the word behind the ready counts, 8 B aligned, cleared with them by isp_align_forward)
    p.origin = ready ? reinterpret_cast<unsigned long long*>(const_cast<int*>(ready) + ((B + 1) & ~1)) : nullptr;
    p.cycles_per_ns = float(khz_launch) * 1e-6f;
    if (g2_pdl >= 1 || ready != nullptr) {
        // programmatic stream serialisation: the grid may be scheduled while the previous one drains (it calls griddepcontrol.wait)
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(unsigned(grid)); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = kSmemTotal; cfg.stream = stream;
        cudaLaunchAttribute at;
        memset(&at, 0, sizeof(at));
        at.id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at.val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = &at; cfg.numAttrs = 1;
        e = cudaLaunchKernelEx(&cfg, mas2_kernel, maps, p);
        if (e != cudaSuccess) return cuda_fail(e, "mas2_kernel launch (programmatic serialisation)");
    } else {
        mas2_kernel<<<grid, kThreads, kSmemTotal, stream>>>(maps, p);
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "mas2_kernel launch");
    return 0;
}

}  // namespace isp
