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
// kernel is therefore built around the latency of ONE row step:
//
//   * One "slot" = one utterance in flight.  A CTA holds 1 or 2 slots; a slot is NS strip
//     warps (32*C text columns each; C = 4 or 8 columns per lane) + one filler warp.  Strip warps get the lowest warp ids,
//     so the strips of co-resident utterances sit on different SM sub-partitions and never
//     compete for an issue port.
//   * Inside a strip warp the wavefront is skewed across LANES: lane l owns C consecutive
//     columns and, at step t, works on row t - l.  The only cross-lane value a row needs
//     (Q[i-1][first column - 1]) was produced by the left neighbour two steps earlier, so
//     its shuffle is issued one step ahead and its latency never sits on the chain.  What
//     is left on the chain per step is FMNMX -> FADD; per cell the step costs FSET + FMNMX
//     (ALU pipe) and FFMA + FADD (FMA pipe): the backpointer bit is accumulated as a float
//     (FSET gives 1.0/0.0, an FFMA tree packs C of them into a mantissa).  A row step is
//     issue-bound (measured: 37 cycles at C=4, 73 at C=8, tools/ubench/chain.cu), so C=4 is used
//     while every strip warp still gets an SM sub-partition to itself, C=8 beyond that.
//   * Strips of an utterance wider than one warp form a second, coarser wavefront: strip s runs
//     >= 31 + 8 rows behind strip s-1 and takes one boundary value per row from a small
//     shared-memory ring (release/acquire progress counters, both directions).
//   * Each strip warp feeds itself: it owns a ring of logit rows in shared memory and, every
//     8 steps, re-arms the stage its last lane has just left with ONE tiled TMA copy per
//     128-column segment (cp.async.bulk.tensor.3d, box = 8 rows x pitch columns), completion
//     on an mbarrier.  No producer warp, no "empty" barriers.  The box width is the smem row
//     pitch; it comes from a small set (36/68/100/132 floats for C=8, 40/72/104/136 for C=4)
//     chosen so that the skewed 16 B reads are bank-conflict-free and ragged utterances fetch
//     little more than their valid T2_b columns.  Rows past T1_b are never requested.
//   * Backpointers cost one bit per cell, in shared memory when the slot's bits fit, else in
//     the caller's workspace (L2-resident).  C=8: one byte per lane and row, i.e. row-major bits
//     (column j at bit j%32 of word j/32).  C=4: one byte per lane and ROW PAIR (low nibble =
//     even row), unpacked by the backtrack when it fetches a window.
//   * filler warp: zero-fills the utterance's dense int16 block and duration row with 16 B
//     streaming stores while the DP runs.
//   * backtrack (strip warp 0): 32 rows per block.  Every lane fetches the 32-column window
//     its row can touch (the path moves at most one column per row); the windows are
//     broadcast through shared memory and the dependent chain runs on a one-hot position
//     register, two ALU levels per row:  R' = (R & ~A) | ((R >> 1) & (A >> 1)); each lane then
//     picks its own row's position out of the 32 chain values with a 5-level select tree (a
//     store per row inside the chain costs 3x more, tools/ubench/chain.cu).  Then all 32 lanes
//     write their row's 1 and the durations of the tokens that start in the block (ballot +
//     clz), so nothing is re-read to build durations.
//
// Bit-exactness: each cell does exactly the reference's one fp32 add on top of an exact
// max; the comparison is the reference's `>=` (ties and -inf >= -inf take the diagonal).

#include <cuda.h>
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "common.cuh"
#include "isp_internal.h"

namespace isp {

constexpr int kR = 8;                 // rows per ring stage (= steps per unrolled chunk = TMA box height)
constexpr int kMaxStages = 16;
constexpr int kMaxStrips = 8;         // ISP_MAS_MAX_T2 / 128
constexpr int kMaxSlots = 2;
constexpr int kMaxThreads = 320;      // slots * (strips + 1) warps <= 10
constexpr int kBnd = 256;             // rows in a strip-boundary ring (power of two, multiple of kR)
constexpr int kMinRing = 48;          // 31 (lane skew) + kR (stage granularity) + kR (read-ahead), rounded up
constexpr int kSlotHdr = 2048;        // barriers, counters, backtrack windows
constexpr int kSeg = 128;             // columns per ring segment (one TMA box wide)
constexpr int kNumBox = 4;            // box widths: base + 32 i floats, base = 36 (C = 8) or 40 (C = 4)

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
    int* status;           // count of utterances with out-of-contract lengths
    long long* probe;      // clock64 stamps of utterance 0 (tools/mas_probe.py)
    int ns;                // strip warps per utterance
    int slots;             // utterances per CTA
    int full_floats;       // allocated ring-row floats of a full strip (all segments)
    int last_floats[2];    // allocated pitch of segment 0 / 1 of the last strip
    int ring_rows;         // rows per strip ring (multiple of kR)
    int bits_pitch;        // bytes per row (C = 8) or row pair (C = 4) of bits, = ns * 32
    int slot_bytes;        // shared memory per slot
    int tma;               // 1: tensor maps are valid -> tiled TMA copies
    int dbg;               // debug/profiling switches (mas.dbg)
};

struct MasMaps { CUtensorMap m[kNumBox]; };

// smallest box (index, width in floats) that covers `cols` (1..128) columns
template <int C> __host__ __device__ inline int box_index(int cols) {
    const int base = C == 8 ? 36 : 40;
    return cols <= base ? 0 : (cols - base + 31) / 32;
}
template <int C> __host__ __device__ inline int box_width(int idx) { return (C == 8 ? 36 : 40) + 32 * idx; }

// ---- zero-fill of [p, p+bytes) with 16 B streaming stores (any alignment) ------------
ISP_DEVINL void warp_zero_fill(char* p, size_t bytes, int lane) {
    size_t head = (16 - (reinterpret_cast<uintptr_t>(p) & 15)) & 15;
    if (head > bytes) head = bytes;
    for (size_t k = lane * 2; k < head; k += 64) *reinterpret_cast<int16_t*>(p + k) = 0;
    char* body = p + head;
    size_t nvec = (bytes - head) >> 4;
    const uint4 z = make_uint4(0, 0, 0, 0);
    size_t v = lane;
    for (; v + 96 < nvec; v += 128) {
        st_cs_v4(body + (v << 4), z);
        st_cs_v4(body + ((v + 32) << 4), z);
        st_cs_v4(body + ((v + 64) << 4), z);
        st_cs_v4(body + ((v + 96) << 4), z);
    }
    for (; v < nvec; v += 32) st_cs_v4(body + (v << 4), z);
    char* tail = body + (nvec << 4);
    size_t tbytes = bytes - head - (nvec << 4);
    for (size_t k = lane * 2; k < tbytes; k += 64) *reinterpret_cast<int16_t*>(tail + k) = 0;
}

ISP_DEVINL void tma_load_box(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;"
        ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
ISP_DEVINL float set_ge(float a, float b) {   // 1.0f if a >= b (false on NaN), else 0.0f: one FSET
    float d;
    asm("set.ge.f32.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
    return d;
}
template <int C> ISP_DEVINL void lds_row(float (&x)[C], uint32_t saddr) {
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x[0]), "=f"(x[1]), "=f"(x[2]), "=f"(x[3]) : "r"(saddr));
    if (C == 8) asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+16];" : "=f"(x[4 % C]), "=f"(x[5 % C]), "=f"(x[6 % C]), "=f"(x[7 % C]) : "r"(saddr));
}
ISP_DEVINL void sts_u8(uint32_t saddr, uint32_t v) {
    asm volatile("st.shared.u8 [%0], %1;" ::"r"(saddr), "r"(v) : "memory");
}

// one backtrack row on a one-hot position: stay where A is 0, move one column down where A is 1.
// Written as two LOP3 levels so that the dependent chain is 2 ALU ops per row, not 3.
ISP_DEVINL uint32_t bt_step(uint32_t R, uint32_t A, uint32_t A1) {
    uint32_t P, Rs = R >> 1, out;
    asm("lop3.b32 %0, %1, %2, 0, 0x30;" : "=r"(P) : "r"(R), "r"(A));              // R & ~A
    asm("lop3.b32 %0, %1, %2, %3, 0xf8;" : "=r"(out) : "r"(P), "r"(Rs), "r"(A1));   // P | (Rs & A1)
    return out;
}
// low nibbles of the 4 bytes of w (after >> shift) packed into 16 bits
ISP_DEVINL uint32_t pack_nibbles(uint32_t w, int shift) {
    uint32_t x = (w >> shift) & 0x0f0f0f0fu;
    x = (x | (x >> 4)) & 0x00ff00ffu;
    return (x | (x >> 8)) & 0xffffu;
}

// ---- one DP row for one lane --------------------------------------------------------
// q[] holds row i-1 on entry and row i on exit.  Returns the lane's C backpointer bits in
// the low bits (bit c set <=> predecessor of column base+c is the diagonal).
template <int C> ISP_DEVINL uint32_t dp_row(float (&q)[C], const float (&x)[C], float left) {
    float s[C];
#pragma unroll
    for (int c = C - 1; c >= 1; --c) {
        s[c] = set_ge(q[c - 1], q[c]);                // mas.py:17 -- ties take j-1
        q[c] = x[c] + fmaxf(q[c - 1], q[c]);          // mas.py:14 -- one fp32 add per cell
    }
    s[0] = set_ge(left, q[0]);                        // left == NaN at global column 0: false, keeps q[0]
    q[0] = x[0] + fmaxf(left, q[0]);
    const float t0 = fmaf(s[1], 2.0f, s[0]), t1 = fmaf(s[3], 2.0f, s[2]);
    float v = fmaf(t1, 4.0f, t0);
    if (C == 8) {
        const float t2 = fmaf(s[5 % C], 2.0f, s[4 % C]), t3 = fmaf(s[7 % C], 2.0f, s[6 % C]);
        v = fmaf(fmaf(t3, 4.0f, t2), 16.0f, v);
    }
    return __float_as_uint(v + 8388608.0f);           // integer 0..2^C-1 in the low mantissa bits
}

template <int C, bool BITS_SMEM, bool MULTI>
__global__ void __launch_bounds__(kMaxThreads, 1)
mas_kernel(const __grid_constant__ MasMaps maps, const MasParams p) {
    constexpr int W = 32 * C;            // columns per strip
    constexpr int NSEG = W / kSeg;       // ring segments per strip (1 or 2)
    constexpr int LPS = kSeg / C;        // lanes per segment
    extern __shared__ __align__(128) unsigned char smem_raw[];

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int ns = p.ns;
    const int nstrip_total = p.slots * ns;
    const bool is_strip = warp < nstrip_total;
    const int slot = is_strip ? warp / ns : warp - nstrip_total;
    const int s = is_strip ? warp - slot * ns : 0;
    const int b = blockIdx.x * p.slots + slot;
    if (b >= p.B) return;                               // the whole slot leaves together
    const uint32_t slot_threads = 32u * (ns + 1);

    // ---- carve the slot's shared memory -------------------------------------------------
    unsigned char* sm = smem_raw + size_t(slot) * p.slot_bytes;
    uint64_t* full_all = reinterpret_cast<uint64_t*>(sm);                       // [kMaxStrips][kMaxStages]
    int* prog = reinterpret_cast<int*>(sm + 1024);                              // [kMaxStrips] rows finished by lane 31
    int* cons = prog + kMaxStrips;                                              // [kMaxStrips] rows entered by lane 0
    uint32_t* winbuf = reinterpret_cast<uint32_t*>(sm + 1024 + 64);             // [32][2] backtrack windows (A, A >> 1)
    float* bnd = reinterpret_cast<float*>(sm + kSlotHdr);                       // [ns-1][kBnd]
    size_t off = kSlotHdr + (MULTI ? sizeof(float) * size_t(ns - 1) * kBnd : 0);
    float* ring_all = reinterpret_cast<float*>(sm + off);
    const int last_total = p.last_floats[0] + p.last_floats[1];
    off += sizeof(float) * (size_t(p.ring_rows) * (size_t(ns - 1) * p.full_floats + last_total) + kSeg);  // + kSeg: lanes past the last valid column read on
    unsigned char* bits_base;
    if (BITS_SMEM) bits_base = sm + off;
    else bits_base = reinterpret_cast<unsigned char*>(p.bits_ws + size_t(b) * p.bits_stride);

    // ---- lengths (read on device; clamped for memory safety, reported via status) ------
    const long long n64 = p.mel_len[b], m64 = p.text_len[b];
    const bool bad = n64 < 1 || n64 > p.T1max || m64 < 1 || m64 > p.T2max;
    const int n = int(n64 < 1 ? 1 : (n64 > p.T1max ? p.T1max : n64));   // frames
    const int m = int(m64 < 1 ? 1 : (m64 > p.T2max ? p.T2max : m64));   // tokens
    const int ns_active = (m + W - 1) / W;
    const int nstg = p.ring_rows / kR;
    const bool probe_w = p.probe != nullptr && b == 0 && is_strip && s == 0;   // warp-uniform
    const bool probe = probe_w && lane == 0;
    long long pc_issue = 0, pc_wait = 0;

    if (is_strip && lane == 0) {
        if (s == 0 && bad) atomicAdd(p.status, 1);
        for (int st = 0; st < nstg; ++st) mbar_init(&full_all[s * kMaxStages + st], 1);
        prog[s] = 0;
        cons[s] = 0;
        fence_mbar_init();
    }
    asm volatile("bar.sync %0, %1;" ::"r"(slot + 1), "r"(slot_threads) : "memory");
    if (probe) p.probe[0] = clock64();

    if (is_strip) {
        if (s < ns_active) {
            // =========================== strip warp: forward DP ===========================
            const int mcols = min(W, m - s * W);                         // valid columns of this strip
            // the strip's ring is NSEG segments of <= 128 columns, each filled by its own TMA box
            const int cols0 = min(kSeg, mcols), cols1 = mcols - cols0;
            const int bi0 = box_index<C>(cols0), bi1 = box_index<C>(max(cols1, 1));
            const int P0 = box_width<C>(bi0), P1 = (NSEG > 1 && cols1 > 0) ? box_width<C>(bi1) : 0;   // row pitch (floats)
            float* ring0 = ring_all + size_t(s) * p.ring_rows * p.full_floats;
            float* ring1 = ring0 + size_t(p.ring_rows) * (s == ns - 1 ? p.last_floats[0] : p.full_floats / NSEG);
            uint64_t* full = full_all + s * kMaxStages;
            const int nchunks = (n + kR - 1) / kR;
            const uint64_t pol = policy_evict_first();
            const bool has_prev = MULTI && s > 0;
            const bool has_next = MULTI && s + 1 < ns_active;
            float* bnd_mine = bnd + size_t(s) * kBnd;
            const float* bnd_prev = bnd + size_t(s > 0 ? s - 1 : 0) * kBnd;
            const uint32_t chunk_tx = uint32_t(kR) * uint32_t(P0 + P1) * 4u;         // a box always lands whole

            // (re)fill the next ring stage with rows [c*kR, c*kR + kR) -- warp-collective; chunks are issued in
            // order, so the stage index just cycles (no runtime division on the chain)
            int issue_st = 0;
            auto issue_chunk = [&](int c) {
                const int st = issue_st;
                issue_st = issue_st + 1 == nstg ? 0 : issue_st + 1;
                const int r0 = c * kR;
                __syncwarp();                          // every lane is done reading this stage
                if (p.tma) {
                    if (lane == 0) {
                        mbar_arrive_expect_tx(&full[st], chunk_tx);
                        tma_load_box(ring0 + size_t(st) * kR * P0, &maps.m[bi0], s * W, r0, b, &full[st], pol);
                        if (NSEG > 1 && cols1 > 0) tma_load_box(ring1 + size_t(st) * kR * P1, &maps.m[bi1], s * W + kSeg, r0, b, &full[st], pol);
                    }
                } else {
                    // tensor maps unavailable (unaligned base or strides): coalesced 4 B loads through registers
                    const int rows = min(kR, n - r0);
                    const float* src = p.logp + int64_t(b) * p.sB + s * W + int64_t(r0) * p.sT1;
                    for (int r = 0; r < rows; ++r) {
                        const float* g = src + int64_t(r) * p.sT1;
                        for (int c2 = lane; c2 < cols0; c2 += 32) ring0[(size_t(st) * kR + r) * P0 + c2] = __ldg(g + c2);
                        if (NSEG > 1) for (int c2 = lane; c2 < cols1; c2 += 32) ring1[(size_t(st) * kR + r) * P1 + c2] = __ldg(g + kSeg + c2);
                    }
                    __syncwarp();
                }
            };
            int wait_st = 0;
            uint32_t wait_ph = 0;
            auto wait_chunk = [&]() {                  // chunks are waited for in order
                if (p.tma) mbar_wait(&full[wait_st], wait_ph);
                if (++wait_st == nstg) { wait_st = 0; wait_ph ^= 1u; }
            };

            for (int c = 0; c < nstg && c < nchunks; ++c) issue_chunk(c);

            // lanes [0, LPS) read segment 0, the rest segment 1 (a lane past the valid columns reads stale rows)
            const bool in1 = NSEG > 1 && lane >= LPS && cols1 > 0;
            const uint32_t ring_sa = smem_u32(in1 ? ring1 : ring0);
            const uint32_t pitchB = uint32_t(in1 ? P1 : P0) * 4u;
            const uint32_t ringB = uint32_t(p.ring_rows) * pitchB;
            // byte offset of this lane's columns in the ring row it reads next (row t+1-lane, one step ahead)
            uint32_t xoff = uint32_t((p.ring_rows - lane) % p.ring_rows) * pitchB + uint32_t(lane % LPS) * (C * 4);
            const float qnan = __int_as_float(0x7fffffff);
            float q[C], xc[C];
#pragma unroll
            for (int c = 0; c < C; ++c) q[c] = -CUDART_INF_F;
            float left_cur = qnan;
            const int gcol0 = s * W + lane * C;
            const uint32_t bpB = uint32_t(p.bits_pitch);
            // C = 8: this lane's bits byte for row (t0 - lane) + k is at bits_* + k * bpB;
            // C = 4: for row pair ((t0 - lane) >> 1) + (k >> 1), written on the pair's odd row.  Advanced once per chunk.
            uint32_t bits_sa = 0;
            unsigned char* bits_g = nullptr;
            {
                const int64_t first = C == 8 ? -int64_t(lane) : -int64_t((lane + 1) >> 1);     // (0 - lane) >> 1, arithmetic
                if (BITS_SMEM) bits_sa = smem_u32(bits_base) + uint32_t(s * 32 + lane) + uint32_t(int32_t(first)) * bpB;
                else bits_g = bits_base + (s * 32 + lane) + first * int64_t(bpB);
            }
            uint32_t nib_lo = 0;                           // C = 4: the even row's nibble, waiting for its odd row
            const bool lane_odd = (lane & 1) != 0;

            wait_chunk();
            lds_row<C>(xc, ring_sa + xoff);
            xoff += pitchB; if (xoff >= ringB) xoff -= ringB;

            const int nsteps = n + 31;
            for (int t0 = 0; t0 < nsteps; t0 += kR) {
                const int ch = t0 / kR;
                long long c0 = 0, c1 = 0, c2 = 0;
                if (probe_w) c0 = clock64();
                {   // the stage lane 31 left during the previous chunk is free: re-arm it
                    const int f = (t0 - 31 >= 0 ? (t0 - 31) / kR : -1) - 1;
                    if (f >= 0 && f + nstg < nchunks) issue_chunk(f + nstg);
                }
                if (probe_w) c1 = clock64();
                if (ch + 1 < nchunks) wait_chunk();      // the read-ahead of this chunk's last step lands there
                if (probe_w) { c2 = clock64(); pc_issue += c1 - c0; pc_wait += c2 - c1; }
                if (MULTI) {
                    if (lane == 0) {
                        uint32_t spins = 0;
                        if (has_prev) {
                            st_release_shared(&cons[s], t0);
                            const int need = min(t0 + kR, n);          // boundary rows this chunk consumes
                            while (ld_acquire_shared(&prog[s - 1]) < need) { if (++spins > (1u << 26)) __trap(); }
                        }
                        if (has_next) {
                            const int need = t0 - 31 + kR - kBnd + 1;  // do not lap the reader of our boundary ring
                            while (ld_acquire_shared(&cons[s + 1]) < need) { if (++spins > (1u << 26)) __trap(); }
                        }
                    }
                    __syncwarp();
                }
                if (t0 >= 32 && t0 + kR <= n) {
                    // ---- steady state: every lane has a valid row >= 1 for all kR steps ----
#pragma unroll
                    for (int k = 0; k < kR; ++k) {
                        const float nxt = __shfl_up_sync(0xffffffffu, q[C - 1], 1);
                        float xn[C];
                        lds_row<C>(xn, ring_sa + xoff);
                        xoff += pitchB; if (xoff >= ringB) xoff -= ringB;
                        float bv = qnan;
                        if (has_prev) bv = bnd_prev[(t0 + k) & (kBnd - 1)];
                        const uint32_t bits = dp_row<C>(q, xc, left_cur);
                        if (C == 8) {
                            if (BITS_SMEM) sts_u8(bits_sa + uint32_t(k) * bpB, bits);
                            else bits_g[int64_t(k) * bpB] = static_cast<unsigned char>(bits);
                        } else {
                            // row parity of this lane at step k is (k + lane) & 1 (t0 is a multiple of 8)
                            const uint32_t byte = nib_lo | (bits << 4);
                            if (((k & 1) != 0) != lane_odd) {
                                if (BITS_SMEM) sts_u8(bits_sa + uint32_t(k >> 1) * bpB, byte);
                                else bits_g[int64_t(k >> 1) * bpB] = static_cast<unsigned char>(byte);
                            }
                            nib_lo = bits & 15u;
                        }
                        if (has_next && lane == 31) bnd_mine[(t0 + k - 31) & (kBnd - 1)] = q[C - 1];
                        left_cur = lane == 0 ? bv : nxt;
#pragma unroll
                        for (int c = 0; c < C; ++c) xc[c] = xn[c];
                    }
                } else {
                    // ---- head and tail: some lanes are before row 0 or past row n-1 ----
                    for (int k = 0; k < kR; ++k) {
                        const int t = t0 + k;
                        const int r = t - lane;
                        const float nxt = __shfl_up_sync(0xffffffffu, q[C - 1], 1);
                        float xn[C];
                        lds_row<C>(xn, ring_sa + xoff);
                        xoff += pitchB; if (xoff >= ringB) xoff -= ringB;
                        float bv = qnan;
                        if (has_prev) bv = bnd_prev[t & (kBnd - 1)];
                        const uint32_t bits = dp_row<C>(q, xc, left_cur);
                        if (r == 0) {
                            // row 0: Q[0][0] = x[0][0], Q[0][j>0] = -inf   (mas.py:11)
#pragma unroll
                            for (int c = 0; c < C; ++c) q[c] = (gcol0 + c == 0) ? xc[c] : -CUDART_INF_F;
                        }
                        if (C == 8) {
                            if (r >= 1 && r < n) {
                                if (BITS_SMEM) sts_u8(bits_sa + uint32_t(k) * bpB, bits);
                                else bits_g[int64_t(k) * bpB] = static_cast<unsigned char>(bits);
                            }
                        } else if (r >= 0 && r < n) {
                            // odd row: the pair is complete; even last row: flush the half pair
                            const uint32_t byte = (r & 1) ? (nib_lo | (bits << 4)) : (bits & 15u);
                            if ((r & 1) || r == n - 1) {
                                const int po = (k + (lane & 1)) >> 1;          // pair of row r, relative to the chunk's base pair
                                if (BITS_SMEM) sts_u8(bits_sa + uint32_t(po) * bpB, byte);
                                else bits_g[int64_t(po) * bpB] = static_cast<unsigned char>(byte);
                            }
                            nib_lo = bits & 15u;
                        }
                        if (has_next && lane == 31 && r >= 0 && r < n) bnd_mine[r & (kBnd - 1)] = q[C - 1];
                        left_cur = lane == 0 ? bv : nxt;
#pragma unroll
                        for (int c = 0; c < C; ++c) xc[c] = xn[c];
                    }
                }
                if (BITS_SMEM) bits_sa += (C == 8 ? kR : kR / 2) * bpB; else bits_g += int64_t(C == 8 ? kR : kR / 2) * bpB;
                if (has_next) {
                    const int done = min(t0 + kR - 31, n);             // rows lane 31 has finished
                    __syncwarp();
                    if (done > 0 && lane == 31) st_release_shared(&prog[s], done);
                }
            }
        }
    } else {
        // =========================== filler warp: zero the outputs =====================
        const size_t cells = size_t(p.T1max) * p.T2max;
        if (!(p.dbg & 4)) warp_zero_fill(reinterpret_cast<char*>(p.hard + size_t(b) * cells), cells * sizeof(int16_t), lane);
        if (p.dur) {
            int64_t* d = p.dur + size_t(b) * p.T2max;
            for (int j = lane; j < p.T2max; j += 32) d[j] = 0;
        }
    }

    // bits (shared or global) and the zero-filled outputs become visible to the slot's warp 0
    if (probe) { p.probe[1] = clock64(); p.probe[4] = pc_issue; p.probe[5] = pc_wait; }
    if (!BITS_SMEM) __threadfence_block();
    asm volatile("bar.sync %0, %1;" ::"r"(slot + 1), "r"(slot_threads) : "memory");
    if (!is_strip || s != 0) return;
    if (probe) p.probe[2] = clock64();

    // =============================== backtrack (strip warp 0) ==========================
    // mas.py:20-24.  Rows i0, i0-1, ..., i0-31 per block; lane t owns row i0-t.
    int16_t* hard_b = p.hard + size_t(b) * p.T1max * p.T2max;
    int64_t* dur_b = p.dur ? p.dur + size_t(b) * p.T2max : nullptr;
    const uint32_t* bits_w = reinterpret_cast<const uint32_t*>(bits_base);
    const int wpr = p.bits_pitch >> 2;      // words per row (C = 8) / row pair (C = 4) of bits
    // the 32 backpointer bits of columns [32 q, 32 q + 32) of row `row`
    auto bits_word = [&](int row, int qq) -> uint32_t {
        if (C == 8) return bits_w[size_t(row) * wpr + qq];
        const uint2 v = *reinterpret_cast<const uint2*>(bits_w + size_t(row >> 1) * wpr + 2 * qq);
        const int sh = (row & 1) * 4;
        return pack_nibbles(v.x, sh) | (pack_nibbles(v.y, sh) << 16);
    };
    int j = m - 1;           // token index on row i0
    int last_start = n;      // first row of token j+1 (exclusive end of token j)
    const uint32_t lt_mask = (1u << lane) - 1u;
    // candidate words of the NEXT block's window are fetched before the chain runs (j moves <= 32)
    uint32_t w0 = 0, w1 = 0, w2 = 0;
    int qw_pref = j >> 5;
    {
        const int row = n - 1 - lane;
        if (row >= 1) {
            w0 = bits_word(row, qw_pref);
            if (qw_pref > 0) w1 = bits_word(row, qw_pref - 1);
        }
    }
    for (int i0 = n - 1; i0 >= 0; i0 -= 32) {
        const int row = i0 - lane;
        // window of this lane's row: bit k <-> column j-31+k  (the path stays within it)
        const int qw = j >> 5;
        const uint32_t hi = qw == qw_pref ? w0 : w1;
        const uint32_t lo = qw == qw_pref ? w1 : w2;
        const uint32_t win = row >= 1 ? __funnelshift_rc(lo, hi, (j & 31) + 1) : 0u;   // row 0 has no predecessor
        *reinterpret_cast<uint2*>(winbuf + 2 * lane) = make_uint2(win, win >> 1);
        __syncwarp();
        // prefetch for the block below: its j is in [j-32, j]  ->  word index in {qw, qw-1, qw-2}
        {
            const int nrow = row - 32;
            w0 = w1 = w2 = 0u;
            if (nrow >= 1) {
                w0 = bits_word(nrow, qw);
                if (qw > 0) w1 = bits_word(nrow, qw - 1);
                if (qw > 1) w2 = bits_word(nrow, qw - 2);
            }
            qw_pref = qw;
        }
        uint32_t Rh[32];
        uint32_t R = 0x80000000u;    // one-hot position inside the window; bit 31 <-> column j
#pragma unroll
        for (int t = 0; t < 32; t += 2) {
            const uint4 a = *reinterpret_cast<const uint4*>(winbuf + 2 * t);    // A_t, A_t >> 1, A_t+1, A_t+1 >> 1
            Rh[t] = R;
            R = bt_step(R, a.x, a.y);
            Rh[t + 1] = R;
            R = bt_step(R, a.z, a.w);
        }
        // this lane's row: position after `lane` rows -- 5-level select tree over the chain values
#pragma unroll
        for (int lvl = 0; lvl < 5; ++lvl) {
            const bool bit = (lane >> lvl) & 1;
#pragma unroll
            for (int i = 0; i < (16 >> lvl); ++i) Rh[i] = bit ? Rh[2 * i + 1] : Rh[2 * i];
        }
        const uint32_t myR = Rh[0];
        const int col = j - __clz(myR);
        const uint32_t dec = __ballot_sync(0xffffffffu, (myR & win) != 0u);   // bit t: path leaves row i0-t diagonally
        if (row >= 0 && !(p.dbg & 1)) hard_b[size_t(row) * p.T2max + col] = 1;
        // a token starts on this row if the path leaves it diagonally, or on row 0
        const bool starts = row >= 0 && (((dec >> lane) & 1u) || row == 0);
        const uint32_t smask = __ballot_sync(0xffffffffu, starts);
        if (starts && dur_b && !(p.dbg & 2)) {
            const uint32_t lower = smask & lt_mask;      // starts of later tokens in this block
            const int next_start = lower ? i0 - (31 - __clz(lower)) : last_start;
            dur_b[col] = int64_t(next_start - row);
        }
        if (smask) last_start = i0 - (31 - __clz(smask));
        j -= __popc(dec);                                   // rows < 1 contribute no bits
    }
    if (probe) p.probe[3] = clock64();
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------

static int g_opt_cols_per_lane = 0;
static int g_opt_ring_rows = 0;
static int g_opt_slots = 0;
static int g_opt_dbg = 0;

int mas_set_option(const char* key, int value, int* prev) {
    if (!strcmp(key, "mas.cols_per_lane")) { *prev = g_opt_cols_per_lane; g_opt_cols_per_lane = value; return 0; }
    if (!strcmp(key, "mas.ring_rows")) { *prev = g_opt_ring_rows; g_opt_ring_rows = value; return 0; }
    if (!strcmp(key, "mas.slots")) { *prev = g_opt_slots; g_opt_slots = value; return 0; }
    if (!strcmp(key, "mas.dbg")) { *prev = g_opt_dbg; g_opt_dbg = value; return 0; }
    return -1;
}

struct MasPlan {
    int C, ns, slots, full_floats, last_floats[2], ring_rows, bits_pitch;
    bool bits_smem;
    size_t slot_bytes;
    size_t bits_ws_words;  // per utterance, when !bits_smem
};

static size_t slot_bytes_for(int ns, int strip_floats_total, int ring_rows, size_t bits_bytes) {
    size_t off = kSlotHdr + (ns > 1 ? sizeof(float) * size_t(ns - 1) * kBnd : 0);
    off += sizeof(float) * (size_t(ring_rows) * strip_floats_total + kSeg);
    off += bits_bytes;
    return (off + 127) & ~size_t(127);
}

template <int C> static void ring_geometry(int T2max, int ns, MasPlan* pl) {
    const int W = 32 * C;
    pl->full_floats = (W / kSeg) * box_width<C>(kNumBox - 1);
    const int last_cols = T2max - (ns - 1) * W;
    const int c0 = last_cols < kSeg ? last_cols : kSeg, c1 = last_cols - c0;
    pl->last_floats[0] = box_width<C>(box_index<C>(c0));
    pl->last_floats[1] = c1 > 0 ? box_width<C>(box_index<C>(c1)) : 0;
}

static int mas_plan(int B, int T1max, int T2max, MasPlan* pl) {
    int sm_count = 148;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
    const size_t smem_limit = 227 * 1024;
    // C = 4 (128-column strips) halves the row step; use it while every strip warp still gets an SM
    // sub-partition to itself, C = 8 (256-column strips) in the throughput regime
    int C = g_opt_cols_per_lane;
    if (C != 4 && C != 8) C = (int64_t(B) * ((T2max + 127) / 128) <= 4 * int64_t(sm_count)) ? 4 : 8;
    if (C == 4 && (T2max + 127) / 128 > kMaxStrips) C = 8;
    const int W = 32 * C;
    const int ns = (T2max + W - 1) / W;
    if (ns > kMaxStrips) return ISP_ERR_UNSUPPORTED;
    pl->C = C;
    pl->ns = ns;
    if (C == 8) ring_geometry<8>(T2max, ns, pl); else ring_geometry<4>(T2max, ns, pl);
    const int floats_total = (ns - 1) * pl->full_floats + pl->last_floats[0] + pl->last_floats[1];
    pl->bits_pitch = ns * 32;
    const size_t bits_bytes = size_t(C == 8 ? T1max : (T1max + 1) / 2) * pl->bits_pitch;
    // one utterance per CTA while every utterance still gets its own SM; two beyond that, so
    // that co-resident chains sit on different sub-partitions of one SM
    int slots = g_opt_slots > 0 ? g_opt_slots : (B > sm_count ? 2 : 1);
    if (slots > kMaxSlots) slots = kMaxSlots;
    while (slots > 1 && 32 * slots * (ns + 1) > kMaxThreads) --slots;
    for (;; --slots) {
        const size_t budget = smem_limit / slots;
        for (int in_smem = 1; in_smem >= 0; --in_smem) {
            const size_t fixed = slot_bytes_for(ns, floats_total, 0, in_smem ? bits_bytes : 0);
            if (fixed + size_t(kMinRing) * floats_total * 4 > budget) continue;
            int rows = int((budget - fixed - 128) / (size_t(floats_total) * 4));
            rows = rows / kR * kR;
            if (rows > kR * kMaxStages) rows = kR * kMaxStages;
            if (in_smem && rows < 72 && slots == 1 && bits_bytes > 64 * 1024) continue;   // long utterances: prefer a deeper ring over resident bits
            if (g_opt_ring_rows > 0) {
                int want = (g_opt_ring_rows + kR - 1) / kR * kR;
                if (want < kMinRing) want = kMinRing;
                if (want < rows) rows = want;
            }
            pl->slots = slots;
            pl->ring_rows = rows;
            pl->bits_smem = in_smem != 0;
            pl->slot_bytes = slot_bytes_for(ns, floats_total, rows, in_smem ? bits_bytes : 0);
            pl->bits_ws_words = in_smem ? 0 : bits_bytes / 4;
            return 0;
        }
        if (slots == 1) break;
    }
    return ISP_ERR_UNSUPPORTED;
}

// tensor maps over the logits: (T2max, T1max, B) fp32, box = (width, kR rows, 1), no swizzle
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static bool make_maps(MasMaps* maps, int C, const float* logp, int64_t sB, int64_t sT1, int B, int T1max, int T2max) {
    static PFN_encodeTiled enc = nullptr;
    if (!enc) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess || !ptr) return false;
        enc = reinterpret_cast<PFN_encodeTiled>(ptr);
    }
    if ((reinterpret_cast<uintptr_t>(logp) & 15) || (sT1 & 3) || (sB & 3)) return false;
    cuuint64_t dims[3] = {cuuint64_t(T2max), cuuint64_t(T1max), cuuint64_t(B)};
    cuuint64_t strides[2] = {cuuint64_t(sT1) * 4, cuuint64_t(B > 1 ? sB : sT1 * T1max) * 4};
    cuuint32_t estr[3] = {1, 1, 1};
    for (int i = 0; i < kNumBox; ++i) {
        cuuint32_t box[3] = {cuuint32_t(C == 8 ? box_width<8>(i) : box_width<4>(i)), cuuint32_t(kR), 1};
        CUresult r = enc(&maps->m[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(logp), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return false;
    }
    return true;
}

size_t mas_workspace_bytes(int B, int T1max, int T2max) {
    if (B <= 0 || T1max <= 0 || T2max <= 0) return 0;
    // one bit per cell, rows padded to whole strips of either width, row pairs rounded up
    const size_t words = (size_t(T1max) + 1) * ((size_t(T2max) + 255) / 256) * 8;
    return 256 + size_t(B) * words * 4;
}

template <int C, bool BS, bool MULTI>
static int launch_one(const MasMaps& maps, const MasParams& p, const MasPlan& pl, cudaStream_t stream) {
    auto kern = mas_kernel<C, BS, MULTI>;
    const size_t smem = pl.slot_bytes * pl.slots;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(mas_kernel)");
    const int grid = (p.B + pl.slots - 1) / pl.slots;
    kern<<<grid, 32 * pl.slots * (pl.ns + 1), smem, stream>>>(maps, p);
    e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "mas_kernel launch");
    return 0;
}

int mas_forward(const float* logp, int64_t sB, int64_t sT1, int64_t sT2,
                const int64_t* text_len, const int64_t* mel_len, int B, int T1max, int T2max,
                int16_t* attn_hard, int64_t* durations, void* ws, size_t ws_bytes, cudaStream_t stream) {
    if (!logp || !text_len || !mel_len || !attn_hard || !ws) { set_error("isp_mas_forward: null pointer"); return ISP_ERR_INVALID; }
    if (B <= 0 || T1max <= 0 || T2max <= 0) { set_error("isp_mas_forward: B, T1max, T2max must be positive"); return ISP_ERR_INVALID; }
    if (sT2 != 1) { set_error("isp_mas_forward: sT2 must be 1 (token axis contiguous), got %lld", (long long)sT2); return ISP_ERR_INVALID; }
    if (sT1 < T2max || (B > 1 && sB < int64_t(T1max - 1) * sT1 + T2max)) { set_error("isp_mas_forward: overlapping strides"); return ISP_ERR_INVALID; }
    if (T2max > ISP_MAS_MAX_T2 || T1max >= (1 << 24)) {
        set_error("isp_mas_forward: T2max=%d > %d or T1max=%d >= 2^24 is not covered", T2max, ISP_MAS_MAX_T2, T1max);
        return ISP_ERR_UNSUPPORTED;
    }
    if (ws_bytes < mas_workspace_bytes(B, T1max, T2max) || (reinterpret_cast<uintptr_t>(ws) & 15)) {
        set_error("isp_mas_forward: workspace too small or not 16 B aligned (%zu < %zu)", ws_bytes, mas_workspace_bytes(B, T1max, T2max));
        return ISP_ERR_WORKSPACE;
    }
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
    p.bits_stride = int64_t(pl.bits_ws_words);
    p.ns = pl.ns; p.slots = pl.slots; p.full_floats = pl.full_floats;
    p.last_floats[0] = pl.last_floats[0]; p.last_floats[1] = pl.last_floats[1];
    p.ring_rows = pl.ring_rows; p.bits_pitch = pl.bits_pitch; p.dbg = g_opt_dbg;
    p.slot_bytes = int(pl.slot_bytes);
    MasMaps maps;
    memset(&maps, 0, sizeof(maps));
    p.tma = make_maps(&maps, pl.C, logp, sB, sT1, B, T1max, T2max) ? 1 : 0;

    cudaError_t e = cudaMemsetAsync(ws, 0, 256, stream);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(status)");

#define ISP_MAS_DISPATCH(CC)                                                                          \
    if (pl.bits_smem) return pl.ns > 1 ? launch_one<CC, true, true>(maps, p, pl, stream)             \
                                       : launch_one<CC, true, false>(maps, p, pl, stream);           \
    return pl.ns > 1 ? launch_one<CC, false, true>(maps, p, pl, stream)                              \
                     : launch_one<CC, false, false>(maps, p, pl, stream);
    if (pl.C == 4) { ISP_MAS_DISPATCH(4) }
    ISP_MAS_DISPATCH(8)
#undef ISP_MAS_DISPATCH
}

}  // namespace isp
