// Monotonic Alignment Search for sm_100a: forward max-plus DP, packed backpointer
// bits, warp-parallel backtrack, dense int16 path and int64 durations -- one launch.
//
// Reference semantics (paths relative to the reference root):
//   tts/modules/aligner/mas.py:8-26      mas_width1 (DP, tie rule, backtrack)
//   tts/modules/aligner/cuda_mas.py:11-46 cuda_b_mas (the GPU kernel this replaces)
//   tts/models/acoustic/modules/alignment.py:275  durations = attn_hard.sum(dim=1)
//
// Shape of the computation.  Q[i][j] = x[i][j] + max(Q[i-1][j-1], Q[i-1][j]) depends on
// row i-1 only, so the text axis is parallel and the frame axis is a serial chain.
// One CTA owns one utterance:
//   * strip warps: warp s owns columns [s*32*C, (s+1)*32*C), lane l owns C consecutive
//     ones, the previous DP row lives in registers.  Per row and lane: one shuffle for
//     the left neighbour, then per cell one compare, one select, one fp32 add, one
//     predicated OR into the lane's backpointer bits -- no block barrier in the row loop.
//     Strips run as a skewed wavefront: strip s trails strip s-1 by one ring stage and
//     takes one boundary value per row from shared memory (acquire/release flag per stage).
//   * producer warp: streams the utterance's logit rows HBM -> shared memory with 1-D
//     bulk async copies (TMA, cp.async.bulk) into a multi-stage ring; completion and
//     slot reuse are tracked with mbarriers.  Only the valid T2_b columns of the valid
//     T1_b rows are ever read.
//   * filler warp: zero-fills the utterance's dense int16 output block and its duration
//     row with 16 B streaming stores while the DP runs.
//   * backtrack (warp 0): 32 rows per step.  Each lane fetches the 32-bit window of
//     backpointer bits its row can touch (the path moves at most one column per row),
//     the windows are broadcast by shuffle and the dependent chain j -= bit runs in
//     registers; then all 32 lanes write their row's 1 and the durations of the tokens
//     that start in this block (ballot + clz), so nothing is re-read to build durations.
// Backpointers cost one bit per cell: row-major, column j at bit j%32 of word j/32, in
// shared memory when an utterance's bits fit, else in the caller's workspace (L2-resident).
//
// Bit-exactness: each cell does exactly the reference's one fp32 add on top of an exact
// max; the comparison is the reference's `>=` (ties and -inf >= -inf take the diagonal).

#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "common.cuh"
#include "isp_internal.h"

namespace isp {

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
    int ns;                // strip warps per CTA
    int pitch;             // ring row pitch, floats (multiple of 4)
    int stage_rows;        // rows per ring stage
    int stages;            // ring stages
    int bits_pitch;        // words per row of bits (= ns * C)
    int tma;               // 1: rows are 16 B aligned in global memory -> bulk copies
};

constexpr int kMaxStages = 8;

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

// ---- one DP row for one lane --------------------------------------------------------
// q[] holds row i-1 on entry and row i on exit.  Returns the lane's C backpointer bits
// (bit c set <=> predecessor of column base+c is the diagonal).
template <int C>
ISP_DEVINL uint32_t dp_row(float (&q)[C], const float (&x)[C], float left) {
    uint32_t bits = 0;
#pragma unroll
    for (int c = C - 1; c >= 1; --c) {
        const float a = q[c - 1], b = q[c];
        const bool diag = a >= b;              // mas.py:17 -- ties take j-1
        bits |= diag ? (1u << c) : 0u;
        q[c] = x[c] + (diag ? a : b);          // mas.py:14 -- one fp32 add per cell
    }
    {
        const float b = q[0];
        const bool diag = left >= b;           // left == NaN at global column 0: false, keeps b
        bits |= diag ? 1u : 0u;
        q[0] = x[0] + (diag ? left : b);
    }
    return bits;
}

template <int C, bool BITS_SMEM, bool MULTI>
__global__ void __launch_bounds__(32 * (8 + 2), 1) mas_kernel(const MasParams p) {
    constexpr int W = 32 * C;  // columns per strip
    extern __shared__ __align__(128) unsigned char smem_raw[];

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int ns = p.ns;
    const int b = blockIdx.x;

    // ---- carve shared memory ----------------------------------------------------------
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
    uint64_t* empty = full + kMaxStages;
    int* prog = reinterpret_cast<int*>(empty + kMaxStages);            // [8]
    const int ring_rows = p.stage_rows * p.stages;
    float* bnd = reinterpret_cast<float*>(smem_raw + 256);              // [ns][ring_rows] (MULTI)
    size_t off = 256 + (MULTI ? sizeof(float) * size_t(ns) * ring_rows : 0);
    off = (off + 127) & ~size_t(127);
    float* ring = reinterpret_cast<float*>(smem_raw + off);             // [ring_rows][pitch] + W pad
    off += sizeof(float) * (size_t(ring_rows) * p.pitch + W);
    off = (off + 15) & ~size_t(15);
    unsigned char* bits_base;
    if (BITS_SMEM) bits_base = smem_raw + off;
    else bits_base = reinterpret_cast<unsigned char*>(p.bits_ws + size_t(b) * p.bits_stride);
    const int bits_row_bytes = p.bits_pitch * 4;

    // ---- lengths (read on device; clamped for memory safety, reported via status) ------
    long long n64 = p.mel_len[b], m64 = p.text_len[b];
    const bool bad = n64 < 1 || n64 > p.T1max || m64 < 1 || m64 > p.T2max;
    const int n = int(n64 < 1 ? 1 : (n64 > p.T1max ? p.T1max : n64));   // frames
    const int m = int(m64 < 1 ? 1 : (m64 > p.T2max ? p.T2max : m64));   // tokens
    const int ns_active = (m + W - 1) / W;
    const int nchunks = (n + p.stage_rows - 1) / p.stage_rows;

    if (threadIdx.x == 0) {
        if (bad) atomicAdd(p.status, 1);
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], ns_active);
        }
        for (int s = 0; s < 8; ++s) prog[s] = 0;
        fence_mbar_init();
    }
    __syncthreads();

    const float* xb = p.logp + int64_t(b) * p.sB;

    if (warp < ns) {
        // =========================== strip warps: forward DP ===========================
        const int s = warp;
        if (s < ns_active) {
            float q[C];
            const int col0 = s * W + lane * C;
            const float qnan = __int_as_float(0x7fffffff);
            float left_carry = qnan;           // boundary value of row r0-1 from strip s-1
            float* bnd_mine = bnd + size_t(s) * ring_rows;
            const float* bnd_prev = bnd + size_t(s > 0 ? s - 1 : 0) * ring_rows;

            for (int ch = 0; ch < nchunks; ++ch) {
                const int st = ch % p.stages;
                const uint32_t ph = (ch / p.stages) & 1;
                const int r0 = ch * p.stage_rows;
                const int rows = min(p.stage_rows, n - r0);
                mbar_wait(&full[st], ph);
                if (MULTI && s > 0) {
                    // strip s-1 must have finished this chunk (its boundary values are ours)
                    if (lane == 0) {
                        uint32_t spins = 0;
                        while (ld_acquire_shared(&prog[s - 1]) < r0 + rows) {
                            if (++spins > (1u << 26)) __trap();
                        }
                    }
                    __syncwarp();
                }
                const float* xs = ring + size_t(st) * p.stage_rows * p.pitch + col0;
                const int slot0 = st * p.stage_rows;
                int r = 0;
                if (ch == 0) {
                    // row 0: Q[0][0] = x[0][0], Q[0][j>0] = -inf   (mas.py:11)
                    float x[C];
#pragma unroll
                    for (int v = 0; v < C / 4; ++v) {
                        const float4 t = *reinterpret_cast<const float4*>(xs + 4 * v);
                        x[4 * v] = t.x; x[4 * v + 1] = t.y; x[4 * v + 2] = t.z; x[4 * v + 3] = t.w;
                    }
#pragma unroll
                    for (int c = 0; c < C; ++c) q[c] = (col0 + c == 0) ? x[c] : -CUDART_INF_F;
                    if (MULTI && lane == 31) bnd_mine[slot0] = q[C - 1];
                    r = 1;
                }
#pragma unroll 4
                for (; r < rows; ++r) {
                    const int i = r0 + r;
                    float left = __shfl_up_sync(0xffffffffu, q[C - 1], 1);
                    if (lane == 0) {
                        left = qnan;
                        if (MULTI && s > 0) left = (r == 0) ? left_carry : bnd_prev[slot0 + r - 1];
                    }
                    float x[C];
                    const float* xr = xs + size_t(r) * p.pitch;
#pragma unroll
                    for (int v = 0; v < C / 4; ++v) {
                        const float4 t = *reinterpret_cast<const float4*>(xr + 4 * v);
                        x[4 * v] = t.x; x[4 * v + 1] = t.y; x[4 * v + 2] = t.z; x[4 * v + 3] = t.w;
                    }
                    uint32_t bits = dp_row<C>(q, x, left);
                    unsigned char* brow = bits_base + size_t(i) * bits_row_bytes + s * (W / 8);
                    if (C == 8) {
                        brow[lane] = static_cast<unsigned char>(bits);
                    } else {  // C == 4: two lanes share a byte
                        bits |= __shfl_down_sync(0xffffffffu, bits, 1) << 4;
                        if ((lane & 1) == 0) brow[lane >> 1] = static_cast<unsigned char>(bits);
                    }
                    if (MULTI && lane == 31) bnd_mine[slot0 + r] = q[C - 1];
                }
                if (MULTI) {
                    if (s > 0 && lane == 0) left_carry = bnd_prev[slot0 + rows - 1];
                    // publish progress: the lane that wrote the boundary values releases them
                    if (lane == 31) st_release_shared(&prog[s], r0 + rows);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[st]);
            }
        }
    } else if (warp == ns) {
        // =========================== producer warp: HBM -> ring ========================
        const uint32_t row_bytes = (uint32_t(m) * 4u + 15u) & ~15u;
        const uint64_t pol = policy_evict_first();
        for (int ch = 0; ch < nchunks; ++ch) {
            const int st = ch % p.stages;
            const int use = ch / p.stages;
            if (use > 0) mbar_wait(&empty[st], (use - 1) & 1);
            const int r0 = ch * p.stage_rows;
            const int rows = min(p.stage_rows, n - r0);
            float* dst = ring + size_t(st) * p.stage_rows * p.pitch;
            const float* src = xb + int64_t(r0) * p.sT1;
            if (p.tma) {
                if (lane == 0) mbar_arrive_expect_tx(&full[st], uint32_t(rows) * row_bytes);
                __syncwarp();
                for (int r = lane; r < rows; r += 32)
                    bulk_g2s_hint(dst + size_t(r) * p.pitch, src + int64_t(r) * p.sT1, row_bytes, &full[st], pol);
            } else {
                // rows not 16 B aligned in global memory: coalesced 4 B loads through registers
                for (int r = 0; r < rows; ++r) {
                    const float* g = src + int64_t(r) * p.sT1;
                    float* d = dst + size_t(r) * p.pitch;
                    int c = lane;
                    for (; c + 96 < m; c += 128) {
                        const float v0 = __ldg(g + c), v1 = __ldg(g + c + 32), v2 = __ldg(g + c + 64), v3 = __ldg(g + c + 96);
                        d[c] = v0; d[c + 32] = v1; d[c + 64] = v2; d[c + 96] = v3;
                    }
                    for (; c < m; c += 32) d[c] = __ldg(g + c);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&full[st]);
            }
        }
    } else {
        // =========================== filler warp: zero the outputs =====================
        const size_t cells = size_t(p.T1max) * p.T2max;
        warp_zero_fill(reinterpret_cast<char*>(p.hard + size_t(b) * cells), cells * sizeof(int16_t), lane);
        if (p.dur) {
            int64_t* d = p.dur + size_t(b) * p.T2max;
            for (int j = lane; j < p.T2max; j += 32) d[j] = 0;
        }
    }

    // bits (shared or global) and the zero-filled outputs become visible to warp 0
    __syncthreads();
    if (warp != 0) return;

    // =============================== backtrack (warp 0) ===============================
    // mas.py:20-24.  Rows i0, i0-1, ..., i0-31 per step; lane t owns row i0-t.
    int16_t* hard_b = p.hard + size_t(b) * p.T1max * p.T2max;
    int64_t* dur_b = p.dur ? p.dur + size_t(b) * p.T2max : nullptr;
    const uint32_t* bits_w = reinterpret_cast<const uint32_t*>(bits_base);
    int j = m - 1;           // token index on row i0
    int last_start = n;      // first row of token j+1 (exclusive end of token j)
    const uint32_t lt_mask = (1u << lane) - 1u;
    for (int i0 = n - 1; i0 >= 0; i0 -= 32) {
        const int row = i0 - lane;
        // window of this lane's row: bit k <-> column j-31+k  (path stays within it)
        uint32_t win = 0;
        if (row >= 1) {   // row 0 has no predecessor (its bits were never written)
            const int qw = j >> 5;
            const uint32_t* wr = bits_w + size_t(row) * p.bits_pitch;
            const uint32_t hi = wr[qw];
            const uint32_t lo = qw > 0 ? wr[qw - 1] : 0u;
            win = __funnelshift_rc(lo, hi, (j & 31) + 1);
        }
        uint32_t rel = 31;   // position of the current column inside the window
        uint32_t dec = 0;    // bit t set <=> the path steps to the diagonal below row i0-t
#pragma unroll
        for (int t = 0; t < 32; ++t) {
            const uint32_t w = __shfl_sync(0xffffffffu, win, t);
            const uint32_t bit = (w >> rel) & 1u;
            rel -= bit;
            dec |= bit << t;
        }
        const int col = j - __popc(dec & lt_mask);       // this lane's token
        if (row >= 0) hard_b[size_t(row) * p.T2max + col] = 1;
        // a token starts on this row if the path leaves it diagonally, or on row 0
        const bool starts = row >= 0 && (((dec >> lane) & 1u) || row == 0);
        const uint32_t smask = __ballot_sync(0xffffffffu, starts);
        if (starts && dur_b) {
            const uint32_t lower = smask & lt_mask;      // starts of later tokens in this block
            const int next_start = lower ? i0 - (31 - __clz(lower)) : last_start;
            dur_b[col] = int64_t(next_start - row);
        }
        if (smask) last_start = i0 - (31 - __clz(smask));
        const int valid = min(32, i0 + 1);
        j -= __popc(valid == 32 ? dec : (dec & ((1u << valid) - 1u)));
    }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------

static int g_opt_cols_per_lane = 0;
static int g_opt_ring_rows = 0;

int mas_set_option(const char* key, int value, int* prev) {
    if (!strcmp(key, "mas.cols_per_lane")) { *prev = g_opt_cols_per_lane; g_opt_cols_per_lane = value; return 0; }
    if (!strcmp(key, "mas.ring_rows")) { *prev = g_opt_ring_rows; g_opt_ring_rows = value; return 0; }
    return -1;
}

struct MasPlan {
    int C, ns, pitch, stage_rows, stages, bits_pitch;
    bool bits_smem, multi;
    size_t smem_bytes;
    size_t bits_ws_words;  // per utterance, when !bits_smem
};

static size_t mas_smem_bytes(const MasPlan& pl, int T1max) {
    const int W = 32 * pl.C;
    const int ring_rows = pl.stage_rows * pl.stages;
    size_t off = 256 + (pl.multi ? sizeof(float) * size_t(pl.ns) * ring_rows : 0);
    off = (off + 127) & ~size_t(127);
    off += sizeof(float) * (size_t(ring_rows) * pl.pitch + W);
    off = (off + 15) & ~size_t(15);
    if (pl.bits_smem) off += size_t(T1max) * pl.bits_pitch * 4;
    return off;
}

static int mas_plan(int B, int T1max, int T2max, MasPlan* pl) {
    int sm_count = 148;
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, 0);
    int C = g_opt_cols_per_lane;
    if (C != 4 && C != 8) {
        // few utterances per SM: the frame chain is the bound -> narrower strips, more warps;
        // many: issue slots are the bound -> fewer instructions per cell.
        C = (B <= sm_count && T2max > 128) ? 4 : 8;
    }
    int ns = (T2max + 32 * C - 1) / (32 * C);
    if (ns > 8) { C = 8; ns = (T2max + 255) / 256; }
    if (ns > 8) return ISP_ERR_UNSUPPORTED;
    pl->C = C;
    pl->ns = ns;
    pl->multi = ns > 1;
    pl->pitch = (T2max + 3) & ~3;
    pl->bits_pitch = ns * C;
    // rows in flight: enough bytes to cover HBM latency at a few utterances per SM
    int ring_rows = g_opt_ring_rows > 0 ? g_opt_ring_rows : 48;
    const size_t row_bytes = size_t(pl->pitch) * 4;
    while (ring_rows > 8 && ring_rows * row_bytes > 64 * 1024) ring_rows -= 8;
    pl->stage_rows = ring_rows >= 32 ? 16 : 8;
    pl->stages = ring_rows / pl->stage_rows;
    if (pl->stages < 2) pl->stages = 2;
    if (pl->stages > kMaxStages) pl->stages = kMaxStages;
    if (pl->stage_rows > T1max) { pl->stage_rows = T1max; }
    // bits in shared memory when the CTA still fits ~2 per SM
    pl->bits_smem = true;
    size_t need = mas_smem_bytes(*pl, T1max);
    if (need > 100 * 1024) {
        pl->bits_smem = false;
        need = mas_smem_bytes(*pl, T1max);
    }
    if (need > 220 * 1024) return ISP_ERR_UNSUPPORTED;
    pl->smem_bytes = need;
    pl->bits_ws_words = pl->bits_smem ? 0 : size_t(T1max) * pl->bits_pitch;
    return 0;
}

size_t mas_workspace_bytes(int B, int T1max, int T2max) {
    if (B <= 0 || T1max <= 0 || T2max <= 0) return 0;
    // sized for the larger of the two strip widths so options can change between calls
    size_t words = size_t(T1max) * ((size_t(T2max) + 255) / 256) * 8;
    return 256 + size_t(B) * words * 4;
}

template <int C, bool BS, bool MULTI>
static int launch_one(const MasParams& p, const MasPlan& pl, cudaStream_t stream) {
    auto kern = mas_kernel<C, BS, MULTI>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(pl.smem_bytes));
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(mas_kernel)");
    kern<<<p.B, 32 * (pl.ns + 2), pl.smem_bytes, stream>>>(p);
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
    p.bits_ws = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(ws) + 256);
    p.bits_stride = int64_t(pl.bits_ws_words);
    p.ns = pl.ns; p.pitch = pl.pitch; p.stage_rows = pl.stage_rows; p.stages = pl.stages; p.bits_pitch = pl.bits_pitch;
    p.tma = ((reinterpret_cast<uintptr_t>(logp) & 15) == 0 && (sB & 3) == 0 && (sT1 & 3) == 0 && (T2max & 3) == 0) ? 1 : 0;

    cudaError_t e = cudaMemsetAsync(ws, 0, 256, stream);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(status)");

#define ISP_MAS_DISPATCH(CC)                                                              \
    if (pl.bits_smem) return pl.multi ? launch_one<CC, true, true>(p, pl, stream)        \
                                      : launch_one<CC, true, false>(p, pl, stream);      \
    else return pl.multi ? launch_one<CC, false, true>(p, pl, stream)                    \
                         : launch_one<CC, false, false>(p, pl, stream);
    if (pl.C == 4) { ISP_MAS_DISPATCH(4) } else { ISP_MAS_DISPATCH(8) }
#undef ISP_MAS_DISPATCH
}

}  // namespace isp
