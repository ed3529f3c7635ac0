// Forward-sum (CTC) alignment loss on the attention log-likelihoods, SURVEY.md section 8 row f-4.
//
// Reference (tts/models/acoustic/loss.py:41-79, AttentionCTCLoss.forward): pad a blank column with value blank_logprob in
// front of attn_logits (:67), log_softmax over the T2max + 1 columns (:69), nn.CTCLoss(zero_infinity=True) with targets
// 1 .. text_len[b], input_lengths = mel_len, target_lengths = text_len (:73-78; mean over the batch of nll_b / text_len[b]).
//
// The same lattice as MAS with (+, x) in place of (max, +): per frame i and token j
//     blank_j' = pB_i  * (blank_j + label_{j-1})
//     label_j' = p_ij  * (label_j + blank_j + label_{j-1})          p = softmax over [blank, tokens] of frame i
// and the likelihood is label_{T2-1} + blank_{T2} after the last frame.  It is a serial chain over the frames, so the kernel
// is built like the MAS one: ONE WARP PER UTTERANCE, lane l owns G consecutive tokens and, at step t, works on frame t - l
// (the wavefront is skewed across lanes, so the one value a lane needs from its left neighbour was produced two steps
// earlier and its shuffle is off the chain; no barrier anywhere).  The variables are base-2 logarithms.  With
// lse(x, y) = max(x, y) + log2(1 + 2^-|x - y|) the blank needs one lse and the label one more on top of it (the inner sum
// is shared), i.e. 2 ex2 + 2 lg2 per cell and a dozen FMA-pipe instructions.  (A linear-domain recursion in block floating
// point -- one ex2 per cell, a power-of-two exponent per token -- was built first and measured 2x slower: its integer
// exponent bookkeeping runs on the half-rate ALU pipe, 40 instructions per cell.)  Unreachable states hold -1e30 instead
// of -inf so that no difference is ever NaN.
//
//   ctc_rownorm_kernel   Z_i = log2(2^blank + sum_j 2^logit_ij) per valid frame (one warp per frame, coalesced)
//   ctc_alpha_kernel     forward variables (kept in the workspace for the gradient) and nll_b
//   ctc_beta_grad_kernel backward variables in the mirrored skew and d nll_b / d attn_logits = p_ij - posterior_ij
//
// Accuracy: fp32 with ex2.approx / lg2.approx (abs error ~2^-22 per lse in log2 units); tests state the tolerance against
// torch's CTC in float64.

#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include <algorithm>

#include "common.cuh"
#include "isp_internal.h"

namespace isp {

constexpr float kCtcLog2e = 1.4426950408889634f;
constexpr float kCtcLn2 = 0.6931471805599453f;

ISP_DEVINL float ctc_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
ISP_DEVINL float ctc_lg2(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
constexpr float kCtcNeg = -1.0e30f;                  // log2 of "no mass"; finite, so that differences are never NaN
// log2(2^x + 2^y)
ISP_DEVINL float ctc_lse2(float x, float y) { return fmaxf(x, y) + ctc_lg2(1.0f + ctc_ex2(-fabsf(x - y))); }

// Per-lane ring of prefetched rows in shared memory: a lane copies the words of its own frames with 4 B async copies and
// reads only what it copied itself, so cp.async.wait_group is all the synchronisation there is.  Depth 8: a step takes a
// few hundred cycles, a load from HBM under load well over a thousand.  A lane's G logits of a frame are 16 B vectors when
// T2max % 4 == 0 (G is a multiple of 4); the forward variables are kept in the workspace in the order the forward pass
// produces them -- (step, lane, token), "skewed" -- and the backward pass at its step s wants exactly the forward step
// n + 30 - s for every lane, so both passes touch them with coalesced 16 B vectors.
constexpr int kCtcDepth = 8;
ISP_DEVINL void ctc_cp4(float* sdst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(sdst)), "l"(gsrc) : "memory");
}
ISP_DEVINL void ctc_cp8(float* sdst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(sdst)), "l"(gsrc) : "memory");
}
ISP_DEVINL void ctc_cp16(float* sdst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(sdst)), "l"(gsrc) : "memory");
}
ISP_DEVINL int ctc_clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// ---- row normaliser ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ctc_rownorm_kernel(const float* __restrict__ logits, const int64_t* __restrict__ mel_len, float* __restrict__ z2,
                   int B, int T1max, int T2max, float blank2 /* blank_logprob * log2(e) */) {
    const int lane = threadIdx.x & 31;
    const long long rows = (long long)B * T1max;
    const long long wstride = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows; row += wstride) {
        const int b = int(row / T1max), i = int(row - (long long)b * T1max);
        if (i >= mel_len[b]) continue;
        const float* x = logits + row * T2max;
        float m = blank2;
        for (int j = lane; j < T2max; j += 32) m = fmaxf(m, __ldg(x + j) * kCtcLog2e);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        float s = lane == 0 ? ctc_ex2(blank2 - m) : 0.0f;
        for (int j = lane; j < T2max; j += 32) s += ctc_ex2(fmaf(__ldg(x + j), kCtcLog2e, -m));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        // {Z_i, M_i}: the recursions use log2 p relative to the row's LARGEST term (M_i), so that the forward and backward
        // variables stay within tens of units instead of drifting by log2 p per frame (fp32 resolution); the softmax
        // probabilities of the gradient use Z_i, and the forward pass adds the sum of Z_i - M_i back into nll
        if (lane == 0) reinterpret_cast<float2*>(z2)[row] = make_float2(m + log2f(s), m);
    }
}

// ---- forward variables ----------------------------------------------------------------------------------------------
// Workspace layout: alpha (B, T1max + 31, 32, G) float (log2 of the labels' forward variables, by forward step), z2 (B, T1max)
// float2 {Z_i, M_i}, l2p (B) float.
template <int G>
__global__ void __launch_bounds__(128)
ctc_alpha_kernel(const float* __restrict__ logits, const float* __restrict__ z2, const int64_t* __restrict__ text_len,
                 const int64_t* __restrict__ mel_len, float* __restrict__ alpha_ws, float* __restrict__ l2p,
                 float* __restrict__ nll, int B, int T1max, int T2max, float blank2) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= B) return;
    const long long n64 = mel_len[b], m64 = text_len[b];
    const int n = int(n64 < 1 ? 1 : (n64 > T1max ? T1max : n64));      // frames
    const int m = int(m64 < 1 ? 1 : (m64 > T2max ? T2max : m64));      // tokens
    const int j0 = lane * G;
    const float* xb = logits + (size_t)b * T1max * T2max;
    const float2* zb = reinterpret_cast<const float2*>(z2) + (size_t)b * T1max;
    float* ab = alpha_ws ? alpha_ws + (size_t)b * (T1max + 31) * (32 * G) : nullptr;      // [step][lane][G]

    float A[G], Bk[G];                 // log2 of label_j, blank_j of the previous frame
#pragma unroll
    for (int g = 0; g < G; ++g) { A[g] = kCtcNeg; Bk[g] = kCtcNeg; }
    if (lane == 0) Bk[0] = 0.0f;       // virtual frame -1: all the mass in the blank before token 0
    // what the right neighbour takes: this lane's last label two steps ago
    float h1v = kCtcNeg, h2v = kCtcNeg;
    // the rows of the lane's next frames, kCtcDepth - 1 steps ahead (addresses clamped; masks applied when the row is used)
    extern __shared__ float ctc_smem[];
    constexpr int W = G + 4;                                        // words per lane and stage: G logits, z (16 B slots)
    float* ring = ctc_smem + (size_t)(threadIdx.x >> 5) * kCtcDepth * 32 * W + lane * W;
    const bool vec = (T2max & 3) == 0 && (reinterpret_cast<uintptr_t>(logits) & 15) == 0;
    auto issue = [&](int step) {
        const int i = ctc_clampi(step - lane, 0, n - 1);
        float* dst = ring + (step & (kCtcDepth - 1)) * 32 * W;
        const float* src = xb + (size_t)i * T2max;
        if (vec) {
#pragma unroll
            for (int g = 0; g < G; g += 4) ctc_cp16(dst + g, src + min(j0 + g, T2max - 4));
        } else {
#pragma unroll
            for (int g = 0; g < G; ++g) ctc_cp4(dst + g, src + min(j0 + g, T2max - 1));
        }
        ctc_cp8(dst + G, zb + i);
        cp_async_commit();
    };
    for (int st = 0; st < kCtcDepth - 1; ++st) issue(st);
    float osum = 0.0f;                                            // sum of Z_i - M_i over the frames (lane 0 sees them all)
    const int steps = n + 31;
    for (int t = 0; t < steps; ++t) {
        const int i = t - lane;                                   // this lane's frame
        const float lv = __shfl_up_sync(0xffffffffu, h2v, 1);
        issue(t + kCtcDepth - 1);
        cp_async_wait_pending(kCtcDepth - 1);                     // the copies of step t have landed
        const float* row = ring + (t & (kCtcDepth - 1)) * 32 * W;
        const float zr = row[G], mr = row[G + 1];
        float lp[G];                                              // log2 p_ij + (Z_i - M_i); "no mass" for tokens >= m
#pragma unroll
        for (int g = 0; g < G; ++g) lp[g] = j0 + g < m ? fmaf(row[g], kCtcLog2e, -mr) : kCtcNeg;
        const float lpB = blank2 - mr;
        if (i >= 0 && i < n) {
            osum += zr - mr;
#pragma unroll
            for (int g = G - 1; g >= 0; --g) {
                const float am1 = g > 0 ? A[g - 1] : (lane == 0 ? kCtcNeg : lv);     // the left neighbour's label, previous frame
                const float s = ctc_lse2(Bk[g], am1);
                A[g] = fmaxf(lp[g] + ctc_lse2(A[g], s), kCtcNeg);
                Bk[g] = fmaxf(lpB + s, kCtcNeg);
            }
            if (ab) {
                // the gradient needs the labels' forward variables only; forward-step order, 16 B vectors
                float4* dst = reinterpret_cast<float4*>(ab + (size_t)t * (32 * G) + j0);
#pragma unroll
                for (int g = 0; g < G; g += 4) dst[g >> 2] = make_float4(A[g], A[g + 1], A[g + 2], A[g + 3]);
            }
        }
        h2v = h1v;
        h1v = A[G - 1];
    }
    // likelihood = label_{m-1} + blank_m after frame n-1 (every lane now holds its frame n-1)
    const int la = (m - 1) / G, lb = m / G;
    float va = kCtcNeg, vb = kCtcNeg;
#pragma unroll
    for (int g = 0; g < G; ++g) {
        if (lane == la && g == (m - 1) - la * G) va = A[g];
        if (lane == lb && g == m - lb * G) vb = Bk[g];
    }
    va = __shfl_sync(0xffffffffu, va, la);
    vb = __shfl_sync(0xffffffffu, vb, lb);
    if (lane == 0) {
        const float l2 = ctc_lse2(va, vb);                                  // log2 of P * 2^osum
        nll[b] = l2 > 0.5f * kCtcNeg ? (osum - l2) * kCtcLn2 : CUDART_INF_F;     // +inf: no alignment exists (mel_len < text_len)
        if (l2p) l2p[b] = l2;
    }
}

// ---- backward variables and the gradient ------------------------------------------------------------------------------
// With H_i(s) = p_i(s) * beta_i(s) (beta_i(s): probability of frames i+1.. given state s at frame i):
//     beta_i(label_j) = H_{i+1}(label_j) + H_{i+1}(blank_{j+1}) + H_{i+1}(label_{j+1})
//     beta_i(blank_j) = H_{i+1}(blank_j) + H_{i+1}(label_j)
// posterior_ij = alpha_i(label_j) beta_i(label_j) / P, and through the log_softmax d nll / d logit_ij = p_ij - posterior_ij
// (the posteriors of a frame sum to 1).  Mirrored skew: lane l works on frame n-1 - (t - (31 - l)) and takes its right
// neighbour's first token from two steps earlier.
template <int G>
__global__ void __launch_bounds__(128)
ctc_beta_grad_kernel(const float* __restrict__ logits, const float* __restrict__ z2, const int64_t* __restrict__ text_len,
                     const int64_t* __restrict__ mel_len, const float* __restrict__ alpha_ws,
                     const float* __restrict__ l2p, const float* __restrict__ nll, const float* __restrict__ grad_scale,
                     float* __restrict__ grad,
                     int B, int T1max, int T2max, float blank2) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= B) return;
    const long long n64 = mel_len[b], m64 = text_len[b];
    const int n = int(n64 < 1 ? 1 : (n64 > T1max ? T1max : n64));
    const int m = int(m64 < 1 ? 1 : (m64 > T2max ? T2max : m64));
    const int j0 = lane * G;
    const float* xb = logits + (size_t)b * T1max * T2max;
    const float2* zb = reinterpret_cast<const float2*>(z2) + (size_t)b * T1max;
    const float* ab = alpha_ws + (size_t)b * (T1max + 31) * (32 * G);                     // [forward step][lane][G]
    float* gb = grad + (size_t)b * T1max * T2max;
    const float nl = nll[b];
    const float gs = grad_scale[b];
    const bool dead = !(nl < CUDART_INF_F) || gs == 0.0f;           // zero_infinity: no gradient for an impossible alignment
    // frames past the utterance (all of them when dead) get a zero gradient
    {
        const int r0 = dead ? 0 : n;
        float* z = gb + (size_t)r0 * T2max;
        const size_t cnt = (size_t)(T1max - r0) * T2max;
        for (size_t k = lane; k < cnt; k += 32) z[k] = 0.0f;
        if (dead) return;
    }
    const float log2P = l2p[b];                                     // log2 of P * 2^(sum of Z_i - M_i): the scale the variables carry

    float Ha[G], Hb[G];                // log2 of H_{i+1}(label_j), H_{i+1}(blank_j)
#pragma unroll
    for (int g = 0; g < G; ++g) { Ha[g] = kCtcNeg; Hb[g] = kCtcNeg; }
    const int lb = m / G;
#pragma unroll
    for (int g = 0; g < G; ++g) if (lane == lb && g == m - lb * G) Hb[g] = 0.0f;      // virtual frame n: blank_m
    float h1a = Ha[0], h1b = Hb[0], h2a = Ha[0], h2b = Hb[0];      // this lane's first token, one and two steps ago
    extern __shared__ float ctc_smem[];
    constexpr int W = 2 * G + 4;                                    // logits, alpha labels, z (16 B slots)
    float* ring = ctc_smem + (size_t)(threadIdx.x >> 5) * kCtcDepth * 32 * W + lane * W;
    const bool vec = (T2max & 3) == 0 && (reinterpret_cast<uintptr_t>(logits) & 15) == 0;
    auto issue = [&](int step) {
        const int i = ctc_clampi(n - 1 - step + 31 - lane, 0, n - 1);
        const int ta = ctc_clampi(n + 30 - step, 0, n + 30);          // the forward step that produced this step's frames
        float* dst = ring + (step & (kCtcDepth - 1)) * 32 * W;
        const float* src = xb + (size_t)i * T2max;
        const float* asrc = ab + (size_t)ta * (32 * G) + j0;
        if (vec) {
#pragma unroll
            for (int g = 0; g < G; g += 4) ctc_cp16(dst + g, src + min(j0 + g, T2max - 4));
        } else {
#pragma unroll
            for (int g = 0; g < G; ++g) ctc_cp4(dst + g, src + min(j0 + g, T2max - 1));
        }
#pragma unroll
        for (int g = 0; g < G; g += 4) ctc_cp16(dst + G + g, asrc + g);
        ctc_cp8(dst + 2 * G, zb + i);
        cp_async_commit();
    };
    for (int st = 0; st < kCtcDepth - 1; ++st) issue(st);
    const int steps = n + 31;
    for (int t = 0; t < steps; ++t) {
        const int i = n - 1 - t + 31 - lane;
        const float rav = __shfl_down_sync(0xffffffffu, h2a, 1);
        const float rbv = __shfl_down_sync(0xffffffffu, h2b, 1);
        issue(t + kCtcDepth - 1);
        cp_async_wait_pending(kCtcDepth - 1);                     // the copies of step t have landed
        const float* row = ring + (t & (kCtcDepth - 1)) * 32 * W;
        const float zr = row[2 * G], mr = row[2 * G + 1];
        float lp[G], a_cur[G];
#pragma unroll
        for (int g = 0; g < G; ++g) {
            lp[g] = j0 + g < T2max ? fmaf(row[g], kCtcLog2e, -mr) : kCtcNeg;       // relative to the row's largest term
            a_cur[g] = row[G + g];
        }
        const float lpB = blank2 - mr;
        const float dz = mr - zr;                                                   // log2 softmax = lp + dz
        if (i >= 0 && i < n) {
            float gr[G];
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const float na = g + 1 < G ? Ha[g + 1] : (lane == 31 ? kCtcNeg : rav);      // the right neighbour's pair, frame i + 1
                const float nb = g + 1 < G ? Hb[g + 1] : (lane == 31 ? kCtcNeg : rbv);
                const float bt_a = ctc_lse2(Ha[g], ctc_lse2(nb, na));
                const float bt_b = ctc_lse2(Hb[g], Ha[g]);
                // posterior of label_j at frame i = alpha * beta / P (a probability: the exponent is <= 0 up to rounding)
                const float post = ctc_ex2(fminf(a_cur[g] + bt_a - log2P, 1.0f));
                gr[g] = gs * (ctc_ex2(lp[g] + dz) - post);
                Ha[g] = fmaxf((j0 + g < m ? lp[g] : kCtcNeg) + bt_a, kCtcNeg);
                Hb[g] = fmaxf(lpB + bt_b, kCtcNeg);
            }
            float* grow = gb + (size_t)i * T2max + j0;
            if (vec && (reinterpret_cast<uintptr_t>(grad) & 15) == 0) {
#pragma unroll
                for (int g = 0; g < G; g += 4)
                    if (j0 + g < T2max) *reinterpret_cast<float4*>(grow + g) = make_float4(gr[g], gr[g + 1], gr[g + 2], gr[g + 3]);
            } else {
#pragma unroll
                for (int g = 0; g < G; ++g) if (j0 + g < T2max) grow[g] = gr[g];
            }
        }
        h2a = h1a; h2b = h1b;
        h1a = Ha[0]; h1b = Hb[0];
    }
}

// ---- host side -------------------------------------------------------------------------------------------------------
static int ctc_group(int T2max) { return (T2max + 1 + 31) / 32; }          // tokens per lane (the virtual token T2 included)
static int ctc_group_padded(int T2max) {
    const int g = ctc_group(T2max);
    return g <= 4 ? 4 : (g <= 8 ? 8 : (g <= 12 ? 12 : (g <= 16 ? 16 : (g <= 20 ? 20 : 0))));
}

// warps (= utterances) per CTA: bounded by shared memory (the backward ring is 1 KB * (2G + 4) per warp), and no more than it takes to
// give every SM a CTA -- a batch of 256 used to sit on 64 of the 148 SMs
static int ctc_warps_per_cta(int G, int B) { return std::max(1, std::min(G <= 8 ? 4 : 2, (B + 147) / 148)); }

size_t ctc_workspace_bytes(int B, int T1max, int T2max) {
    const int G = ctc_group_padded(T2max);
    if (B <= 0 || T1max <= 0 || T2max <= 0 || G == 0) return 0;
    const size_t rows = size_t(B) * T1max, srows = size_t(B) * (T1max + 31);
    return srows * 32 * G * sizeof(float) + rows * sizeof(float2) + size_t(B) * sizeof(float) + 256;
}

struct CtcWs { float* alpha; float* z2; float* l2p; };
static CtcWs ctc_carve(void* ws, int B, int T1max, int G) {
    const size_t srows = size_t(B) * (T1max + 31);
    CtcWs w;
    w.alpha = static_cast<float*>(ws);
    w.z2 = w.alpha + srows * 32 * G;                       // (B, T1max) float2 {Z_i, M_i}
    w.l2p = w.z2 + size_t(B) * T1max * 2;
    return w;
}

static int ctc_check(const char* who, const float* logits, const int64_t* text_len, const int64_t* mel_len, int B, int T1max, int T2max,
                     const void* ws, size_t ws_bytes) {
    if (!logits || !text_len || !mel_len || !ws) { set_error("%s: null pointer", who); return ISP_ERR_INVALID; }
    if (B <= 0 || T1max <= 0 || T2max <= 0) { set_error("%s: sizes must be positive", who); return ISP_ERR_INVALID; }
    if (ctc_group_padded(T2max) == 0) { set_error("%s: T2max=%d > 639 text tokens is not covered", who, T2max); return ISP_ERR_UNSUPPORTED; }
    if (ws_bytes < ctc_workspace_bytes(B, T1max, T2max)) { set_error("%s: workspace of %zu B required, got %zu", who, ctc_workspace_bytes(B, T1max, T2max), ws_bytes); return ISP_ERR_WORKSPACE; }
    if (reinterpret_cast<uintptr_t>(ws) & 15) { set_error("%s: workspace must be 16 B aligned", who); return ISP_ERR_INVALID; }
    return 0;
}

int ctc_forward(const float* logits, const int64_t* text_len, const int64_t* mel_len, int B, int T1max, int T2max,
                float blank_logprob, float* nll, void* ws, size_t ws_bytes, cudaStream_t stream) {
    int rc = ctc_check("isp_ctc_forward", logits, text_len, mel_len, B, T1max, T2max, ws, ws_bytes);
    if (rc) return rc;
    if (!nll) { set_error("isp_ctc_forward: null pointer"); return ISP_ERR_INVALID; }
    const int G = ctc_group_padded(T2max);
    const CtcWs w = ctc_carve(ws, B, T1max, G);
    const float blank2 = blank_logprob * kCtcLog2e;
    const long long rows = (long long)B * T1max;
    ctc_rownorm_kernel<<<int(std::min<long long>((rows + 7) / 8, 148LL * 16)), 256, 0, stream>>>(logits, mel_len, w.z2, B, T1max, T2max, blank2);
    const int wpc = ctc_warps_per_cta(G, B);
    const int grid = (B + wpc - 1) / wpc;
#define ISP_CTC_ALPHA(GG)                                                                                                   \
    {                                                                                                                        \
        const size_t sm = size_t(wpc) * kCtcDepth * 32 * (GG + 4) * sizeof(float);                                         \
        cudaError_t ea = cudaFuncSetAttribute(ctc_alpha_kernel<GG>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sm));    \
        if (ea != cudaSuccess) return cuda_fail(ea, "cudaFuncSetAttribute(ctc_alpha_kernel)");                               \
        ctc_alpha_kernel<GG><<<grid, 32 * wpc, sm, stream>>>(logits, w.z2, text_len, mel_len, w.alpha, w.l2p, nll, B, T1max, T2max, blank2); \
    }
    switch (G) {
        case 4: ISP_CTC_ALPHA(4) break;
        case 8: ISP_CTC_ALPHA(8) break;
        case 12: ISP_CTC_ALPHA(12) break;
        case 16: ISP_CTC_ALPHA(16) break;
        default: ISP_CTC_ALPHA(20) break;
    }
#undef ISP_CTC_ALPHA
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "ctc forward launch");
    return 0;
}

int ctc_backward(const float* logits, const int64_t* text_len, const int64_t* mel_len, int B, int T1max, int T2max,
                 float blank_logprob, const float* nll, const float* grad_scale, float* grad_logits, void* ws, size_t ws_bytes,
                 cudaStream_t stream) {
    int rc = ctc_check("isp_ctc_backward", logits, text_len, mel_len, B, T1max, T2max, ws, ws_bytes);
    if (rc) return rc;
    if (!nll || !grad_scale || !grad_logits) { set_error("isp_ctc_backward: null pointer"); return ISP_ERR_INVALID; }
    const int G = ctc_group_padded(T2max);
    const CtcWs w = ctc_carve(ws, B, T1max, G);
    const float blank2 = blank_logprob * kCtcLog2e;
    const int wpc = ctc_warps_per_cta(G, B);
    const int grid = (B + wpc - 1) / wpc;
#define ISP_CTC_BETA(GG)                                                                                                    \
    {                                                                                                                        \
        const size_t sm = size_t(wpc) * kCtcDepth * 32 * (2 * GG + 4) * sizeof(float);                                     \
        cudaError_t ea = cudaFuncSetAttribute(ctc_beta_grad_kernel<GG>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sm)); \
        if (ea != cudaSuccess) return cuda_fail(ea, "cudaFuncSetAttribute(ctc_beta_grad_kernel)");                            \
        ctc_beta_grad_kernel<GG><<<grid, 32 * wpc, sm, stream>>>(logits, w.z2, text_len, mel_len, w.alpha, w.l2p, nll, grad_scale, \
                                                                 grad_logits, B, T1max, T2max, blank2);                       \
    }
    switch (G) {
        case 4: ISP_CTC_BETA(4) break;
        case 8: ISP_CTC_BETA(8) break;
        case 12: ISP_CTC_BETA(12) break;
        case 16: ISP_CTC_BETA(16) break;
        default: ISP_CTC_BETA(20) break;
    }
#undef ISP_CTC_BETA
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "ctc backward launch");
    return 0;
}

}  // namespace isp
