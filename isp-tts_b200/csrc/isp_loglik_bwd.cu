// Backward of the fused log-likelihood, first stage: the score gradient dS.
//
// Reference semantics (tts/models/acoustic/modules/alignment.py in the reference), differentiated:
//   :196  attn_logits = log_softmax(scale * S, over ALL T2max columns) + log(prior + 1e-6)      (prior: no gradient)
//   :201-206  attn_soft = softmax(attn_logits over the valid text columns) * mask
// With g_l = dL/d attn_logits and g_s = dL/d attn_soft, per frame row i:
//   g_ij  = g_l_ij + soft_ij * (g_s_ij - sum_j g_s_ij soft_ij)             (softmax Jacobian; soft is 0 where masked)
//   dS_ij = scale * (g_ij - softmax_all(scale * S)_ij * sum_j g_ij)         (log-softmax Jacobian)
// and without the prior branch (attention_prior = False, attn_logits = scale * S):  dS_ij = scale * g_ij.
// dQ = dS . K and dK = dS^T . Q are plain batched GEMMs and stay with the library (cuBLAS through torch).
//
// One warp per row, the whole row in registers: every input is read once (16 B/cell), dS written once.
// The kernel is HBM-bound by construction; bench.py reports its GB/s next to the forward kernels.

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "common.cuh"
#include "isp_internal.h"
#include "isp_prior.cuh"

namespace isp {

constexpr int kBwdWarps = 8;
constexpr float kBwdLog2e = 1.4426950408889634f;

ISP_DEVINL float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
ISP_DEVINL float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
ISP_DEVINL float fast_rcp(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
ISP_DEVINL float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// NV: float4 per lane (T2max <= 128 * NV, T2max % 4 == 0).  OUT_BF16: dS as bf16 (the GEMMs that follow run in bf16).
template <int NV, bool OUT_BF16>
__global__ void __launch_bounds__(32 * kBwdWarps)
loglik_bwd_ds_kernel(const float* __restrict__ S, const float* __restrict__ soft, const float* __restrict__ g_logits,
                     const float* __restrict__ g_soft, void* __restrict__ dS, long long rows, int T2max, float scale, int prior) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * kBwdWarps + (threadIdx.x >> 5);
    if (row >= rows) return;
    const size_t base = size_t(row) * T2max;
    const int n4 = T2max >> 2;
    float4 s[NV], a[NV], gl[NV], gs[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = lane + 32 * i;
        const bool ok = c < n4;
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        s[i] = ok ? __ldcs(reinterpret_cast<const float4*>(S + base) + c) : make_float4(-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F);
        a[i] = ok ? __ldcs(reinterpret_cast<const float4*>(soft + base) + c) : z;
        gl[i] = (ok && g_logits) ? __ldcs(reinterpret_cast<const float4*>(g_logits + base) + c) : z;
        gs[i] = (ok && g_soft) ? __ldcs(reinterpret_cast<const float4*>(g_soft + base) + c) : z;
    }
    float dot = 0.f, gsum = 0.f, asum = 0.f, m = -CUDART_INF_F;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        dot += gs[i].x * a[i].x + gs[i].y * a[i].y + gs[i].z * a[i].z + gs[i].w * a[i].w;
        gsum += gl[i].x + gl[i].y + gl[i].z + gl[i].w;
        asum += a[i].x + a[i].y + a[i].z + a[i].w;
        m = fmaxf(m, fmaxf(fmaxf(s[i].x, s[i].y), fmaxf(s[i].z, s[i].w)));
    }
    dot = warp_sum(dot); gsum = warp_sum(gsum); asum = warp_sum(asum); m = warp_max(m);
    const float G = gsum + dot * (1.0f - asum);         // = sum_j g_ij
    const float c2 = scale * kBwdLog2e;
    float e[NV][4], esum = 0.f;
    if (prior) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            e[i][0] = ex2f((s[i].x - m) * c2); e[i][1] = ex2f((s[i].y - m) * c2);
            e[i][2] = ex2f((s[i].z - m) * c2); e[i][3] = ex2f((s[i].w - m) * c2);
            esum += e[i][0] + e[i][1] + e[i][2] + e[i][3];                   // columns past T2max hold -inf: exp = 0
        }
        esum = warp_sum(esum);
    }
    const float k = prior ? G / esum : 0.0f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = lane + 32 * i;
        if (c >= n4) continue;
        float d[4];
        const float av[4] = {a[i].x, a[i].y, a[i].z, a[i].w}, glv[4] = {gl[i].x, gl[i].y, gl[i].z, gl[i].w}, gsv[4] = {gs[i].x, gs[i].y, gs[i].z, gs[i].w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float g = glv[q] + av[q] * (gsv[q] - dot);
            d[q] = scale * (prior ? g - e[i][q] * k : g);
        }
        if (OUT_BF16) {
            __nv_bfloat162 lo = __floats2bfloat162_rn(d[0], d[1]), hi = __floats2bfloat162_rn(d[2], d[3]);
            uint2 pk;
            pk.x = *reinterpret_cast<uint32_t*>(&lo);
            pk.y = *reinterpret_cast<uint32_t*>(&hi);
            *(reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(dS) + base) + c) = pk;
        } else {
            *(reinterpret_cast<float4*>(reinterpret_cast<float*>(dS) + base) + c) = make_float4(d[0], d[1], d[2], d[3]);
        }
    }
}

template <int NV>
static void launch_bwd(const float* S, const float* soft, const float* gl, const float* gs, void* dS, int out_bf16,
                       long long rows, int T2max, float scale, int prior, cudaStream_t stream) {
    const unsigned grid = unsigned((rows + kBwdWarps - 1) / kBwdWarps);
    if (out_bf16) loglik_bwd_ds_kernel<NV, true><<<grid, 32 * kBwdWarps, 0, stream>>>(S, soft, gl, gs, dS, rows, T2max, scale, prior);
    else loglik_bwd_ds_kernel<NV, false><<<grid, 32 * kBwdWarps, 0, stream>>>(S, soft, gl, gs, dS, rows, T2max, scale, prior);
}

int loglik_backward_ds(const float* S, const float* attn_soft, const float* g_logits, const float* g_soft,
                       int B, int T1max, int T2max, float scale, int attention_prior, void* dS, int ds_dtype, cudaStream_t stream) {
    if (!S || !attn_soft || !dS || (!g_logits && !g_soft)) { set_error("isp_loglik_backward_ds: null pointer (at least one incoming gradient is needed)"); return ISP_ERR_INVALID; }
    if (B <= 0 || T1max <= 0 || T2max <= 0) { set_error("isp_loglik_backward_ds: sizes must be positive"); return ISP_ERR_INVALID; }
    if (ds_dtype != ISP_DTYPE_F32 && ds_dtype != ISP_DTYPE_BF16) { set_error("isp_loglik_backward_ds: ds_dtype must be ISP_DTYPE_F32 or ISP_DTYPE_BF16"); return ISP_ERR_INVALID; }
    if (T2max % 4 != 0 || T2max > ISP_LOGLIK_MAX_T2) { set_error("isp_loglik_backward_ds: T2max=%d must be a multiple of 4 and <= %d (pad the token axis)", T2max, ISP_LOGLIK_MAX_T2); return ISP_ERR_UNSUPPORTED; }
    const uintptr_t al = reinterpret_cast<uintptr_t>(S) | reinterpret_cast<uintptr_t>(attn_soft) | reinterpret_cast<uintptr_t>(g_logits) |
                         reinterpret_cast<uintptr_t>(g_soft) | reinterpret_cast<uintptr_t>(dS);
    if (al & 15) { set_error("isp_loglik_backward_ds: all tensors must be 16 B aligned"); return ISP_ERR_INVALID; }
    const long long rows = (long long)B * T1max;
    const int nv = (T2max / 4 + 31) / 32;
    switch (nv) {
        case 1: launch_bwd<1>(S, attn_soft, g_logits, g_soft, dS, ds_dtype == ISP_DTYPE_BF16, rows, T2max, scale, attention_prior, stream); break;
        case 2: launch_bwd<2>(S, attn_soft, g_logits, g_soft, dS, ds_dtype == ISP_DTYPE_BF16, rows, T2max, scale, attention_prior, stream); break;
        case 3: launch_bwd<3>(S, attn_soft, g_logits, g_soft, dS, ds_dtype == ISP_DTYPE_BF16, rows, T2max, scale, attention_prior, stream); break;
        default: launch_bwd<4>(S, attn_soft, g_logits, g_soft, dS, ds_dtype == ISP_DTYPE_BF16, rows, T2max, scale, attention_prior, stream); break;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "loglik_bwd_ds_kernel launch");
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------------
// The same gradient WITHOUT the scores (isp_loglik_backward_from_logits).  Everything the Jacobians need is a function of the
// forward's own output attn_logits and of the prior, which is closed-form:
//   softmax_all(scale * S)_ij = exp(attn_logits_ij) / (prior_ij + 1e-6)          (alignment.py:196 read backwards)
//   attn_soft_ij              = softmax over the valid tokens of attn_logits_ij, 0 on padding   (alignment.py:201-206)
// so neither the score GEMM is recomputed (64 us and a 205 MB round trip at batch 256 x 1000 x 200) nor attn_soft read:
// 12 B/cell in instead of 16, one launch less.  The prior's cells are re-derived exactly as the forward kernel computed them
// (isp_prior.cuh) from the row sums it saved -- the 1e-4 threshold is a comparison, and a cell that fell on the other side of
// it here would be off by a factor of 101.  One warp per frame row, a CTA's rows inside one utterance (the j / T2_b table
// is per utterance).
template <int NV, bool OUT_BF16>
#ifndef ISP_BWD_MINB
#define ISP_BWD_MINB 3
#endif
__global__ void __launch_bounds__(32 * kBwdWarps, ISP_BWD_MINB)
loglik_bwd_logits_kernel(const float* __restrict__ logits, const float* __restrict__ g_logits, const float* __restrict__ g_soft,
                         const float* __restrict__ psum, const int64_t* __restrict__ text_len, const int64_t* __restrict__ mel_len,
                         void* __restrict__ dS, int T1max, int T2max, float scale, int prior) {
    __shared__ __align__(16) float gt[ISP_LOGLIK_MAX_T2];
    const int lane = threadIdx.x & 31, b = blockIdx.y;
    const int i = blockIdx.x * kBwdWarps + (threadIdx.x >> 5);
    const bool row_exists = i < T1max;
    const size_t base = (size_t(b) * T1max + (row_exists ? i : 0)) * T2max;
    const int n4 = T2max >> 2;
    // the row's loads go out first: the lengths and the j / T2_b table (a dependent global load, divisions, a barrier) are set
    // up while they are in flight
    float4 l[NV], gl[NV], gs[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        const int c = lane + 32 * v;
        const bool ok = c < n4 && row_exists;
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        l[v] = ok ? __ldcs(reinterpret_cast<const float4*>(logits + base) + c) : z;
        gl[v] = (ok && g_logits) ? __ldcs(reinterpret_cast<const float4*>(g_logits + base) + c) : z;
        gs[v] = (ok && g_soft) ? __ldcs(reinterpret_cast<const float4*>(g_soft + base) + c) : z;
    }
    const long long n64 = mel_len[b], m64 = text_len[b];
    const int T1b = int(n64 < 1 ? 1 : (n64 > T1max ? T1max : n64));       // clamped as the forward kernel clamps them
    const int T2b = int(m64 < 1 ? 1 : (m64 > T2max ? T2max : m64));
    float psum_row = 0.0f;
    const bool row_valid = i < T1b;
    if (prior) {
        if (row_valid) psum_row = psum[size_t(b) * T1max + i];
        // padded tokens sit far off the diagonal (their raw prior is exp2(-1e6) == 0: below the threshold without a per-cell mask)
        const float t2f = float(T2b);
        const int j0 = threadIdx.x, j1 = threadIdx.x + 32 * kBwdWarps;         // T2max <= 2 * 256
        if (j0 < T2max) gt[j0] = j0 < T2b ? prior_grid(j0, t2f) : 1.0e3f * kPriorScale;
        if (j1 < T2max) gt[j1] = j1 < T2b ? prior_grid(j1, t2f) : 1.0e3f * kPriorScale;
    }
    __syncthreads();
    if (!row_exists) return;
    // ---- one exponential per cell serves both Jacobians: e_j = exp(l_j - mx), mx the row's maximum over the valid tokens.
    //      attn_soft_j = e_j / sum over valid tokens (0 on padding);  softmax_all_j = e_j * exp(mx) / P_j ----
    int nval[NV];                                        // valid tokens among this lane's four columns
    float mx = -CUDART_INF_F, gsum = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        nval[v] = min(4, max(0, T2b - 4 * (lane + 32 * v)));
        const float lv[4] = {l[v].x, l[v].y, l[v].z, l[v].w};
#pragma unroll
        for (int q = 0; q < 4; ++q) if (q < nval[v]) mx = fmaxf(mx, lv[q]);
        gsum += (gl[v].x + gl[v].y) + (gl[v].z + gl[v].w);
    }
    mx = warp_max(mx);                                   // (T2b >= 1: finite)
    const float mxc = mx * kLog2e;
    float e[NV][4], ev[NV][4];
    float z = 0.f, dotz = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        const float lv[4] = {l[v].x, l[v].y, l[v].z, l[v].w}, gv[4] = {gs[v].x, gs[v].y, gs[v].z, gs[v].w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            e[v][q] = fast_ex2(fmaf(lv[q], kLog2e, -mxc));
            ev[v][q] = q < nval[v] ? e[v][q] : 0.0f;
            z += ev[v][q];
            dotz = fmaf(gv[q], ev[v][q], dotz);
        }
    }
    gsum = warp_sum(gsum); z = warp_sum(z); dotz = warp_sum(dotz);
    // a padded frame has attn_soft == 0 (alignment.py:206): no softmax term at all
    const bool use_soft = g_soft != nullptr && row_valid;
    const float inv_z = use_soft ? fast_rcp(z) : 0.0f;
    const float dot = dotz * inv_z;                      // = sum_j g_s_ij soft_ij
    const float asum = use_soft ? 1.0f : 0.0f;           // sum_j soft_ij
    const float G = gsum + dot * (1.0f - asum);          // = sum_j g_ij
    float inv_psum = 0.0f, us = 0.0f, kG = 0.0f;
    if (prior) {
        // alignment.py:34, as the forward divides; a padded frame has no prior at all (0 stays below the threshold)
        inv_psum = row_valid ? 1.0f / (psum_row + 1e-5f) : 0.0f;
        us = prior_grid(i, float(T1b));
        kG = fast_ex2(mxc) * G;                          // softmax_all_j * G = e_j * kG / P_j
    }
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        const int c = lane + 32 * v;
        if (c >= n4) continue;
        const float glv[4] = {gl[v].x, gl[v].y, gl[v].z, gl[v].w}, gsv[4] = {gs[v].x, gs[v].y, gs[v].z, gs[v].w};
        float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (prior) g4 = *reinterpret_cast<const float4*>(gt + 4 * c);
        const float gtv[4] = {g4.x, g4.y, g4.z, g4.w};
        float d[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float g = fmaf(ev[v][q] * inv_z, gsv[q] - dot, glv[q]);
            if (prior) {
                const float P = prior_cell_P(gtv[q], us, inv_psum, true);
                d[q] = scale * fmaf(-e[v][q] * kG, fast_rcp(P), g);
            } else {
                d[q] = scale * g;
            }
        }
        if (OUT_BF16) {
            __nv_bfloat162 lo = __floats2bfloat162_rn(d[0], d[1]), hi = __floats2bfloat162_rn(d[2], d[3]);
            uint2 pk;
            pk.x = *reinterpret_cast<uint32_t*>(&lo);
            pk.y = *reinterpret_cast<uint32_t*>(&hi);
            *(reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(dS) + base) + c) = pk;
        } else {
            *(reinterpret_cast<float4*>(reinterpret_cast<float*>(dS) + base) + c) = make_float4(d[0], d[1], d[2], d[3]);
        }
    }
}

template <int NV>
static void launch_bwd_logits(const float* logits, const float* gl, const float* gs, const float* psum, const int64_t* tl, const int64_t* ml,
                              void* dS, int out_bf16, int B, int T1max, int T2max, float scale, int prior, cudaStream_t stream) {
    const dim3 grid(unsigned((T1max + kBwdWarps - 1) / kBwdWarps), unsigned(B));
    if (out_bf16) loglik_bwd_logits_kernel<NV, true><<<grid, 32 * kBwdWarps, 0, stream>>>(logits, gl, gs, psum, tl, ml, dS, T1max, T2max, scale, prior);
    else loglik_bwd_logits_kernel<NV, false><<<grid, 32 * kBwdWarps, 0, stream>>>(logits, gl, gs, psum, tl, ml, dS, T1max, T2max, scale, prior);
}

int loglik_backward_from_logits(const float* attn_logits, const float* g_logits, const float* g_soft, const float* prior_rowsum,
                                const int64_t* text_len, const int64_t* mel_len, int B, int T1max, int T2max, float scale,
                                int attention_prior, void* dS, int ds_dtype, cudaStream_t stream) {
    if (!attn_logits || !text_len || !mel_len || !dS || (!g_logits && !g_soft)) { set_error("isp_loglik_backward_from_logits: null pointer (at least one incoming gradient is needed)"); return ISP_ERR_INVALID; }
    if (attention_prior && !prior_rowsum) { set_error("isp_loglik_backward_from_logits: the prior's row sums (isp_loglik_forward's workspace) are needed with attention_prior"); return ISP_ERR_INVALID; }
    if (B <= 0 || B > 65535 || T1max <= 0 || T2max <= 0) { set_error("isp_loglik_backward_from_logits: sizes must be positive, B <= 65535"); return ISP_ERR_INVALID; }
    if (ds_dtype != ISP_DTYPE_F32 && ds_dtype != ISP_DTYPE_BF16) { set_error("isp_loglik_backward_from_logits: ds_dtype must be ISP_DTYPE_F32 or ISP_DTYPE_BF16"); return ISP_ERR_INVALID; }
    if (T2max % 4 != 0 || T2max > ISP_LOGLIK_MAX_T2) { set_error("isp_loglik_backward_from_logits: T2max=%d must be a multiple of 4 and <= %d", T2max, ISP_LOGLIK_MAX_T2); return ISP_ERR_UNSUPPORTED; }
    const uintptr_t al = reinterpret_cast<uintptr_t>(attn_logits) | reinterpret_cast<uintptr_t>(g_logits) | reinterpret_cast<uintptr_t>(g_soft) | reinterpret_cast<uintptr_t>(dS);
    if (al & 15) { set_error("isp_loglik_backward_from_logits: all tensors must be 16 B aligned"); return ISP_ERR_INVALID; }
    const int ob = ds_dtype == ISP_DTYPE_BF16;
    switch ((T2max / 4 + 31) / 32) {
        case 1: launch_bwd_logits<1>(attn_logits, g_logits, g_soft, prior_rowsum, text_len, mel_len, dS, ob, B, T1max, T2max, scale, attention_prior, stream); break;
        case 2: launch_bwd_logits<2>(attn_logits, g_logits, g_soft, prior_rowsum, text_len, mel_len, dS, ob, B, T1max, T2max, scale, attention_prior, stream); break;
        case 3: launch_bwd_logits<3>(attn_logits, g_logits, g_soft, prior_rowsum, text_len, mel_len, dS, ob, B, T1max, T2max, scale, attention_prior, stream); break;
        default: launch_bwd_logits<4>(attn_logits, g_logits, g_soft, prior_rowsum, text_len, mel_len, dS, ob, B, T1max, T2max, scale, attention_prior, stream); break;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "loglik_bwd_logits_kernel launch");
    return 0;
}

}  // namespace isp
