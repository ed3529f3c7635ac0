// The log-likelihood epilogue as a stand-alone row kernel: the general path behind shapes the fused kernel (isp_loglik.cu) does
// not cover -- more than 512 text tokens (its accumulator is one TMEM allocation), attention_dim not a multiple of 8 or above 256.
// The scores come from isp_gemm_batched (any shape); this kernel does alignment.py:190-208 of the reference on them:
//   scale, log_softmax over ALL T2max columns (:196), + log(diagonal prior + 1e-6) (:18-37, :196), masked softmax (:201-206).
// One warp per frame row, the row re-read from L1/L2 in five short sweeps (max, sum, logits + masked max, sum, soft), precise
// expf / logf / divisions: 4 B/cell read from DRAM, 8 B/cell written.  A fallback: correctness first, then HBM.

#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include <algorithm>

#include "common.cuh"
#include "isp_internal.h"

namespace isp {

namespace {

ISP_DEVINL float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
ISP_DEVINL float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(256)
loglik_rows_kernel(const float* __restrict__ S, const int64_t* __restrict__ text_len, const int64_t* __restrict__ mel_len,
                   float* __restrict__ logits, float* __restrict__ soft, long long rows, int T1max, int T2max, long long ldS,
                   float scale, int prior) {
    const int lane = threadIdx.x & 31;
    const long long wstride = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows; row += wstride) {
        const int b = int(row / T1max), i = int(row - (long long)b * T1max);
        long long n64 = mel_len[b], m64 = text_len[b];
        const int T1b = int(n64 < 1 ? 1 : (n64 > T1max ? T1max : n64));
        const int T2b = int(m64 < 1 ? 1 : (m64 > T2max ? T2max : m64));
        const bool row_valid = i < T1b;
        const float* s = S + row * ldS;
        float* lg = logits + row * (long long)T2max;
        float* sf = soft + row * (long long)T2max;

        float lse = 0.0f, pnorm = 0.0f;
        const float gm = __fdiv_rn(float(i), float(T1b));                              // alignment.py:24-25
        const float t2f = float(T2b);
        if (prior) {
            float mx = -CUDART_INF_F;
            for (int j = lane; j < T2max; j += 32) mx = fmaxf(mx, fminf(scale * s[j], 3.4028234663852886e38f));
            mx = warp_max(mx);
            float sum = 0.0f, psum = 0.0f;
            for (int j = lane; j < T2max; j += 32) {
                sum += expf(fminf(scale * s[j], 3.4028234663852886e38f) - mx);
                if (row_valid && j < T2b) {
                    const float g = __fdiv_rn(float(j), t2f) - gm;                      // :21-22, :27
                    psum += expf(__fdiv_rn(-(g * g), 0.02f));                           // :29 with gamma = 0.1
                }
            }
            lse = mx + logf(warp_sum(sum));
            pnorm = warp_sum(psum) + 1e-5f;                                             // :34
        }
        float vmax = -CUDART_INF_F;
        for (int j = lane; j < T2max; j += 32) {
            float v = fminf(scale * s[j], 3.4028234663852886e38f);                     // :190, :192
            if (prior) {
                float p = 0.0f;
                if (row_valid && j < T2b) {
                    const float g = __fdiv_rn(float(j), t2f) - gm;
                    p = __fdiv_rn(expf(__fdiv_rn(-(g * g), 0.02f)), pnorm);
                    if (p < 1e-4f) p = 0.0f;                                            // :35
                }
                v = (v - lse) + logf(p + 1e-6f);                                        // :196
            }
            lg[j] = v;                                                                  // :198
            if (j < T2b) vmax = fmaxf(vmax, v);
        }
        vmax = warp_max(vmax);
        float wsum = 0.0f;
        for (int j = lane; j < T2b; j += 32) wsum += expf(lg[j] - vmax);                // the lane re-reads what it wrote
        wsum = warp_sum(wsum);
        for (int j = lane; j < T2max; j += 32)
            sf[j] = (row_valid && j < T2b) ? __fdiv_rn(expf(lg[j] - vmax), wsum) : 0.0f;   // :201-206
    }
}

// x = hi + lo with hi = x rounded to TF32 (10-bit mantissa) and lo = the rounded remainder: hi*hi' + hi*lo' + lo*hi' reproduces the
// fp32 product to ~2^-21 (the dropped lo*lo' term is 2^-22 relative), which the tensor cores accumulate in fp32.  The three terms
// become ONE contraction over 3 D: the frame side is laid out [hi | hi | lo], the token side [hi | lo | hi].
__global__ void __launch_bounds__(256)
split_3xtf32_kernel(const float* __restrict__ x, float* __restrict__ out, long long rows, int D, int role) {
    const long long total = rows * D;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / D;
        const int d = int(i - r * D);
        const float v = x[i];
        uint32_t hb, lb;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(v));
        const float hi = __uint_as_float(hb);
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lb) : "f"(v - hi));
        const float lo = __uint_as_float(lb);
        float* o = out + r * 3 * D + d;
        o[0] = hi;
        o[D] = role == 0 ? hi : lo;
        o[2 * D] = role == 0 ? lo : hi;
    }
}

}  // namespace

int split_3xtf32(const float* x, int64_t rows, int D, int role, float* out, cudaStream_t stream) {
    if (!x || !out) { set_error("isp_split_3xtf32: null pointer"); return ISP_ERR_INVALID; }
    if (rows <= 0 || D <= 0 || (role != 0 && role != 1)) { set_error("isp_split_3xtf32: rows and D must be positive, role 0 (frames) or 1 (tokens)"); return ISP_ERR_INVALID; }
    const long long total = (long long)rows * D;
    const int grid = int(std::min<long long>((total + 255) / 256, 148LL * 16));
    split_3xtf32_kernel<<<grid, 256, 0, stream>>>(x, out, rows, D, role);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "split_3xtf32_kernel launch");
    return 0;
}

int loglik_rows(const float* S, int64_t ldS, const int64_t* text_len, const int64_t* mel_len, int B, int T1max, int T2max,
                float scale, int attention_prior, float* attn_logits, float* attn_soft, cudaStream_t stream) {
    if (!S || !text_len || !mel_len || !attn_logits || !attn_soft) { set_error("isp_loglik_rows: null pointer"); return ISP_ERR_INVALID; }
    if (B <= 0 || T1max <= 0 || T2max <= 0 || ldS < T2max) { set_error("isp_loglik_rows: sizes must be positive and ldS >= T2max"); return ISP_ERR_INVALID; }
    if (!(scale > 0.0f)) { set_error("isp_loglik_rows: scale must be positive"); return ISP_ERR_INVALID; }
    const long long rows = (long long)B * T1max;
    const int grid = int(std::min<long long>((rows + 7) / 8, 148LL * 8 * 4));
    loglik_rows_kernel<<<grid, 256, 0, stream>>>(S, text_len, mel_len, attn_logits, attn_soft, rows, T1max, T2max, ldS, scale, attention_prior ? 1 : 0);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "loglik_rows_kernel launch");
    return 0;
}

}  // namespace isp
