// Element-wise companions of the convolution GEMMs of the projection stacks (SURVEY.md section 8, row f-2;
// reference tts/models/acoustic/modules/alignment.py:40-83,118-154,176-187 and tts/modules/normalization.py:160-208).
//
//   prep_channels_last   mel (B, C, T) / encoded text (B, C, T) or (B, T, C), fp32 or bf16  ->  (B, T, Cp) in the GEMM's
//                        operand type, masked at t >= len (ConvBlock1D masks its input first, alignment.py:75-76), channel
//                        padding Cp - C zero-filled: the K-major operand the implicit-GEMM convolution loads with TMA.
//   instance_norm_apply  masked instance norm (normalization.py:186-206) of a channels-last activation from the column sums
//                        the GEMM's epilogue left (sum y, sum y^2 per 32-row slab over valid frames): mean, biased variance,
//                        (y - mean) / sqrt(var + eps) * weight + bias, re-masked for the next block.
// Both are one read and one write of the activation (HBM-bound, 16 B per lane).

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>

#include "common.cuh"
#include "isp_internal.h"

namespace isp {

namespace {

template <typename T> ISP_DEVINL float to_f(T v);
template <> ISP_DEVINL float to_f<float>(float v) { return v; }
template <> ISP_DEVINL float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> ISP_DEVINL float to_f<__half>(__half v) { return __half2float(v); }
template <typename T> ISP_DEVINL T from_f(float v);
template <> ISP_DEVINL float from_f<float>(float v) { return v; }
template <> ISP_DEVINL __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> ISP_DEVINL __half from_f<__half>(float v) { return __float2half_rn(v); }

// channels-first input: a 32 (t) x 32 (c) tile through shared memory so that both sides are coalesced
template <typename TI, typename TO>
__global__ void __launch_bounds__(256)
prep_cf_kernel(const TI* __restrict__ x, const int64_t* __restrict__ len, TO* __restrict__ out, int C, int T, int Cp) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z, t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 32 x 8
    long long n = len ? len[b] : T;
    const int valid = int(n < 0 ? 0 : (n > T ? T : n));
    const TI* xb = x + size_t(b) * C * T;
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int c = c0 + r, t = t0 + tx;
        tile[r][tx] = (c < C && t < valid) ? to_f<TI>(xb[size_t(c) * T + t]) : 0.0f;
    }
    __syncthreads();
    TO* ob = out + size_t(b) * T * Cp;
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int t = t0 + r, c = c0 + tx;
        if (t < T && c < Cp) ob[size_t(t) * Cp + c] = from_f<TO>(tile[tx][r]);
    }
}

// channels-last input: cast + mask + channel padding
template <typename TI, typename TO>
__global__ void __launch_bounds__(256)
prep_cl_kernel(const TI* __restrict__ x, const int64_t* __restrict__ len, TO* __restrict__ out, int C, int T, int Cp, long long total) {
    for (long long idx = (long long)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (long long)gridDim.x * 256) {
        const int c = int(idx % Cp);
        const long long bt = idx / Cp;
        const int t = int(bt % T);
        const long long b = bt / T;
        long long n = len ? len[b] : T;
        out[idx] = from_f<TO>((c < C && t < n) ? to_f<TI>(x[(b * T + t) * C + c]) : 0.0f);
    }
}

template <typename TI, typename TO>
int prep_launch(const void* x, const int64_t* len, void* out, int channels_first, int B, int C, int T, int Cp, cudaStream_t stream) {
    if (channels_first) {
        const dim3 grid((T + 31) / 32, (Cp + 31) / 32, B);
        prep_cf_kernel<TI, TO><<<grid, 256, 0, stream>>>(static_cast<const TI*>(x), len, static_cast<TO*>(out), C, T, Cp);
    } else {
        const long long total = (long long)B * T * Cp;
        prep_cl_kernel<TI, TO><<<int(std::min<long long>((total + 255) / 256, 148 * 16)), 256, 0, stream>>>(
            static_cast<const TI*>(x), len, static_cast<TO*>(out), C, T, Cp, total);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "prep_channels_last kernel launch");
    return 0;
}

// One CTA = 32 frames x all channels of one utterance.  Per channel: add the slabs' partial sums in slab order, mean and
// rstd in fp32, then the CTA's rows: 8 channels (16 B of bf16) or 4 channels (16 B of fp32) per thread.
template <typename T>
__global__ void __launch_bounds__(256)
instance_norm_kernel(const T* __restrict__ y, const float* __restrict__ stats, const float* __restrict__ weight,
                     const float* __restrict__ bias, const int64_t* __restrict__ len, T* __restrict__ out,
                     int Tmax, int C, int ld_in, int ld_out, int parts, float eps) {
    extern __shared__ float2 s_ab[];                 // per channel: scale a = rstd * w, shift d = bias - mean * a
    const int b = blockIdx.y, t0 = blockIdx.x * 32;
    long long n64 = len ? len[b] : Tmax;
    const int n = int(n64 < 0 ? 0 : (n64 > Tmax ? Tmax : n64));
    const float inv_n = n > 0 ? 1.0f / float(n) : 0.0f;
    const float* st = stats + size_t(b) * parts * C * 2;
    for (int c = threadIdx.x; c < C; c += 256) {
        float s = 0.f, q = 0.f;
        for (int p = 0; p < parts; ++p) {
            const float2 v = *reinterpret_cast<const float2*>(st + (size_t(p) * C + c) * 2);
            s += v.x; q += v.y;
        }
        const float mean = s * inv_n;
        const float var = fmaxf(q * inv_n - mean * mean, 0.0f);
        const float a = rsqrtf(var + eps) * (weight ? weight[c] : 1.0f);
        s_ab[c] = make_float2(a, (bias ? bias[c] : 0.0f) - mean * a);
    }
    __syncthreads();
    constexpr int V = 16 / sizeof(T);               // channels per 16 B
    const int vec_per_row = C / V;                   // host checks C % V == 0
    const int rows = min(32, Tmax - t0);
    for (int idx = threadIdx.x; idx < rows * vec_per_row; idx += 256) {
        const int r = idx / vec_per_row, v = idx - r * vec_per_row;
        const int t = t0 + r;
        T vals[V];
        if (t < n) {
            *reinterpret_cast<uint4*>(vals) = __ldcs(reinterpret_cast<const uint4*>(y + (size_t(b) * Tmax + t) * ld_in + v * V));
#pragma unroll
            for (int k = 0; k < V; ++k) {
                const float2 ab = s_ab[v * V + k];
                vals[k] = from_f<T>(fmaf(to_f<T>(vals[k]), ab.x, ab.y));
            }
        } else {
#pragma unroll
            for (int k = 0; k < V; ++k) vals[k] = from_f<T>(0.0f);
        }
        *reinterpret_cast<uint4*>(out + (size_t(b) * Tmax + t) * ld_out + v * V) = *reinterpret_cast<uint4*>(vals);
    }
}

}  // namespace

int prep_channels_last(const void* x, int in_dtype, int channels_first, const int64_t* len, void* out, int out_dtype,
                       int B, int C, int T, int Cp, cudaStream_t stream) {
    if (!x || !out) { set_error("isp_prep_channels_last: null pointer"); return ISP_ERR_INVALID; }
    if (B <= 0 || C <= 0 || T <= 0 || Cp < C) { set_error("isp_prep_channels_last: sizes must be positive and Cp >= C"); return ISP_ERR_INVALID; }
    if (B > 65535) { set_error("isp_prep_channels_last: B=%d > 65535", B); return ISP_ERR_UNSUPPORTED; }
    if (in_dtype != ISP_DTYPE_F32 && in_dtype != ISP_DTYPE_BF16) { set_error("isp_prep_channels_last: the input must be fp32 or bf16"); return ISP_ERR_INVALID; }
    const bool ib = in_dtype == ISP_DTYPE_BF16;
    switch (out_dtype) {
        case ISP_DTYPE_F32:
            return ib ? prep_launch<__nv_bfloat16, float>(x, len, out, channels_first, B, C, T, Cp, stream)
                      : prep_launch<float, float>(x, len, out, channels_first, B, C, T, Cp, stream);
        case ISP_DTYPE_BF16:
            return ib ? prep_launch<__nv_bfloat16, __nv_bfloat16>(x, len, out, channels_first, B, C, T, Cp, stream)
                      : prep_launch<float, __nv_bfloat16>(x, len, out, channels_first, B, C, T, Cp, stream);
        case ISP_DTYPE_F16:
            return ib ? prep_launch<__nv_bfloat16, __half>(x, len, out, channels_first, B, C, T, Cp, stream)
                      : prep_launch<float, __half>(x, len, out, channels_first, B, C, T, Cp, stream);
        default:
            set_error("isp_prep_channels_last: bad output dtype"); return ISP_ERR_INVALID;
    }
}

template <typename T>
static int norm_launch(const void* y, const float* stats, const float* weight, const float* bias, const int64_t* len, void* out,
                       int B, int T_, int C, int64_t ld_in, int64_t ld_out, int parts, float eps, cudaStream_t stream) {
    const dim3 grid((T_ + 31) / 32, B);
    const size_t smem = size_t(C) * sizeof(float2);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(instance_norm_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(instance_norm_kernel)");
    }
    instance_norm_kernel<T><<<grid, 256, smem, stream>>>(static_cast<const T*>(y), stats, weight, bias, len, static_cast<T*>(out), T_, C,
                                                         int(ld_in), int(ld_out), parts, eps);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "instance_norm_kernel launch");
    return 0;
}

int instance_norm_apply(const void* y, int dtype, const float* stats, int parts, const float* weight, const float* bias,
                        const int64_t* len, void* out, int B, int T, int C, int64_t ld_in, int64_t ld_out, float eps, cudaStream_t stream) {
    if (!y || !stats || !out) { set_error("isp_instance_norm_apply: null pointer"); return ISP_ERR_INVALID; }
    if (B <= 0 || T <= 0 || C <= 0 || parts <= 0) { set_error("isp_instance_norm_apply: sizes must be positive"); return ISP_ERR_INVALID; }
    if (dtype < ISP_DTYPE_F32 || dtype > ISP_DTYPE_F16) { set_error("isp_instance_norm_apply: bad dtype"); return ISP_ERR_INVALID; }
    const int esz = dtype == ISP_DTYPE_F32 ? 4 : 2, V = 16 / esz;
    if (C % V || (ld_in * esz) % 16 || (ld_out * esz) % 16 || ((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(out)) & 15)) {
        set_error("isp_instance_norm_apply: rows must be whole 16 B vectors (C %% %d == 0, 16 B aligned strides)", V); return ISP_ERR_INVALID;
    }
    if (B > 65535 || size_t(C) * sizeof(float2) > 200 * 1024) { set_error("isp_instance_norm_apply: B or C too large"); return ISP_ERR_UNSUPPORTED; }
    if (dtype == ISP_DTYPE_BF16) return norm_launch<__nv_bfloat16>(y, stats, weight, bias, len, out, B, T, C, ld_in, ld_out, parts, eps, stream);
    if (dtype == ISP_DTYPE_F16) return norm_launch<__half>(y, stats, weight, bias, len, out, B, T, C, ld_in, ld_out, parts, eps, stream);
    return norm_launch<float>(y, stats, weight, bias, len, out, B, T, C, ld_in, ld_out, parts, eps, stream);
}

}  // namespace isp
