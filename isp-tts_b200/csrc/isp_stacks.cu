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

// channels-first input: a tile of 32 frames x up to 128 channels through shared memory.  Reads run along t (128 B per channel
// row), writes along c (the tile's 32 output rows are 2 * CW .. 4 * CW bytes each, whole rows when Cp <= 128); a tile past the
// utterance's length is zeros and reads nothing.
constexpr int kPrepCW = 128;
template <typename TI, typename TO>
__global__ void __launch_bounds__(256)
prep_cf_kernel(const TI* __restrict__ x, const int64_t* __restrict__ len, TO* __restrict__ out, int C, int T, int Cp) {
    __shared__ __align__(16) float tile[32][kPrepCW + 4];          // [t][c]; pitch 132 floats: 16 B aligned rows, conflict-free column writes
    const int b = blockIdx.z, t0 = blockIdx.x * 32, c0 = blockIdx.y * kPrepCW;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 32 x 8
    long long n = len ? len[b] : T;
    const int valid = int(n < 0 ? 0 : (n > T ? T : n));
    const int cw = min(kPrepCW, Cp - c0);                            // a multiple of V (host checks Cp % V == 0)
    constexpr int V = 16 / sizeof(TO);                               // channels per 16 B of output
    TO* ob = out + size_t(b) * T * Cp;
    const bool live = t0 < valid;
    if (live) {
        const TI* xb = x + size_t(b) * C * T;
#pragma unroll 4
        for (int r = ty; r < cw; r += 8) {
            const int c = c0 + r, t = t0 + tx;
            tile[tx][r] = (c < C && t < valid) ? to_f<TI>(xb[size_t(c) * T + t]) : 0.0f;
        }
        __syncthreads();
    }
    const int vpr = cw / V;
    for (int idx = threadIdx.x; idx < 32 * vpr; idx += 256) {
        const int r = idx / vpr, v = idx - r * vpr;
        if (t0 + r >= T) continue;
        TO vals[V];
#pragma unroll
        for (int k = 0; k < V; ++k) vals[k] = from_f<TO>(live ? tile[r][v * V + k] : 0.0f);
        *reinterpret_cast<uint4*>(ob + size_t(t0 + r) * Cp + c0 + v * V) = *reinterpret_cast<uint4*>(vals);
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
        const dim3 grid((T + 31) / 32, (Cp + kPrepCW - 1) / kPrepCW, B);
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

// Pass 1: per (utterance, channel) add the slabs' partial sums in slab order (deterministic), mean and rstd in fp32, and leave
// scale a = rstd * w and shift d = bias - mean * a in `ab` (B, C) float2.
__global__ void __launch_bounds__(128)
instance_norm_stats_kernel(const float* __restrict__ stats, const float* __restrict__ weight, const float* __restrict__ bias,
                           const int64_t* __restrict__ len, float2* __restrict__ ab, int Tmax, int C, int parts, float eps) {
    const int b = blockIdx.y, c = blockIdx.x * 128 + threadIdx.x;
    if (c >= C) return;
    long long n64 = len ? len[b] : Tmax;
    const int n = int(n64 < 0 ? 0 : (n64 > Tmax ? Tmax : n64));
    const float inv_n = n > 0 ? 1.0f / float(n) : 0.0f;
    const int used = min(parts, ((n + 127) / 128) * 4);          // slabs of tiles past the length were never written (zeros)
    const float* st = stats + size_t(b) * parts * C * 2;
    float s = 0.f, q = 0.f;
    for (int p = 0; p < used; ++p) {
        const float2 v = *reinterpret_cast<const float2*>(st + (size_t(p) * C + c) * 2);
        s += v.x; q += v.y;
    }
    const float mean = s * inv_n;
    const float var = fmaxf(q * inv_n - mean * mean, 0.0f);
    const float a = rsqrtf(var + eps) * (weight ? weight[c] : 1.0f);
    ab[size_t(b) * C + c] = make_float2(a, (bias ? bias[c] : 0.0f) - mean * a);
}

// Pass 2: one CTA = 64 frames x all channels of one utterance.  A thread owns one 16 B vector of channels (8 of a 2-byte type, 4
// of fp32) -- its scales and shifts sit in registers -- and walks the rows, four loads in flight.  Rows past the length are
// zeros; in place they already are (the GEMM's epilogue masked them), so such rows are not touched at all.
constexpr int kNormRows = 64;
template <typename T>
__global__ void __launch_bounds__(256)
instance_norm_kernel(const T* __restrict__ y, const float2* __restrict__ ab_all, const int64_t* __restrict__ len, T* __restrict__ out,
                     int Tmax, int C, int ld_in, int ld_out) {
    const int b = blockIdx.y, t0 = blockIdx.x * kNormRows;
    long long n64 = len ? len[b] : Tmax;
    const int n = int(n64 < 0 ? 0 : (n64 > Tmax ? Tmax : n64));
    const bool in_place = (const void*)y == (const void*)out;
    if (t0 >= n && in_place) return;
    constexpr int V = 16 / sizeof(T);               // channels per 16 B
    const int vpr = C / V;                           // host checks C % V == 0
    const int rows_par = max(1, 256 / vpr);          // rows handled side by side
    const int t1 = min(t0 + kNormRows, Tmax);
    const int rp = threadIdx.x / vpr, v = threadIdx.x - rp * vpr;
    if (rp >= rows_par) return;
    float2 ab[V];
#pragma unroll
    for (int k = 0; k < V; ++k) ab[k] = ab_all[size_t(b) * C + v * V + k];
    const T* src = y + size_t(b) * Tmax * ld_in + v * V;
    T* dst = out + size_t(b) * Tmax * ld_out + v * V;
    const int tv = min(t1, n);
#pragma unroll 4
    for (int t = t0 + rp; t < tv; t += rows_par) {
        T vals[V];
        *reinterpret_cast<uint4*>(vals) = __ldcs(reinterpret_cast<const uint4*>(src + size_t(t) * ld_in));
#pragma unroll
        for (int k = 0; k < V; ++k) vals[k] = from_f<T>(fmaf(to_f<T>(vals[k]), ab[k].x, ab[k].y));
        *reinterpret_cast<uint4*>(dst + size_t(t) * ld_out) = *reinterpret_cast<uint4*>(vals);
    }
    if (!in_place) {
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
        for (int t = t0 + rp; t < t1; t += rows_par)
            if (t >= n) *reinterpret_cast<uint4*>(dst + size_t(t) * ld_out) = z;
    }
}

}  // namespace

int prep_channels_last(const void* x, int in_dtype, int channels_first, const int64_t* len, void* out, int out_dtype,
                       int B, int C, int T, int Cp, cudaStream_t stream) {
    if (!x || !out) { set_error("isp_prep_channels_last: null pointer"); return ISP_ERR_INVALID; }
    if (B <= 0 || C <= 0 || T <= 0 || Cp < C) { set_error("isp_prep_channels_last: sizes must be positive and Cp >= C"); return ISP_ERR_INVALID; }
    if (B > 65535) { set_error("isp_prep_channels_last: B=%d > 65535", B); return ISP_ERR_UNSUPPORTED; }
    {
        const int v = out_dtype == ISP_DTYPE_F32 ? 4 : 8;
        if (Cp % v || (reinterpret_cast<uintptr_t>(out) & 15)) { set_error("isp_prep_channels_last: Cp must be a multiple of %d and out 16 B aligned", v); return ISP_ERR_INVALID; }
    }
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
static int norm_launch(const void* y, const float2* ab, const int64_t* len, void* out, int B, int T_, int C, int64_t ld_in, int64_t ld_out,
                       cudaStream_t stream) {
    const dim3 grid((T_ + kNormRows - 1) / kNormRows, B);
    instance_norm_kernel<T><<<grid, 256, 0, stream>>>(static_cast<const T*>(y), ab, len, static_cast<T*>(out), T_, C, int(ld_in), int(ld_out));
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "instance_norm_kernel launch");
    return 0;
}

int instance_norm_apply(const void* y, int dtype, const float* stats, int parts, const float* weight, const float* bias,
                        const int64_t* len, void* out, int B, int T, int C, int64_t ld_in, int64_t ld_out, float eps, void* ws,
                        cudaStream_t stream) {
    if (!y || !stats || !out || !ws) { set_error("isp_instance_norm_apply: null pointer"); return ISP_ERR_INVALID; }
    if (B <= 0 || T <= 0 || C <= 0 || parts <= 0) { set_error("isp_instance_norm_apply: sizes must be positive"); return ISP_ERR_INVALID; }
    if (dtype < ISP_DTYPE_F32 || dtype > ISP_DTYPE_F16) { set_error("isp_instance_norm_apply: bad dtype"); return ISP_ERR_INVALID; }
    const int esz = dtype == ISP_DTYPE_F32 ? 4 : 2, V = 16 / esz;
    if (C % V || (ld_in * esz) % 16 || (ld_out * esz) % 16 || ((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(out)) & 15)) {
        set_error("isp_instance_norm_apply: rows must be whole 16 B vectors (C %% %d == 0, 16 B aligned strides)", V); return ISP_ERR_INVALID;
    }
    if (B > 65535 || C / V > 256) { set_error("isp_instance_norm_apply: B > 65535 or more than 256 16 B vectors of channels per row"); return ISP_ERR_UNSUPPORTED; }
    if (reinterpret_cast<uintptr_t>(ws) & 7) { set_error("isp_instance_norm_apply: ws must be 8 B aligned"); return ISP_ERR_INVALID; }
    float2* ab = static_cast<float2*>(ws);
    instance_norm_stats_kernel<<<dim3((C + 127) / 128, B), 128, 0, stream>>>(stats, weight, bias, len, ab, T, C, parts, eps);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "instance_norm_stats_kernel launch");
    if (dtype == ISP_DTYPE_BF16) return norm_launch<__nv_bfloat16>(y, ab, len, out, B, T, C, ld_in, ld_out, stream);
    if (dtype == ISP_DTYPE_F16) return norm_launch<__half>(y, ab, len, out, B, T, C, ld_in, ld_out, stream);
    return norm_launch<float>(y, ab, len, out, B, T, C, ld_in, ld_out, stream);
}

}  // namespace isp
