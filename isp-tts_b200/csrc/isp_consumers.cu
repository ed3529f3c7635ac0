// Consumers of the alignment that take the path straight from the MAS kernel.
//
// Attention binarization loss, tts/models/acoustic/loss.py:97-105 in the reference:
//     log_sum = log(clamp(soft_attention[hard_attention == 1], min=eps)).sum();  loss = -log_sum / hard_attention.sum()
// The boolean-mask gather reads the dense int16 path (2 B/cell) and the dense fp32 attention (4 B/cell) to pick one
// cell per frame.  With the path as a column index per frame (isp_mas_forward_path) the same sums are one gather of
// sum(mel_len) floats: sums[0] = sum log(max(soft[b, i, path[b, i]], eps)), sums[1] = number of frames.

#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>

#include "common.cuh"
#include "isp_internal.h"

namespace isp {

__global__ void __launch_bounds__(256)
bin_loss_kernel(const float* __restrict__ soft, const int16_t* __restrict__ path, const int64_t* __restrict__ mel_len,
                int B, int T1max, int T2max, float eps, float* __restrict__ sums) {
    const long long total = (long long)B * T1max;
    float acc = 0.0f, cnt = 0.0f;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int b = int(idx / T1max), i = int(idx - (long long)b * T1max);
        const long long n = mel_len[b];
        const int col = path[idx];
        if (i < n && col >= 0 && col < T2max) {
            acc += logf(fmaxf(__ldg(soft + size_t(idx) * T2max + col), eps));
            cnt += 1.0f;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc += __shfl_xor_sync(0xffffffffu, acc, o);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    __shared__ float sa[8], sc[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { sa[warp] = acc; sc[warp] = cnt; }
    __syncthreads();
    if (warp == 0) {
        acc = lane < 8 ? sa[lane] : 0.0f;
        cnt = lane < 8 ? sc[lane] : 0.0f;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            acc += __shfl_xor_sync(0xffffffffu, acc, o);
            cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        }
        if (lane == 0) { atomicAdd(sums, acc); atomicAdd(sums + 1, cnt); }
    }
}

int bin_loss_sums(const float* attn_soft, const int16_t* path, const int64_t* mel_len, int B, int T1max, int T2max,
                  float eps, float* sums, cudaStream_t stream) {
    if (!attn_soft || !path || !mel_len || !sums) { set_error("isp_bin_loss_sums: null pointer"); return ISP_ERR_INVALID; }
    if (B <= 0 || T1max <= 0 || T2max <= 0) { set_error("isp_bin_loss_sums: sizes must be positive"); return ISP_ERR_INVALID; }
    cudaError_t e = cudaMemsetAsync(sums, 0, 2 * sizeof(float), stream);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(sums)");
    const long long total = (long long)B * T1max;
    const int grid = int(std::min<long long>((total + 255) / 256, 148 * 8));
    bin_loss_kernel<<<grid, 256, 0, stream>>>(attn_soft, path, mel_len, B, T1max, T2max, eps, sums);
    e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "bin_loss_kernel launch");
    return 0;
}

// Length regulator, tts/models/acoustic/modules/temporal_adaptor.py:411-436 (`durations` branch): the reference builds a
// (T1 x T2) 0/1 matrix from the cumulated durations and multiplies it with x (2*T1*T2*C flops, 4 B/cell of matrix).  Each
// row of that matrix has a single one, at the path's column, so the product is a gather: one 16 B vector per thread,
// x rows come from L2 (consecutive frames repeat the same token), the output is written once with streaming stores.
__global__ void __launch_bounds__(256)
length_regulate_kernel(const uint4* __restrict__ x, const int16_t* __restrict__ path, uint4* __restrict__ out,
                       long long rows, int T1max, int T2max, int vec_per_row) {
    // one warp per frame: the path entry is read once per row, no per-element division
    const int lane = threadIdx.x & 31;
    const long long wstride = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows; row += wstride) {
        const int col = __ldg(path + row);
        uint4* dst = out + row * vec_per_row;
        if (col >= 0 && col < T2max) {
            const uint4* src = x + ((row / T1max) * T2max + col) * vec_per_row;
            for (int v = lane; v < vec_per_row; v += 32) st_cs_v4(dst + v, __ldg(src + v));
        } else {
            for (int v = lane; v < vec_per_row; v += 32) st_cs_v4(dst + v, make_uint4(0u, 0u, 0u, 0u));
        }
    }
}

// Backward: gx[b, j, :] = sum of g[b, t, :] over the token's frames [starts, starts + durations); one CTA per token,
// each thread a float4 of channels, frames read in order (fixed summation order, coalesced rows).
__global__ void __launch_bounds__(128)
length_regulate_bwd_kernel(const float4* __restrict__ g, const int64_t* __restrict__ durations, const int64_t* __restrict__ starts,
                           float4* __restrict__ gx, int T1max, int T2max, int vec_per_row) {
    const long long tok = blockIdx.x;                 // b * T2max + j
    const long long b = tok / T2max;
    long long t0 = starts[tok], n = durations[tok];
    if (t0 < 0) { n += t0; t0 = 0; }
    if (t0 + n > T1max) n = T1max - t0;
    for (int v = threadIdx.x; v < vec_per_row; v += blockDim.x) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4* src = g + (b * T1max + t0) * vec_per_row + v;
        for (long long t = 0; t < n; ++t) {
            const float4 a = __ldcs(src + t * vec_per_row);
            acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
        }
        gx[tok * vec_per_row + v] = acc;
    }
}

int length_regulate(const void* x, const int16_t* path, void* out, int dtype, int B, int T1max, int T2max, int C, cudaStream_t stream) {
    if (!x || !path || !out) { set_error("isp_length_regulate: null pointer"); return ISP_ERR_INVALID; }
    if (B <= 0 || T1max <= 0 || T2max <= 0 || C <= 0) { set_error("isp_length_regulate: sizes must be positive"); return ISP_ERR_INVALID; }
    if (dtype != ISP_DTYPE_F32 && dtype != ISP_DTYPE_BF16) { set_error("isp_length_regulate: dtype must be ISP_DTYPE_F32 or ISP_DTYPE_BF16"); return ISP_ERR_INVALID; }
    const int elem = dtype == ISP_DTYPE_BF16 ? 2 : 4;
    if ((size_t(C) * elem) % 16 != 0 || (reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) % 16 != 0) {
        set_error("isp_length_regulate: rows of x / out must be whole 16 B vectors (C * elem %% 16 == 0, 16 B aligned)");
        return ISP_ERR_INVALID;
    }
    const int vec = int(size_t(C) * elem / 16);
    const long long rows = (long long)B * T1max;
    const int grid = int(std::min<long long>((rows + 7) / 8, 148LL * 8 * 4));
    length_regulate_kernel<<<grid, 256, 0, stream>>>(static_cast<const uint4*>(x), path, static_cast<uint4*>(out), rows, T1max, T2max, vec);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "length_regulate_kernel launch");
    return 0;
}

int length_regulate_backward(const float* g, const int64_t* durations, const int64_t* starts, float* gx,
                             int B, int T1max, int T2max, int C, cudaStream_t stream) {
    if (!g || !durations || !starts || !gx) { set_error("isp_length_regulate_backward: null pointer"); return ISP_ERR_INVALID; }
    if (B <= 0 || T1max <= 0 || T2max <= 0 || C <= 0) { set_error("isp_length_regulate_backward: sizes must be positive"); return ISP_ERR_INVALID; }
    if (C % 4 != 0 || (reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(gx)) % 16 != 0) {
        set_error("isp_length_regulate_backward: C must be a multiple of 4 and g / gx 16 B aligned");
        return ISP_ERR_INVALID;
    }
    const long long toks = (long long)B * T2max;
    if (toks > 0x7fffffffLL) { set_error("isp_length_regulate_backward: B * T2max too large"); return ISP_ERR_INVALID; }
    length_regulate_bwd_kernel<<<unsigned(toks), 128, 0, stream>>>(reinterpret_cast<const float4*>(g), durations, starts,
                                                                   reinterpret_cast<float4*>(gx), T1max, T2max, C / 4);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "length_regulate_bwd_kernel launch");
    return 0;
}


// Per-token average of a frame-level feature (pitch, energy), tts/models/acoustic/modules/temporal_adaptor.py:439-465
// (`durations` branch): out[b, c, j] = sum of x[b, c, t] over the token's frames / number of NON-ZERO x among them, 0 if
// there is none.  The reference takes differences of fp32 running sums over the whole utterance (a dozen kernels and
// cancellation between two large partial sums); here the token boundaries come from one block scan of the durations and
// each token's frames are summed directly.
__global__ void __launch_bounds__(256)
temporal_average_kernel(const float* __restrict__ x, const int64_t* __restrict__ durations, float* __restrict__ out,
                        int C, int T1max, int T2max) {
    extern __shared__ int s_end[];                    // [T2max] inclusive running sum of the durations
    __shared__ int s_part[256];
    const int b = blockIdx.x, tid = threadIdx.x;
    const int64_t* d = durations + size_t(b) * T2max;
    const int per = (T2max + 255) / 256;              // consecutive tokens per thread
    const int j0 = tid * per, j1 = min(T2max, j0 + per);
    int local = 0;
    for (int j = j0; j < j1; ++j) { const long long v = d[j]; local += int(v < 0 ? 0 : (v > T1max ? T1max : v)); }
    s_part[tid] = local;
    __syncthreads();
    for (int o = 1; o < 256; o <<= 1) {               // inclusive scan of the per-thread sums
        const int v = tid >= o ? s_part[tid - o] : 0;
        __syncthreads();
        s_part[tid] += v;
        __syncthreads();
    }
    int run = s_part[tid] - local;
    for (int j = j0; j < j1; ++j) { const long long v = d[j]; run += int(v < 0 ? 0 : (v > T1max ? T1max : v)); s_end[j] = run; }
    __syncthreads();
    for (int idx = tid; idx < C * T2max; idx += 256) {
        const int c = idx / T2max, j = idx - c * T2max;
        const int t1 = min(s_end[j], T1max), t0 = min(j > 0 ? s_end[j - 1] : 0, t1);
        const float* xr = x + (size_t(b) * C + c) * T1max;
        float sum = 0.0f;
        int cnt = 0;
        for (int t = t0; t < t1; ++t) { const float v = __ldg(xr + t); sum += v; cnt += v != 0.0f; }
        out[(size_t(b) * C + c) * T2max + j] = cnt ? sum / float(cnt) : 0.0f;
    }
}

int temporal_average(const float* x, const int64_t* durations, float* out, int B, int C, int T1max, int T2max, cudaStream_t stream) {
    if (!x || !durations || !out) { set_error("isp_temporal_average: null pointer"); return ISP_ERR_INVALID; }
    if (B <= 0 || C <= 0 || T1max <= 0 || T2max <= 0) { set_error("isp_temporal_average: sizes must be positive"); return ISP_ERR_INVALID; }
    if (T2max > 12288) { set_error("isp_temporal_average: T2max=%d > 12288 tokens", T2max); return ISP_ERR_UNSUPPORTED; }
    temporal_average_kernel<<<B, 256, size_t(T2max) * sizeof(int), stream>>>(x, durations, out, C, T1max, T2max);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "temporal_average_kernel launch");
    return 0;
}

}  // namespace isp
