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

}  // namespace isp
