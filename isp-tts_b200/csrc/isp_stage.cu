// Host -> device staging of the encoded operands of isp_loglik_forward, ragged: only the rows below each utterance's
// length cross PCIe; the padding rows of the device tensors are zero-filled here (the operands' contract,
// tts/models/acoustic/modules/alignment.py:75-76 in the reference: projections are exactly 0 at padded positions).
//
// The reference moves whole padded batches with tensor.to(device) (tts/train.py -> trainer: batch.to(device)); with
// LJSpeech-shaped lengths 40 % of those bytes are padding.  A plain gather kernel reads the pinned host tensors through
// their device-visible (UVA) addresses: 16 B per lane, eight loads in flight per thread.  Alone, any grid from 8 CTAs up
// saturates the link (51 GB/s against 54 GB/s for the DMA of the padded tensors); next to the previous step's kernels,
// which is where it runs, a CTA only gets the issue slots and memory pipes its SM has left, and one CTA per SM is what
// keeps the link busy (end to end at cfg3: 8 CTAs 1.39 ms/step, 32: 1.05, 148: 1.00).

#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "common.cuh"
#include "isp_internal.h"

namespace isp {

constexpr int kStageUnroll = 8;
constexpr int kStageThreads = 512;

__global__ void __launch_bounds__(kStageThreads)
stage_operands_kernel(const uint4* __restrict__ hq, const uint4* __restrict__ hk, const int64_t* __restrict__ text_len,
                      const int64_t* __restrict__ mel_len, uint4* __restrict__ dq, uint4* __restrict__ dk,
                      int B, int T1max, int T2max, int vpr /* 16 B vectors per row */) {
    const long long nq = (long long)B * T1max * vpr, nk = (long long)B * T2max * vpr;
    const long long total = nq + nk;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long base = (long long)blockIdx.x * blockDim.x + threadIdx.x; base < total; base += stride * kStageUnroll) {
        uint4 v[kStageUnroll];
        bool ok[kStageUnroll];
#pragma unroll
        for (int u = 0; u < kStageUnroll; ++u) {
            const long long idx = base + u * stride;
            v[u] = make_uint4(0u, 0u, 0u, 0u);
            ok[u] = false;
            if (idx < total) {
                const bool isq = idx < nq;
                const long long e = isq ? idx : idx - nq;
                const int T = isq ? T1max : T2max;
                const long long row = e / vpr;
                const int b = int(row / T), t = int(row - (long long)b * T);
                const long long len = isq ? mel_len[b] : text_len[b];
                ok[u] = true;
                if (t < len) v[u] = isq ? hq[e] : hk[e];          // PCIe read (zero-copy), only for valid rows
            }
        }
#pragma unroll
        for (int u = 0; u < kStageUnroll; ++u) {
            const long long idx = base + u * stride;
            if (ok[u]) { if (idx < nq) dq[idx] = v[u]; else dk[idx - nq] = v[u]; }
        }
    }
}

static int g_opt_stage_ctas = 0;

int stage_set_option(const char* key, int value, int* prev) {
    if (!strcmp(key, "stage.ctas")) { *prev = g_opt_stage_ctas; g_opt_stage_ctas = value; return 0; }
    return -1;
}

int stage_operands(const void* q_host, const void* k_host, int dtype, const int64_t* text_len, const int64_t* mel_len,
                   int B, int T1max, int T2max, int D, void* q_dev, void* k_dev, cudaStream_t stream) {
    if (!q_host || !k_host || !text_len || !mel_len || !q_dev || !k_dev) { set_error("isp_stage_operands: null pointer"); return ISP_ERR_INVALID; }
    if (B <= 0 || T1max <= 0 || T2max <= 0 || D <= 0) { set_error("isp_stage_operands: sizes must be positive"); return ISP_ERR_INVALID; }
    if (dtype != ISP_DTYPE_F32 && dtype != ISP_DTYPE_BF16) { set_error("isp_stage_operands: dtype must be ISP_DTYPE_F32 or ISP_DTYPE_BF16"); return ISP_ERR_INVALID; }
    const int elem = dtype == ISP_DTYPE_BF16 ? 2 : 4;
    if ((size_t(D) * elem) % 16 != 0) { set_error("isp_stage_operands: D * elem = %zu B must be a multiple of 16", size_t(D) * elem); return ISP_ERR_UNSUPPORTED; }
    if ((reinterpret_cast<uintptr_t>(q_host) | reinterpret_cast<uintptr_t>(k_host) | reinterpret_cast<uintptr_t>(q_dev) | reinterpret_cast<uintptr_t>(k_dev)) & 15) {
        set_error("isp_stage_operands: all four tensors must be 16 B aligned"); return ISP_ERR_INVALID;
    }
    // the host tensors must be visible to the device: pinned (cudaHostAlloc / cudaHostRegister) under unified addressing
    const void* hp[2] = {q_host, k_host};
    const void* dev_ptr[2] = {nullptr, nullptr};
    for (int i = 0; i < 2; ++i) {
        cudaPointerAttributes at;
        cudaError_t e = cudaPointerGetAttributes(&at, hp[i]);
        if (e != cudaSuccess) return cuda_fail(e, "cudaPointerGetAttributes(host operand)");
        if (at.type == cudaMemoryTypeUnregistered || at.devicePointer == nullptr) {
            set_error("isp_stage_operands: the host operands must be pinned memory (torch .pin_memory() / cudaHostAlloc)");
            return ISP_ERR_INVALID;
        }
        dev_ptr[i] = at.devicePointer;
    }
    const int vpr = int(size_t(D) * elem / 16);
    stage_operands_kernel<<<g_opt_stage_ctas > 0 ? g_opt_stage_ctas : 148, kStageThreads, 0, stream>>>(static_cast<const uint4*>(dev_ptr[0]), static_cast<const uint4*>(dev_ptr[1]),
                                                           text_len, mel_len, static_cast<uint4*>(q_dev), static_cast<uint4*>(k_dev),
                                                           B, T1max, T2max, vpr);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "stage_operands_kernel launch");
    return 0;
}


// The same operands arriving PACKED: the valid rows of all utterances back to back (utterance b's rows start at the exclusive
// running sum of the lengths), already on the device -- one plain cudaMemcpyAsync on a copy engine moves them, which is what
// scales when several GPUs share the host (zero-copy reads by SMs of 4-8 GPUs through one root complex reached 182 GB/s
// aggregate in round 1; the DMA engines are the platform's ceiling).  This kernel scatters them into the padded (B, T, D)
// tensors the log-likelihood kernel loads, padding rows zero-filled: 16 B per lane, output-driven, offsets from one block scan.
__global__ void __launch_bounds__(1024)
unpack_offsets_kernel(const int64_t* __restrict__ text_len, const int64_t* __restrict__ mel_len, long long* __restrict__ off,
                      int B, int T1max, int T2max) {
    // off[0 .. B): first packed row of utterance b in Q; off[B .. 2B): the same for K.  One CTA, serial chunks of 1024.
    __shared__ long long s_scan[1024];
    __shared__ long long s_carry;
    for (int which = 0; which < 2; ++which) {
        const int64_t* len = which ? text_len : mel_len;
        const int Tmax = which ? T2max : T1max;
        if (threadIdx.x == 0) s_carry = 0;
        __syncthreads();
        for (int b0 = 0; b0 < B; b0 += 1024) {
            const int b = b0 + threadIdx.x;
            long long v = 0;
            if (b < B) { v = len[b]; v = v < 0 ? 0 : (v > Tmax ? Tmax : v); }
            s_scan[threadIdx.x] = v;
            __syncthreads();
            for (int o = 1; o < 1024; o <<= 1) {
                const long long t = threadIdx.x >= o ? s_scan[threadIdx.x - o] : 0;
                __syncthreads();
                s_scan[threadIdx.x] += t;
                __syncthreads();
            }
            if (b < B) off[which * B + b] = s_carry + s_scan[threadIdx.x] - v;
            __syncthreads();
            if (threadIdx.x == 1023) s_carry += s_scan[1023];
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(512)
unpack_operands_kernel(const uint4* __restrict__ pq, const uint4* __restrict__ pk, const int64_t* __restrict__ text_len,
                       const int64_t* __restrict__ mel_len, const long long* __restrict__ off, uint4* __restrict__ dq,
                       uint4* __restrict__ dk, int B, int T1max, int T2max, int vpr) {
    const long long nq = (long long)B * T1max * vpr, total = nq + (long long)B * T2max * vpr;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
        const bool isq = idx < nq;
        const long long e = isq ? idx : idx - nq;
        const int T = isq ? T1max : T2max;
        const long long row = e / vpr;
        const int v = int(e - row * vpr);
        const int b = int(row / T), t = int(row - (long long)b * T);
        const long long len = isq ? mel_len[b] : text_len[b];
        uint4 val = make_uint4(0u, 0u, 0u, 0u);
        if (t < len) val = __ldcs((isq ? pq : pk) + (off[(isq ? 0 : B) + b] + t) * vpr + v);
        (isq ? dq : dk)[e] = val;
    }
}

size_t unpack_workspace_bytes(int B) { return size_t(B) * 2 * sizeof(long long); }

int unpack_operands(const void* q_packed, const void* k_packed, int dtype, const int64_t* text_len, const int64_t* mel_len,
                    int B, int T1max, int T2max, int D, void* q_dev, void* k_dev, void* ws, size_t ws_bytes, cudaStream_t stream) {
    if (!q_packed || !k_packed || !text_len || !mel_len || !q_dev || !k_dev || !ws) { set_error("isp_unpack_operands: null pointer"); return ISP_ERR_INVALID; }
    if (B <= 0 || T1max <= 0 || T2max <= 0 || D <= 0) { set_error("isp_unpack_operands: sizes must be positive"); return ISP_ERR_INVALID; }
    if (dtype != ISP_DTYPE_F32 && dtype != ISP_DTYPE_BF16) { set_error("isp_unpack_operands: dtype must be ISP_DTYPE_F32 or ISP_DTYPE_BF16"); return ISP_ERR_INVALID; }
    const int elem = dtype == ISP_DTYPE_BF16 ? 2 : 4;
    if ((size_t(D) * elem) % 16 != 0) { set_error("isp_unpack_operands: D * elem = %zu B must be a multiple of 16", size_t(D) * elem); return ISP_ERR_UNSUPPORTED; }
    if ((reinterpret_cast<uintptr_t>(q_packed) | reinterpret_cast<uintptr_t>(k_packed) | reinterpret_cast<uintptr_t>(q_dev) | reinterpret_cast<uintptr_t>(k_dev)) & 15) {
        set_error("isp_unpack_operands: all four tensors must be 16 B aligned"); return ISP_ERR_INVALID;
    }
    if (ws_bytes < unpack_workspace_bytes(B) || (reinterpret_cast<uintptr_t>(ws) & 7)) { set_error("isp_unpack_operands: workspace too small or misaligned"); return ISP_ERR_WORKSPACE; }
    long long* off = static_cast<long long*>(ws);
    unpack_offsets_kernel<<<1, 1024, 0, stream>>>(text_len, mel_len, off, B, T1max, T2max);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "unpack_offsets_kernel launch");
    const int vpr = int(size_t(D) * elem / 16);
    unpack_operands_kernel<<<148 * 4, 512, 0, stream>>>(static_cast<const uint4*>(q_packed), static_cast<const uint4*>(k_packed), text_len, mel_len,
                                                        off, static_cast<uint4*>(q_dev), static_cast<uint4*>(k_dev), B, T1max, T2max, vpr);
    e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "unpack_operands_kernel launch");
    return 0;
}

}  // namespace isp
