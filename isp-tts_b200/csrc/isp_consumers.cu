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


// The path as one token index per frame from (rounded) durations, for callers that hold durations but no MAS path -- the
// reference's inference route (temporal_adaptor.py:424-431 builds a T1 x T2 0/1 matrix from the cumulated durations for
// this).  One CTA per utterance: block scan of the durations, then every token writes its own run of frames; frames past
// the utterance's total get -1.
__global__ void __launch_bounds__(256)
path_from_durations_kernel(const int64_t* __restrict__ reps, int16_t* __restrict__ path, int T1max, int T2max) {
    __shared__ int s_part[256];
    __shared__ int s_total;
    const int b = blockIdx.x, tid = threadIdx.x;
    const int64_t* d = reps + size_t(b) * T2max;
    int16_t* out = path + size_t(b) * T1max;
    const int per = (T2max + 255) / 256;
    const int j0 = min(T2max, tid * per), j1 = min(T2max, j0 + per);
    int local = 0;
    for (int j = j0; j < j1; ++j) { const long long v = d[j]; local += int(v < 0 ? 0 : (v > T1max ? T1max : v)); }
    s_part[tid] = local;
    __syncthreads();
    for (int o = 1; o < 256; o <<= 1) {
        const int v = tid >= o ? s_part[tid - o] : 0;
        __syncthreads();
        s_part[tid] += v;
        __syncthreads();
    }
    if (tid == 255) s_total = min(s_part[255], T1max);
    int run = s_part[tid] - local;
    for (int j = j0; j < j1; ++j) {
        const long long v = d[j];
        const int n = int(v < 0 ? 0 : (v > T1max ? T1max : v));
        for (int t = run; t < min(run + n, T1max); ++t) out[t] = int16_t(j);
        run += n;
    }
    __syncthreads();
    for (int t = s_total + tid; t < T1max; t += 256) out[t] = -1;
}

int path_from_durations(const int64_t* reps, int16_t* path, int B, int T1max, int T2max, cudaStream_t stream) {
    if (!reps || !path) { set_error("isp_path_from_durations: null pointer"); return ISP_ERR_INVALID; }
    if (B <= 0 || T1max <= 0 || T2max <= 0 || T2max > 32767) { set_error("isp_path_from_durations: sizes must be positive, T2max <= 32767"); return ISP_ERR_INVALID; }
    path_from_durations_kernel<<<B, 256, 0, stream>>>(reps, path, T1max, T2max);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "path_from_durations_kernel launch");
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
    if (T2max > 12000) { set_error("isp_temporal_average: T2max=%d > 12000 tokens", T2max); return ISP_ERR_UNSUPPORTED; }   // 47 KB dynamic + 1 KB static <= 48 KB
    temporal_average_kernel<<<B, 256, size_t(T2max) * sizeof(int), stream>>>(x, durations, out, C, T1max, T2max);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "temporal_average_kernel launch");
    return 0;
}

}  // namespace isp

namespace isp {

// Per-token average of frame-level features on the soft route, tts/models/acoustic/modules/temporal_adaptor.py:446-449:
//     out[b, c, j] = sum_t x[b, c, t] * a[b, t, j] / (sum_t a[b, t, j] + 1e-5),      a = attn_soft (B, T1max, T2max)
// A contraction with one to four rows (pitch, energy) is a stream over `a`, not a GEMM: each element of `a` is read once
// (16 B per lane, rows of one chunk in flight together) and feeds C + 1 accumulators per column; fp32 throughout, so the
// averages carry no tensor-core rounding.  Pass 1 writes per-chunk partial sums (fixed order -> deterministic), pass 2 adds
// the chunks and divides.  Rows >= row_len[b] (optional) are not read.
constexpr int kSoftAvgRows = 64;     // rows of `a` per CTA
constexpr int kSoftAvgMaxC = 4;

template <int C>
__global__ void __launch_bounds__(64)
soft_average_partial_kernel(const float* __restrict__ x, const float* __restrict__ a, const int64_t* __restrict__ row_len,
                            float* __restrict__ part, int T1max, int T2max, int chunks) {
    const int chunk = blockIdx.x, b = blockIdx.y;
    long long n = row_len ? row_len[b] : T1max;
    const int rows_all = int(n < 0 ? 0 : (n > T1max ? T1max : n));
    const int r0 = chunk * kSoftAvgRows, r1 = min(r0 + kSoftAvgRows, rows_all);
    const int ncg = T2max >> 2;                                                        // T2max % 4 == 0 (host checks)
    const float* ab = a + size_t(b) * T1max * T2max;
    const float* xb = x + size_t(b) * C * T1max;
    float* pb = part + (size_t(b) * chunks + chunk) * (C + 1) * T2max;
    for (int cg = threadIdx.x; cg < ncg; cg += 64) {
        float4 acc[C + 1];
#pragma unroll
        for (int c = 0; c <= C; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
        for (int r = r0; r < r1; ++r) {
            const float4 v = __ldcs(reinterpret_cast<const float4*>(ab + size_t(r) * T2max) + cg);
            acc[C].x += v.x; acc[C].y += v.y; acc[C].z += v.z; acc[C].w += v.w;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const float f = __ldg(xb + size_t(c) * T1max + r);
                acc[c].x = fmaf(f, v.x, acc[c].x); acc[c].y = fmaf(f, v.y, acc[c].y);
                acc[c].z = fmaf(f, v.z, acc[c].z); acc[c].w = fmaf(f, v.w, acc[c].w);
            }
        }
#pragma unroll
        for (int c = 0; c <= C; ++c) reinterpret_cast<float4*>(pb + size_t(c) * T2max)[cg] = acc[c];
    }
}

__global__ void __launch_bounds__(256)
soft_average_finish_kernel(const float* __restrict__ part, float* __restrict__ out, float* __restrict__ colsum,
                           int C, int T2max, int chunks, long long total) {
    // total = B * C * T2max outputs
    for (long long idx = (long long)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (long long)gridDim.x * 256) {
        const int j = int(idx % T2max);
        const long long bc = idx / T2max;
        const int c = int(bc % C);
        const long long b = bc / C;
        const float* pb = part + size_t(b) * chunks * (C + 1) * T2max;
        float num = 0.f, den = 0.f;
        for (int k = 0; k < chunks; ++k) {
            num += pb[(size_t(k) * (C + 1) + c) * T2max + j];
            den += pb[(size_t(k) * (C + 1) + C) * T2max + j];
        }
        out[idx] = num / (den + 1e-5f);
        if (c == 0 && colsum) colsum[b * T2max + j] = den;
    }
}

// d a[b, t, j] = sum_c g[b, c, j] / (S[b, j] + 1e-5) * (x[b, c, t] - out[b, c, j]): one pass, 4 B/cell written.
__global__ void __launch_bounds__(256)
soft_average_bwd_kernel(const float* __restrict__ g, const float* __restrict__ x, const float* __restrict__ out,
                        const float* __restrict__ colsum, float* __restrict__ ga, int C, int T1max, int T2max, long long total4) {
    const int ncg = T2max >> 2;
    for (long long idx = (long long)blockIdx.x * 256 + threadIdx.x; idx < total4; idx += (long long)gridDim.x * 256) {
        const int cg = int(idx % ncg);
        const long long bt = idx / ncg;
        const int t = int(bt % T1max);
        const long long b = bt / T1max;
        const float4 s = *reinterpret_cast<const float4*>(colsum + b * T2max + 4 * cg);
        const float4 inv = make_float4(1.f / (s.x + 1e-5f), 1.f / (s.y + 1e-5f), 1.f / (s.z + 1e-5f), 1.f / (s.w + 1e-5f));
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int c = 0; c < C; ++c) {
            const float f = __ldg(x + (b * C + c) * T1max + t);
            const float4 gg = *reinterpret_cast<const float4*>(g + (b * C + c) * T2max + 4 * cg);
            const float4 oo = *reinterpret_cast<const float4*>(out + (b * C + c) * T2max + 4 * cg);
            acc.x += gg.x * inv.x * (f - oo.x); acc.y += gg.y * inv.y * (f - oo.y);
            acc.z += gg.z * inv.z * (f - oo.z); acc.w += gg.w * inv.w * (f - oo.w);
        }
        __stcs(reinterpret_cast<float4*>(ga) + idx, acc);
    }
}

size_t soft_average_workspace_bytes(int B, int C, int T1max, int T2max) {
    const size_t chunks = (size_t(T1max) + kSoftAvgRows - 1) / kSoftAvgRows;
    return size_t(B) * chunks * (size_t(C) + 1) * size_t(T2max) * sizeof(float);
}

int soft_average(const float* x, const float* attn_soft, const int64_t* row_len, float* out, float* colsum, int B, int C,
                 int T1max, int T2max, void* ws, size_t ws_bytes, cudaStream_t stream) {
    if (!x || !attn_soft || !out || !ws) { set_error("isp_soft_average: null pointer"); return ISP_ERR_INVALID; }
    if (B <= 0 || C <= 0 || T1max <= 0 || T2max <= 0) { set_error("isp_soft_average: sizes must be positive"); return ISP_ERR_INVALID; }
    if (C > kSoftAvgMaxC) { set_error("isp_soft_average: C=%d > %d feature rows", C, kSoftAvgMaxC); return ISP_ERR_UNSUPPORTED; }
    if (T2max % 4 != 0 || (reinterpret_cast<uintptr_t>(attn_soft) & 15) || (reinterpret_cast<uintptr_t>(ws) & 15)) {
        set_error("isp_soft_average: T2max must be a multiple of 4 and attn_soft / ws 16 B aligned"); return ISP_ERR_INVALID;
    }
    if (B > 65535) { set_error("isp_soft_average: B=%d > 65535", B); return ISP_ERR_UNSUPPORTED; }
    if (ws_bytes < soft_average_workspace_bytes(B, C, T1max, T2max)) { set_error("isp_soft_average: workspace too small"); return ISP_ERR_WORKSPACE; }
    const int chunks = (T1max + kSoftAvgRows - 1) / kSoftAvgRows;
    float* part = static_cast<float*>(ws);
    const dim3 grid(chunks, B);
    switch (C) {
        case 1: soft_average_partial_kernel<1><<<grid, 64, 0, stream>>>(x, attn_soft, row_len, part, T1max, T2max, chunks); break;
        case 2: soft_average_partial_kernel<2><<<grid, 64, 0, stream>>>(x, attn_soft, row_len, part, T1max, T2max, chunks); break;
        case 3: soft_average_partial_kernel<3><<<grid, 64, 0, stream>>>(x, attn_soft, row_len, part, T1max, T2max, chunks); break;
        default: soft_average_partial_kernel<4><<<grid, 64, 0, stream>>>(x, attn_soft, row_len, part, T1max, T2max, chunks); break;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "soft_average_partial_kernel launch");
    const long long total = (long long)B * C * T2max;
    soft_average_finish_kernel<<<int(std::min<long long>((total + 255) / 256, 148 * 8)), 256, 0, stream>>>(part, out, colsum, C, T2max, chunks, total);
    e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "soft_average_finish_kernel launch");
    return 0;
}

int soft_average_backward(const float* g, const float* x, const float* out, const float* colsum, float* g_soft, int B, int C,
                          int T1max, int T2max, cudaStream_t stream) {
    if (!g || !x || !out || !colsum || !g_soft) { set_error("isp_soft_average_backward: null pointer"); return ISP_ERR_INVALID; }
    if (B <= 0 || C <= 0 || T1max <= 0 || T2max <= 0 || T2max % 4 != 0) { set_error("isp_soft_average_backward: sizes must be positive, T2max %% 4 == 0"); return ISP_ERR_INVALID; }
    if ((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(colsum) | reinterpret_cast<uintptr_t>(g_soft)) & 15) {
        set_error("isp_soft_average_backward: tensors must be 16 B aligned"); return ISP_ERR_INVALID;
    }
    const long long total4 = (long long)B * T1max * (T2max >> 2);
    soft_average_bwd_kernel<<<int(std::min<long long>((total4 + 255) / 256, 148 * 16)), 256, 0, stream>>>(g, x, out, colsum, g_soft, C, T1max, T2max, total4);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "soft_average_bwd_kernel launch");
    return 0;
}

}  // namespace isp
