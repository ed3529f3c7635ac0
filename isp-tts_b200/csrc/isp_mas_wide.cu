// Monotonic Alignment Search for utterances wider than the strip kernels cover (T2max > 640 text tokens): the slow, general
// path behind isp_mas_forward, so that a long-form batch never fails at run time inside a training step.
//
// Reference semantics: tts/modules/aligner/mas.py:8-26 (mas_width1) and :30-35 (b_mas); the reference's own GPU kernel
// (tts/modules/aligner/cuda_mas.py:11-46) is limited to 256 * k columns per launch configuration and keeps three (B, T1, T2)
// scratch tensors.  Here: one CTA of 1024 threads per utterance, thread t owns the columns t, t + 1024, ...; the previous row
// of Q lives in shared memory (double-buffered, one barrier per row); the backpointer decision Q[i-1][j-1] >= Q[i-1][j] leaves
// the CTA as one bit per cell (a warp ballot is 32 consecutive columns: row-major words in the workspace, 1/8 B per cell);
// the backtrack loads a 32-row x 64-column window of bits with one lane per row and walks it in shared memory, so only
// T1 / 32 dependent global round trips remain.  Bit-exact with the reference: one fp32 add per cell on top of an exact max,
// ties (and -inf >= -inf) take the diagonal.  Roofline: HBM in principle (6 B/cell); in practice one barrier per frame row
// bounds it (~0.1 us per row), which is what a fallback is allowed to cost.

#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "common.cuh"
#include "isp_internal.h"

namespace isp {

namespace {

constexpr int kWT = 1024;                 // threads per CTA
constexpr int kWMaxCols = 16;             // columns per thread -> T2max <= 16384

struct WideParams {
    const float* logp;
    int64_t sB, sT1;
    const int64_t* text_len;
    const int64_t* mel_len;
    int B, T1max, T2max;
    int16_t* hard;
    int64_t* dur;
    int16_t* path;        // (B, T1max): the caller's, or scratch in the workspace
    int path_is_output;   // the caller wants -1 past mel_len
    uint32_t* bits;       // (B, T1max, words) row-major bit matrix
    int words;            // ceil(T2max / 32)
    int* status;
};

__global__ void __launch_bounds__(kWT, 1)
mas_wide_kernel(const WideParams p) {
    extern __shared__ float s_q[];                       // [2][T2max + 1]: s_q[buf][j + 1] = Q[i][j], s_q[buf][0] = -inf
    __shared__ unsigned long long s_win[32];
    __shared__ int s_j;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    long long n64 = p.mel_len[b], m64 = p.text_len[b];
    if (tid == 0 && (n64 < 1 || n64 > p.T1max || m64 < 1 || m64 > p.T2max)) atomicAdd(p.status, 1);
    const int n = int(n64 < 1 ? 1 : (n64 > p.T1max ? p.T1max : n64));
    const int m = int(m64 < 1 ? 1 : (m64 > p.T2max ? p.T2max : m64));
    const int pitch = p.T2max + 1;
    const int ncol = (m + kWT - 1) / kWT;                // columns this utterance needs per thread
    const float* xb = p.logp + size_t(b) * p.sB;
    uint32_t* bits = p.bits + size_t(b) * p.T1max * p.words;

    // ---- dense output: zero fill of the utterance's (T1max, T2max) int16 block (the ones are written at the end) ----------
    if (p.hard) {
        int16_t* hb = p.hard + size_t(b) * p.T1max * p.T2max;
        const size_t total = size_t(p.T1max) * p.T2max;
        if ((reinterpret_cast<uintptr_t>(hb) & 15) == 0) {
            const size_t n16 = total >> 3;
            for (size_t i = tid; i < n16; i += kWT) st_cs_v4(reinterpret_cast<uint4*>(hb) + i, make_uint4(0u, 0u, 0u, 0u));
            for (size_t i = (n16 << 3) + tid; i < total; i += kWT) hb[i] = 0;
        } else {
            for (size_t i = tid; i < total; i += kWT) hb[i] = 0;
        }
    }
    if (p.dur) for (int j = tid; j < p.T2max; j += kWT) p.dur[size_t(b) * p.T2max + j] = 0;

    // ---- forward: row 0 (mas.py:11-12), then one barrier per row --------------------------------------------------------
    for (int j = tid; j <= p.T2max; j += kWT) { s_q[j] = -CUDART_INF_F; s_q[pitch + j] = -CUDART_INF_F; }
    __syncthreads();
    if (tid == 0) s_q[1] = xb[0];
    __syncthreads();
    for (int i = 1; i < n; ++i) {
        const float* prev = s_q + ((i - 1) & 1) * pitch;
        float* cur = s_q + (i & 1) * pitch;
        const float* xr = xb + size_t(i) * p.sT1;
        uint32_t* brow = bits + size_t(i) * p.words;
#pragma unroll 4
        for (int k = 0; k < ncol; ++k) {
            const int j = k * kWT + tid;
            bool diag = false;
            if (j < m) {
                const float a = prev[j];               // Q[i-1][j-1]   (prev[0] = -inf stands left of column 0)
                const float c = prev[j + 1];           // Q[i-1][j]
                diag = j > 0 && a >= c;                // mas.py:17 / cuda_mas.py:27: ties take the diagonal; column 0 stays
                cur[j + 1] = __fadd_rn(__ldg(xr + j), j > 0 ? fmaxf(a, c) : c);
            }
            const uint32_t w = __ballot_sync(0xffffffffu, diag);
            if (lane == 0 && (j >> 5) < p.words) brow[j >> 5] = w;
        }
        __syncthreads();
    }

    // ---- backtrack (mas.py:21-24): 32 rows per round, one lane per row fetches its 64-column window of bits ----------------
    int16_t* path = p.path + size_t(b) * p.T1max;
    __threadfence_block();
    if (tid < 32) {
        int j = m - 1;
        for (int i0 = n - 1; i0 >= 0; i0 -= 32) {
            const int r = i0 - lane;                    // this lane's row
            const int wj = j >> 5;
            unsigned long long win = 0;
            if (r >= 1) {
                const uint32_t hi = bits[size_t(r) * p.words + wj];
                const uint32_t lo = wj > 0 ? bits[size_t(r) * p.words + wj - 1] : 0u;
                win = (static_cast<unsigned long long>(hi) << 32) | lo;
            }
            s_win[lane] = win;
            __syncwarp();
            if (lane == 0) {
                const int base = (wj - 1) * 32;         // column of bit 0 of a window
                const int rows = min(32, i0 + 1);
                for (int l = 0; l < rows; ++l) {
                    path[i0 - l] = int16_t(j);
                    if (i0 - l >= 1) j -= int((s_win[l] >> (j - base)) & 1ull);
                }
                s_j = j;
            }
            __syncwarp();
            j = s_j;
        }
    }
    __syncthreads();

    // ---- outputs: ones, durations (alignment.py:275), -1 past the utterance for a returned path ---------------------------
    for (int i = tid; i < n; i += kWT) {
        const int j = path[i];
        if (p.hard) p.hard[(size_t(b) * p.T1max + i) * p.T2max + j] = 1;
        if (p.dur) {
            // the path is monotone: frame i starts a token's run when it is the first frame or the column changed
            if (i == 0 || path[i - 1] != j) {
                int e = i + 1;
                while (e < n && path[e] == j) ++e;
                p.dur[size_t(b) * p.T2max + j] = e - i;
            }
        }
    }
    if (p.path_is_output) for (int i = n + tid; i < p.T1max; i += kWT) path[i] = -1;
}

}  // namespace

size_t mas_wide_workspace_bytes(int B, int T1max, int T2max) {
    const size_t words = (size_t(T2max) + 31) / 32;
    return 256 + size_t(B) * T1max * words * 4 + ((size_t(B) * T1max * 2 + 15) & ~size_t(15));
}

bool mas_wide_supported(int T2max) { return T2max <= kWT * kWMaxCols; }

int mas_wide_forward(const float* logp, int64_t sB, int64_t sT1, const int64_t* text_len, const int64_t* mel_len, int B, int T1max,
                     int T2max, int16_t* attn_hard, int64_t* durations, int16_t* path, void* ws, cudaStream_t stream) {
    WideParams p;
    p.logp = logp; p.sB = sB; p.sT1 = sT1;
    p.text_len = text_len; p.mel_len = mel_len;
    p.B = B; p.T1max = T1max; p.T2max = T2max;
    p.hard = attn_hard; p.dur = durations;
    p.words = (T2max + 31) / 32;
    p.status = reinterpret_cast<int*>(ws);
    p.bits = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(ws) + 256);
    p.path = path ? path : reinterpret_cast<int16_t*>(reinterpret_cast<char*>(ws) + 256 + size_t(B) * T1max * p.words * 4);
    p.path_is_output = path ? 1 : 0;
    cudaError_t e = cudaMemsetAsync(ws, 0, 256, stream);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(status)");
    const size_t smem = 2 * (size_t(T2max) + 1) * sizeof(float);
    e = cudaFuncSetAttribute(mas_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(mas_wide_kernel)");
    mas_wide_kernel<<<B, kWT, smem, stream>>>(p);
    e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "mas_wide_kernel launch");
    return 0;
}

}  // namespace isp
