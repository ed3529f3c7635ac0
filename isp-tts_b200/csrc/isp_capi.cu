// C ABI (include/isp_tts_b200.h): argument checks, error strings, dispatch.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include "isp_internal.h"

namespace isp {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
    set_error("%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
    return int(e);
}

}  // namespace isp

extern "C" {

int isp_version(void) { return 100; }  // 0.1.0

const char* isp_last_error(void) { return isp::g_err; }

int isp_device_check(void) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return isp::cuda_fail(e, "cudaGetDevice");
    int major = 0;
    e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e != cudaSuccess) return isp::cuda_fail(e, "cudaDeviceGetAttribute");
    if (major != 10) {
        isp::set_error("device %d has compute capability %d.x; these kernels are built for sm_100a only", dev, major);
        return ISP_ERR_DEVICE;
    }
    return ISP_OK;
}

size_t isp_mas_workspace_bytes(int B, int T1max, int T2max) { return isp::mas_workspace_bytes(B, T1max, T2max); }

int isp_mas_forward(const float* logp, int64_t sB, int64_t sT1, int64_t sT2, const int64_t* text_len,
                    const int64_t* mel_len, int B, int T1max, int T2max, int16_t* attn_hard, int64_t* durations,
                    void* ws, size_t ws_bytes, void* stream) {
    return isp::mas_forward(logp, sB, sT1, sT2, text_len, mel_len, B, T1max, T2max, attn_hard, durations, nullptr, ws,
                            ws_bytes, static_cast<cudaStream_t>(stream));
}

int isp_mas_forward_path(const float* logp, int64_t sB, int64_t sT1, int64_t sT2, const int64_t* text_len,
                         const int64_t* mel_len, int B, int T1max, int T2max, int16_t* attn_hard, int64_t* durations,
                         int16_t* path, void* ws, size_t ws_bytes, void* stream) {
    if (!path) { isp::set_error("isp_mas_forward_path: null path"); return ISP_ERR_INVALID; }
    return isp::mas_forward(logp, sB, sT1, sT2, text_len, mel_len, B, T1max, T2max, attn_hard, durations, path, ws,
                            ws_bytes, static_cast<cudaStream_t>(stream));
}

int isp_bin_loss_sums(const float* attn_soft, const int16_t* path, const int64_t* mel_len, int B, int T1max, int T2max,
                      float eps, float* sums, void* stream) {
    return isp::bin_loss_sums(attn_soft, path, mel_len, B, T1max, T2max, eps, sums, static_cast<cudaStream_t>(stream));
}

int isp_mas_status(const void* ws, void* stream) {
    if (!ws) { isp::set_error("isp_mas_status: null workspace"); return -1; }
    int v[2] = {-1, 0};
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e = cudaMemcpyAsync(v, ws, sizeof(v), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { isp::cuda_fail(e, "isp_mas_status"); return -1; }
    if (v[0] != -2) return v[0];                        // isp_mas.cu: a counter
    // isp_mas2.cu: one flag per utterance behind the order array (see mas2_workspace_bytes)
    const int B = v[1];
    if (B <= 0) return -1;
    unsigned char* flags = static_cast<unsigned char*>(malloc(size_t(B)));
    if (!flags) return -1;
    const char* src = static_cast<const char*>(ws) + 256 + ((size_t(B) * 4 + 15) & ~size_t(15));
    e = cudaMemcpyAsync(flags, src, size_t(B), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    int bad = 0;
    if (e == cudaSuccess) for (int i = 0; i < B; ++i) bad += flags[i] != 0;
    free(flags);
    if (e != cudaSuccess) { isp::cuda_fail(e, "isp_mas_status"); return -1; }
    return bad;
}

static int g_align_mode = 0;     // experiments: 0 linked; 1 counts published, MAS launched plainly behind; 2 the plain sequence

// workspace of isp_align_forward: the MAS workspace, then (256 B aligned) the per-utterance ready counts
static size_t align_ready_offset(int B, int T1max, int T2max) { return (isp::mas_workspace_bytes(B, T1max, T2max) + 255) & ~size_t(255); }

size_t isp_align_workspace_bytes(int B, int T1max, int T2max, int D, int dtype) {
    (void)D; (void)dtype;
    if (B <= 0 || T1max <= 0 || T2max <= 0) return 0;
    return align_ready_offset(B, T1max, T2max) + ((size_t(B + 8) * sizeof(int) + 255) & ~size_t(255));       // + the launch's time origin
}

int isp_align_forward(const void* Q, const void* K, int dtype, const int64_t* text_len, const int64_t* mel_len,
                      int B, int T1max, int T2max, int D, float scale, int attention_prior,
                      float* attn_logits, float* attn_soft, int16_t* attn_hard, int64_t* durations, int16_t* path, float* prior_rowsum,
                      void* ws, size_t ws_bytes, int flags, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (!ws || B <= 0 || T1max <= 0 || T2max <= 0) { isp::set_error("isp_align_forward: null workspace or non-positive sizes"); return ISP_ERR_INVALID; }
    if (ws_bytes < isp_align_workspace_bytes(B, T1max, T2max, D, dtype) || (reinterpret_cast<uintptr_t>(ws) & 15)) {
        isp::set_error("isp_align_forward: workspace too small or not 16 B aligned (%zu < %zu)", ws_bytes, isp_align_workspace_bytes(B, T1max, T2max, D, dtype));
        return ISP_ERR_WORKSPACE;
    }
    const size_t off = align_ready_offset(B, T1max, T2max);
    int* ready = reinterpret_cast<int*>(static_cast<char*>(ws) + off);
    const bool link = g_align_mode != 2 && isp::mas_linkable(B, T1max, T2max);
    const size_t ready_bytes = size_t(B + 8) * sizeof(int);      // counts, the MAS launch's time origin, two debug stamps
    if (link && !(flags & ISP_ALIGN_WS_CLEAN)) {
        // the counts are cleared in front of the first kernel, so that the second follows the first directly (a programmatic edge);
        // a workspace that the previous successful call on it left behind is clean already (the MAS kernel resets what it consumed)
        cudaError_t e = cudaMemsetAsync(ready, 0, ready_bytes, st);
        if (e != cudaSuccess) return isp::cuda_fail(e, "cudaMemsetAsync(ready counts)");
    }
    int rc = isp::loglik_forward(Q, K, dtype, text_len, mel_len, B, T1max, T2max, D, scale, attention_prior, attn_logits, attn_soft,
                                 prior_rowsum, prior_rowsum ? size_t(B) * T1max * sizeof(float) : 0, st, link ? ready : nullptr);
    if (rc) return rc;
    rc = isp::mas_forward(attn_logits, int64_t(T1max) * T2max, T2max, 1, text_len, mel_len, B, T1max, T2max, attn_hard, durations, path,
                          ws, off, st, link && g_align_mode == 0 ? ready : nullptr, isp::loglik_tiles_per_utterance(T1max));
    if ((rc || g_align_mode == 1) && link) cudaMemsetAsync(ready, 0, ready_bytes, st);      // nobody consumed the counts: leave the workspace clean
    return rc;
}

int isp_loglik_supported(int T2max, int D, int dtype) { return isp::loglik_supported(T2max, D, dtype) ? 1 : 0; }

size_t isp_loglik_workspace_bytes(int B, int T1max, int T2max, int D, int dtype) {
    return isp::loglik_workspace_bytes(B, T1max, T2max, D, dtype);
}

int isp_loglik_forward(const void* Q, const void* K, int dtype, const int64_t* text_len, const int64_t* mel_len,
                       int B, int T1max, int T2max, int D, float scale, int attention_prior, float* attn_logits,
                       float* attn_soft, void* ws, size_t ws_bytes, void* stream) {
    return isp::loglik_forward(Q, K, dtype, text_len, mel_len, B, T1max, T2max, D, scale, attention_prior,
                               attn_logits, attn_soft, ws, ws_bytes, static_cast<cudaStream_t>(stream));
}

int isp_split_3xtf32(const float* x, int64_t rows, int D, int role, float* out, void* stream) {
    return isp::split_3xtf32(x, rows, D, role, out, static_cast<cudaStream_t>(stream));
}
int isp_loglik_rows(const float* S, int64_t ldS, const int64_t* text_len, const int64_t* mel_len, int B, int T1max, int T2max,
                    float scale, int attention_prior, float* attn_logits, float* attn_soft, void* stream) {
    return isp::loglik_rows(S, ldS, text_len, mel_len, B, T1max, T2max, scale, attention_prior, attn_logits, attn_soft, static_cast<cudaStream_t>(stream));
}

int isp_loglik_backward_from_logits(const float* attn_logits, const float* g_logits, const float* g_soft, const float* prior_rowsum,
                                    const int64_t* text_len, const int64_t* mel_len, int B, int T1max, int T2max, float scale,
                                    int attention_prior, void* dS, int ds_dtype, void* stream) {
    return isp::loglik_backward_from_logits(attn_logits, g_logits, g_soft, prior_rowsum, text_len, mel_len, B, T1max, T2max, scale,
                                            attention_prior, dS, ds_dtype, static_cast<cudaStream_t>(stream));
}

int isp_loglik_backward_ds(const float* S, const float* attn_soft, const float* g_logits, const float* g_soft,
                           int B, int T1max, int T2max, float scale, int attention_prior, void* dS, int ds_dtype, void* stream) {
    return isp::loglik_backward_ds(S, attn_soft, g_logits, g_soft, B, T1max, T2max, scale, attention_prior, dS, ds_dtype,
                                   static_cast<cudaStream_t>(stream));
}

int isp_length_regulate(const void* x, const int16_t* path, void* out, int dtype, int B, int T1max, int T2max, int C, void* stream) {
    return isp::length_regulate(x, path, out, dtype, B, T1max, T2max, C, static_cast<cudaStream_t>(stream));
}

int isp_length_regulate_backward(const float* g, const int64_t* durations, const int64_t* starts, float* gx,
                                 int B, int T1max, int T2max, int C, void* stream) {
    return isp::length_regulate_backward(g, durations, starts, gx, B, T1max, T2max, C, static_cast<cudaStream_t>(stream));
}

int isp_path_from_durations(const int64_t* reps, int16_t* path, int B, int T1max, int T2max, void* stream) {
    return isp::path_from_durations(reps, path, B, T1max, T2max, static_cast<cudaStream_t>(stream));
}

int isp_temporal_average(const float* x, const int64_t* durations, float* out, int B, int C, int T1max, int T2max, void* stream) {
    return isp::temporal_average(x, durations, out, B, C, T1max, T2max, static_cast<cudaStream_t>(stream));
}

size_t isp_ctc_workspace_bytes(int B, int T1max, int T2max) { return isp::ctc_workspace_bytes(B, T1max, T2max); }

int isp_ctc_forward(const float* attn_logits, const int64_t* text_len, const int64_t* mel_len, int B, int T1max, int T2max,
                    float blank_logprob, float* nll, void* ws, size_t ws_bytes, void* stream) {
    return isp::ctc_forward(attn_logits, text_len, mel_len, B, T1max, T2max, blank_logprob, nll, ws, ws_bytes, static_cast<cudaStream_t>(stream));
}

int isp_ctc_backward(const float* attn_logits, const int64_t* text_len, const int64_t* mel_len, int B, int T1max, int T2max,
                     float blank_logprob, const float* nll, const float* grad_scale, float* grad_logits, void* ws, size_t ws_bytes,
                     void* stream) {
    return isp::ctc_backward(attn_logits, text_len, mel_len, B, T1max, T2max, blank_logprob, nll, grad_scale, grad_logits, ws, ws_bytes,
                             static_cast<cudaStream_t>(stream));
}

int isp_stage_operands(const void* q_host, const void* k_host, int dtype, const int64_t* text_len, const int64_t* mel_len,
                       int B, int T1max, int T2max, int D, void* q_dev, void* k_dev, void* stream) {
    return isp::stage_operands(q_host, k_host, dtype, text_len, mel_len, B, T1max, T2max, D, q_dev, k_dev, static_cast<cudaStream_t>(stream));
}

size_t isp_soft_average_workspace_bytes(int B, int C, int T1max, int T2max) { return isp::soft_average_workspace_bytes(B, C, T1max, T2max); }

int isp_soft_average(const float* x, const float* attn_soft, const int64_t* row_len, float* out, float* colsum, int B, int C,
                     int T1max, int T2max, void* ws, size_t ws_bytes, void* stream) {
    return isp::soft_average(x, attn_soft, row_len, out, colsum, B, C, T1max, T2max, ws, ws_bytes, static_cast<cudaStream_t>(stream));
}

int isp_soft_average_backward(const float* g, const float* x, const float* out, const float* colsum, float* g_soft, int B, int C,
                              int T1max, int T2max, void* stream) {
    return isp::soft_average_backward(g, x, out, colsum, g_soft, B, C, T1max, T2max, static_cast<cudaStream_t>(stream));
}

size_t isp_unpack_workspace_bytes(int B) { return isp::unpack_workspace_bytes(B); }

int isp_unpack_operands(const void* q_packed, const void* k_packed, int dtype, const int64_t* text_len, const int64_t* mel_len,
                        int B, int T1max, int T2max, int D, void* q_dev, void* k_dev, void* ws, size_t ws_bytes, void* stream) {
    return isp::unpack_operands(q_packed, k_packed, dtype, text_len, mel_len, B, T1max, T2max, D, q_dev, k_dev, ws, ws_bytes, static_cast<cudaStream_t>(stream));
}

int isp_prep_channels_last(const void* x, int in_dtype, int channels_first, const int64_t* len, void* out, int out_dtype,
                           int B, int C, int T, int Cp, void* stream) {
    return isp::prep_channels_last(x, in_dtype, channels_first, len, out, out_dtype, B, C, T, Cp, static_cast<cudaStream_t>(stream));
}

int isp_instance_norm_apply(const void* y, int dtype, const float* stats, int parts, const float* weight, const float* bias,
                            const int64_t* len, void* out, int B, int T, int C, int64_t ld_in, int64_t ld_out, float eps, void* ws, void* stream) {
    return isp::instance_norm_apply(y, dtype, stats, parts, weight, bias, len, out, B, T, C, ld_in, ld_out, eps, ws, static_cast<cudaStream_t>(stream));
}

int isp_gemm_batched(const isp_gemm_desc* desc, void* stream) { return isp::gemm_batched(desc, static_cast<cudaStream_t>(stream)); }

int isp_set_option(const char* key, int value) {
    if (!key) return ISP_ERR_INVALID;
    int prev = 0;
    if (!strcmp(key, "align.mode")) { prev = g_align_mode; g_align_mode = value; return prev; }
    if (isp::mas_set_option(key, value, &prev) == 0) return prev;
    if (isp::loglik_set_option(key, value, &prev) == 0) return prev;
    if (isp::stage_set_option(key, value, &prev) == 0) return prev;
    isp::set_error("isp_set_option: unknown key '%s'", key);
    return ISP_ERR_INVALID;
}

}  // extern "C"
