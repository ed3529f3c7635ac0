// Internal declarations shared by the kernels and the C-ABI layer.
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include "../../include/isp_tts_b200.h"

namespace isp {

void set_error(const char* fmt, ...);
int  cuda_fail(cudaError_t e, const char* what);

size_t mas_workspace_bytes(int B, int T1max, int T2max);
int    mas_forward(const float* logp, int64_t sB, int64_t sT1, int64_t sT2,
                   const int64_t* text_len, const int64_t* mel_len, int B, int T1max, int T2max,
                   int16_t* attn_hard, int64_t* durations, int16_t* path, void* ws, size_t ws_bytes, cudaStream_t stream,
                   const int* ready = nullptr, int ready_need = 0);
int    bin_loss_sums(const float* attn_soft, const int16_t* path, const int64_t* mel_len, int B, int T1max, int T2max,
                     float eps, float* sums, cudaStream_t stream);
int    mas_set_option(const char* key, int value, int* prev);
// second MAS kernel (isp_mas2.cu): T2max <= 256, backpointer words in shared memory
bool   mas2_supported(int B, int T1max, int T2max);
bool   mas2_linkable(int B, int T1max, int T2max, int dbg);
bool   mas_linkable(int B, int T1max, int T2max);
size_t mas2_workspace_bytes(int B);
int    mas2_forward(const float* logp, int64_t sB, int64_t sT1, const int64_t* text_len, const int64_t* mel_len,
                    int B, int T1max, int T2max, int16_t* attn_hard, int64_t* durations, int16_t* path, void* ws,
                    int no_tma, int ring_rows, int slots, int dbg, cudaStream_t stream, const int* ready = nullptr, int ready_need = 0);
int    mas2_set_option(const char* key, int value, int* prev);
// general kernel for wide utterances (isp_mas_wide.cu): any T2max <= ISP_MAS_WIDE_MAX_T2
bool   mas_wide_supported(int T2max);
size_t mas_wide_workspace_bytes(int B, int T1max, int T2max);
int    mas_wide_forward(const float* logp, int64_t sB, int64_t sT1, const int64_t* text_len, const int64_t* mel_len, int B, int T1max,
                        int T2max, int16_t* attn_hard, int64_t* durations, int16_t* path, void* ws, cudaStream_t stream);

// cluster kernel for long / wide utterances (isp_mas_cluster.cu): T2max <= 1024, one thread-block cluster per utterance
bool   mas_cluster_supported(int B, int T1max, int T2max);
size_t mas_cluster_workspace_bytes(int B, int T1max, int T2max);
int    mas_cluster_set_option(const char* key, int value, int* prev);
bool   mas_cluster_one_wave(int B, int T1max, int T2max);      // all B clusters co-resident (cudaOccupancyMaxActiveClusters)
int    mas_cluster_forward(const float* logp, int64_t sB, int64_t sT1, const int64_t* text_len, const int64_t* mel_len, int B, int T1max,
                           int T2max, int16_t* attn_hard, int64_t* durations, int16_t* path, void* ws, cudaStream_t stream);

size_t loglik_workspace_bytes(int B, int T1max, int T2max, int D, int dtype);
int    loglik_forward(const void* Q, const void* K, int dtype, const int64_t* text_len, const int64_t* mel_len,
                      int B, int T1max, int T2max, int D, float scale, int attention_prior,
                      float* attn_logits, float* attn_soft, void* ws, size_t ws_bytes, cudaStream_t stream, int* ready = nullptr);
bool   loglik_supported(int T2max, int D, int dtype);
int    loglik_tiles_per_utterance(int T1max);     // what `ready[b]` reaches when utterance b's outputs are complete
int    loglik_rows(const float* S, int64_t ldS, const int64_t* text_len, const int64_t* mel_len, int B, int T1max, int T2max,
                   float scale, int attention_prior, float* attn_logits, float* attn_soft, cudaStream_t stream);
int    split_3xtf32(const float* x, int64_t rows, int D, int role, float* out, cudaStream_t stream);
int    loglik_set_option(const char* key, int value, int* prev);
int    loglik_backward_ds(const float* S, const float* attn_soft, const float* g_logits, const float* g_soft,
                          int B, int T1max, int T2max, float scale, int attention_prior, void* dS, int ds_dtype, cudaStream_t stream);

int    loglik_backward_from_logits(const float* attn_logits, const float* g_logits, const float* g_soft, const float* prior_rowsum,
                                   const int64_t* text_len, const int64_t* mel_len, int B, int T1max, int T2max, float scale,
                                   int attention_prior, void* dS, int ds_dtype, cudaStream_t stream);

int    length_regulate(const void* x, const int16_t* path, void* out, int dtype, int B, int T1max, int T2max, int C, cudaStream_t stream);
int    length_regulate_backward(const float* g, const int64_t* durations, const int64_t* starts, float* gx,
                                int B, int T1max, int T2max, int C, cudaStream_t stream);

int    path_from_durations(const int64_t* reps, int16_t* path, int B, int T1max, int T2max, cudaStream_t stream);

int    temporal_average(const float* x, const int64_t* durations, float* out, int B, int C, int T1max, int T2max, cudaStream_t stream);
int    stage_operands(const void* q_host, const void* k_host, int dtype, const int64_t* text_len, const int64_t* mel_len,
                      int B, int T1max, int T2max, int D, void* q_dev, void* k_dev, cudaStream_t stream);

size_t unpack_workspace_bytes(int B);
int    unpack_operands(const void* q_packed, const void* k_packed, int dtype, const int64_t* text_len, const int64_t* mel_len,
                       int B, int T1max, int T2max, int D, void* q_dev, void* k_dev, void* ws, size_t ws_bytes, cudaStream_t stream);

size_t ctc_workspace_bytes(int B, int T1max, int T2max);
int    ctc_forward(const float* logits, const int64_t* text_len, const int64_t* mel_len, int B, int T1max, int T2max,
                   float blank_logprob, float* nll, void* ws, size_t ws_bytes, cudaStream_t stream);
int    ctc_backward(const float* logits, const int64_t* text_len, const int64_t* mel_len, int B, int T1max, int T2max,
                    float blank_logprob, const float* nll, const float* grad_scale, float* grad_logits, void* ws, size_t ws_bytes,
                    cudaStream_t stream);
size_t soft_average_workspace_bytes(int B, int C, int T1max, int T2max);
int    soft_average(const float* x, const float* attn_soft, const int64_t* row_len, float* out, float* colsum, int B, int C,
                    int T1max, int T2max, void* ws, size_t ws_bytes, cudaStream_t stream);
int    soft_average_backward(const float* g, const float* x, const float* out, const float* colsum, float* g_soft, int B, int C,
                             int T1max, int T2max, cudaStream_t stream);
int    prep_channels_last(const void* x, int in_dtype, int channels_first, const int64_t* len, void* out, int out_dtype,
                          int B, int C, int T, int Cp, cudaStream_t stream);
int    instance_norm_apply(const void* y, int dtype, const float* stats, int parts, const float* weight, const float* bias,
                           const int64_t* len, void* out, int B, int T, int C, int64_t ld_in, int64_t ld_out, float eps, void* ws, cudaStream_t stream);
int    gemm_batched(const isp_gemm_desc* d, cudaStream_t stream);
int    stage_set_option(const char* key, int value, int* prev);

}  // namespace isp
