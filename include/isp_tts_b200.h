/*
 * isp_tts_b200.h -- C ABI of the B200-native Aligner hot path of ilya16/isp-tts.
 *
 * The reference is pure Python; its "FFI" for this path is the pair of numba
 * entry points that tts/models/acoustic/modules/alignment.py calls, plus the
 * torch ops around them.  Each function below states the reference interface
 * it replaces (paths relative to the reference root).  INTEGRATION.md shows the
 * ctypes stubs a maintainer adds on the reference side.
 *
 * Conventions (all functions):
 *   - plain pointers and sizes; no torch / numba types;
 *   - every device buffer, workspace included, is allocated and owned by the
 *     caller; the library never allocates, frees or synchronises;
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*; NULL = the
 *     legacy default stream, which is what numba and torch use in the reference);
 *   - pointers must belong to the current CUDA device, which must be sm_100;
 *   - lengths are int64 and are read ON THE DEVICE (no host sync);
 *     the contract is 1 <= len <= Tmax; out-of-range lengths are clamped for
 *     memory safety and reported through isp_mas_status();
 *   - return value: 0 = enqueued; < 0 = ISP_ERR_* (bad argument / unsupported
 *     shape / workspace too small); > 0 = a cudaError_t.  isp_last_error()
 *     returns a thread-local message for the last non-zero return;
 *   - re-entrant, PROVIDED nobody calls isp_set_option concurrently: the tuning knobs
 *     behind it are unsynchronised process-wide globals meant for benchmarks and tests;
 *     the only other process-wide state is a per-device cudaFuncSetAttribute.
 * There is no CPU fallback anywhere behind this ABI.
 */
#ifndef ISP_TTS_B200_H
#define ISP_TTS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ISP_OK               0
#define ISP_ERR_INVALID     (-1)  /* null pointer, non-positive size, bad stride/alignment */
#define ISP_ERR_UNSUPPORTED (-2)  /* shape outside what the kernels cover (see each function) */
#define ISP_ERR_WORKSPACE   (-3)  /* ws_bytes < isp_*_workspace_bytes(...) */
#define ISP_ERR_DEVICE      (-4)  /* current device is not compute capability 10.x */

/* element types of the encoded operands of isp_loglik_forward */
#define ISP_DTYPE_F32  0          /* fp32 in memory, TF32 tensor-core products, fp32 accumulate */
#define ISP_DTYPE_BF16 1          /* bf16 in memory, fp32 accumulate */
#define ISP_DTYPE_F16  2          /* fp16 in memory, fp32 accumulate (isp_gemm_batched and its element-wise companions only) */

#define ISP_MAS_MAX_T2   640      /* text tokens per utterance the fast (strip) MAS kernels cover */
#define ISP_MAS_CLUSTER_MAX_T2 1024 /* beyond ISP_MAS_MAX_T2: one thread-block cluster per utterance (up to 8 CTAs of 128 tokens, DSMEM) */
#define ISP_MAS_WIDE_MAX_T2 16384 /* beyond ISP_MAS_CLUSTER_MAX_T2 a general, slower kernel runs: up to this many tokens */
#define ISP_LOGLIK_MAX_T2 512     /* text tokens per utterance the fused GEMM covers (TMEM columns) */
#define ISP_LOGLIK_MAX_D  256     /* attention_dim */

int         isp_version(void);            /* 10000*major + 100*minor + patch */
const char* isp_last_error(void);
/* 0 if the current device can run the kernels (sm_100), else ISP_ERR_DEVICE / cudaError_t */
int         isp_device_check(void);

/* ------------------------------------------------------------------------------------------
 * Monotonic Alignment Search + hard path + durations.
 *
 * Replaces, in one launch:
 *   cuda_b_mas[(max(64,B),1),(1,256)](log_p, prev_log_p, prev_ind, attn_out, in_lens, out_lens)
 *       tts/modules/aligner/cuda_mas.py:11-46, launched at
 *       tts/models/acoustic/modules/alignment.py:328-330 (with the clone / three
 *       zeros_like at :321-326, which are not needed here), and its CPU twin
 *   b_mas(b_attn_map, in_lens, out_lens)            tts/modules/aligner/mas.py:30-35
 *   attn_hard.sum(dim=1)                            alignment.py:275
 *
 * logp      (B, T1max, T2max) fp32 log-likelihoods, element strides sB/sT1/sT2;
 *           sT2 must be 1.  NOT modified (the reference's GPU route works on a clone).
 *           Rows are frames (mel), columns are text tokens.
 * text_len  (B) int64 = in_lens;  mel_len (B) int64 = out_lens.
 * attn_hard (B, T1max, T2max) int16, contiguous, FULLY written: exactly one 1 per
 *           valid frame, 0 elsewhere (padding included) -- it need not be pre-zeroed.
 * durations (B, T2max) int64, contiguous, fully written; may be NULL.
 * ws        isp_mas_workspace_bytes(B,T1max,T2max) bytes, 16 B aligned; holds the status word,
 *           the packed backpointer bits when they do not fit in shared memory, and the
 *           path's column per frame until the zero-fill of attn_hard has landed.
 * Limits:   T2max <= ISP_MAS_WIDE_MAX_T2.  Up to ISP_MAS_MAX_T2 = 640 tokens an utterance runs on one SM (strip wavefront);
 *           641 .. ISP_MAS_CLUSTER_MAX_T2 = 1024 tokens on a thread-block cluster, the strip boundary and the backtrack maps
 *           crossing CTAs through distributed shared memory (T1max <= ~8000 frames: the backpointer bits stay in shared
 *           memory); beyond, a general kernel (one barrier per frame row).  Same results everywhere.  T1max < 2^24.
 * Results are bit-identical to the reference for NaN-free input.
 */
size_t isp_mas_workspace_bytes(int B, int T1max, int T2max);
int    isp_mas_forward(const float* logp, int64_t sB, int64_t sT1, int64_t sT2,
                       const int64_t* text_len, const int64_t* mel_len,
                       int B, int T1max, int T2max,
                       int16_t* attn_hard, int64_t* durations,
                       void* ws, size_t ws_bytes, void* stream);
/* Same, and also the path as one column index per frame: path (B, T1max) int16, -1 for frames past mel_len.
 * Consumers that only need "which token does frame i belong to" (the binarization loss below, the length
 * regulator, per-token averaging) read this instead of the dense attn_hard.  Here attn_hard may be NULL: the dense
 * (B, T1max, T2max) int16 tensor is then neither zero-filled nor written (2 B/cell less HBM traffic). */
int    isp_mas_forward_path(const float* logp, int64_t sB, int64_t sT1, int64_t sT2,
                            const int64_t* text_len, const int64_t* mel_len,
                            int B, int T1max, int T2max,
                            int16_t* attn_hard, int64_t* durations, int16_t* path,
                            void* ws, size_t ws_bytes, void* stream);
/* Attention binarization loss from the path (tts/models/acoustic/loss.py:97-105):
 * sums[0] = sum over valid frames of log(max(attn_soft[b, i, path[b, i]], eps)), sums[1] = number of valid frames;
 * the loss is -sums[0] / sums[1].  sums: 2 floats on the device, zeroed by the call. */
int    isp_bin_loss_sums(const float* attn_soft, const int16_t* path, const int64_t* mel_len,
                         int B, int T1max, int T2max, float eps, float* sums, void* stream);
/* Length regulator from the path (tts/models/acoustic/modules/temporal_adaptor.py:411-436, `durations` branch):
 * out[b, t, :] = x[b, path[b, t], :] for the utterance's frames, 0 after.  x (B, T2max, C), out (B, T1max, C), fp32 or
 * bf16 (dtype = ISP_DTYPE_*), contiguous, 16 B aligned, C * elem % 16 == 0.
 * Backward (fp32): gx[b, j, :] = sum of g[b, t, :] over starts[b, j] <= t < starts[b, j] + durations[b, j]
 * (starts = exclusive cumulative sum of durations along the token axis; both (B, T2max) int64). */
int    isp_length_regulate(const void* x, const int16_t* path, void* out, int dtype,
                           int B, int T1max, int T2max, int C, void* stream);
int    isp_length_regulate_backward(const float* g, const int64_t* durations, const int64_t* starts, float* gx,
                                    int B, int T1max, int T2max, int C, void* stream);
/* The path (token index per frame, -1 past the utterance's total) from rounded durations, for callers without a MAS path:
 * the reference's inference route builds a (T1 x T2) 0/1 matrix from the cumulated durations instead
 * (tts/models/acoustic/modules/temporal_adaptor.py:424-431).  reps (B, T2max) int64 = (durations + 0.5) truncated;
 * path (B, T1max) int16. */
int    isp_path_from_durations(const int64_t* reps, int16_t* path, int B, int T1max, int T2max, void* stream);
/* Per-token average of frame-level features (tts/models/acoustic/modules/temporal_adaptor.py:439-465, `durations` branch):
 * out[b, c, j] = sum of x[b, c, t] over the token's frames / count of non-zero x among them (0 if none).
 * x (B, C, T1max) fp32, durations (B, T2max) int64 (the MAS durations), out (B, C, T2max) fp32, all contiguous. */
int    isp_temporal_average(const float* x, const int64_t* durations, float* out, int B, int C, int T1max, int T2max, void* stream);
/* The same average on the recipe's soft route (temporal_adaptor.py:446-449, `alignment` = the Aligner's attn_soft,
 * tts/models/acoustic/model.py:154):  out[b, c, j] = sum_t x[b, c, t] * a[b, t, j] / (sum_t a[b, t, j] + 1e-5).
 * One streaming pass over attn_soft in fp32 (no tensor-core rounding), deterministic two-stage sum.
 * x (B, C, T1max) fp32, C <= 4; attn_soft (B, T1max, T2max) fp32, T2max % 4 == 0, 16 B aligned; row_len optional DEVICE
 * int64 (B,): rows >= row_len[b] are known to be zero and are not read; out (B, C, T2max); colsum optional (B, T2max), the
 * denominators without the 1e-5 (the backward needs them); ws: isp_soft_average_workspace_bytes(...) bytes, 16 B aligned.
 * Backward: g_soft[b, t, j] = sum_c g[b, c, j] / (colsum[b, j] + 1e-5) * (x[b, c, t] - out[b, c, j]). */
size_t isp_soft_average_workspace_bytes(int B, int C, int T1max, int T2max);
int    isp_soft_average(const float* x, const float* attn_soft, const int64_t* row_len, float* out, float* colsum,
                        int B, int C, int T1max, int T2max, void* ws, size_t ws_bytes, void* stream);
int    isp_soft_average_backward(const float* g, const float* x, const float* out, const float* colsum, float* g_soft,
                                 int B, int C, int T1max, int T2max, void* stream);
/* Forward-sum (CTC) alignment loss, tts/models/acoustic/loss.py:41-79 (AttentionCTCLoss.forward): blank column of value
 * blank_logprob in front of attn_logits (:67), log_softmax over the T2max + 1 columns (:69), CTC with targets
 * 1 .. text_len[b] over mel_len[b] frames (:73-78).  isp_ctc_forward writes nll[b] (fp32, natural log; +inf when no
 * alignment exists, i.e. mel_len[b] < text_len[b]) and keeps the forward variables in `ws`; isp_ctc_backward, called with
 * the same `ws` afterwards, writes grad_logits[b, i, j] = grad_scale[b] * d nll[b] / d attn_logits[b, i, j] for the whole
 * (B, T1max, T2max) tensor (zeros past mel_len[b]; all zeros for an utterance whose nll is +inf -- zero_infinity).  The
 * reference's loss is mean_b(nll[b] / text_len[b]), so grad_scale[b] = upstream / (B * text_len[b]).
 * attn_logits (B, T1max, T2max) fp32 contiguous; ws 16 B aligned, isp_ctc_workspace_bytes(...) bytes; T2max <= 639. */
size_t isp_ctc_workspace_bytes(int B, int T1max, int T2max);
int    isp_ctc_forward(const float* attn_logits, const int64_t* text_len, const int64_t* mel_len, int B, int T1max, int T2max,
                       float blank_logprob, float* nll, void* ws, size_t ws_bytes, void* stream);
int    isp_ctc_backward(const float* attn_logits, const int64_t* text_len, const int64_t* mel_len, int B, int T1max, int T2max,
                        float blank_logprob, const float* nll, const float* grad_scale, float* grad_logits,
                        void* ws, size_t ws_bytes, void* stream);
/* Reads back (synchronously, after the stream drains) how many utterances had a length
 * outside [1, Tmax] in the last isp_mas_forward that used `ws`.  -1 on error. */
int    isp_mas_status(const void* ws, void* stream);

/* ------------------------------------------------------------------------------------------
 * Pairwise log-likelihood: batched Q.K^T GEMM with the whole epilogue fused.
 *
 * Replaces ConvAttention.forward from the matmul on
 *   tts/models/acoustic/modules/alignment.py:189-208 :
 *   torch.matmul (:189), scale (:190), clamp (:192), batch_diagonal_prior (:18-37, :195),
 *   log_softmax over ALL T2max columns + log(prior + 1e-6) (:196), clone (:198),
 *   masked_fill + softmax (:201-203), mask multiply (:206).
 *
 * Q  (B, T1max, D)  = queries_enc.transpose(1,2) of :187, row-major, D contiguous
 * K  (B, T2max, D)  = keys_enc.transpose(1,2)    of :182, row-major, D contiguous
 *    dtype ISP_DTYPE_F32 or ISP_DTYPE_BF16; base 16 B aligned; D % 8 == 0, D <= ISP_LOGLIK_MAX_D.
 * scale             = attention_dim ** -0.5 (:116)
 * attn_logits, attn_soft  (B, T1max, T2max) fp32 contiguous, fully written (:208).
 * attention_prior   non-zero = the reference default (:194); 0 = plain scaled scores.
 * ws                isp_loglik_workspace_bytes(...) = 4 * B * T1max bytes, 4 B aligned, or NULL with ws_bytes = 0: receives the
 *                   row sums of the un-normalised prior, (B, T1max) fp32, valid frames only -- the one thing
 *                   isp_loglik_backward_from_logits cannot re-derive bit for bit from attn_logits.
 * Limits: T2max <= ISP_LOGLIK_MAX_T2.
 * Accuracy: within 1e-3 relative of the fp32 reference (tests state the tolerance).
 */
/* Ragged host -> device staging of Q and K (what the reference does with batch.to(device), tts/utils/trainer.py, for whole
 * padded tensors).  q_host (B, T1max, D), k_host (B, T2max, D): PINNED host memory, same layout and dtype as the device
 * tensors; text_len, mel_len: DEVICE int64 (B,), already uploaded on `stream`.  Only rows below the lengths are read
 * over PCIe; padding rows of q_dev / k_dev are written as zeros (the operand contract of isp_loglik_forward).
 * D * elem % 16 == 0, all tensors 16 B aligned. */
int    isp_stage_operands(const void* q_host, const void* k_host, int dtype, const int64_t* text_len, const int64_t* mel_len,
                          int B, int T1max, int T2max, int D, void* q_dev, void* k_dev, void* stream);
/* The same operands arriving PACKED and already on the device: q_packed holds the valid rows of all utterances back to back
 * (sum of mel_len rows of D elements; utterance b starts at the exclusive running sum of the lengths), k_packed likewise for
 * text_len.  One plain cudaMemcpyAsync of each packed buffer from pinned host memory runs on a copy engine -- the transfer that
 * scales when several GPUs share the host -- and this call scatters the rows into the padded (B, T, D) operands, padding rows
 * zero-filled.  ws: isp_unpack_workspace_bytes(B) bytes, 8 B aligned. */
size_t isp_unpack_workspace_bytes(int B);
int    isp_unpack_operands(const void* q_packed, const void* k_packed, int dtype, const int64_t* text_len, const int64_t* mel_len,
                           int B, int T1max, int T2max, int D, void* q_dev, void* k_dev, void* ws, size_t ws_bytes, void* stream);
/* The whole hot path in one call -- what Aligner.forward does between the projection stacks and its return
 * (tts/models/acoustic/modules/alignment.py:253 self.attention(...), :267-275 binarize_attention + durations): the two kernels of
 * isp_loglik_forward and isp_mas_forward[_path], linked so that the second does not wait for the first to drain.  The
 * log-likelihood kernel counts, per utterance, the frame tiles whose outputs are in global memory (ws); the MAS kernel is
 * launched with programmatic stream serialisation, becomes resident on the SMs the last wave of the first kernel leaves
 * free, and each of its CTAs starts its utterance as soon as that utterance's count is final.  Arguments, outputs, limits and
 * results are those of the two separate calls (bit-identical); shapes the linked kernels do not cover run as the plain
 * sequence.  path may be NULL; attn_hard may be NULL when path is not.  prior_rowsum: (B, T1max) fp32 or NULL -- what
 * isp_loglik_forward leaves in its workspace for isp_loglik_backward_from_logits.  ws: isp_align_workspace_bytes(...) bytes, 16 B aligned;
 * isp_mas_status(ws, stream) works on it as after isp_mas_forward.  flags: 0, or ISP_ALIGN_WS_CLEAN when the last thing that
 * touched ws was a successful isp_align_forward with the same B, T1max, T2max in the same stream (the call leaves its
 * counters cleared, so the next one needs no memset in front of the kernels: one graph node and ~3 us less per step). */
#define ISP_ALIGN_WS_CLEAN 1
size_t isp_align_workspace_bytes(int B, int T1max, int T2max, int D, int dtype);
int    isp_align_forward(const void* Q, const void* K, int dtype, const int64_t* text_len, const int64_t* mel_len,
                         int B, int T1max, int T2max, int D, float scale, int attention_prior,
                         float* attn_logits, float* attn_soft, int16_t* attn_hard, int64_t* durations, int16_t* path,
                         float* prior_rowsum, void* ws, size_t ws_bytes, int flags, void* stream);
/* 1 when isp_loglik_forward (and isp_align_forward's linked kernels) cover the shape: D a multiple of 8 and <= ISP_LOGLIK_MAX_D,
 * T2max <= ISP_LOGLIK_MAX_T2, and the operands of one frame tile plus one utterance's tokens fit in shared memory (fp32 operands
 * with several hundred tokens at D = 128 do not).  Other shapes: isp_gemm_batched for the scores, then isp_loglik_rows. */
int    isp_loglik_supported(int T2max, int D, int dtype);
size_t isp_loglik_workspace_bytes(int B, int T1max, int T2max, int D, int dtype);
int    isp_loglik_forward(const void* Q, const void* K, int dtype,
                          const int64_t* text_len, const int64_t* mel_len,
                          int B, int T1max, int T2max, int D, float scale, int attention_prior,
                          float* attn_logits, float* attn_soft,
                          void* ws, size_t ws_bytes, void* stream);

/* The same epilogue for shapes isp_loglik_forward does not cover (T2max > ISP_LOGLIK_MAX_T2, attention_dim not a multiple of 8 or
 * above ISP_LOGLIK_MAX_D): S (B, T1max, T2max; row stride ldS) are the UNscaled scores Q.K^T from isp_gemm_batched; the call
 * does alignment.py:190-208 on them (scale, log_softmax over all T2max columns + log prior, masked softmax) one warp per frame
 * row.  Slower than the fused kernel (the scores make a round trip through HBM), same results within the same tolerance. */
/* Faithful fp32 products on the tensor cores ("3xTF32"): the reference's torch.matmul (alignment.py:189) runs true fp32
 * (allow_tf32 is off by default), while one kind::tf32 product keeps 10 mantissa bits.  x (rows, D) fp32 -> out (rows, 3 D):
 * role 0 (frames / Q side) [hi | hi | lo], role 1 (tokens / K side) [hi | lo | hi] with hi = tf32(x), lo = tf32(x - hi), so that
 * ONE TF32 contraction over 3 D (isp_gemm_batched) gives hi.hi' + hi.lo' + lo.hi' = the fp32 product to ~2^-21; isp_loglik_rows
 * then does the epilogue.  (What ConvAttention runs outside autocast; gemm_dtype = "tf32" selects the fused single-pass kernel.) */
int    isp_split_3xtf32(const float* x, int64_t rows, int D, int role, float* out, void* stream);
int    isp_loglik_rows(const float* S, int64_t ldS, const int64_t* text_len, const int64_t* mel_len, int B, int T1max, int T2max,
                       float scale, int attention_prior, float* attn_logits, float* attn_soft, void* stream);

/* ------------------------------------------------------------------------------------------
 * Backward of the log-likelihood, first stage: the gradient with respect to the scores S = Q.K^T.
 *
 * Replaces what autograd records for tts/models/acoustic/modules/alignment.py:190-206 (scale,
 * log_softmax over all T2max columns, + log prior, masked softmax): given
 *   S          (B, T1max, T2max) fp32, the UNscaled scores Q.K^T (recomputed by a library GEMM)
 *   attn_soft  (B, T1max, T2max) fp32, as returned by isp_loglik_forward
 *   g_logits   dL/d attn_logits, g_soft  dL/d attn_soft  (either may be NULL, not both)
 * it writes dS = dL/dS (fp32 or bf16, ds_dtype = ISP_DTYPE_*), from which dQ = dS.K and dK = dS^T.Q
 * are plain batched GEMMs (cuBLAS).  All tensors contiguous and 16 B aligned; T2max % 4 == 0,
 * T2max <= ISP_LOGLIK_MAX_T2.  One pass: every input is read once, dS is written once.
 */
int    isp_loglik_backward_ds(const float* S, const float* attn_soft, const float* g_logits, const float* g_soft,
                              int B, int T1max, int T2max, float scale, int attention_prior,
                              void* dS, int ds_dtype, void* stream);

/* The same gradient without the scores: from the forward's own attn_logits (alignment.py:196 read backwards --
 * softmax_all(scale * S) = exp(attn_logits) / (prior + 1e-6), the prior in closed form from the row sums isp_loglik_forward
 * left in its workspace -- and attn_soft recomputed as the masked softmax of attn_logits).  No score GEMM in front of it, no
 * attn_soft read: 8-12 B/cell in, one launch less.  prior_rowsum (B, T1max) fp32 (may be NULL without attention_prior);
 * text_len, mel_len as in the forward call; the other arguments as isp_loglik_backward_ds.  Only after the FUSED forward kernel
 * (isp_loglik_forward / isp_align_forward): isp_loglik_rows evaluates the prior with other instructions. */
int    isp_loglik_backward_from_logits(const float* attn_logits, const float* g_logits, const float* g_soft, const float* prior_rowsum,
                                       const int64_t* text_len, const int64_t* mel_len, int B, int T1max, int T2max, float scale,
                                       int attention_prior, void* dS, int ds_dtype, void* stream);

/* ------------------------------------------------------------------------------------------
 * Batched, ragged GEMM on the tensor cores (tcgen05.mma, TMA operands, accumulator in TMEM):
 *     C[b] = act(alpha * A[b] . B[b]),   A[b] (M x K), B[b] (K x N), C[b] (M x N, row-major, ldc), fp32 accumulate.
 * The contractions next to the hot path are all instances of it:
 *   scores / dQ / dK of the log-likelihood's backward   tts/models/acoustic/modules/alignment.py:189 (differentiated)
 *   the soft-alignment length regulator and its grads   tts/models/acoustic/modules/temporal_adaptor.py:417-419
 *   the Conv1d layers of the projection stacks          alignment.py:40-83,118-154 (implicit GEMM over `taps` row shifts)
 * Operand storage:  a_mn_major = 0: A[b][m][k] at a + b*a_batch + m*lda + k   (the contraction index contiguous)
 *                   a_mn_major = 1: A[b][m][k] at a + b*a_batch + k*lda + m   (the row index contiguous: a transposed view)
 *                   b_mn_major = 0: B[b][k][n] at b + b*b_batch + n*ldb + k;  b_mn_major = 1: at b + b*b_batch + k*ldb + n
 * a_batch / b_batch = 0 shares the operand between batch entries.  dtype_ab: ISP_DTYPE_BF16, ISP_DTYPE_F16, or ISP_DTYPE_F32
 * (fp32 in memory, TF32 products).  dtype_c: any of the three.  Every pointer and every stride must be a multiple
 * of 16 B.
 * Ragged batches (all optional, DEVICE int64 (batch,)): rows >= m_len[b] and columns >= n_len[b] of C[b] are written as
 * zeros (unless skip_padding, which leaves whole padded tiles untouched); the contraction stops at k_len[b] rounded up to
 * 128 B of K -- the caller guarantees that one operand is zero from k_len[b] on.
 * Convolution: taps > 1 makes it  C[b][m][n] = act(alpha * sum_t sum_k A[b][m + t + tap_shift][k] * B_t[k][n])  with
 * B_t at b + t*b_tap_stride (b_batch must be 0, A K-major); rows outside [0, M) read as zeros.
 * act: 0 none, 1 ReLU, 2 GELU (erf).  col_stats (optional, fp32 (batch, 4*ceil(M/128), N, 2)): per 32-row slab the
 * column sums of C and of C^2 after masking (the masked-instance-norm statistics, tts/modules/normalization.py:160-208);
 * slabs of padded tiles are not written (pre-zero the buffer).
 * bn: tile width up to 256, a multiple of 16 (of 128 B worth of B's elements when B is MN-major); 0 = chosen from N. */
typedef struct isp_gemm_desc {
    const void* a; const void* b; void* c;
    const int64_t* m_len; const int64_t* n_len; const int64_t* k_len;
    float* col_stats;
    int64_t lda, ldb, ldc, a_batch, b_batch, c_batch, b_tap_stride;
    int32_t batch, M, N, K;
    int32_t dtype_ab, dtype_c, a_mn_major, b_mn_major;
    int32_t taps, tap_shift, act, bn, skip_padding;
    float alpha;
    int32_t stages;   /* tuning: depth of the operand ring (2..6), 0 = chosen from the shape */
    void* trace;      /* debug only, normally NULL: device int64 (CTAs, 8), SM-clock stamps of each CTA's phases (tools/gemm_trace.py) */
} isp_gemm_desc;
int    isp_gemm_batched(const isp_gemm_desc* desc, void* stream);

/* Element-wise companions of the convolution GEMMs of the projection stacks (alignment.py:40-83,118-154,176-187).
 * isp_prep_channels_last: x (B, C, T) if channels_first else (B, T, C), fp32 or bf16, contiguous -> out (B, T, Cp) in
 *   out_dtype, zero at t >= len[b] (ConvBlock1D masks its input, alignment.py:75-76; len optional) and in the Cp - C padding
 *   channels: the K-major activation the implicit-GEMM convolution loads.
 * isp_instance_norm_apply: masked instance norm (tts/modules/normalization.py:186-206) of a channels-last activation
 *   y (B, T, C; row stride ld_in) from the column statistics isp_gemm_batched left in `stats` (B, parts, C, 2):
 *   mean = sum / len, var = sumsq / len - mean^2 (biased), out = ((y - mean) / sqrt(var + eps) * weight + bias) for
 *   t < len[b], 0 past it.  weight / bias (C) fp32 or NULL.  ws: B * C * 8 bytes, 8 B aligned (the per-channel scale and shift).
 *   In place (out == y) is allowed; rows past len[b] are then left as they are (the GEMM wrote zeros there). */
int    isp_prep_channels_last(const void* x, int in_dtype, int channels_first, const int64_t* len, void* out, int out_dtype,
                              int B, int C, int T, int Cp, void* stream);
int    isp_instance_norm_apply(const void* y, int dtype, const float* stats, int parts, const float* weight, const float* bias,
                               const int64_t* len, void* out, int B, int T, int C, int64_t ld_in, int64_t ld_out, float eps, void* ws, void* stream);

/* Tuning knobs for benchmarks/tests (process-wide, not part of the drop-in contract).
 *   "mas.ring_rows"      rows of logits kept in flight per strip of 128 tokens, 0 = heuristic
 *   "mas.slots"          utterances per CTA (1 | 2), 0 = heuristic
 *   "mas.bits_global"    1: backpointer bits always in the workspace
 *   "mas.no_tma"         1: 4 B async copies instead of tiled TMA boxes (the path taken when the
 *                        logits are not 16 B aligned or T2max is not a multiple of 4)
 *   "mas.dbg"            profiling switches, see isp_mas.cu
 *   "mas.impl"           0: the kernel the shape calls for; 1: isp_mas.cu; 2: isp_mas2.cu or fail; 3: isp_mas_wide.cu;
 *                        4: isp_mas_cluster.cu or fail (tests force each kernel onto every shape it covers)
 *   "masc.trace"         1: the cluster kernel leaves phase timestamps in the workspace (tools/masc_probe.py)
 * Returns the previous value, or ISP_ERR_INVALID for an unknown key. */
int isp_set_option(const char* key, int value);

#ifdef __cplusplus
}
#endif
#endif /* ISP_TTS_B200_H */
