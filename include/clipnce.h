/*
 * clipnce.h -- C-ABI of the B200-native fused CLIP / InfoNCE hot path.
 *
 * The reference (SrikarK-code/clip-dplm) has no FFI of its own: the boundary of this path is the
 * tail of its model forwards and its loss functions (Python).  The entry points below are exactly
 * what a binding for that tail needs; each one names the reference lines it replaces (paths are
 * relative to the reference checkout).  `INTEGRATION.md` shows the ctypes stub a maintainer adds.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (e.g. the PyTorch caching allocator);
 *     the library never allocates or frees user-visible memory and never synchronises the host:
 *     all work is enqueued on `stream` (a cudaStream_t passed as void*; NULL = legacy default);
 *   - matrices are dense row-major; "hat" buffers hold L2-normalised rows in the compute type:
 *     CLIPNCE_BF16 (tensor-core path, sm_100a tcgen05) or CLIPNCE_F32 (exact check mode, CUDA cores);
 *   - functions return 0 on success or a negative CLIPNCE_E* code; clipnce_last_error() returns a
 *     thread-local message.  No C++ exception crosses this boundary;
 *   - re-entrant: no global mutable state besides a mutex-guarded cache of kernel attributes.
 *
 * Notation:  Ahat_i = A_i / max(|A_i|, 1e-12),  S = s * Ahat Bhat^T,
 *            r_i = logsumexp_j S_ij,  c_j = logsumexp_i S_ij.
 */
#ifndef CLIPNCE_H_
#define CLIPNCE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CLIPNCE_VERSION 100

/* element types */
#define CLIPNCE_BF16 0
#define CLIPNCE_F32  1

/* flags */
#define CLIPNCE_FLAG_FORCE_EXACT 1   /* use the CUDA-core exact (online-max) kernels even for bf16 */

/* error codes */
#define CLIPNCE_OK            0
#define CLIPNCE_EINVAL       -1      /* bad argument (null pointer, bad dtype, bad shape/alignment) */
#define CLIPNCE_EWORKSPACE   -2      /* workspace too small */
#define CLIPNCE_ECUDA        -3      /* CUDA runtime / driver error (message has the detail) */
#define CLIPNCE_EUNSUPPORTED -4      /* device is not sm_100 or shape outside every kernel's range */

int         clipnce_version(void);
const char* clipnce_last_error(void);

/* 1 if (dtype, d, scale, flags) is served by the tcgen05 tensor-core kernels, 0 if by the exact
 * CUDA-core kernels (dtype F32, d % 8 != 0, d > 768, or 2*scale > 86 where exp(S - s) leaves fp32). */
int clipnce_uses_tensor_cores(int dtype, int64_t d, float scale, int flags);

/* Bytes of scratch clipnce_forward / clipnce_backward need for these shapes (max of the two). */
int clipnce_workspace_bytes(int64_t n_rows, int64_t n_cols, int64_t d, int dtype, int flags, size_t* out);

/*
 * Row norms (and, optionally, the normalised rows).  Replaces F.normalize(x, dim=-1):
 *   old/clip.py:63-64,100-101  run1/full.py:47-48,73-74  old/clip_opt.py:93-94
 *   current/rna_clip_codes.ipynb:1948-1949  current/tf_clip_codes (1).ipynb:13146-13148
 *   tong/utils/losses.py:6-7
 * x      [n,d]  in_dtype (BF16 or F32)
 * rinv   [n]    f32   1 / max(|x_i|, 1e-12)
 * x_hat  [n,d]  hat_dtype or NULL: normalised rows (what the reference returns as "*_embeds").
 * The contraction kernels do NOT consume x_hat: they take the raw rows plus rinv and apply
 * rinv_x[i] * rinv_y[j] to the fp32 accumulator ("normalise fused into the GEMM"), so the
 * tensor cores see the caller's exact bf16 values and no second rounding enters the logits.
 */
int clipnce_normalize(const void* x, int in_dtype, int64_t n, int64_t d, float* rinv,
                      void* x_hat, int hat_dtype, void* stream);

/*
 * Operand staging for the contraction kernels: x_c = x converted to c_dtype ([n,d], or NULL when x
 * already has that type) and x_c_t = the same values transposed ([d,ld_t], ld_t >= n, ld_t % 8 == 0,
 * or NULL).  The tensor-core backward streams the transposed copy as its second MMA operand; it is
 * also what is built from the all-gathered embeddings (old/clip_opt.py:102-112, run1/full.py:77-84).
 */
int clipnce_stage_operand(const void* x, int in_dtype, int64_t n, int64_t d, void* x_c, void* x_c_t,
                          int64_t ld_t, int c_dtype, void* stream);

/*
 * Forward statistics of S_ij = s * rinv_x[i] rinv_y[j] <x_i, y_j> without materialising S.  Replaces
 *   torch.matmul(a, b.t()) * logit_scale            old/clip.py:67,104  run1/full.py:50,85
 *                                                   old/clip_opt.py:115-121  rna_clip_codes.ipynb:1951
 *                                                   tf_clip_codes (1).ipynb:13152-13154  tong/utils/losses.py:14
 *   F.cross_entropy(S, arange) / F.cross_entropy(S.t(), arange)   (the log-sum-exp halves of it)
 *                                                   rna_clip_codes.ipynb:1952-1953  old/clip_opt.py:148-149
 *                                                   run1/full.py:98-99,133  tong/utils/losses.py:17-19
 * x [n_rows,d], y [n_cols,d]   raw rows in `dtype`: the rows local to this rank / all (gathered) columns
 * rinv_x [n_rows], rinv_y [n_cols]   from clipnce_normalize
 * diag_offset   column of row 0's positive: the positive of local row i is column i + diag_offset
 *               (0 on one GPU, rank * n_rows under the row-sharded global batch)
 * scale         s = exp(logit_scale), already clamped by the caller (old/clip_opt.py:100)
 * row_lse [n_rows]  r_i, complete (every column is seen locally)
 * col_m, col_l [n_cols]  partial column statistics over the LOCAL rows: c_j = M_j + log(sum over
 *               ranks of col_l_j * exp(col_m_j - M_j)), M_j = max over ranks of col_m_j
 * diag [n_rows]     S_{i, i+diag_offset}
 */
int clipnce_forward(const void* x, const void* y, const float* rinv_x, const float* rinv_y,
                    int64_t n_rows, int64_t n_cols, int64_t d,
                    int64_t diag_offset, float scale, int dtype, int flags,
                    float* row_lse, float* col_m, float* col_l, float* diag,
                    void* workspace, size_t workspace_bytes, void* stream);

/*
 * One side of the backward.  Replaces autograd through cross_entropy + matmul (loss.backward():
 * rna_clip_codes.ipynb:2074, run1/full.py:134, old/clip_opt.py:167, old/ablation.py:16-17).
 * The logit tiles are recomputed; with
 *     G_ij = exp(S_ij + log_u_i) + exp(S_ij + log_v_j) - diag_w * [j == i + diag_offset]
 * it returns  dx_hat = grad_out * s * G Yhat  ([n_rows,d] f32; Yhat_j = rinv_y[j] y_j)  and
 *             *d_scale_sum += grad_out * sum_ij G_ij S_ij   (= dL/dlogit_scale when s = exp(logit_scale)).
 * Symmetric InfoNCE over a global batch N:
 *     log_u_i = -log(2N) - r_i,  log_v_j = -log(2N) - c_j,  diag_w = 1/N
 * one-directional (run1/full.py:133, tong/utils/losses.py:19): log_u_i = -log(N) - r_i, log_v = NULL.
 * Columns without positives (hard-negative cache, old/clip_opt.py:118-121) carry log_v_j = -inf.
 * Call it once as (A, B, B^T, rinv_a, rinv_b, log_u, log_v, +offset) for dAhat and once as
 * (B, A, A^T, rinv_b, rinv_a, log_v, log_u, -offset) for dBhat.
 * y_t [d,ld_t] (clipnce_stage_operand) is only read by the tensor-core path (NULL for the exact path).
 * log_v and d_scale_sum may be NULL.
 */
int clipnce_backward(const void* x, const void* y, const void* y_t, int64_t ld_t,
                     const float* rinv_x, const float* rinv_y,
                     int64_t n_rows, int64_t n_cols, int64_t d, int64_t diag_offset, float scale,
                     const float* log_u, const float* log_v, float diag_w, float grad_out,
                     int dtype, int flags, float* dx_hat, float* d_scale_sum,
                     void* workspace, size_t workspace_bytes, void* stream);

/* out_i = log_coef - lse_i  (lse = +inf -> -inf).  Builds log_u / log_v from row / column LSE. */
int clipnce_log_weights(const float* lse, int64_t n, float log_coef, float* out, void* stream);

/* c_j = M_j + log(l_j * exp(m_j - M_j)) helper for one GPU: col_lse_j = col_m_j + log(col_l_j). */
int clipnce_combine_lse(const float* m, const float* l, int64_t n, float* lse, void* stream);

/* Backward of the normalise: dx_i = rinv_i * (g_i - xhat_i (xhat_i . g_i)), xhat_i = x_i * rinv_i
 * (rows clamped at eps get dx_i = g_i * rinv_i, like clamp_min's sub-gradient).
 * x [n,d] in_dtype, dx_hat [n,d] f32, dx [n,d] out_dtype.  grad_scale: optional DEVICE scalar (f32)
 * multiplied into dx -- the upstream gradient of the loss, read on the device so that autograd's
 * grad_output never forces a host synchronisation; NULL = 1. */
int clipnce_normalize_backward(const void* x, int in_dtype, const float* rinv, const float* dx_hat,
                               const float* grad_scale, int64_t n, int64_t d, void* dx, int out_dtype,
                               void* stream);

/* loss = [ sum_i (row_lse_i - diag_i) + (symmetric ? sum_i (col_lse_{i+diag_offset} - diag_i) : 0) ]
 *        / (symmetric ? 2 n_global : n_global), accumulated into *loss (f32, caller zeroes it).
 * Mean reduction of F.cross_entropy (rna_clip_codes.ipynb:1953). Deterministic (single block). */
int clipnce_loss(const float* row_lse, const float* col_lse, const float* diag, int64_t n_rows,
                 int64_t diag_offset, int64_t n_global, int symmetric, float* loss, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CLIPNCE_H_ */
