/*
 * pn2b200_mlp.h -- C ABI of the fused grouped-MLP kernels of libpn2b200.so.
 *
 * These entry points have no C counterpart in the reference: there the per-point MLP of the SA / FP
 * modules is torch.nn (Conv2d/Conv1d + BatchNorm + ReLU + max: network/models/pointnet_utils.py:
 * 389-403, 443-463, 484-512, 536-590; backbones.py:131-132).  They are what a Python binding of the
 * fused "engine" calls (hotrack_b200/fused.py); conventions as in pn2b200.h: raw device pointers, int
 * sizes, caller-owned buffers, asynchronous on `stream`, int status (0 = ok, see pn2_last_error()).
 *
 * ROW MATRICES: activations are 16-bit matrices X[rows][ld] -- fp16 for every forward quantity (inputs, pre-BatchNorm
 * layer outputs, weights), bf16 for gradients --, channels contiguous (ld % 8 == 0, 16-byte aligned base),
 * rows = B*S*K grouped neighbours or B*N points.  A "row source" is such a
 * matrix plus optional per-channel fp32 scale/shift: when given, the consumer reads
 * relu(x*scale + shift) -- i.e. the producer's BatchNorm + ReLU, applied on the fly.
 */
#ifndef PN2B200_MLP_H_
#define PN2B200_MLP_H_
#include "pn2b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* (B,C,N) fp32 channel-major -> rows [B*N][ld] fp16, columns >= C zero-filled.  sub_sums != NULL: the value
 * sub_sums[c]*sub_scale (a channel mean) is subtracted before rounding ("centred rows"). */
int pn2_to_rows(int b, int c, int n, const float* src, const float* sub_sums, float sub_scale, void* dst, int ld,
                pn2_stream_t stream);

/* Grouped rows of one SA scale (pointnet_utils.py:389-396, 570-575; group-all :170-186).
 * Row (b,s,k), point j = idx[b,s,k] (idx NULL: j = k, k == n):
 *   xyz_first == 0: [ feat[b,j,0..feat_c) | xyz[b,:,j] - new_xyz[b,:,s] | cen[b,s,0..cen_c) | 0.. ]
 *   xyz_first == 1: [ xyz - centre | feat | 0.. ]          (new_xyz NULL: centre = 0)
 * feat / cen: row sources over (B*N) / (B*S) rows, NULL = absent. */
int pn2_sa_build_rows(int b, int n, int s, int k, const float* xyz, const float* new_xyz, const int* idx,
                      const void* feat, int feat_c, int feat_ld, const float* feat_scale, const float* feat_shift,
                      const void* cen, int cen_c, int cen_ld, const float* cen_scale, const float* cen_shift,
                      int xyz_first, void* out, int out_ld, pn2_stream_t stream);

/* Rows of one FP layer (pointnet_utils.py:443-456): [ skip[b,i,:] | sum_j w_j * coarse[b, idx[b,i,j], :] | 0.. ]
 * with w from the three_nn squared distances (w_j = 1/(sqrt(d2_j)+1e-8), normalised); s == 1 broadcasts. */
int pn2_fp_build_rows(int b, int n, int s, const void* skip, int skip_c, int skip_ld, const float* skip_scale,
                      const float* skip_shift, const void* coarse, int coarse_c, int coarse_ld,
                      const float* coarse_scale, const float* coarse_shift, const int* idx, const float* dist2,
                      void* out, int out_ld, pn2_stream_t stream);

/* center[n] = w[n][:] . mean_j act(x[row_j][:]) over <= 16 rows spread over the matrix: a cheap estimate of the
 * per-channel mean of the GEMM output, used to centre it before fp16 rounding (kdim <= 1024).  off0 / off1 (nullable): up to
 * two column segments [start, start + count) whose input rows were themselves stored centred by the per-channel constants
 * off[i] * off_scale (channel sums and 1 / count, as pn2_to_rows takes them); center_true = center + w . offset is what the
 * true (uncentred) output differs from the stored one by -- bn_finalize / bn_eval_affine take center_true. */
int pn2_mlp_center(long long rows, int kdim, int n, const void* x, int x_ld, const float* in_scale,
                   const float* in_shift, const void* w, const float* off0, float off0_scale, int off0_start,
                   int off0_count, const float* off1, float off1_scale, int off1_start, int off1_count, float* center,
                   float* center_true, pn2_stream_t stream);

/* y[rows][n] = act(x)[rows][kdim] * w[n][kdim]^T - center[n]  (fp16 in, fp32 accumulate, fp16 out), act =
 * relu(x*scale+shift) when in_scale != NULL, center NULL = 0.  stats != NULL: stats[0..n) += column sums of y,
 * stats[n..2n) += sums of y^2 (zeroed by the caller) -- the BatchNorm batch statistics (shift-invariant, so the
 * centring only has to be undone in the running mean / eval shift).  kdim % 32 == 0, n % 8 == 0. */
int pn2_mlp_gemm_fwd(long long rows, int kdim, int n, const void* x, int x_ld, const float* in_scale,
                     const float* in_shift, const void* w, const float* center, void* y, int y_ld, float* stats,
                     pn2_stream_t stream);

/* Training-mode BatchNorm constants from the statistics: scale = gamma*rstd, shift = beta - mean*scale, plus
 * the running-statistics update of nn.BatchNorm (momentum, unbiased variance; conv_bias re-added to the mean
 * because the GEMM omits
 * what BatchNorm cancels; likewise the centring constant).  running_* / num_batches_tracked may be NULL. */
int pn2_bn_finalize(int n, long long rows, const float* sums, const float* gamma, const float* beta,
                    const float* conv_bias, const float* center, float momentum, float eps, float* running_mean, float* running_var,
                    long long* num_batches_tracked, float* scale, float* shift, float* mean, float* rstd,
                    pn2_stream_t stream);
/* Eval-mode constants from the running statistics (conv bias folded in). */
int pn2_bn_eval_affine(int n, const float* gamma, const float* beta, const float* conv_bias, const float* center,
                       const float* running_mean, const float* running_var, float eps, float* scale, float* shift,
                       pn2_stream_t stream);

/* BatchNorm+ReLU of the last layer and max over the k rows of each group: out_cm (B,C,S) fp32, optional
 * per-channel sums of the output (chan_sums[C], zeroed by the caller) and arg-max (B,S,C).
 * k == 1: plain BN+ReLU, rows -> channel-major (FP layers, head). */
int pn2_pool_fwd(int b, int s, int k, int c, const void* y, int y_ld, const float* scale, const float* shift,
                 float* out_cm, float* chan_sums, int* argmax, pn2_stream_t stream);

/* Backward of pool_fwd: dz[rows][c] = dout at the arg-max row where the ReLU is active, else 0; sums[0..c) +=
 * sum(dz), sums[c..2c) += sum(dz * xhat)  (zeroed by the caller).  extra_rows [B*S][c] fp32 (nullable) is a second
 * output-gradient term in row form (what fused consumers of this output deliver, see pn2_sa_rows_bwd /
 * pn2_fp_rows_bwd); dout_cm may then be NULL.  extra_rows16 (bf16 rows, a dense consumer's input gradient): k == 1
 * only. */
int pn2_pool_bwd(int b, int s, int k, int c, const float* dout_cm, const float* extra_rows, const void* extra_rows16,
                 int extra16_ld, const void* y, int y_ld,
                 const float* scale,
                 const float* shift, const float* mean, const float* rstd, const int* argmax, void* dz, int dz_ld,
                 float* sums, pn2_stream_t stream);

/* pn2_mlp_gemm_fwd followed by pn2_bn_finalize in ONE launch: the last CTA to finish reads the complete column sums and
 * writes scale / shift / mean / rstd and the running statistics.  counter: one zeroed unsigned int (consumed).
 * next_center (nullable, may alias center / center_true): [n] <- center + batch mean of y, i.e. the batch mean of the
 * un-centred output -- the centring constant a training loop passes as `center` at the next step instead of calling
 * pn2_mlp_center again. */
int pn2_mlp_gemm_fwd_bn(long long rows, int kdim, int n, const void* x, int x_ld, const float* in_scale,
                        const float* in_shift, const void* w, const float* center, void* y, int y_ld, float* stats,
                        unsigned int* counter, const float* gamma, const float* beta, const float* conv_bias,
                        const float* center_true, float momentum, float eps, float* running_mean, float* running_var,
                        long long* num_batches_tracked, float* scale, float* shift, float* mean, float* rstd,
                        float* next_center, pn2_stream_t stream);

/* BatchNorm-backward per-channel coefficients: dY = cA*dz + cB*y + cC; dgamma = sums[c..2c), dbeta = sums[0..c)
 * (accumulate != 0: added to the existing contents -- the parameters' .grad buffers).
 * w != NULL (fp32 conv weight [n][k_true]): also the weights with the coefficients folded in, for pn2_mlp_gemm_dgrad:
 * wa[kp][n] = bf16(S*cA[n]*w[n][k]), wb[kp][n] = fp16(S*cB[n]*w[n][k]) with S a power of two that brings cB into fp16 range
 * (rows k >= k_true zero), negbias[kp] = -sum_n cC[n]*w[n][k], *wb_unscale = 1/S. */
int pn2_bn_bwd_coefs(int n, long long rows, const float* sums, const float* gamma, const float* mean,
                     const float* rstd, float* cA, float* cB, float* cC, float* dgamma, float* dbeta, int accumulate,
                     const float* w, int k_true, int kp, void* wa, void* wb, float* negbias, float* wb_unscale,
                     pn2_stream_t stream);

/* dz_prev[rows][k_out] = dY[rows][n_red] * W[n_red][k_out]  with dY = cA*dz + cB*y + cC, evaluated as
 * *wb_unscale * (dz * wa^T + y * wb^T) - negbias  (wa, wb, negbias, wb_unscale from pn2_bn_bwd_coefs: the big operands
 * are multiplied as stored, bf16 x bf16 and fp16 x fp16, into one fp32 TMEM accumulator).
 * y_prev != NULL: the result is masked by the previous layer's ReLU (y_prev*prev_scale+prev_shift > 0) and sums_prev
 * receives the two BatchNorm-backward sums of the previous layer.  y_prev == NULL: plain input gradient. */
int pn2_mlp_gemm_dgrad(long long rows, int n_red, int k_out, const void* dz, int dz_ld, const void* y, int y_ld,
                       const void* wa, const void* wb, const float* negbias, const float* wb_unscale, const void* y_prev,
                       int y_prev_ld, const float* prev_scale, const float* prev_shift, const float* prev_mean,
                       const float* prev_rstd, void* dz_prev, int dz_prev_ld, float* sums_prev, pn2_stream_t stream);

/* dw[n][k_true] += sum_rows dY[r][n] * act(x)[r][k]  (fp32 atomics into the caller's zeroed buffer). */
int pn2_mlp_gemm_wgrad(long long rows, int n, int kp, int k_true, const void* dz, int dz_ld, const void* y, int y_ld,
                       const float* cA, const float* cB, const float* cC, const void* x, int x_ld,
                       const float* in_scale, const float* in_shift, float* dw, int dw_ld, pn2_stream_t stream);

/* fp32 conv weight [n][k_true] -> fp16 [n][kp] (zero padded) and, if wt_bf16 != NULL, its bf16 transpose [kp][n]. */
int pn2_mlp_prep_weights(int n, int k_true, int kp, const float* w, void* w_f16, void* wt_bf16, pn2_stream_t stream);

/* pn2_mlp_prep_weights (fp16 copy only) for n_layers layers in ONE launch.  descs: device array of n_layers records
 * { const float* w; void* w_f16; int n; int k_true; int kp; int two; } (32 bytes each). */
int pn2_mlp_prep_weights_multi(int n_layers, const void* descs, pn2_stream_t stream);

/* Gradient of sa_build_rows' output rows scattered to the feature tensors (fp32, zeroed by the caller, atomics):
 * dfeat_cm (B,feat_c,N) channel-major -- or, feat_rows_major != 0, rows [B*N][feat_c] (coalesced; the form
 * pn2_pool_bwd's extra_rows takes) --, dcen_cm (B,cen_c,S); either may be NULL. */
int pn2_sa_rows_bwd(int b, int n, int s, int k, const int* idx, const void* dx, int dx_ld, int feat_c,
                    float* dfeat_cm, int feat_rows_major, int cen_c, float* dcen_cm, int xyz_first,
                    pn2_stream_t stream);

/* Gradient of fp_build_rows' output rows: dskip (B,skip_c,N) channel-major, plain stores -- or, skip_rows_major != 0,
 * fp32 rows [B*N][skip_c] zeroed by the caller and ACCUMULATED into (the form pn2_pool_bwd's extra_rows takes) --;
 * dcoarse_rows (B*S, coarse_c) fp32 rows, zeroed by the caller, atomics; either may be NULL. */
int pn2_fp_rows_bwd(int b, int n, int s, const int* idx, const float* dist2, const void* dx, int dx_ld, int skip_c,
                    float* dskip, int skip_rows_major, int coarse_c, float* dcoarse_rows, pn2_stream_t stream);

/* ---- TWO-PLANE ("x2") forward rows -------------------------------------------------------------------------------
 * A row matrix may come as a PAIR of fp16 planes (hi, lo) of the same shape and leading dimension with
 * value = hi + lo (22 significand bits; lo = fp16(v - fp16(v))).  The forward GEMM then evaluates
 * x_hi*w_hi + x_lo*w_hi + x_hi*w_lo into its fp32 accumulator and writes y as such a pair.  The engine uses this
 * for the stacks in front of the backbone's ill-conditioned spot (SA1-SA3 and FP3's first layer: every rounding there
 * is amplified ~40x by the end of the network), where 11-bit fp16 rows miss the 1e-2 parity bar; the backward
 * kernels read the hi planes only.  Each _x2 entry point is its one-plane namesake with nullable lo pointers added
 * (all lo pointers NULL = identical behaviour). */
int pn2_to_rows_x2(int b, int c, int n, const float* src, const float* sub_sums, float sub_scale, void* dst, void* dst_lo,
                   int ld, pn2_stream_t stream);
int pn2_sa_build_rows_x2(int b, int n, int s, int k, const float* xyz, const float* new_xyz, const int* idx,
                         const void* feat, const void* feat_lo, int feat_c, int feat_ld, const float* feat_scale,
                         const float* feat_shift, const void* cen, const void* cen_lo, int cen_c, int cen_ld,
                         const float* cen_scale, const float* cen_shift, int xyz_first, void* out, void* out_lo,
                         int out_ld, pn2_stream_t stream);
int pn2_fp_build_rows_x2(int b, int n, int s, const void* skip, const void* skip_lo, int skip_c, int skip_ld,
                         const float* skip_scale, const float* skip_shift, const void* coarse, const void* coarse_lo,
                         int coarse_c, int coarse_ld, const float* coarse_scale, const float* coarse_shift,
                         const int* idx, const float* dist2, void* out, void* out_lo, int out_ld, pn2_stream_t stream);
/* x_lo and w_lo come together (both or neither); y_lo nullable (the last layer of a stack whose consumer is the
 * pooling kernel still wants it: the max is taken over hi + lo).  in_scale != NULL with two planes: kdim <= 256. */
int pn2_mlp_gemm_fwd_x2(long long rows, int kdim, int n, const void* x, const void* x_lo, int x_ld,
                        const float* in_scale, const float* in_shift, const void* w, const void* w_lo,
                        const float* center, void* y, void* y_lo, int y_ld, float* stats, pn2_stream_t stream);
int pn2_mlp_gemm_fwd_bn_x2(long long rows, int kdim, int n, const void* x, const void* x_lo, int x_ld,
                           const float* in_scale, const float* in_shift, const void* w, const void* w_lo,
                           const float* center, void* y, void* y_lo, int y_ld, float* stats, unsigned int* counter,
                           const float* gamma, const float* beta, const float* conv_bias, const float* center_true,
                           float momentum, float eps, float* running_mean, float* running_var,
                           long long* num_batches_tracked, float* scale, float* shift, float* mean, float* rstd,
                           float* next_center, pn2_stream_t stream);
int pn2_pool_fwd_x2(int b, int s, int k, int c, const void* y, const void* y_lo, int y_ld, const float* scale,
                    const float* shift, float* out_cm, float* chan_sums, int* argmax, pn2_stream_t stream);
/* fp32 conv weight [n][k_true] -> fp16 planes w_hi, w_lo [n][kp] (zero padded).  pn2_mlp_prep_weights_multi: a
 * record whose last int ("two") is non-zero writes the lo plane right behind the hi plane (w_f16 + n*kp). */
int pn2_mlp_prep_weights_x2(int n, int k_true, int kp, const float* w, void* w_hi, void* w_lo, pn2_stream_t stream);

/* ---- MAX-POOL IN THE GEMM EPILOGUE ---------------------------------------------------------------------------------
 * Last layer of a pooled stack (SA scales: pool_k neighbours per group): pn2_mlp_gemm_fwd[_bn]_x2 that additionally takes,
 * per (group, column), the extreme over the group's pool_k rows of (accumulator - center) -- the maximum where gamma >= 0,
 * the minimum where gamma < 0, since BatchNorm's scale has gamma's sign and max_k relu(s*y_k + t) = relu(s*extreme + t) --
 * and the row that holds it (first in row order).  pool_k in {16, 32, 64, 128}, rows % pool_k == 0;
 * pool_val / pool_arg: [rows / pool_k][n].  No pn2_pool_fwd launch, no second pass over y.  (_pool: pool_gamma given
 * explicitly; _bn_pool: the BatchNorm gamma.) */
int pn2_mlp_gemm_fwd_pool(long long rows, int kdim, int n, const void* x, const void* x_lo, int x_ld,
                          const float* in_scale, const float* in_shift, const void* w, const void* w_lo,
                          const float* center, void* y, void* y_lo, int y_ld, float* stats, int pool_k,
                          const float* pool_gamma, float* pool_val, int* pool_arg, pn2_stream_t stream);
int pn2_mlp_gemm_fwd_bn_pool(long long rows, int kdim, int n, const void* x, const void* x_lo, int x_ld,
                             const float* in_scale, const float* in_shift, const void* w, const void* w_lo,
                             const float* center, void* y, void* y_lo, int y_ld, float* stats, unsigned int* counter,
                             const float* gamma, const float* beta, const float* conv_bias, const float* center_true,
                             float momentum, float eps, float* running_mean, float* running_var,
                             long long* num_batches_tracked, float* scale, float* shift, float* mean, float* rstd,
                             float* next_center, int pool_k, float* pool_val, int* pool_arg, pn2_stream_t stream);
/* BatchNorm + ReLU of the pooled extremes -> out_cm (B,C,S) fp32 and optional channel sums; y: the layer's stored (hi
 * plane) output, read at the selected rows for the ReLU decision the backward pass will repeat. */
int pn2_pool_finalize(int b, int s, int k, int c, const float* pool_val, const int* pool_arg, const void* y, int y_ld,
                      const float* scale, const float* shift, float* out_cm, float* chan_sums, pn2_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PN2B200_MLP_H_ */
