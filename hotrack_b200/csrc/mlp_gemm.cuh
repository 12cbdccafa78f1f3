// mlp_gemm.cuh -- argument block shared by the two implementations of the row-matrix GEMM
// (mlp_gemm_tc.cu: tcgen05 / TMEM, the production path; mlp_gemm.cu: warp-level mma.sync, kept as the
// cross-check the tests and tools/dev/gemm_tc_check.cu compare against, PN2_GEMM_IMPL=mma).
#pragma once
#include "mma_common.cuh"

namespace pn2 {

enum { A_PLAIN = 0, A_AFFINE = 1, A_BNBWD = 2 };

struct GemmArgs {
    long long rows;
    int kdim, n;
    const uint16_t* a0; int a0_ld;   // forward: x (fp16);  BNBWD: dz (bf16)
    const uint16_t* a1; int a1_ld;   // BNBWD: y (fp16)
    const float *c0, *c1, *c2;
    const uint16_t* b;               // forward: w (fp16);  BNBWD: W_A = cA o W^T (bf16, [k_out][n_red])
    const uint16_t* b1;              // BNBWD: W_B = yscale^-1 * cB o W^T (fp16, scaled into fp16 range by a power of two)
    const float* yscale;             // BNBWD: device scalar, the factor that undoes W_B's scaling
    const float* center;  // [n] subtracted from the accumulators before 16-bit rounding (nullable)
    uint16_t* out; int out_ld;       // forward: y (fp16);  BNBWD: dz_prev (bf16)
    uint16_t* out_lo;                // forward, two-plane operands (a1 = x_lo, b1 = w_lo): y's lo plane, ld = out_ld (nullable)
    float* sums;
    const uint16_t* yp; int yp_ld;   // MASK: previous layer's y (fp16)
    const float *p_scale, *p_shift, *p_mean, *p_rstd;
    int nst, bres;                   // set by the mma.sync launcher
    // forward + training: BatchNorm finalisation by the last CTA to finish (fin_counter != null; zeroed by the caller)
    unsigned int* fin_counter;
    const float *fin_gamma, *fin_beta, *fin_bias, *fin_center;
    float fin_momentum, fin_eps, fin_unbias, fin_inv_rows;
    float *fin_running_mean, *fin_running_var;
    long long* fin_nbt;
    float *fin_scale, *fin_shift, *fin_mean, *fin_rstd;
    float* fin_next_center;          // nullable: [n] <- batch mean of the un-centred output (next step's centre)
    // forward, last layer of a pooled stack: max over the pool_k rows of each group taken in the EPILOGUE (no pool_fwd
    // launch, no second pass over y).  BatchNorm's scale has the sign of gamma, known before the statistics are:
    // max_k relu(s*y_k + t) = relu(s * (gamma >= 0 ? max_k y_k : min_k y_k) + t), so one extreme per (group, column)
    // suffices.  pool_k in {16, 32, 64, 128} (divides the 128-row tile); pool_val / pool_arg [rows / pool_k][n].
    int pool_k;
    const float* pool_gamma;         // [n] sign source
    float* pool_val;                 // the selected extreme of (acc - center)
    int* pool_arg;                   // its row inside the group (first one in row order)
};

// BatchNorm finalisation of one channel from the column sums (shared by bn_finalize_kernel and the GEMM tail)
__device__ __forceinline__ void bn_finalize_channel(int c, int n, float s1, float s2, float inv_rows, float unbias,
                                                    const float* gamma, const float* beta, const float* bias,
                                                    const float* center, float momentum, float eps, float* running_mean,
                                                    float* running_var, float* scale, float* shift, float* mean_out,
                                                    float* rstd_out) {
    const float mean = s1 * inv_rows;
    const float var = fmaxf(fmaf(-mean, mean, s2 * inv_rows), 0.f);
    const float rstd = rsqrtf(var + eps);
    const float sc = gamma[c] * rstd;
    scale[c] = sc;
    shift[c] = fmaf(-mean, sc, beta[c]);
    mean_out[c] = mean;
    rstd_out[c] = rstd;
    if (running_mean) {
        // the conv bias and the centring constant shift the batch mean BatchNorm sees; the GEMM omits
        // both (BatchNorm cancels them exactly)
        const float m = mean + (bias ? bias[c] : 0.f) + (center ? center[c] : 0.f);
        running_mean[c] = fmaf(momentum, m - running_mean[c], running_mean[c]);
        running_var[c] = fmaf(momentum, var * unbias - running_var[c], running_var[c]);
    }
}

struct WgradArgs {
    long long rows;
    int n, kp, k_true;
    const bf16* dz; int dz_ld;
    const act_t* y; int y_ld;
    const float *cA, *cB, *cC;
    const act_t* x; int x_ld;
    const float *in_scale, *in_shift;
    float* dw; int dw_ld;
};

// tcgen05 weight-gradient kernel (mlp_wgrad_tc.cu); supported when the whole output-channel range fits TMEM (n <= 512).
// Opt-in (env PN2_WGRAD_IMPL=tc): the row reduction makes both operands MN-major, and measured on B200 the tensor core
// is fed MN-major 16-bit tiles at a fraction of the K-major rate (0.13-0.3 us per M=128,K=16 instruction), so the
// warp-level kernel with ldmatrix.trans fragments (mlp_gemm.cu) is the faster one on this network's shapes.
bool wgrad_use_tc();
bool wgrad_tc_supported(const WgradArgs& a);
int launch_wgrad_tc(const WgradArgs& a, cudaStream_t stream);

// tcgen05 implementation (mlp_gemm_tc.cu).  amode: A_PLAIN / A_AFFINE / A_BNBWD; mask: ReLU-mask epilogue.
int launch_gemm_tc(const GemmArgs& a, int amode, bool mask, cudaStream_t stream);
// which implementation pn2_mlp_gemm_fwd / pn2_mlp_gemm_dgrad dispatch to (env PN2_GEMM_IMPL = tc | mma, default tc)
bool gemm_use_tc();

}  // namespace pn2
