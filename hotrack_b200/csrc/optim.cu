// optim.cu -- flat-buffer Adam step for the data-parallel training path, sm_100a.
//
// The reference trains with torch.optim.Adam(lr=1e-4, weight_decay=1e-4)
// (network/trainer.py:66-73, configs/all_config/handtracknet_train_SimGrasp.yml:36-45), which
// launches a multi-tensor kernel group per parameter list.  Here all parameters, gradients and
// moments live in four flat fp32 buffers (hotrack_b200/flat.py), so one step is ONE
// float4-vectorised pass: 16 B/param read x4, 16 B/param written x3 -- HBM-bound, 28 B/param.
// Update rule = torch.optim.Adam (L2 weight decay folded into the gradient, bias-corrected):
//   g += wd*p;  m = b1*m + (1-b1)*g;  v = b2*v + (1-b2)*g*g;
//   p -= (lr / (1-b1^t)) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
// `grad_scale` multiplies the gradient first (1/world_size after a SUM all-reduce).
#include "pn2_common.cuh"

#include <cmath>

namespace pn2 {
namespace {

struct AdamArgs {
    float lr_over_bc1, inv_sqrt_bc2, beta1, beta2, eps, weight_decay, grad_scale;
    float lr;
    const int* step_dev;  // device-resident step counter (CUDA-graph replays): t = *step_dev + 1
};

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, const AdamArgs& a) {
    g = fmaf(a.weight_decay, p, g * a.grad_scale);
    m = fmaf(a.beta1, m, (1.f - a.beta1) * g);
    v = fmaf(a.beta2, v, (1.f - a.beta2) * g * g);
    const float denom = fmaf(sqrtf(v), a.inv_sqrt_bc2, a.eps);
    p -= a.lr_over_bc1 * (m / denom);
}

__global__ void adam_advance_kernel(int* step_dev) { pdl_enter(); *step_dev += 1; }

__global__ void __launch_bounds__(256)
adam_kernel(long long n, float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
            float* __restrict__ v, AdamArgs a) {
    pdl_enter();  // programmatic dependent launch (pn2_common.cuh): first statement, before any memory access
    if (a.step_dev) {  // bias corrections from the device-side counter (same for every thread)
        const float t = (float)(*a.step_dev + 1);
        a.lr_over_bc1 = a.lr / (1.f - powf(a.beta1, t));
        a.inv_sqrt_bc2 = rsqrtf(1.f - powf(a.beta2, t));
    }
    const long long n4 = n >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 pp = reinterpret_cast<float4*>(p)[i];
        const float4 gg = reinterpret_cast<const float4*>(g)[i];
        float4 mm = reinterpret_cast<float4*>(m)[i];
        float4 vv = reinterpret_cast<float4*>(v)[i];
        adam_one(pp.x, gg.x, mm.x, vv.x, a);
        adam_one(pp.y, gg.y, mm.y, vv.y, a);
        adam_one(pp.z, gg.z, mm.z, vv.z, a);
        adam_one(pp.w, gg.w, mm.w, vv.w, a);
        reinterpret_cast<float4*>(p)[i] = pp;
        reinterpret_cast<float4*>(m)[i] = mm;
        reinterpret_cast<float4*>(v)[i] = vv;
    }
    // tail (n % 4) by the first threads of block 0
    const long long t = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) adam_one(p[t], g[t], m[t], v[t], a);
}

}  // namespace
}  // namespace pn2

extern "C" int pn2_adam_step(long long n, float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                             float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                             int* step_dev, float grad_scale, pn2_stream_t stream) {
    using namespace pn2;
    if (n < 0 || (!step_dev && step < 1)) return fail_arg("pn2_adam_step", "n < 0 or step < 1");
    if (n == 0) return 0;
    if (!params || !grads || !exp_avg || !exp_avg_sq) return fail_arg("pn2_adam_step", "null pointer");
    if ((reinterpret_cast<uintptr_t>(params) | reinterpret_cast<uintptr_t>(grads) |
         reinterpret_cast<uintptr_t>(exp_avg) | reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15)
        return fail_arg("pn2_adam_step", "buffers must be 16-byte aligned");
    AdamArgs a;
    const int t = step < 1 ? 1 : step;
    const double bc1 = 1.0 - std::pow((double)beta1, t), bc2 = 1.0 - std::pow((double)beta2, t);
    a.lr_over_bc1 = (float)(lr / bc1);
    a.inv_sqrt_bc2 = (float)(1.0 / std::sqrt(bc2));
    a.lr = lr;
    a.step_dev = step_dev;
    a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay; a.grad_scale = grad_scale;
    long long blocks = ((n >> 2) + 255) / 256;
    if (blocks < 1) blocks = 1;
    if (blocks > 148 * 8) blocks = 148 * 8;
    launch_k(adam_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, n, params, grads, exp_avg, exp_avg_sq, a);
    PN2_CHECK_LAUNCH("adam_kernel");
    if (step_dev) {
        launch_k(adam_advance_kernel, dim3(1), dim3(1), 0, (cudaStream_t)stream, step_dev);
        PN2_CHECK_LAUNCH("adam_advance_kernel");
    }
    return 0;
}
