// interpolate.cu -- three-point feature interpolation (forward + grad), sm_100a.
//
// Replaces three_interpolate_kernel_fast / three_interpolate_grad_kernel_fast
// (reference network/models/pointnet_lib/src/interpolate_gpu.cu:149-169, :192-214).
//
// Forward.  out[b,c,i] = w0*p[idx0] + w1*p[idx1] + w2*p[idx2], evaluated with the
// reference's rounding sequence fma(w2,p2, fma(w0,p0, rn(w1*p1))).  The reference
// launches one thread per (b,c,i) and re-reads idx/weight for every channel
// (C x 24 B per point); its gathers walk a channel row in global memory.  Here a
// thread owns kPtsPerThread points, keeps their idx/weight in registers, and a
// CTA walks a chunk of kChunk channels whose (B,C,M) rows are staged in shared
// memory once, so the three gathers are LDS and the only global traffic per
// output element is its own coalesced 4-byte store: the kernel is bound by the
// B*C*N*4 output bytes (HBM roofline).
//
// Grad.  The reference scatters with three fp32 atomicAdd per (b,c,i) into a
// zeroed (B,C,M) buffer.  sm_100a has no native shared-memory fp32 add
// (ATOMS.CAST.SPIN loop), so the scatter stays in L2 (REDG.ADD.F32) but idx and
// weight are read once per point instead of once per channel.  (The fused engine
// does not call this kernel: its FP stack scatters row-form gradients in fp_rows_bwd, mlp_rows.cu.)
#include "pn2_common.cuh"

namespace pn2 {
namespace {

constexpr int kThreads = 256;
constexpr int kPtsPerThread = 4;
constexpr int kChunk = 32;          // channels per CTA
constexpr int kSmemFloats = 8192;   // 32 KB of staged channel rows

__global__ void __launch_bounds__(kThreads)
three_interpolate_kernel(int c, int m, int n, const float* __restrict__ points, const int* __restrict__ idx,
                         const float* __restrict__ weight, float* __restrict__ out) {
    __shared__ float s_rows[kSmemFloats];
    const int b = blockIdx.z;
    const int c0 = blockIdx.y * kChunk;
    const int cend = min(c, c0 + kChunk);
    const int i0 = blockIdx.x * (kThreads * kPtsPerThread) + threadIdx.x;

    int ix[kPtsPerThread][3];
    float w[kPtsPerThread][3];
#pragma unroll
    for (int p = 0; p < kPtsPerThread; ++p) {
        const int i = i0 + p * kThreads;
        if (i < n) {
            const size_t o = ((size_t)b * n + i) * 3;
#pragma unroll
            for (int j = 0; j < 3; ++j) { ix[p][j] = idx[o + j]; w[p][j] = weight[o + j]; }
        } else {
#pragma unroll
            for (int j = 0; j < 3; ++j) { ix[p][j] = 0; w[p][j] = 0.f; }
        }
    }

    const int rows_per_pass = m <= kSmemFloats ? min(kChunk, kSmemFloats / max(m, 1)) : 0;
    if (rows_per_pass > 0) {
        for (int cb = c0; cb < cend; cb += rows_per_pass) {
            const int nrows = min(rows_per_pass, cend - cb);
            __syncthreads();
            const float* src = points + ((size_t)b * c + cb) * m;  // nrows contiguous rows of m floats
            for (int e = threadIdx.x; e < nrows * m; e += kThreads) s_rows[e] = src[e];
            __syncthreads();
            for (int r = 0; r < nrows; ++r) {
                const float* row = s_rows + r * m;
                float* orow = out + ((size_t)b * c + cb + r) * n;
#pragma unroll
                for (int p = 0; p < kPtsPerThread; ++p) {
                    const int i = i0 + p * kThreads;
                    if (i < n) {
                        const float t = __fmaf_rn(w[p][0], row[ix[p][0]], __fmul_rn(w[p][1], row[ix[p][1]]));
                        orow[i] = __fmaf_rn(w[p][2], row[ix[p][2]], t);
                    }
                }
            }
        }
    } else {  // rows too long for shared memory: gather through L1/L2
        for (int cc = c0; cc < cend; ++cc) {
            const float* row = points + ((size_t)b * c + cc) * m;
            float* orow = out + ((size_t)b * c + cc) * n;
#pragma unroll
            for (int p = 0; p < kPtsPerThread; ++p) {
                const int i = i0 + p * kThreads;
                if (i < n) {
                    const float t = __fmaf_rn(w[p][0], __ldg(row + ix[p][0]), __fmul_rn(w[p][1], __ldg(row + ix[p][1])));
                    orow[i] = __fmaf_rn(w[p][2], __ldg(row + ix[p][2]), t);
                }
            }
        }
    }
}

__global__ void __launch_bounds__(kThreads)
three_interpolate_grad_kernel(int c, int n, int m, const float* __restrict__ grad_out,
                              const int* __restrict__ idx, const float* __restrict__ weight,
                              float* __restrict__ grad_points) {
    const int b = blockIdx.z;
    const int c0 = blockIdx.y * kChunk;
    const int cend = min(c, c0 + kChunk);
    const int i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    const size_t o = ((size_t)b * n + i) * 3;
    const int i0 = idx[o], i1 = idx[o + 1], i2 = idx[o + 2];
    const float w0 = weight[o], w1 = weight[o + 1], w2 = weight[o + 2];
    for (int cc = c0; cc < cend; ++cc) {
        const float g = grad_out[((size_t)b * c + cc) * n + i];
        float* gp = grad_points + ((size_t)b * c + cc) * m;
        // products rounded to fp32 before the add, as interpolate_gpu.cu:211-213
        atomicAdd(gp + i0, __fmul_rn(g, w0));
        atomicAdd(gp + i1, __fmul_rn(g, w1));
        atomicAdd(gp + i2, __fmul_rn(g, w2));
    }
}

}  // namespace
}  // namespace pn2

extern "C" int pn2_three_interpolate(int b, int c, int m, int n, const float* points, const int* idx,
                                     const float* weight, float* out, pn2_stream_t stream) {
    using namespace pn2;
    if (b < 0 || c < 0 || m < 0 || n < 0) return fail_arg("pn2_three_interpolate", "negative size");
    if (b == 0 || c == 0 || n == 0) return 0;
    if (b > 65535 || (c + kChunk - 1) / kChunk > 65535) return fail_arg("pn2_three_interpolate", "b or c too large");
    if (!points || !idx || !weight || !out) return fail_arg("pn2_three_interpolate", "null pointer");
    dim3 grid((n + kThreads * kPtsPerThread - 1) / (kThreads * kPtsPerThread), (c + kChunk - 1) / kChunk, b);
    three_interpolate_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(c, m, n, points, idx, weight, out);
    PN2_CHECK_LAUNCH("three_interpolate_kernel");
    return 0;
}

extern "C" int pn2_three_interpolate_grad(int b, int c, int n, int m, const float* grad_out, const int* idx,
                                          const float* weight, float* grad_points, pn2_stream_t stream) {
    using namespace pn2;
    if (b < 0 || c < 0 || m < 0 || n < 0) return fail_arg("pn2_three_interpolate_grad", "negative size");
    if (b == 0 || c == 0 || n == 0) return 0;
    if (b > 65535 || (c + kChunk - 1) / kChunk > 65535) return fail_arg("pn2_three_interpolate_grad", "b or c too large");
    if (!grad_out || !idx || !weight || !grad_points) return fail_arg("pn2_three_interpolate_grad", "null pointer");
    dim3 grid((n + kThreads - 1) / kThreads, (c + kChunk - 1) / kChunk, b);
    three_interpolate_grad_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(c, n, m, grad_out, idx, weight,
                                                                               grad_points);
    PN2_CHECK_LAUNCH("three_interpolate_grad_kernel");
    return 0;
}
