// group_gather.cu -- index gathers of channel-major features (+ their grads), sm_100a.
//
// Replaces group_points[_grad]_kernel_fast (reference
// network/models/pointnet_lib/src/group_points_gpu.cu:47-66, :8-25) and
// gather_points[_grad]_kernel_fast (src/sampling_gpu.cu:8-24, :46-63).
// gather is group with nsample == 1, so one kernel pair serves both.
//
// The reference launches one thread per (b, c, s, k) and reloads the index for
// every channel.  Here a thread owns one (s,k) slot, loads its index once and
// walks kChunk channels: per output element one L1/L2 gather and one coalesced
// store.  Bound: B*C*S*K*4 output bytes (HBM write roofline).
#include "pn2_common.cuh"

namespace pn2 {
namespace {

constexpr int kThreads = 256;
constexpr int kChunk = 16;

__global__ void __launch_bounds__(kThreads)
group_points_kernel(int c, int n, int slots, const float* __restrict__ points, const int* __restrict__ idx,
                    float* __restrict__ out) {
    pdl_enter();  // programmatic dependent launch (pn2_common.cuh): first statement, before any memory access
    const int b = blockIdx.z;
    const int c0 = blockIdx.y * kChunk;
    const int cend = min(c, c0 + kChunk);
    const int s = blockIdx.x * kThreads + threadIdx.x;
    if (s >= slots) return;
    const int src = idx[(size_t)b * slots + s];
    for (int cc = c0; cc < cend; ++cc)
        out[((size_t)b * c + cc) * slots + s] = __ldg(points + ((size_t)b * c + cc) * n + src);
}

__global__ void __launch_bounds__(kThreads)
group_points_grad_kernel(int c, int n, int slots, const float* __restrict__ grad_out,
                         const int* __restrict__ idx, float* __restrict__ grad_points) {
    const int b = blockIdx.z;
    const int c0 = blockIdx.y * kChunk;
    const int cend = min(c, c0 + kChunk);
    const int s = blockIdx.x * kThreads + threadIdx.x;
    if (s >= slots) return;
    const int dst = idx[(size_t)b * slots + s];
    for (int cc = c0; cc < cend; ++cc)
        atomicAdd(grad_points + ((size_t)b * c + cc) * n + dst, grad_out[((size_t)b * c + cc) * slots + s]);
}

int launch_group(bool grad, int b, int c, int n, long long slots, const float* src, const int* idx, float* dst,
                 cudaStream_t stream, const char* who) {
    if (b < 0 || c < 0 || n < 0 || slots < 0) return fail_arg(who, "negative size");
    if (b == 0 || c == 0 || slots == 0) return 0;
    if (b > 65535 || (c + kChunk - 1) / kChunk > 65535 || slots > 0x7fffffffLL) return fail_arg(who, "size too large");
    if (!src || !idx || !dst) return fail_arg(who, "null pointer");
    dim3 grid((unsigned)((slots + kThreads - 1) / kThreads), (c + kChunk - 1) / kChunk, b);
    if (grad)
        group_points_grad_kernel<<<grid, kThreads, 0, stream>>>(c, n, (int)slots, src, idx, dst);
    else
        launch_k(group_points_kernel, dim3(grid), dim3(kThreads), 0, stream, c, n, (int)slots, src, idx, dst);
    PN2_CHECK_LAUNCH(who);
    return 0;
}

}  // namespace
}  // namespace pn2

extern "C" int pn2_group_points(int b, int c, int n, int npoints, int nsample, const float* points,
                                const int* idx, float* out, pn2_stream_t stream) {
    return pn2::launch_group(false, b, c, n, (long long)npoints * nsample, points, idx, out, (cudaStream_t)stream,
                             "pn2_group_points");
}
extern "C" int pn2_group_points_grad(int b, int c, int n, int npoints, int nsample, const float* grad_out,
                                     const int* idx, float* grad_points, pn2_stream_t stream) {
    return pn2::launch_group(true, b, c, n, (long long)npoints * nsample, grad_out, idx, grad_points,
                             (cudaStream_t)stream, "pn2_group_points_grad");
}
extern "C" int pn2_gather_points(int b, int c, int n, int npoints, const float* points, const int* idx,
                                 float* out, pn2_stream_t stream) {
    return pn2::launch_group(false, b, c, n, npoints, points, idx, out, (cudaStream_t)stream, "pn2_gather_points");
}
extern "C" int pn2_gather_points_grad(int b, int c, int n, int npoints, const float* grad_out, const int* idx,
                                      float* grad_points, pn2_stream_t stream) {
    return pn2::launch_group(true, b, c, n, npoints, grad_out, idx, grad_points, (cudaStream_t)stream,
                             "pn2_gather_points_grad");
}
