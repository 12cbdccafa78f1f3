// ref_shim.cu -- extern "C" entry points onto the REFERENCE's own launchers.
//
// TEST INFRASTRUCTURE ONLY (see oracle/pn2_oracle.c header).  This file holds no
// algorithm: it forwards to the kernel launchers the reference defines in
// network/models/pointnet_lib/src/*_gpu.cu, which oracle/Makefile compiles,
// unmodified and where they lie under /root/reference, into
// oracle/_ref/libpn2_ref.so.  The prototypes come from the reference's own
// headers (-I on the nvcc line): ball_query_gpu.h:12, group_points_gpu.h:13,19,
// interpolate_gpu.h:13,19,26,33, sampling_gpu.h:13,20,26.
#include "ball_query_gpu.h"
#include "group_points_gpu.h"
#include "interpolate_gpu.h"
#include "sampling_gpu.h"

extern "C" {

void ref_furthest_point_sampling(int b, int n, int m, const float* dataset, float* temp, int* idxs, void* stream) {
    furthest_point_sampling_kernel_launcher(b, n, m, dataset, temp, idxs, (cudaStream_t)stream);
}
void ref_ball_query(int b, int n, int m, float radius, int nsample, const float* new_xyz, const float* xyz, int* idx, void* stream) {
    ball_query_kernel_launcher_fast(b, n, m, radius, nsample, new_xyz, xyz, idx, (cudaStream_t)stream);
}
void ref_knn(int b, int n, int m, int k, const float* unknown, const float* known, float* dist2, int* idx, void* stream) {
    knn_kernel_launcher_fast(b, n, m, k, unknown, known, dist2, idx, (cudaStream_t)stream);
}
void ref_three_nn(int b, int n, int m, const float* unknown, const float* known, float* dist2, int* idx, void* stream) {
    three_nn_kernel_launcher_fast(b, n, m, unknown, known, dist2, idx, (cudaStream_t)stream);
}
void ref_three_interpolate(int b, int c, int m, int n, const float* points, const int* idx, const float* weight, float* out, void* stream) {
    three_interpolate_kernel_launcher_fast(b, c, m, n, points, idx, weight, out, (cudaStream_t)stream);
}
void ref_three_interpolate_grad(int b, int c, int n, int m, const float* grad_out, const int* idx, const float* weight, float* grad_points, void* stream) {
    three_interpolate_grad_kernel_launcher_fast(b, c, n, m, grad_out, idx, weight, grad_points, (cudaStream_t)stream);
}
void ref_group_points(int b, int c, int n, int npoints, int nsample, const float* points, const int* idx, float* out, void* stream) {
    group_points_kernel_launcher_fast(b, c, n, npoints, nsample, points, idx, out, (cudaStream_t)stream);
}
void ref_group_points_grad(int b, int c, int n, int npoints, int nsample, const float* grad_out, const int* idx, float* grad_points, void* stream) {
    group_points_grad_kernel_launcher_fast(b, c, n, npoints, nsample, grad_out, idx, grad_points, (cudaStream_t)stream);
}
void ref_gather_points(int b, int c, int n, int npoints, const float* points, const int* idx, float* out, void* stream) {
    gather_points_kernel_launcher_fast(b, c, n, npoints, points, idx, out, (cudaStream_t)stream);
}
void ref_gather_points_grad(int b, int c, int n, int npoints, const float* grad_out, const int* idx, float* grad_points, void* stream) {
    gather_points_grad_kernel_launcher_fast(b, c, n, npoints, grad_out, idx, grad_points, (cudaStream_t)stream);
}

}  // extern "C"
