/*
 * pn2b200.h -- C ABI of libpn2b200.so, the sm_100a replacement for HOTrack's
 * `pointnet2_cuda` extension (reference: network/models/pointnet_lib/src/).
 *
 * One entry point per reference kernel launcher.  Same argument order and
 * meaning as the launcher it replaces; raw device pointers (fp32 / int32,
 * contiguous), int sizes, the CUDA stream last.  Differences from the reference:
 *   - every function returns 0 on success or a non-zero status (a cudaError_t
 *     value, or PN2_EINVAL for a rejected argument) instead of printing and
 *     calling exit(-1) (reference: e.g. ball_query_gpu.cu:62-66);
 *     pn2_last_error() describes the last failure of the calling thread;
 *   - the caller owns every buffer, as in the reference (outputs, the FPS
 *     scratch `temp`, zero-filled ball-query / grad outputs; see
 *     pointnet_lib/pointnet2_utils.py:27-28,71,186,232,262);
 *   - calls are asynchronous on `stream`, never synchronise, never allocate.
 * All paths are relative to /root/reference/network/models/pointnet_lib/.
 */
#ifndef PN2B200_H_
#define PN2B200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define PN2_EINVAL 100001 /* rejected argument (negative size, k too large, ...) */

typedef void* pn2_stream_t; /* cudaStream_t */

/* Library identification.  pn2_version() = 10000*major + 100*minor + patch. */
int pn2_version(void);
/* Number of CUDA kernels this library has launched in this process (all threads). */
long long pn2_launch_count(void);
/* Message for the last non-zero status returned on this thread ("" if none). */
const char* pn2_last_error(void);

/* replaces furthest_point_sampling_kernel_launcher  (src/sampling_gpu.h:26-27,
 * src/sampling_gpu.cu:211-253; kernel :93-209).
 * dataset (B,N,3); temp (B,N) running min distances, read on entry (the caller
 * pre-fills 1e10) and written back on exit -- may be NULL, meaning "start from
 * 1e10, do not write back"; idxs (B,M) int32, idxs[:,0]==0.  Tie rule identical
 * to the reference block reduction (block size from src/cuda_utils.h:10-14). */
int pn2_furthest_point_sampling(int b, int n, int m, const float* dataset, float* temp, int* idxs,
                                pn2_stream_t stream);

/* replaces ball_query_kernel_launcher_fast (src/ball_query_gpu.h:12-13,
 * src/ball_query_gpu.cu:48-67; kernel :9-45).
 * new_xyz (B,M,3) centres, xyz (B,N,3), idx (B,M,nsample) int32.  Rows with no
 * point inside the radius are left untouched (caller zero-fills). */
int pn2_ball_query(int b, int n, int m, float radius, int nsample, const float* new_xyz, const float* xyz,
                   int* idx, pn2_stream_t stream);

/* replaces knn_kernel_launcher_fast (src/interpolate_gpu.h:19-20,
 * src/interpolate_gpu.cu:60-79; kernel :9-57).
 * unknown (B,N,3) queries, known (B,M,3); dist2/idx (B,N,k): the k smallest
 * squared distances ascending, ties by lower index; slots beyond M hold
 * (+inf, 0).  k <= PN2_KNN_MAX_K (reference: k <= 200). */
#define PN2_KNN_MAX_K 1024
int pn2_knn(int b, int n, int m, int k, const float* unknown, const float* known, float* dist2, int* idx,
            pn2_stream_t stream);

/* replaces three_nn_kernel_launcher_fast (src/interpolate_gpu.h:13-14,
 * src/interpolate_gpu.cu:127-146; kernel :81-124). */
int pn2_three_nn(int b, int n, int m, const float* unknown, const float* known, float* dist2, int* idx,
                 pn2_stream_t stream);

/* replaces three_interpolate_kernel_launcher_fast (src/interpolate_gpu.h:26-27,
 * src/interpolate_gpu.cu:171-189; kernel :149-169).
 * points (B,C,M), idx/weight (B,N,3) -> out (B,C,N). */
int pn2_three_interpolate(int b, int c, int m, int n, const float* points, const int* idx, const float* weight,
                          float* out, pn2_stream_t stream);

/* replaces three_interpolate_grad_kernel_launcher_fast (src/interpolate_gpu.h:33-34,
 * src/interpolate_gpu.cu:216-232; kernel :192-214).
 * grad_out (B,C,N) -> accumulated INTO grad_points (B,C,M) (caller zero-fills). */
int pn2_three_interpolate_grad(int b, int c, int n, int m, const float* grad_out, const int* idx,
                               const float* weight, float* grad_points, pn2_stream_t stream);

/* replaces group_points_kernel_launcher_fast (src/group_points_gpu.h:13-14,
 * src/group_points_gpu.cu:69-86; kernel :47-66).
 * points (B,C,N), idx (B,npoints,nsample) -> out (B,C,npoints,nsample). */
int pn2_group_points(int b, int c, int n, int npoints, int nsample, const float* points, const int* idx,
                     float* out, pn2_stream_t stream);

/* replaces group_points_grad_kernel_launcher_fast (src/group_points_gpu.h:19-20,
 * src/group_points_gpu.cu:27-44; kernel :8-25).  Accumulates into grad_points (B,C,N). */
int pn2_group_points_grad(int b, int c, int n, int npoints, int nsample, const float* grad_out, const int* idx,
                          float* grad_points, pn2_stream_t stream);

/* replaces gather_points_kernel_launcher_fast (src/sampling_gpu.h:12-13,
 * src/sampling_gpu.cu:26-44; kernel :8-24).  points (B,C,N), idx (B,M) -> out (B,C,M). */
int pn2_gather_points(int b, int c, int n, int npoints, const float* points, const int* idx, float* out,
                      pn2_stream_t stream);

/* replaces gather_points_grad_kernel_launcher_fast (src/sampling_gpu.h:19-20,
 * src/sampling_gpu.cu:65-83; kernel :46-63).  Accumulates into grad_points (B,C,N). */
int pn2_gather_points_grad(int b, int c, int n, int npoints, const float* grad_out, const int* idx,
                           float* grad_points, pn2_stream_t stream);

/* ---- entry points with no reference C counterpart -------------------------------------- */

/* Flat-buffer Adam step; replaces the torch.optim.Adam the reference trains with
 * (network/trainer.py:66-73).  All four buffers hold n fp32 values, 16-byte aligned.
 * Update rule = torch.optim.Adam with L2 weight decay; `step` counts from 1; the gradient is
 * multiplied by grad_scale first (1/world_size after a SUM all-reduce).  step_dev != NULL: the step
 * number lives on the device (t = *step_dev + 1, incremented after the update), so the call can be
 * captured once in a CUDA graph and replayed; `step` is then ignored. */
int pn2_adam_step(long long n, float* params, const float* grads, float* exp_avg, float* exp_avg_sq, float lr,
                  float beta1, float beta2, float eps, float weight_decay, int step, int* step_dev, float grad_scale,
                  pn2_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PN2B200_H_ */
