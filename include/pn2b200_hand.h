/*
 * pn2b200_hand.h -- C ABI of the hand-frame kernels of libpn2b200.so: what HandTrackNet runs AROUND the pointnet_lib
 * path every frame.  Conventions as in pn2b200.h: raw device pointers, int sizes,
 * caller-owned buffers, asynchronous on `stream`, int status (0 = ok, see pn2_last_error()).
 */
#ifndef PN2B200_HAND_H_
#define PN2B200_HAND_H_
#include "pn2b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Batched rigid alignment  y_i ~ R x_i + t  of b problems of n corresponding 3-D points each (Kabsch), replacing
 * reference network/models/hand_utils.py:42-66 solve_rot_and_trans -- whose 3x3 SVD runs on the CPU (:57-61, "convert to
 * cpu to speed up"): one device<->host round trip per call, three per training step (hand_network.py:100,182-183).
 *   x (n,3) shared by all problems (x_batched == 0) or (b,n,3);  y (b,n,3);  R (b,3,3) row-major;  t (b,3) (= the
 *   reference's (b,3,1)); aux (b,12) scratch the backward needs (nullable when no gradient is wanted).
 * R = V diag(1,1,det(V U^T)) U^T for w = sum (x_i-cx)(y_i-cy)^T = U S V^T: a proper rotation (det +1) also for
 * reflected / coplanar point sets, independent of the sign choices of the SVD. */
int pn2_kabsch_fwd(int b, int n, const float* x, int x_batched, const float* y, float* R, float* t, float* aux,
                   pn2_stream_t stream);

/* Gradient of pn2_kabsch_fwd: grad_R (b,3,3), grad_t (b,3) (either may be NULL = zero) -> grad_y (b,n,3) and, when
 * x is batched, grad_x (b,n,3) (either may be NULL).  The reference gets this from autograd through torch.svd. */
int pn2_kabsch_bwd(int b, int n, const float* x, int x_batched, const float* y, const float* R, const float* aux,
                   const float* grad_R, const float* grad_t, float* grad_x, float* grad_y, pn2_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PN2B200_HAND_H_ */
