// kabsch.cu -- batched rigid alignment (Kabsch) of small point sets on the GPU, forward and backward, sm_100a.
//
// Replaces reference network/models/hand_utils.py:42-66 (solve_rot_and_trans): y ~ R x + t for B pairs of `num`
// corresponding points (the 6 palm keypoints: hand_network.py:100,182-183).  The reference moves the 3x3 covariance to
// the CPU for torch.svd and back -- three device<->host round trips per training step and one per tracked frame, the
// only host synchronisations of HandTrackNet.forward.  Here one thread solves one problem entirely in registers:
//
//   w = sum_i (x_i - cx)(y_i - cy)^T                    (hand_utils.py:53-56)
//   eigen-decomposition of w^T w by cyclic Jacobi rotations in double precision -> right singular vectors v1, v2,
//   v3 = v1 x v2; u_i = w v_i / s_i (i = 1,2; Gram-Schmidt), u3 = u1 x u2
//   R = v1 u1^T + v2 u2^T + v3 u3^T
//
// which is the reference's  R = V diag(1, 1, det(V U^T)) U^T  (:61-63) whatever signs its SVD picks: with both bases
// right-handed the determinant factor is +1, and the signed third singular value s3' = u3 . (w v3) carries the
// reflection case.  t = cy - R cx (:64).
//
// Backward (the reference differentiates through torch.svd; losses hand_pred_r_loss / hand_pred_t_loss of
// hand_network.py:203-204 need it): with M = w^T = R P, P = U diag(s1, s2, s3') U^T the symmetric polar factor,
//   dL/dM = R (K - K^T),   K = U [ (U^T R^T G U)_ij / (s_i + s_j) ] U^T,   G = dL/dR
// (from R^T dM - dM^T R = Omega P + P Omega for the skew Omega = R^T dR), then the chain through w, cx, cy and t.
#include "pn2_common.cuh"
#include "../../include/pn2b200_hand.h"  // every extern "C" definition is checked against its declaration

namespace pn2 {
namespace {

struct M3 {
    double m[3][3];
};
__device__ __forceinline__ M3 mul(const M3& a, const M3& b) {
    M3 c;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) c.m[i][j] = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j] + a.m[i][2] * b.m[2][j];
    return c;
}
__device__ __forceinline__ M3 tr(const M3& a) {
    M3 c;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) c.m[i][j] = a.m[j][i];
    return c;
}
__device__ __forceinline__ void cross(const double* a, const double* b, double* c) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ __forceinline__ double nrm(const double* a) { return sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }

// symmetric 3x3 -> eigenvalues (descending) and eigenvectors (columns of V), cyclic Jacobi
__device__ void eig_sym3(M3 a, double (&lam)[3], M3& v) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) v.m[i][j] = i == j ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 12; ++sweep) {
        const double off = a.m[0][1] * a.m[0][1] + a.m[0][2] * a.m[0][2] + a.m[1][2] * a.m[1][2];
        const double dia = a.m[0][0] * a.m[0][0] + a.m[1][1] * a.m[1][1] + a.m[2][2] * a.m[2][2];
        if (off <= 1e-34 * dia || off == 0.0) break;
#pragma unroll
        for (int pq = 0; pq < 3; ++pq) {
            const int p = pq == 2 ? 1 : 0, q = pq == 0 ? 1 : 2;
            const double apq = a.m[p][q];
            if (apq == 0.0) continue;
            const double theta = (a.m[q][q] - a.m[p][p]) / (2.0 * apq);
            const double tt = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
            const double c = 1.0 / sqrt(tt * tt + 1.0), s = tt * c;
#pragma unroll
            for (int k = 0; k < 3; ++k) {  // A <- A J
                const double akp = a.m[k][p], akq = a.m[k][q];
                a.m[k][p] = c * akp - s * akq;
                a.m[k][q] = s * akp + c * akq;
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) {  // A <- J^T A
                const double apk = a.m[p][k], aqk = a.m[q][k];
                a.m[p][k] = c * apk - s * aqk;
                a.m[q][k] = s * apk + c * aqk;
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) {  // V <- V J
                const double vkp = v.m[k][p], vkq = v.m[k][q];
                v.m[k][p] = c * vkp - s * vkq;
                v.m[k][q] = s * vkp + c * vkq;
            }
        }
    }
    lam[0] = a.m[0][0]; lam[1] = a.m[1][1]; lam[2] = a.m[2][2];
    // sort descending (columns of V follow)
#pragma unroll
    for (int pass = 0; pass < 2; ++pass)
#pragma unroll
        for (int i = 0; i < 2; ++i)
            if (lam[i] < lam[i + 1]) {
                const double tl = lam[i]; lam[i] = lam[i + 1]; lam[i + 1] = tl;
#pragma unroll
                for (int k = 0; k < 3; ++k) { const double tv = v.m[k][i]; v.m[k][i] = v.m[k][i + 1]; v.m[k][i + 1] = tv; }
            }
}

// aux per problem: U (9, row-major, columns u1 u2 u3), s1, s2, s3' (3)
__global__ void kabsch_fwd_kernel(int b, int n, const float* __restrict__ x, long long x_stride,
                                  const float* __restrict__ y, float* __restrict__ R, float* __restrict__ t,
                                  float* __restrict__ aux) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b) return;
    const float* xs = x + (size_t)i * x_stride;
    const float* ys = y + (size_t)i * n * 3;
    double cx[3] = {0, 0, 0}, cy[3] = {0, 0, 0};
    for (int k = 0; k < n; ++k)
#pragma unroll
        for (int d = 0; d < 3; ++d) { cx[d] += xs[k * 3 + d]; cy[d] += ys[k * 3 + d]; }
#pragma unroll
    for (int d = 0; d < 3; ++d) { cx[d] /= n; cy[d] /= n; }
    M3 w;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int c = 0; c < 3; ++c) w.m[a][c] = 0.0;
    for (int k = 0; k < n; ++k)
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int c = 0; c < 3; ++c) w.m[a][c] += ((double)xs[k * 3 + a] - cx[a]) * ((double)ys[k * 3 + c] - cy[c]);
    double lam[3];
    M3 v;
    eig_sym3(mul(tr(w), w), lam, v);
    double v1[3] = {v.m[0][0], v.m[1][0], v.m[2][0]}, v2[3] = {v.m[0][1], v.m[1][1], v.m[2][1]}, v3[3];
    cross(v1, v2, v3);
    double u1[3], u2[3], u3[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        u1[a] = w.m[a][0] * v1[0] + w.m[a][1] * v1[1] + w.m[a][2] * v1[2];
        u2[a] = w.m[a][0] * v2[0] + w.m[a][1] * v2[1] + w.m[a][2] * v2[2];
    }
    double s1 = nrm(u1), s2;
    if (s1 < 1e-300) { u1[0] = 1; u1[1] = 0; u1[2] = 0; s1 = 0; } else { u1[0] /= s1; u1[1] /= s1; u1[2] /= s1; }
    const double d12 = u1[0] * u2[0] + u1[1] * u2[1] + u1[2] * u2[2];
#pragma unroll
    for (int a = 0; a < 3; ++a) u2[a] -= d12 * u1[a];
    s2 = nrm(u2);
    if (s2 < 1e-150 * (s1 + 1e-300)) {  // rank <= 1: any unit vector orthogonal to u1
        const double e[3] = {fabs(u1[0]) < 0.9 ? 1.0 : 0.0, fabs(u1[0]) < 0.9 ? 0.0 : 1.0, 0.0};
        cross(u1, e, u2);
        const double q = nrm(u2);
        u2[0] /= q; u2[1] /= q; u2[2] /= q;
        s2 = 0;
    } else {
        u2[0] /= s2; u2[1] /= s2; u2[2] /= s2;
    }
    cross(u1, u2, u3);
    double wv3[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) wv3[a] = w.m[a][0] * v3[0] + w.m[a][1] * v3[1] + w.m[a][2] * v3[2];
    const double s3 = u3[0] * wv3[0] + u3[1] * wv3[1] + u3[2] * wv3[2];  // signed
    double Rm[3][3];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int c = 0; c < 3; ++c) Rm[a][c] = v1[a] * u1[c] + v2[a] * u2[c] + v3[a] * u3[c];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int c = 0; c < 3; ++c) R[(size_t)i * 9 + a * 3 + c] = (float)Rm[a][c];
        t[(size_t)i * 3 + a] = (float)(cy[a] - (Rm[a][0] * cx[0] + Rm[a][1] * cx[1] + Rm[a][2] * cx[2]));
    }
    if (aux) {
        float* o = aux + (size_t)i * 12;
#pragma unroll
        for (int a = 0; a < 3; ++a) { o[a * 3 + 0] = (float)u1[a]; o[a * 3 + 1] = (float)u2[a]; o[a * 3 + 2] = (float)u3[a]; }
        o[9] = (float)s1; o[10] = (float)s2; o[11] = (float)s3;
    }
}

__global__ void kabsch_bwd_kernel(int b, int n, const float* __restrict__ x, long long x_stride,
                                  const float* __restrict__ y, const float* __restrict__ R,
                                  const float* __restrict__ aux, const float* __restrict__ gR,
                                  const float* __restrict__ gt, float* __restrict__ gx, float* __restrict__ gy) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b) return;
    const float* xs = x + (size_t)i * x_stride;
    const float* ys = y + (size_t)i * n * 3;
    double cx[3] = {0, 0, 0}, cy[3] = {0, 0, 0};
    for (int k = 0; k < n; ++k)
#pragma unroll
        for (int d = 0; d < 3; ++d) { cx[d] += xs[k * 3 + d]; cy[d] += ys[k * 3 + d]; }
#pragma unroll
    for (int d = 0; d < 3; ++d) { cx[d] /= n; cy[d] /= n; }
    M3 Rm, U, G;
    double s[3], g_t[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        g_t[a] = gt ? (double)gt[(size_t)i * 3 + a] : 0.0;
        s[a] = aux[(size_t)i * 12 + 9 + a];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            Rm.m[a][c] = R[(size_t)i * 9 + a * 3 + c];
            U.m[a][c] = aux[(size_t)i * 12 + a * 3 + c];
        }
    }
    // t = cy - R cx: dL/dR gets -gt cx^T
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int c = 0; c < 3; ++c) G.m[a][c] = (gR ? (double)gR[(size_t)i * 9 + a * 3 + c] : 0.0) - g_t[a] * cx[c];
    const M3 H = mul(tr(U), mul(mul(tr(Rm), G), U));  // U^T R^T G U
    M3 Kp;
    const double smax = fmax(fabs(s[0]), 1e-30);
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            double den = s[a] + s[c];
            if (fabs(den) < 1e-12 * smax) den = den < 0 ? -1e-12 * smax : 1e-12 * smax;  // degenerate: the rotation is not unique there
            Kp.m[a][c] = H.m[a][c] / den;
        }
    const M3 K = mul(U, mul(Kp, tr(U)));
    M3 KmKt;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int c = 0; c < 3; ++c) KmKt.m[a][c] = K.m[a][c] - K.m[c][a];
    const M3 dM = mul(Rm, KmKt);  // dL/dM, M = w^T  ->  dL/dw = dM^T
    // centroids: dL/dcy = gt, dL/dcx = -R^T gt
    double gcx[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) gcx[a] = -(Rm.m[0][a] * g_t[0] + Rm.m[1][a] * g_t[1] + Rm.m[2][a] * g_t[2]);
    for (int k = 0; k < n; ++k) {
        double xc[3], yc[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) { xc[d] = (double)xs[k * 3 + d] - cx[d]; yc[d] = (double)ys[k * 3 + d] - cy[d]; }
        if (gy) {
#pragma unroll
            for (int c = 0; c < 3; ++c)  // dL/dyc[c] = sum_a dLdw[a][c] xc[a] = sum_a dM[c][a] xc[a]; sum over k of it is 0
                gy[((size_t)i * n + k) * 3 + c] = (float)(dM.m[c][0] * xc[0] + dM.m[c][1] * xc[1] + dM.m[c][2] * xc[2] + g_t[c] / n);
        }
        if (gx) {
#pragma unroll
            for (int a = 0; a < 3; ++a)  // dL/dxc[a] = sum_c dLdw[a][c] yc[c] = sum_c dM[c][a] yc[c]
                gx[((size_t)i * n + k) * 3 + a] = (float)(dM.m[0][a] * yc[0] + dM.m[1][a] * yc[1] + dM.m[2][a] * yc[2] + gcx[a] / n);
        }
    }
}

}  // namespace
}  // namespace pn2

using namespace pn2;

extern "C" int pn2_kabsch_fwd(int b, int n, const float* x, int x_batched, const float* y, float* R, float* t, float* aux,
                              pn2_stream_t stream) {
    if (b < 0 || n <= 0) return fail_arg("pn2_kabsch_fwd", "bad size");
    if (b == 0) return 0;
    if (!x || !y || !R || !t) return fail_arg("pn2_kabsch_fwd", "null pointer");
    kabsch_fwd_kernel<<<(b + 63) / 64, 64, 0, (cudaStream_t)stream>>>(b, n, x, x_batched ? (long long)n * 3 : 0, y, R, t, aux);
    PN2_CHECK_LAUNCH("kabsch_fwd_kernel");
    return 0;
}

extern "C" int pn2_kabsch_bwd(int b, int n, const float* x, int x_batched, const float* y, const float* R, const float* aux,
                              const float* grad_R, const float* grad_t, float* grad_x, float* grad_y, pn2_stream_t stream) {
    if (b < 0 || n <= 0) return fail_arg("pn2_kabsch_bwd", "bad size");
    if (b == 0) return 0;
    if (!x || !y || !R || !aux || (!grad_R && !grad_t)) return fail_arg("pn2_kabsch_bwd", "null pointer");
    if (grad_x && !x_batched) return fail_arg("pn2_kabsch_bwd", "grad_x needs a per-problem x");
    kabsch_bwd_kernel<<<(b + 63) / 64, 64, 0, (cudaStream_t)stream>>>(b, n, x, x_batched ? (long long)n * 3 : 0, y, R, aux,
                                                                       grad_R, grad_t, grad_x, grad_y);
    PN2_CHECK_LAUNCH("kabsch_bwd_kernel");
    return 0;
}
