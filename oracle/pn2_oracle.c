/*
 * pn2_oracle.c -- CPU restatement of HOTrack's pointnet_lib CUDA kernels.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in hotrack_b200/ may import, link or call
 * this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * reference legs use it, and there only as the checker / timed CPU baseline.
 *
 * Every function follows one reference kernel and reproduces its arithmetic
 * bit-for-bit on the host: the same fp32 operation order, the same FMA
 * contraction nvcc 12.9 -O2 applies to the reference sources for sm_100
 * (checked in the SASS: FADD,FADD,FMUL,FADD,FFMA,FFMA), the same tie rules.
 * Citations are relative to /root/reference/network/models/pointnet_lib/src/.
 *
 * Parity pin: the reference ships no golden vectors or tests for this path
 * (SURVEY.md section 4).  The pin is tests/golden/ *.npz, produced on a B200 by
 * running the reference's own unmodified kernels (oracle/_ref, built by
 * oracle/Makefile from the sources where they lie) via tools/make_golden.py.
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (see oracle/Makefile).
 * -ffp-contract=off matters: every contraction below is explicit.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* Squared distance exactly as the reference kernels evaluate it.
 * Source expression (sampling_gpu.cu:133, ball_query_gpu.cu:33,
 * interpolate_gpu.cu:40,108):  dx*dx + dy*dy + dz*dz
 * nvcc contraction:            fma(dz,dz, fma(dx,dx, rn(dy*dy)))           */
static inline float pn2_dist2(float ax, float ay, float az,
                              float bx, float by, float bz) {
    float dx = ax - bx;
    float dy = ay - by;
    float dz = az - bz;
    float t = dy * dy;
    t = fmaf(dx, dx, t);
    return fmaf(dz, dz, t);
}

/* cuda_utils.h:10-14 opt_n_threads: largest power of two <= work_size, capped
 * at 1024, computed through a double log ratio and truncation. */
int pn2o_opt_n_threads(int work_size) {
    const int pow_2 = (int)(log((double)work_size) / log(2.0));
    int v = 1 << pow_2;
    if (v > 1024) v = 1024;
    if (v < 1) v = 1;
    return v;
}

/* sampling_gpu.cu:86-91 __update: left operand wins ties. */
static inline void fps_update(float *dists, int *dists_i, int i1, int i2) {
    const float v1 = dists[i1], v2 = dists[i2];
    const int a = dists_i[i1], b = dists_i[i2];
    dists[i1] = fmaxf(v1, v2);
    dists_i[i1] = v2 > v1 ? b : a;
}

/* sampling_gpu.cu:93-209 furthest_point_sampling_kernel<block_size>, simulated
 * thread by thread: per-thread strided scan with first-strict-max, then the
 * shared-memory tree reduction.  dataset (B,N,3), temp (B,N) in/out (caller
 * pre-fills 1e10, pointnet2_utils.py:28), idxs (B,M). */
void pn2o_furthest_point_sampling(int b, int n, int m, const float *dataset,
                                  float *temp, int *idxs) {
    if (m <= 0) return;
    const int bs = pn2o_opt_n_threads(n);
    float *dists = (float *)malloc(sizeof(float) * (size_t)bs);
    int *dists_i = (int *)malloc(sizeof(int) * (size_t)bs);
    for (int bi = 0; bi < b; ++bi) {
        const float *ds = dataset + (size_t)bi * n * 3;
        float *tp = temp + (size_t)bi * n;
        int *out = idxs + (size_t)bi * m;
        int old = 0;
        out[0] = old;
        for (int j = 1; j < m; ++j) {
            const float x1 = ds[old * 3 + 0], y1 = ds[old * 3 + 1], z1 = ds[old * 3 + 2];
            for (int tid = 0; tid < bs; ++tid) {
                int besti = 0;
                float best = -1.0f;
                for (int k = tid; k < n; k += bs) {
                    float d = pn2_dist2(ds[k * 3 + 0], ds[k * 3 + 1], ds[k * 3 + 2], x1, y1, z1);
                    float d2 = fminf(d, tp[k]);
                    tp[k] = d2;
                    besti = d2 > best ? k : besti;
                    best = d2 > best ? d2 : best;
                }
                dists[tid] = best;
                dists_i[tid] = besti;
            }
            for (int s = bs / 2; s >= 1; s >>= 1)
                for (int tid = 0; tid < s; ++tid) fps_update(dists, dists_i, tid, tid + s);
            old = dists_i[0];
            out[j] = old;
        }
    }
    free(dists);
    free(dists_i);
}

/* ball_query_gpu.cu:9-45.  idx (B,M,nsample) must be zero-filled by the caller
 * (pointnet2_utils.py:262); rows without a hit stay untouched. */
void pn2o_ball_query(int b, int n, int m, float radius, int nsample,
                     const float *new_xyz, const float *xyz, int *idx) {
    const float radius2 = radius * radius;
    for (int bi = 0; bi < b; ++bi)
        for (int p = 0; p < m; ++p) {
            const float *c = new_xyz + ((size_t)bi * m + p) * 3;
            const float *pts = xyz + (size_t)bi * n * 3;
            int *o = idx + ((size_t)bi * m + p) * nsample;
            int cnt = 0;
            for (int k = 0; k < n; ++k) {
                float d2 = pn2_dist2(c[0], c[1], c[2], pts[k * 3], pts[k * 3 + 1], pts[k * 3 + 2]);
                if (d2 < radius2) {
                    if (cnt == 0)
                        for (int l = 0; l < nsample; ++l) o[l] = k;
                    o[cnt] = k;
                    ++cnt;
                    if (cnt >= nsample) break;
                }
            }
        }
}

/* interpolate_gpu.cu:9-57 knn_kernel_fast: stable ascending insertion with
 * fp64 compares of fp32 values, 1e40 sentinel (-> +inf when stored as fp32). */
void pn2o_knn(int b, int n, int m, int k, const float *unknown, const float *known,
              float *dist2, int *idx) {
    double *best = (double *)malloc(sizeof(double) * (size_t)(k > 0 ? k : 1));
    int *besti = (int *)malloc(sizeof(int) * (size_t)(k > 0 ? k : 1));
    for (int bi = 0; bi < b; ++bi)
        for (int p = 0; p < n; ++p) {
            const float *u = unknown + ((size_t)bi * n + p) * 3;
            const float *kn = known + (size_t)bi * m * 3;
            for (int i = 0; i < k; ++i) { best[i] = 1e40; besti[i] = 0; }
            for (int i = 0; i < m; ++i) {
                float d = pn2_dist2(u[0], u[1], u[2], kn[i * 3], kn[i * 3 + 1], kn[i * 3 + 2]);
                for (int j = 0; j < k; ++j) {
                    if (d < best[j]) {
                        for (int l = k - 1; l > j; --l) { best[l] = best[l - 1]; besti[l] = besti[l - 1]; }
                        best[j] = d;
                        besti[j] = i;
                        break;
                    }
                }
            }
            for (int i = 0; i < k; ++i) {
                idx[((size_t)bi * n + p) * k + i] = besti[i];
                dist2[((size_t)bi * n + p) * k + i] = (float)best[i];
            }
        }
    free(best);
    free(besti);
}

/* interpolate_gpu.cu:81-124 three_nn_kernel_fast. */
void pn2o_three_nn(int b, int n, int m, const float *unknown, const float *known,
                   float *dist2, int *idx) {
    for (int bi = 0; bi < b; ++bi)
        for (int p = 0; p < n; ++p) {
            const float *u = unknown + ((size_t)bi * n + p) * 3;
            const float *kn = known + (size_t)bi * m * 3;
            double best1 = 1e40, best2 = 1e40, best3 = 1e40;
            int i1 = 0, i2 = 0, i3 = 0;
            for (int k = 0; k < m; ++k) {
                float d = pn2_dist2(u[0], u[1], u[2], kn[k * 3], kn[k * 3 + 1], kn[k * 3 + 2]);
                if (d < best1) {
                    best3 = best2; i3 = i2; best2 = best1; i2 = i1; best1 = d; i1 = k;
                } else if (d < best2) {
                    best3 = best2; i3 = i2; best2 = d; i2 = k;
                } else if (d < best3) {
                    best3 = d; i3 = k;
                }
            }
            float *od = dist2 + ((size_t)bi * n + p) * 3;
            int *oi = idx + ((size_t)bi * n + p) * 3;
            od[0] = (float)best1; od[1] = (float)best2; od[2] = (float)best3;
            oi[0] = i1; oi[1] = i2; oi[2] = i3;
        }
}

/* interpolate_gpu.cu:149-169.  Source: w0*p0 + w1*p1 + w2*p2, contracted by
 * nvcc as fma(w2,p2, fma(w0,p0, rn(w1*p1))).  points (B,C,M), idx/weight
 * (B,N,3), out (B,C,N). */
void pn2o_three_interpolate(int b, int c, int m, int n, const float *points,
                            const int *idx, const float *weight, float *out) {
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci) {
            const float *pr = points + ((size_t)bi * c + ci) * m;
            float *orow = out + ((size_t)bi * c + ci) * n;
            for (int p = 0; p < n; ++p) {
                const float *w = weight + ((size_t)bi * n + p) * 3;
                const int *ix = idx + ((size_t)bi * n + p) * 3;
                float t = w[1] * pr[ix[1]];
                t = fmaf(w[0], pr[ix[0]], t);
                orow[p] = fmaf(w[2], pr[ix[2]], t);
            }
        }
}

/* interpolate_gpu.cu:192-214.  The reference scatters with fp32 atomicAdd in an
 * unspecified order into a zeroed buffer; this restatement adds in ascending
 * point order (products rounded to fp32 first, as the kernel does), so
 * comparisons against it are tolerance-based, not bit-exact. */
void pn2o_three_interpolate_grad(int b, int c, int n, int m, const float *grad_out,
                                 const int *idx, const float *weight, float *grad_points) {
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci) {
            const float *g = grad_out + ((size_t)bi * c + ci) * n;
            float *gp = grad_points + ((size_t)bi * c + ci) * m;
            for (int p = 0; p < n; ++p) {
                const float *w = weight + ((size_t)bi * n + p) * 3;
                const int *ix = idx + ((size_t)bi * n + p) * 3;
                gp[ix[0]] += g[p] * w[0];
                gp[ix[1]] += g[p] * w[1];
                gp[ix[2]] += g[p] * w[2];
            }
        }
}

/* group_points_gpu.cu:47-66.  points (B,C,N), idx (B,S,K) -> out (B,C,S,K). */
void pn2o_group_points(int b, int c, int n, int npoints, int nsample,
                       const float *points, const int *idx, float *out) {
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci)
            for (int s = 0; s < npoints * nsample; ++s)
                out[((size_t)bi * c + ci) * npoints * nsample + s] =
                    points[((size_t)bi * c + ci) * n + idx[(size_t)bi * npoints * nsample + s]];
}

/* group_points_gpu.cu:8-25 (atomic scatter; see three_interpolate_grad note). */
void pn2o_group_points_grad(int b, int c, int n, int npoints, int nsample,
                            const float *grad_out, const int *idx, float *grad_points) {
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci)
            for (int s = 0; s < npoints * nsample; ++s)
                grad_points[((size_t)bi * c + ci) * n + idx[(size_t)bi * npoints * nsample + s]] +=
                    grad_out[((size_t)bi * c + ci) * npoints * nsample + s];
}

/* sampling_gpu.cu:8-24.  points (B,C,N), idx (B,M) -> out (B,C,M). */
void pn2o_gather_points(int b, int c, int n, int m, const float *points,
                        const int *idx, float *out) {
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci)
            for (int p = 0; p < m; ++p)
                out[((size_t)bi * c + ci) * m + p] =
                    points[((size_t)bi * c + ci) * n + idx[(size_t)bi * m + p]];
}

/* sampling_gpu.cu:46-63 (atomic scatter). */
void pn2o_gather_points_grad(int b, int c, int n, int m, const float *grad_out,
                             const int *idx, float *grad_points) {
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci)
            for (int p = 0; p < m; ++p)
                grad_points[((size_t)bi * c + ci) * n + idx[(size_t)bi * m + p]] +=
                    grad_out[((size_t)bi * c + ci) * m + p];
}
