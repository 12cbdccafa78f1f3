// gemm_tc_check.cu -- standalone check + timing of the row-matrix GEMM entry points of libpn2b200.so
// (pn2_mlp_gemm_fwd / pn2_mlp_gemm_dgrad) against a scalar CPU restatement, without importing torch.
// Development aid: the parity tests proper are tests/test_fused_gpu.py.
//   nvcc -arch=sm_100a -o gemm_tc_check tools/dev/gemm_tc_check.cu -Lhotrack_b200 -lpn2b200
//   PN2_GEMM_IMPL=tc|mma LD_LIBRARY_PATH=hotrack_b200 ./gemm_tc_check [time]
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/pn2b200_mlp.h"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

static unsigned long long rng = 88172645463325252ull;
static float frand() { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return (float)((rng >> 11) & 0xFFFFFF) / 16777216.f * 2.f - 1.f; }
static float h16(float v) { return __half2float(__float2half_rn(v)); }
static float b16(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

__global__ void flush_read_kernel(const float4* p, size_t n, float* sink) {
    float acc = 0.f;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) { float4 v = p[i]; acc += v.x + v.y + v.z + v.w; }
    if (acc == 123.456f) *sink = acc;
}
static void flush_l2(float* buf, int) { flush_read_kernel<<<1184, 256>>>((const float4*)buf, (size_t)(256 << 20) / 16, buf); }

template <class T> T* dev(const std::vector<T>& h) { T* d; CK(cudaMalloc(&d, h.size() * sizeof(T) + 16)); CK(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice)); return d; }

static int check_fwd(int R, int K, int N, bool affine) {
    std::vector<__half> x((size_t)R * K), w((size_t)N * K);
    std::vector<float> xf((size_t)R * K), wf((size_t)N * K), sc(K), sh(K), cen(N);
    for (size_t i = 0; i < x.size(); ++i) { xf[i] = h16(frand() * 2.f); x[i] = __float2half_rn(xf[i]); }
    for (size_t i = 0; i < w.size(); ++i) { wf[i] = h16(frand() * 0.3f); w[i] = __float2half_rn(wf[i]); }
    for (int k = 0; k < K; ++k) { sc[k] = 0.5f + 0.5f * frand(); sh[k] = 0.3f * frand(); }
    for (int n = 0; n < N; ++n) cen[n] = 0.2f * frand();
    std::vector<float> yref((size_t)R * N), s1(N, 0.f), s2(N, 0.f);
    std::vector<float> xa((size_t)R * K);
    for (size_t i = 0; i < xa.size(); ++i) { int k = i % K; xa[i] = affine ? h16(fmaxf(fmaf(xf[i], sc[k], sh[k]), 0.f)) : xf[i]; }
    for (int r = 0; r < R; ++r)
        for (int n = 0; n < N; ++n) {
            double a = 0;
            for (int k = 0; k < K; ++k) a += (double)xa[(size_t)r * K + k] * wf[(size_t)n * K + k];
            float v = h16((float)a - cen[n]);
            yref[(size_t)r * N + n] = v; s1[n] += v; s2[n] += v * v;
        }
    __half *dx = dev(x), *dw = dev(w); float *dsc = dev(sc), *dsh = dev(sh), *dcen = dev(cen);
    __half* dy; CK(cudaMalloc(&dy, (size_t)R * N * 2)); CK(cudaMemset(dy, 0xff, (size_t)R * N * 2));
    float* dst; CK(cudaMalloc(&dst, 2 * N * 4)); CK(cudaMemset(dst, 0, 2 * N * 4));
    int rc = pn2_mlp_gemm_fwd(R, K, N, dx, K, affine ? dsc : nullptr, affine ? dsh : nullptr, dw, dcen, dy, N, dst, 0);
    if (rc) { printf("fwd R=%d K=%d N=%d: launch failed rc=%d\n", R, K, N, rc); return 1; }
    CK(cudaDeviceSynchronize());
    std::vector<__half> y((size_t)R * N); std::vector<float> st(2 * N);
    CK(cudaMemcpy(y.data(), dy, y.size() * 2, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(st.data(), dst, st.size() * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0, maxs = 0; size_t bad = 0;
    for (size_t i = 0; i < y.size(); ++i) {
        float g = __half2float(y[i]); double e = fabs(g - yref[i]) / (1.0 + fabs(yref[i]));
        if (!(e < 4e-3)) { if (bad < 5) printf("   y[%zu,%zu] = %f ref %f\n", i / N, i % N, g, yref[i]); ++bad; }
        if (e > maxerr) maxerr = e;
    }
    for (int n = 0; n < N; ++n) {
        maxs = fmax(maxs, fabs(st[n] - s1[n]) / (1.0 + fabs(s1[n])));
        maxs = fmax(maxs, fabs(st[N + n] - s2[n]) / (1.0 + fabs(s2[n])));
    }
    const bool ok = bad == 0 && maxs < 5e-3;
    printf("fwd   R=%6d K=%4d N=%4d affine=%d  max rel err %.2e  stats err %.2e  bad %zu  %s\n", R, K, N, affine, maxerr, maxs, bad, ok ? "OK" : "FAIL");
    cudaFree(dx); cudaFree(dw); cudaFree(dsc); cudaFree(dsh); cudaFree(dcen); cudaFree(dy); cudaFree(dst);
    return ok ? 0 : 1;
}

static int check_dgrad(int R, int NR, int KO, bool mask) {
    // dz_prev = (cA*dz + cB*y + cC) * W, evaluated by the kernel as dz*wa^T + y*wb^T - negbias
    std::vector<__nv_bfloat16> dz((size_t)R * NR), wa((size_t)KO * NR);
    std::vector<__half> y((size_t)R * NR), yp((size_t)R * KO), wb((size_t)KO * NR);
    std::vector<float> dzf(dz.size()), wf((size_t)KO * NR), yf(y.size()), ypf(yp.size()), cA(NR), cB(NR), cC(NR), negb(KO), psc(KO), psh(KO), pm(KO), prs(KO);
    for (size_t i = 0; i < dz.size(); ++i) { dzf[i] = b16(frand()); dz[i] = __float2bfloat16_rn(dzf[i]); yf[i] = h16(frand() * 2.f); y[i] = __float2half_rn(yf[i]); }
    for (size_t i = 0; i < wf.size(); ++i) wf[i] = frand() * 0.3f;   // W^T [k][n]
    for (size_t i = 0; i < yp.size(); ++i) { ypf[i] = h16(frand() * 2.f); yp[i] = __float2half_rn(ypf[i]); }
    for (int n = 0; n < NR; ++n) { cA[n] = 0.5f + 0.4f * frand(); cB[n] = 0.1f * frand(); cC[n] = 0.05f * frand(); }
    for (int k = 0; k < KO; ++k) { psc[k] = 0.5f + 0.4f * frand(); psh[k] = 0.2f * frand(); pm[k] = 0.1f * frand(); prs[k] = 1.f + 0.3f * frand(); }
    for (int k = 0; k < KO; ++k) {
        double bsum = 0;
        for (int n = 0; n < NR; ++n) {
            const float wv = wf[(size_t)k * NR + n];
            wa[(size_t)k * NR + n] = __float2bfloat16_rn(cA[n] * wv * 64.f);
            wb[(size_t)k * NR + n] = __float2half_rn(cB[n] * wv * 64.f);
            bsum += (double)cC[n] * wv;
        }
        negb[k] = (float)-bsum;
    }
    std::vector<float> ref((size_t)R * KO), s1(KO, 0.f), s2(KO, 0.f);
    for (int r = 0; r < R; ++r)
        for (int k = 0; k < KO; ++k) {
            double acc = 0;
            for (int n = 0; n < NR; ++n)
                acc += ((double)cA[n] * dzf[(size_t)r * NR + n] + (double)cB[n] * yf[(size_t)r * NR + n] + cC[n]) * wf[(size_t)k * NR + n];
            float v = b16((float)acc);
            if (mask) {
                float yv = ypf[(size_t)r * KO + k];
                if (!(fmaf(yv, psc[k], psh[k]) > 0.f)) v = 0.f;
                s1[k] += v; s2[k] += v * (yv - pm[k]) * prs[k];
            }
            ref[(size_t)r * KO + k] = v;
        }
    auto *ddz = dev(dz), *dwa = dev(wa); auto *dy = dev(y), *dyp = dev(yp), *dwb = dev(wb);
    std::vector<float> uns(1, 1.f / 64.f); float* duns = dev(uns);
    float *dnb = dev(negb), *d1 = dev(psc), *d2 = dev(psh), *d3 = dev(pm), *d4 = dev(prs);
    __nv_bfloat16* dout; CK(cudaMalloc(&dout, (size_t)R * KO * 2)); CK(cudaMemset(dout, 0xff, (size_t)R * KO * 2));
    float* dsum; CK(cudaMalloc(&dsum, 2 * KO * 4)); CK(cudaMemset(dsum, 0, 2 * KO * 4));
    int rc = pn2_mlp_gemm_dgrad(R, NR, KO, ddz, NR, dy, NR, dwa, dwb, dnb, duns, mask ? dyp : nullptr, KO, d1, d2, d3, d4, dout, KO, dsum, 0);
    if (rc) { printf("dgrad R=%d NR=%d KO=%d: launch failed rc=%d\n", R, NR, KO, rc); return 1; }
    CK(cudaDeviceSynchronize());
    std::vector<__nv_bfloat16> out((size_t)R * KO); std::vector<float> st(2 * KO);
    CK(cudaMemcpy(out.data(), dout, out.size() * 2, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(st.data(), dsum, st.size() * 4, cudaMemcpyDeviceToHost));
    double scale = 0; for (float v : ref) scale = fmax(scale, fabs(v));
    double maxerr = 0, maxs = 0, sscale = 0; size_t bad = 0;
    for (size_t i = 0; i < out.size(); ++i) {
        float g = __bfloat162float(out[i]);
        // a value near the ReLU threshold cannot flip (the mask depends on y_prev only); rounding of the folded weights ~ 2^-9
        double e = fabs(g - ref[i]) / (0.2 * scale + fabs(ref[i]));  // bf16 products of rounded folded weights: ~2^-8 of the typical magnitude
        if (!(e < 0.02)) { if (bad < 5) printf("   dz[%zu,%zu] = %f ref %f\n", i / KO, i % KO, g, ref[i]); ++bad; }
        if (e > maxerr) maxerr = e;
    }
    if (mask) {
        for (int k = 0; k < KO; ++k) sscale = fmax(sscale, fmax(fabs(s1[k]), fabs(s2[k])));
        for (int k = 0; k < KO; ++k) {
            maxs = fmax(maxs, fabs(st[k] - s1[k]) / (sscale + 1e-9));
            maxs = fmax(maxs, fabs(st[KO + k] - s2[k]) / (sscale + 1e-9));
        }
    }
    const bool ok = bad == 0 && maxs < 2e-2;
    printf("dgrad R=%6d NR=%4d KO=%4d mask=%d  max rel err %.2e  sums err %.2e  bad %zu  %s\n", R, NR, KO, mask, maxerr, maxs, bad, ok ? "OK" : "FAIL");
    cudaFree(ddz); cudaFree(dwa); cudaFree(dwb); cudaFree(dy); cudaFree(dyp); cudaFree(dnb); cudaFree(d1); cudaFree(d2); cudaFree(d3); cudaFree(d4); cudaFree(dout); cudaFree(dsum);
    return ok ? 0 : 1;
}

static int check_wgrad(int R, int N, int KP, int KT, bool affine) {
    std::vector<__nv_bfloat16> dz((size_t)R * N);
    std::vector<__half> y((size_t)R * N), x((size_t)R * KP);
    std::vector<float> dzf(dz.size()), yf(y.size()), xf(x.size()), cA(N), cB(N), cC(N), sc(KP), sh(KP);
    for (size_t i = 0; i < dz.size(); ++i) { dzf[i] = b16(frand()); dz[i] = __float2bfloat16_rn(dzf[i]); yf[i] = h16(frand() * 2.f); y[i] = __float2half_rn(yf[i]); }
    for (size_t i = 0; i < x.size(); ++i) { xf[i] = h16(frand() * 2.f); x[i] = __float2half_rn(xf[i]); }
    for (int n = 0; n < N; ++n) { cA[n] = 0.5f + 0.4f * frand(); cB[n] = 0.1f * frand(); cC[n] = 0.05f * frand(); }
    for (int k = 0; k < KP; ++k) { sc[k] = 0.5f + 0.4f * frand(); sh[k] = 0.2f * frand(); }
    std::vector<float> dy(dz.size()), xa(x.size());
    for (size_t i = 0; i < dy.size(); ++i) { int n = i % N; dy[i] = b16(fmaf(cA[n], dzf[i], fmaf(cB[n], yf[i], cC[n]))); }
    for (size_t i = 0; i < xa.size(); ++i) { int k = i % KP; xa[i] = b16(affine ? fmaxf(fmaf(xf[i], sc[k], sh[k]), 0.f) : xf[i]); }
    std::vector<double> ref((size_t)N * KT, 0.0);
    for (int r = 0; r < R; ++r)
        for (int n = 0; n < N; ++n) {
            const double d = dy[(size_t)r * N + n];
            for (int k = 0; k < KT; ++k) ref[(size_t)n * KT + k] += d * xa[(size_t)r * KP + k];
        }
    auto *ddz = dev(dz); auto *dy_ = dev(y), *dx = dev(x);
    float *dA = dev(cA), *dB = dev(cB), *dC = dev(cC), *dsc = dev(sc), *dsh = dev(sh);
    float* ddw; CK(cudaMalloc(&ddw, (size_t)N * KT * 4)); CK(cudaMemset(ddw, 0, (size_t)N * KT * 4));
    int rc = pn2_mlp_gemm_wgrad(R, N, KP, KT, ddz, N, dy_, N, dA, dB, dC, dx, KP, affine ? dsc : nullptr, affine ? dsh : nullptr, ddw, KT, 0);
    if (rc) { printf("wgrad R=%d N=%d KP=%d: launch failed rc=%d\n", R, N, KP, rc); return 1; }
    CK(cudaDeviceSynchronize());
    std::vector<float> dw((size_t)N * KT);
    CK(cudaMemcpy(dw.data(), ddw, dw.size() * 4, cudaMemcpyDeviceToHost));
    double scale = 0; for (double v : ref) scale = fmax(scale, fabs(v));
    double maxerr = 0; size_t bad = 0;
    for (size_t i = 0; i < dw.size(); ++i) {
        double e = fabs(dw[i] - ref[i]) / (scale + 1e-9);
        if (!(e < 5e-3)) { if (bad < 5) printf("   dw[%zu,%zu] = %f ref %f\n", i / KT, i % KT, dw[i], ref[i]); ++bad; }
        if (e > maxerr) maxerr = e;
    }
    const bool ok = bad == 0;
    printf("wgrad R=%6d N=%4d KP=%4d KT=%4d affine=%d  max err/scale %.2e (scale %.1f)  bad %zu  %s\n", R, N, KP, KT, affine, maxerr, scale, bad, ok ? "OK" : "FAIL");
    cudaFree(ddz); cudaFree(dy_); cudaFree(dx); cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(dsc); cudaFree(dsh); cudaFree(ddw);
    return ok ? 0 : 1;
}


static void time_wgrad(long long R, int N, int KP, bool affine) {
    void *dz, *y, *x; float *c, *dw, *flush;
    CK(cudaMalloc(&dz, R * N * 2)); CK(cudaMalloc(&y, R * N * 2)); CK(cudaMalloc(&x, R * KP * 2));
    CK(cudaMalloc(&c, 1024 * 4 * 8)); CK(cudaMalloc(&dw, (size_t)N * KP * 4)); CK(cudaMalloc(&flush, 256 << 20));
    CK(cudaMemset(dz, 0, R * N * 2)); CK(cudaMemset(y, 0, R * N * 2)); CK(cudaMemset(x, 0, R * KP * 2)); CK(cudaMemset(c, 0, 1024 * 4 * 8)); CK(cudaMemset(dw, 0, (size_t)N * KP * 4));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9, tot = 0; const int reps = 10;
    for (int i = 0; i < reps + 2; ++i) {
        flush_l2(flush, i);
        cudaEventRecord(e0, 0);
        pn2_mlp_gemm_wgrad(R, N, KP, KP, dz, N, y, N, c, c + 1024, c + 2048, x, KP, affine ? c + 3072 : nullptr, affine ? c + 4096 : nullptr, dw, KP, 0);
        cudaEventRecord(e1, 0); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (i >= 2) { tot += ms; best = fminf(best, ms); }
    }
    const double bytes = (double)R * (2 * N + KP) * 2;
    printf("time wgrad R=%lld N=%d KP=%d affine=%d: avg %.1f us best %.1f us  -> %.0f GB/s (alg %.1f MB)\n", R, N, KP, affine, tot / reps * 1e3, best * 1e3, bytes / (tot / reps * 1e-3) / 1e9, bytes / 1e6);
    cudaFree(dz); cudaFree(y); cudaFree(x); cudaFree(c); cudaFree(dw); cudaFree(flush);
}

static void time_fwd(long long R, int K, int N, bool affine) {
    __half *x, *w, *y; float *sc, *sh, *cen, *st, *flush;
    CK(cudaMalloc(&x, R * K * 2)); CK(cudaMalloc(&w, (size_t)N * K * 2)); CK(cudaMalloc(&y, R * N * 2));
    CK(cudaMalloc(&sc, K * 4)); CK(cudaMalloc(&sh, K * 4)); CK(cudaMalloc(&cen, N * 4)); CK(cudaMalloc(&st, 2 * N * 4));
    CK(cudaMalloc(&flush, 256 << 20));
    CK(cudaMemset(x, 0, R * K * 2)); CK(cudaMemset(w, 0, (size_t)N * K * 2)); CK(cudaMemset(sc, 0, K * 4)); CK(cudaMemset(sh, 0, K * 4)); CK(cudaMemset(cen, 0, N * 4)); CK(cudaMemset(st, 0, 2 * N * 4));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9, tot = 0; const int reps = 10;
    for (int i = 0; i < reps + 2; ++i) {
        flush_l2(flush, i);
        cudaEventRecord(e0, 0);
        pn2_mlp_gemm_fwd(R, K, N, x, K, affine ? sc : nullptr, affine ? sh : nullptr, w, cen, y, N, st, 0);
        cudaEventRecord(e1, 0); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (i >= 2) { tot += ms; best = fminf(best, ms); }
    }
    const double bytes = (double)R * (K + N) * 2;
    printf("time fwd R=%lld K=%d N=%d affine=%d: avg %.1f us best %.1f us  -> %.0f GB/s (alg %.1f MB)\n", R, K, N, affine, tot / reps * 1e3, best * 1e3, bytes / (tot / reps * 1e-3) / 1e9, bytes / 1e6);
    cudaFree(x); cudaFree(w); cudaFree(y); cudaFree(sc); cudaFree(sh); cudaFree(cen); cudaFree(st); cudaFree(flush);
}

static void time_dgrad(long long R, int NR, int KO, bool mask) {
    void *dz, *y, *wt, *wt2, *yp, *out; float *c, *sum, *flush;
    CK(cudaMalloc(&dz, R * NR * 2)); CK(cudaMalloc(&y, R * NR * 2)); CK(cudaMalloc(&wt, (size_t)KO * NR * 2)); CK(cudaMalloc(&wt2, (size_t)KO * NR * 2)); CK(cudaMemset(wt2, 0, (size_t)KO * NR * 2)); CK(cudaMalloc(&yp, R * KO * 2)); CK(cudaMalloc(&out, R * KO * 2));
    CK(cudaMalloc(&c, 1024 * 4 * 8)); CK(cudaMalloc(&sum, 2 * KO * 4)); CK(cudaMalloc(&flush, 256 << 20));
    CK(cudaMemset(dz, 0, R * NR * 2)); CK(cudaMemset(y, 0, R * NR * 2)); CK(cudaMemset(wt, 0, (size_t)KO * NR * 2)); CK(cudaMemset(yp, 0, R * KO * 2)); CK(cudaMemset(c, 0, 1024 * 4 * 8)); CK(cudaMemset(sum, 0, 2 * KO * 4));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9, tot = 0; const int reps = 10;
    for (int i = 0; i < reps + 2; ++i) {
        flush_l2(flush, i);
        cudaEventRecord(e0, 0);
        pn2_mlp_gemm_dgrad(R, NR, KO, dz, NR, y, NR, wt, wt2, c, c + 512, mask ? yp : nullptr, KO, c + 3072, c + 4096, c + 5120, c + 6144, out, KO, sum, 0);
        cudaEventRecord(e1, 0); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (i >= 2) { tot += ms; best = fminf(best, ms); }
    }
    const double bytes = (double)R * (2 * NR + KO + (mask ? KO : 0)) * 2;
    printf("time dgrad R=%lld NR=%d KO=%d mask=%d: avg %.1f us best %.1f us  -> %.0f GB/s (alg %.1f MB)\n", R, NR, KO, mask, tot / reps * 1e3, best * 1e3, bytes / (tot / reps * 1e-3) / 1e9, bytes / 1e6);
    cudaFree(dz); cudaFree(y); cudaFree(wt); cudaFree(yp); cudaFree(out); cudaFree(c); cudaFree(sum); cudaFree(flush);
}

int main(int argc, char** argv) {
    const char* impl = getenv("PN2_GEMM_IMPL");
    printf("PN2_GEMM_IMPL=%s\n", impl ? impl : "(default tc)");
    if (argc > 1 && !strcmp(argv[1], "prof3")) {  // masked dgrad 128 -> 128 (for ncu)
        time_dgrad(131072, 128, 128, true);
        return 0;
    }
    if (argc > 1 && !strcmp(argv[1], "prof4")) {  // a square wgrad, one CTA per SM (for ncu)
        time_wgrad(131072, 128, 128, true);
        return 0;
    }
    if (argc > 1 && !strcmp(argv[1], "prof2")) {  // the smallest wgrad (for ncu)
        time_wgrad(262144, 64, 32, true);
        return 0;
    }
    if (argc > 1 && !strcmp(argv[1], "prof")) {  // the two conv1 GEMMs only (for ncu)
        time_fwd(131072, 128, 384, true);
        time_dgrad(131072, 384, 128, true);
        time_wgrad(131072, 384, 128, true);
        return 0;
    }
    int fails = 0;
    fails += check_fwd(128, 64, 32, false);
    fails += check_fwd(128, 64, 128, false);
    fails += check_fwd(1000, 32, 32, false);
    fails += check_fwd(1000, 128, 128, true);
    fails += check_fwd(777, 64, 64, true);
    fails += check_fwd(4096, 96, 64, false);
    fails += check_fwd(1500, 128, 384, true);
    fails += check_fwd(900, 256, 256, true);
    fails += check_fwd(300, 832, 128, false);
    fails += check_fwd(2049, 128, 512, true);
    fails += check_fwd(20000, 128, 192, true);
    fails += check_dgrad(128, 64, 64, false);
    fails += check_dgrad(1000, 64, 32, true);
    fails += check_dgrad(3000, 128, 96, false);
    fails += check_dgrad(513, 192, 128, true);
    fails += check_dgrad(700, 128, 832, false);
    fails += check_dgrad(2500, 384, 128, true);
    fails += check_dgrad(20000, 256, 256, true);
    fails += check_wgrad(64, 128, 64, 64, false);
    fails += check_wgrad(1000, 32, 32, 3, false);
    fails += check_wgrad(5000, 64, 96, 67, false);
    fails += check_wgrad(999, 128, 128, 128, true);
    fails += check_wgrad(3000, 384, 128, 128, true);
    fails += check_wgrad(2000, 128, 832, 771, false);
    fails += check_wgrad(1500, 512, 128, 128, true);
    fails += check_wgrad(1700, 256, 640, 640, false);
    fails += check_wgrad(2100, 192, 128, 128, true);
    printf("%d failing case(s)\n", fails);
    if (argc > 1 && !strcmp(argv[1], "time")) {
        time_fwd(131072, 128, 384, true);   // conv1
        time_fwd(131072, 192, 128, false);  // fp1 layer 0
        time_fwd(131072, 128, 128, true);   // fp1 layer 1
        time_fwd(262144, 32, 32, false);    // sa1 layer 0
        time_fwd(262144, 32, 64, true);     // sa1 layer 2
        time_fwd(131072, 64, 128, true);    // sa2 layer 2
        time_fwd(43008, 832, 128, false);   // q2 K=64 layer 0
        time_dgrad(131072, 384, 128, true); // conv1 -> fp1
        time_dgrad(131072, 128, 128, true);
        time_dgrad(131072, 128, 192, false);
        time_dgrad(43008, 128, 832, false);
        time_dgrad(262144, 64, 32, true);
        time_wgrad(131072, 384, 128, true);   // conv1
        time_wgrad(131072, 128, 128, true);   // fp1 layer 1
        time_wgrad(131072, 128, 192, false);  // fp1 layer 0
        time_wgrad(262144, 64, 32, true);     // sa1 layer 2
        time_wgrad(43008, 128, 832, false);   // q2 K=64 layer 0
        time_wgrad(43008, 192, 128, true);    // q2 K=64 layer 2
    }
    return fails ? 1 : 0;
}
