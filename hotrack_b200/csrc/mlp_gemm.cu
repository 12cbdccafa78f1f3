// mlp_gemm.cu -- the grouped per-point MLP as 16-bit tensor-core GEMMs over row matrices, sm_100a.
//
// Replaces the per-layer  relu(bn(conv1x1(x)))  chain of the reference's SA / FP modules
// (network/models/pointnet_utils.py:399-403,458-460,505-507,577-580; backbones.py:131-132), which
// runs as conv (cuDNN) + BatchNorm (cuDNN) + ReLU (elementwise) launches, each materialising a
// (B,C,S,K) fp32 tensor, and 5+ more launches per layer in backward.
//
// Layout: activations are ROW matrices  X[R][C]  (R = B*S*K grouped rows or B*N points; channels
// contiguous; 16-bit: fp16 forward, bf16 gradients -- see mma_common.cuh; leading dimension a multiple of 8).  Only the PRE-BatchNorm conv output Y_l of each
// layer is ever stored; BatchNorm + ReLU are applied by the CONSUMER while it stages its A operand
// ("prologue"), and the batch statistics BatchNorm needs are column sums produced by the PRODUCER's
// epilogue.  Per layer that is one read of Y_{l-1} and one write of Y_l (2 B/element each) instead
// of five fp32 round trips.
//
//   gemm_rows_kernel<BN, AMODE, MASK>      C[R][n] = A'[R][k] * B[n][k]^T
//      AMODE PLAIN   A' = A                               (first layer: gathered rows)
//            AFFINE  A' = relu(A*scale + shift)           (forward: BN+ReLU of the previous layer)
//            BNBWD   A' = cA*dZ + cB*Y + cC               (backward: BatchNorm-backward of this layer)
//      epilogue: bf16 tile staged in shared memory, written with coalesced 16-byte stores; column sums
//      sum(v), sum(v*v)  (forward BN statistics)  or, with MASK (backward):  v *= [prev act > 0],
//      sum(v), sum(v*xhat_prev)  -- the two reductions BatchNorm-backward of the previous layer needs.
//   wgrad_kernel<AFFINE>                   dW[n][k] += sum_r dY[r][n] * X'[r][k]   (split over rows,
//      transposed ldmatrix fragments, fp32 atomics into the zeroed gradient).
//
// Centering: a pre-BatchNorm channel can have |mean| >> std (e.g. FP3, whose input is dominated by
// the broadcast global feature), and bf16 rounding of such values destroys what BatchNorm keeps.
// The forward GEMM therefore stores  y - c[n]  with c an estimate of the channel mean (center_kernel:
// W * mean of a few sampled input rows); BatchNorm is shift-invariant, so only the running mean and
// the eval-mode shift need c added back (bn_finalize / bn_eval_affine).
//
// STATUS: the production forward / input-gradient GEMM is mlp_gemm_tc.cu (tcgen05).  gemm_rows_kernel below is kept as
// the forward cross-check (PN2_GEMM_IMPL=mma; tools/dev/gemm_tc_check.cu, tests); its BNBWD mode is no longer
// dispatched (the input gradient now takes coefficient-folded weights, see bn_bwd_coefs_kernel).  wgrad_kernel, the
// per-layer constant kernels and every extern "C" entry point of include/pn2b200_mlp.h's GEMM section live here.
//
// Machine mapping: persistent CTAs (<= 2 per SM) walk 128-row tiles; A is register-prefetched one
// chunk ahead (so the prologue runs once per element, not once per consuming warp), B (weights,
// L2-resident) streams through cp.async; mma.sync.m16n8k16 bf16 with fp32 accumulation.  Every layer
// of this network is far below the tensor roofline (K <= 800, N <= 512): the bound is the HBM
// traffic of the row matrices, which is what the design minimises.
#include "mlp_gemm.cuh"

namespace pn2 {
namespace {

constexpr int BM = 128;
constexpr int kThreads = 256;

__device__ __forceinline__ uint4 ldg128(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void cp_async_wait_ring(int nst) {  // all but the nst-2 most recent groups
    switch (nst) {
        case 2: cp_async_wait<0>(); break;
        case 3: cp_async_wait<1>(); break;
        case 4: cp_async_wait<2>(); break;
        case 5: cp_async_wait<3>(); break;
        default: cp_async_wait<4>(); break;
    }
}

// BKT = K-chunk width (32 or 64 elements: 64 / 128 bytes of every row per chunk).  p.nst = ring depth and
// p.bres = "B slab resident in shared memory" are chosen by the launcher from the shared-memory budget.
// 16 warps per CTA (8 for the 32-column tile): the kernel is one CTA per SM (shared-memory ring), so the
// warp count is what hides instruction and shared-memory latency.
__host__ __device__ constexpr int gemm_threads(int bn) { return bn == 32 ? 256 : 512; }

template <int BN, int BKT, int AMODE, bool MASK>
__global__ void __launch_bounds__(gemm_threads(BN), 1) gemm_rows_kernel(const GemmArgs p) {
    constexpr int kThreads = gemm_threads(BN);
    constexpr int NW = kThreads / 32;
    constexpr int WN = BN / 32, WM = NW / WN, MF = BM / WM / 16;
    static_assert(MF >= 1, "warp tile must hold at least one 16-row fragment");
    constexpr int CLD = BN + 8;
    constexpr int CPR = BN / 8;
    constexpr int RPP = kThreads / CPR;
    constexpr int PASSES = BM / RPP;
    constexpr int NCOEF = AMODE == A_PLAIN ? 0 : (AMODE == A_AFFINE ? 2 : 3);
    constexpr bool FWD = AMODE != A_BNBWD;  // forward GEMMs compute and store fp16, backward ones bf16
    constexpr int LDK = BKT + 8;            // shared-memory row stride of a chunk (odd multiple of 16 bytes)
    constexpr int PPR = BKT / 8;            // 16-byte pieces per row per chunk
    constexpr int APT = BM * PPR / kThreads;
    constexpr int ASZ = BM * LDK, BSZ = BN * LDK;

    const int nst = p.nst;
    const bool bres = p.bres != 0;
    const int ldb = bres ? p.kdim + 8 : LDK;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint16_t* sA = reinterpret_cast<uint16_t*>(smem_raw);               // [nst][BM][LDK]
    uint16_t* sA1 = sA + nst * ASZ;                                      // [nst][BM][LDK], BNBWD only
    uint16_t* sB = sA1 + (AMODE == A_BNBWD ? nst * ASZ : 0);             // [nst][BN][LDK] or [BN][kdim+8]
    uint16_t* sC = sB + (bres ? BN * (p.kdim + 8) : nst * BSZ);          // [BM][CLD]
    float* sCoef = reinterpret_cast<float*>(sC + BM * CLD);
    float* sPrev = sCoef + NCOEF * p.kdim;      // [4][BN], MASK only
    float* sCen = sPrev + (MASK ? 4 * BN : 0);  // [BN]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp / WN, wn = warp % WN;
    const int g = lane >> 2, t4 = lane & 3;
    const int n0 = blockIdx.y * BN;
    const int KT = p.kdim / BKT;
    const long long tiles = (p.rows + BM - 1) / BM;

    if (bres) {  // the whole weight slab of this column tile, once per CTA (oldest cp.async group)
        const int ppr = p.kdim >> 3;
        for (int i = tid; i < BN * ppr; i += kThreads) {
            const int r = i / ppr, ch = i - r * ppr;
            const bool ok = n0 + r < p.n;
            cp_async16(&sB[r * ldb + ch * 8], p.b + (size_t)(ok ? n0 + r : 0) * p.kdim + ch * 8, ok ? 16 : 0);
        }
        cp_async_commit();
    }
    for (int i = tid; i < NCOEF * p.kdim; i += kThreads) {
        const int which = i / p.kdim, c = i - which * p.kdim;
        const float* src = which == 0 ? p.c0 : (which == 1 ? p.c1 : p.c2);
        sCoef[i] = src[c];
    }
    for (int i = tid; i < BN; i += kThreads) sCen[i] = (p.center && n0 + i < p.n) ? p.center[n0 + i] : 0.f;
    if (MASK) {
        for (int i = tid; i < 4 * BN; i += kThreads) {
            const int which = i / BN, c = n0 + (i - which * BN);
            const float* src = which == 0 ? p.p_scale : (which == 1 ? p.p_shift : (which == 2 ? p.p_mean : p.p_rstd));
            sPrev[i] = c < p.n ? src[c] : 0.f;
        }
    }
    __syncthreads();

    // ---- operand pipeline: ring of raw chunks filled by cp.async, deep enough to keep >= 50 KB of HBM
    // loads in flight per SM; A chunks are transformed IN PLACE by the thread that copied them
    const long long my_tiles = blockIdx.x < tiles ? (tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const long long total = my_tiles * KT;

    // issue-side cursor (chunk index, tile, K chunk, ring slot), advanced incrementally: no divisions in the loop
    long long is_it = 0, is_tile = blockIdx.x;
    int is_kc = 0, is_st = 0;
    auto issue = [&]() {
        if (is_it < total) {
            const long long tile = is_tile;
            const int kc = is_kc, st = is_st;
#pragma unroll
            for (int j = 0; j < APT; ++j) {
                const int q = tid + j * kThreads;
                const int r = q / PPR, col = (q % PPR) * 8;
                const long long row = tile * BM + r;
                const bool ok = row < p.rows;
                const long long rr = ok ? row : 0;
                cp_async16(&sA[st * ASZ + r * LDK + col], p.a0 + rr * p.a0_ld + kc * BKT + col, ok ? 16 : 0);
                if (AMODE == A_BNBWD)
                    cp_async16(&sA1[st * ASZ + r * LDK + col], p.a1 + rr * p.a1_ld + kc * BKT + col, ok ? 16 : 0);
            }
            if (!bres) {
                for (int i = tid; i < BN * PPR; i += kThreads) {
                    const int r = i / PPR, ch = i % PPR;
                    const bool ok = n0 + r < p.n;
                    cp_async16(&sB[st * BSZ + r * LDK + ch * 8],
                               p.b + (size_t)(ok ? n0 + r : 0) * p.kdim + kc * BKT + ch * 8, ok ? 16 : 0);
                }
            }
            ++is_it;
            if (++is_kc == KT) { is_kc = 0; is_tile += gridDim.x; }
            if (++is_st == nst) is_st = 0;
        }
        cp_async_commit();  // always: keeps the group count in step with the iteration count
    };
    auto transform_A = [&](long long tile, int kc, int st) {
        if (AMODE == A_PLAIN) return;
#pragma unroll
        for (int j = 0; j < APT; ++j) {
            const int q = tid + j * kThreads;
            const int r = q / PPR, col = (q % PPR) * 8;
            const int c = kc * BKT + col;
            uint4* slot = reinterpret_cast<uint4*>(&sA[st * ASZ + r * LDK + col]);
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (tile * BM + r < p.rows) {
                const uint4 q0 = *slot;
                const uint32_t* x0 = reinterpret_cast<const uint32_t*>(&q0);
                uint32_t* o = reinterpret_cast<uint32_t*>(&v);
                const float4 ka0 = *reinterpret_cast<const float4*>(&sCoef[c]);
                const float4 ka1 = *reinterpret_cast<const float4*>(&sCoef[c + 4]);
                const float4 kb0 = *reinterpret_cast<const float4*>(&sCoef[p.kdim + c]);
                const float4 kb1 = *reinterpret_cast<const float4*>(&sCoef[p.kdim + c + 4]);
                const float k0[8] = {ka0.x, ka0.y, ka0.z, ka0.w, ka1.x, ka1.y, ka1.z, ka1.w};
                const float k1[8] = {kb0.x, kb0.y, kb0.z, kb0.w, kb1.x, kb1.y, kb1.z, kb1.w};
                if (AMODE == A_AFFINE) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float2 a = h2_to_f2(x0[e]);
                        o[e] = f2_to_h2(fmaxf(fmaf(a.x, k0[2 * e], k1[2 * e]), 0.f),
                                        fmaxf(fmaf(a.y, k0[2 * e + 1], k1[2 * e + 1]), 0.f));
                    }
                } else {
                    const uint4 q1 = *reinterpret_cast<const uint4*>(&sA1[st * ASZ + r * LDK + col]);
                    const uint32_t* x1 = reinterpret_cast<const uint32_t*>(&q1);
                    const float4 kc0 = *reinterpret_cast<const float4*>(&sCoef[2 * p.kdim + c]);
                    const float4 kc1 = *reinterpret_cast<const float4*>(&sCoef[2 * p.kdim + c + 4]);
                    const float k2[8] = {kc0.x, kc0.y, kc0.z, kc0.w, kc1.x, kc1.y, kc1.z, kc1.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float2 a = bf2_to_f2(x0[e]);
                        const float2 y = h2_to_f2(x1[e]);
                        o[e] = f2_to_bf2(fmaf(k0[2 * e], a.x, fmaf(k1[2 * e], y.x, k2[2 * e])),
                                         fmaf(k0[2 * e + 1], a.y, fmaf(k1[2 * e + 1], y.y, k2[2 * e + 1])));
                    }
                }
            }
            *slot = v;
        }
    };

    float acc[MF][4][4];
    float s1[8], s2[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) s1[e] = s2[e] = 0.f;
    const int chunk = tid % CPR;
    const int col0 = n0 + chunk * 8;
    uint4 yq[MASK ? PASSES : 1];  // the previous layer's y pieces this thread masks with, fetched at tile start

    for (int s0 = 0; s0 < nst - 1; ++s0) issue();
    long long tile = blockIdx.x;
    int kc = 0, st = 0;
    for (long long it = 0; it < total; ++it) {
        cp_async_wait_ring(nst);  // this thread's copies of chunk `it` (and the resident B slab) have landed
        transform_A(tile, kc, st);
        __syncthreads();  // chunk `it` complete for everyone; everyone is done reading chunk it-1
        issue();          // chunk it+nst-1 refills the buffer chunk it-1 used
        if (kc == 0) {
#pragma unroll
            for (int mf = 0; mf < MF; ++mf)
#pragma unroll
                for (int nf = 0; nf < 4; ++nf)
#pragma unroll
                    for (int e = 0; e < 4; ++e) acc[mf][nf][e] = 0.f;
            if (MASK && col0 < p.n) {
#pragma unroll
                for (int ps = 0; ps < PASSES; ++ps) {
                    const long long grow = tile * BM + tid / CPR + ps * RPP;
                    if (grow < p.rows) yq[ps] = ldg128(p.yp + grow * p.yp_ld + col0);
                }
            }
        }
        const uint16_t* aBase = sA + st * ASZ;
        const uint16_t* bBase = bres ? sB + kc * BKT : sB + st * BSZ;
#pragma unroll
        for (int ks = 0; ks < BKT / 16; ++ks) {
            uint32_t af[MF][4], bfr[2][4];
#pragma unroll
            for (int mf = 0; mf < MF; ++mf)
                ldsm_x4(af[mf], smem_u32(&aBase[(wm * (MF * 16) + mf * 16 + (lane & 15)) * LDK + ks * 16 + (lane >> 4) * 8]));
#pragma unroll
            for (int nb = 0; nb < 2; ++nb)
                ldsm_x4(bfr[nb], smem_u32(&bBase[(wn * 32 + nb * 16 + (lane & 7) + ((lane >> 4) << 3)) * ldb + ks * 16 +
                                                 ((lane >> 3) & 1) * 8]));
#pragma unroll
            for (int mf = 0; mf < MF; ++mf)
#pragma unroll
                for (int nf = 0; nf < 4; ++nf) {
                    if (FWD)
                        mma_f16_16816(acc[mf][nf], af[mf], bfr[nf >> 1][(nf & 1) * 2], bfr[nf >> 1][(nf & 1) * 2 + 1]);
                    else
                        mma_bf16_16816(acc[mf][nf], af[mf], bfr[nf >> 1][(nf & 1) * 2], bfr[nf >> 1][(nf & 1) * 2 + 1]);
                }
        }
        if (kc == KT - 1) {
            // ---- epilogue: fp32 accumulators -> 16-bit tile in shared memory -> coalesced stores + column sums
#pragma unroll
            for (int mf = 0; mf < MF; ++mf)
#pragma unroll
                for (int nf = 0; nf < 4; ++nf) {
                    const int r = wm * (MF * 16) + mf * 16 + g, c = wn * 32 + nf * 8 + t4 * 2;
                    const float ce0 = sCen[c], ce1 = sCen[c + 1];
                    const float v0 = acc[mf][nf][0] - ce0, v1 = acc[mf][nf][1] - ce1;
                    const float v2 = acc[mf][nf][2] - ce0, v3 = acc[mf][nf][3] - ce1;
                    *reinterpret_cast<uint32_t*>(&sC[r * CLD + c]) = FWD ? f2_to_h2(v0, v1) : f2_to_bf2(v0, v1);
                    *reinterpret_cast<uint32_t*>(&sC[(r + 8) * CLD + c]) = FWD ? f2_to_h2(v2, v3) : f2_to_bf2(v2, v3);
                }
            __syncthreads();
            if (col0 < p.n) {
                float ps_[8], ph_[8], pm_[8], pr_[8];
                if (MASK) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const float4 a4 = *reinterpret_cast<const float4*>(&sPrev[chunk * 8 + 4 * h]);
                        const float4 b4 = *reinterpret_cast<const float4*>(&sPrev[BN + chunk * 8 + 4 * h]);
                        const float4 c4 = *reinterpret_cast<const float4*>(&sPrev[2 * BN + chunk * 8 + 4 * h]);
                        const float4 d4 = *reinterpret_cast<const float4*>(&sPrev[3 * BN + chunk * 8 + 4 * h]);
                        ps_[4 * h] = a4.x; ps_[4 * h + 1] = a4.y; ps_[4 * h + 2] = a4.z; ps_[4 * h + 3] = a4.w;
                        ph_[4 * h] = b4.x; ph_[4 * h + 1] = b4.y; ph_[4 * h + 2] = b4.z; ph_[4 * h + 3] = b4.w;
                        pm_[4 * h] = c4.x; pm_[4 * h + 1] = c4.y; pm_[4 * h + 2] = c4.z; pm_[4 * h + 3] = c4.w;
                        pr_[4 * h] = d4.x; pr_[4 * h + 1] = d4.y; pr_[4 * h + 2] = d4.z; pr_[4 * h + 3] = d4.w;
                    }
                }
#pragma unroll
                for (int ps = 0; ps < PASSES; ++ps) {
                    const int r = tid / CPR + ps * RPP;
                    const long long grow = tile * BM + r;
                    if (grow < p.rows) {
                        uint4 v = *reinterpret_cast<const uint4*>(&sC[r * CLD + chunk * 8]);
                        uint32_t* vv = reinterpret_cast<uint32_t*>(&v);
                        if (MASK) {
                            const uint32_t* yy = reinterpret_cast<const uint32_t*>(&yq[ps]);
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                float2 d = bf2_to_f2(vv[e]);  // MASK is a backward epilogue: bf16 tile
                                const float2 y = h2_to_f2(yy[e]);
                                const float a0 = fmaf(y.x, ps_[2 * e], ph_[2 * e]);
                                const float a1 = fmaf(y.y, ps_[2 * e + 1], ph_[2 * e + 1]);
                                d.x = a0 > 0.f ? d.x : 0.f;
                                d.y = a1 > 0.f ? d.y : 0.f;
                                const float h0 = (y.x - pm_[2 * e]) * pr_[2 * e];
                                const float h1 = (y.y - pm_[2 * e + 1]) * pr_[2 * e + 1];
                                s1[2 * e] += d.x; s1[2 * e + 1] += d.y;
                                s2[2 * e] = fmaf(d.x, h0, s2[2 * e]);
                                s2[2 * e + 1] = fmaf(d.y, h1, s2[2 * e + 1]);
                                vv[e] = f2_to_bf2(d.x, d.y);
                            }
                        } else if (p.sums) {
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float2 d = FWD ? h2_to_f2(vv[e]) : bf2_to_f2(vv[e]);
                                s1[2 * e] += d.x; s1[2 * e + 1] += d.y;
                                s2[2 * e] = fmaf(d.x, d.x, s2[2 * e]);
                                s2[2 * e + 1] = fmaf(d.y, d.y, s2[2 * e + 1]);
                            }
                        }
                        *reinterpret_cast<uint4*>(p.out + grow * p.out_ld + col0) = v;
                    }
                }
            }
            // sC is next written after the main-loop barrier of the following iteration
        }
        if (++kc == KT) { kc = 0; tile += gridDim.x; }
        if (++st == nst) st = 0;
    }
    cp_async_wait<0>();

    if (p.sums) {
        // column sums: lanes sharing a chunk within the warp, then the 8 warps through shared memory
        __syncthreads();
        float* red = reinterpret_cast<float*>(sC);  // [NW warps][2][BN]
#pragma unroll
        for (int e = 0; e < 8; ++e) {
#pragma unroll
            for (int o = CPR; o < 32; o <<= 1) {
                s1[e] += __shfl_xor_sync(kFull, s1[e], o);
                s2[e] += __shfl_xor_sync(kFull, s2[e], o);
            }
        }
        if (CPR >= 32 || lane < CPR) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                red[(warp * 2 + 0) * BN + chunk * 8 + e] = s1[e];
                red[(warp * 2 + 1) * BN + chunk * 8 + e] = s2[e];
            }
        }
        __syncthreads();
        for (int i = tid; i < 2 * BN; i += kThreads) {
            const int which = i / BN, c = i - which * BN;
            if (n0 + c < p.n) {
                float s = 0.f;
#pragma unroll
                for (int w = 0; w < NW; ++w) s += red[(w * 2 + which) * BN + c];
                atomicAdd(p.sums + (size_t)which * p.n + n0 + c, s);
            }
        }
    }
}

template <int BN, int BKT, int AMODE, bool MASK>
int launch_gemm(GemmArgs a, cudaStream_t stream) {
    constexpr int NCOEF = AMODE == A_PLAIN ? 0 : (AMODE == A_AFFINE ? 2 : 3);
    constexpr size_t kMaxSmem = 220 * 1024;
    constexpr int LDK = BKT + 8;
    const size_t fixed = (size_t)BM * (BN + 8) * 2 + (size_t)(NCOEF * a.kdim + BN + (MASK ? 4 * BN : 0)) * 4;
    const size_t bres_bytes = (size_t)BN * (a.kdim + 8) * 2;
    const size_t a_stage = (size_t)BM * LDK * 2 * (AMODE == A_BNBWD ? 2 : 1);
    const size_t b_stage = (size_t)BN * LDK * 2;
    // keep the weight slab resident when that still leaves room for >= 3 operand stages
    a.bres = (bres_bytes <= 72 * 1024 && fixed + bres_bytes + 3 * a_stage <= kMaxSmem) ? 1 : 0;
    const size_t stage = a_stage + (a.bres ? 0 : b_stage);
    const size_t avail = kMaxSmem - fixed - (a.bres ? bres_bytes : 0);
    if (fixed + (a.bres ? bres_bytes : 0) > kMaxSmem || avail < 2 * stage)
        return fail_arg("pn2_mlp_gemm", "reduction dimension too large for shared memory");
    long long nst = (long long)(avail / stage);
    if (nst > 6) nst = 6;
    a.nst = (int)nst;
    const size_t smem = fixed + (a.bres ? bres_bytes : 0) + nst * stage;
    static DeviceOnce once;
    if (once.first()) {
        PN2_CHECK(cudaFuncSetAttribute(gemm_rows_kernel<BN, BKT, AMODE, MASK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)kMaxSmem),
                  "gemm: cudaFuncSetAttribute");
    }
    const long long tiles = (a.rows + BM - 1) / BM;
    const int ny = (a.n + BN - 1) / BN;
    const int sms = sm_count();
    long long gx = sms / ny;  // persistent, one CTA per SM: every CTA resident, tiles dealt round-robin
    if (gx < 1) gx = 1;
    if (gx > tiles) gx = tiles;
    dim3 grid((unsigned)gx, ny);
    gemm_rows_kernel<BN, BKT, AMODE, MASK><<<grid, gemm_threads(BN), smem, stream>>>(a);
    PN2_CHECK_LAUNCH("gemm_rows_kernel");
    return 0;
}

template <int AMODE, bool MASK>
int dispatch_bn(const GemmArgs& a, cudaStream_t stream) {
    if (a.kdim % 64 == 0) {
        if (a.n <= 32) return launch_gemm<32, 64, AMODE, MASK>(a, stream);
        if (a.n <= 64) return launch_gemm<64, 64, AMODE, MASK>(a, stream);
        return launch_gemm<128, 64, AMODE, MASK>(a, stream);
    }
    if (a.n <= 32) return launch_gemm<32, 32, AMODE, MASK>(a, stream);
    if (a.n <= 64) return launch_gemm<64, 32, AMODE, MASK>(a, stream);
    return launch_gemm<128, 32, AMODE, MASK>(a, stream);
}

int check_common(const char* who, long long rows, int kdim, int n) {
    if (rows < 0 || kdim <= 0 || n <= 0) return fail_arg(who, "non-positive size");
    if (kdim % 32 != 0) return fail_arg(who, "reduction dimension must be a multiple of 32");
    if (n % 8 != 0) return fail_arg(who, "output columns must be a multiple of 8");
    return 0;
}

// ------------------------------------------------------------------ wgrad -------------------
constexpr int TN = 64, TK = 128;            // CTA tile: 64 output channels x 128 input channels
constexpr int SLDN = TN + 8, SLD = TK + 8;   // shared-memory row strides (odd multiples of 16 bytes: conflict-free ldmatrix)

// Stage = WBR rows of dz / y (64 columns; dz -> dY in place) and x (128 columns; -> X' in place), brought in by
// cp.async WST stages deep.  8 warps and ~108 KB per CTA: TWO CTAs per SM, so one CTA's transform / barrier phase
// overlaps the other's ldmatrix + mma phase (the kernel is latency-bound, not issue- or bandwidth-bound).
constexpr int WBR = 64, WST = 3;
constexpr int kWgradStage = WBR * (2 * SLDN + SLD);  // 16-bit elements
constexpr int kWgradSmem = WST * kWgradStage * 2 + 5 * 128 * 4;

constexpr int kWgradThreads = 256;  // 8 warps: 2 (Cout) x 4 (Cin) warp tiles of 32 x 32

template <bool AFFINE>
__global__ void __launch_bounds__(kWgradThreads, 2) wgrad_kernel(const WgradArgs p) {
    constexpr int kThreads = kWgradThreads;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint16_t* sD = reinterpret_cast<uint16_t*>(smem_raw);  // [WST][WBR][SLDN] dz -> dY (bf16)
    uint16_t* sY = sD + WST * WBR * SLDN;                  // [WST][WBR][SLDN] y (fp16)
    uint16_t* sX = sY + WST * WBR * SLDN;                  // [WST][WBR][SLD]  x (fp16) -> X' (bf16)
    float* sCo = reinterpret_cast<float*>(sX + WST * WBR * SLD);  // [5][128]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n0 = blockIdx.x * TN, k0 = blockIdx.y * TK;
    for (int i = tid; i < 5 * 128; i += kThreads) {
        const int which = i >> 7, c = i & 127;
        float v = 0.f;
        if (which < 3) {
            if (c < TN && n0 + c < p.n) v = (which == 0 ? p.cA : (which == 1 ? p.cB : p.cC))[n0 + c];
        } else if (AFFINE) {
            if (k0 + c < p.kp) v = (which == 3 ? p.in_scale : p.in_shift)[k0 + c];
        }
        sCo[i] = v;
    }
    __syncthreads();

    const long long chunks = (p.rows + WBR - 1) / WBR;
    const long long mine = blockIdx.z < chunks ? (chunks - blockIdx.z + gridDim.z - 1) / gridDim.z : 0;
    const int d_row = tid >> 3, d_col = (tid & 7) * 8;    // dz / y: 2 pieces per thread, rows d_row + 32*j
    const int x_row = tid >> 4, x_col = (tid & 15) * 8;   // x: 4 pieces per thread, rows x_row + 16*j
    const bool ncol_ok = n0 + d_col < p.n, kcol_ok = k0 + x_col < p.kp;

    long long is_i = 0, is_ch = blockIdx.z;
    int is_st = 0;
    auto issue = [&]() {
        if (is_i < mine) {
            const long long ch = is_ch;
            const int st = is_st;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int r = d_row + j * 32;
                const long long row = ch * WBR + r;
                const bool ok = row < p.rows;
                const long long rr = ok ? row : 0;
                const int off = (st * WBR + r) * SLDN + d_col;
                cp_async16(&sD[off], p.dz + rr * p.dz_ld + (ncol_ok ? n0 + d_col : 0), ok && ncol_ok ? 16 : 0);
                cp_async16(&sY[off], p.y + rr * p.y_ld + (ncol_ok ? n0 + d_col : 0), ok && ncol_ok ? 16 : 0);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int r = x_row + j * 16;
                const long long row = ch * WBR + r;
                const bool ok = row < p.rows;
                const long long rr = ok ? row : 0;
                cp_async16(&sX[(st * WBR + r) * SLD + x_col], p.x + rr * p.x_ld + (kcol_ok ? k0 + x_col : 0),
                           ok && kcol_ok ? 16 : 0);
            }
            ++is_i;
            is_ch += gridDim.z;
            if (++is_st == WST) is_st = 0;
        }
        cp_async_commit();
    };
    auto transform = [&](long long ch, int st) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int r = d_row + j * 32;
            const bool ok = ch * WBR + r < p.rows;
            const int off = (st * WBR + r) * SLDN + d_col;
            uint4 vd = make_uint4(0u, 0u, 0u, 0u);
            if (ok) {
                const uint4 qd = *reinterpret_cast<const uint4*>(&sD[off]);
                const uint4 qy = *reinterpret_cast<const uint4*>(&sY[off]);
                const uint32_t* a = reinterpret_cast<const uint32_t*>(&qd);
                const uint32_t* b = reinterpret_cast<const uint32_t*>(&qy);
                uint32_t* od = reinterpret_cast<uint32_t*>(&vd);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int c = d_col + 2 * e;
                    const float2 d = bf2_to_f2(a[e]), y = h2_to_f2(b[e]);
                    // columns beyond n have zero coefficients and zero-filled data: they stay 0
                    od[e] = f2_to_bf2(fmaf(sCo[c], d.x, fmaf(sCo[128 + c], y.x, sCo[256 + c])),
                                      fmaf(sCo[c + 1], d.y, fmaf(sCo[128 + c + 1], y.y, sCo[256 + c + 1])));
                }
            }
            *reinterpret_cast<uint4*>(&sD[off]) = vd;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int r = x_row + j * 16;
            const bool ok = ch * WBR + r < p.rows;
            const int off = (st * WBR + r) * SLD + x_col;
            uint4 vx = make_uint4(0u, 0u, 0u, 0u);
            if (ok) {
                const uint4 qx = *reinterpret_cast<const uint4*>(&sX[off]);
                const uint32_t* x = reinterpret_cast<const uint32_t*>(&qx);
                uint32_t* ox = reinterpret_cast<uint32_t*>(&vx);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int c = x_col + 2 * e;
                    const float2 xv = h2_to_f2(x[e]);
                    if (AFFINE)
                        ox[e] = f2_to_bf2(fmaxf(fmaf(xv.x, sCo[384 + c], sCo[512 + c]), 0.f),
                                          fmaxf(fmaf(xv.y, sCo[384 + c + 1], sCo[512 + c + 1]), 0.f));
                    else
                        ox[e] = f2_to_bf2(xv.x, xv.y);  // the gradient GEMM runs in bf16
                }
            }
            *reinterpret_cast<uint4*>(&sX[off]) = vx;
        }
    };

    const int wn2 = warp >> 2, wk = warp & 3;  // 2 x 4 warps
    const bool warp_on = (n0 + wn2 * 32 < p.n) && (k0 + wk * 32 < p.kp);
    float acc[2][4][4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[a][b][e] = 0.f;

    const int mi = lane >> 3, l7 = lane & 7;
    for (int s0 = 0; s0 < WST - 1; ++s0) issue();
    long long ch = blockIdx.z;
    int st = 0;
    for (long long i = 0; i < mine; ++i, ch += gridDim.z, st = (st + 1 == WST ? 0 : st + 1)) {
        asm volatile("cp.async.wait_group %0;" ::"n"(WST - 2) : "memory");
        transform(ch, st);
        __syncthreads();
        issue();
        if (warp_on) {
#pragma unroll
            for (int rs = 0; rs < WBR; rs += 16) {
                uint32_t af[2][4], bfr[2][4];
#pragma unroll
                for (int mf = 0; mf < 2; ++mf)
                    ldsm_x4_trans(af[mf], smem_u32(&sD[(st * WBR + rs + (mi >> 1) * 8 + l7) * SLDN + wn2 * 32 + mf * 16 +
                                                       (mi & 1) * 8]));
#pragma unroll
                for (int nb = 0; nb < 2; ++nb)
                    ldsm_x4_trans(bfr[nb], smem_u32(&sX[(st * WBR + rs + (mi & 1) * 8 + l7) * SLD + wk * 32 + nb * 16 +
                                                        (mi >> 1) * 8]));
#pragma unroll
                for (int mf = 0; mf < 2; ++mf) {
                    if (n0 + wn2 * 32 + mf * 16 < p.n) {
#pragma unroll
                        for (int nf = 0; nf < 4; ++nf)
                            mma_bf16_16816(acc[mf][nf], af[mf], bfr[nf >> 1][(nf & 1) * 2], bfr[nf >> 1][(nf & 1) * 2 + 1]);
                    }
                }
            }
        }
    }
    if (warp_on) {
        const int g = lane >> 2, t4 = lane & 3;
#pragma unroll
        for (int mf = 0; mf < 2; ++mf)
#pragma unroll
            for (int nf = 0; nf < 4; ++nf)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int n = n0 + wn2 * 32 + mf * 16 + g + (e >> 1) * 8;
                    const int k = k0 + wk * 32 + nf * 8 + t4 * 2 + (e & 1);
                    if (n < p.n && k < p.k_true) atomicAdd(p.dw + (size_t)n * p.dw_ld + k, acc[mf][nf][e]);
                }
    }
}

// ------------------------------------------------------------------ centre estimate ----------
// c[n] = sum_k w[n][k] * mean_j act(x[row_j][k]) over <= 16 rows spread evenly over the matrix;
// c_true[n] = c[n] + sum_k w[n][k] * in_offset[k]: the input rows may themselves be stored centred
// (x_true = x + in_offset per channel), which the GEMM output inherits as a per-channel constant.
// One warp per output column (8 columns per CTA) so that even 512 columns spread over 64 CTAs.
__global__ void __launch_bounds__(256) center_kernel(long long rows, int kdim, int n, const act_t* __restrict__ x,
                                                      int x_ld, const float* __restrict__ sc,
                                                      const float* __restrict__ sh, const act_t* __restrict__ w,
                                                      const float* __restrict__ off0, float off0_scale, int off0_start,
                                                      int off0_count, const float* __restrict__ off1, float off1_scale,
                                                      int off1_start, int off1_count, float* __restrict__ c,
                                                      float* __restrict__ c_true) {
    pdl_enter();  // programmatic dependent launch (pn2_common.cuh): first statement, before any memory access
    __shared__ float sM[1024];
    __shared__ float sO[1024];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ns = rows < 16 ? (int)rows : 16;
    const long long step = rows / ns;
    const float inv = 1.f / (float)ns;
    // this warp's weight row first: its loads do not depend on the sampled rows, so the two memory round trips of
    // the kernel overlap (the kernel is pure latency: a few KB per CTA)
    const int col = blockIdx.x * 8 + warp;
    uint32_t wreg[16];  // kdim <= 1024: 2 x 16 values per lane
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int k = lane * 2 + i * 64;
        wreg[i] = (col < n && k < kdim) ? __ldg(reinterpret_cast<const uint32_t*>(w + (size_t)col * kdim + k)) : 0u;
    }
    for (int k = tid; k < kdim; k += 256) {
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = j < ns ? h_to_f(x[(size_t)(j * step) * x_ld + k]) : 0.f;
        float m = 0.f;
        const float a = sc ? sc[k] : 1.f, b = sc ? sh[k] : 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j)
            if (j < ns) m += sc ? fmaxf(fmaf(v[j], a, b), 0.f) : v[j];
        sM[k] = m * inv;
        // per-column constants the input rows were centred by: up to two column segments (the feature and the centre-feature
        // segment of grouped rows), each sums[] * scale = a channel mean
        float o = 0.f;
        if (off0 && k >= off0_start && k < off0_start + off0_count) o = off0[k - off0_start] * off0_scale;
        else if (off1 && k >= off1_start && k < off1_start + off1_count) o = off1[k - off1_start] * off1_scale;
        sO[k] = o;
    }
    __syncthreads();
    if (col >= n) return;
    float acc = 0.f, acc2 = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int k = lane * 2 + i * 64;  // kdim is a multiple of 32: pairs never straddle the end
        if (k < kdim) {
            const float2 wv = h2_to_f2(wreg[i]);
            acc = fmaf(wv.x, sM[k], fmaf(wv.y, sM[k + 1], acc));
            acc2 = fmaf(wv.x, sO[k], fmaf(wv.y, sO[k + 1], acc2));
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc += __shfl_xor_sync(kFull, acc, o);
        acc2 += __shfl_xor_sync(kFull, acc2, o);
    }
    if (lane == 0) {
        c[col] = acc;
        c_true[col] = acc + acc2;
    }
}

// ------------------------------------------------------------------ small per-channel kernels
__global__ void bn_finalize_kernel(int n, float inv_rows, float unbias, const float* __restrict__ sums,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   const float* __restrict__ bias, const float* __restrict__ center, float momentum,
                                   float eps, float* running_mean, float* running_var,
                                   long long* num_batches_tracked, float* scale, float* shift, float* mean_out,
                                   float* rstd_out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c == 0 && num_batches_tracked) *num_batches_tracked += 1;
    if (c >= n) return;
    bn_finalize_channel(c, n, sums[c], sums[n + c], inv_rows, unbias, gamma, beta, bias, center, momentum, eps, running_mean,
                        running_var, scale, shift, mean_out, rstd_out);
}

__global__ void bn_eval_affine_kernel(int n, const float* __restrict__ gamma, const float* __restrict__ beta,
                                      const float* __restrict__ bias, const float* __restrict__ center,
                                      const float* __restrict__ running_mean, const float* __restrict__ running_var,
                                      float eps, float* scale, float* shift) {
    pdl_enter();  // programmatic dependent launch (pn2_common.cuh): first statement, before any memory access
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    const float sc = gamma[c] * rsqrtf(running_var[c] + eps);
    scale[c] = sc;
    shift[c] = fmaf((bias ? bias[c] : 0.f) + (center ? center[c] : 0.f) - running_mean[c], sc, beta[c]);
}

// BatchNorm-backward coefficients of a layer,  dY = cA*dZ + cB*Y + cC  per output channel, the parameter gradients
// dgamma / dbeta, and (w != null) the layer's weights with the coefficients FOLDED IN for the input-gradient GEMM:
//     dY * W  =  dZ * (cA o W)  +  Y * (cB o W)  +  cC * W
// so that GEMM multiplies the stored dZ and Y rows as they are (no per-element transform of the big operands):
//     wa[k][n] = bf16(S * cA[n] * w[n][k]),  wb[k][n] = fp16(S * cB[n] * w[n][k]),  negbias[k] = -sum_n cC[n] * w[n][k]
// Y is stored fp16 and the tensor core multiplies like with like, so W_B is fp16 too; cB is gradient-sized (1e-8 is
// ordinary), hence the power-of-two S = 2^(4 - ceil(log2 max|cB|)) that brings it into fp16's normal range; W_A carries
// the same S so both products land in one accumulator.  wb_unscale = 1/S is applied by the GEMM's epilogue.
// Block (bx, by) owns the 32 x 32 tile (input channels 32bx.., output channels 32by..) of the folded weights; blocks
// with by == 0 also reduce the bias of their input channels; block (0,0) writes the coefficients and the parameter
// gradients.  Every block recomputes the (cheap) coefficients of all n channels: S needs max|cB|.
__global__ void __launch_bounds__(256) bn_bwd_coefs_kernel(int n, float inv_rows, const float* __restrict__ sums,
                                                           const float* __restrict__ gamma, const float* __restrict__ mean,
                                                           const float* __restrict__ rstd, float* cA, float* cB, float* cC,
                                                           float* dgamma, float* dbeta, int accumulate,
                                                           const float* __restrict__ w, int k_true, int kp, bf16* wa,
                                                           act_t* wb, float* negbias, float* wb_unscale) {
    pdl_enter();  // programmatic dependent launch (pn2_common.cuh): first statement, before any memory access
    extern __shared__ float sco[];  // [3][n]
    __shared__ float tile[32][33];
    __shared__ float sbias[8][32];
    __shared__ float smax[8];
    const int tid = threadIdx.x;
    const bool first = blockIdx.x == 0 && blockIdx.y == 0;
    // this thread's four weights of the tile first: their loads do not depend on the coefficients (latency overlap)
    const int k0 = blockIdx.x * 32, n0 = blockIdx.y * 32, tx = tid & 31, ty = tid >> 5;  // 32 x 8
    float wv[4] = {0.f, 0.f, 0.f, 0.f};
    if (w) {
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {  // w[n0 + ty + 8 jj][k0 + tx]: k contiguous
            const int nn = n0 + ty + 8 * jj, kk = k0 + tx;
            if (nn < n && kk < k_true) wv[jj] = __ldg(w + (size_t)nn * k_true + kk);
        }
    }
    for (int c = tid; c < n; c += 256) {
        const float s1 = sums[c], s2 = sums[n + c];
        const float m1 = s1 * inv_rows, m2 = s2 * inv_rows;
        const float gr = gamma[c] * rstd[c];
        const float a_ = gr, b_ = -gr * m2 * rstd[c], c_ = gr * (m2 * rstd[c] * mean[c] - m1);
        sco[c] = a_; sco[n + c] = b_; sco[2 * n + c] = c_;
        if (first) {
            cA[c] = a_; cB[c] = b_; cC[c] = c_;
            if (accumulate) {  // straight into the parameters' .grad (autograd's "+=" without an extra kernel)
                dgamma[c] += s2;
                dbeta[c] += s1;
            } else {
                dgamma[c] = s2;
                dbeta[c] = s1;
            }
        }
    }
    if (!w) return;
    __syncthreads();
    float mx = 0.f;  // every block derives the same scale from the same coefficients
    for (int c = tid; c < n; c += 256) mx = fmaxf(mx, fabsf(sco[n + c]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(kFull, mx, o));
    if ((tid & 31) == 0) smax[tid >> 5] = mx;
    __syncthreads();
    mx = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) mx = fmaxf(mx, smax[i]);
    int ex = 0;
    if (mx > 0.f && mx < 3e38f) frexpf(mx, &ex);  // mx = f * 2^ex, f in [0.5, 1)
    ex = max(-100, min(100, ex));
    const float wscale = exp2f((float)(4 - ex)), unscale = exp2f((float)(ex - 4));
    if (first && tid == 0) *wb_unscale = unscale;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) tile[ty + 8 * jj][tx] = wv[jj];
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {  // write [k0 + j][n0 + tx]: n contiguous
        const int kk = k0 + j, nn = n0 + tx;
        if (kk < kp && nn < n) {
            const float v = tile[tx][j];
            wa[(size_t)kk * n + nn] = __float2bfloat16(sco[nn] * wscale * v);
            wb[(size_t)kk * n + nn] = f_to_h(sco[n + nn] * wscale * v);
        }
    }
    if (blockIdx.y == 0) {  // negbias[k] = -sum_n cC[n] * w[n][k] for this block's 32 input channels
        const int kk = k0 + tx;
        float bacc = 0.f;
        if (kk < k_true) {
#pragma unroll 4
            for (int nn = ty; nn < n; nn += 8) bacc = fmaf(sco[2 * n + nn], __ldg(w + (size_t)nn * k_true + kk), bacc);
        }
        sbias[ty][tx] = bacc;
        __syncthreads();
        if (ty == 0 && kk < kp) {
            float t = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) t += sbias[i][tx];
            negbias[kk] = -t;
        }
    }
}

__global__ void prep_weights_kernel(int n, int k_true, int kp, const float* __restrict__ w, act_t* __restrict__ wh,
                                    bf16* __restrict__ wt, act_t* __restrict__ wl) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * kp) return;
    const int r = i / kp, c = i - r * kp;
    const float v = c < k_true ? w[(size_t)r * k_true + c] : 0.f;
    const act_t h = f_to_h(v);
    wh[i] = h;
    if (wl) wl[i] = f_to_h(v - h_to_f(h));  // lo plane of the two-plane (hi + lo) form
    if (wt) wt[(size_t)c * n + r] = __float2bfloat16(v);
}

// every layer's weights in one launch: blockIdx.y = layer, grid-stride over its [n][kp] elements
struct PrepDesc {
    const float* w;
    act_t* wh;
    int n, k_true, kp, two;  // two != 0: the lo plane follows the hi plane (wh + n * kp)
};
__global__ void __launch_bounds__(256) prep_weights_multi_kernel(const PrepDesc* __restrict__ descs) {
    pdl_enter();  // programmatic dependent launch (pn2_common.cuh): first statement, before any memory access
    const PrepDesc d = descs[blockIdx.y];
    const int total = d.n * d.kp;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int r = i / d.kp, c = i - r * d.kp;
        const float v = c < d.k_true ? d.w[(size_t)r * d.k_true + c] : 0.f;
        const act_t h = f_to_h(v);
        d.wh[i] = h;
        if (d.two) d.wh[total + i] = f_to_h(v - h_to_f(h));
    }
}

}  // namespace
}  // namespace pn2

using namespace pn2;

extern "C" int pn2_mlp_prep_weights_multi(int n_layers, const void* descs, pn2_stream_t stream) {
    if (n_layers < 0) return fail_arg("pn2_mlp_prep_weights_multi", "negative layer count");
    if (n_layers == 0) return 0;
    if (!descs) return fail_arg("pn2_mlp_prep_weights_multi", "null pointer");
    launch_k(prep_weights_multi_kernel, dim3(dim3(32, n_layers)), dim3(256), 0, (cudaStream_t)stream, (const PrepDesc*)descs);
    PN2_CHECK_LAUNCH("prep_weights_multi_kernel");
    return 0;
}

extern "C" int pn2_mlp_center(long long rows, int kdim, int n, const void* x, int x_ld, const float* in_scale,
                              const float* in_shift, const void* w, const float* off0, float off0_scale, int off0_start,
                              int off0_count, const float* off1, float off1_scale, int off1_start, int off1_count,
                              float* center, float* center_true, pn2_stream_t stream) {
    if (int e = check_common("pn2_mlp_center", rows, kdim, n)) return e;
    if (rows == 0) return 0;
    if (kdim > 1024) return fail_arg("pn2_mlp_center", "kdim > 1024");
    if (!x || !w || !center || !center_true) return fail_arg("pn2_mlp_center", "null pointer");
    launch_k(center_kernel, dim3((n + 7) / 8), dim3(256), 0, (cudaStream_t)stream, rows, kdim, n, (const act_t*)x, x_ld, in_scale, in_shift,
                                                                  (const act_t*)w, off0, off0_scale, off0_start, off0_count, off1,
                                                                  off1_scale, off1_start, off1_count, center, center_true);
    PN2_CHECK_LAUNCH("center_kernel");
    return 0;
}

struct PoolOut {
    int k = 0;
    const float* gamma = nullptr;
    float* val = nullptr;
    int* arg = nullptr;
};
static int check_pool(const char* who, const PoolOut& po, long long rows) {
    if (po.k == 0) return 0;
    if (po.k != 16 && po.k != 32 && po.k != 64 && po.k != 128) return fail_arg(who, "epilogue pooling needs pool_k in {16, 32, 64, 128}");
    if (rows % po.k) return fail_arg(who, "rows must be a multiple of pool_k");
    if (!po.gamma || !po.val || !po.arg) return fail_arg(who, "null pooling pointer");
    return 0;
}

static int gemm_fwd_impl(const char* who, long long rows, int kdim, int n, const void* x, const void* x_lo, int x_ld,
                         const float* in_scale, const float* in_shift, const void* w, const void* w_lo,
                         const float* center, void* y, void* y_lo, int y_ld, float* stats, pn2_stream_t stream,
                         const PoolOut& po = PoolOut()) {
    if (int e = check_pool(who, po, rows)) return e;
    if (int e = check_common(who, rows, kdim, n)) return e;
    if (rows == 0) return 0;
    if (!x || !w || !y) return fail_arg(who, "null pointer");
    if (x_ld % 8 || y_ld % 8 || x_ld < kdim || y_ld < n) return fail_arg(who, "bad leading dimension");
    if ((x_lo != nullptr) != (w_lo != nullptr) || (y_lo && !x_lo)) return fail_arg(who, "two-plane operands come as x_lo AND w_lo");
    GemmArgs a{};
    a.rows = rows; a.kdim = kdim; a.n = n;
    a.a0 = (const uint16_t*)x; a.a0_ld = x_ld;
    a.a1 = (const uint16_t*)x_lo; a.a1_ld = x_ld;
    a.c0 = in_scale; a.c1 = in_shift;
    a.b = (const uint16_t*)w; a.b1 = (const uint16_t*)w_lo;
    a.center = center;
    a.out = (uint16_t*)y; a.out_ld = y_ld; a.out_lo = (uint16_t*)y_lo;
    a.sums = stats;
    a.pool_k = po.k; a.pool_gamma = po.gamma; a.pool_val = po.val; a.pool_arg = po.arg;
    if (gemm_use_tc() || x_lo || po.k) return launch_gemm_tc(a, in_scale ? A_AFFINE : A_PLAIN, false, (cudaStream_t)stream);
    if (in_scale) return dispatch_bn<A_AFFINE, false>(a, (cudaStream_t)stream);
    return dispatch_bn<A_PLAIN, false>(a, (cudaStream_t)stream);
}

extern "C" int pn2_mlp_gemm_fwd(long long rows, int kdim, int n, const void* x, int x_ld, const float* in_scale,
                                const float* in_shift, const void* w, const float* center, void* y, int y_ld,
                                float* stats, pn2_stream_t stream) {
    return gemm_fwd_impl("pn2_mlp_gemm_fwd", rows, kdim, n, x, nullptr, x_ld, in_scale, in_shift, w, nullptr, center, y,
                         nullptr, y_ld, stats, stream);
}

extern "C" int pn2_mlp_gemm_fwd_x2(long long rows, int kdim, int n, const void* x, const void* x_lo, int x_ld,
                                   const float* in_scale, const float* in_shift, const void* w, const void* w_lo,
                                   const float* center, void* y, void* y_lo, int y_ld, float* stats,
                                   pn2_stream_t stream) {
    return gemm_fwd_impl("pn2_mlp_gemm_fwd_x2", rows, kdim, n, x, x_lo, x_ld, in_scale, in_shift, w, w_lo, center, y, y_lo,
                         y_ld, stats, stream);
}

static int gemm_fwd_bn_impl(const char* who, long long rows, int kdim, int n, const void* x, const void* x_lo, int x_ld,
                            const float* in_scale, const float* in_shift, const void* w, const void* w_lo,
                            const float* center, void* y, void* y_lo, int y_ld, float* stats, unsigned int* counter,
                            const float* gamma, const float* beta, const float* conv_bias, const float* center_true,
                            float momentum, float eps, float* running_mean, float* running_var,
                            long long* num_batches_tracked, float* scale, float* shift, float* mean, float* rstd,
                            float* next_center, pn2_stream_t stream, const PoolOut& po = PoolOut()) {
    if (int e = check_common(who, rows, kdim, n)) return e;
    if (int e = check_pool(who, po, rows)) return e;
    if (rows == 0) return fail_arg(who, "BatchNorm statistics of zero rows");
    if (!x || !w || !y || !stats || !counter || !gamma || !beta || !scale || !shift || !mean || !rstd)
        return fail_arg(who, "null pointer");
    if (x_ld % 8 || y_ld % 8 || x_ld < kdim || y_ld < n) return fail_arg(who, "bad leading dimension");
    if ((x_lo != nullptr) != (w_lo != nullptr) || (y_lo && !x_lo)) return fail_arg(who, "two-plane operands come as x_lo AND w_lo");
    GemmArgs a{};
    a.rows = rows; a.kdim = kdim; a.n = n;
    a.a0 = (const uint16_t*)x; a.a0_ld = x_ld;
    a.a1 = (const uint16_t*)x_lo; a.a1_ld = x_ld;
    a.c0 = in_scale; a.c1 = in_shift;
    a.b = (const uint16_t*)w; a.b1 = (const uint16_t*)w_lo;
    a.center = center;
    a.out = (uint16_t*)y; a.out_ld = y_ld; a.out_lo = (uint16_t*)y_lo;
    a.sums = stats;
    a.fin_counter = counter; a.fin_gamma = gamma; a.fin_beta = beta; a.fin_bias = conv_bias; a.fin_center = center_true;
    a.fin_momentum = momentum; a.fin_eps = eps;
    a.fin_unbias = rows > 1 ? (float)((double)rows / (double)(rows - 1)) : 1.f;
    a.fin_inv_rows = (float)(1.0 / (double)rows);
    a.fin_running_mean = running_mean; a.fin_running_var = running_var; a.fin_nbt = num_batches_tracked;
    a.fin_scale = scale; a.fin_shift = shift; a.fin_mean = mean; a.fin_rstd = rstd;
    a.fin_next_center = next_center;
    a.pool_k = po.k; a.pool_gamma = po.gamma; a.pool_val = po.val; a.pool_arg = po.arg;
    return launch_gemm_tc(a, in_scale ? A_AFFINE : A_PLAIN, false, (cudaStream_t)stream);
}

extern "C" int pn2_mlp_gemm_fwd_bn(long long rows, int kdim, int n, const void* x, int x_ld, const float* in_scale,
                                   const float* in_shift, const void* w, const float* center, void* y, int y_ld,
                                   float* stats, unsigned int* counter, const float* gamma, const float* beta,
                                   const float* conv_bias, const float* center_true, float momentum, float eps,
                                   float* running_mean, float* running_var, long long* num_batches_tracked,
                                   float* scale, float* shift, float* mean, float* rstd, float* next_center,
                                   pn2_stream_t stream) {
    return gemm_fwd_bn_impl("pn2_mlp_gemm_fwd_bn", rows, kdim, n, x, nullptr, x_ld, in_scale, in_shift, w, nullptr, center, y,
                            nullptr, y_ld, stats, counter, gamma, beta, conv_bias, center_true, momentum, eps, running_mean,
                            running_var, num_batches_tracked, scale, shift, mean, rstd, next_center, stream);
}

extern "C" int pn2_mlp_gemm_fwd_bn_x2(long long rows, int kdim, int n, const void* x, const void* x_lo, int x_ld,
                                      const float* in_scale, const float* in_shift, const void* w, const void* w_lo,
                                      const float* center, void* y, void* y_lo, int y_ld, float* stats,
                                      unsigned int* counter, const float* gamma, const float* beta,
                                      const float* conv_bias, const float* center_true, float momentum, float eps,
                                      float* running_mean, float* running_var, long long* num_batches_tracked,
                                      float* scale, float* shift, float* mean, float* rstd, float* next_center,
                                      pn2_stream_t stream) {
    return gemm_fwd_bn_impl("pn2_mlp_gemm_fwd_bn_x2", rows, kdim, n, x, x_lo, x_ld, in_scale, in_shift, w, w_lo, center, y,
                            y_lo, y_ld, stats, counter, gamma, beta, conv_bias, center_true, momentum, eps, running_mean,
                            running_var, num_batches_tracked, scale, shift, mean, rstd, next_center, stream);
}

extern "C" int pn2_mlp_gemm_fwd_pool(long long rows, int kdim, int n, const void* x, const void* x_lo, int x_ld,
                                     const float* in_scale, const float* in_shift, const void* w, const void* w_lo,
                                     const float* center, void* y, void* y_lo, int y_ld, float* stats, int pool_k,
                                     const float* pool_gamma, float* pool_val, int* pool_arg, pn2_stream_t stream) {
    PoolOut po;
    po.k = pool_k; po.gamma = pool_gamma; po.val = pool_val; po.arg = pool_arg;
    return gemm_fwd_impl("pn2_mlp_gemm_fwd_pool", rows, kdim, n, x, x_lo, x_ld, in_scale, in_shift, w, w_lo, center, y, y_lo,
                         y_ld, stats, stream, po);
}

extern "C" int pn2_mlp_gemm_fwd_bn_pool(long long rows, int kdim, int n, const void* x, const void* x_lo, int x_ld,
                                        const float* in_scale, const float* in_shift, const void* w, const void* w_lo,
                                        const float* center, void* y, void* y_lo, int y_ld, float* stats,
                                        unsigned int* counter, const float* gamma, const float* beta,
                                        const float* conv_bias, const float* center_true, float momentum, float eps,
                                        float* running_mean, float* running_var, long long* num_batches_tracked,
                                        float* scale, float* shift, float* mean, float* rstd, float* next_center,
                                        int pool_k, float* pool_val, int* pool_arg, pn2_stream_t stream) {
    PoolOut po;
    po.k = pool_k; po.gamma = gamma; po.val = pool_val; po.arg = pool_arg;
    return gemm_fwd_bn_impl("pn2_mlp_gemm_fwd_bn_pool", rows, kdim, n, x, x_lo, x_ld, in_scale, in_shift, w, w_lo, center, y,
                            y_lo, y_ld, stats, counter, gamma, beta, conv_bias, center_true, momentum, eps, running_mean,
                            running_var, num_batches_tracked, scale, shift, mean, rstd, next_center, stream, po);
}

extern "C" int pn2_mlp_gemm_dgrad(long long rows, int n_red, int k_out, const void* dz, int dz_ld, const void* y,
                                  int y_ld, const void* wa, const void* wb, const float* negbias, const float* wb_unscale,
                                  const void* y_prev, int y_prev_ld, const float* prev_scale, const float* prev_shift,
                                  const float* prev_mean, const float* prev_rstd, void* dz_prev, int dz_prev_ld,
                                  float* sums_prev, pn2_stream_t stream) {
    if (int e = check_common("pn2_mlp_gemm_dgrad", rows, n_red, k_out)) return e;
    if (rows == 0) return 0;
    if (!dz || !y || !wa || !wb || !negbias || !wb_unscale || !dz_prev) return fail_arg("pn2_mlp_gemm_dgrad", "null pointer");
    if (dz_ld % 8 || y_ld % 8 || dz_prev_ld % 8 || dz_prev_ld < k_out)
        return fail_arg("pn2_mlp_gemm_dgrad", "bad leading dimension");
    GemmArgs a{};
    a.rows = rows; a.kdim = n_red; a.n = k_out;
    a.a0 = (const uint16_t*)dz; a.a0_ld = dz_ld;
    a.a1 = (const uint16_t*)y; a.a1_ld = y_ld;
    a.b = (const uint16_t*)wa; a.b1 = (const uint16_t*)wb; a.yscale = wb_unscale;
    a.center = negbias;
    a.out = (uint16_t*)dz_prev; a.out_ld = dz_prev_ld;
    if (y_prev) {
        if (!prev_scale || !prev_shift || !prev_mean || !prev_rstd || !sums_prev || y_prev_ld % 8)
            return fail_arg("pn2_mlp_gemm_dgrad", "masking needs the previous layer's constants");
        a.sums = sums_prev;
        a.yp = (const uint16_t*)y_prev; a.yp_ld = y_prev_ld;
        a.p_scale = prev_scale; a.p_shift = prev_shift; a.p_mean = prev_mean; a.p_rstd = prev_rstd;
        return launch_gemm_tc(a, A_BNBWD, true, (cudaStream_t)stream);
    }
    a.sums = nullptr;
    return launch_gemm_tc(a, A_BNBWD, false, (cudaStream_t)stream);
}

extern "C" int pn2_mlp_gemm_wgrad(long long rows, int n, int kp, int k_true, const void* dz, int dz_ld,
                                  const void* y, int y_ld, const float* cA, const float* cB, const float* cC,
                                  const void* x, int x_ld, const float* in_scale, const float* in_shift, float* dw,
                                  int dw_ld, pn2_stream_t stream) {
    if (rows < 0 || n <= 0 || kp <= 0 || k_true <= 0 || k_true > kp) return fail_arg("pn2_mlp_gemm_wgrad", "bad size");
    if (rows == 0) return 0;
    if (n % 8 || kp % 8 || dz_ld % 8 || y_ld % 8 || x_ld % 8) return fail_arg("pn2_mlp_gemm_wgrad", "sizes must be multiples of 8");
    if (!dz || !y || !cA || !cB || !cC || !x || !dw) return fail_arg("pn2_mlp_gemm_wgrad", "null pointer");
    WgradArgs a{};
    a.rows = rows; a.n = n; a.kp = kp; a.k_true = k_true;
    a.dz = (const bf16*)dz; a.dz_ld = dz_ld;
    a.y = (const act_t*)y; a.y_ld = y_ld;
    a.cA = cA; a.cB = cB; a.cC = cC;
    a.x = (const act_t*)x; a.x_ld = x_ld;
    a.in_scale = in_scale; a.in_shift = in_shift;
    a.dw = dw; a.dw_ld = dw_ld;
    const int gx = (n + TN - 1) / TN, gy = (kp + TK - 1) / TK;
    const long long chunks = (rows + WBR - 1) / WBR;
    long long gz = 296 / (gx * gy);  // two CTAs per SM (shared memory), all resident
    if (gz < 1) gz = 1;
    if (gz > chunks) gz = chunks;
    dim3 grid(gx, gy, (unsigned)gz);
    static DeviceOnce once;
    if (once.first()) {
        PN2_CHECK(cudaFuncSetAttribute(wgrad_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgradSmem),
                  "wgrad: cudaFuncSetAttribute");
        PN2_CHECK(cudaFuncSetAttribute(wgrad_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgradSmem),
                  "wgrad: cudaFuncSetAttribute");
    }
    if (wgrad_use_tc() && wgrad_tc_supported(a)) return launch_wgrad_tc(a, (cudaStream_t)stream);
    if (in_scale)
        wgrad_kernel<true><<<grid, kWgradThreads, kWgradSmem, (cudaStream_t)stream>>>(a);
    else
        wgrad_kernel<false><<<grid, kWgradThreads, kWgradSmem, (cudaStream_t)stream>>>(a);
    PN2_CHECK_LAUNCH("wgrad_kernel");
    return 0;
}

extern "C" int pn2_bn_finalize(int n, long long rows, const float* sums, const float* gamma, const float* beta,
                               const float* conv_bias, const float* center, float momentum, float eps, float* running_mean,
                               float* running_var, long long* num_batches_tracked, float* scale, float* shift,
                               float* mean, float* rstd, pn2_stream_t stream) {
    if (n <= 0 || rows <= 0) return fail_arg("pn2_bn_finalize", "non-positive size");
    if (!sums || !gamma || !beta || !scale || !shift || !mean || !rstd) return fail_arg("pn2_bn_finalize", "null pointer");
    const float unbias = rows > 1 ? (float)((double)rows / (double)(rows - 1)) : 1.f;
    bn_finalize_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        n, (float)(1.0 / (double)rows), unbias, sums, gamma, beta, conv_bias, center, momentum, eps, running_mean, running_var,
        num_batches_tracked, scale, shift, mean, rstd);
    PN2_CHECK_LAUNCH("bn_finalize_kernel");
    return 0;
}

extern "C" int pn2_bn_eval_affine(int n, const float* gamma, const float* beta, const float* conv_bias,
                                  const float* center, const float* running_mean, const float* running_var, float eps, float* scale,
                                  float* shift, pn2_stream_t stream) {
    if (n <= 0) return fail_arg("pn2_bn_eval_affine", "non-positive size");
    if (!gamma || !beta || !running_mean || !running_var || !scale || !shift)
        return fail_arg("pn2_bn_eval_affine", "null pointer");
    launch_k(bn_eval_affine_kernel, dim3((n + 127) / 128), dim3(128), 0, (cudaStream_t)stream, n, gamma, beta, conv_bias, center, running_mean,
                                                                            running_var, eps, scale, shift);
    PN2_CHECK_LAUNCH("bn_eval_affine_kernel");
    return 0;
}

extern "C" int pn2_bn_bwd_coefs(int n, long long rows, const float* sums, const float* gamma, const float* mean,
                                const float* rstd, float* cA, float* cB, float* cC, float* dgamma, float* dbeta,
                                int accumulate, const float* w, int k_true, int kp, void* wa, void* wb, float* negbias,
                                float* wb_unscale, pn2_stream_t stream) {
    if (n <= 0 || rows <= 0) return fail_arg("pn2_bn_bwd_coefs", "non-positive size");
    if (!sums || !gamma || !mean || !rstd || !cA || !cB || !cC || !dgamma || !dbeta)
        return fail_arg("pn2_bn_bwd_coefs", "null pointer");
    if (n > 4096) return fail_arg("pn2_bn_bwd_coefs", "n > 4096");
    if (w && (!wa || !wb || !negbias || !wb_unscale || k_true <= 0 || kp < k_true)) return fail_arg("pn2_bn_bwd_coefs", "bad folding arguments");
    const dim3 blocks(w ? (kp + 31) / 32 : 1, w ? (n + 31) / 32 : 1);
    launch_k(bn_bwd_coefs_kernel, dim3(blocks), dim3(256), 3 * n * sizeof(float), (cudaStream_t)stream, 
        n, (float)(1.0 / (double)rows), sums, gamma, mean, rstd, cA, cB, cC, dgamma, dbeta, accumulate, w, k_true, kp, (bf16*)wa,
        (act_t*)wb, negbias, wb_unscale);
    PN2_CHECK_LAUNCH("bn_bwd_coefs_kernel");
    return 0;
}

extern "C" int pn2_mlp_prep_weights(int n, int k_true, int kp, const float* w, void* w_f16, void* wt_bf16,
                                    pn2_stream_t stream) {
    if (n <= 0 || k_true <= 0 || kp < k_true) return fail_arg("pn2_mlp_prep_weights", "bad size");
    if (!w || !w_f16) return fail_arg("pn2_mlp_prep_weights", "null pointer");
    const int total = n * kp;
    prep_weights_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(n, k_true, kp, w, (act_t*)w_f16,
                                                                               (bf16*)wt_bf16, nullptr);
    PN2_CHECK_LAUNCH("prep_weights_kernel");
    return 0;
}

extern "C" int pn2_mlp_prep_weights_x2(int n, int k_true, int kp, const float* w, void* w_hi, void* w_lo,
                                       pn2_stream_t stream) {
    if (n <= 0 || k_true <= 0 || kp < k_true) return fail_arg("pn2_mlp_prep_weights_x2", "bad size");
    if (!w || !w_hi || !w_lo) return fail_arg("pn2_mlp_prep_weights_x2", "null pointer");
    const int total = n * kp;
    prep_weights_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(n, k_true, kp, w, (act_t*)w_hi, nullptr,
                                                                               (act_t*)w_lo);
    PN2_CHECK_LAUNCH("prep_weights_kernel");
    return 0;
}
