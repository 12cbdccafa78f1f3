// mlp_wgrad_tc.cu -- weight gradient of the grouped per-point MLP on tcgen05 tensor cores, sm_100a.
//
//     dW[n][k] += sum_r dY[r][n] * X'[r][k]      dY = cA*dZ + cB*Y + cC  (BatchNorm backward)
//                                                X' = relu(X*scale+shift) | X   (the layer's input)
//
// (the conv1x1 weight gradient cuDNN computes for reference pointnet_utils.py:399-403,458-460,505-507,577-580
// and backbones.py:131 in backward).  The per-channel coefficients are pulled OUT of the row sum,
//
//     dW[n][k] = cA[n] * (dZ^T X')[n][k]  +  cB[n] * (Y^T X')[n][k]  +  cC[n] * (1^T X')[k]
//
// so the two big operands dZ (bf16) and Y (fp16) go from HBM to the tensor core exactly as stored; only X -- usually
// the narrow one -- passes through registers (BatchNorm+ReLU of the producing layer, and a bf16 copy: one
// tcgen05.mma multiplies like with like, so dZ meets X' in bf16 and Y meets it in fp16).  The products accumulate
// in TMEM over ALL rows of the CTA: G1 = dZ^T X', G2 = Y^T X' (MT 128-channel tiles each) and G3 = 1^T X' (a
// constant panel of ones as the A operand); the epilogue combines them and adds into dW with fp32 atomics.
//
// The reduction runs over ROWS, so both operands are "MN-major" for the tensor core (memory rows are K slices):
//     panel = [16 rows][64 columns] 16-bit, 128-byte row pitch, 16-byte pieces XOR-swizzled by row % 8
//             (UMMA canonical MN-major SWIZZLE_128B: LBO = panel pitch 2048 B, SBO = 1024 B per 8 rows)
//     stage = 16 rows: dZ panels (n_pad/64) | Y panels (n_pad/64) | X fp16 panels (kw_pad/64) | X bf16 panels
// Two panels are exactly one 16-byte piece per producer thread.
//
//   8 producer warps  cp.async raw pieces, D stages in flight; when a stage has landed each thread rewrites ITS X
//                     pieces (BN+ReLU, bf16 copy), fence.proxy.async, mbarrier arrive
//   1 MMA warp        one lane: per stage (2 MT + 1) tcgen05.mma (M = 128, N = kw, K = 16); tcgen05.commit frees the
//                     stage; a final commit publishes the accumulators
//   4 epilogue warps  wait for the final commit, tcgen05.ld, combine with cA / cB / cC, atomics into dW
//
// Grid: (row splits, input-channel tiles of kw <= 512 / (2 MT + 1) columns).
// Bound: HBM reads rows*(2n + k)*2 bytes.
#include "mlp_gemm.cuh"
#include "tc_common.cuh"

#include <cstdlib>
#include <cstring>

namespace pn2 {
namespace {

constexpr int WR = 16;            // rows per stage = K of one tcgen05.mma
constexpr int kPanel = WR * 128;  // bytes
constexpr int kWEpiWarps = 4, kWMmaWarp = 4, kWProdWarps = 8;
constexpr int kWProdThreads = kWProdWarps * 32;
constexpr int kWThreads = (kWEpiWarps + 1 + kWProdWarps) * 32;
constexpr int kWSmem = 225 * 1024;
constexpr int kWMaxStages = 8;

struct WgTc {
    WgradArgs a;
    int mt;        // 128-wide output-channel tiles
    int kw;        // input channels per CTA (multiple of 16, (2 mt + 1) * kw <= 512)
    int npd, npx;  // dZ (= Y) panels, X panels (per format) per stage
    int nst;       // ring depth
    int tmem_cols; // power of two >= (2 mt + 1) * kw
};

template <bool AFFINE>
__global__ void __launch_bounds__(kWThreads, 1) wgrad_tc_kernel(const WgTc w) {
    const WgradArgs& p = w.a;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    unsigned char* base = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
    const int stage_bytes = (2 * w.npd + 2 * w.npx) * kPanel;
    unsigned char* sStage = base;
    unsigned char* sOnes = base + w.nst * stage_bytes;                  // two panels (128 columns) of fp16 ones
    float* sCo = reinterpret_cast<float*>(sOnes + 2 * kPanel);              // [3][npd*64] cA cB cC, then [2][npx*64] scale shift
    const int ncol = w.npd * 64, kcol_pad = w.npx * 64;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sCo + 3 * ncol + 2 * kcol_pad);
    uint64_t* full = bars;            // [nst]
    uint64_t* empty = bars + w.nst;   // [nst]
    uint64_t* done = bars + 2 * w.nst;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * w.nst + 1);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int k0 = blockIdx.y * w.kw;                 // first input channel of this CTA
    const int kw_here = min(w.kw, p.kp - k0);         // multiple of 16 (kp % 32 == 0, kw % 16 == 0)
    const long long stages = (p.rows + WR - 1) / WR;
    const long long mine = blockIdx.x < stages ? (stages - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    for (int i = tid; i < 3 * ncol; i += kWThreads) {
        const int which = i / ncol, c = i - which * ncol;
        sCo[i] = c < p.n ? (which == 0 ? p.cA : (which == 1 ? p.cB : p.cC))[c] : 0.f;
    }
    for (int i = tid; i < 2 * kcol_pad; i += kWThreads) {
        const int which = i / kcol_pad, c = i - which * kcol_pad;
        sCo[3 * ncol + i] = (AFFINE && k0 + c < p.kp) ? (which == 0 ? p.in_scale : p.in_shift)[k0 + c] : 0.f;
    }
    for (int i = tid; i < 2 * kPanel / 4; i += kWThreads) reinterpret_cast<uint32_t*>(sOnes)[i] = 0x3C003C00u;  // fp16 1.0 x 2
    if (tid == 0) {
        for (int i = 0; i < w.nst; ++i) {
            mbar_init(&full[i], kWProdThreads);
            mbar_init(&empty[i], 1);
        }
        mbar_init(done, 1);
        mbar_fence_init();
    }
    if (warp == kWMmaWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)w.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async();  // the ones panel (generic-proxy stores) is read by the tensor core
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp > kWMmaWarp) {
        // ================================ producers ================================
        const int pt = tid - (kWMmaWarp + 1) * 32;
        const int pp = pt >> 7;                     // which panel of a pair
        const int pj = pt & 7, pr = (pt & 127) >> 3;  // piece column, row
        const uint32_t poff = pr * 128 + ((pj ^ (pr & 7)) << 4);
        const uint32_t stage0 = smem_u32(sStage);
        // panels that hold real columns; the padding panels of every stage are zeroed once, here, and never written again
        const int npd_v = (p.n + 63) / 64, npx_v = (kw_here + 63) / 64;
        for (int sl = 0; sl < w.nst; ++sl) {
            unsigned char* st = sStage + sl * stage_bytes + poff;
            for (int P = npd_v + pp; P < w.npd; P += 2) {
                *reinterpret_cast<uint4*>(st + P * kPanel) = make_uint4(0u, 0u, 0u, 0u);
                *reinterpret_cast<uint4*>(st + (w.npd + P) * kPanel) = make_uint4(0u, 0u, 0u, 0u);
            }
            for (int P = npx_v + pp; P < w.npx; P += 2) {
                *reinterpret_cast<uint4*>(st + (2 * w.npd + P) * kPanel) = make_uint4(0u, 0u, 0u, 0u);
                *reinterpret_cast<uint4*>(st + (2 * w.npd + w.npx + P) * kPanel) = make_uint4(0u, 0u, 0u, 0u);
            }
        }
        const int D = w.nst - 2;  // one stage of slack between publishing a stage and needing its slot back (see mlp_gemm_tc.cu)
        long long i_s = blockIdx.x, p_s = blockIdx.x;
        int i_slot = 0, p_slot = 0;
        uint32_t i_phase = 0;
        for (long long c = 0; c < mine + D; ++c) {
            if (c < mine) {
                mbar_wait(&empty[i_slot], i_phase ^ 1);
                const uint32_t st = stage0 + i_slot * stage_bytes + poff;
                const long long row = i_s * WR + pr;
                const bool rok = row < p.rows;
                const bf16* dzp = p.dz + row * p.dz_ld + pj * 8;
                const act_t* yp = p.y + row * p.y_ld + pj * 8;
                for (int P = pp; P < npd_v; P += 2) {
                    const bool ok = rok && P * 64 + pj * 8 < p.n;
                    cp_async16_s(st + P * kPanel, ok ? (const void*)(dzp + P * 64) : (const void*)p.dz, ok ? 16 : 0);
                    cp_async16_s(st + (w.npd + P) * kPanel, ok ? (const void*)(yp + P * 64) : (const void*)p.y, ok ? 16 : 0);
                }
                const act_t* xp = p.x + row * p.x_ld + k0 + pj * 8;
                for (int P = pp; P < npx_v; P += 2) {
                    const bool ok = rok && P * 64 + pj * 8 < kw_here;
                    cp_async16_s(st + (2 * w.npd + P) * kPanel, ok ? (const void*)(xp + P * 64) : (const void*)p.x, ok ? 16 : 0);
                }
                i_s += gridDim.x;
                if (++i_slot == w.nst) { i_slot = 0; i_phase ^= 1; }
            }
            cp_async_commit();
            if (c >= D) {
                switch (D) {  // this thread's pieces of stage c - D have landed
                    case 0: cp_wait<0>(); break;
                    case 1: cp_wait<1>(); break;
                    case 2: cp_wait<2>(); break;
                    case 3: cp_wait<3>(); break;
                    case 4: cp_wait<4>(); break;
                    case 5: cp_wait<5>(); break;
                    default: cp_wait<6>(); break;
                }
                unsigned char* st = sStage + p_slot * stage_bytes + poff;
                const bool rok = p_s * WR + pr < p.rows;
                for (int P = pp; P < npx_v; P += 2) {
                    uint4* s16 = reinterpret_cast<uint4*>(st + (2 * w.npd + P) * kPanel);
                    uint4 v16 = make_uint4(0u, 0u, 0u, 0u), vbf = make_uint4(0u, 0u, 0u, 0u);
                    if (rok && P * 64 + pj * 8 < kw_here) {
                        const uint4 qx = *s16;
                        const uint32_t* x = reinterpret_cast<const uint32_t*>(&qx);
                        uint32_t* o16 = reinterpret_cast<uint32_t*>(&v16);
                        uint32_t* obf = reinterpret_cast<uint32_t*>(&vbf);
                        if (AFFINE) {
                            const float* cs = sCo + 3 * ncol + P * 64 + pj * 8;
#pragma unroll
                            for (int hh = 0; hh < 2; ++hh) {
                                const float4 s4 = *reinterpret_cast<const float4*>(cs + 4 * hh);
                                const float4 h4 = *reinterpret_cast<const float4*>(cs + kcol_pad + 4 * hh);
                                const float2 x0 = h2_to_f2(x[2 * hh]), x1 = h2_to_f2(x[2 * hh + 1]);
                                const float a0 = fmaxf(fmaf(x0.x, s4.x, h4.x), 0.f), a1 = fmaxf(fmaf(x0.y, s4.y, h4.y), 0.f);
                                const float a2 = fmaxf(fmaf(x1.x, s4.z, h4.z), 0.f), a3 = fmaxf(fmaf(x1.y, s4.w, h4.w), 0.f);
                                o16[2 * hh] = f2_to_h2(a0, a1); o16[2 * hh + 1] = f2_to_h2(a2, a3);
                                obf[2 * hh] = f2_to_bf2(a0, a1); obf[2 * hh + 1] = f2_to_bf2(a2, a3);
                            }
                        } else {
                            v16 = qx;
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float2 xv = h2_to_f2(x[e]);
                                obf[e] = f2_to_bf2(xv.x, xv.y);
                            }
                        }
                    }
                    if (AFFINE || !rok) *s16 = v16;
                    *reinterpret_cast<uint4*>(st + (2 * w.npd + w.npx + P) * kPanel) = vbf;
                }
#ifndef PN2_FENCE_CONSUMER
                fence_proxy_async();
#endif
                mbar_arrive(&full[p_slot]);
                p_s += gridDim.x;
                if (++p_slot == w.nst) p_slot = 0;
            }
        }
    } else if (warp == kWMmaWarp) {
        // ================================ MMA issuer ================================
        // A = dZ^T / Y^T / 1^T (M = output channels), B = X'^T (N = input channels): both MN-major
        const uint32_t idesc_bf = umma_idesc2(1u, 1u, true, true, 128, kw_here);
        const uint32_t idesc_h = umma_idesc2(0u, 0u, true, true, 128, kw_here);
        const uint32_t g2_col = w.mt * w.kw, g3_col = 2 * w.mt * w.kw;
        const uint64_t ones_desc = umma_desc_sw128(smem_u32(sOnes), kPanel, 1024);
        int slot = 0;
        uint32_t phase = 0;
        for (long long c = 0; c < mine; ++c) {
            mbar_wait(&full[slot], phase);
#ifdef PN2_FENCE_CONSUMER
            fence_proxy_async();
#endif
            tc_fence_after();
            if (lane == 0) {
                const uint32_t sa = smem_u32(sStage + slot * stage_bytes);
                const uint32_t sy = sa + w.npd * kPanel;
                const uint64_t x16 = umma_desc_sw128(sa + 2 * w.npd * kPanel, kPanel, 1024);
                const uint64_t xbf = umma_desc_sw128(sa + (2 * w.npd + w.npx) * kPanel, kPanel, 1024);
                const uint32_t acc = c != 0;
                for (int m = 0; m < w.mt; ++m) {
                    umma_f16(tmem_base + m * w.kw, umma_desc_sw128(sa + 2 * m * kPanel, kPanel, 1024), xbf, idesc_bf, acc);
                    umma_f16(tmem_base + g2_col + m * w.kw, umma_desc_sw128(sy + 2 * m * kPanel, kPanel, 1024), x16, idesc_h, acc);
                }
                umma_f16(tmem_base + g3_col, ones_desc, x16, idesc_h, acc);
                tc_commit(&empty[slot]);
                if (c == mine - 1) tc_commit(done);
            }
            __syncwarp();
            if (++slot == w.nst) { slot = 0; phase ^= 1; }
        }
    } else if (mine > 0) {
        // ================================ epilogue ================================
        mbar_wait(done, 0);
        tc_fence_after();
        const bool vec4 = (p.dw_ld & 3) == 0 && (reinterpret_cast<uintptr_t>(p.dw) & 15) == 0;  // 16-byte aligned rows
        const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
        const uint32_t g2_col = w.mt * w.kw, g3_col = 2 * w.mt * w.kw;
        for (int c16 = 0; c16 < kw_here; c16 += 16) {
            uint32_t g3[32];
            tmem_ld16(tmem_base + lane_base + g3_col + c16, g3);  // every lane holds the same row: 1^T X'
            for (int m = 0; m < w.mt; ++m) {
                if (m * 128 + warp * 32 >= p.n) continue;  // warp-uniform
                const int n = m * 128 + warp * 32 + lane;  // output channel of this thread (TMEM lane)
                uint32_t g1[32], g2[32];
                tmem_ld16(tmem_base + lane_base + m * w.kw + c16, g1);
                tmem_ld16(tmem_base + lane_base + g2_col + m * w.kw + c16, g2);
                if (n < p.n) {
                    const float ca = sCo[n], cb = sCo[ncol + n], cc = sCo[2 * ncol + n];
                    float* dst = p.dw + (size_t)n * p.dw_ld + k0 + c16;
                    float g[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e)
                        g[e] = fmaf(ca, __uint_as_float(g1[e]), fmaf(cb, __uint_as_float(g2[e]), cc * __uint_as_float(g3[e])));
                    if (vec4 && k0 + c16 + 16 <= p.k_true) {
#pragma unroll
                        for (int e = 0; e < 16; e += 4) red_add_v4(dst + e, g[e], g[e + 1], g[e + 2], g[e + 3]);
                    } else {
#pragma unroll
                        for (int e = 0; e < 16; ++e)
                            if (k0 + c16 + e < p.k_true && g[e] != 0.f) atomicAdd(dst + e, g[e]);
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kWMmaWarp) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)w.tmem_cols)
                     : "memory");
    }
}

}  // namespace

bool wgrad_use_tc() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("PN2_WGRAD_IMPL");
        v = (e && strcmp(e, "tc") == 0) ? 1 : 0;
    }
    return v == 1;
}

bool wgrad_tc_supported(const WgradArgs& a) { return a.n <= 512 && a.kp % 32 == 0 && a.n % 8 == 0; }

int launch_wgrad_tc(const WgradArgs& a, cudaStream_t stream) {
    WgTc w;
    w.a = a;
    w.mt = (a.n + 127) / 128;
    w.npd = w.mt * 2;
    int kw_max = 512 / (2 * w.mt + 1) / 16 * 16;
    if (kw_max > 256) kw_max = 256;
    const int ky = (a.kp + kw_max - 1) / kw_max;
    w.kw = ((a.kp + ky - 1) / ky + 15) / 16 * 16;  // even split, multiple of 16, <= kw_max
    if (w.kw > kw_max) w.kw = kw_max;
    w.npx = (w.kw + 63) / 64;
    const int gy = (a.kp + w.kw - 1) / w.kw;
    const int cols = (2 * w.mt + 1) * w.kw;
    w.tmem_cols = 32;
    while (w.tmem_cols < cols) w.tmem_cols <<= 1;
    const size_t stage = (size_t)(2 * w.npd + 2 * w.npx) * kPanel;
    const size_t fixed = (size_t)2 * kPanel + (size_t)(3 * w.npd * 64 + 2 * w.npx * 64) * 4 + 256 + 1024;
    int nst = (int)((kWSmem - fixed) / stage);
    if (nst > kWMaxStages) nst = kWMaxStages;
    if (nst < 3) return fail_arg("pn2_mlp_gemm_wgrad", "stage does not fit shared memory");
    w.nst = nst;
    const size_t smem = fixed + nst * stage;
    static bool configured = false;
    if (!configured) {
        PN2_CHECK(cudaFuncSetAttribute(wgrad_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kWSmem),
                  "wgrad_tc: cudaFuncSetAttribute");
        PN2_CHECK(cudaFuncSetAttribute(wgrad_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kWSmem),
                  "wgrad_tc: cudaFuncSetAttribute");
        configured = true;
    }
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        sms = 148;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const long long stages = (a.rows + WR - 1) / WR;
    long long gx = sms / gy;
    if (gx < 1) gx = 1;
    if (gx > stages) gx = stages;
    if (a.in_scale)
        wgrad_tc_kernel<true><<<dim3((unsigned)gx, gy), kWThreads, smem, stream>>>(w);
    else
        wgrad_tc_kernel<false><<<dim3((unsigned)gx, gy), kWThreads, smem, stream>>>(w);
    PN2_CHECK_LAUNCH("wgrad_tc_kernel");
    return 0;
}

}  // namespace pn2
