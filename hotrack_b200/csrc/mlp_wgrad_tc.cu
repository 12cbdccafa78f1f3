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
// so dZ (bf16) and Y (fp16) reach the tensor core exactly as stored; only X passes through arithmetic (BatchNorm+ReLU
// of the producing layer, and a bf16 copy: one tcgen05.mma multiplies like with like).  G1 = dZ^T X', G2 = Y^T X' (MT
// 128-channel tiles each) and G3 = 1^T X' (a constant tile of ones as A operand) accumulate in TMEM over ALL rows of
// the CTA; the epilogue combines them and adds into dW with fp32 atomics.
//
// The reduction runs over ROWS: memory rows are K slices, which makes both operands MN-major -- and the tensor core is
// fed MN-major 16-bit tiles at a fraction of the K-major rate (measured: 0.13-0.3 us per M=128,K=16 instruction).  So
// the operands are TRANSPOSED on the way in, shared memory to shared memory:
//
//   raw ring   [NR][16 rows][dZ n_c | Y n_c | X kw_c columns] row-major, pitch = an odd number of 16-byte units,
//              filled by cp.async, D = NR - 2 stages in flight
//   T ring     [NT] operand tiles in the UMMA K-major no-swizzle layout: 8x8 core matrices (8 channels x 8 rows,
//              128 contiguous bytes), a channel group's two K halves 128 B apart (LBO), channel groups 256 B apart (SBO)
//   producers  per 16x16 block: ldmatrix.x4.trans from the raw stage (the fragment now holds 8x8 blocks channel-major),
//              [X only: BN+ReLU per channel, fp16 + bf16 copies], stmatrix.x4 = four core matrices
//
//   8 producer warps  as above, fence.proxy.async, mbarrier arrive
//   1 MMA warp        one lane: per stage (2 MT + 1) tcgen05.mma (M = 128, N = kw, K = 16 rows); tcgen05.commit frees
//                     the T slot; a final commit publishes the accumulators
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
constexpr int kWEpiWarps = 4, kWMmaWarp = 4, kWProdWarps = 8;
constexpr int kWProdThreads = kWProdWarps * 32;
constexpr int kWThreads = (kWEpiWarps + 1 + kWProdWarps) * 32;
constexpr int kWSmem = 225 * 1024;
constexpr int kMaxPieces = 10;    // cp.async pieces per producer thread per stage

struct WgTc {
    WgradArgs a;
    int mt;         // 128-wide output-channel tiles
    int kw;         // input channels per CTA (multiple of 16, (2 mt + 1) * kw <= 512)
    int n_c;        // n rounded up to 16
    int nr, nt;     // raw ring depth, T ring depth
    int tmem_cols;  // power of two >= (2 mt + 1) * kw
};

__device__ __forceinline__ void stsm_x4(uint32_t addr, const uint32_t (&r)[4]) {
    asm volatile("stmatrix.sync.aligned.m8n8.x4.shared.b16 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3])
                 : "memory");
}
// K-major, no swizzle: LBO = distance of the two 8-element K halves, SBO = distance of 8-row groups
__device__ __forceinline__ uint64_t umma_desc_k_noswz(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) |
           ((uint64_t)1 << 46);
}
__device__ __forceinline__ void prod_bar() { asm volatile("bar.sync 2, 256;" ::: "memory"); }

template <bool AFFINE>
__global__ void __launch_bounds__(kWThreads, 1) wgrad_tc_kernel(const WgTc w) {
    const WgradArgs& p = w.a;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    unsigned char* base = smem_raw + (((raw_addr + 127u) & ~127u) - raw_addr);
    const int k0 = blockIdx.y * w.kw;                 // first input channel of this CTA
    const int kw_here = min(w.kw, p.kp - k0);         // multiple of 16
    const int n_pad = w.mt * 128;
    const int ppr = (2 * w.n_c + w.kw) >> 3;          // 16-byte pieces per raw row
    const int rp = ppr * 16 + 16;                     // raw row pitch in bytes (odd number of 16-byte units)
    const int raw_bytes = WR * rp;
    const int t_dz = 0, t_y = n_pad * 32, t_x16 = 2 * n_pad * 32, t_xbf = t_x16 + w.kw * 32;
    const int t_bytes = t_xbf + w.kw * 32;            // one T slot: 32 bytes per channel per operand
    unsigned char* sRaw = base;
    unsigned char* sT = sRaw + ((w.nr * raw_bytes + 127) & ~127);
    unsigned char* sOnes = sT + w.nt * t_bytes;       // 128 channels x 16 rows of fp16 ones
    float* sCo = reinterpret_cast<float*>(sOnes + 4096);  // [3][n_pad] cA cB cC, then [2][kw] scale shift
    uint64_t* bars = reinterpret_cast<uint64_t*>(sCo + 3 * n_pad + 2 * w.kw);
    uint64_t* full = bars;            // [nt]
    uint64_t* empty = bars + w.nt;    // [nt]
    uint64_t* done = bars + 2 * w.nt;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * w.nt + 1);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long stages = (p.rows + WR - 1) / WR;
    const long long mine = blockIdx.x < stages ? (stages - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    for (int i = tid; i < 3 * n_pad; i += kWThreads) {
        const int which = i / n_pad, c = i - which * n_pad;
        sCo[i] = c < p.n ? (which == 0 ? p.cA : (which == 1 ? p.cB : p.cC))[c] : 0.f;
    }
    for (int i = tid; i < 2 * w.kw; i += kWThreads) {
        const int which = i / w.kw, c = i - which * w.kw;
        sCo[3 * n_pad + i] = (AFFINE && k0 + c < p.kp) ? (which == 0 ? p.in_scale : p.in_shift)[k0 + c] : 0.f;
    }
    for (int i = tid; i < 4096 / 4; i += kWThreads) reinterpret_cast<uint32_t*>(sOnes)[i] = 0x3C003C00u;  // fp16 1.0 x 2
    // T tiles start zeroed: channel groups beyond n_c / kw_here are never written and must multiply as zero
    for (int i = tid; i < w.nt * t_bytes / 16; i += kWThreads) reinterpret_cast<uint4*>(sT)[i] = make_uint4(0u, 0u, 0u, 0u);
    if (tid == 0) {
        for (int i = 0; i < w.nt; ++i) {
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
    fence_proxy_async();  // the ones tile and the zeroed T tiles (generic-proxy stores) are read by the tensor core
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp > kWMmaWarp) {
        // ================================ producers ================================
        const int pt = tid - (kWMmaWarp + 1) * 32;   // 0..255
        const int pw = pt >> 5;                       // producer warp 0..7
        // this thread's cp.async pieces of a stage: piece q = pt + 256 i  ->  (row, piece column), fixed for the kernel
        int prow[kMaxPieces], ppc[kMaxPieces];
        const int npieces = WR * ppr;
#pragma unroll
        for (int i = 0; i < kMaxPieces; ++i) {
            const int q = pt + 256 * i;
            prow[i] = q < npieces ? q / ppr : -1;
            ppc[i] = q < npieces ? q - prow[i] * ppr : 0;
        }
        const int ncp = w.n_c >> 3;                   // pieces of dZ (= of Y) per row
        const uint32_t raw0 = smem_u32(sRaw), t0 = smem_u32(sT);
        const int D = w.nr - 2;  // a raw slot is refilled two iterations after it was transposed: one producer barrier in between
        const int units = (2 * w.n_c + w.kw) >> 4;    // 16 x 16 blocks per stage
        const int u_y = w.n_c >> 4, u_x = 2 * u_y;    // first Y unit, first X unit
        const int units_x_valid = (kw_here + 15) >> 4;
        long long i_s = blockIdx.x, p_s = blockIdx.x;
        int i_slot = 0, p_slot = 0, t_slot = 0;
        uint32_t t_phase = 0;
        for (long long c = 0; c < mine + D; ++c) {
            if (c < mine) {
                const uint32_t st = raw0 + i_slot * raw_bytes;
                const long long row0 = i_s * WR;
#pragma unroll
                for (int i = 0; i < kMaxPieces; ++i) {
                    if (prow[i] >= 0) {
                        const long long row = row0 + prow[i];
                        const int pc = ppc[i];
                        const void* src;
                        bool ok = row < p.rows;
                        if (pc < ncp) {
                            ok = ok && pc * 8 < p.n;
                            src = p.dz + (ok ? row * p.dz_ld + pc * 8 : 0);
                        } else if (pc < 2 * ncp) {
                            ok = ok && (pc - ncp) * 8 < p.n;
                            src = p.y + (ok ? row * p.y_ld + (pc - ncp) * 8 : 0);
                        } else {
                            ok = ok && (pc - 2 * ncp) * 8 < kw_here;
                            src = p.x + (ok ? row * p.x_ld + k0 + (pc - 2 * ncp) * 8 : 0);
                        }
                        cp_async16_s(st + prow[i] * rp + pc * 16, src, ok ? 16 : 0);
                    }
                }
                i_s += gridDim.x;
                if (++i_slot == w.nr) i_slot = 0;
            }
            cp_async_commit();
            if (c >= D) {
                switch (D) {  // this thread's pieces of stage c - D have landed
                    case 1: cp_wait<1>(); break;
                    case 2: cp_wait<2>(); break;
                    case 3: cp_wait<3>(); break;
                    case 4: cp_wait<4>(); break;
                    case 5: cp_wait<5>(); break;
                    default: cp_wait<6>(); break;
                }
                prod_bar();                                  // ... and everybody else's
                mbar_wait(&empty[t_slot], t_phase ^ 1);      // the MMAs that read this T slot have completed
                const uint32_t rs = raw0 + p_slot * raw_bytes;
                const uint32_t ts = t0 + t_slot * t_bytes;
                const long long srow0 = p_s * WR;
                // ldmatrix lane address: matrix q = lane / 8 -> rows (q & 1) * 8 + lane % 8, columns + (q >> 1) * 8
                const uint32_t ld_off = (uint32_t)(((lane >> 3) & 1) * 8 + (lane & 7)) * rp + (lane >> 4) * 16;
                // stmatrix lane address: core matrix of q: channel group + (q >> 1), K half q & 1, row lane % 8
                const uint32_t st_off = (uint32_t)(lane >> 4) * 256 + ((lane >> 3) & 1) * 128 + (lane & 7) * 16;
                for (int u = pw; u < units; u += kWProdWarps) {
                    if (u >= u_x && u - u_x >= units_x_valid) break;  // padding of the X block
                    uint32_t r[4];
                    ldsm_x4_trans(r, rs + u * 32 + ld_off);
                    if (u < u_x) {  // dZ / Y: straight through
                        const uint32_t dst = ts + (u < u_y ? t_dz + u * 512 : t_y + (u - u_y) * 512);
                        stsm_x4(dst + st_off, r);
                    } else {
                        const int ux = u - u_x;
                        // fragment of matrix q: channel ux*16 + (q >> 1)*8 + lane/4, rows (q & 1)*8 + 2*(lane%4) + {0,1}
                        uint32_t r16[4], rbf[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int ch = ux * 16 + (q >> 1) * 8 + (lane >> 2);
                            const long long row = srow0 + (q & 1) * 8 + 2 * (lane & 3);
                            float2 v = h2_to_f2(r[q]);
                            if (AFFINE) {
                                const float sc = sCo[3 * n_pad + ch], sh = sCo[3 * n_pad + w.kw + ch];
                                v.x = fmaxf(fmaf(v.x, sc, sh), 0.f);
                                v.y = fmaxf(fmaf(v.y, sc, sh), 0.f);
                            }
                            if (row >= p.rows) v.x = 0.f;          // zero-filled rows must stay zero through the ReLU shift
                            if (row + 1 >= p.rows) v.y = 0.f;
                            r16[q] = f2_to_h2(v.x, v.y);
                            rbf[q] = f2_to_bf2(v.x, v.y);
                        }
                        stsm_x4(ts + t_x16 + ux * 512 + st_off, r16);
                        stsm_x4(ts + t_xbf + ux * 512 + st_off, rbf);
                    }
                }
                fence_proxy_async();
                mbar_arrive(&full[t_slot]);
                p_s += gridDim.x;
                if (++p_slot == w.nr) p_slot = 0;
                if (++t_slot == w.nt) { t_slot = 0; t_phase ^= 1; }
            }
        }
    } else if (warp == kWMmaWarp) {
        // ================================ MMA issuer ================================
        // A = dZ^T / Y^T / 1^T (M = output channels), B = X'^T (N = input channels): K-major after the transposition
        const uint32_t idesc_bf = umma_idesc2(1u, 1u, false, false, 128, kw_here);
        const uint32_t idesc_h = umma_idesc2(0u, 0u, false, false, 128, kw_here);
        const uint32_t g2_col = w.mt * w.kw, g3_col = 2 * w.mt * w.kw;
        const uint64_t ones_desc = umma_desc_k_noswz(smem_u32(sOnes));
        int slot = 0;
        uint32_t phase = 0;
        for (long long c = 0; c < mine; ++c) {
            mbar_wait(&full[slot], phase);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t ts = smem_u32(sT + slot * t_bytes);
                const uint64_t x16 = umma_desc_k_noswz(ts + t_x16), xbf = umma_desc_k_noswz(ts + t_xbf);
                const uint32_t acc = c != 0;
                for (int m = 0; m < w.mt; ++m) {
                    umma_f16(tmem_base + m * w.kw, umma_desc_k_noswz(ts + t_dz + m * 4096), xbf, idesc_bf, acc);
                    umma_f16(tmem_base + g2_col + m * w.kw, umma_desc_k_noswz(ts + t_y + m * 4096), x16, idesc_h, acc);
                }
                umma_f16(tmem_base + g3_col, ones_desc, x16, idesc_h, acc);
                tc_commit(&empty[slot]);
                if (c == mine - 1) tc_commit(done);
            }
            __syncwarp();
            if (++slot == w.nt) { slot = 0; phase ^= 1; }
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
                    const float ca = sCo[n], cb = sCo[n_pad + n], cc = sCo[2 * n_pad + n];
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
    w.n_c = (a.n + 15) / 16 * 16;
    int kw_max = 512 / (2 * w.mt + 1) / 16 * 16;
    if (kw_max > 256) kw_max = 256;
    const int ky = (a.kp + kw_max - 1) / kw_max;
    w.kw = ((a.kp + ky - 1) / ky + 15) / 16 * 16;  // even split, multiple of 16, <= kw_max
    if (w.kw > kw_max) w.kw = kw_max;
    const int gy = (a.kp + w.kw - 1) / w.kw;
    const int cols = (2 * w.mt + 1) * w.kw;
    w.tmem_cols = 32;
    while (w.tmem_cols < cols) w.tmem_cols <<= 1;
    const int ppr = (2 * w.n_c + w.kw) / 8;
    if (WR * ppr > kMaxPieces * kWProdThreads) return fail_arg("pn2_mlp_gemm_wgrad", "row too wide for the tcgen05 kernel");
    const size_t raw = (size_t)WR * (ppr * 16 + 16);
    const size_t tb = (size_t)(2 * w.mt * 128 + 2 * w.kw) * 32;
    const size_t fixed = 4096 + (size_t)(3 * w.mt * 128 + 2 * w.kw) * 4 + 256 + 256;
    // T ring: 3 slots (the tensor core is at most two stages behind); raw ring: what is left, 3..8 slots
    w.nt = 3;
    long long nr = ((long long)kWSmem - (long long)fixed - (long long)w.nt * (long long)tb) / (long long)raw;
    if (nr > 8) nr = 8;
    if (nr < 3) {
        w.nt = 2;
        nr = ((long long)kWSmem - (long long)fixed - (long long)w.nt * (long long)tb) / (long long)raw;
        if (nr > 8) nr = 8;
        if (nr < 3) return fail_arg("pn2_mlp_gemm_wgrad", "stage does not fit shared memory");
    }
    w.nr = (int)nr;
    const size_t smem = fixed + ((w.nr * raw + 127) / 128 * 128) + w.nt * tb;
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
