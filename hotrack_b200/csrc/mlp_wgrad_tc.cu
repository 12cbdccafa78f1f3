// mlp_wgrad_tc.cu -- weight gradient of the grouped per-point MLP on tcgen05 tensor cores, sm_100a.
//
//     dW[n][k] += sum_r dY[r][n] * X'[r][k]      dY = cA*dZ + cB*Y + cC  (BatchNorm backward)
//                                                X' = relu(X*scale+shift) | X   (the layer's input)
//
// (the conv1x1 weight gradient cuDNN computes for reference pointnet_utils.py:399-403,458-460,505-507,577-580
// and backbones.py:131 in backward).
//
// The reduction runs over ROWS, so in memory both operands are MN-major (a memory row is one K slice) -- and the tensor
// core is fed MN-major 16-bit tiles at a fraction of the K-major rate (measured: 0.13-0.3 us per M=128,K=16
// instruction).  The operands are therefore TRANSPOSED on the way in, shared memory to shared memory, and since every
// element passes through a thread's registers in that pass anyway, the BatchNorm-backward combination of dZ and Y and the
// BatchNorm+ReLU of X are applied there too: ONE bf16 x bf16 tcgen05.mma per K step, accumulators resident in TMEM over
// all rows of the CTA.
//
//   raw ring   [NR] stages of WR = 64 rows x [dZ 128 | Y 128 | X kw] columns, row-major (pitch = an odd number of 16-byte
//              units: conflict-free ldmatrix)
//   T ring     [NT] slots of 4 sub-tiles (16 rows = one MMA K step each) in the UMMA K-major no-swizzle layout: 8x8 core
//              matrices (8 channels x 8 rows, 128 contiguous bytes), a channel group's two K halves 128 B apart (LBO),
//              channel groups 256 B apart (SBO)
//   4 loader warps    cp.async of the raw stages as far ahead as there are free slots; the copies report their own
//                     completion to the stage's mbarrier (cp.async.mbarrier.arrive.noinc): nobody who computes waits on
//                     memory it has just asked for
//   8 transposer warps  per 16x16 block: ldmatrix.x4.trans (the fragment now holds 8x8 blocks channel-major) of dZ and Y
//                     -> cA*dZ + cB*Y + cC in fp32 -> bf16 -> stmatrix.x4 = four core matrices; X blocks: BatchNorm+ReLU ->
//                     bf16 -> stmatrix; fence.proxy.async, mbarrier arrive; the raw slot goes back to the loaders
//   1 MMA warp        one lane: per stage 4 x tcgen05.mma (M = 128 output channels, N = kw input channels, K = 16 rows),
//                     tcgen05.commit frees the T slot; a final commit publishes the accumulator
//   4 epilogue warps  parked in a named barrier during the main loop (a spinning mbarrier wait took 15 % of the issue
//                     slots), then tcgen05.ld and vector reductions (red.global.add.v4.f32) into dW
//
// 64-row stages: one hand-shake per 4 K steps (the 16-row version of round 1 spent 2.4 us per stage on hand-shakes with
// the MMA warp idle 57 % of the time; 32-row stages measured slower than 64 here too).
// Grid: (row splits, [128-channel output tiles] x [input-channel tiles of kw <= 128]); one CTA per SM.
// Bound: HBM reads rows*(2n + k)*2 bytes (X is re-read once per output tile, dZ / Y once per input tile: L2 hits when the
// tiles of the same rows run side by side).
#include "mlp_gemm.cuh"
#include "tc_common.cuh"

#include <cstdlib>
#include <cstring>
#include <type_traits>

namespace pn2 {
namespace {

constexpr int WR = 64;            // rows per stage (32-row stages measured slower: the per-stage hand-shakes dominate)
constexpr int WSUB = WR / 16;     // MMA K steps per stage
constexpr int MT = 128;           // output channels per CTA = UMMA M
#ifndef PN2_WG_PROD_WARPS
#define PN2_WG_PROD_WARPS 8   // transposer warps
#endif
#ifndef PN2_WG_SPW
#define PN2_WG_SPW 2          // 16-row sub-tiles a transposer warp handles per channel block (independent chains in flight)
#endif
constexpr int kWEpiWarps = 4, kWMmaWarp = 4, kWProdWarps = PN2_WG_PROD_WARPS, kWLoadWarps = 4;
constexpr int kSPW = PN2_WG_SPW;
#ifndef PN2_WARP_ARRIVE
#define PN2_WARP_ARRIVE 0     // 1: one mbarrier arrival per transposer warp (after __syncwarp) instead of one per thread
#endif
constexpr int kFullCount = PN2_WARP_ARRIVE ? kWProdWarps : kWProdWarps * 32;
constexpr int kWProdThreads = kWProdWarps * 32;   // transposers: warps 5..12
constexpr int kWLoadThreads = kWLoadWarps * 32;   // loaders: warps 13..16
constexpr int kWLoadWarp0 = kWEpiWarps + 1 + kWProdWarps;
constexpr int kWThreads = (kWEpiWarps + 1 + kWProdWarps + kWLoadWarps) * 32;
constexpr int kWSmem = 225 * 1024;
constexpr int kMaxKw = 128;   // wider tiles leave no room for a 3-stage raw ring next to the two T slots

struct WgTc {
    WgradArgs a;
    int kw;         // input channels per CTA (multiple of 16)
    int gk;         // input-channel tiles; blockIdx.y = n_tile * gk + k_tile
    int nr, nt;     // raw ring depth, T ring depth
    int tmem_cols;  // power of two >= kw
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

template <bool AFFINE>
__global__ void __launch_bounds__(kWThreads, 1) wgrad_tc_kernel(const WgTc w) {
    const WgradArgs& p = w.a;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    unsigned char* base = smem_raw + (((raw_addr + 127u) & ~127u) - raw_addr);
    const int n0 = (blockIdx.y / w.gk) * MT;          // first output channel of this CTA
    const int k0 = (blockIdx.y % w.gk) * w.kw;        // first input channel
    const int n_here = min(MT, p.n - n0);             // multiple of 8
    const int nb = (n_here + 15) >> 4;                // 16-channel blocks of dZ / Y
    const int kw_here = min(w.kw, p.kp - k0);         // multiple of 16
    const int kb = kw_here >> 4;
    const int ppr = (2 * MT + w.kw) >> 3;             // 16-byte pieces per raw row: [dZ 16][Y 16][X kw/8]
    const int rp = ppr * 16 + 16;                     // raw row pitch in bytes (odd number of 16-byte units)
    const int raw_bytes = WR * rp;
    const int sub_dy = MT * 32, sub_x = w.kw * 32;    // bytes of one 16-row sub-tile: 32 per channel
    const int t_x = WSUB * sub_dy;                    // T slot = [4 x dY sub-tiles][4 x X' sub-tiles]
    const int t_bytes = t_x + WSUB * sub_x;
    unsigned char* sRaw = base;
    unsigned char* sT = sRaw + ((w.nr * raw_bytes + 127) & ~127);
    float* sCo = reinterpret_cast<float*>(sT + w.nt * t_bytes);  // [3][MT] cA cB cC, then [2][kw] scale shift
    uint64_t* bars = reinterpret_cast<uint64_t*>(sCo + 3 * MT + 2 * w.kw);
    uint64_t* full = bars;            // [nt]  transposers -> MMA
    uint64_t* empty = bars + w.nt;    // [nt]  MMA -> transposers
    uint64_t* done = bars + 2 * w.nt;
    uint64_t* rfull = done + 1;       // [nr]  loaders (cp.async completion) -> transposers
    uint64_t* rempty = rfull + w.nr;  // [nr]  transposers -> loaders
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(rempty + w.nr);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long stages = (p.rows + WR - 1) / WR;
    const long long mine = blockIdx.x < stages ? (stages - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    // on-chip setup first (it overlaps the predecessor kernel's tail), global reads after the programmatic-dependent-launch
    // wait (pn2_common.cuh)
    // T tiles start zeroed: channel groups beyond n_here / kw_here are never written and must multiply as zero
    for (int i = tid; i < w.nt * t_bytes / 16; i += kWThreads) reinterpret_cast<uint4*>(sT)[i] = make_uint4(0u, 0u, 0u, 0u);
    if (tid == 0) {
        for (int i = 0; i < w.nt; ++i) {
            mbar_init(&full[i], kFullCount);
            mbar_init(&empty[i], 1);
        }
        mbar_init(done, 1);
        for (int i = 0; i < w.nr; ++i) {
            mbar_init(&rfull[i], kWLoadThreads);
            mbar_init(&rempty[i], kWProdWarps);
        }
        mbar_fence_init();
    }
    if (warp == kWMmaWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)w.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    pdl_enter();
    for (int i = tid; i < 3 * MT; i += kWThreads) {
        const int which = i / MT, c = i - which * MT;
        sCo[i] = c < n_here ? (which == 0 ? p.cA : (which == 1 ? p.cB : p.cC))[n0 + c] : 0.f;
    }
    for (int i = tid; i < 2 * w.kw; i += kWThreads) {
        const int which = i / w.kw, c = i - which * w.kw;
        sCo[3 * MT + i] = (AFFINE && c < kw_here) ? (which == 0 ? p.in_scale : p.in_shift)[k0 + c] : 0.f;
    }
    fence_proxy_async();  // the zeroed T tiles (generic-proxy stores) are read by the tensor core
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp >= kWLoadWarp0) {
        // ================================ loaders ================================
        // cp.async of the raw stages, as far ahead as there are free slots; the copies report their own completion to the
        // stage's mbarrier (cp.async.mbarrier.arrive.noinc), these warps never wait for data.  Warp lw brings rows lw,
        // lw + 4, ...; a lane walks the USEFUL 16-byte pieces of a row -- the dZ and Y columns of this CTA's channel blocks
        // and the X columns of its input-channel tile.
        const int lw = warp - kWLoadWarp0;
        const uint32_t raw0 = smem_u32(sRaw);
        const int nn8 = nb * 2, upr = 2 * nn8 + 2 * kb;  // pieces per row: [dZ nn8][Y nn8][X 2 kb] <= 32 + 16
        // A lane owns the same (at most two) 16-byte pieces of every row: everything that does not depend on the stage is
        // worked out once, the per-row work is one 64-bit add, one 32-bit add and the copy (the first version recomputed
        // 64-bit row offsets per copy: ~590 instructions per stage and warp, which made the LOADERS the kernel's
        // bottleneck -- ncu: 2.5 us per 64-row stage whatever the transposers did)
        static_assert(2 * (MT / 8) + kMaxKw / 8 <= 64, "a lane owns at most two pieces of a raw row");
        const unsigned char* src[2];
        long long step[2];   // 4 rows further (the loader warps interleave rows), in bytes
        uint32_t cb[2];
        bool has[2], ok[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int j = lane + 32 * h;
            has[h] = j < upr;
            long long ldb = 0;
            src[h] = reinterpret_cast<const unsigned char*>(p.dz);
            cb[h] = 0; ok[h] = false;
            if (j < nn8) {
                cb[h] = j * 16; ok[h] = j * 8 < n_here; ldb = 2LL * p.dz_ld;
                src[h] = reinterpret_cast<const unsigned char*>(p.dz + n0 + j * 8);
            } else if (j < 2 * nn8) {
                const int jj = j - nn8;
                cb[h] = 256 + jj * 16; ok[h] = jj * 8 < n_here; ldb = 2LL * p.y_ld;
                src[h] = reinterpret_cast<const unsigned char*>(p.y + n0 + jj * 8);
            } else if (j < upr) {
                const int jj = j - 2 * nn8;
                cb[h] = 512 + jj * 16; ok[h] = true; ldb = 2LL * p.x_ld;
                src[h] = reinterpret_cast<const unsigned char*>(p.x + k0 + jj * 8);
            }
            step[h] = ldb * kWLoadWarps;
            src[h] += ldb * lw;  // this warp's first row of a stage
        }
        constexpr int kRowsPerWarp = WR / kWLoadWarps;
        const uint32_t dstep = (uint32_t)rp * kWLoadWarps;
        long long i_s = blockIdx.x;
        int slot = 0;
        uint32_t phase = 0;
        for (long long c = 0; c < mine; ++c) {
            mbar_wait_parked(&rempty[slot], phase ^ 1);  // the transposers have read this slot's previous stage
            const uint32_t st = raw0 + slot * raw_bytes + lw * rp;
            const long long row0 = i_s * WR;
            const bool full_stage = row0 + WR <= p.rows;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (!has[h]) continue;
                const unsigned char* g = src[h] + (step[h] / kWLoadWarps) * row0;
                uint32_t d = st + cb[h];
                if (ok[h] && full_stage) {
#pragma unroll
                    for (int r = 0; r < kRowsPerWarp; ++r) {
                        cp_async16_s(d, g, 16);
                        g += step[h]; d += dstep;
                    }
                } else {  // the last stage / a padded half block: rows past the end and dead columns are zero-filled
#pragma unroll 4
                    for (int r = 0; r < kRowsPerWarp; ++r) {
                        const bool in = ok[h] && row0 + lw + r * kWLoadWarps < p.rows;
                        cp_async16_s(d, in ? g : reinterpret_cast<const unsigned char*>(p.dz), in ? 16 : 0);
                        g += step[h]; d += dstep;
                    }
                }
            }
            cp_async_mbar_arrive_noinc(&rfull[slot]);
            i_s += gridDim.x;
            if (++slot == w.nr) { slot = 0; phase ^= 1; }
        }
        cp_async_wait_all();  // nothing of this thread's may be in flight when the CTA's shared memory goes away
    } else if (warp > kWMmaWarp) {
        // ================================ transposers ================================
        const int pt = tid - (kWMmaWarp + 1) * 32;   // 0..255
        const int pw = pt >> 5;                       // transposer warp 0..7
        const uint32_t raw0 = smem_u32(sRaw), t0 = smem_u32(sT);
        // a warp owns the sub-tile PAIR pw & 1 (2 x 16 rows) of channel blocks pw >> 1, + 4, ...: no per-unit index
        // arithmetic, the per-channel constants are fetched once per block, and the two sub-tiles are two independent
        // ldmatrix -> arithmetic -> stmatrix chains in flight per warp
        constexpr int kSubGroups = WSUB / kSPW;            // warps sharing a channel block
        constexpr int kBlkStep = kWProdWarps / kSubGroups;  // channel blocks in flight per stage
        static_assert(WSUB % kSPW == 0 && kWProdWarps % kSubGroups == 0, "warp -> (sub-tiles, block) map");
        const int sub0 = (pw % kSubGroups) * kSPW, blk0 = pw / kSubGroups;
        // ldmatrix lane address: matrix q = lane / 8 -> rows (q & 1) * 8 + lane % 8, columns + (q >> 1) * 8
        const uint32_t ld_off = (uint32_t)(sub0 * 16 + ((lane >> 3) & 1) * 8 + (lane & 7)) * rp + (lane >> 4) * 16;
        // stmatrix lane address: core matrix of q: channel group + (q >> 1), K half q & 1, row lane % 8
        const uint32_t st_off = (uint32_t)(lane >> 4) * 256 + ((lane >> 3) & 1) * 128 + (lane & 7) * 16;
        // fragment of matrix q: channel blk*16 + (q >> 1)*8 + lane/4, rows sub*16 + (q & 1)*8 + 2*(lane%4) + {0,1}
        const int ch_lo = lane >> 2, r_lo = sub0 * 16 + 2 * (lane & 3);
        long long p_s = blockIdx.x;
        int p_slot = 0, t_slot = 0;
        uint32_t p_phase = 0, t_phase = 0;
        for (long long c = 0; c < mine; ++c) {
            {
                mbar_wait(&rfull[p_slot], p_phase);          // every copy of this raw stage has landed
                mbar_wait(&empty[t_slot], t_phase ^ 1);      // the MMAs that read this T slot have completed
                const uint32_t rs = raw0 + p_slot * raw_bytes + ld_off;
                const uint32_t ts = t0 + t_slot * t_bytes + st_off;
                const long long srow0 = p_s * WR;
                // rows past the end were zero-filled and must contribute nothing (cC and the ReLU shift alone are not
                // zero): only the last stage can hold any
                const int valid = (int)min((long long)WR, p.rows - srow0) - r_lo;  // this thread's rows r_lo + {0,1,8,9} < valid?
                // the zero-fill masking of rows past the end costs ~20 % of the transposers' instructions: it is compiled
                // into the last stage's copy of the loops only
                auto do_stage = [&](auto tail_c) {
                    constexpr bool tail = decltype(tail_c)::value;
                    for (int blk = blk0; blk < nb; blk += kBlkStep) {
                        const int ch = blk * 16 + ch_lo;
                        const float ca0 = sCo[ch], cb0 = sCo[MT + ch], cc0 = sCo[2 * MT + ch];
                        const float ca1 = sCo[ch + 8], cb1 = sCo[MT + ch + 8], cc1 = sCo[2 * MT + ch + 8];
                        uint32_t rz[kSPW][4], ry[kSPW][4], o[kSPW][4];
    #pragma unroll
                        for (int s2 = 0; s2 < kSPW; ++s2) {
                            ldsm_x4_trans(rz[s2], rs + s2 * 16 * rp + blk * 32);
                            ldsm_x4_trans(ry[s2], rs + s2 * 16 * rp + 256 + blk * 32);
                        }
    #pragma unroll
                        for (int s2 = 0; s2 < kSPW; ++s2)
    #pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const float ca = (q >> 1) ? ca1 : ca0, cb = (q >> 1) ? cb1 : cb0, cc = (q >> 1) ? cc1 : cc0;
                                const float2 dz = bf2_to_f2(rz[s2][q]), yy = h2_to_f2(ry[s2][q]);
                                float v0 = fmaf(ca, dz.x, fmaf(cb, yy.x, cc)), v1 = fmaf(ca, dz.y, fmaf(cb, yy.y, cc));
                                if (tail) {
                                    if (s2 * 16 + (q & 1) * 8 >= valid) v0 = 0.f;
                                    if (s2 * 16 + (q & 1) * 8 + 1 >= valid) v1 = 0.f;
                                }
                                o[s2][q] = f2_to_bf2(v0, v1);
                            }
    #pragma unroll
                        for (int s2 = 0; s2 < kSPW; ++s2) stsm_x4(ts + (sub0 + s2) * sub_dy + blk * 512, o[s2]);
                    }
                    for (int blk = blk0; blk < kb; blk += kBlkStep) {
                        uint32_t r[kSPW][4], o[kSPW][4];
    #pragma unroll
                        for (int s2 = 0; s2 < kSPW; ++s2) ldsm_x4_trans(r[s2], rs + s2 * 16 * rp + 512 + blk * 32);
                        if (AFFINE) {
                            const int ch = blk * 16 + ch_lo;
                            const float sc0 = sCo[3 * MT + ch], sh0 = sCo[3 * MT + w.kw + ch];
                            const float sc1 = sCo[3 * MT + ch + 8], sh1 = sCo[3 * MT + w.kw + ch + 8];
    #pragma unroll
                            for (int s2 = 0; s2 < kSPW; ++s2)
    #pragma unroll
                                for (int q = 0; q < 4; ++q) {
                                    const float sc = (q >> 1) ? sc1 : sc0, sh = (q >> 1) ? sh1 : sh0;
                                    const float2 v = h2_to_f2(r[s2][q]);
                                    float v0 = fmaxf(fmaf(v.x, sc, sh), 0.f), v1 = fmaxf(fmaf(v.y, sc, sh), 0.f);
                                    if (tail) {
                                        if (s2 * 16 + (q & 1) * 8 >= valid) v0 = 0.f;
                                        if (s2 * 16 + (q & 1) * 8 + 1 >= valid) v1 = 0.f;
                                    }
                                    o[s2][q] = f2_to_bf2(v0, v1);
                                }
                        } else {
    #pragma unroll
                            for (int s2 = 0; s2 < kSPW; ++s2)
    #pragma unroll
                                for (int q = 0; q < 4; ++q) {
                                    const float2 v = h2_to_f2(r[s2][q]);
                                    o[s2][q] = f2_to_bf2(v.x, v.y);
                                }
                        }
    #pragma unroll
                        for (int s2 = 0; s2 < kSPW; ++s2) stsm_x4(ts + t_x + (sub0 + s2) * sub_x + blk * 512, o[s2]);
                    }
                };
                if (srow0 + WR > p.rows) do_stage(std::true_type{});
                else do_stage(std::false_type{});
                fence_proxy_async();
#if PN2_WARP_ARRIVE
                __syncwarp();  // every lane's stores and proxy fence are ordered before lane 0's arrival
                if (lane == 0) {
                    mbar_arrive(&full[t_slot]);
                    mbar_arrive(&rempty[p_slot]);
                }
#else
                mbar_arrive(&full[t_slot]);
                __syncwarp();
                if (lane == 0) mbar_arrive(&rempty[p_slot]);  // this warp has read everything it needs from the raw slot
#endif
                p_s += gridDim.x;
                if (++p_slot == w.nr) { p_slot = 0; p_phase ^= 1; }
                if (++t_slot == w.nt) { t_slot = 0; t_phase ^= 1; }
            }
        }
    } else if (warp == kWMmaWarp) {
        // ================================ MMA issuer ================================
        // A = dY^T (M = output channels), B = X'^T (N = input channels): K-major after the transposition
        const uint32_t idesc = umma_idesc2(1u, 1u, false, false, MT, kw_here);
        int slot = 0;
        uint32_t phase = 0;
        for (long long c = 0; c < mine; ++c) {
            mbar_wait_parked(&full[slot], phase);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t ts = smem_u32(sT + slot * t_bytes);
#pragma unroll
                for (int s = 0; s < WSUB; ++s)
                    umma_f16(tmem_base, umma_desc_k_noswz(ts + s * sub_dy), umma_desc_k_noswz(ts + t_x + s * sub_x), idesc,
                             (c | s) != 0);
                tc_commit(&empty[slot]);
                if (c == mine - 1) tc_commit(done);
            }
            __syncwarp();
            if (++slot == w.nt) { slot = 0; phase ^= 1; }
        }
        // the epilogue warps sleep in a named barrier for the whole main loop (a spinning mbarrier wait would take issue
        // slots from the producers); this warp wakes them once the last MMA has completed
        if (mine > 0) mbar_wait(done, 0);
        tc_fence_after();
        tc_fence_before();  // order the completed MMAs before the barrier the epilogue warps leave through
        asm volatile("bar.sync 3, 160;" ::: "memory");
    } else {
        // ================================ epilogue ================================
        asm volatile("bar.sync 3, 160;" ::: "memory");
        tc_fence_after();
        const bool vec4 = (p.dw_ld & 3) == 0 && (reinterpret_cast<uintptr_t>(p.dw) & 15) == 0;  // 16-byte aligned rows
        const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
        const int n = n0 + warp * 32 + lane;  // output channel of this thread (TMEM lane)
        if (mine > 0 && warp * 32 < n_here) {  // warp-uniform
            for (int c16 = 0; c16 < kw_here; c16 += 16) {
                uint32_t g[32];
                tmem_ld16(tmem_base + lane_base + c16, g);
                if (n < p.n) {
                    float* dst = p.dw + (size_t)n * p.dw_ld + k0 + c16;
                    if (vec4 && k0 + c16 + 16 <= p.k_true) {
#pragma unroll
                        for (int e = 0; e < 16; e += 4)
                            red_add_v4(dst + e, __uint_as_float(g[e]), __uint_as_float(g[e + 1]), __uint_as_float(g[e + 2]),
                                       __uint_as_float(g[e + 3]));
                    } else {
#pragma unroll
                        for (int e = 0; e < 16; ++e)
                            if (k0 + c16 + e < p.k_true && g[e] != 0u) atomicAdd(dst + e, __uint_as_float(g[e]));
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
        const char* e = getenv("PN2_WGRAD_IMPL");  // tc (default) | mma: the warp-level mma.sync kernel of mlp_gemm.cu
        v = (e && strcmp(e, "mma") == 0) ? 0 : 1;
    }
    return v == 1;
}

bool wgrad_tc_supported(const WgradArgs& a) { return a.kp % 16 == 0 && a.n % 8 == 0; }

int launch_wgrad_tc(const WgradArgs& a, cudaStream_t stream) {
    WgTc w;
    w.a = a;
    const int gn = (a.n + MT - 1) / MT;
    w.gk = (a.kp + kMaxKw - 1) / kMaxKw;
    w.kw = ((a.kp + w.gk - 1) / w.gk + 15) / 16 * 16;  // even split, multiple of 16, <= kMaxKw
    w.gk = (a.kp + w.kw - 1) / w.kw;
    const int gy = gn * w.gk;
    w.tmem_cols = 32;
    while (w.tmem_cols < w.kw) w.tmem_cols <<= 1;
    const int ppr = (2 * MT + w.kw) / 8;
    const size_t raw = (size_t)WR * (ppr * 16 + 16);
    const size_t tb = (size_t)WSUB * (MT + w.kw) * 32;
    const size_t fixed = (size_t)(3 * MT + 2 * w.kw) * 4 + 256 + 256;  // constants, barriers (<= 19 x 8 bytes), alignment slack
    // T ring: 2 slots (the tensor core is at most one stage behind); raw ring: what is left, 2..6 slots (the loaders run
    // ahead by as many stages as there are free slots)
    w.nt = 2;
    long long nr = ((long long)kWSmem - (long long)fixed - (long long)w.nt * (long long)tb) / (long long)raw;
    if (nr > 6) nr = 6;
    if (nr < 2) return fail_arg("pn2_mlp_gemm_wgrad", "stage does not fit shared memory");
    w.nr = (int)nr;
    const size_t smem = fixed + ((w.nr * raw + 127) / 128 * 128) + w.nt * tb;
    int dev = 0;
    cudaGetDevice(&dev);
    static bool configured[64] = {};
    static int sms[64] = {};
    if (dev < 0 || dev >= 64) dev = 0;
    if (!configured[dev]) {  // per device: the attribute belongs to the device's copy of the function
        PN2_CHECK(cudaFuncSetAttribute(wgrad_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kWSmem),
                  "wgrad_tc: cudaFuncSetAttribute");
        PN2_CHECK(cudaFuncSetAttribute(wgrad_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kWSmem),
                  "wgrad_tc: cudaFuncSetAttribute");
        sms[dev] = 148;
        cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev);
        configured[dev] = true;
    }
    const long long stages = (a.rows + WR - 1) / WR;
    long long gx = sms[dev] / gy;
    if (gx < 1) gx = 1;
    // Few rows: a CTA's fixed cost (prologue, pipeline fill, drain, the red.add epilogue over its whole tile) is worth ~8
    // stages of streaming, and every extra row split adds a full tile of atomics.  At least 8 (PN2_WG_MIN_STAGES) stages per CTA (measured: step 3.24 ms at 1, 3.20 at 4, 3.19 at 8, 3.18 at 16):
    // small layers then occupy a fraction of the SMs and leave the rest to the backward chain running beside them.
    static int min_stages = -1;
    if (min_stages < 0) {
        const char* e = getenv("PN2_WG_MIN_STAGES");
        min_stages = e ? atoi(e) : 8;
        if (min_stages < 1) min_stages = 1;
    }
    if (gx > stages / min_stages) gx = stages / min_stages;
    if (gx < 1) gx = 1;
    if (a.in_scale)
        launch_k(wgrad_tc_kernel<true>, dim3((unsigned)gx, gy), dim3(kWThreads), smem, stream, w);
    else
        launch_k(wgrad_tc_kernel<false>, dim3((unsigned)gx, gy), dim3(kWThreads), smem, stream, w);
    PN2_CHECK_LAUNCH("wgrad_tc_kernel");
    return 0;
}

}  // namespace pn2
