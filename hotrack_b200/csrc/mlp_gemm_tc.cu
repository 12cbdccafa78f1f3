// mlp_gemm_tc.cu -- the row-matrix GEMM of the grouped per-point MLP on the 5th-generation tensor cores
// (tcgen05.mma, accumulators in TMEM), sm_100a.  Same contract as mlp_gemm.cu's gemm_rows_kernel:
//
//     C[R][n] = A'[R][k] * B[n][k]^T          A' = A | relu(A*scale+shift) | cA*dZ + cB*Y + cC
//     epilogue: - centre[n], 16-bit rounding, column sums (BatchNorm statistics) or, with MASK,
//               ReLU mask of the previous layer + the two BatchNorm-backward sums
//
// replacing the conv1x1 -> BatchNorm -> ReLU launches of reference pointnet_utils.py:399-403,458-460,
// 505-507,577-580 and backbones.py:131-132 (and their backward).
//
// Why not TMA for A: the A operand is never the matrix in memory -- BatchNorm+ReLU of the producing
// layer (forward) or the BatchNorm-backward combination of two matrices (backward) is applied on the
// way in, so a thread has to touch every element between HBM and the tensor core.  The kernel is
// therefore warp-specialised around that:
//
//   8 producer warps  cp.async 16-byte pieces of the A (and B) chunk straight into the UMMA canonical
//                     K-major SWIZZLE_128B layout (row pitch 128 B, piece index XOR row%8), D chunks
//                     in flight per thread; when a chunk has landed each thread transforms ITS pieces in
//                     place (fp32 math, one rounding), fence.proxy.async, arrives on the stage's mbarrier
//   1 MMA warp        one elected lane: waits the stage, issues 4 x tcgen05.mma (M=128, N=BN, K=16) per
//                     64-wide chunk, tcgen05.commit -> frees the stage / publishes the accumulator
//   8 epilogue warps  (two per 32-lane TMEM quarter, half the columns each) tcgen05.ld -> - centre -> 16-bit tile in shared memory ->
//                     16-byte coalesced row stores + per-thread column sums (same scheme as the
//                     mma.sync kernel), two TMEM accumulator stages so tile i+1 is multiplied while
//                     tile i is written out
//
// Bound: HBM traffic of the row matrices (rows*(k+n)*2 bytes, + the masks backward); per 128x128 output
// tile the tensor pipe needs ~0.1 us, the producers ~0.4 us of issue slots, the epilogue ~0.35 us.
#include "mlp_gemm.cuh"
#include "tc_common.cuh"

#include <cstdlib>
#include <cstring>

namespace pn2 {
namespace {

constexpr int TM = 128;          // rows per tile = UMMA M
constexpr int TK = 64;           // K chunk: 64 x 16-bit = one 128-byte swizzle row
constexpr int kEpiWarps = 8;     // warps 0..7: TMEM lane quarter = warp % 4, column half = warp / 4
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kMmaWarp = 8;
constexpr int kProdWarps = 8;    // warps 9..16
constexpr int kProdThreads = kProdWarps * 32;
constexpr int kTcThreads = (kEpiWarps + 1 + kProdWarps) * 32;
constexpr int kMaxK = 1024;
constexpr int kMaxKSplit = 256;  // AFFINE + SPLIT: layers > 0 of the SA stacks (kdim <= 128 on this network)
constexpr int kSmemBudget = 225 * 1024;
#ifndef PN2_WARP_ARRIVE
#define PN2_WARP_ARRIVE 0     // 1: transforming producers arrive once per warp (after __syncwarp) instead of once per thread
#endif

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// UMMA shared-memory descriptor of a K-major SWIZZLE_128B operand tile (rows x 64 16-bit elements, 128-byte row
// pitch, 1024-byte aligned): start address >> 4 in [0,14), LBO (ignored for swizzled K-major) = 1 in [16,30),
// SBO = 1024 B (one 8-row group) >> 4 in [32,46), descriptor version 1 in [46,48), layout SWIZZLE_128B = 2 in [61,64).
__device__ __forceinline__ uint64_t umma_desc_k128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

// stage = [A0 16 KB][A1 16 KB, BNBWD / SPLIT only][B BN*128 B][B1, BNBWD / SPLIT only]
//
// SPLIT (forward only): every operand is a pair of fp16 planes (hi, lo) with value = hi + lo -- 22 significand bits --
// and the product is evaluated as A_hi*B_hi + A_lo*B_hi + A_hi*B_lo into the one fp32 accumulator (the lo*lo term is
// below fp32 resolution); the output is written as such a pair too.  Used for the stacks in front of the backbone's
// ill-conditioned spot (FP3 normalises a broadcast global feature: every rounding before it is amplified ~40x by the
// end of the network), where fp16's 11 bits are not enough for the 1e-2 parity bar.
template <int BN, int AMODE, bool SPLIT = false, bool POOL = false>
struct TcCfg {
    static constexpr int kABytes = TM * 128;
    static constexpr bool kTwo = AMODE == A_BNBWD || SPLIT;
    static constexpr int kNA = kTwo ? 2 : 1;
    static constexpr int kNB = kNA;                            // BNBWD: dZ x W_A and Y x W_B accumulate into one tile
    static constexpr int kBBytes = BN * 128;
    static constexpr int kStage = kABytes * kNA + kBBytes * kNB;
    static constexpr int kEpiMax = SPLIT ? 32 : (AMODE == A_BNBWD ? 64 : 128);  // two-operand stages are twice as large: smaller staging tile
    static constexpr int kEpiBN = BN < kEpiMax ? BN : kEpiMax;  // columns per epilogue pass
    static constexpr int kCLD = kEpiBN + 8;                    // 16-bit elements per sC row
    static constexpr int kSC = TM * kCLD * 2 * (SPLIT ? 2 : 1);  // SPLIT: hi tile + lo tile
    static constexpr int kNCoef = AMODE == A_AFFINE ? 2 : 0;
    static constexpr int kCoefK = SPLIT ? kMaxKSplit : kMaxK;
    static constexpr int kPool = POOL ? 9 * BN * 4 : 0;  // epilogue pooling: [BN] sign flags + [4][BN] values + [4][BN] rows
    static constexpr int kFixed = kSC + kNCoef * kCoefK * 4 + 5 * BN * 4 + kPool + 256 + 1024;  // + barriers + alignment slack
    static constexpr int kNstRaw = (kSmemBudget - kFixed) / kStage;
    static constexpr int kNst = kNstRaw > 6 ? 6 : kNstRaw;
    // chunks in flight per producer thread.  NST - 2, not NST - 1: refilling a slot waits for the MMAs of the chunk that
    // used it NST chunks ago; with D = NST - 1 that chunk was published only one iteration earlier and the MMA round
    // trip (mbarrier wake-up, tcgen05.mma, commit) would sit on the producers' critical path every iteration
    static constexpr int kDist = kNst - 2;
    static_assert(kNst >= 3, "not enough shared memory for a three-stage ring");
};

// POOL: the max-pool-in-the-epilogue variant (pn2_mlp_gemm_fwd[_bn]_pool).  A template parameter, not a run-time test of
// p.pool_k: the pooling code is a third of the kernel's instructions, and with it compiled in the hot path of the ordinary
// variants no longer fits the instruction cache (ncu: stall_no_inst on ~15 % of the epilogue's samples).
template <int BN, int AMODE, bool MASK, bool SPLIT = false, bool POOL = false>
__global__ void __launch_bounds__(kTcThreads, 1) gemm_tc_kernel(const GemmArgs p) {
    using Cfg = TcCfg<BN, AMODE, SPLIT, POOL>;
    static_assert(!POOL || AMODE != A_BNBWD, "pooling is a forward epilogue");
    static_assert(!SPLIT || (AMODE != A_BNBWD && !MASK), "SPLIT is a forward mode");
    constexpr bool TWO = Cfg::kTwo;
    constexpr int NST = Cfg::kNst, D = Cfg::kDist;
    constexpr int EBN = Cfg::kEpiBN, CLD = Cfg::kCLD, NH = BN / EBN;
    constexpr int CPR = EBN / 8;          // 16-byte pieces per sC row
    constexpr int RPP = kEpiThreads / CPR;  // rows per epilogue pass
    constexpr int WC = EBN / 2;           // columns of a pass-1 warp (two warps share a 32-lane quarter)
    constexpr int LDW = WC < 32 ? WC : 32;  // columns per tcgen05.ld
    constexpr int ACC = BN;                 // TMEM columns per accumulator stage
    constexpr int PASSES = TM / RPP;
    constexpr bool FWD = AMODE != A_BNBWD;
    constexpr int NCOEF = Cfg::kNCoef;

    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    unsigned char* base = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);  // SWIZZLE_128B tiles: 1024-byte aligned
    unsigned char* sStage = base;
    uint16_t* sC = reinterpret_cast<uint16_t*>(base + NST * Cfg::kStage);
    float* sCoef = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(sC) + Cfg::kSC);
    const int kpad = (p.kdim + TK - 1) / TK * TK;
    float* sPrev = sCoef + NCOEF * kpad;        // [4][BN], MASK only
    float* sCen = sPrev + 4 * BN;               // [BN]
    uint32_t* sNeg = reinterpret_cast<uint32_t*>(sCen + BN);   // [BN] pooling: all-ones where gamma < 0 (the minimum is wanted)
    uint32_t* sPoolV = sNeg + (POOL ? BN : 0);                   // [4][BN] per-quarter extremes (ordered-uint form), pool_k > 32
    int* sPoolA = reinterpret_cast<int*>(sPoolV + (POOL ? 4 * BN : 0));  // [4][BN] their rows
    uint64_t* bars = reinterpret_cast<uint64_t*>(sPoolA + (POOL ? 4 * BN : 0));  // 8-byte aligned: every size above is a multiple of 8
    uint64_t* full = bars;                      // [NST] producers -> MMA
    uint64_t* empty = bars + NST;               // [NST] MMA -> producers
    uint64_t* tfull = bars + 2 * NST;           // [2]   MMA -> epilogue
    uint64_t* tempty = bars + 2 * NST + 2;      // [2]   epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NST + 4);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n0 = blockIdx.y * BN;
    const int KT = kpad / TK;
    const long long tiles = (p.rows + TM - 1) / TM;
    const long long my_tiles = blockIdx.x < tiles ? (tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    // ---- one-time setup.  On-chip part first (barriers, TMEM): it overlaps the predecessor kernel's tail; everything that
    // reads global memory comes after the programmatic-dependent-launch wait (pn2_common.cuh)
    if (tid == 0) {
        for (int i = 0; i < NST; ++i) {
            mbar_init(&full[i], (PN2_WARP_ARRIVE && AMODE == A_AFFINE) ? kProdWarps : kProdThreads);
            mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull[i], 1);
            mbar_init(&tempty[i], kEpiThreads);
        }
        mbar_fence_init();
    }
    constexpr uint32_t kTmemCols = 2 * ACC < 32 ? 32 : 2 * ACC;  // power of two: BN in {32,64,128,256}
    if (warp == kMmaWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    pdl_enter();
    for (int i = tid; i < NCOEF * kpad; i += kTcThreads) {
        const int which = i / kpad, c = i - which * kpad;
        const float* src = which == 0 ? p.c0 : (which == 1 ? p.c1 : p.c2);
        sCoef[i] = c < p.kdim ? src[c] : 0.f;
    }
    for (int i = tid; i < BN; i += kTcThreads) sCen[i] = (p.center && n0 + i < p.n) ? p.center[n0 + i] : 0.f;
    if (POOL)
        for (int i = tid; i < BN; i += kTcThreads) sNeg[i] = (n0 + i < p.n && p.pool_gamma[n0 + i] < 0.f) ? 0xFFFFFFFFu : 0u;
    if (MASK) {
        for (int i = tid; i < 4 * BN; i += kTcThreads) {
            const int which = i / BN, c = n0 + (i - which * BN);
            const float* src = which == 0 ? p.p_scale : (which == 1 ? p.p_shift : (which == 2 ? p.p_mean : p.p_rstd));
            sPrev[i] = c < p.n ? src[c] : 0.f;
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp > kMmaWarp) {
        // ================================ producers ================================
        const int pt = tid - (kMmaWarp + 1) * 32;  // 0..255
        const int pj = pt & 7;                      // 16-byte piece of the 128-byte row
        const int pr = pt >> 3;                     // rows pr + 32*i
        // row r = pr + 32*i of a chunk lives at r*128 + ((pj ^ (r & 7)) << 4): r & 7 == pr & 7, so the 4 (8) pieces of a
        // thread are 4096 bytes apart
        const uint32_t poff = pr * 128 + ((pj ^ (pr & 7)) << 4);
        const uint32_t stage0 = smem_u32(sStage);
        const size_t a0_step = (size_t)32 * p.a0_ld, a1_step = (size_t)32 * p.a1_ld, b_step = (size_t)32 * p.kdim;
        const int n_left = p.n - n0 - pr;           // B rows pr + 32*i < n_left are real
        const long long total = my_tiles * KT;
        long long i_tile = blockIdx.x, p_tile = blockIdx.x;
        int i_kc = 0, i_slot = 0, p_kc = 0, p_slot = 0;
        uint32_t i_phase = 0;
        // Raw operands (forward layer 0, every backward GEMM): nothing happens between landing and multiplying, so the
        // copies report their own completion to the stage barrier (cp.async.mbarrier.arrive.noinc) and the producers
        // run ahead as far as there are free slots.  AFFINE: the thread comes back D chunks later to transform.
        constexpr bool RAW = AMODE != A_AFFINE;
        constexpr int LAG = RAW ? 0 : D;
        for (long long c = 0; c < total + LAG; ++c) {
            if (c < total) {
                mbar_wait(&empty[i_slot], i_phase ^ 1);  // the MMAs that read this slot NST chunks ago have completed
                const uint32_t st = stage0 + i_slot * Cfg::kStage;
                const int kcol = i_kc * TK + pj * 8;
                const bool kok = kcol < p.kdim;
                const long long row0 = i_tile * TM + pr;
                const int rows_left = (int)min((long long)TM, p.rows - i_tile * TM) - pr;  // rows pr + 32*i < rows_left are real
                const uint16_t* a0p = p.a0 + row0 * p.a0_ld + kcol;
                const uint16_t* a1p = TWO ? p.a1 + row0 * p.a1_ld + kcol : nullptr;
#pragma unroll
                for (int i = 0; i < TM / 32; ++i) {
                    const bool ok = kok && 32 * i < rows_left;
                    cp_async16_s(st + poff + i * 4096, ok ? a0p + (size_t)i * a0_step : p.a0, ok ? 16 : 0);
                    if (TWO)
                        cp_async16_s(st + Cfg::kABytes + poff + i * 4096, ok ? a1p + (size_t)i * a1_step : p.a1, ok ? 16 : 0);
                }
                const uint32_t sb = st + Cfg::kABytes * Cfg::kNA;
                const uint16_t* bp = p.b + (size_t)(n0 + pr) * p.kdim + kcol;
#pragma unroll
                for (int i = 0; i < BN / 32; ++i) {
                    const bool ok = kok && 32 * i < n_left;
                    cp_async16_s(sb + poff + i * 4096, ok ? bp + (size_t)i * b_step : p.b, ok ? 16 : 0);
                    if (TWO)
                        cp_async16_s(sb + Cfg::kBBytes + poff + i * 4096, ok ? p.b1 + (bp - p.b) + (size_t)i * b_step : p.b1,
                                     ok ? 16 : 0);
                }
                if (RAW) cp_async_mbar_arrive_noinc(&full[i_slot]);
                if (++i_kc == KT) { i_kc = 0; i_tile += gridDim.x; }
                if (++i_slot == NST) { i_slot = 0; i_phase ^= 1; }
            }
            if (RAW) continue;
            cp_async_commit();  // always: the group count stays in step with c
            if (c >= D) {
                cp_wait<D>();   // this thread's pieces of chunk c - D have landed
                if (AMODE == A_AFFINE) {
                    unsigned char* st = sStage + p_slot * Cfg::kStage + poff;
                    const int cc = p_kc * TK + pj * 8;
                    const float4 ka0 = *reinterpret_cast<const float4*>(&sCoef[cc]);
                    const float4 ka1 = *reinterpret_cast<const float4*>(&sCoef[cc + 4]);
                    const float4 kb0 = *reinterpret_cast<const float4*>(&sCoef[kpad + cc]);
                    const float4 kb1 = *reinterpret_cast<const float4*>(&sCoef[kpad + cc + 4]);
                    const float k0[8] = {ka0.x, ka0.y, ka0.z, ka0.w, ka1.x, ka1.y, ka1.z, ka1.w};
                    const float k1[8] = {kb0.x, kb0.y, kb0.z, kb0.w, kb1.x, kb1.y, kb1.z, kb1.w};
#pragma unroll
                    for (int i = 0; i < TM / 32; ++i) {
                        uint4* slot = reinterpret_cast<uint4*>(st + i * 4096);
                        const uint4 q0 = *slot;
                        const uint32_t* x0 = reinterpret_cast<const uint32_t*>(&q0);
                        uint4 v;
                        uint32_t* o = reinterpret_cast<uint32_t*>(&v);
                        if (SPLIT) {  // value = hi + lo; the activation is split again after BatchNorm + ReLU
                            uint4* slot1 = reinterpret_cast<uint4*>(st + Cfg::kABytes + i * 4096);
                            const uint4 q1 = *slot1;
                            const uint32_t* x1 = reinterpret_cast<const uint32_t*>(&q1);
                            uint4 vl;
                            uint32_t* ol = reinterpret_cast<uint32_t*>(&vl);
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                // the ReLU decision is taken on the hi plane alone -- the plane the backward kernels
                                // read their masks from -- so forward and backward always agree on it (an element
                                // within 2^-12 of the kink passes its tiny negative value instead of 0)
                                const float2 a = h2_to_f2(x0[e]), l = h2_to_f2(x1[e]);
                                const float r0 = fmaf(a.x, k0[2 * e], k1[2 * e]) > 0.f ? fmaf(a.x + l.x, k0[2 * e], k1[2 * e]) : 0.f;
                                const float r1 = fmaf(a.y, k0[2 * e + 1], k1[2 * e + 1]) > 0.f
                                                     ? fmaf(a.y + l.y, k0[2 * e + 1], k1[2 * e + 1]) : 0.f;
                                o[e] = f2_to_h2(r0, r1);
                                const float2 back = h2_to_f2(o[e]);
                                ol[e] = f2_to_h2(r0 - back.x, r1 - back.y);
                            }
                            *slot = v;
                            *slot1 = vl;
                            continue;
                        }
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float2 a = h2_to_f2(x0[e]);
                            o[e] = f2_to_h2(fmaxf(fmaf(a.x, k0[2 * e], k1[2 * e]), 0.f),
                                            fmaxf(fmaf(a.y, k0[2 * e + 1], k1[2 * e + 1]), 0.f));
                        }
                        *slot = v;
                    }
                }
                fence_proxy_async();  // generic-proxy writes (cp.async + the in-place transform) -> visible to the tensor core
#if PN2_WARP_ARRIVE
                __syncwarp();
                if ((tid & 31) == 0) mbar_arrive(&full[p_slot]);
#else
                mbar_arrive(&full[p_slot]);
#endif
                if (++p_kc == KT) { p_kc = 0; p_tile += gridDim.x; }
                if (++p_slot == NST) p_slot = 0;
            }
        }
        (void)p_tile;
    } else if (warp == kMmaWarp) {
        // ================================ MMA issuer ================================
        // forward: fp16 x fp16.  backward: dZ (bf16) x W_A (bf16), then Y (fp16) x W_B (fp16) into the same fp32
        // accumulator -- one instruction multiplies like with like (a mixed fp16 x bf16 tcgen05.mma is an illegal
        // instruction), two instructions of different formats may share the accumulator.  Both weight sets carry the
        // same power-of-two factor (W_B needs it to sit in fp16's range); the epilogue divides it out.
        constexpr uint32_t idesc = FWD ? umma_idesc2(0u, 0u, false, false, TM, BN) : umma_idesc2(1u, 1u, false, false, TM, BN);
        constexpr uint32_t idesc1 = umma_idesc2(0u, 0u, false, false, TM, BN);
        int slot = 0;
        uint32_t phase = 0, as = 0, aphase = 0;
        for (long long t = 0; t < my_tiles; ++t) {
            mbar_wait(&tempty[as], aphase ^ 1);  // the epilogue has drained this accumulator stage
            tc_fence_after();
            for (int kc = 0; kc < KT; ++kc) {
                mbar_wait(&full[slot], phase);
                if (AMODE != A_AFFINE) fence_proxy_async();  // raw stages: no writer-side fence (the copies arrive by themselves)
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t sa = smem_u32(sStage + slot * Cfg::kStage);
                    const uint64_t adesc = umma_desc_k128(sa);
                    const uint64_t bdesc = umma_desc_k128(sa + Cfg::kABytes * Cfg::kNA);
                    const int ksteps = min(TK / 16, (p.kdim - kc * TK + 15) / 16);
                    for (int k4 = 0; k4 < ksteps; ++k4) {  // +32 bytes (>>4 = 2) along K inside the swizzle atom
                        umma_f16(tmem_base + as * ACC, adesc + 2 * k4, bdesc + 2 * k4, idesc, (kc | k4) != 0);
                        if (AMODE == A_BNBWD)
                            umma_f16(tmem_base + as * ACC, umma_desc_k128(sa + Cfg::kABytes) + 2 * k4,
                                     umma_desc_k128(sa + Cfg::kABytes * 2 + Cfg::kBBytes) + 2 * k4, idesc1, 1u);
                        if (SPLIT) {  // + A_lo x B_hi + A_hi x B_lo
                            umma_f16(tmem_base + as * ACC, umma_desc_k128(sa + Cfg::kABytes) + 2 * k4, bdesc + 2 * k4, idesc, 1u);
                            umma_f16(tmem_base + as * ACC, adesc + 2 * k4,
                                     umma_desc_k128(sa + Cfg::kABytes * 2 + Cfg::kBBytes) + 2 * k4, idesc, 1u);
                        }
                    }
                    tc_commit(&empty[slot]);               // arrives when the MMAs above have finished reading the stage
                    if (kc == KT - 1) tc_commit(&tfull[as]);
                }
                __syncwarp();
                if (++slot == NST) { slot = 0; phase ^= 1; }
            }
            as ^= 1;
            if (as == 0) aphase ^= 1;
        }
    } else {
        // ================================ epilogue ================================
        const int chunk = tid % CPR, r0 = tid / CPR;
        const float yscale = AMODE == A_BNBWD ? __ldg(p.yscale) : 0.f;
        float s1[NH][8], s2[NH][8];
#pragma unroll
        for (int h = 0; h < NH; ++h)
#pragma unroll
            for (int e = 0; e < 8; ++e) s1[h][e] = s2[h][e] = 0.f;
        uint32_t as = 0, aphase = 0;
        long long tile = blockIdx.x;
        // the previous layer's y pieces the MASK epilogue reads come from HBM: they are pulled into L2 one column pass
        // ahead (no registers held), the loads proper are issued at the top of the pass that uses them
        auto prefetch_yp = [&](long long tl, int hh) {
            const int c0 = n0 + hh * EBN + chunk * 8;
            if (c0 < p.n) {
#pragma unroll
                for (int ps = 0; ps < (MASK ? PASSES : 1); ++ps) {
                    const long long grow = tl * TM + r0 + ps * RPP;
                    if (grow < p.rows) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.yp + grow * p.yp_ld + c0));
                }
            }
        };
        for (long long t = 0; t < my_tiles; ++t, tile += gridDim.x) {
            if (MASK && t == 0) prefetch_yp(tile, 0);
            mbar_wait(&tfull[as], aphase);
            tc_fence_after();
#pragma unroll
            for (int h = 0; h < NH; ++h) {
                const int col0 = n0 + h * EBN + chunk * 8;  // first of this thread's 8 output columns (pass 2)
                uint4 yq[MASK ? PASSES : 1];
                if (MASK) {
                    if (col0 < p.n) {
#pragma unroll
                        for (int ps = 0; ps < PASSES; ++ps) {
                            const long long grow = tile * TM + r0 + ps * RPP;
                            if (grow < p.rows) yq[ps] = __ldg(reinterpret_cast<const uint4*>(p.yp + grow * p.yp_ld + col0));
                        }
                    }
                    if (h + 1 < NH) prefetch_yp(tile, h + 1);
                    else if (t + 1 < my_tiles) prefetch_yp(tile + gridDim.x, 0);
                }
                // pass 1: TMEM -> registers -> 16-bit tile in shared memory (thread = row; warp = 32 rows x WC columns)
                const int q4 = warp & 3, wc0 = (warp >> 2) * WC;
#pragma unroll
                for (int cw = 0; cw < WC; cw += LDW) {
                    const int cb = wc0 + cw;  // first column (within this EBN-wide pass) of this load
                    if (n0 + h * EBN + cb < p.n) {  // warp-uniform
                        uint32_t v[32];
                        const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) + as * ACC + h * EBN + cb;
                        if (LDW == 32) tmem_ld32(taddr, v); else tmem_ld16(taddr, v);
                        if (AMODE == A_BNBWD) {  // undo the common power-of-two factor of W_A / W_B
#pragma unroll
                            for (int e = 0; e < LDW; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) * yscale);
                        }
                        if constexpr (POOL) {
                            // max-pool in the epilogue: this warp holds 32 consecutive rows (lane = row) of LDW columns.
                            // Values go to an order-preserving unsigned form (inverted where gamma < 0), one
                            // redux.sync.max per column, the first row holding it by ballot.
                            const long long grow = tile * TM + q4 * 32 + lane;
                            const bool live = grow < p.rows;
                            const int k16 = p.pool_k == 16;
                            uint32_t keep_m = 0u;  // the result this lane writes out: column cb + (lane % LDW') of group lane / 16 (k16)
                            int keep_a = 0;
#pragma unroll
                            for (int e = 0; e < LDW; ++e) {
                                uint32_t u = __float_as_uint(__uint_as_float(v[e]) - sCen[h * EBN + cb + e]);
                                u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
                                u ^= sNeg[h * EBN + cb + e];
                                if (!live) u = 0u;
                                if (!k16) {
                                    const uint32_t m = __reduce_max_sync(kFull, u);
                                    const int a = __ffs(__ballot_sync(kFull, u == m)) - 1;
                                    if (lane == e) { keep_m = m; keep_a = a; }
                                } else {  // two 16-row groups per warp
                                    const bool hi16 = lane >= 16;
                                    const uint32_t m0 = __reduce_max_sync(kFull, hi16 ? 0u : u);
                                    const uint32_t m1 = __reduce_max_sync(kFull, hi16 ? u : 0u);
                                    const uint32_t b0 = __ballot_sync(kFull, !hi16 && u == m0);
                                    const uint32_t b1 = __ballot_sync(kFull, hi16 && u == m1);
                                    if (LDW == 16) {  // lanes 0..15 <- group 0, lanes 16..31 <- group 1, column lane % 16
                                        if ((lane & 15) == e) { keep_m = hi16 ? m1 : m0; keep_a = hi16 ? __ffs(b1) - 17 : __ffs(b0) - 1; }
                                    } else {          // 32 columns x 2 groups: stored right away by lanes 0 / 16
                                        if ((lane & 15) == 0) {
                                            const long long g = grow >> 4;
                                            const int col = n0 + h * EBN + cb + e;
                                            uint32_t mm = (hi16 ? m1 : m0) ^ sNeg[h * EBN + cb + e];
                                            mm = (mm & 0x80000000u) ? (mm & 0x7FFFFFFFu) : ~mm;
                                            if (live && col < p.n) {
                                                p.pool_val[g * p.n + col] = __uint_as_float(mm);
                                                p.pool_arg[g * p.n + col] = hi16 ? __ffs(b1) - 17 : __ffs(b0) - 1;
                                            }
                                        }
                                    }
                                }
                            }
                            if (!k16 || LDW == 16) {
                                const int e = k16 ? (lane & 15) : lane;
                                const int colb = h * EBN + cb + e;           // column inside this CTA's BN-wide tile
                                if (e < LDW) {
                                    if (p.pool_k <= 32) {
                                        const long long g = k16 ? (tile * TM + q4 * 32 + (lane & 16)) >> 4 : (tile * TM + q4 * 32) >> 5;
                                        const bool glive = k16 ? (tile * TM + q4 * 32 + (lane & 16)) < p.rows : (tile * TM + q4 * 32) < p.rows;
                                        uint32_t mm = keep_m ^ sNeg[colb];
                                        mm = (mm & 0x80000000u) ? (mm & 0x7FFFFFFFu) : ~mm;
                                        if (glive && n0 + colb < p.n) {
                                            p.pool_val[g * p.n + n0 + colb] = __uint_as_float(mm);
                                            p.pool_arg[g * p.n + n0 + colb] = keep_a;
                                        }
                                    } else {  // groups span quarters: combined after the pass barrier
                                        sPoolV[q4 * BN + colb] = keep_m;
                                        sPoolA[q4 * BN + colb] = keep_a;
                                    }
                                }
                            }
                        }
                        const int r = q4 * 32 + lane;
#pragma unroll
                        for (int g8 = 0; g8 < LDW / 8; ++g8) {
                            const float4 ce0 = *reinterpret_cast<const float4*>(&sCen[h * EBN + cb + g8 * 8]);
                            const float4 ce1 = *reinterpret_cast<const float4*>(&sCen[h * EBN + cb + g8 * 8 + 4]);
                            uint4 q;
                            q.x = FWD ? f2_to_h2(__uint_as_float(v[g8 * 8 + 0]) - ce0.x, __uint_as_float(v[g8 * 8 + 1]) - ce0.y)
                                      : f2_to_bf2(__uint_as_float(v[g8 * 8 + 0]) - ce0.x, __uint_as_float(v[g8 * 8 + 1]) - ce0.y);
                            q.y = FWD ? f2_to_h2(__uint_as_float(v[g8 * 8 + 2]) - ce0.z, __uint_as_float(v[g8 * 8 + 3]) - ce0.w)
                                      : f2_to_bf2(__uint_as_float(v[g8 * 8 + 2]) - ce0.z, __uint_as_float(v[g8 * 8 + 3]) - ce0.w);
                            q.z = FWD ? f2_to_h2(__uint_as_float(v[g8 * 8 + 4]) - ce1.x, __uint_as_float(v[g8 * 8 + 5]) - ce1.y)
                                      : f2_to_bf2(__uint_as_float(v[g8 * 8 + 4]) - ce1.x, __uint_as_float(v[g8 * 8 + 5]) - ce1.y);
                            q.w = FWD ? f2_to_h2(__uint_as_float(v[g8 * 8 + 6]) - ce1.z, __uint_as_float(v[g8 * 8 + 7]) - ce1.w)
                                      : f2_to_bf2(__uint_as_float(v[g8 * 8 + 6]) - ce1.z, __uint_as_float(v[g8 * 8 + 7]) - ce1.w);
                            *reinterpret_cast<uint4*>(&sC[r * CLD + cb + g8 * 8]) = q;
                            if (SPLIT) {  // lo plane: what the fp16 rounding of the hi plane left over
                                const float ce[8] = {ce0.x, ce0.y, ce0.z, ce0.w, ce1.x, ce1.y, ce1.z, ce1.w};
                                const uint32_t* qh = reinterpret_cast<const uint32_t*>(&q);
                                uint4 ql;
                                uint32_t* qlw = reinterpret_cast<uint32_t*>(&ql);
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const float2 back = h2_to_f2(qh[e]);
                                    qlw[e] = f2_to_h2(__uint_as_float(v[g8 * 8 + 2 * e]) - ce[2 * e] - back.x,
                                                      __uint_as_float(v[g8 * 8 + 2 * e + 1]) - ce[2 * e + 1] - back.y);
                                }
                                *reinterpret_cast<uint4*>(&sC[TM * CLD + r * CLD + cb + g8 * 8]) = ql;
                            }
                        }
                    }
                }
                if (h == NH - 1) {  // every tcgen05.ld of this accumulator stage has completed
                    tc_fence_before();
                    mbar_arrive(&tempty[as]);
                }
                epi_bar();
                if (POOL && p.pool_k > 32) {
                    // groups of 64 / 128 rows: combine the quarters' extremes (earlier quarter wins ties: first row in order)
                    const int qpg = p.pool_k >> 5;  // quarters per group: 2 or 4
                    for (int i = tid; i < EBN * (4 / qpg); i += kEpiThreads) {
                        const int gi = i / EBN, colb = h * EBN + (i - gi * EBN);
                        uint32_t best = 0u;
                        int arg = 0;
                        for (int qq = 0; qq < qpg; ++qq) {
                            const uint32_t m = sPoolV[(gi * qpg + qq) * BN + colb];
                            if (m > best) { best = m; arg = qq * 32 + sPoolA[(gi * qpg + qq) * BN + colb]; }
                        }
                        const long long grow0 = tile * TM + gi * p.pool_k;
                        if (grow0 < p.rows && n0 + colb < p.n) {
                            uint32_t mm = best ^ sNeg[colb];
                            mm = (mm & 0x80000000u) ? (mm & 0x7FFFFFFFu) : ~mm;
                            const long long g = grow0 / p.pool_k;
                            p.pool_val[g * p.n + n0 + colb] = __uint_as_float(mm);
                            p.pool_arg[g * p.n + n0 + colb] = arg;
                        }
                    }
                }
                // pass 2: 16-byte pieces, coalesced stores, column sums in registers
                if (col0 < p.n) {
                    float ps_[8], ph_[8];
                    if (MASK) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const int c = h * EBN + chunk * 8 + e;
                            ps_[e] = sPrev[c]; ph_[e] = sPrev[BN + c];
                        }
                    }
#pragma unroll
                    for (int ps = 0; ps < PASSES; ++ps) {
                        const int r = r0 + ps * RPP;
                        const long long grow = tile * TM + r;
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
                                    // sum d*xhat = rstd * (sum d*y - mean * sum d): the per-column constants wait for the end
                                    s1[h][2 * e] += d.x; s1[h][2 * e + 1] += d.y;
                                    s2[h][2 * e] = fmaf(d.x, y.x, s2[h][2 * e]);
                                    s2[h][2 * e + 1] = fmaf(d.y, y.y, s2[h][2 * e + 1]);
                                    vv[e] = f2_to_bf2(d.x, d.y);
                                }
                            } else if (SPLIT) {
                                const uint4 vl = *reinterpret_cast<const uint4*>(&sC[TM * CLD + r * CLD + chunk * 8]);
                                const uint32_t* vlw = reinterpret_cast<const uint32_t*>(&vl);
                                if (p.sums) {
#pragma unroll
                                    for (int e = 0; e < 4; ++e) {
                                        float2 d = h2_to_f2(vv[e]);
                                        const float2 l = h2_to_f2(vlw[e]);
                                        d.x += l.x; d.y += l.y;
                                        s1[h][2 * e] += d.x; s1[h][2 * e + 1] += d.y;
                                        s2[h][2 * e] = fmaf(d.x, d.x, s2[h][2 * e]);
                                        s2[h][2 * e + 1] = fmaf(d.y, d.y, s2[h][2 * e + 1]);
                                    }
                                }
                                if (p.out_lo) *reinterpret_cast<uint4*>(p.out_lo + grow * p.out_ld + col0) = vl;
                            } else if (p.sums) {
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const float2 d = FWD ? h2_to_f2(vv[e]) : bf2_to_f2(vv[e]);
                                    s1[h][2 * e] += d.x; s1[h][2 * e + 1] += d.y;
                                    s2[h][2 * e] = fmaf(d.x, d.x, s2[h][2 * e]);
                                    s2[h][2 * e + 1] = fmaf(d.y, d.y, s2[h][2 * e + 1]);
                                }
                            }
                            *reinterpret_cast<uint4*>(p.out + grow * p.out_ld + col0) = v;
                        }
                    }
                }
                epi_bar();  // sC is rewritten by the next pass 1
            }
            as ^= 1;
            if (as == 0) aphase ^= 1;
        }
        if (p.sums) {
            // column sums: lanes sharing a column piece combine by shuffles, the 8 warps through shared memory,
            // then one atomic per column
            float* red = reinterpret_cast<float*>(sC);  // [8 warps][2][EBN] floats <= 8 KB <= TM*CLD*2 bytes
#pragma unroll
            for (int h = 0; h < NH; ++h) {
#pragma unroll
                for (int e = 0; e < 8; ++e) {
#pragma unroll
                    for (int o = CPR; o < 32; o <<= 1) {
                        s1[h][e] += __shfl_xor_sync(kFull, s1[h][e], o);
                        s2[h][e] += __shfl_xor_sync(kFull, s2[h][e], o);
                    }
                }
                if (lane < CPR) {  // CPR <= 16: lane == chunk for these lanes (CPR divides 32)
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        red[(warp * 2 + 0) * EBN + lane * 8 + e] = s1[h][e];
                        red[(warp * 2 + 1) * EBN + lane * 8 + e] = s2[h][e];
                    }
                }
                epi_bar();
                for (int i = tid; i < 2 * EBN; i += kEpiThreads) {
                    const int which = i / EBN, c = i - which * EBN;
                    const int col = n0 + h * EBN + c;
                    if (col < p.n) {
                        float s = 0.f;
#pragma unroll
                        for (int j = 0; j < kEpiWarps; ++j) s += red[(j * 2 + which) * EBN + c];
                        if (MASK && which == 1) {  // (sum d*y - mean * sum d) * rstd
                            float sd = 0.f;
#pragma unroll
                            for (int j = 0; j < kEpiWarps; ++j) sd += red[(j * 2 + 0) * EBN + c];
                            s = (s - sPrev[2 * BN + h * EBN + c] * sd) * sPrev[3 * BN + h * EBN + c];
                        }
                        atomicAdd(p.sums + (size_t)which * p.n + col, s);
                    }
                }
                epi_bar();
            }
            if (FWD && p.fin_counter) {
                // BatchNorm finalisation by the last CTA to get here: its predecessors' sums are complete (their
                // atomics precede their counter increment) -- one launch less per layer
                __shared__ int s_last;
                __threadfence();
                epi_bar();
                if (tid == 0) s_last = atomicAdd(p.fin_counter, 1u) == gridDim.x * gridDim.y - 1;
                epi_bar();
                if (s_last) {
                    __threadfence();
                    if (tid == 0 && p.fin_nbt) *p.fin_nbt += 1;
                    for (int c = tid; c < p.n; c += kEpiThreads) {
                        const float s1c = __ldcg(p.sums + c);
                        bn_finalize_channel(c, p.n, s1c, __ldcg(p.sums + p.n + c), p.fin_inv_rows, p.fin_unbias,
                                            p.fin_gamma, p.fin_beta, p.fin_bias, p.fin_center, p.fin_momentum, p.fin_eps,
                                            p.fin_running_mean, p.fin_running_var, p.fin_scale, p.fin_shift, p.fin_mean,
                                            p.fin_rstd);
                        // next step's centring constant = this step's batch mean of the un-centred output (may alias
                        // p.center / p.fin_center: every CTA cached the centre at start, this thread read its element above)
                        if (p.fin_next_center) p.fin_next_center[c] = fmaf(s1c, p.fin_inv_rows, p.center ? p.center[c] : 0.f);
                    }
                }
            }
        }
    }

    // ---- teardown
    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

template <int BN, int AMODE, bool MASK, bool SPLIT = false, bool POOL = false>
int launch_tc_(const GemmArgs& a, cudaStream_t stream) {
    using Cfg = TcCfg<BN, AMODE, SPLIT, POOL>;
    const int kpad = (a.kdim + TK - 1) / TK * TK;
    const size_t smem = (size_t)Cfg::kNst * Cfg::kStage + Cfg::kSC + (size_t)Cfg::kNCoef * kpad * 4 + 5 * BN * 4 + Cfg::kPool + 256 + 1024;
    static DeviceOnce once;
    if (once.first()) {
        PN2_CHECK(cudaFuncSetAttribute(gemm_tc_kernel<BN, AMODE, MASK, SPLIT, POOL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       kSmemBudget),
                  "gemm_tc: cudaFuncSetAttribute");
    }
    const int sms = sm_count();
    const long long tiles = (a.rows + TM - 1) / TM;
    const int ny = (a.n + BN - 1) / BN;
    long long gx = sms / ny;  // persistent: one CTA per SM (shared memory, TMEM), tiles dealt round-robin
    if (gx < 1) gx = 1;
    if (gx > tiles) gx = tiles;
    launch_k(gemm_tc_kernel<BN, AMODE, MASK, SPLIT, POOL>, dim3((unsigned)gx, ny), dim3(kTcThreads), smem, stream, a);
    PN2_CHECK_LAUNCH("gemm_tc_kernel");
    return 0;
}

template <int BN, int AMODE, bool MASK, bool SPLIT = false>
int launch_tc(const GemmArgs& a, cudaStream_t stream) {
    if constexpr (AMODE != A_BNBWD && !MASK) {
        if (a.pool_k) return launch_tc_<BN, AMODE, MASK, SPLIT, true>(a, stream);
    }
    return launch_tc_<BN, AMODE, MASK, SPLIT, false>(a, stream);
}

// column-tile width: the narrowest of {32, 64, 128, 256 (forward only)} that covers n, else the widest.  Every extra
// column tile re-reads and re-transforms the whole A operand; padded columns only cost tensor-pipe time (the epilogue
// skips them).
int pick_bn(int n, bool fwd, long long rows) {
    static int fwd_max = -1;
    if (fwd_max < 0) {  // PN2_TC_MAXBN=128: no 256-wide forward tiles (development switch)
        const char* e = getenv("PN2_TC_MAXBN");
        fwd_max = (e && atoi(e) == 128) ? 128 : 256;
    }
    const int widest = fwd ? fwd_max : 128;
    int bn = widest;
    for (int c = 32; c < widest; c <<= 1)
        if (n <= c) { bn = c; break; }
    // Few row tiles (the 4096..10752-row stacks: SA3, FP3, FP2, the 16-neighbour scale of q1 / q2): with one CTA per
    // (row tile, column tile) most SMs would idle while the busy ones each stream the whole weight matrix at one SM's
    // L2 bandwidth.  Narrower column tiles spread the weights over more SMs; re-reading the (small) A tiles is cheap.
    const long long tiles = (rows + TM - 1) / TM;
    while (bn > 32 && tiles * ((n + bn - 1) / bn) < 112) bn >>= 1;
    return bn;
}

template <int AMODE, bool MASK>
int dispatch_tc(const GemmArgs& a, cudaStream_t stream) {
    constexpr bool FWD = AMODE != A_BNBWD;
    switch (pick_bn(a.n, FWD, a.rows)) {
        case 256:
            if constexpr (FWD) return launch_tc<256, AMODE, MASK>(a, stream);
            return launch_tc<128, AMODE, MASK>(a, stream);
        case 128: return launch_tc<128, AMODE, MASK>(a, stream);
        case 64: return launch_tc<64, AMODE, MASK>(a, stream);
        default: return launch_tc<32, AMODE, MASK>(a, stream);
    }
}

// SPLIT: column tiles of at most 128 (the two-plane stage of a 256-wide tile would leave no room for a 3-stage ring)
template <int AMODE>
int dispatch_tc_split(const GemmArgs& a, cudaStream_t stream) {
    switch (pick_bn(a.n, false, a.rows)) {
        case 128: return launch_tc<128, AMODE, false, true>(a, stream);
        case 64: return launch_tc<64, AMODE, false, true>(a, stream);
        default: return launch_tc<32, AMODE, false, true>(a, stream);
    }
}

}  // namespace

bool gemm_use_tc() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("PN2_GEMM_IMPL");
        v = (e && strcmp(e, "mma") == 0) ? 0 : 1;
    }
    return v == 1;
}

int launch_gemm_tc(const GemmArgs& a, int amode, bool mask, cudaStream_t stream) {
    if (a.kdim > kMaxK) return fail_arg("pn2_mlp_gemm", "reduction dimension > 1024");
    if (a.a1 && amode != A_BNBWD) {  // two-plane (hi + lo) forward operands
        if (!a.b1) return fail_arg("pn2_mlp_gemm", "two-plane input rows need two-plane weights");
        if (amode == A_AFFINE && a.kdim > kMaxKSplit) return fail_arg("pn2_mlp_gemm", "two-plane BatchNorm'd input: reduction dimension > 256");
        return amode == A_AFFINE ? dispatch_tc_split<A_AFFINE>(a, stream) : dispatch_tc_split<A_PLAIN>(a, stream);
    }
    if (amode == A_PLAIN) return dispatch_tc<A_PLAIN, false>(a, stream);
    if (amode == A_AFFINE) return dispatch_tc<A_AFFINE, false>(a, stream);
    if (mask) return dispatch_tc<A_BNBWD, true>(a, stream);
    return dispatch_tc<A_BNBWD, false>(a, stream);
}

}  // namespace pn2
