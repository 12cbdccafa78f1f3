// mlp_rows.cu -- building and consuming the ROW matrices of the grouped MLP, sm_100a.
//
// The two ends of every fused SA / FP layer (mlp_gemm.cu holds the middle):
//
//   to_rows          (B,C,N) fp32 channel-major  ->  rows [B*N][ld] fp16        (smem transpose)
//                    (forward rows are fp16, gradient rows bf16: see mma_common.cuh)
//   sa_build_rows    grouping + "- centre" + concat of reference pointnet_utils.py:389-396 (SA-MSG:
//                    [features, xyz - centre]), :570-575 (given centres: ... + centre features
//                    broadcast over K) and :170-186 (group_all: [xyz, features]) in ONE pass, written
//                    as fp16 rows; the source features are themselves rows, optionally still in
//                    pre-BatchNorm form (scale/shift + ReLU applied on the fly).
//   fp_build_rows    three-NN inverse-distance weights (pointnet_utils.py:446-449), three-point
//                    interpolation (interpolate_gpu.cu:149-169) and the [skip, interpolated] concat
//                    (:455-456) in one pass; S == 1 broadcasts (:443-444).
//   pool_fwd         BatchNorm + ReLU of the last layer, max over the K rows of a group (:403,509),
//                    written channel-major fp32 (the module's output layout), plus the arg-max for
//                    backward and per-channel sums: pooled features are handed to the next fused consumer
//                    as fp16 rows CENTRED on their channel mean (to_rows), because a max-pooled feature
//                    is typically a large value with a small spread across groups and plain 16-bit rounding
//                    would eat the spread.  K == 1 is the FP / head case.
//   pool_bwd         routes the output gradient to the arg-max rows through the ReLU mask and reduces
//                    the two BatchNorm-backward sums.
//   sa_rows_bwd / fp_rows_bwd   scatter the gradient of the built rows back to the feature tensors.
//
// All of these are HBM/L2-bound gathers, scatters and transposes: one warp per row with lanes
// along the contiguous channel dimension, shared-memory tiles where a transpose is needed.
#include "mma_common.cuh"

namespace pn2 {
namespace {

constexpr int kThreads = 256;

// ------------------------------------------------------------------ to_rows -----------------
__global__ void __launch_bounds__(256) to_rows_kernel(int c, int n, int ld, const float* __restrict__ src,
                                                       const float* __restrict__ sub_sums, float sub_scale,
                                                       act_t* __restrict__ dst, act_t* __restrict__ dst_lo) {
    pdl_enter();  // programmatic dependent launch (pn2_common.cuh): first statement, before any memory access
    __shared__ float tile[32][33];
    const int b = blockIdx.z, n0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    for (int j = ty; j < 32; j += 8) {
        const int cc = c0 + j, nn = n0 + tx;
        // optional centring: subtract the channel mean (sub_sums[c] * sub_scale) before bf16 rounding
        tile[j][tx] = (cc < c && nn < n)
                          ? src[((size_t)b * c + cc) * n + nn] - (sub_sums ? sub_sums[cc] * sub_scale : 0.f)
                          : 0.f;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
        const int nn = n0 + j, cc = c0 + tx;
        if (nn < n && cc < ld) {
            const float v = tile[tx][j];
            const act_t h = f_to_h(v);
            dst[((size_t)b * n + nn) * ld + cc] = h;
            if (dst_lo) dst_lo[((size_t)b * n + nn) * ld + cc] = f_to_h(v - h_to_f(h));  // two-plane rows: value = hi + lo
        }
    }
}

// ------------------------------------------------------------------ sa_build_rows -----------
struct RowSrc {  // fp16 rows with an optional per-channel affine + ReLU ("still pre-BatchNorm")
    const act_t* p;
    const act_t* lo;  // nullable: second plane of two-plane rows (value = p + lo), same leading dimension
    int c, ld;
    const float *scale, *shift;
};
__device__ __forceinline__ float row_val(const RowSrc& s, size_t row, int ch) {
    float v = h_to_f(s.p[row * s.ld + ch]);
    if (s.lo) {
        // two-plane rows: the ReLU decision is taken on the hi plane alone (what the backward kernels mask by)
        const float full = v + h_to_f(s.lo[row * s.ld + ch]);
        if (!s.scale) return full;
        const float sc = __ldg(s.scale + ch), sh = __ldg(s.shift + ch);
        return fmaf(v, sc, sh) > 0.f ? fmaf(full, sc, sh) : 0.f;
    }
    if (s.scale) v = fmaxf(fmaf(v, __ldg(s.scale + ch), __ldg(s.shift + ch)), 0.f);
    return v;
}

struct SaBuildArgs {
    int b, n, s, k;
    const float *xyz, *new_xyz;  // (B,3,N), (B,3,S) | null (centre 0)
    const int* idx;              // (B,S,K) | null (identity: row k of group <-> point k)
    RowSrc feat, cen;
    int xyz_first;
    act_t* out;
    act_t* out_lo;  // nullable: lo plane of the output rows
    int out_ld;
};

// 8 consecutive channels of a feature row as floats, BatchNorm+ReLU of the producer applied on the fly (16-byte load;
// needs ch % 8 == 0 and ld % 8 == 0)
__device__ __forceinline__ void row_vals8(const RowSrc& s, size_t row, int ch, float (&v)[8]) {
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(s.p + row * s.ld + ch));
    const uint32_t* w = reinterpret_cast<const uint32_t*>(&q);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float2 f = h2_to_f2(w[e]);
        v[2 * e] = f.x; v[2 * e + 1] = f.y;
    }
    if (s.lo) {
        float l[8];
        const uint4 ql = __ldg(reinterpret_cast<const uint4*>(s.lo + row * s.ld + ch));
        const uint32_t* wl = reinterpret_cast<const uint32_t*>(&ql);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 f = h2_to_f2(wl[e]);
            l[2 * e] = f.x; l[2 * e + 1] = f.y;
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            if (s.scale) {  // ReLU decided on the hi plane (see row_val)
                const float sc = __ldg(s.scale + ch + e), sh = __ldg(s.shift + ch + e);
                v[e] = fmaf(v[e], sc, sh) > 0.f ? fmaf(v[e] + l[e], sc, sh) : 0.f;
            } else {
                v[e] += l[e];
            }
        }
        return;
    }
    if (s.scale) {
        const float4 s0 = __ldg(reinterpret_cast<const float4*>(s.scale + ch)), s1 = __ldg(reinterpret_cast<const float4*>(s.scale + ch + 4));
        const float4 h0 = __ldg(reinterpret_cast<const float4*>(s.shift + ch)), h1 = __ldg(reinterpret_cast<const float4*>(s.shift + ch + 4));
        v[0] = fmaxf(fmaf(v[0], s0.x, h0.x), 0.f); v[1] = fmaxf(fmaf(v[1], s0.y, h0.y), 0.f);
        v[2] = fmaxf(fmaf(v[2], s0.z, h0.z), 0.f); v[3] = fmaxf(fmaf(v[3], s0.w, h0.w), 0.f);
        v[4] = fmaxf(fmaf(v[4], s1.x, h1.x), 0.f); v[5] = fmaxf(fmaf(v[5], s1.y, h1.y), 0.f);
        v[6] = fmaxf(fmaf(v[6], s1.z, h1.z), 0.f); v[7] = fmaxf(fmaf(v[7], s1.w, h1.w), 0.f);
    }
}

// A lane owns one 16-byte piece (8 channels) of an output row: rows narrower than 32 pieces share a warp
// (gpr = pieces per row, rpw = rows per warp), wider rows take several passes.  Pieces that lie entirely inside the
// feature segment are one 16-byte gather; pieces straddling a segment boundary (the 3 coordinates shift the centre
// segment off the 8-channel grid) are assembled element by element.
__global__ void __launch_bounds__(kThreads) sa_build_rows_kernel(const SaBuildArgs a) {
    pdl_enter();  // programmatic dependent launch (pn2_common.cuh): first statement, before any memory access
    // centre features (given-centre SA with per-centre features, reference pointnet_utils.py:574-575) are the same for
    // the K rows of a group: when the 8 rows of a CTA share their group the BatchNorm+ReLU'd centre row is staged ONCE,
    // at its OUTPUT columns, and every row copies aligned 16-byte pieces of it
    __shared__ __align__(16) act_t s_cen[1024 + 8];
    __shared__ __align__(16) act_t s_cen_lo[1024 + 8];
    const int lane = threadIdx.x & 31;
    const int gpr = a.out_ld >> 3;
    const int rpw = gpr >= 32 ? 1 : 32 / gpr;
    // 32-bit row arithmetic (the host refuses more than 2^31 rows): the 64-bit divisions this used to do per warp cost
    // more instructions than the row's useful work
    const unsigned total = (unsigned)a.b * (unsigned)a.s * (unsigned)a.k;
    const unsigned wrow = (blockIdx.x * (unsigned)(kThreads / 32) + (threadIdx.x >> 5)) * (unsigned)rpw;
    const int sub = gpr >= 32 ? 0 : lane / gpr;
    const unsigned row = wrow + sub;
    const int fc = a.feat.p ? a.feat.c : 0, cc = a.cen.p ? a.cen.c : 0;
    const int f0 = a.xyz_first ? 3 : 0, x0 = a.xyz_first ? 0 : fc, c0 = fc + 3;
    const bool staged = cc > 0 && rpw == 1 && (a.k & 7) == 0 && a.out_ld <= 1024;
    if (staged) {
        const unsigned grp = (blockIdx.x * (unsigned)(kThreads / 32)) / (unsigned)a.k;  // b * s + s_idx of all 8 rows
        for (int col = (c0 & ~7) + threadIdx.x; col < a.out_ld; col += kThreads) {
            const float v = (col >= c0 && col < c0 + cc) ? row_val(a.cen, (size_t)grp, col - c0) : 0.f;
            const act_t h = f_to_h(v);
            s_cen[col] = h;
            s_cen_lo[col] = f_to_h(v - h_to_f(h));
        }
        __syncthreads();
    }
    if (row >= total || sub >= rpw) return;
    const unsigned bs = row / (unsigned)a.k;
    const int kk = (int)(row - bs * (unsigned)a.k);
    const int b = (int)(bs / (unsigned)a.s), s = (int)(bs - (unsigned)b * (unsigned)a.s);
    const int j = a.idx ? __ldg(a.idx + row) : kk;
    const size_t frow = (size_t)b * a.n + j, crow = (size_t)b * a.s + s;
    const bool fvec = fc > 0 && f0 == 0 && (fc & 7) == 0 && (a.feat.ld & 7) == 0;
    float rel[3] = {0.f, 0.f, 0.f};
    bool have_rel = false;
    act_t* o = a.out + (size_t)row * a.out_ld;
    act_t* ol = a.out_lo ? a.out_lo + (size_t)row * a.out_ld : nullptr;
    for (int g = gpr >= 32 ? lane : lane - sub * gpr; g < gpr; g += 32) {
        const int col = g << 3;
        if (staged && col >= c0 && col + 8 <= c0 + cc) {  // a piece of centre features only
            *reinterpret_cast<uint4*>(o + col) = *reinterpret_cast<const uint4*>(&s_cen[col]);
            if (ol) *reinterpret_cast<uint4*>(ol + col) = *reinterpret_cast<const uint4*>(&s_cen_lo[col]);
            continue;
        }
        if (col >= fc + 3 + cc) {  // zero padding up to the GEMM's chunk granularity
            *reinterpret_cast<uint4*>(o + col) = make_uint4(0u, 0u, 0u, 0u);
            if (ol) *reinterpret_cast<uint4*>(ol + col) = make_uint4(0u, 0u, 0u, 0u);
            continue;
        }
        float v[8];
        if (fvec && col + 8 <= fc) {
            row_vals8(a.feat, frow, col, v);
        } else if ((fc == 0 || (f0 == 0 && col >= fc)) && (cc == 0 || staged)) {
            // A piece without feature channels: relative coordinates, staged centre features, padding.  Kept apart from
            // the general assembly below because the lanes of a warp take these branches side by side: one lane in the
            // general path (eight branchy row_val calls with their global loads) used to cost the whole warp ~400
            // instructions, twice per row of the given-centre stacks (ncu: 45 M warp instructions for 4 M of useful work).
            if (!have_rel && col < x0 + 3 && col + 8 > x0) {
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    const float cx = a.new_xyz ? __ldg(a.new_xyz + ((size_t)b * 3 + d) * a.s + s) : 0.f;
                    rel[d] = __ldg(a.xyz + ((size_t)b * 3 + d) * a.n + j) - cx;
                }
                have_rel = true;
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int c = col + e, d = c - x0;
                float t = 0.f;
                if (d >= 0 && d < 3) t = d == 0 ? rel[0] : (d == 1 ? rel[1] : rel[2]);
                else if (c >= c0 && c < c0 + cc) t = h_to_f(s_cen[c]) + h_to_f(s_cen_lo[c]);
                v[e] = t;
            }
        } else {
            if (!have_rel && col < x0 + 3 && col + 8 > x0) {
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    const float cx = a.new_xyz ? __ldg(a.new_xyz + ((size_t)b * 3 + d) * a.s + s) : 0.f;
                    rel[d] = __ldg(a.xyz + ((size_t)b * 3 + d) * a.n + j) - cx;
                }
                have_rel = true;
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int c = col + e;
                float t = 0.f;
                if (c >= f0 && c < f0 + fc) t = row_val(a.feat, frow, c - f0);
                else if (c >= x0 && c < x0 + 3) t = rel[c - x0];
                else if (c >= c0 && c < c0 + cc) t = staged ? h_to_f(s_cen[c]) + h_to_f(s_cen_lo[c]) : row_val(a.cen, crow, c - c0);
                v[e] = t;
            }
        }
        uint4 q;
        uint32_t* qq = reinterpret_cast<uint32_t*>(&q);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float lo = fminf(fmaxf(v[2 * e], -65504.f), 65504.f), hi = fminf(fmaxf(v[2 * e + 1], -65504.f), 65504.f);
            qq[e] = f2_to_h2(lo, hi);
        }
        *reinterpret_cast<uint4*>(o + col) = q;
        if (ol) {
            uint4 ql;
            uint32_t* qlw = reinterpret_cast<uint32_t*>(&ql);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 back = h2_to_f2(qq[e]);
                qlw[e] = f2_to_h2(v[2 * e] - back.x, v[2 * e + 1] - back.y);
            }
            *reinterpret_cast<uint4*>(ol + col) = ql;
        }
    }
}

// ------------------------------------------------------------------ fp_build_rows -----------
struct FpBuildArgs {
    int b, n, s;
    RowSrc skip, coarse;
    const int* idx;       // (B,N,3)
    const float* dist2;   // (B,N,3) squared distances from three_nn
    act_t* out;
    act_t* out_lo;        // nullable: lo plane of the output rows
    int out_ld;
};
// weights of reference pointnet_utils.py:446-449: w_j = (1/(sqrt(d2_j)+1e-8)) / sum_j(...), in fp32
__device__ __forceinline__ void nn_weights(const float* d2, float (&w)[3]) {
    float r[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) r[j] = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(d2[j]), 1e-8f));
    const float norm = __fadd_rn(__fadd_rn(r[0], r[1]), r[2]);
#pragma unroll
    for (int j = 0; j < 3; ++j) w[j] = __fdiv_rn(r[j], norm);
}

// A (sub-)warp per row: lane l of the row's lane group owns 8-channel pieces l, l + lanes, ... of the COARSE feature rows --
// three 16-byte gathers (the three neighbours) per piece, the producer's BatchNorm+ReLU applied on the fly, one
// interpolation per channel (interpolate_gpu.cu:168 rounding order) -- and writes them at output columns
// skip_c + 8 piece: one 16-byte store when skip_c is a multiple of 8, else eight 2-byte stores (FP1: the skip is the three
// coordinates).  The skip columns and the zero padding are copied by the same lanes afterwards.
// 8 consecutive channels of a feature row with the per-channel constants already in registers (the three neighbours of an
// interpolated point share them)
__device__ __forceinline__ void row_vals8_pre(const RowSrc& s, size_t row, int ch, const float (&sc)[8], const float (&sh)[8],
                                              float (&v)[8]) {
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(s.p + row * s.ld + ch));
    const uint32_t* w = reinterpret_cast<const uint32_t*>(&q);
    float l[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float2 f = h2_to_f2(w[e]);
        v[2 * e] = f.x; v[2 * e + 1] = f.y;
    }
    if (s.lo) {
        const uint4 ql = __ldg(reinterpret_cast<const uint4*>(s.lo + row * s.ld + ch));
        const uint32_t* wl = reinterpret_cast<const uint32_t*>(&ql);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 f = h2_to_f2(wl[e]);
            l[2 * e] = f.x; l[2 * e + 1] = f.y;
        }
    }
    if (!s.lo) {  // one plane: plain BatchNorm + ReLU (two instructions per channel instead of four)
        if (s.scale) {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = fmaxf(fmaf(v[e], sc[e], sh[e]), 0.f);
        }
        return;
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        if (s.scale) v[e] = fmaf(v[e], sc[e], sh[e]) > 0.f ? fmaf(v[e] + l[e], sc[e], sh[e]) : 0.f;  // ReLU decided on the hi plane
        else v[e] += l[e];
    }
}

constexpr int kFpStageLd = 256;  // widest output row the misaligned-store staging buffer holds

__global__ void __launch_bounds__(kThreads) fp_build_rows_kernel(const FpBuildArgs a, int lanes) {
    pdl_enter();  // programmatic dependent launch (pn2_common.cuh): first statement, before any memory access
    // misaligned coarse segment (skip_c % 8 != 0) of a narrow row: the row is assembled in shared memory (2-byte stores)
    // and leaves as 16-byte pieces
    __shared__ __align__(16) act_t s_row[kThreads / 32][2][2][kFpStageLd];  // [warp][row of the warp][plane][column]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rpw = 32 / lanes;                      // rows per warp
    const int l = lane & (lanes - 1), rsub = lane / lanes;
    const unsigned row = (blockIdx.x * (unsigned)(kThreads / 32) + warp) * (unsigned)rpw + rsub;  // < 2^31 rows (host check)
    const unsigned total = (unsigned)a.b * (unsigned)a.n;
    const bool live = row < total;
    const int sc = a.skip.p ? a.skip.c : 0;
    const int cc = a.coarse.c;
    const bool cvec = (cc & 7) == 0 && (a.coarse.ld & 7) == 0;
    const bool staged = cvec && (sc & 7) != 0 && a.out_ld <= kFpStageLd && rpw <= 2;  // warp-uniform
    act_t* o = nullptr;
    act_t* ol = nullptr;
    int b = 0;
    int id[3] = {0, 0, 0};
    float w[3] = {1.f, 0.f, 0.f};
    if (live) {
        b = (int)(row / a.n);
        if (a.s > 1) {
#pragma unroll
            for (int j = 0; j < 3; ++j) id[j] = a.idx[(size_t)row * 3 + j];
            nn_weights(a.dist2 + (size_t)row * 3, w);
        }
        o = staged ? &s_row[warp][rsub][0][0] : a.out + (size_t)row * a.out_ld;
        ol = a.out_lo ? (staged ? &s_row[warp][rsub][1][0] : a.out_lo + (size_t)row * a.out_ld) : nullptr;
    }
    if (live && cvec) {
        const size_t r0 = (size_t)b * a.s + id[0], r1 = (size_t)b * a.s + id[1], r2 = (size_t)b * a.s + id[2];
        for (int cp = l; cp < (cc >> 3); cp += lanes) {
            float ks[8], kh[8], v[8];
            if (a.coarse.scale) {
                const float4 s0 = __ldg(reinterpret_cast<const float4*>(a.coarse.scale + cp * 8)), s1 = __ldg(reinterpret_cast<const float4*>(a.coarse.scale + cp * 8 + 4));
                const float4 h0 = __ldg(reinterpret_cast<const float4*>(a.coarse.shift + cp * 8)), h1 = __ldg(reinterpret_cast<const float4*>(a.coarse.shift + cp * 8 + 4));
                ks[0] = s0.x; ks[1] = s0.y; ks[2] = s0.z; ks[3] = s0.w; ks[4] = s1.x; ks[5] = s1.y; ks[6] = s1.z; ks[7] = s1.w;
                kh[0] = h0.x; kh[1] = h0.y; kh[2] = h0.z; kh[3] = h0.w; kh[4] = h1.x; kh[5] = h1.y; kh[6] = h1.z; kh[7] = h1.w;
            } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) { ks[e] = 1.f; kh[e] = 0.f; }
            }
            if (a.s > 1) {
                float p0[8], p1[8], p2[8];
                row_vals8_pre(a.coarse, r0, cp * 8, ks, kh, p0);
                row_vals8_pre(a.coarse, r1, cp * 8, ks, kh, p1);
                row_vals8_pre(a.coarse, r2, cp * 8, ks, kh, p2);
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = fmaf(w[2], p2[e], fmaf(w[0], p0[e], w[1] * p1[e]));  // interpolate_gpu.cu:168 order
            } else {
                row_vals8_pre(a.coarse, (size_t)b, cp * 8, ks, kh, v);
            }
            uint4 q, ql;
            uint32_t* qq = reinterpret_cast<uint32_t*>(&q);
            uint32_t* qlw = reinterpret_cast<uint32_t*>(&ql);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                qq[e] = f2_to_h2(fminf(fmaxf(v[2 * e], -65504.f), 65504.f), fminf(fmaxf(v[2 * e + 1], -65504.f), 65504.f));
                const float2 back = h2_to_f2(qq[e]);
                qlw[e] = f2_to_h2(v[2 * e] - back.x, v[2 * e + 1] - back.y);
            }
            const int col = sc + cp * 8;
            if ((sc & 7) == 0) {
                *reinterpret_cast<uint4*>(o + col) = q;
                if (ol) *reinterpret_cast<uint4*>(ol + col) = ql;
            } else {
                const act_t* qh = reinterpret_cast<const act_t*>(&q);
                const act_t* qlh = reinterpret_cast<const act_t*>(&ql);
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    o[col + e] = qh[e];
                    if (ol) ol[col + e] = qlh[e];
                }
            }
        }
    }
    // skip columns, (non-vectorisable coarse columns,) zero padding
    const int used = sc + cc, used8 = (used + 7) & ~7;   // columns that carry data; rounded up to whole 16-byte pieces
    if (live) {
        // with vectorised coarse columns only the skip columns and the padding are left: the padding beyond the last
        // piece that carries data goes out as 16-byte zero stores (straight to global memory in the staged case too)
        const int scalar_end = cvec ? used8 : a.out_ld;
        for (int col = l; col < scalar_end; col += lanes) {
            float v = 0.f;
            if (col < sc) {
                v = row_val(a.skip, (size_t)row, col);
            } else if (col < used) {
                if (cvec) continue;
                const int ch = col - sc;
                if (a.s > 1) {
                    const float p0 = row_val(a.coarse, (size_t)b * a.s + id[0], ch);
                    const float p1 = row_val(a.coarse, (size_t)b * a.s + id[1], ch);
                    const float p2 = row_val(a.coarse, (size_t)b * a.s + id[2], ch);
                    v = fmaf(w[2], p2, fmaf(w[0], p0, w[1] * p1));  // interpolate_gpu.cu:168 rounding order
                } else {
                    v = row_val(a.coarse, (size_t)b, ch);
                }
            }
            const act_t h = f_to_h(v);
            o[col] = h;
            if (ol) ol[col] = f_to_h(v - h_to_f(h));
        }
        if (cvec) {
            act_t* go = a.out + (size_t)row * a.out_ld;
            for (int g = (used8 >> 3) + l; g < (a.out_ld >> 3); g += lanes) {
                *reinterpret_cast<uint4*>(go + g * 8) = make_uint4(0u, 0u, 0u, 0u);
                if (a.out_lo) *reinterpret_cast<uint4*>(a.out_lo + (size_t)row * a.out_ld + g * 8) = make_uint4(0u, 0u, 0u, 0u);
            }
        }
    }
    if (staged) {
        __syncwarp();
        if (live) {
            act_t* go = a.out + (size_t)row * a.out_ld;
            for (int g = l; g < (used8 >> 3); g += lanes) {
                *reinterpret_cast<uint4*>(go + g * 8) = *reinterpret_cast<const uint4*>(&s_row[warp][rsub][0][g * 8]);
                if (a.out_lo)
                    *reinterpret_cast<uint4*>(a.out_lo + (size_t)row * a.out_ld + g * 8) =
                        *reinterpret_cast<const uint4*>(&s_row[warp][rsub][1][g * 8]);
            }
        }
    }
}

// ------------------------------------------------------------------ pool fwd / bwd ----------
struct PoolArgs {
    int b, s, k, c;
    const act_t* y; int y_ld;
    const act_t* y_lo;    // forward, nullable: lo plane of two-plane rows (value = y + y_lo)
    const float *scale, *shift, *mean, *rstd;
    float* out_cm;        // (B,C,S)
    float* chan_sums;     // [C] += sum over (b,s) of the output (nullable)
    int* argmax;          // (B,S,C)
    const float* dout_cm; // backward (nullable when extra_rows carries the whole gradient)
    const float* extra_rows;  // backward, k == 1: additional output gradient in row form [B*S][C] (nullable)
    const bf16* extra16; int extra16_ld;  // backward, k == 1: ... and one in bf16 rows (a dense consumer's dx; nullable)
    bf16* dz; int dz_ld;
    float* sums;          // [2][C]
};

// grid (s-tile slots, 64-channel chunks, B).  A CTA walks s-tiles of GT groups (stride gridDim.x); a warp
// owns groups s0+warp, +8, ...; lanes own channel pairs of the chunk.  GT = 8 when there are few groups
// (group-all, the 21 joints) so that the work still spreads over >= 256 CTAs.
template <int GT>
__global__ void __launch_bounds__(kThreads) pool_fwd_kernel(const PoolArgs a) {
    pdl_enter();  // programmatic dependent launch (pn2_common.cuh): first statement, before any memory access
    // A warp reads a group's rows as 16-byte pieces: lane = (row lane rl = lane / 8, piece pc = lane % 8 of the 64-channel
    // chunk); row lane rl takes rows rl, rl + 4, ... (8 independent 16-byte loads per plane in flight at K = 32), the four
    // row lanes are combined by shuffles (first maximum in row order wins, as a sequential scan would have it).
    __shared__ float tile[GT][65];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int pc = lane & 7, rl = lane >> 3;
    const int b = blockIdx.z, c0 = blockIdx.y * 64;
    const int ch = c0 + pc * 8;
    const bool ok = ch < a.c;  // c is a multiple of 8
    float sc[8], sh[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { sc[e] = ok ? a.scale[ch + e] : 0.f; sh[e] = ok ? a.shift[ch + e] : 0.f; }
    float csum[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) csum[j] = 0.f;
    const int s_tiles = (a.s + GT - 1) / GT;
    for (int t = blockIdx.x; t < s_tiles; t += gridDim.x) {
        const int s0 = t * GT;
        for (int gi = warp; gi < GT; gi += 8) {
            const int s = s0 + gi;
            float m[8];
            int mi[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) { m[e] = -1.f; mi[e] = 0x7fffffff; }  // below every ReLU output
            if (ok && s < a.s) {
                const act_t* yr = a.y + ((size_t)(b * a.s + s) * a.k) * a.y_ld + ch;
                const act_t* yl = a.y_lo ? a.y_lo + ((size_t)(b * a.s + s) * a.k) * a.y_ld + ch : nullptr;
#pragma unroll 4
                for (int kk = rl; kk < a.k; kk += 4) {
                    const uint4 q = __ldg(reinterpret_cast<const uint4*>(yr + (size_t)kk * a.y_ld));
                    uint4 ql = make_uint4(0u, 0u, 0u, 0u);
                    if (yl) ql = __ldg(reinterpret_cast<const uint4*>(yl + (size_t)kk * a.y_ld));
                    const uint32_t* w = reinterpret_cast<const uint32_t*>(&q);
                    const uint32_t* wl = reinterpret_cast<const uint32_t*>(&ql);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float2 f = h2_to_f2(w[e]);
                        const float2 fl = h2_to_f2(wl[e]);
                        // two-plane rows: ReLU decided on the hi plane (what pool_bwd masks by), value from hi + lo
                        const float r0 = fmaf(f.x, sc[2 * e], sh[2 * e]) > 0.f ? fmaf(f.x + fl.x, sc[2 * e], sh[2 * e]) : 0.f;
                        const float r1 = fmaf(f.y, sc[2 * e + 1], sh[2 * e + 1]) > 0.f
                                             ? fmaf(f.y + fl.y, sc[2 * e + 1], sh[2 * e + 1]) : 0.f;
                        if (r0 > m[2 * e]) { m[2 * e] = r0; mi[2 * e] = kk; }
                        if (r1 > m[2 * e + 1]) { m[2 * e + 1] = r1; mi[2 * e + 1] = kk; }
                    }
                }
            }
#pragma unroll
            for (int o = 8; o < 32; o <<= 1) {
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float v = __shfl_xor_sync(kFull, m[e], o);
                    const int vi = __shfl_xor_sync(kFull, mi[e], o);
                    if (v > m[e] || (v == m[e] && vi < mi[e])) { m[e] = v; mi[e] = vi; }
                }
            }
            if (rl == 0) {
                if (ok && s < a.s && a.argmax) {
                    int* am = a.argmax + ((size_t)b * a.s + s) * a.c + ch;
                    *reinterpret_cast<int4*>(am) = make_int4(mi[0], mi[1], mi[2], mi[3]);
                    *reinterpret_cast<int4*>(am + 4) = make_int4(mi[4], mi[5], mi[6], mi[7]);
                }
#pragma unroll
                for (int e = 0; e < 8; ++e) tile[gi][pc * 8 + e] = m[e];
            }
        }
        __syncthreads();
        if (lane < GT) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int cc = warp + j * 8, chn = c0 + cc, s = s0 + lane;
                if (chn < a.c && s < a.s) {
                    const float v = tile[lane][cc];
                    a.out_cm[((size_t)b * a.c + chn) * a.s + s] = v;
                    csum[j] += v;
                }
            }
        }
        __syncthreads();
    }
    if (a.chan_sums) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float t = csum[j];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(kFull, t, o);
            const int chn = c0 + warp + j * 8;
            if (lane == 0 && chn < a.c) atomicAdd(a.chan_sums + chn, t);
        }
    }
}

template <int GT>
__global__ void __launch_bounds__(kThreads) pool_bwd_kernel(const PoolArgs a) {
    pdl_enter();  // programmatic dependent launch (pn2_common.cuh): first statement, before any memory access
    __shared__ float tile[GT][65];
    __shared__ float red[8][2][64];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.z, c0 = blockIdx.y * 64;
    const int ch = c0 + lane * 2;
    const bool ok = ch < a.c;
    float sc0 = 0.f, sc1 = 0.f, sh0 = 0.f, sh1 = 0.f, mu0 = 0.f, mu1 = 0.f, rs0 = 0.f, rs1 = 0.f;
    if (ok) {
        sc0 = a.scale[ch]; sc1 = a.scale[ch + 1]; sh0 = a.shift[ch]; sh1 = a.shift[ch + 1];
        mu0 = a.mean[ch]; mu1 = a.mean[ch + 1]; rs0 = a.rstd[ch]; rs1 = a.rstd[ch + 1];
    }
    float p1a = 0.f, p1b = 0.f, p2a = 0.f, p2b = 0.f;
    const int s_tiles = (a.s + GT - 1) / GT;
    for (int t = blockIdx.x; t < s_tiles; t += gridDim.x) {
        const int s0 = t * GT;
        if (lane < GT) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int cc = warp + j * 8, chn = c0 + cc, s = s0 + lane;
                tile[lane][cc] = (a.dout_cm && chn < a.c && s < a.s) ? a.dout_cm[((size_t)b * a.c + chn) * a.s + s] : 0.f;
            }
        }
        __syncthreads();
        for (int gi = warp; gi < GT; gi += 8) {
            const int s = s0 + gi;
            if (!(ok && s < a.s)) continue;
            const size_t g = (size_t)b * a.s + s;
            int i0 = 0, i1 = 0;
            if (a.argmax) { const int2 am = *reinterpret_cast<const int2*>(a.argmax + g * a.c + ch); i0 = am.x; i1 = am.y; }
            float g0 = tile[gi][lane * 2], g1 = tile[gi][lane * 2 + 1];
            if (a.extra_rows) {  // row-form gradients of fused consumers (fp32 [B*S][C], see pn2_sa_rows_bwd)
                const float2 ex = *reinterpret_cast<const float2*>(a.extra_rows + g * a.c + ch);
                g0 += ex.x; g1 += ex.y;
            }
            const act_t* yr = a.y + (g * a.k) * a.y_ld + ch;
            bf16* dr = a.dz + (g * a.k) * a.dz_ld + ch;
            // the two arg-max rows' values first (two independent loads), then a store-only loop over the K rows
            const float y0 = h2_to_f2(*reinterpret_cast<const uint32_t*>(yr + (size_t)i0 * a.y_ld)).x;
            const float y1 = h2_to_f2(*reinterpret_cast<const uint32_t*>(yr + (size_t)i1 * a.y_ld)).y;
            float d0 = 0.f, d1 = 0.f;
            if (fmaf(y0, sc0, sh0) > 0.f) { d0 = g0; p1a += d0; p2a = fmaf(d0, (y0 - mu0) * rs0, p2a); }
            if (fmaf(y1, sc1, sh1) > 0.f) { d1 = g1; p1b += d1; p2b = fmaf(d1, (y1 - mu1) * rs1, p2b); }
#pragma unroll 8
            for (int kk = 0; kk < a.k; ++kk)
                *reinterpret_cast<uint32_t*>(dr + (size_t)kk * a.dz_ld) = f2_to_bf2(kk == i0 ? d0 : 0.f, kk == i1 ? d1 : 0.f);
        }
        __syncthreads();
    }
    red[warp][0][lane * 2] = p1a; red[warp][0][lane * 2 + 1] = p1b;
    red[warp][1][lane * 2] = p2a; red[warp][1][lane * 2 + 1] = p2b;
    __syncthreads();
    if (threadIdx.x < 128) {
        const int which = threadIdx.x >> 6, cc = threadIdx.x & 63;
        if (c0 + cc < a.c) {
            float t = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) t += red[w][which][cc];
            if (t != 0.f) atomicAdd(a.sums + (size_t)which * a.c + c0 + cc, t);
        }
    }
}

// ---- pooled values already taken in the GEMM epilogue (pn2_mlp_gemm_fwd[_bn]_pool): BatchNorm + ReLU of the selected
// extreme, (groups, C) -> channel-major (B, C, S), channel sums.  The ReLU decision is taken on the stored (hi plane) y of
// the selected row -- what pool_bwd masks by --, the value comes from the fp32 accumulator.
__global__ void __launch_bounds__(256) pool_finalize_kernel(int s_count, int k, int c, const float* __restrict__ val,
                                                             const int* __restrict__ arg, const act_t* __restrict__ y, int y_ld,
                                                             const float* __restrict__ scale, const float* __restrict__ shift,
                                                             float* __restrict__ out_cm, float* __restrict__ chan_sums) {
    pdl_enter();  // programmatic dependent launch (pn2_common.cuh): first statement, before any memory access
    __shared__ float tile[32][33];
    const int b = blockIdx.z, s0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    const int ch = c0 + tx;
    float sc = 0.f, sh = 0.f;
    if (ch < c) { sc = scale[ch]; sh = shift[ch]; }
    for (int j = ty; j < 32; j += 8) {
        const int s = s0 + j;
        float r = 0.f;
        if (s < s_count && ch < c) {
            const size_t g = (size_t)b * s_count + s;
            const float yv = h_to_f(y[(g * k + arg[g * c + ch]) * y_ld + ch]);
            r = fmaf(yv, sc, sh) > 0.f ? fmaxf(fmaf(val[g * c + ch], sc, sh), 0.f) : 0.f;
        }
        tile[j][tx] = r;
    }
    __syncthreads();
    float csum = 0.f;
    for (int j = ty; j < 32; j += 8) {
        const int cc = c0 + j, s = s0 + tx;
        if (cc < c && s < s_count) {
            const float v = tile[tx][j];
            out_cm[((size_t)b * c + cc) * s_count + s] = v;
            csum += v;
        }
    }
    if (chan_sums) {  // lanes of a warp hold 32 groups of channel c0 + ty (+ 8, ...): one row j per iteration, summed per j
        // csum mixes the rows j = ty, ty + 8, ...; redo it per channel
        for (int j = ty; j < 32; j += 8) {
            const int cc = c0 + j, s = s0 + tx;
            float v = (cc < c && s < s_count) ? tile[tx][j] : 0.f;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
            if (tx == 0 && cc < c && v != 0.f) atomicAdd(chan_sums + cc, v);
        }
    }
    (void)csum;
}

// ---- few groups (given-centre SA: 21 joints; group-all): one CTA per group at a time, the K rows of the group split
// over the 8 warps, a lane owning 8 consecutive channels (16-byte row pieces), slabs of 256 channels.
constexpr int kSlab = 256;

__global__ void __launch_bounds__(kThreads) pool_fwd_grp_kernel(const PoolArgs a) {
    pdl_enter();  // programmatic dependent launch (pn2_common.cuh): first statement, before any memory access
    __shared__ float sv[8][kSlab];
    __shared__ int si[8][kSlab];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int groups = a.b * a.s;
    for (int c0 = 0; c0 < a.c; c0 += kSlab) {
        const int cw = min(kSlab, a.c - c0);        // channels of this slab (multiple of 8)
        const int ch = c0 + lane * 8;
        const bool on = lane * 8 < cw;
        float sc[8], sh[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) { sc[e] = on ? a.scale[ch + e] : 0.f; sh[e] = on ? a.shift[ch + e] : 0.f; }
        float csum = 0.f;                           // thread tid <-> channel c0 + tid
        for (int g = blockIdx.x; g < groups; g += gridDim.x) {
            float m[8];
            int mi[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) { m[e] = -1.f; mi[e] = 0; }  // below every ReLU output: the first row always wins
            if (on) {
                const act_t* yr = a.y + ((size_t)g * a.k) * a.y_ld + ch;
                const act_t* yl = a.y_lo ? a.y_lo + ((size_t)g * a.k) * a.y_ld + ch : nullptr;
#pragma unroll 4
                for (int kk = warp; kk < a.k; kk += 8) {
                    const uint4 q = __ldg(reinterpret_cast<const uint4*>(yr + (size_t)kk * a.y_ld));
                    uint4 ql = make_uint4(0u, 0u, 0u, 0u);  // fp16 zeros
                    if (yl) ql = __ldg(reinterpret_cast<const uint4*>(yl + (size_t)kk * a.y_ld));
                    const uint32_t* w = reinterpret_cast<const uint32_t*>(&q);
                    const uint32_t* wl = reinterpret_cast<const uint32_t*>(&ql);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float2 f = h2_to_f2(w[e]);
                        const float2 fl = h2_to_f2(wl[e]);
                        // ReLU decided on the hi plane (what pool_bwd masks by), value from hi + lo
                        const float r0 = fmaf(f.x, sc[2 * e], sh[2 * e]) > 0.f ? fmaf(f.x + fl.x, sc[2 * e], sh[2 * e]) : 0.f;
                        const float r1 = fmaf(f.y, sc[2 * e + 1], sh[2 * e + 1]) > 0.f
                                             ? fmaf(f.y + fl.y, sc[2 * e + 1], sh[2 * e + 1]) : 0.f;
                        if (r0 > m[2 * e]) { m[2 * e] = r0; mi[2 * e] = kk; }
                        if (r1 > m[2 * e + 1]) { m[2 * e + 1] = r1; mi[2 * e + 1] = kk; }
                    }
                }
#pragma unroll
                for (int e = 0; e < 8; ++e) { sv[warp][lane * 8 + e] = m[e]; si[warp][lane * 8 + e] = mi[e]; }
            }
            __syncthreads();
            if (tid < cw) {
                float best = sv[0][tid];
                int bi = si[0][tid];
#pragma unroll
                for (int w2 = 1; w2 < 8; ++w2) {  // first maximum in row order: larger value, or equal value and earlier row
                    const float v = sv[w2][tid];
                    const int i2 = si[w2][tid];
                    if (v > best || (v == best && i2 < bi)) { best = v; bi = i2; }
                }
                const int b = g / a.s, s_ = g - b * a.s;
                a.out_cm[((size_t)b * a.c + c0 + tid) * a.s + s_] = best;
                if (a.argmax) a.argmax[(size_t)g * a.c + c0 + tid] = bi;
                csum += best;
            }
            __syncthreads();
        }
        if (a.chan_sums && tid < cw && csum != 0.f) atomicAdd(a.chan_sums + c0 + tid, csum);
    }
}

__global__ void __launch_bounds__(kThreads) pool_bwd_grp_kernel(const PoolArgs a) {
    pdl_enter();  // programmatic dependent launch (pn2_common.cuh): first statement, before any memory access
    __shared__ float sg[kSlab];
    __shared__ int si[kSlab];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int groups = a.b * a.s;
    for (int c0 = 0; c0 < a.c; c0 += kSlab) {
        const int cw = min(kSlab, a.c - c0);
        const int c = c0 + tid;
        const bool mine = tid < cw;
        float sc = 0.f, sh = 0.f, mu = 0.f, rs = 0.f, p1 = 0.f, p2 = 0.f;
        if (mine) { sc = a.scale[c]; sh = a.shift[c]; mu = a.mean[c]; rs = a.rstd[c]; }
        for (int g = blockIdx.x; g < groups; g += gridDim.x) {
            if (mine) {  // the output gradient lands on the arg-max row, through the ReLU mask
                const int b = g / a.s, s_ = g - b * a.s;
                const int kk = a.argmax[(size_t)g * a.c + c];
                const float yv = h_to_f(a.y[((size_t)g * a.k + kk) * a.y_ld + c]);
                float d = a.dout_cm ? a.dout_cm[((size_t)b * a.c + c) * a.s + s_] : 0.f;
                if (a.extra_rows) d += a.extra_rows[(size_t)g * a.c + c];  // row-form gradients of fused consumers
                d = fmaf(yv, sc, sh) > 0.f ? d : 0.f;
                p1 += d;
                p2 = fmaf(d, (yv - mu) * rs, p2);
                sg[tid] = d;
                si[tid] = kk;
            }
            __syncthreads();
            if (lane * 8 < cw) {
                float gv[8];
                int gi[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) { gv[e] = sg[lane * 8 + e]; gi[e] = si[lane * 8 + e]; }
                bf16* dr = a.dz + ((size_t)g * a.k) * a.dz_ld + c0 + lane * 8;
                for (int kk = warp; kk < a.k; kk += 8) {
                    uint4 q;
                    q.x = f2_to_bf2(gi[0] == kk ? gv[0] : 0.f, gi[1] == kk ? gv[1] : 0.f);
                    q.y = f2_to_bf2(gi[2] == kk ? gv[2] : 0.f, gi[3] == kk ? gv[3] : 0.f);
                    q.z = f2_to_bf2(gi[4] == kk ? gv[4] : 0.f, gi[5] == kk ? gv[5] : 0.f);
                    q.w = f2_to_bf2(gi[6] == kk ? gv[6] : 0.f, gi[7] == kk ? gv[7] : 0.f);
                    *reinterpret_cast<uint4*>(dr + (size_t)kk * a.dz_ld) = q;
                }
            }
            __syncthreads();
        }
        if (mine) {
            if (p1 != 0.f) atomicAdd(a.sums + c, p1);
            if (p2 != 0.f) atomicAdd(a.sums + a.c + c, p2);
        }
    }
}

// ---- K == 1 (FP layers, head): BatchNorm+ReLU and rows -> channel-major transpose, wide accesses.
// A CTA owns tiles of 32 consecutive rows x all C channels.  blockDim is a multiple of the C/8 16-byte
// pieces of a row, so thread t always handles piece t % pieces (its 8 channels' constants and partial
// sums live in registers) of rows t / pieces, + blockDim / pieces, ...: row reads/writes are fully
// coalesced.  The shared tile is [32 rows][C + 1] with the column of channel ch permuted to
// (ch % 8) * pieces + ch / 8, which makes both the piece-wise side and the channel-major side
// (128-byte runs of 32 points per channel) bank-conflict free.
constexpr int kRowTile = 32;

__global__ void __launch_bounds__(kThreads) rows_to_cm_kernel(const PoolArgs a) {
    pdl_enter();  // programmatic dependent launch (pn2_common.cuh): first statement, before any memory access
    extern __shared__ __align__(16) float tile_dyn[];  // [32][C + 1]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int pieces = a.c >> 3, ldt = a.c + 1;
    const int pc = tid % pieces, r0 = tid / pieces, rstep = blockDim.x / pieces;
    float sc[8], sh[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { sc[e] = a.scale[pc * 8 + e]; sh[e] = a.shift[pc * 8 + e]; }
    const int s_tiles = (a.s + kRowTile - 1) / kRowTile;
    // persistent CTAs over the B * s_tiles tiles (a grid of one CTA per tile slot would run 1.7 waves at B=32, N=4096)
    for (int gt = blockIdx.x; gt < a.b * s_tiles; gt += gridDim.x) {
        const int b = gt / s_tiles, t = gt - b * s_tiles;
        const int s0 = t * kRowTile;
        // four rows per thread at a time, every load of the batch issued before the first use (memory-level parallelism)
        constexpr int RB = 4;
        for (int rb = r0; rb < kRowTile; rb += RB * rstep) {
            uint4 qb[RB], qlb[RB];
#pragma unroll
            for (int u = 0; u < RB; ++u) {
                const int r = rb + u * rstep;
                qb[u] = qlb[u] = make_uint4(0u, 0u, 0u, 0u);
                if (r < kRowTile && s0 + r < a.s) {
                    qb[u] = __ldg(reinterpret_cast<const uint4*>(a.y + ((size_t)b * a.s + s0 + r) * a.y_ld + pc * 8));
                    if (a.y_lo) qlb[u] = __ldg(reinterpret_cast<const uint4*>(a.y_lo + ((size_t)b * a.s + s0 + r) * a.y_ld + pc * 8));
                }
            }
#pragma unroll
            for (int u = 0; u < RB; ++u) {
                const int r = rb + u * rstep;
                if (!(r < kRowTile && s0 + r < a.s)) continue;
                const uint32_t* v = reinterpret_cast<const uint32_t*>(&qb[u]);
                const uint32_t* vl = reinterpret_cast<const uint32_t*>(&qlb[u]);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float2 f = h2_to_f2(v[e]);
                    const float2 fl = h2_to_f2(vl[e]);
                    tile_dyn[r * ldt + (2 * e) * pieces + pc] =
                        fmaf(f.x, sc[2 * e], sh[2 * e]) > 0.f ? fmaf(f.x + fl.x, sc[2 * e], sh[2 * e]) : 0.f;
                    tile_dyn[r * ldt + (2 * e + 1) * pieces + pc] =
                        fmaf(f.y, sc[2 * e + 1], sh[2 * e + 1]) > 0.f ? fmaf(f.y + fl.y, sc[2 * e + 1], sh[2 * e + 1]) : 0.f;
                }
            }
        }
        __syncthreads();
        for (int ch = warp; ch < a.c; ch += nwarps)
            if (s0 + lane < a.s)
                a.out_cm[((size_t)b * a.c + ch) * a.s + s0 + lane] = tile_dyn[lane * ldt + (ch & 7) * pieces + (ch >> 3)];
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kThreads, 2) cm_to_rows_bwd_kernel(const PoolArgs a) {
    pdl_enter();  // programmatic dependent launch (pn2_common.cuh): first statement, before any memory access
    extern __shared__ __align__(16) float tile_dyn[];  // [32][C + 1] dout tile
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int pieces = a.c >> 3, ldt = a.c + 1;
    const int pc = tid % pieces, r0 = tid / pieces, rstep = blockDim.x / pieces;
    float sc[8], sh[8], mu[8], rs[8], p1[8], p2[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int ch = pc * 8 + e;
        sc[e] = a.scale[ch]; sh[e] = a.shift[ch]; mu[e] = a.mean[ch]; rs[e] = a.rstd[ch];
        p1[e] = p2[e] = 0.f;
    }
    const int s_tiles = (a.s + kRowTile - 1) / kRowTile;
    for (int gt = blockIdx.x; gt < a.b * s_tiles; gt += gridDim.x) {  // persistent CTAs over the B * s_tiles tiles
        const int b = gt / s_tiles, t = gt - b * s_tiles;
        const int s0 = t * kRowTile;
        // channel-major side: 8 independent 128-byte runs in flight per warp (c is a multiple of 8)
        if (a.dout_cm) {
            // 4-byte cp.async straight into the permuted tile: a warp's 48 (C = 384) 128-byte runs are ALL in flight at
            // once and cost no registers (register-staged batches of 8 left ~15 KB per SM in flight: latency-bound)
            const bool in = s0 + lane < a.s;
            const float* src = a.dout_cm + (size_t)b * a.c * a.s + s0 + lane;
            const uint32_t dst = smem_u32(tile_dyn) + (uint32_t)(lane * ldt) * 4u;
            for (int ch = warp; ch < a.c; ch += nwarps) {
                const uint32_t d = dst + (uint32_t)((ch & 7) * pieces + (ch >> 3)) * 4u;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(in ? src + (size_t)ch * a.s : a.dout_cm),
                             "r"(in ? 4 : 0)
                             : "memory");
            }
            cp_async_commit();
            cp_async_wait_all();
        }
        __syncthreads();
        // row side, four rows per thread at a time: every global load of the batch is issued before the first store (a
        // store followed by the next row's loads would serialise one memory latency per row)
        constexpr int RB = 4;
        for (int rb = r0; rb < kRowTile; rb += RB * rstep) {
            uint4 qy[RB], q16[RB];
            float4 e0[RB], e1[RB];
            bool ok[RB];
#pragma unroll
            for (int u = 0; u < RB; ++u) {
                const int r = rb + u * rstep;
                ok[u] = r < kRowTile && s0 + r < a.s;
                e0[u] = e1[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                q16[u] = make_uint4(0u, 0u, 0u, 0u);
                if (ok[u]) {
                    const size_t row = (size_t)b * a.s + s0 + r;
                    qy[u] = __ldg(reinterpret_cast<const uint4*>(a.y + row * a.y_ld + pc * 8));
                    if (a.extra_rows) {  // gradient contributions consumers delivered in row form (sparse gathers)
                        e0[u] = __ldg(reinterpret_cast<const float4*>(a.extra_rows + row * a.c + pc * 8));
                        e1[u] = __ldg(reinterpret_cast<const float4*>(a.extra_rows + row * a.c + pc * 8 + 4));
                    }
                    if (a.extra16)  // gradient a dense consumer (the next fused stack) left in bf16 row form
                        q16[u] = __ldg(reinterpret_cast<const uint4*>(a.extra16 + row * a.extra16_ld + pc * 8));
                }
            }
#pragma unroll
            for (int u = 0; u < RB; ++u) {
                if (!ok[u]) continue;
                const int r = rb + u * rstep;
                const size_t row = (size_t)b * a.s + s0 + r;
                const uint32_t* v = reinterpret_cast<const uint32_t*>(&qy[u]);
                const uint32_t* ew = reinterpret_cast<const uint32_t*>(&q16[u]);
                float ex[8] = {e0[u].x, e0[u].y, e0[u].z, e0[u].w, e1[u].x, e1[u].y, e1[u].z, e1[u].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float2 f = bf2_to_f2(ew[e]);
                    ex[2 * e] += f.x; ex[2 * e + 1] += f.y;
                }
                uint4 o;
                uint32_t* ov = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float2 f = h2_to_f2(v[e]);
                    float d0 = ex[2 * e], d1 = ex[2 * e + 1];
                    if (a.dout_cm) {
                        d0 += tile_dyn[r * ldt + (2 * e) * pieces + pc];
                        d1 += tile_dyn[r * ldt + (2 * e + 1) * pieces + pc];
                    }
                    d0 = fmaf(f.x, sc[2 * e], sh[2 * e]) > 0.f ? d0 : 0.f;
                    d1 = fmaf(f.y, sc[2 * e + 1], sh[2 * e + 1]) > 0.f ? d1 : 0.f;
                    p1[2 * e] += d0; p1[2 * e + 1] += d1;
                    p2[2 * e] = fmaf(d0, (f.x - mu[2 * e]) * rs[2 * e], p2[2 * e]);
                    p2[2 * e + 1] = fmaf(d1, (f.y - mu[2 * e + 1]) * rs[2 * e + 1], p2[2 * e + 1]);
                    ov[e] = f2_to_bf2(d0, d1);
                }
                *reinterpret_cast<uint4*>(a.dz + row * a.dz_ld + pc * 8) = o;
            }
        }
        __syncthreads();
    }
    // rstep threads share a piece: combine them through shared memory, then one atomic per channel per CTA
    float* red = tile_dyn;  // [rstep][2][C]
    for (int e = 0; e < 8; ++e) {
        red[(r0 * 2 + 0) * a.c + pc * 8 + e] = p1[e];
        red[(r0 * 2 + 1) * a.c + pc * 8 + e] = p2[e];
    }
    __syncthreads();
    for (int i = tid; i < 2 * a.c; i += blockDim.x) {
        float tsum = 0.f;
        for (int j = 0; j < rstep; ++j) tsum += red[j * 2 * a.c + i];
        if (tsum != 0.f) atomicAdd(a.sums + i, tsum);
    }
}

template <int GT>
void pool_grid(const PoolArgs& a, dim3& grid) {
    const int s_tiles = (a.s + GT - 1) / GT;
    grid = dim3(a.k == 1 ? (s_tiles < 8 ? s_tiles : 8) : s_tiles, (a.c + 63) / 64, a.b);
}

// ------------------------------------------------------------------ rows backward -----------
struct SaBwdArgs {
    int b, n, s, k;
    const int* idx;
    const bf16* dx; int dx_ld;
    int feat_c; float* dfeat_cm;   // (B,feat_c,N) zeroed, atomics -- or, feat_rows_major, [B*N][feat_c]
    int feat_rows_major;
    int cen_c; float* dcen_cm;     // (B,cen_c,S)  zeroed, atomics
    int xyz_first;
};
// One CTA per group (b, s); warps walk the group's K rows.
//   feature columns  -> scattered to the gathered point's gradient: row-form destination = coalesced 16-byte vector
//                       reductions (lane = 4 channels); channel-major destination = scalar atomics
//   centre columns   -> every row of the group adds into the SAME (b, :, s) column, so the K rows are summed in
//                       registers first and the group issues one atomic per channel (not K same-address atomics)
__global__ void __launch_bounds__(kThreads) sa_rows_bwd_kernel(const SaBwdArgs a) {
    pdl_enter();  // programmatic dependent launch (pn2_common.cuh): first statement, before any memory access
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int grp = blockIdx.x;  // b * s + s_idx
    const int s = grp % a.s, b = grp / a.s;
    const int fc = a.dfeat_cm ? a.feat_c : 0, cc = a.dcen_cm ? a.cen_c : 0;
    const int f0 = a.xyz_first ? 3 : 0, c0 = a.feat_c + 3;
    const bf16* dgrp = a.dx + (size_t)grp * a.k * a.dx_ld;
    // few groups (group-all): gridDim.y CTAs share a group's rows
    const int kchunk = (a.k + gridDim.y - 1) / gridDim.y;
    const int k_lo = blockIdx.y * kchunk, k_hi = min(a.k, k_lo + kchunk);
    if (fc > 0) {
        const bool vec = a.feat_rows_major && (fc & 3) == 0 && (f0 & 3) == 0 && (a.dx_ld & 3) == 0;
        for (int kk = k_lo + warp; kk < k_hi; kk += kThreads / 32) {
            const int j = a.idx ? __ldg(a.idx + (size_t)grp * a.k + kk) : kk;
            const bf16* d = dgrp + (size_t)kk * a.dx_ld + f0;
            if (vec) {
                float* dst = a.dfeat_cm + ((size_t)b * a.n + j) * fc;
                for (int col = lane * 4; col < fc; col += 128) {
                    const uint2 q = __ldg(reinterpret_cast<const uint2*>(d + col));
                    const float2 v0 = bf2_to_f2(q.x), v1 = bf2_to_f2(q.y);
                    if (q.x | q.y) red_add_v4(dst + col, v0.x, v0.y, v1.x, v1.y);
                }
            } else if (a.feat_rows_major) {
                float* dst = a.dfeat_cm + ((size_t)b * a.n + j) * fc;
                for (int col = lane; col < fc; col += 32) {
                    const float v = bf_to_f(d[col]);
                    if (v != 0.f) atomicAdd(dst + col, v);
                }
            } else {
                for (int col = lane; col < fc; col += 32) {
                    const float v = bf_to_f(d[col]);
                    if (v != 0.f) atomicAdd(a.dfeat_cm + ((size_t)b * fc + col) * a.n + j, v);
                }
            }
        }
    }
    if (cc > 0) {
        const bf16* d = dgrp + c0;
        if (((c0 | a.dx_ld | cc) & 1) == 0) {
            for (int col = threadIdx.x * 2; col < cc; col += kThreads * 2) {
                float t0 = 0.f, t1 = 0.f;
#pragma unroll 8
                for (int kk = k_lo; kk < k_hi; ++kk) {
                    const float2 v = bf2_to_f2(__ldg(reinterpret_cast<const uint32_t*>(d + (size_t)kk * a.dx_ld + col)));
                    t0 += v.x; t1 += v.y;
                }
                if (t0 != 0.f) atomicAdd(a.dcen_cm + ((size_t)b * cc + col) * a.s + s, t0);
                if (t1 != 0.f) atomicAdd(a.dcen_cm + ((size_t)b * cc + col + 1) * a.s + s, t1);
            }
        } else {
            for (int col = threadIdx.x; col < cc; col += kThreads) {
                float t = 0.f;
#pragma unroll 8
                for (int kk = k_lo; kk < k_hi; ++kk) t += bf_to_f(d[(size_t)kk * a.dx_ld + col]);
                if (t != 0.f) atomicAdd(a.dcen_cm + ((size_t)b * cc + col) * a.s + s, t);
            }
        }
    }
}

struct FpBwdArgs {
    int b, n, s;
    const int* idx; const float* dist2;
    const bf16* dx; int dx_ld;
    int skip_c; float* dskip_cm;       // (B,skip_c,N) plain stores | rows [B*N][skip_c] accumulated (skip_rows_major) | null
    int skip_rows_major;
    int coarse_c; float* dcoarse_rows; // (B*S, coarse_c) zeroed, atomics | null
};
__global__ void __launch_bounds__(kThreads) fp_rows_bwd_kernel(const FpBwdArgs a) {
    pdl_enter();  // programmatic dependent launch (pn2_common.cuh): first statement, before any memory access
    const int lane = threadIdx.x & 31;
    const unsigned row = blockIdx.x * (unsigned)(kThreads / 32) + (threadIdx.x >> 5);  // < 2^31 rows (host check)
    const unsigned total = (unsigned)a.b * (unsigned)a.n;
    if (row >= total) return;
    const int b = (int)(row / (unsigned)a.n), i = (int)(row - (unsigned)b * (unsigned)a.n);
    const bf16* d = a.dx + (size_t)row * a.dx_ld;
    if (a.dskip_cm) {
        if (a.skip_rows_major) {  // fp32 rows [B*N][skip_c] shared with other consumers of the same producer: accumulate
            float* dst = a.dskip_cm + (size_t)row * a.skip_c;
            if ((a.skip_c & 3) == 0 && (a.dx_ld & 3) == 0 && (reinterpret_cast<uintptr_t>(a.dskip_cm) & 15) == 0) {
                for (int col = lane * 4; col < a.skip_c; col += 128) {
                    const uint2 q = __ldg(reinterpret_cast<const uint2*>(d + col));
                    const float2 v0 = bf2_to_f2(q.x), v1 = bf2_to_f2(q.y);
                    if (q.x | q.y) red_add_v4(dst + col, v0.x, v0.y, v1.x, v1.y);
                }
            } else {
                for (int col = lane; col < a.skip_c; col += 32) {
                    const float v = bf_to_f(d[col]);
                    if (v != 0.f) atomicAdd(dst + col, v);
                }
            }
        } else {
            for (int col = lane; col < a.skip_c; col += 32)
                a.dskip_cm[((size_t)b * a.skip_c + col) * a.n + i] = bf_to_f(d[col]);
        }
    }
    if (a.dcoarse_rows) {
        int id[3] = {0, 0, 0};
        float w[3] = {1.f, 0.f, 0.f};
        if (a.s > 1) {
#pragma unroll
            for (int j = 0; j < 3; ++j) id[j] = a.idx[(size_t)row * 3 + j];
            nn_weights(a.dist2 + (size_t)row * 3, w);
        }
        if ((a.coarse_c & 3) == 0 && (reinterpret_cast<uintptr_t>(a.dcoarse_rows) & 15) == 0) {
            // a lane owns four consecutive channels: 16-byte vector reductions (a quarter of the atomic operations)
            for (int col = lane * 4; col < a.coarse_c; col += 128) {
                const bf16* dp = d + a.skip_c + col;
                const float v0 = bf_to_f(dp[0]), v1 = bf_to_f(dp[1]), v2 = bf_to_f(dp[2]), v3 = bf_to_f(dp[3]);
                if (v0 == 0.f && v1 == 0.f && v2 == 0.f && v3 == 0.f) continue;
                if (a.s > 1) {
#pragma unroll
                    for (int j = 0; j < 3; ++j)
                        red_add_v4(a.dcoarse_rows + ((size_t)b * a.s + id[j]) * a.coarse_c + col, v0 * w[j], v1 * w[j],
                                   v2 * w[j], v3 * w[j]);
                } else {
                    red_add_v4(a.dcoarse_rows + (size_t)b * a.coarse_c + col, v0, v1, v2, v3);
                }
            }
        } else {
            for (int col = lane; col < a.coarse_c; col += 32) {
                const float v = bf_to_f(d[a.skip_c + col]);
                if (v == 0.f) continue;
                if (a.s > 1) {
#pragma unroll
                    for (int j = 0; j < 3; ++j)
                        atomicAdd(a.dcoarse_rows + ((size_t)b * a.s + id[j]) * a.coarse_c + col, v * w[j]);
                } else {
                    atomicAdd(a.dcoarse_rows + (size_t)b * a.coarse_c + col, v);
                }
            }
        }
    }
}

// grid of the K == 1 transposing kernels: as many CTAs as fit the GPU at once (shared memory bound), tiles dealt round-robin
unsigned k1_grid(long long tiles, size_t smem_bytes, int reg_cap) {
    long long per_sm = (long long)(227 * 1024) / (long long)(smem_bytes + 1024);
    if (per_sm < 1) per_sm = 1;
    if (per_sm > reg_cap) per_sm = reg_cap;  // CTAs per SM the register file allows
    const long long slots = 148 * per_sm;
    return (unsigned)(tiles < slots ? tiles : slots);
}

RowSrc mk_src(const void* p, const void* lo, int c, int ld, const float* scale, const float* shift) {
    RowSrc s;
    s.p = (const act_t*)p; s.lo = (const act_t*)lo; s.c = c; s.ld = ld; s.scale = scale; s.shift = shift;
    return s;
}
unsigned warp_blocks(long long rows) { return (unsigned)((rows + (kThreads / 32) - 1) / (kThreads / 32)); }

}  // namespace
}  // namespace pn2

using namespace pn2;

extern "C" int pn2_to_rows_x2(int b, int c, int n, const float* src, const float* sub_sums, float sub_scale, void* dst,
                              void* dst_lo, int ld, pn2_stream_t stream) {
    if (b < 0 || c < 0 || n < 0 || ld < c) return fail_arg("pn2_to_rows", "bad size");
    if (b == 0 || n == 0 || ld == 0) return 0;
    if (!src || !dst) return fail_arg("pn2_to_rows", "null pointer");
    if (b > 65535) return fail_arg("pn2_to_rows", "b > 65535");
    dim3 grid((n + 31) / 32, (ld + 31) / 32, b);
    launch_k(to_rows_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, c, n, ld, src, sub_sums, sub_scale, (act_t*)dst, (act_t*)dst_lo);
    PN2_CHECK_LAUNCH("to_rows_kernel");
    return 0;
}

extern "C" int pn2_to_rows(int b, int c, int n, const float* src, const float* sub_sums, float sub_scale, void* dst,
                           int ld, pn2_stream_t stream) {
    return pn2_to_rows_x2(b, c, n, src, sub_sums, sub_scale, dst, nullptr, ld, stream);
}

extern "C" int pn2_sa_build_rows(int b, int n, int s, int k, const float* xyz, const float* new_xyz, const int* idx,
                                 const void* feat, int feat_c, int feat_ld, const float* feat_scale,
                                 const float* feat_shift, const void* cen, int cen_c, int cen_ld,
                                 const float* cen_scale, const float* cen_shift, int xyz_first, void* out, int out_ld,
                                 pn2_stream_t stream) {
    return pn2_sa_build_rows_x2(b, n, s, k, xyz, new_xyz, idx, feat, nullptr, feat_c, feat_ld, feat_scale, feat_shift, cen,
                                nullptr, cen_c, cen_ld, cen_scale, cen_shift, xyz_first, out, nullptr, out_ld, stream);
}

extern "C" int pn2_sa_build_rows_x2(int b, int n, int s, int k, const float* xyz, const float* new_xyz, const int* idx,
                                    const void* feat, const void* feat_lo, int feat_c, int feat_ld,
                                    const float* feat_scale, const float* feat_shift, const void* cen,
                                    const void* cen_lo, int cen_c, int cen_ld, const float* cen_scale,
                                    const float* cen_shift, int xyz_first, void* out, void* out_lo, int out_ld,
                                    pn2_stream_t stream) {
    if (b < 0 || n <= 0 || s <= 0 || k <= 0) return fail_arg("pn2_sa_build_rows", "bad size");
    if (b == 0) return 0;
    if (!xyz || !out) return fail_arg("pn2_sa_build_rows", "null pointer");
    const int need = (feat ? feat_c : 0) + 3 + (cen ? cen_c : 0);
    if (out_ld < need || out_ld % 8) return fail_arg("pn2_sa_build_rows", "out_ld too small or not a multiple of 8");
    if (xyz_first && cen) return fail_arg("pn2_sa_build_rows", "centre features are not part of the group-all layout");
    if (!idx && k != n) return fail_arg("pn2_sa_build_rows", "identity grouping needs k == n");
    if ((long long)b * s * k >= (1LL << 31)) return fail_arg("pn2_sa_build_rows", "more than 2^31 rows");
    SaBuildArgs a;
    a.b = b; a.n = n; a.s = s; a.k = k; a.xyz = xyz; a.new_xyz = new_xyz; a.idx = idx;
    a.feat = mk_src(feat, feat_lo, feat_c, feat_ld, feat_scale, feat_shift);
    a.cen = mk_src(cen, cen_lo, cen_c, cen_ld, cen_scale, cen_shift);
    a.xyz_first = xyz_first; a.out = (act_t*)out; a.out_lo = (act_t*)out_lo; a.out_ld = out_ld;
    const int gpr = out_ld >> 3, rpw = gpr >= 32 ? 1 : 32 / gpr;
    launch_k(sa_build_rows_kernel, dim3(warp_blocks(((long long)b * s * k + rpw - 1) / rpw)), dim3(kThreads), 0, (cudaStream_t)stream, a);
    PN2_CHECK_LAUNCH("sa_build_rows_kernel");
    return 0;
}

extern "C" int pn2_fp_build_rows(int b, int n, int s, const void* skip, int skip_c, int skip_ld,
                                 const float* skip_scale, const float* skip_shift, const void* coarse, int coarse_c,
                                 int coarse_ld, const float* coarse_scale, const float* coarse_shift, const int* idx,
                                 const float* dist2, void* out, int out_ld, pn2_stream_t stream) {
    return pn2_fp_build_rows_x2(b, n, s, skip, nullptr, skip_c, skip_ld, skip_scale, skip_shift, coarse, nullptr, coarse_c,
                                coarse_ld, coarse_scale, coarse_shift, idx, dist2, out, nullptr, out_ld, stream);
}

extern "C" int pn2_fp_build_rows_x2(int b, int n, int s, const void* skip, const void* skip_lo, int skip_c, int skip_ld,
                                    const float* skip_scale, const float* skip_shift, const void* coarse,
                                    const void* coarse_lo, int coarse_c, int coarse_ld, const float* coarse_scale,
                                    const float* coarse_shift, const int* idx, const float* dist2, void* out,
                                    void* out_lo, int out_ld, pn2_stream_t stream) {
    if (b < 0 || n <= 0 || s <= 0) return fail_arg("pn2_fp_build_rows", "bad size");
    if (b == 0) return 0;
    if (!coarse || !out || (s > 1 && (!idx || !dist2))) return fail_arg("pn2_fp_build_rows", "null pointer");
    if (out_ld < (skip ? skip_c : 0) + coarse_c || out_ld % 8) return fail_arg("pn2_fp_build_rows", "bad out_ld");
    if ((long long)b * n >= (1LL << 31)) return fail_arg("pn2_fp_build_rows", "more than 2^31 rows");
    FpBuildArgs a;
    a.b = b; a.n = n; a.s = s;
    a.skip = mk_src(skip, skip_lo, skip_c, skip_ld, skip_scale, skip_shift);
    a.coarse = mk_src(coarse, coarse_lo, coarse_c, coarse_ld, coarse_scale, coarse_shift);
    a.idx = idx; a.dist2 = dist2; a.out = (act_t*)out; a.out_lo = (act_t*)out_lo; a.out_ld = out_ld;
    int lanes = 32;  // lanes per row: the smallest power of two covering the 8-channel pieces of a coarse row
    while (lanes > 4 && lanes / 2 >= (coarse_c + 7) / 8) lanes >>= 1;
    launch_k(fp_build_rows_kernel, dim3(warp_blocks(((long long)b * n + 32 / lanes - 1) / (32 / lanes))), dim3(kThreads), 0, (cudaStream_t)stream, a, lanes);
    PN2_CHECK_LAUNCH("fp_build_rows_kernel");
    return 0;
}

extern "C" int pn2_pool_fwd(int b, int s, int k, int c, const void* y, int y_ld, const float* scale,
                            const float* shift, float* out_cm, float* chan_sums, int* argmax, pn2_stream_t stream) {
    return pn2_pool_fwd_x2(b, s, k, c, y, nullptr, y_ld, scale, shift, out_cm, chan_sums, argmax, stream);
}

extern "C" int pn2_pool_fwd_x2(int b, int s, int k, int c, const void* y, const void* y_lo, int y_ld, const float* scale,
                               const float* shift, float* out_cm, float* chan_sums, int* argmax, pn2_stream_t stream) {
    if (b < 0 || s <= 0 || k <= 0 || c <= 0 || c % 8) return fail_arg("pn2_pool_fwd", "bad size");
    if (b == 0) return 0;
    if (b > 65535) return fail_arg("pn2_pool_fwd", "b > 65535");
    if (!y || !scale || !shift || !out_cm) return fail_arg("pn2_pool_fwd", "null pointer");
    PoolArgs a{};
    a.b = b; a.s = s; a.k = k; a.c = c; a.y = (const act_t*)y; a.y_lo = (const act_t*)y_lo; a.y_ld = y_ld;
    a.scale = scale; a.shift = shift;
    a.out_cm = out_cm; a.chan_sums = chan_sums; a.argmax = argmax;
    dim3 grid;
    if (k == 1 && !chan_sums && c <= 1024) {
        const int s_tiles = (s + kRowTile - 1) / kRowTile;
        const size_t smem = (size_t)kRowTile * (c + 1) * sizeof(float);
        static DeviceOnce once;
        if (once.first()) {
            PN2_CHECK(cudaFuncSetAttribute(rows_to_cm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           kRowTile * 1025 * 4),
                      "rows_to_cm: cudaFuncSetAttribute");
        }
        const int pieces = c / 8;
        const int threads = pieces >= kThreads ? pieces : (kThreads / pieces) * pieces;  // multiple of pieces
        grid = dim3(k1_grid((long long)b * s_tiles, smem, 4));
        launch_k(rows_to_cm_kernel, dim3(grid), dim3(threads), smem, (cudaStream_t)stream, a);
    } else if (k > 1 && s <= 64) {
        const int groups = b * s;
        launch_k(pool_fwd_grp_kernel, dim3(groups < 592 ? groups : 592), dim3(kThreads), 0, (cudaStream_t)stream, a);
    } else {
        pool_grid<32>(a, grid);
        launch_k(pool_fwd_kernel<32>, dim3(grid), dim3(kThreads), 0, (cudaStream_t)stream, a);
    }
    PN2_CHECK_LAUNCH("pool_fwd_kernel");
    return 0;
}

extern "C" int pn2_pool_finalize(int b, int s, int k, int c, const float* pool_val, const int* pool_arg, const void* y,
                                 int y_ld, const float* scale, const float* shift, float* out_cm, float* chan_sums,
                                 pn2_stream_t stream) {
    if (b < 0 || s <= 0 || k <= 0 || c <= 0) return fail_arg("pn2_pool_finalize", "bad size");
    if (b == 0) return 0;
    if (b > 65535) return fail_arg("pn2_pool_finalize", "b > 65535");
    if (!pool_val || !pool_arg || !y || !scale || !shift || !out_cm) return fail_arg("pn2_pool_finalize", "null pointer");
    dim3 grid((s + 31) / 32, (c + 31) / 32, b);
    launch_k(pool_finalize_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, s, k, c, pool_val, pool_arg, (const act_t*)y, y_ld, scale, shift,
                                                                  out_cm, chan_sums);
    PN2_CHECK_LAUNCH("pool_finalize_kernel");
    return 0;
}

extern "C" int pn2_pool_bwd(int b, int s, int k, int c, const float* dout_cm, const float* extra_rows,
                            const void* extra_rows16, int extra16_ld, const void* y, int y_ld, const float* scale, const float* shift, const float* mean, const float* rstd,
                            const int* argmax, void* dz, int dz_ld, float* sums, pn2_stream_t stream) {
    if (b < 0 || s <= 0 || k <= 0 || c <= 0 || c % 8) return fail_arg("pn2_pool_bwd", "bad size");
    if (b == 0) return 0;
    if (b > 65535) return fail_arg("pn2_pool_bwd", "b > 65535");
    if ((!dout_cm && !extra_rows && !extra_rows16) || !y || !scale || !shift || !mean || !rstd || !dz || !sums)
        return fail_arg("pn2_pool_bwd", "null pointer");
    if (k > 1 && !argmax) return fail_arg("pn2_pool_bwd", "argmax required when k > 1");
    if (extra_rows16 && k != 1) return fail_arg("pn2_pool_bwd", "extra_rows16 needs k == 1");
    if ((extra_rows || extra_rows16) && k == 1 && c > 1024) return fail_arg("pn2_pool_bwd", "extra rows with k == 1 need c <= 1024");
    if (extra_rows16 && (extra16_ld < c || extra16_ld % 8)) return fail_arg("pn2_pool_bwd", "bad extra16_ld");
    PoolArgs a{};
    a.b = b; a.s = s; a.k = k; a.c = c; a.y = (const act_t*)y; a.y_ld = y_ld; a.scale = scale; a.shift = shift;
    a.mean = mean; a.rstd = rstd; a.argmax = const_cast<int*>(argmax); a.dout_cm = dout_cm; a.extra_rows = extra_rows;
    a.extra16 = (const bf16*)extra_rows16; a.extra16_ld = extra16_ld;
    a.dz = (bf16*)dz; a.dz_ld = dz_ld; a.sums = sums;
    dim3 grid;
    if (k == 1 && c <= 1024) {
        const int s_tiles = (s + kRowTile - 1) / kRowTile;
        const size_t smem = (size_t)kRowTile * (c + 1) * sizeof(float);
        static DeviceOnce once;
        if (once.first()) {
            PN2_CHECK(cudaFuncSetAttribute(cm_to_rows_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           kRowTile * 1025 * 4),
                      "cm_to_rows_bwd: cudaFuncSetAttribute");
        }
        const int pieces = c / 8;
        const int threads = pieces >= kThreads ? pieces : (kThreads / pieces) * pieces;
        const size_t red = (size_t)2 * threads * 8 * sizeof(float);  // the final [rstep][2][C] reduction reuses the tile
        grid = dim3(k1_grid((long long)b * s_tiles, smem > red ? smem : red, 2));
        launch_k(cm_to_rows_bwd_kernel, dim3(grid), dim3(threads), smem > red ? smem : red, (cudaStream_t)stream, a);
    } else if (k > 1 && s <= 64) {
        const int groups = b * s;
        launch_k(pool_bwd_grp_kernel, dim3(groups < 592 ? groups : 592), dim3(kThreads), 0, (cudaStream_t)stream, a);
    } else {
        pool_grid<32>(a, grid);
        launch_k(pool_bwd_kernel<32>, dim3(grid), dim3(kThreads), 0, (cudaStream_t)stream, a);
    }
    PN2_CHECK_LAUNCH("pool_bwd_kernel");
    return 0;
}

extern "C" int pn2_sa_rows_bwd(int b, int n, int s, int k, const int* idx, const void* dx, int dx_ld, int feat_c,
                               float* dfeat_cm, int feat_rows_major, int cen_c, float* dcen_cm, int xyz_first,
                               pn2_stream_t stream) {
    if (b < 0 || n <= 0 || s <= 0 || k <= 0) return fail_arg("pn2_sa_rows_bwd", "bad size");
    if (b == 0 || (!dfeat_cm && !dcen_cm)) return 0;
    if (!dx) return fail_arg("pn2_sa_rows_bwd", "null pointer");
    SaBwdArgs a;
    a.b = b; a.n = n; a.s = s; a.k = k; a.idx = idx; a.dx = (const bf16*)dx; a.dx_ld = dx_ld;
    a.feat_c = feat_c; a.dfeat_cm = dfeat_cm; a.feat_rows_major = feat_rows_major;
    a.cen_c = cen_c; a.dcen_cm = dcen_cm; a.xyz_first = xyz_first;
    const long long groups = (long long)b * s;
    long long gy = groups >= 296 ? 1 : (592 + groups - 1) / groups;
    if (gy > (k + 7) / 8) gy = (k + 7) / 8;
    launch_k(sa_rows_bwd_kernel, dim3(dim3((unsigned)groups, (unsigned)gy)), dim3(kThreads), 0, (cudaStream_t)stream, a);
    PN2_CHECK_LAUNCH("sa_rows_bwd_kernel");
    return 0;
}

extern "C" int pn2_fp_rows_bwd(int b, int n, int s, const int* idx, const float* dist2, const void* dx, int dx_ld,
                               int skip_c, float* dskip, int skip_rows_major, int coarse_c, float* dcoarse_rows,
                               pn2_stream_t stream) {
    if (b < 0 || n <= 0 || s <= 0) return fail_arg("pn2_fp_rows_bwd", "bad size");
    if (b == 0 || (!dskip && !dcoarse_rows)) return 0;
    if (!dx || (s > 1 && dcoarse_rows && (!idx || !dist2))) return fail_arg("pn2_fp_rows_bwd", "null pointer");
    if ((long long)b * n >= (1LL << 31)) return fail_arg("pn2_fp_rows_bwd", "more than 2^31 rows");
    FpBwdArgs a;
    a.b = b; a.n = n; a.s = s; a.idx = idx; a.dist2 = dist2; a.dx = (const bf16*)dx; a.dx_ld = dx_ld;
    a.skip_c = skip_c; a.dskip_cm = dskip; a.skip_rows_major = skip_rows_major; a.coarse_c = coarse_c;
    a.dcoarse_rows = dcoarse_rows;
    launch_k(fp_rows_bwd_kernel, dim3(warp_blocks((long long)b * n)), dim3(kThreads), 0, (cudaStream_t)stream, a);
    PN2_CHECK_LAUNCH("fp_rows_bwd_kernel");
    return 0;
}
