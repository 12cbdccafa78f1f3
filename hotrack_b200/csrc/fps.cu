// fps.cu -- farthest point sampling for sm_100a.
//
// Replaces furthest_point_sampling_kernel<bs> (reference
// network/models/pointnet_lib/src/sampling_gpu.cu:93-209).  Same result, index
// for index, including ties; different machine mapping:
//
//  * One persistent CTA per cloud.  Coordinates AND the running min-distance
//    array live in registers for all M iterations (the reference re-reads xyz
//    and read-modify-writes `temp` in global memory every iteration).
//  * The reference's tie behaviour is a property of its block reduction: thread
//    t scans k = t, t+bs, ... keeping the first strict maximum, then a shared
//    memory tree in which the LEFT operand wins ties (sampling_gpu.cu:86-91,
//    143-203).  Net effect: among points with the maximal distance the winner
//    minimises ( bitreverse_{log2 bs}(k mod bs), k div bs ), bs being the
//    reference's block size for this N (cuda_utils.h:10-14).  We therefore
//    renumber points into that priority order once ("position" p) and the
//    arg-max becomes "largest value, lowest position" -- two REDUX
//    instructions per warp instead of a 10-step tree.
//  * One __syncthreads per iteration (double-buffered per-warp partials; every
//    warp redundantly finishes the cross-warp stage), against 11 in the
//    reference at bs=1024.
//  * The winner's coordinates come from a position-indexed float4 copy of the
//    cloud in shared memory: one broadcast LDS.128 on the critical path.
//
// Bound: a serial chain of M-1 dependent arg-max steps; per step the CTA issues
// ~10 instructions per point.  HBM traffic is the one-time 12 B/point read.
#include "pn2_common.cuh"

#include <cmath>
#include <cstdlib>

namespace pn2 {
namespace {

constexpr int kNoPos = 0x7fffffff;
constexpr int kMaxSmem = 227 * 1024 - 2048;  // dynamic smem budget next to the static partials

__device__ __forceinline__ int bitrev(int v, int lg) { return lg == 0 ? 0 : (int)(__brev((unsigned)v) >> (32 - lg)); }

// position of original index k:  p = bitrev(k mod bs) * q_cnt + k div bs
__device__ __forceinline__ int pos_of(int k, int lg_bs, int q_cnt) {
    return bitrev(k & ((1 << lg_bs) - 1), lg_bs) * q_cnt + (k >> lg_bs);
}

struct WarpBest {
    int vbits;  // float bits of the (non-negative) distance; -1.0f bits for "nothing"
    int pos;
};

__device__ __forceinline__ WarpBest warp_argmax(int vbits, int pos) {
    const int m = __reduce_max_sync(kFull, vbits);
    const int p = __reduce_min_sync(kFull, vbits == m ? pos : kNoPos);
    return {m, p};
}

// Register-resident kernel: thread t owns positions t, t+T, ..., t+(PPT-1)T.
template <int PPT, int MAXT = 1024>
__global__ void __launch_bounds__(MAXT, 1)
fps_regs_kernel(int n, int m, int lg_bs, int q_cnt, int n_pos, const float* __restrict__ dataset,
                float* __restrict__ temp, int* __restrict__ idxs) {
    pdl_enter();  // programmatic dependent launch (pn2_common.cuh): first statement, before any memory access
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* spts = reinterpret_cast<float4*>(smem_raw);  // [n_pos] (x, y, z, int k | -1)
    __shared__ int2 s_part[2][32];

    const int T = blockDim.x;
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
    const float* ds = dataset + (size_t)blockIdx.x * n * 3;
    float* tp = temp ? temp + (size_t)blockIdx.x * n : nullptr;
    int* out = idxs + (size_t)blockIdx.x * m;

    // Stage the cloud into position order (coalesced global reads, scattered smem writes).
    for (int p = tid; p < n_pos; p += T) spts[p] = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
    __syncthreads();
    for (int k = tid; k < n; k += T) {
        const float x = ds[k * 3 + 0], y = ds[k * 3 + 1], z = ds[k * 3 + 2];
        spts[pos_of(k, lg_bs, q_cnt)] = make_float4(x, y, z, __int_as_float(k));
    }
    __syncthreads();

    float px[PPT], py[PPT], pz[PPT], td[PPT];
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
        const int p = tid + j * T;
        float4 v = p < n_pos ? spts[p] : make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
        const int k = __float_as_int(v.w);
        px[j] = v.x; py[j] = v.y; pz[j] = v.z;
        // invalid slots carry -1: below every real distance, never selected
        td[j] = k >= 0 ? (tp ? tp[k] : 1e10f) : -1.0f;
    }

    float4 cur = spts[0];  // position 0 is original index 0: idxs[:,0] = 0 (sampling_gpu.cu:113-115)
    if (tid == 0) out[0] = 0;

    for (int it = 1; it < m; ++it) {
        int bv = __float_as_int(-1.0f), bp = kNoPos;
#pragma unroll
        for (int j = 0; j < PPT; ++j) {
            const float d = sqdist(px[j], py[j], pz[j], cur.x, cur.y, cur.z);
            // fminf, like the reference's min(): a NaN distance leaves td unchanged.
            // Invalid slots keep -1 because every real d is >= 0 (or NaN).
            const float d2 = fminf(d, td[j]);
            td[j] = d2;
            const int vb = __float_as_int(d2);
            if (vb > bv) { bv = vb; bp = tid + j * T; }  // ascending positions + strict > => lowest position
        }
        const WarpBest w = warp_argmax(bv, bp);
        if (lane == 0) s_part[it & 1][warp] = make_int2(w.vbits, w.pos);
        __syncthreads();
        int2 e = lane < nwarps ? s_part[it & 1][lane] : make_int2(__float_as_int(-1.0f), kNoPos);
        const WarpBest g = warp_argmax(e.x, e.y);
        cur = spts[g.pos];
        if (tid == 0) out[it] = __float_as_int(cur.w);
    }

    if (tp) {
#pragma unroll
        for (int j = 0; j < PPT; ++j) {
            const int p = tid + j * T;
            if (p < n_pos) {
                const int k = __float_as_int(spts[p].w);
                if (k >= 0) tp[k] = td[j];
            }
        }
    }
}

// Streaming fallback for clouds too large for registers / shared memory: same
// arithmetic and tie rule, but xyz and temp are re-read from global memory (L2)
// every iteration like the reference.  Requires temp != nullptr.
__global__ void __launch_bounds__(1024, 1)
fps_stream_kernel(int n, int m, int lg_bs, int q_cnt, const float* __restrict__ dataset,
                  float* __restrict__ temp, int* __restrict__ idxs) {
    __shared__ int2 s_part[2][32];
    const int T = blockDim.x;
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
    const float* ds = dataset + (size_t)blockIdx.x * n * 3;
    float* tp = temp + (size_t)blockIdx.x * n;
    int* out = idxs + (size_t)blockIdx.x * m;

    int old = 0;
    if (tid == 0) out[0] = 0;
    for (int it = 1; it < m; ++it) {
        const float cx = ds[old * 3 + 0], cy = ds[old * 3 + 1], cz = ds[old * 3 + 2];
        int bv = __float_as_int(-1.0f), bp = kNoPos;
        for (int k = tid; k < n; k += T) {
            const float d = sqdist(ds[k * 3 + 0], ds[k * 3 + 1], ds[k * 3 + 2], cx, cy, cz);
            const float d2 = fminf(d, tp[k]);
            tp[k] = d2;
            const int vb = __float_as_int(d2);
            const int p = pos_of(k, lg_bs, q_cnt);
            if (vb > bv || (vb == bv && p < bp)) { bv = vb; bp = p; }
        }
        const WarpBest w = warp_argmax(bv, bp);
        if (lane == 0) s_part[it & 1][warp] = make_int2(w.vbits, w.pos);
        __syncthreads();
        int2 e = lane < nwarps ? s_part[it & 1][lane] : make_int2(__float_as_int(-1.0f), kNoPos);
        const WarpBest g = warp_argmax(e.x, e.y);
        // invert the position:  k = (p mod q_cnt) * bs + bitrev(p div q_cnt)
        old = (g.pos % q_cnt) * (1 << lg_bs) + bitrev(g.pos / q_cnt, lg_bs);
        if (tid == 0) out[it] = old;
    }
}

// cuda_utils.h:10-14, reproduced operation for operation (double log ratio,
// truncation) because it decides the tie order.
int ref_block_size(int work_size) {
    const int pow_2 = (int)(std::log(static_cast<double>(work_size)) / std::log(2.0));
    int v = 1 << pow_2;
    if (v > 1024) v = 1024;
    if (v < 1) v = 1;
    return v;
}

int ilog2(int v) {
    int l = 0;
    while ((1 << (l + 1)) <= v) ++l;
    return l;
}

template <int PPT, int MAXT = 1024>
int launch_regs(int b, int n, int m, int lg_bs, int q_cnt, int n_pos, int threads, const float* dataset,
                float* temp, int* idxs, cudaStream_t stream) {
    const size_t smem = (size_t)n_pos * sizeof(float4);
    static DeviceOnce once;
    if (once.first()) {
        PN2_CHECK(cudaFuncSetAttribute(fps_regs_kernel<PPT, MAXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem),
                  "fps: cudaFuncSetAttribute");
    }
    launch_k(fps_regs_kernel<PPT, MAXT>, dim3(b), dim3(threads), smem, stream, n, m, lg_bs, q_cnt, n_pos, dataset, temp, idxs);
    PN2_CHECK_LAUNCH("fps_regs_kernel");
    return 0;
}

}  // namespace
}  // namespace pn2

extern "C" int pn2_furthest_point_sampling(int b, int n, int m, const float* dataset, float* temp, int* idxs,
                                           pn2_stream_t stream_) {
    using namespace pn2;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (b < 0 || n < 0) return fail_arg("pn2_furthest_point_sampling", "negative size");
    if (m <= 0 || b == 0) return 0;  // sampling_gpu.cu:101: m <= 0 returns without writing
    if (n == 0) return fail_arg("pn2_furthest_point_sampling", "n == 0 with m > 0");
    if (!dataset || !idxs) return fail_arg("pn2_furthest_point_sampling", "null pointer");

    const int bs = ref_block_size(n);
    const int lg_bs = ilog2(bs);
    const int q_cnt = (n + bs - 1) / bs;
    const int n_pos = bs * q_cnt;

    int threads = 32;
    while (threads < n_pos && threads < 1024) threads <<= 1;
    // 2049..4096 points: 512 threads x 8 points beat 1024 x 4 (16 warps to synchronise and to reduce over instead of 32;
    // measured on the step at N = 4096: 105 -> 87 us; 256 x 16 is within noise of that)
    if (n_pos > 2048 && n_pos <= 4096) threads = 512;
    static int forced = -1;
    if (forced < 0) {  // PN2_FPS_THREADS=256|512|1024: development switch, honoured where the cloud still fits the registers
        const char* e = getenv("PN2_FPS_THREADS");
        forced = e ? atoi(e) : 0;
    }
    if ((forced == 256 || forced == 512 || forced == 1024) && forced <= n_pos && (n_pos + forced - 1) / forced <= (forced == 256 ? 16 : 8))
        threads = forced;
    const int ppt = (n_pos + threads - 1) / threads;
    // 1024 threads cap the register file at 64/thread: 8 points (32 state registers) is the limit
    const bool fits = ppt <= 8 && (size_t)n_pos * sizeof(float4) <= (size_t)kMaxSmem;
    if (threads == 256 && ppt > 8 && ppt <= 16 && (size_t)n_pos * sizeof(float4) <= (size_t)kMaxSmem)  // PN2_FPS_THREADS=256
        return launch_regs<16, 256>(b, n, m, lg_bs, q_cnt, n_pos, threads, dataset, temp, idxs, stream);
    if (fits) {
        if (ppt <= 1) return launch_regs<1>(b, n, m, lg_bs, q_cnt, n_pos, threads, dataset, temp, idxs, stream);
        if (ppt <= 2) return launch_regs<2>(b, n, m, lg_bs, q_cnt, n_pos, threads, dataset, temp, idxs, stream);
        if (ppt <= 4) return launch_regs<4>(b, n, m, lg_bs, q_cnt, n_pos, threads, dataset, temp, idxs, stream);
        return launch_regs<8>(b, n, m, lg_bs, q_cnt, n_pos, threads, dataset, temp, idxs, stream);
    }
    if (!temp) return fail_arg("pn2_furthest_point_sampling", "temp must be non-NULL for n > 8192 points");
    fps_stream_kernel<<<b, 1024, 0, stream>>>(n, m, lg_bs, q_cnt, dataset, temp, idxs);
    PN2_CHECK_LAUNCH("fps_stream_kernel");
    return 0;
}
