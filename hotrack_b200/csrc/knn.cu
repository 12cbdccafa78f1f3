// knn.cu -- k-nearest-neighbour and 3-nearest-neighbour search for sm_100a.
//
// Replaces knn_kernel_fast and three_nn_kernel_fast (reference
// network/models/pointnet_lib/src/interpolate_gpu.cu:9-57 and :81-124).
//
// Result definition shared by both reference kernels: scan the known points in
// ascending index order and insert with a strict `<` -- i.e. the k smallest
// (distance, index) pairs in lexicographic order.  The reference compares in
// fp64, but on exactly converted fp32 values, so fp32 (here: integer bit
// pattern) compares are identical.  Its 1e40 sentinel means a distance of +inf
// or NaN is never inserted and unfilled slots return (+inf, 0).
//
// knn: the reference gives each query ONE THREAD with a 2400-byte local-memory
// insertion sort (HandTrackNet has 21 queries per cloud => 21 live threads per
// CTA).  Here one WARP owns a query: 32 points per step, a 64-bit key
// (dist bits << 32 | index) per point, candidates below the current k-th key
// are ballot-compacted into a shared-memory buffer and folded into the sorted
// top-k by a warp bitonic sort when the buffer fills (expected O(k log(M/k))
// candidates in total).  Known points are staged through shared memory by the
// TMA engine (cp.async.bulk) and shared by the CTA's queries.
//
// three_nn: one thread per query (there are B*N of them); the known set is
// staged once per CTA as float4 so the inner loop is one broadcast LDS.128 +
// 6 FP ops + a branch-free top-3 update.
#include "pn2_common.cuh"

#include <cstdlib>

namespace pn2 {
namespace {

constexpr unsigned long long kMaxKey = ~0ull;
constexpr unsigned kInfBits = 0x7f800000u;

__device__ __forceinline__ unsigned long long make_key(float d, int i) {
    return ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)i;
}

// ---------------------------------------------------------------- knn -------
constexpr int kKnnWarps = 8;
constexpr int kKnnTile = 1024;  // known points per shared-memory tile

// Sorts keys[0..n) ascending; n a power of two >= 64; executed by one warp.
__device__ __forceinline__ void warp_bitonic_sort(unsigned long long* keys, int n, int lane) {
    for (int size = 2; size <= n; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = lane; i < (n >> 1); i += 32) {
                const int lo = 2 * i - (i & (stride - 1));
                const int hi = lo + stride;
                const unsigned long long a = keys[lo], c = keys[hi];
                const bool up = (lo & size) == 0;
                if ((a > c) == up) { keys[lo] = c; keys[hi] = a; }
            }
            __syncwarp();
        }
    }
}

// COOP = false: a warp per query, 8 queries per CTA (many queries).  COOP = true: ONE query per CTA, its 8 warps scan
// interleaved 32-point chunks of every tile into their own top-k lists, which a block-wide bitonic sort merges at the end
// (keys are unique -- the index is part of the key -- so the merged first k are exactly the single-warp result).  For
// the few-queries case (21 joints per cloud: 3 CTAs per cloud, 90 us at B = 32 and at B = 1 alike): 8x shorter chains
// and 8x more CTAs.
template <bool COOP>
__global__ void __launch_bounds__(kKnnWarps * 32)
knn_kernel(int n, int m, int k, int nsort, const float* __restrict__ unknown, const float* __restrict__ known,
           float* __restrict__ dist2, int* __restrict__ idx) {
    pdl_enter();  // programmatic dependent launch (pn2_common.cuh): first statement, before any memory access
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout: [2][kKnnTile*3] floats | [kKnnWarps][nsort] keys
    float* s_pts = reinterpret_cast<float*>(smem_raw);
    unsigned long long* s_keys = reinterpret_cast<unsigned long long*>(smem_raw + 2 * kKnnTile * 3 * sizeof(float));
    __shared__ __align__(8) uint64_t s_bar[2];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y;
    const int q = COOP ? (int)blockIdx.x : (int)blockIdx.x * kKnnWarps + warp;
    const bool active = q < n;
    const float* pts = known + (size_t)b * m * 3;
    const int ntiles = (m + kKnnTile - 1) / kKnnTile;
    const bool use_bulk = (reinterpret_cast<uintptr_t>(pts) & 15) == 0;

    float ux = 0.f, uy = 0.f, uz = 0.f;
    if (active) {
        const float* u = unknown + ((size_t)b * n + q) * 3;
        ux = u[0]; uy = u[1]; uz = u[2];
    }

    unsigned long long* keys = s_keys + (size_t)warp * nsort;
    // sorted part [0,k): sentinel (+inf, 0); everything else: kMaxKey
    const unsigned long long sentinel = (unsigned long long)kInfBits << 32;
    for (int i = lane; i < nsort; i += 32) keys[i] = i < k ? sentinel : kMaxKey;
    __syncwarp();
    unsigned long long thresh = sentinel;  // current k-th smallest key
    const int cap = nsort - k;             // candidate slots keys[k .. nsort)
    int ncand = 0;

    auto tile_count = [&](int t) { return min(kKnnTile, m - t * kKnnTile); };
    auto issue = [&](int t) {
        const int cnt = tile_count(t);
        const uint32_t bytes = ((uint32_t)cnt * 12u) & ~15u;
        if (bytes) {
            mbar_arrive_expect_tx(&s_bar[t & 1], bytes);
            bulk_g2s(s_pts + (t & 1) * kKnnTile * 3, pts + (size_t)t * kKnnTile * 3, bytes, &s_bar[t & 1]);
        } else {
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&s_bar[t & 1])) : "memory");
        }
    };
    if (use_bulk) {
        if (tid == 0) {
            mbar_init(&s_bar[0], 1);
            mbar_init(&s_bar[1], 1);
            mbar_fence_init();
        }
        __syncthreads();
        if (tid == 0) {
            issue(0);
            if (ntiles > 1) issue(1);
        }
    }

    auto fold = [&]() {  // merge candidates into the sorted top-k, refresh the threshold
        warp_bitonic_sort(keys, nsort, lane);
        thresh = keys[k - 1];
        __syncwarp();
        for (int i = k + lane; i < nsort; i += 32) keys[i] = kMaxKey;
        __syncwarp();
        ncand = 0;
    };

    for (int t = 0; t < ntiles; ++t) {
        const int cnt = tile_count(t);
        float* sp = s_pts + (t & 1) * kKnnTile * 3;
        if (use_bulk) {
            const int bulk_floats = (int)((((uint32_t)cnt * 12u) & ~15u) >> 2);
            const int tail = cnt * 3 - bulk_floats;
            if (tid < tail) sp[bulk_floats + tid] = pts[(size_t)t * kKnnTile * 3 + bulk_floats + tid];
            mbar_wait(&s_bar[t & 1], (t >> 1) & 1);
        } else {
            for (int i = tid; i < cnt * 3; i += blockDim.x) sp[i] = pts[(size_t)t * kKnnTile * 3 + i];
        }
        __syncthreads();

        if (active) {
            const int base = t * kKnnTile;
            for (int i = COOP ? warp * 32 : 0; i < cnt; i += COOP ? kKnnWarps * 32 : 32) {
                const int j = i + lane;
                unsigned long long key = kMaxKey;
                if (j < cnt) key = make_key(sqdist(ux, uy, uz, sp[j * 3 + 0], sp[j * 3 + 1], sp[j * 3 + 2]), base + j);
                const bool cand = key < thresh;
                const unsigned mask = __ballot_sync(kFull, cand);
                if (mask) {
                    if (ncand + 32 > cap) fold();  // keep room for a full step
                    const bool still = key < thresh;  // threshold may have tightened
                    const unsigned mask2 = __ballot_sync(kFull, still);
                    if (still) keys[k + ncand + __popc(mask2 & ((1u << lane) - 1u))] = key;
                    ncand += __popc(mask2);
                    __syncwarp();
                }
            }
        }
        __syncthreads();
        if (use_bulk && tid == 0 && t + 2 < ntiles) issue(t + 2);
    }

    if (COOP) {
        if (ncand > 0) fold();  // every warp: [0,k) sorted (sentinels where it found fewer than k), kMaxKey behind
        __syncthreads();
        // block-wide bitonic sort of the 8 lists (8 * nsort keys, a power of two)
        const int total = kKnnWarps * nsort;
        for (int size = 2; size <= total; size <<= 1) {
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                for (int i = tid; i < (total >> 1); i += kKnnWarps * 32) {
                    const int lo = 2 * i - (i & (stride - 1));
                    const int hi = lo + stride;
                    const unsigned long long a = s_keys[lo], c = s_keys[hi];
                    const bool up = (lo & size) == 0;
                    if ((a > c) == up) { s_keys[lo] = c; s_keys[hi] = a; }
                }
                __syncthreads();
            }
        }
        float* od = dist2 + ((size_t)b * n + q) * k;
        int* oi = idx + ((size_t)b * n + q) * k;
        for (int i = tid; i < k; i += kKnnWarps * 32) {
            const unsigned long long key = s_keys[i];
            od[i] = __uint_as_float((unsigned)(key >> 32));
            oi[i] = (int)(unsigned)(key & 0xffffffffu);
        }
        return;
    }
    if (active) {
        if (ncand > 0) fold();
        float* od = dist2 + ((size_t)b * n + q) * k;
        int* oi = idx + ((size_t)b * n + q) * k;
        for (int i = lane; i < k; i += 32) {
            const unsigned long long key = keys[i];
            od[i] = __uint_as_float((unsigned)(key >> 32));
            oi[i] = (int)(unsigned)(key & 0xffffffffu);
        }
    }
}

// ------------------------------------------------------------ three_nn ------
constexpr int kNnThreads = 256;
constexpr int kNnTile = 2048;  // known points per tile (float4 => 32 KB)

__global__ void __launch_bounds__(kNnThreads)
three_nn_kernel(int n, int m, const float* __restrict__ unknown, const float* __restrict__ known,
                float* __restrict__ dist2, int* __restrict__ idx) {
    pdl_enter();  // programmatic dependent launch (pn2_common.cuh): first statement, before any memory access
    __shared__ float4 s_known[kNnTile];
    const int b = blockIdx.y;
    const int q = blockIdx.x * kNnThreads + threadIdx.x;
    const float* pts = known + (size_t)b * m * 3;

    float ux = 0.f, uy = 0.f, uz = 0.f;
    if (q < n) {
        const float* u = unknown + ((size_t)b * n + q) * 3;
        ux = u[0]; uy = u[1]; uz = u[2];
    }
    // 1e40 in the reference: larger than any finite fp32, smaller than +inf.  In
    // bit space: a distance is inserted iff its bits are < the +inf pattern.
    unsigned b1 = kInfBits, b2 = kInfBits, b3 = kInfBits;
    int i1 = 0, i2 = 0, i3 = 0;

    for (int t0 = 0; t0 < m; t0 += kNnTile) {
        const int cnt = min(kNnTile, m - t0);
        __syncthreads();
        for (int i = threadIdx.x; i < cnt; i += kNnThreads) {
            const float* p = pts + (size_t)(t0 + i) * 3;
            s_known[i] = make_float4(p[0], p[1], p[2], 0.f);
        }
        __syncthreads();
#pragma unroll 4
        for (int j = 0; j < cnt; ++j) {
            const float4 p = s_known[j];
            const unsigned d = __float_as_uint(sqdist(ux, uy, uz, p.x, p.y, p.z));
            const int kk = t0 + j;
            // strict-< cascade of interpolate_gpu.cu:109-120, branch-free
            const bool c1 = d < b1, c2 = d < b2, c3 = d < b3;
            b3 = c2 ? b2 : (c3 ? d : b3);
            i3 = c2 ? i2 : (c3 ? kk : i3);
            b2 = c1 ? b1 : (c2 ? d : b2);
            i2 = c1 ? i1 : (c2 ? kk : i2);
            b1 = c1 ? d : b1;
            i1 = c1 ? kk : i1;
        }
    }
    if (q < n) {
        float* od = dist2 + ((size_t)b * n + q) * 3;
        int* oi = idx + ((size_t)b * n + q) * 3;
        od[0] = __uint_as_float(b1); od[1] = __uint_as_float(b2); od[2] = __uint_as_float(b3);
        oi[0] = i1; oi[1] = i2; oi[2] = i3;
    }
}

}  // namespace
}  // namespace pn2

extern "C" int pn2_knn(int b, int n, int m, int k, const float* unknown, const float* known, float* dist2,
                       int* idx, pn2_stream_t stream) {
    using namespace pn2;
    if (b < 0 || n < 0 || m < 0 || k < 0) return fail_arg("pn2_knn", "negative size");
    if (k > PN2_KNN_MAX_K) return fail_arg("pn2_knn", "k > PN2_KNN_MAX_K");
    if (b == 0 || n == 0 || k == 0) return 0;
    if (b > 65535) return fail_arg("pn2_knn", "b > 65535");
    if (!unknown || (!known && m > 0) || !dist2 || !idx) return fail_arg("pn2_knn", "null pointer");
    int kpad = 32;
    while (kpad < k) kpad <<= 1;
    const int nsort = kpad * 2 < 128 ? 128 : kpad * 2;
    const size_t smem = 2 * kKnnTile * 3 * sizeof(float) + (size_t)kKnnWarps * nsort * sizeof(unsigned long long);
    static size_t configured[64] = {};  // per device: the attribute belongs to the device's copy of the function
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    if (smem > 48 * 1024 && smem > configured[dev]) {
        PN2_CHECK(cudaFuncSetAttribute(knn_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                  "knn: cudaFuncSetAttribute");
        PN2_CHECK(cudaFuncSetAttribute(knn_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                  "knn: cudaFuncSetAttribute");
        configured[dev] = smem;
    }
    // few queries (a warp-per-query grid would leave most SMs idle): one query per CTA, its warps share the scan
    const long long wpq_ctas = (long long)((n + kKnnWarps - 1) / kKnnWarps) * b;
    static int force = -2;  // PN2_KNN_COOP=0|1: development switch
    if (force == -2) {
        const char* e = getenv("PN2_KNN_COOP");
        force = e ? atoi(e) : -1;
    }
    // (the cooperative variant does more work in total -- eight looser thresholds, eight lists to fold, the final merge --
    // so it only pays while the warp-per-query grid leaves most SMs idle.  Measured, 21 queries, K = 64: B=1 N=8192
    // 119 -> 68 us, B=4 N=4096 86 -> 44 us, B=8 86 -> 55 us, B=32 86 -> 140 us)
    const bool coop = force >= 0 ? force == 1 : wpq_ctas * 4 <= sm_count();
    if (coop && n <= 65535) {
        launch_k(knn_kernel<true>, dim3(n, b), dim3(kKnnWarps * 32), smem, (cudaStream_t)stream, n, m, k, nsort, unknown, known,
                 dist2, idx);
        PN2_CHECK_LAUNCH("knn_kernel");
        return 0;
    }
    dim3 grid((n + kKnnWarps - 1) / kKnnWarps, b);
    launch_k(knn_kernel<false>, dim3(grid), dim3(kKnnWarps * 32), smem, (cudaStream_t)stream, n, m, k, nsort, unknown, known, dist2, idx);
    PN2_CHECK_LAUNCH("knn_kernel");
    return 0;
}

extern "C" int pn2_three_nn(int b, int n, int m, const float* unknown, const float* known, float* dist2,
                            int* idx, pn2_stream_t stream) {
    using namespace pn2;
    if (b < 0 || n < 0 || m < 0) return fail_arg("pn2_three_nn", "negative size");
    if (b == 0 || n == 0) return 0;
    if (b > 65535) return fail_arg("pn2_three_nn", "b > 65535");
    if (!unknown || (!known && m > 0) || !dist2 || !idx) return fail_arg("pn2_three_nn", "null pointer");
    dim3 grid((n + kNnThreads - 1) / kNnThreads, b);
    launch_k(three_nn_kernel, dim3(grid), dim3(kNnThreads), 0, (cudaStream_t)stream, n, m, unknown, known, dist2, idx);
    PN2_CHECK_LAUNCH("three_nn_kernel");
    return 0;
}
