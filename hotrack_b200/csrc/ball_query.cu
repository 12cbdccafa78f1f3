// ball_query.cu -- radius neighbour search for sm_100a.
//
// Replaces ball_query_kernel_fast (reference
// network/models/pointnet_lib/src/ball_query_gpu.cu:9-45): for every centre the
// first `nsample` point indices (ascending) with d2 < r*r, remaining slots
// padded with the first hit, rows without a hit left untouched.
//
// Mapping (the reference runs ONE THREAD per centre, each streaming all N points
// through L1 with a divergent early exit and an nsample-wide fill loop):
//   * one WARP per centre: 32 candidate points per step, __ballot_sync +
//     prefix popcount give every hit its output slot in ascending index order,
//     so the result is identical and the early exit (nsample found) is
//     warp-uniform;
//   * the cloud is streamed through shared memory in tiles of kTile points by
//     the TMA engine (cp.async.bulk + mbarrier, double-buffered) and shared by
//     the CTA's 8 centres; AoS xyz with stride 3 words is bank-conflict free;
//   * grid = B * ceil(M/8) CTAs (1024 for the SA1 shape: ~7 CTAs per SM).
// Bound: B*M*N fp32 distance evaluations (~9 instructions each); the bytes are
// L2-resident and tiny (12 B/point read once per CTA).
#include "pn2_common.cuh"

namespace pn2 {
namespace {

constexpr int kWarpsPerCta = 8;
constexpr int kTile = 1024;  // points per shared-memory tile (12 KB); two buffers

__global__ void __launch_bounds__(kWarpsPerCta * 32)
ball_query_kernel(int n, int m, float radius, int nsample, const float* __restrict__ new_xyz,
                  const float* __restrict__ xyz, int* __restrict__ idx) {
    pdl_enter();  // programmatic dependent launch (pn2_common.cuh): first statement, before any memory access
    __shared__ __align__(16) float s_pts[2][kTile * 3];
    __shared__ __align__(8) uint64_t s_bar[2];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y;
    const int centre = blockIdx.x * kWarpsPerCta + warp;
    const float* pts = xyz + (size_t)b * n * 3;
    const int ntiles = (n + kTile - 1) / kTile;
    // bulk copies need 16-byte aligned sources; N % 4 != 0 breaks that for odd clouds
    const bool use_bulk = (reinterpret_cast<uintptr_t>(pts) & 15) == 0;

    // radius*radius in fp32, as ball_query_gpu.cu:23
    const float radius2 = __fmul_rn(radius, radius);
    const bool active = centre < m;
    float cx = 0.f, cy = 0.f, cz = 0.f;
    if (active) {
        const float* c = new_xyz + ((size_t)b * m + centre) * 3;
        cx = c[0]; cy = c[1]; cz = c[2];
    }
    int* out = idx + ((size_t)b * m + (active ? centre : 0)) * nsample;

    auto tile_count = [&](int t) { return min(kTile, n - t * kTile); };
    auto issue = [&](int t) {  // one thread: hand tile t to the TMA engine
        const int cnt = tile_count(t);
        const uint32_t bytes = ((uint32_t)cnt * 12u) & ~15u;
        if (bytes) {
            mbar_arrive_expect_tx(&s_bar[t & 1], bytes);
            bulk_g2s(s_pts[t & 1], pts + (size_t)t * kTile * 3, bytes, &s_bar[t & 1]);
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

    int found = 0;      // hits so far (warp-uniform)
    int first = 0;      // first hit index, for padding
    bool done = !active;

    int t = 0;
    for (; t < ntiles; ++t) {
        const int cnt = tile_count(t);
        float* sp = s_pts[t & 1];
        if (use_bulk) {
            // the (< 16 byte) tail of the last tile is not a legal bulk size: plain loads
            const int bulk_floats = (int)((((uint32_t)cnt * 12u) & ~15u) >> 2);
            const int tail = cnt * 3 - bulk_floats;
            if (tid < tail) sp[bulk_floats + tid] = pts[(size_t)t * kTile * 3 + bulk_floats + tid];
            mbar_wait(&s_bar[t & 1], (t >> 1) & 1);
        } else {
            for (int i = tid; i < cnt * 3; i += blockDim.x) sp[i] = pts[(size_t)t * kTile * 3 + i];
        }
        __syncthreads();

        if (!done) {
            const int base = t * kTile;
            for (int i = 0; i < cnt; i += 32) {
                const int j = i + lane;
                bool hit = false;
                if (j < cnt) {
                    const float d2 = sqdist(cx, cy, cz, sp[j * 3 + 0], sp[j * 3 + 1], sp[j * 3 + 2]);
                    hit = d2 < radius2;  // strict; NaN is a miss (ball_query_gpu.cu:34)
                }
                const unsigned mask = __ballot_sync(kFull, hit);
                if (mask) {
                    if (found == 0) first = base + i + __ffs(mask) - 1;
                    const int slot = found + __popc(mask & ((1u << lane) - 1u));
                    if (hit && slot < nsample) out[slot] = base + j;
                    found += __popc(mask);
                    if (found >= nsample) { done = true; break; }
                }
            }
        }
        // every warp is past this buffer before it is refilled; leave early when all are done
        const int all_done = __syncthreads_and(done);
        if (all_done) break;
        if (use_bulk && tid == 0 && t + 2 < ntiles) issue(t + 2);
    }
    // left early with tile t+1 still in flight: the CTA must outlive its bulk copy
    if (use_bulk && tid == 0 && t + 1 < ntiles) mbar_wait(&s_bar[(t + 1) & 1], ((t + 1) >> 1) & 1);

    // pad the remaining slots with the first hit (ball_query_gpu.cu:35-39); an empty ball keeps index 0 in every slot --
    // the reference leaves its zero-initialised output untouched there (pointnet2_utils.py:184 of the reference); written
    // here so that the caller need not zero the buffer first
    if (active && found < nsample)
        for (int s = found + lane; s < nsample; s += 32) out[s] = found > 0 ? first : 0;
}

}  // namespace
}  // namespace pn2

extern "C" int pn2_ball_query(int b, int n, int m, float radius, int nsample, const float* new_xyz,
                              const float* xyz, int* idx, pn2_stream_t stream) {
    using namespace pn2;
    if (b < 0 || n < 0 || m < 0 || nsample < 0) return fail_arg("pn2_ball_query", "negative size");
    if (b == 0 || m == 0 || n == 0 || nsample == 0) return 0;
    if (b > 65535) return fail_arg("pn2_ball_query", "b > 65535");
    if (!new_xyz || !xyz || !idx) return fail_arg("pn2_ball_query", "null pointer");
    dim3 grid((m + kWarpsPerCta - 1) / kWarpsPerCta, b);
    launch_k(ball_query_kernel, dim3(grid), dim3(kWarpsPerCta * 32), 0, (cudaStream_t)stream, n, m, radius, nsample, new_xyz, xyz, idx);
    PN2_CHECK_LAUNCH("ball_query_kernel");
    return 0;
}
