// pn2_common.cuh -- shared device helpers for the sm_100a PointNet++ kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pn2b200.h"

namespace pn2 {

// "once per device" guard for per-function attributes (cudaFuncSetAttribute applies to the CURRENT device's copy of the
// function: a process that drives several GPUs must set it on each) and the SM count of the current device.
struct DeviceOnce {
    bool done[64] = {};
    bool first() {
        int d = 0;
        if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return true;
        if (done[d]) return false;
        done[d] = true;
        return true;
    }
};
inline int sm_count() {
    static int cached[64] = {};
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return 148;
    if (cached[d] == 0) {
        int v = 148;
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, d);
        cached[d] = v;
    }
    return cached[d];
}


constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

// Records the failing call for pn2_last_error() and returns the status code the
// C ABI hands back (the reference prints and exit(-1)s instead, e.g.
// ball_query_gpu.cu:62-66).
int fail(cudaError_t err, const char* where);
int fail_arg(const char* where, const char* what);
// Counts kernel launches issued by this library (pn2_launch_count(); bench.py's gpu_launches).
void count_launch();

#define PN2_CHECK_LAUNCH(where)                                    \
    do {                                                           \
        cudaError_t e__ = cudaGetLastError();                      \
        if (e__ != cudaSuccess) return ::pn2::fail(e__, where);    \
        ::pn2::count_launch();                                     \
    } while (0)

// ---- programmatic dependent launch (PDL) ------------------------------------------------------------------------
// A training step is ~240 dependent kernels, most of them 5..30 us long: the gap between "last CTA of kernel N retires" and
// "first CTA of kernel N+1 runs its first instruction" (grid completion, launch, CTA scheduling) is a measurable part of
// the step.  Kernels launched through launch_k() carry cudaLaunchAttributeProgrammaticStreamSerialization: their CTAs may
// be scheduled as soon as every CTA of the preceding kernel in the stream has STARTED (all of ours call pdl_trigger() in
// their first instructions) and a slot is free, and they block in pdl_wait() -- before ANY global memory access (only on-chip
// set-up such as barrier initialisation and TMEM allocation may precede it) -- their first statement otherwise, before ANY global
// memory access -- until the preceding kernel has completed and its writes are visible.  Semantics are those of plain
// stream order (a kernel that waits cannot complete before its predecessor, so the guarantee is transitive); only the
// launch latency is overlapped.  RULE: a kernel may be launched with launch_k() only if pdl_wait() is its first statement.
// Works eagerly and under stream capture (programmatic graph edges).  PN2_PDL=0 turns the attribute off.
bool pdl_enabled();
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() {
    pdl_wait();
    pdl_trigger();
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

#define PN2_CHECK(call, where)                                     \
    do {                                                           \
        cudaError_t e__ = (call);                                  \
        if (e__ != cudaSuccess) return ::pn2::fail(e__, where);    \
    } while (0)

// Squared distance with the exact rounding sequence of the reference kernels as
// nvcc 12.9 -O2 contracts them for sm_100 (SASS: FADD,FADD,FMUL,FADD,FFMA,FFMA):
//   d = fma(dz,dz, fma(dx,dx, rn(dy*dy)))
// (sampling_gpu.cu:133, ball_query_gpu.cu:33, interpolate_gpu.cu:40,108).
// Written with intrinsics so no compiler flag or version can re-associate it.
__device__ __forceinline__ float sqdist(float ax, float ay, float az, float bx, float by, float bz) {
    const float dx = __fsub_rn(ax, bx);
    const float dy = __fsub_rn(ay, by);
    const float dz = __fsub_rn(az, bz);
    return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// ---- mbarrier + 1-D bulk async copy (TMA engine, SASS UBLKCP) -------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// ... with a suspend-time hint: the waiting thread may be parked by the hardware for up to `ns` nanoseconds per probe and
// is resumed early when the phase completes -- for warps that wait most of the time (an MMA issuer, loader warps), whose
// polling would otherwise take issue slots from the working warps of their scheduler
__device__ __forceinline__ void mbar_wait_parked(uint64_t* bar, uint32_t parity, uint32_t ns = 400) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAITP_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra DONEP_%=;\n"
        "bra WAITP_%=;\n"
        "DONEP_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity), "r"(ns)
        : "memory");
}
// ... for waits that last a long time (an epilogue waiting for the whole main loop): back off between polls so the
// spinning warps do not take issue slots from the working ones
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity, unsigned ns = 256) {
    for (;;) {
        uint32_t ok;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (ok) return;
        __nanosleep(ns);
    }
}
// global -> shared bulk copy; dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

}  // namespace pn2
