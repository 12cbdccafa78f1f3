// errors.cu -- status / error-string plumbing of the C ABI (include/pn2b200.h).
#include "pn2_common.cuh"

#include <atomic>
#include <cstdio>
#include <cstdlib>

namespace pn2 {
namespace {
thread_local char g_last_error[512] = "";
std::atomic<long long> g_launches{0};
}

bool pdl_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("PN2_PDL");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v == 1;
}

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int fail(cudaError_t err, const char* where) {
    snprintf(g_last_error, sizeof(g_last_error), "%s: %s (%s)", where, cudaGetErrorString(err),
             cudaGetErrorName(err));
    return (int)err;
}

int fail_arg(const char* where, const char* what) {
    snprintf(g_last_error, sizeof(g_last_error), "%s: invalid argument: %s", where, what);
    return PN2_EINVAL;
}
}  // namespace pn2

extern "C" int pn2_version(void) { return 201; /* 0.2.1: pn2_fp_rows_bwd takes skip_rows_major; pn2_pool_bwd takes extra_rows for k > 1; ball_query writes empty balls */ }
extern "C" long long pn2_launch_count(void) { return pn2::g_launches.load(std::memory_order_relaxed); }
extern "C" const char* pn2_last_error(void) { return pn2::g_last_error; }
