// errors.cu -- status / error-string plumbing of the C ABI (include/pn2b200.h).
#include "pn2_common.cuh"

#include <cstdio>

namespace pn2 {
namespace {
thread_local char g_last_error[512] = "";
}

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

extern "C" int pn2_version(void) { return 100; /* 0.1.0 */ }
extern "C" const char* pn2_last_error(void) { return pn2::g_last_error; }
