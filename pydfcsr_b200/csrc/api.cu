// api.cu — ABI version and thread-local error reporting for libdfcsr_b200.
#include <stdarg.h>
#include <atomic>
#include <string.h>
#include "common.cuh"

namespace dfcsr {

static thread_local char g_error[512] = "";
static std::atomic<long long> g_launches{0};

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* where) {
    set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), where);
    return DFCSR_ERR_CUDA;
}

}  // namespace dfcsr

extern "C" int dfcsr_abi_version(void) { return DFCSR_ABI_VERSION; }
extern "C" const char* dfcsr_last_error(void) { return dfcsr::g_error; }
extern "C" int64_t dfcsr_launch_count(void) { return (int64_t)dfcsr::g_launches.load(std::memory_order_relaxed); }
