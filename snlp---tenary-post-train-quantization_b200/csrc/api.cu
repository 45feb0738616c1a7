// Library-wide state: error string, ABI version, launch counter.
#include "common.cuh"

#include <stdarg.h>
#include <string.h>

namespace tq {

static thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sm_count() {
    static int cached = 0;
    if (cached == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            cached = n;
        else
            cached = 148;
    }
    return cached;
}

}  // namespace tq

extern "C" int tq_abi_version(void) { return TQ100_ABI_VERSION; }
extern "C" const char* tq_last_error_string(void) { return tq::g_err; }
extern "C" long long tq_launch_count(void) { return tq::g_launches.load(); }
