// Library plumbing of libfastb: version, thread-local error text, launch counter.
#include "fastb_common.cuh"

#include <string.h>

namespace fastb {

static thread_local char g_err[512] = "";
static thread_local long long g_launches = 0;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches += n; }

int fail_cuda(cudaError_t e, const char* what) {
    set_error("%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
    return FASTB_ERR_CUDA;
}

int check_launch(const char* what) {
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail_cuda(e, what);
    count_launch(1);
    return FASTB_OK;
}

}  // namespace fastb

extern "C" int fastb_version(void) { return FASTB_VERSION; }
extern "C" const char* fastb_last_error(void) { return fastb::g_err; }
extern "C" int fastb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}
extern "C" int64_t fastb_launch_count(void) { return fastb::g_launches; }
extern "C" void fastb_reset_launch_count(void) { fastb::g_launches = 0; }
