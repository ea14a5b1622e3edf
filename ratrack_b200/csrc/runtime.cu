// Error plumbing + version for the C-ABI library.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

static thread_local char g_err[512] = "";

void rt_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int rt_check_launch(const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) return RT_OK;
    rt_set_error("%s: CUDA launch failed: %s", what, cudaGetErrorString(e));
    return (int)e;
}

RT_API const char *rt_last_error(void) { return g_err; }
RT_API int rt_abi_version(void) { return 1; }
