// Error plumbing + version for the C-ABI library.
#include <pthread.h>
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

// ---- stream-ordered scratch ---------------------------------------------------------------------------------------
// The C ABI keeps the reference's signatures (no workspace argument), so kernels that need scratch take it from a
// stream-ordered pool.  The DEFAULT pool of a device releases everything back to the driver at the next synchronisation
// (release threshold 0): measured on B200 through tools/bench_ops.py, a cudaMallocAsync / cudaFreeAsync pair in every call
// cost 0.6-1.7 ms of host time -- 5x the kernels it served.  One explicit pool per device with an unlimited release threshold
// keeps the memory cached: after the first call an allocation is a free-list lookup.
static cudaMemPool_t g_pool[64];
static unsigned long long g_pool_ready = 0;   // bit per device

static cudaMemPool_t rt_scratch_pool(int dev) {
    if (dev < 0 || dev >= 64) return nullptr;
    if (!((__atomic_load_n(&g_pool_ready, __ATOMIC_ACQUIRE) >> dev) & 1ull)) {
        static pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
        pthread_mutex_lock(&mu);
        if (!((g_pool_ready >> dev) & 1ull)) {
            cudaMemPoolProps props;
            memset(&props, 0, sizeof(props));
            props.allocType = cudaMemAllocationTypePinned;
            props.handleTypes = cudaMemHandleTypeNone;
            props.location.type = cudaMemLocationTypeDevice;
            props.location.id = dev;
            cudaMemPool_t pool = nullptr;
            if (cudaMemPoolCreate(&pool, &props) == cudaSuccess) {
                unsigned long long keep = ~0ull;
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
                g_pool[dev] = pool;
            } else {
                (void)cudaGetLastError();
                g_pool[dev] = nullptr;   // fall back to the device's default pool
            }
            __atomic_fetch_or(&g_pool_ready, 1ull << dev, __ATOMIC_RELEASE);
        }
        pthread_mutex_unlock(&mu);
    }
    return g_pool[dev];
}

int rt_scratch_alloc(void **p, size_t bytes, cudaStream_t st, const char *what) {
    cudaMemPool_t pool = rt_scratch_pool(rt_current_device());
    const cudaError_t e = pool ? cudaMallocFromPoolAsync(p, bytes, pool, st) : cudaMallocAsync(p, bytes, st);
    if (e != cudaSuccess) {
        rt_set_error("%s: scratch allocation of %zu bytes failed: %s", what, bytes, cudaGetErrorString(e));
        *p = nullptr;
        return (int)e;
    }
    return RT_OK;
}
void rt_scratch_free(void *p, cudaStream_t st) {
    if (p) cudaFreeAsync(p, st);
}

RT_API const char *rt_last_error(void) { return g_err; }
RT_API int rt_abi_version(void) { return 1; }
