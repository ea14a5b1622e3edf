// Shared device/host helpers for the ratrack_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define RT_API extern "C" __attribute__((visibility("default")))

// ---- error plumbing: every C-ABI entry returns 0 or a negative/ CUDA error code; text via rt_last_error()
enum {
    RT_OK = 0,
    RT_ERR_INVALID = -1,      // bad argument (null pointer, negative size, k > 200 ...)
    RT_ERR_UNSUPPORTED = -2,  // shape outside what the kernels were built for
};

void rt_set_error(const char *fmt, ...);
int rt_check_launch(const char *what);  // cudaGetLastError() -> code (+ message)
// stream-ordered scratch from a per-device pool that keeps its memory cached (runtime.cu)
int rt_scratch_alloc(void **p, size_t bytes, cudaStream_t st, const char *what);
void rt_scratch_free(void *p, cudaStream_t st);

#define RT_REQUIRE(cond, ...)                \
    do {                                     \
        if (!(cond)) {                       \
            rt_set_error(__VA_ARGS__);       \
            return RT_ERR_INVALID;           \
        }                                    \
    } while (0)

static inline int rt_divup(long long a, long long b) { return (int)((a + b - 1) / b); }

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a PER-DEVICE property of a kernel: a process that drives several
// GPUs (nn.DataParallel threads, one engine per device) must opt in on each of them.  One bit per device ordinal,
// thread-safe; a device is marked only after the caller's attribute calls succeeded.
struct RtPerDevice {
    unsigned long long mask[4] = {0, 0, 0, 0};   // device ordinals 0..255
    bool done(int dev) const {
        return dev >= 0 && dev < 256 && (__atomic_load_n(&mask[dev >> 6], __ATOMIC_ACQUIRE) >> (dev & 63)) & 1ull;
    }
    void mark(int dev) {
        if (dev >= 0 && dev < 256) __atomic_fetch_or(&mask[dev >> 6], 1ull << (dev & 63), __ATOMIC_RELEASE);
    }
};
static inline int rt_current_device() {
    int dev = 0;
    cudaGetDevice(&dev);
    return dev;
}

// largest power of two <= n, capped at 1024: the CTA size the reference picks for FPS
// (reference: src/lib/src/cuda_utils.h:10-14); it fixes the arg-max tie-break order.
static inline int rt_ref_block_size(int n) {
    int p = 1;
    while (p * 2 <= n && p < 1024) p *= 2;
    return p;
}

#ifdef __CUDACC__
// ---- squared distance in the reference's sm_100 evaluation order:
//      t = dy*dy (rounded); t = fma(dx,dx,t); d = fma(dz,dz,t)
__device__ __forceinline__ float rt_sqdist(float ax, float ay, float az, float bx, float by, float bz) {
    const float dx = ax - bx, dy = ay - by, dz = az - bz;
    float t = __fmul_rn(dy, dy);
    t = __fmaf_rn(dx, dx, t);
    return __fmaf_rn(dz, dz, t);
}

__device__ __forceinline__ uint32_t rt_smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

// ---- mbarrier + 1-D bulk async copy (TMA engine, SASS UBLKCP) -------------------------------
__device__ __forceinline__ void rt_mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(rt_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void rt_fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void rt_fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void rt_mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(rt_smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void rt_mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(rt_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void rt_mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(rt_smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy; dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void rt_bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     rt_smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(rt_smem_u32(bar))
                 : "memory");
}

// Stage `nfloats` contiguous floats into shared memory with the bulk-copy engine when the
// source is 16-byte aligned (always true for cloud bases when N % 4 == 0), else with plain
// loads.  Whole CTA must call; returns after the data is visible to every thread.
// `bar` must have been initialised with count 1 and is used once per call phase `parity`.
__device__ __forceinline__ void rt_stage_floats(float *s_dst, const float *g_src, int nfloats, uint64_t *bar,
                                                uint32_t parity) {
    const int bulk = ((reinterpret_cast<uintptr_t>(g_src) & 15) == 0) ? (nfloats & ~3) : 0;
    if (threadIdx.x == 0) {
        if (bulk > 0) {
            rt_mbar_expect_tx(bar, (uint32_t)bulk * 4u);
            // split into <= 64 KB pieces (engine-friendly, keeps expect_tx within range)
            for (int off = 0; off < bulk; off += 16384) {
                const int n = min(16384, bulk - off);
                rt_bulk_g2s(s_dst + off, g_src + off, (uint32_t)n * 4u, bar);
            }
        } else {
            rt_mbar_arrive(bar);
        }
    }
    for (int i = bulk + threadIdx.x; i < nfloats; i += blockDim.x) s_dst[i] = g_src[i];
    rt_mbar_wait(bar, parity);
    __syncthreads();
}

// 256-bit read-only global load (SASS LDG.E.ENL2.256): one instruction per 32-byte sector.  For thread-per-row access
// patterns (every lane in a different row) the L1 data pipe spends a wavefront per lane per instruction whatever its
// width, so twice the bytes per instruction halve the wavefronts.  `p` must be 32-byte aligned.
__device__ __forceinline__ void rt_ldg256(const float *p, float4 &lo, float4 &hi) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(lo.x), "=f"(lo.y), "=f"(lo.z), "=f"(lo.w), "=f"(hi.x), "=f"(hi.y), "=f"(hi.z), "=f"(hi.w)
                 : "l"(p));
}

// 256-bit global store (SASS STG.E.ENL2.256), the store-side counterpart of rt_ldg256; `p` must be 32-byte aligned
__device__ __forceinline__ void rt_stg256(float *p, float4 lo, float4 hi) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(lo.x), "f"(lo.y), "f"(lo.z), "f"(lo.w), "f"(hi.x),
                 "f"(hi.y), "f"(hi.z), "f"(hi.w)
                 : "memory");
}

// packed fp32 pairs (SASS FFMA2 / FMUL2 / FADD2): every half is an ordinary round-to-nearest IEEE operation, so results are
// bit-identical to the scalar forms; one issue slot instead of two
__device__ __forceinline__ float2 rt_ffma2(float2 a, float2 b, float2 c) {
    unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a), rb = *reinterpret_cast<unsigned long long *>(&b),
                       rc = *reinterpret_cast<unsigned long long *>(&c), rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2 *>(&rd);
}
__device__ __forceinline__ float2 rt_fmul2(float2 a, float2 b) {
    unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a), rb = *reinterpret_cast<unsigned long long *>(&b), rd;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    return *reinterpret_cast<float2 *>(&rd);
}
__device__ __forceinline__ float2 rt_fadd2(float2 a, float2 b) {
    unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a), rb = *reinterpret_cast<unsigned long long *>(&b), rd;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    return *reinterpret_cast<float2 *>(&rd);
}

__device__ __forceinline__ uint32_t rt_redux_max_u32(uint32_t v) {
    uint32_t r;
    asm volatile("redux.sync.max.u32 %0, %1, 0xffffffff;" : "=r"(r) : "r"(v));
    return r;
}
__device__ __forceinline__ uint32_t rt_redux_min_u32(uint32_t v) {
    uint32_t r;
    asm volatile("redux.sync.min.u32 %0, %1, 0xffffffff;" : "=r"(r) : "r"(v));
    return r;
}
// monotone map float -> uint32 (total order, -0 < +0, NaN sorted high/low by sign bit; callers never pass NaN)
__device__ __forceinline__ uint32_t rt_float_ordered(float f) {
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
#endif
