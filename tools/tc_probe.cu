// tcgen05 bring-up probe (development tool, not part of the product library).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/tc_probe tools/tc_probe.cu && ./tools/tc_probe
// Checks, on real hardware, the conventions the cost-volume kernel relies on:
//   P1  kind::f16 MMA, A and B from shared memory, K-major no-swizzle core-matrix layout, LBO/SBO meaning
//   P2  A operand from TMEM written with tcgen05.st 32x32b (row = lane, fp16 pairs packed per column)
//   P3  D of one MMA written over the (dead) A columns of an earlier MMA in the same issue stream
//   P4  MMA issue throughput (cycles per M128 N256 K16 instruction)
// Every mbarrier wait is bounded, so a wrong guess produces an error line, not a hang.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ bool mbar_wait_bounded(uint64_t *bar, uint32_t parity, int *err) {
    for (int it = 0; it < (1 << 22); ++it) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (ok) return true;
    }
    *err = 1;
    return false;
}
__device__ __forceinline__ void tmem_alloc(uint32_t *slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
// 32 lanes x 32 consecutive columns: thread t of warp w <-> lane 32*(w%4)+t
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t *r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t *r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor, K-major, no swizzle (mma_sm100_desc.hpp: SmemDescriptor)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;  // version = 1 (Blackwell)
    return d;                // base_offset 0, lbo_mode 0, layout_type 0 = SWIZZLE_NONE
}
// instruction descriptor: F16 x F16 -> F32, both K-major
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

constexpr int M = 128, N = 256, K = 64;   // K = 4 MMA k-steps of 16
// canonical no-swizzle K-major: element (r,k) at ((k/8)*(R/8) + r/8)*128 + (r%8)*16 + (k%8)*2 bytes
__host__ __device__ inline int core_off(int r, int k, int R) { return ((k / 8) * (R / 8) + r / 8) * 64 + (r % 8) * 8 + (k % 8); }

// mode 0: A,B from smem, LBO = K-stride, SBO = MN-stride      (expected correct)
// mode 1: A,B from smem, LBO/SBO swapped
// mode 2: A from TMEM (tcgen05.st), B from smem (mode-0 convention)
// mode 3: like 2, then a second MMA chain writes its D over the A columns (K=16 only, A2 from smem)
__global__ void __launch_bounds__(192) probe_kernel(int mode, const __half *A, const __half *B, float *D, float *D2, int *err) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __half *sA = reinterpret_cast<__half *>(smem);                  // 128*64*2 = 16 KB
    __half *sB = reinterpret_cast<__half *>(smem + 16384);          // 256*64*2 = 32 KB
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < M * K; i += blockDim.x) sA[core_off(i / K, i % K, M)] = A[i];
    for (int i = threadIdx.x; i < N * K; i += blockDim.x) sB[core_off(i / K, i % K, N)] = B[i];
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> async proxy (UMMA)
    if (warp == 4) tmem_alloc(&tmem_slot, 512);
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tm = tmem_slot;
    const uint32_t tD = tm, tA = tm + 256;
    if (mode >= 2 && warp < 4) {
        // A -> TMEM: lane = row, column c holds (k=2c, k=2c+1)
        const int row = warp * 32 + lane;
        for (int c0 = 0; c0 < K / 2; c0 += 8) {
            uint32_t r[8];
            for (int j = 0; j < 8; ++j) {
                __half2 h = __halves2half2(A[row * K + 2 * (c0 + j)], A[row * K + 2 * (c0 + j) + 1]);
                r[j] = *reinterpret_cast<uint32_t *>(&h);
            }
            tmem_st8(tA + ((uint32_t)(warp * 32) << 16) + c0, r);
        }
        tmem_st_wait();
        fence_before();
    }
    __syncthreads();
    if (warp == 5) {
        fence_after();
        if (lane == 0) {
            const uint32_t idesc = make_idesc(M, N);
            const uint32_t kstrideA = (M / 8) * 128, kstrideB = (N / 8) * 128;
            for (int k = 0; k < K / 16; ++k) {
                uint64_t da, db;
                if (mode == 1) {
                    da = make_desc(smem_u32(sA) + k * 2 * kstrideA, 128, kstrideA);
                    db = make_desc(smem_u32(sB) + k * 2 * kstrideB, 128, kstrideB);
                } else {
                    da = make_desc(smem_u32(sA) + k * 2 * kstrideA, kstrideA, 128);
                    db = make_desc(smem_u32(sB) + k * 2 * kstrideB, kstrideB, 128);
                }
                if (mode >= 2) mma_ts(tD, tA + k * 8, db, idesc, k > 0);
                else mma_ss(tD, da, db, idesc, k > 0);
            }
            if (mode == 3) {
                // second product D2 = A[:, 0:16] * B[:, 0:16]^T written over the A columns (cols 256..511)
                uint64_t da = make_desc(smem_u32(sA), kstrideA, 128);
                uint64_t db = make_desc(smem_u32(sB), kstrideB, 128);
                mma_ss(tA, da, db, idesc, 0);
            }
            umma_commit(&bar);
        }
        __syncwarp();
    }
    if (warp < 4) {
        if (mbar_wait_bounded(&bar, 0, err)) {
            fence_after();
            const int row = warp * 32 + lane;
            for (int c0 = 0; c0 < N; c0 += 32) {
                uint32_t r[32];
                tmem_ld32(tD + ((uint32_t)(warp * 32) << 16) + c0, r);
                for (int j = 0; j < 32; ++j) D[row * N + c0 + j] = __uint_as_float(r[j]);
            }
            if (mode == 3)
                for (int c0 = 0; c0 < N; c0 += 32) {
                    uint32_t r[32];
                    tmem_ld32(tA + ((uint32_t)(warp * 32) << 16) + c0, r);
                    for (int j = 0; j < 32; ++j) D2[row * N + c0 + j] = __uint_as_float(r[j]);
                }
        }
        fence_before();
    }
    __syncthreads();
    if (warp == 4) tmem_dealloc(tm, 512);
}

// P4: issue `iters` x (M128 N256 K16) MMAs back to back, report cycles per instruction
__global__ void __launch_bounds__(192) throughput_kernel(int iters, int from_tmem, long long *cycles, int *err) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 49152 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 4) tmem_alloc(&tmem_slot, 512);
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tm = tmem_slot;
    if (warp == 5 && lane == 0) {
        const uint32_t idesc = make_idesc(M, N);
        const uint32_t kA = (M / 8) * 128, kB = (N / 8) * 128;
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const int k = i & 3;
            uint64_t da = make_desc(smem_u32(smem) + k * 2 * kA, kA, 128);
            uint64_t db = make_desc(smem_u32(smem + 16384) + k * 2 * kB, kB, 128);
            if (from_tmem) mma_ts(tm, tm + 256 + k * 8, db, idesc, 1);
            else mma_ss(tm, da, db, idesc, 1);
        }
        umma_commit(&bar);
        mbar_wait_bounded(&bar, 0, err);
        *cycles = clock64() - t0;
    }
    __syncthreads();
    if (warp == 4) tmem_dealloc(tm, 512);
}

int main() {
    __half *hA = (__half *)malloc(M * K * 2), *hB = (__half *)malloc(N * K * 2);
    float *ref = (float *)malloc(M * N * 4), *ref2 = (float *)malloc(M * N * 4), *out = (float *)malloc(M * N * 4);
    srand(1);
    for (int i = 0; i < M * K; ++i) hA[i] = __float2half((rand() % 2001 - 1000) / 500.0f);
    for (int i = 0; i < N * K; ++i) hB[i] = __float2half((rand() % 2001 - 1000) / 500.0f);
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            double s = 0, s2 = 0;
            for (int k = 0; k < K; ++k) {
                double p = (double)__half2float(hA[m * K + k]) * (double)__half2float(hB[n * K + k]);
                s += p;
                if (k < 16) s2 += p;
            }
            ref[m * N + n] = (float)s;
            ref2[m * N + n] = (float)s2;
        }
    __half *dA, *dB;
    float *dD, *dD2;
    int *dErr;
    long long *dCyc;
    CK(cudaMalloc(&dA, M * K * 2)); CK(cudaMalloc(&dB, N * K * 2)); CK(cudaMalloc(&dD, M * N * 4)); CK(cudaMalloc(&dD2, M * N * 4));
    CK(cudaMalloc(&dErr, 4)); CK(cudaMalloc(&dCyc, 8));
    CK(cudaMemcpy(dA, hA, M * K * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, hB, N * K * 2, cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    CK(cudaFuncSetAttribute(throughput_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    for (int mode = 0; mode < 4; ++mode) {
        if (mode == 1) continue;  // swapped LBO/SBO: faults (illegal address) on hardware -- convention settled by mode 0
        CK(cudaMemset(dD, 0xff, M * N * 4)); CK(cudaMemset(dD2, 0xff, M * N * 4)); CK(cudaMemset(dErr, 0, 4));
        probe_kernel<<<1, 192, 49152>>>(mode, dA, dB, dD, dD2, dErr);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d: CUDA error %s\n", mode, cudaGetErrorString(e)); return 1; }
        int err; CK(cudaMemcpy(&err, dErr, 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(out, dD, M * N * 4, cudaMemcpyDeviceToHost));
        double mx = 0; int bad = 0;
        for (int i = 0; i < M * N; ++i) { double d = fabs((double)out[i] - ref[i]); if (!(d <= 1e-2)) ++bad; if (d > mx) mx = d; }
        printf("mode %d: timeout=%d  max|err|=%.3e  mismatches=%d/%d  (D[0]=%f ref=%f, D[129]=%f ref=%f)\n", mode, err, mx, bad, M * N,
               out[0], ref[0], out[129], ref[129]);
        if (mode == 3) {
            CK(cudaMemcpy(out, dD2, M * N * 4, cudaMemcpyDeviceToHost));
            mx = 0; bad = 0;
            for (int i = 0; i < M * N; ++i) { double d = fabs((double)out[i] - ref2[i]); if (!(d <= 1e-2)) ++bad; if (d > mx) mx = d; }
            printf("mode 3 (D2 over A columns): max|err|=%.3e mismatches=%d\n", mx, bad);
        }
    }
    for (int from_tmem = 0; from_tmem < 2; ++from_tmem)
        for (int iters = 64; iters <= 1024; iters *= 4) {
            CK(cudaMemset(dErr, 0, 4));
            throughput_kernel<<<1, 192, 49152>>>(iters, from_tmem, dCyc, dErr);
            CK(cudaDeviceSynchronize());
            long long c; int err;
            CK(cudaMemcpy(&c, dCyc, 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&err, dErr, 4, cudaMemcpyDeviceToHost));
            printf("throughput A-from-%s iters=%d: %lld cycles = %.1f cyc/MMA (M128 N256 K16 f16) timeout=%d\n",
                   from_tmem ? "tmem" : "smem", iters, c, (double)c / iters, err);
        }
    return 0;
}
