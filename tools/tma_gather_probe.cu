// Hardware probe for the TMA gather4 path (sm_100a): cp.async.bulk.tensor.2d ... tile::gather4 of four arbitrary rows of a
// row-major fp32 matrix into shared memory, with and without the 128-byte swizzle.  Prints where every source element landed.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tools/tma_gather_probe tools/tma_gather_probe.cu -lcuda && tools/tma_gather_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__global__ void probe(const __grid_constant__ CUtensorMap tmap, int col, int r0, int r1, int r2, int r3, int bytes, float *out, int *status) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar), dst = (uint32_t)__cvta_generic_to_shared(smem);
    for (int i = threadIdx.x; i < 2048 / 4; i += blockDim.x) reinterpret_cast<float *>(smem)[i] = -1.0f;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                     ::"r"(dst), "l"(&tmap), "r"(bar_a), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
    }
    // bounded wait
    int ok = 0;
    for (int it = 0; it < (1 << 20) && !ok; ++it) {
        uint32_t p;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(p) : "r"(bar_a) : "memory");
        ok = p;
    }
    if (threadIdx.x == 0) *status = ok;
    __syncthreads();
    for (int i = threadIdx.x; i < 2048 / 4; i += blockDim.x) out[i] = reinterpret_cast<float *>(smem)[i];
}

int main() {
    const int R = 1024, C = 256;
    float *h = (float *)malloc(sizeof(float) * R * C);
    for (int r = 0; r < R; ++r) for (int c = 0; c < C; ++c) h[r * C + c] = r * 1000.0f + c;   // value encodes (row, col)
    float *d, *out; int *status;
    CK(cudaMalloc(&d, sizeof(float) * R * C)); CK(cudaMalloc(&out, 2048)); CK(cudaMalloc(&status, 4));
    CK(cudaMemcpy(d, h, sizeof(float) * R * C, cudaMemcpyHostToDevice));
    for (int variant = 0; variant < 4; variant += 2) {   // (box rows 4 is an illegal instruction on sm_100a: the four rows ARE the box height)
        const int box_rows = (variant & 1) ? 4 : 1;
        const CUtensorMapSwizzle sw = (variant & 2) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE;
        CUtensorMap tmap;
        cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)R}, strides[1] = {(cuuint64_t)C * 4};
        cuuint32_t box[2] = {32, (cuuint32_t)box_rows}, estr[2] = {1, 1};
        CUresult cr = cuTensorMapEncodeTiled(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                                             CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("== variant %d: box rows %d, swizzle %s: encode rc=%d\n", variant, box_rows, (variant & 2) ? "128B" : "none", (int)cr);
        if (cr != CUDA_SUCCESS) continue;
        CK(cudaMemset(status, 0, 4));
        probe<<<1, 128, 2048>>>(tmap, 64, 5, 17, 900, 3, 4 * 32 * 4, out, status);   // destination = dynamic shared memory base (1024-byte aligned)
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("   kernel error: %s\n", cudaGetErrorString(e)); cudaGetLastError(); continue; }
        int st; float ho[512];
        CK(cudaMemcpy(&st, status, 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(ho, out, 2048, cudaMemcpyDeviceToHost));
        printf("   barrier completed: %d\n", st);
        for (int row = 0; row < 4; ++row) {   // 128-byte lines of the destination
            printf("   smem line %d:", row);
            for (int ch = 0; ch < 8; ++ch) printf(" %8.0f", ho[row * 32 + ch * 4]);   // first float of every 16-byte chunk
            printf("\n");
        }
    }
    return 0;
}
