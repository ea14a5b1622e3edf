// Index-driven copy kernels for sm_100a: gather_points, group_points, three_interpolate and
// their gradients.
//
// Replace gather_points_kernel_fast / gather_points_grad_kernel_fast
// (reference: src/lib/src/sampling_gpu.cu:8-24, 46-63), group_points_kernel_fast /
// group_points_grad_kernel_fast (src/lib/src/group_points_gpu.cu:47-66, 8-25) and
// three_interpolate_kernel_fast / three_interpolate_grad_kernel_fast
// (src/lib/src/interpolate_gpu.cu:149-169, 192-214).
//
// The reference launches one thread per OUTPUT ELEMENT PER CHANNEL, so every index (and
// weight) is re-read C times.  Here a thread owns 4 consecutive output positions, loads its
// indices / weights once as 128-bit vectors, then streams over a slab of channels writing
// 128-bit coalesced stores; the gathers hit a (B,C,N) row that stays in L1/L2.
// All base offsets are formed in 64 bits (the reference's int32 offsets overflow at large B).
#include "common.cuh"

bool rt_segsum_supported(int n_dst, long long e_total);   // segsum.cu
int rt_launch_segmented_scatter(int b, int c, int n_dst, long long e_total, int src_div, const float *grad_out, const int *idx,
                                const float *weight, float *grad_points, cudaStream_t st, const char *what);

namespace {

constexpr int G_THREADS = 256;
constexpr int G_CH_SLAB = 16;  // channels per CTA (grid.y = ceil(C / slab))

// out[b,c,e] = points[b,c,idx[b,e]]   for e in [0, E): gather_points (E = npoints) and
// group_points (E = npoints*nsample) are the same operation.
__global__ void __launch_bounds__(G_THREADS) gather_rows_kernel(int c, int n, long long e_total, int slab,
                                                                const float *__restrict__ points,
                                                                const int *__restrict__ idx,
                                                                float *__restrict__ out) {
    const int b = blockIdx.z;
    const int c0 = blockIdx.y * slab;
    const int c1 = min(c, c0 + slab);
    const long long e0 = ((long long)blockIdx.x * G_THREADS + threadIdx.x) * 4;
    if (e0 >= e_total) return;
    const int *ix = idx + (size_t)b * e_total + e0;
    const float *src = points + ((size_t)b * c + c0) * n;
    float *dst = out + ((size_t)b * c + c0) * e_total + e0;
    const bool vec = (e0 + 3 < e_total) && ((e_total & 3) == 0);
    if (vec) {
        const int4 i4 = __ldg(reinterpret_cast<const int4 *>(ix));
#pragma unroll 4
        for (int ch = c0; ch < c1; ++ch) {
            float4 v;
            v.x = __ldg(src + i4.x);
            v.y = __ldg(src + i4.y);
            v.z = __ldg(src + i4.z);
            v.w = __ldg(src + i4.w);
            __stcs(reinterpret_cast<float4 *>(dst), v);
            src += n;
            dst += e_total;
        }
    } else {
        const int cnt = (int)min(4ll, e_total - e0);
        for (int ch = c0; ch < c1; ++ch) {
            for (int q = 0; q < cnt; ++q) dst[q] = __ldg(src + __ldg(ix + q));
            src += n;
            dst += e_total;
        }
    }
}

// grad_points[b,c,idx[b,e]] += grad_out[b,c,e]  (same accumulation primitive as the reference:
// fp32 atomic add, order not defined)
__global__ void __launch_bounds__(G_THREADS) scatter_rows_kernel(int c, int n, long long e_total,
                                                                 const float *__restrict__ grad_out,
                                                                 const int *__restrict__ idx,
                                                                 float *__restrict__ grad_points) {
    const int b = blockIdx.z;
    const int c0 = blockIdx.y * G_CH_SLAB;
    const int c1 = min(c, c0 + G_CH_SLAB);
    const long long e = (long long)blockIdx.x * G_THREADS + threadIdx.x;
    if (e >= e_total) return;
    const int id = __ldg(idx + (size_t)b * e_total + e);
    const float *g = grad_out + ((size_t)b * c + c0) * e_total + e;
    float *dst = grad_points + ((size_t)b * c + c0) * n + id;
    for (int ch = c0; ch < c1; ++ch) {
        atomicAdd(dst, __ldg(g));
        g += e_total;
        dst += n;
    }
}

// out[b,c,j] = fma(w2, p[i2], fma(w0, p[i0], w1*p[i1]))  -- the reference's sm_100 evaluation order
__global__ void __launch_bounds__(G_THREADS) three_interpolate_kernel(int c, int m, int n,
                                                                      const float *__restrict__ points,
                                                                      const int *__restrict__ idx,
                                                                      const float *__restrict__ weight,
                                                                      float *__restrict__ out) {
    const int b = blockIdx.z;
    const int c0 = blockIdx.y * G_CH_SLAB;
    const int c1 = min(c, c0 + G_CH_SLAB);
    const int j = blockIdx.x * G_THREADS + threadIdx.x;
    if (j >= n) return;
    const int *ix = idx + ((size_t)b * n + j) * 3;
    const float *w = weight + ((size_t)b * n + j) * 3;
    const int i0 = __ldg(ix + 0), i1 = __ldg(ix + 1), i2 = __ldg(ix + 2);
    const float w0 = __ldg(w + 0), w1 = __ldg(w + 1), w2 = __ldg(w + 2);
    const float *src = points + ((size_t)b * c + c0) * m;
    float *dst = out + ((size_t)b * c + c0) * n + j;
#pragma unroll 4
    for (int ch = c0; ch < c1; ++ch) {
        float t = __fmul_rn(w1, __ldg(src + i1));
        t = __fmaf_rn(w0, __ldg(src + i0), t);
        *dst = __fmaf_rn(w2, __ldg(src + i2), t);
        src += m;
        dst += n;
    }
}

// The same interpolation with the slab of source rows (slab x m floats) staged in shared memory: the three gathers per
// output become shared-memory reads (a few bank conflicts) instead of L1 requests that touch up to 32 sectors each, a
// thread owns 4 consecutive outputs (indices / weights loaded once as 128-bit vectors) and stores 128 bits per channel.
// (A point-major slab -- one 128-bit shared-memory load per source and 4 channels -- was measured slower: 28.6 vs 22.5 us
// at C=128, the transposing stores of the staging cost more than the gathers save.)
constexpr int TI_THREADS = 256;
__global__ void __launch_bounds__(TI_THREADS) three_interpolate_smem_kernel(int c, int m, int n, int slab,
                                                                               const float *__restrict__ points,
                                                                               const int *__restrict__ idx,
                                                                               const float *__restrict__ weight,
                                                                               float *__restrict__ out) {
    extern __shared__ __align__(16) float s_src[];   // slab x m
    const int b = blockIdx.z, c0 = blockIdx.y * slab, nc = min(slab, c - c0);
    const float *src = points + ((size_t)b * c + c0) * m;
    const int total = nc * m;
    if ((m & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        for (int i = threadIdx.x; i < total / 4; i += TI_THREADS) reinterpret_cast<float4 *>(s_src)[i] = __ldg(reinterpret_cast<const float4 *>(src) + i);
    } else {
        for (int i = threadIdx.x; i < total; i += TI_THREADS) s_src[i] = __ldg(src + i);
    }
    __syncthreads();
    // n % 4 == 0 (checked by the launcher): groups of 4 outputs never straddle the end
    for (int j0 = (blockIdx.x * TI_THREADS + threadIdx.x) * 4; j0 < n; j0 += gridDim.x * TI_THREADS * 4) {
        const int4 *ix = reinterpret_cast<const int4 *>(idx + ((size_t)b * n + j0) * 3);
        const float4 *wp = reinterpret_cast<const float4 *>(weight + ((size_t)b * n + j0) * 3);
        const int4 ia = __ldg(ix), ib = __ldg(ix + 1), ic = __ldg(ix + 2);
        const float4 wa = __ldg(wp), wb = __ldg(wp + 1), wc = __ldg(wp + 2);
        const int i[12] = {ia.x, ia.y, ia.z, ia.w, ib.x, ib.y, ib.z, ib.w, ic.x, ic.y, ic.z, ic.w};
        const float w[12] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w, wc.x, wc.y, wc.z, wc.w};
        float *dst = out + ((size_t)b * c + c0) * n + j0;
        const float *row = s_src;
        for (int ch = 0; ch < nc; ++ch) {
            float v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                // fma(w2, p[i2], fma(w0, p[i0], w1 * p[i1])): the reference's sm_100 evaluation order
                float t = __fmul_rn(w[3 * q + 1], row[i[3 * q + 1]]);
                t = __fmaf_rn(w[3 * q + 0], row[i[3 * q + 0]], t);
                v[q] = __fmaf_rn(w[3 * q + 2], row[i[3 * q + 2]], t);
            }
            __stcs(reinterpret_cast<float4 *>(dst), make_float4(v[0], v[1], v[2], v[3]));
            dst += n;
            row += m;
        }
    }
}

__global__ void __launch_bounds__(G_THREADS) three_interpolate_grad_kernel(int c, int n, int m,
                                                                           const float *__restrict__ grad_out,
                                                                           const int *__restrict__ idx,
                                                                           const float *__restrict__ weight,
                                                                           float *__restrict__ grad_points) {
    const int b = blockIdx.z;
    const int c0 = blockIdx.y * G_CH_SLAB;
    const int c1 = min(c, c0 + G_CH_SLAB);
    const int j = blockIdx.x * G_THREADS + threadIdx.x;
    if (j >= n) return;
    const int *ix = idx + ((size_t)b * n + j) * 3;
    const float *w = weight + ((size_t)b * n + j) * 3;
    const int i0 = __ldg(ix + 0), i1 = __ldg(ix + 1), i2 = __ldg(ix + 2);
    const float w0 = __ldg(w + 0), w1 = __ldg(w + 1), w2 = __ldg(w + 2);
    const float *g = grad_out + ((size_t)b * c + c0) * n + j;
    float *dst = grad_points + ((size_t)b * c + c0) * m;
    for (int ch = c0; ch < c1; ++ch) {
        const float gv = __ldg(g);
        atomicAdd(dst + i0, __fmul_rn(gv, w0));
        atomicAdd(dst + i1, __fmul_rn(gv, w1));
        atomicAdd(dst + i2, __fmul_rn(gv, w2));
        g += n;
        dst += m;
    }
}

// ---- gradients through shared memory ------------------------------------------------------------------------------
// The scatter kernels above issue one global fp32 atomic per (channel, element): ~200 G atomics/s on B200, the same
// primitive and the same cost as the reference.  For three_interpolate_grad, when a slab of channels of the destination
// (slab x m floats) fits in shared memory, a CTA that owns (cloud, slab) accumulates there with shared-memory atomics --
// no L2 round trip -- and adds the finished slab to the caller's buffer with plain coalesced read-modify-writes (it is
// the only writer of those rows): 128 -> 79 us at C=128, n=1024.  Accumulation order stays undefined, as in the reference.
constexpr int GS_THREADS = 512;
constexpr int GS_SMEM_MAX = 96 * 1024;

// grad_points[b, c0+ch, idx[b,j,q]] += grad_out[b, c0+ch, j] * weight[b,j,q],  q = 0..2
__global__ void __launch_bounds__(GS_THREADS) three_interpolate_grad_smem_kernel(int c, int n, int m, int slab,
                                                                                 const float *__restrict__ grad_out,
                                                                                 const int *__restrict__ idx,
                                                                                 const float *__restrict__ weight,
                                                                                 float *__restrict__ grad_points) {
    extern __shared__ float s_acc[];   // slab x m
    const int b = blockIdx.y, c0 = blockIdx.x * slab, nc = min(slab, c - c0);
    for (int i = threadIdx.x; i < nc * m; i += GS_THREADS) s_acc[i] = 0.0f;
    __syncthreads();
    const float *g = grad_out + ((size_t)b * c + c0) * n;
    for (int j = threadIdx.x; j < n; j += GS_THREADS) {
        const int *ix = idx + ((size_t)b * n + j) * 3;
        const float *w = weight + ((size_t)b * n + j) * 3;
        const int i0 = __ldg(ix + 0), i1 = __ldg(ix + 1), i2 = __ldg(ix + 2);
        const float w0 = __ldg(w + 0), w1 = __ldg(w + 1), w2 = __ldg(w + 2);
        for (int ch = 0; ch < nc; ++ch) {
            const float gv = __ldg(g + (size_t)ch * n + j);
            float *a = s_acc + ch * m;
            atomicAdd(a + i0, __fmul_rn(gv, w0));
            atomicAdd(a + i1, __fmul_rn(gv, w1));
            atomicAdd(a + i2, __fmul_rn(gv, w2));
        }
    }
    __syncthreads();
    float *dst = grad_points + ((size_t)b * c + c0) * m;
    for (int i = threadIdx.x; i < nc * m; i += GS_THREADS) dst[i] += s_acc[i];
}

// channels per CTA so that slab x n floats fit; 0 = destination too wide for shared memory (global-atomic kernels)
inline int smem_slab(int c, int n) {
    if (n <= 0) return 0;
    int slab = GS_SMEM_MAX / (int)sizeof(float) / n;
    if (slab < 1) return 0;
    slab = slab > 16 ? 16 : slab;
    return slab > c ? c : slab;
}
template <typename K>
inline bool smem_opt_in(K kernel) {
    static RtPerDevice done;   // one instance per kernel (template instantiation); per device, thread-safe
    const int dev = rt_current_device();
    if (done.done(dev)) return true;
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GS_SMEM_MAX) != cudaSuccess) return false;
    done.mark(dev);
    return true;
}

int launch_gather(int b, int c, int n, long long e_total, const float *points, const int *idx, float *out,
                  cudaStream_t st, const char *what) {
    if (b == 0 || c == 0 || e_total == 0) return RT_OK;
    RT_REQUIRE(b <= 65535, "%s: batch > 65535", what);
    // channels per CTA: 16 amortise the index loads; narrow tensors (xyz: C = 3) take one channel per CTA instead, or the
    // whole op is a handful of half-empty CTAs walking the channels serially (gather_points C=3: 9.1 -> 7.0 us)
    const int slab = c >= 32 ? G_CH_SLAB : (c >= 8 ? 4 : 1);
    dim3 grid(rt_divup(e_total, (long long)G_THREADS * 4), rt_divup(c, slab), b);
    gather_rows_kernel<<<grid, G_THREADS, 0, st>>>(c, n, e_total, slab, points, idx, out);
    return rt_check_launch(what);
}

int launch_scatter(int b, int c, int n, long long e_total, const float *grad_out, const int *idx, float *grad_points,
                   cudaStream_t st, const char *what) {
    if (b == 0 || c == 0 || e_total == 0) return RT_OK;
    RT_REQUIRE(b <= 65535, "%s: batch > 65535", what);
    // default: deterministic segmented sum over the inverse index (segsum.cu); the atomic kernel below is the reference's
    // primitive, kept for shapes the inverse index does not cover and for A/B timing (RT_GRAD_ATOMIC=1)
    if (rt_segsum_supported(n, e_total))
        return rt_launch_segmented_scatter(b, c, n, e_total, 1, grad_out, idx, nullptr, grad_points, st, what);
    // (a shared-memory variant of this scatter, like three_interpolate_grad_smem_kernel below, was measured SLOWER for
    // grouping gradients -- 421 vs 374 us at C=64, ns=32: ball-query padding repeats one index through the tail of a
    // group, and 32 lanes adding to one shared-memory word serialise; the L2 atomic units absorb that better)
    dim3 grid(rt_divup(e_total, G_THREADS), rt_divup(c, G_CH_SLAB), b);
    scatter_rows_kernel<<<grid, G_THREADS, 0, st>>>(c, n, e_total, grad_out, idx, grad_points);
    return rt_check_launch(what);
}

}  // namespace

// replaces gather_points_wrapper_fast (reference: src/lib/src/sampling.cpp:12-21)
RT_API int rt_gather_points(int b, int c, int n, int npoints, const float *points, const int *idx, float *out,
                            void *stream) {
    RT_REQUIRE(b >= 0 && c >= 0 && n >= 0 && npoints >= 0 && points && idx && out, "gather_points: bad arguments");
    return launch_gather(b, c, n, npoints, points, idx, out, (cudaStream_t)stream, "gather_points");
}

// replaces gather_points_grad_wrapper_fast (reference: src/lib/src/sampling.cpp:24-34)
RT_API int rt_gather_points_grad(int b, int c, int n, int npoints, const float *grad_out, const int *idx,
                                 float *grad_points, void *stream) {
    RT_REQUIRE(b >= 0 && c >= 0 && n >= 0 && npoints >= 0 && grad_out && idx && grad_points,
               "gather_points_grad: bad arguments");
    return launch_scatter(b, c, n, npoints, grad_out, idx, grad_points, (cudaStream_t)stream, "gather_points_grad");
}

// replaces group_points_wrapper_fast (reference: src/lib/src/group_points.cpp:27-38)
RT_API int rt_group_points(int b, int c, int n, int npoints, int nsample, const float *points, const int *idx,
                           float *out, void *stream) {
    RT_REQUIRE(b >= 0 && c >= 0 && n >= 0 && npoints >= 0 && nsample >= 0 && points && idx && out,
               "group_points: bad arguments");
    return launch_gather(b, c, n, (long long)npoints * nsample, points, idx, out, (cudaStream_t)stream,
                         "group_points");
}

// replaces group_points_grad_wrapper_fast (reference: src/lib/src/group_points.cpp:13-24)
RT_API int rt_group_points_grad(int b, int c, int n, int npoints, int nsample, const float *grad_out, const int *idx,
                                float *grad_points, void *stream) {
    RT_REQUIRE(b >= 0 && c >= 0 && n >= 0 && npoints >= 0 && nsample >= 0 && grad_out && idx && grad_points,
               "group_points_grad: bad arguments");
    return launch_scatter(b, c, n, (long long)npoints * nsample, grad_out, idx, grad_points, (cudaStream_t)stream,
                          "group_points_grad");
}

// replaces three_interpolate_wrapper_fast (reference: src/lib/src/interpolate.cpp:39-52)
RT_API int rt_three_interpolate(int b, int c, int m, int n, const float *points, const int *idx, const float *weight,
                                float *out, void *stream) {
    RT_REQUIRE(b >= 0 && c >= 0 && m >= 0 && n >= 0 && points && idx && weight && out,
               "three_interpolate: bad arguments");
    if (b == 0 || c == 0 || n == 0) return RT_OK;
    RT_REQUIRE(b <= 65535, "three_interpolate: batch > 65535");
    const int slab = smem_slab(c, m);
    const bool aligned = (n & 3) == 0 && (reinterpret_cast<uintptr_t>(idx) & 15) == 0 && (reinterpret_cast<uintptr_t>(weight) & 15) == 0 &&
                         (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    if (slab > 0 && aligned && smem_opt_in(three_interpolate_smem_kernel)) {
        // split the outputs of a (cloud, slab) over CTAs only while there are too few (cloud, slab) pairs to fill the GPU
        const int pairs = b * rt_divup(c, slab);
        int split = pairs >= 296 ? 1 : rt_divup(296, pairs);
        const int max_split = rt_divup(n, TI_THREADS * 4);
        split = split > max_split ? max_split : split;
        dim3 sgrid(split, rt_divup(c, slab), b);
        three_interpolate_smem_kernel<<<sgrid, TI_THREADS, (size_t)slab * m * sizeof(float), (cudaStream_t)stream>>>(c, m, n, slab, points,
                                                                                                              idx, weight, out);
        return rt_check_launch("three_interpolate");
    }
    dim3 grid(rt_divup(n, G_THREADS), rt_divup(c, G_CH_SLAB), b);
    three_interpolate_kernel<<<grid, G_THREADS, 0, (cudaStream_t)stream>>>(c, m, n, points, idx, weight, out);
    return rt_check_launch("three_interpolate");
}

// replaces three_interpolate_grad_wrapper_fast (reference: src/lib/src/interpolate.cpp:54-67)
RT_API int rt_three_interpolate_grad(int b, int c, int n, int m, const float *grad_out, const int *idx,
                                     const float *weight, float *grad_points, void *stream) {
    RT_REQUIRE(b >= 0 && c >= 0 && m >= 0 && n >= 0 && grad_out && idx && weight && grad_points,
               "three_interpolate_grad: bad arguments");
    if (b == 0 || c == 0 || n == 0) return RT_OK;
    RT_REQUIRE(b <= 65535, "three_interpolate_grad: batch > 65535");
    if (rt_segsum_supported(m, (long long)n * 3))
        return rt_launch_segmented_scatter(b, c, m, (long long)n * 3, 3, grad_out, idx, weight, grad_points, (cudaStream_t)stream,
                                           "three_interpolate_grad");
    const int slab = smem_slab(c, m);
    if (slab > 0 && smem_opt_in(three_interpolate_grad_smem_kernel)) {
        dim3 sgrid(rt_divup(c, slab), b);
        three_interpolate_grad_smem_kernel<<<sgrid, GS_THREADS, (size_t)slab * m * sizeof(float), (cudaStream_t)stream>>>(
            c, n, m, slab, grad_out, idx, weight, grad_points);
        return rt_check_launch("three_interpolate_grad");
    }
    dim3 grid(rt_divup(n, G_THREADS), rt_divup(c, G_CH_SLAB), b);
    three_interpolate_grad_kernel<<<grid, G_THREADS, 0, (cudaStream_t)stream>>>(c, n, m, grad_out, idx, weight,
                                                                                grad_points);
    return rt_check_launch("three_interpolate_grad");
}
