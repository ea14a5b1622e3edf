// Deterministic gradients of the gather-type ops (group_points, gather_points, three_interpolate).
//
// The reference scatters with fp32 atomicAdd (src/lib/src/group_points_gpu.cu:8-25, sampling_gpu.cu:46-63,
// interpolate_gpu.cu:192-214): the summation order -- and therefore the low bits of every gradient -- changes from run to run.
// Here a scatter-add is evaluated as a SEGMENTED SUM over the inverse index:
//   1. inverse_index_kernel   per cloud, a stable counting sort of the E source positions by their destination: histogram
//                             (integer shared-memory atomics: exact), exclusive scan, then an in-order placement in which
//                             equal keys inside a 32-wide chunk are ranked with match.any -- so every destination's sources
//                             end up in increasing source order, always;
//   2. segment_sum_kernel     one thread per destination (coalesced over destinations), 8 channels at a time: it adds its
//                             sources in that fixed order, products rounded like the reference's (gv * w, then +=).
// No floating-point atomics anywhere: two runs give bit-identical gradients.  Scratch (E + n + 1 ints per cloud) is
// stream-ordered (cudaMallocAsync), so the C ABI keeps the reference's signatures.
#include <stdlib.h>

#include "common.cuh"

namespace {

// one warp per cloud; shared: n_dst counters
__global__ void __launch_bounds__(32) inverse_index_kernel(int n_dst, long long e_total, const int *__restrict__ idx_all,
                                                           int *__restrict__ order_all, int *__restrict__ seg_all) {
    extern __shared__ int s_cnt[];
    const int lane = threadIdx.x;
    const int *idx = idx_all + (size_t)blockIdx.x * e_total;
    int *order = order_all + (size_t)blockIdx.x * e_total;
    int *seg = seg_all + (size_t)blockIdx.x * (n_dst + 1);
    for (int i = lane; i < n_dst; i += 32) s_cnt[i] = 0;
    __syncwarp();
    for (long long e = lane; e < e_total; e += 32) {
        const int k = __ldg(idx + e);
        if (k >= 0 && k < n_dst) atomicAdd(&s_cnt[k], 1);     // an out-of-range index is dropped (the reference would write out of bounds)
    }
    __syncwarp();
    int run = 0;
    for (int base = 0; base < n_dst; base += 32) {
        const int i = base + lane;
        const int v = i < n_dst ? s_cnt[i] : 0;
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int up = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += up;
        }
        if (i < n_dst) {
            seg[i] = run + inc - v;
            s_cnt[i] = run + inc - v;          // from here on: next free slot of this destination
        }
        run += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) seg[n_dst] = run;
    __syncwarp();
    for (long long e0 = 0; e0 < e_total; e0 += 32) {
        const long long e = e0 + lane;
        int k = e < e_total ? __ldg(idx + e) : -1;
        const bool valid = k >= 0 && k < n_dst;
        if (!valid) k = -1 - lane;              // unique: matches nobody
        const unsigned same = __match_any_sync(0xffffffffu, k);
        const int rank = __popc(same & ((1u << lane) - 1u));
        const int base = valid ? s_cnt[k] : 0;
        __syncwarp();
        if (valid) {
            order[base + rank] = (int)e;
            if (rank == 0) s_cnt[k] = base + __popc(same);
        }
        __syncwarp();
    }
}

constexpr int SS_THREADS = 128, SS_CH = 8;

// grad_points[b, ch, t] += sum_{j in seg(b,t)} weight[b, order[j]] * grad_out[b, ch, order[j] / src_div]
__global__ void __launch_bounds__(SS_THREADS) segment_sum_kernel(int c, int n_dst, long long e_total, int src_div,
                                                                 const float *__restrict__ grad_out, const float *__restrict__ weight,
                                                                 const int *__restrict__ order_all, const int *__restrict__ seg_all,
                                                                 float *__restrict__ grad_points) {
    const int b = blockIdx.z, c0 = blockIdx.y * SS_CH, t = blockIdx.x * SS_THREADS + threadIdx.x;
    if (t >= n_dst) return;
    const int nc = min(SS_CH, c - c0);
    const long long e_src = e_total / src_div;
    const int *order = order_all + (size_t)b * e_total;
    const int *seg = seg_all + (size_t)b * (n_dst + 1);
    const float *g = grad_out + ((size_t)b * c + c0) * e_src;
    const float *w = weight ? weight + (size_t)b * e_total : nullptr;
    const int s0 = __ldg(seg + t), s1 = __ldg(seg + t + 1);
    float acc[SS_CH];
#pragma unroll
    for (int i = 0; i < SS_CH; ++i) acc[i] = 0.0f;
    for (int j = s0; j < s1; ++j) {
        const int e = __ldg(order + j);
        const long long col = src_div == 1 ? e : e / src_div;
        const float wv = w ? __ldg(w + e) : 1.0f;
#pragma unroll
        for (int i = 0; i < SS_CH; ++i)
            if (i < nc) {
                const float gv = __ldg(g + (size_t)i * e_src + col);
                acc[i] += w ? __fmul_rn(gv, wv) : gv;
            }
    }
    if (s1 > s0) {
        float *dst = grad_points + ((size_t)b * c + c0) * n_dst + t;
#pragma unroll
        for (int i = 0; i < SS_CH; ++i)
            if (i < nc) dst[(size_t)i * n_dst] += acc[i];
    }
}

}  // namespace

// 1 = the deterministic path can serve this shape (RT_GRAD_ATOMIC=1 forces the reference-style atomic kernels for A/B timing)
bool rt_segsum_supported(int n_dst, long long e_total) {
    static int atomic_env = -1;
    if (atomic_env < 0) {
        const char *env = getenv("RT_GRAD_ATOMIC");
        atomic_env = (env && atoi(env) == 1) ? 1 : 0;
    }
    return !atomic_env && n_dst >= 1 && n_dst <= 48 * 1024 && e_total < (1ll << 31);
}

// grad_points (b, c, n_dst) += scatter of grad_out (b, c, e_total / src_div) through idx (b, e_total) [x weight (b, e_total)]
int rt_launch_segmented_scatter(int b, int c, int n_dst, long long e_total, int src_div, const float *grad_out, const int *idx,
                                const float *weight, float *grad_points, cudaStream_t st, const char *what) {
    if (b == 0 || c == 0 || e_total == 0 || n_dst == 0) return RT_OK;
    RT_REQUIRE(b <= 65535, "%s: batch > 65535", what);
    static RtPerDevice attr;
    const int dev = rt_current_device();
    if (!attr.done(dev)) {
        const cudaError_t e = cudaFuncSetAttribute(inverse_index_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) {
            rt_set_error("%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e));
            return (int)e;
        }
        attr.mark(dev);
    }
    int *scratch = nullptr;
    const size_t ints = (size_t)b * ((size_t)e_total + n_dst + 1);
    const cudaError_t me = cudaMallocAsync(&scratch, ints * sizeof(int), st);
    if (me != cudaSuccess) {
        rt_set_error("%s: cudaMallocAsync(%zu): %s", what, ints * sizeof(int), cudaGetErrorString(me));
        return (int)me;
    }
    int *order = scratch, *seg = scratch + (size_t)b * e_total;
    inverse_index_kernel<<<b, 32, (size_t)n_dst * sizeof(int), st>>>(n_dst, e_total, idx, order, seg);
    int rc = rt_check_launch(what);
    if (rc == RT_OK) {
        dim3 grid(rt_divup(n_dst, SS_THREADS), rt_divup(c, SS_CH), b);
        segment_sum_kernel<<<grid, SS_THREADS, 0, st>>>(c, n_dst, e_total, src_div, grad_out, weight, order, seg, grad_points);
        rc = rt_check_launch(what);
    }
    cudaFreeAsync(scratch, st);
    return rc;
}
