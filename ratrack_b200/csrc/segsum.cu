// Deterministic gradients of the gather-type ops (group_points, gather_points, three_interpolate).
//
// The reference scatters with fp32 atomicAdd (src/lib/src/group_points_gpu.cu:8-25, sampling_gpu.cu:46-63,
// interpolate_gpu.cu:192-214): the summation order -- and therefore the low bits of every gradient -- changes from run to run.
// Here a scatter-add is evaluated as a SEGMENTED SUM over the inverse index:
//   1. inverse_index_kernel   per cloud (one CTA, 16 warps), a stable counting sort of the E source positions by their destination: histogram
//                             (integer shared-memory atomics: exact), exclusive scan, then an in-order placement in which
//                             equal keys inside a 32-wide chunk are ranked with match.any -- so every destination's sources
//                             end up in increasing source order, always;
//   2. segment_sum_kernel     per (cloud, channel group): inverse index in shared memory, every grad_out row staged with
//                             coalesced loads, a few lanes per destination adding its sources in a fixed order (products
//                             rounded like the reference's gv * w).
// No floating-point atomics anywhere: two runs give bit-identical gradients.  Scratch (E + n + 1 ints per cloud) is
// stream-ordered (rt_scratch_alloc: a cached per-device pool), so the C ABI keeps the reference's signatures.
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int II_THREADS = 512, II_WARPS = II_THREADS / 32;

// One CTA per cloud.  Shared: rows x n_dst counters (row = a warp's private histogram of its contiguous slice of the
// sources) + n_dst totals.  Stability: warp w owns sources [w * chunk, (w + 1) * chunk); inside a warp the sources are
// placed 32 at a time in increasing order, equal keys of a 32-wide step ranked with match.any -- so every destination's
// sources end up in increasing source order, always.
__global__ void __launch_bounds__(II_THREADS) inverse_index_kernel(int n_dst, long long e_total, int rows, const int *__restrict__ idx_all,
                                                                   int *__restrict__ order_all, int *__restrict__ seg_all) {
    extern __shared__ int s_ii[];
    int *s_cnt = s_ii;                               // [rows][n_dst]
    int *s_tot = s_ii + (size_t)rows * n_dst;        // [n_dst]
    __shared__ int s_wsum[II_WARPS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int *idx = idx_all + (size_t)blockIdx.x * e_total;
    int *order = order_all + (size_t)blockIdx.x * e_total;
    int *seg = seg_all + (size_t)blockIdx.x * (n_dst + 1);
    for (int i = tid; i < rows * n_dst; i += II_THREADS) s_cnt[i] = 0;
    __syncthreads();
    // slice of this warp (only the first `rows` warps take part in the histogram / placement passes)
    const long long chunk = ((e_total + rows - 1) / rows + 31) / 32 * 32;
    const long long e_beg = (long long)warp * chunk, e_end = min(e_total, e_beg + chunk);
    int *mine = s_cnt + (size_t)warp * n_dst;
    if (warp < rows)
        for (long long e = e_beg + lane; e < e_end; e += 32) {
            const int k = __ldg(idx + e);
            if (k >= 0 && k < n_dst) atomicAdd(&mine[k], 1);     // integer: exact.  An out-of-range index is dropped
        }
    __syncthreads();
    // per destination: exclusive scan over the warps' rows, total
    for (int k = tid; k < n_dst; k += II_THREADS) {
        int run = 0;
        for (int r = 0; r < rows; ++r) {
            const int v = s_cnt[(size_t)r * n_dst + k];
            s_cnt[(size_t)r * n_dst + k] = run;
            run += v;
        }
        s_tot[k] = run;
    }
    __syncthreads();
    // exclusive scan of the totals over the destinations: thread t owns a contiguous run of destinations
    const int per = (n_dst + II_THREADS - 1) / II_THREADS;
    const int k0 = tid * per, k1 = min(n_dst, k0 + per);
    int local = 0;
    for (int k = k0; k < k1; ++k) local += s_tot[k];
    int inc = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += up;
    }
    if (lane == 31) s_wsum[warp] = inc;
    __syncthreads();
    int before = inc - local;
    for (int w = 0; w < warp; ++w) before += s_wsum[w];
    for (int k = k0; k < k1; ++k) {
        const int v = s_tot[k];
        seg[k] = before;
        s_tot[k] = before;
        before += v;
    }
    if (tid == II_THREADS - 1) seg[n_dst] = before;
    __syncthreads();
    if (warp >= rows) return;
    for (int k = lane; k < n_dst; k += 32) mine[k] += s_tot[k];   // next free slot of (this warp, destination k)
    __syncwarp();
    for (long long e0 = e_beg; e0 < e_end; e0 += 32) {
        const long long e = e0 + lane;
        int k = e < e_end ? __ldg(idx + e) : -1;
        const bool valid = k >= 0 && k < n_dst;
        if (!valid) k = -1 - lane;              // unique: matches nobody
        const unsigned same = __match_any_sync(0xffffffffu, k);
        const int rank = __popc(same & ((1u << lane) - 1u));
        const int base = valid ? mine[k] : 0;
        __syncwarp();
        if (valid) {
            order[base + rank] = (int)e;
            if (rank == 0) mine[k] = base + __popc(same);
        }
        __syncwarp();
    }
}

constexpr int SS_THREADS = 512;

// grad_points[b, ch, t] += sum_{j in seg(b,t)} weight[b, order[j]] * grad_out[b, ch, order[j] / src_div]
// One CTA per (cloud, group of `cg` channels).  The cloud's inverse index (and weights) sit in shared memory for all
// channels of the group; CPS channels at a time, the grad_out rows are staged with coalesced 128-bit loads and gathered
// from there (the global gather of the first version read one 4-byte value per 32-byte sector: 0.02 of the HBM peak).
// LPD lanes share a destination: lane i adds sources i, i + LPD, ... in order, then a fixed shuffle tree combines the
// lanes -- a fixed summation order for given inputs, whatever the scheduling: two runs give bit-identical gradients.
// The sums go through shared memory and are added to the caller's rows in one coalesced pass (the CTA is their only writer).
template <int LPD, int CPS>
__global__ void __launch_bounds__(SS_THREADS) segment_sum_kernel(int c, int cg, int n_dst, int e_total, int src_div,
                                                                 const float *__restrict__ grad_out, const float *__restrict__ weight,
                                                                 const int *__restrict__ order_all, const int *__restrict__ seg_all,
                                                                 float *__restrict__ grad_points) {
    extern __shared__ __align__(16) int s_ss[];
    const int e_src = e_total / src_div, e_pad = (e_src + 3) & ~3;
    int *s_order = s_ss;                                                   // [e_total] source column (already divided by src_div)
    float *s_w = reinterpret_cast<float *>(s_ss + e_total);                // [e_total] (weighted form only)
    float *s_row = reinterpret_cast<float *>(s_ss + (weight ? 2 : 1) * (size_t)e_total);   // [CPS][e_pad], 16-byte aligned (e_total % 4 == 0)
    int *s_seg = reinterpret_cast<int *>(s_row + (size_t)CPS * e_pad);     // [n_dst + 1]
    float *s_out = reinterpret_cast<float *>(s_seg + n_dst + 1);           // [CPS][n_dst]
    const int b = blockIdx.y, c0 = blockIdx.x * cg, tid = threadIdx.x;
    const int nc = min(cg, c - c0);
    const int *order = order_all + (size_t)b * e_total;
    const int *seg = seg_all + (size_t)b * (n_dst + 1);
    for (int j = tid; j < e_total; j += SS_THREADS) {
        const int e = __ldg(order + j);
        s_order[j] = src_div == 1 ? e : e / src_div;
        if (weight) s_w[j] = __ldg(weight + (size_t)b * e_total + e);
    }
    for (int t = tid; t <= n_dst; t += SS_THREADS) s_seg[t] = __ldg(seg + t);
    const int sub = tid % LPD, grp = tid / LPD;
    constexpr int GROUPS = SS_THREADS / LPD;
    for (int ch = 0; ch < nc; ch += CPS) {
        __syncthreads();   // s_order / s_w / s_seg written; previous rows and sums no longer read
#pragma unroll
        for (int k = 0; k < CPS; ++k) {
            if (ch + k >= nc) break;
            const float *g = grad_out + ((size_t)b * c + c0 + ch + k) * e_src;
            float *row = s_row + (size_t)k * e_pad;
            if ((e_src & 3) == 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0) {
                for (int i = tid; i < e_src / 4; i += SS_THREADS) reinterpret_cast<float4 *>(row)[i] = __ldg(reinterpret_cast<const float4 *>(g) + i);
            } else {
                for (int i = tid; i < e_src; i += SS_THREADS) row[i] = __ldg(g + i);
            }
        }
        __syncthreads();
        for (int t0 = 0; t0 < n_dst; t0 += GROUPS) {   // uniform trip count: the shuffles below need every lane
            const int t = t0 + grp;
            const int s0 = t < n_dst ? s_seg[t] : 0, s1 = t < n_dst ? s_seg[t + 1] : 0;
            float acc[CPS];
#pragma unroll
            for (int k = 0; k < CPS; ++k) acc[k] = 0.0f;
            for (int j = s0 + sub; j < s1; j += LPD) {
                const int col = s_order[j];
                const float wv = weight ? s_w[j] : 1.0f;
#pragma unroll
                for (int k = 0; k < CPS; ++k) {
                    const float gv = s_row[(size_t)k * e_pad + col];   // (rows past nc hold stale data: their sums are never stored)
                    acc[k] += weight ? __fmul_rn(gv, wv) : gv;
                }
            }
#pragma unroll
            for (int k = 0; k < CPS; ++k) {
#pragma unroll
                for (int o = LPD / 2; o >= 1; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
                if (sub == 0 && t < n_dst) s_out[(size_t)k * n_dst + t] = acc[k];
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < CPS; ++k) {
            if (ch + k >= nc) break;
            float *dst = grad_points + ((size_t)b * c + c0 + ch + k) * n_dst;
            for (int t = tid; t < n_dst; t += SS_THREADS)
                if (s_seg[t + 1] > s_seg[t]) dst[t] += s_out[(size_t)k * n_dst + t];
        }
    }
}

// first version (global gathers, one thread per destination): kept for shapes whose index does not fit in shared memory
constexpr int SG_THREADS = 128, SG_CH = 8;
__global__ void __launch_bounds__(SG_THREADS) segment_sum_global_kernel(int c, int n_dst, long long e_total, int src_div,
                                                                        const float *__restrict__ grad_out, const float *__restrict__ weight,
                                                                        const int *__restrict__ order_all, const int *__restrict__ seg_all,
                                                                        float *__restrict__ grad_points) {
    const int b = blockIdx.z, c0 = blockIdx.y * SG_CH, t = blockIdx.x * SG_THREADS + threadIdx.x;
    if (t >= n_dst) return;
    const int nc = min(SG_CH, c - c0);
    const long long e_src = e_total / src_div;
    const int *order = order_all + (size_t)b * e_total;
    const int *seg = seg_all + (size_t)b * (n_dst + 1);
    const float *g = grad_out + ((size_t)b * c + c0) * e_src;
    const float *w = weight ? weight + (size_t)b * e_total : nullptr;
    const int s0 = __ldg(seg + t), s1 = __ldg(seg + t + 1);
    float acc[SG_CH];
#pragma unroll
    for (int i = 0; i < SG_CH; ++i) acc[i] = 0.0f;
    for (int j = s0; j < s1; ++j) {
        const int e = __ldg(order + j);
        const long long col = src_div == 1 ? e : e / src_div;
        const float wv = w ? __ldg(w + e) : 1.0f;
#pragma unroll
        for (int i = 0; i < SG_CH; ++i)
            if (i < nc) {
                const float gv = __ldg(g + (size_t)i * e_src + col);
                acc[i] += w ? __fmul_rn(gv, wv) : gv;
            }
    }
    if (s1 > s0) {
        float *dst = grad_points + ((size_t)b * c + c0) * n_dst + t;
#pragma unroll
        for (int i = 0; i < SG_CH; ++i)
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
    return !atomic_env && n_dst >= 1 && n_dst <= 24 * 1024 && e_total < (1ll << 31);
}

// the inverse index alone (group_rows.cu): order (b, e_total) = source positions sorted by destination, stable; seg (b, n_dst + 1)
int rt_launch_inverse_index(int b, int n_dst, long long e_total, const int *idx, int *order, int *seg, cudaStream_t st, const char *what) {
    constexpr size_t kSmemMax = 200 * 1024;
    static RtPerDevice attr;
    const int dev = rt_current_device();
    if (!attr.done(dev)) {
        const cudaError_t e = cudaFuncSetAttribute(inverse_index_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax);
        if (e != cudaSuccess) {
            rt_set_error("%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e));
            return (int)e;
        }
        attr.mark(dev);
    }
    int rows = (int)(kSmemMax / sizeof(int) / (size_t)n_dst) - 1;
    rows = rows > II_WARPS ? II_WARPS : rows;
    rows = rows < 1 ? 1 : rows;
    inverse_index_kernel<<<b, II_THREADS, (size_t)(rows + 1) * n_dst * sizeof(int), st>>>(n_dst, e_total, rows, idx, order, seg);
    return rt_check_launch(what);
}

// grad_points (b, c, n_dst) += scatter of grad_out (b, c, e_total / src_div) through idx (b, e_total) [x weight (b, e_total)]
int rt_launch_segmented_scatter(int b, int c, int n_dst, long long e_total, int src_div, const float *grad_out, const int *idx,
                                const float *weight, float *grad_points, cudaStream_t st, const char *what) {
    if (b == 0 || c == 0 || e_total == 0 || n_dst == 0) return RT_OK;
    RT_REQUIRE(b <= 65535, "%s: batch > 65535", what);
    constexpr size_t kSmemMax = 200 * 1024;
    static RtPerDevice attr;
    const int dev = rt_current_device();
    if (!attr.done(dev)) {
        cudaError_t e = cudaFuncSetAttribute(inverse_index_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(segment_sum_kernel<8, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(segment_sum_kernel<8, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(segment_sum_kernel<8, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(segment_sum_kernel<32, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(segment_sum_kernel<32, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(segment_sum_kernel<32, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax);
        if (e != cudaSuccess) {
            rt_set_error("%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e));
            return (int)e;
        }
        attr.mark(dev);
    }
    int *scratch = nullptr;
    const size_t ints = (size_t)b * ((size_t)e_total + n_dst + 1);
    const int ae = rt_scratch_alloc((void **)&scratch, ints * sizeof(int), st, what);
    if (ae != RT_OK) return ae;
    int *order = scratch, *seg = scratch + (size_t)b * e_total;
    // histogram rows: as many warps as the counters leave room for (a second array holds the totals)
    int rows = (int)(kSmemMax / sizeof(int) / (size_t)n_dst) - 1;
    rows = rows > II_WARPS ? II_WARPS : rows;
    rows = rows < 1 ? 1 : rows;
    inverse_index_kernel<<<b, II_THREADS, (size_t)(rows + 1) * n_dst * sizeof(int), st>>>(n_dst, e_total, rows, idx, order, seg);
    int rc = rt_check_launch(what);
    if (rc == RT_OK) {
        const long long e_src = e_total / src_div;
        const size_t e_pad = (size_t)((e_src + 3) & ~3ll);
        auto smem_for = [&](int cps) { return ((weight ? 2 : 1) * (size_t)e_total + cps * e_pad + (size_t)(cps + 1) * n_dst + 1) * 4; };
        if (smem_for(1) <= kSmemMax && (e_total & 3) == 0) {
            // channels staged per step: as many as fit (fewer CTA barriers and row-load round trips per channel)
            const int cps = smem_for(4) <= kSmemMax ? 4 : (smem_for(2) <= kSmemMax ? 2 : 1);
            // channels per CTA: the index is re-staged by every CTA of a cloud, so groups are as large as still fills the GPU
            int cg = 8;
            while (cg < 64 && (long long)b * rt_divup(c, 2 * cg) >= 296) cg *= 2;
            cg = cg < cps ? cps : cg;
            dim3 grid(rt_divup(c, cg), b);
            const bool few = e_total / n_dst <= 48;   // 8 lanes per destination keep four destinations of a warp in flight
#define RT_SS_LAUNCH(LPD, CPS) \
    segment_sum_kernel<LPD, CPS><<<grid, SS_THREADS, smem_for(CPS), st>>>(c, cg, n_dst, (int)e_total, src_div, grad_out, weight, order, seg, grad_points)
            if (few) {
                if (cps == 4) RT_SS_LAUNCH(8, 4); else if (cps == 2) RT_SS_LAUNCH(8, 2); else RT_SS_LAUNCH(8, 1);
            } else {
                if (cps == 4) RT_SS_LAUNCH(32, 4); else if (cps == 2) RT_SS_LAUNCH(32, 2); else RT_SS_LAUNCH(32, 1);
            }
#undef RT_SS_LAUNCH
        } else {
            dim3 grid(rt_divup(n_dst, SG_THREADS), rt_divup(c, SG_CH), b);
            segment_sum_global_kernel<<<grid, SG_THREADS, 0, st>>>(c, n_dst, e_total, src_div, grad_out, weight, order, seg, grad_points);
        }
        rc = rt_check_launch(what);
    }
    rt_scratch_free(scratch, st);
    return rc;
}
