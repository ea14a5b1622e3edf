// Grouping in the channels-innermost layout of the training path: QueryAndGroup's result
//     [grouped_xyz - new_xyz ; grouped_features]            (reference: src/lib/pointnet2_utils.py:269-292)
// written directly as rows (b, npoint, nsample, 3 + c) -- the layout the dense layers consume (csrc/dense_tc.cu) -- instead of
// the reference's chain group_points(xyz) -> subtract -> group_points(features) -> cat, which builds three channel-major
// (b, C, npoint, nsample) tensors that the first 1x1 convolution then has to transpose.  torch.profiler on the training step
// (profiles/r2_train_profile.txt) showed 15 % of the device time in exactly those copies.
//
//   group_rows_kernel       one output row per LPR lanes (LPR = 8 / 16 / 32 by row width): the index is read once per row,
//                           the neighbour's feature row and the output row are contiguous -> coalesced on both sides;
//   group_rows_grad_kernel  gradient w.r.t. the feature rows as a SEGMENTED SUM over the stable inverse index
//                           (segsum.cu inverse_index_kernel): every destination point adds its sources in increasing source
//                           order -- bit-repeatable, no atomics (the reference: fp32 atomicAdd, group_points_gpu.cu:8-25).
#include "common.cuh"

bool rt_segsum_supported(int n_dst, long long e_total);   // segsum.cu
int rt_launch_inverse_index(int b, int n_dst, long long e_total, const int *idx, int *order, int *seg, cudaStream_t st, const char *what);

namespace {

constexpr int GR_THREADS = 256;

template <int LPR>
__global__ void __launch_bounds__(GR_THREADS) group_rows_kernel(long long rows, int c, int n, int npoint, int nsample,
                                                                const float *__restrict__ xyz, const float *__restrict__ new_xyz,
                                                                const float *__restrict__ feat, const int *__restrict__ idx,
                                                                float *__restrict__ out) {
    const int sub = threadIdx.x % LPR;
    const long long row = ((long long)blockIdx.x * GR_THREADS + threadIdx.x) / LPR;
    if (row >= rows) return;
    // (cloud, centre): a 32-bit division whenever the row index fits (a 64-bit one costs ~100 instructions per thread)
    const unsigned centre = rows <= 0xffffffffll ? (unsigned)row / (unsigned)nsample : (unsigned)(row / nsample);
    const unsigned cloud = centre / (unsigned)npoint;
    const int i = __ldg(idx + row);
    const long long g = (long long)cloud * n + i;
    const int ld = c + 3;
    float *o = out + row * ld;
    const float *f = feat + g * c - 3;
    for (int ch = sub; ch < ld; ch += LPR)
        o[ch] = ch < 3 ? __ldg(xyz + g * 3 + ch) - __ldg(new_xyz + (long long)centre * 3 + ch) : __ldg(f + ch);
}

// grad_feat[cloud, t, ch] = sum over the sources e of destination t (increasing e) of grad_rows[cloud, e, 3 + ch]
template <int LPR>
__global__ void __launch_bounds__(GR_THREADS) group_rows_grad_kernel(int b, int c, int n, long long e_total,
                                                                     const float *__restrict__ grad_rows, const int *__restrict__ order,
                                                                     const int *__restrict__ seg, float *__restrict__ grad_feat) {
    const int sub = threadIdx.x % LPR;
    const long long dst = ((long long)blockIdx.x * GR_THREADS + threadIdx.x) / LPR;   // (cloud, point)
    if (dst >= (long long)b * n) return;
    const bool small = (long long)b * n <= 0xffffffffll;
    const int cloud = small ? (int)((unsigned)dst / (unsigned)n) : (int)(dst / n);
    const int t = (int)(dst - (long long)cloud * n);
    const int *sg = seg + (long long)cloud * (n + 1);
    const int j0 = __ldg(sg + t), j1 = __ldg(sg + t + 1);
    const int *ord = order + (long long)cloud * e_total;
    const int ld = c + 3;
    const float *base = grad_rows + (long long)cloud * e_total * ld + 3;
    for (int ch = sub; ch < c; ch += LPR) {
        float s = 0.0f;
        for (int j = j0; j < j1; ++j) s += __ldg(base + (long long)__ldg(ord + j) * ld + ch);
        grad_feat[dst * c + ch] = s;
    }
}

int gr_lpr(int width) { return width <= 8 ? 8 : (width <= 16 ? 16 : 32); }

}  // namespace

RT_API int rt_group_rows(int b, int c, int n, int npoint, int nsample, const float *xyz, const float *new_xyz, const float *feat_rows,
                         const int *idx, float *out, void *stream) {
    RT_REQUIRE(b >= 0 && c >= 1 && n >= 1 && npoint >= 0 && nsample >= 1, "rt_group_rows: bad sizes");
    RT_REQUIRE(xyz && new_xyz && feat_rows && idx && out, "rt_group_rows: null pointer");
    const long long rows = (long long)b * npoint * nsample;
    if (rows == 0) return RT_OK;
    RT_REQUIRE(rows / nsample < (1ll << 32), "rt_group_rows: more than 2^32 centres");
    cudaStream_t st = (cudaStream_t)stream;
    const int lpr = gr_lpr(c + 3);
    const long long blocks = (rows * lpr + GR_THREADS - 1) / GR_THREADS;
    RT_REQUIRE(blocks < (1ll << 31), "rt_group_rows: grid too large");
    if (lpr == 8) group_rows_kernel<8><<<(unsigned)blocks, GR_THREADS, 0, st>>>(rows, c, n, npoint, nsample, xyz, new_xyz, feat_rows, idx, out);
    else if (lpr == 16) group_rows_kernel<16><<<(unsigned)blocks, GR_THREADS, 0, st>>>(rows, c, n, npoint, nsample, xyz, new_xyz, feat_rows, idx, out);
    else group_rows_kernel<32><<<(unsigned)blocks, GR_THREADS, 0, st>>>(rows, c, n, npoint, nsample, xyz, new_xyz, feat_rows, idx, out);
    return rt_check_launch("group_rows_kernel");
}

RT_API int rt_group_rows_grad(int b, int c, int n, int npoint, int nsample, const float *grad_rows, const int *idx, float *grad_feat_rows,
                              void *stream) {
    RT_REQUIRE(b >= 0 && c >= 1 && n >= 1 && npoint >= 0 && nsample >= 1, "rt_group_rows_grad: bad sizes");
    RT_REQUIRE(grad_rows && idx && grad_feat_rows, "rt_group_rows_grad: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const long long e_total = (long long)npoint * nsample;
    if (b == 0) return RT_OK;
    if (e_total == 0) {
        cudaMemsetAsync(grad_feat_rows, 0, sizeof(float) * (size_t)b * n * c, st);
        return RT_OK;
    }
    if (!rt_segsum_supported(n, e_total) || b > 65535) {
        rt_set_error("rt_group_rows_grad: n = %d points / %lld grouped rows per cloud is outside the inverse-index kernel's range", n, e_total);
        return RT_ERR_UNSUPPORTED;
    }
    int *scratch = nullptr;
    const size_t ints = (size_t)b * ((size_t)e_total + n + 1);
    int rc = rt_scratch_alloc((void **)&scratch, ints * sizeof(int), st, "rt_group_rows_grad");
    if (rc != RT_OK) return rc;
    int *order = scratch, *seg = scratch + (size_t)b * e_total;
    rc = rt_launch_inverse_index(b, n, e_total, idx, order, seg, st, "rt_group_rows_grad");
    if (rc == RT_OK) {
        const int lpr = gr_lpr(c);
        const long long blocks = ((long long)b * n * lpr + GR_THREADS - 1) / GR_THREADS;
        if (lpr == 8) group_rows_grad_kernel<8><<<(unsigned)blocks, GR_THREADS, 0, st>>>(b, c, n, e_total, grad_rows, order, seg, grad_feat_rows);
        else if (lpr == 16) group_rows_grad_kernel<16><<<(unsigned)blocks, GR_THREADS, 0, st>>>(b, c, n, e_total, grad_rows, order, seg, grad_feat_rows);
        else group_rows_grad_kernel<32><<<(unsigned)blocks, GR_THREADS, 0, st>>>(b, c, n, e_total, grad_rows, order, seg, grad_feat_rows);
        rc = rt_check_launch("group_rows_grad_kernel");
    }
    rt_scratch_free(scratch, st);
    return rc;
}
