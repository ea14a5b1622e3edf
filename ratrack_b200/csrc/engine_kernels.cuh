// Device kernels of the fused Track4D.backbone engine (eval mode, BN folded) -- declarations of the
// host-side launchers.  Internal layout convention: every feature tensor is ROW-MAJOR (points x channels),
// clouds concatenated along rows; the reference's channel-major (B,C,N) tensors exist only at the API edge.
#pragma once
#include "common.cuh"

enum { RT_ACT_NONE = 0, RT_ACT_RELU = 1, RT_ACT_LEAKY01 = 2 };

// One K-segment of a row GEMM: X rows (row stride ldx) times W (nout x k, row stride ldw)
struct RtSeg {
    const float *x;
    int ldx;
    int k;
    const float *w;
    int ldw;
};

// Y[r, o] = act( sum_seg X_seg[r, :] . W_seg[o, :] + bias[o] + cloud_bias[r / rows_per_cloud, o] )
struct RtRowGemm {
    int rows, nout;
    int nseg;
    RtSeg seg[4];
    const float *bias;        // nout or null
    const float *cloud_bias;  // (clouds x nout) or null
    int rows_per_cloud;
    int act;
    float *y;
    int ldy;
};
int rt_launch_rowgemm(const RtRowGemm &a, cudaStream_t st);

// out[(cloud,p,s), c] = act( Y[cloud_y, idx[cloud,p,s], yoff + c] + Wx[c,:].(xyz_in[idx] - xyz_c[p]) + bias[c] + Q[(cloud,p), c] )
struct RtGatherCombine {
    int clouds, npts, ns, c;
    const float *y;     // (clouds x n_in) rows, row stride ldy
    int ldy, yoff, n_in;
    const int *idx;     // (clouds, npts, ns)
    const float *xyz_in;  // (clouds, n_in, 3)
    const float *xyz_c;   // (clouds, npts, 3)
    const float *wx;      // (c x 3)
    const float *bias;    // c or null
    const float *q;       // (clouds*npts x c) or null
    int act;
    float *out;           // (clouds*npts*ns x c)
};
int rt_launch_gather_combine(const RtGatherCombine &a, cudaStream_t st);

// pooled[(cloud,p), coff + c] = max_s x[((cloud,p),s), c]
int rt_launch_maxpool_rows(int groups, int ns, int c, const float *x, float *pooled, int ldp, int coff, cudaStream_t st);

// WeightNet (3 -> 8 -> 8 -> c, ReLU each) weighted sum over the ns neighbours:
// out[(cloud,p), ch] = sum_s wn(xyz_in[idx[cloud,p,s]] - xyz_c[p])[ch] * V_s[ch]
//   V_s = v[((cloud,p),s), ch]            when gather_v == 0  (v holds one row per (point, neighbour))
//   V_s = v[(cloud, idx[cloud,p,s]), ch]  when gather_v == 1  (v holds one row per point)
struct RtWeightedSum {
    int clouds, npts, ns, c, n_in, gather_v;
    const int *idx;
    const float *xyz_in, *xyz_c;
    const float *wa, *ba, *wb, *bb, *wc, *bc;  // (8x3),(8),(8x8),(8),(c x 8),(c)
    const float *v;
    float *out;
    const int *perm;   // optional processing order: slot s of the grid handles point perm[s] (global row index); null = identity
};
int rt_launch_weighted_sum(const RtWeightedSum &a, cudaStream_t st);

// expanded-form k-NN used by the cost volume (reference: utils/model_utils/model_utils.py:17-39, 85-99):
// d = max((-2 q.s + |q|^2) + |s|^2, 0) in the reference's fp32 rounding order; idx (clouds, n, k) ascending.
// mcounts (optional): valid search points per cloud of a padded variable-size batch (cloud stride stays m)
int rt_launch_knn_expanded(int clouds, int n, int m, int k, const float *q, const float *s, int *idx, cudaStream_t st,
                           const int *mcounts = nullptr);

// inverse-distance weights of three_nn (reference: lib/pointnet2_modules.py:141-144), in place dist2 -> weight
int rt_launch_nn_weights(long long rows, float *dist2_to_w, cudaStream_t st);
// out[(cloud,p), c] = fma(w2, F[i2,c], fma(w0, F[i0,c], w1*F[i1,c]))   F (clouds x m rows, stride ldf)
int rt_launch_interp3(int clouds, int n, int m, int c, const float *f, int ldf, const int *idx, const float *w,
                      float *out, int ldo, cudaStream_t st);
// rows (clouds x npts_out) gathered from (clouds x n_in) rows of width c
int rt_launch_gather_rows(int clouds, int npts_out, int n_in, int c, const float *src, const int *idx, float *dst,
                          cudaStream_t st);
// g[cloud, c] = max_p F[(cloud,p), c]
// counts (optional): valid rows per cloud of a padded batch
int rt_launch_cloud_max(int clouds, int npts, int c, const float *f, int ldf, float *g, cudaStream_t st, const int *counts = nullptr);
// cb[cloud, o] = W[o, :k] . g[cloud, :k] + bias[o]
int rt_launch_cloud_matvec(int clouds, int nout, int k, const float *w, int ldw, const float *g, int ldg,
                           const float *bias, float *cb, cudaStream_t st);
// (B, C, N) channel-major  <->  (B*N, C) row-major (optionally into a column window of wider rows)
int rt_launch_cm_to_rows(int b, int c, int n, const float *src, float *dst, int ldd, int coff, cudaStream_t st);
int rt_launch_rows_to_cm(int b, int c, int n, const float *src, int lds, int soff, float *dst, int dst_c, int dst_coff,
                         cudaStream_t st);
// dst[(b, c, n)] = g[b, c] for all n (broadcast rows of the global feature into a channel-major output)
int rt_launch_broadcast_cm(int b, int c, int n, const float *g, float *dst, int dst_c, int dst_coff, cudaStream_t st);
// 5-layer GRU, one step: x (b,128), h_in / h_out (5, *, 128) with layer stride h_stride floats; y = h_out[4]
int rt_launch_gru_hh(int b, const float *h_in, const float *whh, const float *bhh, size_t h_stride, float *gh, cudaStream_t st);
int rt_launch_gru(int b, const float *x, const float *h_in, const float *wih, const float *bih, const float *gh, float *h_out,
                  size_t h_stride, cudaStream_t st);
// cls[(b,n)] = sigmoid(lin_w . (W4 . h3[(b,n), :32]) + lin_b);  flow written by rowgemm + rows_to_cm
int rt_launch_cls_tail(long long rows, const float *h3, const float *w4, const float *lin_w, const float *lin_b,
                       float *cls, cudaStream_t st);
int rt_launch_fill(float *p, long long n, float v, cudaStream_t st);
// if *status != 0: flow, cls and h_out (5 x b x 128, layer stride h_stride) become NaN (fp16-range guard fired)
int rt_launch_poison_on_status(const int *status, float *flow, long long nflow, float *cls, long long ncls, float *h_out, int b,
                               size_t h_stride, cudaStream_t st);
// Spatial processing order: perm[cloud*n + j] = cloud*n + (index of the j-th point of the cloud along a 30-bit Morton curve).
// Kernels that gather neighbour rows walk their points in this order so that the points of a tile share neighbours (L1 / L2
// hits instead of repeated 1 KB row fetches); every point is still computed independently and written to its own row, so the
// results are bit-identical to the identity order.
int rt_launch_morton_perm(int clouds, int n, const float *xyz, int *perm, cudaStream_t st);
// neighbors.cu: ball query of two radii over the same centres in one launch
// counts (optional): valid points per searched cloud of a padded variable-size batch (cloud stride stays n)
int rt_launch_ball_query2(int b, int n, int m, float radius_a, int nsample_a, int *idx_a, float radius_b, int nsample_b,
                          int *idx_b, const float *new_xyz, const float *xyz, int zero_fill, cudaStream_t st,
                          const int *counts = nullptr);
// fps.cu: FPS from the reference's initial state that also emits new_xyz = xyz[idx]; RT_ERR_UNSUPPORTED -> use the 3-launch path
int rt_launch_fps_fused(int b, int n, int m, const float *xyz, int *idx, float *new_xyz, cudaStream_t st);
// variable-size batch: cloud i has counts_host[i] (== counts_dev[i]) valid points, cloud stride n_stride.  Every cloud is
// sampled exactly as a stand-alone cloud of its own size would be (the tie-break depends on the reference's CTA size for
// that size), one launch per size class.  list_dev: scratch for b cloud indices.
int rt_launch_fps_fused_varlen(int b, int n_stride, int m, const float *xyz, const int *counts_host, const int *counts_dev,
                               int *list_dev, int *idx, float *new_xyz, cudaStream_t st);
// dst (b, c, n): zero the columns at and beyond counts[b]
int rt_launch_mask_cm(int b, int c, int n, const int *counts, float *dst, cudaStream_t st);
