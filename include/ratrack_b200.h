/*
 * ratrack_b200.h -- C ABI of libratrack_b200.so (hand-written sm_100a kernels for RaTrack's
 * per-frame point-cloud hot path).
 *
 * Conventions (all entry points):
 *   - plain C: device pointers + sizes, no torch / ATen types;
 *   - every tensor is contiguous fp32 (float) or int32 (int) in DEVICE memory, laid out exactly
 *     as the reference lays it out (shapes given per function);
 *   - the caller owns every buffer; the library never allocates on behalf of these calls;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); work is
 *     enqueued and the call returns immediately -- no synchronisation;
 *   - return value: 0 on success, <0 for an argument/shape error, >0 = the cudaError_t of a
 *     failed launch.  rt_last_error() returns a thread-local description.  The library never
 *     calls exit() (the reference's launchers do: e.g. src/lib/src/ball_query_gpu.cu:62-65).
 *
 * Section 1 is the drop-in boundary for the reference's compiled module `pointnet2_cuda`
 * (src/lib/src/pointnet2_api.cpp:11-24): one rt_* per pybind entry, same argument order.
 * Section 2 is the fused inference engine for Track4D.backbone.
 */
#ifndef RATRACK_B200_H
#define RATRACK_B200_H

#ifdef __cplusplus
extern "C" {
#endif

int rt_abi_version(void);
const char *rt_last_error(void);

/* ---------------------------------------------------------------------------------------------
 * Section 1 -- pointnet2 ops (Seam A)
 * ------------------------------------------------------------------------------------------ */

/* replaces ball_query_wrapper_fast       (reference: src/lib/src/ball_query.cpp:18-29,
 *                                          kernel src/lib/src/ball_query_gpu.cu:9-45)
 * new_xyz (b,m,3), xyz (b,n,3) -> idx (b,m,nsample).  First `nsample` points in index order with
 * d2 < radius*radius; unused slots repeat the first hit; rows with no hit are left untouched
 * (the reference's Python zero-fills idx first: src/lib/pointnet2_utils.py:246). */
int rt_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xyz, const float *xyz, int *idx,
                  void *stream);

/* replaces group_points_wrapper_fast      (reference: src/lib/src/group_points.cpp:27-38)
 * points (b,c,n), idx (b,npoints,nsample) -> out (b,c,npoints,nsample) */
int rt_group_points(int b, int c, int n, int npoints, int nsample, const float *points, const int *idx, float *out,
                    void *stream);

/* replaces group_points_grad_wrapper_fast (reference: src/lib/src/group_points.cpp:13-24)
 * grad_out (b,c,npoints,nsample), idx -> grad_points (b,c,n) accumulated into (caller zero-fills) */
int rt_group_points_grad(int b, int c, int n, int npoints, int nsample, const float *grad_out, const int *idx,
                         float *grad_points, void *stream);

/* replaces gather_points_wrapper_fast     (reference: src/lib/src/sampling.cpp:12-21)
 * points (b,c,n), idx (b,npoints) -> out (b,c,npoints) */
int rt_gather_points(int b, int c, int n, int npoints, const float *points, const int *idx, float *out, void *stream);

/* replaces gather_points_grad_wrapper_fast (reference: src/lib/src/sampling.cpp:24-34) */
int rt_gather_points_grad(int b, int c, int n, int npoints, const float *grad_out, const int *idx, float *grad_points,
                          void *stream);

/* replaces furthest_point_sampling_wrapper (reference: src/lib/src/sampling.cpp:37-47,
 *                                           kernel src/lib/src/sampling_gpu.cu:94-253)
 * xyz (b,n,3), temp (b,n) [caller fills with 1e10; holds final min distances on return]
 * -> idx (b,m), idx[:,0] = 0.  Bit-identical to the reference including arg-max ties. */
int rt_furthest_point_sampling(int b, int n, int m, const float *xyz, float *temp, int *idx, void *stream);

/* replaces knn_wrapper_fast               (reference: src/lib/src/interpolate.cpp:27-36)
 * unknown (b,n,3), known (b,m,3) -> dist2 (b,n,k) squared, idx (b,n,k), ascending, ties stable.
 * k > 200 is rejected (the reference silently overflows its stack: interpolate_gpu.cu:30-31). */
int rt_knn(int b, int n, int m, int k, const float *unknown, const float *known, float *dist2, int *idx,
           void *stream);

/* replaces three_nn_wrapper_fast          (reference: src/lib/src/interpolate.cpp:16-25)
 * unknown (b,n,3), known (b,m,3) -> dist2 (b,n,3) squared, idx (b,n,3) */
int rt_three_nn(int b, int n, int m, const float *unknown, const float *known, float *dist2, int *idx, void *stream);

/* replaces three_interpolate_wrapper_fast (reference: src/lib/src/interpolate.cpp:39-52)
 * points (b,c,m), idx (b,n,3), weight (b,n,3) -> out (b,c,n) */
int rt_three_interpolate(int b, int c, int m, int n, const float *points, const int *idx, const float *weight,
                         float *out, void *stream);

/* replaces three_interpolate_grad_wrapper_fast (reference: src/lib/src/interpolate.cpp:54-67)
 * grad_out (b,c,n), idx, weight -> grad_points (b,c,m) accumulated into (caller zero-fills) */
int rt_three_interpolate_grad(int b, int c, int n, int m, const float *grad_out, const int *idx, const float *weight,
                              float *grad_points, void *stream);

/* replaces square_distance + knn_point   (reference: src/utils/model_utils/model_utils.py:17-39, 85-99)
 * query (b,n,3), search (b,m,3) row-major -> idx (b,n,k) int32, k <= 32: the k smallest expanded-form distances
 * max((-2 q.s + |q|^2) + |s|^2, 0), evaluated in the fp32 rounding order of torch.matmul / torch.sum on both CPU
 * and B200 (fma(z,z',fma(y,y',x*x')), (x^2+y^2)+z^2); ascending, ties -> lower index. */
int rt_knn_expanded(int b, int n, int m, int k, const float *query, const float *search, int *idx, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Section 2 -- fused inference engine for Track4D.backbone (Seam C)
 *
 * replaces, in eval mode, the whole of Track4D.backbone (reference: src/models/track4d.py:67-106):
 * feature_extraction_head (two PNHead passes, src/utils/model_utils/model_utils.py:393-424),
 * FeatureCorrelator.forward (:193-250) and FlowDecoder.forward (:281-305).
 * ------------------------------------------------------------------------------------------ */
typedef struct rt_engine rt_engine;

/* number of entries of the weight-pointer table rt_engine_create expects */
int rt_engine_num_weights(void);

/* `weights`: table of DEVICE pointers to fp32 tensors, BatchNorm already folded, in the order documented in
 * ratrack_b200/engine.py (WEIGHT_ORDER) == struct EngineW of csrc/engine.cu.  The engine keeps the
 * pointers (no copy): the caller owns the tensors and must keep them alive.  npoint = FPS sample count of
 * all three SA levels (reference: configs.yaml:25 `npoints`). */
int rt_engine_create(rt_engine **out, int npoint, const void *const *weights, int nweights);
void rt_engine_destroy(rt_engine *e);

/* bytes of scratch device memory rt_backbone_forward needs for a batch of b pairs of n points (< 0 on error) */
long long rt_engine_workspace_bytes(const rt_engine *e, int b, int n);

/* optional: two cudaEvent_t recorded on the launch stream around the dominant kernel (the dense
 * cost-volume MLP) of every forward; pass NULL, NULL to disable */
int rt_engine_set_profile_events(rt_engine *e, void *start_event, void *stop_event);

/* flags (default 571 = bits 0,1,3,4,5,9): bit 0 = cost volume on the tcgen05 tensor-core kernel, bit 1 = every other dense layer on the
 * tcgen05 MLP kernel; cleared bits select the fp32 SIMT kernels of the same dataflow (A/B parity of the split-fp16
 * arithmetic).  Scheduling bits, all bit-identical in their results: bit 2 = split batches of >= 8 pairs over two concurrent lanes
 * (stream sets); bit 3 = every FPS CTA claims a whole SM; bit 4 = the cost-volume kNN starts beside the FPS chain; bit 5 = the
 * feature path runs on an engine-owned stream of middle priority; bit 6 = Morton-ordered cost-volume tiles; bit 7 = two clouds per
 * FPS CTA; bit 8 = self-kNN deferred behind the SA levels; bit 9 = FPS levels 2/3 by the parallel identity check where it
 * proves the serial sampler's result.  RT_ENGINE_FLAGS in the environment overrides the default at creation. */
int rt_engine_set_flags(rt_engine *e, int flags);

/* how many lanes rt_backbone_forward will use for a batch of b pairs (1 or 2); lane 0 gets ceil(b/2) pairs */
int rt_engine_num_lanes(const rt_engine *e, int b);

/* blocking: device status word of the last rt_backbone_forward.  0 = ok; bit 1 = an activation of the cost-volume
 * MLP left the fp16 hi/lo range (|x| >= 65000): that call's outputs are invalid */
int rt_engine_last_status(rt_engine *e, int *status_out);

/* non-blocking form: enqueues on `stream` the copy of the status words (one per lane) into host_status[0..1], which must be
 * pinned host memory; read them after synchronising with `stream`.  Independently of either call, a forward whose guard fired
 * overwrites its flow / cls / h outputs with NaN on the device, so the failure is loud even when the status is never read. */
int rt_engine_status_async(rt_engine *e, int *host_status, void *stream);

/* kernels launched by this engine since creation */
long long rt_engine_launch_count(const rt_engine *e);

/* stage profile: on != 0 makes the next forwards record CUDA timing events at their stage boundaries (feature-path
 * stream); rt_engine_stage_times blocks on the last forward and writes up to `cap` intervals in milliseconds (and their
 * names, static strings) -- returns how many.  Measurement aid: no effect on results. */
int rt_engine_stage_profile(rt_engine *e, int on);
int rt_engine_stage_times(rt_engine *e, float *ms, const char **names, int cap);

/* pc1, pc2 (b,3,n); ft1, ft2 (b,2,n); h_in (5,b,128)  ->  flow (b,3,n), h_out (5,b,128), cls (b,n),
 * cor (b,256,n), f1, f2 (b,256,n), prop (b,128,n): the 7-tuple of Track4D.backbone (track4d.py:86).
 * knn12 / knn11 (b,n,16) int32, optional (NULL to skip): the cost volume's neighbour sets pc1->pc2, pc1->pc1
 * (model_utils.py:216,239), exposed for tie-aware parity checks.  All device pointers; workspace must be
 * 256-byte aligned.  Asynchronous on `stream`. */
int rt_backbone_forward(rt_engine *e, int b, int n, const float *pc1, const float *pc2, const float *ft1,
                        const float *ft2, const float *h_in, float *flow, float *h_out, float *cls, float *cor,
                        float *f1, float *f2, float *prop, int *knn12, int *knn11, void *workspace,
                        long long workspace_bytes, void *stream);

/* Variable-size batch.  Real radar frames have 240..350 points each (reference: src/dataset_classes/track_vod_3d.py:49-122)
 * and the reference runs them one at a time (batch 1, src/models/track4d.py:56).  Here b pairs of different sizes share one
 * call: the clouds are zero-padded to n columns and npts1 / npts2 -- HOST arrays of b ints -- give the valid points of every
 * pc1 / pc2 cloud (16 <= npts <= n).  Padded points do not exist for the network: every neighbour search is restricted to a
 * cloud's own points and every cloud is sampled with the tie-break of its own size, so each pair's outputs equal, bit for
 * bit, what rt_backbone_forward returns for that pair alone at its own size; output columns of padded points are zero. */
int rt_backbone_forward_varlen(rt_engine *e, int b, int n, const int *npts1, const int *npts2, const float *pc1,
                               const float *pc2, const float *ft1, const float *ft2, const float *h_in, float *flow,
                               float *h_out, float *cls, float *cor, float *f1, float *f2, float *prop, int *knn12,
                               int *knn11, void *workspace, long long workspace_bytes, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Section 3 -- association
 * ------------------------------------------------------------------------------------------ */

/* replaces log_optimal_transport + the matching of Track4D.sinkhorn_module
 *   (reference: src/models/utils/track4d_utils.py:405-434, src/models/track4d.py:166-180)
 * aff (b,m,n) affinities of m previous x n current objects -> scores (b,m+1,n+1) log-couplings after `iters` log-space
 * Sinkhorn iterations with dustbin score `alpha` [optional, may be NULL], indices0 (b,m) [optional] and indices1 (b,n):
 * the mutually-best partner of every object, -1 where there is none (int64, as torch returns them).  Up to 127 objects on a
 * side the coupling matrix lives in shared memory; up to 2047 it moves to a stream-ordered scratch (cudaMallocAsync). */
int rt_sinkhorn_match(int b, int m, int n, const float *aff, float alpha, int iters, float *scores, long long *indices0,
                      long long *indices1, void *stream);

/* replaces the per-object embedding assembly of Track4D.affinity_module (reference: src/models/track4d.py:202-216)
 * points (139, p_total) channel-major: the feature columns of nobj objects side by side, object i = columns
 * [offsets[i], offsets[i+1]) (offsets: nobj + 1 device ints, every object non-empty) -> out (nobj, 141) =
 * [mean position (3), position variance (3), max feature (128), mean flow (3), mean rrv (2), rrv variance (2)],
 * variances = population variances, two-pass. */
int rt_object_embeddings(int nobj, int p_total, const float *points, const int *offsets, float *out, void *stream);

/* replaces the host-side sklearn call of Track4D.clustering (reference: src/models/track4d.py:36,108-126)
 * x (b,n,d) fp32 feature rows -> labels (b,n) int32: exactly sklearn.cluster.DBSCAN(eps, min_samples).fit_predict per set
 * (cluster numbers in order of first core point, border points to the lowest-numbered adjacent cluster, noise = -1).
 * d <= 16; n <= 1024 keeps the adjacency in shared memory, n <= 16384 uses a stream-ordered scratch (cudaMallocAsync). */
int rt_dbscan(int b, int n, int d, const float *x, float eps, int min_samples, int *labels, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Section 4 -- dense layers of the training step on the tcgen05 tensor cores
 *
 * The reference trains through torch autograd: every 1x1 convolution / Linear of SharedMLP, FeatureCorrelator,
 * FlowPredictor, ClsPredictor and PNHead (reference: src/lib/pytorch_utils.py:35-101,
 * src/utils/model_utils/model_utils.py:223-231, 308-357, 393-424) is a (rows x k).(n x k)^T product over all
 * (batch, point, neighbour) rows, served by cuDNN / cuBLAS fp32 kernels in forward, dgrad and wgrad.  These entry
 * points are those three products: fp32 in / out, split-fp16 operands with fp32 TMEM accumulation (fp32-class
 * accuracy), bit-repeatable.  Row-major operands with explicit leading dimensions (in elements).
 * ------------------------------------------------------------------------------------------ */

/* amax_out[0] = max |x[i]|, i < count (device scalar; the kernels below read it on the device to scale an operand of
 * unbounded dynamic range -- a gradient -- into the range of the fp16 planes) */
int rt_absmax(const float *x, long long count, float *amax_out, void *stream);

/* y[r, j] = act(sum_i x[r, i] * W'[j, i] + bias[j]),  W'[j, i] = w[j * w_sn + i * w_sk],  r < rows, i < k, j < n.
 * forward: w = the layer's (n x k) weight, w_sn = k, w_sk = 1.  dgrad (dX = dY . W): x = dY, k <-> n, w_sn = 1,
 * w_sk = (row length of W).  bias (n) may be NULL.  x_amax / w_amax: device scalars from rt_absmax (max |x|, max |w|);
 * with them the operands are scaled by powers of two into the fp16 planes' range and any finite fp32 input is valid.
 * NULL = unscaled x / 2^10 w: then a value beyond the fp16 range (|x| >= 65504, |w| >= 64) yields inf / NaN in y,
 * never a silently saturated result.  act: 0 none, 1 ReLU, 2 LeakyReLU(0.1). */
int rt_dense_tc_forward(long long rows, int k, int n, const float *x, long long ldx, const float *w, long long w_sn,
                        long long w_sk, const float *bias, const float *x_amax, const float *w_amax, int act, float *y,
                        long long ldy, void *stream);

/* dw[j, i] = sum_r dy[r, j] * x[r, i]  (n x k, contiguous, overwritten).  dy_amax, x_amax: device scalars or NULL.
 * The rows are split over the grid and the partial sums added in a fixed order (scratch from the library's
 * stream-ordered pool). */
int rt_dense_tc_wgrad(long long rows, int n, int k, const float *dy, long long lddy, const float *x, long long ldx,
                      const float *dy_amax, const float *x_amax, float *dw, void *stream);

/* replaces, in the channels-innermost layout of the training path, QueryAndGroup's chain
 *   group_points(xyz) - new_xyz ; group_points(features) ; cat     (reference: src/lib/pointnet2_utils.py:269-292)
 * xyz (b,n,3), new_xyz (b,npoint,3), feat_rows (b,n,c) [features with the channels innermost], idx (b,npoint,nsample)
 * -> out (b,npoint,nsample,3+c): row = [xyz[idx] - new_xyz (3), feat_rows[idx] (c)] */
int rt_group_rows(int b, int c, int n, int npoint, int nsample, const float *xyz, const float *new_xyz,
                  const float *feat_rows, const int *idx, float *out, void *stream);

/* its gradient w.r.t. the features (the reference: group_points_grad_kernel_fast, fp32 atomicAdd,
 * src/lib/src/group_points_gpu.cu:8-25): grad_rows (b,npoint,nsample,3+c), idx -> grad_feat_rows (b,n,c), overwritten;
 * a segmented sum in increasing source order: bit-repeatable. */
int rt_group_rows_grad(int b, int c, int n, int npoint, int nsample, const float *grad_rows, const int *idx,
                       float *grad_feat_rows, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Section 5 -- cost volume of the training step, channels innermost
 *
 * replaces, for FeatureCorrelator.forward and its autograd backward (reference:
 * src/utils/model_utils/model_utils.py:193-250), the op-by-op chain over (B,256,16,N) tensors by fused kernels
 * on row-major tensors (point, neighbour, channel).  c = channels (multiple of 32, <= 1024), k <= 32 neighbours.
 * Every gradient is a fixed-order sum: bit-repeatable.
 * ------------------------------------------------------------------------------------------ */

/* first layer, project-then-gather (model_utils.py:216-231):
 * out[p,j,:] = LeakyReLU_0.1(p2[nbr(p,j),:] + p1[p,:] + wd.(xyz2[nbr(p,j)] - xyz1[p]) + bias)
 * p1 (b,n1,c), p2 (b,n2,c) per-point projections; xyz1 (b,n1,3), xyz2 (b,n2,3); idx (b,n1,k) neighbours in cloud 2;
 * wd (c,3); bias (c) or NULL -> out (b,n1,k,c), dir_out (b,n1,k,3) [optional]: the direction vectors */
int rt_cv1_forward(int b, int n1, int n2, int k, int c, const float *p1, const float *p2, const float *xyz1,
                   const float *xyz2, const int *idx, const float *wd, const float *bias, float *out, float *dir_out,
                   void *stream);

/* dout, out (b,n1,k,c), dir (b,n1,k,3), idx -> dp1 (b,n1,c), dp2 (b,n2,c), dwd (c,3), dbias (c), all overwritten */
int rt_cv1_backward(int b, int n1, int n2, int k, int c, const float *dout, const float *out, const float *dir,
                    const int *idx, float *dp1, float *dp2, float *dwd, float *dbias, void *stream);

/* backward prologue of a dense layer whose activation was fused into the GEMM epilogue (rt_dense_tc_forward act 1 / 2):
 * g = dy * act'(y) over (rows, n) contiguous, amax_out[0] = max |g|, colsum (n) = column sums of g (the bias gradient) */
int rt_act_grad(long long rows, int n, int act, const float *y, const float *dy, float *g, float *amax_out,
                float *colsum, void *stream);

/* WeightNet-weighted neighbour sum (model_utils.py:232-236, 239-248):
 * out[p,:] = sum_j ReLU(w3.h2[p,j,:] + b3) * X[p,j,:],  X[p,j] = x[p,j] (idx NULL; x (b,n,k,c)) or x[idx[p,j]]
 * (x (b,n,c) per-point rows, idx (b,n,k)); h2 (b,n,k,8) = WeightNet's hidden rows, w3 (c,8), b3 (c) -> out (b,n,c) */
int rt_wsum_forward(int b, int n, int k, int c, const float *x, const int *idx, const float *h2, const float *w3,
                    const float *b3, float *out, void *stream);

/* dout (b,n,c) -> dx (shape of x), dh2 (b,n,k,8), dw3 (c,8), db3 (c), all overwritten */
int rt_wsum_backward(int b, int n, int k, int c, const float *x, const int *idx, const float *h2, const float *w3,
                     const float *b3, const float *dout, float *dx, float *dh2, float *dw3, float *db3, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* RATRACK_B200_H */
