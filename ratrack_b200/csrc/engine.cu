// Fused inference engine for Track4D.backbone (reference: src/models/track4d.py:67-106 and
// src/utils/model_utils/model_utils.py:166-424), eval mode: BatchNorm folded into the preceding 1x1
// convolutions by the host side (ratrack_b200/engine.py), every stage a CUDA kernel of this library.
//
// What is different from the reference's dataflow (same results up to fp32 re-association):
//   * geometry once: FPS / ball-query / three_nn / kNN depend only on xyz, so pc1's index structures are
//     computed once and shared by pn_head(pc1) and the FlowDecoder's second PNHead (`mse`);
//   * project-then-gather: the first 1x1 conv of every SA scale and of the cost volume is linear in the
//     gathered features, W.[dxyz; f_j] = Wx.dxyz + Wf.f_j, so Wf.f is evaluated once per POINT and the
//     grouped (B,C,S,ns) / (B,515,16,N) tensors of the reference are never built;
//   * channels that are constant over a cloud (global max-pool features, the GRU output) enter the next
//     layer as a per-cloud bias instead of being broadcast and concatenated.
#include <new>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../../include/ratrack_b200.h"
#include "engine_kernels.cuh"
#include "mlp_tc.cuh"

void rt_fps_set_exclusive(int on);  // fps.cu
void rt_fps_set_skip(const int *flags);
int rt_launch_fps_identity(int b, int n, const float *xyz, float *temp, int *ok, int *idx_a, float *xyz_a, int *idx_b, float *xyz_b,
                           cudaStream_t st);
int rt_launch_costvol_mlp(int rows, const float *x1, const float *w2, const float *b2, const float *w3, const float *b3,
                          float *xa, float *xb, cudaStream_t st);  // engine-internal, below
// costvol_tc.cu: gather + 3-layer MLP + WeightNet-weighted neighbour sum on tcgen05
int rt_launch_costvol_tc(int total_pts, int n, const float *p1, const float *p2, const float *xyz1, const float *xyz2,
                         const int *knn, const int *perm, const float *w1x, const void *wpack, const void *wcpack, const float *b2,
                         const float *b3, const float *bc, const float *wa, const float *ba, const float *wb, const float *bb,
                         float *out, int *status, cudaStream_t st);

namespace {

struct SaScaleW { const float *wx, *b1, *w2, *b2, *w3, *b3; };
struct HeadW {
    const float *wf_ft, *wf_loc, *wf_glob, *wf_cor;
    SaScaleW l1[2];
    const float *lin1_w, *lin1_b;
    const float *wf2;
    SaScaleW l2[2];
    const float *lin2_w, *lin2_b;
    const float *wf3;
    SaScaleW l3[2];
    const float *lin3_w, *lin3_b;
    const float *fp3_wi, *fp3_ws, *fp3_b, *fp2_wi, *fp2_ws, *fp2_b, *fp1_wi, *fp1_b;
};
struct WeightNetW { const float *wa, *ba, *wb, *bb, *wc, *bc; };
struct CvW {
    const float *w1_l1, *w1_g1, *w1_l2, *w1_g2, *w1_x, *b1, *w2, *b2, *w3, *b3;
    WeightNetW wn1, wn2;
    const float *w23_pack, *wc1_pack;  // fp16 hi/lo planes in UMMA core-matrix layout (costvol_tc.cu), made by engine.py
};
struct ClsW { const float *w1, *b1, *w2, *b2, *w3, *b3, *w4, *lin_w, *lin_b; };
struct FlowW { const float *w1_l, *w1_g, *b1, *w2, *b2, *w3, *b3, *w4; };
struct GruW { const float *wih, *whh, *bih, *bhh; };
struct EngineW {
    HeadW pn, mse;
    CvW cv;
    ClsW cp;
    FlowW fp;
    GruW gru;
};
constexpr int kNumWeights = sizeof(EngineW) / sizeof(const float *);
static_assert(sizeof(EngineW) == kNumWeights * sizeof(const float *), "EngineW must be a plain pointer table");
static_assert(kNumWeights == 157, "weight table order is mirrored by ratrack_b200/engine.py");

// PNHead geometry, hard-coded in the reference (utils/model_utils/model_utils.py:397-399)
struct LevelCfg { float radius[2]; int ns[2]; int c1[2], c2[2], c3[2]; int lin_out; };
const LevelCfg kLevels[3] = {
    {{2.f, 4.f}, {4, 8}, {16, 16}, {16, 16}, {32, 32}, 32},
    {{4.f, 8.f}, {8, 16}, {32, 32}, {32, 64}, {0, 0}, 64},
    {{8.f, 16.f}, {16, 32}, {64, 64}, {64, 64}, {0, 0}, 64},
};
constexpr int kKnn = 16;

// bump allocator over the caller's workspace; run with base == nullptr to size it
struct Carver {
    char *base;
    size_t off = 0;
    explicit Carver(void *b) : base((char *)b) {}
    template <typename T>
    T *take(size_t count) {
        off = (off + 255) & ~(size_t)255;
        T *p = base ? reinterpret_cast<T *>(base + off) : nullptr;
        off += count * sizeof(T);
        return p;
    }
};

struct Ws {
    float *xyz0, *ft0, *xyz[3], *temp;
    int *fps[3], *bq[3][2], *nn_idx[3], *knn12, *knn11, *perm1, *fps_ok, *npts, *fps_list;
    float *nn_w[3];
    float *proj, *xa, *xb, *pooled, *l1, *l2, *l3, *l2p, *l1p, *interp, *feat, *prop;
    float *gmax, *gprop, *cb_a, *cb_b, *p1, *p2, *cost1, *cor, *h1, *h2, *h3, *flow_rows, *gru_h, *gru_gh;
    int *status;
};

void carve(Carver &c, Ws &w, int b, int n, int S) {
    const size_t B2 = 2 * (size_t)b;
    w.xyz0 = c.take<float>(B2 * n * 3);
    w.ft0 = c.take<float>(B2 * n * 2);
    for (int l = 0; l < 3; ++l) w.xyz[l] = c.take<float>(B2 * S * 3);
    w.temp = c.take<float>(3 * B2 * (size_t)max(n, S));   // one FPS scratch per level
    for (int l = 0; l < 3; ++l) w.fps[l] = c.take<int>(B2 * S);
    w.fps_ok = c.take<int>(B2);
    w.npts = c.take<int>(B2);       // variable-size batches: valid points per cloud (pc1 clouds, then pc2 clouds)
    w.fps_list = c.take<int>(B2);
    for (int l = 0; l < 3; ++l)
        for (int s = 0; s < 2; ++s) w.bq[l][s] = c.take<int>(B2 * S * kLevels[l].ns[s]);
    // three_nn of FP3 (xyz2 <- xyz3), FP2 (xyz1 <- xyz2), FP1 (xyz0 <- xyz1)
    const size_t nn_rows[3] = {B2 * S, B2 * S, B2 * (size_t)n};
    for (int l = 0; l < 3; ++l) {
        w.nn_idx[l] = c.take<int>(nn_rows[l] * 3);
        w.nn_w[l] = c.take<float>(nn_rows[l] * 3);
    }
    w.knn12 = c.take<int>((size_t)b * n * kKnn);
    w.knn11 = c.take<int>((size_t)b * n * kKnn);
    w.perm1 = c.take<int>((size_t)b * n);
    const size_t head_rows = B2 * S * 32 * 64;  // widest SA activation: ns=32 x 64 channels
    const size_t cv_rows = (size_t)b * n * kKnn * 256;
    const size_t xbuf = head_rows > cv_rows ? head_rows : cv_rows;
    w.proj = c.take<float>(B2 * (size_t)max(n * 32, S * 128));
    w.xa = c.take<float>(xbuf);
    w.xb = c.take<float>(xbuf);
    w.pooled = c.take<float>(B2 * S * 128);
    w.l1 = c.take<float>(B2 * S * 32);
    w.l2 = c.take<float>(B2 * S * 64);
    w.l3 = c.take<float>(B2 * S * 64);
    w.l2p = c.take<float>(B2 * S * 128);
    w.l1p = c.take<float>(B2 * S * 128);
    w.interp = c.take<float>(B2 * (size_t)max(n, S) * 128);
    w.feat = c.take<float>(B2 * n * 128);
    w.prop = c.take<float>((size_t)b * n * 128);
    w.gmax = c.take<float>(B2 * 128);
    w.gprop = c.take<float>((size_t)b * 128);
    w.cb_a = c.take<float>((size_t)b * 256);
    w.cb_b = c.take<float>((size_t)b * 256);
    w.p1 = c.take<float>((size_t)b * n * 256);
    w.p2 = c.take<float>((size_t)b * n * 256);
    w.cost1 = c.take<float>((size_t)b * n * 256);
    w.cor = c.take<float>((size_t)b * n * 256);
    w.h1 = c.take<float>((size_t)b * n * 128);
    w.h2 = c.take<float>((size_t)b * n * 64);
    w.h3 = c.take<float>((size_t)b * n * 32);
    w.flow_rows = c.take<float>((size_t)b * n * 3);
    w.gru_h = c.take<float>((size_t)5 * b * 128);
    w.gru_gh = c.take<float>((size_t)5 * b * 384);
    w.status = c.take<int>(64);
}

#define RT_TRY(expr)              \
    do {                          \
        int rc_ = (expr);         \
        if (rc_ != RT_OK) return rc_; \
    } while (0)

RtRowGemm gemm1(long long rows, int nout, const float *x, int ldx, int k, const float *w, const float *bias, int act, float *y,
                int ldy) {
    RtRowGemm g{};
    g.rows = (int)rows;
    g.nout = nout;
    g.nseg = 1;
    g.seg[0] = RtSeg{x, ldx, k, w, k};
    g.bias = bias;
    g.rows_per_cloud = 1;
    g.act = act;
    g.y = y;
    g.ldy = ldy;
    return g;
}

}  // namespace

// tensor-core weight packs (fp16 hi/lo planes, mlp_tc.cuh layout), built on the device at engine creation
struct HeadPacks { void *proj[3], *w2[3][2], *w3[2], *lin[3], *fp3, *fp2, *fp1; };
struct Packs { HeadPacks pn, mse; void *p1, *p2, *cp[3], *fw[4]; };

// Streams and events of one in-flight sub-batch.  A forward call splits its batch over two lanes so the latency-bound
// stretches of one half (the FPS chain, small tail kernels) are filled by the throughput-bound kernels of the other.
struct Lane {
    // geometry runs beside the feature path on three streams: [0] the dependent FPS chain (latency-bound, 2B CTAs),
    // [1] ball queries + three_nn (need only the FPS result of their level), [2] the cost-volume kNN
    cudaStream_t main_stream = nullptr, geo_stream[3] = {nullptr, nullptr, nullptr};
    // output stream: API-layout transposes and the cls head hang off the feature path, nothing waits for them until the end
    cudaStream_t aux_stream = nullptr;
    cudaEvent_t ev_in = nullptr, ev_fps[3] = {nullptr, nullptr, nullptr}, ev_lvl[3] = {nullptr, nullptr, nullptr},
                ev_nn = nullptr, ev_knn = nullptr, ev_knn11 = nullptr, ev_sa3 = nullptr, ev_gh = nullptr, ev_feat = nullptr, ev_cor = nullptr, ev_prop = nullptr, ev_aux = nullptr,
                ev_done = nullptr;
};

struct rt_engine {
    EngineW w;
    Packs packs;
    void *arena = nullptr;
    int npoint;
    cudaEvent_t prof_start = nullptr, prof_stop = nullptr;
    long long launches = 0;
    Lane lanes[2];
    cudaEvent_t ev_fork = nullptr;
    int flags = 1595;              // bit 0: tensor-core cost volume (costvol_tc.cu); bit 1: tensor-core MLP chains
                                   // (mlp_tc.cu); cleared bits select the fp32 SIMT kernels of the same dataflow;
                                   // bit 2 (off by default: measured 4 % slower at 32 pairs, neutral at 128): split batches of >= 8 pairs over two lanes
                                   // bit 3: FPS CTAs claim a whole SM each (fps.cu launch_reg) so co-running kernels cannot stretch the chain
                                   // bit 4: cost-volume kNN starts with the FPS chain instead of behind it (pair with bit 3)
                                   // bit 7: with bit 3, two clouds share one FPS CTA (2b clouds block b SMs instead of 2b)
                                   // bit 6: the cost-volume kernels walk pc1 in Morton order (tiles of spatial neighbours share gathered rows)
                                   // bit 9: FPS levels 2 / 3 take the identity shortcut where fps_identity_kernel proves it (fps.cu)
                                   // bit 5: the feature path runs on an engine-owned stream of middle priority (geometry above it, kNN
                                   //        and API-layout outputs below it) forked from / joined to the caller's stream
    const int *last_status[2] = {nullptr, nullptr};
    std::vector<int> npts_host[2];   // per lane: point counts of a variable-size batch, kept alive across the async copy
    // optional stage profile (rt_engine_stage_profile): timing events recorded on the feature-path stream of lane 0 at the
    // boundaries of the forward, so the critical path can be read without a profiler's launch overhead distorting it
    static constexpr int kStages = 16;
    cudaEvent_t stage_ev[kStages] = {};
    int stage_n = 0;
    bool stage_on = false;
};

namespace {

// name of the interval that ENDS at mark i (mark 0 = inputs are row-major, geometry forked)
const char *const kStageNames[] = {"inputs->rows", "pn_head SA1 (waits for FPS-1 + ball query 1)", "pn_head SA2", "pn_head SA3", "pn_head FP",
                                   "global max + P1/P2 projections", "wait for the cost-volume kNN", "cost volume (costvol_tc)",
                                   "patch-to-patch sum", "mse SA1", "mse SA2", "mse SA3", "mse FP", "cloud max + GRU",
                                   "flow head + output join", "-"};
inline void stage_mark(rt_engine *e, int lane_idx, cudaStream_t st) {
    if (e->stage_on && lane_idx == 0 && e->stage_n < rt_engine::kStages) cudaEventRecord(e->stage_ev[e->stage_n++], st);
}

// geometry of all 2b clouds: FPS chain, ball queries, three_nn (+weights), cost-volume kNN.
// Every stage records an event the feature path (and the dependent geometry stages) wait on.
int run_geometry(rt_engine *e, Lane &L, Ws &w, int b, int n, const int *npts_host) {
    const int B2 = 2 * b, S = e->npoint;
    cudaStream_t s_fps = L.geo_stream[0], s_nbr = L.geo_stream[1], s_knn = L.geo_stream[2];
    const float *lvl_in[3] = {w.xyz0, w.xyz[0], w.xyz[1]};
    const int lvl_n[3] = {n, S, S};
    for (int g = 0; g < 3; ++g) cudaStreamWaitEvent(L.geo_stream[g], L.ev_in, 0);
    const float *pc1 = w.xyz0, *pc2 = w.xyz0 + (size_t)b * n * 3;
    const bool knn_early = (e->flags & 16) != 0;
    // Morton processing order of pc1 for the cost-volume kernels (bit 6); needs xyz only
    if (e->flags & 64) {
        RT_TRY(rt_launch_morton_perm(b, n, pc1, w.perm1, s_knn));
        e->launches += 1;
    }
    // bit 8: the self-kNN (pc1 -> pc1) is needed only by the patch-to-patch sum: it is launched by the feature path after
    // the last SA level of pn_head (run_self_knn) and fills the under-occupied FP / projection stages instead of competing
    // with the SA kernels
    const bool knn11_late = (e->flags & 256) != 0 && (e->flags & 2) != 0 && !npts_host;   // the hook lives in the tensor-core head
    // variable-size batch: per-cloud point counts on the device (pc1 clouds first, then pc2 clouds); nullptr = all n
    const int *cnt1 = npts_host ? w.npts : nullptr, *cnt2 = npts_host ? w.npts + b : nullptr;
    if (knn_early) {
        RT_TRY(rt_launch_knn_expanded(b, n, n, kKnn, pc1, pc2, w.knn12, s_knn, cnt2));
        cudaEventRecord(L.ev_knn, s_knn);
        if (!knn11_late) {
            RT_TRY(rt_launch_knn_expanded(b, n, n, kKnn, pc1, pc1, w.knn11, s_knn, cnt1));
            cudaEventRecord(L.ev_knn11, s_knn);
        }
        e->launches += 2;
    }
    rt_fps_set_exclusive(((e->flags & 8) ? 1 : 0) | ((e->flags & 128) ? 2 : 0));
    // Levels 2 and 3 sample S of the S points level 1 selected: the identity permutation unless a round ties, which
    // fps_identity_kernel checks exactly (fps.cu).  Clouds that pass are skipped by the serial sampler of both levels:
    // the dependent FPS chain shrinks from 3 x 511 rounds to 511 rounds + one parallel check.
    const bool identity = (e->flags & 512) != 0 && S >= 128 && S <= 1024;
    for (int l = 0; l < 3; ++l) {
        if (l == 1 && identity) {
            RT_TRY(rt_launch_fps_identity(B2, S, w.xyz[0], nullptr, w.fps_ok, w.fps[1], w.xyz[1], w.fps[2], w.xyz[2], s_fps));
            e->launches += 1;
        }
        rt_fps_set_skip(l >= 1 && identity ? w.fps_ok : nullptr);
        // one launch per level on the dependent chain: min-distance init, sampling and the gather of new_xyz are fused
        int rc = (l == 0 && npts_host)
                     ? rt_launch_fps_fused_varlen(B2, n, S, lvl_in[0], npts_host, w.npts, w.fps_list, w.fps[0], w.xyz[0], s_fps)
                     : rt_launch_fps_fused(B2, lvl_n[l], S, lvl_in[l], w.fps[l], w.xyz[l], s_fps);
        rt_fps_set_skip(nullptr);
        if (rc == RT_ERR_UNSUPPORTED && l == 0 && npts_host) {
            rt_set_error("backbone_forward_varlen: a cloud size has no register-resident FPS kernel (1 <= n <= 4096 expected)");
            return rc;
        }
        if (rc == RT_ERR_UNSUPPORTED) {   // cloud too large for the register-resident kernel
            RT_TRY(rt_launch_fill(w.temp + (size_t)l * B2 * max(n, S), (long long)B2 * lvl_n[l], 1e10f, s_fps));
            RT_TRY(rt_furthest_point_sampling(B2, lvl_n[l], S, lvl_in[l], w.temp + (size_t)l * B2 * max(n, S), w.fps[l], s_fps));
            RT_TRY(rt_launch_gather_rows(B2, S, lvl_n[l], 3, lvl_in[l], w.fps[l], w.xyz[l], s_fps));
            e->launches += 2;
            rc = RT_OK;
        }
        RT_TRY(rc);
        cudaEventRecord(L.ev_fps[l], s_fps);
        cudaStreamWaitEvent(s_nbr, L.ev_fps[l], 0);
        // both radii of the level in one pass; rows without a hit are zero-filled by the kernel (the reference zero-fills
        // the buffer from Python, lib/pointnet2_utils.py:246)
        RT_TRY(rt_launch_ball_query2(B2, lvl_n[l], S, kLevels[l].radius[0], kLevels[l].ns[0], w.bq[l][0], kLevels[l].radius[1],
                                     kLevels[l].ns[1], w.bq[l][1], w.xyz[l], lvl_in[l], 1, s_nbr, l == 0 ? cnt1 : nullptr));
        cudaEventRecord(L.ev_lvl[l], s_nbr);
        e->launches += 2;
    }
    rt_fps_set_exclusive(0);
    // FP3: unknown xyz[1] <- known xyz[2];  FP2: xyz[0] <- xyz[1];  FP1: xyz0 <- xyz[0]
    const float *unk[3] = {w.xyz[1], w.xyz[0], w.xyz0};
    const float *kn[3] = {w.xyz[2], w.xyz[1], w.xyz[0]};
    const int un[3] = {S, S, n};
    for (int l = 0; l < 3; ++l) {
        RT_TRY(rt_three_nn(B2, un[l], S, unk[l], kn[l], w.nn_w[l], w.nn_idx[l], s_nbr));
        RT_TRY(rt_launch_nn_weights((long long)B2 * un[l], w.nn_w[l], s_nbr));
        e->launches += 2;
    }
    cudaEventRecord(L.ev_nn, s_nbr);
    // cost-volume kNN: only after the FPS chain.  FPS is a chain of ~1500 dependent rounds on 2b CTAs; measured on
    // B200, any kernel sharing its SMs stretches every round 2-3x (438 us instead of 148 us for FPS-1 with the kNN
    // beside it), so the throughput-bound kNN is ordered behind it and overlaps the SA / FP kernels instead.
    if (!knn_early) {
        cudaStreamWaitEvent(s_knn, L.ev_fps[2], 0);
        RT_TRY(rt_launch_knn_expanded(b, n, n, kKnn, pc1, pc2, w.knn12, s_knn, cnt2));
        cudaEventRecord(L.ev_knn, s_knn);
        if (!knn11_late) {
            RT_TRY(rt_launch_knn_expanded(b, n, n, kKnn, pc1, pc1, w.knn11, s_knn, cnt1));
            cudaEventRecord(L.ev_knn11, s_knn);
        }
        e->launches += 2;
    }
    return RT_OK;
}

// One PNHead (utils/model_utils/model_utils.py:409-424) over the first `clouds` clouds of the geometry.
// Level-1 input features arrive as row segments; `cloud_bias1` carries cloud-constant channels.
int run_head(rt_engine *e, Lane &L, const HeadW &hw, Ws &w, int clouds, int n, const RtSeg *segs, int nseg, const float *cloud_bias1,
             float *out, cudaStream_t st) {
    const int S = e->npoint;
    const float *lvl_xyz_in[3] = {w.xyz0, w.xyz[0], w.xyz[1]};
    const int lvl_n[3] = {n, S, S};
    float *lvl_out[3] = {w.l1, w.l2, w.l3};
    const float *lvl_feat_in[3] = {nullptr, w.l1, w.l2};
    const int lvl_cin[3] = {0, 32, 64};
    const float *wf[3] = {nullptr, hw.wf2, hw.wf3};
    const SaScaleW *sw[3] = {hw.l1, hw.l2, hw.l3};
    const float *lin_w[3] = {hw.lin1_w, hw.lin2_w, hw.lin3_w};
    const float *lin_b[3] = {hw.lin1_b, hw.lin2_b, hw.lin3_b};
    for (int l = 0; l < 3; ++l) {
        const LevelCfg &cfg = kLevels[l];
        const int c1tot = cfg.c1[0] + cfg.c1[1];
        // projection of every input point's features through the feature columns of both scales' first conv
        RtRowGemm pg{};
        pg.rows = clouds * lvl_n[l];
        pg.nout = c1tot;
        if (l == 0) {
            pg.nseg = nseg;
            for (int i = 0; i < nseg; ++i) pg.seg[i] = segs[i];
            pg.cloud_bias = cloud_bias1;
            pg.rows_per_cloud = n;
        } else {
            pg.nseg = 1;
            pg.seg[0] = RtSeg{lvl_feat_in[l], lvl_cin[l], lvl_cin[l], wf[l], lvl_cin[l]};
            pg.rows_per_cloud = 1;
        }
        pg.act = RT_ACT_NONE;
        pg.y = w.proj;
        pg.ldy = c1tot;
        RT_TRY(rt_launch_rowgemm(pg, st));
        e->launches += 1;
        cudaStreamWaitEvent(st, L.ev_lvl[l], 0);   // FPS + ball query of this level
        const int pooled_c = (cfg.c3[0] ? cfg.c3[0] : cfg.c2[0]) + (cfg.c3[1] ? cfg.c3[1] : cfg.c2[1]);
        int coff = 0;
        for (int s = 0; s < 2; ++s) {
            const int ns = cfg.ns[s];
            const long long rows = (long long)clouds * S * ns;
            RtGatherCombine gc{};
            gc.clouds = clouds; gc.npts = S; gc.ns = ns; gc.c = cfg.c1[s];
            gc.y = w.proj; gc.ldy = c1tot; gc.yoff = s ? cfg.c1[0] : 0; gc.n_in = lvl_n[l];
            gc.idx = w.bq[l][s]; gc.xyz_in = lvl_xyz_in[l]; gc.xyz_c = w.xyz[l];
            gc.wx = sw[l][s].wx; gc.bias = sw[l][s].b1; gc.q = nullptr; gc.act = RT_ACT_RELU; gc.out = w.xa;
            RT_TRY(rt_launch_gather_combine(gc, st));
            RT_TRY(rt_launch_rowgemm(gemm1(rows, cfg.c2[s], w.xa, cfg.c1[s], cfg.c1[s], sw[l][s].w2, sw[l][s].b2, RT_ACT_RELU, w.xb, cfg.c2[s]), st));
            const float *last = w.xb;
            int clast = cfg.c2[s];
            e->launches += 2;
            if (cfg.c3[s]) {
                RT_TRY(rt_launch_rowgemm(gemm1(rows, cfg.c3[s], w.xb, cfg.c2[s], cfg.c2[s], sw[l][s].w3, sw[l][s].b3, RT_ACT_RELU, w.xa, cfg.c3[s]), st));
                last = w.xa;
                clast = cfg.c3[s];
                e->launches += 1;
            }
            RT_TRY(rt_launch_maxpool_rows(clouds * S, ns, clast, last, w.pooled, pooled_c, coff, st));
            e->launches += 1;
            coff += clast;
        }
        RT_TRY(rt_launch_rowgemm(gemm1((long long)clouds * S, cfg.lin_out, w.pooled, pooled_c, pooled_c, lin_w[l], lin_b[l], RT_ACT_NONE, lvl_out[l], cfg.lin_out), st));
        e->launches += 1;
    }
    // feature propagation (lib/pointnet2_modules.py:129-158): interpolate, concat skip, 1-layer SharedMLP
    cudaStreamWaitEvent(st, L.ev_nn, 0);
    {   // FP3: l2 <- l3
        RT_TRY(rt_launch_interp3(clouds, S, S, 64, w.l3, 64, w.nn_idx[0], w.nn_w[0], w.interp, 64, st));
        RtRowGemm g{};
        g.rows = clouds * S; g.nout = 128; g.nseg = 2;
        g.seg[0] = RtSeg{w.interp, 64, 64, hw.fp3_wi, 64};
        g.seg[1] = RtSeg{w.l2, 64, 64, hw.fp3_ws, 64};
        g.bias = hw.fp3_b; g.rows_per_cloud = 1; g.act = RT_ACT_RELU; g.y = w.l2p; g.ldy = 128;
        RT_TRY(rt_launch_rowgemm(g, st));
    }
    {   // FP2: l1 <- l2'
        RT_TRY(rt_launch_interp3(clouds, S, S, 128, w.l2p, 128, w.nn_idx[1], w.nn_w[1], w.interp, 128, st));
        RtRowGemm g{};
        g.rows = clouds * S; g.nout = 128; g.nseg = 2;
        g.seg[0] = RtSeg{w.interp, 128, 128, hw.fp2_wi, 128};
        g.seg[1] = RtSeg{w.l1, 32, 32, hw.fp2_ws, 32};
        g.bias = hw.fp2_b; g.rows_per_cloud = 1; g.act = RT_ACT_RELU; g.y = w.l1p; g.ldy = 128;
        RT_TRY(rt_launch_rowgemm(g, st));
    }
    {   // FP1: l0 <- l1' (no skip features)
        RT_TRY(rt_launch_interp3(clouds, n, S, 128, w.l1p, 128, w.nn_idx[2], w.nn_w[2], w.interp, 128, st));
        RT_TRY(rt_launch_rowgemm(gemm1((long long)clouds * n, 128, w.interp, 128, 128, hw.fp1_wi, hw.fp1_b, RT_ACT_RELU, out, 128), st));
    }
    e->launches += 6;
    return RT_OK;
}

}  // namespace

namespace {

RtMlpTc mlp_rows(long long rows, const float *x, int ldx, int k) {
    RtMlpTc m{};
    m.rows = rows;
    m.load_mode = RT_MLP_LOAD_ROWS;
    m.out_mode = RT_MLP_OUT_ROWS;
    m.nseg = 1;
    m.seg[0] = RtMlpSeg{x, ldx, k, nullptr, nullptr, 0, 0};
    m.rows_per_cloud = 1;
    return m;
}
void mlp_layer(RtMlpTc &m, const void *pack, const float *bias, int k, int n, int act) {
    m.layer[m.nlayers++] = RtMlpLayer{pack, bias, (k + 15) / 16 * 16, (n + 15) / 16 * 16, act};
}
void mlp_out(RtMlpTc &m, float *out, int ldo, int ooff, int n_out) {
    m.out = out; m.ldo = ldo; m.ooff = ooff; m.n_out = n_out;
}

// PNHead on the tensor cores: per level  projection GEMM -> 2 x [gather + conv chain + max-pool] -> linear
int run_head_tc(rt_engine *e, Lane &L, int lane_idx, const HeadW &hw, const HeadPacks &pk, Ws &w, int clouds, int n, const RtMlpSeg *segs,
                int nseg, const float *cloud_bias1, float *out, cudaStream_t st, bool self_knn_after_sa = false) {
    const int S = e->npoint;
    const float *lvl_xyz_in[3] = {w.xyz0, w.xyz[0], w.xyz[1]};
    const int lvl_n[3] = {n, S, S};
    float *lvl_out[3] = {w.l1, w.l2, w.l3};
    const SaScaleW *sw[3] = {hw.l1, hw.l2, hw.l3};
    const float *lin_b[3] = {hw.lin1_b, hw.lin2_b, hw.lin3_b};
    for (int l = 0; l < 3; ++l) {
        const LevelCfg &cfg = kLevels[l];
        const int c1tot = cfg.c1[0] + cfg.c1[1];
        if (l == 0) {
            // projection of the level-1 input features through the feature columns of both scales' first conv
            // (levels 2 and 3: fused with the previous level's Linear, below)
            RtMlpTc pg = mlp_rows((long long)clouds * n, nullptr, 0, 0);
            int k0 = 0;
            pg.nseg = nseg;
            for (int i = 0; i < nseg; ++i) { pg.seg[i] = segs[i]; k0 += (segs[i].k + 15) / 16 * 16; }
            pg.cloud_bias = cloud_bias1; pg.rows_per_cloud = n; pg.cloud_bias_ld = c1tot;
            mlp_layer(pg, pk.proj[0], nullptr, k0, c1tot, RT_ACT_NONE);
            mlp_out(pg, w.proj, c1tot, 0, c1tot);
            pg.status = w.status;
            RT_TRY(rt_launch_mlp_tc(pg, st));
            e->launches += 1;
        }
        cudaStreamWaitEvent(st, L.ev_lvl[l], 0);   // FPS + ball query of this level
        const int pooled_c = (cfg.c3[0] ? cfg.c3[0] : cfg.c2[0]) + (cfg.c3[1] ? cfg.c3[1] : cfg.c2[1]);
        int coff = 0;
        for (int s = 0; s < 2; ++s) {
            RtMlpTc m{};
            m.rows = (long long)clouds * S * cfg.ns[s];
            m.load_mode = RT_MLP_LOAD_GATHER; m.out_mode = RT_MLP_OUT_MAXPOOL;
            m.y = w.proj; m.ldy = c1tot; m.yoff = s ? cfg.c1[0] : 0; m.n_in = lvl_n[l];
            m.idx = w.bq[l][s]; m.xyz_in = lvl_xyz_in[l]; m.xyz_c = w.xyz[l]; m.wx = sw[l][s].wx; m.b1 = sw[l][s].b1;
            m.npts = S; m.ns = cfg.ns[s]; m.c1 = cfg.c1[s]; m.rows_per_cloud = 1;
            mlp_layer(m, pk.w2[l][s], sw[l][s].b2, cfg.c1[s], cfg.c2[s], RT_ACT_RELU);
            int clast = cfg.c2[s];
            if (cfg.c3[s]) { mlp_layer(m, pk.w3[s], sw[l][s].b3, cfg.c2[s], cfg.c3[s], RT_ACT_RELU); clast = cfg.c3[s]; }
            mlp_out(m, w.pooled, pooled_c, coff, clast);
            m.status = w.status;
            // bit 10: scales with one tensor-core layer gather from a shared-memory copy of the cloud's projections (sa_tc.cu)
            int rc = (e->flags & 1024) ? rt_launch_sa_tc(m, clouds, st) : RT_ERR_UNSUPPORTED;
            if (rc == RT_ERR_UNSUPPORTED) rc = rt_launch_mlp_tc(m, st);
            RT_TRY(rc);
            coff += clast;
        }
        // Linear after the max-pool; for levels 1 and 2 the next level's projection rides in the same launch as a second
        // layer (the Linear's fp32 output is kept too: the FP layers read it as skip features)
        RtMlpTc lg = mlp_rows((long long)clouds * S, w.pooled, pooled_c, pooled_c);
        mlp_layer(lg, pk.lin[l], lin_b[l], pooled_c, cfg.lin_out, RT_ACT_NONE);
        if (l < 2) {
            const int c1next = kLevels[l + 1].c1[0] + kLevels[l + 1].c1[1];
            mlp_layer(lg, pk.proj[l + 1], nullptr, cfg.lin_out, c1next, RT_ACT_NONE);
            lg.mid_out = lvl_out[l]; lg.mid_ldo = cfg.lin_out; lg.mid_layer = 0;
            mlp_out(lg, w.proj, c1next, 0, c1next);
        } else {
            mlp_out(lg, lvl_out[l], cfg.lin_out, 0, cfg.lin_out);
        }
        lg.status = w.status;
        RT_TRY(rt_launch_mlp_tc(lg, st));
        e->launches += 3;
        stage_mark(e, lane_idx, st);
        if (l == 2 && self_knn_after_sa) {   // late self-kNN: from here on the feature path runs small kernels
            cudaEventRecord(L.ev_sa3, st);
            cudaStreamWaitEvent(L.geo_stream[2], L.ev_sa3, 0);
            RT_TRY(rt_launch_knn_expanded(clouds / 2, n, n, kKnn, w.xyz0, w.xyz0, w.knn11, L.geo_stream[2]));
            cudaEventRecord(L.ev_knn11, L.geo_stream[2]);
        }
    }
    cudaStreamWaitEvent(st, L.ev_nn, 0);
    // the three-point interpolation is evaluated inside the GEMM's operand loader (no interp buffer, no extra launch)
    {   // FP3: l2 <- l3
        RtMlpTc m = mlp_rows((long long)clouds * S, nullptr, 0, 0);
        m.seg[0] = RtMlpSeg{w.l3, 64, 64, w.nn_idx[0], w.nn_w[0], S, S};
        m.nseg = 2; m.seg[1] = RtMlpSeg{w.l2, 64, 64, nullptr, nullptr, 0, 0};
        mlp_layer(m, pk.fp3, hw.fp3_b, 128, 128, RT_ACT_RELU);
        mlp_out(m, w.l2p, 128, 0, 128); m.status = w.status;
        RT_TRY(rt_launch_mlp_tc(m, st));
    }
    {   // FP2: l1 <- l2'
        RtMlpTc m = mlp_rows((long long)clouds * S, nullptr, 0, 0);
        m.seg[0] = RtMlpSeg{w.l2p, 128, 128, w.nn_idx[1], w.nn_w[1], S, S};
        m.nseg = 2; m.seg[1] = RtMlpSeg{w.l1, 32, 32, nullptr, nullptr, 0, 0};
        mlp_layer(m, pk.fp2, hw.fp2_b, 160, 128, RT_ACT_RELU);
        mlp_out(m, w.l1p, 128, 0, 128); m.status = w.status;
        RT_TRY(rt_launch_mlp_tc(m, st));
    }
    {   // FP1: l0 <- l1'
        RtMlpTc m = mlp_rows((long long)clouds * n, nullptr, 0, 0);
        m.seg[0] = RtMlpSeg{w.l1p, 128, 128, w.nn_idx[2], w.nn_w[2], n, S};
        mlp_layer(m, pk.fp1, hw.fp1_b, 128, 128, RT_ACT_RELU);
        mlp_out(m, out, 128, 0, 128); m.status = w.status;
        RT_TRY(rt_launch_mlp_tc(m, st));
    }
    e->launches += 3;
    return RT_OK;
}

// ---- weight packing at engine creation -----------------------------------------------------------
struct PackJob { void **dst; int n; int nseg; RtPackSeg seg[4]; };

void add_job(PackJob *jobs, int &nj, void **dst, int n, const float *w, int k) {
    jobs[nj] = PackJob{dst, n, 1, {RtPackSeg{w, k, k}}};
    ++nj;
}

int head_jobs(PackJob *jobs, int &nj, const HeadW &hw, HeadPacks &pk, bool mse) {
    // level-1 projection: both scales' feature columns, segment by segment (ft | local | cor; the cloud-constant
    // `glob` block is applied as a per-cloud bias)
    PackJob j{&pk.proj[0], 32, 1, {RtPackSeg{hw.wf_ft, 2, 2}}};
    if (mse) {
        j.nseg = 3;
        j.seg[1] = RtPackSeg{hw.wf_loc, 128, 128};
        j.seg[2] = RtPackSeg{hw.wf_cor, 256, 256};
    }
    jobs[nj++] = j;
    add_job(jobs, nj, &pk.proj[1], 64, hw.wf2, 32);
    add_job(jobs, nj, &pk.proj[2], 128, hw.wf3, 64);
    const SaScaleW *sw[3] = {hw.l1, hw.l2, hw.l3};
    for (int l = 0; l < 3; ++l)
        for (int s = 0; s < 2; ++s) {
            add_job(jobs, nj, &pk.w2[l][s], kLevels[l].c2[s], sw[l][s].w2, kLevels[l].c1[s]);
            if (l == 0) add_job(jobs, nj, &pk.w3[s], kLevels[0].c3[s], sw[0][s].w3, kLevels[0].c2[s]);
        }
    add_job(jobs, nj, &pk.lin[0], 32, hw.lin1_w, 64);
    add_job(jobs, nj, &pk.lin[1], 64, hw.lin2_w, 96);
    add_job(jobs, nj, &pk.lin[2], 64, hw.lin3_w, 128);
    jobs[nj++] = PackJob{&pk.fp3, 128, 2, {RtPackSeg{hw.fp3_wi, 64, 64}, RtPackSeg{hw.fp3_ws, 64, 64}}};
    jobs[nj++] = PackJob{&pk.fp2, 128, 2, {RtPackSeg{hw.fp2_wi, 128, 128}, RtPackSeg{hw.fp2_ws, 32, 32}}};
    add_job(jobs, nj, &pk.fp1, 128, hw.fp1_wi, 128);
    return RT_OK;
}

int build_packs(rt_engine *e) {
    PackJob jobs[64];
    int nj = 0;
    head_jobs(jobs, nj, e->w.pn, e->packs.pn, false);
    head_jobs(jobs, nj, e->w.mse, e->packs.mse, true);
    add_job(jobs, nj, &e->packs.p1, 256, e->w.cv.w1_l1, 128);
    add_job(jobs, nj, &e->packs.p2, 256, e->w.cv.w1_l2, 128);
    add_job(jobs, nj, &e->packs.cp[0], 128, e->w.cp.w1, 256);
    add_job(jobs, nj, &e->packs.cp[1], 64, e->w.cp.w2, 128);
    add_job(jobs, nj, &e->packs.cp[2], 32, e->w.cp.w3, 64);
    add_job(jobs, nj, &e->packs.fw[0], 128, e->w.fp.w1_l, 128);
    add_job(jobs, nj, &e->packs.fw[1], 64, e->w.fp.w2, 128);
    add_job(jobs, nj, &e->packs.fw[2], 32, e->w.fp.w3, 64);
    add_job(jobs, nj, &e->packs.fw[3], 3, e->w.fp.w4, 32);
    size_t total = 0;
    size_t offs[64];
    for (int i = 0; i < nj; ++i) {
        int kp = 0;
        for (int s = 0; s < jobs[i].nseg; ++s) kp += (jobs[i].seg[s].k + 15) / 16 * 16;
        const int np = (jobs[i].n + 15) / 16 * 16;
        offs[i] = total;
        total += ((size_t)4 * np * kp + 255) & ~(size_t)255;
    }
    cudaError_t err = cudaMalloc(&e->arena, total);
    if (err != cudaSuccess) {
        rt_set_error("engine_create: cudaMalloc(%zu) for weight packs: %s", total, cudaGetErrorString(err));
        return (int)err;
    }
    for (int i = 0; i < nj; ++i) {
        *jobs[i].dst = (char *)e->arena + offs[i];
        RT_TRY(rt_launch_pack_umma(*jobs[i].dst, jobs[i].n, (jobs[i].n + 15) / 16 * 16, jobs[i].seg, jobs[i].nseg, nullptr));
    }
    err = cudaDeviceSynchronize();
    if (err != cudaSuccess) {
        rt_set_error("engine_create: packing weights: %s", cudaGetErrorString(err));
        return (int)err;
    }
    return RT_OK;
}

}  // namespace

// v0 of the dense cost-volume MLP (two 256x256 layers with LeakyReLU over b*n*16 rows): two row GEMMs.
int rt_launch_costvol_mlp(int rows, const float *x1, const float *w2, const float *b2, const float *w3, const float *b3,
                          float *xa, float *xb, cudaStream_t st) {
    RT_TRY(rt_launch_rowgemm(gemm1(rows, 256, x1, 256, 256, w2, b2, RT_ACT_LEAKY01, xb, 256), st));
    RT_TRY(rt_launch_rowgemm(gemm1(rows, 256, xb, 256, 256, w3, b3, RT_ACT_LEAKY01, xa, 256), st));
    return RT_OK;
}

RT_API int rt_engine_num_lanes(const rt_engine *e, int b);
RT_API int rt_engine_num_weights(void) { return kNumWeights; }

RT_API int rt_engine_create(rt_engine **out, int npoint, const void *const *weights, int nweights) {
    RT_REQUIRE(out && weights, "engine_create: null argument");
    RT_REQUIRE(nweights == kNumWeights, "engine_create: expected %d weight pointers, got %d", kNumWeights, nweights);
    RT_REQUIRE(npoint >= 1, "engine_create: npoint=%d", npoint);
    rt_engine *e = new (std::nothrow) rt_engine();
    RT_REQUIRE(e, "engine_create: out of host memory");
    memcpy(&e->w, weights, sizeof(EngineW));
    e->npoint = npoint;
    if (const char *env = getenv("RT_ENGINE_FLAGS")) e->flags = atoi(env);   // A/B timing without touching the caller
    // pointers that may legitimately be null: pn_head has only the `ft` feature segment; levels 2-3 have no third conv
    const float *const *tab = reinterpret_cast<const float *const *>(&e->w);
    const int head = sizeof(HeadW) / sizeof(const float *);
    for (int i = 0; i < kNumWeights; ++i) {
        if (tab[i]) continue;
        const int in_head = i < 2 * head ? i % head : -1;
        const bool opt = (i < head && in_head >= 1 && in_head <= 3) ||              // pn.wf_loc/glob/cor
                         (in_head >= 0 && (in_head == 23 || in_head == 24 || in_head == 29 || in_head == 30 ||  // l2 w3/b3
                                           in_head == 38 || in_head == 39 || in_head == 44 || in_head == 45));  // l3 w3/b3
        if (!opt) {
            delete e;
            rt_set_error("engine_create: weight pointer %d is null", i);
            return RT_ERR_INVALID;
        }
    }
    cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming);
    // priorities: the latency-bound geometry chain (FPS, ball query, three_nn) first, the feature path next, the
    // throughput-bound cost-volume kNN and the API-layout output copies last -- they fill whatever is idle
    int prio_least = 0, prio_greatest = 0;
    cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest);
    const int prio_mid = prio_greatest < prio_least - 1 ? prio_greatest + 1 : prio_greatest;
    for (Lane &L : e->lanes) {
        cudaStreamCreateWithPriority(&L.main_stream, cudaStreamNonBlocking, prio_mid);
        cudaStreamCreateWithPriority(&L.geo_stream[0], cudaStreamNonBlocking, prio_greatest);
        cudaStreamCreateWithPriority(&L.geo_stream[1], cudaStreamNonBlocking, prio_greatest);
        cudaStreamCreateWithPriority(&L.geo_stream[2], cudaStreamNonBlocking, prio_least);
        cudaStreamCreateWithPriority(&L.aux_stream, cudaStreamNonBlocking, prio_least);
        cudaEvent_t *evs[] = {&L.ev_in, &L.ev_fps[0], &L.ev_fps[1], &L.ev_fps[2], &L.ev_lvl[0], &L.ev_lvl[1], &L.ev_lvl[2],
                              &L.ev_nn, &L.ev_knn, &L.ev_knn11, &L.ev_sa3, &L.ev_gh, &L.ev_feat, &L.ev_cor, &L.ev_prop, &L.ev_aux, &L.ev_done};
        for (cudaEvent_t *ev : evs) cudaEventCreateWithFlags(ev, cudaEventDisableTiming);
    }
    const int rc = build_packs(e);
    if (rc != RT_OK) {
        if (e->arena) cudaFree(e->arena);
        delete e;
        return rc;
    }
    *out = e;
    return RT_OK;
}

RT_API void rt_engine_destroy(rt_engine *e) {
    if (!e) return;
    if (e->arena) cudaFree(e->arena);
    if (e->ev_fork) cudaEventDestroy(e->ev_fork);
    for (cudaEvent_t ev : e->stage_ev)
        if (ev) cudaEventDestroy(ev);
    for (Lane &L : e->lanes) {
        if (L.main_stream) cudaStreamDestroy(L.main_stream);
        for (int g = 0; g < 3; ++g)
            if (L.geo_stream[g]) cudaStreamDestroy(L.geo_stream[g]);
        if (L.aux_stream) cudaStreamDestroy(L.aux_stream);
        cudaEvent_t evs[] = {L.ev_in, L.ev_fps[0], L.ev_fps[1], L.ev_fps[2], L.ev_lvl[0], L.ev_lvl[1], L.ev_lvl[2],
                             L.ev_nn, L.ev_knn, L.ev_knn11, L.ev_sa3, L.ev_gh, L.ev_feat, L.ev_cor, L.ev_prop, L.ev_aux, L.ev_done};
        for (cudaEvent_t ev : evs)
            if (ev) cudaEventDestroy(ev);
    }
    delete e;
}

RT_API int rt_engine_num_lanes(const rt_engine *e, int b) { return (e && (e->flags & 4) && b >= 8) ? 2 : 1; }

RT_API long long rt_engine_workspace_bytes(const rt_engine *e, int b, int n) {
    if (!e || b < 1 || n < 1) return -1;
    Ws w;
    if (rt_engine_num_lanes(e, b) == 1) {
        Carver c(nullptr);
        carve(c, w, b, n, e->npoint);
        return (long long)c.off + 256;
    }
    const int b0 = (b + 1) / 2;
    Carver c0(nullptr), c1(nullptr);
    carve(c0, w, b0, n, e->npoint);
    carve(c1, w, b - b0, n, e->npoint);
    return (long long)(((c0.off + 255) & ~(size_t)255) + c1.off + 512);
}

RT_API int rt_engine_set_profile_events(rt_engine *e, void *start, void *stop) {
    RT_REQUIRE(e, "engine_set_profile_events: null engine");
    e->prof_start = (cudaEvent_t)start;
    e->prof_stop = (cudaEvent_t)stop;
    return RT_OK;
}

RT_API long long rt_engine_launch_count(const rt_engine *e) { return e ? e->launches : -1; }

// Stage profile of the NEXT forwards (lane 0's feature-path stream): on = 1 creates timing events and records one at every
// stage boundary; rt_engine_stage_times synchronises and returns the milliseconds between consecutive marks of the last
// forward (ms[i] = duration of stage names[i+1]).  Returns the number of intervals written.
RT_API int rt_engine_stage_profile(rt_engine *e, int on) {
    RT_REQUIRE(e, "engine_stage_profile: null engine");
    if (on && !e->stage_ev[0])
        for (cudaEvent_t &ev : e->stage_ev) cudaEventCreate(&ev);
    e->stage_on = on != 0;
    e->stage_n = 0;
    return RT_OK;
}
RT_API int rt_engine_stage_times(rt_engine *e, float *ms, const char **names, int cap) {
    if (!e || !ms || e->stage_n < 2) return 0;
    cudaEventSynchronize(e->stage_ev[e->stage_n - 1]);
    int n = 0;
    for (int i = 1; i < e->stage_n && n < cap; ++i, ++n) {
        cudaEventElapsedTime(&ms[n], e->stage_ev[i - 1], e->stage_ev[i]);
        if (names) names[n] = kStageNames[i < 16 ? i : 15];
    }
    return n;
}

RT_API int rt_engine_set_flags(rt_engine *e, int flags) {
    RT_REQUIRE(e, "engine_set_flags: null engine");
    e->flags = flags;
    return RT_OK;
}

// blocking read of the device status word of the last forward: 0 = ok, bit 1 = an activation left the fp16x2
// range of the tensor-core cost volume (|x| >= 65000) -- results of that call must not be used
RT_API int rt_engine_last_status(rt_engine *e, int *status_out) {
    RT_REQUIRE(e && status_out, "engine_last_status: null argument");
    *status_out = 0;
    for (const int *p : e->last_status) {
        if (!p) continue;
        int v = 0;
        cudaError_t err = cudaMemcpy(&v, p, sizeof(int), cudaMemcpyDeviceToHost);
        if (err != cudaSuccess) {
            rt_set_error("engine_last_status: %s", cudaGetErrorString(err));
            return (int)err;
        }
        *status_out |= v;
    }
    return RT_OK;
}

// non-blocking companion of rt_engine_last_status: enqueues, on `stream`, the copy of the status words of the last forward into
// host_status[0..1] (pinned host memory; unused lanes are written as 0).  The caller reads them after an event / sync of its own.
RT_API int rt_engine_status_async(rt_engine *e, int *host_status, void *stream) {
    RT_REQUIRE(e && host_status, "engine_status_async: null argument");
    for (int l = 0; l < 2; ++l) {
        host_status[l] = 0;
        if (!e->last_status[l]) continue;
        const cudaError_t err = cudaMemcpyAsync(host_status + l, e->last_status[l], sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream);
        if (err != cudaSuccess) {
            rt_set_error("engine_status_async: %s", cudaGetErrorString(err));
            return (int)err;
        }
    }
    return RT_OK;
}

// One lane: b pairs of a call of b_total pairs.  All tensor pointers are already offset to the lane's first pair;
// h_in / h_out are (5, b_total, 128), so their layer stride is h_stride = b_total * 128.
static int forward_lane(rt_engine *e, Lane &L, int lane_idx, int b, size_t h_stride, int n, const float *pc1, const float *pc2,
                        const float *ft1, const float *ft2, const float *h_in, float *flow, float *h_out, float *cls, float *cor,
                        float *f1, float *f2, float *prop, int *knn12, int *knn11, void *workspace, cudaStream_t st,
                        const int *npts_host = nullptr) {
    Carver c(workspace);
    Ws w;
    carve(c, w, b, n, e->npoint);
    const int B2 = 2 * b;
    const size_t half3 = (size_t)b * n * 3, half2 = (size_t)b * n * 2, half128 = (size_t)b * n * 128;

    // inputs (B,3,N)/(B,2,N) channel-major -> row-major, pc1 clouds first then pc2
    RT_TRY(rt_launch_cm_to_rows(b, 3, n, pc1, w.xyz0, 3, 0, st));
    RT_TRY(rt_launch_cm_to_rows(b, 3, n, pc2, w.xyz0 + half3, 3, 0, st));
    RT_TRY(rt_launch_cm_to_rows(b, 2, n, ft1, w.ft0, 2, 0, st));
    RT_TRY(rt_launch_cm_to_rows(b, 2, n, ft2, w.ft0 + half2, 2, 0, st));
    e->launches += 4;
    // fork: geometry depends only on xyz, so it runs on its own stream and the feature path joins stage by stage
    if (lane_idx == 0) e->stage_n = 0;
    stage_mark(e, lane_idx, st);   // (the inputs->rows kernels precede this mark; "start" is recorded by the caller below)
    if (npts_host) {   // 2b counts (pc1 clouds, then pc2 clouds) of a padded variable-size batch
        const cudaError_t ce = cudaMemcpyAsync(w.npts, npts_host, sizeof(int) * 2 * (size_t)b, cudaMemcpyHostToDevice, st);
        if (ce != cudaSuccess) {
            rt_set_error("backbone_forward_varlen: %s", cudaGetErrorString(ce));
            return (int)ce;
        }
    }
    const int *cnt1 = npts_host ? w.npts : nullptr, *cnt2 = npts_host ? w.npts + b : nullptr;
    cudaEventRecord(L.ev_in, st);
    RT_TRY(run_geometry(e, L, w, b, n, npts_host));
    // hidden half of the GRU (depends on h_in only): off the critical path, on the output stream
    cudaStreamWaitEvent(L.aux_stream, L.ev_in, 0);
    RT_TRY(rt_launch_gru_hh(b, h_in, e->w.gru.whh, e->w.gru.bhh, h_stride, w.gru_gh, L.aux_stream));
    cudaEventRecord(L.ev_gh, L.aux_stream);

    // feature_extraction_head: pn_head over both clouds of every pair at once (track4d.py:102-106)
    const bool tc_mlp = (e->flags & 2) != 0;
    cudaMemsetAsync(w.status, 0, 64 * sizeof(int), st);
    e->last_status[lane_idx] = w.status;
    if (tc_mlp) {
        RtMlpSeg seg_ft{w.ft0, 2, 2, nullptr, nullptr, 0, 0};
        RT_TRY(run_head_tc(e, L, lane_idx, e->w.pn, e->packs.pn, w, B2, n, &seg_ft, 1, nullptr, w.feat, st, (e->flags & 256) != 0 && (e->flags & 2) != 0));
    } else {
        RtSeg seg_ft{w.ft0, 2, 2, e->w.pn.wf_ft, 2};
        RT_TRY(run_head(e, L, e->w.pn, w, B2, n, &seg_ft, 1, nullptr, w.feat, st));
    }
    stage_mark(e, lane_idx, st);   // pn_head FP done
    RT_TRY(rt_launch_cloud_max(B2, n, 128, w.feat, 128, w.gmax, st, cnt1));   // cnt1 = the 2b counts, pc1 clouds then pc2 clouds
    // API outputs pc1_features / pc2_features = cat(local, broadcast global) (track4d.py:89-95): off the critical path
    cudaStream_t aux = tc_mlp ? L.aux_stream : st;
    if (aux != st) {
        cudaEventRecord(L.ev_feat, st);
        cudaStreamWaitEvent(aux, L.ev_feat, 0);
    }
    RT_TRY(rt_launch_rows_to_cm(b, 128, n, w.feat, 128, 0, f1, 256, 0, aux));
    RT_TRY(rt_launch_rows_to_cm(b, 128, n, w.feat + half128, 128, 0, f2, 256, 0, aux));
    RT_TRY(rt_launch_broadcast_cm(b, 128, n, w.gmax, f1, 256, 128, aux));
    RT_TRY(rt_launch_broadcast_cm(b, 128, n, w.gmax + (size_t)b * 128, f2, 256, 128, aux));
    if (npts_host) {   // columns of padded points are defined: zero
        RT_TRY(rt_launch_mask_cm(b, 256, n, cnt1, f1, aux));
        RT_TRY(rt_launch_mask_cm(b, 256, n, cnt2, f2, aux));
    }
    e->launches += 5;

    // FeatureCorrelator (model_utils.py:193-250)
    const CvW &cv = e->w.cv;
    const float *x1 = w.xyz0, *x2 = w.xyz0 + half3;
    RT_TRY(rt_launch_cloud_matvec(b, 256, 128, cv.w1_g1, 128, w.gmax, 128, cv.b1, w.cb_a, st));
    RT_TRY(rt_launch_cloud_matvec(b, 256, 128, cv.w1_g2, 128, w.gmax + (size_t)b * 128, 128, nullptr, w.cb_b, st));
    if (tc_mlp) {
        RtMlpTc m = mlp_rows((long long)b * n, w.feat, 128, 128);
        mlp_layer(m, e->packs.p1, nullptr, 128, 256, RT_ACT_NONE);
        mlp_out(m, w.p1, 256, 0, 256);
        m.cloud_bias = w.cb_a; m.rows_per_cloud = n; m.cloud_bias_ld = 256; m.status = w.status;
        RT_TRY(rt_launch_mlp_tc(m, st));
        m = mlp_rows((long long)b * n, w.feat + half128, 128, 128);
        mlp_layer(m, e->packs.p2, nullptr, 128, 256, RT_ACT_NONE);
        mlp_out(m, w.p2, 256, 0, 256);
        m.cloud_bias = w.cb_b; m.rows_per_cloud = n; m.cloud_bias_ld = 256; m.status = w.status;
        RT_TRY(rt_launch_mlp_tc(m, st));
    } else {
        RtRowGemm g = gemm1((long long)b * n, 256, w.feat, 128, 128, cv.w1_l1, nullptr, RT_ACT_NONE, w.p1, 256);
        g.cloud_bias = w.cb_a; g.rows_per_cloud = n;
        RT_TRY(rt_launch_rowgemm(g, st));
        g = gemm1((long long)b * n, 256, w.feat + half128, 128, 128, cv.w1_l2, nullptr, RT_ACT_NONE, w.p2, 256);
        g.cloud_bias = w.cb_b; g.rows_per_cloud = n;
        RT_TRY(rt_launch_rowgemm(g, st));
    }
    stage_mark(e, lane_idx, st);   // global max + P1/P2 done
    cudaStreamWaitEvent(st, L.ev_knn, 0);   // join: everything the geometry stream produced is now ordered before `st`
    stage_mark(e, lane_idx, st);   // kNN joined
    if (e->flags & 1) {
        if (e->prof_start && lane_idx == 0) cudaEventRecord(e->prof_start, st);
        RT_TRY(rt_launch_costvol_tc(b * n, n, w.p1, w.p2, x1, x2, w.knn12, (e->flags & 64) ? w.perm1 : nullptr, cv.w1_x, cv.w23_pack, cv.wc1_pack, cv.b2, cv.b3,
                                    cv.wn1.bc, cv.wn1.wa, cv.wn1.ba, cv.wn1.wb, cv.wn1.bb, w.cost1, w.status, st));
        if (e->prof_stop && lane_idx == 0) cudaEventRecord(e->prof_stop, st);
    } else {
        RtGatherCombine gc{};
        gc.clouds = b; gc.npts = n; gc.ns = kKnn; gc.c = 256;
        gc.y = w.p2; gc.ldy = 256; gc.yoff = 0; gc.n_in = n; gc.idx = w.knn12; gc.xyz_in = x2; gc.xyz_c = x1;
        gc.wx = cv.w1_x; gc.bias = nullptr; gc.q = w.p1; gc.act = RT_ACT_LEAKY01; gc.out = w.xa;
        RT_TRY(rt_launch_gather_combine(gc, st));
        if (e->prof_start && lane_idx == 0) cudaEventRecord(e->prof_start, st);
        RT_TRY(rt_launch_costvol_mlp(b * n * kKnn, w.xa, cv.w2, cv.b2, cv.w3, cv.b3, w.xa, w.xb, st));
        if (e->prof_stop && lane_idx == 0) cudaEventRecord(e->prof_stop, st);
    }
    {
        RtWeightedSum ws{};
        ws.clouds = b; ws.npts = n; ws.ns = kKnn; ws.c = 256; ws.n_in = n; ws.gather_v = 0;
        ws.idx = w.knn12; ws.xyz_in = x2; ws.xyz_c = x1;
        ws.wa = cv.wn1.wa; ws.ba = cv.wn1.ba; ws.wb = cv.wn1.wb; ws.bb = cv.wn1.bb; ws.wc = cv.wn1.wc; ws.bc = cv.wn1.bc;
        ws.v = w.xa; ws.out = w.cost1;
        ws.perm = (e->flags & 64) ? w.perm1 : nullptr;
        if (!(e->flags & 1)) RT_TRY(rt_launch_weighted_sum(ws, st));
        stage_mark(e, lane_idx, st);   // cost volume done
        cudaStreamWaitEvent(st, L.ev_knn11, 0);
        ws.gather_v = 1; ws.idx = w.knn11; ws.xyz_in = x1;
        ws.wa = cv.wn2.wa; ws.ba = cv.wn2.ba; ws.wb = cv.wn2.wb; ws.bb = cv.wn2.bb; ws.wc = cv.wn2.wc; ws.bc = cv.wn2.bc;
        ws.v = w.cost1; ws.out = w.cor;
        RT_TRY(rt_launch_weighted_sum(ws, st));
    }
    stage_mark(e, lane_idx, st);   // patch-to-patch sum done
    if (aux != st) {
        cudaEventRecord(L.ev_cor, st);
        cudaStreamWaitEvent(aux, L.ev_cor, 0);
    }
    RT_TRY(rt_launch_rows_to_cm(b, 256, n, w.cor, 256, 0, cor, 256, 0, aux));
    if (npts_host) RT_TRY(rt_launch_mask_cm(b, 256, n, cnt1, cor, aux));
    e->launches += 10;

    // FlowDecoder (model_utils.py:281-305): cls head on the cost volume (independent of the mse head: output stream)
    const ClsW &cp = e->w.cp;
    const long long pts = (long long)b * n;
    if (tc_mlp) {
        RtMlpTc m = mlp_rows(pts, w.cor, 256, 256);
        mlp_layer(m, e->packs.cp[0], cp.b1, 256, 128, RT_ACT_RELU);
        mlp_layer(m, e->packs.cp[1], cp.b2, 128, 64, RT_ACT_RELU);
        mlp_layer(m, e->packs.cp[2], cp.b3, 64, 32, RT_ACT_RELU);
        mlp_out(m, w.h3, 32, 0, 32);
        m.status = w.status;
        RT_TRY(rt_launch_mlp_tc(m, aux));
    } else {
        RT_TRY(rt_launch_rowgemm(gemm1(pts, 128, w.cor, 256, 256, cp.w1, cp.b1, RT_ACT_RELU, w.h1, 128), st));
        RT_TRY(rt_launch_rowgemm(gemm1(pts, 64, w.h1, 128, 128, cp.w2, cp.b2, RT_ACT_RELU, w.h2, 64), st));
        RT_TRY(rt_launch_rowgemm(gemm1(pts, 32, w.h2, 64, 64, cp.w3, cp.b3, RT_ACT_RELU, w.h3, 32), st));
    }
    RT_TRY(rt_launch_cls_tail(pts, w.h3, cp.w4, cp.lin_w, cp.lin_b, cls, aux));
    if (npts_host) RT_TRY(rt_launch_mask_cm(b, 1, n, cnt1, cls, aux));
    // second PNHead over embeddings = cat(feature1, pc1_features, cor_features) on pc1's geometry
    const HeadW &mse = e->w.mse;
    RT_TRY(rt_launch_cloud_matvec(b, 32, 128, mse.wf_glob, 128, w.gmax, 128, nullptr, w.cb_a, st));
    if (tc_mlp) {
        RtMlpSeg segs[3] = {RtMlpSeg{w.ft0, 2, 2, nullptr, nullptr, 0, 0}, RtMlpSeg{w.feat, 128, 128, nullptr, nullptr, 0, 0},
                            RtMlpSeg{w.cor, 256, 256, nullptr, nullptr, 0, 0}};
        RT_TRY(run_head_tc(e, L, lane_idx, mse, e->packs.mse, w, b, n, segs, 3, w.cb_a, w.prop, st));
    } else {
        RtSeg segs[3] = {RtSeg{w.ft0, 2, 2, mse.wf_ft, 2}, RtSeg{w.feat, 128, 128, mse.wf_loc, 128},
                         RtSeg{w.cor, 256, 256, mse.wf_cor, 256}};
        RT_TRY(run_head(e, L, mse, w, b, n, segs, 3, w.cb_a, w.prop, st));
    }
    stage_mark(e, lane_idx, st);   // mse FP done
    if (aux != st) {
        cudaEventRecord(L.ev_prop, st);
        cudaStreamWaitEvent(aux, L.ev_prop, 0);
    }
    RT_TRY(rt_launch_rows_to_cm(b, 128, n, w.prop, 128, 0, prop, 128, 0, aux));
    if (npts_host) RT_TRY(rt_launch_mask_cm(b, 128, n, cnt1, prop, aux));
    if (aux != st) cudaEventRecord(L.ev_aux, aux);
    RT_TRY(rt_launch_cloud_max(b, n, 128, w.prop, 128, w.gprop, st, cnt1));
    cudaStreamWaitEvent(st, L.ev_gh, 0);
    RT_TRY(rt_launch_gru(b, w.gprop, h_in, e->w.gru.wih, e->w.gru.bih, w.gru_gh, h_out, h_stride, st));
    stage_mark(e, lane_idx, st);   // GRU done
    // FlowPredictor on cat(prop_features, broadcast GRU output)
    const FlowW &fp = e->w.fp;
    RT_TRY(rt_launch_cloud_matvec(b, 128, 128, fp.w1_g, 128, h_out + 4 * h_stride, 128, fp.b1, w.cb_b, st));
    if (tc_mlp) {
        RtMlpTc m = mlp_rows(pts, w.prop, 128, 128);
        mlp_layer(m, e->packs.fw[0], nullptr, 128, 128, RT_ACT_RELU);
        mlp_layer(m, e->packs.fw[1], fp.b2, 128, 64, RT_ACT_RELU);
        mlp_layer(m, e->packs.fw[2], fp.b3, 64, 32, RT_ACT_RELU);
        mlp_layer(m, e->packs.fw[3], nullptr, 32, 3, RT_ACT_NONE);
        mlp_out(m, w.flow_rows, 3, 0, 3);
        m.cloud_bias = w.cb_b; m.rows_per_cloud = n; m.cloud_bias_ld = 128; m.status = w.status;
        RT_TRY(rt_launch_mlp_tc(m, st));
    } else {
        RtRowGemm g = gemm1(pts, 128, w.prop, 128, 128, fp.w1_l, nullptr, RT_ACT_RELU, w.h1, 128);
        g.cloud_bias = w.cb_b; g.rows_per_cloud = n;
        RT_TRY(rt_launch_rowgemm(g, st));
        RT_TRY(rt_launch_rowgemm(gemm1(pts, 64, w.h1, 128, 128, fp.w2, fp.b2, RT_ACT_RELU, w.h2, 64), st));
        RT_TRY(rt_launch_rowgemm(gemm1(pts, 32, w.h2, 64, 64, fp.w3, fp.b3, RT_ACT_RELU, w.h3, 32), st));
        RT_TRY(rt_launch_rowgemm(gemm1(pts, 3, w.h3, 32, 32, fp.w4, nullptr, RT_ACT_NONE, w.flow_rows, 3), st));
    }
    RT_TRY(rt_launch_rows_to_cm(b, 3, n, w.flow_rows, 3, 0, flow, 3, 0, st));
    if (npts_host) RT_TRY(rt_launch_mask_cm(b, 3, n, cnt1, flow, st));
    e->launches += 11;
    if (aux != st) cudaStreamWaitEvent(st, L.ev_aux, 0);   // join the output stream
    // fp16-range guard fired anywhere in this lane's step -> its results become NaN (never silently saturated values)
    if (e->flags & 3) {
        RT_TRY(rt_launch_poison_on_status(w.status, flow, (long long)b * 3 * n, cls, (long long)b * n, h_out, b, h_stride, st));
        e->launches += 1;
    }
    stage_mark(e, lane_idx, st);   // flow head + joined outputs
    if (knn12) cudaMemcpyAsync(knn12, w.knn12, sizeof(int) * (size_t)b * n * kKnn, cudaMemcpyDeviceToDevice, st);
    if (knn11) cudaMemcpyAsync(knn11, w.knn11, sizeof(int) * (size_t)b * n * kKnn, cudaMemcpyDeviceToDevice, st);
    return rt_check_launch("backbone_forward");
}

static int backbone_forward_impl(rt_engine *e, int b, int n, const float *pc1, const float *pc2, const float *ft1,
                                 const float *ft2, const float *h_in, float *flow, float *h_out, float *cls, float *cor,
                                 float *f1, float *f2, float *prop, int *knn12, int *knn11, void *workspace,
                                 long long workspace_bytes, void *stream, const int *npts1, const int *npts2) {
    RT_REQUIRE(e && pc1 && pc2 && ft1 && ft2 && h_in && flow && h_out && cls && cor && f1 && f2 && prop && workspace,
               "backbone_forward: null argument");
    RT_REQUIRE(b >= 1 && n >= 1, "backbone_forward: b=%d n=%d", b, n);
    RT_REQUIRE(2 * b <= 65535, "backbone_forward: batch > 32767");
    RT_REQUIRE(workspace_bytes >= rt_engine_workspace_bytes(e, b, n), "backbone_forward: workspace too small");
    RT_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "backbone_forward: workspace must be 256-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t hs = (size_t)b * 128;
    e->last_status[0] = e->last_status[1] = nullptr;
    // variable-size batch: per lane, the counts of its pc1 clouds followed by those of its pc2 clouds (host memory that
    // stays alive until the lane's cudaMemcpyAsync has staged it)
    const bool varlen = npts1 != nullptr;
    auto lane_counts = [&](int lane, int o, int bl) -> const int * {
        if (!varlen) return nullptr;
        std::vector<int> &v = e->npts_host[lane];
        v.assign(npts1 + o, npts1 + o + bl);
        v.insert(v.end(), npts2 + o, npts2 + o + bl);
        return v.data();
    };
    if (varlen)
        for (int i = 0; i < b; ++i)
            RT_REQUIRE(npts1[i] >= kKnn && npts1[i] <= n && npts2[i] >= kKnn && npts2[i] <= n,
                       "backbone_forward_varlen: pair %d has %d / %d points (need %d..%d)", i, npts1[i], npts2[i], kKnn, n);
    if (rt_engine_num_lanes(e, b) == 1) {
        if (!(e->flags & 32))
            return forward_lane(e, e->lanes[0], 0, b, hs, n, pc1, pc2, ft1, ft2, h_in, flow, h_out, cls, cor, f1, f2, prop, knn12,
                                knn11, workspace, st, lane_counts(0, 0, b));
        Lane &L = e->lanes[0];
        cudaEventRecord(e->ev_fork, st);
        cudaStreamWaitEvent(L.main_stream, e->ev_fork, 0);
        RT_TRY(forward_lane(e, L, 0, b, hs, n, pc1, pc2, ft1, ft2, h_in, flow, h_out, cls, cor, f1, f2, prop, knn12, knn11,
                            workspace, L.main_stream, lane_counts(0, 0, b)));
        cudaEventRecord(L.ev_done, L.main_stream);
        cudaStreamWaitEvent(st, L.ev_done, 0);
        return rt_check_launch("backbone_forward");
    }
    // two lanes: pairs [0, b0) and [b0, b) run concurrently on their own stream sets, forked from / joined to `st`
    const int b0 = (b + 1) / 2;
    Carver c0(nullptr);
    Ws w0;
    carve(c0, w0, b0, n, e->npoint);
    const size_t ws0 = (c0.off + 255) & ~(size_t)255;
    cudaEventRecord(e->ev_fork, st);
    for (int l = 0; l < 2; ++l) {
        Lane &L = e->lanes[l];
        const int bl = l ? b - b0 : b0;
        const size_t o = l ? b0 : 0;   // first pair of the lane
        cudaStreamWaitEvent(L.main_stream, e->ev_fork, 0);
        RT_TRY(forward_lane(e, L, l, bl, hs, n, pc1 + o * 3 * n, pc2 + o * 3 * n, ft1 + o * 2 * n, ft2 + o * 2 * n, h_in + o * 128,
                            flow + o * 3 * n, h_out + o * 128, cls + o * n, cor + o * 256 * n, f1 + o * 256 * n, f2 + o * 256 * n,
                            prop + o * 128 * n, knn12 ? knn12 + o * n * kKnn : nullptr, knn11 ? knn11 + o * n * kKnn : nullptr,
                            (char *)workspace + (l ? ws0 : 0), L.main_stream, lane_counts(l, (int)o, bl)));
        cudaEventRecord(L.ev_done, L.main_stream);
        cudaStreamWaitEvent(st, L.ev_done, 0);
    }
    return rt_check_launch("backbone_forward");
}

RT_API int rt_backbone_forward(rt_engine *e, int b, int n, const float *pc1, const float *pc2, const float *ft1,
                               const float *ft2, const float *h_in, float *flow, float *h_out, float *cls, float *cor,
                               float *f1, float *f2, float *prop, int *knn12, int *knn11, void *workspace,
                               long long workspace_bytes, void *stream) {
    return backbone_forward_impl(e, b, n, pc1, pc2, ft1, ft2, h_in, flow, h_out, cls, cor, f1, f2, prop, knn12, knn11, workspace,
                                 workspace_bytes, stream, nullptr, nullptr);
}

// Variable-size batch (real radar frames have 240..350 points each: src/dataset_classes/track_vod_3d.py:49-122): the clouds
// are padded to n columns, npts1 / npts2 (HOST arrays of b ints) give the valid points of every pc1 / pc2 cloud.  Every pair is
// computed exactly as rt_backbone_forward computes it alone at its own size (FPS tie-breaks included: one FPS launch per size
// class); output columns of padded points are zero.
RT_API int rt_backbone_forward_varlen(rt_engine *e, int b, int n, const int *npts1, const int *npts2, const float *pc1, const float *pc2,
                                      const float *ft1, const float *ft2, const float *h_in, float *flow, float *h_out, float *cls,
                                      float *cor, float *f1, float *f2, float *prop, int *knn12, int *knn11, void *workspace,
                                      long long workspace_bytes, void *stream) {
    RT_REQUIRE(npts1 && npts2, "backbone_forward_varlen: null point counts");
    return backbone_forward_impl(e, b, n, pc1, pc2, ft1, ft2, h_in, flow, h_out, cls, cor, f1, f2, prop, knn12, knn11, workspace,
                                 workspace_bytes, stream, npts1, npts2);
}
