// Cost-volume core on the 5th-generation tensor cores (tcgen05 / TMEM / bulk-copy engine), sm_100a.
//
// Replaces, for FeatureCorrelator.forward (reference: src/utils/model_utils/model_utils.py:216-236), the chain
//     gather f2 / xyz of the 16 neighbours -> concat 515 channels -> conv 515->256 -> conv 256->256 -> conv 256->256
//     (LeakyReLU 0.1 each) -> WeightNet(dxyz) weighting -> sum over the 16 neighbours
// by ONE persistent kernel that never materialises a (B,515,16,N) / (B,256,16,N) tensor:
//   tile      = 128 rows = 8 query points x 16 neighbours; the row index is the TMEM lane;
//   layer 1   = P1[point] + P2[neighbour] + Wx.dxyz  (P1/P2 = per-POINT projections computed once by a plain GEMM),
//               evaluated by the worker warps straight into TMEM as the A operand of layer 2;
//   layer 2/3 = tcgen05.mma kind::f16, M128 N256 K16, accumulator in TMEM columns [0,256); A operand read from
//               TMEM columns [256,512); B operand (weights) streamed by the bulk-copy engine (UBLKCP) through a
//               4-stage shared-memory ring in the canonical K-major core-matrix layout (pre-packed on the host);
//   precision = every fp32 operand x is split x = hi + 2^-11 lo' (two fp16 planes, weights pre-scaled by 2^10) and each
//               product is evaluated as (lo'*hi + hi*lo') * 2^-11 + hi*hi with fp32 accumulation, corrections first
//               (tools/tc_precision.cu: the rms error of an fp32 FMA chain x 1.7) at 1/3 of the fp16 tensor rate;
//   WeightNet = its last layer (8 -> 256) is a K=16 MMA whose accumulator lands on the dead A columns;
//   epilogue  = bias + LeakyReLU, weight, and a 16-lane butterfly that leaves the neighbour sum in registers.
// TMEM (512 columns) is exactly full: 256 accumulator + 128 A_hi + 128 A_lo, so a tile's MMA and epilogue phases
// alternate (DESIGN.md discusses the resulting tensor-pipe ceiling and the cta_group::2 follow-up).
#include <cuda_fp16.h>
#include <stdlib.h>

#include "engine_kernels.cuh"

namespace {

constexpr int CT_ROWS = 128, CT_PTS = 8, CT_NS = 16, CT_C = 256;
constexpr int CT_KC = 32;                          // K per streamed weight chunk

constexpr int CT_STAGES_MAX = 6;   // ring depth is a template parameter (4 or 6 stages of 32 KB); the carve-up reserves room for 6
constexpr int CT_GROUPS = CT_C / CT_KC;            // K groups per layer: the unit of the A-operand hand-off (8)
constexpr int CT_PLANE_BYTES = CT_C * CT_KC * 2;   // one fp16 plane of a chunk: 16 KB
constexpr int CT_STAGE_BYTES = 2 * CT_PLANE_BYTES; // hi + lo
constexpr int CT_WC_PLANE = CT_C * 16 * 2;         // WeightNet last layer, K padded 8 -> 16: 8 KB
constexpr int CT_AW_PLANE = CT_ROWS * 16 * 2;      // its A operand: 4 KB
constexpr int CT_WORKER_WARPS = 8;
constexpr int CT_THREADS = 32 * (CT_WORKER_WARPS + 2);
constexpr float CT_WINV = 1.0f / 1024.0f;          // weights are packed as 2^10 * W

// shared memory carve-up (bytes)
constexpr int SM_STAGES = 0;
constexpr int SM_WC = SM_STAGES + CT_STAGES_MAX * CT_STAGE_BYTES;
constexpr int SM_AW = SM_WC + 2 * CT_WC_PLANE;
constexpr int SM_B2 = SM_AW + 2 * CT_AW_PLANE;
constexpr int SM_B3 = SM_B2 + CT_C * 4;
constexpr int SM_BC = SM_B3 + CT_C * 4;
constexpr int SM_WX = SM_BC + CT_C * 4;            // [3][256]
constexpr int SM_WN = SM_WX + 3 * CT_C * 4;        // wa(24) ba(8) wb(64) bb(8)
constexpr int SM_BAR = SM_WN + 128 * 4;
constexpr int SM_TOTAL = SM_BAR + 32 * 8;

struct CostVolTcArgs {
    int total_pts, n;
    const float *p1, *p2, *xyz1, *xyz2;
    const int *knn;
    const int *perm;   // processing order (slot -> point), see rt_launch_morton_perm; null = identity
    const float *w1x;
    const __half *wpack, *wcpack;
    const float *b2, *b3, *bc, *wa, *ba, *wb, *bb;
    float *out;
    int *status;
};

// ---- PTX wrappers ------------------------------------------------------------------------------------
__device__ __forceinline__ void ct_mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = rt_smem_u32(bar);
    for (uint32_t it = 0;; ++it) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
        if (ok) return;
        __nanosleep(it < 4 ? 32 : 96);  // back off: a spinning warp steals issue slots from the warps that have work
        if (it > (1u << 22)) __trap();  // a protocol bug must fail loudly, never hang the device
    }
}
__device__ __forceinline__ void ct_tmem_alloc(uint32_t *slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(rt_smem_u32(slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void ct_tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void ct_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void ct_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void ct_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(rt_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void ct_mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void ct_mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
// D = A*B + D * 2^-11 (scale-input-d): folds the 2^11 of the scaled lo planes back when the main products start
__device__ __forceinline__ void ct_mma_ts_rescale(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p, 11;\n\t}"
                 ::"r"(d), "r"(a), "l"(b), "r"(idesc) : "memory");
}
__device__ __forceinline__ void ct_mma_ss_rescale(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p, 11;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc) : "memory");
}
__device__ __forceinline__ void ct_ld32(uint32_t taddr, uint32_t *r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void ct_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void ct_st8(uint32_t taddr, const uint32_t *r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void ct_st16(uint32_t taddr, const uint32_t *r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                   "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void ct_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// K-major, no-swizzle shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp SmemDescriptor):
// LBO = byte distance between core matrices adjacent in K, SBO = between 8-row groups (verified by tools/tc_probe.cu)
__device__ __forceinline__ uint64_t ct_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3fff) << 32) | ((uint64_t)1 << 46);
}
// K groups are consumed in the order both worker halves produce them: 0,4,1,5,... (half 0 owns groups 0-3, half 1 4-7)
__device__ __forceinline__ int ct_group_at(int i) { return (i >> 1) + 4 * (i & 1); }
// F16 x F16 -> F32, A and B K-major, M = 128, N = 256
constexpr uint32_t CT_IDESC = (1u << 4) | ((uint32_t)(CT_C >> 3) << 17) | ((uint32_t)(CT_ROWS >> 4) << 24);

// LeakyReLU(0.1) on a packed pair: max(v, 0.1 v) -- two instructions per pair, no select
__device__ __forceinline__ float2 leaky01x2(float2 v) {
    const float2 m = rt_fmul2(v, make_float2(0.1f, 0.1f));
    return make_float2(fmaxf(v.x, m.x), fmaxf(v.y, m.y));
}

// x = hi + 2^-11 * lo' with two fp16 (x0 in the low half: K even)
__device__ __forceinline__ void split2(float x0, float x1, uint32_t &hi, uint32_t &lo, float &amax) {
    amax = fmaxf(amax, fmaxf(fabsf(x0), fabsf(x1)));
    const __half2 h = __floats2half2_rn(x0, x1);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn((x0 - hf.x) * 2048.0f, (x1 - hf.y) * 2048.0f);   // lo plane stored as 2^11 * lo: normal whenever hi is
    hi = *reinterpret_cast<const uint32_t *>(&h);
    lo = *reinterpret_cast<const uint32_t *>(&l);
}

// Per-row operands of a tile's layer 1.  They are fetched one tile ahead (head) and one 16-channel chunk ahead (body)
// so the index -> xyz -> P1/P2 load chain hides under the MMA phases instead of stalling the prologue.
struct CtRow {
    int p, pc;
    bool valid;
    size_t g2;
    float dx, dy, dz;
    float4 u[4], v[4];   // chunk of P2[neighbour] and P1[point]
};
__device__ __forceinline__ void ct_issue_chunk(const CostVolTcArgs &a, const CtRow &r, int c0, float4 *u, float4 *v) {
    const float4 *p1r = reinterpret_cast<const float4 *>(a.p1 + (size_t)r.pc * CT_C + c0);
    const float4 *p2r = reinterpret_cast<const float4 *>(a.p2 + r.g2 * CT_C + c0);
    // 256-bit loads: every lane reads a different neighbour row, so the L1 data pipe pays per instruction, not per byte
    rt_ldg256(reinterpret_cast<const float *>(p2r), u[0], u[1]);
    rt_ldg256(reinterpret_cast<const float *>(p2r + 2), u[2], u[3]);
    rt_ldg256(reinterpret_cast<const float *>(p1r), v[0], v[1]);
    rt_ldg256(reinterpret_cast<const float *>(p1r + 2), v[2], v[3]);
}
__device__ __forceinline__ void ct_issue_row(const CostVolTcArgs &a, int tile, int row, int cbeg, CtRow &r) {
    const int slot = tile * CT_PTS + (row >> 4);
    r.valid = slot < a.total_pts;
    r.p = r.valid ? (a.perm ? __ldg(a.perm + slot) : slot) : a.total_pts;
    r.pc = r.valid ? r.p : a.total_pts - 1;
    const int cloud = r.pc / a.n;
    const int nbr = __ldg(a.knn + (size_t)r.pc * CT_NS + (row & 15));
    r.g2 = (size_t)cloud * a.n + nbr;
    r.dx = __ldg(a.xyz2 + r.g2 * 3 + 0) - __ldg(a.xyz1 + (size_t)r.pc * 3 + 0);
    r.dy = __ldg(a.xyz2 + r.g2 * 3 + 1) - __ldg(a.xyz1 + (size_t)r.pc * 3 + 1);
    r.dz = __ldg(a.xyz2 + r.g2 * 3 + 2) - __ldg(a.xyz1 + (size_t)r.pc * 3 + 2);
    ct_issue_chunk(a, r, cbeg, r.u, r.v);
}

template <int CT_STAGES>
__global__ void __launch_bounds__(CT_THREADS, 1) costvol_tc_v2_kernel(CostVolTcArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *s_stage = smem + SM_STAGES;
    uint8_t *s_wc = smem + SM_WC;
    uint8_t *s_aw = smem + SM_AW;
    float *s_b2 = reinterpret_cast<float *>(smem + SM_B2);
    float *s_b3 = reinterpret_cast<float *>(smem + SM_B3);
    float *s_bc = reinterpret_cast<float *>(smem + SM_BC);
    float *s_wx = reinterpret_cast<float *>(smem + SM_WX);
    float *s_wn = reinterpret_cast<float *>(smem + SM_WN);
    uint64_t *bar_full = reinterpret_cast<uint64_t *>(smem + SM_BAR);
    uint64_t *bar_empty = bar_full + CT_STAGES_MAX;
    uint64_t *bar_a = bar_empty + CT_STAGES_MAX;   // [CT_GROUPS]: A columns of one 32-wide K group are in TMEM
    uint64_t *bar_d = bar_a + CT_GROUPS;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar_d + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntiles = (a.total_pts + CT_PTS - 1) / CT_PTS;

    // ---- one-time setup --------------------------------------------------------------------------
    if (threadIdx.x == 0) {
        for (int s = 0; s < CT_STAGES; ++s) {
            rt_mbar_init(&bar_full[s], 1);
            rt_mbar_init(&bar_empty[s], 1);
        }
        for (int g = 0; g < CT_GROUPS; ++g) rt_mbar_init(&bar_a[g], CT_WORKER_WARPS / 2);  // the 4 warps owning that K half
        rt_mbar_init(bar_d, 1);
        rt_fence_mbar_init();
    }
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(a.wcpack);
        uint4 *dst = reinterpret_cast<uint4 *>(s_wc);
        for (int i = threadIdx.x; i < 2 * CT_WC_PLANE / 16; i += CT_THREADS) dst[i] = __ldg(src + i);
        uint4 *aw = reinterpret_cast<uint4 *>(s_aw);
        for (int i = threadIdx.x; i < 2 * CT_AW_PLANE / 16; i += CT_THREADS) aw[i] = make_uint4(0, 0, 0, 0);
        for (int i = threadIdx.x; i < CT_C; i += CT_THREADS) {
            s_b2[i] = __ldg(a.b2 + i);
            s_b3[i] = __ldg(a.b3 + i);
            s_bc[i] = __ldg(a.bc + i);
            s_wx[i] = __ldg(a.w1x + i * 3 + 0);
            s_wx[CT_C + i] = __ldg(a.w1x + i * 3 + 1);
            s_wx[2 * CT_C + i] = __ldg(a.w1x + i * 3 + 2);
        }
        if (threadIdx.x < 24) s_wn[threadIdx.x] = __ldg(a.wa + threadIdx.x);
        if (threadIdx.x < 8) s_wn[24 + threadIdx.x] = __ldg(a.ba + threadIdx.x);
        if (threadIdx.x < 64) s_wn[32 + threadIdx.x] = __ldg(a.wb + threadIdx.x);
        if (threadIdx.x < 8) s_wn[96 + threadIdx.x] = __ldg(a.bb + threadIdx.x);
    }
    rt_fence_proxy_async();  // generic-proxy writes of s_wc / s_aw -> visible to the tensor-core (async) proxy
    if (warp == CT_WORKER_WARPS + 1) ct_tmem_alloc(tmem_slot, 512);
    ct_fence_before();
    __syncthreads();
    ct_fence_after();
    const uint32_t tm = *tmem_slot;
    const uint32_t tD = tm, tAhi = tm + 256, tAlo = tm + 384, tW = tm + 256;

    if (warp == CT_WORKER_WARPS) {
        // ===== weight producer: one elected lane feeds the ring with the bulk-copy engine =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const uint8_t *wp = reinterpret_cast<const uint8_t *>(a.wpack);
                for (int layer = 0; layer < 2; ++layer) {
                    // correction pass: [hi, lo] planes of one K chunk per stage, in the order the A groups arrive
                    for (int c = 0; c < CT_GROUPS; ++c) {
                        ct_mbar_wait(&bar_empty[stage], phase ^ 1);
                        rt_mbar_expect_tx(&bar_full[stage], CT_STAGE_BYTES);
                        const int chunk = layer * CT_GROUPS + ct_group_at(c);
                        rt_bulk_g2s(s_stage + stage * CT_STAGE_BYTES, wp + (size_t)chunk * CT_STAGE_BYTES, CT_STAGE_BYTES, &bar_full[stage]);
                        if (++stage == CT_STAGES) { stage = 0; phase ^= 1; }
                    }
                    // main pass: the hi planes of two consecutive K chunks per stage
                    for (int c = 0; c < CT_GROUPS / 2; ++c) {
                        ct_mbar_wait(&bar_empty[stage], phase ^ 1);
                        rt_mbar_expect_tx(&bar_full[stage], CT_STAGE_BYTES);
                        for (int pl = 0; pl < 2; ++pl) {
                            const int chunk = layer * CT_GROUPS + 2 * c + pl;
                            rt_bulk_g2s(s_stage + stage * CT_STAGE_BYTES + pl * CT_PLANE_BYTES, wp + (size_t)chunk * CT_STAGE_BYTES,
                                        CT_PLANE_BYTES, &bar_full[stage]);
                        }
                        if (++stage == CT_STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == CT_WORKER_WARPS + 1) {
        // ===== MMA issuer: a single thread drives the tensor core =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0, a_phase = 0;
            const uint32_t aw_hi = rt_smem_u32(s_aw), aw_lo = aw_hi + CT_AW_PLANE;
            const uint32_t wc_hi = rt_smem_u32(s_wc), wc_lo = wc_hi + CT_WC_PLANE;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                for (int layer = 0; layer < 2; ++layer) {
                    for (int c = 0; c < CT_GROUPS; ++c) {
                        const int g = ct_group_at(c);
                        ct_mbar_wait(&bar_a[g], a_phase);   // this K group of the A operand has landed in TMEM:
                        ct_fence_after();                   // the MMAs start while the workers still produce later groups
                        ct_mbar_wait(&bar_full[stage], phase);
                        ct_fence_after();
                        const uint32_t base = rt_smem_u32(s_stage + stage * CT_STAGE_BYTES);
                        // Correction products first (lo*hi + hi*lo over all of K, carrying the 2^11 of the scaled lo planes),
                        // main products (hi*hi) last, the first of them rescaling the partial sum by 2^-11: the tensor core
                        // truncates the fp32 accumulator after every k-step, so the number of accumulation steps taken at
                        // full magnitude sets the error (tools/tc_precision.cu: 3x smaller rms than interleaving).
#pragma unroll
                        for (int j = 0; j < CT_KC / 16; ++j) {
                            const int kk = g * (CT_KC / 16) + j;
                            // chunk plane = [kc = K/8][row group = 32][8 rows][8 halfs]: kc stride 4096 B, row group 128 B
                            const uint64_t bhi = ct_desc(base + j * 8192, 4096, 128);
                            const uint64_t blo = ct_desc(base + CT_PLANE_BYTES + j * 8192, 4096, 128);
                            ct_mma_ts(tD, tAlo + 8 * kk, bhi, CT_IDESC, (c | j) > 0);
                            ct_mma_ts(tD, tAhi + 8 * kk, blo, CT_IDESC, 1);
                        }
                        ct_commit(&bar_empty[stage]);  // frees the ring slot when these MMAs have read it
                        if (++stage == CT_STAGES) { stage = 0; phase ^= 1; }
                    }
                    for (int c = 0; c < CT_GROUPS / 2; ++c) {
                        ct_mbar_wait(&bar_full[stage], phase);
                        ct_fence_after();
                        const uint32_t base = rt_smem_u32(s_stage + stage * CT_STAGE_BYTES);
#pragma unroll
                        for (int j = 0; j < 2 * (CT_KC / 16); ++j) {
                            const int kk = 2 * c * (CT_KC / 16) + j;   // the stage holds the hi planes of K chunks 2c, 2c+1 back to back
                            if (kk == 0) ct_mma_ts_rescale(tD, tAhi, ct_desc(base, 4096, 128), CT_IDESC);   // D = A*B + D * 2^-11
                            else ct_mma_ts(tD, tAhi + 8 * kk, ct_desc(base + j * 8192, 4096, 128), CT_IDESC, 1);
                        }
                        ct_commit(&bar_empty[stage]);
                        if (++stage == CT_STAGES) { stage = 0; phase ^= 1; }
                    }
                    if (layer == 1) {
                        // WeightNet last layer: [128 x 16] . [256 x 16]^T -> columns [256,512) (the A planes are dead now;
                        // tcgen05.mma executes in issue order)
                        ct_mma_ss(tW, ct_desc(aw_lo, 2048, 128), ct_desc(wc_hi, 4096, 128), CT_IDESC, 0);
                        ct_mma_ss(tW, ct_desc(aw_hi, 2048, 128), ct_desc(wc_lo, 4096, 128), CT_IDESC, 1);
                        ct_mma_ss_rescale(tW, ct_desc(aw_hi, 2048, 128), ct_desc(wc_hi, 4096, 128), CT_IDESC);
                    }
                    ct_commit(bar_d);
                    a_phase ^= 1;
                }
            }
        }
        __syncwarp();
    } else {
        // ===== worker warps: layer 1 (gather + combine) -> TMEM, mid epilogue, final epilogue =====
        const int q = warp & 3, hlf = warp >> 2;
        const int row = 32 * q + lane;
        const uint32_t lane_base = (uint32_t)(32 * q) << 16;
        const int cbeg = 128 * hlf;
        uint32_t d_phase = 0;
        float amax = 0.0f;
        CtRow cur;
        if ((int)blockIdx.x < ntiles) ct_issue_row(a, blockIdx.x, row, cbeg, cur);
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            // ---------- layer 1 ----------  (row operands + first chunk were issued during the previous tile)
            const int p = cur.p;
            const bool valid = cur.valid;
            const float dx = cur.dx, dy = cur.dy, dz = cur.dz;
            const float2 dx2 = make_float2(dx, dx), dy2 = make_float2(dy, dy), dz2 = make_float2(dz, dz);
            if (hlf == 0) {
                // WeightNet trunk 3 -> 8 -> 8 (ReLU), result = A operand (K = 8, padded to 16) of the last-layer MMA
                float h1[8], h2[8];
#pragma unroll
                for (int o = 0; o < 8; ++o) {
                    float v = s_wn[24 + o];
                    v = fmaf(s_wn[o * 3 + 0], dx, v);
                    v = fmaf(s_wn[o * 3 + 1], dy, v);
                    v = fmaf(s_wn[o * 3 + 2], dz, v);
                    h1[o] = fmaxf(v, 0.0f);
                }
#pragma unroll
                for (int o = 0; o < 8; ++o) {
                    float v = s_wn[96 + o];
#pragma unroll
                    for (int k = 0; k < 8; ++k) v = fmaf(s_wn[32 + o * 8 + k], h1[k], v);
                    h2[o] = fmaxf(v, 0.0f);
                }
                uint32_t hi[4], lo[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) split2(h2[2 * k], h2[2 * k + 1], hi[k], lo[k], amax);
                const int off = (row >> 3) * 128 + (row & 7) * 16;  // core-matrix row of K block 0
                *reinterpret_cast<uint4 *>(s_aw + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<uint4 *>(s_aw + CT_AW_PLANE + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
            for (int c0 = cbeg; c0 < cbeg + 128; c0 += 16) {
                float4 un[4], vn[4];
                if (c0 + 16 < cbeg + 128) ct_issue_chunk(a, cur, c0 + 16, un, vn);   // next chunk's loads before this chunk's math
                uint32_t hi[8], lo[8];
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const int c = c0 + 4 * g;
                    const float4 u = cur.u[g], v1 = cur.v[g];
                    const float4 wx = *reinterpret_cast<const float4 *>(s_wx + c);
                    const float4 wy = *reinterpret_cast<const float4 *>(s_wx + CT_C + c);
                    const float4 wz = *reinterpret_cast<const float4 *>(s_wx + 2 * CT_C + c);
                    // packed pairs, same per-channel operation order as the scalar form
                    // leaky(u + fma(wz, dz, fma(wy, dy, wx * dx)) + v1)
                    float2 ta = rt_fmul2(make_float2(wx.x, wx.y), dx2), tb = rt_fmul2(make_float2(wx.z, wx.w), dx2);
                    ta = rt_ffma2(make_float2(wy.x, wy.y), dy2, ta);
                    tb = rt_ffma2(make_float2(wy.z, wy.w), dy2, tb);
                    ta = rt_ffma2(make_float2(wz.x, wz.y), dz2, ta);
                    tb = rt_ffma2(make_float2(wz.z, wz.w), dz2, tb);
                    ta = rt_fadd2(rt_fadd2(make_float2(u.x, u.y), ta), make_float2(v1.x, v1.y));
                    tb = rt_fadd2(rt_fadd2(make_float2(u.z, u.w), tb), make_float2(v1.z, v1.w));
                    ta = leaky01x2(ta);
                    tb = leaky01x2(tb);
                    split2(ta.x, ta.y, hi[2 * g], lo[2 * g], amax);
                    split2(tb.x, tb.y, hi[2 * g + 1], lo[2 * g + 1], amax);
                }
                ct_st8(tAhi + lane_base + c0 / 2, hi);
                ct_st8(tAlo + lane_base + c0 / 2, lo);
                if (c0 + 16 < cbeg + 128) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        cur.u[g] = un[g];
                        cur.v[g] = vn[g];
                    }
                }
                if ((c0 & 16) != 0) {
                    // a 32-column K group is complete: hand it to the MMA warp, which starts layer 2 on it right away
                    ct_st_wait();
                    ct_fence_before();
                    rt_fence_proxy_async();  // s_aw stores -> async proxy (needed before the WeightNet MMA much later)
                    __syncwarp();
                    if (lane == 0) rt_mbar_arrive(&bar_a[c0 >> 5]);
                }
            }

            // ---------- mid epilogue: layer-2 accumulator -> bias, LeakyReLU -> A operand of layer 3 ----------
            ct_mbar_wait(bar_d, d_phase);
            d_phase ^= 1;
            ct_fence_after();
            for (int c0 = cbeg; c0 < cbeg + 128; c0 += 32) {
                uint32_t r[32], hi[16], lo[16];
                ct_ld32(tD + lane_base + c0, r);
                ct_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float2 b = *reinterpret_cast<const float2 *>(s_b2 + c0 + 2 * i);
                    const float2 x = leaky01x2(rt_ffma2(make_float2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])),
                                                        make_float2(CT_WINV, CT_WINV), b));
                    split2(x.x, x.y, hi[i], lo[i], amax);
                }
                ct_st16(tAhi + lane_base + c0 / 2, hi);
                ct_st16(tAlo + lane_base + c0 / 2, lo);
            }
            ct_st_wait();
            ct_fence_before();
            // layer 3 overwrites the accumulator every worker is still reading: release all K groups only when all 8
            // worker warps are done with it
            asm volatile("bar.sync 1, %0;" ::"n"(32 * CT_WORKER_WARPS) : "memory");
            if (lane == 0)
                for (int g = 4 * hlf; g < 4 * hlf + 4; ++g) rt_mbar_arrive(&bar_a[g]);
            // next tile's index / xyz / first P1,P2 chunk: in flight during the layer-3 MMAs and the final epilogue
            const int p_out = p;
            const bool valid_out = valid;
            if (tile + (int)gridDim.x < ntiles) ct_issue_row(a, tile + gridDim.x, row, cbeg, cur);

            // ---------- final epilogue: LeakyReLU(layer 3) * ReLU(WeightNet), summed over the 16 neighbours ----------
            ct_mbar_wait(bar_d, d_phase);
            d_phase ^= 1;
            ct_fence_after();
            for (int c0 = cbeg; c0 < cbeg + 128; c0 += 32) {
                uint32_t rd[32], rw[32];
                ct_ld32(tD + lane_base + c0, rd);
                ct_ld32(tW + lane_base + c0, rw);
                ct_ld_wait();
                float v[32];
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    const float2 winv = make_float2(CT_WINV, CT_WINV);
                    const float2 x = leaky01x2(rt_ffma2(make_float2(__uint_as_float(rd[i]), __uint_as_float(rd[i + 1])), winv,
                                                        *reinterpret_cast<const float2 *>(s_b3 + c0 + i)));
                    const float2 wl = rt_ffma2(make_float2(__uint_as_float(rw[i]), __uint_as_float(rw[i + 1])), winv,
                                               *reinterpret_cast<const float2 *>(s_bc + c0 + i));
                    const float2 pr = rt_fmul2(make_float2(fmaxf(wl.x, 0.0f), fmaxf(wl.y, 0.0f)), x);
                    v[i] = pr.x;
                    v[i + 1] = pr.y;
                }
                // butterfly over the 16 lanes of a point: every step halves what a lane keeps
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const bool up = lane & 8;
                    const float send = up ? v[i] : v[i + 16], keep = up ? v[i + 16] : v[i];
                    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const bool up = lane & 4;
                    const float send = up ? v[i] : v[i + 8], keep = up ? v[i + 8] : v[i];
                    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const bool up = lane & 2;
                    const float send = up ? v[i] : v[i + 4], keep = up ? v[i + 4] : v[i];
                    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
                }
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const bool up = lane & 1;
                    const float send = up ? v[i] : v[i + 2], keep = up ? v[i + 2] : v[i];
                    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
                }
                if (valid_out) {
                    const int col = c0 + ((lane & 8) ? 16 : 0) + ((lane & 4) ? 8 : 0) + ((lane & 2) ? 4 : 0) + ((lane & 1) ? 2 : 0);
                    *reinterpret_cast<float2 *>(a.out + (size_t)p_out * CT_C + col) = make_float2(v[0], v[1]);
                }
            }
            ct_fence_before();
            // nobody may overwrite the A planes / WeightNet accumulator before every worker has read them
            asm volatile("bar.sync 1, %0;" ::"n"(32 * CT_WORKER_WARPS) : "memory");
        }
        // fp16 range guard (|x| < 65504): report, never silently saturate
        if (!(amax < 65000.0f)) atomicOr(a.status, 2);
    }
    ct_fence_before();
    __syncthreads();
    if (warp == CT_WORKER_WARPS + 1) ct_tmem_dealloc(tm, 512);
}

}  // namespace

// engine-internal launcher.  wpack = [layer 2,3][8 chunks][hi,lo][kc 4][row group 32][8][8] fp16 of 2^10 * W;
// wcpack = [hi,lo][kc 2][row group 32][8][8] fp16 of 2^10 * Wc (K 8 padded to 16).
int rt_launch_costvol_tc_v2(int total_pts, int n, const float *p1, const float *p2, const float *xyz1, const float *xyz2,
                         const int *knn, const int *perm, const float *w1x, const void *wpack, const void *wcpack, const float *b2,
                         const float *b3, const float *bc, const float *wa, const float *ba, const float *wb, const float *bb,
                         float *out, int *status, cudaStream_t st) {
    if (total_pts <= 0) return RT_OK;
    static RtPerDevice attr_set;
    if (!attr_set.done(rt_current_device())) {
        cudaError_t e = cudaFuncSetAttribute(costvol_tc_v2_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(costvol_tc_v2_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL);
        if (e != cudaSuccess) {
            rt_set_error("costvol_tc: cannot reserve %d bytes of shared memory: %s", SM_TOTAL, cudaGetErrorString(e));
            return (int)e;
        }
        attr_set.mark(rt_current_device());
    }
    static int stages = 0;   // RT_CV_STAGES=4|6 (A/B timing)
    if (!stages) {
        const char *env = getenv("RT_CV_STAGES");
        stages = (env && atoi(env) == 6) ? 6 : 4;   // measured equal (467.5 vs 469.5 us): the ring is not what bounds the kernel
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int ntiles = (total_pts + CT_PTS - 1) / CT_PTS;
    CostVolTcArgs a{total_pts, n, p1, p2, xyz1, xyz2, knn, perm, w1x, (const __half *)wpack, (const __half *)wcpack,
                    b2, b3, bc, wa, ba, wb, bb, out, status};
    if (stages == 4) costvol_tc_v2_kernel<4><<<ntiles < sms ? ntiles : sms, CT_THREADS, SM_TOTAL, st>>>(a);
    else costvol_tc_v2_kernel<6><<<ntiles < sms ? ntiles : sms, CT_THREADS, SM_TOTAL, st>>>(a);
    return rt_check_launch("costvol_tc_v2_kernel");
}
