// Cost-volume core on the 5th-generation tensor cores (tcgen05 / TMEM / bulk-copy engine), sm_100a.
//
// Replaces, for FeatureCorrelator.forward (reference: src/utils/model_utils/model_utils.py:216-236), the chain
//     gather f2 / xyz of the 16 neighbours -> concat 515 channels -> conv 515->256 -> conv 256->256 -> conv 256->256
//     (LeakyReLU 0.1 each) -> WeightNet(dxyz) weighting -> sum over the 16 neighbours
// by ONE persistent kernel that never materialises a (B,515,16,N) / (B,256,16,N) tensor:
//   tile      = 128 rows = 8 query points x 16 neighbours; the row index is the TMEM lane;
//   layer 1   = P1[point] + P2[neighbour] + Wx.dxyz  (P1/P2 = per-POINT projections computed once by a plain GEMM),
//               evaluated by the worker warps straight into TMEM as the A operand of layer 2; the 8 P1 rows of a tile
//               arrive in shared memory through the bulk-copy engine (16 rows share one), P2 rows are 256-bit gathers;
//   layer 2/3 = tcgen05.mma kind::f16, M128 N256 K16, A operand read from TMEM, B operand (weights, canonical K-major
//               core-matrix layout, pre-packed on the host) streamed by the bulk-copy engine (UBLKCP): the hi planes of a
//               layer (128 KB) stay resident in shared memory from its correction pass to its main pass, the lo planes go
//               through a 3-stage ring -- 512 KB of L2 -> SM traffic per tile instead of 768 KB.  Measured with the phase
//               knock-outs (RT_CV_DEBUG): the weight stream alone, all SMs pulling 768 KB per tile, ran at the chip's
//               L2 -> SM limit (12.8 TB/s) and took 252 of the kernel's 443 us;
//   precision = every fp32 operand x is split x = hi + 2^-11 lo' (two fp16 planes, weights pre-scaled by 2^10) and each
//               product is evaluated as (lo'*hi + hi*lo') * 2^-11 + hi*hi with fp32 accumulation, corrections first
//               (tools/tc_precision.cu: the rms error of an fp32 FMA chain x 1.7) at 1/3 of the fp16 tensor rate;
//   WeightNet = its last layer (8 -> 256) is a K=16 MMA whose accumulator lands on the dead layer-3 A columns;
//   epilogue  = bias + LeakyReLU, weight, and a 16-lane butterfly that leaves the neighbour sum in registers.
//
// TMEM (512 columns) is two 256-column regions that swap roles inside a tile, so the tensor pipe and the workers overlap:
//     R1: layer-2 A operand   -> layer-3 accumulator
//     R0: layer-2 accumulator -> layer-3 A operand (written IN PLACE by the mid epilogue) -> WeightNet accumulator
// A K step (16 values) of an A operand lives where the 16 accumulator columns it was computed from lived (8 columns of
// hi halves, then 8 of lo halves), so every worker warp only ever touches its own lanes x its own columns: no CTA-wide
// barrier between the phases of a tile or between tiles.  The layer-3 MMAs of a K group start as soon as the mid
// epilogue has converted it (they write R1, which is dead), exactly as layer 2 trails the layer-1 production.
// 16 worker warps (4 TMEM lane quarters x 4 column sets) + 1 bulk-copy producer warp + 1 MMA-issuer warp.  Column set p
// owns the K steps p, p+4, p+8, p+12, so the four sets complete a 64-column K group together every quarter of a phase
// and the tensor pipe trails the workers by one group; the main products of layer 2 are issued as two N halves so the
// mid epilogue starts on the first 128 columns while the second half is still being accumulated.
#include <cuda_fp16.h>
#include <stdlib.h>

#include "engine_kernels.cuh"

namespace {

constexpr int CT_ROWS = 128, CT_PTS = 8, CT_NS = 16, CT_C = 256;
constexpr int CT_KC = 32;                          // K per streamed weight chunk = one K group of the A-operand hand-off
constexpr int CT_LSTAGES = 3;                      // ring depth of the lo planes (the hi planes of a layer stay resident, see below)
constexpr int CT_CHUNKS = CT_C / CT_KC;            // 8 streamed weight chunks per pass
constexpr int CT_GROUPS = 4;                       // A-operand hand-off units: 64 columns = 4 K steps, one per column set
constexpr int CT_PLANE_BYTES = CT_C * CT_KC * 2;   // one fp16 plane of a chunk: 16 KB
constexpr int CT_STAGE_BYTES = 2 * CT_PLANE_BYTES; // hi + lo
constexpr int CT_WC_PLANE = CT_C * 16 * 2;         // WeightNet last layer, K padded 8 -> 16: 8 KB
constexpr int CT_AW_PLANE = CT_ROWS * 16 * 2;      // its A operand: 4 KB
constexpr int CT_PARTS = 4;                        // column sets: set p owns K steps p, p + 4, p + 8, p + 12
constexpr int CT_WORKER_WARPS = 4 * CT_PARTS;
constexpr int CT_THREADS = 32 * (CT_WORKER_WARPS + 2);
constexpr int CT_P1_BYTES = CT_PTS * CT_C * 4;     // the P1 rows of one tile: 8 KB
constexpr float CT_WINV = 1.0f / 1024.0f;          // weights are packed as 2^10 * W

// shared memory carve-up (bytes)
constexpr int SM_HI = 0;                                        // hi planes of all 8 K chunks of the current layer: 128 KB
constexpr int SM_LO = SM_HI + CT_CHUNKS * CT_PLANE_BYTES;       // lo-plane ring
constexpr int SM_P1 = SM_LO + CT_LSTAGES * CT_PLANE_BYTES;      // [2][8][256] fp32
constexpr int SM_WC = SM_P1 + 2 * CT_P1_BYTES;
constexpr int SM_AW = SM_WC + 2 * CT_WC_PLANE;
constexpr int SM_B2 = SM_AW + 2 * CT_AW_PLANE;
constexpr int SM_B3 = SM_B2 + CT_C * 4;
constexpr int SM_BC = SM_B3 + CT_C * 4;
constexpr int SM_WX = SM_BC + CT_C * 4;            // [3][256]
constexpr int SM_WN = SM_WX + 3 * CT_C * 4;        // wa(24) ba(8) wb(64) bb(8)
constexpr int SM_BAR = SM_WN + 128 * 4;
constexpr int SM_TOTAL = SM_BAR + 40 * 8;

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
    unsigned long long *trace;   // RT_CV_TRACE: clock64 stamps of CTA 0 (worker warp 0, MMA thread), 32 per tile per role
    int debug;   // timing experiments only (RT_CV_DEBUG): 1 no MMAs, 2 no layer-1 work, 4 no mid epilogue, 8 no final epilogue, 16 / 32 no lo / hi weight traffic
};

#define CT_TRACE(role, slot)                                                                                      \
    do {                                                                                                          \
        if (DBG && a.trace && blockIdx.x == 0 && trace_it < 6) a.trace[(trace_it * 2 + (role)) * 32 + (slot)] = clock64(); \
    } while (0)

// ---- PTX wrappers ------------------------------------------------------------------------------------
// workers back off between probes (a spinning warp steals issue slots from the warps that have work) ...
__device__ __forceinline__ void ct_mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = rt_smem_u32(bar);
    for (uint32_t it = 0;; ++it) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
        if (ok) return;
        if (it >= 8) __nanosleep(it < 64 ? 20 : 96);   // try_wait already suspends for a while; sleep only on long waits
        if (it > (1u << 22)) __trap();  // a protocol bug must fail loudly, never hang the device
    }
}
// ... the single MMA-issuer thread does not: its wake-up latency is on the tensor pipe's critical path
__device__ __forceinline__ void ct_mbar_spin(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = rt_smem_u32(bar);
    for (uint32_t it = 0;; ++it) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
        if (ok) return;
        if (it > (1u << 26)) __trap();
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
__device__ __forceinline__ void ct_ld16(uint32_t taddr, uint32_t *r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
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
// TMEM column (inside a region) of the hi plane of K step kk (16 K values = 8 columns); the lo plane is 8 columns further
__device__ __forceinline__ uint32_t ct_acol(int kk) { return 16u * (uint32_t)kk; }
// F16 x F16 -> F32, A and B K-major, M = 128, N = 256 / 128
constexpr uint32_t CT_IDESC = (1u << 4) | ((uint32_t)(CT_C >> 3) << 17) | ((uint32_t)(CT_ROWS >> 4) << 24);
constexpr uint32_t CT_IDESC_H = (1u << 4) | ((uint32_t)(CT_C >> 4) << 17) | ((uint32_t)(CT_ROWS >> 4) << 24);

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
    const float2 d = rt_fmul2(rt_fadd2(make_float2(x0, x1), make_float2(-hf.x, -hf.y)), make_float2(2048.0f, 2048.0f));
    const __half2 l = __floats2half2_rn(d.x, d.y);   // lo plane stored as 2^11 * lo: normal whenever hi is
    hi = *reinterpret_cast<const uint32_t *>(&h);
    lo = *reinterpret_cast<const uint32_t *>(&l);
}

// Per-row operands of a tile's layer 1.  They are fetched one tile ahead (head) and one 16-channel chunk ahead (body)
// so the index -> xyz -> P2 load chain hides under the MMA phases instead of stalling the prologue.
struct CtRow {
    int p;
    bool valid;
    size_t g2;
    float dx, dy, dz;
    float4 u[4];   // chunk of P2[neighbour]
};
__device__ __forceinline__ void ct_issue_chunk(const CostVolTcArgs &a, const CtRow &r, int c0, float4 *u) {
    const float *p2r = a.p2 + r.g2 * CT_C + c0;
    // 256-bit loads: every lane reads a different neighbour row, so the L1 data pipe pays per instruction, not per byte
    rt_ldg256(p2r, u[0], u[1]);
    rt_ldg256(p2r + 8, u[2], u[3]);
}
__device__ __forceinline__ void ct_issue_row(const CostVolTcArgs &a, int tile, int row, int cbeg, CtRow &r) {
    const int slot = tile * CT_PTS + (row >> 4);
    r.valid = slot < a.total_pts;
    r.p = r.valid ? (a.perm ? __ldg(a.perm + slot) : slot) : a.total_pts - 1;
    const int cloud = r.p / a.n;
    const int nbr = __ldg(a.knn + (size_t)r.p * CT_NS + (row & 15));
    r.g2 = (size_t)cloud * a.n + nbr;
    r.dx = __ldg(a.xyz2 + r.g2 * 3 + 0) - __ldg(a.xyz1 + (size_t)r.p * 3 + 0);
    r.dy = __ldg(a.xyz2 + r.g2 * 3 + 1) - __ldg(a.xyz1 + (size_t)r.p * 3 + 1);
    r.dz = __ldg(a.xyz2 + r.g2 * 3 + 2) - __ldg(a.xyz1 + (size_t)r.p * 3 + 2);
    ct_issue_chunk(a, r, cbeg, r.u);
}

template <bool DBG>
__global__ void __launch_bounds__(CT_THREADS, 1) costvol_tc_kernel(CostVolTcArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *s_hi = smem + SM_HI;
    uint8_t *s_lo = smem + SM_LO;
    uint8_t *s_p1 = smem + SM_P1;
    uint8_t *s_wc = smem + SM_WC;
    uint8_t *s_aw = smem + SM_AW;
    float *s_b2 = reinterpret_cast<float *>(smem + SM_B2);
    float *s_b3 = reinterpret_cast<float *>(smem + SM_B3);
    float *s_bc = reinterpret_cast<float *>(smem + SM_BC);
    float *s_wx = reinterpret_cast<float *>(smem + SM_WX);
    float *s_wn = reinterpret_cast<float *>(smem + SM_WN);
    uint64_t *bar_hfull = reinterpret_cast<uint64_t *>(smem + SM_BAR);   // [CT_CHUNKS] hi plane of K chunk c has landed
    uint64_t *bar_hempty = bar_hfull + CT_CHUNKS;                          // [CT_CHUNKS] ... and the layer's last MMA on it has read it
    uint64_t *bar_full = bar_hempty + CT_CHUNKS;                           // [CT_LSTAGES] lo-plane ring
    uint64_t *bar_empty = bar_full + CT_LSTAGES;
    uint64_t *bar_a = bar_empty + CT_LSTAGES;  // [CT_GROUPS]: the A columns of one 64-wide K group are in TMEM
    uint64_t *bar_d2 = bar_a + CT_GROUPS;      // [2] layer-2 accumulator, N half h, is complete
    uint64_t *bar_d3 = bar_d2 + 2;             // layer-3 + WeightNet accumulators are complete
    uint64_t *bar_epi = bar_d3 + 1;            // every worker warp has read the tile's last accumulators
    uint64_t *bar_pf = bar_epi + 1;            // [2] P1 rows of a tile have landed
    uint64_t *bar_pe = bar_pf + 2;             // [2] ... and have been consumed
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar_pe + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntiles = (a.total_pts + CT_PTS - 1) / CT_PTS;

    // ---- one-time setup --------------------------------------------------------------------------
    if (threadIdx.x == 0) {
        for (int s = 0; s < CT_LSTAGES; ++s) {
            rt_mbar_init(&bar_full[s], 1);
            rt_mbar_init(&bar_empty[s], 1);
        }
        for (int c = 0; c < CT_CHUNKS; ++c) {
            rt_mbar_init(&bar_hfull[c], 1);
            rt_mbar_init(&bar_hempty[c], 1);
        }
        for (int g = 0; g < CT_GROUPS; ++g) rt_mbar_init(&bar_a[g], CT_WORKER_WARPS);   // every warp owns one K step of a group
        rt_mbar_init(&bar_d2[0], 1);
        rt_mbar_init(&bar_d2[1], 1);
        rt_mbar_init(bar_d3, 1);
        rt_mbar_init(bar_epi, CT_WORKER_WARPS);
        for (int s = 0; s < 2; ++s) {
            rt_mbar_init(&bar_pf[s], 1);
            rt_mbar_init(&bar_pe[s], CT_WORKER_WARPS);
        }
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
    const uint32_t tR0 = tm, tR1 = tm + 256;

    if (warp == CT_WORKER_WARPS) {
        // ===== producer: one elected lane feeds the weight ring and the P1 rows with the bulk-copy engine =====
        if (lane == 0) {
            int stage = 0, it = 0;
            uint32_t phase = 0;
            auto issue_p1 = [&](int tile, int buf) {
                rt_mbar_expect_tx(&bar_pf[buf], CT_P1_BYTES);
                for (int i = 0; i < CT_PTS; ++i) {
                    const int slot = tile * CT_PTS + i;
                    const int p = slot < a.total_pts ? (a.perm ? __ldg(a.perm + slot) : slot) : a.total_pts - 1;
                    rt_bulk_g2s(s_p1 + buf * CT_P1_BYTES + i * (CT_C * 4), a.p1 + (size_t)p * CT_C, CT_C * 4, &bar_pf[buf]);
                }
            };
            if ((int)blockIdx.x < ntiles) issue_p1(blockIdx.x, 0);
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
                if (tile + (int)gridDim.x < ntiles) {
                    // the buffer of tile it+1 was last used by tile it-1, whose layer 1 finished long ago (this thread is
                    // at most CT_STAGES weight chunks ahead of the MMAs): the wait never blocks in practice
                    const int nb = (it + 1) & 1, fill = (it + 1) >> 1;
                    if (fill >= 1) ct_mbar_wait(&bar_pe[nb], (uint32_t)(fill - 1) & 1u);
                    issue_p1(tile + gridDim.x, nb);
                }
                const uint8_t *wp = reinterpret_cast<const uint8_t *>(a.wpack);
                for (int layer = 0; layer < 2; ++layer) {
                    // one pass per layer: chunk c = [hi plane -> its resident slot][lo plane -> ring]; the main products
                    // re-read the resident hi planes, nothing is streamed twice
                    const int fill = 2 * it + layer;   // how many times every hi slot has been filled before
                    for (int c = 0; c < CT_CHUNKS; ++c) {
                        const uint8_t *src = wp + (size_t)(layer * CT_CHUNKS + c) * CT_STAGE_BYTES;
                        if (fill >= 1) ct_mbar_wait(&bar_hempty[c], (uint32_t)(fill - 1) & 1u);
                        if (DBG && (a.debug & 32)) rt_mbar_arrive(&bar_hfull[c]);   // timing experiment: no hi-plane traffic
                        else {
                            rt_mbar_expect_tx(&bar_hfull[c], CT_PLANE_BYTES);
                            rt_bulk_g2s(s_hi + c * CT_PLANE_BYTES, src, CT_PLANE_BYTES, &bar_hfull[c]);
                        }
                        ct_mbar_wait(&bar_empty[stage], phase ^ 1);
                        if (DBG && (a.debug & 16)) rt_mbar_arrive(&bar_full[stage]);   // timing experiment: no lo-plane traffic
                        else {
                            rt_mbar_expect_tx(&bar_full[stage], CT_PLANE_BYTES);
                            rt_bulk_g2s(s_lo + stage * CT_PLANE_BYTES, src + CT_PLANE_BYTES, CT_PLANE_BYTES, &bar_full[stage]);
                        }
                        if (++stage == CT_LSTAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == CT_WORKER_WARPS + 1) {
        // ===== MMA issuer: a single thread drives the tensor core =====
        if (lane == 0) {
            int stage = 0, fill = 0;   // fill = (tile iteration, layer) counter: parity of the resident hi slots
            uint32_t phase = 0, a_phase = 0, e_phase = 0;
            const uint32_t aw_hi = rt_smem_u32(s_aw), aw_lo = aw_hi + CT_AW_PLANE;
            const uint32_t wc_hi = rt_smem_u32(s_wc), wc_lo = wc_hi + CT_WC_PLANE;
            const uint32_t hi0 = rt_smem_u32(s_hi);
            int trace_it = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++trace_it) {
                CT_TRACE(1, 0);
                if (tile != (int)blockIdx.x) {
                    // layer 2 accumulates into R0, which holds the previous tile's WeightNet accumulator until every
                    // worker warp has read it
                    ct_mbar_spin(bar_epi, e_phase);
                    e_phase ^= 1;
                    ct_fence_after();
                }
                CT_TRACE(1, 1);
                for (int layer = 0; layer < 2; ++layer, ++fill) {
                    const uint32_t tD = layer == 0 ? tR0 : tR1, tA = layer == 0 ? tR1 : tR0;
                    // Correction products first (lo*hi + hi*lo over all of K, carrying the 2^11 of the scaled lo planes),
                    // main products (hi*hi) last, the first of them rescaling the partial sum by 2^-11: the tensor core
                    // truncates the fp32 accumulator after every k-step, so the number of accumulation steps taken at
                    // full magnitude sets the error (tools/tc_precision.cu: 3x smaller rms than interleaving).
                    for (int c = 0; c < CT_CHUNKS; ++c) {
                        if ((c & 1) == 0) {
                            ct_mbar_spin(&bar_a[c >> 1], a_phase);   // K steps 4g .. 4g+3 of the A operand have landed in TMEM:
                            ct_fence_after();                        // the MMAs start while the workers still produce later groups
                            CT_TRACE(1, 2 + 12 * layer + (c >> 1));   // group g arrived
                        }
                        ct_mbar_spin(&bar_hfull[c], (uint32_t)fill & 1u);
                        ct_mbar_spin(&bar_full[stage], phase);
                        ct_fence_after();
                        const uint32_t bh = hi0 + c * CT_PLANE_BYTES, bl = rt_smem_u32(s_lo + stage * CT_PLANE_BYTES);
#pragma unroll
                        for (int j = 0; j < CT_KC / 16; ++j) {
                            const uint32_t ahi = tA + ct_acol(2 * c + j);
                            // chunk plane = [kc = K/8][row group = 32][8 rows][8 halfs]: kc stride 4096 B, row group 128 B
                            if (DBG && (a.debug & 1)) continue;
                            ct_mma_ts(tD, ahi + 8, ct_desc(bh + j * 8192, 4096, 128), CT_IDESC, (c | j) > 0);
                            ct_mma_ts(tD, ahi, ct_desc(bl + j * 8192, 4096, 128), CT_IDESC, 1);
                        }
                        ct_commit(&bar_empty[stage]);  // frees the lo ring slot when these MMAs have read it
                        if (++stage == CT_LSTAGES) { stage = 0; phase ^= 1; }
                        if (c & 1) CT_TRACE(1, 6 + 12 * layer + (c >> 1));   // corrections of group g issued
                    }
                    if (layer == 0) {
                        // main products of layer 2 as two N halves over the whole of K (all hi planes are resident):
                        // the mid epilogue starts on columns [0,128) while [128,256) is still being accumulated
                        for (int half = 0; half < 2; ++half) {
                            for (int c = 0; c < CT_CHUNKS; ++c) {
                                const uint32_t base = hi0 + c * CT_PLANE_BYTES + half * 2048;   // row groups 16h ..
#pragma unroll
                                for (int j = 0; j < CT_KC / 16; ++j) {
                                    const int kk = 2 * c + j;
                                    if (DBG && (a.debug & 1)) continue;
                                    if (kk == 0) ct_mma_ts_rescale(tD + 128 * half, tA + ct_acol(0), ct_desc(base, 4096, 128), CT_IDESC_H);
                                    else ct_mma_ts(tD + 128 * half, tA + ct_acol(kk), ct_desc(base + j * 8192, 4096, 128), CT_IDESC_H, 1);
                                }
                                if (half == 1) ct_commit(&bar_hempty[c]);   // the next layer's hi plane may land here
                            }
                            ct_commit(&bar_d2[half]);
                            CT_TRACE(1, 10 + half);
                        }
                    } else {
                        for (int c = 0; c < CT_CHUNKS; ++c) {
                            const uint32_t base = hi0 + c * CT_PLANE_BYTES;
#pragma unroll
                            for (int j = 0; j < CT_KC / 16; ++j) {
                                const int kk = 2 * c + j;
                                if (DBG && (a.debug & 1)) continue;
                                if (kk == 0) ct_mma_ts_rescale(tD, tA + ct_acol(0), ct_desc(base, 4096, 128), CT_IDESC);   // D = A*B + D * 2^-11
                                else ct_mma_ts(tD, tA + ct_acol(kk), ct_desc(base + j * 8192, 4096, 128), CT_IDESC, 1);
                            }
                            ct_commit(&bar_hempty[c]);
                        }
                        // WeightNet last layer: [128 x 16] . [256 x 16]^T -> R0 (the layer-3 A operand is dead now;
                        // tcgen05.mma executes in issue order)
                        ct_mma_ss(tR0, ct_desc(aw_lo, 2048, 128), ct_desc(wc_hi, 4096, 128), CT_IDESC, 0);
                        ct_mma_ss(tR0, ct_desc(aw_hi, 2048, 128), ct_desc(wc_lo, 4096, 128), CT_IDESC, 1);
                        ct_mma_ss_rescale(tR0, ct_desc(aw_hi, 2048, 128), ct_desc(wc_hi, 4096, 128), CT_IDESC);
                        ct_commit(bar_d3);
                        CT_TRACE(1, 22);
                    }
                    a_phase ^= 1;
                }
            }
        }
        __syncwarp();
    } else {
        // ===== worker warps: layer 1 (gather + combine) -> TMEM, mid epilogue, final epilogue =====
        const int q = warp & 3, part = warp >> 2;
        const int row = 32 * q + lane;
        const uint32_t lane_base = (uint32_t)(32 * q) << 16;
        uint32_t d_phase = 0;
        int it = 0;
        float amax = 0.0f;
        CtRow cur;
        if ((int)blockIdx.x < ntiles) ct_issue_row(a, blockIdx.x, row, 16 * part, cur);
        const int trace_it_base = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int trace_it = (warp == 0 && lane == 0) ? it + trace_it_base : 99;
            CT_TRACE(0, 0);
            // ---------- layer 1 ----------  (row operands + first chunk were issued during the previous tile)
            const int p_out = cur.p;
            const bool valid_out = cur.valid;
            const float2 dx2 = make_float2(cur.dx, cur.dx), dy2 = make_float2(cur.dy, cur.dy), dz2 = make_float2(cur.dz, cur.dz);
            if (part == 0) {
                // WeightNet trunk 3 -> 8 -> 8 (ReLU), result = A operand (K = 8, padded to 16) of the last-layer MMA.
                // (The previous tile's WeightNet MMA has completed: this warp has passed that tile's bar_d3.)
                float h1[8], h2[8];
#pragma unroll
                for (int o = 0; o < 8; ++o) {
                    float v = s_wn[24 + o];
                    v = fmaf(s_wn[o * 3 + 0], cur.dx, v);
                    v = fmaf(s_wn[o * 3 + 1], cur.dy, v);
                    v = fmaf(s_wn[o * 3 + 2], cur.dz, v);
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
            const int buf = it & 1;
            ct_mbar_wait(&bar_pf[buf], (uint32_t)(it >> 1) & 1u);
            const float *p1row = reinterpret_cast<const float *>(s_p1 + buf * CT_P1_BYTES) + (row >> 4) * CT_C;
            CT_TRACE(0, 1);
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                const int c0 = 16 * (part + 4 * ch);   // K step part + 4 ch
                float4 un[4];
                if (ch < 3 && !(DBG && (a.debug & 2))) ct_issue_chunk(a, cur, c0 + 64, un);   // next chunk's loads before this chunk's math
                uint32_t hi[8], lo[8];
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    if (DBG && (a.debug & 2)) { hi[2 * g] = hi[2 * g + 1] = lo[2 * g] = lo[2 * g + 1] = 0; continue; }
                    const int c = c0 + 4 * g;
                    const float4 u = cur.u[g];
                    const float4 v1 = *reinterpret_cast<const float4 *>(p1row + c);   // 16 lanes share a row: broadcast
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
                const uint32_t t = tR1 + lane_base + c0;   // = ct_acol(K step): hi halves, then lo halves
                if (!(DBG && (a.debug & 2))) {
                    ct_st8(t, hi);
                    ct_st8(t + 8, lo);
                }
                if (ch < 3) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) cur.u[g] = un[g];
                }
                // this warp's K step of group ch is complete: the MMA warp starts layer 2 on the group once all 16 have arrived
                ct_st_wait();
                ct_fence_before();
                if (part == 0 && ch == 0) rt_fence_proxy_async();  // s_aw stores -> async proxy (WeightNet MMA, much later)
                __syncwarp();
                if (lane == 0) {
                    rt_mbar_arrive(&bar_a[ch]);
                    if (ch == 3) rt_mbar_arrive(&bar_pe[buf]);   // this warp is done with the tile's P1 rows
                }
                CT_TRACE(0, 2 + ch);
            }

            // ---------- mid epilogue: layer-2 accumulator -> bias, LeakyReLU -> A operand of layer 3, in place ----------
            // two K steps (one per N half ... of the same half) per TMEM round trip: load both, convert both, store both
#pragma unroll
            for (int hp = 0; hp < 2; ++hp) {
                ct_mbar_wait(&bar_d2[hp], d_phase);   // columns [128 hp, 128 hp + 128)
                ct_fence_after();
                CT_TRACE(0, 6 + hp);
                const int ca = 16 * (part + 8 * hp), cb = ca + 64;   // K steps part + 8 hp and part + 8 hp + 4
                uint32_t r[32], hi[16], lo[16];
                if (!(DBG && (a.debug & 4))) {
                    ct_ld16(tR0 + lane_base + ca, r);
                    ct_ld16(tR0 + lane_base + cb, r + 16);
                    ct_ld_wait();
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) r[i] = 0;
                }
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    if (DBG && (a.debug & 4)) { hi[i] = lo[i] = 0; continue; }
                    const int c = (i < 8 ? ca : cb - 16) + 2 * i;
                    const float2 b = *reinterpret_cast<const float2 *>(s_b2 + c);
                    const float2 x = leaky01x2(rt_ffma2(make_float2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])),
                                                        make_float2(CT_WINV, CT_WINV), b));
                    split2(x.x, x.y, hi[i], lo[i], amax);
                }
                if (!(DBG && (a.debug & 4))) {
                    ct_st8(tR0 + lane_base + ca, hi);
                    ct_st8(tR0 + lane_base + ca + 8, lo);
                    ct_st8(tR0 + lane_base + cb, hi + 8);
                    ct_st8(tR0 + lane_base + cb + 8, lo + 8);
                }
                ct_st_wait();
                ct_fence_before();
                __syncwarp();
                if (lane == 0) {   // layer 3 starts on these K groups (its accumulator is R1: dead)
                    rt_mbar_arrive(&bar_a[2 * hp]);
                    rt_mbar_arrive(&bar_a[2 * hp + 1]);
                }
                CT_TRACE(0, 8 + 2 * hp);
                CT_TRACE(0, 9 + 2 * hp);
            }
            // next tile's index / xyz / first P2 chunk: in flight during the layer-3 MMAs and the final epilogue
            if (tile + (int)gridDim.x < ntiles) ct_issue_row(a, tile + gridDim.x, row, 16 * part, cur);

            // ---------- final epilogue: LeakyReLU(layer 3) * ReLU(WeightNet), summed over the 16 neighbours ----------
            CT_TRACE(0, 12);
            ct_mbar_wait(bar_d3, d_phase);
            d_phase ^= 1;
            ct_fence_after();
            CT_TRACE(0, 13);
#pragma unroll
            for (int hc = 0; hc < 2; ++hc) {
                if (DBG && (a.debug & 8)) continue;
                // two of the warp's K-step column blocks per pass: 32 values per row, as in the 32-column butterfly
                const int ca = 16 * (part + 8 * hc), cb = ca + 64;
                uint32_t rd[32], rw[32];
                ct_ld16(tR1 + lane_base + ca, rd);
                ct_ld16(tR1 + lane_base + cb, rd + 16);
                ct_ld16(tR0 + lane_base + ca, rw);
                ct_ld16(tR0 + lane_base + cb, rw + 16);
                ct_ld_wait();
                float v[32];
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    const int c = (i < 16 ? ca : cb - 16) + i;
                    const float2 winv = make_float2(CT_WINV, CT_WINV);
                    const float2 x = leaky01x2(rt_ffma2(make_float2(__uint_as_float(rd[i]), __uint_as_float(rd[i + 1])), winv,
                                                        *reinterpret_cast<const float2 *>(s_b3 + c)));
                    const float2 wl = rt_ffma2(make_float2(__uint_as_float(rw[i]), __uint_as_float(rw[i + 1])), winv,
                                               *reinterpret_cast<const float2 *>(s_bc + c));
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
                    // value index inside the 32: lanes with bit 3 hold the second block (cb), the rest of the bits the column
                    const int col = ((lane & 8) ? cb : ca) + ((lane & 4) ? 8 : 0) + ((lane & 2) ? 4 : 0) + ((lane & 1) ? 2 : 0);
                    *reinterpret_cast<float2 *>(a.out + (size_t)p_out * CT_C + col) = make_float2(v[0], v[1]);
                }
            }
            // the MMA warp may overwrite R0 (next tile's layer-2 accumulator) once every worker warp has passed here; this
            // warp's own next layer 1 only writes the R1 cells it has just read
            ct_fence_before();
            __syncwarp();
            if (lane == 0) rt_mbar_arrive(bar_epi);
            CT_TRACE(0, 14);
        }
        // fp16 range guard (|x| < 65504): report, never silently saturate
        if (!(amax < 65000.0f) && !(DBG && (a.debug & 48))) atomicOr(a.status, 2);   // (stale weights of the traffic knock-outs are not data)
    }
    ct_fence_before();
    __syncthreads();
    if (warp == CT_WORKER_WARPS + 1) ct_tmem_dealloc(tm, 512);
}

}  // namespace

// engine-internal launcher.  wpack = [layer 2,3][8 chunks][hi,lo][kc 4][row group 32][8][8] fp16 of 2^10 * W;
// wcpack = [hi,lo][kc 2][row group 32][8][8] fp16 of 2^10 * Wc (K 8 padded to 16).
int rt_launch_costvol_tc(int total_pts, int n, const float *p1, const float *p2, const float *xyz1, const float *xyz2,
                         const int *knn, const int *perm, const float *w1x, const void *wpack, const void *wcpack, const float *b2,
                         const float *b3, const float *bc, const float *wa, const float *ba, const float *wb, const float *bb,
                         float *out, int *status, cudaStream_t st) {
    if (total_pts <= 0) return RT_OK;
    static RtPerDevice attr_set;
    if (!attr_set.done(rt_current_device())) {
        cudaError_t e = cudaFuncSetAttribute(costvol_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(costvol_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL);
        if (e != cudaSuccess) {
            rt_set_error("costvol_tc: cannot reserve %d bytes of shared memory: %s", SM_TOTAL, cudaGetErrorString(e));
            return (int)e;
        }
        attr_set.mark(rt_current_device());
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int ntiles = (total_pts + CT_PTS - 1) / CT_PTS;
    if (const char *env = getenv("RT_CV_GRID")) sms = atoi(env) > 0 ? atoi(env) : sms;   // experiment: fewer CTAs (is the kernel SM-local or L2-bound?)
    CostVolTcArgs a{total_pts, n, p1, p2, xyz1, xyz2, knn, perm, w1x, (const __half *)wpack, (const __half *)wcpack,
                    b2, b3, bc, wa, ba, wb, bb, out, status, nullptr, 0};
    if (const char *env = getenv("RT_CV_DEBUG")) a.debug = atoi(env);
    static unsigned long long *trace_buf = nullptr;
    const bool trace = getenv("RT_CV_TRACE") != nullptr;
    if (trace) {
        if (!trace_buf) cudaMalloc(&trace_buf, 6 * 2 * 32 * 8);
        cudaMemsetAsync(trace_buf, 0, 6 * 2 * 32 * 8, st);
        a.trace = trace_buf;
        if (!a.debug) a.debug = 64;
    }
    if (a.debug) costvol_tc_kernel<true><<<ntiles < sms ? ntiles : sms, CT_THREADS, SM_TOTAL, st>>>(a);   // phase knock-outs: timing only
    else costvol_tc_kernel<false><<<ntiles < sms ? ntiles : sms, CT_THREADS, SM_TOTAL, st>>>(a);
    if (trace) {   // debugging aid: per-phase clock64 stamps of CTA 0, relative to the first stamp of each tile
        unsigned long long h[6 * 2 * 32];
        cudaStreamSynchronize(st);
        cudaMemcpy(h, trace_buf, sizeof(h), cudaMemcpyDeviceToHost);
        for (int it = 1; it < 4; ++it) {
            const unsigned long long t0 = h[(it * 2) * 32];
            fprintf(stderr, "[costvol trace] tile %d worker:", it);
            for (int i = 0; i < 15; ++i) fprintf(stderr, " %lld", h[(it * 2) * 32 + i] ? (long long)(h[(it * 2) * 32 + i] - t0) : -1ll);
            fprintf(stderr, "\n[costvol trace] tile %d mma   :", it);
            for (int i = 0; i < 23; ++i) fprintf(stderr, " %lld", h[(it * 2 + 1) * 32 + i] ? (long long)(h[(it * 2 + 1) * 32 + i] - t0) : -1ll);
            fprintf(stderr, "\n");
        }
    }
    return rt_check_launch("costvol_tc_kernel");
}
