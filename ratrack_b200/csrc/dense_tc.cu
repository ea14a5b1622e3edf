// Dense layers of the TRAINING path on the 5th-generation tensor cores (tcgen05 / TMEM), sm_100a.
//
// The reference trains through torch autograd over cuDNN / cuBLAS fp32 kernels: every 1x1 convolution and Linear of
// SharedMLP, FeatureCorrelator, Flow/ClsPredictor and PNHead (reference: src/lib/pytorch_utils.py:35-101,
// src/utils/model_utils/model_utils.py:223-231, 308-357, 393-424) is a (rows x K) . (N x K)^T product with rows = every
// (batch, point, neighbour) position -- millions of rows, K, N <= 528.  Three products per layer and step:
//     forward   Y  = X  . W^T (+ bias)          rt_dense_tc_forward
//     dgrad     dX = dY . W                     rt_dense_tc_forward on dY with the transposed weight view
//     wgrad     dW = dY^T . X                   rt_dense_tc_wgrad
// fp32 in, fp32 out, fp32-class accuracy: every operand is split x = hi + 2^-11 lo' into two fp16 planes and a product is
// hi*hi + 2^-11 (lo'*hi + hi*lo') with the two parts in SEPARATE fp32 TMEM accumulators (the tensor core truncates the
// accumulator after every K step -- tools/tc_precision.cu -- so the small terms never ride on the large ones).
//
// lin_tc_kernel (forward / dgrad): persistent CTAs, tile = 128 rows x one N tile (<= 128 columns, the CTA's weights stay
// resident in shared memory as fp16 planes in the K-major core-matrix layout).  Warp roles:
//     16 loader warps  : coalesced 128-bit row reads (4 rows x 128 contiguous bytes per warp instruction), fp32 -> hi/lo
//                        fp16, conflict-free 64-bit stores into a 3-stage ring of K-major core-matrix planes (64 K per stage)
//     1 MMA thread     : tcgen05.mma kind::f16 M128 N<=128 K16, A and B from shared memory, 3 products per K step
//     4 epilogue warps : TMEM -> registers -> (main + 2^-11 corr) / scale + bias -> 256-bit row stores
// with two accumulator pairs in TMEM, so loading tile t+1, the MMAs of tile t and the epilogue of tile t-1 overlap.
// The N tiles of a row tile are neighbouring CTAs of the grid: they run at the same time, the second read of the rows is an
// L2 hit.
//
// wgrad_tc_kernel: the reduction runs over ROWS, so both operands are staged TRANSPOSED: a loader thread owns one channel
// and 8 consecutive rows (8 coalesced 32-bit loads per warp instruction -> one 16-byte core-matrix row), tile = 128 dY
// channels x <= 256 X channels, 64 rows per stage, split over the rows across the grid; the per-CTA partial sums are added
// in a fixed order by a second kernel (bit-repeatable, like every other gradient of this library).
//
// Operands with an unbounded dynamic range (gradients) are scaled by a power of two taken from their absolute maximum
// (rt_absmax, one pass, read on the device: no host synchronisation) so that the fp16 planes stay normal; a value outside
// the fp16 range turns into inf / NaN in the result -- loud, never silently saturated.
#include <cuda_fp16.h>

#include "common.cuh"

namespace {

constexpr int LT_ROWS = 128, LT_KC = 64, LT_NST = 3;
// A-operand stage: K-major core matrices (8 rows x 16 bytes), 8-row groups 128 bytes apart (SBO), K-adjacent core matrices
// LT_LBO bytes apart.  LBO is 2048 + 32 so that the 16 lanes of a half warp -- two rows x eight 16-byte chunks of a coalesced
// 128-bit row read -- store their 8-byte half rows to 16 distinct 8-byte bank slots (conflict-free 64-bit stores).
constexpr int LT_LBO = 16 * LT_ROWS + 32;
constexpr int LT_PLANE = (LT_KC / 8) * LT_LBO;  // one fp16 plane of a stage
constexpr int LT_STAGE = 2 * LT_PLANE;
constexpr int LT_EPI_WARPS = 4, LT_LOAD_WARPS = 16;
constexpr int LT_TASKS = 64 / LT_LOAD_WARPS;   // warp tasks (4 rows x 128 bytes) per loader warp and stage
constexpr int LT_THREADS = 32 * (LT_EPI_WARPS + LT_LOAD_WARPS + 1);
constexpr float LT_WSCALE = 1024.0f;            // weights are staged as 2^10 W: their lo plane stays normal
constexpr float LT_LO = 2048.0f;                // lo planes are stored as 2^11 lo

// ---- PTX wrappers -------------------------------------------------------------------------------------------------
__device__ __forceinline__ void dt_wait(uint64_t *bar, uint32_t parity) {   // workers: back off between probes
    const uint32_t addr = rt_smem_u32(bar);
    for (uint32_t it = 0;; ++it) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
        if (ok) return;
        if (it >= 8) __nanosleep(it < 64 ? 20 : 96);
        if (it > (1u << 22)) __trap();   // a protocol bug must fail loudly, never hang the device
    }
}
__device__ __forceinline__ void dt_spin(uint64_t *bar, uint32_t parity) {   // the MMA thread: wake-up latency matters
    const uint32_t addr = rt_smem_u32(bar);
    for (uint32_t it = 0;; ++it) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
        if (ok) return;
        if (it > (1u << 26)) __trap();
    }
}
__device__ __forceinline__ void dt_tmem_alloc(uint32_t *slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(rt_smem_u32(slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void dt_tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void dt_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void dt_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void dt_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(rt_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void dt_mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void dt_ld16(uint32_t taddr, uint32_t *r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void dt_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// K-major, no-swizzle shared-memory matrix descriptor: LBO = byte distance between core matrices adjacent in K,
// SBO = between 8-row groups (conventions verified on hardware by tools/tc_probe.cu, profiles/r1_tcgen05_probe.txt)
__device__ __forceinline__ uint64_t dt_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3fff) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ uint32_t dt_idesc(int n) {   // F16 x F16 -> F32, A and B K-major, M = 128
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
// x = hi + 2^-11 lo' (two fp16, x0 in the low half)
__device__ __forceinline__ void dt_split2(float x0, float x1, uint32_t &hi, uint32_t &lo) {
    const __half2 h = __floats2half2_rn(x0, x1);
    const float2 hf = __half22float2(h);
    const float2 d = rt_fmul2(rt_fadd2(make_float2(x0, x1), make_float2(-hf.x, -hf.y)), make_float2(LT_LO, LT_LO));
    const __half2 l = __floats2half2_rn(d.x, d.y);
    hi = *reinterpret_cast<const uint32_t *>(&h);
    lo = *reinterpret_cast<const uint32_t *>(&l);
}
__device__ __forceinline__ void dt_split8(const float *v, float s, uint4 &hi, uint4 &lo) {
    const float2 s2 = make_float2(s, s);
    const float2 a = rt_fmul2(make_float2(v[0], v[1]), s2), b = rt_fmul2(make_float2(v[2], v[3]), s2);
    const float2 c = rt_fmul2(make_float2(v[4], v[5]), s2), d = rt_fmul2(make_float2(v[6], v[7]), s2);
    dt_split2(a.x, a.y, hi.x, lo.x);
    dt_split2(b.x, b.y, hi.y, lo.y);
    dt_split2(c.x, c.y, hi.z, lo.z);
    dt_split2(d.x, d.y, hi.w, lo.w);
}
// power-of-two scale that brings an absolute maximum (its fp32 bit pattern) to [2^14, 2^15); 1 for a null pointer
__device__ __forceinline__ float dt_scale_from_amax(const float *amax) {
    if (amax == nullptr) return 1.0f;
    const uint32_t e = (__float_as_uint(__ldg(amax)) >> 23) & 0xffu;
    if (e == 0 || e == 255) return 1.0f;                  // all zero (or not finite: the result will say so)
    int f = 268 - (int)e;                                 // exponent field of 2^(14 - (e - 127))
    f = f > 253 ? 253 : (f < 1 ? 1 : f);                  // keep the scale and its inverse normal
    return __uint_as_float((uint32_t)f << 23);
}

struct LinTcArgs {
    long long rows;
    int k, n;             // real sizes
    int kp, np, n_tiles;  // K padded to 16; width of one N tile (multiple of 16, <= 128); number of N tiles
    const float *x;
    long long ldx;
    const float *w;       // W'[n][k] = w[n * w_sn + k * w_sk]
    long long w_sn, w_sk;
    const float *bias;    // n entries or null
    const float *x_amax;  // device scalar: max |x| (null: x is used unscaled)
    const float *w_amax;  // device scalar: max |w| (null: the weights are staged as 2^10 w)
    float *y;
    long long ldy;
    int act;              // 0 none, 1 ReLU, 2 LeakyReLU(0.1) on the output
};

__global__ void __launch_bounds__(LT_THREADS, 1) lin_tc_kernel(LinTcArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int plane_w = a.kp * a.np * 2;                  // bytes of one weight plane
    uint8_t *s_ring = smem;
    uint8_t *s_whi = smem + LT_NST * LT_STAGE, *s_wlo = s_whi + plane_w;
    float *s_bias = reinterpret_cast<float *>(s_wlo + plane_w);
    uint64_t *bar_full = reinterpret_cast<uint64_t *>(s_bias + 128);   // [NST] a stage of A planes has been written
    uint64_t *bar_empty = bar_full + LT_NST;                           // [NST] ... and read by the MMAs
    uint64_t *bar_dfull = bar_empty + LT_NST;                          // [2] accumulators of a tile are complete
    uint64_t *bar_dempty = bar_dfull + 2;                              // [2] ... and have been read by the epilogue
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar_dempty + 2);

    const int n_tile = blockIdx.x % a.n_tiles, n0 = n_tile * a.np;
    const long long tile0 = blockIdx.x / a.n_tiles, tstep = gridDim.x / a.n_tiles;
    const long long ntiles = (a.rows + LT_ROWS - 1) / LT_ROWS;
    const int ksteps = a.kp / 16, nchunks = (a.kp + LT_KC - 1) / LT_KC;
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < 4 * a.np) tmem_cols <<= 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < LT_NST; ++s) {
            rt_mbar_init(&bar_full[s], LT_LOAD_WARPS);
            rt_mbar_init(&bar_empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            rt_mbar_init(&bar_dfull[b], 1);
            rt_mbar_init(&bar_dempty[b], LT_EPI_WARPS);
        }
        rt_fence_mbar_init();
    }
    const float ws = a.w_amax ? dt_scale_from_amax(a.w_amax) : LT_WSCALE;
    // this CTA's weight tile -> fp16 hi/lo planes, [kg = k/8][n/8][8 rows][8 halfs] (LBO = np * 16, SBO = 128)
    for (int i = threadIdx.x; i < a.np * (a.kp / 8); i += LT_THREADS) {
        const int nl = i % a.np, kg = i / a.np, n = n0 + nl;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int k = 8 * kg + j;
            v[j] = (n < a.n && k < a.k) ? __ldg(a.w + n * a.w_sn + k * a.w_sk) : 0.0f;
        }
        uint4 hi, lo;
        dt_split8(v, ws, hi, lo);
        const int off = kg * (a.np * 16) + (nl >> 3) * 128 + (nl & 7) * 16;
        *reinterpret_cast<uint4 *>(s_whi + off) = hi;
        *reinterpret_cast<uint4 *>(s_wlo + off) = lo;
    }
    for (int i = threadIdx.x; i < a.np; i += LT_THREADS) s_bias[i] = (a.bias && n0 + i < a.n) ? __ldg(a.bias + n0 + i) : 0.0f;
    rt_fence_proxy_async();   // generic-proxy writes of the weight planes -> visible to the tensor core (async proxy)
    if (warp == LT_EPI_WARPS + LT_LOAD_WARPS) dt_tmem_alloc(tmem_slot, tmem_cols);
    dt_fence_before();
    __syncthreads();
    dt_fence_after();
    const uint32_t tm = *tmem_slot;
    const float xs = dt_scale_from_amax(a.x_amax);

    if (warp == LT_EPI_WARPS + LT_LOAD_WARPS) {
        // ===== MMA issuer =====
        if (lane == 0) {
            const uint32_t idesc = dt_idesc(a.np), lbo_w = (uint32_t)a.np * 16u;
            const uint32_t whi = rt_smem_u32(s_whi), wlo = rt_smem_u32(s_wlo), ring = rt_smem_u32(s_ring);
            uint32_t it = 0, tcount = 0;
            for (long long tile = tile0; tile < ntiles; tile += tstep, ++tcount) {
                const uint32_t buf = tcount & 1u;
                dt_spin(&bar_dempty[buf], ((tcount >> 1) & 1u) ^ 1u);   // the epilogue has drained this accumulator pair
                dt_fence_after();
                const uint32_t tM = tm + buf * 2u * (uint32_t)a.np, tC = tM + (uint32_t)a.np;
                for (int kc = 0; kc < nchunks; ++kc, ++it) {
                    const uint32_t s = it % LT_NST;
                    dt_spin(&bar_full[s], (it / LT_NST) & 1u);
                    dt_fence_after();
                    const uint32_t st = ring + s * LT_STAGE;
                    const int steps = min(LT_KC / 16, ksteps - kc * (LT_KC / 16));
                    for (int j = 0; j < steps; ++j) {
                        const uint32_t kk = (uint32_t)(kc * (LT_KC / 16) + j);
                        const uint64_t a_hi = dt_desc(st + j * 2 * LT_LBO, LT_LBO, 128), a_lo = dt_desc(st + LT_PLANE + j * 2 * LT_LBO, LT_LBO, 128);
                        const uint64_t b_hi = dt_desc(whi + kk * 2u * lbo_w, lbo_w, 128), b_lo = dt_desc(wlo + kk * 2u * lbo_w, lbo_w, 128);
                        dt_mma_ss(tC, a_lo, b_hi, idesc, kk > 0);
                        dt_mma_ss(tC, a_hi, b_lo, idesc, 1);
                        dt_mma_ss(tM, a_hi, b_hi, idesc, kk > 0);
                    }
                    dt_commit(&bar_empty[s]);   // the stage is free once these MMAs have read it
                }
                dt_commit(&bar_dfull[buf]);
            }
        }
        __syncwarp();
    } else if (warp >= LT_EPI_WARPS) {
        // ===== loader warps: rows -> fp16 hi/lo planes in the K-major core-matrix layout =====
        // The (tile, K chunk) items of this CTA form one flat sequence; the loads of item i+1 are in flight while item i is
        // converted and stored (two register sets).  ncu on the first version (8 warps, generic addressing): 92 % of the
        // kernel's instructions were the loaders', 115 per 32-byte task, two warps per scheduler -> instruction-latency bound
        // at a third of the HBM rate.  Hence 16 warps and a lean path (whole tile in range, 16-byte aligned rows, full K chunks)
        // whose per-task cost is one pointer add, one 128-bit load, the split of four values and two 64-bit stores.
        const int lw = warp - LT_EPI_WARPS;
        const bool al16 = (a.ldx & 3) == 0 && (reinterpret_cast<uintptr_t>(a.x) & 15) == 0;
        const bool lean_k = al16 && (a.k % LT_KC) == 0;
        // a warp task = one 128-bit load instruction = 4 rows x 128 contiguous bytes (8 lanes per row: one full line per
        // quarter warp -- the L1 data pipe moves 128 bytes per wavefront; the first version's 256-bit loads with 4 lanes per
        // row cost a wavefront per 32-byte sector and held that pipe at 77 %)
        int t_r[LT_TASKS], t_k[LT_TASKS], t_off[LT_TASKS];
        long long t_src[LT_TASKS];
#pragma unroll
        for (int i = 0; i < LT_TASKS; ++i) {
            const int wt = lw + LT_LOAD_WARPS * i, seg = wt & 1, j = lane & 7;
            t_r[i] = 4 * (wt >> 1) + (lane >> 3);
            t_k[i] = seg * 32 + 4 * j;
            t_off[i] = (seg * 4 + (j >> 1)) * LT_LBO + (t_r[i] >> 3) * 128 + (t_r[i] & 7) * 16 + (j & 1) * 8;
            t_src[i] = (long long)t_r[i] * a.ldx + t_k[i];
        }
        auto issue = [&](long long tile, int kc, float4 *v) {
            const long long row0 = tile * LT_ROWS;
            if (lean_k && row0 + LT_ROWS <= a.rows) {
                const float *p = a.x + row0 * a.ldx + kc * LT_KC;
#pragma unroll
                for (int i = 0; i < LT_TASKS; ++i) v[i] = __ldg(reinterpret_cast<const float4 *>(p + t_src[i]));
            } else {
#pragma unroll
                for (int i = 0; i < LT_TASKS; ++i) {
                    const long long row = row0 + t_r[i];
                    const int kcol = kc * LT_KC + t_k[i];
                    const float *p = a.x + row * a.ldx + kcol;
                    if (row >= a.rows || kcol >= a.k) v[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                    else if (al16 && kcol + 4 <= a.k) v[i] = __ldg(reinterpret_cast<const float4 *>(p));
                    else v[i] = make_float4(__ldg(p), kcol + 1 < a.k ? __ldg(p + 1) : 0.0f, kcol + 2 < a.k ? __ldg(p + 2) : 0.0f,
                                            kcol + 3 < a.k ? __ldg(p + 3) : 0.0f);
                }
            }
        };
        const float2 xs2 = make_float2(xs, xs);
        auto convert = [&](uint32_t it, const float4 *v) {
            const uint32_t s = it % LT_NST;
            dt_wait(&bar_empty[s], ((it / LT_NST) & 1u) ^ 1u);
            uint8_t *st = s_ring + s * LT_STAGE;
#pragma unroll
            for (int i = 0; i < LT_TASKS; ++i) {
                const float2 p0 = rt_fmul2(make_float2(v[i].x, v[i].y), xs2), p1 = rt_fmul2(make_float2(v[i].z, v[i].w), xs2);
                uint2 hi, lo;
                dt_split2(p0.x, p0.y, hi.x, lo.x);
                dt_split2(p1.x, p1.y, hi.y, lo.y);
                *reinterpret_cast<uint2 *>(st + t_off[i]) = hi;
                *reinterpret_cast<uint2 *>(st + LT_PLANE + t_off[i]) = lo;
            }
            rt_fence_proxy_async();
            __syncwarp();
            if (lane == 0) rt_mbar_arrive(&bar_full[s]);
        };
        auto advance = [&](long long &tile, int &kc) {
            if (++kc == nchunks) { kc = 0; tile += tstep; }
        };
        float4 va[LT_TASKS], vb[LT_TASKS];
        long long tile = tile0;
        int kc = 0;
        if (tile < ntiles) issue(tile, kc, va);
        for (uint32_t it = 0; tile < ntiles; it += 2) {   // two items per trip: the register sets swap roles without copies
            advance(tile, kc);
            if (tile < ntiles) issue(tile, kc, vb);
            convert(it, va);
            if (tile >= ntiles) break;
            advance(tile, kc);
            if (tile < ntiles) issue(tile, kc, va);
            convert(it + 1, vb);
        }
    } else {
        // ===== epilogue warps: one thread per row (TMEM lane) =====
        const uint32_t lane_base = (uint32_t)(32 * warp) << 16;
        const float inv_x = 1.0f / xs, inv_w = 1.0f / ws;   // powers of two: two exact multiplies, no overflow of xs * ws
        const float slope = a.act == 1 ? 0.0f : (a.act == 2 ? 0.1f : 1.0f);
        const bool st256 = (a.ldy & 7) == 0 && (reinterpret_cast<uintptr_t>(a.y) & 31) == 0 && (n0 & 7) == 0;
        uint32_t tcount = 0;
        for (long long tile = tile0; tile < ntiles; tile += tstep, ++tcount) {
            const uint32_t buf = tcount & 1u;
            dt_wait(&bar_dfull[buf], (tcount >> 1) & 1u);
            dt_fence_after();
            const uint32_t tM = tm + buf * 2u * (uint32_t)a.np + lane_base, tC = tM + (uint32_t)a.np;
            const long long row = tile * LT_ROWS + 32 * warp + lane;
            float *yrow = a.y + row * a.ldy + n0;
            for (int c = 0; c < a.np; c += 16) {
                uint32_t m[16], cr[16];
                dt_ld16(tM + c, m);
                dt_ld16(tC + c, cr);
                dt_ld_wait();
                float v[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float t = fmaf(fmaf(__uint_as_float(cr[i]), 1.0f / LT_LO, __uint_as_float(m[i])) * inv_x, inv_w, s_bias[c + i]);
                    v[i] = fmaxf(t, t * slope);
                }
                if (row < a.rows) {
                    if (st256 && n0 + c + 16 <= a.n) {
                        rt_stg256(yrow + c, make_float4(v[0], v[1], v[2], v[3]), make_float4(v[4], v[5], v[6], v[7]));
                        rt_stg256(yrow + c + 8, make_float4(v[8], v[9], v[10], v[11]), make_float4(v[12], v[13], v[14], v[15]));
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (n0 + c + i < a.n) yrow[c + i] = v[i];
                    }
                }
            }
            dt_fence_before();
            __syncwarp();
            if (lane == 0) rt_mbar_arrive(&bar_dempty[buf]);
        }
    }
    dt_fence_before();
    __syncthreads();
    if (warp == LT_EPI_WARPS + LT_LOAD_WARPS) dt_tmem_dealloc(tm, tmem_cols);
}

// ---- wgrad ----------------------------------------------------------------------------------------------------------
constexpr int WG_RCH = 64;                       // rows per stage = 4 K steps
constexpr int WG_NST = 2;
constexpr int WG_MAX_CHAIN = 16384;               // rows per accumulation chain (1024 K steps)
constexpr int WG_LOAD_WARPS = 16;
constexpr int WG_TASKS = 6;                     // warp tasks per loader warp and stage: 2 of dy (32 / 16) + up to 4 of x (64 / 16)
constexpr int WG_THREADS = 32 * (WG_LOAD_WARPS + 1);
constexpr int WG_APLANE = 128 * WG_RCH * 2;      // 16 KB

struct WgradArgs {
    long long rows, rows_per_split;
    int n, k;                 // channels of dy / of x
    int m_tiles, k_tiles, ktw;   // 128-channel tiles of dy; tiles of x of width ktw (multiple of 16, <= 256)
    const float *dy;
    long long lddy;
    const float *x;
    long long ldx;
    const float *dy_amax, *x_amax;   // device scalars max |dy|, max |x| (null: unscaled)
    float *part;              // [splits][m_tiles * 128][k_tiles * ktw]
};

__global__ void __launch_bounds__(WG_THREADS, 1) wgrad_tc_kernel(WgradArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bplane = a.ktw * WG_RCH * 2, stage_bytes = 2 * WG_APLANE + 2 * bplane;
    uint64_t *bar_full = reinterpret_cast<uint64_t *>(smem + WG_NST * stage_bytes);
    uint64_t *bar_empty = bar_full + WG_NST;
    uint64_t *bar_done = bar_empty + WG_NST;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar_done + 1);

    const int mt = blockIdx.y / a.k_tiles, kt = blockIdx.y % a.k_tiles;
    const int m0 = mt * 128, k0 = kt * a.ktw;
    const long long r_begin = (long long)blockIdx.x * a.rows_per_split;
    const long long r_end = min(a.rows, r_begin + a.rows_per_split);
    const int nstages = (int)((r_end - r_begin + WG_RCH - 1) / WG_RCH);   // >= 1 by construction of the grid
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < 2 * a.ktw) tmem_cols <<= 1;
    const uint32_t corr_off = tmem_cols / 2;

    if (threadIdx.x == 0) {
        for (int s = 0; s < WG_NST; ++s) {
            rt_mbar_init(&bar_full[s], WG_LOAD_WARPS);
            rt_mbar_init(&bar_empty[s], 1);
        }
        rt_mbar_init(bar_done, 1);
        rt_fence_mbar_init();
    }
    if (warp == WG_LOAD_WARPS) dt_tmem_alloc(tmem_slot, tmem_cols);
    dt_fence_before();
    __syncthreads();
    dt_fence_after();
    const uint32_t tm = *tmem_slot;
    const float ds = dt_scale_from_amax(a.dy_amax), xsc = dt_scale_from_amax(a.x_amax);

    if (warp == WG_LOAD_WARPS) {
        if (lane == 0) {
            const uint32_t idesc = dt_idesc(a.ktw), lbo_b = (uint32_t)a.ktw * 16u;
            const uint32_t base = rt_smem_u32(smem);
            for (int it = 0; it < nstages; ++it) {
                const uint32_t s = (uint32_t)it % WG_NST;
                dt_spin(&bar_full[s], ((uint32_t)it / WG_NST) & 1u);
                dt_fence_after();
                const uint32_t sa = base + s * stage_bytes, sb = sa + 2 * WG_APLANE;
#pragma unroll
                for (int j = 0; j < WG_RCH / 16; ++j) {
                    const uint64_t a_hi = dt_desc(sa + j * 4096, 2048, 128), a_lo = dt_desc(sa + WG_APLANE + j * 4096, 2048, 128);
                    const uint64_t b_hi = dt_desc(sb + j * 2u * lbo_b, lbo_b, 128), b_lo = dt_desc(sb + bplane + j * 2u * lbo_b, lbo_b, 128);
                    const uint32_t acc = (it | j) > 0;
                    dt_mma_ss(tm + corr_off, a_lo, b_hi, idesc, acc);
                    dt_mma_ss(tm + corr_off, a_hi, b_lo, idesc, 1);
                    dt_mma_ss(tm, a_hi, b_hi, idesc, acc);
                }
                dt_commit(&bar_empty[s]);
            }
            dt_commit(bar_done);
        }
        __syncwarp();
    } else {
        // ===== loaders: thread = one channel x 8 consecutive rows -> one 16-byte core-matrix row (K = rows) =====
        // warp task = 32 channels x 8 rows (8 coalesced 128-byte reads).  Per stage: 32 tasks of the dy operand (128 channels x
        // 8 row groups: tasks u = 0, 1 of each of the 16 warps) and up to 64 of the x operand (u = 2 .. 5); all of a warp's
        // loads of a stage are issued before the first value is converted.  ncu on the first version: 22 instructions per
        // load in generic addressing and predicates (issue-bound at 72 %), hence the lean path for stages and channel blocks
        // that are entirely in range: a pointer walk, no predicates.
        const int nbt = (a.ktw + 31) / 32 * 8;
        int t_c[WG_TASKS], t_g[WG_TASKS], t_off[WG_TASKS];
        bool t_live[WG_TASKS], t_full[WG_TASKS], t_ok[WG_TASKS];
        const float *t_src[WG_TASKS];
#pragma unroll
        for (int u = 0; u < WG_TASKS; ++u) {
            const bool isa = u < 2;
            const int t2 = warp + WG_LOAD_WARPS * (isa ? u : u - 2);
            const int c = (t2 >> 3) * 32 + lane, g = t2 & 7;
            const int cb = (t2 >> 3) * 32;
            t_c[u] = c;
            t_g[u] = g;
            t_live[u] = isa ? m0 + cb < a.n : t2 < nbt;   // a 32-channel block of dy beyond n stays zero (cleared once below)
            const int ch = (isa ? m0 : k0) + c;
            t_ok[u] = t_live[u] && (isa ? ch < a.n : (c < a.ktw && ch < a.k));
            t_full[u] = t_live[u] && (isa ? m0 + cb + 32 <= a.n : (cb + 32 <= a.ktw && k0 + cb + 32 <= a.k));   // warp-uniform
            t_off[u] = g * (isa ? 2048 : a.ktw * 16) + (c >> 3) * 128 + (c & 7) * 16;
            t_src[u] = (isa ? a.dy : a.x) + (long long)(8 * g) * (isa ? a.lddy : a.ldx) + ch;
        }
        if (m0 + 128 > a.n) {   // some dy blocks are dead: their planes must read as zeros
            for (int st_ = 0; st_ < WG_NST; ++st_)
                for (int i = threadIdx.x; i < 2 * WG_APLANE / 16; i += 32 * WG_LOAD_WARPS)
                    reinterpret_cast<uint4 *>(smem + st_ * stage_bytes)[i] = make_uint4(0, 0, 0, 0);
            rt_fence_proxy_async();
            asm volatile("bar.sync 1, %0;" ::"n"(32 * WG_LOAD_WARPS) : "memory");   // loader warps only
        }
        for (int it = 0; it < nstages; ++it) {
            const uint32_t s = (uint32_t)it % WG_NST;
            const long long r0 = r_begin + (long long)it * WG_RCH;
            const bool rows_full = r0 + WG_RCH <= r_end;
            float v[WG_TASKS][8];
#pragma unroll
            for (int u = 0; u < WG_TASKS; ++u) {
                const long long ld = u < 2 ? a.lddy : a.ldx;
                const float *p = t_src[u] + r0 * ld;
                if (rows_full && t_full[u]) {
#pragma unroll
                    for (int i = 0; i < 8; ++i, p += ld) v[u][i] = __ldg(p);
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i, p += ld) v[u][i] = (t_ok[u] && r0 + 8 * t_g[u] + i < r_end) ? __ldg(p) : 0.0f;
                }
            }
            dt_wait(&bar_empty[s], (((uint32_t)it / WG_NST) & 1u) ^ 1u);
            uint8_t *sa = smem + s * stage_bytes, *sb = sa + 2 * WG_APLANE;
#pragma unroll
            for (int u = 0; u < WG_TASKS; ++u) {
                const bool isa = u < 2;
                if (!t_live[u] || (!isa && t_c[u] >= a.ktw)) continue;
                uint4 hi, lo;
                dt_split8(v[u], isa ? ds : xsc, hi, lo);
                uint8_t *dst = (isa ? sa : sb) + t_off[u];
                *reinterpret_cast<uint4 *>(dst) = hi;
                *reinterpret_cast<uint4 *>(dst + (isa ? WG_APLANE : bplane)) = lo;
            }
            rt_fence_proxy_async();
            __syncwarp();
            if (lane == 0) rt_mbar_arrive(&bar_full[s]);
        }
        // ===== epilogue (warps 0..3): lane = dy channel, columns = x channels =====
        if (warp < 4) {
            dt_wait(bar_done, 0);
            dt_fence_after();
            const uint32_t lane_base = (uint32_t)(32 * warp) << 16;
            const float inv = 1.0f / ds, inv2 = 1.0f / xsc;
            const long long NP = (long long)a.m_tiles * 128, KP = (long long)a.k_tiles * a.ktw;
            float *prow = a.part + ((long long)blockIdx.x * NP + m0 + 32 * warp + lane) * KP + k0;
            for (int c = 0; c < a.ktw; c += 16) {
                uint32_t m[16], cr[16];
                dt_ld16(tm + lane_base + c, m);
                dt_ld16(tm + corr_off + lane_base + c, cr);
                dt_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; i += 4) {
                    float4 o;
                    o.x = fmaf(__uint_as_float(cr[i + 0]), 1.0f / LT_LO, __uint_as_float(m[i + 0])) * inv * inv2;
                    o.y = fmaf(__uint_as_float(cr[i + 1]), 1.0f / LT_LO, __uint_as_float(m[i + 1])) * inv * inv2;
                    o.z = fmaf(__uint_as_float(cr[i + 2]), 1.0f / LT_LO, __uint_as_float(m[i + 2])) * inv * inv2;
                    o.w = fmaf(__uint_as_float(cr[i + 3]), 1.0f / LT_LO, __uint_as_float(m[i + 3])) * inv * inv2;
                    *reinterpret_cast<float4 *>(prow + c + i) = o;
                }
            }
        }
    }
    dt_fence_before();
    __syncthreads();
    if (warp == WG_LOAD_WARPS) dt_tmem_dealloc(tm, tmem_cols);
}

// dw[n][k] = sum over the splits in a FIXED order (bit-repeatable): a CTA owns 32 consecutive output elements, warp g adds
// the splits g, g + 8, ... in increasing order (coalesced 128-byte reads), then the 8 warps' sums are added in warp order.
// (One thread per element walked up to a few hundred partials serially: 23 us per layer, 3.5 % of the training step.)
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float *part, int splits, long long NP, long long KP, int n, int k, float *dw) {
    __shared__ float s_sum[8][32];
    const int g = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long i = (long long)blockIdx.x * 32 + lane;
    float s = 0.0f;
    if (i < (long long)n * k) {
        const int r = (int)(i / k), c = (int)(i % k);
        for (int sp = g; sp < splits; sp += 8) s += part[((long long)sp * NP + r) * KP + c];
    }
    s_sum[g][lane] = s;
    __syncthreads();
    if (g == 0 && i < (long long)n * k) {
        float t = s_sum[0][lane];
#pragma unroll
        for (int u = 1; u < 8; ++u) t += s_sum[u][lane];
        dw[i] = t;
    }
}

// max |x| as an fp32 bit pattern (non-negative floats order like unsigned integers)
__global__ void absmax_kernel(const float *x, long long count, unsigned int *out) {
    float m = 0.0f;
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if ((reinterpret_cast<uintptr_t>(x) & 15) == 0) {
        const long long n4 = count >> 2;
        const float4 *x4 = reinterpret_cast<const float4 *>(x);
        for (long long j = i; j < n4; j += stride) {
            const float4 v = __ldg(x4 + j);
            m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
        }
        for (long long j = (n4 << 2) + i; j < count; j += stride) m = fmaxf(m, fabsf(__ldg(x + j)));
    } else {
        for (long long j = i; j < count; j += stride) m = fmaxf(m, fabsf(__ldg(x + j)));
    }
    const uint32_t r = rt_redux_max_u32(__float_as_uint(m));
    if ((threadIdx.x & 31) == 0 && r) atomicMax(out, r);
}

int dt_sm_count() {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms;
}

}  // namespace

RT_API int rt_absmax(const float *x, long long count, float *amax_out, void *stream) {
    RT_REQUIRE(x && amax_out && count >= 0, "rt_absmax: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(amax_out, 0, 4, st);
    if (count == 0) return RT_OK;
    const int blocks = (int)min((long long)dt_sm_count() * 8, (count + 4095) / 4096);
    absmax_kernel<<<blocks, 256, 0, st>>>(x, count, reinterpret_cast<unsigned int *>(amax_out));
    return rt_check_launch("absmax_kernel");
}

RT_API int rt_dense_tc_forward(long long rows, int k, int n, const float *x, long long ldx, const float *w, long long w_sn, long long w_sk,
                               const float *bias, const float *x_amax, const float *w_amax, int act, float *y, long long ldy, void *stream) {
    RT_REQUIRE(rows >= 0 && k >= 1 && n >= 1 && x && w && y, "rt_dense_tc_forward: bad arguments");
    RT_REQUIRE(ldx >= k && ldy >= n, "rt_dense_tc_forward: leading dimensions smaller than the row width");
    RT_REQUIRE(act >= 0 && act <= 2, "rt_dense_tc_forward: act must be 0 (none), 1 (ReLU) or 2 (LeakyReLU 0.1)");
    if (rows == 0) return RT_OK;
    const int kp = (k + 15) / 16 * 16;
    // N tile: a multiple of 16, at most 128 columns, and its two weight planes (4 kp np bytes) within 128 KB
    const int n16 = (n + 15) / 16 * 16;
    int cap = min(128, (128 * 1024) / (4 * kp) / 16 * 16);
    if (cap < 16) {
        rt_set_error("rt_dense_tc_forward: K = %d does not fit the resident-weight kernel (K <= 2048)", k);
        return RT_ERR_UNSUPPORTED;
    }
    const int n_tiles = (n16 + cap - 1) / cap;
    const int np = ((n16 + n_tiles - 1) / n_tiles + 15) / 16 * 16;
    size_t smem = (size_t)LT_NST * LT_STAGE + (size_t)4 * kp * np + 128 * 4 + 16 * 8;
    if (smem < 120 * 1024) smem = 120 * 1024;   // one CTA per SM, whatever the shape: co-resident CTAs would wait on each other's TMEM
    static RtPerDevice attr_set;
    if (!attr_set.done(rt_current_device())) {
        const cudaError_t e = cudaFuncSetAttribute(lin_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) {
            rt_set_error("rt_dense_tc_forward: cannot reserve shared memory: %s", cudaGetErrorString(e));
            return (int)e;
        }
        attr_set.mark(rt_current_device());
    }
    const long long ntiles = (rows + LT_ROWS - 1) / LT_ROWS;
    long long grid_rows = dt_sm_count() / n_tiles;
    if (grid_rows < 1) grid_rows = 1;
    if (grid_rows > ntiles) grid_rows = ntiles;
    LinTcArgs a{rows, k, n, kp, np, n_tiles, x, ldx, w, w_sn, w_sk, bias, x_amax, w_amax, y, ldy, act};
    lin_tc_kernel<<<(unsigned)(grid_rows * n_tiles), LT_THREADS, smem, (cudaStream_t)stream>>>(a);
    return rt_check_launch("lin_tc_kernel");
}

RT_API int rt_dense_tc_wgrad(long long rows, int n, int k, const float *dy, long long lddy, const float *x, long long ldx,
                             const float *dy_amax, const float *x_amax, float *dw, void *stream) {
    RT_REQUIRE(rows >= 0 && k >= 1 && n >= 1 && dy && x && dw, "rt_dense_tc_wgrad: bad arguments");
    RT_REQUIRE(lddy >= n && ldx >= k, "rt_dense_tc_wgrad: leading dimensions smaller than the row width");
    cudaStream_t st = (cudaStream_t)stream;
    if (rows == 0) {
        cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)n * k, st);
        return RT_OK;
    }
    const int kp = (k + 15) / 16 * 16;
    const int k_tiles = (kp + 255) / 256;
    const int ktw = ((kp + k_tiles - 1) / k_tiles + 15) / 16 * 16;
    const int m_tiles = (n + 127) / 128;
    const int tiles = m_tiles * k_tiles;
    // Split of the rows: at least one CTA per SM, and no accumulation chain longer than WG_MAX_CHAIN rows -- the tensor core
    // truncates its fp32 accumulator at every K step (16 rows), so the error of a chain grows linearly with its length
    // (measured: 1.9e-5 of max |dW| at 14 k rows per CTA, 7e-5 at 57 k); the partial sums are added in fp32 round-to-nearest.
    long long splits = dt_sm_count() / tiles;
    if (splits < 1) splits = 1;
    const long long max_splits = (rows + 8 * WG_RCH - 1) / (8 * WG_RCH);   // at least 8 stages of work per CTA
    if (splits > max_splits) splits = max_splits;
    if ((rows + splits - 1) / splits > WG_MAX_CHAIN) {
        splits = (rows + WG_MAX_CHAIN - 1) / WG_MAX_CHAIN;
        const long long wave = dt_sm_count() / tiles > 0 ? dt_sm_count() / tiles : 1;
        splits = (splits + wave - 1) / wave * wave;                         // whole waves of CTAs
    }
    long long rps = ((rows + splits - 1) / splits + WG_RCH - 1) / WG_RCH * WG_RCH;
    splits = (rows + rps - 1) / rps;
    const size_t stage = 2 * WG_APLANE + (size_t)2 * ktw * WG_RCH * 2;
    size_t smem = WG_NST * stage + 8 * 8;
    if (smem < 120 * 1024) smem = 120 * 1024;
    static RtPerDevice attr_set;
    if (!attr_set.done(rt_current_device())) {
        const cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) {
            rt_set_error("rt_dense_tc_wgrad: cannot reserve shared memory: %s", cudaGetErrorString(e));
            return (int)e;
        }
        attr_set.mark(rt_current_device());
    }
    const long long NP = (long long)m_tiles * 128, KP = (long long)k_tiles * ktw;
    float *part = nullptr;
    int rc = rt_scratch_alloc((void **)&part, sizeof(float) * (size_t)(splits * NP * KP), st, "rt_dense_tc_wgrad");
    if (rc != RT_OK) return rc;
    WgradArgs a{rows, rps, n, k, m_tiles, k_tiles, ktw, dy, lddy, x, ldx, dy_amax, x_amax, part};
    wgrad_tc_kernel<<<dim3((unsigned)splits, (unsigned)tiles), WG_THREADS, smem, st>>>(a);
    rc = rt_check_launch("wgrad_tc_kernel");
    if (rc == RT_OK) {
        const long long total = (long long)n * k;
        wgrad_reduce_kernel<<<(unsigned)((total + 31) / 32), 256, 0, st>>>(part, (int)splits, NP, KP, n, k, dw);
        rc = rt_check_launch("wgrad_reduce_kernel");
    }
    rt_scratch_free(part, st);
    return rc;
}
