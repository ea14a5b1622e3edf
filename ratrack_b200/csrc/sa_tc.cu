// Set-abstraction scale on the tensor cores with the cloud's projected features resident in shared memory (sm_100a).
//
// One scale of PointnetSAModuleMSG (reference: QueryAndGroup + SharedMLP + max-pool, src/lib/pointnet2_utils.py:269-292,
// src/lib/pytorch_utils.py:5-32, src/lib/pointnet2_modules.py:35-47) whose SharedMLP has two convolutions:
//     h[row, :]  = relu(Y[idx[row], :] + Wx . (xyz_in[idx[row]] - centre) + b1)      first conv, projected per point
//     out[c, :]  = max over the NS rows of centre c of relu(W2 . h + b2)             second conv on tcgen05 + max-pool
// The generic row-tile kernel (mlp_tc.cu, gather mode) fetches Y rows from L2 with one thread per row: every neighbour
// row is read NS x (overlap of the balls) times through L1 (ncu: 70 % L1 data-pipe, long-scoreboard stalls on the
// index -> row chain).  A cloud has only n_in <= 1024 points, so its whole Y table (n_in x C1 fp32 = 64..128 KB), its
// xyz and the centres fit in shared memory: this kernel loads them once per cloud segment and gathers from there.
//   grid     = persistent, one CTA per SM; CTA b owns a contiguous range of the (cloud, tile) list: at most 2-3 table loads;
//   CTA      = 4 independent tile slots x 4 warps (one per TMEM lane quarter); a slot owns C1 + C2 TMEM columns and one
//              mbarrier; the slot's first thread issues its MMAs (no separate issuer warp: one named barrier per tile);
//   shapes   = compile-time (C1, C2, NS): no address arithmetic or loop control left in the inner loops;
//   numerics = identical to mlp_tc.cu: fp16 hi/lo split of activations and 2^10-scaled weights, corrections first, then
//              the main products with scale-input-d (tools/tc_precision.cu).
#include <cuda_fp16.h>

#include "engine_kernels.cuh"
#include "mlp_tc.cuh"

namespace {

constexpr int SA_SLOTS = 4;
constexpr int SA_THREADS = 128 * SA_SLOTS;
constexpr float SA_WINV = 1.0f / 1024.0f;
constexpr float SA_LO_SCALE = 2048.0f;

struct SaTcArgs {
    int clouds, npts, n_in;      // centres per cloud, candidate points per cloud
    const float *y;              // [clouds * n_in, ldy], this scale's columns start at yoff
    int ldy, yoff;
    const int *idx;              // [clouds * npts * NS]
    const float *xyz_in, *xyz_c; // [clouds * n_in, 3], [clouds * npts, 3]
    const float *wx, *b1;        // [C1, 3], [C1]
    const void *wpack;           // second conv, fp16 hi/lo planes of 2^10 W in core-matrix layout
    const float *b2;
    const void *wpack3;          // optional third conv (C3 > 0) and its bias
    const float *b3;
    float *out;                  // out[(cloud * npts + centre) * ldo + ooff + c]
    int ldo, ooff;
    int *status;
};

__device__ __forceinline__ void sa_mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = rt_smem_u32(bar);
    for (uint32_t it = 0;; ++it) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
        if (ok) return;
        __nanosleep(32);
        if (it > (1u << 22)) __trap();
    }
}
__device__ __forceinline__ void sa_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void sa_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void sa_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(rt_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void sa_mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void sa_mma_ts_rescale(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p, 11;\n\t}"
                 ::"r"(d), "r"(a), "l"(b), "r"(idesc) : "memory");
}
__device__ __forceinline__ void sa_ld16(uint32_t taddr, uint32_t *r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void sa_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void sa_st8(uint32_t taddr, const uint32_t *r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void sa_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint64_t sa_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3fff) << 32) | ((uint64_t)1 << 46);
}
constexpr uint32_t sa_idesc(int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); }

__device__ __forceinline__ void sa_split2(float2 x, uint32_t &hi, uint32_t &lo, __half2 &amax) {
    const __half2 h = __floats2half2_rn(x.x, x.y);
    amax = __hmax2(amax, __habs2(h));
    const float2 hf = __half22float2(h);
    const float2 r = rt_fmul2(rt_fadd2(x, make_float2(-hf.x, -hf.y)), make_float2(SA_LO_SCALE, SA_LO_SCALE));
    const __half2 l = __floats2half2_rn(r.x, r.y);
    hi = *reinterpret_cast<const uint32_t *>(&h);
    lo = *reinterpret_cast<const uint32_t *>(&l);
}

// max over the NS consecutive lanes of a centre for 16 values per lane (halving butterfly, see mlp_tc.cu)
template <int NS>
__device__ __forceinline__ void sa_group_max(float *v, int lane, int &col0, int &cnt, bool &writer) {
    col0 = 0;
    writer = true;
    constexpr int STEPS = NS == 32 ? 5 : NS == 16 ? 4 : NS == 8 ? 3 : NS == 4 ? 2 : NS == 2 ? 1 : 0;
    int c = 16;
#pragma unroll
    for (int st = 0; st < STEPS; ++st) {
        const int off = NS >> (st + 1);
        const int half = 16 >> (st + 1);
        if (half >= 1) {
            const bool up = (lane & off) != 0;
#pragma unroll
            for (int i = 0; i < half; ++i) {
                const float send = up ? v[i] : v[i + half], keep = up ? v[i + half] : v[i];
                v[i] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, off));
            }
            col0 += up ? half : 0;
            c = half;
        } else {
            v[0] = fmaxf(v[0], __shfl_xor_sync(0xffffffffu, v[0], off));
            writer = writer && (lane & off) == 0;
        }
    }
    cnt = c;
}

template <int C1>
constexpr int sa_ystride() { return C1 + 4; }   // floats; = 4 (mod 32): 8 lanes reading 16 B of rows r..r+7 hit 32 distinct banks

template <int C1, int C2, int C3>
constexpr int sa_slot_cols() {   // TMEM columns of one tile slot: accumulator + A planes, a power of two so 4 slots tile the 512
    int need = (C2 > C3 ? C2 : C3) + (C1 > C2 || C3 == 0 ? C1 : C2), c = 32;
    while (c < need) c *= 2;
    return c;
}

template <int C1, int C2, int C3, int NS>
__global__ void __launch_bounds__(SA_THREADS, 1) sa_tc_kernel(SaTcArgs a) {
    constexpr int NS_SHIFT = NS == 32 ? 5 : NS == 16 ? 4 : NS == 8 ? 3 : NS == 4 ? 2 : NS == 2 ? 1 : 0;
    constexpr int YS = sa_ystride<C1>();
    constexpr int SLOT_COLS = sa_slot_cols<C1, C2, C3>();
    constexpr int DCOLS = C2 > C3 ? C2 : C3;                       // accumulator columns
    constexpr int ACOLS = (C1 > C2 || C3 == 0) ? C1 : C2;          // A operand columns (hi plane + lo plane)
    constexpr int CL = C3 > 0 ? C3 : C2;                           // channels of the pooled output
    constexpr int WBYTES = 4 * C1 * C2, WBYTES3 = 4 * C2 * C3;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar_d[SA_SLOTS];
    __shared__ uint32_t tmem_slot;
    // carve-up: weights | gather constants | bias | xyz of the points | centres | Y table
    uint8_t *s_w = smem;
    uint8_t *s_w3 = smem + WBYTES;
    float4 *s_gc = reinterpret_cast<float4 *>(smem + WBYTES + WBYTES3);     // [C1 / 2][2]: {wx0,wx1,wy0,wy1} {wz0,wz1,b0,b1}
    float *s_b2 = reinterpret_cast<float *>(s_gc + C1);
    float *s_b3 = s_b2 + C2;
    float4 *s_xyz = reinterpret_cast<float4 *>(s_b3 + (C3 > 0 ? C3 : 4));
    float4 *s_ctr = s_xyz + a.n_in;
    float *s_y = reinterpret_cast<float *>(s_ctr + a.npts);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int slot = warp >> 2, q = warp & 3;
    const int rit = 32 * q + lane;   // row in tile = TMEM lane

    {   // one-time: weights, constants
        const uint4 *src = reinterpret_cast<const uint4 *>(a.wpack);
        uint4 *dst = reinterpret_cast<uint4 *>(s_w);
        for (int i = tid; i < WBYTES / 16; i += SA_THREADS) dst[i] = __ldg(src + i);
        if (C3 > 0) {
            const uint4 *src3 = reinterpret_cast<const uint4 *>(a.wpack3);
            uint4 *dst3 = reinterpret_cast<uint4 *>(s_w3);
            for (int i = tid; i < WBYTES3 / 16; i += SA_THREADS) dst3[i] = __ldg(src3 + i);
            for (int i = tid; i < C3; i += SA_THREADS) s_b3[i] = a.b3 ? __ldg(a.b3 + i) : 0.0f;
        }
        for (int i = tid; i < C1 / 2; i += SA_THREADS) {
            const int c = 2 * i;
            s_gc[2 * i] = make_float4(__ldg(a.wx + c * 3 + 0), __ldg(a.wx + c * 3 + 3), __ldg(a.wx + c * 3 + 1), __ldg(a.wx + c * 3 + 4));
            s_gc[2 * i + 1] = make_float4(__ldg(a.wx + c * 3 + 2), __ldg(a.wx + c * 3 + 5), __ldg(a.b1 + c), __ldg(a.b1 + c + 1));
        }
        for (int i = tid; i < C2; i += SA_THREADS) s_b2[i] = a.b2 ? __ldg(a.b2 + i) : 0.0f;
        if (tid == 0) {
            for (int s = 0; s < SA_SLOTS; ++s) rt_mbar_init(&bar_d[s], 1);
            rt_fence_mbar_init();
        }
    }
    rt_fence_proxy_async();   // s_w (generic-proxy stores) -> visible to the tensor core's async proxy
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(rt_smem_u32(&tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    sa_fence_before();
    __syncthreads();
    sa_fence_after();
    const uint32_t tm = tmem_slot;
    const uint32_t tD = tm + slot * SLOT_COLS, tAhi = tD + DCOLS, tAlo = tAhi + ACOLS / 2;
    const uint32_t lane_base = (uint32_t)(32 * q) << 16;

    const int T = (a.npts * NS) / 128;   // tiles per cloud (the launcher guarantees divisibility)
    const long long G = (long long)a.clouds * T;
    const long long lo = G * blockIdx.x / gridDim.x, hi = G * (blockIdx.x + 1) / gridDim.x;
    uint32_t d_phase = 0;
    __half2 amax = __floats2half2_rn(0.0f, 0.0f);

    // the slot's first index is fetched before the table load, later ones one tile ahead (t_next = the tile j_next belongs to)
    long long t_next = lo + slot;
    int j_next = t_next < hi ? __ldg(a.idx + t_next * 128 + rit) : 0;

    for (long long t0 = lo; t0 < hi;) {
        const int cloud = (int)(t0 / T);
        const long long seg_end = min(hi, (long long)(cloud + 1) * T);
        __syncthreads();   // every slot is done gathering from the previous cloud's table
        {
            const float *ysrc = a.y + (size_t)cloud * a.n_in * a.ldy + a.yoff;
            for (int i = tid; i < a.n_in * (C1 / 4); i += SA_THREADS) {
                const int r = i / (C1 / 4), c4 = i % (C1 / 4);
                *reinterpret_cast<float4 *>(s_y + r * YS + 4 * c4) = __ldg(reinterpret_cast<const float4 *>(ysrc + (size_t)r * a.ldy) + c4);
            }
            const float *xs = a.xyz_in + (size_t)cloud * a.n_in * 3;
            for (int i = tid; i < a.n_in; i += SA_THREADS) s_xyz[i] = make_float4(__ldg(xs + 3 * i), __ldg(xs + 3 * i + 1), __ldg(xs + 3 * i + 2), 0.0f);
            const float *cs = a.xyz_c + (size_t)cloud * a.npts * 3;
            for (int i = tid; i < a.npts; i += SA_THREADS) s_ctr[i] = make_float4(__ldg(cs + 3 * i), __ldg(cs + 3 * i + 1), __ldg(cs + 3 * i + 2), 0.0f);
        }
        __syncthreads();
        for (long long t = t0 + slot; t < seg_end; t += SA_SLOTS) {
            const int j = t_next == t ? j_next : __ldg(a.idx + t * 128 + rit);   // (a short first segment can skip this slot)
            // next tile of this slot (possibly in the next cloud segment): its index load flies under this tile
            t_next = t + SA_SLOTS < seg_end ? t + SA_SLOTS : seg_end + slot;
            if (t_next < hi) j_next = __ldg(a.idx + t_next * 128 + rit);
            const int row_in_cloud = (int)(t - (long long)cloud * T) * 128 + rit;
            const int centre = row_in_cloud >> NS_SHIFT;
            const float4 pj = s_xyz[j], pc = s_ctr[centre];
            const float dx = pj.x - pc.x, dy = pj.y - pc.y, dz = pj.z - pc.z;
            const float2 dx2 = make_float2(dx, dx), dy2 = make_float2(dy, dy), dz2 = make_float2(dz, dz);
            const float *yrow = s_y + j * YS;
            // ---------- first conv: gather + combine -> A operand ----------
#pragma unroll
            for (int ch = 0; ch < C1 / 16; ++ch) {
                float4 yv[4];
#pragma unroll
                for (int g = 0; g < 4; ++g) yv[g] = *reinterpret_cast<const float4 *>(yrow + 16 * ch + 4 * g);
                uint32_t hi_[8], lo_[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 ga = s_gc[2 * (8 * ch + i)], gb = s_gc[2 * (8 * ch + i) + 1];
                    const float4 yq = yv[i >> 1];
                    const float2 y2 = (i & 1) ? make_float2(yq.z, yq.w) : make_float2(yq.x, yq.y);
                    // fma(wz, dz, fma(wy, dy, fma(wx, dx, y + b1))), the operation order of mlp_tc.cu's gather mode
                    float2 v = rt_fadd2(y2, make_float2(gb.z, gb.w));
                    v = rt_ffma2(make_float2(ga.x, ga.y), dx2, v);
                    v = rt_ffma2(make_float2(ga.z, ga.w), dy2, v);
                    v = rt_ffma2(make_float2(gb.x, gb.y), dz2, v);
                    v = make_float2(fmaxf(v.x, 0.0f), fmaxf(v.y, 0.0f));
                    sa_split2(v, hi_[i], lo_[i], amax);
                }
                sa_st8(tAhi + lane_base + 8 * ch, hi_);
                sa_st8(tAlo + lane_base + 8 * ch, lo_);
            }
            sa_st_wait();
            sa_fence_before();
            // the 4 warps of the slot: A is complete and everybody has finished reading the previous accumulator
            asm volatile("bar.sync %0, 128;" ::"r"(1 + slot) : "memory");
            if (rit == 0) {
                sa_fence_after();
                constexpr uint32_t LBO = (C2 / 8) * 128;
                const uint32_t hi_base = rt_smem_u32(s_w), lo_base = hi_base + 2 * C1 * C2;
                constexpr uint32_t IDESC = sa_idesc(C2);
                // corrections (lo*hi + hi*lo, carrying 2^11) first, then the main products, the first of them rescaling
#pragma unroll
                for (int kk = 0; kk < C1 / 16; ++kk) {
                    sa_mma_ts(tD, tAlo + 8 * kk, sa_desc(hi_base + kk * 2 * LBO, LBO, 128), IDESC, kk > 0);
                    sa_mma_ts(tD, tAhi + 8 * kk, sa_desc(lo_base + kk * 2 * LBO, LBO, 128), IDESC, 1);
                }
                sa_mma_ts_rescale(tD, tAhi, sa_desc(hi_base, LBO, 128), IDESC);
#pragma unroll
                for (int kk = 1; kk < C1 / 16; ++kk) sa_mma_ts(tD, tAhi + 8 * kk, sa_desc(hi_base + kk * 2 * LBO, LBO, 128), IDESC, 1);
                sa_commit(&bar_d[slot]);
            }
            __syncwarp();
            sa_mbar_wait(&bar_d[slot], d_phase);
            d_phase ^= 1;
            sa_fence_after();
            if (C3 > 0) {
                // ---------- mid epilogue: bias, ReLU -> A operand of the third conv; its MMAs; wait ----------
#pragma unroll
                for (int c0 = 0; c0 < C2; c0 += 16) {
                    uint32_t r[16], hi_[8], lo_[8];
                    sa_ld16(tD + lane_base + c0, r);
                    sa_ld_wait();
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float2 x = rt_ffma2(make_float2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])),
                                                  make_float2(SA_WINV, SA_WINV), *reinterpret_cast<const float2 *>(s_b2 + c0 + 2 * i));
                        sa_split2(make_float2(fmaxf(x.x, 0.0f), fmaxf(x.y, 0.0f)), hi_[i], lo_[i], amax);
                    }
                    sa_st8(tAhi + lane_base + c0 / 2, hi_);
                    sa_st8(tAlo + lane_base + c0 / 2, lo_);
                }
                sa_st_wait();
                sa_fence_before();
                asm volatile("bar.sync %0, 128;" ::"r"(1 + slot) : "memory");   // A complete, accumulator read by all four warps
                if (rit == 0) {
                    sa_fence_after();
                    constexpr uint32_t LBO3 = ((C3 > 0 ? C3 : 8) / 8) * 128;
                    const uint32_t hi_base = rt_smem_u32(s_w3), lo_base = hi_base + 2 * C2 * C3;
                    constexpr uint32_t IDESC3 = sa_idesc(C3 > 0 ? C3 : 8);
#pragma unroll
                    for (int kk = 0; kk < C2 / 16; ++kk) {
                        sa_mma_ts(tD, tAlo + 8 * kk, sa_desc(hi_base + kk * 2 * LBO3, LBO3, 128), IDESC3, kk > 0);
                        sa_mma_ts(tD, tAhi + 8 * kk, sa_desc(lo_base + kk * 2 * LBO3, LBO3, 128), IDESC3, 1);
                    }
                    sa_mma_ts_rescale(tD, tAhi, sa_desc(hi_base, LBO3, 128), IDESC3);
#pragma unroll
                    for (int kk = 1; kk < C2 / 16; ++kk) sa_mma_ts(tD, tAhi + 8 * kk, sa_desc(hi_base + kk * 2 * LBO3, LBO3, 128), IDESC3, 1);
                    sa_commit(&bar_d[slot]);
                }
                __syncwarp();
                sa_mbar_wait(&bar_d[slot], d_phase);
                d_phase ^= 1;
                sa_fence_after();
            }
            // ---------- epilogue: bias, ReLU, max over the NS rows of a centre ----------
            const float *s_bl = C3 > 0 ? s_b3 : s_b2;
            float *orow = a.out + ((size_t)cloud * a.npts + centre) * a.ldo + a.ooff;
#pragma unroll
            for (int c0 = 0; c0 < CL; c0 += 16) {
                uint32_t r[16];
                sa_ld16(tD + lane_base + c0, r);
                sa_ld_wait();
                float v[16];
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    const float2 x = rt_ffma2(make_float2(__uint_as_float(r[i]), __uint_as_float(r[i + 1])), make_float2(SA_WINV, SA_WINV),
                                              *reinterpret_cast<const float2 *>(s_bl + c0 + i));
                    v[i] = fmaxf(x.x, 0.0f);
                    v[i + 1] = fmaxf(x.y, 0.0f);
                }
                if (NS == 32) {
                    // every value is >= +0 after ReLU: IEEE order = unsigned order of the bit patterns, one redux per column
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(rt_redux_max_u32(__float_as_uint(v[i]) & 0x7fffffffu));
                    if (lane == 0) {
#pragma unroll
                        for (int g = 0; g < 4; ++g)
                            *reinterpret_cast<float4 *>(orow + c0 + 4 * g) = make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
                    }
                } else {
                    int col0, cnt;
                    bool writer;
                    sa_group_max<NS>(v, lane, col0, cnt, writer);
                    if (writer) {
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (i < cnt) orow[c0 + col0 + i] = v[i];
                    }
                }
            }
            sa_fence_before();
        }
        t0 = seg_end;
    }
    if (!(fmaxf(__low2float(amax), __high2float(amax)) < 65000.0f) && a.status) atomicOr(a.status, 2);
    sa_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
}

template <int C1, int C2, int C3, int NS>
int sa_launch(const SaTcArgs &a, size_t smem_bytes, int grid, cudaStream_t st) {
    static RtPerDevice attr_set;
    if (!attr_set.done(rt_current_device())) {
        const cudaError_t e = cudaFuncSetAttribute(sa_tc_kernel<C1, C2, C3, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 256);
        if (e != cudaSuccess) {
            rt_set_error("sa_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return (int)e;
        }
        attr_set.mark(rt_current_device());
    }
    sa_tc_kernel<C1, C2, C3, NS><<<grid, SA_THREADS, smem_bytes, st>>>(a);
    return rt_check_launch("sa_tc_kernel");
}

}  // namespace

// Tries the shared-memory-table kernel for a gather-mode / max-pool RtMlpTc job with one or two tensor-core layers.  Returns
// RT_ERR_UNSUPPORTED (without setting an error text) when the shape is not instantiated or the table does not fit: the
// caller then launches the generic kernel.
int rt_launch_sa_tc(const RtMlpTc &m, int clouds, cudaStream_t st) {
    if (m.load_mode != RT_MLP_LOAD_GATHER || m.out_mode != RT_MLP_OUT_MAXPOOL || m.nlayers < 1 || m.nlayers > 2 || m.cloud_bias || m.mid_out)
        return RT_ERR_UNSUPPORTED;
    const int c1 = m.c1, c2 = m.layer[0].n, c3 = m.nlayers == 2 ? m.layer[1].n : 0, ns = m.ns;
    if (m.layer[0].k != c1 || m.layer[0].act != RT_ACT_RELU || m.n_out != (c3 ? c3 : c2)) return RT_ERR_UNSUPPORTED;
    if (c3 && (m.layer[1].k != c2 || m.layer[1].act != RT_ACT_RELU)) return RT_ERR_UNSUPPORTED;
    if (clouds <= 0 || m.rows != (long long)clouds * m.npts * ns || (m.npts * ns) % 128 != 0) return RT_ERR_UNSUPPORTED;
    if ((m.ldy & 3) || (m.yoff & 3) || (reinterpret_cast<uintptr_t>(m.y) & 15) || (m.ldo & 3) || (m.ooff & 3) ||
        (reinterpret_cast<uintptr_t>(m.out) & 15))
        return RT_ERR_UNSUPPORTED;
    const size_t smem_bytes = (size_t)4 * c1 * c2 + (size_t)4 * c2 * c3 + (size_t)c1 * 16 + (size_t)c2 * 4 + (size_t)(c3 ? c3 : 4) * 4 +
                              (size_t)(m.n_in + m.npts) * 16 + (size_t)m.n_in * (c1 + 4) * 4;
    if (smem_bytes > 227 * 1024 - 256) return RT_ERR_UNSUPPORTED;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long tiles = m.rows / 128;
    // a CTA pays one table load per cloud segment: give it at least ~8 tiles
    long long grid = tiles / 8;
    grid = grid < 1 ? 1 : (grid > sms ? sms : grid);
    SaTcArgs a{clouds, m.npts, m.n_in, m.y, m.ldy, m.yoff, m.idx, m.xyz_in, m.xyz_c, m.wx, m.b1, m.layer[0].wpack, m.layer[0].bias,
               c3 ? m.layer[1].wpack : nullptr, c3 ? m.layer[1].bias : nullptr, m.out, m.ldo, m.ooff, m.status};
#define SA_CASE(C1, C2, C3, NS) \
    if (c1 == C1 && c2 == C2 && c3 == C3 && ns == NS) return sa_launch<C1, C2, C3, NS>(a, smem_bytes, (int)grid, st)
    SA_CASE(16, 16, 32, 4);    // level 1 (three convolutions)
    SA_CASE(16, 16, 32, 8);
    SA_CASE(32, 32, 0, 8);     // level 2
    SA_CASE(32, 64, 0, 16);
    SA_CASE(64, 64, 0, 16);    // level 3
    SA_CASE(64, 64, 0, 32);
#undef SA_CASE
    return RT_ERR_UNSUPPORTED;
}
