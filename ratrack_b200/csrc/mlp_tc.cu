// Row-tile MLP chains on the tcgen05 tensor cores (sm_100a): the workhorse behind every dense layer of the
// fused engine except the streamed 256x256 cost-volume layers (costvol_tc.cu).
//
//   tile      = 128 rows (row = TMEM lane);  persistent CTAs loop over tiles;
//   input     = MT_LOAD_ROWS:   concatenation of up to 4 row-major fp32 segments (each padded to 16 columns), or
//               MT_LOAD_GATHER: the first set-abstraction layer evaluated on the fly from per-point projections,
//                               relu(Y[idx] + Wx.(xyz[idx] - centre) + b1)   (reference: QueryAndGroup + first conv,
//                               src/lib/pointnet2_utils.py:269-292, src/lib/pytorch_utils.py:5-32) -- the grouped
//                               (B,C,S,ns) tensor is never built;
//   layers    = 1..4 x [tcgen05.mma kind::f16 M128 N<=256, A from TMEM, weights resident in shared memory in the
//               K-major core-matrix layout, fp16 hi/lo split with 3 MMAs per product (fp32-class accuracy)]
//               with bias / ReLU / LeakyReLU epilogues that write the next layer's A operand straight back to TMEM;
//   output    = MT_OUT_ROWS: fp32 rows, or MT_OUT_MAXPOOL: max over the ns consecutive rows of a centre
//               (reference: F.max_pool2d over nsample, src/lib/pointnet2_modules.py:42-44) by warp shuffles.
#include <cuda_fp16.h>
#include <stdlib.h>

#include "engine_kernels.cuh"
#include "mlp_tc.cuh"

namespace {

// worker warps: NW / 4 per TMEM lane quarter, they split the 16-column chunks of a row between them.  8 for the narrow shapes
// that keep 2-3 CTAs resident per SM; 16 for the shapes that fill an SM's TMEM / shared memory with ONE CTA (N = 256 layers),
// where 9 warps leave the SM's schedulers idle between the TMEM round trips of the loader and the epilogues
constexpr int mt_threads(int nw) { return 32 * (nw + 1); }
constexpr float MT_WINV = 1.0f / 1024.0f;
// x = hi + lo with two fp16 planes; the lo plane is stored as 2^11 * lo so that it is a NORMAL fp16 number whenever hi is
// (an unscaled lo of an activation below 0.25 is subnormal and loses up to 10 bits: tools/tc_precision.cu, data set 2)
constexpr float MT_LO_SCALE = 2048.0f;

// A failed probe backs off with nanosleep: a spinning warp otherwise burns issue slots of the SM sub-partition it shares
// with the warps (of this and the co-resident CTAs) that have work -- ncu counted a third of this kernel's issued
// instructions in these loops (profiles/r1_ncu_mlp_tc_sa3.txt: ISETP + BRA + SYNCS + YIELD).
__device__ __forceinline__ void mt_mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = rt_smem_u32(bar);
    for (uint32_t it = 0;; ++it) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
        if (ok) return;
        __nanosleep(it < 4 ? 32 : 96);
        if (it > (1u << 22)) __trap();
    }
}
__device__ __forceinline__ void mt_tmem_alloc(uint32_t *slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(rt_smem_u32(slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void mt_tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void mt_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void mt_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void mt_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(rt_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mt_mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
// D = A*B + D * 2^-11 (the instruction's scale-input-d operand): folds the 2^11 scaling of the correction products back
__device__ __forceinline__ void mt_mma_ts_rescale(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p, 11;\n\t}"
                 ::"r"(d), "r"(a), "l"(b), "r"(idesc) : "memory");
}
__device__ __forceinline__ void mt_ld16(uint32_t taddr, uint32_t *r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void mt_st8(uint32_t taddr, const uint32_t *r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void mt_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint64_t mt_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3fff) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ uint32_t mt_idesc(int n) {  // F16 x F16 -> F32, K-major, M = 128
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
// branch-free activation: max(v, slope * v) with slope 0 (ReLU), 0.1 (LeakyReLU) or 1 (none)
__device__ __forceinline__ float mt_slope(int act) { return act == RT_ACT_RELU ? 0.0f : (act == RT_ACT_LEAKY01 ? 0.1f : 1.0f); }
__device__ __forceinline__ float mt_act(float v, float slope) { return fmaxf(v, v * slope); }
// range guard: the running max of |hi| is kept as a half2 (one HMNMX2 per pair); an fp32 input beyond the fp16 range
// rounds to inf and trips it
__device__ __forceinline__ void mt_split2(float x0, float x1, uint32_t &hi, uint32_t &lo, __half2 &amax) {
    const __half2 h = __floats2half2_rn(x0, x1);
    amax = __hmax2(amax, __habs2(h));
    const float2 hf = __half22float2(h);
    // packed fp32 pairs (FADD2 / FMUL2): both halves are ordinary round-to-nearest operations
    const float2 r = rt_fmul2(rt_fadd2(make_float2(x0, x1), make_float2(-hf.x, -hf.y)), make_float2(MT_LO_SCALE, MT_LO_SCALE));
    const __half2 l = __floats2half2_rn(r.x, r.y);
    hi = *reinterpret_cast<const uint32_t *>(&h);
    lo = *reinterpret_cast<const uint32_t *>(&l);
}
// 16 fp32 values -> hi/lo fp16 planes, 8 TMEM columns each
__device__ __forceinline__ void mt_store16(const float *v, uint32_t t_hi, uint32_t t_lo, __half2 &amax) {
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) mt_split2(v[2 * i], v[2 * i + 1], hi[i], lo[i], amax);
    mt_st8(t_hi, hi);
    mt_st8(t_lo, lo);
}

// max over the NS consecutive lanes of a centre for 16 values per lane.  Halving butterfly: at every step a lane
// keeps half of its values and hands the other half to its partner, so NS-lane groups need ~16 shuffles, not 16*log2(NS).
// On return v[0..cnt) are the maxima of columns col0 .. col0+cnt of the chunk; lanes with `writer` store them.
template <int NS>
__device__ __forceinline__ void mt_group_max(float *v, int lane, int &col0, int &cnt, bool &writer) {
    col0 = 0;
    writer = true;
    constexpr int STEPS = NS == 32 ? 5 : NS == 16 ? 4 : NS == 8 ? 3 : NS == 4 ? 2 : NS == 2 ? 1 : 0;
    int c = 16;
#pragma unroll
    for (int st = 0; st < STEPS; ++st) {
        const int off = NS >> (st + 1);
        const int half = 16 >> (st + 1);   // compile-time after unrolling
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

// Gather-mode operands of one row, fetched one tile ahead so that the index -> xyz -> projection-row load chain of
// tile t+1 (three dependent L2 round trips) runs underneath the MMA and epilogue of tile t.
struct MtGatherRow {
    const float *yrow;   // projected row of the neighbour (this scale's columns)
    float dx, dy, dz;
};
__device__ __forceinline__ void mt_gather_issue(const RtMlpTc &a, long long tile, int row_in_tile, MtGatherRow &r) {
    uint32_t row = (uint32_t)tile * 128u + (uint32_t)row_in_tile;
    if (row >= (uint32_t)a.rows) row = (uint32_t)a.rows - 1u;
    const uint32_t cp = row >> a.ns_shift;            // (cloud, centre)
    const uint32_t cloud = cp / (uint32_t)a.npts;
    const int j = __ldg(a.idx + row);
    const long long g = (long long)cloud * a.n_in + j;
    const float *pc = a.xyz_c + (size_t)cp * 3;
    r.dx = __ldg(a.xyz_in + g * 3 + 0) - __ldg(pc + 0);
    r.dy = __ldg(a.xyz_in + g * 3 + 1) - __ldg(pc + 1);
    r.dz = __ldg(a.xyz_in + g * 3 + 2) - __ldg(pc + 2);
    r.yrow = a.y + g * a.ldy + a.yoff;
}

template <int LOAD_MODE, int MT_WORKER_WARPS>
__global__ void __launch_bounds__(mt_threads(MT_WORKER_WARPS), MT_WORKER_WARPS == 8 ? 3 : 1) mlp_tc_kernel(RtMlpTc a) {
    constexpr int MT_THREADS = mt_threads(MT_WORKER_WARPS), NPART = MT_WORKER_WARPS / 4;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar_a, bar_d, bar_w;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long ntiles = (a.rows + 127) / 128;
    // Programmatic dependent launch: the next kernel of the stream may be scheduled as soon as this grid's CTAs are all
    // resident, and this kernel's own prologue below (weights -> shared memory, biases, TMEM allocation: none of it produced
    // by the previous kernel) runs while the previous grid drains.  griddepcontrol.wait, further down, is the point after
    // which data written by the previous kernel may be read and anything may be written.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    // weights of all layers -> shared memory (already in core-matrix layout); layer l starts at w_off[l]
    int w_off[RT_MLP_MAX_LAYERS + 1];
    w_off[0] = 0;
#pragma unroll
    for (int l = 0; l < RT_MLP_MAX_LAYERS; ++l) w_off[l + 1] = w_off[l] + (l < a.nlayers ? 4 * a.layer[l].k * a.layer[l].n : 0);
    // the bulk-copy engine (UBLKCP) brings every layer's weights in while the workers already fetch their first rows
    if (threadIdx.x == 0) {
        rt_mbar_init(&bar_w, 1);
        rt_fence_mbar_init();
        rt_mbar_expect_tx(&bar_w, (uint32_t)w_off[RT_MLP_MAX_LAYERS]);
        for (int l = 0; l < a.nlayers; ++l) {
            const uint8_t *src = reinterpret_cast<const uint8_t *>(a.layer[l].wpack);
            const int bytes = w_off[l + 1] - w_off[l];
            for (int o = 0; o < bytes; o += 32768)
                rt_bulk_g2s(smem + w_off[l] + o, src + o, (uint32_t)min(32768, bytes - o), &bar_w);
        }
    }
    // per-layer biases -> shared memory (zeros when a layer has none), 256 floats per layer
    float *s_bias = reinterpret_cast<float *>(smem + w_off[RT_MLP_MAX_LAYERS]);
    for (int l = 0; l < a.nlayers; ++l)
        for (int i = threadIdx.x; i < a.layer[l].n; i += MT_THREADS)
            s_bias[l * 256 + i] = a.layer[l].bias ? __ldg(a.layer[l].bias + i) : 0.0f;
    // gather mode: (wx, wy, wz, b1) of every first-conv channel as one float4
    float4 *s_gc = reinterpret_cast<float4 *>(s_bias + RT_MLP_MAX_LAYERS * 256);
    if (LOAD_MODE == RT_MLP_LOAD_GATHER)
        for (int i = threadIdx.x; i < a.c1; i += MT_THREADS)
            s_gc[i] = make_float4(__ldg(a.wx + i * 3 + 0), __ldg(a.wx + i * 3 + 1), __ldg(a.wx + i * 3 + 2), __ldg(a.b1 + i));
    if (threadIdx.x == 0) {
        rt_mbar_init(&bar_a, MT_WORKER_WARPS);
        rt_mbar_init(&bar_d, 1);
        rt_fence_mbar_init();
    }
    rt_fence_proxy_async();
    if (warp == MT_WORKER_WARPS) mt_tmem_alloc(&tmem_slot, a.tmem_cols);
    mt_fence_before();
    __syncthreads();
    mt_fence_after();
    const uint32_t tm = tmem_slot;
    const uint32_t tD = tm, tAhi = tm + a.nsplit * a.d_cols, tAlo = tAhi + a.a_cols;
    asm volatile("griddepcontrol.wait;" ::: "memory");   // the previous kernel's outputs (our inputs) are complete and visible

    if (warp == MT_WORKER_WARPS) {
        // ===== MMA issuer =====
        if (lane == 0) {
            uint32_t a_phase = 0;
            mt_mbar_wait(&bar_w, 0);   // weights have landed in shared memory
            for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                for (int l = 0; l < a.nlayers; ++l) {
                    mt_mbar_wait(&bar_a, a_phase);
                    a_phase ^= 1;
                    mt_fence_after();
                    const int k = a.layer[l].k, n = a.layer[l].n;
                    const uint32_t lbo = (uint32_t)(n / 8) * 128;
                    const uint32_t hi_base = rt_smem_u32(smem + w_off[l]), lo_base = hi_base + 2 * k * n;
                    const uint32_t idesc = mt_idesc(n);
                    // Issue order matters for accuracy: the tensor core TRUNCATES the fp32 accumulator after every
                    // k-step (measured, tools/tc_precision.cu: the error grows with the number of accumulation steps taken
                    // at full magnitude and is biased toward zero).  So: (1) the correction products lo*hi + hi*lo first,
                    // both carrying the 2^11 of the scaled lo planes; (2) the first main product rescales that partial sum
                    // with scale-input-d (D = A*B + D * 2^-11); (3) the remaining k/16 - 1 main products.  With a second
                    // accumulator (a.nsplit == 2) the second half of K accumulates separately and the epilogue adds the two
                    // in fp32 round-to-nearest: half as many truncations at full magnitude.
                    const int ksteps = k / 16;
                    const int khalf = (a.nsplit == 2 && ksteps >= 4) ? ksteps / 2 : ksteps;
                    for (int kk = 0; kk < ksteps; ++kk) {
                        mt_mma_ts(tD, tAlo + 8 * kk, mt_desc(hi_base + kk * 2 * lbo, lbo, 128), idesc, kk > 0);
                        mt_mma_ts(tD, tAhi + 8 * kk, mt_desc(lo_base + kk * 2 * lbo, lbo, 128), idesc, 1);
                    }
                    mt_mma_ts_rescale(tD, tAhi, mt_desc(hi_base, lbo, 128), idesc);
                    for (int kk = 1; kk < khalf; ++kk)
                        mt_mma_ts(tD, tAhi + 8 * kk, mt_desc(hi_base + kk * 2 * lbo, lbo, 128), idesc, 1);
                    for (int kk = khalf; kk < ksteps; ++kk)
                        mt_mma_ts(tD + a.d_cols, tAhi + 8 * kk, mt_desc(hi_base + kk * 2 * lbo, lbo, 128), idesc, kk > khalf);
                    mt_commit(&bar_d);
                }
            }
        }
        __syncwarp();
    } else {
        // ===== worker warps: one thread per row =====
        const int q = warp & 3, hlf = warp >> 2;        // TMEM lane quarter, chunk residue (mod NPART)
        const int row_in_tile = 32 * q + lane;
        const uint32_t lane_base = (uint32_t)(32 * q) << 16;
        uint32_t d_phase = 0;
        const bool out_al32 = (reinterpret_cast<uintptr_t>(a.out) & 31) == 0;
        __half2 amax = __floats2half2_rn(0.0f, 0.0f);
        MtGatherRow pre;
        if (LOAD_MODE == RT_MLP_LOAD_GATHER && (long long)blockIdx.x < ntiles) mt_gather_issue(a, blockIdx.x, row_in_tile, pre);
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const long long row = tile * 128 + row_in_tile;
            const bool valid = row < a.rows;
            const long long rc = valid ? row : a.rows - 1;
            // ---------- layer-0 input ----------
            if (LOAD_MODE == RT_MLP_LOAD_ROWS) {
                int c0 = 0;
                for (int s = 0; s < a.nseg; ++s) {
                    const int ks = a.seg[s].k;
                    if (a.seg[s].nn_idx) {
                        // interpolated segment: three gathered rows, the reference kernel's fma order
                        const int *ix = a.seg[s].nn_idx + rc * 3;
                        const float *ww = a.seg[s].nn_w + rc * 3;
                        const long long cbase = (rc / a.seg[s].nn_n) * a.seg[s].nn_m;
                        const float *r0 = a.seg[s].x + (cbase + __ldg(ix + 0)) * a.seg[s].ldx;
                        const float *r1 = a.seg[s].x + (cbase + __ldg(ix + 1)) * a.seg[s].ldx;
                        const float *r2 = a.seg[s].x + (cbase + __ldg(ix + 2)) * a.seg[s].ldx;
                        const float w0 = __ldg(ww + 0), w1 = __ldg(ww + 1), w2 = __ldg(ww + 2);
                        for (int o = 0; o < ks; o += 16, c0 += 16) {
                            if (((c0 >> 4) % NPART) != hlf) continue;
                            float v[16];
                            float4 q0[4], q1[4], q2[4];
                            if ((a.seg[s].ldx & 7) == 0) {   // rows are 32-byte aligned: 256-bit loads
#pragma unroll
                                for (int g = 0; g < 4; g += 2) {
                                    rt_ldg256(r0 + o + 4 * g, q0[g], q0[g + 1]);
                                    rt_ldg256(r1 + o + 4 * g, q1[g], q1[g + 1]);
                                    rt_ldg256(r2 + o + 4 * g, q2[g], q2[g + 1]);
                                }
                            } else {
#pragma unroll
                                for (int g = 0; g < 4; ++g) {
                                    q0[g] = __ldg(reinterpret_cast<const float4 *>(r0 + o) + g);
                                    q1[g] = __ldg(reinterpret_cast<const float4 *>(r1 + o) + g);
                                    q2[g] = __ldg(reinterpret_cast<const float4 *>(r2 + o) + g);
                                }
                            }
#pragma unroll
                            for (int g = 0; g < 4; ++g) {
                                const float4 p0 = q0[g], p1 = q1[g], p2 = q2[g];
                                v[4 * g + 0] = __fmaf_rn(w2, p2.x, __fmaf_rn(w0, p0.x, __fmul_rn(w1, p1.x)));
                                v[4 * g + 1] = __fmaf_rn(w2, p2.y, __fmaf_rn(w0, p0.y, __fmul_rn(w1, p1.y)));
                                v[4 * g + 2] = __fmaf_rn(w2, p2.z, __fmaf_rn(w0, p0.z, __fmul_rn(w1, p1.z)));
                                v[4 * g + 3] = __fmaf_rn(w2, p2.w, __fmaf_rn(w0, p0.w, __fmul_rn(w1, p1.w)));
                            }
                            mt_store16(v, tAhi + lane_base + c0 / 2, tAlo + lane_base + c0 / 2, amax);
                        }
                        continue;
                    }
                    const float *x = a.seg[s].x + rc * a.seg[s].ldx;
                    const bool vec = (a.seg[s].ldx & 3) == 0;
                    for (int o = 0; o < ks; o += 16, c0 += 16) {
                        if (((c0 >> 4) % NPART) != hlf) continue;
                        float v[16];
                        if (vec && o + 16 <= ks) {
                            float4 t[4];
                            if ((a.seg[s].ldx & 7) == 0) {
                                rt_ldg256(x + o, t[0], t[1]);
                                rt_ldg256(x + o + 8, t[2], t[3]);
                            } else {
#pragma unroll
                                for (int g = 0; g < 4; ++g) t[g] = __ldg(reinterpret_cast<const float4 *>(x + o) + g);
                            }
#pragma unroll
                            for (int g = 0; g < 4; ++g) {
                                v[4 * g] = t[g].x; v[4 * g + 1] = t[g].y; v[4 * g + 2] = t[g].z; v[4 * g + 3] = t[g].w;
                            }
                        } else {
#pragma unroll
                            for (int i = 0; i < 16; ++i) v[i] = (o + i < ks) ? __ldg(x + o + i) : 0.0f;
                        }
                        mt_store16(v, tAhi + lane_base + c0 / 2, tAlo + lane_base + c0 / 2, amax);
                    }
                }
            } else {
                // index, xyz difference and row pointer were fetched during the previous tile; the row itself now
                for (int c0 = 16 * hlf; c0 < a.c1; c0 += 16 * NPART) {
                    float v[16];
                    float4 t[4];
                    if ((a.ldy & 7) == 0 && (a.yoff & 7) == 0) {
                        rt_ldg256(pre.yrow + c0, t[0], t[1]);
                        rt_ldg256(pre.yrow + c0 + 8, t[2], t[3]);
                    } else {
#pragma unroll
                        for (int gq = 0; gq < 4; ++gq) t[gq] = __ldg(reinterpret_cast<const float4 *>(pre.yrow + c0) + gq);
                    }
#pragma unroll
                    for (int gq = 0; gq < 4; ++gq) {
                        v[4 * gq] = t[gq].x; v[4 * gq + 1] = t[gq].y; v[4 * gq + 2] = t[gq].z; v[4 * gq + 3] = t[gq].w;
                    }
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float4 w = s_gc[c0 + i];
                        v[i] = fmaxf(fmaf(w.z, pre.dz, fmaf(w.y, pre.dy, fmaf(w.x, pre.dx, v[i] + w.w))), 0.0f);
                    }
                    mt_store16(v, tAhi + lane_base + c0 / 2, tAlo + lane_base + c0 / 2, amax);
                }
            }
            mt_st_wait();
            mt_fence_before();
            __syncwarp();
            if (lane == 0) rt_mbar_arrive(&bar_a);
            if (LOAD_MODE == RT_MLP_LOAD_GATHER && tile + gridDim.x < ntiles)
                mt_gather_issue(a, tile + gridDim.x, row_in_tile, pre);   // next tile's index -> xyz chain flies under this tile's MMA

            for (int l = 0; l < a.nlayers; ++l) {
                mt_mbar_wait(&bar_d, d_phase);
                d_phase ^= 1;
                mt_fence_after();
                const int n = a.layer[l].n;
                const float slope = mt_slope(a.layer[l].act);
                const float *bias = s_bias + l * 256;
                const float *cb = (l == 0 && a.cloud_bias)
                                      ? a.cloud_bias + (size_t)((uint32_t)rc / (uint32_t)a.rows_per_cloud) * a.cloud_bias_ld : nullptr;
                const bool last = l == a.nlayers - 1;
                for (int c0 = 16 * hlf; c0 < n; c0 += 16 * NPART) {
                    uint32_t r[16];
                    mt_ld16(tD + lane_base + c0, r);
                    if (a.nsplit == 2 && a.layer[l].k >= 64) {   // second K half accumulated separately (see the issuer)
                        uint32_t r2[16];
                        mt_ld16(tD + a.d_cols + lane_base + c0, r2);
#pragma unroll
                        for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) + __uint_as_float(r2[i]));
                    }
                    float v[16];
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const float4 bq = *reinterpret_cast<const float4 *>(bias + c0 + 4 * g);
                        v[4 * g + 0] = fmaf(__uint_as_float(r[4 * g + 0]), MT_WINV, bq.x);
                        v[4 * g + 1] = fmaf(__uint_as_float(r[4 * g + 1]), MT_WINV, bq.y);
                        v[4 * g + 2] = fmaf(__uint_as_float(r[4 * g + 2]), MT_WINV, bq.z);
                        v[4 * g + 3] = fmaf(__uint_as_float(r[4 * g + 3]), MT_WINV, bq.w);
                    }
                    if (cb) {
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            const float4 cq = __ldg(reinterpret_cast<const float4 *>(cb + c0) + g);
                            v[4 * g + 0] += cq.x; v[4 * g + 1] += cq.y; v[4 * g + 2] += cq.z; v[4 * g + 3] += cq.w;
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = mt_act(v[i], slope);
                    if (!last) {
                        mt_store16(v, tAhi + lane_base + c0 / 2, tAlo + lane_base + c0 / 2, amax);
                        if (a.mid_out && l == a.mid_layer && valid) {
                            float *o = a.mid_out + row * a.mid_ldo + c0;
#pragma unroll
                            for (int g = 0; g < 4; ++g)
                                reinterpret_cast<float4 *>(o)[g] = make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
                        }
                    } else if (a.out_mode == RT_MLP_OUT_ROWS) {
                        if (valid) {
                            float *o = a.out + row * a.ldo + a.ooff + c0;
                            if (c0 + 16 <= a.n_out && (a.ldo & 7) == 0 && (a.ooff & 7) == 0 && out_al32) {
                                // thread-per-row stores pay one L1 wavefront per lane per instruction: 256-bit stores
                                rt_stg256(o, make_float4(v[0], v[1], v[2], v[3]), make_float4(v[4], v[5], v[6], v[7]));
                                rt_stg256(o + 8, make_float4(v[8], v[9], v[10], v[11]), make_float4(v[12], v[13], v[14], v[15]));
                            } else if (c0 + 16 <= a.n_out && (a.ldo & 3) == 0 && (a.ooff & 3) == 0) {
#pragma unroll
                                for (int g = 0; g < 4; ++g)
                                    reinterpret_cast<float4 *>(o)[g] = make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
                            } else {
#pragma unroll
                                for (int i = 0; i < 16; ++i)
                                    if (c0 + i < a.n_out) o[i] = v[i];
                            }
                        }
                    } else {
                        // max over the ns consecutive rows (= lanes) of a centre; 128 % ns == 0 so groups never straddle tiles
                        int col0 = 0, cnt = 16;
                        bool writer = true;
                        if (a.ns == 32 && a.layer[l].act == RT_ACT_RELU) {
                            // the 32 lanes of the warp are one centre and every value is >= +0 after ReLU, so IEEE order is
                            // unsigned-integer order of the bit patterns: one redux.sync.max per column (16 instructions
                            // instead of the 31 shuffles + selects of the butterfly)
#pragma unroll
                            for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(rt_redux_max_u32(__float_as_uint(v[i]) & 0x7fffffffu));   // -0.0 -> +0.0
                            writer = lane == 0;
                        } else
                        switch (a.ns) {
                            case 32: mt_group_max<32>(v, lane, col0, cnt, writer); break;
                            case 16: mt_group_max<16>(v, lane, col0, cnt, writer); break;
                            case 8: mt_group_max<8>(v, lane, col0, cnt, writer); break;
                            case 4: mt_group_max<4>(v, lane, col0, cnt, writer); break;
                            case 2: mt_group_max<2>(v, lane, col0, cnt, writer); break;
                            default: break;
                        }
                        if (valid && writer) {
                            float *o = a.out + (size_t)((uint32_t)row >> a.ns_shift) * a.ldo + a.ooff + c0 + col0;
#pragma unroll
                            for (int i = 0; i < 16; ++i)
                                if (i < cnt && c0 + col0 + i < a.n_out) o[i] = v[i];
                        }
                    }
                }
                if (!last) {
                    mt_st_wait();
                    mt_fence_before();
                    __syncwarp();
                    if (lane == 0) rt_mbar_arrive(&bar_a);
                }
            }
            mt_fence_before();
        }
        if (!(fmaxf(__low2float(amax), __high2float(amax)) < 65000.0f) && a.status) atomicOr(a.status, 2);
    }
    mt_fence_before();
    __syncthreads();
    if (warp == MT_WORKER_WARPS) mt_tmem_dealloc(tm, a.tmem_cols);
}

}  // namespace

int rt_launch_mlp_tc(RtMlpTc a, cudaStream_t st) {
    if (a.rows <= 0) return RT_OK;
    RT_REQUIRE(a.nlayers >= 1 && a.nlayers <= RT_MLP_MAX_LAYERS, "mlp_tc: nlayers=%d", a.nlayers);
    int dmax = 0, kmax = 0, wbytes = 0;
    for (int l = 0; l < a.nlayers; ++l) {
        const int k = a.layer[l].k, n = a.layer[l].n;
        RT_REQUIRE(k >= 16 && k % 16 == 0 && n >= 16 && n % 16 == 0 && n <= 256, "mlp_tc: layer %d has k=%d n=%d", l, k, n);
        RT_REQUIRE(l == 0 || k == a.layer[l - 1].n, "mlp_tc: layer %d k=%d does not chain", l, k);
        dmax = n > dmax ? n : dmax;
        kmax = k > kmax ? k : kmax;
        wbytes += 4 * k * n;
    }
    if (a.load_mode == RT_MLP_LOAD_ROWS) {
        int ktot = 0;
        for (int s = 0; s < a.nseg; ++s) ktot += (a.seg[s].k + 15) / 16 * 16;
        RT_REQUIRE(ktot == a.layer[0].k, "mlp_tc: segments cover %d columns, layer 0 expects %d", ktot, a.layer[0].k);
        for (int s = 0; s < a.nseg; ++s)
            RT_REQUIRE(!a.seg[s].nn_idx || ((a.seg[s].k & 15) == 0 && (a.seg[s].ldx & 3) == 0 && a.seg[s].nn_w),
                       "mlp_tc: interpolated segment %d needs k %% 16 == 0 and ldx %% 4 == 0", s);
    } else {
        RT_REQUIRE(a.c1 == a.layer[0].k && a.c1 <= 64 && (a.ldy & 3) == 0 && (a.yoff & 3) == 0, "mlp_tc: gather layout");
    }
    RT_REQUIRE(a.rows < (1ll << 31) - 256, "mlp_tc: %lld rows (32-bit row arithmetic)", a.rows);
    RT_REQUIRE(!a.mid_out || (a.mid_layer >= 0 && a.mid_layer < a.nlayers - 1 && (a.mid_ldo & 3) == 0 &&
                              (reinterpret_cast<uintptr_t>(a.mid_out) & 15) == 0),
               "mlp_tc: mid output needs a non-final layer, ld %% 4 == 0 and 16-byte alignment");
    if (a.load_mode == RT_MLP_LOAD_ROWS) {
        for (int s = 0; s < a.nseg; ++s)
            RT_REQUIRE((a.seg[s].ldx & 3) != 0 || (reinterpret_cast<uintptr_t>(a.seg[s].x) & 31) == 0,
                       "mlp_tc: segment %d must be 32-byte aligned", s);
    } else {
        RT_REQUIRE((reinterpret_cast<uintptr_t>(a.y) & 31) == 0, "mlp_tc: y must be 32-byte aligned");
    }
    RT_REQUIRE(!a.cloud_bias || ((a.cloud_bias_ld & 3) == 0 && (reinterpret_cast<uintptr_t>(a.cloud_bias) & 15) == 0),
               "mlp_tc: cloud_bias must be 16-byte aligned with a leading dimension that is a multiple of 4");
    a.ns_shift = 0;
    if (a.out_mode == RT_MLP_OUT_MAXPOOL || a.load_mode == RT_MLP_LOAD_GATHER) {
        RT_REQUIRE(a.ns >= 1 && a.ns <= 32 && (a.ns & (a.ns - 1)) == 0, "mlp_tc: ns=%d must be a power of two <= 32", a.ns);
        while ((1 << a.ns_shift) < a.ns) ++a.ns_shift;
    }
    const int need = dmax + kmax;  // accumulator columns + two A planes of kmax/2 columns
    RT_REQUIRE(need <= 512, "mlp_tc: %d TMEM columns needed", need);
    int cols = 32;
    while (cols < need) cols *= 2;
    // a second accumulator for the second half of K (accuracy, see the issuer) where it is free: the power-of-two TMEM
    // allocation already has the room.  The choice depends on the layer shapes only, never on the row count: a pair's
    // results must not depend on the batch it travels in.
    a.nsplit = (kmax >= 64 && 2 * dmax + kmax <= cols) ? 2 : 1;
    a.tmem_cols = cols;
    a.d_cols = dmax;
    a.a_cols = kmax / 2;
    RT_REQUIRE(wbytes <= 200 * 1024, "mlp_tc: %d bytes of weights do not fit in shared memory", wbytes);
    static RtPerDevice smem_set;
    if (!smem_set.done(rt_current_device())) {
        cudaError_t e = cudaFuncSetAttribute(mlp_tc_kernel<RT_MLP_LOAD_ROWS, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 204 * 1024);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(mlp_tc_kernel<RT_MLP_LOAD_GATHER, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 204 * 1024);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(mlp_tc_kernel<RT_MLP_LOAD_ROWS, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 204 * 1024);
        if (e != cudaSuccess) {
            rt_set_error("mlp_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return (int)e;
        }
        smem_set.mark(rt_current_device());
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // resident CTAs per SM: bounded by TMEM columns, shared memory and (to keep tail effects small) 4
    int per_sm = 512 / cols;
    const bool gather = a.load_mode == RT_MLP_LOAD_GATHER;
    const int gc_bytes = RT_MLP_MAX_LAYERS * 256 * 4 + (gather ? 16 * a.c1 : 0);   // biases + gather constants
    const int by_smem = (220 * 1024) / (wbytes + gc_bytes + 2048);
    per_sm = per_sm < by_smem ? per_sm : by_smem;
    // register file: the row-loading variant (72 regs x 288 threads) fits 3 CTAs per SM, the gather variant, which keeps
    // the next tile's operands in flight in registers (96 regs), fits 2
    const int by_regs = 3;
    per_sm = per_sm < 1 ? 1 : (per_sm > by_regs ? by_regs : per_sm);
    const long long ntiles = (a.rows + 127) / 128;
    long long grid = (long long)sms * per_sm;
    if (grid > ntiles) grid = ntiles;
    cudaLaunchConfig_t cfg = {};
    // one CTA per SM anyway (TMEM or shared memory): give it 16 worker warps
    static int wide_env = -1;
    if (wide_env < 0) {
        const char *env = getenv("RT_MLP_WIDE");   // RT_MLP_WIDE=0: always 8 worker warps (A/B timing)
        wide_env = (env && atoi(env) == 0) ? 0 : 1;
    }
    const bool wide = wide_env && per_sm == 1 && !gather;
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(mt_threads(wide ? 16 : 8));
    cfg.dynamicSmemBytes = (size_t)(wbytes + gc_bytes);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // see the kernel prologue
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    // Measured on B200 (tools/ab.sh, profiles/r2_ab_pdl_ring_fps.txt): with the attribute the 32-pair step is 1 % SLOWER
    // (2063 vs 2042 us): the early-resident CTAs of the next launch hold registers / shared memory while they wait, which
    // costs the running grid more than the overlapped prologue saves.  Off by default; RT_MLP_PDL=1 enables it for A/B.
    static int use_pdl = -1;
    if (use_pdl < 0) {
        const char *env = getenv("RT_MLP_PDL");
        use_pdl = (env && atoi(env) == 1) ? 1 : 0;
    }
    cfg.numAttrs = use_pdl ? 1 : 0;
    const cudaError_t le = gather ? cudaLaunchKernelEx(&cfg, mlp_tc_kernel<RT_MLP_LOAD_GATHER, 8>, a)
                           : wide ? cudaLaunchKernelEx(&cfg, mlp_tc_kernel<RT_MLP_LOAD_ROWS, 16>, a)
                                  : cudaLaunchKernelEx(&cfg, mlp_tc_kernel<RT_MLP_LOAD_ROWS, 8>, a);
    if (le != cudaSuccess) {
        rt_set_error("mlp_tc_kernel: launch failed: %s", cudaGetErrorString(le));
        return (int)le;
    }
    return rt_check_launch("mlp_tc_kernel");
}

namespace {
struct PackArgs { __half *dst; int n_real, n_pad, k_pad, nseg; RtPackSeg seg[4]; };

__global__ void __launch_bounds__(256) pack_umma_kernel(PackArgs a) {
    const int total = a.n_pad * a.k_pad;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int n = e / a.k_pad, k = e % a.k_pad;
        float w = 0.0f;
        if (n < a.n_real) {
            int c0 = 0;
            for (int s = 0; s < a.nseg; ++s) {
                const int kp = (a.seg[s].k + 15) / 16 * 16;
                if (k >= c0 && k < c0 + kp) {
                    if (k - c0 < a.seg[s].k) w = a.seg[s].w[(long long)n * a.seg[s].ldw + (k - c0)];
                    break;
                }
                c0 += kp;
            }
        }
        w *= 1024.0f;
        const __half hi = __float2half_rn(w);
        const __half lo = __float2half_rn((w - __half2float(hi)) * MT_LO_SCALE);   // scaled like the activations' lo plane
        // [plane][kc = k/8][rg = n/8][n%8][k%8]
        const size_t off = ((size_t)(k / 8) * (a.n_pad / 8) + n / 8) * 64 + (n % 8) * 8 + (k % 8);
        a.dst[off] = hi;
        a.dst[(size_t)a.n_pad * a.k_pad + off] = lo;
    }
}
}  // namespace

int rt_launch_pack_umma(void *dst, int n_real, int n_pad, const RtPackSeg *segs, int nseg, cudaStream_t st) {
    RT_REQUIRE(dst && nseg >= 1 && nseg <= 4 && n_pad % 16 == 0 && n_real <= n_pad, "pack_umma: bad arguments");
    PackArgs a{};
    a.dst = (__half *)dst;
    a.n_real = n_real;
    a.n_pad = n_pad;
    a.nseg = nseg;
    a.k_pad = 0;
    for (int s = 0; s < nseg; ++s) {
        a.seg[s] = segs[s];
        a.k_pad += (segs[s].k + 15) / 16 * 16;
    }
    pack_umma_kernel<<<rt_divup((long long)a.n_pad * a.k_pad, 256), 256, 0, st>>>(a);
    return rt_check_launch("pack_umma_kernel");
}
