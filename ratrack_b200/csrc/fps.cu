// Furthest point sampling for sm_100a.
//
// Replaces furthest_point_sampling_kernel / furthest_point_sampling_kernel_launcher
// (reference: src/lib/src/sampling_gpu.cu:94-253) behind the same contract: idx[0] = 0, `temp`
// holds the running min-distance (caller-initialised, 1e10 in src/lib/pointnet2_utils.py:26)
// and is left holding the final values, results bit-identical to the reference including ties.
//
// Design (B200-first, not a translation):
//   * one CTA per cloud, the whole cloud lives in REGISTERS (coordinates + running min
//     distance), so a round touches no memory except one shared-memory exchange;
//   * the reference's block-wide arg-max is a shared-memory tree with log2(bs)+1 barriers.
//     Its tie-break is equivalent to "among maxima, minimise (bitrev(k mod bs), k div bs)"
//     (bs = the reference's CTA size for this n).  Thread t here owns the residue class
//     bitrev(t), and walks its points in that priority order, so plain "first strict maximum"
//     scans plus "lowest lane / lowest warp wins" reproduce the reference bit for bit with
//     ONE barrier per round: redux.sync.max + ballot inside a warp, one double-buffered
//     shared-memory exchange across warps;
//   * the cloud's xyz is staged once into shared memory by the bulk-copy engine (TMA, UBLKCP)
//     for the broadcast read of the newly selected point each round.
//   A generic kernel (any n) keeps the same ordering rule with 64-bit packed keys.
#include <stdlib.h>

#include "common.cuh"

namespace {

int g_fps_exclusive = 0;
int g_fps_pair = 0;
const int *g_fps_skip = nullptr;   // per-cloud flags: clouds already served by fps_identity_kernel (their CTAs exit at once)
const int *g_fps_counts = nullptr, *g_fps_list = nullptr;   // variable-size batch: per-cloud point counts, clouds of the launch
int g_fps_stride = 0;                                        // ... and the cloud stride (0: the launch's own n)
constexpr size_t kFpsHogBytes = 226 * 1024;   // + static + the 1 KB per-CTA reserve = the SM's 228 KB: no other CTA fits beside it

__host__ __device__ __forceinline__ unsigned bitrev_bits(unsigned v, int bits) {
    unsigned r = 0;
    for (int i = 0; i < bits; ++i) r |= ((v >> i) & 1u) << (bits - 1 - i);
    return r;
}

// T threads; each thread owns PPT = Q*PH points.  bs = T*Q is the reference CTA size.
// Slot s (priority order) = a*PH + ph  ->  point k = r + (bitrev_q(a) + Q*ph)*T,  r = bitrev_logT(t).
// FUSED (engine-internal): the running min-distance starts at the reference's 1e10 in registers and is never stored
// (no temp buffer, no fill kernel), and the coordinates of every selected point are written to new_xyz as it is picked
// (no gather kernel) -- two dependent launches fewer per level on the latency-critical FPS chain.
// CPB clouds per CTA (engine-internal, with the SM-exclusive mode): CPB independent groups of T threads, each with its own
// named barrier, share one SM -- two latency-bound chains interleave almost for free, and the sampling of 2b clouds blocks
// b SMs instead of 2b.
template <int T, int Q, int PH, bool FUSED, int CPB>
__global__ void __launch_bounds__(T *CPB) fps_reg_kernel(int nclouds, int n, int m, const float *__restrict__ xyz_all,
                                                         float *__restrict__ temp_all, int *__restrict__ idx_all,
                                                         float *__restrict__ new_xyz_all, const int *__restrict__ skip,
                                                         const int *__restrict__ counts, const int *__restrict__ list) {
    constexpr int PPT = Q * PH;
    constexpr int NW = (T + 31) / 32;
    constexpr int LOGT = (T == 32) ? 5 : (T == 64) ? 6 : (T == 128) ? 7 : (T == 256) ? 8 : (T == 512) ? 9 : 10;
    constexpr int LOGQ = (Q == 1) ? 0 : (Q == 2) ? 1 : (Q == 4) ? 2 : (Q == 8) ? 3 : (Q == 16) ? 4 : 5;

    extern __shared__ __align__(16) float s_xyz_all[];  // CPB x roundup4(n*3) floats
    __shared__ uint2 s_red_all[CPB][2][NW];
    __shared__ __align__(8) uint64_t s_bar_all[CPB];

    const int grp = (CPB == 1) ? 0 : (int)threadIdx.x / T;       // which cloud of the CTA this thread works on
    const int slot = blockIdx.x * CPB + grp;
    if (slot >= nclouds) return;                                 // whole group leaves together (named barriers are per group)
    // variable-size batch: `list` names the clouds of this launch's size class, `counts` their point counts; the cloud
    // stride stays n (the padded size), everything else of the round loop sees the cloud's own count
    const int cloud = list ? __ldg(list + slot) : slot;
    if (skip && skip[cloud]) return;
    const int n_stride = n;
    if (counts) n = __ldg(counts + cloud);
    float *s_xyz = s_xyz_all + (size_t)grp * ((n_stride * 3 + 3) & ~3);
    // exchange slots, as plain pointers computed once (indexing the 3-D array inside the round loop made the compiler
    // re-derive the address behind a branch every round: +75 cycles per round)
    uint2 *const red_base = &s_red_all[grp][0][0];
    uint64_t &s_bar = s_bar_all[grp];
    const float *xyz = xyz_all + (size_t)cloud * n_stride * 3;
    float *temp = FUSED ? nullptr : temp_all + (size_t)cloud * n_stride;
    int *idx = idx_all + (size_t)cloud * m;
    float *new_xyz = FUSED ? new_xyz_all + (size_t)cloud * m * 3 : nullptr;
    const int t = (CPB == 1) ? (int)threadIdx.x : (int)threadIdx.x % T;   // thread within the group
    const int lane = t & 31, warp = t >> 5;
    uint2 *const red_mine = red_base + warp;
    // group-wide barrier: id 1 + grp, T threads (bar.sync with an id lets the CPB groups of a CTA run independently)
    // (CPB == 1 keeps the plain CTA barrier: a barrier named through a register costs ~75 cycles more per round)
    auto group_sync = [&]() {
        if (CPB == 1) __syncthreads();
        else asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "n"(T) : "memory");
    };

    if (t == 0) {
        rt_mbar_init(&s_bar, 1);
        rt_fence_mbar_init();
    }
    group_sync();
    {   // stage the cloud's xyz: bulk copy (TMA engine) when 16-byte aligned, plain loads for the tail / unaligned clouds
        const int nfl = n * 3;
        const int bulk = ((reinterpret_cast<uintptr_t>(xyz) & 15) == 0) ? (nfl & ~3) : 0;
        if (t == 0) {
            if (bulk > 0) {
                rt_mbar_expect_tx(&s_bar, (uint32_t)bulk * 4u);
                for (int off = 0; off < bulk; off += 16384)
                    rt_bulk_g2s(s_xyz + off, xyz + off, (uint32_t)min(16384, bulk - off) * 4u, &s_bar);
            } else {
                rt_mbar_arrive(&s_bar);
            }
        }
        for (int i = bulk + t; i < nfl; i += T) s_xyz[i] = xyz[i];
        rt_mbar_wait(&s_bar, 0);
        group_sync();
    }

    const int r = (int)bitrev_bits((unsigned)t, LOGT);
    float px[PPT], py[PPT], pz[PPT], td[PPT];
    int pk[PPT];
#pragma unroll
    for (int s = 0; s < PPT; ++s) {
        const int a = s / PH, ph = s % PH;
        const int k = r + ((int)bitrev_bits((unsigned)a, LOGQ) + Q * ph) * T;
        pk[s] = (k < n) ? k : -1;
        const int kk = (k < n) ? k : 0;
        px[s] = s_xyz[kk * 3 + 0];
        py[s] = s_xyz[kk * 3 + 1];
        pz[s] = s_xyz[kk * 3 + 2];
        td[s] = (k < n) ? (FUSED ? 1e10f : temp[k]) : -2.0f;  // padding never beats the reference's initial best of -1
    }

    int old = 0;
    if (t == 0) idx[0] = 0;

    for (int j = 1; j < m; ++j) {
        const float x1 = s_xyz[old * 3 + 0], y1 = s_xyz[old * 3 + 1], z1 = s_xyz[old * 3 + 2];
        // fused gather: the coordinates of the point picked in the previous round are in registers right now
        if (FUSED && t < 3) new_xyz[(j - 1) * 3 + t] = t == 0 ? x1 : (t == 1 ? y1 : z1);
        // per-thread arg-max as a tournament over the priority-ordered slots (left wins ties == "first strict
        // maximum" of the reference's scan), so the dependent chain is log2(PPT) deep instead of PPT
        float cd[PPT];
        int ci[PPT];
#pragma unroll
        for (int s = 0; s < PPT; ++s) {
            const float d = rt_sqdist(px[s], py[s], pz[s], x1, y1, z1);
            // padding slots hold td = -2 and stay there (fminf keeps the smaller)
            const float d2 = fminf(d, td[s]);
            td[s] = d2;
            cd[s] = (d2 > -1.0f) ? d2 : -2.0f;   // the reference's best starts at -1: NaN / padding never win
            ci[s] = pk[s];
        }
#pragma unroll
        for (int stride = 1; stride < PPT; stride *= 2)
#pragma unroll
            for (int i = 0; i + stride < PPT; i += 2 * stride)
                if (cd[i + stride] > cd[i]) {
                    cd[i] = cd[i + stride];
                    ci[i] = ci[i + stride];
                }
        const bool any = cd[0] > -1.0f;
        const float best = any ? cd[0] : -1.0f;
        const int besti = any ? ci[0] : 0;
        const uint32_t key = rt_float_ordered(best);
        const uint32_t wmax = rt_redux_max_u32(key);
        const uint32_t vote = __ballot_sync(0xffffffffu, key == wmax);
        const int src = __ffs(vote) - 1;
        const int wi = __shfl_sync(0xffffffffu, besti, src);
        int win;
        if (NW == 1) {
            win = wi;
        } else {
            const int boff = (j & 1) * NW;
            if (lane == 0) red_mine[boff] = make_uint2(wmax, (uint32_t)wi);
            group_sync();
            const uint2 *rb = red_base + boff;
            // every warp picks the winner of the NW warp maxima with one more redux / ballot / shuffle (lane w reads warp
            // w's entry; the lowest warp wins ties, as a serial `>` scan would).  ncu had half of a round's stall samples on
            // the serial scan of the 8 entries (profiles/r2_ncu_fps.txt).  Key 0 is below every ordered key of a finite value.
            const uint2 mine = lane < NW ? rb[lane] : make_uint2(0u, 0u);
            const uint32_t kmax = rt_redux_max_u32(mine.x);
            const uint32_t v2 = __ballot_sync(0xffffffffu, lane < NW && mine.x == kmax);
            win = (int)__shfl_sync(0xffffffffu, mine.y, __ffs(v2) - 1);
        }
        old = win;
        if (t == 0) idx[j] = old;
    }
    if (FUSED && t < 3) new_xyz[(m - 1) * 3 + t] = s_xyz[old * 3 + t];

    if (!FUSED) {
#pragma unroll
        for (int s = 0; s < PPT; ++s)
            if (pk[s] >= 0) temp[pk[s]] = td[s];
    }
}

// One WARP per cloud (32 <= n <= 1024): PPL = bs*PH/32 points per lane in registers, no block barrier and no shared-memory
// exchange at all -- a round is the distance update, a per-lane tournament, redux.sync.max + ballot + one shuffle.  Lane l
// owns the points of priority ranks [l*PPL, (l+1)*PPL) (rank = bitrev(k mod bs) * PH + k div bs, the reference's tie order),
// so "first strict maximum inside the lane, lowest lane among equals" is again the reference's winner.  A CTA carries
// FW_WARPS independent clouds, one per SM sub-partition: every warp has a scheduler to itself, and the sampling of 2b
// clouds blocks 2b/4 SMs instead of 2b (SM-exclusive mode).
constexpr int FW_WARPS = 4;

template <int PPL, bool FUSED>
__global__ void __launch_bounds__(32 * FW_WARPS) fps_warp_kernel(int nclouds, int n, int m, int logbs, int ph_shift,
                                                                 const float *__restrict__ xyz_all, float *__restrict__ temp_all,
                                                                 int *__restrict__ idx_all, float *__restrict__ new_xyz_all) {
    extern __shared__ __align__(16) float s_xyz_all[];  // FW_WARPS x roundup4(n*3) floats
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cloud = blockIdx.x * FW_WARPS + warp;
    if (cloud >= nclouds) return;                       // warps are independent: no CTA-wide barrier anywhere
    float *s_xyz = s_xyz_all + (size_t)warp * ((n * 3 + 3) & ~3);
    const float *xyz = xyz_all + (size_t)cloud * n * 3;
    float *temp = FUSED ? nullptr : temp_all + (size_t)cloud * n;
    int *idx = idx_all + (size_t)cloud * m;
    float *new_xyz = FUSED ? new_xyz_all + (size_t)cloud * m * 3 : nullptr;
    for (int i = lane; i < n * 3; i += 32) s_xyz[i] = __ldg(xyz + i);
    __syncwarp();

    const int bs = 1 << logbs, ph_mask = (1 << ph_shift) - 1;
    auto point_of = [&](int rank) {                     // priority rank -> point index (may be >= n: padding)
        const int a = rank >> ph_shift, ph = rank & ph_mask;
        return (int)(__brev((unsigned)a) >> (32 - logbs)) + ph * bs;
    };
    float px[PPL], py[PPL], pz[PPL], td[PPL];
#pragma unroll
    for (int s = 0; s < PPL; ++s) {
        const int k = point_of(lane * PPL + s);
        const int kk = (k < n) ? k : 0;
        px[s] = s_xyz[kk * 3 + 0];
        py[s] = s_xyz[kk * 3 + 1];
        pz[s] = s_xyz[kk * 3 + 2];
        td[s] = (k < n) ? (FUSED ? 1e10f : temp[k]) : -2.0f;  // padding never beats the reference's initial best of -1
    }
    int old = 0;
    if (lane == 0) idx[0] = 0;
    for (int j = 1; j < m; ++j) {
        const float x1 = s_xyz[old * 3 + 0], y1 = s_xyz[old * 3 + 1], z1 = s_xyz[old * 3 + 2];
        if (FUSED && lane < 3) new_xyz[(j - 1) * 3 + lane] = lane == 0 ? x1 : (lane == 1 ? y1 : z1);
        float cd[PPL];
        int ci[PPL];
#pragma unroll
        for (int s = 0; s < PPL; ++s) {
            const float d2 = fminf(rt_sqdist(px[s], py[s], pz[s], x1, y1, z1), td[s]);
            td[s] = d2;
            cd[s] = (d2 > -1.0f) ? d2 : -2.0f;   // the reference's best starts at -1: NaN / padding never win
            ci[s] = s;
        }
#pragma unroll
        for (int stride = 1; stride < PPL; stride *= 2)
#pragma unroll
            for (int i = 0; i + stride < PPL; i += 2 * stride)
                if (cd[i + stride] > cd[i]) {      // left wins ties: first strict maximum in priority order
                    cd[i] = cd[i + stride];
                    ci[i] = ci[i + stride];
                }
        const bool any = cd[0] > -1.0f;
        const uint32_t key = rt_float_ordered(any ? cd[0] : -1.0f);
        const uint32_t wmax = rt_redux_max_u32(key);
        const uint32_t vote = __ballot_sync(0xffffffffu, key == wmax);
        const int src = __ffs(vote) - 1;
        const int wslot = __shfl_sync(0xffffffffu, any ? ci[0] : -1, src);
        old = wslot >= 0 ? point_of(src * PPL + wslot) : 0;
        if (lane == 0) idx[j] = old;
    }
    if (FUSED && lane < 3) new_xyz[(m - 1) * 3 + lane] = s_xyz[old * 3 + lane];
    if (!FUSED) {
#pragma unroll
        for (int s = 0; s < PPL; ++s) {
            const int k = point_of(lane * PPL + s);
            if (k < n) temp[k] = td[s];
        }
    }
}

template <int PPL, bool FUSED>
int launch_warp2(int b, int n, int m, int logbs, int ph_shift, const float *xyz, float *temp, int *idx, float *new_xyz,
                 cudaStream_t st) {
    size_t smem = (size_t)FW_WARPS * ((n * 3 + 3) & ~3) * sizeof(float);
    if (g_fps_exclusive && smem < kFpsHogBytes) smem = kFpsHogBytes;
    static RtPerDevice attr;   // per kernel instantiation, per device
    if (smem > 40 * 1024 && !attr.done(rt_current_device())) {
        const cudaError_t e = cudaFuncSetAttribute(fps_warp_kernel<PPL, FUSED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFpsHogBytes);
        if (e != cudaSuccess) {
            rt_set_error("fps: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return (int)e;
        }
        attr.mark(rt_current_device());
    }
    fps_warp_kernel<PPL, FUSED><<<(b + FW_WARPS - 1) / FW_WARPS, 32 * FW_WARPS, smem, st>>>(b, n, m, logbs, ph_shift, xyz, temp,
                                                                                            idx, new_xyz);
    return rt_check_launch("fps_warp_kernel");
}
// returns 1 if launched
template <bool FUSED>
int launch_warp(int b, int n, int m, const float *xyz, float *temp, int *idx, float *new_xyz, cudaStream_t st, int *launched) {
    *launched = 0;
    if (n < 32 || n > 1024) return RT_OK;
    const int bs = rt_ref_block_size(n);
    int logbs = 0;
    while ((1 << logbs) < bs) ++logbs;
    const int ph = (n + bs - 1) / bs;                 // 1 or 2 for n <= 1024
    if (ph > 2) return RT_OK;
    const int ppl = bs * ph / 32;
    *launched = 1;
    switch (ppl) {
        case 1: return launch_warp2<1, FUSED>(b, n, m, logbs, ph - 1, xyz, temp, idx, new_xyz, st);
        case 2: return launch_warp2<2, FUSED>(b, n, m, logbs, ph - 1, xyz, temp, idx, new_xyz, st);
        case 4: return launch_warp2<4, FUSED>(b, n, m, logbs, ph - 1, xyz, temp, idx, new_xyz, st);
        case 8: return launch_warp2<8, FUSED>(b, n, m, logbs, ph - 1, xyz, temp, idx, new_xyz, st);
        case 16: return launch_warp2<16, FUSED>(b, n, m, logbs, ph - 1, xyz, temp, idx, new_xyz, st);
        case 32: return launch_warp2<32, FUSED>(b, n, m, logbs, ph - 1, xyz, temp, idx, new_xyz, st);
        default: break;
    }
    *launched = 0;
    return RT_OK;
}

// Generic kernel: any n >= 1, temp kept in global memory (L1/L2 resident), 256 threads.
// Candidate = (ordered(d2) << 32) | ~priority, priority = (bitrev_L(k mod bs) << 21) | (k div bs);
// block-wide max of the 64-bit candidate gives the reference winner.
__global__ void __launch_bounds__(256) fps_generic_kernel(int n, int m, int bs, int logbs,
                                                          const float *__restrict__ xyz_all,
                                                          float *__restrict__ temp_all, int *__restrict__ idx_all) {
    __shared__ unsigned long long s_red[2][8];
    const int cloud = blockIdx.x;
    const float *xyz = xyz_all + (size_t)cloud * n * 3;
    float *temp = temp_all + (size_t)cloud * n;
    int *idx = idx_all + (size_t)cloud * m;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const unsigned long long none = ((unsigned long long)rt_float_ordered(-1.0f) << 32) | 0xffffffffull;

    int old = 0;
    if (t == 0) idx[0] = 0;
    for (int j = 1; j < m; ++j) {
        const float x1 = __ldg(xyz + old * 3 + 0), y1 = __ldg(xyz + old * 3 + 1), z1 = __ldg(xyz + old * 3 + 2);
        unsigned long long cand = none;
        for (int k = t; k < n; k += 256) {
            const float d = rt_sqdist(__ldg(xyz + k * 3 + 0), __ldg(xyz + k * 3 + 1), __ldg(xyz + k * 3 + 2), x1, y1, z1);
            const float d2 = fminf(d, temp[k]);
            temp[k] = d2;
            if (d2 > -1.0f) {  // the reference's per-thread best starts at -1: anything <= -1 (or NaN) never wins
                const unsigned pr = (bitrev_bits((unsigned)(k & (bs - 1)), logbs) << 21) | (unsigned)(k >> logbs);
                const unsigned long long c = ((unsigned long long)rt_float_ordered(d2) << 32) | (unsigned)(~pr);
                cand = c > cand ? c : cand;
            }
        }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            const unsigned long long c = __shfl_xor_sync(0xffffffffu, cand, o);
            cand = c > cand ? c : cand;
        }
        const int buf = j & 1;
        if (lane == 0) s_red[buf][warp] = cand;
        __syncthreads();
        unsigned long long b = s_red[buf][0];
#pragma unroll
        for (int w = 1; w < 8; ++w) {
            const unsigned long long c = s_red[buf][w];
            b = c > b ? c : b;
        }
        if (b == none) {
            old = 0;
        } else {
            const unsigned pr = ~(unsigned)(b & 0xffffffffull);
            old = (int)(((pr & 0x1fffffu) << logbs) | bitrev_bits(pr >> 21, logbs));
        }
        if (t == 0) idx[j] = old;
    }
}

template <int T, int Q, int PH, bool FUSED, int CPB>
int launch_reg3(int b, int n, int m, const float *xyz, float *temp, int *idx, float *new_xyz, cudaStream_t st) {
    const int n_kernel = g_fps_stride ? g_fps_stride : n;   // variable-size batch: clouds are strided by the padded size
    size_t smem = (size_t)CPB * ((n_kernel * 3 + 3) & ~3) * sizeof(float);
    // SM-exclusive mode (rt_fps_set_exclusive): every round of this kernel is a dependent latency chain, and any CTA
    // sharing the SM stretches each round 2-3x (measured on B200).  Asking for (nearly) the whole shared memory of the
    // SM keeps every other CTA off it, so the chain runs at its stand-alone speed whatever else is in flight.
    if (g_fps_exclusive && smem < kFpsHogBytes) smem = kFpsHogBytes;
    static RtPerDevice attr;   // per kernel instantiation, per device
    if (smem > 40 * 1024 && !attr.done(rt_current_device())) {
        const cudaError_t e = cudaFuncSetAttribute(fps_reg_kernel<T, Q, PH, FUSED, CPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFpsHogBytes);
        if (e != cudaSuccess) {
            rt_set_error("fps: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return (int)e;
        }
        attr.mark(rt_current_device());
    }
    fps_reg_kernel<T, Q, PH, FUSED, CPB><<<(b + CPB - 1) / CPB, T * CPB, smem, st>>>(b, n_kernel, m, xyz, temp, idx, new_xyz, g_fps_skip, g_fps_counts, g_fps_list);
    return rt_check_launch("fps_reg_kernel");
}
template <int T, int Q, int PH, bool FUSED>
int launch_reg2(int b, int n, int m, const float *xyz, float *temp, int *idx, float *new_xyz, cudaStream_t st) {
    // two clouds per CTA only where it pays: SM-exclusive mode (each CTA blocks a whole SM) and 256-thread groups
    if (g_fps_exclusive && g_fps_pair && T == 256 && b >= 2 && (size_t)2 * ((n * 3 + 3) & ~3) * sizeof(float) <= kFpsHogBytes)
        return launch_reg3<T, Q, PH, FUSED, (T == 256 ? 2 : 1)>(b, n, m, xyz, temp, idx, new_xyz, st);
    return launch_reg3<T, Q, PH, FUSED, 1>(b, n, m, xyz, temp, idx, new_xyz, st);
}
template <int T, int Q, int PH>
int launch_reg(int b, int n, int m, const float *xyz, float *temp, int *idx, float *new_xyz, cudaStream_t st) {
    return new_xyz ? launch_reg2<T, Q, PH, true>(b, n, m, xyz, temp, idx, new_xyz, st)
                   : launch_reg2<T, Q, PH, false>(b, n, m, xyz, temp, idx, nullptr, st);
}


// ---- FPS of a cloud that already IS in FPS order ------------------------------------------------------------------------
// Levels 2 and 3 of PNHead sample npoint = 512 points out of the 512 points level 1 has just selected, starting again at
// point 0 (reference: src/utils/model_utils/model_utils.py:397-399, lib/pointnet2_modules.py:30-35).  Let Q_0..Q_{n-1} be a
// cloud in FPS order and M_j[k] = min(init_k, min_{i<j} d(Q_k, Q_i)) the running distance before round j.  The sampler picks
// idx[j] = argmax_k M_j[k].  Because Q is in FPS order, Q_j attained that maximum over a SUPERSET of Q when it was selected,
// with bit-identical distance arithmetic -- so the result is the identity permutation unless another Q_k ties with Q_j at
// round j and wins the reference's tie-break (lowest (bitrev(k mod bs), k div bs)).  This kernel CHECKS exactly that
// condition, round by round but for all rounds in parallel (every (k, j) test is independent given the prefix minima):
//   pass 1   D[k] = M_k[k]                       (what candidate k holds at the round it must win)
//   pass 2   for j < k: M_j[k] < D[j], or M_j[k] == D[j] and k loses the tie-break to j     (k not yet selected)
//            for j > k: M_j[k] = 0 <= D[j]; conservatively D[j] > 0 is required (no exhausted / duplicate rounds)
// A cloud that passes gets idx = 0..n-1 (and new_xyz = xyz, twice when the next level repeats the same question) and its
// flag set, so the serial sampler (511 dependent rounds, ~146 us) skips it; a cloud that fails is left to the serial kernel.
// Not a heuristic: where the flag is set the serial kernel would have produced the same indices bit for bit.
__global__ void __launch_bounds__(1024) fps_identity_kernel(int n, int bs, int logbs, const float *__restrict__ xyz_all,
                                                            float *__restrict__ temp_all, int *__restrict__ ok_all,
                                                            int *__restrict__ idx_a, float *__restrict__ xyz_a,
                                                            int *__restrict__ idx_b, float *__restrict__ xyz_b) {
    extern __shared__ __align__(16) float s_id[];   // n x 3 coordinates, n diagonal values
    float *q = s_id, *dg = s_id + 3 * n;
    __shared__ int s_fail;
    const int cloud = blockIdx.x, k = threadIdx.x;
    const float *xyz = xyz_all + (size_t)cloud * n * 3;
    if (k == 0) s_fail = 0;
    for (int i = k; i < 3 * n; i += blockDim.x) q[i] = xyz[i];
    __syncthreads();
    const bool live = k < n;
    const float px = live ? q[3 * k] : 0.0f, py = live ? q[3 * k + 1] : 0.0f, pz = live ? q[3 * k + 2] : 0.0f;
    const float init = (temp_all && live) ? temp_all[(size_t)cloud * n + k] : 1e10f;
    bool fail = false;
    float m = init;
    if (live) {
        for (int i = 0; i < k; ++i) m = fminf(m, rt_sqdist(px, py, pz, q[3 * i], q[3 * i + 1], q[3 * i + 2]));
        dg[k] = m;
        // NaN / infinite coordinates, exhausted rounds (duplicates) and a non-standard initial state go to the serial kernel
        fail = !(fabsf(px) < 1e18f && fabsf(py) < 1e18f && fabsf(pz) < 1e18f) || !(init >= 0.0f) || (k >= 1 && !(m > 0.0f));
    }
    __syncthreads();
    if (live && !fail) {
        const unsigned pk = (__brev((unsigned)(k & (bs - 1))) >> (32 - logbs)) * 4096u + (unsigned)(k >> logbs);
        m = init;
        for (int j = 1; j < k; ++j) {
            m = fminf(m, rt_sqdist(px, py, pz, q[3 * j - 3], q[3 * j - 2], q[3 * j - 1]));   // M_j[k]
            const float d = dg[j];
            if (m >= d) {
                const unsigned pj = (__brev((unsigned)(j & (bs - 1))) >> (32 - logbs)) * 4096u + (unsigned)(j >> logbs);
                fail = fail || m > d || pk < pj;
            }
        }
    }
    if (fail) s_fail = 1;
    __syncthreads();
    const bool ok = s_fail == 0;
    if (k == 0) ok_all[cloud] = ok ? 1 : 0;
    if (!ok) return;
    for (int i = k; i < n; i += blockDim.x) {
        idx_a[(size_t)cloud * n + i] = i;
        if (idx_b) idx_b[(size_t)cloud * n + i] = i;
    }
    for (int i = k; i < 3 * n; i += blockDim.x) {
        if (xyz_a) xyz_a[(size_t)cloud * n * 3 + i] = q[i];
        if (xyz_b) xyz_b[(size_t)cloud * n * 3 + i] = q[i];
    }
    // what the serial kernel leaves in `temp`: the running minimum after the last round (every point but the last has met itself)
    if (temp_all && live) temp_all[(size_t)cloud * n + k] = k < n - 1 ? fminf(init, 0.0f) : dg[k];
}

}  // namespace

// engine-internal: 1 = FPS CTAs claim a whole SM each (see launch_reg)
void rt_fps_set_exclusive(int on) { g_fps_exclusive = on & 1; g_fps_pair = (on >> 1) & 1; }   // bit 1: two clouds per CTA
// engine-internal: clouds whose flag is set are skipped by the next register-resident FPS launches (nullptr = none)
void rt_fps_set_skip(const int *flags) { g_fps_skip = flags; }

// engine-internal: identity test of FPS(n points -> n samples) for clouds already in FPS order (see fps_identity_kernel).
// ok (b) flags; idx_a / xyz_a (and optionally idx_b / xyz_b for a second identical level) are written for the clouds that pass.
// temp: the caller's running-minimum buffer (C ABI) or nullptr for the reference's initial 1e10.
int rt_launch_fps_identity(int b, int n, const float *xyz, float *temp, int *ok, int *idx_a, float *xyz_a, int *idx_b, float *xyz_b,
                           cudaStream_t st) {
    if (b == 0) return RT_OK;
    RT_REQUIRE(n >= 32 && n <= 1024, "fps_identity: n=%d", n);
    const int bs = rt_ref_block_size(n);
    int logbs = 0;
    while ((1 << logbs) < bs) ++logbs;
    const int threads = (n + 31) / 32 * 32;
    fps_identity_kernel<<<b, threads, (size_t)4 * n * sizeof(float), st>>>(n, bs, logbs, xyz, temp, ok, idx_a, xyz_a, idx_b, xyz_b);
    return rt_check_launch("fps_identity_kernel");
}

// register-resident variants; returns 1 when one was launched, 0 when the shape needs the generic kernel, < 0 / > 0 on error
static int fps_dispatch(int b, int n, int m, const float *xyz, float *temp, int *idx, float *new_xyz, cudaStream_t st, int *launched) {
    // warp-per-cloud kernel (n <= 1024) only with RT_FPS_WARP=1 (A/B timing): measured on B200 it frees 3/4 of the SMs the
    // sampling blocks but its rounds are slower (245 vs 138 us at n=1024, 204 vs 118 us at n=512: one warp issues its 32
    // point updates far below one instruction per cycle), and the longer chain costs the step 6 %
    static int use_warp = -1;
    if (use_warp < 0) {
        const char *env = getenv("RT_FPS_WARP");
        use_warp = (env && atoi(env) == 1) ? 1 : 0;
    }
    if (use_warp && !g_fps_skip) {
        const int rc = new_xyz ? launch_warp<true>(b, n, m, xyz, temp, idx, new_xyz, st, launched)
                               : launch_warp<false>(b, n, m, xyz, temp, idx, nullptr, st, launched);
        if (rc != RT_OK || *launched) return rc;
    }
    const int bs = rt_ref_block_size(n);
    const int ph = (n + bs - 1) / bs;
    // T threads, Q = bs / T residues per thread, PH = ceil(n / bs) passes.  T = 256 halves the per-thread work of a round
    // (the rounds are a pure latency chain); RT_FPS_THREADS=128 selects the 4-warp variant for A/B timing.
    static int pref_t = 0;
    if (!pref_t) {
        const char *env = getenv("RT_FPS_THREADS");
        pref_t = (env && atoi(env) == 128) ? 128 : 256;
    }
    *launched = 1;
#define RT_FPS_CASE(TT, QQ, PP) \
    if (t == TT && q == QQ && ph == PP) return launch_reg<TT, QQ, PP>(b, n, m, xyz, temp, idx, new_xyz, st);
    if (bs >= 128 && ph <= 4) {
        const int t = (bs >= 256 && pref_t == 256) ? 256 : 128;
        const int q = bs / t;
        RT_FPS_CASE(256, 1, 1) RT_FPS_CASE(256, 1, 2)
        RT_FPS_CASE(256, 2, 1) RT_FPS_CASE(256, 2, 2)
        RT_FPS_CASE(256, 4, 1) RT_FPS_CASE(256, 4, 2) RT_FPS_CASE(256, 4, 3) RT_FPS_CASE(256, 4, 4)
        RT_FPS_CASE(128, 1, 1) RT_FPS_CASE(128, 1, 2)
        RT_FPS_CASE(128, 2, 1) RT_FPS_CASE(128, 2, 2)
        RT_FPS_CASE(128, 4, 1) RT_FPS_CASE(128, 4, 2)
        RT_FPS_CASE(128, 8, 1) RT_FPS_CASE(128, 8, 2)
        if (bs == 1024 && ph <= 4) {  // 2048 < n <= 4096 with the 128-thread preference: 256 threads x 16 points
            if (ph == 3) return launch_reg<256, 4, 3>(b, n, m, xyz, temp, idx, new_xyz, st);
            return launch_reg<256, 4, 4>(b, n, m, xyz, temp, idx, new_xyz, st);
        }
    }
#undef RT_FPS_CASE
    *launched = 0;
    return RT_OK;
}

// template selection by n_class (the size class's largest cloud), cloud stride / staging by n_stride
static int fps_dispatch_sized(int b, int n_class, int n_stride, int m, const float *xyz, float *temp, int *idx, float *new_xyz,
                              cudaStream_t st, int *launched) {
    g_fps_stride = n_stride;
    const int rc = fps_dispatch(b, n_class, m, xyz, temp, idx, new_xyz, st, launched);
    g_fps_stride = 0;
    return rc;
}

// engine-internal: FPS from the reference's initial state (temp = 1e10) that also emits new_xyz (b,m,3) = xyz[idx].
// Returns RT_ERR_UNSUPPORTED when the shape has no register-resident variant (caller falls back to fill + FPS + gather).
int rt_launch_fps_fused(int b, int n, int m, const float *xyz, int *idx, float *new_xyz, cudaStream_t st) {
    if (m <= 0 || b == 0) return RT_OK;
    int launched = 0;
    const int rc = fps_dispatch(b, n, m, xyz, nullptr, idx, new_xyz, st, &launched);
    if (rc != RT_OK) return rc;
    return launched ? RT_OK : RT_ERR_UNSUPPORTED;
}

// engine-internal: variable-size batch (see engine_kernels.cuh).  One launch per size class bs = rt_ref_block_size(count):
// the register-resident kernel for (bs, passes of the class's largest cloud) reproduces, for every cloud of the class, exactly
// what a stand-alone launch for that cloud's own size computes (slots beyond a cloud's count are padding).
int rt_launch_fps_fused_varlen(int b, int n_stride, int m, const float *xyz, const int *counts_host, const int *counts_dev,
                               int *list_dev, int *idx, float *new_xyz, cudaStream_t st) {
    if (m <= 0 || b == 0) return RT_OK;
    RT_REQUIRE(b <= 65535, "fps_varlen: batch > 65535");
    int *order = (int *)malloc(sizeof(int) * (size_t)b);
    RT_REQUIRE(order, "fps_varlen: out of host memory");
    int filled = 0, rc = RT_OK;
    for (int bs = 1; bs <= 1024 && rc == RT_OK; bs *= 2) {
        const int first = filled;
        int nmax = 0;
        for (int i = 0; i < b; ++i)
            if (rt_ref_block_size(counts_host[i]) == bs) {
                order[filled++] = i;
                nmax = counts_host[i] > nmax ? counts_host[i] : nmax;
            }
        if (filled == first) continue;
        const cudaError_t ce = cudaMemcpyAsync(list_dev + first, order + first, sizeof(int) * (size_t)(filled - first), cudaMemcpyHostToDevice, st);
        if (ce != cudaSuccess) {
            rt_set_error("fps_varlen: %s", cudaGetErrorString(ce));
            rc = (int)ce;
            break;
        }
        // dispatch on the class's largest cloud (it fixes bs and the number of passes); the kernel strides clouds by n_stride
        g_fps_counts = counts_dev;
        g_fps_list = list_dev + first;
        int launched = 0;
        rc = fps_dispatch_sized(filled - first, nmax, n_stride, m, xyz, nullptr, idx, new_xyz, st, &launched);
        g_fps_counts = g_fps_list = nullptr;
        if (rc == RT_OK && !launched) rc = RT_ERR_UNSUPPORTED;
    }
    // pageable host memory: the copies above were staged before cudaMemcpyAsync returned
    free(order);
    if (rc == RT_OK && filled != b) {
        rt_set_error("fps_varlen: a cloud has %s points", "0 or more than the register-resident kernels cover");
        rc = RT_ERR_UNSUPPORTED;
    }
    return rc;
}

// C-ABI.  replaces furthest_point_sampling_wrapper (reference: src/lib/src/sampling.cpp:37-47)
RT_API int rt_furthest_point_sampling(int b, int n, int m, const float *xyz, float *temp, int *idx, void *stream) {
    RT_REQUIRE(b >= 0 && n >= 1 && xyz && temp && idx, "furthest_point_sampling: bad arguments (b=%d n=%d)", b, n);
    if (m <= 0 || b == 0) return RT_OK;  // reference kernel returns early on m <= 0
    cudaStream_t st = (cudaStream_t)stream;
    int launched = 0;
    // sampling ALL points of a cloud (the second and third set-abstraction levels: 512 of 512) is the identity whenever the
    // cloud is already in FPS order and no round ties -- checked exactly by fps_identity_kernel; clouds that fail the check
    // (and every other shape) take the serial kernel
    int *ok = nullptr;
    if (m == n && n >= 128 && n <= 1024 && b <= 65535) {
        if (rt_scratch_alloc((void **)&ok, sizeof(int) * (size_t)b, st, "furthest_point_sampling") == RT_OK) {
            const int rc0 = rt_launch_fps_identity(b, n, xyz, temp, ok, idx, nullptr, nullptr, nullptr, st);
            if (rc0 != RT_OK) {
                rt_scratch_free(ok, st);
                return rc0;
            }
        } else {
            ok = nullptr;   // no scratch: the serial kernel alone is always correct
            (void)cudaGetLastError();
        }
    }
    g_fps_skip = ok;
    const int rc = fps_dispatch(b, n, m, xyz, temp, idx, nullptr, st, &launched);
    g_fps_skip = nullptr;
    rt_scratch_free(ok, st);
    if (rc != RT_OK || launched) return rc;
    RT_REQUIRE(!ok, "furthest_point_sampling: internal (identity shortcut without a register-resident kernel)");
    const int bs = rt_ref_block_size(n);
    int logbs = 0;
    while ((1 << logbs) < bs) ++logbs;
    RT_REQUIRE((n >> logbs) < (1 << 21), "furthest_point_sampling: n too large");
    fps_generic_kernel<<<b, 256, 0, st>>>(n, m, bs, logbs, xyz, temp, idx);
    return rt_check_launch("fps_generic_kernel");
}
