// Building-block kernels of the fused Track4D.backbone engine (sm_100a).  See engine_kernels.cuh for the
// contracts and engine.cu for how they are chained.  Everything here is fp32 SIMT; the dense 256-wide
// cost-volume MLP has its own tensor-core kernel (costvol_tc.cu).
#include <math.h>

#include <stdlib.h>

#include "engine_kernels.cuh"

namespace {

__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == RT_ACT_RELU) return fmaxf(v, 0.0f);
    if (act == RT_ACT_LEAKY01) return v > 0.0f ? v : 0.1f * v;
    return v;
}

// ------------------------------------------------------------------------------------------------
// row GEMM: CTA tile 64 rows x 64 outputs, 256 threads, 4x4 register tile, K staged 16 at a time.
constexpr int RG_T = 64, RG_K = 16, RG_LD = RG_T + 4;

__global__ void __launch_bounds__(256) rowgemm_kernel(RtRowGemm a) {
    __shared__ __align__(16) float sX[RG_K][RG_LD];
    __shared__ __align__(16) float sW[RG_K][RG_LD];
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const long long row0 = (long long)blockIdx.x * RG_T;
    const int out0 = blockIdx.y * RG_T;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

    for (int s = 0; s < a.nseg; ++s) {
        const RtSeg sg = a.seg[s];
        for (int k0 = 0; k0 < sg.k; k0 += RG_K) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int e = tid + 256 * i;
                const int r = e >> 4, kk = e & 15;
                const long long gr = row0 + r;
                const int gk = k0 + kk;
                float xv = 0.0f, wv = 0.0f;
                if (gk < sg.k) {
                    if (gr < a.rows) xv = __ldg(sg.x + gr * sg.ldx + gk);
                    if (out0 + r < a.nout) wv = __ldg(sg.w + (long long)(out0 + r) * sg.ldw + gk);
                }
                sX[kk][r] = xv;
                sW[kk][r] = wv;
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < RG_K; ++kk) {
                const float4 xa = *reinterpret_cast<const float4 *>(&sX[kk][4 * ty]);
                const float4 wb = *reinterpret_cast<const float4 *>(&sW[kk][4 * tx]);
                const float xr[4] = {xa.x, xa.y, xa.z, xa.w};
                const float wr[4] = {wb.x, wb.y, wb.z, wb.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xr[i], wr[j], acc[i][j]);
            }
            __syncthreads();
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long gr = row0 + 4 * ty + i;
        if (gr >= a.rows) continue;
        const float *cb = a.cloud_bias ? a.cloud_bias + (gr / a.rows_per_cloud) * a.nout : nullptr;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int o = out0 + 4 * tx + j;
            if (o >= a.nout) continue;
            float v = acc[i][j];
            if (a.bias) v += __ldg(a.bias + o);
            if (cb) v += __ldg(cb + o);
            a.y[gr * a.ldy + o] = apply_act(v, a.act);
        }
    }
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gather_combine_kernel(RtGatherCombine a) {
    const long long total = (long long)a.clouds * a.npts * a.ns * a.c;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int ch = (int)(e % a.c);
        const long long row = e / a.c;              // (cloud, p, s)
        const long long cp = row / a.ns;            // (cloud, p)
        const int cloud = (int)(cp / a.npts);
        const int j = __ldg(a.idx + row);
        const float *pi = a.xyz_in + ((long long)cloud * a.n_in + j) * 3;
        const float *pc = a.xyz_c + cp * 3;
        const float dx = __ldg(pi + 0) - __ldg(pc + 0), dy = __ldg(pi + 1) - __ldg(pc + 1), dz = __ldg(pi + 2) - __ldg(pc + 2);
        float v = __ldg(a.y + ((long long)cloud * a.n_in + j) * a.ldy + a.yoff + ch);
        const float *w = a.wx + ch * 3;
        v += fmaf(__ldg(w + 2), dz, fmaf(__ldg(w + 1), dy, __ldg(w + 0) * dx));
        if (a.bias) v += __ldg(a.bias + ch);
        if (a.q) v += __ldg(a.q + cp * a.c + ch);
        a.out[e] = apply_act(v, a.act);
    }
}

__global__ void __launch_bounds__(256) maxpool_rows_kernel(long long groups, int ns, int c, const float *__restrict__ x,
                                                           float *__restrict__ pooled, int ldp, int coff) {
    const long long total = groups * c;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int ch = (int)(e % c);
        const long long g = e / c;
        const float *src = x + g * ns * c + ch;
        float m = __ldg(src);
        for (int s = 1; s < ns; ++s) m = fmaxf(m, __ldg(src + (long long)s * c));
        pooled[g * ldp + coff + ch] = m;
    }
}

// ------------------------------------------------------------------------------------------------
// WeightNet-weighted neighbour sum.  CTA = 8 points, 256 threads.  Phase 1: one thread per (point,
// neighbour) evaluates the 3->8->8 trunk into shared memory; phase 2: one thread per channel.
constexpr int WS_PTS = 8;


__global__ void __launch_bounds__(256) weighted_sum_generic_kernel(RtWeightedSum a) {
    extern __shared__ __align__(16) float s_h2[];  // WS_PTS * ns * 8 trunk outputs, each stored twice (h, h): FFMA2 operands
    __shared__ int s_idx[WS_PTS * 32];
    const long long cp0 = (long long)blockIdx.x * WS_PTS;   // first SLOT of the CTA; slot -> point through a.perm
    const long long ncp = (long long)a.clouds * a.npts;
    const int pairs = WS_PTS * a.ns;
    __shared__ long long s_cp[WS_PTS];
    if (threadIdx.x < WS_PTS) {
        const long long slot = cp0 + threadIdx.x;
        s_cp[threadIdx.x] = slot < ncp ? (a.perm ? (long long)__ldg(a.perm + slot) : slot) : ncp;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < pairs; t += blockDim.x) {
        const long long cp = s_cp[t / a.ns];
        float h2[8];
        int j = 0;
        if (cp < ncp) {
            const int cloud = (int)(cp / a.npts);
            j = __ldg(a.idx + cp * a.ns + (t % a.ns));
            const float *pi = a.xyz_in + ((long long)cloud * a.n_in + j) * 3;
            const float *pc = a.xyz_c + cp * 3;
            const float d[3] = {__ldg(pi + 0) - __ldg(pc + 0), __ldg(pi + 1) - __ldg(pc + 1), __ldg(pi + 2) - __ldg(pc + 2)};
            float h1[8];
#pragma unroll
            for (int o = 0; o < 8; ++o) {
                float v = __ldg(a.ba + o);
#pragma unroll
                for (int k = 0; k < 3; ++k) v = fmaf(__ldg(a.wa + o * 3 + k), d[k], v);
                h1[o] = fmaxf(v, 0.0f);
            }
#pragma unroll
            for (int o = 0; o < 8; ++o) {
                float v = __ldg(a.bb + o);
#pragma unroll
                for (int k = 0; k < 8; ++k) v = fmaf(__ldg(a.wb + o * 8 + k), h1[k], v);
                h2[o] = fmaxf(v, 0.0f);
            }
        } else {
#pragma unroll
            for (int o = 0; o < 8; ++o) h2[o] = 0.0f;
        }
#pragma unroll
        for (int o = 0; o < 8; ++o) reinterpret_cast<float2 *>(s_h2)[t * 8 + o] = make_float2(h2[o], h2[o]);
        s_idx[t] = j;
    }
    __syncthreads();
    if ((a.c & 3) == 0 && a.c <= 256) {
        // 4 channels per thread (128-bit gathers: 4x the bytes in flight per thread), 256 / (c/4) points side by side
        const int lanes = a.c >> 2;                 // threads per point
        const int pslots = blockDim.x / lanes;      // points processed concurrently
        const int cg = threadIdx.x % lanes, ps = threadIdx.x / lanes;
        if (ps < pslots) {
            // channel pairs (4cg, 4cg+1) and (4cg+2, 4cg+3): the 8 -> c layer runs as packed FFMA2, same per-channel
            // operation order as a scalar fmaf chain (w = bc; w = fma(wc[k], h2[k], w), k ascending)
            float2 wc[2][8], bc[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    wc[j][k] = make_float2(__ldg(a.wc + (4 * cg + 2 * j) * 8 + k), __ldg(a.wc + (4 * cg + 2 * j + 1) * 8 + k));
                bc[j] = make_float2(__ldg(a.bc + 4 * cg + 2 * j), __ldg(a.bc + 4 * cg + 2 * j + 1));
            }
            for (int p = ps; p < WS_PTS; p += pslots) {
                const long long cp = s_cp[p];
                if (cp >= ncp) break;
                const int cloud = (int)(cp / a.npts);
                float2 acc[2] = {make_float2(0.0f, 0.0f), make_float2(0.0f, 0.0f)};
                for (int s0 = 0; s0 < a.ns; s0 += 8) {
                    float4 vv[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int s = s0 + u;
                        vv[u] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                        if (s < a.ns) {
                            const float *row = a.gather_v ? a.v + ((long long)cloud * a.n_in + s_idx[p * a.ns + s]) * a.c
                                                          : a.v + (cp * a.ns + s) * a.c;
                            vv[u] = __ldg(reinterpret_cast<const float4 *>(row) + cg);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int s = s0 + u;
                        if (s < a.ns) {
                            const float4 *h2 = reinterpret_cast<const float4 *>(s_h2) + (p * a.ns + s) * 4;   // 8 x (h, h)
                            float2 w0 = bc[0], w1 = bc[1];
#pragma unroll
                            for (int k2 = 0; k2 < 4; ++k2) {
                                const float4 hh = h2[k2];
                                w0 = rt_ffma2(wc[0][2 * k2], make_float2(hh.x, hh.y), w0);
                                w1 = rt_ffma2(wc[1][2 * k2], make_float2(hh.x, hh.y), w1);
                                w0 = rt_ffma2(wc[0][2 * k2 + 1], make_float2(hh.z, hh.w), w0);
                                w1 = rt_ffma2(wc[1][2 * k2 + 1], make_float2(hh.z, hh.w), w1);
                            }
                            acc[0] = rt_ffma2(make_float2(fmaxf(w0.x, 0.0f), fmaxf(w0.y, 0.0f)), make_float2(vv[u].x, vv[u].y), acc[0]);
                            acc[1] = rt_ffma2(make_float2(fmaxf(w1.x, 0.0f), fmaxf(w1.y, 0.0f)), make_float2(vv[u].z, vv[u].w), acc[1]);
                        }
                    }
                }
                reinterpret_cast<float4 *>(a.out + cp * a.c)[cg] = make_float4(acc[0].x, acc[0].y, acc[1].x, acc[1].y);
            }
        }
        return;
    }
    for (int ch = threadIdx.x; ch < a.c; ch += blockDim.x) {
        float wc[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) wc[k] = __ldg(a.wc + ch * 8 + k);
        const float bc = __ldg(a.bc + ch);
        for (int p = 0; p < WS_PTS; ++p) {
            const long long cp = s_cp[p];
            if (cp >= ncp) break;
            const int cloud = (int)(cp / a.npts);
            float acc = 0.0f;
            for (int s = 0; s < a.ns; ++s) {
                const float *h2 = s_h2 + (p * a.ns + s) * 16;   // (h, h) pairs
                float w = bc;
#pragma unroll
                for (int k = 0; k < 8; ++k) w = fmaf(wc[k], h2[2 * k], w);
                const float v = a.gather_v ? __ldg(a.v + ((long long)cloud * a.n_in + s_idx[p * a.ns + s]) * a.c + ch)
                                           : __ldg(a.v + (cp * a.ns + s) * a.c + ch);
                acc = fmaf(fmaxf(w, 0.0f), v, acc);
            }
            a.out[cp * a.c + ch] = acc;
        }
    }
}

// Persistent variant for c % 4 == 0, c <= 256 (the cost volume: c = 256).  A CTA loops over groups of WS_PTS points; the
// 8 -> c weights of its channel pairs stay in REGISTERS for the whole kernel and the trunk weights in shared memory,
// loaded once per CTA -- in the one-group-per-CTA kernel above the per-CTA reload of wc (a 128-byte-strided, fully
// uncoalesced read: 32 L1 wavefronts per load) made up 85 % of the L1 data-pipe traffic that bounds this operation.
__global__ void __launch_bounds__(256) weighted_sum_kernel(RtWeightedSum a) {
    extern __shared__ __align__(16) float s_h2[];  // WS_PTS * ns * 8 trunk outputs, each stored twice (h, h): FFMA2 operands
    __shared__ int s_idx[WS_PTS * 32];
    __shared__ long long s_cp[WS_PTS];
    __shared__ float s_w[104];                     // wa (24) ba (8) wb (64) bb (8)
    const long long ncp = (long long)a.clouds * a.npts;
    const long long ngroups = (ncp + WS_PTS - 1) / WS_PTS;
    const int pairs = WS_PTS * a.ns;
    const int lanes = a.c >> 2;                 // threads per point: 4 channels each
    const int pslots = blockDim.x / lanes;      // points processed side by side
    const int cg = threadIdx.x % lanes, ps = threadIdx.x / lanes;
    if (threadIdx.x < 24) s_w[threadIdx.x] = __ldg(a.wa + threadIdx.x);
    if (threadIdx.x < 8) s_w[24 + threadIdx.x] = __ldg(a.ba + threadIdx.x);
    if (threadIdx.x < 64) s_w[32 + threadIdx.x] = __ldg(a.wb + threadIdx.x);
    if (threadIdx.x < 8) s_w[96 + threadIdx.x] = __ldg(a.bb + threadIdx.x);
    // channel pairs (4cg, 4cg+1) and (4cg+2, 4cg+3): the 8 -> c layer runs as packed FFMA2, same per-channel
    // operation order as a scalar fmaf chain (w = bc; w = fma(wc[k], h2[k], w), k ascending)
    float2 wc[2][8], bc[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
            wc[j][k] = make_float2(__ldg(a.wc + (4 * cg + 2 * j) * 8 + k), __ldg(a.wc + (4 * cg + 2 * j + 1) * 8 + k));
        bc[j] = make_float2(__ldg(a.bc + 4 * cg + 2 * j), __ldg(a.bc + 4 * cg + 2 * j + 1));
    }
    for (long long grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
        const long long cp0 = grp * WS_PTS;   // first SLOT of the group; slot -> point through a.perm
        __syncthreads();                      // previous group's phase 2 is done with s_h2 / s_idx / s_cp (and s_w is written)
        if (threadIdx.x < WS_PTS) {
            const long long slot = cp0 + threadIdx.x;
            s_cp[threadIdx.x] = slot < ncp ? (a.perm ? (long long)__ldg(a.perm + slot) : slot) : ncp;
        }
        __syncthreads();
        // phase 1: one thread per (point, neighbour) evaluates the 3 -> 8 -> 8 trunk
        for (int t = threadIdx.x; t < pairs; t += blockDim.x) {
            const long long cp = s_cp[t / a.ns];
            float h2[8];
            int j = 0;
            if (cp < ncp) {
                const int cloud = (int)(cp / a.npts);
                j = __ldg(a.idx + cp * a.ns + (t % a.ns));
                const float *pi = a.xyz_in + ((long long)cloud * a.n_in + j) * 3;
                const float *pc = a.xyz_c + cp * 3;
                const float d[3] = {__ldg(pi + 0) - __ldg(pc + 0), __ldg(pi + 1) - __ldg(pc + 1), __ldg(pi + 2) - __ldg(pc + 2)};
                float h1[8];
#pragma unroll
                for (int o = 0; o < 8; ++o) {
                    float v = s_w[24 + o];
#pragma unroll
                    for (int k = 0; k < 3; ++k) v = fmaf(s_w[o * 3 + k], d[k], v);
                    h1[o] = fmaxf(v, 0.0f);
                }
#pragma unroll
                for (int o = 0; o < 8; ++o) {
                    float v = s_w[96 + o];
#pragma unroll
                    for (int k = 0; k < 8; ++k) v = fmaf(s_w[32 + o * 8 + k], h1[k], v);
                    h2[o] = fmaxf(v, 0.0f);
                }
            } else {
#pragma unroll
                for (int o = 0; o < 8; ++o) h2[o] = 0.0f;
            }
#pragma unroll
            for (int o = 0; o < 8; ++o) reinterpret_cast<float2 *>(s_h2)[t * 8 + o] = make_float2(h2[o], h2[o]);
            s_idx[t] = j;
        }
        __syncthreads();
        // phase 2: 128-bit gathers of the value rows, weights from the trunk outputs, sum over the neighbours
        if (ps < pslots) {
            for (int p = ps; p < WS_PTS; p += pslots) {
                const long long cp = s_cp[p];
                if (cp >= ncp) break;
                const int cloud = (int)(cp / a.npts);
                float2 acc[2] = {make_float2(0.0f, 0.0f), make_float2(0.0f, 0.0f)};
                for (int s0 = 0; s0 < a.ns; s0 += 8) {
                    float4 vv[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int s = s0 + u;
                        vv[u] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                        if (s < a.ns) {
                            const float *row = a.gather_v ? a.v + ((long long)cloud * a.n_in + s_idx[p * a.ns + s]) * a.c
                                                          : a.v + (cp * a.ns + s) * a.c;
                            vv[u] = __ldg(reinterpret_cast<const float4 *>(row) + cg);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int s = s0 + u;
                        if (s < a.ns) {
                            const float4 *h2 = reinterpret_cast<const float4 *>(s_h2) + (p * a.ns + s) * 4;   // 8 x (h, h)
                            float2 w0 = bc[0], w1 = bc[1];
#pragma unroll
                            for (int k2 = 0; k2 < 4; ++k2) {
                                const float4 hh = h2[k2];
                                w0 = rt_ffma2(wc[0][2 * k2], make_float2(hh.x, hh.y), w0);
                                w1 = rt_ffma2(wc[1][2 * k2], make_float2(hh.x, hh.y), w1);
                                w0 = rt_ffma2(wc[0][2 * k2 + 1], make_float2(hh.z, hh.w), w0);
                                w1 = rt_ffma2(wc[1][2 * k2 + 1], make_float2(hh.z, hh.w), w1);
                            }
                            acc[0] = rt_ffma2(make_float2(fmaxf(w0.x, 0.0f), fmaxf(w0.y, 0.0f)), make_float2(vv[u].x, vv[u].y), acc[0]);
                            acc[1] = rt_ffma2(make_float2(fmaxf(w1.x, 0.0f), fmaxf(w1.y, 0.0f)), make_float2(vv[u].z, vv[u].w), acc[1]);
                        }
                    }
                }
                reinterpret_cast<float4 *>(a.out + cp * a.c)[cg] = make_float4(acc[0].x, acc[0].y, acc[1].x, acc[1].y);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// expanded-form kNN: one thread per query, search cloud staged as (x, y, z, |p|^2) in shared memory.
constexpr int KX_THREADS = 128, KX_TILE = 1024;

template <int KMAX>
__global__ void __launch_bounds__(KX_THREADS) knn_expanded_kernel(int n, int m, int k, const float *__restrict__ q,
                                                                  const float *__restrict__ s, int *__restrict__ idx) {
    __shared__ float4 s_pts[KX_TILE];
    const int cloud = blockIdx.y;
    const int i = blockIdx.x * KX_THREADS + threadIdx.x;
    const bool ok = i < n;
    const float *qp = q + ((long long)cloud * n + (ok ? i : 0)) * 3;
    const float qx = __ldg(qp + 0), qy = __ldg(qp + 1), qz = __ldg(qp + 2);
    // torch.sum(src ** 2, -1) on the permuted view: (x*x + y*y) + z*z, every op rounded (no contraction)
    const float q2 = __fadd_rn(__fadd_rn(__fmul_rn(qx, qx), __fmul_rn(qy, qy)), __fmul_rn(qz, qz));
    const float inf = __int_as_float(0x7f800000);
    float best[KMAX];
    int besti[KMAX];
#pragma unroll
    for (int l = 0; l < KMAX; ++l) {
        best[l] = inf;
        besti[l] = 0;
    }
    float worst = inf;
    for (int base = 0; base < m; base += KX_TILE) {
        const int tn = min(KX_TILE, m - base);
        __syncthreads();
        for (int t = threadIdx.x; t < tn; t += KX_THREADS) {
            const float *p = s + ((long long)cloud * m + base + t) * 3;
            const float x = __ldg(p + 0), y = __ldg(p + 1), z = __ldg(p + 2);
            s_pts[t] = make_float4(x, y, z, __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
        }
        __syncthreads();
        for (int t = 0; t < tn; ++t) {
            const float4 p = s_pts[t];
            // matmul with K = 3 (cuBLAS and MKL agree): fma(z,z', fma(y,y', x*x'))
            const float dot = __fmaf_rn(qz, p.z, __fmaf_rn(qy, p.y, __fmul_rn(qx, p.x)));
            float d = __fadd_rn(__fadd_rn(__fmul_rn(-2.0f, dot), q2), p.w);
            d = fmaxf(d, 0.0f);
            if (d < worst) {
                int pos = k - 1;
#pragma unroll
                for (int l = KMAX - 1; l > 0; --l) {
                    if (l < k && best[l - 1] > d) {
                        best[l] = best[l - 1];
                        besti[l] = besti[l - 1];
                        pos = l - 1;
                    }
                }
#pragma unroll
                for (int l = 0; l < KMAX; ++l)
                    if (l == pos) {
                        best[l] = d;
                        besti[l] = base + t;
                    }
#pragma unroll
                for (int l = 0; l < KMAX; ++l)
                    if (l == k - 1) worst = best[l];
            }
        }
    }
    if (ok) {
        int *o = idx + ((long long)cloud * n + i) * k;
#pragma unroll
        for (int l = 0; l < KMAX; ++l)
            if (l < k) o[l] = besti[l];
    }
}

// Warp-cooperative variant (k <= 32): one warp per query, 16 queries per CTA.  The search cloud is staged
// in tiles of 1024 points; lane L holds candidates L, L+32, ... of the tile in registers plus one carried
// winner of the previous tiles.  The k smallest are then extracted one by one with redux.sync.min
// (distance bits first, then index: ties -> lower index).  No divergent per-thread insertion lists.
constexpr int KW_WARPS = 8, KW_QPW = 2, KW_TILE = 1024;

__global__ void __launch_bounds__(32 * KW_WARPS) knn_expanded_warp_kernel(int n, int m_stride, int k, const float *__restrict__ q,
                                                                           const float *__restrict__ s, int *__restrict__ idx,
                                                                           const int *__restrict__ mcounts) {
    __shared__ float4 s_pts[KW_TILE];
    const int cloud = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // mcounts (optional): valid search points per cloud of a padded batch; the cloud stride stays m_stride
    const int m = mcounts ? min(m_stride, __ldg(mcounts + cloud)) : m_stride;
    const int q0 = (blockIdx.x * KW_WARPS + warp) * KW_QPW;
    float qx[KW_QPW], qy[KW_QPW], qz[KW_QPW], q2[KW_QPW];
    uint32_t car_d[KW_QPW], car_i[KW_QPW];   // carried best list: lane l < k holds the l-th smallest so far
#pragma unroll
    for (int u = 0; u < KW_QPW; ++u) {
        const int qi = min(q0 + u, n - 1);
        const float *qp = q + ((long long)cloud * n + qi) * 3;
        qx[u] = __ldg(qp + 0); qy[u] = __ldg(qp + 1); qz[u] = __ldg(qp + 2);
        q2[u] = __fadd_rn(__fadd_rn(__fmul_rn(qx[u], qx[u]), __fmul_rn(qy[u], qy[u])), __fmul_rn(qz[u], qz[u]));
        car_d[u] = 0x7f800000u;
        car_i[u] = 0xffffffffu;
    }
    for (int base = 0; base < m; base += KW_TILE) {
        const int tn = min(KW_TILE, m - base);
        __syncthreads();
        for (int t = threadIdx.x; t < tn; t += 32 * KW_WARPS) {
            const float *p = s + ((long long)cloud * m_stride + base + t) * 3;
            const float x = __ldg(p + 0), y = __ldg(p + 1), z = __ldg(p + 2);
            s_pts[t] = make_float4(x, y, z, __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < KW_QPW; ++u) {
            uint32_t d[32];   // distance bits (d >= 0, so uint order == float order); +inf marks "absent / taken"
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const int t = lane + 32 * i;
                if (t < tn) {
                    const float4 p = s_pts[t];
                    const float dot = __fmaf_rn(qz[u], p.z, __fmaf_rn(qy[u], p.y, __fmul_rn(qx[u], p.x)));
                    d[i] = __float_as_uint(fmaxf(__fadd_rn(__fadd_rn(__fmul_rn(-2.0f, dot), q2[u]), p.w), 0.0f));
                } else {
                    d[i] = 0x7f800000u;
                }
            }
            uint32_t cd = car_d[u], ci = car_i[u], nd = 0x7f800000u, ni = 0xffffffffu;
            // ---- prune: U = k-th smallest of the 32 lane minima.  The k lanes with the smallest minima each own a
            // candidate <= U, so the k nearest are all <= U; typically only ~k..2k candidates survive.
            uint32_t lmin = cd;
#pragma unroll
            for (int i = 0; i < 32; ++i) lmin = min(lmin, d[i]);
            uint32_t sv = lmin;   // bitonic sort of the 32 lane minima across the warp (ascending by lane)
#pragma unroll
            for (int k2 = 2; k2 <= 32; k2 <<= 1)
#pragma unroll
                for (int j = k2 >> 1; j > 0; j >>= 1) {
                    const uint32_t other = __shfl_xor_sync(0xffffffffu, sv, j);
                    const bool asc = (lane & k2) == 0, low = (lane & j) == 0;
                    sv = (asc == low) ? min(sv, other) : max(sv, other);
                }
            const uint32_t U = __shfl_sync(0xffffffffu, sv, k - 1);
            // survivors of this lane, in index order, into 4 register slots
            uint32_t sd0 = 0x7f800000u, sd1 = sd0, sd2 = sd0, sd3 = sd0, si0 = 0, si1 = 0, si2 = 0, si3 = 0;
            int cnt = 0;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                if (d[i] <= U && d[i] != 0x7f800000u) {
                    const uint32_t id = (uint32_t)(base + lane + 32 * i);
                    if (cnt == 0) { sd0 = d[i]; si0 = id; }
                    else if (cnt == 1) { sd1 = d[i]; si1 = id; }
                    else if (cnt == 2) { sd2 = d[i]; si2 = id; }
                    else if (cnt == 3) { sd3 = d[i]; si3 = id; }
                    ++cnt;
                }
            }
            if (__any_sync(0xffffffffu, cnt > 4)) {
                // rare (heavy ties / clustered duplicates): extract from the full register tile
                for (int r = 0; r < k; ++r) {
                    uint32_t ld = cd, li = ci;
                    int lslot = -1;
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (d[i] < ld) { ld = d[i]; li = (uint32_t)(base + lane + 32 * i); lslot = i; }
                    const uint32_t md = rt_redux_min_u32(ld);
                    const uint32_t mi = rt_redux_min_u32(ld == md ? li : 0xffffffffu);
                    if (ld == md && li == mi) {
                        if (lslot < 0) cd = 0x7f800000u;
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (i == lslot) d[i] = 0x7f800000u;
                    }
                    if (lane == r) { nd = md; ni = mi; }
                }
            } else {
                for (int r = 0; r < k; ++r) {
                    // lane-local minimum: carried entry first (older => lower index wins ties), then slots in index order
                    uint32_t ld = cd, li = ci;
                    int lslot = -1;
                    if (sd0 < ld) { ld = sd0; li = si0; lslot = 0; }
                    if (sd1 < ld) { ld = sd1; li = si1; lslot = 1; }
                    if (sd2 < ld) { ld = sd2; li = si2; lslot = 2; }
                    if (sd3 < ld) { ld = sd3; li = si3; lslot = 3; }
                    const uint32_t md = rt_redux_min_u32(ld);
                    const uint32_t mi = rt_redux_min_u32(ld == md ? li : 0xffffffffu);
                    if (ld == md && li == mi) {           // exactly one lane owns the winner: retire it
                        if (lslot < 0) cd = 0x7f800000u;
                        else if (lslot == 0) sd0 = 0x7f800000u;
                        else if (lslot == 1) sd1 = 0x7f800000u;
                        else if (lslot == 2) sd2 = 0x7f800000u;
                        else sd3 = 0x7f800000u;
                    }
                    if (lane == r) { nd = md; ni = mi; }
                }
            }
            car_d[u] = nd;
            car_i[u] = ni;
        }
    }
#pragma unroll
    for (int u = 0; u < KW_QPW; ++u)
        if (q0 + u < n && lane < k) idx[((long long)cloud * n + q0 + u) * k + lane] = (int)car_i[u];
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) nn_weights_kernel(long long rows, float *__restrict__ d) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    float *p = d + r * 3;
    // dist = sqrt(d2); recip = 1/(dist + 1e-8); w = recip / (r0 + r1 + r2)   (lib/pointnet2_modules.py:141-144)
    const float r0 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(p[0]), 1e-8f));
    const float r1 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(p[1]), 1e-8f));
    const float r2 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(p[2]), 1e-8f));
    const float norm = __fadd_rn(__fadd_rn(r0, r1), r2);
    p[0] = __fdiv_rn(r0, norm);
    p[1] = __fdiv_rn(r1, norm);
    p[2] = __fdiv_rn(r2, norm);
}

__global__ void __launch_bounds__(256) interp3_kernel(int clouds, int n, int m, int c, const float *__restrict__ f, int ldf,
                                                      const int *__restrict__ idx, const float *__restrict__ w,
                                                      float *__restrict__ out, int ldo) {
    const long long total = (long long)clouds * n * c;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int ch = (int)(e % c);
        const long long row = e / c;
        const int cloud = (int)(row / n);
        const int *ix = idx + row * 3;
        const float *ww = w + row * 3;
        const float *base = f + (long long)cloud * m * ldf + ch;
        float t = __fmul_rn(__ldg(ww + 1), __ldg(base + (long long)__ldg(ix + 1) * ldf));
        t = __fmaf_rn(__ldg(ww + 0), __ldg(base + (long long)__ldg(ix + 0) * ldf), t);
        out[row * ldo + ch] = __fmaf_rn(__ldg(ww + 2), __ldg(base + (long long)__ldg(ix + 2) * ldf), t);
    }
}

__global__ void __launch_bounds__(256) gather_rows3_kernel(int clouds, int npts_out, int n_in, int c,
                                                           const float *__restrict__ src, const int *__restrict__ idx,
                                                           float *__restrict__ dst) {
    const long long total = (long long)clouds * npts_out * c;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int ch = (int)(e % c);
        const long long row = e / c;
        const int cloud = (int)(row / npts_out);
        dst[e] = __ldg(src + ((long long)cloud * n_in + __ldg(idx + row)) * c + ch);
    }
}

__global__ void __launch_bounds__(256) cloud_max_kernel(int npts, int c, const float *__restrict__ f, int ldf, float *__restrict__ g,
                                                        const int *__restrict__ counts) {
    // grid (cloud, channel-block of 32); 256 threads = 32 point-lanes x 8 channel quads (128-bit loads), four independent
    // loads in flight per thread (the first version walked its points with one dependent 32-bit load at a time: 22 us for
    // 33 MB).  max is exact in any order.
    // counts (optional): valid points per cloud of a padded batch (rows beyond are ignored; the row stride stays npts)
    __shared__ float4 s[32][9];
    const int cloud = blockIdx.x, cq = threadIdx.x & 7, pl = threadIdx.x >> 3;
    const int ch = blockIdx.y * 32 + 4 * cq;
    const int np = counts ? min(npts, __ldg(counts + cloud)) : npts;
    const float ninf = -__int_as_float(0x7f800000);
    float4 m = make_float4(ninf, ninf, ninf, ninf);
    const float *base = f + (long long)cloud * npts * ldf + ch;
    if (ch + 3 < c && (ldf & 3) == 0 && (reinterpret_cast<uintptr_t>(base) & 15) == 0) {
        int p = pl;
        for (; p + 96 < np; p += 128) {
            const float4 a0 = __ldg(reinterpret_cast<const float4 *>(base + (long long)p * ldf));
            const float4 a1 = __ldg(reinterpret_cast<const float4 *>(base + (long long)(p + 32) * ldf));
            const float4 a2 = __ldg(reinterpret_cast<const float4 *>(base + (long long)(p + 64) * ldf));
            const float4 a3 = __ldg(reinterpret_cast<const float4 *>(base + (long long)(p + 96) * ldf));
            m.x = fmaxf(fmaxf(m.x, a0.x), fmaxf(fmaxf(a1.x, a2.x), a3.x));
            m.y = fmaxf(fmaxf(m.y, a0.y), fmaxf(fmaxf(a1.y, a2.y), a3.y));
            m.z = fmaxf(fmaxf(m.z, a0.z), fmaxf(fmaxf(a1.z, a2.z), a3.z));
            m.w = fmaxf(fmaxf(m.w, a0.w), fmaxf(fmaxf(a1.w, a2.w), a3.w));
        }
        for (; p < np; p += 32) {
            const float4 a0 = __ldg(reinterpret_cast<const float4 *>(base + (long long)p * ldf));
            m.x = fmaxf(m.x, a0.x); m.y = fmaxf(m.y, a0.y); m.z = fmaxf(m.z, a0.z); m.w = fmaxf(m.w, a0.w);
        }
    } else {
        for (int p = pl; p < np; p += 32) {
            if (ch + 0 < c) m.x = fmaxf(m.x, __ldg(base + (long long)p * ldf + 0));
            if (ch + 1 < c) m.y = fmaxf(m.y, __ldg(base + (long long)p * ldf + 1));
            if (ch + 2 < c) m.z = fmaxf(m.z, __ldg(base + (long long)p * ldf + 2));
            if (ch + 3 < c) m.w = fmaxf(m.w, __ldg(base + (long long)p * ldf + 3));
        }
    }
    s[pl][cq] = m;
    __syncthreads();
    if (pl == 0) {
#pragma unroll 8
        for (int i = 1; i < 32; ++i) {
            const float4 o = s[i][cq];
            m.x = fmaxf(m.x, o.x); m.y = fmaxf(m.y, o.y); m.z = fmaxf(m.z, o.z); m.w = fmaxf(m.w, o.w);
        }
        float *o = g + (long long)cloud * c + ch;
        if (ch + 0 < c) o[0] = m.x;
        if (ch + 1 < c) o[1] = m.y;
        if (ch + 2 < c) o[2] = m.z;
        if (ch + 3 < c) o[3] = m.w;
    }
}

__global__ void __launch_bounds__(256) cloud_matvec_kernel(int nout, int k, const float *__restrict__ w, int ldw,
                                                           const float *__restrict__ g, int ldg, const float *__restrict__ bias,
                                                           float *__restrict__ cb) {
    // grid (cloud, output block of 8); one warp per output, lanes stride over k
    const int cloud = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int o = blockIdx.y * 8 + warp; o < nout && o < blockIdx.y * 8 + 8; o += 8) {
        float acc = 0.0f;
        for (int kk = lane; kk < k; kk += 32) acc = fmaf(__ldg(w + (long long)o * ldw + kk), __ldg(g + (long long)cloud * ldg + kk), acc);
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
        if (lane == 0) cb[(long long)cloud * nout + o] = acc + (bias ? __ldg(bias + o) : 0.0f);
    }
}

// (b, c, n) <-> (b*n, ld) transposes through a 32x33 tile
__global__ void __launch_bounds__(256) cm_to_rows_kernel(int c, int n, const float *__restrict__ src, float *__restrict__ dst, int ldd, int coff) {
    __shared__ float t[32][33];
    const int b = blockIdx.z, n0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int i = ty; i < 32; i += 8)
        if (c0 + i < c && n0 + tx < n) t[i][tx] = __ldg(src + ((long long)b * c + c0 + i) * n + n0 + tx);
    __syncthreads();
    for (int i = ty; i < 32; i += 8)
        if (n0 + i < n && c0 + tx < c) dst[((long long)b * n + n0 + i) * ldd + coff + c0 + tx] = t[tx][i];
}
__global__ void __launch_bounds__(256) rows_to_cm_kernel(int c, int n, const float *__restrict__ src, int lds, int soff,
                                                         float *__restrict__ dst, int dst_c, int dst_coff) {
    __shared__ float t[32][33];
    const int b = blockIdx.z, n0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int i = ty; i < 32; i += 8)
        if (n0 + i < n && c0 + tx < c) t[i][tx] = __ldg(src + ((long long)b * n + n0 + i) * lds + soff + c0 + tx);
    __syncthreads();
    for (int i = ty; i < 32; i += 8)
        if (c0 + i < c && n0 + tx < n) dst[((long long)b * dst_c + dst_coff + c0 + i) * n + n0 + tx] = t[tx][i];
}
__global__ void __launch_bounds__(256) broadcast_cm_kernel(int c, int n, const float *__restrict__ g, float *__restrict__ dst,
                                                           int dst_c, int dst_coff) {
    const int b = blockIdx.z, ch = blockIdx.y;
    const float v = __ldg(g + (long long)b * c + ch);
    float *row = dst + ((long long)b * dst_c + dst_coff + ch) * n;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) row[i] = v;
}

// Weights of the GRU kernels below arrive TRANSPOSED (128 x 384) so consecutive threads read consecutive rows of one k.
// hidden half of all five GRU layers: gh[l][b][row] = W_hh[l][row,:] . h_in[l][b] + b_hh[l][row].  grid (3, B, 5) x 128.
__global__ void __launch_bounds__(128) gru_hh_kernel(int bsz, const float *__restrict__ h_in, const float *__restrict__ whh_t,
                                                     const float *__restrict__ bhh, size_t h_stride, float *__restrict__ gh) {
    __shared__ float s_h[128];
    const int l = blockIdx.z, b = blockIdx.y, row = blockIdx.x * 128 + threadIdx.x;
    s_h[threadIdx.x] = h_in[(size_t)l * h_stride + (size_t)b * 128 + threadIdx.x];
    __syncthreads();
    const float *w = whh_t + (size_t)l * 128 * 384 + row;
    float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
#pragma unroll 8
    for (int k = 0; k < 128; k += 4) {
        a0 = fmaf(__ldg(w + (k + 0) * 384), s_h[k + 0], a0);
        a1 = fmaf(__ldg(w + (k + 1) * 384), s_h[k + 1], a1);
        a2 = fmaf(__ldg(w + (k + 2) * 384), s_h[k + 2], a2);
        a3 = fmaf(__ldg(w + (k + 3) * 384), s_h[k + 3], a3);
    }
    gh[((size_t)l * bsz + b) * 384 + row] = (a0 + a1) + (a2 + a3) + __ldg(bhh + l * 384 + row);
}

// input half + gates of all five layers of one batch element: 384 threads = the r, z, n rows of W_ih.
__global__ void __launch_bounds__(384) gru_chain_kernel(int bsz, const float *__restrict__ x, const float *__restrict__ h_in,
                                                        const float *__restrict__ wih_t, const float *__restrict__ bih,
                                                        const float *__restrict__ gh, float *__restrict__ h_out, size_t h_stride) {
    __shared__ float s_x[128], s_gi[384];
    const int b = blockIdx.x, t = threadIdx.x;
    if (t < 128) s_x[t] = x[(size_t)b * 128 + t];
    __syncthreads();
    for (int l = 0; l < 5; ++l) {
        const float *w = wih_t + (size_t)l * 128 * 384 + t;
        float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
        // the chain is bound by the latency of streaming 196 KB of weights per layer from L2: keep 32 loads in flight
#pragma unroll 1
        for (int k0 = 0; k0 < 128; k0 += 32) {
            float wv[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) wv[i] = __ldg(w + (k0 + i) * 384);
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                a0 = fmaf(wv[i + 0], s_x[k0 + i + 0], a0);
                a1 = fmaf(wv[i + 1], s_x[k0 + i + 1], a1);
                a2 = fmaf(wv[i + 2], s_x[k0 + i + 2], a2);
                a3 = fmaf(wv[i + 3], s_x[k0 + i + 3], a3);
            }
        }
        s_gi[t] = (a0 + a1) + (a2 + a3) + __ldg(bih + l * 384 + t);
        __syncthreads();
        if (t < 128) {
            const float *g = gh + ((size_t)l * bsz + b) * 384;
            const float r = 1.0f / (1.0f + expf(-(s_gi[t] + g[t])));
            const float z = 1.0f / (1.0f + expf(-(s_gi[128 + t] + g[128 + t])));
            const float nn = tanhf(s_gi[256 + t] + r * g[256 + t]);
            const float h = (1.0f - z) * nn + z * h_in[(size_t)l * h_stride + (size_t)b * 128 + t];
            h_out[(size_t)l * h_stride + (size_t)b * 128 + t] = h;
            s_x[t] = h;
        }
        __syncthreads();
    }
}

// Cluster variant of the chain: 4 CTAs (one thread-block cluster) per batch element.  CTA r owns hidden units
// [32r, 32r+32) = 96 of the 384 gate rows, 4 threads per row (a quarter of k each, all 32 weight loads in flight at once);
// the 32 new hidden values of a CTA are written into the next-layer input buffer of ALL four CTAs through distributed
// shared memory, one cluster barrier per layer.  The chain is bound by streaming W_ih from L2 (196 KB per layer): four SMs
// pull it in parallel, and every load of a layer is in flight together -- 43 -> ~10 us for the five layers.
__device__ __forceinline__ uint32_t gru_cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void gru_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void gru_st_remote(float *local_addr, uint32_t rank, float v) {
    uint32_t la = rt_smem_u32(local_addr), ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(rank));
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(ra), "f"(v) : "memory");
}
__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(384) gru_chain_cluster_kernel(
    int bsz, const float *__restrict__ x, const float *__restrict__ h_in, const float *__restrict__ wih_t,
    const float *__restrict__ bih, const float *__restrict__ gh, float *__restrict__ h_out, size_t h_stride) {
    __shared__ float s_x[2][128], s_part[4][96];
    const int t = threadIdx.x;
    const uint32_t rank = gru_cluster_rank();
    const int b = blockIdx.x / 4;
    const int rl = t % 96, kq = t / 96;                     // local row (gate-major: 32 rows per gate), k quarter
    const int row = (rl / 32) * 128 + 32 * (int)rank + (rl % 32);
    if (t < 128) s_x[0][t] = x[(size_t)b * 128 + t];
    gru_cluster_sync();                                     // every CTA's buffers exist before anybody writes remotely
    int cur = 0;
    for (int l = 0; l < 5; ++l) {
        const float *w = wih_t + (size_t)l * 128 * 384 + (size_t)(kq * 32) * 384 + row;
        float wv[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) wv[i] = __ldg(w + (size_t)i * 384);
        float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
            a0 = fmaf(wv[i + 0], s_x[cur][kq * 32 + i + 0], a0);
            a1 = fmaf(wv[i + 1], s_x[cur][kq * 32 + i + 1], a1);
            a2 = fmaf(wv[i + 2], s_x[cur][kq * 32 + i + 2], a2);
            a3 = fmaf(wv[i + 3], s_x[cur][kq * 32 + i + 3], a3);
        }
        s_part[kq][rl] = (a0 + a1) + (a2 + a3);
        __syncthreads();
        if (t < 32) {
            const int u = 32 * (int)rank + t;               // hidden unit
            float gi[3];
#pragma unroll
            for (int g = 0; g < 3; ++g)
                gi[g] = ((s_part[0][g * 32 + t] + s_part[1][g * 32 + t]) + (s_part[2][g * 32 + t] + s_part[3][g * 32 + t])) +
                        __ldg(bih + l * 384 + g * 128 + u);
            const float *g_h = gh + ((size_t)l * bsz + b) * 384;
            const float r = 1.0f / (1.0f + expf(-(gi[0] + g_h[u])));
            const float z = 1.0f / (1.0f + expf(-(gi[1] + g_h[128 + u])));
            const float nn = tanhf(gi[2] + r * g_h[256 + u]);
            const float h = (1.0f - z) * nn + z * h_in[(size_t)l * h_stride + (size_t)b * 128 + u];
            h_out[(size_t)l * h_stride + (size_t)b * 128 + u] = h;
#pragma unroll
            for (uint32_t c = 0; c < 4; ++c) gru_st_remote(&s_x[cur ^ 1][u], c, h);   // next layer's input, in all four CTAs
        }
        gru_cluster_sync();                                 // remote stores visible; s_part / s_x[cur] free again
        cur ^= 1;
    }
}

__global__ void __launch_bounds__(256) cls_tail_kernel(long long rows, const float *__restrict__ h3, const float *__restrict__ w4,
                                                       const float *__restrict__ lin_w, const float *__restrict__ lin_b,
                                                       float *__restrict__ cls) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const float *h = h3 + r * 32;
    float o[3] = {0.0f, 0.0f, 0.0f};
#pragma unroll 8
    for (int k = 0; k < 32; ++k) {
        const float v = __ldg(h + k);
        o[0] = fmaf(__ldg(w4 + k), v, o[0]);
        o[1] = fmaf(__ldg(w4 + 32 + k), v, o[1]);
        o[2] = fmaf(__ldg(w4 + 64 + k), v, o[2]);
    }
    const float y = fmaf(__ldg(lin_w + 2), o[2], fmaf(__ldg(lin_w + 1), o[1], __ldg(lin_w + 0) * o[0])) + __ldg(lin_b);
    cls[r] = 1.0f / (1.0f + expf(-y));
}

// Morton-order permutation of one cloud per CTA: bounding box -> 10 bits per axis -> interleave -> bitonic sort of
// (code << 32 | index) in shared memory.  Any xyz (NaN included) still yields a permutation: the index rides in the key.
constexpr int MP_MAX = 4096;
__device__ __forceinline__ uint32_t mp_spread3(uint32_t v) {   // 10 bits -> every third bit
    v &= 0x3ffu;
    v = (v | (v << 16)) & 0x030000ffu;
    v = (v | (v << 8)) & 0x0300f00fu;
    v = (v | (v << 4)) & 0x030c30c3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}
__global__ void __launch_bounds__(256) morton_perm_kernel(int n, int npad, const float *__restrict__ xyz_all, int *__restrict__ perm) {
    extern __shared__ unsigned long long s_key[];   // npad
    __shared__ float s_lo[3][8], s_hi[3][8];
    const int cloud = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const float *xyz = xyz_all + (size_t)cloud * n * 3;
    float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    for (int i = t; i < n; i += 256)
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float v = __ldg(xyz + i * 3 + a);
            lo[a] = fminf(lo[a], v);
            hi[a] = fmaxf(hi[a], v);
        }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
        }
        if (lane == 0) { s_lo[a][warp] = lo[a]; s_hi[a][warp] = hi[a]; }
    }
    __syncthreads();
    float scale[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float l = s_lo[a][0], h = s_hi[a][0];
#pragma unroll
        for (int w = 1; w < 8; ++w) { l = fminf(l, s_lo[a][w]); h = fmaxf(h, s_hi[a][w]); }
        lo[a] = l;
        scale[a] = (h > l) ? 1023.0f / (h - l) : 0.0f;
    }
    for (int i = t; i < npad; i += 256) {
        unsigned long long key = ~0ull;
        if (i < n) {
            uint32_t q[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const float f = (__ldg(xyz + i * 3 + a) - lo[a]) * scale[a];
                q[a] = (uint32_t)fminf(fmaxf(f, 0.0f), 1023.0f);   // NaN -> 0 (fmaxf drops it)
            }
            const uint32_t code = mp_spread3(q[0]) | (mp_spread3(q[1]) << 1) | (mp_spread3(q[2]) << 2);
            key = ((unsigned long long)code << 32) | (uint32_t)i;
        }
        s_key[i] = key;
    }
    __syncthreads();
    for (int k = 2; k <= npad; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = t; i < npad; i += 256) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long x = s_key[i], y = s_key[ixj];
                    const bool up = (i & k) == 0;
                    if ((x > y) == up) { s_key[i] = y; s_key[ixj] = x; }
                }
            }
            __syncthreads();
        }
    for (int i = t; i < n; i += 256) perm[(size_t)cloud * n + i] = cloud * n + (int)(uint32_t)(s_key[i] & 0xffffffffull);
}
__global__ void __launch_bounds__(256) iota_kernel(long long n, int *__restrict__ p) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = (int)i;
}

__global__ void __launch_bounds__(256) fill_kernel(float *p, long long n, float v) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}

inline int grid_for(long long total, int threads, int cap = 148 * 16) {
    long long g = (total + threads - 1) / threads;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

int rt_launch_rowgemm(const RtRowGemm &a, cudaStream_t st) {
    if (a.rows <= 0 || a.nout <= 0) return RT_OK;
    dim3 grid(rt_divup(a.rows, RG_T), rt_divup(a.nout, RG_T));
    rowgemm_kernel<<<grid, 256, 0, st>>>(a);
    return rt_check_launch("rowgemm_kernel");
}
int rt_launch_gather_combine(const RtGatherCombine &a, cudaStream_t st) {
    const long long total = (long long)a.clouds * a.npts * a.ns * a.c;
    if (total <= 0) return RT_OK;
    gather_combine_kernel<<<grid_for(total, 256, 148 * 32), 256, 0, st>>>(a);
    return rt_check_launch("gather_combine_kernel");
}
int rt_launch_maxpool_rows(int groups, int ns, int c, const float *x, float *pooled, int ldp, int coff, cudaStream_t st) {
    if (groups <= 0) return RT_OK;
    maxpool_rows_kernel<<<grid_for((long long)groups * c, 256), 256, 0, st>>>(groups, ns, c, x, pooled, ldp, coff);
    return rt_check_launch("maxpool_rows_kernel");
}
int rt_launch_weighted_sum(const RtWeightedSum &a, cudaStream_t st) {
    const long long ncp = (long long)a.clouds * a.npts;
    if (ncp <= 0) return RT_OK;
    RT_REQUIRE(a.ns <= 32, "weighted_sum: nsample > 32");
    const size_t smem = WS_PTS * a.ns * 16 * sizeof(float);
    if ((a.c & 3) == 0 && a.c >= 4 && a.c <= 256) {
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const long long ngroups = (ncp + WS_PTS - 1) / WS_PTS;
        const long long grid = ngroups < 2ll * sms ? ngroups : 2ll * sms;   // 2 resident CTAs per SM (126 registers x 256 threads)
        weighted_sum_kernel<<<(int)grid, 256, smem, st>>>(a);
        return rt_check_launch("weighted_sum_kernel");
    }
    weighted_sum_generic_kernel<<<rt_divup(ncp, WS_PTS), 256, smem, st>>>(a);
    return rt_check_launch("weighted_sum_generic_kernel");
}
int rt_launch_knn_expanded(int clouds, int n, int m, int k, const float *q, const float *s, int *idx, cudaStream_t st,
                           const int *mcounts) {
    if (clouds <= 0 || n <= 0) return RT_OK;
    RT_REQUIRE(k >= 1 && k <= 32 && m >= 1, "knn_expanded: k=%d outside [1,32] or empty search cloud", k);
    RT_REQUIRE(k <= m, "knn_expanded: k=%d > search points %d", k, m);
    dim3 grid(rt_divup(n, KW_WARPS * KW_QPW), clouds);
    knn_expanded_warp_kernel<<<grid, 32 * KW_WARPS, 0, st>>>(n, m, k, q, s, idx, mcounts);
    return rt_check_launch("knn_expanded_warp_kernel");
}
int rt_launch_nn_weights(long long rows, float *d, cudaStream_t st) {
    if (rows <= 0) return RT_OK;
    nn_weights_kernel<<<rt_divup(rows, 256), 256, 0, st>>>(rows, d);
    return rt_check_launch("nn_weights_kernel");
}
int rt_launch_interp3(int clouds, int n, int m, int c, const float *f, int ldf, const int *idx, const float *w, float *out,
                      int ldo, cudaStream_t st) {
    const long long total = (long long)clouds * n * c;
    if (total <= 0) return RT_OK;
    interp3_kernel<<<grid_for(total, 256, 148 * 32), 256, 0, st>>>(clouds, n, m, c, f, ldf, idx, w, out, ldo);
    return rt_check_launch("interp3_kernel");
}
int rt_launch_gather_rows(int clouds, int npts_out, int n_in, int c, const float *src, const int *idx, float *dst, cudaStream_t st) {
    const long long total = (long long)clouds * npts_out * c;
    if (total <= 0) return RT_OK;
    gather_rows3_kernel<<<grid_for(total, 256), 256, 0, st>>>(clouds, npts_out, n_in, c, src, idx, dst);
    return rt_check_launch("gather_rows3_kernel");
}
int rt_launch_cloud_max(int clouds, int npts, int c, const float *f, int ldf, float *g, cudaStream_t st, const int *counts) {
    if (clouds <= 0) return RT_OK;
    dim3 grid(clouds, rt_divup(c, 32));
    cloud_max_kernel<<<grid, 256, 0, st>>>(npts, c, f, ldf, g, counts);
    return rt_check_launch("cloud_max_kernel");
}
int rt_launch_cloud_matvec(int clouds, int nout, int k, const float *w, int ldw, const float *g, int ldg, const float *bias,
                           float *cb, cudaStream_t st) {
    if (clouds <= 0) return RT_OK;
    cloud_matvec_kernel<<<dim3(clouds, rt_divup(nout, 8)), 256, 0, st>>>(nout, k, w, ldw, g, ldg, bias, cb);
    return rt_check_launch("cloud_matvec_kernel");
}
int rt_launch_cm_to_rows(int b, int c, int n, const float *src, float *dst, int ldd, int coff, cudaStream_t st) {
    if (b <= 0) return RT_OK;
    dim3 grid(rt_divup(n, 32), rt_divup(c, 32), b);
    cm_to_rows_kernel<<<grid, 256, 0, st>>>(c, n, src, dst, ldd, coff);
    return rt_check_launch("cm_to_rows_kernel");
}
int rt_launch_rows_to_cm(int b, int c, int n, const float *src, int lds, int soff, float *dst, int dst_c, int dst_coff,
                         cudaStream_t st) {
    if (b <= 0) return RT_OK;
    dim3 grid(rt_divup(n, 32), rt_divup(c, 32), b);
    rows_to_cm_kernel<<<grid, 256, 0, st>>>(c, n, src, lds, soff, dst, dst_c, dst_coff);
    return rt_check_launch("rows_to_cm_kernel");
}
int rt_launch_broadcast_cm(int b, int c, int n, const float *g, float *dst, int dst_c, int dst_coff, cudaStream_t st) {
    if (b <= 0) return RT_OK;
    dim3 grid(rt_divup(n, 256), c, b);
    broadcast_cm_kernel<<<grid, 256, 0, st>>>(c, n, g, dst, dst_c, dst_coff);
    return rt_check_launch("broadcast_cm_kernel");
}
// 5-layer GRU, sequence length 1 (torch.nn.GRU equations; reference: FlowDecoder.torchGRU, model_utils.py:279,294-297).
// Only the input half of every layer (W_ih . x_l, x_l = h_out[l-1]) is a dependent chain; the hidden half
// (W_hh . h_in[l] + b_hh) depends on the call's inputs alone, so rt_launch_gru_hh evaluates it for all five layers at
// the start of the step on a side stream, and the chain kernel (one CTA per batch element, all five layers, one
// launch instead of five) only adds the input half.  Summation order of every dot product = gru_layer_kernel's.
int rt_launch_gru_hh(int b, const float *h_in, const float *whh, const float *bhh, size_t h_stride, float *gh, cudaStream_t st) {
    if (b <= 0) return RT_OK;
    gru_hh_kernel<<<dim3(3, b, 5), 128, 0, st>>>(b, h_in, whh, bhh, h_stride, gh);
    return rt_check_launch("gru_hh_kernel");
}
int rt_launch_gru(int b, const float *x, const float *h_in, const float *wih, const float *bih, const float *gh, float *h_out,
                  size_t h_stride, cudaStream_t st) {
    if (b <= 0) return RT_OK;
    static int use_cluster = -1;
    if (use_cluster < 0) {
        const char *env = getenv("RT_GRU_CLUSTER");
        use_cluster = (env && atoi(env) == 0) ? 0 : 1;
    }
    if (use_cluster) {
        gru_chain_cluster_kernel<<<4 * b, 384, 0, st>>>(b, x, h_in, wih, bih, gh, h_out, h_stride);
        return rt_check_launch("gru_chain_cluster_kernel");
    }
    gru_chain_kernel<<<b, 384, 0, st>>>(b, x, h_in, wih, bih, gh, h_out, h_stride);
    return rt_check_launch("gru_chain_kernel");
}
int rt_launch_cls_tail(long long rows, const float *h3, const float *w4, const float *lin_w, const float *lin_b, float *cls,
                       cudaStream_t st) {
    if (rows <= 0) return RT_OK;
    cls_tail_kernel<<<rt_divup(rows, 256), 256, 0, st>>>(rows, h3, w4, lin_w, lin_b, cls);
    return rt_check_launch("cls_tail_kernel");
}
int rt_launch_morton_perm(int clouds, int n, const float *xyz, int *perm, cudaStream_t st) {
    if (clouds <= 0 || n <= 0) return RT_OK;
    RT_REQUIRE((long long)clouds * n < (1ll << 31), "morton_perm: too many points");
    if (n > MP_MAX) {   // larger clouds keep the input order (still correct, just no locality gain)
        const long long total = (long long)clouds * n;
        iota_kernel<<<grid_for(total, 256), 256, 0, st>>>(total, perm);
        return rt_check_launch("iota_kernel");
    }
    int npad = 32;
    while (npad < n) npad *= 2;
    morton_perm_kernel<<<clouds, 256, npad * sizeof(unsigned long long), st>>>(n, npad, xyz, perm);
    return rt_check_launch("morton_perm_kernel");
}
// Loud failure instead of silent garbage: when the fp16-range guard of the tensor-core kernels fired (status != 0), the
// step's results are overwritten with NaN on the device, so a caller that never reads the status word cannot consume
// saturated values.  No host synchronisation; when the status is clean the kernel reads one word and exits.
__global__ void __launch_bounds__(256) poison_on_status_kernel(const int *status, float *flow, long long nflow, float *cls,
                                                               long long ncls, float *h_out, int b, size_t h_stride) {
    if (*status == 0) return;
    const float nan = __int_as_float(0x7fc00000);
    const long long stride = (long long)gridDim.x * blockDim.x, t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (long long i = t0; i < nflow; i += stride) flow[i] = nan;
    for (long long i = t0; i < ncls; i += stride) cls[i] = nan;
    for (long long i = t0; i < (long long)5 * b * 128; i += stride) h_out[(i / (b * 128)) * h_stride + i % (b * 128)] = nan;
}
int rt_launch_poison_on_status(const int *status, float *flow, long long nflow, float *cls, long long ncls, float *h_out, int b,
                               size_t h_stride, cudaStream_t st) {
    poison_on_status_kernel<<<64, 256, 0, st>>>(status, flow, nflow, cls, ncls, h_out, b, h_stride);
    return rt_check_launch("poison_on_status_kernel");
}
// dst (b, c, n) channel-major: columns at and beyond counts[b] are set to zero (padded points of a variable-size batch)
__global__ void __launch_bounds__(256) mask_cm_kernel(int c, int n, const int *__restrict__ counts, float *__restrict__ dst) {
    const int b = blockIdx.z, ch = blockIdx.y, n0 = __ldg(counts + b);
    for (int p = n0 + blockIdx.x * 256 + threadIdx.x; p < n; p += gridDim.x * 256) dst[((long long)b * c + ch) * n + p] = 0.0f;
}
int rt_launch_mask_cm(int b, int c, int n, const int *counts, float *dst, cudaStream_t st) {
    if (b <= 0 || c <= 0 || n <= 0) return RT_OK;
    mask_cm_kernel<<<dim3(2, c, b), 256, 0, st>>>(c, n, counts, dst);
    return rt_check_launch("mask_cm_kernel");
}
int rt_launch_fill(float *p, long long n, float v, cudaStream_t st) {
    if (n <= 0) return RT_OK;
    fill_kernel<<<grid_for(n, 256), 256, 0, st>>>(p, n, v);
    return rt_check_launch("fill_kernel");
}

// C ABI: cost-volume neighbour search.  replaces square_distance + knn_point
// (reference: src/utils/model_utils/model_utils.py:17-39, 85-99): expanded-form fp32 distances in torch's own
// rounding order, clamp at 0, the k smallest per query (ascending, ties -> lower index; torch.topk leaves
// both the order and the tie choice undefined).
RT_API int rt_knn_expanded(int b, int n, int m, int k, const float *query, const float *search, int *idx, void *stream) {
    RT_REQUIRE(b >= 0 && n >= 0 && m >= 0 && query && search && idx, "knn_expanded: bad arguments");
    RT_REQUIRE(b <= 65535, "knn_expanded: batch > 65535");
    RT_REQUIRE(k <= m, "knn_expanded: k=%d > number of search points %d", k, m);
    return rt_launch_knn_expanded(b, n, m, k, query, search, idx, (cudaStream_t)stream);
}
