// Cost volume of the TRAINING path, channels innermost (rows), without the reference's (B,515,16,N) / (B,256,16,N)
// intermediates (reference: FeatureCorrelator.forward, src/utils/model_utils/model_utils.py:193-250).
//
// The reference (and round 1 of this package) evaluates the chain with one torch op per step on channel-major tensors:
// gather, add, add, add, LeakyReLU, layout copy, conv, LeakyReLU, conv, LeakyReLU, WeightNet's 8 -> 256 conv, ReLU, multiply,
// sum -- about thirty passes over 4.3 GB tensors at batch 256, and as many again in backward (torch.profiler: a third of the
// training step).  Here the chain is four fused ops, every thread owning ONE CHANNEL so that all global accesses are
// contiguous 1 KB rows:
//   cv1 (forward / backward)   x1[p,k,:] = LeakyReLU(P2[nbr(p,k),:] + P1[p,:] + Wd.(xyz2[nbr] - xyz1[p]) + b)
//                              P1 / P2 = the per-POINT projections of the first convolution (dense_tc.cu GEMMs); backward:
//                              dP1 = sum over k, dWd / db from per-CTA partials added in a fixed order, dP2 as a segmented sum
//                              over the stable inverse index (segsum.cu) -- no atomics, bit-repeatable;
//   act_grad                   g = dy * act'(y), max |g| and the column sums of g (the bias gradient) in ONE pass: the
//                              backward prologue of a dense layer whose activation was fused into the GEMM epilogue;
//   wsum (forward / backward)  out[p,:] = sum_k ReLU(W3.h2[p,k,:] + b3) * x[p,k,:]   (x rows, or rows gathered from per-point
//                              features for the patch-to-patch step): WeightNet's last layer evaluated in registers from its
//                              8-channel hidden rows, the (B,256,16,N) weight tensor never exists; backward recomputes it.
#include "common.cuh"

bool rt_segsum_supported(int n_dst, long long e_total);   // segsum.cu
int rt_launch_inverse_index(int b, int n_dst, long long e_total, const int *idx, int *order, int *seg, cudaStream_t st, const char *what);

namespace {

constexpr int CV_MAXK = 32;     // neighbours per point (the model uses 16)
constexpr int CV_H = 8;         // WeightNet hidden width (reference: model_utils.py:359-390, hidden_unit=[8, 8])

// point / destination index -> cloud: a 32-bit division whenever the index fits (the 64-bit one costs ~100 instructions per thread)
__device__ __forceinline__ long long cv_div(long long a, int b) {
    return a <= 0xffffffffll ? (long long)((unsigned)a / (unsigned)b) : a / b;
}
__device__ __forceinline__ float cv_leaky(float v) { return fmaxf(v, 0.1f * v); }
__device__ __forceinline__ float cv_slope(float y) { return y > 0.0f ? 1.0f : 0.1f; }

int cv_sms() {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms;
}

// ---- cv1 -------------------------------------------------------------------------------------------------------------
// KT = 16: the model's neighbour count as a compile-time constant (loops fully unrolled: all of a point's row loads are in flight
// together -- the runtime-K loops kept ~4 loads per thread in flight and ran at a third of the HBM rate); KT = 0: any K
template <int KT>
__global__ void cv1_fwd_kernel(long long total_pts, int n1, int n2, int Krt, int C, const float *__restrict__ p1, const float *__restrict__ p2,
                               const float *__restrict__ xyz1, const float *__restrict__ xyz2, const int *__restrict__ idx,
                               const float *__restrict__ wd, const float *__restrict__ bias, float *__restrict__ out,
                               float *__restrict__ dir_out) {
    const int c = threadIdx.x;
    const int K = KT ? KT : Krt;
    const float wx = __ldg(wd + c * 3), wy = __ldg(wd + c * 3 + 1), wz = __ldg(wd + c * 3 + 2), bv = bias ? __ldg(bias + c) : 0.0f;
    for (long long p = blockIdx.x; p < total_pts; p += gridDim.x) {
        const long long cloud = cv_div(p, n1);
        const float qx = __ldg(xyz1 + p * 3), qy = __ldg(xyz1 + p * 3 + 1), qz = __ldg(xyz1 + p * 3 + 2);
        const float p1v = __ldg(p1 + p * C + c);
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const long long g = cloud * n2 + __ldg(idx + p * K + k);
            const float dx = __ldg(xyz2 + g * 3) - qx, dy = __ldg(xyz2 + g * 3 + 1) - qy, dz = __ldg(xyz2 + g * 3 + 2) - qz;
            const float v = ((__ldg(p2 + g * C + c) + p1v) + fmaf(wz, dz, fmaf(wy, dy, wx * dx))) + bv;
            out[(p * K + k) * C + c] = cv_leaky(v);
            if (c == 0 && dir_out) {
                float *d = dir_out + (p * K + k) * 3;
                d[0] = dx; d[1] = dy; d[2] = dz;
            }
        }
    }
}

// dP1[p,c] = sum_k g; per-CTA partials of dWd (3 per channel) and db, g = dout * slope(out)
__global__ void cv1_bwd_point_kernel(long long total_pts, int K, int C, const float *__restrict__ dout, const float *__restrict__ out,
                                     const float *__restrict__ dir, float *__restrict__ dp1, float *__restrict__ part) {
    const int c = threadIdx.x;
    float ax = 0.0f, ay = 0.0f, az = 0.0f, ab = 0.0f;
    for (long long p = blockIdx.x; p < total_pts; p += gridDim.x) {
        float s = 0.0f;
        for (int k = 0; k < K; ++k) {
            const long long r = p * K + k;
            const float g = __ldg(dout + r * C + c) * cv_slope(__ldg(out + r * C + c));
            s += g;
            ax = fmaf(g, __ldg(dir + r * 3), ax);
            ay = fmaf(g, __ldg(dir + r * 3 + 1), ay);
            az = fmaf(g, __ldg(dir + r * 3 + 2), az);
        }
        dp1[p * C + c] = s;
        ab += s;
    }
    float *pp = part + (long long)blockIdx.x * 4 * C;
    pp[c] = ax; pp[C + c] = ay; pp[2 * C + c] = az; pp[3 * C + c] = ab;
}

// dP2[cloud, t, c] = sum over the sources e = (point, k) of destination t, in increasing e, of g[cloud, e, c]
__global__ void cv1_bwd_scatter_kernel(long long total_dst, int n2, long long e_total, int C, const float *__restrict__ dout,
                                       const float *__restrict__ out, const int *__restrict__ order, const int *__restrict__ seg,
                                       float *__restrict__ dp2) {
    const int c = threadIdx.x;
    for (long long d = blockIdx.x; d < total_dst; d += gridDim.x) {
        const long long cloud = cv_div(d, n2);
        const int t = (int)(d - cloud * n2);
        const int *sg = seg + cloud * (n2 + 1);
        const int j0 = __ldg(sg + t), j1 = __ldg(sg + t + 1);
        const int *ord = order + cloud * e_total;
        float s = 0.0f;
        for (int j = j0; j < j1; ++j) {
            const long long r = cloud * e_total + __ldg(ord + j);
            s += __ldg(dout + r * C + c) * cv_slope(__ldg(out + r * C + c));
        }
        dp2[d * C + c] = s;
    }
}

// out[m] = sum over g < G of part[g * M + m], in order
__global__ void cv_reduce_partials_kernel(const float *__restrict__ part, int G, int M, float *__restrict__ out) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    float s = 0.0f;
    for (int g = 0; g < G; ++g) s += part[(long long)g * M + m];
    out[m] = s;
}

// ---- act_grad --------------------------------------------------------------------------------------------------------
// g = dy * act'(y) (act 1: ReLU, 2: LeakyReLU 0.1), amax = max |g| (fp32 bit pattern, atomicMax: order-independent),
// per-CTA column sums of g.  Thread = column (n <= blockDim), CTA = a contiguous block of rows.
__global__ void act_grad_kernel(long long rows, int n, int act, const float *__restrict__ y, const float *__restrict__ dy,
                                float *__restrict__ g, unsigned int *__restrict__ amax, float *__restrict__ part) {
    const int c = threadIdx.x % n, sub = threadIdx.x / n, nsub = blockDim.x / n;   // nsub rows in flight per step
    const long long per = (rows + gridDim.x - 1) / gridDim.x;
    const long long r0 = (long long)blockIdx.x * per, r1 = min(rows, r0 + per);
    float colsum = 0.0f, m = 0.0f;
    if (sub < nsub)
        for (long long r = r0 + sub; r < r1; r += nsub) {
            const float yv = __ldg(y + r * n + c), d = __ldg(dy + r * n + c);
            const float gv = d * (yv > 0.0f ? 1.0f : (act == 2 ? 0.1f : 0.0f));
            g[r * n + c] = gv;
            colsum += gv;
            m = fmaxf(m, fabsf(gv));
        }
    const uint32_t wm = rt_redux_max_u32(__float_as_uint(m));
    if ((threadIdx.x & 31) == 0 && wm) atomicMax(amax, wm);
    // column sums: the nsub row slots of a column are added in slot order through shared memory
    extern __shared__ float s_cs[];
    s_cs[threadIdx.x] = colsum;
    __syncthreads();
    if (threadIdx.x < n) {
        float s = 0.0f;
        for (int u = 0; u < nsub; ++u) s += s_cs[u * n + threadIdx.x];
        part[(long long)blockIdx.x * n + threadIdx.x] = s;
    }
}

// ---- wsum --------------------------------------------------------------------------------------------------------------
// out[p,c] = sum_k ReLU(W3[c,:].h2[p,k,:] + b3[c]) * X[p,k,c];  X = x[(p*K+k)] (idx == null) or xpts[cloud*n + idx[p,k]]
template <int KT>
__global__ void wsum_fwd_kernel(long long total_pts, int n, int Krt, int C, const float *__restrict__ x, const int *__restrict__ idx,
                                const float *__restrict__ h2, const float *__restrict__ w3, const float *__restrict__ b3,
                                float *__restrict__ out) {
    const int c = threadIdx.x;
    float w[CV_H];
#pragma unroll
    for (int i = 0; i < CV_H; ++i) w[i] = __ldg(w3 + c * CV_H + i);
    const float bv = __ldg(b3 + c);
    const int K = KT ? KT : Krt;
    if constexpr (KT > 0) {
        // The point's hidden rows (and neighbour indices) are the same for every channel: the CTA stages them in shared memory
        // one point ahead, so a thread's only global loads are its KT independent x values -- all in flight together -- instead of
        // KT dependent index -> row chains plus 2 KT broadcast loads that the register budget serialised.
        __shared__ __align__(16) float s_h2[2][KT * CV_H];
        __shared__ int s_idx[2][KT];
        auto stage = [&](long long q, int buf) {
            for (int i = threadIdx.x; i < KT * (CV_H + 1); i += blockDim.x) {
                if (i < KT * CV_H) s_h2[buf][i] = __ldg(h2 + q * KT * CV_H + i);
                else if (idx) s_idx[buf][i - KT * CV_H] = __ldg(idx + q * KT + (i - KT * CV_H));
            }
        };
        long long p = blockIdx.x;
        int buf = 0;
        if (p < total_pts) stage(p, 0);
        __syncthreads();
        for (; p < total_pts; p += gridDim.x, buf ^= 1) {
            const long long cloud = cv_div(p, n);
            float xv[KT];
#pragma unroll
            for (int k = 0; k < KT; ++k) {
                const long long xr = idx ? cloud * n + s_idx[buf][k] : p * KT + k;
                xv[k] = __ldg(x + xr * C + c);
            }
            if (p + gridDim.x < total_pts) stage(p + gridDim.x, buf ^ 1);   // the other buffer was last read before the previous barrier
            float s = 0.0f;
#pragma unroll
            for (int k = 0; k < KT; ++k) {
                const float4 ha = *reinterpret_cast<const float4 *>(&s_h2[buf][k * CV_H]), hb = *reinterpret_cast<const float4 *>(&s_h2[buf][k * CV_H + 4]);
                float wn = bv;
                wn = fmaf(w[0], ha.x, wn); wn = fmaf(w[1], ha.y, wn); wn = fmaf(w[2], ha.z, wn); wn = fmaf(w[3], ha.w, wn);
                wn = fmaf(w[4], hb.x, wn); wn = fmaf(w[5], hb.y, wn); wn = fmaf(w[6], hb.z, wn); wn = fmaf(w[7], hb.w, wn);
                s = fmaf(fmaxf(wn, 0.0f), xv[k], s);
            }
            out[p * C + c] = s;
            __syncthreads();
        }
        return;
    }
    for (long long p = blockIdx.x; p < total_pts; p += gridDim.x) {
        const long long cloud = cv_div(p, n);
        float s = 0.0f;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const long long r = p * K + k;
            const float4 ha = __ldg(reinterpret_cast<const float4 *>(h2 + r * CV_H)), hb = __ldg(reinterpret_cast<const float4 *>(h2 + r * CV_H) + 1);
            float wn = bv;
            wn = fmaf(w[0], ha.x, wn); wn = fmaf(w[1], ha.y, wn); wn = fmaf(w[2], ha.z, wn); wn = fmaf(w[3], ha.w, wn);
            wn = fmaf(w[4], hb.x, wn); wn = fmaf(w[5], hb.y, wn); wn = fmaf(w[6], hb.z, wn); wn = fmaf(w[7], hb.w, wn);
            const long long xr = idx ? cloud * n + __ldg(idx + r) : r;
            s = fmaf(fmaxf(wn, 0.0f), __ldg(x + xr * C + c), s);
        }
        out[p * C + c] = s;
    }
}

// backward: dx rows (idx == null), dh2 (p,k,8), per-CTA partials of dW3 (C x 8) and db3 (C).  Thread = channel.
// dh2[p,k,i] = sum_c W3[c,i] * dwn[p,k,c] crosses the threads: the point's dwn (K x C) goes through shared memory and the
// first K * 8 threads add it over the channels in channel order.
template <int KT>
__global__ void wsum_bwd_kernel(long long total_pts, int n, int Krt, int C, const float *__restrict__ x, const int *__restrict__ idx,
                                const float *__restrict__ h2, const float *__restrict__ w3, const float *__restrict__ b3,
                                const float *__restrict__ dout, float *__restrict__ dx, float *__restrict__ dh2, float *__restrict__ part) {
    extern __shared__ float s_mem[];
    float *s_w3 = s_mem;                       // [C][8]
    float *s_dwn = s_mem + C * CV_H;           // [K][C + 1]
    const int c = threadIdx.x;
    float w[CV_H], aw[CV_H];
#pragma unroll
    for (int i = 0; i < CV_H; ++i) {
        w[i] = __ldg(w3 + c * CV_H + i);
        s_w3[c * CV_H + i] = w[i];
        aw[i] = 0.0f;
    }
    const float bv = __ldg(b3 + c);
    const int K = KT ? KT : Krt;
    float ab = 0.0f;
    __syncthreads();
    constexpr int KS = KT > 0 ? KT : 1;
    __shared__ __align__(16) float s_h2[2][KS * CV_H];   // KT > 0: the point's hidden rows / neighbour indices, staged one point ahead
    __shared__ int s_idx[2][KS];
    auto stage = [&](long long q, int buf) {
        for (int i = threadIdx.x; i < KS * (CV_H + 1); i += blockDim.x) {
            if (i < KS * CV_H) s_h2[buf][i] = __ldg(h2 + q * KS * CV_H + i);
            else if (idx) s_idx[buf][i - KS * CV_H] = __ldg(idx + q * KS + (i - KS * CV_H));
        }
    };
    int buf = 0;
    if (KT > 0 && (long long)blockIdx.x < total_pts) {
        stage(blockIdx.x, 0);
        __syncthreads();
    }
    for (long long p = blockIdx.x; p < total_pts; p += gridDim.x, buf ^= 1) {
        const long long cloud = cv_div(p, n);
        const float dv = __ldg(dout + p * C + c);
        float xv[KS];
        if constexpr (KT > 0) {
#pragma unroll
            for (int k = 0; k < KT; ++k) {
                const long long xr = idx ? cloud * n + s_idx[buf][k] : p * KT + k;
                xv[k] = __ldg(x + xr * C + c);     // KT independent loads, all in flight
            }
            if (p + gridDim.x < total_pts) stage(p + gridDim.x, buf ^ 1);   // read next after the two barriers below
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const long long r = p * K + k;
            float4 ha, hb;
            if constexpr (KT > 0) {
                ha = *reinterpret_cast<const float4 *>(&s_h2[buf][k * CV_H]);
                hb = *reinterpret_cast<const float4 *>(&s_h2[buf][k * CV_H + 4]);
            } else {
                ha = __ldg(reinterpret_cast<const float4 *>(h2 + r * CV_H));
                hb = __ldg(reinterpret_cast<const float4 *>(h2 + r * CV_H) + 1);
            }
            float wn = bv;
            wn = fmaf(w[0], ha.x, wn); wn = fmaf(w[1], ha.y, wn); wn = fmaf(w[2], ha.z, wn); wn = fmaf(w[3], ha.w, wn);
            wn = fmaf(w[4], hb.x, wn); wn = fmaf(w[5], hb.y, wn); wn = fmaf(w[6], hb.z, wn); wn = fmaf(w[7], hb.w, wn);
            float xk;
            if constexpr (KT > 0) xk = xv[k];
            else xk = __ldg(x + (idx ? cloud * n + __ldg(idx + r) : r) * C + c);
            if (dx) dx[r * C + c] = fmaxf(wn, 0.0f) * dv;
            const float dwn = wn > 0.0f ? xk * dv : 0.0f;
            s_dwn[k * (C + 1) + c] = dwn;
            ab += dwn;
            aw[0] = fmaf(dwn, ha.x, aw[0]); aw[1] = fmaf(dwn, ha.y, aw[1]); aw[2] = fmaf(dwn, ha.z, aw[2]); aw[3] = fmaf(dwn, ha.w, aw[3]);
            aw[4] = fmaf(dwn, hb.x, aw[4]); aw[5] = fmaf(dwn, hb.y, aw[5]); aw[6] = fmaf(dwn, hb.z, aw[6]); aw[7] = fmaf(dwn, hb.w, aw[7]);
        }
        __syncthreads();
        // two threads per (k, i): each adds half of the channels in channel order, the halves are added by one shuffle.
        // Every thread of the CTA runs every trip (the shuffle names the full warp): work beyond K * 16 is predicated off.
        for (int t0 = 0; t0 < K * CV_H * 2; t0 += blockDim.x) {
            const int t = t0 + threadIdx.x;
            const bool live = t < K * CV_H * 2;
            const int pair = t >> 1, half = t & 1;
            const int k = live ? pair / CV_H : 0, i = pair % CV_H;
            const int c0 = half * (C / 2), c1 = live ? c0 + C / 2 : c0;
            float s = 0.0f;
            for (int cc = c0; cc < c1; ++cc) s = fmaf(s_w3[cc * CV_H + i], s_dwn[k * (C + 1) + cc], s);
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            if (live && half == 0) dh2[(p * K + k) * CV_H + i] = s;
        }
        __syncthreads();
    }
    float *pp = part + (long long)blockIdx.x * (CV_H + 1) * C;
#pragma unroll
    for (int i = 0; i < CV_H; ++i) pp[c * CV_H + i] = aw[i];
    pp[CV_H * C + c] = ab;
}

// gathered variant: dxpts[cloud, t, c] = sum over the sources e = (point, k) of t of ReLU(W3.h2[e] + b3)[c] * dout[point, c]
__global__ void wsum_bwd_scatter_kernel(long long total_dst, int n, int K, int C, const float *__restrict__ h2, const float *__restrict__ w3,
                                        const float *__restrict__ b3, const float *__restrict__ dout, const int *__restrict__ order,
                                        const int *__restrict__ seg, float *__restrict__ dxpts) {
    const int c = threadIdx.x;
    float w[CV_H];
#pragma unroll
    for (int i = 0; i < CV_H; ++i) w[i] = __ldg(w3 + c * CV_H + i);
    const float bv = __ldg(b3 + c);
    const long long e_total = (long long)n * K;
    for (long long d = blockIdx.x; d < total_dst; d += gridDim.x) {
        const long long cloud = cv_div(d, n);
        const int t = (int)(d - cloud * n);
        const int *sg = seg + cloud * (n + 1);
        const int j0 = __ldg(sg + t), j1 = __ldg(sg + t + 1);
        const int *ord = order + cloud * e_total;
        float s = 0.0f;
        for (int j = j0; j < j1; ++j) {
            const long long r = cloud * e_total + __ldg(ord + j);
            const float4 ha = __ldg(reinterpret_cast<const float4 *>(h2 + r * CV_H)), hb = __ldg(reinterpret_cast<const float4 *>(h2 + r * CV_H) + 1);
            float wn = bv;
            wn = fmaf(w[0], ha.x, wn); wn = fmaf(w[1], ha.y, wn); wn = fmaf(w[2], ha.z, wn); wn = fmaf(w[3], ha.w, wn);
            wn = fmaf(w[4], hb.x, wn); wn = fmaf(w[5], hb.y, wn); wn = fmaf(w[6], hb.z, wn); wn = fmaf(w[7], hb.w, wn);
            s = fmaf(fmaxf(wn, 0.0f), __ldg(dout + cv_div(r, K) * C + c), s);
        }
        dxpts[d * C + c] = s;
    }
}

bool cv_bad_c(int C) { return C < 32 || C > 1024 || (C & 31) != 0; }

}  // namespace

RT_API int rt_cv1_forward(int b, int n1, int n2, int k, int c, const float *p1, const float *p2, const float *xyz1, const float *xyz2,
                          const int *idx, const float *wd, const float *bias, float *out, float *dir_out, void *stream) {
    RT_REQUIRE(b >= 0 && n1 >= 1 && n2 >= 1 && k >= 1 && k <= CV_MAXK, "rt_cv1_forward: bad sizes");
    RT_REQUIRE(!cv_bad_c(c), "rt_cv1_forward: channels must be a multiple of 32 in [32, 1024]");
    RT_REQUIRE(p1 && p2 && xyz1 && xyz2 && idx && wd && out, "rt_cv1_forward: null pointer");
    const long long pts = (long long)b * n1;
    if (pts == 0) return RT_OK;
    const int grid = (int)min(pts, (long long)cv_sms() * 8);
    if (k == 16) cv1_fwd_kernel<16><<<grid, c, 0, (cudaStream_t)stream>>>(pts, n1, n2, k, c, p1, p2, xyz1, xyz2, idx, wd, bias, out, dir_out);
    else cv1_fwd_kernel<0><<<grid, c, 0, (cudaStream_t)stream>>>(pts, n1, n2, k, c, p1, p2, xyz1, xyz2, idx, wd, bias, out, dir_out);
    return rt_check_launch("cv1_fwd_kernel");
}

RT_API int rt_cv1_backward(int b, int n1, int n2, int k, int c, const float *dout, const float *out, const float *dir, const int *idx,
                           float *dp1, float *dp2, float *dwd, float *dbias, void *stream) {
    RT_REQUIRE(b >= 0 && n1 >= 1 && n2 >= 1 && k >= 1 && k <= CV_MAXK, "rt_cv1_backward: bad sizes");
    RT_REQUIRE(!cv_bad_c(c), "rt_cv1_backward: channels must be a multiple of 32 in [32, 1024]");
    RT_REQUIRE(dout && out && dir && idx && dp1 && dp2 && dwd && dbias, "rt_cv1_backward: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const long long pts = (long long)b * n1, e_total = (long long)n1 * k;
    if (pts == 0) return RT_OK;
    if (!rt_segsum_supported(n2, e_total) || b > 65535) {
        rt_set_error("rt_cv1_backward: n2 = %d / %lld rows per cloud is outside the inverse-index kernel's range", n2, e_total);
        return RT_ERR_UNSUPPORTED;
    }
    const int grid = (int)min(pts, (long long)cv_sms() * 4);
    float *part = nullptr;
    int *scratch = nullptr;
    int rc = rt_scratch_alloc((void **)&part, sizeof(float) * (size_t)grid * 4 * c, st, "rt_cv1_backward");
    if (rc != RT_OK) return rc;
    rc = rt_scratch_alloc((void **)&scratch, sizeof(int) * (size_t)b * ((size_t)e_total + n2 + 1), st, "rt_cv1_backward");
    if (rc != RT_OK) { rt_scratch_free(part, st); return rc; }
    int *order = scratch, *seg = scratch + (size_t)b * e_total;
    cv1_bwd_point_kernel<<<grid, c, 0, st>>>(pts, k, c, dout, out, dir, dp1, part);
    rc = rt_check_launch("cv1_bwd_point_kernel");
    if (rc == RT_OK) {
        // part = [grid][wx | wy | wz | b][c] -> dwd (c,3) needs a transpose: reduce into a small staging row, then scatter
        float *red = nullptr;
        rc = rt_scratch_alloc((void **)&red, sizeof(float) * 4 * c, st, "rt_cv1_backward");
        if (rc == RT_OK) {
            cv_reduce_partials_kernel<<<(4 * c + 255) / 256, 256, 0, st>>>(part, grid, 4 * c, red);
            cudaMemcpy2DAsync(dwd + 0, 3 * sizeof(float), red, sizeof(float), sizeof(float), c, cudaMemcpyDeviceToDevice, st);
            cudaMemcpy2DAsync(dwd + 1, 3 * sizeof(float), red + c, sizeof(float), sizeof(float), c, cudaMemcpyDeviceToDevice, st);
            cudaMemcpy2DAsync(dwd + 2, 3 * sizeof(float), red + 2 * c, sizeof(float), sizeof(float), c, cudaMemcpyDeviceToDevice, st);
            cudaMemcpyAsync(dbias, red + 3 * c, sizeof(float) * c, cudaMemcpyDeviceToDevice, st);
            rc = rt_check_launch("cv_reduce_partials_kernel");
            rt_scratch_free(red, st);
        }
    }
    if (rc == RT_OK) rc = rt_launch_inverse_index(b, n2, e_total, idx, order, seg, st, "rt_cv1_backward");
    if (rc == RT_OK) {
        const long long dsts = (long long)b * n2;
        cv1_bwd_scatter_kernel<<<(int)min(dsts, (long long)cv_sms() * 8), c, 0, st>>>(dsts, n2, e_total, c, dout, out, order, seg, dp2);
        rc = rt_check_launch("cv1_bwd_scatter_kernel");
    }
    rt_scratch_free(scratch, st);
    rt_scratch_free(part, st);
    return rc;
}

RT_API int rt_act_grad(long long rows, int n, int act, const float *y, const float *dy, float *g, float *amax_out, float *colsum, void *stream) {
    RT_REQUIRE(rows >= 0 && n >= 1 && n <= 1024 && (act == 1 || act == 2), "rt_act_grad: bad arguments");
    RT_REQUIRE(y && dy && g && amax_out && colsum, "rt_act_grad: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(amax_out, 0, 4, st);
    if (rows == 0) {
        cudaMemsetAsync(colsum, 0, sizeof(float) * n, st);
        return RT_OK;
    }
    const int nsub = n >= 256 ? 1 : 256 / n;
    const int threads = (nsub * n + 31) / 32 * 32;
    const int grid = (int)min((rows + 63) / 64, (long long)cv_sms() * 8);
    float *part = nullptr;
    int rc = rt_scratch_alloc((void **)&part, sizeof(float) * (size_t)grid * n, st, "rt_act_grad");
    if (rc != RT_OK) return rc;
    act_grad_kernel<<<grid, threads, sizeof(float) * threads, st>>>(rows, n, act, y, dy, g, reinterpret_cast<unsigned int *>(amax_out), part);
    rc = rt_check_launch("act_grad_kernel");
    if (rc == RT_OK) {
        cv_reduce_partials_kernel<<<(n + 255) / 256, 256, 0, st>>>(part, grid, n, colsum);
        rc = rt_check_launch("cv_reduce_partials_kernel");
    }
    rt_scratch_free(part, st);
    return rc;
}

RT_API int rt_wsum_forward(int b, int n, int k, int c, const float *x, const int *idx, const float *h2, const float *w3, const float *b3,
                           float *out, void *stream) {
    RT_REQUIRE(b >= 0 && n >= 1 && k >= 1 && k <= CV_MAXK, "rt_wsum_forward: bad sizes");
    RT_REQUIRE(!cv_bad_c(c), "rt_wsum_forward: channels must be a multiple of 32 in [32, 1024]");
    RT_REQUIRE(x && h2 && w3 && b3 && out, "rt_wsum_forward: null pointer");
    const long long pts = (long long)b * n;
    if (pts == 0) return RT_OK;
    const int grid = (int)min(pts, (long long)cv_sms() * 8);
    if (k == 16) wsum_fwd_kernel<16><<<grid, c, 0, (cudaStream_t)stream>>>(pts, n, k, c, x, idx, h2, w3, b3, out);
    else wsum_fwd_kernel<0><<<grid, c, 0, (cudaStream_t)stream>>>(pts, n, k, c, x, idx, h2, w3, b3, out);
    return rt_check_launch("wsum_fwd_kernel");
}

RT_API int rt_wsum_backward(int b, int n, int k, int c, const float *x, const int *idx, const float *h2, const float *w3, const float *b3,
                            const float *dout, float *dx, float *dh2, float *dw3, float *db3, void *stream) {
    RT_REQUIRE(b >= 0 && n >= 1 && k >= 1 && k <= CV_MAXK, "rt_wsum_backward: bad sizes");
    RT_REQUIRE(!cv_bad_c(c), "rt_wsum_backward: channels must be a multiple of 32 in [32, 1024]");
    RT_REQUIRE(x && h2 && w3 && b3 && dout && dx && dh2 && dw3 && db3, "rt_wsum_backward: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const long long pts = (long long)b * n, e_total = (long long)n * k;
    if (pts == 0) return RT_OK;
    if (idx && (!rt_segsum_supported(n, e_total) || b > 65535)) {
        rt_set_error("rt_wsum_backward: n = %d / %lld rows per cloud is outside the inverse-index kernel's range", n, e_total);
        return RT_ERR_UNSUPPORTED;
    }
    const int grid = (int)min(pts, (long long)cv_sms() * 4);
    const int M = (CV_H + 1) * c;
    float *part = nullptr;
    int rc = rt_scratch_alloc((void **)&part, sizeof(float) * (size_t)(grid + 1) * M, st, "rt_wsum_backward");
    if (rc != RT_OK) return rc;
    const size_t smem = sizeof(float) * ((size_t)c * CV_H + (size_t)k * (c + 1));
    static RtPerDevice attr_set;
    if (smem > 48 * 1024 && !attr_set.done(rt_current_device())) {
        cudaError_t e = cudaFuncSetAttribute(wsum_bwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(wsum_bwd_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) {
            rt_set_error("rt_wsum_backward: cannot reserve shared memory: %s", cudaGetErrorString(e));
            rt_scratch_free(part, st);
            return (int)e;
        }
        attr_set.mark(rt_current_device());
    }
    if (k == 16) wsum_bwd_kernel<16><<<grid, c, smem, st>>>(pts, n, k, c, x, idx, h2, w3, b3, dout, idx ? nullptr : dx, dh2, part);
    else wsum_bwd_kernel<0><<<grid, c, smem, st>>>(pts, n, k, c, x, idx, h2, w3, b3, dout, idx ? nullptr : dx, dh2, part);
    rc = rt_check_launch("wsum_bwd_kernel");
    if (rc == RT_OK) {
        float *red = part + (size_t)grid * M;
        cv_reduce_partials_kernel<<<(M + 255) / 256, 256, 0, st>>>(part, grid, M, red);
        cudaMemcpyAsync(dw3, red, sizeof(float) * CV_H * c, cudaMemcpyDeviceToDevice, st);
        cudaMemcpyAsync(db3, red + CV_H * c, sizeof(float) * c, cudaMemcpyDeviceToDevice, st);
        rc = rt_check_launch("cv_reduce_partials_kernel");
    }
    if (rc == RT_OK && idx) {
        int *scratch = nullptr;
        rc = rt_scratch_alloc((void **)&scratch, sizeof(int) * (size_t)b * ((size_t)e_total + n + 1), st, "rt_wsum_backward");
        if (rc == RT_OK) {
            int *order = scratch, *seg = scratch + (size_t)b * e_total;
            rc = rt_launch_inverse_index(b, n, e_total, idx, order, seg, st, "rt_wsum_backward");
            if (rc == RT_OK) {
                wsum_bwd_scatter_kernel<<<(int)min(pts, (long long)cv_sms() * 8), c, 0, st>>>(pts, n, k, c, h2, w3, b3, dout, order, seg, dx);
                rc = rt_check_launch("wsum_bwd_scatter_kernel");
            }
            rt_scratch_free(scratch, st);
        }
    }
    rt_scratch_free(part, st);
    return rc;
}
