// Neighbour-search kernels for sm_100a: ball_query, three_nn, knn.
//
// Replace ball_query_kernel_fast (reference: src/lib/src/ball_query_gpu.cu:9-45),
// three_nn_kernel_fast (src/lib/src/interpolate_gpu.cu:81-124) and knn_kernel_fast
// (src/lib/src/interpolate_gpu.cu:9-57) with bit-identical results.
//
// The reference walks the whole cloud from global memory with one thread per query
// (12-byte strided, uncoalesced, divergent early exit).  Here the searched cloud is staged
// into shared memory in tiles by the bulk-copy engine (TMA, UBLKCP) and
//   * ball_query uses one WARP per query: 32 candidates tested per step, ballot + popcount
//     gives each hit its slot in index order (the reference's "first nsample in scan order"),
//     whole-warp early exit once nsample hits are found;
//   * three_nn / knn use one thread per query reading the tile with shared-memory broadcasts.
#include "common.cuh"

namespace {

constexpr int TILE_PTS = 2048;  // 24 KB of xyz per tile (two CTAs per SM stay under the 48 KB default)

// ------------------------------------------------------------------------------------------
// ball_query: grid (ceil(m / QPB), b), 256 threads = 8 warps, each warp walks QPB/8 queries.
constexpr int BQ_THREADS = 256;
constexpr int BQ_QPB = 64;

// Two radii in one pass (the MSG set-abstraction layers always query the same centres with two radii): the distance
// of a candidate is evaluated once and feeds two independent ordered compactions.  nsample_b == 0 disables the second.
__global__ void __launch_bounds__(BQ_THREADS) ball_query_kernel(int n, int m, float radius_a, int nsample_a, int *__restrict__ idx_a,
                                                                float radius_b, int nsample_b, int *__restrict__ idx_b,
                                                                const float *__restrict__ new_xyz,
                                                                const float *__restrict__ xyz) {
    extern __shared__ __align__(16) float s_pts[];  // min(n, TILE_PTS)*3
    __shared__ __align__(8) uint64_t s_bar;
    const int cloud = blockIdx.y;
    const float *pts = xyz + (size_t)cloud * n * 3;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float r2a = __fmul_rn(radius_a, radius_a), r2b = __fmul_rn(radius_b, radius_b);
    constexpr int QPW = BQ_QPB / (BQ_THREADS / 32);

    if (threadIdx.x == 0) {
        rt_mbar_init(&s_bar, 1);
        rt_fence_mbar_init();
    }
    __syncthreads();

    const int q0 = blockIdx.x * BQ_QPB + warp * QPW;
    float qx[QPW], qy[QPW], qz[QPW];
    int cnt_a[QPW], first_a[QPW], cnt_b[QPW], first_b[QPW];
#pragma unroll
    for (int i = 0; i < QPW; ++i) {
        const int q = q0 + i;
        const bool ok = q < m;
        const float *c = new_xyz + ((size_t)cloud * m + (ok ? q : 0)) * 3;
        qx[i] = __ldg(c + 0);
        qy[i] = __ldg(c + 1);
        qz[i] = __ldg(c + 2);
        cnt_a[i] = ok ? 0 : nsample_a;  // out-of-range queries are "done"
        cnt_b[i] = ok ? 0 : nsample_b;
        first_a[i] = first_b[i] = -1;
    }

    uint32_t parity = 0;
    for (int base = 0; base < n; base += TILE_PTS) {
        const int tn = min(TILE_PTS, n - base);
        if (base > 0) __syncthreads();  // everyone finished reading the previous tile
        rt_stage_floats(s_pts, pts + (size_t)base * 3, tn * 3, &s_bar, parity);
        parity ^= 1;
#pragma unroll
        for (int i = 0; i < QPW; ++i) {
            int ca = cnt_a[i], cb = cnt_b[i], fa = first_a[i], fb = first_b[i];
            if (ca >= nsample_a && cb >= nsample_b) continue;  // warp-uniform
            int *oa = idx_a + ((size_t)cloud * m + (q0 + i)) * nsample_a;
            int *ob = idx_b + ((size_t)cloud * m + (q0 + i)) * nsample_b;
            for (int k0 = 0; k0 < tn && (ca < nsample_a || cb < nsample_b); k0 += 32) {
                const int k = k0 + lane;
                bool hit_a = false, hit_b = false;
                if (k < tn) {
                    const float d2 = rt_sqdist(qx[i], qy[i], qz[i], s_pts[k * 3 + 0], s_pts[k * 3 + 1], s_pts[k * 3 + 2]);
                    hit_a = d2 < r2a;
                    hit_b = d2 < r2b;
                }
                if (ca < nsample_a) {
                    const uint32_t vote = __ballot_sync(0xffffffffu, hit_a);
                    if (vote) {
                        if (fa < 0) fa = base + k0 + __ffs(vote) - 1;
                        const int slot = ca + __popc(vote & ((1u << lane) - 1u));
                        if (hit_a && slot < nsample_a) oa[slot] = base + k;
                        ca += __popc(vote);
                    }
                }
                if (cb < nsample_b) {
                    const uint32_t vote = __ballot_sync(0xffffffffu, hit_b);
                    if (vote) {
                        if (fb < 0) fb = base + k0 + __ffs(vote) - 1;
                        const int slot = cb + __popc(vote & ((1u << lane) - 1u));
                        if (hit_b && slot < nsample_b) ob[slot] = base + k;
                        cb += __popc(vote);
                    }
                }
            }
            cnt_a[i] = ca; cnt_b[i] = cb;
            first_a[i] = fa; first_b[i] = fb;
        }
    }
    // pad unused slots with the first hit; queries with no hit leave the caller's buffer untouched
#pragma unroll
    for (int i = 0; i < QPW; ++i) {
        if (q0 + i >= m) continue;
        if (first_a[i] >= 0) {
            int *out = idx_a + ((size_t)cloud * m + (q0 + i)) * nsample_a;
            for (int s = cnt_a[i] + lane; s < nsample_a; s += 32) out[s] = first_a[i];
        }
        if (first_b[i] >= 0) {
            int *out = idx_b + ((size_t)cloud * m + (q0 + i)) * nsample_b;
            for (int s = cnt_b[i] + lane; s < nsample_b; s += 32) out[s] = first_b[i];
        }
    }
}

// ------------------------------------------------------------------------------------------
// three_nn: one thread per unknown point; known cloud in shared-memory tiles.
constexpr int NN_THREADS = 128;

__global__ void __launch_bounds__(NN_THREADS) three_nn_kernel(int n, int m, const float *__restrict__ unknown,
                                                              const float *__restrict__ known,
                                                              float *__restrict__ dist2, int *__restrict__ idx) {
    extern __shared__ __align__(16) float s_pts[];
    __shared__ __align__(8) uint64_t s_bar;
    const int cloud = blockIdx.y;
    const int j = blockIdx.x * NN_THREADS + threadIdx.x;
    const bool ok = j < n;
    const float *u = unknown + ((size_t)cloud * n + (ok ? j : 0)) * 3;
    const float ux = __ldg(u + 0), uy = __ldg(u + 1), uz = __ldg(u + 2);
    if (threadIdx.x == 0) {
        rt_mbar_init(&s_bar, 1);
        rt_fence_mbar_init();
    }
    __syncthreads();
    // the reference keeps double-typed bests initialised to 1e40: every finite fp32 distance is
    // smaller, +inf/NaN never are -- identical to fp32 bests initialised to +inf.
    float b1 = __int_as_float(0x7f800000), b2 = b1, b3 = b1;
    int i1 = 0, i2 = 0, i3 = 0;
    uint32_t parity = 0;
    for (int base = 0; base < m; base += TILE_PTS) {
        const int tn = min(TILE_PTS, m - base);
        if (base > 0) __syncthreads();
        rt_stage_floats(s_pts, known + ((size_t)cloud * m + base) * 3, tn * 3, &s_bar, parity);
        parity ^= 1;
#pragma unroll 4
        for (int k = 0; k < tn; ++k) {
            const float d = rt_sqdist(ux, uy, uz, s_pts[k * 3 + 0], s_pts[k * 3 + 1], s_pts[k * 3 + 2]);
            if (d < b3) {
                const int kk = base + k;
                if (d < b1) {
                    b3 = b2; i3 = i2;
                    b2 = b1; i2 = i1;
                    b1 = d; i1 = kk;
                } else if (d < b2) {
                    b3 = b2; i3 = i2;
                    b2 = d; i2 = kk;
                } else {
                    b3 = d; i3 = kk;
                }
            }
        }
    }
    if (ok) {
        float *od = dist2 + ((size_t)cloud * n + j) * 3;
        int *oi = idx + ((size_t)cloud * n + j) * 3;
        od[0] = b1; od[1] = b2; od[2] = b3;
        oi[0] = i1; oi[1] = i2; oi[2] = i3;
    }
}

// ------------------------------------------------------------------------------------------
// knn (k <= 200, ascending, stable on ties): one thread per query, sorted list in per-thread
// scratch carved from shared memory when it fits (k <= 32) else local memory; candidates
// are rejected against the current k-th best before any list traffic.
template <int KMAX>
__global__ void __launch_bounds__(NN_THREADS) knn_kernel(int n, int m, int k, const float *__restrict__ unknown,
                                                         const float *__restrict__ known,
                                                         float *__restrict__ dist2, int *__restrict__ idx) {
    extern __shared__ __align__(16) float s_pts[];
    __shared__ __align__(8) uint64_t s_bar;
    const int cloud = blockIdx.y;
    const int j = blockIdx.x * NN_THREADS + threadIdx.x;
    const bool ok = j < n;
    const float *u = unknown + ((size_t)cloud * n + (ok ? j : 0)) * 3;
    const float ux = __ldg(u + 0), uy = __ldg(u + 1), uz = __ldg(u + 2);
    if (threadIdx.x == 0) {
        rt_mbar_init(&s_bar, 1);
        rt_fence_mbar_init();
    }
    __syncthreads();
    const float inf = __int_as_float(0x7f800000);
    float best[KMAX];
    int besti[KMAX];
#pragma unroll
    for (int i = 0; i < KMAX; ++i) {
        best[i] = inf;
        besti[i] = 0;
    }
    float worst = (k > 0) ? inf : -inf;
    uint32_t parity = 0;
    for (int base = 0; base < m; base += TILE_PTS) {
        const int tn = min(TILE_PTS, m - base);
        if (base > 0) __syncthreads();
        rt_stage_floats(s_pts, known + ((size_t)cloud * m + base) * 3, tn * 3, &s_bar, parity);
        parity ^= 1;
        for (int p = 0; p < tn; ++p) {
            const float d = rt_sqdist(ux, uy, uz, s_pts[p * 3 + 0], s_pts[p * 3 + 1], s_pts[p * 3 + 2]);
            if (d < worst) {
                // insert after every entry <= d (strict '<' in the reference => stable)
                int pos = k - 1;
                if (KMAX <= 32) {
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
                            besti[l] = base + p;
                        }
                } else {
                    while (pos > 0 && best[pos - 1] > d) {
                        best[pos] = best[pos - 1];
                        besti[pos] = besti[pos - 1];
                        --pos;
                    }
                    best[pos] = d;
                    besti[pos] = base + p;
                }
                worst = best[k - 1];
            }
        }
    }
    if (ok) {
        float *od = dist2 + ((size_t)cloud * n + j) * k;
        int *oi = idx + ((size_t)cloud * n + j) * k;
        if (KMAX <= 32) {
#pragma unroll
            for (int i = 0; i < KMAX; ++i)
                if (i < k) {
                    od[i] = best[i];
                    oi[i] = besti[i];
                }
        } else {
            for (int i = 0; i < k; ++i) {
                od[i] = best[i];
                oi[i] = besti[i];
            }
        }
    }
}

inline size_t tile_bytes(int n) { return (size_t)min(n, TILE_PTS) * 3 * sizeof(float); }

}  // namespace

// replaces ball_query_wrapper_fast (reference: src/lib/src/ball_query.cpp:18-29)
RT_API int rt_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xyz, const float *xyz,
                         int *idx, void *stream) {
    RT_REQUIRE(b >= 0 && n >= 0 && m >= 0 && nsample >= 0 && new_xyz && xyz && idx, "ball_query: bad arguments");
    if (b == 0 || m == 0 || nsample == 0 || n == 0) return RT_OK;
    RT_REQUIRE(b <= 65535, "ball_query: batch > 65535");
    dim3 grid(rt_divup(m, BQ_QPB), b);
    ball_query_kernel<<<grid, BQ_THREADS, tile_bytes(n), (cudaStream_t)stream>>>(n, m, radius, nsample, idx, 0.0f, 0, idx,
                                                                                 new_xyz, xyz);
    return rt_check_launch("ball_query_kernel");
}

// engine-internal: two radii over the same centres in one launch (same results as two rt_ball_query calls)
int rt_launch_ball_query2(int b, int n, int m, float radius_a, int nsample_a, int *idx_a, float radius_b, int nsample_b,
                          int *idx_b, const float *new_xyz, const float *xyz, cudaStream_t st) {
    if (b == 0 || m == 0 || n == 0) return RT_OK;
    RT_REQUIRE(b <= 65535 && nsample_a > 0 && nsample_b > 0, "ball_query2: bad arguments");
    dim3 grid(rt_divup(m, BQ_QPB), b);
    ball_query_kernel<<<grid, BQ_THREADS, tile_bytes(n), st>>>(n, m, radius_a, nsample_a, idx_a, radius_b, nsample_b, idx_b,
                                                              new_xyz, xyz);
    return rt_check_launch("ball_query_kernel(2 radii)");
}

// replaces three_nn_wrapper_fast (reference: src/lib/src/interpolate.cpp:16-25)
RT_API int rt_three_nn(int b, int n, int m, const float *unknown, const float *known, float *dist2, int *idx,
                       void *stream) {
    RT_REQUIRE(b >= 0 && n >= 0 && m >= 0 && unknown && known && dist2 && idx, "three_nn: bad arguments");
    if (b == 0 || n == 0) return RT_OK;
    RT_REQUIRE(b <= 65535, "three_nn: batch > 65535");
    dim3 grid(rt_divup(n, NN_THREADS), b);
    three_nn_kernel<<<grid, NN_THREADS, tile_bytes(m), (cudaStream_t)stream>>>(n, m, unknown, known, dist2, idx);
    return rt_check_launch("three_nn_kernel");
}

// replaces knn_wrapper_fast (reference: src/lib/src/interpolate.cpp:27-36); k <= 200 as in
// interpolate_gpu.cu:30-31, but rejected with an error instead of overflowing the stack.
RT_API int rt_knn(int b, int n, int m, int k, const float *unknown, const float *known, float *dist2, int *idx,
                  void *stream) {
    RT_REQUIRE(b >= 0 && n >= 0 && m >= 0 && unknown && known && dist2 && idx, "knn: bad arguments");
    RT_REQUIRE(k >= 0 && k <= 200, "knn: k=%d outside [0, 200] (reference limit, interpolate_gpu.cu:30-31)", k);
    if (b == 0 || n == 0 || k == 0) return RT_OK;
    RT_REQUIRE(b <= 65535, "knn: batch > 65535");
    dim3 grid(rt_divup(n, NN_THREADS), b);
    cudaStream_t st = (cudaStream_t)stream;
    if (k <= 16)
        knn_kernel<16><<<grid, NN_THREADS, tile_bytes(m), st>>>(n, m, k, unknown, known, dist2, idx);
    else if (k <= 32)
        knn_kernel<32><<<grid, NN_THREADS, tile_bytes(m), st>>>(n, m, k, unknown, known, dist2, idx);
    else
        knn_kernel<200><<<grid, NN_THREADS, tile_bytes(m), st>>>(n, m, k, unknown, known, dist2, idx);
    return rt_check_launch("knn_kernel");
}
