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
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int TILE_PTS = 2048;  // 24 KB of xyz per tile (two CTAs per SM stay under the 48 KB default)

// ------------------------------------------------------------------------------------------
// ball_query: grid (ceil(m / QPB), b), 256 threads = 8 warps, each warp walks QPB/8 queries.
constexpr int BQ_THREADS = 256;
constexpr int BQ_QPB = 64;

// Two radii in one pass (the MSG set-abstraction layers always query the same centres with two radii): the distance
// of a candidate is evaluated once and feeds two independent ordered compactions.  nsample_b == 0 disables the second.
// A warp walks its queries two at a time over the same 32 candidates (one set of shared-memory loads, two independent
// dependency chains); the common step -- no candidate of the 32 inside the larger radius of either query -- costs one
// ballot per query and no branch into the compaction code.  The tile is padded to a multiple of 32 with +inf points
// (never inside any radius), so the scan has no tail test.
struct BqState { int ca, cb, fa, fb; };

__device__ __forceinline__ void bq_compact(bool hit_a, bool hit_b, int cand, int lane, int nsample_a, int nsample_b, int *oa,
                                           int *ob, BqState &st) {
    if (st.ca < nsample_a) {
        const uint32_t vote = __ballot_sync(0xffffffffu, hit_a);
        if (vote) {
            if (st.fa < 0) st.fa = cand - lane + __ffs(vote) - 1;
            const int slot = st.ca + __popc(vote & ((1u << lane) - 1u));
            if (hit_a && slot < nsample_a) oa[slot] = cand;
            st.ca += __popc(vote);
        }
    }
    if (st.cb < nsample_b) {
        const uint32_t vote = __ballot_sync(0xffffffffu, hit_b);
        if (vote) {
            if (st.fb < 0) st.fb = cand - lane + __ffs(vote) - 1;
            const int slot = st.cb + __popc(vote & ((1u << lane) - 1u));
            if (hit_b && slot < nsample_b) ob[slot] = cand;
            st.cb += __popc(vote);
        }
    }
}

__global__ void __launch_bounds__(BQ_THREADS) ball_query_kernel(int n, int m, float radius_a, int nsample_a, int *__restrict__ idx_a,
                                                                float radius_b, int nsample_b, int *__restrict__ idx_b,
                                                                const float *__restrict__ new_xyz,
                                                                const float *__restrict__ xyz, int zero_fill,
                                                                const int *__restrict__ counts) {
    extern __shared__ __align__(16) float s_pts[];  // roundup32(min(n, TILE_PTS))*3
    __shared__ __align__(8) uint64_t s_bar;
    const int cloud = blockIdx.y;
    const float *pts = xyz + (size_t)cloud * n * 3;
    if (counts) n = min(n, __ldg(counts + cloud));   // padded variable-size batch: only the cloud's own points are candidates
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float r2a = __fmul_rn(radius_a, radius_a), r2b = __fmul_rn(radius_b, radius_b);
    const float r2max = nsample_b > 0 ? fmaxf(r2a, r2b) : r2a;
    constexpr int QPW = BQ_QPB / (BQ_THREADS / 32);
    static_assert(QPW % 2 == 0, "queries are walked in pairs");

    if (threadIdx.x == 0) {
        rt_mbar_init(&s_bar, 1);
        rt_fence_mbar_init();
    }
    __syncthreads();

    const int q0 = blockIdx.x * BQ_QPB + warp * QPW;
    float qx[QPW], qy[QPW], qz[QPW];
    BqState st[QPW];
#pragma unroll
    for (int i = 0; i < QPW; ++i) {
        const int q = q0 + i;
        const bool ok = q < m;
        const float *c = new_xyz + ((size_t)cloud * m + (ok ? q : 0)) * 3;
        qx[i] = __ldg(c + 0);
        qy[i] = __ldg(c + 1);
        qz[i] = __ldg(c + 2);
        st[i].ca = ok ? 0 : nsample_a;  // out-of-range queries are "done"
        st[i].cb = ok ? 0 : nsample_b;
        st[i].fa = st[i].fb = -1;
    }

    uint32_t parity = 0;
    for (int base = 0; base < n; base += TILE_PTS) {
        const int tn = min(TILE_PTS, n - base);
        const int tpad = (tn + 31) & ~31;
        if (base > 0) __syncthreads();  // everyone finished reading the previous tile
        for (int t = tn * 3 + threadIdx.x; t < tpad * 3; t += BQ_THREADS) s_pts[t] = __int_as_float(0x7f800000);
        rt_stage_floats(s_pts, pts + (size_t)base * 3, tn * 3, &s_bar, parity);
        parity ^= 1;
#pragma unroll
        for (int i = 0; i < QPW; i += 2) {
            BqState s0 = st[i], s1 = st[i + 1];
            bool live0 = s0.ca < nsample_a || s0.cb < nsample_b, live1 = s1.ca < nsample_a || s1.cb < nsample_b;  // warp-uniform
            if (!live0 && !live1) continue;
            int *oa0 = idx_a + ((size_t)cloud * m + (q0 + i)) * nsample_a, *oa1 = oa0 + nsample_a;
            int *ob0 = idx_b + ((size_t)cloud * m + (q0 + i)) * nsample_b, *ob1 = ob0 + nsample_b;
            for (int k0 = 0; k0 < tpad && (live0 || live1); k0 += 32) {
                const int k = k0 + lane;
                const float px = s_pts[k * 3 + 0], py = s_pts[k * 3 + 1], pz = s_pts[k * 3 + 2];
                const float d0 = rt_sqdist(qx[i], qy[i], qz[i], px, py, pz);
                const float d1 = rt_sqdist(qx[i + 1], qy[i + 1], qz[i + 1], px, py, pz);
                const bool any0 = __any_sync(0xffffffffu, d0 < r2max), any1 = __any_sync(0xffffffffu, d1 < r2max);
                if (any0 && live0) {
                    bq_compact(d0 < r2a, d0 < r2b, base + k, lane, nsample_a, nsample_b, oa0, ob0, s0);
                    live0 = s0.ca < nsample_a || s0.cb < nsample_b;
                }
                if (any1 && live1) {
                    bq_compact(d1 < r2a, d1 < r2b, base + k, lane, nsample_a, nsample_b, oa1, ob1, s1);
                    live1 = s1.ca < nsample_a || s1.cb < nsample_b;
                }
            }
            st[i] = s0;
            st[i + 1] = s1;
        }
    }
    // pad unused slots with the first hit; queries with no hit leave the caller's buffer untouched (the reference's
    // Python zero-fills it first) unless `zero_fill` asks this kernel to write the zeros itself (engine: no memset launch)
#pragma unroll
    for (int i = 0; i < QPW; ++i) {
        if (q0 + i >= m) continue;
        if (st[i].fa >= 0 || zero_fill) {
            int *out = idx_a + ((size_t)cloud * m + (q0 + i)) * nsample_a;
            const int fill = st[i].fa >= 0 ? st[i].fa : 0;
            for (int s = st[i].ca + lane; s < nsample_a; s += 32) out[s] = fill;
        }
        if (st[i].fb >= 0 || zero_fill) {
            int *out = idx_b + ((size_t)cloud * m + (q0 + i)) * nsample_b;
            const int fill = st[i].fb >= 0 ? st[i].fb : 0;
            for (int s = st[i].cb + lane; s < nsample_b; s += 32) out[s] = fill;
        }
    }
}

// ------------------------------------------------------------------------------------------
// ball_query, one THREAD per query (round 2).  The warp-per-query kernel above pays a ballot / popcount / prefix step
// for every 32 candidates of every query (~13 issued instructions per distance); here a thread owns a query and walks
// the cloud from shared memory two candidates at a time with packed fp32 pairs (FADD2 / FMUL2 / FFMA2: every half is
// the same round-to-nearest operation as rt_sqdist, so the distances are bit-identical): 6 arithmetic instructions and
// one combined test per pair.  A hit (rare per step) branches into an in-order append, which is the reference's
// "first nsample in scan order" by construction.  A full list disables itself by dropping its threshold to -1 (no
// distance is negative); the warp leaves the scan when all of its 32 queries are full.
// Candidates are staged NEGATED and as pairs: (-x0,-x1,-y0,-y1) / (-z0,-z1), so q - p is one packed add; slots past
// the cloud hold -inf (distance +inf, never a hit).
constexpr int BQT_THREADS = 128;

__device__ __forceinline__ void bqt_append(float d, int k, float &ta, float &tb, int &ca, int &cb, int &fa, int &fb, int nsa, int nsb,
                                           int *__restrict__ oa, int *__restrict__ ob) {
    if (d < ta) {
        if (fa < 0) fa = k;
        oa[ca] = k;
        if (++ca == nsa) ta = -1.0f;
    }
    if (d < tb) {
        if (fb < 0) fb = k;
        ob[cb] = k;
        if (++cb == nsb) tb = -1.0f;
    }
}

__global__ void __launch_bounds__(BQT_THREADS) ball_query_thread_kernel(int n, int m, float radius_a, int nsample_a, int *__restrict__ idx_a,
                                                                         float radius_b, int nsample_b, int *__restrict__ idx_b,
                                                                         const float *__restrict__ new_xyz,
                                                                         const float *__restrict__ xyz, int zero_fill,
                                                                         const int *__restrict__ counts) {
    __shared__ float4 s_xy[TILE_PTS / 2];
    __shared__ float2 s_z[TILE_PTS / 2];
    const int cloud = blockIdx.y;
    const float *pts = xyz + (size_t)cloud * n * 3;
    if (counts) n = min(n, __ldg(counts + cloud));   // padded variable-size batch: only the cloud's own points are candidates
    const int q = blockIdx.x * BQT_THREADS + threadIdx.x;
    const bool ok = q < m;
    const float *c = new_xyz + ((size_t)cloud * m + (ok ? q : 0)) * 3;
    const float qx = __ldg(c + 0), qy = __ldg(c + 1), qz = __ldg(c + 2);
    const float2 qx2 = make_float2(qx, qx), qy2 = make_float2(qy, qy), qz2 = make_float2(qz, qz);
    // thresholds: squared radius while the list is open, -1 once it is full (or for an out-of-range query / disabled list)
    float ta = ok ? __fmul_rn(radius_a, radius_a) : -1.0f;
    float tb = (ok && nsample_b > 0) ? __fmul_rn(radius_b, radius_b) : -1.0f;
    int ca = 0, cb = 0, fa = -1, fb = -1;
    int *oa = idx_a + ((size_t)cloud * m + (ok ? q : 0)) * nsample_a;
    int *ob = idx_b + ((size_t)cloud * m + (ok ? q : 0)) * nsample_b;
    const float ninf = __int_as_float(0xff800000);

    for (int base = 0; base < n; base += TILE_PTS) {
        const int tn = min(TILE_PTS, n - base);
        const int npairs = (((tn + 1) >> 1) + 3) & ~3;   // pairs, padded to the unroll factor
        if (base > 0) __syncthreads();  // everyone finished reading the previous tile
        for (int p = threadIdx.x; p < npairs; p += BQT_THREADS) {
            const int k0 = 2 * p, k1 = 2 * p + 1;
            const float *g = pts + (size_t)(base + k0) * 3;
            const float x0 = k0 < tn ? -__ldg(g + 0) : ninf, y0 = k0 < tn ? -__ldg(g + 1) : ninf, z0 = k0 < tn ? -__ldg(g + 2) : ninf;
            const float x1 = k1 < tn ? -__ldg(g + 3) : ninf, y1 = k1 < tn ? -__ldg(g + 4) : ninf, z1 = k1 < tn ? -__ldg(g + 5) : ninf;
            s_xy[p] = make_float4(x0, x1, y0, y1);
            s_z[p] = make_float2(z0, z1);
        }
        __syncthreads();
        for (int p0 = 0; p0 < npairs; p0 += 4) {
            if (__all_sync(0xffffffffu, fmaxf(ta, tb) < 0.0f)) break;   // every query of the warp has both lists full
            float2 d[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float4 xy = s_xy[p0 + u];
                const float2 z = s_z[p0 + u];
                const float2 dx = rt_fadd2(qx2, make_float2(xy.x, xy.y)), dy = rt_fadd2(qy2, make_float2(xy.z, xy.w)),
                             dz = rt_fadd2(qz2, z);
                d[u] = rt_ffma2(dz, dz, rt_ffma2(dx, dx, rt_fmul2(dy, dy)));   // rt_sqdist's order: dy*dy, +dx*dx, +dz*dz
            }
            const float tmax = fmaxf(ta, tb);
            const float dmin = fminf(fminf(fminf(d[0].x, d[0].y), fminf(d[1].x, d[1].y)), fminf(fminf(d[2].x, d[2].y), fminf(d[3].x, d[3].y)));
            if (dmin < tmax) {   // (NaN distances are never hits: fminf drops them, as `d < r2` does)
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int k = base + 2 * (p0 + u);
                    bqt_append(d[u].x, k, ta, tb, ca, cb, fa, fb, nsample_a, nsample_b, oa, ob);
                    bqt_append(d[u].y, k + 1, ta, tb, ca, cb, fa, fb, nsample_a, nsample_b, oa, ob);
                }
            }
        }
        if (__syncthreads_and(fmaxf(ta, tb) < 0.0f)) break;   // (uniform exit: the loop head has a CTA barrier)
    }
    // pad unused slots with the first hit; queries with no hit leave the caller's buffer untouched (the reference's
    // Python zero-fills it first) unless `zero_fill` asks this kernel to write the zeros itself (engine: no memset launch)
    if (!ok) return;
    if (fa >= 0 || zero_fill) {
        const int fill = fa >= 0 ? fa : 0;
        for (int s2 = ca; s2 < nsample_a; ++s2) oa[s2] = fill;
    }
    if (nsample_b > 0 && (fb >= 0 || zero_fill)) {
        const int fill = fb >= 0 ? fb : 0;
        for (int s2 = cb; s2 < nsample_b; ++s2) ob[s2] = fill;
    }
}

// ------------------------------------------------------------------------------------------
// three_nn: NN_G = 8 lanes per unknown point, NN_QPG = 2 points per lane group, known cloud in shared-memory tiles.
// Lane g of a group scans candidates g, g+8, ... (ascending, so its private best-3 list obeys the reference's strict-'<'
// cascade: among equal distances the lower index sits in the better slot); the eight lists are then merged by three
// rounds of "group minimum of (distance bits, index)", which is the same lexicographic order the reference's serial
// scan produces.  8x the warps of a thread-per-point scan: the per-candidate dependent chain is hidden by occupancy.
constexpr int NN_THREADS = 128;   // knn_kernel below: one thread per query
constexpr int NN3_THREADS = 256, NN_G = 8, NN_QPG = 2;
constexpr int NN3_QPB = NN3_THREADS / NN_G * NN_QPG;   // 64 unknown points per CTA

__global__ void __launch_bounds__(NN3_THREADS) three_nn_kernel(int n, int m, const float *__restrict__ unknown,
                                                               const float *__restrict__ known,
                                                               float *__restrict__ dist2, int *__restrict__ idx) {
    extern __shared__ __align__(16) float s_pts[];
    __shared__ __align__(8) uint64_t s_bar;
    const int cloud = blockIdx.y;
    const int g = threadIdx.x & (NN_G - 1);
    const int q0 = blockIdx.x * NN3_QPB + (threadIdx.x / NN_G) * NN_QPG;
    float ux[NN_QPG], uy[NN_QPG], uz[NN_QPG];
    const float inf = __int_as_float(0x7f800000);
    // the reference keeps double-typed bests initialised to 1e40: every finite fp32 distance is smaller, +inf/NaN never
    // are -- identical to fp32 bests initialised to +inf.
    float b1[NN_QPG], b2[NN_QPG], b3[NN_QPG];
    int i1[NN_QPG], i2[NN_QPG], i3[NN_QPG];
#pragma unroll
    for (int u = 0; u < NN_QPG; ++u) {
        const int j = min(q0 + u, n - 1);
        const float *p = unknown + ((size_t)cloud * n + j) * 3;
        ux[u] = __ldg(p + 0); uy[u] = __ldg(p + 1); uz[u] = __ldg(p + 2);
        b1[u] = b2[u] = b3[u] = inf;
        i1[u] = i2[u] = i3[u] = 0;
    }
    if (threadIdx.x == 0) {
        rt_mbar_init(&s_bar, 1);
        rt_fence_mbar_init();
    }
    __syncthreads();
    uint32_t parity = 0;
    for (int base = 0; base < m; base += TILE_PTS) {
        const int tn = min(TILE_PTS, m - base);
        if (base > 0) __syncthreads();
        rt_stage_floats(s_pts, known + ((size_t)cloud * m + base) * 3, tn * 3, &s_bar, parity);
        parity ^= 1;
#pragma unroll 2
        for (int k = g; k < tn; k += NN_G) {
            const float px = s_pts[k * 3 + 0], py = s_pts[k * 3 + 1], pz = s_pts[k * 3 + 2];
            const int kk = base + k;
#pragma unroll
            for (int u = 0; u < NN_QPG; ++u) {
                const float d = rt_sqdist(ux[u], uy[u], uz[u], px, py, pz);
                if (d < b3[u]) {
                    if (d < b1[u]) {
                        b3[u] = b2[u]; i3[u] = i2[u];
                        b2[u] = b1[u]; i2[u] = i1[u];
                        b1[u] = d; i1[u] = kk;
                    } else if (d < b2[u]) {
                        b3[u] = b2[u]; i3[u] = i2[u];
                        b2[u] = d; i2[u] = kk;
                    } else {
                        b3[u] = d; i3[u] = kk;
                    }
                }
            }
        }
    }
    // merge the eight private lists: d >= 0 and never NaN here, so the float bits order like the values
#pragma unroll
    for (int u = 0; u < NN_QPG; ++u) {
        float od[3];
        int oi[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const unsigned long long mine = ((unsigned long long)__float_as_uint(b1[u]) << 32) | (unsigned)i1[u];
            unsigned long long best = mine;
#pragma unroll
            for (int o = NN_G / 2; o >= 1; o >>= 1) {
                const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
                best = other < best ? other : best;
            }
            od[r] = __uint_as_float((unsigned)(best >> 32));
            oi[r] = (int)(unsigned)(best & 0xffffffffull);
            // exactly one lane owns a finite winner (indices are unique); sentinels (+inf, 0) tie harmlessly
            if (mine == best && b1[u] < inf) {
                b1[u] = b2[u]; i1[u] = i2[u];
                b2[u] = b3[u]; i2[u] = i3[u];
                b3[u] = inf; i3[u] = 0;
            }
        }
        if (g == 0 && q0 + u < n) {
            float *pd = dist2 + ((size_t)cloud * n + q0 + u) * 3;
            int *pi = idx + ((size_t)cloud * n + q0 + u) * 3;
            pd[0] = od[0]; pd[1] = od[1]; pd[2] = od[2];
            pi[0] = oi[0]; pi[1] = oi[1]; pi[2] = oi[2];
        }
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
inline size_t tile_bytes_pad32(int n) { return (size_t)((min(n, TILE_PTS) + 31) & ~31) * 3 * sizeof(float); }

}  // namespace

// RT_BQ_WARP=1 selects the round-1 warp-per-query kernel (A/B timing)
static bool bq_use_thread_kernel() {
    static int v = -1;
    if (v < 0) {
        const char *env = getenv("RT_BQ_WARP");
        v = (env && atoi(env) == 1) ? 0 : 1;
    }
    return v == 1;
}

// replaces ball_query_wrapper_fast (reference: src/lib/src/ball_query.cpp:18-29)
RT_API int rt_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xyz, const float *xyz,
                         int *idx, void *stream) {
    RT_REQUIRE(b >= 0 && n >= 0 && m >= 0 && nsample >= 0 && new_xyz && xyz && idx, "ball_query: bad arguments");
    if (b == 0 || m == 0 || nsample == 0 || n == 0) return RT_OK;
    RT_REQUIRE(b <= 65535, "ball_query: batch > 65535");
    if (bq_use_thread_kernel()) {
        dim3 grid(rt_divup(m, BQT_THREADS), b);
        ball_query_thread_kernel<<<grid, BQT_THREADS, 0, (cudaStream_t)stream>>>(n, m, radius, nsample, idx, 0.0f, 0, idx, new_xyz, xyz, 0,
                                                                                  nullptr);
        return rt_check_launch("ball_query_thread_kernel");
    }
    dim3 grid(rt_divup(m, BQ_QPB), b);
    ball_query_kernel<<<grid, BQ_THREADS, tile_bytes_pad32(n), (cudaStream_t)stream>>>(n, m, radius, nsample, idx, 0.0f, 0, idx,
                                                                                 new_xyz, xyz, 0, nullptr);
    return rt_check_launch("ball_query_kernel");
}

// engine-internal: two radii over the same centres in one launch (same results as two rt_ball_query calls)
int rt_launch_ball_query2(int b, int n, int m, float radius_a, int nsample_a, int *idx_a, float radius_b, int nsample_b,
                          int *idx_b, const float *new_xyz, const float *xyz, int zero_fill, cudaStream_t st, const int *counts) {
    if (b == 0 || m == 0 || n == 0) return RT_OK;
    RT_REQUIRE(b <= 65535 && nsample_a > 0 && nsample_b > 0, "ball_query2: bad arguments");
    if (bq_use_thread_kernel()) {
        dim3 grid(rt_divup(m, BQT_THREADS), b);
        ball_query_thread_kernel<<<grid, BQT_THREADS, 0, st>>>(n, m, radius_a, nsample_a, idx_a, radius_b, nsample_b, idx_b, new_xyz, xyz,
                                                                zero_fill, counts);
        return rt_check_launch("ball_query_thread_kernel(2 radii)");
    }
    dim3 grid(rt_divup(m, BQ_QPB), b);
    ball_query_kernel<<<grid, BQ_THREADS, tile_bytes_pad32(n), st>>>(n, m, radius_a, nsample_a, idx_a, radius_b, nsample_b, idx_b,
                                                              new_xyz, xyz, zero_fill, counts);
    return rt_check_launch("ball_query_kernel(2 radii)");
}

// replaces three_nn_wrapper_fast (reference: src/lib/src/interpolate.cpp:16-25)
RT_API int rt_three_nn(int b, int n, int m, const float *unknown, const float *known, float *dist2, int *idx,
                       void *stream) {
    RT_REQUIRE(b >= 0 && n >= 0 && m >= 0 && unknown && known && dist2 && idx, "three_nn: bad arguments");
    if (b == 0 || n == 0) return RT_OK;
    RT_REQUIRE(b <= 65535, "three_nn: batch > 65535");
    dim3 grid(rt_divup(n, NN3_QPB), b);
    three_nn_kernel<<<grid, NN3_THREADS, tile_bytes(m), (cudaStream_t)stream>>>(n, m, unknown, known, dist2, idx);
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
