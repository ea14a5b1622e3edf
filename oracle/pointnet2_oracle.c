/*
 * pointnet2_oracle.c -- CPU restatement of RaTrack's native pointnet2 ops.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library.  The product
 * (ratrack_b200/) never links, imports or falls back to it.
 *
 * The reference has no CPU implementation of these ops (they exist only as CUDA
 * kernels under /root/reference/src/lib/src), so this file restates each kernel's
 * algorithm as plain C, including the floating-point evaluation order the
 * reference's sm_100 build uses (checked with cuobjdump -sass on oracle/_ref):
 *
 *     d2 = fmaf(dz, dz, fmaf(dx, dx, dy * dy))            all four distance kernels
 *     out = fmaf(w2, p2, fmaf(w0, p0, w1 * p1))           three_interpolate
 *
 * Compile with -ffp-contract=off so the compiler adds no contraction of its own.
 * Pinning: tests/golden/ref_gpu_*.npz hold outputs of the reference's own kernels
 * run on a B200 (oracle/gen_golden_ref_gpu.py); tests/test_oracle_golden.py checks
 * this file against them bit for bit.
 *
 * Layouts are the reference's: row-major contiguous fp32 / int32.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* reference: lib/src/ball_query_gpu.cu:33, interpolate_gpu.cu:40,108, sampling_gpu.cu:133 */
static inline float sqdist3(float ax, float ay, float az, float bx, float by, float bz) {
    float dx = ax - bx, dy = ay - by, dz = az - bz;
    float t = dy * dy;
    t = fmaf(dx, dx, t);
    return fmaf(dz, dz, t);
}

/* reference: lib/src/cuda_utils.h:10-14 (largest power of two <= n, capped at 1024) */
ORC_API int orc_opt_n_threads(int work_size) {
    int pow_2 = (int)(log((double)work_size) / log(2.0));
    int v = 1 << pow_2;
    if (v > 1024) v = 1024;
    if (v < 1) v = 1;
    return v;
}

/*
 * Furthest point sampling.  reference: lib/src/sampling_gpu.cu:94-209 (kernel),
 * :86-91 (__update), :211-253 (block size choice).  Literal simulation of one CTA
 * of `bs` threads: per-thread strided scan keeping the first strict maximum, then
 * the shared-memory tree where slot p takes slot p+s only when strictly greater.
 * temp is caller-initialised (1e10 in lib/pointnet2_utils.py:26) and updated in place.
 */
ORC_API void orc_furthest_point_sampling(int b, int n, int m, const float *xyz, float *temp, int *idx) {
    if (m <= 0) return;
    const int bs = orc_opt_n_threads(n);
#pragma omp parallel for schedule(dynamic, 1)
    for (int bi = 0; bi < b; ++bi) {
        const float *p = xyz + (size_t)bi * n * 3;
        float *tmp = temp + (size_t)bi * n;
        int *out = idx + (size_t)bi * m;
        float *dists = (float *)malloc(sizeof(float) * bs);
        int *dists_i = (int *)malloc(sizeof(int) * bs);
        int old = 0;
        out[0] = 0;
        for (int j = 1; j < m; ++j) {
            const float x1 = p[old * 3 + 0], y1 = p[old * 3 + 1], z1 = p[old * 3 + 2];
            for (int tid = 0; tid < bs; ++tid) {
                int besti = 0;
                float best = -1.0f;
                for (int k = tid; k < n; k += bs) {
                    float d = sqdist3(p[k * 3 + 0], p[k * 3 + 1], p[k * 3 + 2], x1, y1, z1);
                    float d2 = fminf(d, tmp[k]);
                    tmp[k] = d2;
                    if (d2 > best) { besti = k; best = d2; }
                }
                dists[tid] = best;
                dists_i[tid] = besti;
            }
            for (int s = bs / 2; s >= 1; s >>= 1) {
                for (int t = 0; t < s; ++t) {
                    float v1 = dists[t], v2 = dists[t + s];
                    int i1 = dists_i[t], i2 = dists_i[t + s];
                    dists[t] = fmaxf(v1, v2);
                    dists_i[t] = (v2 > v1) ? i2 : i1;
                }
            }
            old = dists_i[0];
            out[j] = old;
        }
        free(dists);
        free(dists_i);
    }
}

/* reference: lib/src/sampling_gpu.cu:8-24.  out[b,c,j] = points[b,c,idx[b,j]] */
ORC_API void orc_gather_points(int b, int c, int n, int m, const float *points, const int *idx, float *out) {
#pragma omp parallel for collapse(2)
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci) {
            const float *src = points + ((size_t)bi * c + ci) * n;
            const int *ix = idx + (size_t)bi * m;
            float *dst = out + ((size_t)bi * c + ci) * m;
            for (int j = 0; j < m; ++j) dst[j] = src[ix[j]];
        }
}

/* reference: lib/src/sampling_gpu.cu:46-63 (atomicAdd scatter; here in ascending j order) */
ORC_API void orc_gather_points_grad(int b, int c, int n, int m, const float *grad_out, const int *idx,
                                    float *grad_points) {
#pragma omp parallel for collapse(2)
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci) {
            const float *g = grad_out + ((size_t)bi * c + ci) * m;
            const int *ix = idx + (size_t)bi * m;
            float *dst = grad_points + ((size_t)bi * c + ci) * n;
            for (int j = 0; j < m; ++j) dst[ix[j]] += g[j];
        }
}

/*
 * reference: lib/src/ball_query_gpu.cu:9-45.  First `nsample` points (index order)
 * with d2 < radius*radius (strict, fp32); the first hit fills every slot; untouched
 * slots keep whatever the caller put there (zeros: lib/pointnet2_utils.py:246).
 */
ORC_API void orc_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xyz,
                            const float *xyz, int *idx) {
    const float radius2 = radius * radius;
#pragma omp parallel for collapse(2) schedule(static)
    for (int bi = 0; bi < b; ++bi)
        for (int j = 0; j < m; ++j) {
            const float *q = new_xyz + ((size_t)bi * m + j) * 3;
            const float *p = xyz + (size_t)bi * n * 3;
            int *o = idx + ((size_t)bi * m + j) * nsample;
            int cnt = 0;
            for (int k = 0; k < n; ++k) {
                float d2 = sqdist3(q[0], q[1], q[2], p[k * 3 + 0], p[k * 3 + 1], p[k * 3 + 2]);
                if (d2 < radius2) {
                    if (cnt == 0)
                        for (int l = 0; l < nsample; ++l) o[l] = k;
                    o[cnt] = k;
                    ++cnt;
                    if (cnt >= nsample) break;
                }
            }
        }
}

/* reference: lib/src/group_points_gpu.cu:47-66.  out[b,c,p,s] = points[b,c,idx[b,p,s]] */
ORC_API void orc_group_points(int b, int c, int n, int npoints, int nsample, const float *points,
                              const int *idx, float *out) {
    const size_t ps = (size_t)npoints * nsample;
#pragma omp parallel for collapse(2)
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci) {
            const float *src = points + ((size_t)bi * c + ci) * n;
            const int *ix = idx + (size_t)bi * ps;
            float *dst = out + ((size_t)bi * c + ci) * ps;
            for (size_t e = 0; e < ps; ++e) dst[e] = src[ix[e]];
        }
}

/* reference: lib/src/group_points_gpu.cu:8-25 (atomicAdd scatter; here ascending element order) */
ORC_API void orc_group_points_grad(int b, int c, int n, int npoints, int nsample, const float *grad_out,
                                   const int *idx, float *grad_points) {
    const size_t ps = (size_t)npoints * nsample;
#pragma omp parallel for collapse(2)
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci) {
            const float *g = grad_out + ((size_t)bi * c + ci) * ps;
            const int *ix = idx + (size_t)bi * ps;
            float *dst = grad_points + ((size_t)bi * c + ci) * n;
            for (size_t e = 0; e < ps; ++e) dst[ix[e]] += g[e];
        }
}

/*
 * reference: lib/src/interpolate_gpu.cu:81-124.  Three smallest fp32 d2 kept in
 * double-typed slots initialised to 1e40, strict '<' cascade, results narrowed to float.
 */
ORC_API void orc_three_nn(int b, int n, int m, const float *unknown, const float *known, float *dist2,
                          int *idx) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int bi = 0; bi < b; ++bi)
        for (int j = 0; j < n; ++j) {
            const float *u = unknown + ((size_t)bi * n + j) * 3;
            const float *kn = known + (size_t)bi * m * 3;
            double best1 = 1e40, best2 = 1e40, best3 = 1e40;
            int i1 = 0, i2 = 0, i3 = 0;
            for (int k = 0; k < m; ++k) {
                float d = sqdist3(u[0], u[1], u[2], kn[k * 3 + 0], kn[k * 3 + 1], kn[k * 3 + 2]);
                if (d < best1) {
                    best3 = best2; i3 = i2;
                    best2 = best1; i2 = i1;
                    best1 = d; i1 = k;
                } else if (d < best2) {
                    best3 = best2; i3 = i2;
                    best2 = d; i2 = k;
                } else if (d < best3) {
                    best3 = d; i3 = k;
                }
            }
            float *od = dist2 + ((size_t)bi * n + j) * 3;
            int *oi = idx + ((size_t)bi * n + j) * 3;
            od[0] = (float)best1; od[1] = (float)best2; od[2] = (float)best3;
            oi[0] = i1; oi[1] = i2; oi[2] = i3;
        }
}

/*
 * reference: lib/src/interpolate_gpu.cu:9-57.  Sorted insertion (strict '<') into
 * k double-typed slots initialised to 1e40 / index 0; the reference caps k at 200.
 */
ORC_API int orc_knn(int b, int n, int m, int k, const float *unknown, const float *known, float *dist2,
                    int *idx) {
    if (k > 200 || k < 0) return -1;
#pragma omp parallel for collapse(2) schedule(static)
    for (int bi = 0; bi < b; ++bi)
        for (int j = 0; j < n; ++j) {
            const float *u = unknown + ((size_t)bi * n + j) * 3;
            const float *kn = known + (size_t)bi * m * 3;
            double best[200];
            int besti[200];
            for (int i = 0; i < k; ++i) { best[i] = 1e40; besti[i] = 0; }
            for (int i = 0; i < m; ++i) {
                float d = sqdist3(u[0], u[1], u[2], kn[i * 3 + 0], kn[i * 3 + 1], kn[i * 3 + 2]);
                for (int s = 0; s < k; ++s) {
                    if (d < best[s]) {
                        for (int l = k - 1; l > s; --l) { best[l] = best[l - 1]; besti[l] = besti[l - 1]; }
                        best[s] = d;
                        besti[s] = i;
                        break;
                    }
                }
            }
            float *od = dist2 + ((size_t)bi * n + j) * k;
            int *oi = idx + ((size_t)bi * n + j) * k;
            for (int i = 0; i < k; ++i) { oi[i] = besti[i]; od[i] = (float)best[i]; }
        }
    return 0;
}

/* reference: lib/src/interpolate_gpu.cu:149-169; sm_100 SASS order: w1*p1, fma(w0,p0,.), fma(w2,p2,.) */
ORC_API void orc_three_interpolate(int b, int c, int m, int n, const float *points, const int *idx,
                                   const float *weight, float *out) {
#pragma omp parallel for collapse(2)
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci) {
            const float *src = points + ((size_t)bi * c + ci) * m;
            float *dst = out + ((size_t)bi * c + ci) * n;
            for (int j = 0; j < n; ++j) {
                const int *ix = idx + ((size_t)bi * n + j) * 3;
                const float *w = weight + ((size_t)bi * n + j) * 3;
                float t = w[1] * src[ix[1]];
                t = fmaf(w[0], src[ix[0]], t);
                dst[j] = fmaf(w[2], src[ix[2]], t);
            }
        }
}

/* reference: lib/src/interpolate_gpu.cu:192-214 (3 atomicAdds per element; here ascending j, tap 0,1,2) */
ORC_API void orc_three_interpolate_grad(int b, int c, int n, int m, const float *grad_out, const int *idx,
                                        const float *weight, float *grad_points) {
#pragma omp parallel for collapse(2)
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci) {
            const float *g = grad_out + ((size_t)bi * c + ci) * n;
            float *dst = grad_points + ((size_t)bi * c + ci) * m;
            for (int j = 0; j < n; ++j) {
                const int *ix = idx + ((size_t)bi * n + j) * 3;
                const float *w = weight + ((size_t)bi * n + j) * 3;
                dst[ix[0]] += g[j] * w[0];
                dst[ix[1]] += g[j] * w[1];
                dst[ix[2]] += g[j] * w[2];
            }
        }
}
