// DBSCAN labels on the device.  The reference clusters the moving points of every frame with
// scikit-learn on the HOST (`self.dbscan.fit_predict(f_detached)`, src/models/track4d.py:36,108-126: a device->host copy, a
// CPU tree query and a sync per frame).  This kernel reproduces sklearn's labels exactly:
//   core[i]   = #{j : |x_i - x_j| <= eps} >= min_samples            (the point itself counts)
//   clusters  = connected components of the core points under the eps graph, numbered in the order of their smallest
//               core index (sklearn's dbscan_inner walks the points in index order and fully expands one cluster before
//               it opens the next);
//   border    = a non-core point within eps of core points joins the lowest-numbered of their clusters (the first one to
//               reach it in sklearn's traversal);  noise = -1.
// One CTA per set; the n x n adjacency lives in shared memory as bit rows (n <= 1024: 128 KB) -- or, for larger sets (the
// reference has no limit; config-5 frames of ~3000 points can have more than 1024 moving points), in a stream-ordered global
// scratch with the points read through L1 -- components come from min-label propagation with pointer jumping.  Distances are accumulated in fp64 from the fp32 inputs, like sklearn's trees.
#include "common.cuh"

namespace {

constexpr int DB_THREADS = 1024;
constexpr int DB_MAX_N_SMEM = 1024;
constexpr int DB_MAX_N = 16384;
constexpr int DB_MAX_D = 16;

__global__ void __launch_bounds__(DB_THREADS) dbscan_kernel(int n, int d, const float *__restrict__ x_all, double eps2, int min_samples,
                                                            int *__restrict__ labels_all, uint32_t *__restrict__ adj_scratch) {
    extern __shared__ __align__(16) unsigned char db_smem[];
    const int words = (n + 31) >> 5;
    const bool big = adj_scratch != nullptr;
    uint32_t *adj = big ? adj_scratch + (size_t)blockIdx.x * n * words : reinterpret_cast<uint32_t *>(db_smem);   // n x words
    float *xs_s = reinterpret_cast<float *>(db_smem + (big ? 0 : sizeof(uint32_t) * (size_t)n * words));           // n x d (small sets)
    int *label = reinterpret_cast<int *>(xs_s + (big ? 0 : (size_t)n * d));   // n: smallest core index of the component (cores only)
    int *cid = label + n;                                            // n: cluster number of a core point
    uint32_t *corem = reinterpret_cast<uint32_t *>(cid + n);         // words: core bit mask
    uint32_t *rootm = corem + words;                                 // words: component-root bit mask
    __shared__ int s_changed;
    const int t = threadIdx.x;
    const float *x = x_all + (size_t)blockIdx.x * n * d;
    int *labels = labels_all + (size_t)blockIdx.x * n;

    const float *xs = big ? x : xs_s;
    if (!big)
        for (int i = t; i < n * d; i += DB_THREADS) xs_s[i] = x[i];
    for (int i = t; i < words; i += DB_THREADS) { corem[i] = 0u; rootm[i] = 0u; }
    __syncthreads();
    // adjacency rows + neighbour counts: one row per thread, so every candidate read is a shared-memory broadcast
    // (a (row, word)-per-thread split was tried for occupancy at small n: bank conflicts / extra row reloads made it slower)
    for (int i = t; i < n; i += DB_THREADS) {
        float xi[DB_MAX_D];
#pragma unroll
        for (int c = 0; c < DB_MAX_D; ++c) xi[c] = c < d ? xs[i * d + c] : 0.0f;
        int count = 0;
        for (int w = 0; w < words; ++w) {
            uint32_t bits = 0u;
            const int jend = min(32, n - 32 * w);
            for (int b = 0; b < jend; ++b) {
                const float *xj = xs + (size_t)(32 * w + b) * d;
                double acc = 0.0;
#pragma unroll
                for (int c = 0; c < DB_MAX_D; ++c)
                    if (c < d) {
                        const double df = (double)xi[c] - (double)xj[c];
                        acc += df * df;
                    }
                if (acc <= eps2) bits |= 1u << b;
            }
            adj[(size_t)i * words + w] = bits;
            count += __popc(bits);
        }
        const bool core = count >= min_samples;
        label[i] = core ? i : 0x7fffffff;
        if (core) atomicOr(&corem[i >> 5], 1u << (i & 31));
    }
    __syncthreads();
    // connected components of the core graph: label = smallest core index reachable
    for (int iter = 0; iter < 4 * n + 8; ++iter) {
        if (t == 0) s_changed = 0;
        __syncthreads();
        for (int i = t; i < n; i += DB_THREADS) {
            if (!((corem[i >> 5] >> (i & 31)) & 1u)) continue;
            int best = label[i];
            for (int w = 0; w < words; ++w) {
                uint32_t bits = adj[(size_t)i * words + w] & corem[w];
                while (bits) {
                    const int j = 32 * w + __ffs(bits) - 1;
                    bits &= bits - 1;
                    best = min(best, label[j]);
                }
            }
            best = min(best, label[best]);          // pointer jumping (labels only ever decrease: races are benign)
            if (best < label[i]) {
                label[i] = best;
                s_changed = 1;
            }
        }
        __syncthreads();
        if (!s_changed) break;
        __syncthreads();
    }
    // number the components by their smallest core index
    for (int i = t; i < n; i += DB_THREADS)
        if (label[i] == i) atomicOr(&rootm[i >> 5], 1u << (i & 31));
    __syncthreads();
    for (int i = t; i < n; i += DB_THREADS) {
        int c = -1;
        if (label[i] != 0x7fffffff) {
            const int r = label[i];
            c = 0;
            for (int w = 0; w < (r >> 5); ++w) c += __popc(rootm[w]);
            c += __popc(rootm[r >> 5] & ((1u << (r & 31)) - 1u));
        }
        cid[i] = c;
    }
    __syncthreads();
    // border points: lowest cluster number among the adjacent core points
    for (int i = t; i < n; i += DB_THREADS) {
        int c = cid[i];
        if (c < 0) {
            int best = 0x7fffffff;
            for (int w = 0; w < words; ++w) {
                uint32_t bits = adj[(size_t)i * words + w] & corem[w];
                while (bits) {
                    const int j = 32 * w + __ffs(bits) - 1;
                    bits &= bits - 1;
                    best = min(best, cid[j]);
                }
            }
            c = best == 0x7fffffff ? -1 : best;
        }
        labels[i] = c;
    }
}

}  // namespace

// C ABI.  x (b,n,d) fp32 -> labels (b,n) int32, identical to sklearn.cluster.DBSCAN(eps, min_samples).fit_predict on every
// (n,d) set (the reference: Track4D.clustering, src/models/track4d.py:108-126, eps 1.5, min_samples = min_obj_points).
RT_API int rt_dbscan(int b, int n, int d, const float *x, float eps, int min_samples, int *labels, void *stream) {
    RT_REQUIRE(b >= 0 && n >= 0 && d >= 1 && x && labels && min_samples >= 1 && eps >= 0.0f, "dbscan: bad arguments");
    RT_REQUIRE(n <= DB_MAX_N && d <= DB_MAX_D, "dbscan: at most %d points of %d dimensions per set", DB_MAX_N, DB_MAX_D);
    if (b == 0 || n == 0) return RT_OK;
    const int words = (n + 31) >> 5;
    const bool big = n > DB_MAX_N_SMEM;
    const size_t adj_bytes = sizeof(uint32_t) * (size_t)n * words;
    const size_t smem = (big ? 0 : adj_bytes + sizeof(float) * (size_t)n * d) + sizeof(uint32_t) * 2 * (size_t)words + sizeof(int) * 2 * (size_t)n;
    static RtPerDevice attr;
    const int dev = rt_current_device();
    if (!attr.done(dev)) {
        const cudaError_t e = cudaFuncSetAttribute(dbscan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        if (e != cudaSuccess) {
            rt_set_error("dbscan: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return (int)e;
        }
        attr.mark(dev);
    }
    RT_REQUIRE(smem <= 220 * 1024, "dbscan: %zu bytes of shared memory", smem);
    cudaStream_t st = (cudaStream_t)stream;
    uint32_t *scratch = nullptr;
    if (big) {
        const int ae = rt_scratch_alloc((void **)&scratch, adj_bytes * (size_t)b, st, "dbscan");
        if (ae != RT_OK) return ae;
    }
    dbscan_kernel<<<b, DB_THREADS, smem, st>>>(n, d, x, (double)eps * (double)eps, min_samples, labels, scratch);
    const int rc = rt_check_launch("dbscan_kernel");
    rt_scratch_free(scratch, st);
    return rc;
}
