// Sinkhorn association on the device: the reference's `log_optimal_transport` (500 log-space Sinkhorn iterations on the
// (m+1) x (n+1) coupling matrix, src/models/utils/track4d_utils.py:405-434) followed by the mutual-arg-max matching of
// `Track4D.sinkhorn_module` (src/models/track4d.py:166-180).  SURVEY.md section 8(f) row 1.
//
// The reference runs this as ~2000 tiny torch kernels per frame (two logsumexp + broadcasts per iteration).  Here ONE CTA
// per batch element keeps the coupling matrix, u and v in shared memory for all iterations: a warp owns a row (u update)
// or a column (v update; a transposed copy makes that a row walk too), the logsumexp is a shuffle max + shuffle sum.
#include "common.cuh"

namespace {

constexpr int SK_THREADS = 256;
constexpr int SK_MAX_SMEM = 128;   // (m+1), (n+1) <= SK_MAX_SMEM: coupling matrix and its transpose live in shared memory
constexpr int SK_MAX = 2048;       // beyond that, up to SK_MAX, they live in a stream-ordered global scratch (L1/L2-resident)

// log(sum_k exp(row[k] + add[k])), k < len  (torch.logsumexp: max-shifted; -inf rows stay -inf), evaluated by the LPR
// consecutive lanes that share the row (LPR = 8 for small matrices: four rows per warp step, 3-step shuffles)
template <int LPR>
__device__ __forceinline__ float sk_lse(const float *row, const float *add, int len, int sub) {
    float mx = -INFINITY;
    for (int k = sub; k < len; k += LPR) mx = fmaxf(mx, row[k] + add[k]);
#pragma unroll
    for (int o = LPR / 2; o >= 1; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    const float shift = (mx == -INFINITY || mx == INFINITY) ? 0.0f : mx;
    float sum = 0.0f;
    for (int k = sub; k < len; k += LPR) sum += expf(row[k] + add[k] - shift);
#pragma unroll
    for (int o = LPR / 2; o >= 1; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    return logf(sum) + shift;
}

// THREADS: 256, or 64 for matrices of up to 8 x 8 (two warps hold all rows: the two CTA barriers of an iteration -- the
// kernel's critical path at the reference's few-objects-per-frame operating point -- synchronise 2 warps instead of 8)
template <int LPR, int THREADS>
__global__ void __launch_bounds__(THREADS) sinkhorn_match_kernel(int m, int n, const float *__restrict__ aff_all, float alpha,
                                                                     int iters, float *__restrict__ scores_all,
                                                                     long long *__restrict__ idx0_all, long long *__restrict__ idx1_all,
                                                                     float *__restrict__ zscratch) {
    extern __shared__ float sm[];
    const int M = m + 1, N = n + 1, ld = N + 1, ldt = M + 1;   // +1: rows start on different banks
    const size_t zfloats = (size_t)M * ld + (size_t)N * ldt;
    float *Z = zscratch ? zscratch + blockIdx.x * zfloats : sm;   // M x ld
    float *ZT = Z + (size_t)M * ld;                               // N x ldt (transposed copy)
    float *u = zscratch ? sm : ZT + (size_t)N * ldt;              // M
    float *v = u + M;                 // N
    float *log_mu = v + N;            // M
    float *log_nu = log_mu + M;       // N
    float *max0v = log_nu + N;        // m   row maxima of the final scores (without the dustbins)
    float *max1v = max0v + M;         // n
    int *max0i = reinterpret_cast<int *>(max1v + N);
    int *max1i = max0i + M;
    const int b = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5, nw = THREADS / 32;
    const float *aff = aff_all + (size_t)b * m * n;

    // couplings = [[scores, alpha], [alpha, alpha]]; log_mu / log_nu as in log_optimal_transport
    for (int e = t; e < M * N; e += THREADS) {
        const int i = e / N, j = e % N;
        const float z = (i < m && j < n) ? aff[i * n + j] : alpha;
        Z[i * ld + j] = z;
        ZT[j * ldt + i] = z;
    }
    const float norm = -logf((float)m + (float)n);
    for (int i = t; i < M; i += THREADS) {
        u[i] = 0.0f;
        log_mu[i] = i < m ? norm : logf((float)n) + norm;
    }
    for (int j = t; j < N; j += THREADS) {
        v[j] = 0.0f;
        log_nu[j] = j < n ? norm : logf((float)m) + norm;
    }
    __syncthreads();
    // (a single-warp variant for <= 32 objects -- rows in registers, u / v by shuffles, no CTA barrier -- was measured slower:
    //  1.7 vs 0.68 ms at 20 x 20; the serial exp/log chain of one warp costs more than the two barriers it removes)
    {
    constexpr int RPW = 32 / LPR;                       // rows per warp step
    const int sub = lane % LPR, rsel = lane / LPR;
    const int rows_per_step = nw * RPW;
    for (int it = 0; it < iters; ++it) {
        for (int i0 = 0; i0 < M; i0 += rows_per_step) {   // trip count is CTA-uniform: every lane reaches the shuffles
            const int i = i0 + warp * RPW + rsel;
            const int ic = i < M ? i : M - 1;
            const float l = sk_lse<LPR>(Z + ic * ld, v, N, sub);
            if (sub == 0 && i < M) u[i] = log_mu[i] - l;
        }
        __syncthreads();
        for (int j0 = 0; j0 < N; j0 += rows_per_step) {
            const int j = j0 + warp * RPW + rsel;
            const int jc = j < N ? j : N - 1;
            const float l = sk_lse<LPR>(ZT + jc * ldt, u, M, sub);
            if (sub == 0 && j < N) v[j] = log_nu[j] - l;
        }
        __syncthreads();
    }
    }
    // Z + u + v - norm
    for (int e = t; e < M * N; e += THREADS) {
        const int i = e / N, j = e % N;
        const float s = Z[i * ld + j] + u[i] + v[j] - norm;
        Z[i * ld + j] = s;
        ZT[j * ldt + i] = s;
        if (scores_all) scores_all[(size_t)b * M * N + e] = s;
    }
    __syncthreads();
    // arg-max over the real rows / columns (first maximum wins, as torch.max does)
    for (int i = warp; i < m; i += nw) {
        float bv = -INFINITY;
        int bi = 0x7fffffff;
        for (int j = lane; j < n; j += 32) {
            const float s = Z[i * ld + j];
            if (s > bv || (s == bv && j < bi)) { bv = s; bi = j; }
        }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) { max0v[i] = bv; max0i[i] = bi == 0x7fffffff ? 0 : bi; }
    }
    for (int j = warp; j < n; j += nw) {
        float bv = -INFINITY;
        int bi = 0x7fffffff;
        for (int i = lane; i < m; i += 32) {
            const float s = ZT[j * ldt + i];
            if (s > bv || (s == bv && i < bi)) { bv = s; bi = i; }
        }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) { max1v[j] = bv; max1i[j] = bi == 0x7fffffff ? 0 : bi; }
    }
    __syncthreads();
    // mutual check and matching threshold (track4d.py:170-180)
    for (int i = t; i < m; i += THREADS) {
        const bool mutual0 = max1i[max0i[i]] == i;
        const bool valid0 = mutual0 && expf(max0v[i]) > 0.0f;
        if (idx0_all) idx0_all[(size_t)b * m + i] = valid0 ? max0i[i] : -1;
    }
    for (int j = t; j < n; j += THREADS) {
        const int i = max1i[j];
        const bool mutual1 = max0i[i] == j;
        const bool mutual0 = max1i[max0i[i]] == i;
        const bool valid0 = mutual0 && expf(max0v[i]) > 0.0f;
        idx1_all[(size_t)b * n + j] = (mutual1 && valid0) ? i : -1;
    }
}

}  // namespace

// C ABI.  aff (b,m,n) fp32 affinities -> scores (b,m+1,n+1) [optional], indices0 (b,m) [optional], indices1 (b,n) int64
// (index of the matched previous object, -1 = none): log_optimal_transport(aff, alpha, iters) + the mutual matching of
// Track4D.sinkhorn_module (reference: src/models/utils/track4d_utils.py:405-434, src/models/track4d.py:166-180).
RT_API int rt_sinkhorn_match(int b, int m, int n, const float *aff, float alpha, int iters, float *scores, long long *indices0,
                             long long *indices1, void *stream) {
    RT_REQUIRE(b >= 0 && m >= 1 && n >= 1 && aff && indices1 && iters >= 0, "sinkhorn_match: bad arguments (b=%d m=%d n=%d)", b, m, n);
    RT_REQUIRE(m + 1 <= SK_MAX && n + 1 <= SK_MAX, "sinkhorn_match: at most %d x %d objects", SK_MAX - 1, SK_MAX - 1);
    if (b == 0) return RT_OK;
    const int M = m + 1, N = n + 1;
    const bool in_smem = M <= SK_MAX_SMEM && N <= SK_MAX_SMEM;
    const size_t zbytes = sizeof(float) * ((size_t)M * (N + 1) + (size_t)N * (M + 1));
    const size_t vbytes = sizeof(float) * 4 * (size_t)(M + N) + sizeof(int) * (size_t)(M + N);
    const size_t smem = vbytes + (in_smem ? zbytes : 0);
    static RtPerDevice attr;
    const int dev = rt_current_device();
    if (!attr.done(dev)) {
        cudaError_t e = cudaFuncSetAttribute(sinkhorn_match_kernel<8, SK_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(sinkhorn_match_kernel<32, SK_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        if (e != cudaSuccess) {
            rt_set_error("sinkhorn: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return (int)e;
        }
        attr.mark(dev);
    }
    cudaStream_t st = (cudaStream_t)stream;
    float *zscratch = nullptr;
    if (!in_smem) {   // more than 127 objects on a side: the reference has no limit (track4d.py:166-180), so neither has this entry
        const int ae = rt_scratch_alloc((void **)&zscratch, zbytes * (size_t)b, st, "sinkhorn_match");
        if (ae != RT_OK) return ae;
    }
    if (M <= 8 && N <= 8)
        sinkhorn_match_kernel<8, 64><<<b, 64, smem, st>>>(m, n, aff, alpha, iters, scores, indices0, indices1, zscratch);
    else if (M <= 96 && N <= 96)
        sinkhorn_match_kernel<8, SK_THREADS><<<b, SK_THREADS, smem, st>>>(m, n, aff, alpha, iters, scores, indices0, indices1, zscratch);
    else
        sinkhorn_match_kernel<32, SK_THREADS><<<b, SK_THREADS, smem, st>>>(m, n, aff, alpha, iters, scores, indices0, indices1, zscratch);
    const int rc = rt_check_launch("sinkhorn_match_kernel");
    rt_scratch_free(zscratch, st);
    return rc;
}
