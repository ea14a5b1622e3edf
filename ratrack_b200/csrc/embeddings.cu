// Object embeddings of the association step in ONE launch (reference: Track4D.affinity_module, src/models/track4d.py:202-216,
// which builds them per object with ~15 torch calls: mean / var of the position, max of the 128 features, mean flow, mean and
// var of the radial velocity).  Input: the feature columns of all objects side by side, (139, P) channel-major, object i owning
// the columns [off[i], off[i+1]).  One CTA per object, one thread per channel: a channel's values of an object are contiguous.
#include "common.cuh"

namespace {

constexpr int EMB_CH = 139, EMB_OUT = 141;

__global__ void __launch_bounds__(160) object_embeddings_kernel(int p_total, const float *__restrict__ points, const int *__restrict__ off,
                                                                float *__restrict__ out) {
    const int obj = blockIdx.x, ch = threadIdx.x + 3;      // channels 0..2 (the warped position) do not enter the embedding
    if (ch >= EMB_CH) return;
    const int c0 = __ldg(off + obj), c1 = __ldg(off + obj + 1), n = c1 - c0;
    const float *x = points + (size_t)ch * p_total;
    float *o = out + (size_t)obj * EMB_OUT;
    if (ch >= 11) {                                        // features 11..138 -> max
        float m = -INFINITY;
        for (int j = c0; j < c1; ++j) m = fmaxf(m, __ldg(x + j));
        o[6 + (ch - 11)] = m;
        return;
    }
    float s = 0.0f;
    for (int j = c0; j < c1; ++j) s += __ldg(x + j);
    const float mean = s / (float)n;
    const bool pos = ch < 6, rrv = ch >= 9;
    if (pos) o[ch - 3] = mean;                             // 0..2 centre
    else if (!rrv) o[134 + (ch - 6)] = mean;               // 134..136 mean flow
    else o[137 + (ch - 9)] = mean;                         // 137..138 mean radial velocity
    if (pos || rrv) {                                      // population variance, two-pass (as torch.var(unbiased=False))
        float v = 0.0f;
        for (int j = c0; j < c1; ++j) {
            const float d = __ldg(x + j) - mean;
            v = fmaf(d, d, v);
        }
        v /= (float)n;
        if (pos) o[3 + (ch - 3)] = v;                      // 3..5
        else o[139 + (ch - 9)] = v;                        // 139..140
    }
}

}  // namespace

RT_API int rt_object_embeddings(int nobj, int p_total, const float *points, const int *offsets, float *out, void *stream) {
    RT_REQUIRE(nobj >= 0 && p_total >= 0, "rt_object_embeddings: bad sizes");
    if (nobj == 0) return RT_OK;
    RT_REQUIRE(points && offsets && out, "rt_object_embeddings: null pointer");
    object_embeddings_kernel<<<nobj, 160, 0, (cudaStream_t)stream>>>(p_total, points, offsets, out);
    return rt_check_launch("object_embeddings_kernel");
}
