// Launcher contract of the tensor-core row-tile MLP kernel (mlp_tc.cu).
#pragma once
#include "common.cuh"

enum { RT_MLP_LOAD_ROWS = 0, RT_MLP_LOAD_GATHER = 1 };
enum { RT_MLP_OUT_ROWS = 0, RT_MLP_OUT_MAXPOOL = 1 };
constexpr int RT_MLP_MAX_LAYERS = 4;

// weights of one layer: fp16 hi/lo planes of 2^10 * W in the K-major core-matrix layout
// [plane hi, lo][kc = k/8][row group = n/8][8 rows][8 halfs]   (k, n already padded to multiples of 16)
struct RtMlpLayer {
    const void *wpack;
    const float *bias;  // n entries (padded) or null
    int k, n, act;
};
struct RtMlpSeg {
    const float *x;
    int ldx, k;         // k real columns; occupies ceil16(k) columns of layer 0's K
    // optional three-point interpolation of the segment (PointnetFPModule, reference lib/pointnet2_modules.py:141-146):
    // value[row, c] = fma(w2, x[i2, c], fma(w0, x[i0, c], w1 * x[i1, c])) with (i, w) = nn_idx / nn_w[row, 0..2] and
    // x rows taken from cloud (row / nn_n) of nn_m points.  nn_idx == nullptr: plain rows.
    const int *nn_idx;
    const float *nn_w;
    int nn_n, nn_m;
};
struct RtMlpTc {
    long long rows;
    int load_mode, out_mode;
    // RT_MLP_LOAD_ROWS
    int nseg;
    RtMlpSeg seg[4];
    // RT_MLP_LOAD_GATHER: relu(y[cloud, idx[row], yoff + c] + wx[c,:].(xyz_in[idx] - xyz_c[row / ns]) + b1[c]), c < c1
    const float *y;
    int ldy, yoff, n_in;
    const int *idx;
    const float *xyz_in, *xyz_c, *wx, *b1;
    int npts, ns, c1;
    int nlayers;
    RtMlpLayer layer[RT_MLP_MAX_LAYERS];
    const float *cloud_bias;   // added to layer 0's output: cloud_bias[row / rows_per_cloud, c]
    int rows_per_cloud, cloud_bias_ld;
    float *out;                // rows: out[row * ldo + ooff + c]; maxpool: out[(row / ns) * ldo + ooff + c], c < n_out
    int ldo, ooff, n_out;
    float *mid_out;            // optional: the (activated) output of layer `mid_layer` (not the last) is also written as fp32 rows,
    int mid_ldo, mid_layer;    //           mid_out[row * mid_ldo + c], c < layer[mid_layer].n -- two chained GEMMs, both results kept
    int *status;               // optional device status word (bit 1: fp16 range exceeded)
    int tmem_cols, d_cols, a_cols, ns_shift, nsplit;  // filled by the launcher (nsplit: accumulators per layer, 1 or 2)
};
int rt_launch_mlp_tc(RtMlpTc a, cudaStream_t st);
// sa_tc.cu: the same job for a gather / max-pool scale with one tensor-core layer, gathering from a shared-memory copy of the
// cloud's projected features; RT_ERR_UNSUPPORTED (no error text) when the shape is not covered -> use rt_launch_mlp_tc
int rt_launch_sa_tc(const RtMlpTc &m, int clouds, cudaStream_t st);

// device-side weight packing: W (n_real x sum k_s, given as column segments of row-major fp32 matrices) ->
// the layout above with every segment padded to 16 columns and n padded to n_pad rows (zeros)
struct RtPackSeg { const float *w; int ldw, k; };
int rt_launch_pack_umma(void *dst, int n_real, int n_pad, const RtPackSeg *segs, int nseg, cudaStream_t st);
