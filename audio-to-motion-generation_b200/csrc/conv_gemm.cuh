// Internal API of the tap-offset implicit-GEMM convolution engine (csrc/conv_gemm.cu).
//
// Every dense contraction of the generator (Conv1d k3/k4, ConvTranspose1d, Conv2d of the audio
// encoder, 1x1 convs, Linear) is expressed as
//     D[m, n] = act( sum_{tap} sum_{c} A_src(tap)[coords(m) + offset(tap), c] * W[n, k(tap, c)] + bias[n] )
// where A is a channels-last bf16 activation tensor seen through a <= 5-D TMA tensor map
// (dim 0 = channels) and a "tap" is a coordinate offset in dims 1..4.  TMA's out-of-bounds zero
// fill supplies the convolution halo and keeps taps from bleeding across clip boundaries, so no
// im2col buffer ever exists.  Up to two A sources give the skip-concat K loop (model_layers.py:364,371)
// or the two row parities of a stride-2 Conv2d.
#pragma once
#include <cuda.h>
#include <vector>
#include "a2m_common.cuh"

namespace a2m {

constexpr int kMaxTaps = 24;
constexpr int kBlockM = 128;
constexpr int kBlockK = 64;

enum Act { kActNone = 0, kActLeaky = 1, kActRelu = 2 };
enum OutType { kOutBf16 = 0, kOutF32 = 1 };

struct Tap {
    int src;            // 0 or 1: which A tensor map
    int off[4];         // coordinate offset in dims 1..4
    int channels;       // channels of that source (multiple of 64)
    long long w_off;    // element offset of this tap's [n=0, c=0] weight in the fp32 source weight
};

// Describes one A source: a strided view of a bf16 tensor, dim 0 = channels (unit stride).
struct AView {
    const void* ptr;
    int rank;                   // 2..5
    long long dims[5];          // dims[0] = channels
    long long strides[5];       // in elements; strides[0] = 1
};

struct ConvGemmDesc {
    AView a[2];
    int n_src;
    int box[4];                 // rows of one M tile along dims 1..4 (product = 128)
    int m_extent[4];            // number of output positions along dims 1..4 (tiles = ceil(extent / box))
    std::vector<Tap> taps;
    int N;                      // output channels
    // output addressing: element offset = out_base + sum_i coord_i * out_stride[i] + n
    long long out_stride[4];
    long long out_base;
    int act;
    int out_type;
    int block_n_hint = 0;       // 256: use 128 x 256 tiles if the layer allows it (N % 256 == 0, bf16 output);
                                // 512: 256 x 256 tiles on CTA pairs (same conditions + TMA-storable output)
    const float* ln_gamma = nullptr;    // LayerNorm over the N = 256 output channels of a row, applied to the fp32 accumulator + bias
    const float* ln_beta = nullptr;     // in the epilogue (real_motion_model.py:205,257 after proj_out); the planner reports in
                                        // ConvGemmPlan::ln_fused whether this layer's tiling can do it (one 256-wide tile per row)
    int split_k = 1;            // > 1: K blocks split over gridDim.z; split z writes its fp32 partial sums at
    long long split_stride = 0; // out + z * split_stride (act must be none; bias added by split 0); the consumer
                                // sums the planes in a fixed order (deterministic, batch-invariant)
};

// Device-side launch record (kernel parameter), built once per (layer, batch size).
struct ConvGemmParams {
    CUtensorMap a_map[2];
    CUtensorMap b_map;
    CUtensorMap c_map;          // bf16 output as a 5-D strided view (TMA store epilogue); valid iff c_tma
    int c_tma;
    int n_taps;
    int tap_src[kMaxTaps];
    int tap_chunks[kMaxTaps];
    int tap_off[kMaxTaps][4];
    int box[4];
    int tiles[4];
    int m_extent[4];
    long long out_stride[4];
    long long out_base;
    int N;
    int k_blocks;
    int act;
    int out_type;
    int split_k;
    long long split_stride;
    const float* ln_gamma;      // non-null: LayerNorm(256) fused into the epilogue (pair kernel, one tile per cluster)
    const float* ln_beta;
};

struct ConvGemmPlan {
    ConvGemmParams p;
    const void* w_packed;       // bf16 [N_pad, K]
    const float* bias;          // fp32 [N] (folded), may be null
    void* out;
    int block_n;
    int stages;                 // shared-memory ring depth of the 128 x 128 variant (2 or 3)
    int ln_fused;               // 1: the requested LayerNorm runs in this launch's epilogue
    int pair;                   // the CTA-pair kernel (cta_group::2, 256 x 256 tiles): 1 persistent, 2 one tile per cluster
    dim3 grid;
    long long flops;            // 2 * M * N * K of the valid output rows
};

// K = sum over taps of channels
long long conv_gemm_k(const ConvGemmDesc& d);

// Pack fp32 weights to the engine's bf16 [n_pad, K] K-major layout (K order = taps x channels),
// folding an optional per-output-channel scale (BatchNorm) into the rows.
//   w element (n, tap, c) = w_src[n * w_stride_n + tap.w_off + c * w_stride_c]
//   destination element (n, k) = w_packed[n * dst_ld + dst_col + k]   (dst_ld = 0 -> dense [N, K])
int pack_weights(const float* w_src, long long w_stride_n, long long w_stride_c, const std::vector<Tap>& taps, int N,
                 const float* scale /* nullable [N] */, __nv_bfloat16* w_packed, cudaStream_t stream,
                 long long dst_ld = 0, long long dst_col = 0);

// bias'[n] = (conv_bias[n] - mean[n]) * gamma[n] / sqrt(var[n] + eps) + beta[n];  scale[n] = gamma / sqrt(var + eps)
int fold_batchnorm(const float* conv_bias, const float* gamma, const float* beta, const float* mean, const float* var,
                   float eps, int N, float* scale_out, float* bias_out, cudaStream_t stream);

int conv_gemm_plan(const ConvGemmDesc& d, const void* w_packed, const float* bias, void* out, ConvGemmPlan* plan);
int conv_gemm_launch(const ConvGemmPlan& plan, int* err_flag, cudaStream_t stream);

}  // namespace a2m
