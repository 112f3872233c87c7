// Internal API of the non-GEMM kernels of the generator forward (csrc/layers.cu).  All activations
// are channels-last bf16 ([B, T, C]); reductions, softmax, LayerNorm and the attention coefficients are
// evaluated in fp32.
#pragma once
#include <memory>
#include "a2m_common.cuh"

namespace a2m {

// AudioEncoder conv 0: Conv2d(1 -> 64, k4, s2, p1) + BatchNorm(eval) + LeakyReLU(0.2)  (model_layers.py:252)
//   mel [B, T, F] fp32 (element strides stride_b, stride_t, 1: the D2 adapter slice is read in place)
//   -> out [B, T/2, F/2, 64] bf16.  w_folded [16 taps][64 channels] fp32 (BN scale folded), bias_folded [64].
//   n_inner > 0: clip b is window (b % n_inner) of stream (b / n_inner), at mel + stream * stride_outer + window * stride_b
int launch_conv0(const float* mel, long long stride_b, long long stride_t, int B, int T, int F, const float* w_folded,
                 const float* bias_folded, __nv_bfloat16* out, cudaStream_t stream, int n_inner = 0, long long stride_outer = 0);

// LeakyReLU(0.2) + F.interpolate(size=(T,1), mode='bilinear') + squeeze (model_layers.py:107,277-279) applied
// to the centre column computed by the last encoder conv (pre-activation, split-K sums):
// in [n_planes, B, Hc, C] fp32 (summed over the planes in order) -> out [B, T, C] bf16.
int launch_time_interp(const float* in, int n_planes, int B, int Hc, int T, int C, __nv_bfloat16* out, cudaStream_t stream);

// SelfAttention (model_layers.py:133-146) after the fused q|k|v 1x1-conv GEMM:
//   qkv [B, T, 2*d + C] bf16 (q: d, k: d, v: C; d = C/8), x [B, T, C] bf16
//   out = gamma * softmax(q k^T) v + x (+ res2 if given: the ResBlock skip, :190)
int launch_attention(const __nv_bfloat16* qkv, const __nv_bfloat16* x, const __nv_bfloat16* res2, const float* gamma,
                     int B, int T, int C, __nv_bfloat16* out, cudaStream_t stream);

// The same block fused with its q|k|v projection for the 256-channel decoder layers (csrc/attn_fused.cu):
// w_qkv [320, 256] bf16 (q: 32, k: 32, v: 256 rows), bias_qkv [320]; T must divide 128.
struct AttnFusedPlan;
bool attn_fused_supported(int T, int C);
int attn_fused_plan(const __nv_bfloat16* w_qkv, const float* bias_qkv, const float* gamma, const __nv_bfloat16* x,
                    const __nv_bfloat16* res2, int B, int T, int C, __nv_bfloat16* out, std::shared_ptr<AttnFusedPlan>* plan,
                    const float* ca_w0 = nullptr, const float* ca_b0 = nullptr, const float* ca_w2 = nullptr,
                    const float* ca_b2 = nullptr);      // ca_*: optional ChannelAttention (hidden 32) fused into the epilogue
int attn_fused_launch(const AttnFusedPlan& plan, int* err_flag, cudaStream_t stream);

// Tensor-core attention core for the wide UNet attentions (csrc/attn_core.cu), after the q|k|v GEMM: T must divide
// 128, C a multiple of 256 with C / 8 in {64, 128, 192, 256}.
struct AttnCorePlan;
bool attn_core_supported(int T, int C);
int attn_core_plan(const __nv_bfloat16* qkv, const float* gamma, const __nv_bfloat16* x, const __nv_bfloat16* res2, int B, int T,
                   int C, __nv_bfloat16* out, std::shared_ptr<AttnCorePlan>* plan);
int attn_core_launch(const AttnCorePlan& plan, int* err_flag, cudaStream_t stream);

// ChannelAttention (model_layers.py:167-174): x * (sigmoid(mlp(avg_T x)) + sigmoid(mlp(max_T x))), C = 256, hidden 32;
// w0 = fc.0.weight [hidden, C], w2 = fc.2.weight TRANSPOSED to [hidden, C]
int launch_channel_attention(const __nv_bfloat16* x, int B, int T, int C, int hidden, const float* w0, const float* b0,
                             const float* w2, const float* b2, __nv_bfloat16* out, cudaStream_t stream);

// LayerNorm over the last dim (256) of [rows, C] bf16 -> bf16 (real_motion_model.py:205,257)
int launch_layernorm(const __nv_bfloat16* x, long long rows, int C, const float* gamma, const float* beta,
                     __nv_bfloat16* out, cudaStream_t stream);

// Static skeleton graph shared by every (clip, frame): neighbours by node (symmetric tree, no self loops)
struct GraphTopo {
    int n_nodes;
    const int* nbr;      // device [n_nodes][kMaxDeg], -1 padded
    const int* deg;      // device [n_nodes]
};
constexpr int kMaxDeg = 6;
constexpr int kJointFeat = 64;
constexpr int kGatHeads = 4;

// Fused five-layer GNN stack (csrc/gnn_fused.cu): GAT, GraphConv, GAT, GraphConv, GAT, each followed by
// LayerNorm(64) -> LeakyReLU(0.2) -> + residual, x_in/x_out [n_graphs * J, 64] bf16.
struct GnnFusedWeights {
    const __nv_bfloat16* gat_w[3];      // [272, 64] bf16: lin.weight + 16 folded attention rows (gat_fold_attention)
    const float* gat_bias[3];
    const __nv_bfloat16* gc_w[2];       // [64, 128] bf16 = [W_rel | W_root]
    const float* gc_bias[2];
    const float *ln_w[5], *ln_b[5];
};
struct GnnFusedPlan;
// wext [272][64] bf16 with rows 0..255 = lin.weight already packed: fills rows 256..271 with W_h^T att_src / att_dst
// (hi and lo bf16 parts) for the four heads.  att_src / att_dst: fp32 [4][64].
int gat_fold_attention(__nv_bfloat16* wext, const float* att_src, const float* att_dst, cudaStream_t stream);
// Graphs are tiled per group of `group_graphs` consecutive graphs (a clip's T frames): a graph's position inside
// its 128-row tile -- and with it the tensor-core accumulation order -- never depends on how clips are batched.
int gnn_fused_plan(const GnnFusedWeights& w, GraphTopo topo, long long n_groups, int group_graphs, const __nv_bfloat16* x_in,
                   __nv_bfloat16* x_out, std::shared_ptr<GnnFusedPlan>* out);
int gnn_fused_launch(const GnnFusedPlan& plan, int* err_flag, cudaStream_t stream);

// Internal losses of SelfAttention_G.forward (real_motion_model.py:307-461) on pose [B, T, 104] fp32.
// triples: device int [n][3] (hand first, joints already offset by 10, then body);
// losses_out[0] = angle loss, losses_out[1] = bone loss (only if real_pose != NULL); scratch: 4 doubles.
int launch_pose_losses(const float* pose, const float* real_pose, int B, int T, const int* triples, int n_hand,
                       int n_body, const int* parents, double* scratch, float* losses_out, cudaStream_t stream);

// Small kernels of the discriminator forward (SelfAttention_D, real_motion_model.py:580-642)
int launch_pose_pad(const float* pose, int B, int T, int T_alloc, int C, int C_pad, __nv_bfloat16* out, cudaStream_t stream);
int launch_mean_time(const __nv_bfloat16* x, int B, int T, int C, __nv_bfloat16* out, cudaStream_t stream);
// one GATConv(64, 64, heads=4, concat=False) layer alone; wt = lin.weight transposed [64][256] fp32
int launch_gat_single(const __nv_bfloat16* x, long long n_graphs, int J, const float* wt, const float* att_src,
                      const float* att_dst, const float* bias, const int* nbr, const int* deg, __nv_bfloat16* out,
                      cudaStream_t stream);
// Conv1d(2 C -> 1, k3 p1) over cat([x, graph features repeated over time]); w [3][2 C] fp32 tap-major
int launch_disc_logits(const __nv_bfloat16* x, const __nv_bfloat16* xg, const float* w, const float* bias, int B, int T,
                       int C, float* out, cudaStream_t stream);

int launch_f32_to_bf16(const float* in, long long n, __nv_bfloat16* out, cudaStream_t stream);
int launch_bf16_to_f32(const __nv_bfloat16* in, long long n, float* out, cudaStream_t stream);
// [B, C, T] fp32 (reference NCW layout) <-> [B, T, C] bf16 for the AudioEncoder / UNet1D drop-in surfaces
int launch_ncw_to_btc(const float* in, int B, int C, int T, __nv_bfloat16* out, cudaStream_t stream);
int launch_btc_to_ncw(const __nv_bfloat16* in, int B, int C, int T, float* out, cudaStream_t stream);

}  // namespace a2m
