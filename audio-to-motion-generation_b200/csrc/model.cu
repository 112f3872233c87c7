// SelfAttention_G forward (real_motion_model.py:154-278, eval mode) as a resident launch program:
// a2m_model_create folds BatchNorm, packs every weight to bf16 once and keeps it in HBM;
// a2m_model_forward builds (and caches per input shape) an activation arena plus the list of kernel
// launches -- tcgen05 implicit-GEMM convolutions / linears (conv_gemm.cu) and the small fused kernels
// of layers.cu -- and enqueues them on the caller's stream.  Decisions D1 (up_attention before the
// skip concat) and D3 (eval semantics) of SURVEY.md section 8 apply.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>
#include "conv_gemm.cuh"
#include "layers.cuh"

using namespace a2m;

namespace {

// pats/data_loading/skeleton.py:94-110
const int kParents[52] = {-1, 0, 1, 2, 0, 4, 5, 0, 7, 7, 6, 10, 11, 12, 13, 10, 15, 16, 17, 10, 19, 20, 21, 10, 23, 24,
                          25, 10, 27, 28, 29, 3, 31, 32, 33, 34, 31, 36, 37, 38, 31, 40, 41, 42, 31, 44, 45, 46, 31, 48,
                          49, 50};
constexpr int kBodyJoints = 10, kHandJoints = 42, kPoseFeats = 104, kBodyFeats = 20;
constexpr float kBnEps = 1e-5f;
constexpr int kConv4Splits = 8;           // split-K planes of the last encoder conv (K = 24 taps x 512)

struct Param {
    const float* f32 = nullptr;
    const long long* i64 = nullptr;
    int ndim = 0;
    long long shape[4] = {0, 0, 0, 0};
    long long numel() const { long long n = 1; for (int i = 0; i < ndim; ++i) n *= shape[i]; return n; }
};

struct LayerW {                 // one packed GEMM layer
    __nv_bfloat16* w = nullptr; // [N, K]
    float* bias = nullptr;      // [N] folded (may be null)
    int N = 0;
    std::vector<Tap> taps;
    int act = kActNone;
};

struct AttnW { LayerW qkv; const float* gamma = nullptr; int C = 0; };
struct ChanW { const float *w0, *b0, *w2, *b2; int C, hidden; };
struct GatW { LayerW lin; const float* bias; };
struct LnW { const float *w, *b; };
struct ResW { LayerW c1, c2; AttnW attn; };

struct DecoderW {
    ResW pre_res; LayerW pre_conv; AttnW pre_attn; ChanW pre_chan; bool chan_first;
    LayerW proj_in;
    GatW gat[3]; LayerW gconv[2]; LnW ln64[5];
    LayerW proj_out; LnW norm;
    ResW post_res; LayerW post_conv; AttnW post_attn; ChanW post_chan; bool has_post_chan;
    LayerW logits;
    int joints;
    int *nbr, *deg;             // device topology tables
};

struct ForwardPlan {
    int B, T, F;
    int T_out = 0;                              // AudioEncoder.forward(x, time_steps): length of the encoder output (= T in the generator)
    unsigned long long last_use = 0;            // plan cache: least recently used goes first
    void* arena = nullptr;
    std::vector<std::function<int(cudaStream_t)>> ops;
    std::vector<int> op_is_gemm;                // parallel to ops
    std::vector<std::string> op_name;           // parallel to ops (measurement aid)
    std::vector<long long> op_flops;            // parallel to ops: 2*M*N*K for GEMMs, 0 otherwise
    int enc_end = 0, unet_end = 0;              // op index ranges: [0,enc_end) encoder, [enc_end,unet_end) unet,
    int body_end = 0;                           // [unet_end,body_end) body decoder, [body_end,size) hand decoder
    float* mel_in = nullptr;                    // staging (fp32) used by the stage entry points
    __nv_bfloat16 *enc_out = nullptr, *unet_out = nullptr;
    float* pose_stage = nullptr;                // fp32 [B, T, 104] written by the two logits GEMMs
    long long gemm_flops = 0;
    ~ForwardPlan() { if (arena) cudaFree(arena); }
};

}  // namespace

struct a2m_model {
    int device = 0;
    std::map<std::string, Param> params;
    std::vector<void*> owned;                   // device allocations freed on destroy
    // encoder
    float *conv0_w = nullptr, *conv0_b = nullptr;
    LayerW enc[5];                              // [1..4] used
    // unet
    LayerW ds[4], bott, up0_even, up0_odd, up1, up2_even, up2_odd, up3, final_conv;
    AttnW bott_attn, up_attn;
    DecoderW dec[2];
    // discriminator (a2m_disc_*): SelfAttention_D, real_motion_model.py:464-642
    bool has_disc = false;
    int disc_down = 0, disc_c = 0;              // n_downsampling; channels after conv2 (512 for the defaults)
    std::vector<LayerW> disc_conv;              // conv1 (2), conv2 (2 per stage), conv3 (3)
    AttnW disc_attn;
    LayerW disc_proj[2], disc_out[2];
    const float *disc_gat_wt[2] = {nullptr, nullptr}, *disc_gat_src[2] = {nullptr, nullptr},
                *disc_gat_dst[2] = {nullptr, nullptr}, *disc_gat_bias[2] = {nullptr, nullptr};
    int *disc_nbr[2] = {nullptr, nullptr}, *disc_deg[2] = {nullptr, nullptr};
    const float *disc_logit_w = nullptr, *disc_logit_b = nullptr;
    // stand-alone building block (a2m_block_*): one layer class of model_layers.py with its own forward
    int blk_kind = 0, blk_cin = 0, blk_cout = 0;
    LayerW blk_conv, blk_even, blk_odd;
    AttnW blk_attn;
    ChanW blk_chan;
    ResW blk_res;
    int* triples = nullptr; int n_hand_triples = 0, n_body_triples = 0;
    int* parents = nullptr;
    double* loss_scratch = nullptr;
    float* denorm = nullptr;                    // [2][104] mean | std of the optional output de-normalisation
    // optional timeline (a2m_model_timeline_*): one event after every op of every recorded forward
    std::vector<cudaEvent_t> tl_events;
    int tl_capacity = 0, tl_steps = 0, tl_ops = 0;
    bool denorm_on = false;
    int* err_flag = nullptr;
    bool has_encoder = false, has_unet = false, has_decoders = false;
    // per-forward mutable slots read by the op closures
    const float* cur_mel = nullptr;
    long long cur_stride_b = 0, cur_stride_t = 0;
    int cur_n_inner = 0;                        // > 0: the batch is (streams x windows), a2m_model_forward_windows
    long long cur_stride_outer = 0;
    int next_n_inner = 0;
    long long next_stride_outer = 0;
    float* cur_pose = nullptr;
    std::map<std::string, std::unique_ptr<ForwardPlan>> plans;
    unsigned long long plan_clock = 0;
    long long plan_generation = 0;              // bumped whenever a cached plan (and its arena) is released
    std::string build_error;
    // the two decoder branches are independent: the body branch runs on a side stream, forked / joined with events
    cudaStream_t side_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
};

namespace {

struct Builder {
    a2m_model* m;
    cudaStream_t s;
    int rc = A2M_OK;

    template <typename T>
    T* alloc(size_t n) {
        void* p = nullptr;
        if (rc != A2M_OK) return nullptr;
        cudaError_t e = cudaMalloc(&p, n * sizeof(T) + 256);
        if (e != cudaSuccess) { a2m_set_error("model: cudaMalloc(%zu) failed: %s", n * sizeof(T), cudaGetErrorString(e)); rc = (int)e; return nullptr; }
        m->owned.push_back(p);
        return static_cast<T*>(p);
    }
    const Param* find(const std::string& name, bool required = true) {
        auto it = m->params.find(name);
        if (it == m->params.end()) {
            if (required && rc == A2M_OK) { a2m_set_error("model: missing tensor '%s' in the state_dict", name.c_str()); rc = A2M_ERR_ARGUMENT; }
            return nullptr;
        }
        return &it->second;
    }
    const float* f32(const std::string& name, long long numel) {
        const Param* p = find(name);
        if (!p) return nullptr;
        if (!p->f32 || p->numel() != numel) {
            if (rc == A2M_OK) { a2m_set_error("model: tensor '%s' has %lld elements, expected %lld fp32", name.c_str(), p->numel(), numel); rc = A2M_ERR_ARGUMENT; }
            return nullptr;
        }
        return p->f32;
    }
    void check(int r) { if (rc == A2M_OK && r != A2M_OK) rc = r; }
    // small fp32 tensors the kernels read directly: copied, so the handle never aliases caller memory
    const float* keep(const std::string& name, long long numel) {
        const float* src = f32(name, numel);
        float* dst = alloc<float>(static_cast<size_t>(numel));
        if (rc != A2M_OK) return nullptr;
        cudaMemcpyAsync(dst, src, numel * sizeof(float), cudaMemcpyDeviceToDevice, s);
        return dst;
    }

    // BatchNorm(eval) folded into per-channel scale / bias; returns scale (device) and fills *bias
    const float* fold_bn(const std::string& conv, const std::string& bn, int N, float** bias_out) {
        float* scale = alloc<float>(N);
        *bias_out = alloc<float>(N);
        const float *cb = f32(conv + ".bias", N), *g = f32(bn + ".weight", N), *b = f32(bn + ".bias", N),
                    *mu = f32(bn + ".running_mean", N), *var = f32(bn + ".running_var", N);
        if (rc != A2M_OK) return nullptr;
        check(fold_batchnorm(cb, g, b, mu, var, kBnEps, N, scale, *bias_out, s));
        return scale;
    }
    void pack(LayerW& L, const float* w, long long sn, long long sc, const float* scale) {
        if (rc != A2M_OK) return;
        long long K = 0;
        for (const Tap& t : L.taps) K += t.channels;
        L.w = alloc<__nv_bfloat16>(static_cast<size_t>(L.N) * K);
        if (rc != A2M_OK) return;
        check(pack_weights(w, sn, sc, L.taps, L.N, scale, L.w, s));
    }
    static Tap tap(int src, int o1, int o2, int o3, int channels, long long w_off) {
        Tap t; t.src = src; t.off[0] = o1; t.off[1] = o2; t.off[2] = o3; t.off[3] = 0; t.channels = channels; t.w_off = w_off;
        return t;
    }

    // ConvNormRelu 1-D, k3 s1 p1 over one or two concatenated sources (model_layers.py:60-66,94-118)
    void conv1d_k3(LayerW& L, const std::string& p, int c0, int c1, int N) {
        const int ct = c0 + c1;
        for (int j = 0; j < 3; ++j) L.taps.push_back(tap(0, j - 1, 0, 0, c0, j));
        if (c1) for (int j = 0; j < 3; ++j) L.taps.push_back(tap(1, j - 1, 0, 0, c1, static_cast<long long>(c0) * 3 + j));
        L.N = N; L.act = kActLeaky;
        const float* scale = fold_bn(p + ".conv", p + ".norm", N, &L.bias);
        pack(L, f32(p + ".conv.weight", static_cast<long long>(N) * ct * 3), ct * 3LL, 3, scale);
    }
    // ConvNormRelu 1-D downsample, k4 s2 p1: view [B, T/2, 2, C]; tap j reads input 2t + j - 1
    void conv1d_k4s2(LayerW& L, const std::string& p, int C, int N) {
        L.taps = {tap(0, 1, -1, 0, C, 0), tap(0, 0, 0, 0, C, 1), tap(0, 1, 0, 0, C, 2), tap(0, 0, 1, 0, C, 3)};
        L.N = N; L.act = kActLeaky;
        const float* scale = fold_bn(p + ".conv", p + ".norm", N, &L.bias);
        pack(L, f32(p + ".conv.weight", static_cast<long long>(N) * C * 4), C * 4LL, 4, scale);
    }
    // Conv1d + BatchNorm1d + LeakyReLU of the discriminator's nn.Sequential stacks (real_motion_model.py:504-550), names
    // given explicitly.  kind 0: k4 s2 p1 (view [B, T/2, 2, C]); 1: k4 s1 p1 (taps -1..2, one step shorter); 2: k3 s1 p1.
    // c_pad > c_in: the input tensor carries zero channels up to c_pad (104 pose features -> 128).
    void conv_bn(LayerW& L, const std::string& conv, const std::string& bn, int kind, int c_in, int c_pad, int N) {
        const int k = kind == 2 ? 3 : 4;
        if (kind == 0) L.taps = {tap(0, 1, -1, 0, c_pad, 0), tap(0, 0, 0, 0, c_pad, 1), tap(0, 1, 0, 0, c_pad, 2), tap(0, 0, 1, 0, c_pad, 3)};
        else for (int j = 0; j < k; ++j) L.taps.push_back(tap(0, j - 1, 0, 0, c_pad, j));
        L.N = N; L.act = kActLeaky;
        const float* scale = fold_bn(conv, bn, N, &L.bias);
        const float* w = f32(conv + ".weight", static_cast<long long>(N) * c_in * k);
        if (rc != A2M_OK) return;
        if (c_pad != c_in) {                                   // zero-padded copy [N][c_pad][k]
            float* wp = alloc<float>(static_cast<size_t>(N) * c_pad * k);
            if (rc != A2M_OK) return;
            cudaMemsetAsync(wp, 0, static_cast<size_t>(N) * c_pad * k * sizeof(float), s);
            cudaMemcpy2DAsync(wp, static_cast<size_t>(c_pad) * k * 4, w, static_cast<size_t>(c_in) * k * 4,
                              static_cast<size_t>(c_in) * k * 4, N, cudaMemcpyDeviceToDevice, s);
            w = wp;
        }
        pack(L, w, static_cast<long long>(c_pad) * k, k, scale);
    }
    // ConvTranspose1D k3 s2 p1 op1 + BN + ReLU (model_layers.py:200-215) as two parity GEMMs:
    //   out[2j] = W[:,:,1] x[j];  out[2j+1] = W[:,:,2] x[j] + W[:,:,0] x[j+1];  weight layout [C_in, C_out, 3]
    void conv_transpose(LayerW& even, LayerW& odd, const std::string& p, int C, int N) {
        float* bias = nullptr;
        const float* scale = fold_bn(p + ".conv_transpose", p + ".bn", N, &bias);
        const float* w = f32(p + ".conv_transpose.weight", static_cast<long long>(C) * N * 3);
        even.taps = {tap(0, 0, 0, 0, C, 1)};
        odd.taps = {tap(0, 0, 0, 0, C, 2), tap(0, 1, 0, 0, C, 0)};
        even.N = odd.N = N; even.act = odd.act = kActRelu; even.bias = odd.bias = bias;
        pack(even, w, 3, N * 3LL, scale);
        pack(odd, w, 3, N * 3LL, scale);
    }
    // nn.Linear / Conv1d k1: weight [N, C]
    void linear(LayerW& L, const std::string& wname, const std::string& bname, int C, int N, int act) {
        L.taps = {tap(0, 0, 0, 0, C, 0)};
        L.N = N; L.act = act;
        if (!bname.empty()) L.bias = const_cast<float*>(keep(bname, N));
        pack(L, f32(wname, static_cast<long long>(N) * C), C, 1, nullptr);
    }
    // SelfAttention: q | k | v 1x1 convs fused along N (model_layers.py:126-128)
    void attention(AttnW& A, const std::string& p, int C) {
        const int d = C / 8, N = 2 * d + C;
        A.C = C;
        A.gamma = keep(p + ".gamma", 1);
        LayerW& L = A.qkv;
        L.taps = {tap(0, 0, 0, 0, C, 0)};
        L.N = N; L.act = kActNone;
        L.w = alloc<__nv_bfloat16>(static_cast<size_t>(N) * C);
        L.bias = alloc<float>(N);
        const float *wq = f32(p + ".query_conv.weight", static_cast<long long>(d) * C), *wk = f32(p + ".key_conv.weight", static_cast<long long>(d) * C),
                    *wv = f32(p + ".value_conv.weight", static_cast<long long>(C) * C);
        const float *bq = f32(p + ".query_conv.bias", d), *bk = f32(p + ".key_conv.bias", d), *bv = f32(p + ".value_conv.bias", C);
        if (rc != A2M_OK) return;
        check(pack_weights(wq, C, 1, L.taps, d, nullptr, L.w, s));
        check(pack_weights(wk, C, 1, L.taps, d, nullptr, L.w + static_cast<size_t>(d) * C, s));
        check(pack_weights(wv, C, 1, L.taps, C, nullptr, L.w + static_cast<size_t>(2 * d) * C, s));
        cudaMemcpyAsync(L.bias, bq, d * 4, cudaMemcpyDeviceToDevice, s);
        cudaMemcpyAsync(L.bias + d, bk, d * 4, cudaMemcpyDeviceToDevice, s);
        cudaMemcpyAsync(L.bias + 2 * d, bv, C * 4, cudaMemcpyDeviceToDevice, s);
    }
    void channel(ChanW& Cw, const std::string& p, int C) {
        Cw.C = C; Cw.hidden = C / 8;
        Cw.w0 = keep(p + ".fc.0.weight", static_cast<long long>(Cw.hidden) * C); Cw.b0 = keep(p + ".fc.0.bias", Cw.hidden);
        Cw.b2 = keep(p + ".fc.2.bias", C);
        // fc.2.weight [C, hidden] is kept transposed ([hidden, C]) so that channel-per-thread reads coalesce
        const float* w2 = f32(p + ".fc.2.weight", static_cast<long long>(C) * Cw.hidden);
        float* w2t = alloc<float>(static_cast<size_t>(C) * Cw.hidden);
        if (rc != A2M_OK) return;
        std::vector<float> h(static_cast<size_t>(C) * Cw.hidden), ht(h.size());
        cudaError_t e = cudaMemcpy(h.data(), w2, h.size() * sizeof(float), cudaMemcpyDeviceToHost);
        for (int c = 0; c < C; ++c)
            for (int u = 0; u < Cw.hidden; ++u) ht[static_cast<size_t>(u) * C + c] = h[static_cast<size_t>(c) * Cw.hidden + u];
        if (e == cudaSuccess) e = cudaMemcpy(w2t, ht.data(), ht.size() * sizeof(float), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { a2m_set_error("model: transposing '%s.fc.2.weight': %s", p.c_str(), cudaGetErrorString(e)); rc = (int)e; return; }
        Cw.w2 = w2t;
    }
    void resblock(ResW& R, const std::string& p, int C) {
        conv1d_k3(R.c1, p + ".conv1", C, 0, C);
        conv1d_k3(R.c2, p + ".conv2", C, 0, C);
        attention(R.attn, p + ".attention", C);
    }
    void gat(GatW& G, const std::string& p) {
        const Param* w = find(p + ".lin.weight", false);
        std::string wname = p + ".lin.weight";
        if (!w) {                                           // torch_geometric version-dependent key names
            for (const char* alt : {".lin_src.weight", ".lin_l.weight"})
                if (find(p + alt, false)) { wname = p + alt; break; }
        }
        // extended weight [272][64]: rows 0..255 = lin.weight (bf16), rows 256..271 = the attention vectors folded
        // through W (a_src . (W_h x) = (W_h^T a_src) . x), so the logits come out of the same MMA as H
        LayerW& L = G.lin;
        L.taps = {tap(0, 0, 0, 0, kJointFeat, 0)};
        L.N = kGatHeads * kJointFeat; L.act = kActNone;
        L.w = alloc<__nv_bfloat16>(static_cast<size_t>(272) * kJointFeat);
        const float* wsrc = f32(wname, static_cast<long long>(L.N) * kJointFeat);
        const float *as = f32(p + ".att_src", kGatHeads * kJointFeat), *ad = f32(p + ".att_dst", kGatHeads * kJointFeat);
        if (rc != A2M_OK) return;
        check(pack_weights(wsrc, kJointFeat, 1, L.taps, L.N, nullptr, L.w, s));
        check(gat_fold_attention(L.w, as, ad, s));
        G.bias = keep(p + ".bias", kJointFeat);
    }
    // GraphConv: lin_rel(sum_j x_j) + lin_root(x_i) = [agg | x] [W_rel | W_root]^T + b_rel  (two A sources)
    void graph_conv(LayerW& L, const std::string& p) {
        L.taps = {tap(0, 0, 0, 0, kJointFeat, 0), tap(1, 0, 0, 0, kJointFeat, 0)};
        L.N = kJointFeat; L.act = kActNone;
        L.w = alloc<__nv_bfloat16>(kJointFeat * 2 * kJointFeat);
        L.bias = const_cast<float*>(keep(p + ".lin_rel.bias", kJointFeat));
        const float *wr = f32(p + ".lin_rel.weight", kJointFeat * kJointFeat), *wo = f32(p + ".lin_root.weight", kJointFeat * kJointFeat);
        if (rc != A2M_OK) return;
        std::vector<Tap> one = {tap(0, 0, 0, 0, kJointFeat, 0)};
        check(pack_weights(wr, kJointFeat, 1, one, kJointFeat, nullptr, L.w, s, 2 * kJointFeat, 0));
        check(pack_weights(wo, kJointFeat, 1, one, kJointFeat, nullptr, L.w, s, 2 * kJointFeat, kJointFeat));
    }
    void topology(DecoderW& D, const std::string& buf, int J) {
        const Param* p = find(buf);
        if (!p) return;
        if (!p->i64 || p->ndim != 2 || p->shape[0] != 2) { a2m_set_error("model: '%s' must be int64 [2, E]", buf.c_str()); rc = A2M_ERR_ARGUMENT; return; }
        const long long E = p->shape[1];
        std::vector<long long> e(2 * E);
        cudaError_t err = cudaMemcpy(e.data(), p->i64, 2 * E * 8, cudaMemcpyDeviceToHost);
        if (err != cudaSuccess) { a2m_set_error("model: reading '%s': %s", buf.c_str(), cudaGetErrorString(err)); rc = (int)err; return; }
        std::vector<int> nbr(J * kMaxDeg, -1), deg(J, 0);
        for (long long k = 0; k < E; ++k) {
            const long long src = e[k], dst = e[E + k];                // row 0 = source j, row 1 = target i
            if (src == dst) continue;
            if (src < 0 || dst < 0 || src >= J || dst >= J || deg[dst] >= kMaxDeg) {
                a2m_set_error("model: '%s' edge %lld->%lld outside a %d-node graph of degree <= %d", buf.c_str(), src, dst, J, kMaxDeg);
                rc = A2M_ERR_UNSUPPORTED; return;
            }
            nbr[dst * kMaxDeg + deg[dst]++] = static_cast<int>(src);
        }
        D.nbr = alloc<int>(nbr.size()); D.deg = alloc<int>(deg.size());
        if (rc != A2M_OK) return;
        cudaMemcpy(D.nbr, nbr.data(), nbr.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(D.deg, deg.data(), deg.size() * 4, cudaMemcpyHostToDevice);
    }
    void decoder(DecoderW& D, const std::string& part, int J, int n_feat) {
        const int C = 256;
        D.joints = J;
        const std::string pre = part + "_decoder_pre", post = part + "_decoder_post";
        resblock(D.pre_res, pre + ".0", C);
        conv1d_k3(D.pre_conv, pre + ".1", C, 0, C);
        D.chan_first = part == "body";                              // real_motion_model.py:70-75 vs :96-101
        channel(D.pre_chan, pre + (D.chan_first ? ".2" : ".3"), C);
        attention(D.pre_attn, pre + (D.chan_first ? ".3" : ".2"), C);
        linear(D.proj_in, part + "_proj_in.weight", part + "_proj_in.bias", C, J * kJointFeat, kActNone);
        for (int i = 0; i < 3; ++i) gat(D.gat[i], part + "_gcn" + std::to_string(2 * i + 1));
        for (int i = 0; i < 2; ++i) graph_conv(D.gconv[i], part + "_gcn" + std::to_string(2 * i + 2));
        for (int i = 0; i < 5; ++i) {
            D.ln64[i].w = keep(part + "_layer_norms." + std::to_string(i) + ".weight", kJointFeat);
            D.ln64[i].b = keep(part + "_layer_norms." + std::to_string(i) + ".bias", kJointFeat);
        }
        linear(D.proj_out, part + "_proj_out.weight", part + "_proj_out.bias", J * kJointFeat, C, kActNone);
        D.norm.w = keep(part + "_norm.weight", C); D.norm.b = keep(part + "_norm.bias", C);
        resblock(D.post_res, post + ".0", C);
        conv1d_k3(D.post_conv, post + ".1", C, 0, C);
        attention(D.post_attn, post + ".2", C);
        D.has_post_chan = part == "hand";
        if (D.has_post_chan) channel(D.post_chan, post + ".3", C);
        linear(D.logits, part + "_logits.weight", part + "_logits.bias", C, n_feat, kActNone);
        topology(D, part + "_edge_index_template", J);
    }
    void encoder() {
        // conv 0 (1 -> 64, k4 s2 p1): tiny, folded on the host into fp32 [16 taps][64 channels] + bias[64]
        const std::string p = "audio_encoder.conv.0";
        const float *w = f32(p + ".conv.weight", 64 * 16), *cb = f32(p + ".conv.bias", 64), *g = f32(p + ".norm.weight", 64),
                    *b = f32(p + ".norm.bias", 64), *mu = f32(p + ".norm.running_mean", 64), *var = f32(p + ".norm.running_var", 64);
        if (rc != A2M_OK) return;
        std::vector<float> hw(64 * 16), hcb(64), hg(64), hb(64), hmu(64), hvar(64);
        cudaMemcpy(hw.data(), w, hw.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(hcb.data(), cb, 256, cudaMemcpyDeviceToHost);
        cudaMemcpy(hg.data(), g, 256, cudaMemcpyDeviceToHost); cudaMemcpy(hb.data(), b, 256, cudaMemcpyDeviceToHost);
        cudaMemcpy(hmu.data(), mu, 256, cudaMemcpyDeviceToHost); cudaMemcpy(hvar.data(), var, 256, cudaMemcpyDeviceToHost);
        std::vector<float> fb(64), wt(16 * 64);          // wt: tap-major [16][64] so a warp's weight loads coalesce
        for (int c = 0; c < 64; ++c) {
            const float sc = hg[c] / sqrtf(hvar[c] + kBnEps);
            for (int i = 0; i < 16; ++i) wt[i * 64 + c] = hw[c * 16 + i] * sc;
            fb[c] = (hcb[c] - hmu[c]) * sc + hb[c];
        }
        hw = wt;
        m->conv0_w = alloc<float>(64 * 16); m->conv0_b = alloc<float>(64);
        if (rc != A2M_OK) return;
        cudaMemcpy(m->conv0_w, hw.data(), hw.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(m->conv0_b, fb.data(), 256, cudaMemcpyHostToDevice);
        // convs 1, 2: k4x4 s2 p1; dims (C, pw, W/2, H/2, B), row parity = A source
        const int cin[5] = {1, 64, 128, 256, 512}, cout[5] = {64, 128, 256, 512, 256};
        const int par[4][2] = {{1, -1}, {0, 0}, {1, 0}, {0, 1}};   // kernel index -> (parity, shift)
        for (int l = 1; l <= 2; ++l) {
            LayerW& L = m->enc[l];
            for (int i = 0; i < 4; ++i)
                for (int j = 0; j < 4; ++j)
                    L.taps.push_back(tap(par[i][0], par[j][0], par[j][1], par[i][1], cin[l], i * 4 + j));
            L.N = cout[l]; L.act = kActLeaky;
            const std::string q = "audio_encoder.conv." + std::to_string(l);
            const float* scale = fold_bn(q + ".conv", q + ".norm", L.N, &L.bias);
            pack(L, f32(q + ".conv.weight", static_cast<long long>(L.N) * cin[l] * 16), cin[l] * 16LL, 16, scale);
        }
        {   // conv 3: k3x3 s1 p1; dims (C, W, H, B)
            LayerW& L = m->enc[3];
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) L.taps.push_back(tap(0, j - 1, i - 1, 0, cin[3], i * 3 + j));
            L.N = cout[3]; L.act = kActLeaky;
            const std::string q = "audio_encoder.conv.3";
            const float* scale = fold_bn(q + ".conv", q + ".norm", L.N, &L.bias);
            pack(L, f32(q + ".conv.weight", static_cast<long long>(L.N) * cin[3] * 9), cin[3] * 9LL, 9, scale);
        }
        {   // conv 4: k(3,8) p(1,3); only the centre output column survives the (T,1) resize, so the W
            // offset of every tap is fixed up at plan time (depends on F): off[0] = j + (w_centre - 3)
            LayerW& L = m->enc[4];
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 8; ++j) L.taps.push_back(tap(0, j, i - 1, 0, cin[4], i * 8 + j));
            L.N = cout[4]; L.act = kActLeaky;
            const std::string q = "audio_encoder.conv.4";
            const float* scale = fold_bn(q + ".conv", q + ".norm", L.N, &L.bias);
            pack(L, f32(q + ".conv.weight", static_cast<long long>(L.N) * cin[4] * 24), cin[4] * 24LL, 24, scale);
        }
    }
    void unet() {
        const int c = 256;
        conv1d_k3(m->ds[0], "unet.downsample_layers.0", c, 0, 2 * c);
        conv1d_k4s2(m->ds[1], "unet.downsample_layers.1", 2 * c, 2 * c);
        conv1d_k3(m->ds[2], "unet.downsample_layers.2", 2 * c, 0, 4 * c);
        conv1d_k4s2(m->ds[3], "unet.downsample_layers.3", 4 * c, 4 * c);
        conv1d_k3(m->bott, "unet.bottleneck", 4 * c, 0, 8 * c);
        attention(m->bott_attn, "unet.bottleneck_attention", 8 * c);
        conv_transpose(m->up0_even, m->up0_odd, "unet.upsample_layers.0", 8 * c, 4 * c);
        attention(m->up_attn, "unet.up_attention", 4 * c);
        conv1d_k3(m->up1, "unet.upsample_layers.1", 4 * c, 4 * c, 4 * c);
        conv_transpose(m->up2_even, m->up2_odd, "unet.upsample_layers.2", 4 * c, 2 * c);
        conv1d_k3(m->up3, "unet.upsample_layers.3", 2 * c, 2 * c, 2 * c);
        linear(m->final_conv, "unet.final_conv.weight", "unet.final_conv.bias", 2 * c, c, kActNone);
    }
    void losses() {
        // real_motion_model.py:280-304: (parent, joint, first child) triples; hand joints are offset by 10
        std::vector<int> t;
        int nh = 0, nb = 0;
        for (int i = 0; i < kHandJoints; ++i) {
            const int par = kParents[i + 10] >= 10 ? kParents[i + 10] - 10 : -1;
            if (par == -1) continue;
            for (int j = i + 1; j < kHandJoints; ++j)
                if (kParents[j + 10] - 10 == i) { t.insert(t.end(), {par + 10, i + 10, j + 10}); ++nh; break; }
        }
        for (int i = 0; i < kBodyJoints; ++i) {
            const int par = kParents[i] < kBodyJoints ? kParents[i] : -1;
            if (par == -1) continue;
            for (int j = i + 1; j < kBodyJoints; ++j)
                if (kParents[j] == i) { t.insert(t.end(), {par, i, j}); ++nb; break; }
        }
        m->n_hand_triples = nh; m->n_body_triples = nb;
        m->triples = alloc<int>(t.size()); m->parents = alloc<int>(52); m->loss_scratch = alloc<double>(4);
        m->err_flag = alloc<int>(1);
        if (rc != A2M_OK) return;
        cudaMemcpy(m->triples, t.data(), t.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(m->parents, kParents, 52 * 4, cudaMemcpyHostToDevice);
        cudaMemset(m->err_flag, 0, 4);
    }
};

int pow2_at_least(int x) { int p = 1; while (p < x) p <<= 1; return p; }

}  // namespace

// The plan is built in two passes over the same code: pass 0 sizes the arena, pass 1 emits the ops
// with real pointers.
namespace {

struct Bufs {
    size_t off = 0;
    unsigned char* base = nullptr;
    template <typename T>
    T* get(size_t n) {
        const size_t bytes = (n * sizeof(T) + 255) & ~size_t(255);
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += bytes;
        return p;
    }
};

struct Emit {
    a2m_model* m;
    ForwardPlan* P;
    bool dry;
    int rc = A2M_OK;
    std::string tag;                            // label of the ops emitted next (per-op profile)

    // ln: LayerNorm(256) that follows this layer; fused into the epilogue where the tiling allows it (the return value says so)
    bool gemm(const LayerW& L, std::vector<Tap> taps, const AView* views, int n_src, const int box[4], const int ext[4],
              void* out, const long long ostride[4], long long obase, int out_type, int act = -1, int split_k = 1,
              long long split_stride = 0, const LnW* ln = nullptr) {
        if (dry || rc != A2M_OK) return false;
        ConvGemmDesc d;
        d.n_src = n_src;
        for (int s = 0; s < n_src; ++s) d.a[s] = views[s];
        for (int i = 0; i < 4; ++i) { d.box[i] = box[i]; d.m_extent[i] = ext[i]; d.out_stride[i] = ostride[i]; }
        d.taps = std::move(taps);
        d.N = L.N; d.out_base = obase; d.act = act >= 0 ? act : L.act; d.out_type = out_type; d.split_k = split_k; d.split_stride = split_stride;
        {   // layers whose width is a multiple of 256 run 256 x 256 tiles on CTA pairs (the planner falls back to
            // 128 x 256 / 128 x 128 tiles where the output cannot be stored by TMA or there is a single M tile)
            long long m_tiles = 1;
            for (int i = 0; i < 4; ++i) m_tiles *= (ext[i] + box[i] - 1) / box[i];
            if (L.N % 256 == 0 && L.N >= 512 && m_tiles * (L.N / 256) >= 120) d.block_n_hint = 256;
            if (L.N % 256 == 0 && (m_tiles >= 2 || ln != nullptr)) d.block_n_hint = 512;
        }
        if (ln) { d.ln_gamma = ln->w; d.ln_beta = ln->b; }
        auto plan = std::make_shared<ConvGemmPlan>();
        rc = conv_gemm_plan(d, L.w, L.bias, out, plan.get());
        if (rc != A2M_OK) return false;
        P->gemm_flops += plan->flops;
        int* flag = m->err_flag;
        P->ops.push_back([plan, flag](cudaStream_t s) { return conv_gemm_launch(*plan, flag, s); });
        P->op_is_gemm.push_back(1);
        P->op_name.push_back(tag + ".gemm");
        P->op_flops.push_back(plan->flops);
        return plan->ln_fused != 0;
    }
    // rows = (L, B) of a [B, L, C] tensor; taps shift along L
    void conv_rows(const LayerW& L, const __nv_bfloat16* a0, int c0, const __nv_bfloat16* a1, int c1, int len, int B,
                   void* out, long long out_row_stride, long long out_clip_stride, long long obase, int out_type) {
        AView v[2];
        v[0].ptr = a0; v[0].rank = 3; v[0].dims[0] = c0; v[0].dims[1] = len; v[0].dims[2] = B;
        v[0].strides[0] = 1; v[0].strides[1] = c0; v[0].strides[2] = static_cast<long long>(len) * c0;
        if (a1) { v[1] = v[0]; v[1].ptr = a1; v[1].dims[0] = c1; v[1].strides[1] = c1; v[1].strides[2] = static_cast<long long>(len) * c1; }
        const int lt = std::min(128, pow2_at_least(len));
        const int box[4] = {lt, 128 / lt, 1, 1}, ext[4] = {len, B, 1, 1};
        const long long os[4] = {out_row_stride, out_clip_stride, 0, 0};
        gemm(L, L.taps, v, a1 ? 2 : 1, box, ext, out, os, obase, out_type);
    }
    void conv_k3(const LayerW& L, const __nv_bfloat16* a0, int c0, const __nv_bfloat16* a1, int c1, int len, int B,
                 __nv_bfloat16* out) {
        conv_rows(L, a0, c0, a1, c1, len, B, out, L.N, static_cast<long long>(len) * L.N, 0, kOutBf16);
    }
    void conv_k4s2(const LayerW& L, const __nv_bfloat16* a, int C, int len, int B, __nv_bfloat16* out) {
        const int lo = len / 2;
        AView v;
        v.ptr = a; v.rank = 4; v.dims[0] = C; v.dims[1] = 2; v.dims[2] = lo; v.dims[3] = B;
        v.strides[0] = 1; v.strides[1] = C; v.strides[2] = 2LL * C; v.strides[3] = static_cast<long long>(len) * C;
        const int lt = std::min(128, pow2_at_least(lo));
        const int box[4] = {1, lt, 128 / lt, 1}, ext[4] = {1, lo, B, 1};
        const long long os[4] = {0, L.N, static_cast<long long>(lo) * L.N, 0};
        gemm(L, L.taps, &v, 1, box, ext, out, os, 0, kOutBf16);
    }
    // input [B, in_alloc, C] with rows [0, in_len) valid -> rows [0, out_len) of out [B, out_alloc, N]; taps shift along the rows
    void conv_rows_io(const LayerW& L, const __nv_bfloat16* a, int C, int in_len, int in_alloc, int out_len, int out_alloc,
                      int B, __nv_bfloat16* out) {
        AView v;
        v.ptr = a; v.rank = 3; v.dims[0] = C; v.dims[1] = in_len; v.dims[2] = B;
        v.strides[0] = 1; v.strides[1] = C; v.strides[2] = static_cast<long long>(in_alloc) * C;
        const int lt = std::min(128, pow2_at_least(out_len));
        const int box[4] = {lt, 128 / lt, 1, 1}, ext[4] = {out_len, B, 1, 1};
        const long long os[4] = {L.N, static_cast<long long>(out_alloc) * L.N, 0, 0};
        gemm(L, L.taps, &v, 1, box, ext, out, os, 0, kOutBf16);
    }
    // the same for k4 s2 p1: in_alloc is even and rows [in_len, in_alloc) are zero
    void conv_k4s2_io(const LayerW& L, const __nv_bfloat16* a, int C, int in_alloc, int out_len, int out_alloc, int B,
                      __nv_bfloat16* out) {
        AView v;
        v.ptr = a; v.rank = 4; v.dims[0] = C; v.dims[1] = 2; v.dims[2] = in_alloc / 2; v.dims[3] = B;
        v.strides[0] = 1; v.strides[1] = C; v.strides[2] = 2LL * C; v.strides[3] = static_cast<long long>(in_alloc) * C;
        const int lt = std::min(128, pow2_at_least(out_len));
        const int box[4] = {1, lt, 128 / lt, 1}, ext[4] = {1, out_len, B, 1};
        const long long os[4] = {0, L.N, static_cast<long long>(out_alloc) * L.N, 0};
        gemm(L, L.taps, &v, 1, box, ext, out, os, 0, kOutBf16);
    }
    void conv_transpose(const LayerW& even, const LayerW& odd, const __nv_bfloat16* a, int C, int len, int B,
                        __nv_bfloat16* out) {
        const int N = even.N;
        conv_rows(even, a, C, nullptr, 0, len, B, out, 2LL * N, 2LL * len * N, 0, kOutBf16);
        conv_rows(odd, a, C, nullptr, 0, len, B, out, 2LL * N, 2LL * len * N, N, kOutBf16);
    }
    bool linear_rows(const LayerW& L, const __nv_bfloat16* a0, const __nv_bfloat16* a1, int C, long long rows, void* out,
                     long long ldc, long long col, int out_type, const LnW* ln = nullptr) {
        AView v[2];
        v[0].ptr = a0; v[0].rank = 2; v[0].dims[0] = C; v[0].dims[1] = rows; v[0].strides[0] = 1; v[0].strides[1] = C;
        if (a1) { v[1] = v[0]; v[1].ptr = a1; }
        const int box[4] = {128, 1, 1, 1}, ext[4] = {static_cast<int>(rows), 1, 1, 1};
        const long long os[4] = {ldc, 0, 0, 0};
        return gemm(L, L.taps, v, a1 ? 2 : 1, box, ext, out, os, col, out_type, -1, 1, 0, ln);
    }
    void op(std::function<int(cudaStream_t)> f) {
        if (!dry && rc == A2M_OK) {
            P->ops.push_back(std::move(f)); P->op_is_gemm.push_back(0);
            P->op_name.push_back(tag); P->op_flops.push_back(0);
        }
    }

    // SelfAttention followed by ChannelAttention (the hand decoder's order): one kernel when the fused block applies
    bool attention_then_channel(const AttnW& A, const ChanW& W, const __nv_bfloat16* x, int len, int B, __nv_bfloat16* out) {
        if (!attn_fused_supported(len, A.C) || W.C != A.C || W.hidden != 32) return false;
        attention(A, x, nullptr, len, B, nullptr, out, &W);
        return true;
    }
    void attention(const AttnW& A, const __nv_bfloat16* x, const __nv_bfloat16* res2, int len, int B, __nv_bfloat16* qkv,
                   __nv_bfloat16* out, const ChanW* chan = nullptr) {
        const int C = A.C, ld = 2 * (C / 8) + C;
        const float* gamma = A.gamma;
        if (attn_fused_supported(len, C)) {               // decoder blocks: projection + attention in one kernel
            if (dry || rc != A2M_OK) return;
            std::shared_ptr<AttnFusedPlan> ap;
            rc = chan ? attn_fused_plan(A.qkv.w, A.qkv.bias, gamma, x, res2, B, len, C, out, &ap, chan->w0, chan->b0, chan->w2, chan->b2)
                      : attn_fused_plan(A.qkv.w, A.qkv.bias, gamma, x, res2, B, len, C, out, &ap);
            if (rc != A2M_OK) return;
            int* flag = m->err_flag;
            op([ap, flag](cudaStream_t s) { return attn_fused_launch(*ap, flag, s); });
            return;
        }
        linear_rows(A.qkv, x, nullptr, C, static_cast<long long>(B) * len, qkv, ld, 0, kOutBf16);
        if (attn_core_supported(len, C)) {                // wide UNet attentions: S, P.v on the tensor cores
            if (dry || rc != A2M_OK) return;
            std::shared_ptr<AttnCorePlan> cp;
            rc = attn_core_plan(qkv, gamma, x, res2, B, len, C, out, &cp);
            if (rc != A2M_OK) return;
            int* flag = m->err_flag;
            op([cp, flag](cudaStream_t s) { return attn_core_launch(*cp, flag, s); });
            return;
        }
        op([=](cudaStream_t s) { return launch_attention(qkv, x, res2, gamma, B, len, C, out, s); });
    }
    void channel(const ChanW& W, const __nv_bfloat16* x, int len, int B, __nv_bfloat16* out) {
        op([=](cudaStream_t s) { return launch_channel_attention(x, B, len, W.C, W.hidden, W.w0, W.b0, W.w2, W.b2, out, s); });
    }
    // ResBlock (model_layers.py:185-190): x -> conv1 -> conv2 -> attention -> + x
    void resblock(const ResW& R, const __nv_bfloat16* x, int len, int B, __nv_bfloat16* t1, __nv_bfloat16* t2,
                  __nv_bfloat16* qkv, __nv_bfloat16* out) {
        // (a one-kernel ResBlock with the activations resident in shared memory was built and retired:
        // tools/probes/retired/resblock_fused.cu)
        const int C = R.attn.C;
        conv_k3(R.c1, x, C, nullptr, 0, len, B, t1);
        conv_k3(R.c2, t1, C, nullptr, 0, len, B, t2);
        attention(R.attn, t2, x, len, B, qkv, out);
    }
};

int build_plan(a2m_model* m, ForwardPlan* P) {
    const int B = P->B, T = P->T, F = P->F, T_out = P->T_out;
    // three k4 s2 p1 stages: H -> floor(H / 2) (model_layers.py:252-256).  T % 4 == 0 makes H1 and H2 exact; H2 may be odd,
    // then the third stage drops its last row pair like the reference does.  The parity view of a stride-2 stage needs an
    // even number of allocated rows: a1 gets one zero row of padding per clip when H2 is odd.
    const int H1 = T / 2, W1 = F / 2, H2 = T / 4, W2 = F / 4, H3 = H2 / 2, W3 = F / 8;
    const int H2a = H2 + (H2 & 1);
    const int w_out = W3 - 1;                                   // conv 4: W + 2*3 - 8 + 1
    A2M_ARG_CHECK(w_out >= 1 && (w_out % 2) == 1, "model: F = %d gives an even number (%d) of encoder output columns; "
                  "only odd widths (F/8 even) are implemented", F, w_out);
    const int w_centre = (w_out - 1) / 2;
    Bufs bufs;
    for (int pass = 0; pass < 2; ++pass) {
        Emit E{m, P, pass == 0};
        if (pass == 1) {
            P->ops.clear();
            P->op_is_gemm.clear();
            P->op_name.clear();
            P->op_flops.clear();
            P->gemm_flops = 0;
            cudaError_t e = cudaMalloc(&P->arena, bufs.off + 256);
            if (e != cudaSuccess) { a2m_set_error("model: arena cudaMalloc(%zu) failed: %s", bufs.off, cudaGetErrorString(e)); return (int)e; }
            bufs.base = static_cast<unsigned char*>(P->arena);
        }
        bufs.off = 0;
        const size_t BT = static_cast<size_t>(B) * T;
        auto* mel_stage = bufs.get<float>(BT * F);
        auto* a0 = bufs.get<__nv_bfloat16>(static_cast<size_t>(B) * H1 * W1 * 64);
        auto* a1 = bufs.get<__nv_bfloat16>(static_cast<size_t>(B) * H2a * W2 * 128);
        if (pass == 1 && H2a != H2)               // the padding rows are never written: zero them once
            A2M_CUDA_CHECK(cudaMemset(a1, 0, static_cast<size_t>(B) * H2a * W2 * 128 * sizeof(__nv_bfloat16)));
        auto* a2 = bufs.get<__nv_bfloat16>(static_cast<size_t>(B) * H3 * W3 * 256);
        auto* a3 = bufs.get<__nv_bfloat16>(static_cast<size_t>(B) * H3 * W3 * 512);
        auto* a4 = bufs.get<float>(static_cast<size_t>(kConv4Splits) * B * H3 * 256);
        auto* e0 = bufs.get<__nv_bfloat16>(static_cast<size_t>(B) * std::max(T, T_out) * 256);
        auto* s0 = bufs.get<__nv_bfloat16>(BT * 512);
        auto* u1 = bufs.get<__nv_bfloat16>(BT / 2 * 512);
        auto* s1 = bufs.get<__nv_bfloat16>(BT / 2 * 1024);
        auto* u3 = bufs.get<__nv_bfloat16>(BT / 4 * 1024);
        auto* u4 = bufs.get<__nv_bfloat16>(BT / 4 * 2048);
        auto* qkvb = bufs.get<__nv_bfloat16>(BT / 4 * 2560);
        auto* u5 = bufs.get<__nv_bfloat16>(BT / 4 * 2048);
        auto* u6 = bufs.get<__nv_bfloat16>(BT / 2 * 1024);
        auto* qkvu = bufs.get<__nv_bfloat16>(BT / 2 * 1280);
        auto* u7 = bufs.get<__nv_bfloat16>(BT / 2 * 1024);
        auto* u8 = bufs.get<__nv_bfloat16>(BT / 2 * 1024);
        auto* u9 = bufs.get<__nv_bfloat16>(BT * 512);
        auto* u10 = bufs.get<__nv_bfloat16>(BT * 512);
        auto* r = bufs.get<__nv_bfloat16>(BT * 256);
        auto* pose_stage = bufs.get<float>(BT * kPoseFeats);
        P->mel_in = mel_stage; P->enc_out = e0; P->unet_out = r; P->pose_stage = pose_stage;

        // ---------------- AudioEncoder (model_layers.py:267-280) ----------------
        if (m->has_encoder) {
            a2m_model* mm = m;
            E.tag = "enc.conv0";
            E.op([=](cudaStream_t s) {
                return launch_conv0(mm->cur_mel, mm->cur_stride_b, mm->cur_stride_t, B, T, F, mm->conv0_w, mm->conv0_b, a0, s,
                                    mm->cur_n_inner, mm->cur_stride_outer);
            });
            // in: [B, H (allocated rows, even), W, C]; out: [B, Ho_alloc, W / 2, N], rows [0, Ho) written
            auto conv2d_s2 = [&](const LayerW& L, const __nv_bfloat16* in, int H, int W, int C, __nv_bfloat16* out, int Ho, int Ho_alloc) {
                const int Wo = W / 2;
                AView v[2];
                for (int ph = 0; ph < 2; ++ph) {
                    v[ph].ptr = in + static_cast<size_t>(ph) * W * C; v[ph].rank = 5;
                    const long long dims[5] = {C, 2, Wo, H / 2, B}, str[5] = {1, C, 2LL * C, 2LL * W * C, static_cast<long long>(H) * W * C};
                    for (int i = 0; i < 5; ++i) { v[ph].dims[i] = dims[i]; v[ph].strides[i] = str[i]; }
                }
                const int wb = std::min(128, pow2_at_least(Wo)), hb = std::min(128 / wb, pow2_at_least(Ho));
                const int box[4] = {1, wb, hb, 128 / (wb * hb)}, ext[4] = {1, Wo, Ho, B};
                const long long os[4] = {0, L.N, static_cast<long long>(Wo) * L.N, static_cast<long long>(Ho_alloc) * Wo * L.N};
                E.gemm(L, L.taps, v, 2, box, ext, out, os, 0, kOutBf16);
            };
            E.tag = "enc.conv1";
            conv2d_s2(m->enc[1], a0, H1, W1, 64, a1, H2, H2a);
            E.tag = "enc.conv2";
            conv2d_s2(m->enc[2], a1, H2a, W2, 128, a2, H3, H3);
            E.tag = "enc.conv3";
            {   // conv 3
                AView v; v.ptr = a2; v.rank = 4;
                const long long dims[4] = {256, W3, H3, B}, str[4] = {1, 256, 256LL * W3, 256LL * W3 * H3};
                for (int i = 0; i < 4; ++i) { v.dims[i] = dims[i]; v.strides[i] = str[i]; }
                const int wb = std::min(128, pow2_at_least(W3)), hb = std::min(128 / wb, pow2_at_least(H3));
                const int box[4] = {wb, hb, 128 / (wb * hb), 1}, ext[4] = {W3, H3, B, 1};
                const long long os[4] = {512, 512LL * W3, 512LL * W3 * H3, 0};
                E.gemm(m->enc[3], m->enc[3].taps, &v, 1, box, ext, a3, os, 0, kOutBf16);
            }
            E.tag = "enc.conv4";
            {   // conv 4, centre column only -> fp32 [B, H3, 256]
                AView v; v.ptr = a3; v.rank = 4;
                const long long dims[4] = {512, W3, H3, B}, str[4] = {1, 512, 512LL * W3, 512LL * W3 * H3};
                for (int i = 0; i < 4; ++i) { v.dims[i] = dims[i]; v.strides[i] = str[i]; }
                std::vector<Tap> taps = m->enc[4].taps;
                for (Tap& t : taps) t.off[0] += w_centre - 3;
                const int hb = std::min(128, pow2_at_least(H3));
                const int box[4] = {1, hb, 128 / hb, 1}, ext[4] = {1, H3, B, 1};
                const long long os[4] = {0, 256, 256LL * H3, 0};
                // few M tiles (B*H3/128) but K = 12288: K is split over gridDim.z (a fixed count, so results do
                // not depend on the batch size); each split writes its own fp32 plane and the interpolation
                // kernel sums the planes in order, then applies LeakyReLU
                E.gemm(m->enc[4], taps, &v, 1, box, ext, a4, os, 0, kOutF32, kActNone, kConv4Splits,
                       static_cast<long long>(B) * H3 * 256);
            }
            E.tag = "enc.interp";
            E.op([=](cudaStream_t s) { return launch_time_interp(a4, kConv4Splits, B, H3, T_out, 256, e0, s); });
        }
        if (pass == 1) P->enc_end = static_cast<int>(P->ops.size());

        // ---------------- UNet1D (model_layers.py:341-374, D1) ----------------
        if (m->has_unet) {
        E.tag = "unet.ds0";
        E.conv_k3(m->ds[0], e0, 256, nullptr, 0, T, B, s0);
        E.tag = "unet.ds1";
        E.conv_k4s2(m->ds[1], s0, 512, T, B, u1);
        E.tag = "unet.ds2";
        E.conv_k3(m->ds[2], u1, 512, nullptr, 0, T / 2, B, s1);
        E.tag = "unet.ds3";
        E.conv_k4s2(m->ds[3], s1, 1024, T / 2, B, u3);
        E.tag = "unet.bottleneck";
        E.conv_k3(m->bott, u3, 1024, nullptr, 0, T / 4, B, u4);
        E.tag = "unet.bott_attn";
        E.attention(m->bott_attn, u4, nullptr, T / 4, B, qkvb, u5);
        E.tag = "unet.up0";
        E.conv_transpose(m->up0_even, m->up0_odd, u5, 2048, T / 4, B, u6);
        E.tag = "unet.up_attn";
        E.attention(m->up_attn, u6, nullptr, T / 2, B, qkvu, u7);            // D1: before the concat
        E.tag = "unet.up1";
        E.conv_k3(m->up1, u7, 1024, s1, 1024, T / 2, B, u8);
        E.tag = "unet.up2";
        E.conv_transpose(m->up2_even, m->up2_odd, u8, 1024, T / 2, B, u9);
        E.tag = "unet.up3";
        E.conv_k3(m->up3, u9, 512, s0, 512, T, B, u10);
        E.tag = "unet.final";
        E.linear_rows(m->final_conv, u10, nullptr, 512, static_cast<long long>(BT), r, 256, 0, kOutBf16);
        }
        if (pass == 1) P->unet_end = static_cast<int>(P->ops.size());

        // ---------------- decoders (real_motion_model.py:160-262) ----------------
        for (int part = 0; part < 2 && m->has_decoders; ++part) {
            const DecoderW& D = m->dec[part];
            const int J = D.joints;
            const size_t nodes = BT * J;
            auto* t1 = bufs.get<__nv_bfloat16>(BT * 256);
            auto* t2 = bufs.get<__nv_bfloat16>(BT * 256);
            auto* t3 = bufs.get<__nv_bfloat16>(BT * 256);
            auto* t4 = bufs.get<__nv_bfloat16>(BT * 256);
            auto* qkv = bufs.get<__nv_bfloat16>(BT * 320);
            auto* xa = bufs.get<__nv_bfloat16>(nodes * 64);
            auto* xb = bufs.get<__nv_bfloat16>(nodes * 64);
            const std::string dn = part == 0 ? "body" : "hand";
            E.tag = dn + ".pre_res";
            E.resblock(D.pre_res, r, T, B, t1, t2, qkv, t3);
            E.tag = dn + ".pre_conv";
            E.conv_k3(D.pre_conv, t3, 256, nullptr, 0, T, B, t1);
            E.tag = dn + ".pre_attn_chan";
            if (D.chan_first) {
                E.channel(D.pre_chan, t1, T, B, t2);
                E.attention(D.pre_attn, t2, nullptr, T, B, qkv, t3);
            } else {
                if (!E.attention_then_channel(D.pre_attn, D.pre_chan, t1, T, B, t3)) {
                    E.attention(D.pre_attn, t1, nullptr, T, B, qkv, t2);
                    E.channel(D.pre_chan, t2, T, B, t3);
                }
            }
            E.tag = dn + ".proj_in";
            E.linear_rows(D.proj_in, t3, nullptr, 256, static_cast<long long>(BT), xa, static_cast<long long>(J) * 64, 0, kOutBf16);
            E.tag = dn + ".gnn";
            GraphTopo topo{J, D.nbr, D.deg};
            if (!E.dry && E.rc == A2M_OK) {
                GnnFusedWeights gw;
                for (int i = 0; i < 3; ++i) {
                    gw.gat_w[i] = D.gat[i].lin.w; gw.gat_bias[i] = D.gat[i].bias;
                }
                for (int i = 0; i < 2; ++i) { gw.gc_w[i] = D.gconv[i].w; gw.gc_bias[i] = D.gconv[i].bias; }
                for (int i = 0; i < 5; ++i) { gw.ln_w[i] = D.ln64[i].w; gw.ln_b[i] = D.ln64[i].b; }
                std::shared_ptr<GnnFusedPlan> gp;
                E.rc = gnn_fused_plan(gw, topo, B, T, xa, xb, &gp);      // tiles never straddle clips
                int* flag = m->err_flag;
                if (E.rc == A2M_OK) E.op([gp, flag](cudaStream_t s) { return gnn_fused_launch(*gp, flag, s); });
            }
            __nv_bfloat16* cur = xb;
            E.tag = dn + ".proj_out";
            const LnW nw = D.norm;                       // LayerNorm(256) in proj_out's epilogue where its tiling holds whole rows
            if (!E.linear_rows(D.proj_out, cur, nullptr, J * 64, static_cast<long long>(BT), t2, 256, 0, kOutBf16, &nw)) {
                E.tag = dn + ".norm";
                E.op([=](cudaStream_t s) { return launch_layernorm(t2, static_cast<long long>(BT), 256, nw.w, nw.b, t2, s); });
            }
            E.tag = dn + ".post_res";
            E.resblock(D.post_res, t2, T, B, t1, t3, qkv, t4);
            E.tag = dn + ".post_conv";
            E.conv_k3(D.post_conv, t4, 256, nullptr, 0, T, B, t1);
            E.tag = dn + ".post_attn";
            const __nv_bfloat16* last = t2;
            if (D.has_post_chan && E.attention_then_channel(D.post_attn, D.post_chan, t1, T, B, t3)) {
                last = t3;
            } else {
                E.attention(D.post_attn, t1, nullptr, T, B, qkv, t2);
                if (D.has_post_chan) { E.channel(D.post_chan, t2, T, B, t3); last = t3; }
            }
            E.tag = dn + ".logits";
            // logits: fp32 straight into pose[B, T, 104] at this branch's column block
            E.linear_rows(D.logits, last, nullptr, 256, static_cast<long long>(BT), pose_stage, kPoseFeats,
                          part == 0 ? 0 : kBodyFeats, kOutF32);
            if (pass == 1 && part == 0) P->body_end = static_cast<int>(P->ops.size());
        }
        if (E.rc != A2M_OK) return E.rc;
    }
    return A2M_OK;
}

ForwardPlan* get_plan(a2m_model* m, int B, int T, int F, int* rc_out, int T_out = 0) {
    if (T_out <= 0) T_out = T;
    char key[80];
    snprintf(key, sizeof(key), "%d:%d:%d:%d", B, T, F, T_out);
    auto it = m->plans.find(key);
    if (it != m->plans.end()) { *rc_out = A2M_OK; it->second->last_use = ++m->plan_clock; return it->second.get(); }
    std::unique_ptr<ForwardPlan> P(new ForwardPlan());
    P->B = B; P->T = T; P->F = F; P->T_out = T_out;
    *rc_out = build_plan(m, P.get());
    if (*rc_out != A2M_OK) return nullptr;
    ForwardPlan* raw = P.get();
    raw->last_use = ++m->plan_clock;
    // bound the cache: evict the least recently used shape only (a CUDA graph captured over a plan bakes its arena
    // pointers in: a2m_model_plan_generation lets the caller notice that one of its shapes was evicted)
    if (m->plans.size() >= 16) {
        auto victim = m->plans.begin();
        for (auto jt = m->plans.begin(); jt != m->plans.end(); ++jt)
            if (jt->second->last_use < victim->second->last_use) victim = jt;
        m->plans.erase(victim);
        ++m->plan_generation;
    }
    m->plans[key] = std::move(P);
    return raw;
}

int check_shape(long long B, int T, int F, const char* who) {
    A2M_ARG_CHECK(B >= 1 && B <= 65535, "%s: batch %lld out of range [1, 65535]", who, (long long)B);
    A2M_ARG_CHECK(T >= 8 && T % 4 == 0, "%s: T = %d; the time axis must be a multiple of 4 (the reference's UNet skip "
                  "concats, model_layers.py:341-374) and at least 8 (three stride-2 encoder stages)", who, T);
    A2M_ARG_CHECK(T <= 4096, "%s: T = %d; this build implements T <= 4096", who, T);
    A2M_ARG_CHECK(F >= 16 && F % 8 == 0, "%s: F = %d; the mel axis must be a multiple of 8 and >= 16", who, F);
    return A2M_OK;
}

int run_ops(ForwardPlan* P, int lo, int hi, cudaStream_t s, a2m_model* m = nullptr) {
    const bool rec = m != nullptr && m->tl_steps < m->tl_capacity && m->tl_ops == static_cast<int>(P->ops.size());
    for (int i = lo; i < hi; ++i) {
        const int rc = P->ops[i](s);
        if (rc != A2M_OK) return rc;
        if (rec) A2M_CUDA_CHECK(cudaEventRecord(m->tl_events[static_cast<size_t>(m->tl_steps) * (m->tl_ops + 1) + 1 + i], s));
    }
    return A2M_OK;
}

cudaEvent_t timeline_reference(int device) {          // one reference event per device, shared by every handle (lanes)
    static cudaEvent_t ref[64] = {};
    if (device < 0 || device >= 64) return nullptr;
    if (!ref[device]) {
        if (cudaEventCreate(&ref[device]) != cudaSuccess) return nullptr;
        cudaEventRecord(ref[device], 0);
        cudaEventSynchronize(ref[device]);
    }
    return ref[device];
}

}  // namespace

extern "C" int a2m_model_create(const a2m_tensor_desc* tensors, int n_tensors, int device, a2m_model** out) {
    A2M_ARG_CHECK(out != nullptr && tensors != nullptr && n_tensors > 0, "a2m_model_create: NULL argument");
    *out = nullptr;
    A2M_CUDA_CHECK(cudaSetDevice(device));
    std::unique_ptr<a2m_model> m(new a2m_model());
    m->device = device;
    for (int i = 0; i < n_tensors; ++i) {
        const a2m_tensor_desc& t = tensors[i];
        A2M_ARG_CHECK(t.name != nullptr && t.ndim >= 0 && t.ndim <= 4, "a2m_model_create: bad descriptor %d", i);
        Param p;
        p.ndim = t.ndim;
        for (int k = 0; k < t.ndim; ++k) p.shape[k] = t.shape[k];
        if (t.dtype == A2M_DTYPE_F32) p.f32 = static_cast<const float*>(t.data);
        else if (t.dtype == A2M_DTYPE_I64) p.i64 = static_cast<const long long*>(t.data);
        else { a2m_set_error("a2m_model_create: tensor '%s' has unsupported dtype %d", t.name, t.dtype); return A2M_ERR_ARGUMENT; }
        m->params[t.name] = p;
    }
    Builder b{m.get(), nullptr};
    // a full SelfAttention_G state_dict has all three sections; the standalone AudioEncoder / UNet1D
    // drop-in modules pass only their own
    m->has_encoder = m->params.count("audio_encoder.conv.0.conv.weight") > 0;
    m->has_unet = m->params.count("unet.final_conv.weight") > 0;
    m->has_decoders = m->params.count("body_logits.weight") > 0;
    A2M_ARG_CHECK(m->has_encoder || m->has_unet || m->has_decoders, "a2m_model_create: no known section in the state_dict");
    if (m->has_encoder) b.encoder();
    if (m->has_unet) b.unet();
    if (m->has_decoders) {
        b.decoder(m->dec[0], "body", kBodyJoints, kBodyFeats);
        b.decoder(m->dec[1], "hand", kHandJoints, kPoseFeats - kBodyFeats);
    }
    b.losses();
    cudaError_t e = cudaDeviceSynchronize();
    if (b.rc == A2M_OK && e != cudaSuccess) { a2m_set_error("a2m_model_create: %s", cudaGetErrorString(e)); b.rc = (int)e; }
    if (b.rc != A2M_OK) {
        for (void* p : m->owned) cudaFree(p);
        return b.rc;
    }
    m->params.clear();                       // the caller's tensors are not referenced after this point
    A2M_CUDA_CHECK(cudaStreamCreateWithFlags(&m->side_stream, cudaStreamNonBlocking));
    A2M_CUDA_CHECK(cudaEventCreateWithFlags(&m->ev_fork, cudaEventDisableTiming));
    A2M_CUDA_CHECK(cudaEventCreateWithFlags(&m->ev_join, cudaEventDisableTiming));
    *out = m.release();
    return A2M_OK;
}

extern "C" void a2m_model_destroy(a2m_model* m) {
    if (!m) return;
    cudaSetDevice(m->device);
    cudaDeviceSynchronize();
    m->plans.clear();
    for (void* p : m->owned) cudaFree(p);
    for (cudaEvent_t e : m->tl_events) cudaEventDestroy(e);
    if (m->side_stream) cudaStreamDestroy(m->side_stream);
    if (m->ev_fork) cudaEventDestroy(m->ev_fork);
    if (m->ev_join) cudaEventDestroy(m->ev_join);
    delete m;
}

extern "C" int a2m_model_forward(a2m_model* m, const float* mel, int64_t mel_stride_b, int64_t mel_stride_t, int64_t B,
                                 int T, int F, float* pose, float* losses, const float* real_pose, void* stream) {
    A2M_ARG_CHECK(m != nullptr && mel != nullptr && pose != nullptr, "a2m_model_forward: NULL argument");
    A2M_ARG_CHECK(m->has_encoder && m->has_unet && m->has_decoders, "a2m_model_forward: the handle was created from a partial state_dict");
    int rc = check_shape(B, T, F, "a2m_model_forward");
    if (rc != A2M_OK) return rc;
    ForwardPlan* P = get_plan(m, static_cast<int>(B), T, F, &rc);
    if (!P) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    A2M_ARG_CHECK(mel_stride_t >= F && (B == 1 || mel_stride_b >= mel_stride_t), "a2m_model_forward: mel strides (%lld, %lld)",
                  (long long)mel_stride_b, (long long)mel_stride_t);
    m->cur_mel = mel; m->cur_stride_b = mel_stride_b; m->cur_stride_t = mel_stride_t;
    m->cur_n_inner = m->next_n_inner; m->cur_stride_outer = m->next_stride_outer;
    m->next_n_inner = 0; m->next_stride_outer = 0;
    // trunk on the caller's stream, then body decoder (side stream) || hand decoder (caller's stream)
    const bool timeline = m->tl_steps < m->tl_capacity && m->tl_ops == static_cast<int>(P->ops.size());
    if (timeline) A2M_CUDA_CHECK(cudaEventRecord(m->tl_events[static_cast<size_t>(m->tl_steps) * (m->tl_ops + 1)], s));    // step start
    rc = run_ops(P, 0, P->unet_end, s, m);
    if (rc != A2M_OK) return rc;
    A2M_CUDA_CHECK(cudaEventRecord(m->ev_fork, s));
    A2M_CUDA_CHECK(cudaStreamWaitEvent(m->side_stream, m->ev_fork, 0));
    rc = run_ops(P, P->unet_end, P->body_end, m->side_stream, m);
    if (rc != A2M_OK) return rc;
    A2M_CUDA_CHECK(cudaEventRecord(m->ev_join, m->side_stream));
    rc = run_ops(P, P->body_end, static_cast<int>(P->ops.size()), s, m);
    if (rc != A2M_OK) return rc;
    A2M_CUDA_CHECK(cudaStreamWaitEvent(s, m->ev_join, 0));
    if (timeline) ++m->tl_steps;
    float* stage = P->pose_stage;
    if (m->denorm_on) {     // x * std + mean (generate_motion_video.py:259-260) instead of the plain copy out of the arena
        rc = a2m_pose_denormalize_f32(stage, m->denorm, m->denorm + kPoseFeats, static_cast<int64_t>(B) * T, pose, stream);
        if (rc != A2M_OK) return rc;
    } else {
        A2M_CUDA_CHECK(cudaMemcpyAsync(pose, stage, static_cast<size_t>(B) * T * kPoseFeats * sizeof(float), cudaMemcpyDeviceToDevice, s));
    }
    if (losses) {
        rc = launch_pose_losses(stage, real_pose, static_cast<int>(B), T, m->triples, m->n_hand_triples, m->n_body_triples,
                                m->parents, m->loss_scratch, losses, s);
        if (rc != A2M_OK) return rc;
    }
    return A2M_OK;
}

// Sliding-window generation over long streams (BASELINE config 4; the reference's window arithmetic, dataUtils.py:585-620,
// 648-654): clip (s, w) reads log-mel rows start_w + t * stride_t of stream s IN PLACE -- one launch program for all
// n_streams * n_windows clips, no window is ever gathered.
extern "C" int a2m_model_forward_windows(a2m_model* m, const float* mel, int64_t stride_stream, int64_t stride_window,
                                         int64_t stride_t, int64_t n_streams, int64_t n_windows, int T, int F, float* pose,
                                         float* losses, void* stream) {
    A2M_ARG_CHECK(m != nullptr, "a2m_model_forward_windows: NULL model");
    A2M_ARG_CHECK(n_streams >= 1 && n_windows >= 1 && n_streams * n_windows <= 65535, "a2m_model_forward_windows: %lld streams x "
                  "%lld windows (at most 65535 clips per call)", (long long)n_streams, (long long)n_windows);
    A2M_ARG_CHECK(stride_window >= 1 && stride_stream >= 1 && stride_t >= F, "a2m_model_forward_windows: strides (%lld, %lld, %lld)",
                  (long long)stride_stream, (long long)stride_window, (long long)stride_t);
    A2M_ARG_CHECK(stride_window >= stride_t, "a2m_model_forward_windows: window stride %lld below the row stride %lld",
                  (long long)stride_window, (long long)stride_t);
    m->next_n_inner = static_cast<int>(n_windows);
    m->next_stride_outer = stride_stream;
    const int rc = a2m_model_forward(m, mel, stride_window, stride_t, n_streams * n_windows, T, F, pose, losses, nullptr, stream);
    m->next_n_inner = 0;
    return rc;
}

// Timeline of the next `steps` forwards of shape (B, T, F): an event after every op (and one at the start of each
// forward), read back as milliseconds since a per-device reference event shared by all handles, so the lanes of a
// pipeline line up.  Diagnostic (the events add small gaps); there is no nsys in this environment.
extern "C" int a2m_model_timeline_begin(a2m_model* m, int64_t B, int T, int F, int steps) {
    A2M_ARG_CHECK(m != nullptr && steps >= 1 && steps <= 64, "a2m_model_timeline_begin: bad argument");
    int rc = check_shape(B, T, F, "a2m_model_timeline_begin");
    if (rc != A2M_OK) return rc;
    ForwardPlan* P = get_plan(m, static_cast<int>(B), T, F, &rc);
    if (!P) return rc;
    A2M_CUDA_CHECK(cudaSetDevice(m->device));
    A2M_ARG_CHECK(timeline_reference(m->device) != nullptr, "a2m_model_timeline_begin: no reference event");
    for (cudaEvent_t e : m->tl_events) cudaEventDestroy(e);
    m->tl_ops = static_cast<int>(P->ops.size());
    m->tl_events.assign(static_cast<size_t>(steps) * (m->tl_ops + 1), nullptr);
    for (auto& e : m->tl_events) A2M_CUDA_CHECK(cudaEventCreate(&e));
    m->tl_capacity = steps;
    m->tl_steps = 0;
    return A2M_OK;
}

// out_ms_host: [recorded steps][n_ops + 1] (column 0 = the forward's start); returns the layout through the pointers.
// Synchronises the device.  Ops [0, unet_end) and [body_end, n_ops) run on the caller's stream, [unet_end, body_end) on
// the handle's side stream.
extern "C" int a2m_model_timeline_read(a2m_model* m, float* out_ms_host, int capacity, int* steps_host, int* n_ops_host,
                                       int* unet_end_host, int* body_end_host, int64_t B, int T, int F) {
    A2M_ARG_CHECK(m != nullptr && steps_host && n_ops_host && unet_end_host && body_end_host, "a2m_model_timeline_read: NULL");
    int rc = check_shape(B, T, F, "a2m_model_timeline_read");
    if (rc != A2M_OK) return rc;
    ForwardPlan* P = get_plan(m, static_cast<int>(B), T, F, &rc);
    if (!P) return rc;
    A2M_CUDA_CHECK(cudaSetDevice(m->device));
    A2M_CUDA_CHECK(cudaDeviceSynchronize());
    *steps_host = m->tl_steps; *n_ops_host = m->tl_ops; *unet_end_host = P->unet_end; *body_end_host = P->body_end;
    cudaEvent_t ref = timeline_reference(m->device);
    const size_t n = static_cast<size_t>(m->tl_steps) * (m->tl_ops + 1);
    for (size_t i = 0; i < n && static_cast<int>(i) < capacity; ++i) {
        float ms = 0.f;
        A2M_CUDA_CHECK(cudaEventElapsedTime(&ms, ref, m->tl_events[i]));
        out_ms_host[i] = ms;
    }
    m->tl_capacity = 0;                                 // recording stops; the events stay until the next begin / destroy
    return A2M_OK;
}

extern "C" int a2m_model_set_output_denorm(a2m_model* m, const float* mean, const float* stdv, void* stream) {
    A2M_ARG_CHECK(m != nullptr, "a2m_model_set_output_denorm: NULL model");
    A2M_ARG_CHECK((mean == nullptr) == (stdv == nullptr), "a2m_model_set_output_denorm: give both mean and std, or neither");
    if (mean == nullptr) { m->denorm_on = false; return A2M_OK; }
    A2M_CUDA_CHECK(cudaSetDevice(m->device));
    if (!m->denorm) {
        A2M_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&m->denorm), 2 * kPoseFeats * sizeof(float)));
        m->owned.push_back(m->denorm);
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    A2M_CUDA_CHECK(cudaMemcpyAsync(m->denorm, mean, kPoseFeats * sizeof(float), cudaMemcpyDeviceToDevice, s));
    A2M_CUDA_CHECK(cudaMemcpyAsync(m->denorm + kPoseFeats, stdv, kPoseFeats * sizeof(float), cudaMemcpyDeviceToDevice, s));
    m->denorm_on = true;
    return A2M_OK;
}

extern "C" int a2m_model_encoder_forward_ex(a2m_model* m, const float* mel, int64_t B, int T, int F, int time_steps,
                                            float* out_nct, void* stream) {
    A2M_ARG_CHECK(m != nullptr && mel != nullptr && out_nct != nullptr, "a2m_model_encoder_forward: NULL argument");
    A2M_ARG_CHECK(m->has_encoder, "a2m_model_encoder_forward: no audio_encoder.* tensors in the state_dict");
    int rc = check_shape(B, T, F, "a2m_model_encoder_forward");
    if (rc != A2M_OK) return rc;
    if (time_steps <= 0) time_steps = T;
    A2M_ARG_CHECK(time_steps <= 65536, "a2m_model_encoder_forward: time_steps %d", time_steps);
    ForwardPlan* P = get_plan(m, static_cast<int>(B), T, F, &rc, time_steps);
    if (!P) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    m->cur_mel = mel; m->cur_stride_b = static_cast<long long>(T) * F; m->cur_stride_t = F;
    m->cur_n_inner = 0; m->cur_stride_outer = 0;
    rc = run_ops(P, 0, P->enc_end, s);
    if (rc != A2M_OK) return rc;
    return launch_btc_to_ncw(P->enc_out, static_cast<int>(B), 256, time_steps, out_nct, s);
}

extern "C" int a2m_model_encoder_forward(a2m_model* m, const float* mel, int64_t B, int T, int F, float* out_nct, void* stream) {
    return a2m_model_encoder_forward_ex(m, mel, B, T, F, T, out_nct, stream);
}

extern "C" int64_t a2m_model_plan_generation(a2m_model* m) { return m ? m->plan_generation : -1; }

extern "C" int a2m_model_unet_forward(a2m_model* m, const float* x_nct, int64_t B, int T, float* out_nct, void* stream) {
    A2M_ARG_CHECK(m != nullptr && x_nct != nullptr && out_nct != nullptr, "a2m_model_unet_forward: NULL argument");
    A2M_ARG_CHECK(m->has_unet, "a2m_model_unet_forward: no unet.* tensors in the state_dict");
    int rc = check_shape(B, T, 64, "a2m_model_unet_forward");
    if (rc != A2M_OK) return rc;
    ForwardPlan* P = get_plan(m, static_cast<int>(B), T, 64, &rc);
    if (!P) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    rc = launch_ncw_to_btc(x_nct, static_cast<int>(B), 256, T, P->enc_out, s);
    if (rc != A2M_OK) return rc;
    rc = run_ops(P, P->enc_end, P->unet_end, s);
    if (rc != A2M_OK) return rc;
    return launch_btc_to_ncw(P->unet_out, static_cast<int>(B), 256, T, out_nct, s);
}

// Unit-test / diagnostic surface of the fused graph stack: packs nothing (the handle's weights are used),
// converts, launches and synchronises.
extern "C" int a2m_model_gnn_forward(a2m_model* m, int part, const float* x, int64_t n_graphs, float* out, void* stream) {
    A2M_ARG_CHECK(m != nullptr && x != nullptr && out != nullptr, "a2m_model_gnn_forward: NULL argument");
    A2M_ARG_CHECK(m->has_decoders, "a2m_model_gnn_forward: no decoder tensors in the state_dict");
    A2M_ARG_CHECK(part == 0 || part == 1, "a2m_model_gnn_forward: part %d (0 = body, 1 = hand)", part);
    A2M_ARG_CHECK(n_graphs >= 1 && n_graphs <= (1 << 24), "a2m_model_gnn_forward: %lld graphs", (long long)n_graphs);
    const DecoderW& D = m->dec[part];
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const long long n = n_graphs * D.joints * kJointFeat;
    __nv_bfloat16* buf = nullptr;
    A2M_CUDA_CHECK(cudaMalloc(&buf, static_cast<size_t>(2 * n) * sizeof(__nv_bfloat16) + 512));
    GnnFusedWeights gw;
    for (int i = 0; i < 3; ++i) { gw.gat_w[i] = D.gat[i].lin.w; gw.gat_bias[i] = D.gat[i].bias; }
    for (int i = 0; i < 2; ++i) { gw.gc_w[i] = D.gconv[i].w; gw.gc_bias[i] = D.gconv[i].bias; }
    for (int i = 0; i < 5; ++i) { gw.ln_w[i] = D.ln64[i].w; gw.ln_b[i] = D.ln64[i].b; }
    __nv_bfloat16* xo = buf + ((n + 127) & ~127LL);
    std::shared_ptr<GnnFusedPlan> gp;
    int rc = launch_f32_to_bf16(x, n, buf, s);
    if (rc == A2M_OK) rc = gnn_fused_plan(gw, GraphTopo{D.joints, D.nbr, D.deg}, 1, static_cast<int>(n_graphs), buf, xo, &gp);
    if (rc == A2M_OK) rc = gnn_fused_launch(*gp, m->err_flag, s);
    if (rc == A2M_OK) rc = launch_bf16_to_f32(xo, n, out, s);
    cudaError_t e = cudaStreamSynchronize(s);
    cudaFree(buf);
    if (rc != A2M_OK) return rc;
    if (e != cudaSuccess) { a2m_set_error("a2m_model_gnn_forward: %s", cudaGetErrorString(e)); return (int)e; }
    return a2m_model_status(m);
}

extern "C" int a2m_model_status(a2m_model* m) {
    A2M_ARG_CHECK(m != nullptr, "a2m_model_status: NULL");
    int flag = 0;
    A2M_CUDA_CHECK(cudaDeviceSynchronize());
    A2M_CUDA_CHECK(cudaMemcpy(&flag, m->err_flag, sizeof(int), cudaMemcpyDeviceToHost));
    if (flag != 0) {
        cudaMemset(m->err_flag, 0, sizeof(int));
        a2m_set_error("a2m_model_status: a conv_gemm pipeline barrier wait expired (role %d)", flag);
        return A2M_ERR_PIPELINE;
    }
    return A2M_OK;
}

extern "C" int64_t a2m_model_gemm_flops(a2m_model* m, int64_t B, int T, int F) {
    if (!m) return -1;
    int rc = A2M_OK;
    if (check_shape(B, T, F, "a2m_model_gemm_flops") != A2M_OK) return -1;
    ForwardPlan* P = get_plan(m, static_cast<int>(B), T, F, &rc);
    return P ? P->gemm_flops : -1;
}

// Per-op CUDA-event timing of the forward program (bench.py's roofline leg): every launch is
// bracketed by events on `stream`; per_op (nullable, n entries) receives each launch's average ms.
namespace {
int profile_impl(a2m_model* m, const float* mel, int64_t mel_stride_b, int64_t mel_stride_t, int64_t B, int T, int F,
                 int iters, ForwardPlan** plan_out, std::vector<double>* per_op, cudaStream_t s) {
    A2M_ARG_CHECK(m != nullptr && mel != nullptr && iters >= 1, "a2m_model_profile: bad argument");
    A2M_ARG_CHECK(m->has_encoder && m->has_unet && m->has_decoders, "a2m_model_profile: partial handle");
    int rc = check_shape(B, T, F, "a2m_model_profile");
    if (rc != A2M_OK) return rc;
    ForwardPlan* P = get_plan(m, static_cast<int>(B), T, F, &rc);
    if (!P) return rc;
    m->cur_mel = mel; m->cur_stride_b = mel_stride_b; m->cur_stride_t = mel_stride_t;
    m->cur_n_inner = 0; m->cur_stride_outer = 0;
    const int n = static_cast<int>(P->ops.size());
    std::vector<cudaEvent_t> ev(2 * n);
    for (auto& e : ev) A2M_CUDA_CHECK(cudaEventCreate(&e));
    per_op->assign(n, 0.0);
    for (int it = 0; it < iters && rc == A2M_OK; ++it) {
        for (int i = 0; i < n && rc == A2M_OK; ++i) {
            cudaEventRecord(ev[2 * i], s);
            rc = P->ops[i](s);
            cudaEventRecord(ev[2 * i + 1], s);
        }
        cudaError_t e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) { a2m_set_error("a2m_model_profile: %s", cudaGetErrorString(e)); rc = (int)e; break; }
        for (int i = 0; i < n; ++i) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, ev[2 * i], ev[2 * i + 1]);
            (*per_op)[i] += ms / iters;
        }
    }
    for (auto& e : ev) cudaEventDestroy(e);
    *plan_out = P;
    return rc;
}
}  // namespace

extern "C" int a2m_model_profile(a2m_model* m, const float* mel, int64_t mel_stride_b, int64_t mel_stride_t, int64_t B, int T,
                                 int F, int iters, float* out_ms_host, int* n_gemm_host, void* stream) {
    A2M_ARG_CHECK(out_ms_host != nullptr, "a2m_model_profile: NULL output");
    ForwardPlan* P = nullptr;
    std::vector<double> per_op;
    const int rc = profile_impl(m, mel, mel_stride_b, mel_stride_t, B, T, F, iters, &P, &per_op, static_cast<cudaStream_t>(stream));
    if (rc != A2M_OK) return rc;
    double gemm = 0.0, other = 0.0;
    int n_gemm = 0;
    for (size_t i = 0; i < per_op.size(); ++i) {
        (P->op_is_gemm[i] ? gemm : other) += per_op[i];
        n_gemm += P->op_is_gemm[i];
    }
    out_ms_host[0] = static_cast<float>(gemm + other);
    out_ms_host[1] = static_cast<float>(gemm);
    out_ms_host[2] = static_cast<float>(other);
    if (n_gemm_host) *n_gemm_host = n_gemm;
    return A2M_OK;
}

extern "C" int a2m_model_profile_ops(a2m_model* m, const float* mel, int64_t mel_stride_b, int64_t mel_stride_t, int64_t B,
                                     int T, int F, int iters, float* out_ms_host, int capacity, int* n_ops_host, void* stream) {
    A2M_ARG_CHECK(out_ms_host != nullptr && n_ops_host != nullptr && capacity >= 1, "a2m_model_profile_ops: bad argument");
    ForwardPlan* P = nullptr;
    std::vector<double> per_op;
    const int rc = profile_impl(m, mel, mel_stride_b, mel_stride_t, B, T, F, iters, &P, &per_op, static_cast<cudaStream_t>(stream));
    if (rc != A2M_OK) return rc;
    *n_ops_host = static_cast<int>(per_op.size());
    for (int i = 0; i < capacity && i < *n_ops_host; ++i) out_ms_host[i] = static_cast<float>(per_op[i]);
    return A2M_OK;
}

extern "C" const char* a2m_model_op_name(a2m_model* m, int64_t B, int T, int F, int index, int64_t* flops_host) {
    if (!m || check_shape(B, T, F, "a2m_model_op_name") != A2M_OK) return nullptr;
    int rc = A2M_OK;
    ForwardPlan* P = get_plan(m, static_cast<int>(B), T, F, &rc);
    if (!P || index < 0 || index >= static_cast<int>(P->op_name.size())) return nullptr;
    if (flops_host) *flops_host = P->op_flops[index];
    return P->op_name[index].c_str();
}


// ---------------------------------------------------------------------------------------------------------------
// Stand-alone building blocks: the layer classes of model_layers.py with their own forward (ConvNormRelu :112-118,
// SelfAttention :133-146, ChannelAttention :167-174, ResBlock :185-190, ConvTranspose1D :211-215), on the same
// kernels the generator uses.  x, out: fp32 [B, C, T] (the reference's NCW layout).
// ---------------------------------------------------------------------------------------------------------------
namespace {

int block_out_length(int kind, int T) {
    return kind == A2M_BLOCK_CONV_K4S2 ? T / 2 : kind == A2M_BLOCK_CONV_TRANSPOSE ? 2 * T : T;
}

int build_block_plan(a2m_model* m, ForwardPlan* P) {
    const int B = P->B, T = P->T, kind = m->blk_kind, cin = m->blk_cin, cout = m->blk_cout;
    const int To = block_out_length(kind, T);
    Bufs bufs;
    for (int pass = 0; pass < 2; ++pass) {
        Emit E{m, P, pass == 0};
        if (pass == 1) {
            P->ops.clear(); P->op_is_gemm.clear(); P->op_name.clear(); P->op_flops.clear();
            P->gemm_flops = 0;
            cudaError_t e = cudaMalloc(&P->arena, bufs.off + 256);
            if (e != cudaSuccess) { a2m_set_error("block: arena cudaMalloc(%zu) failed: %s", bufs.off, cudaGetErrorString(e)); return (int)e; }
            bufs.base = static_cast<unsigned char*>(P->arena);
        }
        bufs.off = 0;
        const size_t BT = static_cast<size_t>(B) * T;
        auto* xin = bufs.get<__nv_bfloat16>(BT * cin);
        auto* xout = bufs.get<__nv_bfloat16>(static_cast<size_t>(B) * To * cout);
        auto* t1 = bufs.get<__nv_bfloat16>(BT * cin);
        auto* t2 = bufs.get<__nv_bfloat16>(BT * cin);
        auto* qkv = bufs.get<__nv_bfloat16>(BT * (cin + 2 * (cin / 8)));
        P->enc_out = xin; P->unet_out = xout;
        E.tag = "block";
        switch (kind) {
            case A2M_BLOCK_CONV_K3: E.conv_k3(m->blk_conv, xin, cin, nullptr, 0, T, B, xout); break;
            case A2M_BLOCK_CONV_K4S2: E.conv_k4s2(m->blk_conv, xin, cin, T, B, xout); break;
            case A2M_BLOCK_CONV_TRANSPOSE: E.conv_transpose(m->blk_even, m->blk_odd, xin, cin, T, B, xout); break;
            case A2M_BLOCK_SELF_ATTENTION: E.attention(m->blk_attn, xin, nullptr, T, B, qkv, xout); break;
            case A2M_BLOCK_CHANNEL_ATTENTION: E.channel(m->blk_chan, xin, T, B, xout); break;
            case A2M_BLOCK_RESBLOCK: E.resblock(m->blk_res, xin, T, B, t1, t2, qkv, xout); break;
            default: a2m_set_error("block: kind %d", kind); return A2M_ERR_ARGUMENT;
        }
        if (E.rc != A2M_OK) return E.rc;
    }
    return A2M_OK;
}

}  // namespace

extern "C" int a2m_block_create(int kind, const a2m_tensor_desc* tensors, int n_tensors, int in_channels, int out_channels,
                                int leaky, int device, a2m_model** out) {
    A2M_ARG_CHECK(out != nullptr && tensors != nullptr && n_tensors > 0, "a2m_block_create: NULL argument");
    *out = nullptr;
    A2M_ARG_CHECK(kind >= A2M_BLOCK_CONV_K3 && kind <= A2M_BLOCK_RESBLOCK, "a2m_block_create: kind %d", kind);
    A2M_ARG_CHECK(in_channels >= 64 && in_channels % 64 == 0 && in_channels <= 4096, "a2m_block_create: in_channels %d must be a "
                  "multiple of 64 (one 128-byte swizzle row of bf16 per K step)", in_channels);
    A2M_ARG_CHECK(out_channels >= 8 && out_channels % 8 == 0 && out_channels <= 8192, "a2m_block_create: out_channels %d", out_channels);
    const bool same = kind == A2M_BLOCK_SELF_ATTENTION || kind == A2M_BLOCK_CHANNEL_ATTENTION || kind == A2M_BLOCK_RESBLOCK;
    A2M_ARG_CHECK(!same || in_channels == out_channels, "a2m_block_create: this block keeps its channel count");
    A2M_ARG_CHECK(kind != A2M_BLOCK_CHANNEL_ATTENTION || in_channels <= 1024, "a2m_block_create: ChannelAttention up to 1024 channels");
    A2M_CUDA_CHECK(cudaSetDevice(device));
    std::unique_ptr<a2m_model> m(new a2m_model());
    m->device = device;
    for (int i = 0; i < n_tensors; ++i) {
        const a2m_tensor_desc& t = tensors[i];
        A2M_ARG_CHECK(t.name != nullptr && t.ndim >= 0 && t.ndim <= 4 && t.dtype == A2M_DTYPE_F32, "a2m_block_create: bad descriptor %d", i);
        Param p;
        p.ndim = t.ndim;
        for (int k = 0; k < t.ndim; ++k) p.shape[k] = t.shape[k];
        p.f32 = static_cast<const float*>(t.data);
        m->params[t.name] = p;
    }
    m->blk_kind = kind; m->blk_cin = in_channels; m->blk_cout = out_channels;
    Builder b{m.get(), nullptr};
    const std::string p = "blk";
    switch (kind) {
        case A2M_BLOCK_CONV_K3: b.conv1d_k3(m->blk_conv, p, in_channels, 0, out_channels); break;
        case A2M_BLOCK_CONV_K4S2: b.conv1d_k4s2(m->blk_conv, p, in_channels, out_channels); break;
        case A2M_BLOCK_CONV_TRANSPOSE: b.conv_transpose(m->blk_even, m->blk_odd, p, in_channels, out_channels); break;
        case A2M_BLOCK_SELF_ATTENTION: b.attention(m->blk_attn, p, in_channels); break;
        case A2M_BLOCK_CHANNEL_ATTENTION: b.channel(m->blk_chan, p, in_channels); break;
        default: b.resblock(m->blk_res, p, in_channels); break;
    }
    if (kind == A2M_BLOCK_CONV_K3 || kind == A2M_BLOCK_CONV_K4S2) m->blk_conv.act = leaky ? kActLeaky : kActRelu;
    m->err_flag = b.alloc<int>(1);
    if (b.rc == A2M_OK) cudaMemset(m->err_flag, 0, 4);
    cudaError_t e = cudaDeviceSynchronize();
    if (b.rc == A2M_OK && e != cudaSuccess) { a2m_set_error("a2m_block_create: %s", cudaGetErrorString(e)); b.rc = (int)e; }
    if (b.rc != A2M_OK) {
        for (void* q : m->owned) cudaFree(q);
        return b.rc;
    }
    m->params.clear();
    *out = m.release();
    return A2M_OK;
}

extern "C" int a2m_block_forward(a2m_model* m, const float* x_nct, int64_t B, int T, float* out_nct, void* stream) {
    A2M_ARG_CHECK(m != nullptr && x_nct != nullptr && out_nct != nullptr, "a2m_block_forward: NULL argument");
    A2M_ARG_CHECK(m->blk_kind != 0, "a2m_block_forward: the handle is not a building block");
    A2M_ARG_CHECK(B >= 1 && B <= 65535 && T >= 1 && T <= 4096, "a2m_block_forward: B = %lld, T = %d", (long long)B, T);
    A2M_ARG_CHECK(m->blk_kind != A2M_BLOCK_CONV_K4S2 || T % 2 == 0, "a2m_block_forward: the stride-2 block needs an even T (got %d)", T);
    char key[64];
    snprintf(key, sizeof(key), "blk:%lld:%d", (long long)B, T);
    ForwardPlan* P = nullptr;
    auto it = m->plans.find(key);
    if (it != m->plans.end()) {
        P = it->second.get();
    } else {
        std::unique_ptr<ForwardPlan> np(new ForwardPlan());
        np->B = static_cast<int>(B); np->T = T; np->F = 0;
        const int rc = build_block_plan(m, np.get());
        if (rc != A2M_OK) return rc;
        if (m->plans.size() >= 16) m->plans.clear();
        P = np.get();
        m->plans[key] = std::move(np);
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int rc = launch_ncw_to_btc(x_nct, static_cast<int>(B), m->blk_cin, T, P->enc_out, s);
    if (rc != A2M_OK) return rc;
    rc = run_ops(P, 0, static_cast<int>(P->ops.size()), s);
    if (rc != A2M_OK) return rc;
    return launch_btc_to_ncw(P->unet_out, static_cast<int>(B), m->blk_cout, block_out_length(m->blk_kind, T), out_nct, s);
}


// ---------------------------------------------------------------------------------------------------------------
// Discriminator: SelfAttention_D.forward(x) (real_motion_model.py:580-642, eval mode, audio = None, aux_labels = None --
// the two optional arguments cannot work as shipped: the concat with audio has 6144 channels where `logits` takes 4096,
// and the aux classifier is handed a [B] tensor).  pose [B, T, 104] fp32 -> scores [B, T'] fp32.
// Grouped convolutions with groups = 1 (the only value the reference instantiates); every Conv1d + BatchNorm + LeakyReLU
// is one tcgen05 implicit GEMM; the k4 s1 p1 layers shorten the sequence by one step each.
// ---------------------------------------------------------------------------------------------------------------
namespace {

struct DiscLengths { int len[12]; int n; };      // sequence length after the input and after every conv

int disc_lengths(int T, int n_down, DiscLengths* L) {
    int t = T < 4 ? T + (4 - T % 4) : T;         // F.pad(x, (0, 4 - T % 4)) for T < 4 (:583-584)
    int n = 0;
    L->len[n++] = t;
    auto s2 = [&](int v) { return (v + 2 - 4) / 2 + 1; };
    auto s1 = [&](int v) { return v + 2 - 4 + 1; };
    t = s2(t); L->len[n++] = t;
    t = s1(t); L->len[n++] = t;
    for (int i = 0; i < n_down; ++i) { t = s2(t); L->len[n++] = t; t = s1(t); L->len[n++] = t; }
    t = s1(t); L->len[n++] = t;
    t = s1(t); L->len[n++] = t;
    L->len[n++] = t;                              // k3 s1 p1 keeps the length
    L->n = n;
    for (int i = 0; i < n; ++i)
        if (L->len[i] < 1) return -1;
    return t;
}

int build_disc_plan(a2m_model* m, ForwardPlan* P) {
    const int B = P->B, T = P->T, nd = m->disc_down, c = m->disc_c;
    DiscLengths L;
    const int To = disc_lengths(T, nd, &L);
    A2M_ARG_CHECK(To >= 1, "discriminator: %d steps are too few for %d downsampling stages", T, nd);
    Bufs bufs;
    for (int pass = 0; pass < 2; ++pass) {
        Emit E{m, P, pass == 0};
        if (pass == 1) {
            P->ops.clear(); P->op_is_gemm.clear(); P->op_name.clear(); P->op_flops.clear();
            P->gemm_flops = 0;
            cudaError_t e = cudaMalloc(&P->arena, bufs.off + 256);
            if (e != cudaSuccess) { a2m_set_error("discriminator: arena cudaMalloc(%zu) failed: %s", bufs.off, cudaGetErrorString(e)); return (int)e; }
            // rows past a sequence's length (the padding that makes stride-2 views even) are never written: zero once
            e = cudaMemset(P->arena, 0, bufs.off + 256);
            if (e != cudaSuccess) { a2m_set_error("discriminator: arena memset: %s", cudaGetErrorString(e)); return (int)e; }
            bufs.base = static_cast<unsigned char*>(P->arena);
        }
        bufs.off = 0;
        auto even = [](int v) { return v + (v & 1); };
        const int n_conv = static_cast<int>(m->disc_conv.size());
        // buffer i holds the input of conv i (i = 0: the padded pose); channels per buffer
        std::vector<int> chan(n_conv + 1);
        chan[0] = 128;
        for (int i = 0; i < n_conv; ++i) chan[i + 1] = m->disc_conv[i].N;
        std::vector<__nv_bfloat16*> buf(n_conv + 1);
        for (int i = 0; i <= n_conv; ++i) buf[i] = bufs.get<__nv_bfloat16>(static_cast<size_t>(B) * even(L.len[i]) * chan[i]);
        const int C4 = 4 * c, Tl = To;                 // 2048 channels, final length
        auto* attn_out = bufs.get<__nv_bfloat16>(static_cast<size_t>(B) * Tl * C4);
        auto* qkv = bufs.get<__nv_bfloat16>(static_cast<size_t>(B) * Tl * (C4 + 2 * (C4 / 8)));
        auto* pooled = bufs.get<__nv_bfloat16>(static_cast<size_t>(B) * C4);
        auto* nodes_in = bufs.get<__nv_bfloat16>(static_cast<size_t>(B) * (kBodyJoints + kHandJoints) * kJointFeat);
        auto* nodes_out = bufs.get<__nv_bfloat16>(static_cast<size_t>(B) * (kBodyJoints + kHandJoints) * kJointFeat);
        auto* xg = bufs.get<__nv_bfloat16>(static_cast<size_t>(B) * C4);
        auto* scores = bufs.get<float>(static_cast<size_t>(B) * Tl);
        P->enc_out = buf[0]; P->pose_stage = scores; P->T_out = Tl;

        // conv stack.  Layer kinds in order: conv1 (s2, s1), conv2 stages (s2, s1), conv3 (s1, s1, [attention], k3)
        int li = 0;
        auto s2 = [&](const char* tag) {
            E.tag = tag;
            E.conv_k4s2_io(m->disc_conv[li], buf[li], chan[li], even(L.len[li]), L.len[li + 1], even(L.len[li + 1]), B, buf[li + 1]);
            ++li;
        };
        auto s1 = [&](const char* tag, __nv_bfloat16* in = nullptr, int in_alloc = 0) {
            E.tag = tag;
            E.conv_rows_io(m->disc_conv[li], in ? in : buf[li], chan[li], L.len[li], in ? in_alloc : even(L.len[li]), L.len[li + 1],
                           even(L.len[li + 1]), B, buf[li + 1]);
            ++li;
        };
        s2("disc.conv1.0"); s1("disc.conv1.4");
        for (int i = 0; i < nd; ++i) { s2("disc.conv2.a"); s1("disc.conv2.b"); }
        s1("disc.conv3.0");
        // conv3.4 writes a dense [B, Tl, 4c] tensor (the attention's layout): out_alloc = Tl
        E.tag = "disc.conv3.4";
        E.conv_rows_io(m->disc_conv[li], buf[li], chan[li], L.len[li], even(L.len[li]), L.len[li + 1], L.len[li + 1], B, buf[li + 1]);
        ++li;
        E.tag = "disc.conv3.8";
        E.attention(m->disc_attn, buf[li], nullptr, Tl, B, qkv, attn_out);
        E.tag = "disc.conv3.9";
        E.conv_rows_io(m->disc_conv[li], attn_out, C4, Tl, Tl, Tl, Tl, B, buf[li + 1]);
        __nv_bfloat16* feat = buf[li + 1];             // [B, Tl, 4c] dense (k3 keeps the length; allocated even(Tl) rows, used Tl)
        ++li;
        E.tag = "disc.pool";
        E.op([=](cudaStream_t st) { return launch_mean_time(feat, B, Tl, C4, pooled, st); });
        // graph branches: body = channels [0, 2c), hand = [2c, 4c)  (:595-616)
        for (int part = 0; part < 2; ++part) {
            const int J = part == 0 ? kBodyJoints : kHandJoints;
            __nv_bfloat16* nin = nodes_in + (part == 0 ? 0 : static_cast<size_t>(B) * kBodyJoints * kJointFeat);
            __nv_bfloat16* nout = nodes_out + (part == 0 ? 0 : static_cast<size_t>(B) * kBodyJoints * kJointFeat);
            {
                E.tag = part == 0 ? "disc.body_proj" : "disc.hand_proj";
                AView v;
                v.ptr = pooled + part * 2 * c; v.rank = 2; v.dims[0] = 2 * c; v.dims[1] = B; v.strides[0] = 1; v.strides[1] = C4;
                const int box[4] = {128, 1, 1, 1}, ext[4] = {B, 1, 1, 1};
                const long long os[4] = {static_cast<long long>(J) * kJointFeat, 0, 0, 0};
                E.gemm(m->disc_proj[part], m->disc_proj[part].taps, &v, 1, box, ext, nin, os, 0, kOutBf16);
            }
            E.tag = part == 0 ? "disc.body_gat" : "disc.hand_gat";
            const float *wt = m->disc_gat_wt[part], *as = m->disc_gat_src[part], *ad = m->disc_gat_dst[part], *gb = m->disc_gat_bias[part];
            const int *nbr = m->disc_nbr[part], *deg = m->disc_deg[part];
            E.op([=](cudaStream_t st) { return launch_gat_single(nin, B, J, wt, as, ad, gb, nbr, deg, nout, st); });
            E.tag = part == 0 ? "disc.body_graph_out" : "disc.hand_graph_out";
            E.linear_rows(m->disc_out[part], nout, nullptr, J * kJointFeat, B, xg, C4, part * 2 * c, kOutBf16);
        }
        E.tag = "disc.logits";
        const float *lw = m->disc_logit_w, *lb = m->disc_logit_b;
        E.op([=](cudaStream_t st) { return launch_disc_logits(feat, xg, lw, lb, B, Tl, C4, scores, st); });
        if (E.rc != A2M_OK) return E.rc;
    }
    return A2M_OK;
}

}  // namespace

extern "C" int a2m_disc_create(const a2m_tensor_desc* tensors, int n_tensors, int n_downsampling, int device, a2m_model** out) {
    A2M_ARG_CHECK(out != nullptr && tensors != nullptr && n_tensors > 0, "a2m_disc_create: NULL argument");
    *out = nullptr;
    A2M_ARG_CHECK(n_downsampling >= 0 && n_downsampling <= 2, "a2m_disc_create: n_downsampling %d (0..2: the channel count "
                  "grows by 2^n per stage)", n_downsampling);
    A2M_CUDA_CHECK(cudaSetDevice(device));
    std::unique_ptr<a2m_model> m(new a2m_model());
    m->device = device;
    for (int i = 0; i < n_tensors; ++i) {
        const a2m_tensor_desc& t = tensors[i];
        A2M_ARG_CHECK(t.name != nullptr && t.ndim >= 0 && t.ndim <= 4, "a2m_disc_create: bad descriptor %d", i);
        Param p;
        p.ndim = t.ndim;
        for (int k = 0; k < t.ndim; ++k) p.shape[k] = t.shape[k];
        if (t.dtype == A2M_DTYPE_F32) p.f32 = static_cast<const float*>(t.data);
        else if (t.dtype == A2M_DTYPE_I64) p.i64 = static_cast<const long long*>(t.data);
        else { a2m_set_error("a2m_disc_create: tensor '%s' has unsupported dtype %d", t.name, t.dtype); return A2M_ERR_ARGUMENT; }
        m->params[t.name] = p;
    }
    Builder b{m.get(), nullptr};
    const Param* w0 = b.find("conv1.0.weight");
    A2M_ARG_CHECK(w0 && w0->ndim == 3 && w0->shape[1] == kPoseFeats && w0->shape[2] == 4 && w0->shape[0] == 64,
                  "a2m_disc_create: conv1.0.weight must be [64, 104, 4] (in_channels 104, out_channels 64, groups 1)");
    m->has_disc = true; m->disc_down = n_downsampling;
    int c = 64;
    m->disc_conv.resize(2 + 2 * n_downsampling + 3);
    int li = 0;
    b.conv_bn(m->disc_conv[li++], "conv1.0", "conv1.1", 0, kPoseFeats, 128, c);
    b.conv_bn(m->disc_conv[li++], "conv1.4", "conv1.5", 1, c, c, c);
    for (int n = 1; n <= n_downsampling; ++n) {
        const int mul = 1 << n;
        const std::string p = "conv2." + std::to_string(n - 1);
        b.conv_bn(m->disc_conv[li++], p + ".0", p + ".1", 0, c, c, c * mul);
        b.conv_bn(m->disc_conv[li++], p + ".4", p + ".5", 1, c * mul, c * mul, c * mul);
        c *= mul;
    }
    m->disc_c = c;
    b.conv_bn(m->disc_conv[li++], "conv3.0", "conv3.1", 1, c, c, 2 * c);
    b.conv_bn(m->disc_conv[li++], "conv3.4", "conv3.5", 1, 2 * c, 2 * c, 4 * c);
    b.attention(m->disc_attn, "conv3.8", 4 * c);
    b.conv_bn(m->disc_conv[li++], "conv3.9", "conv3.10", 2, 4 * c, 4 * c, 4 * c);
    const char* parts[2] = {"body", "hand"};
    for (int part = 0; part < 2 && b.rc == A2M_OK; ++part) {
        const int J = part == 0 ? kBodyJoints : kHandJoints;
        const std::string p = parts[part];
        b.linear(m->disc_proj[part], p + "_proj.weight", p + "_proj.bias", 2 * c, J * kJointFeat, kActNone);
        b.linear(m->disc_out[part], p + "_graph_out.weight", p + "_graph_out.bias", J * kJointFeat, 2 * c, kActNone);
        m->disc_gat_src[part] = b.keep(p + "_gat.att_src", kGatHeads * kJointFeat);
        m->disc_gat_dst[part] = b.keep(p + "_gat.att_dst", kGatHeads * kJointFeat);
        m->disc_gat_bias[part] = b.keep(p + "_gat.bias", kJointFeat);
        std::string wname = p + "_gat.lin.weight";
        if (!b.find(wname, false))
            for (const char* alt : {"_gat.lin_src.weight", "_gat.lin_l.weight"})
                if (b.find(p + alt, false)) { wname = p + alt; break; }
        const float* w = b.f32(wname, static_cast<long long>(kGatHeads) * kJointFeat * kJointFeat);
        float* wt = b.alloc<float>(static_cast<size_t>(kGatHeads) * kJointFeat * kJointFeat);
        if (b.rc == A2M_OK) {                          // transpose [256][64] -> [64][256] on the host (64 KB, once)
            std::vector<float> h(256 * 64), ht(256 * 64);
            cudaError_t e = cudaMemcpy(h.data(), w, h.size() * 4, cudaMemcpyDeviceToHost);
            for (int o = 0; o < 256; ++o)
                for (int f = 0; f < 64; ++f) ht[f * 256 + o] = h[o * 64 + f];
            if (e == cudaSuccess) e = cudaMemcpy(wt, ht.data(), ht.size() * 4, cudaMemcpyHostToDevice);
            if (e != cudaSuccess) { a2m_set_error("a2m_disc_create: %s", cudaGetErrorString(e)); b.rc = (int)e; }
        }
        m->disc_gat_wt[part] = wt;
        DecoderW topo;
        b.topology(topo, p + "_edge_index_template", J);
        m->disc_nbr[part] = topo.nbr; m->disc_deg[part] = topo.deg;
    }
    if (b.rc == A2M_OK) {                              // logits Conv1d(8c -> 1, k3): [1][8c][3] -> tap-major [3][8c]
        const int C8 = 8 * c;
        const Param* lp = b.find("logits.weight");
        if (lp && (lp->ndim != 3 || lp->shape[0] != 1 || lp->shape[1] != C8 || lp->shape[2] != 3)) {
            a2m_set_error("a2m_disc_create: logits.weight must be [1, %d, 3] (out_shape 1, groups 1)", C8);
            b.rc = A2M_ERR_UNSUPPORTED;
        }
        const float* w = b.f32("logits.weight", 3LL * C8);
        float* wt = b.alloc<float>(3 * static_cast<size_t>(C8));
        if (b.rc == A2M_OK) {
            std::vector<float> h(3 * static_cast<size_t>(C8)), ht(h.size());
            cudaError_t e = cudaMemcpy(h.data(), w, h.size() * 4, cudaMemcpyDeviceToHost);
            for (int ch = 0; ch < C8; ++ch)
                for (int k = 0; k < 3; ++k) ht[static_cast<size_t>(k) * C8 + ch] = h[static_cast<size_t>(ch) * 3 + k];
            if (e == cudaSuccess) e = cudaMemcpy(wt, ht.data(), ht.size() * 4, cudaMemcpyHostToDevice);
            if (e != cudaSuccess) { a2m_set_error("a2m_disc_create: %s", cudaGetErrorString(e)); b.rc = (int)e; }
        }
        m->disc_logit_w = wt;
        m->disc_logit_b = b.keep("logits.bias", 1);
    }
    m->err_flag = b.alloc<int>(1);
    if (b.rc == A2M_OK) cudaMemset(m->err_flag, 0, 4);
    cudaError_t e = cudaDeviceSynchronize();
    if (b.rc == A2M_OK && e != cudaSuccess) { a2m_set_error("a2m_disc_create: %s", cudaGetErrorString(e)); b.rc = (int)e; }
    if (b.rc != A2M_OK) {
        for (void* q : m->owned) cudaFree(q);
        return b.rc;
    }
    m->params.clear();
    *out = m.release();
    return A2M_OK;
}

extern "C" int a2m_disc_out_length(int T, int n_downsampling) {
    DiscLengths L;
    return T >= 1 && n_downsampling >= 0 && n_downsampling <= 2 ? disc_lengths(T, n_downsampling, &L) : -1;
}

extern "C" int a2m_disc_forward(a2m_model* m, const float* pose, int64_t B, int T, float* scores, void* stream) {
    A2M_ARG_CHECK(m != nullptr && pose != nullptr && scores != nullptr, "a2m_disc_forward: NULL argument");
    A2M_ARG_CHECK(m->has_disc, "a2m_disc_forward: the handle is not a discriminator");
    A2M_ARG_CHECK(B >= 1 && B <= 65535 && T >= 4 && T <= 4096, "a2m_disc_forward: B = %lld, T = %d (4 <= T <= 4096)", (long long)B, T);
    char key[64];
    snprintf(key, sizeof(key), "disc:%lld:%d", (long long)B, T);
    ForwardPlan* P = nullptr;
    auto it = m->plans.find(key);
    if (it != m->plans.end()) {
        P = it->second.get();
    } else {
        std::unique_ptr<ForwardPlan> np(new ForwardPlan());
        np->B = static_cast<int>(B); np->T = T; np->F = 0;
        const int rc = build_disc_plan(m, np.get());
        if (rc != A2M_OK) return rc;
        if (m->plans.size() >= 16) m->plans.clear();
        P = np.get();
        m->plans[key] = std::move(np);
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    DiscLengths L;
    disc_lengths(T, m->disc_down, &L);
    int rc = launch_pose_pad(pose, static_cast<int>(B), T, L.len[0] + (L.len[0] & 1), kPoseFeats, 128, P->enc_out, s);
    if (rc != A2M_OK) return rc;
    rc = run_ops(P, 0, static_cast<int>(P->ops.size()), s);
    if (rc != A2M_OK) return rc;
    A2M_CUDA_CHECK(cudaMemcpyAsync(scores, P->pose_stage, static_cast<size_t>(B) * P->T_out * sizeof(float), cudaMemcpyDeviceToDevice, s));
    return A2M_OK;
}
