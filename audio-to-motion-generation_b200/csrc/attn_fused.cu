// Fused SelfAttention block for the 256-channel decoder layers (model_layers.py:121-146):
//     q | k | v = 1x1 convs of x (one GEMM, N = 32 + 32 + 256)      -> tcgen05, TMA-fed K pipeline
//     S = q k^T  (no 1/sqrt(d): model_layers.py:140)                 -> tcgen05, all clips of the tile at once
//     P = exp(S - rowmax) restricted to the row's own clip           -> CUDA cores, one thread per row
//     O = P v                                                        -> tcgen05 (B = v read MN-major)
//     out = gamma * O / rowsum + x (+ res2: the ResBlock skip, :190) -> bf16
// Optionally the ChannelAttention that follows the block in the hand decoder (model_layers.py:167-174) is applied in
// the epilogue: the CTA holds whole clips, so the per-clip average / maximum over time, the 256 -> 32 -> 256 MLP and the
// per-channel rescale need nothing from outside the tile (one launch and one [B, T, 256] round trip less).
// One CTA owns 128 consecutive rows = 128 / T whole clips (T in {8, 16, 32, 64}); q, k, v, S and O never leave
// the SM (TMEM accumulators, bf16 operand tiles in shared memory).  Replaces a GEMM launch + an attention
// launch and the [B, T, 320] bf16 round trip through L2 between them.
#include <cuda.h>
#include <cstring>
#include "conv_gemm.cuh"
#include "layers.cuh"

void a2m_count_launch();

namespace a2m {

int make_weight_map(CUtensorMap* map, const void* w, long long n_rows, long long k, int box_rows);   // conv_gemm.cu

namespace {

constexpr int kThreads = 512;
constexpr int kC = 256, kD = 32, kNqkv = 2 * kD + kC;      // 320
constexpr int kStages = 3;
constexpr int kABytes = 128 * 64 * 2;                      // 16 KB
constexpr int kBBytes = kNqkv * 64 * 2;                    // 40 KB
constexpr int kStageBytes = kABytes + kBBytes;             // 56 KB
// operand tiles that alias the (drained) pipeline ring
constexpr int kOffQ = 0, kOffK = 16384, kOffP = 32768, kOffV = 65536;       // Q, K [128][64]; P [128][128]; V 4 x [128][64]
constexpr int kOffSum = kStages * kStageBytes;             // row sums [128] fp32
constexpr int kOffBias = kOffSum + 512;                    // q | k | v bias [320] fp32
constexpr int kOffBar = kOffBias + kNqkv * 4;
constexpr int kSmemBytes = kOffBar + 128 + 1024;
constexpr uint32_t kColS = 320;                            // TMEM: QKV [0,320) (later O [0,256)), S [320,448)
static_assert(kOffV + 65536 <= kStages * kStageBytes, "operand tiles must fit in the ring");

struct AttnParams {
    CUtensorMap x_map;        // [rows][256] bf16, box 64 x 128
    CUtensorMap w_map;        // [320][256] bf16, box 64 x 160
    const float* bias;        // [320]
    const float* gamma;       // device scalar
    const __nv_bfloat16* x;   // residual
    const __nv_bfloat16* res2;
    __nv_bfloat16* out;
    long long n_rows;
    int T;
    // optional ChannelAttention applied to the block's output (model_layers.py:167-174), hand decoder order
    // SelfAttention -> ChannelAttention: out * (sigmoid(mlp(avg_T out)) + sigmoid(mlp(max_T out))); hidden = 32
    const float* ca_w0;       // fc.0.weight [32][256]; NULL = no channel attention
    const float* ca_b0;       // [32]
    const float* ca_w2;       // fc.2.weight transposed [32][256]
    const float* ca_b2;       // [256]
};

__device__ __forceinline__ uint32_t pack2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ int sw128_off(int r, int c) { return r * 128 + ((c ^ (r & 7)) << 4); }
// MN-major SW128 descriptor with an explicit leading-dimension byte offset (stride between 64-element atoms along N)
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}

__global__ void __launch_bounds__(kThreads, 1)
attn_fused_kernel(const __grid_constant__ AttnParams p, int* __restrict__ err_flag) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    unsigned char* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    float* s_sum = reinterpret_cast<float*>(smem + kOffSum);
    float* s_bias = reinterpret_cast<float*>(smem + kOffBias);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kOffBar);      // [kStages]
    uint64_t* free0_bar = full_bar + kStages;                              // stage 0 drained by the first K block
    uint64_t* gemm_bar = free0_bar + 1;                                    // q | k | v accumulators complete
    uint64_t* s_bar = gemm_bar + 1;                                        // S = q k^T complete
    uint64_t* o_bar = s_bar + 1;                                           // O = P v complete
    uint64_t* qk_bar = o_bar + 1;                                          // 128 arrivals: q | k staged
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(qk_bar + 1);

    const int tid = threadIdx.x, warp = tid >> 5;
    const int r = tid & 127, q = tid >> 7, quad = warp & 3;
    const long long row0 = static_cast<long long>(blockIdx.x) * 128;
    const int T = p.T;

    pdl_launch_dependents();
    if (tid == 0) {
        tma_prefetch_desc(&p.x_map);
        tma_prefetch_desc(&p.w_map);
        for (int s = 0; s < kStages; ++s) mbar_init(&full_bar[s], 1);
        mbar_init(free0_bar, 1);
        mbar_init(gemm_bar, 1);
        mbar_init(s_bar, 1);
        mbar_init(o_bar, 1);
        mbar_init(qk_bar, 128);
        mbar_fence_init();
    }
    if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
    if (tid < kNqkv) s_bias[tid] = __ldg(p.bias + tid);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const uint32_t ring = smem_u32(smem);
    pdl_wait();                                           // x (and res2) come from the previous kernels
    if (tid == 0) {
        for (int kb = 0; kb < kStages; ++kb) {           // the first kStages K blocks; the 4th reuses stage 0 below
            unsigned char* st = smem + kb * kStageBytes;
            mbar_expect_tx(&full_bar[kb], kStageBytes);
            tma_load_5d(st, &p.x_map, &full_bar[kb], kb * 64, static_cast<int>(row0), 0, 0, 0);
            tma_load_5d(st + kABytes, &p.w_map, &full_bar[kb], kb * 64, 0, 0, 0, 0);
            tma_load_5d(st + kABytes + 160 * 128, &p.w_map, &full_bar[kb], kb * 64, 160, 0, 0, 0);
        }
    }

    // ---------------- q | k | v = x Wqkv^T : 4 K blocks through a 3-stage ring ----------------
    if (tid == 0) {
        const uint32_t id256 = umma_idesc_bf16(128, 256), id64 = umma_idesc_bf16(128, 64);
        for (int kb = 0; kb < 4; ++kb) {
            const int s = kb % kStages;
            mbar_wait(&full_bar[s], (kb / kStages) & 1, err_flag, 21);
            tc_fence_after();
            const uint32_t a = ring + s * kStageBytes, b = a + kABytes;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                umma_bf16(tmem_base, umma_desc_sw128(a + k * 32), umma_desc_sw128(b + k * 32), id256, (kb | k) != 0);
                umma_bf16(tmem_base + 256, umma_desc_sw128(a + k * 32), umma_desc_sw128(b + 32768 + k * 32), id64, (kb | k) != 0);
            }
            if (kb == 0) {                                // stage 0 is free once these MMAs have read it: 4th K block
                umma_commit(free0_bar);
                mbar_wait(free0_bar, 0, err_flag, 22);
                unsigned char* st = smem;
                mbar_expect_tx(&full_bar[0], kStageBytes);
                tma_load_5d(st, &p.x_map, &full_bar[0], 3 * 64, static_cast<int>(row0), 0, 0, 0);
                tma_load_5d(st + kABytes, &p.w_map, &full_bar[0], 3 * 64, 0, 0, 0, 0);
                tma_load_5d(st + kABytes + 160 * 128, &p.w_map, &full_bar[0], 3 * 64, 160, 0, 0, 0);
            }
        }
        umma_commit(gemm_bar);
    }
    mbar_wait(gemm_bar, 0, err_flag, 23);
    tc_fence_after();

    // ---------------- stage q, k (quarter 0) and v (all quarters) as bf16 MMA operands ----------------
    if (q == 0) {
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {                  // hf 0: q -> Qs, hf 1: k -> Ks; columns 32..63 of both are zero
            uint32_t t[32];
            tmem_ld_32x32(tmem_lane + hf * 32, t);
            tmem_ld_wait();
            unsigned char* dst = smem + (hf == 0 ? kOffQ : kOffK);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint4 o;
                const float4 b0 = *reinterpret_cast<const float4*>(s_bias + hf * 32 + c * 8);
                const float4 b1 = *reinterpret_cast<const float4*>(s_bias + hf * 32 + c * 8 + 4);
                o.x = pack2(__uint_as_float(t[c * 8]) + b0.x, __uint_as_float(t[c * 8 + 1]) + b0.y);
                o.y = pack2(__uint_as_float(t[c * 8 + 2]) + b0.z, __uint_as_float(t[c * 8 + 3]) + b0.w);
                o.z = pack2(__uint_as_float(t[c * 8 + 4]) + b1.x, __uint_as_float(t[c * 8 + 5]) + b1.y);
                o.w = pack2(__uint_as_float(t[c * 8 + 6]) + b1.z, __uint_as_float(t[c * 8 + 7]) + b1.w);
                *reinterpret_cast<uint4*>(dst + sw128_off(r, c)) = o;
                *reinterpret_cast<uint4*>(dst + sw128_off(r, c + 4)) = make_uint4(0, 0, 0, 0);
            }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(qk_bar);
    }
    if (tid == 128) {                                     // S = q k^T as soon as q | k are staged (v staging overlaps)
        mbar_wait(qk_bar, 0, err_flag, 24);
        tc_fence_after();
        const uint32_t id128 = umma_idesc_bf16(128, 128);
#pragma unroll
        for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base + kColS, umma_desc_sw128(ring + kOffQ + k * 32), umma_desc_sw128(ring + kOffK + k * 32), id128, k != 0);
        umma_commit(s_bar);
    }
    {   // v: this quarter's 64 channels -> Vs chunk q ([128 rows (time)][64 channels], read MN-major by the P.v MMA)
        unsigned char* dst = smem + kOffV + q * 16384;
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
            uint32_t t[32];
            tmem_ld_32x32(tmem_lane + 64 + q * 64 + hf * 32, t);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint4 o;
                const float4 b0 = *reinterpret_cast<const float4*>(s_bias + 64 + q * 64 + hf * 32 + c * 8);
                const float4 b1 = *reinterpret_cast<const float4*>(s_bias + 64 + q * 64 + hf * 32 + c * 8 + 4);
                o.x = pack2(__uint_as_float(t[c * 8]) + b0.x, __uint_as_float(t[c * 8 + 1]) + b0.y);
                o.y = pack2(__uint_as_float(t[c * 8 + 2]) + b0.z, __uint_as_float(t[c * 8 + 3]) + b0.w);
                o.z = pack2(__uint_as_float(t[c * 8 + 4]) + b1.x, __uint_as_float(t[c * 8 + 5]) + b1.y);
                o.w = pack2(__uint_as_float(t[c * 8 + 6]) + b1.z, __uint_as_float(t[c * 8 + 7]) + b1.w);
                *reinterpret_cast<uint4*>(dst + sw128_off(r, hf * 4 + c)) = o;
            }
        }
    }
    // zero P (quarters 1..3; the row threads then write their own clip's block)
    if (q != 0) {
        uint4* pz = reinterpret_cast<uint4*>(smem + kOffP);
        for (int i = tid - 128; i < 32768 / 16; i += kThreads - 128) pz[i] = make_uint4(0, 0, 0, 0);
    }
    __syncthreads();                                      // P zeroed before the row threads write into it

    // ---------------- softmax of each row over its own clip (quarter 0: one thread per row) ----------------
    if (q == 0) {
        mbar_wait(s_bar, 0, err_flag, 25);
        tc_fence_after();
        const int win = T < 32 ? 32 : T;                  // warp-uniform column window that covers the warp's clips
        const int win0 = ((quad * 32) / win) * win;
        const int c_lo = (r / T) * T - win0, c_hi = c_lo + T;    // my clip's columns inside the window
        float e[64];
        float m = -INFINITY;
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
            if (hf * 32 < win) {
                uint32_t t[32];
                tmem_ld_32x32(tmem_lane + kColS + win0 + hf * 32, t);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int c = hf * 32 + j;
                    e[c] = (c >= c_lo && c < c_hi) ? __uint_as_float(t[j]) : -INFINITY;
                    m = fmaxf(m, e[c]);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) e[hf * 32 + j] = -INFINITY;
            }
        }
        float sum = 0.f;
#pragma unroll
        for (int c = 0; c < 64; ++c) { e[c] = __expf(e[c] - m); sum += e[c]; }       // exp(-inf) = 0 outside the clip
        s_sum[r] = sum;
        // P row: 8-column chunks of the window that intersect my clip (T is a multiple of 8, so whole chunks)
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
            const int c = ch * 8;
            if (c >= c_lo && c < c_hi) {
                uint4 o;
                o.x = pack2(e[c], e[c + 1]); o.y = pack2(e[c + 2], e[c + 3]);
                o.z = pack2(e[c + 4], e[c + 5]); o.w = pack2(e[c + 6], e[c + 7]);
                const int col = win0 + c;                 // absolute column (time row of the tile)
                *reinterpret_cast<uint4*>(smem + kOffP + (col >> 6) * 16384 + sw128_off(r, (col & 63) >> 3)) = o;
            }
        }
    }
    tc_fence_before();
    fence_proxy_async_smem();
    __syncthreads();                                      // v and P staged, all TMEM reads of q | k | v done

    // ---------------- O = P v ----------------
    if (tid == 0) {
        tc_fence_after();
        const uint32_t id_o = umma_idesc_bf16(128, 256) | (1u << 16);         // B (v) is MN-major
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
            umma_bf16(tmem_base, umma_desc_sw128(ring + kOffP + (kk >> 2) * 16384 + (kk & 3) * 32),
                      umma_desc_mn(ring + kOffV + kk * 2048, 16384), id_o, kk != 0);
        umma_commit(o_bar);
    }
    // the residual rows are fetched while the P.v MMA runs (16 independent 16-byte loads per thread in flight).
    // x / res2 were written by earlier kernels of the stream that this (programmatically launched) kernel
    // overlapped with: read them through L2 (ld.global.cg), never through the non-coherent path, whose L1
    // lines may predate those writes
    const long long row = row0 + r;
    const bool live = row < p.n_rows;
    const long long base = row * kC + q * 64;
    uint4 xv[8], rv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        xv[i] = make_uint4(0, 0, 0, 0);
        rv[i] = make_uint4(0, 0, 0, 0);
        if (live) {
            xv[i] = __ldcg(reinterpret_cast<const uint4*>(p.x + base + i * 8));
            if (p.res2) rv[i] = __ldcg(reinterpret_cast<const uint4*>(p.res2 + base + i * 8));
        }
    }
    mbar_wait(o_bar, 0, err_flag, 26);
    tc_fence_after();

    // ---------------- out = gamma * O / rowsum + x (+ res2) ----------------
    const bool chan = p.ca_w0 != nullptr;                 // kernel-uniform
    // with channel attention the bf16 output tile goes to the (drained) operand ring first, rows padded to 528 B so that
    // both the row-wise 16-byte stores and the column walks of the pooling are conflict-free
    constexpr int kRowB = 2 * kC + 16;
    unsigned char* s_o = smem;
    {
        const float scale = __ldg(p.gamma) / s_sum[r];
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
            uint32_t t[32];
            tmem_ld_32x32(tmem_lane + q * 64 + hf * 32, t);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const uint4 xq = xv[hf * 4 + c], rq = rv[hf * 4 + c];
                const uint32_t xs[4] = {xq.x, xq.y, xq.z, xq.w}, rs[4] = {rq.x, rq.y, rq.z, rq.w};
                uint32_t os[4];
#pragma unroll
                for (int e2 = 0; e2 < 4; ++e2) {
                    const float a = scale * __uint_as_float(t[c * 8 + 2 * e2]) + __uint_as_float(xs[e2] << 16) +
                                    __uint_as_float(rs[e2] << 16);
                    const float b = scale * __uint_as_float(t[c * 8 + 2 * e2 + 1]) + __uint_as_float(xs[e2] & 0xffff0000u) +
                                    __uint_as_float(rs[e2] & 0xffff0000u);
                    os[e2] = pack2(a, b);
                }
                const uint4 o4 = make_uint4(os[0], os[1], os[2], os[3]);
                if (chan) *reinterpret_cast<uint4*>(s_o + r * kRowB + (q * 64 + hf * 32 + c * 8) * 2) = o4;
                else if (live) *reinterpret_cast<uint4*>(p.out + base + hf * 32 + c * 8) = o4;
            }
        }
    }
    if (chan) {
        float* s_avg = reinterpret_cast<float*>(smem + 128 * kRowB);       // [clips][256]
        const int clips = 128 / T;
        float* s_max = s_avg + clips * kC;                                 // [clips][256]
        float* s_scale = s_max + clips * kC;                               // [clips][256]
        float* s_hid = s_scale + clips * kC;                               // [clips][2][32]
        __syncthreads();
        for (int item = tid; item < clips * kC; item += kThreads) {         // per clip and channel: mean and max over time
            const int clip = item >> 8, c = item & (kC - 1);
            const unsigned char* col = s_o + (clip * T) * kRowB + c * 2;
            float sum = 0.f, mx = -INFINITY;
            for (int t = 0; t < T; ++t) {
                const float v = __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(col + t * kRowB));
                sum += v;
                mx = fmaxf(mx, v);
            }
            s_avg[item] = sum / static_cast<float>(T);
            s_max[item] = mx;
        }
        __syncthreads();
        for (int u = warp; u < clips * 64; u += kThreads / 32) {            // hidden units: relu(w0 . pooled + b0), 32 per branch
            const int clip = u >> 6, which = (u >> 5) & 1, unit = u & 31;
            const float* src = (which ? s_max : s_avg) + clip * kC;
            const int lane = tid & 31;
            float acc = 0.f;
            for (int k = lane; k < kC; k += 32) acc = fmaf(__ldg(p.ca_w0 + unit * kC + k), src[k], acc);
            acc = warp_sum(acc);
            if (lane == 0) s_hid[u] = fmaxf(acc + __ldg(p.ca_b0 + unit), 0.f);
        }
        __syncthreads();
        for (int item = tid; item < clips * kC; item += kThreads) {
            const int clip = item >> 8, c = item & (kC - 1);
            const float* h = s_hid + clip * 64;
            float za = __ldg(p.ca_b2 + c), zm = za;
#pragma unroll 8
            for (int u = 0; u < 32; ++u) {
                const float w = __ldg(p.ca_w2 + u * kC + c);
                za = fmaf(w, h[u], za);
                zm = fmaf(w, h[32 + u], zm);
            }
            s_scale[item] = 1.f / (1.f + __expf(-za)) + 1.f / (1.f + __expf(-zm));       // sigmoid each, then add
        }
        __syncthreads();
        if (live) {
            const float* sc = s_scale + (r / T) * kC + q * 64;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint4 mine = *reinterpret_cast<const uint4*>(s_o + r * kRowB + (q * 64 + i * 8) * 2);
                const uint32_t ov[4] = {mine.x, mine.y, mine.z, mine.w};
                uint32_t o4[4];
#pragma unroll
                for (int e2 = 0; e2 < 4; ++e2) {
                    const uint32_t v = ov[e2];
                    const int c = i * 8 + e2 * 2;
                    o4[e2] = pack2(__uint_as_float(v << 16) * sc[c], __uint_as_float(v & 0xffff0000u) * sc[c + 1]);
                }
                *reinterpret_cast<uint4*>(p.out + base + i * 8) = make_uint4(o4[0], o4[1], o4[2], o4[3]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

}  // namespace

struct AttnFusedPlan {
    AttnParams p;
    int grid;
};

bool attn_fused_supported(int T, int C) { return C == kC && T >= 8 && T <= 64 && (128 % T) == 0; }

int attn_fused_plan(const __nv_bfloat16* w_qkv, const float* bias_qkv, const float* gamma, const __nv_bfloat16* x,
                    const __nv_bfloat16* res2, int B, int T, int C, __nv_bfloat16* out, std::shared_ptr<AttnFusedPlan>* plan_out,
                    const float* ca_w0, const float* ca_b0, const float* ca_w2, const float* ca_b2) {
    A2M_ARG_CHECK(attn_fused_supported(T, C), "attn_fused: T = %d, C = %d not supported", T, C);
    auto plan = std::make_shared<AttnFusedPlan>();
    AttnParams& p = plan->p;
    memset(&p, 0, sizeof(p));
    const long long rows = static_cast<long long>(B) * T;
    int rc = make_weight_map(&p.x_map, x, rows, kC, 128);
    if (rc != A2M_OK) return rc;
    rc = make_weight_map(&p.w_map, w_qkv, kNqkv, kC, 160);
    if (rc != A2M_OK) return rc;
    p.bias = bias_qkv; p.gamma = gamma; p.x = x; p.res2 = res2; p.out = out; p.n_rows = rows; p.T = T;
    p.ca_w0 = ca_w0; p.ca_b0 = ca_b0; p.ca_w2 = ca_w2; p.ca_b2 = ca_b2;
    plan->grid = static_cast<int>((rows + 127) / 128);
    *plan_out = plan;
    return A2M_OK;
}

int attn_fused_launch(const AttnFusedPlan& plan, int* err_flag, cudaStream_t stream) {
    static A2mPerDeviceOnce configured;
    if (configured.first()) {
        A2M_CUDA_CHECK(cudaFuncSetAttribute(attn_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    }
    A2M_CUDA_CHECK(a2m_launch_pdl(attn_fused_kernel, dim3(plan.grid), dim3(kThreads), kSmemBytes, stream, plan.p, err_flag));
    a2m_count_launch();
    return A2M_OK;
}

}  // namespace a2m
