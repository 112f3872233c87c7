// EXPERIMENTAL (opt-in with A2M_RESBLOCK_FUSION=1; see the note at its call site in model.cu and DESIGN.md section 9).
// Fused ResBlock of the decoder stacks (model_layers.py:177-190, with the ConvNormRelu / SelfAttention it is made of:
// :51-118, :121-146), T = 64, 256 channels:
//     t1 = LeakyReLU(BN(conv3(x)))   t2 = LeakyReLU(BN(conv3(t1)))   out = gamma * softmax(q k^T) v + t2 + x
// in ONE kernel per 128 rows (two whole clips).  Replaces two implicit-GEMM launches and the fused attention launch: the
// three are each latency-bound (15-45 us in flight for ~5 us of tensor work, profiles/r1_timeline_2lanes.txt), and
// everything a clip needs is local to it -- k = 3 convolutions along T with zero padding at the clip ends, attention
// within the clip.
//
// The activations never leave the SM between the layers.  A 256-channel tile lives in shared memory as four 64-channel
// chunks of 132 128-byte rows with the 128-byte swizzle, i.e. directly as the K-major A operand of tcgen05.mma, with the
// two clips' time steps INTERLEAVED: rows 0, 1 zero | row 2 (t + 1) + clip | rows 130, 131 zero.  Then
// (tools/probes/umma_m64_probe.cu):
//   * a convolution tap is the same tile read two rows earlier / later: the descriptor's start address moves by whole
//     rows (base-offset field 0 -- the swizzle is a function of the absolute shared-memory address);
//   * one M = 128 MMA per tap covers both clips, and the zero rows at the two ends are exactly the reference's zero
//     padding of both (a first version used one M = 64 MMA per clip on a clip-after-clip layout with lane-offset
//     accumulators: twice the MMAs at half rate each, 19 us of tensor time per launch);
//   * TMEM lane i holds (t = i >> 1, clip = i & 1).
// Only the weights stream from L2 (32 blocks of 20-32 KB through a four-stage TMA ring fed by a dedicated producer
// warp).  (Tried: 2-CTA clusters multicasting the weight blocks -- no faster, the stream is bound by the bytes in flight
// per SM, not by L2 bandwidth, and the cluster barriers cost 4 us.)  The epilogue of a layer writes the next layer's operand tile; from the q | k | v projection on the kernel is
// csrc/attn_fused.cu's algorithm (S and O as M = 128 MMAs over unpadded Q / K / V tiles).
#include <cuda.h>
#include <cstring>
#include "conv_gemm.cuh"
#include "layers.cuh"

void a2m_count_launch();

namespace a2m {

int make_weight_map(CUtensorMap* map, const void* w, long long n_rows, long long k, int box_rows);   // conv_gemm.cu

namespace {

constexpr int kWorkers = 512;
constexpr int kThreadsRb = kWorkers + 32;                  // + one producer warp
constexpr int kC = 256, kD = 32, kNqkv = 2 * kD + kC, kT = 64;
constexpr int kChunkBytes = 136 * 128;                     // 132 rows used; 17 KB keeps every chunk 1024-byte aligned
constexpr int kTileBytes = 4 * kChunkBytes;
constexpr int kOffA0 = 0;                                  // x, then t1, then t2 (each layer's MMAs are done before its epilogue
                                                           // overwrites the tile in place); later V (4 x [128][64])
constexpr int kOffRing = kTileBytes;                       // weight ring; later Q, K [128][64] and P [128][128]
constexpr int kStages = 4;
constexpr int kStageBytes = 32768;                         // conv K block 32 KB; q | k | v half block (160 rows) 20 KB
constexpr int kOffSum = kOffRing + kStages * kStageBytes;  // softmax row sums [128] fp32
constexpr int kOffBias = kOffSum + 512;                    // conv1 [256] | conv2 [256] | q k v [320] fp32
constexpr int kOffBar = kOffBias + (2 * kC + kNqkv) * 4;
constexpr int kSmemRb = kOffBar + 128 + 1024;
constexpr int kOffQ = kOffRing, kOffK = kOffRing + 16384, kOffP = kOffRing + 32768, kOffV = kOffA0;
constexpr uint32_t kColS = 320;                            // TMEM: conv / q|k|v accumulators [0,320), later O [0,256); S [320,448)
constexpr int kConvBlocks = 12, kBlocks = 2 * kConvBlocks + 8;
static_assert(kOffRing % 1024 == 0 && kStageBytes % 1024 == 0, "swizzled tiles need 1024 B alignment");
static_assert(kSmemRb <= 227 * 1024, "shared memory budget");

struct ResblockParams {
    CUtensorMap w1_map, w2_map;   // conv weights [256][768] bf16 (K = tap x channel), box 64 x 128
    CUtensorMap wq_map;           // q | k | v weights [320][256] bf16, box 64 x 160 (one box per ring stage)
    const float* bias1;           // [256] BatchNorm folded
    const float* bias2;           // [256]
    const float* bias_qkv;        // [320]
    const float* gamma;           // device scalar
    const __nv_bfloat16* x;       // [rows][256]
    __nv_bfloat16* t2;            // [rows][256] scratch: the attention's own residual is re-read from here
    __nv_bfloat16* out;           // [rows][256]
    long long n_rows;
};

__device__ __forceinline__ uint32_t pack2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ int sw128_off(int r, int c) { return r * 128 + ((c ^ (r & 7)) << 4); }
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes) {      // MN-major SW128
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}
__device__ __forceinline__ float leaky(float x) { return x > 0.f ? x : kLeakySlope * x; }

__global__ void __launch_bounds__(kThreadsRb, 1)
resblock_fused_kernel(const __grid_constant__ ResblockParams p, int* __restrict__ err_flag) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    unsigned char* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    float* s_sum = reinterpret_cast<float*>(smem + kOffSum);
    float* s_bias = reinterpret_cast<float*>(smem + kOffBias);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kOffBar);      // [kStages] weight stage landed
    uint64_t* empty_bar = full_bar + kStages;                              // [kStages] stage read by its MMAs
    uint64_t* acc_bar = empty_bar + kStages;                                     // a layer's accumulators complete (three phases)
    uint64_t* s_bar = acc_bar + 1;                                         // S = q k^T complete
    uint64_t* o_bar = s_bar + 1;                                           // O = P v complete
    uint64_t* qk_bar = o_bar + 1;                                          // 128 arrivals: q | k staged
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(qk_bar + 1);

    const int tid = threadIdx.x, warp = tid >> 5;
    const long long row0 = static_cast<long long>(blockIdx.x) * 128;

    pdl_launch_dependents();
    if (tid == 0) {
        tma_prefetch_desc(&p.w1_map);
        tma_prefetch_desc(&p.w2_map);
        tma_prefetch_desc(&p.wq_map);
        for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(acc_bar, 1);
        mbar_init(s_bar, 1);
        mbar_init(o_bar, 1);
        mbar_init(qk_bar, 128);
        mbar_fence_init();
    }
    if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
    for (int i = tid; i < 2 * kC + kNqkv; i += kThreadsRb)
        s_bias[i] = __ldg(i < kC ? p.bias1 + i : i < 2 * kC ? p.bias2 + (i - kC) : p.bias_qkv + (i - 2 * kC));
    {   // the padding rows of the activation tile (rows 0, 1, 130, 131 of every chunk) stay zero through all three layers
        const uint4 z = make_uint4(0, 0, 0, 0);
        for (int i = tid; i < 4 * 4 * 8; i += kThreadsRb) {
            const int chunk = (i >> 5) & 3, which = (i >> 3) & 3, c16 = i & 7;
            const int prow = which == 0 ? 0 : which == 1 ? 1 : which == 2 ? 130 : 131;
            *reinterpret_cast<uint4*>(smem + kOffA0 + chunk * kChunkBytes + prow * 128 + (c16 << 4)) = z;
        }
    }
    fence_proxy_async_smem();                              // the zero rows are read by the tensor core (async proxy)
    tc_fence_before();
    __syncthreads();                                       // the only CTA-wide barriers are this one and the last one
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();                                            // x comes from the previous kernel

    if (warp == kWorkers / 32) {
        // ================= producer warp: 28 weight K blocks through the two-stage ring =================
        if ((tid & 31) == 0) {
            for (int g = 0; g < kBlocks; ++g) {
                const int s = g % kStages;
                if (!mbar_wait(&empty_bar[s], ((g / kStages) & 1) ^ 1, err_flag, 51)) break;
                unsigned char* st = smem + kOffRing + s * kStageBytes;
                if (g < 2 * kConvBlocks) {
                    const CUtensorMap* wm = g < kConvBlocks ? &p.w1_map : &p.w2_map;
                    const int kb = g < kConvBlocks ? g : g - kConvBlocks;
                    mbar_expect_tx(&full_bar[s], 32768);
                    tma_load_5d(st, wm, &full_bar[s], kb * 64, 0, 0, 0, 0);
                    tma_load_5d(st + 16384, wm, &full_bar[s], kb * 64, 128, 0, 0, 0);
                } else {                                   // q | k | v: K block kb, rows 160 j .. 160 j + 159
                    const int hb = g - 2 * kConvBlocks, kb = hb >> 1, j = hb & 1;
                    mbar_expect_tx(&full_bar[s], 160 * 128);
                    tma_load_5d(st, &p.wq_map, &full_bar[s], kb * 64, j * 160, 0, 0, 0);
                }
            }
        }
    } else {
        // ================= 512 workers: four per row =================
        const int r = tid & 127, q = tid >> 7, quad = warp & 3;
        const uint32_t tmem_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
        const uint32_t base = smem_u32(smem);
        // tile row of this thread in the two lane orders: natural (S, O accumulators over the unpadded Q / K / V tiles; also
        // the x load) and interleaved (conv / q|k|v accumulators over the padded tile: lane i = time step i >> 1 of clip i & 1)
        const int clip_m = r & 1, t_m = r >> 1;
        const int row_m = clip_m * kT + t_m;               // natural tile row held by my lane in the interleaved layout
        const int prow_m = r + 2;                          // its padded row
        auto worker_sync = [] { named_barrier(1, kWorkers); };

        {   // ---- x -> A0 (my natural row r, channel quarter q = chunk q)
            const long long row = row0 + r;
            const int prow = 2 * ((r & 63) + 1) + (r >> 6);
            unsigned char* dst = smem + kOffA0 + q * kChunkBytes;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                uint4 v = make_uint4(0, 0, 0, 0);
                if (row < p.n_rows) v = __ldcg(reinterpret_cast<const uint4*>(p.x + row * kC + q * 64 + i * 8));
                *reinterpret_cast<uint4*>(dst + sw128_off(prow, i)) = v;
            }
        }
        fence_proxy_async_smem();
        worker_sync();

        int g = 0;                                         // weight K block counter (MMA thread)
        // one k = 3 convolution: accumulators [128 lanes][256 columns] from the tile at `a_off`
        auto conv_mma = [&](int a_off) {
            const uint32_t idesc = umma_idesc_bf16(128, 256);
            for (int kb = 0; kb < kConvBlocks; ++kb, ++g) {
                const int s = g % kStages, tap = kb >> 2, chunk = kb & 3;       // K offset kb * 64 = tap * 256 + chunk * 64
                if (!mbar_wait(&full_bar[s], (g / kStages) & 1, err_flag, 52)) break;
                tc_fence_after();
                const uint32_t b_addr = base + kOffRing + s * kStageBytes;
                const uint32_t a_addr = base + a_off + chunk * kChunkBytes + (2 * tap) * 128;      // time steps t - 1 + tap of both clips
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16(tmem_base, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(b_addr + k * 32), idesc, (kb | k) != 0);
                umma_commit(&empty_bar[s]);
            }
            umma_commit(acc_bar);
        };
        // + folded bias, LeakyReLU, bf16 -> my row of the tile at `dst_off` (chunk q) and optionally to global memory
        auto conv_epilogue = [&](const float* bias, int dst_off, __nv_bfloat16* gdst) {
            unsigned char* dst = smem + dst_off + q * kChunkBytes;
            const long long row = row0 + row_m;
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                uint32_t t[32];
                tmem_ld_32x32(tmem_lane + q * 64 + hf * 32, t);
                tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float4 b0 = *reinterpret_cast<const float4*>(bias + q * 64 + hf * 32 + c * 8);
                    const float4 b1 = *reinterpret_cast<const float4*>(bias + q * 64 + hf * 32 + c * 8 + 4);
                    uint4 o;
                    o.x = pack2(leaky(__uint_as_float(t[c * 8]) + b0.x), leaky(__uint_as_float(t[c * 8 + 1]) + b0.y));
                    o.y = pack2(leaky(__uint_as_float(t[c * 8 + 2]) + b0.z), leaky(__uint_as_float(t[c * 8 + 3]) + b0.w));
                    o.z = pack2(leaky(__uint_as_float(t[c * 8 + 4]) + b1.x), leaky(__uint_as_float(t[c * 8 + 5]) + b1.y));
                    o.w = pack2(leaky(__uint_as_float(t[c * 8 + 6]) + b1.z), leaky(__uint_as_float(t[c * 8 + 7]) + b1.w));
                    *reinterpret_cast<uint4*>(dst + sw128_off(prow_m, hf * 4 + c)) = o;
                    if (gdst != nullptr && row < p.n_rows)
                        *reinterpret_cast<uint4*>(gdst + row * kC + q * 64 + hf * 32 + c * 8) = o;
                }
            }
        };

        // ---------------- conv1: x -> t1 ----------------
        if (tid == 0) { tc_fence_after(); conv_mma(kOffA0); }
        mbar_wait(acc_bar, 0, err_flag, 53);
        tc_fence_after();
        conv_epilogue(s_bias, kOffA0, nullptr);
        tc_fence_before();
        fence_proxy_async_smem();
        worker_sync();
        // ---------------- conv2: t1 -> t2 (shared memory for the projection, global memory for the residual) ----------------
        if (tid == 0) { tc_fence_after(); conv_mma(kOffA0); }
        mbar_wait(acc_bar, 1, err_flag, 54);
        tc_fence_after();
        conv_epilogue(s_bias + kC, kOffA0, p.t2);
        tc_fence_before();
        fence_proxy_async_smem();
        worker_sync();
        // ---------------- q | k | v = t2 Wqkv^T ----------------
        if (tid == 0) {
            tc_fence_after();
            const uint32_t id160 = umma_idesc_bf16(128, 160);
            for (int hb = 0; hb < 8; ++hb, ++g) {          // K block hb >> 1, weight rows (= accumulator columns) 160 (hb & 1) ...
                const int s = g % kStages, kb = hb >> 1, j = hb & 1;
                if (!mbar_wait(&full_bar[s], (g / kStages) & 1, err_flag, 55)) break;
                tc_fence_after();
                const uint32_t b_addr = base + kOffRing + s * kStageBytes;
                const uint32_t a_addr = base + kOffA0 + kb * kChunkBytes + 2 * 128;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16(tmem_base + j * 160, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(b_addr + k * 32), id160, (kb | k) != 0);
                umma_commit(&empty_bar[s]);
            }
            umma_commit(acc_bar);
        }
        mbar_wait(acc_bar, 0, err_flag, 56);
        tc_fence_after();

        // ---------------- stage q, k (quarter 0) and v (all quarters) as bf16 operands, rows in natural tile order ----------------
        if (q == 0) {
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {               // hf 0: q -> Q, hf 1: k -> K; columns 32..63 of both are zero
                uint32_t t[32];
                tmem_ld_32x32(tmem_lane + hf * 32, t);
                tmem_ld_wait();
                unsigned char* dst = smem + (hf == 0 ? kOffQ : kOffK);
                const float* bq = s_bias + 2 * kC + hf * 32;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint4 o;
                    const float4 b0 = *reinterpret_cast<const float4*>(bq + c * 8), b1 = *reinterpret_cast<const float4*>(bq + c * 8 + 4);
                    o.x = pack2(__uint_as_float(t[c * 8]) + b0.x, __uint_as_float(t[c * 8 + 1]) + b0.y);
                    o.y = pack2(__uint_as_float(t[c * 8 + 2]) + b0.z, __uint_as_float(t[c * 8 + 3]) + b0.w);
                    o.z = pack2(__uint_as_float(t[c * 8 + 4]) + b1.x, __uint_as_float(t[c * 8 + 5]) + b1.y);
                    o.w = pack2(__uint_as_float(t[c * 8 + 6]) + b1.z, __uint_as_float(t[c * 8 + 7]) + b1.w);
                    *reinterpret_cast<uint4*>(dst + sw128_off(row_m, c)) = o;
                    *reinterpret_cast<uint4*>(dst + sw128_off(row_m, c + 4)) = make_uint4(0, 0, 0, 0);
                }
            }
            fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(qk_bar);
        }
        if (tid == 128) {                                  // S = q k^T as soon as q | k are staged (v staging overlaps)
            mbar_wait(qk_bar, 0, err_flag, 57);
            tc_fence_after();
            const uint32_t id128 = umma_idesc_bf16(128, 128);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                umma_bf16(tmem_base + kColS, umma_desc_sw128(base + kOffQ + k * 32), umma_desc_sw128(base + kOffK + k * 32), id128, k != 0);
            umma_commit(s_bar);
        }
        {   // v: columns 64..319 of the projection (weights packed q | k | v as for csrc/attn_fused.cu)
            unsigned char* dst = smem + kOffV + q * 16384;
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                uint32_t t[32];
                tmem_ld_32x32(tmem_lane + 64 + q * 64 + hf * 32, t);
                tmem_ld_wait();
                const float* bv = s_bias + 2 * kC + 64 + q * 64 + hf * 32;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint4 o;
                    const float4 b0 = *reinterpret_cast<const float4*>(bv + c * 8), b1 = *reinterpret_cast<const float4*>(bv + c * 8 + 4);
                    o.x = pack2(__uint_as_float(t[c * 8]) + b0.x, __uint_as_float(t[c * 8 + 1]) + b0.y);
                    o.y = pack2(__uint_as_float(t[c * 8 + 2]) + b0.z, __uint_as_float(t[c * 8 + 3]) + b0.w);
                    o.z = pack2(__uint_as_float(t[c * 8 + 4]) + b1.x, __uint_as_float(t[c * 8 + 5]) + b1.y);
                    o.w = pack2(__uint_as_float(t[c * 8 + 6]) + b1.z, __uint_as_float(t[c * 8 + 7]) + b1.w);
                    *reinterpret_cast<uint4*>(dst + sw128_off(row_m, hf * 4 + c)) = o;
                }
            }
        }
        if (q != 0) {                                      // zero P; the row threads then write their own clip's block
            uint4* pz = reinterpret_cast<uint4*>(smem + kOffP);
            for (int i = tid - 128; i < 32768 / 16; i += kWorkers - 128) pz[i] = make_uint4(0, 0, 0, 0);
        }
        worker_sync();

        // ---------------- softmax of each row over its own clip (quarter 0: one thread per natural row) ----------------
        if (q == 0) {
            mbar_wait(s_bar, 0, err_flag, 58);
            tc_fence_after();
            const int win0 = (r >> 6) * kT;                // my clip's 64 columns
            float e[64];
            float m = -INFINITY;
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                uint32_t t[32];
                tmem_ld_32x32(tmem_lane + kColS + win0 + hf * 32, t);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) { e[hf * 32 + j] = __uint_as_float(t[j]); m = fmaxf(m, e[hf * 32 + j]); }
            }
            float sum = 0.f;
#pragma unroll
            for (int c = 0; c < 64; ++c) { e[c] = __expf(e[c] - m); sum += e[c]; }
            s_sum[r] = sum;
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) {
                uint4 o;
                const int c = ch * 8;
                o.x = pack2(e[c], e[c + 1]); o.y = pack2(e[c + 2], e[c + 3]);
                o.z = pack2(e[c + 4], e[c + 5]); o.w = pack2(e[c + 6], e[c + 7]);
                const int col = win0 + c;
                *reinterpret_cast<uint4*>(smem + kOffP + (col >> 6) * 16384 + sw128_off(r, (col & 63) >> 3)) = o;
            }
        }
        tc_fence_before();
        fence_proxy_async_smem();
        worker_sync();                                     // v and P staged, all TMEM reads of q | k | v done

        // ---------------- O = P v ----------------
        if (tid == 0) {
            tc_fence_after();
            const uint32_t id_o = umma_idesc_bf16(128, 256) | (1u << 16);         // B (v) is MN-major
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
                umma_bf16(tmem_base, umma_desc_sw128(base + kOffP + (kk >> 2) * 16384 + (kk & 3) * 32),
                          umma_desc_mn(base + kOffV + kk * 2048, 16384), id_o, kk != 0);
            umma_commit(o_bar);
        }
        // residual rows (natural order) while the MMA runs: t2 (written above by other threads of this CTA, ordered by the
        // barriers since) and x; both through L2
        const long long row = row0 + r;
        const bool live = row < p.n_rows;
        const long long gofs = row * kC + q * 64;
        uint4 xv[8], rv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            xv[i] = make_uint4(0, 0, 0, 0);
            rv[i] = make_uint4(0, 0, 0, 0);
            if (live) {
                xv[i] = __ldcg(reinterpret_cast<const uint4*>(p.t2 + gofs + i * 8));
                rv[i] = __ldcg(reinterpret_cast<const uint4*>(p.x + gofs + i * 8));
            }
        }
        mbar_wait(o_bar, 0, err_flag, 59);
        tc_fence_after();
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
                    if (live) *reinterpret_cast<uint4*>(p.out + gofs + hf * 32 + c * 8) = make_uint4(os[0], os[1], os[2], os[3]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

}  // namespace

struct ResblockFusedPlan {
    ResblockParams p;
    int grid;
};

bool resblock_fused_supported(int T, int C) { return T == kT && C == kC; }

// w_qkv / bias_qkv: the attention's projection packed [q (32) | k (32) | v (256)][256], as for csrc/attn_fused.cu
int resblock_fused_plan(const __nv_bfloat16* w1, const float* bias1, const __nv_bfloat16* w2, const float* bias2,
                        const __nv_bfloat16* w_qkv, const float* bias_qkv, const float* gamma, const __nv_bfloat16* x,
                        __nv_bfloat16* t2, int B, int T, int C, __nv_bfloat16* out, std::shared_ptr<ResblockFusedPlan>* plan_out) {
    A2M_ARG_CHECK(resblock_fused_supported(T, C), "resblock_fused: T = %d, C = %d not supported", T, C);
    auto plan = std::make_shared<ResblockFusedPlan>();
    ResblockParams& p = plan->p;
    memset(&p, 0, sizeof(p));
    int rc = make_weight_map(&p.w1_map, w1, kC, 3 * kC, 128);
    if (rc != A2M_OK) return rc;
    rc = make_weight_map(&p.w2_map, w2, kC, 3 * kC, 128);
    if (rc != A2M_OK) return rc;
    rc = make_weight_map(&p.wq_map, w_qkv, kNqkv, kC, 160);
    if (rc != A2M_OK) return rc;
    p.bias1 = bias1; p.bias2 = bias2; p.bias_qkv = bias_qkv; p.gamma = gamma; p.x = x; p.t2 = t2; p.out = out;
    p.n_rows = static_cast<long long>(B) * T;
    plan->grid = static_cast<int>((p.n_rows + 127) / 128);
    *plan_out = plan;
    return A2M_OK;
}

int resblock_fused_launch(const ResblockFusedPlan& plan, int* err_flag, cudaStream_t stream) {
    static A2mPerDeviceOnce configured;
    if (configured.first())
        A2M_CUDA_CHECK(cudaFuncSetAttribute(resblock_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemRb));
    A2M_CUDA_CHECK(a2m_launch_pdl(resblock_fused_kernel, dim3(plan.grid), dim3(kThreadsRb), kSmemRb, stream, plan.p, err_flag));
    a2m_count_launch();
    return A2M_OK;
}

}  // namespace a2m
