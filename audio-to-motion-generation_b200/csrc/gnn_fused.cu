// Fused static-skeleton GNN stack: the five layers of real_motion_model.py:172-201 /
// :224-253  --  GAT, GraphConv, GAT, GraphConv, GAT, each followed by LayerNorm(64) -> LeakyReLU(0.2)
// -> + residual  --  in ONE kernel, node features never leaving the SM between layers.
//
// torch_geometric semantics restated (PyG is not a dependency):
//   GATConv(64, 64, heads=4, concat=False): h = W x (no bias), e_ij = LeakyReLU_0.2(a_src.h_j + a_dst.h_i) over
//     j in N(i) + {i}, alpha = softmax_j(e_ij), out_i = mean_heads(sum_j alpha_ij h_j) + bias
//   GraphConv(64, 64), aggr = add:        out_i = W_rel (sum_{j in N(i)} x_j) + b_rel + W_root x_i
//
// Everything that is a contraction runs on the tensor cores, including the neighbourhood aggregation: a tile is 128 node
// rows = whole graphs (3 hand graphs or 12 body graphs).  The attention logits come straight from the node tile
// (S = X U^T with U = W^T a_src / W^T a_dst folded at load, hi + lo bf16 split), the softmax rows (fp32, CUDA cores,
// <= 7 entries per row) are scattered as bf16 into a dense block-diagonal [128 x 128] matrix P^h in shared memory, and
// the layer is evaluated aggregate-first: Z^h = P^h X, OUT = sum_h bf16(Z^h) W_h^T with Z^h handed back to the tensor
// cores through TENSOR MEMORY (see the comment above the kernel).  The row threads (residual stream in fp32 registers)
// do the softmax, bias, LayerNorm, LeakyReLU and residual.  Persistent CTAs; node tiles arrive by TMA and leave by TMA
// store; per-layer weights stream through shared memory by TMA, prefetched as soon as the MMA that read the previous
// slice has completed.  Earlier generations (one CTA per SM; linear-first with H staged through shared memory) are in
// tools/probes/retired/gnn_fused_v2_v3.cu.
#include <cuda.h>
#include <cstdlib>
#include <cstring>
#include "conv_gemm.cuh"
#include "layers.cuh"
#include "fft_math.cuh"        // packed fp32 pairs (add / mul / fma .f32x2) for the row epilogue

void a2m_count_launch();

namespace a2m {

namespace {

constexpr int kRows = 128;
constexpr int kGatRows = 272;                          // 256 W rows + 16 folded attention rows

struct GnnParams {
    CUtensorMap w_head[3];               // GAT [272][64] bf16, box 64 x 64 (head slices)
    CUtensorMap w_att[3];                // same tensor, box 64 x 16 (rows 256..271)
    CUtensorMap w_gc[2];                 // [64][128] bf16, box 64 x 64
    CUtensorMap x_in, x_out;             // [groups][group_rows][64] bf16, box 64 x rows_per_tile x 1
    const float* gat_bias[3];
    const float* gc_bias[2];
    const float* ln_w[5];
    const float* ln_b[5];
    const int* nbr;
    const int* deg;
    int J, gpc;
    long long n_groups;                  // graphs are tiled per group (= clip), so a graph's slot in its tile -- and with
    int group_graphs, tiles_per_group;   // it the MMA accumulation order -- never depends on how clips are batched
};

__device__ __forceinline__ float leaky(float x) { return x > 0.f ? x : kLeakySlope * x; }
__device__ __forceinline__ float bf_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
// byte offset of 16-byte chunk `c` (8 features) of row r in a [128][64] bf16 SW128 tile
__device__ __forceinline__ int sw128_off(int r, int c) { return r * 128 + ((c ^ (r & 7)) << 4); }
// byte offset of element (row r, column col) of a [128][128] bf16 matrix stored as two K-major SW128 chunks
__device__ __forceinline__ int p_off(int r, int col) {
    return (col >> 6) * 16384 + r * 128 + ((((col & 63) >> 3) ^ (r & 7)) << 4) + (col & 7) * 2;
}
__device__ __forceinline__ uint32_t idesc_b_mn(uint32_t m, uint32_t n) { return umma_idesc_bf16(m, n) | (1u << 16); }


// =====================================================================================================
// Aggregate first, A operand from tensor memory.  out_i = sum_h W_h (sum_j alpha^h_ij / 4 x_j) + b, so per GAT layer
//   S   = X U^T                      (attention logits straight from the node tile: all four softmaxes start at once)
//   Z^h = P^h X                      (B = the node tile itself, MN-major; two heads per round, two P buffers)
//   OUT = sum_h bf16(Z^h) W_h^T      (A = Z^h written back to TENSOR MEMORY as packed bf16 by the row threads)
// No H staging through shared memory, no per-head barrier chain: six CTA barriers and four MMA round trips per GAT
// layer (the linear-first generation needed ten and ~six).  GraphConv: AGG = Adj X -> bf16 in TMEM -> OUT = AGG W_rel^T + X W_root^T.
// TMEM (256 columns, two CTAs per SM): OUT [0,64)  Zf [64,192) (two heads, fp32)  Zb [192,256) (two heads, bf16 A
// operand; the 16 logit columns S alias its start -- S is dead before the first Zb store).
// Weight slices (8 KB: one head / W_rel / W_root) stream through a three-slot ring, three ahead.
// Tried and dropped: four threads per row (512 threads, 16 features and one head each, 64 registers per thread, still two
// CTAs per SM).  Twice the warps did not hide the round trips: 717 us instead of 509 us for the hand stack (spills, twice
// the barrier participants); one warp polling the mbarriers while the rest park on bar.sync changed nothing either.
// =====================================================================================================
namespace v4 {

#ifdef A2M_GNN_TRACE            // probe build only (tools/probes/gnn_trace.sh): clock stamps of CTA 0's MMA issuer
__device__ long long g_gnn_trace[128];
#define GNN_STAMP(i) do { if (blockIdx.x == 0 && tid == 0 && it == 1) g_gnn_trace[(i)] = clock64(); } while (0)
#define GNN_TILE_STAMP(i) do { if (blockIdx.x == 0 && tid == 0 && it < 20) g_gnn_trace[64 + it * 3 + (i)] = clock64(); } while (0)
#else
#define GNN_STAMP(i) do { } while (0)
#define GNN_TILE_STAMP(i) do { } while (0)
#endif

constexpr int kThreads4 = 256;
constexpr int kOffW4 = 0;                               // 3 x 8 KB weight slots
constexpr int kOffU4 = 24576;                           // GAT attention rows [16][64] bf16
constexpr int kOffX4 = 26624;                           // node tile, bf16 [128][64] SW128
constexpr int kOffP4 = kOffX4 + 16384;                  // 2 x attention / adjacency matrix [128][128] bf16
constexpr int kOffS4 = kOffP4 + 2 * 32768;              // s_src [128][4] fp32; LayerNorm partials [128][2][2] alias it
constexpr int kOffTopo4 = kOffS4 + kRows * 4 * 4;       // nbr [48][6], deg [48]
constexpr int kOffPar4 = kOffTopo4 + 48 * kMaxDeg * 4 + 48 * 4;   // this layer's bias[64], ln_w[64], ln_b[64]
constexpr int kOffBar4 = kOffPar4 + 192 * 4;
constexpr int kSmemBytes4 = kOffBar4 + 128 + 1024;
constexpr uint32_t kColOut4 = 0, kColZf = 64, kColZb = 192, kColS4 = 192;
static_assert(kOffX4 % 1024 == 0 && kOffP4 % 1024 == 0 && kOffU4 % 1024 == 0, "swizzled tiles need 1024 B alignment");
static_assert(2 * (kSmemBytes4 + 1024) <= 228 * 1024, "two CTAs per SM");

__device__ __forceinline__ uint64_t desc_add(uint64_t desc, uint32_t bytes) { return desc + (bytes >> 4); }

__global__ void __launch_bounds__(kThreads4, 2)
gnn4_kernel(const __grid_constant__ GnnParams p, int* __restrict__ err_flag) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    unsigned char* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    unsigned char* s_x = smem + kOffX4;
    unsigned char* s_p = smem + kOffP4;
    float* s_src = reinterpret_cast<float*>(smem + kOffS4);
    float* s_ln = s_src;
    int* s_nbr = reinterpret_cast<int*>(smem + kOffTopo4);
    int* s_deg = s_nbr + 48 * kMaxDeg;
    float* s_par = reinterpret_cast<float*>(smem + kOffPar4);
    uint64_t* w_bar = reinterpret_cast<uint64_t*>(smem + kOffBar4);    // [3] weight slots (thread 0 only)
    uint64_t* u_bar = w_bar + 3;                                       // attention rows landed (thread 0 only)
    uint64_t* s_bar = w_bar + 4;                                       // logits MMA done
    uint64_t* z_bar = w_bar + 5;                                       // aggregation MMAs done (P buffers free, Zf valid)
    uint64_t* o_bar = w_bar + 6;                                       // OUT complete
    uint64_t* x_bar = w_bar + 7;                                       // node tile landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 8);

    const int tid = threadIdx.x, warp = tid >> 5;
    const int r = tid & 127, half = tid >> 7, quad = warp & 3;
    const int J = p.J, rows_per_tile = p.gpc * J;
    const int group_rows = p.group_graphs * J;
    const long long n_tiles = p.n_groups * p.tiles_per_group;
    const long long my_tiles = (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
    const uint32_t tile_bytes = static_cast<uint32_t>(rows_per_tile) * 128u;
    const long long n_items = my_tiles * 16;            // weight slices this CTA consumes, 16 per tile

    // weight slice `item` (within a tile: GAT heads 0-3 | GC rel, root | GAT | GC | GAT) into slot item % 3
    auto load_item = [&](long long item) {             // thread 0 only
        if (item >= n_items) return;
        const int i = static_cast<int>(item & 15), slot = static_cast<int>(item % 3);
        const int layer = i < 4 ? 0 : i < 6 ? 1 : i < 10 ? 2 : i < 12 ? 3 : 4;
        unsigned char* dst = smem + kOffW4 + slot * 8192;
        mbar_expect_tx(&w_bar[slot], 8192);
        if ((layer & 1) == 0) {
            const int h = i - (layer == 0 ? 0 : layer == 2 ? 6 : 12);
            tma_load_5d(dst, &p.w_head[layer >> 1], &w_bar[slot], 0, h * 64, 0, 0, 0);
        } else {
            tma_load_5d(dst, &p.w_gc[layer >> 1], &w_bar[slot], (i & 1) * 64, 0, 0, 0, 0);   // even item: W_rel (k 0..63), odd: W_root
        }
    };
    auto load_u = [&](int gat) {                       // thread 0 only
        mbar_expect_tx(u_bar, 2048);
        tma_load_5d(smem + kOffU4, &p.w_att[gat], u_bar, 0, 256, 0, 0, 0);
    };

    pdl_launch_dependents();
    if (tid == 0) {
        for (int i = 0; i < 3; ++i) { tma_prefetch_desc(&p.w_head[i]); tma_prefetch_desc(&p.w_att[i]); }
        for (int i = 0; i < 2; ++i) tma_prefetch_desc(&p.w_gc[i]);
        tma_prefetch_desc(&p.x_in);
        tma_prefetch_desc(&p.x_out);
        for (int i = 0; i < 8; ++i) mbar_init(&w_bar[i], 1);
        mbar_fence_init();
        load_u(0);                                     // weights are constants: no need to wait for the predecessor
        load_item(0);
        load_item(1);
        load_item(2);
    }
    if (warp == 1) { tmem_alloc(tmem_slot, 256); tmem_relinquish(); }
    for (int i = tid; i < J * kMaxDeg; i += kThreads4) s_nbr[i] = p.nbr[i];
    for (int i = tid; i < J; i += kThreads4) s_deg[i] = p.deg[i];
    {
        uint4 z = make_uint4(0, 0, 0, 0);
        for (int i = tid; i < 2 * 32768 / 16; i += kThreads4) reinterpret_cast<uint4*>(s_p)[i] = z;
        for (int i = tid; i < 16384 / 16; i += kThreads4)
            if ((i >> 3) >= rows_per_tile) reinterpret_cast<uint4*>(s_x)[i] = z;
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    // bounded waits (a2m_common.cuh); after a thread's first time-out the rest of its waits are skipped, so that one lost
    // signal ends this launch within seconds -- with the error code in *err_flag -- instead of costing two seconds per wait
    bool dead = false;
    auto wait = [&](uint64_t* bar, uint32_t parity, int code) {
        if (!dead && !mbar_wait(bar, parity, err_flag, code)) dead = true;
    };
    pdl_wait();                                        // node features come from the previous kernel (proj_in)
    auto load_tile = [&](long long tile) {             // thread 0 only
        const int group = static_cast<int>(tile / p.tiles_per_group);
        const int row0 = static_cast<int>(tile - static_cast<long long>(group) * p.tiles_per_group) * rows_per_tile;
        mbar_expect_tx(x_bar, tile_bytes);
        tma_load_5d(s_x, &p.x_in, x_bar, 0, row0, group, 0, 0);
    };
    if (tid == 0) load_tile(blockIdx.x);
    const uint32_t idesc_h = umma_idesc_bf16(128, 64), idesc_s = umma_idesc_bf16(128, 16);
    const uint32_t idesc_agg = idesc_b_mn(128, 64);
    const uint64_t w_desc = umma_desc_sw128(smem_u32(smem + kOffW4)), u_desc = umma_desc_sw128(smem_u32(smem + kOffU4)),
                   x_desc = umma_desc_sw128(smem_u32(s_x)), p_desc = umma_desc_sw128(smem_u32(s_p));

    const bool valid_row = r < rows_per_tile;
    const int jloc = r % J, g0 = r - jloc;
    const int dg = valid_row ? s_deg[jloc] : 0;
    int idx[kMaxDeg + 1], pofs[kMaxDeg + 1];
    idx[0] = r;
#pragma unroll
    for (int k = 0; k < kMaxDeg; ++k) idx[k + 1] = k < dg ? g0 + s_nbr[jloc * kMaxDeg + k] : r;
#pragma unroll
    for (int k = 0; k <= kMaxDeg; ++k) pofs[k] = half * 32768 + p_off(r, idx[k]);     // my P buffer is buffer `half`

    // softmax over {self} + neighbours for head h (the head mean 1/4 folded in) -> my P buffer
    // the same in two steps: the coefficients first (registers), the stores once the P buffer is free again
    auto softmax_p = [&](int h, float sd, float (&alpha)[kMaxDeg + 1]) {
        float m = -INFINITY;
#pragma unroll
        for (int k = 0; k <= kMaxDeg; ++k) {
            alpha[k] = k <= dg ? leaky(s_src[idx[k] * 4 + h] + sd) : -INFINITY;
            m = fmaxf(m, alpha[k]);
        }
        float den = 0.f;
#pragma unroll
        for (int k = 0; k <= kMaxDeg; ++k) { alpha[k] = k <= dg ? __expf(alpha[k] - m) : 0.f; den += alpha[k]; }
        const float inv = 0.25f / den;
#pragma unroll
        for (int k = 0; k <= kMaxDeg; ++k) alpha[k] *= inv;
    };
    auto store_p = [&](const float (&alpha)[kMaxDeg + 1]) {
#pragma unroll
        for (int k = 0; k <= kMaxDeg; ++k)
            if (k <= dg) *reinterpret_cast<__nv_bfloat16*>(s_p + pofs[k]) = __float2bfloat16_rn(alpha[k]);
    };
    auto write_p = [&](int h, float sd) {
        float alpha[kMaxDeg + 1];
        float m = -INFINITY;
#pragma unroll
        for (int k = 0; k <= kMaxDeg; ++k) {
            alpha[k] = k <= dg ? leaky(s_src[idx[k] * 4 + h] + sd) : -INFINITY;
            m = fmaxf(m, alpha[k]);
        }
        float den = 0.f;
#pragma unroll
        for (int k = 0; k <= kMaxDeg; ++k) { alpha[k] = k <= dg ? __expf(alpha[k] - m) : 0.f; den += alpha[k]; }
        const float inv = 0.25f / den;
#pragma unroll
        for (int k = 0; k <= kMaxDeg; ++k)
            if (k <= dg) *reinterpret_cast<__nv_bfloat16*>(s_p + pofs[k]) = __float2bfloat16_rn(alpha[k] * inv);
    };
    // my 32 fp32 columns of Zf (starting at column c0 of the Zf region) -> packed bf16 -> 16 columns of Zb at cb
    auto convert = [&](uint32_t c0, uint32_t cb) {
        uint32_t t[32], o[16];
        tmem_ld_32x32(tmem_lane + kColZf + c0, t);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) o[i] = pack_bf16(__uint_as_float(t[2 * i]), __uint_as_float(t[2 * i + 1]));
        tmem_st_32x16(tmem_lane + kColZb + cb, o);
    };

    // the same for 64 columns (one head): both loads in flight before the single wait
    auto convert2 = [&](uint32_t c0, uint32_t cb) {
        uint32_t ta[32], tb[32], o[16];
        tmem_ld_32x32(tmem_lane + kColZf + c0, ta);
        tmem_ld_32x32(tmem_lane + kColZf + c0 + 32, tb);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) o[i] = pack_bf16(__uint_as_float(ta[2 * i]), __uint_as_float(ta[2 * i + 1]));
        tmem_st_32x16(tmem_lane + kColZb + cb, o);
#pragma unroll
        for (int i = 0; i < 16; ++i) o[i] = pack_bf16(__uint_as_float(tb[2 * i]), __uint_as_float(tb[2 * i + 1]));
        tmem_st_32x16(tmem_lane + kColZb + cb + 16, o);
    };

    uint32_t wpar[3] = {0, 0, 0}, upar = 0;             // thread 0 only
    uint32_t spar = 0, zpar = 0, opar = 0;
    long long item = 0;                                // first weight slice of the current layer (thread 0's view)
    auto wait_item = [&](long long it_) {              // thread 0 only
        const int slot = static_cast<int>(it_ % 3);
        wait(&w_bar[slot], wpar[slot], 41);
        wpar[slot] ^= 1;
        return desc_add(w_desc, static_cast<uint32_t>(slot) * 8192u);
    };

    for (long long it = 0; it < my_tiles; ++it) {
        const long long tile = blockIdx.x + it * gridDim.x;
        const int group = static_cast<int>(tile / p.tiles_per_group);
        const int row0 = static_cast<int>(tile - static_cast<long long>(group) * p.tiles_per_group) * rows_per_tile;
        const bool live = valid_row && row0 + r < group_rows;
        GNN_TILE_STAMP(0);
        wait(x_bar, static_cast<uint32_t>(it & 1), 40);
        GNN_TILE_STAMP(1);
        using a2m_fft::pair_t;
        pair_t x[16];                                  // residual stream: this thread's 32 features in fp32, as 16 packed pairs
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const uint4 u = *reinterpret_cast<const uint4*>(s_x + sw128_off(r, half * 4 + c));
            const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) x[c * 4 + e] = a2m_fft::pack(bf_lo(w4[e]), bf_hi(w4[e]));
        }

#pragma unroll 1
        for (int layer = 0; layer < 5; ++layer) {
            pair_t v[16];
            if (tid < 192) {                            // this layer's bias | ln_w | ln_b (read after several barriers)
                const float* src = tid < 64 ? ((layer & 1) ? p.gc_bias[layer >> 1] : p.gat_bias[layer >> 1])
                                            : tid < 128 ? p.ln_w[layer] : p.ln_b[layer];
                s_par[tid] = __ldg(src + (tid & 63));
            }
            GNN_STAMP(layer * 12);
            if ((layer & 1) == 0) {
                // ================= GATConv =================
                if (tid == 0) {
                    tc_fence_after();
                    wait(u_bar, upar, 42); upar ^= 1;
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16(tmem_base + kColS4, desc_add(x_desc, k * 32), desc_add(u_desc, k * 32), idesc_s, k != 0);
                    umma_commit(s_bar);
                }
                wait(s_bar, spar, 43);
                spar ^= 1;
                tc_fence_after();
                GNN_STAMP(layer * 12 + 1);
                if (tid == 0 && !(layer == 4 && it + 1 == my_tiles)) load_u(layer == 4 ? 0 : (layer >> 1) + 1);
                float sd0, sd1;                          // destination logits of my heads: half, half + 2
                {
                    uint32_t t[16];
                    tmem_ld_32x16(tmem_lane + kColS4, t);
                    tmem_ld_wait();
                    sd0 = half == 0 ? __uint_as_float(t[4]) + __uint_as_float(t[12]) : __uint_as_float(t[5]) + __uint_as_float(t[13]);
                    sd1 = half == 0 ? __uint_as_float(t[6]) + __uint_as_float(t[14]) : __uint_as_float(t[7]) + __uint_as_float(t[15]);
                    if (half == 0) {
                        *reinterpret_cast<float4*>(s_src + r * 4) =
                            make_float4(__uint_as_float(t[0]) + __uint_as_float(t[8]), __uint_as_float(t[1]) + __uint_as_float(t[9]),
                                        __uint_as_float(t[2]) + __uint_as_float(t[10]), __uint_as_float(t[3]) + __uint_as_float(t[11]));
                    }
                }
                tc_fence_before();
                __syncthreads();                        // s_src visible
                GNN_STAMP(layer * 12 + 2);
                write_p(half, sd0);                     // round 1: heads 0 (buffer 0) and 1 (buffer 1)
                fence_proxy_async_smem();
                __syncthreads();
                GNN_STAMP(layer * 12 + 3);
                if (tid == 0) {
                    tc_fence_after();
#pragma unroll
                    for (int b = 0; b < 2; ++b)
#pragma unroll
                        for (int kk = 0; kk < 8; ++kk)
                            umma_bf16(tmem_base + kColZf + b * 64, desc_add(p_desc, b * 32768 + (kk >> 2) * 16384 + (kk & 3) * 32),
                                      desc_add(x_desc, kk * 2048), idesc_agg, kk != 0);
                    umma_commit(z_bar);
                }
                float alpha2[kMaxDeg + 1];
                softmax_p(half + 2, sd1, alpha2);       // round 2's coefficients while round 1's MMAs run
                wait(z_bar, zpar, 44);
                zpar ^= 1;
                tc_fence_after();
                GNN_STAMP(layer * 12 + 4);
                store_p(alpha2);                        // round 2: heads 2 and 3 (the MMAs that read round 1 are done)
                convert2(half * 64, half * 32);
                tmem_st_wait();
                tc_fence_before();
                fence_proxy_async_smem();
                __syncthreads();
                GNN_STAMP(layer * 12 + 5);
                if (tid == 0) {
                    tc_fence_after();
#pragma unroll
                    for (int b = 0; b < 2; ++b) {       // OUT = Z^0 W_0^T + Z^1 W_1^T
                        const uint64_t wd = wait_item(item + b);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_bf16_ts(tmem_base + kColOut4, tmem_base + kColZb + b * 32 + k * 8, desc_add(wd, k * 32), idesc_h, (b | k) != 0);
                    }
#pragma unroll
                    for (int b = 0; b < 2; ++b)
#pragma unroll
                        for (int kk = 0; kk < 8; ++kk)
                            umma_bf16(tmem_base + kColZf + b * 64, desc_add(p_desc, b * 32768 + (kk >> 2) * 16384 + (kk & 3) * 32),
                                      desc_add(x_desc, kk * 2048), idesc_agg, kk != 0);
                    umma_commit(z_bar);
                }
                wait(z_bar, zpar, 45);
                zpar ^= 1;
                tc_fence_after();
                GNN_STAMP(layer * 12 + 6);
                if (tid == 0) { load_item(item + 3); load_item(item + 4); }     // slices 0, 1 are consumed
                if (layer < 4 && half == 0) {              // P buffer 0 is drained: the adjacency (no self loops) of the
                    const __nv_bfloat16 one = __float2bfloat16_rn(1.f), zero = __float2bfloat16_rn(0.f);    // GraphConv layer that follows
                    *reinterpret_cast<__nv_bfloat16*>(s_p + pofs[0]) = zero;
#pragma unroll
                    for (int k = 1; k <= kMaxDeg; ++k)
                        if (k <= dg) *reinterpret_cast<__nv_bfloat16*>(s_p + pofs[k]) = one;
                }
                convert2(half * 64, half * 32);
                tmem_st_wait();
                tc_fence_before();
                fence_proxy_async_smem();
                __syncthreads();
                GNN_STAMP(layer * 12 + 7);
                if (tid == 0) {
                    tc_fence_after();
#pragma unroll
                    for (int b = 0; b < 2; ++b) {       // OUT += Z^2 W_2^T + Z^3 W_3^T
                        const uint64_t wd = wait_item(item + 2 + b);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_bf16_ts(tmem_base + kColOut4, tmem_base + kColZb + b * 32 + k * 8, desc_add(wd, k * 32), idesc_h, 1);
                    }
                    umma_commit(o_bar);
                }
                wait(o_bar, opar, 46);
                opar ^= 1;
                tc_fence_after();
                GNN_STAMP(layer * 12 + 8);
                if (tid == 0) { load_item(item + 5); load_item(item + 6); item += 4; }
            } else {
                // ================= GraphConv =================
                // P buffer 0 already holds the adjacency: the preceding GAT layer wrote it once its own aggregation was done
                if (tid == 0) {
                    tc_fence_after();
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk)       // AGG = Adj . X (exact: 0/1 weights, fp32 accumulation)
                        umma_bf16(tmem_base + kColZf, desc_add(p_desc, (kk >> 2) * 16384 + (kk & 3) * 32),
                                  desc_add(x_desc, kk * 2048), idesc_agg, kk != 0);
                    umma_commit(z_bar);
                }
                wait(z_bar, zpar, 47);
                zpar ^= 1;
                tc_fence_after();
                GNN_STAMP(layer * 12 + 4);
                convert(half * 32, half * 16);
                tmem_st_wait();
                tc_fence_before();
                __syncthreads();
                GNN_STAMP(layer * 12 + 7);
                if (tid == 0) {
                    tc_fence_after();
                    const uint64_t wrel = wait_item(item), wroot = wait_item(item + 1);
#pragma unroll
                    for (int k = 0; k < 4; ++k)           // W_rel . agg
                        umma_bf16_ts(tmem_base + kColOut4, tmem_base + kColZb + k * 8, desc_add(wrel, k * 32), idesc_h, k != 0);
#pragma unroll
                    for (int k = 0; k < 4; ++k)           // + W_root . x
                        umma_bf16(tmem_base + kColOut4, desc_add(x_desc, k * 32), desc_add(wroot, k * 32), idesc_h, 1);
                    umma_commit(o_bar);
                }
                wait(o_bar, opar, 48);
                opar ^= 1;
                tc_fence_after();
                GNN_STAMP(layer * 12 + 8);
                if (tid == 0) { load_item(item + 3); load_item(item + 4); item += 2; }
            }
            {
                uint32_t t[32];
                tmem_ld_32x32(tmem_lane + kColOut4 + half * 32, t);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = a2m_fft::pack(__uint_as_float(t[2 * i]), __uint_as_float(t[2 * i + 1]));
            }
            // ---- + bias, LayerNorm(64) over the two 32-feature halves of the node -> LeakyReLU -> + residual; every step on
            // packed pairs (FADD2 / FMUL2 / FFMA2: half the issue slots of this phase, the longest of the layer)
            {
                using namespace a2m_fft;
                const float4* par = reinterpret_cast<const float4*>(s_par + half * 32);
#pragma unroll
                for (int i4 = 0; i4 < 8; ++i4) {
                    const float4 b4 = par[i4];
                    v[2 * i4] = add2(v[2 * i4], pack(b4.x, b4.y));
                    v[2 * i4 + 1] = add2(v[2 * i4 + 1], pack(b4.z, b4.w));
                }
                pair_t s2 = pack(0.f, 0.f), q2 = pack(0.f, 0.f);
#pragma unroll
                for (int i = 0; i < 16; ++i) { s2 = add2(s2, v[i]); q2 = fma2(v[i], v[i], q2); }
                *reinterpret_cast<float2*>(s_ln + (r * 2 + half) * 2) = make_float2(lo(s2) + hi(s2), lo(q2) + hi(q2));
                switch (quad) {                            // only the two warps that share these 32 rows (warp, warp + 4);
                    case 0: named_barrier(1, 64); break;   // literal ids keep the kernel at five hardware barriers
                    case 1: named_barrier(2, 64); break;
                    case 2: named_barrier(3, 64); break;
                    default: named_barrier(4, 64); break;
                }
                const float4 a = *reinterpret_cast<const float4*>(s_ln + r * 4);
                const float mean = (a.x + a.z) * (1.f / 64.f);
                const float rstd = rsqrtf(fmaxf((a.y + a.w) * (1.f / 64.f) - mean * mean, 0.f) + 1e-5f);
                const pair_t nmean = bcast(-mean), rs = bcast(rstd), slope = bcast(kLeakySlope);
#pragma unroll
                for (int i4 = 0; i4 < 8; ++i4) {
                    const float4 w4 = par[16 + i4], b4 = par[32 + i4];
#pragma unroll
                    for (int h2 = 0; h2 < 2; ++h2) {
                        const pair_t w = h2 ? pack(w4.z, w4.w) : pack(w4.x, w4.y), b = h2 ? pack(b4.z, b4.w) : pack(b4.x, b4.y);
                        const pair_t y = fma2(add2(v[2 * i4 + h2], nmean), mul2(w, rs), b);
                        const pair_t z = mul2(y, slope);                           // LeakyReLU(0.2) = max(y, 0.2 y)
                        x[2 * i4 + h2] = add2(x[2 * i4 + h2], pack(fmaxf(lo(y), lo(z)), fmaxf(hi(y), hi(z))));
                    }
                }
            }
            if (!live) {
#pragma unroll
                for (int i = 0; i < 16; ++i) x[i] = a2m_fft::pack(0.f, 0.f);
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint4 o;
                o.x = pack_bf16(a2m_fft::lo(x[c * 4]), a2m_fft::hi(x[c * 4])); o.y = pack_bf16(a2m_fft::lo(x[c * 4 + 1]), a2m_fft::hi(x[c * 4 + 1]));
                o.z = pack_bf16(a2m_fft::lo(x[c * 4 + 2]), a2m_fft::hi(x[c * 4 + 2])); o.w = pack_bf16(a2m_fft::lo(x[c * 4 + 3]), a2m_fft::hi(x[c * 4 + 3]));
                *reinterpret_cast<uint4*>(s_x + sw128_off(r, half * 4 + c)) = o;
            }
            tc_fence_before();
            fence_proxy_async_smem();
            __syncthreads();
            GNN_STAMP(layer * 12 + 9);
        }
        GNN_TILE_STAMP(2);
        if (tid == 0) {                                // store this tile, then (same buffer) fetch the next one
            tma_store_5d(&p.x_out, s_x, 0, row0, group, 0, 0);
            tma_store_commit();
            if (it + 1 < my_tiles) {
                tma_store_wait_read();
                load_tile(tile + gridDim.x);
            }
        }
    }
    if (tid == 0) tma_store_wait_read();
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 256); }
}

}  // namespace v4



// U rows of the extended GAT weight: the attention logits a_src . (W_h x) = (W_h^T a_src) . x come out of the
// same MMA as H.  Rows 256+h: src (hi), 260+h: dst (hi), 264+h: src (lo), 268+h: dst (lo); hi + lo carries
// ~16 mantissa bits of the fp32 fold, which is evaluated on the bf16-rounded W the MMA itself uses.
__global__ void gat_fold_attention_kernel(__nv_bfloat16* __restrict__ wext, const float* __restrict__ att_src,
                                          const float* __restrict__ att_dst) {
    const int f = threadIdx.x & 63, h = (threadIdx.x >> 6) & 3, which = threadIdx.x >> 8;       // 512 threads
    const float* a = which == 0 ? att_src : att_dst;
    float u = 0.f;
    for (int o = 0; o < kJointFeat; ++o)
        u = fmaf(a[h * kJointFeat + o], __bfloat162float(wext[(h * kJointFeat + o) * kJointFeat + f]), u);
    const __nv_bfloat16 hi = __float2bfloat16_rn(u);
    const __nv_bfloat16 lo = __float2bfloat16_rn(u - __bfloat162float(hi));
    wext[(256 + which * 4 + h) * kJointFeat + f] = hi;
    wext[(264 + which * 4 + h) * kJointFeat + f] = lo;
}

}  // namespace

// Host side ---------------------------------------------------------------------------------------
struct GnnFusedPlan {
    GnnParams p;
    int grid;
};

int make_weight_map(CUtensorMap* map, const void* w, long long n_rows, long long k, int box_rows);   // conv_gemm.cu

int gat_fold_attention(__nv_bfloat16* wext, const float* att_src, const float* att_dst, cudaStream_t stream) {
    gat_fold_attention_kernel<<<1, 512, 0, stream>>>(wext, att_src, att_dst);
    a2m_count_launch();
    A2M_LAUNCH_CHECK();
    return A2M_OK;
}

int make_map_bf16(CUtensorMap* map, const void* ptr, int rank, const long long* dims, const long long* strides,
                  const int* box, const char* what);                                                   // conv_gemm.cu

int gnn_fused_plan(const GnnFusedWeights& w, GraphTopo topo, long long n_groups, int group_graphs, const __nv_bfloat16* x_in,
                   __nv_bfloat16* x_out, std::shared_ptr<GnnFusedPlan>* out) {
    A2M_ARG_CHECK(topo.n_nodes >= 1 && topo.n_nodes <= 48, "gnn: %d nodes per graph (max 48)", topo.n_nodes);
    A2M_ARG_CHECK(n_groups >= 1 && group_graphs >= 1 && n_groups * group_graphs * topo.n_nodes <= 0x7fffffffLL &&
                  static_cast<long long>(group_graphs) * topo.n_nodes <= 0x7fffffffLL && n_groups <= 0x7fffffffLL,
                  "gnn: %lld groups of %d graphs", n_groups, group_graphs);
    auto plan = std::make_shared<GnnFusedPlan>();
    GnnParams& p = plan->p;
    memset(&p, 0, sizeof(p));
    int rc;
    for (int i = 0; i < 3; ++i) {
        rc = make_weight_map(&p.w_head[i], w.gat_w[i], kGatRows, 64, 64);
        if (rc != A2M_OK) return rc;
        rc = make_weight_map(&p.w_att[i], w.gat_w[i], kGatRows, 64, 16);
        if (rc != A2M_OK) return rc;
        p.gat_bias[i] = w.gat_bias[i];
    }
    for (int i = 0; i < 2; ++i) {
        rc = make_weight_map(&p.w_gc[i], w.gc_w[i], 64, 128, 64);
        if (rc != A2M_OK) return rc;
        p.gc_bias[i] = w.gc_bias[i];
    }
    for (int i = 0; i < 5; ++i) { p.ln_w[i] = w.ln_w[i]; p.ln_b[i] = w.ln_b[i]; }
    p.nbr = topo.nbr; p.deg = topo.deg; p.J = topo.n_nodes; p.gpc = kRows / topo.n_nodes;
    p.n_groups = n_groups; p.group_graphs = group_graphs;
    p.tiles_per_group = (group_graphs + p.gpc - 1) / p.gpc;
    const int rows_per_tile = p.gpc * p.J;
    const long long group_rows = static_cast<long long>(group_graphs) * p.J;
    const long long dims[3] = {64, group_rows, n_groups}, strides[3] = {1, 64, group_rows * 64};
    const int box[5] = {64, rows_per_tile, 1, 1, 1};
    rc = make_map_bf16(&p.x_in, x_in, 3, dims, strides, box, "gnn node features (in)");
    if (rc != A2M_OK) return rc;
    rc = make_map_bf16(&p.x_out, x_out, 3, dims, strides, box, "gnn node features (out)");
    if (rc != A2M_OK) return rc;
    const long long tiles = n_groups * p.tiles_per_group;
    const long long resident = 2LL * a2m_num_sms();                        // two CTAs per SM
    plan->grid = static_cast<int>(tiles < resident ? tiles : resident);
    *out = plan;
    return A2M_OK;
}

#ifdef A2M_GNN_TRACE
extern "C" int a2m_gnn_trace_read(long long* out) {
    return cudaMemcpyFromSymbol(out, v4::g_gnn_trace, sizeof(long long) * 128) == cudaSuccess ? 0 : -1;
}
#endif

int gnn_fused_launch(const GnnFusedPlan& plan, int* err_flag, cudaStream_t stream) {
    static A2mPerDeviceOnce configured;
    if (configured.first())
        A2M_CUDA_CHECK(cudaFuncSetAttribute(v4::gnn4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, v4::kSmemBytes4));
    A2M_CUDA_CHECK(a2m_launch_pdl(v4::gnn4_kernel, dim3(plan.grid), dim3(v4::kThreads4), v4::kSmemBytes4, stream, plan.p, err_flag));
    a2m_count_launch();
    return A2M_OK;
}

}  // namespace a2m
