// Fused static-skeleton GNN stack: the five layers of real_motion_model.py:172-201 /
// :224-253  --  GAT, GraphConv, GAT, GraphConv, GAT, each followed by LayerNorm(64) -> LeakyReLU(0.2)
// -> + residual  --  in ONE kernel, node features never leaving the SM between layers.
//
// torch_geometric semantics restated (PyG is not a dependency):
//   GATConv(64, 64, heads=4, concat=False): h = W x (no bias), e_ij = LeakyReLU_0.2(a_src.h_j + a_dst.h_i) over
//     j in N(i) + {i}, alpha = softmax_j(e_ij), out_i = mean_heads(sum_j alpha_ij h_j) + bias
//   GraphConv(64, 64), aggr = add:        out_i = W_rel (sum_{j in N(i)} x_j) + b_rel + W_root x_i
//
// One persistent CTA per SM keeps all five weight matrices resident in shared memory (128 KB, loaded
// once with TMA in the 128B-swizzled K-major layout) and walks over tiles of whole graphs (<= 128 node
// rows: 3 hand graphs or 12 body graphs).  Per layer the shared linear runs on the tensor cores
// (tcgen05.mma, A = the bf16 node tile in shared memory, D = [128 x 256] fp32 in TMEM); the epilogue
// threads (2 per node) read their TMEM lane, reduce the attention scalars, stage h as bf16 in shared
// memory, and do the neighbour softmax / aggregation / head mean / LayerNorm / residual with the
// residual stream held in fp32 registers across all five layers.
#include <cuda.h>
#include <cstring>
#include "conv_gemm.cuh"
#include "layers.cuh"

void a2m_count_launch();

namespace a2m {

namespace {

constexpr int kThreads = 256;
constexpr int kRows = 128;
constexpr int kHStride = 528;                         // bytes per staged h row (512 + 16: conflict-free 16 B accesses)
constexpr int kOffW = 0;                              // 3 x 32 KB GAT + 2 x 16 KB GraphConv
constexpr int kOffX = 131072;                         // node tile, bf16 [128][64] SW128
constexpr int kOffH = kOffX + 16384;                  // h rows (GAT) / aggregated tile (GraphConv)
constexpr int kOffS = kOffH + kRows * kHStride;       // s_src, s_dst [128][4] fp32 each
constexpr int kOffLn = kOffS + 2 * kRows * 4 * 4;     // LayerNorm partials [128][2][2] fp32
constexpr int kOffTopo = kOffLn + kRows * 2 * 2 * 4;  // nbr [J][6], deg [J]
constexpr int kOffBar = kOffTopo + 48 * kMaxDeg * 4 + 48 * 4;
constexpr int kSmemBytes = kOffBar + 64 + 1024;

struct GnnParams {
    CUtensorMap w_gat[3];
    CUtensorMap w_gc[2];
    const float* att_src[3];
    const float* att_dst[3];
    const float* gat_bias[3];
    const float* gc_bias[2];
    const float* ln_w[5];
    const float* ln_b[5];
    const int* nbr;
    const int* deg;
    int J, gpc;
    long long n_graphs;
    const __nv_bfloat16* x_in;
    __nv_bfloat16* x_out;
};

__device__ __forceinline__ float leaky(float x) { return x > 0.f ? x : kLeakySlope * x; }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ float bf_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
// byte offset of 16-byte chunk `c` (8 features) of row r in a [128][64] bf16 SW128 K-major tile
__device__ __forceinline__ int sw128_off(int r, int c) { return r * 128 + ((c ^ (r & 7)) << 4); }

// LayerNorm(64) over the two 32-feature halves of a node held by threads (r, 0) and (r, 1)
__device__ __forceinline__ void layernorm_pair(float (&v)[32], float* s_ln, int r, int half, const float* __restrict__ w,
                                               const float* __restrict__ b) {
    float s = 0.f, q = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) { s += v[i]; q = fmaf(v[i], v[i], q); }
    s_ln[(r * 2 + half) * 2] = s;
    s_ln[(r * 2 + half) * 2 + 1] = q;
    __syncthreads();
    const float ts = s + s_ln[(r * 2 + (half ^ 1)) * 2], tq = q + s_ln[(r * 2 + (half ^ 1)) * 2 + 1];
    const float mean = ts * (1.f / 64.f);
    const float rstd = rsqrtf(fmaxf(tq * (1.f / 64.f) - mean * mean, 0.f) + 1e-5f);
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = (v[i] - mean) * rstd * __ldg(w + half * 32 + i) + __ldg(b + half * 32 + i);
}

__global__ void __launch_bounds__(kThreads, 1)
gnn_fused_kernel(const __grid_constant__ GnnParams p, int* __restrict__ err_flag) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    unsigned char* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    unsigned char* s_x = smem + kOffX;
    unsigned char* s_h = smem + kOffH;
    float* s_src = reinterpret_cast<float*>(smem + kOffS);
    float* s_dst = s_src + kRows * 4;
    float* s_ln = reinterpret_cast<float*>(smem + kOffLn);
    int* s_nbr = reinterpret_cast<int*>(smem + kOffTopo);
    int* s_deg = s_nbr + 48 * kMaxDeg;
    uint64_t* w_bar = reinterpret_cast<uint64_t*>(smem + kOffBar);
    uint64_t* mma_bar = w_bar + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mma_bar + 1);

    const int tid = threadIdx.x, warp = tid >> 5;
    const int r = tid & 127, half = tid >> 7, quad = warp & 3;
    const int J = p.J, rows_per_tile = p.gpc * J;

    if (tid == 0) {
        for (int i = 0; i < 3; ++i) tma_prefetch_desc(&p.w_gat[i]);
        for (int i = 0; i < 2; ++i) tma_prefetch_desc(&p.w_gc[i]);
        mbar_init(w_bar, 1);
        mbar_init(mma_bar, 1);
        mbar_fence_init();
        // all five weight matrices, once per CTA
        mbar_expect_tx(w_bar, 131072);
        for (int i = 0; i < 3; ++i) tma_load_5d(smem + kOffW + i * 32768, &p.w_gat[i], w_bar, 0, 0, 0, 0, 0);
        for (int i = 0; i < 2; ++i) {
            tma_load_5d(smem + kOffW + 98304 + i * 16384, &p.w_gc[i], w_bar, 0, 0, 0, 0, 0);          // W_rel  (k 0..63)
            tma_load_5d(smem + kOffW + 98304 + i * 16384 + 8192, &p.w_gc[i], w_bar, 64, 0, 0, 0, 0);  // W_root (k 64..127)
        }
    }
    if (warp == 1) { tmem_alloc(tmem_slot, 256); tmem_relinquish(); }
    for (int i = tid; i < J * kMaxDeg; i += kThreads) s_nbr[i] = p.nbr[i];
    for (int i = tid; i < J; i += kThreads) s_deg[i] = p.deg[i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    mbar_wait(w_bar, 0, err_flag, 11);      // bounded; on expiry the flag is set and the CTA still runs to completion
    uint32_t mma_parity = 0;
    const uint32_t idesc_gat = umma_idesc_bf16(128, 256), idesc_gc = umma_idesc_bf16(128, 64);
    const uint32_t x_addr = smem_u32(s_x), h_addr = smem_u32(s_h), w_addr = smem_u32(smem + kOffW);

    const long long n_tiles = (p.n_graphs + p.gpc - 1) / p.gpc;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long row0 = tile * rows_per_tile;
        const long long rows_left = p.n_graphs * J - row0;
        const int n_here = static_cast<int>(rows_left < rows_per_tile ? rows_left : rows_per_tile);
        const bool live = r < n_here;
        const int jloc = r % J, g0 = r - jloc;
        const int dg = live ? s_deg[jloc] : 0;

        // ---- load this thread's 32 features (fp32 residual stream in registers + bf16 MMA operand tile)
        float x[32];
        {
            uint4 q[4] = {};
            if (live) {
                const uint4* src = reinterpret_cast<const uint4*>(p.x_in + (row0 + r) * 64 + half * 32);
#pragma unroll
                for (int c = 0; c < 4; ++c) q[c] = __ldg(src + c);
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                *reinterpret_cast<uint4*>(s_x + sw128_off(r, half * 4 + c)) = q[c];
                const uint32_t u[4] = {q[c].x, q[c].y, q[c].z, q[c].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) { x[c * 8 + 2 * e] = bf_lo(u[e]); x[c * 8 + 2 * e + 1] = bf_hi(u[e]); }
            }
        }
        fence_async_smem();
        __syncthreads();

#pragma unroll 1
        for (int layer = 0; layer < 5; ++layer) {
            float v[32];
            if ((layer & 1) == 0) {
                // ================= GATConv =================
                const int gi = layer >> 1;
                if (tid == 0) {
                    tc_fence_after();
                    const uint32_t wa = w_addr + gi * 32768;
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16(tmem_base, umma_desc_sw128(x_addr + k * 32), umma_desc_sw128(wa + k * 32), idesc_gat, k != 0);
                    umma_commit(mma_bar);
                }
                mbar_wait(mma_bar, mma_parity, err_flag, 12);
                mma_parity ^= 1;
                tc_fence_after();
                // phase A: my node's h for heads {2*half, 2*half+1}: attention scalars + bf16 staging
                const float* a_src = p.att_src[gi];
                const float* a_dst = p.att_dst[gi];
#pragma unroll 1
                for (int hh = 0; hh < 2; ++hh) {
                    float ps = 0.f, pd = 0.f;
#pragma unroll
                    for (int cc = 0; cc < 2; ++cc) {
                        const int col0 = half * 128 + hh * 64 + cc * 32;
                        uint32_t t[32];
                        tmem_ld_32x32(tmem_lane + col0, t);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const float f = __uint_as_float(t[i]);
                            ps = fmaf(f, __ldg(a_src + col0 + i), ps);
                            pd = fmaf(f, __ldg(a_dst + col0 + i), pd);
                        }
                        unsigned char* hrow = s_h + r * kHStride + col0 * 2;
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            uint4 q;
                            q.x = pack_bf16(__uint_as_float(t[c * 8]), __uint_as_float(t[c * 8 + 1]));
                            q.y = pack_bf16(__uint_as_float(t[c * 8 + 2]), __uint_as_float(t[c * 8 + 3]));
                            q.z = pack_bf16(__uint_as_float(t[c * 8 + 4]), __uint_as_float(t[c * 8 + 5]));
                            q.w = pack_bf16(__uint_as_float(t[c * 8 + 6]), __uint_as_float(t[c * 8 + 7]));
                            *reinterpret_cast<uint4*>(hrow + c * 16) = q;
                        }
                    }
                    s_src[r * 4 + half * 2 + hh] = ps;
                    s_dst[r * 4 + half * 2 + hh] = pd;
                }
                tc_fence_before();
                __syncthreads();
                // phase B: softmax over {self} + neighbours per head, weighted aggregation of my 32 features
                float alpha[kMaxDeg + 1][4];
                int idx[kMaxDeg + 1];
                idx[0] = r;
#pragma unroll
                for (int k = 0; k < kMaxDeg; ++k) idx[k + 1] = k < dg ? g0 + s_nbr[jloc * kMaxDeg + k] : r;
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    const float di = s_dst[r * 4 + h];
                    float m = -INFINITY;
#pragma unroll
                    for (int k = 0; k <= kMaxDeg; ++k) {
                        alpha[k][h] = k <= dg ? leaky(s_src[idx[k] * 4 + h] + di) : -INFINITY;
                        m = fmaxf(m, alpha[k][h]);
                    }
                    float den = 0.f;
#pragma unroll
                    for (int k = 0; k <= kMaxDeg; ++k) { alpha[k][h] = k <= dg ? __expf(alpha[k][h] - m) : 0.f; den += alpha[k][h]; }
                    const float inv = 0.25f / den;                      // softmax normaliser and the head mean
#pragma unroll
                    for (int k = 0; k <= kMaxDeg; ++k) alpha[k][h] *= inv;
                }
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = 0.f;
#pragma unroll
                for (int k = 0; k <= kMaxDeg; ++k) {
                    if (k <= dg) {
                        const unsigned char* hrow = s_h + idx[k] * kHStride + half * 64;
#pragma unroll
                        for (int h = 0; h < 4; ++h) {
                            const float a = alpha[k][h];
#pragma unroll
                            for (int c = 0; c < 4; ++c) {
                                const uint4 q = *reinterpret_cast<const uint4*>(hrow + h * 128 + c * 16);
                                v[c * 8 + 0] = fmaf(a, bf_lo(q.x), v[c * 8 + 0]); v[c * 8 + 1] = fmaf(a, bf_hi(q.x), v[c * 8 + 1]);
                                v[c * 8 + 2] = fmaf(a, bf_lo(q.y), v[c * 8 + 2]); v[c * 8 + 3] = fmaf(a, bf_hi(q.y), v[c * 8 + 3]);
                                v[c * 8 + 4] = fmaf(a, bf_lo(q.z), v[c * 8 + 4]); v[c * 8 + 5] = fmaf(a, bf_hi(q.z), v[c * 8 + 5]);
                                v[c * 8 + 6] = fmaf(a, bf_lo(q.w), v[c * 8 + 6]); v[c * 8 + 7] = fmaf(a, bf_hi(q.w), v[c * 8 + 7]);
                            }
                        }
                    }
                }
                const float* gb = p.gat_bias[gi];
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] += __ldg(gb + half * 32 + i);
            } else {
                // ================= GraphConv =================
                const int ci = layer >> 1;
                // aggregated neighbour tile (bf16, SW128) into the h buffer
                float agg[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) agg[i] = 0.f;
#pragma unroll
                for (int k = 0; k < kMaxDeg; ++k) {
                    if (k < dg) {
                        const int j = g0 + s_nbr[jloc * kMaxDeg + k];
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const uint4 q = *reinterpret_cast<const uint4*>(s_x + sw128_off(j, half * 4 + c));
                            agg[c * 8 + 0] += bf_lo(q.x); agg[c * 8 + 1] += bf_hi(q.x); agg[c * 8 + 2] += bf_lo(q.y); agg[c * 8 + 3] += bf_hi(q.y);
                            agg[c * 8 + 4] += bf_lo(q.z); agg[c * 8 + 5] += bf_hi(q.z); agg[c * 8 + 6] += bf_lo(q.w); agg[c * 8 + 7] += bf_hi(q.w);
                        }
                    }
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint4 q;
                    q.x = pack_bf16(agg[c * 8], agg[c * 8 + 1]); q.y = pack_bf16(agg[c * 8 + 2], agg[c * 8 + 3]);
                    q.z = pack_bf16(agg[c * 8 + 4], agg[c * 8 + 5]); q.w = pack_bf16(agg[c * 8 + 6], agg[c * 8 + 7]);
                    *reinterpret_cast<uint4*>(s_h + sw128_off(r, half * 4 + c)) = q;
                }
                fence_async_smem();
                __syncthreads();
                if (tid == 0) {
                    tc_fence_after();
                    const uint32_t wa = w_addr + 98304 + ci * 16384;
#pragma unroll
                    for (int k = 0; k < 4; ++k)       // W_rel . agg
                        umma_bf16(tmem_base, umma_desc_sw128(h_addr + k * 32), umma_desc_sw128(wa + k * 32), idesc_gc, k != 0);
#pragma unroll
                    for (int k = 0; k < 4; ++k)       // + W_root . x
                        umma_bf16(tmem_base, umma_desc_sw128(x_addr + k * 32), umma_desc_sw128(wa + 8192 + k * 32), idesc_gc, 1);
                    umma_commit(mma_bar);
                }
                mbar_wait(mma_bar, mma_parity, err_flag, 13);
                mma_parity ^= 1;
                tc_fence_after();
                uint32_t t[32];
                tmem_ld_32x32(tmem_lane + half * 32, t);
                tmem_ld_wait();
                tc_fence_before();
                const float* cb = p.gc_bias[ci];
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(t[i]) + __ldg(cb + half * 32 + i);
            }
            // ---- LayerNorm(64) -> LeakyReLU -> + residual; refresh the bf16 operand tile
            layernorm_pair(v, s_ln, r, half, p.ln_w[layer], p.ln_b[layer]);
#pragma unroll
            for (int i = 0; i < 32; ++i) x[i] += leaky(v[i]);
            if (!live) {
#pragma unroll
                for (int i = 0; i < 32; ++i) x[i] = 0.f;
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint4 q;
                q.x = pack_bf16(x[c * 8], x[c * 8 + 1]); q.y = pack_bf16(x[c * 8 + 2], x[c * 8 + 3]);
                q.z = pack_bf16(x[c * 8 + 4], x[c * 8 + 5]); q.w = pack_bf16(x[c * 8 + 6], x[c * 8 + 7]);
                *reinterpret_cast<uint4*>(s_x + sw128_off(r, half * 4 + c)) = q;
                if (layer == 4 && live) *reinterpret_cast<uint4*>(p.x_out + (row0 + r) * 64 + half * 32 + c * 8) = q;
            }
            fence_async_smem();
            __syncthreads();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 256); }
}

}  // namespace

// Host side ---------------------------------------------------------------------------------------
struct GnnFusedPlan {
    GnnParams p;
    int grid;
};

int make_weight_map(CUtensorMap* map, const void* w, long long n_rows, long long k, int box_rows);   // conv_gemm.cu

int gnn_fused_plan(const GnnFusedWeights& w, GraphTopo topo, long long n_graphs, const __nv_bfloat16* x_in,
                   __nv_bfloat16* x_out, std::shared_ptr<GnnFusedPlan>* out) {
    A2M_ARG_CHECK(topo.n_nodes >= 1 && topo.n_nodes <= 48, "gnn: %d nodes per graph (max 48)", topo.n_nodes);
    auto plan = std::make_shared<GnnFusedPlan>();
    GnnParams& p = plan->p;
    memset(&p, 0, sizeof(p));
    for (int i = 0; i < 3; ++i) {
        const int rc = make_weight_map(&p.w_gat[i], w.gat_w[i], 256, 64, 256);
        if (rc != A2M_OK) return rc;
        p.att_src[i] = w.att_src[i]; p.att_dst[i] = w.att_dst[i]; p.gat_bias[i] = w.gat_bias[i];
    }
    for (int i = 0; i < 2; ++i) {
        const int rc = make_weight_map(&p.w_gc[i], w.gc_w[i], 64, 128, 64);
        if (rc != A2M_OK) return rc;
        p.gc_bias[i] = w.gc_bias[i];
    }
    for (int i = 0; i < 5; ++i) { p.ln_w[i] = w.ln_w[i]; p.ln_b[i] = w.ln_b[i]; }
    p.nbr = topo.nbr; p.deg = topo.deg; p.J = topo.n_nodes; p.gpc = kRows / topo.n_nodes;
    p.n_graphs = n_graphs; p.x_in = x_in; p.x_out = x_out;
    const long long tiles = (n_graphs + p.gpc - 1) / p.gpc;
    const int sms = a2m_num_sms();
    plan->grid = static_cast<int>(tiles < sms ? tiles : sms);
    *out = plan;
    return A2M_OK;
}

int gnn_fused_launch(const GnnFusedPlan& plan, int* err_flag, cudaStream_t stream) {
    static bool configured = false;
    if (!configured) {
        A2M_CUDA_CHECK(cudaFuncSetAttribute(gnn_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        configured = true;
    }
    gnn_fused_kernel<<<plan.grid, kThreads, kSmemBytes, stream>>>(plan.p, err_flag);
    a2m_count_launch();
    A2M_LAUNCH_CHECK();
    return A2M_OK;
}

}  // namespace a2m
