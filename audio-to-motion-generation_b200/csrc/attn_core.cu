// SelfAttention core for the wide UNet attentions (model_layers.py:133-146 with C = 2048 / 1024, D1 placement),
// after their q | k | v projection GEMM:  out = gamma * softmax(q k^T) v + x.
//   S = q k^T                         -> tcgen05 (K = d = C / 8), all clips of a 128-row tile at once
//   P = exp(S - rowmax) on own clip   -> CUDA cores, one thread per row, bf16 into a block-diagonal [128 x 128] tile
//   O = P v  for one 256-channel slab -> tcgen05 (B = v slab read MN-major straight from its TMA tile)
//   out = gamma * O / rowsum + x      -> bf16
// grid = (128-row tiles, C / 256 channel slabs): every CTA recomputes the cheap S / P of its tile (16 MMAs) and owns
// one slab of v, so the two launches fill the machine (256 CTAs) instead of 32 / 64 tiles.
// Replaces the CUDA-core attention_kernel for these shapes (29 + 34 us -> see DESIGN.md).
#include <cuda.h>
#include <cstring>
#include "conv_gemm.cuh"
#include "layers.cuh"

void a2m_count_launch();

namespace a2m {

int make_weight_map(CUtensorMap* map, const void* w, long long n_rows, long long k, int box_rows);   // conv_gemm.cu

namespace {

constexpr int kThreads = 512;
constexpr int kOffQ = 0;                    // q tile: d / 64 chunks of [128][64] bf16 (<= 64 KB)
constexpr int kOffK = 65536;                // k tile
constexpr int kOffP = 131072;               // P [128][128] bf16, two K-major chunks
constexpr int kOffV = 163840;               // v slab: 4 chunks of [128 (time)][64 (channels)]
constexpr int kOffSum = kOffV + 65536;      // row sums [128] fp32
constexpr int kOffBar = kOffSum + 512;
constexpr int kSmemBytes = kOffBar + 64 + 1024;
constexpr uint32_t kColS = 0, kColO = 128;  // TMEM: S [0,128), O [128,384)
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");

struct CoreParams {
    CUtensorMap qkv_map;      // [rows][2 d + C] bf16, box 64 x 128
    const float* gamma;
    const __nv_bfloat16* x;
    const __nv_bfloat16* res2;
    __nv_bfloat16* out;
    long long n_rows;
    int T, C, d;
};

__device__ __forceinline__ uint32_t pack2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ int sw128_off(int r, int c) { return r * 128 + ((c ^ (r & 7)) << 4); }
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
attn_core_kernel(const __grid_constant__ CoreParams p, int* __restrict__ err_flag) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    unsigned char* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    float* s_sum = reinterpret_cast<float*>(smem + kOffSum);
    uint64_t* qk_bar = reinterpret_cast<uint64_t*>(smem + kOffBar);
    uint64_t* v_bar = qk_bar + 1;
    uint64_t* s_bar = qk_bar + 2;
    uint64_t* o_bar = qk_bar + 3;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(qk_bar + 4);

    const int tid = threadIdx.x, warp = tid >> 5;
    const int r = tid & 127, q = tid >> 7, quad = warp & 3;
    const long long row0 = static_cast<long long>(blockIdx.x) * 128;
    const int slab = blockIdx.y;                          // 256 channels of v / out
    const int T = p.T, d = p.d, n_dc = p.d >> 6;          // d / 64 chunks of q and of k

    pdl_launch_dependents();
    if (tid == 0) {
        tma_prefetch_desc(&p.qkv_map);
        mbar_init(qk_bar, 1);
        mbar_init(v_bar, 1);
        mbar_init(s_bar, 1);
        mbar_init(o_bar, 1);
        mbar_fence_init();
    }
    if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
    {   // zero P: only the row's own clip block is ever written
        uint4* pz = reinterpret_cast<uint4*>(smem + kOffP);
        for (int i = tid; i < 32768 / 16; i += kThreads) pz[i] = make_uint4(0, 0, 0, 0);
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const uint32_t base_addr = smem_u32(smem);
    pdl_wait();                                           // q | k | v come from the projection GEMM just before
    if (tid == 0) {
        mbar_expect_tx(qk_bar, static_cast<uint32_t>(2 * n_dc * 16384));
        for (int i = 0; i < n_dc; ++i) {
            tma_load_5d(smem + kOffQ + i * 16384, &p.qkv_map, qk_bar, i * 64, static_cast<int>(row0), 0, 0, 0);
            tma_load_5d(smem + kOffK + i * 16384, &p.qkv_map, qk_bar, d + i * 64, static_cast<int>(row0), 0, 0, 0);
        }
        mbar_expect_tx(v_bar, 65536);
        for (int i = 0; i < 4; ++i)
            tma_load_5d(smem + kOffV + i * 16384, &p.qkv_map, v_bar, 2 * d + slab * 256 + i * 64, static_cast<int>(row0), 0, 0, 0);
        // S = q k^T
        mbar_wait(qk_bar, 0, err_flag, 51);
        tc_fence_after();
        const uint32_t id_s = umma_idesc_bf16(128, 128);
        for (int i = 0; i < n_dc; ++i)
#pragma unroll
            for (int k = 0; k < 4; ++k)
                umma_bf16(tmem_base + kColS, umma_desc_sw128(base_addr + kOffQ + i * 16384 + k * 32),
                          umma_desc_sw128(base_addr + kOffK + i * 16384 + k * 32), id_s, (i | k) != 0);
        umma_commit(s_bar);
    }
    // ---------------- softmax of each row over its own clip (quarter 0: one thread per row) ----------------
    if (q == 0) {
        mbar_wait(s_bar, 0, err_flag, 52);
        tc_fence_after();
        const int win = T < 32 ? 32 : T;                  // warp-uniform column window that covers the warp's clips
        const int win0 = ((quad * 32) / win) * win;
        const int c_lo = (r / T) * T - win0, c_hi = c_lo + T;
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
        for (int c = 0; c < 64; ++c) { e[c] = __expf(e[c] - m); sum += e[c]; }
        s_sum[r] = sum;
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
            const int c = ch * 8;
            if (c >= c_lo && c < c_hi) {
                uint4 o;
                o.x = pack2(e[c], e[c + 1]); o.y = pack2(e[c + 2], e[c + 3]);
                o.z = pack2(e[c + 4], e[c + 5]); o.w = pack2(e[c + 6], e[c + 7]);
                const int col = win0 + c;
                *reinterpret_cast<uint4*>(smem + kOffP + (col >> 6) * 16384 + sw128_off(r, (col & 63) >> 3)) = o;
            }
        }
    }
    // the residual rows of this slab are fetched while the MMAs run: quarters 1-3 get here at once, the softmax
    // threads once their row is done (the 16 vectors must not be live across the softmax registers).  They are read
    // through L2: the kernel is launched programmatically and may have started before the producers of x finished
    const long long row = row0 + r;
    const bool live = row < p.n_rows;
    const long long obase = row * p.C + slab * 256 + q * 64;
    uint4 xv[8], rv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        xv[i] = make_uint4(0, 0, 0, 0);
        rv[i] = make_uint4(0, 0, 0, 0);
        if (live) {
            xv[i] = __ldcg(reinterpret_cast<const uint4*>(p.x + obase + i * 8));
            if (p.res2) rv[i] = __ldcg(reinterpret_cast<const uint4*>(p.res2 + obase + i * 8));
        }
    }
    tc_fence_before();
    fence_proxy_async_smem();
    __syncthreads();
    // ---------------- O = P v (this slab) ----------------
    if (tid == 0) {
        mbar_wait(v_bar, 0, err_flag, 53);
        tc_fence_after();
        const uint32_t id_o = umma_idesc_bf16(128, 256) | (1u << 16);         // B (v) is MN-major
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
            umma_bf16(tmem_base + kColO, umma_desc_sw128(base_addr + kOffP + (kk >> 2) * 16384 + (kk & 3) * 32),
                      umma_desc_mn(base_addr + kOffV + kk * 2048, 16384), id_o, kk != 0);
        umma_commit(o_bar);
    }
    mbar_wait(o_bar, 0, err_flag, 54);
    tc_fence_after();
    {
        const float scale = __ldg(p.gamma) / s_sum[r];
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
            uint32_t t[32];
            tmem_ld_32x32(tmem_lane + kColO + q * 64 + hf * 32, t);
            tmem_ld_wait();
            if (live) {
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
                    *reinterpret_cast<uint4*>(p.out + obase + hf * 32 + c * 8) = make_uint4(os[0], os[1], os[2], os[3]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

}  // namespace

struct AttnCorePlan {
    CoreParams p;
    dim3 grid;
};

bool attn_core_supported(int T, int C) {
    return T >= 8 && T <= 64 && (128 % T) == 0 && C % 256 == 0 && C >= 512 && (C / 8) % 64 == 0 && C / 8 <= 256;
}

int attn_core_plan(const __nv_bfloat16* qkv, const float* gamma, const __nv_bfloat16* x, const __nv_bfloat16* res2, int B, int T,
                   int C, __nv_bfloat16* out, std::shared_ptr<AttnCorePlan>* plan_out) {
    A2M_ARG_CHECK(attn_core_supported(T, C), "attn_core: T = %d, C = %d not supported", T, C);
    auto plan = std::make_shared<AttnCorePlan>();
    CoreParams& p = plan->p;
    memset(&p, 0, sizeof(p));
    const long long rows = static_cast<long long>(B) * T;
    const int d = C / 8;
    const int rc = make_weight_map(&p.qkv_map, qkv, rows, 2 * d + C, 128);
    if (rc != A2M_OK) return rc;
    p.gamma = gamma; p.x = x; p.res2 = res2; p.out = out; p.n_rows = rows; p.T = T; p.C = C; p.d = d;
    plan->grid = dim3(static_cast<unsigned>((rows + 127) / 128), static_cast<unsigned>(C / 256), 1);
    *plan_out = plan;
    return A2M_OK;
}

int attn_core_launch(const AttnCorePlan& plan, int* err_flag, cudaStream_t stream) {
    static A2mPerDeviceOnce configured;
    if (configured.first()) {
        A2M_CUDA_CHECK(cudaFuncSetAttribute(attn_core_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    }
    A2M_CUDA_CHECK(a2m_launch_pdl(attn_core_kernel, plan.grid, dim3(kThreads), kSmemBytes, stream, plan.p, err_flag));
    a2m_count_launch();
    return A2M_OK;
}

}  // namespace a2m
