// Retired generations of the fused graph kernel (v2: one 512-thread CTA per SM; v3: linear first per head, two CTAs
// per SM).  Superseded by the kernel in audio-to-motion-generation_b200/csrc/gnn_fused.cu (aggregate first, A operand
// from tensor memory); kept for reference only -- not compiled into liba2m_b200.so.

constexpr int kThreads = 512;
constexpr int kRows = 128;
constexpr int kGatRows = 272;                          // 256 W rows + 16 folded attention rows
constexpr int kOffW = 0;                               // per-layer weights: GAT 34816 B / GraphConv 16384 B
constexpr int kOffX = 34816;                           // node tile, bf16 [128][64] SW128, x 2 buffers
constexpr int kOffH = kOffX + 2 * 16384;               // GAT: H^h tiles (4 x 16 KB); GraphConv: AGG tile
constexpr int kOffP = kOffH + 4 * 16384;               // attention / adjacency matrices, 2 x [128][128] bf16
constexpr int kOffS = kOffP + 2 * 32768;               // s_src [128][4] fp32
constexpr int kOffLn = kOffS + kRows * 4 * 4;          // LayerNorm partials [128][4][2] fp32
constexpr int kOffTopo = kOffLn + kRows * 4 * 2 * 4;   // nbr [48][6], deg [48]
constexpr int kOffPar = kOffTopo + 48 * kMaxDeg * 4 + 48 * 4;   // per layer: bias[64], ln_w[64], ln_b[64] fp32
constexpr int kOffBar = kOffPar + 5 * 192 * 4;
constexpr int kSmemBytes = kOffBar + 64 + 1024;
constexpr uint32_t kColS = 256, kColOut = 288;         // TMEM columns: H [0,256), S [256,272), OUT [288,352)

static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");

struct GnnParams {
    CUtensorMap w_gat[3];                // [272][64] bf16, box 64 x 136
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

__global__ void __launch_bounds__(kThreads, 1)
gnn_fused_kernel(const __grid_constant__ GnnParams p, int* __restrict__ err_flag) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    unsigned char* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    unsigned char* s_h = smem + kOffH;
    unsigned char* s_p = smem + kOffP;
    float* s_src = reinterpret_cast<float*>(smem + kOffS);
    float* s_ln = reinterpret_cast<float*>(smem + kOffLn);
    int* s_nbr = reinterpret_cast<int*>(smem + kOffTopo);
    int* s_deg = s_nbr + 48 * kMaxDeg;
    float* s_par = reinterpret_cast<float*>(smem + kOffPar);
    uint64_t* w_bar = reinterpret_cast<uint64_t*>(smem + kOffBar);
    uint64_t* mma_bar = w_bar + 1;
    uint64_t* x_bar = w_bar + 2;                       // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 4);

    const int tid = threadIdx.x, warp = tid >> 5;
    const int r = tid & 127, q = tid >> 7, quad = warp & 3;
    const int J = p.J, rows_per_tile = p.gpc * J;
    const int group_rows = p.group_graphs * J;
    const long long n_tiles = p.n_groups * p.tiles_per_group;
    const long long my_tiles = (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;    // >= 1 by construction of the grid
    const uint32_t tile_bytes = static_cast<uint32_t>(rows_per_tile) * 128u;

    auto load_weights = [&](int layer) {               // one thread
        if ((layer & 1) == 0) {
            mbar_expect_tx(w_bar, kGatRows * 128);
            tma_load_5d(smem + kOffW, &p.w_gat[layer >> 1], w_bar, 0, 0, 0, 0, 0);
            tma_load_5d(smem + kOffW + 136 * 128, &p.w_gat[layer >> 1], w_bar, 0, 136, 0, 0, 0);
        } else {
            mbar_expect_tx(w_bar, 16384);
            tma_load_5d(smem + kOffW, &p.w_gc[layer >> 1], w_bar, 0, 0, 0, 0, 0);            // W_rel  (k 0..63)
            tma_load_5d(smem + kOffW + 8192, &p.w_gc[layer >> 1], w_bar, 64, 0, 0, 0, 0);    // W_root (k 64..127)
        }
    };

    pdl_launch_dependents();
    if (tid == 0) {
        for (int i = 0; i < 3; ++i) tma_prefetch_desc(&p.w_gat[i]);
        for (int i = 0; i < 2; ++i) tma_prefetch_desc(&p.w_gc[i]);
        tma_prefetch_desc(&p.x_in);
        tma_prefetch_desc(&p.x_out);
        mbar_init(w_bar, 1);
        mbar_init(mma_bar, 1);
        mbar_init(&x_bar[0], 1);
        mbar_init(&x_bar[1], 1);
        mbar_fence_init();
        load_weights(0);                               // weights are constants: no need to wait for the predecessor
    }
    if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
    for (int i = tid; i < J * kMaxDeg; i += kThreads) s_nbr[i] = p.nbr[i];
    for (int i = tid; i < J; i += kThreads) s_deg[i] = p.deg[i];
    for (int i = tid; i < 5 * 192; i += kThreads) {
        const int layer = i / 192, j = i - layer * 192;
        const float* src = j < 64 ? ((layer & 1) ? p.gc_bias[layer >> 1] : p.gat_bias[layer >> 1]) : j < 128 ? p.ln_w[layer] : p.ln_b[layer];
        s_par[i] = src[j & 63];
    }
    {   // zero the attention matrices (only the static neighbour positions are ever rewritten) and the rows of
        // the node tiles that TMA never writes (rows_per_tile .. 127): 0 x garbage must not become NaN
        uint4 z = make_uint4(0, 0, 0, 0);
        for (int i = tid; i < 2 * 32768 / 16; i += kThreads) reinterpret_cast<uint4*>(s_p)[i] = z;
        for (int i = tid; i < 2 * 16384 / 16; i += kThreads) {
            const int row = (i & 1023) >> 3;
            if (row >= rows_per_tile) reinterpret_cast<uint4*>(smem + kOffX)[i] = z;
        }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    pdl_wait();                                        // node features come from the previous kernel (proj_in)
    if (tid == 0) {
        mbar_expect_tx(&x_bar[0], tile_bytes);
        tma_load_5d(smem + kOffX, &p.x_in, &x_bar[0], 0, static_cast<int>(blockIdx.x % p.tiles_per_group) * rows_per_tile,
                    static_cast<int>(blockIdx.x / p.tiles_per_group), 0, 0);
    }
    const uint32_t idesc_h = umma_idesc_bf16(128, 256), idesc_s = umma_idesc_bf16(128, 16);
    const uint32_t idesc_agg = idesc_b_mn(128, 64), idesc_gc = umma_idesc_bf16(128, 64);
    const uint32_t w_addr = smem_u32(smem + kOffW), h_addr = smem_u32(s_h), p_addr = smem_u32(s_p);

    // static per-thread topology: the node this row holds, its neighbours' rows and their positions in P
    const bool valid_row = r < rows_per_tile;
    const int jloc = r % J, g0 = r - jloc;
    const int dg = valid_row ? s_deg[jloc] : 0;
    int idx[kMaxDeg + 1], pofs[kMaxDeg + 1];
    idx[0] = r;
#pragma unroll
    for (int k = 0; k < kMaxDeg; ++k) idx[k + 1] = k < dg ? g0 + s_nbr[jloc * kMaxDeg + k] : r;
#pragma unroll
    for (int k = 0; k <= kMaxDeg; ++k) pofs[k] = p_off(r, idx[k]);

    uint32_t mma_parity = 0, w_parity = 0;
    for (long long it = 0; it < my_tiles; ++it) {
        const long long tile = blockIdx.x + it * gridDim.x;
        const int group = static_cast<int>(tile / p.tiles_per_group);
        const int row0 = static_cast<int>(tile - static_cast<long long>(group) * p.tiles_per_group) * rows_per_tile;   // within the group
        const int buf = static_cast<int>(it & 1);
        unsigned char* s_x = smem + kOffX + buf * 16384;
        const uint32_t x_addr = smem_u32(s_x);
        const bool live = valid_row && row0 + r < group_rows;      // rows past the group are zero-filled on load, clipped on store
        if (tid == 0 && it + 1 < my_tiles) {           // prefetch the next tile into the other buffer
            tma_store_wait_read();                     // ... once the store that last read it has drained
            mbar_expect_tx(&x_bar[buf ^ 1], tile_bytes);
            const long long nt = tile + gridDim.x;
            tma_load_5d(smem + kOffX + (buf ^ 1) * 16384, &p.x_in, &x_bar[buf ^ 1], 0,
                        static_cast<int>(nt % p.tiles_per_group) * rows_per_tile, static_cast<int>(nt / p.tiles_per_group), 0, 0);
        }
        mbar_wait(&x_bar[buf], static_cast<uint32_t>((it >> 1) & 1), err_flag, 10);
        // residual stream: this thread's 16 features in fp32 registers
        float x[16];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const uint4 u = *reinterpret_cast<const uint4*>(s_x + sw128_off(r, q * 2 + c));
            const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) { x[c * 8 + 2 * e] = bf_lo(w4[e]); x[c * 8 + 2 * e + 1] = bf_hi(w4[e]); }
        }

#pragma unroll 1
        for (int layer = 0; layer < 5; ++layer) {
            float v[16];
            const bool more_weights = !(layer == 4 && it + 1 == my_tiles);
            // only the MMA-issuing thread consumes the weights (through the tensor core), so only it waits for them;
            // the other threads never see w_bar and cannot fall a phase behind it
            if (tid == 0) { mbar_wait(w_bar, w_parity, err_flag, 11); w_parity ^= 1; }
            if ((layer & 1) == 0) {
                // ================= GATConv =================
                if (tid == 0) {
                    tc_fence_after();
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        umma_bf16(tmem_base, umma_desc_sw128(x_addr + k * 32), umma_desc_sw128(w_addr + k * 32), idesc_h, k != 0);
                        umma_bf16(tmem_base + kColS, umma_desc_sw128(x_addr + k * 32), umma_desc_sw128(w_addr + 32768 + k * 32),
                                  idesc_s, k != 0);
                    }
                    umma_commit(mma_bar);
                }
                mbar_wait(mma_bar, mma_parity, err_flag, 12);
                mma_parity ^= 1;
                tc_fence_after();
                if (tid == 0 && more_weights) load_weights((layer + 1) % 5);
                // ---- phase A: attention logits of my node, my head's H row staged as bf16 (MMA B operand)
                float s_dst_q;
                {
                    uint32_t t[16];
                    tmem_ld_32x16(tmem_lane + kColS, t);
                    tmem_ld_wait();
                    const uint32_t dh = q == 0 ? t[4] : q == 1 ? t[5] : q == 2 ? t[6] : t[7];
                    const uint32_t dl = q == 0 ? t[12] : q == 1 ? t[13] : q == 2 ? t[14] : t[15];
                    s_dst_q = __uint_as_float(dh) + __uint_as_float(dl);
                    if (q == 0) {
                        *reinterpret_cast<float4*>(s_src + r * 4) =
                            make_float4(__uint_as_float(t[0]) + __uint_as_float(t[8]), __uint_as_float(t[1]) + __uint_as_float(t[9]),
                                        __uint_as_float(t[2]) + __uint_as_float(t[10]), __uint_as_float(t[3]) + __uint_as_float(t[11]));
                    }
                }
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                    uint32_t t[32];
                    tmem_ld_32x32(tmem_lane + q * 64 + cc * 32, t);
                    tmem_ld_wait();
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        uint4 o;
                        o.x = pack_bf16(__uint_as_float(t[c * 8]), __uint_as_float(t[c * 8 + 1]));
                        o.y = pack_bf16(__uint_as_float(t[c * 8 + 2]), __uint_as_float(t[c * 8 + 3]));
                        o.z = pack_bf16(__uint_as_float(t[c * 8 + 4]), __uint_as_float(t[c * 8 + 5]));
                        o.w = pack_bf16(__uint_as_float(t[c * 8 + 6]), __uint_as_float(t[c * 8 + 7]));
                        *reinterpret_cast<uint4*>(s_h + q * 16384 + sw128_off(r, cc * 4 + c)) = o;
                    }
                }
                tc_fence_before();
                fence_proxy_async_smem();
                __syncthreads();
                // ---- softmax over {self} + neighbours for head q (the head mean 1/4 is folded in)
                float alpha[kMaxDeg + 1];
                {
                    float m = -INFINITY;
#pragma unroll
                    for (int k = 0; k <= kMaxDeg; ++k) {
                        alpha[k] = k <= dg ? leaky(s_src[idx[k] * 4 + q] + s_dst_q) : -INFINITY;
                        m = fmaxf(m, alpha[k]);
                    }
                    float den = 0.f;
#pragma unroll
                    for (int k = 0; k <= kMaxDeg; ++k) { alpha[k] = k <= dg ? __expf(alpha[k] - m) : 0.f; den += alpha[k]; }
                    const float inv = 0.25f / den;
#pragma unroll
                    for (int k = 0; k <= kMaxDeg; ++k) alpha[k] *= inv;
                }
                // ---- two rounds over the two P buffers: heads {0,1}, then heads {2,3}
#pragma unroll 1
                for (int round = 0; round < 2; ++round) {
                    if ((q >> 1) == round) {
                        unsigned char* pb = s_p + (q & 1) * 32768;
#pragma unroll
                        for (int k = 0; k <= kMaxDeg; ++k)
                            if (k <= dg) *reinterpret_cast<__nv_bfloat16*>(pb + pofs[k]) = __float2bfloat16_rn(alpha[k]);
                    }
                    fence_proxy_async_smem();
                    __syncthreads();
                    if (tid == 0) {
                        tc_fence_after();
#pragma unroll
                        for (int hh = 0; hh < 2; ++hh) {
                            const int h = round * 2 + hh;
#pragma unroll
                            for (int kk = 0; kk < 8; ++kk)
                                umma_bf16(tmem_base + kColOut,
                                          umma_desc_sw128(p_addr + hh * 32768 + (kk >> 2) * 16384 + (kk & 3) * 32),
                                          umma_desc_sw128(h_addr + h * 16384 + kk * 2048), idesc_agg, (h | kk) != 0);
                        }
                        umma_commit(mma_bar);
                    }
                    mbar_wait(mma_bar, mma_parity, err_flag, 13);
                    mma_parity ^= 1;
                }
                tc_fence_after();
                {
                    uint32_t t[16];
                    tmem_ld_32x16(tmem_lane + kColOut + q * 16, t);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(t[i]);
                }
            } else {
                // ================= GraphConv =================
                if (q == 0) {                              // adjacency (no self loops) into P buffer 0
                    const __nv_bfloat16 one = __float2bfloat16_rn(1.f), zero = __float2bfloat16_rn(0.f);
                    *reinterpret_cast<__nv_bfloat16*>(s_p + pofs[0]) = zero;
#pragma unroll
                    for (int k = 1; k <= kMaxDeg; ++k)
                        if (k <= dg) *reinterpret_cast<__nv_bfloat16*>(s_p + pofs[k]) = one;
                }
                fence_proxy_async_smem();
                __syncthreads();
                if (tid == 0) {
                    tc_fence_after();
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk)       // AGG = Adj . X
                        umma_bf16(tmem_base + kColOut, umma_desc_sw128(p_addr + (kk >> 2) * 16384 + (kk & 3) * 32),
                                  umma_desc_sw128(x_addr + kk * 2048), idesc_agg, kk != 0);
                    umma_commit(mma_bar);
                }
                mbar_wait(mma_bar, mma_parity, err_flag, 14);
                mma_parity ^= 1;
                tc_fence_after();
                {
                    uint32_t t[16];
                    tmem_ld_32x16(tmem_lane + kColOut + q * 16, t);
                    tmem_ld_wait();
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        uint4 o;
                        o.x = pack_bf16(__uint_as_float(t[c * 8]), __uint_as_float(t[c * 8 + 1]));
                        o.y = pack_bf16(__uint_as_float(t[c * 8 + 2]), __uint_as_float(t[c * 8 + 3]));
                        o.z = pack_bf16(__uint_as_float(t[c * 8 + 4]), __uint_as_float(t[c * 8 + 5]));
                        o.w = pack_bf16(__uint_as_float(t[c * 8 + 6]), __uint_as_float(t[c * 8 + 7]));
                        *reinterpret_cast<uint4*>(s_h + sw128_off(r, q * 2 + c)) = o;
                    }
                }
                tc_fence_before();
                fence_proxy_async_smem();
                __syncthreads();
                if (tid == 0) {
                    tc_fence_after();
#pragma unroll
                    for (int k = 0; k < 4; ++k)           // W_rel . agg
                        umma_bf16(tmem_base, umma_desc_sw128(h_addr + k * 32), umma_desc_sw128(w_addr + k * 32), idesc_gc, k != 0);
#pragma unroll
                    for (int k = 0; k < 4; ++k)           // + W_root . x
                        umma_bf16(tmem_base, umma_desc_sw128(x_addr + k * 32), umma_desc_sw128(w_addr + 8192 + k * 32), idesc_gc, 1);
                    umma_commit(mma_bar);
                }
                mbar_wait(mma_bar, mma_parity, err_flag, 15);
                mma_parity ^= 1;
                tc_fence_after();
                if (tid == 0 && more_weights) load_weights((layer + 1) % 5);
                {
                    uint32_t t[16];
                    tmem_ld_32x16(tmem_lane + q * 16, t);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(t[i]);
                }
            }
            // ---- LayerNorm(64) over the four 16-feature quarters of the node -> LeakyReLU -> + residual
            {
                const float4* par = reinterpret_cast<const float4*>(s_par + layer * 192 + q * 16);   // warp-uniform: broadcasts
#pragma unroll
                for (int i4 = 0; i4 < 4; ++i4) {
                    const float4 b4 = par[i4];
                    v[i4 * 4] += b4.x; v[i4 * 4 + 1] += b4.y; v[i4 * 4 + 2] += b4.z; v[i4 * 4 + 3] += b4.w;
                }
                float s = 0.f, sq = 0.f;
#pragma unroll
                for (int i = 0; i < 16; ++i) { s += v[i]; sq = fmaf(v[i], v[i], sq); }
                *reinterpret_cast<float2*>(s_ln + (r * 4 + q) * 2) = make_float2(s, sq);
                __syncthreads();
                const float4 a = *reinterpret_cast<const float4*>(s_ln + r * 8);
                const float4 b = *reinterpret_cast<const float4*>(s_ln + r * 8 + 4);
                const float mean = (a.x + a.z + b.x + b.z) * (1.f / 64.f);
                const float rstd = rsqrtf(fmaxf((a.y + a.w + b.y + b.w) * (1.f / 64.f) - mean * mean, 0.f) + 1e-5f);
#pragma unroll
                for (int i4 = 0; i4 < 4; ++i4) {
                    const float4 w4 = par[16 + i4], b4 = par[32 + i4];
                    x[i4 * 4] += leaky((v[i4 * 4] - mean) * rstd * w4.x + b4.x);
                    x[i4 * 4 + 1] += leaky((v[i4 * 4 + 1] - mean) * rstd * w4.y + b4.y);
                    x[i4 * 4 + 2] += leaky((v[i4 * 4 + 2] - mean) * rstd * w4.z + b4.z);
                    x[i4 * 4 + 3] += leaky((v[i4 * 4 + 3] - mean) * rstd * w4.w + b4.w);
                }
            }
            if (!live) {
#pragma unroll
                for (int i = 0; i < 16; ++i) x[i] = 0.f;
            }
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint4 o;
                o.x = pack_bf16(x[c * 8], x[c * 8 + 1]); o.y = pack_bf16(x[c * 8 + 2], x[c * 8 + 3]);
                o.z = pack_bf16(x[c * 8 + 4], x[c * 8 + 5]); o.w = pack_bf16(x[c * 8 + 6], x[c * 8 + 7]);
                *reinterpret_cast<uint4*>(s_x + sw128_off(r, q * 2 + c)) = o;
            }
            tc_fence_before();
            fence_proxy_async_smem();
            __syncthreads();
        }
        if (tid == 0) {                                // the tile's final node features leave by TMA (rows past the
            tma_store_5d(&p.x_out, s_x, 0, row0, group, 0, 0);                 // end of the group are clipped)
            tma_store_commit();
        }
    }
    if (tid == 0) tma_store_wait_read();
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// =====================================================================================================
// v3: the same layer math, restructured so that TWO CTAs fit on one SM (<= 111 KB shared memory and 256 TMEM
// columns each) and hide each other's MMA / barrier latency -- the v2 kernel above is bound by the serial
// chain  MMA -> stage -> softmax -> MMA -> epilogue  of a single tile (issue slots ~30 % used).
//   * 256 threads = 2 per node row (32 features each);
//   * GAT heads are processed one at a time: H^h = X W_h^T lands in one of two 64-column TMEM buffers, is staged
//     as bf16 into one of two 16 KB shared-memory tiles, and P^h (single 32 KB buffer) multiplies it into OUT;
//     the per-head 8 KB weight slices stream through a two-slot ring, always two slices ahead;
//   * the attention logits (S = X U^T) and everything else are as in v2.
// Tried and dropped: a ninth, dedicated issuer warp with mbarrier hand-offs instead of __syncthreads.  Nine warps per
// CTA put three warps of each CTA on one SM sub-partition, whose 16 K registers then hold only one CTA's worth at
// 112 registers per thread: occupancy fell to one CTA per SM and the kernel was 37 % slower (profiles/).
// =====================================================================================================
namespace v3 {

constexpr int kThreads3 = 256;
constexpr int kOffW3 = 0;                               // 2 x 8 KB weight slots (GAT head slices / GraphConv W_rel, W_root)
constexpr int kOffU3 = 16384;                           // GAT attention rows [16][64] bf16
constexpr int kOffX3 = 18432;                           // node tile, bf16 [128][64] SW128
constexpr int kOffH3 = kOffX3 + 16384;                  // 2 x 16 KB: H^h tiles (GraphConv: AGG tile in slot 0)
constexpr int kOffP3 = kOffH3 + 2 * 16384;              // attention / adjacency matrix [128][128] bf16
constexpr int kOffS3 = kOffP3 + 32768;                  // s_src [128][4] fp32
constexpr int kOffLn3 = kOffS3 + kRows * 4 * 4;         // LayerNorm partials [128][2][2] fp32
constexpr int kOffTopo3 = kOffLn3 + kRows * 2 * 2 * 4;  // nbr [48][6], deg [48]
constexpr int kOffPar3 = kOffTopo3 + 48 * kMaxDeg * 4 + 48 * 4;
constexpr int kOffBar3 = kOffPar3 + 5 * 192 * 4;
constexpr int kSmemBytes3 = kOffBar3 + 64 + 1024;
constexpr uint32_t kColHB = 0, kColS3 = 128, kColOut3 = 160;        // TMEM: HB0 [0,64) HB1 [64,128) S [128,144) OUT [160,224)
static_assert(kOffX3 % 1024 == 0 && kOffH3 % 1024 == 0 && kOffP3 % 1024 == 0, "swizzled tiles need 1024 B alignment");
static_assert(2 * (kSmemBytes3 + 1024) <= 228 * 1024, "two CTAs per SM");

struct Gnn3Params {
    CUtensorMap w_head[3];               // GAT [272][64] bf16, box 64 x 64 (head slices)
    CUtensorMap w_att[3];                // same tensor, box 64 x 16 (rows 256..271)
    CUtensorMap w_gc[2];                 // [64][128] bf16, box 64 x 64
    CUtensorMap x_in, x_out;
    const float* gat_bias[3];
    const float* gc_bias[2];
    const float* ln_w[5];
    const float* ln_b[5];
    const int* nbr;
    const int* deg;
    int J, gpc;
    long long n_groups;
    int group_graphs, tiles_per_group;
};

__global__ void __launch_bounds__(kThreads3, 2)
gnn3_kernel(const __grid_constant__ Gnn3Params p, int* __restrict__ err_flag) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    unsigned char* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    unsigned char* s_x = smem + kOffX3;
    unsigned char* s_h = smem + kOffH3;
    unsigned char* s_p = smem + kOffP3;
    float* s_src = reinterpret_cast<float*>(smem + kOffS3);
    float* s_ln = reinterpret_cast<float*>(smem + kOffLn3);
    int* s_nbr = reinterpret_cast<int*>(smem + kOffTopo3);
    int* s_deg = s_nbr + 48 * kMaxDeg;
    float* s_par = reinterpret_cast<float*>(smem + kOffPar3);
    uint64_t* w_bar = reinterpret_cast<uint64_t*>(smem + kOffBar3);    // [2] weight slots (waited on by thread 0 only)
    uint64_t* h_bar = w_bar + 2;                                       // [2] H buffer b holds a finished MMA
    uint64_t* o_bar = w_bar + 4;                                       // OUT updated / P and H tile consumed
    uint64_t* x_bar = w_bar + 5;                                       // node tile landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 6);

    const int tid = threadIdx.x, warp = tid >> 5;
    const int r = tid & 127, half = tid >> 7, quad = warp & 3;
    const int J = p.J, rows_per_tile = p.gpc * J;
    const int group_rows = p.group_graphs * J;
    const long long n_tiles = p.n_groups * p.tiles_per_group;
    const long long my_tiles = (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
    const uint32_t tile_bytes = static_cast<uint32_t>(rows_per_tile) * 128u;
    const long long n_items = my_tiles * 16;            // weight slices this CTA consumes, 16 per tile

    // weight slice `item` (0..15 within a tile: GAT heads 0-3 | GC rel, root | GAT | GC | GAT) into slot item & 1
    auto load_item = [&](long long item) {             // thread 0 only
        if (item >= n_items) return;
        const int i = static_cast<int>(item & 15), slot = i & 1;
        const int layer = i < 4 ? 0 : i < 6 ? 1 : i < 10 ? 2 : i < 12 ? 3 : 4;
        unsigned char* dst = smem + kOffW3 + slot * 8192;
        if ((layer & 1) == 0) {
            const int h = i - (layer == 0 ? 0 : layer == 2 ? 6 : 12);
            mbar_expect_tx(&w_bar[slot], h == 0 ? 8192 + 2048 : 8192);
            tma_load_5d(dst, &p.w_head[layer >> 1], &w_bar[slot], 0, h * 64, 0, 0, 0);
            if (h == 0) tma_load_5d(smem + kOffU3, &p.w_att[layer >> 1], &w_bar[slot], 0, 256, 0, 0, 0);
        } else {
            mbar_expect_tx(&w_bar[slot], 8192);
            tma_load_5d(dst, &p.w_gc[layer >> 1], &w_bar[slot], slot * 64, 0, 0, 0, 0);   // slot 0: W_rel (k 0..63), slot 1: W_root
        }
    };

    pdl_launch_dependents();
    if (tid == 0) {
        for (int i = 0; i < 3; ++i) { tma_prefetch_desc(&p.w_head[i]); tma_prefetch_desc(&p.w_att[i]); }
        for (int i = 0; i < 2; ++i) tma_prefetch_desc(&p.w_gc[i]);
        tma_prefetch_desc(&p.x_in);
        tma_prefetch_desc(&p.x_out);
        for (int i = 0; i < 6; ++i) mbar_init(&w_bar[i], 1);
        mbar_fence_init();
        load_item(0);                                  // weights are constants: no need to wait for the predecessor
        load_item(1);
    }
    if (warp == 1) { tmem_alloc(tmem_slot, 256); tmem_relinquish(); }
    for (int i = tid; i < J * kMaxDeg; i += kThreads3) s_nbr[i] = p.nbr[i];
    for (int i = tid; i < J; i += kThreads3) s_deg[i] = p.deg[i];
    for (int i = tid; i < 5 * 192; i += kThreads3) {
        const int layer = i / 192, j = i - layer * 192;
        const float* src = j < 64 ? ((layer & 1) ? p.gc_bias[layer >> 1] : p.gat_bias[layer >> 1]) : j < 128 ? p.ln_w[layer] : p.ln_b[layer];
        s_par[i] = src[j & 63];
    }
    {
        uint4 z = make_uint4(0, 0, 0, 0);
        for (int i = tid; i < 32768 / 16; i += kThreads3) reinterpret_cast<uint4*>(s_p)[i] = z;
        for (int i = tid; i < 16384 / 16; i += kThreads3)
            if ((i >> 3) >= rows_per_tile) reinterpret_cast<uint4*>(s_x)[i] = z;
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
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
    const uint32_t w_addr = smem_u32(smem + kOffW3), u_addr = smem_u32(smem + kOffU3), x_addr = smem_u32(s_x),
                   h_addr = smem_u32(s_h), p_addr = smem_u32(s_p);

    const bool valid_row = r < rows_per_tile;
    const int jloc = r % J, g0 = r - jloc;
    const int dg = valid_row ? s_deg[jloc] : 0;
    int idx[kMaxDeg + 1];
    idx[0] = r;
#pragma unroll
    for (int k = 0; k < kMaxDeg; ++k) idx[k + 1] = k < dg ? g0 + s_nbr[jloc * kMaxDeg + k] : r;

    uint32_t wpar[2] = {0, 0}, hpar[2] = {0, 0}, opar = 0;     // wpar is used by thread 0 only
    long long item = 0;                                // first weight slice of the current layer (thread 0's view)
    for (long long it = 0; it < my_tiles; ++it) {
        const long long tile = blockIdx.x + it * gridDim.x;
        const int group = static_cast<int>(tile / p.tiles_per_group);
        const int row0 = static_cast<int>(tile - static_cast<long long>(group) * p.tiles_per_group) * rows_per_tile;
        const bool live = valid_row && row0 + r < group_rows;
        mbar_wait(x_bar, static_cast<uint32_t>(it & 1), err_flag, 30);
        float x[32];                                   // residual stream: this thread's 32 features in fp32
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const uint4 u = *reinterpret_cast<const uint4*>(s_x + sw128_off(r, half * 4 + c));
            const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) { x[c * 8 + 2 * e] = bf_lo(w4[e]); x[c * 8 + 2 * e + 1] = bf_hi(w4[e]); }
        }

#pragma unroll 1
        for (int layer = 0; layer < 5; ++layer) {
            float v[32];
            if ((layer & 1) == 0) {
                // ================= GATConv, one head at a time =================
                if (tid == 0) {
                    tc_fence_after();
                    mbar_wait(&w_bar[0], wpar[0], err_flag, 31); wpar[0] ^= 1;          // head 0 slice + attention rows
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        umma_bf16(tmem_base + kColS3, umma_desc_sw128(x_addr + k * 32), umma_desc_sw128(u_addr + k * 32), idesc_s, k != 0);
                        umma_bf16(tmem_base + kColHB, umma_desc_sw128(x_addr + k * 32), umma_desc_sw128(w_addr + k * 32), idesc_h, k != 0);
                    }
                    umma_commit(&h_bar[0]);
                    mbar_wait(&w_bar[1], wpar[1], err_flag, 31); wpar[1] ^= 1;          // head 1 slice
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16(tmem_base + kColHB + 64, umma_desc_sw128(x_addr + k * 32), umma_desc_sw128(w_addr + 8192 + k * 32), idesc_h, k != 0);
                    umma_commit(&h_bar[1]);
                }
                float s_dst2[2] = {0.f, 0.f};           // attention logit (destination part) of my two heads: half, half + 2
#pragma unroll 1
                for (int h = 0; h < 4; ++h) {
                    const int b = h & 1;
                    mbar_wait(&h_bar[b], hpar[b], err_flag, 32);
                    hpar[b] ^= 1;
                    tc_fence_after();
                    if (tid == 0) load_item(item + h + 2);             // slot b is free: MMA1_h has read it
                    if (h == 0) {
                        uint32_t t[16];
                        tmem_ld_32x16(tmem_lane + kColS3, t);
                        tmem_ld_wait();
                        s_dst2[0] = half == 0 ? __uint_as_float(t[4]) + __uint_as_float(t[12]) : __uint_as_float(t[5]) + __uint_as_float(t[13]);
                        s_dst2[1] = half == 0 ? __uint_as_float(t[6]) + __uint_as_float(t[14]) : __uint_as_float(t[7]) + __uint_as_float(t[15]);
                        if (half == 0) {
                            *reinterpret_cast<float4*>(s_src + r * 4) =
                                make_float4(__uint_as_float(t[0]) + __uint_as_float(t[8]), __uint_as_float(t[1]) + __uint_as_float(t[9]),
                                            __uint_as_float(t[2]) + __uint_as_float(t[10]), __uint_as_float(t[3]) + __uint_as_float(t[11]));
                        }
                    }
                    {   // stage my half of H^h (32 features) as bf16 into tile b
                        uint32_t t[32];
                        tmem_ld_32x32(tmem_lane + kColHB + b * 64 + half * 32, t);
                        tmem_ld_wait();
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            uint4 o;
                            o.x = pack_bf16(__uint_as_float(t[c * 8]), __uint_as_float(t[c * 8 + 1]));
                            o.y = pack_bf16(__uint_as_float(t[c * 8 + 2]), __uint_as_float(t[c * 8 + 3]));
                            o.z = pack_bf16(__uint_as_float(t[c * 8 + 4]), __uint_as_float(t[c * 8 + 5]));
                            o.w = pack_bf16(__uint_as_float(t[c * 8 + 6]), __uint_as_float(t[c * 8 + 7]));
                            *reinterpret_cast<uint4*>(s_h + b * 16384 + sw128_off(r, half * 4 + c)) = o;
                        }
                    }
                    tc_fence_before();
                    fence_proxy_async_smem();
                    __syncthreads();                    // H^h staged (and, for h = 0, s_src visible); TMEM buffer b is free
                    // P^h is written by the threads of half (h & 1); it needs the previous head's MMA to have drained P
                    if (h > 0) { mbar_wait(o_bar, opar, err_flag, 33); opar ^= 1; }
                    if (half == b) {
                        const float sd = (h >> 1) ? s_dst2[1] : s_dst2[0];
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
                        const float inv = 0.25f / den;                  // softmax normaliser and the head mean
#pragma unroll
                        for (int k = 0; k <= kMaxDeg; ++k)
                            if (k <= dg) *reinterpret_cast<__nv_bfloat16*>(s_p + p_off(r, idx[k])) = __float2bfloat16_rn(alpha[k] * inv);
                    }
                    fence_proxy_async_smem();
                    __syncthreads();
                    if (tid == 0) {
                        tc_fence_after();
#pragma unroll
                        for (int kk = 0; kk < 8; ++kk)
                            umma_bf16(tmem_base + kColOut3, umma_desc_sw128(p_addr + (kk >> 2) * 16384 + (kk & 3) * 32),
                                      umma_desc_sw128(h_addr + b * 16384 + kk * 2048), idesc_agg, (h | kk) != 0);
                        umma_commit(o_bar);
                        if (h + 2 < 4) {                // H^{h+2} into the TMEM buffer drained at the barrier above; its weight
                            mbar_wait(&w_bar[b], wpar[b], err_flag, 31); wpar[b] ^= 1;      // slice had this whole round to land
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                umma_bf16(tmem_base + kColHB + b * 64, umma_desc_sw128(x_addr + k * 32),
                                          umma_desc_sw128(w_addr + b * 8192 + k * 32), idesc_h, k != 0);
                            umma_commit(&h_bar[b]);
                        }
                    }
                }
                mbar_wait(o_bar, opar, err_flag, 34);
                opar ^= 1;
                tc_fence_after();
                if (tid == 0) item += 4;
                {
                    uint32_t t[32];
                    tmem_ld_32x32(tmem_lane + kColOut3 + half * 32, t);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(t[i]);
                }
            } else {
                // ================= GraphConv =================
                if (half == 0) {                           // adjacency (no self loops) into P
                    const __nv_bfloat16 one = __float2bfloat16_rn(1.f), zero = __float2bfloat16_rn(0.f);
                    *reinterpret_cast<__nv_bfloat16*>(s_p + p_off(r, idx[0])) = zero;
#pragma unroll
                    for (int k = 1; k <= kMaxDeg; ++k)
                        if (k <= dg) *reinterpret_cast<__nv_bfloat16*>(s_p + p_off(r, idx[k])) = one;
                }
                fence_proxy_async_smem();
                __syncthreads();
                if (tid == 0) {
                    tc_fence_after();
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk)       // AGG = Adj . X (exact: 0/1 weights, fp32 accumulation)
                        umma_bf16(tmem_base + kColOut3, umma_desc_sw128(p_addr + (kk >> 2) * 16384 + (kk & 3) * 32),
                                  umma_desc_sw128(x_addr + kk * 2048), idesc_agg, kk != 0);
                    umma_commit(o_bar);
                }
                mbar_wait(o_bar, opar, err_flag, 35);
                opar ^= 1;
                tc_fence_after();
                {
                    uint32_t t[32];
                    tmem_ld_32x32(tmem_lane + kColOut3 + half * 32, t);
                    tmem_ld_wait();
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        uint4 o;
                        o.x = pack_bf16(__uint_as_float(t[c * 8]), __uint_as_float(t[c * 8 + 1]));
                        o.y = pack_bf16(__uint_as_float(t[c * 8 + 2]), __uint_as_float(t[c * 8 + 3]));
                        o.z = pack_bf16(__uint_as_float(t[c * 8 + 4]), __uint_as_float(t[c * 8 + 5]));
                        o.w = pack_bf16(__uint_as_float(t[c * 8 + 6]), __uint_as_float(t[c * 8 + 7]));
                        *reinterpret_cast<uint4*>(s_h + sw128_off(r, half * 4 + c)) = o;
                    }
                }
                tc_fence_before();
                fence_proxy_async_smem();
                __syncthreads();
                if (tid == 0) {
                    tc_fence_after();
                    mbar_wait(&w_bar[0], wpar[0], err_flag, 31); wpar[0] ^= 1;
                    mbar_wait(&w_bar[1], wpar[1], err_flag, 31); wpar[1] ^= 1;
#pragma unroll
                    for (int k = 0; k < 4; ++k)           // W_rel . agg
                        umma_bf16(tmem_base + kColHB, umma_desc_sw128(h_addr + k * 32), umma_desc_sw128(w_addr + k * 32), idesc_h, k != 0);
#pragma unroll
                    for (int k = 0; k < 4; ++k)           // + W_root . x
                        umma_bf16(tmem_base + kColHB, umma_desc_sw128(x_addr + k * 32), umma_desc_sw128(w_addr + 8192 + k * 32), idesc_h, 1);
                    umma_commit(&h_bar[0]);
                }
                mbar_wait(&h_bar[0], hpar[0], err_flag, 36);
                hpar[0] ^= 1;
                tc_fence_after();
                if (tid == 0) { load_item(item + 2); load_item(item + 3); item += 2; }
                {
                    uint32_t t[32];
                    tmem_ld_32x32(tmem_lane + kColHB + half * 32, t);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(t[i]);
                }
            }
            // ---- + bias, LayerNorm(64) over the two 32-feature halves of the node -> LeakyReLU -> + residual
            {
                const float4* par = reinterpret_cast<const float4*>(s_par + layer * 192 + half * 32);
#pragma unroll
                for (int i4 = 0; i4 < 8; ++i4) {
                    const float4 b4 = par[i4];
                    v[i4 * 4] += b4.x; v[i4 * 4 + 1] += b4.y; v[i4 * 4 + 2] += b4.z; v[i4 * 4 + 3] += b4.w;
                }
                float s = 0.f, sq = 0.f;
#pragma unroll
                for (int i = 0; i < 32; ++i) { s += v[i]; sq = fmaf(v[i], v[i], sq); }
                *reinterpret_cast<float2*>(s_ln + (r * 2 + half) * 2) = make_float2(s, sq);
                __syncthreads();
                const float4 a = *reinterpret_cast<const float4*>(s_ln + r * 4);
                const float mean = (a.x + a.z) * (1.f / 64.f);
                const float rstd = rsqrtf(fmaxf((a.y + a.w) * (1.f / 64.f) - mean * mean, 0.f) + 1e-5f);
#pragma unroll
                for (int i4 = 0; i4 < 8; ++i4) {
                    const float4 w4 = par[16 + i4], b4 = par[32 + i4];
                    x[i4 * 4] += leaky((v[i4 * 4] - mean) * rstd * w4.x + b4.x);
                    x[i4 * 4 + 1] += leaky((v[i4 * 4 + 1] - mean) * rstd * w4.y + b4.y);
                    x[i4 * 4 + 2] += leaky((v[i4 * 4 + 2] - mean) * rstd * w4.z + b4.z);
                    x[i4 * 4 + 3] += leaky((v[i4 * 4 + 3] - mean) * rstd * w4.w + b4.w);
                }
            }
            if (!live) {
#pragma unroll
                for (int i = 0; i < 32; ++i) x[i] = 0.f;
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint4 o;
                o.x = pack_bf16(x[c * 8], x[c * 8 + 1]); o.y = pack_bf16(x[c * 8 + 2], x[c * 8 + 3]);
                o.z = pack_bf16(x[c * 8 + 4], x[c * 8 + 5]); o.w = pack_bf16(x[c * 8 + 6], x[c * 8 + 7]);
                *reinterpret_cast<uint4*>(s_x + sw128_off(r, half * 4 + c)) = o;
            }
            tc_fence_before();
            fence_proxy_async_smem();
            __syncthreads();
        }
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

}  // namespace v3
