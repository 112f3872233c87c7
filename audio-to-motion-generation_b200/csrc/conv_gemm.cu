// Tap-offset implicit-GEMM convolution engine on the 5th-generation tensor cores (SURVEY.md K4/K5/K8,
// and K2 for the 2-D encoder convolutions).  See conv_gemm.cuh for the formulation.
//
// One CTA computes a 128 x BLOCK_N output tile:
//   warp 0    : TMA producer  -- per K block (one tap, 64 channels) a 5-D activation box (128 rows
//               x 64 bf16, OOB rows zero-filled = halo / batch tail) and a 2-D weight box
//               (BLOCK_N x 64 bf16), both 128B-swizzled, into a STAGES-deep shared-memory ring
//   warp 1    : tcgen05.mma issuer (one elected thread), fp32 accumulator in TMEM; tcgen05.commit
//               releases ring slots and finally signals the epilogue
//   warps 2-5 : epilogue -- tcgen05.ld the accumulator (one TMEM lane = one output row per thread),
//               + folded bias, LeakyReLU(0.2) / ReLU, bf16 (or fp32) vector stores, strided so
//               ConvTranspose parities and the [B,T,104] pose layout are written in place
// Two CTAs fit per SM (3 stages x 32 KB), so one CTA's epilogue overlaps the other's main loop.
//
// Three kernels share that structure (the planner picks, conv_gemm_plan):
//   conv_gemm_kernel<BLOCK_N, STAGES>  one 128 x {32, 64, 128, 256} tile per CTA (cta_group::1): narrow / ragged widths,
//                                      fp32 and split-K outputs
//   conv_gemm_pair_single_kernel       one 256 x 256 tile per 2-CTA cluster (cta_group::2), layers with one round of tiles
//   conv_gemm_pair_kernel              the same tile, persistent over a tile list, accumulator double-buffered in tensor
//                                      memory so that a tile's epilogue runs under the next tile's MMAs
// The two pair kernels can also apply LayerNorm(256) to the accumulator rows in their epilogue (proj_out).
#include <cstdlib>
#include <cstring>
#include <mutex>
#include "conv_gemm.cuh"

void a2m_count_launch();

namespace a2m {

namespace {

constexpr int kThreads = 192;
constexpr int kABytes = kBlockM * kBlockK * 2;          // 16 KB

// 128 x 128 tiles: 3 stages x 32 KB, two CTAs per SM (one CTA's epilogue overlaps the other's main loop).
// 128 x 256 tiles (wide layers): 4 stages x 48 KB, one CTA per SM -- 33 % fewer operand bytes per FLOP from L2
// (85 instead of 64 FLOP/B), which is what bounds the large-K UNet layers.
template <int BLOCK_N, int STAGES>
struct Smem {
    static constexpr int kStages = STAGES;
    static constexpr int kBBytes = BLOCK_N * kBlockK * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kBarOffset = kStages * kStageBytes;
    static constexpr int kBiasOffset = kBarOffset + 128;        // folded bias of this N tile, fp32 [BLOCK_N]
    static constexpr int kTotal = kBiasOffset + BLOCK_N * 4 + 1024;      // + alignment slack
};

// (A 2-CTA-cluster variant that multicast the shared weight tile was measured and retired: the coupling of the two CTAs
// cost more than the halved weight traffic saved on every layer class of this model -- DESIGN.md section 4.)
template <int BLOCK_N, int STAGES>
__global__ void __launch_bounds__(kThreads)
conv_gemm_kernel(const __grid_constant__ ConvGemmParams p, const float* __restrict__ bias, void* __restrict__ out,
                 int* __restrict__ err_flag) {
    using S = Smem<BLOCK_N, STAGES>;
    constexpr int kStages = S::kStages;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    unsigned char* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);      // 1024 B: SWIZZLE_128B atom
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::kBarOffset);
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* accum_bar = empty_bar + kStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr uint32_t kTmemCols = BLOCK_N < 32 ? 32 : BLOCK_N;

    // tile -> base coordinates in dims 1..4 (dim 1 fastest)
    int base[4];
    {
        int t = blockIdx.x;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int ti = t % p.tiles[i];
            t /= p.tiles[i];
            base[i] = ti * p.box[i];
        }
    }
    const int n0 = blockIdx.y * BLOCK_N;
    const int kb_lo = static_cast<int>(static_cast<long long>(blockIdx.z) * p.k_blocks / p.split_k);
    const int kb_hi = static_cast<int>(static_cast<long long>(blockIdx.z + 1) * p.k_blocks / p.split_k);

    pdl_launch_dependents();
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.a_map[0]);
        tma_prefetch_desc(&p.a_map[1]);
        tma_prefetch_desc(&p.b_map);
        for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(accum_bar, 1);
        mbar_fence_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, kTmemCols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();                                      // the prologue above overlapped the previous kernel's tail

    if (warp == 0) {
        if (lane == 0) {
            int kb = 0;
            bool ok = true;
            for (int t = 0; t < p.n_taps && ok; ++t) {
                const CUtensorMap* amap = &p.a_map[p.tap_src[t]];
                const int c1 = base[0] + p.tap_off[t][0], c2 = base[1] + p.tap_off[t][1];
                const int c3 = base[2] + p.tap_off[t][2], c4 = base[3] + p.tap_off[t][3];
                for (int ch = 0; ch < p.tap_chunks[t]; ++ch, ++kb) {
                    if (kb < kb_lo || kb >= kb_hi) continue;
                    const int lk = kb - kb_lo;
                    const int s = lk % kStages;
                    if (!mbar_wait(&empty_bar[s], ((lk / kStages) & 1) ^ 1, err_flag, 1)) { ok = false; break; }
                    mbar_expect_tx(&full_bar[s], S::kStageBytes);
                    unsigned char* stage = smem + s * S::kStageBytes;
                    tma_load_5d(stage, amap, &full_bar[s], ch * kBlockK, c1, c2, c3, c4);
                    tma_load_5d(stage + kABytes, &p.b_map, &full_bar[s], kb * kBlockK, n0, 0, 0, 0);   // all maps are rank 5
                    if (BLOCK_N > 128)               // the weight box is 128 rows: second half of a 256-wide tile
                        tma_load_5d(stage + kABytes + 128 * kBlockK * 2, &p.b_map, &full_bar[s], kb * kBlockK, n0 + 128, 0, 0, 0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_bf16(kBlockM, BLOCK_N);
            for (int kb = 0; kb < kb_hi - kb_lo; ++kb) {
                const int s = kb % kStages;
                if (!mbar_wait(&full_bar[s], (kb / kStages) & 1, err_flag, 2)) break;
                tc_fence_after();
                const uint32_t a_addr = smem_u32(smem + s * S::kStageBytes);
                const uint32_t b_addr = a_addr + kABytes;
#pragma unroll
                for (int k = 0; k < kBlockK / 16; ++k) {
                    // 16 bf16 = 32 B along K inside the 128 B swizzle row
                    umma_bf16(tmem_base, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(b_addr + k * 32), idesc,
                              (kb | k) != 0);
                }
                umma_commit(&empty_bar[s]);          // slot reusable once these MMAs have read it
            }
            umma_commit(accum_bar);                  // accumulator complete
        }
    } else {
        const int quad = warp & 3;                   // TMEM lane quadrant this warp may access
        const int row = quad * 32 + lane;
        // the tile's bias goes to shared memory while the main loop runs (the epilogue then reads broadcasts)
        float* s_bias = reinterpret_cast<float*>(smem + S::kBiasOffset);
        for (int i = threadIdx.x - 64; i < BLOCK_N; i += 128)
            s_bias[i] = (bias != nullptr && n0 + i < p.N && blockIdx.z == 0) ? __ldg(bias + n0 + i) : 0.f;
        named_barrier(2, 128);
        long long off = p.out_base;
        bool valid = true;
        {
            int r = row;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int ii = r % p.box[i];
                r /= p.box[i];
                const int c = base[i] + ii;
                valid = valid && (c < p.m_extent[i]);
                off += static_cast<long long>(c) * p.out_stride[i];
            }
        }
        off += static_cast<long long>(blockIdx.z) * p.split_stride;     // split-K: one fp32 partial plane per split
        mbar_wait(accum_bar, 0, err_flag, 3);
        tc_fence_after();
        if (BLOCK_N >= 64 && p.c_tma) {
            // bf16 output through shared memory + TMA store: the (now idle) operand ring is reused as
            // [128 rows][64 cols] 128B-swizzled sub-tiles, so global writes are whole 128-byte rows.
            tma_prefetch_desc(&p.c_map);
#pragma unroll 1
            for (int c0 = 0; c0 < BLOCK_N; c0 += 64) {
                const int n = n0 + c0;
                if (n >= p.N) break;                  // uniform over the epilogue warps
                unsigned char* st = smem + (c0 >> 6) * 16384;
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    uint32_t v[32];
                    tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + c0 + hf * 32, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        float f[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            float x = __uint_as_float(v[j + e]) + s_bias[c0 + hf * 32 + j + e];
                            if (p.act == kActLeaky) x = x > 0.f ? x : kLeakySlope * x;
                            else if (p.act == kActRelu) x = fmaxf(x, 0.f);
                            f[e] = x;
                        }
                        uint4 q;
                        __nv_bfloat162 h0 = __floats2bfloat162_rn(f[0], f[1]);
                        __nv_bfloat162 h1 = __floats2bfloat162_rn(f[2], f[3]);
                        __nv_bfloat162 h2 = __floats2bfloat162_rn(f[4], f[5]);
                        __nv_bfloat162 h3 = __floats2bfloat162_rn(f[6], f[7]);
                        q.x = *reinterpret_cast<uint32_t*>(&h0); q.y = *reinterpret_cast<uint32_t*>(&h1);
                        q.z = *reinterpret_cast<uint32_t*>(&h2); q.w = *reinterpret_cast<uint32_t*>(&h3);
                        const int chunk = hf * 4 + (j >> 3);
                        *reinterpret_cast<uint4*>(st + row * 128 + ((chunk ^ (row & 7)) << 4)) = q;
                    }
                }
                fence_proxy_async_smem();
                named_barrier(1, 128);
                if (warp == 2 && lane == 0) {
                    tma_store_5d(&p.c_map, st, n, base[0], base[1], base[2], base[3]);
                    tma_store_commit();
                }
            }
            if (warp == 2 && lane == 0) tma_store_wait_read();
        } else {
#pragma unroll 1
        for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
            const int n = n0 + c0;
            if (n >= p.N) break;                      // warp-uniform
            uint32_t v[32];
            tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + c0, v);
            tmem_ld_wait();
            float f[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                float x = __uint_as_float(v[j]) + s_bias[c0 + j];
                if (p.act == kActLeaky) x = x > 0.f ? x : kLeakySlope * x;
                else if (p.act == kActRelu) x = fmaxf(x, 0.f);
                f[j] = x;
            }
            if (!valid) continue;

            if (p.out_type == kOutBf16) {
                __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out) + off + n;
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                    if (n + j < p.N) {                // N % 8 == 0 is enforced by the planner
                        uint4 q;
                        __nv_bfloat162 h0 = __floats2bfloat162_rn(f[j], f[j + 1]);
                        __nv_bfloat162 h1 = __floats2bfloat162_rn(f[j + 2], f[j + 3]);
                        __nv_bfloat162 h2 = __floats2bfloat162_rn(f[j + 4], f[j + 5]);
                        __nv_bfloat162 h3 = __floats2bfloat162_rn(f[j + 6], f[j + 7]);
                        q.x = *reinterpret_cast<uint32_t*>(&h0); q.y = *reinterpret_cast<uint32_t*>(&h1);
                        q.z = *reinterpret_cast<uint32_t*>(&h2); q.w = *reinterpret_cast<uint32_t*>(&h3);
                        *reinterpret_cast<uint4*>(o + j) = q;
                    }
                }
            } else {
                float* o = reinterpret_cast<float*>(out) + off + n;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    if (n + j + 3 < p.N) {
                        *reinterpret_cast<float4*>(o + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            if (n + j + e < p.N) o[j + e] = f[j + e];
                    }
                }
            }
        }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

// LayerNorm statistics of one accumulator row (256 fp32 columns of my tensor-memory lane, + bias): two passes, thread-local,
// so a row's result never depends on which rows share its tile or its launch.
__device__ __forceinline__ void ln256_row_stats(uint32_t tmem_row, const float* s_bias, float& mean, float& rstd) {
    float sum = 0.f;
#pragma unroll 1
    for (int c = 0; c < 256; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_row + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) sum += __uint_as_float(v[j]) + s_bias[c + j];
    }
    mean = sum * (1.f / 256.f);
    float sq = 0.f;
#pragma unroll 1
    for (int c = 0; c < 256; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_row + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) { const float dlt = __uint_as_float(v[j]) + s_bias[c + j] - mean; sq = fmaf(dlt, dlt, sq); }
    }
    rstd = rsqrtf(sq * (1.f / 256.f) + 1e-5f);
}

// ---------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2), persistent: two neighbouring M tiles and a 256-wide N tile are ONE 256 x 256 MMA tile.
// Each CTA stages its own 128 activation rows and its own half of the weight tile (128 of the 256 rows): 32 KB per K
// block instead of 48 KB for the same 128 x 256 x 64 of work per SM.  The leader (even) CTA's elected thread issues the
// MMAs for the pair; tcgen05.commit is multicast, so each CTA's producer sees its own ring slots released and each
// CTA's epilogue warps see their own accumulator half (rows 0..127 in the leader's tensor memory, 128..255 in the
// peer's).  Both producers count their bytes on the LEADER's "full" barrier.
// One cluster per SM pair walks the tile list (tile = cluster + i * clusters).  The accumulator is double-buffered in
// tensor memory (2 x 256 columns): the epilogue of tile i (tcgen05.ld -> bias / activation -> bf16 -> swizzled staging
// -> TMA store) runs while the MMAs of tile i + 1 fill the other buffer; both CTAs' epilogues hand a drained buffer back
// to the leader's MMA thread through a cluster-scope mbarrier arrive.  Five 32 KB ring stages + 2 x 16 KB staging.
// Layers whose tile list fits one round (every cluster has exactly one tile: the short decoder layers) use
// conv_gemm_pair_single_kernel below instead.
// ---------------------------------------------------------------------------------------------
struct SmemPair {
    static constexpr int kStages = 5;
    static constexpr int kStageBytes = 2 * kABytes;             // A: 128 rows, B: my 128 of the tile's 256 rows
    static constexpr int kStagingOffset = kStages * kStageBytes;            // 2 x [128 rows][64 cols] bf16, 128B-swizzled
    static constexpr int kBarOffset = kStagingOffset + 2 * 16384;
    static constexpr int kBiasOffset = kBarOffset + 256;
    static constexpr int kTotal = kBiasOffset + 3 * 256 * 4 + 1024;        // bias | LN gamma | LN beta
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads)
conv_gemm_pair_kernel(const __grid_constant__ ConvGemmParams p, const float* __restrict__ bias, int* __restrict__ err_flag) {
    using S = SmemPair;
    constexpr int kStages = S::kStages;
    constexpr int kN = 256;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    unsigned char* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::kBarOffset);     // waited on in the leader only
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* acc_full = empty_bar + kStages;          // [2] accumulator buffer complete (multicast commit)
    uint64_t* acc_empty = acc_full + 2;                // [2] buffer drained by BOTH CTAs' epilogues (leader's copy is used)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int m_tiles = p.tiles[0] * p.tiles[1] * p.tiles[2] * p.tiles[3];
    const int n_tiles = p.N / kN;
    const int work_total = ((m_tiles + 1) >> 1) * n_tiles;
    const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

    // work item -> my M tile's base coordinates (dims 1..4), the N offset, and whether my half exists (odd tile count)
    auto decode = [&](int w, int (&base)[4], int& n0) {
        n0 = (w % n_tiles) * kN;
        int t = (w / n_tiles) * 2 + static_cast<int>(rank);
        const bool valid = t < m_tiles;
        if (!valid) t = 0;                              // the idle half of the last pair recomputes tile 0 and stores nothing
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            base[i] = (t % p.tiles[i]) * p.box[i];
            t /= p.tiles[i];
        }
        return valid;
    };

    pdl_launch_dependents();
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.a_map[0]);
        tma_prefetch_desc(&p.a_map[1]);
        tma_prefetch_desc(&p.b_map);
        tma_prefetch_desc(&p.c_map);
        for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 2); }
        mbar_fence_init();
    }
    if (warp == 1) {
        tmem_alloc_pair(tmem_slot, 512);
        tmem_relinquish_pair();
    }
    tc_fence_before();
    cluster_sync_all();                               // both CTAs' barriers exist before either signals the other
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();

    if (warp == 0) {
        if (lane == 0) {
            int g = 0;                                // K blocks issued so far (ring position carries over from tile to tile)
            bool ok = true;
            for (int w = cluster_id; w < work_total && ok; w += n_clusters) {
                int base[4], n0;
                decode(w, base, n0);
                int kb = 0;
                for (int t = 0; t < p.n_taps && ok; ++t) {
                    const CUtensorMap* amap = &p.a_map[p.tap_src[t]];
                    const int c1 = base[0] + p.tap_off[t][0], c2 = base[1] + p.tap_off[t][1];
                    const int c3 = base[2] + p.tap_off[t][2], c4 = base[3] + p.tap_off[t][3];
                    for (int ch = 0; ch < p.tap_chunks[t]; ++ch, ++kb, ++g) {
                        const int s = g % kStages;
                        if (!mbar_wait(&empty_bar[s], ((g / kStages) & 1) ^ 1, err_flag, 11)) { ok = false; break; }
                        if (leader) mbar_expect_tx(&full_bar[s], 2 * S::kStageBytes);       // both CTAs' bytes
                        const uint32_t full_addr = cluster_map_shared(smem_u32(&full_bar[s]), 0);
                        unsigned char* stage = smem + s * S::kStageBytes;
                        tma_load_5d_pair(stage, amap, full_addr, ch * kBlockK, c1, c2, c3, c4);
                        tma_load_5d_pair(stage + kABytes, &p.b_map, full_addr, kb * kBlockK, n0 + static_cast<int>(rank) * 128, 0, 0, 0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && leader) {
            const uint32_t idesc = umma_idesc_bf16(256, kN);
            int g = 0, j = 0;
            bool ok = true;
            for (int w = cluster_id; w < work_total && ok; w += n_clusters, ++j) {
                const int buf = j & 1;
                if (!mbar_wait(&acc_empty[buf], ((j >> 1) & 1) ^ 1, err_flag, 14)) break;     // both epilogues drained it
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(buf * kN);
                for (int kb = 0; kb < p.k_blocks; ++kb, ++g) {
                    const int s = g % kStages;
                    if (!mbar_wait(&full_bar[s], (g / kStages) & 1, err_flag, 12)) { ok = false; break; }
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(smem + s * S::kStageBytes);
                    const uint32_t b_addr = a_addr + kABytes;
#pragma unroll
                    for (int k = 0; k < kBlockK / 16; ++k)
                        umma_bf16_pair(d_tmem, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(b_addr + k * 32), idesc, (kb | k) != 0);
                    umma_commit_pair(&empty_bar[s], 0x3);
                }
                if (ok) umma_commit_pair(&acc_full[buf], 0x3);
            }
        }
    } else {
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        float* s_bias = reinterpret_cast<float*>(smem + S::kBiasOffset);
        float* s_gamma = s_bias + kN;
        float* s_beta = s_gamma + kN;
        const bool ln = p.ln_gamma != nullptr;          // LayerNorm over the row's 256 channels (N = 256: the whole row is mine)
        unsigned char* staging = smem + S::kStagingOffset;
        const uint32_t acc_empty_leader0 = cluster_map_shared(smem_u32(&acc_empty[0]), 0);
        int j = 0, chunk_no = 0;
        for (int w = cluster_id; w < work_total; w += n_clusters, ++j) {
            int base[4], n0;
            const bool tile_valid = decode(w, base, n0);
            const int buf = j & 1;
            for (int i = threadIdx.x - 64; i < kN; i += 128) {
                s_bias[i] = bias != nullptr ? __ldg(bias + n0 + i) : 0.f;
                if (ln) { s_gamma[i] = __ldg(p.ln_gamma + i); s_beta[i] = __ldg(p.ln_beta + i); }
            }
            named_barrier(2, 128);
            if (!mbar_wait(&acc_full[buf], (j >> 1) & 1, err_flag, 13)) break;
            tc_fence_after();
            float mean = 0.f, rstd = 1.f;
            if (ln) ln256_row_stats(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + buf * kN, s_bias, mean, rstd);
#pragma unroll 1
            for (int c0 = 0; c0 < kN; c0 += 64, ++chunk_no) {
                unsigned char* st = staging + (chunk_no & 1) * 16384;
                if (warp == 2 && lane == 0) tma_store_wait_read_keep1();       // the store issued from this buffer two chunks ago
                named_barrier(1, 128);
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    uint32_t v[32];
                    tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + buf * kN + c0 + hf * 32, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int jj = 0; jj < 32; jj += 8) {
                        float f[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            float x = __uint_as_float(v[jj + e]) + s_bias[c0 + hf * 32 + jj + e];
                            if (ln) x = (x - mean) * rstd * s_gamma[c0 + hf * 32 + jj + e] + s_beta[c0 + hf * 32 + jj + e];
                            if (p.act == kActLeaky) x = x > 0.f ? x : kLeakySlope * x;
                            else if (p.act == kActRelu) x = fmaxf(x, 0.f);
                            f[e] = x;
                        }
                        uint4 q;
                        __nv_bfloat162 h0 = __floats2bfloat162_rn(f[0], f[1]);
                        __nv_bfloat162 h1 = __floats2bfloat162_rn(f[2], f[3]);
                        __nv_bfloat162 h2 = __floats2bfloat162_rn(f[4], f[5]);
                        __nv_bfloat162 h3 = __floats2bfloat162_rn(f[6], f[7]);
                        q.x = *reinterpret_cast<uint32_t*>(&h0); q.y = *reinterpret_cast<uint32_t*>(&h1);
                        q.z = *reinterpret_cast<uint32_t*>(&h2); q.w = *reinterpret_cast<uint32_t*>(&h3);
                        const int chunk = hf * 4 + (jj >> 3);
                        *reinterpret_cast<uint4*>(st + row * 128 + ((chunk ^ (row & 7)) << 4)) = q;
                    }
                }
                fence_proxy_async_smem();
                tc_fence_before();
                named_barrier(1, 128);
                if (warp == 2 && lane == 0) {
                    if (tile_valid) {
                        tma_store_5d(&p.c_map, st, n0 + c0, base[0], base[1], base[2], base[3]);
                        tma_store_commit();
                    }
                    // after the last chunk every epilogue thread of this CTA has read its accumulator rows: hand the buffer back
                    if (c0 + 64 == kN) mbar_arrive_cluster(acc_empty_leader0 + buf * 8);
                }
            }
        }
        if (warp == 2 && lane == 0) tma_store_wait_read();
    }
    tc_fence_before();
    cluster_sync_all();                               // neither CTA leaves (or frees tensor memory) while the other can still signal it
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, 512);
    }
}

// The same CTA-pair tile for layers whose tile list fits one round: one tile per cluster, single accumulator buffer, six
// ring stages, and the drained ring itself is the epilogue's staging area (measured 2 us per launch faster on the
// decoder layers than running the persistent kernel for a single round).
constexpr int kPairSingleSmem = 6 * 2 * kABytes + 128 + 3 * 256 * 4 + 1024;      // ring, barriers, bias | LN gamma | LN beta, slack
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads)
conv_gemm_pair_single_kernel(const __grid_constant__ ConvGemmParams p, const float* __restrict__ bias, int* __restrict__ err_flag) {
    constexpr int kStages = 6;
    constexpr int kStageBytes = 2 * kABytes;
    constexpr int kBarOffset = kStages * kStageBytes;
    constexpr int kBiasOffset = kBarOffset + 128;
    constexpr int kN = 256;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    unsigned char* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kBarOffset);
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* accum_bar = empty_bar + kStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int n_tiles = p.N / kN;
    int base[4];
    bool tile_valid;
    int n0;
    {
        const int w = blockIdx.x >> 1;
        n0 = (w % n_tiles) * kN;
        int t = (w / n_tiles) * 2 + static_cast<int>(rank);
        const int m_tiles = p.tiles[0] * p.tiles[1] * p.tiles[2] * p.tiles[3];
        tile_valid = t < m_tiles;
        if (!tile_valid) t = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            base[i] = (t % p.tiles[i]) * p.box[i];
            t /= p.tiles[i];
        }
    }
    pdl_launch_dependents();
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.a_map[0]);
        tma_prefetch_desc(&p.a_map[1]);
        tma_prefetch_desc(&p.b_map);
        for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(accum_bar, 1);
        mbar_fence_init();
    }
    if (warp == 1) {
        tmem_alloc_pair(tmem_slot, kN);
        tmem_relinquish_pair();
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();
    if (warp == 0) {
        if (lane == 0) {
            int kb = 0;
            bool ok = true;
            for (int t = 0; t < p.n_taps && ok; ++t) {
                const CUtensorMap* amap = &p.a_map[p.tap_src[t]];
                const int c1 = base[0] + p.tap_off[t][0], c2 = base[1] + p.tap_off[t][1];
                const int c3 = base[2] + p.tap_off[t][2], c4 = base[3] + p.tap_off[t][3];
                for (int ch = 0; ch < p.tap_chunks[t]; ++ch, ++kb) {
                    const int s = kb % kStages;
                    if (!mbar_wait(&empty_bar[s], ((kb / kStages) & 1) ^ 1, err_flag, 11)) { ok = false; break; }
                    if (leader) mbar_expect_tx(&full_bar[s], 2 * kStageBytes);
                    const uint32_t full_addr = cluster_map_shared(smem_u32(&full_bar[s]), 0);
                    unsigned char* stage = smem + s * kStageBytes;
                    tma_load_5d_pair(stage, amap, full_addr, ch * kBlockK, c1, c2, c3, c4);
                    tma_load_5d_pair(stage + kABytes, &p.b_map, full_addr, kb * kBlockK, n0 + static_cast<int>(rank) * 128, 0, 0, 0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && leader) {
            const uint32_t idesc = umma_idesc_bf16(256, kN);
            for (int kb = 0; kb < p.k_blocks; ++kb) {
                const int s = kb % kStages;
                if (!mbar_wait(&full_bar[s], (kb / kStages) & 1, err_flag, 12)) break;
                tc_fence_after();
                const uint32_t a_addr = smem_u32(smem + s * kStageBytes);
                const uint32_t b_addr = a_addr + kABytes;
#pragma unroll
                for (int k = 0; k < kBlockK / 16; ++k)
                    umma_bf16_pair(tmem_base, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(b_addr + k * 32), idesc, (kb | k) != 0);
                umma_commit_pair(&empty_bar[s], 0x3);
            }
            umma_commit_pair(accum_bar, 0x3);
        }
    } else {
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        float* s_bias = reinterpret_cast<float*>(smem + kBiasOffset);
        float* s_gamma = s_bias + kN;
        float* s_beta = s_gamma + kN;
        const bool ln = p.ln_gamma != nullptr;          // LayerNorm over the row's 256 channels (one N tile: the whole row is mine)
        for (int i = threadIdx.x - 64; i < kN; i += 128) {
            s_bias[i] = (bias != nullptr && n0 + i < p.N) ? __ldg(bias + n0 + i) : 0.f;
            if (ln) { s_gamma[i] = __ldg(p.ln_gamma + i); s_beta[i] = __ldg(p.ln_beta + i); }
        }
        named_barrier(2, 128);
        mbar_wait(accum_bar, 0, err_flag, 13);
        tc_fence_after();
        tma_prefetch_desc(&p.c_map);
        float mean = 0.f, rstd = 1.f;
        if (ln) ln256_row_stats(tmem_base + (static_cast<uint32_t>(quad * 32) << 16), s_bias, mean, rstd);
#pragma unroll 1
        for (int c0 = 0; c0 < kN; c0 += 64) {
            unsigned char* st = smem + (c0 >> 6) * 16384;
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                uint32_t v[32];
                tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + c0 + hf * 32, v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                    float f[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        float x = __uint_as_float(v[j + e]) + s_bias[c0 + hf * 32 + j + e];
                        if (ln) x = (x - mean) * rstd * s_gamma[c0 + hf * 32 + j + e] + s_beta[c0 + hf * 32 + j + e];
                        if (p.act == kActLeaky) x = x > 0.f ? x : kLeakySlope * x;
                        else if (p.act == kActRelu) x = fmaxf(x, 0.f);
                        f[e] = x;
                    }
                    uint4 q;
                    __nv_bfloat162 h0 = __floats2bfloat162_rn(f[0], f[1]);
                    __nv_bfloat162 h1 = __floats2bfloat162_rn(f[2], f[3]);
                    __nv_bfloat162 h2 = __floats2bfloat162_rn(f[4], f[5]);
                    __nv_bfloat162 h3 = __floats2bfloat162_rn(f[6], f[7]);
                    q.x = *reinterpret_cast<uint32_t*>(&h0); q.y = *reinterpret_cast<uint32_t*>(&h1);
                    q.z = *reinterpret_cast<uint32_t*>(&h2); q.w = *reinterpret_cast<uint32_t*>(&h3);
                    const int chunk = hf * 4 + (j >> 3);
                    *reinterpret_cast<uint4*>(st + row * 128 + ((chunk ^ (row & 7)) << 4)) = q;
                }
            }
            fence_proxy_async_smem();
            named_barrier(1, 128);
            if (warp == 2 && lane == 0 && tile_valid) {
                tma_store_5d(&p.c_map, st, n0 + c0, base[0], base[1], base[2], base[3]);
                tma_store_commit();
            }
        }
        if (warp == 2 && lane == 0) tma_store_wait_read();
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, kN);
    }
}

// ---------------------------------------------------------------------------------------------
// weight packing / BatchNorm folding (run once at model load)
// ---------------------------------------------------------------------------------------------
struct PackTaps {
    int n_taps;
    int k_start[kMaxTaps + 1];
    long long w_off[kMaxTaps];
};

__global__ void pack_weights_kernel(const float* __restrict__ w, long long sn, long long sc, PackTaps pt, int N,
                                    long long K, const float* __restrict__ scale, __nv_bfloat16* __restrict__ dst,
                                    long long dst_ld, long long dst_col) {
    const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= static_cast<long long>(N) * K) return;
    const int n = static_cast<int>(idx / K);
    const int k = static_cast<int>(idx - static_cast<long long>(n) * K);
    int t = 0;
    while (t + 1 < pt.n_taps && k >= pt.k_start[t + 1]) ++t;
    const int c = k - pt.k_start[t];
    float v = w[n * sn + pt.w_off[t] + c * sc];
    if (scale) v *= scale[n];
    dst[n * dst_ld + dst_col + k] = __float2bfloat16_rn(v);
}

__global__ void fold_bn_kernel(const float* __restrict__ conv_bias, const float* __restrict__ gamma,
                               const float* __restrict__ beta, const float* __restrict__ mean,
                               const float* __restrict__ var, float eps, int N, float* __restrict__ scale_out,
                               float* __restrict__ bias_out) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float s = gamma[n] / sqrtf(var[n] + eps);
    scale_out[n] = s;
    bias_out[n] = ((conv_bias ? conv_bias[n] : 0.f) - mean[n]) * s + beta[n];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

int make_map(CUtensorMap* map, const void* ptr, int rank, const long long* dims, const long long* strides,
             const int* box, CUtensorMapL2promotion promo, const char* what) {
    EncodeTiledFn enc = encode_tiled();
    if (!enc) { a2m_set_error("cuTensorMapEncodeTiled is not available from the driver"); return A2M_ERR_STATE; }
    cuuint64_t gdim[5], gstr[4];
    cuuint32_t bdim[5], estr[5] = {1, 1, 1, 1, 1};
    for (int i = 0; i < 5; ++i) {
        gdim[i] = i < rank ? static_cast<cuuint64_t>(dims[i]) : 1;
        bdim[i] = static_cast<cuuint32_t>(box[i]);
    }
    long long last = 0;
    for (int i = 1; i < 5; ++i) {
        long long s = i < rank ? strides[i] * 2 : last;          // bytes
        if (i >= rank && s == 0) s = 16;
        if (s % 16 != 0 || s <= 0) { a2m_set_error("%s: stride of dim %d (%lld B) is not a positive multiple of 16", what, i, s); return A2M_ERR_ARGUMENT; }
        gstr[i - 1] = static_cast<cuuint64_t>(s);
        last = s * static_cast<long long>(gdim[i]);
    }
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0) { a2m_set_error("%s: base pointer not 16-byte aligned", what); return A2M_ERR_ARGUMENT; }
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(ptr), gdim, gstr, bdim, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, promo,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        a2m_set_error("%s: cuTensorMapEncodeTiled failed (CUresult %d) dims=[%llu,%llu,%llu,%llu,%llu] box=[%u,%u,%u,%u,%u]",
                      what, (int)r, (unsigned long long)gdim[0], (unsigned long long)gdim[1], (unsigned long long)gdim[2],
                      (unsigned long long)gdim[3], (unsigned long long)gdim[4], bdim[0], bdim[1], bdim[2], bdim[3], bdim[4]);
        return A2M_ERR_ARGUMENT;
    }
    return A2M_OK;
}

template <int BLOCK_N, int STAGES>
int launch_variant(const ConvGemmPlan& plan, int* err_flag, cudaStream_t stream) {
    static A2mPerDeviceOnce configured;
    if (configured.first()) {
        A2M_CUDA_CHECK(cudaFuncSetAttribute(conv_gemm_kernel<BLOCK_N, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            Smem<BLOCK_N, STAGES>::kTotal));
    }
    A2M_CUDA_CHECK(a2m_launch_pdl(conv_gemm_kernel<BLOCK_N, STAGES>, plan.grid, dim3(kThreads), Smem<BLOCK_N, STAGES>::kTotal,
                                  stream, plan.p, plan.bias, plan.out, err_flag));
    a2m_count_launch();
    return A2M_OK;
}

}  // namespace

// generic bf16 view (<= 5-D, dim 0 = unit-stride channels) as a 128B-swizzled TMA map
int make_map_bf16(CUtensorMap* map, const void* ptr, int rank, const long long* dims, const long long* strides,
                  const int* box, const char* what) {
    return make_map(map, ptr, rank, dims, strides, box, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, what);
}

// 2-D bf16 weight matrix [n_rows, k] (K-major) as a 128B-swizzled TMA map with box (64, box_rows)
int make_weight_map(CUtensorMap* map, const void* w, long long n_rows, long long k, int box_rows) {
    long long wd[2] = {k, n_rows}, ws[2] = {1, k};
    int wb[5] = {kBlockK, box_rows, 1, 1, 1};
    return make_map(map, w, 2, wd, ws, wb, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, "weight map");
}

long long conv_gemm_k(const ConvGemmDesc& d) {
    long long k = 0;
    for (const Tap& t : d.taps) k += t.channels;
    return k;
}

int pack_weights(const float* w_src, long long w_stride_n, long long w_stride_c, const std::vector<Tap>& taps, int N,
                 const float* scale, __nv_bfloat16* w_packed, cudaStream_t stream, long long dst_ld, long long dst_col) {
    A2M_ARG_CHECK(!taps.empty() && static_cast<int>(taps.size()) <= kMaxTaps, "pack_weights: %zu taps", taps.size());
    PackTaps pt;
    pt.n_taps = static_cast<int>(taps.size());
    int k = 0;
    for (int t = 0; t < pt.n_taps; ++t) { pt.k_start[t] = k; pt.w_off[t] = taps[t].w_off; k += taps[t].channels; }
    pt.k_start[pt.n_taps] = k;
    const long long total = static_cast<long long>(N) * k;
    pack_weights_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(w_src, w_stride_n, w_stride_c,
                                                                                         pt, N, k, scale, w_packed,
                                                                                         dst_ld > 0 ? dst_ld : k, dst_col);
    a2m_count_launch();
    A2M_LAUNCH_CHECK();
    return A2M_OK;
}

int fold_batchnorm(const float* conv_bias, const float* gamma, const float* beta, const float* mean, const float* var,
                   float eps, int N, float* scale_out, float* bias_out, cudaStream_t stream) {
    fold_bn_kernel<<<(N + 127) / 128, 128, 0, stream>>>(conv_bias, gamma, beta, mean, var, eps, N, scale_out, bias_out);
    a2m_count_launch();
    A2M_LAUNCH_CHECK();
    return A2M_OK;
}

int conv_gemm_plan(const ConvGemmDesc& d, const void* w_packed, const float* bias, void* out, ConvGemmPlan* plan) {
    A2M_ARG_CHECK(plan != nullptr, "conv_gemm_plan: NULL plan");
    A2M_ARG_CHECK(d.n_src == 1 || d.n_src == 2, "conv_gemm_plan: n_src %d", d.n_src);
    A2M_ARG_CHECK(!d.taps.empty() && static_cast<int>(d.taps.size()) <= kMaxTaps, "conv_gemm_plan: %zu taps (max %d)",
                  d.taps.size(), kMaxTaps);
    A2M_ARG_CHECK(d.N >= 1, "conv_gemm_plan: N %d", d.N);
    A2M_ARG_CHECK(d.out_type == kOutF32 || d.N % 8 == 0, "conv_gemm_plan: bf16 output needs N %% 8 == 0 (N = %d)", d.N);
    long long rows = 1;
    for (int i = 0; i < 4; ++i) {
        A2M_ARG_CHECK(d.box[i] >= 1 && d.box[i] <= 256 && d.m_extent[i] >= 1, "conv_gemm_plan: box/extent of dim %d", i + 1);
        rows *= d.box[i];
    }
    A2M_ARG_CHECK(rows == kBlockM, "conv_gemm_plan: tile has %lld rows, must be %d", rows, kBlockM);

    ConvGemmParams& p = plan->p;
    memset(&p, 0, sizeof(p));
    int box5[5] = {kBlockK, d.box[0], d.box[1], d.box[2], d.box[3]};
    for (int s = 0; s < 2; ++s) {
        const AView& v = d.a[s < d.n_src ? s : 0];
        A2M_ARG_CHECK(v.ptr != nullptr && v.rank >= 2 && v.rank <= 5 && v.strides[0] == 1 && v.dims[0] % kBlockK == 0,
                      "conv_gemm_plan: bad A view %d (rank %d, channels %lld)", s, v.rank, v.dims[0]);
        const int rc = make_map(&p.a_map[s], v.ptr, v.rank, v.dims, v.strides, box5, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                "activation map");
        if (rc != A2M_OK) return rc;
    }
    const long long K = conv_gemm_k(d);
    int block_n = 128;
    if (d.N <= 32) block_n = 32;
    else if (d.N <= 64) block_n = 64;
    else if (d.block_n_hint >= 256 && d.N % 256 == 0 && d.out_type == kOutBf16 && d.split_k == 1) block_n = 256;
    {
        long long wd[2] = {K, d.N}, ws[2] = {1, K};
        int wb[5] = {kBlockK, block_n > 128 ? 128 : block_n, 1, 1, 1};
        const int rc = make_map(&p.b_map, w_packed, 2, wd, ws, wb, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, "weight map");
        if (rc != A2M_OK) return rc;
    }
    p.c_tma = 0;
    if (d.out_type == kOutBf16 && block_n >= 64) {
        // output as a strided 5-D view: dim 0 = channels, dims 1..4 = the M coordinates
        long long cd[5], cs[5];
        cd[0] = d.N; cs[0] = 1;
        bool ok = (d.out_base % 8) == 0;
        for (int i = 0; i < 4; ++i) {
            cd[i + 1] = d.m_extent[i];
            cs[i + 1] = d.m_extent[i] > 1 ? d.out_stride[i] : 8;      // extent-1 dims: any legal stride
            ok = ok && cs[i + 1] > 0 && (cs[i + 1] % 8) == 0;
        }
        if (ok) {
            int cb[5] = {64, d.box[0], d.box[1], d.box[2], d.box[3]};
            const int rc = make_map(&p.c_map, static_cast<__nv_bfloat16*>(out) + d.out_base, 5, cd, cs, cb,
                                    CU_TENSOR_MAP_L2_PROMOTION_NONE, "output map");
            if (rc != A2M_OK) return rc;
            p.c_tma = 1;
        }
    }
    p.n_taps = static_cast<int>(d.taps.size());
    int kb = 0;
    for (int t = 0; t < p.n_taps; ++t) {
        const Tap& tp = d.taps[t];
        A2M_ARG_CHECK(tp.src >= 0 && tp.src < d.n_src && tp.channels > 0 && tp.channels % kBlockK == 0 &&
                          tp.channels <= d.a[tp.src].dims[0],
                      "conv_gemm_plan: tap %d (src %d, channels %d)", t, tp.src, tp.channels);
        p.tap_src[t] = tp.src;
        p.tap_chunks[t] = tp.channels / kBlockK;
        for (int i = 0; i < 4; ++i) p.tap_off[t][i] = tp.off[i];
        kb += p.tap_chunks[t];
    }
    p.k_blocks = kb;
    long long m_tiles = 1, m_valid = 1;
    for (int i = 0; i < 4; ++i) {
        p.box[i] = d.box[i];
        p.m_extent[i] = d.m_extent[i];
        p.tiles[i] = (d.m_extent[i] + d.box[i] - 1) / d.box[i];
        p.out_stride[i] = d.out_stride[i];
        m_tiles *= p.tiles[i];
        m_valid *= d.m_extent[i];
    }
    p.out_base = d.out_base;
    p.N = d.N;
    p.act = d.act;
    p.out_type = d.out_type;
    A2M_ARG_CHECK(d.split_k >= 1 && d.split_k <= 64 && d.split_k <= kb, "conv_gemm_plan: split_k %d (k blocks %d)", d.split_k, kb);
    A2M_ARG_CHECK(d.split_k == 1 || (d.out_type == kOutF32 && d.act == kActNone),
                  "conv_gemm_plan: split-K needs an fp32 output without activation");
    p.split_k = d.split_k;
    p.split_stride = d.split_stride;
    A2M_ARG_CHECK(m_tiles <= 0x7fffffffLL, "conv_gemm_plan: too many M tiles");
    plan->w_packed = w_packed;
    plan->bias = bias;
    plan->out = out;
    plan->block_n = block_n;
    {   // short K loops (decoder layers): two 32 KB stages let a third CTA -- typically the next layer's, launched
        // programmatically -- become resident while this layer's CTAs drain
        plan->stages = (block_n == 128 && kb <= 16) ? 2 : 3;
    }
    plan->grid = dim3(static_cast<unsigned>(m_tiles), static_cast<unsigned>((d.N + block_n - 1) / block_n),
                      static_cast<unsigned>(d.split_k));
    const bool want_ln = d.ln_gamma != nullptr && d.ln_beta != nullptr && d.N == 256;
    // (a layer with a fused LayerNorm takes the pair kernel even for a single M tile -- half the pair idles -- so that a
    // row's arithmetic is the same for every batch size)
    plan->pair = (d.block_n_hint == 512 && block_n == 256 && p.c_tma && (m_tiles >= 2 || want_ln)) ? 1 : 0;
    if (plan->pair) {                                   // persistent: one cluster (CTA pair) per SM pair, or fewer
        const long long work = (m_tiles + 1) / 2 * (d.N / 256);
        const long long slots = a2m_num_sms() / 2;
        // the tile list takes ceil(work / slots) rounds whatever the grid: launch only as many clusters as fill those
        // rounds evenly (128 tiles -> 64 clusters x 2 rounds, not 74) and leave the other SMs to the kernels of the
        // second stream lane
        const long long rounds = (work + slots - 1) / slots;
        const long long clusters = (work + rounds - 1) / rounds;
        plan->grid = dim3(static_cast<unsigned>(2 * clusters), 1, 1);
        if (rounds == 1) plan->pair = 2;                // one tile per cluster
    }
    plan->ln_fused = (want_ln && plan->pair != 0) ? 1 : 0;
    p.ln_gamma = plan->ln_fused ? d.ln_gamma : nullptr;
    p.ln_beta = plan->ln_fused ? d.ln_beta : nullptr;
    plan->flops = 2 * m_valid * d.N * K;
    return A2M_OK;
}

static int launch_pair(const ConvGemmPlan& plan, int* err_flag, cudaStream_t stream) {
    if (plan.pair == 2) {
        static A2mPerDeviceOnce configured;
        if (configured.first())
            A2M_CUDA_CHECK(cudaFuncSetAttribute(conv_gemm_pair_single_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPairSingleSmem));
        A2M_CUDA_CHECK(a2m_launch_pdl(conv_gemm_pair_single_kernel, plan.grid, dim3(kThreads), kPairSingleSmem, stream, plan.p,
                                      plan.bias, err_flag));
    } else {
        static A2mPerDeviceOnce configured;
        if (configured.first())
            A2M_CUDA_CHECK(cudaFuncSetAttribute(conv_gemm_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SmemPair::kTotal));
        A2M_CUDA_CHECK(a2m_launch_pdl(conv_gemm_pair_kernel, plan.grid, dim3(kThreads), SmemPair::kTotal, stream, plan.p, plan.bias,
                                      err_flag));
    }
    a2m_count_launch();
    return A2M_OK;
}

int conv_gemm_launch(const ConvGemmPlan& plan, int* err_flag, cudaStream_t stream) {
    if (plan.pair) return launch_pair(plan, err_flag, stream);
    switch (plan.block_n) {
        case 32: return launch_variant<32, 3>(plan, err_flag, stream);
        case 64: return launch_variant<64, 3>(plan, err_flag, stream);
        case 128: return plan.stages == 2 ? launch_variant<128, 2>(plan, err_flag, stream) : launch_variant<128, 3>(plan, err_flag, stream);
        case 256: return launch_variant<256, 4>(plan, err_flag, stream);
        default: a2m_set_error("conv_gemm_launch: block_n %d", plan.block_n); return A2M_ERR_STATE;
    }
}

}  // namespace a2m

// ---------------------------------------------------------------------------------------------
// C ABI: the generic operator, used directly by the unit tests (one layer vs the oracle's conv)
// ---------------------------------------------------------------------------------------------
extern "C" int a2m_gemm_taps(const a2m_gemm_desc* g, const float* w_src, int64_t w_stride_n, int64_t w_stride_c,
                             const float* scale, const float* bias, void* out, void* stream) {
    using namespace a2m;
    A2M_ARG_CHECK(g != nullptr && w_src != nullptr && out != nullptr, "a2m_gemm_taps: NULL argument");
    A2M_ARG_CHECK(g->n_taps >= 1 && g->n_taps <= kMaxTaps, "a2m_gemm_taps: n_taps %d (max %d)", g->n_taps, kMaxTaps);
    ConvGemmDesc d;
    d.n_src = g->n_src;
    for (int s = 0; s < 2; ++s) {
        d.a[s].ptr = g->a_ptr[s];
        d.a[s].rank = g->a_rank[s];
        for (int i = 0; i < 5; ++i) { d.a[s].dims[i] = g->a_dims[s][i]; d.a[s].strides[i] = g->a_strides[s][i]; }
    }
    for (int i = 0; i < 4; ++i) {
        d.box[i] = g->box[i]; d.m_extent[i] = g->m_extent[i]; d.out_stride[i] = g->out_stride[i];
    }
    for (int t = 0; t < g->n_taps; ++t) {
        Tap tp;
        tp.src = g->tap_src[t];
        for (int i = 0; i < 4; ++i) tp.off[i] = g->tap_off[t][i];
        tp.channels = g->tap_channels[t];
        tp.w_off = g->tap_w_off[t];
        d.taps.push_back(tp);
    }
    d.N = g->N; d.out_base = g->out_base; d.act = g->act; d.out_type = g->out_type;
    d.block_n_hint = g->tile_hint;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const long long K = conv_gemm_k(d);
    __nv_bfloat16* wp = nullptr;
    int* flag = nullptr;
    A2M_CUDA_CHECK(cudaMalloc(&wp, static_cast<size_t>(d.N) * K * 2));
    cudaError_t e = cudaMalloc(&flag, sizeof(int));
    if (e != cudaSuccess) { cudaFree(wp); a2m_set_error("cudaMalloc failed"); return (int)e; }
    cudaMemsetAsync(flag, 0, sizeof(int), s);
    int rc = pack_weights(w_src, w_stride_n, w_stride_c, d.taps, d.N, scale, wp, s);
    ConvGemmPlan plan;
    if (rc == A2M_OK) rc = conv_gemm_plan(d, wp, bias, out, &plan);
    if (rc == A2M_OK) rc = conv_gemm_launch(plan, flag, s);
    int host_flag = 0;
    e = cudaStreamSynchronize(s);
    if (e == cudaSuccess) e = cudaMemcpy(&host_flag, flag, sizeof(int), cudaMemcpyDeviceToHost);
    cudaFree(wp);
    cudaFree(flag);
    if (rc != A2M_OK) return rc;
    if (e != cudaSuccess) { a2m_set_error("a2m_gemm_taps: %s", cudaGetErrorString(e)); return (int)e; }
    if (host_flag != 0) {
        a2m_set_error("a2m_gemm_taps: pipeline barrier wait expired (role %d)", host_flag);
        return A2M_ERR_PIPELINE;
    }
    return A2M_OK;
}
