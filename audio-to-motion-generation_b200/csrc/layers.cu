// Non-GEMM kernels of the generator forward: first encoder conv, time interpolation, self-attention,
// channel attention, LayerNorm, the static-graph GAT / GraphConv tails and the pose losses.
// See layers.cuh for the contracts and the reference lines each kernel restates.
#include "layers.cuh"
#include "fft_math.cuh"

void a2m_count_launch();

namespace a2m {

namespace {

__device__ __forceinline__ float leaky(float x) { return x > 0.f ? x : kLeakySlope * x; }

__device__ __forceinline__ void bf16x8_to_float(const uint4& q, float (&f)[8]) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 t = __bfloat1622float2(h[i]);
        f[2 * i] = t.x;
        f[2 * i + 1] = t.y;
    }
}
__device__ __forceinline__ uint4 float_to_bf16x8(const float (&f)[8]) {
    uint4 q;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&q);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    return q;
}

// ------------------------------------------------------------------------------------------ conv0
// One CTA = kConv0Rows output rows of one clip: the 2R+2 input rows are staged (zero padded) in shared
// memory.  A warp computes four adjacent output columns of one output row at a time: its 32 lanes own the
// 32 channel pairs (so every store instruction writes one whole 128-byte channel row) and all lanes read the
// same 4 x 10 input window as broadcast vector loads (12 loads for 8 x 16 FMAs per lane).
constexpr int kConv0Rows = 8;
__global__ void __launch_bounds__(256)
conv0_kernel(const float* __restrict__ mel, long long stride_b, long long stride_t, int n_inner, long long stride_outer,
             int T, int F, const float* __restrict__ w, const float* __restrict__ bias, __nv_bfloat16* __restrict__ out) {
    extern __shared__ __align__(16) float s_rows[];   // [2R + 2][stride]: column c of the input at index c + 1
    const int ho0 = blockIdx.x * kConv0Rows;
    const long long b = blockIdx.y;
    const int Ho = T / 2, Wo = F / 2;
    const int stride = ((F + 2 + 3) / 4) * 4 + 4;     // multiple of 4 floats, room for the 10-wide window of the last group
    const int n_in = 2 * kConv0Rows + 2;
    // clip b = window (b % n_inner) of stream (b / n_inner): sliding windows over one log-mel per stream are read in place
    const float* clip = mel + (b / n_inner) * stride_outer + (b % n_inner) * stride_b;
    for (int i = threadIdx.x; i < n_in * stride; i += blockDim.x) {
        const int r = i / stride, col = i - r * stride - 1;
        const int h = 2 * ho0 + r - 1;
        float v = 0.f;
        if (h >= 0 && h < T && col >= 0 && col < F) v = __ldg(clip + h * stride_t + col);
        s_rows[i] = v;
    }
    const int cp = threadIdx.x & 31, warp = threadIdx.x >> 5;
    using a2m_fft::pair_t;                            // my two channels ride one packed register: FFMA2, half the issue slots
    pair_t wp[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {                    // w is tap-major [16][64]: one coalesced 256-byte row per tap
        const float2 ww = __ldg(reinterpret_cast<const float2*>(w + i * 64) + cp);
        wp[i] = a2m_fft::pack(ww.x, ww.y);
    }
    const pair_t bp = a2m_fft::pack(__ldg(bias + 2 * cp), __ldg(bias + 2 * cp + 1));
    __syncthreads();
    const int groups = (Wo + 3) / 4;
    for (int item = warp; item < kConv0Rows * groups; item += 8) {
        const int r = item / groups, wo0 = (item - r * groups) * 4;
        const int ho = ho0 + r;
        if (ho >= Ho) break;
        pair_t acc[4] = {bp, bp, bp, bp};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float* row = s_rows + (2 * r + i) * stride + 2 * wo0;       // 16-byte aligned: stride and 2*wo0 are multiples of 4
            const float4 x0 = *reinterpret_cast<const float4*>(row);
            const float4 x1 = *reinterpret_cast<const float4*>(row + 4);
            const float2 x2 = *reinterpret_cast<const float2*>(row + 8);
            const float xs[10] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w, x2.x, x2.y};
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[u] = a2m_fft::fma2(a2m_fft::bcast(xs[2 * u + j]), wp[i * 4 + j], acc[u]);
        }
        __nv_bfloat16* o = out + ((b * Ho + ho) * Wo + wo0) * 64 + 2 * cp;
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (wo0 + u < Wo)
                *reinterpret_cast<__nv_bfloat162*>(o + u * 64) =
                    __floats2bfloat162_rn(leaky(a2m_fft::lo(acc[u])), leaky(a2m_fft::hi(acc[u])));
    }
}

// ------------------------------------------------------------------------------------ time interp
// in: raw conv-4 sums (+ folded bias) as n_planes split-K planes of fp32 [B, Hc, C].  One thread owns 8 channels
// of one source row g of one clip and emits the T / Hc output rows that interpolate around it: it sums the planes
// of rows g-1, g, g+1 in a fixed order (deterministic), applies LeakyReLU (the split-K GEMM cannot), then the
// bilinear (T, 1) resize of model_layers.py:277 (align_corners=False: src = (t + 0.5) * Hc / T - 0.5, clamped at 0).
constexpr int kMaxPlanes = 8;
__global__ void __launch_bounds__(256)
time_interp_kernel(const float* __restrict__ in, int n_planes, long long plane_stride, int Hc, int T, int C,
                   long long total, __nv_bfloat16* __restrict__ out) {
    const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int c8 = C / 8;
    const int c = static_cast<int>(idx % c8) * 8;
    const long long bg = idx / c8;
    const int g = static_cast<int>(bg % Hc);
    const long long b = bg / Hc;
    const int rows[3] = {g > 0 ? g - 1 : 0, g, g < Hc - 1 ? g + 1 : Hc - 1};
    float v[3][8];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int e = 0; e < 8; ++e) v[r][e] = 0.f;
#pragma unroll
    for (int pl = 0; pl < kMaxPlanes; ++pl) {             // fixed summation order over the split-K planes
        if (pl < n_planes) {
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const float4* src = reinterpret_cast<const float4*>(in + pl * plane_stride + (b * Hc + rows[r]) * C + c);
                const float4 a0 = __ldg(src), a1 = __ldg(src + 1);
                v[r][0] += a0.x; v[r][1] += a0.y; v[r][2] += a0.z; v[r][3] += a0.w;
                v[r][4] += a1.x; v[r][5] += a1.y; v[r][6] += a1.z; v[r][7] += a1.w;
            }
        }
    }
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int e = 0; e < 8; ++e) v[r][e] = leaky(v[r][e]);
    const int rep = T / Hc;
    const float scale = static_cast<float>(Hc) / static_cast<float>(T);
    for (int u = 0; u < rep; ++u) {
        const int t = g * rep + u;
        float src = scale * (static_cast<float>(t) + 0.5f) - 0.5f;
        if (src < 0.f) src = 0.f;
        const int i0 = static_cast<int>(src);
        const int i1 = i0 + (i0 < Hc - 1 ? 1 : 0);
        const float l1 = src - static_cast<float>(i0), l0 = 1.f - l1;
        // i0, i1 are in {g-1, g, g+1} (clamped): pick the staged rows
        const int s0 = i0 < g ? 0 : (i0 == g ? 1 : 2), s1 = i1 < g ? 0 : (i1 == g ? 1 : 2);
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const float a = s0 == 0 ? v[0][e] : (s0 == 1 ? v[1][e] : v[2][e]);
            const float d = s1 == 0 ? v[0][e] : (s1 == 1 ? v[1][e] : v[2][e]);
            f[e] = l0 * a + l1 * d;
        }
        *reinterpret_cast<uint4*>(out + (b * T + t) * C + c) = float_to_bf16x8(f);
    }
}

// -------------------------------------------------------------------------------------- attention
// One CTA per clip.  q, k staged as fp32 in shared memory; S = q k^T register-tiled 4x4 per thread;
// softmax by one warp per row, written TRANSPOSED (P^T[j][t]) so that the P.V loop reads, for a fixed
// source step j, the weights of all T output steps as broadcast 16-byte loads; each thread owns one
// channel and keeps its T accumulators in registers (T FMAs per 1 global + T/4 shared loads).
template <int T>
__global__ void __launch_bounds__(256, 2)
attention_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ x,
                 const __nv_bfloat16* __restrict__ res2, const float* __restrict__ gamma_p, int C,
                 __nv_bfloat16* __restrict__ out) {
    extern __shared__ __align__(16) float s_attn[];
    const int d = C / 8, ld = 2 * d + C, qs = d + 4;              // row stride d+4: 16 B aligned, conflict-free
    float* s_q = s_attn;                       // [T][d+4]
    float* s_k = s_q + T * qs;                 // [T][d+4]
    float* s_p = s_k + T * qs;                 // S [T][T+1], then P^T [T][T]
    const long long b = blockIdx.x;
    const __nv_bfloat16* base = qkv + b * T * ld;
    for (int i = threadIdx.x; i < T * (d / 8); i += blockDim.x) {          // 8 bf16 per 16-byte load
        const int t = i / (d / 8), c8 = i - t * (d / 8);
        float f[8];
        bf16x8_to_float(*reinterpret_cast<const uint4*>(base + static_cast<long long>(t) * ld + c8 * 8), f);
        *reinterpret_cast<float4*>(s_q + t * qs + c8 * 8) = make_float4(f[0], f[1], f[2], f[3]);
        *reinterpret_cast<float4*>(s_q + t * qs + c8 * 8 + 4) = make_float4(f[4], f[5], f[6], f[7]);
        bf16x8_to_float(*reinterpret_cast<const uint4*>(base + static_cast<long long>(t) * ld + d + c8 * 8), f);
        *reinterpret_cast<float4*>(s_k + t * qs + c8 * 8) = make_float4(f[0], f[1], f[2], f[3]);
        *reinterpret_cast<float4*>(s_k + t * qs + c8 * 8 + 4) = make_float4(f[4], f[5], f[6], f[7]);
    }
    __syncthreads();
    constexpr int kTiles = (T / 4) * (T / 4);                     // 4x4 output tiles of S
    for (int tile = threadIdx.x; tile < kTiles; tile += blockDim.x) {
        const int i0 = (tile / (T / 4)) * 4, j0 = (tile % (T / 4)) * 4;
        float acc[4][4] = {};
        for (int c = 0; c < d; c += 4) {
            float4 qv[4], kv[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                qv[a] = *reinterpret_cast<const float4*>(s_q + (i0 + a) * qs + c);
                kv[a] = *reinterpret_cast<const float4*>(s_k + (j0 + a) * qs + c);
            }
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    acc[a][e] += qv[a].x * kv[e].x + qv[a].y * kv[e].y + qv[a].z * kv[e].z + qv[a].w * kv[e].w;
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int e = 0; e < 4; ++e) s_p[(i0 + a) * (T + 1) + j0 + e] = acc[a][e];     // no 1/sqrt(d): model_layers.py:140
    }
    __syncthreads();
    // softmax over j for each row i; results kept in registers, then written transposed after a barrier
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int kPerLane = (T + 31) / 32, kRowsPerWarp = (T + 7) / 8;
    float pr[kRowsPerWarp][kPerLane];
#pragma unroll
    for (int rr = 0; rr < kRowsPerWarp; ++rr) {
        const int i = warp + rr * 8;
        float m = -INFINITY;
#pragma unroll
        for (int u = 0; u < kPerLane; ++u) {
            const int j = lane + 32 * u;
            pr[rr][u] = (i < T && j < T) ? s_p[i * (T + 1) + j] : -INFINITY;
            m = fmaxf(m, pr[rr][u]);
        }
        m = warp_max(m);
        float s = 0.f;
#pragma unroll
        for (int u = 0; u < kPerLane; ++u) { pr[rr][u] = __expf(pr[rr][u] - m); s += pr[rr][u]; }
        s = warp_sum(s);
        const float inv = 1.f / s;
#pragma unroll
        for (int u = 0; u < kPerLane; ++u) pr[rr][u] *= inv;
    }
    __syncthreads();
#pragma unroll
    for (int rr = 0; rr < kRowsPerWarp; ++rr) {
        const int i = warp + rr * 8;
#pragma unroll
        for (int u = 0; u < kPerLane; ++u) {
            const int j = lane + 32 * u;
            if (i < T && j < T) s_p[j * T + i] = pr[rr][u];      // P^T[j][i]
        }
    }
    __syncthreads();
    const float gamma = *gamma_p;
    const __nv_bfloat16* vbase = base + 2 * d;
    if constexpr (T <= 32) {
        // each thread owns two adjacent channels (4-byte loads / stores, two independent accumulator sets)
        for (int c = 2 * threadIdx.x; c < C; c += 2 * blockDim.x) {
            float acc0[T], acc1[T];
#pragma unroll
            for (int t = 0; t < T; ++t) { acc0[t] = 0.f; acc1[t] = 0.f; }
#pragma unroll 4
            for (int j = 0; j < T; ++j) {
                const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(vbase + static_cast<long long>(j) * ld + c));
                const float4* pt = reinterpret_cast<const float4*>(s_p + j * T);
#pragma unroll
                for (int t4 = 0; t4 < T / 4; ++t4) {
                    const float4 p4 = pt[t4];
                    acc0[4 * t4] = fmaf(p4.x, v.x, acc0[4 * t4]);         acc1[4 * t4] = fmaf(p4.x, v.y, acc1[4 * t4]);
                    acc0[4 * t4 + 1] = fmaf(p4.y, v.x, acc0[4 * t4 + 1]); acc1[4 * t4 + 1] = fmaf(p4.y, v.y, acc1[4 * t4 + 1]);
                    acc0[4 * t4 + 2] = fmaf(p4.z, v.x, acc0[4 * t4 + 2]); acc1[4 * t4 + 2] = fmaf(p4.z, v.y, acc1[4 * t4 + 2]);
                    acc0[4 * t4 + 3] = fmaf(p4.w, v.x, acc0[4 * t4 + 3]); acc1[4 * t4 + 3] = fmaf(p4.w, v.y, acc1[4 * t4 + 3]);
                }
            }
#pragma unroll
            for (int t = 0; t < T; ++t) {
                const long long o = (b * T + t) * C + c;
                const float2 xv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(x + o));
                float r0 = gamma * acc0[t] + xv.x, r1 = gamma * acc1[t] + xv.y;
                if (res2) {
                    const float2 rv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(res2 + o));
                    r0 += rv.x; r1 += rv.y;
                }
                *reinterpret_cast<__nv_bfloat162*>(out + o) = __floats2bfloat162_rn(r0, r1);
            }
        }
    } else {
        // long sequences: one channel per thread keeps the T accumulators in registers
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            float acc[T];
#pragma unroll
            for (int t = 0; t < T; ++t) acc[t] = 0.f;
#pragma unroll 4
            for (int j = 0; j < T; ++j) {
                const float v = __bfloat162float(vbase[static_cast<long long>(j) * ld + c]);
                const float4* pt = reinterpret_cast<const float4*>(s_p + j * T);
#pragma unroll
                for (int t4 = 0; t4 < T / 4; ++t4) {
                    const float4 p4 = pt[t4];
                    acc[4 * t4] = fmaf(p4.x, v, acc[4 * t4]);
                    acc[4 * t4 + 1] = fmaf(p4.y, v, acc[4 * t4 + 1]);
                    acc[4 * t4 + 2] = fmaf(p4.z, v, acc[4 * t4 + 2]);
                    acc[4 * t4 + 3] = fmaf(p4.w, v, acc[4 * t4 + 3]);
                }
            }
#pragma unroll
            for (int t = 0; t < T; ++t) {
                const long long o = (b * T + t) * C + c;
                float r = gamma * acc[t] + __bfloat162float(x[o]);
                if (res2) r += __bfloat162float(res2[o]);
                out[o] = __float2bfloat16_rn(r);
            }
        }
    }
}

// Any sequence length up to 64 (the UNet attentions run at T/4 and T/2 of the clip length, e.g. 2, 6, 14): plain
// loops, one CTA per clip; only the odd sizes the tiled kernel above does not cover come here.
__global__ void __launch_bounds__(256)
attention_generic_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ x,
                         const __nv_bfloat16* __restrict__ res2, const float* __restrict__ gamma_p, int T, int C,
                         __nv_bfloat16* __restrict__ out) {
    __shared__ float s_p[64 * 64];
    const int d = C / 8, ld = 2 * d + C;
    const long long b = blockIdx.x;
    const __nv_bfloat16* base = qkv + b * T * ld;
    for (int idx = threadIdx.x; idx < T * T; idx += blockDim.x) {
        const int i = idx / T, j = idx - i * T;
        const __nv_bfloat16* qi = base + static_cast<long long>(i) * ld;
        const __nv_bfloat16* kj = base + static_cast<long long>(j) * ld + d;
        float acc = 0.f;
        for (int c = 0; c < d; ++c) acc = fmaf(__bfloat162float(qi[c]), __bfloat162float(kj[c]), acc);
        s_p[idx] = acc;                                   // no 1/sqrt(d): model_layers.py:140
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = warp; i < T; i += 8) {
        float v0 = lane < T ? s_p[i * T + lane] : -INFINITY, v1 = lane + 32 < T ? s_p[i * T + lane + 32] : -INFINITY;
        const float m = warp_max(fmaxf(v0, v1));
        v0 = lane < T ? __expf(v0 - m) : 0.f;
        v1 = lane + 32 < T ? __expf(v1 - m) : 0.f;
        const float inv = 1.f / warp_sum(v0 + v1);
        if (lane < T) s_p[i * T + lane] = v0 * inv;
        if (lane + 32 < T) s_p[i * T + lane + 32] = v1 * inv;
    }
    __syncthreads();
    const float gamma = *gamma_p;
    const __nv_bfloat16* vbase = base + 2 * d;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        for (int t = 0; t < T; ++t) {
            float acc = 0.f;
            for (int j = 0; j < T; ++j) acc = fmaf(s_p[t * T + j], __bfloat162float(vbase[static_cast<long long>(j) * ld + c]), acc);
            const long long o = (b * T + t) * C + c;
            float r = gamma * acc + __bfloat162float(x[o]);
            if (res2) r += __bfloat162float(res2[o]);
            out[o] = __float2bfloat16_rn(r);
        }
    }
}

// Sequences longer than 64 steps (the reference has no limit, model_layers.py:133-146): one CTA per (clip, 8 query
// rows); the 8 x T score rows live in shared memory, softmax by one warp per row, then P.v with two channels per thread.
constexpr int kLongRows = 8;
__global__ void __launch_bounds__(256)
attention_long_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ x,
                      const __nv_bfloat16* __restrict__ res2, const float* __restrict__ gamma_p, int T, int C,
                      __nv_bfloat16* __restrict__ out) {
    extern __shared__ __align__(16) float s_long[];
    const int d = C / 8, ld = 2 * d + C;
    float* s_q = s_long;                                   // [8][d]
    float* s_p = s_long + kLongRows * d;                   // [8][T]
    const long long b = blockIdx.x;
    const int i0 = blockIdx.y * kLongRows, rows = min(kLongRows, T - i0);
    const __nv_bfloat16* base = qkv + b * T * ld;
    for (int idx = threadIdx.x; idx < rows * d; idx += blockDim.x) {
        const int r = idx / d, c = idx - r * d;
        s_q[idx] = __bfloat162float(base[static_cast<long long>(i0 + r) * ld + c]);
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp < rows) {
        const float* q = s_q + warp * d;
        float* prow = s_p + warp * T;
        float m = -INFINITY;
        for (int j = lane; j < T; j += 32) {
            const uint4* k8 = reinterpret_cast<const uint4*>(base + static_cast<long long>(j) * ld + d);   // d % 8 == 0, ld % 8 == 0
            float acc = 0.f;
            for (int c = 0; c < d / 8; ++c) {
                const uint4 u = __ldg(k8 + c);
                const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    acc = fmaf(q[c * 8 + 2 * e], __uint_as_float(w4[e] << 16), acc);
                    acc = fmaf(q[c * 8 + 2 * e + 1], __uint_as_float(w4[e] & 0xffff0000u), acc);
                }
            }
            prow[j] = acc;                                  // no 1/sqrt(d): model_layers.py:140
            m = fmaxf(m, acc);
        }
        m = warp_max(m);
        float sum = 0.f;
        for (int j = lane; j < T; j += 32) { const float e = __expf(prow[j] - m); prow[j] = e; sum += e; }
        const float inv = 1.f / warp_sum(sum);
        for (int j = lane; j < T; j += 32) prow[j] *= inv;
    }
    __syncthreads();
    const float gamma = *gamma_p;
    const __nv_bfloat16* vbase = base + 2 * d;
    for (int c = 2 * threadIdx.x; c < C; c += 2 * blockDim.x) {
        float acc[kLongRows][2];
#pragma unroll
        for (int r = 0; r < kLongRows; ++r) acc[r][0] = acc[r][1] = 0.f;
        for (int j = 0; j < T; ++j) {
            const uint32_t u = __ldg(reinterpret_cast<const uint32_t*>(vbase + static_cast<long long>(j) * ld + c));
            const float v0 = __uint_as_float(u << 16), v1 = __uint_as_float(u & 0xffff0000u);
#pragma unroll
            for (int r = 0; r < kLongRows; ++r) {
                const float pj = s_p[r * T + j];
                acc[r][0] = fmaf(pj, v0, acc[r][0]);
                acc[r][1] = fmaf(pj, v1, acc[r][1]);
            }
        }
#pragma unroll
        for (int r = 0; r < kLongRows; ++r) {
            if (r < rows) {
                const long long o = (b * T + i0 + r) * C + c;
                float r0 = gamma * acc[r][0] + __bfloat162float(x[o]), r1 = gamma * acc[r][1] + __bfloat162float(x[o + 1]);
                if (res2) { r0 += __bfloat162float(res2[o]); r1 += __bfloat162float(res2[o + 1]); }
                *reinterpret_cast<__nv_bfloat162*>(out + o) = __floats2bfloat162_rn(r0, r1);
            }
        }
    }
}

// ------------------------------------------------------------------------------ channel attention
// One CTA (256 threads = 8 warps) per clip.  Pooling and the final scaling move 8 channels per thread as 16-byte
// vectors, warp w taking the time steps w, w + 8, ...; the per-warp partial sums / maxima meet in shared memory.
__global__ void __launch_bounds__(256)
channel_attention_kernel(const __nv_bfloat16* __restrict__ x, int T, int C, int hidden, const float* __restrict__ w0,
                         const float* __restrict__ b0, const float* __restrict__ w2, const float* __restrict__ b2,
                         __nv_bfloat16* __restrict__ out) {
    extern __shared__ __align__(16) float s_ca[];
    float* s_psum = s_ca;                // [8][C]
    float* s_pmax = s_psum + 8 * C;      // [8][C]
    float* s_avg = s_pmax + 8 * C;       // [C]
    float* s_max = s_avg + C;            // [C]
    float* s_scale = s_max + C;          // [C]
    float* s_h = s_scale + C;            // [2][hidden]
    const long long b = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const __nv_bfloat16* xb = x + b * T * C;
    const int groups = C / 8;
    for (int cg = lane; cg < groups; cg += 32) {
        float sum[8], mx[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) { sum[e] = 0.f; mx[e] = -INFINITY; }
        for (int t = warp; t < T; t += 8) {
            float f[8];
            bf16x8_to_float(__ldg(reinterpret_cast<const uint4*>(xb + static_cast<long long>(t) * C + cg * 8)), f);
#pragma unroll
            for (int e = 0; e < 8; ++e) { sum[e] += f[e]; mx[e] = fmaxf(mx[e], f[e]); }
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) { s_psum[warp * C + cg * 8 + e] = sum[e]; s_pmax[warp * C + cg * 8 + e] = mx[e]; }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float sum = 0.f, mx = -INFINITY;
#pragma unroll
        for (int w = 0; w < 8; ++w) { sum += s_psum[w * C + c]; mx = fmaxf(mx, s_pmax[w * C + c]); }
        s_avg[c] = sum / static_cast<float>(T);
        s_max[c] = mx;
    }
    __syncthreads();
    for (int u = warp; u < 2 * hidden; u += 8) {
        const int unit = u % hidden;
        const float* src = u < hidden ? s_avg : s_max;
        float acc = 0.f;
        for (int k = lane; k < C; k += 32) acc = fmaf(__ldg(w0 + unit * C + k), src[k], acc);
        acc = warp_sum(acc);
        if (lane == 0) s_h[u] = fmaxf(acc + __ldg(b0 + unit), 0.f);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float za = __ldg(b2 + c), zm = za;
        for (int u = 0; u < hidden; ++u) {
            const float w = __ldg(w2 + u * C + c);           // w2 is stored transposed [hidden][C]
            za = fmaf(w, s_h[u], za);
            zm = fmaf(w, s_h[hidden + u], zm);
        }
        s_scale[c] = 1.f / (1.f + __expf(-za)) + 1.f / (1.f + __expf(-zm));      // sigmoid each, then add
    }
    __syncthreads();
    __nv_bfloat16* ob = out + b * T * C;
    for (int cg = lane; cg < groups; cg += 32) {
        const float4 s0 = *reinterpret_cast<const float4*>(s_scale + cg * 8), s1 = *reinterpret_cast<const float4*>(s_scale + cg * 8 + 4);
        const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
        for (int t = warp; t < T; t += 8) {
            const long long o = static_cast<long long>(t) * C + cg * 8;
            float f[8];
            bf16x8_to_float(__ldg(reinterpret_cast<const uint4*>(xb + o)), f);
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] *= sc[e];
            *reinterpret_cast<uint4*>(ob + o) = float_to_bf16x8(f);
        }
    }
}

// -------------------------------------------------------------------------------------- layernorm
__global__ void __launch_bounds__(256)
layernorm256_kernel(const __nv_bfloat16* __restrict__ x, long long rows, const float* __restrict__ gamma,
                    const float* __restrict__ beta, __nv_bfloat16* __restrict__ out) {
    const long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    float f[8];
    bf16x8_to_float(*reinterpret_cast<const uint4*>(x + row * 256 + lane * 8), f);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += f[i];
    const float mean = warp_sum(s) * (1.f / 256.f);
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { const float dlt = f[i] - mean; v = fmaf(dlt, dlt, v); }
    const float rstd = rsqrtf(warp_sum(v) * (1.f / 256.f) + 1e-5f);
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = (f[i] - mean) * rstd * gamma[lane * 8 + i] + beta[lane * 8 + i];
    *reinterpret_cast<uint4*>(out + row * 256 + lane * 8) = float_to_bf16x8(f);
}

// ------------------------------------------------------------------------------------ pose losses
__global__ void __launch_bounds__(256)
angle_loss_kernel(const float* __restrict__ pose, long long n_frames, const int* __restrict__ triples, int n_hand,
                  int n_body, double* __restrict__ scratch) {
    // one thread per (frame, triple): 35 triples x B*T frames of independent atan2 work
    __shared__ float s_hand[8], s_body[8];
    const int n_tri = n_hand + n_body;
    const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    float hand = 0.f, body = 0.f;
    if (idx < n_frames * n_tri) {
        const long long f = idx / n_tri;
        const int t = static_cast<int>(idx - f * n_tri);
        const float* p = pose + f * 104;                 // interleaved (x, y) per joint: view(B,T,52,2)
        const float pi = 3.14159265358979323846f;
        const int jp = __ldg(triples + 3 * t), jj = __ldg(triples + 3 * t + 1), jc = __ldg(triples + 3 * t + 2);
        const float2 pp = __ldg(reinterpret_cast<const float2*>(p) + jp), pj = __ldg(reinterpret_cast<const float2*>(p) + jj),
                     pc = __ldg(reinterpret_cast<const float2*>(p) + jc);
        const float ax = pj.x - pp.x, ay = pj.y - pp.y;
        const float bx = pc.x - pj.x, by = pc.y - pj.y;
        const float th = atan2f(ax * by - ay * bx, ax * bx + ay * by);
        const float lo = t < n_hand ? 0.f : -0.5f * pi;
        const float pen = fmaxf(lo - th, 0.f) + fmaxf(th - pi, 0.f);
        if (t < n_hand) hand = pen; else body = pen;
    }
    hand = warp_sum(hand);
    body = warp_sum(body);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { s_hand[warp] = hand; s_body[warp] = body; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double h = 0.0, bsum = 0.0;
        for (int w = 0; w < 8; ++w) { h += s_hand[w]; bsum += s_body[w]; }
        atomicAdd(&scratch[0], h);
        atomicAdd(&scratch[1], bsum);
    }
}

__global__ void __launch_bounds__(64)
bone_loss_kernel(const float* __restrict__ gen, const float* __restrict__ real, int T, const int* __restrict__ parents,
                 double* __restrict__ scratch) {
    __shared__ float s_sq[2];
    const long long b = blockIdx.x;
    const int j = threadIdx.x;                            // joint; bones are joints with a parent
    float sq = 0.f;
    if (j < 52 && parents[j] >= 0) {
        const int par = parents[j];
        float lg = 0.f, lr = 0.f;
        for (int t = 0; t < T; ++t) {
            const float* g = gen + (b * T + t) * 104;
            const float* r = real + (b * T + t) * 104;
            const float gx = g[2 * j] - g[2 * par], gy = g[2 * j + 1] - g[2 * par + 1];
            const float rx = r[2 * j] - r[2 * par], ry = r[2 * j + 1] - r[2 * par + 1];
            lg += sqrtf(gx * gx + gy * gy);
            lr += sqrtf(rx * rx + ry * ry);
        }
        const float dlt = (lg - lr) / static_cast<float>(T);
        sq = dlt * dlt;
    }
    sq = warp_sum(sq);
    if ((threadIdx.x & 31) == 0) s_sq[threadIdx.x >> 5] = sq;
    __syncthreads();
    if (threadIdx.x == 0) atomicAdd(&scratch[2], static_cast<double>(s_sq[0]) + static_cast<double>(s_sq[1]));
}

__global__ void finalize_losses_kernel(const double* __restrict__ scratch, long long n_frames, long long n_clips,
                                       int n_hand, int n_body, int n_bones, bool with_bone, float* __restrict__ out) {
    const double hand = n_hand > 0 ? scratch[0] / (static_cast<double>(n_frames) * n_hand) : 0.0;
    const double body = n_body > 0 ? scratch[1] / (static_cast<double>(n_frames) * n_body) : 0.0;
    out[0] = static_cast<float>(0.7 * hand + 0.3 * body);
    out[1] = with_bone ? static_cast<float>(scratch[2] / (static_cast<double>(n_clips) * n_bones)) : 0.f;
}

// ------------------------------------------------------------------------------- layout / dtype
__global__ void f32_to_bf16_kernel(const float* __restrict__ in, long long n, __nv_bfloat16* __restrict__ out) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) out[i] = __float2bfloat16_rn(in[i]);
}
__global__ void bf16_to_f32_kernel(const __nv_bfloat16* __restrict__ in, long long n, float* __restrict__ out) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) out[i] = __bfloat162float(in[i]);
}
__global__ void ncw_to_btc_kernel(const float* __restrict__ in, int C, int T, __nv_bfloat16* __restrict__ out) {
    __shared__ float tile[32][33];
    const long long b = blockIdx.z;
    const int c0 = blockIdx.y * 32, t0 = blockIdx.x * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, t = t0 + threadIdx.x;
        tile[i][threadIdx.x] = (c < C && t < T) ? in[(b * C + c) * T + t] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int t = t0 + i, c = c0 + threadIdx.x;
        if (t < T && c < C) out[(b * T + t) * C + c] = __float2bfloat16_rn(tile[threadIdx.x][i]);
    }
}
__global__ void btc_to_ncw_kernel(const __nv_bfloat16* __restrict__ in, int C, int T, float* __restrict__ out) {
    __shared__ float tile[32][33];
    const long long b = blockIdx.z;
    const int c0 = blockIdx.y * 32, t0 = blockIdx.x * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int t = t0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (c < C && t < T) ? __bfloat162float(in[(b * T + t) * C + c]) : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, t = t0 + threadIdx.x;
        if (t < T && c < C) out[(b * C + c) * T + t] = tile[threadIdx.x][i];
    }
}

}  // namespace

#define A2M_AFTER_LAUNCH()  \
    a2m_count_launch();     \
    A2M_LAUNCH_CHECK();     \
    return A2M_OK

int launch_conv0(const float* mel, long long stride_b, long long stride_t, int B, int T, int F, const float* w_folded,
                 const float* bias_folded, __nv_bfloat16* out, cudaStream_t stream, int n_inner, long long stride_outer) {
    if (n_inner <= 0) { n_inner = B > 0 ? B : 1; stride_outer = 0; }
    A2M_ARG_CHECK(T % 2 == 0 && F % 2 == 0 && B <= 65535, "conv0: T %d, F %d, B %d", T, F, B);
    const int stride = ((F + 2 + 3) / 4) * 4 + 4;
    conv0_kernel<<<dim3((T / 2 + kConv0Rows - 1) / kConv0Rows, B), 256, (2 * kConv0Rows + 2) * stride * sizeof(float), stream>>>(
        mel, stride_b, stride_t, n_inner, stride_outer, T, F, w_folded, bias_folded, out);
    A2M_AFTER_LAUNCH();
}

// Any output length (T not a multiple of Hc: T % 8 != 0, or AudioEncoder.forward(x, time_steps != T),
// model_layers.py:267-279): one thread per (clip, output step, 8 channels), the two source rows read directly.
__global__ void __launch_bounds__(256)
time_interp_general_kernel(const float* __restrict__ in, int n_planes, long long plane_stride, int Hc, int T, int C,
                           long long total, __nv_bfloat16* __restrict__ out) {
    const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int c8 = C / 8;
    const int c = static_cast<int>(idx % c8) * 8;
    const long long bt = idx / c8;
    const int t = static_cast<int>(bt % T);
    const long long b = bt / T;
    float src = (static_cast<float>(Hc) / static_cast<float>(T)) * (static_cast<float>(t) + 0.5f) - 0.5f;
    if (src < 0.f) src = 0.f;
    const int i0 = min(static_cast<int>(src), Hc - 1);
    const int i1 = i0 + (i0 < Hc - 1 ? 1 : 0);
    const float l1 = src - static_cast<float>(i0), l0 = 1.f - l1;
    float v[2][8];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int e = 0; e < 8; ++e) v[r][e] = 0.f;
    for (int pl = 0; pl < n_planes; ++pl) {                   // fixed summation order over the split-K planes
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const float4* p4 = reinterpret_cast<const float4*>(in + pl * plane_stride + (b * Hc + (r ? i1 : i0)) * C + c);
            const float4 a0 = __ldg(p4), a1 = __ldg(p4 + 1);
            v[r][0] += a0.x; v[r][1] += a0.y; v[r][2] += a0.z; v[r][3] += a0.w;
            v[r][4] += a1.x; v[r][5] += a1.y; v[r][6] += a1.z; v[r][7] += a1.w;
        }
    }
    uint4 o;
    uint32_t w[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float f0 = l0 * leaky(v[0][2 * e]) + l1 * leaky(v[1][2 * e]);
        const float f1 = l0 * leaky(v[0][2 * e + 1]) + l1 * leaky(v[1][2 * e + 1]);
        __nv_bfloat162 h = __floats2bfloat162_rn(f0, f1);
        w[e] = *reinterpret_cast<uint32_t*>(&h);
    }
    o.x = w[0]; o.y = w[1]; o.z = w[2]; o.w = w[3];
    *reinterpret_cast<uint4*>(out + (b * T + t) * C + c) = o;
}

int launch_time_interp(const float* in, int n_planes, int B, int Hc, int T, int C, __nv_bfloat16* out, cudaStream_t stream) {
    A2M_ARG_CHECK(C % 8 == 0 && Hc >= 1 && T >= 1 && n_planes >= 1 && n_planes <= kMaxPlanes,
                  "time_interp: C = %d, Hc = %d, T = %d, planes = %d", C, Hc, T, n_planes);
    if (T % Hc != 0) {
        const long long total = static_cast<long long>(B) * T * (C / 8);
        time_interp_general_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(
            in, n_planes, static_cast<long long>(B) * Hc * C, Hc, T, C, total, out);
        A2M_AFTER_LAUNCH();
    }
    const long long total = static_cast<long long>(B) * Hc * (C / 8);
    time_interp_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(in, n_planes, static_cast<long long>(B) * Hc * C, Hc, T, C, total, out);
    A2M_AFTER_LAUNCH();
}

template <int T>
static int launch_attention_t(const __nv_bfloat16* qkv, const __nv_bfloat16* x, const __nv_bfloat16* res2,
                              const float* gamma, int B, int C, __nv_bfloat16* out, cudaStream_t stream) {
    const int d = C / 8;
    const size_t smem = (2 * static_cast<size_t>(T) * (d + 4) + static_cast<size_t>(T) * (T + 1)) * sizeof(float);
    static A2mPerDeviceOnce configured;
    if (configured.first()) {
        A2M_CUDA_CHECK(cudaFuncSetAttribute(attention_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    }
    A2M_ARG_CHECK(smem <= 100 * 1024, "attention: T %d x C %d needs %zu B of shared memory", T, C, smem);
    attention_kernel<T><<<B, 256, smem, stream>>>(qkv, x, res2, gamma, C, out);
    A2M_AFTER_LAUNCH();
}

int launch_attention(const __nv_bfloat16* qkv, const __nv_bfloat16* x, const __nv_bfloat16* res2, const float* gamma,
                     int B, int T, int C, __nv_bfloat16* out, cudaStream_t stream) {
    A2M_ARG_CHECK(C % 64 == 0, "attention: C = %d must be a multiple of 64", C);
    // the tiled kernel keeps q, k ([T][C/8 + 4] each) and the scores in shared memory; wide layers at long T fall through
    const size_t tiled_smem = (2 * static_cast<size_t>(T) * (C / 8 + 4) + static_cast<size_t>(T) * (T + 1)) * sizeof(float);
    switch (tiled_smem <= 100 * 1024 ? T : 0) {
        case 8: return launch_attention_t<8>(qkv, x, res2, gamma, B, C, out, stream);
        case 16: return launch_attention_t<16>(qkv, x, res2, gamma, B, C, out, stream);
        case 24: return launch_attention_t<24>(qkv, x, res2, gamma, B, C, out, stream);
        case 32: return launch_attention_t<32>(qkv, x, res2, gamma, B, C, out, stream);
        case 40: return launch_attention_t<40>(qkv, x, res2, gamma, B, C, out, stream);
        case 48: return launch_attention_t<48>(qkv, x, res2, gamma, B, C, out, stream);
        case 56: return launch_attention_t<56>(qkv, x, res2, gamma, B, C, out, stream);
        case 64: return launch_attention_t<64>(qkv, x, res2, gamma, B, C, out, stream);
        default:
            break;
    }
    A2M_ARG_CHECK(T >= 1 && T <= 4096, "attention: T = %d; this build implements T <= 4096", T);
    if (T <= 64) {
        attention_generic_kernel<<<B, 256, 0, stream>>>(qkv, x, res2, gamma, T, C, out);
        A2M_AFTER_LAUNCH();
    }
    const size_t smem = static_cast<size_t>(kLongRows) * (C / 8 + T) * sizeof(float);
    A2M_ARG_CHECK(smem <= 200 * 1024 && C % 64 == 0, "attention: T %d x C %d needs %zu B of shared memory", T, C, smem);
    static A2mPerDeviceOnce configured;
    if (configured.first())
        A2M_CUDA_CHECK(cudaFuncSetAttribute(attention_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attention_long_kernel<<<dim3(B, (T + kLongRows - 1) / kLongRows), 256, smem, stream>>>(qkv, x, res2, gamma, T, C, out);
    A2M_AFTER_LAUNCH();
}

int launch_channel_attention(const __nv_bfloat16* x, int B, int T, int C, int hidden, const float* w0, const float* b0,
                             const float* w2, const float* b2, __nv_bfloat16* out, cudaStream_t stream) {
    A2M_ARG_CHECK(C % 8 == 0 && C <= 1024 && hidden >= 1 && hidden <= 256, "channel attention: C = %d hidden = %d", C, hidden);
    channel_attention_kernel<<<B, 256, (19 * C + 2 * hidden) * sizeof(float), stream>>>(x, T, C, hidden, w0, b0, w2, b2, out);
    A2M_AFTER_LAUNCH();
}

int launch_layernorm(const __nv_bfloat16* x, long long rows, int C, const float* gamma, const float* beta,
                     __nv_bfloat16* out, cudaStream_t stream) {
    A2M_ARG_CHECK(C == 256, "layernorm: C = %d (this build implements 256)", C);
    layernorm256_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, stream>>>(x, rows, gamma, beta, out);
    A2M_AFTER_LAUNCH();
}

int launch_pose_losses(const float* pose, const float* real_pose, int B, int T, const int* triples, int n_hand,
                       int n_body, const int* parents, double* scratch, float* losses_out, cudaStream_t stream) {
    const long long frames = static_cast<long long>(B) * T;
    A2M_CUDA_CHECK(cudaMemsetAsync(scratch, 0, 4 * sizeof(double), stream));
    const long long items = frames * (n_hand + n_body);
    angle_loss_kernel<<<static_cast<unsigned>((items + 255) / 256), 256, 0, stream>>>(pose, frames, triples, n_hand, n_body, scratch);
    a2m_count_launch();
    A2M_LAUNCH_CHECK();
    if (real_pose) {
        bone_loss_kernel<<<B, 64, 0, stream>>>(pose, real_pose, T, parents, scratch);
        a2m_count_launch();
        A2M_LAUNCH_CHECK();
    }
    finalize_losses_kernel<<<1, 1, 0, stream>>>(scratch, frames, B, n_hand, n_body, 51, real_pose != nullptr, losses_out);
    A2M_AFTER_LAUNCH();
}

int launch_f32_to_bf16(const float* in, long long n, __nv_bfloat16* out, cudaStream_t stream) {
    f32_to_bf16_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(in, n, out);
    A2M_AFTER_LAUNCH();
}
int launch_bf16_to_f32(const __nv_bfloat16* in, long long n, float* out, cudaStream_t stream) {
    bf16_to_f32_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(in, n, out);
    A2M_AFTER_LAUNCH();
}
// ------------------------------------------------------------------------------ discriminator (SelfAttention_D)
// pose [B, T, 104] fp32 -> [B, T_alloc, 128] bf16 (channels 104..127 zero; rows T..T_alloc-1 are never written)
__global__ void pose_pad_kernel(const float* __restrict__ pose, long long total, int T, int T_alloc, int C, int C_pad,
                                __nv_bfloat16* __restrict__ out) {
    const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int c = static_cast<int>(idx % C_pad);
    const long long bt = idx / C_pad;
    const int t = static_cast<int>(bt % T);
    const long long b = bt / T;
    out[(b * T_alloc + t) * C_pad + c] = __float2bfloat16_rn(c < C ? pose[bt * C + c] : 0.f);
}
// mean over the T rows of [B, T, C] bf16 -> [B, C] bf16 (x.mean(dim=2) of real_motion_model.py:599,609)
__global__ void mean_time_kernel(const __nv_bfloat16* __restrict__ x, long long total, int T, int C, __nv_bfloat16* __restrict__ out) {
    const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int c = static_cast<int>(idx % C);
    const long long b = idx / C;
    float s = 0.f;
    for (int t = 0; t < T; ++t) s += __bfloat162float(x[(b * T + t) * C + c]);
    out[idx] = __float2bfloat16_rn(s / static_cast<float>(T));
}
// One GATConv(64, 64, heads=4, concat=False) layer on its own (no LayerNorm / residual: real_motion_model.py:604,616),
// one CTA per graph, fp32 arithmetic.  wt: lin.weight transposed to [64][256]; att_src / att_dst: [4][64].
__global__ void __launch_bounds__(256)
gat_single_kernel(const __nv_bfloat16* __restrict__ x, int J, const float* __restrict__ wt, const float* __restrict__ att_src,
                  const float* __restrict__ att_dst, const float* __restrict__ bias, const int* __restrict__ nbr,
                  const int* __restrict__ deg, __nv_bfloat16* __restrict__ out) {
    extern __shared__ __align__(16) float s_gat[];
    float* s_x = s_gat;                          // [J][64]
    float* s_h = s_x + J * 64;                   // [J][256]
    float* s_src = s_h + J * 256;                // [J][4]
    float* s_dst = s_src + J * 4;                // [J][4]
    const long long g = blockIdx.x;
    const int tid = threadIdx.x;
    for (int i = tid; i < J * 64; i += 256) s_x[i] = __bfloat162float(x[g * J * 64 + i]);
    __syncthreads();
    {   // h[j][o] = sum_f W[o][f] x[j][f]; thread o keeps its weight row in registers
        float w[64];
#pragma unroll
        for (int f = 0; f < 64; ++f) w[f] = __ldg(wt + f * 256 + tid);
        for (int j = 0; j < J; ++j) {
            float acc = 0.f;
#pragma unroll
            for (int f = 0; f < 64; ++f) acc = fmaf(w[f], s_x[j * 64 + f], acc);
            s_h[j * 256 + tid] = acc;
        }
    }
    __syncthreads();
    const int warp = tid >> 5, lane = tid & 31;
    for (int item = warp; item < J * 4; item += 8) {          // attention logits per (node, head)
        const int j = item >> 2, h = item & 3;
        const float h0 = s_h[j * 256 + h * 64 + lane], h1 = s_h[j * 256 + h * 64 + 32 + lane];
        const float a = warp_sum(h0 * __ldg(att_src + h * 64 + lane) + h1 * __ldg(att_src + h * 64 + 32 + lane));
        const float d = warp_sum(h0 * __ldg(att_dst + h * 64 + lane) + h1 * __ldg(att_dst + h * 64 + 32 + lane));
        if (lane == 0) { s_src[item] = a; s_dst[item] = d; }
    }
    __syncthreads();
    for (int idx = tid; idx < J * 64; idx += 256) {           // out[i][o] = mean_h sum_j alpha^h_ij h[j][h][o] + bias[o]
        const int i = idx >> 6, o = idx & 63;
        const int dg = deg[i];
        float acc = 0.f;
        for (int h = 0; h < 4; ++h) {
            const float sd = s_dst[i * 4 + h];
            float e[kMaxDeg + 1];
            float m = leaky(s_src[i * 4 + h] + sd);
            e[0] = m;
            for (int k = 0; k < dg; ++k) { e[k + 1] = leaky(s_src[nbr[i * kMaxDeg + k] * 4 + h] + sd); m = fmaxf(m, e[k + 1]); }
            float den = 0.f, num = 0.f;
            for (int k = 0; k <= dg; ++k) {
                const float w = __expf(e[k] - m);
                const int j = k == 0 ? i : nbr[i * kMaxDeg + k - 1];
                den += w;
                num = fmaf(w, s_h[j * 256 + h * 64 + o], num);
            }
            acc += num / den;
        }
        out[g * J * 64 + idx] = __float2bfloat16_rn(0.25f * acc + __ldg(bias + o));
    }
}
// logits Conv1d(2 C -> 1, k3 p1) over cat([x, graph features repeated over time]) (real_motion_model.py:619-630):
// one warp per (clip, step).  w: [3][2 C] fp32 (tap-major), x [B, T, C] bf16, xg [B, C] bf16 -> out [B, T] fp32
__global__ void __launch_bounds__(256)
disc_logits_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ xg, const float* __restrict__ w,
                   const float* __restrict__ bias, long long n_items, int T, int C, float* __restrict__ out) {
    const long long item = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
    if (item >= n_items) return;
    const int lane = threadIdx.x & 31;
    const int t = static_cast<int>(item % T);
    const long long b = item / T;
    float acc = 0.f;
    for (int tap = 0; tap < 3; ++tap) {
        const int tt = t + tap - 1;
        if (tt < 0 || tt >= T) continue;                      // zero padding covers the graph channels too
        const __nv_bfloat16* xr = x + (b * T + tt) * C;
        const __nv_bfloat16* gr = xg + b * C;
        const float* w0 = w + tap * 2 * C;
        for (int c = lane; c < C; c += 32)
            acc = fmaf(__ldg(w0 + c), __bfloat162float(xr[c]), fmaf(__ldg(w0 + C + c), __bfloat162float(gr[c]), acc));
    }
    acc = warp_sum(acc);
    if (lane == 0) out[item] = acc + __ldg(bias);
}

int launch_pose_pad(const float* pose, int B, int T, int T_alloc, int C, int C_pad, __nv_bfloat16* out, cudaStream_t stream) {
    const long long total = static_cast<long long>(B) * T * C_pad;
    pose_pad_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(pose, total, T, T_alloc, C, C_pad, out);
    A2M_AFTER_LAUNCH();
}
int launch_mean_time(const __nv_bfloat16* x, int B, int T, int C, __nv_bfloat16* out, cudaStream_t stream) {
    const long long total = static_cast<long long>(B) * C;
    mean_time_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(x, total, T, C, out);
    A2M_AFTER_LAUNCH();
}
int launch_gat_single(const __nv_bfloat16* x, long long n_graphs, int J, const float* wt, const float* att_src,
                      const float* att_dst, const float* bias, const int* nbr, const int* deg, __nv_bfloat16* out,
                      cudaStream_t stream) {
    A2M_ARG_CHECK(J >= 1 && J <= 48 && n_graphs >= 1 && n_graphs <= 0x7fffffffLL, "gat: %lld graphs of %d nodes", n_graphs, J);
    const size_t smem = static_cast<size_t>(J) * (64 + 256 + 8) * sizeof(float);
    static A2mPerDeviceOnce configured;
    if (configured.first())
        A2M_CUDA_CHECK(cudaFuncSetAttribute(gat_single_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * (64 + 256 + 8) * 4));
    gat_single_kernel<<<static_cast<unsigned>(n_graphs), 256, smem, stream>>>(x, J, wt, att_src, att_dst, bias, nbr, deg, out);
    A2M_AFTER_LAUNCH();
}
int launch_disc_logits(const __nv_bfloat16* x, const __nv_bfloat16* xg, const float* w, const float* bias, int B, int T,
                       int C, float* out, cudaStream_t stream) {
    const long long n_items = static_cast<long long>(B) * T;
    disc_logits_kernel<<<static_cast<unsigned>((n_items + 7) / 8), 256, 0, stream>>>(x, xg, w, bias, n_items, T, C, out);
    A2M_AFTER_LAUNCH();
}

int launch_ncw_to_btc(const float* in, int B, int C, int T, __nv_bfloat16* out, cudaStream_t stream) {
    ncw_to_btc_kernel<<<dim3((T + 31) / 32, (C + 31) / 32, B), dim3(32, 8), 0, stream>>>(in, C, T, out);
    A2M_AFTER_LAUNCH();
}
int launch_btc_to_ncw(const __nv_bfloat16* in, int B, int C, int T, float* out, cudaStream_t stream) {
    btc_to_ncw_kernel<<<dim3((T + 31) / 32, (C + 31) / 32, B), dim3(32, 8), 0, stream>>>(in, C, T, out);
    A2M_AFTER_LAUNCH();
}

}  // namespace a2m
