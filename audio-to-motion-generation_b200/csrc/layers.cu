// Non-GEMM kernels of the generator forward: first encoder conv, time interpolation, self-attention,
// channel attention, LayerNorm, the static-graph GAT / GraphConv tails and the pose losses.
// See layers.cuh for the contracts and the reference lines each kernel restates.
#include "layers.cuh"

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
// memory; a warp owns one output column at a time (shared-memory reads are broadcasts) and its 32 lanes
// own the 32 channel pairs, so every store instruction writes one whole 128-byte channel row.
constexpr int kConv0Rows = 8;
__global__ void __launch_bounds__(256)
conv0_kernel(const float* __restrict__ mel, long long stride_b, long long stride_t, int T, int F,
             const float* __restrict__ w, const float* __restrict__ bias, __nv_bfloat16* __restrict__ out) {
    extern __shared__ float s_rows[];                 // [2R + 2][F + 2], zero padded left/right/top/bottom
    const int ho0 = blockIdx.x * kConv0Rows;
    const long long b = blockIdx.y;
    const int Ho = T / 2, Wo = F / 2, stride = F + 2;
    const int n_in = 2 * kConv0Rows + 2;
    for (int i = threadIdx.x; i < n_in * stride; i += blockDim.x) {
        const int r = i / stride, col = i - r * stride - 1;
        const int h = 2 * ho0 + r - 1;
        float v = 0.f;
        if (h >= 0 && h < T && col >= 0 && col < F) v = __ldg(mel + b * stride_b + h * stride_t + col);
        s_rows[i] = v;
    }
    const int cp = threadIdx.x & 31, grp = threadIdx.x >> 5;
    float w0[16], w1[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { w0[i] = __ldg(w + (2 * cp) * 16 + i); w1[i] = __ldg(w + (2 * cp + 1) * 16 + i); }
    const float b0 = __ldg(bias + 2 * cp), b1 = __ldg(bias + 2 * cp + 1);
    __syncthreads();
    for (int r = 0; r < kConv0Rows; ++r) {
        const int ho = ho0 + r;
        if (ho >= Ho) break;
        __nv_bfloat16* o = out + ((b * Ho + ho) * Wo) * 64 + 2 * cp;
        for (int wo = grp; wo < Wo; wo += 8) {
            float a0 = b0, a1 = b1;
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float x = s_rows[(2 * r + i) * stride + 2 * wo + j];
                    a0 = fmaf(x, w0[i * 4 + j], a0);
                    a1 = fmaf(x, w1[i * 4 + j], a1);
                }
            *reinterpret_cast<__nv_bfloat162*>(o + static_cast<long long>(wo) * 64) = __floats2bfloat162_rn(leaky(a0), leaky(a1));
        }
    }
}

// ------------------------------------------------------------------------------------ time interp
// in: raw conv-4 sums (+ folded bias) fp32 [B, Hc, C]; LeakyReLU is applied here (the split-K GEMM
// cannot), then the bilinear (T, 1) resize of model_layers.py:277; 8 channels per thread.
__global__ void time_interp_kernel(const float* __restrict__ in, int n_planes, long long plane_stride, int Hc, int T,
                                   int C, long long total8, __nv_bfloat16* __restrict__ out) {
    const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= total8) return;
    const int c8 = C / 8;
    const int c = static_cast<int>(idx % c8) * 8;
    const long long bt = idx / c8;
    const int t = static_cast<int>(bt % T);
    const long long b = bt / T;
    // torch upsample_bilinear2d, align_corners=False: src = scale * (dst + 0.5) - 0.5, clamped at 0
    const float scale = static_cast<float>(Hc) / static_cast<float>(T);
    float src = scale * (static_cast<float>(t) + 0.5f) - 0.5f;
    if (src < 0.f) src = 0.f;
    const int i0 = static_cast<int>(src);
    const int i1 = i0 + (i0 < Hc - 1 ? 1 : 0);
    const float l1 = src - static_cast<float>(i0), l0 = 1.f - l1;
    float x0[8] = {}, x1[8] = {};
    for (int pl = 0; pl < n_planes; ++pl) {           // fixed summation order over the split-K planes
        const float4* r0 = reinterpret_cast<const float4*>(in + pl * plane_stride + (b * Hc + i0) * C + c);
        const float4* r1 = reinterpret_cast<const float4*>(in + pl * plane_stride + (b * Hc + i1) * C + c);
        const float4 a0 = __ldg(r0), a1 = __ldg(r0 + 1), c0 = __ldg(r1), c1 = __ldg(r1 + 1);
        x0[0] += a0.x; x0[1] += a0.y; x0[2] += a0.z; x0[3] += a0.w; x0[4] += a1.x; x0[5] += a1.y; x0[6] += a1.z; x0[7] += a1.w;
        x1[0] += c0.x; x1[1] += c0.y; x1[2] += c0.z; x1[3] += c0.w; x1[4] += c1.x; x1[5] += c1.y; x1[6] += c1.z; x1[7] += c1.w;
    }
    float f[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = l0 * leaky(x0[e]) + l1 * leaky(x1[e]);
    *reinterpret_cast<uint4*>(out + (bt * C + c)) = float_to_bf16x8(f);
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

// ------------------------------------------------------------------------------ channel attention
__global__ void __launch_bounds__(1024)
channel_attention_kernel(const __nv_bfloat16* __restrict__ x, int T, int C, int hidden, const float* __restrict__ w0,
                         const float* __restrict__ b0, const float* __restrict__ w2, const float* __restrict__ b2,
                         __nv_bfloat16* __restrict__ out) {
    extern __shared__ float s_ca[];
    float* s_avg = s_ca;                 // [C]
    float* s_max = s_avg + C;            // [C]
    float* s_h = s_max + C;              // [2][hidden]
    const long long b = blockIdx.x;
    const int c = threadIdx.x;           // blockDim.x == C
    const __nv_bfloat16* xb = x + b * T * C;
    float sum = 0.f, mx = -INFINITY;
    for (int t = 0; t < T; ++t) {
        const float v = __bfloat162float(xb[static_cast<long long>(t) * C + c]);
        sum += v;
        mx = fmaxf(mx, v);
    }
    s_avg[c] = sum / static_cast<float>(T);
    s_max[c] = mx;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
    for (int u = warp; u < 2 * hidden; u += n_warps) {
        const int unit = u % hidden;
        const float* src = u < hidden ? s_avg : s_max;
        float acc = 0.f;
        for (int k = lane; k < C; k += 32) acc = fmaf(w0[unit * C + k], src[k], acc);
        acc = warp_sum(acc);
        if (lane == 0) s_h[u] = fmaxf(acc + b0[unit], 0.f);
    }
    __syncthreads();
    float za = b2[c], zm = b2[c];
    for (int u = 0; u < hidden; ++u) {
        const float w = w2[c * hidden + u];
        za = fmaf(w, s_h[u], za);
        zm = fmaf(w, s_h[hidden + u], zm);
    }
    const float scale = 1.f / (1.f + __expf(-za)) + 1.f / (1.f + __expf(-zm));      // sigmoid each, then add
    __nv_bfloat16* ob = out + b * T * C;
    for (int t = 0; t < T; ++t) {
        const long long o = static_cast<long long>(t) * C + c;
        ob[o] = __float2bfloat16_rn(__bfloat162float(xb[o]) * scale);
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
    __shared__ float s_hand[8], s_body[8];
    const long long f = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    float hand = 0.f, body = 0.f;
    if (f < n_frames) {
        const float* p = pose + f * 104;                 // interleaved (x, y) per joint: view(B,T,52,2)
        const float pi = 3.14159265358979323846f;
        for (int t = 0; t < n_hand + n_body; ++t) {
            const int jp = triples[3 * t], jj = triples[3 * t + 1], jc = triples[3 * t + 2];
            const float ax = p[2 * jj] - p[2 * jp], ay = p[2 * jj + 1] - p[2 * jp + 1];
            const float bx = p[2 * jc] - p[2 * jj], by = p[2 * jc + 1] - p[2 * jj + 1];
            const float th = atan2f(ax * by - ay * bx, ax * bx + ay * by);
            const float lo = t < n_hand ? 0.f : -0.5f * pi;
            const float pen = fmaxf(lo - th, 0.f) + fmaxf(th - pi, 0.f);
            if (t < n_hand) hand += pen; else body += pen;
        }
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
                 const float* bias_folded, __nv_bfloat16* out, cudaStream_t stream) {
    A2M_ARG_CHECK(T % 2 == 0 && F % 2 == 0 && B <= 65535, "conv0: T %d, F %d, B %d", T, F, B);
    conv0_kernel<<<dim3((T / 2 + kConv0Rows - 1) / kConv0Rows, B), 256, (2 * kConv0Rows + 2) * (F + 2) * sizeof(float), stream>>>(
        mel, stride_b, stride_t, T, F, w_folded, bias_folded, out);
    A2M_AFTER_LAUNCH();
}

int launch_time_interp(const float* in, int n_planes, int B, int Hc, int T, int C, __nv_bfloat16* out, cudaStream_t stream) {
    A2M_ARG_CHECK(C % 8 == 0, "time_interp: C = %d", C);
    const long long total = static_cast<long long>(B) * T * (C / 8);
    time_interp_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(in, n_planes, static_cast<long long>(B) * Hc * C, Hc, T, C, total, out);
    A2M_AFTER_LAUNCH();
}

template <int T>
static int launch_attention_t(const __nv_bfloat16* qkv, const __nv_bfloat16* x, const __nv_bfloat16* res2,
                              const float* gamma, int B, int C, __nv_bfloat16* out, cudaStream_t stream) {
    const int d = C / 8;
    const size_t smem = (2 * static_cast<size_t>(T) * (d + 4) + static_cast<size_t>(T) * (T + 1)) * sizeof(float);
    static bool configured = false;
    if (!configured) {
        A2M_CUDA_CHECK(cudaFuncSetAttribute(attention_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        configured = true;
    }
    A2M_ARG_CHECK(smem <= 100 * 1024, "attention: T %d x C %d needs %zu B of shared memory", T, C, smem);
    attention_kernel<T><<<B, 256, smem, stream>>>(qkv, x, res2, gamma, C, out);
    A2M_AFTER_LAUNCH();
}

int launch_attention(const __nv_bfloat16* qkv, const __nv_bfloat16* x, const __nv_bfloat16* res2, const float* gamma,
                     int B, int T, int C, __nv_bfloat16* out, cudaStream_t stream) {
    A2M_ARG_CHECK(C % 64 == 0, "attention: C = %d must be a multiple of 64", C);
    switch (T) {
        case 8: return launch_attention_t<8>(qkv, x, res2, gamma, B, C, out, stream);
        case 16: return launch_attention_t<16>(qkv, x, res2, gamma, B, C, out, stream);
        case 24: return launch_attention_t<24>(qkv, x, res2, gamma, B, C, out, stream);
        case 32: return launch_attention_t<32>(qkv, x, res2, gamma, B, C, out, stream);
        case 40: return launch_attention_t<40>(qkv, x, res2, gamma, B, C, out, stream);
        case 48: return launch_attention_t<48>(qkv, x, res2, gamma, B, C, out, stream);
        case 56: return launch_attention_t<56>(qkv, x, res2, gamma, B, C, out, stream);
        case 64: return launch_attention_t<64>(qkv, x, res2, gamma, B, C, out, stream);
        default:
            a2m_set_error("attention: T = %d; this build implements T in {8, 16, ..., 64}", T);
            return A2M_ERR_UNSUPPORTED;
    }
}

int launch_channel_attention(const __nv_bfloat16* x, int B, int T, int C, int hidden, const float* w0, const float* b0,
                             const float* w2, const float* b2, __nv_bfloat16* out, cudaStream_t stream) {
    A2M_ARG_CHECK(C % 32 == 0 && C <= 1024 && hidden >= 1, "channel attention: C = %d hidden = %d", C, hidden);
    channel_attention_kernel<<<B, C, (2 * C + 2 * hidden) * sizeof(float), stream>>>(x, T, C, hidden, w0, b0, w2, b2, out);
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
    angle_loss_kernel<<<static_cast<unsigned>((frames + 255) / 256), 256, 0, stream>>>(pose, frames, triples, n_hand, n_body, scratch);
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
int launch_ncw_to_btc(const float* in, int B, int C, int T, __nv_bfloat16* out, cudaStream_t stream) {
    ncw_to_btc_kernel<<<dim3((T + 31) / 32, (C + 31) / 32, B), dim3(32, 8), 0, stream>>>(in, C, T, out);
    A2M_AFTER_LAUNCH();
}
int launch_btc_to_ncw(const __nv_bfloat16* in, int B, int C, int T, float* out, cudaStream_t stream) {
    btc_to_ncw_kernel<<<dim3((T + 31) / 32, (C + 31) / 32, B), dim3(32, 8), 0, stream>>>(in, C, T, out);
    A2M_AFTER_LAUNCH();
}

}  // namespace a2m
