// Mel spectrogram front end for the long FFT of the PATS-native features: replaces
// pats/data_loading/audio.py:58-79 log_mel_512 = librosa.feature.melspectrogram(y, sr, n_fft=2048, hop_length=512)
// (centred frames with reflect / zero padding, periodic Hann, power spectrum, 128 Slaney bands) -> zero floor -> log,
// one launch for a batch of clips.  See include/a2m_b200.h (a2m_melspec_*).
//
// One 256-thread CTA transforms one frame at a time and walks a chunk of consecutive frames of one clip (their 75 %
// overlap is an L1/L2 hit, so HBM sees every sample about once):
//   load 2048 samples (padding resolved by index arithmetic) x window, packed even/odd as 1024 complex points
//   -> five radix-4 Stockham passes in shared memory (one butterfly per thread per pass, twiddles from a table)
//   -> real-input untangle, |X|^2 (or |X|) for the 1025 bins
//   -> mel bands as contiguous runs of float4 weight groups, two threads per band -> log -> 512-byte output row.
#include <cmath>
#include <vector>
#include <cstring>
#include "a2m_common.cuh"
#include "fft_math.cuh"

void a2m_count_launch();

namespace {

constexpr int kC = 1024;                 // complex points = nfft / 2
constexpr int kNfftW = 2 * kC;
constexpr int kBinsW = kC + 1;
constexpr int kThreadsW = kC / 4;
constexpr int kMaxMelW = 128;
constexpr int kChunk = 8;                // consecutive frames per work item

struct WideGeom {
    int hop, n_mel, nnz, power, pad_mode, log_mode;
    float log_offset;
    long long n_samples, wav_stride, frames;
    int chunks_per_clip;
    long long n_items;
};

struct WideTables {
    const float* window;      // [2048]
    const float2* tw;         // [1024] exp(-2 pi i m / 1024)
    const float2* unt;        // [1024] exp(-2 pi i k / 2048)
    const int4* col_meta;     // [n_mel] {first bin (multiple of 4), 4-bin groups, offset into weights, 0}
    const float* weights;     // [nnz]
};

// transform buffers are padded by one point every 16: the scattered stores of the first two Stockham passes (stride 4
// and 16 points) then spread over all banks
__device__ __forceinline__ int padi(int i) { return i + (i >> 4); }
constexpr int kCPad = kC + kC / 16;

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}

__global__ void __launch_bounds__(kThreadsW)
melspec_wide_kernel(const float* __restrict__ wav, WideTables tab, WideGeom g, float* __restrict__ out) {
    // shared memory: transform ping-pong (16 KB), raw-sample ring (8 KB: consecutive frames share nfft - hop samples, only
    // the hop new ones are fetched), FFT twiddles (8 KB), bin powers, mel weights.  The window coefficients a thread
    // needs never change (fixed thread -> point mapping): eight registers.
    extern __shared__ __align__(16) unsigned char smem_w[];
    float2* s_a = reinterpret_cast<float2*>(smem_w);
    float2* s_b = s_a + kCPad;
    float* s_ring = reinterpret_cast<float*>(s_b + kCPad);
    float2* s_tw = reinterpret_cast<float2*>(s_ring + kNfftW);
    float* s_pw = reinterpret_cast<float*>(s_tw + kC);     // 1028 floats (bins 1025..1027 stay zero)
    float* s_wt = s_pw + 1028;
    const float2* __restrict__ g_unt = tab.unt;
    const int4* __restrict__ g_meta = tab.col_meta;
    const int tid = threadIdx.x;
    // twiddles regrouped per pass so that consecutive k are consecutive addresses: pass p holds W^{t k kC/(4p)} at
    // [p - 4 + (t - 1) * p + k], t = 1..3, k < p  (offsets 0, 12, 60, 252; 1020 entries)
    for (int p = 4; p < kC; p *= 4)
        for (int i = tid; i < 3 * p; i += kThreadsW) {
            const int t = i / p + 1, k = i - (t - 1) * p;
            s_tw[p - 4 + i] = tab.tw[t * k * (kC / (4 * p))];
        }
    for (int i = tid; i < g.nnz; i += kThreadsW) s_wt[i] = tab.weights[i];
    if (tid < 3) s_pw[kBinsW + tid] = 0.f;
    float2 win[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) win[t] = *reinterpret_cast<const float2*>(tab.window + 2 * (tid + kThreadsW * t));
    __syncthreads();

    const long long N = g.n_samples;
    const int lead = g.pad_mode ? kC : 0;                  // padded coordinate s_p = s + lead
    for (long long item = blockIdx.x; item < g.n_items; item += gridDim.x) {
        const long long clip = item / g.chunks_per_clip;
        const long long f0 = (item - clip * g.chunks_per_clip) * kChunk;
        const long long f1 = f0 + kChunk < g.frames ? f0 + kChunk : g.frames;
        const float* y = wav + clip * g.wav_stride;
        const bool can_prefetch = g.hop <= 2 * kThreadsW;
        float pre0 = 0.f, pre1 = 0.f;
        for (long long f = f0; f < f1; ++f) {
            // ---- bring the ring up to date: ring[s_p & 2047] = padded sample s_p for s_p in [f * hop, f * hop + 2048).
            // The hop new samples of frame f + 1 were fetched into registers while frame f was being transformed.
            const long long base = f * g.hop;
            auto sample = [&](long long sp) {
                long long sidx = sp - lead;
                if (g.pad_mode == 1) { if (sidx < 0) sidx = -sidx; if (sidx >= N) sidx = 2 * (N - 1) - sidx; }      // np.pad mode='reflect'
                return (sidx >= 0 && sidx < N) ? __ldg(y + sidx) : 0.f;                                     // zeros otherwise
            };
            if (f == f0 || !can_prefetch) {
                const long long fresh = (f == f0 || g.hop >= kNfftW) ? base : base + kNfftW - g.hop;
                for (long long sp = fresh + tid; sp < base + kNfftW; sp += kThreadsW) s_ring[sp & (kNfftW - 1)] = sample(sp);
            } else {
                const long long fresh = base + kNfftW - g.hop;
                if (tid < g.hop) s_ring[(fresh + tid) & (kNfftW - 1)] = pre0;
                if (tid + kThreadsW < g.hop) s_ring[(fresh + tid + kThreadsW) & (kNfftW - 1)] = pre1;
            }
            __syncthreads();
            if (can_prefetch && f + 1 < f1) {              // in flight during the transform below
                const long long nxt = base + kNfftW;       // first new sample of frame f + 1
                pre0 = tid < g.hop ? sample(nxt + tid) : 0.f;
                pre1 = tid + kThreadsW < g.hop ? sample(nxt + tid + kThreadsW) : 0.f;
            }
            // ---- frame x window, even/odd packed: z[n] = x[2n] + i x[2n+1]
            const int rb = static_cast<int>(base & (kNfftW - 1));
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int n = tid + kThreadsW * t;
                float2 x;
                if ((rb & 1) == 0) x = *reinterpret_cast<const float2*>(s_ring + ((rb + 2 * n) & (kNfftW - 1)));
                else x = make_float2(s_ring[(rb + 2 * n) & (kNfftW - 1)], s_ring[(rb + 2 * n + 1) & (kNfftW - 1)]);
                s_a[padi(n)] = make_float2(x.x * win[t].x, x.y * win[t].y);
            }
            __syncthreads();
            // ---- 1024-point complex FFT: radix-4 Stockham autosort, p = 1, 4, 16, 64, 256
            float2* src = s_a;
            float2* dst = s_b;
#pragma unroll
            for (int p = 1; p < kC; p *= 4) {
                const int k = tid & (p - 1);
                const int j = ((tid - k) << 2) + k;
                const float2 u0 = src[padi(tid)];
                float2 u1 = src[padi(tid + kThreadsW)], u2 = src[padi(tid + 2 * kThreadsW)], u3 = src[padi(tid + 3 * kThreadsW)];
                if (p > 1) {
                    const float2* tw = s_tw + (p - 4) + k;
                    u1 = cmul(u1, tw[0]); u2 = cmul(u2, tw[p]); u3 = cmul(u3, tw[2 * p]);
                }
                const float2 s02 = make_float2(u0.x + u2.x, u0.y + u2.y), d02 = make_float2(u0.x - u2.x, u0.y - u2.y);
                const float2 s13 = make_float2(u1.x + u3.x, u1.y + u3.y), d13 = make_float2(u1.x - u3.x, u1.y - u3.y);
                dst[padi(j)] = make_float2(s02.x + s13.x, s02.y + s13.y);
                dst[padi(j + p)] = make_float2(d02.x + d13.y, d02.y - d13.x);            // u0 - i u1 - u2 + i u3
                dst[padi(j + 2 * p)] = make_float2(s02.x - s13.x, s02.y - s13.y);
                dst[padi(j + 3 * p)] = make_float2(d02.x - d13.y, d02.y + d13.x);        // u0 + i u1 - u2 - i u3
                __syncthreads();
                float2* tmp = src; src = dst; dst = tmp;
            }
            // ---- untangle the real-input transform (Z = src), power per bin
            for (int k = tid; k <= kC / 2; k += kThreadsW) {
                const float2 zk = src[padi(k)], zm = src[padi((kC - k) & (kC - 1))];
                const float2 e = make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y));      // (Zk + conj Zm) / 2
                const float2 o = make_float2(0.5f * (zk.y + zm.y), 0.5f * (zm.x - zk.x));      // (Zk - conj Zm) / (2i)
                const float2 wo = cmul(k == 0 ? make_float2(1.f, 0.f) : __ldg(g_unt + k), o);
                const float2 xk = make_float2(e.x + wo.x, e.y + wo.y), xm = make_float2(e.x - wo.x, e.y - wo.y);   // X[k], conj X[1024-k]
                const float pk = fmaf(xk.x, xk.x, xk.y * xk.y), pm = fmaf(xm.x, xm.x, xm.y * xm.y);
                s_pw[k] = g.power == 2 ? pk : sqrtf(pk);
                s_pw[kC - k] = g.power == 2 ? pm : sqrtf(pm);
            }
            __syncthreads();
            // ---- mel bands: two threads per band over its 4-bin groups
            {
                const int c = tid >> 1, part = tid & 1;
                float acc0 = 0.f, acc1 = 0.f;
                if (c < g.n_mel) {
                    const int4 mt = __ldg(g_meta + c);
                    const float4* pg = reinterpret_cast<const float4*>(s_pw + mt.x);
                    const float4* w = reinterpret_cast<const float4*>(s_wt + mt.z);
                    for (int q = part; q < mt.y; q += 2) {
                        const float4 a = pg[q], b = w[q];
                        acc0 = fmaf(a.x, b.x, acc0); acc1 = fmaf(a.y, b.y, acc1); acc0 = fmaf(a.z, b.z, acc0); acc1 = fmaf(a.w, b.w, acc1);
                    }
                }
                float acc = acc0 + acc1;
                acc += __shfl_xor_sync(0xffffffffu, acc, 1);
                if (c < g.n_mel && part == 0) {
                    const float v = g.log_mode ? (acc == 0.f ? g.log_offset : acc) : acc + g.log_offset;
                    out[(clip * g.frames + f) * g.n_mel + c] = logf(v);
                }
            }
            // the next frame's ring update touches only samples this frame no longer reads after the barrier above, and
            // its packing pass is behind a barrier of its own; s_pw is rewritten only after four more barriers
        }
        __syncthreads();                                   // the next item refills the whole ring
    }
}


// =====================================================================================================
// melspec2048_kernel: the same transform on the register-radix / packed-f32x2 scheme of the 512-point kernel (logmel.cu).
//   * 64 threads own a PAIR of consecutive frames (A, B): every register holds the same quantity of both frames, every
//     butterfly is one FADD2 / FMUL2 / FFMA2, every twiddle / window / untangle constant serves two frames.
//   * 1024-point complex FFT of the even/odd-packed frame = 16 x 16 x 4: n = t + 64 m, k = k1 + 16 (q + 16 r)
//       pass A  thread t:        DFT16 over m, twiddle W1024^(t k1)                      -> exchange 1 (shared memory)
//       pass B  thread (k1, v):  DFT16 over u (t = 4 u + v), twiddle W64^(v q)           -> exchange 2
//       pass C  thread (k1, v'): DFT4 over v for q = 4 v' .. 4 v' + 3                    -> Z[k1 + 16 q + 256 r]
//     two exchanges through one padded 17 KB region per group (16-byte units, layouts chosen so that every quarter
//     warp hits eight different bank groups both ways) instead of five Stockham passes; group-level named barriers only.
//   * untangle of the bin pairs (j, 1024 - j), j = t + 64 i, from Z in the same region; |2 X|^2 (or |2 X|) goes back as
//     (A, B) pairs and each thread projects two mel bands (b and n_mel - 1 - b: the wide and the narrow triangles
//     balance), the factor 1/4 (1/2) folded into the weights.
// Four groups per CTA, two CTAs per SM; any hop, all padding modes.
// =====================================================================================================
namespace fast {

using a2m_fft::pair_t;
using a2m_fft::cpx;

constexpr int kGroupThreads = 64;
constexpr int kGroups = 4;
constexpr int kThreadsF = kGroupThreads * kGroups;
constexpr int kRow = 68;                                 // 16-byte units per exchange row (64 + 4: rows 4 bank groups apart)
constexpr int kRegionBytes = 16 * kRow * 16;             // 17 408 B per group
constexpr int kPwEntries = 1028;                         // bins 0..1024 + 3 zeros (the 4-bin groups of the mel runs)
static_assert(kPwEntries * 8 <= kRegionBytes && (1024 + 32) * 16 <= kRegionBytes, "Z and the powers alias the exchange region");

struct FastTab {
    const float2* win2;       // [1024] (w[2 n], w[2 n + 1])
    const float2* tw1;        // [16][64] exp(-2 pi i t k1 / 1024) at [k1][t]
    const float2* tw2;        // [16][4]  exp(-2 pi i v q / 64) at [q][v]
    const float2* unt;        // [516]    (-sin, -cos)(2 pi j / 2048), j = 0..512
    const float2* uv;         // [kPwPadded] per bin (padded index): weight into band g(k) ("rising"), into band g(k) - 1 ("falling"),
                              // x 1/4 (power) or 1/2 (magnitude): the kernel keeps |2 X|
    const int* seg;           // [n_mel + 2] first bin of segment g; band c = rising over segment c + falling over segment c + 1
    const int2* chunk;        // [kMaxChunks] {first bin, end bin}: the segments cut into pieces of <= 8 bins, in bin order
    const int* seg_chunk;     // [n_mel + 2] first chunk of segment g
    int n_chunks;
};
constexpr int kMaxChunks = 320;
constexpr int kChunkBins = 8;
constexpr int kPwPadded = 1030;                          // bins 0..1024 (+ slack: the projection reads pairs of bins)
__device__ __forceinline__ int pw_idx(int k) { return k; }

constexpr int kTablesBytes = (1024 + 1024 + 520 + kPwPadded + kMaxChunks) * 8 + 132 * 4;
constexpr int kSmemFast = kTablesBytes + kGroups * kRegionBytes;
static_assert(kTablesBytes % 16 == 0, "the exchange regions are accessed in 16-byte units");
static_assert(kPwPadded * 8 % 16 == 0 && kPwPadded * 8 + kMaxChunks * 16 <= kRegionBytes, "powers + chunk partials alias the exchange region");

__device__ __forceinline__ void group_sync(int grp) { a2m::named_barrier(1 + grp, kGroupThreads); }
__device__ __forceinline__ void st_cpx(unsigned char* base, int unit, cpx v) {
    *reinterpret_cast<ulonglong2*>(base + unit * 16) = make_ulonglong2(v.re, v.im);
}
__device__ __forceinline__ cpx ld_cpx(const unsigned char* base, int unit) {
    const ulonglong2 u = *reinterpret_cast<const ulonglong2*>(base + unit * 16);
    return a2m_fft::make(u.x, u.y);
}
__device__ __forceinline__ int z_unit(int k) { return k + 2 * (k >> 6); }      // Z[k]: two pad units per 64
// natural log of a positive normal number (the caller adds an offset or floors zeros first): one MUFU + one multiply
__device__ __forceinline__ float fast_log(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y * 0.69314718055994530942f;
}

__global__ void __launch_bounds__(kThreadsF, 2)
melspec2048_kernel(const float* __restrict__ wav, FastTab tab, WideGeom g, float* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_f[];
    float2* s_win = reinterpret_cast<float2*>(smem_f);
    float2* s_tw1 = s_win + 1024;
    float2* s_unt = s_tw1 + 1024;
    float2* s_uv = s_unt + 520;
    int2* s_chunk = reinterpret_cast<int2*>(s_uv + kPwPadded);
    int* s_seg = reinterpret_cast<int*>(s_chunk + kMaxChunks);           // first chunk of each segment
    const int tid = threadIdx.x, grp = tid >> 6, t = tid & 63;
    unsigned char* region = smem_f + kTablesBytes + grp * kRegionBytes;
    for (int i = tid; i < 1024; i += kThreadsF) { s_win[i] = tab.win2[i]; s_tw1[i] = tab.tw1[i]; }
    for (int i = tid; i < 516; i += kThreadsF) s_unt[i] = tab.unt[i];
    for (int i = tid; i < kPwPadded; i += kThreadsF) s_uv[i] = tab.uv[i];
    for (int i = tid; i < kMaxChunks; i += kThreadsF) s_chunk[i] = tab.chunk[i];
    for (int i = tid; i < g.n_mel + 2; i += kThreadsF) s_seg[i] = tab.seg_chunk[i];
    __syncthreads();

    const long long N = g.n_samples;
    const int lead = g.pad_mode ? kC : 0;
    const long long pairs_per_clip = (g.frames + 1) >> 1;
    const long long n_items = pairs_per_clip * (g.n_items / g.chunks_per_clip);      // clips x frame pairs
    const int k1b = t >> 2, vb = t & 3;                  // my (k1, v) of passes B and C

    for (long long item = static_cast<long long>(blockIdx.x) * kGroups + grp; item < n_items;
         item += static_cast<long long>(gridDim.x) * kGroups) {
        const long long clip = item / pairs_per_clip;
        const long long fa = (item - clip * pairs_per_clip) * 2, fb = fa + 1;
        const bool has_b = fb < g.frames;
        const float* y = wav + clip * g.wav_stride;
        auto sample = [&](long long sidx) {
            if (g.pad_mode == 1) { if (sidx < 0) sidx = -sidx; if (sidx >= N) sidx = 2 * (N - 1) - sidx; }      // np.pad mode='reflect'
            return (sidx >= 0 && sidx < N) ? __ldg(y + sidx) : 0.f;
        };
        // ---- load both frames x window, even/odd packed: z[n] = x[2 n] + i x[2 n + 1], n = t + 64 m
        cpx z[16];
        {
            const long long sa = fa * g.hop - lead, sb = fb * g.hop - lead;
            const bool interior = sa >= 0 && has_b && sb + kNfftW <= N;      // both frames inside the clip: no padding logic,
            if (interior) {                                                  // 64 independent loads in flight
                const float* pa = y + sa + 2 * t;
                const float* pb = y + sb + 2 * t;
                float ax[16], ay[16], bx[16], by[16];
#pragma unroll
                for (int m = 0; m < 16; ++m) {
                    ax[m] = __ldg(pa + 128 * m); ay[m] = __ldg(pa + 128 * m + 1);
                    bx[m] = __ldg(pb + 128 * m); by[m] = __ldg(pb + 128 * m + 1);
                }
#pragma unroll
                for (int m = 0; m < 16; ++m) {
                    const float2 w = s_win[t + 64 * m];
                    z[m] = a2m_fft::make(a2m_fft::mul2(a2m_fft::pack(ax[m], bx[m]), a2m_fft::bcast(w.x)),
                                         a2m_fft::mul2(a2m_fft::pack(ay[m], by[m]), a2m_fft::bcast(w.y)));
                }
            } else {                                     // frames that touch the padding (two at each end of a clip)
#pragma unroll
                for (int m = 0; m < 16; ++m) {
                    const int n = t + 64 * m;
                    const float ax = sample(sa + 2 * n), ay = sample(sa + 2 * n + 1);
                    const float bx = has_b ? sample(sb + 2 * n) : 0.f, by = has_b ? sample(sb + 2 * n + 1) : 0.f;
                    const float2 w = s_win[n];
                    z[m] = a2m_fft::make(a2m_fft::mul2(a2m_fft::pack(ax, bx), a2m_fft::bcast(w.x)),
                                         a2m_fft::mul2(a2m_fft::pack(ay, by), a2m_fft::bcast(w.y)));
                }
            }
        }
        // ---- pass A: DFT16 over m, twiddle W1024^(t k1)
        a2m_fft::dft16(z);
#pragma unroll
        for (int k1 = 1; k1 < 16; ++k1) {
            const float2 w = s_tw1[k1 * 64 + t];
            z[k1] = a2m_fft::mul_scalar(z[k1], w.x, w.y);
        }
        group_sync(grp);                                 // the previous pair's mel projection has read the region
#pragma unroll
        for (int k1 = 0; k1 < 16; ++k1) st_cpx(region, k1 * kRow + t, z[k1]);
        group_sync(grp);
#pragma unroll
        for (int u = 0; u < 16; ++u) z[u] = ld_cpx(region, k1b * kRow + 4 * u + vb);
        // ---- pass B: DFT16 over u, twiddle W64^(v q)
        a2m_fft::dft16(z);
        if (vb != 0) {
#pragma unroll
            for (int q = 1; q < 16; ++q) {
                const float2 w = __ldg(tab.tw2 + q * 4 + vb);
                z[q] = a2m_fft::mul_scalar(z[q], w.x, w.y);
            }
        }
        group_sync(grp);                                 // every exchange-1 read is done
#pragma unroll
        for (int q = 0; q < 16; ++q) st_cpx(region, k1b * kRow + (q >> 2) * 17 + (q & 3) * 4 + vb, z[q]);
        group_sync(grp);
        // ---- pass C: for q = 4 v' + qi, DFT4 over v -> Z[k1 + 16 q + 256 r] in z[4 qi + r]
#pragma unroll
        for (int qi = 0; qi < 4; ++qi) {
#pragma unroll
            for (int v = 0; v < 4; ++v) z[4 * qi + v] = ld_cpx(region, k1b * kRow + vb * 17 + qi * 4 + v);
            a2m_fft::dft4(z[4 * qi], z[4 * qi + 1], z[4 * qi + 2], z[4 * qi + 3]);
        }
        group_sync(grp);                                 // every exchange-2 read is done
#pragma unroll
        for (int qi = 0; qi < 4; ++qi)
#pragma unroll
            for (int r = 0; r < 4; ++r) st_cpx(region, z_unit(k1b + 16 * (4 * vb + qi) + 256 * r), z[4 * qi + r]);
        group_sync(grp);
        // ---- untangle the bin pairs (j, 1024 - j), j = t + 64 i (i = 0..7) and j = 512 (thread 0)
        pair_t sq_lo[9], sq_hi[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            const int j = i < 8 ? t + 64 * i : 512;
            sq_lo[i] = sq_hi[i] = a2m_fft::pack(0.f, 0.f);
            if (i < 8 || t == 0) {
                const cpx zk = ld_cpx(region, z_unit(j)), zp = ld_cpx(region, z_unit((kC - j) & (kC - 1)));
                const float2 tu = s_unt[j];
                a2m_fft::untangle_pair_sq(zk, zp, tu.x, tu.y, sq_lo[i], sq_hi[i]);
            }
        }
        group_sync(grp);                                 // every Z read is done: the region becomes the (A, B) powers per bin
        {
            float2* s_pw = reinterpret_cast<float2*>(region);
#pragma unroll
            for (int i = 0; i < 9; ++i) {
                const int j = i < 8 ? t + 64 * i : 512;
                if (i < 8 || t == 0) {
                    float2 a = make_float2(a2m_fft::lo(sq_lo[i]), a2m_fft::hi(sq_lo[i]));
                    float2 b = make_float2(a2m_fft::lo(sq_hi[i]), a2m_fft::hi(sq_hi[i]));
                    if (g.power != 2) { a.x = sqrtf(a.x); a.y = sqrtf(a.y); b.x = sqrtf(b.x); b.y = sqrtf(b.y); }
                    s_pw[pw_idx(j)] = a;
                    s_pw[pw_idx(kC - j)] = b;
                }
            }
        }
        group_sync(grp);
        // ---- mel projection in two steps.  (1) every thread takes chunks t, t + 64, ... of <= 8 consecutive bins of one
        // segment and sums both edges: R = sum u P (into band g), F = sum v P (into band g - 1); consecutive lanes read
        // consecutive bins, every thread has the same three or four chunks of work.  (2) band c = the R partials of
        // segment c + the F partials of segment c + 1, summed in chunk order (deterministic).
        {
            const float2* s_pw = reinterpret_cast<const float2*>(region);
            float4* s_part = reinterpret_cast<float4*>(region + kPwPadded * 8);
            for (int ci = t; ci < tab.n_chunks; ci += kGroupThreads) {
                const int2 ch = s_chunk[ci];
                pair_t r = a2m_fft::pack(0.f, 0.f), f = a2m_fft::pack(0.f, 0.f);
#pragma unroll
                for (int ii = 0; ii < kChunkBins; ++ii) {
                    const int k = ch.x + ((ii + t) & (kChunkBins - 1));       // rotated per lane: lanes 64 B apart spread over the banks
                    if (k < ch.y) {
                        const float2 pw = s_pw[k], uv = s_uv[k];
                        const pair_t pp = a2m_fft::pack(pw.x, pw.y);
                        r = a2m_fft::fma2(pp, a2m_fft::bcast(uv.x), r);
                        f = a2m_fft::fma2(pp, a2m_fft::bcast(uv.y), f);
                    }
                }
                s_part[ci] = make_float4(a2m_fft::lo(r), a2m_fft::hi(r), a2m_fft::lo(f), a2m_fft::hi(f));
            }
            group_sync(grp);
#pragma unroll
            for (int side = 0; side < 2; ++side) {
                const int c = t + 64 * side;
                if (c >= g.n_mel) continue;
                float va = 0.f, vbv = 0.f;
                const int c0 = s_seg[c], c1 = s_seg[c + 1], c2 = s_seg[c + 2];
                float va2 = 0.f, vb2 = 0.f;              // two chains: the wide bands have up to 14 partials per edge
                for (int ci = c0; ci + 1 < c1; ci += 2) {
                    const float4 p0 = s_part[ci], p1 = s_part[ci + 1];
                    va += p0.x; vbv += p0.y; va2 += p1.x; vb2 += p1.y;
                }
                if ((c1 - c0) & 1) { const float4 pp = s_part[c1 - 1]; va += pp.x; vbv += pp.y; }
                for (int ci = c1; ci + 1 < c2; ci += 2) {
                    const float4 p0 = s_part[ci], p1 = s_part[ci + 1];
                    va += p0.z; vbv += p0.w; va2 += p1.z; vb2 += p1.w;
                }
                if ((c2 - c1) & 1) { const float4 pp = s_part[c2 - 1]; va += pp.z; vbv += pp.w; }
                va += va2; vbv += vb2;
                const float la = g.log_mode ? (va == 0.f ? g.log_offset : va) : va + g.log_offset;
                const float lb = g.log_mode ? (vbv == 0.f ? g.log_offset : vbv) : vbv + g.log_offset;
                out[(clip * g.frames + fa) * g.n_mel + c] = fast_log(la);
                if (has_b) out[(clip * g.frames + fb) * g.n_mel + c] = fast_log(lb);
            }
        }
    }
}

}  // namespace fast

constexpr int kSmemFixedWide = (2 * kCPad) * 8 + kNfftW * 4 + kC * 8 + 1028 * 4;      // + 4 * nnz mel weights

}  // namespace

struct a2m_melspec_plan {
    int device, nfft, hop, n_mel, nnz, power, pad_mode, log_mode;
    float log_offset;
    void* blob;
    WideTables tab;
    fast::FastTab ftab;
    int fast_ok;              // the filterbank has the segment structure melspec2048_kernel projects with
};

extern "C" int a2m_melspec_plan_create(int nfft, int hop, int n_mel, int power, int pad_mode, const double* window_host,
                                       const double* mel_weights_host, double log_offset, int log_mode, int device,
                                       a2m_melspec_plan** out) {
    A2M_ARG_CHECK(out != nullptr, "a2m_melspec_plan_create: out is NULL");
    *out = nullptr;
    if (nfft != kNfftW) {
        a2m_set_error("a2m_melspec_plan_create: fft length %d not supported (this kernel implements nfft = %d; "
                      "nfft = 512 is a2m_mel_plan_create)", nfft, kNfftW);
        return A2M_ERR_UNSUPPORTED;
    }
    A2M_ARG_CHECK(hop >= 1, "a2m_melspec_plan_create: hop %d must be >= 1", hop);
    A2M_ARG_CHECK(n_mel >= 1 && n_mel <= kMaxMelW, "a2m_melspec_plan_create: n_mel %d must be in [1, %d]", n_mel, kMaxMelW);
    A2M_ARG_CHECK(power == 1 || power == 2, "a2m_melspec_plan_create: power %d (1 = magnitude, 2 = power)", power);
    A2M_ARG_CHECK(pad_mode >= A2M_PAD_NONE && pad_mode <= A2M_PAD_ZEROS, "a2m_melspec_plan_create: pad_mode %d", pad_mode);
    A2M_ARG_CHECK(log_mode == A2M_LOG_ADD_OFFSET || log_mode == A2M_LOG_FLOOR_ZEROS, "a2m_melspec_plan_create: log_mode %d", log_mode);
    A2M_ARG_CHECK(window_host && mel_weights_host, "a2m_melspec_plan_create: NULL table");

    // column-compressed mel matrix: every band's support is one contiguous run of bins, padded to whole 4-bin groups
    std::vector<int4> meta(kMaxMelW, make_int4(0, 0, 0, 0));
    std::vector<float> weights;
    for (int c = 0; c < n_mel; ++c) {
        int first = -1, last = -1;
        for (int k = 0; k < kBinsW; ++k)
            if (mel_weights_host[static_cast<size_t>(k) * n_mel + c] != 0.0) {
                if (first >= 0 && k != last + 1) {
                    a2m_set_error("a2m_melspec_plan_create: mel column %d is not one contiguous run of bins", c);
                    return A2M_ERR_UNSUPPORTED;
                }
                if (first < 0) first = k;
                last = k;
            }
        if (first < 0) continue;                            // empty band: log of the offset / floor
        const int lo = first & ~3, hi = last | 3;           // hi <= 1027: the three bins past Nyquist are zero in shared memory
        meta[c] = make_int4(lo, (hi - lo + 1) / 4, static_cast<int>(weights.size()), 0);
        for (int k = lo; k <= hi; ++k)
            weights.push_back(k < kBinsW ? static_cast<float>(mel_weights_host[static_cast<size_t>(k) * n_mel + c]) : 0.f);
    }
    const int nnz = static_cast<int>(weights.size());
    std::vector<float> win(kNfftW);
    for (int i = 0; i < kNfftW; ++i) win[i] = static_cast<float>(window_host[i]);
    std::vector<float2> tw(kC), unt(kC);
    const double two_pi = 6.283185307179586476925286766559;
    for (int m = 0; m < kC; ++m) {
        tw[m] = make_float2(static_cast<float>(std::cos(two_pi * m / kC)), static_cast<float>(-std::sin(two_pi * m / kC)));
        unt[m] = make_float2(static_cast<float>(std::cos(two_pi * m / kNfftW)), static_cast<float>(-std::sin(two_pi * m / kNfftW)));
    }
    A2M_CUDA_CHECK(cudaSetDevice(device));
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~size_t(255); return o; };
    const size_t o_win = carve(kNfftW * 4), o_tw = carve(kC * 8), o_unt = carve(kC * 8), o_cm = carve(kMaxMelW * 16),
                 o_wt = carve((nnz + 4) * 4);
    // tables of the register-radix kernel
    const size_t o_tw1 = carve(1024 * 8), o_tw2 = carve(64 * 8), o_unt2 = carve(516 * 8), o_uv = carve(fast::kPwPadded * 8),
                 o_seg = carve(132 * 4), o_chunk = carve(fast::kMaxChunks * 8), o_sc = carve(132 * 4);
    std::vector<float2> tw1(1024), tw2(64), unt2(516, make_float2(0.f, 0.f));
    for (int k1 = 0; k1 < 16; ++k1)
        for (int t = 0; t < 64; ++t) {
            const double a = two_pi * ((t * k1) % 1024) / 1024.0;
            tw1[k1 * 64 + t] = make_float2(static_cast<float>(std::cos(a)), static_cast<float>(-std::sin(a)));
        }
    for (int q = 0; q < 16; ++q)
        for (int v = 0; v < 4; ++v) {
            const double a = two_pi * ((v * q) % 64) / 64.0;
            tw2[q * 4 + v] = make_float2(static_cast<float>(std::cos(a)), static_cast<float>(-std::sin(a)));
        }
    for (int j = 0; j <= 512; ++j) {
        const double a = two_pi * j / 2048.0;
        unt2[j] = make_float2(static_cast<float>(-std::sin(a)), static_cast<float>(-std::cos(a)));
    }
    // the filterbank as segments: bin k of segment g feeds band g ("rising", weight u) and band g - 1 ("falling", weight v).
    // True of every triangular bank (Slaney, HTK); a matrix that is not of that shape runs the generic kernel.
    std::vector<float2> uv(fast::kPwPadded, make_float2(0.f, 0.f));
    std::vector<int> seg(132, kBinsW);
    bool fast_ok = true;
    {
        const float fold = power == 2 ? 0.25f : 0.5f;       // the kernel keeps |2 X|
        std::vector<int> peak(n_mel, -1);
        for (int c = 0; c < n_mel; ++c) {
            double best = 0.0;
            for (int k = 0; k < kBinsW; ++k) {
                const double w = mel_weights_host[static_cast<size_t>(k) * n_mel + c];
                if (w > best) { best = w; peak[c] = k; }
            }
        }
        std::vector<int> gk(kBinsW, 0);
        int prev = 0;
        for (int k = 0; k < kBinsW && fast_ok; ++k) {
            int first = -1, count = 0;
            for (int c = 0; c < n_mel; ++c)
                if (mel_weights_host[static_cast<size_t>(k) * n_mel + c] != 0.0) { if (first < 0) first = c; ++count; }
            int gseg = prev;
            if (count == 2) {
                if (mel_weights_host[static_cast<size_t>(k) * n_mel + first + 1] == 0.0) fast_ok = false;     // not adjacent
                gseg = first + 1;
            } else if (count == 1) {
                gseg = k <= peak[first] ? first : first + 1;
            } else if (count > 2) {
                fast_ok = false;
            }
            if (gseg < prev) fast_ok = false;
            gk[k] = gseg;
            prev = gseg;
            const int e = k;
            const double u = gseg < n_mel ? mel_weights_host[static_cast<size_t>(k) * n_mel + gseg] : 0.0;
            const double v = gseg >= 1 ? mel_weights_host[static_cast<size_t>(k) * n_mel + gseg - 1] : 0.0;
            uv[e] = make_float2(static_cast<float>(u) * fold, static_cast<float>(v) * fold);
        }
        for (int gi = 0; gi <= n_mel + 1; ++gi) {           // first bin with g(k) >= gi
            int k = 0;
            while (k < kBinsW && gk[k] < gi) ++k;
            seg[gi] = k;
        }
    }
    std::vector<int2> chunks;
    std::vector<int> seg_chunk(132, 0);
    if (fast_ok) {
        for (int gi = 0; gi <= n_mel; ++gi) {
            seg_chunk[gi] = static_cast<int>(chunks.size());
            for (int k = seg[gi]; k < seg[gi + 1]; k += fast::kChunkBins)
                chunks.push_back(make_int2(k, k + fast::kChunkBins < seg[gi + 1] ? k + fast::kChunkBins : seg[gi + 1]));
        }
        for (int gi = n_mel + 1; gi < 132; ++gi) seg_chunk[gi] = static_cast<int>(chunks.size());
        if (chunks.size() > static_cast<size_t>(fast::kMaxChunks)) fast_ok = false;
    }
    const int n_chunks = static_cast<int>(chunks.size());
    chunks.resize(fast::kMaxChunks, make_int2(0, 0));
    unsigned char* blob = nullptr;
    A2M_CUDA_CHECK(cudaMalloc(&blob, off));
    std::vector<unsigned char> host(off, 0);
    memcpy(host.data() + o_win, win.data(), kNfftW * 4);
    memcpy(host.data() + o_tw, tw.data(), kC * 8);
    memcpy(host.data() + o_unt, unt.data(), kC * 8);
    memcpy(host.data() + o_cm, meta.data(), kMaxMelW * 16);
    if (nnz) memcpy(host.data() + o_wt, weights.data(), nnz * 4);
    memcpy(host.data() + o_tw1, tw1.data(), 1024 * 8);
    memcpy(host.data() + o_tw2, tw2.data(), 64 * 8);
    memcpy(host.data() + o_unt2, unt2.data(), 516 * 8);
    memcpy(host.data() + o_uv, uv.data(), fast::kPwPadded * 8);
    memcpy(host.data() + o_seg, seg.data(), 132 * 4);
    memcpy(host.data() + o_chunk, chunks.data(), fast::kMaxChunks * 8);
    memcpy(host.data() + o_sc, seg_chunk.data(), 132 * 4);
    cudaError_t e = cudaMemcpy(blob, host.data(), off, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(blob); a2m_set_error("a2m_melspec_plan_create: upload failed: %s", cudaGetErrorString(e)); return (int)e; }
    a2m_melspec_plan* p = new a2m_melspec_plan();
    p->device = device; p->nfft = nfft; p->hop = hop; p->n_mel = n_mel; p->nnz = nnz; p->power = power; p->pad_mode = pad_mode;
    p->log_mode = log_mode; p->log_offset = static_cast<float>(log_offset); p->blob = blob;
    p->tab.window = reinterpret_cast<const float*>(blob + o_win);
    p->tab.tw = reinterpret_cast<const float2*>(blob + o_tw);
    p->tab.unt = reinterpret_cast<const float2*>(blob + o_unt);
    p->tab.col_meta = reinterpret_cast<const int4*>(blob + o_cm);
    p->tab.weights = reinterpret_cast<const float*>(blob + o_wt);
    p->ftab.win2 = reinterpret_cast<const float2*>(blob + o_win);         // (w[2 n], w[2 n + 1]) = the window read as float2
    p->ftab.tw1 = reinterpret_cast<const float2*>(blob + o_tw1);
    p->ftab.tw2 = reinterpret_cast<const float2*>(blob + o_tw2);
    p->ftab.unt = reinterpret_cast<const float2*>(blob + o_unt2);
    p->ftab.uv = reinterpret_cast<const float2*>(blob + o_uv);
    p->ftab.seg = reinterpret_cast<const int*>(blob + o_seg);
    p->ftab.chunk = reinterpret_cast<const int2*>(blob + o_chunk);
    p->ftab.seg_chunk = reinterpret_cast<const int*>(blob + o_sc);
    p->ftab.n_chunks = n_chunks;
    p->fast_ok = fast_ok ? 1 : 0;
    *out = p;
    return A2M_OK;
}

extern "C" void a2m_melspec_plan_destroy(a2m_melspec_plan* plan) {
    if (!plan) return;
    cudaFree(plan->blob);
    delete plan;
}

extern "C" int64_t a2m_melspec_num_frames(const a2m_melspec_plan* plan, int64_t n_samples) {
    if (!plan || n_samples < 0) return -1;
    if (plan->pad_mode != A2M_PAD_NONE) return 1 + n_samples / plan->hop;       // centred: padded length n + nfft
    if (n_samples < plan->nfft) return -1;
    return 1 + (n_samples - plan->nfft) / plan->hop;
}

extern "C" int a2m_melspec_f32(const a2m_melspec_plan* plan, const float* wav, int64_t n_clips, int64_t n_samples,
                               int64_t wav_stride, float* out, void* stream) {
    A2M_ARG_CHECK(plan != nullptr, "a2m_melspec_f32: plan is NULL");
    A2M_ARG_CHECK(n_clips >= 0 && n_samples >= 0, "a2m_melspec_f32: negative size");
    A2M_ARG_CHECK(n_clips <= 1 || wav_stride >= n_samples, "a2m_melspec_f32: wav_stride %lld < n_samples %lld",
                  (long long)wav_stride, (long long)n_samples);
    const int64_t frames = a2m_melspec_num_frames(plan, n_samples);
    A2M_ARG_CHECK(frames >= 1, "a2m_melspec_f32: %lld samples are fewer than one %d-sample frame", (long long)n_samples, plan->nfft);
    A2M_ARG_CHECK(plan->pad_mode != A2M_PAD_REFLECT || n_samples > plan->nfft / 2,
                  "a2m_melspec_f32: reflect padding needs more than %d samples, got %lld", plan->nfft / 2, (long long)n_samples);
    if (n_clips == 0) return A2M_OK;
    A2M_ARG_CHECK(wav != nullptr && out != nullptr, "a2m_melspec_f32: NULL buffer");
    WideGeom g;
    g.hop = plan->hop; g.n_mel = plan->n_mel; g.nnz = plan->nnz; g.power = plan->power; g.pad_mode = plan->pad_mode;
    g.log_mode = plan->log_mode; g.log_offset = plan->log_offset;
    g.n_samples = n_samples; g.wav_stride = wav_stride; g.frames = frames;
    g.chunks_per_clip = static_cast<int>((frames + kChunk - 1) / kChunk);
    g.n_items = static_cast<long long>(g.chunks_per_clip) * n_clips;
    if (!plan->fast_ok) {                                  // a filterbank that is not made of adjacent triangles: generic kernel
        const int smem = kSmemFixedWide + 4 * ((plan->nnz + 3) & ~3);
        A2M_ARG_CHECK(smem <= 100 * 1024, "a2m_melspec_f32: %d bytes of shared memory", smem);
        static A2mPerDeviceOnce wide_attr_set;
        if (wide_attr_set.first())
            A2M_CUDA_CHECK(cudaFuncSetAttribute(melspec_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        long long wgrid = 4LL * a2m_num_sms();
        if (wgrid > g.n_items) wgrid = g.n_items;
        melspec_wide_kernel<<<static_cast<unsigned>(wgrid), kThreadsW, smem, static_cast<cudaStream_t>(stream)>>>(
            wav, plan->tab, g, out);
        a2m_count_launch();
        A2M_LAUNCH_CHECK();
        return A2M_OK;
    }
    static A2mPerDeviceOnce attr_set;
    if (attr_set.first())
        A2M_CUDA_CHECK(cudaFuncSetAttribute(fast::melspec2048_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, fast::kSmemFast));
    const long long pairs = (frames + 1) / 2 * n_clips;
    long long grid = 2LL * a2m_num_sms();                   // two CTAs per SM, four frame pairs per CTA and turn
    if (grid * fast::kGroups > pairs) grid = (pairs + fast::kGroups - 1) / fast::kGroups;
    fast::melspec2048_kernel<<<static_cast<unsigned>(grid), fast::kThreadsF, fast::kSmemFast, static_cast<cudaStream_t>(stream)>>>(
        wav, plan->ftab, g, out);
    a2m_count_launch();
    A2M_LAUNCH_CHECK();
    return A2M_OK;
}
