// Fused log-mel front end: framing -> periodic Hann -> real FFT -> |.| -> sparse mel projection ->
// log(. + offset), one launch for a whole batch of clips.  Replaces pose_video/mel_features.py:192-223 (see
// include/a2m_b200.h).  HBM-bound by contract: every sample is read once (the overlap between tiles is an L2 hit) and
// every output element written once; everything in between lives in registers / shared memory.
//
// Two kernels:
//
//  logmel512_kernel  (fft length 512, the hot path's 25 ms window at 16 kHz)
//    * a WARP is an independent pipeline over tiles of 4 consecutive frames of one clip; there is no CTA-wide barrier
//      in the main loop.  Lane 0 stages a tile's samples with ONE bulk asynchronous copy (cp.async.bulk, 16-byte
//      aligned: the span start is aligned down and the reads are offset; a head / tail that would leave the caller's
//      buffer is copied by hand) into the warp's own buffer, completing on the warp's mbarrier; the next tile's copy is
//      issued as soon as the current tile's samples are in registers, so it overlaps the whole transform.
//    * 16 lanes own a PAIR of frames (A, B): every register holds the same quantity of both frames and every
//      butterfly is one packed fp32 instruction (FADD2 / FMUL2 / FFMA2 -- half the issue slots of scalar code, and
//      every twiddle / window / untangle constant fetched from shared memory serves two frames).
//    * 256-point complex FFT of the even/odd-packed frame = two register-resident radix-16 passes with one exchange
//      through padded shared memory (16-byte accesses, conflict-free both ways).
//    * the real-input untangle needs Z[256 - k] next to Z[k]: lane l and lane 16 - l swap HALF of their registers
//      with warp shuffles and each computes both |X[k]| and |X[256 - k]| from one (A, B) pair -- no spectrum dump.
//    * magnitudes of the pair go to shared memory once ([bin] -> (A, B)); lane l sums bands l, l + 16, ... with one
//      16-byte load per two bins (both frames) and writes log(mel + offset) straight to global memory.
//    fp32 and int16 PCM input (the reference accepts any real dtype; int16 halves the HBM / PCIe bytes per sample).
//
//  logmel_generic_kernel  (any power-of-two fft length 64 .. 4096; also |STFT| only)
//    one CTA per frame, radix-2 Stockham FFT of the even/odd-packed frame in shared memory.  Correct for every
//    geometry log_mel_spectrogram can ask for (its 8 kHz default -> 256, 22.05 kHz -> 1024, 44.1 / 48 kHz -> 2048);
//    not tuned.
#include <cmath>
#include <vector>
#include <cstring>
#include "a2m_common.cuh"
#include "fft_math.cuh"

using a2m_fft::cpx;
using a2m_fft::pair_t;

void a2m_count_launch();

namespace {

constexpr int kMaxMel = 128;
constexpr int kFastNfft = 512;
constexpr int kWarps = 8;                             // warps per CTA of the 512-point kernel
constexpr int kTileFrames = 4;                        // frames per warp tile: two 16-lane groups x (A, B)
constexpr int kXchgRow = 17;                          // 16-byte units per exchange row (16 + 1 pad)
constexpr int kXchgGroupBytes = 16 * kXchgRow * 16;   // 4352 B per 16-lane group; later the group's magnitudes
constexpr int kMagEntries = 258;                      // (A, B) magnitudes of bins 0..256 (+1 pad) alias the exchange rows
static_assert(kMagEntries * 8 <= kXchgGroupBytes, "magnitudes must fit the exchange region");

struct FastTables {               // device-resident constants of a plan (512-point kernel)
    const float* window;          // [512]  window, zero padded past the window length
    const float2* tw;             // [16][16] exp(-2 pi i m2 k1 / 256) at [k1][m2]
    const float2* untangle;       // [256]  (-sin, -cos)(2 pi k / 512)
    const int4* col_meta;         // [n_mel] {first bin (even), bin pairs, offset into weights (even), 0}
    const float* weights;         // [nnz]  0.5 * mel weight (the kernel keeps |2 X|)
};

struct FastGeom {
    int window, hop, n_mel, nnz;
    int n_m1;                     // rows of the 16 x 16 input that are not all zero padding: ceil(window / 32)
    int span_bytes;               // per-warp sample buffer (multiple of 16)
    int tables_bytes;             // offset of the first warp's private region
    float log_offset;
    int log_mode;                 // 0: log(x + offset) (mel_features.py:223); 1: log(x == 0 ? offset : x) (pats/data_loading/audio.py:117-119)
};

__device__ __forceinline__ float fast_sqrt(float x) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ void wait_or_trap(uint64_t* bar, uint32_t parity) {
    const long long t0 = clock64();
    while (!a2m::mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) asm volatile("trap;");     // a pipeline bug must never hang the GPU box
    }
}

template <typename InT> __device__ __forceinline__ float to_f32(InT v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<int16_t>(int16_t v) { return static_cast<float>(v); }

// samples (x[i], x[i + 1]) of a staged span; `aligned`: i is even (fp32: 8-byte aligned, int16: 4-byte aligned)
template <typename InT>
__device__ __forceinline__ float2 load_pair(const InT* __restrict__ s, int i, bool aligned);
template <>
__device__ __forceinline__ float2 load_pair<float>(const float* __restrict__ s, int i, bool aligned) {
    if (aligned) return *reinterpret_cast<const float2*>(s + i);
    return make_float2(s[i], s[i + 1]);
}
template <>
__device__ __forceinline__ float2 load_pair<int16_t>(const int16_t* __restrict__ s, int i, bool aligned) {
    if (aligned) {
        const uint32_t u = *reinterpret_cast<const uint32_t*>(s + i);
        return make_float2(static_cast<float>(static_cast<int16_t>(u & 0xffffu)), static_cast<float>(static_cast<int16_t>(u >> 16)));
    }
    return make_float2(static_cast<float>(s[i]), static_cast<float>(s[i + 1]));
}

template <typename InT>
__global__ void __launch_bounds__(kWarps * 32, 2)
logmel512_kernel(const InT* __restrict__ wav, long long wav_stride, long long n_samples, long long n_clips,
                 int frames_per_clip, int tiles_per_clip, long long n_tiles, FastTables tab, FastGeom g,
                 float* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem[];
    float* s_window = reinterpret_cast<float*>(smem);
    float2* s_tw = reinterpret_cast<float2*>(smem + 2048);
    float2* s_unt = reinterpret_cast<float2*>(smem + 4096);
    int4* s_meta = reinterpret_cast<int4*>(smem + 6144);
    float* s_weights = reinterpret_cast<float*>(smem + 8192);

    const int tid = threadIdx.x;
    for (int i = tid; i < kFastNfft; i += kWarps * 32) s_window[i] = tab.window[i];
    for (int i = tid; i < 256; i += kWarps * 32) { s_tw[i] = tab.tw[i]; s_unt[i] = tab.untangle[i]; }
    for (int i = tid; i < g.n_mel; i += kWarps * 32) s_meta[i] = tab.col_meta[i];
    for (int i = tid; i < g.nnz; i += kWarps * 32) s_weights[i] = tab.weights[i];

    const int warp = tid >> 5, lane = tid & 31;
    const int l = lane & 15, grp = lane >> 4;
    const int warp_bytes = 2 * kXchgGroupBytes + g.span_bytes + 16;
    unsigned char* wbase = smem + g.tables_bytes + warp * warp_bytes;
    unsigned char* xchg = wbase + grp * kXchgGroupBytes;                     // this group's exchange rows / magnitudes
    unsigned char* s_samples = wbase + 2 * kXchgGroupBytes;
    uint64_t* bar = reinterpret_cast<uint64_t*>(s_samples + g.span_bytes);
    if (lane == 0) {
        a2m::mbar_init(bar, 1);
        a2m::mbar_fence_init();
    }
    // never-written shared memory may hold NaN bit patterns; reads past a tile's span are multiplied by window zeros
    for (int i = lane; i < g.span_bytes / 4; i += 32) reinterpret_cast<uint32_t*>(s_samples)[i] = 0u;
    __syncthreads();                                                         // tables and barriers are in place

    const uintptr_t buf_lo = (reinterpret_cast<uintptr_t>(wav) + 15) & ~uintptr_t(15);          // 16-byte blocks fully
    const uintptr_t buf_hi = reinterpret_cast<uintptr_t>(wav + (n_clips - 1) * wav_stride + n_samples) & ~uintptr_t(15);  // inside the caller's buffer

    // Stage the samples of one tile: element i of the tile's span lands at byte (src & 15) + i * sizeof(InT) of the
    // warp's buffer.  Whole warp; lane 0 issues the bulk copy.
    auto stage = [&](long long tile) {
        const unsigned clip = static_cast<unsigned>(tile) / static_cast<unsigned>(tiles_per_clip);         // n_tiles < 2^31
        const int f0 = static_cast<int>(static_cast<unsigned>(tile) - clip * tiles_per_clip) * kTileFrames;
        const int nf = min(kTileFrames, frames_per_clip - f0);
        const InT* src = wav + static_cast<long long>(clip) * wav_stride + static_cast<long long>(f0) * g.hop;
        const uintptr_t a = reinterpret_cast<uintptr_t>(src);
        const uintptr_t a_end = a + (static_cast<uintptr_t>(nf - 1) * g.hop + g.window) * sizeof(InT);
        const uintptr_t a0 = a & ~uintptr_t(15);
        uintptr_t lo = a0, hi = (a_end + 15) & ~uintptr_t(15);
        if (lo < buf_lo) lo = buf_lo;
        if (hi > buf_hi) hi = buf_hi;
        const bool bulk = hi > lo;
        if (lane == 0) {
            a2m::fence_proxy_async_smem();                  // the warp's reads of the previous tile precede the async writes
            a2m::mbar_expect_tx(bar, bulk ? static_cast<uint32_t>(hi - lo) : 0u);
            if (bulk) a2m::bulk_load_1d(s_samples + (lo - a0), reinterpret_cast<const void*>(lo), static_cast<uint32_t>(hi - lo), bar);
        }
        if (!bulk) { lo = a_end; hi = a_end; }             // everything by hand
        for (uintptr_t p = a + lane * sizeof(InT); p < lo && p < a_end; p += 32 * sizeof(InT))          // head
            *reinterpret_cast<InT*>(s_samples + (p - a0)) = *reinterpret_cast<const InT*>(p);
        for (uintptr_t p = (hi > a ? hi : a) + lane * sizeof(InT); p < a_end; p += 32 * sizeof(InT))     // tail
            *reinterpret_cast<InT*>(s_samples + (p - a0)) = *reinterpret_cast<const InT*>(p);
    };

    const long long warp_global = static_cast<long long>(blockIdx.x) * kWarps + warp;
    const long long n_warps = static_cast<long long>(gridDim.x) * kWarps;
    long long tile = warp_global;
    if (tile < n_tiles) stage(tile);
    uint32_t phase = 0;
    const ulonglong2* xq = reinterpret_cast<const ulonglong2*>(xchg);
    ulonglong2* xw = reinterpret_cast<ulonglong2*>(xchg);
    float2* mag = reinterpret_cast<float2*>(xchg);
    const int pl = (16 - l) & 15;                                            // lane holding Z[256 - k] of this lane's k
    const bool lane0 = l == 0;

    for (; tile < n_tiles; tile += n_warps) {
        const unsigned clip = static_cast<unsigned>(tile) / static_cast<unsigned>(tiles_per_clip);
        const int f0 = static_cast<int>(static_cast<unsigned>(tile) - clip * tiles_per_clip) * kTileFrames;
        const InT* src = wav + static_cast<long long>(clip) * wav_stride + static_cast<long long>(f0) * g.hop;
        const int shift = static_cast<int>((reinterpret_cast<uintptr_t>(src) & 15) / sizeof(InT));
        wait_or_trap(bar, phase);
        phase ^= 1;
        __syncwarp();                                                        // hand-copied head / tail of the other lanes

        // ---- samples * window -> z[16 m1 + l] of frames A = f0 + 2 grp and B = A + 1 ---------------------------
        cpx v[16];
        {
            const InT* s = reinterpret_cast<const InT*>(s_samples);
            const int rel_a = shift + 2 * grp * g.hop, rel_b = rel_a + g.hop;
            const bool al_a = (rel_a & 1) == 0, al_b = (rel_b & 1) == 0;     // warp-uniform
            const int last = g.n_m1 - 1;
#pragma unroll
            for (int m1 = 0; m1 < 16; ++m1) {
                if (m1 <= last) {                                            // uniform
                    const int n = 32 * m1 + 2 * l;
                    const float2 w = *reinterpret_cast<const float2*>(s_window + n);
                    // past the window (last row) the buffer holds samples of later frames or of an earlier tile: finite
                    // values (the buffer is zeroed once, then only ever holds copied input), times the zero padding
                    const float2 xa = load_pair<InT>(s, rel_a + n, al_a);
                    const float2 xb = load_pair<InT>(s, rel_b + n, al_b);
                    v[m1] = a2m_fft::make(a2m_fft::pack(xa.x * w.x, xb.x * w.x), a2m_fft::pack(xa.y * w.y, xb.y * w.y));
                } else {
                    v[m1] = a2m_fft::make(a2m_fft::pack(0.f, 0.f), a2m_fft::pack(0.f, 0.f));
                }
            }
        }
        __syncwarp();                                                        // every lane has its samples in registers
        if (tile + n_warps < n_tiles) stage(tile + n_warps);                 // overlaps everything below

        // ---- pass 1: DFT16 over m1, twiddle W256^(l k1), exchange [k1][m2] ----------------------------------------
        a2m_fft::dft16(v);
#pragma unroll
        for (int k1 = 1; k1 < 16; ++k1) {
            const float2 w = s_tw[k1 * 16 + l];
            v[k1] = a2m_fft::mul_scalar(v[k1], w.x, w.y);
        }
#pragma unroll
        for (int k1 = 0; k1 < 16; ++k1) xw[k1 * kXchgRow + l] = make_ulonglong2(v[k1].re, v[k1].im);
        __syncwarp();
#pragma unroll
        for (int m2 = 0; m2 < 16; ++m2) {
            const ulonglong2 q = xq[l * kXchgRow + m2];
            v[m2] = a2m_fft::make(q.x, q.y);
        }
        __syncwarp();                                                        // the rows are overwritten by magnitudes below

        // ---- pass 2: DFT16 over m2 -> v[k2] = Z[l + 16 k2] --------------------------------------------------------
        a2m_fft::dft16(v);

        // ---- untangle: lane l and lane 16 - l swap registers 8..15; |2 X[k]| and |2 X[256 - k]| per pair -----------
        {   // bin 128 pairs with itself: |X[128]| = |Z[128]| (register 8 of lane 0)
            const pair_t sq = a2m_fft::fma2(v[8].im, v[8].im, a2m_fft::mul2(v[8].re, v[8].re));
            if (lane0) {
                mag[128] = make_float2(2.f * fast_sqrt(a2m_fft::lo(sq)), 2.f * fast_sqrt(a2m_fft::hi(sq)));
                mag[257] = make_float2(0.f, 0.f);
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int i = 15 - j;                                            // the register the partner hands over
            // lane 0 pairs with itself one register further on (bin 16 j with bin 256 - 16 j = 16 (16 - j))
            cpx snd;
            if (i == 15) {
                snd.re = lane0 ? v[0].re : v[15].re;
                snd.im = lane0 ? v[0].im : v[15].im;
            } else {
                snd.re = lane0 ? v[i + 1].re : v[i].re;
                snd.im = lane0 ? v[i + 1].im : v[i].im;
            }
            cpx zp;
            zp.re = __shfl_sync(0xffffffffu, snd.re, pl, 16);
            zp.im = __shfl_sync(0xffffffffu, snd.im, pl, 16);
            const int k = l + 16 * j;
            const float2 t = s_unt[k];
            pair_t sq_k, sq_m;
            a2m_fft::untangle_pair_sq(v[j], zp, t.x, t.y, sq_k, sq_m);
            mag[k] = make_float2(fast_sqrt(a2m_fft::lo(sq_k)), fast_sqrt(a2m_fft::hi(sq_k)));
            mag[256 - k] = make_float2(fast_sqrt(a2m_fft::lo(sq_m)), fast_sqrt(a2m_fft::hi(sq_m)));
        }
        __syncwarp();

        // ---- mel bands l, l + 16, ...: one 16-byte load = two bins of both frames ----------------------------------
        {
            const int fa = f0 + 2 * grp;
            float* row_a = out + (static_cast<long long>(clip) * frames_per_clip + fa) * g.n_mel;
            const bool has_a = fa < frames_per_clip, has_b = fa + 1 < frames_per_clip;
            for (int c = l; c < g.n_mel; c += 16) {
                const int4 m = s_meta[c];
                const ulonglong2* mg = reinterpret_cast<const ulonglong2*>(mag + m.x);
                const float2* w = reinterpret_cast<const float2*>(s_weights + m.z);
                pair_t acc = a2m_fft::pack(0.f, 0.f);
                for (int j = 0; j < m.y; ++j) {
                    const ulonglong2 q = mg[j];
                    const float2 ww = w[j];
                    acc = a2m_fft::fma2(q.x, a2m_fft::bcast(ww.x), acc);
                    acc = a2m_fft::fma2(q.y, a2m_fft::bcast(ww.y), acc);
                }
                const float ea = a2m_fft::lo(acc), eb = a2m_fft::hi(acc);
                if (has_a) row_a[c] = __logf(g.log_mode ? (ea == 0.f ? g.log_offset : ea) : ea + g.log_offset);
                if (has_b) row_a[g.n_mel + c] = __logf(g.log_mode ? (eb == 0.f ? g.log_offset : eb) : eb + g.log_offset);
            }
        }
        __syncwarp();                                                        // the magnitudes are exchange rows again
    }
}

// ---------------------------------------------------------------------------------------------------------------
// generic kernel
// ---------------------------------------------------------------------------------------------------------------
struct GenTables {
    const float* window;          // [nfft] zero padded
    const float2* tw;             // [n/2]  exp(-2 pi i t / n), n = nfft / 2
    const float2* untangle;       // [n+1]  (-sin, -cos)(2 pi k / nfft)
    const int4* col_meta;         // [n_mel] {first bin, bins, offset into weights, 0}
    const float* weights;         // [nnz]
};
struct GenGeom {
    int window, hop, nfft, n_mel;
    float log_offset;
    int log_mode;
};
constexpr int kGenThreads = 128;

template <typename InT, bool kMagOnly>
__global__ void __launch_bounds__(kGenThreads)
logmel_generic_kernel(const InT* __restrict__ wav, long long wav_stride, long long n_clips, long long frames_per_clip,
                      GenTables tab, GenGeom g, float* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int n = g.nfft / 2;
    float2* buf0 = reinterpret_cast<float2*>(smem);
    float2* buf1 = buf0 + n;
    float* mag = reinterpret_cast<float*>(buf1 + n);                         // [n + 1]
    const int tid = threadIdx.x;
    const long long total = n_clips * frames_per_clip;
    for (long long fr = blockIdx.x; fr < total; fr += gridDim.x) {
        const long long clip = fr / frames_per_clip;
        const InT* src = wav + clip * wav_stride + (fr - clip * frames_per_clip) * g.hop;
        for (int m = tid; m < n; m += kGenThreads) {                         // even/odd packing, window, zero padding
            const int i = 2 * m;
            const float a = i < g.window ? to_f32<InT>(src[i]) * __ldg(tab.window + i) : 0.f;
            const float b = i + 1 < g.window ? to_f32<InT>(src[i + 1]) * __ldg(tab.window + i + 1) : 0.f;
            buf0[m] = make_float2(a, b);
        }
        __syncthreads();
        float2* x = buf0;
        float2* y = buf1;
        for (int len = 1; len < n; len <<= 1) {                              // radix-2 Stockham, autosort
            const int tw_step = n / (2 * len);
            for (int j = tid; j < n / 2; j += kGenThreads) {
                const int k = j & (len - 1);
                const float2 w = __ldg(tab.tw + k * tw_step);
                const float2 a = x[j], b0 = x[j + n / 2];
                const float2 b = make_float2(b0.x * w.x - b0.y * w.y, b0.x * w.y + b0.y * w.x);
                const int idx = ((j - k) << 1) + k;
                y[idx] = make_float2(a.x + b.x, a.y + b.y);
                y[idx + len] = make_float2(a.x - b.x, a.y - b.y);
            }
            __syncthreads();
            float2* t = x; x = y; y = t;
        }
        for (int k = tid; k <= n; k += kGenThreads) {                        // real-input untangle, bins 0..n
            const float2 zk = x[k & (n - 1)], zq = x[(n - k) & (n - 1)];
            const float2 t = __ldg(tab.untangle + k);
            const float ar = zk.x + zq.x, ai = zk.y - zq.y, br = zk.x - zq.x, bi = zk.y + zq.y;
            const float xr = ar + (t.x * br - t.y * bi), xi = ai + (t.x * bi + t.y * br);
            mag[k] = 0.5f * sqrtf(xr * xr + xi * xi);
        }
        __syncthreads();
        if (kMagOnly) {
            float* o = out + fr * (n + 1);
            for (int k = tid; k <= n; k += kGenThreads) o[k] = mag[k];
        } else {
            float* o = out + fr * g.n_mel;
            for (int c = tid; c < g.n_mel; c += kGenThreads) {
                const int4 m = __ldg(tab.col_meta + c);
                float acc = 0.f;
                for (int j = 0; j < m.y; ++j) acc = fmaf(mag[m.x + j], __ldg(tab.weights + m.z + j), acc);
                o[c] = logf(g.log_mode ? (acc == 0.f ? g.log_offset : acc) : acc + g.log_offset);
            }
        }
        __syncthreads();                                                     // mag / buffers are rewritten by the next frame
    }
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// plan
// ---------------------------------------------------------------------------------------------------------------
struct a2m_mel_plan {
    int device;
    int window, hop, nfft, n_mel;
    float log_offset;
    int log_mode;
    void* blob;           // one device allocation holding all tables
    bool fast;            // the 512-point kernel's tables are present
    int fast_nnz;
    FastTables ftab;
    GenTables gtab;
};

extern "C" int a2m_mel_plan_create(int window, int hop, int nfft, int n_mel, const double* hann_host,
                                   const double* mel_weights_host, double log_offset, int device,
                                   a2m_mel_plan** out) {
    return a2m_mel_plan_create_ex(window, hop, nfft, n_mel, hann_host, mel_weights_host, log_offset, A2M_LOG_ADD_OFFSET,
                                  device, out);
}

extern "C" int a2m_mel_plan_create_ex(int window, int hop, int nfft, int n_mel, const double* hann_host,
                                      const double* mel_weights_host, double log_offset, int log_mode, int device,
                                      a2m_mel_plan** out) {
    A2M_ARG_CHECK(out != nullptr, "a2m_mel_plan_create: out is NULL");
    *out = nullptr;
    A2M_ARG_CHECK(log_mode == A2M_LOG_ADD_OFFSET || log_mode == A2M_LOG_FLOOR_ZEROS, "a2m_mel_plan_create: log_mode %d", log_mode);
    if (nfft < 64 || nfft > 4096 || (nfft & (nfft - 1)) != 0) {
        a2m_set_error("a2m_mel_plan_create: fft length %d not supported (powers of two from 64 to 4096)", nfft);
        return A2M_ERR_UNSUPPORTED;
    }
    A2M_ARG_CHECK(window >= 1 && window <= nfft, "a2m_mel_plan_create: window %d must be in [1, %d]", window, nfft);
    A2M_ARG_CHECK(hop >= 1, "a2m_mel_plan_create: hop %d must be >= 1", hop);
    A2M_ARG_CHECK(n_mel >= 1 && n_mel <= kMaxMel, "a2m_mel_plan_create: n_mel %d must be in [1, %d]", n_mel, kMaxMel);
    A2M_ARG_CHECK(hann_host && mel_weights_host, "a2m_mel_plan_create: NULL table");
    const int bins = nfft / 2 + 1, n = nfft / 2;

    // column-compressed mel matrix: per band the run of bins from its first to its last non-zero weight
    std::vector<int> first(n_mel, -1), last(n_mel, -1);
    for (int c = 0; c < n_mel; ++c) {
        for (int k = 0; k < bins; ++k)
            if (mel_weights_host[static_cast<size_t>(k) * n_mel + c] != 0.0) { if (first[c] < 0) first[c] = k; last[c] = k; }
    }
    auto weight = [&](int k, int c) { return k < bins ? static_cast<float>(mel_weights_host[static_cast<size_t>(k) * n_mel + c]) : 0.f; };

    // generic kernel: exact runs
    std::vector<int4> gmeta(kMaxMel, make_int4(0, 0, 0, 0));
    std::vector<float> gweights;
    for (int c = 0; c < n_mel; ++c) {
        if (first[c] < 0) continue;                         // empty band: log(offset)
        gmeta[c] = make_int4(first[c], last[c] - first[c] + 1, static_cast<int>(gweights.size()), 0);
        for (int k = first[c]; k <= last[c]; ++k) gweights.push_back(weight(k, c));
    }
    // 512-point kernel: runs padded (zero weights) to whole bin PAIRS starting on an even bin; 0.5 folded in (exact)
    const bool fast = nfft == kFastNfft;
    std::vector<int4> fmeta(kMaxMel, make_int4(0, 0, 0, 0));
    std::vector<float> fweights;
    if (fast) {
        for (int c = 0; c < n_mel; ++c) {
            if (first[c] < 0) continue;
            const int lo = first[c] & ~1, hi = last[c] | 1;                 // hi <= 257
            fmeta[c] = make_int4(lo, (hi - lo + 1) / 2, static_cast<int>(fweights.size()), 0);
            for (int k = lo; k <= hi; ++k) fweights.push_back(0.5f * weight(k, c));
        }
    }

    const double two_pi = 6.283185307179586476925286766559;
    std::vector<float> win(nfft, 0.f);                      // zero padded: samples past the window contribute nothing
    for (int i = 0; i < window; ++i) win[i] = static_cast<float>(hann_host[i]);
    std::vector<float2> gtw(n / 2), gunt(n + 1);
    for (int t = 0; t < n / 2; ++t)
        gtw[t] = make_float2(static_cast<float>(std::cos(two_pi * t / n)), static_cast<float>(-std::sin(two_pi * t / n)));
    for (int k = 0; k <= n; ++k)
        gunt[k] = make_float2(static_cast<float>(-std::sin(two_pi * k / nfft)), static_cast<float>(-std::cos(two_pi * k / nfft)));
    std::vector<float2> ftw(256), funt(256);
    for (int k1 = 0; k1 < 16; ++k1)
        for (int m2 = 0; m2 < 16; ++m2) {                   // [k1][m2]: a 16-lane group reads one contiguous row
            const int e = (k1 * m2) & 255;
            ftw[k1 * 16 + m2] = make_float2(static_cast<float>(std::cos(two_pi * e / 256.0)), static_cast<float>(-std::sin(two_pi * e / 256.0)));
        }
    for (int e = 0; e < 256; ++e)
        funt[e] = make_float2(static_cast<float>(-std::sin(two_pi * e / 512.0)), static_cast<float>(-std::cos(two_pi * e / 512.0)));

    A2M_CUDA_CHECK(cudaSetDevice(device));
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~size_t(255); return o; };
    const size_t o_win = carve(nfft * 4), o_gtw = carve((n / 2) * 8), o_gunt = carve((n + 1) * 8), o_gmeta = carve(kMaxMel * 16),
                 o_gw = carve((gweights.size() + 4) * 4), o_ftw = carve(256 * 8), o_funt = carve(256 * 8),
                 o_fmeta = carve(kMaxMel * 16), o_fw = carve((fweights.size() + 4) * 4);
    unsigned char* blob = nullptr;
    A2M_CUDA_CHECK(cudaMalloc(&blob, off));
    std::vector<unsigned char> host(off, 0);
    memcpy(host.data() + o_win, win.data(), nfft * 4);
    memcpy(host.data() + o_gtw, gtw.data(), gtw.size() * 8);
    memcpy(host.data() + o_gunt, gunt.data(), gunt.size() * 8);
    memcpy(host.data() + o_gmeta, gmeta.data(), kMaxMel * 16);
    if (!gweights.empty()) memcpy(host.data() + o_gw, gweights.data(), gweights.size() * 4);
    memcpy(host.data() + o_ftw, ftw.data(), 256 * 8);
    memcpy(host.data() + o_funt, funt.data(), 256 * 8);
    memcpy(host.data() + o_fmeta, fmeta.data(), kMaxMel * 16);
    if (!fweights.empty()) memcpy(host.data() + o_fw, fweights.data(), fweights.size() * 4);
    cudaError_t e = cudaMemcpy(blob, host.data(), off, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(blob); a2m_set_error("a2m_mel_plan_create: upload failed: %s", cudaGetErrorString(e)); return (int)e; }

    a2m_mel_plan* p = new a2m_mel_plan();
    p->device = device; p->window = window; p->hop = hop; p->nfft = nfft; p->n_mel = n_mel;
    p->log_offset = static_cast<float>(log_offset); p->log_mode = log_mode; p->blob = blob;
    p->fast = fast; p->fast_nnz = static_cast<int>(fweights.size());
    p->gtab.window = reinterpret_cast<const float*>(blob + o_win);
    p->gtab.tw = reinterpret_cast<const float2*>(blob + o_gtw);
    p->gtab.untangle = reinterpret_cast<const float2*>(blob + o_gunt);
    p->gtab.col_meta = reinterpret_cast<const int4*>(blob + o_gmeta);
    p->gtab.weights = reinterpret_cast<const float*>(blob + o_gw);
    p->ftab.window = p->gtab.window;
    p->ftab.tw = reinterpret_cast<const float2*>(blob + o_ftw);
    p->ftab.untangle = reinterpret_cast<const float2*>(blob + o_funt);
    p->ftab.col_meta = reinterpret_cast<const int4*>(blob + o_fmeta);
    p->ftab.weights = reinterpret_cast<const float*>(blob + o_fw);
    *out = p;
    return A2M_OK;
}

extern "C" void a2m_mel_plan_destroy(a2m_mel_plan* plan) {
    if (!plan) return;
    cudaFree(plan->blob);
    delete plan;
}

extern "C" int64_t a2m_mel_num_frames(const a2m_mel_plan* plan, int64_t n_samples) {
    if (!plan) return -1;
    const long long d = n_samples - plan->window;
    // floor division (python semantics) for negative numerators
    long long q = d / plan->hop;
    if ((d % plan->hop != 0) && (d < 0)) --q;
    return 1 + q;
}

namespace {

constexpr int kSmemCap = 227 * 1024;

// shared memory of the 512-point kernel for this geometry; 0 if it cannot run (the generic kernel takes over)
template <typename InT>
int fast_geometry(const a2m_mel_plan* plan, FastGeom* g) {
    if (!plan->fast) return 0;
    g->window = plan->window; g->hop = plan->hop; g->n_mel = plan->n_mel; g->nnz = plan->fast_nnz;
    g->n_m1 = (plan->window + 31) / 32;
    g->log_offset = plan->log_offset; g->log_mode = plan->log_mode;
    // the furthest element a tile reads: alignment shift (< 16 bytes) + 3 hops + the padded rows of the last frame
    const long long span_elems = 16 / static_cast<int>(sizeof(InT)) + 3LL * plan->hop + 32LL * g->n_m1;
    const long long span_bytes = (span_elems * static_cast<long long>(sizeof(InT)) + 15) & ~15LL;
    g->tables_bytes = (8192 + ((plan->fast_nnz + 3) / 4) * 16 + 15) & ~15;
    const long long total = g->tables_bytes + static_cast<long long>(kWarps) * (2 * kXchgGroupBytes + span_bytes + 16);
    if (total > kSmemCap) return 0;
    g->span_bytes = static_cast<int>(span_bytes);
    return static_cast<int>(total);
}

template <typename InT, bool kMagOnly>
int launch_logmel(const a2m_mel_plan* plan, const InT* wav, int64_t n_clips, int64_t n_samples, int64_t wav_stride,
                  float* out, void* stream, const char* who) {
    A2M_ARG_CHECK(plan != nullptr, "%s: plan is NULL", who);
    A2M_ARG_CHECK(n_clips >= 0 && n_samples >= 0, "%s: negative size", who);
    A2M_ARG_CHECK(n_clips <= 1 || wav_stride >= n_samples, "%s: wav_stride %lld < n_samples %lld", who, (long long)wav_stride, (long long)n_samples);
    const int64_t frames = a2m_mel_num_frames(plan, n_samples);
    A2M_ARG_CHECK(frames >= 0, "%s: negative dimensions are not allowed (%lld samples, window %d)", who,
                  (long long)n_samples, plan->window);
    if (n_clips == 0 || frames == 0) return A2M_OK;           // empty result, nothing to launch
    A2M_ARG_CHECK(wav != nullptr && out != nullptr, "%s: NULL buffer", who);
    A2M_ARG_CHECK((reinterpret_cast<uintptr_t>(wav) % sizeof(InT)) == 0, "%s: misaligned waveform pointer", who);
    A2M_ARG_CHECK(frames <= 0x7fffffffLL, "%s: %lld frames per clip", who, (long long)frames);
    cudaStream_t st = static_cast<cudaStream_t>(stream);

    FastGeom fg;
    const int fast_smem = kMagOnly ? 0 : fast_geometry<InT>(plan, &fg);
    const int tiles_per_clip = static_cast<int>((frames + kTileFrames - 1) / kTileFrames);
    const long long n_tiles = static_cast<long long>(tiles_per_clip) * n_clips;
    if (fast_smem > 0 && n_tiles <= 0x7fffffffLL) {
        static A2mPerDeviceOnce attr_set;                     // one instance per template instantiation
        if (attr_set.first())
            A2M_CUDA_CHECK(cudaFuncSetAttribute(logmel512_kernel<InT>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemCap));
        const int ctas_per_sm = 2 * (fast_smem + 1024) <= 228 * 1024 ? 2 : 1;
        long long grid = static_cast<long long>(ctas_per_sm) * a2m_num_sms();
        const long long need = (n_tiles + kWarps - 1) / kWarps;
        if (grid > need) grid = need;
        logmel512_kernel<InT><<<static_cast<unsigned>(grid), kWarps * 32, fast_smem, st>>>(
            wav, wav_stride, n_samples, n_clips, static_cast<int>(frames), tiles_per_clip, n_tiles, plan->ftab, fg, out);
    } else {
        GenGeom g;
        g.window = plan->window; g.hop = plan->hop; g.nfft = plan->nfft; g.n_mel = plan->n_mel;
        g.log_offset = plan->log_offset; g.log_mode = plan->log_mode;
        const int smem = plan->nfft * 8 + (plan->nfft / 2 + 4) * 4;
        static A2mPerDeviceOnce attr_set;
        if (attr_set.first())
            A2M_CUDA_CHECK(cudaFuncSetAttribute(logmel_generic_kernel<InT, kMagOnly>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        long long grid = n_clips * frames;
        const long long cap = 16LL * a2m_num_sms();
        if (grid > cap) grid = cap;
        logmel_generic_kernel<InT, kMagOnly><<<static_cast<unsigned>(grid), kGenThreads, smem, st>>>(
            wav, wav_stride, n_clips, frames, plan->gtab, g, out);
    }
    a2m_count_launch();
    A2M_LAUNCH_CHECK();
    return A2M_OK;
}

}  // namespace

extern "C" int a2m_logmel_f32(const a2m_mel_plan* plan, const float* wav, int64_t n_clips, int64_t n_samples,
                              int64_t wav_stride, float* out, void* stream) {
    return launch_logmel<float, false>(plan, wav, n_clips, n_samples, wav_stride, out, stream, "a2m_logmel_f32");
}

extern "C" int a2m_logmel_i16(const a2m_mel_plan* plan, const int16_t* wav, int64_t n_clips, int64_t n_samples,
                              int64_t wav_stride, float* out, void* stream) {
    return launch_logmel<int16_t, false>(plan, wav, n_clips, n_samples, wav_stride, out, stream, "a2m_logmel_i16");
}

extern "C" int a2m_stft_magnitude_f32(const a2m_mel_plan* plan, const float* wav, int64_t n_clips,
                                      int64_t n_samples, int64_t wav_stride, float* out, void* stream) {
    return launch_logmel<float, true>(plan, wav, n_clips, n_samples, wav_stride, out, stream, "a2m_stft_magnitude_f32");
}
