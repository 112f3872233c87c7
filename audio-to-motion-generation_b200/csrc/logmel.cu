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
//    * magnitudes of the pair go to shared memory once ([bin] -> (A, B)) and are read back ONCE: a triangular filterbank
//      feeds every bin to at most two adjacent bands, so bin k belongs to "segment" g (band g - 1 with weight v_k, band g
//      with weight u_k); lane l sums segments l, l + 16, ... (R = sum u mag, F = sum v mag), band g - 1 = R[g - 1] + F[g]
//      arrives by one shuffle.  The order in which a lane visits its bins is fixed at plan time by an edge colouring of
//      the (lane, bin mod 16) graph, so the 16 lanes of a step always hit 16 different banks; log(mel + offset) goes
//      straight to global memory.
//    fp32 and int16 PCM input (the reference accepts any real dtype; int16 halves the HBM / PCIe bytes per sample).
//
//  logmel_generic_kernel  (any power-of-two fft length 64 .. 4096; also |STFT| only)
//    one CTA per frame, radix-2 Stockham FFT of the even/odd-packed frame in shared memory, in fp64 like the
//    reference's numpy path.  Correct for every geometry log_mel_spectrogram can ask for (its 8 kHz default -> 256,
//    22.05 kHz -> 1024, 44.1 / 48 kHz -> 2048); not tuned.
#include <cmath>
#include <mutex>
#include <utility>
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
constexpr int kMagZeroSlots = 272;                    // (A, B) magnitudes of bins 0..256 alias the exchange rows; entries
constexpr int kMagEntries = kMagZeroSlots + 16;       // 272..287 are zeros, one per bank: what idle schedule slots read
static_assert(kMagEntries * 8 <= kXchgGroupBytes, "magnitudes must fit the exchange region");

struct FastTables {               // device-resident constants of a plan (512-point kernel)
    const float* window;          // [512]  window, zero padded past the window length
    const float2* tw;             // [16][16] exp(-2 pi i m2 k1 / 256) at [k1][m2]
    const float2* untangle;       // [256]  (-sin, -cos)(2 pi k / 512)
    const int* sched_bin;         // [steps][16] bin a lane reads at a step (kMagZeroSlots + bank = none: zero magnitude, zero weights)
    const float2* sched_uv;       // [steps][16] (u, v) = 0.5 * (weight into band g, weight into band g - 1); the kernel keeps |2 X|
    const int* round_steps;       // [n_rounds] schedule rows of each round
};
constexpr int kMaxRounds = 9;                         // ceil((128 + 1) / 16)

struct FastGeom {
    int window, hop, n_mel;
    int n_steps;                  // rows of the mel schedule
    int n_rounds;                 // segment rounds: round r, lane l -> segment 16 r + l
    int dist_last;                // the last round holds one segment only: its bins are spread over the lanes instead
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

// natural log of a positive normal number (every caller adds an offset or floors zeros first): one MUFU + one multiply
__device__ __forceinline__ float fast_log(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y * 0.69314718055994530942f;
}

__device__ __forceinline__ void wait_or_trap(uint64_t* bar, uint32_t parity) {
    const long long t0 = clock64();
    while (!a2m::mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) asm volatile("trap;");     // a pipeline bug must never hang the GPU box
    }
}

// samples (x[i], x[i + 1]) of a staged span; `aligned`: i is even (fp32: 8-byte aligned, int16: 4-byte aligned).
// Unaligned fp32 pairs are two 4-byte loads; the two 16-lane groups of the warp (grp) take them in opposite order, so
// that one instruction covers the even banks from one group and the odd banks from the other instead of colliding.
template <typename InT>
__device__ __forceinline__ float2 load_pair(const InT* __restrict__ s, int i, bool aligned, int grp);
template <>
__device__ __forceinline__ float2 load_pair<float>(const float* __restrict__ s, int i, bool aligned, int grp) {
    if (aligned) return *reinterpret_cast<const float2*>(s + i);
    const float first = s[i + grp], second = s[i + 1 - grp];
    return grp ? make_float2(second, first) : make_float2(first, second);
}
template <>
__device__ __forceinline__ float2 load_pair<int16_t>(const int16_t* __restrict__ s, int i, bool aligned, int grp) {
    (void)grp;
    if (aligned) {
        const uint32_t u = *reinterpret_cast<const uint32_t*>(s + i);
        return make_float2(static_cast<float>(static_cast<int16_t>(u & 0xffffu)), static_cast<float>(static_cast<int16_t>(u >> 16)));
    }
    return make_float2(static_cast<float>(s[i]), static_cast<float>(s[i + 1]));
}

// samples * window of rows m1 < kRows (kRows == 0: g.n_m1 rows, decided at run time) -> z[16 m1 + l] of frames A and B
template <typename InT, int kRows, bool kAligned>
__device__ __forceinline__ void load_windowed(cpx (&v)[16], const InT* __restrict__ s, const float* __restrict__ s_window,
                                              int rel_a, int rel_b, int l, int grp, int n_m1) {
    const bool al_a = kAligned || (rel_a & 1) == 0, al_b = kAligned || (rel_b & 1) == 0;       // warp-uniform
#pragma unroll
    for (int m1 = 0; m1 < 16; ++m1) {
        if (kRows ? m1 < kRows : m1 < n_m1) {                                // uniform
            const int n = 32 * m1 + 2 * l;
            const float2 w = *reinterpret_cast<const float2*>(s_window + n);
            // past the window (last row) the buffer holds samples of later frames or of an earlier tile: finite values
            // (the buffer is zeroed once, then only ever holds copied input), times the window's zero padding
            const float2 xa = load_pair<InT>(s, rel_a + n, al_a, grp);
            const float2 xb = load_pair<InT>(s, rel_b + n, al_b, grp);
            v[m1] = a2m_fft::make(a2m_fft::pack(xa.x * w.x, xb.x * w.x), a2m_fft::pack(xa.y * w.y, xb.y * w.y));
        } else {
            v[m1] = a2m_fft::make(a2m_fft::pack(0.f, 0.f), a2m_fft::pack(0.f, 0.f));
        }
    }
}

template <typename InT, int kRows>
__global__ void __launch_bounds__(kWarps * 32, 2)
logmel512_kernel(const InT* __restrict__ wav, long long wav_stride, long long n_samples, long long n_clips,
                 int frames_per_clip, int tiles_per_clip, long long n_tiles, FastTables tab, FastGeom g,
                 float* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem[];
    float* s_window = reinterpret_cast<float*>(smem);
    float2* s_tw = reinterpret_cast<float2*>(smem + 2048);
    float2* s_unt = reinterpret_cast<float2*>(smem + 4096);
    float2* s_uv = reinterpret_cast<float2*>(smem + 6144);                   // [n_steps][16]
    int* s_off = reinterpret_cast<int*>(smem + 6144 + g.n_steps * 128);      // [n_steps][16] byte offset of the bin's magnitudes
    int* s_round = s_off + g.n_steps * 16;                                   // [n_rounds]

    const int tid = threadIdx.x;
    for (int i = tid; i < kFastNfft; i += kWarps * 32) s_window[i] = tab.window[i];
    for (int i = tid; i < 256; i += kWarps * 32) { s_tw[i] = tab.tw[i]; s_unt[i] = tab.untangle[i]; }
    for (int i = tid; i < g.n_steps * 16; i += kWarps * 32) { s_uv[i] = tab.sched_uv[i]; s_off[i] = 8 * tab.sched_bin[i]; }
    if (tid < g.n_rounds) s_round[tid] = tab.round_steps[tid];

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

    const uintptr_t wav_addr = reinterpret_cast<uintptr_t>(wav);
    const uintptr_t buf_lo = (wav_addr + 15) & ~uintptr_t(15);               // 16-byte blocks fully inside the caller's buffer
    const uintptr_t buf_hi = reinterpret_cast<uintptr_t>(wav + (n_clips - 1) * wav_stride + n_samples) & ~uintptr_t(15);

    // tiles of this warp: tile = clip * tiles_per_clip + tq, advanced by n_warps without a division
    const unsigned n_warps = gridDim.x * kWarps;
    const unsigned step_clip = n_warps / static_cast<unsigned>(tiles_per_clip), step_tq = n_warps % static_cast<unsigned>(tiles_per_clip);
    const unsigned first = blockIdx.x * kWarps + warp;
    unsigned clip = first / static_cast<unsigned>(tiles_per_clip), tq = first % static_cast<unsigned>(tiles_per_clip);
    auto advance = [&](unsigned& c, unsigned& q) {
        c += step_clip;
        q += step_tq;
        if (q >= static_cast<unsigned>(tiles_per_clip)) { q -= tiles_per_clip; ++c; }
    };
    auto tile_addr = [&](unsigned c, unsigned q) {
        return wav_addr + (static_cast<long long>(c) * wav_stride + static_cast<long long>(q) * (kTileFrames * g.hop)) * static_cast<long long>(sizeof(InT));
    };

    // Stage the samples of one tile: element i of the tile's span lands at byte (src & 15) + i * sizeof(InT) of the
    // warp's buffer.  Lane 0 issues one bulk copy; the 16-byte blocks at the two ends of the CALLER's buffer that the
    // copy must not touch (first tile of a misaligned buffer, last tile) are copied by hand by the whole warp.
    auto stage = [&](unsigned c, unsigned q) {
        const uintptr_t a = tile_addr(c, q);
        const int nf = min(kTileFrames, frames_per_clip - static_cast<int>(q) * kTileFrames);
        const uintptr_t a_end = a + static_cast<uintptr_t>((nf - 1) * g.hop + g.window) * sizeof(InT);
        const uintptr_t a0 = a & ~uintptr_t(15);
        uintptr_t lo = a0, hi = (a_end + 15) & ~uintptr_t(15);
        const bool edge = lo < buf_lo || hi > buf_hi;
        if (edge) {
            if (lo < buf_lo) lo = buf_lo;
            if (hi > buf_hi) hi = buf_hi;
        }
        const bool bulk = hi > lo;
        if (lane == 0) {
            a2m::fence_proxy_async_smem();                  // the warp's reads of the previous tile precede the async writes
            a2m::mbar_expect_tx(bar, bulk ? static_cast<uint32_t>(hi - lo) : 0u);
            if (bulk) a2m::bulk_load_1d(s_samples + (lo - a0), reinterpret_cast<const void*>(lo), static_cast<uint32_t>(hi - lo), bar);
        }
        if (edge) {
            if (!bulk) { lo = a_end; hi = a_end; }         // everything by hand
            for (uintptr_t p = a + lane * sizeof(InT); p < lo && p < a_end; p += 32 * sizeof(InT))          // head
                *reinterpret_cast<InT*>(s_samples + (p - a0)) = *reinterpret_cast<const InT*>(p);
            for (uintptr_t p = (hi > a ? hi : a) + lane * sizeof(InT); p < a_end; p += 32 * sizeof(InT))     // tail
                *reinterpret_cast<InT*>(s_samples + (p - a0)) = *reinterpret_cast<const InT*>(p);
        }
    };

    if (clip < n_clips) stage(clip, tq);
    uint32_t phase = 0;
    const ulonglong2* xq = reinterpret_cast<const ulonglong2*>(xchg);
    ulonglong2* xw = reinterpret_cast<ulonglong2*>(xchg);
    float2* mag = reinterpret_cast<float2*>(xchg);
    const int pl = (16 - l) & 15;                                            // lane holding Z[256 - k] of this lane's k
    const bool lane0 = l == 0;
    const int n_mel = g.n_mel, n_rounds = g.n_rounds;
    const float log_offset = g.log_offset;
    const bool floor_zeros = g.log_mode != 0, dist_last = g.dist_last != 0;

    for (; clip < n_clips; advance(clip, tq)) {
        const int f0 = static_cast<int>(tq) * kTileFrames;
        const int shift = static_cast<int>((tile_addr(clip, tq) & 15) / sizeof(InT));
        wait_or_trap(bar, phase);
        phase ^= 1;
        __syncwarp();                                                        // hand-copied head / tail of the other lanes

        // ---- samples * window -> z[16 m1 + l] of frames A = f0 + 2 grp and B = A + 1 ---------------------------
        cpx v[16];
        {
            const InT* s = reinterpret_cast<const InT*>(s_samples);
            const int rel_a = shift + 2 * grp * g.hop, rel_b = rel_a + g.hop;
            if (((shift | g.hop) & 1) == 0) load_windowed<InT, kRows, true>(v, s, s_window, rel_a, rel_b, l, grp, g.n_m1);
            else load_windowed<InT, kRows, false>(v, s, s_window, rel_a, rel_b, l, grp, g.n_m1);
        }
        __syncwarp();                                                        // every lane has its samples in registers
        {
            unsigned c2 = clip, q2 = tq;
            advance(c2, q2);
            if (c2 < n_clips) stage(c2, q2);                                 // overlaps everything below
        }

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
            if (lane0) mag[128] = make_float2(2.f * fast_sqrt(a2m_fft::lo(sq)), 2.f * fast_sqrt(a2m_fft::hi(sq)));
            mag[kMagZeroSlots + l] = make_float2(0.f, 0.f);                  // what the idle slots of the mel schedule read
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

        // ---- mel: segment g = 16 r + l per round; band g - 1 = R[g - 1] + F[g] ---------------------------------------
        // The schedule rows of a round are padded to PAIRS (idle slots read zeros); the (offset, weights) of the next pair
        // are fetched while the current pair's magnitudes are in flight.
        {
            const int fa = f0 + 2 * grp;
            float* out_p = out + (static_cast<long long>(clip) * frames_per_clip + fa) * n_mel + (l - 1);   // band 16 r + l - 1
            const bool has_a = fa < frames_per_clip, has_b = fa + 1 < frames_per_clip;
            const pair_t zero = a2m_fft::pack(0.f, 0.f);
            pair_t carry = zero;                                             // R of lane 15 of the previous round
            const unsigned char* magb = reinterpret_cast<const unsigned char*>(mag);
            const int* off_p = s_off + l;
            const float2* uv_p = s_uv + l;
            int off0 = off_p[0], off1 = off_p[16];
            float2 uv0 = uv_p[0], uv1 = uv_p[16];
            int band = l - 1;
            for (int r = 0; r < n_rounds; ++r) {
                const int n_pairs = s_round[r];
                pair_t acc_r = zero, acc_f = zero;
                for (int t = 0; t < n_pairs; ++t) {
                    const pair_t m0 = *reinterpret_cast<const pair_t*>(magb + off0);
                    const pair_t m1 = *reinterpret_cast<const pair_t*>(magb + off1);
                    const float2 w0 = uv0, w1 = uv1;
                    off_p += 32;
                    uv_p += 32;
                    off0 = off_p[0]; off1 = off_p[16];                       // one pair of rows past the end exists (zeros)
                    uv0 = uv_p[0]; uv1 = uv_p[16];
                    acc_r = a2m_fft::fma2(m0, a2m_fft::bcast(w0.x), acc_r);
                    acc_f = a2m_fft::fma2(m0, a2m_fft::bcast(w0.y), acc_f);
                    acc_r = a2m_fft::fma2(m1, a2m_fft::bcast(w1.x), acc_r);
                    acc_f = a2m_fft::fma2(m1, a2m_fft::bcast(w1.y), acc_f);
                }
                pair_t e;
                bool write = band >= 0 && band < n_mel;
                float* dst = out_p;
                if (dist_last && r == n_rounds - 1) {                        // one segment spread over the lanes: band n_mel - 1
#pragma unroll
                    for (int o = 8; o > 0; o >>= 1) acc_f = a2m_fft::add2(acc_f, __shfl_xor_sync(0xffffffffu, acc_f, o, 16));
                    e = a2m_fft::add2(carry, acc_f);
                    write = lane0;                                           // band 16 r - 1 = n_mel - 1: lane 0's slot
                } else {
                    pair_t r_prev = __shfl_sync(0xffffffffu, acc_r, (l + 15) & 15, 16);
                    if (lane0) r_prev = carry;
                    carry = __shfl_sync(0xffffffffu, acc_r, 15, 16);
                    e = a2m_fft::add2(r_prev, acc_f);
                }
                if (write) {
                    const float ea = a2m_fft::lo(e), eb = a2m_fft::hi(e);
                    if (has_a) dst[0] = fast_log(floor_zeros ? (ea == 0.f ? log_offset : ea) : ea + log_offset);
                    if (has_b) dst[n_mel] = fast_log(floor_zeros ? (eb == 0.f ? log_offset : eb) : eb + log_offset);
                }
                band += 16;
                out_p += 16;
            }
        }
        __syncwarp();                                                        // the magnitudes are exchange rows again
    }
}

// ---------------------------------------------------------------------------------------------------------------
// generic kernel: fp64 arithmetic like the reference's numpy path (window, FFT, magnitude, mel sum), fp32 in / out
// ---------------------------------------------------------------------------------------------------------------
struct GenTables {
    const double* window;         // [nfft] zero padded
    const double2* tw;            // [n/2]  exp(-2 pi i t / n), n = nfft / 2
    const double2* untangle;      // [n+1]  (-sin, -cos)(2 pi k / nfft)
    const int4* col_meta;         // [n_mel] {first bin, bins, offset into weights, 0}
    const double* weights;        // [nnz]
};
struct GenGeom {
    int window, hop, nfft, n_mel;
    double log_offset;
    int log_mode;
};
constexpr int kGenThreads = 128;

template <typename InT, bool kMagOnly>
__global__ void __launch_bounds__(kGenThreads)
logmel_generic_kernel(const InT* __restrict__ wav, long long wav_stride, long long n_clips, long long frames_per_clip,
                      GenTables tab, GenGeom g, float* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int n = g.nfft / 2;
    double2* buf0 = reinterpret_cast<double2*>(smem);
    double2* buf1 = buf0 + n;
    double* mag = reinterpret_cast<double*>(buf1 + n);                       // [n + 1]
    const int tid = threadIdx.x;
    const long long total = n_clips * frames_per_clip;
    for (long long fr = blockIdx.x; fr < total; fr += gridDim.x) {
        const long long clip = fr / frames_per_clip;
        const InT* src = wav + clip * wav_stride + (fr - clip * frames_per_clip) * g.hop;
        for (int m = tid; m < n; m += kGenThreads) {                         // even/odd packing, window, zero padding
            const int i = 2 * m;
            const double a = i < g.window ? static_cast<double>(src[i]) * __ldg(tab.window + i) : 0.0;
            const double b = i + 1 < g.window ? static_cast<double>(src[i + 1]) * __ldg(tab.window + i + 1) : 0.0;
            buf0[m] = make_double2(a, b);
        }
        __syncthreads();
        double2* x = buf0;
        double2* y = buf1;
        for (int len = 1; len < n; len <<= 1) {                              // radix-2 Stockham, autosort
            const int tw_step = n / (2 * len);
            for (int j = tid; j < n / 2; j += kGenThreads) {
                const int k = j & (len - 1);
                const double2 w = __ldg(tab.tw + k * tw_step);
                const double2 a = x[j], b0 = x[j + n / 2];
                const double2 b = make_double2(b0.x * w.x - b0.y * w.y, b0.x * w.y + b0.y * w.x);
                const int idx = ((j - k) << 1) + k;
                y[idx] = make_double2(a.x + b.x, a.y + b.y);
                y[idx + len] = make_double2(a.x - b.x, a.y - b.y);
            }
            __syncthreads();
            double2* t = x; x = y; y = t;
        }
        for (int k = tid; k <= n; k += kGenThreads) {                        // real-input untangle, bins 0..n
            const double2 zk = x[k & (n - 1)], zq = x[(n - k) & (n - 1)];
            const double2 t = __ldg(tab.untangle + k);
            const double ar = zk.x + zq.x, ai = zk.y - zq.y, br = zk.x - zq.x, bi = zk.y + zq.y;
            const double xr = ar + (t.x * br - t.y * bi), xi = ai + (t.x * bi + t.y * br);
            mag[k] = 0.5 * sqrt(xr * xr + xi * xi);
        }
        __syncthreads();
        if (kMagOnly) {
            float* o = out + fr * (n + 1);
            for (int k = tid; k <= n; k += kGenThreads) o[k] = static_cast<float>(mag[k]);
        } else {
            float* o = out + fr * g.n_mel;
            for (int c = tid; c < g.n_mel; c += kGenThreads) {
                const int4 m = __ldg(tab.col_meta + c);
                double acc = 0.0;
                for (int j = 0; j < m.y; ++j) acc = fma(mag[m.x + j], __ldg(tab.weights + m.z + j), acc);
                o[c] = static_cast<float>(log(g.log_mode ? (acc == 0.0 ? g.log_offset : acc) : acc + g.log_offset));
            }
        }
        __syncthreads();                                                     // mag / buffers are rewritten by the next frame
    }
}

}  // namespace


// ---------------------------------------------------------------------------------------------------------------
// mel schedule of the 512-point kernel (host, plan time)
// ---------------------------------------------------------------------------------------------------------------
namespace {

struct MelSchedule {
    bool ok = false;
    std::vector<int> bin;          // [steps][16]
    std::vector<float2> uv;        // [steps][16]
    int n_rounds = 0, dist_last = 0;
    int round_steps[kMaxRounds + 3] = {};
};

// Proper edge colouring of a bipartite multigraph (16 lanes x 16 residues, one edge per bin) with max-degree many
// colours (Koenig): colour = the step at which the lane reads the bin.  Returns the number of colours.
int colour_edges(const std::vector<int>& eu, const std::vector<int>& ev, std::vector<int>& colour) {
    const int n_e = static_cast<int>(eu.size());
    int deg_u[16] = {}, deg_v[16] = {}, delta = 0;
    for (int e = 0; e < n_e; ++e) { ++deg_u[eu[e]]; ++deg_v[ev[e]]; }
    for (int i = 0; i < 16; ++i) { if (deg_u[i] > delta) delta = deg_u[i]; if (deg_v[i] > delta) delta = deg_v[i]; }
    std::vector<int> at_u(16 * (delta + 1), -1), at_v(16 * (delta + 1), -1);   // [vertex][colour] -> edge
    colour.assign(n_e, -1);
    for (int e = 0; e < n_e; ++e) {
        const int u = eu[e], v = ev[e];
        int a = 0, b = 0;
        while (at_u[u * delta + a] >= 0) ++a;               // free at u
        while (at_v[v * delta + b] >= 0) ++b;               // free at v
        if (a != b) {
            // flip the a/b alternating path that starts at v with colour a (it cannot reach u: the graph is bipartite)
            std::vector<int> path;
            int x = v, c = a;
            bool on_v = true;
            for (;;) {
                const int f = on_v ? at_v[x * delta + c] : at_u[x * delta + c];
                if (f < 0) break;
                path.push_back(f);
                x = on_v ? eu[f] : ev[f];
                on_v = !on_v;
                c = (c == a) ? b : a;
            }
            for (int f : path) { at_u[eu[f] * delta + colour[f]] = -1; at_v[ev[f] * delta + colour[f]] = -1; }
            for (int f : path) {
                colour[f] = (colour[f] == a) ? b : a;
                at_u[eu[f] * delta + colour[f]] = f;
                at_v[ev[f] * delta + colour[f]] = f;
            }
        }
        colour[e] = a;
        at_u[u * delta + a] = e;
        at_v[v * delta + a] = e;
    }
    return delta;
}

// weights: fp64 [257, n_mel] row-major.  Fails (ok = false) when a bin feeds more than two bands or two bands that are
// not adjacent -- not a triangular filterbank; the generic kernel handles such a matrix.
MelSchedule build_mel_schedule(const double* w, int n_mel) {
    MelSchedule s;
    const int bins = 257, n_seg = n_mel + 1;
    std::vector<int> seg(bins, -1);
    std::vector<float> u(bins, 0.f), v(bins, 0.f);
    std::vector<int> peak(n_mel, 0);                        // bin of each band's largest weight
    for (int c = 0; c < n_mel; ++c)
        for (int k = 0; k < bins; ++k)
            if (w[static_cast<size_t>(k) * n_mel + c] > w[static_cast<size_t>(peak[c]) * n_mel + c]) peak[c] = k;
    for (int k = 0; k < bins; ++k) {
        int c0 = -1, c1 = -1, cnt = 0;
        for (int c = 0; c < n_mel; ++c)
            if (w[static_cast<size_t>(k) * n_mel + c] != 0.0) { if (cnt == 0) c0 = c; c1 = c; ++cnt; }
        if (cnt == 0) continue;
        if (cnt > 2 || c1 - c0 > 1) return s;
        // segment g feeds band g - 1 with v and band g with u: a bin of two bands (c, c + 1) is in segment c + 1; a bin
        // of one band c is in segment c (rising edge, u) or c + 1 (falling edge past the peak, v)
        if (cnt == 2) {
            seg[k] = c1;
            v[k] = 0.5f * static_cast<float>(w[static_cast<size_t>(k) * n_mel + c0]);      // 0.5: the kernel keeps |2 X|
            u[k] = 0.5f * static_cast<float>(w[static_cast<size_t>(k) * n_mel + c1]);
        } else if (k > peak[c0]) {
            seg[k] = c0 + 1;
            v[k] = 0.5f * static_cast<float>(w[static_cast<size_t>(k) * n_mel + c0]);
        } else {
            seg[k] = c0;
            u[k] = 0.5f * static_cast<float>(w[static_cast<size_t>(k) * n_mel + c0]);
        }
    }
    s.n_rounds = (n_seg + 15) / 16;
    // a last round that holds segment n_mel alone (n_mel a multiple of 16: the falling edge of the last band) would
    // keep one lane busy for all its bins: spread them over the 16 lanes, the kernel adds the lanes up
    int n_last = 0;
    for (int k = 0; k < bins; ++k) n_last += seg[k] == n_mel;
    s.dist_last = (n_seg % 16 == 1 && n_last > 1) ? 1 : 0;
    for (int r = 0; r < s.n_rounds; ++r) {
        std::vector<int> eu, ev, ek, colour;
        int spread = 0;
        for (int k = 0; k < bins; ++k)
            if (seg[k] >= 0 && seg[k] / 16 == r) {
                eu.push_back(s.dist_last && r == s.n_rounds - 1 ? (spread++) % 16 : seg[k] % 16);
                ev.push_back(k % 16);
                ek.push_back(k);
            }
        const int used_steps = eu.empty() ? 0 : colour_edges(eu, ev, colour);
        const int steps = (used_steps + 1) & ~1;            // the kernel walks the rows in pairs
        if (steps > 254) return s;
        s.round_steps[r] = steps / 2;
        const size_t base = s.bin.size();
        s.bin.resize(base + static_cast<size_t>(steps) * 16, -1);
        s.uv.resize(base + static_cast<size_t>(steps) * 16, make_float2(0.f, 0.f));
        for (size_t e = 0; e < eu.size(); ++e) {
            const size_t slot = base + static_cast<size_t>(colour[e]) * 16 + eu[e];
            if (s.bin[slot] != -1) return s;                // colouring bug guard
            s.bin[slot] = ek[e];
            s.uv[slot] = make_float2(u[ek[e]], v[ek[e]]);
        }
        for (int t = 0; t < steps; ++t) {                   // idle slots read a zero entry in a bank no lane of the step uses
            bool used[16] = {};
            for (int ln = 0; ln < 16; ++ln) { const int k = s.bin[base + t * 16 + ln]; if (k >= 0) used[k % 16] = true; }
            int free_bank = 0;
            for (int ln = 0; ln < 16; ++ln) {
                if (s.bin[base + t * 16 + ln] >= 0) continue;
                while (used[free_bank]) ++free_bank;
                used[free_bank] = true;
                s.bin[base + t * 16 + ln] = kMagZeroSlots + free_bank;
            }
        }
    }
    // the kernel's look-ahead reads one pair of rows past the last round
    for (int t = 0; t < 2; ++t)
        for (int ln = 0; ln < 16; ++ln) { s.bin.push_back(kMagZeroSlots + ln); s.uv.push_back(make_float2(0.f, 0.f)); }
    s.ok = true;
    return s;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// plan
// ---------------------------------------------------------------------------------------------------------------
struct a2m_mel_plan {
    int device;
    int window, hop, nfft, n_mel;
    float log_offset;
    double log_offset64;
    int log_mode;
    void* blob;           // one device allocation holding all tables
    bool fast;            // the 512-point kernel's tables are present
    int n_steps, n_rounds, dist_last;
    FastTables ftab;
    GenTables gtab;
};

// Host-only diagnostic: the mel schedule the 512-point kernel would use for this filterbank (tests check on the CPU
// that it reproduces the matrix and that the 16 lanes of a step read 16 different banks).
extern "C" int a2m_mel_schedule_host(const double* mel_weights_host, int n_mel, int capacity_steps, int* bin_out,
                                     float* uv_out, int* round_steps_out, int* n_rounds_out, int* dist_last_out) {
    A2M_ARG_CHECK(mel_weights_host && bin_out && uv_out && round_steps_out && n_rounds_out && dist_last_out,
                  "a2m_mel_schedule_host: NULL argument");
    A2M_ARG_CHECK(n_mel >= 1 && n_mel <= kMaxMel, "a2m_mel_schedule_host: n_mel %d must be in [1, %d]", n_mel, kMaxMel);
    const MelSchedule s = build_mel_schedule(mel_weights_host, n_mel);
    if (!s.ok) {
        a2m_set_error("a2m_mel_schedule_host: not a triangular filterbank (a bin feeds more than two adjacent bands)");
        return A2M_ERR_UNSUPPORTED;
    }
    const int steps = static_cast<int>(s.bin.size() / 16) - 2;            // without the look-ahead rows
    A2M_ARG_CHECK(steps <= capacity_steps, "a2m_mel_schedule_host: %d steps, capacity %d", steps, capacity_steps);
    for (size_t i = 0; i < static_cast<size_t>(steps) * 16; ++i) { bin_out[i] = s.bin[i]; uv_out[2 * i] = s.uv[i].x; uv_out[2 * i + 1] = s.uv[i].y; }
    for (int r = 0; r < s.n_rounds; ++r) round_steps_out[r] = 2 * s.round_steps[r];     // rows (the kernel walks them in pairs)
    *n_rounds_out = s.n_rounds;
    *dist_last_out = s.dist_last;
    return steps;
}

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
    auto weight = [&](int k, int c) { return k < bins ? mel_weights_host[static_cast<size_t>(k) * n_mel + c] : 0.0; };

    // generic kernel: exact runs
    std::vector<int4> gmeta(kMaxMel, make_int4(0, 0, 0, 0));
    std::vector<double> gweights;
    for (int c = 0; c < n_mel; ++c) {
        if (first[c] < 0) continue;                         // empty band: log(offset)
        gmeta[c] = make_int4(first[c], last[c] - first[c] + 1, static_cast<int>(gweights.size()), 0);
        for (int k = first[c]; k <= last[c]; ++k) gweights.push_back(weight(k, c));
    }
    // 512-point kernel: the mel sum as a plan-time schedule (see build_mel_schedule); a matrix that is not a triangular
    // filterbank runs on the generic kernel
    MelSchedule sched;
    if (nfft == kFastNfft) sched = build_mel_schedule(mel_weights_host, n_mel);
    const bool fast = sched.ok;

    const double two_pi = 6.283185307179586476925286766559;
    std::vector<float> win(nfft, 0.f);                      // zero padded: samples past the window contribute nothing
    std::vector<double> gwin(nfft, 0.0);
    for (int i = 0; i < window; ++i) { win[i] = static_cast<float>(hann_host[i]); gwin[i] = hann_host[i]; }
    std::vector<double2> gtw(n / 2), gunt(n + 1);
    for (int t = 0; t < n / 2; ++t) gtw[t] = make_double2(std::cos(two_pi * t / n), -std::sin(two_pi * t / n));
    for (int k = 0; k <= n; ++k) gunt[k] = make_double2(-std::sin(two_pi * k / nfft), -std::cos(two_pi * k / nfft));
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
    const size_t o_win = carve(nfft * 4), o_gwin = carve(nfft * 8), o_gtw = carve((n / 2) * 16), o_gunt = carve((n + 1) * 16),
                 o_gmeta = carve(kMaxMel * 16), o_gw = carve((gweights.size() + 4) * 8), o_ftw = carve(256 * 8),
                 o_funt = carve(256 * 8), o_sbin = carve((sched.bin.size() + 4) * 4), o_suv = carve((sched.uv.size() + 4) * 8),
                 o_srnd = carve(sizeof(sched.round_steps));
    unsigned char* blob = nullptr;
    A2M_CUDA_CHECK(cudaMalloc(&blob, off));
    std::vector<unsigned char> host(off, 0);
    memcpy(host.data() + o_win, win.data(), nfft * 4);
    memcpy(host.data() + o_gwin, gwin.data(), nfft * 8);
    memcpy(host.data() + o_gtw, gtw.data(), gtw.size() * 16);
    memcpy(host.data() + o_gunt, gunt.data(), gunt.size() * 16);
    memcpy(host.data() + o_gmeta, gmeta.data(), kMaxMel * 16);
    if (!gweights.empty()) memcpy(host.data() + o_gw, gweights.data(), gweights.size() * 8);
    memcpy(host.data() + o_ftw, ftw.data(), 256 * 8);
    memcpy(host.data() + o_funt, funt.data(), 256 * 8);
    if (!sched.bin.empty()) {
        memcpy(host.data() + o_sbin, sched.bin.data(), sched.bin.size() * 4);
        memcpy(host.data() + o_suv, sched.uv.data(), sched.uv.size() * 8);
    }
    memcpy(host.data() + o_srnd, sched.round_steps, sizeof(sched.round_steps));
    cudaError_t e = cudaMemcpy(blob, host.data(), off, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(blob); a2m_set_error("a2m_mel_plan_create: upload failed: %s", cudaGetErrorString(e)); return (int)e; }

    a2m_mel_plan* p = new a2m_mel_plan();
    p->device = device; p->window = window; p->hop = hop; p->nfft = nfft; p->n_mel = n_mel;
    p->log_offset = static_cast<float>(log_offset); p->log_offset64 = log_offset; p->log_mode = log_mode; p->blob = blob;
    p->fast = fast; p->n_steps = static_cast<int>(sched.bin.size() / 16); p->n_rounds = sched.n_rounds;
    p->dist_last = sched.dist_last;
    p->gtab.window = reinterpret_cast<const double*>(blob + o_gwin);
    p->gtab.tw = reinterpret_cast<const double2*>(blob + o_gtw);
    p->gtab.untangle = reinterpret_cast<const double2*>(blob + o_gunt);
    p->gtab.col_meta = reinterpret_cast<const int4*>(blob + o_gmeta);
    p->gtab.weights = reinterpret_cast<const double*>(blob + o_gw);
    p->ftab.window = reinterpret_cast<const float*>(blob + o_win);
    p->ftab.tw = reinterpret_cast<const float2*>(blob + o_ftw);
    p->ftab.untangle = reinterpret_cast<const float2*>(blob + o_funt);
    p->ftab.sched_bin = reinterpret_cast<const int*>(blob + o_sbin);
    p->ftab.sched_uv = reinterpret_cast<const float2*>(blob + o_suv);
    p->ftab.round_steps = reinterpret_cast<const int*>(blob + o_srnd);
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

// cudaFuncSetAttribute is per (kernel, device): true the first time this pair is seen
bool first_use(const void* kernel) {
    static std::mutex mu;
    static std::vector<std::pair<const void*, int>> seen;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    for (const auto& e : seen)
        if (e.first == kernel && e.second == dev) return false;
    seen.emplace_back(kernel, dev);
    return true;
}

// shared memory of the 512-point kernel for this geometry; 0 if it cannot run (the generic kernel takes over)
template <typename InT>
int fast_geometry(const a2m_mel_plan* plan, FastGeom* g) {
    if (!plan->fast) return 0;
    g->window = plan->window; g->hop = plan->hop; g->n_mel = plan->n_mel;
    g->n_steps = plan->n_steps; g->n_rounds = plan->n_rounds; g->dist_last = plan->dist_last;
    g->n_m1 = (plan->window + 31) / 32;
    g->log_offset = plan->log_offset; g->log_mode = plan->log_mode;
    // the furthest element a tile reads: alignment shift (< 16 bytes) + 3 hops + the padded rows of the last frame
    const long long span_elems = 16 / static_cast<int>(sizeof(InT)) + 3LL * plan->hop + 32LL * g->n_m1;
    const long long span_bytes = (span_elems * static_cast<long long>(sizeof(InT)) + 15) & ~15LL;
    g->tables_bytes = (6144 + plan->n_steps * (128 + 64) + (kMaxRounds + 3) * 4 + 15) & ~15;
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
        const int ctas_per_sm = 2 * (fast_smem + 1024) <= 228 * 1024 ? 2 : 1;
        long long grid = static_cast<long long>(ctas_per_sm) * a2m_num_sms();
        const long long need = (n_tiles + kWarps - 1) / kWarps;
        if (grid > need) grid = need;
        auto launch = [&](auto kernel) -> int {
            if (first_use(reinterpret_cast<const void*>(kernel)))
                A2M_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemCap));
            kernel<<<static_cast<unsigned>(grid), kWarps * 32, fast_smem, st>>>(
                wav, wav_stride, n_samples, n_clips, static_cast<int>(frames), tiles_per_clip, n_tiles, plan->ftab, fg, out);
            return A2M_OK;
        };
        // the rows of the 16 x 16 input that hold window samples are a compile-time constant for the two geometries of the
        // path (25 ms at 16 kHz: 400 samples = 13 rows; log_mel_400: the 512-sample centred window = 16 rows)
        int rc;
        if (fg.n_m1 == 13) rc = launch(logmel512_kernel<InT, 13>);
        else if (fg.n_m1 == 16) rc = launch(logmel512_kernel<InT, 16>);
        else rc = launch(logmel512_kernel<InT, 0>);
        if (rc != A2M_OK) return rc;
    } else {
        GenGeom g;
        g.window = plan->window; g.hop = plan->hop; g.nfft = plan->nfft; g.n_mel = plan->n_mel;
        g.log_offset = plan->log_offset64; g.log_mode = plan->log_mode;
        const int smem = plan->nfft * 16 + (plan->nfft / 2 + 4) * 8;
        static A2mPerDeviceOnce attr_set;
        if (attr_set.first())
            A2M_CUDA_CHECK(cudaFuncSetAttribute(logmel_generic_kernel<InT, kMagOnly>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
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
