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

constexpr int kSmemFixedWide = (2 * kCPad) * 8 + kNfftW * 4 + kC * 8 + 1028 * 4;      // + 4 * nnz mel weights

}  // namespace

struct a2m_melspec_plan {
    int device, nfft, hop, n_mel, nnz, power, pad_mode, log_mode;
    float log_offset;
    void* blob;
    WideTables tab;
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
    unsigned char* blob = nullptr;
    A2M_CUDA_CHECK(cudaMalloc(&blob, off));
    std::vector<unsigned char> host(off, 0);
    memcpy(host.data() + o_win, win.data(), kNfftW * 4);
    memcpy(host.data() + o_tw, tw.data(), kC * 8);
    memcpy(host.data() + o_unt, unt.data(), kC * 8);
    memcpy(host.data() + o_cm, meta.data(), kMaxMelW * 16);
    if (nnz) memcpy(host.data() + o_wt, weights.data(), nnz * 4);
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
    const int smem = kSmemFixedWide + 4 * ((plan->nnz + 3) & ~3);
    A2M_ARG_CHECK(smem <= 100 * 1024, "a2m_melspec_f32: %d bytes of shared memory", smem);
    static A2mPerDeviceOnce attr_set;
    if (attr_set.first())                                  // the cap of the argument check above: covers every plan
        A2M_CUDA_CHECK(cudaFuncSetAttribute(melspec_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    long long grid = 4LL * a2m_num_sms();
    if (grid > g.n_items) grid = g.n_items;
    melspec_wide_kernel<<<static_cast<unsigned>(grid), kThreadsW, smem, static_cast<cudaStream_t>(stream)>>>(
        wav, plan->tab, g, out);
    a2m_count_launch();
    A2M_LAUNCH_CHECK();
    return A2M_OK;
}
