// Fused log-mel front end: framing -> periodic Hann -> 512-point real FFT ->
// |.| -> sparse mel projection -> log(. + offset), one launch for a whole batch of clips.
//
// Replaces pose_video/mel_features.py:192-223 (see include/a2m_b200.h).  HBM-bound by contract:
// every sample is read once (the 240-sample overlap between tiles is an L2 hit) and every output
// element written once; everything in between lives in registers / shared memory.
//
// Mapping: 16 threads own one frame (two register-resident radix-16 passes of a 256-point complex
// FFT of the even/odd-packed frame, one padded shared-memory exchange in between, partner bins of
// the real-input untangle fetched with warp shuffles).  A 256-thread CTA works on 16 frames at a
// time and walks over tiles of <= 32 consecutive frames of one clip; samples of a tile are staged
// once in shared memory, results are staged and written back as whole 256-byte rows.
#include <cmath>
#include <vector>
#include <cstring>
#include "a2m_common.cuh"
#include "fft_math.cuh"

using a2m_fft::cpx;

namespace {

constexpr int kThreads = 256;
constexpr int kSlots = kThreads / 16;         // frames in flight per CTA
constexpr int kMaxTileFrames = 32;
constexpr int kSpanFloats = 5632;             // staged samples per tile (22 KB)
constexpr int kXchgStride = 18;               // float2 units; 144 B rows keep LDS.128 conflict-free
constexpr int kMaxMel = 128;
constexpr int kNfft = 512;
constexpr int kBins = kNfft / 2 + 1;

struct MelTables {                // device-resident constants of a plan
    const float* window;          // [window]
    const float2* w256;           // [256]  exp(-2 pi i e / 256)
    const float2* untangle;       // [256]  (-sin, -cos)(2 pi k / 512)
    const int* col_start;         // [n_mel] first spectrogram bin with a non-zero weight
    const int* col_count;         // [n_mel]
    const int* col_ptr;           // [n_mel] offset into weights
    const float* weights;         // [nnz]
};

struct MelGeom {
    int window, hop, n_mel, nnz;
    int tile_frames;              // frames per tile
    int max_bin;                  // highest spectrogram bin with a non-zero mel weight
    float log_offset;
};

struct SmemLayout {
    int samples, xchg, out, window, w256, untangle, col_start, col_count, col_ptr, weights, total;
};

__host__ __device__ inline SmemLayout smem_layout(int n_mel, int nnz, bool mag_only) {
    SmemLayout L;
    int off = 0;
    L.samples = off;   off += kSpanFloats * 4;
    L.xchg = off;      off += kSlots * 16 * kXchgStride * 8;
    L.out = off;       off += mag_only ? 0 : kMaxTileFrames * n_mel * 4;
    L.window = off;    off += kNfft * 4;
    L.w256 = off;      off += 256 * 8;
    L.untangle = off;  off += 256 * 8;
    L.col_start = off; off += kMaxMel * 4;
    L.col_count = off; off += kMaxMel * 4;
    L.col_ptr = off;   off += kMaxMel * 4;
    L.weights = off;   off += ((nnz + 3) / 4) * 16;
    L.total = off;
    return L;
}

// One frame by 16 lanes.  `lane16` = position in the 16-lane group, `frame_smp` = first sample of
// the frame in the staged span.  On return mag[] (the group's exchange slot, reused) holds |X[k]|
// for k < 16 * n_k2 (and k = 256 at index 256 when kMagOnly).
template <bool kMagOnly>
__device__ __forceinline__ void frame_spectrum(const float* __restrict__ s_samples, int frame_smp, int window,
                                               const float* __restrict__ s_window,
                                               const float2* __restrict__ s_w256,
                                               const float2* __restrict__ s_unt, float2* __restrict__ slot,
                                               int lane16, int n_k2) {
    cpx v[16];
    // ---- pass 1: thread m2 = lane16 loads z[16 m1 + m2] = (x[32 m1 + 2 m2], x[.. + 1]) * hann
#pragma unroll
    for (int m1 = 0; m1 < 16; ++m1) {
        const int n = 32 * m1 + 2 * lane16;
        float a = 0.f, b = 0.f;
        if (n < window) a = s_samples[frame_smp + n] * s_window[n];
        if (n + 1 < window) b = s_samples[frame_smp + n + 1] * s_window[n + 1];
        v[m1] = a2m_fft::make(a, b);
    }
    a2m_fft::dft16(v);                                      // Y[m2][k1], k1 = register index
#pragma unroll
    for (int k1 = 1; k1 < 16; ++k1) {                       // * W256^(m2 k1)
        const float2 w = s_w256[lane16 * k1];
        v[k1] = a2m_fft::mul(v[k1], a2m_fft::make(w.x, w.y));
    }
    // ---- exchange: slot[k1 * stride + m2]  (writes: 16 lanes contiguous; reads: one 144 B row each)
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) slot[k1 * kXchgStride + lane16] = make_float2(v[k1].x, v[k1].y);
    __syncwarp();
    {
        const float4* row = reinterpret_cast<const float4*>(slot + lane16 * kXchgStride);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 q = row[j];
            v[2 * j] = a2m_fft::make(q.x, q.y);
            v[2 * j + 1] = a2m_fft::make(q.z, q.w);
        }
    }
    __syncwarp();                                           // slot is reused for magnitudes below
    // ---- pass 2: thread k1 = lane16, DFT over m2 -> Z[k1 + 16 k2] in v[k2]
    a2m_fft::dft16(v);
    // ---- untangle: partner bin 256-k lives in lane (16-k1)&15 at register 15-k2 (k1 != 0)
    float* mag = reinterpret_cast<float*>(slot);
    const int src = (threadIdx.x & 16) | ((16 - lane16) & 15);
    if (kMagOnly && lane16 == 0) mag[256] = fabsf(v[0].x - v[0].y);      // X[256] = Re Z0 - Im Z0
#pragma unroll
    for (int k2 = 0; k2 < 16; ++k2) {
        float px = __shfl_sync(0xffffffffu, v[15 - k2].x, src);
        float py = __shfl_sync(0xffffffffu, v[15 - k2].y, src);
        if (lane16 == 0) {                                  // k1 == 0: partner is own bin 16*((16-k2)&15)
            px = v[(16 - k2) & 15].x;
            py = v[(16 - k2) & 15].y;
        }
        if (k2 < n_k2) {
            const int k = lane16 + 16 * k2;
            const float2 t = s_unt[k];
            const cpx x2 = a2m_fft::untangle2(v[k2], a2m_fft::make(px, py), a2m_fft::make(t.x, t.y));
            mag[k] = a2m_fft::half_magnitude(x2);
        }
    }
    __syncwarp();
}

template <bool kMagOnly>
__global__ void __launch_bounds__(kThreads, 2)
logmel_kernel(const float* __restrict__ wav, long long n_clips, long long n_samples, long long wav_stride,
              long long frames_per_clip, int tiles_per_clip, long long n_tiles, MelTables tab, MelGeom g,
              float* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem[];
    const SmemLayout L = smem_layout(g.n_mel, g.nnz, kMagOnly);
    float* s_samples = reinterpret_cast<float*>(smem + L.samples);
    float2* s_xchg = reinterpret_cast<float2*>(smem + L.xchg);
    float* s_out = reinterpret_cast<float*>(smem + L.out);
    float* s_window = reinterpret_cast<float*>(smem + L.window);
    float2* s_w256 = reinterpret_cast<float2*>(smem + L.w256);
    float2* s_unt = reinterpret_cast<float2*>(smem + L.untangle);
    int* s_col_start = reinterpret_cast<int*>(smem + L.col_start);
    int* s_col_count = reinterpret_cast<int*>(smem + L.col_count);
    int* s_col_ptr = reinterpret_cast<int*>(smem + L.col_ptr);
    float* s_weights = reinterpret_cast<float*>(smem + L.weights);

    const int tid = threadIdx.x;
    for (int i = tid; i < g.window; i += kThreads) s_window[i] = tab.window[i];
    for (int i = tid; i < 256; i += kThreads) { s_w256[i] = tab.w256[i]; s_unt[i] = tab.untangle[i]; }
    if (!kMagOnly) {
        for (int i = tid; i < g.n_mel; i += kThreads) {
            s_col_start[i] = tab.col_start[i]; s_col_count[i] = tab.col_count[i]; s_col_ptr[i] = tab.col_ptr[i];
        }
        for (int i = tid; i < g.nnz; i += kThreads) s_weights[i] = tab.weights[i];
    }

    const int lane16 = tid & 15;
    const int slot_id = tid >> 4;
    float2* slot = s_xchg + slot_id * 16 * kXchgStride;
    const int n_k2 = kMagOnly ? 16 : (g.max_bin >> 4) + 1;
    const int out_width = kMagOnly ? kBins : g.n_mel;

    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long clip = tile / tiles_per_clip;
        const int f0 = static_cast<int>(tile - clip * tiles_per_clip) * g.tile_frames;
        const int nf = static_cast<int>(min(static_cast<long long>(g.tile_frames), frames_per_clip - f0));
        const int span = (nf - 1) * g.hop + g.window;
        const float* src = wav + clip * wav_stride + static_cast<long long>(f0) * g.hop;
        __syncthreads();                                      // previous tile fully consumed / tables visible
        for (int i = tid; i < span; i += kThreads) s_samples[i] = __ldg(src + i);
        __syncthreads();
        float* out_tile = out + (clip * frames_per_clip + f0) * out_width;
        for (int fb = 0; fb < nf; fb += kSlots) {
            const int f = fb + slot_id;
            // the shuffles inside frame_spectrum use the full mask: an idle 16-lane group (ragged tile
            // end) still runs the FFT on frame 0 of the span and discards the result
            const bool active = f < nf;
            const int fs = active ? f * g.hop : 0;
            frame_spectrum<kMagOnly>(s_samples, fs, g.window, s_window, s_w256, s_unt, slot, lane16, n_k2);
            const float* mag = reinterpret_cast<const float*>(slot);
            if (kMagOnly) {
                if (active)
                    for (int k = lane16; k < kBins; k += 16) out_tile[static_cast<long long>(f) * kBins + k] = mag[k];
            } else if (active) {
                for (int c = lane16; c < g.n_mel; c += 16) {
                    const int b0 = s_col_start[c], cnt = s_col_count[c];
                    const float* w = s_weights + s_col_ptr[c];
                    float acc = 0.f;
                    for (int j = 0; j < cnt; ++j) acc = fmaf(mag[b0 + j], w[j], acc);
                    s_out[f * g.n_mel + c] = logf(acc + g.log_offset);
                }
            }
            __syncwarp();                                     // mag (slot) is overwritten by the next frame
        }
        if (!kMagOnly) {
            __syncthreads();
            const int n_out = nf * g.n_mel;
            if ((reinterpret_cast<uintptr_t>(out_tile) & 15) == 0 && (n_out & 3) == 0) {
                const float4* s4 = reinterpret_cast<const float4*>(s_out);
                float4* o4 = reinterpret_cast<float4*>(out_tile);
                for (int i = tid; i < n_out / 4; i += kThreads) __stcs(o4 + i, s4[i]);
            } else {
                for (int i = tid; i < n_out; i += kThreads) out_tile[i] = s_out[i];
            }
        }
    }
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// plan
// ---------------------------------------------------------------------------------------------
struct a2m_mel_plan {
    int device;
    int window, hop, nfft, n_mel, nnz, max_bin;
    float log_offset;
    void* blob;           // one device allocation holding all tables
    MelTables tab;
};

extern "C" int a2m_mel_plan_create(int window, int hop, int nfft, int n_mel, const double* hann_host,
                                   const double* mel_weights_host, double log_offset, int device,
                                   a2m_mel_plan** out) {
    A2M_ARG_CHECK(out != nullptr, "a2m_mel_plan_create: out is NULL");
    *out = nullptr;
    if (nfft != kNfft) {
        a2m_set_error("a2m_mel_plan_create: fft length %d not supported (this build implements nfft = %d, "
                      "i.e. 257 spectrogram bins)", nfft, kNfft);
        return A2M_ERR_UNSUPPORTED;
    }
    A2M_ARG_CHECK(window >= 1 && window <= nfft, "a2m_mel_plan_create: window %d must be in [1, %d]", window, nfft);
    A2M_ARG_CHECK(hop >= 1, "a2m_mel_plan_create: hop %d must be >= 1", hop);
    A2M_ARG_CHECK(n_mel >= 1 && n_mel <= kMaxMel, "a2m_mel_plan_create: n_mel %d must be in [1, %d]", n_mel, kMaxMel);
    A2M_ARG_CHECK(hann_host && mel_weights_host, "a2m_mel_plan_create: NULL table");
    A2M_ARG_CHECK(window + 0 <= kSpanFloats, "a2m_mel_plan_create: window too long");

    // column-compressed mel matrix; every column's support must be one contiguous run of bins
    std::vector<int> col_start(kMaxMel, 0), col_count(kMaxMel, 0), col_ptr(kMaxMel, 0);
    std::vector<float> weights;
    int max_bin = 0;
    for (int c = 0; c < n_mel; ++c) {
        int first = -1, last = -1;
        for (int k = 0; k < kBins; ++k)
            if (mel_weights_host[static_cast<size_t>(k) * n_mel + c] != 0.0) { if (first < 0) first = k; last = k; }
        col_ptr[c] = static_cast<int>(weights.size());
        if (first < 0) continue;                            // empty band: log(offset)
        if (last > 255) {
            a2m_set_error("a2m_mel_plan_create: mel column %d uses the Nyquist bin, not supported", c);
            return A2M_ERR_UNSUPPORTED;
        }
        col_start[c] = first;
        col_count[c] = last - first + 1;
        for (int k = first; k <= last; ++k) weights.push_back(static_cast<float>(mel_weights_host[static_cast<size_t>(k) * n_mel + c]));
        if (last > max_bin) max_bin = last;
    }
    const int nnz = static_cast<int>(weights.size());

    std::vector<float> win(window);
    for (int i = 0; i < window; ++i) win[i] = static_cast<float>(hann_host[i]);
    std::vector<float2> w256(256), unt(256);
    const double two_pi = 6.283185307179586476925286766559;
    for (int e = 0; e < 256; ++e) {
        w256[e] = make_float2(static_cast<float>(std::cos(two_pi * e / 256.0)), static_cast<float>(-std::sin(two_pi * e / 256.0)));
        unt[e] = make_float2(static_cast<float>(-std::sin(two_pi * e / 512.0)), static_cast<float>(-std::cos(two_pi * e / 512.0)));
    }

    A2M_CUDA_CHECK(cudaSetDevice(device));
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~size_t(255); return o; };
    const size_t o_win = carve(window * 4), o_w = carve(256 * 8), o_u = carve(256 * 8), o_cs = carve(kMaxMel * 4),
                 o_cc = carve(kMaxMel * 4), o_cp = carve(kMaxMel * 4), o_wt = carve((nnz + 4) * 4);
    unsigned char* blob = nullptr;
    A2M_CUDA_CHECK(cudaMalloc(&blob, off));
    std::vector<unsigned char> host(off, 0);
    memcpy(host.data() + o_win, win.data(), window * 4);
    memcpy(host.data() + o_w, w256.data(), 256 * 8);
    memcpy(host.data() + o_u, unt.data(), 256 * 8);
    memcpy(host.data() + o_cs, col_start.data(), kMaxMel * 4);
    memcpy(host.data() + o_cc, col_count.data(), kMaxMel * 4);
    memcpy(host.data() + o_cp, col_ptr.data(), kMaxMel * 4);
    if (nnz) memcpy(host.data() + o_wt, weights.data(), nnz * 4);
    cudaError_t e = cudaMemcpy(blob, host.data(), off, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(blob); a2m_set_error("a2m_mel_plan_create: upload failed: %s", cudaGetErrorString(e)); return (int)e; }

    a2m_mel_plan* p = new a2m_mel_plan();
    p->device = device; p->window = window; p->hop = hop; p->nfft = nfft; p->n_mel = n_mel; p->nnz = nnz;
    p->max_bin = max_bin; p->log_offset = static_cast<float>(log_offset); p->blob = blob;
    p->tab.window = reinterpret_cast<const float*>(blob + o_win);
    p->tab.w256 = reinterpret_cast<const float2*>(blob + o_w);
    p->tab.untangle = reinterpret_cast<const float2*>(blob + o_u);
    p->tab.col_start = reinterpret_cast<const int*>(blob + o_cs);
    p->tab.col_count = reinterpret_cast<const int*>(blob + o_cc);
    p->tab.col_ptr = reinterpret_cast<const int*>(blob + o_cp);
    p->tab.weights = reinterpret_cast<const float*>(blob + o_wt);
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

void a2m_count_launch();

template <bool kMagOnly>
static int launch_logmel(const a2m_mel_plan* plan, const float* wav, int64_t n_clips, int64_t n_samples,
                         int64_t wav_stride, float* out, void* stream, const char* who) {
    A2M_ARG_CHECK(plan != nullptr, "%s: plan is NULL", who);
    A2M_ARG_CHECK(n_clips >= 0 && n_samples >= 0, "%s: negative size", who);
    A2M_ARG_CHECK(n_clips <= 1 || wav_stride >= n_samples, "%s: wav_stride %lld < n_samples %lld", who, (long long)wav_stride, (long long)n_samples);
    const int64_t frames = a2m_mel_num_frames(plan, n_samples);
    A2M_ARG_CHECK(frames >= 0, "%s: negative dimensions are not allowed (%lld samples, window %d)", who,
                  (long long)n_samples, plan->window);
    if (n_clips == 0 || frames == 0) return A2M_OK;           // empty result, nothing to launch
    A2M_ARG_CHECK(wav != nullptr && out != nullptr, "%s: NULL buffer", who);

    MelGeom g;
    g.window = plan->window; g.hop = plan->hop; g.n_mel = plan->n_mel; g.nnz = plan->nnz;
    g.max_bin = plan->max_bin; g.log_offset = plan->log_offset;
    long long tf = 1 + (kSpanFloats - plan->window) / plan->hop;
    if (tf > kMaxTileFrames) tf = kMaxTileFrames;
    if (tf > frames) tf = frames;
    g.tile_frames = static_cast<int>(tf);
    const int tiles_per_clip = static_cast<int>((frames + tf - 1) / tf);
    const long long n_tiles = static_cast<long long>(tiles_per_clip) * n_clips;

    const SmemLayout L = smem_layout(g.n_mel, g.nnz, kMagOnly);
    static bool attr_set[2] = {false, false};
    if (!attr_set[kMagOnly]) {
        A2M_CUDA_CHECK(cudaFuncSetAttribute(logmel_kernel<kMagOnly>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        attr_set[kMagOnly] = true;
    }
    long long grid = 2LL * a2m_num_sms();
    if (grid > n_tiles) grid = n_tiles;
    logmel_kernel<kMagOnly><<<static_cast<unsigned>(grid), kThreads, L.total, static_cast<cudaStream_t>(stream)>>>(
        wav, n_clips, n_samples, wav_stride, frames, tiles_per_clip, n_tiles, plan->tab, g, out);
    a2m_count_launch();
    A2M_LAUNCH_CHECK();
    return A2M_OK;
}

extern "C" int a2m_logmel_f32(const a2m_mel_plan* plan, const float* wav, int64_t n_clips, int64_t n_samples,
                              int64_t wav_stride, float* out, void* stream) {
    return launch_logmel<false>(plan, wav, n_clips, n_samples, wav_stride, out, stream, "a2m_logmel_f32");
}

extern "C" int a2m_stft_magnitude_f32(const a2m_mel_plan* plan, const float* wav, int64_t n_clips,
                                      int64_t n_samples, int64_t wav_stride, float* out, void* stream) {
    return launch_logmel<true>(plan, wav, n_clips, n_samples, wav_stride, out, stream, "a2m_stft_magnitude_f32");
}
