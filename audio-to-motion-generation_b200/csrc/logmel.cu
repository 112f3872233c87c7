// Fused log-mel front end: framing -> periodic Hann -> 512-point real FFT ->
// |.| -> sparse mel projection -> log(. + offset), one launch for a whole batch of clips.
//
// Replaces pose_video/mel_features.py:192-223 (see include/a2m_b200.h).  HBM-bound by contract:
// every sample is read once (the 240-sample overlap between tiles is an L2 hit) and every output
// element written once; everything in between lives in registers / shared memory.
//
// Mapping: 16 threads own one frame (two register-resident radix-16 passes of a 256-point complex
// FFT of the even/odd-packed frame, one padded shared-memory exchange in between; the spectrum Z is
// dumped to shared memory so the partner bin of the real-input untangle is one conflict-free load).
// A 256-thread CTA works on a tile of 16 consecutive frames of one clip; a tile's samples are staged
// in shared memory with cp.async into one of two buffers while the previous tile is transformed, and
// results are staged and written back as whole 256-byte rows.
#include <cmath>
#include <vector>
#include <cstring>
#include "a2m_common.cuh"
#include "fft_math.cuh"

using a2m_fft::cpx;

namespace {

constexpr int kThreads = 256;
constexpr int kSlots = kThreads / 16;         // frames in flight per CTA = frames per tile
constexpr int kSpanFloats = 2816;             // staged samples per tile buffer (15 * 160 + 416), two buffers
constexpr int kXchgStride = 18;               // float2 units; 144 B rows keep LDS.128 conflict-free
constexpr int kSlotZ = 16 * kXchgStride;      // float2 per slot: exchange rows, later Z[0..255]
constexpr int kSlotMag = 260;                 // floats per slot: |X[k]|, k = 0..256
constexpr int kMaxMel = 128;
constexpr int kNfft = 512;
constexpr int kBins = kNfft / 2 + 1;

struct MelTables {                // device-resident constants of a plan
    const float* window;          // [512]  periodic Hann, zero padded past the window length
    const float2* tw;             // [16][16] exp(-2 pi i m2 k1 / 256) at [k1][m2]
    const float2* untangle;       // [256]  (-sin, -cos)(2 pi k / 512)
    const int4* col_meta;         // [n_mel] {first bin (multiple of 4), 4-bin groups, offset into weights (multiple of 4), 0}
    const float* weights;         // [nnz]
};

struct MelGeom {
    int window, hop, n_mel, nnz;
    int tile_frames;              // frames per tile (<= kSlots)
    int max_bin;                  // highest spectrogram bin with a non-zero mel weight
    float log_offset;
    int log_mode;                 // 0: log(x + offset) (mel_features.py:223); 1: log(x == 0 ? offset : x) (pats/data_loading/audio.py:117-119)
};

struct SmemLayout {
    int samples, slot_z, slot_mag, out, window, tw, untangle, col_meta, weights, total;
};

__host__ __device__ inline SmemLayout smem_layout(int n_mel, int nnz, bool mag_only) {
    SmemLayout L;
    int off = 0;
    L.samples = off;   off += 2 * kSpanFloats * 4;
    L.slot_z = off;    off += kSlots * kSlotZ * 8;
    L.slot_mag = off;  off += kSlots * kSlotMag * 4;
    L.out = off;       off += mag_only ? 0 : kSlots * n_mel * 4;
    L.window = off;    off += kNfft * 4;
    L.tw = off;        off += 256 * 8;
    L.untangle = off;  off += 256 * 8;
    L.col_meta = off;  off += kMaxMel * 16;
    L.weights = off;   off += ((nnz + 3) / 4) * 16;
    L.total = off;
    return L;
}

__device__ __forceinline__ float fast_sqrt(float x) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void cp_async4(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(a2m::smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// One frame by 16 lanes.  `lane16` = position in the 16-lane group, s = first sample of the frame in the staged
// span.  On return mag[k] = |X[k]| for k < 16 * n_k2 (and k = 256 when kMagOnly).
//   pass 1: lane m2 holds z[16 m1 + m2] = (x[32 m1 + 2 m2], x[.. + 1]) * hann, DFT16 over m1, * W256^(m2 k1)
//   exchange through shared memory ([k1][m2], padded rows)
//   pass 2: lane k1, DFT16 over m2 -> Z[k1 + 16 k2]
//   untangle: Z is dumped to shared memory so that the partner bin Z[(256 - k) & 255] is one conflict-free load
template <bool kMagOnly>
__device__ __forceinline__ void frame_spectrum(const float* __restrict__ s, int window, bool even_hop,
                                               const float* __restrict__ s_window, const float2* __restrict__ s_tw,
                                               const float2* __restrict__ s_unt, float2* __restrict__ zs,
                                               float* __restrict__ mag, int lane16, int n_k2) {
    cpx v[16];
    const int n_m1 = (window + 31) >> 5;                    // rows of the 16 x 16 input that are not all padding
#pragma unroll
    for (int m1 = 0; m1 < 16; ++m1) {
        if (m1 < n_m1) {                                    // uniform
            const int n = 32 * m1 + 2 * lane16;
            const float2 w = *reinterpret_cast<const float2*>(s_window + n);          // zero past the window
            float2 x;
            if (even_hop) x = *reinterpret_cast<const float2*>(s + n);
            else x = make_float2(s[n], s[n + 1]);
            v[m1] = a2m_fft::make(x.x * w.x, x.y * w.y);
        } else {
            v[m1] = a2m_fft::make(0.f, 0.f);
        }
    }
    a2m_fft::dft16(v);                                      // Y[m2][k1], k1 = register index
#pragma unroll
    for (int k1 = 1; k1 < 16; ++k1) {                       // * W256^(m2 k1)
        const float2 w = s_tw[k1 * 16 + lane16];
        v[k1] = a2m_fft::mul(v[k1], a2m_fft::make(w.x, w.y));
    }
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) zs[k1 * kXchgStride + lane16] = make_float2(v[k1].x, v[k1].y);
    __syncwarp();
    {
        const float4* row = reinterpret_cast<const float4*>(zs + lane16 * kXchgStride);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 q = row[j];
            v[2 * j] = a2m_fft::make(q.x, q.y);
            v[2 * j + 1] = a2m_fft::make(q.z, q.w);
        }
    }
    __syncwarp();                                           // the exchange rows are overwritten by Z below
    a2m_fft::dft16(v);                                      // v[k2] = Z[lane16 + 16 k2]
#pragma unroll
    for (int k2 = 0; k2 < 16; ++k2) zs[lane16 + 16 * k2] = make_float2(v[k2].x, v[k2].y);
    __syncwarp();
    if (kMagOnly && lane16 == 0) mag[256] = fabsf(v[0].x - v[0].y);      // X[256] = Re Z0 - Im Z0
#pragma unroll
    for (int k2 = 0; k2 < 16; ++k2) {
        if (k2 < n_k2) {                                    // uniform
            const int k = lane16 + 16 * k2;
            const float2 pz = zs[(256 - k) & 255];          // k = 0 pairs with itself
            const float2 t = s_unt[k];
            const cpx x2 = a2m_fft::untangle2(v[k2], a2m_fft::make(pz.x, pz.y), a2m_fft::make(t.x, t.y));
            mag[k] = 0.5f * fast_sqrt(x2.x * x2.x + x2.y * x2.y);
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
    float* s_out = reinterpret_cast<float*>(smem + L.out);
    float* s_window = reinterpret_cast<float*>(smem + L.window);
    float2* s_tw = reinterpret_cast<float2*>(smem + L.tw);
    float2* s_unt = reinterpret_cast<float2*>(smem + L.untangle);
    int4* s_meta = reinterpret_cast<int4*>(smem + L.col_meta);
    float* s_weights = reinterpret_cast<float*>(smem + L.weights);

    const int tid = threadIdx.x;
    for (int i = tid; i < kNfft; i += kThreads) s_window[i] = tab.window[i];
    for (int i = tid; i < 256; i += kThreads) { s_tw[i] = tab.tw[i]; s_unt[i] = tab.untangle[i]; }
    if (!kMagOnly) {
        for (int i = tid; i < g.n_mel; i += kThreads) s_meta[i] = tab.col_meta[i];
        for (int i = tid; i < g.nnz; i += kThreads) s_weights[i] = tab.weights[i];
    }
    // samples read past a tile's span are multiplied by the zero padding of the window: keep them finite
    for (int i = tid; i < 2 * kSpanFloats; i += kThreads) s_samples[i] = 0.f;

    const int lane16 = tid & 15;
    const int slot_id = tid >> 4;
    float2* zs = reinterpret_cast<float2*>(smem + L.slot_z) + slot_id * kSlotZ;
    float* mag = reinterpret_cast<float*>(smem + L.slot_mag) + slot_id * kSlotMag;
    const int n_k2 = kMagOnly ? 16 : (g.max_bin >> 4) + 1;
    const int out_width = kMagOnly ? kBins : g.n_mel;
    const bool even_hop = (g.hop & 1) == 0;

    auto stage = [&](long long tile, int buf) {             // asynchronous copy of one tile's samples
        const unsigned clip = static_cast<unsigned>(tile) / static_cast<unsigned>(tiles_per_clip);     // n_tiles < 2^31
        const int f0 = static_cast<int>(static_cast<unsigned>(tile) - clip * tiles_per_clip) * g.tile_frames;
        const int nf = min(g.tile_frames, static_cast<int>(frames_per_clip) - f0);
        const int span = (nf - 1) * g.hop + g.window;
        const float* src = wav + static_cast<long long>(clip) * wav_stride + static_cast<long long>(f0) * g.hop;
        float* dst = s_samples + buf * kSpanFloats;
        for (int i = tid; i < span; i += kThreads) cp_async4(dst + i, src + i);
        cp_async_commit();
    };

    __syncthreads();                                        // tables and the zero fill are in place
    long long tile = blockIdx.x;
    if (tile < n_tiles) stage(tile, 0);
    int buf = 0;
    for (; tile < n_tiles; tile += gridDim.x, buf ^= 1) {
        const unsigned clip = static_cast<unsigned>(tile) / static_cast<unsigned>(tiles_per_clip);
        const int f0 = static_cast<int>(static_cast<unsigned>(tile) - clip * tiles_per_clip) * g.tile_frames;
        const int nf = min(g.tile_frames, static_cast<int>(frames_per_clip) - f0);
        cp_async_wait_all();
        __syncthreads();                                    // this tile's samples landed; the other buffer and s_out are free
        if (tile + gridDim.x < n_tiles) stage(tile + gridDim.x, buf ^ 1);      // overlaps the FFTs below
        float* out_tile = out + (static_cast<long long>(clip) * frames_per_clip + f0) * out_width;
        // the warp-level syncs inside frame_spectrum need every lane: an idle 16-lane group (ragged tile end)
        // runs the FFT on frame 0 of the span and discards the result
        const bool active = slot_id < nf;
        const int fs = active ? slot_id * g.hop : 0;
        frame_spectrum<kMagOnly>(s_samples + buf * kSpanFloats + fs, g.window, even_hop, s_window, s_tw, s_unt, zs, mag,
                                 lane16, n_k2);
        if (kMagOnly) {
            if (active)
                for (int k = lane16; k < kBins; k += 16) out_tile[static_cast<long long>(slot_id) * kBins + k] = mag[k];
        } else {
            if (active) {
                for (int c = lane16; c < g.n_mel; c += 16) {
                    // column support padded to whole 4-bin groups (zero weights): two 16-byte loads per 4 FMAs
                    const int4 m = s_meta[c];                               // {first bin (multiple of 4), groups, weight offset, 0}
                    const float4* mg = reinterpret_cast<const float4*>(mag + m.x);
                    const float4* w = reinterpret_cast<const float4*>(s_weights + m.z);
                    float acc0 = 0.f, acc1 = 0.f;
                    for (int j = 0; j < m.y; ++j) {
                        const float4 a = mg[j], b = w[j];
                        acc0 = fmaf(a.x, b.x, acc0); acc1 = fmaf(a.y, b.y, acc1);
                        acc0 = fmaf(a.z, b.z, acc0); acc1 = fmaf(a.w, b.w, acc1);
                    }
                    const float e = acc0 + acc1;
                    s_out[slot_id * g.n_mel + c] = __logf(g.log_mode ? (e == 0.f ? g.log_offset : e) : e + g.log_offset);
                }
            }
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
    cp_async_wait_all();
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// plan
// ---------------------------------------------------------------------------------------------
struct a2m_mel_plan {
    int device;
    int window, hop, nfft, n_mel, nnz, max_bin;
    float log_offset;
    int log_mode;
    void* blob;           // one device allocation holding all tables
    MelTables tab;
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

    // column-compressed mel matrix; every column's support must be one contiguous run of bins.  Each run is padded
    // (zero weights) to start and end on a multiple of 4 bins so the kernel reads magnitudes and weights as float4
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
        const int lo = first & ~3, hi = (last | 3);         // hi <= 255
        col_start[c] = lo;
        col_count[c] = (hi - lo + 1) / 4;                   // groups of 4 bins
        for (int k = lo; k <= hi; ++k) weights.push_back(static_cast<float>(mel_weights_host[static_cast<size_t>(k) * n_mel + c]));
        if (hi > max_bin) max_bin = hi;
    }
    const int nnz = static_cast<int>(weights.size());

    std::vector<float> win(kNfft, 0.f);                     // zero padded: samples past the window contribute nothing
    for (int i = 0; i < window; ++i) win[i] = static_cast<float>(hann_host[i]);
    std::vector<float2> tw(256), unt(256);
    const double two_pi = 6.283185307179586476925286766559;
    for (int k1 = 0; k1 < 16; ++k1)
        for (int m2 = 0; m2 < 16; ++m2) {                   // [k1][m2]: a 16-lane group reads one contiguous row
            const int e = (k1 * m2) & 255;
            tw[k1 * 16 + m2] = make_float2(static_cast<float>(std::cos(two_pi * e / 256.0)), static_cast<float>(-std::sin(two_pi * e / 256.0)));
        }
    for (int e = 0; e < 256; ++e)
        unt[e] = make_float2(static_cast<float>(-std::sin(two_pi * e / 512.0)), static_cast<float>(-std::cos(two_pi * e / 512.0)));
    std::vector<int4> meta(kMaxMel);
    for (int c = 0; c < kMaxMel; ++c) meta[c] = make_int4(col_start[c], col_count[c], col_ptr[c], 0);

    A2M_CUDA_CHECK(cudaSetDevice(device));
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~size_t(255); return o; };
    const size_t o_win = carve(kNfft * 4), o_w = carve(256 * 8), o_u = carve(256 * 8), o_cm = carve(kMaxMel * 16),
                 o_wt = carve((nnz + 4) * 4);
    unsigned char* blob = nullptr;
    A2M_CUDA_CHECK(cudaMalloc(&blob, off));
    std::vector<unsigned char> host(off, 0);
    memcpy(host.data() + o_win, win.data(), kNfft * 4);
    memcpy(host.data() + o_w, tw.data(), 256 * 8);
    memcpy(host.data() + o_u, unt.data(), 256 * 8);
    memcpy(host.data() + o_cm, meta.data(), kMaxMel * 16);
    if (nnz) memcpy(host.data() + o_wt, weights.data(), nnz * 4);
    cudaError_t e = cudaMemcpy(blob, host.data(), off, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(blob); a2m_set_error("a2m_mel_plan_create: upload failed: %s", cudaGetErrorString(e)); return (int)e; }

    a2m_mel_plan* p = new a2m_mel_plan();
    p->device = device; p->window = window; p->hop = hop; p->nfft = nfft; p->n_mel = n_mel; p->nnz = nnz;
    p->max_bin = max_bin; p->log_offset = static_cast<float>(log_offset); p->log_mode = log_mode; p->blob = blob;
    p->tab.window = reinterpret_cast<const float*>(blob + o_win);
    p->tab.tw = reinterpret_cast<const float2*>(blob + o_w);
    p->tab.untangle = reinterpret_cast<const float2*>(blob + o_u);
    p->tab.col_meta = reinterpret_cast<const int4*>(blob + o_cm);
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
    g.max_bin = plan->max_bin; g.log_offset = plan->log_offset; g.log_mode = plan->log_mode;
    // a frame's reads extend to the next multiple of 32 past its window (zero window weights there)
    long long tf = 1 + (kSpanFloats - ((plan->window + 31) / 32) * 32) / plan->hop;
    if (tf > kSlots) tf = kSlots;
    if (tf > frames) tf = frames;
    g.tile_frames = static_cast<int>(tf);
    const int tiles_per_clip = static_cast<int>((frames + tf - 1) / tf);
    const long long n_tiles = static_cast<long long>(tiles_per_clip) * n_clips;

    A2M_ARG_CHECK(n_tiles <= 0x7fffffffLL && frames <= 0x7fffffffLL, "%s: %lld tiles", who, n_tiles);
    const SmemLayout L = smem_layout(g.n_mel, g.nnz, kMagOnly);
    static A2mPerDeviceOnce attr_set;                       // one instance per template instantiation
    if (attr_set.first()) {
        A2M_CUDA_CHECK(cudaFuncSetAttribute(logmel_kernel<kMagOnly>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
    }
    long long grid = 2LL * a2m_num_sms();
    if (grid > n_tiles) grid = n_tiles;
    A2M_ARG_CHECK(L.total <= 110 * 1024, "%s: %d bytes of shared memory", who, L.total);
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
