// Band-limited sinc resampling: the first step of the PATS front end log_mel_400,
// pats/data_loading/audio.py:87  y = librosa.core.resample(y, orig_sr=sr, target_sr=16000).
//
// librosa hands this to resampy ('kaiser_best', its default up to 0.9); both are third-party, absent and unpinned in the
// reference, so the algorithm is restated from resampy's published one (oracle/pats_oracle.py, parity unpinned):
// a Kaiser-windowed sinc table (64 zero crossings, 512 entries per crossing) interpolated linearly between entries,
//   y[t] = sum_i (win[off + i step] + eta dwin[off + i step]) x[n - i]  +  the mirrored right wing over x[n + 1 + k],
//   n = floor(t / ratio), step = int(min(1, ratio) * 512); output length ceil(n_in * ratio) (librosa's fix_length),
//   samples past resampy's int(n_in * ratio) are zero.
// One thread per output sample; the position t / ratio is evaluated in fp64, the taps in fp32 with the table in L1/L2.
// Downsampling 44.1 kHz -> 16 kHz is 2 x 177 taps per output sample: ~17 MFLOP per 4.27 s clip, FP32-bound, far from HBM.
#include <cmath>
#include <vector>
#include "a2m_common.cuh"

void a2m_count_launch();

struct a2m_resample_plan {
    int device;
    double ratio;
    float scale;
    int num_table, index_step, n_win;
    float* table;            // [2][n_win]: window (scaled by ratio when downsampling), then its forward difference
};

namespace {

__global__ void __launch_bounds__(256)
resample_kernel(const float* __restrict__ x, long long n_clips, long long n_in, long long in_stride, long long n_out,
                long long n_valid, double inv_ratio, float scale, int num_table, int index_step, int n_win,
                const float* __restrict__ win, const float* __restrict__ dwin, float* __restrict__ out) {
    const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= n_clips * n_out) return;
    const long long clip = idx / n_out, t = idx - clip * n_out;
    if (t >= n_valid) { out[idx] = 0.f; return; }
    const float* xc = x + clip * in_stride;
    const double time_register = static_cast<double>(t) * inv_ratio;
    const long long n = static_cast<long long>(time_register);
    float frac = scale * static_cast<float>(time_register - static_cast<double>(n));
    float index_frac = frac * num_table;
    int offset = static_cast<int>(index_frac);
    float eta = index_frac - offset;
    float acc = 0.f;
    long long i_max = (n_win - offset) / index_step;
    if (i_max > n + 1) i_max = n + 1;
    for (long long i = 0; i < i_max; ++i) {
        const int k = offset + static_cast<int>(i) * index_step;
        acc = fmaf(fmaf(eta, __ldg(dwin + k), __ldg(win + k)), __ldg(xc + n - i), acc);
    }
    frac = scale - frac;
    index_frac = frac * num_table;
    offset = static_cast<int>(index_frac);
    eta = index_frac - offset;
    long long k_max = (n_win - offset) / index_step;
    if (k_max > n_in - n - 1) k_max = n_in - n - 1;
    for (long long k = 0; k < k_max; ++k) {
        const int j = offset + static_cast<int>(k) * index_step;
        acc = fmaf(fmaf(eta, __ldg(dwin + j), __ldg(win + j)), __ldg(xc + n + k + 1), acc);
    }
    out[idx] = acc;
}

}  // namespace

extern "C" int a2m_resample_plan_create(double orig_sr, double target_sr, int device, a2m_resample_plan** out) {
    A2M_ARG_CHECK(out != nullptr, "a2m_resample_plan_create: out is NULL");
    *out = nullptr;
    A2M_ARG_CHECK(orig_sr > 0 && target_sr > 0 && target_sr / orig_sr <= 64.0 && orig_sr / target_sr <= 64.0,
                  "a2m_resample_plan_create: sample rates %g -> %g", orig_sr, target_sr);
    // resampy.filters.sinc_window(num_zeros=64, precision=9, window=kaiser(beta), rolloff): 'kaiser_best'
    const int num_zeros = 64, precision = 9;
    const double beta = 14.769656459379492, rolloff = 0.9475937167399596, pi = 3.14159265358979323846;
    const int num_bits = 1 << precision, n = num_bits * num_zeros, n_win = n + 1;
    const double ratio = target_sr / orig_sr;
    auto bessel_i0 = [](double v) {                     // power series, converges fast for the arguments met here
        double sum = 1.0, term = 1.0;
        for (int k = 1; k < 200; ++k) { term *= (v / (2.0 * k)) * (v / (2.0 * k)); sum += term; if (term < 1e-18 * sum) break; }
        return sum;
    };
    std::vector<double> w(n_win);
    const double i0b = bessel_i0(beta);
    for (int i = 0; i <= n; ++i) {
        const double tpos = rolloff * (static_cast<double>(num_zeros) * i / n);      // rolloff * linspace(0, num_zeros, n + 1)
        const double sinc = tpos == 0.0 ? 1.0 : std::sin(pi * tpos) / (pi * tpos);
        const double u = static_cast<double>(i) / n;                                   // np.kaiser(2n + 1, beta)[n + i]
        const double taper = bessel_i0(beta * std::sqrt(std::max(0.0, 1.0 - u * u))) / i0b;
        w[i] = taper * rolloff * sinc * (ratio < 1.0 ? ratio : 1.0);
    }
    std::vector<float> host(2 * static_cast<size_t>(n_win), 0.f);
    for (int i = 0; i < n_win; ++i) host[i] = static_cast<float>(w[i]);
    for (int i = 0; i + 1 < n_win; ++i) host[n_win + i] = static_cast<float>(w[i + 1] - w[i]);
    A2M_CUDA_CHECK(cudaSetDevice(device));
    float* table = nullptr;
    A2M_CUDA_CHECK(cudaMalloc(&table, host.size() * sizeof(float)));
    cudaError_t e = cudaMemcpy(table, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(table); a2m_set_error("a2m_resample_plan_create: %s", cudaGetErrorString(e)); return (int)e; }
    a2m_resample_plan* p = new a2m_resample_plan();
    p->device = device; p->ratio = ratio; p->scale = static_cast<float>(ratio < 1.0 ? ratio : 1.0);
    p->num_table = num_bits; p->index_step = static_cast<int>((ratio < 1.0 ? ratio : 1.0) * num_bits); p->n_win = n_win;
    p->table = table;
    if (p->index_step < 1) { cudaFree(table); delete p; a2m_set_error("a2m_resample_plan_create: ratio %g too small", ratio); return A2M_ERR_UNSUPPORTED; }
    *out = p;
    return A2M_OK;
}

extern "C" void a2m_resample_plan_destroy(a2m_resample_plan* plan) {
    if (!plan) return;
    cudaFree(plan->table);
    delete plan;
}

extern "C" int64_t a2m_resample_out_length(const a2m_resample_plan* plan, int64_t n_samples) {
    if (!plan || n_samples < 0) return -1;
    return static_cast<int64_t>(std::ceil(static_cast<double>(n_samples) * plan->ratio));
}

extern "C" int a2m_resample_f32(const a2m_resample_plan* plan, const float* wav, int64_t n_clips, int64_t n_samples,
                                int64_t wav_stride, float* out, void* stream) {
    A2M_ARG_CHECK(plan != nullptr, "a2m_resample_f32: plan is NULL");
    A2M_ARG_CHECK(n_clips >= 0 && n_samples >= 0, "a2m_resample_f32: negative size");
    A2M_ARG_CHECK(n_clips <= 1 || wav_stride >= n_samples, "a2m_resample_f32: wav_stride %lld < n_samples %lld",
                  (long long)wav_stride, (long long)n_samples);
    const long long n_out = a2m_resample_out_length(plan, n_samples);
    if (n_clips == 0 || n_out == 0) return A2M_OK;
    A2M_ARG_CHECK(wav != nullptr && out != nullptr, "a2m_resample_f32: NULL buffer");
    const long long n_valid = static_cast<long long>(static_cast<double>(n_samples) * plan->ratio);     // resampy's int(n * ratio)
    const long long total = n_clips * n_out;
    A2M_ARG_CHECK(total / 256 < 0x7fffffffLL, "a2m_resample_f32: %lld output samples", total);
    resample_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        wav, n_clips, n_samples, wav_stride, n_out, n_valid, 1.0 / plan->ratio, plan->scale, plan->num_table, plan->index_step,
        plan->n_win, plan->table, plan->table + plan->n_win, out);
    a2m_count_launch();
    A2M_LAUNCH_CHECK();
    return A2M_OK;
}
