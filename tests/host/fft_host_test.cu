// CPU emulation of the 512-point log-mel kernel's per-tile algorithm (csrc/logmel.cu): 16 "lanes" own a PAIR of
// frames, the same fft_math.cuh building blocks (compiled here with plain-C++ pairs instead of the packed
// FADD2 / FMUL2 / FFMA2 registers), the same exchange layout and the same partner rule of the untangle
// (lane l <-> lane 16 - l swap registers 8..15, lane 0 pairs with itself one register further on).
// Used by tests/test_fft_host.py to validate the FFT factorisation, the index algebra and its fp32 accuracy
// without a GPU.
//   usage: fft_host_test <wav.f32> <n_samples> <window> <hop> <hann.f64> <out.f32 [frames,257]>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define A2M_FFT_HOST_EMULATION 1
#include "../../audio-to-motion-generation_b200/csrc/fft_math.cuh"

using namespace a2m_fft;

int main(int argc, char** argv) {
    if (argc != 7) return 2;
    const int n = atoi(argv[2]), window = atoi(argv[3]), hop = atoi(argv[4]);
    std::vector<float> wav(n + 1024, 0.f);              // the kernel reads past the last window (times window zeros)
    std::vector<double> hann(window);
    FILE* f = fopen(argv[1], "rb"); if (!f || fread(wav.data(), 4, n, f) != (size_t)n) return 3; fclose(f);
    f = fopen(argv[5], "rb"); if (!f || fread(hann.data(), 8, window, f) != (size_t)window) return 3; fclose(f);
    const int frames = 1 + (n - window) / hop;
    const double two_pi = 6.283185307179586476925286766559;
    std::vector<float> win(512, 0.f);
    for (int i = 0; i < window; ++i) win[i] = (float)hann[i];
    float twx[256], twy[256], ux[256], uy[256];
    for (int k1 = 0; k1 < 16; ++k1)
        for (int m2 = 0; m2 < 16; ++m2) {
            const int e = (k1 * m2) & 255;
            twx[k1 * 16 + m2] = (float)std::cos(two_pi * e / 256.0);
            twy[k1 * 16 + m2] = (float)-std::sin(two_pi * e / 256.0);
        }
    for (int e = 0; e < 256; ++e) {
        ux[e] = (float)-std::sin(two_pi * e / 512.0);
        uy[e] = (float)-std::cos(two_pi * e / 512.0);
    }
    const int n_m1 = (window + 31) / 32;
    std::vector<float> out((size_t)frames * 257);
    for (int fa = 0; fa < frames; fa += 2) {             // frames (A, B) = (fa, fa + 1); B may not exist
        const float* sa = wav.data() + (size_t)fa * hop;
        const float* sb = sa + hop;
        cpx xchg[16][16];   // [k1][m2]
        cpx Z[16][16];      // [lane][k2]
        for (int lane = 0; lane < 16; ++lane) {            // pass 1
            cpx v[16];
            for (int m1 = 0; m1 < 16; ++m1) {
                if (m1 < n_m1) {
                    const int i = 32 * m1 + 2 * lane;
                    v[m1] = make(pack(sa[i] * win[i], sb[i] * win[i]), pack(sa[i + 1] * win[i + 1], sb[i + 1] * win[i + 1]));
                } else {
                    v[m1] = make(pack(0.f, 0.f), pack(0.f, 0.f));
                }
            }
            dft16(v);
            for (int k1 = 1; k1 < 16; ++k1) v[k1] = mul_scalar(v[k1], twx[k1 * 16 + lane], twy[k1 * 16 + lane]);
            for (int k1 = 0; k1 < 16; ++k1) xchg[k1][lane] = v[k1];
        }
        for (int lane = 0; lane < 16; ++lane) {            // pass 2
            cpx v[16];
            for (int m2 = 0; m2 < 16; ++m2) v[m2] = xchg[lane][m2];
            dft16(v);
            for (int k2 = 0; k2 < 16; ++k2) Z[lane][k2] = v[k2];
        }
        float mag[2][258];
        for (int lane = 0; lane < 16; ++lane) {            // untangle with the kernel's partner rule
            const int pl = (16 - lane) & 15;
            if (lane == 0) {
                const pair_t sq = fma2(Z[0][8].im, Z[0][8].im, mul2(Z[0][8].re, Z[0][8].re));
                mag[0][128] = 2.f * std::sqrt(lo(sq));
                mag[1][128] = 2.f * std::sqrt(hi(sq));
            }
            for (int j = 0; j < 8; ++j) {
                const int i = 15 - j;
                // what lane `pl` sends in slot i
                const cpx zp = pl == 0 ? (i == 15 ? Z[0][0] : Z[0][i + 1]) : Z[pl][i];
                const int k = lane + 16 * j;
                pair_t sq_k, sq_m;
                untangle_pair_sq(Z[lane][j], zp, ux[k], uy[k], sq_k, sq_m);
                mag[0][k] = std::sqrt(lo(sq_k)); mag[1][k] = std::sqrt(hi(sq_k));
                mag[0][256 - k] = std::sqrt(lo(sq_m)); mag[1][256 - k] = std::sqrt(hi(sq_m));
            }
        }
        for (int h = 0; h < 2 && fa + h < frames; ++h)
            for (int k = 0; k <= 256; ++k) out[(size_t)(fa + h) * 257 + k] = 0.5f * mag[h][k];   // the kernel folds 0.5 into the mel weights
    }
    f = fopen(argv[6], "wb"); fwrite(out.data(), 4, out.size(), f); fclose(f);
    return 0;
}
