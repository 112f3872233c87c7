// CPU emulation of the log-mel kernel's per-frame algorithm (16 "lanes" per frame, the same
// fft_math.cuh building blocks and the same index algebra as csrc/logmel.cu), used by
// tests/test_fft_host.py to validate the FFT factorisation and its fp32 accuracy without a GPU.
//   usage: fft_host_test <wav.f32> <n_samples> <window> <hop> <hann.f64> <out.f32 [frames,257]>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../audio-to-motion-generation_b200/csrc/fft_math.cuh"

using namespace a2m_fft;

int main(int argc, char** argv) {
    if (argc != 7) return 2;
    const int n = atoi(argv[2]), window = atoi(argv[3]), hop = atoi(argv[4]);
    std::vector<float> wav(n);
    std::vector<double> hann(window);
    FILE* f = fopen(argv[1], "rb"); if (!f || fread(wav.data(), 4, n, f) != (size_t)n) return 3; fclose(f);
    f = fopen(argv[5], "rb"); if (!f || fread(hann.data(), 8, window, f) != (size_t)window) return 3; fclose(f);
    const int frames = 1 + (n - window) / hop;
    const double two_pi = 6.283185307179586476925286766559;
    std::vector<cpx> w256(256), unt(256);
    for (int e = 0; e < 256; ++e) {
        w256[e] = make((float)std::cos(two_pi * e / 256.0), (float)-std::sin(two_pi * e / 256.0));
        unt[e] = make((float)-std::sin(two_pi * e / 512.0), (float)-std::cos(two_pi * e / 512.0));
    }
    std::vector<float> out((size_t)frames * 257);
    for (int fr = 0; fr < frames; ++fr) {
        const float* s = wav.data() + (size_t)fr * hop;
        cpx xchg[16][16];   // [k1][m2]
        cpx Z[16][16];      // [lane k1][k2]
        for (int lane = 0; lane < 16; ++lane) {            // pass 1
            cpx v[16];
            for (int m1 = 0; m1 < 16; ++m1) {
                const int i = 32 * m1 + 2 * lane;
                float a = 0.f, b = 0.f;
                if (i < window) a = s[i] * (float)hann[i];
                if (i + 1 < window) b = s[i + 1] * (float)hann[i + 1];
                v[m1] = make(a, b);
            }
            dft16(v);
            for (int k1 = 1; k1 < 16; ++k1) v[k1] = mul(v[k1], w256[(lane * k1) & 255]);
            for (int k1 = 0; k1 < 16; ++k1) xchg[k1][lane] = v[k1];
        }
        for (int lane = 0; lane < 16; ++lane) {            // pass 2
            cpx v[16];
            for (int m2 = 0; m2 < 16; ++m2) v[m2] = xchg[lane][m2];
            dft16(v);
            for (int k2 = 0; k2 < 16; ++k2) Z[lane][k2] = v[k2];
        }
        float* o = out.data() + (size_t)fr * 257;
        o[256] = std::fabs(Z[0][0].x - Z[0][0].y);
        for (int lane = 0; lane < 16; ++lane) {            // untangle with the kernel's partner rule
            for (int k2 = 0; k2 < 16; ++k2) {
                const int k = lane + 16 * k2;
                const int pk = (256 - k) & 255;             // the kernel reads Z[(256 - k) & 255] from its shared-memory dump
                const cpx p = Z[pk & 15][pk >> 4];
                o[k] = half_magnitude(untangle2(Z[lane][k2], p, unt[k]));
            }
        }
    }
    f = fopen(argv[6], "wb"); fwrite(out.data(), 4, out.size(), f); fclose(f);
    return 0;
}
