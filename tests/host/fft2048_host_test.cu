// Host emulation of melspec2048_kernel's index algebra (csrc/melspec_wide.cu): the 16 x 16 x 4 split of the 1024-point
// complex FFT of an even/odd-packed 2048-sample frame, the two exchange layouts, and the untangle into |X[k]|^2 --
// checked against a direct fp64 DFT.  Build + run: tests/test_fft_host.py (g++/nvcc, no GPU).
#define A2M_FFT_HOST_EMULATION
#include <cmath>
#include <cstdio>
#include <vector>
#include "../../audio-to-motion-generation_b200/csrc/fft_math.cuh"

using namespace a2m_fft;

static int z_unit(int k) { return k + 2 * (k >> 6); }

int main() {
    const int kC = 1024, kRow = 68;
    const double two_pi = 6.283185307179586476925286766559;
    std::vector<float> xa(2048), xb(2048);
    unsigned s = 12345u;
    for (int i = 0; i < 2048; ++i) {
        s = s * 1664525u + 1013904223u; xa[i] = ((s >> 8) & 0xffff) / 65536.f - 0.5f;
        s = s * 1664525u + 1013904223u; xb[i] = ((s >> 8) & 0xffff) / 65536.f - 0.5f;
    }
    std::vector<cpx> region(16 * kRow);
    std::vector<std::vector<cpx>> regs(64, std::vector<cpx>(16));
    // pass A
    for (int t = 0; t < 64; ++t) {
        cpx z[16];
        for (int m = 0; m < 16; ++m) {
            const int n = t + 64 * m;
            z[m] = make(pack(xa[2 * n], xb[2 * n]), pack(xa[2 * n + 1], xb[2 * n + 1]));
        }
        dft16(z);
        for (int k1 = 1; k1 < 16; ++k1) {
            const double a = two_pi * ((t * k1) % 1024) / 1024.0;
            z[k1] = mul_scalar(z[k1], (float)std::cos(a), (float)-std::sin(a));
        }
        for (int k1 = 0; k1 < 16; ++k1) region[k1 * kRow + t] = z[k1];
    }
    // pass B
    for (int t = 0; t < 64; ++t) {
        const int k1b = t >> 2, vb = t & 3;
        cpx z[16];
        for (int u = 0; u < 16; ++u) z[u] = region[k1b * kRow + 4 * u + vb];
        dft16(z);
        if (vb != 0)
            for (int q = 1; q < 16; ++q) {
                const double a = two_pi * ((vb * q) % 64) / 64.0;
                z[q] = mul_scalar(z[q], (float)std::cos(a), (float)-std::sin(a));
            }
        for (int q = 0; q < 16; ++q) regs[t][q] = z[q];
    }
    for (int t = 0; t < 64; ++t) {
        const int k1b = t >> 2, vb = t & 3;
        for (int q = 0; q < 16; ++q) region[k1b * kRow + (q >> 2) * 17 + (q & 3) * 4 + vb] = regs[t][q];
    }
    // pass C
    for (int t = 0; t < 64; ++t) {
        const int k1b = t >> 2, vb = t & 3;
        for (int qi = 0; qi < 4; ++qi) {
            cpx z[4];
            for (int v = 0; v < 4; ++v) z[v] = region[k1b * kRow + vb * 17 + qi * 4 + v];
            dft4(z[0], z[1], z[2], z[3]);
            for (int r = 0; r < 4; ++r) regs[t][4 * qi + r] = z[r];
        }
    }
    std::vector<cpx> Z(1024 + 40);
    for (int t = 0; t < 64; ++t) {
        const int k1b = t >> 2, vb = t & 3;
        for (int qi = 0; qi < 4; ++qi)
            for (int r = 0; r < 4; ++r) Z[z_unit(k1b + 16 * (4 * vb + qi) + 256 * r)] = regs[t][4 * qi + r];
    }
    // untangle
    std::vector<float> pa(1025), pb(1025);
    for (int t = 0; t < 64; ++t)
        for (int i = 0; i < 9; ++i) {
            const int j = i < 8 ? t + 64 * i : 512;
            if (!(i < 8 || t == 0)) continue;
            const double a = two_pi * j / 2048.0;
            pair_t lo_, hi_;
            untangle_pair_sq(Z[z_unit(j)], Z[z_unit((kC - j) & (kC - 1))], (float)-std::sin(a), (float)-std::cos(a), lo_, hi_);
            pa[j] = 0.25f * lo(lo_); pb[j] = 0.25f * hi(lo_);
            pa[kC - j] = 0.25f * lo(hi_); pb[kC - j] = 0.25f * hi(hi_);
        }
    // reference
    double worst = 0, scale = 0;
    for (int k = 0; k <= 1024; k += 1) {
        double ra = 0, ia = 0, rb = 0, ib = 0;
        for (int n = 0; n < 2048; ++n) {
            const double a = two_pi * ((long long)k * n % 2048) / 2048.0;
            ra += xa[n] * std::cos(a); ia -= xa[n] * std::sin(a);
            rb += xb[n] * std::cos(a); ib -= xb[n] * std::sin(a);
        }
        const double qa = ra * ra + ia * ia, qb = rb * rb + ib * ib;
        worst = std::fmax(worst, std::fmax(std::fabs(qa - pa[k]), std::fabs(qb - pb[k])));
        scale = std::fmax(scale, std::fmax(qa, qb));
    }
    printf("fft2048: max |power error| %.3e of max power %.3e\n", worst, scale);
    if (!(worst <= 2e-5 * scale)) { printf("FAIL\n"); return 1; }
    printf("OK\n");
    return 0;
}
