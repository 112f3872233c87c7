// Register-level FFT building blocks of the fused log-mel kernel (host + device, so the index
// algebra is unit-tested on the CPU by tests/host/fft_host_test.cpp before it ever runs on a GPU).
//
// A 512-point real frame x[n] is packed as z[m] = x[2m] + i x[2m+1] (m < 256); Z = FFT256(z) is
// computed as 16 x 16 (two register-resident radix-16 passes with one exchange in between), then
// "untangled" into the real-input spectrum X[k], k = 0..256.
#pragma once
#include <cuda_runtime.h>

#ifndef A2M_HD
#define A2M_HD __host__ __device__ __forceinline__
#endif

namespace a2m_fft {

struct cpx {
    float x, y;
};
A2M_HD cpx make(float x, float y) { cpx c; c.x = x; c.y = y; return c; }
A2M_HD cpx add(cpx a, cpx b) { return make(a.x + b.x, a.y + b.y); }
A2M_HD cpx sub(cpx a, cpx b) { return make(a.x - b.x, a.y - b.y); }
A2M_HD cpx mul(cpx a, cpx b) { return make(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
A2M_HD cpx mul_neg_i(cpx a) { return make(a.y, -a.x); }     // a * (-i)

// 4-point DFT, natural order in and out: y[c] = sum_a x[a] * (-i)^(a c)
A2M_HD void dft4(cpx& x0, cpx& x1, cpx& x2, cpx& x3) {
    const cpx s02 = add(x0, x2), d02 = sub(x0, x2);
    const cpx s13 = add(x1, x3), d13 = mul_neg_i(sub(x1, x3));
    x0 = add(s02, s13);
    x1 = add(d02, d13);
    x2 = sub(s02, s13);
    x3 = sub(d02, d13);
}

// W16^e = exp(-2 pi i e / 16) for the exponents b*c, b,c in 0..3
#define A2M_C1 0.92387953251128674f   /* cos(pi/8) */
#define A2M_S1 0.38268343236508977f   /* sin(pi/8) */
#define A2M_R2 0.70710678118654752f   /* sqrt(1/2) */
template <int E>
A2M_HD cpx mul_w16(cpx a) {
    if (E == 0) return a;
    if (E == 1) return mul(a, make(A2M_C1, -A2M_S1));
    if (E == 2) return make((a.x + a.y) * A2M_R2, (a.y - a.x) * A2M_R2);
    if (E == 3) return mul(a, make(A2M_S1, -A2M_C1));
    if (E == 4) return mul_neg_i(a);
    if (E == 6) return make((a.y - a.x) * A2M_R2, -(a.x + a.y) * A2M_R2);
    if (E == 9) return mul(a, make(-A2M_C1, A2M_S1));
    return a;
}

// 16-point DFT in place, natural order in and out (n = 4a + b, k = c + 4d).
A2M_HD void dft16(cpx (&v)[16]) {
    // step 1: for each b, DFT4 over a  -> v[4c + b] holds t[b][c]
    dft4(v[0], v[4], v[8], v[12]);
    dft4(v[1], v[5], v[9], v[13]);
    dft4(v[2], v[6], v[10], v[14]);
    dft4(v[3], v[7], v[11], v[15]);
    // twiddle t[b][c] *= W16^(b c)
    v[5] = mul_w16<1>(v[5]);   v[6] = mul_w16<2>(v[6]);   v[7] = mul_w16<3>(v[7]);
    v[9] = mul_w16<2>(v[9]);   v[10] = mul_w16<4>(v[10]); v[11] = mul_w16<6>(v[11]);
    v[13] = mul_w16<3>(v[13]); v[14] = mul_w16<6>(v[14]); v[15] = mul_w16<9>(v[15]);
    // step 2: for each c, DFT4 over b -> X[c + 4d] lands in v[4c + d]
    dft4(v[0], v[1], v[2], v[3]);
    dft4(v[4], v[5], v[6], v[7]);
    dft4(v[8], v[9], v[10], v[11]);
    dft4(v[12], v[13], v[14], v[15]);
    // v[4c + d] = X[c + 4d]  ->  transpose the 4x4 index to natural order
    cpx t;
    t = v[1];  v[1] = v[4];   v[4] = t;
    t = v[2];  v[2] = v[8];   v[8] = t;
    t = v[3];  v[3] = v[12];  v[12] = t;
    t = v[6];  v[6] = v[9];   v[9] = t;
    t = v[7];  v[7] = v[13];  v[13] = t;
    t = v[11]; v[11] = v[14]; v[14] = t;
}

// Real-input untangle for one bin k (0 <= k < 256):
//   X[k] = 0.5 * ( (Z[k] + P) + t_k * (Z[k] - P) ),  P = conj(Z[(256-k) & 255]),
//   t_k = -i * exp(-2 pi i k / 512) = (-sin th, -cos th), th = 2 pi k / 512.
// Returns 2 * X[k] (the factor 0.5 is folded into the magnitude).
A2M_HD cpx untangle2(cpx zk, cpx zpartner, cpx tk) {
    const cpx p = make(zpartner.x, -zpartner.y);
    return add(add(zk, p), mul(tk, sub(zk, p)));
}
A2M_HD float half_magnitude(cpx twoX) {      // |X| from 2X, IEEE sqrt
    return 0.5f * sqrtf(twoX.x * twoX.x + twoX.y * twoX.y);
}

}  // namespace a2m_fft
