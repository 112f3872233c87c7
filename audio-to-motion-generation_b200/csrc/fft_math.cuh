// Register-level FFT building blocks of the fused log-mel kernel.
//
// A 512-point real frame x[n] is packed as z[m] = x[2m] + i x[2m+1] (m < 256); Z = FFT256(z) is computed as 16 x 16
// (two register-resident radix-16 passes with one exchange in between), then "untangled" into the real-input
// spectrum X[k], k = 0..256.
//
// Every value is a PAIR: the same quantity of two frames (A, B) that one thread transforms in lock step.  On the
// device the pair is one 64-bit register and the arithmetic is the packed fp32 instructions of sm_100
// (add/sub/mul/fma .f32x2 -> FADD2 / FMUL2 / FFMA2, one issue slot for both frames; constants and per-lane twiddles
// enter as scalar-broadcast operands).  On the host (tests/host/fft_host_test.cu) the pair is a struct of two floats
// with the same operation order, so the index algebra is unit-tested on the CPU before it ever runs on a GPU.
#pragma once
#include <cuda_runtime.h>
#include <cmath>

// product build: device functions on 64-bit register pairs.  A2M_FFT_HOST_EMULATION (the host test): plain C++.
#ifdef A2M_FFT_HOST_EMULATION
#define A2M_HD inline
#else
#define A2M_HD __device__ __forceinline__
#endif

namespace a2m_fft {

#ifndef A2M_FFT_HOST_EMULATION
typedef unsigned long long pair_t;                    // (A, B) = (low, high) fp32 halves
__device__ __forceinline__ pair_t pack(float a, float b) { pair_t d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ pair_t bcast(float c) { return pack(c, c); }                 // becomes a scalar-broadcast operand
__device__ __forceinline__ float lo(pair_t p) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(p)); (void)b; return a; }
__device__ __forceinline__ float hi(pair_t p) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(p)); (void)a; return b; }
__device__ __forceinline__ pair_t add2(pair_t a, pair_t b) { pair_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ pair_t sub2(pair_t a, pair_t b) { pair_t d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ pair_t mul2(pair_t a, pair_t b) { pair_t d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ pair_t fma2(pair_t a, pair_t b, pair_t c) { pair_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
#else
struct pair_t { float a, b; };
inline pair_t pack(float a, float b) { pair_t p; p.a = a; p.b = b; return p; }
inline pair_t bcast(float c) { return pack(c, c); }
inline float lo(pair_t p) { return p.a; }
inline float hi(pair_t p) { return p.b; }
inline pair_t add2(pair_t x, pair_t y) { return pack(x.a + y.a, x.b + y.b); }
inline pair_t sub2(pair_t x, pair_t y) { return pack(x.a - y.a, x.b - y.b); }
inline pair_t mul2(pair_t x, pair_t y) { return pack(x.a * y.a, x.b * y.b); }
inline pair_t fma2(pair_t x, pair_t y, pair_t z) { return pack(fmaf(x.a, y.a, z.a), fmaf(x.b, y.b, z.b)); }
#endif

struct cpx {                 // a complex number of each of the two frames
    pair_t re, im;
};
A2M_HD cpx make(pair_t re, pair_t im) { cpx c; c.re = re; c.im = im; return c; }
A2M_HD cpx add(cpx a, cpx b) { return make(add2(a.re, b.re), add2(a.im, b.im)); }
A2M_HD cpx sub(cpx a, cpx b) { return make(sub2(a.re, b.re), sub2(a.im, b.im)); }
// a * (wr + i wi) with a scalar (both frames share the twiddle): 2 mul + 2 fma; the negated scalar is an operand
// modifier on the device
A2M_HD cpx mul_scalar(cpx a, float wr, float wi) {
    const pair_t R = bcast(wr), I = bcast(wi), NI = bcast(-wi);
    return make(fma2(a.im, NI, mul2(a.re, R)), fma2(a.re, I, mul2(a.im, R)));
}

// 4-point DFT, natural order in and out: y[c] = sum_a x[a] * (-i)^(a c).  No explicit negation anywhere: the factor
// -i of the odd outputs is spelled out in the adds.
A2M_HD void dft4(cpx& x0, cpx& x1, cpx& x2, cpx& x3) {
    const cpx s02 = add(x0, x2), d02 = sub(x0, x2);
    const cpx s13 = add(x1, x3), d13 = sub(x1, x3);
    x0 = add(s02, s13);
    x2 = sub(s02, s13);
    x1 = make(add2(d02.re, d13.im), sub2(d02.im, d13.re));          // d02 + (-i) d13
    x3 = make(sub2(d02.re, d13.im), add2(d02.im, d13.re));          // d02 - (-i) d13
}
// The same with x2 standing for (-i) * x2 (the W16^4 twiddle of the radix-16 kernel, folded in)
A2M_HD void dft4_x2_times_neg_i(cpx& x0, cpx& x1, cpx& x2, cpx& x3) {
    const cpx s02 = make(add2(x0.re, x2.im), sub2(x0.im, x2.re));   // x0 + (-i) x2
    const cpx d02 = make(sub2(x0.re, x2.im), add2(x0.im, x2.re));   // x0 - (-i) x2
    const cpx s13 = add(x1, x3), d13 = sub(x1, x3);
    x0 = add(s02, s13);
    x2 = sub(s02, s13);
    x1 = make(add2(d02.re, d13.im), sub2(d02.im, d13.re));
    x3 = make(sub2(d02.re, d13.im), add2(d02.im, d13.re));
}

#define A2M_C1 0.92387953251128674f   /* cos(pi/8) */
#define A2M_S1 0.38268343236508977f   /* sin(pi/8) */
#define A2M_R2 0.70710678118654752f   /* sqrt(1/2) */
// a * W16^e, W16 = exp(-2 pi i / 16), for the exponents b*c (b, c in 0..3) other than 0 and 4
template <int E>
A2M_HD cpx mul_w16(cpx a) {
    if (E == 1) return mul_scalar(a, A2M_C1, -A2M_S1);
    if (E == 2) return make(mul2(add2(a.re, a.im), bcast(A2M_R2)), mul2(sub2(a.im, a.re), bcast(A2M_R2)));
    if (E == 3) return mul_scalar(a, A2M_S1, -A2M_C1);
    if (E == 6) return make(mul2(sub2(a.im, a.re), bcast(A2M_R2)), mul2(add2(a.re, a.im), bcast(-A2M_R2)));
    if (E == 9) return mul_scalar(a, -A2M_C1, A2M_S1);
    return a;
}

// 16-point DFT in place, natural order in and out (n = 4a + b, k = c + 4d).
A2M_HD void dft16(cpx (&v)[16]) {
    // step 1: for each b, DFT4 over a  -> v[4c + b] holds t[b][c]
    dft4(v[0], v[4], v[8], v[12]);
    dft4(v[1], v[5], v[9], v[13]);
    dft4(v[2], v[6], v[10], v[14]);
    dft4(v[3], v[7], v[11], v[15]);
    // twiddle t[b][c] *= W16^(b c); W16^4 = -i of t[2][2] is folded into the DFT4 that consumes it
    v[5] = mul_w16<1>(v[5]);   v[6] = mul_w16<2>(v[6]);   v[7] = mul_w16<3>(v[7]);
    v[9] = mul_w16<2>(v[9]);                              v[11] = mul_w16<6>(v[11]);
    v[13] = mul_w16<3>(v[13]); v[14] = mul_w16<6>(v[14]); v[15] = mul_w16<9>(v[15]);
    // step 2: for each c, DFT4 over b -> X[c + 4d] lands in v[4c + d]
    dft4(v[0], v[1], v[2], v[3]);
    dft4(v[4], v[5], v[6], v[7]);
    dft4_x2_times_neg_i(v[8], v[9], v[10], v[11]);
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

// Real-input untangle of the bin pair (k, 256 - k), 0 <= k <= 128, from Z[k] and Z[(256 - k) & 255]:
//   A = Z[k] + conj(Zp), B = Z[k] - conj(Zp), t_k = -i exp(-2 pi i k / 512) = (tx, ty) = (-sin th, -cos th)
//   2 X[k] = A + t_k B,   2 conj(X[256 - k]) = A - t_k B
// Returns |2 X[k]|^2 and |2 X[256 - k]|^2 (the factor 1/2 of the magnitude is folded into the mel weights).
A2M_HD void untangle_pair_sq(cpx zk, cpx zp, float tx, float ty, pair_t& sq_k, pair_t& sq_mirror) {
    const pair_t a_re = add2(zk.re, zp.re), a_im = sub2(zk.im, zp.im);
    const pair_t b_re = sub2(zk.re, zp.re), b_im = add2(zk.im, zp.im);
    const pair_t TX = bcast(tx), TY = bcast(ty), NTY = bcast(-ty);
    const pair_t tb_re = fma2(b_im, NTY, mul2(b_re, TX));
    const pair_t tb_im = fma2(b_re, TY, mul2(b_im, TX));
    const pair_t p_re = add2(a_re, tb_re), p_im = add2(a_im, tb_im);
    const pair_t q_re = sub2(a_re, tb_re), q_im = sub2(a_im, tb_im);
    sq_k = fma2(p_im, p_im, mul2(p_re, p_re));
    sq_mirror = fma2(q_im, q_im, mul2(q_re, q_re));
}

}  // namespace a2m_fft
