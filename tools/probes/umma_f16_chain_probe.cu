// Probe: can a tcgen05.mma write its accumulator as FP16 and a second tcgen05.mma read those same tensor-memory columns
// as its A operand (kind::f16, A format F16), with no thread touching the data in between?  (Graph stack: Z^h = P^h X
// followed by OUT += Z^h W_h^T; today the row threads convert Z from fp32 to packed bf16 through registers.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -I include -o /tmp/f16chain tools/probes/umma_f16_chain_probe.cu && /tmp/f16chain
// MMA 1: Z[128 x 64] (F16, TMEM columns 64..) = P[128 x 64] (bf16, K-major SW128) . X[64 x 64] (bf16, MN-major SW128)
// MMA 2: OUT[128 x 64] (F32, columns 0..63) = Z (TMEM, F16) . W^T, W[64 x 64] K-major SW128, bf16 (mode 0) or fp16 (mode 1)
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "../../audio-to-motion-generation_b200/csrc/a2m_common.cuh"

void a2m_set_error(const char*, ...) {}
int a2m_num_sms() { return 148; }
using namespace a2m;

__host__ __device__ inline float pval(int r, int k) { return ((r * 7 + k * 3) % 11) / 16.f; }
__host__ __device__ inline float xval(int k, int f) { return ((k * 5 + f * 2) % 13 - 6) / 8.f; }
__host__ __device__ inline float wval(int n, int f) { return ((n * 3 + f) % 7 - 3) / 4.f; }

__global__ void probe(int mode, uint32_t* z_raw /*[128][32]*/, float* out /*[128][64]*/) {
    extern __shared__ unsigned char raw[];
    const uint32_t r0 = smem_u32(raw);
    unsigned char* smem = raw + (((r0 + 1023u) & ~1023u) - r0);
    unsigned char* sP = smem;                  // [128][64] bf16
    unsigned char* sX = smem + 16384;          // [64 k][64 f] bf16 (MN-major B operand)
    unsigned char* sW = smem + 16384 + 8192;   // [64 n][64 f] 16-bit
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 32768);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
    const int tid = threadIdx.x, warp = tid >> 5;
    auto off = [](int r, int c) { return r * 128 + (((c >> 3) ^ (r & 7)) << 4) + (c & 7) * 2; };
    const bool in_f16 = mode == 2 || mode == 3 || mode == 5;
    for (int i = tid; i < 128 * 64; i += 128) {
        if (in_f16) *reinterpret_cast<__half*>(sP + off(i / 64, i % 64)) = __float2half_rn(pval(i / 64, i % 64));
        else *reinterpret_cast<__nv_bfloat16*>(sP + off(i / 64, i % 64)) = __float2bfloat16_rn(pval(i / 64, i % 64));
    }
    for (int i = tid; i < 64 * 64; i += 128) {
        if (in_f16) *reinterpret_cast<__half*>(sX + off(i / 64, i % 64)) = __float2half_rn(xval(i / 64, i % 64));
        else *reinterpret_cast<__nv_bfloat16*>(sX + off(i / 64, i % 64)) = __float2bfloat16_rn(xval(i / 64, i % 64));
    }
    for (int i = tid; i < 64 * 64; i += 128) {
        if ((mode & 1) == 0) *reinterpret_cast<__nv_bfloat16*>(sW + off(i / 64, i % 64)) = __float2bfloat16_rn(wval(i / 64, i % 64));
        else *reinterpret_cast<__half*>(sW + off(i / 64, i % 64)) = __float2half_rn(wval(i / 64, i % 64));
    }
    if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_fence_init(); }
    if (warp == 0) { tmem_alloc(slot, 128); tmem_relinquish(); }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot;
    const uint32_t lane_base = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    if (tid == 0) {
        const uint32_t mn = ((64u >> 3) << 17) | ((128u >> 4) << 24);
        // modes 0, 1: P and X bf16 -> Z F16; modes 2, 3: P and X fp16 -> Z F16 (the PTX tables may tie an F16 accumulator to
        // F16 inputs); mode & 1: W fp16 instead of bf16; mode >= 4: MMA 1 alone (mode 4: bf16 inputs, mode 5: fp16 inputs)
        const uint32_t in16 = (mode == 2 || mode == 3 || mode == 5) ? 0u : 1u;
        const uint32_t idesc1 = (0u << 4) | (in16 << 7) | (in16 << 10) | (1u << 16) | mn;       // D F16, B MN-major
        const uint32_t idesc2 = (1u << 4) | (0u << 7) | (((mode & 1) ? 0u : 1u) << 10) | mn;    // D F32, A F16 (TMEM), B bf16 / f16
        const uint64_t pd = umma_desc_sw128(smem_u32(sP)), xd = umma_desc_sw128(smem_u32(sX)), wd = umma_desc_sw128(smem_u32(sW));
        for (int kk = 0; kk < 4; ++kk) umma_bf16(tmem + 64, pd + ((kk * 32) >> 4), xd + ((kk * 2048) >> 4), idesc1, kk != 0);
        if (mode < 4)
            for (int k = 0; k < 4; ++k) umma_bf16_ts(tmem, tmem + 64 + k * 8, wd + ((k * 32) >> 4), idesc2, k != 0);
        umma_commit(&bar[0]);
    }
    int err = 0;
    mbar_wait(&bar[0], 0, &err, 1);
    tc_fence_after();
    uint32_t t[32];
    tmem_ld_32x32(lane_base + 64, t);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) z_raw[tid * 32 + i] = t[i];
    tmem_ld_32x32(lane_base, t);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) out[tid * 64 + i] = __uint_as_float(t[i]);
    tmem_ld_32x32(lane_base + 32, t);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) out[tid * 64 + 32 + i] = __uint_as_float(t[i]);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 128); }
}

int main(int argc, char** argv) {
    const int only = argc > 1 ? atoi(argv[1]) : 0;
    uint32_t* dz; float* dout;
    cudaMalloc(&dz, 128 * 32 * 4); cudaMalloc(&dout, 128 * 64 * 4);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
    for (int mode = only; mode <= only; ++mode) {
        cudaMemset(dz, 0, 128 * 32 * 4); cudaMemset(dout, 0, 128 * 64 * 4);
        probe<<<1, 128, 40000>>>(mode, dz, dout);
        cudaError_t e = cudaDeviceSynchronize();
        printf("mode %d: %s\n", mode, cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
        static uint32_t hz[128 * 32]; static float hout[128 * 64];
        cudaMemcpy(hz, dz, sizeof(hz), cudaMemcpyDeviceToHost); cudaMemcpy(hout, dout, sizeof(hout), cudaMemcpyDeviceToHost);
        // reference
        static float Z[128][64];
        double zerr = 0, oerr = 0, omax = 0; int packed_ok = 0, packed_bad = 0;
        for (int r = 0; r < 128; ++r)
            for (int f = 0; f < 64; ++f) {
                float s = 0;
                for (int k = 0; k < 64; ++k) s += pval(r, k) * xval(k, f);
                Z[r][f] = __half2float(__float2half_rn(s));
            }
        for (int r = 0; r < 128; ++r)
            for (int c = 0; c < 32; ++c) {
                __half lo = __ushort_as_half(static_cast<unsigned short>(hz[r * 32 + c] & 0xffff));
                __half hi = __ushort_as_half(static_cast<unsigned short>(hz[r * 32 + c] >> 16));
                const double d0 = fabs(__half2float(lo) - Z[r][2 * c]), d1 = fabs(__half2float(hi) - Z[r][2 * c + 1]);
                zerr = fmax(zerr, fmax(d0, d1));
                (d0 < 0.02 && d1 < 0.02) ? ++packed_ok : ++packed_bad;
            }
        for (int r = 0; r < 128; ++r)
            for (int n = 0; n < 64; ++n) {
                float s = 0;
                for (int f = 0; f < 64; ++f) s += Z[r][f] * wval(n, f);
                oerr = fmax(oerr, fabs(s - hout[r * 64 + n])); omax = fmax(omax, fabs(s));
            }
        printf("  Z as packed half2 (column c = features 2c, 2c+1): max |err| %.4g, %d ok / %d bad\n", zerr, packed_ok, packed_bad);
        printf("  row 0 raw Z columns 0..3: %08x %08x %08x %08x (expected %.3f %.3f %.3f %.3f ...)\n", hz[0], hz[1], hz[2], hz[3],
               Z[0][0], Z[0][1], Z[0][2], Z[0][3]);
        printf("  OUT = Z W^T: max |err| %.4g (max |value| %.4g)\n", oerr, omax);
    }
    return 0;
}
