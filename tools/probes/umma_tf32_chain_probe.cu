// Probe: does a kind::tf32 tcgen05.mma accept, as its tensor-memory A operand, the FP32 accumulator columns an earlier
// kind::f16 tcgen05.mma wrote -- with no thread touching the data in between?  (Graph stack: Z^h = P^h X, then
// OUT += Z^h W_h^T.  An F16 accumulator cannot be chained: it needs F16 inputs and is stored one value per 32-bit column,
// see umma_f16_chain_probe.cu.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -I include -o /tmp/tf32chain tools/probes/umma_tf32_chain_probe.cu && /tmp/tf32chain
// MMA 1: Z[128 x 64] (F32, TMEM columns 64..127) = P[128 x 64] (bf16, K-major SW128) . X[64 x 64] (bf16, MN-major SW128)
// MMA 2: OUT[128 x 64] (F32, columns 0..63) = Z (TMEM, read as TF32) . W^T, W[64 n][64 k] fp32 as two K-major SW128
//        slices [64 n][32 k]; eight K = 8 steps, step j reads Z columns 8 j .. 8 j + 7
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cuda_runtime.h>
#include "../../audio-to-motion-generation_b200/csrc/a2m_common.cuh"

void a2m_set_error(const char*, ...) {}
int a2m_num_sms() { return 148; }
using namespace a2m;

__host__ __device__ inline float pval(int r, int k) { return ((r * 7 + k * 3) % 11) / 16.f; }
__host__ __device__ inline float xval(int k, int f) { return ((k * 5 + f * 2) % 13 - 6) / 8.f; }
__device__ int g_identity;
__host__ __device__ inline float wval_dense(int n, int f) { return ((n * 3 + f) % 7 - 3) / 4.f; }

__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__global__ void probe(float* z_out /*[128][64]*/, float* out /*[128][64]*/) {
    extern __shared__ unsigned char raw[];
    const uint32_t r0 = smem_u32(raw);
    unsigned char* smem = raw + (((r0 + 1023u) & ~1023u) - r0);
    unsigned char* sP = smem;                  // [128][64] bf16
    unsigned char* sX = smem + 16384;          // [64 k][64 f] bf16 (MN-major B operand)
    unsigned char* sW = smem + 16384 + 8192;   // 2 x [64 n][32 k] fp32
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384 + 8192 + 16384);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
    const int tid = threadIdx.x, warp = tid >> 5;
    auto off = [](int r, int c) { return r * 128 + (((c >> 3) ^ (r & 7)) << 4) + (c & 7) * 2; };
    for (int i = tid; i < 128 * 64; i += 128) *reinterpret_cast<__nv_bfloat16*>(sP + off(i / 64, i % 64)) = __float2bfloat16_rn(pval(i / 64, i % 64));
    for (int i = tid; i < 64 * 64; i += 128) *reinterpret_cast<__nv_bfloat16*>(sX + off(i / 64, i % 64)) = __float2bfloat16_rn(xval(i / 64, i % 64));
    for (int i = tid; i < 64 * 64; i += 128) {
        const int n = i / 64, k = i % 64, slice = k >> 5, kk = k & 31;     // 16-byte chunk = 4 floats
        *reinterpret_cast<float*>(sW + slice * 8192 + n * 128 + (((kk >> 2) ^ (n & 7)) << 4) + (kk & 3) * 4) = (g_identity & 1) ? (n == k ? 1.f : 0.f) : wval_dense(n, k);
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
        const uint32_t idesc1 = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | mn;       // D F32, A/B bf16, B MN-major
        const uint32_t idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | mn;                    // D F32, A/B TF32, K-major
        const uint64_t pd = umma_desc_sw128(smem_u32(sP)), xd = umma_desc_sw128(smem_u32(sX)), wd = umma_desc_sw128(smem_u32(sW));
        for (int kk = 0; kk < 4; ++kk) umma_bf16(tmem + 64, pd + ((kk * 32) >> 4), xd + ((kk * 2048) >> 4), idesc1, kk != 0);
        if (g_identity >= 2) {                      // commit + wait between the producer and the consumer MMA
            int e2 = 0;
            umma_commit(&bar[1]);
            mbar_wait(&bar[1], 0, &e2, 2);
            tc_fence_after();
        }
        for (int j = 0; j < 8; ++j)
            umma_tf32_ts(tmem, tmem + 64 + j * 8, wd + (((j >> 2) * 8192 + (j & 3) * 32) >> 4), idesc2, j != 0);
        umma_commit(&bar[0]);
    }
    int err = 0;
    mbar_wait(&bar[0], 0, &err, 1);
    tc_fence_after();
    uint32_t t[32];
    for (int h = 0; h < 2; ++h) {
        tmem_ld_32x32(lane_base + 64 + h * 32, t);
        tmem_ld_wait();
        for (int i = 0; i < 32; ++i) z_out[tid * 64 + h * 32 + i] = __uint_as_float(t[i]);
        tmem_ld_32x32(lane_base + h * 32, t);
        tmem_ld_wait();
        for (int i = 0; i < 32; ++i) out[tid * 64 + h * 32 + i] = __uint_as_float(t[i]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 128); }
}

int main(int argc, char** argv) {
    const int identity = argc > 1 ? atoi(argv[1]) : 0;
    cudaMemcpyToSymbol(g_identity, &identity, sizeof(int));
    float *dz, *dout;
    cudaMalloc(&dz, 128 * 64 * 4); cudaMalloc(&dout, 128 * 64 * 4);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 48000);
    static float hz[128 * 64], hout[128 * 64];
    double zerr = 0, oerr = 0, omax = 0;
    int bad_launches = 0;
    const int n_launches = 200;
    for (int it = 0; it < n_launches; ++it) {
        cudaMemset(dz, 0, 128 * 64 * 4); cudaMemset(dout, 0, 128 * 64 * 4);
        probe<<<1, 128, 48000>>>(dz, dout);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("tf32 chain: %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(hz, dz, sizeof(hz), cudaMemcpyDeviceToHost); cudaMemcpy(hout, dout, sizeof(hout), cudaMemcpyDeviceToHost);
        double this_err = 0;
        for (int r = 0; r < 128; ++r)
            for (int f = 0; f < 64; ++f) {
                float s = 0;
                for (int k = 0; k < 64; ++k) s += pval(r, k) * xval(k, f);
                zerr = fmax(zerr, fabs(s - hz[r * 64 + f]));
            }
        for (int r = 0; r < 128; ++r)
            for (int n = 0; n < 64; ++n) {
                double s = 0;
                for (int f = 0; f < 64; ++f) s += static_cast<double>(hz[r * 64 + f]) * ((identity & 1) ? (n == f ? 1.0 : 0.0) : wval_dense(n, f));
                this_err = fmax(this_err, fabs(s - hout[r * 64 + n])); omax = fmax(omax, fabs(s));
            }
        oerr = fmax(oerr, this_err);
        if (this_err > 0.01) ++bad_launches;
    }
    printf("mode %d: %d of %d launches wrong\n", identity, bad_launches, n_launches);
    printf("  Z = P X (fp32 accumulator): max |err| %.4g\n", zerr);
    printf("  OUT = Z W^T with Z read from tensor memory as TF32: max |err| %.4g (max |value| %.4g; tf32 rounding of Z and W ~ 1e-3 relative)\n", oerr, omax);
    if (identity & 1) {
        for (int r : {0, 5, 37, 100}) {
            printf("  row %3d Z  :", r); for (int f = 0; f < 16; ++f) printf(" %7.3f", hz[r * 64 + f]); printf("\n");
            printf("  row %3d OUT:", r); for (int f = 0; f < 16; ++f) printf(" %7.3f", hout[r * 64 + f]); printf("\n");
        }
    }
    return 0;
}
