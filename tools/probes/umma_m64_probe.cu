// Probe (groundwork for a fused decoder-stack kernel, DESIGN.md "what comes next"): where does tcgen05.mma with M = 64
// (cta_group::1) put its accumulator rows in tensor memory, and does a K-major SWIZZLE_128B A operand whose start
// address is shifted by whole rows (conv taps on a shared-memory-resident activation tile) need the descriptor's
// base-offset field?
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -I include -o /tmp/probe tools/probes/umma_m64_probe.cu && /tmp/probe
// A[r][c] (bf16, [136 rows][64 cols], SW128 K-major, 1024-aligned) = r + 1 for c == 0 else 0; B[n][k] = 1 for n == k == 0.
// D[m][0] then names the A row that fed output row m; every TMEM lane is read back and printed.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../audio-to-motion-generation_b200/csrc/a2m_common.cuh"

void a2m_set_error(const char*, ...) {}
int a2m_num_sms() { return 148; }
using namespace a2m;

__device__ __forceinline__ uint64_t desc_sw128_base(uint32_t smem_addr, uint32_t base_offset) {
    uint64_t d = umma_desc_sw128(smem_addr);
    d |= static_cast<uint64_t>(base_offset & 7) << 49;
    return d;
}

__global__ void probe(int M, int row_shift, int use_base_offset, int d_lane, float* out /*[128]*/) {
    extern __shared__ unsigned char raw[];
    const uint32_t r0 = smem_u32(raw);
    unsigned char* smem = raw + (((r0 + 1023u) & ~1023u) - r0);
    __nv_bfloat16* A = reinterpret_cast<__nv_bfloat16*>(smem);                  // 136 rows x 128 B
    __nv_bfloat16* Bm = reinterpret_cast<__nv_bfloat16*>(smem + 18432);         // 16 rows x 128 B (1024-aligned)
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 18432 + 2048);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 136 * 64; i += 128) {
        const int r = i / 64, c = i % 64;
        const int chunk = c >> 3;                                               // 16-byte chunk, swizzled with the row
        const int off = r * 128 + ((chunk ^ (r & 7)) << 4) + (c & 7) * 2;
        *reinterpret_cast<__nv_bfloat16*>(smem + off) = __float2bfloat16_rn(c == 0 ? static_cast<float>(r + 1) : 0.f);
    }
    for (int i = tid; i < 16 * 64; i += 128) {
        const int r = i / 64, c = i % 64;
        const int off = r * 128 + (((c >> 3) ^ (r & 7)) << 4) + (c & 7) * 2;
        *reinterpret_cast<__nv_bfloat16*>(smem + 18432 + off) = __float2bfloat16_rn((r == 0 && c == 0) ? 1.f : 0.f);
    }
    if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); }
    if (warp == 0) { tmem_alloc(slot, 32); tmem_relinquish(); }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot;
    {   // zero the 32 columns of my lane
        uint32_t z[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        tmem_st_32x16(tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16), z);
        tmem_st_32x16(tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16) + 16, z);
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem) + row_shift * 128;
        const uint32_t bo = use_base_offset ? ((a_addr >> 7) & 7) : 0;
        umma_bf16(tmem + (static_cast<uint32_t>(d_lane) << 16), desc_sw128_base(a_addr, bo), umma_desc_sw128(smem_u32(smem + 18432)), umma_idesc_bf16(M, 16), 0);
        umma_commit(bar);
    }
    mbar_wait(bar, 0, nullptr, 0);
    tc_fence_after();
    uint32_t v[8];
    tmem_ld_32x8(tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16), v);
    tmem_ld_wait();
    out[tid] = __uint_as_float(v[0]);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 32); }
}

int main() {
    float* d;
    cudaMalloc(&d, 128 * sizeof(float));
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 24 * 1024);
    const int cases[][4] = {{128, 0, 0, 0}, {64, 0, 0, 0}, {64, 1, 0, 0}, {64, 1, 1, 0}, {64, 66, 0, 0}, {128, 3, 0, 0}, {64, 66, 0, 16}};
    for (auto& c : cases) {
        cudaMemset(d, 0, 512);
        probe<<<1, 128, 24 * 1024>>>(c[0], c[1], c[2], c[3], d);
        cudaError_t e = cudaDeviceSynchronize();
        float h[128];
        cudaMemcpy(h, d, 512, cudaMemcpyDeviceToHost);
        printf("M=%d row_shift=%d base_offset_field=%s D lane offset %d : %s\n", c[0], c[1], c[2] ? "set" : "0", c[3], cudaGetErrorString(e));
        for (int l = 0; l < 128; ++l) printf("%s%g", l % 32 == 0 ? "\n  lanes " : " ", h[l]);
        printf("\n");
        if (e != cudaSuccess) break;
    }
    return 0;
}
