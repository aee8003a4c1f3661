// micro-benchmark: cost of a tcgen05.mma (M128 x N64 x K16, bf16) by the layout of the shared-memory A operand:
// K-major vs MN-major (a_major = 1, two 64-wide M chunks 16 KB apart), B MN-major in both.  One CTA per SM.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../spt_proto_b200/csrc/tc.cuh"
using namespace spt::tc;
namespace spt { thread_local char g_last_error[512]; std::atomic<uint64_t> g_launch_count{0}; }
template <int AMN, int BMN>
__global__ void __launch_bounds__(128, 1) k(long long *out, int iters) {
    extern __shared__ unsigned char raw[];
    const uint32_t base = (smem_u32(raw) + 1023) & ~1023u;
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_fence_init(); }
    if (threadIdx.x < 32) tmem_alloc<256>(smem_u32(&slot));
    fence_before_sync(); __syncthreads(); fence_after_sync();
    const uint32_t tm = slot;
    if (threadIdx.x == 0) {
        constexpr uint32_t idesc = idesc_bf16(128, 64, AMN, BMN);
        const uint64_t da = AMN ? desc_mnmajor(base, 0, 16384) : desc_kmajor(base, 0);
        const uint64_t db = BMN ? desc_mnmajor(base + 32768, 0, 16384) : desc_kmajor(base + 32768, 0);
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
                umma_bf16(tm, da + (AMN ? kk * MNMAJOR_K16 : (uint64_t)((kk >> 2) * 1024 + (kk & 3) * 2)),
                          db + (BMN ? kk * MNMAJOR_K16 : (uint64_t)((kk & 3) * 2)), idesc, 1);
        }
        umma_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 0);
        out[blockIdx.x] = clock64() - t0;
    }
    fence_before_sync(); __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc<256>(tm);
}
template <int AMN, int BMN> void run() {
    long long *d; cudaMalloc(&d, 148 * 8);
    const int iters = 1000;
    cudaFuncSetAttribute(k<AMN, BMN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    for (int rep = 0; rep < 2; ++rep) k<AMN, BMN><<<148, 128, 96 * 1024>>>(d, iters);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, d, 148 * 8, cudaMemcpyDeviceToHost);
    printf("A %s, B %s: %.1f clk per M128xN64xK16 MMA  err=%s\n", AMN ? "MN-major" : "K-major ", BMN ? "MN-major" : "K-major ",
           (double)h[0] / (iters * 8), cudaGetErrorString(cudaGetLastError()));
}
int main() { run<0, 0>(); run<0, 1>(); run<1, 1>(); run<1, 0>(); return 0; }
