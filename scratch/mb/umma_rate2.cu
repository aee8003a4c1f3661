// micro-benchmark 2: is the ~93-clk cost of a small tcgen05.mma per issuing thread or per SM?  1 vs 2 CTAs per SM.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../spt_proto_b200/csrc/tc.cuh"
using namespace spt::tc;
namespace spt { thread_local char g_last_error[512]; std::atomic<uint64_t> g_launch_count{0}; }
template <int N, int TS>
__global__ void __launch_bounds__(128, 2) k(long long *out, int iters) {
    extern __shared__ unsigned char raw[];
    const uint32_t base = (smem_u32(raw) + 1023) & ~1023u;
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_fence_init(); }
    if (threadIdx.x < 32) tmem_alloc<256>(smem_u32(&slot));
    fence_before_sync(); __syncthreads(); fence_after_sync();
    const uint32_t tm = slot;
    if (threadIdx.x == 0) {
        constexpr uint32_t idesc = idesc_bf16(128, N, 0, 0);
        const uint64_t da = desc_kmajor(base, 0), db = desc_kmajor(base + 16384, 0);
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                if (TS) umma_bf16_ts(tm + (kk & 1) * 64, tm + 192 + kk * 8, db + kk * 2, idesc, 1);
                else umma_bf16(tm + (kk & 1) * 64, da + kk * 2, db + kk * 2, idesc, 1);
            }
        }
        umma_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 0);
        out[blockIdx.x] = clock64() - t0;
    }
    fence_before_sync(); __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc<256>(tm);
}
template <int N, int TS> void run(int ctas) {
    long long *d; cudaMalloc(&d, 296 * 8);
    const int iters = 2000;
    cudaFuncSetAttribute(k<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int rep = 0; rep < 2; ++rep) k<N, TS><<<ctas, 128, 64 * 1024>>>(d, iters);
    cudaDeviceSynchronize();
    long long h[296]; cudaMemcpy(h, d, ctas * 8, cudaMemcpyDeviceToHost);
    printf("N=%3d %s, %3d CTAs: %.1f clk per MMA per CTA  err=%s\n", N, TS ? "A tmem" : "A smem", ctas, (double)h[0] / (iters * 4), cudaGetErrorString(cudaGetLastError()));
}
int main() { run<32, 0>(148); run<32, 0>(296); run<32, 1>(296); run<64, 0>(148); run<64, 0>(296); run<64, 1>(148); run<64, 1>(296); run<128, 0>(148); run<128, 0>(296); return 0; }
