// micro-benchmark: back-to-back tcgen05.mma (cta_group::1, M128 N{64,128,256} K16, bf16) on fixed smem operands
#include <cstdio>
#include <cuda_runtime.h>
#include "../../spt_proto_b200/csrc/tc.cuh"
using namespace spt::tc;
namespace spt { thread_local char g_last_error[512]; std::atomic<uint64_t> g_launch_count{0}; }
template <int N>
__global__ void __launch_bounds__(128, 1) k(long long *out, int iters, int a_tmem, int n_acc) {
    extern __shared__ unsigned char raw[];
    const uint32_t base = (smem_u32(raw) + 1023) & ~1023u;
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_fence_init(); }
    if (threadIdx.x < 32) tmem_alloc<512>(smem_u32(&slot));
    fence_before_sync(); __syncthreads(); fence_after_sync();
    const uint32_t tm = slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = idesc_bf16(128, N, 0, 0);
        const uint64_t da = desc_kmajor(base, 0), db = desc_kmajor(base + 32768, 0);
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const uint32_t acc = tm + ((i * 4 + kk) % n_acc) * N;     // rotate over n_acc independent accumulators
                if (a_tmem) umma_bf16_ts(acc, tm + 480 + (kk & 1) * 8, db + kk * 2, idesc, 1);
                else umma_bf16(acc, da + kk * 2, db + kk * 2, idesc, 1);
            }
        }
        umma_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 0);
        long long t1 = clock64();
        out[blockIdx.x] = t1 - t0;
    }
    fence_before_sync(); __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc<512>(tm);
}
template <int N> void run(int a_tmem, int n_acc) {
    long long *d; cudaMalloc(&d, 148 * 8);
    const int iters = 2000;
    cudaFuncSetAttribute(k<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    for (int rep = 0; rep < 2; ++rep) k<N><<<148, 128, 100 * 1024>>>(d, iters, a_tmem, n_acc);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double clk = (double)h[0] / (iters * 4);
    printf("N=%3d acc=%d A=%s: %.1f clk per MMA (M128 K16) -> %.0f flop/clk/SM  err=%s\n", N, n_acc, a_tmem ? "tmem" : "smem", clk, 2.0 * 128 * N * 16 / clk, cudaGetErrorString(cudaGetLastError()));
}
int main() { for (int na : {1, 2, 4, 8}) { run<32>(0, na); run<64>(0, na); run<128>(0, na < 3 ? na : 3); } run<256>(0, 1); run<32>(1, 1); run<32>(1, 2); run<32>(1, 4); run<64>(1, 1); run<64>(1, 2); run<64>(1, 4); run<128>(1, 2); run<256>(1, 1); return 0; }
