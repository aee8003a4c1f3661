// micro-benchmark 4: round-trip latency of the MMA <-> math-warp handshake used by the attention kernels:
//   issuer: tcgen05.mma (N128) + tcgen05.commit(bar1) -> N_W warps: wait bar1 [, tcgen05.ld] -> elected arrive bar2 -> issuer: wait bar2.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../spt_proto_b200/csrc/tc.cuh"
using namespace spt::tc;
namespace spt { thread_local char g_last_error[512]; std::atomic<uint64_t> g_launch_count{0}; }

template <int MODE>   // 0: wait + arrive only; 1: + tcgen05.ld x32 + wait::ld + fence; 2: like 1 + tcgen05.st x16
__global__ void __launch_bounds__(576, 1) k(long long *out, int iters, int n_w, int n_mma) {
    extern __shared__ unsigned char raw[];
    const uint32_t base = (smem_u32(raw) + 1023) & ~1023u;
    __shared__ uint64_t bars[2];
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar1 = smem_u32(&bars[0]), bar2 = smem_u32(&bars[1]);
    if (threadIdx.x == 0) { mbar_init(bar1, 1); mbar_init(bar2, n_w); mbar_fence_init(); }
    if (warp == 17) tmem_alloc<512>(smem_u32(&slot));
    fence_before_sync(); __syncthreads(); fence_after_sync();
    const uint32_t tm = slot;
    if (warp == 17) {
        constexpr uint32_t id_s = idesc_bf16(128, 128, 0, 0);
        const uint64_t da = desc_kmajor(base, 0), db = desc_kmajor(base + 16384, 0);
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            if (i > 0) mbar_wait(bar2, (i - 1) & 1);
            fence_after_sync();
            if (elect_one()) {
                for (int kk = 0; kk < n_mma; ++kk) umma_bf16(tm, da + kk * 2, db + kk * 2, id_s, kk != 0);
                umma_commit(bar1);
            }
            __syncwarp();
        }
        mbar_wait(bar2, (iters - 1) & 1);
        if (lane == 0) out[blockIdx.x] = clock64() - t0;
    } else if (warp < n_w) {
        const uint32_t lane_base = tm + ((uint32_t)((warp & 3) * 32) << 16);
        uint32_t acc = 0;
        for (int i = 0; i < iters; ++i) {
            mbar_wait(bar1, i & 1);
            if (MODE >= 1) {
                fence_after_sync();
                uint32_t r[32];
                tmem_ld32(lane_base + (warp >> 2) * 32, r);
                acc += r[lane & 31 ? 3 : 5];
                if (MODE >= 2) {
                    uint32_t pk[16];
#pragma unroll
                    for (int u = 0; u < 16; ++u) pk[u] = r[2 * u] & acc;
                    tmem_st16(lane_base + 256 + (warp >> 2) * 16, pk);
                    tmem_st_wait();
                }
                fence_before_sync();
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar2);
        }
        if (acc == 0x12345) out[300] = acc;
    }
    fence_before_sync(); __syncthreads();
    if (warp == 17) tmem_dealloc<512>(tm);
}
template <int MODE> void run(int n_w, int n_mma) {
    long long *d; cudaMalloc(&d, 400 * 8);
    const int iters = 2000;
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
    for (int rep = 0; rep < 2; ++rep) k<MODE><<<148, 576, 80 * 1024>>>(d, iters, n_w, n_mma);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("mode %d, %2d waiting warps, %d MMAs per round: %.0f clk per round trip  err=%s\n", MODE, n_w, n_mma, (double)h[0] / iters, cudaGetErrorString(cudaGetLastError()));
    cudaFree(d);
}
int main() {
    for (int n_w : {1, 4, 16}) { run<0>(n_w, 1); run<0>(n_w, 4); run<1>(n_w, 4); run<2>(n_w, 4); }
    return 0;
}
