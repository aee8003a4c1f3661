// micro-benchmark 3: the MMA patterns of the 128 x 128-tile attention kernels issued by one thread of one CTA per SM,
// alone and with N_LD warps streaming tcgen05.ld / tcgen05.st over other TMEM columns (contention for TMEM).
#include <cstdio>
#include <cuda_runtime.h>
#include "../../spt_proto_b200/csrc/tc.cuh"
using namespace spt::tc;
namespace spt { thread_local char g_last_error[512]; std::atomic<uint64_t> g_launch_count{0}; }

// PAT 0: fwd128  : 4 x SS N128 -> S ; 8 x TS N64 -> O                    (single chains)
// PAT 1: kv128   : 4 x (SS N128 -> S, SS N128 -> dP) ; 8 x (TS N64 -> dV, TS N64 -> dK)
// PAT 2: q128    : 4 x (SS -> S, SS -> dP) ; 8 x TS N64 -> dQ
// PAT 3: fwd, interleaved: k-steps of S (into alternate buffer) and of PV alternate
// PAT 4: only 8 x TS N64 -> O (same accumulator)
// PAT 5: only 4 x SS N128 -> S (same accumulator)
template <int PAT>
__global__ void __launch_bounds__(576, 1) k(long long *out, int iters, int n_ld) {
    extern __shared__ unsigned char raw[];
    const uint32_t base = (smem_u32(raw) + 1023) & ~1023u;
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    __shared__ int stop;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_fence_init(); stop = 0; }
    if (warp == 17) tmem_alloc<512>(smem_u32(&slot));
    fence_before_sync(); __syncthreads(); fence_after_sync();
    const uint32_t tm = slot;
    if (threadIdx.x == 17 * 32) {
        constexpr uint32_t id_s = idesc_bf16(128, 128, 0, 0), id_a = idesc_bf16(128, 64, 0, 1);
        const uint64_t da = desc_kmajor(base, 0), db = desc_kmajor(base + 16384, 0), dc = desc_kmajor(base + 32768, 0),
                       dd = desc_kmajor(base + 49152, 0), dbt = desc_mnmajor(base + 16384, 0, 16384), ddt = desc_mnmajor(base + 49152, 0, 16384);
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            if (PAT == 0) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) umma_bf16(tm, da + kk * 2, db + kk * 2, id_s, kk != 0);
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) umma_bf16_ts(tm + 256, tm + 128 + kk * 8, dbt + kk * 128, id_a, 1);
            } else if (PAT == 1) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) { umma_bf16(tm, da + kk * 2, db + kk * 2, id_s, kk != 0); umma_bf16(tm + 128, dc + kk * 2, dd + kk * 2, id_s, kk != 0); }
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) { umma_bf16_ts(tm + 384, tm + 256 + kk * 8, dbt + kk * 128, id_a, 1); umma_bf16_ts(tm + 448, tm + 320 + kk * 8, ddt + kk * 128, id_a, 1); }
            } else if (PAT == 2) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) { umma_bf16(tm, da + kk * 2, db + kk * 2, id_s, kk != 0); umma_bf16(tm + 128, dc + kk * 2, dd + kk * 2, id_s, kk != 0); }
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) umma_bf16_ts(tm + 384, tm + 256 + kk * 8, dbt + kk * 128, id_a, 1);
            } else if (PAT == 3) {
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {
                    if (kk < 4) umma_bf16(tm + (i & 1) * 128, da + kk * 2, db + kk * 2, id_s, kk != 0);
                    umma_bf16_ts(tm + 384, tm + 256 + kk * 8, dbt + kk * 128, id_a, 1);
                }
            } else if (PAT == 4) {
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) umma_bf16_ts(tm + 256, tm + 128 + kk * 8, dbt + kk * 128, id_a, 1);
            } else {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) umma_bf16(tm, da + kk * 2, db + kk * 2, id_s, kk != 0);
            }
        }
        umma_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 0);
        out[blockIdx.x] = clock64() - t0;
        *(volatile int *)&stop = 1;
    } else if (warp < n_ld) {
        // TMEM traffic like the math warps': per "tile" one 32-column load + one 16-column store, then ~1000 clk of ALU work
        const uint32_t lane_base = tm + ((uint32_t)((warp & 3) * 32) << 16);
        uint32_t r[32], acc = 0;
        while (!*(volatile int *)&stop) {
            tmem_ld32(lane_base + (warp >> 2) * 32, r);
#pragma unroll
            for (int i = 0; i < 32; ++i) acc += r[i];
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) pk[i] = acc + i;
            tmem_st16(lane_base + 480 + (warp >> 2) * 4, pk);   // columns 480.. : outside every operand / accumulator
            tmem_st_wait();
            long long t = clock64();
            while (clock64() - t < 600) {}
        }
        if (acc == 0x12345) out[200] = acc;
    }
    fence_before_sync(); __syncthreads();
    if (warp == 17) tmem_dealloc<512>(tm);
}
template <int PAT> void run(const char *name, int n_ld) {
    long long *d; cudaMalloc(&d, 400 * 8);
    const int iters = 500;
    cudaFuncSetAttribute(k<PAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
    for (int rep = 0; rep < 2; ++rep) k<PAT><<<148, 576, 80 * 1024>>>(d, iters, n_ld);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-28s ld warps %2d: %.0f clk per tile  err=%s\n", name, n_ld, (double)h[0] / iters, cudaGetErrorString(cudaGetLastError()));
    cudaFree(d);
}
int main() {
    for (int n_ld : {0, 16}) {
        run<0>("fwd128 (S chain, O chain)", n_ld);
        run<3>("fwd interleaved S/PV", n_ld);
        run<1>("kv128 (2+2 accumulators)", n_ld);
        run<2>("q128", n_ld);
        run<4>("8 x TS N64 same acc", n_ld);
        run<5>("4 x SS N128 same acc", n_ld);
    }
    return 0;
}
