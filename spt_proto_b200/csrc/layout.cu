// Layout copies of the stage path: the reference layer moves q, k, v to head-major with
// `transpose(1, 2).contiguous()` before its kernels and the product back afterwards
// (naive_gpt/layers/sparse/attention.py:92-95, 138-142).  torch's strided copy reads 2-byte elements one by one
// (44 us per 33.5 MB operand = 1.5 TB/s); here a row of E elements moves as 16-byte words, four rows in flight per
// thread, and the last-two-dims transpose of the shipped output layout goes through a padded shared-memory tile.
// (The fused path needs neither: its kernels stride over [N, S, H, E] through 4-D tensor maps.)
#include "common.cuh"

namespace spt {
namespace layout {

// out[a, c, b, :] = in[a, b, c, :], rows of `w16` 16-byte words.  One thread per (output row, word), consecutive threads
// = consecutive words of consecutive output rows (stores fully coalesced, loads in whole rows of >= 64 bytes).
constexpr int SW_UNROLL = 4;
// I = unsigned where the word count fits 32 bits: the three divisions per word are the kernel's instruction budget
// (62 % issue-active with 64-bit indices), 32-bit ones cost a quarter of that
template <typename I>
__device__ __forceinline__ I swap12_src(I o, I Bd, I Cd, I w16) {
    const I row = o / w16, w = o - row * w16;
    const I ac = row / Bd, b = row - ac * Bd;     // output row = (a, c, b)
    const I a = ac / Cd, c = ac - a * Cd;
    return ((a * Bd + b) * Cd + c) * w16 + w;
}
template <typename I>
__global__ void __launch_bounds__(256)
swap12_kernel(const int4 *__restrict__ in, int4 *__restrict__ out, long long n_words_ll, int Bd_, int Cd_, int w16_) {
    const I n_words = (I)n_words_ll, Bd = (I)Bd_, Cd = (I)Cd_, w16 = (I)w16_;
    const I stride = (I)gridDim.x * blockDim.x;
    I i = (I)blockIdx.x * blockDim.x + threadIdx.x;
    // (i + 3 * stride cannot wrap: the launcher keeps n_words + 4 * stride below 2^32 for the 32-bit instance)
    for (; i + (SW_UNROLL - 1) * stride < n_words; i += SW_UNROLL * stride) {
        int4 v[SW_UNROLL];
#pragma unroll
        for (int u = 0; u < SW_UNROLL; ++u) v[u] = ld_stream(in + swap12_src<I>(i + u * stride, Bd, Cd, w16));
#pragma unroll
        for (int u = 0; u < SW_UNROLL; ++u) st_stream(out + (i + u * stride), v[u]);
    }
    for (; i < n_words; i += stride) st_stream(out + i, ld_stream(in + swap12_src<I>(i, Bd, Cd, w16)));
}

// out[b, c, r] = in[b, r, c] for 2-byte (T = uint16_t) or 4-byte elements: 64 x 64 tiles through shared memory
template <typename T>
__global__ void __launch_bounds__(256)
transpose_tile_kernel(const T *__restrict__ in, T *__restrict__ out, int R, int C) {
    __shared__ T tile[64][64 + (sizeof(T) == 2 ? 2 : 1)];
    const long long base = (long long)blockIdx.z * R * C;
    const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;   // 64 x 4
#pragma unroll 4
    for (int j = ty; j < 64; j += 4)
        if (r0 + j < R && c0 + tx < C) tile[j][tx] = in[base + (long long)(r0 + j) * C + c0 + tx];
    __syncthreads();
#pragma unroll 4
    for (int j = ty; j < 64; j += 4)
        if (c0 + j < C && r0 + tx < R) out[base + (long long)(c0 + j) * R + r0 + tx] = tile[tx][j];
}

}  // namespace layout
}  // namespace spt

using namespace spt;

extern "C" int spt_swap_dims12(const void *in, void *out, int64_t A, int64_t B, int64_t C, int64_t row_bytes,
                               spt_stream_t stream) {
    SPT_REQUIRE(in && out, "swap_dims12: null pointer");
    SPT_REQUIRE(A >= 1 && B >= 1 && C >= 1 && B < (1ll << 31) && C < (1ll << 31), "swap_dims12: bad sizes");
    SPT_REQUIRE(row_bytes >= 16 && row_bytes % 16 == 0 && row_bytes / 16 < (1 << 20),
                "swap_dims12: rows must be a multiple of 16 bytes (got %lld)", (long long)row_bytes);
    SPT_REQUIRE(((uintptr_t)in | (uintptr_t)out) % 16 == 0, "swap_dims12: operands must be 16-byte aligned");
    const int w16 = (int)(row_bytes / 16);
    const long long n_words = (long long)A * B * C * w16;
    long long blocks = (n_words + 256ll * layout::SW_UNROLL - 1) / (256ll * layout::SW_UNROLL);
    const long long cap = (long long)num_sms() * 32;
    if (blocks > cap) blocks = cap;
    if (n_words + 4 * blocks * 256 < (1ll << 32))
        layout::swap12_kernel<unsigned><<<(unsigned)blocks, 256, 0, as_stream(stream)>>>((const int4 *)in, (int4 *)out, n_words,
                                                                                      (int)B, (int)C, w16);
    else
        layout::swap12_kernel<long long><<<(unsigned)blocks, 256, 0, as_stream(stream)>>>((const int4 *)in, (int4 *)out, n_words,
                                                                                       (int)B, (int)C, w16);
    return after_launch("swap12_kernel");
}

extern "C" int spt_transpose_last2(const void *in, void *out, int64_t batch, int R, int C, int elem_bytes,
                                   spt_stream_t stream) {
    SPT_REQUIRE(in && out, "transpose_last2: null pointer");
    SPT_REQUIRE(batch >= 1 && batch <= 65535 && R >= 1 && C >= 1, "transpose_last2: bad sizes");
    SPT_REQUIRE((R + 63) / 64 <= 65535, "transpose_last2: too many rows");
    const dim3 grid((C + 63) / 64, (R + 63) / 64, (unsigned)batch);
    if (elem_bytes == 2)
        layout::transpose_tile_kernel<uint16_t><<<grid, 256, 0, as_stream(stream)>>>((const uint16_t *)in, (uint16_t *)out, R, C);
    else if (elem_bytes == 4)
        layout::transpose_tile_kernel<uint32_t><<<grid, 256, 0, as_stream(stream)>>>((const uint32_t *)in, (uint32_t *)out, R, C);
    else
        return fail(SPT_ERR_INVALID_ARGUMENT, "transpose_last2: element size %d not supported (2 or 4)", elem_bytes);
    return after_launch("transpose_tile_kernel");
}
