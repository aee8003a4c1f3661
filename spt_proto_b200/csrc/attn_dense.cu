// Fused PQ-sparse attention, forward + backward, as MASKED DENSE TILES on the tensor cores.
//
// What it replaces: the stage chain  sddmm -> clamp_(scaling * ., -10, 10) -> causal CSR softmax ->
// spmm  of SparseVanillaAttentionV2._get_attn/_apply_attn (reference naive_gpt/layers/sparse/
// attention.py:122-141) and its autograd backward (kernels/spmm.py:23-49, softmax.py:21-30,
// sddmm.py:25-51), including the transposed products dK = dS^T Q, dV = P^T dO.
//
// Why dense tiles: the gathered formulation moves one 128-byte K/V row per selected (query, key)
// pair through L1/shared memory — ~1 SM-cycle per pair and per gather, six gathers per fwd+bwd —
// and needs a CSR->CSC transpose for dK/dV.  On B200 the tensor pipe makes the *dense causal* tile
// product cheaper than that gather even though only 1/4 of the causal entries are selected
// (hard part 4 "decide by measurement": stage path measured at 4.1 ms per
// 2048-token sequence, see profiles/).  The selection enters as a per-row bitmask in the lookup
// kernel's native "lane-major" layout (S % 128 == 0):
//     mask[b][r][4 g + t] bit i  <=>  key 128 g + 4 i + t is one of row r's lookup candidates
// plus extra0[b][r] = number of zero-padding slots of the row (they all alias key 0 and, like in the
// reference, take part in the softmax).  Both are emitted by the lookup kernel directly, so the
// int32 index tensor (64 MB per sequence) is never materialised on this path.
//
//   w[r][j] = mask bit (+ extra0[r] for j == 0)
//   e[r][j] = w * exp(clamp(scale * q_r.k_j, -10, 10)),  Z_r = max(1e-9, sum_j e),  y_r = sum_j e/Z v_j
// No running max is needed (the clamp bounds the exponent), so there is no rescaling pass.
// Backward recomputes e from q,k (flash style), only Z [B,S] is saved:
//   dP = dO V^T,  D_r = dO_r . y_r,  dS = (e/Z) * (dP - D) * [|scale s| <= 10] * scale
//   dV = P^T dO, dK = dS^T Q  (kernel "kv": one CTA per 64-key tile, computes S^T, dP^T)
//   dQ = dS K                 (kernel "q" : one CTA per 64-row tile)
// Deterministic: no atomics anywhere.
//
// Tensor-core path: bf16 mma.sync.m16n8k16 with fp32 accumulation, operands staged in XOR-swizzled
// shared memory by cp.async and fetched with ldmatrix (DESIGN.md section 5 explains why this
// kernel family is not on tcgen05 yet).  bf16 only, head dim 64.
#include "common.cuh"

namespace spt {
namespace attn {

constexpr int D = 64;        // head dim
constexpr int BM = 64;       // tile rows (queries for fwd / dq, keys for dkv)
constexpr int BN = 64;       // tile cols
constexpr int NWARP = 4;     // 16 tile-rows per warp
constexpr int THREADS = NWARP * 32;
constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&h);
}
__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// A [64][64] bf16 tile in shared memory: 128-byte rows, 16-byte chunk c of row r stored at c ^ (r & 7).
__device__ __forceinline__ uint32_t tile_addr(uint32_t base, int row, int chunk) {
    return base + row * (D * 2) + ((chunk ^ (row & 7)) << 4);
}

// cp.async a [64][64] tile (rows row0.. of a [S][64] matrix); rows >= S are zero-filled by the caller's
// guarantee S % 64 == 0, so no predicate is needed.
__device__ __forceinline__ void load_tile_async(uint32_t s_base, const __nv_bfloat16 *g, int row0, int rs) {
    // 64 rows * 8 chunks = 512 chunks, 128 threads -> 4 each
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int idx = threadIdx.x + i * THREADS;
        const int r = idx >> 3, c = idx & 7;
        cp_async16(tile_addr(s_base, r, c), g + (size_t)(row0 + r) * rs + c * 8);
    }
}

// A-operand fragments of a 16-row slab (rows row0..row0+15) for all 4 k-steps of D = 64.
__device__ __forceinline__ void load_a_frags(uint32_t (&a)[4][4], uint32_t s_base, int row0, int lane) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
        ldsm_x4(a[ks], tile_addr(s_base, row0 + (lane & 15), ks * 2 + (lane >> 4)));
}

// acc[n][.] (16 x 64, n = 8 column tiles) = A(16 x 64 over d) * T^T where T is a [64 cols][64 d] tile.
__device__ __forceinline__ void gemm_nt(float (&acc)[8][4], const uint32_t (&a)[4][4], uint32_t s_tile, int lane) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
        for (int np = 0; np < 4; ++np) {  // pairs of n-tiles
            uint32_t b[4];
            ldsm_x4(b, tile_addr(s_tile, np * 16 + (lane & 7) + ((lane >> 4) << 3), ks * 2 + ((lane >> 3) & 1)));
            mma_bf16(acc[2 * np], a[ks], b[0], b[1]);
            mma_bf16(acc[2 * np + 1], a[ks], b[2], b[3]);
        }
    }
}

// acc[n][.] (16 x 64 over d) += P(16 x 64 over the tile's rows, as packed A frags) * T, T = [64 rows][64 d].
__device__ __forceinline__ void gemm_nn(float (&acc)[8][4], const uint32_t (&p)[4][4], uint32_t s_tile, int lane) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {      // 16 tile-rows per k-step
#pragma unroll
        for (int np = 0; np < 4; ++np) {  // pairs of d n-tiles
            uint32_t b[4];
            ldsm_x4_t(b, tile_addr(s_tile, ks * 16 + (lane & 7) + (((lane >> 3) & 1) << 3), np * 2 + (lane >> 4)));
            mma_bf16(acc[2 * np], p[ks], b[0], b[1]);
            mma_bf16(acc[2 * np + 1], p[ks], b[2], b[3]);
        }
    }
}

// fp32 accumulator tile (16 x 64) -> bf16 A-operand fragments for the next GEMM
__device__ __forceinline__ void acc_to_a(uint32_t (&p)[4][4], const float (&acc)[8][4]) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        p[ks][0] = pack_bf16(acc[2 * ks][0], acc[2 * ks][1]);
        p[ks][1] = pack_bf16(acc[2 * ks][2], acc[2 * ks][3]);
        p[ks][2] = pack_bf16(acc[2 * ks + 1][0], acc[2 * ks + 1][1]);
        p[ks][3] = pack_bf16(acc[2 * ks + 1][2], acc[2 * ks + 1][3]);
    }
}

__device__ __forceinline__ void zero_acc(float (&acc)[8][4]) {
#pragma unroll
    for (int n = 0; n < 8; ++n)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[n][i] = 0.0f;
}

// Store a warp's 16 x 64 fp32 accumulator slab (scaled) as bf16 rows of a [S][64] matrix, through the
// warp's own 16-row slab of a shared tile so that global stores are 16-byte coalesced.
__device__ __forceinline__ void store_slab_bf16(const float (&acc)[8][4], float s_lo, float s_hi, uint32_t s_base,
                                                unsigned char *s_ptr, int slab_row0, __nv_bfloat16 *g, int g_row0,
                                                int lane, int rs) {
    const int g4 = lane >> 2, t4 = lane & 3;
#pragma unroll
    for (int n = 0; n < 8; ++n) {
        // element (row g4 / g4+8, cols 8n + 2 t4, +1): chunk n, byte offset 4 * t4 inside the chunk
        const int r0 = slab_row0 + g4, r1 = r0 + 8;
        *reinterpret_cast<uint32_t *>(s_ptr + r0 * (D * 2) + ((n ^ (r0 & 7)) << 4) + t4 * 4) =
            pack_bf16(acc[n][0] * s_lo, acc[n][1] * s_lo);
        *reinterpret_cast<uint32_t *>(s_ptr + r1 * (D * 2) + ((n ^ (r1 & 7)) << 4) + t4 * 4) =
            pack_bf16(acc[n][2] * s_hi, acc[n][3] * s_hi);
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {  // 16 rows * 8 chunks = 128 chunks, 32 lanes -> 4 each
        const int idx = lane + i * 32;
        const int r = slab_row0 + (idx >> 3), c = idx & 7;
        const uint4 v = *reinterpret_cast<const uint4 *>(s_ptr + r * (D * 2) + ((c ^ (r & 7)) << 4));
        *reinterpret_cast<uint4 *>(g + (size_t)(g_row0 + (idx >> 3)) * rs + c * 8) = v;
    }
    (void)s_base;
}

// ---------------------------------------------------------------------------------------------------
// Forward: one CTA per (64-query tile, head); warps own 16 query rows; loop over causal key tiles.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(THREADS)
attn_fwd_kernel(const __nv_bfloat16 *__restrict__ q, const __nv_bfloat16 *__restrict__ k,
                const __nv_bfloat16 *__restrict__ v, const uint32_t *__restrict__ mask,
                const int32_t *__restrict__ extra0, __nv_bfloat16 *__restrict__ y, float *__restrict__ zsum, int S, int H,
                float scale_log2, float clamp_log2) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char *s_q = smem;                       // 8 KB
    unsigned char *s_k = smem + 8192;                // 2 x 8 KB
    unsigned char *s_v = smem + 8192 * 3;            // 2 x 8 KB
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g4 = lane >> 2, t4 = lane & 3;
    const int tile = gridDim.x - 1 - blockIdx.x;     // heaviest tiles first
    const int b = blockIdx.y;
    const int m0 = tile * BM;
    const size_t head = (size_t)b * S;
    const int rs = H * D;                              // row stride: [N, S, H, D] interleaved heads (H = 1: [B, S, D])
    const size_t hoff = ((size_t)(b / H) * S * H + (b % H)) * D;
    const __nv_bfloat16 *qh = q + hoff, *kh = k + hoff, *vh = v + hoff;
    const int words = S / 32;
    const int n_tiles = tile + 1;                    // key tiles 0..tile (causal)

    load_tile_async(smem_u32(s_q), qh, m0, rs);
    load_tile_async(smem_u32(s_k), kh, 0, rs);
    load_tile_async(smem_u32(s_v), vh, 0, rs);
    cp_async_commit();

    const int row_lo = m0 + warp * 16 + g4, row_hi = row_lo + 8;
    const uint32_t *mrow_lo = mask + (head + row_lo) * words, *mrow_hi = mask + (head + row_hi) * words;
    const float ex_lo = (float)extra0[head + row_lo], ex_hi = (float)extra0[head + row_hi];

    float o[8][4];
    zero_acc(o);
    float sum_lo = 0.0f, sum_hi = 0.0f;
    uint32_t aq[4][4];

    for (int jt = 0; jt < n_tiles; ++jt) {
        const int buf = jt & 1;
        if (jt + 1 < n_tiles) {
            load_tile_async(smem_u32(s_k + (buf ^ 1) * 8192), kh, (jt + 1) * BN, rs);
            load_tile_async(smem_u32(s_v + (buf ^ 1) * 8192), vh, (jt + 1) * BN, rs);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        if (jt == 0) load_a_frags(aq, smem_u32(s_q), warp * 16, lane);

        float s[8][4];
        zero_acc(s);
        gemm_nt(s, aq, smem_u32(s_k + buf * 8192), lane);

        // tile column c = 8 n + 2 t4 + i is key 64 jt + c: lane t = 2 (t4 & 1) + i, bit 16 (jt & 1) + 2 n + (t4 >> 1)
        const int wbase = (jt >> 1) * 4 + 2 * (t4 & 1), bbase = 16 * (jt & 1) + (t4 >> 1);
        const uint32_t w_lo0 = __ldg(mrow_lo + wbase), w_lo1 = __ldg(mrow_lo + wbase + 1);
        const uint32_t w_hi0 = __ldg(mrow_hi + wbase), w_hi1 = __ldg(mrow_hi + wbase + 1);
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            const int bit = bbase + 2 * n;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const uint32_t wl = i ? w_lo1 : w_lo0, wh = i ? w_hi1 : w_hi0;
                float wgt_lo = (float)((wl >> bit) & 1u), wgt_hi = (float)((wh >> bit) & 1u);
                if (jt == 0 && n == 0 && t4 == 0 && i == 0) { wgt_lo += ex_lo; wgt_hi += ex_hi; }  // key 0
                const float e_lo = wgt_lo * ex2(fminf(fmaxf(s[n][i] * scale_log2, -clamp_log2), clamp_log2));
                const float e_hi = wgt_hi * ex2(fminf(fmaxf(s[n][2 + i] * scale_log2, -clamp_log2), clamp_log2));
                s[n][i] = e_lo;
                s[n][2 + i] = e_hi;
                sum_lo += e_lo;
                sum_hi += e_hi;
            }
        }
        uint32_t p[4][4];
        acc_to_a(p, s);
        gemm_nn(o, p, smem_u32(s_v + buf * 8192), lane);
        __syncthreads();  // everyone is done with buffer `buf` before it is refilled
    }
    sum_lo += __shfl_xor_sync(FULL, sum_lo, 1);
    sum_lo += __shfl_xor_sync(FULL, sum_lo, 2);
    sum_hi += __shfl_xor_sync(FULL, sum_hi, 1);
    sum_hi += __shfl_xor_sync(FULL, sum_hi, 2);
    sum_lo = fmaxf(sum_lo, 1e-9f);
    sum_hi = fmaxf(sum_hi, 1e-9f);
    if (t4 == 0) {
        zsum[head + row_lo] = sum_lo;
        zsum[head + row_hi] = sum_hi;
    }
    // s_q is free (fragments are in registers): reuse it as the store staging tile
    store_slab_bf16(o, 1.0f / sum_lo, 1.0f / sum_hi, smem_u32(s_q), s_q, warp * 16, y + hoff, m0 + warp * 16, lane, rs);
}

// D_r = dO_r . y_r  (one warp per row pair; bf16 inputs, fp32 out)
__global__ void __launch_bounds__(256)
attn_delta_kernel(const __nv_bfloat16 *__restrict__ dy, const __nv_bfloat16 *__restrict__ y,
                  float *__restrict__ delta, int64_t rows, int S, int H) {
    const int64_t row = (int64_t)blockIdx.x * 32 + (threadIdx.x >> 3);  // 8 lanes per row (8 x 16 B = 128 B)
    if (row >= rows) return;
    const int sub = threadIdx.x & 7;
    const int64_t b = row / S, r = row % S;                              // delta is head-major [B, S]
    const int64_t off = (((b / H) * S + r) * H + (b % H)) * D;
    float a[8], c[8];
    Vec16<__nv_bfloat16>::load(dy + off + sub * 8, a);
    Vec16<__nv_bfloat16>::load(y + off + sub * 8, c);
    float acc = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc = fmaf(a[i], c[i], acc);
    acc = group_sum<8>(acc);
    if (sub == 0) delta[row] = acc;
}

// Shared epilogue math of the two backward kernels for one accumulator element.
//   s_raw: q.k   dp: dO.v   returns (p, ds) with p = w*exp(clamp(scale s))/Z, ds = p*(dp-delta)*inside
__device__ __forceinline__ void bwd_elem(float s_raw, float dp, float wgt, float inv_z, float delta,
                                         float scale_log2, float clamp_log2, float &p, float &ds) {
    const float t = s_raw * scale_log2;
    const float e = ex2(fminf(fmaxf(t, -clamp_log2), clamp_log2));
    p = wgt * e * inv_z;
    const float inside = (t >= -clamp_log2 && t <= clamp_log2) ? 1.0f : 0.0f;
    ds = p * (dp - delta) * inside;
}

// ---------------------------------------------------------------------------------------------------
// Backward, dK / dV: one CTA per (64-key tile, head); warps own 16 keys; loop over query tiles >= tile.
// Works on the transposed tiles S^T = K Q^T and dP^T = V dO^T so that P^T and dS^T come out of the
// accumulators directly as A operands of  dV += P^T dO  and  dK += dS^T Q.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(THREADS)
attn_bwd_kv_kernel(const __nv_bfloat16 *__restrict__ q, const __nv_bfloat16 *__restrict__ k,
                   const __nv_bfloat16 *__restrict__ v, const __nv_bfloat16 *__restrict__ dy,
                   const uint32_t *__restrict__ mask, const int32_t *__restrict__ extra0,
                   const float *__restrict__ zsum, const float *__restrict__ delta, __nv_bfloat16 *__restrict__ dk,
                   __nv_bfloat16 *__restrict__ dv, int S, int H, float scale, float scale_log2, float clamp_log2) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char *s_k = smem;                  // 8 KB (also store staging)
    unsigned char *s_v = smem + 8192;           // 8 KB (also store staging)
    unsigned char *s_q = smem + 8192 * 2;       // 2 x 8 KB
    unsigned char *s_dy = smem + 8192 * 4;      // 2 x 8 KB
    float *s_invz = reinterpret_cast<float *>(smem + 8192 * 6);       // 2 x 64
    float *s_delta = s_invz + 2 * BM;                                  // 2 x 64
    float *s_ex0 = s_delta + 2 * BM;                                   // 2 x 64
    uint32_t *s_mask = reinterpret_cast<uint32_t *>(s_ex0 + 2 * BM);   // 2 x 64 x 4 words (lane t = 0..3)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g4 = lane >> 2, t4 = lane & 3;
    const int jt = blockIdx.x;                   // key tile (early key tiles are the heaviest)
    const int b = blockIdx.y;
    const int n0 = jt * BN;
    const size_t head = (size_t)b * S;
    const int rs = H * D;
    const size_t hoff = ((size_t)(b / H) * S * H + (b % H)) * D;
    const __nv_bfloat16 *qh = q + hoff, *kh = k + hoff, *vh = v + hoff, *dyh = dy + hoff;
    const int words = S / 32;
    const int n_q_tiles = S / BM;

    auto load_row_tile = [&](int it, int buf) {
        load_tile_async(smem_u32(s_q + buf * 8192), qh, it * BM, rs);
        load_tile_async(smem_u32(s_dy + buf * 8192), dyh, it * BM, rs);
        if (threadIdx.x < BM) {
            const size_t r = head + it * BM + threadIdx.x;
            s_invz[buf * BM + threadIdx.x] = 1.0f / zsum[r];
            s_delta[buf * BM + threadIdx.x] = delta[r];
            s_ex0[buf * BM + threadIdx.x] = (float)extra0[r];
            *reinterpret_cast<uint4 *>(s_mask + (buf * BM + threadIdx.x) * 4) =
                *reinterpret_cast<const uint4 *>(mask + r * words + (jt >> 1) * 4);
        }
    };

    load_tile_async(smem_u32(s_k), kh, n0, rs);
    load_tile_async(smem_u32(s_v), vh, n0, rs);
    load_row_tile(jt, 0);
    cp_async_commit();

    float acc_dk[8][4], acc_dv[8][4];
    zero_acc(acc_dk);
    zero_acc(acc_dv);
    uint32_t ak[4][4], av[4][4];
    // this thread's two keys inside the 64-key tile and their bit positions in the 2 mask words
    const int key_lo = warp * 16 + g4, key_hi = key_lo + 8;

    for (int it = jt; it < n_q_tiles; ++it) {
        const int buf = (it - jt) & 1;
        if (it + 1 < n_q_tiles) {
            load_row_tile(it + 1, buf ^ 1);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        if (it == jt) {
            load_a_frags(ak, smem_u32(s_k), warp * 16, lane);
            load_a_frags(av, smem_u32(s_v), warp * 16, lane);
        }
        float st[8][4], dpt[8][4];
        zero_acc(st);
        zero_acc(dpt);
        gemm_nt(st, ak, smem_u32(s_q + buf * 8192), lane);     // S^T  [16 keys x 64 rows]
        gemm_nt(dpt, av, smem_u32(s_dy + buf * 8192), lane);   // dP^T [16 keys x 64 rows]
        const float *invz = s_invz + buf * BM, *dl = s_delta + buf * BM, *ex0 = s_ex0 + buf * BM;
        const uint32_t *mk = s_mask + buf * BM * 4;
        // key (in tile) -> lane t = key & 3 (same for key_lo and key_hi = key_lo + 8), bit 16 (jt & 1) + (key >> 2)
        const int kt = key_lo & 3, kb_lo = 16 * (jt & 1) + (key_lo >> 2), kb_hi = kb_lo + 2;
#pragma unroll
        for (int n = 0; n < 8; ++n) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int r = n * 8 + t4 * 2 + i;           // query row inside the tile
                const uint32_t wd = mk[r * 4 + kt];
                float wgt_lo = (float)((wd >> kb_lo) & 1u), wgt_hi = (float)((wd >> kb_hi) & 1u);
                if (jt == 0 && key_lo == 0) wgt_lo += ex0[r];  // key 0 carries the zero-padding multiplicity
                float p_lo, ds_lo, p_hi, ds_hi;
                bwd_elem(st[n][i], dpt[n][i], wgt_lo, invz[r], dl[r], scale_log2, clamp_log2, p_lo, ds_lo);
                bwd_elem(st[n][2 + i], dpt[n][2 + i], wgt_hi, invz[r], dl[r], scale_log2, clamp_log2, p_hi, ds_hi);
                st[n][i] = p_lo;
                st[n][2 + i] = p_hi;
                dpt[n][i] = ds_lo;
                dpt[n][2 + i] = ds_hi;
            }
        }
        uint32_t pa[4][4];
        acc_to_a(pa, st);
        gemm_nn(acc_dv, pa, smem_u32(s_dy + buf * 8192), lane);  // dV += P^T dO
        acc_to_a(pa, dpt);
        gemm_nn(acc_dk, pa, smem_u32(s_q + buf * 8192), lane);   // dK += dS^T Q
        __syncthreads();
    }
    store_slab_bf16(acc_dv, 1.0f, 1.0f, smem_u32(s_v), s_v, warp * 16, dv + hoff, n0 + warp * 16, lane, rs);
    store_slab_bf16(acc_dk, scale, scale, smem_u32(s_k), s_k, warp * 16, dk + hoff, n0 + warp * 16, lane, rs);
}

// ---------------------------------------------------------------------------------------------------
// Backward, dQ: one CTA per (64-query tile, head), same tiling as the forward.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(THREADS)
attn_bwd_q_kernel(const __nv_bfloat16 *__restrict__ q, const __nv_bfloat16 *__restrict__ k,
                  const __nv_bfloat16 *__restrict__ v, const __nv_bfloat16 *__restrict__ dy,
                  const uint32_t *__restrict__ mask, const int32_t *__restrict__ extra0,
                  const float *__restrict__ zsum, const float *__restrict__ delta, __nv_bfloat16 *__restrict__ dq,
                  int S, int H, float scale, float scale_log2, float clamp_log2) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char *s_q = smem;               // 8 KB (also store staging)
    unsigned char *s_dy = smem + 8192;       // 8 KB
    unsigned char *s_k = smem + 8192 * 2;    // 2 x 8 KB
    unsigned char *s_v = smem + 8192 * 4;    // 2 x 8 KB
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g4 = lane >> 2, t4 = lane & 3;
    const int tile = gridDim.x - 1 - blockIdx.x;
    const int b = blockIdx.y;
    const int m0 = tile * BM;
    const size_t head = (size_t)b * S;
    const int rs = H * D;
    const size_t hoff = ((size_t)(b / H) * S * H + (b % H)) * D;
    const __nv_bfloat16 *qh = q + hoff, *kh = k + hoff, *vh = v + hoff, *dyh = dy + hoff;
    const int words = S / 32;
    const int n_tiles = tile + 1;

    load_tile_async(smem_u32(s_q), qh, m0, rs);
    load_tile_async(smem_u32(s_dy), dyh, m0, rs);
    load_tile_async(smem_u32(s_k), kh, 0, rs);
    load_tile_async(smem_u32(s_v), vh, 0, rs);
    cp_async_commit();

    const int row_lo = m0 + warp * 16 + g4, row_hi = row_lo + 8;
    const uint32_t *mrow_lo = mask + (head + row_lo) * words, *mrow_hi = mask + (head + row_hi) * words;
    const float ex_lo = (float)extra0[head + row_lo], ex_hi = (float)extra0[head + row_hi];
    const float iz_lo = 1.0f / zsum[head + row_lo], iz_hi = 1.0f / zsum[head + row_hi];
    const float dl_lo = delta[head + row_lo], dl_hi = delta[head + row_hi];

    float acc[8][4];
    zero_acc(acc);
    uint32_t aq[4][4], ady[4][4];
    for (int jt = 0; jt < n_tiles; ++jt) {
        const int buf = jt & 1;
        if (jt + 1 < n_tiles) {
            load_tile_async(smem_u32(s_k + (buf ^ 1) * 8192), kh, (jt + 1) * BN, rs);
            load_tile_async(smem_u32(s_v + (buf ^ 1) * 8192), vh, (jt + 1) * BN, rs);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        if (jt == 0) {
            load_a_frags(aq, smem_u32(s_q), warp * 16, lane);
            load_a_frags(ady, smem_u32(s_dy), warp * 16, lane);
        }
        float s[8][4], dp[8][4];
        zero_acc(s);
        zero_acc(dp);
        gemm_nt(s, aq, smem_u32(s_k + buf * 8192), lane);
        gemm_nt(dp, ady, smem_u32(s_v + buf * 8192), lane);
        // tile column c = 8 n + 2 t4 + i is key 64 jt + c: lane t = 2 (t4 & 1) + i, bit 16 (jt & 1) + 2 n + (t4 >> 1)
        const int wbase = (jt >> 1) * 4 + 2 * (t4 & 1), bbase = 16 * (jt & 1) + (t4 >> 1);
        const uint32_t w_lo0 = __ldg(mrow_lo + wbase), w_lo1 = __ldg(mrow_lo + wbase + 1);
        const uint32_t w_hi0 = __ldg(mrow_hi + wbase), w_hi1 = __ldg(mrow_hi + wbase + 1);
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            const int bit = bbase + 2 * n;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const uint32_t wl = i ? w_lo1 : w_lo0, wh = i ? w_hi1 : w_hi0;
                float wgt_lo = (float)((wl >> bit) & 1u), wgt_hi = (float)((wh >> bit) & 1u);
                if (jt == 0 && n == 0 && t4 == 0 && i == 0) { wgt_lo += ex_lo; wgt_hi += ex_hi; }
                float p_lo, ds_lo, p_hi, ds_hi;
                bwd_elem(s[n][i], dp[n][i], wgt_lo, iz_lo, dl_lo, scale_log2, clamp_log2, p_lo, ds_lo);
                bwd_elem(s[n][2 + i], dp[n][2 + i], wgt_hi, iz_hi, dl_hi, scale_log2, clamp_log2, p_hi, ds_hi);
                dp[n][i] = ds_lo;
                dp[n][2 + i] = ds_hi;
            }
        }
        uint32_t pa[4][4];
        acc_to_a(pa, dp);
        gemm_nn(acc, pa, smem_u32(s_k + buf * 8192), lane);   // dQ += dS K
        __syncthreads();
    }
    store_slab_bf16(acc, scale, scale, smem_u32(s_q), s_q, warp * 16, dq + hoff, m0 + warp * 16, lane, rs);
}

constexpr int FWD_SMEM = 8192 * 5;
constexpr int BWD_KV_SMEM = 8192 * 6 + (3 * 2 * BM) * 4 + 2 * BM * 4 * 4;
constexpr int BWD_Q_SMEM = 8192 * 6;

}  // namespace attn
}  // namespace spt

using namespace spt;

static int check_attn_args(const char *what, int B, int S, int d, int dtype) {
    if (dtype != SPT_BF16) return fail(SPT_ERR_UNSUPPORTED, "%s: only bf16 is supported on the fused path", what);
    if (d != attn::D) return fail(SPT_ERR_UNSUPPORTED, "%s: head dim %d not supported (64 only)", what, d);
    if (B < 1 || B > 65535 || S < 128 || S % 128 != 0)
        return fail(SPT_ERR_INVALID_ARGUMENT, "%s: need 1 <= B <= 65535 and S a positive multiple of 128 (B=%d S=%d)", what, B, S);
    return SPT_OK;
}

extern "C" int spt_sparse_attn_fwd(const void *q, const void *k, const void *v, const uint32_t *mask,
                                   const int32_t *extra0, void *y, float *zsum, int B, int S, int d, int H,
                                   float scale, float clamp, int dtype, spt_stream_t stream) {
    SPT_REQUIRE(q && k && v && mask && extra0 && y && zsum, "sparse_attn_fwd: null pointer");
    SPT_REQUIRE(H >= 1 && B % H == 0, "sparse_attn_fwd: B=%d must be a multiple of the interleaved head count H=%d", B, H);
    int rc = check_attn_args("sparse_attn_fwd", B, S, d, dtype);
    if (rc != SPT_OK) return rc;
    using bf = __nv_bfloat16;
    cudaFuncSetAttribute(attn::attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, attn::FWD_SMEM);
    attn::attn_fwd_kernel<<<dim3(S / attn::BM, B), attn::THREADS, attn::FWD_SMEM, as_stream(stream)>>>(
        (const bf *)q, (const bf *)k, (const bf *)v, mask, extra0, (bf *)y, zsum, S, H, scale * attn::LOG2E,
        clamp * attn::LOG2E);
    return after_launch("attn_fwd_kernel");
}

extern "C" size_t spt_sparse_attn_bwd_workspace_bytes(int B, int S) { return (size_t)B * S * sizeof(float); }

extern "C" int spt_sparse_attn_bwd(const void *q, const void *k, const void *v, const void *y, const void *grad_y,
                                   const uint32_t *mask, const int32_t *extra0, const float *zsum, void *grad_q,
                                   void *grad_k, void *grad_v, void *workspace, int B, int S, int d, int H,
                                   float scale, float clamp, int dtype, spt_stream_t stream) {
    SPT_REQUIRE(q && k && v && y && grad_y && mask && extra0 && zsum && grad_q && grad_k && grad_v && workspace,
                "sparse_attn_bwd: null pointer");
    int rc = check_attn_args("sparse_attn_bwd", B, S, d, dtype);
    if (rc != SPT_OK) return rc;
    SPT_REQUIRE(H >= 1 && B % H == 0, "sparse_attn_bwd: B=%d must be a multiple of H=%d", B, H);
    using bf = __nv_bfloat16;
    cudaStream_t st = as_stream(stream);
    float *delta = (float *)workspace;
    const int64_t rows = (int64_t)B * S;
    attn::attn_delta_kernel<<<(unsigned)((rows + 31) / 32), 256, 0, st>>>((const bf *)grad_y, (const bf *)y, delta, rows, S, H);
    SPT_LAUNCH_CHECK("attn_delta_kernel");
    cudaFuncSetAttribute(attn::attn_bwd_kv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, attn::BWD_KV_SMEM);
    attn::attn_bwd_kv_kernel<<<dim3(S / attn::BN, B), attn::THREADS, attn::BWD_KV_SMEM, st>>>(
        (const bf *)q, (const bf *)k, (const bf *)v, (const bf *)grad_y, mask, extra0, zsum, delta, (bf *)grad_k,
        (bf *)grad_v, S, H, scale, scale * attn::LOG2E, clamp * attn::LOG2E);
    SPT_LAUNCH_CHECK("attn_bwd_kv_kernel");
    cudaFuncSetAttribute(attn::attn_bwd_q_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, attn::BWD_Q_SMEM);
    attn::attn_bwd_q_kernel<<<dim3(S / attn::BM, B), attn::THREADS, attn::BWD_Q_SMEM, st>>>(
        (const bf *)q, (const bf *)k, (const bf *)v, (const bf *)grad_y, mask, extra0, zsum, delta, (bf *)grad_q, S, H,
        scale, scale * attn::LOG2E, clamp * attn::LOG2E);
    SPT_LAUNCH_CHECK("attn_bwd_q_kernel");
    return SPT_OK;
}
