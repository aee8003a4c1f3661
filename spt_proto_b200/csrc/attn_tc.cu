// Fused PQ-sparse attention, forward + backward, as MASKED DENSE TILES on the 5th-generation tensor
// cores: tcgen05.mma with accumulators in TMEM, operands staged by TMA, the probability tile fed back
// to the second GEMM as a TMEM A-operand (never through shared memory or HBM).
//
// What it replaces: the stage chain  sddmm -> clamp_(scaling * ., -10, 10) -> causal CSR softmax ->
// spmm  of SparseVanillaAttentionV2._get_attn/_apply_attn (reference naive_gpt/layers/sparse/
// attention.py:122-141) and its autograd backward (kernels/spmm.py:23-49, softmax.py:21-30,
// sddmm.py:25-51), including the transposed products dK = dS^T Q, dV = P^T dO.
//
// Why dense tiles: the gathered formulation moves one 128-byte K/V row per selected (query, key)
// pair through L1/shared memory and needs a CSR->CSC transpose for dK/dV; on B200 the dense causal
// tile product is cheaper even though only 1/4 of the causal entries are selected.  The
// selection enters as the lookup kernel's lane-major bitmask (S % 128 == 0):
//     mask[b][r][4 g + t] bit i  <=>  key 128 g + 4 i + t is one of row r's lookup candidates
// plus extra0[b][r] = number of zero-padding slots of the row (they alias key 0 and, like in the
// reference, take part in the softmax).
//
//   w[r][j] = mask bit (+ extra0[r] for j == 0)
//   e[r][j] = w * exp(clamp(scale * q_r.k_j, -10, 10)),  Z_r = max(1e-9, sum_j e),  y_r = sum_j e/Z v_j
// No running max is needed (the clamp bounds the exponent), so there is no rescaling pass.
// Backward recomputes e from q, k (flash style); only Z [B,S] is saved.  With dO' = dO / Z and
// delta' = (dO . y) / Z prepared by a small row kernel:
//   dP' = dO' V^T,  dS = e * (dP' - delta') * [|scale s| <= 10],  dQ = scale dS K
//   dV = E^T dO',   dK = scale dS^T Q      (E = w * exp(...), unnormalised)
// Deterministic: no atomics anywhere.
//
// Kernel anatomy (all three kernels): a CTA owns one 128-row "owner" tile (queries for fwd / dQ, keys
// for dK/dV) and loops over 64-row "other" tiles.  192 threads:
//   warps 0-3 : one thread per owner row = one TMEM lane; tcgen05.ld the score row, mask/exp math,
//               tcgen05.st the bf16 probability row back to TMEM
//   warp  4   : TMA producer (cp.async.bulk.tensor 4-D straight from the [N, S, H, d] layout)
//   warp  5   : TMEM allocator + tcgen05.mma issuer (one thread)
// Two CTAs per SM (256 TMEM columns each) so one CTA's MMAs run under the other's exp math.
#include "tc.cuh"

namespace spt {
namespace attn_tc {

using namespace tc;

constexpr int D = 64;          // head dim
constexpr int BM = 128;        // owner tile rows  (= TMEM lanes)
constexpr int BN = 64;         // other tile rows per iteration
constexpr int STAGES = 3;
constexpr int THREADS = 192;
constexpr int OWN_BYTES = BM * D * 2;   // 16 KB
constexpr int T_BYTES = BN * D * 2;     // 8 KB
constexpr int TMEM_COLS = 256;
constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&h);
}

struct Smem {
    uint32_t base;            // 1024-aligned shared address
    unsigned char *ptr;       // generic pointer to the same byte
};
__device__ __forceinline__ Smem align_smem(unsigned char *raw) {
    const uint32_t a = smem_u32(raw);
    const uint32_t base = (a + 1023) & ~1023u;
    return {base, raw + (base - a)};
}

// store 32 fp32 accumulator values (scaled) as 32 bf16 = 64 contiguous bytes
__device__ __forceinline__ void store_row32(__nv_bfloat16 *dst, const uint32_t (&r)[32], float s) {
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
        float t[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) t[u] = __uint_as_float(r[i + u]) * s;
        Vec16<__nv_bfloat16>::store(dst + i, t);
    }
}

// ---------------------------------------------------------------------------------------------------
// Forward.  TMEM columns: S[2] 0/64 (fp32 128x64), P[2] 128/160 (bf16 pairs, 32 cols), O 192 (128x64).
// ---------------------------------------------------------------------------------------------------
constexpr int FWD_SMEM = OWN_BYTES + 2 * STAGES * T_BYTES + 1024 + 256;

__global__ void __launch_bounds__(THREADS, 2)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                   const __grid_constant__ CUtensorMap map_v, const uint32_t *__restrict__ mask,
                   const int32_t *__restrict__ extra0, __nv_bfloat16 *__restrict__ y, float *__restrict__ zsum, int S,
                   int H, float scale_log2, float clamp_log2) {
    extern __shared__ unsigned char smem_raw[];
    const Smem sm = align_smem(smem_raw);
    const uint32_t s_q = sm.base, s_k = s_q + OWN_BYTES, s_v = s_k + STAGES * T_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(sm.ptr + OWN_BYTES + 2 * STAGES * T_BYTES);
    const uint32_t bar0 = smem_u32(bars);
    const uint32_t q_full = bar0, o_full = bar0 + 8;
    auto k_full = [&](int s) { return bar0 + 16 + s * 8; };
    auto k_empty = [&](int s) { return bar0 + 16 + (STAGES + s) * 8; };
    auto v_full = [&](int s) { return bar0 + 16 + (2 * STAGES + s) * 8; };
    auto v_empty = [&](int s) { return bar0 + 16 + (3 * STAGES + s) * 8; };
    auto s_full = [&](int b) { return bar0 + 16 + (4 * STAGES + b) * 8; };
    auto p_full = [&](int b) { return bar0 + 16 + (4 * STAGES + 2 + b) * 8; };
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 + 4 * STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = gridDim.x - 1 - blockIdx.x;      // heaviest (last) query tiles first
    const int b = blockIdx.y;
    const int m0 = tile * BM;
    const int n_tiles = (m0 + BM) / BN;                // key tiles 0 .. (causal)
    const int hn = b / H, hh = b % H;

    if (threadIdx.x == 0) {
        mbar_init(q_full, 1);
        mbar_init(o_full, 1);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(k_full(s), 1);
            mbar_init(k_empty(s), 1);
            mbar_init(v_full(s), 1);
            mbar_init(v_empty(s), 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(s_full(i), 1);
            mbar_init(p_full(i), 128);
        }
        mbar_fence_init();
    }
    if (warp == 5) tmem_alloc<TMEM_COLS>(smem_u32(tmem_slot));
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    constexpr uint32_t COL_S = 0, COL_P = 128, COL_O = 192;

    if (warp == 4) {
        // ===== TMA producer =====
        if (lane == 0) {
            mbar_expect_tx(q_full, OWN_BYTES);
            tma_load_4d(s_q, &map_q, q_full, 0, hh, m0, hn);
            tma_load_4d(s_q + T_BYTES, &map_q, q_full, 0, hh, m0 + BN, hn);
            for (int j = 0; j < n_tiles; ++j) {
                const int st = j % STAGES;
                const uint32_t ph = (j / STAGES) & 1;
                mbar_wait(k_empty(st), ph ^ 1);
                mbar_expect_tx(k_full(st), T_BYTES);
                tma_load_4d(s_k + st * T_BYTES, &map_k, k_full(st), 0, hh, j * BN, hn);
                mbar_wait(v_empty(st), ph ^ 1);
                mbar_expect_tx(v_full(st), T_BYTES);
                tma_load_4d(s_v + st * T_BYTES, &map_v, v_full(st), 0, hh, j * BN, hn);
            }
        }
    } else if (warp == 5) {
        // ===== MMA issuer =====
        if (lane == 0) {
            constexpr uint32_t id_s = idesc_bf16(BM, BN, 0, 0);   // S = Q K^T   (both K-major)
            constexpr uint32_t id_o = idesc_bf16(BM, D, 0, 1);    // O += P V    (A from TMEM, V MN-major)
            auto issue_s = [&](int j) {
                const int st = j % STAGES;
                mbar_wait(k_full(st), (j / STAGES) & 1);
                fence_after_sync();
#pragma unroll
                for (int k = 0; k < D / 16; ++k)
                    umma_bf16(tmem_base + COL_S + (j & 1) * BN, desc_kmajor(s_q, k), desc_kmajor(s_k + st * T_BYTES, k),
                              id_s, k != 0);
                umma_commit(s_full(j & 1));
                umma_commit(k_empty(st));
            };
            mbar_wait(q_full, 0);
            issue_s(0);
            for (int j = 0; j < n_tiles; ++j) {
                if (j + 1 < n_tiles) issue_s(j + 1);   // S buffer (j+1)&1 was released by p_full of tile j-1
                mbar_wait(p_full(j & 1), (j >> 1) & 1);
                const int st = j % STAGES;
                mbar_wait(v_full(st), (j / STAGES) & 1);
                fence_after_sync();
#pragma unroll
                for (int k = 0; k < BN / 16; ++k)
                    umma_bf16_ts(tmem_base + COL_O, tmem_base + COL_P + (j & 1) * 32 + k * 8,
                                 desc_mnmajor(s_v + st * T_BYTES, k, T_BYTES), id_o, (j | k) != 0);
                umma_commit(v_empty(st));
            }
            umma_commit(o_full);
        }
    } else {
        // ===== softmax warps: thread = query row = TMEM lane =====
        const int row = m0 + warp * 32 + lane;
        const size_t grow = (size_t)b * S + row;
        const uint4 *mrow = reinterpret_cast<const uint4 *>(mask + grow * (S / 32));
        const float ex0 = (float)extra0[grow];
        const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
        float sum = 0.0f;
        uint4 mw = __ldg(mrow), mw_next = mw;
        for (int j = 0; j < n_tiles; ++j) {
            const int bsel = j & 1;
            if (bsel == 0 && j + 2 < n_tiles) mw_next = __ldg(mrow + (j >> 1) + 1);
            mbar_wait(s_full(bsel), (j >> 1) & 1);
            fence_after_sync();
#pragma unroll
            for (int c32 = 0; c32 < 2; ++c32) {
                uint32_t r[32];
                tmem_ld32(lane_base + COL_S + bsel * BN + c32 * 32, r);
                // tile column c = 32 c32 + i is key 64 j + c of group j >> 1: word c & 3, bit 16 (j & 1) + (c >> 2)
                const int b0 = 16 * bsel + 8 * c32;
                const uint32_t w[4] = {mw.x >> b0, mw.y >> b0, mw.z >> b0, mw.w >> b0};
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    const float t0 = fminf(fmaxf(__uint_as_float(r[i]) * scale_log2, -clamp_log2), clamp_log2);
                    const float t1 = fminf(fmaxf(__uint_as_float(r[i + 1]) * scale_log2, -clamp_log2), clamp_log2);
                    float e0 = ex2(t0), e1 = ex2(t1);
                    const uint32_t wa = w[i & 3], wb = w[(i + 1) & 3];
                    if (i == 0) {
                        float wgt = (float)(wa & 1u);
                        if (j == 0 && c32 == 0) wgt += ex0;      // key 0 carries the zero-padding multiplicity
                        e0 *= wgt;
                    } else {
                        e0 = ((wa >> (i >> 2)) & 1u) ? e0 : 0.0f;
                    }
                    e1 = ((wb >> (i >> 2)) & 1u) ? e1 : 0.0f;
                    sum += e0 + e1;
                    pk[i >> 1] = pack_bf16(e0, e1);
                }
                tmem_st16(lane_base + COL_P + bsel * 32 + c32 * 16, pk);
            }
            tmem_st_wait();
            fence_before_sync();
            mbar_arrive(p_full(bsel));
            if (bsel == 1) mw = mw_next;
        }
        sum = fmaxf(sum, 1e-9f);
        zsum[grow] = sum;
        const float inv = 1.0f / sum;
        mbar_wait(o_full, 0);
        fence_after_sync();
        __nv_bfloat16 *dst = y + (((size_t)hn * S + row) * H + hh) * D;
#pragma unroll
        for (int c32 = 0; c32 < 2; ++c32) {
            uint32_t r[32];
            tmem_ld32(lane_base + COL_O + c32 * 32, r);
            store_row32(dst + c32 * 32, r, inv);
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 5) tmem_dealloc<TMEM_COLS>(tmem_base);
}

// ---------------------------------------------------------------------------------------------------
// Backward prologue: delta'_r = (dO_r . y_r) / Z_r  and  dO'_r = dO_r / Z_r (bf16, same layout as dO).
// 8 lanes per row (8 x 16 B = one 128-byte row).
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
attn_bwd_prep_kernel(const __nv_bfloat16 *__restrict__ dy, const __nv_bfloat16 *__restrict__ y,
                     const float *__restrict__ zsum, float *__restrict__ delta, __nv_bfloat16 *__restrict__ dys,
                     int64_t rows, int S, int H) {
    const int64_t row = (int64_t)blockIdx.x * 32 + (threadIdx.x >> 3);
    if (row >= rows) return;
    const int sub = threadIdx.x & 7;
    const int64_t b = row / S, r = row % S;                               // delta, zsum are head-major [B, S]
    const int64_t off = (((b / H) * S + r) * H + (b % H)) * D + sub * 8;
    float a[8], c[8];
    Vec16<__nv_bfloat16>::load(dy + off, a);
    Vec16<__nv_bfloat16>::load(y + off, c);
    float acc = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc = fmaf(a[i], c[i], acc);
    acc = group_sum<8>(acc);
    const float inv = 1.0f / zsum[row];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] *= inv;
    Vec16<__nv_bfloat16>::store(dys + off, a);
    if (sub == 0) delta[row] = acc * inv;
}

// shared element math of the two backward kernels
//   s_raw = q.k, dp = dO'.v, wgt = mask weight  ->  e = wgt * exp(clamp(scale s)), ds = e * (dp - delta') * inside
__device__ __forceinline__ void bwd_elem(float s_raw, float dp, float wgt, float delta, float scale_log2,
                                         float clamp_log2, float &e, float &ds) {
    const float t = s_raw * scale_log2;
    e = wgt * ex2(fminf(fmaxf(t, -clamp_log2), clamp_log2));
    const float g = e * (dp - delta);
    ds = (fabsf(t) <= clamp_log2) ? g : 0.0f;
}

// ---------------------------------------------------------------------------------------------------
// Backward, dQ: owner = 128 query rows (Q, dO'), loop over 64-key tiles (K_j, V_j).
// TMEM columns: S 0 (128x64 fp32; dS bf16 pairs are written back over columns 0..31), dP' 64, dQ 128.
// ---------------------------------------------------------------------------------------------------
constexpr int BWD_SMEM = 2 * OWN_BYTES + 2 * STAGES * T_BYTES + STAGES * (BN * 16 + BN * 4 + BN * 4) + 1024 + 256;

__global__ void __launch_bounds__(THREADS, 2)
attn_bwd_q_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                     const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_dys,
                     const uint32_t *__restrict__ mask, const int32_t *__restrict__ extra0,
                     const float *__restrict__ delta, __nv_bfloat16 *__restrict__ dq, int S, int H, float scale,
                     float scale_log2, float clamp_log2) {
    extern __shared__ unsigned char smem_raw[];
    const Smem sm = align_smem(smem_raw);
    const uint32_t s_q = sm.base, s_dy = s_q + OWN_BYTES, s_k = s_dy + OWN_BYTES, s_v = s_k + STAGES * T_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(sm.ptr + 2 * OWN_BYTES + 2 * STAGES * T_BYTES);
    const uint32_t bar0 = smem_u32(bars);
    const uint32_t own_full = bar0, acc_full = bar0 + 8, sc_full = bar0 + 16, p_full = bar0 + 24;
    auto kv_full = [&](int s) { return bar0 + 32 + s * 8; };
    auto kv_empty = [&](int s) { return bar0 + 32 + (STAGES + s) * 8; };
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 4 + 2 * STAGES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = gridDim.x - 1 - blockIdx.x;
    const int b = blockIdx.y;
    const int m0 = tile * BM;
    const int n_tiles = (m0 + BM) / BN;
    const int hn = b / H, hh = b % H;

    if (threadIdx.x == 0) {
        mbar_init(own_full, 1);
        mbar_init(acc_full, 1);
        mbar_init(sc_full, 1);
        mbar_init(p_full, 128);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(kv_full(s), 1);
            mbar_init(kv_empty(s), 1);
        }
        mbar_fence_init();
    }
    if (warp == 5) tmem_alloc<TMEM_COLS>(smem_u32(tmem_slot));
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    constexpr uint32_t COL_S = 0, COL_DP = 64, COL_DQ = 128;

    if (warp == 4) {
        if (lane == 0) {
            mbar_expect_tx(own_full, 2 * OWN_BYTES);
            tma_load_4d(s_q, &map_q, own_full, 0, hh, m0, hn);
            tma_load_4d(s_q + T_BYTES, &map_q, own_full, 0, hh, m0 + BN, hn);
            tma_load_4d(s_dy, &map_dys, own_full, 0, hh, m0, hn);
            tma_load_4d(s_dy + T_BYTES, &map_dys, own_full, 0, hh, m0 + BN, hn);
            for (int j = 0; j < n_tiles; ++j) {
                const int st = j % STAGES;
                mbar_wait(kv_empty(st), ((j / STAGES) & 1) ^ 1);
                mbar_expect_tx(kv_full(st), 2 * T_BYTES);
                tma_load_4d(s_k + st * T_BYTES, &map_k, kv_full(st), 0, hh, j * BN, hn);
                tma_load_4d(s_v + st * T_BYTES, &map_v, kv_full(st), 0, hh, j * BN, hn);
            }
        }
    } else if (warp == 5) {
        if (lane == 0) {
            constexpr uint32_t id_s = idesc_bf16(BM, BN, 0, 0);   // S = Q K^T, dP' = dO' V^T
            constexpr uint32_t id_a = idesc_bf16(BM, D, 0, 1);    // dQ += dS K   (A from TMEM, K MN-major)
            mbar_wait(own_full, 0);
            for (int j = 0; j < n_tiles; ++j) {
                const int st = j % STAGES;
                mbar_wait(kv_full(st), (j / STAGES) & 1);
                fence_after_sync();
#pragma unroll
                for (int k = 0; k < D / 16; ++k)
                    umma_bf16(tmem_base + COL_S, desc_kmajor(s_q, k), desc_kmajor(s_k + st * T_BYTES, k), id_s, k != 0);
#pragma unroll
                for (int k = 0; k < D / 16; ++k)
                    umma_bf16(tmem_base + COL_DP, desc_kmajor(s_dy, k), desc_kmajor(s_v + st * T_BYTES, k), id_s, k != 0);
                umma_commit(sc_full);
                mbar_wait(p_full, j & 1);
                fence_after_sync();
#pragma unroll
                for (int k = 0; k < BN / 16; ++k)
                    umma_bf16_ts(tmem_base + COL_DQ, tmem_base + COL_S + k * 8,
                                 desc_mnmajor(s_k + st * T_BYTES, k, T_BYTES), id_a, (j | k) != 0);
                umma_commit(kv_empty(st));
            }
            umma_commit(acc_full);
        }
    } else {
        const int row = m0 + warp * 32 + lane;
        const size_t grow = (size_t)b * S + row;
        const uint4 *mrow = reinterpret_cast<const uint4 *>(mask + grow * (S / 32));
        const float ex0 = (float)extra0[grow];
        const float dl = delta[grow];
        const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
        uint4 mw = __ldg(mrow), mw_next = mw;
        for (int j = 0; j < n_tiles; ++j) {
            const int bsel = j & 1;
            if (bsel == 0 && j + 2 < n_tiles) mw_next = __ldg(mrow + (j >> 1) + 1);
            mbar_wait(sc_full, j & 1);
            fence_after_sync();
#pragma unroll
            for (int c32 = 0; c32 < 2; ++c32) {
                uint32_t r[32], g[32];
                tmem_ld32_nowait(lane_base + COL_S + c32 * 32, r);
                tmem_ld32_nowait(lane_base + COL_DP + c32 * 32, g);
                tmem_ld_wait();
                const int b0 = 16 * bsel + 8 * c32;
                const uint32_t w[4] = {mw.x >> b0, mw.y >> b0, mw.z >> b0, mw.w >> b0};
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    float wg0 = (float)((w[i & 3] >> (i >> 2)) & 1u);
                    const float wg1 = (float)((w[(i + 1) & 3] >> (i >> 2)) & 1u);
                    if (i == 0 && j == 0 && c32 == 0) wg0 += ex0;
                    float e0, e1, d0, d1;
                    bwd_elem(__uint_as_float(r[i]), __uint_as_float(g[i]), wg0, dl, scale_log2, clamp_log2, e0, d0);
                    bwd_elem(__uint_as_float(r[i + 1]), __uint_as_float(g[i + 1]), wg1, dl, scale_log2, clamp_log2, e1, d1);
                    pk[i >> 1] = pack_bf16(d0, d1);
                }
                tmem_st16(lane_base + COL_S + c32 * 16, pk);   // dS over the consumed score columns
            }
            tmem_st_wait();
            fence_before_sync();
            mbar_arrive(p_full);
            if (bsel == 1) mw = mw_next;
        }
        mbar_wait(acc_full, 0);
        fence_after_sync();
        __nv_bfloat16 *dst = dq + (((size_t)hn * S + row) * H + hh) * D;
#pragma unroll
        for (int c32 = 0; c32 < 2; ++c32) {
            uint32_t r[32];
            tmem_ld32(lane_base + COL_DQ + c32 * 32, r);
            store_row32(dst + c32 * 32, r, scale);
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 5) tmem_dealloc<TMEM_COLS>(tmem_base);
}

// ---------------------------------------------------------------------------------------------------
// Backward, dK / dV: owner = 128 keys (K, V) = one lane-major mask group, loop over 64-row query tiles
// (Q_j, dO'_j) at and below the diagonal.  Works on the transposed tiles S^T = K Q^T, dP'^T = V dO'^T
// so that E^T and dS^T come out with keys on the TMEM lanes, ready to be the A operands of
//   dV += E^T dO'_j   and   dK += dS^T Q_j.
// TMEM columns: S^T 0 (E^T bf16 written back over 0..31), dP'^T 64 (dS^T over 64..95), dV 128, dK 192.
// Per-row quantities of the 64 query rows (mask words of this key group, delta', extra0) are staged in
// shared memory by the producer warp, one slot per pipeline stage.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(THREADS, 2)
attn_bwd_kv_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                      const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_dys,
                      const uint32_t *__restrict__ mask, const int32_t *__restrict__ extra0,
                      const float *__restrict__ delta, __nv_bfloat16 *__restrict__ dk, __nv_bfloat16 *__restrict__ dv,
                      int S, int H, float scale, float scale_log2, float clamp_log2) {
    extern __shared__ unsigned char smem_raw[];
    const Smem sm = align_smem(smem_raw);
    const uint32_t s_k = sm.base, s_v = s_k + OWN_BYTES, s_q = s_v + OWN_BYTES, s_dy = s_q + STAGES * T_BYTES;
    unsigned char *rowq = sm.ptr + 2 * OWN_BYTES + 2 * STAGES * T_BYTES;   // per stage: mask uint4[64], delta[64], ex0[64]
    constexpr int ROWQ_BYTES = BN * 16 + BN * 4 + BN * 4;
    uint64_t *bars = reinterpret_cast<uint64_t *>(rowq + STAGES * ROWQ_BYTES);
    const uint32_t bar0 = smem_u32(bars);
    const uint32_t own_full = bar0, acc_full = bar0 + 8, sc_full = bar0 + 16, p_full = bar0 + 24;
    auto qd_full = [&](int s) { return bar0 + 32 + s * 8; };
    auto qd_empty = [&](int s) { return bar0 + 32 + (STAGES + s) * 8; };
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 4 + 2 * STAGES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kt = blockIdx.x;                         // key tile (early key tiles are the heaviest)
    const int b = blockIdx.y;
    const int n0 = kt * BM;
    const int j0 = n0 / BN;                            // first query tile that can see these keys
    const int n_tiles = S / BN - j0;
    const int hn = b / H, hh = b % H;
    const size_t head = (size_t)b * S;
    const int words = S / 32;

    if (threadIdx.x == 0) {
        mbar_init(own_full, 1);
        mbar_init(acc_full, 1);
        mbar_init(sc_full, 1);
        mbar_init(p_full, 128);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(qd_full(s), 1 + 32);             // expect_tx arrive + the 32 producer lanes' row data
            mbar_init(qd_empty(s), 1);
        }
        mbar_fence_init();
    }
    if (warp == 5) tmem_alloc<TMEM_COLS>(smem_u32(tmem_slot));
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    constexpr uint32_t COL_S = 0, COL_DP = 64, COL_DV = 128, COL_DK = 192;

    if (warp == 4) {
        if (lane == 0) {
            mbar_expect_tx(own_full, 2 * OWN_BYTES);
            tma_load_4d(s_k, &map_k, own_full, 0, hh, n0, hn);
            tma_load_4d(s_k + T_BYTES, &map_k, own_full, 0, hh, n0 + BN, hn);
            tma_load_4d(s_v, &map_v, own_full, 0, hh, n0, hn);
            tma_load_4d(s_v + T_BYTES, &map_v, own_full, 0, hh, n0 + BN, hn);
        }
        for (int j = 0; j < n_tiles; ++j) {
            const int st = j % STAGES;
            const int r0 = (j0 + j) * BN;
            mbar_wait(qd_empty(st), ((j / STAGES) & 1) ^ 1);
            if (lane == 0) {
                mbar_expect_tx(qd_full(st), 2 * T_BYTES);
                tma_load_4d(s_q + st * T_BYTES, &map_q, qd_full(st), 0, hh, r0, hn);
                tma_load_4d(s_dy + st * T_BYTES, &map_dys, qd_full(st), 0, hh, r0, hn);
            }
            unsigned char *slot = rowq + st * ROWQ_BYTES;
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int rr = lane + 32 * u;
                const size_t gr = head + r0 + rr;
                reinterpret_cast<uint4 *>(slot)[rr] = __ldg(reinterpret_cast<const uint4 *>(mask + gr * words) + kt);
                reinterpret_cast<float *>(slot + BN * 16)[rr] = delta[gr];
                reinterpret_cast<float *>(slot + BN * 16 + BN * 4)[rr] = (float)extra0[gr];
            }
            mbar_arrive(qd_full(st));
        }
    } else if (warp == 5) {
        if (lane == 0) {
            constexpr uint32_t id_s = idesc_bf16(BM, BN, 0, 0);   // S^T = K Q^T, dP'^T = V dO'^T
            constexpr uint32_t id_a = idesc_bf16(BM, D, 0, 1);    // dV += E^T dO', dK += dS^T Q (B MN-major)
            mbar_wait(own_full, 0);
            for (int j = 0; j < n_tiles; ++j) {
                const int st = j % STAGES;
                mbar_wait(qd_full(st), (j / STAGES) & 1);
                fence_after_sync();
#pragma unroll
                for (int k = 0; k < D / 16; ++k)
                    umma_bf16(tmem_base + COL_S, desc_kmajor(s_k, k), desc_kmajor(s_q + st * T_BYTES, k), id_s, k != 0);
#pragma unroll
                for (int k = 0; k < D / 16; ++k)
                    umma_bf16(tmem_base + COL_DP, desc_kmajor(s_v, k), desc_kmajor(s_dy + st * T_BYTES, k), id_s, k != 0);
                umma_commit(sc_full);
                mbar_wait(p_full, j & 1);
                fence_after_sync();
#pragma unroll
                for (int k = 0; k < BN / 16; ++k)
                    umma_bf16_ts(tmem_base + COL_DV, tmem_base + COL_S + k * 8,
                                 desc_mnmajor(s_dy + st * T_BYTES, k, T_BYTES), id_a, (j | k) != 0);
#pragma unroll
                for (int k = 0; k < BN / 16; ++k)
                    umma_bf16_ts(tmem_base + COL_DK, tmem_base + COL_DP + k * 8,
                                 desc_mnmajor(s_q + st * T_BYTES, k, T_BYTES), id_a, (j | k) != 0);
                umma_commit(qd_empty(st));
            }
            umma_commit(acc_full);
        }
    } else {
        // thread = key n0 + kk = TMEM lane kk: lane-major word kk & 3 of the row's group kt, bit kk >> 2
        const int kk = warp * 32 + lane;
        const int wsel = kk & 3, bit = kk >> 2;
        const bool key0 = (n0 + kk) == 0;
        const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
        for (int j = 0; j < n_tiles; ++j) {
            const int st = j % STAGES;
            const unsigned char *slot = rowq + st * ROWQ_BYTES;
            const uint32_t *s_mask = reinterpret_cast<const uint32_t *>(slot);
            const float *s_delta = reinterpret_cast<const float *>(slot + BN * 16);
            const float *s_ex0 = s_delta + BN;
            mbar_wait(qd_full(st), (j / STAGES) & 1);   // the producer lanes' row data of this stage
            mbar_wait(sc_full, j & 1);
            fence_after_sync();
#pragma unroll
            for (int c32 = 0; c32 < 2; ++c32) {
                uint32_t r[32], g[32];
                tmem_ld32_nowait(lane_base + COL_S + c32 * 32, r);
                tmem_ld32_nowait(lane_base + COL_DP + c32 * 32, g);
                tmem_ld_wait();
                uint32_t pe[16], pd[16];
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    const int c = c32 * 32 + i;                      // query row inside the tile
                    float wg0 = (float)((s_mask[c * 4 + wsel] >> bit) & 1u);
                    float wg1 = (float)((s_mask[(c + 1) * 4 + wsel] >> bit) & 1u);
                    if (key0) {
                        wg0 += s_ex0[c];
                        wg1 += s_ex0[c + 1];
                    }
                    float e0, e1, d0, d1;
                    bwd_elem(__uint_as_float(r[i]), __uint_as_float(g[i]), wg0, s_delta[c], scale_log2, clamp_log2, e0, d0);
                    bwd_elem(__uint_as_float(r[i + 1]), __uint_as_float(g[i + 1]), wg1, s_delta[c + 1], scale_log2,
                             clamp_log2, e1, d1);
                    pe[i >> 1] = pack_bf16(e0, e1);
                    pd[i >> 1] = pack_bf16(d0, d1);
                }
                tmem_st16(lane_base + COL_S + c32 * 16, pe);    // E^T over the consumed S^T columns
                tmem_st16(lane_base + COL_DP + c32 * 16, pd);   // dS^T over the consumed dP'^T columns
            }
            tmem_st_wait();
            fence_before_sync();
            mbar_arrive(p_full);
        }
        mbar_wait(acc_full, 0);
        fence_after_sync();
        const size_t off = (((size_t)hn * S + n0 + kk) * H + hh) * D;
#pragma unroll
        for (int c32 = 0; c32 < 2; ++c32) {
            uint32_t r[32];
            tmem_ld32(lane_base + COL_DV + c32 * 32, r);
            store_row32(dv + off + c32 * 32, r, 1.0f);
            tmem_ld32(lane_base + COL_DK + c32 * 32, r);
            store_row32(dk + off + c32 * 32, r, scale);
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 5) tmem_dealloc<TMEM_COLS>(tmem_base);
}

// ---- host -----------------------------------------------------------------------------------------
// 4-D bf16 tensor map over a [N, S, H, D] tensor (H = 1: head-major [B, S, D]); box = 64 rows of one head
static int make_map(CUtensorMap *map, const void *base, int N, int S, int H) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return fail(SPT_ERR_CUDA, "sparse_attn: cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[4] = {(cuuint64_t)D, (cuuint64_t)H, (cuuint64_t)S, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)D * 2, (cuuint64_t)H * D * 2, (cuuint64_t)S * H * D * 2};
    cuuint32_t box[4] = {(cuuint32_t)D, 1, (cuuint32_t)BN, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(SPT_ERR_CUDA, "sparse_attn: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return SPT_OK;
}

}  // namespace attn_tc
}  // namespace spt

using namespace spt;

static int check_attn_args(const char *what, int B, int S, int d, int H, int dtype) {
    if (dtype != SPT_BF16) return fail(SPT_ERR_UNSUPPORTED, "%s: only bf16 is supported on the fused path", what);
    if (d != attn_tc::D) return fail(SPT_ERR_UNSUPPORTED, "%s: head dim %d not supported (64 only)", what, d);
    if (B < 1 || B > 65535 || S < 128 || S % 128 != 0)
        return fail(SPT_ERR_INVALID_ARGUMENT, "%s: need 1 <= B <= 65535 and S a positive multiple of 128 (B=%d S=%d)", what, B, S);
    if (H < 1 || B % H != 0)
        return fail(SPT_ERR_INVALID_ARGUMENT, "%s: B=%d must be a multiple of the interleaved head count H=%d", what, B, H);
    return SPT_OK;
}

extern "C" int spt_sparse_attn_fwd(const void *q, const void *k, const void *v, const uint32_t *mask,
                                   const int32_t *extra0, void *y, float *zsum, int B, int S, int d, int H,
                                   float scale, float clamp, int dtype, spt_stream_t stream) {
    SPT_REQUIRE(q && k && v && mask && extra0 && y && zsum, "sparse_attn_fwd: null pointer");
    int rc = check_attn_args("sparse_attn_fwd", B, S, d, H, dtype);
    if (rc != SPT_OK) return rc;
    SPT_REQUIRE(((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)y) % 16 == 0, "sparse_attn_fwd: operands must be 16-byte aligned");
    using bf = __nv_bfloat16;
    CUtensorMap mq, mk, mv;
    if ((rc = attn_tc::make_map(&mq, q, B / H, S, H)) != SPT_OK) return rc;
    if ((rc = attn_tc::make_map(&mk, k, B / H, S, H)) != SPT_OK) return rc;
    if ((rc = attn_tc::make_map(&mv, v, B / H, S, H)) != SPT_OK) return rc;
    cudaFuncSetAttribute(attn_tc::attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, attn_tc::FWD_SMEM);
    attn_tc::attn_fwd_tc_kernel<<<dim3(S / attn_tc::BM, B), attn_tc::THREADS, attn_tc::FWD_SMEM, as_stream(stream)>>>(
        mq, mk, mv, mask, extra0, (bf *)y, zsum, S, H, scale * attn_tc::LOG2E, clamp * attn_tc::LOG2E);
    return after_launch("attn_fwd_tc_kernel");
}

// workspace: delta' [B, S] fp32, then dO' (bf16, same shape as grad_y)
extern "C" size_t spt_sparse_attn_bwd_workspace_bytes(int B, int S) {
    return (size_t)B * S * sizeof(float) + (size_t)B * S * attn_tc::D * 2;
}

extern "C" int spt_sparse_attn_bwd(const void *q, const void *k, const void *v, const void *y, const void *grad_y,
                                   const uint32_t *mask, const int32_t *extra0, const float *zsum, void *grad_q,
                                   void *grad_k, void *grad_v, void *workspace, int B, int S, int d, int H,
                                   float scale, float clamp, int dtype, spt_stream_t stream) {
    SPT_REQUIRE(q && k && v && y && grad_y && mask && extra0 && zsum && grad_q && grad_k && grad_v && workspace,
                "sparse_attn_bwd: null pointer");
    int rc = check_attn_args("sparse_attn_bwd", B, S, d, H, dtype);
    if (rc != SPT_OK) return rc;
    SPT_REQUIRE(((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)grad_y | (uintptr_t)workspace) % 16 == 0,
                "sparse_attn_bwd: operands must be 16-byte aligned");
    using bf = __nv_bfloat16;
    cudaStream_t st = as_stream(stream);
    float *delta = (float *)workspace;
    bf *dys = (bf *)((char *)workspace + (size_t)B * S * sizeof(float));
    const int64_t rows = (int64_t)B * S;
    attn_tc::attn_bwd_prep_kernel<<<(unsigned)((rows + 31) / 32), 256, 0, st>>>((const bf *)grad_y, (const bf *)y, zsum,
                                                                                delta, dys, rows, S, H);
    SPT_LAUNCH_CHECK("attn_bwd_prep_kernel");
    CUtensorMap mq, mk, mv, md;
    if ((rc = attn_tc::make_map(&mq, q, B / H, S, H)) != SPT_OK) return rc;
    if ((rc = attn_tc::make_map(&mk, k, B / H, S, H)) != SPT_OK) return rc;
    if ((rc = attn_tc::make_map(&mv, v, B / H, S, H)) != SPT_OK) return rc;
    if ((rc = attn_tc::make_map(&md, dys, B / H, S, H)) != SPT_OK) return rc;
    cudaFuncSetAttribute(attn_tc::attn_bwd_kv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, attn_tc::BWD_SMEM);
    attn_tc::attn_bwd_kv_tc_kernel<<<dim3(S / attn_tc::BM, B), attn_tc::THREADS, attn_tc::BWD_SMEM, st>>>(
        mq, mk, mv, md, mask, extra0, delta, (bf *)grad_k, (bf *)grad_v, S, H, scale, scale * attn_tc::LOG2E,
        clamp * attn_tc::LOG2E);
    SPT_LAUNCH_CHECK("attn_bwd_kv_tc_kernel");
    cudaFuncSetAttribute(attn_tc::attn_bwd_q_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, attn_tc::BWD_SMEM);
    attn_tc::attn_bwd_q_tc_kernel<<<dim3(S / attn_tc::BM, B), attn_tc::THREADS, attn_tc::BWD_SMEM, st>>>(
        mq, mk, mv, md, mask, extra0, delta, (bf *)grad_q, S, H, scale, scale * attn_tc::LOG2E, clamp * attn_tc::LOG2E);
    SPT_LAUNCH_CHECK("attn_bwd_q_tc_kernel");
    return SPT_OK;
}
