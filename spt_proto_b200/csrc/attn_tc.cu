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
// for dK/dV) and loops over 64-row "other" tiles.  320 threads:
//   warps 0-7 : two warpgroups; thread = (owner row = TMEM lane, column half): tcgen05.ld its 32 score
//               columns, mask/exp math, tcgen05.st the bf16 probability columns back to TMEM
//   warp  8   : TMA producer (cp.async.bulk.tensor 4-D straight from the [N, S, H, d] layout)
//   warp  9   : TMEM allocator + tcgen05.mma issuer (one thread)
// Two CTAs per SM (256 TMEM columns each): 16 math warps per SM keep the XU/FMA pipes fed while the
// other CTA's MMAs and barrier round trips are in flight.  Score tiles of the next iteration are
// issued while the current one is still in the math phase (double-buffered S in the forward; in the
// dQ kernel the score columns are released as soon as they are in registers).
#include "tc.cuh"

namespace spt {
namespace attn_tc {

using namespace tc;

constexpr int BM = 128;        // owner tile rows  (= TMEM lanes)
constexpr int BN = 64;         // other tile rows per iteration
constexpr int STAGES = 3;
constexpr int N_MATH = 256;      // warps 0-7
constexpr int THREADS = 320;
// Head dim D in {64, 128}.  Operand tiles are stored as D / 64 sub-tiles of [rows][64] bf16 (one TMA box
// each, 128-byte swizzled): an owner tile is [D/64][128 rows][64], an "other" tile [D/64][64 rows][64].
// D = 64 needs 256 TMEM columns (two CTAs per SM), D = 128 takes the whole TMEM (one CTA per SM).
template <int D>
struct Dim {
    static_assert(D == 64 || D == 128, "head dim must be 64 or 128");
    static constexpr int NSUB = D / 64;
    static constexpr int OWN_SUB = BM * 64 * 2, T_SUB = BN * 64 * 2;          // 16 KB, 8 KB
    static constexpr int OWN_BYTES = NSUB * OWN_SUB, T_BYTES = NSUB * T_SUB;
    static constexpr int TMEM_COLS = D == 64 ? 256 : 512;
    static constexpr int CTAS = D == 64 ? 2 : 1;
    static constexpr int FWD_SMEM = OWN_BYTES + 2 * STAGES * T_BYTES + 1024 /*row sums*/ + 1024 /*align*/ + 256 /*barriers*/;
    static constexpr int BWD_SMEM = 2 * OWN_BYTES + 2 * STAGES * T_BYTES + STAGES * ((BN + 4) * 16 + BN * 4 + BN * 4) + 1024 + 256;
};
// descriptor offset (16-byte units) of K16 slice k of a K-major tile whose 64-wide sub-tiles are SUB bytes apart
template <int SUB>
__device__ constexpr uint64_t kslice(int k) { return (uint64_t)((k >> 2) * (SUB >> 4) + (k & 3) * 2); }
// TMA: 128-row owner tile / 64-row other tile of head (hn, hh) starting at sequence row `row0`
template <int D>
__device__ __forceinline__ void tma_owner(uint32_t dst, const CUtensorMap *map, uint32_t bar, int hh, int row0, int hn) {
#pragma unroll
    for (int dh = 0; dh < Dim<D>::NSUB; ++dh) {
        tma_load_4d(dst + dh * Dim<D>::OWN_SUB, map, bar, dh * 64, hh, row0, hn);
        tma_load_4d(dst + dh * Dim<D>::OWN_SUB + Dim<D>::T_SUB, map, bar, dh * 64, hh, row0 + BN, hn);
    }
}
template <int D>
__device__ __forceinline__ void tma_other(uint32_t dst, const CUtensorMap *map, uint32_t bar, int hh, int row0, int hn) {
#pragma unroll
    for (int dh = 0; dh < Dim<D>::NSUB; ++dh) tma_load_4d(dst + dh * Dim<D>::T_SUB, map, bar, dh * 64, hh, row0, hn);
}
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
// The element math is written for pipe balance: per element the XU pipe (ex2) is the floor, so
// everything else is pushed onto the FMA pipe and kept off the half-rate ALU pipe.
//   clamp:  u = sat(s * a + 0.5) with a = scale / (2 clamp)  (FFMA.SAT), arg = u * 2L - L  (FFMA)
//           replaces FMUL + 2 FMNMX (ALU);  |scale s| <= clamp  <=>  the unsaturated value equals u
//   mask :  bit test straight into a predicate (one LOP3), then a predicated multiply by zero (FMA
//           pipe) instead of shift + and + int->float + FSEL
__device__ __forceinline__ float fma_sat(float a, float b, float c) {
    float d;
    asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
__device__ __forceinline__ float sat01(float v) {
    float d;
    asm("add.sat.f32 %0, %1, 0f00000000;" : "=f"(d) : "f"(v));
    return d;
}
__device__ __forceinline__ float keep_if_bit(float e, uint32_t w, uint32_t bitmask) {   // e if w & bitmask, else 0
    asm("{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b32 t;\n\t"
        "and.b32 t, %1, %2;\n\t"
        "setp.eq.u32 p, t, 0;\n\t"
        "@p mul.f32 %0, %0, 0f00000000;\n\t"
        "}" : "+f"(e) : "r"(w), "r"(bitmask));
    return e;
}
__device__ __forceinline__ float zero_if_ne(float g, float u, float v) {   // g if u == v, else 0
    asm("{\n\t"
        ".reg .pred p;\n\t"
        "setp.neu.f32 p, %1, %2;\n\t"
        "@p mul.f32 %0, %0, 0f00000000;\n\t"
        "}" : "+f"(g) : "f"(u), "f"(v));
    return g;
}


struct ClampK {     // constants of the FMA-pipe clamp: u = sat(s * a + 0.5), arg = u * two_l + neg_l
    float a, two_l, neg_l;
};
__device__ __forceinline__ ClampK make_clamp(float scale_log2, float clamp_log2) {
    return {scale_log2 / (2.0f * clamp_log2), 2.0f * clamp_log2, -clamp_log2};
}

// Forward: 32 score columns of one row -> 16 packed bf16 pairs of e = w * exp(clamp(scale s)), row sum.
// w[t] = lane-major mask word t shifted so that column i tests bit i >> 2.  FIRST: column 0 is key 0.
template <bool FIRST>
__device__ __forceinline__ void fwd_chunk32(const uint32_t (&r)[32], const uint32_t (&w)[4], const ClampK ck, float ex0,
                                            float &sum, uint32_t (&pk)[16]) {
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
        float e0 = ex2(fmaf(fma_sat(__uint_as_float(r[i]), ck.a, 0.5f), ck.two_l, ck.neg_l));
        float e1 = ex2(fmaf(fma_sat(__uint_as_float(r[i + 1]), ck.a, 0.5f), ck.two_l, ck.neg_l));
        if (FIRST && i == 0) e0 *= (float)(w[0] & 1u) + ex0;     // key 0 carries the zero-padding multiplicity
        else e0 = keep_if_bit(e0, w[i & 3], 1u << (i >> 2));
        e1 = keep_if_bit(e1, w[(i + 1) & 3], 1u << (i >> 2));
        sum += e0 + e1;
        pk[i >> 1] = pack_bf16(e0, e1);
    }
}

// Backward element: s = q.k, dp = dO'.v, masked weight applied by the caller through `e`.
//   e_raw = exp(clamp(scale s));   ds = e * (dp - delta') * [|scale s| <= clamp]
__device__ __forceinline__ float bwd_exp(float s_raw, const ClampK ck, float &u, float &v) {
    v = fmaf(s_raw, ck.a, 0.5f);
    u = sat01(v);
    return ex2(fmaf(u, ck.two_l, ck.neg_l));
}
__device__ __forceinline__ float bwd_ds(float e, float dp, float delta, float u, float v) {
    return zero_if_ne(e * (dp - delta), u, v);
}

// dQ kernel: 32 columns (keys) of one query row.  w[t] as in fwd_chunk32.
template <bool FIRST>
__device__ __forceinline__ void bwdq_chunk32(const uint32_t (&r)[32], const uint32_t (&g)[32], const uint32_t (&w)[4],
                                             const ClampK ck, float delta, float ex0, uint32_t (&pk)[16]) {
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
        float u0, v0, u1, v1;
        float e0 = bwd_exp(__uint_as_float(r[i]), ck, u0, v0);
        float e1 = bwd_exp(__uint_as_float(r[i + 1]), ck, u1, v1);
        if (FIRST && i == 0) e0 *= (float)(w[0] & 1u) + ex0;
        else e0 = keep_if_bit(e0, w[i & 3], 1u << (i >> 2));
        e1 = keep_if_bit(e1, w[(i + 1) & 3], 1u << (i >> 2));
        pk[i >> 1] = pack_bf16(bwd_ds(e0, __uint_as_float(g[i]), delta, u0, v0),
                               bwd_ds(e1, __uint_as_float(g[i + 1]), delta, u1, v1));
    }
}

// dK/dV kernel: 16 columns (query rows c0 .. c0+15) of one key.  mrow = this key's lane-major word of every
// row, [64] in shared memory; bitmask selects the key's bit.  KEY0: this thread is key 0 (adds extra0[row]).
template <bool KEY0>
__device__ __forceinline__ void bwdkv_chunk16(const uint32_t (&r)[16], const uint32_t (&g)[16], const uint32_t *mrow,
                                              const float *s_delta, const float *s_ex0, int c0, uint32_t bitmask,
                                              const ClampK ck, uint32_t (&pe)[8], uint32_t (&pd)[8]) {
#pragma unroll
    for (int q4 = 0; q4 < 16; q4 += 4) {
        const uint4 mw = *reinterpret_cast<const uint4 *>(mrow + c0 + q4);
        const float4 dl = *reinterpret_cast<const float4 *>(s_delta + c0 + q4);
        const uint32_t mwv[4] = {mw.x, mw.y, mw.z, mw.w};
        const float dlv[4] = {dl.x, dl.y, dl.z, dl.w};
        float e[4], d[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            float u, v;
            e[t] = bwd_exp(__uint_as_float(r[q4 + t]), ck, u, v);
            if (KEY0) e[t] *= (float)((mwv[t] & bitmask) != 0u) + s_ex0[c0 + q4 + t];
            else e[t] = keep_if_bit(e[t], mwv[t], bitmask);
            d[t] = bwd_ds(e[t], __uint_as_float(g[q4 + t]), dlv[t], u, v);
        }
        pe[(q4 >> 1)] = pack_bf16(e[0], e[1]);
        pe[(q4 >> 1) + 1] = pack_bf16(e[2], e[3]);
        pd[(q4 >> 1)] = pack_bf16(d[0], d[1]);
        pd[(q4 >> 1) + 1] = pack_bf16(d[2], d[3]);
    }
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
__device__ __forceinline__ void math_warps_sync() { asm volatile("bar.sync 1, %0;" ::"n"(N_MATH) : "memory"); }

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

template <int D>
__global__ void __launch_bounds__(THREADS, Dim<D>::CTAS)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                   const __grid_constant__ CUtensorMap map_v, const uint32_t *__restrict__ mask,
                   const int32_t *__restrict__ extra0, __nv_bfloat16 *__restrict__ y, float *__restrict__ zsum, int S,
                   int H, float scale_log2, float clamp_log2) {
    constexpr int OWN_BYTES = Dim<D>::OWN_BYTES, T_BYTES = Dim<D>::T_BYTES, OWN_SUB = Dim<D>::OWN_SUB, T_SUB = Dim<D>::T_SUB,
                  TMEM_COLS = Dim<D>::TMEM_COLS;
    extern __shared__ unsigned char smem_raw[];
    const Smem sm = align_smem(smem_raw);
    const uint32_t s_q = sm.base, s_k = s_q + OWN_BYTES, s_v = s_k + STAGES * T_BYTES;
    float *s_part = reinterpret_cast<float *>(sm.ptr + OWN_BYTES + 2 * STAGES * T_BYTES);   // [2][128] partial row sums
    uint64_t *bars = reinterpret_cast<uint64_t *>(sm.ptr + OWN_BYTES + 2 * STAGES * T_BYTES + 1024);
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
            mbar_init(p_full(i), N_MATH);
        }
        mbar_fence_init();
    }
    if (warp == 9) tmem_alloc<TMEM_COLS>(smem_u32(tmem_slot));
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    constexpr uint32_t COL_S = 0, COL_P = 128, COL_O = 192;

    if (warp == 8) {
        // ===== TMA producer =====
        if (lane == 0) {
            mbar_expect_tx(q_full, OWN_BYTES);
            tma_owner<D>(s_q, &map_q, q_full, hh, m0, hn);
            for (int j = 0; j < n_tiles; ++j) {
                const int st = j % STAGES;
                const uint32_t ph = (j / STAGES) & 1;
                mbar_wait(k_empty(st), ph ^ 1);
                mbar_expect_tx(k_full(st), T_BYTES);
                tma_other<D>(s_k + st * T_BYTES, &map_k, k_full(st), hh, j * BN, hn);
                mbar_wait(v_empty(st), ph ^ 1);
                mbar_expect_tx(v_full(st), T_BYTES);
                tma_other<D>(s_v + st * T_BYTES, &map_v, v_full(st), hh, j * BN, hn);
            }
        }
    } else if (warp == 9) {
        // ===== MMA issuer: the whole warp runs the (uniform) loop, one elected lane issues =====
        constexpr uint32_t id_s = idesc_bf16(BM, BN, 0, 0);   // S = Q K^T   (both K-major)
        constexpr uint32_t id_o = idesc_bf16(BM, D, 0, 1);    // O += P V    (A from TMEM, V MN-major)
        const uint64_t dq0 = desc_kmajor(s_q, 0), dk0 = desc_kmajor(s_k, 0), dv0 = desc_mnmajor(s_v, 0, T_SUB);
        auto issue_s = [&](int j) {
            const int st = j % STAGES;
            mbar_wait(k_full(st), (j / STAGES) & 1);
            fence_after_sync();
            if (elect_one()) {
                const uint64_t dk = dk0 + (uint64_t)(st * (T_BYTES >> 4));
#pragma unroll
                for (int k = 0; k < D / 16; ++k)
                    umma_bf16(tmem_base + COL_S + (j & 1) * BN, dq0 + kslice<OWN_SUB>(k), dk + kslice<T_SUB>(k), id_s, k != 0);
                umma_commit(s_full(j & 1));
                umma_commit(k_empty(st));
            }
            __syncwarp();
        };
        mbar_wait(q_full, 0);
        issue_s(0);
        for (int j = 0; j < n_tiles; ++j) {
            if (j + 1 < n_tiles) issue_s(j + 1);   // S buffer (j+1)&1 was released by p_full of tile j-1
            mbar_wait(p_full(j & 1), (j >> 1) & 1);
            const int st = j % STAGES;
            mbar_wait(v_full(st), (j / STAGES) & 1);
            fence_after_sync();
            if (elect_one()) {
                const uint64_t dv = dv0 + (uint64_t)(st * (T_BYTES >> 4));
#pragma unroll
                for (int k = 0; k < BN / 16; ++k)
                    umma_bf16_ts(tmem_base + COL_O, tmem_base + COL_P + (j & 1) * 32 + k * 8, dv + k * MNMAJOR_K16, id_o,
                                 (j | k) != 0);
                umma_commit(v_empty(st));
                if (j + 1 == n_tiles) umma_commit(o_full);
            }
            __syncwarp();
        }
    } else {
        // ===== softmax warps: thread = (query row = TMEM lane, column half) =====
        const int quarter = warp & 3, half = warp >> 2;
        const int rt = quarter * 32 + lane;               // row inside the tile
        const int row = m0 + rt;
        const size_t grow = (size_t)b * S + row;
        const uint4 *mrow = reinterpret_cast<const uint4 *>(mask + grow * (S / 32));
        const float ex0 = (float)extra0[grow];
        const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
        float sum = 0.0f;
        const ClampK ck = make_clamp(scale_log2, clamp_log2);
        uint4 mw = __ldg(mrow), mw_next = mw;
        for (int j = 0; j < n_tiles; ++j) {
            const int bsel = j & 1;
            if (bsel == 0 && j + 2 < n_tiles) mw_next = __ldg(mrow + (j >> 1) + 1);
            mbar_wait(s_full(bsel), (j >> 1) & 1);
            fence_after_sync();
            uint32_t r[32];
            tmem_ld32(lane_base + COL_S + bsel * BN + half * 32, r);
            // tile column c = 32 half + i is key 64 j + c of group j >> 1: word c & 3, bit 16 (j & 1) + (c >> 2)
            const int b0 = 16 * bsel + 8 * half;
            const uint32_t w[4] = {mw.x >> b0, mw.y >> b0, mw.z >> b0, mw.w >> b0};
            uint32_t pk[16];
            if (j == 0 && half == 0) fwd_chunk32<true>(r, w, ck, ex0, sum, pk);
            else fwd_chunk32<false>(r, w, ck, ex0, sum, pk);
            tmem_st16(lane_base + COL_P + bsel * 32 + half * 16, pk);
            tmem_st_wait();
            fence_before_sync();
            mbar_arrive(p_full(bsel));
            if (bsel == 1) mw = mw_next;
        }
        s_part[half * BM + rt] = sum;
        math_warps_sync();
        sum = fmaxf(s_part[rt] + s_part[BM + rt], 1e-9f);
        if (half == 0) zsum[grow] = sum;
        const float inv = 1.0f / sum;
        mbar_wait(o_full, 0);
        fence_after_sync();
        __nv_bfloat16 *dst = y + (((size_t)hn * S + row) * H + hh) * D + half * (D / 2);
#pragma unroll
        for (int c = 0; c < D / 64; ++c) {
            uint32_t r[32];
            tmem_ld32(lane_base + COL_O + half * (D / 2) + c * 32, r);
            store_row32(dst + c * 32, r, inv);
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 9) tmem_dealloc<TMEM_COLS>(tmem_base);
}

// ---------------------------------------------------------------------------------------------------
// Backward prologue: delta'_r = (dO_r . y_r) / Z_r  and  dO'_r = dO_r / Z_r (bf16, same layout as dO).
// 8 lanes per row (8 x 16 B = one 128-byte row).
// ---------------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(256)
attn_bwd_prep_kernel(const __nv_bfloat16 *__restrict__ dy, const __nv_bfloat16 *__restrict__ y,
                     const float *__restrict__ zsum, float *__restrict__ delta, __nv_bfloat16 *__restrict__ dys,
                     int64_t rows, int S, int H) {
    constexpr int LPR = D / 8;                                             // lanes per row (16 B each)
    const int64_t row = (int64_t)blockIdx.x * (256 / LPR) + (threadIdx.x / LPR);
    if (row >= rows) return;
    const int sub = threadIdx.x % LPR;
    const int64_t b = row / S, r = row % S;                               // delta, zsum are head-major [B, S]
    const int64_t off = (((b / H) * S + r) * H + (b % H)) * D + sub * 8;
    float a[8], c[8];
    Vec16<__nv_bfloat16>::load(dy + off, a);
    Vec16<__nv_bfloat16>::load(y + off, c);
    float acc = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc = fmaf(a[i], c[i], acc);
    acc = group_sum<LPR>(acc);
    const float inv = 1.0f / zsum[row];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] *= inv;
    Vec16<__nv_bfloat16>::store(dys + off, a);
    if (sub == 0) delta[row] = acc * inv;
}

// ---------------------------------------------------------------------------------------------------
// Backward, dQ: owner = 128 query rows (Q, dO'), loop over 64-key tiles (K_j, V_j).
// TMEM columns: S 0, dP' 64 (128x64 fp32 each), dS[2] 128/160 (bf16 pairs), dQ 192.
// The score columns are released (s_read) as soon as every math thread holds its 32+32 values in
// registers, so the tensor core computes the next tile's scores under this tile's exp math.
// ---------------------------------------------------------------------------------------------------

template <int D>
__global__ void __launch_bounds__(THREADS, Dim<D>::CTAS)
attn_bwd_q_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                     const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_dys,
                     const uint32_t *__restrict__ mask, const int32_t *__restrict__ extra0,
                     const float *__restrict__ delta, __nv_bfloat16 *__restrict__ dq, int S, int H, float scale,
                     float scale_log2, float clamp_log2) {
    constexpr int OWN_BYTES = Dim<D>::OWN_BYTES, T_BYTES = Dim<D>::T_BYTES, OWN_SUB = Dim<D>::OWN_SUB, T_SUB = Dim<D>::T_SUB,
                  TMEM_COLS = Dim<D>::TMEM_COLS;
    extern __shared__ unsigned char smem_raw[];
    const Smem sm = align_smem(smem_raw);
    const uint32_t s_q = sm.base, s_dy = s_q + OWN_BYTES, s_k = s_dy + OWN_BYTES, s_v = s_k + STAGES * T_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(sm.ptr + 2 * OWN_BYTES + 2 * STAGES * T_BYTES);
    const uint32_t bar0 = smem_u32(bars);
    const uint32_t own_full = bar0, acc_full = bar0 + 8, sc_full = bar0 + 16, s_read = bar0 + 24;
    auto p_full = [&](int i) { return bar0 + 32 + i * 8; };
    auto kv_full = [&](int s) { return bar0 + 48 + s * 8; };
    auto kv_empty = [&](int s) { return bar0 + 48 + (STAGES + s) * 8; };
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 6 + 2 * STAGES);

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
        mbar_init(s_read, N_MATH);
        mbar_init(p_full(0), N_MATH);
        mbar_init(p_full(1), N_MATH);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(kv_full(s), 1);
            mbar_init(kv_empty(s), 1);
        }
        mbar_fence_init();
    }
    if (warp == 9) tmem_alloc<TMEM_COLS>(smem_u32(tmem_slot));
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    constexpr uint32_t COL_S = 0, COL_DP = 64, COL_DS = 128, COL_DQ = 192;

    if (warp == 8) {
        if (lane == 0) {
            mbar_expect_tx(own_full, 2 * OWN_BYTES);
            tma_owner<D>(s_q, &map_q, own_full, hh, m0, hn);
            tma_owner<D>(s_dy, &map_dys, own_full, hh, m0, hn);
            for (int j = 0; j < n_tiles; ++j) {
                const int st = j % STAGES;
                mbar_wait(kv_empty(st), ((j / STAGES) & 1) ^ 1);
                mbar_expect_tx(kv_full(st), 2 * T_BYTES);
                tma_other<D>(s_k + st * T_BYTES, &map_k, kv_full(st), hh, j * BN, hn);
                tma_other<D>(s_v + st * T_BYTES, &map_v, kv_full(st), hh, j * BN, hn);
            }
        }
    } else if (warp == 9) {
        constexpr uint32_t id_s = idesc_bf16(BM, BN, 0, 0);   // S = Q K^T, dP' = dO' V^T
        constexpr uint32_t id_a = idesc_bf16(BM, D, 0, 1);    // dQ += dS K   (A from TMEM, K MN-major)
        const uint64_t dq0 = desc_kmajor(s_q, 0), ddy0 = desc_kmajor(s_dy, 0), dk0 = desc_kmajor(s_k, 0),
                       dv0 = desc_kmajor(s_v, 0), dkt0 = desc_mnmajor(s_k, 0, T_SUB);
        auto issue_acc = [&](int j, bool last) {                  // dQ += dS_j K_j
            const int st = j % STAGES;
            mbar_wait(p_full(j & 1), (j >> 1) & 1);
            fence_after_sync();
            if (elect_one()) {
                const uint64_t dkt = dkt0 + (uint64_t)(st * (T_BYTES >> 4));
#pragma unroll
                for (int k = 0; k < BN / 16; ++k)
                    umma_bf16_ts(tmem_base + COL_DQ, tmem_base + COL_DS + (j & 1) * 32 + k * 8, dkt + k * MNMAJOR_K16, id_a,
                                 (j | k) != 0);
                umma_commit(kv_empty(st));
                if (last) umma_commit(acc_full);
            }
            __syncwarp();
        };
        mbar_wait(own_full, 0);
        for (int j = 0; j < n_tiles; ++j) {
            const int st = j % STAGES;
            mbar_wait(kv_full(st), (j / STAGES) & 1);
            if (j > 0) mbar_wait(s_read, (j - 1) & 1);        // tile j-1's scores are in registers
            fence_after_sync();
            if (elect_one()) {
                const uint64_t off = (uint64_t)(st * (T_BYTES >> 4));
                // S and dP' are independent accumulators: alternating their k-steps keeps consecutive MMAs independent
                // (back-to-back MMAs into one accumulator are serialised by the accumulate dependency, ~93 clk each)
#pragma unroll
                for (int k = 0; k < D / 16; ++k) {
                    umma_bf16(tmem_base + COL_S, dq0 + kslice<OWN_SUB>(k), dk0 + off + kslice<T_SUB>(k), id_s, k != 0);
                    umma_bf16(tmem_base + COL_DP, ddy0 + kslice<OWN_SUB>(k), dv0 + off + kslice<T_SUB>(k), id_s, k != 0);
                }
                umma_commit(sc_full);
            }
            __syncwarp();
            if (j > 0) issue_acc(j - 1, false);
        }
        issue_acc(n_tiles - 1, true);
    } else {
        const int quarter = warp & 3, half = warp >> 2;
        const int row = m0 + quarter * 32 + lane;
        const size_t grow = (size_t)b * S + row;
        const uint4 *mrow = reinterpret_cast<const uint4 *>(mask + grow * (S / 32));
        const float ex0 = (float)extra0[grow];
        const float dl = delta[grow];
        const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
        const ClampK ck = make_clamp(scale_log2, clamp_log2);
        uint4 mw = __ldg(mrow), mw_next = mw;
        for (int j = 0; j < n_tiles; ++j) {
            const int bsel = j & 1;
            if (bsel == 0 && j + 2 < n_tiles) mw_next = __ldg(mrow + (j >> 1) + 1);
            mbar_wait(sc_full, j & 1);
            fence_after_sync();
            uint32_t r[32], g[32];
            tmem_ld32_nowait(lane_base + COL_S + half * 32, r);
            tmem_ld32_nowait(lane_base + COL_DP + half * 32, g);
            tmem_ld_wait();
            fence_before_sync();
            mbar_arrive(s_read);
            const int b0 = 16 * bsel + 8 * half;
            const uint32_t w[4] = {mw.x >> b0, mw.y >> b0, mw.z >> b0, mw.w >> b0};
            uint32_t pk[16];
            if (j == 0 && half == 0) bwdq_chunk32<true>(r, g, w, ck, dl, ex0, pk);
            else bwdq_chunk32<false>(r, g, w, ck, dl, ex0, pk);
            tmem_st16(lane_base + COL_DS + bsel * 32 + half * 16, pk);
            tmem_st_wait();
            fence_before_sync();
            mbar_arrive(p_full(bsel));
            if (bsel == 1) mw = mw_next;
        }
        mbar_wait(acc_full, 0);
        fence_after_sync();
        __nv_bfloat16 *dst = dq + (((size_t)hn * S + row) * H + hh) * D + half * (D / 2);
#pragma unroll
        for (int c = 0; c < D / 64; ++c) {
            uint32_t r[32];
            tmem_ld32(lane_base + COL_DQ + half * (D / 2) + c * 32, r);
            store_row32(dst + c * 32, r, scale);
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 9) tmem_dealloc<TMEM_COLS>(tmem_base);
}

// ---------------------------------------------------------------------------------------------------
// Backward, dK / dV: owner = 128 keys (K, V) = one lane-major mask group, loop over 64-row query tiles
// (Q_j, dO'_j) at and below the diagonal.  Works on the transposed tiles S^T = K Q^T, dP'^T = V dO'^T
// so that E^T and dS^T come out with keys on the TMEM lanes, ready to be the A operands of
//   dV += E^T dO'_j   and   dK += dS^T Q_j.
// The 64-row tiles are consumed as two 32-row sub-tiles with two score buffers in TMEM (S^T 32 + dP'^T 32
// columns each, then dV, dK): the tensor core computes the next sub-tile's scores under this one's exp math.
// E^T / dS^T (bf16 pairs) are written back over score columns the same thread has already consumed.
// Per-row quantities of the 64 query rows (mask words of this key group, delta', extra0) are staged in
// shared memory by the producer warp, one slot per pipeline stage.
// ---------------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(THREADS, Dim<D>::CTAS)
attn_bwd_kv_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                      const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_dys,
                      const uint32_t *__restrict__ mask, const int32_t *__restrict__ extra0,
                      const float *__restrict__ delta, __nv_bfloat16 *__restrict__ dk, __nv_bfloat16 *__restrict__ dv,
                      int S, int H, float scale, float scale_log2, float clamp_log2) {
    constexpr int OWN_BYTES = Dim<D>::OWN_BYTES, T_BYTES = Dim<D>::T_BYTES, OWN_SUB = Dim<D>::OWN_SUB, T_SUB = Dim<D>::T_SUB,
                  TMEM_COLS = Dim<D>::TMEM_COLS;
    extern __shared__ unsigned char smem_raw[];
    const Smem sm = align_smem(smem_raw);
    const uint32_t s_k = sm.base, s_v = s_k + OWN_BYTES, s_q = s_v + OWN_BYTES, s_dy = s_q + STAGES * T_BYTES;
    unsigned char *rowq = sm.ptr + 2 * OWN_BYTES + 2 * STAGES * T_BYTES;   // per stage: mask^T [4][64 + 4], delta[64], ex0[64]
    constexpr int MT = BN + 4;                         // padded row of the transposed mask: the 4 words hit distinct banks
    constexpr int ROWQ_BYTES = MT * 16 + BN * 4 + BN * 4;
    uint64_t *bars = reinterpret_cast<uint64_t *>(rowq + STAGES * ROWQ_BYTES);
    const uint32_t bar0 = smem_u32(bars);
    const uint32_t own_full = bar0, acc_full = bar0 + 8;
    auto sc_full = [&](int i) { return bar0 + 16 + i * 8; };
    auto p_full = [&](int i) { return bar0 + 32 + i * 8; };
    auto qd_full = [&](int s) { return bar0 + 48 + s * 8; };
    auto qd_empty = [&](int s) { return bar0 + 48 + (STAGES + s) * 8; };
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 6 + 2 * STAGES);

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
        for (int i = 0; i < 2; ++i) {
            mbar_init(sc_full(i), 1);
            mbar_init(p_full(i), N_MATH);
        }
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(qd_full(s), 1 + 32);             // expect_tx arrive + the 32 producer lanes' row data
            mbar_init(qd_empty(s), 1);
        }
        mbar_fence_init();
    }
    if (warp == 9) tmem_alloc<TMEM_COLS>(smem_u32(tmem_slot));
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    // score buffers b = 0, 1 (one per 32-row sub-tile in flight): S^T at 64 b, dP'^T at 64 b + 32
    constexpr uint32_t COL_SC = 0, COL_DV = 128, COL_DK = 128 + D;
    constexpr int SUBN = 32;

    if (warp == 8) {
        if (lane == 0) {
            mbar_expect_tx(own_full, 2 * OWN_BYTES);
            tma_owner<D>(s_k, &map_k, own_full, hh, n0, hn);
            tma_owner<D>(s_v, &map_v, own_full, hh, n0, hn);
        }
        for (int j = 0; j < n_tiles; ++j) {
            const int st = j % STAGES;
            const int r0 = (j0 + j) * BN;
            mbar_wait(qd_empty(st), ((j / STAGES) & 1) ^ 1);
            if (lane == 0) {
                mbar_expect_tx(qd_full(st), 2 * T_BYTES);
                tma_other<D>(s_q + st * T_BYTES, &map_q, qd_full(st), hh, r0, hn);
                tma_other<D>(s_dy + st * T_BYTES, &map_dys, qd_full(st), hh, r0, hn);
            }
            unsigned char *slot = rowq + st * ROWQ_BYTES;
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int rr = lane + 32 * u;
                const size_t gr = head + r0 + rr;
                const uint4 mw = __ldg(reinterpret_cast<const uint4 *>(mask + gr * words) + kt);
                uint32_t *mt = reinterpret_cast<uint32_t *>(slot);        // transposed: [word t][row], rows padded to MT
                mt[rr] = mw.x;
                mt[MT + rr] = mw.y;
                mt[2 * MT + rr] = mw.z;
                mt[3 * MT + rr] = mw.w;
                reinterpret_cast<float *>(slot + MT * 16)[rr] = delta[gr];
                reinterpret_cast<float *>(slot + MT * 16 + BN * 4)[rr] = (float)extra0[gr];
            }
            mbar_arrive(qd_full(st));
        }
    } else if (warp == 9) {
        // Sub-tiles t = 2 j + h of 32 query rows: scores of sub-tile t+1 are issued before the math warps have
        // finished sub-tile t (two score buffers), the accumulating MMAs of t follow once its E^T / dS^T are written.
        constexpr uint32_t id_s = idesc_bf16(BM, SUBN, 0, 0); // S^T = K Q^T, dP'^T = V dO'^T   (N = 32 query rows)
        constexpr uint32_t id_a = idesc_bf16(BM, D, 0, 1);    // dV += E^T dO', dK += dS^T Q (B MN-major, K = 32)
        const uint64_t dk0 = desc_kmajor(s_k, 0), dv0 = desc_kmajor(s_v, 0), dq0 = desc_kmajor(s_q, 0),
                       ddy0 = desc_kmajor(s_dy, 0), dqt0 = desc_mnmajor(s_q, 0, T_SUB),
                       ddyt0 = desc_mnmajor(s_dy, 0, T_SUB);
        const int n_sub = 2 * n_tiles;
        auto issue_scores = [&](int t) {
            const int j = t >> 1, h = t & 1, st = j % STAGES;
            if (h == 0) mbar_wait(qd_full(st), (j / STAGES) & 1);
            fence_after_sync();
            if (elect_one()) {
                // rows 32 h .. 32 h + 31 of the stage's 64-row tiles: + 32 rows * 128 B inside every 64-wide sub-tile
                const uint64_t off = (uint64_t)((st * T_BYTES + h * SUBN * 128) >> 4);
                const uint32_t col = tmem_base + COL_SC + (t & 1) * 64;
                // S^T and dP'^T are independent accumulators: alternate their k-steps (see the dQ kernel)
#pragma unroll
                for (int k = 0; k < D / 16; ++k) {
                    umma_bf16(col, dk0 + kslice<OWN_SUB>(k), dq0 + off + kslice<T_SUB>(k), id_s, k != 0);
                    umma_bf16(col + 32, dv0 + kslice<OWN_SUB>(k), ddy0 + off + kslice<T_SUB>(k), id_s, k != 0);
                }
                umma_commit(sc_full(t & 1));
            }
            __syncwarp();
        };
        mbar_wait(own_full, 0);
        issue_scores(0);
        for (int t = 0; t < n_sub; ++t) {
            if (t + 1 < n_sub) issue_scores(t + 1);   // its buffer was last read by the accumulating MMAs of t - 1 (in order)
            const int j = t >> 1, h = t & 1, st = j % STAGES;
            mbar_wait(p_full(t & 1), (t >> 1) & 1);
            fence_after_sync();
            if (elect_one()) {
                const uint64_t off = (uint64_t)((st * T_BYTES) >> 4);
                const uint32_t col = tmem_base + COL_SC + (t & 1) * 64;
                // query rows 16 k .. 16 k + 15 of the sub-tile sit in columns 16 k .. 16 k + 7 (the writer's own half)
#pragma unroll
                for (int k = 0; k < SUBN / 16; ++k) {   // dV and dK alternate as well
                    umma_bf16_ts(tmem_base + COL_DV, col + k * 16, ddyt0 + off + (2 * h + k) * MNMAJOR_K16, id_a, (t | k) != 0);
                    umma_bf16_ts(tmem_base + COL_DK, col + 32 + k * 16, dqt0 + off + (2 * h + k) * MNMAJOR_K16, id_a, (t | k) != 0);
                }
                if (h == 1) umma_commit(qd_empty(st));
                if (t + 1 == n_sub) umma_commit(acc_full);
            }
            __syncwarp();
        }
    } else {
        // thread = (key n0 + kk = TMEM lane kk, column half): lane-major word kk & 3 of the row's group kt, bit kk >> 2
        const int quarter = warp & 3, half = warp >> 2;
        const int kk = quarter * 32 + lane;
        const int wsel = kk & 3;
        const uint32_t bitmask = 1u << (kk >> 2);
        const bool key0 = (n0 + kk) == 0;
        const ClampK ck = make_clamp(scale_log2, clamp_log2);
        const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
        for (int t = 0; t < 2 * n_tiles; ++t) {
            const int j = t >> 1, h = t & 1, st = j % STAGES;
            const unsigned char *slot = rowq + st * ROWQ_BYTES;
            const uint32_t *mrow = reinterpret_cast<const uint32_t *>(slot) + wsel * MT;   // this key's word of every row
            const float *s_delta = reinterpret_cast<const float *>(slot + MT * 16);
            const float *s_ex0 = s_delta + BN;
            if (h == 0) mbar_wait(qd_full(st), (j / STAGES) & 1);   // the producer lanes' row data of this stage
            mbar_wait(sc_full(t & 1), (t >> 1) & 1);
            fence_after_sync();
            const uint32_t col = lane_base + COL_SC + (t & 1) * 64 + half * 16;   // this thread's 16 query rows
            const int c0 = h * SUBN + half * 16;                                   // ... = rows c0 .. c0 + 15 of the tile
            uint32_t r[16], g[16];
            tmem_ld16_nowait(col, r);
            tmem_ld16_nowait(col + 32, g);
            tmem_ld_wait();
            uint32_t pe[8], pd[8];
            if (key0) bwdkv_chunk16<true>(r, g, mrow, s_delta, s_ex0, c0, bitmask, ck, pe, pd);
            else bwdkv_chunk16<false>(r, g, mrow, s_delta, s_ex0, c0, bitmask, ck, pe, pd);
            tmem_st8(col, pe);         // E^T over the S^T columns this thread has consumed
            tmem_st8(col + 32, pd);    // dS^T over its dP'^T columns
            tmem_st_wait();
            fence_before_sync();
            mbar_arrive(p_full(t & 1));
        }
        mbar_wait(acc_full, 0);
        fence_after_sync();
        const size_t off = (((size_t)hn * S + n0 + kk) * H + hh) * D + half * (D / 2);
#pragma unroll
        for (int c = 0; c < D / 64; ++c) {
            uint32_t r[32];
            tmem_ld32(lane_base + COL_DV + half * (D / 2) + c * 32, r);
            store_row32(dv + off + c * 32, r, 1.0f);
            tmem_ld32(lane_base + COL_DK + half * (D / 2) + c * 32, r);
            store_row32(dk + off + c * 32, r, scale);
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 9) tmem_dealloc<TMEM_COLS>(tmem_base);
}

// ---- host -----------------------------------------------------------------------------------------
// 4-D bf16 tensor map over a [N, S, H, D] tensor (H = 1: head-major [B, S, D]); box = 64 rows x 64 elements of one head
static int make_map(CUtensorMap *map, const void *base, int N, int S, int H, int D) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return fail(SPT_ERR_CUDA, "sparse_attn: cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[4] = {(cuuint64_t)D, (cuuint64_t)H, (cuuint64_t)S, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)D * 2, (cuuint64_t)H * D * 2, (cuuint64_t)S * H * D * 2};
    cuuint32_t box[4] = {64, 1, (cuuint32_t)BN, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(SPT_ERR_CUDA, "sparse_attn: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return SPT_OK;
}

template <int D>
static int launch_fwd(const CUtensorMap &mq, const CUtensorMap &mk, const CUtensorMap &mv, const uint32_t *mask,
                      const int32_t *extra0, __nv_bfloat16 *y, float *zsum, int B, int S, int H, float scale, float clamp,
                      cudaStream_t st) {
    cudaFuncSetAttribute(attn_fwd_tc_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, Dim<D>::FWD_SMEM);
    attn_fwd_tc_kernel<D><<<dim3(S / BM, B), THREADS, Dim<D>::FWD_SMEM, st>>>(mq, mk, mv, mask, extra0, y, zsum, S, H,
                                                                              scale * LOG2E, clamp * LOG2E);
    return after_launch("attn_fwd_tc_kernel");
}

template <int D>
static int launch_prep(const __nv_bfloat16 *y, const __nv_bfloat16 *grad_y, const float *zsum, float *delta,
                       __nv_bfloat16 *dys, int B, int S, int H, cudaStream_t st) {
    const int64_t rows = (int64_t)B * S;
    constexpr int RPB = 256 / (D / 8);
    attn_bwd_prep_kernel<D><<<(unsigned)((rows + RPB - 1) / RPB), 256, 0, st>>>(grad_y, y, zsum, delta, dys, rows, S, H);
    return after_launch("attn_bwd_prep_kernel");
}

template <int D>
static int launch_bwd(const CUtensorMap &mq, const CUtensorMap &mk, const CUtensorMap &mv, const CUtensorMap &md,
                      const uint32_t *mask, const int32_t *extra0, const float *delta, __nv_bfloat16 *gq,
                      __nv_bfloat16 *gk, __nv_bfloat16 *gv, int B, int S, int H, float scale, float clamp, cudaStream_t st) {
    cudaFuncSetAttribute(attn_bwd_kv_tc_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, Dim<D>::BWD_SMEM);
    attn_bwd_kv_tc_kernel<D><<<dim3(S / BM, B), THREADS, Dim<D>::BWD_SMEM, st>>>(
        mq, mk, mv, md, mask, extra0, delta, gk, gv, S, H, scale, scale * LOG2E, clamp * LOG2E);
    SPT_LAUNCH_CHECK("attn_bwd_kv_tc_kernel");
    cudaFuncSetAttribute(attn_bwd_q_tc_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, Dim<D>::BWD_SMEM);
    attn_bwd_q_tc_kernel<D><<<dim3(S / BM, B), THREADS, Dim<D>::BWD_SMEM, st>>>(
        mq, mk, mv, md, mask, extra0, delta, gq, S, H, scale, scale * LOG2E, clamp * LOG2E);
    SPT_LAUNCH_CHECK("attn_bwd_q_tc_kernel");
    return SPT_OK;
}

}  // namespace attn_tc
}  // namespace spt

using namespace spt;

static int check_attn_args(const char *what, int B, int S, int d, int H, int dtype) {
    if (dtype != SPT_BF16) return fail(SPT_ERR_UNSUPPORTED, "%s: only bf16 is supported on the fused path", what);
    if (d != 64 && d != 128) return fail(SPT_ERR_UNSUPPORTED, "%s: head dim %d not supported (64 or 128)", what, d);
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
    if ((rc = attn_tc::make_map(&mq, q, B / H, S, H, d)) != SPT_OK) return rc;
    if ((rc = attn_tc::make_map(&mk, k, B / H, S, H, d)) != SPT_OK) return rc;
    if ((rc = attn_tc::make_map(&mv, v, B / H, S, H, d)) != SPT_OK) return rc;
    if (d == 64) return attn_tc::launch_fwd<64>(mq, mk, mv, mask, extra0, (bf *)y, zsum, B, S, H, scale, clamp, as_stream(stream));
    return attn_tc::launch_fwd<128>(mq, mk, mv, mask, extra0, (bf *)y, zsum, B, S, H, scale, clamp, as_stream(stream));
}

// workspace: delta' [B, S] fp32, then dO' (bf16, same shape as grad_y; sized for the largest head dim)
extern "C" size_t spt_sparse_attn_bwd_workspace_bytes(int B, int S) {
    return (size_t)B * S * sizeof(float) + (size_t)B * S * 128 * 2;
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
    float *delta = (float *)workspace;
    bf *dys = (bf *)((char *)workspace + (size_t)B * S * sizeof(float));
    // the row kernel goes first: besides producing dO' and delta' it is a runtime-API launch, which binds the
    // device's primary context to this (autograd worker) thread before the driver-API tensor-map encoder runs
    rc = d == 64 ? attn_tc::launch_prep<64>((const bf *)y, (const bf *)grad_y, zsum, delta, dys, B, S, H, as_stream(stream))
                 : attn_tc::launch_prep<128>((const bf *)y, (const bf *)grad_y, zsum, delta, dys, B, S, H, as_stream(stream));
    if (rc != SPT_OK) return rc;
    CUtensorMap mq, mk, mv, md;
    if ((rc = attn_tc::make_map(&mq, q, B / H, S, H, d)) != SPT_OK) return rc;
    if ((rc = attn_tc::make_map(&mk, k, B / H, S, H, d)) != SPT_OK) return rc;
    if ((rc = attn_tc::make_map(&mv, v, B / H, S, H, d)) != SPT_OK) return rc;
    if ((rc = attn_tc::make_map(&md, dys, B / H, S, H, d)) != SPT_OK) return rc;
    if (d == 64)
        return attn_tc::launch_bwd<64>(mq, mk, mv, md, mask, extra0, delta, (bf *)grad_q, (bf *)grad_k, (bf *)grad_v, B, S,
                                       H, scale, clamp, as_stream(stream));
    return attn_tc::launch_bwd<128>(mq, mk, mv, md, mask, extra0, delta, (bf *)grad_q, (bf *)grad_k, (bf *)grad_v, B, S, H,
                                    scale, clamp, as_stream(stream));
}
