// Shared pieces of the fused sparse-attention kernels (attn_tc.cu: 128 x 64 tiles, two CTAs per SM;
// attn_tc128.cu: 128 x 128 tiles, one CTA per SM): tile geometry, TMA helpers, the packed element math,
// optional phase timers.
#pragma once
#include <cstdlib>
#include <cstring>

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
// ---- element math --------------------------------------------------------------------------------
// Per score element the three kernels need  e = w * exp(clamp(scale s, -10, 10))  (and, in the backward,
// ds = e (dp - delta') [|scale s| <= 10]).  The first version spent ~11-15 instructions per element on it and was
// issue-bound; this one spends ~5-7:
//   * fp32 pairs are processed with the packed sm_100 instructions (mul.f32x2 / add.f32x2 / fma.rn.f32x2 = FMUL2 /
//     FADD2 / FFMA2: two elements per issue slot), on the register pairs tcgen05.ld delivers;
//   * the clamp is resolved per warp and 32-column chunk: one FMNMX3 per two elements tracks max |s|; only when some
//     lane of the warp holds a score beyond the clamp (a vote) does the chunk take the exact path with min/max and the
//     zero-gradient indicator.  Otherwise clamp(x) = x and the indicator is 1 — bit-identical results either way;
//   * the selection mask is applied to the PACKED bf16 pair with one LOP3: the mask bits of four columns are moved to
//     the sign bits of the four bytes of a register (one shift per four elements), and PRMT's sign-replicate mode
//     expands two of them into a 0xFFFF / 0x0000 pair mask (one PRMT per two elements);
//   * a fraction of the exp2 (the XU pipe, 16 / clk / SM, is the floor once the rest is this cheap) is evaluated on the
//     FMA pipe: Cody-Waite range reduction + a cubic (max relative error 7.5e-5, bf16 keeps 3.9e-3).
__device__ __forceinline__ uint64_t pk2(float lo, float hi) {
    uint64_t d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
    return d;
}
__device__ __forceinline__ void up2(uint64_t v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("mul.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}
__device__ __forceinline__ float max3abs(float a, float b, float c) {   // max(a, |b|, |c|): one FMNMX3
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(fabsf(b)), "f"(fabsf(c)));
    return d;
}

struct MathK {      // arg = s * c (log2 units);  |scale s| <= clamp  <=>  |s| <= thr  <=>  |arg| <= L
    float c, thr, L;
};
__device__ __forceinline__ MathK make_math(float scale_log2, float clamp_log2) {
    return {scale_log2, clamp_log2 / scale_log2, clamp_log2};
}
// does any lane of the warp hold a raw score beyond the clamp among its N values?  (warp-uniform result)
template <int N>
__device__ __forceinline__ bool warp_needs_clamp(const uint32_t (&r)[N], float thr) {
    float mx = 0.0f;
#pragma unroll
    for (int i = 0; i < N; i += 2) mx = max3abs(mx, __uint_as_float(r[i]), __uint_as_float(r[i + 1]));
    return __any_sync(0xffffffffu, mx > thr);
}
// exp2 of a pair.  POLY: on the FMA pipe.  r = a + 1.5 * 2^23 holds round(a) in its low mantissa bits; f = a - round(a)
// in [-0.5, 0.5]; 2^f by a cubic; 2^round(a) by adding round(a) << 23 to the exponent field.
template <bool POLY>
__device__ __forceinline__ void ex2_pair(uint64_t a, float &e0, float &e1) {
    if constexpr (!POLY) {
        float a0, a1;
        up2(a, a0, a1);
        e0 = ex2(a0);
        e1 = ex2(a1);
    } else {
        const uint64_t magic = pk2(12582912.0f, 12582912.0f), nmagic = pk2(-12582912.0f, -12582912.0f);
        const uint64_t r = add2(a, magic);
        const uint64_t f = fma2(add2(r, nmagic), pk2(-1.0f, -1.0f), a);
        uint64_t p = fma2(pk2(0.0551716685295105f, 0.0551716685295105f), f, pk2(0.2426111251115799f, 0.2426111251115799f));
        p = fma2(p, f, pk2(0.6932609677314758f, 0.6932609677314758f));
        p = fma2(p, f, pk2(0.9999280571937561f, 0.9999280571937561f));
        float r0, r1, p0, p1;
        up2(r, r0, r1);
        up2(p, p0, p1);
        e0 = __int_as_float(__float_as_int(p0) + (__float_as_int(r0) << 23));
        e1 = __int_as_float(__float_as_int(p1) + (__float_as_int(r1) << 23));
    }
}
// which pairs of a chunk go to the FMA pipe: pair index p (0 .. N/2-1); POLY_MOD = 0: none, else every POLY_MOD-th
#ifndef SPT_ATTN_POLY_MOD
#define SPT_ATTN_POLY_MOD 0
#endif
__device__ __forceinline__ constexpr bool poly_pair(int p) {
    constexpr int mod = SPT_ATTN_POLY_MOD > 0 ? SPT_ATTN_POLY_MOD : 1;
    return SPT_ATTN_POLY_MOD > 0 && (p % mod) == mod - 1;
}

// e of one pair of raw scores: exp2(clamp(s c)); CLAMP: exact path, also returns the per-element gradient indicators
template <bool CLAMP, bool POLY>
__device__ __forceinline__ void exp_pair(float s0, float s1, const MathK mk, float &e0, float &e1, bool &in0, bool &in1) {
    uint64_t a = mul2(pk2(s0, s1), pk2(mk.c, mk.c));
    if constexpr (CLAMP) {
        float a0, a1;
        up2(a, a0, a1);
        in0 = fabsf(s0) <= mk.thr;
        in1 = fabsf(s1) <= mk.thr;
        a = pk2(fminf(fmaxf(a0, -mk.L), mk.L), fminf(fmaxf(a1, -mk.L), mk.L));
    }
    ex2_pair<POLY>(a, e0, e1);
}

// The 32 score columns [32 half, 32 half + 32) of key tile j of one query row: gathers, from the row's four lane-major
// mask words of key group j >> 1, the byte (index 2 (j & 1) + half) that covers them.  Result X: byte t bit n <=> column
// 4 n + t of the chunk.
__device__ __forceinline__ uint32_t chunk_mask_bytes(const uint4 mw, int byte_idx) {
    const uint32_t sel = (uint32_t)(((4 + byte_idx) << 4) | byte_idx);
    return prmt(prmt(mw.x, mw.y, sel), prmt(mw.z, mw.w, sel), 0x5410u);
}
// pair masks of columns (4 n, 4 n + 1) and (4 n + 2, 4 n + 3) of a chunk
__device__ __forceinline__ void pair_masks(uint32_t X, int n, uint32_t &m01, uint32_t &m23) {
    const uint32_t Y = X << (7 - n);
    m01 = prmt(Y, 0u, 0x9988u);
    m23 = prmt(Y, 0u, 0xBBAAu);
}
__device__ __forceinline__ uint64_t unpack_bf16x2(uint32_t p) { return pk2(__uint_as_float(p << 16), __uint_as_float(p & 0xffff0000u)); }

// Forward: 32 score columns of one row -> 16 masked packed bf16 pairs of e = w * exp(clamp(scale s)); the row sum
// accumulates the bf16-rounded values (exactly the weights the P V product uses).  FIRST: column 0 is key 0, which also
// carries the row's zero-padding multiplicity ex0 (mult0 = bit + ex0 when > 0).
template <bool FIRST, bool CLAMP>
__device__ __forceinline__ void fwd_chunk32(const uint32_t (&r)[32], uint32_t X, const MathK mk, float ex0, uint64_t &sum2,
                                            uint32_t (&pk)[16]) {
    float mult0 = 1.0f;
    if (FIRST) {
        mult0 = (float)(X & 1u) + ex0;
        if (mult0 > 0.0f) X |= 1u;
        else mult0 = 1.0f;
    }
#pragma unroll
    for (int n = 0; n < 8; ++n) {
        uint32_t m01, m23;
        pair_masks(X, n, m01, m23);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int i = 4 * n + 2 * h;
            float e0, e1;
            bool in0, in1;
            if (poly_pair(i >> 1)) exp_pair<CLAMP, true>(__uint_as_float(r[i]), __uint_as_float(r[i + 1]), mk, e0, e1, in0, in1);
            else exp_pair<CLAMP, false>(__uint_as_float(r[i]), __uint_as_float(r[i + 1]), mk, e0, e1, in0, in1);
            if (FIRST && i == 0) e0 *= mult0;
            const uint32_t p = pack_bf16(e0, e1) & (h ? m23 : m01);
            sum2 = add2(sum2, unpack_bf16x2(p));
            pk[i >> 1] = p;
        }
    }
}

// dQ kernel: 32 columns (keys) of one query row -> 16 masked packed pairs of ds = e (dp - delta') [unclamped]; ndelta = -delta'.
template <bool FIRST, bool CLAMP>
__device__ __forceinline__ void bwdq_chunk32(const uint32_t (&r)[32], const uint32_t (&g)[32], uint32_t X, const MathK mk,
                                             float ndelta, float ex0, uint32_t (&pk)[16]) {
    float mult0 = 1.0f;
    if (FIRST) {
        mult0 = (float)(X & 1u) + ex0;
        if (mult0 > 0.0f) X |= 1u;
        else mult0 = 1.0f;
    }
    const uint64_t nd2 = pk2(ndelta, ndelta);
#pragma unroll
    for (int n = 0; n < 8; ++n) {
        uint32_t m01, m23;
        pair_masks(X, n, m01, m23);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int i = 4 * n + 2 * h;
            float e0, e1;
            bool in0 = true, in1 = true;
            if (poly_pair(i >> 1)) exp_pair<CLAMP, true>(__uint_as_float(r[i]), __uint_as_float(r[i + 1]), mk, e0, e1, in0, in1);
            else exp_pair<CLAMP, false>(__uint_as_float(r[i]), __uint_as_float(r[i + 1]), mk, e0, e1, in0, in1);
            if (FIRST && i == 0) e0 *= mult0;
            float d0, d1;
            up2(mul2(pk2(e0, e1), add2(pk2(__uint_as_float(g[i]), __uint_as_float(g[i + 1])), nd2)), d0, d1);
            if (CLAMP) {
                d0 = in0 ? d0 : 0.0f;
                d1 = in1 ? d1 : 0.0f;
            }
            pk[i >> 1] = pack_bf16(d0, d1) & (h ? m23 : m01);
        }
    }
}

// dK/dV kernel: 16 columns (query rows c0 .. c0+15) of one key.  mrow = this key's lane-major word of every row of the
// stage ([64] in shared memory), shl moves the key's bit to bit 31; s_ndelta = -delta' of the rows.
// KEY0: this thread is key 0 (adds the rows' zero-padding multiplicity s_ex0).
// r / g hold NR values of which [OFF, OFF + 16) are processed; c0 = tile row of r[OFF].
template <bool KEY0, bool CLAMP, int OFF = 0, int NR = 16, typename EX0 = float>
__device__ __forceinline__ void bwdkv_chunk16(const uint32_t (&rr)[NR], const uint32_t (&gg)[NR], const uint32_t *mrow,
                                              const float *s_ndelta, const EX0 *s_ex0, int c0, int shl, const MathK mk,
                                              uint32_t (&pe)[8], uint32_t (&pd)[8]) {
    uint32_t r[16], g[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        r[i] = rr[OFF + i];
        g[i] = gg[OFF + i];
    }
#pragma unroll
    for (int q4 = 0; q4 < 16; q4 += 4) {
        const uint4 mw = *reinterpret_cast<const uint4 *>(mrow + c0 + q4);
        const float4 nd = *reinterpret_cast<const float4 *>(s_ndelta + c0 + q4);
        const uint32_t mwv[4] = {mw.x << shl, mw.y << shl, mw.z << shl, mw.w << shl};
        const float ndv[4] = {nd.x, nd.y, nd.z, nd.w};
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int i = q4 + 2 * h;
            float e0, e1;
            bool in0 = true, in1 = true;
            if (poly_pair(i >> 1)) exp_pair<CLAMP, true>(__uint_as_float(r[i]), __uint_as_float(r[i + 1]), mk, e0, e1, in0, in1);
            else exp_pair<CLAMP, false>(__uint_as_float(r[i]), __uint_as_float(r[i + 1]), mk, e0, e1, in0, in1);
            uint32_t ma = mwv[2 * h], mb = mwv[2 * h + 1];
            if (KEY0) {     // key 0: weight = bit + zero-padding multiplicity of the row
                const float w0 = (float)(ma >> 31) + (float)s_ex0[c0 + i], w1 = (float)(mb >> 31) + (float)s_ex0[c0 + i + 1];
                e0 *= w0;
                e1 *= w1;
                ma = w0 > 0.0f ? 0x80000000u : 0u;
                mb = w1 > 0.0f ? 0x80000000u : 0u;
            }
            const uint32_t m = prmt(ma, mb, 0xFFBBu);      // sign of ma -> low half, sign of mb -> high half
            float d0, d1;
            up2(mul2(pk2(e0, e1), add2(pk2(__uint_as_float(g[i]), __uint_as_float(g[i + 1])), pk2(ndv[2 * h], ndv[2 * h + 1]))),
                d0, d1);
            if (CLAMP) {
                d0 = in0 ? d0 : 0.0f;
                d1 = in1 ? d1 : 0.0f;
            }
            pe[i >> 1] = pack_bf16(e0, e1) & m;
            pd[i >> 1] = pack_bf16(d0, d1) & m;
        }
    }
}

// ---- optional in-kernel phase timers (build with -DSPT_ATTN_PROF; read with spt_debug_attn_prof) ----------------------
// One math thread (warp 0, lane 0) and the MMA-issuer lane of every CTA accumulate clock64() deltas per phase and add them
// to g_prof[kernel][slot] at exit.  Slots 0-7: math thread (0 wait scores, 1 tcgen05.ld, 2 element math, 3 tcgen05.st +
// arrive, 4 iterations, 5 whole loop); 8-15: issuer (8 wait operands, 9 wait math, 10 issue, 11 whole loop, 12 iterations).
#ifdef SPT_ATTN_PROF
static __device__ unsigned long long g_prof[3][16];   // one copy per translation unit
struct Prof {
    long long t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long last;
    __device__ __forceinline__ void start() { last = clock64(); }
    __device__ __forceinline__ void lap(int slot) {
        const long long now = clock64();
        t[slot] += now - last;
        last = now;
    }
    __device__ __forceinline__ void flush(int kernel, int base, bool on) {
        if (on)
            for (int i = 0; i < 8; ++i) atomicAdd(&g_prof[kernel][base + i], (unsigned long long)t[i]);
    }
};
#define PROF(...) __VA_ARGS__
#else
#define PROF(...)
#endif

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

}  // namespace attn_tc

// 128 x 128-tile kernels (attn_tc128.cu), head dim 64
namespace attn_tc128 {
int launch_bwd_kv128(const CUtensorMap &mq, const CUtensorMap &mk, const CUtensorMap &mv, const CUtensorMap &md,
                     const uint32_t *mask, const int32_t *extra0, const float *ndelta, __nv_bfloat16 *gk,
                     __nv_bfloat16 *gv, int B, int S, int H, float scale, float clamp, cudaStream_t st);
int launch_bwd_q128(const CUtensorMap &mq, const CUtensorMap &mk, const CUtensorMap &mv, const CUtensorMap &md,
                    const uint32_t *mask, const int32_t *extra0, const float *ndelta, __nv_bfloat16 *gq, int B, int S,
                    int H, float scale, float clamp, cudaStream_t st);
int launch_bwd_fused128(const CUtensorMap &mq, const CUtensorMap &mk, const CUtensorMap &mv, const CUtensorMap &md,
                        const uint32_t *mask, const int32_t *extra0, const float *ndelta, float *dq_acc,
                        __nv_bfloat16 *gq, __nv_bfloat16 *gk, __nv_bfloat16 *gv, int B, int S, int H, float scale,
                        float clamp, cudaStream_t st);
int launch_fwd128(const CUtensorMap &mq, const CUtensorMap &mk, const CUtensorMap &mv, const uint32_t *mask,
                  const int32_t *extra0, __nv_bfloat16 *y, float *zsum, int B, int S, int H, float scale, float clamp,
                  int y_transposed, cudaStream_t st);
int read_prof(unsigned long long *out48, int reset);
}  // namespace attn_tc128
}  // namespace spt
