// bf16 sddmm / spmm / transposed spmm on the CSR (CSC) pattern with mma.sync (m16n8k16, fp32 accumulation) fed straight
// from registers — reference call sites extension/sddmm.cpp:27-69, extension/spmm.cpp:27-69 (cuSPARSE generic SDDMM / SpMM).
//
// The warp-per-row SIMT kernels of csr.cu are bound by the LSU pipe, not by arithmetic: per gathered (row, column) pair
// they issue one 128-byte row read (one L1 wavefront: the floor) plus ~1.75 SHUFFLES (reduction across the 8 lanes that
// share a row, routing of indices and results) — and shuffles go through the same LSU data pipe (ncu: 2.8 LSU wavefronts
// per pair, 68 % of the pipe's peak, 84 % issue-active).  Staging the rows in shared memory for ldmatrix was measured
// SLOWER (the data then crosses the unified L1 / shared-memory path three times).  What works is to load the rows in the
// register layout the tensor core wants, using the freedom to PERMUTE the contraction index:
//   sddmm : lane (g, t) = (lane / 4, lane % 4) reads a contiguous quarter of gathered key rows g and g + 8 with 256-bit
//           loads (the four lanes of a row cover its 128 bytes in ONE wavefront) and uses register pair (2 ks, 2 ks + 1)
//           as the A fragment of k-step ks: contraction position (t, ks, j) <-> element 16 t + 4 ks + j.  B is the
//           query row under the same permutation, replicated in every column, so every lane ends up holding the score
//           of row g in its accumulator — no reduction, no routing.
//   spmm  : contraction = the 16 entries of a group.  Lane (g, t) loads features 8 g .. 8 g + 7 of the rows of entries
//           2 t, 2 t + 1, 2 t + 8, 2 t + 9 (128-bit loads: the eight g-lanes cover one row per wavefront) and PRMTs the
//           pairs (entry 2 t, entry 2 t + 1) of one feature into B fragments; feature tile j, column n <-> feature
//           8 n + j, so that lane t of the result owns 16 contiguous features.  Row 0 / row 1 of A are the high / low
//           bf16 halves of the fp32 weights (w = hi + lo keeps ~16 bits: the stage API takes fp32 values).
// ~1.2 - 1.5 LSU wavefronts and < 3 instructions per pair.  Summation order is fixed (entry order inside a warp, warps
// in index order) => bit-reproducible like the SIMT kernels.
#include <cstdlib>

#include "common.cuh"

namespace spt {
namespace csr_mma {

using bf16 = __nv_bfloat16;

__device__ __forceinline__ void mma16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldg256(const void *p, uint32_t (&r)[8]) {       // 32-byte aligned
    asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(p));
}
__device__ __forceinline__ uint4 ldg128(const void *p) { return __ldg(reinterpret_cast<const uint4 *>(p)); }

constexpr int WARPS = 8;

// ---- sddmm ----------------------------------------------------------------------------------------------------------
// a batch = 32 entries = two m16 groups; lane (g, t) holds quarter t of the key rows of entries g, g + 8, g + 16, g + 24
template <int D>
__global__ void __launch_bounds__(WARPS * 32)
sddmm_mma_kernel(const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices, const bf16 *__restrict__ q,
                 const bf16 *__restrict__ k, float *__restrict__ values, int B, int S, int64_t nnz, float scale, float clamp) {
    constexpr int NV = D / 64;                              // 256-bit loads per row quarter
    constexpr int KS = D / 16;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row_id = (int64_t)blockIdx.x * WARPS + warp;
    if (row_id >= (int64_t)B * S) return;
    const int b = (int)(row_id / S), r = (int)(row_id % S);
    const int e0 = indptr[r], e1 = indptr[r + 1];
    if (e0 >= e1) return;
    const int g = lane >> 2, t = lane & 3;
    const int32_t *ip = indices + (size_t)b * nnz;
    float *vp = values + (size_t)b * nnz;
    const bf16 *kq = k + (size_t)b * S * D + t * (D / 4);   // this lane's quarter of every key row
    uint32_t bq[2 * KS];                                    // the query row's quarter t: B fragments (b0, b1) of k-step ks = words 2 ks, 2 ks + 1
    {
        const bf16 *qp = q + ((size_t)b * S + r) * D + t * (D / 4);
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            uint32_t tmp[8];
            ldg256(qp + 16 * v, tmp);
#pragma unroll
            for (int i = 0; i < 8; ++i) bq[8 * v + i] = tmp[i];
        }
    }
    auto load_idx = [&](int base, int (&ix)[4]) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int e = base + g + 8 * u;
            ix[u] = e < e1 ? __ldg(ip + e) : 0;
        }
    };
    int ix[4];
    load_idx(e0, ix);
    for (int base = e0; base < e1; base += 32) {
        uint32_t rows[4][2 * KS];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                uint32_t tmp[8];
                ldg256(kq + (size_t)ix[u] * D + 16 * v, tmp);
#pragma unroll
                for (int i = 0; i < 8; ++i) rows[u][8 * v + i] = tmp[i];
            }
        }
        if (base + 32 < e1) load_idx(base + 32, ix);       // the next batch's indices travel under this batch's rows
#pragma unroll
        for (int gi = 0; gi < 2; ++gi) {
            float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
            for (int ks = 0; ks < KS; ++ks)
                mma16816(acc, rows[2 * gi][2 * ks], rows[2 * gi + 1][2 * ks], rows[2 * gi][2 * ks + 1], rows[2 * gi + 1][2 * ks + 1],
                         bq[2 * ks], bq[2 * ks + 1]);
            if (t == 0) {                                   // acc[0] = score of entry g of the group, acc[2] = entry g + 8
                const int e = base + 16 * gi + g;
                float v0 = acc[0] * scale, v1 = acc[2] * scale;
                if (clamp > 0.0f) {
                    v0 = fminf(fmaxf(v0, -clamp), clamp);
                    v1 = fminf(fmaxf(v1, -clamp), clamp);
                }
                if (e < e1) vp[e] = v0;
                if (e + 8 < e1) vp[e + 8] = v1;
            }
        }
    }
}

// ---- spmm / transposed spmm -------------------------------------------------------------------------------------------
// Per-lane view of a batch of 32 entries (two groups): entries 2 t, 2 t + 1, 2 t + 8, 2 t + 9 of each group.
struct Batch {
    int row[8];         // source rows of the lane's 8 entries
    float w[8];         // their weights
};

// accumulate one batch: acc[j] = m16n8 accumulator of feature tile j (row 0: hi weights, row 1: lo weights)
template <int D>
__device__ __forceinline__ void spmm_batch(float (&acc)[D / 8][4], const bf16 *__restrict__ xg /* x + 8 g */, const Batch &bt,
                                           int g) {
    constexpr int NC = D / 64;                              // 16-byte feature chunks per lane and row
    uint4 rows[8][NC];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int c = 0; c < NC; ++c) rows[i][c] = ldg128(xg + (size_t)bt.row[i] * D + 64 * c);
#pragma unroll
    for (int gi = 0; gi < 2; ++gi) {
        // A fragments: row 0 (lanes g = 0) the high bf16 halves of the weights, row 1 (g = 1) the low halves
        uint32_t a0 = 0u, a2 = 0u;
        {
            float w[4] = {bt.w[4 * gi], bt.w[4 * gi + 1], bt.w[4 * gi + 2], bt.w[4 * gi + 3]};
            uint32_t h[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const bf16 hi = __float2bfloat16(w[i]);
                const bf16 lo = __float2bfloat16(w[i] - __bfloat162float(hi));
                h[i] = g == 0 ? __bfloat16_as_ushort(hi) : g == 1 ? __bfloat16_as_ushort(lo) : 0u;
            }
            a0 = h[0] | (h[1] << 16);
            a2 = h[2] | (h[3] << 16);
        }
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const uint32_t ra[4] = {rows[4 * gi][c].x, rows[4 * gi][c].y, rows[4 * gi][c].z, rows[4 * gi][c].w};
            const uint32_t rb[4] = {rows[4 * gi + 1][c].x, rows[4 * gi + 1][c].y, rows[4 * gi + 1][c].z, rows[4 * gi + 1][c].w};
            const uint32_t rc[4] = {rows[4 * gi + 2][c].x, rows[4 * gi + 2][c].y, rows[4 * gi + 2][c].z, rows[4 * gi + 2][c].w};
            const uint32_t rd[4] = {rows[4 * gi + 3][c].x, rows[4 * gi + 3][c].y, rows[4 * gi + 3][c].z, rows[4 * gi + 3][c].w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {                   // feature 64 c + 8 g + j: (entry 2 t, entry 2 t + 1) and (2 t + 8, 2 t + 9)
                const uint32_t sel = (j & 1) ? 0x7632u : 0x5410u;
                const uint32_t b0 = __byte_perm(ra[j >> 1], rb[j >> 1], sel), b1 = __byte_perm(rc[j >> 1], rd[j >> 1], sel);
                mma16816(acc[8 * c + j], a0, 0u, a2, 0u, b0, b1);
            }
        }
    }
}

// y of the warp = accumulator row 0 + row 1; afterwards lane t < 4 owns features 64 c + 16 t + {0..15}:
// out[c][i] (i < 8) = acc[8 c + i][0], out[c][8 + i] = acc[8 c + i][1]
template <int D>
__device__ __forceinline__ void spmm_fold(float (&acc)[D / 8][4]) {
#pragma unroll
    for (int j = 0; j < D / 8; ++j) {
        acc[j][0] += __shfl_down_sync(FULL, acc[j][0], 4);
        acc[j][1] += __shfl_down_sync(FULL, acc[j][1], 4);
    }
}

template <typename TO>
__device__ __forceinline__ void store8(TO *p, const float (&v)[8]);
template <>
__device__ __forceinline__ void store8<float>(float *p, const float (&v)[8]) {
    reinterpret_cast<float4 *>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4 *>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
template <>
__device__ __forceinline__ void store8<bf16>(bf16 *p, const float (&v)[8]) { Vec16<bf16>::store(p, v); }

template <int D, typename TO>
__device__ __forceinline__ void store_lane_features(TO *yrow, const float (&acc)[D / 8][4], int t) {
#pragma unroll
    for (int c = 0; c < D / 64; ++c) {
        float lo[8], hi[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            lo[i] = acc[8 * c + i][0];
            hi[i] = acc[8 * c + i][1];
        }
        store8<TO>(yrow + 64 * c + 16 * t, lo);
        store8<TO>(yrow + 64 * c + 16 * t + 8, hi);
    }
}

// the lane's 8 entries of the batch starting at `base`: (source row, weight); PERM: weights through the CSC permutation
template <bool PERM>
__device__ __forceinline__ void load_batch(Batch &bt, const int32_t *__restrict__ ip, const int32_t *__restrict__ pm,
                                           const float *__restrict__ vp, int base, int e1, int t) {
#pragma unroll
    for (int gi = 0; gi < 2; ++gi)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int e = base + 16 * gi + 2 * t + (i & 1) + 8 * (i >> 1);
            const bool ok = e < e1;
            bt.row[4 * gi + i] = ok ? __ldg(ip + e) : 0;
            bt.w[4 * gi + i] = ok ? __ldg(vp + (PERM ? __ldg(pm + e) : e)) : 0.0f;
        }
}

// CSR: one warp per output row
template <int D, typename TO>
__global__ void __launch_bounds__(WARPS * 32)
spmm_mma_kernel(const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices, const float *__restrict__ values,
                const bf16 *__restrict__ x, TO *__restrict__ y, int B, int S, int64_t nnz) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row_id = (int64_t)blockIdx.x * WARPS + warp;
    if (row_id >= (int64_t)B * S) return;
    const int b = (int)(row_id / S), r = (int)(row_id % S);
    const int g = lane >> 2, t = lane & 3;
    const int e0 = indptr[r], e1 = indptr[r + 1];
    const int32_t *ip = indices + (size_t)b * nnz;
    const float *vp = values + (size_t)b * nnz;
    const bf16 *xg = x + (size_t)b * S * D + 8 * g;
    float acc[D / 8][4];
#pragma unroll
    for (int j = 0; j < D / 8; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.0f;
    Batch cur, nxt;
    if (e0 < e1) load_batch<false>(cur, ip, nullptr, vp, e0, e1, t);
    for (int base = e0; base < e1; base += 32) {
        if (base + 32 < e1) load_batch<false>(nxt, ip, nullptr, vp, base + 32, e1, t);
        spmm_batch<D>(acc, xg, cur, g);
        cur = nxt;
    }
    spmm_fold<D>(acc);
    if (lane < 4) store_lane_features<D, TO>(y + ((size_t)b * S + r) * D, acc, t);
}

// CSC: one block per output column; the block's warps take the column's batches round-robin, their partial rows are
// added in warp order (column 0 of a lookup pattern collects every row's zero padding: tens of thousands of entries)
template <int D, typename TO>
__global__ void __launch_bounds__(WARPS * 32)
spmm_t_mma_kernel(const int32_t *__restrict__ col_ptr, const int32_t *__restrict__ row_idx, const int32_t *__restrict__ perm,
                  const float *__restrict__ values, const bf16 *__restrict__ x, TO *__restrict__ y, int S, int64_t nnz) {
    __shared__ __align__(16) float s_part[WARPS][D];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t col_id = blockIdx.x;
    const int b = (int)(col_id / S), c = (int)(col_id % S);
    const int g = lane >> 2, t = lane & 3;
    const int32_t *cp = col_ptr + (size_t)b * (S + 1);
    const int e0 = cp[c], e1 = cp[c + 1];
    const int32_t *ip = row_idx + (size_t)b * nnz;
    const int32_t *pm = perm + (size_t)b * nnz;
    const float *vp = values + (size_t)b * nnz;
    const bf16 *xg = x + (size_t)b * S * D + 8 * g;
    float acc[D / 8][4];
#pragma unroll
    for (int j = 0; j < D / 8; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.0f;
    Batch cur, nxt;
    int base = e0 + 32 * warp;
    if (base < e1) load_batch<true>(cur, ip, pm, vp, base, e1, t);
    for (; base < e1; base += 32 * WARPS) {
        if (base + 32 * WARPS < e1) load_batch<true>(nxt, ip, pm, vp, base + 32 * WARPS, e1, t);
        spmm_batch<D>(acc, xg, cur, g);
        cur = nxt;
    }
    spmm_fold<D>(acc);
    if (lane < 4) {
#pragma unroll
        for (int cc = 0; cc < D / 64; ++cc)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                s_part[warp][64 * cc + 16 * t + i] = acc[8 * cc + i][0];
                s_part[warp][64 * cc + 16 * t + 8 + i] = acc[8 * cc + i][1];
            }
    }
    __syncthreads();
    for (int i = threadIdx.x * 8; i < D; i += WARPS * 32 * 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = 0.0f;
#pragma unroll
        for (int w = 0; w < WARPS; ++w)                     // fixed order
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] += s_part[w][i + u];
        store8<TO>(y + ((size_t)b * S + c) * D + i, v);
    }
}

template <int D>
static int launch_sddmm_d(const int32_t *indptr, const int32_t *indices, const bf16 *q, const bf16 *k, float *values, int B,
                          int S, int64_t nnz, float scale, float clamp, cudaStream_t st) {
    constexpr int W = WARPS;
    const int64_t rows = (int64_t)B * S;
    sddmm_mma_kernel<D><<<(unsigned)((rows + W - 1) / W), W * 32, 0, st>>>(indptr, indices, q, k, values, B, S, nnz, scale, clamp);
    return after_launch("sddmm_mma_kernel");
}

template <int D, typename TO>
static int launch_spmm_d(bool trans, const int32_t *ptr, const int32_t *src_idx, const int32_t *perm, const float *values,
                         const bf16 *x, TO *y, int B, int S, int64_t nnz, cudaStream_t st) {
    constexpr int W = WARPS;
    const int64_t rows = (int64_t)B * S;
    if (trans) {
        spmm_t_mma_kernel<D, TO><<<(unsigned)rows, W * 32, 0, st>>>(ptr, src_idx, perm, values, x, y, S, nnz);
        return after_launch("spmm_t_mma_kernel");
    }
    spmm_mma_kernel<D, TO><<<(unsigned)((rows + W - 1) / W), W * 32, 0, st>>>(ptr, src_idx, values, x, y, B, S, nnz);
    return after_launch("spmm_mma_kernel");
}

bool supported(int d, const void *a, const void *b) {
    static const bool off = [] { const char *e = getenv("SPT_CSR_MMA"); return e && atoi(e) == 0; }();   // A/B switch
    return !off && (d == 64 || d == 128) && ((uintptr_t)a % 32 == 0) && ((uintptr_t)b % 32 == 0);
}

int launch_sddmm(const int32_t *indptr, const int32_t *indices, const bf16 *q, const bf16 *k, float *values, int B, int S,
                 int d, int64_t nnz, float scale, float clamp, cudaStream_t st) {
    return d == 64 ? launch_sddmm_d<64>(indptr, indices, q, k, values, B, S, nnz, scale, clamp, st)
                   : launch_sddmm_d<128>(indptr, indices, q, k, values, B, S, nnz, scale, clamp, st);
}

int launch_spmm(bool trans, const int32_t *ptr, const int32_t *src_idx, const int32_t *perm, const float *values,
                const bf16 *x, void *y, bool y_bf16, int B, int S, int d, int64_t nnz, cudaStream_t st) {
    if (d == 64)
        return y_bf16 ? launch_spmm_d<64, bf16>(trans, ptr, src_idx, perm, values, x, (bf16 *)y, B, S, nnz, st)
                      : launch_spmm_d<64, float>(trans, ptr, src_idx, perm, values, x, (float *)y, B, S, nnz, st);
    return y_bf16 ? launch_spmm_d<128, bf16>(trans, ptr, src_idx, perm, values, x, (bf16 *)y, B, S, nnz, st)
                  : launch_spmm_d<128, float>(trans, ptr, src_idx, perm, values, x, (float *)y, B, S, nnz, st);
}

}  // namespace csr_mma
}  // namespace spt
