// (3) sddmm, (4) CSR softmax, (5) spmm / transposed spmm, (a-7) CSR->CSC on a batched CSR whose
// indptr is shared by the whole batch (reference: extension/sddmm.cpp, softmax.cu, spmm.cpp,
// legacy/csr2csc.cpp; call sites naive_gpt/kernels/{sddmm,softmax,spmm}.py).
//
// All kernels are warp-per-row (or warp-per-column for the transposed product) over a general CSR:
// arbitrary indptr, unsorted and possibly duplicated column indices (the lookup stage emits bucket
// order and zero padding).  Memory-bound stages: the index / value streams are read with coalesced
// 128-byte warp accesses; dense rows of q/k/v/x are gathered with 16-byte lane loads, L lanes per
// gathered row (L * 16 B >= row bytes) so that each gathered 128-byte line is touched once.
#include <cstdlib>

#include "common.cuh"

namespace spt {

constexpr int CSR_WARPS = 8;  // warps per block for the warp-per-row kernels

// ---- sddmm -------------------------------------------------------------------------------------
// values[b,e] = scale * <q[b,row(e),:], k[b,idx[b,e],:]> (optionally clamped).
// L lanes per gathered key row, G = 32 / L keys in flight per warp instruction.
template <typename T, int L>
__global__ void __launch_bounds__(CSR_WARPS * 32)
sddmm_kernel(const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices, const T *__restrict__ q,
             const T *__restrict__ k, float *__restrict__ values, int B, int S, int d, int64_t nnz, float scale,
             float clamp) {
    constexpr int VEC = Vec16<T>::N;
    constexpr int G = 32 / L;
    const int lane = threadIdx.x & 31;
    const int64_t row_id = (int64_t)blockIdx.x * CSR_WARPS + (threadIdx.x >> 5);
    if (row_id >= (int64_t)B * S) return;
    const int b = (int)(row_id / S), r = (int)(row_id % S);
    const int sub = lane % L, grp = lane / L;
    const bool has = sub * VEC < d;
    float qv[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) qv[i] = 0.0f;
    if (has) Vec16<T>::load(q + ((size_t)b * S + r) * d + sub * VEC, qv);
    const int e0 = indptr[r], e1 = indptr[r + 1];
    const int32_t *ip = indices + (size_t)b * nnz;
    float *vp = values + (size_t)b * nnz;
    const T *kb = k + (size_t)b * S * d;
    for (int base = e0; base < e1; base += 32) {
        const int e = base + lane;
        const int my_idx = e < e1 ? ip[e] : 0;
        float mine = 0.0f;
        const int cnt = min(32, e1 - base);
#pragma unroll
        for (int st = 0; st < L; ++st) {  // 32 entries = L steps of G entries
            if (st * G >= cnt) break;     // warp-uniform
            const int col = __shfl_sync(FULL, my_idx, st * G + grp);
            float kv[VEC], acc = 0.0f;
            if (has) {
                Vec16<T>::load(kb + (size_t)col * d + sub * VEC, kv);
#pragma unroll
                for (int i = 0; i < VEC; ++i) acc = fmaf(qv[i], kv[i], acc);
            }
            acc = group_sum<L>(acc);
            // entry (st*G + g) was computed by group g; lane l wants entry l
            const float got = __shfl_sync(FULL, acc, (lane % G) * L);
            if (lane / G == st) mine = got;
        }
        if (e < e1) {
            float v = mine * scale;
            if (clamp > 0.0f) v = fminf(fmaxf(v, -clamp), clamp);
            vp[e] = v;
        }
    }
}

// fp32 rows with 256-bit lane loads (sm_100 LDG.E.256): 8 floats per lane, so a d = 64 row takes 8 lanes instead of 16 —
// one shuffle step less in every group sum and half as many steps per batch (the kernel is LSU / shuffle bound)
template <int L>
__global__ void __launch_bounds__(CSR_WARPS * 32)
sddmm_f32w_kernel(const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices, const float *__restrict__ q,
                  const float *__restrict__ k, float *__restrict__ values, int B, int S, int d, int64_t nnz, float scale,
                  float clamp) {
    constexpr int VEC = 8;
    constexpr int G = 32 / L;
    const int lane = threadIdx.x & 31;
    const int64_t row_id = (int64_t)blockIdx.x * CSR_WARPS + (threadIdx.x >> 5);
    if (row_id >= (int64_t)B * S) return;
    const int b = (int)(row_id / S), r = (int)(row_id % S);
    const int sub = lane % L, grp = lane / L;
    const bool has = sub * VEC < d;
    auto ld8 = [](const float *p, float (&v)[8]) {
        uint32_t u[8];
        asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]) : "l"(p));
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(u[i]);
    };
    float qv[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) qv[i] = 0.0f;
    if (has) ld8(q + ((size_t)b * S + r) * d + sub * VEC, qv);
    const int e0 = indptr[r], e1 = indptr[r + 1];
    const int32_t *ip = indices + (size_t)b * nnz;
    float *vp = values + (size_t)b * nnz;
    const float *kb = k + (size_t)b * S * d;
    for (int base = e0; base < e1; base += 32) {
        const int e = base + lane;
        const int my_idx = e < e1 ? ip[e] : 0;
        float mine = 0.0f;
        const int cnt = min(32, e1 - base);
#pragma unroll
        for (int st = 0; st < L; ++st) {  // 32 entries = L steps of G entries
            if (st * G >= cnt) break;     // warp-uniform
            const int col = __shfl_sync(FULL, my_idx, st * G + grp);
            float kv[VEC], acc = 0.0f;
            if (has) {
                ld8(kb + (size_t)col * d + sub * VEC, kv);
#pragma unroll
                for (int i = 0; i < VEC; ++i) acc = fmaf(qv[i], kv[i], acc);
            }
            acc = group_sum<L>(acc);
            const float got = __shfl_sync(FULL, acc, (lane % G) * L);
            if (lane / G == st) mine = got;
        }
        if (e < e1) {
            float v = mine * scale;
            if (clamp > 0.0f) v = fminf(fmaxf(v, -clamp), clamp);
            vp[e] = v;
        }
    }
}

// scalar fallback for head dims that are not a multiple of the 16-byte vector
template <typename T>
__global__ void __launch_bounds__(CSR_WARPS * 32)
sddmm_scalar_kernel(const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices, const T *__restrict__ q,
                    const T *__restrict__ k, float *__restrict__ values, int B, int S, int d, int64_t nnz,
                    float scale, float clamp) {
    const int lane = threadIdx.x & 31;
    const int64_t row_id = (int64_t)blockIdx.x * CSR_WARPS + (threadIdx.x >> 5);
    if (row_id >= (int64_t)B * S) return;
    const int b = (int)(row_id / S), r = (int)(row_id % S);
    const T *qp = q + ((size_t)b * S + r) * d;
    for (int e = indptr[r] + lane; e < indptr[r + 1]; e += 32) {
        const T *kp = k + ((size_t)b * S + indices[(size_t)b * nnz + e]) * d;
        float acc = 0.0f;
        for (int i = 0; i < d; ++i) acc = fmaf(to_f32(qp[i]), to_f32(kp[i]), acc);
        float v = acc * scale;
        if (clamp > 0.0f) v = fminf(fmaxf(v, -clamp), clamp);
        values[(size_t)b * nnz + e] = v;
    }
}

// 32 bytes per lane through ONE 256-bit load (sm_100 LDG.E.256): a bf16 d = 64 row takes 4 lanes instead of 8 (fp32: 8
// instead of 16), i.e. twice the gathered rows per load instruction and half the shuffles / index arithmetic per pair —
// the gather kernels are LSU- and issue-bound, not bandwidth-bound
__device__ __forceinline__ void ldg256_u32(const void *p, uint32_t (&u)[8]) {
    asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]) : "l"(p));
}
template <typename T>
struct Vec32;
template <>
struct Vec32<float> {
    static constexpr int N = 8;
    __device__ __forceinline__ static void load(const float *p, float (&o)[8]) {
        uint32_t u[8];
        ldg256_u32(p, u);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = __uint_as_float(u[i]);
    }
};
template <>
struct Vec32<__nv_bfloat16> {
    static constexpr int N = 16;
    __device__ __forceinline__ static void load(const __nv_bfloat16 *p, float (&o)[16]) {
        uint32_t u[8];
        ldg256_u32(p, u);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            o[2 * i] = __uint_as_float(u[i] << 16);
            o[2 * i + 1] = __uint_as_float(u[i] & 0xffff0000u);
        }
    }
};
// N fp32 values -> N consecutive elements of TO (16-byte stores)
template <typename TO, int N>
__device__ __forceinline__ void store_vec(TO *p, const float (&o)[N]) {
    constexpr int V = Vec16<TO>::N;
    if constexpr (N % V == 0) {
#pragma unroll
        for (int i = 0; i < N; i += V) {
            float t[V];
#pragma unroll
            for (int u = 0; u < V; ++u) t[u] = o[i + u];
            Vec16<TO>::store(p + i, t);
        }
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i) p[i] = from_f32<TO>(o[i]);
    }
}

// ---- spmm (y = A x) and transposed spmm (y = A^T x through the CSC) ------------------------------
// One warp per output row.  TRANS = false: entries e in [indptr[r], indptr[r+1]), source row
// indices[b,e], weight values[b,e].  TRANS = true: entries e' in [col_ptr[b,c], col_ptr[b,c+1]),
// source row row_idx[b,e'], weight values[b, perm[b,e']].  Accumulation order inside a group is the
// entry order; groups are combined by a fixed butterfly => deterministic.
template <typename T, typename TO, int L, bool TRANS, typename VT = Vec16<T>>
__global__ void __launch_bounds__(CSR_WARPS * 32)
spmm_kernel(const int32_t *__restrict__ ptr, const int32_t *__restrict__ src_idx, const int32_t *__restrict__ perm,
            const float *__restrict__ values, const T *__restrict__ x, TO *__restrict__ y, int B, int S, int d,
            int64_t nnz) {
    constexpr int VEC = VT::N;
    constexpr int G = 32 / L;
    const int lane = threadIdx.x & 31;
    const int64_t row_id = (int64_t)blockIdx.x * CSR_WARPS + (threadIdx.x >> 5);
    if (row_id >= (int64_t)B * S) return;
    const int b = (int)(row_id / S), r = (int)(row_id % S);
    const int sub = lane % L, grp = lane / L;
    const bool has = sub * VEC < d;
    const int32_t *pp = TRANS ? ptr + (size_t)b * (S + 1) : ptr;
    const int e0 = pp[r], e1 = pp[r + 1];
    const int32_t *ip = src_idx + (size_t)b * nnz;
    const int32_t *pm = TRANS ? perm + (size_t)b * nnz : nullptr;
    const float *vp = values + (size_t)b * nnz;
    const T *xb = x + (size_t)b * S * d;
    float acc[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[i] = 0.0f;
    for (int base = e0; base < e1; base += 32) {
        const int e = base + lane;
        int my_idx = 0;
        float my_val = 0.0f;
        if (e < e1) {
            my_idx = ip[e];
            my_val = TRANS ? vp[pm[e]] : vp[e];
        }
        const int cnt = min(32, e1 - base);
#pragma unroll
        for (int st = 0; st < L; ++st) {
            if (st * G >= cnt) break;  // warp-uniform
            const int src = st * G + grp;
            const int col = __shfl_sync(FULL, my_idx, src);
            const float w = __shfl_sync(FULL, my_val, src);  // 0 for padding lanes
            if (has) {
                float xv[VEC];
                VT::load(xb + (size_t)col * d + sub * VEC, xv);
#pragma unroll
                for (int i = 0; i < VEC; ++i) acc[i] = fmaf(w, xv[i], acc[i]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
#pragma unroll
        for (int o = L; o < 32; o <<= 1) acc[i] += __shfl_xor_sync(FULL, acc[i], o);
    }
    if (grp == 0 && has) store_vec<TO, VEC>(y + ((size_t)b * S + r) * d + sub * VEC, acc);
}

template <typename T, typename TO, bool TRANS>
__global__ void __launch_bounds__(CSR_WARPS * 32)
spmm_scalar_kernel(const int32_t *__restrict__ ptr, const int32_t *__restrict__ src_idx,
                   const int32_t *__restrict__ perm, const float *__restrict__ values, const T *__restrict__ x,
                   TO *__restrict__ y, int B, int S, int d, int64_t nnz) {
    const int lane = threadIdx.x & 31;
    const int64_t row_id = (int64_t)blockIdx.x * CSR_WARPS + (threadIdx.x >> 5);
    if (row_id >= (int64_t)B * S) return;
    const int b = (int)(row_id / S), r = (int)(row_id % S);
    const int32_t *pp = TRANS ? ptr + (size_t)b * (S + 1) : ptr;
    for (int i = lane; i < d; i += 32) {
        float acc = 0.0f;
        for (int e = pp[r]; e < pp[r + 1]; ++e) {
            const size_t be = (size_t)b * nnz + e;
            const float w = TRANS ? values[(size_t)b * nnz + perm[be]] : values[be];
            acc = fmaf(w, to_f32(x[((size_t)b * S + src_idx[be]) * d + i]), acc);
        }
        y[((size_t)b * S + r) * d + i] = from_f32<TO>(acc);
    }
}

// ---- CSR softmax -------------------------------------------------------------------------------
// Warp per row; entries strided over lanes (coalesced).  exp without max subtraction, causal
// predicate as a 0/1 factor, denominator >= 1e-9 (softmax.cu:16-46).
__global__ void __launch_bounds__(CSR_WARPS * 32)
softmax_fwd_kernel(const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                   const float *__restrict__ values, float *__restrict__ output, int B, int S, int64_t nnz) {
    const int lane = threadIdx.x & 31;
    const int64_t row_id = (int64_t)blockIdx.x * CSR_WARPS + (threadIdx.x >> 5);
    if (row_id >= (int64_t)B * S) return;
    const int b = (int)(row_id / S), r = (int)(row_id % S);
    const int e0 = indptr[r], e1 = indptr[r + 1];
    const int32_t *ip = indices + (size_t)b * nnz;
    const float *vp = values + (size_t)b * nnz;
    float *op = output + (size_t)b * nnz;
    constexpr int KEEP = 8;  // rows up to 256 entries stay in registers (the layer's k = S/8 at S = 2048)
    float ex[KEEP];
    float sum = 0.0f;
    int it = 0;
    for (int e = e0 + lane; e < e1; e += 32, ++it) {
        const float v = (ip[e] <= r) ? __expf(vp[e]) : 0.0f;
        if (it < KEEP) ex[it] = v;
        sum += v;
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / fmaxf(1e-9f, sum);
    it = 0;
    for (int e = e0 + lane; e < e1; e += 32, ++it) {
        const float v = it < KEEP ? ex[it] : ((ip[e] <= r) ? __expf(vp[e]) : 0.0f);
        op[e] = v * inv;
    }
}

// dv = y * (dy - sum(y * dy)) on kept entries (true gradient; see header comment in spt_b200.h).
__global__ void __launch_bounds__(CSR_WARPS * 32)
softmax_bwd_kernel(const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                   const float *__restrict__ output, const float *__restrict__ grad_output,
                   float *__restrict__ grad_values, int B, int S, int64_t nnz, int reference_clamp,
                   const float *__restrict__ clamped, float scale, float clamp) {
    const int lane = threadIdx.x & 31;
    const int64_t row_id = (int64_t)blockIdx.x * CSR_WARPS + (threadIdx.x >> 5);
    if (row_id >= (int64_t)B * S) return;
    const int b = (int)(row_id / S), r = (int)(row_id % S);
    const int e0 = indptr[r], e1 = indptr[r + 1];
    const int32_t *ip = indices + (size_t)b * nnz;
    const float *yp = output + (size_t)b * nnz;
    const float *gp = grad_output + (size_t)b * nnz;
    float *op = grad_values + (size_t)b * nnz;
    // clamped != nullptr: the input of the softmax was v = clamp(scale * raw, -clamp, clamp) (the stage layer's scores) and the
    // gradient w.r.t. raw is wanted: the clamp's zero-gradient mask and the scale are applied to the stored value (same
    // roundings as softmax_bwd followed by clamp_scale_bwd: one pass instead of two over the nnz arrays)
    const float *cp = clamped ? clamped + (size_t)b * nnz : nullptr;
    constexpr int KEEP = 8;
    float ys[KEEP], gs[KEEP];
    float sum = 0.0f;
    int it = 0;
    for (int e = e0 + lane; e < e1; e += 32, ++it) {
        const bool keep = ip[e] <= r;
        const float yv = keep ? yp[e] : 0.0f, gv = gp[e];
        if (it < KEEP) { ys[it] = yv; gs[it] = gv; }
        sum = fmaf(yv, gv, sum);
    }
    sum = warp_sum(sum);
    if (reference_clamp) sum = fmaxf(1e-9f, sum);   // opt-in: the shipped kernel's clamp (softmax.cu:69), for A/B parity runs
    it = 0;
    for (int e = e0 + lane; e < e1; e += 32, ++it) {
        float yv, gv;
        if (it < KEEP) { yv = ys[it]; gv = gs[it]; }
        else { yv = (ip[e] <= r) ? yp[e] : 0.0f; gv = gp[e]; }
        float o = yv * (gv - sum);
        if (cp) o = (clamp <= 0.0f || fabsf(cp[e]) < clamp) ? o * scale : 0.0f;
        op[e] = o;
    }
}

// ---- CSR -> CSC ----------------------------------------------------------------------------------
// Deterministic, stable counting sort by column, tiled over rows:
//   K1  per (tile of TR rows, batch): shared-memory histogram of the tile's columns -> tile_cnt[b][tile][c]
//   K2  exclusive scan over tiles for every (batch, column) in parallel (tile_cnt becomes the start offset
//       of (tile, column)), then a per-batch exclusive scan over the column totals -> col_ptr[b][c]
//   K3  placement: staged through shared memory (csr2csc_place_staged_kernel below) when it fits, else
//       one warp per (tile, batch) walking the tile's rows in order, 32 entries at a time, equal columns
//       ranked by lane with __match_any_sync.  Either way the result does not depend on scheduling.
constexpr int C2C_TR = 64;

__global__ void __launch_bounds__(256)
csr2csc_count_kernel(const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                     int32_t *__restrict__ tile_cnt, int S, int64_t nnz, int n_tiles) {
    extern __shared__ int32_t s_cnt[];
    const int tile = blockIdx.x, b = blockIdx.y;
    for (int c = threadIdx.x; c < S; c += blockDim.x) s_cnt[c] = 0;
    __syncthreads();
    const int r0 = tile * C2C_TR, r1 = min(S, r0 + C2C_TR);
    const int e0 = indptr[r0], e1 = indptr[r1];
    const int32_t *ip = indices + (size_t)b * nnz;
    for (int e = e0 + threadIdx.x; e < e1; e += blockDim.x) {
        const int c = ip[e];
        if ((unsigned)c < (unsigned)S) atomicAdd(&s_cnt[c], 1);
    }
    __syncthreads();
    int32_t *out = tile_cnt + ((size_t)b * n_tiles + tile) * S;
    for (int c = threadIdx.x; c < S; c += blockDim.x) out[c] = s_cnt[c];
}

__global__ void csr2csc_place_kernel(const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                                     const int32_t *__restrict__ tile_cnt, const int32_t *__restrict__ col_ptr,
                                     int32_t *__restrict__ row_idx, int32_t *__restrict__ perm, int S, int64_t nnz,
                                     int n_tiles) {
    extern __shared__ int32_t s_all[];  // [warps][S] running cursors
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int tile = blockIdx.x * nw + wid, b = blockIdx.y;
    if (tile >= n_tiles) return;
    int32_t *cur = s_all + (size_t)wid * S;
    const int32_t *tc = tile_cnt + ((size_t)b * n_tiles + tile) * S;
    const int32_t *cp = col_ptr + (size_t)b * (S + 1);
    for (int c = lane; c < S; c += 32) cur[c] = cp[c] + tc[c];
    __syncwarp();
    const int32_t *ip = indices + (size_t)b * nnz;
    int32_t *ro = row_idx + (size_t)b * nnz, *po = perm + (size_t)b * nnz;
    const int r0 = tile * C2C_TR, r1 = min(S, r0 + C2C_TR);
    for (int r = r0; r < r1; ++r) {
        const int e0 = indptr[r], e1 = indptr[r + 1];
        for (int base = e0; base < e1; base += 32) {
            const int e = base + lane;
            const int c = e < e1 ? ip[e] : -1;
            const bool ok = (unsigned)c < (unsigned)S;  // out-of-range columns are dropped (as in K1)
            const unsigned active = __ballot_sync(FULL, ok);
            if (ok) {
                const unsigned same = __match_any_sync(active, c);
                const int rank = __popc(same & ((1u << lane) - 1u));
                const int leader = __ffs(same) - 1;
                int32_t start = 0;
                if (lane == leader) {
                    start = cur[c];
                    cur[c] = start + __popc(same);
                }
                start = __shfl_sync(same, start, leader);
                ro[start + rank] = r;
                po[start + rank] = e;
            }
            __syncwarp();
        }
    }
}

// Parallel first half of the scan: one thread per (batch, column) turns the tile counts into exclusive
// tile offsets and leaves the column total in col_tot[b][c] (grid (ceil(S/256), B)); the per-batch
// exclusive scan over columns (csr2csc_colscan_kernel) then gives col_ptr.
__global__ void __launch_bounds__(256)
csr2csc_tilescan_kernel(int32_t *__restrict__ tile_cnt, int32_t *__restrict__ col_tot, int S, int n_tiles) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (c >= S) return;
    int32_t *tc = tile_cnt + (size_t)b * n_tiles * S + c;
    int32_t run = 0;
    for (int t = 0; t < n_tiles; ++t) {
        const int32_t v = tc[(size_t)t * S];
        tc[(size_t)t * S] = run;
        run += v;
    }
    col_tot[(size_t)b * (S + 1) + c] = run;
}

// block-wide exclusive scan of n int32 values in shared memory (in place); returns the total
__device__ __forceinline__ int32_t block_exclusive_scan(int32_t *vals, int n, int32_t *s_warp, int32_t *s_carry) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (threadIdx.x == 0) *s_carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += blockDim.x) {
        const int c = base + threadIdx.x;
        const int32_t v = c < n ? vals[c] : 0;
        int32_t inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int32_t t = __shfl_up_sync(FULL, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) s_warp[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            int32_t w = lane < nw ? s_warp[lane] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int32_t t = __shfl_up_sync(FULL, w, o);
                if (lane >= o) w += t;
            }
            s_warp[lane] = w;  // inclusive
        }
        __syncthreads();
        const int32_t carry = *s_carry;
        const int32_t warp_off = wid > 0 ? s_warp[wid - 1] : 0;
        if (c < n) vals[c] = carry + warp_off + inc - v;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) *s_carry = carry + s_warp[nw - 1];
        __syncthreads();
    }
    return *s_carry;
}

__global__ void __launch_bounds__(1024)
csr2csc_colscan_kernel(int32_t *__restrict__ col_ptr, int S) {
    extern __shared__ int32_t s_tot[];
    __shared__ int32_t s_warp[32];
    __shared__ int32_t s_carry;
    int32_t *cp = col_ptr + (size_t)blockIdx.x * (S + 1);
    for (int c = threadIdx.x; c < S; c += blockDim.x) s_tot[c] = cp[c];
    __syncthreads();
    const int32_t total = block_exclusive_scan(s_tot, S, s_warp, &s_carry);
    for (int c = threadIdx.x; c < S; c += blockDim.x) cp[c] = s_tot[c];
    if (threadIdx.x == 0) cp[S] = total;
}

// Staged placement.  A block owns one tile of C2C_TR rows; the tile's entries are consumed in chunks of
// CAP entries.  Inside a chunk the 8 warps own 8 consecutive entry ranges (so (warp, position) order =
// entry order = row-major order: the sort is stable) and place (row, entry, column) into shared-memory
// arrays sorted by column; the block then writes the sorted chunk out, consecutive threads to consecutive
// addresses of a column's run — instead of the two scattered 4-byte stores per entry of a direct
// placement.  Counting uses shared-memory atomics (order-free).  Placement takes its slot from an atomic
// cursor when the 32 columns of a warp instruction are all different (checked through a tag array: the
// common case, every key of a row is distinct), and ranks equal columns by lane with __match_any_sync
// otherwise (zero padding runs) — so the result never depends on scheduling.  Deterministic.
constexpr int C2C_WARPS = 8;

__global__ void __launch_bounds__(C2C_WARPS * 32)
csr2csc_place_staged_kernel(const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                            const int32_t *__restrict__ tile_cnt, const int32_t *__restrict__ col_ptr,
                            int32_t *__restrict__ row_idx, int32_t *__restrict__ perm, int S, int64_t nnz,
                            int n_tiles, int CAP, const int *__restrict__ run_if) {
    if (run_if && *run_if == 0) return;      // the bit-matrix placement below handled this call
    extern __shared__ __align__(16) unsigned char c2c_smem[];
    __shared__ int32_t s_warp[32];
    __shared__ int32_t s_carry;
    __shared__ int32_t s_indptr[C2C_TR + 1];
    int32_t *s_goff = reinterpret_cast<int32_t *>(c2c_smem);                 // [S] next global slot of column c
    int32_t *s_lstart = s_goff + S;                                          // [S] chunk-local start of column c
    int32_t *s_perm = s_lstart + S;                                          // [CAP]
    int32_t *s_cur = s_perm + CAP;                                           // [8][S] per-warp counts -> cursors
    uint16_t *s_col = reinterpret_cast<uint16_t *>(s_cur + (size_t)C2C_WARPS * S);   // [CAP]
    uint8_t *s_row = reinterpret_cast<uint8_t *>(s_col + CAP);               // [CAP]
    uint8_t *s_tag = s_row + CAP;                                            // [8][S] last lane that touched column c
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int tile = blockIdx.x, b = blockIdx.y;
    const int r0 = tile * C2C_TR, r1 = min(S, r0 + C2C_TR);
    const int32_t *tc = tile_cnt + ((size_t)b * n_tiles + tile) * S;
    const int32_t *cp = col_ptr + (size_t)b * (S + 1);
    const int32_t *ip = indices + (size_t)b * nnz;
    int32_t *ro = row_idx + (size_t)b * nnz, *po = perm + (size_t)b * nnz;
    for (int c = threadIdx.x; c < S; c += blockDim.x) s_goff[c] = cp[c] + tc[c];
    for (int i = threadIdx.x; i <= r1 - r0; i += blockDim.x) s_indptr[i] = indptr[r0 + i];
    __syncthreads();
    const int e_begin = s_indptr[0], e_end = s_indptr[r1 - r0];
    int32_t *cur = s_cur + (size_t)wid * S;
    uint8_t *tag = s_tag + (size_t)wid * S;
    for (int cb = e_begin; cb < e_end; cb += CAP) {
        const int ce = min(e_end, cb + CAP);
        const int per_warp = ((ce - cb + C2C_WARPS * 32 - 1) / (C2C_WARPS * 32)) * 32;   // multiple of 32
        const int w0 = cb + wid * per_warp, w1 = min(ce, w0 + per_warp);
        for (int i = threadIdx.x; i < C2C_WARPS * S; i += blockDim.x) s_cur[i] = 0;
        __syncthreads();
        // pass 1: per-warp column counts of its entry range
        for (int e = w0 + lane; e < w1; e += 32) {
            const int c = ip[e];
            if ((unsigned)c < (unsigned)S) atomicAdd(&cur[c], 1);
        }
        __syncthreads();
        // exclusive prefix over the warps for every column, column totals -> local starts
        for (int c = threadIdx.x; c < S; c += blockDim.x) {
            int run = 0;
#pragma unroll
            for (int w = 0; w < C2C_WARPS; ++w) {
                const int v = s_cur[(size_t)w * S + c];
                s_cur[(size_t)w * S + c] = run;
                run += v;
            }
            s_lstart[c] = run;
        }
        __syncthreads();
        const int placed = block_exclusive_scan(s_lstart, S, s_warp, &s_carry);
        // pass 2: placement into the sorted shared-memory chunk
        for (int base = w0; base < w1; base += 32) {
            const int e = base + lane;
            const int c = e < w1 ? ip[e] : -1;
            const bool ok = (unsigned)c < (unsigned)S;
            if (ok) tag[c] = (uint8_t)lane;
            __syncwarp();
            const bool dup = ok && tag[c] != (uint8_t)lane;        // another lane of this instruction has my column
            int slot = 0;
            if (__any_sync(FULL, dup)) {
                const unsigned active = __ballot_sync(FULL, ok);
                if (ok) {
                    const unsigned same = __match_any_sync(active, c);
                    const int leader = __ffs(same) - 1;
                    int start = 0;
                    if (lane == leader) {
                        start = cur[c];
                        cur[c] = start + __popc(same);
                    }
                    slot = __shfl_sync(same, start, leader) + __popc(same & ((1u << lane) - 1u));
                }
            } else if (ok) {
                slot = atomicAdd(&cur[c], 1);
            }
            if (ok) {
                const int pos = s_lstart[c] + slot;
                int lo = 0, hi = r1 - r0;                  // row of entry e: last i with s_indptr[i] <= e
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (s_indptr[mid] <= e) lo = mid; else hi = mid;
                }
                s_perm[pos] = e;
                s_col[pos] = (uint16_t)c;
                s_row[pos] = (uint8_t)lo;
            }
            __syncwarp();
        }
        __syncthreads();
        // write-out: entry i of the sorted chunk goes to its column's run in the global CSC
        for (int i = threadIdx.x; i < placed; i += blockDim.x) {
            const int c = s_col[i];
            const int g = s_goff[c] + (i - s_lstart[c]);
            ro[g] = r0 + s_row[i];
            po[g] = s_perm[i];
        }
        __syncthreads();
        for (int c = threadIdx.x; c < S; c += blockDim.x) {
            const int next = c + 1 < S ? s_lstart[c + 1] : placed;
            s_goff[c] += next - s_lstart[c];
        }
        __syncthreads();
    }
}

// Placement through a per-tile BIT MATRIX (the common pattern: S <= 2048, rows of up to 256 entries, no column repeated
// inside a row except column 0 — what the lookup stage emits: distinct keys + zero padding).  For the tile's 64 rows,
//   MT[c]   (64 bits)  = the rows that contain column c        (shared-memory atomicOr, order-free)
//   P[r][c] (1 byte)   = position of column c inside row r
// give everything the sorted output needs without sorting: after an exclusive scan of popc(MT[c]) over the columns, output
// slot i of the tile belongs to the column found by a binary search over the scan, its row is the rank-th set bit of
// MT[c], its CSR position comes from P — consecutive threads write consecutive slots of a column's run.  Column 0 (zero
// padding: many entries per row) goes through a row-major list built with warp ballots.  No per-warp counters, no
// tags, no __match_any, 32 warps per block instead of 8.  Anything else (a repeated non-zero column, a longer row) raises
// `fallback` and the staged kernel above redoes the call: deterministic and stable either way.
constexpr int C2B_THREADS = 1024;
constexpr int C2B_SLOTS = C2C_TR * 256;          // entries of a tile

__global__ void __launch_bounds__(C2B_THREADS)
csr2csc_place_bits_kernel(const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                          const int32_t *__restrict__ tile_cnt, const int32_t *__restrict__ col_ptr,
                          int32_t *__restrict__ row_idx, int32_t *__restrict__ perm, int S, int64_t nnz, int n_tiles,
                          int *__restrict__ fallback) {
    extern __shared__ __align__(16) unsigned char c2b_smem[];
    __shared__ int32_t s_warp[32];
    __shared__ int32_t s_carry;
    __shared__ int32_t s_indptr[C2C_TR + 1];
    __shared__ int32_t s_cnt0[C2C_TR], s_zstart[C2C_TR + 1];
    uint32_t *mt_lo = reinterpret_cast<uint32_t *>(c2b_smem);          // [S] rows 0 .. 31 of the tile; later: global offset of slot 0 of the column
    uint32_t *mt_hi = mt_lo + S;                                       // [S] rows 32 .. 63
    int32_t *lstart = reinterpret_cast<int32_t *>(mt_hi + S);          // [S] exclusive scan of the column counts
    uint16_t *s_col = reinterpret_cast<uint16_t *>(lstart + S);        // [slots] column of a sorted slot
    uint8_t *s_row = reinterpret_cast<uint8_t *>(s_col + C2B_SLOTS);   // [slots] its row inside the tile
    uint8_t *s_pos = s_row + C2B_SLOTS;                                // [slots] its position inside the row
    uint8_t *P = s_pos + C2B_SLOTS;                                    // [C2C_TR][S]
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int tile = blockIdx.x, b = blockIdx.y;
    const int r0 = tile * C2C_TR, r1 = min(S, r0 + C2C_TR), n_rows = r1 - r0;
    const int32_t *ip = indices + (size_t)b * nnz;
    for (int c = threadIdx.x; c < S; c += C2B_THREADS) mt_lo[c] = mt_hi[c] = 0u;
    for (int i = threadIdx.x; i <= n_rows; i += C2B_THREADS) s_indptr[i] = indptr[r0 + i];
    if (threadIdx.x < C2C_TR) s_cnt0[threadIdx.x] = 0;
    __syncthreads();
    // pass 1: bit matrix, positions, zero counts; one warp per row
    bool bad = false;
    for (int r = wid; r < n_rows; r += C2B_THREADS / 32) {
        const int e0 = s_indptr[r];
        int e1 = s_indptr[r + 1];
        if (e1 - e0 > 256) {                                         // positions would not fit a byte: fallback (warp-uniform)
            bad = true;
            e1 = e0 + 256;
        }
        uint32_t *word = r < 32 ? mt_lo : mt_hi;
        const uint32_t bit = 1u << (r & 31);
        int zeros = 0;
        for (int base = e0; base < e1; base += 32) {
            const int e = base + lane;
            const int c = e < e1 ? ip[e] : -1;
            const bool ok = (unsigned)c < (unsigned)S;               // out-of-range columns are dropped (as in the count kernel)
            zeros += __popc(__ballot_sync(FULL, ok && c == 0));
            if (ok && c != 0) {
                const uint32_t old = atomicOr(&word[c], bit);
                bad |= (old & bit) != 0;                               // the column twice in one row
                P[(size_t)r * S + c] = (uint8_t)(e - e0);
            }
        }
        if (lane == 0) s_cnt0[r] = zeros;
    }
    if (__any_sync(FULL, bad)) {
        if (lane == 0) atomicOr(fallback, 1);
    }
    __syncthreads();
    // column counts -> exclusive scan; zero-list offsets per row
    if (wid == 0) {
        int run = 0;
        for (int base = 0; base < C2C_TR; base += 32) {
            const int v = base + lane < n_rows ? s_cnt0[base + lane] : 0;
            int inc = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(FULL, inc, o);
                if (lane >= o) inc += t;
            }
            if (base + lane < C2C_TR) s_zstart[base + lane] = run + inc - v;
            run += __shfl_sync(FULL, inc, 31);
        }
        if (lane == 0) s_zstart[C2C_TR] = run;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < S; c += C2B_THREADS) lstart[c] = c == 0 ? s_zstart[C2C_TR] : __popc(mt_lo[c]) + __popc(mt_hi[c]);
    __syncthreads();
    const int total = block_exclusive_scan(lstart, S, s_warp, &s_carry);
    // pass 2: the column-0 entries in row-major order take the first slots
    for (int r = wid; r < n_rows; r += C2B_THREADS / 32) {
        if (s_cnt0[r] == 0) continue;                                  // warp-uniform
        const int e0 = s_indptr[r], e1 = min(s_indptr[r + 1], e0 + 256);
        int at = s_zstart[r];
        for (int base = e0; base < e1; base += 32) {
            const int e = base + lane;
            const bool z = e < e1 && ip[e] == 0;
            const unsigned m = __ballot_sync(FULL, z);
            if (z) {
                const int slot = at + __popc(m & ((1u << lane) - 1u));
                s_row[slot] = (uint8_t)r;
                s_pos[slot] = (uint8_t)(e - e0);
                s_col[slot] = 0;
            }
            at += __popc(m);
        }
    }
    // expand: every other column lists its rows (ascending) into its slots
    const int32_t *tc = tile_cnt + ((size_t)b * n_tiles + tile) * S;
    const int32_t *cp = col_ptr + (size_t)b * (S + 1);
    for (int c = threadIdx.x; c < S; c += C2B_THREADS) {
        const int first = lstart[c];
        if (c != 0) {
            int slot = first;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint32_t m = half ? mt_hi[c] : mt_lo[c];
                while (m) {
                    const int r = 32 * half + __ffs(m) - 1;
                    m &= m - 1;
                    s_row[slot] = (uint8_t)r;
                    s_pos[slot] = P[(size_t)r * S + c];
                    s_col[slot] = (uint16_t)c;
                    ++slot;
                }
            }
        }
        mt_lo[c] = (uint32_t)(cp[c] + tc[c] - first);                  // global position of the column's slot 0 minus its first slot
    }
    __syncthreads();
    if (*reinterpret_cast<volatile int *>(fallback)) return;           // some block of this launch bailed out: the staged kernel redoes everything
    // write-out: consecutive threads -> consecutive slots of a column's run
    int32_t *ro = row_idx + (size_t)b * nnz, *po = perm + (size_t)b * nnz;
    for (int i = threadIdx.x; i < total; i += C2B_THREADS) {
        const int r = s_row[i];
        const int g = (int)mt_lo[s_col[i]] + i;
        ro[g] = r0 + r;
        po[g] = s_indptr[r] + s_pos[i];
    }
}

// ---- transposed spmm through the CSC, one BLOCK per output column -----------------------------------
// Column lengths are very uneven (column 0 collects every row's zero padding: tens of thousands of entries
// against ~S/8 for a typical column), so a warp per column leaves the whole launch waiting for a few
// warps.  Here the 8 warps of a block split the column's entries into contiguous ranges, and their
// partial sums are added in warp order through shared memory: balanced and deterministic.
template <typename T, typename TO, int L>
__global__ void __launch_bounds__(CSR_WARPS * 32)
spmm_t_block_kernel(const int32_t *__restrict__ col_ptr, const int32_t *__restrict__ row_idx,
                    const int32_t *__restrict__ perm, const float *__restrict__ values, const T *__restrict__ x,
                    TO *__restrict__ y, int B, int S, int d, int64_t nnz) {
    constexpr int VEC = Vec16<T>::N;
    constexpr int G = 32 / L;
    __shared__ float s_part[CSR_WARPS][L * VEC];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int b = blockIdx.x / S, c = blockIdx.x % S;
    const int sub = lane % L, grp = lane / L;
    const bool has = sub * VEC < d;
    const int32_t *pp = col_ptr + (size_t)b * (S + 1);
    const int c0 = pp[c], c1 = pp[c + 1];
    const int per_warp = ((c1 - c0 + CSR_WARPS * 32 - 1) / (CSR_WARPS * 32)) * 32;
    const int e0 = c0 + wid * per_warp, e1 = min(c1, e0 + per_warp);
    const int32_t *ip = row_idx + (size_t)b * nnz;
    const int32_t *pm = perm + (size_t)b * nnz;
    const float *vp = values + (size_t)b * nnz;
    const T *xb = x + (size_t)b * S * d;
    float acc[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[i] = 0.0f;
    for (int base = e0; base < e1; base += 32) {
        const int e = base + lane;
        int my_idx = 0;
        float my_val = 0.0f;
        if (e < e1) {
            my_idx = ip[e];
            my_val = vp[pm[e]];
        }
        const int cnt = min(32, e1 - base);
#pragma unroll
        for (int st = 0; st < L; ++st) {
            if (st * G >= cnt) break;  // warp-uniform
            const int src = st * G + grp;
            const int row = __shfl_sync(FULL, my_idx, src);
            const float w = __shfl_sync(FULL, my_val, src);  // 0 for padding lanes
            if (has) {
                float xv[VEC];
                Vec16<T>::load(xb + (size_t)row * d + sub * VEC, xv);
#pragma unroll
                for (int i = 0; i < VEC; ++i) acc[i] = fmaf(w, xv[i], acc[i]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
#pragma unroll
        for (int o = L; o < 32; o <<= 1) acc[i] += __shfl_xor_sync(FULL, acc[i], o);
    }
    if (grp == 0) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) s_part[wid][sub * VEC + i] = acc[i];
    }
    __syncthreads();
    if (wid == 0 && grp == 0 && has) {
        float out[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            float t = 0.0f;
#pragma unroll
            for (int w = 0; w < CSR_WARPS; ++w) t += s_part[w][sub * VEC + i];
            out[i] = t;
        }
        TO *yp = y + ((size_t)b * S + c) * d + sub * VEC;
#pragma unroll
        for (int i = 0; i < VEC; ++i) yp[i] = from_f32<TO>(out[i]);
    }
}

// ---- transposed spmm, second version: columns grouped by length ------------------------------------------------------
// spmm_t_block_kernel gives every column a block of 8 warps; for the typical column (a few hundred entries) that is one
// batch per warp plus a block-wide reduction.  Here a block owns CSR_WARPS consecutive output columns: a column of up to
// SPMM_T_HEAVY entries is summed by ONE warp; longer ones (column 0 of a lookup pattern collects every row's zero
// padding, the first ~k columns are selected by almost every row) are split among the block's warps afterwards, whose
// partial sums are added in warp order.  Same per-entry arithmetic and order inside a warp as above.
// (Also tried for both products: all row loads of a batch issued before any arithmetic + packed fma.rn.f32x2 — 102
// registers, 23 % occupancy, 1.4x SLOWER than the simple loop; dropped.)
template <typename T, int L, bool PERM, typename VT>
__device__ __forceinline__ void spmm_accumulate(float (&acc)[VT::N], const int32_t *__restrict__ ip,
                                                const int32_t *__restrict__ pm, const float *__restrict__ vp,
                                                const T *__restrict__ xb, int e0, int e1, int lane, int d, bool has) {
    constexpr int VEC = VT::N, G = 32 / L;
    const int sub = lane % L, grp = lane / L;
    for (int base = e0; base < e1; base += 32) {
        const int e = base + lane;
        int my_idx = 0;
        float my_val = 0.0f;
        if (e < e1) {
            my_idx = ip[e];
            my_val = vp[PERM ? pm[e] : e];
        }
        const int cnt = min(32, e1 - base);
#pragma unroll
        for (int st = 0; st < L; ++st) {
            if (st * G >= cnt) break;  // warp-uniform
            const int src = st * G + grp;
            const int row = __shfl_sync(FULL, my_idx, src);
            const float w = __shfl_sync(FULL, my_val, src);  // 0 for padding lanes
            if (has) {
                float xv[VEC];
                VT::load(xb + (size_t)row * d + sub * VEC, xv);
#pragma unroll
                for (int i = 0; i < VEC; ++i) acc[i] = fmaf(w, xv[i], acc[i]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
#pragma unroll
        for (int o = L; o < 32; o <<= 1) acc[i] += __shfl_xor_sync(FULL, acc[i], o);
    }
}

constexpr int SPMM_T_HEAVY = 1024;
template <typename T, typename TO, int L, typename VT = Vec16<T>>
__global__ void __launch_bounds__(CSR_WARPS * 32)
spmm2_t_kernel(const int32_t *__restrict__ col_ptr, const int32_t *__restrict__ row_idx, const int32_t *__restrict__ perm,
               const float *__restrict__ values, const T *__restrict__ x, TO *__restrict__ y, int S, int d, int64_t nnz) {
    constexpr int VEC = VT::N;
    __shared__ float s_part[CSR_WARPS][L * VEC];
    __shared__ int s_len[CSR_WARPS];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int blocks_per_head = (S + CSR_WARPS - 1) / CSR_WARPS;
    const int b = blockIdx.x / blocks_per_head, cb = (blockIdx.x % blocks_per_head) * CSR_WARPS;
    const int sub = lane % L;
    const bool has = sub * VEC < d;
    const int32_t *pp = col_ptr + (size_t)b * (S + 1);
    const int32_t *ip = row_idx + (size_t)b * nnz;
    const int32_t *pm = perm + (size_t)b * nnz;
    const float *vp = values + (size_t)b * nnz;
    const T *xb = x + (size_t)b * S * d;
    {
        const int c = cb + wid;
        const int c0 = c < S ? pp[c] : 0, c1 = c < S ? pp[c + 1] : 0;
        if (lane == 0) s_len[wid] = c1 - c0;
        if (c < S && c1 - c0 <= SPMM_T_HEAVY) {
            float acc[VEC];
#pragma unroll
            for (int i = 0; i < VEC; ++i) acc[i] = 0.0f;
            spmm_accumulate<T, L, true, VT>(acc, ip, pm, vp, xb, c0, c1, lane, d, has);
            if (lane < L && has) store_vec<TO, VEC>(y + ((size_t)b * S + c) * d + sub * VEC, acc);
        }
    }
    __syncthreads();
    for (int u = 0; u < CSR_WARPS; ++u) {
        if (s_len[u] <= SPMM_T_HEAVY) continue;                       // block-uniform
        const int c = cb + u;
        const int c0 = pp[c], c1 = pp[c + 1];
        const int per_warp = ((c1 - c0 + CSR_WARPS * 32 - 1) / (CSR_WARPS * 32)) * 32;
        const int e0 = c0 + wid * per_warp, e1 = min(c1, e0 + per_warp);
        float acc[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[i] = 0.0f;
        spmm_accumulate<T, L, true, VT>(acc, ip, pm, vp, xb, e0, e1, lane, d, has);
        if (lane < L) {
#pragma unroll
            for (int i = 0; i < VEC; ++i) s_part[wid][sub * VEC + i] = acc[i];
        }
        __syncthreads();
        if (wid == 0 && lane < L && has) {
            float sum[VEC];
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                float t = 0.0f;
#pragma unroll
                for (int w = 0; w < CSR_WARPS; ++w) t += s_part[w][sub * VEC + i];
                sum[i] = t;
            }
            store_vec<TO, VEC>(y + ((size_t)b * S + c) * d + sub * VEC, sum);
        }
        __syncthreads();
    }
}

// bf16 tensor-core variants (csr_mma.cu)
namespace csr_mma {
bool supported(int d, const void *a, const void *b);
int launch_sddmm(const int32_t *indptr, const int32_t *indices, const __nv_bfloat16 *q, const __nv_bfloat16 *k, float *values,
                 int B, int S, int d, int64_t nnz, float scale, float clamp, cudaStream_t st);
int launch_spmm(bool trans, const int32_t *ptr, const int32_t *src_idx, const int32_t *perm, const float *values,
                const __nv_bfloat16 *x, void *y, bool y_bf16, int B, int S, int d, int64_t nnz, cudaStream_t st);
}  // namespace csr_mma
// dense 64 x 64 tiles from the CSC pattern (csr_dense.cu): the transposed product at attention densities
namespace csr_dense {
bool supported(int d, int S, const void *x, const void *y);
int launch_spmm_t(const int32_t *col_ptr, const int32_t *row_idx, const int32_t *perm, const float *values,
                  const __nv_bfloat16 *x, void *y, bool y_bf16, int B, int S, int d, int64_t nnz, cudaStream_t st);
}  // namespace csr_dense

// ---- launch helpers ------------------------------------------------------------------------------
static inline int lanes_for(int d, int vec) {
    int need = (d + vec - 1) / vec, L = 1;
    while (L < need) L <<= 1;
    return L;
}

template <typename T>
static int launch_sddmm(const int32_t *indptr, const int32_t *indices, const T *q, const T *k, float *values, int B,
                        int S, int d, int64_t nnz, float scale, float clamp, cudaStream_t st) {
    constexpr int VEC = Vec16<T>::N;
    const int64_t rows = (int64_t)B * S;
    if constexpr (sizeof(T) == 2) {
        if (csr_mma::supported(d, q, k) && rows < ((int64_t)1 << 31))
            return csr_mma::launch_sddmm(indptr, indices, q, k, values, B, S, d, nnz, scale, clamp, st);
    }
    const unsigned grid = (unsigned)((rows + CSR_WARPS - 1) / CSR_WARPS);
    if constexpr (sizeof(T) == 4) {
        static const bool narrow = [] { const char *e = getenv("SPT_SDDMM_F32_NARROW"); return e && atoi(e) == 1; }();   // A/B switch
        if (!narrow && d % 8 == 0 && d <= 256 && ((uintptr_t)q % 32 == 0) && ((uintptr_t)k % 32 == 0)) {
            switch (lanes_for(d, 8)) {
                case 1: sddmm_f32w_kernel<1><<<grid, CSR_WARPS * 32, 0, st>>>(indptr, indices, q, k, values, B, S, d, nnz, scale, clamp); break;
                case 2: sddmm_f32w_kernel<2><<<grid, CSR_WARPS * 32, 0, st>>>(indptr, indices, q, k, values, B, S, d, nnz, scale, clamp); break;
                case 4: sddmm_f32w_kernel<4><<<grid, CSR_WARPS * 32, 0, st>>>(indptr, indices, q, k, values, B, S, d, nnz, scale, clamp); break;
                case 8: sddmm_f32w_kernel<8><<<grid, CSR_WARPS * 32, 0, st>>>(indptr, indices, q, k, values, B, S, d, nnz, scale, clamp); break;
                case 16: sddmm_f32w_kernel<16><<<grid, CSR_WARPS * 32, 0, st>>>(indptr, indices, q, k, values, B, S, d, nnz, scale, clamp); break;
                default: sddmm_f32w_kernel<32><<<grid, CSR_WARPS * 32, 0, st>>>(indptr, indices, q, k, values, B, S, d, nnz, scale, clamp); break;
            }
            return after_launch("sddmm_f32w_kernel");
        }
    }
    const bool vec_ok = (d % VEC == 0) && (d <= 32 * VEC) && ((uintptr_t)q % 16 == 0) && ((uintptr_t)k % 16 == 0);
    if (!vec_ok) {
        sddmm_scalar_kernel<T><<<grid, CSR_WARPS * 32, 0, st>>>(indptr, indices, q, k, values, B, S, d, nnz, scale, clamp);
        return after_launch("sddmm_scalar_kernel");
    }
#define SPT_SDDMM_CASE(LL)                                                                                       \
    case LL:                                                                                                     \
        sddmm_kernel<T, LL><<<grid, CSR_WARPS * 32, 0, st>>>(indptr, indices, q, k, values, B, S, d, nnz, scale, clamp); \
        break;
    switch (lanes_for(d, VEC)) {
        SPT_SDDMM_CASE(1)
        SPT_SDDMM_CASE(2)
        SPT_SDDMM_CASE(4)
        SPT_SDDMM_CASE(8)
        SPT_SDDMM_CASE(16)
        SPT_SDDMM_CASE(32)
    }
#undef SPT_SDDMM_CASE
    return after_launch("sddmm_kernel");
}

template <typename T, typename TO, bool TRANS>
static int launch_spmm(const int32_t *ptr, const int32_t *src_idx, const int32_t *perm, const float *values,
                       const T *x, TO *y, int B, int S, int d, int64_t nnz, cudaStream_t st) {
    constexpr int VEC = Vec16<T>::N;
    const int64_t rows = (int64_t)B * S;
    if constexpr (sizeof(T) == 2 && TRANS) {
        if (csr_dense::supported(d, S, x, y) && rows < ((int64_t)1 << 31) && nnz < ((int64_t)1 << 31)) {
            const int rc = csr_dense::launch_spmm_t(ptr, src_idx, perm, values, x, y, sizeof(TO) == 2, B, S, d, nnz, st);
            if (rc != SPT_ERR_UNSUPPORTED) return rc;              // unaligned lists: the gathered kernel below
        }
    }
    if constexpr (sizeof(T) == 2) {
        // the tensor-core spmm is correct but measured slower than the SIMT kernels (its B-fragment layout forces
        // 16-byte loads that coalesce per 32 bytes only: 4 LSU wavefronts per gathered row); opt-in for experiments
        static const bool on = [] { const char *e = getenv("SPT_CSR_MMA_SPMM"); return e && atoi(e) == 1; }();
        if (on && csr_mma::supported(d, x, y) && rows < ((int64_t)1 << 31))
            return csr_mma::launch_spmm(TRANS, ptr, src_idx, perm, values, x, y, sizeof(TO) == 2, B, S, d, nnz, st);
    }
    const unsigned grid = (unsigned)((rows + CSR_WARPS - 1) / CSR_WARPS);
    {
        // 256-bit lane loads when the rows allow it (see Vec32)
        static const bool narrow = [] { const char *e = getenv("SPT_SPMM_NARROW"); return e && atoi(e) == 1; }();   // A/B switch
        constexpr int WV = Vec32<T>::N;
        // measured at the bench shapes: CSR product 0.84 -> 0.73 ms (bf16), 0.33 -> 0.27 (fp32); transposed product
        // 1.10 -> 0.94 ms (fp32) but 1.33 -> 1.52 (bf16: 4 lanes per row leave a 3-step, 16-value reduction per column)
        const bool use_wide = !narrow && (!TRANS || sizeof(T) == 4);
        if (use_wide && d % WV == 0 && d <= 16 * WV && ((uintptr_t)x % 32 == 0) && ((uintptr_t)y % 16 == 0) && rows < ((int64_t)1 << 31)) {
            const unsigned tgrid = (unsigned)(B * ((S + CSR_WARPS - 1) / CSR_WARPS));
#define SPT_SPMM_WIDE(LL)                                                                                        \
    case LL:                                                                                                     \
        if (TRANS)                                                                                               \
            spmm2_t_kernel<T, TO, LL, Vec32<T>><<<tgrid, CSR_WARPS * 32, 0, st>>>(ptr, src_idx, perm, values, x, y, S, d, nnz); \
        else                                                                                                     \
            spmm_kernel<T, TO, LL, false, Vec32<T>><<<grid, CSR_WARPS * 32, 0, st>>>(ptr, src_idx, perm, values, x, y, B, S, d, nnz); \
        break;
            switch (lanes_for(d, WV)) {
                SPT_SPMM_WIDE(1)
                SPT_SPMM_WIDE(2)
                SPT_SPMM_WIDE(4)
                SPT_SPMM_WIDE(8)
                SPT_SPMM_WIDE(16)
            }
#undef SPT_SPMM_WIDE
            return after_launch("spmm_kernel(wide)");
        }
    }
    const bool vec_ok = (d % VEC == 0) && (d <= 32 * VEC) && ((uintptr_t)x % 16 == 0) && ((uintptr_t)y % 16 == 0);
    if (!vec_ok) {
        spmm_scalar_kernel<T, TO, TRANS><<<grid, CSR_WARPS * 32, 0, st>>>(ptr, src_idx, perm, values, x, y, B, S, d, nnz);
        return after_launch("spmm_scalar_kernel");
    }
    static const bool v1 = [] { const char *e = getenv("SPT_SPMM_V1"); return e && atoi(e) == 1; }();   // A/B switch
#define SPT_SPMM_CASE(LL)                                                                                        \
    case LL:                                                                                                     \
        if (TRANS && !v1 && rows < ((int64_t)1 << 31))                                                           \
            spmm2_t_kernel<T, TO, LL><<<(unsigned)(B * ((S + CSR_WARPS - 1) / CSR_WARPS)), CSR_WARPS * 32, 0, st>>>( \
                ptr, src_idx, perm, values, x, y, S, d, nnz);                                                    \
        else if (TRANS && rows < ((int64_t)1 << 31))                                                             \
            spmm_t_block_kernel<T, TO, LL><<<(unsigned)rows, CSR_WARPS * 32, 0, st>>>(ptr, src_idx, perm, values, x, y, \
                                                                                      B, S, d, nnz);            \
        else                                                                                                     \
            spmm_kernel<T, TO, LL, TRANS><<<grid, CSR_WARPS * 32, 0, st>>>(ptr, src_idx, perm, values, x, y, B, S, d, nnz); \
        break;
    switch (lanes_for(d, VEC)) {
        SPT_SPMM_CASE(1)
        SPT_SPMM_CASE(2)
        SPT_SPMM_CASE(4)
        SPT_SPMM_CASE(8)
        SPT_SPMM_CASE(16)
        SPT_SPMM_CASE(32)
    }
#undef SPT_SPMM_CASE
    return after_launch("spmm_kernel");
}

template <bool TRANS>
static int dispatch_spmm(const int32_t *ptr, const int32_t *src_idx, const int32_t *perm, const float *values,
                         const void *x, void *y, int B, int S, int d, int64_t nnz, int dtype, int out_dtype,
                         cudaStream_t st) {
    using bf16 = __nv_bfloat16;
    if (dtype == SPT_F32 && out_dtype == SPT_F32)
        return launch_spmm<float, float, TRANS>(ptr, src_idx, perm, values, (const float *)x, (float *)y, B, S, d, nnz, st);
    if (dtype == SPT_BF16 && out_dtype == SPT_F32)
        return launch_spmm<bf16, float, TRANS>(ptr, src_idx, perm, values, (const bf16 *)x, (float *)y, B, S, d, nnz, st);
    if (dtype == SPT_BF16 && out_dtype == SPT_BF16)
        return launch_spmm<bf16, bf16, TRANS>(ptr, src_idx, perm, values, (const bf16 *)x, (bf16 *)y, B, S, d, nnz, st);
    return fail(SPT_ERR_UNSUPPORTED, "spmm: dtype combination (%d -> %d) not supported", dtype, out_dtype);
}

static int c2c_warps_per_block(int S) {
    int w = (int)((160 * 1024) / ((size_t)S * 4));
    return w < 1 ? 1 : (w > 4 ? 4 : w);
}

}  // namespace spt

using namespace spt;

#define SPT_CHECK_CSR(name)                                                                              \
    SPT_REQUIRE(B >= 1 && S >= 1 && nnz >= 0, name ": bad sizes B=%d S=%d nnz=%lld", B, S, (long long)nnz); \
    SPT_REQUIRE((int64_t)B * S < ((int64_t)1 << 31) * CSR_WARPS, name ": B*S too large")

extern "C" int spt_sddmm_fwd(const int32_t *indptr, const int32_t *indices, const void *query, const void *key,
                             float *values, int B, int S, int d, int64_t nnz, float scale, float clamp, int dtype,
                             spt_stream_t stream) {
    SPT_REQUIRE(indptr && indices && query && key && values, "sddmm_fwd: null pointer");
    SPT_CHECK_CSR("sddmm_fwd");
    SPT_REQUIRE(d >= 1, "sddmm_fwd: bad head dim %d", d);
    if (nnz == 0) return SPT_OK;
    if (dtype == SPT_F32)
        return launch_sddmm(indptr, indices, (const float *)query, (const float *)key, values, B, S, d, nnz, scale, clamp, as_stream(stream));
    if (dtype == SPT_BF16)
        return launch_sddmm(indptr, indices, (const __nv_bfloat16 *)query, (const __nv_bfloat16 *)key, values, B, S, d, nnz, scale, clamp, as_stream(stream));
    return fail(SPT_ERR_INVALID_ARGUMENT, "sddmm_fwd: unknown dtype %d", dtype);
}

// gradient of  v = clamp(scale * raw, -clamp, clamp)  w.r.t. raw, from the clamped values:  scale * g where |v| < clamp, else 0
// (torch's clamp backward, attention.py:125-127 via autograd); one streaming pass instead of five elementwise kernels
__global__ void __launch_bounds__(256)
clamp_scale_bwd_kernel(const float4 *__restrict__ grad, const float4 *__restrict__ clamped, float4 *__restrict__ out,
                       int64_t n4, float scale, float clamp) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 g = grad[i], v = clamped[i];
        float4 o;
        o.x = (clamp <= 0.0f || fabsf(v.x) < clamp) ? g.x * scale : 0.0f;
        o.y = (clamp <= 0.0f || fabsf(v.y) < clamp) ? g.y * scale : 0.0f;
        o.z = (clamp <= 0.0f || fabsf(v.z) < clamp) ? g.z * scale : 0.0f;
        o.w = (clamp <= 0.0f || fabsf(v.w) < clamp) ? g.w * scale : 0.0f;
        out[i] = o;
    }
}

extern "C" int spt_clamp_scale_bwd(const float *grad, const float *clamped, float *out, int64_t n, float scale, float clamp,
                                   spt_stream_t stream) {
    SPT_REQUIRE(grad && clamped && out, "clamp_scale_bwd: null pointer");
    SPT_REQUIRE(n >= 0 && n % 4 == 0, "clamp_scale_bwd: element count must be a multiple of 4 (got %lld)", (long long)n);
    SPT_REQUIRE(((uintptr_t)grad | (uintptr_t)clamped | (uintptr_t)out) % 16 == 0, "clamp_scale_bwd: operands must be 16-byte aligned");
    if (n == 0) return SPT_OK;
    const int64_t n4 = n / 4;
    const int64_t want = (n4 + 255) / 256;
    const int grid = (int)(want < (int64_t)num_sms() * 16 ? want : (int64_t)num_sms() * 16);
    clamp_scale_bwd_kernel<<<grid, 256, 0, as_stream(stream)>>>((const float4 *)grad, (const float4 *)clamped, (float4 *)out, n4,
                                                              scale, clamp);
    return after_launch("clamp_scale_bwd_kernel");
}

extern "C" int spt_spmm_fwd(const int32_t *indptr, const int32_t *indices, const float *values, const void *x, void *y,
                            int B, int S, int d, int64_t nnz, int dtype, int out_dtype, spt_stream_t stream) {
    SPT_REQUIRE(indptr && indices && values && x && y, "spmm_fwd: null pointer");
    SPT_CHECK_CSR("spmm_fwd");
    SPT_REQUIRE(d >= 1, "spmm_fwd: bad feature dim %d", d);
    return dispatch_spmm<false>(indptr, indices, nullptr, values, x, y, B, S, d, nnz, dtype, out_dtype, as_stream(stream));
}

extern "C" int spt_spmm_t_fwd(const int32_t *col_ptr, const int32_t *row_idx, const int32_t *perm, const float *values,
                              const void *x, void *y, int B, int S, int d, int64_t nnz, int dtype, int out_dtype,
                              spt_stream_t stream) {
    SPT_REQUIRE(col_ptr && row_idx && perm && values && x && y, "spmm_t_fwd: null pointer");
    SPT_CHECK_CSR("spmm_t_fwd");
    SPT_REQUIRE(d >= 1, "spmm_t_fwd: bad feature dim %d", d);
    return dispatch_spmm<true>(col_ptr, row_idx, perm, values, x, y, B, S, d, nnz, dtype, out_dtype, as_stream(stream));
}

extern "C" int spt_softmax_fwd(const int32_t *indptr, const int32_t *indices, const float *values, float *output, int B,
                               int S, int64_t nnz, spt_stream_t stream) {
    SPT_REQUIRE(indptr && indices && values && output, "softmax_fwd: null pointer");
    SPT_CHECK_CSR("softmax_fwd");
    if (nnz == 0) return SPT_OK;
    const int64_t rows = (int64_t)B * S;
    softmax_fwd_kernel<<<(unsigned)((rows + CSR_WARPS - 1) / CSR_WARPS), CSR_WARPS * 32, 0, as_stream(stream)>>>(
        indptr, indices, values, output, B, S, nnz);
    return after_launch("softmax_fwd_kernel");
}

extern "C" int spt_softmax_bwd_ex(const int32_t *indptr, const int32_t *indices, const float *output,
                                  const float *grad_output, float *grad_values, int B, int S, int64_t nnz,
                                  int reference_clamp, spt_stream_t stream) {
    SPT_REQUIRE(indptr && indices && output && grad_output && grad_values, "softmax_bwd: null pointer");
    SPT_CHECK_CSR("softmax_bwd");
    if (nnz == 0) return SPT_OK;
    const int64_t rows = (int64_t)B * S;
    softmax_bwd_kernel<<<(unsigned)((rows + CSR_WARPS - 1) / CSR_WARPS), CSR_WARPS * 32, 0, as_stream(stream)>>>(
        indptr, indices, output, grad_output, grad_values, B, S, nnz, reference_clamp, nullptr, 1.0f, 0.0f);
    return after_launch("softmax_bwd_kernel");
}

extern "C" int spt_softmax_clamp_bwd(const int32_t *indptr, const int32_t *indices, const float *output,
                                     const float *grad_output, const float *clamped, float *grad_raw, int B, int S,
                                     int64_t nnz, float scale, float clamp, int reference_clamp, spt_stream_t stream) {
    SPT_REQUIRE(indptr && indices && output && grad_output && clamped && grad_raw, "softmax_clamp_bwd: null pointer");
    SPT_CHECK_CSR("softmax_clamp_bwd");
    if (nnz == 0) return SPT_OK;
    const int64_t rows = (int64_t)B * S;
    softmax_bwd_kernel<<<(unsigned)((rows + CSR_WARPS - 1) / CSR_WARPS), CSR_WARPS * 32, 0, as_stream(stream)>>>(
        indptr, indices, output, grad_output, grad_raw, B, S, nnz, reference_clamp, clamped, scale, clamp);
    return after_launch("softmax_bwd_kernel");
}

extern "C" int spt_softmax_bwd(const int32_t *indptr, const int32_t *indices, const float *output,
                               const float *grad_output, float *grad_values, int B, int S, int64_t nnz,
                               spt_stream_t stream) {
    return spt_softmax_bwd_ex(indptr, indices, output, grad_output, grad_values, B, S, nnz, 0, stream);
}

extern "C" size_t spt_csr2csc_workspace_bytes(int B, int S, int64_t nnz) {
    (void)nnz;
    const size_t n_tiles = (size_t)(S + C2C_TR - 1) / C2C_TR;
    return (size_t)B * n_tiles * S * sizeof(int32_t) + 16;     // tile counts + the placement kernels' fallback flag
}

extern "C" int spt_csr2csc(const int32_t *indptr, const int32_t *indices, int32_t *col_ptr, int32_t *row_idx,
                           int32_t *perm, void *workspace, int B, int S, int64_t nnz, spt_stream_t stream) {
    SPT_REQUIRE(indptr && indices && col_ptr && row_idx && perm && workspace, "csr2csc: null pointer");
    SPT_CHECK_CSR("csr2csc");
    SPT_REQUIRE(B <= 65535, "csr2csc: batch %d exceeds grid limit", B);
    SPT_REQUIRE((size_t)S * 4 <= 200 * 1024, "csr2csc: S=%d too large for the shared-memory column cursors", S);
    SPT_REQUIRE(nnz < ((int64_t)1 << 31), "csr2csc: nnz too large");
    cudaStream_t st = as_stream(stream);
    const int n_tiles = (S + C2C_TR - 1) / C2C_TR;
    int32_t *tile_cnt = (int32_t *)workspace;
    const size_t smem_s = (size_t)S * 4;
    if (smem_s > 48 * 1024) {
        cudaFuncSetAttribute(csr2csc_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s);
        cudaFuncSetAttribute(csr2csc_colscan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s);
    }
    csr2csc_count_kernel<<<dim3(n_tiles, B), 256, smem_s, st>>>(indptr, indices, tile_cnt, S, nnz, n_tiles);
    SPT_LAUNCH_CHECK("csr2csc_count_kernel");
    csr2csc_tilescan_kernel<<<dim3((S + 255) / 256, B), 256, 0, st>>>(tile_cnt, col_ptr, S, n_tiles);
    SPT_LAUNCH_CHECK("csr2csc_tilescan_kernel");
    csr2csc_colscan_kernel<<<B, 1024, smem_s, st>>>(col_ptr, S);
    SPT_LAUNCH_CHECK("csr2csc_colscan_kernel");
    // staged placement when its shared-memory arrays fit (S <= 2048 with 16384-entry chunks, S <= 4096 with 4096)
    const int cap = S <= 2048 ? 16384 : 4096;
    const size_t smem_st = (size_t)S * 8 + (size_t)cap * 4 + (size_t)C2C_WARPS * S * 4 + (size_t)cap * 2 + (size_t)cap +
                           (size_t)C2C_WARPS * S;
    if (smem_st <= 220 * 1024) {
        // bit-matrix placement first (S <= 2048); it raises the flag behind the tile counts when the pattern is not its
        // kind, and the staged kernel then redoes the call (it returns at once otherwise)
        static const bool no_bits = [] { const char *e = getenv("SPT_CSR2CSC_BITS"); return e && atoi(e) == 0; }();   // A/B switch
        const size_t smem_b = (size_t)S * 12 + (size_t)C2B_SLOTS * 4 + (size_t)C2C_TR * S;
        int *flag = nullptr;
        if (!no_bits && S <= 2048 && S % 4 == 0 && smem_b <= 220 * 1024) {
            flag = reinterpret_cast<int *>(reinterpret_cast<char *>(workspace) + (size_t)B * n_tiles * S * sizeof(int32_t));
            cudaError_t e = cudaMemsetAsync(flag, 0, 16, st);
            if (e != cudaSuccess) return fail(SPT_ERR_CUDA, "csr2csc: memset: %s", cudaGetErrorString(e));
            cudaFuncSetAttribute(csr2csc_place_bits_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b);
            csr2csc_place_bits_kernel<<<dim3(n_tiles, B), C2B_THREADS, smem_b, st>>>(indptr, indices, tile_cnt, col_ptr, row_idx,
                                                                                   perm, S, nnz, n_tiles, flag);
            SPT_LAUNCH_CHECK("csr2csc_place_bits_kernel");
        }
        cudaFuncSetAttribute(csr2csc_place_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_st);
        csr2csc_place_staged_kernel<<<dim3(n_tiles, B), C2C_WARPS * 32, smem_st, st>>>(indptr, indices, tile_cnt, col_ptr,
                                                                                      row_idx, perm, S, nnz, n_tiles, cap, flag);
        SPT_LAUNCH_CHECK("csr2csc_place_staged_kernel");
        return SPT_OK;
    }
    const int nw = c2c_warps_per_block(S);
    const size_t smem_p = smem_s * nw;
    if (smem_p > 48 * 1024)
        cudaFuncSetAttribute(csr2csc_place_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_p);
    csr2csc_place_kernel<<<dim3((n_tiles + nw - 1) / nw, B), nw * 32, smem_p, st>>>(indptr, indices, tile_cnt, col_ptr,
                                                                                   row_idx, perm, S, nnz, n_tiles);
    SPT_LAUNCH_CHECK("csr2csc_place_kernel");
    return SPT_OK;
}
