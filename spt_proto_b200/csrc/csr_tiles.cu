// Tile index of a batched CSR pattern + the transposed sparse product on it (bf16 x, head dim 64 / 128) — reference call
// sites extension/spmm.cpp:27-69 with trans_lhs (dV = P^T dO and dK = dS^T Q of the backward passes, kernels/spmm.py:42-47,
// kernels/sddmm.py:44-49).
//
// The dense-tile transposed product (csr_dense.cu) only ever asks one question of the CSC: "which entries fall into column
// tile ct and row chunk rc?".  A full CSR -> CSC transposition answers much more than that (a stable sort by column: 1.15 ms
// at the bench shape, more than any of the products) and its answer has to be dug out again by cursors walking 64 sorted
// column lists.  The tile index stores exactly the answer:
//     tile_ptr[b][ct * n_rc + rc]  ->  the bucket of entries of head b with column in [64 ct, 64 ct + 64) and row in
//                                      [64 rc, 64 rc + 64)   (ct-major: one column tile's chunks are contiguous)
//     tile_ent[b][slot]            =   c_local | r_local << 6 | e << 12     (e = position of the entry in the head's CSR)
// one 32-bit word per entry (the CSC needs 8 bytes) — which caps a head at 2^20 entries.  It is built by a counting pass and
// a placing pass over the indices, each block owning one 64-row chunk (histogram over at most 128 column tiles in shared
// memory, `__match_any_sync` hands a warp's entries of one tile consecutive slots, so the duplicates of a row — the
// lookup's zero padding, all on column 0 — stay adjacent).  The order of the entries inside a bucket is the order of the
// shared-memory atomics: run-to-run it can differ, the SET cannot.
//
// The product: a block owns a column tile of one head and walks its non-empty buckets; bucket entries are read with
// coalesced loads (no cursors, no sorted lists), their values gathered through e, summed over runs of adjacent equal cells
// with warp shuffles (fp32 shared-memory atomics are CAS loops: the padding runs would serialise them) and added into a
// 64 x 64 fp32 tile; tile (bf16 hi + lo) x the staged x chunk runs on mma.sync as in csr_dense.cu.  Tiles and x chunks are
// double-buffered (two block barriers per chunk) and the next bucket's entries + values are fetched before the MMAs of the
// current one.
#include <climits>
#include <cstdlib>

#include "common.cuh"

namespace spt {
namespace csr_tiles {

using bf16 = __nv_bfloat16;

constexpr int DT = 64;
constexpr int THREADS = 256;
constexpr int PS = 72;                 // fp32 tile row stride (see csr_dense.cu)
constexpr int MAX_CT = 128;            // S <= 8192
constexpr int PRE = 4;                 // bucket entries per thread fetched ahead (PRE * THREADS = 1024)

// ---- index build -----------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(THREADS)
tiles_count_kernel(const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices, int32_t *__restrict__ tile_ptr,
                   int S, int64_t nnz, int n_ct, int n_rc) {
    __shared__ int hist[THREADS / 32][MAX_CT];
    const int rc = blockIdx.x, b = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (THREADS / 32) * MAX_CT; i += THREADS) (&hist[0][0])[i] = 0;
    __syncthreads();
    const int32_t *ix = indices + (size_t)b * nnz;
    for (int r = rc * DT + warp; r < min(S, rc * DT + DT); r += THREADS / 32) {
        const int e0 = indptr[r], e1 = indptr[r + 1];
        for (int e = e0 + lane; e < e1; e += 32) {
            const int c = ix[e];
            if ((unsigned)c < (unsigned)S) atomicAdd(&hist[warp][c >> 6], 1);
        }
    }
    __syncthreads();
    for (int ct = threadIdx.x; ct < n_ct; ct += THREADS) {
        int n = 0;
#pragma unroll
        for (int w = 0; w < THREADS / 32; ++w) n += hist[w][ct];
        tile_ptr[(size_t)b * (n_ct * n_rc + 1) + (size_t)ct * n_rc + rc] = n;
    }
}

// in-place exclusive scan of a head's n = n_ct * n_rc counts; tile_ptr[n] = total
__global__ void __launch_bounds__(1024)
tiles_scan_kernel(int32_t *__restrict__ tile_ptr, int n) {
    __shared__ int warp_sum[32];
    int32_t *p = tile_ptr + (size_t)blockIdx.x * (n + 1);
    const int per = (n + 1023) / 1024;
    const int i0 = threadIdx.x * per, i1 = min(n, i0 + per);
    int sum = 0;
    for (int i = i0; i < i1; ++i) sum += p[i];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += up;
    }
    if (lane == 31) warp_sum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = warp_sum[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int up = __shfl_up_sync(0xffffffffu, w, d);
            if (lane >= d) w += up;
        }
        warp_sum[lane] = w;
    }
    __syncthreads();
    int run = inc - sum + (warp ? warp_sum[warp - 1] : 0);     // exclusive prefix of this thread's segment
    for (int i = i0; i < i1; ++i) {
        const int c = p[i];
        p[i] = run;
        run += c;
    }
    if (threadIdx.x == 1023) p[n] = warp_sum[31];
}

__global__ void __launch_bounds__(THREADS)
tiles_place_kernel(const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices, const int32_t *__restrict__ tile_ptr,
                   uint32_t *__restrict__ tile_ent, int S, int64_t nnz, int n_ct, int n_rc) {
    __shared__ int cursor[MAX_CT];
    const int rc = blockIdx.x, b = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int32_t *tp = tile_ptr + (size_t)b * (n_ct * n_rc + 1);
    for (int ct = threadIdx.x; ct < n_ct; ct += THREADS) cursor[ct] = tp[(size_t)ct * n_rc + rc];
    __syncthreads();
    const int32_t *ix = indices + (size_t)b * nnz;
    uint32_t *out = tile_ent + (size_t)b * nnz;
    const uint32_t lt = (1u << lane) - 1;
    for (int r = rc * DT + warp; r < min(S, rc * DT + DT); r += THREADS / 32) {
        const int e0 = indptr[r], e1 = indptr[r + 1];
        for (int base = e0; base < e1; base += 4 * 32) {              // warp-uniform trip count; four index loads in flight
            int cs[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) cs[j] = base + j * 32 + lane < e1 ? ix[base + j * 32 + lane] : -1;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (base + j * 32 >= e1) break;
                const int e = base + j * 32 + lane, c = cs[j];
                const bool ok = (unsigned)c < (unsigned)S;
                const int key = ok ? (c >> 6) : (MAX_CT + lane);       // lanes without an entry: a group of their own
                const uint32_t peers = __match_any_sync(0xffffffffu, key);
                const int leader = __ffs(peers) - 1;
                int slot = 0;
                if (ok && lane == leader) slot = atomicAdd(&cursor[key], __popc(peers));
                slot = __shfl_sync(0xffffffffu, slot, leader) + __popc(peers & lt);
                if (ok) out[slot] = (uint32_t)(c & 63) | ((uint32_t)(r - rc * DT) << 6) | ((uint32_t)e << 12);
            }
        }
    }
}

// ---- the product -----------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, bool valid) {   // !valid: zero fill
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void split2(float2 v, uint32_t &hi, uint32_t &lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(v.x, v.y);
    const __nv_bfloat162 l = __floats2bfloat162_rn(v.x - __low2float(h), v.y - __high2float(h));
    hi = *reinterpret_cast<const uint32_t *>(&h);
    lo = *reinterpret_cast<const uint32_t *>(&l);
}
template <typename TO>
__device__ __forceinline__ void store2(TO *p, float a, float b);
template <>
__device__ __forceinline__ void store2<float>(float *p, float a, float b) { *reinterpret_cast<float2 *>(p) = make_float2(a, b); }
template <>
__device__ __forceinline__ void store2<bf16>(bf16 *p, float a, float b) { *reinterpret_cast<__nv_bfloat162 *>(p) = __floats2bfloat162_rn(a, b); }

// one warp-wide batch of 32 consecutive bucket entries -> the tile.  cell = c_local | r_local << 6 (or -1: no entry).
template <bool TRANS>
__device__ __forceinline__ void scatter_batch(float *P, int cell, float v, int lane) {
    const int prev = __shfl_up_sync(0xffffffffu, cell, 1), next = __shfl_down_sync(0xffffffffu, cell, 1);
    const bool head = lane == 0 || prev != cell;
    if (__any_sync(0xffffffffu, !head && cell >= 0)) {               // runs of equal cells (duplicates): sum them first
        int start = head ? lane : 0;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) start = max(start, __shfl_up_sync(0xffffffffu, start, d) * (lane >= d));
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const float up = __shfl_up_sync(0xffffffffu, v, d);
            if (lane - d >= start) v += up;
        }
        if (lane != 31 && next == cell) cell = -1;                   // only the last lane of a run adds
    }
    // tile row = the owner's index (column for the transposed product, row for the direct one), tile column = contraction
    if (cell >= 0) atomicAdd(P + (TRANS ? (cell & 63) * PS + (cell >> 6) : (cell >> 6) * PS + (cell & 63)), v);
}

// TRANS: y[c] = sum_r A[r, c] x[r] (owner = column tile, walks row chunks); !TRANS: y[r] = sum_c A[r, c] x[c] (owner = row
// tile, walks column chunks).  Same index, same buckets: only the walk and the orientation of the tile differ.
template <int D, typename TO, bool TRANS>
__global__ void __launch_bounds__(THREADS, 4)
spmm_tiles_kernel(const int32_t *__restrict__ tile_ptr, const uint32_t *__restrict__ tile_ent, const float *__restrict__ values,
                  const bf16 *__restrict__ x, TO *__restrict__ y, int B, int S, int64_t nnz, int n_ct, int n_rc) {
    constexpr int XS = (D + 8) * 2;            // bytes per staged x row
    constexpr int NT = D / 16;
    extern __shared__ __align__(16) unsigned char smem[];
    float *P0 = reinterpret_cast<float *>(smem);                                   // [2][64 columns][PS]
    unsigned char *X0 = smem + 2 * DT * PS * 4;                                    // [2][64 rows][XS]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // the heaviest owner of every head first: under a causal pattern column tile 0 / the last row tile touch every chunk
    const int b = blockIdx.x % B, ct = TRANS ? blockIdx.x / B : n_rc - 1 - blockIdx.x / B;       // ct = the owner's tile index
    const int32_t *tp = tile_ptr + (size_t)b * (n_ct * n_rc + 1) + (TRANS ? (size_t)ct * n_rc : (size_t)ct);
    constexpr int one = 1;
    const int tstep = TRANS ? one : n_rc;                          // bucket (owner, other) -> tp[other * tstep]
    const uint32_t *ent = tile_ent + (size_t)b * nnz;
    const float *vp = values + (size_t)b * nnz;
    const bf16 *xb = x + (size_t)b * S * D;
    const uint32_t x_s = (uint32_t)__cvta_generic_to_shared(X0);

    const int g = lane >> 2, t = lane & 3;
    const int mt = warp & 3, nh = warp >> 2;
    float acc[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[j][i] = 0.0f;
    const uint32_t b_off = ((lane & 7) + ((lane >> 3) & 1) * 8) * XS + (nh * (D / 2) + (lane >> 4) * 8) * 2;

    auto clear_tile = [&](float *P) {
        float4 *p4 = reinterpret_cast<float4 *>(P);
        for (int i = tid; i < DT * PS / 4; i += THREADS) p4[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    };
    auto load_x = [&](int rc, int buf) {
        const int r0 = rc * DT;
        for (int i = tid; i < DT * (D / 8); i += THREADS) {
            const int rr = i / (D / 8), ch = i % (D / 8);
            const bool ok = r0 + rr < S;
            cp_async16(x_s + buf * DT * XS + rr * XS + ch * 16, xb + (size_t)(ok ? r0 + rr : 0) * D + ch * 8, ok);
        }
    };
    // next non-empty bucket at or after rc
    auto next_chunk = [&](int rc) {
        while (rc < n_rc && tp[(size_t)rc * tstep + 1] == tp[(size_t)rc * tstep]) ++rc;
        return rc;
    };
    int cell_n[PRE];
    float val_n[PRE];
    auto prefetch = [&](int rc) {                                     // the first PRE * THREADS entries of bucket rc
        const int p0 = tp[(size_t)rc * tstep], p1 = tp[(size_t)rc * tstep + 1];
#pragma unroll
        for (int j = 0; j < PRE; ++j) {
            const int q = p0 + j * THREADS + tid;
            const uint32_t w = q < p1 ? ent[q] : 0xffffffffu;
            cell_n[j] = q < p1 ? (int)(w & 0xfffu) : -1;
            val_n[j] = q < p1 ? vp[w >> 12] : 0.0f;
        }
    };

    int rc = next_chunk(0), it = 0;
    if (rc < n_rc) {
        clear_tile(P0);
        load_x(rc, 0);
        prefetch(rc);
    }
    __syncthreads();
    for (; rc < n_rc; ++it) {
        const int buf = it & 1;
        float *P = P0 + buf * DT * PS;
        // (A) this bucket -> tile[buf]
        {
            const int p0 = tp[(size_t)rc * tstep], p1 = tp[(size_t)rc * tstep + 1];
#pragma unroll
            for (int j = 0; j < PRE; ++j)
                if (p0 + j * THREADS + (tid & ~31) < p1) scatter_batch<TRANS>(P, cell_n[j], val_n[j], lane);      // warp-uniform
            // long buckets (the padding column's early chunks hold 16 k entries): four independent entry -> value chains
            // per thread and trip (one chain at a time made the first column tile's block the critical path of the kernel)
            for (int q0 = p0 + PRE * THREADS + (tid & ~31); q0 < p1; q0 += 4 * THREADS) {
                uint32_t w[4];
                float v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int q = q0 + j * THREADS + lane;
                    w[j] = q < p1 ? ent[q] : 0xffffffffu;
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] = w[j] != 0xffffffffu ? vp[w[j] >> 12] : 0.0f;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (q0 + j * THREADS < p1) scatter_batch<TRANS>(P, w[j] != 0xffffffffu ? (int)(w[j] & 0xfffu) : -1, v[j], lane);
            }
        }
        const int rc_next = next_chunk(rc + 1);
        if (rc_next < n_rc) prefetch(rc_next);                        // lands under the MMAs below
        cp_async_wait_all();
        __syncthreads();
        // (B) acc += tile[buf] (bf16 hi + lo) x chunk[buf]; the other tile is cleared, the next chunk's rows requested
        if (rc_next < n_rc) {
            clear_tile(P0 + (buf ^ 1) * DT * PS);
            load_x(rc_next, buf ^ 1);
        }
        const uint32_t b_addr = x_s + buf * DT * XS + b_off;
#pragma unroll
        for (int ks = 0; ks < DT / 16; ++ks) {
            const float *pa = P + (mt * 16 + g) * PS + ks * 16 + 2 * t;
            uint32_t ah[4], al[4];
            split2(*reinterpret_cast<const float2 *>(pa), ah[0], al[0]);
            split2(*reinterpret_cast<const float2 *>(pa + 8 * PS), ah[1], al[1]);
            split2(*reinterpret_cast<const float2 *>(pa + 8), ah[2], al[2]);
            split2(*reinterpret_cast<const float2 *>(pa + 8 * PS + 8), ah[3], al[3]);
#pragma unroll
            for (int jp = 0; jp < NT / 2; ++jp) {
                uint32_t bb[4];
                ldmatrix_x4_trans(b_addr + ks * 16 * XS + jp * 32, bb);
                mma16816(acc[2 * jp], ah, bb[0], bb[1]);
                mma16816(acc[2 * jp], al, bb[0], bb[1]);
                mma16816(acc[2 * jp + 1], ah, bb[2], bb[3]);
                mma16816(acc[2 * jp + 1], al, bb[2], bb[3]);
            }
        }
        __syncthreads();
        rc = rc_next;
    }
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int c = ct * DT + mt * 16 + g + half * 8;
        if (c < S) {
            TO *dst = y + ((size_t)b * S + c) * D + nh * (D / 2) + 2 * t;
#pragma unroll
            for (int j = 0; j < NT; ++j) store2<TO>(dst + 8 * j, acc[j][2 * half], acc[j][2 * half + 1]);
        }
    }
}

template <int D, typename TO, bool TRANS>
static int launch_dt(const int32_t *tile_ptr, const uint32_t *tile_ent, const float *values, const bf16 *x, TO *y, int B, int S,
                     int64_t nnz, cudaStream_t st) {
    const int n_ct = (S + DT - 1) / DT, n_rc = n_ct;
    const size_t smem = 2 * (size_t)DT * PS * 4 + 2 * (size_t)DT * (D + 8) * 2;
    cudaFuncSetAttribute(spmm_tiles_kernel<D, TO, TRANS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    spmm_tiles_kernel<D, TO, TRANS><<<(unsigned)((int64_t)B * n_ct), THREADS, smem, st>>>(tile_ptr, tile_ent, values, x, y, B, S,
                                                                                        nnz, n_ct, n_rc);
    return after_launch("spmm_tiles_kernel");
}
template <int D, typename TO>
static int launch_d(bool trans, const int32_t *tile_ptr, const uint32_t *tile_ent, const float *values, const bf16 *x, TO *y,
                    int B, int S, int64_t nnz, cudaStream_t st) {
    return trans ? launch_dt<D, TO, true>(tile_ptr, tile_ent, values, x, y, B, S, nnz, st)
                 : launch_dt<D, TO, false>(tile_ptr, tile_ent, values, x, y, B, S, nnz, st);
}


// ---- sddmm on the tile index -------------------------------------------------------------------------------------------
// values[e] = clamp(scale <q[row(e)], k[col(e)]>).  A block owns a 64-row tile of one head (its Q rows stay in shared
// memory, their A fragments in registers for head dim 64) and walks the non-empty buckets of its row: S = Q K_chunk^T on
// mma.sync (bf16 operands are exact, fp32 accumulation) into a double-buffered fp32 tile, then every entry of the bucket
// picks its cell and stores it at its CSR position.  The gathered kernel (csr_mma.cu) moves a 128-byte key row per entry
// through L2; this one reads a 4-byte index word and writes 4 bytes per entry — the key chunk is staged once per 64 rows.
// One block barrier per chunk.
__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}

template <int D>
__global__ void __launch_bounds__(THREADS, D == 64 ? 3 : 2)
sddmm_tiles_kernel(const int32_t *__restrict__ tile_ptr, const uint32_t *__restrict__ tile_ent, const bf16 *__restrict__ q,
                   const bf16 *__restrict__ k, float *__restrict__ values, int B, int S, int64_t nnz, int n_ct, int n_rc,
                   float scale, float clamp) {
    constexpr int XS = (D + 8) * 2;            // bytes per staged row
    constexpr int KS = D / 16;                 // k-steps
    constexpr bool HOLD_A = D == 64;           // Q fragments live in registers across chunks
    extern __shared__ __align__(16) unsigned char smem[];
    unsigned char *Q0 = smem;                                                        // [64 rows][XS]
    unsigned char *K0 = smem + DT * XS;                                              // [2][64 keys][XS]
    float *S0 = reinterpret_cast<float *>(smem + 3 * DT * XS);                       // [2][64 rows][PS]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x % B, rt = n_rc - 1 - blockIdx.x / B;   // under a causal pattern the last row tile is the heaviest
    const int32_t *tp = tile_ptr + (size_t)b * (n_ct * n_rc + 1) + rt;               // bucket (ct, rt) -> tp[ct * n_rc]
    const uint32_t *ent = tile_ent + (size_t)b * nnz;
    float *vp = values + (size_t)b * nnz;
    const bf16 *qb = q + (size_t)b * S * D, *kb = k + (size_t)b * S * D;
    const uint32_t q_s = (uint32_t)__cvta_generic_to_shared(Q0), k_s = (uint32_t)__cvta_generic_to_shared(K0);

    auto load_rows = [&](uint32_t dst, const bf16 *src, int r0) {                     // 64 rows -> shared memory, zero beyond S
        for (int i = tid; i < DT * (D / 8); i += THREADS) {
            const int rr = i / (D / 8), ch = i % (D / 8);
            const bool ok = r0 + rr < S;
            cp_async16(dst + rr * XS + ch * 16, src + (size_t)(ok ? r0 + rr : 0) * D + ch * 8, ok);
        }
    };
    auto next_chunk = [&](int ct) {
        while (ct < n_ct && tp[(size_t)ct * n_rc + 1] == tp[(size_t)ct * n_rc]) ++ct;
        return ct;
    };
    const int g = lane >> 2, t = lane & 3;
    const int mt = warp & 3, nh = warp >> 2;                                         // 16 rows x 32 columns of the tile per warp
    const uint32_t a_addr = q_s + (mt * 16 + (lane & 15)) * XS + (lane >> 4) * 16;
    const uint32_t b_off = (nh * 32 + (lane & 7) + ((lane >> 4) & 1) * 8) * XS + ((lane >> 3) & 1) * 16;

    int ct = next_chunk(0), it = 0;
    if (ct >= n_ct) return;                                                          // a row tile without entries (block-uniform)
    load_rows(q_s, qb, rt * DT);
    load_rows(k_s, kb, ct * DT);
    cp_async_wait_all();
    __syncthreads();
    uint32_t qa[HOLD_A ? KS : 1][4];
    if (HOLD_A) {
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) ldmatrix_x4(a_addr + ks * 32, qa[HOLD_A ? ks : 0]);
    }
    for (; ct < n_ct; ++it) {
        const int buf = it & 1;
        const int ct_next = next_chunk(ct + 1);
        if (ct_next < n_ct) load_rows(k_s + (buf ^ 1) * DT * XS, kb, ct_next * DT);  // its last readers finished before the previous barrier
        const int p0 = tp[(size_t)ct * n_rc], p1 = tp[(size_t)ct * n_rc + 1];
        uint32_t w[PRE];                                                             // this bucket's first entries, under the MMAs
#pragma unroll
        for (int j = 0; j < PRE; ++j) w[j] = p0 + j * THREADS + tid < p1 ? ent[p0 + j * THREADS + tid] : 0xffffffffu;
        // S[buf] = Q K_chunk^T
        float acc[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[j][i] = 0.0f;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            uint32_t a[4];
            if (HOLD_A) {
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = qa[HOLD_A ? ks : 0][i];
            } else {
                ldmatrix_x4(a_addr + ks * 32, a);
            }
#pragma unroll
            for (int jp = 0; jp < 2; ++jp) {
                uint32_t bb[4];
                ldmatrix_x4(k_s + buf * DT * XS + b_off + jp * 16 * XS + ks * 32, bb);
                mma16816(acc[2 * jp], a, bb[0], bb[1]);
                mma16816(acc[2 * jp + 1], a, bb[2], bb[3]);
            }
        }
        float *Sb = S0 + buf * DT * PS;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float *dst = Sb + (mt * 16 + g) * PS + nh * 32 + j * 8 + 2 * t;
            *reinterpret_cast<float2 *>(dst) = make_float2(acc[j][0], acc[j][1]);
            *reinterpret_cast<float2 *>(dst + 8 * PS) = make_float2(acc[j][2], acc[j][3]);
        }
        cp_async_wait_all();
        __syncthreads();
        // every entry of the bucket takes its cell
        auto emit = [&](uint32_t word) {
            if (word != 0xffffffffu) {
                float v = Sb[((word >> 6) & 63u) * PS + (word & 63u)] * scale;
                if (clamp > 0.0f) v = fminf(fmaxf(v, -clamp), clamp);
                vp[word >> 12] = v;
            }
        };
#pragma unroll
        for (int j = 0; j < PRE; ++j) emit(w[j]);
        for (int qi = p0 + PRE * THREADS + tid; qi < p1; qi += THREADS) emit(ent[qi]);
        ct = ct_next;
    }
}

template <int D>
static int launch_sddmm_d(const int32_t *tile_ptr, const uint32_t *tile_ent, const bf16 *q, const bf16 *k, float *values, int B,
                          int S, int64_t nnz, float scale, float clamp, cudaStream_t st) {
    const int n_t = (S + DT - 1) / DT;
    const size_t smem = 3 * (size_t)DT * (D + 8) * 2 + 2 * (size_t)DT * PS * 4;
    cudaFuncSetAttribute(sddmm_tiles_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    sddmm_tiles_kernel<D><<<(unsigned)((int64_t)B * n_t), THREADS, smem, st>>>(tile_ptr, tile_ent, q, k, values, B, S, nnz, n_t, n_t,
                                                                             scale, clamp);
    return after_launch("sddmm_tiles_kernel");
}

}  // namespace csr_tiles
}  // namespace spt

using namespace spt;

// tile_ptr: [B, n_t * n_t + 1] int32 with n_t = ceil(S / 64); tile_ent: [B, nnz] uint32.  No workspace.
extern "C" int spt_csr_tiles_supported(int S, int64_t nnz) { return S >= 1 && S <= csr_tiles::MAX_CT * csr_tiles::DT && nnz <= ((int64_t)1 << 20); }

extern "C" int64_t spt_csr_tiles_ptr_len(int S) {
    const int64_t n = (S + csr_tiles::DT - 1) / csr_tiles::DT;
    return n * n + 1;
}

extern "C" int spt_csr_tiles(const int32_t *indptr, const int32_t *indices, int32_t *tile_ptr, uint32_t *tile_ent, int B, int S,
                             int64_t nnz, spt_stream_t stream) {
    SPT_REQUIRE(indptr && indices && tile_ptr && tile_ent, "csr_tiles: null pointer");
    SPT_REQUIRE(B >= 1 && B <= 65535 && S >= 1 && nnz >= 0, "csr_tiles: bad sizes B=%d S=%d nnz=%lld", B, S, (long long)nnz);
    if (!spt_csr_tiles_supported(S, nnz))
        return fail(SPT_ERR_UNSUPPORTED, "csr_tiles: S=%d / nnz=%lld beyond the 32-bit entry format (S <= 8192, nnz <= 2^20 per head)",
                    S, (long long)nnz);
    const int n_t = (S + csr_tiles::DT - 1) / csr_tiles::DT;
    cudaStream_t st = as_stream(stream);
    csr_tiles::tiles_count_kernel<<<dim3(n_t, B), csr_tiles::THREADS, 0, st>>>(indptr, indices, tile_ptr, S, nnz, n_t, n_t);
    SPT_LAUNCH_CHECK("tiles_count_kernel");
    csr_tiles::tiles_scan_kernel<<<B, 1024, 0, st>>>(tile_ptr, n_t * n_t);
    SPT_LAUNCH_CHECK("tiles_scan_kernel");
    csr_tiles::tiles_place_kernel<<<dim3(n_t, B), csr_tiles::THREADS, 0, st>>>(indptr, indices, tile_ptr, tile_ent, S, nnz, n_t, n_t);
    return after_launch("tiles_place_kernel");
}

extern "C" int spt_spmm_tiles_fwd(const int32_t *tile_ptr, const uint32_t *tile_ent, const float *values, const void *x, void *y,
                                  int B, int S, int d, int64_t nnz, int dtype, int out_dtype, int trans, spt_stream_t stream) {
    SPT_REQUIRE(tile_ptr && tile_ent && values && x && y, "spmm_tiles_fwd: null pointer");
    SPT_REQUIRE(B >= 1 && S >= 1 && nnz >= 0, "spmm_tiles_fwd: bad sizes B=%d S=%d nnz=%lld", B, S, (long long)nnz);
    if (dtype != SPT_BF16 || (d != 64 && d != 128) || !spt_csr_tiles_supported(S, nnz) || (out_dtype != SPT_BF16 && out_dtype != SPT_F32))
        return fail(SPT_ERR_UNSUPPORTED, "spmm_tiles_fwd: bf16 x with head dim 64 / 128 only (dtype %d, d %d)", dtype, d);
    SPT_REQUIRE((uintptr_t)x % 16 == 0 && (uintptr_t)y % 8 == 0, "spmm_tiles_fwd: x must be 16-byte, y 8-byte aligned");
    using bf = __nv_bfloat16;
    cudaStream_t st = as_stream(stream);
    const bool t = trans != 0;
    if (d == 64)
        return out_dtype == SPT_BF16 ? csr_tiles::launch_d<64, bf>(t, tile_ptr, tile_ent, values, (const bf *)x, (bf *)y, B, S, nnz, st)
                                     : csr_tiles::launch_d<64, float>(t, tile_ptr, tile_ent, values, (const bf *)x, (float *)y, B, S, nnz, st);
    return out_dtype == SPT_BF16 ? csr_tiles::launch_d<128, bf>(t, tile_ptr, tile_ent, values, (const bf *)x, (bf *)y, B, S, nnz, st)
                                 : csr_tiles::launch_d<128, float>(t, tile_ptr, tile_ent, values, (const bf *)x, (float *)y, B, S, nnz, st);
}

extern "C" int spt_spmm_t_tiles_fwd(const int32_t *tile_ptr, const uint32_t *tile_ent, const float *values, const void *x, void *y,
                                    int B, int S, int d, int64_t nnz, int dtype, int out_dtype, spt_stream_t stream) {
    return spt_spmm_tiles_fwd(tile_ptr, tile_ent, values, x, y, B, S, d, nnz, dtype, out_dtype, 1, stream);
}

/* values[b, e] = clamp(scale * <query[b, row(e)], key[b, col(e)]>, -clamp, clamp) (clamp <= 0: none) for every entry of
 * the tile index — the spt_sddmm_fwd product for bf16 operands with head dim 64 / 128.  Entries whose column is outside
 * [0, S) are not in the index: their values are left untouched. */
extern "C" int spt_sddmm_tiles_fwd(const int32_t *tile_ptr, const uint32_t *tile_ent, const void *query, const void *key,
                                   float *values, int B, int S, int d, int64_t nnz, float scale, float clamp, int dtype,
                                   spt_stream_t stream) {
    SPT_REQUIRE(tile_ptr && tile_ent && query && key && values, "sddmm_tiles_fwd: null pointer");
    SPT_REQUIRE(B >= 1 && S >= 1 && nnz >= 0, "sddmm_tiles_fwd: bad sizes B=%d S=%d nnz=%lld", B, S, (long long)nnz);
    if (dtype != SPT_BF16 || (d != 64 && d != 128) || !spt_csr_tiles_supported(S, nnz))
        return fail(SPT_ERR_UNSUPPORTED, "sddmm_tiles_fwd: bf16 operands with head dim 64 / 128 only (dtype %d, d %d)", dtype, d);
    SPT_REQUIRE((uintptr_t)query % 16 == 0 && (uintptr_t)key % 16 == 0, "sddmm_tiles_fwd: operands must be 16-byte aligned");
    using bf = __nv_bfloat16;
    cudaStream_t st = as_stream(stream);
    return d == 64 ? csr_tiles::launch_sddmm_d<64>(tile_ptr, tile_ent, (const bf *)query, (const bf *)key, values, B, S, nnz, scale, clamp, st)
                   : csr_tiles::launch_sddmm_d<128>(tile_ptr, tile_ent, (const bf *)query, (const bf *)key, values, B, S, nnz, scale, clamp, st);
}
