// Transposed sparse product  y[c] = sum_r A[r, c] x[r]  (bf16 x, head dim 64 / 128) on DENSE 64 x 64 tiles built from the
// CSC pattern — reference call sites extension/spmm.cpp:27-69 with trans_lhs (the dV = P^T dO and dK = dS^T Q products of
// the backward passes, kernels/spmm.py, kernels/sddmm.py).
//
// Why not gather: at the densities this path runs at (top-k 256 of at most 2048 causal keys: 1/8 of the square, 1/4 and
// more of the causal triangle) the gathered kernel (csr.cu: spmm2_t_kernel) moves one 128-byte x row, one 32-byte value
// sector and 8 index bytes PER ENTRY through L2 -> SM: 67 M entries x 168 B = 11 GB per call at the bench shape, i.e. it
// runs at the L2 bandwidth (1.3 ms).  Here a block owns 64 columns of one head and walks the 64-row chunks its columns
// touch: the entries of every (column, chunk) are one contiguous run of the column's CSC list (csr2csc emits rows in
// ascending order), scattered with shared-memory atomics into a 64 x 64 fp32 tile — duplicates (the zero-padding
// entries all point at column 0) simply add up — and the tile times the x chunk (64 rows staged ONCE per 64 columns) is
// 128 mma.sync (m16n8k16) per warp.  Traffic per entry: 8 index bytes + the value sector; x: 8 KB per tile.
//   * the weights stay fp32-accurate: the tile is split into bf16 hi + lo halves in registers (two MMAs), x is bf16
//     already, accumulation is fp32;
//   * lists are walked by 4 lanes per column that hold an aligned window of 16 entries (rows / positions) in registers:
//     a chunk's run is consumed in one or two steps of independent gathers;
//   * columns with very long lists (column 0 collects every short row's padding: 32 K entries at S 2048) would serialise
//     on their 4 lanes: up to DT_MAXH of them per block are first accumulated by the whole block into per-row strips.
// Summation order: chunks ascending, tensor-core order inside a chunk, atomics inside a (column, row) cell — the result
// is deterministic except for the fp32 order in which duplicate entries of one cell add up.
#include <climits>
#include <cstdlib>

#include "common.cuh"

namespace spt {
namespace csr_dense {

using bf16 = __nv_bfloat16;

constexpr int DT = 64;                 // tile edge (columns per block, rows per chunk)
constexpr int DT_THREADS = 256;
constexpr int PS = 72;                 // fp32 tile row stride: (8 g + 2 t) -> the 16 lanes of an LDS.64 phase hit distinct banks
constexpr int DT_HEAVY = 2048;         // entries: longer column lists go through a strip
constexpr int DT_MAXH = 2;

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
// (v0, v1) fp32 -> packed bf16 high parts and packed bf16 remainders
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

template <int D, typename TO>
__global__ void __launch_bounds__(DT_THREADS, 4)
spmm_t_dense_kernel(const int32_t *__restrict__ col_ptr, const int32_t *__restrict__ row_idx, const int32_t *__restrict__ perm,
                    const float *__restrict__ values, const bf16 *__restrict__ x, TO *__restrict__ y, int B, int S, int64_t nnz,
                    int paired) {
    constexpr int XS = (D + 8) * 2;            // bytes per staged x row (16-byte pad: conflict-free ldmatrix)
    constexpr int NT = D / 16;                 // n-tiles (8 features) per warp: a warp owns 16 columns x D / 2 features
    extern __shared__ __align__(16) unsigned char smem[];
    float *P = reinterpret_cast<float *>(smem);                                   // [64 columns][PS]
    unsigned char *X = smem + DT * PS * 4;                                         // [64 rows][XS]
    float *strip = reinterpret_cast<float *>(X + DT * XS);                         // [DT_MAXH][S]
    __shared__ int s_cp[DT + 1];
    __shared__ int s_heavy[DT_MAXH];
    __shared__ int s_nheavy, s_rmin, s_rmax, s_bad;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // paired (opt-in): a block takes column tiles p and n_tiles - 1 - p of one head, one after the other
    const int n_tiles = (S + DT - 1) / DT;
    const int b = blockIdx.x % B, pair = blockIdx.x / B;
    const int32_t *pp = col_ptr + (size_t)b * (S + 1);
    const int32_t *ip = row_idx + (size_t)b * nnz;
    const int32_t *pm = perm + (size_t)b * nnz;
    const float *vp = values + (size_t)b * nnz;
    const bf16 *xb = x + (size_t)b * S * D;

  for (int pass = 0; pass < (paired ? 2 : 1); ++pass) {
    const int tile = pass ? n_tiles - 1 - pair : pair;
    if (pass && tile <= pair) break;                               // odd tile count: the middle tile is done once
    const int c0 = tile * DT;
    __syncthreads();                                               // the previous pass is done with the shared state
    if (tid <= DT) s_cp[tid] = pp[min(c0 + tid, S)];
    if (tid == 0) {
        s_nheavy = 0;
        s_rmin = INT_MAX;
        s_rmax = -1;
        s_bad = 0;
    }
    __syncthreads();
    if (tid < DT) {
        const int e0 = s_cp[tid], e1 = s_cp[tid + 1];
        if (e1 > e0) {
            atomicMin(&s_rmin, ip[e0]);
            atomicMax(&s_rmax, ip[e1 - 1]);
            if (e1 - e0 > DT_HEAVY) {
                const int slot = atomicAdd(&s_nheavy, 1);
                if (slot < DT_MAXH) s_heavy[slot] = tid;
            }
        }
    }
    __syncthreads();
    const int n_heavy = min(s_nheavy, DT_MAXH);
    // long lists: the whole block adds the column's entries into a per-row strip
    for (int h = 0; h < n_heavy; ++h) {
        float *sp = strip + (size_t)h * S;
        for (int r = tid; r < S; r += DT_THREADS) sp[r] = 0.0f;
        __syncthreads();
        const int e0 = s_cp[s_heavy[h]], e1 = s_cp[s_heavy[h] + 1];
        int lo = INT_MAX, hi = -1;
        // Four independent (row, position -> value) chains per thread and trip.  The list is sorted by row and column 0's
        // holds each early row hundreds of times (zero padding): fp32 shared-memory atomics are CAS loops
        // (ATOMS.CAST.SPIN), so 32 lanes adding to one address serialise 32-fold — the first version of this loop was
        // the critical path of the whole kernel (0.6 ms for one block).  Each warp first sums the runs of adjacent equal
        // rows among its 32 entries with shuffles; only the last lane of a run touches the strip.
        const int e_round = e0 + ((e1 - e0 + 4 * DT_THREADS - 1) / (4 * DT_THREADS)) * (4 * DT_THREADS);
        for (int e = e0 + tid; e < e_round; e += 4 * DT_THREADS) {
            int r[4], q[4];
            float v[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const bool ok = e + i * DT_THREADS < e1;
                r[i] = ok ? ip[e + i * DT_THREADS] : -1;
                q[i] = ok ? pm[e + i * DT_THREADS] : 0;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = r[i] >= 0 ? vp[q[i]] : 0.0f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int prev = __shfl_up_sync(0xffffffffu, r[i], 1), next = __shfl_down_sync(0xffffffffu, r[i], 1);
                int start = (lane == 0 || prev != r[i]) ? lane : 0;           // first lane of this lane's run
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) start = max(start, __shfl_up_sync(0xffffffffu, start, d) * (lane >= d));
                float sum = v[i];
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const float up = __shfl_up_sync(0xffffffffu, sum, d);
                    if (lane - d >= start) sum += up;
                }
                if ((lane == 31 || next != r[i]) && (unsigned)r[i] < (unsigned)S) {
                    atomicAdd(sp + r[i], sum);
                    lo = min(lo, r[i]);
                    hi = max(hi, r[i]);
                }
            }
        }
        if (hi >= 0) {                                   // whatever the order of this list, its rows are inside the chunk range
            atomicMin(&s_rmin, lo);
            atomicMax(&s_rmax, hi);
        }
    }
    __syncthreads();
    const int r_begin = (max(s_rmin, 0) / DT) * DT, r_last = min(s_rmax, S - 1);

    // list walkers: 4 lanes per column.  The quad holds an ALIGNED window of 16 entries of its column's list in registers
    // (lane `sub`: entries wbase + 4 sub .. + 3, one 16-byte load each for rows and positions — ncu: with scalar loads of
    // an unaligned window the index fetches alone were a quarter of the kernel's LSU wavefronts); `cur` is the first
    // entry not yet consumed.  A chunk consumes the leading run of pending entries whose rows lie below its end — rows
    // ascend, so that run is everything the chunk owns — and a window is refetched only when it is used up.
    const int cl = tid >> 2, sub = tid & 3;
    const int quad_shift = lane & ~3;
    int my_strip = -1;
    for (int h = 0; h < n_heavy; ++h)
        if (s_heavy[h] == cl) my_strip = h;
    int cur = s_cp[cl];
    const int e_end = my_strip >= 0 ? cur : s_cp[cl + 1];          // strip columns have nothing to walk
    int wbase = cur & ~15;
    // Two windows in flight (the first version fetched a window and gathered its values only when a chunk asked for them:
    // two dependent DRAM latencies per pass, which made a block's duration the critical path — 0.42 ms for 32 heads):
    // the CURRENT window holds rows + the gathered VALUES, the NEXT one rows + positions.  Advancing turns the next window
    // into the current one (its value gathers go out, their addresses are there already) and fetches a new next window;
    // all of it completes under the MMAs of the chunks in between.
    int rw[4], rwn[4], psn[4];
    float wv[4];
    auto load_idx = [&](int base, int (&r)[4], int (&p)[4]) {
        const int q0 = base + 4 * sub;
        if (q0 < e_end && (int64_t)q0 + 4 <= nnz) {
            const int4 r4 = __ldg(reinterpret_cast<const int4 *>(ip + q0)), p4 = __ldg(reinterpret_cast<const int4 *>(pm + q0));
            r[0] = r4.x; r[1] = r4.y; r[2] = r4.z; r[3] = r4.w;
            p[0] = p4.x; p[1] = p4.y; p[2] = p4.z; p[3] = p4.w;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (q0 + i >= e_end) {                                 // beyond the list: never consumed
                r[i] = INT_MAX;
                p[i] = 0;
            }
    };
    auto gather = [&](const int (&r)[4], const int (&p)[4]) {
#pragma unroll
        for (int i = 0; i < 4; ++i) wv[i] = r[i] != INT_MAX ? vp[p[i]] : 0.0f;
    };
    load_idx(wbase, rw, psn);
    gather(rw, psn);
    load_idx(wbase + 16, rwn, psn);

    // MMA roles: warp -> 16 columns (mt) x half of the features (nh)
    const int g = lane >> 2, t = lane & 3;
    const int mt = warp & 3, nh = warp >> 2;
    float acc[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[j][i] = 0.0f;
    const uint32_t x_s = (uint32_t)__cvta_generic_to_shared(X);
    const uint32_t b_addr = x_s + ((lane & 7) + ((lane >> 3) & 1) * 8) * XS + (nh * (D / 2) + (lane >> 4) * 8) * 2;
    float *prow = P + cl * PS + sub * 16;                          // this lane's 16 cells of its column's tile row

    for (int r0 = r_begin; r0 <= r_last; r0 += DT) {
        // (1) x chunk -> shared memory (asynchronous), rows beyond S read as zero
        for (int i = tid; i < DT * (D / 8); i += DT_THREADS) {
            const int rr = i / (D / 8), ch = i % (D / 8);
            const bool ok = r0 + rr < S;
            cp_async16(x_s + rr * XS + ch * 16, xb + (size_t)(ok ? r0 + rr : 0) * D + ch * 8, ok);
        }
        // (2) the warp's 8 tile rows (its 8 columns: 576 contiguous floats) are cleared by the warp itself — consecutive
        // 16-byte stores, no block barrier needed before its lanes scatter into them; strip columns are then overwritten
        {
            float4 *pw = reinterpret_cast<float4 *>(P + warp * 8 * PS);
#pragma unroll
            for (int i = lane; i < 8 * PS / 4; i += 32) pw[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        }
        __syncwarp();
        if (my_strip >= 0) {
            const float *sp = strip + (size_t)my_strip * S + r0 + sub * 16;
#pragma unroll
            for (int i = 0; i < 16; ++i) prow[i] = (r0 + sub * 16 + i < S) ? sp[i] : 0.0f;
        }
        for (;;) {
            const int lim = r0 + DT;
            const int q0 = wbase + 4 * sub;
            // leading entries of this lane that are consumed already or belong to this chunk
            int mine = 0;
#pragma unroll
            for (int i = 3; i >= 0; --i) mine = (q0 + i < cur || rw[i] < lim) ? mine + 1 : 0;
            const uint32_t full = (__ballot_sync(0xffffffffu, mine == 4) >> quad_shift) & 15u;
            const int lead = __ffs(~full & 15u | 16u) - 1;            // lanes 0 .. lead - 1 of the quad are all-in
            const int take = sub <= lead ? mine : 0;
            float v[4] = {wv[0], wv[1], wv[2], wv[3]};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (i < take && q0 + i >= cur) {
                    const int cell = rw[i] - r0;
                    // adjacent entries of one cell (duplicates) are summed in registers: one atomic per run and lane
                    if (i < 3 && i + 1 < take && rw[i + 1] == rw[i]) v[i + 1] += v[i];
                    else if (cell >= 0) atomicAdd(P + cl * PS + cell, v[i]);
                    else s_bad = 1;                                  // a row below the chunk: the list is not ascending
                }
            }
            // new cursor = end of the leading run (uniform in the quad)
            const int lead_mine = __shfl_sync(0xffffffffu, mine, quad_shift + min(lead, 3));
            const int ncur = max(cur, lead == 4 ? wbase + 16 : wbase + 4 * lead + lead_mine);
            const bool next_window = ncur >= wbase + 16 && ncur < e_end;
            cur = ncur;
            if (next_window) {
                wbase += 16;
#pragma unroll
                for (int i = 0; i < 4; ++i) rw[i] = rwn[i];
                gather(rw, psn);
                load_idx(wbase + 16, rwn, psn);
            }
            if (__ballot_sync(0xffffffffu, next_window) == 0) break;  // warp-uniform: every quad has reached its chunk's end
        }
        cp_async_wait_all();
        __syncthreads();
        // (3) acc += tile (bf16 hi + lo) x chunk
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
    }
    if (cur < e_end) s_bad = 1;                                      // entries left behind the last chunk: same reason
    __syncthreads();
    if (s_bad) {
        // The lists of this block are not in ascending row order (not an spt_csr2csc output): plain gathered product,
        // a warp per column, lane = D / 32 features.
        constexpr int F = D / 32;
        for (int u = 0; u < DT / 8; ++u) {
            const int c = c0 + warp * (DT / 8) + u;
            if (c >= S) break;
            float a[F];
#pragma unroll
            for (int i = 0; i < F; ++i) a[i] = 0.0f;
            for (int q = s_cp[c - c0]; q < s_cp[c - c0 + 1]; ++q) {
                const int r = ip[q];
                if ((unsigned)r >= (unsigned)S) continue;
                const float v = vp[pm[q]];
#pragma unroll
                for (int i = 0; i < F; ++i) a[i] = fmaf(v, __bfloat162float(xb[(size_t)r * D + lane * F + i]), a[i]);
            }
#pragma unroll
            for (int i = 0; i < F; i += 2) store2<TO>(y + ((size_t)b * S + c) * D + lane * F + i, a[i], a[i + 1]);
        }
        continue;
    }
    // epilogue: rows g / g + 8 of the warp's 16 columns, features nh D/2 + 8 j + 2 t
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int c = c0 + mt * 16 + g + half * 8;
        if (c < S) {
            TO *dst = y + ((size_t)b * S + c) * D + nh * (D / 2) + 2 * t;
#pragma unroll
            for (int j = 0; j < NT; ++j) store2<TO>(dst + 8 * j, acc[j][2 * half], acc[j][2 * half + 1]);
        }
    }
  }
}

static size_t smem_bytes(int D, int S) { return (size_t)DT * PS * 4 + (size_t)DT * (D + 8) * 2 + (size_t)DT_MAXH * S * 4; }

bool supported(int d, int S, const void *x, const void *y) {
    static const bool off = [] { const char *e = getenv("SPT_SPMM_T_DENSE"); return e && atoi(e) == 0; }();   // A/B switch
    return !off && (d == 64 || d == 128) && ((uintptr_t)x % 16 == 0) && ((uintptr_t)y % 8 == 0) && smem_bytes(d, S) <= 100 * 1024;
}
// the 16-byte window loads need every head's lists 16-byte aligned
static bool lists_aligned(const void *row_idx, const void *perm, int64_t nnz) {
    return nnz % 4 == 0 && ((uintptr_t)row_idx % 16 == 0) && ((uintptr_t)perm % 16 == 0);
}

template <int D, typename TO>
static int launch_d(const int32_t *col_ptr, const int32_t *row_idx, const int32_t *perm, const float *values, const bf16 *x,
                    TO *y, int B, int S, int64_t nnz, cudaStream_t st) {
    const size_t smem = smem_bytes(D, S);
    cudaFuncSetAttribute(spmm_t_dense_kernel<D, TO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    // SPT_SPMM_T_PAIR=1: a block takes column tiles p and n - 1 - p (equal chunk counts under a causal pattern).  Measured
    // at S 2048: no gain — 128 heads 0.82 ms paired against 0.76 ms, 32 heads 0.43 against 0.42 (a block's duration is
    // set by the latency of its window passes, not by the balance) — so one tile per block is the default.
    static const int paired = [] { const char *e = getenv("SPT_SPMM_T_PAIR"); return e && atoi(e) == 1 ? 1 : 0; }();
    const int n_tiles = (S + DT - 1) / DT;
    const int64_t blocks = (int64_t)B * (paired ? (n_tiles + 1) / 2 : n_tiles);
    spmm_t_dense_kernel<D, TO><<<(unsigned)blocks, DT_THREADS, smem, st>>>(col_ptr, row_idx, perm, values, x, y, B, S, nnz, paired);
    return after_launch("spmm_t_dense_kernel");
}

int launch_spmm_t(const int32_t *col_ptr, const int32_t *row_idx, const int32_t *perm, const float *values, const bf16 *x,
                  void *y, bool y_bf16, int B, int S, int d, int64_t nnz, cudaStream_t st) {
    if (!lists_aligned(row_idx, perm, nnz)) return SPT_ERR_UNSUPPORTED;      // caller falls back to the gathered kernel
    if (d == 64)
        return y_bf16 ? launch_d<64, bf16>(col_ptr, row_idx, perm, values, x, (bf16 *)y, B, S, nnz, st)
                      : launch_d<64, float>(col_ptr, row_idx, perm, values, x, (float *)y, B, S, nnz, st);
    return y_bf16 ? launch_d<128, bf16>(col_ptr, row_idx, perm, values, x, (bf16 *)y, B, S, nnz, st)
                  : launch_d<128, float>(col_ptr, row_idx, perm, values, x, (float *)y, B, S, nnz, st);
}

}  // namespace csr_dense
}  // namespace spt
