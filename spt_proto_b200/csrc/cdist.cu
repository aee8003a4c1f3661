// (1) cdist: L1 distance from PQ sub-vectors to codewords + argmin -> PQ codes.
//
// Replaces cdist_forward_cuda / cdist_backward_cuda (reference extension/cdist.cu:185-333) and
// fuses PQBase.forward(mode='encode') (naive_gpt/layers/basic/quantizer.py:26-77).
//
// Numerics contract (bit-exact codes): distance = sum_i |q_i - t_i| accumulated over i ascending
// in fp32 (no FMA is possible: there is no multiply), running strict-'<' minimum starting at 1e13
// => lowest codeword index wins ties (cdist.cu:29,46-54).  bf16 inputs are widened to fp32 first.
//
// This is L1, not a GEMM: it stays on the CUDA cores (hard part 1).  Per
// (sub-vector, codeword, element) it costs one FADD + one FADD-with-|.| source modifier.
// Bound: issue-bound SIMT (6*n*d*c flop vs 2*n*d*e bytes) — see DESIGN.md section 5.
#include "common.cuh"

namespace spt {

constexpr int CDIST_THREADS = 256;

// ---- forward, reference layout: query [m, n, dc] ---------------------------------------------
// One thread per (subspace, query).  The sub-space's codebook slice lives in shared memory and is
// read as a warp-wide broadcast.  DC > 0: compile-time sub-vector length (query in registers);
// DC == 0: generic path, query element re-read through L1.
template <typename T, int DC>
__global__ void __launch_bounds__(CDIST_THREADS)
cdist_fwd_kernel(const T *__restrict__ query, const float *__restrict__ table, float *__restrict__ distance,
                 int32_t *__restrict__ indices, int64_t n, int c, int dc_rt) {
    extern __shared__ float s_table[];  // [c][dc]
    const int dc = DC > 0 ? DC : dc_rt;
    const int s = blockIdx.y;
    for (int i = threadIdx.x; i < c * dc; i += blockDim.x) s_table[i] = table[(size_t)s * c * dc + i];
    __syncthreads();
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const T *qp = query + ((size_t)s * n + q) * dc;
    float qv[DC > 0 ? DC : 1];
    if (DC > 0) {
#pragma unroll
        for (int i = 0; i < DC; ++i) qv[i] = to_f32(qp[i]);
    }
    float *dp = distance ? distance + ((size_t)s * n + q) * c : nullptr;
    const bool vec_store = dp && (c % 4 == 0);
    int min_index = 0;
    float min_distance = 1e13f;
    float4 pack;
    for (int w = 0; w < c; ++w) {
        const float *tp = s_table + w * dc;
        float reduced = 0.0f;
        if (DC > 0) {
#pragma unroll
            for (int i = 0; i < DC; ++i) reduced += fabsf(qv[i] - tp[i]);
        } else {
            for (int i = 0; i < dc; ++i) reduced += fabsf(to_f32(qp[i]) - tp[i]);
        }
        if (reduced < min_distance) {
            min_distance = reduced;
            min_index = w;
        }
        if (vec_store) {
            (&pack.x)[w & 3] = reduced;
            if ((w & 3) == 3) st_stream(reinterpret_cast<float4 *>(dp + w - 3), pack);
        } else if (dp) {
            dp[w] = reduced;
        }
    }
    indices[(size_t)s * n + q] = min_index;
}

// ---- fused encode, layer layout: z [rows, m*dc] -> codes [rows, m] ----------------------------
// Thread g handles (row = g / m, subspace = g % m): consecutive threads read consecutive dc-element
// chunks (fully coalesced 16/32-byte loads) and write consecutive int32 codes.  L1 distance is not a
// GEMM; the kernel is bound by the fp32 add pipe (2 adds per (element, codeword)), so the adds are
// issued as packed FADD2 (sm_100 add.f32x2) over PAIRS OF CODEWORDS: the two halves carry codewords
// w and w+1, the element loop stays sequential in i, hence every distance is still the reference's
// i-ascending fp32 sum and, with the strict `<`, the codes are bit-identical to the reference kernel
// (extension/cdist.cu:42-54).  |x| is a LOP3 on the otherwise idle ALU pipe.  The codebook is staged in
// shared memory as [codeword pair][element][subspace] float2 (lanes of a warp differ in subspace:
// distinct banks, equal-subspace lanes broadcast).  blockIdx.y selects one of up to two tensors (q and
// k of a layer share the codebook and are encoded by one launch).
__device__ __forceinline__ uint64_t sub_f32x2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("sub.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t add_f32x2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

template <typename T, int DC>
__global__ void __launch_bounds__(CDIST_THREADS)
pq_encode_kernel(const T *__restrict__ z0, const T *__restrict__ z1, const float *__restrict__ table,
                 int32_t *__restrict__ codes0, int32_t *__restrict__ codes1, int64_t total /* rows*m */, int m, int c) {
    extern __shared__ __align__(16) float s_table[];  // [(c+1)/2][DC][m][2]; an odd last codeword is paired with +inf
    const int cp = (c + 1) / 2;
    for (int i = threadIdx.x; i < cp * DC * m * 2; i += blockDim.x) {
        const int h = i & 1, s = (i >> 1) % m, e = ((i >> 1) / m) % DC, wp = (i >> 1) / (m * DC);
        const int w = 2 * wp + h;
        s_table[i] = w < c ? table[((size_t)s * c + w) * DC + e] : __int_as_float(0x7f800000);
    }
    __syncthreads();
    const T *z = blockIdx.y ? z1 : z0;
    int32_t *codes = blockIdx.y ? codes1 : codes0;
    // grid-stride over (row, subspace) items: the codebook is staged once per resident block
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (int64_t)gridDim.x * blockDim.x) {
        const int s = (int)(g % m);
        float qv[DC];
        const T *zp = z + (size_t)g * DC;
        if constexpr (DC % Vec16<T>::N == 0) {
#pragma unroll
            for (int i = 0; i < DC; i += Vec16<T>::N) {
                float tmp[Vec16<T>::N];
                Vec16<T>::load(zp + i, tmp);
#pragma unroll
                for (int j = 0; j < Vec16<T>::N; ++j) qv[i + j] = tmp[j];
            }
        } else {
#pragma unroll
            for (int i = 0; i < DC; ++i) qv[i] = to_f32(zp[i]);
        }
        uint64_t q2[DC];   // (q_i, q_i)
#pragma unroll
        for (int i = 0; i < DC; ++i) q2[i] = ((uint64_t)__float_as_uint(qv[i]) << 32) | __float_as_uint(qv[i]);
        int min_index = 0;
        float min_distance = 1e13f;
        const uint64_t *tp = reinterpret_cast<const uint64_t *>(s_table) + s;
#pragma unroll 2
        for (int wp = 0; wp < cp; ++wp, tp += (size_t)DC * m) {
            uint64_t acc = 0;   // (+0.0f, +0.0f)
#pragma unroll
            for (int i = 0; i < DC; ++i)
                acc = add_f32x2(acc, sub_f32x2(q2[i], tp[(size_t)i * m]) & 0x7fffffff7fffffffull);
            const float d0 = __uint_as_float((uint32_t)acc), d1 = __uint_as_float((uint32_t)(acc >> 32));
            if (d0 < min_distance) {
                min_distance = d0;
                min_index = 2 * wp;
            }
            if (d1 < min_distance) {   // an odd c pairs its last codeword with +inf: never selected
                min_distance = d1;
                min_index = 2 * wp + 1;
            }
        }
        codes[g] = min_index;
    }
}

// ---- backward wrt query: gq[s,n,i] = sum_c sgn(q_i - t_ci) * g[s,n,c]  (cdist.cu:72-131) --------
template <int DC>
__global__ void __launch_bounds__(CDIST_THREADS)
cdist_bwd_query_kernel(const float *__restrict__ query, const float *__restrict__ table,
                       const float *__restrict__ grad, float *__restrict__ grad_query, int64_t n, int c,
                       int dc_rt) {
    extern __shared__ float s_table[];
    const int dc = DC > 0 ? DC : dc_rt;
    const int s = blockIdx.y;
    for (int i = threadIdx.x; i < c * dc; i += blockDim.x) s_table[i] = table[(size_t)s * c * dc + i];
    __syncthreads();
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const float *qp = query + ((size_t)s * n + q) * dc;
    const float *gp = grad + ((size_t)s * n + q) * c;
    float *op = grad_query + ((size_t)s * n + q) * dc;
    if (DC > 0) {
        float qv[DC > 0 ? DC : 1], acc[DC > 0 ? DC : 1];
#pragma unroll
        for (int i = 0; i < DC; ++i) { qv[i] = qp[i]; acc[i] = 0.0f; }
        for (int w = 0; w < c; ++w) {
            const float g = gp[w];
            const float *tp = s_table + w * DC;
#pragma unroll
            for (int i = 0; i < DC; ++i) acc[i] += (qv[i] - tp[i]) > 0.0f ? g : -g;
        }
#pragma unroll
        for (int i = 0; i < DC; ++i) op[i] = acc[i];
    } else {
        for (int i = 0; i < dc; ++i) {
            const float qi = qp[i];
            float acc = 0.0f;
            for (int w = 0; w < c; ++w) acc += (qi - s_table[w * dc + i]) > 0.0f ? gp[w] : -gp[w];
            op[i] = acc;
        }
    }
}

// ---- backward wrt table: gt[s,c,i] = -sum_n sgn(q_ni - t_ci) * g[s,n,c]  (cdist.cu:134-182) -----
// Stage 1: each block reduces a chunk of CHUNK queries for one subspace into a partial [c*dc]
// (the reference runs 16 blocks serially over all n, cdist.cu:320-329).  Stage 2 adds the partials
// in chunk order => deterministic.
constexpr int BWD_T_CHUNK = 512;
constexpr int BWD_T_TILE = 32;

__global__ void __launch_bounds__(CDIST_THREADS)
cdist_bwd_table_stage1(const float *__restrict__ query, const float *__restrict__ table,
                       const float *__restrict__ grad, float *__restrict__ partial, int64_t n, int c, int dc,
                       int n_chunks) {
    extern __shared__ float smem[];
    float *s_q = smem;                      // [TILE][dc]
    float *s_g = smem + BWD_T_TILE * dc;    // [TILE][c]
    const int s = blockIdx.y, chunk = blockIdx.x;
    const int64_t n0 = (int64_t)chunk * BWD_T_CHUNK;
    const int64_t n1 = min(n, n0 + BWD_T_CHUNK);
    const int n_out = c * dc;
    // each thread owns outputs o = tid, tid + blockDim, ... ; keep up to 4 in registers
    constexpr int MAX_OWN = 4;
    float acc[MAX_OWN], tv[MAX_OWN];
    int own = 0;
    for (int o = threadIdx.x; o < n_out && own < MAX_OWN; o += blockDim.x, ++own) {
        acc[own] = 0.0f;
        tv[own] = table[(size_t)s * n_out + o];
    }
    for (int64_t t0 = n0; t0 < n1; t0 += BWD_T_TILE) {
        const int rows = (int)min((int64_t)BWD_T_TILE, n1 - t0);
        for (int i = threadIdx.x; i < rows * dc; i += blockDim.x) s_q[i] = query[((size_t)s * n + t0) * dc + i];
        for (int i = threadIdx.x; i < rows * c; i += blockDim.x) s_g[i] = grad[((size_t)s * n + t0) * c + i];
        __syncthreads();
        int k = 0;
        for (int o = threadIdx.x; o < n_out && k < MAX_OWN; o += blockDim.x, ++k) {
            const int w = o / dc, i = o % dc;
            float a = acc[k];
            const float t = tv[k];
            for (int r = 0; r < rows; ++r) {
                const float g = s_g[r * c + w];
                a -= (s_q[r * dc + i] - t) > 0.0f ? g : -g;
            }
            acc[k] = a;
        }
        __syncthreads();
    }
    int k = 0;
    for (int o = threadIdx.x; o < n_out && k < MAX_OWN; o += blockDim.x, ++k)
        partial[((size_t)s * n_chunks + chunk) * n_out + o] = acc[k];
}

__global__ void cdist_bwd_table_stage2(const float *__restrict__ partial, float *__restrict__ grad_table,
                                       int n_out, int n_chunks) {
    const int s = blockIdx.y;
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= n_out) return;
    float a = 0.0f;
    for (int ch = 0; ch < n_chunks; ++ch) a += partial[((size_t)s * n_chunks + ch) * n_out + o];
    grad_table[(size_t)s * n_out + o] = a;
}

template <typename T>
static int launch_cdist_fwd(const T *query, const float *table, float *distance, int32_t *indices, int m,
                            int64_t n, int c, int dc, cudaStream_t st) {
    dim3 grid((unsigned)((n + CDIST_THREADS - 1) / CDIST_THREADS), m);
    size_t smem = (size_t)c * dc * sizeof(float);
#define SPT_CDIST_CASE(D)                                                                               \
    case D:                                                                                             \
        if (smem > 48 * 1024)                                                                           \
            cudaFuncSetAttribute(cdist_fwd_kernel<T, D>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                 (int)smem);                                                            \
        cdist_fwd_kernel<T, D><<<grid, CDIST_THREADS, smem, st>>>(query, table, distance, indices, n, c, dc); \
        break;
    switch (dc) {
        SPT_CDIST_CASE(4)
        SPT_CDIST_CASE(8)
        SPT_CDIST_CASE(16)
        SPT_CDIST_CASE(32)
        default:
            if (smem > 48 * 1024)
                cudaFuncSetAttribute(cdist_fwd_kernel<T, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem);
            cdist_fwd_kernel<T, 0><<<grid, CDIST_THREADS, smem, st>>>(query, table, distance, indices, n, c, dc);
    }
#undef SPT_CDIST_CASE
    return after_launch("cdist_fwd_kernel");
}

template <typename T>
static int launch_pq_encode(const T *z0, const T *z1, const float *table, int32_t *codes0, int32_t *codes1,
                            int64_t rows, int m, int c, int dc, cudaStream_t st) {
    const int64_t total = rows * m;
    const int64_t want = (total + CDIST_THREADS - 1) / CDIST_THREADS;
    const int64_t cap = (int64_t)num_sms() * 8 / (z1 ? 2 : 1);      // resident blocks: one codebook staging each
    dim3 grid((unsigned)(want < cap ? want : cap), z1 ? 2 : 1);
    size_t smem = (size_t)m * ((c + 1) / 2) * 2 * dc * sizeof(float);
#define SPT_ENC_CASE(D)                                                                                 \
    case D:                                                                                             \
        if (smem > 48 * 1024)                                                                           \
            cudaFuncSetAttribute(pq_encode_kernel<T, D>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                 (int)smem);                                                            \
        pq_encode_kernel<T, D><<<grid, CDIST_THREADS, smem, st>>>(z0, z1, table, codes0, codes1, total, m, c); \
        break;
    switch (dc) {
        SPT_ENC_CASE(4)
        SPT_ENC_CASE(8)
        SPT_ENC_CASE(16)
        SPT_ENC_CASE(32)
        default:
            return fail(SPT_ERR_UNSUPPORTED, "pq_encode: d_codeword %d not in {4,8,16,32}", dc);
    }
#undef SPT_ENC_CASE
    return after_launch("pq_encode_kernel");
}

}  // namespace spt

using namespace spt;

extern "C" int spt_cdist_fwd(const void *query, const float *table, float *distance, int32_t *indices, int m,
                             int64_t n, int c, int dc, int dtype, spt_stream_t stream) {
    SPT_REQUIRE(query && table && indices, "cdist_fwd: null pointer");
    SPT_REQUIRE(m >= 1 && n >= 0 && c >= 1 && dc >= 1 && dc <= 64, "cdist_fwd: bad sizes m=%d n=%lld c=%d dc=%d",
                m, (long long)n, c, dc);
    SPT_REQUIRE(m <= 65535, "cdist_fwd: n_subspaces %d exceeds grid limit", m);
    SPT_REQUIRE((size_t)c * dc * 4 <= 200 * 1024, "cdist_fwd: codebook slice %d x %d does not fit shared memory", c, dc);
    if (n == 0) return SPT_OK;
    if (dtype == SPT_F32)
        return launch_cdist_fwd((const float *)query, table, distance, indices, m, n, c, dc, as_stream(stream));
    if (dtype == SPT_BF16)
        return launch_cdist_fwd((const __nv_bfloat16 *)query, table, distance, indices, m, n, c, dc, as_stream(stream));
    return fail(SPT_ERR_INVALID_ARGUMENT, "cdist_fwd: unknown dtype %d", dtype);
}

static int pq_encode_impl(const void *z0, const void *z1, const float *table, int32_t *codes0, int32_t *codes1,
                          int64_t rows, int m, int c, int dc, int dtype, spt_stream_t stream) {
    SPT_REQUIRE(z0 && table && codes0 && (!z1 == !codes1), "pq_encode: null pointer");
    SPT_REQUIRE(m >= 1 && rows >= 0 && c >= 1 && dc >= 1, "pq_encode: bad sizes");
    SPT_REQUIRE((size_t)m * c * dc * 4 <= 200 * 1024, "pq_encode: codebook %d x %d x %d does not fit shared memory", m, c, dc);
    if (rows == 0) return SPT_OK;
    if (dtype == SPT_F32)
        return launch_pq_encode((const float *)z0, (const float *)z1, table, codes0, codes1, rows, m, c, dc, as_stream(stream));
    if (dtype == SPT_BF16)
        return launch_pq_encode((const __nv_bfloat16 *)z0, (const __nv_bfloat16 *)z1, table, codes0, codes1, rows, m, c,
                                dc, as_stream(stream));
    return fail(SPT_ERR_INVALID_ARGUMENT, "pq_encode: unknown dtype %d", dtype);
}

extern "C" int spt_pq_encode(const void *z, const float *table, int32_t *codes, int64_t rows, int m, int c, int dc,
                             int dtype, spt_stream_t stream) {
    return pq_encode_impl(z, nullptr, table, codes, nullptr, rows, m, c, dc, dtype, stream);
}

extern "C" int spt_pq_encode_pair(const void *z0, const void *z1, const float *table, int32_t *codes0, int32_t *codes1,
                                  int64_t rows, int m, int c, int dc, int dtype, spt_stream_t stream) {
    SPT_REQUIRE(z1 && codes1, "pq_encode_pair: null pointer");
    return pq_encode_impl(z0, z1, table, codes0, codes1, rows, m, c, dc, dtype, stream);
}

extern "C" size_t spt_cdist_bwd_workspace_bytes(int m, int64_t n, int c, int dc) {
    const int64_t n_chunks = (n + BWD_T_CHUNK - 1) / BWD_T_CHUNK;
    return (size_t)m * (size_t)n_chunks * c * dc * sizeof(float);
}

extern "C" int spt_cdist_bwd(const float *query, const float *table, const float *grad_distance, float *grad_query,
                             float *grad_table, void *workspace, int m, int64_t n, int c, int dc,
                             spt_stream_t stream) {
    SPT_REQUIRE(query && table && grad_distance && grad_query && grad_table, "cdist_bwd: null pointer");
    SPT_REQUIRE(m >= 1 && m <= 65535 && n >= 1 && c >= 1 && dc >= 1 && dc <= 64, "cdist_bwd: bad sizes");
    SPT_REQUIRE(c * dc <= 4 * CDIST_THREADS, "cdist_bwd: c*dc = %d exceeds %d", c * dc, 4 * CDIST_THREADS);
    SPT_REQUIRE(workspace, "cdist_bwd: workspace required");
    cudaStream_t st = as_stream(stream);
    {
        dim3 grid((unsigned)((n + CDIST_THREADS - 1) / CDIST_THREADS), m);
        size_t smem = (size_t)c * dc * sizeof(float);
        switch (dc) {
            case 4: cdist_bwd_query_kernel<4><<<grid, CDIST_THREADS, smem, st>>>(query, table, grad_distance, grad_query, n, c, dc); break;
            case 8: cdist_bwd_query_kernel<8><<<grid, CDIST_THREADS, smem, st>>>(query, table, grad_distance, grad_query, n, c, dc); break;
            case 16: cdist_bwd_query_kernel<16><<<grid, CDIST_THREADS, smem, st>>>(query, table, grad_distance, grad_query, n, c, dc); break;
            case 32: cdist_bwd_query_kernel<32><<<grid, CDIST_THREADS, smem, st>>>(query, table, grad_distance, grad_query, n, c, dc); break;
            default: cdist_bwd_query_kernel<0><<<grid, CDIST_THREADS, smem, st>>>(query, table, grad_distance, grad_query, n, c, dc);
        }
        SPT_LAUNCH_CHECK("cdist_bwd_query_kernel");
    }
    {
        const int n_chunks = (int)((n + BWD_T_CHUNK - 1) / BWD_T_CHUNK);
        dim3 grid(n_chunks, m);
        size_t smem = (size_t)BWD_T_TILE * (dc + c) * sizeof(float);
        if (smem > 48 * 1024)
            cudaFuncSetAttribute(cdist_bwd_table_stage1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cdist_bwd_table_stage1<<<grid, CDIST_THREADS, smem, st>>>(query, table, grad_distance, (float *)workspace, n,
                                                                  c, dc, n_chunks);
        SPT_LAUNCH_CHECK("cdist_bwd_table_stage1");
        dim3 grid2((c * dc + 127) / 128, m);
        cdist_bwd_table_stage2<<<grid2, 128, 0, st>>>((const float *)workspace, grad_table, c * dc, n_chunks);
        SPT_LAUNCH_CHECK("cdist_bwd_table_stage2");
    }
    return SPT_OK;
}
