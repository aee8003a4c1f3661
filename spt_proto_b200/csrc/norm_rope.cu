// The two elementwise layers that sit between the SPT operators in a LLaMA-style block, each as ONE kernel per
// direction (the reference, and the first version here, run them as chains of 5-10 torch elementwise kernels per
// call; at the fine-tuning shape of BASELINE configs[3] those chains were ~1.5 ms of a 15.7 ms step):
//
//   rmsnorm : LlamaRMSNorm (reference naive_gpt/layers/basic/utils.py:22-38), bf16 in / bf16 weight / bf16 out,
//             statistics in fp32, the same intermediate roundings as the torch expression:
//                 inv = rsqrt(mean(x^2) + eps);  y16 = bf16(x * inv);  out = bf16(w * y16)
//             backward: gy = bf16(g * w);  dx = bf16(gy * inv - x * inv^3 * sum_c(gy x) / C);
//                       dw partial[block][c] = sum over the block's rows of g * y16   (fp32, summed by the caller)
//   rope    : RotaryEmbedding (reference naive_gpt/layers/basic/position.py:5-48) on [N, S, H, E] bf16:
//                 out = bf16(bf16(x * cos) + bf16(rotate_half(x) * sin)),  rotate_half(x) = cat(-x_hi, x_lo)
//             backward is the transposed rotation: dx = bf16(bf16(g * cos) + rotate_half^T(bf16(g * sin))).
#include "common.cuh"

namespace spt {
namespace glue {

constexpr int NT = 256;
using bf = __nv_bfloat16;

__device__ __forceinline__ float bf_round(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

// block-wide sum, result broadcast to every thread
__device__ __forceinline__ float block_sum_all(float v, float *s_red) {
    v = warp_sum(v);
    __syncthreads();                       // s_red may still be read from the previous call
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.0f;
#pragma unroll
    for (int i = 0; i < NT / 32; ++i) t += s_red[i];
    return t;
}

// one block per row; CH = 16-byte chunks per thread (C <= CH * NT * 8)
template <int CH>
__global__ void __launch_bounds__(NT)
rmsnorm_fwd_kernel(const bf *__restrict__ x, const bf *__restrict__ w, bf *__restrict__ out, float *__restrict__ inv_rms,
                   int C, float eps) {
    __shared__ float s_red[NT / 32];
    const size_t row = blockIdx.x;
    float xv[CH][8];
    float ss = 0.0f;
#pragma unroll
    for (int u = 0; u < CH; ++u) {
        const int c = (u * NT + threadIdx.x) * 8;
        if (c < C) {
            Vec16<bf>::load(x + row * C + c, xv[u]);
#pragma unroll
            for (int i = 0; i < 8; ++i) ss = fmaf(xv[u][i], xv[u][i], ss);
        }
    }
    ss = block_sum_all(ss, s_red);
    const float inv = rsqrtf(ss / (float)C + eps);
    if (threadIdx.x == 0) inv_rms[row] = inv;
#pragma unroll
    for (int u = 0; u < CH; ++u) {
        const int c = (u * NT + threadIdx.x) * 8;
        if (c < C) {
            float wv[8], o[8];
            Vec16<bf>::load(w + c, wv);
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = wv[i] * bf_round(xv[u][i] * inv);
            Vec16<bf>::store(out + row * C + c, o);
        }
    }
}

// persistent blocks over rows; a thread keeps the weight-gradient partials of its columns in registers
template <int CH>
__global__ void __launch_bounds__(NT)
rmsnorm_bwd_kernel(const bf *__restrict__ g, const bf *__restrict__ x, const bf *__restrict__ w,
                   const float *__restrict__ inv_rms, bf *__restrict__ dx, float *__restrict__ dw_partial, int64_t R,
                   int C) {
    __shared__ float s_red[NT / 32];
    float wv[CH][8], dw[CH][8];
#pragma unroll
    for (int u = 0; u < CH; ++u) {
        const int c = (u * NT + threadIdx.x) * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) dw[u][i] = 0.0f, wv[u][i] = 0.0f;
        if (c < C) Vec16<bf>::load(w + c, wv[u]);
    }
    for (int64_t row = blockIdx.x; row < R; row += gridDim.x) {
        float xv[CH][8], gy[CH][8];
        const float inv = inv_rms[row];
        float dot = 0.0f;
#pragma unroll
        for (int u = 0; u < CH; ++u) {
            const int c = (u * NT + threadIdx.x) * 8;
            if (c < C) {
                float gv[8];
                Vec16<bf>::load(x + row * C + c, xv[u]);
                Vec16<bf>::load(g + row * C + c, gv);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    gy[u][i] = bf_round(gv[i] * wv[u][i]);
                    dot = fmaf(gy[u][i], xv[u][i], dot);
                    dw[u][i] = fmaf(gv[i], bf_round(xv[u][i] * inv), dw[u][i]);
                }
            }
        }
        dot = block_sum_all(dot, s_red);
        const float coef = dot * inv * inv * inv / (float)C;
#pragma unroll
        for (int u = 0; u < CH; ++u) {
            const int c = (u * NT + threadIdx.x) * 8;
            if (c < C) {
                float o[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i] = fmaf(gy[u][i], inv, -xv[u][i] * coef);
                Vec16<bf>::store(dx + row * C + c, o);
            }
        }
    }
#pragma unroll
    for (int u = 0; u < CH; ++u) {
        const int c = (u * NT + threadIdx.x) * 8;
        if (c < C) {
            *reinterpret_cast<float4 *>(dw_partial + (size_t)blockIdx.x * C + c) = make_float4(dw[u][0], dw[u][1], dw[u][2], dw[u][3]);
            *reinterpret_cast<float4 *>(dw_partial + (size_t)blockIdx.x * C + c + 4) = make_float4(dw[u][4], dw[u][5], dw[u][6], dw[u][7]);
        }
    }
}

// thread = one 8-element chunk of the low half of a (n, s, h) row and its partner in the high half
__global__ void __launch_bounds__(NT)
rope_kernel(const bf *__restrict__ x, const bf *__restrict__ cs, const bf *__restrict__ sn, bf *__restrict__ out,
            int64_t rows, int S, int H, int E, int transpose) {
    const int per_row = E / 16;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * per_row) return;
    const int64_t row = idx / per_row;
    const int e0 = (int)(idx % per_row) * 8, e1 = e0 + E / 2;
    const int s = (int)((row / H) % S);
    float xl[8], xh[8], cl[8], ch[8], sl[8], sh[8], ol[8], oh[8];
    Vec16<bf>::load(x + row * E + e0, xl);
    Vec16<bf>::load(x + row * E + e1, xh);
    Vec16<bf>::load(cs + (size_t)s * E + e0, cl);
    Vec16<bf>::load(cs + (size_t)s * E + e1, ch);
    Vec16<bf>::load(sn + (size_t)s * E + e0, sl);
    Vec16<bf>::load(sn + (size_t)s * E + e1, sh);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        if (!transpose) {   // out_lo = x_lo cos_lo - x_hi sin_lo ;  out_hi = x_hi cos_hi + x_lo sin_hi
            ol[i] = bf_round(xl[i] * cl[i]) + bf_round(-xh[i] * sl[i]);
            oh[i] = bf_round(xh[i] * ch[i]) + bf_round(xl[i] * sh[i]);
        } else {            // dx_lo = g_lo cos_lo + g_hi sin_hi ;  dx_hi = g_hi cos_hi - g_lo sin_lo
            ol[i] = bf_round(xl[i] * cl[i]) + bf_round(xh[i] * sh[i]);
            oh[i] = bf_round(xh[i] * ch[i]) - bf_round(xl[i] * sl[i]);
        }
    }
    Vec16<bf>::store(out + row * E + e0, ol);
    Vec16<bf>::store(out + row * E + e1, oh);
}

static int rms_chunks(int C) {   // 16-byte chunks per thread, 0 = unsupported
    if (C < 8 || C % 8 != 0) return 0;
    const int ch = (C / 8 + NT - 1) / NT;
    return ch <= 1 ? 1 : ch <= 2 ? 2 : ch <= 4 ? 4 : 0;
}
static int rms_bwd_grid(int64_t R) {
    const int64_t cap = (int64_t)num_sms() * 2;
    return (int)(R < cap ? R : cap);
}

}  // namespace glue
}  // namespace spt

using namespace spt;

extern "C" int spt_rmsnorm_bwd_blocks(int64_t R) { return R < 1 ? 0 : glue::rms_bwd_grid(R); }

extern "C" int spt_rmsnorm_fwd_bf16(const void *x, const void *w, void *out, float *inv_rms, int64_t R, int C, float eps,
                                    spt_stream_t stream) {
    SPT_REQUIRE(x && w && out && inv_rms, "rmsnorm_fwd: null pointer");
    const int ch = glue::rms_chunks(C);
    SPT_REQUIRE(R >= 1 && R <= 0x7fffffff && ch != 0, "rmsnorm_fwd: need 1 <= rows < 2^31 and C a multiple of 8 up to 8192 (R=%lld C=%d)",
                (long long)R, C);
    using glue::bf;
    cudaStream_t st = as_stream(stream);
    const unsigned grid = (unsigned)R;
    if (ch == 1) glue::rmsnorm_fwd_kernel<1><<<grid, glue::NT, 0, st>>>((const bf *)x, (const bf *)w, (bf *)out, inv_rms, C, eps);
    else if (ch == 2) glue::rmsnorm_fwd_kernel<2><<<grid, glue::NT, 0, st>>>((const bf *)x, (const bf *)w, (bf *)out, inv_rms, C, eps);
    else glue::rmsnorm_fwd_kernel<4><<<grid, glue::NT, 0, st>>>((const bf *)x, (const bf *)w, (bf *)out, inv_rms, C, eps);
    return after_launch("rmsnorm_fwd_kernel");
}

// dw_partial: [spt_rmsnorm_bwd_blocks(R), C] fp32, summed over dim 0 by the caller
extern "C" int spt_rmsnorm_bwd_bf16(const void *g, const void *x, const void *w, const float *inv_rms, void *dx,
                                    float *dw_partial, int64_t R, int C, spt_stream_t stream) {
    SPT_REQUIRE(g && x && w && inv_rms && dx && dw_partial, "rmsnorm_bwd: null pointer");
    const int ch = glue::rms_chunks(C);
    SPT_REQUIRE(R >= 1 && ch != 0, "rmsnorm_bwd: need rows >= 1 and C a multiple of 8 up to 8192 (R=%lld C=%d)", (long long)R, C);
    using glue::bf;
    cudaStream_t st = as_stream(stream);
    const unsigned grid = (unsigned)glue::rms_bwd_grid(R);
    if (ch == 1) glue::rmsnorm_bwd_kernel<1><<<grid, glue::NT, 0, st>>>((const bf *)g, (const bf *)x, (const bf *)w, inv_rms, (bf *)dx, dw_partial, R, C);
    else if (ch == 2) glue::rmsnorm_bwd_kernel<2><<<grid, glue::NT, 0, st>>>((const bf *)g, (const bf *)x, (const bf *)w, inv_rms, (bf *)dx, dw_partial, R, C);
    else glue::rmsnorm_bwd_kernel<4><<<grid, glue::NT, 0, st>>>((const bf *)g, (const bf *)x, (const bf *)w, inv_rms, (bf *)dx, dw_partial, R, C);
    return after_launch("rmsnorm_bwd_kernel");
}

// x, out: [N, S, H, E] bf16 (rows = N * S * H); cos, sin: [S, E] bf16 (already gathered by position)
extern "C" int spt_rope_bf16(const void *x, const void *cos, const void *sin, void *out, int64_t rows, int S, int H, int E,
                             int transpose, spt_stream_t stream) {
    SPT_REQUIRE(x && cos && sin && out, "rope: null pointer");
    SPT_REQUIRE(rows >= 1 && S >= 1 && H >= 1 && E >= 16 && E % 16 == 0 && rows % ((int64_t)S * H) == 0,
                "rope: need E a multiple of 16 and rows = N * S * H (rows=%lld S=%d H=%d E=%d)", (long long)rows, S, H, E);
    using glue::bf;
    const int64_t n = rows * (E / 16);
    SPT_REQUIRE((n + glue::NT - 1) / glue::NT < (1ll << 31), "rope: too many rows");
    glue::rope_kernel<<<(unsigned)((n + glue::NT - 1) / glue::NT), glue::NT, 0, as_stream(stream)>>>(
        (const bf *)x, (const bf *)cos, (const bf *)sin, (bf *)out, rows, S, H, E, transpose);
    return after_launch("rope_kernel");
}
