// Row-wise normalisation kernels (HBM-bound): RMSNorm, LayerNorm (+fused residual add), GroupNorm statistics and
// GroupNorm+ReLU apply over NHWC maps.  bf16 in/out, fp32 math, 16-byte vector loads, one warp-shuffle tree per row.
#include "ptx.cuh"
#include "gn_common.cuh"
#include "capi_common.h"

namespace gb {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum for blockDim.x <= 1024 (result broadcast to all threads).
__device__ __forceinline__ float block_sum(float v, float* red) {
    v = warp_sum(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    float t = (l < nw) ? red[l] : 0.f;
    t = warp_sum(t);
    return t;
}

// y = bf16( w * bf16( h * rsqrt(mean(h^2)+eps) ) ),  h = bf16(x + r) if r  (HF LlamaRMSNorm, modeling_llama.py:53-70)
// One block per row; dim % 8 == 0; VPT vectors of 8 per thread kept in registers.
template <int VPT>
__global__ void rmsnorm_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ r,
                               const float* __restrict__ w, __nv_bfloat16* __restrict__ y,
                               __nv_bfloat16* __restrict__ h_out, int dim, float eps) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // a PDL-launched GEMM may start prefetching its weights now
    __shared__ float red[32];
    const long long row = blockIdx.x;
    const int nvec = dim >> 3;
    float vals[VPT][8];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
        const int v = threadIdx.x + i * blockDim.x;
        if (v < nvec) {
            uint4 a = *reinterpret_cast<const uint4*>(x + row * dim + v * 8);
            const __nv_bfloat162* a2 = reinterpret_cast<const __nv_bfloat162*>(&a);
            uint4 b = make_uint4(0, 0, 0, 0);
            if (r) b = *reinterpret_cast<const uint4*>(r + row * dim + v * 8);
            const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&b);
            uint32_t hp[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                float2 f = __bfloat1622float2(a2[t]);
                if (r) {
                    const float2 g = __bfloat1622float2(b2[t]);
                    f.x = bf16_round(f.x + g.x);
                    f.y = bf16_round(f.y + g.y);
                }
                vals[i][2 * t] = f.x;
                vals[i][2 * t + 1] = f.y;
                ss += f.x * f.x + f.y * f.y;
                hp[t] = pack_bf16x2(f.x, f.y);
            }
            if (h_out) *reinterpret_cast<uint4*>(h_out + row * dim + v * 8) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
        }
    }
    ss = block_sum(ss, red);
    const float rs = rsqrtf(ss / dim + eps);
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
        const int v = threadIdx.x + i * blockDim.x;
        if (v < nvec) {
            const float4 w0 = *reinterpret_cast<const float4*>(w + v * 8);
            const float4 w1 = *reinterpret_cast<const float4*>(w + v * 8 + 4);
            const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
            uint32_t o[4];
#pragma unroll
            for (int t = 0; t < 4; ++t)
                o[t] = pack_bf16x2(ww[2 * t] * bf16_round(vals[i][2 * t] * rs), ww[2 * t + 1] * bf16_round(vals[i][2 * t + 1] * rs));
            *reinterpret_cast<uint4*>(y + row * dim + v * 8) = make_uint4(o[0], o[1], o[2], o[3]);
        }
    }
}

// y = LN(x (+ r)) * w + b ; biased variance; two-pass in registers.  One block per row.
template <int VPT>
__global__ void layernorm_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ r,
                                 const float* __restrict__ w, const float* __restrict__ b,
                                 __nv_bfloat16* __restrict__ y, int dim, float eps, long long x_row_stride,
                                 long long y_row_stride) {
    __shared__ float red[32];
    const long long row = blockIdx.x;
    const int nvec = dim >> 3;
    float vals[VPT][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
        const int v = threadIdx.x + i * blockDim.x;
        if (v < nvec) {
            uint4 a = *reinterpret_cast<const uint4*>(x + row * x_row_stride + v * 8);
            const __nv_bfloat162* a2 = reinterpret_cast<const __nv_bfloat162*>(&a);
            uint4 c = make_uint4(0, 0, 0, 0);
            if (r) c = *reinterpret_cast<const uint4*>(r + row * x_row_stride + v * 8);
            const __nv_bfloat162* c2 = reinterpret_cast<const __nv_bfloat162*>(&c);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                float2 f = __bfloat1622float2(a2[t]);
                if (r) {
                    const float2 g = __bfloat1622float2(c2[t]);
                    f.x = bf16_round(f.x + g.x);
                    f.y = bf16_round(f.y + g.y);
                }
                vals[i][2 * t] = f.x;
                vals[i][2 * t + 1] = f.y;
                s += f.x + f.y;
            }
        }
    }
    const float mean = block_sum(s, red) / dim;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
        const int v = threadIdx.x + i * blockDim.x;
        if (v < nvec) {
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const float d = vals[i][t] - mean;
                q += d * d;
            }
        }
    }
    const float rstd = rsqrtf(block_sum(q, red) / dim + eps);
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
        const int v = threadIdx.x + i * blockDim.x;
        if (v < nvec) {
            float ww[8], bb[8];
            *reinterpret_cast<float4*>(ww) = *reinterpret_cast<const float4*>(w + v * 8);
            *reinterpret_cast<float4*>(ww + 4) = *reinterpret_cast<const float4*>(w + v * 8 + 4);
            *reinterpret_cast<float4*>(bb) = *reinterpret_cast<const float4*>(b + v * 8);
            *reinterpret_cast<float4*>(bb + 4) = *reinterpret_cast<const float4*>(b + v * 8 + 4);
            uint32_t o[4];
#pragma unroll
            for (int t = 0; t < 4; ++t)
                o[t] = pack_bf16x2((vals[i][2 * t] - mean) * rstd * ww[2 * t] + bb[2 * t],
                                   (vals[i][2 * t + 1] - mean) * rstd * ww[2 * t + 1] + bb[2 * t + 1]);
            *reinterpret_cast<uint4*>(y + row * y_row_stride + v * 8) = make_uint4(o[0], o[1], o[2], o[3]);
        }
    }
}

// GroupNorm statistics over NHWC bf16 maps [B, P, C] (P pixels), G groups of C/G consecutive channels.
// Deterministic two-stage: partial sums per (b, pixel-chunk) then a fixed-order combine.
// stage 1: grid (chunks, B); each block reduces `pix_per_chunk` pixels for all groups.
__global__ void gn_partial_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ part, int P, int C, int G,
                                  int pix_per_chunk) {
    // thread t handles channel vector (8 ch) v = t % (C/8) ... blockDim.x = C/8 * rows_par
    const int nvec = C >> 3;
    const int rows_par = blockDim.x / nvec;
    const int v = threadIdx.x % nvec, rr = threadIdx.x / nvec;
    const int b = blockIdx.y, chunk = blockIdx.x;
    const int p0 = chunk * pix_per_chunk, p1 = min(P, p0 + pix_per_chunk);
    float s = 0.f, q = 0.f;
    for (int pix = p0 + rr; pix < p1; pix += rows_par) {
        const uint4 a = *reinterpret_cast<const uint4*>(x + ((long long)b * P + pix) * C + v * 8);
        const __nv_bfloat162* a2 = reinterpret_cast<const __nv_bfloat162*>(&a);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const float2 f = __bfloat1622float2(a2[t]);
            s += f.x + f.y;
            q += f.x * f.x + f.y * f.y;
        }
    }
    // reduce over the threads that share a group: cpg = C/G channels = cpg/8 vectors (consecutive v), and rows_par
    extern __shared__ float sh[];  // [2][blockDim.x]
    sh[threadIdx.x] = s;
    sh[blockDim.x + threadIdx.x] = q;
    __syncthreads();
    const int vpg = (C / G) >> 3;  // vectors per group
    if (threadIdx.x < G) {
        const int g = threadIdx.x;
        float ts = 0.f, tq = 0.f;
        for (int r2 = 0; r2 < rows_par; ++r2)
            for (int k = 0; k < vpg; ++k) {
                ts += sh[r2 * nvec + g * vpg + k];
                tq += sh[blockDim.x + r2 * nvec + g * vpg + k];
            }
        float* dst = part + (((long long)b * gridDim.x + chunk) * G + g) * 2;
        dst[0] = ts;
        dst[1] = tq;
    }
}
// stage 2: stats[b][g] = {mean, rstd}
__global__ void gn_finalize_kernel(const float* __restrict__ part, float* __restrict__ stats, int chunks, int G,
                                   float count, float eps) {
    const int b = blockIdx.x, g = threadIdx.x;
    if (g >= G) return;
    double s = 0.0, q = 0.0;
    for (int c = 0; c < chunks; ++c) {
        const float* src = part + (((long long)b * chunks + c) * G + g) * 2;
        s += src[0];
        q += src[1];
    }
    const double mean = s / count;
    double var = q / count - mean * mean;
    if (var < 0) var = 0;
    stats[(b * G + g) * 2] = (float)mean;
    stats[(b * G + g) * 2 + 1] = (float)(1.0 / sqrt(var + (double)eps));
}
// y = relu((x-mean)*rstd*gamma+beta)
__global__ void gn_relu_apply_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ stats,
                                     const float* __restrict__ gamma, const float* __restrict__ beta,
                                     __nv_bfloat16* __restrict__ y, long long P, int C, int G, long long total_vec) {
    const int nvec = C >> 3;
    const int cpg = C / G;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total_vec;
         i += (long long)gridDim.x * blockDim.x) {
        const int v = i % nvec;
        const long long pix = i / nvec;
        const int b = pix / P;
        const int g = (v * 8) / cpg;
        const float mean = stats[(b * G + g) * 2], rstd = stats[(b * G + g) * 2 + 1];
        const uint4 a = *reinterpret_cast<const uint4*>(x + i * 8);
        *reinterpret_cast<uint4*>(y + i * 8) = gn_relu8(a, mean, rstd, gn_load_affine8(gamma, beta, v * 8));
    }
}

}  // namespace gb
using namespace gb;

GROMA_API int32_t groma_rmsnorm(const void* x, const void* residual, const float* w, void* y, void* h_out,
                                int64_t rows, int32_t dim, float eps, void* stream) {
    if (!x || !w || !y || rows <= 0 || dim <= 0) return GROMA_ERR_ARG;
    if (dim & 7) return GROMA_ERR_ALIGN;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int nvec = dim >> 3;
    auto X = reinterpret_cast<const __nv_bfloat16*>(x);
    auto R = reinterpret_cast<const __nv_bfloat16*>(residual);
    auto Y = reinterpret_cast<__nv_bfloat16*>(y);
    auto H = reinterpret_cast<__nv_bfloat16*>(h_out);
    if (nvec <= 128) rmsnorm_kernel<1><<<(unsigned)rows, ((nvec + 31) / 32) * 32, 0, st>>>(X, R, w, Y, H, dim, eps);
    else if (nvec <= 512) rmsnorm_kernel<2><<<(unsigned)rows, ((nvec + 63) / 64) * 32, 0, st>>>(X, R, w, Y, H, dim, eps);
    else if (nvec <= 2048) rmsnorm_kernel<4><<<(unsigned)rows, 512, 0, st>>>(X, R, w, Y, H, dim, eps);
    else return GROMA_ERR_UNSUPPORTED;
    return GROMA_LAUNCH_CHECK();
}

GROMA_API int32_t groma_layernorm(const void* x, const void* residual, const float* w, const float* b, void* y,
                                  int64_t rows, int32_t dim, float eps, int64_t x_row_stride, int64_t y_row_stride,
                                  void* stream) {
    if (!x || !w || !b || !y || rows <= 0 || dim <= 0) return GROMA_ERR_ARG;
    if ((dim & 7) || (x_row_stride & 7) || (y_row_stride & 7)) return GROMA_ERR_ALIGN;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int nvec = dim >> 3;
    auto X = reinterpret_cast<const __nv_bfloat16*>(x);
    auto R = reinterpret_cast<const __nv_bfloat16*>(residual);
    auto Y = reinterpret_cast<__nv_bfloat16*>(y);
    if (nvec <= 128)
        layernorm_kernel<1><<<(unsigned)rows, ((nvec + 31) / 32) * 32, 0, st>>>(X, R, w, b, Y, dim, eps, x_row_stride, y_row_stride);
    else if (nvec <= 512)
        layernorm_kernel<2><<<(unsigned)rows, ((nvec + 63) / 64) * 32, 0, st>>>(X, R, w, b, Y, dim, eps, x_row_stride, y_row_stride);
    else return GROMA_ERR_UNSUPPORTED;
    return GROMA_LAUNCH_CHECK();
}

// GroupNorm(G) + ReLU over NHWC [B, P, C]; `part` is fp32 scratch of B*chunks*G*2, `stats` fp32 [B,G,2].
static int32_t gn_check(int32_t B, int64_t P, int32_t C, int32_t G) {
    if (B <= 0 || P <= 0 || C <= 0 || G <= 0) return GROMA_ERR_ARG;
    if ((C & 7) || (C % G) || ((C / G) & 7)) return GROMA_ERR_ALIGN;
    return GROMA_OK;
}
GROMA_API int32_t groma_groupnorm_stats(const void* x, float* part, float* stats, int32_t B, int64_t P, int32_t C, int32_t G,
                                        float eps, int32_t chunks, void* stream) {
    if (!x || !part || !stats) return GROMA_ERR_ARG;
    if (int32_t rc = gn_check(B, P, C, G)) return rc;
    if (chunks < 1) return GROMA_ERR_ALIGN;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int nvec = C >> 3;
    int rows_par = 1024 / nvec;
    if (rows_par < 1) return GROMA_ERR_UNSUPPORTED;
    if (rows_par > 8) rows_par = 8;
    const int threads = nvec * rows_par;
    if (G > threads) return GROMA_ERR_UNSUPPORTED;
    const int ppc = (int)((P + chunks - 1) / chunks);
    gn_partial_kernel<<<dim3(chunks, B), threads, 2 * threads * sizeof(float), st>>>(
        reinterpret_cast<const __nv_bfloat16*>(x), part, (int)P, C, G, ppc);
    gn_finalize_kernel<<<B, ((G + 31) / 32) * 32, 0, st>>>(part, stats, chunks, G, (float)((double)P * (C / G)), eps);
    return GROMA_LAUNCH_CHECK();
}
GROMA_API int32_t groma_groupnorm_apply_relu(const void* x, const float* stats, const float* gamma, const float* beta, void* y,
                                             int32_t B, int64_t P, int32_t C, int32_t G, void* stream) {
    if (!x || !stats || !gamma || !beta || !y) return GROMA_ERR_ARG;
    if (int32_t rc = gn_check(B, P, C, G)) return rc;
    const long long total_vec = (long long)B * P * (C >> 3);
    int blocks = (int)((total_vec + 255) / 256);
    if (blocks > 148 * 32) blocks = 148 * 32;
    gn_relu_apply_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const __nv_bfloat16*>(x), stats, gamma, beta, reinterpret_cast<__nv_bfloat16*>(y), P, C, G, total_vec);
    return GROMA_LAUNCH_CHECK();
}
GROMA_API int32_t groma_groupnorm_relu(const void* x, const float* gamma, const float* beta, void* y, float* part,
                                       float* stats, int32_t B, int64_t P, int32_t C, int32_t G, float eps,
                                       int32_t chunks, void* stream) {
    if (!gamma || !beta || !y) return GROMA_ERR_ARG;
    if (int32_t rc = groma_groupnorm_stats(x, part, stats, B, P, C, G, eps, chunks, stream)) return rc;
    return groma_groupnorm_apply_relu(x, stats, gamma, beta, y, B, P, C, G, stream);
}
