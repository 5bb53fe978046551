// Fused epilogues of the decode step (one token per sequence).  Each consumes the fp32 split-K partials of a swap-AB
// GEMM stored token-major, ws[split][token][feature] (GROMA_GF_PARTIAL_T), sums the splits in fixed order and applies
// what the reference does next, so a LLaMA layer needs 4 small launches instead of 7:
//   qkv  : + RoPE (rotate-half, position from a device scalar) + KV-cache append     (modeling_llama.py:138-168,225-289)
//   o    : + residual add + post-attention RMSNorm                                   (modeling_llama.py:292-340, 53-70)
//   g/u  : SwiGLU over interleaved (gate, up) features                               (modeling_llama.py:171-184)
//   down : + residual add + the NEXT layer's input RMSNorm (or the final norm)
// All kernels are programmatic-dependent-launch aware: they trigger their dependents at entry and wait for their
// producer before touching memory, so launch latency overlaps the previous kernel (they are a few microseconds long).
#include "ptx.cuh"
#include "decode_common.cuh"
#include "capi_common.h"
#include <cooperative_groups.h>

namespace gb {

__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ float block_sum_f(float v, float* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    float t = (l < nw) ? red[l] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    return t;
}

// x[b,:] = bf16(sum_s ws[s][b][:] + x[b,:]);  y[b,:] = w * bf16(x * rsqrt(mean(x^2)+eps)).  One CTA per token, N % 4 == 0.
template <int VPT>
__global__ void reduce_residual_rmsnorm_kernel(const float* __restrict__ ws, int S, int B, int N, __nv_bfloat16* __restrict__ x,
                                               const float* __restrict__ w, __nv_bfloat16* __restrict__ y, float eps) {
    pdl_trigger();
    pdl_wait();
    __shared__ float red[32];
    const int b = blockIdx.x;
    const int nvec = N >> 2;
    float h[VPT][4];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
        const int v = threadIdx.x + i * blockDim.x;
        h[i][0] = h[i][1] = h[i][2] = h[i][3] = 0.f;
        if (v < nvec) {
            for (int s = 0; s < S; ++s) {
                const float4 t = __ldcg(reinterpret_cast<const float4*>(ws + ((long long)s * B + b) * N + v * 4));
                h[i][0] += t.x; h[i][1] += t.y; h[i][2] += t.z; h[i][3] += t.w;
            }
            __nv_bfloat16* xp = x + (long long)b * N + v * 4;
            const uint2 xr = *reinterpret_cast<const uint2*>(xp);
            const float2 x01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&xr.x));
            const float2 x23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&xr.y));
            h[i][0] = bf16_round(h[i][0] + x01.x); h[i][1] = bf16_round(h[i][1] + x01.y);
            h[i][2] = bf16_round(h[i][2] + x23.x); h[i][3] = bf16_round(h[i][3] + x23.y);
            *reinterpret_cast<uint2*>(xp) = make_uint2(pack_bf16x2(h[i][0], h[i][1]), pack_bf16x2(h[i][2], h[i][3]));
            ss += h[i][0] * h[i][0] + h[i][1] * h[i][1] + h[i][2] * h[i][2] + h[i][3] * h[i][3];
        }
    }
    ss = block_sum_f(ss, red);
    const float rs = rsqrtf(ss / N + eps);
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
        const int v = threadIdx.x + i * blockDim.x;
        if (v < nvec) {
            const float4 w4 = *reinterpret_cast<const float4*>(w + v * 4);
            *reinterpret_cast<uint2*>(y + (long long)b * N + v * 4) =
                make_uint2(pack_bf16x2(w4.x * bf16_round(h[i][0] * rs), w4.y * bf16_round(h[i][1] * rs)),
                           pack_bf16x2(w4.z * bf16_round(h[i][2] * rs), w4.w * bf16_round(h[i][3] * rs)));
        }
    }
}

// Cluster version: RN_CL CTAs (one thread-block cluster) per token, each owning N/RN_CL features; the per-slice sums of
// squares are exchanged through distributed shared memory and added in rank order (deterministic), so the 13 x 16 KB of
// split-K partials of one token are pulled by 8 SMs instead of one and the kernel stays off the decode critical path.
constexpr int RN_CL = 8;
__global__ void __cluster_dims__(RN_CL, 1, 1) __launch_bounds__(128)
reduce_residual_rmsnorm_cluster_kernel(const float* __restrict__ ws, int S, int B, int N, __nv_bfloat16* __restrict__ x,
                                       const float* __restrict__ w, __nv_bfloat16* __restrict__ y, float eps) {
    pdl_trigger();
    cluster_arrive_relaxed();   // DSMEM rule: peers must be running before their shared memory is written (waited on below)
    pdl_wait();
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int crank = (int)cluster.block_rank();
    __shared__ float red[32];
    __shared__ float part[RN_CL];
    const int b = blockIdx.y;
    const int slice = N / RN_CL;                 // features per CTA (multiple of 4)
    const int v = threadIdx.x;                   // one float4 per thread: slice <= 512
    const int col = crank * slice + v * 4;
    const bool act = v * 4 < slice;
    float h0 = 0.f, h1 = 0.f, h2 = 0.f, h3 = 0.f;
    if (act) {
        for (int s = 0; s < S; ++s) {
            const float4 t = __ldcg(reinterpret_cast<const float4*>(ws + ((long long)s * B + b) * N + col));
            h0 += t.x; h1 += t.y; h2 += t.z; h3 += t.w;
        }
        __nv_bfloat16* xp = x + (long long)b * N + col;
        const uint2 xr = *reinterpret_cast<const uint2*>(xp);
        const float2 x01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&xr.x));
        const float2 x23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&xr.y));
        h0 = bf16_round(h0 + x01.x); h1 = bf16_round(h1 + x01.y); h2 = bf16_round(h2 + x23.x); h3 = bf16_round(h3 + x23.y);
        *reinterpret_cast<uint2*>(xp) = make_uint2(pack_bf16x2(h0, h1), pack_bf16x2(h2, h3));
    }
    const float ss_local = block_sum_f(act ? (h0 * h0 + h1 * h1 + h2 * h2 + h3 * h3) : 0.f, red);
    cluster_wait();
    if (threadIdx.x < RN_CL) {
        float* peer = cluster.map_shared_rank(part, threadIdx.x);
        peer[crank] = ss_local;                  // every CTA publishes its slice sum to all peers
    }
    cluster.sync();
    float ss = 0.f;
#pragma unroll
    for (int r = 0; r < RN_CL; ++r) ss += part[r];
    const float rs = rsqrtf(ss / N + eps);
    if (act) {
        const float4 w4 = *reinterpret_cast<const float4*>(w + col);
        *reinterpret_cast<uint2*>(y + (long long)b * N + col) =
            make_uint2(pack_bf16x2(w4.x * bf16_round(h0 * rs), w4.y * bf16_round(h1 * rs)),
                       pack_bf16x2(w4.z * bf16_round(h2 * rs), w4.w * bf16_round(h3 * rs)));
    }
}

// out[b, j] = bf16(silu(g) * u), (g, u) = sum_s ws[s][b][2j], ws[s][b][2j+1]
__global__ void reduce_swiglu_kernel(const float* __restrict__ ws, int S, int B, int N, __nv_bfloat16* __restrict__ out) {
    pdl_trigger();
    pdl_wait();
    const int NO = N >> 1;
    const long long total = (long long)B * (NO >> 1);   // two outputs (one float4 of partials) per thread
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int b = i / (NO >> 1);
        const int j2 = i - (long long)b * (NO >> 1);
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int s = 0; s < S; ++s) {
            const float4 t = __ldcg(reinterpret_cast<const float4*>(ws + ((long long)s * B + b) * N + j2 * 4));
            a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
        }
        *reinterpret_cast<uint32_t*>(out + (long long)b * NO + j2 * 2) = pack_bf16x2(silu(a.x) * a.y, silu(a.z) * a.w);
    }
}

// qkv partials [S][B][3*H*D] -> q_out[b, h*D + d] (RoPE), cache_k[b,h,pos,:] (RoPE), cache_v[b,h,pos,:]
__global__ void reduce_rope_kv_kernel(const float* __restrict__ ws, int S, int B, int H, int D, __nv_bfloat16* __restrict__ q_out,
                                      __nv_bfloat16* __restrict__ cache_k, __nv_bfloat16* __restrict__ cache_v,
                                      const float* __restrict__ cos_t, const float* __restrict__ sin_t,
                                      const int* __restrict__ pos_ptr, long long cap) {
    pdl_trigger();
    pdl_wait();
    const int half = D >> 1;
    const int N = 3 * H * D;
    const int pos = *pos_ptr;
    const long long total = (long long)B * H * half;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int j = i % half;
        long long r = i / half;
        const int h = r % H;
        const int b = r / H;
        float q1, q2, k1, k2, v1, v2;
        splitk_pair(ws, S, B, N, b, h * D, j, half, q1, q2);
        splitk_pair(ws, S, B, N, b, H * D + h * D, j, half, k1, k2);
        splitk_pair(ws, S, B, N, b, 2 * H * D + h * D, j, half, v1, v2);
        // the un-fused path stores qkv in bf16 before RoPE: keep that rounding point
        q1 = bf16_round(q1); q2 = bf16_round(q2); k1 = bf16_round(k1); k2 = bf16_round(k2);
        const float c = cos_t[(long long)pos * half + j], sn = sin_t[(long long)pos * half + j];
        __nv_bfloat16* qo = q_out + (long long)b * H * D + h * D;
        rope_pair(q1, q2, c, sn, qo[j], qo[j + half]);
        const long long co = (((long long)b * H + h) * cap + pos) * D;
        rope_pair(k1, k2, c, sn, cache_k[co + j], cache_k[co + j + half]);
        cache_v[co + j] = __float2bfloat16_rn(v1);
        cache_v[co + j + half] = __float2bfloat16_rn(v2);
    }
}

template <typename... KArgs, typename... Args>
static int launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t st, int pdl, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    if (pdl) { cfg.attrs = attr; cfg.numAttrs = 1; }
    return cudaLaunchKernelEx(&cfg, kernel, args...) == cudaSuccess ? GROMA_OK : GROMA_ERR_CUDA;
}

}  // namespace gb
using namespace gb;

GROMA_API int32_t groma_decode_reduce_norm(const float* ws, int32_t splits, int32_t B, int32_t N, void* x, const float* w,
                                           void* y, float eps, int32_t pdl, void* stream) {
    if (!ws || !x || !w || !y || splits < 1 || B <= 0 || N <= 0) return GROMA_ERR_ARG;
    if (N & 3) return GROMA_ERR_ALIGN;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int nvec = N >> 2;
    auto X = reinterpret_cast<__nv_bfloat16*>(x);
    auto Y = reinterpret_cast<__nv_bfloat16*>(y);
    if (N % (RN_CL * 4) == 0 && N / RN_CL <= 512 && N >= 1024)
        return launch_pdl(reduce_residual_rmsnorm_cluster_kernel, dim3(RN_CL, B), dim3(128), st, pdl, ws, splits, B, N, X, w, Y, eps);
    if (nvec <= 1024) return launch_pdl(reduce_residual_rmsnorm_kernel<1>, dim3(B), dim3(((nvec + 31) / 32) * 32), st, pdl, ws, splits, B, N, X, w, Y, eps);
    if (nvec <= 4096) return launch_pdl(reduce_residual_rmsnorm_kernel<4>, dim3(B), dim3(1024), st, pdl, ws, splits, B, N, X, w, Y, eps);
    return GROMA_ERR_UNSUPPORTED;
}

GROMA_API int32_t groma_decode_reduce_swiglu(const float* ws, int32_t splits, int32_t B, int32_t N, void* out, int32_t pdl,
                                             void* stream) {
    if (!ws || !out || splits < 1 || B <= 0 || N <= 0) return GROMA_ERR_ARG;
    if (N & 3) return GROMA_ERR_ALIGN;
    const long long total = (long long)B * (N >> 2);
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    return launch_pdl(reduce_swiglu_kernel, dim3(blocks), dim3(256), reinterpret_cast<cudaStream_t>(stream), pdl, ws, splits, B, N,
                      reinterpret_cast<__nv_bfloat16*>(out));
}

GROMA_API int32_t groma_decode_reduce_rope_kv(const float* ws, int32_t splits, int32_t B, int32_t H, int32_t D, void* q_out,
                                              void* cache_k, void* cache_v, const float* cos_t, const float* sin_t,
                                              const int32_t* pos_ptr, int64_t cap, int32_t pdl, void* stream) {
    if (!ws || !q_out || !cache_k || !cache_v || !cos_t || !sin_t || !pos_ptr || (D & 1)) return GROMA_ERR_ARG;
    const long long total = (long long)B * H * (D >> 1);
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    return launch_pdl(reduce_rope_kv_kernel, dim3(blocks), dim3(256), reinterpret_cast<cudaStream_t>(stream), pdl, ws, splits, B, H, D,
                      reinterpret_cast<__nv_bfloat16*>(q_out), reinterpret_cast<__nv_bfloat16*>(cache_k),
                      reinterpret_cast<__nv_bfloat16*>(cache_v), cos_t, sin_t, pos_ptr, (long long)cap);
}
