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
__global__ void __launch_bounds__(1024) reduce_residual_rmsnorm_kernel(const float* __restrict__ ws, int S, int B, int N, __nv_bfloat16* __restrict__ x,
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
            const float4 t = splitk_sum4<4>(ws + (long long)b * N + v * 4, (long long)B * N, S);   // 1024-thread blocks: keep the register budget
            h[i][0] = t.x; h[i][1] = t.y; h[i][2] = t.z; h[i][3] = t.w;
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
    // the norm weight is a parameter, not a product of the previous kernel: fetch it while that kernel is still running (it used
    // to be loaded after the cluster barrier, one exposed L2 round trip on the critical path of each of the 64 calls per step)
    float4 w4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (act) w4 = __ldg(reinterpret_cast<const float4*>(w + col));
    pdl_wait();
    float h0 = 0.f, h1 = 0.f, h2 = 0.f, h3 = 0.f;
    if (act) {
        __nv_bfloat16* xp = x + (long long)b * N + col;
        const uint2 xr = *reinterpret_cast<const uint2*>(xp);     // in flight together with the partials
        const float4 t = splitk_sum4(ws + (long long)b * N + col, (long long)B * N, S);
        h0 = t.x; h1 = t.y; h2 = t.z; h3 = t.w;
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
        const float4 a = splitk_sum4(ws + (long long)b * N + j2 * 4, (long long)B * N, S);
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
        const int cols[3] = {h * D, H * D + h * D, 2 * H * D + h * D};
        float s1[3], s2[3];
        splitk_pairs<3>(ws, S, B, N, b, cols, j, half, s1, s2);
        float q1 = s1[0], q2 = s2[0], k1 = s1[1], k2 = s2[1];
        const float v1 = s1[2], v2 = s2[2];
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

// Tail of the step in one launch (was splitk_reduce + argmax + decode_advance): logits[b, :] = sum_s ws[s][b][:] (split order, so
// bit-identical to the generic reduce), greedy argmax with torch.argmax's first-index tie-break written to ids[b] in place, and
// the position / length bookkeeping for the next step.  HA_CL CTAs (one cluster) per row share the vocabulary; the per-CTA
// candidates meet in rank 0's shared memory through DSMEM.
constexpr int HA_CL = 8;
__device__ __forceinline__ void argmax_merge(float& best, int& bi, float ov, int oi) {
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
}
__global__ void __cluster_dims__(HA_CL, 1, 1) __launch_bounds__(512)
reduce_head_argmax_cluster_kernel(const float* __restrict__ ws, int S, int B, int V, float* __restrict__ logits,
                                  long long* __restrict__ ids, int* __restrict__ pos, int* __restrict__ kv_len) {
    pdl_trigger();
    cluster_arrive_relaxed();
    pdl_wait();
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int crank = (int)cluster.block_rank();
    __shared__ float sv[32];
    __shared__ int si[32];
    __shared__ float pv[HA_CL];
    __shared__ int pi[HA_CL];
    const int b = blockIdx.y;
    const int per = (V + HA_CL - 1) / HA_CL;
    const int lo = crank * per, hi = min(V, lo + per);
    float best = -INFINITY;
    int bi = 0x7fffffff;
    constexpr int C = 4, U = 8;   // columns per thread per pass x splits in flight
    for (int c0 = lo + threadIdx.x; c0 < hi; c0 += C * blockDim.x) {
        float acc[C];
#pragma unroll
        for (int c = 0; c < C; ++c) acc[c] = 0.f;
        for (int s0 = 0; s0 < S; s0 += U) {
            float t[U][C];
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const int col = c0 + c * blockDim.x;
                    if (s0 + u < S && col < hi) t[u][c] = __ldcg(ws + ((long long)(s0 + u) * B + b) * V + col);
                }
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int c = 0; c < C; ++c)
                    if (s0 + u < S && c0 + c * blockDim.x < hi) acc[c] += t[u][c];
        }
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const int col = c0 + c * blockDim.x;
            if (col < hi) {
                logits[(long long)b * V + col] = acc[c];
                argmax_merge(best, bi, acc[c], col);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) argmax_merge(best, bi, __shfl_xor_sync(0xffffffffu, best, o), __shfl_xor_sync(0xffffffffu, bi, o));
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { sv[w] = best; si[w] = bi; }
    __syncthreads();
    cluster_wait();
    if (w == 0) {
        const int nw = blockDim.x >> 5;
        best = l < nw ? sv[l] : -INFINITY;
        bi = l < nw ? si[l] : 0x7fffffff;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) argmax_merge(best, bi, __shfl_xor_sync(0xffffffffu, best, o), __shfl_xor_sync(0xffffffffu, bi, o));
        if (l == 0) {
            *cluster.map_shared_rank(&pv[crank], 0) = best;
            *cluster.map_shared_rank(&pi[crank], 0) = bi;
        }
    }
    cluster.sync();
    if (crank == 0 && threadIdx.x == 0) {
        best = -INFINITY; bi = 0x7fffffff;
#pragma unroll
        for (int r = 0; r < HA_CL; ++r) argmax_merge(best, bi, pv[r], pi[r]);
        ids[b] = bi;
        kv_len[b] += 1;
        if (b == 0) *pos += 1;
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

GROMA_API int32_t groma_decode_head_argmax(const float* ws, int32_t splits, int32_t B, int32_t V, float* logits, int64_t* ids,
                                           int32_t* pos, int32_t* kv_len, int32_t pdl, void* stream) {
    if (!ws || !logits || !ids || !pos || !kv_len || splits < 1 || B <= 0 || V <= 0) return GROMA_ERR_ARG;
    return launch_pdl(reduce_head_argmax_cluster_kernel, dim3(HA_CL, B), dim3(512), reinterpret_cast<cudaStream_t>(stream), pdl, ws,
                      splits, B, V, logits, reinterpret_cast<long long*>(ids), pos, kv_len);
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
