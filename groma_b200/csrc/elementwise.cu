// Data-movement / pointwise kernels of the Groma forward path (all HBM-bound; 16-byte vectors, grid-stride).
#include "ptx.cuh"
#include "decode_common.cuh"
#include "capi_common.h"

namespace gb {

static inline int grid_for(long long n, int threads) {
    long long b = (n + threads - 1) / threads;
    if (b > 148LL * 32) b = 148LL * 32;
    if (b < 1) b = 1;
    return (int)b;
}

// ---- ViT patchify: images [B,3,S,S] (fp32) -> patches [B*(S/14)^2, ld] bf16, k = c*196 + ky*14 + kx, zero pad to ld
__global__ void patchify_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ out, int B, int S, int ld) {
    const int G = S / 14;
    const long long total = (long long)B * G * G * ld;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int k = i % ld;
        const long long t = i / ld;
        float v = 0.f;
        if (k < 588) {
            const int c = k / 196, r = k % 196, ky = r / 14, kx = r % 14;
            const int px = t % G, py = (t / G) % G, b = t / ((long long)G * G);
            v = img[(((long long)b * 3 + c) * S + (py * 14 + ky)) * S + (px * 14 + kx)];
        }
        out[i] = __float2bfloat16_rn(v);
    }
}

// ---- tokens[b,0,:] = cls + pos[0]; tokens[b,1+p,:] = patch[b,p,:] + pos[1+p]   (modeling_dinov2.py:96-116)
__global__ void vit_embed_kernel(const __nv_bfloat16* __restrict__ patch, const float* __restrict__ cls,
                                 const float* __restrict__ pos, __nv_bfloat16* __restrict__ out, int B, int NP, int C) {
    const long long total = (long long)B * (NP + 1) * C;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = i % C;
        const long long t = i / C;
        const int tok = t % (NP + 1), b = t / (NP + 1);
        float v = pos[(long long)tok * C + c];
        if (tok == 0) v += cls[c];
        else v += __bfloat162float(patch[((long long)b * NP + tok - 1) * C + c]);
        out[i] = __float2bfloat16_rn(v);
    }
}

// ---- mean of n (<=4) token tensors [B, T, C] with the first `skip` tokens dropped -> [B, T-skip, C]
__global__ void mean_tokens_kernel(const __nv_bfloat16* __restrict__ a0, const __nv_bfloat16* __restrict__ a1,
                                   const __nv_bfloat16* __restrict__ a2, const __nv_bfloat16* __restrict__ a3, int n,
                                   __nv_bfloat16* __restrict__ out, int B, int T, int C, int skip) {
    const int nvec = C >> 3;
    const long long total = (long long)B * (T - skip) * nvec;
    const __nv_bfloat16* src[4] = {a0, a1, a2, a3};
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int v = i % nvec;
        const long long t = i / nvec;
        const int tok = t % (T - skip), b = t / (T - skip);
        const long long off = ((long long)b * T + tok + skip) * C + v * 8;
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int k = 0; k < n; ++k) {
            const uint4 a = *reinterpret_cast<const uint4*>(src[k] + off);
            const __nv_bfloat162* a2v = reinterpret_cast<const __nv_bfloat162*>(&a);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float2 f = __bfloat1622float2(a2v[q]);
                acc[2 * q] += f.x;
                acc[2 * q + 1] += f.y;
            }
        }
        const float inv = 1.0f / n;
        *reinterpret_cast<uint4*>(out + i * 8) =
            make_uint4(pack_bf16x2(acc[0] * inv, acc[1] * inv), pack_bf16x2(acc[2] * inv, acc[3] * inv),
                       pack_bf16x2(acc[4] * inv, acc[5] * inv), pack_bf16x2(acc[6] * inv, acc[7] * inv));
    }
}

// ---- 2x2 space-to-depth of the patch tokens (groma.py:227-237): in [B, 1+g*g, C] (CLS first) ->
//      out [B, (g/2)^2, 4C], channel blocks ordered (row,col) parity (0,0),(1,0),(0,1),(1,1)
__global__ void space_to_depth_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, int B,
                                      int g, int C) {
    const int nvec = C >> 3, h = g / 2;
    const long long total = (long long)B * h * h * 4 * nvec;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int v = i % nvec;
        long long t = i / nvec;
        const int blk = t % 4; t /= 4;
        const int ox = t % h, oy = (t / h) % h, b = t / ((long long)h * h);
        const int dy = (blk == 1 || blk == 3) ? 1 : 0, dx = (blk >= 2) ? 1 : 0;
        const long long src = ((long long)b * (g * g + 1) + 1 + (2 * oy + dy) * g + (2 * ox + dx)) * C + v * 8;
        *reinterpret_cast<uint4*>(out + i * 8) = *reinterpret_cast<const uint4*>(in + src);
    }
}

// ---- row gather / scatter of bf16 rows (embedding lookup, visual-token splice)
//      gather:  out[i,:] = (idx[i] < split ? t0[idx[i]] : t1[idx[i]-split])
__global__ void gather_rows_kernel(const long long* __restrict__ idx, const __nv_bfloat16* __restrict__ t0,
                                   const __nv_bfloat16* __restrict__ t1, long long split,
                                   __nv_bfloat16* __restrict__ out, long long n, int D) {
    const int nvec = D >> 3;
    const long long total = n * nvec;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int v = i % nvec;
        const long long r = i / nvec;
        const long long id = idx[r];
        const __nv_bfloat16* src = (id < split || !t1) ? t0 + id * D : t1 + (id - split) * D;
        *reinterpret_cast<uint4*>(out + r * D + v * 8) = *reinterpret_cast<const uint4*>(src + v * 8);
    }
}
//      scatter: out[idx[i],:] = src[i,:]
__global__ void scatter_rows_kernel(const long long* __restrict__ idx, const __nv_bfloat16* __restrict__ src,
                                    __nv_bfloat16* __restrict__ out, long long n, int D) {
    const int nvec = D >> 3;
    const long long total = n * nvec;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int v = i % nvec;
        const long long r = i / nvec;
        *reinterpret_cast<uint4*>(out + idx[r] * D + v * 8) = *reinterpret_cast<const uint4*>(src + r * D + v * 8);
    }
}

// ---- c = bf16(a + b) (a, b bf16; b may be fp32-broadcast rows handled elsewhere)
__global__ void add_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                           __nv_bfloat16* __restrict__ c, long long nvec) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
        const uint4 x = *reinterpret_cast<const uint4*>(a + i * 8);
        const uint4 y = *reinterpret_cast<const uint4*>(b + i * 8);
        const __nv_bfloat162* x2 = reinterpret_cast<const __nv_bfloat162*>(&x);
        const __nv_bfloat162* y2 = reinterpret_cast<const __nv_bfloat162*>(&y);
        uint32_t o[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const float2 f = __bfloat1622float2(x2[t]), g = __bfloat1622float2(y2[t]);
            o[t] = pack_bf16x2(f.x + g.x, f.y + g.y);
        }
        *reinterpret_cast<uint4*>(c + i * 8) = make_uint4(o[0], o[1], o[2], o[3]);
    }
}
// ---- c[r,:] = bf16(a[r,:] + b[r % period,:])  (positional-embedding broadcast add)
__global__ void add_bcast_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                                 __nv_bfloat16* __restrict__ c, long long rows, long long period, int D) {
    const int nvec = D >> 3;
    const long long total = rows * nvec;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int v = i % nvec;
        const long long r = i / nvec;
        const uint4 x = *reinterpret_cast<const uint4*>(a + i * 8);
        const uint4 y = *reinterpret_cast<const uint4*>(b + (r % period) * D + v * 8);
        const __nv_bfloat162* x2 = reinterpret_cast<const __nv_bfloat162*>(&x);
        const __nv_bfloat162* y2 = reinterpret_cast<const __nv_bfloat162*>(&y);
        uint32_t o[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const float2 f = __bfloat1622float2(x2[t]), g = __bfloat1622float2(y2[t]);
            o[t] = pack_bf16x2(f.x + g.x, f.y + g.y);
        }
        *reinterpret_cast<uint4*>(c + i * 8) = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

// ---- LLaMA rotate-half RoPE on the fused QKV rows + KV-cache append (modeling_llama.py:138-168)
//   qkv [B*T, 3*H*D] bf16 ; q_out [B*T, H*D] ; cache_k/v [B, H, ctx_cap, D]; position = pos0 + t
//   cos/sin tables fp32 [max_pos, D/2]
__global__ void rope_kv_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ q_out,
                               __nv_bfloat16* __restrict__ cache_k, __nv_bfloat16* __restrict__ cache_v,
                               const float* __restrict__ cos_t, const float* __restrict__ sin_t, int B, int T, int H,
                               int D, int pos0, const int* __restrict__ pos_ptr, long long ctx_cap) {
    const int half = D / 2;
    if (pos_ptr) pos0 = *pos_ptr;
    const int hv = half >> 3;   // 16-byte vectors per half head
    const long long total = (long long)B * T * H * hv;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int jv = i % hv;
        long long r = i / hv;
        const int h = r % H; r /= H;
        const int t = r % T, b = r / T;
        const int pos = pos0 + t, j = jv * 8;
        float c[8], s[8];
        *reinterpret_cast<float4*>(c) = *reinterpret_cast<const float4*>(cos_t + (long long)pos * half + j);
        *reinterpret_cast<float4*>(c + 4) = *reinterpret_cast<const float4*>(cos_t + (long long)pos * half + j + 4);
        *reinterpret_cast<float4*>(s) = *reinterpret_cast<const float4*>(sin_t + (long long)pos * half + j);
        *reinterpret_cast<float4*>(s + 4) = *reinterpret_cast<const float4*>(sin_t + (long long)pos * half + j + 4);
        const long long row = ((long long)b * T + t) * 3 * H * D;
        const __nv_bfloat16* qp = qkv + row + h * D + j;
        const __nv_bfloat16* kp = qkv + row + (long long)H * D + h * D + j;
        const __nv_bfloat16* vp = qkv + row + 2LL * H * D + h * D + j;
        const uint4 q1v = *reinterpret_cast<const uint4*>(qp), q2v = *reinterpret_cast<const uint4*>(qp + half);
        const uint4 k1v = *reinterpret_cast<const uint4*>(kp), k2v = *reinterpret_cast<const uint4*>(kp + half);
        const __nv_bfloat162* q1 = reinterpret_cast<const __nv_bfloat162*>(&q1v);
        const __nv_bfloat162* q2 = reinterpret_cast<const __nv_bfloat162*>(&q2v);
        const __nv_bfloat162* k1 = reinterpret_cast<const __nv_bfloat162*>(&k1v);
        const __nv_bfloat162* k2 = reinterpret_cast<const __nv_bfloat162*>(&k2v);
        uint32_t qa[4], qb[4], ka[4], kb[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float2 a = __bfloat1622float2(q1[u]), bq = __bfloat1622float2(q2[u]);
            const float2 e = __bfloat1622float2(k1[u]), f = __bfloat1622float2(k2[u]);
            __nv_bfloat16 r[8];
            rope_pair(a.x, bq.x, c[2 * u], s[2 * u], r[0], r[2]);
            rope_pair(a.y, bq.y, c[2 * u + 1], s[2 * u + 1], r[1], r[3]);
            rope_pair(e.x, f.x, c[2 * u], s[2 * u], r[4], r[6]);
            rope_pair(e.y, f.y, c[2 * u + 1], s[2 * u + 1], r[5], r[7]);
            qa[u] = (uint32_t)__bfloat16_as_ushort(r[0]) | ((uint32_t)__bfloat16_as_ushort(r[1]) << 16);
            qb[u] = (uint32_t)__bfloat16_as_ushort(r[2]) | ((uint32_t)__bfloat16_as_ushort(r[3]) << 16);
            ka[u] = (uint32_t)__bfloat16_as_ushort(r[4]) | ((uint32_t)__bfloat16_as_ushort(r[5]) << 16);
            kb[u] = (uint32_t)__bfloat16_as_ushort(r[6]) | ((uint32_t)__bfloat16_as_ushort(r[7]) << 16);
        }
        __nv_bfloat16* qo = q_out + ((long long)b * T + t) * H * D + h * D + j;
        *reinterpret_cast<uint4*>(qo) = make_uint4(qa[0], qa[1], qa[2], qa[3]);
        *reinterpret_cast<uint4*>(qo + half) = make_uint4(qb[0], qb[1], qb[2], qb[3]);
        const long long co = (((long long)b * H + h) * ctx_cap + pos) * D + j;
        *reinterpret_cast<uint4*>(cache_k + co) = make_uint4(ka[0], ka[1], ka[2], ka[3]);
        *reinterpret_cast<uint4*>(cache_k + co + half) = make_uint4(kb[0], kb[1], kb[2], kb[3]);
        *reinterpret_cast<uint4*>(cache_v + co) = *reinterpret_cast<const uint4*>(vp);
        *reinterpret_cast<uint4*>(cache_v + co + half) = *reinterpret_cast<const uint4*>(vp + half);
    }
}

// ---- tiny-K linear in fp32: out[m,n] = act(sum_k x[m,k]*w[n,k] + b[n]) -> bf16  (roi_align.py:255 Linear(4,256))
__global__ void linear_smallk_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                                     __nv_bfloat16* __restrict__ out, long long M, int N, int K, int relu) {
    const long long total = M * N;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int n = i % N;
        const long long m = i / N;
        float acc = 0.f;
        for (int k = 0; k < K; ++k) acc += x[m * K + k] * w[(long long)n * K + k];
        if (b) acc += b[n];
        if (relu) acc = fmaxf(acc, 0.f);
        out[i] = __float2bfloat16_rn(acc);
    }
}

// ---- decode bookkeeping on the device (so the step can live in a CUDA graph): pos += 1, kv_len[b] += 1
__global__ void decode_advance_kernel(int* pos, int* kv_len, int B) {
    const int i = threadIdx.x;
    if (i == 0) *pos += 1;
    if (i < B) kv_len[i] += 1;
}

// ---- greedy argmax over fp32 logits [rows, V] -> int64 ids (first maximal index, like torch.argmax)
__global__ void argmax_kernel(const float* __restrict__ logits, long long* __restrict__ out, int V, long long ld) {
    __shared__ float sv[32];
    __shared__ int si[32];
    const float* row = logits + blockIdx.x * ld;
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = threadIdx.x; i < V; i += blockDim.x) {
        const float v = row[i];
        if (v > best || (v == best && i < bi)) { best = v; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { sv[w] = best; si[w] = bi; }
    __syncthreads();
    if (w == 0) {
        const int nw = blockDim.x >> 5;
        best = l < nw ? sv[l] : -INFINITY;
        bi = l < nw ? si[l] : 0x7fffffff;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if (l == 0) out[blockIdx.x] = bi;
    }
}

// ---- fp32 -> bf16 / bf16 -> fp32 casts
__global__ void f32_to_bf16_kernel(const float* __restrict__ a, __nv_bfloat16* __restrict__ b, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        b[i] = __float2bfloat16_rn(a[i]);
}
__global__ void bf16_to_f32_kernel(const __nv_bfloat16* __restrict__ a, float* __restrict__ b, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        b[i] = __bfloat162float(a[i]);
}

}  // namespace gb
using namespace gb;
#define ST reinterpret_cast<cudaStream_t>(stream)
#define BF(p) reinterpret_cast<__nv_bfloat16*>(p)
#define CBF(p) reinterpret_cast<const __nv_bfloat16*>(p)

GROMA_API int32_t groma_vit_patchify(const float* images, void* patches, int32_t B, int32_t S, int32_t ld, void* stream) {
    if (!images || !patches || B <= 0 || S % 14 || ld < 588 || (ld & 7)) return GROMA_ERR_ARG;
    const long long total = (long long)B * (S / 14) * (S / 14) * ld;
    patchify_kernel<<<grid_for(total, 256), 256, 0, ST>>>(images, BF(patches), B, S, ld);
    return GROMA_LAUNCH_CHECK();
}
GROMA_API int32_t groma_vit_embed(const void* patch, const float* cls, const float* pos, void* out, int32_t B,
                                  int32_t NP, int32_t C, void* stream) {
    if (!patch || !cls || !pos || !out) return GROMA_ERR_ARG;
    vit_embed_kernel<<<grid_for((long long)B * (NP + 1) * C, 256), 256, 0, ST>>>(CBF(patch), cls, pos, BF(out), B, NP, C);
    return GROMA_LAUNCH_CHECK();
}
GROMA_API int32_t groma_mean_tokens(const void* a0, const void* a1, const void* a2, const void* a3, int32_t n,
                                    void* out, int32_t B, int32_t T, int32_t C, int32_t skip, void* stream) {
    if (!a0 || !out || n < 1 || n > 4 || (C & 7)) return GROMA_ERR_ARG;
    mean_tokens_kernel<<<grid_for((long long)B * (T - skip) * (C / 8), 256), 256, 0, ST>>>(
        CBF(a0), CBF(a1), CBF(a2), CBF(a3), n, BF(out), B, T, C, skip);
    return GROMA_LAUNCH_CHECK();
}
GROMA_API int32_t groma_space_to_depth(const void* in, void* out, int32_t B, int32_t g, int32_t C, void* stream) {
    if (!in || !out || (g & 1) || (C & 7)) return GROMA_ERR_ARG;
    space_to_depth_kernel<<<grid_for((long long)B * g * g * (C / 8), 256), 256, 0, ST>>>(CBF(in), BF(out), B, g, C);
    return GROMA_LAUNCH_CHECK();
}
GROMA_API int32_t groma_gather_rows(const int64_t* idx, const void* t0, const void* t1, int64_t split, void* out,
                                    int64_t n, int32_t D, void* stream) {
    if (!idx || !t0 || !out || (D & 7)) return GROMA_ERR_ARG;
    if (n == 0) return GROMA_OK;
    gather_rows_kernel<<<grid_for(n * (D / 8), 256), 256, 0, ST>>>(reinterpret_cast<const long long*>(idx), CBF(t0),
                                                                    CBF(t1), split, BF(out), n, D);
    return GROMA_LAUNCH_CHECK();
}
GROMA_API int32_t groma_scatter_rows(const int64_t* idx, const void* src, void* out, int64_t n, int32_t D, void* stream) {
    if (!idx || !src || !out || (D & 7)) return GROMA_ERR_ARG;
    if (n == 0) return GROMA_OK;
    scatter_rows_kernel<<<grid_for(n * (D / 8), 256), 256, 0, ST>>>(reinterpret_cast<const long long*>(idx), CBF(src),
                                                                     BF(out), n, D);
    return GROMA_LAUNCH_CHECK();
}
GROMA_API int32_t groma_add(const void* a, const void* b, void* c, int64_t n, void* stream) {
    if (!a || !b || !c || (n & 7)) return GROMA_ERR_ARG;
    add_kernel<<<grid_for(n / 8, 256), 256, 0, ST>>>(CBF(a), CBF(b), BF(c), n / 8);
    return GROMA_LAUNCH_CHECK();
}
GROMA_API int32_t groma_add_bcast(const void* a, const void* b, void* c, int64_t rows, int64_t period, int32_t D,
                                  void* stream) {
    if (!a || !b || !c || (D & 7) || period < 1) return GROMA_ERR_ARG;
    add_bcast_kernel<<<grid_for(rows * (D / 8), 256), 256, 0, ST>>>(CBF(a), CBF(b), BF(c), rows, period, D);
    return GROMA_LAUNCH_CHECK();
}
GROMA_API int32_t groma_rope_kv(const void* qkv, void* q_out, void* cache_k, void* cache_v, const float* cos_t,
                                const float* sin_t, int32_t B, int32_t T, int32_t H, int32_t D, int32_t pos0,
                                const int32_t* pos_ptr, int64_t ctx_cap, void* stream) {
    if (!qkv || !q_out || !cache_k || !cache_v || !cos_t || !sin_t || (D & 15)) return GROMA_ERR_ARG;
    if (!pos_ptr && pos0 + T > ctx_cap) return GROMA_ERR_ARG;
    rope_kv_kernel<<<grid_for((long long)B * T * H * (D / 16), 256), 256, 0, ST>>>(
        CBF(qkv), BF(q_out), BF(cache_k), BF(cache_v), cos_t, sin_t, B, T, H, D, pos0, pos_ptr, ctx_cap);
    return GROMA_LAUNCH_CHECK();
}
GROMA_API int32_t groma_argmax(const float* logits, int64_t* out, int32_t rows, int32_t V, int64_t ld, void* stream) {
    if (!logits || !out || rows <= 0 || V <= 0) return GROMA_ERR_ARG;
    argmax_kernel<<<rows, 1024, 0, ST>>>(logits, reinterpret_cast<long long*>(out), V, ld);
    return GROMA_LAUNCH_CHECK();
}
GROMA_API int32_t groma_cast_f32_bf16(const float* a, void* b, int64_t n, void* stream) {
    if (!a || !b) return GROMA_ERR_ARG;
    f32_to_bf16_kernel<<<grid_for(n, 256), 256, 0, ST>>>(a, BF(b), n);
    return GROMA_LAUNCH_CHECK();
}
GROMA_API int32_t groma_cast_bf16_f32(const void* a, float* b, int64_t n, void* stream) {
    if (!a || !b) return GROMA_ERR_ARG;
    bf16_to_f32_kernel<<<grid_for(n, 256), 256, 0, ST>>>(CBF(a), b, n);
    return GROMA_LAUNCH_CHECK();
}

GROMA_API int32_t groma_linear_smallk(const float* x, const float* w, const float* b, void* out, int64_t M, int32_t N,
                                      int32_t K, int32_t relu, void* stream) {
    if (!x || !w || !out || K <= 0 || K > 64) return GROMA_ERR_ARG;
    if (M == 0) return GROMA_OK;
    linear_smallk_kernel<<<grid_for(M * N, 256), 256, 0, ST>>>(x, w, b, BF(out), M, N, K, relu);
    return GROMA_LAUNCH_CHECK();
}
GROMA_API int32_t groma_decode_advance(int32_t* pos, int32_t* kv_len, int32_t B, void* stream) {
    if (!pos || !kv_len || B <= 0 || B > 1024) return GROMA_ERR_ARG;
    decode_advance_kernel<<<1, ((B + 31) / 32) * 32, 0, ST>>>(pos, kv_len, B);
    return GROMA_LAUNCH_CHECK();
}
