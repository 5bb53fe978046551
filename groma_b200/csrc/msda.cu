// Multi-scale deformable attention sampling (Deformable-DETR).  Follows the SenseTime kernel the reference binds
// (mmcv/ops/csrc/common/cuda/ms_deform_attn_cuda_kernel.cuh:19-66,203-256) and its PyTorch twin
// (mmcv/ops/multi_scale_deform_attn.py:93-150): bilinear taps at loc*size-0.5 with zero padding, weighted by the
// per-head softmax over (levels x points).  Fused here: softmax of the attention logits and the
// reference-point + offset location math (modeling_deformable_detr.py:586-610), so the only HBM traffic is
// value (bf16), the fp32 projection row, the reference points and the bf16 output.
// One warp per (batch, query, head); lane = channel of the 32-wide head -> every tap is one 64-byte coalesced read.
#include "ptx.cuh"
#include "capi_common.h"

namespace gb {

struct MsdaParams {
    const __nv_bfloat16* value;  // [B, S, nH, 32]
    const float* proj;           // [B*Q, nH*L*P*2 + nH*L*P]  (sampling offsets | attention logits)
    const float* ref;            // [B, Q, ref_dim]  (ref_dim 2: per-level normalised xy shared by all levels; 4: cxcywh)
    __nv_bfloat16* out;          // [B, Q, nH*32]
    int B, Q, S, nH, L, P, ref_dim;
    int lvl_h[4], lvl_w[4], lvl_start[4];
};

__global__ void msda_kernel(const MsdaParams p) {
    const int lane = threadIdx.x & 31;
    const long long warp_global = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long total = (long long)p.B * p.Q * p.nH;
    if (warp_global >= total) return;
    const int head = warp_global % p.nH;
    const long long bq = warp_global / p.nH;
    const int b = bq / p.Q;
    const int LP = p.L * p.P;
    const int n_off = p.nH * LP * 2;
    const float* prow = p.proj + bq * (long long)(n_off + p.nH * LP);
    const float* off = prow + head * LP * 2;
    const float* logit = prow + n_off + head * LP;
    const float* ref = p.ref + bq * p.ref_dim;

    // softmax over L*P logits (<= 16), every lane computes it redundantly (broadcast loads)
    float mx = -INFINITY;
    for (int i = 0; i < LP; ++i) mx = fmaxf(mx, logit[i]);
    float den = 0.f;
    for (int i = 0; i < LP; ++i) den += __expf(logit[i] - mx);
    const float inv_den = 1.f / den;

    const __nv_bfloat16* vbase = p.value + ((long long)b * p.S) * p.nH * 32 + head * 32 + lane;
    const long long vstride = (long long)p.nH * 32;
    float acc = 0.f;
    for (int l = 0; l < p.L; ++l) {
        const int H = p.lvl_h[l], W = p.lvl_w[l];
        const __nv_bfloat16* vl = vbase + (long long)p.lvl_start[l] * vstride;
        for (int k = 0; k < p.P; ++k) {
            const int i = l * p.P + k;
            const float ox = off[i * 2], oy = off[i * 2 + 1];
            float lx, ly;
            if (p.ref_dim == 2) {
                lx = ref[0] + ox / (float)W;
                ly = ref[1] + oy / (float)H;
            } else {
                lx = ref[0] + ox / (float)p.P * ref[2] * 0.5f;
                ly = ref[1] + oy / (float)p.P * ref[3] * 0.5f;
            }
            const float w_attn = __expf(logit[i] - mx) * inv_den;
            const float h_im = ly * H - 0.5f, w_im = lx * W - 0.5f;
            if (h_im > -1.f && w_im > -1.f && h_im < (float)H && w_im < (float)W) {
                const int h_low = (int)floorf(h_im), w_low = (int)floorf(w_im);
                const int h_high = h_low + 1, w_high = w_low + 1;
                const float lh = h_im - h_low, lw = w_im - w_low, hh = 1.f - lh, hw = 1.f - lw;
                float v1 = 0.f, v2 = 0.f, v3 = 0.f, v4 = 0.f;
                if (h_low >= 0 && w_low >= 0) v1 = __bfloat162float(vl[(long long)(h_low * W + w_low) * vstride]);
                if (h_low >= 0 && w_high <= W - 1) v2 = __bfloat162float(vl[(long long)(h_low * W + w_high) * vstride]);
                if (h_high <= H - 1 && w_low >= 0) v3 = __bfloat162float(vl[(long long)(h_high * W + w_low) * vstride]);
                if (h_high <= H - 1 && w_high <= W - 1) v4 = __bfloat162float(vl[(long long)(h_high * W + w_high) * vstride]);
                acc += w_attn * (hh * hw * v1 + hh * lw * v2 + lh * hw * v3 + lh * lw * v4);
            }
        }
    }
    p.out[bq * (long long)(p.nH * 32) + head * 32 + lane] = __float2bfloat16_rn(acc);
}


// ---- single-level fast path: the value slice of one (image, head) -- S x 32 bf16 = 64 KB for the 32x32 map -- is staged
// in shared memory once (coalesced 16-byte loads) and every query of that image/head samples from it, so HBM/L2 sees the
// algorithmic bytes only (value once, projection row once, output once) instead of 16 scattered 64-byte gathers per
// (query, head).  grid (heads, B, q_splits); one warp per query, lane = channel.
__global__ void __launch_bounds__(512) msda_smem_kernel(const MsdaParams p, int q_per_cta) {
    extern __shared__ __align__(16) uint8_t sm_val[];
    __nv_bfloat16* sv = reinterpret_cast<__nv_bfloat16*>(sm_val);   // [S][32]
    const int head = blockIdx.x, b = blockIdx.y;
    const int H = p.lvl_h[0], W = p.lvl_w[0];
    // stage value[b, :, head, :]  (rows of 64 bytes at stride nH*64 bytes)
    const __nv_bfloat16* vsrc = p.value + ((long long)b * p.S) * p.nH * 32 + head * 32;
    for (int i = threadIdx.x; i < p.S * 4; i += blockDim.x) {
        const int r = i >> 2, c = i & 3;
        *reinterpret_cast<uint4*>(sv + r * 32 + c * 8) = *reinterpret_cast<const uint4*>(vsrc + (long long)r * p.nH * 32 + c * 8);
    }
    __syncthreads();
    // Two phases per chunk of 32 queries (P == 4 on this path):
    //  A  lane = query : 3 coalesced-per-lane float4 loads of its projection row (+ reference point), softmax of the 4
    //                    logits, location math, and the 16 (tap x corner) bilinear weights / clamped smem row indices
    //  B  half-warp = query, lane pair = 2 channels: weights and indices are broadcast by shuffles, values come from
    //                    shared memory as bf16x2 -- no global traffic at all in the sampling loop
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const int n_off = p.nH * 8;                       // P = 4: 8 offsets + 4 logits per head
    const int row_len = n_off + p.nH * 4;
    const int q0 = blockIdx.z * q_per_cta, q1 = min(p.Q, q0 + q_per_cta);
    const float fW = (float)W, fH = (float)H;
    for (int qc = q0 + warp * 32; qc < q1; qc += nw * 32) {
        const int q = qc + lane;
        float wgt[16];
        int idx[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) { wgt[i] = 0.f; idx[i] = 0; }
        if (q < q1) {
            const long long bq = (long long)b * p.Q + q;
            const float* prow = p.proj + bq * row_len;
            const float4 o0 = *reinterpret_cast<const float4*>(prow + head * 8);
            const float4 o1 = *reinterpret_cast<const float4*>(prow + head * 8 + 4);
            const float4 lg = *reinterpret_cast<const float4*>(prow + n_off + head * 4);
            const float* ref = p.ref + bq * p.ref_dim;
            const float r0 = ref[0], r1 = ref[1];
            float sx = 1.f / fW, sy = 1.f / fH;       // ref_dim 2: loc = ref + off / (W, H)
            const float ox[4] = {o0.x, o0.z, o1.x, o1.z}, oy[4] = {o0.y, o0.w, o1.y, o1.w};
            float lx[4], ly[4];
            if (p.ref_dim == 2) {
#pragma unroll
                for (int k = 0; k < 4; ++k) { lx[k] = r0 + ox[k] / fW; ly[k] = r1 + oy[k] / fH; }
            } else {
                const float r2 = ref[2], r3 = ref[3];
#pragma unroll
                for (int k = 0; k < 4; ++k) { lx[k] = r0 + ox[k] / 4.f * r2 * 0.5f; ly[k] = r1 + oy[k] / 4.f * r3 * 0.5f; }
            }
            (void)sx; (void)sy;
            const float l4[4] = {lg.x, lg.y, lg.z, lg.w};
            const float mx = fmaxf(fmaxf(l4[0], l4[1]), fmaxf(l4[2], l4[3]));
            float e4[4], den = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) { e4[k] = __expf(l4[k] - mx); den += e4[k]; }
            const float inv_den = 1.f / den;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float w_attn = e4[k] * inv_den;
                const float h_im = ly[k] * fH - 0.5f, w_im = lx[k] * fW - 0.5f;
                if (h_im > -1.f && w_im > -1.f && h_im < fH && w_im < fW) {
                    const int h_low = (int)floorf(h_im), w_low = (int)floorf(w_im);
                    const int h_high = h_low + 1, w_high = w_low + 1;
                    const float lh = h_im - h_low, lw = w_im - w_low, hh = 1.f - lh, hw = 1.f - lw;
                    const bool t_ok = h_low >= 0, b_ok = h_high <= H - 1, l_ok = w_low >= 0, r_ok = w_high <= W - 1;
                    const int hl = max(h_low, 0), hhh = min(h_high, H - 1), wl = max(w_low, 0), whh = min(w_high, W - 1);
                    wgt[k * 4 + 0] = (t_ok && l_ok) ? w_attn * hh * hw : 0.f; idx[k * 4 + 0] = hl * W + wl;
                    wgt[k * 4 + 1] = (t_ok && r_ok) ? w_attn * hh * lw : 0.f; idx[k * 4 + 1] = hl * W + whh;
                    wgt[k * 4 + 2] = (b_ok && l_ok) ? w_attn * lh * hw : 0.f; idx[k * 4 + 2] = hhh * W + wl;
                    wgt[k * 4 + 3] = (b_ok && r_ok) ? w_attn * lh * lw : 0.f; idx[k * 4 + 3] = hhh * W + whh;
                }
            }
        }
        const int half = lane >> 4, cl = (lane & 15) * 2;
#pragma unroll 1
        for (int pr = 0; pr < 16; ++pr) {
            const int src = 2 * pr + half;
            float ax = 0.f, ay = 0.f;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float wv = __shfl_sync(0xffffffffu, wgt[i], src);
                const int iv = __shfl_sync(0xffffffffu, idx[i], src);
                const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(sv + iv * 32 + cl));
                // same evaluation order as the reference kernel: sum over taps of w_attn * (4-corner bilinear)
                ax += wv * f.x;
                ay += wv * f.y;
            }
            const int qq = qc + src;
            if (qq < q1)
                *reinterpret_cast<uint32_t*>(p.out + ((long long)b * p.Q + qq) * (long long)(p.nH * 32) + head * 32 + cl) = pack_bf16x2(ax, ay);
        }
    }
}

}  // namespace gb
using namespace gb;

// C-ABI twin of mmcv `ms_deform_attn_forward` (mmcv/ops/csrc/pytorch/pybind.cpp:162,765) with the softmax and
// location arithmetic folded in; head_dim fixed at 32 (d_model 256 / 8 heads).
GROMA_API int32_t groma_msda_forward(const void* value, const float* proj, const float* ref, void* out, int32_t B,
                                     int32_t Q, int32_t S, int32_t n_heads, int32_t n_levels, int32_t n_points,
                                     int32_t ref_dim, const int32_t* level_hw, const int32_t* level_start, void* stream) {
    if (!value || !proj || !ref || !out || !level_hw || !level_start) return GROMA_ERR_ARG;
    if (n_levels < 1 || n_levels > 4 || n_points < 1 || n_levels * n_points > 16 || (ref_dim != 2 && ref_dim != 4))
        return GROMA_ERR_UNSUPPORTED;
    MsdaParams p;
    p.value = reinterpret_cast<const __nv_bfloat16*>(value); p.proj = proj; p.ref = ref;
    p.out = reinterpret_cast<__nv_bfloat16*>(out);
    p.B = B; p.Q = Q; p.S = S; p.nH = n_heads; p.L = n_levels; p.P = n_points; p.ref_dim = ref_dim;
    for (int l = 0; l < 4; ++l) {
        p.lvl_h[l] = l < n_levels ? level_hw[2 * l] : 0;
        p.lvl_w[l] = l < n_levels ? level_hw[2 * l + 1] : 0;
        p.lvl_start[l] = l < n_levels ? level_start[l] : 0;
    }
    if (n_levels == 1 && n_points == 4 && (long long)S * 64 <= 200 * 1024) {
        const int smem = S * 64;
        static int configured = 0;
        if (smem > 48 * 1024 && smem > configured) {
            if (cudaFuncSetAttribute(msda_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return GROMA_ERR_CUDA;
            configured = smem;
        }
        // split the queries of one (image, head) over enough CTAs to fill the 148 SMs about twice
        // 64 KB of shared memory per CTA -> 3 CTAs resident per SM: keep the whole grid in ONE wave
        int qs = (3 * 148) / (B * n_heads);
        if (qs < 1) qs = 1;
        if (qs > (Q + 63) / 64) qs = (Q + 63) / 64;
        const int q_per_cta = (Q + qs - 1) / qs;
        msda_smem_kernel<<<dim3(n_heads, B, qs), 512, smem, reinterpret_cast<cudaStream_t>(stream)>>>(p, q_per_cta);
        return GROMA_LAUNCH_CHECK();
    }
    const long long warps = (long long)B * Q * n_heads;
    const int threads = 256;
    const long long blocks = (warps * 32 + threads - 1) / threads;
    msda_kernel<<<(unsigned)blocks, threads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    return GROMA_LAUNCH_CHECK();
}
