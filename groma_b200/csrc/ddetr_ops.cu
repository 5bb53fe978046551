// Small fp32 kernels of the Deformable-DETR proposer glue (groma/model/ddetr_transformer.py:546-568,692-715 and
// groma/model/groma.py:246-249,268).  Kept in fp32 end to end because their outputs decide integer results
// (top-k indices, NMS keep set).
#include "ptx.cuh"
#include "capi_common.h"

namespace gb {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float inverse_sigmoid_(float x) {  // modeling_deformable_detr.py:981-985, eps 1e-5
    x = fminf(fmaxf(x, 0.f), 1.f);
    const float x1 = fmaxf(x, 1e-5f), x2 = fmaxf(1.f - x, 1e-5f);
    return logf(x1 / x2);
}

// Two-stage proposal selection: for the k top-scoring encoder tokens gather coord logits (= delta + grid proposal),
// emit reference points (sigmoid) and the 4x128 sine embedding of the proposal (get_proposal_pos_embed, :432-446).
//   delta [B,S,4] fp32, proposals [S,4] fp32 (inverse-sigmoid grid, +inf when invalid), topk [B,k] int64
//   ref_out [B,k,4] fp32, pos_out [B,k,4*npf] bf16
__global__ void ddetr_select_kernel(const float* __restrict__ delta, const float* __restrict__ proposals,
                                    const long long* __restrict__ topk, float* __restrict__ ref_out,
                                    __nv_bfloat16* __restrict__ pos_out, int B, int S, int k, int npf) {
    const long long total = (long long)B * k * 4 * npf;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int f = i % npf;
        long long r = i / npf;
        const int c = r % 4; r /= 4;
        const int q = r % k, b = r / k;
        const long long tok = topk[(long long)b * k + q];
        const float logit = delta[((long long)b * S + tok) * 4 + c] + proposals[tok * 4 + c];
        const float sg = sigmoidf_(logit);
        if (f == 0) ref_out[((long long)b * k + q) * 4 + c] = sg;
        const float dim_t = powf(10000.f, 2.f * (float)(f / 2) / (float)npf);
        const float v = sg * 6.283185307179586f / dim_t;
        pos_out[i] = __float2bfloat16_rn((f & 1) ? cosf(v) : sinf(v));
    }
}

// Final heads: pred = sigmoid(d5 + inv_sigmoid(sigmoid(d4 + inv_sigmoid(ref0))))   
//              score = sigmoid(coco)^0.4 * sigmoid(sa1b)^0.6 ; xyxy = center_to_corners(pred)
__global__ void ddetr_finalize_kernel(const float* __restrict__ d4, const float* __restrict__ d5,
                                      const float* __restrict__ ref0, const float* __restrict__ coco,
                                      const float* __restrict__ sa1b, float* __restrict__ pred_cxcywh,
                                      float* __restrict__ pred_xyxy, float* __restrict__ score, long long n,
                                      long long out_stride_boxes, long long out_stride_scores, int per_img) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float bx[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float r1 = sigmoidf_(d4[i * 4 + c] + inverse_sigmoid_(ref0[i * 4 + c]));
            bx[c] = sigmoidf_(d5[i * 4 + c] + inverse_sigmoid_(r1));
        }
        const long long img = i / per_img, q = i % per_img;
        float* pc = pred_cxcywh + (img * out_stride_boxes + q) * 4;
        float* px = pred_xyxy + (img * out_stride_boxes + q) * 4;
        pc[0] = bx[0]; pc[1] = bx[1]; pc[2] = bx[2]; pc[3] = bx[3];
        px[0] = bx[0] - 0.5f * bx[2]; px[1] = bx[1] - 0.5f * bx[3];
        px[2] = bx[0] + 0.5f * bx[2]; px[3] = bx[1] + 0.5f * bx[3];
        score[img * out_stride_scores + q] = powf(sigmoidf_(coco[i]), 0.4f) * powf(sigmoidf_(sa1b[i]), 0.6f);
    }
}

// rows of x [B,S,D] whose valid[s]==0 are zeroed (gen_encoder_output_proposals masking, :424-426)
__global__ void mask_rows_kernel(__nv_bfloat16* __restrict__ x, const unsigned char* __restrict__ valid, int B, int S, int D) {
    const long long total = (long long)B * S * D;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int s = (i / D) % S;
        if (!valid[s]) x[i] = __float2bfloat16_rn(0.f);
    }
}

}  // namespace gb
using namespace gb;

static inline int grid_n(long long n) { long long b = (n + 255) / 256; if (b > 148 * 16) b = 148 * 16; if (b < 1) b = 1; return (int)b; }

GROMA_API int32_t groma_ddetr_select(const float* delta, const float* proposals, const int64_t* topk, float* ref_out,
                                     void* pos_out, int32_t B, int32_t S, int32_t k, int32_t num_pos_feats, void* stream) {
    if (!delta || !proposals || !topk || !ref_out || !pos_out) return GROMA_ERR_ARG;
    ddetr_select_kernel<<<grid_n((long long)B * k * 4 * num_pos_feats), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        delta, proposals, reinterpret_cast<const long long*>(topk), ref_out, reinterpret_cast<__nv_bfloat16*>(pos_out), B, S, k,
        num_pos_feats);
    return GROMA_LAUNCH_CHECK();
}
GROMA_API int32_t groma_ddetr_finalize(const float* d4, const float* d5, const float* ref0, const float* coco,
                                       const float* sa1b, float* pred_cxcywh, float* pred_xyxy, float* score,
                                       int32_t B, int32_t Q, int64_t out_stride_boxes, int64_t out_stride_scores,
                                       void* stream) {
    if (!d4 || !d5 || !ref0 || !coco || !sa1b || !pred_cxcywh || !pred_xyxy || !score) return GROMA_ERR_ARG;
    ddetr_finalize_kernel<<<grid_n((long long)B * Q), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        d4, d5, ref0, coco, sa1b, pred_cxcywh, pred_xyxy, score, (long long)B * Q, out_stride_boxes, out_stride_scores, Q);
    return GROMA_LAUNCH_CHECK();
}
GROMA_API int32_t groma_mask_rows(void* x, const uint8_t* valid, int32_t B, int32_t S, int32_t D, void* stream) {
    if (!x || !valid) return GROMA_ERR_ARG;
    mask_rows_kernel<<<grid_n((long long)B * S * D), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<__nv_bfloat16*>(x), valid, B, S, D);
    return GROMA_LAUNCH_CHECK();
}
