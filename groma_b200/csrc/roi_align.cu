// RoIAlign (avg pool, `aligned` flag, fixed sampling ratio) over NHWC bf16 maps, fp32 arithmetic.
// Semantics = the reference's CUDA kernel (mmcv/ops/csrc/common/cuda/roi_align_cuda_kernel.cuh:17-108 with
// bilinear_interpolate of common_cuda_helper.hpp:28-70): no rounding of RoI coordinates, negative RoI extents are
// legal, samples outside [-1, H] x [-1, W] contribute 0.
// Layout is what the B200 path wants instead of the reference's NCHW: channels innermost, so each bilinear tap is
// one coalesced 16-byte-per-thread read, and the output can be written straight into the zero-bordered
// [K, PH+2, PW+2, C] buffer the implicit-GEMM 3x3 conv consumes (pad=1) -- no separate padding pass.
#include "ptx.cuh"
#include "capi_common.h"

namespace gb {

struct RoiParams {
    const __nv_bfloat16* input;  // [N, H, W, C]
    const float* rois;           // [K, 5] (batch_idx, x1, y1, x2, y2)
    __nv_bfloat16* output;       // [K, PH+2p, PW+2p, C]
    int K, C, H, W, PH, PW, sampling_ratio, aligned, pad;
    float spatial_scale;
};

struct Tap { int y_low, y_high, x_low, x_high; float w1, w2, w3, w4; bool valid; };

__device__ __forceinline__ Tap make_tap(float y, float x, int H, int W) {
    Tap t;
    t.valid = !(y < -1.0f || y > (float)H || x < -1.0f || x > (float)W);
    if (y <= 0.f) y = 0.f;
    if (x <= 0.f) x = 0.f;
    int y_low = (int)y, x_low = (int)x, y_high, x_high;
    if (y_low >= H - 1) { y_high = y_low = H - 1; y = (float)y_low; } else y_high = y_low + 1;
    if (x_low >= W - 1) { x_high = x_low = W - 1; x = (float)x_low; } else x_high = x_low + 1;
    const float ly = y - y_low, lx = x - x_low, hy = 1.f - ly, hx = 1.f - lx;
    t.y_low = y_low; t.y_high = y_high; t.x_low = x_low; t.x_high = x_high;
    t.w1 = hy * hx; t.w2 = hy * lx; t.w3 = ly * hx; t.w4 = ly * lx;
    return t;
}

// generic path (adaptive sampling grid, or maps too large for 32-bit element offsets): every thread recomputes every tap.
// grid (K, PH + 2*pad); block = C/8 threads (each owns 8 channels) -- loops over the PW bins of one output row.
__global__ void roi_align_nhwc_generic_kernel(const RoiParams p) {
    const int n = blockIdx.x;
    const int yy = blockIdx.y;  // padded row
    const int OW = p.PW + 2 * p.pad;
    const int nvec = p.C >> 3;
    __nv_bfloat16* orow = p.output + (((long long)n * (p.PH + 2 * p.pad) + yy) * OW) * p.C;
    const int ph = yy - p.pad;
    if (ph < 0 || ph >= p.PH) {
        for (int i = threadIdx.x; i < OW * nvec; i += blockDim.x) *reinterpret_cast<uint4*>(orow + (long long)i * 8) = make_uint4(0, 0, 0, 0);
        return;
    }
    const float* r = p.rois + n * 5;
    const int bi = (int)r[0];
    const float offset = p.aligned ? 0.5f : 0.0f;
    const float roi_start_w = r[1] * p.spatial_scale - offset;
    const float roi_start_h = r[2] * p.spatial_scale - offset;
    const float roi_end_w = r[3] * p.spatial_scale - offset;
    const float roi_end_h = r[4] * p.spatial_scale - offset;
    float roi_width = roi_end_w - roi_start_w, roi_height = roi_end_h - roi_start_h;
    if (!p.aligned) { roi_width = fmaxf(roi_width, 1.f); roi_height = fmaxf(roi_height, 1.f); }
    const float bin_h = roi_height / (float)p.PH, bin_w = roi_width / (float)p.PW;
    const int gh = p.sampling_ratio > 0 ? p.sampling_ratio : (int)ceilf(roi_height / p.PH);
    const int gw = p.sampling_ratio > 0 ? p.sampling_ratio : (int)ceilf(roi_width / p.PW);
    const float count = (float)max(gh * gw, 1);
    const __nv_bfloat16* in = p.input + (long long)bi * p.H * p.W * p.C;

    if (p.pad) {
        for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
            *reinterpret_cast<uint4*>(orow + (long long)i * 8) = make_uint4(0, 0, 0, 0);
            *reinterpret_cast<uint4*>(orow + ((long long)(OW - 1) * p.C) + i * 8) = make_uint4(0, 0, 0, 0);
        }
    }
    for (int pw = 0; pw < p.PW; ++pw) {
        for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
            float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (int iy = 0; iy < gh; ++iy) {
                const float y = roi_start_h + ph * bin_h + (iy + .5f) * bin_h / (float)gh;
                for (int ix = 0; ix < gw; ++ix) {
                    const float x = roi_start_w + pw * bin_w + (ix + .5f) * bin_w / (float)gw;
                    const Tap t = make_tap(y, x, p.H, p.W);
                    if (!t.valid) continue;
                    const uint4 a1 = *reinterpret_cast<const uint4*>(in + ((long long)t.y_low * p.W + t.x_low) * p.C + v * 8);
                    const uint4 a2 = *reinterpret_cast<const uint4*>(in + ((long long)t.y_low * p.W + t.x_high) * p.C + v * 8);
                    const uint4 a3 = *reinterpret_cast<const uint4*>(in + ((long long)t.y_high * p.W + t.x_low) * p.C + v * 8);
                    const uint4 a4 = *reinterpret_cast<const uint4*>(in + ((long long)t.y_high * p.W + t.x_high) * p.C + v * 8);
                    const __nv_bfloat162* b1 = reinterpret_cast<const __nv_bfloat162*>(&a1);
                    const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&a2);
                    const __nv_bfloat162* b3 = reinterpret_cast<const __nv_bfloat162*>(&a3);
                    const __nv_bfloat162* b4 = reinterpret_cast<const __nv_bfloat162*>(&a4);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float2 f1 = __bfloat1622float2(b1[q]), f2 = __bfloat1622float2(b2[q]);
                        const float2 f3 = __bfloat1622float2(b3[q]), f4 = __bfloat1622float2(b4[q]);
                        acc[2 * q] += t.w1 * f1.x + t.w2 * f2.x + t.w3 * f3.x + t.w4 * f4.x;
                        acc[2 * q + 1] += t.w1 * f1.y + t.w2 * f2.y + t.w3 * f3.y + t.w4 * f4.y;
                    }
                }
            }
            *reinterpret_cast<uint4*>(orow + ((long long)(pw + p.pad) * p.C) + v * 8) =
                make_uint4(pack_bf16x2(acc[0] / count, acc[1] / count), pack_bf16x2(acc[2] / count, acc[3] / count),
                           pack_bf16x2(acc[4] / count, acc[5] / count), pack_bf16x2(acc[6] / count, acc[7] / count));
        }
    }
}

// Fixed sampling grid (the path Groma takes: sampling_ratio = 2).  ncu on the first version showed the kernel bound by
// instruction issue, not by its gathers (profiles/r02_region_ops.md: 83 % issue-active, 24 % DRAM, 546 warp instructions per
// 16-byte output vector): every thread recomputed the tap indices / weights of every sample and spent two instructions per
// bf16 pair on unpacking plus five on the weighted sum.  Here the PW * gh * gw taps of the output row are computed ONCE per
// block into shared memory (valid taps packed first, in the reference's (iy, ix) order), the channel loop reads them back as
// two broadcast LDS.128 per tap, and the weighted sums run as packed fp32x2 FMAs (same rounding and the same order as the
// scalar form: mul, fma, fma, fma, add per element).
__global__ void roi_align_nhwc_kernel(const RoiParams p) {
    extern __shared__ __align__(16) unsigned char roi_smem[];
    const int n = blockIdx.x;
    const int yy = blockIdx.y;  // padded row
    const int OW = p.PW + 2 * p.pad;
    const int nvec = p.C >> 3;
    __nv_bfloat16* orow = p.output + (((long long)n * (p.PH + 2 * p.pad) + yy) * OW) * p.C;
    const int ph = yy - p.pad;
    if (ph < 0 || ph >= p.PH) {
        for (int i = threadIdx.x; i < OW * nvec; i += blockDim.x) *reinterpret_cast<uint4*>(orow + (long long)i * 8) = make_uint4(0, 0, 0, 0);
        return;
    }
    const int g = p.sampling_ratio, spb = g * g;           // samples per bin
    int4* tap_off = reinterpret_cast<int4*>(roi_smem);                       // [PW][spb] element offsets of the four corners
    float4* tap_w = reinterpret_cast<float4*>(tap_off + p.PW * spb);         // [PW][spb] bilinear weights
    int* tap_cnt = reinterpret_cast<int*>(tap_w + p.PW * spb);               // [PW] valid taps of the bin
    const float* r = p.rois + n * 5;
    const int bi = (int)r[0];
    if (threadIdx.x < p.PW) {
        const int pw = threadIdx.x;
        const float offset = p.aligned ? 0.5f : 0.0f;
        const float roi_start_w = r[1] * p.spatial_scale - offset;
        const float roi_start_h = r[2] * p.spatial_scale - offset;
        const float roi_end_w = r[3] * p.spatial_scale - offset;
        const float roi_end_h = r[4] * p.spatial_scale - offset;
        float roi_width = roi_end_w - roi_start_w, roi_height = roi_end_h - roi_start_h;
        if (!p.aligned) { roi_width = fmaxf(roi_width, 1.f); roi_height = fmaxf(roi_height, 1.f); }
        const float bin_h = roi_height / (float)p.PH, bin_w = roi_width / (float)p.PW;
        int cnt = 0;
        for (int iy = 0; iy < g; ++iy) {
            const float y = roi_start_h + ph * bin_h + (iy + .5f) * bin_h / (float)g;
            for (int ix = 0; ix < g; ++ix) {
                const float x = roi_start_w + pw * bin_w + (ix + .5f) * bin_w / (float)g;
                const Tap t = make_tap(y, x, p.H, p.W);
                if (!t.valid) continue;
                tap_off[pw * spb + cnt] = make_int4((t.y_low * p.W + t.x_low) * p.C, (t.y_low * p.W + t.x_high) * p.C,
                                                    (t.y_high * p.W + t.x_low) * p.C, (t.y_high * p.W + t.x_high) * p.C);
                tap_w[pw * spb + cnt] = make_float4(t.w1, t.w2, t.w3, t.w4);
                ++cnt;
            }
        }
        tap_cnt[pw] = cnt;
    }
    if (p.pad) {
        for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
            *reinterpret_cast<uint4*>(orow + (long long)i * 8) = make_uint4(0, 0, 0, 0);
            *reinterpret_cast<uint4*>(orow + ((long long)(OW - 1) * p.C) + i * 8) = make_uint4(0, 0, 0, 0);
        }
    }
    __syncthreads();
    const float count = (float)max(spb, 1);
    const bool pow2 = (spb & (spb - 1)) == 0;
    const float2 inv2 = make_float2(1.0f / count, 1.0f / count);
    const __nv_bfloat16* in = p.input + (long long)bi * p.H * p.W * p.C;
    for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
        const __nv_bfloat16* inv = in + v * 8;
#pragma unroll 2
        for (int pw = 0; pw < p.PW; ++pw) {
            float2 acc[4] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
            const int cnt = tap_cnt[pw];
            for (int s = 0; s < cnt; ++s) {
                const int4 o = tap_off[pw * spb + s];
                const float4 w = tap_w[pw * spb + s];
                const uint4 a1 = *reinterpret_cast<const uint4*>(inv + o.x);
                const uint4 a2 = *reinterpret_cast<const uint4*>(inv + o.y);
                const uint4 a3 = *reinterpret_cast<const uint4*>(inv + o.z);
                const uint4 a4 = *reinterpret_cast<const uint4*>(inv + o.w);
                const __nv_bfloat162* b1 = reinterpret_cast<const __nv_bfloat162*>(&a1);
                const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&a2);
                const __nv_bfloat162* b3 = reinterpret_cast<const __nv_bfloat162*>(&a3);
                const __nv_bfloat162* b4 = reinterpret_cast<const __nv_bfloat162*>(&a4);
                const float2 w1 = make_float2(w.x, w.x), w2 = make_float2(w.y, w.y), w3 = make_float2(w.z, w.z), w4 = make_float2(w.w, w.w);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float2 t = __fmul2_rn(w1, __bfloat1622float2(b1[q]));
                    t = __ffma2_rn(w2, __bfloat1622float2(b2[q]), t);
                    t = __ffma2_rn(w3, __bfloat1622float2(b3[q]), t);
                    t = __ffma2_rn(w4, __bfloat1622float2(b4[q]), t);
                    acc[q] = __fadd2_rn(acc[q], t);
                }
            }
            if (pow2) {   // x / 2^k == x * 2^-k bit for bit: skip the eight IEEE divisions
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[q] = __fmul2_rn(acc[q], inv2);
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) { acc[q].x = acc[q].x / count; acc[q].y = acc[q].y / count; }
            }
            *reinterpret_cast<uint4*>(orow + ((long long)(pw + p.pad) * p.C) + v * 8) =
                make_uint4(pack_bf16x2(acc[0].x, acc[0].y), pack_bf16x2(acc[1].x, acc[1].y),
                           pack_bf16x2(acc[2].x, acc[2].y), pack_bf16x2(acc[3].x, acc[3].y));
        }
    }
}

}  // namespace gb
using namespace gb;

// C-ABI twin of mmcv `roi_align_forward` (mmcv/ops/csrc/pytorch/pybind.cpp:191,611), avg-pool mode only:
// caller-allocated output, NHWC bf16 maps, optional 1-pixel zero border on the output.
GROMA_API int32_t groma_roi_align_forward(const void* input, const float* rois, void* output, int32_t K, int32_t C,
                                          int32_t H, int32_t W, int32_t pooled_h, int32_t pooled_w,
                                          float spatial_scale, int32_t sampling_ratio, int32_t aligned,
                                          int32_t out_pad, void* stream) {
    if (!input || !rois || !output || C <= 0 || H <= 0 || W <= 0 || pooled_h <= 0 || pooled_w <= 0) return GROMA_ERR_ARG;
    if (C & 7) return GROMA_ERR_ALIGN;
    if (K == 0) return GROMA_OK;
    RoiParams p;
    p.input = reinterpret_cast<const __nv_bfloat16*>(input); p.rois = rois;
    p.output = reinterpret_cast<__nv_bfloat16*>(output);
    p.K = K; p.C = C; p.H = H; p.W = W; p.PH = pooled_h; p.PW = pooled_w; p.sampling_ratio = sampling_ratio;
    p.aligned = aligned; p.pad = out_pad ? 1 : 0; p.spatial_scale = spatial_scale;
    int threads = C / 8;
    if (threads > 256) threads = 256;
    threads = ((threads + 31) / 32) * 32;
    dim3 grid(K, pooled_h + 2 * p.pad);
    const long long tap_bytes = (long long)pooled_w * sampling_ratio * sampling_ratio * 32 + (long long)pooled_w * 4;
    const bool fixed_grid = sampling_ratio > 0 && tap_bytes <= 40 * 1024 && pooled_w <= threads &&
                            (long long)H * W * C < (1LL << 31);
    if (fixed_grid)
        roi_align_nhwc_kernel<<<grid, threads, (size_t)tap_bytes, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    else
        roi_align_nhwc_generic_kernel<<<grid, threads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    return GROMA_LAUNCH_CHECK();
}
