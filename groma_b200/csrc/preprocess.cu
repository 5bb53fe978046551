// Image preprocessing in front of the ViT: the reference does this on the CPU per image --
//   Image.open(f).convert('RGB').resize((448, 448))                        groma/eval/run_groma.py:77-78 (Pillow, BICUBIC, uint8)
//   vis_processor.preprocess(image, return_tensors='pt')['pixel_values']   groma/eval/run_groma.py:79 (rescale 1/255, mean/std)
// Here: raw uint8 HWC bytes go to the device once and three small kernels reproduce Pillow's two-pass fixed-point resampler
// bit for bit (src/libImaging/Resample.c: coefficients in double -> 22-bit fixed point, uint8 between the passes) and apply
// the rescale+normalize as a 3x256 float32 table.  Byte/integer work, HBM/L2 bound, no tensor cores.
#include "capi_common.h"
#include <cstdint>
#include <cuda_runtime.h>

namespace gb {

constexpr int PP_PRECISION_BITS = 32 - 8 - 2;
constexpr int PP_KMAX = 64;          // taps per output sample: ceil(2*scale)*2+1 <= 64  ->  input up to 15.5x the output size

// Pillow's bicubic_filter (a = -0.5).  Explicit round-to-nearest intrinsics: no FMA contraction, same doubles as the C code.
__device__ __forceinline__ double pp_bicubic(double x) {
    if (x < 0.0) x = -x;
    if (x < 1.0) {
        double t = __dmul_rn(1.5, x);
        t = __dsub_rn(t, 2.5);
        t = __dmul_rn(t, x);
        t = __dmul_rn(t, x);
        return __dadd_rn(t, 1.0);
    }
    if (x < 2.0) {
        double t = __dsub_rn(x, 5.0);
        t = __dmul_rn(t, x);
        t = __dadd_rn(t, 8.0);
        t = __dmul_rn(t, x);
        t = __dsub_rn(t, 4.0);
        return __dmul_rn(t, -0.5);
    }
    return 0.0;
}

// precompute_coeffs + normalize_coeffs_8bpc for both axes.  blockIdx.x: 0 = horizontal (in = W), 1 = vertical (in = H).
// coef layout per axis: int32 bounds[out][2] (first input index, taps) followed by int32 kk[out][PP_KMAX].
__global__ void pp_coeffs_kernel(int in_w, int in_h, int out_size, int* __restrict__ coef) {
    const int axis = blockIdx.x;
    const int in_size = axis == 0 ? in_w : in_h;
    int* bounds = coef + (size_t)axis * out_size * (2 + PP_KMAX);
    int* kk = bounds + 2 * out_size;
    const double scale = __ddiv_rn((double)in_size, (double)out_size);
    const double filterscale = scale < 1.0 ? 1.0 : scale;
    const double support = __dmul_rn(2.0, filterscale);
    const double ss = __ddiv_rn(1.0, filterscale);
    for (int xx = threadIdx.x; xx < out_size; xx += blockDim.x) {
        const double center = __dmul_rn(__dadd_rn((double)xx, 0.5), scale);
        int xmin = __double2int_rz(__dadd_rn(__dsub_rn(center, support), 0.5));
        if (xmin < 0) xmin = 0;
        int xmax = __double2int_rz(__dadd_rn(__dadd_rn(center, support), 0.5));
        if (xmax > in_size) xmax = in_size;
        xmax -= xmin;
        if (xmax > PP_KMAX) xmax = PP_KMAX;   // unreachable: the launcher rejects scales that need more taps
        double k[PP_KMAX];
        double ww = 0.0;
        for (int x = 0; x < xmax; ++x) {
            const double w = pp_bicubic(__dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ss));
            k[x] = w;
            ww = __dadd_rn(ww, w);
        }
        for (int x = 0; x < PP_KMAX; ++x) {
            int v = 0;
            if (x < xmax) {
                const double kn = (ww != 0.0) ? __ddiv_rn(k[x], ww) : k[x];
                const double f = __dmul_rn(kn, (double)(1 << PP_PRECISION_BITS));
                v = kn < 0 ? __double2int_rz(__dadd_rn(-0.5, f)) : __double2int_rz(__dadd_rn(0.5, f));
            }
            kk[(size_t)xx * PP_KMAX + x] = v;
        }
        bounds[2 * xx] = xmin;
        bounds[2 * xx + 1] = xmax;
    }
}

__device__ __forceinline__ uint8_t pp_clip8(int acc) {
    const int v = acc >> PP_PRECISION_BITS;   // arithmetic shift, as Pillow's clip8 lookup index
    return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// Horizontal pass: one CTA per input row, the row staged in shared memory.  tmp[row][xx][c].
__global__ void __launch_bounds__(256) pp_horizontal_kernel(const uint8_t* __restrict__ img, int W, long long row_stride, int out_size,
                                                            const int* __restrict__ coef, uint8_t* __restrict__ tmp) {
    extern __shared__ uint8_t row[];
    const int r = blockIdx.x;
    const uint8_t* src = img + (long long)r * row_stride;
    for (int i = threadIdx.x; i < W * 3; i += blockDim.x) row[i] = src[i];
    __syncthreads();
    const int* bounds = coef;
    const int* kk = coef + 2 * out_size;
    for (int xx = threadIdx.x; xx < out_size; xx += blockDim.x) {
        const int x0 = bounds[2 * xx], n = bounds[2 * xx + 1];
        const int* k = kk + (size_t)xx * PP_KMAX;
        int a0 = 1 << (PP_PRECISION_BITS - 1), a1 = a0, a2 = a0;
        for (int t = 0; t < n; ++t) {
            const int kv = k[t];
            const uint8_t* p = row + (x0 + t) * 3;
            a0 += p[0] * kv; a1 += p[1] * kv; a2 += p[2] * kv;
        }
        uint8_t* o = tmp + ((long long)r * out_size + xx) * 3;
        o[0] = pp_clip8(a0); o[1] = pp_clip8(a1); o[2] = pp_clip8(a2);
    }
}

// Vertical pass + rescale/normalize table: one CTA per output row.  out_f32 [3][S][S] (pixel_values), out_u8 [S][S][3].
__global__ void __launch_bounds__(256) pp_vertical_kernel(const uint8_t* __restrict__ tmp, int out_size, const int* __restrict__ coef,
                                                          const float* __restrict__ lut, float* __restrict__ out_f32,
                                                          uint8_t* __restrict__ out_u8) {
    const int yy = blockIdx.x;
    const int* bounds = coef + (size_t)out_size * (2 + PP_KMAX);
    const int* kk = bounds + 2 * out_size;
    const int y0 = bounds[2 * yy], n = bounds[2 * yy + 1];
    __shared__ int ks[PP_KMAX];
    if (threadIdx.x < PP_KMAX) ks[threadIdx.x] = kk[(size_t)yy * PP_KMAX + threadIdx.x];
    __syncthreads();
    for (int xx = threadIdx.x; xx < out_size; xx += blockDim.x) {
        int a0 = 1 << (PP_PRECISION_BITS - 1), a1 = a0, a2 = a0;
        for (int t = 0; t < n; ++t) {
            const uint8_t* p = tmp + ((long long)(y0 + t) * out_size + xx) * 3;
            const int kv = ks[t];
            a0 += p[0] * kv; a1 += p[1] * kv; a2 += p[2] * kv;
        }
        const uint8_t v0 = pp_clip8(a0), v1 = pp_clip8(a1), v2 = pp_clip8(a2);
        if (out_u8) {
            uint8_t* o = out_u8 + ((long long)yy * out_size + xx) * 3;
            o[0] = v0; o[1] = v1; o[2] = v2;
        }
        if (out_f32) {
            const long long plane = (long long)out_size * out_size, off = (long long)yy * out_size + xx;
            out_f32[off] = lut[v0];
            out_f32[plane + off] = lut[256 + v1];
            out_f32[2 * plane + off] = lut[512 + v2];
        }
    }
}

}  // namespace gb

GROMA_API int32_t groma_preprocess_image(const uint8_t* img, int32_t H, int32_t W, int64_t row_stride, const float* lut,
                                         int32_t out_size, uint8_t* tmp, int32_t* coef, float* out_f32, uint8_t* out_u8,
                                         void* stream) {
    if (!img || !tmp || !coef || (!out_f32 && !out_u8) || (out_f32 && !lut) || H <= 0 || W <= 0 || out_size <= 0) return GROMA_ERR_ARG;
    if (row_stride < (int64_t)W * 3) return GROMA_ERR_ARG;
    // taps per sample = ceil(2*scale)*2+1 must fit PP_KMAX; 15x covers 6720-pixel sides at 448
    if ((int64_t)H > 15LL * out_size || (int64_t)W > 15LL * out_size || W * 3 > 48 * 1024) return GROMA_ERR_UNSUPPORTED;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int threads = 256;
    gb::pp_coeffs_kernel<<<2, threads, 0, st>>>(W, H, out_size, coef);
    const size_t smem = (size_t)W * 3;
    gb::pp_horizontal_kernel<<<H, threads, smem, st>>>(img, W, row_stride, out_size, coef, tmp);
    gb::pp_vertical_kernel<<<out_size, threads, 0, st>>>(tmp, out_size, coef, lut, out_f32, out_u8);
    return cudaGetLastError() == cudaSuccess ? GROMA_OK : GROMA_ERR_CUDA;
}
