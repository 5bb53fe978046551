// Bilinear (align_corners=True) resampling kernels of the region encoder (groma/model/roi_align.py:118-126,150-178,
// 215-228), NHWC bf16, fp32 interpolation, 16-byte vectors.  Formula and index clamping follow
// torch.nn.functional.interpolate(mode='bilinear', align_corners=True).
#include "ptx.cuh"
#include "gn_common.cuh"
#include "capi_common.h"

namespace gb {

struct Lerp { int i0, i1; float l0, l1; };
__device__ __forceinline__ Lerp make_lerp(int dst, int in_size, int out_size) {
    Lerp r;
    const float scale = out_size > 1 ? (float)(in_size - 1) / (float)(out_size - 1) : 0.f;
    const float src = scale * dst;
    r.i0 = (int)src;
    if (r.i0 > in_size - 1) r.i0 = in_size - 1;
    r.i1 = r.i0 + ((r.i0 < in_size - 1) ? 1 : 0);
    r.l1 = src - r.i0;
    r.l0 = 1.f - r.l1;
    return r;
}

__device__ __forceinline__ uint4 bilerp8(const __nv_bfloat16* p00, const __nv_bfloat16* p01, const __nv_bfloat16* p10,
                                         const __nv_bfloat16* p11, const Lerp& ly, const Lerp& lx) {
    const uint4 a = *reinterpret_cast<const uint4*>(p00), b = *reinterpret_cast<const uint4*>(p01);
    const uint4 c = *reinterpret_cast<const uint4*>(p10), d = *reinterpret_cast<const uint4*>(p11);
    const __nv_bfloat162* a2 = reinterpret_cast<const __nv_bfloat162*>(&a);
    const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&b);
    const __nv_bfloat162* c2 = reinterpret_cast<const __nv_bfloat162*>(&c);
    const __nv_bfloat162* d2 = reinterpret_cast<const __nv_bfloat162*>(&d);
    uint32_t o[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const float2 fa = __bfloat1622float2(a2[t]), fb = __bfloat1622float2(b2[t]);
        const float2 fc = __bfloat1622float2(c2[t]), fd = __bfloat1622float2(d2[t]);
        const float x = ly.l0 * (lx.l0 * fa.x + lx.l1 * fb.x) + ly.l1 * (lx.l0 * fc.x + lx.l1 * fd.x);
        const float y = ly.l0 * (lx.l0 * fa.y + lx.l1 * fb.y) + ly.l1 * (lx.l0 * fc.y + lx.l1 * fd.y);
        o[t] = pack_bf16x2(x, y);
    }
    return make_uint4(o[0], o[1], o[2], o[3]);
}

// tokens [B, skip + g*g, C] -> out [B, Ho, Wo, ld]: channels [0,C) bilinear, C = x coord, C+1 = y coord, rest 0
__global__ void upsample_coords_kernel(const __nv_bfloat16* __restrict__ tok, int skip, int g, int C,
                                       __nv_bfloat16* __restrict__ out, int B, int Ho, int Wo, int ld,
                                       const float* __restrict__ xs, const float* __restrict__ ys) {
    const int nvec = ld >> 3, cvec = C >> 3;
    const long long total = (long long)B * Ho * Wo * nvec;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int v = i % nvec;
        long long r = i / nvec;
        const int x = r % Wo; r /= Wo;
        const int y = r % Ho;
        const int b = r / Ho;
        uint4 o;
        if (v < cvec) {
            const Lerp ly = make_lerp(y, g, Ho), lx = make_lerp(x, g, Wo);
            const __nv_bfloat16* base = tok + ((long long)b * (skip + g * g) + skip) * C + v * 8;
            o = bilerp8(base + (long long)(ly.i0 * g + lx.i0) * C, base + (long long)(ly.i0 * g + lx.i1) * C,
                        base + (long long)(ly.i1 * g + lx.i0) * C, base + (long long)(ly.i1 * g + lx.i1) * C, ly, lx);
        } else if (v == cvec) {
            o = make_uint4(pack_bf16x2(xs[x], ys[y]), 0, 0, 0);
        } else {
            o = make_uint4(0, 0, 0, 0);
        }
        *reinterpret_cast<uint4*>(out + i * 8) = o;
    }
}

// One target level of MLVLFuseModule._single_shuffle (roi_align.py:150-178):
//   out[b, 1+y, 1+x, 0:C/2)      = tar[b,y,x, 0:C/2)
//   out[..., C/2:3C/4)           = resize(top[..., 3C/4:C))   to (Ht,Wt)
//   out[..., 3C/4:C)             = resize(down[..., C/2:3C/4)) to (Ht,Wt)
// out is the zero-bordered [B, Ht+2, Wt+2, C] buffer the 3x3 implicit-GEMM conv reads; borders are (re)written as 0.
__global__ void fuse_shuffle_kernel(const __nv_bfloat16* __restrict__ tar, const __nv_bfloat16* __restrict__ top,
                                    const __nv_bfloat16* __restrict__ down, __nv_bfloat16* __restrict__ out, int B,
                                    int C, int Ht, int Wt, int Htop, int Wtop, int Hdn, int Wdn) {
    const int nvec = C >> 3;
    const int Hp = Ht + 2, Wp = Wt + 2;
    const long long total = (long long)B * Hp * Wp * nvec;
    const int q = nvec / 4;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int v = i % nvec;
        long long r = i / nvec;
        const int xp = r % Wp; r /= Wp;
        const int yp = r % Hp;
        const int b = r / Hp;
        uint4 o = make_uint4(0, 0, 0, 0);
        if (xp >= 1 && xp <= Wt && yp >= 1 && yp <= Ht) {
            const int x = xp - 1, y = yp - 1;
            if (v < 2 * q) {
                o = *reinterpret_cast<const uint4*>(tar + (((long long)b * Ht + y) * Wt + x) * C + v * 8);
            } else if (v < 3 * q) {
                const int sv = v + q;  // channels [3C/4, C) of top
                const Lerp ly = make_lerp(y, Htop, Ht), lx = make_lerp(x, Wtop, Wt);
                const __nv_bfloat16* base = top + (long long)b * Htop * Wtop * C + sv * 8;
                o = bilerp8(base + (long long)(ly.i0 * Wtop + lx.i0) * C, base + (long long)(ly.i0 * Wtop + lx.i1) * C,
                            base + (long long)(ly.i1 * Wtop + lx.i0) * C, base + (long long)(ly.i1 * Wtop + lx.i1) * C, ly, lx);
            } else {
                const int sv = v - q;  // channels [C/2, 3C/4) of down
                const Lerp ly = make_lerp(y, Hdn, Ht), lx = make_lerp(x, Wdn, Wt);
                const __nv_bfloat16* base = down + (long long)b * Hdn * Wdn * C + sv * 8;
                o = bilerp8(base + (long long)(ly.i0 * Wdn + lx.i0) * C, base + (long long)(ly.i0 * Wdn + lx.i1) * C,
                            base + (long long)(ly.i1 * Wdn + lx.i0) * C, base + (long long)(ly.i1 * Wdn + lx.i1) * C, ly, lx);
            }
        }
        *reinterpret_cast<uint4*>(out + i * 8) = o;
    }
}

// bilerp8 over four taps already in registers
__device__ __forceinline__ uint4 bilerp8v(const uint4& a, const uint4& b, const uint4& c, const uint4& d, const Lerp& ly, const Lerp& lx) {
    const __nv_bfloat162* a2 = reinterpret_cast<const __nv_bfloat162*>(&a);
    const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&b);
    const __nv_bfloat162* c2 = reinterpret_cast<const __nv_bfloat162*>(&c);
    const __nv_bfloat162* d2 = reinterpret_cast<const __nv_bfloat162*>(&d);
    uint32_t o[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const float2 fa = __bfloat1622float2(a2[t]), fb = __bfloat1622float2(b2[t]);
        const float2 fc = __bfloat1622float2(c2[t]), fd = __bfloat1622float2(d2[t]);
        const float x = ly.l0 * (lx.l0 * fa.x + lx.l1 * fb.x) + ly.l1 * (lx.l0 * fc.x + lx.l1 * fd.x);
        const float y = ly.l0 * (lx.l0 * fa.y + lx.l1 * fb.y) + ly.l1 * (lx.l0 * fc.y + lx.l1 * fd.y);
        o[t] = pack_bf16x2(x, y);
    }
    return make_uint4(o[0], o[1], o[2], o[3]);
}

// fuse_shuffle_kernel over the RAW conv outputs of the previous fusion round: relu(GroupNorm) of that round (per-level statistics
// st_* = [B, G, (mean, rstd)], one shared gamma / beta: the three levels go through the same ConvModule, roi_align.py:133-143)
// is applied to every tap on its way in and rounded to bf16 exactly as gn_relu_apply_kernel would have stored it, so the result
// equals apply -> store -> fuse_shuffle bit for bit while the apply pass (one read + one write of every map) disappears.
__global__ void fuse_shuffle_gn_kernel(const __nv_bfloat16* __restrict__ tar, const __nv_bfloat16* __restrict__ top,
                                       const __nv_bfloat16* __restrict__ down, __nv_bfloat16* __restrict__ out, int B,
                                       int C, int Ht, int Wt, int Htop, int Wtop, int Hdn, int Wdn,
                                       const float* __restrict__ st_tar, const float* __restrict__ st_top,
                                       const float* __restrict__ st_dn, const float* __restrict__ gamma,
                                       const float* __restrict__ beta, int G) {
    const int nvec = C >> 3;
    const int Hp = Ht + 2, Wp = Wt + 2;
    const long long total = (long long)B * Hp * Wp * nvec;
    const int q = nvec / 4;
    const int cpg = C / G;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int v = i % nvec;
        long long r = i / nvec;
        const int xp = r % Wp; r /= Wp;
        const int yp = r % Hp;
        const int b = r / Hp;
        uint4 o = make_uint4(0, 0, 0, 0);
        if (xp >= 1 && xp <= Wt && yp >= 1 && yp <= Ht) {
            const int x = xp - 1, y = yp - 1;
            if (v < 2 * q) {
                const int g = (v * 8) / cpg;
                const float mean = st_tar[(b * G + g) * 2], rstd = st_tar[(b * G + g) * 2 + 1];
                o = gn_relu8(*reinterpret_cast<const uint4*>(tar + (((long long)b * Ht + y) * Wt + x) * C + v * 8), mean, rstd,
                             gn_load_affine8(gamma, beta, v * 8));
            } else {
                const bool from_top = v < 3 * q;
                const int sv = from_top ? v + q : v - q;   // channels [3C/4, C) of top | [C/2, 3C/4) of down
                const int Hs = from_top ? Htop : Hdn, Ws = from_top ? Wtop : Wdn;
                const float* stt = from_top ? st_top : st_dn;
                const int g = (sv * 8) / cpg;
                const float mean = stt[(b * G + g) * 2], rstd = stt[(b * G + g) * 2 + 1];
                const GnAffine8 af = gn_load_affine8(gamma, beta, sv * 8);
                const Lerp ly = make_lerp(y, Hs, Ht), lx = make_lerp(x, Ws, Wt);
                const __nv_bfloat16* base = (from_top ? top : down) + (long long)b * Hs * Ws * C + sv * 8;
                const uint4 t00 = *reinterpret_cast<const uint4*>(base + (long long)(ly.i0 * Ws + lx.i0) * C);
                const uint4 t01 = *reinterpret_cast<const uint4*>(base + (long long)(ly.i0 * Ws + lx.i1) * C);
                const uint4 t10 = *reinterpret_cast<const uint4*>(base + (long long)(ly.i1 * Ws + lx.i0) * C);
                const uint4 t11 = *reinterpret_cast<const uint4*>(base + (long long)(ly.i1 * Ws + lx.i1) * C);
                o = bilerp8v(gn_relu8(t00, mean, rstd, af), gn_relu8(t01, mean, rstd, af), gn_relu8(t10, mean, rstd, af),
                             gn_relu8(t11, mean, rstd, af), ly, lx);
            }
        }
        *reinterpret_cast<uint4*>(out + i * 8) = o;
    }
}

}  // namespace gb
using namespace gb;
static inline int grid_rs(long long n) { long long b = (n + 255) / 256; if (b > 148 * 32) b = 148 * 32; if (b < 1) b = 1; return (int)b; }

GROMA_API int32_t groma_upsample_coords(const void* tokens, int32_t skip, int32_t g, int32_t C, void* out, int32_t B,
                                        int32_t Ho, int32_t Wo, int32_t ld, const float* xs, const float* ys, void* stream) {
    if (!tokens || !out || !xs || !ys || (C & 7) || (ld & 7) || ld < C + 8) return GROMA_ERR_ARG;
    upsample_coords_kernel<<<grid_rs((long long)B * Ho * Wo * (ld / 8)), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const __nv_bfloat16*>(tokens), skip, g, C, reinterpret_cast<__nv_bfloat16*>(out), B, Ho, Wo, ld, xs, ys);
    return GROMA_LAUNCH_CHECK();
}
GROMA_API int32_t groma_fuse_shuffle(const void* tar, const void* top, const void* down, void* out, int32_t B, int32_t C,
                                     int32_t Ht, int32_t Wt, int32_t Htop, int32_t Wtop, int32_t Hdn, int32_t Wdn, void* stream) {
    if (!tar || !top || !down || !out || (C % 32)) return GROMA_ERR_ARG;
    fuse_shuffle_kernel<<<grid_rs((long long)B * (Ht + 2) * (Wt + 2) * (C / 8)), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const __nv_bfloat16*>(tar), reinterpret_cast<const __nv_bfloat16*>(top),
        reinterpret_cast<const __nv_bfloat16*>(down), reinterpret_cast<__nv_bfloat16*>(out), B, C, Ht, Wt, Htop, Wtop, Hdn, Wdn);
    return GROMA_LAUNCH_CHECK();
}
GROMA_API int32_t groma_fuse_shuffle_gn(const void* tar, const void* top, const void* down, void* out, int32_t B, int32_t C,
                                        int32_t Ht, int32_t Wt, int32_t Htop, int32_t Wtop, int32_t Hdn, int32_t Wdn,
                                        const float* stats_tar, const float* stats_top, const float* stats_down,
                                        const float* gamma, const float* beta, int32_t G, void* stream) {
    if (!tar || !top || !down || !out || !stats_tar || !stats_top || !stats_down || !gamma || !beta || (C % 32) || G <= 0) return GROMA_ERR_ARG;
    if ((C % G) || ((C / G) & 7)) return GROMA_ERR_ALIGN;
    fuse_shuffle_gn_kernel<<<grid_rs((long long)B * (Ht + 2) * (Wt + 2) * (C / 8)), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const __nv_bfloat16*>(tar), reinterpret_cast<const __nv_bfloat16*>(top),
        reinterpret_cast<const __nv_bfloat16*>(down), reinterpret_cast<__nv_bfloat16*>(out), B, C, Ht, Wt, Htop, Wtop, Hdn, Wdn,
        stats_tar, stats_top, stats_down, gamma, beta, G);
    return GROMA_LAUNCH_CHECK();
}
