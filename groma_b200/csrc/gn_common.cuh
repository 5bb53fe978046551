// relu(GroupNorm) of one 16-byte vector of 8 bf16 channels, shared by the stand-alone apply kernel (norm.cu) and by the
// resampling kernel that applies it tap by tap (resample.cu): explicit intrinsics, so both contract the same way and the fused
// path stays bit-identical to "apply, store bf16, resample" (mmcv ConvModule norm + act, mmcv/cnn/bricks/conv_module.py:196-206).
#pragma once
#include "ptx.cuh"

namespace gb {

struct GnAffine8 { float g[8], b[8]; };
__device__ __forceinline__ GnAffine8 gn_load_affine8(const float* __restrict__ gamma, const float* __restrict__ beta, int c0) {
    GnAffine8 a;
    *reinterpret_cast<float4*>(a.g) = *reinterpret_cast<const float4*>(gamma + c0);
    *reinterpret_cast<float4*>(a.g + 4) = *reinterpret_cast<const float4*>(gamma + c0 + 4);
    *reinterpret_cast<float4*>(a.b) = *reinterpret_cast<const float4*>(beta + c0);
    *reinterpret_cast<float4*>(a.b + 4) = *reinterpret_cast<const float4*>(beta + c0 + 4);
    return a;
}
__device__ __forceinline__ float gn_relu1(float x, float mean, float rstd, float g, float b) {
    return fmaxf(__fmaf_rn(__fmul_rn(__fsub_rn(x, mean), rstd), g, b), 0.f);
}
__device__ __forceinline__ uint4 gn_relu8(uint4 a, float mean, float rstd, const GnAffine8& af) {
    const __nv_bfloat162* a2 = reinterpret_cast<const __nv_bfloat162*>(&a);
    uint32_t o[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const float2 f = __bfloat1622float2(a2[t]);
        o[t] = pack_bf16x2(gn_relu1(f.x, mean, rstd, af.g[2 * t], af.b[2 * t]), gn_relu1(f.y, mean, rstd, af.g[2 * t + 1], af.b[2 * t + 1]));
    }
    return make_uint4(o[0], o[1], o[2], o[3]);
}

}  // namespace gb
