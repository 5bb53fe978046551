// Shared arithmetic of the decode-step kernels, so that a fused kernel and its stand-alone counterpart run the very same
// instruction sequence (bit-identical results).
#pragma once
#include "ptx.cuh"

namespace gb {

// sum over the S split-K partials ws[s][b][col + {j, j+half}] of the token-major partial layout [S][B][N], in split order
__device__ __forceinline__ void splitk_pair(const float* __restrict__ ws, int S, int B, int N, int b, int col, int j, int half,
                                            float& a1, float& a2) {
    a1 = 0.f; a2 = 0.f;
    for (int s = 0; s < S; ++s) {
        const float* row = ws + ((long long)s * B + b) * N + col + j;
        a1 += __ldcg(row);
        a2 += __ldcg(row + half);
    }
}

// rotate-half RoPE of one (j, j+half) pair ($HF/models/llama/modeling_llama.py:138-168); inputs already rounded to bf16
__device__ __forceinline__ void rope_pair(float x1, float x2, float c, float sn, __nv_bfloat16& o1, __nv_bfloat16& o2) {
    o1 = __float2bfloat16_rn(x1 * c - x2 * sn);
    o2 = __float2bfloat16_rn(x2 * c + x1 * sn);
}

}  // namespace gb
