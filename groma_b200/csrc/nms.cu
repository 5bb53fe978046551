// Batched, fully on-device greedy NMS and descending top-k (one CTA per image, everything in shared memory).
// NMS follows the reference's op exactly (mmcv/ops/nms.py:14-33 + csrc/common/cuda/nms_cuda_kernel.cuh:18-74 +
// csrc/pytorch/cuda/nms_cuda.cu:5-54): optional score filter (score > thr, only when thr > 0), sort by score
// descending, suppress j>i when interS > thr*(Sa+Sb-interS), keep the first max_num.  What changes is the shape of
// the computation: the reference launches per image from a Python loop and copies the bitmask to the host for the
// serial sweep; here sort, bitmask and sweep stay in one kernel for the whole batch, so no D2H sync.
// Ties in score are broken by lower original index first (stable), which is what torch's CPU sort does.
#include "ptx.cuh"
#include "capi_common.h"

namespace gb {

// Bitonic sort of (key desc, idx asc) pairs in shared memory; n_pow2 power of two, all threads participate.
__device__ __forceinline__ bool before(float ka, int ia, float kb, int ib) {
    return (ka > kb) || (ka == kb && ia < ib);
}
__device__ void bitonic_sort_desc(float* key, int* idx, int n_pow2) {
    for (int k = 2; k <= n_pow2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < n_pow2; t += blockDim.x) {
                const int ixj = t ^ j;
                if (ixj > t) {
                    const bool up = (t & k) == 0;  // ascending position order within this subsequence
                    const float ka = key[t], kb = key[ixj];
                    const int ia = idx[t], ib = idx[ixj];
                    const bool a_first = before(ka, ia, kb, ib);
                    if (up ? !a_first : a_first) {
                        key[t] = kb; key[ixj] = ka;
                        idx[t] = ib; idx[ixj] = ia;
                    }
                }
            }
            __syncthreads();
        }
    }
}

__device__ __forceinline__ bool dev_iou_gt(const float* a, const float* b, float offset, float thr) {
    const float left = fmaxf(a[0], b[0]), right = fminf(a[2], b[2]);
    const float top = fmaxf(a[1], b[1]), bottom = fminf(a[3], b[3]);
    const float width = fmaxf(right - left + offset, 0.f), height = fmaxf(bottom - top + offset, 0.f);
    const float interS = width * height;
    const float Sa = (a[2] - a[0] + offset) * (a[3] - a[1] + offset);
    const float Sb = (b[2] - b[0] + offset) * (b[3] - b[1] + offset);
    return interS > thr * (Sa + Sb - interS);
}

// boxes [B, N, 4] xyxy, scores [B, N]; counts[b] (optional) = number of valid entries of image b (<= N).
// out keep [B, max_out] (original indices in score order, -1 padded), num_keep [B], argmax_idx [B] (first index of the
// maximal score among the valid entries: the reference's fallback when nothing survives the score filter).
__global__ void nms_batched_kernel(const float* __restrict__ boxes, const float* __restrict__ scores,
                                   const int* __restrict__ counts, int N, int n_pow2, float iou_thr, float score_thr,
                                   int offset_i, int max_num, long long* __restrict__ keep, int max_out,
                                   int* __restrict__ num_keep, int* __restrict__ argmax_idx) {
    extern __shared__ __align__(16) uint8_t sm_nms[];
    float* key = reinterpret_cast<float*>(sm_nms);                   // [n_pow2]
    int* idx = reinterpret_cast<int*>(key + n_pow2);                  // [n_pow2]
    float* sbox = reinterpret_cast<float*>(idx + n_pow2);             // [n_pow2*4] sorted boxes
    const int words = (n_pow2 + 63) / 64;
    unsigned long long* mask = reinterpret_cast<unsigned long long*>(sbox + n_pow2 * 4);  // [n_pow2][words]
    __shared__ int s_n;
    const int b = blockIdx.x;
    const int n_in = counts ? min(counts[b], N) : N;
    const float* bx = boxes + (long long)b * N * 4;
    const float* sc = scores + (long long)b * N;
    const bool filter = score_thr > 0.f;
    const float offset = (float)offset_i;

    for (int t = threadIdx.x; t < n_pow2; t += blockDim.x) {
        bool ok = t < n_in;
        float s = ok ? sc[t] : 0.f;
        if (ok && filter && !(s > score_thr)) ok = false;
        key[t] = ok ? s : -INFINITY;
        idx[t] = ok ? t : 0x7fffffff;  // invalid entries sort last
    }
    __syncthreads();
    // argmax over the valid inputs (first max), before sorting
    if (threadIdx.x < 32) {
        float best = -INFINITY; int bi = 0x7fffffff;
        for (int t = threadIdx.x; t < n_in; t += 32) {
            const float s = sc[t];
            if (s > best || (s == best && t < bi)) { best = s; bi = t; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if (threadIdx.x == 0) argmax_idx[b] = (n_in > 0) ? bi : -1;
    }
    bitonic_sort_desc(key, idx, n_pow2);
    if (threadIdx.x == 0) {
        // number of valid (finite-key) entries: they are a prefix after sorting
        int lo = 0, hi = n_pow2;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (idx[mid] != 0x7fffffff) lo = mid + 1; else hi = mid; }
        s_n = lo;
    }
    __syncthreads();
    const int n = s_n;
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
        const float4 v = *reinterpret_cast<const float4*>(bx + (long long)idx[t] * 4);
        sbox[t * 4] = v.x; sbox[t * 4 + 1] = v.y; sbox[t * 4 + 2] = v.z; sbox[t * 4 + 3] = v.w;
    }
    __syncthreads();
    // suppression bitmask: bit j of row i set iff j > i and IoU(i, j) > thr
    for (int w = threadIdx.x; w < n * words; w += blockDim.x) {
        const int i = w / words, cw = w % words;
        unsigned long long bits = 0ull;
        const int j0 = cw * 64;
        if (j0 + 63 > i) {
            for (int jj = 0; jj < 64; ++jj) {
                const int j = j0 + jj;
                if (j > i && j < n && dev_iou_gt(sbox + i * 4, sbox + j * 4, offset, iou_thr)) bits |= 1ull << jj;
            }
        }
        mask[i * words + cw] = bits;
    }
    __syncthreads();
    // serial sweep by warp 0: lane l owns remv word l (words <= 32 -> n_pow2 <= 2048)
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        unsigned long long remv = 0ull;
        int kept = 0;
        const int limit = max_num > 0 ? min(max_num, max_out) : max_out;
        for (int i = 0; i < n && kept < limit; ++i) {
            const unsigned long long wv = __shfl_sync(0xffffffffu, remv, i >> 6);
            if (!((wv >> (i & 63)) & 1ull)) {
                if (lane == 0) keep[(long long)b * max_out + kept] = idx[i];
                ++kept;
                if (lane < words) remv |= mask[i * words + lane];
            }
        }
        for (int t = kept + lane; t < max_out; t += 32) keep[(long long)b * max_out + t] = -1;
        if (lane == 0) num_keep[b] = kept;
    }
}

// scores [B, ld] -> idx [B, k] of the k largest (descending; ties: lower index first)
__global__ void topk_desc_kernel(const float* __restrict__ scores, long long ld, int N, int n_pow2, int k,
                                 long long* __restrict__ out) {
    extern __shared__ __align__(16) uint8_t sm_topk[];
    float* key = reinterpret_cast<float*>(sm_topk);
    int* idx = reinterpret_cast<int*>(key + n_pow2);
    const float* sc = scores + blockIdx.x * ld;
    for (int t = threadIdx.x; t < n_pow2; t += blockDim.x) {
        key[t] = t < N ? sc[t] : -INFINITY;
        idx[t] = t < N ? t : 0x7fffffff;
    }
    __syncthreads();
    bitonic_sort_desc(key, idx, n_pow2);
    for (int t = threadIdx.x; t < k; t += blockDim.x) out[(long long)blockIdx.x * k + t] = idx[t];
}

static int next_pow2(int n) { int p = 1; while (p < n) p <<= 1; return p; }

}  // namespace gb
using namespace gb;

// C-ABI twin of mmcv `nms` (pybind.cpp:175,596) + the Python-side filtering of mmcv/ops/nms.py:14-33, batched.
GROMA_API int32_t groma_nms_batched(const float* boxes, const float* scores, const int32_t* counts, int32_t B,
                                    int32_t N, float iou_threshold, float score_threshold, int32_t offset,
                                    int32_t max_num, int64_t* keep, int32_t max_out, int32_t* num_keep,
                                    int32_t* argmax_idx, void* stream) {
    if (!boxes || !scores || !keep || !num_keep || !argmax_idx || B <= 0 || N <= 0 || max_out <= 0) return GROMA_ERR_ARG;
    if (reinterpret_cast<uintptr_t>(boxes) & 15) return GROMA_ERR_ALIGN;
    const int np2 = next_pow2(N);
    if (np2 > 1024) return GROMA_ERR_UNSUPPORTED;
    const int words = (np2 + 63) / 64;
    const size_t smem = (size_t)np2 * (4 + 4 + 16) + (size_t)np2 * words * 8;
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        if (cudaFuncSetAttribute(nms_batched_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return GROMA_ERR_CUDA;
        configured = smem;
    }
    nms_batched_kernel<<<B, 512, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
        boxes, scores, counts, N, np2, iou_threshold, score_threshold, offset, max_num,
        reinterpret_cast<long long*>(keep), max_out, num_keep, argmax_idx);
    return GROMA_LAUNCH_CHECK();
}

GROMA_API int32_t groma_topk_desc(const float* scores, int64_t ld, int32_t B, int32_t N, int32_t k, int64_t* out_idx,
                                  void* stream) {
    if (!scores || !out_idx || B <= 0 || N <= 0 || k <= 0 || k > N) return GROMA_ERR_ARG;
    const int np2 = next_pow2(N);
    if (np2 > 4096) return GROMA_ERR_UNSUPPORTED;
    topk_desc_kernel<<<B, 512, (size_t)np2 * 8, reinterpret_cast<cudaStream_t>(stream)>>>(
        scores, ld, N, np2, k, reinterpret_cast<long long*>(out_idx));
    return GROMA_LAUNCH_CHECK();
}
