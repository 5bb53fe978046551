// Flash-attention forward on the 5th-generation tensor cores (sm_100a): S = Q K^T and O = P V are tcgen05.mma with fp32
// accumulators in TMEM; Q/K/V tiles arrive by TMA (128B swizzle); the softmax runs on 128 threads (one query row per TMEM
// lane), writes P as bf16 straight into the swizzled shared-memory operand layout, and keeps the running output in
// registers (rescaled once per KV tile).  Replaces the mma.sync kernel for head dims 64 (DINOv2) and 128 (LLaMA prefill):
//   $HF/models/llama/modeling_llama.py:199-289 (causal + key-padding), $HF/models/dinov2/modeling_dinov2.py:153-179.
//
// CTA = 192 threads: warp 0 TMA producer, warp 1 MMA issuer (one thread), warps 2-5 softmax / output.
// Per KV tile j (128 keys):   QK_j -> [softmax_j: 2 passes over S in TMEM, P_j -> smem] -> PV_j -> O_j added in registers,
// with QK_{j+1} issued right behind PV_j so the tensor pipe works while the softmax warps fold O_j.
// V is consumed in its natural [key][d] layout as an MN-major B operand (no transpose pass).
#include "ptx.cuh"
#include "capi_common.h"

namespace gb {

constexpr int FA_BM = 128, FA_BN = 128, FA_THREADS = 192, FA_KV_STAGES = 2;

struct FaParams {
    CUtensorMap tma_q, tma_k, tma_v;        // 2D [rows, cols] bf16, box {64, 128}
    __nv_bfloat16* o; long long o_ld;       // o[(b*Sq + i) * o_ld + h*D + d]
    const int* kv_len;                      // optional [B]
    int Sq, Sk, H, B;
    int q_batch_rows, q_head_cols;          // Q row = b*q_batch_rows + i, col = h*q_head_cols
    int k_batch_rows, k_head_rows, k_head_cols;   // K/V row = b*k_batch_rows + h*k_head_rows + j, col = h*k_head_cols
    int v_batch_rows, v_head_rows, v_head_cols;
    int q_pos0, causal;
    float scale_log2;
};

// UMMA smem descriptor, SWIZZLE_128B, explicit LBO/SBO (bytes)
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
    uint64_t desc = 0;
    desc |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
    desc |= static_cast<uint64_t>((lbo >> 4) & 0x3FFF) << 16;
    desc |= static_cast<uint64_t>((sbo >> 4) & 0x3FFF) << 32;
    desc |= static_cast<uint64_t>(1) << 46;
    desc |= static_cast<uint64_t>(2) << 61;
    return desc;
}
__host__ __device__ constexpr uint32_t make_idesc_bf16_ex(uint32_t m, uint32_t n, uint32_t b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (0u << 15) | (b_mn_major << 16) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

template <int D>
__global__ void __launch_bounds__(FA_THREADS, 1) attention_tcgen05_kernel(const __grid_constant__ FaParams p) {
    static_assert(D == 64 || D == 128, "head dim");
    constexpr int DBLK = D / 64;                        // 64-column (128-byte) blocks per row of Q/K/V
    constexpr int Q_BYTES = FA_BM * D * 2;
    constexpr int KV_BYTES = FA_BN * D * 2;
    constexpr int P_BYTES = FA_BM * FA_BN * 2;          // two 64-key blocks of [128 rows][128 B]
    constexpr int TMEM_COLS = 256;                      // S: cols [0,128), O tile: cols [128, 128+D)
    extern __shared__ uint8_t smem_raw_fa[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw_fa) + 1023) & ~uintptr_t(1023));
    uint8_t* Qs = smem;
    uint8_t* Ks = Qs + Q_BYTES;                         // [stages][KV_BYTES]
    uint8_t* Vs = Ks + FA_KV_STAGES * KV_BYTES;
    uint8_t* Ps = Vs + FA_KV_STAGES * KV_BYTES;
    uint64_t* q_full = reinterpret_cast<uint64_t*>(Ps + P_BYTES);
    uint64_t* kv_full = q_full + 1;                     // [stages]
    uint64_t* kv_empty = kv_full + FA_KV_STAGES;        // [stages]
    uint64_t* s_full = kv_empty + FA_KV_STAGES;
    uint64_t* p_ready = s_full + 1;
    uint64_t* o_full = p_ready + 1;
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(o_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int q0 = qb * FA_BM;
    int sk = p.Sk;
    if (p.kv_len) sk = min(sk, p.kv_len[b]);
    int k_end = sk;
    if (p.causal) k_end = min(sk, p.q_pos0 + min(q0 + FA_BM, p.Sq));
    const int n_tiles = (k_end + FA_BN - 1) / FA_BN;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.tma_q); tma_prefetch_desc(&p.tma_k); tma_prefetch_desc(&p.tma_v);
    }
    if (warp == 1) {
        if (lane == 0) {
            mbar_init(q_full, 1);
            for (int i = 0; i < FA_KV_STAGES; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
            mbar_init(s_full, 1);
            mbar_init(p_ready, 4);
            mbar_init(o_full, 1);
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc<TMEM_COLS>(tmem_holder);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    const uint32_t tmem_s = tmem_base, tmem_o = tmem_base + 128;

    if (warp == 0) {
        if (lane == 0 && n_tiles > 0) {
            // ---------------- TMA producer
            mbar_expect_tx(q_full, Q_BYTES);
#pragma unroll
            for (int blk = 0; blk < DBLK; ++blk)
                tma_load_2d(Qs + blk * (FA_BM * 128), &p.tma_q, q_full, h * p.q_head_cols + blk * 64, b * p.q_batch_rows + q0);
            for (int j = 0; j < n_tiles; ++j) {
                const int st = j % FA_KV_STAGES;
                const uint32_t ph = (j / FA_KV_STAGES) & 1;
                mbar_wait(&kv_empty[st], ph ^ 1);
                mbar_expect_tx(&kv_full[st], 2 * KV_BYTES);
                const int krow = b * p.k_batch_rows + h * p.k_head_rows + j * FA_BN;
                const int vrow = b * p.v_batch_rows + h * p.v_head_rows + j * FA_BN;
#pragma unroll
                for (int blk = 0; blk < DBLK; ++blk) {
                    tma_load_2d(Ks + st * KV_BYTES + blk * (FA_BN * 128), &p.tma_k, &kv_full[st], h * p.k_head_cols + blk * 64, krow);
                    tma_load_2d(Vs + st * KV_BYTES + blk * (FA_BN * 128), &p.tma_v, &kv_full[st], h * p.v_head_cols + blk * 64, vrow);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && n_tiles > 0) {
            // ---------------- MMA issuer
            constexpr uint32_t idesc_qk = make_idesc_bf16_ex(FA_BM, FA_BN, 0);   // S[128 x 128] = Q (K-major) . K^T (K-major)
            constexpr uint32_t idesc_pv = make_idesc_bf16_ex(FA_BM, D, 1);       // O[128 x D]   = P (K-major) . V (MN-major)
            const uint32_t q_addr = smem_u32(Qs), p_addr = smem_u32(Ps);
            auto issue_qk = [&](int j) {
                const int st = j % FA_KV_STAGES;
                mbar_wait(&kv_full[st], (j / FA_KV_STAGES) & 1);
                tc_fence_after();
                const uint32_t k_addr = smem_u32(Ks + st * KV_BYTES);
#pragma unroll
                for (int kk = 0; kk < D / 16; ++kk) {
                    const uint32_t off = (kk >> 2) * (FA_BM * 128) + (kk & 3) * 32;
                    umma_bf16(tmem_s, make_desc_sw128(q_addr + off, 0, 1024), make_desc_sw128(k_addr + off, 0, 1024), idesc_qk,
                              kk > 0 ? 1u : 0u);
                }
                umma_commit(s_full);
            };
            mbar_wait(q_full, 0);
            issue_qk(0);
            for (int j = 0; j < n_tiles; ++j) {
                const int st = j % FA_KV_STAGES;
                mbar_wait(p_ready, j & 1);          // P_j is in smem, S_j and O_{j-1} have been consumed
                tc_fence_after();
                const uint32_t v_addr = smem_u32(Vs + st * KV_BYTES);
#pragma unroll
                for (int kk = 0; kk < FA_BN / 16; ++kk) {
                    const uint64_t da = make_desc_sw128(p_addr + (kk >> 2) * (FA_BM * 128) + (kk & 3) * 32, 0, 1024);
                    // V tile: [128 keys][D] as D/64 blocks of [128 rows][128 B]; MN-major: SBO = 8 k-rows, LBO = next 64-wide d block
                    const uint64_t db = make_desc_sw128(v_addr + kk * 2048, FA_BN * 128, 1024);
                    umma_bf16(tmem_o, da, db, idesc_pv, kk > 0 ? 1u : 0u);
                }
                umma_commit(o_full);
                umma_commit(&kv_empty[st]);         // K_j / V_j (and P_j) free once everything issued so far has completed
                if (j + 1 < n_tiles) issue_qk(j + 1);
            }
        }
    } else {
        // ---------------- softmax + output (thread = query row = TMEM lane)
        const int qw = warp & 3;
        const int r = qw * 32 + lane;
        const int qi = q0 + r;
        const uint32_t lane_sel = uint32_t(qw * 32) << 16;
        float o_acc[D];
#pragma unroll
        for (int i = 0; i < D; ++i) o_acc[i] = 0.f;
        float m_run = -INFINITY, l_run = 0.f;
        const int q_limit = p.causal ? (p.q_pos0 + qi) : 0x7fffffff;   // last visible key for this row
        for (int j = 0; j < n_tiles; ++j) {
            const int j0 = j * FA_BN;
            mbar_wait(s_full, j & 1);
            tc_fence_after();
            float mx = -INFINITY;
#pragma unroll
            for (int c = 0; c < FA_BN / 16; ++c) {
                uint32_t v[32];
                tmem_ld16(tmem_s + lane_sel + c * 16, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int key = j0 + c * 16 + i;
                    const float x = (key < sk && key <= q_limit) ? __uint_as_float(v[i]) * p.scale_log2 : -INFINITY;
                    mx = fmaxf(mx, x);
                }
            }
            const float m_new = fmaxf(m_run, mx);
            const float m_safe = (m_new == -INFINITY) ? 0.f : m_new;
            const float corr = exp2f(m_run - m_safe);
            m_run = m_new;
            float rowsum = 0.f;
#pragma unroll
            for (int c = 0; c < FA_BN / 16; ++c) {
                uint32_t v[32];
                tmem_ld16(tmem_s + lane_sel + c * 16, v);
                tmem_ld_wait();
                uint32_t pk[8];
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    const int key = j0 + c * 16 + i;
                    const float x0 = (key < sk && key <= q_limit) ? __uint_as_float(v[i]) * p.scale_log2 : -INFINITY;
                    const float x1 = (key + 1 < sk && key + 1 <= q_limit) ? __uint_as_float(v[i + 1]) * p.scale_log2 : -INFINITY;
                    const float p0 = exp2f(x0 - m_safe), p1 = exp2f(x1 - m_safe);
                    rowsum += p0 + p1;
                    pk[i >> 1] = pack_bf16x2(p0, p1);
                }
                // 16 keys = 32 bytes = 2 chunks of the 128-byte swizzled row of the 64-key block (c >> 2)
                uint8_t* prow = Ps + (c >> 2) * (FA_BM * 128) + r * 128;
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    const int chunk = ((c & 3) * 2 + t) ^ (r & 7);
                    *reinterpret_cast<uint4*>(prow + chunk * 16) = make_uint4(pk[4 * t], pk[4 * t + 1], pk[4 * t + 2], pk[4 * t + 3]);
                }
            }
            l_run = l_run * corr + rowsum;
            fence_proxy_async();        // generic-proxy smem writes -> visible to the tensor core (async proxy)
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_ready);
            mbar_wait(o_full, j & 1);
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < D / 16; ++c) {
                uint32_t v[32];
                tmem_ld16(tmem_o + lane_sel + c * 16, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) o_acc[c * 16 + i] = o_acc[c * 16 + i] * corr + __uint_as_float(v[i]);
            }
            tc_fence_before();
        }
        if (qi < p.Sq) {
            const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
            __nv_bfloat16* dst = p.o + ((long long)b * p.Sq + qi) * p.o_ld + (long long)h * D;
#pragma unroll
            for (int i = 0; i < D; i += 8)
                *reinterpret_cast<uint4*>(dst + i) =
                    make_uint4(pack_bf16x2(o_acc[i] * inv, o_acc[i + 1] * inv), pack_bf16x2(o_acc[i + 2] * inv, o_acc[i + 3] * inv),
                               pack_bf16x2(o_acc[i + 4] * inv, o_acc[i + 5] * inv), pack_bf16x2(o_acc[i + 6] * inv, o_acc[i + 7] * inv));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<TMEM_COLS>(tmem_base);
    }
}

typedef CUresult (*PFN_encodeTiledFa)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiledFa fa_encode_fn() {
    static PFN_encodeTiledFa fn = nullptr;
    if (fn) return fn;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = reinterpret_cast<PFN_encodeTiledFa>(ptr);
    return fn;
}
static int fa_make_map(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld) {
    PFN_encodeTiledFa enc = fa_encode_fn();
    if (!enc) return GROMA_ERR_DRIVER;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld * 2};
    cuuint32_t box[2] = {64, 128};
    cuuint32_t estr[2] = {1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS
               ? GROMA_OK : GROMA_ERR_TMA_ENCODE;
}

template <int D>
static int launch_fa(const FaParams& p, cudaStream_t st) {
    constexpr int SMEM = FA_BM * D * 2 + 2 * FA_KV_STAGES * FA_BN * D * 2 + FA_BM * FA_BN * 2 + 1024 + 256;
    static bool set = false;
    if (!set) {
        if (cudaFuncSetAttribute(attention_tcgen05_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM) != cudaSuccess) return GROMA_ERR_CUDA;
        set = true;
    }
    dim3 grid((p.Sq + FA_BM - 1) / FA_BM, p.H, p.B);
    attention_tcgen05_kernel<D><<<grid, FA_THREADS, SMEM, st>>>(p);
    return GROMA_LAUNCH_CHECK();
}

}  // namespace gb
using namespace gb;

// q/k/v are 2-D bf16 row-major views [rows, cols] with row stride ld (elements); element (b, i|j, h, d) lives at
//   q: row b*q_batch_rows + i,                   col h*q_head_cols + d
//   k: row b*k_batch_rows + h*k_head_rows + j,   col h*k_head_cols + d      (v alike)
// which covers both the KV cache [B,H,cap,D] and fused qkv activations [B*S, 3*H*D].  o: [B*Sq, o_ld], col h*D + d.
GROMA_API int32_t groma_attention_tc(const void* q, int64_t q_rows, int64_t q_cols, int64_t q_ld, int32_t q_batch_rows, int32_t q_head_cols,
                                     const void* k, int64_t k_rows, int64_t k_cols, int64_t k_ld, int32_t k_batch_rows, int32_t k_head_rows,
                                     int32_t k_head_cols, const void* v, int64_t v_rows, int64_t v_cols, int64_t v_ld, int32_t v_batch_rows,
                                     int32_t v_head_rows, int32_t v_head_cols, void* o, int64_t o_ld, const int32_t* kv_len, int32_t B,
                                     int32_t H, int32_t Sq, int32_t Sk, int32_t D, int32_t causal, int32_t q_pos0, float scale, void* stream) {
    if (!q || !k || !v || !o || B <= 0 || H <= 0 || Sq <= 0 || Sk <= 0) return GROMA_ERR_ARG;
    if (D != 64 && D != 128) return GROMA_ERR_UNSUPPORTED;
    if ((q_ld & 7) || (k_ld & 7) || (v_ld & 7) || (o_ld & 7) || (reinterpret_cast<uintptr_t>(q) & 15) || (reinterpret_cast<uintptr_t>(k) & 15) ||
        (reinterpret_cast<uintptr_t>(v) & 15) || (reinterpret_cast<uintptr_t>(o) & 15))
        return GROMA_ERR_ALIGN;
    FaParams p;
    int rc = fa_make_map(&p.tma_q, q, (uint64_t)q_rows, (uint64_t)q_cols, (uint64_t)q_ld); if (rc) return rc;
    rc = fa_make_map(&p.tma_k, k, (uint64_t)k_rows, (uint64_t)k_cols, (uint64_t)k_ld); if (rc) return rc;
    rc = fa_make_map(&p.tma_v, v, (uint64_t)v_rows, (uint64_t)v_cols, (uint64_t)v_ld); if (rc) return rc;
    p.o = reinterpret_cast<__nv_bfloat16*>(o); p.o_ld = o_ld; p.kv_len = kv_len;
    p.Sq = Sq; p.Sk = Sk; p.H = H; p.B = B;
    p.q_batch_rows = q_batch_rows; p.q_head_cols = q_head_cols;
    p.k_batch_rows = k_batch_rows; p.k_head_rows = k_head_rows; p.k_head_cols = k_head_cols;
    p.v_batch_rows = v_batch_rows; p.v_head_rows = v_head_rows; p.v_head_cols = v_head_cols;
    p.q_pos0 = q_pos0; p.causal = causal; p.scale_log2 = scale * 1.4426950408889634f;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    return D == 128 ? launch_fa<128>(p, st) : launch_fa<64>(p, st);
}
