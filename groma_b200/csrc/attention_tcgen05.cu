// Flash-attention forward on the 5th-generation tensor cores (sm_100a) for head dims 64 (DINOv2, Deformable-DETR is 32 and
// stays on the small kernel) and 128 (LLaMA prefill):
//   $HF/models/llama/modeling_llama.py:199-289 (causal + key-padding), $HF/models/dinov2/modeling_dinov2.py:153-179.
//
// One CTA = TWO 128-row query tiles (A, B) of one (batch, head) in ping-pong over the same K/V stream:
//   warp 0        TMA producer: Q_A, Q_B once, then K_j / V_j tiles of 128 keys into separate 2-deep rings
//                 (warps 0-3 form one warpgroup that hands its registers to the softmax warpgroups with setmaxnreg)
//   warp 1        one thread issues every tcgen05.mma:  S_X = Q_X K_j^T (operands in shared memory), O_X += P_X V_j with
//                 P_X read straight from TENSOR MEMORY (it overwrites the first 64 columns of S_X as packed bf16) and V
//                 consumed in its natural [key][d] layout as an MN-major B operand
//   warps 4-7     softmax of tile A, warps 8-11 softmax of tile B: thread = query row = TMEM lane; ONE pass over the 128 scores
//                 of the row held in registers (max, exp2, row sum, bf16 pack, tcgen05.st)
// Issue order  QK_A(j+1) right behind PV_A(j), QK_B(j+1) behind PV_B(j): while the softmax warps of one tile work, the tensor
// pipe runs the other tile's PV and next QK.  The running output stays in TMEM (fp32, accumulated by the MMA); it is only
// touched by the softmax warps when the running row maximum has grown by more than 2^8 since the value the exponentials are
// currently referenced to (then O and the row sum are rescaled once) and at the end (O / l -> bf16).  Everything else --
// masks, scale, log2(e) folding -- is one FFMA + MUFU.EX2 per score.
// TMEM: S_A [0,128) S_B [128,256) O_A [256,256+D) O_B [384,384+D) = all 512 columns, one CTA per SM.
#include "ptx.cuh"
#include "capi_common.h"
#include "groma_b200.h"
#include <cstdlib>

namespace gb {

constexpr int FA_BM = 128, FA_BN = 128, FA_THREADS = 384, FA_STAGES = 2;

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
    // raw views of the same tensors for the few keys past the last full 128-key tile (k_tail_max > 0: they are folded into the
    // epilogue on CUDA cores instead of costing a whole masked tile pass); o_batch_rows = rows per batch of `o` (>= Sq)
    const __nv_bfloat16 *q_ptr, *k_ptr, *v_ptr;
    long long q_ld, k_ld, v_ld;
    int k_tail_max, o_batch_rows;
    int n_main;                             // query-tile pairs (grid.x)
    int q_tail;                             // 1: query row n_main * 256 (the ViT's 1025th token) runs on the two idle warps of the
                                            // producer warpgroup of CTA blockIdx.x == 0 of its head, against the K / V tiles in smem
};
__device__ __forceinline__ float2 bf16x2_to_f2(uint32_t w) { return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u)); }
constexpr int FA_KTAIL = 8;   // at most this many trailing keys are handled in the epilogue (non-causal only)

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
// D[tmem] (+)= A[tmem] * B[smem desc]: the A operand (P, bf16 pairs packed in 32-bit columns, row = lane) comes from TMEM
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]),
        "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]),
        "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32_raw(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int D>
struct FaSmem {
    static constexpr int DBLK = D / 64;                 // 64-column (128-byte) blocks per row of Q/K/V
    static constexpr int TILE_BYTES = 128 * D * 2;      // one 128-row tile of Q, K or V
    static constexpr int TAIL_BYTES = 2 * FA_KTAIL * D * 2;   // trailing K and V rows, staged for the epilogue
    // q_tail: the query row (bf16, then fp32) + one warp's (m, l, acc[D]) for the merge + each warp's 64 rounded probabilities
    static constexpr int QT_BYTES = D * 2 + D * 4 + (D + 2) * 4 + 8 + 2 * 64 * 4;
    static constexpr int BYTES = 2 * TILE_BYTES + 2 * FA_STAGES * TILE_BYTES + 1024 + 256 + TAIL_BYTES + QT_BYTES;
};

template <int D>
__global__ void __launch_bounds__(FA_THREADS, 1) attention_fa2q_kernel(const __grid_constant__ FaParams p) {
    static_assert(D == 64 || D == 128, "head dim");
    using L = FaSmem<D>;
    constexpr int DBLK = L::DBLK, TILE = L::TILE_BYTES;
    extern __shared__ uint8_t smem_raw_fa[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw_fa) + 1023) & ~uintptr_t(1023));
    uint8_t* Qs = smem;                                 // [2 tiles][DBLK][128 rows][128 B]
    uint8_t* Ks = Qs + 2 * TILE;                        // [stages][DBLK][128][128 B]
    uint8_t* Vs = Ks + FA_STAGES * TILE;
    uint64_t* q_full = reinterpret_cast<uint64_t*>(Vs + FA_STAGES * TILE);
    uint64_t* k_full = q_full + 1;                      // [stages]
    uint64_t* k_empty = k_full + FA_STAGES;
    uint64_t* v_full = k_empty + FA_STAGES;
    uint64_t* v_empty = v_full + FA_STAGES;
    uint64_t* s_full = v_empty + FA_STAGES;             // [2 tiles]
    uint64_t* p_ready = s_full + 2;                     // [2]
    uint64_t* o_done = p_ready + 2;                     // [2]
    uint64_t* tail_full = o_done + 2;
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tail_full + 1);
    __nv_bfloat16* tail_k = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<uint8_t*>(q_full) + 256);   // [FA_KTAIL][D]
    __nv_bfloat16* tail_v = tail_k + FA_KTAIL * D;
    __nv_bfloat16* qt_q = tail_v + FA_KTAIL * D;                                  // [D]   the tail query row
    float* qt_qf = reinterpret_cast<float*>(qt_q + D);                             // [D]   the same row in fp32
    float* qt_merge = qt_qf + D;                                                   // [2 + D] warp 3's partial state (+ 2 pad)
    float* qt_p = qt_merge + D + 4;                                                // [2 warps][64] bf16-rounded probabilities of a tile

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // heavy (late, for causal) query blocks first: they have the most key tiles
    const int h = blockIdx.y, b = blockIdx.z;
    const int qb = p.n_main - 1 - (int)blockIdx.x;
    const int q0 = qb * 2 * FA_BM;
    int sk = p.Sk;
    if (p.kv_len) sk = min(sk, p.kv_len[b]);
    // S = 1025 (ViT: 1024 patches + [CLS]) would spend a ninth, 127/128-masked tile pass on ONE key: up to k_tail_max trailing
    // keys are instead added by each row's own thread in the epilogue (exact online-softmax merge, fp32 dot products)
    const int k_tail = (!p.causal && (sk % FA_BN) <= p.k_tail_max) ? (sk % FA_BN) : 0;
    sk -= k_tail;
    int n_x[2];                                         // key tiles each query tile has to visit
#pragma unroll
    for (int X = 0; X < 2; ++X) {
        const int first = q0 + X * FA_BM;
        int k_end = sk;
        if (p.causal) k_end = min(sk, p.q_pos0 + min(first + FA_BM, p.Sq));
        n_x[X] = (first < p.Sq) ? (k_end + FA_BN - 1) / FA_BN : 0;
    }
    const int n_max = max(n_x[0], n_x[1]);
    // q_tail: this CTA also owns the one query row past the last full pair of tiles (non-causal, D = 64: the ViT)
    const bool tail_cta = (D == 64) && p.q_tail > 0 && blockIdx.x == 0;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.tma_q); tma_prefetch_desc(&p.tma_k); tma_prefetch_desc(&p.tma_v);
    }
    if (warp == 1) {
        if (lane == 0) {
            mbar_init(q_full, 1);
            // a stage is free once the MMAs that read it are complete AND, in a tail CTA, both tail warps are done with it
            const uint32_t readers = 1 + (tail_cta ? 2 : 0);
            for (int i = 0; i < FA_STAGES; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], readers); mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], readers); }
            for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_ready[i], 4); mbar_init(&o_done[i], 1); }
            mbar_init(tail_full, 1);
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc<512>(tmem_holder);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    // register hand-over between warpgroups: the softmax threads keep a whole 128-score row in registers
    if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 48;" ::: "memory");
    if (warp == 0) {
        if (lane == 0 && n_max > 0) {
            // ---------------- TMA producer
            mbar_expect_tx(q_full, 2 * TILE);
#pragma unroll
            for (int X = 0; X < 2; ++X)
#pragma unroll
                for (int blk = 0; blk < DBLK; ++blk)
                    tma_load_2d(Qs + X * TILE + blk * (FA_BM * 128), &p.tma_q, q_full, h * p.q_head_cols + blk * 64, b * p.q_batch_rows + q0 + X * FA_BM);
            for (int j = 0; j < n_max; ++j) {
                const int st = j % FA_STAGES;
                const uint32_t ph = (j / FA_STAGES) & 1;
                const int krow = b * p.k_batch_rows + h * p.k_head_rows + j * FA_BN;
                const int vrow = b * p.v_batch_rows + h * p.v_head_rows + j * FA_BN;
                mbar_wait(&k_empty[st], ph ^ 1);
                mbar_expect_tx(&k_full[st], TILE);
#pragma unroll
                for (int blk = 0; blk < DBLK; ++blk)
                    tma_load_2d(Ks + st * TILE + blk * (FA_BN * 128), &p.tma_k, &k_full[st], h * p.k_head_cols + blk * 64, krow);
                mbar_wait(&v_empty[st], ph ^ 1);
                mbar_expect_tx(&v_full[st], TILE);
#pragma unroll
                for (int blk = 0; blk < DBLK; ++blk)
                    tma_load_2d(Vs + st * TILE + blk * (FA_BN * 128), &p.tma_v, &v_full[st], h * p.v_head_cols + blk * 64, vrow);
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && n_max > 0) {
            // ---------------- MMA issuer
            constexpr uint32_t idesc_qk = make_idesc_bf16_ex(FA_BM, FA_BN, 0);   // S[128 x 128] = Q (K-major) . K^T (K-major)
            constexpr uint32_t idesc_pv = make_idesc_bf16_ex(FA_BM, D, 1);       // O[128 x D]   = P (TMEM)    . V (MN-major)
            auto issue_qk = [&](int X, int j) {
                const int st = j % FA_STAGES;
                const uint32_t q_addr = smem_u32(Qs + X * TILE), k_addr = smem_u32(Ks + st * TILE);
#pragma unroll
                for (int kk = 0; kk < D / 16; ++kk) {
                    const uint32_t off = (kk >> 2) * (FA_BM * 128) + (kk & 3) * 32;
                    umma_bf16(tmem_base + X * 128, make_desc_sw128(q_addr + off, 0, 1024), make_desc_sw128(k_addr + off, 0, 1024), idesc_qk,
                              kk > 0 ? 1u : 0u);
                }
                umma_commit(&s_full[X]);
            };
            auto issue_pv = [&](int X, int j) {
                const int st = j % FA_STAGES;
                const uint32_t v_addr = smem_u32(Vs + st * TILE);
#pragma unroll
                for (int kk = 0; kk < FA_BN / 16; ++kk) {
                    // V tile: [128 keys][D] as D/64 blocks of [128 rows][128 B]; MN-major: SBO = 8 k-rows, LBO = next 64-wide d block
                    const uint64_t db = make_desc_sw128(v_addr + kk * 2048, FA_BN * 128, 1024);
                    umma_bf16_ts(tmem_base + 256 + X * 128, tmem_base + X * 128 + kk * 8, db, idesc_pv, (j > 0 || kk > 0) ? 1u : 0u);
                }
                umma_commit(&o_done[X]);
            };
            mbar_wait(q_full, 0);
            mbar_wait(&k_full[0], 0);
            tc_fence_after();
            if (n_x[0] > 0) issue_qk(0, 0);
            if (n_x[1] > 0) issue_qk(1, 0);
            umma_commit(&k_empty[0]);     // K_0 is released once both S_X(0) are complete (commit tracks every MMA issued so far)
            for (int j = 0; j < n_max; ++j) {
                const int st = j % FA_STAGES;
                const uint32_t ph = (j / FA_STAGES) & 1;
                const int stn = (j + 1) % FA_STAGES;
                const uint32_t phn = ((j + 1) / FA_STAGES) & 1;
                mbar_wait(&v_full[st], ph);
                bool k_next_ready = false;
#pragma unroll
                for (int X = 0; X < 2; ++X) {
                    if (j < n_x[X]) {
                        mbar_wait(&p_ready[X], j & 1);      // P_X(j) is in TMEM, S_X(j) consumed, O_X rescaled if it had to be
                        tc_fence_after();
                        issue_pv(X, j);
                    }
                    if (j + 1 < n_x[X]) {
                        if (!k_next_ready) { mbar_wait(&k_full[stn], phn); tc_fence_after(); k_next_ready = true; }
                        issue_qk(X, j + 1);
                    }
                }
                umma_commit(&v_empty[st]);                  // V_j free once both PV(j) are done
                if (k_next_ready) umma_commit(&k_empty[stn]);   // K_{j+1} free once both QK(j+1) are done
            }
        }
    } else if (tail_cta) {
        if constexpr (D == 64) {
        // ---------------- warps 2, 3 of a tail CTA: query row qi = n_main * 256 on CUDA cores, out of the K / V tiles of the ring.
        // Scores: lane = key (the 128-byte K row of a key is read as eight swizzled 16-byte chunks, the query row as broadcast
        // chunks); PV: lane = dimension pair (P goes lane -> warp by shuffle, V as one 4-byte read per key).  Each warp takes 64
        // keys of every tile; exponentials are referenced to a warp-uniform running maximum; the two warps' (m, l, acc) meet in
        // shared memory; the keys past the last full tile (k_tail) are folded in last, as in the tensor-core rows' epilogue.
        const int w = warp - 2;
        const int qi = p.n_main * 2 * FA_BM;
        const uint32_t ks_s = smem_u32(Ks), vs_s = smem_u32(Vs), qf_s = smem_u32(qt_qf), pp_s = smem_u32(qt_p + w * 64);
        if (w == 0) {
            const __nv_bfloat16* qrow = p.q_ptr + ((long long)b * p.q_batch_rows + qi) * p.q_ld + (long long)h * p.q_head_cols;
            const uint32_t qw = *reinterpret_cast<const uint32_t*>(qrow + 2 * lane);
            *reinterpret_cast<uint32_t*>(qt_q + 2 * lane) = qw;
            *reinterpret_cast<float2*>(qt_qf + 2 * lane) = bf16x2_to_f2(qw);     // converted once, not once per key
        }
        asm volatile("bar.sync 5, 64;" ::: "memory");
        float m_run = -INFINITY, l_lane = 0.f, acc0 = 0.f, acc1 = 0.f;
        for (int j = 0; j < n_max; ++j) {
            const int st = j % FA_STAGES;
            const uint32_t ph = (j / FA_STAGES) & 1;
            mbar_wait(&k_full[st], ph);
            float sc[2];
#pragma unroll
            for (int ps = 0; ps < 2; ++ps) {
                const int r = w * 64 + ps * 32 + lane;               // key row inside the tile
                const uint32_t krow = ks_s + st * TILE + r * 128;
                float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;       // four independent chains
#pragma unroll 2                                                     // (48 registers per thread in this warpgroup: a full unroll spills)
                for (int c = 0; c < 8; ++c) {
                    const uint4 kv4 = ld_shared_v4(krow + ((c ^ (r & 7)) << 4));
                    const uint4 qa = ld_shared_v4(qf_s + (c << 5)), qb = ld_shared_v4(qf_s + (c << 5) + 16);   // broadcast
                    const float2 k0 = bf16x2_to_f2(kv4.x), k1 = bf16x2_to_f2(kv4.y), k2 = bf16x2_to_f2(kv4.z), k3 = bf16x2_to_f2(kv4.w);
                    d0 = fmaf(__uint_as_float(qa.x), k0.x, d0); d1 = fmaf(__uint_as_float(qa.y), k0.y, d1);
                    d2 = fmaf(__uint_as_float(qa.z), k1.x, d2); d3 = fmaf(__uint_as_float(qa.w), k1.y, d3);
                    d0 = fmaf(__uint_as_float(qb.x), k2.x, d0); d1 = fmaf(__uint_as_float(qb.y), k2.y, d1);
                    d2 = fmaf(__uint_as_float(qb.z), k3.x, d2); d3 = fmaf(__uint_as_float(qb.w), k3.y, d3);
                }
                sc[ps] = (j * FA_BN + r < sk) ? ((d0 + d1) + (d2 + d3)) * p.scale_log2 : -INFINITY;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&k_empty[st]);
            float tmax = fmaxf(sc[0], sc[1]);
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, off));
            const float m_new = fmaxf(m_run, tmax);                  // warp-uniform
            mbar_wait(&v_full[st], ph);
            if (m_new != -INFINITY) {
                const float corr = ex2_approx(m_run - m_new);        // m_run = -inf -> 0
                const float p0 = ex2_approx(sc[0] - m_new), p1 = ex2_approx(sc[1] - m_new);
                l_lane = l_lane * corr + (p0 + p1);
                acc0 *= corr; acc1 *= corr;
                m_run = m_new;
                // P enters PV in bf16, as on the tensor-core path; lane -> warp through shared memory (16 broadcast LDS.128, not 64 shuffles)
                asm volatile("st.shared.f32 [%0], %1;" ::"r"(pp_s + lane * 4), "f"(__bfloat162float(__float2bfloat16_rn(p0))) : "memory");
                asm volatile("st.shared.f32 [%0], %1;" ::"r"(pp_s + 128 + lane * 4), "f"(__bfloat162float(__float2bfloat16_rn(p1))) : "memory");
                __syncwarp();
                const uint32_t vbase = vs_s + st * TILE + (w * 64) * 128 + (uint32_t)(lane & 3) * 4;
                const uint32_t vch = (uint32_t)lane >> 2;
#pragma unroll 4
                for (int k4 = 0; k4 < 16; ++k4) {
                    const uint4 p4 = ld_shared_v4(pp_s + (k4 << 4));
                    const float pk[4] = {__uint_as_float(p4.x), __uint_as_float(p4.y), __uint_as_float(p4.z), __uint_as_float(p4.w)};
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int rr = k4 * 4 + u;                    // (w * 64 + rr) & 7 == rr & 7
                        uint32_t vw;
                        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(vw) : "r"(vbase + rr * 128 + ((vch ^ (rr & 7)) << 4)) : "memory");
                        const float2 vf = bf16x2_to_f2(vw);
                        acc0 = fmaf(pk[u], vf.x, acc0);
                        acc1 = fmaf(pk[u], vf.y, acc1);
                    }
                }
            }
            __syncwarp();                                            // (also: everyone is done with qt_p before the next tile rewrites it)
            if (lane == 0) mbar_arrive(&v_empty[st]);
        }
        float l_w = l_lane;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) l_w += __shfl_xor_sync(0xffffffffu, l_w, off);
        if (w == 1) {
            if (lane == 0) { qt_merge[0] = m_run; qt_merge[1] = l_w; }
            qt_merge[2 + 2 * lane] = acc0;
            qt_merge[3 + 2 * lane] = acc1;
        }
        asm volatile("bar.sync 5, 64;" ::: "memory");
        if (w == 0) {
            const float m_o = qt_merge[0];
            float m_fin = fmaxf(m_run, m_o);
            float f_a = (m_fin == -INFINITY) ? 0.f : ex2_approx(m_run - m_fin), f_b = (m_fin == -INFINITY) ? 0.f : ex2_approx(m_o - m_fin);
            float l_fin = l_w * f_a + qt_merge[1] * f_b;
            acc0 = acc0 * f_a + qt_merge[2 + 2 * lane] * f_b;
            acc1 = acc1 * f_a + qt_merge[3 + 2 * lane] * f_b;
            if (k_tail > 0) {
                mbar_wait(tail_full, 0);
                const float2 qf = bf16x2_to_f2(*reinterpret_cast<const uint32_t*>(qt_q + 2 * lane));
                for (int t = 0; t < k_tail; ++t) {
                    const float2 kf = bf16x2_to_f2(*reinterpret_cast<const uint32_t*>(tail_k + t * D + 2 * lane));
                    float x = fmaf(qf.x, kf.x, qf.y * kf.y);
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
                    x *= p.scale_log2;
                    const float m2 = fmaxf(m_fin, x);
                    const float f = (m_fin == -INFINITY) ? 0.f : ex2_approx(m_fin - m2);
                    const float pt = ex2_approx(x - m2);
                    l_fin = l_fin * f + pt;
                    const float ptr_ = __bfloat162float(__float2bfloat16_rn(pt));
                    const float2 vf = bf16x2_to_f2(*reinterpret_cast<const uint32_t*>(tail_v + t * D + 2 * lane));
                    acc0 = fmaf(ptr_, vf.x, acc0 * f);
                    acc1 = fmaf(ptr_, vf.y, acc1 * f);
                    m_fin = m2;
                }
            }
            const float inv = (l_fin > 0.f) ? 1.f / l_fin : 0.f;
            __nv_bfloat16* dst = p.o + ((long long)b * p.o_batch_rows + qi) * p.o_ld + (long long)h * D;
            *reinterpret_cast<uint32_t*>(dst + 2 * lane) = pack_bf16x2(acc0 * inv, acc1 * inv);
        }
        }
    }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 224;" ::: "memory");
        // ---------------- softmax + output: thread = query row = TMEM lane
        const int X = (warp - 4) >> 2;
        const int qw = warp & 3;
        const int r = qw * 32 + lane;
        const int qi = q0 + X * FA_BM + r;
        const uint32_t lane_sel = uint32_t(qw * 32) << 16;
        const uint32_t tmem_s = tmem_base + X * 128 + lane_sel, tmem_o = tmem_base + 256 + X * 128 + lane_sel;
        const int nt = n_x[X];
        // one softmax warp stages the trailing key / value rows for the epilogue while the first S tile is still on its way (the
        // global latency is paid here, once, under the TMA + first QK latency)
        if (warp == 4) {
        if (k_tail > 0) {
            constexpr int CPR = D / 8, RPP = 32 / CPR;       // 16-byte chunks per row; rows copied per warp pass
            const int ch = lane % CPR;
            const __nv_bfloat16* ksrc = p.k_ptr + ((long long)b * p.k_batch_rows + (long long)h * p.k_head_rows + sk) * p.k_ld + (long long)h * p.k_head_cols + ch * 8;
            const __nv_bfloat16* vsrc = p.v_ptr + ((long long)b * p.v_batch_rows + (long long)h * p.v_head_rows + sk) * p.v_ld + (long long)h * p.v_head_cols + ch * 8;
            for (int t = lane / CPR; t < k_tail; t += RPP) {
                const uint4 kk = *reinterpret_cast<const uint4*>(ksrc + (long long)t * p.k_ld);
                const uint4 vv = *reinterpret_cast<const uint4*>(vsrc + (long long)t * p.v_ld);
                *reinterpret_cast<uint4*>(tail_k + t * D + ch * 8) = kk;
                *reinterpret_cast<uint4*>(tail_v + t * D + ch * 8) = vv;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(tail_full);
        }
        }
        float m_used = 0.f, l_run = 0.f;
        const int q_limit = p.causal ? (p.q_pos0 + qi) : 0x7fffffff;   // last visible key for this row
        const int tile_first_limit = p.causal ? (p.q_pos0 + q0 + X * FA_BM) : 0x7fffffff;   // smallest q_limit of the tile
        for (int j = 0; j < nt; ++j) {
            const int j0 = j * FA_BN;
            mbar_wait(&s_full[X], j & 1);
            tc_fence_after();
            uint32_t sv[128];
#pragma unroll
            for (int c = 0; c < 4; ++c) tmem_ld32_raw(tmem_s + c * 32, sv + c * 32);
            tmem_ld_wait();
            const bool need_mask = (j0 + FA_BN > sk) || (j0 + FA_BN - 1 > tile_first_limit);   // warp-uniform
            if (need_mask) {
#pragma unroll
                for (int i = 0; i < 128; ++i) {
                    const int key = j0 + i;
                    if (!(key < sk && key <= q_limit)) sv[i] = 0xff800000u;   // -inf
                }
            }
            // four independent chains each: a 128-long dependent max / add chain is 500+ cycles of pure latency per tile
            float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
            for (int i = 0; i < 128; i += 4) {
#pragma unroll
                for (int t = 0; t < 4; ++t) mx4[t] = fmaxf(mx4[t], __uint_as_float(sv[i + t]));
            }
            const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
            float m_new = mx * p.scale_log2;                 // scale > 0
            if (m_new == -INFINITY) m_new = (j == 0) ? 0.f : m_used;
            if (j == 0) {
                m_used = m_new;
            } else {
                // lazy rescale: exponentials stay referenced to m_used until the row maximum has grown by more than 2^8
                const bool grow = m_new - m_used > 8.0f;
                if (__any_sync(0xffffffffu, grow)) {
                    // s_full(j) fired after PV_X(j-1) (same issue thread, in order), so O_X is complete and nobody writes it now
                    const float m_next = fmaxf(m_used, m_new);
                    const float f = ex2_approx(m_used - m_next);
                    m_used = m_next;
                    l_run *= f;
#pragma unroll
                    for (int c = 0; c < D / 32; ++c) {
                        uint32_t ov[32];
                        tmem_ld32_raw(tmem_o + c * 32, ov);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * f);
                        tmem_st32(tmem_o + c * 32, ov);
                    }
                }
            }
            const float neg_m = -m_used;
            float rs4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int i = 0; i < 128; i += 8) {
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    // (moving half of the exponentials to a degree-3 polynomial on the FMA pipe, FlashAttention-4 style, measured SLOWER
                    //  here -- 474 -> 394 TFLOP/s: this loop is issue-bound, not MUFU-bound, on B200)
                    const float p0 = ex2_approx(fmaf(__uint_as_float(sv[i + 2 * t]), p.scale_log2, neg_m));
                    const float p1 = ex2_approx(fmaf(__uint_as_float(sv[i + 2 * t + 1]), p.scale_log2, neg_m));
                    rs4[t] += p0 + p1;
                    sv[(i >> 1) + t] = pack_bf16x2(p0, p1);
                }
            }
            l_run += (rs4[0] + rs4[1]) + (rs4[2] + rs4[3]);
            tmem_st32(tmem_s, sv);
            tmem_st32(tmem_s + 32, sv + 32);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_ready[X]);
        }
        // final O / l -> bf16.  The tensor-memory loads are warp-collective: every lane executes them, only the global stores
        // are guarded (rows past Sq in a partially filled tile).
        if (nt > 0) {
            mbar_wait(&o_done[X], (nt - 1) & 1);
            tc_fence_after();
        }
        // trailing keys (k_tail <= FA_KTAIL): x_t = scale * q_i . k_t in fp32, merged as one more online-softmax step
        float f_old = 1.f, pt[FA_KTAIL];
        const bool row_live = qi < p.Sq;
        if (k_tail > 0) mbar_wait(tail_full, 0);          // completed long ago; acquires warp 2's staging stores
        if (k_tail > 0 && row_live) {
            uint32_t qr[D / 2];
            if (n_max > 0) {
                // the row sits in the Q tile the tensor core read: 128B-swizzled, 16-byte chunk c of row r at chunk c ^ (r & 7)
                mbar_wait(q_full, 0);          // long complete; orders this thread's reads after the TMA writes
                const uint8_t* qt = Qs + X * TILE + r * 128;
#pragma unroll
                for (int i = 0; i < D / 8; ++i)
                    *reinterpret_cast<uint4*>(&qr[4 * i]) = *reinterpret_cast<const uint4*>(qt + (i >> 3) * (FA_BM * 128) + (((i & 7) ^ (r & 7)) << 4));
            } else {
                const __nv_bfloat16* qrow = p.q_ptr + ((long long)b * p.q_batch_rows + qi) * p.q_ld + (long long)h * p.q_head_cols;
#pragma unroll
                for (int i = 0; i < D / 8; ++i) *reinterpret_cast<uint4*>(&qr[4 * i]) = *reinterpret_cast<const uint4*>(qrow + 8 * i);
            }
            float xt[FA_KTAIL];
            float m_fin = (nt > 0) ? m_used : -INFINITY;
#pragma unroll
            for (int t = 0; t < FA_KTAIL; ++t) {
                xt[t] = -INFINITY;
                if (t < k_tail) {
                    const __nv_bfloat16* krow = tail_k + t * D;          // staged by warp 2; same address in every lane: broadcast
                    float a0 = 0.f, a1 = 0.f;
#pragma unroll
                    for (int i = 0; i < D / 8; ++i) {
                        const uint4 kv4 = *reinterpret_cast<const uint4*>(krow + 8 * i);
                        const uint32_t kw[4] = {kv4.x, kv4.y, kv4.z, kv4.w};
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const float2 qf = bf16x2_to_f2(qr[4 * i + u]);
                            const float2 kf = bf16x2_to_f2(kw[u]);
                            a0 = fmaf(qf.x, kf.x, a0);
                            a1 = fmaf(qf.y, kf.y, a1);
                        }
                    }
                    xt[t] = (a0 + a1) * p.scale_log2;
                    m_fin = fmaxf(m_fin, xt[t]);
                }
            }
            f_old = (nt > 0) ? ex2_approx(m_used - m_fin) : 0.f;
            l_run *= f_old;
#pragma unroll
            for (int t = 0; t < FA_KTAIL; ++t) {
                pt[t] = (t < k_tail) ? ex2_approx(xt[t] - m_fin) : 0.f;
                l_run += pt[t];
                pt[t] = __bfloat162float(__float2bfloat16_rn(pt[t]));   // P enters the PV product in bf16, as in the tensor-core path
            }
        } else {
#pragma unroll
            for (int t = 0; t < FA_KTAIL; ++t) pt[t] = 0.f;
        }
        const float inv = (l_run > 0.f) ? 1.f / l_run : 0.f;
        __nv_bfloat16* dst = p.o + ((long long)b * p.o_batch_rows + qi) * p.o_ld + (long long)h * D;
#pragma unroll
        for (int c = 0; c < D / 32; ++c) {
            uint32_t ov[32];
            if (nt > 0) { tmem_ld32_raw(tmem_o + c * 32, ov); tmem_ld_wait(); }
            else {
#pragma unroll
                for (int i = 0; i < 32; ++i) ov[i] = 0u;
            }
            if (k_tail > 0 && row_live) {
#pragma unroll
                for (int i = 0; i < 32; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * f_old);
#pragma unroll
                for (int t = 0; t < FA_KTAIL; ++t) {
                    if (t >= k_tail) break;
                    const __nv_bfloat16* vrow = tail_v + t * D + c * 32;
#pragma unroll
                    for (int i = 0; i < 32; i += 8) {
                        const uint4 vv = *reinterpret_cast<const uint4*>(vrow + i);
                        const uint32_t vw[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const float2 vf = bf16x2_to_f2(vw[u]);
                            ov[i + 2 * u] = __float_as_uint(fmaf(pt[t], vf.x, __uint_as_float(ov[i + 2 * u])));
                            ov[i + 2 * u + 1] = __float_as_uint(fmaf(pt[t], vf.y, __uint_as_float(ov[i + 2 * u + 1])));
                        }
                    }
                }
            }
            if (qi < p.Sq) {
#pragma unroll
                for (int i = 0; i < 32; i += 8)
                    *reinterpret_cast<uint4*>(dst + c * 32 + i) =
                        make_uint4(pack_bf16x2(__uint_as_float(ov[i]) * inv, __uint_as_float(ov[i + 1]) * inv),
                                   pack_bf16x2(__uint_as_float(ov[i + 2]) * inv, __uint_as_float(ov[i + 3]) * inv),
                                   pack_bf16x2(__uint_as_float(ov[i + 4]) * inv, __uint_as_float(ov[i + 5]) * inv),
                                   pack_bf16x2(__uint_as_float(ov[i + 6]) * inv, __uint_as_float(ov[i + 7]) * inv));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        // re-read from shared memory: keeping the address live across the 48-register warpgroup's branches made ptxas spill it
        tmem_dealloc<512>(*reinterpret_cast<volatile uint32_t*>(tmem_holder));
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
    constexpr int SMEM = FaSmem<D>::BYTES;
    static bool set = false;
    if (!set) {
        if (cudaFuncSetAttribute(attention_fa2q_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM) != cudaSuccess) return GROMA_ERR_CUDA;
        set = true;
    }
    dim3 grid(p.n_main, p.H, p.B);
    attention_fa2q_kernel<D><<<grid, FA_THREADS, SMEM, st>>>(p);
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
    p.q_ptr = reinterpret_cast<const __nv_bfloat16*>(q); p.k_ptr = reinterpret_cast<const __nv_bfloat16*>(k);
    p.v_ptr = reinterpret_cast<const __nv_bfloat16*>(v);
    p.q_ld = q_ld; p.k_ld = k_ld; p.v_ld = v_ld; p.o_batch_rows = Sq;
    static const int tails = [] { const char* e = getenv("GROMA_FA_TAILS"); return e ? atoi(e) : 2; }();   // 0: plain tiling, 1: key tail only (A/B)
    p.k_tail_max = tails ? FA_KTAIL : 0;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    // (Moving the one query row past the last full tile pair -- row 1024 of the ViT's 1025 -- to CUDA-core CTAs of the same grid or
    // to a second launch was measured and bought 2 us of 175, because a separate pass has to pull all of K and V again for one
    // row; profiles/r02_fa_tails.md.  q_tail below reads them from the shared-memory ring instead.)
    p.n_main = (p.Sq + 2 * FA_BM - 1) / (2 * FA_BM);
    p.q_tail = 0;
    // GROMA_FA_TAILS >= 2 (default): ONE query row past the last full pair of tiles (S = 1025) rides on the idle warps of the
    // producer warpgroup instead of costing a fifth CTA per head -- 1024 CTAs are 6.92 waves of 148 SMs, 1280 were 8.65
    if (tails >= 2 && D == 64 && !causal && p.Sq % (2 * FA_BM) == 1 && p.Sq > 2 * FA_BM) {
        p.n_main = p.Sq / (2 * FA_BM);
        p.q_tail = 1;
    }
    return D == 128 ? launch_fa<128>(p, st) : launch_fa<64>(p, st);
}
