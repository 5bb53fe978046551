// One persistent kernel per greedy decode step (reference groma/model/groma.py:376-402 + the HF greedy loop): all 32
// LlamaDecoderLayers, both heads and the argmax in ONE launch of one CTA per SM, replacing the ~294-launch CUDA graph.
//
// Why: the step is HBM-bound (13.2 GB of weights + B*ctx*0.5 MiB of KV per step); with one kernel per GEMM every launch
// pays a ramp and a tail in which the weight stream drains (129 GEMMs x ~3.5 us) and every split-K reduce is a dependent
// launch.  Here the weight / KV stream of a CTA is ONE in-order sequence of 16 KB shared-memory slots filled by a single
// TMA producer lane that never waits for activations: weights and the cached K/V rows do not depend on the step's
// arithmetic, only the tiny activation operand (16 tokens x 64 k = 2 KB per slot) does.  So while a CTA sits in a data
// dependency (split-K reduce, RMSNorm, softmax merge) its ring keeps filling with the NEXT phase's weights and the HBM
// pipe stays busy across what used to be kernel boundaries.
//
// Work decomposition (static, no scheduler):
//   GEMM (swap-AB: 128 weight rows = MMA M, <=16 tokens = MMA N): the (row-tile, k-block) units of a projection form one
//     flat list that is cut into gridDim.x equal contiguous ranges ("stream-K"): every CTA streams the same number of
//     bytes.  A range touches <= a few row tiles; each (CTA, tile) segment accumulates in TMEM and is written as an fp32
//     partial ws[tile][contributor][token][128]; contributors of a tile are consecutive CTAs, summed in CTA order
//     (deterministic).  tcgen05.mma 128x16x16, accumulators in 4 TMEM buffers.
//   attention: items (row b, head h, key segment) round-robin over CTAs; K and V stream through the same ring as 64-key
//     slots; 16 half-warps x 4 keys per slot, fp32 online softmax, P rounded to bf16 (same rounding points as
//     decode_attention_tma_kernel); segments of one (b,h) are merged by whichever CTA finishes last (fixed order).
//   reduces: one CTA per token sums the o / down partials, adds the residual, applies RMSNorm (HF's two roundings);
//     SwiGLU blocks and logit tiles round-robin.
// Dependencies are monotonic counters in global memory (zeroed by the caller before each step): producers of data
// release-increment after their stores, consumers acquire-poll.  Every CTA walks the phases in the same order and only
// ever waits for work of an EARLIER phase, so the schedule cannot deadlock as long as all CTAs are co-resident
// (grid <= #SMs, one CTA per SM by shared-memory footprint).  A watchdog turns a broken invariant into an error code
// instead of a hung GPU.
#include <cstdlib>
#include "ptx.cuh"
#include "decode_common.cuh"
#include "capi_common.h"
#include "groma_b200.h"

namespace gb {

constexpr int MK_STAGES = 12;
constexpr int MK_A_BYTES = 16384;   // 128 x 64 bf16 weight tile | 64 keys x 128 dims of K or of V
constexpr int MK_B_BYTES = 2048;    // 16 tokens x 64 k bf16
constexpr int MK_STAGE_BYTES = MK_A_BYTES + MK_B_BYTES;
constexpr int MK_TOK = 16;          // MMA N = max rows per step
constexpr int MK_WORKERS = 8;       // worker warps: TMEM epilogues (first 4), attention math, reduces
constexpr int MK_WTHREADS = MK_WORKERS * 32;
constexpr int MK_THREADS = 64 + MK_WTHREADS + 64;   // warp 0 activation loader, 1 MMA issuer, 2-9 workers, 10 stream loader, 11 L2 prefetcher
constexpr int MK_TL = 32;           // timeline events per (CTA, role) when MkParams::timeline is set
constexpr int MK_ACC = 4;           // TMEM accumulator buffers of 16 columns
constexpr int MK_MAXC = 8;          // contributors per weight-row tile (checked on the host)
constexpr int MK_D = 128;           // head dim
constexpr int MK_KEYS = 64;         // keys per K / V slot
constexpr int MK_PART = 132;        // floats per attention partial: m, l, pad, pad, acc[128]
constexpr unsigned long long MK_TIMEOUT_NS = 2000000000ull;

struct MkParams {
    CUtensorMap map_w;      // [L*(3Hd+Hd+2I) + V, Hd] bf16: per layer qkv | o | gate/up interleaved, then lm_head|extra
    CUtensorMap map_wd;     // [L*Hd, I] down projections
    CUtensorMap map_yattn;  // [B, Hd] box {64, 16}: input of qkv / head
    CUtensorMap map_ymlp;   // [B, Hd]: input of gate/up
    CUtensorMap map_a;      // [B, Hd]: attention output, input of o
    CUtensorMap map_gu;     // [B, I]: SwiGLU output, input of down
    int L, B, H, Hd, I, V, vocab, S_att;
    int pf_slots;                 // how many 16 KB slots the L2 prefetcher runs ahead of the stream loader (0 = off)
    long long cap;
    float scale_log2, eps;
    const __nv_bfloat16* embed;
    const __nv_bfloat16* new_embed;
    const float* ln_w;            // [2L+1][Hd]: input_layernorm_l, post_attention_layernorm_l, ..., final norm
    __nv_bfloat16* kv;            // [L][2][B][H][cap][128]
    const float* rope_cos;
    const float* rope_sin;
    long long* ids;               // [B] in: token to embed, out: greedy next token
    int* pos;                     // [1] position of this step's token (advanced at the end)
    int* kv_len;                  // [B] keys to attend (incl. this step's), advanced at the end
    __nv_bfloat16 *x, *y_attn, *y_mlp, *a, *gu;
    float* logits;                // [B][V]
    float *ws_qkv, *ws_o, *ws_gu, *ws_down, *ws_head;   // [tiles][MK_MAXC][16][128] fp32
    float* att_part;              // [B*H*S_att][MK_PART]
    float* cand_val;              // [tiles_head][16]
    int* cand_idx;
    int* flags;                   // dependency counters, zero at entry (layout: MkDims)
    int* status;                  // [0] abort code (0 = ok), [1..7] diagnostics of the first watchdog hit, then per (CTA, role)
    unsigned long long* timeline; // optional [grid][4][MK_TL] globaltimer stamps (tools/decode_mega.py --timeline)
};

struct MkDims {
    int tq, to, tg, td, th;       // row tiles of qkv, o, gate/up, down, head
    int kbh, kbi;                 // k-blocks (64) of Hd, I
    int RW;                       // weight-arena rows per layer
    int f_tq, f_to, f_tg, f_td, f_item, f_head, f_gub, lstride;   // counter offsets inside a layer block
    int f_th, f_logits, f_tok;    // after the L layer blocks: head tiles, logits-tiles-done, token_done[(L+1)*2]
};

__device__ __forceinline__ MkDims mk_dims(const MkParams& p) {
    MkDims d;
    d.tq = 3 * p.Hd / 128; d.to = p.Hd / 128; d.tg = 2 * p.I / 128; d.td = p.Hd / 128; d.th = (p.V + 127) / 128;
    d.kbh = p.Hd / 64; d.kbi = p.I / 64;
    d.RW = 4 * p.Hd + 2 * p.I;
    d.f_tq = 0; d.f_to = d.f_tq + d.tq; d.f_tg = d.f_to + d.to; d.f_td = d.f_tg + d.tg;
    d.f_item = d.f_td + d.td; d.f_head = d.f_item + p.B * p.H; d.f_gub = d.f_head + p.H; d.lstride = d.f_gub + d.tg;
    d.f_th = p.L * d.lstride; d.f_logits = d.f_th + d.th; d.f_tok = d.f_logits + 1;
    return d;
}

// ---------------------------------------------------------------------------------------------- small PTX helpers
__device__ __forceinline__ unsigned long long mk_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ int ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ int ld_relaxed(const int* p) {
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_add(int* p, int v) {
    asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}\n"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_expect_tx_only(uint64_t* bar, uint32_t bytes) {   // tx-count += bytes, no arrival
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* d, int32_t c0, int32_t c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(d)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_prefetch_1d(const void* gptr, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gptr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void bar_workers() { asm volatile("bar.sync 1, %0;" ::"n"(MK_WTHREADS) : "memory"); }
__device__ __forceinline__ void bar_epilogue() { asm volatile("bar.sync 2, 128;" ::: "memory"); }

// Watchdog: a wait that outlives MK_TIMEOUT_NS records who / where and raises the abort flag; every wait loop polls the
// flag, so the whole grid drains within one more timeout instead of hanging the GPU.
struct MkCtx {
    const MkParams* p;
    volatile int* abort_s;   // shared-memory mirror of the abort state of this CTA
    int role;                // 0 stream loader, 1 mma, 2 worker, 3 activation loader
};
__device__ __forceinline__ void mk_tl(const MkCtx& c, int idx) {
    if (c.p->timeline != nullptr && idx < MK_TL) c.p->timeline[((long long)blockIdx.x * 4 + c.role) * MK_TL + idx] = mk_now();
}
// roles: 0 stream loader (weights / KV), 1 MMA issuer, 2 workers, 3 activation loader
// status layout: [0] first abort code, [1..6] its details, [8 + (cta*4 + role)*4 ..] = (code, info, aux, thread) of the wait each
// (CTA, role) was parked in when the abort reached it -- enough to see which dependency never arrived.
__device__ __noinline__ void mk_note(const MkCtx& c, int code, long long a, long long b) {
    int* slot = c.p->status + 8 + ((int)blockIdx.x * 4 + c.role) * 4;
    if (atomicCAS(slot, 0, code) == 0) { slot[1] = (int)a; slot[2] = (int)b; slot[3] = threadIdx.x; }
}
__device__ __noinline__ void mk_fail(const MkCtx& c, int code, long long a, long long b) {
    if (atomicCAS(c.p->status, 0, code) == 0) {
        c.p->status[1] = blockIdx.x; c.p->status[2] = c.role; c.p->status[3] = threadIdx.x;
        c.p->status[4] = (int)a; c.p->status[5] = (int)b; c.p->status[6] = (int)(a >> 32);
        __threadfence();
    }
    mk_note(c, code, a, b);
    *c.abort_s = 1;
}
__device__ __forceinline__ bool mk_aborted(const MkCtx& c) { return *c.abort_s != 0; }
// returns false when the wait was abandoned
// mbarrier.try_wait suspends the thread in hardware until the phase completes or a time limit passes -- unlike a test_wait spin it
// does not keep hammering the shared-memory pipeline the TMA writes and the UMMA operand reads go through (128-256 worker
// threads parked on a barrier for a whole GEMM phase did exactly that in the first version).
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2, 0x989680;\n\t"     // suspend-time hint: 10 ms
        "selp.u32 %0, 1, 0, P;\n\t}\n"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mk_wait_mbar(const MkCtx& c, uint64_t* bar, uint32_t parity, int code, long long info) {
    if (mbar_test(bar, parity)) return true;
    if (mk_aborted(c)) return false;
    const unsigned long long t0 = mk_now();
    while (!mbar_try(bar, parity)) {
        if (mk_aborted(c) || ld_relaxed(c.p->status) != 0) { mk_note(c, code, info, parity); *c.abort_s = 1; return false; }
        if (mk_now() - t0 > MK_TIMEOUT_NS) { mk_fail(c, code, info, parity); return false; }
    }
    return true;
}
// warp-collective form: lane 0 waits, the warp re-converges, every lane then observes the (already completed) phase itself so that
// the asynchronous-proxy writes the barrier tracks are visible to it
__device__ __forceinline__ bool mk_wait_mbar_warp(const MkCtx& c, uint64_t* bar, uint32_t parity, int code, long long info) {
    bool ok = true;
    if ((threadIdx.x & 31) == 0) ok = mk_wait_mbar(c, bar, parity, code, info);
    ok = __shfl_sync(0xffffffffu, ok ? 1 : 0, 0) != 0;
    if (ok) ok = mbar_test(bar, parity);
    return ok;
}
__device__ __forceinline__ bool mk_wait_flag(const MkCtx& c, const int* flag, int need, int code, long long info) {
    if (ld_acquire(flag) >= need) return true;
    if (mk_aborted(c)) return false;
    const unsigned long long t0 = mk_now();
    int spins = 0;
    while (ld_acquire(flag) < need) {
        __nanosleep(20);
        if ((++spins & 63) == 0) {
            if (mk_aborted(c) || ld_relaxed(c.p->status) != 0) { mk_note(c, code, info, need); *c.abort_s = 1; return false; }
            if (mk_now() - t0 > MK_TIMEOUT_NS) { mk_fail(c, code, info, need); return false; }
        }
    }
    return true;
}

// ---------------------------------------------------------------------------------------------- static schedule
struct MkGemm {
    int which;      // 0 qkv, 1 o, 2 gate/up, 3 down, 4 head
    int tiles, kb;  // row tiles, k-blocks
    int row0;       // first weight row in its tensor map
    int u0, u1;     // this CTA's unit range
};
__device__ __forceinline__ MkGemm mk_gemm(const MkParams& p, const MkDims& d, int layer, int which) {
    MkGemm g;
    g.which = which;
    switch (which) {
        case 0: g.tiles = d.tq; g.kb = d.kbh; g.row0 = layer * d.RW; break;
        case 1: g.tiles = d.to; g.kb = d.kbh; g.row0 = layer * d.RW + 3 * p.Hd; break;
        case 2: g.tiles = d.tg; g.kb = d.kbh; g.row0 = layer * d.RW + 4 * p.Hd; break;
        case 3: g.tiles = d.td; g.kb = d.kbi; g.row0 = layer * p.Hd; break;
        default: g.tiles = d.th; g.kb = d.kbh; g.row0 = p.L * d.RW; break;
    }
    const long long U = (long long)g.tiles * g.kb;
    g.u0 = (int)((long long)blockIdx.x * U / gridDim.x);
    g.u1 = (int)(((long long)blockIdx.x + 1) * U / gridDim.x);
    return g;
}
// CTA owning unit u of a list of U units cut into G equal ranges [c*U/G, (c+1)*U/G)
__device__ __forceinline__ int mk_owner(long long u, long long U, int G) { return (int)(((u + 1) * G - 1) / U); }
__device__ __forceinline__ void mk_contrib(int tile, int kb, int tiles, int G, int& first, int& n) {
    const long long U = (long long)tiles * kb;
    first = mk_owner((long long)tile * kb, U, G);
    n = mk_owner((long long)(tile + 1) * kb - 1, U, G) - first + 1;
}

__device__ __forceinline__ int g_tiles_gu(const MkDims& d) { return d.tg; }

// key range of attention item i = seg * (B*H) + (b*H + h): segment-major, so the round-robin i -> CTA mixes long and short
// segments on every CTA; the 64-key chunks of a row are dealt to its S segments as evenly as possible
struct MkItem { int b, h, seg, k_begin, n, nch; };
__device__ __forceinline__ MkItem mk_item(const MkParams& p, const int* kvlen_s, int i) {
    MkItem it;
    const int BH = p.B * p.H;
    it.seg = i / BH;
    const int bh = i - it.seg * BH;
    it.b = bh / p.H; it.h = bh - it.b * p.H;
    const int n_all = kvlen_s[it.b];
    const int C = (n_all + MK_KEYS - 1) / MK_KEYS;
    const int c0 = it.seg * C / p.S_att, c1 = (it.seg + 1) * C / p.S_att;
    it.k_begin = min(c0 * MK_KEYS, n_all);
    it.n = min(c1 * MK_KEYS, n_all) - it.k_begin;
    it.nch = c1 - c0;
    return it;
}
__device__ __forceinline__ int mk_kv_slots(const MkParams& p, const int* kvlen_s) {
    int n = 0;
    const int NI = p.B * p.H * p.S_att;
    for (int i = blockIdx.x; i < NI; i += gridDim.x) n += 2 * mk_item(p, kvlen_s, i).nch;
    return n;
}

// ---------------------------------------------------------------------------------------------- kernel
struct MkSmemTail {
    alignas(16) float st_acc[16][MK_D];
    alignas(16) __nv_bfloat16 s_new[3][MK_D];   // read back as uint4
    alignas(8) uint64_t full_bar[MK_STAGES];
    uint64_t empty_bar[MK_STAGES];
    uint64_t tfull_bar[MK_ACC];
    uint64_t tempty_bar[MK_ACC];
    uint32_t tmem_holder[2];
    int abort_flag;
    int last_flag;
    volatile int seq_a;          // slots whose weight / KV load has been issued (stream loader -> activation loader)
    volatile int attn_done;      // layers whose attention items this CTA's workers have finished (workers -> MMA issuer)
    int kv_rel[MK_STAGES];
    int kvlen[MK_TOK];
    float red[MK_WORKERS];
    float st_m[16], st_l[16];
};
static_assert(sizeof(MkSmemTail) % 16 == 0, "tail layout");
constexpr int MK_SMEM_BYTES = MK_STAGES * MK_STAGE_BYTES + 1024 + (int)sizeof(MkSmemTail);

__device__ __forceinline__ float mk_block_sum(float v, float* red) {   // over the MK_WTHREADS worker threads
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = (threadIdx.x >> 5) - 2, l = threadIdx.x & 31;
    bar_workers();
    if (l == 0) red[w] = v;
    bar_workers();
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < MK_WORKERS; ++i) t += red[i];
    return t;
}

// sum of the contributions to feature `row` (0..127) of `tile` for token `tok`, in contributor order
__device__ __forceinline__ float mk_sum1(const float* ws, int tile, int first_n, int tok, int row) {
    float a = 0.f;
    const float* src = ws + (((long long)tile * MK_MAXC) * MK_TOK + tok) * 128 + row;
    for (int s = 0; s < first_n; ++s) a += __ldcg(src + (long long)s * MK_TOK * 128);
    return a;
}

__global__ void __launch_bounds__(MK_THREADS, 1) decode_step_megakernel(const __grid_constant__ MkParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    MkSmemTail& T = *reinterpret_cast<MkSmemTail*>(smem + MK_STAGES * MK_STAGE_BYTES);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const MkDims d = mk_dims(p);
    const int G = gridDim.x, cta = blockIdx.x;

    if (threadIdx.x == 0) {
        for (int i = 0; i < MK_STAGES; ++i) { mbar_init(&T.full_bar[i], 1); mbar_init(&T.empty_bar[i], 1); T.kv_rel[i] = 0; }
        for (int i = 0; i < MK_ACC; ++i) { mbar_init(&T.tfull_bar[i], 1); mbar_init(&T.tempty_bar[i], 4); }
        T.abort_flag = 0; T.last_flag = 0; T.seq_a = 0; T.attn_done = 0;
        fence_barrier_init();
        tma_prefetch_desc(&p.map_w); tma_prefetch_desc(&p.map_wd); tma_prefetch_desc(&p.map_yattn);
        tma_prefetch_desc(&p.map_ymlp); tma_prefetch_desc(&p.map_a); tma_prefetch_desc(&p.map_gu);
    }
    if (threadIdx.x < MK_TOK) T.kvlen[threadIdx.x] = (threadIdx.x < p.B) ? p.kv_len[threadIdx.x] : 0;
    if (warp == 1) tmem_alloc<64>(T.tmem_holder);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = T.tmem_holder[0];
    const int pos = *p.pos;
    const int nkv = mk_kv_slots(p, T.kvlen);      // K/V slots of this CTA per layer
    MkCtx ctx{&p, &T.abort_flag, warp == 0 ? 3 : (warp == 1 ? 1 : (warp >= 10 ? 0 : 2))};

    if (warp == 10) {
        // =========================================================== stream loader: weights and cached K/V, never waits for data.
        // One lane, a handful of instructions per 16 KB slot (no divisions in the loops): it has to sustain one slot per ~0.3 us.
        if (lane == 0) {
            int seq = 0;
            bool ok = true;
            mk_tl(ctx, 0);
            auto slot_ptr = [&](int sq) -> uint8_t* { return smem + (sq % MK_STAGES) * MK_STAGE_BYTES; };
            for (int layer = 0; layer <= p.L && ok; ++layer) {
                for (int ph = 0; ph < 5 && ok; ++ph) {
                    if (layer == p.L && ph > 0) break;
                    if (ph == 1) {
                        const int NI = p.B * p.H * p.S_att;
                        const long long plane = (long long)p.B * p.H * p.cap * MK_D;       // elements of one layer's K (or V)
                        const __nv_bfloat16* kl = p.kv + (long long)layer * 2 * plane;
                        for (int i = cta; i < NI && ok; i += G) {
                            const MkItem it = mk_item(p, T.kvlen, i);
                            const __nv_bfloat16* kb = kl + (((long long)it.b * p.H + it.h) * p.cap + it.k_begin) * MK_D;
                            int left = it.n;
                            for (int ch = 0; ch < it.nch; ++ch, left -= MK_KEYS, kb += MK_KEYS * MK_D) {
                                const uint32_t bytes = (uint32_t)min(MK_KEYS, left) * MK_D * 2;
#pragma unroll
                                for (int kv = 0; kv < 2; ++kv) {
                                    const int s = seq % MK_STAGES;
                                    if (!mk_wait_mbar(ctx, &T.empty_bar[s], (uint32_t)(((seq / MK_STAGES) & 1) ^ 1), 101, seq)) { ok = false; break; }
                                    mbar_expect_tx(&T.full_bar[s], bytes);
                                    bulk_load_1d(slot_ptr(seq), kv ? kb + plane : kb, bytes, &T.full_bar[s]);
                                    T.seq_a = ++seq;
                                }
                                if (!ok) break;
                            }
                        }
                        continue;
                    }
                    const MkGemm g = mk_gemm(p, d, layer, layer == p.L ? 4 : (ph == 0 ? 0 : ph - 1));
                    const CUtensorMap* amap = (g.which == 3) ? &p.map_wd : &p.map_w;
                    int tile = g.u0 / g.kb, k = g.u0 - tile * g.kb;
                    for (int u = g.u0; u < g.u1; ++u) {
                        const int s = seq % MK_STAGES;
                        if (!mk_wait_mbar(ctx, &T.empty_bar[s], (uint32_t)(((seq / MK_STAGES) & 1) ^ 1), 101, seq)) { ok = false; break; }
                        mbar_expect_tx_only(&T.full_bar[s], MK_A_BYTES);
                        tma_load_2d(slot_ptr(seq), amap, &T.full_bar[s], k * 64, g.row0 + tile * 128);
                        T.seq_a = ++seq;       // after the expect_tx above: the activation loader may now arrive on this slot's barrier
                        if (++k == g.kb) { k = 0; ++tile; }
                    }
                }
            }
            mk_tl(ctx, 1);
        }
    } else if (warp == 11) {
        // =========================================================== L2 prefetcher: the same slot list as the stream loader, pf_slots
        // ahead of it.  HBM latency under load (~3 us) x 7 TB/s is more than the 200 KB a ring can keep in flight per SM; with the
        // tiles already in L2 the ring only has to cover L2 latency.
        if (lane == 0 && p.pf_slots > 0) {
            int seq = 0;
            bool ok = true;
            auto throttle = [&]() {
                if (seq - T.seq_a < p.pf_slots) return;
                int spins = 0;
                while (seq - T.seq_a >= p.pf_slots) {
                    if ((++spins & 1023) == 0 && (mk_aborted(ctx) || ld_relaxed(p.status) != 0)) { ok = false; return; }
                }
            };
            for (int layer = 0; layer <= p.L && ok; ++layer) {
                for (int ph = 0; ph < 5 && ok; ++ph) {
                    if (layer == p.L && ph > 0) break;
                    if (ph == 1) {
                        const int NI = p.B * p.H * p.S_att;
                        const long long plane = (long long)p.B * p.H * p.cap * MK_D;
                        const __nv_bfloat16* kl = p.kv + (long long)layer * 2 * plane;
                        for (int i = cta; i < NI && ok; i += G) {
                            const MkItem it = mk_item(p, T.kvlen, i);
                            const __nv_bfloat16* kb = kl + (((long long)it.b * p.H + it.h) * p.cap + it.k_begin) * MK_D;
                            int left = it.n;
                            for (int ch = 0; ch < it.nch && ok; ++ch, left -= MK_KEYS, kb += MK_KEYS * MK_D) {
                                const uint32_t bytes = (uint32_t)min(MK_KEYS, left) * MK_D * 2;
                                throttle();
                                bulk_prefetch_1d(kb, bytes);
                                bulk_prefetch_1d(kb + plane, bytes);
                                seq += 2;
                            }
                        }
                        continue;
                    }
                    const MkGemm g = mk_gemm(p, d, layer, layer == p.L ? 4 : (ph == 0 ? 0 : ph - 1));
                    const CUtensorMap* amap = (g.which == 3) ? &p.map_wd : &p.map_w;
                    int tile = g.u0 / g.kb, k = g.u0 - tile * g.kb;
                    for (int u = g.u0; u < g.u1 && ok; ++u, ++seq) {
                        throttle();
                        tma_prefetch_2d(amap, k * 64, g.row0 + tile * 128);
                        if (++k == g.kb) { k = 0; ++tile; }
                    }
                }
            }
        }
    } else if (warp == 0) {
        // =========================================================== activation loader: the 2 KB B operand of every GEMM slot, as soon as
        // the phase that produces it has published it (one counter per dependency, polled once per phase)
        if (lane == 0) {
            int seq = 0;
            bool ok = true;
            for (int layer = 0; layer <= p.L && ok; ++layer) {
                int* fl = p.flags + layer * d.lstride;
                for (int ph = 0; ph < 5 && ok; ++ph) {
                    if (layer == p.L && ph > 0) break;
                    if (ph == 1) { seq += nkv; continue; }
                    const MkGemm g = mk_gemm(p, d, layer, layer == p.L ? 4 : (ph == 0 ? 0 : ph - 1));
                    const CUtensorMap* bmap; const int* flag; int need;
                    switch (g.which) {
                        case 0: bmap = &p.map_yattn; flag = p.flags + d.f_tok + layer * 2; need = p.B; break;
                        case 1: bmap = &p.map_a; flag = fl + d.f_head; need = p.B * p.H; break;
                        case 2: bmap = &p.map_ymlp; flag = p.flags + d.f_tok + layer * 2 + 1; need = p.B; break;
                        case 3: bmap = &p.map_gu; flag = fl + d.f_gub; need = g_tiles_gu(d); break;
                        default: bmap = &p.map_yattn; flag = p.flags + d.f_tok + p.L * 2; need = p.B; break;
                    }
                    if (g.u0 < g.u1) {
                        ok = mk_wait_flag(ctx, flag, need, 102, seq);
                        if (!ok) break;
                        fence_proxy_async_all();
                    }
                    int k = g.u0 % g.kb;
                    for (int u = g.u0; u < g.u1; ++u, ++seq) {
                        if (T.seq_a <= seq) {      // the slot's weight load (and its expect_tx) must have been issued first
                            const unsigned long long t0 = mk_now();
                            int spins = 0;
                            while (T.seq_a <= seq) {
                                if ((++spins & 255) == 0) {
                                    if (mk_aborted(ctx) || ld_relaxed(p.status) != 0) { mk_note(ctx, 103, seq, T.seq_a); T.abort_flag = 1; ok = false; break; }
                                    if (mk_now() - t0 > MK_TIMEOUT_NS) { mk_fail(ctx, 103, seq, T.seq_a); ok = false; break; }
                                }
                            }
                            if (!ok) break;
                        }
                        const int s = seq % MK_STAGES;
                        mbar_expect_tx(&T.full_bar[s], MK_B_BYTES);
                        tma_load_2d(smem + s * MK_STAGE_BYTES + MK_A_BYTES, bmap, &T.full_bar[s], k * 64, 0);
                        if (++k == g.kb) k = 0;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // =========================================================== MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(128, MK_TOK);
            int seq = 0;
            int acc = 0; uint32_t acc_phase = 0;
            bool ok = true;
            for (int layer = 0; layer <= p.L && ok; ++layer) {
                for (int ph = 0; ph < 5 && ok; ++ph) {
                    if (layer == p.L && ph > 0) break;
                    if (ph == 1) {
                        // The K/V slots of this phase are consumed by the workers.  The issuer must not look at the full barrier of
                        // a LATER use of a ring slot before the earlier use has gone through (parity waits cannot tell phases two
                        // rounds apart): wait until the workers have drained this layer's attention slots.
                        seq += nkv;
                        if (T.attn_done <= layer) {
                            const unsigned long long t0 = mk_now();
                            int spins = 0;
                            while (T.attn_done <= layer) {
                                if ((++spins & 255) == 0) {
                                    if (mk_aborted(ctx) || ld_relaxed(p.status) != 0) { mk_note(ctx, 203, layer, seq); T.abort_flag = 1; ok = false; break; }
                                    if (mk_now() - t0 > MK_TIMEOUT_NS) { mk_fail(ctx, 203, layer, seq); ok = false; break; }
                                }
                            }
                        }
                        continue;
                    }
                    const MkGemm g = mk_gemm(p, d, layer, layer == p.L ? 4 : (ph == 0 ? 0 : ph - 1));
                    if (layer == 0 || layer == p.L) mk_tl(ctx, layer == 0 ? ph * 2 : 10);
                    int u = g.u0;
                    while (u < g.u1 && ok) {
                        const int tile = u / g.kb;
                        const int uend = min(g.u1, (tile + 1) * g.kb);
                        ok = mk_wait_mbar(ctx, &T.tempty_bar[acc], acc_phase ^ 1, 201, seq);
                        if (!ok) break;
                        tc_fence_after();
                        const uint32_t d_tmem = tmem_base + acc * MK_TOK;
                        const int ustart = u;
                        for (; u < uend; ++u, ++seq) {
                            const int s = seq % MK_STAGES;
                            ok = mk_wait_mbar(ctx, &T.full_bar[s], (uint32_t)((seq / MK_STAGES) & 1), 202, seq);
                            if (!ok) break;
                            tc_fence_after();
                            const uint32_t sa = smem_u32(smem + s * MK_STAGE_BYTES);
                            const uint64_t da = make_sw128_kmajor_desc(sa);
                            const uint64_t db = make_sw128_kmajor_desc(sa + MK_A_BYTES);
#pragma unroll
                            for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, da + 2 * k, db + 2 * k, idesc, (u > ustart || k > 0) ? 1u : 0u);
                            umma_commit(&T.empty_bar[s]);
                        }
                        if (!ok) break;
                        umma_commit(&T.tfull_bar[acc]);
                        if (++acc == MK_ACC) { acc = 0; acc_phase ^= 1; }
                    }
                    if (layer == 0 || layer == p.L) mk_tl(ctx, layer == 0 ? ph * 2 + 1 : 11);
                }
            }
        }
    } else {
        // =========================================================== workers
        const int w = warp - 2;                 // 0..7
        const int tid = threadIdx.x - 64;       // 0..255
        const int grp = lane >> 4, l16 = lane & 15;
        const int hw = w * 2 + grp;             // half-warp id 0..15
        const int q4 = warp & 3;                // TMEM lane quarter of this warp
        int seq = 0;
        int acc = 0; uint32_t acc_phase = 0;
        bool ok = true;
        int* tokdone = p.flags + d.f_tok;

        // Abort discipline: worker warps never leave the static schedule early -- every named barrier below is reached by all
        // 256 worker threads even after a watchdog hit (waits return at once, `ok` gates global side effects), so a broken
        // invariant ends in an error code, not in a CTA parked at bar.sync.
        auto sync_ok = [&]() { bar_workers(); ok = ok && !mk_aborted(ctx); };

        // ---- split-K contributions -> global partials (first four worker warps), one call per projection
        auto epilogue = [&](const MkGemm& g, float* ws, int* tile_done) {
            int u = g.u0;
            while (u < g.u1) {
                const int tile = u / g.kb;
                const int uend = min(g.u1, (tile + 1) * g.kb);
                if (w < 4) {
                    bool okw = mk_wait_mbar_warp(ctx, &T.tfull_bar[acc], acc_phase, 301, tile) && ok;
                    okw = __all_sync(0xffffffffu, okw);
                    if (okw) {
                        tc_fence_after();
                        uint32_t v[32];
                        tmem_ld16(tmem_base + (uint32_t(q4 * 32) << 16) + uint32_t(acc * MK_TOK), v);
                        tmem_ld_wait();
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&T.tempty_bar[acc]);
                        int first, n;
                        mk_contrib(tile, g.kb, g.tiles, G, first, n);
                        float* dst = ws + (((long long)tile * MK_MAXC + (cta - first)) * MK_TOK) * 128 + q4 * 32 + lane;
#pragma unroll
                        for (int j = 0; j < MK_TOK; ++j) __stcg(dst + j * 128, __uint_as_float(v[j]));
                    } else {
                        ok = false;
                    }
                    bar_epilogue();      // CTA-scope order of the 128 threads' stores before the one gpu-scope release below (cumulative)
                    if (okw && w == 0 && lane == 0) red_release_add(tile_done + tile, 1);
                }
                seq += uend - u;
                u = uend;
                if (++acc == MK_ACC) { acc = 0; acc_phase ^= 1; }
            }
        };
        // ---- y = w * bf16(x * rsqrt(mean(x^2)+eps)) for token `cta`; x = bf16(x + sum of partials) (or the embedding row)
        auto reduce_norm = [&](const float* ws, const int* tile_done, int tiles, int kb, const float* lnw, __nv_bfloat16* y, int* done_flag) {
            const int tok = cta;
            float h[8][4];
            float ss = 0.f;
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int f = (it * MK_WTHREADS + tid) * 4;
                h[it][0] = h[it][1] = h[it][2] = h[it][3] = 0.f;
                if (f < p.Hd && ok) {
                    __nv_bfloat16* xp = p.x + (long long)tok * p.Hd + f;
                    float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
                    if (ws != nullptr) {
                        const int tile = f >> 7;
                        int first, n;
                        mk_contrib(tile, kb, tiles, G, first, n);
                        ok = mk_wait_flag(ctx, tile_done + tile, n, 401, tile);
                        if (ok) {
                            const float* src = ws + (((long long)tile * MK_MAXC) * MK_TOK + tok) * 128 + (f & 127);
                            for (int s = 0; s < n; ++s) {
                                const float4 t = __ldcg(reinterpret_cast<const float4*>(src + (long long)s * MK_TOK * 128));
                                v0 += t.x; v1 += t.y; v2 += t.z; v3 += t.w;
                            }
                            const uint2 xr = *reinterpret_cast<const uint2*>(xp);
                            const float2 x01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&xr.x));
                            const float2 x23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&xr.y));
                            v0 = bf16_round(v0 + x01.x); v1 = bf16_round(v1 + x01.y); v2 = bf16_round(v2 + x23.x); v3 = bf16_round(v3 + x23.y);
                        }
                    } else {
                        const long long id = p.ids[tok];
                        const __nv_bfloat16* er = (id < p.vocab) ? p.embed + id * p.Hd : p.new_embed + (id - p.vocab) * p.Hd;
                        const uint2 xr = *reinterpret_cast<const uint2*>(er + f);
                        const float2 x01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&xr.x));
                        const float2 x23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&xr.y));
                        v0 = x01.x; v1 = x01.y; v2 = x23.x; v3 = x23.y;
                    }
                    if (ok) *reinterpret_cast<uint2*>(xp) = make_uint2(pack_bf16x2(v0, v1), pack_bf16x2(v2, v3));
                    h[it][0] = v0; h[it][1] = v1; h[it][2] = v2; h[it][3] = v3;
                    ss += v0 * v0 + v1 * v1 + v2 * v2 + v3 * v3;
                }
            }
            ss = mk_block_sum(ss, T.red);
            ok = ok && !mk_aborted(ctx);
            const float rs = rsqrtf(ss / p.Hd + p.eps);
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int f = (it * MK_WTHREADS + tid) * 4;
                if (f < p.Hd && ok) {
                    const float4 w4 = *reinterpret_cast<const float4*>(lnw + f);
                    *reinterpret_cast<uint2*>(y + (long long)tok * p.Hd + f) =
                        make_uint2(pack_bf16x2(w4.x * bf16_round(h[it][0] * rs), w4.y * bf16_round(h[it][1] * rs)),
                                   pack_bf16x2(w4.z * bf16_round(h[it][2] * rs), w4.w * bf16_round(h[it][3] * rs)));
                }
            }
            sync_ok();
            if (ok && tid == 0) { fence_proxy_async_all(); red_release_add(done_flag, 1); }
        };

        // ---- step start: embedding row + input RMSNorm of layer 0
        if (cta < p.B) reduce_norm(nullptr, nullptr, 0, 0, p.ln_w, p.y_attn, tokdone);

        for (int layer = 0; layer < p.L; ++layer) {
            int* fl = p.flags + layer * d.lstride;
            // ---------------- qkv
            const MkGemm gq = mk_gemm(p, d, layer, 0);
            if (tid == 0 && layer == 0) mk_tl(ctx, 0);
            epilogue(gq, p.ws_qkv, fl + d.f_tq);
            if (tid == 0 && layer == 0) mk_tl(ctx, 1);
            // ---------------- attention items
            {
                const int NI = p.B * p.H * p.S_att;
                __nv_bfloat16* kc = p.kv + ((long long)layer * 2) * p.B * p.H * p.cap * MK_D;
                __nv_bfloat16* vc = kc + (long long)p.B * p.H * p.cap * MK_D;
                for (int i = cta; i < NI; i += G) {
                    const MkItem it = mk_item(p, T.kvlen, i);
                    const int n_all = T.kvlen[it.b];
                    int newidx = -1;
                    if (pos >= it.k_begin && pos < it.k_begin + it.n) newidx = pos - it.k_begin;
                    const bool append = newidx >= 0 || (pos >= n_all && it.seg == p.S_att - 1);
                    // q (and, for the segment that owns this step's position, the new K / V row) from the qkv partials
                    if (tid < 64 && ok) {
                        const int j = tid;
                        int f0, n0;
                        mk_contrib(it.h, gq.kb, gq.tiles, G, f0, n0);
                        ok = mk_wait_flag(ctx, fl + d.f_tq + it.h, n0, 501, it.h);
                        if (ok) {
                            const float c = p.rope_cos[(long long)pos * 64 + j], sn = p.rope_sin[(long long)pos * 64 + j];
                            float q1 = bf16_round(mk_sum1(p.ws_qkv, it.h, n0, it.b, j));
                            float q2 = bf16_round(mk_sum1(p.ws_qkv, it.h, n0, it.b, j + 64));
                            rope_pair(q1, q2, c, sn, T.s_new[0][j], T.s_new[0][j + 64]);
                            if (append) {
                                int f1, n1, f2, n2;
                                mk_contrib(p.H + it.h, gq.kb, gq.tiles, G, f1, n1);
                                mk_contrib(2 * p.H + it.h, gq.kb, gq.tiles, G, f2, n2);
                                ok = mk_wait_flag(ctx, fl + d.f_tq + p.H + it.h, n1, 502, it.h) &&
                                     mk_wait_flag(ctx, fl + d.f_tq + 2 * p.H + it.h, n2, 503, it.h);
                                if (ok) {
                                    float k1 = bf16_round(mk_sum1(p.ws_qkv, p.H + it.h, n1, it.b, j));
                                    float k2 = bf16_round(mk_sum1(p.ws_qkv, p.H + it.h, n1, it.b, j + 64));
                                    rope_pair(k1, k2, c, sn, T.s_new[1][j], T.s_new[1][j + 64]);
                                    T.s_new[2][j] = __float2bfloat16_rn(mk_sum1(p.ws_qkv, 2 * p.H + it.h, n2, it.b, j));
                                    T.s_new[2][j + 64] = __float2bfloat16_rn(mk_sum1(p.ws_qkv, 2 * p.H + it.h, n2, it.b, j + 64));
                                    const long long ro = (((long long)it.b * p.H + it.h) * p.cap + pos) * MK_D;
                                    kc[ro + j] = T.s_new[1][j]; kc[ro + j + 64] = T.s_new[1][j + 64];
                                    vc[ro + j] = T.s_new[2][j]; vc[ro + j + 64] = T.s_new[2][j + 64];
                                }
                            }
                        }
                    }
                    sync_ok();
                    float qf[8];
                    {
                        const uint4 qv = *reinterpret_cast<const uint4*>(&T.s_new[0][l16 * 8]);
                        const __nv_bfloat162* q2 = reinterpret_cast<const __nv_bfloat162*>(&qv);
#pragma unroll
                        for (int t = 0; t < 4; ++t) {
                            const float2 f = __bfloat1622float2(q2[t]);
                            qf[2 * t] = f.x * p.scale_log2;
                            qf[2 * t + 1] = f.y * p.scale_log2;
                        }
                    }
                    float m = -INFINITY, lsum = 0.f, accv[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                    for (int ch = 0; ch < it.nch; ++ch) {
                        const int left = it.n - ch * MK_KEYS;
                        float sc[4];
                        // ---- K slot: 4 keys per half-warp (key j = hw + 16*u)
                        {
                            const int s = seq % MK_STAGES;
                            ok = mk_wait_mbar_warp(ctx, &T.full_bar[s], (uint32_t)((seq / MK_STAGES) & 1), 504, seq) && ok;
                            const __nv_bfloat16* ks = reinterpret_cast<const __nv_bfloat16*>(smem + s * MK_STAGE_BYTES);
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const int j = hw + 16 * u;
                                const bool valid = j < left;
                                uint4 kk = valid ? *reinterpret_cast<const uint4*>(ks + j * MK_D + l16 * 8) : make_uint4(0, 0, 0, 0);
                                if (ch * MK_KEYS + j == newidx) kk = *reinterpret_cast<const uint4*>(&T.s_new[1][l16 * 8]);
                                const __nv_bfloat162* k2 = reinterpret_cast<const __nv_bfloat162*>(&kk);
                                float dt = 0.f;
#pragma unroll
                                for (int t = 0; t < 4; ++t) {
                                    const float2 f = __bfloat1622float2(k2[t]);
                                    dt += f.x * qf[2 * t] + f.y * qf[2 * t + 1];
                                }
                                dt += __shfl_xor_sync(0xffffffffu, dt, 8);
                                dt += __shfl_xor_sync(0xffffffffu, dt, 4);
                                dt += __shfl_xor_sync(0xffffffffu, dt, 2);
                                dt += __shfl_xor_sync(0xffffffffu, dt, 1);
                                sc[u] = valid ? dt : -INFINITY;
                            }
                            __syncwarp();
                            if (lane == 0 && ok && atomicAdd(&T.kv_rel[s], 1) == MK_WORKERS - 1) { T.kv_rel[s] = 0; mbar_arrive(&T.empty_bar[s]); }
                            ++seq;
                        }
                        // ---- V slot
                        {
                            const int s = seq % MK_STAGES;
                            ok = mk_wait_mbar_warp(ctx, &T.full_bar[s], (uint32_t)((seq / MK_STAGES) & 1), 505, seq) && ok;
                            const __nv_bfloat16* vs = reinterpret_cast<const __nv_bfloat16*>(smem + s * MK_STAGE_BYTES);
                            uint4 vv[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const int j = hw + 16 * u;
                                vv[u] = (j < left) ? *reinterpret_cast<const uint4*>(vs + j * MK_D + l16 * 8) : make_uint4(0, 0, 0, 0);
                                if (ch * MK_KEYS + j == newidx) vv[u] = *reinterpret_cast<const uint4*>(&T.s_new[2][l16 * 8]);
                            }
                            __syncwarp();
                            if (lane == 0 && ok && atomicAdd(&T.kv_rel[s], 1) == MK_WORKERS - 1) { T.kv_rel[s] = 0; mbar_arrive(&T.empty_bar[s]); }
                            ++seq;
                            float mn = m;
#pragma unroll
                            for (int u = 0; u < 4; ++u) mn = fmaxf(mn, sc[u]);
                            const float mref = (mn == -INFINITY) ? 0.f : mn;
                            const float corr = exp2f(m - mref);
                            m = mn;
                            lsum *= corr;
#pragma unroll
                            for (int t = 0; t < 8; ++t) accv[t] *= corr;
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const float pe = exp2f(sc[u] - mref);
                                lsum += pe;
                                const float pr = bf16_round(pe);
                                const __nv_bfloat162* v2 = reinterpret_cast<const __nv_bfloat162*>(&vv[u]);
#pragma unroll
                                for (int t = 0; t < 4; ++t) {
                                    const float2 f = __bfloat1622float2(v2[t]);
                                    accv[2 * t] += pr * f.x;
                                    accv[2 * t + 1] += pr * f.y;
                                }
                            }
                        }
                    }
                    // ---- merge the 16 half-warp states (fixed order)
                    if (l16 == 0) { T.st_m[hw] = m; T.st_l[hw] = lsum; }
#pragma unroll
                    for (int t = 0; t < 8; ++t) T.st_acc[hw][l16 * 8 + t] = accv[t];
                    sync_ok();
                    float M = -INFINITY, num = 0.f, den = 0.f;
                    if (tid < MK_D) {
#pragma unroll
                        for (int q = 0; q < 16; ++q) M = fmaxf(M, T.st_m[q]);
#pragma unroll
                        for (int q = 0; q < 16; ++q) {
                            const float c = (T.st_m[q] == -INFINITY) ? 0.f : exp2f(T.st_m[q] - M);
                            num += c * T.st_acc[q][tid];
                            den += c * T.st_l[q];
                        }
                    }
                    const long long bh = (long long)it.b * p.H + it.h;
                    __nv_bfloat16* aout = p.a + bh * MK_D;
                    bool finish = true;          // this CTA writes the final row
                    if (p.S_att > 1) {
                        if (tid < MK_D && ok) {
                            float* pp = p.att_part + ((long long)bh * p.S_att + it.seg) * MK_PART;
                            if (tid == 0) { pp[0] = M; pp[1] = den; }
                            pp[4 + tid] = num;
                        }
                        sync_ok();
                        if (tid == 0) { __threadfence(); T.last_flag = (ok && atomicAdd(fl + d.f_item + (int)bh, 1) == p.S_att - 1) ? 1 : 0; __threadfence(); }
                        sync_ok();
                        finish = T.last_flag != 0;
                        if (finish && tid < MK_D) {
                            const float* pb = p.att_part + (long long)bh * p.S_att * MK_PART;
                            float MM = -INFINITY;
                            for (int s = 0; s < p.S_att; ++s) MM = fmaxf(MM, __ldcg(pb + (long long)s * MK_PART));
                            num = 0.f; den = 0.f;
                            for (int s = 0; s < p.S_att; ++s) {
                                const float ms = __ldcg(pb + (long long)s * MK_PART);
                                const float c = (ms == -INFINITY) ? 0.f : exp2f(ms - MM);
                                num += c * __ldcg(pb + (long long)s * MK_PART + 4 + tid);
                                den += c * __ldcg(pb + (long long)s * MK_PART + 1);
                            }
                        }
                    }
                    if (finish && ok && tid < MK_D) aout[tid] = __float2bfloat16_rn(den > 0.f ? num / den : 0.f);
                    sync_ok();                   // also protects st_* / s_new reuse by the next item
                    if (finish && ok && tid == 0) { fence_proxy_async_all(); red_release_add(fl + d.f_head, 1); }
                }
            }
            if (tid == 0) T.attn_done = layer + 1;      // every worker warp is past the last item's final barrier: the K/V slots are drained
            if (tid == 0 && layer == 0) mk_tl(ctx, 2);
            // ---------------- o projection, residual + post-attention RMSNorm
            const MkGemm go = mk_gemm(p, d, layer, 1);
            epilogue(go, p.ws_o, fl + d.f_to);
            if (tid == 0 && layer == 0) mk_tl(ctx, 3);
            if (cta < p.B)
                reduce_norm(p.ws_o, fl + d.f_to, go.tiles, go.kb, p.ln_w + (long long)(2 * layer + 1) * p.Hd, p.y_mlp, tokdone + layer * 2 + 1);
            if (tid == 0 && layer == 0) mk_tl(ctx, 4);
            // ---------------- gate/up, SwiGLU blocks (64 outputs = one 128-row tile of interleaved gate/up rows)
            const MkGemm gg = mk_gemm(p, d, layer, 2);
            epilogue(gg, p.ws_gu, fl + d.f_tg);
            if (tid == 0 && layer == 0) mk_tl(ctx, 5);
            for (int j = cta; j < gg.tiles; j += G) {
                int first, n;
                mk_contrib(j, gg.kb, gg.tiles, G, first, n);
                if (tid == 0 && ok) ok = mk_wait_flag(ctx, fl + d.f_tg + j, n, 601, j);
                sync_ok();
                const int tok = tid >> 4, o4 = (tid & 15) * 4;      // 4 outputs = 8 interleaved (gate, up) rows
                if (tok < p.B && ok) {
                    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
                    const float* src = p.ws_gu + (((long long)j * MK_MAXC) * MK_TOK + tok) * 128 + o4 * 2;
                    for (int s = 0; s < n; ++s) {
                        const float4 t0 = __ldcg(reinterpret_cast<const float4*>(src + (long long)s * MK_TOK * 128));
                        const float4 t1 = __ldcg(reinterpret_cast<const float4*>(src + (long long)s * MK_TOK * 128 + 4));
                        a0.x += t0.x; a0.y += t0.y; a0.z += t0.z; a0.w += t0.w;
                        a1.x += t1.x; a1.y += t1.y; a1.z += t1.z; a1.w += t1.w;
                    }
                    *reinterpret_cast<uint2*>(p.gu + (long long)tok * p.I + j * 64 + o4) =
                        make_uint2(pack_bf16x2(silu(a0.x) * a0.y, silu(a0.z) * a0.w), pack_bf16x2(silu(a1.x) * a1.y, silu(a1.z) * a1.w));
                }
                sync_ok();
                if (ok && tid == 0) { fence_proxy_async_all(); red_release_add(fl + d.f_gub, 1); }
            }
            if (tid == 0 && layer == 0) mk_tl(ctx, 6);
            // ---------------- down projection, residual + the next layer's input RMSNorm (or the final norm)
            const MkGemm gd = mk_gemm(p, d, layer, 3);
            epilogue(gd, p.ws_down, fl + d.f_td);
            if (tid == 0 && layer == 0) mk_tl(ctx, 7);
            if (cta < p.B)
                reduce_norm(p.ws_down, fl + d.f_td, gd.tiles, gd.kb, p.ln_w + (long long)(2 * layer + 2) * p.Hd, p.y_attn, tokdone + (layer + 1) * 2);
            if (tid == 0 && layer == 0) mk_tl(ctx, 8);
        }
        // ---------------- heads: logits tiles + per-tile argmax candidates, then the final argmax and the position advance
        {
            const MkGemm gh = mk_gemm(p, d, p.L, 4);
            int* th_done = p.flags + d.f_th;
            if (tid == 0) mk_tl(ctx, 9);
            epilogue(gh, p.ws_head, th_done);
            if (tid == 0) mk_tl(ctx, 10);
            for (int t = cta; t < gh.tiles; t += G) {
                int first, n;
                mk_contrib(t, gh.kb, gh.tiles, G, first, n);
                if (tid == 0 && ok) ok = mk_wait_flag(ctx, th_done + t, n, 701, t);
                sync_ok();
                const int tok = tid >> 4, f8 = (tid & 15) * 8;
                float best = -INFINITY; int besti = 0x7fffffff;
                if (tok < p.B && ok) {
                    float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                    const float* src = p.ws_head + (((long long)t * MK_MAXC) * MK_TOK + tok) * 128 + f8;
                    for (int s = 0; s < n; ++s) {
                        const float4 t0 = __ldcg(reinterpret_cast<const float4*>(src + (long long)s * MK_TOK * 128));
                        const float4 t1 = __ldcg(reinterpret_cast<const float4*>(src + (long long)s * MK_TOK * 128 + 4));
                        v[0] += t0.x; v[1] += t0.y; v[2] += t0.z; v[3] += t0.w; v[4] += t1.x; v[5] += t1.y; v[6] += t1.z; v[7] += t1.w;
                    }
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const int col = t * 128 + f8 + e;
                        if (col < p.V) {
                            p.logits[(long long)tok * p.V + col] = v[e];
                            if (v[e] > best) { best = v[e]; besti = col; }     // first maximum (torch.argmax)
                        }
                    }
                }
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) {     // 16 lanes of a half-warp hold one token's tile
                    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
                    if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
                }
                if ((tid & 15) == 0 && tok < p.B && ok) { p.cand_val[t * MK_TOK + tok] = best; p.cand_idx[t * MK_TOK + tok] = besti; }
                sync_ok();
                if (ok && tid == 0) red_release_add(p.flags + d.f_logits, 1);
            }
            if (tid == 0) mk_tl(ctx, 11);
            if (cta == 0) {
                if (tid == 0 && ok) ok = mk_wait_flag(ctx, p.flags + d.f_logits, gh.tiles, 702, 0);
                sync_ok();
                const int tok = tid >> 4;
                float best = -INFINITY; int besti = 0x7fffffff;
                if (tok < p.B && ok)
                    for (int t = (tid & 15); t < gh.tiles; t += 16) {
                        const float v = __ldcg(p.cand_val + t * MK_TOK + tok);
                        const int ix = __ldcg(p.cand_idx + t * MK_TOK + tok);
                        if (v > best || (v == best && ix < besti)) { best = v; besti = ix; }
                    }
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) {
                    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
                    if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
                }
                if ((tid & 15) == 0 && tok < p.B && ok) { p.ids[tok] = besti; p.kv_len[tok] = T.kvlen[tok] + 1; }
                if (tid == 0 && ok) *p.pos = pos + 1;
                if (tid == 0) mk_tl(ctx, 12);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<64>(tmem_base);
    }
}

// host side ------------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiledMk)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiledMk mk_encode_fn() {
    static PFN_encodeTiledMk fn = nullptr;
    if (fn) return fn;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = reinterpret_cast<PFN_encodeTiledMk>(ptr);
    return fn;
}
static int mk_map(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    PFN_encodeTiledMk enc = mk_encode_fn();
    if (!enc) return GROMA_ERR_DRIVER;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * 2};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? GROMA_OK : GROMA_ERR_TMA_ENCODE;
}

}  // namespace gb
using namespace gb;

GROMA_API int32_t groma_decode_step_layout(int32_t L, int32_t B, int32_t H, int32_t Hd, int32_t I, int32_t V, int64_t* layout, void* /*stream*/) {
    if (!layout || L < 1 || B < 1 || H < 1 || Hd < 128 || I < 64 || V < 1) return GROMA_ERR_ARG;
    const long long tq = 3 * Hd / 128, to = Hd / 128, tg = 2 * I / 128, td = Hd / 128, th = (V + 127) / 128;
    const long long lstride = tq + to + tg + td + (long long)B * H + H + tg;
    layout[0] = L * lstride + th + 1 + (L + 1) * 2;      // == MkDims::f_tok + (L + 1) * 2
    layout[1] = (long long)MK_MAXC * MK_TOK * 128;
    layout[2] = MK_PART;
    layout[3] = MK_TOK;
    return GROMA_OK;
}

GROMA_API int32_t groma_decode_step_fused(const groma_decode_step_args* a, void* stream) {
    if (!a || !a->w_arena || !a->w_down || !a->kv || !a->flags || !a->status) return GROMA_ERR_ARG;
    if (a->B < 1 || a->B > MK_TOK || a->H < 1 || a->L < 1 || a->S_att < 1) return GROMA_ERR_UNSUPPORTED;
    if (a->Hd != a->H * MK_D || (a->Hd % 128) || (a->I % 64) || ((2 * a->I) % 128) || a->Hd > 8192) return GROMA_ERR_UNSUPPORTED;
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int grid = a->grid > 0 ? a->grid : sms;
    if (grid > sms) return GROMA_ERR_ARG;      // all CTAs must be co-resident (one per SM)
    // every row tile may receive partials from at most MK_MAXC consecutive CTAs
    {
        const long long tiles[5] = {3LL * a->Hd / 128, a->Hd / 128, 2LL * a->I / 128, a->Hd / 128, (a->V + 127) / 128};
        const long long kbs[5] = {a->Hd / 64, a->Hd / 64, a->Hd / 64, a->I / 64, a->Hd / 64};
        for (int g = 0; g < 5; ++g) {
            const long long U = tiles[g] * kbs[g];
            if (U < grid) return GROMA_ERR_UNSUPPORTED;   // every CTA owns >= 1 unit, so the contributors of a tile are consecutive CTAs
            for (long long t = 0; t < tiles[g]; ++t) {
                const long long f = ((t * kbs[g] + 1) * grid - 1) / U, l = (((t + 1) * kbs[g]) * grid - 1) / U;
                if (l - f + 1 > MK_MAXC) return GROMA_ERR_UNSUPPORTED;
            }
        }
    }
    MkParams p;
    const uint64_t RW = 4ull * a->Hd + 2ull * a->I;
    int rc;
    if ((rc = mk_map(&p.map_w, a->w_arena, (uint64_t)a->L * RW + a->V, a->Hd, 128))) return rc;
    if ((rc = mk_map(&p.map_wd, a->w_down, (uint64_t)a->L * a->Hd, a->I, 128))) return rc;
    if ((rc = mk_map(&p.map_yattn, a->y_attn, a->B, a->Hd, MK_TOK))) return rc;
    if ((rc = mk_map(&p.map_ymlp, a->y_mlp, a->B, a->Hd, MK_TOK))) return rc;
    if ((rc = mk_map(&p.map_a, a->a, a->B, a->Hd, MK_TOK))) return rc;
    if ((rc = mk_map(&p.map_gu, a->gu, a->B, a->I, MK_TOK))) return rc;
    p.pf_slots = a->l2_prefetch_slots; p.L = a->L; p.B = a->B; p.H = a->H; p.Hd = a->Hd; p.I = a->I; p.V = a->V; p.vocab = a->vocab; p.S_att = a->S_att;
    p.cap = a->cap; p.scale_log2 = a->scale * 1.4426950408889634f; p.eps = a->eps;
    p.embed = reinterpret_cast<const __nv_bfloat16*>(a->embed); p.new_embed = reinterpret_cast<const __nv_bfloat16*>(a->new_embed);
    p.ln_w = a->ln_w; p.kv = reinterpret_cast<__nv_bfloat16*>(a->kv); p.rope_cos = a->rope_cos; p.rope_sin = a->rope_sin;
    p.ids = reinterpret_cast<long long*>(a->ids); p.pos = a->pos; p.kv_len = a->kv_len;
    p.x = reinterpret_cast<__nv_bfloat16*>(a->x); p.y_attn = reinterpret_cast<__nv_bfloat16*>(a->y_attn);
    p.y_mlp = reinterpret_cast<__nv_bfloat16*>(a->y_mlp); p.a = reinterpret_cast<__nv_bfloat16*>(a->a); p.gu = reinterpret_cast<__nv_bfloat16*>(a->gu);
    p.logits = a->logits; p.ws_qkv = a->ws_qkv; p.ws_o = a->ws_o; p.ws_gu = a->ws_gu; p.ws_down = a->ws_down; p.ws_head = a->ws_head;
    p.att_part = a->att_part; p.cand_val = a->cand_val; p.cand_idx = a->cand_idx; p.flags = a->flags; p.status = a->status;
    p.timeline = reinterpret_cast<unsigned long long*>(a->timeline);
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(decode_step_megakernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MK_SMEM_BYTES) != cudaSuccess) return GROMA_ERR_CUDA;
        attr_set = true;
    }
    decode_step_megakernel<<<grid, MK_THREADS, MK_SMEM_BYTES, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    return cudaGetLastError() == cudaSuccess ? GROMA_OK : GROMA_ERR_CUDA;
}
