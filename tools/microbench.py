"""Kernel micro-benchmarks (CUDA events, L2 flushed between iterations). Usage: python tools/microbench.py [gemm|attn|decode|conv]"""
import sys, math, json
import torch
sys.path.insert(0, ".")
from groma_b200 import ops as G

flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")

def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]

def bench_gemm():
    for (M, N, K, bn) in [(8192, 8192, 8192, 256), (8192, 8192, 8192, 128), (15456, 4096, 4096, 0), (15456, 12288, 4096, 0), (15456, 22016, 4096, 0),
                          (15456, 4096, 11008, 0), (16400, 1024, 1024, 0), (16400, 3072, 1024, 0), (16400, 4096, 1024, 0), (16400, 1024, 4096, 0)]:
        a = torch.randn(M, K, device="cuda").bfloat16(); w = torch.randn(N, K, device="cuda").bfloat16()
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        ms = timeit(lambda: G.gemm(a, w, out=out, block_n=bn))
        ms_t = timeit(lambda: torch.matmul(a, w.t(), out=out))
        fl = 2.0 * M * N * K
        print(f"gemm M={M} N={N} K={K} bn={bn}: {ms:.3f} ms {fl/ms/1e9:.0f} TFLOP/s | cublas {ms_t:.3f} ms {fl/ms_t/1e9:.0f} TFLOP/s", flush=True)

def bench_conv():
    for (n_img, H, C) in [(4, 128, 1024), (16, 64, 1024), (16, 32, 1024)]:
        hp = H + 2
        x = torch.randn(n_img * hp * hp, C, device="cuda").bfloat16(); w = (torch.randn(C, 9 * C, device="cuda") * 0.02).bfloat16()
        ms = timeit(lambda: G.conv3x3_flat(x, w, n_img, hp, hp))
        fl = 2.0 * n_img * H * H * C * C * 9
        print(f"conv3x3 n={n_img} {H}x{H} C={C}: {ms:.3f} ms {fl/ms/1e9:.0f} TFLOP/s (useful)", flush=True)

def bench_attn():
    for (B, H, S, D, causal) in [(16, 32, 966, 128, True), (16, 16, 1025, 64, False), (16, 8, 300, 32, False)]:
        q = torch.randn(B, S, H, D, device="cuda").bfloat16(); k = torch.randn(B, H, S, D, device="cuda").bfloat16(); v = torch.randn(B, H, S, D, device="cuda").bfloat16()
        ms = timeit(lambda: G.attention(q, k, v, causal=causal, scale=1 / math.sqrt(D)))
        fl = 4.0 * B * H * S * S * D * (0.5 if causal else 1.0)
        print(f"attn(mma.sync) B={B} H={H} S={S} D={D} causal={causal}: {ms:.3f} ms {fl/ms/1e9:.0f} TFLOP/s", flush=True)
        if D in (64, 128):
            ms = timeit(lambda: G.attention_tc(q, k, v, causal=causal, scale=1 / math.sqrt(D)))
            print(f"attn(tcgen05)  B={B} H={H} S={S} D={D} causal={causal}: {ms:.3f} ms {fl/ms/1e9:.0f} TFLOP/s", flush=True)
    B, H, D, ctx = 16, 32, 128, 1024
    q = torch.randn(B, 1, H, D, device="cuda").bfloat16(); k = torch.randn(B, H, ctx, D, device="cuda").bfloat16(); v = torch.randn(B, H, ctx, D, device="cuda").bfloat16()
    ms = timeit(lambda: G.attention(q, k, v, causal=True, scale=0.1, q_pos0=ctx - 1))
    print(f"decode attn B={B} ctx={ctx}: {ms:.3f} ms {2*B*H*ctx*D*2/ms/1e6:.0f} GB/s", flush=True)

def bench_decode():
    for (M, N, K, split) in [(16, 12288, 4096, 1), (16, 12288, 4096, 2), (16, 4096, 4096, 4), (16, 4096, 4096, 8), (16, 22016, 4096, 1), (16, 4096, 11008, 4), (16, 4096, 11008, 8), (16, 32128, 4096, 1)]:
        x = torch.randn(M, K, device="cuda").bfloat16(); w = torch.randn(N, K, device="cuda").bfloat16()
        ws = torch.empty(split, N, M, device="cuda", dtype=torch.float32)
        ms = timeit(lambda: G.gemm_swap_ab(x, w, ws, split_k=split))
        print(f"swapAB M={M} N={N} K={K} split={split}: {ms*1000:.1f} us {N*K*2/ms/1e6:.0f} GB/s", flush=True)

which = sys.argv[1:] or ["gemm", "conv", "attn", "decode"]
for wname in which:
    {"gemm": bench_gemm, "conv": bench_conv, "attn": bench_attn, "decode": bench_decode}[wname]()
