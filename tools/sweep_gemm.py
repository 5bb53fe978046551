import sys, torch
sys.path.insert(0, ".")
from groma_b200 import ops as G
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    ts.sort(); return ts[len(ts) // 2]
SHAPES = [(15456, 4096, 4096), (15456, 12288, 4096), (15456, 22016, 4096), (15456, 4096, 11008), (8192, 8192, 8192), (16400, 1024, 1024), (16400, 1024, 4096), (16400, 3072, 1024), (16400, 4096, 1024), (4800, 256, 256), (16384, 256, 1024), (1600, 4096, 1024)]
if len(sys.argv) > 1 and sys.argv[1] == "vit":      # the short-K ViT projections + one LLaMA shape (epilogue A/B runs)
    SHAPES = [(16400, 1024, 1024), (16400, 1024, 4096), (16400, 3072, 1024), (16400, 4096, 1024), (15456, 4096, 4096)]
for (M, N, K) in SHAPES:
    a = torch.randn(M, K, device="cuda").bfloat16(); w = torch.randn(N, K, device="cuda").bfloat16()
    res = torch.randn(M, N, device="cuda").bfloat16(); gamma = torch.randn(N, device="cuda"); bias = torch.randn(N, device="cuda")
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    line = f"M={M} N={N} K={K}: "
    for bn in (128, 256, 512):
        if min(bn, 256) > N: continue
        ms = timeit(lambda: G.gemm(a, w, out=out, block_n=bn, bias=bias, gamma=gamma, residual=res))
        line += f"bn{bn} {ms*1000:.0f}us ({2.0*M*N*K/ms/1e9:.0f} TF)  "
    ms = timeit(lambda: torch.matmul(a, w.t(), out=out))
    print(line + f"| cublas {ms*1000:.0f}us", flush=True)
