"""One tcgen05 GEMM shape under ncu:  python tools/one_gemm.py M N K [res]   (res = bias + LayerScale + residual epilogue, as the
ViT o / fc2 projections run it)"""
import sys, torch
sys.path.insert(0, ".")
from groma_b200 import ops as G
M, N, K = [int(x) for x in sys.argv[1:4]]
res_epi = len(sys.argv) > 4 and sys.argv[4] == "res"
a = torch.randn(M, K, device="cuda").bfloat16(); w = torch.randn(N, K, device="cuda").bfloat16()
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
kw = {}
if res_epi:
    kw = dict(bias=torch.randn(N, device="cuda"), gamma=torch.randn(N, device="cuda"), residual=torch.randn(M, N, device="cuda").bfloat16())
for _ in range(4):
    G.gemm(a, w, out=out, **kw)
torch.cuda.synchronize()
