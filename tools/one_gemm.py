import sys, torch
sys.path.insert(0, ".")
from groma_b200 import ops as G
M, N, K = [int(x) for x in sys.argv[1:4]]
a = torch.randn(M, K, device="cuda").bfloat16(); w = torch.randn(N, K, device="cuda").bfloat16()
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
for _ in range(4):
    G.gemm(a, w, out=out)
torch.cuda.synchronize()
