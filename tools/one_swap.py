import sys, torch
sys.path.insert(0, ".")
from groma_b200 import ops as G
M, N, K, split = [int(x) for x in sys.argv[1:5]]
x = torch.randn(M, K, device="cuda").bfloat16()
ws_ = [torch.randn(N, K, device="cuda").bfloat16() for _ in range(6)]   # rotate weights: cold L2 like the real decode
ws = torch.empty(split, N, M, device="cuda", dtype=torch.float32)
for i in range(12):
    G.gemm_swap_ab(x, ws_[i % 6], ws.view(split, M, N), split_k=split, transposed=True)
torch.cuda.synchronize()
