"""One launch shape of the tcgen05 flash attention for ncu:  python tools/one_attn.py B H S D causal"""
import math, sys, torch
sys.path.insert(0, ".")
from groma_b200 import ops as G
B, H, S, D, causal = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), bool(int(sys.argv[5]))
q = torch.randn(B, S, H, D, device="cuda").bfloat16(); k = torch.randn(B, H, S, D, device="cuda").bfloat16(); v = torch.randn(B, H, S, D, device="cuda").bfloat16()
for _ in range(4):
    G.attention_tc(q, k, v, causal=causal, scale=1 / math.sqrt(D))
torch.cuda.synchronize()
