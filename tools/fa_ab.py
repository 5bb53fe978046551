"""Steady-state timing of one flash-attention shape (CUDA events around batches of 20 launches, 15 batches; min and median):
    python tools/fa_ab.py B H S D causal       (GROMA_FA_TAILS=0 for the plain tiling)"""
import math, os, sys, torch
sys.path.insert(0, ".")
from groma_b200 import ops as G
B, H, S, D, causal = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), bool(int(sys.argv[5]))
q = torch.randn(B, S, H, D, device="cuda").bfloat16(); k = torch.randn(B, H, S, D, device="cuda").bfloat16(); v = torch.randn(B, H, S, D, device="cuda").bfloat16()
o = torch.empty(B, S, H * D, device="cuda", dtype=torch.bfloat16)
f = lambda: G.attention_tc(q, k, v, causal=causal, scale=1 / math.sqrt(D), out=o)
for _ in range(30): f()
ts = []
for _ in range(15):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): f()
    e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) / 20)
ts.sort()
fl = 4.0 * B * H * S * S * D * (0.5 if causal else 1.0)
print(f"[tails={os.environ.get('GROMA_FA_TAILS', '2')}] B={B} H={H} S={S} D={D} causal={causal}: min {ts[0]*1e3:.1f} us  median {ts[7]*1e3:.1f} us  "
      f"-> {fl / ts[7] / 1e9:.0f} TFLOP/s (median)", flush=True)
