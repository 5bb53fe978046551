"""The kernels of one decode step, one eager launch per distinct shape, for `ncu --set full` (DRAM traffic per launch):
the five swap-AB GEMM shapes (qkv, o, gate/up, down, heads) with the step's split-K factors and the decode attention at ctx 1030.
    ncu --set full --clock-control none -k regex:'gemm_bf16_tcgen05|decode_attention' -c 6 -o gpurun_out/r2_decode_traffic python tools/decode_traffic.py"""
import math, sys, torch
sys.path.insert(0, ".")
from groma_b200 import ops as G
from groma_b200.config import PathConfig
from groma_b200.engine import GromaEngine

cfg = PathConfig()
B, ctx = 16, 1030
Hd, I, V, H = cfg.llm_hidden, cfg.llm_inter, cfg.vocab + cfg.num_new_token, cfg.llm_heads
eng = GromaEngine.__new__(GromaEngine)      # only _decode_splits() is needed: no weights are packed
eng.cfg, eng._splits = cfg, None
sp = eng._decode_splits()
shapes = [("qkv", 3 * Hd, Hd, sp["qkv"]), ("o", Hd, Hd, sp["o"]), ("gu", 2 * I, Hd, sp["gu"]), ("down", Hd, I, sp["down"]), ("head", V, Hd, sp["head"])]
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
print("splits", sp)
for name, N, K, s in shapes:
    W = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
    x = torch.randn(B, K, device="cuda").bfloat16()
    ws = torch.empty(s, B, N, dtype=torch.float32, device="cuda")
    flush.zero_()
    G.gemm_swap_ab(x, W, ws, split_k=s, pdl=False, transposed=True)
    torch.cuda.synchronize()
    print(name, N, K, s, "algorithmic weight bytes", N * K * 2)
kc = torch.randn(B, H, ctx + 64, 128, device="cuda").bfloat16(); vc = torch.randn_like(kc)
q = torch.randn(B, H * 128, device="cuda").bfloat16(); out = torch.empty_like(q)
kv_len = torch.full((B,), ctx + 1, dtype=torch.int32, device="cuda")
flush.zero_()
G.decode_attention(q, kc, vc, kv_len, 1 / math.sqrt(128), out)
torch.cuda.synchronize()
print("attention algorithmic K/V bytes", 2 * B * H * (ctx + 1) * 128 * 2)
