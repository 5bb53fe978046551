"""BASELINE.json configs[1] (DINOv2-L only, B=64) and configs[2] (region tokenizer: proposer + RoIAlign + projector, B=32,
300 proposals, R=100) on one B200, plus the MSDA sampling kernel against its algorithmic bytes."""
import sys, json, torch
sys.path.insert(0, ".")
from groma_b200 import ops as G
from groma_b200.config import PathConfig
from groma_b200.synth import make_state_dict
from groma_b200.engine import GromaEngine

def timeit(fn, iters=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    ts.sort(); return ts[len(ts) // 2]

cfg = PathConfig(box_score_thres=0.0)
# vision-only: keep the LLM tiny so the state dict stays small
small = PathConfig(box_score_thres=0.0, llm_layers=1, vocab=1024)
small.llm_hidden = cfg.llm_hidden
eng = GromaEngine(small, make_state_dict(small, 0, perturb_norms=False, dtype=torch.bfloat16, device="cuda"))
out = {}
only_region = len(sys.argv) > 1 and sys.argv[1] == "region"      # A/B runs of the region tokenizer alone
if not only_region:
    B = 64
    img = torch.randn(B, 3, 448, 448, device="cuda")
    ms = timeit(lambda: eng.vit(img))
    out["config2_vit_B64"] = {"ms": ms, "images_per_s": B / ms * 1e3, "tflops": B * 723.6e9 / ms / 1e9,
                              "frac_of_sustained_bf16_peak": B * 723.6e9 / ms / 1e9 / 1441.5}
B = 32
img = torch.randn(B, 3, 448, 448, device="cuda")
hs = eng.vit(img)
g = torch.Generator().manual_seed(0)
boxes = [torch.rand(100, 4, generator=g) * 0.6 + 0.2 for _ in range(B)]
def region():
    pc, px, sc, _ = eng.proposer(hs)
    return eng.region_encoder(hs, boxes)
ms = timeit(region, iters=3)
fl = B * (14.4e9 + 2074.6e9 + 100 * 11.5e9)
out["config3_region_tokenizer_B32_R100"] = {"ms": ms, "images_per_s": B / ms * 1e3, "tflops": fl / ms / 1e9, "lower_bound_ms": 72.0}
if only_region:
    print(json.dumps(out, indent=1)); sys.exit(0)
ms_p = timeit(lambda: eng.proposer(hs), iters=3)
out["proposer_only_B32"] = {"ms": ms_p}
# MSDA kernel alone (encoder shape and decoder shape), algorithmic bytes per SURVEY 8d
for name, Q, ref_dim in (("enc", 1024, 2), ("dec", 300, 4)):
    value = torch.randn(B, 1024, 8, 32, device="cuda").bfloat16()
    proj = torch.randn(B * Q, 96, device="cuda")
    ref = torch.rand(B, Q, ref_dim, device="cuda")
    o = torch.empty(B, Q, 256, device="cuda", dtype=torch.bfloat16)
    ms = timeit(lambda: G.msda(value, proj, ref, [(32, 32)], 8, 4, out=o), iters=20, warm=5)
    alg = B * (1024 * 256 * 2 + Q * 96 * 4 + Q * ref_dim * 4 + Q * 256 * 2)
    out[f"msda_{name}_B32"] = {"us": ms * 1e3, "algorithmic_MB": alg / 1e6, "GBps": alg / ms / 1e6, "frac_hbm_peak": alg / ms / 1e6 / 6577.7}
print(json.dumps(out, indent=1))
