"""A/B of the persistent decode kernel against the multi-kernel CUDA-graph step at Groma-7B shapes (B=16, ctx ~1030):
same inputs, same KV cache -> logits / next ids / appended K,V rows compared, then both timed.
    python tools/decode_mega.py [ctx] [B]"""
import os, sys, math, torch
sys.path.insert(0, ".")
from groma_b200.config import PathConfig, tiny_config
from groma_b200.synth import make_state_dict
from groma_b200.engine import GromaEngine

ctx = int(sys.argv[1]) if len(sys.argv) > 1 else 1030
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
tiny = os.environ.get("TINY", "0") == "1"
over = {k[3:].lower(): int(v) for k, v in os.environ.items() if k.startswith("MK_")}     # e.g. MK_LLM_LAYERS=1 MK_LLM_HIDDEN=1024 MK_LLM_HEADS=8
cfg = tiny_config(box_score_thres=0.0, **over) if tiny else PathConfig(box_score_thres=0.0, **over)
sd = make_state_dict(cfg, seed=0, perturb_norms=True, dtype=torch.bfloat16, device="cuda")
eng = GromaEngine(cfg, sd)
del sd
cap = ctx + 70
eng.alloc_kv(B, cap)
eng.ensure_rope(cap)
torch.manual_seed(1)
eng.kv.normal_(0, 0.5)
kv0 = eng.kv.clone()
d = eng._decode_buffers(B)
ids0 = torch.randint(1000 if cfg.vocab > 2000 else 10, cfg.vocab, (B,), device="cuda")

def reset():
    d["ids"].copy_(ids0); d["pos"].fill_(ctx); d["kv_len"].fill_(ctx + 1)

def run(mega, steps=1):
    eng.use_megakernel = mega
    eng.kv.copy_(kv0)
    reset()
    outs = []
    for _ in range(steps):
        lg = eng.decode_step(B).clone()
        outs.append((lg, d["ids"].clone()))
    torch.cuda.synchronize()
    eng.check_decode_status()
    return outs, eng.kv[:, :, :, :, ctx:ctx + steps].clone(), int(d["pos"].item()), d["kv_len"].clone()

ref, kv_ref, pos_r, kvl_r = run(False, 3)
got, kv_got, pos_g, kvl_g = run(True, 3)
print("grid", eng._mk["grid"], "S_att", eng._mk["s_att"], "pos", pos_r, pos_g, "kv_len equal", torch.equal(kvl_r, kvl_g))
for s, ((l0, i0), (l1, i1)) in enumerate(zip(ref, got)):
    e = ((l0 - l1).abs().max() / l0.abs().max()).item()
    print(f"step {s}: logits nrel {e:.3e}  ids equal {torch.equal(i0, i1)}  {i0[:6].tolist()} {i1[:6].tolist()}")
ek = ((kv_ref.float() - kv_got.float()).abs().max() / kv_ref.float().abs().max()).item()
print(f"appended K/V rows nrel {ek:.3e}")

if os.environ.get("TIMELINE", "0") == "1":
    # phase-boundary stamps (%globaltimer) of one eager megakernel step: [grid][role][event], microseconds since the first stamp
    eng.use_megakernel = True
    st = eng._mega_state(B)
    st["timeline"] = torch.zeros((st["grid"] * 4 * 32,), dtype=torch.int64, device="cuda")
    for _ in range(2):
        eng.kv.copy_(kv0); reset(); st["timeline"].zero_()
        eng.decode_step(B); torch.cuda.synchronize()
    eng.check_decode_status()
    tl = st["timeline"].cpu().view(st["grid"], 4, 32).double()
    t0 = tl[tl > 0].min()
    names = {2: ["L0 start", "qkv epi done", "attn done", "o epi done", "norm1 done", "gu epi done", "swiglu done", "down epi done", "norm2 done",
                 "head start", "head epi done", "logits done", "argmax done"],
             1: ["qkv mma start", "qkv mma issued", "", "", "o mma start", "o mma issued", "gu mma start", "gu mma issued", "down mma start", "down mma issued",
                 "head mma start", "head mma issued"],
             0: ["stream start", "stream all issued"]}
    for role, nm in names.items():
        for e, n in enumerate(nm):
            v = tl[:, role, e]
            v = v[v > 0]
            if n and len(v):
                v = (v - t0) / 1e3
                print(f"  role {role} {n:>18s}: min {v.min():8.1f}  median {v.median():8.1f}  max {v.max():8.1f} us  (n={len(v)})")
    st["timeline"] = None


def graph_of(fn):
    g = torch.cuda.CUDAGraph(); st = torch.cuda.Stream(); st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        fn()
        with torch.cuda.graph(g, stream=st, capture_error_mode="thread_local"):
            fn()
    torch.cuda.current_stream().wait_stream(st)
    return g

def timeit(g, reps=20):
    reset()
    for _ in range(3): g.replay()
    ts = []
    for _ in range(reps):
        reset()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort(); return ts[len(ts) // 2], ts[0]

Hd, I, V, L = cfg.llm_hidden, cfg.llm_inter, cfg.vocab + cfg.num_new_token, cfg.llm_layers
wbytes = L * (4 * Hd * Hd + 3 * Hd * I) * 2 + V * Hd * 2
kvbytes = L * 2 * B * (ctx + 1) * Hd * 2
for mega in (False, True):
    eng.use_megakernel = mega
    reset()
    ms, best = timeit(graph_of(lambda: eng.decode_step(B)))
    eng.check_decode_status()
    print(f"{'megakernel' if mega else 'graph step'}: median {ms:.3f} ms  best {best:.3f} ms  ({(wbytes + kvbytes) / ms / 1e6:.0f} GB/s of {((wbytes + kvbytes) / 1e9):.2f} GB)", flush=True)
