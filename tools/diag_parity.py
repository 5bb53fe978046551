"""Diagnostic: logits distance GPU vs oracle(bf16) vs oracle(fp32) on the tiny config."""
import sys, torch
sys.path.insert(0, ".")
from oracle.config import tiny_config, SyntheticTokenizer
from oracle.groma_oracle import Oracle
from oracle.weights import make_state_dict
from groma.model.groma import GromaConfig, GromaModel

def nrel(a, b): a, b = a.float().cpu(), b.float().cpu(); return ((a - b).abs().max() / b.abs().max()).item()
def rms(a, b): a, b = a.float().cpu(), b.float().cpu(); return ((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt()).item()

cfg = tiny_config(box_score_thres=0.0)
sd = make_state_dict(cfg, 0)
tok = SyntheticTokenizer(cfg.vocab)
ob = Oracle(cfg, sd, "bf16"); ob.init_special_token_id(tok)
of = Oracle(cfg, sd, "fp32"); of.init_special_token_id(tok)
m = GromaModel(GromaConfig.from_path_config(cfg), state_dict=sd, path_config=cfg); m.init_special_token_id(tok)
g = torch.Generator().manual_seed(0)
B, Tt = 2, 24
images = torch.randn(B, 3, 448, 448, generator=g)
ids = torch.randint(10, cfg.vocab, (B, Tt), generator=g); ids[:, 3] = tok.map["<image>"]; ids[:, 9] = tok.map["<region>"]
boxes = [torch.rand(5, 4, generator=g) * 0.6 + 0.2, torch.rand(9, 4, generator=g) * 0.6 + 0.2]
outb = ob.forward_prefill(ids.clone(), images, selected_override=boxes)
outf = of.forward_prefill(ids.clone(), images, selected_override=boxes)
res = m.forward(input_ids=ids.clone(), images=images.cuda(), use_cache=True, return_dict=True, _selected_override=boxes)
lg = res.logits.cpu()
print("logits: gpu vs oracle-bf16  nrel %.2e rms %.2e" % (nrel(lg, outb["logits"]), rms(lg, outb["logits"])))
print("logits: gpu vs oracle-fp32  nrel %.2e rms %.2e" % (nrel(lg, outf["logits"]), rms(lg, outf["logits"])))
print("logits: oracle-bf16 vs fp32 nrel %.2e rms %.2e" % (nrel(outb["logits"], outf["logits"]), rms(outb["logits"], outf["logits"])))
# teacher-forced LLM only: feed the oracle's inputs_embeds to the GPU LLM
x = outb["inputs_embeds"].to(torch.bfloat16).cuda().reshape(-1, cfg.llm_hidden).contiguous()
T = outb["input_ids"].shape[1]
m.engine.alloc_kv(B, T + 1)
kv_len = outb["attention_mask"].sum(1).to(torch.int32).cuda()
l2 = m.engine.llm_prefill(x, B, T, kv_len).reshape(B, T, -1).cpu()
print("LLM-only (teacher-forced embeds): nrel %.2e rms %.2e" % (nrel(l2, outb["logits"]), rms(l2, outb["logits"])))
print("embeds gpu vs oracle: nrel %.2e rms %.2e" % (nrel(m.engine.stages.get("x", x), outb["inputs_embeds"]), 0))
# single GEMM precision sanity: the head on identical hidden
