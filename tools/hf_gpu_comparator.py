"""Informational comparator (NOT part of bench.py's contract): stock PyTorch / HF transformers modules of the two big
sub-models of the path on the same GPU, bf16, same shapes as the bench workload (B=16, 448x448 images, prefill T=966,
128 greedy tokens, random-init weights).  It leaves out the DDETR proposer, NMS, the region encoder (~2 TFLOP/img of
3x3 convs) and the sequence assembly, so it is a LOWER bound on what a PyTorch-on-GPU build of the reference spends
(SURVEY.md §8d "comparator").  Prints one JSON line.
    python tools/hf_gpu_comparator.py [attn_implementation]"""
import json, sys, time
import torch
from transformers import Dinov2Config, Dinov2Model, LlamaConfig, LlamaForCausalLM

attn = sys.argv[1] if len(sys.argv) > 1 else "sdpa"
B, T, NEW = 16, 966, 128
dev = torch.device("cuda:0")
torch.manual_seed(0)
torch.set_default_dtype(torch.bfloat16)
with torch.device(dev):
    vit = Dinov2Model(Dinov2Config(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, mlp_ratio=4, image_size=518,
                                   patch_size=14, attn_implementation="sdpa")).eval()
    llm = LlamaForCausalLM(LlamaConfig(hidden_size=4096, intermediate_size=11008, num_hidden_layers=32, num_attention_heads=32,
                                       num_key_value_heads=32, vocab_size=32000, max_position_embeddings=4096, rms_norm_eps=1e-5,
                                       attn_implementation=attn)).eval()
torch.set_default_dtype(torch.float32)
images = torch.randn(B, 3, 448, 448, device=dev, dtype=torch.bfloat16)
emb = (torch.randn(B, T, 4096, device=dev) * 0.02).to(torch.bfloat16)
mask = torch.ones(B, T, dtype=torch.long, device=dev)


def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e


@torch.inference_mode()
def step(new):
    e0 = ev()
    vit(pixel_values=images, output_hidden_states=True)
    e1 = ev()
    out = llm.generate(inputs_embeds=emb, attention_mask=mask, max_new_tokens=new, min_new_tokens=new, do_sample=False, use_cache=True,
                       pad_token_id=0)
    e2 = ev()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1), e1.elapsed_time(e2), out.shape


step(8); step(8)
res = [step(NEW) for _ in range(2)]
vit_ms = min(r[0] for r in res); llm_ms = min(r[1] for r in res)
print(json.dumps({"what": "HF transformers %s on the same B200: Dinov2Model (bf16, sdpa) + LlamaForCausalLM.generate (bf16, %s, DynamicCache, greedy), "
                          "B=16, T=966 prefill + 128 new tokens; no DDETR / region encoder / NMS" % (__import__("transformers").__version__, attn),
                  "vit_ms": vit_ms, "llm_prefill_plus_decode_ms": llm_ms, "images_per_sec_upper_bound": B / ((vit_ms + llm_ms) / 1000.0),
                  "generated_shape": list(res[0][2])}))
