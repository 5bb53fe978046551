"""GPU comparator arm of BASELINE.md section 4b: the Groma forward path written the way the reference writes it -- stock PyTorch /
HF transformers modules, one op per line -- run on the SAME B200 in bf16: HF `Dinov2Model` (sdpa), the Deformable-DETR proposer
with the grid_sample MSDA core HF falls back to, `torchvision.ops.nms` per image with the reference's host loop and
`torch.randperm`, the region encoder with cuDNN convs + `torchvision.ops.roi_align`, HF `LlamaForCausalLM.generate`
(flash-attention-2 when importable, else sdpa; DynamicCache; greedy).  Random-init weights of the Groma-7B architecture,
BASELINE.json configs 2 (ViT only, B=64), 3 (region tokenizer, B=32, R=100) and 4 (end to end, B=16, 512-token prompt, 128 new
tokens).  NOT part of bench.py's contract and none of this repo's kernels are on its path; it prints one JSON object that
profiles/r02_gpu_comparator.json records next to this repo's numbers for the same configs.
    python tools/gpu_comparator.py [2|3|4 ...]"""
import json
import math
import sys
import time

import torch
import torch.nn as nn
import torch.nn.functional as F
import torchvision
from transformers import Dinov2Config, Dinov2Model, LlamaConfig, LlamaForCausalLM

dev = torch.device("cuda:0")
torch.backends.cudnn.benchmark = True
torch.backends.cuda.matmul.allow_tf32 = True
torch.backends.cudnn.allow_tf32 = True
BF = torch.bfloat16


def lin(i, o, bias=True):
    m = nn.Linear(i, o, bias=bias)
    nn.init.normal_(m.weight, std=0.02)
    return m


class MSDA(nn.Module):
    """DeformableDetrMultiscaleDeformableAttention, 1 level, with the pure-PyTorch sampling core ($HF .../modeling_deformable_detr.py:171-222)."""

    def __init__(self, d=256, heads=8, points=4):
        super().__init__()
        self.h, self.p = heads, points
        self.sampling_offsets, self.attention_weights = lin(d, heads * points * 2), lin(d, heads * points)
        self.value_proj, self.output_proj = lin(d, d), lin(d, d)

    def forward(self, query, value_src, ref, g):
        B, Q, D = query.shape
        h, p, hd = self.h, self.p, D // self.h
        value = self.value_proj(value_src).view(B, g * g, h, hd)
        off = self.sampling_offsets(query).view(B, Q, h, 1, p, 2).float()
        aw = F.softmax(self.attention_weights(query).view(B, Q, h, p).float(), -1)
        if ref.shape[-1] == 2:
            loc = ref[:, :, None, None, None, :] + off / torch.tensor([g, g], device=dev, dtype=torch.float32)
        else:
            loc = ref[:, :, None, None, None, :2] + off / p * ref[:, :, None, None, None, 2:] * 0.5
        grid = (2 * loc - 1)[:, :, :, 0].permute(0, 2, 1, 3, 4).reshape(B * h, Q, p, 2)
        v = value.permute(0, 2, 3, 1).reshape(B * h, hd, g, g).float()
        samp = F.grid_sample(v, grid, mode="bilinear", padding_mode="zeros", align_corners=False)        # [B*h, hd, Q, p]
        out = (samp * aw.permute(0, 2, 1, 3).reshape(B * h, 1, Q, p)).sum(-1).view(B, h * hd, Q).transpose(1, 2)
        return self.output_proj(out.to(query.dtype))


class EncLayer(nn.Module):
    def __init__(self, d=256, ffn=1024):
        super().__init__()
        self.self_attn, self.ln1, self.fc1, self.fc2, self.ln2 = MSDA(d), nn.LayerNorm(d), lin(d, ffn), lin(ffn, d), nn.LayerNorm(d)

    def forward(self, x, pos, ref, g):
        x = self.ln1(x + self.self_attn(x + pos, x, ref, g))
        return self.ln2(x + self.fc2(F.relu(self.fc1(x))))


class DecLayer(nn.Module):
    def __init__(self, d=256, ffn=1024, heads=8):
        super().__init__()
        self.h = heads
        self.q, self.k, self.v, self.o = lin(d, d), lin(d, d), lin(d, d), lin(d, d)
        self.ln1, self.cross, self.ln2, self.fc1, self.fc2, self.ln3 = nn.LayerNorm(d), MSDA(d), nn.LayerNorm(d), lin(d, ffn), lin(ffn, d), nn.LayerNorm(d)

    def forward(self, x, qpos, memory, ref, g):
        B, Q, D = x.shape
        qk = x + qpos
        sp = lambda t: t.view(B, Q, self.h, D // self.h).transpose(1, 2)
        a = F.scaled_dot_product_attention(sp(self.q(qk)), sp(self.k(qk)), sp(self.v(x))).transpose(1, 2).reshape(B, Q, D)
        x = self.ln1(x + self.o(a))
        x = self.ln2(x + self.cross(x + qpos, memory, ref, g))
        return self.ln3(x + self.fc2(F.relu(self.fc1(x))))


def mlp3(d, o):
    return nn.Sequential(lin(d, d), nn.ReLU(), lin(d, d), nn.ReLU(), lin(d, o))


class Proposer(nn.Module):
    """groma/model/ddetr.py:111-155 + ddetr_transformer.py:484-728 (two-stage, box refine, refs never advance)."""

    def __init__(self, C=1024, d=256, g=32, nq=300):
        super().__init__()
        self.g, self.nq, self.d = g, nq, d
        self.inproj, self.inln = lin(C, d), nn.LayerNorm(d, eps=1e-6)
        self.enc = nn.ModuleList(EncLayer(d) for _ in range(6))
        self.dec = nn.ModuleList(DecLayer(d) for _ in range(6))
        self.level_embed = nn.Parameter(torch.randn(1, d))
        self.tgt = nn.Embedding(nq, d)
        self.enc_output, self.enc_ln, self.pos_trans, self.pos_ln = lin(d, d), nn.LayerNorm(d), lin(2 * d, 2 * d), nn.LayerNorm(2 * d)
        self.cls_enc, self.cls_coco, self.cls_sa1b = lin(d, 1), lin(d, 1), lin(d, 1)
        self.bbox = nn.ModuleList(mlp3(d, 4) for _ in range(7))
        lin1 = (torch.arange(g, dtype=torch.float32) + 0.5) / g
        gy, gx = torch.meshgrid(lin1, lin1, indexing="ij")
        self.register_buffer("ref2", torch.stack([gx.reshape(-1), gy.reshape(-1)], -1), persistent=False)
        prop = torch.cat([self.ref2, torch.full((g * g, 2), 0.05)], -1)
        self.register_buffer("prop_logit", torch.log(prop / (1 - prop)), persistent=False)
        dim_t = 10000 ** (2 * torch.div(torch.arange(d // 2, dtype=torch.float32), 2, rounding_mode="floor") / (d // 2))
        self.register_buffer("dim_t", dim_t, persistent=False)
        ones = torch.ones(1, g, g)
        ye, xe = (ones.cumsum(1) - 0.5) / (g + 1e-6) * 2 * math.pi, (ones.cumsum(2) - 0.5) / (g + 1e-6) * 2 * math.pi
        px, py = xe[..., None] / dim_t, ye[..., None] / dim_t
        sine = lambda t: torch.stack((t[..., 0::2].sin(), t[..., 1::2].cos()), 4).flatten(3)
        self.register_buffer("pos", torch.cat((sine(py), sine(px)), 3).reshape(1, g * g, d), persistent=False)

    def forward(self, hs):
        B, g = hs[0].shape[0], self.g
        x = self.inln(self.inproj(torch.stack(hs[-4:]).mean(0)[:, 1:]))
        pos = (self.pos + self.level_embed).to(x.dtype)
        ref = self.ref2[None].expand(B, -1, -1)
        for l in self.enc:
            x = l(x, pos, ref, g)
        memory = x
        eo = self.enc_ln(self.enc_output(memory))
        cls = self.cls_enc(eo)[..., 0].float()
        coord = self.bbox[6](eo).float() + self.prop_logit
        topk = torch.topk(cls, self.nq, dim=1)[1]
        ref4 = torch.gather(coord, 1, topk[..., None].expand(-1, -1, 4)).sigmoid()
        pp = (ref4 * 2 * math.pi)[..., None] / self.dim_t
        pe = torch.stack((pp[..., 0::2].sin(), pp[..., 1::2].cos()), 4).flatten(2).to(x.dtype)
        qpos = self.pos_ln(self.pos_trans(pe))[..., :self.d]
        h = self.tgt.weight[None].expand(B, -1, -1)
        inter = []
        for l in self.dec:
            h = l(h, qpos, memory, ref4, g)
            inter.append(h)
        inv = lambda t: torch.log(t.clamp(1e-5, 1) / (1 - t).clamp(1e-5, 1))
        r1 = (self.bbox[4](inter[4]).float() + inv(ref4)).sigmoid()
        pred = (self.bbox[5](inter[5]).float() + inv(r1)).sigmoid()
        score = self.cls_coco(inter[5])[..., 0].float().sigmoid() ** 0.4 * self.cls_sa1b(inter[5])[..., 0].float().sigmoid() ** 0.6
        return pred, score


class RegionEncoder(nn.Module):
    """groma/model/roi_align.py:97-327: upsample + coord concat + 1x1, 5 shared fuse rounds (conv3x3 -> GN64 -> ReLU), 3-level RoIAlign,
    3 conv3x3, flatten linear, box MLP, updims."""

    def __init__(self, C=1024, out=4096, g=32):
        super().__init__()
        self.g, self.C = g, C
        self.input_conv = nn.ModuleList(nn.Conv2d(C + 2, C, 1) for _ in range(3))
        self.fuse = nn.ModuleList(nn.Sequential(nn.Conv2d(C, C, 3, padding=1, bias=False), nn.GroupNorm(64, C), nn.ReLU()) for _ in range(5))
        self.pconvs = nn.ModuleList(nn.Conv2d(C, C, 3, padding=1) for _ in range(3))
        self.flatten_linear, self.updims = lin(C * 196, 1024), lin(1024, out)
        self.pos = nn.Sequential(lin(4, 256), nn.ReLU(), nn.LayerNorm(256), lin(256, 1024), nn.ReLU(), nn.LayerNorm(1024))

    def forward(self, hs, boxes):
        B, g, C = hs[0].shape[0], self.g, self.C
        xs = []
        for l, s in enumerate((4 * g, 2 * g, g)):
            f = F.interpolate(hs[l - 3][:, 1:].reshape(B, g, g, C).permute(0, 3, 1, 2), size=(s, s), mode="bilinear", align_corners=True)
            r = torch.linspace(-1, 1, s, device=dev, dtype=f.dtype)
            yy, xx = torch.meshgrid(r, r, indexing="ij")
            xs.append(self.input_conv[l](torch.cat([f, torch.stack([xx, yy])[None].expand(B, -1, -1, -1)], 1)))
        q = C // 4
        for conv in self.fuse:
            new = []
            for l in range(3):
                s = xs[l].shape[-1]
                top, dn = xs[min(l + 1, 2)], xs[max(l - 1, 0)]
                ft = F.interpolate(top[:, 3 * q:].float(), size=(s, s), mode="bilinear", align_corners=True).to(xs[l].dtype)
                fd = F.interpolate(dn[:, 2 * q:3 * q].float(), size=(s, s), mode="bilinear", align_corners=True).to(xs[l].dtype)
                new.append(conv(torch.cat([xs[l][:, :2 * q], ft, fd], 1)))
            xs = new
        rois = torch.cat([torch.cat([torch.full((len(b), 1), float(i), device=dev), b.float() * 448], 1) for i, b in enumerate(boxes)])
        acc = 0
        for l in range(3):
            rf = torchvision.ops.roi_align(xs[l].float(), rois, 14, (8, 4, 2)[l] / 14.0, 2, True)      # RoIAlign runs in fp32 in the reference
            acc = acc + self.pconvs[l](rf.to(xs[l].dtype))
        flat = self.flatten_linear(F.relu(acc).flatten(1))
        return self.updims(flat + self.pos(torch.cat(boxes).to(flat.dtype)))


class TorchGroma(nn.Module):
    def __init__(self, with_llm=True):
        super().__init__()
        self.vit = Dinov2Model(Dinov2Config(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, mlp_ratio=4, image_size=518, patch_size=14,
                                            attn_implementation="sdpa"))
        self.bridge = nn.Sequential(lin(4096, 4096), nn.GELU(), lin(4096, 4096))
        self.proposer, self.region = Proposer(), RegionEncoder()
        self.attn = None
        if with_llm:
            try:
                import flash_attn  # noqa: F401
                self.attn = "flash_attention_2"
            except Exception:
                self.attn = "sdpa"
            self.llm = LlamaForCausalLM(LlamaConfig(hidden_size=4096, intermediate_size=11008, num_hidden_layers=32, num_attention_heads=32,
                                                    num_key_value_heads=32, vocab_size=32000, max_position_embeddings=4096, rms_norm_eps=1e-5,
                                                    attn_implementation=self.attn))
            self.new_embs = nn.Embedding(114, 4096)

    def vision(self, images):
        return self.vit(pixel_values=images, output_hidden_states=True).hidden_states[-4:]

    def select(self, pred, score, thr=0.0):
        """groma.py:251-280: python loop over the batch, torchvision nms on fp32 xyxy, randperm on the CPU generator."""
        out = []
        for i in range(pred.shape[0]):
            b = pred[i]
            xyxy = torch.cat([b[:, :2] - 0.5 * b[:, 2:], b[:, :2] + 0.5 * b[:, 2:]], -1)
            keep = torchvision.ops.nms(xyxy, score[i], 0.6)
            keep = keep[score[i][keep] > thr][:100]
            out.append(b[keep][torch.randperm(len(keep))] if len(keep) > 0 else b[score[i].argmax()][None])
        return out

    @torch.inference_mode()
    def generate(self, images, ids, new=128):
        hs = self.vision(images)
        B, g = images.shape[0], 32
        f = hs[-1][:, 1:].reshape(B, g, g, 1024)
        img_tok = self.bridge(torch.cat([f[:, 0::2, 0::2], f[:, 1::2, 0::2], f[:, 0::2, 1::2], f[:, 1::2, 1::2]], -1).reshape(B, 256, 4096))
        pred, score = self.proposer(hs)
        boxes = self.select(pred, score)
        reg = self.region(hs, boxes)
        # sequence assembly + splice (groma.py:317-369): 256 <image> slots, R x (<r_j>, <region>) slots, text elsewhere
        emb, off = [], 0
        for i in range(B):
            R = len(boxes[i])
            t = self.llm.model.embed_tokens(ids[i])
            r = torch.stack([self.new_embs.weight[14:14 + R], reg[off:off + R]], 1).reshape(2 * R, -1)
            emb.append(torch.cat([t[:4], img_tok[i], t[5:256], r, t[257:]]))
            off += R
        emb = nn.utils.rnn.pad_sequence(emb, batch_first=True)
        mask = torch.ones(emb.shape[:2], dtype=torch.long, device=dev)
        return self.llm.generate(inputs_embeds=emb, attention_mask=mask, max_new_tokens=new, min_new_tokens=new, do_sample=False, use_cache=True,
                                 pad_token_id=0), emb.shape[1]


def timed(fn, iters=3, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


def main():
    which = [int(a) for a in sys.argv[1:]] or [2, 3, 4]
    torch.manual_seed(0)
    out = {"torch": torch.__version__, "transformers": __import__("transformers").__version__, "torchvision": torchvision.__version__}
    m = TorchGroma(with_llm=4 in which).to(dev).to(BF).eval()
    out["llama_attn_implementation"] = m.attn
    g = torch.Generator().manual_seed(0)
    with torch.inference_mode():
        if 2 in which:
            img = torch.randn(64, 3, 448, 448, device=dev, dtype=BF)
            ms = timed(lambda: m.vision(img))
            out["config2_vit_B64"] = {"ms": ms, "images_per_s": 64 / ms * 1e3}
        if 3 in which:
            img = torch.randn(32, 3, 448, 448, device=dev, dtype=BF)
            hs = m.vision(img)
            boxes = [(torch.rand(100, 4, generator=g) * 0.6 + 0.2).to(dev) for _ in range(32)]
            ms = timed(lambda: (m.proposer(hs), m.region(hs, boxes)))
            out["config3_region_tokenizer_B32_R100"] = {"ms": ms, "images_per_s": 32 / ms * 1e3, "proposer_ms": timed(lambda: m.proposer(hs))}
        if 4 in which:
            img = torch.randn(16, 3, 448, 448, device=dev, dtype=BF)
            ids = torch.randint(1000, 32000, (16, 512), device=dev)
            T = [0]

            def step():
                torch.manual_seed(0)
                seq, T[0] = m.generate(img, ids, 128)
                return seq
            ms = timed(step, iters=2, warm=1)
            out["config4_e2e_B16"] = {"ms": ms, "images_per_s": 16 / ms * 1e3, "prefill_tokens": T[0], "new_tokens": 128}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
