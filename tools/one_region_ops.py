"""RoIAlign (bench shape: 16 images, 1600 RoIs with Groma's cxcywh-fed-as-xyxy quirk, level 0 = 128x128x1024 maps) and the MSDA sampling
kernel (B=32, encoder shape) as single launches for `ncu --set full`:   python tools/one_region_ops.py"""
import sys, torch
sys.path.insert(0, ".")
from groma_b200 import ops as G
torch.manual_seed(0)
B, C, R = 16, 1024, 100
g = torch.Generator().manual_seed(0)
boxes = torch.rand(B * R, 4, generator=g) * 0.6 + 0.2          # cxcywh in (0,1), as the bench's proposals look after NMS
img = torch.arange(B).repeat_interleave(R).float()[:, None]
rois = torch.cat([img, boxes * 448.0], 1).cuda().contiguous()   # fed as xyxy
for s, scale in ((128, 8 / 14.0), (64, 4 / 14.0), (32, 2 / 14.0)):
    feat = torch.randn(B, s, s, C, device="cuda").bfloat16()
    out = torch.empty(B * R, 16, 16, C, device="cuda", dtype=torch.bfloat16)
    for _ in range(2):
        G.roi_align(feat, rois, 14, scale, 2, True, pad=True, out=out)
    torch.cuda.synchronize()
    if "--time" in sys.argv:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ts = []
        for _ in range(10):
            e0.record(); G.roi_align(feat, rois, 14, scale, 2, True, pad=True, out=out); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        wr = out.numel() * 2
        print(f"roi_align {s}x{s}x{C}: best {min(ts):.1f} us  median {sorted(ts)[5]:.1f} us  "
              f"(output {wr / 1e6:.0f} MB -> {wr / min(ts) / 1e6:.2f} TB/s written)", flush=True)
Bm, Q = 32, 1024
value = torch.randn(Bm, 1024, 8, 32, device="cuda").bfloat16()
proj = torch.randn(Bm * Q, 96, device="cuda")
ref = torch.rand(Bm, Q, 2, device="cuda")
o = torch.empty(Bm, Q, 256, device="cuda", dtype=torch.bfloat16)
for _ in range(2):
    G.msda(value, proj, ref, [(32, 32)], 8, 4, out=o)
torch.cuda.synchronize()
