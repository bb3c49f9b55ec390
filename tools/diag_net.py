"""Diagnostic: where does the CUDA CNN diverge from the bf16-emulating oracle?  (GPU box)"""
import sys
from pathlib import Path

import torch
import torch.nn.functional as F

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from mvlm_b200 import build, ops  # noqa: E402
from mvlm_b200.weights import seeded_state_dict  # noqa: E402
from oracle.hourglass_ref import HourglassOracle  # noqa: E402

build.build()
torch.backends.cudnn.allow_tf32 = False
# 1) single conv accumulate precision (fp32 out), K = 2304
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn((2, 32, 32, 256), generator=g, device="cuda").to(torch.bfloat16)
w = torch.randn((128, 256, 3, 3), generator=g, device="cuda") / 48.0
wp = ops.pack_conv_weight(w, 128, 256)
out = torch.zeros((2, 128, 32, 32), device="cuda")
ops.conv2d_bf16(x, wp, n_tile=128, out_f32=out, cout_real=128)
ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.to(torch.bfloat16).float(), padding=1)
ref64 = F.conv2d(x.double().permute(0, 3, 1, 2), w.to(torch.bfloat16).double(), padding=1)
print("conv K=2304: scale", ref.abs().max().item(), "std", ref.std().item())
print("  cuda vs fp64:   max %.3e mean %.3e" % ((out.double() - ref64).abs().max().item(), (out.double() - ref64).abs().mean().item()))
print("  cudnn32 vs fp64: max %.3e mean %.3e" % ((ref.double() - ref64).abs().max().item(), (ref.double() - ref64).abs().mean().item()))

# 2) network probes
sd = seeded_state_dict(73, "RGB+depth", 1234)
gen = torch.Generator().manual_seed(5)
img_u8 = torch.randint(0, 256, (2, 64, 64, 4), generator=gen, dtype=torch.uint8)
net = ops.Hourglass(sd, 73, 4, 2, 64, 64, keep_probes=True)
peaks, hm = net.forward(img_u8.cuda(), want_heatmaps=True)
xin = (img_u8.float() / 255).permute(0, 3, 1, 2).contiguous()
r32, i32 = HourglassOracle(sd).forward(xin, return_intermediates=True)
r16, i16 = HourglassOracle(sd, emulate_bf16=True).forward(xin, return_intermediates=True)
for name in ("r3", "hg1", "sum_temp", "x10"):
    got = net.probe(name).float().permute(0, 3, 1, 2).cpu()[:, : i16[name].shape[1]]
    s = i32[name].std().item()
    print(f"{name:9s} std {s:7.3f}  cuda-emu mean {((got - i16[name]).abs().mean() / s).item():.5f} max {((got - i16[name]).abs().max() / s).item():.5f}"
          f" | emu-fp32 mean {((i16[name] - i32[name]).abs().mean() / s).item():.5f} | cuda-fp32 mean {((got - i32[name]).abs().mean() / s).item():.5f}"
          f" | frac elems differing cuda-emu {(got != i16[name]).float().mean().item():.4f}")
s = r32.std().item()
hm = hm.cpu()
print(f"heatmap   std {s:7.3f}  cuda-emu mean {((hm - r16).abs().mean() / s).item():.5f} | emu-fp32 {((r16 - r32).abs().mean() / s).item():.5f} | cuda-fp32 {((hm - r32).abs().mean() / s).item():.5f}")
