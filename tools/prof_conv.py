"""Runs a few launches of one conv shape (for `ncu --set full -k regex:conv_umma`).  GPU box only.
usage: python tools/prof_conv.py H CIN COUT NTILE [V]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from mvlm_b200 import build, ops  # noqa: E402

build.build()
h, cin, cout, nt = (int(x) for x in sys.argv[1:5])
v = int(sys.argv[5]) if len(sys.argv) > 5 else 100
k = int(sys.argv[6]) if len(sys.argv) > 6 else 3
epi = sys.argv[7] if len(sys.argv) > 7 else "rb"
x = torch.randn((v, h, h, cin), device="cuda").to(torch.bfloat16)
w = torch.randn((cout, cin, k, k), device="cuda") / (cin * k * k) ** 0.5
wp = ops.pack_conv_weight(w, cout, cin)
big = torch.zeros((v, h, h, 256), device="cuda", dtype=torch.bfloat16)
act = torch.zeros((v, h, h, max(cout, 64)), device="cuda", dtype=torch.bfloat16)
s = torch.ones(cout, device="cuda")
t = torch.zeros(cout, device="cuda")
keys = torch.zeros((v * cout,), device="cuda", dtype=torch.int64)
if epi == "raw":
    kw = dict(out_raw=(big, 0))
elif epi == "head":
    kw = dict(argmax_keys=keys, cout_real=cout - 7, up=(2, 2, 0, 1), y_off0=-1, x_off0=0)
elif epi == "res":
    kw = dict(res1=(big, 0), out_raw=(big, 0))
else:
    kw = dict(pre=(s, t, act, 0), res1=(big, 0), out_raw=(big, 0), post=(s, t, big, 0))
for _ in range(3):
    ops.conv2d_bf16(x, wp, n_tile=nt, kh=k, kw=k, **kw)
torch.cuda.synchronize()
print("ok")

# role timing (mvlm_debug_conv_profile)
import ctypes as C  # noqa: E402
from mvlm_b200 import _lib  # noqa: E402
lib = _lib.load()
buf = torch.zeros((148 + 128, 8), dtype=torch.int64, device="cuda")  # role counters + CTA-0 tile timeline
lib.mvlm_debug_conv_profile.argtypes = [C.c_void_p]
lib.mvlm_debug_conv_profile(buf.data_ptr())
ops.conv2d_bf16(x, wp, n_tile=nt, kh=k, kw=k, **kw)
torch.cuda.synchronize()
lib.mvlm_debug_conv_profile(None)
b = buf[:148].double().mean(0).cpu().numpy()
names = ["prod wait A-empty", "prod wait B-empty", "mma wait operands", "mma wait acc-free", "mma total", "epi wait acc-full", "epi total", "prod total"]
for n, v in zip(names, b):
    print(f"{n:20s} {v / 1e3:10.1f} kcycles")
tiles = v * (h // 16) ** 2 * max(1, cout // 128)
print("tiles/CTA", tiles / 148, "cycles/tile", b[4] / (tiles / 148))
