"""Experiment: tensor-pipe + smem-read throughput of the conv kernel without TMA traffic (debug mode 1)."""
import ctypes as C
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from mvlm_b200 import _lib, build, ops  # noqa: E402

build.build()
lib = _lib.load()
lib.mvlm_debug_conv_profile.argtypes = [C.c_void_p]
lib.mvlm_debug_conv_mode.argtypes = [C.c_int]
v, h = 100, 128
for cin, cout, nt in ((256, 128, 128),):
    x = torch.randn((v, h, h, cin), device="cuda").to(torch.bfloat16)
    w = torch.randn((cout, cin, 3, 3), device="cuda") / 48
    wp = ops.pack_conv_weight(w, cout, cin)
    out = torch.zeros((v, h, h, cout), device="cuda", dtype=torch.bfloat16)
    for mode in (0, 1, 0, 1):
        buf = torch.zeros((148 + 128, 8), dtype=torch.int64, device="cuda")  # role counters + CTA-0 tile timeline
        lib.mvlm_debug_conv_mode(mode)
        lib.mvlm_debug_conv_profile(buf.data_ptr())
        ops.conv2d_bf16(x, wp, n_tile=nt, out_raw=(out, 0))
        torch.cuda.synchronize()
        lib.mvlm_debug_conv_profile(None)
        lib.mvlm_debug_conv_mode(0)
        b = buf[:148].double().mean(0).cpu().numpy()
        n_mma = v * (h // 16) ** 2 / 148 * (cin // 64) * 9 * 4
        print(f"{cin}->{cout}@{h}^2 mode={mode}: mma total {b[4] / 1e3:8.1f} kcyc, wait operands {b[2] / 1e3:8.1f}, wait acc {b[3] / 1e3:7.1f} "
              f"-> {b[4] / n_mma:6.1f} cycles per MMA over the whole run, {(b[4] - b[2] - b[3]) / n_mma:6.1f} without the timed waits (floor 128)")
