"""Experiment: cycles per MMA inside the conv kernel for the cout <= 64 (.ws) layers against the 128-channel ones.
debug modes: 0 = normal, 1 = no TMA traffic / operand waits skipped, 2 = epilogue dropped, 3 = both."""
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
v = 100
for h, cin, cout in ((256, 64, 64), (256, 32, 32), (256, 16, 64)):
    x = torch.randn((v, h, h, cin), device="cuda").to(torch.bfloat16)
    w = torch.randn((cout, cin, 3, 3), device="cuda") / 48
    wp = ops.pack_conv_weight(w, cout, cin)
    out = torch.zeros((v, h, h, cout), device="cuda", dtype=torch.bfloat16)
    for mode in (0, 3, 7):
        buf = torch.zeros((148 + 128, 8), dtype=torch.int64, device="cuda")  # role counters + CTA-0 tile timeline
        lib.mvlm_debug_conv_mode(mode)
        lib.mvlm_debug_conv_profile(buf.data_ptr())
        ops.conv2d_bf16(x, wp, n_tile=128, out_raw=(out, 0))
        torch.cuda.synchronize()
        lib.mvlm_debug_conv_profile(None)
        lib.mvlm_debug_conv_mode(0)
        b = buf[:148].double().mean(0).cpu().numpy()
        tiles = v * (h // 8) * (h // 32) * max(1, cout // 128) / 148
        n_mma = tiles * ((cin + 63) // 64) * 9 * min(4, cin // 16)
        print(f"{cin}->{cout}@{h}^2 mode={mode}: mma total {b[4] / 1e3:8.1f} kcyc, wait operands {b[2] / 1e3:8.1f}, wait acc {b[3] / 1e3:7.1f} "
              f"-> {b[4] / n_mma:6.1f} cycles per MMA, {(b[4] - b[2] - b[3]) / n_mma:6.1f} without the timed waits; {b[4] / tiles:7.0f} cycles per tile")
