"""Micro-benchmark of the tcgen05 conv kernel on the CNN's dominant shapes (V views batched).

Usage (GPU box): python tools/bench_conv.py [V]
Prints per shape: ms, algorithmic TFLOP/s (2*Co*Ci*kh*kw*H*W*V) and fraction of the measured
sustained bf16 peak.  CUDA events on the launching stream, 3 warm-ups, 10 timed launches;
inputs (>= 100 MB at V=100) exceed... small shapes are L2-resident by design of the network.
"""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from mvlm_b200 import build, ops  # noqa: E402

SHAPES = [
    # name, H, Cin, Cout, n_tile, k, epilogue
    ("256->256@128 conv5", 128, 256, 256, 128, 3, "pre"),
    ("256->128@128 rb.conv1", 128, 256, 128, 128, 3, "rb"),
    ("256->128@64 rb.conv1", 64, 256, 128, 128, 3, "rb"),
    ("128->64@128 rb.conv2", 128, 128, 64, 64, 3, "rb"),
    ("64->64@128 rb.conv3", 128, 64, 64, 64, 3, "rb"),
    ("256->80@128 conv6", 128, 256, 80, 80, 3, "raw"),
    ("80->256@128 conv7", 128, 80, 256, 128, 3, "raw"),
    ("64->64@256 c2.conv1", 256, 64, 64, 64, 3, "rb"),
    ("256->128@32", 32, 256, 128, 128, 3, "rb"),
    ("256->128@16", 16, 256, 128, 128, 3, "rb"),
    ("256->128@8", 8, 256, 128, 128, 3, "rb"),
    ("256->128@4", 4, 256, 128, 128, 3, "rb"),
    ("64->128@256 1x1", 256, 64, 128, 128, 1, "raw"),
]


def main():
    build.build()
    v = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    peaks = {}
    pk = Path(__file__).resolve().parents[1] / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    peak = peaks.get("bf16_tflops_sustained", 1386.9)
    rows = []
    for name, h, cin, cout, n_tile, k, epi in SHAPES:
        x = torch.randn((v, h, h, cin), device="cuda").to(torch.bfloat16)
        w = torch.randn((cout, cin, k, k), device="cuda") / (cin * k * k) ** 0.5
        wp = ops.pack_conv_weight(w, cout, cin)
        big = torch.zeros((v, h, h, 256), device="cuda", dtype=torch.bfloat16)
        act = torch.zeros((v, h, h, max(cout, 64)), device="cuda", dtype=torch.bfloat16)
        s = torch.ones(cout, device="cuda")
        t = torch.zeros(cout, device="cuda")
        kw = dict(n_tile=n_tile, kh=k, kw=k)
        if epi == "raw":
            kw.update(out_raw=(big, 0))
        elif epi == "pre":
            kw.update(pre=(s, t, big, 0))
        else:
            kw.update(pre=(s, t, act, 0), res1=(big, 0), out_raw=(big, 0), post=(s, t, big, 0))
        for _ in range(3):
            ops.conv2d_bf16(x, wp, **kw)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        iters = 10
        e0.record()
        for _ in range(iters):
            ops.conv2d_bf16(x, wp, **kw)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        flop = 2.0 * cout * cin * k * k * h * h * v
        tf = flop / ms / 1e9
        rows.append((name, ms, tf, tf / peak))
        print(f"{name:28s} {ms:8.3f} ms  {tf:8.1f} TFLOP/s  {100 * tf / peak:5.1f}% of sustained peak", flush=True)
        del x, big, act
    return rows


if __name__ == "__main__":
    main()
