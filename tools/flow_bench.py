"""CNN stage time of the dataflow plan vs the per-layer plan at the headline workload (GPU box).

usage: python tools/flow_bench.py [V] [S] [--cfg=lo,hi,tiles,k ...]   (several --cfg allowed)
"""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from mvlm_b200 import build, ops  # noqa: E402
from mvlm_b200.weights import seeded_state_dict  # noqa: E402

build.build()
pos = [a for a in sys.argv[1:] if not a.startswith("--")]
V = int(pos[0]) if len(pos) > 0 else 100
S = int(pos[1]) if len(pos) > 1 else 256
cfgs = [a.split("=")[1] for a in sys.argv[1:] if a.startswith("--cfg=")] or ["1,32,64,3"]
sd = seeded_state_dict(73, "RGB+depth", 1234)
img = torch.randint(0, 256, (V, S, S, 4), dtype=torch.uint8, device="cuda")


def timed(net, reps=10):
    for _ in range(3):
        net.forward(img, graph=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        pk, _ = net.forward(img, graph=True)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, pk.clone()


os.environ["MVLM_FLOW"] = "0"
ref = ops.Hourglass(sd, 73, 4, V, S, S)
t_ref, pk_ref = timed(ref)
print(f"per-layer plan: {t_ref:.3f} ms  ({ref.num_launches} ops, workspace {ref.workspace.numel() / 1e9:.2f} GB)", flush=True)
del ref
torch.cuda.empty_cache()
for cfg in cfgs:
    lo, hi, tiles, k = cfg.split(",")
    os.environ.update({"MVLM_FLOW": "1", "MVLM_FLOW_LO": lo, "MVLM_FLOW_HI": hi, "MVLM_FLOW_TILES": tiles, "MVLM_FLOW_K": k})
    net = ops.Hourglass(sd, 73, 4, V, S, S)
    t, pk = timed(net)
    same = torch.equal(pk.view(torch.int32), pk_ref.view(torch.int32))
    print(f"dataflow plan rows {lo}..{hi} tiles={tiles} k={k}: {t:.3f} ms  ({net.num_segments} segments, workspace "
          f"{net.workspace.numel() / 1e9:.2f} GB)  peaks identical: {same}", flush=True)
    del net
    torch.cuda.empty_cache()
