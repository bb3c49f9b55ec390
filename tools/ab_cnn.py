"""A/B timing of the CNN stage for two (or more) builds of libmvlm_b200.so on ONE box, alternating, fresh process each
(boxes differ by several per cent in sustained clocks, so numbers from different gpurun calls do not compare).

usage: python tools/ab_cnn.py libA.so libB.so [--rounds=3] [--views=100] [--size=256]
       python tools/ab_cnn.py --child   (internal)
"""
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def child():
    import torch
    from mvlm_b200 import ops
    from mvlm_b200.weights import seeded_state_dict
    v, s = int(os.environ.get("AB_VIEWS", "100")), int(os.environ.get("AB_SIZE", "256"))
    net = ops.Hourglass(seeded_state_dict(73, "RGB+depth", 1234), 73, 4, v, s, s)
    img = torch.randint(0, 256, (v, s, s, 4), dtype=torch.uint8, device="cuda")
    for _ in range(5):
        net.forward(img, graph=True)
    torch.cuda.synchronize()
    best, tot = 1e9, 0.0
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            net.forward(img, graph=True)
        e1.record()
        torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / 10
        best = min(best, t)
        tot += t
    print(f"{best:.4f} {tot / 3:.4f}")


if "--child" in sys.argv:
    child()
    sys.exit(0)
libs = [a for a in sys.argv[1:] if not a.startswith("--")]
opt = {a.split("=")[0][2:]: a.split("=")[1] for a in sys.argv[1:] if a.startswith("--") and "=" in a}
rounds = int(opt.get("rounds", "3"))
res = {lib: [] for lib in libs}
for r in range(rounds):
    for lib in libs:
        env = dict(os.environ, MVLM_B200_LIB=str(Path(lib).resolve()), AB_VIEWS=opt.get("views", "100"), AB_SIZE=opt.get("size", "256"))
        out = subprocess.run([sys.executable, __file__, "--child"], env=env, capture_output=True, text=True)
        if out.returncode != 0:
            print(lib, "FAILED", out.stderr[-500:])
            continue
        best, mean = (float(x) for x in out.stdout.split()[-2:])
        res[lib].append((best, mean))
        print(f"round {r} {lib}: best {best:.3f} ms  mean {mean:.3f} ms", flush=True)
for lib in libs:
    if res[lib]:
        print(f"{lib}: best-of-rounds {min(b for b, _ in res[lib]):.3f} ms, mean {sum(m for _, m in res[lib]) / len(res[lib]):.3f} ms")
