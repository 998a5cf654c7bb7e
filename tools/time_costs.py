"""Time of one fused cost launch vs number of candidates (BASELINE config 3 shape)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "embodied-active-learning-vision_b200"), os.path.join(ROOT, "tests")]
import torch
import workloads as wl
from test_gpu_full_size import build

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
s = build("c3", n, 3_000)
ctx = s["ctx_for"](s["samples"], s["p_raw"], n)
ctx.set_history(s["hist"])
U = wl.random_controls((8, s["H"], s["D"]), seed=1).to(s["dev"])
for g in (1, 2, 4, 8):
    for _ in range(3):
        ctx.costs(U[:g])
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        ctx.costs(U[:g])
    b.record()
    torch.cuda.synchronize()
    t = a.elapsed_time(b) / 20 * 1e3
    print(f"G={g}: {t:.1f} us per launch, {g * s['H'] * n / t / 1e6:.3f}e12 forward pairs/s")
for _ in range(3):
    ctx.gradient(U[0])
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20):
    ctx.gradient(U[0])
b.record()
torch.cuda.synchronize()
print(f"gradient eval: {a.elapsed_time(b) / 20 * 1e3:.1f} us")
