"""Quick GPU probe of the target decoder: error against the recorded reference vectors + timing."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "embodied-active-learning-vision_b200"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
from cases import TARGET_CASES, DecoderModel, target_samples  # noqa: E402
from control_torch.target_decoder import DeviceTarget  # noqa: E402

for name in TARGET_CASES:
    case = TARGET_CASES[name]
    g = np.load(os.path.join(ROOT, "tests", "golden", f"target_{name}.npz"))
    model = DecoderModel(case, g["z_rows"], g["seed_x"], out_extra=int(g["out_features"]) - case["nl"])
    dev = DeviceTarget(model)
    p = dev.pdf_torch(target_samples(case)).cpu().numpy()
    torch.cuda.synchronize()
    err = np.abs(p - g["p"]) / np.abs(g["p"])
    print(name, "max rel err", err.max(), "fault", int(dev._fault.item()), "first", p[:4], g["p"][:4], flush=True)

if len(sys.argv) > 1:
    n = int(float(sys.argv[1]))
    case = dict(TARGET_CASES["default"], n=n)
    model = DecoderModel(case, torch.randn(1, 16))
    dev = DeviceTarget(model)
    s = torch.rand(n, 3, device="cuda") * 2 - 1
    for _ in range(3):
        p = dev.pdf_torch(s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        p = dev.pdf_torch(s)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    flop = 2.0 * n * 256 * 512 * 3
    print(f"n={n}: {ms:.3f} ms per pdf (pack included), {flop / ms / 1e9:.1f} TFLOP/s tf32 issued, {n / ms / 1e3:.1f} Msamples/s", flush=True)
