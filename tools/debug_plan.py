import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "embodied-active-learning-vision_b200")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from types import SimpleNamespace
import control_torch.klerg as kk
from tests.test_gpu_parity import make_robot
ct = SimpleNamespace(kk=kk)
orig_select = kk.line_search_select
def traced(costs, windows, idx, lam0, J0):
    out = orig_select(costs, windows, idx, lam0, J0)
    print("  host select idx", idx, "J0", float(J0), "Js", [float(c) for c in costs], "windows", windows, "->", out[0], out[1], out[3])
    return out
kk.line_search_select = traced
for dl in (False, True):
    r, case = make_robot(ct, "c1_xy")
    r.device_loop = dl
    for k in range(2):
        r.step(case["n"], case["m"], save_update=True)
        print("device_loop", dl, "step", k, "last_cost", float(r.last_cost), "stats", r.stats)
        if dl:
            print("  pack head", r.ctx.buf.plan_result[:8].tolist())
            import ctypes
            raw = r.ctx.buf.plan_scratch[:96].cpu().numpy().view(np.int32)
            print("  state", raw[:24].tolist(), "last_cost", raw[:1].view(np.float32))
            cp = r.ctx.buf.plan_scratch.cpu().numpy()
