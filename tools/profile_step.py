"""cProfile of Robot.step() at a small workload (host overhead of the planner loop)."""
import cProfile, os, pstats, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "embodied-active-learning-vision_b200")]
import torch
import workloads as wl
from control_torch.klerg import Robot

name = sys.argv[1] if len(sys.argv) > 1 else "c1"
w = wl.WORKLOADS[name]
lims = [wl.LIMS[s] for s in w["states"]]
target = wl.make_target(w["target"], lims, seed=1, device="cuda")
kw = wl.robot_kwargs(name, target)
torch.manual_seed(7)
r = Robot(**kw)
r.test(1000)
for row in wl.random_walk_history(name, min(w["M"], r.memory_buffer.capacity), seed=5):
    r.memory_buffer.push(row)
for _ in range(5):
    r.step(w["N"], w["M"], save_update=True)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(30):
    r.step(w["N"], w["M"], save_update=True)
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(28)
