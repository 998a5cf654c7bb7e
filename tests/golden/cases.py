"""Synthetic workloads shared by make_golden.py (reference side) and the parity tests."""
import numpy as np
import torch


class MixtureTarget:
    """Synthetic target density: diagonal Gaussian mixture + 1e-5."""

    def __init__(self, dim, n_comp=3, seed=1, lo=-0.8, hi=0.8):
        g = torch.Generator().manual_seed(seed)
        self.mu = torch.rand(n_comp, dim, generator=g) * (hi - lo) + lo
        self.var = torch.rand(n_comp, dim, generator=g) * 0.1 + 0.02
        self.device = "cpu"
        self.dtype = torch.float32

    def pdf_torch(self, x):
        d = x.unsqueeze(1) - self.mu.unsqueeze(0)
        return torch.exp(-0.5 * (d * d / self.var.unsqueeze(0)).sum(2)).sum(1) + 1e-5

    def init_uniform_grid(self, x):
        v = torch.ones(x.shape[0])
        v /= v.sum()
        return v + 1e-5


ROBOT_CASES = {
    # name: ctor kwargs, (n_target, n_hist), n_steps
    "c1_xy": dict(states="xy", D=2, horizon=20, cap=500, n=1000, m=500, steps=12, std=0.05, x0=[0.1, -0.2, 0.0, 0.0]),
    "xyz_small": dict(states="xyz", D=3, horizon=12, cap=64, n=600, m=40, steps=10, std=0.08, x0=[0.3, 0.2, -0.1, 0, 0, 0]),
    "xyw_plot": dict(states="xyw", D=3, horizon=10, cap=50, n=400, m=30, steps=8, std=0.08, x0=[0.0, 0.5, 0.2, 0, 0, 0], plot=True),
    "xyzrpw": dict(states="xyzrpw", D=6, horizon=10, cap=40, n=500, m=30, steps=8, std=0.3,
                   x0=[0.1, 0.0, -0.2, 3.0, 0.1, 0.3, 0, 0, 0, 0, 0, 0]),
    "xyXY_vel": dict(states="xyXY", D=4, horizon=10, cap=40, n=500, m=30, steps=8, std=0.1, x0=[0.2, 0.1, 0.0, 0.0], vel=True),
    "xy_weightenv": dict(states="xy", D=2, horizon=10, cap=40, n=300, m=30, steps=8, std=0.05, x0=[0.5, 0.5, 0, 0], weight_env=True),
    "xy_uniform": dict(states="xy", D=2, horizon=10, cap=40, n=300, m=30, steps=6, std=0.05, x0=[-0.5, 0.5, 0, 0], uniform=True),
}

LIMS = dict(x=[-1.0, 1.0], y=[-1.0, 1.0], z=[-1.0, 1.0], r=[2.39, 3.89], p=[-0.75, 0.75], w=[-2.0, 2.0])
CTRL = dict(x=[-1.25, 1.25], y=[-1.25, 1.25], z=[-1.25, 1.25], r=[-0.5, 0.5], p=[-0.5, 0.5], w=[-1.25, 1.25])


def robot_kwargs(case, target):
    st = case["states"]
    pos = [s for s in st if s == s.lower()]
    lim = [LIMS[s] for s in pos]
    ctrl = [CTRL[s] for s in pos]
    if case.get("vel"):
        lim = lim + [CTRL[s.lower()] for s in st if s != s.lower()]
    return dict(
        x0=np.array(case["x0"], dtype=np.float64), robot_lim=np.array(lim), explr_idx=list(range(len(st))),
        explr_robot_lim_scale=1.15, target_dist=target, dt=0.2, R=0.5, horizon=case["horizon"],
        buffer_capacity=case["cap"], std=case["std"], std_plot=case["std"], states=st,
        plot_states=st[:2], tray_lim=np.array(lim), robot_ctrl_lim=np.array(ctrl),
        plot_data=bool(case.get("plot")), uniform_tdist=bool(case.get("uniform")), vel_states=bool(case.get("vel")),
    )


def seed_buffer_states(state0, case):
    """Short random walk pushed into the buffer before the first step."""
    g = torch.Generator().manual_seed(5)
    n = case["D"] if not case.get("vel") else 2
    rows = []
    for _ in range(5):
        s = state0.clone()
        s[:n] += 0.05 * torch.randn(n, generator=g)
        rows.append(s)
    return rows
