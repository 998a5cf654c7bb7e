"""Synthetic workloads of BASELINE.json / SURVEY.md section 8(d), shared by bench.py and tests.

Nothing here is on the hot path: these build *inputs* (limits, history states,
control sequences, target densities) with fixed seeds.
"""
import math

import numpy as np
import torch

LIMS = dict(x=[-1.0, 1.0], y=[-1.0, 1.0], z=[-1.0, 1.0], r=[2.39, 3.89], p=[-0.75, 0.75], w=[-2.0, 2.0])
CTRL = dict(x=[-1.25, 1.25], y=[-1.25, 1.25], z=[-1.25, 1.25], r=[-0.5, 0.5], p=[-0.5, 0.5], w=[-1.25, 1.25])

# name -> states, horizon, samples, history rows drawn, buffer capacity, candidates, targets
WORKLOADS = {
    "c1": dict(states="xy", H=20, N=1_000, M=500, cap=500, B=1, K=1, target="gmm"),
    "c2": dict(states="xyz", H=50, N=100_000, M=3_000, cap=3_000, B=1, K=1, target="vae"),
    "c3": dict(states="xyz", H=50, N=1_000_000, M=3_000, cap=3_000, B=1024, K=1, target="gmm"),
    "c4": dict(states="xyzrpw", H=50, N=10_000_000, M=100_000, cap=100_000, B=1, K=1, target="gmm"),
    "c5": dict(states="xyz", H=50, N=1_000_000, M=3_000, cap=3_000, B=1, K=16, target="gmm"),
}


def std_from_ratio(lims, n_samples, ratio=0.1):
    """Kernel scale rule of the reference's load_config.get_std (scripts/load_config.py:131-138)."""
    lims = np.asarray(lims, dtype=np.float64)
    n = lims.shape[0]
    vol = np.prod(lims[:, 1] - lims[:, 0])
    return float((ratio / n_samples * vol * math.gamma(n / 2 + 1) / math.pi ** (n / 2)) ** (1 / n))


class MixtureTarget:
    """Diagonal Gaussian-mixture target density + 1e-5 (duck type of the reference's target_dist)."""

    def __init__(self, lims, n_comp=3, seed=1, device="cpu"):
        g = torch.Generator().manual_seed(seed)
        lims = torch.as_tensor(np.asarray(lims), dtype=torch.float32)
        span = lims[:, 1] - lims[:, 0]
        self.mu = (lims[:, 0] + span * (0.1 + 0.8 * torch.rand(n_comp, len(lims), generator=g))).to(device)
        self.var = ((0.02 + 0.1 * torch.rand(n_comp, len(lims), generator=g)) * (span / 2) ** 2).to(device)
        self.device = device
        self.dtype = torch.float32

    def pdf_torch(self, x):
        d = x.to(self.device).unsqueeze(1) - self.mu.unsqueeze(0)
        return torch.exp(-0.5 * (d * d / self.var.unsqueeze(0)).sum(2)).sum(1) + 1e-5

    def init_uniform_grid(self, x):
        v = torch.ones(x.shape[0], device=self.device)
        v /= v.sum()
        return v + 1e-5


class SyntheticVAETarget:
    """Random-init stand-in for the reference's CVAE as the controller sees it (scripts/vae/vae.py:11-110,244-275):
    the attributes ``pdf_torch`` reads - ``decode`` (Linear z+s -> 256, ReLU, Linear 256 -> 512, ReLU, Linear
    512 -> 1+feat), ``z_samples``, ``ylogvar_dim``, ``logvar_lims``, ``init`` - and a torch-CPU ``pdf_torch`` with
    the reference's operators for the CPU legs.  On a GPU the controller wraps it in ``DeviceTarget``."""

    def __init__(self, s_dim, z_dim=16, hidden=(512, 256), seed=0, feat=8):
        g = torch.Generator().manual_seed(seed)
        dims = [z_dim + s_dim] + list(reversed(hidden)) + [1 + feat]
        layers = []
        for i, (a, b) in enumerate(zip(dims[:-1], dims[1:])):
            bound = 1.0 / math.sqrt(a)
            lin = torch.nn.Linear(a, b)
            with torch.no_grad():
                lin.weight.copy_((torch.rand(b, a, generator=g) * 2 - 1) * bound * 3.0)
                lin.bias.copy_((torch.rand(b, generator=g) * 2 - 1) * bound)
            layers.append(lin)
            if i + 2 < len(dims):
                layers.append(torch.nn.ReLU())
        self.decode = torch.nn.Sequential(*layers).requires_grad_(False)
        self.z_samples = torch.randn(1, z_dim, generator=g)
        self.ylogvar_dim = torch.tensor(1)
        self.logvar_lims = (-10, 2)
        self.init = torch.tensor([True])
        self.dx = False
        self.use_buffer = False
        self.device = "cpu"
        self.dtype = torch.float32

    @torch.no_grad()
    def pdf_torch(self, x):
        x = x.to(device="cpu", dtype=torch.float32)
        latent = torch.cat([self.z_samples.repeat(x.shape[0], 1), x], dim=1)
        y = self.decode(latent)[:, :int(self.ylogvar_dim)]
        return torch.amax(torch.exp(torch.clamp(y, *self.logvar_lims)), 1).squeeze()

    def init_uniform_grid(self, x):
        return x.sum(1) ** 0


def make_target(kind, lims, seed=1, device="cpu"):
    if kind == "vae":
        model = SyntheticVAETarget(len(lims), seed=seed)
        if torch.device(device).type == "cuda":
            from control_torch.target_decoder import DeviceTarget
            return DeviceTarget(model, device)
        return model
    return MixtureTarget(lims, seed=seed, device=device)


def robot_kwargs(name, target, n_samples=None, horizon=None, cap=None):
    w = WORKLOADS[name]
    st = w["states"]
    lim = [LIMS[s] for s in st]
    ctrl = [CTRL[s] for s in st]
    n = n_samples or w["N"]
    x0 = np.array([0.5 * (a + b) for a, b in lim] + [0.0] * len(st))
    return dict(x0=x0, robot_lim=np.array(lim), explr_idx=list(range(len(st))), explr_robot_lim_scale=1.15,
                target_dist=target, dt=0.2, R=0.5, horizon=horizon or w["H"], buffer_capacity=cap or w["cap"],
                std=std_from_ratio(lim, n), std_plot=std_from_ratio(lim, n), states=st, plot_states=st[:2],
                tray_lim=np.array(lim), robot_ctrl_lim=np.array(ctrl))


def random_walk_history(name, rows, seed=0):
    """History states: a bounded random walk of the double integrator inside the limits, [rows, 2D]."""
    st = WORKLOADS[name]["states"]
    lim = torch.tensor([LIMS[s] for s in st])
    g = torch.Generator().manual_seed(seed)
    D = len(st)
    mid, half = lim.mean(1), (lim[:, 1] - lim[:, 0]) / 2
    steps = torch.randn(rows, D, generator=g) * 0.05
    pos = torch.cumsum(steps, 0)
    # reflect into [-1, 1] (triangle wave) then map to the limits
    pos = torch.abs(((pos + 1) % 4) - 2) - 1
    vel = torch.vstack([torch.zeros(1, D), (pos[1:] - pos[:-1]) / 0.2]).clamp(-1, 1)
    return torch.hstack([mid + half * pos, vel * half]).to(torch.float32).contiguous()


def random_controls(shape, seed=0):
    g = torch.Generator().manual_seed(seed)
    return ((torch.rand(*shape, generator=g) * 2 - 1) * 0.5).to(torch.float32)
