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
    "xyXY_speed": dict(states="xyXY", D=4, horizon=10, cap=40, n=500, m=30, steps=8, std=0.1, x0=[0.2, 0.1, 0.0, 0.0], vel=True,
                       magnitude=True),
    "xy_weightenv": dict(states="xy", D=2, horizon=10, cap=40, n=300, m=30, steps=8, std=0.05, x0=[0.5, 0.5, 0, 0], weight_env=True),
    "xy_uniform": dict(states="xy", D=2, horizon=10, cap=40, n=300, m=30, steps=6, std=0.05, x0=[-0.5, 0.5, 0, 0], uniform=True),
    # non-default robot_config.yaml flags, set on the constructed controller
    "xy_saturate_fixedlam": dict(states="xy", D=2, horizon=10, cap=40, n=300, m=30, steps=8, std=0.05, x0=[0.4, -0.3, 0, 0],
                                 flags=dict(saturate=True, fixed_lam=True, lam=2)),
    "xy_noappsearch": dict(states="xy", D=2, horizon=10, cap=40, n=300, m=30, steps=8, std=0.05, x0=[-0.2, 0.6, 0, 0],
                           flags=dict(ctrlAppSearch=False)),
    "xyz_nearloc_recent": dict(states="xyz", D=3, horizon=12, cap=64, n=600, m=40, steps=8, std=0.08, x0=[0.3, 0.2, -0.1, 0, 0, 0],
                               flags=dict(sample_near_current_loc=True, add_recent_history=True)),
    "xyz_prior": dict(states="xyz", D=3, horizon=10, cap=40, n=400, m=30, steps=6, std=0.08, x0=[0.0, 0.1, 0.2, 0, 0, 0],
                      flags=dict(use_prior=True)),
    # state-feedback default policies (default_policies.py:53-119); the BarrierPush case starts on the walls moving outwards
    "xy_lqr": dict(states="xy", D=2, horizon=10, cap=40, n=300, m=30, steps=8, std=0.05, x0=[0.6, -0.5, 0.3, 0.2], policy="LQR"),
    "xyz_barrierpush": dict(states="xyz", D=3, horizon=12, cap=40, n=400, m=30, steps=8, std=0.08,
                            x0=[0.97, -0.98, 0.2, 0.6, -0.5, 0.1], policy="BarrierPush"),
    "xyw_plot_corners": dict(states="xyw", D=3, horizon=10, cap=50, n=400, m=30, steps=6, std=0.08, x0=[0.2, -0.4, 0.1, 0, 0, 0],
                             plot=True, flags=dict(test_corners=True)),
}


def apply_case_flags(r, case):
    """Flags the reference reads from robot_config.yaml, set on a constructed controller (reference, oracle or the
    B200 mirror alike) together with what Robot.__init__ would have created for them (klerg.py:174-182)."""
    if case.get("weight_env"):
        r.weight_env, r.weight_temp = True, False
    flags = case.get("flags", {})
    for k, v in flags.items():
        setattr(r, k, v)
    name = case.get("policy")
    if name:
        if hasattr(r, "set_policy"):  # the oracle
            r.set_policy(name)
        else:  # reference / B200 mirror: the class of that name from the module the controller took its policy from
            import sys
            r.policy = getattr(sys.modules[type(r.policy).__module__], name)(r.planner, r.horizon)
    if flags.get("sample_near_current_loc") and not hasattr(r, "loc_sampler"):
        r.loc_sampler = torch.distributions.Normal(torch.zeros_like(r.std), r.std * 4.)
    if flags.get("use_prior") and not hasattr(r.prior_dist, "device"):
        r.prior_dist.device = "cpu"  # klerg.py:460 reads it, PriorDist never sets it

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
        use_magnitude=bool(case.get("magnitude")),
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


# ---- VAE target density (vae/vae.py:244-275) ---------------------------------------------------------------
# name: s_dim, z_dim, hidden_dim (reference order: encoder side first, the decoder reverses it), logvar columns,
#       z-buffer rows (0 = plain z_samples), dx model, #samples (ragged against the 128-row tiles), weight gain
TARGET_CASES = {
    "default": dict(sd=3, zd=16, hidden=[512, 256], nl=1, zbuf=0, dx=False, n=1000, gain=3.0),
    "pose6_rgbvar": dict(sd=6, zd=8, hidden=[64, 32], nl=3, zbuf=0, dx=False, n=333, gain=4.0),
    "zbuffer_dx": dict(sd=2, zd=6, hidden=[96, 40], nl=1, zbuf=3, dx=True, n=257, gain=6.5),
    "wide_logvar": dict(sd=3, zd=4, hidden=[160, 24], nl=9, zbuf=2, dx=False, n=130, gain=1.2),
}


def decoder_weights(case, seed=11, out_extra=5):
    """Seeded decoder weights [(W, b)] * 3 in torch.nn.Linear layout; `gain` spreads the log-variance over the
    clamp range so that both clamp limits are exercised."""
    g = torch.Generator().manual_seed(seed)
    h2, h1 = case["hidden"]  # decoder = reversed(hidden_dim) (vae.py:79)
    dims = [case["zd"] + case["sd"], h1, h2, case["nl"] + out_extra]
    out = []
    for a, b in zip(dims[:-1], dims[1:]):
        bound = 1.0 / (a ** 0.5)
        w = (torch.rand(b, a, generator=g) * 2 - 1) * bound * case["gain"]
        bias = (torch.rand(b, generator=g) * 2 - 1) * bound
        out.append((w.contiguous(), bias.contiguous()))
    return out


def target_samples(case, seed=3):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(case["n"], case["sd"], generator=g) * 2.3 - 1.15).to(torch.float32)


class _ZRows:
    def __init__(self, rows):
        self.rows = rows

    def get_samples(self):
        return self.rows.detach().clone()


class DecoderModel:
    """Stand-in with the attributes of the reference's VAE that pdf_torch reads (vae/vae.py:11-110), evaluated
    on the CPU by the oracle.  Used where the reference itself cannot travel (GPU box)."""

    def __init__(self, case, z_rows, seed_x=None, seed=11, initialised=True, out_extra=5):
        ws = decoder_weights(case, seed, out_extra)
        layers = []
        for i, (w, b) in enumerate(ws):
            lin = torch.nn.Linear(w.shape[1], w.shape[0])
            with torch.no_grad():
                lin.weight.copy_(w)
                lin.bias.copy_(b)
            layers.append(lin)
            if i + 1 < len(ws):
                layers.append(torch.nn.ReLU())
        self.decode = torch.nn.Sequential(*layers)
        self.weights = ws
        z_rows = torch.as_tensor(z_rows, dtype=torch.float32).reshape(-1, case["zd"])
        self.use_buffer = case["zbuf"] > 0
        self.z_buff = _ZRows(z_rows)
        self.z_samples = z_rows[:1].clone()
        self.ylogvar_dim = torch.tensor(case["nl"])
        self.logvar_lims = (-10, 2)
        self.init = torch.tensor([bool(initialised)])
        self.dx = bool(case["dx"])
        self.seed_x = torch.zeros(1, case["sd"]) if seed_x is None else torch.as_tensor(seed_x, dtype=torch.float32).reshape(1, -1)
        self.device = "cpu"
        self.dtype = torch.float32

    def z_rows(self):
        return self.z_buff.get_samples() if self.use_buffer else self.z_samples

    def pdf_torch(self, samples):
        from oracle import target_oracle
        return target_oracle.vae_pdf(samples.to("cpu"), self.weights, self.z_rows(), int(self.ylogvar_dim),
                                     self.logvar_lims, self.seed_x if self.dx else None, bool(self.init))

    def init_uniform_grid(self, x):
        return x.sum(1) ** 0
