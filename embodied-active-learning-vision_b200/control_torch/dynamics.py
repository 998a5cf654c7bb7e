"""Integrator models of the planner (reference control_torch/dynamics.py) on the GPU.

Same classes, constructor arguments and stateful ``reset/step/get_lin`` API as
the reference; every step is one klerg_rollout launch (H = 1).  The planner does
not go through these objects for its horizon rollouts - it calls the batched
kernel directly - they exist for the callers that drive ``robot.robot`` /
``robot.planner`` by hand and for the single real-robot step per control tick.
"""
import torch

from . import _cabi as cabi
from . import engine

def rk4_integrate(f, dt, xt, ut):
    """Generic RK4 step of x' = f(x, u) for a caller-supplied f (reference dynamics.py:7-13).  The planner's own
    models never go through it (their RK4 step is closed-form inside the CUDA kernels); it is kept for callers
    that integrate their own vector field and runs on whatever device the tensors live on."""
    stages = []
    for c in (0.0, 0.5, 0.5, 1.0):  # classical Butcher tableau
        x_stage = xt if not stages else xt + c * stages[-1]
        stages.append(dt * f(x_stage, ut))
    return xt + (stages[0] + 2.0 * (stages[1] + stages[2]) + stages[3]) / 6.0


_KINDS = {"single": cabi.DYN_SINGLE, "double": cabi.DYN_DOUBLE, "speed": cabi.DYN_SPEED, "roll": cabi.DYN_ROLL}


class BaseIntegratorEnv(torch.nn.Module):
    kind = "double"

    def __init__(self, num_states, num_actions, dt, rk4, states, dtype=torch.float32, rpw=(0, 0, 0), ang_map=None):
        super().__init__()
        if not rk4:
            raise NotImplementedError("only rk4=True is ported (the controller never uses Euler steps)")
        cabi.require_cuda()
        self.num_states = num_states
        self.num_actions = num_actions
        self.dt = torch.tensor(dt, dtype=dtype)
        self.rk4 = rk4
        self.states = states
        self.dtype = dtype
        self.state_dist = torch.distributions.Uniform(0.0, 0.1)
        self.R = torch.eye(3)
        self.spec = cabi.dyn_spec(_KINDS[self.kind], num_states, num_actions, float(dt), rpw, ang_map)

    # -- device plumbing ------------------------------------------------------
    def _launch(self, u, want_lin=False):
        x = self.state.detach().to(device="cuda", dtype=torch.float32)
        uu = torch.as_tensor(u).detach().to(device="cuda", dtype=torch.float32).reshape(1, 1, self.num_actions)
        R0 = self.R.to(device="cuda", dtype=torch.float32).reshape(9).contiguous() if self.kind == "roll" else None
        return engine.rollout(self.spec, None, x.contiguous(), uu, R0=R0, want_lin=want_lin, want_R=self.kind == "roll")

    # -- reference API ---------------------------------------------------------
    def reset(self, state=None):
        if state is None:
            self.state = self.state_dist.sample((self.num_states,))
        elif isinstance(state, torch.Tensor):
            self.state = state[: self.num_states].clone()
        else:
            self.state = torch.tensor(state[: self.num_states], dtype=self.dtype)
        return self.state.clone()

    def step(self, u, save=True):
        out = self._launch(u)
        nxt = out["traj"][0, 1].to(device="cpu", dtype=self.dtype)
        if out["R"] is not None:  # the reference advances R even when save=False (dynamics.py:274-276)
            self.R = out["R"][0].reshape(3, 3).cpu()
        if save:
            self.state = nxt.clone()
        return nxt

    def get_lin(self, x, u):
        return self.fdx(x, u), self.fdu(x, u)

    def fdx(self, x, u):
        a, n = self.num_actions, self.num_states
        A = torch.zeros((n, n), dtype=self.dtype)
        if self.kind == "single":
            return A
        if self.kind == "roll":
            P = self._launch(torch.zeros(a), want_lin=True)["P"][0, 0].reshape(a, a).to("cpu", self.dtype)
            R_keep = self.R  # linearisation must not advance R
            A[:a, a: 2 * a] = P
            self.R = R_keep
        else:
            A[:a, a: 2 * a] = torch.eye(a, dtype=self.dtype) * 0.8
        return A

    def fdu(self, x, u):
        a, n = self.num_actions, self.num_states
        B = torch.zeros((n, a), dtype=self.dtype)
        if self.kind == "single":
            return torch.eye(a, dtype=self.dtype)
        B[a: 2 * a] = torch.eye(a, dtype=self.dtype)
        if self.kind == "speed":
            sg = torch.as_tensor(x)[a: 2 * a].sign().to(self.dtype)
            sg[sg == 0] = 1.0
            B[2 * a:] = torch.diag(sg)
        return B

    def f(self, x, u):
        raise NotImplementedError("continuous-time f(x,u) is folded into the device RK4 step; use step()")


class SingleIntegratorEnv(BaseIntegratorEnv):
    kind = "single"

    def __init__(self, dt=0.1, x0=torch.zeros(2), states=None, rk4=True, dtype=torch.float32):
        dim = len(x0)
        super().__init__(dim, dim, dt, rk4, states, dtype=dtype)
        self.reset(x0)


class DoubleIntegratorEnv(BaseIntegratorEnv):
    kind = "double"

    def __init__(self, dt=0.1, x0=torch.zeros(4), states=None, rk4=True, dtype=torch.float32):
        dim = len(x0)
        super().__init__(dim, int(dim / 2), dt, rk4, states.lower() + states.upper(), dtype=dtype)
        self.reset(x0)


class DoubleIntegratorSpeedEnv(BaseIntegratorEnv):
    kind = "speed"

    def __init__(self, dt=0.1, x0=torch.zeros(6), states=None, rk4=True, dtype=torch.float32):
        dim = len(x0)
        super().__init__(dim, int(dim / 3), dt, rk4, states.lower() + "v" * len(states) + states.upper(), dtype=dtype)
        self.reset(x0)

    def reset(self, state=None):
        x = super().reset(state)
        if len(x) < self.num_states:
            a = self.num_actions
            self.state = torch.hstack([self.state, torch.abs(self.state[a: 2 * a])])
        return self.state.clone()


class DoubleIntegratorRollEnv(BaseIntegratorEnv):
    """6-D pose model: angles are advanced on SO(3) (reference dynamics.py:224-315)."""
    kind = "roll"

    def __init__(self, dt=0.1, x0=torch.zeros(12), states=None, rk4=True, dtype=torch.float32,
                 rot_to_angles_fn=None, angles_to_rot_fn=None):
        dim = len(x0)
        a = int(dim / 2)
        rot = [-1, -1, -1]
        rest = ""
        for idx, key in enumerate(states):
            if key in "rpw":
                rot["rpw".index(key)] = idx
            else:
                rest += key
        assert all(r >= 0 for r in rot), f"need roll, pitch, and yaw to use this dynamics model, got states {states}"
        rest += "rpw"
        ang_map = None
        if rot_to_angles_fn is not None:
            # the reference passes Lambda(ws_conversion, (robot_lim[rpw], tray_lim[rpw])) (klerg.py:147-149)
            rot_lim, ang_lim = rot_to_angles_fn.vars
            ang_map = (torch.as_tensor(rot_lim).tolist(), torch.as_tensor(ang_lim).tolist())
        self.rpw = torch.tensor(rot)
        self.d_rpw = self.rpw + a
        super().__init__(dim, a, dt, rk4, rest.lower() + rest.upper(), dtype=dtype, rpw=rot, ang_map=ang_map)
        self.reset(x0)

    def reset(self, state=None):
        x = super().reset(state)
        # R = euler_XYZ(angles) is rebuilt on the device at the next launch; keep a host copy for callers
        out = engine.rollout(self.spec, None, self.state.to("cuda", torch.float32).contiguous(),
                             torch.zeros((1, 0, self.num_actions), device="cuda"), want_R=True)
        self.R = out["R"][0].reshape(3, 3).cpu()
        return x
