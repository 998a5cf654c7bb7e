"""Workspace barriers (reference control_torch/barrier.py) evaluated on the GPU: BarrierFunction (what setup_barrier
builds), and the two variants the reference keeps beside it, TiltBarrierFunction and VelocityBarrier."""
import os
from argparse import Namespace

import torch
import yaml

from . import _cabi as cabi
from . import engine

base_path = os.path.dirname(os.path.abspath(__file__))


def _read_flags(uniform):
    name = "robot_config_uniform.yaml" if uniform else "robot_config.yaml"
    with open(os.path.join(base_path, name)) as f:
        return yaml.load(f, Loader=yaml.FullLoader)


def setup_barrier(states, robot_lim, robot_ctrl_lim, non_vel_locs, dtype, extra_args, rot_states=False,
                  override_use_barrier=False, uniform=False):
    """Build the barrier from the yaml flags; returns (barrier, barr_lim) like reference barrier.py:8-37."""
    cfg = Namespace(**_read_flags(uniform))
    if override_use_barrier:
        # the reference sets this before the yaml loop, so the yaml value still wins (barrier.py:17-20)
        pass
    barr_lim = torch.tensor(robot_lim[non_vel_locs].tolist() + robot_ctrl_lim.tolist(), dtype=dtype)
    if not cfg.use_barrier:
        return NoBarrier(), barr_lim
    n = len(states)
    if cfg.position_barrier and not cfg.velocity_barrier:
        weight = [cfg.barr_weight] * n + [0] * n
    elif cfg.velocity_barrier and not cfg.position_barrier:
        weight = [0] * n + [cfg.barr_weight] * n
    else:
        weight = cfg.barr_weight
    return BarrierFunction(b_lim=barr_lim, barr_weight=weight, b_buff=0.1, power=[4.0] * (2 * n)), barr_lim


class BarrierFunction(torch.nn.Module):
    """sum_i w_i [(x_i<=lo_i)(x_i-lo_i)^p_i + (x_i>=hi_i)(x_i-hi_i)^p_i], limits shrunk by b_buff."""

    def __init__(self, b_lim, power=4, barr_weight=100.0, b_buff=0.01):
        super().__init__()
        self.ergodic_dim = len(b_lim)
        self.b_buff = b_buff
        if not isinstance(power, list):
            power = [power] * self.ergodic_dim
        self.power = torch.tensor(power, dtype=torch.float32).unsqueeze(1)
        if not isinstance(barr_weight, list):
            barr_weight = [barr_weight] * self.ergodic_dim
        self.barr_weight = torch.tensor(barr_weight, dtype=torch.float32).unsqueeze(1)
        self.update_lims(b_lim)

    def update_lims(self, b_lim):
        lim = torch.as_tensor(b_lim, dtype=torch.float32).clone()
        self.b_lim = lim.clone()
        self.b_lim[:, 0] = lim[:, 0] + self.b_buff
        self.b_lim[:, 1] = lim[:, 1] - self.b_buff

    def update_ergodic_dim(self, new_ergodic_dim):
        self.ergodic_dim = new_ergodic_dim
        self.b_lim = self.b_lim[:new_ergodic_dim]
        self.barr_weight = self.barr_weight[:new_ergodic_dim]
        self.power = self.power[:new_ergodic_dim]

    def spec(self):
        """C-ABI description (read at call time, so in-place edits of the tensors are honoured)."""
        n = self.ergodic_dim
        return cabi.barrier_spec(self.b_lim[:n, 0].tolist(), self.b_lim[:n, 1].tolist(),
                                 self.barr_weight[:n, 0].tolist(), self.power[:n, 0].tolist())

    def _eval(self, x, value, grad):
        cabi.require_cuda()
        rows = x.detach().to(device="cuda", dtype=torch.float32)
        rows = rows.unsqueeze(0) if rows.dim() == 1 else rows
        v, g = engine.barrier_eval(self.spec(), rows.contiguous(), want_value=value, want_grad=grad)
        return v, g

    def barr(self, x):
        v, _ = self._eval(x, True, False)
        return v[0].to(device=x.device, dtype=x.dtype)

    def dbarr(self, x):
        _, g = self._eval(x, False, True)
        return g[0].to(device=x.device, dtype=x.dtype)

    def __call__(self, x):
        v, _ = self._eval(x, True, False)
        return v.to(device=x.device, dtype=x.dtype)

    def __repr__(self):
        return f"BarrierFunction(b_lim={self.b_lim.tolist()}, weight={self.barr_weight.flatten().tolist()})"


def dummy_conversion(x):
    return x


def _affine_limits(fn):
    """(in_lim, out_lim) [n,2] of a ``Lambda(ws_conversion, (in_lim, out_lim))`` angle map, None for the identity."""
    if fn is None or fn is dummy_conversion:
        return None
    vars_ = getattr(fn, "vars", None)
    if vars_ is None or len(vars_) != 2:
        raise NotImplementedError("TiltBarrierFunction: rot_to_angles_fn must be a Lambda(ws_conversion, (in_lim, out_lim)) "
                                  "as Robot builds it (klerg.py:147-149), or None")
    return [torch.as_tensor(v, dtype=torch.float32).reshape(-1, 2) for v in vars_]


class TiltBarrierFunction(torch.nn.Module):
    """Wall barrier + a tilt term (reference barrier.py:95-144; its only use, barrier.py:34-35, is commented out):
    tilt = acos(cos r cos p), the wrapped barrier's yaw limits become tilt/pi * their original values, and
    [tilt <= tilt_lim] * weight * (tilt - tilt_lim)^power is added.  Evaluated on the GPU (klerg_barrier_eval_ext)."""

    def __init__(self, other_bar, states, tilt_lim, tilt_power=4, tilt_weight=10.0, pitch_control=False,
                 rot_to_angles_fn=None, angles_to_rot_fn=None):
        super().__init__()
        self.r_idx = states.rfind('r')
        self.p_idx = states.rfind('p')
        self.b_lim = tilt_lim
        self.power = tilt_power
        self.weight = tilt_weight
        self.other_bar = other_bar
        self.w_idx = states.rfind('w')
        self.w_b_lim = other_bar.b_lim[self.w_idx].clone()
        self.rot_to_angles_fn = rot_to_angles_fn if rot_to_angles_fn is not None else dummy_conversion
        self.rpw = torch.tensor([self.r_idx, self.p_idx, self.w_idx], dtype=int)

    def update_lims(self, b_lim):
        self.other_bar.update_lims(b_lim)

    def update_ergodic_dim(self, new_ergodic_dim):
        self.other_bar.update_ergodic_dim(new_ergodic_dim)

    def _tilt_spec(self):
        t = cabi.TiltSpec()
        t.r_idx, t.p_idx, t.w_idx = int(self.r_idx), int(self.p_idx), int(self.w_idx)
        t.w_lo, t.w_hi = float(self.w_b_lim[0]), float(self.w_b_lim[1])
        t.tilt_lim, t.power, t.weight = float(self.b_lim), float(self.power), float(self.weight)
        lims = _affine_limits(self.rot_to_angles_fn)
        t.has_map = int(lims is not None)
        if lims is not None:
            for k in range(2):  # roll, pitch
                t.rot_lo[k], t.rot_hi[k] = float(lims[0][k][0]), float(lims[0][k][1])
                t.ang_lo[k], t.ang_hi[k] = float(lims[1][k][0]), float(lims[1][k][1])
        return t

    def _eval(self, x, value, grad):
        cabi.require_cuda()
        rows = x.detach().to(device="cuda", dtype=torch.float32)
        rows = rows.unsqueeze(0) if rows.dim() == 1 else rows
        v, g, tilt = engine.barrier_eval_ext(self.other_bar.spec(), rows.contiguous(), tilt=self._tilt_spec(),
                                             want_value=value, want_grad=grad)
        # the reference leaves the yaw limits of the last evaluated row in the wrapped barrier
        self.other_bar.b_lim[self.w_idx] = tilt[-1].cpu() / torch.pi * self.w_b_lim.clone()
        return v, g

    def barr(self, x):
        v, _ = self._eval(x, True, False)
        return v[0].to(device=x.device, dtype=x.dtype)

    def dbarr(self, x):
        _, g = self._eval(x, False, True)
        return g[0].to(device=x.device, dtype=x.dtype)

    def __call__(self, x):
        v, _ = self._eval(x, True, False)
        return v.to(device=x.device, dtype=x.dtype)

    def __repr__(self):
        return f"TiltBarrierFunction(tilt_lim={self.b_lim}, power={self.power}, weight={self.weight}, other={self.other_bar!r})"


class VelocityBarrier(torch.nn.Module):
    """Band of +-b_lim around the OTHER state on the velocity (upper-case) rows (reference barrier.py:162-205, never
    instantiated there).  ``barr(x_new, x_old)`` / ``dbarr`` take single states; ``__call__(x_old, x_new)`` hands its
    rows to ``barr`` in the order given - x_old in the role of x_new - exactly like the reference."""

    def __init__(self, planner_states, b_lim=0.1, power=4, barr_weight=100.0):
        super().__init__()
        self.planner_states = planner_states
        self.skip = [s.lower() == s for s in planner_states]  # only velocities
        self.state_dim = len(planner_states)
        self.power = power if isinstance(power, list) else [power] * self.state_dim
        self.barr_weight = barr_weight if isinstance(barr_weight, list) else [barr_weight] * self.state_dim
        if isinstance(b_lim, float):
            b_lim = torch.tile(torch.tensor([[-1., 1.]]), (self.state_dim, 1)) * b_lim
        self.b_lim = b_lim

    def spec(self):
        lim = torch.as_tensor(self.b_lim, dtype=torch.float32)
        w = [0.0 if sk else float(wt) for sk, wt in zip(self.skip, self.barr_weight)]
        return cabi.barrier_spec(lim[:, 0].tolist(), lim[:, 1].tolist(), w, [float(p) for p in self.power])

    def _eval(self, x_new, x_old, value, grad):
        cabi.require_cuda()
        a = x_new.detach().to(device="cuda", dtype=torch.float32)
        b = x_old.detach().to(device="cuda", dtype=torch.float32)
        a = a.unsqueeze(0) if a.dim() == 1 else a
        b = b.unsqueeze(0) if b.dim() == 1 else b
        v, g, _ = engine.barrier_eval_ext(self.spec(), a.contiguous(), x_ref=b.contiguous(), want_value=value, want_grad=grad)
        return v, g

    def barr(self, x_new, x_old):
        v, _ = self._eval(x_new, x_old, True, False)
        return v[0].to(device=x_new.device, dtype=x_new.dtype)

    def dbarr(self, x_new, x_old):
        _, g = self._eval(x_new, x_old, False, True)
        return g[0].to(device=x_new.device, dtype=x_new.dtype)

    def __call__(self, x_old, x_new):
        v, _ = self._eval(x_old, x_new, True, False)  # sic: barr(xt_old, xt_new), barrier.py:198-199
        return v.to(device=x_old.device, dtype=x_old.dtype)

    def update_ergodic_dim(self, new_ergodic_dim):
        pass

    def __repr__(self):
        return f"VelocityBarrier(states={self.planner_states!r}, b_lim={torch.as_tensor(self.b_lim).tolist()})"


class NoBarrier(torch.nn.Module):
    """Barrier switched off (reference barrier.py:147-159)."""

    def __init__(self):
        pass

    def spec(self):
        return None

    def barr(self, x, others=None):
        return 0.0

    def dbarr(self, x, others=None):
        return torch.zeros(len(x))

    def __call__(self, x, others=None):
        return torch.zeros(len(x))

    def update_ergodic_dim(self, new_ergodic_dim):
        pass

    def update_lims(self, b_lim):
        pass

    def __repr__(self):
        return "NoBarrier (Dummy Function)"
