"""Workspace barrier (reference control_torch/barrier.py) evaluated on the GPU."""
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
