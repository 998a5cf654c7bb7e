"""Default policies of the planner (reference control_torch/default_policies.py).

Only control-sequence bookkeeping lives here (no arithmetic).  ``Roll`` (default)
and ``Zero`` have dmu/dx = 0, which the device adjoint sweep assumes;
``BarrierPush`` and ``LQR`` are not ported (never enabled by the shipped configs).
"""
import torch


class Roll(torch.nn.Module):
    """Replay the current plan; ``reset(x, u, k<0)`` shifts it left by |k| and zero-fills."""

    def __init__(self, model, horizon):
        super().__init__()
        self.num_actions = model.num_actions
        self.dtype = model.dtype
        self._dx = torch.zeros([model.num_actions, model.num_states], dtype=self.dtype)
        self.u = iter([])

    def reset(self, x=None, u=None, iter_idx=0):
        if iter_idx < 0:
            u = torch.roll(u, iter_idx, 0)
            u[iter_idx:] = 0.0
        self.u = iter(u)
        return u

    def dx(self, x=None, u=None):
        return self._dx.clone()

    def __call__(self, x=None):
        try:
            return next(self.u)
        except StopIteration:
            print("out of controls")
            return torch.zeros(self.num_actions, dtype=self.dtype)


class Zero(Roll):
    """Like Roll, but a negative ``iter_idx`` clears the plan instead of shifting it."""

    def reset(self, x=None, u=None, iter_idx=0):
        if iter_idx < 0:
            u = torch.zeros_like(u)
        self.u = iter(u)
        return u

    def __call__(self, x=None):
        try:
            return next(self.u)
        except StopIteration:
            return torch.zeros(self.num_actions, dtype=self.dtype)


def _not_ported(name):
    class _Missing:
        def __init__(self, *a, **k):
            raise NotImplementedError(f"default policy {name!r} is not ported to the B200 controller "
                                      "(its dmu/dx != 0 needs the general adjoint); use Roll or Zero")
    _Missing.__name__ = name
    return _Missing


BarrierPush = _not_ported("BarrierPush")
LQR = _not_ported("LQR")
