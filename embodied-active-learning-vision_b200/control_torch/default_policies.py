"""Default policies of the planner (reference control_torch/default_policies.py).

``Roll`` (default) and ``Zero`` are control-sequence bookkeeping with dmu/dx = 0, which the fused evals assume.
``BarrierPush`` and ``LQR`` feed the state back: the planner evaluates their closed loop and the adjoint with
dmu/dx on the device (klerg_policy_rollout / klerg_adjoint_policy), the classes carry the parameters.
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


class BarrierPush(torch.nn.Module):
    """Replay the plan (iterations after the first) or start from zeros, and push back with -weight * velocity
    wherever a position sits on its +-1 wall moving outwards (reference default_policies.py:53-97).  The planner
    evaluates the closed loop on the device (klerg_policy_rollout); ``__call__`` / ``dx`` are the same rule on host
    tensors for callers that step the policy themselves."""
    feedback = True

    def __init__(self, model, horizon):
        super().__init__()
        self.num_actions = model.num_actions
        self.num_states = model.num_states
        self.dtype = model.dtype
        self._dx = torch.zeros([model.num_actions, model.num_states], dtype=self.dtype)
        self.u = iter([])
        self.b_lim = [[-1., 1.]] * model.num_states
        self.skip = [s.upper() == s for s in model.states]  # positions only
        self.weight = 5.
        if any(not sk for sk in self.skip[model.num_actions:]):
            # lower-case velocity-magnitude states ('v' of the speed model) index past the controls in the reference
            raise NotImplementedError("BarrierPush: the speed-state model is not supported (IndexError in the reference)")
        self._use_u = False

    def reset(self, x=None, u=None, iter_idx=0):
        self._use_u = iter_idx > 0
        self.u = iter(u) if self._use_u else iter([])
        return u

    def _pushed(self, x):
        """Mask [A]: positions on (or beyond) a wall of their +-1 box with the velocity pointing outwards."""
        a = self.num_actions
        box = torch.as_tensor(self.b_lim[:a], dtype=x.dtype)
        pos, vel = x[:a], x[a:2 * a]
        free = torch.tensor(self.skip[:a])
        return (((pos >= box[:, 1]) & (vel > 0)) | ((pos <= box[:, 0]) & (vel < 0))) & ~free

    def clipped(self, x, u):
        hit = self._pushed(x)
        u[hit] = -self.weight * x[self.num_actions:2 * self.num_actions][hit]  # in place: u is a row of the plan
        return u

    def dx(self, x=None, u=None):
        out = self._dx.clone()
        rows = self._pushed(x).nonzero().flatten()
        out[rows, rows + self.num_actions] = -self.weight
        return out

    def __call__(self, x=None):
        u = next(self.u, None)
        if u is None:
            u = torch.zeros(self.num_actions, dtype=self.dtype)
        return self.clipped(x, u)

    def device_spec(self):
        from . import _cabi as cabi
        return cabi.policy_spec(cabi.POLICY_BARRIER_PUSH, use_u=self._use_u, weight=self.weight)


class LQR(torch.nn.Module):
    """u = -K x with the continuous-time LQR gain of the model linearised at ones (reference
    default_policies.py:100-119; the Riccati solve is scipy's, on the host, once per controller)."""
    feedback = True

    def __init__(self, model, horizon):
        super().__init__()
        import numpy as np
        from scipy.linalg import solve_continuous_are
        A, B = model.get_lin(torch.ones(model.num_states), torch.ones(model.num_actions))
        A = A.cpu().numpy()
        B = B.cpu().numpy()
        if A.shape[0] != 2 * model.num_actions:
            raise NotImplementedError("LQR: Q is sized for a double-integrator state (positions + velocities)")
        Q = np.diag([5.] * model.num_actions + [1.] * model.num_actions)
        Rm = np.eye(model.num_actions) * 100. * horizon
        Pm = solve_continuous_are(A, B, Q, Rm, balanced=False)
        self.Klqr = torch.as_tensor(np.linalg.inv(Rm) @ B.T @ Pm, dtype=model.dtype)

    def reset(self, x=None, u=None, iter_idx=0):
        return u

    def dx(self, x=None, u=None):
        return -self.Klqr.clone()

    def __call__(self, x):
        return -self.Klqr @ x

    def device_spec(self):
        from . import _cabi as cabi
        return cabi.policy_spec(cabi.POLICY_LQR, K=self.Klqr.tolist())
