"""CPU oracle for the KL-ergodic hot path.  TEST INFRASTRUCTURE ONLY.

This module is a torch-CPU fp32 restatement of the algorithm implemented by the
reference in ``franka_test/scripts/control_torch`` (klerg.py, klerg_utils.py,
barrier.py, dynamics.py, memory_buffer.py, default_policies.py, rotations.py).
It exists so that the CUDA path can be checked on a GPU box where the reference
itself is not present.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the
product package never does.

Pinning: ``tests/test_oracle_golden.py`` compares every function below against
golden vectors that ``tests/golden/make_golden.py`` produced by importing the
live reference from ``/root/reference`` (the reference ships no tests or
fixtures of its own), and ``tests/test_oracle_live.py``
re-checks against the live reference whenever ``/root/reference`` exists.

Each function cites the reference lines it restates (paths relative to
``franka_test/scripts/control_torch``).  The arithmetic deliberately uses the
same torch operators in the same order as the reference so that results are
bit-identical on the same host; the structure (free functions + one planner
class) is this repository's own.
"""
from __future__ import annotations

import itertools
import math

import numpy as np
import torch

# ----------------------------------------------------------------------------
# a1-a6: pairwise kernel utilities (klerg_utils.py)
# ----------------------------------------------------------------------------


def kernel_matrix(states_explr, samples, scale, nu):
    """psi[i, j] = exp(-0.5 * sum_d (t_jd - s_id)^2 / scale_d) / nu.

    klerg_utils.py:7-10.  ``scale`` divides the squared difference un-squared.
    states_explr: [1, T, D] (or broadcastable), samples: [N, 1, D] -> [N, T].
    """
    inner = torch.square(states_explr - samples) / scale
    return torch.exp(-0.5 * torch.sum(inner, 2)) / nu


PAIR_CHUNK_BYTES = 1 << 28  # the [N, T, D] broadcast of the reference is evaluated in sample chunks of this size


def _sample_chunks(samples, n_states):
    """The reference materialises [N, T, D] (klerg_utils.py:8): 2.4e13 B at BASELINE config 4.  Rows of the result are
    independent, so the oracle walks the sample axis in chunks - bit-identical per sample."""
    per = max(1, PAIR_CHUNK_BYTES // (4 * max(1, n_states) * max(1, samples.shape[1])))
    return [samples] if samples.shape[0] <= per else list(torch.split(samples, per))


def footprint_sum(states, samples, explr, scale, nu):
    """q_i = sum_j psi[i, j] (klerg_utils.py:17-22)."""
    sub = states[:, explr].unsqueeze(0)
    return torch.cat([torch.sum(kernel_matrix(sub, c.unsqueeze(1), torch.abs(scale), nu), 1)
                      for c in _sample_chunks(samples, states.shape[0])])


def spread_max(states, samples, explr, scale, nu):
    """max_j psi[i, j] (klerg_utils.py:24-29)."""
    sub = states[:, explr].unsqueeze(0)
    return torch.cat([torch.amax(kernel_matrix(sub, c.unsqueeze(1), torch.abs(scale), nu), 1)
                      for c in _sample_chunks(samples, states.shape[0])])


def kl_gradient(x, samples, explr, scale, weights, nu):
    """dg/dx for one planner state (klerg_utils.py:12-15, 31-36).

    g[explr] = sum_i w_i * (-(x - s_i)/|scale|) * psi(x, s_i); other slots zero.
    """
    out = torch.zeros((samples.shape[0], x.shape[0]), dtype=x.dtype)
    xe = x[explr]
    lever = -(xe - samples) / torch.abs(scale)
    k = kernel_matrix(xe.unsqueeze(0).unsqueeze(0), samples.unsqueeze(1), torch.abs(scale), nu)
    out[:, explr] = lever * k
    return torch.sum(out * weights.unsqueeze(1), dim=0)


def unit_mass(dist):
    """In place: NaN -> 1e-6 then divide by the sum (cost_norm, klerg_utils.py:38-42)."""
    dist[torch.isnan(dist)] = 1e-6
    dist /= torch.sum(dist)
    return dist


def renormalize(dist, dim=None, floor=1e-6):
    """x/sum -> clamp(floor) -> log -> minus max -> exp (klerg_utils.py:45-58)."""
    if dim is not None:
        dist = dist / torch.sum(dist, dim, keepdims=True)
        dist = torch.clamp(dist, floor, None)
        dist = torch.log(dist)
        # NB: the reference subtracts the (values, indices) namedtuple of torch.max
        # here, which raises for dim != None; we restate the evident intent.
        dist = dist - torch.max(dist, dim, keepdims=True)[0]
        return torch.exp(dist)
    dist = dist / torch.sum(dist)
    dist = torch.clamp(dist, floor, None)
    dist = torch.log(dist)
    dist = dist - torch.max(dist)
    return torch.exp(dist)


def renormalize_closed_form(dist, floor=1e-6):
    """Algebraic equivalent used by the CUDA path: c/max(c), c = max(x/sum, floor)."""
    c = torch.clamp(dist / torch.sum(dist), floor, None)
    return c / torch.max(c)


# ----------------------------------------------------------------------------
# a15: barrier (barrier.py:40-90, setup at :8-37)
# ----------------------------------------------------------------------------


class OracleBarrier:
    """Quartic wall penalty; limits are shrunk by ``buff`` on both sides."""

    def __init__(self, limits, weight=5.0, power=4.0, buff=0.1):
        n = len(limits)
        self.n = n
        self.buff = buff
        self.power = torch.tensor([power] * n if not isinstance(power, list) else power).unsqueeze(1)
        self.weight = torch.tensor([weight] * n if not isinstance(weight, list) else weight).unsqueeze(1)
        self.set_limits(limits)

    def set_limits(self, limits):
        # barrier.py:58-62
        self.lim = limits.clone()
        for i in range(self.n):
            self.lim[i][0] = limits[i][0] + self.buff
            self.lim[i][1] = limits[i][1] - self.buff

    def _mask(self, xc):
        return torch.stack([(xc <= self.lim[:, 0]), (xc >= self.lim[:, 1])]).T.to(int)

    def value(self, x):
        # barrier.py:70-75
        xc = x[: self.n]
        body = self.weight * (xc.unsqueeze(1) - self.lim) ** self.power
        return torch.sum(self._mask(xc) * body)

    def grad(self, x):
        # barrier.py:77-84
        out = torch.zeros_like(x)
        xc = x[: self.n]
        body = self.power * self.weight * (xc.unsqueeze(1) - self.lim) ** (self.power - 1)
        out[: self.n] = torch.sum(self._mask(xc) * body, 1)
        return out

    def rows(self, xs):
        # barrier.py:86-87
        return torch.stack([self.value(x) for x in xs])


class OracleNoBarrier:
    """barrier.py:147-159."""

    def value(self, x):
        return 0.0

    def grad(self, x):
        return torch.zeros(len(x))

    def rows(self, xs):
        return torch.zeros(len(xs))


def make_barrier(states, robot_lim, robot_ctrl_lim, non_vel_locs, dtype, cfg):
    """barrier.py:8-37 with the yaml flags passed in ``cfg``."""
    lim = torch.tensor(robot_lim[non_vel_locs].tolist() + np.asarray(robot_ctrl_lim).tolist(), dtype=dtype)
    if not cfg["use_barrier"]:
        return OracleNoBarrier(), lim
    n = len(states)
    if cfg["position_barrier"] and not cfg["velocity_barrier"]:
        w = [cfg["barr_weight"]] * n + [0] * n
    elif cfg["velocity_barrier"] and not cfg["position_barrier"]:
        w = [0] * n + [cfg["barr_weight"]] * n
    else:
        w = cfg["barr_weight"]
    return OracleBarrier(lim, weight=w, power=[4.0] * (2 * n), buff=0.1), lim


# ----------------------------------------------------------------------------
# a16-a18: dynamics (dynamics.py) and the rotation helpers (rotations.py)
# ----------------------------------------------------------------------------


def rk4(f, dt, x, u):
    """dynamics.py:7-13."""
    k1 = dt * f(x, u)
    k2 = dt * f(x + k1 / 2.0, u)
    k3 = dt * f(x + k2 / 2.0, u)
    k4 = dt * f(x + k3, u)
    return x + (1 / 6.0) * (k1 + 2.0 * k2 + 2.0 * k3 + k4)


def _axis_rot(axis, angle):
    # rotations.py:37-67
    c = torch.cos(angle)
    s = torch.sin(angle)
    one = torch.ones_like(angle)
    zero = torch.zeros_like(angle)
    flat = {
        "X": (one, zero, zero, zero, c, -s, zero, s, c),
        "Y": (c, zero, s, zero, one, zero, -s, zero, c),
        "Z": (c, -s, zero, s, c, zero, zero, zero, one),
    }[axis]
    return torch.stack(flat, -1).reshape(angle.shape + (3, 3))


def euler_xyz_to_matrix(rot):
    """rotations.py:70-96 with convention 'XYZ' (order flipped: Rz @ Ry @ Rx)."""
    m = [_axis_rot(c, e) for c, e in zip("XYZ", torch.unbind(rot, -1))]
    return torch.matmul(torch.matmul(m[2], m[1]), m[0])


def matrix_to_euler_xyz(mat):
    """rotations.py:142-181 with convention 'XYZ' (flipped to match scipy)."""
    roll = torch.atan2(mat[..., 2, 1], mat[..., 2, 2])
    pitch = torch.asin(mat[..., 2, 0] * -1.0)
    yaw = torch.atan2(mat[..., 1, 0], mat[..., 0, 0])
    return torch.stack((roll, pitch, yaw), -1)


def skew(w):
    # dynamics.py:172-187
    m = torch.zeros(3, 3, dtype=w.dtype)
    m[0, 1] = -w[2]
    m[0, 2] = w[1]
    m[1, 0] = w[2]
    m[1, 2] = -w[0]
    m[2, 0] = -w[1]
    m[2, 1] = w[0]
    return m


def euler_rate_block(rot, R):
    """dynamics.py:189-211 (mutates ``rot[1]`` by +1e-5 like the reference)."""
    rot[1] += 1e-5
    blk = torch.eye(3, dtype=rot.dtype)
    s0, c0 = torch.sin(rot[0]), torch.cos(rot[0])
    t1, c1 = torch.tan(rot[1]), torch.cos(rot[1])
    blk[0, 1] = s0 * t1
    blk[0, 2] = c0 * t1
    blk[1, 1] = c0
    blk[1, 2] = -s0
    blk[2, 1] = s0 / c1
    blk[2, 2] = c0 / c1
    return (blk @ R).ravel()


def wrap_angles(rot):
    # dynamics.py:219-222
    rot[0] = rot[0] % (2 * torch.pi)
    rot[1:] = (rot[1:] + torch.pi) % (2 * torch.pi) - torch.pi
    return rot


def advance_rotation(R, omega, dt):
    # dynamics.py:213-217
    Rn = torch.matrix_exp(skew(omega) * dt) @ R
    return Rn, wrap_angles(matrix_to_euler_xyz(Rn))


class OracleDynamics:
    """kind in {'single', 'double', 'speed', 'roll'} (dynamics.py:16-315).

    ``ang_in``/``ang_out`` are optional [3,2] limit tensors for the affine
    rot<->angle rescale the reference wraps in ``Lambda(ws_conversion, ...)``
    (klerg.py:147-149); ``None`` means identity.
    """

    def __init__(self, kind, dt, x0, states, dtype=torch.float32, ang_in=None, ang_out=None):
        self.kind = kind
        self.dtype = dtype
        dim = len(x0)
        self.num_states = dim
        self.dt = torch.tensor(dt, dtype=dtype)
        if kind == "single":
            self.num_actions = dim
            self.A = torch.zeros([dim, dim])
            self.B = torch.eye(dim)
            self.states = states
        elif kind in ("double", "roll"):
            a = int(dim / 2)
            self.num_actions = a
            self.A = torch.vstack([torch.hstack([torch.zeros((a, a)), torch.eye(a) * 0.8]), torch.zeros((a, dim))]).to(dtype)
            self.B = torch.vstack([torch.zeros((a, a)), torch.eye(a)]).to(dtype)
            if kind == "double":
                self.states = states.lower() + states.upper()
            else:
                rot = torch.ones(3, dtype=int) * -1
                rest = ""
                for i, key in enumerate(states):
                    if key in "rpw":
                        rot["rpw".index(key)] = i
                    else:
                        rest += key
                assert not any(rot < 0)
                self.rpw = rot
                self.d_rpw = rot + a
                rest += "rpw"
                self.states = rest.lower() + rest.upper()
                self.R = torch.eye(3)
                self.rot_idxs = torch.tensor([[r, c] for r in self.rpw for c in self.d_rpw])
                self.ang_in, self.ang_out = ang_in, ang_out
        elif kind == "speed":
            a = int(dim / 3)
            self.num_actions = a
            self.A = torch.vstack([torch.hstack([torch.zeros((a, a)), torch.eye(a) * 0.8, torch.eye(a) * 0.0]), torch.zeros((a * 2, dim))])
            self.B = torch.vstack([torch.zeros((a, a)), torch.eye(a), torch.eye(a)])
            self.states = states.lower() + "v" * len(states) + states.upper()
        else:
            raise ValueError(kind)
        self.reset(x0)

    # -- affine angle maps (franka_utils.ws_conversion via Lambda) --
    def _to_angles(self, v):
        if self.ang_in is None:
            return v
        i, o = self.ang_in, self.ang_out
        return (v - i[:, 0]) / (i[:, 1] - i[:, 0]) * (o[:, 1] - o[:, 0]) + o[:, 0]

    def _from_angles(self, v):
        if self.ang_in is None:
            return v
        i, o = self.ang_out, self.ang_in
        return (v - i[:, 0]) / (i[:, 1] - i[:, 0]) * (o[:, 1] - o[:, 0]) + o[:, 0]

    def fdx(self, x, u):
        if self.kind != "roll":
            return self.A.clone()
        # dynamics.py:283-289: uses self.state, not x
        A = self.A.clone()
        rot = self._to_angles(self.state[self.rpw])
        A[self.rot_idxs[:, 0], self.rot_idxs[:, 1]] = euler_rate_block(rot, self.R)
        return A

    def fdu(self, x, u):
        if self.kind != "speed":
            return self.B.clone()
        # dynamics.py:111-117
        a = self.num_actions
        mod = torch.ones_like(x)
        sg = x[a : 2 * a].sign()
        sg[sg == 0] = 1.0
        mod[2 * a :] = sg
        return mod.unsqueeze(1) * self.B.clone()

    def get_lin(self, x, u):
        return self.fdx(x, u), self.fdu(x, u)

    def f(self, x, u):
        return self.fdx(x, u) @ x + self.fdu(x, u) @ u

    def reset(self, state):
        n = self.num_states
        if isinstance(state, torch.Tensor):
            self.state = state[:n] if self.kind == "roll" else state[:n].clone()
        else:
            self.state = torch.tensor(state[:n], dtype=self.dtype)
        if self.kind == "speed" and len(self.state) < n:
            a = self.num_actions
            self.state = torch.hstack([self.state, torch.abs(self.state[a : 2 * a])])
        if self.kind == "roll":
            self.R = euler_xyz_to_matrix(self._to_angles(self.state[self.rpw]))
        return self.state.clone()

    def step(self, u, save=True):
        nxt = rk4(self.f, self.dt, self.state, u)
        if self.kind == "speed":
            a = self.num_actions
            nxt[-a:] = nxt[a : 2 * a].abs()
        elif self.kind == "roll":
            # dynamics.py:291-301: R always advances, angles are overwritten
            self.R, new_rot = advance_rotation(self.R, self.state[self.d_rpw], self.dt)
            nxt[self.rpw] = self._from_angles(new_rot)
        if save:
            self.state = nxt.clone()
        return nxt


# ----------------------------------------------------------------------------
# a19: memory buffer (memory_buffer.py:38-92)
# ----------------------------------------------------------------------------


class OracleBuffer:
    def __init__(self, capacity, state_dim, dtype=torch.float32):
        self.capacity = capacity
        self.position = 0
        self.full = False
        self.data = torch.empty((capacity, state_dim), dtype=dtype)

    def push(self, state):
        if (self.position + 1) == self.capacity:
            self.full = True
        self.data[self.position] = state
        self.position = (self.position + 1) % self.capacity

    def __len__(self):
        return self.capacity if self.full else self.position

    def draw_indices(self, count):
        """Indices in the reference's RNG order: randperm(len)[:count]."""
        if len(self) == 0:
            return []
        n = self.capacity if self.full else self.position
        count = min(count, n)
        return torch.randperm(n)[:count]

    def sample(self, count):
        return self.data[self.draw_indices(count), :].clone()

    def get_recent(self, count):
        if self.position > count:
            return self.data[self.position - count : self.position].clone()
        if self.full:
            return torch.vstack([self.data[: self.position], self.data[self.position - count :]])
        return self.data[: self.position].clone()

    def get_all(self):
        return self.data.clone() if self.full else self.data[: self.position].clone()

    def reset(self):
        self.position = 0
        self.full = False


# ----------------------------------------------------------------------------
# a20: default policy 'Roll' (default_policies.py:5-28)
# ----------------------------------------------------------------------------


def roll_controls(u, shift):
    """``Roll.reset`` for shift < 0: roll left and zero the tail; else unchanged."""
    if shift < 0:
        u = torch.roll(u, shift, 0)
        u[shift:] = 0.0
    return u


def tilt_barrier(x, lo, hi, weight, power, r_idx, p_idx, w_idx, w_lim, tilt_lim, tilt_power=4, tilt_weight=10.0,
                 ang_map=None):
    """TiltBarrierFunction.barr / dbarr for one state (barrier.py:95-144) -> (value, grad, tilt).  lo / hi / weight /
    power: the wrapped BarrierFunction's rows (limits already shrunk); w_lim: its ORIGINAL yaw limits [2];
    ang_map = (in_lim, out_lim) [3,2] of the rot -> angle map or None."""
    r, p = x[r_idx], x[p_idx]
    if ang_map is not None:
        i, o = ang_map
        r = (r - i[0, 0]) / (i[0, 1] - i[0, 0]) * (o[0, 1] - o[0, 0]) + o[0, 0]
        p = (p - i[1, 0]) / (i[1, 1] - i[1, 0]) * (o[1, 1] - o[1, 0]) + o[1, 0]
    tilt = torch.arccos(torch.cos(r) * torch.cos(p))
    lo, hi = lo.clone(), hi.clone()
    lo[w_idx], hi[w_idx] = tilt / torch.pi * w_lim[0], tilt / torch.pi * w_lim[1]
    n = len(lo)
    below, above = (x[:n] <= lo).to(x.dtype), (x[:n] >= hi).to(x.dtype)
    value = torch.sum(weight * (below * (x[:n] - lo) ** power + above * (x[:n] - hi) ** power))
    grad = torch.zeros_like(x)
    grad[:n] = power * weight * (below * (x[:n] - lo) ** (power - 1) + above * (x[:n] - hi) ** (power - 1))
    if tilt <= tilt_lim:
        value = value + tilt_weight * (tilt - tilt_lim) ** tilt_power
        common = tilt_power * tilt_weight * (tilt - tilt_lim) ** (tilt_power - 1) / torch.sqrt(1 - torch.cos(p) ** 2 * torch.cos(r) ** 2)
        grad[r_idx] += common * torch.sin(r) * torch.cos(p)
        grad[p_idx] += common * torch.sin(p) * torch.cos(r)
    return value, grad, tilt


def velocity_barrier(x_new, x_old, band, weight, power, skip):
    """VelocityBarrier.barr / dbarr for one pair of states (barrier.py:162-205): band [S,2] around x_old on the rows
    that are not skipped -> (value, grad wrt x_new)."""
    lim = x_old.unsqueeze(1) + band
    use = torch.tensor([0.0 if s else 1.0 for s in skip], dtype=x_new.dtype)
    above, below = (x_new >= lim[:, 1]).to(x_new.dtype), (x_new <= lim[:, 0]).to(x_new.dtype)
    value = torch.sum(use * weight * (above * (x_new - lim[:, 1]) ** power + below * (x_new - lim[:, 0]) ** power))
    grad = use * power * weight * (above * (x_new - lim[:, 1]) ** (power - 1) + below * (x_new - lim[:, 0]) ** (power - 1))
    return value, grad


class OracleFeedback:
    """The state-feedback default policies (default_policies.py:53-119) as one rule ``act(x, planned) -> (u, dmudx)``.

    'BarrierPush': the planned control (zeros on the first inner iteration), with -5 v_i on every position that sits on
    its +-1 wall moving outwards; 'LQR': u = -K x, K = R^-1 B^T P of the continuous-time Riccati solution for the
    model linearised at ones (Q = diag(5 on positions, 1 on velocities), R = 100 * horizon * I)."""

    def __init__(self, name, model, horizon):
        self.name = name
        self.a, self.n, self.dtype = model.num_actions, model.num_states, model.dtype
        if name == "LQR":
            from scipy.linalg import solve_continuous_are
            A, B = (m.numpy() for m in model.get_lin(torch.ones(self.n), torch.ones(self.a)))
            Rm = np.eye(self.a) * 100.0 * horizon
            Pm = solve_continuous_are(A, B, np.diag([5.0] * self.a + [1.0] * self.a), Rm, balanced=False)
            self.K = torch.as_tensor(np.linalg.inv(Rm) @ B.T @ Pm, dtype=self.dtype)
        elif name == "BarrierPush":
            self.positions = [i for i, s in enumerate(model.states) if s.upper() != s]
        else:
            raise ValueError(name)

    def uses_plan(self, idx):
        return self.name == "BarrierPush" and idx > 0

    def act(self, x, planned):
        dmudx = torch.zeros([self.a, self.n], dtype=self.dtype)
        if self.name == "LQR":
            return -self.K @ x, -self.K.clone()
        u = planned.clone()
        for i in self.positions:
            v = x[i + self.a]
            if (x[i] >= 1.0 and v > 0) or (x[i] <= -1.0 and v < 0):
                u[i] = -5.0 * v
                dmudx[i, i + self.a] = -5.0
        return u, dmudx


# ----------------------------------------------------------------------------
# line-search window logic (klerg.py:712-751), separated from the cost calls
# ----------------------------------------------------------------------------


def line_search_windows(t_app, idx, horizon, max_app_dur=5):
    """All (tau_i, tau_f) windows the reference would try, in order."""
    if t_app == 0 or t_app == horizon - 1:
        lam = min(horizon, max_app_dur)
    elif t_app == idx:
        lam = min(horizon - t_app, max_app_dur)
    else:
        lam = min(t_app - idx, horizon - t_app - idx, int(math.ceil(max_app_dur / 2)))
    lam = max(lam, 1)
    lam0 = lam
    out = []
    while lam > 0:
        if t_app == idx:
            out.append((t_app, lam + 1))
        elif t_app == horizon - 1:
            out.append((lam - 1, t_app))
        else:
            out.append((t_app - lam, t_app + lam + 1))
        lam -= 1
    return lam0, out


# ----------------------------------------------------------------------------
# a7-a14: the planner (klerg.py Robot)
# ----------------------------------------------------------------------------

DEFAULT_CFG = dict(
    DefaultPolicy="Roll", test_corners=False, use_barrier=True, force_thresh=5, barr_weight=5.0,
    position_barrier=True, velocity_barrier=True, add_recent_history=False, optimize_samples=False,
    sample_near_current_loc=False, weight_env=False, weight_temp=True, ctrlAppSearch=True,
    full_cost=False, fixed_lam=False, lam=1, saturate=False, pct_of_horizon_for_inner_loop=0.5, alpha=1.0,
)  # robot_config.yaml


class OraclePrior:
    """PriorDist (klerg.py:27-50): two fixed Gaussians over the named states, diagonal covariance, + 1e-5."""

    def __init__(self, states):
        base = "xyzrpw"
        duck = [-0.8, -0.8, -0.15, 3.6, 0.5, 0.0]
        ball = [0.6, 0.9, -0.15, 2.6, -0.5, 0.0]
        cov = [0.2, 0.2, 0.5, 0.2, 0.2, 0.5]
        pick = lambda tab, dflt: [tab[base.rfind(s)] if s in base else dflt for s in states]
        covar = torch.diag(torch.FloatTensor(pick(cov, 1.0)))
        self.tdists = [torch.distributions.MultivariateNormal(torch.FloatTensor(pick(m, 0.0)), covar) for m in (duck, ball)]
        self.device = "cpu"  # the reference reads prior_dist.device (klerg.py:460) but never sets it: callers must

    def pdf_torch(self, samples):
        return torch.sum(torch.stack([d.log_prob(samples).exp() for d in self.tdists]), 0) + 1e-5

    def pdf(self, x):
        return self.pdf_torch(torch.as_tensor(x)).numpy()


class OracleRobot:
    """Restatement of ``Robot`` (klerg.py:85-751) for the default 'Roll' policy.

    Constructor arguments follow the reference.  ``cfg`` overrides yaml flags.
    ``trace`` (optional list) collects per-eval records for parity tests.
    """

    def __init__(self, x0, robot_lim, explr_idx, explr_robot_lim_scale=1.0, target_dist=None, dt=0.1,
                 R=0.01, use_vel=True, pybullet=False, horizon=10, buffer_capacity=100, std=0.05,
                 std_plot=0.05, plot_data=False, plot_extra=False, states="xy", plot_states="xy",
                 tray_lim=None, robot_ctrl_lim=None, uniform_tdist=False, vel_states=False,
                 use_magnitude=False, cfg=None):
        self.cfg = dict(DEFAULT_CFG)
        if uniform_tdist:
            # robot_config_uniform.yaml differs only in these two flags
            self.cfg.update(weight_env=True, weight_temp=False)
        if cfg:
            self.cfg.update(cfg)
        for k, v in self.cfg.items():
            setattr(self, k, v)
        self.target_dist = target_dist
        self.dtype = getattr(target_dist, "dtype", torch.float32)
        torch.set_default_dtype(self.dtype)
        self.use_prior = False
        self.prior_dist = OraclePrior(states)
        self.pybullet = pybullet
        self.robot_lim = torch.tensor(robot_lim, dtype=self.dtype)
        self.explr_idx = torch.tensor(explr_idx)
        self.states = states
        self.robot_ctrl_lim = robot_ctrl_lim
        self.uniform_tdist = uniform_tdist
        self.vel_states = vel_states
        self.horizon = horizon
        self.plot_extra = plot_extra
        self.plot_smooth = True
        self.use_magnitude = use_magnitude
        self.use_vel = use_vel
        self.tray_lim = torch.tensor(tray_lim, dtype=self.dtype) if tray_lim is not None else None
        if len(states) == len(plot_states):
            self.plot_extra = False
            self.plot_smooth = False

        # klerg.py:134-157
        if vel_states:
            locs = [i for i, s in enumerate(states) if s == s.lower()]
            self.non_vel_locs = np.array(locs)
            self.vel_locs = [i for i, s in enumerate(states) if s == s.upper()]
            dyn_states = "".join(states[i] for i in locs)
            x0 = np.hstack([np.array(x0)[self.non_vel_locs], np.zeros(len(locs))])
        else:
            self.non_vel_locs = list(range(len(states)))
            self.use_magnitude = False
            dyn_states = states
        extra = {}
        if sum(r in states for r in "rpw") > 1:
            self.rot_states = True
            rpw = [i for i, k in enumerate(states) if k in "rpw"]
            if not torch.all(self.robot_lim[rpw] == self.tray_lim[rpw]):
                extra = dict(ang_in=self.robot_lim[rpw], ang_out=self.tray_lim[rpw])
            kind = "roll"
        else:
            self.rot_states = False
            if self.use_magnitude:
                kind = "speed"
                x0 = np.hstack([x0, np.zeros(len(self.non_vel_locs))])
            else:
                kind = "double"
        scale = 1.0 if use_vel else 3.0
        self.robot = OracleDynamics(kind, dt * scale, x0, dyn_states, self.dtype, **extra)
        self.explr_locs = torch.tensor([i for i, s in enumerate(self.robot.states) if s in states])
        self.planner = OracleDynamics(kind, dt, x0, dyn_states, self.dtype, **extra)

        # klerg.py:168-173
        self.lims = self.robot_lim.clone()
        self.lims += torch.tile(torch.tensor([[-1.0, 1.0]]), (len(self.lims), 1)) * (self.lims[:, [1]] - self.lims[:, [0]]) * (explr_robot_lim_scale - 1.0) / 2.0
        if self.use_magnitude:
            self.lims[self.vel_locs, 0] = 0.0
        self._set_sampler()

        # klerg.py:184-197
        self.num_iters_per_step = max(1, int(self.pct_of_horizon_for_inner_loop * horizon))
        self.std = torch.tensor([1.0 if s.lower() == s else 5.0 for s in states], dtype=self.dtype) * std
        self.std_plot = torch.tensor([1.0 if s.lower() == s else 5.0 for s in states], dtype=self.dtype) * std_plot
        if isinstance(R, (int, float)):
            R = [R] * self.robot.num_actions
        self.R_inv = torch.inverse(torch.diag(torch.tensor(R, dtype=self.dtype)))
        self.u = torch.zeros((horizon, self.planner.num_actions), dtype=self.dtype)
        self.memory_buffer = OracleBuffer(buffer_capacity, self.planner.num_states, dtype=self.dtype)
        self.control_lim = torch.tensor([[-0.5, 0.5] if s in "z" else [-1.0, 1.0] for s in dyn_states], dtype=self.dtype)
        self.plot_data = plot_data
        self.plot_states = plot_states
        self.barrier, self.barr_lim = make_barrier(dyn_states, self.robot_lim, self.robot_ctrl_lim, self.non_vel_locs, self.dtype, self.cfg)
        self.trace = None
        self.n_cost_evals = 0
        self.n_grad_evals = 0
        self.feedback = None
        if self.DefaultPolicy in ("BarrierPush", "LQR"):
            self.set_policy(self.DefaultPolicy)

    def set_policy(self, name):
        """DefaultPolicy of robot_config.yaml (klerg.py:200-202): 'Roll' / 'Zero' replay the plan, the others feed back."""
        self.DefaultPolicy = name
        self.feedback = OracleFeedback(name, self.planner, self.horizon) if name in ("BarrierPush", "LQR") else None

    def _set_sampler(self):
        lo, hi = self.lims[self.explr_idx].T
        self.sample_lo, self.sample_hi = lo, hi
        self.env_sampler = torch.distributions.Uniform(lo, hi)

    # -- klerg.py:223-277 --
    def corners_for(self, plot_idx):
        c = torch.tensor(list(itertools.product(*self.lims[plot_idx])), dtype=self.dtype)
        if len(self.explr_idx) > 2:
            tmp = torch.zeros((c.shape[0], len(self.explr_idx)), dtype=self.dtype)
            tmp[:, plot_idx] = c
            c = tmp
        return c

    def setup_plotting(self, num_samples=100):
        if self.plot_data:
            state = self.robot.state.clone()
            num_samples += 4
            samples = self.env_sampler.sample((num_samples,))
            dummy = renormalize(torch.ones(num_samples))
            locs = torch.tile(state[self.explr_locs].unsqueeze(0), (self.horizon + 1, 1))
            self.plot_data = [samples] + [dummy] * 2 + [locs] + [dummy] * 2 + [torch.tensor([1000.0])]
            self.all_plot_states = [a + b for a, b in itertools.combinations(self.states, 2)]
            self.all_plot_idx = [torch.tensor([self.states.rfind(s) for s in ps]) for ps in self.all_plot_states]
            self.all_corner_samples = [self.corners_for(pi) for pi in self.all_plot_idx]
            self.desired_plot_idx = np.argwhere(np.array(self.all_plot_states) == self.plot_states).item()
            self.plot_idx = self.all_plot_idx[self.desired_plot_idx]
            self.corner_samples = self.all_corner_samples[self.desired_plot_idx]
            self.corners = torch.ones(len(self.corner_samples), dtype=self.dtype)
        else:
            self.plot_idx = torch.tensor([self.states.rfind(s) for s in self.plot_states])
            self.plot_data = None
            self.test_corners = False
        self.last_plan = torch.vstack([self.robot.state] + [self.robot.step(ut) for ut in self.u])

    def update_lims(self, idx, lims):
        if not isinstance(lims, torch.Tensor):
            lims = torch.tensor(lims, dtype=self.dtype)
        self.lims[idx] = lims
        if self.use_magnitude:
            self.lims[self.vel_locs, 0] = 0.0
        self._set_sampler()
        self.corner_samples = self.corners_for(self.plot_idx)
        if self.use_barrier:
            lim = torch.tensor(self.lims[self.non_vel_locs].tolist() + np.asarray(self.robot_ctrl_lim).tolist(), dtype=self.dtype)
            self.barrier.set_limits(lim)

    # -- klerg.py:326-340: consumes RNG exactly like the reference --
    def test(self, num_target_samples=100, N=10):
        N = torch.as_tensor(N)
        torch.randn(N + self.horizon, self.robot.num_states)
        self.env_sampler.sample((num_target_samples,))
        self.env_sampler.sample((num_target_samples,))
        self.setup_plotting(num_target_samples)

    # -- klerg.py:279-291 --
    def step(self, num_target_samples=50, num_traj_samples=30, save_update=False, temp=1.0):
        self.kldiv_planner(num_target_samples, num_traj_samples, temp)
        ctrl = self.u[0].clone()
        if not save_update:
            state = self.robot.step(ctrl, save=False)
        else:
            state = self.robot.step(ctrl)
            self.save_update(state, save=True)
        vel = state[self.planner.num_actions :]
        return state[self.explr_locs].numpy(), vel.numpy(), ctrl.numpy()

    # -- klerg.py:293-323 --
    def save_update(self, full_state, force=0.0, save=True):
        if not isinstance(full_state, torch.Tensor):
            full_state = torch.tensor(full_state, dtype=self.dtype)
        if torch.any(torch.isnan(full_state)):
            return
        if self.pybullet:
            k = torch.norm(self.last_plan[:, self.non_vel_locs] - full_state[self.non_vel_locs], dim=1).argmin().item()
        else:
            k = torch.norm(self.last_plan - full_state, dim=1).argmin().item()
        planned = self.last_plan[k]
        a = self.planner.num_actions
        smooth = 0.5 if self.pybullet else 0.8
        full_state[a:] = smooth * full_state[a:] + (1 - smooth) * planned[a:]
        x = self.robot.reset(full_state)
        if self.feedback is None:  # BarrierPush.reset / LQR.reset hand the plan back unchanged
            self.u = roll_controls(self.u.clone(), -k)
        self.last_policy_idx = k
        if save:
            self.memory_buffer.push(x.clone())

    # -- klerg.py:342-349 --
    def saturate_control(self, u, app_thresh=0.1):
        return torch.tanh(u / app_thresh) * self.control_lim[:, 1]

    # -- klerg.py:367-407 (optimize_samples=False) --
    def get_samples(self, n_target, n_hist):
        if self.add_recent_history:
            recent = self.memory_buffer.get_recent(self.horizon)
            n_target -= len(recent)
        if self.sample_near_current_loc:
            n_target = int(n_target * 0.9)
        samples = self.env_sampler.sample((n_target,))
        if self.sample_near_current_loc:  # klerg.py:181-182: Normal(0, 4 std) around the current exploration state
            near = torch.distributions.Normal(torch.zeros_like(self.std), self.std * 4.0)
            samples = torch.vstack([samples, near.sample((int(n_target / 0.9 * 0.1),)) + self.robot.state[self.explr_locs].clone()])
        if self.add_recent_history:
            samples = torch.vstack([samples, recent[:, self.explr_locs]])
        if self.test_corners:
            samples = torch.vstack([samples, self.corner_samples])
        hist = self.memory_buffer.sample(n_hist)
        return samples, hist, torch.ones(1)

    # -- klerg.py:452-486 --
    def get_target_dist(self, samples, temp, uniform=False, plot=False):
        outside = ((samples < self.robot_lim[:, 0]) | (samples > self.robot_lim[:, 1])).sum(1).gt(0)
        if uniform:
            p = renormalize(self.target_dist.init_uniform_grid(samples.clone()).squeeze())
        elif self.use_prior:  # klerg.py:459-461
            p = renormalize(self.prior_dist.pdf_torch(samples.clone()).squeeze())
        else:
            p = self.target_dist.pdf_torch(samples.clone()).squeeze()
        if self.weight_env or self.weight_temp or plot:
            if len(self.memory_buffer) > 0:
                spread = spread_max(self.memory_buffer.get_all(), samples, self.explr_idx, self.std, nu=1.0)
                spread /= torch.max(spread)
                spread[outside] = 1.0
            else:
                spread = torch.zeros(1)
            if self.weight_env and not plot:
                p += (1 - spread) * p.min()
            elif self.weight_temp or plot:
                p = p ** torch.mean(spread)
            p = renormalize(p)
        return p ** temp

    # -- klerg.py:409-431 --
    def forward(self, idx):
        x = self.planner.reset(self.robot.state.clone())
        u_tmp = roll_controls(self.u.clone(), idx) if self.feedback is None else self.u.clone()
        pending = iter(u_tmp)
        lin, traj = [], []
        for t in range(self.horizon):
            dmudx = torch.zeros([self.planner.num_actions, self.planner.num_states], dtype=self.dtype)
            if self.feedback is None:
                u_tmp[t] = next(pending)
            else:
                planned = u_tmp[t] if self.feedback.uses_plan(idx) else torch.zeros(self.planner.num_actions, dtype=self.dtype)
                u_tmp[t], dmudx = self.feedback.act(x, planned)
            A, B = self.planner.get_lin(x, u_tmp[t])
            lin.append((A, B, self.barrier.grad(x), dmudx))
            traj.append(x)
            x = self.planner.step(u_tmp[t])
        return u_tmp, lin, torch.vstack(traj)

    # -- klerg.py:433-450, 590-593 --
    def backward(self, samples, p, q, nu, lin, traj):
        rho = torch.zeros_like(self.planner.state)
        w = p / q
        du_all = torch.zeros_like(self.u)
        djdlam = torch.zeros(self.horizon, dtype=self.dtype)
        dgdx_all = []

        def rho_dot(r, g):
            dgdx, A, B, dbarr, dmudx = g
            return dgdx - dbarr - (A + B @ dmudx).T @ r

        for t in reversed(range(self.horizon)):
            A, B, dbarr, dmudx = lin[t]
            dgdx = kl_gradient(traj[t], samples, self.explr_locs, self.std, w, nu)
            dgdx_all.append(dgdx)
            rho = rk4(rho_dot, -self.planner.dt, rho, [dgdx, *lin[t]])
            du = -self.R_inv @ B.T @ rho
            du_all[t] = du
            if self.ctrlAppSearch:
                djdlam[t] = rho @ B @ du
        self._last_dgdx = torch.vstack(dgdx_all[::-1])
        return du_all, djdlam

    # -- klerg.py:686-710 --
    def get_cost(self, samples, p, q_base, u_test):
        self.n_cost_evals += 1
        self.planner.reset(self.robot.state.clone())
        traj = torch.vstack([self.planner.step(ut) for ut in u_test])
        q_iter = footprint_sum(traj, samples, self.explr_locs, self.std, torch.ones(1))
        q = renormalize(q_base + q_iter)
        p = unit_mass(p)
        q = unit_mass(q)
        dkl = torch.sum(p * torch.log(p / q)) / torch.ones(1)
        cost = dkl + 0.0 + torch.sum(self.barrier.rows(traj) * 1.0)
        if self.trace is not None:
            self.trace.append(dict(kind="cost", u=u_test.clone(), x0=self.robot.state.clone(), cost=cost.clone(), dkl=dkl.clone(), traj=traj.clone()))
        return cost

    # -- klerg.py:712-751 --
    def line_search(self, t_app, u_app, p, q_base, samples, idx, J0):
        lam, windows = line_search_windows(t_app, idx, self.horizon)
        Jn = J0 * 2
        tau_i, tau_f = idx, lam
        done = False
        k = 0
        while not done and lam > 0:
            tau_last = [tau_i, tau_f]
            Jn_last = Jn
            tau_i, tau_f = windows[k]
            k += 1
            cand = self.u.clone()
            cand[tau_i:tau_f] = u_app
            Jn = self.get_cost(samples.clone(), p.clone(), q_base.clone(), cand)
            lam -= 1
            if (Jn_last < J0) and (Jn > Jn_last):
                done = True
        if (not done) and (Jn < J0):
            tau_last = [tau_i, tau_f]
            done = True
        return tau_last, done

    # -- klerg.py:489-588 (full_cost=False) --
    def kldiv_planner(self, n_target, n_hist, temp=1.0):
        samples, hist, nu = self.get_samples(n_target, n_hist)
        with torch.no_grad():
            p = self.get_target_dist(samples, temp, uniform=self.uniform_tdist)
            q_base = footprint_sum(hist.clone(), samples.clone(), self.explr_locs, self.std.clone(), nu)
            if len(hist) == 0:
                q_base = torch.zeros_like(q_base)
            last_cost = self.get_cost(samples.clone(), p.clone(), q_base.clone(), self.u.clone())
            traj_samples = hist.clone()
            q = renormalize(q_base.clone())
            self._step_inputs = dict(samples=samples.clone(), hist=hist.clone(), p=p.clone(), q_base=q_base.clone())
            for idx in range(self.num_iters_per_step):
                u_tmp, lin, traj = self.forward(idx)
                prev_traj_samples, prev_q = traj_samples.clone(), q.clone()
                traj_samples = torch.vstack([hist, traj]).to(self.dtype)
                q_iter = footprint_sum(traj.clone(), samples.clone(), self.explr_locs, self.std.clone(), nu)
                q = renormalize(q_base + q_iter)
                du, djdlam = self.backward(samples.clone(), p.clone(), q.clone(), nu, lin, traj)
                self.n_grad_evals += 1
                if self.saturate:
                    u_star = self.saturate_control(u_tmp + self.alpha * du)
                else:
                    u_star = torch.clamp(u_tmp + self.alpha * du, *self.control_lim.T)
                t_app = torch.argmin(djdlam).item()
                if self.trace is not None:
                    self.trace.append(dict(kind="grad", idx=idx, u=u_tmp.clone(), x0=self.robot.state.clone(), traj=traj.clone(), q=q.clone(), du=du.clone(), djdlam=djdlam.clone(), dgdx=self._last_dgdx.clone(), t_app=t_app))
                if not self.ctrlAppSearch:  # klerg.py:562-563: take the whole saturated step
                    u_tmp = u_star
                elif djdlam[t_app] < 0:
                    if self.fixed_lam:  # klerg.py:553-554
                        u_tmp[t_app : t_app + self.lam] = u_star[t_app].clone()
                    else:
                        tau, ok = self.line_search(t_app, u_star[t_app].clone(), p.clone(), q_base.clone(), samples.clone(), idx, last_cost)
                        if ok:
                            u_tmp[tau[0] : tau[1]] = u_star[t_app].clone()
                else:
                    q, traj_samples = prev_q, prev_traj_samples
                    break
                cost = self.get_cost(samples.clone(), p.clone(), q_base.clone(), u_tmp.clone())
                if idx > 0 and last_cost <= cost:
                    q, traj_samples = prev_q, prev_traj_samples
                    break
                last_cost = cost.clone()
                self.u = u_tmp
            self.u = torch.nan_to_num(self.u)
            x = self.planner.reset(self.robot.state.clone())
            self.last_plan = torch.vstack([x] + [self.planner.step(ut) for ut in self.u])
            self._step_outputs = dict(q=q, traj_samples=traj_samples, last_cost=last_cost)
            if self.plot_data is not None:
                self.update_plots(traj_samples, samples, p, q, temp, nu)

    # -- klerg.py:602-682 (single=True, use_fut=False) --
    def update_plots(self, traj_samples, samples, p, q, temp, nu):
        base = self.robot.state[self.explr_locs].expand_as(samples).clone()
        if self.plot_extra:
            self.extra_pplot, self.extra_qplot = [], []
            for pi in self.all_plot_idx:
                ps = base.clone()
                ps[:, pi] = samples[:, pi].clone()
                pp = self.get_target_dist(ps, temp, plot=True)
                qp = renormalize(footprint_sum(traj_samples.clone(), ps, self.explr_locs, self.std_plot, nu))
                if self.test_corners:
                    self.extra_pplot.append(pp)
                    self.extra_qplot.append(qp)
                else:
                    self.extra_pplot.append(torch.hstack([pp, self.corners * torch.min(pp)]))
                    self.extra_qplot.append(torch.hstack([qp, self.corners * torch.min(qp)]))
        elif self.plot_smooth:
            ps = base
            ps[:, self.plot_idx] = samples[:, self.plot_idx].clone()
            pp = self.get_target_dist(ps, temp, plot=True)
            qp = renormalize(footprint_sum(traj_samples.clone(), ps, self.explr_locs, self.std_plot, nu))
            if self.test_corners:
                self.extra_pplot, self.extra_qplot = pp, qp
            else:
                self.extra_pplot = torch.hstack([pp, self.corners * torch.min(pp)])
                self.extra_qplot = torch.hstack([qp, self.corners * torch.min(qp)])
        if self.uniform_tdist:
            p = self.get_target_dist(samples, temp, uniform=False, plot=True)
        if self.test_corners:
            self.plot_data[0], self.plot_data[1], self.plot_data[2] = samples.clone(), p.clone(), q.clone()
        else:
            self.plot_data[0] = torch.vstack([samples, self.corner_samples])
            self.plot_data[1] = torch.hstack([p, self.corners * torch.min(p)])
            self.plot_data[2] = torch.hstack([q, self.corners * torch.min(q)])
        self.plot_data[3] = self.last_plan[:, self.explr_locs].clone()
        if self.plot_extra:
            self.plot_data[4] = self.extra_pplot[self.desired_plot_idx].clone()
            self.plot_data[5] = self.extra_qplot[self.desired_plot_idx].clone()
        elif self.plot_smooth:
            self.plot_data[4] = self.extra_pplot.clone()
            self.plot_data[5] = self.extra_qplot.clone()
        else:
            self.plot_data[4] = self.plot_data[1].clone()
            self.plot_data[5] = self.plot_data[2].clone()
        pt, qt = unit_mass(p), unit_mass(q)
        self.plot_data[6] = torch.sum(pt * torch.log(pt / qt))


# ----------------------------------------------------------------------------
# Float64 closed-form evaluation used to bound the
# error of both fp32 implementations in parity tests.
# ----------------------------------------------------------------------------


def eval_closed_form_f64(samples, p, q_base, traj_pre, scale, explr):
    """Return (q, w, dgdx[H,D]) in float64 for a pre-step trajectory."""
    s = samples.double()
    sc = torch.abs(scale).double()
    x = traj_pre[:, explr].double()
    d = x.unsqueeze(0) - s.unsqueeze(1)  # [N,H,D]
    k = torch.exp(-0.5 * (d * d / sc).sum(2))  # [N,H]
    v = q_base.double() + k.sum(1)
    c = torch.clamp(v / v.sum(), 1e-6, None)
    q = c / c.max()
    w = p.double() / q
    g = (-(d / sc) * (k * w.unsqueeze(1)).unsqueeze(2)).sum(0)  # [H,D]
    return q, w, g
