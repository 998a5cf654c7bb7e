"""KL-ergodic MPC planner with the reference's ``Robot`` API, running on a B200.

Mirrors franka_test/scripts/control_torch/klerg.py:85-751 (constructor arguments,
methods, mutable attributes, RNG call order and return types) so the experiment
loops, the fingerprint scripts and the VAE trainer can keep calling it.  What is
different is where the work happens:

* the host keeps only bookkeeping - configuration, the small control sequence
  ``u`` [H, A], the data-dependent control flow of ``kldiv_planner`` /
  ``line_search`` and the torch CPU RNG draws (bit-exact samples and
  memory-buffer indices);
* every arithmetic step - target-density weighting, history footprint, rollout,
  barrier, footprint, KL cost, importance ratio, gradient, adjoint - is a CUDA
  kernel behind ``libklerg_b200.so``; the <= 5 line-search candidates of an
  iteration are evaluated as ONE batched launch instead of a serial Python loop.

The non-default ``robot_config.yaml`` flags ``saturate``, ``fixed_lam``,
``ctrlAppSearch: False``, ``sample_near_current_loc``, ``add_recent_history``,
``test_corners``, ``PriorDist`` / ``use_prior`` and the BarrierPush / LQR default
policies are built (golden sequences from the live
reference).  Not ported: ``optimize_samples`` (an autograd/Adam pre-conditioner of
the samples through the target model), ``full_cost`` (does not run in the
reference itself), ``average_prior`` (reads an attribute the reference never
defines).
"""
import itertools
import math
import os

import numpy as np
import torch
import yaml

from . import _cabi as cabi
from . import engine
from .barrier import setup_barrier
from .default_policies import BarrierPush, LQR, Roll, Zero  # noqa: F401  (looked up by name from the yaml)
from .dynamics import DoubleIntegratorEnv, DoubleIntegratorRollEnv, DoubleIntegratorSpeedEnv
from .klerg_utils import Lambda
from .memory_buffer import MemoryBuffer_torch
from .planner import PlannerContext
from .target_decoder import DeviceTarget, decoder_supported, is_decoder_model

base_path = os.path.dirname(os.path.abspath(__file__))

try:
    from franka.franka_utils import find_non_vel_locs, ws_conversion
except ImportError:  # package used outside the reference's scripts/ layout
    import sys
    sys.path.insert(0, os.path.dirname(base_path))
    from franka.franka_utils import find_non_vel_locs, ws_conversion


class PriorDist():
    """Fixed two-object prior over the named states (klerg.py:27-50): two Gaussians with a shared diagonal covariance,
    + 1e-5.  Evaluated where the samples live (closed form of MultivariateNormal.log_prob for a diagonal covariance).
    The reference reads ``prior_dist.device`` without ever setting it; here it is set."""

    def __init__(self, states):
        base_states = 'xyzrpw'
        base_duck = [-0.8, -0.8, -0.15, 3.6, 0.5, 0.]
        base_ball = [0.6, 0.9, -0.15, 2.6, -0.5, 0.]
        base_covar = [0.2, 0.2, 0.5, 0.2, 0.2, 0.5]
        pick = lambda tab, dflt: [tab[base_states.rfind(s)] if (s in base_states) else dflt for s in states]
        self.locs = torch.tensor([pick(base_duck, 0.), pick(base_ball, 0.)], dtype=torch.float32)
        self.var = torch.tensor(pick(base_covar, 1.), dtype=torch.float32)
        self.device = 'cuda'

    def pdf(self, x):
        return self.pdf_torch(torch.as_tensor(x)).cpu().numpy()

    def pdf_torch(self, samples):
        locs, var = self.locs.to(samples.device), self.var.to(samples.device)
        half_log_det = 0.5 * torch.log(var).sum() + 0.5 * len(var) * math.log(2 * math.pi)
        d = samples.unsqueeze(0) - locs.unsqueeze(1)
        return torch.exp(-0.5 * (d * d / var).sum(2) - half_log_det).sum(0) + 1e-5


class DeviceSamples:
    """Workspace samples that only exist on the device: this rank's rows + the row count over all ranks."""

    def __init__(self, local, n_total):
        self.local = local
        self.shape = (int(n_total), local.shape[1])


def line_search_windows(t_app, idx, horizon, max_app_dur=5):
    """Windows [tau_i, tau_f) the reference's line_search would try, in order (klerg.py:714-738)."""
    if t_app == 0 or t_app == horizon - 1:
        lam = min(horizon, max_app_dur)
    elif t_app == idx:
        lam = min(horizon - t_app, max_app_dur)
    else:
        lam = min(t_app - idx, horizon - t_app - idx, int(math.ceil(max_app_dur / 2)))
    lam = max(lam, 1)
    lam0, out = lam, []
    while lam > 0:
        if t_app == idx:
            out.append((t_app, lam + 1))
        elif t_app == horizon - 1:
            out.append((lam - 1, t_app))
        else:
            out.append((t_app - lam, t_app + lam + 1))
        lam -= 1
    return lam0, out


def line_search_select(costs, windows, idx, lam0, J0):
    """Replay the reference's sequential accept/stop rule (klerg.py:723-751) on precomputed costs.

    Returns (tau, success, n_evaluated, chosen_cost_index or None).
    """
    Jn = J0 * 2
    tau_i, tau_f = idx, lam0
    done = False
    k = 0
    tau_last = [tau_i, tau_f]
    last_k = None
    while not done and k < len(windows):
        tau_last, Jn_last, last_k_prev = [tau_i, tau_f], Jn, (k - 1 if k > 0 else None)
        tau_i, tau_f = windows[k]
        Jn = costs[k]
        k += 1
        if (Jn_last < J0) and (Jn > Jn_last):
            done = True
            last_k = last_k_prev
    if (not done) and (Jn < J0):
        return [tau_i, tau_f], True, k, k - 1
    if done:
        return tau_last, True, k, last_k
    return tau_last, False, k, None


class Robot(object):
    """ Robot class that runs the KL-Erg MPC Planner (B200 build) """

    def __init__(self, x0, robot_lim, explr_idx, explr_robot_lim_scale=1.0, target_dist=None, dt=0.1,
                 R=0.01, use_vel=True, pybullet=False,
                 horizon=10, buffer_capacity=100, std=0.05, std_plot=0.05, plot_data=False, plot_extra=False,
                 states='xy', plot_states='xy', tray_lim=None, robot_ctrl_lim=None,
                 uniform_tdist=False, vel_states=False, use_magnitude=False, process_group=None):
        cabi.require_cuda()
        cabi.load()
        self.load_yaml(uniform_tdist)
        self._check_flags()

        self.target_dist = target_dist
        for attr, val in zip(['dtype', 'device'], [torch.float32, 'cpu']):
            if not hasattr(self.target_dist, attr):
                setattr(self.target_dist, attr, val)
        self.dtype = self.target_dist.dtype
        if self.dtype != torch.float32:
            raise NotImplementedError("the B200 controller computes in fp32")
        self.device = 'cpu'  # device of the tensors handed back to callers (as in the reference)
        self.cuda = torch.device("cuda", torch.cuda.current_device())
        torch.set_default_dtype(self.dtype)
        self.group = engine.ShardGroup(process_group)

        self.prior_dist = PriorDist(states)
        self.use_prior = False
        self.average_prior = False
        self.pybullet = pybullet

        # workspace
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
        self.tray_lim = torch.tensor(tray_lim, dtype=self.dtype) if tray_lim is not None else tray_lim
        if len(states) == len(plot_states):
            self.plot_extra = False
            self.plot_smooth = False

        if self.vel_states:
            self.non_vel_locs, self.vel_locs, states = find_non_vel_locs(self.states)
            x0 = np.hstack([np.array(x0)[self.non_vel_locs], np.zeros(len(self.non_vel_locs))])
        else:
            self.non_vel_locs = list(range(len(self.states)))
            self.use_magnitude = False

        extra_args = {}
        if sum([rot in self.states for rot in 'rpw']) > 1:
            self.rot_states = True
            rpw = [idx for idx, key in enumerate(self.states) if key in 'rpw']
            if not torch.all(self.robot_lim[rpw] == self.tray_lim[rpw]):
                extra_args['rot_to_angles_fn'] = Lambda(ws_conversion, (self.robot_lim[rpw], self.tray_lim[rpw]))
                extra_args['angles_to_rot_fn'] = Lambda(ws_conversion, (self.tray_lim[rpw], self.robot_lim[rpw]))
            dynamics = DoubleIntegratorRollEnv
        else:
            self.rot_states = False
            if self.use_magnitude:
                dynamics = DoubleIntegratorSpeedEnv
                x0 = np.hstack([x0, np.zeros(len(self.non_vel_locs))])
            else:
                dynamics = DoubleIntegratorEnv

        dt_scale = 1. if self.use_vel else 3.
        x0 = np.asarray(x0, dtype=np.float32)
        self.robot = dynamics(dt=dt * dt_scale, x0=x0, states=states, dtype=self.dtype, **extra_args)
        self.explr_locs = torch.tensor([idx for idx, s in enumerate(self.robot.states) if s in self.states])
        self.planner = dynamics(dt=dt, x0=x0, states=states, dtype=self.dtype, **extra_args)
        if 'b' in self.robot.states:
            self.bv_idx = self.robot.states.rfind('B')

        # sampling box (klerg.py:169-173)
        self.lims = self.robot_lim.clone()
        half = (self.lims[:, [1]] - self.lims[:, [0]]) * (explr_robot_lim_scale - 1.) / 2.
        self.lims += torch.tile(torch.tensor([[-1., 1.]]), (len(self.lims), 1)) * half
        if self.use_magnitude:
            self.lims[self.vel_locs, 0] = 0.
        self.env_sampler = torch.distributions.Uniform(*self.lims[self.explr_idx].T)

        self.num_iters_per_step = max(1, int(self.pct_of_horizon_for_inner_loop * self.horizon))
        vel_scale = [1. if state.lower() == state else 5. for state in self.states]
        self.std = torch.tensor(vel_scale, dtype=self.dtype) * std
        self.std_plot = torch.tensor(vel_scale, dtype=self.dtype) * std_plot
        if isinstance(R, (int, float)):
            R = [R] * self.robot.num_actions
        self.R_inv = torch.inverse(torch.diag(torch.tensor(R, dtype=self.dtype)))
        self.u = torch.zeros((self.horizon, self.planner.num_actions), dtype=self.dtype)
        self.memory_buffer = MemoryBuffer_torch(buffer_capacity, self.planner.num_states, dtype=self.dtype)
        self.control_lim = torch.tensor([[-0.5, 0.5] if state in 'z' else [-1.0, 1.0] for state in states],
                                        dtype=self.dtype)

        policies = {"Roll": Roll, "Zero": Zero, "BarrierPush": BarrierPush, "LQR": LQR}
        if self.DefaultPolicy not in policies:
            raise NotImplementedError(f"DefaultPolicy {self.DefaultPolicy!r} is unknown; use Roll, Zero, BarrierPush or LQR")
        self.policy = policies[self.DefaultPolicy](self.planner, self.horizon)

        self.plot_data = plot_data
        self.plot_states = plot_states
        self.barrier, self.barr_lim = setup_barrier(states, self.robot_lim, self.robot_ctrl_lim, self.non_vel_locs,
                                                    self.dtype, extra_args, self.rot_states, uniform=uniform_tdist)
        self.count = 0
        self.stats = dict(cost_evals=0, grad_evals=0, steps=0)

    def load_yaml(self, uniform):
        file = 'robot_config_uniform.yaml' if uniform else 'robot_config.yaml'
        with open(os.path.join(base_path, file)) as f:
            yaml_config = yaml.load(f, Loader=yaml.FullLoader)
        for k, v in yaml_config.items():
            setattr(self, k, v)

    # ------------------------------------------------------------------ plotting
    def setup_plotting(self, num_samples=100):
        if self.plot_data:
            state = self.robot.state.clone()
            num_samples += 4
            samples = self.env_sampler.sample((num_samples,))
            dummy_qp = torch.ones(num_samples)
            dummy_locs = torch.tile(state[self.explr_locs].unsqueeze(0), (self.horizon + 1, 1))
            dummy_cost = torch.tensor([1000.])
            self.plot_data = [samples] + [dummy_qp] * 2 + [dummy_locs] + [dummy_qp] * 2 + [dummy_cost]
            self.all_plot_states = [x[0] + x[1] for x in itertools.combinations(self.states, 2)]
            self.all_plot_idx = [torch.tensor([self.states.rfind(s) for s in ps]) for ps in self.all_plot_states]
            if any(torch.hstack(self.all_plot_idx) == -1):
                raise ValueError('robot controller (klerg) did not find requested plot state')
            self.all_corner_samples = [self.get_corners(ps) for ps in self.all_plot_idx]
            self.desired_plot_idx = np.argwhere(np.array(self.all_plot_states) == self.plot_states).item()
            self.plot_idx = self.all_plot_idx[self.desired_plot_idx]
            self.corner_samples = self.all_corner_samples[self.desired_plot_idx]
            self.corners = torch.ones(len(self.corner_samples), dtype=self.dtype)
        else:
            self.plot_idx = torch.tensor([self.states.rfind(s) for s in self.plot_states])
            self.plot_data = None
            self.test_corners = False
        self.last_plan = torch.vstack([self.robot.state] + [self.robot.step(ut) for ut in self.u])

    @torch.no_grad()
    def update_lims(self, idx, lims):
        if not isinstance(lims, torch.Tensor):
            lims = torch.tensor(lims, dtype=self.dtype)
        self.lims[idx] = lims
        if self.use_magnitude:
            self.lims[self.vel_locs, 0] = 0.
        self.env_sampler = torch.distributions.Uniform(*self.lims[self.explr_idx].T)
        self.update_corners()
        if self.use_barrier:
            barr_lim = torch.tensor(self.lims[self.non_vel_locs].tolist() + self.robot_ctrl_lim.tolist(), dtype=self.dtype)
            self.barrier.update_lims(barr_lim)

    @torch.no_grad()
    def update_corners(self):
        self.corner_samples = self.get_corners(self.plot_idx)

    @torch.no_grad()
    def get_corners(self, plot_idx):
        corner_samples = torch.tensor(list(itertools.product(*self.lims[plot_idx])), dtype=self.dtype)
        if len(self.explr_idx) > 2:
            tmp = torch.zeros((corner_samples.shape[0], len(self.explr_idx)), dtype=self.dtype)
            tmp[:, plot_idx] = corner_samples
            corner_samples = tmp
        return corner_samples

    def _check_flags(self):
        for flag in ("optimize_samples", "full_cost", "average_prior"):
            if getattr(self, flag, False):
                raise NotImplementedError(f"robot_config flag {flag!r} is not ported to the B200 controller")

    # ------------------------------------------------------------------ public API
    def step(self, num_target_samples=50, num_traj_samples=30, save_update=False, temp=1.0):
        self.kldiv_planner(num_target_samples=num_target_samples, num_traj_samples=num_traj_samples, temp=temp)
        ctrl = self.u[0].clone()
        if not save_update:
            state = self.robot.step(ctrl, save=False)
        else:
            state = self.robot.step(ctrl)
            self.save_update(state, save=True)
        vel = state[self.planner.num_actions:]
        return state[self.explr_locs].numpy(), vel.numpy(), ctrl.numpy()

    @torch.no_grad()
    def save_update(self, full_state, force=0., save=True):
        if not isinstance(full_state, torch.Tensor):
            full_state = torch.tensor(full_state, dtype=self.dtype)
        if torch.any(torch.isnan(full_state)):
            print('got nan in full_state')
            return
        if 'b' in self.states:
            full_state[self.bv_idx] = self.last_plan[1][self.bv_idx]
        # closest planned state (index selection on [H+1,S] host data)
        if self.pybullet:
            dist = torch.norm(self.last_plan[:, self.non_vel_locs] - full_state[self.non_vel_locs], dim=1)
        else:
            dist = torch.norm(self.last_plan - full_state, dim=1)
        policy_idx = dist.argmin().item()
        planned_state = self.last_plan[policy_idx]
        vel_smoothing = 0.5 if self.pybullet else 0.8
        a = self.planner.num_actions
        full_state[a:] = vel_smoothing * full_state[a:] + (1 - vel_smoothing) * planned_state[a:]
        x = self.robot.reset(full_state)
        self.u = self.policy.reset(x, self.u.clone(), -policy_idx)
        if save:
            self.memory_buffer.push(x.clone())

    @torch.no_grad()
    def test(self, num_target_samples=100, N=10):
        """Warm-up.  Consumes the CPU RNG exactly like the reference (klerg.py:327-340)."""
        N = torch.as_tensor(N)
        torch.randn(N + self.horizon, self.robot.num_states)
        samples = self.env_sampler.sample((num_target_samples,))
        ratio = self.env_sampler.sample((num_target_samples,)).sum(1)
        # touch the two hot kernels once so that module load / first-launch cost is paid here
        spec = cabi.kernel_spec(len(self.explr_locs), self.planner.num_states, self.explr_locs.tolist(),
                                self.std.tolist(), 1.0)
        s_dev = samples.to(self.cuda)
        packed = engine.pack_samples(spec, s_dev)
        dummy = torch.zeros((self.horizon, self.planner.num_states), device=self.cuda)
        engine.footprint(spec, 0, dummy, packed, s_dev.shape[0])
        engine.kl_gradient(spec, dummy[:1], packed, s_dev.shape[0], ratio.to(self.cuda))
        self.traj_footprint_vec_jit = None  # the reference creates its jitted kernel here
        self.setup_plotting(num_target_samples)

    # ------------------------------------------------------------------ sampling / target
    def get_samples(self, num_target_samples, num_traj_samples):
        """Host RNG draws in the reference's order: uniform samples, then the buffer permutation."""
        self._check_flags()
        if self.add_recent_history:
            recent = self.memory_buffer.get_recent(self.horizon)
            num_target_samples -= len(recent)
        if self.sample_near_current_loc:
            num_target_samples = int(num_target_samples * 0.9)
        n_near = int(num_target_samples / 0.9 * 0.1) if self.sample_near_current_loc else 0
        fixed = []  # appended rows that need no draw
        if self.add_recent_history:
            fixed.append(recent[:, self.explr_locs])
        if self.test_corners:
            fixed.append(self.corner_samples)
        n_total = num_target_samples + n_near + sum(len(e) for e in fixed)

        def near_current():
            """Normal(0, 4 std) around the current exploration state, drawn AFTER the uniform samples (RNG order of
            klerg.py:375-392; the sampler is what Robot.__init__ creates at klerg.py:181-182)."""
            if not n_near:
                return []
            if not hasattr(self, "loc_sampler"):
                self.loc_sampler = torch.distributions.Normal(torch.zeros_like(self.std), self.std * 4.)
            return [self.loc_sampler.sample((n_near,)) + self.robot.state[self.explr_locs].clone()]

        if self._device_draw_ok() and num_target_samples * len(self.explr_idx) >= self.device_draw_min_numbers:
            # the same draw, continued on the device from the host generator's state (bit-exact samples, the host
            # generator ends where the host draw would have left it - device_uniform returns after handing the
            # state back): no 4*N*D-byte host pass and copy per step
            lo, hi = self.group.shard_bounds(n_total)
            draw = (num_target_samples, self.env_sampler.low, self.env_sampler.high,
                    min(lo, num_target_samples), min(hi, num_target_samples), self.cuda)
            if self.prefetch_draw and self._prefetch is None:
                self._prefetch = engine.UniformPrefetch()
            parts = [engine.device_uniform(*draw, prefetch=self._prefetch if self.prefetch_draw else None)]
            off = num_target_samples
            for e in near_current() + fixed:  # appended rows that fall into this rank's slice
                a, b = max(lo, off) - off, min(hi, off + len(e)) - off
                if b > a:
                    parts.append(e[a:b].to(self.cuda, non_blocking=True))
                off += len(e)
            samples = DeviceSamples(torch.cat(parts).contiguous() if len(parts) > 1 else parts[0], n_total)
        else:
            samples = self.env_sampler.sample((num_target_samples,))
            extras = near_current() + fixed
            if extras:
                samples = torch.vstack([samples] + extras)
        hist_dev, hist_idx = self.memory_buffer.sample_device(num_traj_samples)
        self.last_hist_idx = hist_idx
        # The generator now stands where the next step's draw will find it (unless somebody else draws in between):
        # enqueue that draw speculatively.  Large workspaces at once - one serial CTA, it overlaps this step's history
        # pass; small ones after the plan came back (_after_plan), so that the side kernel never delays this step.
        self._next_draw = None
        if self.prefetch_draw and isinstance(samples, DeviceSamples):
            self._next_draw = draw
            if num_target_samples * len(self.explr_idx) >= self.prefetch_early_numbers:
                self._after_plan()
        return samples, hist_dev, torch.ones(1)

    prefetch_draw = True              # speculative draw of the next step's samples (engine.UniformPrefetch)
    prefetch_early_numbers = 2_000_000  # from this many numbers per draw the speculation starts before the history pass
    _prefetch = None
    _next_draw = None

    def _after_plan(self):
        draw, self._next_draw = self._next_draw, None
        if draw is not None and self._prefetch is not None:
            self._prefetch.launch(*draw)

    device_rng = True  # draw the workspace samples on the device when nothing needs them on the host
    device_draw_min_numbers = 16_384  # below this the host draw + copy is cheaper than the state hand-over (config 1)

    def _device_draw_ok(self):
        """The samples can stay on the device when the target density is evaluated there and no plot data (host
        tensors by contract, klerg.py:659-682) is kept."""
        if not self.device_rng or self.plot_data is not None or self.uniform_tdist:
            return False
        if self.use_prior:
            return True  # PriorDist is evaluated on the device
        target = self._device_target()
        return torch.device(getattr(target, "device", "cpu")).type == "cuda"

    def _pdf(self, samples_host, uniform, samples_dev=None):
        """The target density is an INPUT of the path (VAE / belief grid); evaluated where it lives.

        Returns (values for this rank's slice, pre_renorm).  ``pdf_torch`` is point-wise (vae.py:244-275), so a
        sharded controller evaluates it on its own slice only - reusing the samples already on the device when
        the target lives there; ``init_uniform_grid`` normalises over the whole batch and is evaluated in full."""
        if uniform:
            tdev = torch.device(self.target_dist.device)
            full = self.target_dist.init_uniform_grid(samples_host.clone().to(tdev)).squeeze()
            return self._shard(full), True
        if self.use_prior:  # klerg.py:459-461: the fixed prior, renormalised, instead of the model
            if samples_dev is not None:
                return self.prior_dist.pdf_torch(samples_dev).squeeze(), True
            return self.prior_dist.pdf_torch(self._shard(samples_host).to(self.cuda)).squeeze(), True
        target = self._device_target()
        tdev = torch.device(target.device)
        if samples_dev is not None and tdev.type == samples_dev.device.type and tdev.index in (None, samples_dev.device.index):
            return target.pdf_torch(samples_dev.clone()).squeeze(), False
        return target.pdf_torch(self._shard(samples_host).clone().to(tdev)).squeeze(), False

    def _device_target(self):
        """The reference's VAE (vae/vae.py) handed in as ``target_dist`` is evaluated on the device
        (target_decoder.DeviceTarget); any other ``pdf_torch`` provider is called where it lives."""
        td = self.target_dist
        if isinstance(td, DeviceTarget) or not is_decoder_model(td):
            return td
        if not decoder_supported(td):  # e.g. a deeper hidden_dim list or use_chunk_decode: the model's own pdf_torch
            return td
        cached = getattr(self, "_wrapped_target", None)
        if cached is None or cached.model is not td:
            cached = self._wrapped_target = DeviceTarget(td, self.cuda)
        return cached

    def _shard(self, t):
        lo, hi = self.group.shard_bounds(t.shape[0])
        return t[lo:hi]

    def _target_on_device(self, samples_host, samples_dev, scale_spec, temp, uniform=False, plot=False, spread=None):
        """get_target_dist (klerg.py:452-486) -> (p, p_stats) on the device for this rank's sample slice.
        ``spread``: max_j psi over all buffer rows, when the caller already has it (fused history pass)."""
        p_raw, pre_renorm = self._pdf(samples_host, uniform, samples_dev)
        p_raw = p_raw.detach().to(torch.float32).to(self.cuda, non_blocking=True).contiguous()
        n_total = samples_host.shape[0]
        lo = self.robot_lim[:, 0].tolist()
        hi = self.robot_lim[:, 1].tolist()
        if pre_renorm:  # uniform branch renormalises before weighting (klerg.py:456-458)
            p_raw, _, _ = engine.target_weight(2, samples_dev, lo, hi, None, p_raw, n_total, 1.0, True, self.group)
        weighted = self.weight_env or self.weight_temp or plot
        if not weighted:
            spread = None
        elif spread is None and len(self.memory_buffer) > 0:
            spec = cabi.kernel_spec(len(self.explr_idx), self.planner.num_states, self.explr_idx.tolist(),
                                    self.std.tolist(), 1.0)  # NB explr_idx, not explr_locs (klerg.py:473)
            packed = scale_spec["packed_std"]
            out, _ = engine.footprint(spec, 1, self.memory_buffer.get_all_device(), packed, samples_dev.shape[0])
            spread = out[0]
        if not weighted:
            mode = 2
        elif self.weight_env and not plot:
            mode = 1
        else:
            mode = 0
        p, p_stats, _ = engine.target_weight(mode, samples_dev, lo, hi, spread, p_raw, n_total, temp, weighted,
                                             self.group)
        return p, p_stats

    def _fused_history_ok(self, hist_dev):
        """Both memory-buffer passes in one: needs the spread at all (a weighting flag), rows to visit, and the same
        kernel for both (the reference passes explr_idx to one and explr_locs to the other, klerg.py:473 vs :496)."""
        return bool((self.weight_env or self.weight_temp) and hist_dev.shape[0] > 0
                    and self.explr_idx.tolist() == self.explr_locs.tolist())

    # ------------------------------------------------------------------ planner
    def _context(self):
        bar = self.barrier.spec()
        ctx = PlannerContext(self.planner.spec, bar, self.explr_locs.tolist(), self.horizon,
                             torch.diagonal(self.R_inv).tolist(), self.control_lim[:, 0].tolist(),
                             self.control_lim[:, 1].tolist(), alpha=self.alpha, group=self.group)
        return ctx

    def _costs(self, ctx, U_host):
        """Batched get_cost: U_host [B,H,A] on the host -> [B] costs on the host (one sync)."""
        B = U_host.shape[0]
        U = U_host.to(self.cuda, non_blocking=True)
        c = ctx.costs(U, view=True)
        if ctx.fused and B <= ctx.buf.max_g:  # costs + the kernel's copy of the fault word in one D2H
            pack = ctx.buf.cost_pack.cpu()
            if pack[-1] != 0:
                raise RuntimeError("klerg_eval_costs: an in-kernel wait timed out (lost peer or a launch that was not "
                                   "co-resident); the costs of this step are void")
            c = pack[:B].clone()
        else:
            c = c.cpu()
        self.stats["cost_evals"] += B
        return c

    def kldiv_planner(self, num_target_samples, num_traj_samples, temp=1.0):
        # u* = tanh((u + alpha du) / 0.1) * control_lim[:, 1] instead of the clamp (klerg.py:342-349, 519-522)
        engine.set_saturate(0.1 if self.saturate else 0.0)
        with engine.nvtx_range("klerg.get_samples"):
            samples, hist_dev, nu = self.get_samples(num_target_samples, num_traj_samples)
        with torch.no_grad():
            H = self.horizon
            prev = getattr(self, "ctx", None)
            ctx = self.ctx = self._context()
            if prev is not None:
                ctx.buf = prev.buf  # eval scratch / output buffers are reused from step to step when the shapes match
            if isinstance(samples, DeviceSamples):
                samples_dev = samples.local
            else:
                samples_dev = self._shard(samples).to(self.cuda, non_blocking=True).contiguous()
            with engine.nvtx_range("klerg.pack_samples"):
                ctx.set_samples(samples_dev, self.std.tolist(), 1.0, n_total=samples.shape[0])
                ctx.set_state(self.robot.state.to(self.cuda, non_blocking=True))
            spread = None
            with engine.nvtx_range("klerg.history_footprint"):
                if self._fused_history_ok(hist_dev):
                    # the spread of get_target_dist (all buffer rows, klerg.py:470-475) and the history footprint q_base
                    # (the drawn rows, klerg.py:496) visit the same rows: ONE pass over the squared distances
                    rows, t_sum = self.memory_buffer.partition_device(self.last_hist_idx)
                    ctx.q_base, spread, _ = engine.footprint_sum_max(ctx.spec, rows, t_sum, ctx.packed, ctx.n)
                else:
                    ctx.set_history(hist_dev)
            with engine.nvtx_range("klerg.target_density"):
                p, p_stats = self._target_on_device(samples, samples_dev, dict(packed_std=ctx.packed), temp,
                                                    uniform=self.uniform_tdist, spread=spread)
                ctx.set_target(p, p_stats)

            feedback = bool(getattr(self.policy, "feedback", False))  # BarrierPush / LQR: dmu/dx != 0
            if self.device_loop and ctx.fused and self.plot_data is None and self.ctrlAppSearch and not feedback:
                with engine.nvtx_range("klerg.plan_optimize"):
                    self._optimize_on_device(ctx)
                return
            last_cost = self._costs(ctx, self.u.unsqueeze(0))[0]
            accepted = None  # forward output (v, totals) of the last gradient eval, for plot_data
            prev_accepted = None
            for idx in range(self.num_iters_per_step):
                # forward(idx): the Roll/Zero policy replays self.u unchanged for idx >= 0
                u_tmp = self.policy.reset(None, self.u.clone(), idx)
                if feedback:  # the policy decides the controls along the closed loop (klerg.py:418-420)
                    g = ctx.gradient(u_tmp.to(self.cuda, non_blocking=True), keep=self.plot_data is not None,
                                     policy=self.policy.device_spec())
                    u_tmp = g["u_eff"].cpu()
                else:
                    g = ctx.gradient(u_tmp.to(self.cuda, non_blocking=True), keep=self.plot_data is not None)
                self.stats["grad_evals"] += 1
                prev_accepted, accepted = accepted, g
                if "host_pack" in g:  # fused eval: djdlam, u* and the fault word share one buffer -> a single D2H copy
                    pack = g["host_pack"].cpu()
                    if pack[-1] != 0:
                        raise RuntimeError("klerg_eval_gradient: an in-kernel wait timed out (lost peer or a launch that "
                                           "was not co-resident); the gradient of this step is void")
                    djdlam, u_star = pack[:H], pack[H:-1].view(H, -1)
                else:
                    djdlam = g["djdlam"].cpu()
                    u_star = g["u_star"].cpu()
                t_app = torch.argmin(djdlam).item()
                if not self.ctrlAppSearch:  # klerg.py:562-563: the whole saturated / clamped step
                    u_tmp = u_star.clone()
                    cost = self._costs(ctx, u_tmp.unsqueeze(0))[0]
                elif djdlam[t_app] < 0:
                    u_app = u_star[t_app]
                    if self.fixed_lam:
                        u_tmp[t_app:t_app + self.lam] = u_app.clone()
                        cost = self._costs(ctx, u_tmp.unsqueeze(0))[0]
                    else:
                        lam0, windows = line_search_windows(t_app, idx, H)
                        cands = self.u.unsqueeze(0).repeat(len(windows), 1, 1)
                        for k, (ti, tf) in enumerate(windows):
                            cands[k, ti:tf] = u_app
                        Js = self._costs(ctx, cands)
                        tau, success, _, chosen = line_search_select(Js, windows, idx, lam0, last_cost)
                        if success:
                            u_tmp[tau[0]:tau[1]] = u_app.clone()
                        if feedback:  # u_tmp is the policy's sequence, not self.u: its cost is an evaluation of its own
                            cost = self._costs(ctx, u_tmp.unsqueeze(0))[0]
                        elif success:
                            # get_cost(u_tmp) is the evaluation already made for that window
                            cost = Js[chosen] if chosen is not None else self._costs(ctx, u_tmp.unsqueeze(0))[0]
                        else:
                            cost = last_cost  # u_tmp == self.u: identical evaluation
                else:
                    accepted = prev_accepted
                    break
                if (idx > 0) and (last_cost <= cost):
                    accepted = prev_accepted
                    break
                last_cost = cost.clone()
                self.u = u_tmp
            self.u = torch.nan_to_num(self.u)
            self.last_cost = last_cost

            ro = engine.rollout(self.planner.spec, None, ctx.x0, self.u.to(self.cuda, non_blocking=True))
            self.last_plan = ro["traj"][0].cpu()
            self.stats["steps"] += 1

            if self.plot_data is not None:
                self.update_plots(ctx, accepted, samples, samples_dev, hist_dev, p, temp)
            wrapped = getattr(self, "_wrapped_target", None)
            if wrapped is not None:
                wrapped.check_fault()  # a timed-out wait inside the decoder kernel must not pass silently
            self._after_plan()

    device_loop = True  # run the optimisation loop on the device (one D2H per step) unless plot data is kept

    def _optimize_on_device(self, ctx):
        """Iterations, line searches and accept rules of klerg.py:505-576 decided on the device (klerg_plan_optimize):
        the same evals in the same order as the host loop below, one packed read-back instead of two per iteration."""
        H, A = self.horizon, self.planner.num_actions
        u_dev = self.u.to(self.cuda, non_blocking=True).reshape(H, A).contiguous()
        pack_dev = ctx.optimize(u_dev, self.num_iters_per_step, self.fixed_lam, getattr(self, "lam", 1))
        # the next step's speculative draw is enqueued while the loop runs: its side stream waits for everything
        # enqueued so far, so the draw kernel starts after the loop, and the host pays for the launch while it would
        # otherwise sit in the read-back below
        self._after_plan()
        pack = pack_dev.cpu()
        if pack[1] != 0:
            raise RuntimeError("klerg_plan_optimize: an in-kernel wait timed out (lost peer or a launch that was not "
                               "co-resident); the plan of this step is void")
        self.last_cost = pack[0].clone()
        self.stats["cost_evals"] += int(pack[2])
        self.stats["grad_evals"] += int(pack[3])
        ctx.evals["cost"] += int(pack[2])
        ctx.evals["grad"] += int(pack[3])
        self.u = pack[8:8 + H * A].view(H, A).clone()
        self.last_plan = pack[8 + H * A:].view(H + 1, -1).clone()
        self.stats["steps"] += 1
        wrapped = getattr(self, "_wrapped_target", None)
        if wrapped is not None:
            wrapped.check_fault()
        self._after_plan()

    # ------------------------------------------------------------------ plots (klerg.py:602-682)
    @torch.no_grad()
    def check_plots(self):
        return None

    def _gather_full(self, t_dev):
        if self.group.world == 1:
            return t_dev.cpu()
        import torch.distributed as dist
        n_local = torch.tensor([t_dev.shape[0]], device=self.cuda)
        sizes = self.group.gather_blocks(n_local).reshape(-1).tolist()
        pad = torch.zeros(max(sizes), dtype=t_dev.dtype, device=self.cuda)
        pad[: t_dev.shape[0]] = t_dev
        allv = self.group.gather_blocks(pad)
        return torch.cat([allv[r, : sizes[r]] for r in range(self.group.world)]).cpu()

    @torch.no_grad()
    def update_plots(self, ctx, accepted, samples, samples_dev, hist_dev, p_dev, temp):
        n = samples_dev.shape[0]
        if accepted is not None:
            q_dev = ctx.q_from(accepted["v"], accepted["totals"])
            traj_dev = torch.vstack([hist_dev, accepted["traj"]])
        else:
            _, tot = engine.footprint(ctx.spec, 0, hist_dev[:0], ctx.packed, n, add_in=ctx.q_base)
            q_dev = ctx.q_from(ctx.q_base, self.group.gather_blocks(tot))
            traj_dev = hist_dev

        def smooth_pair(plot_idx):
            ps = self.robot.state[self.explr_locs].expand_as(samples).clone()
            ps[:, plot_idx] = samples[:, plot_idx].clone()
            ps_dev = self._shard(ps).to(self.cuda).contiguous()
            spec_std = cabi.kernel_spec(len(self.explr_locs), self.planner.num_states, self.explr_locs.tolist(),
                                        self.std.tolist(), 1.0)
            pp, _ = self._target_on_device(ps, ps_dev, dict(packed_std=engine.pack_samples(spec_std, ps_dev)), temp,
                                           plot=True)
            spec_plot = cabi.kernel_spec(len(self.explr_locs), self.planner.num_states, self.explr_locs.tolist(),
                                         self.std_plot.tolist(), 1.0)
            out, tot = engine.footprint(spec_plot, 0, traj_dev, engine.pack_samples(spec_plot, ps_dev), n)
            qp = engine.renormalize_sharded(out[0, :n], self.group.gather_blocks(tot).reshape(self.group.world, -1),
                                            ctx.floor)
            pp, qp = self._gather_full(pp), self._gather_full(qp)
            if self.test_corners:
                return pp, qp
            return torch.hstack([pp, self.corners * torch.min(pp)]), torch.hstack([qp, self.corners * torch.min(qp)])

        if self.plot_extra:
            pairs = [smooth_pair(pi) for pi in self.all_plot_idx]
            self.extra_pplot, self.extra_qplot = [a for a, _ in pairs], [b for _, b in pairs]
        elif self.plot_smooth:
            self.extra_pplot, self.extra_qplot = smooth_pair(self.plot_idx)
        if self.uniform_tdist:
            p_dev, _ = self._target_on_device(samples, samples_dev, dict(packed_std=ctx.packed), temp, uniform=False,
                                              plot=True)
        p, q = self._gather_full(p_dev), self._gather_full(q_dev)
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
        # D_KL of the displayed p, q (klerg.py:679-682), computed on the device
        v_q = q_dev.unsqueeze(0).contiguous()
        _, tot = engine.footprint(ctx.spec, 0, hist_dev[:0], ctx.packed, n, add_in=q_dev)
        p_stats = engine.combine_blocks(self.group.gather_blocks(engine.vector_stats(p_dev)[:1]), [cabi.RED_SUM])
        dkl = engine.kl_cost(v_q, n, self.group.gather_blocks(tot), p_dev, p_stats, None, self.group, floor=0.0)
        self.plot_data[6] = dkl[0].cpu()
