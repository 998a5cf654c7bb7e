"""GPU parity of the target-density decoder (klerg_target_decoder_pdf, tcgen05 3xTF32) against the vectors recorded
from the reference's VAE.pdf_torch and against the CPU oracle.  Tolerance: 1e-4 relative (BASELINE.json north_star;
the 3xTF32 split lands around 1e-6)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from cases import TARGET_CASES, DecoderModel, MixtureTarget, ROBOT_CASES, robot_kwargs, seed_buffer_states, target_samples  # noqa: E402
from oracle import klerg_oracle as ko  # noqa: E402

RTOL = 1e-4
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def model_for(name, **kw):
    case = TARGET_CASES[name]
    g = np.load(os.path.join(GOLD, f"target_{name}.npz"))
    return case, g, DecoderModel(case, g["z_rows"], g["seed_x"], out_extra=int(g["out_features"]) - case["nl"], **kw)


@pytest.fixture(scope="module")
def td():
    import control_torch.target_decoder as m
    return m


@pytest.mark.parametrize("name", sorted(TARGET_CASES))
def test_pdf_matches_reference_vectors(td, name):
    case, g, model = model_for(name)
    dev = td.DeviceTarget(model)
    p = dev.pdf_torch(target_samples(case))
    dev.check_fault()
    assert p.is_cuda and p.shape == (case["n"],)
    np.testing.assert_allclose(p.cpu().numpy(), g["p"], rtol=RTOL, atol=0)
    # the numpy entry point (vae.py:238-242)
    np.testing.assert_allclose(dev.pdf(target_samples(case).numpy()), g["p"], rtol=RTOL, atol=0)


def test_uninitialised_model_is_uniform(td):
    case, g, model = model_for("default", initialised=False)
    p = td.DeviceTarget(model).pdf_torch(target_samples(case))
    np.testing.assert_array_equal(p.cpu().numpy(), g["p_uninit"])


@pytest.mark.parametrize("n", [0, 1, 127, 128, 129, 148 * 128 + 5, 40_000])
def test_ragged_sizes_vs_oracle(td, n):
    case = dict(TARGET_CASES["default"], n=n)
    g = torch.Generator().manual_seed(n + 1)
    model = DecoderModel(case, torch.randn(1, case["zd"], generator=g))
    dev = td.DeviceTarget(model)
    s = target_samples(case, seed=n)
    p = dev.pdf_torch(s)
    dev.check_fault()
    ref = model.pdf_torch(s)
    assert tuple(p.shape) == tuple(ref.shape)
    np.testing.assert_allclose(p.cpu().numpy(), ref.numpy(), rtol=RTOL, atol=0)


def test_weights_are_reread_every_call(td):
    """The trainer keeps updating the model behind the controller (sensor_main_module.py:311-339)."""
    case, g, model = model_for("pose6_rgbvar")
    dev = td.DeviceTarget(model)
    s = target_samples(case)
    p0 = dev.pdf_torch(s).cpu()
    with torch.no_grad():
        model.decode[2].weight.mul_(0.5)
        model.weights[1] = (model.decode[2].weight.detach().clone(), model.weights[1][1])
        model.z_samples.add_(0.3)
    p1 = dev.pdf_torch(s).cpu()
    dev.check_fault()
    assert not torch.allclose(p0, p1)
    np.testing.assert_allclose(p1.numpy(), model.pdf_torch(s).numpy(), rtol=RTOL, atol=0)
    # unchanged weights are not packed again (version counters / storage pointers), a replaced parameter is
    packs = dev.stats["packs"]
    p2 = dev.pdf_torch(s).cpu()
    assert dev.stats["packs"] == packs and torch.equal(p1, p2)
    with torch.no_grad():
        model.decode[0].bias = torch.nn.Parameter(model.decode[0].bias.detach() + 0.05)
        model.weights[0] = (model.weights[0][0], model.decode[0].bias.detach().clone())
    p3 = dev.pdf_torch(s).cpu()
    assert dev.stats["packs"] == packs + 1 and not torch.allclose(p2, p3)
    np.testing.assert_allclose(p3.numpy(), model.pdf_torch(s).numpy(), rtol=RTOL, atol=0)


def test_unsupported_decoder_raises(td):
    case, g, model = model_for("default")
    model.decode = torch.nn.Sequential(torch.nn.Linear(19, 8), torch.nn.Tanh(), torch.nn.Linear(8, 4))
    with pytest.raises(NotImplementedError):
        td.DeviceTarget(model).pdf_torch(target_samples(case))
    for hidden in ([100, 40], [288, 40]):  # h2 = 100: not a multiple of 32; 288: two passes of 144 columns
        case2 = dict(TARGET_CASES["default"], hidden=hidden)
        m2 = DecoderModel(case2, torch.zeros(1, case2["zd"]))
        with pytest.raises(RuntimeError, match="second hidden width"):
            td.DeviceTarget(m2).pdf_torch(target_samples(case2))


def test_robot_step_with_model_target():
    """Robot(target_dist=<VAE-like model>) evaluates p on the device and steps like the oracle controller that calls
    the model's own pdf_torch on the CPU."""
    from control_torch.klerg import Robot
    rc = ROBOT_CASES["xyz_small"]
    case = dict(TARGET_CASES["default"])
    out = []
    for cls in (ko.OracleRobot, Robot):
        torch.manual_seed(7)
        model = DecoderModel(case, torch.randn(1, case["zd"], generator=torch.Generator().manual_seed(2)))
        r = cls(**robot_kwargs(rc, model))
        r.test(rc["n"])
        for s in seed_buffer_states(r.robot.state, rc):
            r.memory_buffer.push(s)
        res = [r.step(rc["n"], rc["m"], save_update=True) for _ in range(3)]
        out.append((res, r.u.clone()))
        if cls is Robot:
            assert r._wrapped_target.stats["evals"] >= 3, "the device decoder was not used"
            r._wrapped_target.check_fault()
    (ra, ua), (rb, ub) = out
    for (sa, va, ca), (sb, vb, cb) in zip(ra, rb):
        np.testing.assert_allclose(sb, sa, rtol=1e-3, atol=1e-5)
        np.testing.assert_allclose(cb, ca, rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(ub.numpy(), ua.numpy(), rtol=1e-3, atol=1e-5)


def test_full_size_properties(td):
    """BASELINE sizes (1e7 samples, default decoder): size-independent properties - every value inside the clamp
    image, tile-permutation invariance (p is point-wise), and a strided subset equal to the oracle."""
    case = dict(TARGET_CASES["default"], n=10_000_000)
    model = DecoderModel(case, torch.randn(1, case["zd"], generator=torch.Generator().manual_seed(5)))
    dev = td.DeviceTarget(model)
    g = torch.Generator(device="cuda").manual_seed(1)
    s = torch.rand(case["n"], case["sd"], generator=g, device="cuda") * 2.3 - 1.15
    p = dev.pdf_torch(s)
    dev.check_fault()
    assert float(p.min()) >= np.exp(-10.0) * (1 - 1e-6) and float(p.max()) <= np.exp(2.0) * (1 + 1e-6)
    perm = torch.randperm(case["n"], device="cuda", generator=g)
    p_perm = dev.pdf_torch(s[perm])
    assert torch.equal(p_perm, p[perm]), "p must not depend on where a sample sits in a tile"
    idx = torch.arange(0, case["n"], 9973, device="cuda")
    ref = model.pdf_torch(s[idx].cpu())
    np.testing.assert_allclose(p[idx].cpu().numpy(), ref.numpy(), rtol=RTOL, atol=0)


def test_trainer_spread_grade_block(td):
    """the trainer's per-iteration spread / grade block (dist_modules/trainer_module.py:511-538)
    written exactly as the reference writes it, but on CUDA tensors with this package's drop-ins
    (control_torch.klerg_utils.traj_spread_vec and the model wrapped in DeviceTarget), against the CPU oracle."""
    from control_torch.klerg_utils import traj_spread_vec
    from oracle import target_oracle
    case = dict(TARGET_CASES["default"], n=5000)
    g = torch.Generator().manual_seed(12)
    model = DecoderModel(case, torch.randn(1, case["zd"], generator=g))
    samples = target_samples(case, seed=4)
    traj = torch.rand(700, 3, generator=g) * 2 - 1
    std = 0.05
    ref = target_oracle.trainer_spread_grade(traj, samples, std, model.pdf_torch)

    device = torch.device("cuda")
    dev_model = td.DeviceTarget(model)
    s_dev, t_dev = samples.to(device), traj.to(device)
    # ---- trainer_module.py:511-538, verbatim modulo `self.` ----
    dim = s_dev.shape[1]
    explr_idx = torch.arange(dim, device=device)
    std_t = torch.tensor([std] * dim, device=device)
    max_q = traj_spread_vec(t_dev, s_dev, explr_idx, std_t, nu=1.)
    max_q /= torch.max(max_q)
    spread = max_q.mean()
    entropy_dist = dev_model.pdf_torch(s_dev)
    entropy_dist = entropy_dist**spread
    entropy_dist /= entropy_dist.max()
    grade = torch.clamp(10.**(-torch.log10(entropy_dist.min()) - 4), max=0.01)
    # ----
    dev_model.check_fault()
    assert max_q.is_cuda and entropy_dist.is_cuda
    np.testing.assert_allclose(max_q.cpu().numpy(), ref[2].numpy(), rtol=RTOL, atol=1e-30)
    np.testing.assert_allclose(float(spread), float(ref[0]), rtol=RTOL)
    np.testing.assert_allclose(entropy_dist.cpu().numpy(), ref[3].numpy(), rtol=RTOL)
    np.testing.assert_allclose(float(grade), float(ref[1]), rtol=1e-3)


SHAPES = [
    # sd, zd, hidden (reference order: [h2, h1]), nl, zbuf, n
    (1, 1, [32, 8], 1, 0, 300),        # smallest widths
    (7, 3, [512, 1024], 1, 0, 700),    # widest first layer the kernel takes (32 operand stages per pass), 7 conditioning dims
    (3, 16, [256, 72], 2, 0, 500),     # single accumulator pass, K-steps not a multiple of the stage (9 = 4 + 4 + 1)
    (2, 5, [320, 40], 3, 4, 260),      # two passes of 160 columns, z buffer of 4
    (6, 16, [512, 256], 15, 0, 390),   # most logvar columns, wide conditioning
    (4, 8, [448, 136], 5, 6, 129),     # 224-column passes, z buffer of 6 with the wide tables (ring shrinks to fit)
    (3, 16, [512, 256], 1, 8, 1000),   # the default decoder with a full z buffer of 8
]


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "sd%d_z%d_h%dx%d_nl%d_zb%d" % (s[0], s[1], s[2][1], s[2][0], s[3], s[4]))
def test_decoder_shapes_vs_oracle(td, shape):
    """Every code path of the kernel that depends on the decoder's shape: one or two accumulator passes, short last
    operand stage, 4- or 8-wide first-layer rows, 4- or 16-wide epilogue rows, packed single-column epilogue, z buffers."""
    sd, zd, hidden, nl, zbuf, n = shape
    case = dict(sd=sd, zd=zd, hidden=hidden, nl=nl, zbuf=zbuf, dx=bool(zbuf), n=n, gain=2.5)
    g = torch.Generator().manual_seed(sum(hidden) + nl)
    z = torch.randn(max(zbuf, 1), zd, generator=g)
    model = DecoderModel(case, z, seed_x=torch.rand(1, sd, generator=g) * 0.2)
    dev = td.DeviceTarget(model)
    s = target_samples(case, seed=n)
    p = dev.pdf_torch(s)
    dev.check_fault()
    ref = model.pdf_torch(s)
    assert float(ref.std()) > 0, "degenerate case: the clamp swallowed everything"
    np.testing.assert_allclose(p.cpu().numpy(), ref.numpy(), rtol=RTOL, atol=0)


def test_tables_too_large_is_refused(td):
    """A z buffer whose first-layer tables cannot share the SM with one W2 ring raises instead of mis-computing."""
    case = dict(sd=6, zd=4, hidden=[512, 1024], nl=1, zbuf=8, dx=False, n=64, gain=1.0)
    model = DecoderModel(case, torch.zeros(8, 4))
    with pytest.raises(RuntimeError, match="do not fit"):
        td.DeviceTarget(model).pdf_torch(target_samples(case))


def test_robot_falls_back_for_decoders_the_kernel_is_not_built_for(td):
    """A VAE the reference accepts but the kernel does not cover (here: a Tanh decoder with two layers) is evaluated
    through the model's own pdf_torch - Robot.step() must keep working (only DeviceTarget itself refuses)."""
    from control_torch.klerg import Robot
    rc = ROBOT_CASES["xyz_small"]
    case = dict(TARGET_CASES["default"])
    model = DecoderModel(case, torch.zeros(1, case["zd"]))
    assert td.decoder_supported(model)
    model.decode = torch.nn.Sequential(torch.nn.Linear(19, 8), torch.nn.Tanh(), torch.nn.Linear(8, 4))
    assert not td.decoder_supported(model)
    calls = []
    model.pdf_torch = lambda s: (calls.append(s.shape[0]), torch.full((s.shape[0],), 0.5) + 0.1 * s[:, 0].cpu().abs())[1]
    torch.manual_seed(7)
    r = Robot(**robot_kwargs(rc, model))
    r.test(rc["n"])
    assert r._device_target() is model
    out = r.step(rc["n"], rc["m"], save_update=True)
    assert calls and all(np.isfinite(o).all() for o in out)
