"""Error behaviour of the C ABI (no GPU needed: arguments are validated on the host before any launch).
The reference raises Python exceptions for bad arguments; the library returns a negative status and a
message through klerg_last_error(), and the ctypes layer turns that into RuntimeError."""
import ctypes as C

import pytest

from control_torch import _cabi as cabi


@pytest.fixture(scope="module")
def lib():
    return cabi.load()


def spec(D=2, S=4, scale=0.05):
    return cabi.kernel_spec(D, S, list(range(D)), [scale] * D, 1.0)


def err(lib):
    return lib.klerg_last_error().decode()


def test_kernel_spec_validation(lib):
    bad = spec()
    bad.D = 9  # > KLERG_MAX_D
    assert lib.klerg_pack_samples(C.byref(bad), None, 0, None, 0, None) == -1
    assert "out of range" in err(lib)
    bad = spec(scale=0.0)
    assert lib.klerg_pack_samples(C.byref(bad), None, 4, None, 4, None) == -1
    assert "non-zero" in err(lib)
    bad = spec()
    bad.explr[1] = 7  # column outside the state row
    assert lib.klerg_footprint(C.byref(bad), 0, None, 1, 0, 0, None, 4, 4, None, None, 4, None, None, None) == -1
    assert "explr" in err(lib)


def test_size_validation(lib):
    k = spec()
    assert lib.klerg_pack_samples(C.byref(k), None, 8, None, 6, None) == -1  # ld < N
    assert lib.klerg_footprint(C.byref(k), 2, None, 1, 0, 0, None, 4, 4, None, C.c_void_p(8), 4, C.c_void_p(8),
                               C.c_void_p(8), None) == -1  # mode must be 0 / 1
    assert "mode" in err(lib)
    assert lib.klerg_kl_gradient(C.byref(k), None, 0, None, 4, 4, None, None, None, None) == -1  # H < 1
    dyn = cabi.dyn_spec(cabi.DYN_DOUBLE, 5, 2, 0.1)  # S must be 2A
    assert lib.klerg_rollout(C.byref(dyn), None, None, None, None, 1, 3, None, None, None, None, None, None) == -1
    assert "inconsistent" in err(lib)


def test_fused_eval_validation(lib):
    k = spec(D=3, S=6)
    dyn = cabi.dyn_spec(cabi.DYN_DOUBLE, 6, 3, 0.2)
    three = cabi.farr([1.0] * 3)
    args = (C.byref(k), C.byref(dyn), None, None, None, None, None)
    # horizon out of range
    assert lib.klerg_eval_gradient(*args, 0, None, 8, 8, None, None, None, 1e-6, three, 1.0, three, three, None, None,
                                   None, None, None, None, None, None, None, None, None, None) == -1
    assert "H out of range" in err(lib)
    # more candidates than one fused launch takes
    assert lib.klerg_eval_costs(*args, 9, 4, None, 8, 8, None, None, None, 1e-6, None, None, None, None, None, None, None) == -1
    assert "G must be" in err(lib)
    # sample leading dimension must be a multiple of 4 and >= N
    assert lib.klerg_eval_costs(*args, 2, 4, None, 8, 6, None, None, None, 1e-6, None, None, None, None, None, None, None) == -1
    peers = cabi.Peers()
    peers.world, peers.rank = 9, 0
    assert lib.klerg_eval_costs(C.byref(k), C.byref(dyn), None, C.byref(peers), None, None, None, 2, 4, None, 8, 8, None,
                                None, None, 1e-6, None, None, None, None, None, None, None) == -1
    assert "world/rank" in err(lib)
    # options / emulation plumbing (no launches)
    assert lib.klerg_set_option(99, 1) == -1
    assert lib.klerg_set_option(cabi.OPT_GRID_LIMIT, 64) == 0 and lib.klerg_get_option(cabi.OPT_GRID_LIMIT) == 64
    assert lib.klerg_set_option(cabi.OPT_GRID_LIMIT, 0) == 0
    assert lib.klerg_emu_launch(None) == -1  # nothing recorded
    assert "emulation" in err(lib)


def test_python_layer_raises_without_gpu_or_on_status():
    import torch
    with pytest.raises(RuntimeError, match="status -1"):
        cabi.check(-1, "klerg_something")
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            cabi.require_cuda()
        from control_torch import klerg_utils as ku
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            ku.traj_footprint_vec(torch.zeros(3, 4), torch.zeros(5, 2), [0, 1], torch.tensor([0.1, 0.1]), 1.0)


def test_every_declared_symbol_is_exported_and_bound(lib):
    """include/klerg_b200.h is the boundary: every function it declares is exported by the .so and has a ctypes
    prototype in _cabi.SIGNATURES (and nothing is bound that the header does not declare)."""
    import os
    import re
    header = open(os.path.join(cabi.INCLUDE, "klerg_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(klerg_[a-z0-9_]+)\s*\(", header))
    assert len(declared) > 30
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in klerg_b200.h but not exported"
    assert declared == set(cabi.SIGNATURES), declared ^ set(cabi.SIGNATURES)


def test_target_decoder_validation(lib):
    """VAE target decoder (vae/vae.py:244-275): shapes the tensor-core kernel is not built for are refused loudly."""
    assert lib.klerg_target_decoder_packed_bytes(3, 16, 1, 256, 512, 1) > 4 * 256 * 512 * 2
    assert lib.klerg_target_decoder_packed_bytes(3, 16, 1, 256, 100, 1) == 0  # second hidden width % 32
    assert "multiple of 32" in err(lib)
    assert lib.klerg_target_decoder_packed_bytes(3, 16, 1, 256, 288, 1) == 0  # two accumulator passes need h2 % 64
    assert lib.klerg_target_decoder_packed_bytes(3, 16, 1, 256, 320, 1) > 0
    assert lib.klerg_target_decoder_packed_bytes(9, 16, 1, 256, 512, 1) == 0
    assert "s_dim" in err(lib)
    assert lib.klerg_target_decoder_packed_bytes(3, 16, 1, 250, 512, 1) == 0
    assert lib.klerg_target_decoder_packed_bytes(3, 16, 1, 256, 512, 16) == 0
    assert "ylogvar_dim" in err(lib)
    assert lib.klerg_target_decoder_pdf(None, 3, 16, 1, 256, 512, 1, None, 5, None, -10.0, 2.0, None, None, None) == -1
    assert "null" in err(lib)
    assert lib.klerg_target_decoder_pdf(None, 3, 16, 1, 256, 512, 1, None, 0, None, -10.0, 2.0, None, None, None) == 0  # empty
    assert lib.klerg_target_decoder_pdf(C.c_void_p(64), 3, 16, 1, 256, 512, 1, C.c_void_p(64), 5, None, -10.0, 2.0,
                                        C.c_void_p(64), None, None) == -1
    assert "aligned" in err(lib)
    assert lib.klerg_target_decoder_pack(None, None, None, None, None, None, None, 3, 16, 1, 256, 512, 1, None, None) == -1


def test_targets_gradient_validation(lib):
    """Shared-psi K-target gradient: the limits of the tensor-core kernel are checked before any launch."""
    k = spec(D=3, S=6)
    dummy = C.c_void_p(256)
    args_ok = (dummy, 10, dummy, 100, 100, dummy, dummy, 1, dummy)
    assert lib.klerg_kl_gradient_targets(C.byref(k), dummy, 65, dummy, 100, 100, dummy, dummy, 1, dummy, 4, 100, 1e-6,
                                         dummy, None, dummy, None, None) == -2  # H > 64
    assert "H must be" in err(lib)
    assert lib.klerg_kl_gradient_targets(C.byref(k), *args_ok, 33, 100, 1e-6, dummy, None, dummy, None, None) == -2  # K > 32
    k6 = spec(D=6, S=12)
    assert lib.klerg_kl_gradient_targets(C.byref(k6), *args_ok, 19, 100, 1e-6, dummy, None, dummy, None, None) == -2  # 19*7 rows
    assert "128 rows" in err(lib)
    assert lib.klerg_kl_gradient_targets(C.byref(k), dummy, 10, dummy, 100, 98, dummy, dummy, 1, dummy, 4, 100, 1e-6,
                                         dummy, None, dummy, None, None) == -1  # ld < N
    assert lib.klerg_kl_gradient_targets(C.byref(k), None, 10, dummy, 100, 100, dummy, dummy, 1, dummy, 4, 100, 1e-6,
                                         dummy, None, dummy, None, None) == -1  # null states
    assert lib.klerg_kl_gradient_targets_scratch_bytes(50, 16) >= 128 * 64 * 4
    dyn = cabi.dyn_spec(cabi.DYN_DOUBLE, 6, 3, 0.2)
    three = cabi.farr([1.0] * 3)
    assert lib.klerg_adjoint_targets(C.byref(dyn), C.byref(k), 10, 0, None, 1, None, None, None, None, three, 1.0, three,
                                     three, None, None, None, None, None) == -1  # K < 1
    assert "K out of range" in err(lib)


def test_policy_entry_points_validation(lib):
    """klerg_policy_rollout / klerg_adjoint_policy (state-feedback default policies): argument checks run before any
    launch, so they can be exercised without a GPU."""
    dyn = cabi.dyn_spec(cabi.DYN_DOUBLE, 4, 2, 0.1)
    pol = cabi.policy_spec(cabi.POLICY_LQR, K=[[1.0, 0.0, 0.5, 0.0], [0.0, 1.0, 0.0, 0.5]])
    assert lib.klerg_policy_rollout(C.byref(dyn), None, None, None, None, 10, None, None, None) != 0
    assert b"policy spec is null" in lib.klerg_last_error()
    bad = cabi.policy_spec(7)
    assert lib.klerg_policy_rollout(C.byref(dyn), C.byref(bad), None, None, None, 10, None, None, None) != 0
    assert b"unknown kind" in lib.klerg_last_error()
    assert lib.klerg_policy_rollout(C.byref(dyn), C.byref(pol), None, None, None, 0, None, None, None) != 0
    assert b"H out of range" in lib.klerg_last_error()
    assert lib.klerg_policy_rollout(C.byref(dyn), C.byref(pol), None, None, None, 10, None, None, None) != 0
    assert b"null argument" in lib.klerg_last_error()
    speed = cabi.dyn_spec(cabi.DYN_SPEED, 6, 2, 0.1)
    assert lib.klerg_policy_rollout(C.byref(speed), C.byref(pol), None, None, None, 10, None, None, None) != 0
    assert b"speed-state model" in lib.klerg_last_error()
    assert lib.klerg_adjoint_policy(C.byref(speed), 10, None, None, None, None, None, None, 1.0, None, None, None, None, None, None) != 0
    assert b"speed-state model" in lib.klerg_last_error()
    assert lib.klerg_adjoint_policy(C.byref(dyn), 10, None, None, None, None, None, None, 1.0, None, None, None, None, None, None) != 0
    assert b"null argument" in lib.klerg_last_error()
    with pytest.raises(ValueError):
        cabi.policy_spec(cabi.POLICY_LQR, K=[[0.0] * 25] * 8)
