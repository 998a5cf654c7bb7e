"""CPU: the target-density oracle against vectors recorded from the live reference VAE
(tests/golden/make_golden_target.py), and against the live reference when it is present."""
import os

import numpy as np
import pytest
import torch

from cases import TARGET_CASES, DecoderModel, decoder_weights, target_samples
from oracle import target_oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_case(name):
    case = TARGET_CASES[name]
    g = np.load(os.path.join(GOLD, f"target_{name}.npz"))
    ws = decoder_weights(case, out_extra=int(g["out_features"]) - case["nl"])
    return case, g, ws


@pytest.mark.parametrize("name", sorted(TARGET_CASES))
def test_oracle_matches_reference_vectors(name):
    case, g, ws = load_case(name)
    samples = target_samples(case)
    z = torch.from_numpy(g["z_rows"])
    shift = torch.from_numpy(g["seed_x"]) if case["dx"] else None
    p = target_oracle.vae_pdf(samples, ws, z, case["nl"], (-10, 2), shift)
    # same torch operators on the same host: identical up to GEMM blocking of the two builds of the weights
    np.testing.assert_allclose(p.numpy(), g["p"], rtol=2e-6, atol=0)
    p0 = target_oracle.vae_pdf(samples, ws, z, case["nl"], (-10, 2), shift, initialised=False)
    np.testing.assert_array_equal(p0.numpy(), g["p_uninit"])


@pytest.mark.parametrize("name", sorted(TARGET_CASES))
def test_decoder_model_stub_matches_reference_vectors(name):
    """The stand-in model the GPU tests wrap reproduces the reference's pdf_torch."""
    case, g, ws = load_case(name)
    m = DecoderModel(case, g["z_rows"], g["seed_x"], out_extra=int(g["out_features"]) - case["nl"])
    np.testing.assert_allclose(m.pdf_torch(target_samples(case)).numpy(), g["p"], rtol=2e-6, atol=0)


@pytest.mark.skipif(not os.path.isdir("/root/reference/franka_test/scripts/vae"), reason="live reference not present")
def test_oracle_matches_live_reference():
    import sys
    sys.path.insert(0, GOLD)
    import make_golden_target as mg
    VAE = mg.import_reference_vae()
    case = TARGET_CASES["zbuffer_dx"]
    vae, ws = mg.build_reference(VAE, case)
    g = torch.Generator().manual_seed(21)
    for _ in range(2):
        vae.update_dist(torch.rand(1, case["sd"], generator=g), torch.rand(1, 3, 64, 64, generator=g))
    samples = target_samples(case, seed=8)
    with torch.no_grad():
        ref = vae.pdf_torch(samples.clone())
    p = target_oracle.vae_pdf(samples, ws, vae.z_buff.get_samples(), case["nl"], vae.logvar_lims, vae.seed_x)
    np.testing.assert_allclose(p.numpy(), ref.numpy(), rtol=2e-6, atol=0)
