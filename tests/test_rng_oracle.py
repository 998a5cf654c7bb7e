"""oracle/mt19937_oracle.py (numpy restatement of torch's CPU generator: MT19937 + the float32 uniform transform)
against torch itself - the generator is a third-party dependency of the reference, so the oracle is pinned on the
reference's own call, ``Uniform(low, high).sample((N,))`` (klerg.py:173,375), run here."""
import numpy as np
import pytest
import torch

from oracle import mt19937_oracle as mo


def _state():
    return torch.get_rng_state().numpy().tobytes()


@pytest.mark.parametrize("seed,warm,n,D", [(0, 0, 1000, 3), (5, 17, 7, 2), (11, 3, 624 * 3 + 5, 6), (3, 623, 100_003, 6),
                                           (9, 624, 1, 1), (2, 1, 0, 4)])
def test_uniform_draw_and_state_bit_exact(seed, warm, n, D):
    torch.manual_seed(seed)
    if warm:
        torch.rand(warm)
    blob = _state()
    low = torch.tensor([-1.15, -1.15, 2.1, -0.9, -2.3, 0.3][:D])
    high = low + torch.tensor([2.3, 2.3, 1.7, 1.8, 4.6, 0.2][:D])
    want = torch.distributions.Uniform(low, high).sample((n,))
    after = _state()
    got, blob2 = mo.uniform_samples(blob, n, low.numpy(), high.numpy())
    assert np.array_equal(got, want.numpy().reshape(n, D))
    assert blob2 == after
    # the host generator continues identically from the packed state (what the memory buffer's randperm sees)
    torch.set_rng_state(torch.tensor(list(blob2), dtype=torch.uint8))
    a = torch.randperm(97)
    torch.set_rng_state(torch.tensor(list(after), dtype=torch.uint8))
    assert torch.equal(a, torch.randperm(97))


def test_block_regeneration_matches_known_answer():
    """First outputs of MT19937 seeded with 5489 (the reference implementation's default seed): 3499211612, ..."""
    torch.manual_seed(5489)
    state, left, nxt = mo.parse_state(_state())
    out, *_ = mo.draw_u32(state, left, nxt, 3)
    assert out.tolist() == [3499211612, 581869302, 3890346734]
