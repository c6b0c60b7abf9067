"""The "chaos" argument behind the deck tolerances as an ASSERTION (CPU, oracle only): the
oracle differs from itself in the noise-driven field energy when only the particle order --
i.e. the summation order of the deposited current -- changes, by the amounts recorded in
tests/deck_tolerances.py, while the particle energies stay far inside the 1 % contract."""
import numpy as np
import pytest

from deck_tolerances import SELF_DEV, SELF_DEV_PARTICLES, field_rtol
from decks import DECKS
from test_gpu_decks import _oracle_run


@pytest.mark.parametrize("name", list(SELF_DEV))
def test_oracle_differs_from_itself_when_only_the_particle_order_changes(name):
    d = DECKS[name]()
    ref, n = _oracle_run(d, 1000)
    rng = np.random.default_rng(5)
    p, off = d["prts"].copy(), d["off"]
    for k in range(len(off) - 1):
        rng.shuffle(p[off[k]:off[k + 1]])
    ref2, n2 = _oracle_run(dict(d, prts=p), 1000)
    assert n == n2
    f1, f2 = ref[:, :6].sum(axis=1), ref2[:, :6].sum(axis=1)
    dev = np.abs(f2[1:] / f1[1:] - 1).max()
    # the recorded self-deviation is what the oracle shows today (within a factor two) ...
    assert 0.5 * SELF_DEV[name] <= dev <= 2. * SELF_DEV[name], dev
    # ... it is above the 1 % a naive reading of the contract would demand of the device,
    # and the bound the device is held to has room above it
    assert dev > 1e-2 and field_rtol(name) >= 2. * dev
    # measured against the total energy the same wander is tiny: that bound stays at 1 %
    assert np.abs(f2 - f1).max() < 2e-3 * ref.sum(axis=1).min()
    # particle energies: reordering moves them by a fraction of a per mille
    dp = np.abs(ref2[:, 6:].sum(axis=1) / ref[:, 6:].sum(axis=1) - 1).max()
    assert dp <= 3. * SELF_DEV_PARTICLES[name] and dp < 2e-3, dp
