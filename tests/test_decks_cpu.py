"""The deck-shaped test states (tests/decks.py) on the CPU: quasi-neutral, every particle
inside its patch, and a few steps of the oracle conserve the particle number and the total
energy -- so that the GPU-marked 1000-step comparisons start from sane inputs."""
import numpy as np
import pytest

import oracle_lib as ol
from decks import DECKS


@pytest.mark.parametrize("name", list(DECKS))
def test_deck_state_is_sane(name):
    d = DECKS[name]()
    og, prts, off, flds = d["og"], d["prts"], d["off"], d["flds"]
    assert off[-1] == len(prts) and len(off) == og.n_patches + 1
    # quasi-neutral (the loaders neutralise cell by cell; weights round in float)
    assert abs(float(prts["qni_wni"].astype(np.float64).sum())) < 1e-3 * len(prts)
    # patch-relative positions inside the patch, cell centres in the variant directions
    for dim in range(3):
        hi = og.ldims[dim] * og.dx[dim]
        assert prts["x"][:, dim].min() >= 0 and prts["x"][:, dim].max() < hi
    assert np.isfinite(flds).all() and np.isfinite(prts["u"]).all()
    assert set(np.unique(prts["kind"])) <= set(range(len(og.kinds)))
    # sortable: every particle indexes into its patch
    p, o = prts.copy(), off.copy()
    assert ol.lib().po_sort(og.byref(), ol.ptr(p), ol.ptr(o), None) == 0

    f = flds.copy()
    ol.fill_ghosts(og, f, 0, 9)
    e0 = ol.energies(og, f, p, o).sum()
    for s in range(1, 21):
        p, o = ol.step(og, f, p, o, sort_now=(s % d["sort_interval"] == 0))
    assert len(p) == len(prts)
    e1 = ol.energies(og, f, p, o).sum()
    assert abs(e1 / e0 - 1) < 5e-2
