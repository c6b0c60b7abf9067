"""The gapped particle store (psc_b200/csrc/gap.cuh): psc_b200_step never runs the boundary
exchange + sort as a pass of its own; the push writes every particle that stays in its
cell to its final slot and only the movers are placed afterwards.  What the store stands
for must be exactly what the reference's BndParticles + SortCountsort2 produce: with the
fields held fixed (push_fields = 0) the particle update is independent of the deposit's
summation order, so after k consecutive gapped steps the store is compared BIT-EXACT with
the CPU oracle (sort, push, exchange) x k -- records, per-patch offsets, per-cell counts.
Tight slack forces the re-layout path, a tiny mover list the redo-on-the-eager-path one."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as ol
from b200_helpers import gpu_state
from gen import random_fields, thermal_plasma

pytestmark = pytest.mark.gpu

KINDS = ((-1., 1.), (1., 100.))
PER = dict()
WALL_Z = dict(bc_fld_lo=[1, 1, 2], bc_fld_hi=[1, 1, 2], bc_prt_lo=[1, 1, 0], bc_prt_hi=[1, 1, 0])
ABSORB_Z = dict(bc_fld_lo=[1, 1, 2], bc_fld_hi=[1, 1, 2], bc_prt_lo=[1, 1, 2], bc_prt_hi=[1, 1, 2])
CASES = {
    # name: (grid kwargs, options)
    "xyz_2x1x2": (dict(gdims=(16, 16, 16), length=(16., 16., 16.), np_=(2, 1, 2)), {}),
    "xyz_single_patch": (dict(gdims=(8, 8, 8), length=(8., 8., 8.), np_=(1, 1, 1)), {}),
    "xyz_aniso_dyn": (dict(gdims=(8, 24, 16), length=(10., 20., 7.), np_=(1, 3, 1)), {}),
    "xyz_small_tiles": (dict(gdims=(16, 16, 16), length=(16., 16., 16.), np_=(2, 2, 1)), dict(tile=4)),
    "xyz_wall_z": (dict(gdims=(16, 8, 16), length=(16., 8., 16.), np_=(1, 1, 2), **WALL_Z), {}),
    "xyz_absorbing_z": (dict(gdims=(8, 8, 16), length=(8., 8., 16.), np_=(1, 1, 2), **ABSORB_Z), {}),
    "yz_var1": (dict(gdims=(1, 32, 48), length=(1., 40., 30.), np_=(1, 2, 3)), {}),
    "yz_split": (dict(gdims=(1, 32, 32), length=(1., 32., 32.), np_=(1, 2, 1), deposit=ol.DEPOSIT_SPLIT), {}),
    "yz_wall_z": (dict(gdims=(1, 16, 32), length=(1., 16., 32.), np_=(1, 1, 2), **WALL_Z), {}),
}


def _run_oracle(og, flds, prts, off, k):
    f, p, o = flds.copy(), prts.copy(), off.copy()
    L, G = ol.lib(), og.byref()
    n_drop = 0
    for _ in range(k):
        assert L.po_sort(G, ol.ptr(p), ol.ptr(o), None) == 0
        L.po_push_mprts(G, ol.ptr(f), ol.ptr(p), ol.ptr(o))
        p, o, nd = ol.bnd_particles(og, p, o)
        n_drop += nd
        L.po_bndf_add_ghosts_J(G, ol.ptr(f))
        L.po_add_ghosts(G, ol.ptr(f), 9, 0, 3)
        L.po_fill_ghosts(G, ol.ptr(f), 9, 0, 3)
    assert L.po_sort(G, ol.ptr(p), ol.ptr(o), None) == 0
    return f, p, o, n_drop


@pytest.mark.parametrize("slack", [0, 1], ids=["auto_slack", "slack1"])
@pytest.mark.parametrize("k", [1, 2, 5])
@pytest.mark.parametrize("vth", [0.05, 0.5])
@pytest.mark.parametrize("name", list(CASES))
def test_gapped_steps_bit_exact(name, vth, k, slack):
    import psc_b200 as pb
    gkw, opts = CASES[name]
    dx = [l / g for l, g in zip(gkw["length"], gkw["gdims"])]
    dt = 0.45 * min(d for d, g in zip(dx, gkw["gdims"]) if g > 1)
    og = ol.Grid(dt=dt, kinds=KINDS, nicell=6, **gkw)
    flds = random_fields(og, seed=11)
    prts, off = thermal_plasma(og, ppc=6, seed=12, vth=(vth, vth / 10))
    rf, rp, ro, n_drop = _run_oracle(og, flds, prts, off, k)

    grid, mprts, mflds = gpu_state(og, flds, prts, off, dict(opts, gapped=1, gap_slack=slack))
    prm = pb.StepParams(sort=1, marder_loop=0, marder_diffusion=0., push_fields=0, checks=0)
    for _ in range(k):
        pb.check(grid.lib.psc_b200_step(grid.ctx, C.byref(prm)))
    assert grid.get_stat("gap_steps") == k and grid.get_stat("gap_redone") == 0
    if slack == 1 and vth > 0.1 and k > 1:
        assert grid.get_stat("gap_relayouts") > 0
    # sizes are known without materialising the store
    assert np.array_equal(mprts.sizeByPatch(), np.diff(ro))
    assert mprts.size() == len(rp)
    j = mflds.download(0, 3)
    got, got_off = mprts.get()
    assert np.array_equal(got_off, ro)
    assert got.tobytes() == rp.tobytes(), "gapped store differs from sort+push+exchange of the oracle"
    assert np.array_equal(ol.count_by_cell(og, got, got_off), ol.count_by_cell(og, rp, ro))
    assert grid.get_stat("n_dropped") == n_drop
    if "absorbing" in name and vth > 0.1:
        assert n_drop > 0
    scale = np.abs(rf[:, :3]).max()
    assert np.abs(j - rf[:, :3]).max() <= 1e-5 * scale
    # the compacted store is a plain cell-ordered one: one more step from it, gapped again
    pb.check(grid.lib.psc_b200_step(grid.ctx, C.byref(prm)))
    rf2, rp2, ro2, _ = _run_oracle(og, flds, rp, ro, 1)
    got2, got_off2 = mprts.get()
    assert np.array_equal(got_off2, ro2) and got2.tobytes() == rp2.tobytes()
    grid.close()


def test_gapped_equals_eager_fused_with_fields():
    """full Psc::step (fields evolving): gapped and eager paths agree to round-off over
    several steps and conserve the particle number"""
    import psc_b200 as pb
    og = ol.Grid(gdims=(16, 16, 16), length=(16., 16., 16.), np_=(2, 2, 2), dt=0.4, kinds=KINDS, nicell=8)
    flds = random_fields(og, seed=8, amp_e=0.02, amp_b=0.05)
    ol.fill_ghosts(og, flds, 3, 9)
    prts, off = thermal_plasma(og, ppc=8, seed=9, vth=(0.2, 0.02), margin=0.05)
    res = []
    for gapped in (0, 1):
        grid, mprts, mflds = gpu_state(og, flds, prts, off, dict(gapped=gapped))
        psc = pb.Psc(grid, mflds, mprts, sort_interval=1, fused=True)
        for _ in range(6):
            psc.step()
        assert grid.get_stat("gap_steps") == (6 if gapped else 0)
        res.append((mprts.get(), mflds.download(), pb.api.energies(grid)))
        grid.close()
    (p0, o0), f0, e0 = res[0]
    (p1, o1), f1, e1 = res[1]
    assert len(p0) == len(p1) == len(prts)
    assert np.array_equal(o0, o1)
    assert np.abs(f0 - f1).max() <= 2e-5 * np.abs(f0).max()
    assert np.array_equal(p0["kind"], p1["kind"])
    assert np.abs(p0["x"] - p1["x"]).max() <= 1e-5 * 16
    assert np.abs(p0["u"] - p1["u"]).max() <= 1e-5
    np.testing.assert_allclose(e0, e1, rtol=1e-5)
