"""Device binary collisions (psc_b200/csrc/collision.cu) against the CPU oracle: the reference's
known-answer test on the device (RngFake), and -- because oracle and device share the
counter-based streams -- a thermal plasma compared particle for particle."""
import numpy as np
import pytest

import oracle_lib as ol
from b200_helpers import gpu_state
from gen import thermal_plasma

pytestmark = pytest.mark.gpu


def test_collision_test1_on_device():
    """CollisionTest.Test1 (src/libpsc/tests/test_collision.cxx:130-172), eps 1e-5"""
    import psc_b200 as pb
    og = ol.Grid(gdims=(1, 16, 16), length=(160., 160., 160.), np_=(1, 1, 1), dt=1., kinds=((1., 1.),), nicell=200)
    prts = np.zeros(2, dtype=ol.PRT_DTYPE)
    prts["x"] = [[5., 5., 5.], [5., 5., 5.]]
    prts["u"] = [[1., 0., 0.], [0., 0., 0.]]
    prts["qni_wni"] = 1.
    off = np.array([0, 2], dtype=np.uint32)
    grid, mprts, _ = gpu_state(og, None, prts, off)
    grid.cori = 1. / 200
    coll = pb.Collision(grid, 1, 1., rng=0)
    assert coll(mprts, step=1) == 1
    got, _ = mprts.get()
    eps = 1e-5
    u0, u1 = got["u"][0], got["u"][1]
    assert abs(u0[0] + u1[0] - 1.) < eps and abs(u0[1] + u1[1]) < eps and abs(u0[2] + u1[2]) < eps
    assert abs(u0[0] - 0.96226911) < eps and abs(u0[1]) < eps and abs(abs(u0[2]) - 0.17342988) < eps
    assert abs(u1[0] - 0.03773088) < eps and abs(u1[1]) < eps and abs(abs(u1[2]) - 0.17342988) < eps
    grid.close()


CASES = {
    "xyz_odd": (dict(gdims=(8, 8, 8), length=(8., 8., 8.), np_=(2, 1, 1)), 7),      # triangles in every cell
    "xyz_even": (dict(gdims=(8, 8, 8), length=(8., 8., 8.), np_=(1, 2, 1)), 8),
    "yz": (dict(gdims=(1, 16, 16), length=(1., 16., 16.), np_=(1, 2, 2)), 25),
    "big_cells": (dict(gdims=(2, 2, 2), length=(2., 2., 2.), np_=(1, 1, 1)), 700),  # 1400 per cell > the 1024 cap
}


@pytest.mark.parametrize("name", list(CASES))
def test_collide_matches_oracle(name):
    import psc_b200 as pb
    gkw, ppc = CASES[name]
    kinds = ((-1., 1.), (1., 25.))
    og = ol.Grid(dt=0.5, kinds=kinds, nicell=10, **gkw)
    prts, off = thermal_plasma(og, ppc=ppc, seed=3, vth=(0.3, 0.05))
    grid, mprts, _ = gpu_state(og, None, prts, off)
    grid.cori = 0.1
    coll = pb.Collision(grid, 10, 0.3, rng=1, seed=5)
    n_gpu = coll(mprts, step=20)  # sorts the store first (CollisionHost asserts the cell order)
    got, got_off = mprts.get()

    ref, ro = prts.copy(), off.copy()
    rc, _ = ol.sort(og, ref, ro)
    assert rc == 0
    before = ref["u"].copy()
    n_ref = ol.collide(og, ref, ro, 10, 0.3, 0.1, rng=ol.RNG_HASH, seed=5, step=20)
    assert n_gpu == n_ref > 0
    assert np.array_equal(got_off, ro)
    assert got["x"].tobytes() == ref["x"].tobytes() and np.array_equal(got["kind"], ref["kind"])
    assert np.abs(ref["u"] - before).max() > 1e-3
    # same pairs, same random numbers, same operation order; the libm transcendentals
    # (atan, log, sin, cos, acos) differ by an ulp or two between glibc and CUDA
    err = np.abs(got["u"].astype(np.float64) - ref["u"]).max()
    assert err < 2e-5 * np.abs(ref["u"]).max(), err
    grid.close()


def test_psc_steps_with_collisions_conserve():
    """collisions inside the step loop (every 2nd step, after the sort: psc.hxx:356-371): the particle
    number never changes, and a collide call leaves the kinetic energy of the particles where it was
    (every pair conserves its energy; the species exchange it)"""
    import psc_b200 as pb
    from b200_helpers import make_gpu_grid
    kinds = ((-1., 1.), (1., 25.))
    og = ol.Grid(gdims=(16, 16, 16), length=(16., 16., 16.), np_=(2, 2, 1), dt=0.4, kinds=kinds, nicell=8)
    grid = make_gpu_grid(og)
    grid.cori = 1. / 8
    mprts, mflds = pb.Mparticles(grid), pb.MfieldsState(grid)
    mprts.setup_thermal(8, [0.2, 0.02], seed=3)
    coll = pb.Collision(grid, 2, 0.05, seed=11)
    psc = pb.Psc(grid, mflds, mprts, sort_interval=1, fused=True, collision=coll)
    psc.initialize()
    n0 = mprts.size()
    for _ in range(8):
        psc.step()
    assert mprts.size() == n0 and coll.n_collisions > 0
    e_before = pb.energies(grid)[6:8]
    u_before = mprts.get()[0]["u"].copy()
    coll(mprts, step=99)
    e_after = pb.energies(grid)[6:8]
    assert np.abs(mprts.get()[0]["u"] - u_before).max() > 1e-3
    assert abs(e_after.sum() - e_before.sum()) < 1e-5 * e_before.sum()
    assert abs(e_after[0] - e_before[0]) > 0  # electrons and ions did exchange energy
    grid.close()


SPOT = dict(zl=2., zh=6., xc=4., yc=5., rH=3., T=[0.04, 0., 0.01], Mi=25.)


@pytest.mark.parametrize("case", [dict(gdims=(8, 8, 8), length=(8., 8., 8.), np_=(2, 1, 2)),
                                  dict(gdims=(1, 16, 16), length=(1., 10., 8.), np_=(1, 2, 2), corner=(0., -1., 0.))],
                         ids=["xyz", "yz"])
@pytest.mark.parametrize("rH", [3., 0.])
def test_heating_matches_oracle(case, rH):
    """Heating__ + HeatingSpotFoil (psc_heating_impl.hxx:27-76, heating_spot_foil.hxx:22-89): same
    particles kicked, same kicks (shared counter-based streams; libm differs by an ulp or two)"""
    import psc_b200 as pb
    kinds = ((-1., 1.), (-1., 1.), (1., 25.))  # flatfoil's three kinds; the second is not heated (T = 0)
    og = ol.Grid(dt=0.5, kinds=kinds, nicell=10, **case)
    prts, off = thermal_plasma(og, ppc=5, seed=3, vth=(0.1, 0.1, 0.02))
    spot = dict(SPOT, rH=rH)
    grid, mprts, _ = gpu_state(og, None, prts, off)
    heat = pb.Heating(grid, 20, spot, seed=7)
    n_gpu = heat(mprts, step=40)
    got, got_off = mprts.get()
    ref = prts.copy()
    n_ref = ol.heating(og, ref, off, 20, spot, seed=7, step=40)
    assert n_gpu == n_ref and 0 < n_ref < len(prts)
    changed_ref = np.any(ref["u"] != prts["u"], axis=1)
    changed_gpu = np.any(got["u"] != prts["u"], axis=1)
    assert np.array_equal(changed_ref, changed_gpu)
    assert not changed_ref[prts["kind"] == 1].any()
    assert got["x"].tobytes() == prts["x"].tobytes()
    assert np.abs(got["u"].astype(np.float64) - ref["u"]).max() < 2e-5 * np.abs(ref["u"]).max()
    grid.close()
