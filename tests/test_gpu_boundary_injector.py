"""BoundaryInjectorB200 on the device against the oracle's restatement of
BoundaryInjector::inject (src/include/boundary_injector.hxx:93-160): the accepted particles
byte for byte, the deposited current to float rounding (global atomics: the order of the
additions to one J value is not fixed; 1e-6 of the largest value), and the reference's own
integration tests (src/libpsc/tests/test_boundary_injector.cxx) replayed step by step."""
import numpy as np
import pytest

import oracle_lib as ol
from b200_helpers import gpu_state, make_gpu_grid
from gen import thermal_plasma
from injector_cases import TestGenerator, injector_grid_kw, run_oracle, CHECK_EPS

pytestmark = pytest.mark.gpu

J_RTOL = 1e-6

GRIDS = {
    "yz_var1": dict(gdims=(1, 16, 8), length=(1., 8., 6.), np_=(1, 2, 2)),
    "yz_split": dict(gdims=(1, 16, 8), length=(1., 8., 6.), np_=(1, 2, 2), deposit=ol.DEPOSIT_SPLIT),
    "xyz": dict(gdims=(8, 8, 4), length=(4., 6., 4.), np_=(2, 2, 1)),
}


def _open_y_grid(name, dt=0.3):
    kw = injector_grid_kw(dt=dt)
    kw.update(GRIDS[name])
    if kw["gdims"][0] > 1:
        kw["bc_fld_lo"][0] = kw["bc_fld_hi"][0] = ol.BND_FLD_PERIODIC
    return ol.Grid(**kw)


def _random_candidates(og, n_per_cell, seed):
    rng = np.random.default_rng(seed)
    cand = []
    for p in range(og.n_patches):
        if og.patch_off(p)[1] != 0:
            continue
        for i in range(og.ldims[0]):
            for k in range(og.ldims[2]):
                for _ in range(n_per_cell):
                    idx = (i, -1, k)
                    x = [(ii + rng.random()) * dx for ii, dx in zip(idx, og.dx)]
                    if og.gdims[0] == 1:
                        x[0] = 0.0
                    u = rng.normal(size=3) * 0.8 + [0., 0.9, 0.]  # some do not make it in
                    cand.append((p, idx, tuple(x), tuple(u), float(rng.uniform(0.5, 2.)), int(rng.integers(0, 2))))
    return cand


@pytest.mark.parametrize("name", list(GRIDS))
def test_inject_matches_oracle(name):
    import psc_b200 as pb
    og = _open_y_grid(name)
    prts, off = thermal_plasma(og, ppc=3, seed=2, vth=(0.2, 0.05))
    flds = og.zeros_fields()
    flds[:, ol.JXI:ol.JZI + 1] = np.random.default_rng(1).normal(size=flds[:, :3].shape).astype(np.float32) * 0.01
    cand = _random_candidates(og, 5, seed=4)
    ref = flds.copy()
    rprts, roff = ol.boundary_inject(og, ref, prts, off, cand)
    n_in = len(rprts) - len(prts)
    assert 0 < n_in < len(cand)  # some entered, some were turned away
    grid, mprts, mflds = gpu_state(og, flds, prts, off)
    inj = pb.BoundaryInjector(None, grid)
    assert inj.inject(mprts, mflds, cand=cand) == n_in
    got, got_off = mprts.get()
    assert np.array_equal(got_off, roff)
    assert got.tobytes() == rprts.tobytes()
    gf = mflds.download()
    scale = np.abs(ref[:, :3]).max()
    assert np.abs(gf - ref).max() <= J_RTOL * scale
    assert (gf[:, 3:] == flds[:, 3:]).all()
    assert np.abs(ref[:, :3] - flds[:, :3]).max() > 1e-3  # something was deposited
    grid.close()


class _Replay:
    """injector that replays the draws the oracle run recorded (one list per step)"""

    def __init__(self, grid, draws_by_step, which):
        import psc_b200 as pb
        self.inj = pb.BoundaryInjector(None, grid)
        self.draws, self.which, self.step = draws_by_step, which, 0

    def inject(self, mprts, mflds):
        self.inj.inject(mprts, mflds, cand=self.draws[self.step][self.which])
        self.step += 1


def _sorted_records(prts, off):
    out = []
    for p in range(len(off) - 1):
        a = prts[off[p]:off[p + 1]]
        key = np.lexsort([a["u"][:, 2], a["u"][:, 1], a["x"][:, 2], a["x"][:, 1], a["x"][:, 0], a["kind"]])
        out.append(a[key])
    return np.concatenate(out) if out else prts


@pytest.mark.parametrize("case", ["one_particle", "many_particles", "many_species"])
@pytest.mark.parametrize("fused", [0, 1])
def test_reference_integration_tests(case, fused):
    """test_boundary_injector.cxx:106-283 on the device, every step against the oracle"""
    import psc_b200 as pb
    gens = {"one_particle": lambda: [TestGenerator(1, 1)],
            "many_particles": lambda: [TestGenerator(-1, 1)],
            "many_species": lambda: [TestGenerator(-1, 1), TestGenerator(-1, 0)]}[case]
    og = ol.Grid(**injector_grid_kw())
    grid = make_gpu_grid(og)
    mprts, mflds = pb.Mparticles(grid), pb.MfieldsState(grid)
    checks = pb.Checks(grid, continuity_interval=1, gauss_interval=1)
    psc = pb.Psc(grid, mflds, mprts, sort_interval=0, checks=checks, fused=bool(fused))
    record = []
    n_steps = 2
    # the oracle run, step by step (so that the device can be compared after each one)
    o_flds, o_prts, o_off = og.zeros_fields(), np.zeros(0, dtype=ol.PRT_DTYPE), np.zeros(2, dtype=np.uint32)
    for i in range(len(gens())):
        psc.add_injector(_Replay(grid, record, i))
    generators = gens()
    psc.initialize()
    for step in range(n_steps):
        o_prts, o_off, errs, o_flds = run_oracle(og, generators, 1, prts=o_prts, off=o_off, flds=o_flds,
                                                 record=record)
        psc.step()
        got, got_off = mprts.get()
        assert np.array_equal(got_off, o_off)
        assert _sorted_records(got, got_off).tobytes() == _sorted_records(o_prts, o_off).tobytes()
        gf = mflds.download()
        assert np.abs(gf - o_flds).max() <= 2e-6 * max(np.abs(o_flds).max(), 1e-30)
        assert checks.continuity.last_max_err < CHECK_EPS and checks.gauss.last_max_err < CHECK_EPS
        assert abs(checks.continuity.last_max_err - errs[0][0]) < 1e-6
        assert abs(checks.gauss.last_max_err - errs[0][1]) < 1e-6
    n = mprts.size()
    if case == "one_particle":
        assert n == 1
    else:
        assert n > 1
    if case == "many_species":
        assert (got["kind"] == 0).any() and (got["kind"] == 1).any()
    grid.close()


def test_generated_on_the_host_like_the_reference():
    """the whole injector (draws included) through Psc::step with a Maxwellian generator on a
    multi-patch grid: only the patches on the lower y wall inject, and the checks hold"""
    import psc_b200 as pb
    kw = injector_grid_kw(gdims=(1, 16, 8), length=(1., 16., 8.), np_=(1, 2, 2), dt=0.5)
    og = ol.Grid(**kw)
    grid = make_gpu_grid(og)
    mprts, mflds = pb.Mparticles(grid), pb.MfieldsState(grid)
    checks = pb.Checks(grid, continuity_interval=1, gauss_interval=1)
    psc = pb.Psc(grid, mflds, mprts, sort_interval=2, checks=checks)
    gen = pb.ParticleGeneratorMaxwellian(1, grid.kinds[1], [0., 0.3, 0.], [0.01, 0.01, 0.01],
                                         rng=np.random.default_rng(3))
    inj = pb.BoundaryInjector(gen, grid, n_in_cell=lambda: 3)
    psc.add_injector(inj)
    psc.initialize()
    total = 0
    for _ in range(4):
        psc.step()
        total += inj.n_injected
        assert checks.continuity.last_max_err < CHECK_EPS and checks.gauss.last_max_err < CHECK_EPS
    prts, off = mprts.get()
    assert total > 0 and len(prts) == total  # nothing has reached the far wall yet
    n_by_patch = np.diff(off)
    lower = [p for p in range(grid.n_patches()) if grid.at_boundary_lo(p, 1)]
    assert n_by_patch[lower].sum() == total
    assert (prts["x"][:, 1] > 0).all() and (prts["kind"] == 1).all()
    grid.close()


def test_deposit_j_rejects_a_bad_list():
    import psc_b200 as pb
    og = ol.Grid(**injector_grid_kw())
    grid = make_gpu_grid(og)
    arr = (pb.JPath * 1)()
    arr[0].patch = 5
    with pytest.raises(pb.PscB200Error, match="names patch 5 of 1"):
        pb.check(grid.lib.psc_b200_deposit_j(grid.ctx, arr, 1))
    pb.check(grid.lib.psc_b200_deposit_j(grid.ctx, None, 0))  # an empty list is fine
    grid.close()
