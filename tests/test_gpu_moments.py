"""Device-side 1st-order moments: Moment_n_1st, Moment_v_1st,
Moment_p_1st, Moment_T_1st, Moments_1st (cell-centred) and Moment_rho_1st_nc against the
oracle's restatement of psc/moment.hxx + psc/deposit.hxx + fields_item.hxx (ghost add and
reflecting-wall folds included), and on the reference's own known-answer cases
(src/libpsc/tests/test_moments.cxx:149-402).  The deposit uses atomics, so only the
summation order differs from the CPU: compared at 2e-6 of the largest value."""
import numpy as np
import pytest

import oracle_lib as ol
from b200_helpers import gpu_state
from gen import thermal_plasma
from golden_cases import MOMENT_CASES, moment_case_grid, moment_case_expected, moment_interior_comp0

pytestmark = pytest.mark.gpu

KINDS = ((-1., 1.), (1., 100.), (-1., 1.))
WALL_Z = dict(bc_fld_lo=[1, 1, 2], bc_fld_hi=[1, 1, 2], bc_prt_lo=[1, 1, 0], bc_prt_hi=[1, 1, 0])
WALL_YZ = dict(bc_fld_lo=[1, 2, 2], bc_fld_hi=[1, 2, 2], bc_prt_lo=[1, 0, 0], bc_prt_hi=[1, 0, 0])
GRIDS = {
    "xyz": dict(gdims=(16, 8, 16), length=(16., 8., 16.), np_=(2, 1, 2)),
    "xyz_wall_z": dict(gdims=(8, 8, 16), length=(8., 8., 16.), np_=(1, 1, 2), **WALL_Z),
    "yz": dict(gdims=(1, 16, 32), length=(1., 20., 30.), np_=(1, 2, 2)),
    "yz_wall_yz": dict(gdims=(1, 16, 16), length=(1., 16., 16.), np_=(1, 2, 1), **WALL_YZ),
}
MOMENTS = {"n": ol.MOM_N, "v": ol.MOM_V, "p": ol.MOM_P, "T": ol.MOM_T, "all": ol.MOM_ALL, "rho_nc": ol.MOM_RHO_NC}


@pytest.mark.parametrize("store", ["unordered", "cell_ordered"])
@pytest.mark.parametrize("mom", list(MOMENTS))
@pytest.mark.parametrize("name", list(GRIDS))
def test_moment_matches_oracle(name, mom, store):
    """unordered store: one thread per particle, global atomics per value and corner; cell-ordered
    store: one warp per cell, the sums of a cell in registers, one add per cell, target and component"""
    import psc_b200 as pb
    og = ol.Grid(dt=0.4, kinds=KINDS, nicell=6, **GRIDS[name])
    prts, off = thermal_plasma(og, ppc=5, seed=21, vth=(0.4, 0.03, 0.2))
    prts["qni_wni"] *= (0.5 + np.random.default_rng(3).random(len(prts))).astype(np.float32)  # weights != 1
    ref = ol.moment_1st(og, prts, off, MOMENTS[mom])
    grid, mprts, _ = gpu_state(og, None, prts, off)
    if store == "cell_ordered":
        pb.Sort()(mprts)
    assert grid.get_stat("sorted") == (store == "cell_ordered")
    item = pb.Moment(grid, MOMENTS[mom])
    assert item.n_comps() == ref.shape[1]
    got = item(mprts).download()
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() <= 2e-6 * np.abs(ref).max()
    if mom == "rho_nc":
        # the dedicated entry point Marder / checks use gives the same array
        rho = pb.Mfields(grid, 1)
        pb.check(grid.lib.psc_b200_moment_rho_1st_nc(grid.ctx, rho.id))
        assert np.abs(rho.download() - ref).max() <= 2e-6 * np.abs(ref).max()
    grid.close()


@pytest.mark.parametrize("store", ["unordered", "cell_ordered"])
@pytest.mark.parametrize("dim", ["xyz", "yz"])
@pytest.mark.parametrize("case", MOMENT_CASES, ids=[c["name"] for c in MOMENT_CASES])
def test_moment_known_answers(case, dim, store):
    """test_moments.cxx: one particle of weight .4, nicell 200, dx 10"""
    import psc_b200 as pb
    og, prts, off = moment_case_grid(case, dim)
    grid, mprts, _ = gpu_state(og, None, prts, off)
    if store == "cell_ordered":
        pb.Sort()(mprts)
    got = moment_interior_comp0(og, pb.Moment(grid, case["which"])(mprts).download())
    exp = moment_case_expected(case, dim, og)
    assert np.abs(got - exp).max() < 1e-6
    grid.close()
