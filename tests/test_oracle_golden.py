"""Pins the CPU oracle (and the reference-header build, when present) against
every golden vector the reference's own tests hold for the hot path."""
import ctypes as C

import numpy as np
import pytest

import golden_cases as gc
import oracle_lib as ol
from golden_cases import (GOLDEN, run_push_case, deposit_grid, rho_nc_norm,
                          push_fixture_grid, inject)

PUSH_IDS = [c["name"] for c in GOLDEN["push_cases"]]
DEP_IDS = [c["name"] for c in GOLDEN["deposit_cases"]]

BACKENDS = [("oracle", ol.push_mprts)]
if ol.ref_available():
    BACKENDS.append(("ref", ol.ref_push_mprts))


@pytest.mark.parametrize("backend", BACKENDS, ids=[b[0] for b in BACKENDS])
@pytest.mark.parametrize("dim", ["xyz", "yz"])
@pytest.mark.parametrize("case", GOLDEN["push_cases"], ids=PUSH_IDS)
def test_single_particle_push(case, dim, backend):
    # test_push_particles.cxx is typed over Config1vbecSplit for yz and xyz
    run_push_case(case, dim, backend[1])


@pytest.mark.parametrize("dim", ["yz"])
@pytest.mark.parametrize("case", [c for c in GOLDEN["push_cases"]], ids=PUSH_IDS)
def test_single_particle_push_var1(case, dim):
    """production yz deposit (Current1vbVar1, psc_config.hxx:47-72) must satisfy the
    same known answers (unpinned upstream: the tests only type over Split)."""
    import golden_cases as gc
    orig = gc.push_fixture_grid
    try:
        gc.push_fixture_grid = lambda d, deposit=ol.DEPOSIT_VAR1, np3=(1, 1, 1): orig(d, ol.DEPOSIT_VAR1, np3)
        run_push_case(case, dim, ol.push_mprts)
    finally:
        gc.push_fixture_grid = orig


def _calc_j(impl, grid, real, xm, xp, vxi):
    dt = np.float64 if real else np.float32
    f = np.zeros((9, grid.im[2], grid.im[1], grid.im[0]), dtype=dt)
    if impl == "oracle":
        fn = ol.lib().po_calc_j_d if real else ol.lib().po_calc_j_f
        a = [np.array(v, dtype=dt) for v in (xm, xp, vxi)]
        fn(grid.byref(), ol.ptr(f), ol.ptr(a[0]), ol.ptr(a[1]), ol.ptr(a[2]), 1.0)
    else:
        g = grid.g
        rc = ol.ref().psc_ref_calc_j(int(real), 1 if grid.is_yz else 0, g.deposit, g.gdims,
                                     g.length, g.dt, g.fnqs, ol.ptr(f), g.im, g.ib,
                                     ol.d3(*xm), ol.d3(*xp), ol.d3(*vxi), 1.0)
        assert rc == 0
    return f


def _div_j(J, yz):
    """CalcDivNc (test_current_deposition.cxx:14-44): backward differences, dx = 1"""
    jx, jy, jz = J[0], J[1], J[2]
    if yz:
        return (jy[1:, 1:, :] - jy[1:, :-1, :]) + (jz[1:, 1:, :] - jz[:-1, 1:, :])
    return ((jx[1:, 1:, 1:] - jx[1:, 1:, :-1]) + (jy[1:, 1:, 1:] - jy[1:, :-1, 1:])
            + (jz[1:, 1:, 1:] - jz[:-1, 1:, 1:]))


IMPLS = ["oracle"] + (["ref"] if ol.ref_available() else [])


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("deposit", [ol.DEPOSIT_SPLIT, ol.DEPOSIT_VAR1], ids=["split", "var1"])
@pytest.mark.parametrize("dim", ["yz", "xyz"])
@pytest.mark.parametrize("case", GOLDEN["deposit_cases"], ids=DEP_IDS)
def test_current_deposition_double(case, dim, deposit, impl):
    """test_current_deposition.cxx:127-152: exact J arrays (yz) and discrete
    continuity d_rho + div J = 0 to 2*DBL_EPSILON (yz and xyz)."""
    if deposit == ol.DEPOSIT_VAR1 and dim == "xyz":
        pytest.skip("Current1vbVar1 is yz only (inc_curr_1vb_var1.cxx:162-166)")
    eps = 2 * np.finfo(np.float64).eps
    grid = deposit_grid(dim, deposit)
    yz = dim == "yz"
    xm, xp, vxi = case["xm"], case["xp"], case["vxi"]
    f = _calc_j(impl, grid, True, xm, xp, vxi)
    # upstream allocates no ghosts (ibn = 0); our grids carry ibn = 2: take the interior
    ib, l3 = grid.ib, grid.ldims
    f = f[:, -ib[2]:-ib[2] + l3[2], -ib[1]:-ib[1] + l3[1], -ib[0]:-ib[0] + l3[0]]
    # continuity (:104-125).  In yz the x-displacement does not enter rho.
    ld = grid.ldims
    d_rho = rho_nc_norm(ld, xp, yz) - rho_nc_norm(ld, xm, yz)
    if yz:
        d_rho = d_rho[1:ld[2], 1:ld[1], 0:1]
    else:
        d_rho = d_rho[1:ld[2], 1:ld[1], 1:ld[0]]
    div = _div_j(f[0:3], yz)
    assert np.abs(d_rho + div).max() < eps
    if yz and "jyi_ref_zy" in case and not (case["split_only"] and deposit != ol.DEPOSIT_SPLIT and False):
        for m, key in ((0, "jxi_ref_zy"), (1, "jyi_ref_zy"), (2, "jzi_ref_zy")):
            ref = np.array(case[key])
            assert np.abs(f[m, :, :, 0] - ref).max() < eps, key


@pytest.mark.parametrize("dim", ["yz", "xyz"])
@pytest.mark.parametrize("case", GOLDEN["deposit_cases"], ids=DEP_IDS)
def test_current_deposition_float_oracle_vs_ref(case, dim):
    if not ol.ref_available():
        pytest.skip("no _ref")
    for deposit in (ol.DEPOSIT_SPLIT, ol.DEPOSIT_VAR1):
        if deposit == ol.DEPOSIT_VAR1 and dim == "xyz":
            continue
        grid = deposit_grid(dim, deposit)
        a = _calc_j("oracle", grid, False, case["xm"], case["xp"], case["vxi"])
        b = _calc_j("ref", grid, False, case["xm"], case["xp"], case["vxi"])
        assert a.tobytes() == b.tobytes()


def test_sort_known_answer():
    """test_collision_cuda.cxx:104-190: cell indices, stable permutation, offsets"""
    sc = GOLDEN["sort_case"]
    grid = ol.Grid(gdims=sc["gdims"], length=sc["length"], np_=sc["np"], dt=1.,
                   kinds=[(1., 1.)], nicell=200)
    prts, off = inject(grid, [(e["patch"], e["x"], (e["ux"], 0., 0.), 1., 0)
                              for e in sc["inject"]])
    n_cells = grid.n_cells
    patch_of = np.repeat(np.arange(grid.n_patches), np.diff(off))
    idx = [int(patch_of[i]) * n_cells + ol.cell_index(grid, prts["x"][i]) for i in range(len(prts))]
    assert idx == sc["idx_before"]
    ids_ux = prts["u"][:, 0].copy()
    rc, perm = ol.sort(grid, prts, off, want_perm=True)
    assert rc == 0
    idx2 = [int(patch_of[i]) * n_cells + ol.cell_index(grid, prts["x"][i]) for i in range(len(prts))]
    assert idx2 == sc["idx_after"]
    # identify particles by their ux tag
    got_id = [int(np.argmin(np.abs(ids_ux - prts["u"][i, 0]))) for i in range(len(prts))]
    assert got_id == sc["id_after"]
    glob_perm = [int(off[patch_of[i]] + perm[i]) for i in range(len(prts))]
    assert glob_perm == sc["id_after"]
    cnt = ol.count_by_cell(grid, prts, off)
    offs = np.concatenate([[0], np.cumsum(cnt)])
    assert offs[1] == 2 and all(offs[2:10] == 6) and all(offs[10:83] == 9)
    assert offs[83] == 11 and all(offs[84:92] == 13) and all(offs[92:257] == 15)


@pytest.mark.parametrize("dim", ["xyz", "yz"])
@pytest.mark.parametrize("case", gc.MOMENT_CASES, ids=[c["name"] for c in gc.MOMENT_CASES])
def test_oracle_moment_known_answers(case, dim):
    """the oracle's moment family on the reference's known-answer cases
    (src/libpsc/tests/test_moments.cxx:149-402, eps 1e-6)"""
    og, prts, off = gc.moment_case_grid(case, dim)
    got = gc.moment_interior_comp0(og, ol.moment_1st(og, prts, off, case["which"]))
    exp = gc.moment_case_expected(case, dim, og)
    assert np.abs(got - exp).max() < 1e-6


@pytest.mark.parametrize("dim", ["xyz", "yz"])
@pytest.mark.parametrize("name", list(gc.FIELD_CASES))
def test_oracle_field_known_answers(name, dim):
    """the oracle's Yee / Marder-correct / div / ghost operators on the reference's own
    known-answer tests (src/libpsc/tests/test_push_fields.cxx:26-260, test_bnd.cxx:104-303)"""
    got, exp, tol = gc.FIELD_CASES[name](dim, gc.OracleFieldOps)
    if tol == 0.:
        assert np.array_equal(got, exp)
    else:
        assert np.abs(got - exp).max() < tol
