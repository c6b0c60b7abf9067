"""CPU-only check of the product's per-particle arithmetic (psc_b200/csrc/
pic_math.cuh compiled for the host by tests/hostcheck/) against the oracle:
the explicit-stack trajectory split, the Var1 piece walker, the gather/Boris/move
and the boundary classification must be bit-identical to the reference restatement.
The CUDA kernels run exactly this code (test_gpu_*.py repeat the comparison on
the device through the C ABI)."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as ol
from b200_helpers import desc_from_grid, hostcheck
from gen import random_fields, thermal_plasma

CASES = [
    ("xyz_split", dict(gdims=(8, 8, 8), length=(8., 8., 8.), deposit=ol.DEPOSIT_SPLIT)),
    ("xyz_split_aniso", dict(gdims=(8, 4, 12), length=(10., 3., 7.), deposit=ol.DEPOSIT_SPLIT)),
    ("yz_var1", dict(gdims=(1, 16, 16), length=(1., 20., 12.), deposit=ol.DEPOSIT_VAR1)),
    ("yz_split", dict(gdims=(1, 16, 16), length=(1., 20., 12.), deposit=ol.DEPOSIT_SPLIT)),
]


@pytest.mark.parametrize("name,kw", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("vth", [0.05, 0.6])
def test_device_math_on_host_bit_exact(name, kw, vth):
    kinds = ((-1., 1.), (1., 100.))
    dt = 0.4 * min(l / g for l, g in zip(kw["length"], kw["gdims"]) if g > 1)
    grid = ol.Grid(dt=dt, kinds=kinds, nicell=50, **kw)
    flds = random_fields(grid, seed=3)
    prts, off = thermal_plasma(grid, ppc=8, seed=5, vth=(vth, vth / 10))
    f1, p1 = flds.copy(), prts.copy()
    f2, p2 = flds.copy(), prts.copy()
    ol.push_mprts(grid, f1, p1, off)
    d = desc_from_grid(grid)
    rc = hostcheck().hc_push_mprts(C.byref(d), ol.ptr(f2), ol.ptr(p2), ol.ptr(off))
    assert rc == 0
    assert p1.tobytes() == p2.tobytes()
    # sequential host accumulation in the same order => J identical too
    assert f1.tobytes() == f2.tobytes()


@pytest.mark.parametrize("bc", ["periodic", "reflecting", "absorbing"])
@pytest.mark.parametrize("dim", ["xyz", "yz"])
def test_bnd_classify_matches_process_patch(dim, bc):
    gd = (1, 8, 8) if dim == "yz" else (8, 8, 8)
    np3 = (1, 2, 2) if dim == "yz" else (2, 2, 1)
    kw = {}
    if bc != "periodic":
        prt_bc = ol.BND_PRT_REFLECTING if bc == "reflecting" else ol.BND_PRT_ABSORBING
        kw = dict(bc_fld_lo=[ol.BND_FLD_PERIODIC, ol.BND_FLD_CONDUCTING_WALL, ol.BND_FLD_PERIODIC],
                  bc_fld_hi=[ol.BND_FLD_PERIODIC, ol.BND_FLD_CONDUCTING_WALL, ol.BND_FLD_PERIODIC],
                  bc_prt_lo=[ol.BND_PRT_PERIODIC, prt_bc, ol.BND_PRT_PERIODIC],
                  bc_prt_hi=[ol.BND_PRT_PERIODIC, prt_bc, ol.BND_PRT_PERIODIC])
    grid = ol.Grid(gdims=gd, length=(3., 5., 7.), np_=np3, dt=0.2, kinds=((-1., 1.),), nicell=10, **kw)
    prts, off = thermal_plasma(grid, ppc=6, seed=11, vth=(0.5,))
    # scatter positions so that a good fraction lies just outside the patch
    rng = np.random.default_rng(2)
    kick = (rng.random(prts["x"].shape) - 0.5) * np.array(grid.dx) * 1.6
    for d in range(3):
        if not grid.g.invar[d]:
            prts["x"][:, d] += kick[:, d].astype(np.float32)
    # a few exact edge cases (bnd_particles_impl.hxx:139-143,199-202)
    prts["x"][0, 1] = np.float32(-1e-8)
    prts["x"][1, 2] = np.float32(-5e-7)
    prts["x"][2, 1] = np.float32(grid.ldims[1] * grid.dx[1])
    want, want_off, n_dropped = ol.bnd_particles(grid, prts, off)

    p2 = prts.copy()
    dirs = np.zeros((len(prts), 3), dtype=np.int32)
    flag = np.zeros(len(prts), dtype=np.int32)
    d = desc_from_grid(grid)
    rc = hostcheck().hc_bnd_classify(C.byref(d), ol.ptr(p2), ol.ptr(off), ol.ptr(dirs), ol.ptr(flag))
    assert rc == 0
    assert int((flag == 2).sum()) == n_dropped
    # rebuild the exchange result from the per-particle classification and compare
    # with the oracle's [stayers | arrivals in direction order] layout
    npch = grid.n_patches
    patch_of = np.repeat(np.arange(npch), np.diff(off))
    out = []
    for p in range(npch):
        mine = np.where((patch_of == p) & (flag != 2) & (dirs == 0).all(axis=1))[0]
        out.append(p2[mine])
        for dz in (-1, 0, 1):
            for dy in (-1, 0, 1):
                for dx in (-1, 0, 1):
                    if (dx, dy, dz) == (0, 0, 0):
                        continue
                    nei = ol.lib().po_neighbor_patch(grid.byref(), p, ol.i3(dx, dy, dz))
                    if nei < 0:
                        continue
                    sel = np.where((patch_of == nei) & (flag != 2) & (dirs[:, 0] == -dx)
                                   & (dirs[:, 1] == -dy) & (dirs[:, 2] == -dz))[0]
                    out.append(p2[sel])
    got = np.concatenate(out)
    assert got.tobytes() == want.tobytes()


@pytest.mark.parametrize("n_ranks,n_patches", [(1, 5), (2, 2), (3, 17), (8, 64), (8, 9), (5, 1536)])
def test_best_mapping_matches_reference_bisection(n_ranks, n_patches):
    """psc_b200_best_mapping (pure host function of the C ABI) vs the oracle's restatement
    of best_mapping_recursive (psc_balance_impl.hxx:99-160); the reference's own test
    (test_balance.cxx:234-253) only prints the mapping, so this is pinned by restatement"""
    import psc_b200
    L = psc_b200.load()
    rng = np.random.default_rng(n_ranks * 1000 + n_patches)
    for trial in range(20):
        loads = rng.uniform(0.1, 100., size=n_patches) ** (1 + trial % 3)
        cap = np.ones(n_ranks)
        got = np.zeros(n_ranks, dtype=np.int32)
        ref = np.zeros(n_ranks, dtype=np.int32)
        assert L.psc_b200_best_mapping(n_ranks, ol.ptr(cap), n_patches, ol.ptr(loads), ol.ptr(got)) == 0
        ol.lib().po_best_mapping(n_ranks, ol.ptr(cap), n_patches, ol.ptr(loads), ol.ptr(ref))
        assert got.tolist() == ref.tolist()
        assert got.sum() == n_patches and got.min() >= 1


def test_best_mapping_known_answer():
    """equal loads split evenly; one heavy patch gets a rank to itself"""
    import psc_b200
    L = psc_b200.load()
    out = np.zeros(4, dtype=np.int32)
    cap = np.ones(4)
    loads = np.ones(8)
    assert L.psc_b200_best_mapping(4, ol.ptr(cap), 8, ol.ptr(loads), ol.ptr(out)) == 0
    assert out.tolist() == [2, 2, 2, 2]
    loads = np.array([1., 1., 1., 100., 1., 1., 1., 1.])
    assert L.psc_b200_best_mapping(4, ol.ptr(cap), 8, ol.ptr(loads), ol.ptr(out)) == 0
    assert out.sum() == 8 and out.min() >= 1
    # the heavy patch (index 3) sits alone or nearly alone on its rank
    start = np.concatenate([[0], np.cumsum(out)])
    r = np.searchsorted(start, 3, side="right") - 1
    assert out[r] <= 2


def test_grid_setup_accepts_open_and_rejects_absorbing_field_bc():
    """BND_FLD_OPEN is implemented (psc_bnd_fields_impl.hxx:210-300,535-640); BND_FLD_ABSORBING is not --
    the reference asserts on it -- and must be refused instead of running with untouched ghost cells"""
    import ctypes as C
    import psc_b200
    from b200_helpers import desc_from_grid
    import oracle_lib as ol
    lib = psc_b200.load()
    for bc, ok_expected in ((0, True), (3, False)):
        og = ol.Grid(gdims=(8, 8, 8), length=(8., 8., 8.), np_=(1, 1, 1), dt=0.1, kinds=((-1., 1.),), nicell=1,
                     bc_fld_lo=[1, 1, bc], bc_fld_hi=[1, 1, bc], bc_prt_lo=[1, 1, 2], bc_prt_hi=[1, 1, 2])
        d = desc_from_grid(og)
        ctx = C.c_void_p()
        rc = lib.psc_b200_create(C.byref(d), C.byref(ctx))
        msg = lib.psc_b200_last_error().decode()
        if rc == 0:
            lib.psc_b200_destroy(ctx)
        if ok_expected:
            # (no GPU here: creation fails later, on the missing device, not on the grid)
            assert rc == 0 or "ABSORBING" not in msg
        else:
            assert rc != 0 and "ABSORBING" in msg
