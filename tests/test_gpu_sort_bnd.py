"""Sort and particle boundary exchange on the device vs the CPU oracle: bit-exact
permutations, per-cell counts and migration (BASELINE.json north_star)."""
import numpy as np
import pytest

import oracle_lib as ol
from b200_helpers import gpu_state
from gen import thermal_plasma
from golden_cases import GOLDEN, inject

pytestmark = pytest.mark.gpu

KINDS = ((-1., 1.), (1., 100.))


def _grid(dim, np3=None, **kw):
    if dim == "yz":
        return ol.Grid(gdims=(1, 24, 16), length=(1., 30., 10.), np_=np3 or (1, 3, 2), dt=0.2,
                       kinds=KINDS, nicell=10, **kw)
    return ol.Grid(gdims=(8, 12, 16), length=(5., 9., 20.), np_=np3 or (2, 1, 2), dt=0.2,
                   kinds=KINDS, nicell=10, **kw)


@pytest.mark.parametrize("ppc", [1, 7, 40])
@pytest.mark.parametrize("dim", ["xyz", "yz"])
def test_sort_matches_oracle(dim, ppc):
    import psc_b200 as pb
    og = _grid(dim)
    prts, off = thermal_plasma(og, ppc=ppc, seed=3, vth=(0.1, 0.01), shuffle=True)
    grid, mprts, _ = gpu_state(og, None, prts, off)
    pb.Sort()(mprts)
    got, got_off = mprts.get()
    ref = prts.copy()
    rc, perm = ol.sort(og, ref, off, want_perm=True)
    assert rc == 0
    assert np.array_equal(got_off, off)
    # records are distinct, so equal arrays <=> equal permutations
    assert got.tobytes() == ref.tobytes()
    assert np.array_equal(ol.count_by_cell(og, got, got_off), ol.count_by_cell(og, ref, off))
    # idempotent
    pb.Sort()(mprts)
    again, _ = mprts.get()
    assert again.tobytes() == got.tobytes()
    assert grid.get_stat("sorted") == 1
    grid.close()


def test_sort_empty_and_ragged():
    import psc_b200 as pb
    og = _grid("xyz")
    prts, off = thermal_plasma(og, ppc=3, seed=4, vth=(0.1, 0.01))
    # keep particles of patches 1 and 3 only, and only a ragged subset
    keep = np.zeros(len(prts), bool)
    keep[off[1]:off[1] + 17] = True
    keep[off[3] + 5:off[4]] = True
    n_by = [int(keep[off[p]:off[p + 1]].sum()) for p in range(og.n_patches)]
    prts = prts[keep]
    off = ol.off_from_counts(n_by)
    grid, mprts, _ = gpu_state(og, None, prts, off)
    pb.Sort()(mprts)
    got, got_off = mprts.get()
    ref = prts.copy()
    ol.sort(og, ref, off)
    assert np.array_equal(got_off, off) and got.tobytes() == ref.tobytes()
    # no particles at all
    grid2, mprts2, _ = gpu_state(og, None, prts[:0], ol.off_from_counts([0] * og.n_patches))
    pb.Sort()(mprts2)
    assert mprts2.size() == 0
    grid.close()
    grid2.close()


def test_sort_rejects_particle_outside_patch():
    import psc_b200 as pb
    og = _grid("xyz")
    prts, off = thermal_plasma(og, ppc=1, seed=4, vth=(0.1, 0.01))
    prts["x"][5, 1] = -0.5
    grid, mprts, _ = gpu_state(og, None, prts, off)
    with pytest.raises(pb.PscB200Error):
        pb.Sort()(mprts)
    grid.close()


def test_sort_known_answer():
    """test_collision_cuda.cxx:104-190 (the reference's only sort vector)"""
    import psc_b200 as pb
    sc = GOLDEN["sort_case"]
    og = ol.Grid(gdims=sc["gdims"], length=sc["length"], np_=sc["np"], dt=1., kinds=[(1., 1.)],
                 nicell=200)
    prts, off = inject(og, [(e["patch"], e["x"], (e["ux"], 0., 0.), 1., 0) for e in sc["inject"]])
    ids_ux = prts["u"][:, 0].copy()
    grid, mprts, _ = gpu_state(og, None, prts, off)
    pb.Sort()(mprts)
    got, _ = mprts.get()
    got_id = [int(np.argmin(np.abs(ids_ux - got["u"][i, 0]))) for i in range(len(got))]
    assert got_id == sc["id_after"]
    grid.close()


BCS = {
    "periodic": {},
    "reflecting_y": dict(bc_fld_lo=[1, 2, 1], bc_fld_hi=[1, 2, 1], bc_prt_lo=[1, 0, 1], bc_prt_hi=[1, 0, 1]),
    "absorbing_z": dict(bc_fld_lo=[1, 1, 2], bc_fld_hi=[1, 1, 2], bc_prt_lo=[1, 1, 2], bc_prt_hi=[1, 1, 2]),
}


@pytest.mark.parametrize("np3", [None, "single"])
@pytest.mark.parametrize("bc", list(BCS))
@pytest.mark.parametrize("dim", ["xyz", "yz"])
def test_bnd_particles_matches_oracle(dim, bc, np3):
    import psc_b200 as pb
    og = _grid(dim, np3=(1, 1, 1) if np3 else None, **BCS[bc])
    prts, off = thermal_plasma(og, ppc=9, seed=5, vth=(0.3, 0.03))
    # move a good fraction out of their patches (less than one patch size)
    rng = np.random.default_rng(6)
    kick = (rng.random(prts["x"].shape) - 0.5) * np.array(og.dx) * 3.
    for d in range(3):
        if not og.g.invar[d]:
            prts["x"][:, d] += kick[:, d].astype(np.float32)
    ref, ref_off, n_drop = ol.bnd_particles(og, prts, off)
    grid, mprts, _ = gpu_state(og, None, prts, off)
    pb.BndParticles(grid)(mprts)
    got, got_off = mprts.get()
    assert np.array_equal(got_off, ref_off)
    assert got.tobytes() == ref.tobytes()
    assert grid.get_stat("n_dropped") == n_drop
    if bc == "absorbing_z":
        assert n_drop > 0
    # every particle is inside its patch again: sorting must succeed and agree
    pb.Sort()(mprts)
    got2, _ = mprts.get()
    ol.sort(og, ref, ref_off)
    assert got2.tobytes() == ref.tobytes()
    grid.close()


@pytest.mark.parametrize("dim", ["xyz", "yz"])
def test_inject_appends_per_patch(dim):
    og = _grid(dim)
    a, off_a = thermal_plasma(og, ppc=2, seed=1, vth=(0.1, 0.01))
    b, off_b = thermal_plasma(og, ppc=1, seed=2, vth=(0.1, 0.01))
    grid, mprts, _ = gpu_state(og, None, a, off_a)
    mprts.inject(b, np.diff(off_b))
    got, got_off = mprts.get()
    exp = np.concatenate([np.concatenate([a[off_a[p]:off_a[p + 1]], b[off_b[p]:off_b[p + 1]]])
                          for p in range(og.n_patches)])
    assert np.array_equal(np.diff(got_off), np.diff(off_a) + np.diff(off_b))
    assert got.tobytes() == exp.tobytes()
    assert np.array_equal(mprts.sizeByPatch(), np.diff(got_off))
    grid.close()


def test_fused_step_equals_separate_operators():
    """psc_b200_step (boundary exchange fused with the next sort) leaves exactly the
    store that bnd_particles followed by sort produces"""
    import psc_b200 as pb
    from gen import random_fields
    og = _grid("xyz")
    flds = random_fields(og, seed=2)
    prts, off = thermal_plasma(og, ppc=10, seed=9, vth=(0.4, 0.04))
    res = []
    for fused in (0, 1):
        grid, mprts, mflds = gpu_state(og, flds, prts, off, dict(fused_sort=fused, gapped=0))
        prm = pb.StepParams(sort=1, marder_loop=0, marder_diffusion=0., push_fields=0, checks=0)
        import ctypes as C
        for _ in range(3):
            pb.check(grid.lib.psc_b200_step(grid.ctx, C.byref(prm)))
        if not fused:
            pb.Sort()(mprts)
        res.append(mprts.get())
        if fused:
            assert grid.get_stat("fused_steps") >= 2
        grid.close()
    assert np.array_equal(res[0][1], res[1][1])
    assert res[0][0].tobytes() == res[1][0].tobytes()


@pytest.mark.parametrize("dim", ["xyz", "yz"])
def test_keep_sorted_deck_cadence(dim):
    """A deck that sorts every 10th step (psc_bubble_yz.cxx / psc_harris_yz.cxx
    sort_interval = 10) through the separate operators: with keep_sorted (default) every step
    takes the tiled push + fused exchange/sort, without it the unordered-store kernels run
    between sorts.  Same particles (as multisets per patch: only the order may differ), same
    counts, same J."""
    import psc_b200 as pb
    from gen import random_fields
    og = _grid(dim)
    flds = random_fields(og, seed=3)
    prts, off = thermal_plasma(og, ppc=8, seed=4, vth=(0.3, 0.03))
    res = []
    for keep in (0, 1):
        grid, mprts, mflds = gpu_state(og, flds, prts, off, dict(keep_sorted=keep))
        sort_, pushp, bndp = pb.Sort(), pb.PushParticles(), pb.BndParticles(grid)
        for step in range(1, 8):
            if step % 5 == 0:
                sort_(mprts)
            pushp.push_mprts(mprts, mflds)
            bndp(mprts)
        n_fused = grid.get_stat("fused_steps")
        assert n_fused == (7 if keep else 0), n_fused
        res.append((mprts.get(), mflds.download(0, 3)))
        grid.close()
    (p0, o0), j0 = res[0]
    (p1, o1), j1 = res[1]
    assert np.array_equal(o0, o1)
    for p in range(og.n_patches):
        a, b = p0[o0[p]:o0[p + 1]], p1[o1[p]:o1[p + 1]]
        ka = np.lexsort((a["u"][:, 2], a["u"][:, 1], a["u"][:, 0], a["x"][:, 1], a["x"][:, 2], a["kind"]))
        kb = np.lexsort((b["u"][:, 2], b["u"][:, 1], b["u"][:, 0], b["x"][:, 1], b["x"][:, 2], b["kind"]))
        # fields are fixed here (no field push), so every particle's update is independent of
        # the others: bit-identical records
        assert a[ka].tobytes() == b[kb].tobytes()
    assert np.abs(j0 - j1).max() <= 1e-5 * np.abs(j0).max()


def test_fused_sort_falls_back_for_a_huge_cell():
    """the destination-count planes are 16 bit: a cell with more than 65535 particles breaks
    the fused pass's precondition and the step takes the general exchange + sort -- same
    result (bit-exact against the oracle), counted as a fallback"""
    import ctypes as C
    import psc_b200 as pb
    from gen import random_fields
    og = ol.Grid(gdims=(8, 8, 8), length=(8., 8., 8.), np_=(1, 1, 1), dt=0.3, kinds=KINDS, nicell=4)
    flds = random_fields(og, seed=1)
    prts, off = thermal_plasma(og, ppc=2, seed=2, vth=(0.2, 0.02))
    rng = np.random.default_rng(3)
    n_big = 70000
    big = np.zeros(n_big, dtype=prts.dtype)
    big["x"] = (np.array([3., 4., 5.]) + 0.05 + 0.9 * rng.random((n_big, 3))).astype(np.float32)
    big["u"] = (0.2 * rng.standard_normal((n_big, 3))).astype(np.float32)
    big["kind"] = 0
    big["qni_wni"] = np.float32(-1.)
    prts = np.concatenate([prts, big])
    off = ol.off_from_counts([len(prts)])
    grid, mprts, mflds = gpu_state(og, flds, prts, off)
    prm = pb.StepParams(sort=1, marder_loop=0, marder_diffusion=0., push_fields=0, checks=0)
    rp, ro = prts.copy(), off.copy()
    L, G = ol.lib(), og.byref()
    for _ in range(2):
        pb.check(grid.lib.psc_b200_step(grid.ctx, C.byref(prm)))
        assert L.po_sort(G, ol.ptr(rp), ol.ptr(ro), None) == 0
        L.po_push_mprts(G, ol.ptr(flds.copy()), ol.ptr(rp), ol.ptr(ro))
        rp, ro, _ = ol.bnd_particles(og, rp, ro)
    assert L.po_sort(G, ol.ptr(rp), ol.ptr(ro), None) == 0
    assert grid.get_stat("fused_fallbacks") >= 1
    got, got_off = mprts.get()
    assert np.array_equal(got_off, ro)
    assert got.tobytes() == rp.tobytes()
    grid.close()
