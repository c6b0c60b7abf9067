"""OutputFieldsItem's hand-off (src/include/output_fields.hxx:150-236) on the device: the three
primitives bit for bit against numpy, and OutputFields / OutputMoments over a run against the
same bookkeeping done on the oracle's states."""
import numpy as np
import pytest

import oracle_lib as ol
from b200_helpers import gpu_state, make_gpu_grid
from gen import thermal_plasma

pytestmark = pytest.mark.gpu

KINDS = ((-1., 1.), (1., 25.))
GRIDS = {
    "xyz": dict(gdims=(8, 8, 16), length=(8., 8., 16.), np_=(1, 2, 2)),    # im^3 % 4 == 0: 128-bit path
    "yz_odd": dict(gdims=(1, 10, 21), length=(1., 10., 21.), np_=(1, 2, 3)),  # 9 * 11 points: scalar path
}


def _interior(og, a):
    b = og.ibn
    sl = [slice(b[d], a.shape[4 - d] - b[d]) for d in range(3)]
    return a[:, :, sl[2], sl[1], sl[0]]


@pytest.mark.parametrize("name", list(GRIDS))
def test_add_scale_interior_bit_exact(name):
    import psc_b200 as pb
    og = ol.Grid(dt=0.3, kinds=KINDS, nicell=4, **GRIDS[name])
    grid = make_gpu_grid(og)
    rng = np.random.default_rng(5)
    a, b = pb.Mfields(grid, 5), pb.Mfields(grid, 3)
    ha = rng.normal(size=a.shape()).astype(np.float32)
    hb = rng.normal(size=b.shape()).astype(np.float32)
    a.upload(ha)
    b.upload(hb)
    a.add(b, mb=1, other_mb=0, n_comps=3)   # tfd = tfd + pfd on a component window
    ha[:, 1:4] = ha[:, 1:4] + hb
    assert a.download().tobytes() == ha.tobytes()
    a.add(b, mb=4, other_mb=2, n_comps=1)
    ha[:, 4] = ha[:, 4] + hb[:, 2]
    a.scale(1. / 7, mb=1, me=5)             # (1. / naccum) * tfd: double scalar, float data
    ha[:, 1:5] = (np.float64(1. / 7) * ha[:, 1:5].astype(np.float64)).astype(np.float32)
    assert a.download().tobytes() == ha.tobytes()
    got = a.download_interior(1, 4)
    assert got.shape == (og.n_patches, 3) + tuple(og.ldims[::-1])
    assert got.tobytes() == np.ascontiguousarray(_interior(og, ha)[:, 1:4]).tobytes()
    assert b.download().tobytes() == hb.tobytes()  # the source is untouched
    with pytest.raises(pb.PscB200Error, match="overlap"):
        a.add(a, mb=1, other_mb=2, n_comps=2)
    with pytest.raises(pb.PscB200Error, match="out of bounds"):
        a.add(b, mb=3, other_mb=0, n_comps=3)
    grid.close()


def test_output_fields_over_a_run():
    """pfield every 2 steps, tfield every 6 averaging the last 4 steps sampled every 2: the
    writers receive exactly the reference's schedule, names and the oracle's numbers"""
    import psc_b200 as pb
    og = ol.Grid(gdims=(1, 16, 16), length=(1., 16., 16.), np_=(1, 2, 2), dt=0.4, kinds=KINDS, nicell=6)
    prts, off = thermal_plasma(og, ppc=6, seed=3, vth=(0.2, 0.02))
    f = og.zeros_fields()
    f[:, ol.HX] = 0.1
    g = og.g
    grid = pb.Grid(gdims=tuple(g.gdims), length=tuple(g.length), np=tuple(g.np), dt=g.dt,
                   kinds=((-1., 1., "e"), (1., 25., "i")), fnqs=g.fnqs)
    mprts, mflds = pb.Mparticles(grid), pb.MfieldsState(grid)
    mprts.set(prts, np.diff(off))
    mflds.upload(f)
    psc = pb.Psc(grid, mflds, mprts, sort_interval=2, fused=True)
    pf = dict(out_interval=2)
    tf = dict(out_interval=6, average_length=4, sample_interval=2)
    outf = pb.OutputFields(pb.OutputFieldItemParams(**pf), pb.OutputFieldItemParams(**tf))
    outm = pb.OutputMoments(grid, pb.OutputFieldItemParams(**pf), pb.OutputFieldItemParams(**tf))
    psc.add_diagnostic(outf)
    psc.add_diagnostic(outm)
    n_steps = 6
    psc.integrate(n_steps)
    # the same bookkeeping on the oracle
    ol.fill_ghosts(og, f, 0, 9)
    rp, ro = prts.copy(), off.copy()
    pfd, acc_f, acc_m, n_acc = {0: _interior(og, f).copy()}, None, None, 0
    for step in range(1, n_steps + 1):
        rp, ro = ol.step(og, f, rp, ro, sort_now=(step % 2 == 0))
        if step % 2 == 0:
            pfd[step] = _interior(og, f).copy()
        if step in (4, 6):  # next_out - t in {2, 0}: < 4 and even
            jeh, mom = _interior(og, f).copy(), _interior(og, ol.moment_1st(og, rp, ro, ol.MOM_ALL)).copy()
            acc_f = jeh if acc_f is None else acc_f + jeh
            acc_m = mom if acc_m is None else acc_m + mom
            n_acc += 1
    mean_f = (np.float64(1. / n_acc) * acc_f.astype(np.float64)).astype(np.float32)
    mean_m = (np.float64(1. / n_acc) * acc_m.astype(np.float64)).astype(np.float32)

    assert outf.io_pfd.pfx == "pfd" and outf.io_tfd.pfx == "tfd"
    assert outm.io_pfd.pfx == "pfd_moments" and outm.io_tfd.pfx == "tfd_moments"
    assert [s["timestep"] for s in outf.io_pfd.steps] == [0, 2, 4, 6]
    assert [s["timestep"] for s in outf.io_tfd.steps] == [6] and [s["timestep"] for s in outm.io_tfd.steps] == [6]
    for s in outf.io_pfd.steps:
        assert s["name"] == "jeh" and s["comp_names"][3] == "ex_ec"
        ref = pfd[s["timestep"]]
        assert np.abs(s["data"] - ref).max() <= 2e-5 * max(np.abs(ref).max(), 1e-30)
    t = outf.io_tfd.steps[0]
    assert np.abs(t["data"] - mean_f).max() <= 2e-5 * np.abs(mean_f).max()
    m = outm.io_tfd.steps[0]
    assert m["name"] == "all_1st_cc" and m["comp_names"][:2] == ["rho_e", "jx_e"] and len(m["comp_names"]) == 26
    assert np.abs(m["data"] - mean_m).max() <= 1e-4 * np.abs(mean_m).max()
    # the running sum was cleared after the write (output_fields.hxx:227-229)
    assert outf.naccum == 0 and not outf.tfd.download().any()
    grid.close()
