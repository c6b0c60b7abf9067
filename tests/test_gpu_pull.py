"""Pull mode (option `pull`, off by default; push_lean_pull.cuh): the push of
step n + 1 completes the sort of step n -- only the particles that changed cell are moved
after a push, the stayers cross over to the other buffer inside the next push.  The store must
come out byte for byte as the scatter path leaves it.  With the fields frozen (push_fields = 0)
the trajectories do not depend on the order in which J was summed, so k-step runs compare
exactly; with the fields running the two paths agree to the rounding of J."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

KINDS = ((-1., 1.), (1., 25.))


def _run(pull, gdims, np3, ppc, steps, bc=None, push_fields=0, get_at=None, vth=(0.3, 0.06), cap=0):
    import psc_b200 as pb
    kw = {}
    if bc:
        kw = dict(bc_fld_lo=bc[0], bc_fld_hi=bc[0], bc_prt_lo=bc[1], bc_prt_hi=bc[1])
    grid = pb.Grid(gdims=gdims, length=tuple(float(g) for g in gdims), np=np3, dt=0.5, kinds=KINDS, nicell=ppc, **kw)
    grid.set_option("pull", pull)
    grid.set_option("pull_cap", cap)
    mprts, mflds = pb.Mparticles(grid), pb.MfieldsState(grid)
    mprts.setup_thermal(ppc, list(vth), seed=1)
    psc = pb.Psc(grid, mflds, mprts, sort_interval=1, fused=True)
    psc.initialize()
    mid = None
    for step in range(steps):
        prm = pb.StepParams(sort=1, marder_loop=0, marder_diffusion=0.9, push_fields=push_fields, checks=0)
        pb.check(grid.lib.psc_b200_step(grid.ctx, C.byref(prm)))
        if get_at == step:
            mid = mprts.get()  # forces the pending sort to be completed outside a push
    stats = {k: grid.get_stat(k) for k in ("pull_steps", "pull_materialized", "pull_overflows", "fused_fallbacks",
                                           "n_dropped")}
    prts, off = mprts.get()
    f = mflds.download()
    grid.close()
    return prts, off, f, stats, mid


CASES = {
    "xyz_2x1x1": dict(gdims=(64, 32, 32), np3=(2, 1, 1), ppc=8, steps=5),
    "xyz_2x2x2": dict(gdims=(64, 64, 64), np3=(2, 2, 2), ppc=6, steps=4),
    "xyz_1_patch_list_overflows": dict(gdims=(32, 32, 32), np3=(1, 1, 1), ppc=8, steps=4, cap=5000),
    "yz_2x2": dict(gdims=(1, 64, 64), np3=(1, 2, 2), ppc=16, steps=6),
    "xyz_open_z": dict(gdims=(32, 32, 64), np3=(1, 1, 2), ppc=6, steps=6, bc=([1, 1, 0], [1, 1, 3])),
    "xyz_cool": dict(gdims=(32, 32, 32), np3=(1, 1, 1), ppc=16, steps=5, vth=(0.05, 0.005)),
}


@pytest.mark.parametrize("name", list(CASES))
def test_pull_store_is_the_scatter_store(name):
    kw = CASES[name]
    a = _run(1, **kw)
    b = _run(0, **kw)
    assert b[3]["pull_steps"] == 0
    assert a[3]["fused_fallbacks"] == 0 and a[3]["n_dropped"] == b[3]["n_dropped"]
    if "overflows" in name:
        assert a[3]["pull_overflows"] > 0  # every step took the full scatter after all
    else:
        assert a[3]["pull_steps"] > 0 and a[3]["pull_overflows"] == 0
    assert np.array_equal(a[1], b[1])
    assert a[0].tobytes() == b[0].tobytes()
    scale = np.abs(b[2][:, :3]).max()
    assert np.abs(a[2][:, :3] - b[2][:, :3]).max() <= 2e-6 * scale  # J: summation order only


def test_pull_completed_outside_a_push():
    """anything that reads the store between two steps completes the pending sort (pull_materialize)"""
    kw = dict(gdims=(64, 32, 32), np3=(2, 1, 1), ppc=8, steps=5, get_at=2)
    a = _run(1, **kw)
    b = _run(0, **kw)
    assert a[3]["pull_materialized"] >= 1  # the read in the middle (the stats are taken before the last one)
    assert a[4][0].tobytes() == b[4][0].tobytes() and np.array_equal(a[4][1], b[4][1])
    assert a[0].tobytes() == b[0].tobytes()


def test_pull_with_running_fields_agrees_to_rounding():
    kw = dict(gdims=(64, 32, 32), np3=(2, 1, 1), ppc=8, steps=4, push_fields=1, vth=(0.1, 0.02))
    a = _run(1, **kw)
    b = _run(0, **kw)
    assert np.array_equal(a[1], b[1])
    assert np.abs(a[0]["x"] - b[0]["x"]).max() < 1e-4 and np.abs(a[0]["u"] - b[0]["u"]).max() < 1e-5
    assert np.abs(a[2] - b[2]).max() <= 2e-5 * np.abs(b[2]).max()
