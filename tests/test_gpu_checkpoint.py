"""Checkpoint / restart (psc_b200_checkpoint_write / _read; write_checkpoint / read_checkpoint,
src/include/checkpoint.hxx:14-82): what is read back is byte for byte what was written, the
file carries PSC's variable decomposition (size_by_patch + one array per particle component),
and a restarted run continues exactly like the uninterrupted one."""
import ctypes as C
import os
import struct

import numpy as np
import pytest

import oracle_lib as ol
from b200_helpers import gpu_state
from gen import random_fields, thermal_plasma

pytestmark = pytest.mark.gpu
KINDS = ((-1., 1.), (1., 100.))


@pytest.mark.parametrize("case", [dict(gdims=(16, 16, 16), length=(16., 16., 16.), np_=(2, 1, 2)),
                                  dict(gdims=(1, 32, 32), length=(1., 40., 30.), np_=(1, 2, 2))], ids=["xyz", "yz"])
def test_restart_continues_like_the_uninterrupted_run(case, tmp_path):
    import psc_b200 as pb
    og = ol.Grid(dt=0.4, kinds=KINDS, nicell=8, **case)
    flds = random_fields(og, seed=5, amp_e=0.02, amp_b=0.05)
    ol.fill_ghosts(og, flds, 3, 9)
    prts, off = thermal_plasma(og, ppc=8, seed=6, vth=(0.3, 0.03))
    # fields held fixed: the particle update is then independent of J's summation order and the
    # comparison can be byte for byte (with evolving fields two identical runs already differ
    # in the last bits of E)
    prm = pb.StepParams(sort=1, marder_loop=0, marder_diffusion=0., push_fields=0, checks=0)

    grid, mprts, mflds = gpu_state(og, flds, prts, off)
    for _ in range(3):
        pb.check(grid.lib.psc_b200_step(grid.ctx, C.byref(prm)))
    grid.timestep = 3
    path = pb.write_checkpoint(grid, str(tmp_path / "checkpoint_3.b200"))
    p_at, o_at = mprts.get()
    f_at = mflds.download()
    for _ in range(2):
        pb.check(grid.lib.psc_b200_step(grid.ctx, C.byref(prm)))
    p_end, o_end = mprts.get()
    grid.close()

    # the file: header, then size_by_patch and the components as PSC's checkpoint lays them out
    raw = open(path + ".0", "rb").read()
    magic, version, hbytes = struct.unpack_from("<QII", raw, 0)
    assert magic == int.from_bytes(b"PSCB200C", "little") and version == 1
    n_p = og.n_patches
    sbp = np.frombuffer(raw, dtype=np.uint32, count=n_p, offset=hbytes)
    assert np.array_equal(sbp, np.diff(o_at))
    n = int(sbp.sum())
    comp = np.frombuffer(raw, dtype=np.float32, count=8 * n, offset=hbytes + 4 * n_p).reshape(8, n)
    assert comp[0].tobytes() == np.ascontiguousarray(p_at["x"][:, 0]).tobytes()
    assert comp[5].tobytes() == np.ascontiguousarray(p_at["u"][:, 2]).tobytes()
    assert np.array_equal(comp[6].view(np.int32), p_at["kind"])
    assert comp[7].tobytes() == p_at["qni_wni"].tobytes()

    # restart in a fresh context
    grid2, mprts2, mflds2 = gpu_state(og, None, None, None)
    assert pb.read_checkpoint(path, grid2) == 3 and grid2.timestep == 3
    p_r, o_r = mprts2.get()
    assert np.array_equal(o_r, o_at) and p_r.tobytes() == p_at.tobytes()
    assert mflds2.download().tobytes() == f_at.tobytes()
    for _ in range(2):
        pb.check(grid2.lib.psc_b200_step(grid2.ctx, C.byref(prm)))
    p_end2, o_end2 = mprts2.get()
    assert np.array_equal(o_end2, o_end) and p_end2.tobytes() == p_end.tobytes()
    grid2.close()


def test_checkpoint_of_another_grid_is_refused(tmp_path):
    import psc_b200 as pb
    og = ol.Grid(gdims=(8, 8, 8), length=(8., 8., 8.), np_=(1, 1, 1), dt=0.4, kinds=KINDS, nicell=4)
    prts, off = thermal_plasma(og, ppc=4, seed=6, vth=(0.3, 0.03))
    grid, _, _ = gpu_state(og, og.zeros_fields(), prts, off)
    path = pb.write_checkpoint(grid, str(tmp_path / "cp.b200"))
    grid.close()
    og2 = ol.Grid(gdims=(8, 8, 8), length=(8., 8., 9.), np_=(1, 1, 1), dt=0.4, kinds=KINDS, nicell=4)
    grid2, _, _ = gpu_state(og2, None, None, None)
    with pytest.raises(pb.PscB200Error):
        pb.read_checkpoint(path, grid2)
    with pytest.raises(pb.PscB200Error):
        pb.read_checkpoint(str(tmp_path / "missing.b200"), grid2)
    grid2.close()
