"""Pipelined host I/O (psc_b200_step_begin / mflds_*_async / io_wait / step_end): the same
numbers as the synchronous sequence upload E,B -> psc_b200_step -> download J, with the
transfers overlapped with the particle re-sort.  Compared bit for bit with the synchronous
loop (same kernels, same order per stream) and, through it, with the oracle (test_gpu_steps)."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as ol
from b200_helpers import gpu_state
from gen import random_fields, thermal_plasma

pytestmark = pytest.mark.gpu
KINDS = ((-1., 1.), (1., 100.))


def _pinned(shape):
    import torch
    return torch.empty(shape, dtype=torch.float32, pin_memory=True).numpy()


@pytest.mark.parametrize("push_fields", [0, 1])
@pytest.mark.parametrize("case", [dict(gdims=(16, 16, 16), length=(16., 16., 16.), np_=(2, 1, 2)),
                                  dict(gdims=(1, 32, 32), length=(1., 32., 32.), np_=(1, 2, 2))], ids=["xyz", "yz"])
def test_pipelined_steps_equal_synchronous_steps(case, push_fields):
    import psc_b200 as pb
    og = ol.Grid(dt=0.4, kinds=KINDS, nicell=8, **case)
    flds = random_fields(og, seed=5, amp_e=0.02, amp_b=0.05)
    ol.fill_ghosts(og, flds, 3, 9)
    prts, off = thermal_plasma(og, ppc=8, seed=6, vth=(0.3, 0.03))
    prm = pb.StepParams(sort=1, marder_loop=0, marder_diffusion=0., push_fields=push_fields, checks=0, energies=1)
    n_steps = 4
    out = []
    for pipelined in (False, True):
        grid, mprts, mflds = gpu_state(og, flds, prts, off)
        lib, ctx = grid.lib, grid.ctx
        h_eb = _pinned(mflds.shape(6))
        h_eb[:] = mflds.download(pb.EX, pb.EX + 6)
        h_j = _pinned(mflds.shape(3))
        js, ens = [], []
        en = np.zeros(8)
        for n in range(n_steps):
            if not pipelined:
                pb.check(lib.psc_b200_mflds_upload(ctx, 0, pb.EX, pb.EX + 6, h_eb.ctypes.data_as(C.c_void_p)))
                pb.check(lib.psc_b200_step(ctx, C.byref(prm)))
                pb.check(lib.psc_b200_mflds_download(ctx, 0, pb.JXI, pb.JXI + 3, h_j.ctypes.data_as(C.c_void_p)))
                # DiagEnergies as a separate pass over the particles ...
                pb.check(lib.psc_b200_energies(ctx, en.ctypes.data_as(C.c_void_p)))
            else:
                if n == 0:
                    pb.check(lib.psc_b200_mflds_upload(ctx, 0, pb.EX, pb.EX + 6, h_eb.ctypes.data_as(C.c_void_p)))
                pb.check(lib.psc_b200_step_begin(ctx, C.byref(prm)))
                pb.check(lib.psc_b200_mflds_download_async(ctx, 0, pb.JXI, pb.JXI + 3, h_j.ctypes.data_as(C.c_void_p)))
                pb.check(lib.psc_b200_io_wait(ctx))          # J is on the host, the sort still runs
                js.append(h_j.copy())
                # "the host's field solver" (slowly varying, so that every step sees different
                # fields): next step's E,B go up behind the sort
                h_eb *= np.float32(1.01)
                pb.check(lib.psc_b200_mflds_upload_async(ctx, 0, pb.EX, pb.EX + 6, h_eb.ctypes.data_as(C.c_void_p)))
                pb.check(lib.psc_b200_step_end(ctx))
                # ... and reduced inside the step (fields behind the field chain, particles behind the sort)
                pb.check(lib.psc_b200_last_energies(ctx, en.ctypes.data_as(C.c_void_p)))
            ens.append(en.copy())
            if not pipelined:
                js.append(h_j.copy())
                h_eb *= np.float32(1.01)
        assert grid.get_stat("fused_steps") == n_steps
        got, got_off = mprts.get()
        out.append((js, got, got_off, ens))
        grid.close()
    (j0, p0, o0, e0), (j1, p1, o1, e1) = out
    np.testing.assert_allclose(np.array(e1), np.array(e0), rtol=1e-6, atol=1e-30)
    assert np.array_equal(o0, o1) and p0.tobytes() == p1.tobytes()
    for a, b in zip(j0, j1):
        # J: same kernels, atomics order differs from run to run
        assert np.abs(a - b).max() <= 1e-5 * np.abs(a).max()


def test_forgotten_step_end_is_completed_by_the_next_call():
    import psc_b200 as pb
    og = ol.Grid(gdims=(16, 16, 16), length=(16., 16., 16.), np_=(2, 2, 1), dt=0.4, kinds=KINDS, nicell=8)
    flds = random_fields(og, seed=5, amp_e=0.02, amp_b=0.05)
    prts, off = thermal_plasma(og, ppc=8, seed=6, vth=(0.3, 0.03))
    prm = pb.StepParams(sort=1, marder_loop=0, marder_diffusion=0., push_fields=0, checks=0)
    grid, mprts, mflds = gpu_state(og, flds, prts, off)
    pb.check(grid.lib.psc_b200_step_begin(grid.ctx, C.byref(prm)))
    got, got_off = mprts.get()  # completes the pending step first
    L, G = ol.lib(), og.byref()
    rp, ro = prts.copy(), off.copy()
    rf = flds.copy()
    assert L.po_sort(G, ol.ptr(rp), ol.ptr(ro), None) == 0
    L.po_push_mprts(G, ol.ptr(rf), ol.ptr(rp), ol.ptr(ro))
    rp, ro, _ = ol.bnd_particles(og, rp, ro)
    assert L.po_sort(G, ol.ptr(rp), ol.ptr(ro), None) == 0
    assert np.array_equal(got_off, ro) and got.tobytes() == rp.tobytes()
    grid.close()
