"""The DEFAULT headline path of psc_b200_step -- k_push_lean (tensor-map TMA, moment deposit,
class counts for the leavers) followed by the fused boundary exchange + sort -- against the
CPU oracle, directly: with the fields held fixed (push_fields = 0) the particle update does
not depend on the deposit's summation order, so after k consecutive steps the store must be
BIT-EXACT what (sort, push, exchange) x k of the oracle gives: records, per-patch offsets,
per-cell counts, drop counts.  The same cases run on the older kernels (lean = 0) and on the
packed two-particles-per-lane variant (lean = 2).

One case has the BASELINE.json workload's shape (32^3-cell patches, 64 particles per cell,
two patches per direction): chunks of 32 particles straddle cells at production density,
tiles sit on patch faces, and every cell keeps its neighbours busy."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as ol
from b200_helpers import gpu_state
from gen import random_fields, thermal_plasma
from test_gpu_gapped import CASES, KINDS, _run_oracle

pytestmark = pytest.mark.gpu


def _lean_grid(gkw):
    """the same case with patches 16 cells wide in x (k_push_lean's tile is 16 x 4 x 4 cells;
    narrower patches take k_push_tiled's run-time geometry), cell size unchanged"""
    gkw = dict(gkw)
    g, n, l = list(gkw["gdims"]), gkw["np_"], list(gkw["length"])
    if g[0] > 1 and (g[0] // n[0]) % 16:
        f = 16 * n[0] / g[0]
        g[0], l[0] = 16 * n[0], l[0] * f
    gkw["gdims"], gkw["length"] = tuple(g), tuple(l)
    return gkw


def _steps_vs_oracle(og, flds, prts, off, k, opts, jtol=1e-5, expect_lean=None):
    import psc_b200 as pb
    rf, rp, ro, n_drop = _run_oracle(og, flds, prts, off, k)
    grid, mprts, mflds = gpu_state(og, flds, prts, off, dict(opts, gapped=0))
    prm = pb.StepParams(sort=1, marder_loop=0, marder_diffusion=0., push_fields=0, checks=0)
    for _ in range(k):
        pb.check(grid.lib.psc_b200_step(grid.ctx, C.byref(prm)))
    assert grid.get_stat("fused_steps") == k and grid.get_stat("fused_fallbacks") == 0
    if expect_lean is not None:
        assert grid.get_stat("lean_pushes") == (k if expect_lean else 0)
    j = mflds.download(0, 3)
    got, got_off = mprts.get()
    assert np.array_equal(got_off, ro)
    assert got.tobytes() == rp.tobytes(), "store differs from (sort, push, exchange) x k of the oracle"
    assert np.array_equal(ol.count_by_cell(og, got, got_off), ol.count_by_cell(og, rp, ro))
    assert grid.get_stat("n_dropped") == n_drop
    scale = np.abs(rf[:, :3]).max()
    assert np.abs(j - rf[:, :3]).max() <= jtol * scale
    grid.close()
    return n_drop


@pytest.mark.parametrize("lean", [0, 1, 2])
@pytest.mark.parametrize("k", [1, 3])
@pytest.mark.parametrize("vth", [0.05, 0.5])
@pytest.mark.parametrize("name", list(CASES))
def test_default_path_steps_bit_exact(name, vth, k, lean):
    gkw, opts = CASES[name]
    takes_lean = bool(lean) and "tile" not in opts
    if takes_lean:
        gkw = _lean_grid(gkw)
    dx = [l / g for l, g in zip(gkw["length"], gkw["gdims"])]
    dt = 0.45 * min(d for d, g in zip(dx, gkw["gdims"]) if g > 1)
    og = ol.Grid(dt=dt, kinds=KINDS, nicell=6, **gkw)
    flds = random_fields(og, seed=11)
    prts, off = thermal_plasma(og, ppc=6, seed=12, vth=(vth, vth / 10))
    n_drop = _steps_vs_oracle(og, flds, prts, off, k, dict(opts, lean=lean), expect_lean=takes_lean)
    if "absorbing" in name and vth > 0.1:
        assert n_drop > 0


@pytest.mark.parametrize("lean", [1, 2])
def test_baseline_shape_two_steps_bit_exact(lean):
    """S3D's shape: 2 x 2 x 2 patches of 32^3 cells, 32 + 32 particles per cell, dt = 0.75 / sqrt 3,
    B_z = 0.1 plus a random E/B perturbation so that every field component matters"""
    og = ol.Grid(gdims=(64, 64, 64), length=(64., 64., 64.), np_=(2, 2, 2), dt=0.75 / np.sqrt(3.), kinds=KINDS,
                 nicell=32)
    flds = random_fields(og, seed=21, amp_e=0.02, amp_b=0.02)
    flds[:, ol.HZ] += 0.1
    ol.fill_ghosts(og, flds, 3, 9)
    prts, off = thermal_plasma(og, ppc=32, seed=22, vth=(0.05, 0.005))
    assert len(prts) == 64 ** 3 * 64
    _steps_vs_oracle(og, flds, prts, off, 2, dict(lean=lean), expect_lean=True)
