#!/usr/bin/env python
"""Timings of the operators built from SURVEY.md 8(f) (the rows either side of the hot path)
at the S3D working-set size, each against the bytes it has to move: binary collisions, heating,
the 1st-order moments, Marder, the continuity / Gauss checks, DiagEnergies, the OutputFields
hand-off (running sum, mean, interior copy to the host), the boundary injector's deposit and
checkpoint write / read.  One JSON line.

  python tools/aux_bench.py [--cells 256 --ppc 64 --reps 3]

CUDA events on the library's stream (psc_b200_timer_start / _stop); algorithmic bytes are
stated per operator in the output; peak = MEASURED_PEAKS.json when present."""
import argparse
import ctypes as C
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

KINDS = ((-1., 1.), (1., 100.))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=256)
    ap.add_argument("--ppc", type=int, default=64)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    import psc_b200 as pb
    n = args.cells
    npd = max(1, n // 32)
    grid = pb.Grid(gdims=(n, n, n), length=(float(n),) * 3, np=(npd,) * 3, dt=0.75 / np.sqrt(3.), kinds=KINDS,
                   nicell=args.ppc // 2)
    mprts, mflds = pb.Mparticles(grid), pb.MfieldsState(grid)
    mprts.setup_thermal(args.ppc // 2, [0.05, 0.005], seed=1)
    n_prts = mprts.size()
    n_cells = n ** 3
    fld_pts = grid.n_patches() * int(np.prod(grid.im))
    psc = pb.Psc(grid, mflds, mprts, sort_interval=1, fused=True)
    psc.initialize()
    for _ in range(2):
        psc.step()
    pb.Sort()(mprts)
    peak = None
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak = float(peak.get("hbm_gbs") or peak.get("hbm_gb_s") or 0) or None
    except Exception:
        pass
    peak = peak or 6445.3

    def timed(fn, reps=args.reps, warm=1):
        for _ in range(warm):
            fn()
        grid.sync()
        ms = []
        for _ in range(reps):
            grid.timer_start()
            fn()
            ms.append(grid.timer_stop())
        return float(np.median(ms))

    out = {}

    def rec(name, ms, nbytes, what):
        out[name] = dict(ms=round(ms, 3), algorithmic_GB=round(nbytes / 1e9, 3), GBps=round(nbytes / ms / 1e6, 1),
                         frac_of_hbm_peak=round(nbytes / ms / 1e6 / peak, 3), bytes=what)

    # ---- collisions (psc.hxx:363-371): every particle read and written once
    coll = pb.Collision(grid, 10, 0.1)
    rec("collide", timed(lambda: coll(mprts, step=10)), 64. * n_prts, "u read + written, x read: 48 B/particle (counted 64: both records)")
    # ---- heating (a spot covering an eighth of the box)
    heat = pb.Heating(grid, 20, dict(zl=0.25 * n, zh=0.75 * n, xc=0.5 * n, yc=0.5 * n, rH=0.25 * n, T=[0.04, 0.04], Mi=100.))
    rec("heating_spot_foil", timed(lambda: heat(mprts, step=20)), 32. * n_prts,
        "x read of every particle, u read + written inside the spot: >= 16 B/particle (counted 32)")
    pb.Sort()(mprts)
    # ---- moments
    mom = pb.Moment(grid, pb.MOMENT_ALL)
    rec("moments_1st_all", timed(lambda: mom(mprts)), 32. * n_prts + 2 * 4. * mom.n_comps() * fld_pts,
        "32 B/particle read + the 26-component result written and ghost-added")
    mom_n = pb.Moment(grid, pb.MOMENT_N)
    rec("moment_n_1st", timed(lambda: mom_n(mprts)), 32. * n_prts + 2 * 4. * mom_n.n_comps() * fld_pts, "32 B/particle + 2 components")
    # ---- Marder (3 loops), checks, energies
    marder = pb.Marder(grid, 0.9, 3)
    rec("marder_3_loops", timed(lambda: marder(mflds, mprts)), 32. * n_prts + 3 * 8 * 4. * fld_pts,
        "rho: 32 B/particle; per loop: E read + written, rho, div, res")
    e = C.c_double()
    rec("check_gauss", timed(lambda: pb.check(grid.lib.psc_b200_check_gauss(grid.ctx, C.byref(e)))),
        32. * n_prts + 6 * 4. * fld_pts, "rho: 32 B/particle; div E")
    en = np.zeros(8)
    rec("energies", timed(lambda: pb.check(grid.lib.psc_b200_energies(grid.ctx, en.ctypes.data_as(C.c_void_p)))),
        32. * n_prts + 6 * 4. * fld_pts, "x.w + u of every particle, E and H")
    # ---- OutputFields hand-off
    tfd = pb.Mfields(grid, 9)
    rec("outf_accumulate_jeh", timed(lambda: tfd.add(mflds)), 3 * 9 * 4. * fld_pts, "9 components: 2 reads + 1 write")
    rec("outf_mean_jeh", timed(lambda: tfd.scale(0.5)), 2 * 9 * 4. * fld_pts, "9 components: read + write")
    t0 = time.perf_counter()
    host = tfd.download_interior()
    dt = time.perf_counter() - t0
    out["outf_interior_to_host_jeh"] = dict(ms=round(dt * 1e3, 2), GB=round(host.nbytes / 1e9, 3),
                                            GBps=round(host.nbytes / dt / 1e9, 1),
                                            what="pack on the device + one D2H into PAGEABLE host memory (wall clock)")
    # ---- boundary injector's deposit: 1e6 trajectories
    m = 1_000_000
    rng = np.random.default_rng(0)
    paths = (pb.JPath * m)()
    a = np.frombuffer(paths, dtype=np.dtype([("patch", "<i4"), ("lg", "<i4", (3,)), ("xm", "<f4", (3,)),
                                             ("xp", "<f4", (3,)), ("v", "<f4", (3,)), ("q", "<f4")]))
    a["patch"] = rng.integers(0, grid.n_patches(), m)
    lg = rng.integers(0, 32, (m, 3))
    a["lg"] = lg
    a["xm"] = lg + rng.random((m, 3))
    a["xp"] = a["xm"] + rng.normal(size=(m, 3)) * 0.2
    a["v"] = 0.1
    a["q"] = 1.
    t0 = time.perf_counter()
    pb.check(grid.lib.psc_b200_deposit_j(grid.ctx, paths, m))
    grid.sync()
    dt = time.perf_counter() - t0
    out["deposit_j_1e6_paths"] = dict(ms=round(dt * 1e3, 2), what="H2D of 52-byte records from pageable memory + one thread "
                                      "per trajectory (wall clock)", paths_per_s=round(m / dt))
    # ---- checkpoint
    d = tempfile.mkdtemp()
    t0 = time.perf_counter()
    path = pb.write_checkpoint(grid, os.path.join(d, "ck"))
    t1 = time.perf_counter()
    size = sum(os.path.getsize(os.path.join(d, f)) for f in os.listdir(d))
    pb.read_checkpoint(path, grid)
    grid.sync()
    t2 = time.perf_counter()
    out["checkpoint"] = dict(GB=round(size / 1e9, 2), write_s=round(t1 - t0, 2), read_s=round(t2 - t1, 2),
                             write_GBps=round(size / (t1 - t0) / 1e9, 2), read_GBps=round(size / (t2 - t1) / 1e9, 2),
                             what="local tmp file system of the box; D2H + fwrite / fread + H2D")
    for f in os.listdir(d):
        os.remove(os.path.join(d, f))
    os.rmdir(d)
    print(json.dumps(dict(workload="S3D-thermal %d^3 cells x %d ppc, %d particles, 32^3-cell patches" % (n, args.ppc, n_prts),
                          hbm_peak_GBps=peak, operators=out)))
    grid.close()


if __name__ == "__main__":
    main()
