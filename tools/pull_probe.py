"""pull mode on / off: same particles byte for byte, same fields to rounding; step time"""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np
import psc_b200 as pb


def run(pull, gdims, np3, ppc, steps, bc=None, push_fields=1, prof=False):
    kw = {}
    if bc:
        kw = dict(bc_fld_lo=bc[0], bc_fld_hi=bc[0], bc_prt_lo=bc[1], bc_prt_hi=bc[1])
    grid = pb.Grid(gdims=gdims, length=tuple(float(g) for g in gdims), np=np3, dt=0.5,
                   kinds=((-1., 1.), (1., 25.)), nicell=ppc, **kw)
    grid.set_option("pull", pull)
    mprts, mflds = pb.Mparticles(grid), pb.MfieldsState(grid)
    mprts.setup_thermal(ppc, [0.3, 0.06], seed=1)
    psc = pb.Psc(grid, mflds, mprts, sort_interval=1, fused=True)
    psc.initialize()
    grid.sync()
    t0 = time.time()
    import ctypes as C
    for _ in range(steps):
        if push_fields:
            psc.step()
        else:
            prm = pb.StepParams(sort=1, marder_loop=0, marder_diffusion=0.9, push_fields=0, checks=0)
            pb.check(grid.lib.psc_b200_step(grid.ctx, C.byref(prm)))
    grid.sync()
    if prof:
        pass  # print("   profile", {k: round(v[0] / steps, 3) for k, v in grid.profile().items()})
    dt = (time.time() - t0) / steps * 1e3
    stats = {k: grid.get_stat(k) for k in ("pull_steps", "pull_materialized", "pull_overflows", "fused_steps",
                                           "fused_fallbacks", "n_dropped")}
    prts, off = mprts.get()
    f = mflds.download()
    grid.close()
    return prts, off, f, stats, dt



def frozen(gdims, np3, ppc, steps, bc=None):
    a = run(1, gdims, np3, ppc, steps, bc, push_fields=0, prof=True)
    b = run(0, gdims, np3, ppc, steps, bc, push_fields=0, prof=True)
    same = a[0].tobytes() == b[0].tobytes() and np.array_equal(a[1], b[1])
    jd = np.abs(a[2][:, :3] - b[2][:, :3]).max() / np.abs(b[2][:, :3]).max()
    print(gdims, np3, "frozen fields, %d steps: particles identical: %s, J rel diff %.2e" % (steps, same, jd), a[3], "ms %.2f vs %.2f" % (a[4], b[4]), flush=True)




def celldiag(gd, np3, ppc, steps):
    a = run(1, gd, np3, ppc, steps, None, push_fields=0)
    b = run(0, gd, np3, ppc, steps, None, push_fields=0)
    pa, pb_ = a[0], b[0]
    print(gd, np3, steps, a[3])
    va, vb = pa.view("V32").ravel(), pb_.view("V32").ravel()
    print(" multiset equal", np.sort(va).tobytes() == np.sort(vb).tobytes(), "n", len(va), len(vb))
    diff = np.nonzero(va != vb)[0]
    print(" differing", len(diff), diff[:5], diff[-5:] if len(diff) else "")
    if not len(diff):
        return
    ld = [g // n for g, n in zip(gd, np3)]
    def cell(x):
        return (int(x[2]) * ld[1] + int(x[1])) * ld[0] + int(x[0])
    i = diff[0]
    lo = max(0, i - 12)
    for k in range(lo, i + 12):
        ca, cb = cell(pa[k]["x"]), cell(pb_[k]["x"])
        print("  %d  pull: cell %d x %s | ref: cell %d x %s %s" % (k, ca, np.round(pa[k]["x"], 4), cb, np.round(pb_[k]["x"], 4), "" if va[k] == vb[k] else "<<<"))



for gd, np3, ppc, st in (((64, 32, 32), (2, 1, 1), 8, 6), ((64, 64, 64), (2, 2, 2), 16, 5), ((1, 64, 64), (1, 2, 2), 16, 6),
                         ((96, 32, 32), (3, 1, 1), 8, 6)):
    frozen(gd, np3, ppc, st)
celldiag((64, 64, 64), (2, 2, 2), 8, 6)
