"""Runs the transcribed reference known-answer tests (tests/golden/psc_golden.json)
against any backend exposing  push(grid, flds, prts, off) -> None (in place).
Used with the CPU oracle (test_oracle_golden.py) and with the CUDA path through
the C ABI (test_gpu_golden.py)."""
import json
import os

import numpy as np

import oracle_lib as ol

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "psc_golden.json")) as f:
    GOLDEN = json.load(f)

COMP = dict(JXI=0, JYI=1, JZI=2, EX=3, EY=4, EZ=5, HX=6, HY=7, HZ=8)


def push_fixture_grid(dim, deposit=ol.DEPOSIT_SPLIT, np3=(1, 1, 1)):
    """testing.hxx:137-170 (make_psc)"""
    fx = GOLDEN["push_fixture"]
    gdims = list(fx["gdims"])
    if dim == "yz":
        gdims[0] = 1
    return ol.Grid(gdims=gdims, length=(fx["L"],) * 3, np_=np3, dt=fx["dt"],
                   kinds=[tuple(k) for k in fx["kinds"]], nicell=fx["nicell"],
                   deposit=deposit)


def setup_fields(grid, spec):
    """setup_fields.hxx:19-45 + centering.hxx:25-41: E/J on edge centres, H on
    face centres, every point including ghosts."""
    f = grid.zeros_fields()
    if not spec:
        return f
    ld, ibn, dx = grid.ldims, grid.ibn, grid.dx
    n_ghosts = max(ibn)
    for p in range(grid.n_patches):
        xb = grid.patch_xb(p)
        V = grid.fview(f, p)
        rng = [range(0, ld[d]) if grid.g.invar[d] else range(-n_ghosts, ld[d] + n_ghosts)
               for d in range(3)]
        for name, (kind, arg) in spec.items():
            m = COMP[name]
            c = (m - 3) % 3
            is_h = m >= 6
            for k in rng[2]:
                for j in rng[1]:
                    for i in rng[0]:
                        idx = (i, j, k)
                        pos = []
                        for d in range(3):
                            cc = (d != c) if is_h else (d == c)
                            if cc:
                                pos.append(xb[d] + (idx[d] + np.float32(.5)) * dx[d])
                            else:
                                pos.append(xb[d] + idx[d] * dx[d])
                        val = arg if kind == "const" else pos[arg]
                        V[m, i, j, k] += np.float32(val)
    return f


def inject(grid, injections):
    """InjectorSimple (injector_simple.hxx:24-43): x -> patch-relative float,
    qni_wni = w * q.  injections: list of (patch, x_global, u, w, kind)."""
    by_patch = [[] for _ in range(grid.n_patches)]
    for p, x, u, w, kind in injections:
        xb = grid.patch_xb(p)
        rec = np.zeros(1, dtype=ol.PRT_DTYPE)
        rec["x"][0] = np.asarray(x, dtype=np.float64).astype(np.float32) - \
            np.asarray(xb, dtype=np.float64).astype(np.float32)
        rec["u"][0] = np.asarray(u, dtype=np.float32)
        rec["kind"] = kind
        rec["qni_wni"] = np.float32(w * grid.kinds[kind][0])
        by_patch[p].append(rec)
    prts = np.concatenate([r for bp in by_patch for r in bp]) if injections else \
        np.zeros(0, dtype=ol.PRT_DTYPE)
    off = ol.off_from_counts([len(bp) for bp in by_patch])
    return prts, off


def run_push_case(case, dim, push):
    """runSingleParticleTest (testing.hxx:172-219) + checkCurrent (:232-251)"""
    eps = GOLDEN["push_fixture"]["eps"]
    grid = push_fixture_grid(dim)
    flds = setup_fields(grid, case["fields"])
    p0 = case["prt0"]
    prts, off = inject(grid, [(0, p0["x"], p0["u"], p0["w"], p0["kind"])])
    push(grid, flds, prts, off)
    exp = case["expect"][dim]
    xb = grid.patch_xb(0)
    q = grid.kinds[0][0]
    got_pos = prts["x"][0].astype(np.float64) + np.array(xb)
    np.testing.assert_allclose(prts["u"][0], exp["u"], atol=eps, rtol=0)
    np.testing.assert_allclose(prts["qni_wni"][0] / q, exp["w"], atol=eps, rtol=0)
    np.testing.assert_allclose(got_pos, exp["x"], atol=eps, rtol=0)
    assert prts["kind"][0] == exp["kind"]
    if "curr_ref" in case:
        ref = grid.zeros_fields()
        V = grid.fview(ref, 0)
        for m, pos, val in case["curr_ref"][dim]:
            i = 0 if dim == "yz" else pos[0]
            if dim == "yz" and m == 0 and False:
                continue
            V[m, i, pos[1], pos[2]] += np.float32(val)
        # the upstream check spans the whole array; E/B are zero in these cases
        assert np.abs(flds - ref).max() < eps
    return grid, flds, prts


def deposit_grid(dim, deposit):
    fx = GOLDEN["deposit_fixture"]
    gdims = list(fx["gdims"])
    if dim == "yz":
        gdims[0] = 1
    g = ol.Grid(gdims=gdims, length=[float(v) for v in gdims], dt=fx["dt"],
                fnqs=fx["fnqs"], deposit=deposit, kinds=[(1., 1.)])
    return g


def rho_nc_norm(ldims, x, yz):
    """psc::deposit::norm::nc (psc/deposit.hxx:172-191,24-50) with ib = 0, val 1,
    double precision, on an array of shape ldims (no ghosts): used by
    check_continuity (test_current_deposition.cxx:104-125)."""
    rho = np.zeros((ldims[2] + 1, ldims[1] + 1, ldims[0] + 1))
    l = [int(np.floor(v)) for v in x]
    h = [v - li for v, li in zip(x, l)]
    if yz:
        for dz in (0, 1):
            for dy in (0, 1):
                w = (h[1] if dy else 1 - h[1]) * (h[2] if dz else 1 - h[2])
                rho[l[2] + dz, l[1] + dy, 0] += w
    else:
        for dz in (0, 1):
            for dy in (0, 1):
                for dx in (0, 1):
                    w = ((h[0] if dx else 1 - h[0]) * (h[1] if dy else 1 - h[1])
                         * (h[2] if dz else 1 - h[2]))
                    rho[l[2] + dz, l[1] + dy, l[0] + dx] += w
    return rho
