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


# ----------------------------------------------------------------------------
# moments: src/libpsc/tests/test_moments.cxx:52-143 fixture (16^3 cells or 1 x 16 x 16,
# L = 160 => dx = 10, one kind q = m = 1, nicell 200, one particle of weight .4) and the
# expectations of its TYPED_TESTs: only component 0 is checked, over the patch interior,
# tolerance eps = 1e-6
MOMENT_W, MOMENT_NICELL = .4, 200
MOMENT_CASES = [
    # name, moment, particle x, u, expected (cell or node) -> value factor for xyz / yz
    dict(name="Moment_n_1", which=ol.MOM_N, x=(5., 5., 5.), u=(0., 0., 1.), ref="test_moments.cxx:149-176",
         xyz={(0, 0, 0): 1.}, yz={(0, 0, 0): 1.}),
    dict(name="Moments_1st", which=ol.MOM_ALL, x=(5., 5., 5.), u=(0., 0., 1.), ref=":178-206",
         xyz={(0, 0, 0): 1.}, yz={(0, 0, 0): 1.}),
    dict(name="Moment_n_2", which=ol.MOM_N, x=(25., 5., 5.), u=(0., 0., 1.), ref=":236-264",
         xyz={(2, 0, 0): 1.}, yz={(0, 0, 0): 1.}),
    dict(name="Moment_v_1st", which=ol.MOM_V, x=(5., 5., 5.), u=(.001, .002, .003), ref=":266-292",
         xyz={(0, 0, 0): .001}, yz={(0, 0, 0): .001}),
    dict(name="Moment_p_1st", which=ol.MOM_P, x=(5., 5., 5.), u=(.001, .002, .003), ref=":294-320",
         xyz={(0, 0, 0): .001}, yz={(0, 0, 0): .001}),
    dict(name="Moment_rho_1st_nc_cc", which=ol.MOM_RHO_NC, x=(5., 5., 5.), u=(0., 0., 0.), ref=":322-364",
         xyz={(i, j, k): 1. / 8. for i in (0, 1) for j in (0, 1) for k in (0, 1)},
         yz={(0, j, k): 1. / 4. for j in (0, 1) for k in (0, 1)}),
    dict(name="Moment_rho_1st_nc_nc", which=ol.MOM_RHO_NC, x=(10., 10., 10.), u=(0., 0., 0.), ref=":366-402",
         xyz={(1, 1, 1): 1.}, yz={(0, 1, 1): 1.}),
]


def moment_case_grid(case, dim):
    yz = dim == "yz"
    g = ol.Grid(gdims=(1 if yz else 16, 16, 16), length=(160., 160., 160.), np_=(1, 1, 1), dt=1.,
                kinds=((1., 1.),), nicell=MOMENT_NICELL)
    prts = np.zeros(1, dtype=ol.PRT_DTYPE)
    prts["x"][0] = case["x"]
    prts["u"][0] = case["u"]
    prts["kind"][0] = 0
    prts["qni_wni"][0] = MOMENT_W  # q = 1
    return g, prts, ol.off_from_counts([1])


def moment_case_expected(case, dim, grid):
    """component 0 over the interior, as the reference's loops check it"""
    ld, ib = grid.ldims, grid.ib
    exp = np.zeros((ld[2], ld[1], ld[0]))
    for (i, j, k), f in case[dim].items():
        exp[k, j, i] = MOMENT_W / MOMENT_NICELL * f
    return exp


def moment_interior_comp0(grid, arr):
    ld, ib = grid.ldims, grid.ib
    return arr[0, 0, -ib[2]:-ib[2] + ld[2], -ib[1]:-ib[1] + ld[1], -ib[0]:-ib[0] + ld[0]]


# ----------------------------------------------------------------------------
# field operators: src/libpsc/tests/test_push_fields.cxx (fixture testing.hxx:137-170:
# 16^3 or 1 x 16 x 16 cells, L = 160, dt = 1, periodic) and test_bnd.cxx (gdims (2|1) x 8 x 4,
# dx = 10, patches {1, 2, 1}, B = 2 ghosts).  Each case returns (got, expected, tolerance)
# for an `ops` object with push_E / push_H / div_nc / marder_apply / fill_ghosts /
# add_ghosts working on numpy arrays in the oracle's layout.

def _interior(grid, a):
    ld, ib = grid.ldims, grid.ib
    return a[..., -ib[2]:-ib[2] + ld[2], -ib[1]:-ib[1] + ld[1], -ib[0]:-ib[0] + ld[0]]


def _coords(grid, p, stagger):
    """global coordinates (z, y, x meshgrid, ghosts included) of points with the given
    per-direction stagger (0 = node, .5 = cell centre)"""
    xb, dx, ib, im = grid.patch_xb(p), grid.dx, grid.ib, grid.im
    ax = [xb[d] + (np.arange(im[d]) + ib[d] + stagger[d]) * dx[d] for d in range(3)]
    return np.meshgrid(ax[2], ax[1], ax[0], indexing="ij")


def field_case_pushf1(dim, ops):
    """test_push_fields.cxx:26-64: EY = sin(kz z), push_H(1) -> HX = kz cos(kz z_cc), eps 1e-2"""
    import decks
    g = push_fixture_grid(dim)
    kz = 2. * np.pi / g.length[2]
    f = decks.setup_fields(g, lambda m, x, y, z: np.sin(kz * z) if m == 4 else np.zeros(z.shape))
    ops.push_H(g, f, 1.)
    z, _, _ = _coords(g, 0, (0, 0, .5))
    return _interior(g, f[0, 6]), _interior(g, kz * np.cos(kz * z)), 1e-2


def field_case_pushf2(dim, ops):
    """:66-110: HX = cos(ky y), push_E(1) -> EZ = ky sin(ky y_cc), eps 1e-2"""
    import decks
    g = push_fixture_grid(dim)
    ky = 2. * np.pi / g.length[1]
    f = decks.setup_fields(g, lambda m, x, y, z: np.cos(ky * y) if m == 6 else np.zeros(y.shape))
    ops.push_E(g, f, 1.)
    _, y, _ = _coords(g, 0, (0, .5, 0))
    return _interior(g, f[0, 5]), _interior(g, ky * np.sin(ky * y)), 1e-2


def field_case_marder_correct(dim, ops):
    """:123-172: EZ = sin(kz z), phi = sin(kz z_node), correct(diffusion 5) ->
    EZ = sin(kz z) + .5 dt diffusion kz cos(kz z) at the EZ points, norm_linf < 1e-3"""
    import decks
    g = push_fixture_grid(dim)
    kz = 2. * np.pi / g.length[2]
    diffusion = 5.
    f = decks.setup_fields(g, lambda m, x, y, z: np.sin(kz * z) if m == 5 else np.zeros(z.shape))
    ref = decks.setup_fields(
        g, lambda m, x, y, z: (np.sin(kz * z) + .5 * g.dt * diffusion * kz * np.cos(kz * z)) if m == 5
        else np.zeros(z.shape))
    z, _, _ = _coords(g, 0, (0, 0, 0))
    # init_phi (:112-121): z = (k - bnd) * dz, i.e. the patch-local node coordinate
    phi = np.sin(kz * (z - g.patch_xb(0)[2])).astype(np.float32)[None, None]
    assert np.abs(_interior(g, f) - _interior(g, ref)).max() > 1e-3
    ops.marder_apply(g, f, np.ascontiguousarray(phi), diffusion)
    return _interior(g, f[0, 3:6]), _interior(g, ref[0, 3:6]), 1e-3


def field_case_div(dim, ops, m0):
    """:191-260 ItemDivE (m0 = EX) / ItemDivJ (m0 = JXI): Y-comp = cos(ky y), Z-comp = sin(kz z) ->
    div = -ky sin(ky y) + kz cos(kz z) at the nodes, norm_linf < 1e-2"""
    import decks
    g = push_fixture_grid(dim)
    ky, kz = 2. * np.pi / g.length[1], 2. * np.pi / g.length[2]
    f = decks.setup_fields(
        g, lambda m, x, y, z: np.cos(ky * y) if m == m0 + 1 else (np.sin(kz * z) if m == m0 + 2 else np.zeros(z.shape)),
        comps=range(m0, m0 + 3))
    got = ops.div_nc(g, f, m0)
    z, y, _ = _coords(g, 0, (0, 0, 0))
    return _interior(g, got[0, 0]), _interior(g, -ky * np.sin(ky * y) + kz * np.cos(kz * z)), 1e-2


def bnd_grid(dim):
    """test_bnd.cxx:15-41"""
    yz = dim == "yz"
    return ol.Grid(gdims=(1 if yz else 2, 8, 4), length=(10. if yz else 20., 80., 40.), np_=(1, 2, 1), dt=.1,
                   kinds=((1., 1.),), nicell=1)


def bnd_case_fill_ghosts(dim, ops):
    """test_bnd.cxx:104-174: interior = 100 ii + 10 jj + kk (global indices), fill_ghosts ->
    every point holds the value of its periodic image; exact"""
    g = bnd_grid(dim)
    gd, ld, ib, im = g.gdims, g.ldims, g.ib, g.im
    f = g.zeros_fields(1)
    exp = g.zeros_fields(1)
    for p in range(g.n_patches):
        off = g.patch_off(p)
        idx = [np.arange(im[d]) + ib[d] + off[d] for d in range(3)]
        kk, jj, ii = np.meshgrid(idx[2], idx[1], idx[0], indexing="ij")
        full = 100 * (ii % gd[0]) + 10 * (jj % gd[1]) + (kk % gd[2])
        exp[p, 0] = full
        inner = np.zeros(full.shape, dtype=bool)
        inner[-ib[2]:-ib[2] + ld[2], -ib[1]:-ib[1] + ld[1], -ib[0]:-ib[0] + ld[0]] = True
        f[p, 0] = np.where(inner, full, 0)
    ops.fill_ghosts(g, f)
    ops.fill_ghosts(g, f)  # "let's do it again" (:173)
    return f, exp, 0.


def bnd_case_add_ghosts(dim, ops):
    """test_bnd.cxx:232-303: every point (ghosts included) = 1, add_ghosts -> interior points
    hold 1 + the number of ghost images that fold onto them, ghosts keep 1; exact"""
    g = bnd_grid(dim)
    ld, ib, im = g.ldims, g.ib, g.im
    f = g.zeros_fields(1)
    f[:] = 1
    exp = np.ones_like(f)
    B = 2
    for p in range(g.n_patches):
        for k in range(ld[2]):
            for j in range(ld[1]):
                for i in range(ld[0]):
                    nx = 0 if dim == "yz" else int(i < B) + int(i >= ld[0] - B)
                    ny = int(j < B) + int(j >= ld[1] - B)
                    nz = int(k < B) + int(k >= ld[2] - B)
                    exp[p, 0, k - ib[2], j - ib[1], i - ib[0]] = (nx + 1) * (ny + 1) * (nz + 1)
    ops.add_ghosts(g, f)
    return f, exp, 0.


FIELD_CASES = {
    "Pushf1": field_case_pushf1,
    "Pushf2": field_case_pushf2,
    "MarderCorrect": field_case_marder_correct,
    "ItemDivE": lambda dim, ops: field_case_div(dim, ops, 3),
    "ItemDivJ": lambda dim, ops: field_case_div(dim, ops, 0),
    "BndFillGhosts": bnd_case_fill_ghosts,
    "BndAddGhosts": bnd_case_add_ghosts,
}


class OracleFieldOps:
    """the field operators of the CPU oracle on numpy arrays"""
    push_E = staticmethod(lambda g, f, dt_fac: ol.push_E(g, f, dt_fac))
    push_H = staticmethod(lambda g, f, dt_fac: ol.push_H(g, f, dt_fac))
    div_nc = staticmethod(lambda g, f, m0: ol.div_nc(g, f, m0))
    marder_apply = staticmethod(lambda g, f, res, diffusion: ol.lib().po_marder_apply(g.byref(), ol.ptr(f), ol.ptr(res), diffusion))
    fill_ghosts = staticmethod(lambda g, f: ol.fill_ghosts(g, f, 0, f.shape[1]))
    add_ghosts = staticmethod(lambda g, f: ol.add_ghosts(g, f, 0, f.shape[1]))
