"""ctypes bindings for the CPU oracle (oracle/libpsc_oracle.so) and, when built,
the reference-header library (oracle/_ref/libpsc_ref.so).

TEST INFRASTRUCTURE ONLY: imported from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs -- never from psc_b200/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

JXI, JYI, JZI, EX, EY, EZ, HX, HY, HZ, NR_FIELDS = range(10)
BND_FLD_OPEN, BND_FLD_PERIODIC, BND_FLD_CONDUCTING_WALL, BND_FLD_ABSORBING = range(4)
BND_PRT_REFLECTING, BND_PRT_PERIODIC, BND_PRT_ABSORBING, BND_PRT_OPEN = range(4)
DEPOSIT_VAR1, DEPOSIT_SPLIT = 0, 1
MAX_KINDS = 10

# ParticleSimple<float> (particle_simple.hxx:10-42): 32-byte AoS record
PRT_DTYPE = np.dtype(
    [("x", "<f4", (3,)), ("u", "<f4", (3,)), ("kind", "<i4"), ("qni_wni", "<f4")]
)
assert PRT_DTYPE.itemsize == 32

i3 = C.c_int * 3
d3 = C.c_double * 3


class PoGrid(C.Structure):
    _fields_ = [
        ("gdims", i3), ("np", i3),
        ("length", d3), ("corner", d3),
        ("dt", C.c_double), ("fnqs", C.c_double), ("eta", C.c_double),
        ("n_kinds", C.c_int),
        ("q", C.c_double * MAX_KINDS), ("m", C.c_double * MAX_KINDS),
        ("bc_fld_lo", i3), ("bc_fld_hi", i3), ("bc_prt_lo", i3), ("bc_prt_hi", i3),
        ("deposit", C.c_int),
        ("ldims", i3), ("ibn", i3), ("im", i3), ("ib", i3), ("invar", i3),
        ("dx", d3), ("dx_inv", d3),
        ("n_patches", C.c_int),
        ("periodic", i3),
    ]


def build_oracle():
    """(re)build libpsc_oracle.so (and _ref when /root/reference is present)."""
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "CC=gcc", "CXX=g++"],
                          stdout=subprocess.DEVNULL)


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(ORACLE_DIR, "libpsc_oracle.so")
        src = [os.path.join(ORACLE_DIR, f) for f in
               ("psc_oracle.c", "psc_oracle.h", "psc_oracle_deposit.inc", "psc_oracle_collision.inc")]
        if (not os.path.exists(path)
                or os.path.getmtime(path) < max(os.path.getmtime(s) for s in src)):
            build_oracle()
        L = C.CDLL(path)
        P = C.c_void_p
        G = C.POINTER(PoGrid)
        L.po_grid_setup.argtypes = [G]
        L.po_neighbor_patch.argtypes = [G, C.c_int, i3]
        L.po_neighbor_patch.restype = C.c_int
        L.po_fld_patch_len.argtypes = [G]
        L.po_fld_patch_len.restype = C.c_long
        L.po_push_mprts.argtypes = [G, P, P, P]
        L.po_push_mprts_range.argtypes = [G, P, P, P, C.c_int, C.c_int]
        L.po_calc_j_f.argtypes = [G, P, P, P, P, C.c_float]
        L.po_calc_j_d.argtypes = [G, P, P, P, P, C.c_double]
        L.po_sort.argtypes = [G, P, P, P]
        L.po_sort.restype = C.c_int
        L.po_sort_range.argtypes = [G, P, P, P, C.c_int, C.c_int]
        L.po_sort_range.restype = C.c_int
        L.po_cell_index.argtypes = [G, P]
        L.po_cell_index.restype = C.c_int
        L.po_count_by_cell.argtypes = [G, P, P, P]
        L.po_bnd_particles.argtypes = [G, P, P, P, P, P, P]
        L.po_fill_ghosts.argtypes = [G, P, C.c_int, C.c_int, C.c_int]
        L.po_add_ghosts.argtypes = [G, P, C.c_int, C.c_int, C.c_int]
        L.po_push_E.argtypes = [G, P, C.c_double]
        L.po_push_H.argtypes = [G, P, C.c_double]
        L.po_bndf_fill_ghosts_E.argtypes = [G, P]
        L.po_bndf_fill_ghosts_H.argtypes = [G, P]
        L.po_bndf_add_ghosts_J.argtypes = [G, P]
        L.po_moment_rho_1st_nc.argtypes = [G, P, P, P]
        L.po_div_nc.argtypes = [G, P, C.c_int, C.c_int, P]
        L.po_moment_n_comps.argtypes = [G, C.c_int]
        L.po_moment_1st.argtypes = [G, P, P, C.c_int, P]
        L.po_continuity.argtypes = [G, P, P, P]
        L.po_continuity.restype = C.c_double
        L.po_gauss.argtypes = [G, P, P]
        L.po_gauss.restype = C.c_double
        L.po_marder_correct.argtypes = [G, P, P, P, C.c_double, C.c_int]
        L.po_marder_apply.argtypes = [G, P, P, C.c_double]
        L.po_energies.argtypes = [G, P, P, P, P]
        L.po_best_mapping.argtypes = [C.c_int, P, C.c_int, P, P]
        L.po_get_loads.argtypes = [G, P, C.c_double, P]
        L.po_binary_collision_f.restype = C.c_float
        L.po_binary_collision_f.argtypes = [P, P] + [C.c_float] * 7
        L.po_binary_collision_d.restype = C.c_double
        L.po_binary_collision_d.argtypes = [P, P] + [C.c_double] * 7
        L.po_collide.restype = C.c_long
        L.po_collide.argtypes = [G, P, P, C.c_int, C.c_double, C.c_double, C.c_int, C.c_uint64, C.c_uint64, C.c_int]
        L.po_describe.restype = C.c_char_p
        _lib = L
    return _lib


def ref_available():
    return os.path.exists(os.path.join(ORACLE_DIR, "_ref", "libpsc_ref.so"))


def ref():
    global _ref
    if _ref is None:
        L = C.CDLL(os.path.join(ORACLE_DIR, "_ref", "libpsc_ref.so"))
        P = C.c_void_p
        L.psc_ref_push_mprts.argtypes = [
            C.c_int, C.c_int, i3, d3, C.c_double, C.c_double, C.c_double, C.c_int,
            P, P, P, i3, i3, C.c_int, P, P]
        L.psc_ref_push_mprts.restype = C.c_int
        L.psc_ref_calc_j.argtypes = [
            C.c_int, C.c_int, C.c_int, i3, d3, C.c_double, C.c_double, P, i3, i3,
            d3, d3, d3, C.c_double]
        L.psc_ref_calc_j.restype = C.c_int
        L.psc_ref_push_p.argtypes = [P, P, P, C.c_float]
        L.psc_ref_describe.restype = C.c_char_p
        _ref = L
    return _ref


def ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Grid:
    """Host-side description of a PSC grid (Grid_t: grid.hxx:68-101) for the oracle."""

    def __init__(self, gdims, length, np_=(1, 1, 1), dt=1.0, kinds=((1.0, 1.0),),
                 nicell=None, fnqs=None, eta=1.0, corner=(0, 0, 0), deposit=None,
                 bc_fld_lo=None, bc_fld_hi=None, bc_prt_lo=None, bc_prt_hi=None):
        g = PoGrid()
        g.gdims = i3(*gdims)
        g.np = i3(*np_)
        g.length = d3(*[float(v) for v in length])
        g.corner = d3(*[float(v) for v in corner])
        g.dt = dt
        # Grid_t::Normalization with dimensionless params (grid.hxx:205-220,265-293):
        # cori = 1/nicell, alpha = wp/wl = 1, eta = 1  => fnqs = 1/nicell
        if fnqs is None:
            fnqs = 1.0 / nicell if nicell else 1.0
        g.fnqs = fnqs
        g.eta = eta
        g.n_kinds = len(kinds)
        for k, (q, m) in enumerate(kinds):
            g.q[k] = q
            g.m[k] = m
        g.bc_fld_lo = i3(*(bc_fld_lo or [BND_FLD_PERIODIC] * 3))
        g.bc_fld_hi = i3(*(bc_fld_hi or [BND_FLD_PERIODIC] * 3))
        g.bc_prt_lo = i3(*(bc_prt_lo or [BND_PRT_PERIODIC] * 3))
        g.bc_prt_hi = i3(*(bc_prt_hi or [BND_PRT_PERIODIC] * 3))
        yz = gdims[0] == 1
        if deposit is None:
            # psc_config.hxx:47-72: dim_yz -> Var1, everything else -> Split
            deposit = DEPOSIT_VAR1 if yz else DEPOSIT_SPLIT
        g.deposit = deposit
        lib().po_grid_setup(C.byref(g))
        self.g = g
        self.kinds = list(kinds)

    # convenience accessors -------------------------------------------------
    @property
    def gdims(self): return tuple(self.g.gdims)
    @property
    def np3(self): return tuple(self.g.np)
    @property
    def ldims(self): return tuple(self.g.ldims)
    @property
    def ibn(self): return tuple(self.g.ibn)
    @property
    def im(self): return tuple(self.g.im)
    @property
    def ib(self): return tuple(self.g.ib)
    @property
    def dx(self): return tuple(self.g.dx)
    @property
    def length(self): return tuple(self.g.length)
    @property
    def n_patches(self): return self.g.n_patches
    @property
    def dt(self): return self.g.dt
    @property
    def deposit(self): return self.g.deposit
    @property
    def is_yz(self): return self.g.gdims[0] == 1
    @property
    def n_cells(self): return self.ldims[0] * self.ldims[1] * self.ldims[2]

    def byref(self):
        return C.byref(self.g)

    def patch_off(self, p):
        npx, npy, _ = self.np3
        idx3 = (p % npx, (p // npx) % npy, p // (npx * npy))
        return tuple(i * l for i, l in zip(idx3, self.ldims))

    def patch_xb(self, p):
        off = self.patch_off(p)
        return tuple(o * dx + c for o, dx, c in zip(off, self.dx, self.g.corner))

    def zeros_fields(self, n_comps=NR_FIELDS, dtype=np.float32):
        im = self.im
        return np.zeros((self.n_patches, n_comps, im[2], im[1], im[0]), dtype=dtype)

    def fview(self, flds, p=0):
        """index helper: returns f(m, i, j, k) accessor honouring ib."""
        ib = self.ib

        class V:
            def __getitem__(s, idx):
                m, i, j, k = idx
                return flds[p, m, k - ib[2], j - ib[1], i - ib[0]]

            def __setitem__(s, idx, v):
                m, i, j, k = idx
                flds[p, m, k - ib[2], j - ib[1], i - ib[0]] = v
        return V()


# ---------------------------------------------------------------------------
# oracle operations on numpy arrays


def off_from_counts(n_by_patch):
    off = np.zeros(len(n_by_patch) + 1, dtype=np.uint32)
    off[1:] = np.cumsum(n_by_patch)
    return off


def push_mprts(grid, flds, prts, off):
    lib().po_push_mprts(grid.byref(), ptr(flds), ptr(prts), ptr(off))


def ref_push_mprts(grid, flds, prts, off):
    g = grid.g
    q = np.array([k[0] for k in grid.kinds], dtype=np.float64)
    m = np.array([k[1] for k in grid.kinds], dtype=np.float64)
    rc = ref().psc_ref_push_mprts(
        1 if grid.is_yz else 0, g.deposit, g.gdims, g.length, g.dt, g.fnqs, g.eta,
        g.n_kinds, ptr(q), ptr(m), ptr(flds), g.im, g.ib, g.n_patches, ptr(prts),
        ptr(off))
    assert rc == 0


def sort(grid, prts, off, want_perm=False):
    perm = np.zeros(len(prts), dtype=np.uint32) if want_perm else None
    rc = lib().po_sort(grid.byref(), ptr(prts), ptr(off), ptr(perm))
    return rc, perm


def cell_index(grid, x):
    x = np.ascontiguousarray(x, dtype=np.float32)
    return lib().po_cell_index(grid.byref(), ptr(x))


def count_by_cell(grid, prts, off):
    cnt = np.zeros(grid.n_patches * grid.n_cells, dtype=np.uint32)
    lib().po_count_by_cell(grid.byref(), ptr(prts), ptr(off), ptr(cnt))
    return cnt


def bnd_particles(grid, prts, off, rank_of_patch=None):
    out = np.zeros(len(prts), dtype=PRT_DTYPE)
    off_out = np.zeros(grid.n_patches + 1, dtype=np.uint32)
    nd = np.zeros(1, dtype=np.uint32)
    rop = None if rank_of_patch is None else np.ascontiguousarray(rank_of_patch, dtype=np.int32)
    lib().po_bnd_particles(grid.byref(), ptr(prts), ptr(off), ptr(out), ptr(off_out),
                           ptr(rop), ptr(nd))
    return out[: off_out[-1]].copy(), off_out, int(nd[0])


def fill_ghosts(grid, flds, mb, me):
    lib().po_fill_ghosts(grid.byref(), ptr(flds), flds.shape[1], mb, me)


def add_ghosts(grid, flds, mb, me):
    lib().po_add_ghosts(grid.byref(), ptr(flds), flds.shape[1], mb, me)


def push_E(grid, flds, dt_fac):
    lib().po_push_E(grid.byref(), ptr(flds), dt_fac)


def push_H(grid, flds, dt_fac):
    lib().po_push_H(grid.byref(), ptr(flds), dt_fac)


def moment_rho(grid, prts, off):
    rho = grid.zeros_fields(1)
    lib().po_moment_rho_1st_nc(grid.byref(), ptr(prts), ptr(off), ptr(rho))
    return rho


MOM_N, MOM_V, MOM_P, MOM_T, MOM_ALL, MOM_RHO_NC = range(6)


def moment_1st(grid, prts, off, which):
    """Moment_{n,v,p,T}_1st / Moments_1st (cell-centred) / Moment_rho_1st_nc"""
    nc = lib().po_moment_n_comps(grid.byref(), which)
    out = grid.zeros_fields(nc)
    lib().po_moment_1st(grid.byref(), ptr(prts), ptr(off), which, ptr(out))
    return out


def div_nc(grid, flds, m0):
    div = grid.zeros_fields(1)
    lib().po_div_nc(grid.byref(), ptr(flds), flds.shape[1], m0, ptr(div))
    return div


def continuity(grid, rho_m, rho_p, flds):
    return lib().po_continuity(grid.byref(), ptr(rho_m), ptr(rho_p), ptr(flds))


def gauss(grid, rho, flds):
    return lib().po_gauss(grid.byref(), ptr(rho), ptr(flds))


def marder(grid, flds, prts, off, diffusion, loop):
    lib().po_marder_correct(grid.byref(), ptr(flds), ptr(prts), ptr(off), diffusion, loop)


class PoHeating(C.Structure):
    """po_heating_prm"""
    _fields_ = [("zl", C.c_double), ("zh", C.c_double), ("xc", C.c_double), ("yc", C.c_double), ("rH", C.c_double),
                ("T", C.c_double * 10), ("Mi", C.c_double), ("n_kinds", C.c_int), ("interval", C.c_int),
                ("seed", C.c_uint64), ("step", C.c_uint64)]


def heating(grid, prts, off, interval, spot, seed=0, step=0, patch_begin=0):
    """Heating__::operator() with HeatingSpotFoil (in place); returns the number of particles kicked"""
    hp = PoHeating(zl=spot["zl"], zh=spot["zh"], xc=spot["xc"], yc=spot["yc"], rH=spot["rH"], Mi=spot["Mi"],
                   n_kinds=len(spot["T"]), interval=interval, seed=seed, step=step)
    for k, t in enumerate(spot["T"]):
        hp.T[k] = t
    L = lib()
    L.po_heating.restype = C.c_long
    L.po_heating.argtypes = [C.POINTER(PoGrid), C.POINTER(PoHeating), C.c_void_p, C.c_void_p, C.c_int]
    return L.po_heating(grid.byref(), C.byref(hp), ptr(prts), ptr(off), patch_begin)


RNG_FAKE, RNG_HASH = 0, 1


def collide(grid, prts, off, interval, nu, cori, rng=RNG_HASH, seed=0, step=0, patch_begin=0):
    """CollisionHost::operator() on a cell-sorted store (in place); returns the number of binary collisions"""
    n = lib().po_collide(grid.byref(), ptr(prts), ptr(off), interval, nu, cori, rng, seed, step, patch_begin)
    assert n >= 0, "the store is not ordered by cell"
    return n


def energies(grid, flds, prts, off):
    out = np.zeros(8, dtype=np.float64)
    lib().po_energies(grid.byref(), ptr(flds), ptr(prts), ptr(off), ptr(out))
    return out


class PoInjectCand(C.Structure):
    _fields_ = [("patch", C.c_int), ("idx", C.c_int * 3), ("x", C.c_double * 3), ("u", C.c_double * 3),
                ("w", C.c_double), ("kind", C.c_int)]


def boundary_inject(grid, flds, prts, off, cand):
    """BoundaryInjector::inject on the generator's draws `cand` = [(patch, idx, x, u, w, kind)]:
    returns (prts, off) with the accepted particles appended to their patches (push_back order)
    and the current of their way in added to flds' J."""
    L = lib()
    L.po_boundary_inject.restype = C.c_long
    arr = (PoInjectCand * max(len(cand), 1))()
    for i, (p, idx, x, u, w, kind) in enumerate(cand):
        arr[i].patch, arr[i].w, arr[i].kind = p, w, kind
        for d in range(3):
            arr[i].idx[d], arr[i].x[d], arr[i].u[d] = idx[d], x[d], u[d]
    out = np.zeros(max(len(cand), 1), dtype=PRT_DTYPE)
    out_patch = np.zeros(max(len(cand), 1), dtype=np.int32)
    n = L.po_boundary_inject(grid.byref(), ptr(flds), arr, C.c_long(len(cand)), ptr(out), ptr(out_patch))
    out, out_patch = out[:n], out_patch[:n]
    parts = []
    for p in range(grid.n_patches):
        parts.append(prts[off[p]:off[p + 1]])
        parts.append(out[out_patch == p])
    new = np.concatenate(parts) if parts else prts
    n_by = [int(off[p + 1] - off[p]) + int((out_patch == p).sum()) for p in range(grid.n_patches)]
    return np.ascontiguousarray(new), off_from_counts(n_by)


def step(grid, flds, prts, off, sort_now=True, marder_loop=0, marder_diffusion=0.9, inject=None):
    """Psc::step ordering (psc.hxx:321-486) with the oracle's operators; `inject(flds, prts,
    off) -> (prts, off)` stands for the injectors' slot between push and exchange (:391-399)."""
    L = lib()
    G = grid.byref()
    if sort_now:
        L.po_sort(G, ptr(prts), ptr(off), None)
    L.po_push_mprts(G, ptr(flds), ptr(prts), ptr(off))
    if inject is not None:
        prts, off = inject(flds, prts, off)
    prts, off, _ = bnd_particles(grid, prts, off)
    L.po_bndf_add_ghosts_J(G, ptr(flds))
    L.po_add_ghosts(G, ptr(flds), NR_FIELDS, JXI, JXI + 3)
    L.po_fill_ghosts(G, ptr(flds), NR_FIELDS, JXI, JXI + 3)
    L.po_push_H(G, ptr(flds), .5)
    L.po_bndf_fill_ghosts_H(G, ptr(flds))
    L.po_fill_ghosts(G, ptr(flds), NR_FIELDS, HX, HX + 3)
    L.po_push_E(G, ptr(flds), 1.)
    L.po_bndf_fill_ghosts_E(G, ptr(flds))
    L.po_fill_ghosts(G, ptr(flds), NR_FIELDS, EX, EX + 3)
    if marder_loop:
        marder(grid, flds, prts, off, marder_diffusion, marder_loop)
    L.po_push_H(G, ptr(flds), .5)
    L.po_bndf_fill_ghosts_H(G, ptr(flds))
    L.po_fill_ghosts(G, ptr(flds), NR_FIELDS, HX, HX + 3)
    return prts, off
