"""Scaled-down versions of the reference's case decks (BASELINE.json configs) as test
inputs: same geometry class, boundary conditions, kinds, field / density / drift profiles
and cadence as the deck, on a grid small enough for the CPU oracle to run 1000 steps in
under a minute.  Particles are loaded the way SetupParticles does
(src/include/setup_particles.hxx:108-122, 244-252, 307-331): n_in_cell = max(1,
int(n * nicell + .5)) particles at the CELL CENTRE, weight n * nicell / n_in_cell,
momentum ~ N(drift, sqrt(T / m)); fields are sampled at their Yee positions
(src/include/setup_fields.hxx:19-45).  The random stream is numpy's (the decks' own
std::default_random_engine stream is not needed: both sides get the same arrays)."""
import numpy as np

import oracle_lib as ol
from oracle_lib import PRT_DTYPE, EX, HX, off_from_counts

JXI, JYI, JZI, EXc, EYc, EZc, HXc, HYc, HZc = range(9)
# Yee offsets (in cells) of every component: E_d / J_d staggered along d, H_d along the
# other two (src/libpsc/bits/discretization.txt:4-12)
STAGGER = {
    0: (.5, 0, 0), 1: (0, .5, 0), 2: (0, 0, .5),
    3: (.5, 0, 0), 4: (0, .5, 0), 5: (0, 0, .5),
    6: (0, .5, .5), 7: (.5, 0, .5), 8: (.5, .5, 0),
}


def setup_fields(og, func, comps=range(3, 9)):
    """f[p, m, k, j, i] = func(m, x, y, z) at the component's Yee position (global
    coordinates), ghosts included"""
    f = og.zeros_fields()
    dx, ib, im = og.dx, og.ib, og.im
    for p in range(og.n_patches):
        xb = og.patch_xb(p)
        for m in comps:
            s = STAGGER[m]
            ax = []
            for d in range(3):
                idx = np.arange(im[d]) + ib[d]
                st = 0. if og.g.invar[d] else s[d]
                ax.append(xb[d] + (idx + st) * dx[d])  # patch_xb is a global coordinate
            z, y, x = np.meshgrid(ax[2], ax[1], ax[0], indexing="ij")
            f[p, m] = func(m, x, y, z).astype(np.float32)
    return f


def setup_particles(og, nicell, npt, seed, neutralizing=None):
    """npt(kind, x, y, z) -> (n, p[3], T[3]) arrays over the cell centres of a patch;
    neutralizing: the (last) kind whose particle count per cell balances the charge of the
    others (setup_particles.hxx:176-186)"""
    rng = np.random.default_rng(seed)
    ld, dx = og.ldims, og.dx
    chunks, counts = [], []
    for p in range(og.n_patches):
        xb = og.patch_xb(p)
        ax = [(np.arange(ld[d]) + .5) * dx[d] for d in range(3)]
        zc, yc, xc = np.meshgrid(ax[2], ax[1], ax[0], indexing="ij")
        loc = np.stack([xc.ravel(), yc.ravel(), zc.ravel()], axis=1)
        glob = loc + np.array([xb[d] for d in range(3)])  # patch_xb is a global coordinate
        per_kind = []
        n_q_in_cell = np.zeros(len(loc), dtype=np.int64)
        for kind, (q, m) in enumerate(og.kinds):
            n, pd, T = npt(kind, glob[:, 0], glob[:, 1], glob[:, 2])
            n = np.broadcast_to(np.asarray(n, dtype=np.float64), (len(loc),))
            if kind == neutralizing:
                assert kind == len(og.kinds) - 1
                n_in_cell = (-n_q_in_cell / q).astype(np.int64)
            else:
                n_in_cell = np.where(n > 0, np.maximum(1, (n * nicell + .5).astype(np.int64)), 0)
                n_q_in_cell += (q * n_in_cell).astype(np.int64)
            w = np.where(n_in_cell > 0, n * nicell / np.maximum(n_in_cell, 1), 0.)
            rep = np.repeat(np.arange(len(loc)), n_in_cell)
            a = np.zeros(len(rep), dtype=PRT_DTYPE)
            a["x"] = loc[rep].astype(np.float32)
            u = np.empty((len(rep), 3))
            for d in range(3):
                drift = np.broadcast_to(np.asarray(pd[d], dtype=np.float64), (len(loc),))[rep]
                Td = np.broadcast_to(np.asarray(T[d], dtype=np.float64), (len(loc),))[rep]
                u[:, d] = drift + rng.standard_normal(len(rep)) * np.sqrt(Td / m)
            a["u"] = u.astype(np.float32)
            a["kind"] = kind
            a["qni_wni"] = (q * w[rep]).astype(np.float32)
            per_kind.append((rep, a))
        # the loader visits the cells in order and the populations inside a cell
        rep_all = np.concatenate([r for r, _ in per_kind])
        kind_all = np.concatenate([np.full(len(r), k) for k, (r, _) in enumerate(per_kind)])
        a_all = np.concatenate([a for _, a in per_kind])
        order = np.lexsort((kind_all, rep_all))
        chunks.append(a_all[order])
        counts.append(len(order))
    return np.concatenate(chunks), off_from_counts(counts)


def courant_dt(og, cfl=.75):
    """courant_length (src/include/grid.hxx): 1 / sqrt(sum 1/dx_d^2) over the variant dims"""
    inv = sum(1. / og.dx[d] ** 2 for d in range(3) if not og.g.invar[d])
    return cfl / np.sqrt(inv)


def _grid(dt=None, **kw):
    g0 = ol.Grid(dt=1., **kw)
    return ol.Grid(dt=courant_dt(g0), **kw)


def bubble_yz(nicell=64, seed=1, gdims=(1, 64, 96), np_=(1, 2, 3)):
    """psc_bubble_yz (src/psc_bubble_yz.cxx:79-105 parameters, :117-146 grid, :155-204
    particles, :209-268 fields), 1 x 64 x 96 cells instead of 1 x 1024 x 1536"""
    BB, nnb, nn0, MMach, TTe, TTi, MMi = .07, .1, 1., 3., .02, .02, 100.
    LLn = 12.5
    LLB = LLn / 6.
    LLy, LLz = 2. * LLn, 3. * LLn
    og = _grid(gdims=gdims, length=(LLn, LLy, LLz), corner=(0., -.5 * LLy, -.5 * LLz),
               np_=np_, kinds=((-1., 1.), (1., MMi)), nicell=nicell)
    V0 = MMach * np.sqrt(TTe / MMi)

    def npt(kind, x, y, z):
        n = np.full(y.shape, nnb)
        p = [np.zeros(y.shape) for _ in range(3)]
        for ys in (y + .5 * LLy, y - .5 * LLy):
            r = np.sqrt(z ** 2 + ys ** 2)
            ins = r < LLn
            n = n + np.where(ins, (nn0 - nnb) * np.cos(np.pi / 2. * r / LLn) ** 2, 0.)
            rs = np.where(r > 0, r, 1.)
            p[2] = p[2] + np.where(ins & (r > 0), V0 * np.sin(np.pi * r / LLn) * z / rs, 0.)
            p[1] = p[1] + np.where(ins & (r > 0), V0 * np.sin(np.pi * r / LLn) * ys / rs, 0.)
        if kind == 0:
            for ys in (y + .5 * LLy, y - .5 * LLy):
                r = np.sqrt(z ** 2 + ys ** 2)
                sh = (r <= LLn) & (r >= LLn - 2. * LLB)
                p[0] = np.where(sh, -BB * np.pi / (2. * LLB) * np.cos(np.pi * (LLn - r) / (2. * LLB)) / n, p[0])
        T = TTe if kind == 0 else TTi
        return n, p, (T, T, T)

    def fld(m, x, y, z):
        rv = np.zeros(y.shape)
        for ys in (y + .5 * LLy, y - .5 * LLy):
            r = np.sqrt(z ** 2 + ys ** 2)
            sh = (r < LLn) & (r > LLn - 2. * LLB)
            rs = np.where(r > 0, r, 1.)
            s = np.sin(np.pi * (LLn - r) / (2. * LLB))
            if m == HZc:
                rv += np.where(sh, -BB * s * ys / rs, 0.)
            elif m == HYc:
                rv += np.where(sh, BB * s * z / rs, 0.)
            elif m == EXc:
                rv += np.where(sh, MMach * np.sqrt(TTe / MMi) * BB * s * np.sin(np.pi * r / LLn), 0.)
        return rv

    flds = setup_fields(og, fld)
    prts, off = setup_particles(og, nicell, npt, seed)
    return dict(og=og, flds=flds, prts=prts, off=off, sort_interval=10, marder_interval=0)


def harris_yz(nicell=32, seed=2, gdims=(1, 64, 128), np_=(1, 2, 4)):
    """psc_harris_yz (src/psc_harris_yz.cxx:209-245 grid / kinds / boundary conditions,
    :262-330 Harris sheet + background, :337-360 fields, :370,395 cadence), 1 x 64 x 128
    cells: B_z = B0 tanh(y / L) with a perturbation, conducting walls in y, periodic z"""
    mass_ratio, Ti_Te, L_di, nb_n0, wpe_wce, dby_b0 = 25., 5., .5, .05, 2., .03
    b0 = 1. / wpe_wce
    di = np.sqrt(mass_ratio)
    L = L_di * di
    Ly, Lz = 12.8 * L_di * di, 25.6 * L_di * di
    Te = b0 ** 2 / (2. * (1. + Ti_Te))
    Ti = Te * Ti_Te
    og = _grid(gdims=gdims, length=(1., Ly, Lz), corner=(0., -.5 * Ly, 0.), np_=np_,
               kinds=((1., mass_ratio), (-1., 1.)), nicell=nicell,
               bc_fld_lo=[1, 2, 1], bc_fld_hi=[1, 2, 1], bc_prt_lo=[1, 0, 1], bc_prt_hi=[1, 0, 1])
    # drift speeds that carry the sheet current, split by temperature (Harris equilibrium)
    vdri = 2. * Ti / (b0 * L)
    vdre = -2. * Te / (b0 * L)

    def npt(kind, x, y, z):
        sheet = 1. / np.cosh(y / L) ** 2
        n = sheet + nb_n0
        px = (vdri if kind == 0 else vdre) * sheet / n
        T = Ti if kind == 0 else Te
        return n, (px, 0., 0.), (T, T, T)

    def fld(m, x, y, z):
        if m == HZc:
            return b0 * np.tanh(y / L) + dby_b0 * b0 * np.pi / Lz * Ly * np.cos(2. * np.pi * (z - .5 * Lz) / Lz) * \
                np.sin(np.pi * y / Ly) * 0.5
        if m == HYc:
            return -dby_b0 * b0 * np.sin(2. * np.pi * (z - .5 * Lz) / Lz) * np.cos(np.pi * y / Ly)
        return np.zeros(y.shape)

    flds = setup_fields(og, fld)
    prts, off = setup_particles(og, nicell, npt, seed)
    return dict(og=og, flds=flds, prts=prts, off=off, sort_interval=10, marder_interval=100)


def kh_xyz(nicell=8, seed=3, gdims=(16, 16, 16), np_=(2, 2, 2), length=(8., 8., 8.)):
    """psc_kelvin_helmholtz (src/psc_kelvin_helmholtz.cxx:66-170: four kinds -- two electron
    and two ion populations --, conducting walls in y, periodic x / z, a sheared E x B flow
    across y), 3D 16 x 16 x 16 cells in 2 x 2 x 2 patches"""
    mi, Te, Ti, B0, v0, delta = 25., .02, .02, .5, .1, 1.2
    Lx, Ly, Lz = length
    og = _grid(gdims=gdims, length=(Lx, Ly, Lz), corner=(0., -.5 * Ly, 0.), np_=np_,
               kinds=((-1., 1.), (1., mi), (-1., 1.), (1., mi)), nicell=nicell,
               bc_fld_lo=[1, 2, 1], bc_fld_hi=[1, 2, 1], bc_prt_lo=[1, 0, 1], bc_prt_hi=[1, 0, 1])

    def vz(y):
        return v0 * np.tanh(y / delta)

    def npt(kind, x, y, z):
        # populations 0/1 fill y < 0, populations 2/3 fill y > 0 (two-fluid tagging)
        lower = y < 0
        n = np.where(lower if kind < 2 else ~lower, 1., 0.)
        T = Te if kind % 2 == 0 else Ti
        return n, (0., 0., vz(y)), (T, T, T)

    def fld(m, x, y, z):
        if m == HXc:
            return np.full(y.shape, B0)
        if m == EYc:
            return -vz(y) * B0  # E = -v x B
        return np.zeros(y.shape)

    flds = setup_fields(og, fld)
    prts, off = setup_particles(og, nicell, npt, seed)
    return dict(og=og, flds=flds, prts=prts, off=off, sort_interval=10, marder_interval=0)


def flatfoil_yz(nicell=25, seed=4, gdims=(1, 32, 96), np_=(1, 2, 6), length=(1., 32., 96.)):
    """psc_flatfoil_yz (src/psc_flatfoil_yz.cxx:258-279 parameters, :289-343 grid: periodic
    yz, three kinds he_e / e / i with the ions neutralizing, :352-386 background + foil
    target, :110-161 InjectFoil, :461,501 cadence), 1 x 32 x 96 cells instead of the deck's
    1 x 80 x 240 (CASE_2D_SMALL); heating, injection and collisions off.  No
    initial fields (BB = 0): the field energy is noise, the particle energy is the signal."""
    mass_ratio, target_n, he_ratio = 100., 2.5, .01
    T_target, bg_n, bg_T = .001, .002, .001
    d_i = np.sqrt(mass_ratio)
    zw = 1. * d_i
    _, Ly, Lz = length
    og = _grid(gdims=gdims, length=(1., Ly, Lz), corner=(-.5, -.5 * Ly, -.5 * Lz), np_=np_,
               kinds=((-1., 1.), (-1., 1.), (1., mass_ratio)), nicell=nicell)

    def npt(kind, x, y, z):
        inside = np.abs(z) <= zw
        if kind == 2:      # ions
            n = np.where(inside, target_n, bg_n)
        elif kind == 0:    # high-energy electrons: only in the foil
            n = np.where(inside, he_ratio * target_n, 0.)
        else:
            n = np.where(inside, (1. - he_ratio) * target_n, bg_n)
        T = np.where(inside, T_target, bg_T)
        return n, (0., 0., 0.), (T, T, T)

    flds = og.zeros_fields()
    prts, off = setup_particles(og, nicell, npt, seed, neutralizing=2)
    return dict(og=og, flds=flds, prts=prts, off=off, sort_interval=10, marder_interval=100)


DECKS = {"flatfoil_yz": flatfoil_yz, "bubble_yz": bubble_yz, "harris_yz": harris_yz, "kelvin_helmholtz_xyz": kh_xyz}
