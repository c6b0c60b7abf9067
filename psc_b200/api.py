"""Host-side mirror of PSC's PscConfig plugin surface over the psc_b200 C ABI.

Same type names, argument meaning and call order as the reference's types
(SURVEY.md 8b): Grid ~ Grid_t, Mparticles, MfieldsState, PushParticles, Sort,
BndParticles, Bnd, BndFields, PushFields, Marder, Checks, and Psc whose step()
follows Psc::step (src/include/psc.hxx:321-486).  Every method is one C-ABI call
into the CUDA library; nothing is computed here.  Errors raise PscB200Error
(the reference aborts: libpsc/bits.hxx:35-40)."""
import ctypes as C

import numpy as np

from ._lib import GridDesc, StepParams, CollisionParams, HeatingParams, JPath, PscB200Error, load, check, i3, d3, MAX_KINDS

JXI, JYI, JZI, EX, EY, EZ, HX, HY, HZ, NR_FIELDS = range(10)
BND_FLD_OPEN, BND_FLD_PERIODIC, BND_FLD_CONDUCTING_WALL, BND_FLD_ABSORBING = range(4)
BND_PRT_REFLECTING, BND_PRT_PERIODIC, BND_PRT_ABSORBING, BND_PRT_OPEN = range(4)
DEPOSIT_DEFAULT, DEPOSIT_VAR1, DEPOSIT_SPLIT = -1, 0, 1

# ParticleSimple<float>, src/include/particle_simple.hxx:10-42
PRT_DTYPE = np.dtype(
    [("x", "<f4", (3,)), ("u", "<f4", (3,)), ("kind", "<i4"), ("qni_wni", "<f4")])


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Grid:
    """Grid_t (src/include/grid.hxx:68-101): domain, BCs, kinds, normalisation, dt.
    Owns the device context of this rank."""

    def __init__(self, gdims, length, np=(1, 1, 1), dt=1.0, kinds=((1.0, 1.0),),
                 nicell=None, fnqs=None, eta=1.0, corner=(0., 0., 0.),
                 bc_fld_lo=None, bc_fld_hi=None, bc_prt_lo=None, bc_prt_hi=None,
                 deposit=DEPOSIT_DEFAULT, rank=0, n_ranks=1, n_patches_by_rank=None,
                 device=-1, max_n_prts=0):
        d = GridDesc()
        d.gdims = i3(*gdims)
        d.np = i3(*np)
        d.length = d3(*[float(v) for v in length])
        d.corner = d3(*[float(v) for v in corner])
        d.dt = dt
        if fnqs is None:
            # Grid_t::Normalization, dimensionless (grid.hxx:205-220,265-293): 1/nicell
            fnqs = 1.0 / nicell if nicell else 1.0
        d.fnqs, d.eta = fnqs, eta
        self.cori = 1.0 / nicell if nicell else 1.0  # grid.norm.cori (grid.hxx:288)
        self.prts_per_unit_density = float(nicell) if nicell else 1.0  # grid.norm (grid.hxx:291)
        self.corner, self.length = tuple(float(v) for v in corner), tuple(float(v) for v in length)
        assert len(kinds) <= MAX_KINDS
        d.n_kinds = len(kinds)
        for k, kind in enumerate(kinds):
            d.q[k], d.m[k] = kind[0], kind[1]
        d.bc_fld_lo = i3(*(bc_fld_lo or [BND_FLD_PERIODIC] * 3))
        d.bc_fld_hi = i3(*(bc_fld_hi or [BND_FLD_PERIODIC] * 3))
        d.bc_prt_lo = i3(*(bc_prt_lo or [BND_PRT_PERIODIC] * 3))
        d.bc_prt_hi = i3(*(bc_prt_hi or [BND_PRT_PERIODIC] * 3))
        d.deposit = deposit
        d.rank, d.n_ranks = rank, n_ranks
        self._npr = None
        if n_patches_by_rank is not None:
            self._npr = (C.c_int * n_ranks)(*n_patches_by_rank)
            d.n_patches_by_rank = self._npr
        d.device = device
        d.max_n_prts = max_n_prts
        self.desc = d
        self.kinds = [tuple(k[:2]) for k in kinds]
        # Grid_t::Kind::name (grid.hxx:18-33), the suffix of per-kind output components
        self.kind_names = [k[2] if len(k) > 2 else "k%d" % i for i, k in enumerate(kinds)]
        self.lib = load()
        self.ctx = C.c_void_p()
        check(self.lib.psc_b200_create(C.byref(d), C.byref(self.ctx)))
        ld, ibn = i3(), i3()
        check(self.lib.psc_b200_get_ldims(self.ctx, ld, ibn))
        self.ldims, self.ibn = tuple(ld), tuple(ibn)
        self.im = tuple(l + 2 * b for l, b in zip(self.ldims, self.ibn))
        self.ib = tuple(-b for b in self.ibn)
        self.gdims, self.np3 = tuple(gdims), tuple(np)
        self.dx = tuple(L / g for L, g in zip(length, gdims))
        self.dt = dt
        self.timestep = 0

    def n_patches(self):
        return self.lib.psc_b200_n_patches(self.ctx)

    def patch_begin(self):
        return self.lib.psc_b200_patch_begin(self.ctx)

    def patch_off(self, p):
        """cell offset of local patch p ("bydim" order, mrc_domain_lib.c:21-35)"""
        g = self.patch_begin() + p
        idx3 = (g % self.np3[0], (g // self.np3[0]) % self.np3[1], g // (self.np3[0] * self.np3[1]))
        return tuple(i * l for i, l in zip(idx3, self.ldims))

    def patch_xb(self, p):
        """Grid_t::Patch::xb (grid.hxx:82-86)"""
        return tuple(o * dx + c for o, dx, c in zip(self.patch_off(p), self.dx, self.corner))

    def at_boundary_lo(self, p, d):
        """Grid_t::atBoundaryLo (grid.hxx:142-145)"""
        return self.patch_off(p)[d] == 0

    def sync(self):
        check(self.lib.psc_b200_sync(self.ctx))

    def set_option(self, name, value):
        check(self.lib.psc_b200_set_option(self.ctx, name.encode(), float(value)))

    def get_stat(self, name):
        v = C.c_double()
        check(self.lib.psc_b200_get_stat(self.ctx, name.encode(), C.byref(v)))
        return v.value

    def timer_start(self):
        check(self.lib.psc_b200_timer_start(self.ctx))

    def timer_stop(self):
        ms = C.c_float()
        check(self.lib.psc_b200_timer_stop(self.ctx, C.byref(ms)))
        return ms.value

    def profile(self):
        n = 64
        names = (C.c_char_p * n)()
        ms = (C.c_float * n)()
        cnt = (C.c_uint64 * n)()
        k = self.lib.psc_b200_prof_get(self.ctx, n, names, ms, cnt)
        return {names[i].decode(): (ms[i], cnt[i]) for i in range(k)}

    def profile_reset(self):
        check(self.lib.psc_b200_prof_reset(self.ctx))

    def nccl_init(self, unique_id_bytes):
        buf = (C.c_char * 128).from_buffer_copy(unique_id_bytes)
        check(self.lib.psc_b200_nccl_init(self.ctx, buf))

    @staticmethod
    def nccl_unique_id():
        buf = (C.c_char * 128)()
        check(load().psc_b200_nccl_unique_id(buf))
        return bytes(buf)

    def balance(self, factor_fields=1.0):
        """Balance::operator() (psc_balance_impl.hxx:770-1026): redistributes whole patches
        over the ranks by load = n_prts + factor_fields * n_cells; True if anything moved"""
        ch = C.c_int()
        check(self.lib.psc_b200_balance(self.ctx, float(factor_fields), C.byref(ch)))
        return bool(ch.value)

    def close(self):
        if self.ctx:
            self.lib.psc_b200_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Mparticles:
    """MparticlesB200 ~ MparticlesSimple (src/include/particles_simple.hxx:123-247)"""

    def __init__(self, grid):
        self.grid_ = grid

    def grid(self):
        return self.grid_

    def n_patches(self):
        return self.grid_.n_patches()

    def size(self):
        n = C.c_uint64()
        check(self.grid_.lib.psc_b200_mprts_size(self.grid_.ctx, C.byref(n)))
        return n.value

    def sizeByPatch(self):
        n = np.zeros(self.n_patches(), dtype=np.uint32)
        check(self.grid_.lib.psc_b200_mprts_size_by_patch(self.grid_.ctx, _ptr(n)))
        return n

    def set(self, prts, n_by_patch):
        prts = np.ascontiguousarray(prts, dtype=PRT_DTYPE)
        n = np.ascontiguousarray(n_by_patch, dtype=np.uint32)
        assert len(n) == self.n_patches() and int(n.sum()) == len(prts)
        check(self.grid_.lib.psc_b200_mprts_set(self.grid_.ctx, _ptr(prts), _ptr(n)))

    def inject(self, prts, n_by_patch):
        """injector()[p](...) for every patch at once (injector_simple.hxx:24-43)"""
        prts = np.ascontiguousarray(prts, dtype=PRT_DTYPE)
        n = np.ascontiguousarray(n_by_patch, dtype=np.uint32)
        assert len(n) == self.n_patches() and int(n.sum()) == len(prts)
        check(self.grid_.lib.psc_b200_mprts_inject(self.grid_.ctx, _ptr(prts), _ptr(n)))

    def get(self):
        """get_as<MparticlesSingle>() (src/include/particles.hxx:35-54): (records, off)"""
        n = self.size()
        prts = np.zeros(n, dtype=PRT_DTYPE)
        off = np.zeros(self.n_patches() + 1, dtype=np.uint32)
        check(self.grid_.lib.psc_b200_mprts_get(self.grid_.ctx, _ptr(prts), _ptr(off)))
        return prts, off

    def setup_thermal(self, ppc, vth, seed=0):
        v = np.ascontiguousarray(vth, dtype=np.float64)
        assert len(v) == len(self.grid_.kinds)
        if np.ndim(ppc) == 0:
            check(self.grid_.lib.psc_b200_mprts_setup_thermal(self.grid_.ctx, int(ppc), _ptr(v), seed))
        else:  # one value per local patch
            pp = np.ascontiguousarray(ppc, dtype=np.int32)
            assert len(pp) == self.n_patches()
            check(self.grid_.lib.psc_b200_mprts_setup_thermal_by_patch(self.grid_.ctx, _ptr(pp), _ptr(v), seed))


class Mfields:
    """MfieldsB200 (scratch field container with n_comps components)"""

    def __init__(self, grid, n_comps, field_id=None):
        self.grid_ = grid
        self.n_comps_ = n_comps
        if field_id is None:
            fid = C.c_int()
            check(grid.lib.psc_b200_mflds_create(grid.ctx, n_comps, C.byref(fid)))
            field_id = fid.value
        self.id = field_id

    def grid(self):
        return self.grid_

    def n_comps(self):
        return self.n_comps_

    def n_patches(self):
        return self.grid_.n_patches()

    def shape(self, n_comps=None):
        im = self.grid_.im
        return (self.n_patches(), n_comps or self.n_comps_, im[2], im[1], im[0])

    def upload(self, host, mb=0, me=None):
        me = self.n_comps_ if me is None else me
        host = np.ascontiguousarray(host, dtype=np.float32)
        assert host.shape == self.shape(me - mb), (host.shape, self.shape(me - mb))
        check(self.grid_.lib.psc_b200_mflds_upload(self.grid_.ctx, self.id, mb, me, _ptr(host)))

    def download(self, mb=0, me=None):
        me = self.n_comps_ if me is None else me
        host = np.zeros(self.shape(me - mb), dtype=np.float32)
        check(self.grid_.lib.psc_b200_mflds_download(self.grid_.ctx, self.id, mb, me, _ptr(host)))
        return host

    def zero(self, mb=0, me=None):
        me = self.n_comps_ if me is None else me
        check(self.grid_.lib.psc_b200_mflds_zero(self.grid_.ctx, self.id, mb, me))

    def fill(self, m, value):
        check(self.grid_.lib.psc_b200_mflds_fill(self.grid_.ctx, self.id, m, value))

    def add(self, other, mb=0, other_mb=0, n_comps=None):
        """self[mb + m] += other[other_mb + m] (output_fields.hxx:203)"""
        n = min(self.n_comps_ - mb, other.n_comps_ - other_mb) if n_comps is None else n_comps
        check(self.grid_.lib.psc_b200_mflds_add(self.grid_.ctx, self.id, mb, other.id, other_mb, n))

    def scale(self, a, mb=0, me=None):
        """self[m] = float32(a * float64(self[m])) (output_fields.hxx:221)"""
        me = self.n_comps_ if me is None else me
        check(self.grid_.lib.psc_b200_mflds_scale(self.grid_.ctx, self.id, mb, me, float(a)))

    def download_interior(self, mb=0, me=None):
        """psc::mflds::interior on the host: [p][m][k][j][i] over the patches' own cells"""
        me = self.n_comps_ if me is None else me
        ld = self.grid_.ldims
        host = np.zeros((self.n_patches(), me - mb, ld[2], ld[1], ld[0]), dtype=np.float32)
        check(self.grid_.lib.psc_b200_mflds_download_interior(self.grid_.ctx, self.id, mb, me, _ptr(host)))
        return host


class MfieldsState(Mfields):
    """MfieldsStateB200: 9 components JXI..HZ, ghosts grid.ibn (fields3d.hxx:415-464)"""

    def __init__(self, grid):
        super().__init__(grid, NR_FIELDS, field_id=0)


class PushParticles:
    """PushParticlesB200<Dim>: push_particles_1vb.hxx:27-84"""

    def push_mprts(self, mprts, mflds):
        g = mprts.grid()
        check(g.lib.psc_b200_push_mprts(g.ctx))


class Sort:
    """SortB200: SortCountsort2 semantics (psc_sort_impl.hxx:65-124)"""

    def __call__(self, mprts):
        g = mprts.grid()
        check(g.lib.psc_b200_sort(g.ctx))


class Collision:
    """CollisionB200: Collision_<Mparticles, ...> (psc_collision_impl.hxx:20-274): binary Coulomb
    collisions inside every cell, called every `interval` steps right after the sort
    (psc.hxx:363-371).  rng = 1: counter-based streams keyed by (seed, step, cell, pair);
    rng = 0: RngFake (the reference's known-answer setting)."""

    def __init__(self, grid, interval, nu, rng=1, seed=0):
        assert nu > 0.  # psc_collision_impl.hxx:51
        self.grid_, self.interval_, self.nu, self.rng, self.seed = grid, interval, nu, rng, seed
        self.n_collisions = 0

    def interval(self):
        return self.interval_

    def __call__(self, mprts, step=None):
        g = mprts.grid()
        prm = CollisionParams(interval=self.interval_, nu=self.nu, cori=g.cori, rng=self.rng, seed=self.seed,
                              step=g.timestep if step is None else step)
        n = C.c_uint64()
        check(g.lib.psc_b200_collide(g.ctx, C.byref(prm), C.byref(n)))
        self.n_collisions = n.value
        return n.value


class Heating:
    """HeatingB200: Heating__ with the HeatingSpotFoil profile (psc_heating_impl.hxx:27-76,
    heating_spot_foil.hxx:6-89); `spot` = dict(zl, zh, xc, yc, rH, T=[per kind], Mi)"""

    def __init__(self, grid, interval, spot, seed=0):
        self.grid_, self.interval_, self.spot, self.seed = grid, interval, dict(spot), seed
        self.n_kicked = 0

    def __call__(self, mprts, step=None):
        g = mprts.grid()
        sp = self.spot
        prm = HeatingParams(zl=sp["zl"], zh=sp["zh"], xc=sp["xc"], yc=sp["yc"], rH=sp["rH"], Mi=sp["Mi"],
                            n_kinds=len(sp["T"]), interval=self.interval_, seed=self.seed,
                            step=g.timestep if step is None else step)
        for k, t in enumerate(sp["T"]):
            prm.T[k] = t
        n = C.c_uint64()
        check(g.lib.psc_b200_heating_spot_foil(g.ctx, C.byref(prm), C.byref(n)))
        self.n_kicked = n.value
        return n.value


class ParticleGeneratorMaxwellian:
    """ParticleGeneratorMaxwellian (src/include/boundary_injector.hxx:16-57): uniform position
    inside the cell, each momentum component normal(mean_u, sqrt(T / m)); `kind` = (q, m).
    Host code in the reference (rng::Uniform / rng::Normal), host code here (numpy)."""

    def __init__(self, kind_idx, kind, mean_u, temperature, correct_gamma=False, rng=None):
        self.kind_idx, self.correct_gamma = kind_idx, correct_gamma
        self.mean_u = [float(v) for v in mean_u]
        self.stdev_u = [float(np.sqrt(t / kind[1])) for t in temperature]
        self.rng = rng or np.random.default_rng(0)

    def get(self, min_pos, pos_range):
        x = [m + self.rng.random() * r for m, r in zip(min_pos, pos_range)]
        u = [m + s * self.rng.standard_normal() for m, s in zip(self.mean_u, self.stdev_u)]
        if self.correct_gamma:  # vel_to_4vel (setup_particles.hxx:39-43)
            gamma = 1.0 / np.sqrt(1.0 - sum(c * c for c in u))
            u = [c * gamma for c in u]
        return x, u, 1.0, self.kind_idx


class BoundaryInjector:
    """BoundaryInjectorB200: BoundaryInjector<ParticleGenerator, PushParticles>
    (src/include/boundary_injector.hxx:66-167).  Every step, for every ghost cell just below
    the lower y wall of the patches that touch it, `n_in_cell()` particles of an imaginary
    unit-density population are drawn from `generator.get(min_pos, pos_range)` -- returning
    (x[3], u[3], w, kind), positions patch-local -- advanced one step in y, and those that
    enter the patch are injected, with the current of their way in deposited into J.

    The generator and the cell loop are host code, exactly as in the reference; the device
    takes the accepted particles (Mparticles.inject) and the deposit (psc_b200_deposit_j).
    The arithmetic between the draw and the hand-over runs in the configuration's real_t
    (float32).  `n_in_cell` defaults to get_n_in_cell(1, prts_per_unit_density, true)
    (setup_particles.hxx:110-122): nicell + a uniform draw, truncated."""

    INJECT_DIM = 1

    def __init__(self, generator, grid, n_in_cell=None, rng=None):
        self.generator, self.grid_ = generator, grid
        self.rng = rng or np.random.default_rng(0)
        self.n_in_cell = n_in_cell or (lambda: int(np.float32(grid.prts_per_unit_density) +
                                                   np.float32(self.rng.random())))
        self.n_injected = 0

    def candidates(self):
        """the draws of one inject() call: list of (patch, idx, x, u, w, kind)"""
        g = self.grid_
        out = []
        D = self.INJECT_DIM
        for p in range(g.n_patches()):
            if not g.at_boundary_lo(p, D):
                continue
            ilo, ihi = [0, 0, 0], list(g.ldims)
            ilo[D], ihi[D] = -1, 0
            # VecRange(ilo, ihi) (kg/VecRange.hxx): the last index runs fastest
            for i in range(ilo[0], ihi[0]):
                for j in range(ilo[1], ihi[1]):
                    for k in range(ilo[2], ihi[2]):
                        idx = (i, j, k)
                        # Real3 cell_corner = Double3(idx) * dx: narrowed to real_t
                        corner = [float(np.float32(ii * dx)) for ii, dx in zip(idx, g.dx)]
                        for _ in range(self.n_in_cell()):
                            x, u, w, kind = self.generator.get(corner, g.dx)
                            out.append((p, idx, tuple(x), tuple(u), float(w), int(kind)))
        return out

    def inject(self, mprts, mflds, cand=None):
        g = self.grid_
        D = self.INJECT_DIM
        cand = self.candidates() if cand is None else cand
        f32 = np.float32
        dt = f32(g.dt)
        dxi = [f32(n / L) for n, L in zip(g.gdims, g.length)]  # Real3 dxi = grid.domain.dx_inv (domain.hxx:44)
        n_by_patch = np.zeros(g.n_patches(), dtype=np.uint32)
        recs, paths = [], []
        for (p, idx, x, u, w, kind) in sorted(cand, key=lambda c: c[0]):  # (stable: patch order)
            uf = [f32(c) for c in u]
            root = f32(1.) / np.sqrt(f32(1.) + uf[0] * uf[0] + uf[1] * uf[1] + uf[2] * uf[2])
            v = [c * root for c in uf]
            x0 = [f32(c) for c in x]
            x1 = list(x0)
            x1[D] = x1[D] + f32(1.) * dt * v[D]
            if x1[D] < 0:
                continue  # did not enter the patch
            xb = g.patch_xb(p)
            rec = np.zeros((), dtype=PRT_DTYPE)
            rec["x"] = [f32(float(a) + b) - f32(b) for a, b in zip(x1, xb)]
            rec["u"] = uf
            rec["kind"] = kind
            rec["qni_wni"] = f32(w * g.kinds[kind][0])
            recs.append(rec)
            n_by_patch[p] += 1
            jp = JPath()
            jp.patch = p
            for d in range(3):
                jp.lg[d] = idx[d]
                jp.xm[d] = x0[d] * dxi[d]
                jp.xp[d] = x1[d] * dxi[d]
                jp.v[d] = v[d]
            jp.qni_wni = f32(g.kinds[kind][0] * w)
            paths.append(jp)
        if recs:
            mprts.inject(np.array(recs, dtype=PRT_DTYPE), n_by_patch)
            arr = (JPath * len(paths))(*paths)
            check(g.lib.psc_b200_deposit_j(g.ctx, arr, len(paths)))
        self.n_injected = len(recs)
        return len(recs)


class BndParticles:
    """BndParticlesB200 (bnd_particles_impl.hxx:234-247)"""

    def __init__(self, grid):
        self.grid_ = grid

    def __call__(self, mprts):
        g = mprts.grid()
        check(g.lib.psc_b200_bnd_particles(g.ctx))


class Bnd:
    """BndB200 (psc_bnd_impl.hxx:105-158)"""

    def add_ghosts(self, mflds, mb, me):
        g = mflds.grid()
        check(g.lib.psc_b200_bnd_add_ghosts(g.ctx, mflds.id, mb, me))

    def fill_ghosts(self, mflds, mb, me):
        g = mflds.grid()
        check(g.lib.psc_b200_bnd_fill_ghosts(g.ctx, mflds.id, mb, me))


class BndFields:
    """BndFieldsB200<Dim> (psc_bnd_fields_impl.hxx:27-188)"""

    def fill_ghosts_E(self, mflds):
        g = mflds.grid()
        check(g.lib.psc_b200_bndf_fill_ghosts_E(g.ctx))

    def fill_ghosts_H(self, mflds):
        g = mflds.grid()
        check(g.lib.psc_b200_bndf_fill_ghosts_H(g.ctx))

    def add_ghosts_J(self, mflds):
        g = mflds.grid()
        check(g.lib.psc_b200_bndf_add_ghosts_J(g.ctx))


class PushFields:
    """PushFieldsB200 (psc_push_fields_impl.hxx:134-178)"""

    def push_E(self, mflds, dt_fac, dim=None):
        g = mflds.grid()
        check(g.lib.psc_b200_push_E(g.ctx, dt_fac))

    def push_H(self, mflds, dt_fac, dim=None):
        g = mflds.grid()
        check(g.lib.psc_b200_push_H(g.ctx, dt_fac))


MOMENT_N, MOMENT_V, MOMENT_P, MOMENT_T, MOMENT_ALL, MOMENT_RHO_NC = range(6)


class Moment:
    """ItemMoment (include/fields_item.hxx:97-134) for the 1st-order moments of
    fields_item_moments_1st.hxx:9-37: Moment(grid, MOMENT_N)(mprts) -> Mfields with the
    moment's components (ghost add and reflecting folds done), like Moment_n_1st{grid}(mprts)"""

    NAMES = {MOMENT_N: "n_1st_cc", MOMENT_V: "v_1st_cc", MOMENT_P: "p_1st_cc", MOMENT_T: "T_1st_cc",
             MOMENT_ALL: "all_1st_cc", MOMENT_RHO_NC: "rho_1st_nc"}

    def __init__(self, grid, which):
        self.grid_, self.which = grid, which
        self.n_comps_ = grid.lib.psc_b200_moment_n_comps(grid.ctx, which)
        if self.n_comps_ <= 0:
            raise ValueError("unknown moment %r" % (which,))
        self.mres = Mfields(grid, self.n_comps_)

    # per-kind component stems (psc/moment.hxx:130-133,158-161,185-188,217-220,246-249,281-286)
    STEMS = {MOMENT_N: ["n"], MOMENT_V: ["vx", "vy", "vz"], MOMENT_P: ["px", "py", "pz"],
             MOMENT_T: ["Txx", "Tyy", "Tzz", "Txy", "Txz", "Tyz"],
             MOMENT_ALL: ["rho", "jx", "jy", "jz", "px", "py", "pz", "txx", "tyy", "tzz", "txy", "tyz", "tzx"]}

    def name(self):
        return self.NAMES[self.which]

    def n_comps(self):
        return self.n_comps_

    def comp_names(self):
        """addKindSuffix (fields_item.hxx:22-32): kinds outermost; rho_1st_nc is just "rho" """
        if self.which == MOMENT_RHO_NC:
            return ["rho"]
        return ["%s_%s" % (stem, kn) for kn in self.grid_.kind_names for stem in self.STEMS[self.which]]

    def __call__(self, mprts):
        check(self.grid_.lib.psc_b200_moment_1st(self.grid_.ctx, self.mres.id, self.which))
        return self.mres


class ItemJeh:
    """Item_jeh (fields_item_fields.hxx:14-31): the state fields themselves, no copy"""

    @staticmethod
    def name():
        return "jeh"

    @staticmethod
    def comp_names():
        return ["jx_ec", "jy_ec", "jz_ec", "ex_ec", "ey_ec", "ez_ec", "hx_fc", "hy_fc", "hz_fc"]

    def __call__(self, mflds):
        return mflds


class OutputFieldItemParams:
    """BaseOutputFieldItemParams + OutputTfieldItemParams (output_fields.hxx:63-101)"""

    def __init__(self, out_interval=0, data_dir=".", rn=(0, 0, 0), rx=(10000000,) * 3,
                 average_length=1000000, sample_interval=1):
        self.out_interval, self.data_dir, self.rn, self.rx = out_interval, data_dir, tuple(rn), tuple(rx)
        self.average_length, self.sample_interval = average_length, sample_interval

    def enabled(self):
        return self.out_interval > 0

    def do_out(self, timestep):
        return self.enabled() and timestep % self.out_interval == 0

    def do_accum(self, timestep):
        """:88-100 -- on the sampling steps of the averaging window that ends at the next output"""
        if not self.enabled():
            return False
        n_intervals_elapsed = int((timestep - 1) / self.out_interval)  # (C++ integer division truncates)
        next_out = self.out_interval * (n_intervals_elapsed + 1)
        in_averaging_range = next_out - timestep < self.average_length
        on_averaging_step = (next_out - timestep) % self.sample_interval == 0
        return in_averaging_range and on_averaging_step


class WriterMemory:
    """stands where WriterMRC / WriterADIOS2 stand (writer_mrc.hxx:8-123): open(pfx, dir) once,
    write_step(grid, rn, rx, data, name, comp_names) per output; keeps what it was given.
    `data` is the item's interior on the host, [p][m][k][j][i]."""

    def __init__(self):
        self.pfx = self.dir = None
        self.steps = []

    def __bool__(self):
        return self.pfx is not None

    def open(self, pfx, dir="."):
        assert self.pfx is None
        self.pfx, self.dir = pfx, dir

    def write_step(self, grid, rn, rx, data, name, comp_names):
        self.steps.append(dict(timestep=grid.timestep, rn=rn, rx=rx, data=data, name=name,
                               comp_names=list(comp_names)))


class OutputFieldsItem:
    """OutputFieldsItem<Mfields, MfieldsState, Mparticles, GetItem, Writer>
    (output_fields.hxx:150-236; DiagnosticBase::perform_diagnostic).  The item is evaluated on
    the device (`get_item(mprts, mflds) -> (Mfields, name, comp_names)`), the running sum for
    the time average lives and is updated there, and only what a writer is handed crosses to
    the host: the interior of the item (pfd) or of the mean (tfd)."""

    def __init__(self, get_item, suffix, pfield=None, tfield=None, writer=WriterMemory):
        self.get_item, self.suffix = get_item, suffix
        self.pfield, self.tfield = pfield or OutputFieldItemParams(), tfield or OutputFieldItemParams()
        self.io_pfd, self.io_tfd = writer(), writer()
        self.tfd, self.naccum = None, 0

    def perform_diagnostic(self, mprts, mflds):
        grid = mflds.grid()
        timestep = grid.timestep
        do_pfield = self.pfield.do_out(timestep)
        do_tfield = self.tfield.do_out(timestep)
        do_tfield_accum = self.tfield.do_accum(timestep)
        if not (do_pfield or do_tfield_accum):
            return
        item, name, comp_names = self.get_item(mprts, mflds)
        if do_pfield:
            if not self.io_pfd:
                self.io_pfd.open("pfd" + self.suffix, self.pfield.data_dir)
            self.io_pfd.write_step(grid, self.pfield.rn, self.pfield.rx, item.download_interior(), name, comp_names)
        if do_tfield_accum:
            if self.tfd is None:
                self.tfd = Mfields(grid, item.n_comps())
            self.tfd.add(item)
            self.naccum += 1
        if do_tfield and self.naccum > 0:
            # (naccum == 0 happens at the initial output when average_length < out_interval;
            # the reference dereferences its unallocated tfd_ there)
            if not self.io_tfd:
                self.io_tfd.open("tfd" + self.suffix, self.tfield.data_dir)
            # convert accumulated values to correct temporal mean
            self.tfd.scale(1. / self.naccum)
            self.io_tfd.write_step(grid, self.tfield.rn, self.tfield.rx, self.tfd.download_interior(), name,
                                   comp_names)
            self.naccum = 0
            self.tfd.zero()

    __call__ = perform_diagnostic


def OutputFields(pfield=None, tfield=None, writer=WriterMemory):
    """OutputFields<MfieldsState, Mparticles> = OutputFieldsItem<..., GetItemJeh> (:238-243)"""
    item = ItemJeh()
    return OutputFieldsItem(lambda mprts, mflds: (item(mflds), item.name(), item.comp_names()), "",
                            pfield, tfield, writer)


def OutputMoments(grid, pfield=None, tfield=None, writer=WriterMemory, which=MOMENT_ALL):
    """OutputMoments<MfieldsState, Mparticles, Dim> = OutputFieldsItem<..., GetItemMoments<Dim>>
    (:245-250): Moments_1st evaluated on the device"""
    item = Moment(grid, which)
    return OutputFieldsItem(lambda mprts, mflds: (item(mprts), item.name(), item.comp_names()), "_moments",
                            pfield, tfield, writer)


class Marder:
    """MarderB200 (marder_impl.hxx:150-264): Marder(grid, diffusion, loop, dump)"""

    def __init__(self, grid, diffusion, loop, dump=False):
        self.diffusion, self.loop = diffusion, loop

    def correct_gauss(self, mflds, mprts):
        g = mflds.grid()
        check(g.lib.psc_b200_marder(g.ctx, self.diffusion, self.loop))

    __call__ = correct_gauss


class _Continuity:
    def __init__(self, grid, interval):
        self.grid_, self.check_interval, self.last_max_err = grid, interval, 0.0

    def should_do_check(self, timestep):
        return self.check_interval > 0 and timestep % self.check_interval == 0

    def before_particle_push(self, mprts):
        if self.should_do_check(self.grid_.timestep):
            check(self.grid_.lib.psc_b200_check_continuity_begin(self.grid_.ctx))

    def after_particle_push(self, mprts, mflds):
        if self.should_do_check(self.grid_.timestep):
            e = C.c_double()
            check(self.grid_.lib.psc_b200_check_continuity_end(self.grid_.ctx, C.byref(e)))
            self.last_max_err = e.value


class _Gauss:
    def __init__(self, grid, interval):
        self.grid_, self.check_interval, self.last_max_err = grid, interval, 0.0

    def should_do_check(self, timestep):
        return self.check_interval > 0 and timestep % self.check_interval == 0

    def __call__(self, mprts, mflds):
        if self.should_do_check(self.grid_.timestep):
            e = C.c_double()
            check(self.grid_.lib.psc_b200_check_gauss(self.grid_.ctx, C.byref(e)))
            self.last_max_err = e.value


class Checks:
    """ChecksB200 (checks_impl.hxx:33-215; ChecksParams: checks_params.hxx)"""

    def __init__(self, grid, continuity_interval=0, gauss_interval=0):
        self.continuity = _Continuity(grid, continuity_interval)
        self.gauss = _Gauss(grid, gauss_interval)


def write_checkpoint(grid, path=None):
    """write_checkpoint (checkpoint.hxx:14-46): grid + mprts + mflds of this rank -> <path>.<rank>;
    default name as the reference's, "checkpoint_<timestep>.b200" """
    path = path or "checkpoint_%d.b200" % grid.timestep
    check(grid.lib.psc_b200_checkpoint_write(grid.ctx, path.encode(), grid.timestep))
    return path


def read_checkpoint(path, grid):
    """read_checkpoint (checkpoint.hxx:52-82) into a context created for the same grid; restores
    the time step too"""
    t = C.c_int64()
    check(grid.lib.psc_b200_checkpoint_read(grid.ctx, path.encode(), C.byref(t)))
    grid.timestep = t.value
    return t.value


def energies(grid):
    """DiagEnergies: [EX2 EY2 EZ2 HX2 HY2 HZ2 E_electron E_ion]"""
    out = np.zeros(8, dtype=np.float64)
    check(grid.lib.psc_b200_energies(grid.ctx, _ptr(out)))
    return out


class Psc:
    """Psc<PscConfig1vbecB200>: step() sequences the operators exactly as
    Psc::step does (src/include/psc.hxx:321-486); PscParams cadence fields:
    sort_interval, marder_interval (psc.hxx:66-83).  `fused=True` issues the same
    sequence through the single C-ABI call psc_b200_step (host C++ driver)."""

    def __init__(self, grid, mflds, mprts, sort_interval=1, marder_interval=0,
                 marder_diffusion=0.9, marder_loop=3, checks=None, fused=False, collision=None):
        self.collision = collision
        self.injectors_, self.diagnostics_ = [], []
        self.grid_, self.mflds_, self.mprts_ = grid, mflds, mprts
        self.sort_interval, self.marder_interval = sort_interval, marder_interval
        self.marder = Marder(grid, marder_diffusion, marder_loop)
        self.checks = checks or Checks(grid)
        self.fused = fused
        self.sort_, self.pushp_, self.pushf_ = Sort(), PushParticles(), PushFields()
        self.bnd_, self.bndf, self.bndp_ = Bnd(), BndFields(), BndParticles(grid)

    def add_injector(self, injector):
        """Psc::add_injector (psc.hxx:172-176): anything with inject(mprts, mflds).  Injectors
        act between the push and the particle exchange (psc.hxx:391-399), so a step with
        injectors is issued operator by operator rather than as the single fused call."""
        assert injector is not None
        self.injectors_.append(injector)

    def add_diagnostic(self, diagnostic):
        """Psc::add_diagnostic (psc.hxx:184-198): anything with perform_diagnostic(mprts, mflds)"""
        assert diagnostic is not None
        self.diagnostics_.append(diagnostic)

    def perform_diagnostics(self):
        """psc.hxx:514-528"""
        for diagnostic in self.diagnostics_:
            diagnostic.perform_diagnostic(self.mprts_, self.mflds_)

    def integrate(self, nmax):
        """Psc::integrate (psc.hxx:243-310): ghost fills, initial diagnostics, then step +
        diagnostics until timestep nmax"""
        self.initialize()
        self.perform_diagnostics()
        while self.grid_.timestep < nmax:
            self.step()
            self.perform_diagnostics()

    def initialize(self):
        """psc.hxx:220-238 pre_first_step: fill H, J, E ghosts"""
        self.bndf.fill_ghosts_H(self.mflds_)
        self.bnd_.fill_ghosts(self.mflds_, HX, HX + 3)
        self.bnd_.fill_ghosts(self.mflds_, JXI, JXI + 3)
        self.bndf.fill_ghosts_E(self.mflds_)
        self.bnd_.fill_ghosts(self.mflds_, EX, EX + 3)

    def step(self):
        g, mflds, mprts = self.grid_, self.mflds_, self.mprts_
        g.timestep += 1
        t = g.timestep
        do_sort = self.sort_interval > 0 and t % self.sort_interval == 0
        do_marder = self.marder_interval > 0 and t % self.marder_interval == 0
        if self.collision is not None and self.collision.interval() > 0 and t % self.collision.interval() == 0:
            # psc.hxx:356-371: sort, then collide (the pairing walks cell runs)
            self.sort_(mprts)
            self.collision(mprts, step=t)
        if self.fused and not self.injectors_:
            prm = StepParams(sort=int(do_sort), marder_loop=self.marder.loop if do_marder else 0,
                             marder_diffusion=self.marder.diffusion, push_fields=1,
                             checks=int(self.checks.continuity.should_do_check(t)))
            check(g.lib.psc_b200_step(g.ctx, C.byref(prm)))
            if prm.checks:
                cont, gauss = C.c_double(), C.c_double()
                check(g.lib.psc_b200_last_checks(g.ctx, C.byref(cont), C.byref(gauss)))
                self.checks.continuity.last_max_err, self.checks.gauss.last_max_err = cont.value, gauss.value
            return
        if do_sort:
            self.sort_(mprts)                                    # psc.hxx:356-361
        self.checks.continuity.before_particle_push(mprts)       # :379-384
        self.pushp_.push_mprts(mprts, mflds)                     # :389
        for injector in self.injectors_:                         # :391-399
            injector.inject(mprts, mflds)
        self.bndp_(mprts)                                        # :412
        self.bndf.add_ghosts_J(mflds)                            # :417
        self.bnd_.add_ghosts(mflds, JXI, JXI + 3)                # :418
        self.bnd_.fill_ghosts(mflds, JXI, JXI + 3)               # :419
        self.pushf_.push_H(mflds, .5)                            # :426
        self.bndf.fill_ghosts_H(mflds)                           # :428
        self.bnd_.fill_ghosts(mflds, HX, HX + 3)                 # :432
        self.pushf_.push_E(mflds, 1.)                            # :439
        self.bndf.fill_ghosts_E(mflds)                           # :441
        self.bnd_.fill_ghosts(mflds, EX, EX + 3)                 # :445
        if do_marder:
            self.marder(mflds, mprts)                            # :448-455
        self.pushf_.push_H(mflds, .5)                            # :461
        self.bndf.fill_ghosts_H(mflds)                           # :463
        self.bnd_.fill_ghosts(mflds, HX, HX + 3)                 # :467
        self.checks.continuity.after_particle_push(mprts, mflds)  # :471-476
        self.checks.gauss(mprts, mflds)                          # :479-483
