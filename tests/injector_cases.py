"""Shared pieces of the boundary-injector tests (CPU oracle tests and GPU parity tests):
the reference's test generator and grid (src/libpsc/tests/test_boundary_injector.cxx:45-104)
and a Psc::step loop over the oracle with the injectors in their slot."""
import numpy as np

import oracle_lib as ol

# float32 state: the checks hold to rounding (the reference's thresholds are for its double config)
CHECK_EPS = 1e-5


class TestGenerator:
    """struct ParticleGenerator of the reference's test (:76-104): particles just below the
    wall (0.999 of the ghost cell) moving in at u_y = 2; after `max_n_injected` draws the
    rest get u_y = 0 and therefore never enter"""
    __test__ = False

    def __init__(self, max_n_injected, kind_idx):
        self.max_n_injected, self.kind_idx, self.n_injected = max_n_injected, kind_idx, 0

    def get(self, min_pos, pos_range):
        uy = 2.0
        if self.max_n_injected > 0:
            self.n_injected += 1
            if self.n_injected > self.max_n_injected:
                uy = 0.0
        x = [m + r * f for m, r, f in zip(min_pos, pos_range, (0.0, 0.999, 0.0))]
        return x, [0.0, uy, 0.0], 1.0, self.kind_idx


def injector_grid_kw(gdims=(1, 8, 2), length=(1., 8., 2.), np_=(1, 1, 1), dt=1.0):
    """setupGrid() (:45-74): 1 x 8 x 2 cells of size 1, open in y (fields and particles),
    electrons (-1, 1) and ions (1, 1), nicell = 1, dt = 1"""
    return dict(gdims=gdims, length=length, np_=np_, dt=dt, kinds=((-1., 1.), (1., 1.)), nicell=1,
                bc_fld_lo=[ol.BND_FLD_PERIODIC, ol.BND_FLD_OPEN, ol.BND_FLD_PERIODIC],
                bc_fld_hi=[ol.BND_FLD_PERIODIC, ol.BND_FLD_OPEN, ol.BND_FLD_PERIODIC],
                bc_prt_lo=[ol.BND_PRT_PERIODIC, ol.BND_PRT_OPEN, ol.BND_PRT_PERIODIC],
                bc_prt_hi=[ol.BND_PRT_PERIODIC, ol.BND_PRT_OPEN, ol.BND_PRT_PERIODIC])


class OracleGeom:
    """what BoundaryInjector.candidates() asks of a grid, answered by the oracle's grid"""

    def __init__(self, og, nicell=1):
        self.og = og
        self.ldims, self.dx, self.prts_per_unit_density = og.ldims, og.dx, float(nicell)

    def n_patches(self):
        return self.og.n_patches

    def at_boundary_lo(self, p, d):
        return self.og.patch_off(p)[d] == 0


def make_injectors(og, generators, n_in_cell=None):
    """one BoundaryInjector (host side only: draws) per generator"""
    from psc_b200.api import BoundaryInjector
    return [BoundaryInjector(gen, OracleGeom(og), n_in_cell=n_in_cell, rng=np.random.default_rng(11 + i))
            for i, gen in enumerate(generators)]


def oracle_checks(og, flds, rho_m, prts, off):
    rho_p = ol.moment_rho(og, prts, off)
    return ol.continuity(og, rho_m, rho_p, flds), ol.gauss(og, rho_p, flds)


def run_oracle(og, generators, n_steps, n_in_cell=None, prts=None, off=None, flds=None, record=None):
    """Psc::step n_steps times with one injector per generator (psc.hxx:321-486, injectors at
    :391-399); returns (prts, off, [(continuity, gauss) per step], flds).  `record`, if given,
    receives every step's list of draws per injector (to replay them on the device)."""
    flds = og.zeros_fields() if flds is None else flds
    prts = np.zeros(0, dtype=ol.PRT_DTYPE) if prts is None else prts
    off = np.zeros(og.n_patches + 1, dtype=np.uint32) if off is None else off
    injectors = make_injectors(og, generators, n_in_cell)
    errs = []
    for _ in range(n_steps):
        rho_m = ol.moment_rho(og, prts, off)
        draws = []

        def inject(flds_, prts_, off_):
            for inj in injectors:
                cand = inj.candidates()
                draws.append(cand)
                prts_, off_ = ol.boundary_inject(og, flds_, prts_, off_, cand)
            return prts_, off_

        prts, off = ol.step(og, flds, prts, off, sort_now=False, inject=inject)
        if record is not None:
            record.append(draws)
        errs.append(oracle_checks(og, flds, rho_m, prts, off))
    return prts, off, errs, flds
