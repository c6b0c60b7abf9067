"""The reference's multi-step known-answer tests (src/libpsc/tests/test_push_particles_2.cxx):
Accel (:23-90) -- 131 particles at rest in E = (1, 2, 3): after n steps u = (1, 2, 3) n to
1e-5 -- and Cyclo (:95-170) -- u = (1, 1, 1) gyrating in H_z = 2 pi / 64 for 64 steps, u
follows the analytic rotation to 1e-2 --, each step = push, particle boundary exchange,
J ghost add + fill, continuity check below 1e-7.  Fixture: testing.hxx:137-170 (16^3 or
1 x 16 x 16 cells, L = 160, dt = 1, nicell 200, periodic).  Run on the CPU oracle (pins it)
and on the device through the operator wrappers."""
import numpy as np
import pytest

import oracle_lib as ol
from golden_cases import push_fixture_grid

N_PRTS = 131
L = 160.


def _grid(dim, q):
    g0 = push_fixture_grid(dim)
    return ol.Grid(gdims=g0.gdims, length=g0.length, np_=(1, 1, 1), dt=g0.dt, kinds=((q, 1.),),
                   fnqs=g0.g.fnqs, eta=g0.g.eta, deposit=g0.deposit)


def _case(name, dim):
    rng = np.random.default_rng(7)
    if name == "accel":
        og = _grid(dim, 1.)
        flds = og.zeros_fields()
        for m, v in ((ol.EX, 1.), (ol.EX + 1, 2.), (ol.EX + 2, 3.)):
            flds[:, m] = v
        u0, w, n_steps, eps = (0., 0., 0.), np.ones(N_PRTS), 10, 1e-5

        def expect(n):
            return np.array([1., 2., 3.]) * (n + 1)
    else:
        og = _grid(dim, 2.)
        n_steps, eps = 64, 1e-2
        flds = og.zeros_fields()
        flds[:, ol.HX + 2] = 2. * np.pi / n_steps
        u0, w = (1., 1., 1.), rng.random(N_PRTS)

        def expect(n):
            a = 2 * np.pi * (0.125 * n_steps - (n + 1)) / n_steps
            a0 = 2 * np.pi * (0.125 * n_steps) / n_steps
            return np.array([np.cos(a) / np.cos(a0), np.sin(a) / np.sin(a0), 1.])
    prts = np.zeros(N_PRTS, dtype=ol.PRT_DTYPE)
    x = rng.random((N_PRTS, 3)) * L
    if dim == "yz":
        x[:, 0] = 0.5 * L  # invariant direction
    prts["x"] = x.astype(np.float32)
    prts["x"] = np.minimum(prts["x"], np.nextafter(np.float32(L), np.float32(0)))
    prts["u"] = np.array(u0, dtype=np.float32)
    prts["kind"] = 0
    prts["qni_wni"] = (w * og.kinds[0][0]).astype(np.float32)
    return og, flds, prts, ol.off_from_counts([N_PRTS]), n_steps, eps, expect


@pytest.mark.parametrize("dim", ["xyz", "yz"])
@pytest.mark.parametrize("name", ["accel", "cyclo"])
def test_oracle_accel_cyclo(name, dim):
    og, flds, prts, off, n_steps, eps, expect = _case(name, dim)
    L_, G = ol.lib(), og.byref()
    for n in range(n_steps):
        rho_m = ol.moment_rho(og, prts, off)
        L_.po_push_mprts(G, ol.ptr(flds), ol.ptr(prts), ol.ptr(off))
        prts, off, n_drop = ol.bnd_particles(og, prts, off)
        assert n_drop == 0 and len(prts) == N_PRTS
        ol.add_ghosts(og, flds, 0, 3)
        ol.fill_ghosts(og, flds, 0, 3)
        rho_p = ol.moment_rho(og, prts, off)
        assert ol.continuity(og, rho_m, rho_p, flds) < 1e-7
        assert np.abs(prts["u"] - expect(n)).max() < eps, n


@pytest.mark.gpu
@pytest.mark.parametrize("keep_sorted", [0, 1])
@pytest.mark.parametrize("dim", ["xyz", "yz"])
@pytest.mark.parametrize("name", ["accel", "cyclo"])
def test_gpu_accel_cyclo(name, dim, keep_sorted):
    import psc_b200 as pb
    from b200_helpers import gpu_state
    og, flds, prts, off, n_steps, eps, expect = _case(name, dim)
    grid, mprts, mflds = gpu_state(og, flds, prts, off, dict(keep_sorted=keep_sorted))
    pushp, bndp, bnd = pb.PushParticles(), pb.BndParticles(grid), pb.Bnd()
    checks = pb.Checks(grid, continuity_interval=1)
    for n in range(n_steps):
        grid.timestep = n + 1
        checks.continuity.before_particle_push(mprts)
        pushp.push_mprts(mprts, mflds)
        bndp(mprts)
        bnd.add_ghosts(mflds, pb.JXI, pb.JXI + 3)
        bnd.fill_ghosts(mflds, pb.JXI, pb.JXI + 3)
        checks.continuity.after_particle_push(mprts, mflds)
        assert checks.continuity.last_max_err < 1e-7, (n, checks.continuity.last_max_err)
        got, got_off = mprts.get()
        assert len(got) == N_PRTS
        assert np.abs(got["u"] - expect(n)).max() < eps, n
    grid.close()
