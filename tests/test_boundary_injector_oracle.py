"""The oracle's restatement of BoundaryInjector::inject (src/include/boundary_injector.hxx:93-160)
run through the reference's own tests for it (src/libpsc/tests/test_boundary_injector.cxx), CPU only.
The reference has no tabulated numbers for this path: its tests assert particle counts and that the
continuity and Gauss checks hold with particles entering through an open wall, which is what pins
the deposit of the way in (a wrong current shows up in both checks at once)."""
import numpy as np

import oracle_lib as ol

from injector_cases import TestGenerator, injector_grid_kw, run_oracle, CHECK_EPS


def test_particle_generator_maxwellian():
    """BoundaryInjectorTest.ParticleGeneratorMaxwellianTest (test_boundary_injector.cxx:11-30):
    zero temperature, zero range => exactly (pos, mean_u), w = 1, the kind it was built with"""
    from psc_b200.api import ParticleGeneratorMaxwellian
    gen = ParticleGeneratorMaxwellian(15, (1.0, 1836.0), [0.0, 5.0, 15.0], [0.0, 0.0, 0.0])
    x, u, w, kind = gen.get([1.0, 2.0, 5.0], [0.0, 0.0, 0.0])
    assert kind == 15 and w == 1.0
    assert list(u) == [0.0, 5.0, 15.0]
    assert list(x) == [1.0, 2.0, 5.0]


def test_particle_generator_maxwellian_moments():
    """finite temperature: mean and spread of the draws (stdev = sqrt(T / m))"""
    from psc_b200.api import ParticleGeneratorMaxwellian
    gen = ParticleGeneratorMaxwellian(0, (-1.0, 4.0), [0.1, 0.0, -0.2], [0.04, 0.01, 0.16],
                                      rng=np.random.default_rng(7))
    draws = [gen.get([1.0, -1.0, 3.0], [1.0, 2.0, 0.5]) for _ in range(20000)]
    x = np.array([d[0] for d in draws])
    u = np.array([d[1] for d in draws])
    assert (x >= [1.0, -1.0, 3.0]).all() and (x < [2.0, 1.0, 3.5]).all()
    np.testing.assert_allclose(u.mean(0), [0.1, 0.0, -0.2], atol=5e-3)
    np.testing.assert_allclose(u.std(0), [0.1, 0.05, 0.2], rtol=3e-2)


def test_integration_1_particle():
    """BoundaryInjectorTest.Integration1Particle (:106-160): the generator lets one particle in;
    after two steps there is one particle and both checks held at every step"""
    og = ol.Grid(**injector_grid_kw())
    prts, off, errs, _ = run_oracle(og, [TestGenerator(1, 1)], n_steps=2)
    assert len(prts) == 1
    for cont, gauss in errs:
        assert cont < CHECK_EPS and gauss < CHECK_EPS


def test_integration_many_particles():
    """BoundaryInjectorTest.IntegrationManyParticles (:162-216)"""
    og = ol.Grid(**injector_grid_kw())
    prts, off, errs, _ = run_oracle(og, [TestGenerator(-1, 1)], n_steps=2)
    assert len(prts) > 1
    for cont, gauss in errs:
        assert cont < CHECK_EPS and gauss < CHECK_EPS


def test_integration_many_species():
    """BoundaryInjectorTest.IntegrationManySpecies (:218-283): one injector per species"""
    og = ol.Grid(**injector_grid_kw())
    prts, off, errs, _ = run_oracle(og, [TestGenerator(-1, 1), TestGenerator(-1, 0)], n_steps=2)
    assert (prts["kind"] == 0).any() and (prts["kind"] == 1).any()
    for cont, gauss in errs:
        assert cont < CHECK_EPS and gauss < CHECK_EPS


def test_rejected_particles_leave_no_trace():
    """a particle that fails to enter the patch (:126-129) is neither injected nor deposited"""
    og = ol.Grid(**injector_grid_kw())
    flds = og.zeros_fields()
    prts = np.zeros(0, dtype=ol.PRT_DTYPE)
    off = np.zeros(og.n_patches + 1, dtype=np.uint32)
    cand = [(0, (0, -1, k), (0.0, -0.001, k + 0.5), (0.0, -2.0, 0.0), 1.0, 1) for k in range(2)]
    prts, off = ol.boundary_inject(og, flds, prts, off, cand)
    assert len(prts) == 0 and not flds.any()


def test_deposit_carries_the_charge_in():
    """the deposited J_y across the wall face carries exactly the part of the particle's charge
    that the wall node gains: sum over the face of J_y * dt = q w (1 - what stays on node -1)"""
    og = ol.Grid(**injector_grid_kw(gdims=(1, 8, 4), length=(1., 8., 4.), dt=0.5))
    flds = og.zeros_fields()
    prts = np.zeros(0, dtype=ol.PRT_DTYPE)
    off = np.zeros(og.n_patches + 1, dtype=np.uint32)
    cand = [(0, (0, -1, 1), (0.0, -0.25, 1.5), (0.0, 3.0, 0.0), 1.0, 1)]
    prts, off = ol.boundary_inject(og, flds, prts, off, cand)
    assert len(prts) == 1
    v = og.fview(flds)
    vy = np.float32(3.0) / np.sqrt(np.float32(10.0))
    y1 = np.float32(-0.25) + np.float32(0.5) * vy
    assert y1 > 0 and abs(prts["x"][0][1] - y1) < 1e-7
    # edge (j = -1 -> 0): the particle travels from -0.25 to 0 inside cell -1, z split .5/.5
    jy_face = sum(v[ol.JYI, 0, -1, k] for k in range(-1, 4))
    assert abs(jy_face * og.dt - 0.25) < 1e-6
    jy_in = sum(v[ol.JYI, 0, 0, k] for k in range(-1, 4))
    assert abs(jy_in * og.dt - float(y1)) < 1e-6
