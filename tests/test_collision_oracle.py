"""The oracle's restatement of the binary collision operator pinned on the reference's own
known-answer tests (src/libpsc/tests/test_collision.cxx), CPU only."""
import numpy as np

import oracle_lib as ol


def test_binary_collision_test1_double():
    """BinaryCollision.Test1 (test_collision.cxx:32-54): RngFake, nudt1 = .1, momentum conserved to 1e-14"""
    u1, u2 = np.array([1., 0., 0.]), np.zeros(3)
    ol.lib().po_binary_collision_d(ol.ptr(u1), ol.ptr(u2), 1., 1., 1., 1., .1, .5, .5)
    assert abs(u1[0] + u2[0] - 1.) < 1e-14 and abs(u1[1] + u2[1]) < 1e-14 and abs(u1[2] + u2[2]) < 1e-14
    assert np.abs(u1 - [1., 0., 0.]).max() > 0.1  # it did scatter


def _two_particle_grid():
    # make_psc<dim_yz> (test_collision.cxx:61-90): 1 x 16 x 16 cells, length 160, nicell 200, dt 1
    return ol.Grid(gdims=(1, 16, 16), length=(160., 160., 160.), np_=(1, 1, 1), dt=1., kinds=((1., 1.),), nicell=200)


def test_collision_test1_single():
    """CollisionTest.Test1 (test_collision.cxx:130-172): two particles in one cell, interval 1, nu 1,
    RngFake: the tabulated post-collision momenta, eps 1e-5"""
    og = _two_particle_grid()
    prts = np.zeros(2, dtype=ol.PRT_DTYPE)
    prts["x"] = [[5., 5., 5.], [5., 5., 5.]]
    prts["u"] = [[1., 0., 0.], [0., 0., 0.]]
    prts["qni_wni"] = 1.
    off = np.array([0, 2], dtype=np.uint32)
    n = ol.collide(og, prts, off, 1, 1., 1. / 200, rng=ol.RNG_FAKE)
    assert n == 1
    eps = 1e-5
    u0, u1 = prts["u"][0], prts["u"][1]
    assert abs(u0[0] + u1[0] - 1.) < eps and abs(u0[1] + u1[1]) < eps and abs(u0[2] + u1[2]) < eps
    assert abs(u0[0] - 0.96226911) < eps and abs(u0[1]) < eps and abs(abs(u0[2]) - 0.17342988) < eps
    assert abs(u1[0] - 0.03773088) < eps and abs(u1[1]) < eps and abs(abs(u1[2]) - 0.17342988) < eps


def test_collisions_conserve_momentum_and_energy():
    """every binary collision conserves the pair's momentum and energy (equal weights): a property of the
    operator the reference's tests only check for one pair"""
    from gen import thermal_plasma
    kinds = ((-1., 1.), (1., 25.))
    og = ol.Grid(gdims=(8, 8, 8), length=(8., 8., 8.), np_=(2, 1, 1), dt=0.5, kinds=kinds, nicell=10)
    prts, off = thermal_plasma(og, ppc=7, seed=3, vth=(0.3, 0.05))  # odd populations: triangles
    rc, _ = ol.sort(og, prts, off)
    assert rc == 0
    m = np.array([k[1] for k in kinds])[prts["kind"]]
    before = prts["u"].astype(np.float64).copy()
    n = ol.collide(og, prts, off, 10, 0.3, 0.1, rng=ol.RNG_HASH, seed=5, step=20)
    assert n > 0
    after = prts["u"].astype(np.float64)
    assert np.abs(after - before).max() > 1e-3
    cell = np.array([ol.cell_index(og, x) for x in prts["x"]]) + og.n_cells * np.repeat(np.arange(og.n_patches), np.diff(off))
    for d in range(3):
        pb = np.bincount(cell, weights=m * before[:, d])
        pa = np.bincount(cell, weights=m * after[:, d])
        assert np.abs(pa - pb).max() < 2e-5
    eb = np.bincount(cell, weights=m * (np.sqrt(1. + (before ** 2).sum(1)) - 1.))
    ea = np.bincount(cell, weights=m * (np.sqrt(1. + (after ** 2).sum(1)) - 1.))
    assert np.abs(ea - eb).max() < 2e-5 * max(1., eb.max())
    # same streams, same result
    p2 = prts.copy()
    p2["u"] = before.astype(np.float32)
    ol.collide(og, p2, off, 10, 0.3, 0.1, rng=ol.RNG_HASH, seed=5, step=20)
    assert p2.tobytes() == prts.tobytes()
