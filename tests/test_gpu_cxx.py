"""The C++ PscConfig wrapper types (include/psc_b200/psc_config_b200.hxx) driven like a PSC
deck by tests/cxx/test_wrappers.cxx, compared with the CPU oracle stepping the same initial
state: same particle migration (exact per-patch counts), x/u and fields within the
multi-step tolerances of test_gpu_fields.py, continuity at round-off."""
import os
import struct
import subprocess

import numpy as np
import pytest

import oracle_lib as ol
from psc_b200.api import PRT_DTYPE

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CXX_DIR = os.path.join(ROOT, "tests", "cxx")
KINDS = ((-1., 1.), (1., 100.))
N_STEPS = 4


def build_driver():
    subprocess.check_call(["make", "-s", "-C", CXX_DIR])
    return os.path.join(CXX_DIR, "test_wrappers")


def test_wrapper_driver_builds():
    """CPU-side check: the header compiles against a Grid_t look-alike and links against
    the C-ABI library (no device needed)"""
    exe = build_driver()
    assert os.access(exe, os.X_OK)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 1 and "usage" in out.stderr


def _read_dump(path):
    with open(path, "rb") as f:
        buf = f.read()
    hdr = struct.unpack_from("8i", buf, 0)
    n_patches, n0, n1, nf = hdr[:4]
    pos = 32
    def take(dtype, count):
        nonlocal pos
        a = np.frombuffer(buf, dtype=dtype, count=count, offset=pos).copy()
        pos += a.nbytes
        return a
    off0 = take(np.uint32, n_patches + 1)
    prts0 = take(PRT_DTYPE, n0)
    flds0 = take(np.float32, nf)
    off1 = take(np.uint32, n_patches + 1)
    prts1 = take(PRT_DTYPE, n1)
    flds1 = take(np.float32, nf)
    tail = take(np.float64, 10)
    nn = take(np.int32, 2)
    mom_n = take(np.float32, int(nn[0]))
    mom_all = take(np.float32, int(nn[1]))
    no = take(np.int32, 3)
    outputs = [take(np.float32, int(k)) for k in no]  # tfd jeh, tfd moments, pfd jeh of step 2
    return hdr, off0, prts0, flds0, off1, prts1, flds1, tail, mom_n, mom_all, outputs


def _interior(og, a):
    b = og.ibn
    sl = [slice(b[d], a.shape[4 - d] - b[d]) for d in range(3)]
    return a[:, :, sl[2], sl[1], sl[0]]


@pytest.mark.gpu
@pytest.mark.parametrize("fused", [0, 1], ids=["operators", "fused_step"])
@pytest.mark.parametrize("dim", ["xyz", "yz"])
def test_wrappers_match_oracle(dim, fused, tmp_path):
    exe = build_driver()
    out = str(tmp_path / "dump.bin")
    r = subprocess.run([exe, out, dim, str(N_STEPS), str(fused)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    hdr, off0, prts0, flds0, off1, prts1, flds1, tail, mom_n, mom_all, outputs = _read_dump(out)
    if dim == "yz":
        og = ol.Grid(gdims=(1, 16, 32), length=(1., 20., 30.), np_=(1, 2, 2), dt=0.3, kinds=KINDS, nicell=6)
    else:
        og = ol.Grid(gdims=(16, 8, 16), length=(16., 8., 16.), np_=(2, 1, 2), dt=0.4, kinds=KINDS, nicell=6)
    assert hdr[0] == og.n_patches
    n_total = og.n_patches * og.n_cells * 12
    assert len(prts0) == n_total and off0[-1] == n_total
    assert abs(tail[1] - n_total) < 1e-3  # sum of w through the accessor: every w is 1
    shape = og.zeros_fields().shape
    f = flds0.reshape(shape).copy()
    # Psc::initialize: ghost fills before the first step (psc.hxx:220-238)
    ol.fill_ghosts(og, f, 0, 9)
    rp, ro = prts0.copy(), off0.copy()
    # OutputFields / OutputMoments (output_fields.hxx:150-236) with pfield every 2, tfield every 4
    # averaging the last 3 steps: what the writers must have been handed
    tfd_jeh = tfd_mom = pfd2 = None
    for step in range(1, N_STEPS + 1):
        rp, ro = ol.step(og, f, rp, ro, sort_now=(step % 2 == 0))
        if step >= 2:
            jeh, mom = _interior(og, f).copy(), _interior(og, ol.moment_1st(og, rp, ro, ol.MOM_ALL))
            tfd_jeh = jeh if tfd_jeh is None else tfd_jeh + jeh
            tfd_mom = mom.copy() if tfd_mom is None else tfd_mom + mom
            if step == 2:
                pfd2 = jeh
    tfd_jeh = (np.float64(1. / 3) * tfd_jeh.astype(np.float64)).astype(np.float32)
    tfd_mom = (np.float64(1. / 3) * tfd_mom.astype(np.float64)).astype(np.float32)
    for got, ref, tol in ((outputs[0], tfd_jeh, 2e-5), (outputs[1], tfd_mom, 1e-4), (outputs[2], pfd2, 2e-5)):
        assert got.size == ref.size
        assert np.abs(got.reshape(ref.shape) - ref).max() <= tol * np.abs(ref).max()
    assert tail[0] < 1e-5, "continuity residual %g" % tail[0]
    if fused:
        # the fused step performs the sort of the following step early when one is due;
        # N_STEPS is even and sort_interval 2, so step N_STEPS+1 would not sort: orders agree
        pass
    assert np.array_equal(off1, ro), "per-patch particle counts differ from the oracle"
    gf = flds1.reshape(shape)
    assert np.abs(gf - f).max() <= 2e-5 * np.abs(f).max()
    # same order only if both sorted at the same times; compare as multisets per patch
    for p in range(og.n_patches):
        a, b = prts1[off1[p]:off1[p + 1]], rp[ro[p]:ro[p + 1]]
        ka = np.lexsort((a["x"][:, 1], a["x"][:, 2], a["kind"]))
        kb = np.lexsort((b["x"][:, 1], b["x"][:, 2], b["kind"]))
        assert np.abs(a["x"][ka] - b["x"][kb]).max() <= 1e-5 * max(og.length)
        assert np.abs(a["u"][ka] - b["u"][kb]).max() <= 1e-5
    ref_en = ol.energies(og, f, rp, ro)
    np.testing.assert_allclose(tail[2:10], ref_en, rtol=1e-4)
    # Moment_n_1st / Moments_1st wrappers on the final particles (the dump's, so the inputs
    # are identical; only the summation order differs: atomics)
    for got, which in ((mom_n, ol.MOM_N), (mom_all, ol.MOM_ALL)):
        ref = ol.moment_1st(og, prts1, off1, which)
        got = got.reshape(ref.shape)
        assert np.abs(got - ref).max() <= 2e-5 * np.abs(ref).max()


def test_injector_driver_builds():
    """CPU-side check of BoundaryInjectorB200 / Step::add_injector: compiles and links"""
    build_driver()
    exe = os.path.join(CXX_DIR, "test_injector")
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 1 and "usage" in out.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("fused", [0, 1], ids=["operators", "fused_flag"])
@pytest.mark.parametrize("case", ["one_particle", "many_particles", "many_species"])
def test_boundary_injector_reference_tests(case, fused, tmp_path):
    """the reference's BoundaryInjector integration tests (test_boundary_injector.cxx:106-283)
    through the C++ wrapper types; the driver makes the reference's assertions (counts, species,
    continuity and Gauss after every step), the final state is compared with the oracle"""
    from injector_cases import TestGenerator, injector_grid_kw, run_oracle
    build_driver()
    exe = os.path.join(CXX_DIR, "test_injector")
    out = str(tmp_path / "inj.bin")
    r = subprocess.run([exe, out, case, str(fused)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    with open(out, "rb") as f:
        buf = f.read()
    n_patches, n_prts, n_f, _ = struct.unpack_from("4i", buf, 0)
    pos = 16
    off = np.frombuffer(buf, dtype=np.uint32, count=n_patches + 1, offset=pos)
    pos += off.nbytes
    prts = np.frombuffer(buf, dtype=PRT_DTYPE, count=n_prts, offset=pos)
    pos += prts.nbytes
    flds = np.frombuffer(buf, dtype=np.float32, count=n_f, offset=pos)
    gens = {"one_particle": [TestGenerator(1, 1)], "many_particles": [TestGenerator(-1, 1)],
            "many_species": [TestGenerator(-1, 1), TestGenerator(-1, 0)]}[case]
    og = ol.Grid(**injector_grid_kw())
    rp, ro, errs, rf = run_oracle(og, gens, 2)
    assert np.array_equal(off, ro)
    key = lambda a: np.lexsort([a["u"][:, 1], a["x"][:, 2], a["x"][:, 1], a["kind"]])
    assert prts[key(prts)].tobytes() == rp[key(rp)].tobytes()
    assert np.abs(flds.reshape(rf.shape) - rf).max() <= 2e-6 * np.abs(rf).max()
