"""OutputFieldsParamsTest.DoOut / Tfield_DoAccum (src/libpsc/tests/test_mfields_io.cxx:232-259):
the reference's known answers for the output cadence, on the Python mirror and on the C++
wrapper header's parameter structs.  CPU only."""
import os
import subprocess

from psc_b200.api import OutputFieldItemParams, Moment, MOMENT_ALL, MOMENT_N, MOMENT_RHO_NC, ItemJeh

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_do_out():
    prm = OutputFieldItemParams()
    assert not prm.do_out(0)  # should be disabled
    prm.out_interval = 10     # now enabled
    assert prm.do_out(0)


def test_tfield_do_accum():
    prm = OutputFieldItemParams()  # default: use every step between outs
    assert not prm.do_accum(0)     # should be disabled
    prm.out_interval = 100         # now enabled
    assert prm.do_accum(0)         # accum on out step itself
    prm.average_length = 50
    assert not prm.do_accum(0)
    assert not prm.do_accum(50)
    assert prm.do_accum(51)
    assert prm.do_accum(52)
    assert prm.do_accum(53)
    prm.sample_interval = 2
    assert not prm.do_accum(51)
    assert prm.do_accum(52)
    assert not prm.do_accum(53)
    assert prm.do_accum(100)


def test_cxx_params_known_answers():
    cxx = os.path.join(ROOT, "tests", "cxx")
    subprocess.check_call(["make", "-s", "-C", cxx])
    r = subprocess.run([os.path.join(cxx, "test_wrappers"), "selftest"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def test_component_names():
    """Item_jeh (fields_item_fields.hxx:24-28) and addKindSuffix over the moments' stems
    (fields_item.hxx:22-32, psc/moment.hxx:130-133, 281-286): kinds outermost"""
    assert ItemJeh.comp_names() == ["jx_ec", "jy_ec", "jz_ec", "ex_ec", "ey_ec", "ez_ec", "hx_fc", "hy_fc", "hz_fc"]

    class G:  # what comp_names reads
        kind_names = ["e", "i"]
    for which, first, n in ((MOMENT_ALL, ["rho_e", "jx_e"], 26), (MOMENT_N, ["n_e", "n_i"], 2), (MOMENT_RHO_NC, ["rho"], 1)):
        m = Moment.__new__(Moment)
        m.grid_, m.which = G, which
        names = m.comp_names()
        assert len(names) == n and names[:len(first)] == first
    m = Moment.__new__(Moment)
    m.grid_, m.which = G, MOMENT_ALL
    assert m.comp_names()[13:15] == ["rho_i", "jx_i"]
