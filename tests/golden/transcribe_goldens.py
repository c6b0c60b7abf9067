#!/usr/bin/env python
"""Transcribes the known-answer tests that psc-code/psc holds for the PIC hot path
into tests/golden/psc_golden.json.  The numbers are the reference's own test
constants; the few expected values its tests compute at run time (vx/vy/vz,
push_x, analytic currents) are evaluated here in double exactly as the test
source does.  Run from the repo root:  python tests/golden/transcribe_goldens.py

Sources (all under /root/reference/src/libpsc/tests/):
  test_push_particles.cxx:93-494   SingleParticlePushp1..16 (fixture testing.hxx:119-272)
  test_current_deposition.cxx:179-456  13 deposit cases, double, dx = 1, fnqs = 1
  test_collision_cuda.cxx:104-190  15-particle stable sort by cell
"""
import json
import math
import os

JXI, JYI, JZI, EX, EY, EZ, HX, HY, HZ = range(9)


def gamma_inv(u):
    return 1. / math.sqrt(1. + u[0] ** 2 + u[1] ** 2 + u[2] ** 2)


def make_push_cases():
    fnq = .05   # testing.hxx:133  (= fnqs/dt * dx = (1/200) * 10)
    dd = 10.    # testing.hxx:134
    cases = []

    def add(name, cite, fields, x0, u0, dim_rule, curr=None):
        """dim_rule(dim, prt0) -> (x1, u1) exactly as the TYPED_TEST body does"""
        c = dict(name=name, cite=cite, fields=fields,
                 prt0=dict(x=x0, u=u0, w=1., kind=0), expect={})
        for dim in ("xyz", "yz"):
            inv = (dim == "yz", False, False)
            x1, u1 = dim_rule(inv, list(x0), list(u0))
            c["expect"][dim] = dict(x=x1, u=u1, w=1., kind=0)
        if curr is not None:
            c["curr_ref"] = {}
            for dim in ("xyz", "yz"):
                inv = (dim == "yz", False, False)
                x1, u1 = dim_rule(inv, list(x0), list(u0))
                # push_x() (testing.hxx:254-266) returns the un-masked xi1
                g = gamma_inv(u1)
                xi1 = [x0[d] + g * u1[d] for d in range(3)]
                c["curr_ref"][dim] = curr(x0, xi1)
        cases.append(c)

    def vz_only(inv, x, u):  # prt1.x[2] += vz(prt1)
        x[2] += gamma_inv(u) * u[2]
        return x, u

    def push_x(inv, x, u):   # testing.hxx:254-266
        g = gamma_inv(u)
        for d in range(3):
            if not inv[d]:
                x[d] = x[d] + g * u[d]
        return x, u

    # :93-107
    add("Pushp1", "test_push_particles.cxx:93-107", {}, [5., 5., 5.], [0., 0., 1.], vz_only)

    # :114-130  EZ = 2
    def r2(inv, x, u):
        u[2] = 3.
        return vz_only(inv, x, u)
    add("Pushp2", "test_push_particles.cxx:114-130", {"EZ": ["const", 2.]},
        [5., 5., 5.], [0., 0., 1.], r2)

    # :140-156  EZ = z
    def r3(inv, x, u):
        u[2] = 6.
        return vz_only(inv, x, u)
    add("Pushp3", "test_push_particles.cxx:140-156", {"EZ": ["crd", 2]},
        [5., 5., 5.], [0., 0., 1.], r3)

    # :163-183  EZ = y
    def r4(inv, x, u):
        if not inv[1]:
            u[2] = 5.
        return push_x(inv, x, u)
    add("Pushp4", "test_push_particles.cxx:163-183", {"EZ": ["crd", 1]},
        [5., 4., 5.], [0., 0., 1.], r4)

    # :190-211  EZ = x
    def r5(inv, x, u):
        u[2] = 4.
        if inv[0]:
            u[2] = 1.
        return vz_only(inv, x, u)
    add("Pushp5", "test_push_particles.cxx:190-211", {"EZ": ["crd", 0]},
        [3., 5., 5.], [0., 0., 1.], r5)

    # :218-239
    add("Pushp6", "test_push_particles.cxx:218-239", {}, [1., 2., 3.], [1., 1., 1.], push_x)

    # :246-261  EZ = z, other block
    def r7(inv, x, u):
        u[2] = 156.
        return push_x(inv, x, u)
    add("Pushp7", "test_push_particles.cxx:246-261", {"EZ": ["crd", 2]},
        [151., 152., 155.], [1., 1., 1.], r7)

    # :262-282
    add("Pushp8", "test_push_particles.cxx:262-282", {}, [10., 10., 10.], [0., 0., 1.], push_x,
        lambda x0, xi1: [[JZI, [1, 1, 1], fnq / dd * (xi1[2] - x0[2])]])
    # :289-311
    add("Pushp9", "test_push_particles.cxx:289-311", {}, [10., 10., 19.5], [0., 0., 1.], push_x,
        lambda x0, xi1: [[JZI, [1, 1, 1], fnq / dd * (20. - x0[2])],
                         [JZI, [1, 1, 2], fnq / dd * (xi1[2] - 20.)]])
    # :318-342
    add("Pushp10", "test_push_particles.cxx:318-342", {}, [10., 19.5, 10.], [0., 1., 0.], push_x,
        lambda x0, xi1: [[JYI, [1, 1, 1], .05 / 10. * (20. - x0[1])],
                         [JYI, [1, 2, 1], .05 / 10. * (xi1[1] - 20.)]])
    # :349-369  (x move: deposits only when x is not invariant)
    add("Pushp11", "test_push_particles.cxx:349-369", {}, [10., 10., 10.], [1., 0., 0.], push_x,
        lambda x0, xi1: [[JXI, [1, 1, 1], fnq / dd * (xi1[0] - x0[0])]])
    # :376-399
    add("Pushp12", "test_push_particles.cxx:376-399", {}, [10., 10., 10.], [0., 1., 1.], push_x,
        lambda x0, xi1: [[JYI, [1, 1, 1], 0.00280342], [JYI, [1, 1, 2], 8.333333e-05],
                         [JZI, [1, 1, 1], 0.00280342], [JZI, [1, 2, 1], 8.333333e-05]])
    # :406-428
    add("Pushp13", "test_push_particles.cxx:406-428", {}, [10., 19.5, 10.], [0., 1., 1.], push_x,
        lambda x0, xi1: [[JYI, [1, 1, 1], 0.00243749], [JZI, [1, 1, 1], 6.25e-5],
                         [JYI, [1, 2, 1], 0.00036592], [JZI, [1, 2, 1], 0.00282275],
                         [JYI, [1, 1, 2], 6.25e-5], [JYI, [1, 2, 2], 2.08e-5]])
    # :435-457
    add("Pushp14", "test_push_particles.cxx:435-457", {}, [10., 10., 19.5], [0., 1., 1.], push_x,
        lambda x0, xi1: [[JZI, [1, 1, 1], 0.00243749], [JYI, [1, 1, 1], 6.25e-5],
                         [JZI, [1, 2, 1], 6.25e-5], [JZI, [1, 1, 2], 0.00036592],
                         [JYI, [1, 1, 2], 0.00282275], [JZI, [1, 2, 2], 2.08e-5]])
    # :464-477, :484-497
    add("Pushp15", "test_push_particles.cxx:464-477", {}, [5., 5., 39.5], [0., 0., 1.], push_x)
    add("Pushp16", "test_push_particles.cxx:484-497", {}, [5., 5., 159.5], [0., 0., 1.], push_x)
    return cases


def make_deposit_cases():
    Z = [[0.] * 4 for _ in range(4)]

    def arr(rows):  # reference arrays are [z][y]
        return [list(map(float, r)) for r in rows]

    cases = []

    def add(name, cite, xm, xp, vxi, jx=None, jy=None, jz=None, split_only=False):
        c = dict(name=name, cite=cite, xm=xm, xp=xp, vxi=vxi, split_only=split_only)
        if jy is not None or jx is not None or jz is not None:
            c["jxi_ref_zy"] = arr(jx or Z)
            c["jyi_ref_zy"] = arr(jy or Z)
            c["jzi_ref_zy"] = arr(jz or Z)
        cases.append(c)

    t = "test_current_deposition.cxx:"
    add("CurrentNotMoving", t + "179-202", [.5, 1., 1.], [.5, 1., 1.], [0., 0., 0.], Z, Z, Z)
    add("CurrentY", t + "204-227", [.5, 1., 1.], [.5, 1.2, 1.], [0., .2, 0.],
        jy=[[0, 0, 0, 0], [0, .2, 0, 0], [0, 0, 0, 0], [0, 0, 0, 0]])
    add("CurrentYShift", t + "229-252", [.5, 1., 1.7], [.5, 1.2, 1.7], [0., .2, 0.],
        jy=[[0, 0, 0, 0], [0, .06, 0, 0], [0, .14, 0, 0], [0, 0, 0, 0]])
    add("CurrentZ", t + "254-277", [.5, 1., 1.3], [.5, 1., 1.6], [0., 0., 3.],
        jz=[[0, 0, 0, 0], [0, .3, 0, 0], [0, 0, 0, 0], [0, 0, 0, 0]])
    add("CurrentYCross", t + "279-303", [.5, 1.9, 1.], [.5, 2.1, 1.], [0., .2, 0.],
        jy=[[0, 0, 0, 0], [0, .1, .1, 0], [0, 0, 0, 0], [0, 0, 0, 0]])
    add("CurrentYCrossShift", t + "305-329", [.5, 1.9, 1.3], [.5, 2.1, 1.3], [0., .2, 0.],
        jy=[[0, 0, 0, 0], [0, .07, .07, 0], [0, .03, .03, 0], [0, 0, 0, 0]])
    add("CurrentYZ", t + "331-355", [.5, 1.2, 1.1], [.5, 1.4, 1.4], [0., .2, .3],
        jy=[[0, 0, 0, 0], [0, .15, 0, 0], [0, .05, 0, 0], [0, 0, 0, 0]],
        jz=[[0, 0, 0, 0], [0, .21, .09, 0], [0, 0, 0, 0], [0, 0, 0, 0]])
    add("CurrentYZCrossShift", t + "357-381", [.5, 1.9, 1.3], [.5, 2.1, 1.4], [0., .2, 0.],
        jy=[[0, 0, 0, 0], [0, .0675, .0625, 0], [0, .0325, .0375, 0], [0, 0, 0, 0]],
        jz=[[0, 0, 0, 0], [0, .0025, .095, .0025], [0, 0, 0, 0], [0, 0, 0, 0]])
    add("CurrentYZCrossYZ", t + "383-412", [.5, 1.9, 1.6], [.5, 2.1, 2.2], [0., .2, 0.],
        jy=[[0, 0, 0, 0], [0, .025, 1. / 600., 0], [0, .075, .09 + 1. / 600, 0],
            [0, 0, 2. / 300., 0]],
        jz=[[0, 0, 0, 0], [0, .015, .38 + 1 / 300., 1. / 600.],
            [0, 0, .18 + 2. / 300., 4. / 300.], [0, 0, 0, 0]], split_only=True)
    # continuity-only cases (no reference arrays), also run in xyz
    add("CurrentX", t + "414-423", [1.3, 1., 1.], [1.6, 1., 1.], [.3, 0., 0.])
    add("CurrentXY", t + "425-434", [1.1, 1.2, 1.3], [1.4, 1.8, 1.3], [.3, .6, 0.])
    add("CurrentXYZ", t + "436-445", [1.1, 1.2, 1.3], [1.4, 1.6, 1.8], [.3, .4, .5])
    add("CurrentXYZCrossXYZ", t + "447-456", [1.9, 1.8, 1.7], [2.2, 2.4, 2.6], [.3, .6, .9])
    return cases


def make_sort_case():
    # test_collision_cuda.cxx:70-103 grid: 1x16x16 cells, L 160, 1x2x2 patches (dim_yz)
    inj = [
        [0, [5., 5., 5.], 0.], [0, [5., 5., 5.], .01],
        [0, [5., 15., 15.], .02], [0, [5., 15., 15.], .03], [0, [5., 15., 15.], .04],
        [0, [5., 15., 5.], .05], [0, [5., 15., 5.], .06], [0, [5., 15., 5.], .07],
        [0, [5., 15., 5.], .08],
        [1, [5., 105., 25.], .09], [1, [5., 105., 25.], .10],
        [1, [5., 115., 35.], .11], [1, [5., 115., 35.], .12],
        [1, [5., 115., 25.], .13], [1, [5., 115., 25.], .14],
    ]
    return dict(
        cite="test_collision_cuda.cxx:104-190",
        gdims=[1, 16, 16], length=[160., 160., 160.], np=[1, 2, 2],
        inject=[dict(patch=p, x=x, ux=ux) for p, x, ux in inj],
        # global cell index = patch * 64 + cell (cmprts n_cells = 256)
        idx_before=[0, 0, 9, 9, 9, 1, 1, 1, 1, 82, 82, 91, 91, 83, 83],
        idx_after=[0, 0, 1, 1, 1, 1, 9, 9, 9, 82, 82, 83, 83, 91, 91],
        id_after=[0, 1, 5, 6, 7, 8, 2, 3, 4, 9, 10, 13, 14, 11, 12],
    )


def main():
    out = dict(
        note="golden vectors transcribed from psc-code/psc's own tests by "
             "tests/golden/transcribe_goldens.py",
        push_fixture=dict(cite="testing.hxx:119-170", gdims=[16, 16, 16], L=160., dt=1.,
                          kinds=[[1., 1.]], nicell=200, eps=1e-5),
        push_cases=make_push_cases(),
        deposit_fixture=dict(cite="test_current_deposition.cxx:56-101", gdims=[4, 4, 4],
                             dt=1., fnqs=1., qni_wni=1.),
        deposit_cases=make_deposit_cases(),
        sort_case=make_sort_case(),
    )
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "psc_golden.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
