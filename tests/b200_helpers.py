"""helpers shared by the tests that exercise the product (host logic on CPU, CUDA
path on the GPU box)."""
import ctypes as C
import os
import subprocess

import numpy as np

import oracle_lib as ol

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def desc_from_grid(grid, rank=0, n_ranks=1, n_patches_by_rank=None, max_n_prts=0):
    """psc_b200_grid_desc for the same grid the oracle uses"""
    from psc_b200._lib import GridDesc, i3, d3
    g = grid.g
    d = GridDesc()
    d.gdims = i3(*g.gdims)
    d.np = i3(*g.np)
    d.length = d3(*g.length)
    d.corner = d3(*g.corner)
    d.dt, d.fnqs, d.eta = g.dt, g.fnqs, g.eta
    d.n_kinds = g.n_kinds
    for k in range(g.n_kinds):
        d.q[k] = g.q[k]
        d.m[k] = g.m[k]
    d.bc_fld_lo = i3(*g.bc_fld_lo)
    d.bc_fld_hi = i3(*g.bc_fld_hi)
    d.bc_prt_lo = i3(*g.bc_prt_lo)
    d.bc_prt_hi = i3(*g.bc_prt_hi)
    d.deposit = g.deposit
    d.rank, d.n_ranks = rank, n_ranks
    if n_patches_by_rank is not None:
        arr = (C.c_int * n_ranks)(*n_patches_by_rank)
        d.n_patches_by_rank = arr
        d._keep = arr
    d.device = -1
    d.max_n_prts = max_n_prts
    return d


_hc = None


def hostcheck():
    """host build of the product's per-particle math (tests/hostcheck/), CPU only"""
    global _hc
    if _hc is None:
        src = os.path.join(ROOT, "tests", "hostcheck", "pic_math_host.cpp")
        so = os.path.join(ROOT, "tests", "hostcheck", "libpic_math_host.so")
        deps = [src, os.path.join(ROOT, "psc_b200", "csrc", "pic_math.cuh"),
                os.path.join(ROOT, "psc_b200", "csrc", "grid.hpp"),
                os.path.join(ROOT, "include", "psc_b200.h")]
        if not os.path.exists(so) or os.path.getmtime(so) < max(map(os.path.getmtime, deps)):
            subprocess.check_call(["g++", "-std=c++17", "-O3", "-ffp-contract=off", "-fPIC",
                                   "-shared", src, "-o", so])
        L = C.CDLL(so)
        P = C.c_void_p
        L.hc_push_mprts.argtypes = [P, P, P, P]
        L.hc_bnd_classify.argtypes = [P, P, P, P, P]
        L.hc_grid_info.argtypes = [P, P, P, P, P]
        L.hc_neighbor_table.argtypes = [P, P, P]
        _hc = L
    return _hc


def make_gpu_grid(og, **kw):
    """psc_b200.Grid (device context) for the same grid the oracle uses"""
    import psc_b200 as pb
    g = og.g
    return pb.Grid(gdims=tuple(g.gdims), length=tuple(g.length), np=tuple(g.np), dt=g.dt,
                   kinds=og.kinds, fnqs=g.fnqs, eta=g.eta, corner=tuple(g.corner),
                   bc_fld_lo=list(g.bc_fld_lo), bc_fld_hi=list(g.bc_fld_hi),
                   bc_prt_lo=list(g.bc_prt_lo), bc_prt_hi=list(g.bc_prt_hi),
                   deposit=g.deposit, **kw)


def gpu_state(og, flds, prts, off, options=None):
    """device context loaded with the given fields and particles"""
    import psc_b200 as pb
    grid = make_gpu_grid(og)
    for k, v in (options or {}).items():
        grid.set_option(k, v)
    mprts, mflds = pb.Mparticles(grid), pb.MfieldsState(grid)
    if prts is not None:
        mprts.set(prts, np.diff(off))
    if flds is not None:
        mflds.upload(flds)
    return grid, mprts, mflds


def gpu_push(options=None, sort_first=False, expect_lean=None):
    """push(grid, flds, prts, off) backend running the CUDA path through the C ABI;
    expect_lean: assert that k_push_lean did (True) / did not (False) take the push"""
    import psc_b200 as pb

    def push(og, flds, prts, off):
        grid, mprts, mflds = gpu_state(og, flds, prts, off, options)
        if sort_first:
            pb.Sort()(mprts)
        pb.PushParticles().push_mprts(mprts, mflds)
        if expect_lean is not None:
            assert (grid.get_stat("lean_pushes") >= 1) == expect_lean, "k_push_lean %s" % (
                "did not run" if expect_lean else "ran")
        got, got_off = mprts.get()
        assert np.array_equal(got_off, off)
        prts[:] = got
        flds[:] = mflds.download()
        grid.close()
    return push
