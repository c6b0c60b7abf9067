"""Worker of tests/test_dist_gloo.py: one process per rank, gloo backend, CPU only.

Checks the host-side multi-rank logic the NCCL exchanges are planned from, the way the
ranks of a real run would see it:
  1. every rank derives the same decomposition; the ranges tile the global patch list
  2. neighbour tables are mutually consistent: if my patch P sees patch Q of rank r in
     direction d, rank r's Q sees P in direction -d -- both sides enumerate (patch
     ascending, direction ascending), which is what lets the exchange go without headers;
     the per-peer message counts computed locally match what the peer computes (all_to_all)
  3. best_mapping is deterministic: every rank computes the same new patch counts from
     all-gathered loads (psc_balance_impl.hxx:99-160), and they match the oracle's mapping
"""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle_lib as ol  # noqa: E402
from b200_helpers import desc_from_grid, hostcheck  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    hc = hostcheck() if rank == 0 else None
    dist.barrier()  # rank 0 builds the helper library once
    hc = hostcheck()
    for gd, np3 in (((16, 16, 32), (2, 2, 4)), ((1, 32, 48), (1, 2, 3)), ((8, 8, 24), (1, 1, 3))):
        og = ol.Grid(gdims=gd, length=tuple(float(g) for g in gd), np_=np3, dt=0.3,
                     kinds=((-1., 1.), (1., 100.)), nicell=4)
        n_global = og.n_patches
        d = desc_from_grid(og, rank=rank, n_ranks=world)
        n_p, p_begin, ld, nei0 = C.c_int(), C.c_int(), (C.c_int * 3)(), (C.c_int * 27)()
        assert hc.hc_grid_info(C.byref(d), C.byref(n_p), C.byref(p_begin), ld, nei0) == 0
        # 1. ranges tile the patch list
        mine = torch.tensor([p_begin.value, n_p.value], dtype=torch.int64)
        allr = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(allr, mine)
        pos = 0
        for r in range(world):
            assert int(allr[r][0]) == pos and int(allr[r][1]) >= 1
            pos += int(allr[r][1])
        assert pos == n_global
        owner = np.concatenate([np.full(int(a[1]), r) for r, a in enumerate(allr)])
        # 2. neighbour tables
        n_loc = n_p.value
        nei_gp = np.zeros(n_loc * 27, dtype=np.int32)
        nei_rk = np.zeros(n_loc * 27, dtype=np.int32)
        assert hc.hc_neighbor_table(C.byref(d), ol.ptr(nei_gp), ol.ptr(nei_rk)) == n_loc
        ok = nei_gp >= 0
        assert np.array_equal(nei_rk[ok], owner[nei_gp[ok]])
        full = [None] * world
        dist.all_gather_object(full, (p_begin.value, nei_gp.reshape(n_loc, 27)))
        table = np.concatenate([t for _, t in sorted(full, key=lambda x: x[0])])
        for gp in range(n_global):
            for di in range(27):
                q = table[gp, di]
                if q >= 0:
                    assert table[q, 26 - di] == gp, (gp, di, q)
        # per-peer message counts: (patch, dir) pairs whose neighbour lives on another rank
        send = np.zeros(world, dtype=np.int64)
        for k in range(n_loc * 27):
            if k % 27 != 13 and nei_gp[k] >= 0 and nei_rk[k] != rank:
                send[nei_rk[k]] += 1
        recv = torch.zeros(world, dtype=torch.int64)
        outs = list(torch.tensor(send).split(1))
        ins = list(recv.split(1))
        # gloo has no all_to_all on CPU tensors in every build: emulate with all_gather
        allsend = [torch.zeros(world, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(allsend, torch.tensor(send))
        for r in range(world):
            # what r sends me must equal what I expect from r = what I send r (symmetry of
            # the neighbour relation)
            assert int(allsend[r][rank]) == int(send[r]), (r, allsend[r], send)
        # 3. balancer mapping
        rng = np.random.default_rng(5)
        loads_global = rng.uniform(1., 10., size=n_global)
        my_loads = torch.tensor(loads_global[p_begin.value:p_begin.value + n_loc])
        gl = [None] * world
        dist.all_gather_object(gl, my_loads.numpy())
        loads = np.concatenate(gl)
        assert np.array_equal(loads, loads_global)
        import psc_b200
        L = psc_b200.load()
        cap = np.ones(world)
        out = np.zeros(world, dtype=np.int32)
        assert L.psc_b200_best_mapping(world, ol.ptr(cap), n_global, ol.ptr(loads), ol.ptr(out)) == 0
        res = [None] * world
        dist.all_gather_object(res, out.tolist())
        assert all(r == res[0] for r in res)
        assert sum(res[0]) == n_global and min(res[0]) >= 1
        ref = np.zeros(world, dtype=np.int32)
        ol.lib().po_best_mapping(world, ol.ptr(cap), n_global, ol.ptr(loads), ol.ptr(ref))
        assert out.tolist() == ref.tolist(), (out, ref)
    dist.barrier()
    if rank == 0:
        print("dist_gloo ok")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
