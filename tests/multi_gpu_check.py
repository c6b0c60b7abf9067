"""Multi-GPU parity check, one process per GPU (NCCL):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29533 tests/multi_gpu_check.py

Every rank owns a contiguous range of the global patch list.  Rank 0 gathers the other
ranks' particles and fields and compares them with the CPU oracle run on the whole
domain with the same rank map (remote arrivals are appended behind local ones:
ddc_particles.hxx:456-468).  Run by tests/test_gpu_multi.py when >= 2 GPUs are visible
and by hand under `gpurun --gpus 2`."""
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
import psc_b200 as pb  # noqa: E402
from gen import random_fields, thermal_plasma  # noqa: E402

KINDS = ((-1., 1.), (1., 100.))
# (xyz patches are 16 x 8 x 8 cells: multiples of k_push_lean's 16 x 4 x 4 tile, so that the multi-rank
# cases run the headline kernel and its listing of the remote leavers)
CASES = {
    "xyz_periodic_slabs": dict(gdims=(32, 16, 32), length=(32., 16., 32.), np_=(2, 2, 4)),
    "yz_periodic": dict(gdims=(1, 32, 64), length=(1., 32., 64.), np_=(1, 2, 4)),
    "xyz_wall_z": dict(gdims=(32, 16, 32), length=(32., 16., 32.), np_=(2, 2, 4),
                       bc_fld_lo=[1, 1, 2], bc_fld_hi=[1, 1, 2], bc_prt_lo=[1, 1, 0], bc_prt_hi=[1, 1, 0]),
}


def bench_parity_cases(rank, world, local_rank):
    """the cases bench.py runs ahead of its timed region at world > 1 (process group already up):
    3 fused steps on periodic slabs and on a walled box, and an uneven decomposition that is
    rebalanced after the first step.  Returns (all ok, [per-case records]) on rank 0."""
    out, ok = [], True
    npg = 16
    small = 3 if world <= 4 else 1
    uneven = [npg - small * (world - 1)] + [small] * (world - 1)
    todo = [("xyz_periodic_slabs", dict(fused=True, n_steps=1)), ("xyz_periodic_slabs", dict(fused=True)),
            ("xyz_wall_z", dict(fused=True)), ("yz_periodic", dict(fused=True, n_steps=1))]
    todo.append(("xyz_periodic_slabs", dict(fused=True, pipelined=True)))
    if uneven:
        todo.append(("xyz_periodic_slabs", dict(fused=True, n_steps=4, n_by_rank=uneven, balance_step=0)))
    for name, kw in todo:
        if CASES[name]["np_"][2] * CASES[name]["np_"][1] * CASES[name]["np_"][0] < world:
            continue
        o, info = run_case(name, CASES[name], rank, world, local_rank, want_info=True, **kw)
        ok = ok and o
        if info:
            out.append(info)
    return ok, out


def gather_obj(obj, rank, world):
    out = [None] * world if rank == 0 else None
    dist.gather_object(obj, out, dst=0)
    return out


# one rank takes the general exchange + sort while its peers stay on the fused pass (a particle
# that moves two cells breaks the fused pass's precondition on the rank that owns it): dz = 0.25 <
# v dt.  Both paths issue exactly one particle exchange per call over the same neighbour tables,
# which is what keeps the ranks in step (fused_sort.cu, fused_bnd_sort).
FALLBACK_CASE = dict(gdims=(32, 16, 32), length=(32., 16., 8.), np_=(2, 2, 4))


def run_case(name, gkw, rank, world, local_rank, fused, n_steps=3, n_by_rank=None, balance_step=None,
             want_info=False, vth=(0.5, 0.05), hot_particle=False, pipelined=False):
    og = ol.Grid(dt=0.35, kinds=KINDS, nicell=8, **gkw)
    npg = og.n_patches
    if n_by_rank is None:
        per, remn = divmod(npg, world)
        n_by_rank = [per + (r < remn) for r in range(world)]
    starts = np.concatenate([[0], np.cumsum(n_by_rank)])
    rank_of_patch = np.repeat(np.arange(world), n_by_rank).astype(np.int32)
    flds = random_fields(og, seed=3, amp_e=0.02, amp_b=0.05)
    ol.fill_ghosts(og, flds, 3, 9)
    prts, off = thermal_plasma(og, ppc=6, seed=4, vth=vth, margin=0.02)
    if hot_particle:
        # patch 0 (rank 0): from cell z = 2 to cell z = 4 in one step
        prts["x"][off[0] + 5] = (3.5, 3.5, 0.74)
        prts["u"][off[0] + 5] = (0., 0., 10.)

    g = og.g
    grid = pb.Grid(gdims=tuple(g.gdims), length=tuple(g.length), np=tuple(g.np), dt=g.dt, kinds=og.kinds,
                   fnqs=g.fnqs, eta=g.eta, bc_fld_lo=list(g.bc_fld_lo), bc_fld_hi=list(g.bc_fld_hi),
                   bc_prt_lo=list(g.bc_prt_lo), bc_prt_hi=list(g.bc_prt_hi), rank=rank, n_ranks=world,
                   n_patches_by_rank=n_by_rank, device=local_rank)
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(pb.Grid.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    grid.nccl_init(bytes(idt.cpu().numpy().tobytes()))
    p0, p1 = starts[rank], starts[rank + 1]
    assert grid.n_patches() == p1 - p0 and grid.patch_begin() == p0
    mprts, mflds = pb.Mparticles(grid), pb.MfieldsState(grid)
    mprts.set(prts[off[p0]:off[p1]], np.diff(off[p0:p1 + 1]))
    mflds.upload(flds[p0:p1])
    psc = pb.Psc(grid, mflds, mprts, sort_interval=1, fused=fused)

    # reference: the oracle's step with the same rank map
    rf, rp, ro = flds.copy(), prts.copy(), off.copy()
    L, G = ol.lib(), og.byref()
    for _ in range(n_steps):
        L.po_sort(G, ol.ptr(rp), ol.ptr(ro), None)
        L.po_push_mprts(G, ol.ptr(rf), ol.ptr(rp), ol.ptr(ro))
        rp, ro, _ = ol.bnd_particles(og, rp, ro, rank_of_patch=rank_of_patch)
        L.po_bndf_add_ghosts_J(G, ol.ptr(rf))
        L.po_add_ghosts(G, ol.ptr(rf), 9, 0, 3)
        L.po_fill_ghosts(G, ol.ptr(rf), 9, 0, 3)
        L.po_push_H(G, ol.ptr(rf), .5)
        L.po_bndf_fill_ghosts_H(G, ol.ptr(rf))
        L.po_fill_ghosts(G, ol.ptr(rf), 9, 6, 9)
        L.po_push_E(G, ol.ptr(rf), 1.)
        L.po_bndf_fill_ghosts_E(G, ol.ptr(rf))
        L.po_fill_ghosts(G, ol.ptr(rf), 9, 3, 6)
        L.po_push_H(G, ol.ptr(rf), .5)
        L.po_bndf_fill_ghosts_H(G, ol.ptr(rf))
        L.po_fill_ghosts(G, ol.ptr(rf), 9, 6, 9)
        if pipelined:
            # the step through psc_b200_step_begin / _step_end (pipelined host I/O): J comes down
            # while the exchange's scatter is still running, DiagEnergies are reduced inside
            prm = pb.StepParams(sort=1, marder_loop=0, marder_diffusion=0., push_fields=1, checks=0, energies=1)
            h_j = np.zeros(mflds.shape(3), dtype=np.float32)
            pb.check(grid.lib.psc_b200_step_begin(grid.ctx, C.byref(prm)))
            pb.check(grid.lib.psc_b200_mflds_download_async(grid.ctx, 0, pb.JXI, pb.JXI + 3,
                                                           h_j.ctypes.data_as(C.c_void_p)))
            pb.check(grid.lib.psc_b200_io_wait(grid.ctx))
            pb.check(grid.lib.psc_b200_step_end(grid.ctx))
            en_in_step = np.zeros(8)
            pb.check(grid.lib.psc_b200_last_energies(grid.ctx, en_in_step.ctypes.data_as(C.c_void_p)))
        else:
            psc.step()
        if balance_step is not None and _ == balance_step:
            # Balance::operator(): loads -> best_mapping -> whole patches move between GPUs
            loads = (np.diff(ro) + 1.0 * og.n_cells).astype(np.float64)
            new_n = np.zeros(world, dtype=np.int32)
            cap = np.ones(world)
            L.po_best_mapping(world, ol.ptr(cap), npg, ol.ptr(loads), ol.ptr(new_n))
            changed = grid.balance(1.0)
            assert changed == (list(new_n) != list(n_by_rank)), (changed, new_n, n_by_rank)
            assert grid.n_patches() == new_n[rank], (grid.n_patches(), new_n)
            assert grid.patch_begin() == int(np.sum(new_n[:rank]))
            n_by_rank = [int(x) for x in new_n]
            rank_of_patch = np.repeat(np.arange(world), n_by_rank).astype(np.int32)
            # the operator types re-attach to the new decomposition (psc_balance_generation_cnt)
            mprts, mflds = pb.Mparticles(grid), pb.MfieldsState(grid)
            psc = pb.Psc(grid, mflds, mprts, sort_interval=1, fused=fused)
    # the fused step already did the next sort, and so did the separate operators (keep_sorted:
    # the exchange after a sorted push is the fused exchange + sort); Sort is then a no-op
    L.po_sort(G, ol.ptr(rp), ol.ptr(ro), None)
    pb.Sort()(mprts)
    gp, go = mprts.get()
    gf = mflds.download()
    en = pb.energies(grid)
    if pipelined:
        # J as it came down (interior + ghosts of this rank's patches) against the device's copy,
        # and the energies of the last step as reduced inside it
        j_dev = mflds.download(pb.JXI, pb.JXI + 3)
        assert np.array_equal(h_j, j_dev), "J downloaded inside the step differs from the device's"
    stats = dict(fused=grid.get_stat("fused_steps"), fallbacks=grid.get_stat("fused_fallbacks"),
                 lean=grid.get_stat("lean_pushes"))
    res = gather_obj((gp, go, gf, stats, en_in_step if pipelined else None), rank, world)
    ok = True
    info = None
    if rank == 0 and pipelined:
        # (psc_b200_last_energies is this rank's share; psc_b200_energies reduces over the ranks)
        en_sum = np.sum([r[4] for r in res], axis=0)
        ok = ok and bool(np.allclose(en_sum, en, rtol=1e-5, atol=1e-30))
    if rank == 0:
        counts = np.concatenate([np.diff(r[1]) for r in res])
        ref_counts = np.diff(ro)
        all_p = np.concatenate([r[0] for r in res])
        all_f = np.concatenate([r[2] for r in res])
        same_counts = np.array_equal(counts, ref_counts)
        ferr = np.abs(all_f - rf).max() / np.abs(rf).max()
        perr = uerr = float("nan")
        if same_counts:
            perr = float(np.abs(all_p["x"] - rp["x"]).max())
            uerr = float(np.abs(all_p["u"] - rp["u"]).max())
        ref_e = ol.energies(og, rf, rp, ro)
        eerr = float(np.abs(en - ref_e).max() / np.abs(ref_e).max())
        ok = ok and same_counts and ferr < 3e-5 and perr < 1e-4 and uerr < 1e-5 and eerr < 1e-4
        # per-cell counts and the migration order are exact.  After ONE step the particle records
        # are byte-identical to the oracle's (same fields in, same arithmetic, same order); over
        # several steps x/u carry the round-off of J's summation order through E
        bytes_equal = bool(same_counts and all_p.tobytes() == rp.tobytes())
        if n_steps == 1:
            ok = ok and bytes_equal
        if hot_particle and fused:
            ok = ok and res[0][3]["fallbacks"] >= 1 and all(r[3]["fallbacks"] == 0 for r in res[1:])
        info = dict(name=name, fused=int(fused), steps=n_steps, ranks=world, counts_equal=bool(same_counts),
                    particles_byte_exact=bytes_equal, fld_rel=float(ferr), x_abs=perr, u_abs=uerr,
                    energies_rel=eerr, balanced=balance_step is not None, pipelined=bool(pipelined), ok=bool(ok))
        print("%-28s fused=%d  counts %s  fld rel %.2e  x abs %.2e  u abs %.2e  energies rel %.2e  %s  %s" % (
            name, fused, "same" if same_counts else "DIFFER", ferr, perr, uerr, eerr,
            [r[3] for r in res], "ok" if ok else "FAIL"), flush=True)
    grid.close()
    if want_info:
        return ok, info
    return ok


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    cases = CASES
    ok = True
    for name, kw in cases.items():
        for fused in (False, True):
            ok = run_case(name, kw, rank, world, local_rank, fused) and ok
            ok = run_case(name + "_1step", kw, rank, world, local_rank, fused, n_steps=1) and ok
    # the pipelined step (step_begin / step_end: the scatter behind the exchange is deferred)
    for name in ("xyz_periodic_slabs", "xyz_wall_z", "yz_periodic"):
        ok = run_case(name + "_pipelined", cases[name], rank, world, local_rank, True, pipelined=True) and ok
        ok = run_case(name + "_pipelined_1step", cases[name], rank, world, local_rank, True, n_steps=1,
                      pipelined=True) and ok
    # uneven patch distribution (what the balancer produces)
    npg = 16
    small = 3 if world <= 4 else 1
    uneven = [npg - small * (world - 1)] + [small] * (world - 1)
    ok = run_case("xyz_uneven_ranks", cases["xyz_periodic_slabs"], rank, world, local_rank, True,
                  n_by_rank=uneven) and ok
    # one rank on the general path, the others on the fused pass
    ok = run_case("xyz_one_rank_falls_back", FALLBACK_CASE, rank, world, local_rank, True, n_steps=2,
                  vth=(0.05, 0.005), hot_particle=True) and ok
    # load balancing: start uneven, rebalance after the first step, keep stepping
    for fused in (False, True):
        ok = run_case("xyz_balance_after_step1", cases["xyz_periodic_slabs"], rank, world, local_rank,
                      fused, n_steps=4, n_by_rank=uneven, balance_step=0) and ok
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_GPU_CHECK", "PASS" if flag.item() else "FAIL", flush=True)
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
