#!/usr/bin/env python
"""Deck-scale runs of BASELINE.json configs[0..3] on the device: the reference decks' DEFAULT
grids (tests/decks.py restates their profiles, kinds, boundary conditions and cadence), with the
deck's collisions / heating where it has them, for --steps steps.  One JSON line per deck:
particle-steps/s, ms per step, total-energy drift, continuity residual of a checked step,
particle count before / after.

  python tools/deck_scale.py --deck bubble_yz --steps 1000
  python -m torch.distributed.run --nproc-per-node 4 ... tools/deck_scale.py --deck harris_yz

Multi-GPU: one process per GPU, patches split evenly (PSC's rank ranges); every rank builds
the whole initial state on the host and keeps its own patches."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

# the decks' own grids (file:line of the reference)
SCALE = {
    # psc_bubble_yz.cxx:119-146: 1 x 1024 x 1536 cells in 32 x 48 patches, nicell 100; collisions every 10 (nu .1, :303-306)
    "bubble_yz": dict(kw=dict(gdims=(1, 1024, 1536), np_=(1, 32, 48), nicell=100), collision=(10, .1)),
    # psc_flatfoil_yz.cxx:299-302 (CASE_2D): 1 x 1600 x 4800 cells in 50 x 150 patches, nicell 100;
    # collisions every 10 (:464-467), heating every 20 (:556-570)
    "flatfoil_yz": dict(kw=dict(gdims=(1, 1600, 4800), np_=(1, 50, 150), length=(1., 800., 2400.), nicell=100),
                        collision=(10, 3.76 * 0.001 ** 2 / 2.5 / 20.), heating=20),
    # psc_harris_yz.cxx:226-235: 1 x 128 x 512 cells, nicell 100; split into 4 x 16 patches so that 8 GPUs get work
    "harris_yz": dict(kw=dict(gdims=(1, 128, 512), np_=(1, 4, 16), nicell=100)),
    # psc_kelvin_helmholtz.cxx (3D variant of BASELINE configs[3]): 128^3 cells in 4^3 patches, nicell 50, four kinds
    "kelvin_helmholtz_xyz": dict(kw=dict(gdims=(128, 128, 128), np_=(4, 4, 4), length=(64., 64., 64.), nicell=50)),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--deck", required=True, choices=list(SCALE))
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--physics", type=int, default=1, help="0: no collisions / heating (the parity runs' setting)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    import decks
    import psc_b200 as pb
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    cfg = SCALE[args.deck]
    t0 = time.time()
    st = decks.DECKS[args.deck](**cfg["kw"])
    og, g = st["og"], st["og"].g
    t_build = time.time() - t0
    npg = og.n_patches
    per, rem = divmod(npg, world)
    n_by_rank = [per + (r < rem) for r in range(world)]
    p0 = sum(n_by_rank[:rank])
    p1 = p0 + n_by_rank[rank]
    grid = pb.Grid(gdims=tuple(g.gdims), length=tuple(g.length), np=tuple(g.np), dt=g.dt, kinds=og.kinds,
                   fnqs=g.fnqs, eta=g.eta, corner=tuple(g.corner), bc_fld_lo=list(g.bc_fld_lo),
                   bc_fld_hi=list(g.bc_fld_hi), bc_prt_lo=list(g.bc_prt_lo), bc_prt_hi=list(g.bc_prt_hi),
                   deposit=g.deposit, rank=rank, n_ranks=world, n_patches_by_rank=n_by_rank, device=local_rank)
    grid.cori = 1. / cfg["kw"]["nicell"]
    if world > 1:
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(pb.Grid.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        grid.nccl_init(bytes(idt.cpu().numpy().tobytes()))
    off = st["off"]
    mprts, mflds = pb.Mparticles(grid), pb.MfieldsState(grid)
    mprts.set(st["prts"][off[p0]:off[p1]], np.diff(off[p0:p1 + 1]))
    mflds.upload(st["flds"][p0:p1])
    n_total = int(off[-1])
    del st

    coll = heat = None
    if args.physics and "collision" in cfg:
        coll = pb.Collision(grid, cfg["collision"][0], cfg["collision"][1], seed=1)
    if args.physics and "heating" in cfg:
        # psc_flatfoil_yz.cxx:556-570: the heating spot over the foil, electrons only
        d_i = 10.
        heat = pb.Heating(grid, cfg["heating"], dict(zl=-10. * d_i, zh=10. * d_i, xc=0., yc=0., rH=3. * d_i,
                                                     T=[.04, .04, 0.], Mi=100.), seed=2)
    psc = pb.Psc(grid, mflds, mprts, sort_interval=st_sort(args.deck), marder_interval=0, fused=True, collision=coll)
    psc.initialize()

    def allsum(v):
        if not dist:
            return np.asarray(v, dtype=np.float64)
        t = torch.tensor(np.asarray(v, dtype=np.float64), device="cuda")
        dist.all_reduce(t)
        return t.cpu().numpy()

    e0 = pb.energies(grid)  # (already summed over the ranks by the library)
    grid.set_option("profile", 1)
    grid.profile_reset()
    grid.sync()
    if dist:
        dist.barrier()
    grid.timer_start()
    for n in range(args.steps):
        psc.step()
        if heat is not None and grid.timestep % heat.interval_ == 0:
            heat(mprts)
    ms = grid.timer_stop()
    if dist:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    e1 = pb.energies(grid)
    prof = {k: round(v[0] / args.steps, 3) for k, v in grid.profile().items()}
    grid.set_option("profile", 0)
    n_after = int(allsum([mprts.size()])[0])
    # one checked step: continuity residual (psc.hxx:379-384,471-476)
    chk = pb.Psc(grid, mflds, mprts, sort_interval=1, fused=True,
                 checks=pb.Checks(grid, continuity_interval=1, gauss_interval=0))
    chk.step()
    cont = chk.checks.continuity.last_max_err
    if rank == 0:
        print(json.dumps({
            "deck": args.deck, "grid": list(g.gdims), "patches": list(g.np), "nicell": cfg["kw"]["nicell"],
            "n_gpus": world, "particles": n_total, "particles_after": n_after, "steps": args.steps,
            "physics": {"collisions": bool(coll), "heating": bool(heat)},
            "ms_per_step": ms / args.steps, "particle_steps_per_s": n_total * args.steps / (ms * 1e-3),
            "energy_total_start": float(e0.sum()), "energy_total_end": float(e1.sum()),
            "energy_drift": float(e1.sum() / e0.sum() - 1.), "field_energy_start": float(e0[:6].sum()),
            "field_energy_end": float(e1[:6].sum()), "continuity_max_err": cont,
            "fused_steps": grid.get_stat("fused_steps"), "fused_fallbacks": grid.get_stat("fused_fallbacks"),
            "kernels_ms_per_step": prof, "host_build_s": round(t_build, 1)}), flush=True)
    grid.close()
    if dist:
        dist.destroy_process_group()


def st_sort(deck):
    return 10  # PscParams::sort_interval of every deck (keep_sorted keeps the store ordered in between)


if __name__ == "__main__":
    main()
