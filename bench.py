#!/usr/bin/env python
"""bench.py -- particle-steps/s of the PIC hot path (push + deposit + boundary exchange +
sort, plus the field half of Psc::step) on synthetic 3D thermal plasma (SURVEY.md 8d,
BASELINE.json configs[4]): per GPU 256^3 cells x 64 particles per cell in 8x8x8 patches of
32^3, periodic, uniform B_z; ranks stack their slabs along z (weak scaling).

  python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo (CUDA, C ABI)
  python bench.py --impl reference ...                           # CPU arm: the reference's
        1vb path on the box's host cores (oracle/_ref not needed: the plain-C port of
        oracle/ is what runs; see DESIGN.md "CPU baseline")

One JSON line on rank 0.  `value` times K steps with the state resident in HBM (CUDA
events on the context's stream, max over ranks); `e2e` is the same step through the
operator-level C ABI with HOST buffers in the loop every step (E/B uploaded from pinned
host memory, deposited J and energies read back); `roofline` is the dominant kernel
against the measured HBM copy bandwidth; `cpu_baseline` is the oracle port on host cores.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

B_PUSH = 64.0    # algorithmic bytes per particle of push+deposit: read 32 B, write 32 B
B_SORT = 72.0    # key write+read 8 B, particle read+write 64 B
B_STEP = B_PUSH + B_SORT
KINDS = ((-1., 1.), (1., 100.))
VTH = (0.05, 0.005)


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            d = json.load(f)
        for k in ("hbm_gbs", "hbm_gb_s", "hbm_GBs"):
            if k in d:
                return float(d[k]), "measured (MEASURED_PEAKS.json)"
        for v in d.values():
            if isinstance(v, dict) and "hbm_gbs" in v:
                return float(v["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons during the timed region"""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.stop_ev, self.th, self.proc = index, [], threading.Event(), None, None

    def _run(self):
        # one streaming nvidia-smi (a sample every 100 ms) instead of one process per sample;
        # falls back to polling if the loop mode is not available
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                line = line.strip()
                if line:
                    self.rows.append([x.strip() for x in line.split(",")])
                if self.stop_ev.is_set():
                    break
            if self.rows or self.stop_ev.is_set():
                return
        except Exception:
            pass
        while not self.stop_ev.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True,
                                     text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_ev.wait(0.2)

    def start(self):
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()
        time.sleep(0.15)  # let the stream deliver its first sample before the timed region

    def stop(self):
        self.stop_ev.set()
        if self.proc is not None:
            try:
                self.proc.terminate()
            except Exception:
                pass
        if self.th:
            self.th.join(timeout=6)
        sm = [int(r[0]) for r in self.rows if r and r[0].isdigit()]
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows for i in range(4)
                          if len(r) > 2 + i and r[2 + i].lower().startswith("active")})
        return {"sm_mhz": int(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


# ---------------------------------------------------------------------------- CPU arm

def cpu_arm(cells, ppc, n_patches_target, steps, warmup, threads):
    """the reference's CPU 1vb path (oracle port: push_particles_1vb.hxx + psc_sort_impl.hxx
    + bnd_particles_impl.hxx restated in oracle/psc_oracle.c), patches spread over host
    threads the way PSC spreads them over MPI ranks"""
    # torchrun exports OMP_NUM_THREADS=1 to its workers: the oracle's OpenMP loops (boundary
    # exchange) must see the cores this process may use, whatever the launcher set
    os.environ["OMP_NUM_THREADS"] = str(threads)
    import oracle_lib as ol
    from gen import thermal_plasma
    from concurrent.futures import ThreadPoolExecutor
    try:
        C.CDLL("libgomp.so.1", mode=C.RTLD_GLOBAL).omp_set_num_threads(int(threads))
    except OSError:
        pass
    pz = max(1, n_patches_target // 4)
    og = ol.Grid(gdims=(cells * 2, cells * 2, cells * pz), length=(2. * cells, 2. * cells, 1. * cells * pz),
                 np_=(2, 2, pz), dt=0.75 / np.sqrt(3.), kinds=KINDS, nicell=ppc // 2)
    flds = og.zeros_fields()
    flds[:, ol.HZ] = 0.1
    prts, off = thermal_plasma(og, ppc=ppc // 2, seed=1234, vth=VTH, shuffle=False)
    n = len(prts)
    L, G = ol.lib(), og.byref()
    npch = og.n_patches
    threads = max(1, min(threads, npch))
    bounds = [(t * npch // threads, (t + 1) * npch // threads) for t in range(threads)]
    pool = ThreadPoolExecutor(threads)

    # the exchange writes into a second, preallocated store (PSC reuses its buffers too);
    # the domain is periodic, so the particle number does not change
    spare = [np.zeros_like(prts), np.zeros_like(off)]
    nd = np.zeros(1, dtype=np.uint32)

    # the push + deposit is the reference's own code (oracle/_ref: push_particles_1vb.hxx and
    # the headers it pulls in, compiled unmodified) whenever that library was built; sort and
    # boundary exchange are the plain-C restatement either way
    use_ref = ol.ref_available()
    if use_ref:
        R = ol.ref()
        g = og.g
        qk = np.array([k[0] for k in og.kinds], dtype=np.float64)
        mk = np.array([k[1] for k in og.kinds], dtype=np.float64)
        slot_len = int(np.prod(flds.shape[1:]))

        def push_range(prts, off, p0, p1):
            fl = flds[p0:p1]
            rc = R.psc_ref_push_mprts(0, g.deposit, g.gdims, g.length, g.dt, g.fnqs, g.eta, g.n_kinds,
                                      ol.ptr(qk), ol.ptr(mk), ol.ptr(fl), g.im, g.ib, p1 - p0, ol.ptr(prts),
                                      ol.ptr(off[p0:p1 + 1]))
            assert rc == 0
    else:
        def push_range(prts, off, p0, p1):
            L.po_push_mprts_range(G, ol.ptr(flds), ol.ptr(prts), ol.ptr(off), p0, p1)

    def one_step(prts, off):
        def work(b):
            L.po_sort_range(G, ol.ptr(prts), ol.ptr(off), None, b[0], b[1])
            push_range(prts, off, b[0], b[1])
        list(pool.map(work, bounds))
        p2, o2 = spare
        L.po_bnd_particles(G, ol.ptr(prts), ol.ptr(off), ol.ptr(p2), ol.ptr(o2), None, ol.ptr(nd))  # OpenMP over patches
        assert int(o2[-1]) == len(prts)
        spare[0], spare[1] = prts, off
        return p2, o2

    for _ in range(warmup):
        prts, off = one_step(prts, off)
    t0 = time.perf_counter()
    for _ in range(steps):
        prts, off = one_step(prts, off)
    dt = time.perf_counter() - t0
    pool.shutdown()
    kind = "_ref+port" if use_ref else "port"
    sample = ("%d patches of %d^3 cells x %d ppc = %d particles, %d steps of sort+push+deposit+"
              "boundary exchange, %d host threads over patches; push+deposit = %s" %
              (npch, cells, ppc, n, steps, threads,
               "the reference's own headers (oracle/_ref)" if use_ref else "plain-C port"))
    return n * steps / dt, dt / steps * 1e3, sample, threads, kind


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    cells, ppc = args.ref_cells, 64
    # ~4 patches per thread keeps a step around a few seconds
    val, ms, sample, used, kind = cpu_arm(cells, ppc, max(4, 2 * cores), args.steps, min(args.warmup, 1), cores)
    line = {
        "impl": "reference", "metric": "particle-steps/sec (push+deposit+sort)", "value": val,
        "unit": "particle-steps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "S3D-thermal (SURVEY.md 8d), bounded sample: " + sample,
                   "scope": "all host cores of the box, whatever --gpus says: the CPU arm does not scale with N"},
        "cpu_baseline": {"value": val, "unit": "particle-steps/s", "cores": used, "kind": kind,
                         "sample": sample},
        "e2e": {"value": val, "unit": "particle-steps/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------- load balancing

def balance_phase(args, rank, world, local_rank, dist, torch):
    """BASELINE.json configs[3] / north_star "load balancing redistributes patches by particle
    count": a non-uniform plasma (density bump along z, centred in rank 0's slab) on an even
    patch split, stepped, rebalanced with psc_b200_balance (Balance_::operator(),
    psc_balance_impl.hxx:770-1026: load = n_prts + factor_fields * n_cells, recursive bisection
    of the patch list, whole patches moved GPU to GPU over NCCL), stepped again."""
    import psc_b200 as pb
    n, pe = args.balance_cells, 32
    npd = n // pe
    gdims, np3 = (n, n, n * world), (npd, npd, npd * world)
    grid = pb.Grid(gdims=gdims, length=tuple(float(g) for g in gdims), np=np3, dt=0.75 / np.sqrt(3.),
                   kinds=KINDS, nicell=32, rank=rank, n_ranks=world, device=local_rank)
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(pb.Grid.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    grid.nccl_init(bytes(idt.cpu().numpy().tobytes()))
    # density by patch: 12 + 84 exp(-((z - z0) / w)^2) particles per cell and kind, z0 = middle
    # of rank 0's slab, w = 0.35 slabs
    p0, npl = grid.patch_begin(), grid.n_patches()
    iz = (p0 + np.arange(npl)) // (npd * npd)
    zc = (iz + .5) * pe
    ppc = np.round(12 + 84 * np.exp(-((zc - .5 * n) / (.35 * n)) ** 2)).astype(np.int32)
    mprts, mflds = pb.Mparticles(grid), pb.MfieldsState(grid)
    mprts.setup_thermal(ppc, list(VTH), seed=99)
    mflds.fill(pb.HZ, 0.1)
    psc = pb.Psc(grid, mflds, mprts, sort_interval=1, fused=True)
    psc.initialize()

    def gather(v):
        t = torch.zeros(world, dtype=torch.float64, device="cuda")
        t[rank] = float(v)
        dist.all_reduce(t)
        return [float(x) for x in t.cpu()]

    def timed_steps(k):
        for _ in range(2):
            psc.step()
        grid.sync()
        dist.barrier()
        torch.cuda.synchronize()
        grid.timer_start()
        for _ in range(k):
            psc.step()
        ms = grid.timer_stop()
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / k

    n_before, patches_before = gather(mprts.size()), gather(grid.n_patches())
    ms_before = timed_steps(5)
    grid.sync()
    dist.barrier()
    t0 = time.perf_counter()
    changed = grid.balance(1.0)
    grid.sync()
    dist.barrier()
    ms_balance = (time.perf_counter() - t0) * 1e3
    # the operator types re-attach to the new decomposition (psc_balance_generation_cnt)
    mprts, mflds = pb.Mparticles(grid), pb.MfieldsState(grid)
    psc = pb.Psc(grid, mflds, mprts, sort_interval=1, fused=True)
    n_after, patches_after = gather(mprts.size()), gather(grid.n_patches())
    ms_after = timed_steps(5)
    n_end = gather(mprts.size())
    grid.close()
    tot = sum(n_before)
    moved = int(sum(abs(a - b) for a, b in zip(patches_before, patches_after)) // 2)
    return {
        "workload": "density bump along z on %d^3 cells per GPU, %d^3-cell patches: %d..%d particles per cell; "
                    "even patch split, 5 timed steps, psc_b200_balance(factor_fields=1), 5 timed steps" %
                    (n, pe, 2 * 12, 2 * 96),
        "changed": bool(changed), "particles_total": int(tot), "particles_conserved": int(sum(n_end)) == int(tot),
        "particles_by_rank_before": [int(x) for x in n_before], "particles_by_rank_after": [int(x) for x in n_after],
        "patches_by_rank_before": [int(x) for x in patches_before], "patches_by_rank_after": [int(x) for x in patches_after],
        "patches_moved": moved,
        "imbalance_before": max(n_before) / (tot / world), "imbalance_after": max(n_after) / (tot / world),
        "ms_per_step_before": ms_before, "ms_per_step_after": ms_after, "ms_balance": ms_balance,
        "particle_steps_per_s_before": tot / (ms_before * 1e-3), "particle_steps_per_s_after": tot / (ms_after * 1e-3),
    }


# ---------------------------------------------------------------------------- GPU arm

def bind_to_gpu_numa_node(local_rank):
    """pin this process (and with it the pinned host buffers it allocates) to the NUMA node the GPU
    hangs off: at 8 ranks the host<->device copies of the e2e loop otherwise cross the sockets"""
    try:
        out = subprocess.run(["nvidia-smi", "-i", str(local_rank), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=10).stdout.strip().lower()
        bus = out[-12:] if len(out) >= 12 else out  # 0000:1b:00.0
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


def run_b200(args, rank, world, local_rank):
    numa_node = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    import torch
    import psc_b200 as pb
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # ---- multi-rank parity, ahead of the timed region: small grids through the same NCCL
    # halo / migration / balance paths, compared with the CPU oracle run on the whole domain
    parity = None
    if world > 1 and not args.no_parity:
        import multi_gpu_check as mgc
        p_ok, p_cases = mgc.bench_parity_cases(rank, world, local_rank)
        parity = {"pass": bool(p_ok), "cases": len(p_cases), "detail": p_cases,
                  "what": "tests/multi_gpu_check.py cases vs the CPU oracle on the whole domain (rank 0 gathers)"}

    n, ppc, pe = args.cells, args.ppc, args.patch
    npd = n // pe
    yz = args.dim == "yz"
    if yz:
        # S2D-thermal: 1 x n x n cells per GPU, stacked along z
        gdims, np3 = (1, n, n * world), (1, npd, npd * world)
        n_cells_gpu, dt = n * n, 0.75 / np.sqrt(2.)
    else:
        gdims, np3 = (n, n, n * world), (npd, npd, npd * world)
        n_cells_gpu, dt = n ** 3, 0.75 / np.sqrt(3.)
    grid = pb.Grid(gdims=gdims, length=tuple(float(g) for g in gdims), np=np3,
                   dt=dt, kinds=KINDS, nicell=ppc // 2, rank=rank, n_ranks=world,
                   device=local_rank, max_n_prts=int(n_cells_gpu * ppc * 1.02))
    if world > 1:
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(pb.Grid.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        grid.nccl_init(bytes(idt.cpu().numpy().tobytes()))
    for k, v in (("fma", args.fma), ("tiled", args.tiled), ("tma", args.tma),
                 ("fused_sort", args.fused_sort), ("gapped", args.gapped), ("gap_slack", args.gap_slack), ("overlap", args.overlap)):
        grid.set_option(k, v)
    if args.tile:
        grid.set_option("tile", args.tile)
    if args.threads:
        grid.set_option("threads", args.threads)
    if args.min_blocks:
        grid.set_option("min_blocks", args.min_blocks)
    extra_opts = {}
    for kv in args.opt:
        k, v = kv.split("=")
        extra_opts[k] = float(v)
        grid.set_option(k, float(v))
    mprts, mflds = pb.Mparticles(grid), pb.MfieldsState(grid)
    mprts.setup_thermal(ppc // 2, list(VTH), seed=1234)
    mflds.fill(pb.HX if yz else pb.HZ, 0.1)
    n_prts = mprts.size()
    psc = pb.Psc(grid, mflds, mprts, sort_interval=args.sort_interval, fused=True)
    psc.initialize()

    def barrier():
        grid.sync()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        psc.step()
    grid.set_option("profile", 1)
    grid.profile_reset()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()  # before the barrier: its start-up must not delay rank 0 inside the timed region
    barrier()
    l0 = grid.get_stat("n_launches")
    grid.timer_start()
    for _ in range(args.steps):
        psc.step()
    ms = grid.timer_stop()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = int(grid.get_stat("n_launches") - l0)
    prof = grid.profile()
    grid.set_option("profile", 0)
    n_after = mprts.size()
    if dist:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        cnt = torch.tensor([float(n_prts), float(n_after)], dtype=torch.float64, device="cuda")
        dist.all_reduce(cnt)
        n_total, n_after_total = int(cnt[0].item()), int(cnt[1].item())
    else:
        n_total, n_after_total = n_prts, n_after
    assert n_after_total == n_total, "periodic run lost particles: %d -> %d" % (n_total, n_after_total)
    value = n_total * args.steps / (ms * 1e-3)

    # ---- e2e: operator-level C ABI with host buffers in the loop
    e2e = None
    if not args.no_e2e:
        shape_eb, shape_j = mflds.shape(6), mflds.shape(3)
        h_eb = torch.empty(shape_eb, dtype=torch.float32, pin_memory=True).numpy()
        h_eb[:] = mflds.download(pb.EX, pb.EX + 6)
        h_j = torch.empty(shape_j, dtype=torch.float32, pin_memory=True).numpy()
        lib, ctx = grid.lib, grid.ctx
        k_e2e = max(1, min(args.steps, args.e2e_steps))
        sort_, pushp, bndp, bnd, bndf = pb.Sort(), pb.PushParticles(), pb.BndParticles(grid), pb.Bnd(), pb.BndFields()
        out_en = np.zeros(8)
        prm = pb.StepParams(sort=1, marder_loop=0, marder_diffusion=0., push_fields=1, checks=0)
        prm_en = pb.StepParams(sort=1, marder_loop=0, marder_diffusion=0., push_fields=1, checks=0, energies=1)
        p_eb, p_j, p_en = h_eb.ctypes.data_as(C.c_void_p), h_j.ctypes.data_as(C.c_void_p), out_en.ctypes.data_as(C.c_void_p)

        def e2e_loop(pipelined, k_e2e=k_e2e):
            parts = np.zeros(4)
            barrier()
            t0 = time.perf_counter()
            for k in range(k_e2e):
                ta = time.perf_counter()
                if not pipelined or k == 0:
                    pb.check(lib.psc_b200_mflds_upload(ctx, 0, pb.EX, pb.EX + 6, p_eb))
                tb = time.perf_counter()
                if pipelined:
                    # push, then the particle re-sort and the field chain side by side; J comes down
                    # and the next step's E,B go up on the field stream while the sort runs
                    pb.check(lib.psc_b200_step_begin(ctx, C.byref(prm_en)))
                    pb.check(lib.psc_b200_mflds_download_async(ctx, 0, pb.JXI, pb.JXI + 3, p_j))
                    pb.check(lib.psc_b200_io_wait(ctx))
                    tc = time.perf_counter()
                    pb.check(lib.psc_b200_mflds_upload_async(ctx, 0, pb.EX, pb.EX + 6, p_eb))
                    pb.check(lib.psc_b200_step_end(ctx))
                    td = time.perf_counter()
                else:
                    pb.check(lib.psc_b200_step(ctx, C.byref(prm)))
                    grid.sync()
                    tc = time.perf_counter()
                    pb.check(lib.psc_b200_mflds_download(ctx, 0, pb.JXI, pb.JXI + 3, p_j))
                    td = time.perf_counter()
                if pipelined:
                    pb.check(lib.psc_b200_last_energies(ctx, p_en))  # reduced inside the step
                else:
                    pb.check(lib.psc_b200_energies(ctx, p_en))
                te = time.perf_counter()
                parts += (tb - ta, tc - tb, td - tc, te - td)
            barrier()
            dt = time.perf_counter() - t0
            if dist:
                t = torch.tensor([dt], dtype=torch.float64, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            return dt, parts

        e2e_loop(False, 1)  # untimed: pinned landing zones, stream and event creation, first-touch costs
        e2e_loop(True, 2)
        dt_sync, parts_sync = e2e_loop(False)
        dt_e2e, parts = e2e_loop(True)
        names = ("first_upload", "push..J_on_host", "upload+sort_end", "energies")
        e2e = {"value": n_total * k_e2e / dt_e2e, "unit": "particle-steps/s",
               "h2d_bytes_per_step": int(h_eb.nbytes) * world, "d2h_bytes_per_step": (int(h_j.nbytes) + 64) * world,
               "steps": k_e2e, "numa_node_of_rank0": numa_node,
               "ms_per_step": {k: round(float(v) / k_e2e * 1e3, 2) for k, v in zip(names, parts)},
               "what": "per step through the C ABI with pinned HOST buffers: psc_b200_step_begin (push + deposit, then "
                       "sort || J ghosts + Yee), J (3 comps, all patches) down as soon as it is final, the next "
                       "step's E,B (6 comps) up behind the running sort, psc_b200_step_end; the energies are reduced "
                       "inside the step (fields behind the Yee update, particles behind the sort) and read back",
               "synchronous": {"value": n_total * k_e2e / dt_sync,
                               "ms_per_step": {k: round(float(v) / k_e2e * 1e3, 2)
                                               for k, v in zip(("upload", "step", "download", "energies"), parts_sync)},
                               "what": "upload E,B; psc_b200_step; download J; energies -- one after the other"}}

    grid_closed = False
    balance = None
    if world > 1 and not args.no_balance:
        # (after the timed region; the headline context is released first: both do not fit)
        grid.close()
        grid_closed = True
        try:
            balance = balance_phase(args, rank, world, local_rank, dist, torch)
        except Exception as e:  # reported, never fatal for the headline line
            balance = {"error": str(e)[:300]}
    if rank != 0:
        if not grid_closed:
            grid.close()
        return
    peak, peak_src = measured_peak()
    kernels = {k: {"ms_per_step": v[0] / args.steps, "launches_per_step": v[1] / args.steps}
               for k, v in prof.items()}
    push_key = next((k for k in ("push_lean", "push_gap", "push_tiled_tma", "push_tiled", "push_general") if k in prof), None)
    roofline = None
    if push_key:
        t_push = prof[push_key][0] / max(1, prof[push_key][1]) * 1e-3
        ach = B_PUSH * n_prts / t_push / 1e9
        roofline = {"bound": "hbm", "kernel": push_key, "achieved": ach, "peak": peak, "unit": "GB/s",
                    "frac": ach / peak, "traffic": None, "peak_source": peak_src,
                    "algorithmic_bytes_per_particle": B_PUSH,
                    "step_frac": (value / world) * B_STEP / 1e9 / peak,
                    "step_algorithmic_bytes_per_particle": B_STEP}
    # DRAM bytes of one launch of the dominant kernel from the last `ncu --set full`
    # capture of this workload (profiles/push_traffic.json, from tools/ncu_summary.py's dram__bytes lines)
    if roofline:
        try:
            with open(os.path.join(ROOT, "profiles", "push_traffic.json")) as f:
                tr = json.load(f)
            # (only if the capture is of THIS kernel and THIS workload: a stale file reads as null)
            if tr.get("particles") == n_prts and tr.get("kernel_key") == push_key:
                roofline["traffic"] = tr["dram_bytes_per_launch"]
                roofline["traffic_source"] = tr.get("source")
        except Exception:
            pass
    cpu = None
    if not args.no_cpu and world == 1:  # (the contract: rank 0 at N = 1 only; `--impl reference` times it at any N)
        cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        v, _, sample, used, kind = cpu_arm(32, 64, max(4, 2 * cores), 2, 1, cores)
        cpu = {"value": v, "unit": "particle-steps/s", "cores": used, "kind": kind, "sample": sample}
    line = {
        "metric": "particle-steps/sec (push+deposit+sort)", "value": value, "unit": "particle-steps/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": (("S2D-thermal (yz): 1 x %d^2 cells x %d ppc per GPU, %d^2-cell patches, periodic, "
                                 if yz else
                                 "S3D-thermal: %d^3 cells x %d ppc per GPU, %d^3-cell patches, periodic, ")
                                + "full Psc::step (sort+push+deposit+exchange+J ghosts+Yee E/H), sort every %s")
                               % (n, ppc, pe, "step" if args.sort_interval == 1 else "%d steps" % args.sort_interval),
                   "particles_per_gpu": n_prts, "cells_per_gpu": n_cells_gpu, "parallelism": "slabs along z, %d rank(s)" % world,
                   "l2": "working set (%.1f GB of particles per GPU) >> 126 MB L2, no flush needed" % (n_prts * 32 / 1e9),
                   "fma": args.fma, "options": {"tiled": args.tiled, "tma": args.tma,
                                                "fused_sort": args.fused_sort, "gapped": args.gapped, "overlap": args.overlap,
                                                **extra_opts}},
        "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
        "cpu_baseline": cpu, "kernels": kernels,
    }
    if parity is not None:
        line["parity_check"] = parity
    if balance is not None:
        line["balance"] = balance
    print(json.dumps(line), flush=True)
    if not grid_closed:
        grid.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--dim", default="xyz", choices=["xyz", "yz"],
                    help="xyz: the headline S3D workload; yz: S2D-thermal (use --cells 4096 --ppc 100)")
    ap.add_argument("--cells", type=int, default=256, help="cells per GPU edge")
    ap.add_argument("--patch", type=int, default=32, help="cells per patch edge")
    ap.add_argument("--ppc", type=int, default=64)
    ap.add_argument("--fma", type=int, default=0)
    ap.add_argument("--tiled", type=int, default=1)
    ap.add_argument("--tma", type=int, default=1)
    ap.add_argument("--fused-sort", dest="fused_sort", type=int, default=1)
    ap.add_argument("--sort-interval", dest="sort_interval", type=int, default=1,
                    help="PscParams::sort_interval; the headline metric sorts every step, PSC's decks every 10th")
    ap.add_argument("--gapped", type=int, default=0, help="gapped particle store (no sort pass)")
    ap.add_argument("--gap-slack", dest="gap_slack", type=int, default=0)
    ap.add_argument("--overlap", type=int, default=0, help="field chain on a second stream next to the sort")
    ap.add_argument("--tile", type=int, default=0)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--min-blocks", dest="min_blocks", type=int, default=0)
    ap.add_argument("--e2e-steps", dest="e2e_steps", type=int, default=8)
    ap.add_argument("--ref-cells", dest="ref_cells", type=int, default=32,
                    help="--impl reference: cells per patch edge of the bounded sample")
    ap.add_argument("--opt", action="append", default=[], metavar="NAME=VALUE",
                    help="psc_b200_set_option(NAME, VALUE), repeatable (e.g. --opt lean=0)")
    ap.add_argument("--no-balance", dest="no_balance", action="store_true",
                    help="world > 1: skip the load-balancing measurement that runs after the timed region")
    ap.add_argument("--balance-cells", dest="balance_cells", type=int, default=128,
                    help="cells per GPU edge of the load-balancing workload")
    ap.add_argument("--no-parity", dest="no_parity", action="store_true",
                    help="world > 1: skip the multi-rank oracle comparison that runs ahead of the timed region")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, world, local_rank)
        if world > 1:
            import torch.distributed as dist
            if dist.is_initialized():
                dist.destroy_process_group()


if __name__ == "__main__":
    main()
