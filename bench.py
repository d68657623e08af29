#!/usr/bin/env python3
"""bench.py -- force-step throughput of the B200 TreePM engine.

One "step" = one full gravity force step on a PM step of the reference
(run.c:519-548 with SplitGravityTimestepsOn=0): gravpm_force + force_tree_full
+ grav_short_tree over all particles of a synthetic 256^3 dark-matter box
(BASELINE.json configs[1]: Nmesh 768, Asmth 1.5, TreeRcut 6, ErrTolForceAcc
0.002, relative opening criterion fed by the previous step's accelerations).

  value : particles / second, inputs resident in HBM, timed with CUDA events on
          the engine's stream (max over ranks).
  e2e   : the same step through b200_force_step_aos on the reference's 160-byte
          particle records in pinned HOST memory, H2D + D2H inside the timed region.
  --impl reference : the CPU path (the reference's own tree C from oracle/_ref when
          built, else the oracle port; PM = oracle restatement + pocketfft) on all
          host threads, on a bounded sample of the same workload.

Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

G = 43.0071
METRIC = "particles/sec per force step (PM+tree)"
UNIT = "particles/s"


# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full captures
# (profiles/r01_*_ncu_summary.txt), 256^3 workload
NCU_TRAFFIC = {"k_grav_pairs": 6.85e9, "k_grav_walk": 6.26e9}


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_force_step(ng, steps=1, warmup=0):
    """The CPU path on an ng^3 Zel'dovich box with the bench's parameters.
    Returns (particles_per_second, kind, seconds_per_step)."""
    import numpy as np
    import oracle
    ics = importlib.import_module("mp-gadget_b200.ics")
    box = float(ng)
    nmesh = ics.default_nmesh(ng)
    pos, mass = ics.zeldovich_lattice(ng, box)
    n = len(mass)
    par = ics.tree_params(box, n, treeusebh=0)
    ref = None
    try:
        refmod = importlib.import_module("oracle.ref")
        ref = refmod.load()
    except Exception:
        ref = None
    kind = "reference" if ref is not None else "port"
    nthr = host_threads()
    os.environ.setdefault("OMP_NUM_THREADS", str(nthr))
    oldacc = None
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        gpm, _, _ = oracle.pm_force(pos, mass, box, nmesh, 1.5, G, workers=nthr)
        if ref is not None:
            acc = ref.tree_gravity(pos, mass, box, nmesh, 1.5, G, par, oldacc)
        else:
            tr = oracle.OracleTree(pos, mass, box)
            acc, _, _ = tr.grav_short_tree(par, G, nmesh, 1.5, oldacc=oldacc)
        dt = time.perf_counter() - t0
        oldacc = acc + gpm
        if it >= warmup:
            times.append(dt)
    t = sum(times) / len(times)
    return n / t, kind, t, nmesh


def run_reference(args, rank):
    if rank != 0:
        return
    ng = args.cpu_ng
    v, kind, t, nm = cpu_force_step(ng, steps=max(1, args.steps), warmup=min(args.warmup, 1))
    sample = "%d^3 Zel'dovich box, Nmesh %d, full force step (PM + tree build + walk), one OMP process" % (ng, nm)
    out = {"metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": 1e3 * t, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
           "data": "synthetic", "impl": "reference",
           "config": {"workload": "256^3 DM-only TreePM force step (sampled at %d^3 on CPU)" % ng, "Nmesh": 768,
                      "Asmth": 1.5, "TreeRcut": 6.0, "ErrTolForceAcc": 0.002},
           "cpu_baseline": {"value": v, "unit": UNIT, "cores": host_threads(), "kind": kind, "sample": sample},
           "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def sharded_config(world, ng_per_gpu):
    """Global lattice, mesh and top-tree depth for `world` GPUs at ~ng_per_gpu^3 particles each."""
    ng_tot = int(round(ng_per_gpu * world ** (1.0 / 3.0)))
    ng_tot += (-ng_tot) % world                      # lattice planes split evenly over the ranks
    # cell = spacing/3 as for 256^3/768; FFT-friendly (2^a 3^b 5^c) multiple of 2*world nearest to 3*ng_tot
    cands = [2 ** a * 3 ** b * 5 ** c for a in range(1, 13) for b in range(0, 7) for c in range(0, 4)]
    cands = [m for m in cands if m % (2 * world) == 0]
    nmesh = min(cands, key=lambda m: abs(m - 3 * ng_tot))
    rcut_spacings = 6.0 * 1.5 * ng_tot / nmesh
    d = 1
    while (1 << (d + 1)) * rcut_spacings * 1.02 < 1.001 * ng_tot and d < 8:
        d += 1
    while (1 << d) % world:
        d -= 1
    return ng_tot, nmesh, d


def run_sharded(args, rank, world, local, dist, pkg, ics, W, K):
    import numpy as np
    import torch
    sh = importlib.import_module("mp-gadget_b200.sharded")
    dev = torch.device("cuda", local)
    ng_tot, nmesh, topdepth = sharded_config(world, args.ng)
    box = float(ng_tot)
    par = ics.tree_params(box, ng_tot ** 3, treeusebh=1)
    e = pkg.Engine(local)
    stream = torch.cuda.ExternalStream(e.stream(), device=dev)
    s = sh.ShardedTreePM(e, box, nmesh, 1.5, G, topdepth, dist=dist, device="cuda:%d" % local)
    # own particles: lattice planes of my x-range (+3 planes margin), kept where the displaced x is mine
    per = ng_tot // world
    pos, mass = ics.planewave_lattice(ng_tot, box, xplanes=(rank * per - 3, (rank + 1) * per + 3), device="cuda:%d" % local)
    keep = s.dom.owner_of(pos[:, 0]) == rank
    pos, mass = pos[keep].contiguous(), mass[keep].contiguous()
    n_own = pos.shape[0]
    ntot = torch.tensor([n_own], dtype=torch.int64, device=dev)
    dist.all_reduce(ntot)
    total = int(ntot.item())
    assert total == ng_tot ** 3, (total, ng_tot ** 3)
    oldacc = None

    def step_dev():
        nonlocal oldacc
        s.load(pos, mass, oldacc=oldacc, rcut_cells=par["Rcut"])     # ghost exchange is part of the step
        gpm, acc, pot = s.force_step(par)
        oldacc = (acc + gpm).contiguous()

    def barrier():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    step_dev()
    par["TreeUseBH"] = 0
    for _ in range(W):
        step_dev()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    l0 = e.kernel_launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    phase = {}
    ev0.record(stream)
    t0 = time.perf_counter()
    for _ in range(K):
        step_dev()
        for k, v in e.timings().items():
            phase[k] = phase.get(k, 0.0) + v / K
    torch.cuda.synchronize()
    wall_dev = (time.perf_counter() - t0) * 1e3
    ev1.record(stream)
    barrier()
    # the harness interleaves torch (default stream) and engine (own stream) work with syncs in
    # between, so wall-clock between device syncs is the honest per-rank step time
    ms_dev = wall_dev
    launches = e.kernel_launches() - l0
    nghost = s.n_tot - s.n_own

    # e2e: own particles from pinned host memory in, accelerations back to pinned host memory
    hpos = torch.empty((n_own, 3), dtype=torch.float64).pin_memory(); hpos.copy_(pos)
    hmass = torch.empty(n_own, dtype=torch.float32).pin_memory(); hmass.copy_(mass)
    hacc = torch.empty((n_own, 3), dtype=torch.float64).pin_memory()
    hgpm = torch.empty((n_own, 3), dtype=torch.float64).pin_memory()
    hpot = torch.empty(n_own, dtype=torch.float64).pin_memory()
    hold = torch.empty((n_own, 3), dtype=torch.float64).pin_memory(); hold.copy_(oldacc)

    def step_e2e():
        dp = hpos.to(dev, non_blocking=True); dm = hmass.to(dev, non_blocking=True); do = hold.to(dev, non_blocking=True)
        s.load(dp, dm, oldacc=do, rcut_cells=par["Rcut"])
        gpm, acc, pot = s.force_step(par)
        hgpm.copy_(gpm, non_blocking=True); hacc.copy_(acc, non_blocking=True); hpot.copy_(pot, non_blocking=True)
        torch.cuda.synchronize()

    step_e2e()
    barrier()
    Ke = max(2, min(K, 5))
    t0 = time.perf_counter()
    for _ in range(Ke):
        step_e2e()
    ms_e2e = (time.perf_counter() - t0) * 1e3
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    tt = torch.tensor([ms_dev, ms_e2e, float(nghost), float(n_own)], dtype=torch.float64, device=dev)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e, max_ghost, max_own = [float(x) for x in tt]
    tl = torch.tensor([float(launches)], dtype=torch.float64, device=dev)
    dist.all_reduce(tl)
    if rank == 0:
        hbm, how = peaks()
        pairs_ms = phase["walk_post"]
        pairs_bytes = 4.0 * phase["walk_pieces"] + 124.0 * n_own
        out = {
            "metric": METRIC, "value": total * K / (ms_dev * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%d^3 DM-only TreePM force step sharded over %d GPUs (~%d^3 per GPU)" % (ng_tot, world, args.ng),
                       "particles_total": total, "particles_per_gpu_max": int(max_own), "ghosts_per_gpu_max": int(max_ghost),
                       "Nmesh": nmesh, "Asmth": 1.5, "TreeRcut": 6.0, "ErrTolForceAcc": 0.002, "toptree_depth": topdepth,
                       "opening": "relative (TreeUseBH=0) after one BH pass",
                       "ics": "lattice + periodic plane-wave displacement field, rms 0.2 spacing",
                       "l2": "inputs larger than L2",
                       "parallelism": "x-slab domain of top-tree cell layers; ghost-layer import + top-moment all-reduce (tree), "
                                      "slab FFT with NCCL all-to-all transposes + halo planes (PM)"},
            "e2e": {"value": total * Ke / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": n_own * (24 + 4 + 24),
                    "d2h_bytes_per_step": n_own * 56, "ms_per_step": ms_e2e / Ke,
                    "api": "ShardedTreePM.load + force_step (pinned host pos/mass/oldacc in, acc/gpm/pot out), per rank"},
            "gpu_launches": int(tl.item()),
            "clocks": clocks,
            "roofline": {"kernel": "k_grav_pairs", "bound": "hbm", "achieved": pairs_bytes / (pairs_ms * 1e-3) / 1e9, "peak": hbm,
                         "unit": "GB/s", "frac": pairs_bytes / (pairs_ms * 1e-3) / 1e9 / hbm, "traffic": None, "peak_source": how,
                         "note": "rank 0; pair summation bound by the fp64/conversion pipes and the L1 data path, not HBM; "
                                 "compulsory bytes only (SURVEY 8d K8)"},
            "phases_ms": phase,
            "timing": "wall clock between device synchronisations, max over ranks (engine stream + torch stream interleave)",
        }
        print(json.dumps(out), flush=True)
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--ng", type=int, default=256, help="particles per dimension per GPU")
    ap.add_argument("--cpu-ng", type=int, default=128, help="CPU-baseline sample size per dimension")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-hydro", action="store_true", help="skip the SPH density + hydro timing (configs[2] gas part)")
    ap.add_argument("--steploop", action="store_true", help="also time hierarchical KDK sub-steps with the particle state resident "
                    "in HBM (b200_step_*; off by default until its first hardware run)")
    ap.add_argument("--ics", default="planewave", choices=["planewave", "fft"])
    ap.add_argument("--rms", type=float, default=0.2)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import numpy as np
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this engine has no CPU fallback")
    pkg = importlib.import_module("mp-gadget_b200")
    ics = importlib.import_module("mp-gadget_b200.ics")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    W = max(args.warmup, 3)
    K = args.steps
    if world > 1:
        return run_sharded(args, rank, world, local, dist, pkg, ics, W, K)
    ng = args.ng
    box = float(ng)
    nmesh = ics.default_nmesh(ng)
    # same generator as the multi-GPU arm: lattice + periodic plane-wave displacement field
    if args.ics == "fft":
        pos, mass = ics.zeldovich_lattice(ng, box)
        d_pos, d_mass = torch.from_numpy(pos).cuda(), torch.from_numpy(mass).cuda()
    else:
        d_pos, d_mass = ics.planewave_lattice(ng, box, device="cuda", rms=args.rms)
        pos = d_pos.cpu().numpy(); mass = d_mass.cpu().numpy()
    n = len(mass)
    par = ics.tree_params(box, n, treeusebh=1)

    e = pkg.Engine(local)
    stream = torch.cuda.ExternalStream(e.stream(), device=torch.device("cuda", local))
    e.gravpm_init_periodic(box, 1.5, nmesh, G)

    # ---- device-resident arm ------------------------------------------------
    d_acc = torch.empty((n, 3), dtype=torch.float64, device="cuda")
    d_gpm = torch.empty((n, 3), dtype=torch.float64, device="cuda")
    d_pot = torch.empty(n, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    e.set_particles_dev(d_pos.data_ptr(), d_mass.data_ptr(), n)

    def step_dev():
        # gravpm_force + force_tree_full + grav_short_tree in one call: the PM step runs on a
        # second stream concurrently with the tree build and walk
        e.force_step_dev(par, d_gpm.data_ptr(), d_acc.data_ptr(), d_pot.data_ptr())
        e.oldacc_from_last_step()

    def step_dev_serial():
        e.gravpm_force_dev(d_gpm.data_ptr(), None)
        e.force_tree_full(box)
        e.grav_short_tree_dev(par, d_acc.data_ptr(), d_pot.data_ptr())
        e.oldacc_from_last_step()

    def barrier():
        torch.cuda.synchronize()

    step_dev_serial()                # first pass uses the Barnes-Hut angle (TreeUseBH=2 semantics, gravshort-tree.c:148-151)
    par["TreeUseBH"] = 0
    for _ in range(W):
        step_dev()
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    l0 = e.kernel_launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    phase = {}
    ev0.record(stream)
    for _ in range(K):
        step_dev()
    ev1.record(stream)
    barrier()
    ms_dev = ev0.elapsed_time(ev1)
    launches = e.kernel_launches() - l0
    # per-kernel durations for the roofline table: the same K steps issued serially on one
    # stream (in the timed region above the PM kernels overlap the walk, which stretches both)
    ev4, ev5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev4.record(stream)
    for _ in range(K):
        step_dev_serial()
        for k, v in e.timings().items():
            phase[k] = phase.get(k, 0.0) + v
    ev5.record(stream)
    barrier()
    ms_serial = ev4.elapsed_time(ev5) / K
    info = e.tree_info
    for k in phase:
        phase[k] /= K

    # ---- end-to-end arm: the reference's AoS in pinned host memory --------------
    P = np.zeros(n, dtype=pkg.PARTICLE_DTYPE)
    P["Pos"] = pos; P["Mass"] = mass; P["Type"] = 1; P["ID"] = np.arange(n)
    acc_h = d_acc.cpu().numpy(); gpm_h = d_gpm.cpu().numpy()
    P["FullTreeGravAccel"] = acc_h; P["GravPM"] = gpm_h
    pinned = torch.empty(n * 160, dtype=torch.uint8).pin_memory()
    pinned.numpy()[:] = P.view(np.uint8).reshape(-1)
    del P, d_pos, d_mass
    for _ in range(2):
        e.force_step_aos(None, par, ptr=pinned.data_ptr(), n=n)
    barrier()
    t0 = time.perf_counter()
    ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev2.record(stream)
    Ke = max(2, min(K, 5))
    for _ in range(Ke):
        e.force_step_aos(None, par, ptr=pinned.data_ptr(), n=n)
    ev3.record(stream)
    barrier()
    ms_e2e = ev2.elapsed_time(ev3)
    wall_e2e = time.perf_counter() - t0
    clocks = sampler.stop()
    te2e = e.timings()
    # check the AoS result against the device arm (same inputs up to the oldacc refresh)
    Pout = pinned.numpy().view(pkg.PARTICLE_DTYPE)
    chk = float(np.abs(Pout["GravPM"][:1000] - gpm_h[:1000]).max() / (np.abs(gpm_h[:1000]).max() + 1e-300))

    total = n * world
    value = total * K / (ms_dev * 1e-3)
    e2e_v = total * Ke / (ms_e2e * 1e-3)
    hbm, how = peaks()
    walk_ms = phase["walk"]
    pairs_ms = phase["walk_post"]
    nn = int(info.numnodes)
    pieces = phase["walk_pieces"]
    # SURVEY 8d K8 split over the two kernels.  k_grav_walk: 40 B in per target + node rows read once
    # (80 B here) + its outputs (36 B partial sums/count per target, 4 B per queued leaf piece).
    # k_grav_pairs: the same lists and partial sums in, the source rows read once (32 B), 24 B
    # position in and 32 B result out per target.
    walk_bytes = 40.0 * n + 80.0 * nn + 36.0 * n + 4.0 * pieces
    pairs_bytes = 4.0 * pieces + 36.0 * n + 32.0 * n + 24.0 * n + 32.0 * n
    pairs_gbs = pairs_bytes / (pairs_ms * 1e-3) / 1e9
    pair_slots = 8.0 * pieces          # pair evaluations issued (one source slot per lane)
    N3 = float(nmesh) ** 3
    Mc = float(nmesh) ** 2 * (nmesh // 2 + 1)
    kern = {
        "k_grav_walk": {"ms": walk_ms, "alg_bytes": walk_bytes},
        "k_grav_pairs": {"ms": pairs_ms, "alg_bytes": pairs_bytes, "pair_evaluations": pair_slots,
                         "Gpairs_per_s": pair_slots / (pairs_ms * 1e-3) / 1e9},
        "k_pm_deposit+clear": {"ms": phase["pm_deposit"], "alg_bytes": 156.0 * n + 8 * N3},
        "cufft_d2z": {"ms": phase["pm_fft_forward"], "alg_bytes": 16 * N3},
        "k_pm_potential_transfer": {"ms": phase["pm_transfer"], "alg_bytes": 32 * Mc},
        "cufft_z2d": {"ms": phase["pm_fft_inverse"], "alg_bytes": 16 * N3},
        # difference + readout fused: the potential mesh is read once (8 B/cell) instead of the
        # 32 B/cell of the three force meshes, plus the particle side of the gather
        "k_pm_readout_fused": {"ms": phase["pm_gradient"] + phase["pm_readout"], "alg_bytes": 8 * N3 + (24.0 + 32.0) * n},
        "tree_build": {"ms": phase["tree_total"], "alg_bytes": (96 + 28 + 32) * float(n) + 80.0 * nn},
    }
    for k in kern.values():
        k["GBps"] = k["alg_bytes"] / (k["ms"] * 1e-3) / 1e9 if k["ms"] > 0 else None
        k["frac"] = k["GBps"] / hbm if k["GBps"] else None
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "%d^3 DM-only TreePM force step per GPU (gravpm_force + force_tree_full + grav_short_tree)" % ng,
                   "particles_per_gpu": n, "Nmesh": nmesh, "Asmth": 1.5, "TreeRcut": 6.0, "ErrTolForceAcc": 0.002,
                   "opening": "relative (TreeUseBH=0) after one BH pass", "ics": "lattice + periodic plane-wave displacement field, rms 0.2 spacing",
                   "l2": "inputs larger than L2 (particle arrays %.0f MB, mesh %.1f GB)" % (n * 28 / 1e6, N3 * 8 / 1e9),
                   "parallelism": "single GPU" if world == 1 else "independent replicas, one box per GPU (no data-path collective yet)"},
        "e2e": {"value": e2e_v, "unit": UNIT, "h2d_bytes_per_step": n * 160, "d2h_bytes_per_step": n * 160,
                "ms_per_step": ms_e2e / Ke, "wall_ms_per_step": 1e3 * wall_e2e / Ke,
                "h2d_ms": te2e["h2d"], "d2h_ms": te2e["d2h"], "api": "b200_force_step_aos (pinned host AoS, 160 B/particle)",
                "check_vs_device_arm": chk},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"kernel": "k_grav_pairs", "bound": "hbm", "achieved": pairs_gbs, "peak": hbm, "unit": "GB/s",
                     "frac": pairs_gbs / hbm, "traffic": NCU_TRAFFIC.get("k_grav_pairs"), "peak_source": how,
                     "note": "dominant kernel of the step; a pair summation whose operands hit L1/L2, bound by the fp64 and "
                             "conversion pipes and the L1 data path, not by HBM (ncu: profiles/); bytes = SURVEY 8d K8 compulsory "
                             "traffic of this kernel; see kernels[] for the HBM-bound PM kernels"},
        "phases_ms": phase, "ms_per_step_serial": ms_serial,
        "kernels": kern,
        "tree": {"numnodes": nn, "maxdepth": int(info.maxdepth)},
    }
    if not args.no_hydro and world == 1:
        # BASELINE.json configs[2], gas part: SPH density (with the smoothing-length iteration) and
        # hydro force on ng^3 gas particles (gas-only tree = force_tree_rebuild_mask(GASMASK), run.c:466-489)
        try:
            rng = np.random.default_rng(1)
            vel = rng.standard_normal((n, 3)) * 0.05
            h0 = np.full(n, 3.0 * box / ng * 0.8)
            sp = pkg.sph_params(KernelType=2, DensityIndependentSphOn=1, MinGasHsml=1e-4, atime=0.1, hubble=3.0, dloga_bin=0.01)
            e.set_particles(pos, mass, type=np.zeros(n, np.uint8))
            rec = {}
            for rep in range(3):
                e.force_tree_build(box, mask=1)
                tree_ms = e.timings()["tree_total"]
                e.sph_set_gas(h0, vel=vel, entropy=np.ones(n))
                d = e.density(sp, update_hsml=1, DoEgyDensity=1); tm_d = e.timings()["sph_density"]
                h = e.hydro_force(sp); tm_h = e.timings()["sph_hydro"]
                rec = {"n_gas": n, "kernel": "quintic, 113 neighbours, pressure-entropy", "tree_ms": tree_ms, "density_ms": tm_d,
                       "hydro_ms": tm_h, "density_passes_mean": float(d["niter"].mean()), "density_passes_max": int(d["niter"].max()),
                       "neighbours_mean": float(d["ninteract"].mean()), "hydro_candidates_mean": float(h["ninteract"].mean()),
                       "gas_per_s_density": n / (tm_d * 1e-3), "gas_per_s_hydro": n / (tm_h * 1e-3),
                       "gas_per_s_sph_step": n / ((tree_ms + tm_d + tm_h) * 1e-3)}
            out["hydro"] = rec
        except Exception as ex:
            out["hydro"] = {"failed": repr(ex)}
    if args.steploop and world == 1:
        # sub-steps of the hierarchical integrator (run.c:355-800) without the 160-byte record round trip
        try:
            SL = importlib.import_module("mp-gadget_b200.steploop")
            cosmo = ics.FlatLCDM()
            rng = np.random.default_rng(2)
            S = SL.StepEngine(e, cosmo.sync, cosmo.factor, cosmo.hubble, Omega0=cosmo.Omega0, Hubble=cosmo.Hubble, G=G)
            S.set_particles(pos, mass, np.ones(n, np.uint8), box, vel=0.05 * rng.standard_normal((n, 3)))
            S.set_gravity(ics.tree_params(box, n, treeusebh=2), G, nmesh, 1.5)
            S.set_times(np.zeros(7, np.int64), np.zeros(47, np.int64), np.zeros(47, np.int64))
            sub = []
            for k in range(9):
                t0 = time.perf_counter(); bad, info = S.advance(first=(k == 0), pm=True); dt = time.perf_counter() - t0
                sub.append({"wall_ms": 1e3 * dt, "active": int(info[1]), "is_pm": int(info[2]), "bad": bad})
            out["steploop"] = {"substeps": sub, "host_bytes_per_substep": "scalars only", "aos_roundtrip_bytes_per_force_call": 2 * 160 * n}
        except Exception as ex:
            out["steploop"] = {"failed": repr(ex)}
    if not args.no_cpu and world == 1:
        try:
            v, kind, t, nm = cpu_force_step(args.cpu_ng, steps=1, warmup=0)
            out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": host_threads(), "kind": kind,
                                   "sample": "%d^3 Zel'dovich box, Nmesh %d, one full force step (%.1f s)" % (args.cpu_ng, nm, t)}
        except Exception as ex:      # the baseline is reported, never required
            out["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": host_threads(), "kind": "port", "sample": "failed: %r" % (ex,)}
    print(json.dumps(out), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
