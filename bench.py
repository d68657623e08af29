#!/usr/bin/env python3
"""bench.py -- force-step throughput of the B200 TreePM engine.

One "step" = one full gravity force step on a PM step of the reference (run.c:519-548 with
SplitGravityTimestepsOn=0): gravpm_force + force_tree_full + grav_short_tree over all particles of a synthetic
256^3 dark-matter box per GPU (BASELINE.json configs[1]: Nmesh 768, Asmth 1.5, TreeRcut 6, ErrTolForceAcc 0.002,
relative opening criterion fed by the previous step's accelerations).

  value : particles / second, inputs resident in HBM, timed with CUDA events on the engine's stream (max over ranks).
  e2e   : the same step through b200_force_step_aos on the reference's 160-byte particle records in pinned HOST
          memory, H2D + D2H inside the timed region.
  cpu_baseline / --impl reference : the CPU path on the SAME box and mesh (same generator, same seed): the reference's
          own tree C compiled unmodified (oracle/_ref) + the restated PM (OpenMP C + pocketfft; PFFT is not in the
          image), all host threads.  At N=1 the GPU result of the same step is compared with it particle by
          particle ("parity").
  states: the default box is the displaced lattice (rms 1 spacing) whose leaf occupancy does not depend on the box size;
          the z9 lattice state (rms 0.2 spacing; 8 particles in every leaf on a 2^k lattice) and the clustered state
          of SURVEY 8d are timed as extra entries of the same line.

Prints ONE JSON line on rank 0.
"""
import argparse
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
CPU_KIND = "reference tree + restated PM"       # forcetree.c/treewalk.c/gravshort-tree.c compiled unmodified; PM = oracle_pm.c + pocketfft

# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full captures (profiles/),
# 256^3 workload; None where no capture of the current kernel exists
NCU_TRAFFIC = {"k_grav_pairs": 7.45e9, "k_grav_walk": 10.70e9}      # profiles/r02_grav_{pairs,walk}_ncu_summary.txt
FP64_DFMA_PER_SM_CLK = 64.0                      # B200: 64 fp64 FMA lanes per SM


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), float(p.get("sm_max_mhz", 1965.0)), "measured"
    except Exception:
        return 6650.0, 1965.0, "fallback"


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


def claim_host_threads():
    """The CPU arm uses every core this process may run on.  torch.distributed.run exports OMP_NUM_THREADS=1, so the
    value is overwritten (not defaulted) and must be in place before libgomp is loaded by the oracle libraries."""
    n = host_threads()
    os.environ["OMP_NUM_THREADS"] = str(n)
    os.environ.pop("OMP_PROC_BIND", None)
    return n


class CpuArm:
    """The reference CPU path on one box: tree = the reference's own C where oracle/_ref travelled, else the oracle port."""

    def __init__(self, nthr):
        import oracle
        self.oracle = oracle
        self.nthr = nthr
        self.ref = None
        try:
            refmod = importlib.import_module("oracle.ref")
            if refmod.available():
                self.ref = refmod.Ref(nthreads=nthr)          # omp_set_num_threads(nthr) inside ref_init
        except Exception:
            self.ref = None
        self.kind = CPU_KIND if self.ref is not None else "port"

    def step(self, pos, mass, box, nmesh, par, oldacc):
        """-> (tree acc[n,3], GravPM[n,3], seconds tree, seconds pm)"""
        t0 = time.perf_counter()
        gpm, _, _ = self.oracle.pm_force(pos, mass, box, nmesh, 1.5, G, workers=self.nthr)
        t1 = time.perf_counter()
        if self.ref is not None:
            acc = self.ref.tree_gravity(pos, mass, box, nmesh, 1.5, G, par, oldacc)
        else:
            tr = self.oracle.OracleTree(pos, mass, box)
            acc, _, _ = tr.grav_short_tree(par, G, nmesh, 1.5, oldacc=oldacc)
        t2 = time.perf_counter()
        return acc, gpm, t2 - t1, t1 - t0


def workload_config(ng, nmesh, state):
    return {"workload": "%d^3 DM-only TreePM force step per GPU (gravpm_force + force_tree_full + grav_short_tree)" % ng,
            "Nmesh": nmesh, "Asmth": 1.5, "TreeRcut": 6.0, "ErrTolForceAcc": 0.002,
            "opening": "relative (TreeUseBH=0) after one Barnes-Hut pass",
            "ics": {"displaced": "lattice + periodic plane-wave displacement field, rms 1.0 spacing (seed 181170)",
                    "z9": "lattice + periodic plane-wave displacement field, rms 0.2 spacing (seed 181170)",
                    "clustered": "tests/test_gravity.c:288-302 recipe"}[state], "state": state}


def run_reference(args, rank, world):
    if rank != 0:
        return
    nthr = claim_host_threads()
    import numpy as np
    import torch
    torch.set_num_threads(nthr)
    ics = importlib.import_module("mp-gadget_b200.ics")
    ng = args.ng
    box = float(ng)
    nmesh = ics.default_nmesh(ng)
    pos_t, mass_t = ics.bench_ics(args.state, ng, box, device="cpu")
    pos, mass = pos_t.numpy(), mass_t.numpy()
    n = len(mass)
    cpu = CpuArm(nthr)
    par = ics.tree_params(box, n, treeusebh=1)
    # one Barnes-Hut pass supplies the old accelerations (TreeUseBH = 2 semantics); it is the warm-up step
    acc, gpm, _, _ = cpu.step(pos, mass, box, nmesh, par, None)
    par["TreeUseBH"] = 0
    oldacc = acc + gpm
    times, t_tree, t_pm = [], 0.0, 0.0
    budget = time.perf_counter() + args.cpu_budget
    for it in range(max(1, args.steps)):
        t0 = time.perf_counter()
        acc, gpm, tt, tp = cpu.step(pos, mass, box, nmesh, par, oldacc)
        times.append(time.perf_counter() - t0); t_tree += tt; t_pm += tp
        oldacc = acc + gpm
        if time.perf_counter() > budget:
            break
    t = sum(times) / len(times)
    v = n / t
    cfg = workload_config(ng, nmesh, args.state)
    sample = ("the whole %d^3 / Nmesh %d workload of one GPU, same generator and seed as the GPU arm; %d timed steps of %.1f s "
              "(tree %.1f s, PM %.1f s) after one Barnes-Hut pass" % (ng, nmesh, len(times), t, t_tree / len(times), t_pm / len(times)))
    if world > 1:
        cfg["note"] = "rank 0 times one GPU's share (%d^3) of the %d-GPU weak-scaling workload on the host cores" % (ng, world)
    out = {"metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times), "steps_requested": args.steps,
           "warmup": 1, "ms_per_step": 1e3 * t, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
           "data": "synthetic", "impl": "reference", "config": cfg, "same_config": True,
           "cpu_baseline": {"value": v, "unit": UNIT, "cores": nthr, "kind": cpu.kind, "sample": sample},
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
    # a top cell must stay wider than Rcut also across the periodic seam, where the 1.001 Box root cell costs 0.001 Box
    while 1.001 * ng_tot / (1 << (d + 1)) - 0.001 * ng_tot > rcut_spacings * 1.02 and d < 8:
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
    s = sh.ShardedTreePM(e, box, nmesh, 1.5, G, topdepth, dist=dist, device="cuda:%d" % local, rcut_cells=par["Rcut"])
    # own particles: lattice planes of my x-range (+ margin for the displacement), kept where the displaced x is mine
    per = ng_tot // world
    margin = 8 if args.state == "displaced" else 3
    pos, mass = ics.bench_ics(args.state, ng_tot, box, xplanes=(rank * per - margin, (rank + 1) * per + margin), device="cuda:%d" % local)
    keep = s.dom.owner_of(pos[:, 0]) == rank
    pos, mass = pos[keep].contiguous(), mass[keep].contiguous()
    n_own = pos.shape[0]
    ntot = torch.tensor([n_own], dtype=torch.int64, device=dev)
    dist.all_reduce(ntot)
    total = int(ntot.item())
    assert total == ng_tot ** 3, (total, ng_tot ** 3)
    oldacc = None
    last = {}

    def step_dev():
        nonlocal oldacc
        gpm, acc, pot = s.force_step(pos, mass, oldacc, par)      # ghost exchange is part of the step
        oldacc = (acc + gpm)
        last["gpm"], last["acc"] = gpm, acc

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
    for _ in range(K):
        step_dev()
        for k, v in e.timings().items():
            phase[k] = phase.get(k, 0.0) + v / K
    ev1.record(stream)
    barrier()
    ms_dev = ev0.elapsed_time(ev1)
    launches = e.kernel_launches() - l0
    nghost = s.n_tot - s.n_own
    sharded_phases = s.phase_ms() if hasattr(s, "phase_ms") else {}

    # inputs and results of the last timed step, which the parity check below repeats (the result tensors are views of the
    # engine's output buffers: later steps overwrite them)
    last_oldacc = s.last_oldacc.clone() if getattr(s, "last_oldacc", None) is not None else None
    if not args.no_parity:
        last = {k: v.clone() for k, v in last.items()}

    # e2e: own particles from pinned host memory in, accelerations back to pinned host memory
    hpos = torch.empty((n_own, 3), dtype=torch.float64).pin_memory(); hpos.copy_(pos)
    hmass = torch.empty(n_own, dtype=torch.float32).pin_memory(); hmass.copy_(mass)
    hacc = torch.empty((n_own, 3), dtype=torch.float64).pin_memory()
    hgpm = torch.empty((n_own, 3), dtype=torch.float64).pin_memory()
    hpot = torch.empty(n_own, dtype=torch.float64).pin_memory()
    hold = torch.empty((n_own, 3), dtype=torch.float64).pin_memory(); hold.copy_(oldacc)

    def step_e2e():
        with torch.cuda.stream(stream):
            dp = hpos.to(dev, non_blocking=True); dm = hmass.to(dev, non_blocking=True); do = hold.to(dev, non_blocking=True)
        gpm, acc, pot = s.force_step(dp, dm, do, par)
        with torch.cuda.stream(stream):
            hgpm.copy_(gpm, non_blocking=True); hacc.copy_(acc, non_blocking=True); hpot.copy_(pot, non_blocking=True)
        torch.cuda.synchronize()

    step_e2e()
    barrier()
    Ke = max(2, min(K, 5))
    ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev2.record(stream)
    for _ in range(Ke):
        step_e2e()
    ev3.record(stream)
    barrier()
    ms_e2e = ev2.elapsed_time(ev3)
    clocks = sampler.stop() if rank == 0 else None
    hydro = None
    if args.sharded_hydro:
        del hpos, hmass, hacc, hgpm, hpot, hold
        # gas cells one level coarser than the gravity cut: the ghost layer must be wider than the largest smoothing length
        hd = topdepth - 1
        while hd > 0 and (1 << hd) % world:
            hd -= 1
        hydro = sharded_hydro_entry(rank, world, dev, dist, pkg, sh, e, pos, mass, box, ng_tot, max(hd, 1), stream)
    # parity: every rank checks sampled own targets of the last timed step against a single-engine walk of the same global
    # tree (last: it gathers the whole box on every rank)
    parity = None
    if not args.no_parity:
        s.last_oldacc = last_oldacc
        parity = s.parity_check(pos, mass, last["acc"], last["gpm"], par, nsample=4096)
    tt = torch.tensor([ms_dev, ms_e2e, float(nghost), float(n_own)] + ([parity["acc_max_err_over_mean"]] if parity else [0.0]),
                      dtype=torch.float64, device=dev)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e, max_ghost, max_own, par_max = [float(x) for x in tt]
    tl = torch.tensor([float(launches)], dtype=torch.float64, device=dev)
    dist.all_reduce(tl)
    if rank == 0:
        hbm, smhz, how = peaks()
        pairs_ms = phase["walk_post"]
        pairs_bytes = 4.0 * phase["walk_pieces"] + 124.0 * n_own
        cfg = workload_config(args.ng, nmesh, args.state)
        cfg.update({"workload": "%d^3 DM-only TreePM force step sharded over %d GPUs (~%d^3 per GPU)" % (ng_tot, world, args.ng),
                    "particles_total": total, "particles_per_gpu_max": int(max_own), "ghosts_per_gpu_max": int(max_ghost),
                    "toptree_depth": topdepth, "l2": "inputs larger than L2",
                    "parallelism": "x-slab domain of top-tree cell layers; ghost-layer import + top-moment all-reduce (tree), "
                                   "slab FFT with NCCL all-to-all transposes + halo planes (PM), all issued from C on the engine's streams"})
        out = {
            "metric": METRIC, "value": total * K / (ms_dev * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": cfg,
            "e2e": {"value": total * Ke / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": n_own * (24 + 4 + 24),
                    "d2h_bytes_per_step": n_own * 56, "ms_per_step": ms_e2e / Ke,
                    "api": "b200_sharded_force_step (pinned host pos/mass/oldacc in, acc/gpm/pot out), per rank"},
            "gpu_launches": int(tl.item()),
            "clocks": clocks,
            "roofline": {"kernel": "k_grav_pairs", "bound": "hbm", "achieved": pairs_bytes / (pairs_ms * 1e-3) / 1e9, "peak": hbm,
                         "unit": "GB/s", "frac": pairs_bytes / (pairs_ms * 1e-3) / 1e9 / hbm, "traffic": None, "peak_source": how,
                         "note": "rank 0; pair summation bound by the fp64/conversion pipes and the L1 data path, not HBM; "
                                 "compulsory bytes only (SURVEY 8d K8)"},
            "phases_ms": phase, "sharded_phases_ms": sharded_phases,
            "parity": None if parity is None else dict(parity, acc_max_err_over_mean_all_ranks=par_max),
            "timing": "CUDA events on the engine's stream around K steps, max over ranks",
        }
        if hydro is not None:
            out["hydro"] = hydro
        print(json.dumps(out), flush=True)
    dist.destroy_process_group()


def sharded_hydro_entry(rank, world, dev, dist, pkg, sh, e, pos, mass, box, ng_tot, topdepth, stream, K=2):
    """BASELINE.json configs[4] (gas part): the displaced box as GAS, sharded like the gravity step -- ghost-layer import,
    density with the smoothing-length iteration for the own particles, converged state of the exported particles sent
    back, hmax refresh, symmetric hydro pass -- with the whole SPH state resident in HBM (ShardedSPH).  Timed: the step a
    run takes every time (Hsml predicted from the last step: converged values +- 1 %), all particles active, and a
    mixed-time-bin sub-step in which a quarter of the particles is active and the rest are sources with their old state."""
    import numpy as np
    import torch
    n = pos.shape[0]
    gen = torch.Generator(device=dev); gen.manual_seed(1000 + rank)
    vel = 0.05 * torch.randn((n, 3), dtype=torch.float64, device=dev, generator=gen)
    ent = torch.ones(n, dtype=torch.float64, device=dev)
    h0 = torch.full((n,), 3.0 * 0.8 * box / ng_tot, dtype=torch.float64, device=dev)
    sp = pkg.sph_params(KernelType=2, DensityIndependentSphOn=1, MinGasHsml=1e-4, atime=0.1, hubble=3.0, dloga_bin=0.01)
    s = sh.ShardedSPH(e, box, topdepth, dist=dist, device=str(dev))

    def step(hs, active=None, bins=None):
        nghost = s.load(pos, mass, hs, vel=vel, entropy=ent)
        if bins is not None:
            tab = {k: np.zeros(47) for k in ("gravkick", "hydrokick", "dloga_pred", "drift")}; tab["dloga_bin"] = np.full(47, 0.01)
            s.set_mixed(bins, tab, active)
        d = s.density(sp, DoEgyDensity=1)
        h = s.hydro_force(sp)
        return nghost, d, h

    def timed(fn, reps=3):
        """best of `reps` single steps (max over ranks each): the steps are short and share the node with host-side work"""
        fn()                                        # untimed: sizes the piece pools for this mode
        best, out = None, None
        for _ in range(reps):
            torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(stream)
            out = fn()
            ev1.record(stream)
            torch.cuda.synchronize()
            t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            best = float(t.item()) if best is None else min(best, float(t.item()))
        return best, out

    try:
        nghost, d, h = step(h0)                                         # cold start: converges Hsml (and sizes the pools)
        hconv = d["hsml"].clone()
        hw = (hconv * (1.0 + 0.01 * torch.randn(n, dtype=torch.float64, device=dev, generator=gen))).contiguous()
        ms_cold, _ = timed(lambda: step(h0))
        ms_warm, (nghost, dw, hwf) = timed(lambda: step(hw))
        passes = float(dw["niter"].to(torch.float64).mean().item())
        # mixed time bins: every fourth own particle active (bin 10), the others (bin 12) keep the state of the last full pass
        idx = torch.arange(n, device=dev)
        act = idx[(idx % 4) == 0].to(torch.int32).contiguous()
        bins = torch.where((idx % 4) == 0, 10, 12).to(torch.uint8)
        ms_mixed, _ = timed(lambda: step(hw, active=act, bins=bins))
        tot = torch.tensor([float(n), float(nghost)], dtype=torch.float64, device=dev)
        dist.all_reduce(tot)
        return {"workload": "%d^3 gas (displaced box) sharded over %d GPUs: ghost import + gas tree + density + state return + hydro, state resident in HBM" % (ng_tot, world),
                "n_gas_total": int(tot[0].item()), "ghosts_total": int(tot[1].item()), "kernel": "quintic, 113 neighbours, pressure-entropy",
                "sph_step_cold_ms": ms_cold, "sph_step_ms": ms_warm, "density_passes_mean": passes,
                "gas_per_s": tot[0].item() / (ms_warm * 1e-3),
                "mixed_bin_substep_ms": ms_mixed, "mixed_bin_active_fraction": 0.25,
                "active_gas_per_s_mixed": 0.25 * tot[0].item() / (ms_mixed * 1e-3),
                "timing": "CUDA events on the engine's stream around single steps, max over ranks, best of 3; Hsml of the timed step = converged values +- 1 % (what drift.c:60-70 hands density() every step after the first)"}
    except Exception as ex:
        # a failure on one rank leaves the others inside a collective: say so at once and take the job down (torchrun
        # stops the other ranks) instead of waiting for the NCCL watchdog
        import traceback
        sys.stderr.write("[rank %d] sharded hydro entry failed: %r\n%s\n" % (rank, ex, traceback.format_exc()))
        sys.stderr.flush()
        os._exit(17)


def gpu_state_line(pkg, ics, e, state, ng, box, nmesh, K=2):
    """ms per device-resident force step on another particle state (same sizes): extra entry of the JSON line."""
    import torch
    d_pos, d_mass = ics.bench_ics(state, ng, box, device="cuda")
    n = d_mass.shape[0]
    par = ics.tree_params(box, n, treeusebh=1)
    g = torch.empty((n, 3), dtype=torch.float64, device="cuda"); a = torch.empty_like(g); p = torch.empty(n, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    e.set_particles_dev(d_pos.data_ptr(), d_mass.data_ptr(), n)
    stream = torch.cuda.ExternalStream(e.stream())
    e.force_step_dev(par, g.data_ptr(), a.data_ptr(), p.data_ptr()); e.oldacc_from_last_step()
    par["TreeUseBH"] = 0
    e.force_step_dev(par, g.data_ptr(), a.data_ptr(), p.data_ptr()); e.oldacc_from_last_step()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(K):
        e.force_step_dev(par, g.data_ptr(), a.data_ptr(), p.data_ptr()); e.oldacc_from_last_step()
    ev1.record(stream)
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / K
    tm = e.timings()
    return {"ms_per_step": ms, "particles_per_s": n / (ms * 1e-3),
            "walk_ms": tm["walk"], "pairs_ms": tm["walk_post"], "tree_ms": tm["tree_total"], "pm_ms": tm["pm_total"], "n": n}


def hydro_entry(pkg, ics, e, pos, mass, ng, box, nmesh, par, no_cpu=False):
    """BASELINE.json configs[2]: 256^3 dark matter + 256^3 gas on one GPU: the gravity of all 2 n particles (PM + tree
    + walk) followed by the gas tree (force_tree_rebuild_mask(GASMASK)), density (with the smoothing-length iteration)
    and hydro force of the gas (run.c:466-548)."""
    import numpy as np
    import torch
    n = len(mass)
    try:
        rng = np.random.default_rng(1)
        fb = 0.0472 / 0.288
        gas_pos = np.mod(pos + 0.5 * box / ng, box)                    # half-cell-offset copy of the lattice
        allpos = np.concatenate([pos, gas_pos]); del gas_pos
        allmass = np.concatenate([mass * (1 - fb), mass * fb]).astype(np.float32)
        typ = np.concatenate([np.ones(n, np.uint8), np.zeros(n, np.uint8)])
        N = 2 * n
        vel = rng.standard_normal((N, 3)) * 0.05
        h0 = np.full(N, 3.0 * box / ng * 0.8)
        sp = pkg.sph_params(KernelType=2, DensityIndependentSphOn=1, MinGasHsml=1e-4, atime=0.1, hubble=3.0, dloga_bin=0.01)
        gpar = dict(par); gpar["TreeUseBH"] = 1
        e.set_particles(allpos, allmass, type=typ)
        d_g = torch.empty((N, 3), dtype=torch.float64, device="cuda"); d_a = torch.empty_like(d_g); d_p = torch.empty(N, dtype=torch.float64, device="cuda")
        best = None
        for rep in range(3):
            t0 = time.perf_counter()
            e.force_step_dev(gpar, d_g.data_ptr(), d_a.data_ptr(), d_p.data_ptr()); e.oldacc_from_last_step()
            torch.cuda.synchronize()
            grav_ms = 1e3 * (time.perf_counter() - t0)
            gpar["TreeUseBH"] = 0
            e.force_tree_build(box, mask=1)
            tree_ms = e.timings()["tree_total"]
            e.sph_set_gas(h0, vel=vel, entropy=np.ones(N))
            d = e.density(sp, update_hsml=1, DoEgyDensity=1); tm_d = e.timings()["sph_density"]
            h = e.hydro_force(sp); tm_h = e.timings()["sph_hydro"]
            cur = {"n_dm": n, "n_gas": n, "kernel": "quintic, 113 neighbours, pressure-entropy", "gravity_ms": grav_ms, "gas_tree_ms": tree_ms,
                   "density_ms": tm_d, "hydro_ms": tm_h, "density_passes_mean": float(d["niter"][n:].mean()),
                   "density_passes_max": int(d["niter"][n:].max()), "neighbours_mean": float(d["ninteract"][n:].mean()),
                   "hydro_candidates_mean": float(h["ninteract"][n:].mean()),
                   "gas_per_s_density": n / (tm_d * 1e-3), "gas_per_s_hydro": n / (tm_h * 1e-3),
                   "gas_per_s_sph_step": n / ((tree_ms + tm_d + tm_h) * 1e-3),
                   "dm_gas_step_ms": grav_ms + tree_ms + tm_d + tm_h,
                   "particles_per_s_dm_gas_step": N / ((grav_ms + tree_ms + tm_d + tm_h) * 1e-3)}
            if rep > 0 and (best is None or cur["dm_gas_step_ms"] < best["dm_gas_step_ms"]):
                best = cur
        rec = best
        # The step a run takes every time after the first: density() starts from the smoothing lengths drift.c:60-70
        # predicted from the last step (here: the converged values +- 1 %), not from the initial guess.
        hw = d["hsml"] * (1.0 + 0.01 * rng.standard_normal(N))
        for rep in range(2):
            e.force_tree_build(box, mask=1)
            e.sph_set_gas(hw, vel=vel, entropy=np.ones(N))
            dw = e.density(sp, update_hsml=1, DoEgyDensity=1); tm_dw = e.timings()["sph_density"]
        rec.update({"density_predicted_hsml_ms": tm_dw, "density_predicted_hsml_passes_mean": float(dw["niter"][n:].mean()),
                    "dm_gas_step_predicted_hsml_ms": rec["gravity_ms"] + rec["gas_tree_ms"] + tm_dw + rec["hydro_ms"],
                    "gas_per_s_sph_step_predicted_hsml": n / ((rec["gas_tree_ms"] + tm_dw + rec["hydro_ms"]) * 1e-3),
                    "predicted_hsml": "converged Hsml +- 1 % (rms): what density() is handed on every step after the first"})
        hbm, smhz, how = peaks()
        # compulsory bytes per SURVEY 8d K10/K12 (query + result per pass and active gas particle)
        for key, b, ms in (("roofline_density", (80.0 + 96.0) * n * rec["density_passes_mean"], rec["density_ms"]),
                           ("roofline_hydro", (136.0 + 40.0) * n, rec["hydro_ms"])):
            gbs = b / (ms * 1e-3) / 1e9
            rec[key] = {"bound": "hbm", "alg_bytes": b, "unit": "GB/s", "peak": hbm, "achieved": gbs, "frac": gbs / hbm,
                        "note": "neighbour gathers hit L1/L2; the pair kernels are bound by the fp64 pipe (ncu: profiles/), not HBM"}
    except Exception as ex:
        return {"failed": repr(ex)}
    if not no_cpu:
        try:    # the reference's own density.c / hydra.c on the host cores, 64^3 gas sample of the same recipe
            from oracle import ref as R
            r = R.Ref(nthreads=host_threads()) if R.available() else None
            if r is not None:
                ngs = 64
                p_t, m_t = ics.bench_ics("displaced", ngs, float(ngs), device="cpu")
                ps, ms_ = p_t.numpy(), m_t.numpy(); ns = len(ms_)
                vs = np.random.default_rng(1).standard_normal((ns, 3)) * 0.05
                t0 = time.perf_counter(); dc = r.sph_density(ps, ms_, float(ngs), np.full(ns, 3.0 * 0.8), vel=vs, kerneltype=2, mingashsml_frac=1e-4, DoEgyDensity=1)
                t1 = time.perf_counter(); r.sph_hydro(atime=0.1, hubble=3.0, dloga_bin=0.01, DensityIndependentSphOn=1); t2 = time.perf_counter()
                hws = dc["hsml"] * (1.0 + 0.01 * np.random.default_rng(2).standard_normal(ns))
                t3 = time.perf_counter(); r.sph_density(ps, ms_, float(ngs), hws, vel=vs, kerneltype=2, mingashsml_frac=1e-4, DoEgyDensity=1); t4 = time.perf_counter()
                rec["cpu_baseline"] = {"kind": "reference (density.c, hydra.c compiled unmodified)", "cores": host_threads(), "unit": "gas particles/s",
                                       "sample": "64^3 gas of the same recipe (%.1f s density, %.1f s hydro, %.1f s density from predicted Hsml)" % (t1 - t0, t2 - t1, t4 - t3),
                                       "density": ns / (t1 - t0), "hydro": ns / (t2 - t1), "value": ns / (t2 - t0),
                                       "density_predicted_hsml": ns / (t4 - t3), "value_predicted_hsml": ns / (t4 - t3 + t2 - t1)}
        except Exception as ex:
            rec["cpu_baseline"] = {"failed": repr(ex)}
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--ng", type=int, default=256, help="particles per dimension per GPU")
    ap.add_argument("--state", default="displaced", choices=["displaced", "z9", "clustered"], help="particle state of the timed box (ics.bench_ics)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU arm (cpu_baseline + parity of the 256^3 step)")
    ap.add_argument("--cpu-budget", type=float, default=240.0, help="seconds after which --impl reference stops adding timed steps")
    ap.add_argument("--no-hydro", action="store_true", help="skip the SPH density + hydro timing (configs[2])")
    ap.add_argument("--sharded-hydro", action="store_true", help="multi-GPU: add the sharded SPH entry (configs[4], gas part) to the line")
    ap.add_argument("--no-states", action="store_true", help="skip the extra z9 / clustered state entries")
    ap.add_argument("--no-steploop", action="store_true", help="skip the device-resident step-loop entry")
    ap.add_argument("--no-extras", action="store_true", help="skip the generic inverse PM pass and friends-of-friends probes (tools/c2r_probe.py, fof_probe.py)")
    ap.add_argument("--no-parity", action="store_true", help="multi-GPU: skip the sampled parity check")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world == 1 and not args.no_cpu:
        claim_host_threads()             # before numpy / the oracle libraries come in

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
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=240))

    W = max(args.warmup, 3)
    K = args.steps
    if world > 1:
        return run_sharded(args, rank, world, local, dist, pkg, ics, W, K)
    ng = args.ng
    box = float(ng)
    nmesh = ics.default_nmesh(ng)
    d_pos, d_mass = ics.bench_ics(args.state, ng, box, device="cuda")
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
        # gravpm_force + force_tree_full + grav_short_tree in one call
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
    # per-kernel durations for the roofline table: the same K steps issued as the three separate calls
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

    # ---- parity + CPU baseline: the step the GPU just did, redone by the CPU path on the same box ----------------------
    # GPU pass A leaves FullTreeGravAccel + GravPM; pass B (relative criterion fed by A) is the compared one.
    parity = None
    cpu_rec = None
    if not args.no_cpu:
        try:
            step_dev()
            oldacc_h = (d_acc + d_gpm).cpu().numpy()
            step_dev()
            acc_b = d_acc.cpu().numpy(); gpm_b = d_gpm.cpu().numpy()
            nthr = host_threads()
            cpu = CpuArm(nthr)
            cacc, cgpm, t_tree, t_pm = cpu.step(pos, mass, box, nmesh, par, oldacc_h)
            tot_c = cacc + cgpm
            norm = np.sqrt((tot_c ** 2).sum(1))
            mean = float(norm.mean())
            d = np.abs((acc_b + gpm_b) - tot_c).max(axis=1)
            parity = {"compared": "every particle of the %d^3 step: GPU (acc + GravPM) vs %s on the same inputs" % (ng, cpu.kind),
                      "acc_max_err_over_mean": float(d.max() / mean), "tree_max_err_over_mean": float(np.abs(acc_b - cacc).max() / mean),
                      "pm_max_err_over_max": float(np.abs(gpm_b - cgpm).max() / np.abs(cgpm).max()),
                      "per_particle_rel_p999": float(np.quantile(d / np.maximum(norm, 1e-300), 0.999)),
                      "tolerance": 1e-6, "ok": bool(d.max() / mean < 1e-6)}
            t = t_tree + t_pm
            cpu_rec = {"value": n / t, "unit": UNIT, "cores": nthr, "kind": cpu.kind, "same_config": True,
                       "sample": "one force step of the whole %d^3 / Nmesh %d workload on the same particles (%.1f s: tree %.1f s, PM %.1f s)"
                                 % (ng, nmesh, t, t_tree, t_pm)}
            del cacc, cgpm, tot_c, cpu
        except Exception as ex:      # the baseline is reported, never required
            cpu_rec = {"value": None, "unit": UNIT, "cores": host_threads(), "kind": "port", "sample": "failed: %r" % (ex,)}

    # ---- end-to-end arm: the reference's AoS in pinned host memory --------------
    P = np.zeros(n, dtype=pkg.PARTICLE_DTYPE)
    P["Pos"] = pos; P["Mass"] = mass; P["Type"] = 1; P["ID"] = np.arange(n)
    acc_h = d_acc.cpu().numpy(); gpm_h = d_gpm.cpu().numpy()
    P["FullTreeGravAccel"] = acc_h; P["GravPM"] = gpm_h
    pinned = torch.empty(n * 160, dtype=torch.uint8).pin_memory()
    pinned.numpy()[:] = P.view(np.uint8).reshape(-1)
    del P, d_pos, d_mass
    e2e_bytes = e.force_step_aos_bytes()
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
    # the AoS result against the device arm (same inputs up to the oldacc refresh)
    Pout = pinned.numpy().view(pkg.PARTICLE_DTYPE)
    chk = float(np.abs(Pout["GravPM"][:100000] - gpm_h[:100000]).max() / (np.abs(gpm_h[:100000]).max() + 1e-300))
    del pinned, Pout

    total = n * world
    value = total * K / (ms_dev * 1e-3)
    e2e_v = total * Ke / (ms_e2e * 1e-3)
    hbm, smhz, how = peaks()
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
    # fp64-pipe roof of the pair kernel ("interactions/s vs the fp64 FMA roofline"): 26 fp64-pipe instructions
    # per pair slot (3 sub + 3 fma r2, 5 rsqrt refinement, 5 mul, 1 sub + 2 sub + 2 fma window, 4 fma sums; SASS of k_grav_pairs)
    fp64_peak_slots = 148 * FP64_DFMA_PER_SM_CLK * smhz * 1e6 / 26.0
    N3 = float(nmesh) ** 3
    Mc = float(nmesh) ** 2 * (nmesh // 2 + 1)
    own_fft = e.pm_transform_kind() == 1
    kern = {
        "k_grav_walk": {"ms": walk_ms, "alg_bytes": walk_bytes},
        "k_grav_pairs": {"ms": pairs_ms, "alg_bytes": pairs_bytes, "pair_evaluations": pair_slots,
                         "Gpairs_per_s": pair_slots / (pairs_ms * 1e-3) / 1e9,
                         "fp64_pipe_frac": pair_slots / (pairs_ms * 1e-3) / fp64_peak_slots,
                         "fp64_pipe_note": "26 fp64-pipe instructions per pair slot against 148 SMs x 64 lanes x %.0f MHz" % smhz},
        "k_pm_deposit+clear": {"ms": phase["pm_deposit"], "alg_bytes": 156.0 * n + 8 * N3},
        # the transforms.  Own passes (csrc/pm_fft.cu): every pass reads and writes each value once -- z pass 8 B/cell real
        # side + 16 B/mode spectrum side, column passes 32 B/mode.  cuFFT: 16 B/cell per transform is the floor it is held to.
        **({"k_fft_z_forward+k_fft_columns<y>": {"ms": phase["pm_fft_forward"], "alg_bytes": 8 * N3 + 48 * Mc},
            "k_fft_columns<x forward, potential_transfer, x inverse>": {"ms": phase["pm_transfer"], "alg_bytes": 32 * Mc},
            "k_fft_columns<y inverse>+k_fft_z_inverse": {"ms": phase["pm_fft_inverse"], "alg_bytes": 8 * N3 + 48 * Mc}}
           if own_fft else
           {"cufft_d2z": {"ms": phase["pm_fft_forward"], "alg_bytes": 16 * N3},
            "k_pm_potential_transfer": {"ms": phase["pm_transfer"], "alg_bytes": 32 * Mc},
            "cufft_z2d": {"ms": phase["pm_fft_inverse"], "alg_bytes": 16 * N3}}),
        # difference + readout fused: the potential mesh is read once (8 B/cell) instead of the
        # 32 B/cell of the three force meshes, plus the particle side of the gather
        "k_pm_readout_fused": {"ms": phase["pm_gradient"] + phase["pm_readout"], "alg_bytes": 8 * N3 + (24.0 + 32.0) * n},
        "tree_build": {"ms": phase["tree_total"], "alg_bytes": (96 + 28 + 32) * float(n) + 80.0 * nn},
    }
    for k in kern.values():
        k["GBps"] = k["alg_bytes"] / (k["ms"] * 1e-3) / 1e9 if k["ms"] > 0 else None
        k["frac"] = k["GBps"] / hbm if k["GBps"] else None
    cfg = workload_config(ng, nmesh, args.state)
    cfg.update({"particles_per_gpu": n, "l2": "inputs larger than L2 (particle arrays %.0f MB, mesh %.1f GB)" % (n * 28 / 1e6, N3 * 8 / 1e9),
                "parallelism": "single GPU",
                "pm_transforms": "own shared-memory passes, Green's function fused (csrc/pm_fft.cu)" if own_fft else "cuFFT D2Z/Z2D + k_pm_potential_transfer"})
    # the dominant kernel of the step: whichever half of the short-range tree gravity took longer
    dom = max(("k_grav_walk", "k_grav_pairs"), key=lambda k: kern[k]["ms"])
    roofline = {"kernel": dom, "bound": "hbm", "achieved": kern[dom]["GBps"], "peak": hbm, "unit": "GB/s",
                "frac": kern[dom]["frac"], "traffic": NCU_TRAFFIC.get(dom), "peak_source": how,
                "note": "dominant kernel of the step by time.  Neither half of the tree gravity streams HBM: k_grav_walk is bound by "
                        "instruction issue (ncu: issue-active 70 %, integer / fp32 decision arithmetic, operands in shared memory and L2), "
                        "k_grav_pairs by the fp64 and conversion pipes (kernels.k_grav_pairs.fp64_pipe_frac).  bytes = SURVEY 8d compulsory "
                        "traffic of the kernel (targets + nodes + the piece lists it writes or reads); see kernels[] for the HBM-bound PM kernels"}
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": cfg,
        "e2e": {"value": e2e_v, "unit": UNIT, "h2d_bytes_per_step": n * e2e_bytes[0], "d2h_bytes_per_step": n * e2e_bytes[1],
                "ms_per_step": ms_e2e / Ke, "wall_ms_per_step": 1e3 * wall_e2e / Ke,
                "h2d_ms": te2e["h2d"], "d2h_ms": te2e["d2h"], "api": "b200_force_step_aos (the reference's 160-byte particle records in pinned host memory; %d B/particle in and %d B/particle "
                                                                     "out cross PCIe as strided copies)" % e2e_bytes,
                "check_vs_device_arm": chk},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "phases_ms": phase, "ms_per_step_serial": ms_serial,
        "kernels": kern,
        "tree": {"numnodes": nn, "maxdepth": int(info.maxdepth)},
        "parity": parity,
    }
    if cpu_rec is not None:
        out["cpu_baseline"] = cpu_rec
    if not args.no_states:
        # the other particle states of SURVEY 8d at the same size, device-resident step
        st = {}
        for state in ("z9", "clustered"):
            if state == args.state:
                continue
            try:
                st[state] = gpu_state_line(pkg, ics, e, state, ng, box, nmesh)
            except Exception as ex:
                st[state] = {"failed": repr(ex)}
        out["states"] = st
    if not args.no_hydro:
        out["hydro"] = hydro_entry(pkg, ics, e, pos, mass, ng, box, nmesh, par, no_cpu=args.no_cpu)
    if not args.no_steploop:
        # sub-steps of the hierarchical integrator (run.c:355-800) without the 160-byte record round trip
        try:
            SL = importlib.import_module("mp-gadget_b200.steploop")
            cosmo = ics.FlatLCDM()
            rng = np.random.default_rng(2)
            S = SL.StepEngine(e, cosmo.sync, cosmo.factor, cosmo.hubble, Omega0=cosmo.Omega0, Hubble=cosmo.Hubble, G=G)
            vel0 = 0.05 * rng.standard_normal((n, 3))
            for warm in (True, False):
                # one untimed PM sub-step first: it sizes the walk's piece pool and chunk tables for this particle set (the
                # entries above leave them sized for other trees, and a walk that outgrows them is repeated)
                S.set_particles(pos, mass, np.ones(n, np.uint8), box, vel=vel0)
                S.set_gravity(ics.tree_params(box, n, treeusebh=2), G, nmesh, 1.5)
                S.set_times(np.zeros(7, np.int64), np.zeros(47, np.int64), np.zeros(47, np.int64))
                if warm:
                    S.advance(first=True, pm=True)
            sub = []
            for k in range(7):
                S.prof = {}
                t0 = time.perf_counter(); bad, sinfo = S.advance(first=(k == 0), pm=True); dt = time.perf_counter() - t0
                sub.append({"wall_ms": 1e3 * dt, "active": int(sinfo[1]), "is_pm": int(sinfo[2]), "bad": bad,
                            "stages_ms": {a: round(b, 2) for a, b in S.prof.items()}})
            out["steploop"] = {"substeps": sub, "host_bytes_per_substep": "scalars only", "aos_roundtrip_bytes_per_force_call": 2 * 160 * n,
                               "what": "hierarchical KDK sub-steps (timestep.c:296-598) with the particle state resident in HBM; one untimed PM sub-step before"}
        except Exception as ex:
            out["steploop"] = {"failed": repr(ex)}
    if not args.no_extras:
        # SURVEY 8f ranks 3 and 4 at bench size, each in its own process (its own context; nothing it does can touch the line above)
        out["extras"] = {"pm_c2r_readout": run_probe("c2r_probe.py", [str(ng), str(nmesh)]), "fof_primary": run_probe("fof_probe.py", [str(ng), "clustered"])}
    print(json.dumps(out), flush=True)


def run_probe(script, argv, timeout=150):
    """The JSON line a tools/ probe prints, or what went wrong."""
    import subprocess
    try:
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", script)] + argv, capture_output=True, text=True, timeout=timeout)
        lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
        if r.returncode != 0 or not lines:
            return {"failed": (r.stderr or r.stdout)[-400:]}
        return json.loads(lines[-1])
    except Exception as ex:
        return {"failed": repr(ex)}


if __name__ == "__main__":
    main()
