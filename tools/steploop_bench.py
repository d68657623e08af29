"""Step-loop timing on a GPU box (not yet run: written after round 1's GPU minutes were spent).
Hierarchical KDK sub-steps with the particle state resident in HBM (mp-gadget_b200/steploop.py) on a
plane-wave-displaced 128^3 / 256^3 DM box: wall time per sub-step against the number of
gravitationally active particles, the host<->device bytes per sub-step (scalars only), and for comparison
the 160-byte-record round trip the host-resident loop would pay per force call (b200_force_step_aos).
The cosmology callables (what the reference host takes from cosmology.c / timefac.c) are ics.FlatLCDM.
usage: python tools/steploop_bench.py [ng ...]   -> gpurun_out/steploop_bench.json"""
import importlib
import json
import os
import sys
import time
import numpy as np

sys.path.insert(0, "."); sys.path.insert(0, "tests")
pkg = importlib.import_module("mp-gadget_b200"); ics = importlib.import_module("mp-gadget_b200.ics")
SL = importlib.import_module("mp-gadget_b200.steploop")
G = 43.0071
cosmo = ics.FlatLCDM()
sync = cosmo.sync
out = {}
for ng in [int(a) for a in sys.argv[1:]] or [128, 256]:
    box = 1000.0 * ng
    pos_t, mass_t = ics.planewave_lattice(ng, box, device="cuda", seed=5)
    pos = pos_t.cpu().numpy(); mass = mass_t.cpu().numpy(); n = len(mass)
    del pos_t, mass_t
    rng = np.random.default_rng(2)
    vel = 30.0 * rng.standard_normal((n, 3))
    e = pkg.Engine(0)
    S = SL.StepEngine(e, sync, cosmo.factor, cosmo.hubble, Omega0=cosmo.Omega0, Hubble=cosmo.Hubble, G=G)
    S.set_particles(pos, mass, np.ones(n, np.uint8), box, vel=vel)
    S.set_gravity(ics.tree_params(box, n, treeusebh=2), G, 3 * ng, 1.5)
    S.set_times(np.zeros(7, np.int64), np.zeros(47, np.int64), np.zeros(47, np.int64))
    steps = []
    for s in range(9):
        l0 = e.kernel_launches() if hasattr(e, "kernel_launches") else 0
        t0 = time.perf_counter()
        bad, info = S.advance(first=(s == 0), pm=True)
        dt = time.perf_counter() - t0
        scal = S.get_times()[0]
        steps.append(dict(step=s, wall_ms=1e3 * dt, active=int(info[1]), is_pm=int(info[2]), mintimebin=int(scal[0]), maxtimebin=int(scal[1]), bad=bad))
        print(ng, steps[-1], flush=True)
    bins = np.bincount(S.get()["bin_grav"], minlength=47)
    out["ng%d" % ng] = dict(n=n, steps=steps, bins={int(b): int(c) for b, c in enumerate(bins) if c},
                            aos_roundtrip_bytes_per_call=2 * 160 * n)
    e.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/steploop_bench.json", "w"), indent=1)
