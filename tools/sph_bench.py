"""SPH density + hydro timing on a gas Zel'dovich box (BASELINE.json configs[2] gas part)."""
import importlib, json, sys, time
import numpy as np
sys.path.insert(0, ".")
pkg = importlib.import_module("mp-gadget_b200"); ics = importlib.import_module("mp-gadget_b200.ics")
import torch
out = {}
e = pkg.Engine(0)
for ng in (128, 256):
    box = float(ng)
    pos_t, mass_t = ics.planewave_lattice(ng, box, device="cuda", seed=5)
    pos = pos_t.cpu().numpy(); mass = mass_t.cpu().numpy(); n = len(mass)
    del pos_t, mass_t
    rng = np.random.default_rng(1)
    vel = rng.standard_normal((n, 3)) * 0.05
    ent = np.ones(n)
    h0 = np.full(n, 3.0 * box / ng * 0.8)      # quintic support ~ 3 eta spacing
    sp = pkg.sph_params(KernelType=2, DensityIndependentSphOn=1, MinGasHsml=1e-4, atime=0.1, hubble=3.0, dloga_bin=0.01)
    e.set_particles(pos, mass, type=np.zeros(n, np.uint8))
    rec = {}
    for rep in range(3):
        t0 = time.time(); e.force_tree_build(box, mask=1); t1 = time.time()
        e.sph_set_gas(h0, vel=vel, entropy=ent)
        d = e.density(sp, update_hsml=1, DoEgyDensity=1); tm_d = e.timings()["sph_density"]
        h = e.hydro_force(sp); tm_h = e.timings()["sph_hydro"]
        rec = dict(n=n, tree_ms=e.timings()["tree_total"], density_ms=tm_d, hydro_ms=tm_h, niter_mean=float(d["niter"].mean()),
                   niter_max=int(d["niter"].max()), ngb_last_mean=float(d["ninteract"].mean()), hydro_cand_mean=float(h["ninteract"].mean()),
                   hsml_mean=float(d["hsml"].mean()), gas_per_s_density=n / (tm_d * 1e-3), gas_per_s_hydro=n / (tm_h * 1e-3))
    out["gpu_%d" % ng] = rec
    print(ng, rec, flush=True)
# CPU baseline: the reference's own density.c/hydra.c if built, else the oracle port, on 64^3
try:
    import oracle
    from oracle import ref as R
    ng = 64; box = float(ng)
    pos, mass = ics.zeldovich_lattice(ng, box, seed=5); n = len(mass)
    vel = np.random.default_rng(1).standard_normal((n, 3)) * 0.05
    h0 = np.full(n, 3.0 * 0.8)
    r = R.load()
    if r is not None:
        t0 = time.time(); rd = r.sph_density(pos, mass, box, h0, vel=vel, kerneltype=2, mingashsml_frac=1e-4, DoEgyDensity=1); t1 = time.time()
        rh = r.sph_hydro(atime=0.1, hubble=3.0, dloga_bin=0.01, DensityIndependentSphOn=1); t2 = time.time()
        import os
        out["cpu_reference_64"] = dict(n=n, density_s=t1 - t0, hydro_s=t2 - t1, cores=len(os.sched_getaffinity(0)),
                                        gas_per_s_density=n / (t1 - t0), gas_per_s_hydro=n / (t2 - t1))
        print(out["cpu_reference_64"])
except Exception as ex:
    print("cpu baseline failed", ex)
json.dump(out, open("gpurun_out/sph_bench.json", "w"), indent=1)
