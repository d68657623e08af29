"""SPH density + hydro on a displaced gas box: one run for ncu launch lists / captures.  usage: sph_prof.py [ng] [reps]"""
import importlib, sys, time
import numpy as np
sys.path.insert(0, ".")
pkg = importlib.import_module("mp-gadget_b200"); ics = importlib.import_module("mp-gadget_b200.ics")
ng = int(sys.argv[1]) if len(sys.argv) > 1 else 128
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
box = float(ng)
pos_t, mass_t = ics.bench_ics("displaced", ng, box, device="cuda")
pos = pos_t.cpu().numpy(); mass = mass_t.cpu().numpy(); n = len(mass)
rng = np.random.default_rng(1)
vel = rng.standard_normal((n, 3)) * 0.05
h0 = np.full(n, 3.0 * box / ng * 0.8)
sp = pkg.sph_params(KernelType=2, DensityIndependentSphOn=1, MinGasHsml=1e-4, atime=0.1, hubble=3.0, dloga_bin=0.01)
e = pkg.Engine(0)
e.set_particles(pos, mass, type=np.zeros(n, np.uint8))
for rep in range(reps):
    e.force_tree_build(box, mask=1)
    e.sph_set_gas(h0, vel=vel, entropy=np.ones(n))
    d = e.density(sp, update_hsml=1, DoEgyDensity=1); td = e.timings()["sph_density"]
    h = e.hydro_force(sp); th = e.timings()["sph_hydro"]
    print("cold: density %.2f ms (%.2f passes, max %d), hydro %.2f ms, ngb %.1f cand %.1f" % (td, d["niter"].mean(), d["niter"].max(), th, d["ninteract"].mean(), h["ninteract"].mean()), flush=True)
    # warm start: the converged lengths drifted by DtHsml over a step, as density() finds them in a run (drift.c:60-70)
    hw = d["hsml"] * (1.0 + 0.02 * rng.standard_normal(n))
    e.force_tree_build(box, mask=1)
    e.sph_set_gas(hw, vel=vel, entropy=np.ones(n))
    d2 = e.density(sp, update_hsml=1, DoEgyDensity=1); td2 = e.timings()["sph_density"]
    print("warm: density %.2f ms (%.2f passes, max %d)" % (td2, d2["niter"].mean(), d2["niter"].max()), flush=True)
