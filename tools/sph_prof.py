"""one density + hydro pass on a 128^3 gas box (for ncu launch lists)"""
import importlib, sys, numpy as np
sys.path.insert(0, ".")
pkg = importlib.import_module("mp-gadget_b200"); ics = importlib.import_module("mp-gadget_b200.ics")
ng = int(sys.argv[1]) if len(sys.argv) > 1 else 128
box = float(ng)
pos_t, mass_t = ics.planewave_lattice(ng, box, device="cuda", seed=5)
pos = pos_t.cpu().numpy(); mass = mass_t.cpu().numpy(); n = len(mass)
vel = np.random.default_rng(1).standard_normal((n, 3)) * 0.05
h0 = np.full(n, 3.0 * 0.8)
sp = pkg.sph_params(KernelType=2, DensityIndependentSphOn=1, MinGasHsml=1e-4, atime=0.1, hubble=3.0, dloga_bin=0.01)
e = pkg.Engine(0)
e.set_particles(pos, mass, type=np.zeros(n, np.uint8))
e.force_tree_build(box, mask=1)
e.sph_set_gas(h0, vel=vel, entropy=np.ones(n))
d = e.density(sp, update_hsml=1, DoEgyDensity=1)
h = e.hydro_force(sp)
print("density ms", e.timings()["sph_density"], "hydro ms", e.timings()["sph_hydro"])
