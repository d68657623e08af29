#!/usr/bin/env python3
"""Generate tests/golden/ref_sph_quartic.npz: the reference's own density.c / hydra.c / densitykernel.c with the QUARTIC
spline (DENSITY_KERNEL_QUARTIC_SPLINE = 4, densitykernel.h:17-21) on the zeldovich16 fixture of ref_sph.npz, both SPH
formulations.  Run in the build container:  make -C oracle ref && python tests/golden/make_golden_sph_quartic.py"""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref as R          # noqa: E402


def main():
    r = R.load(nthreads=1)
    g = np.load(os.path.join(HERE, "ref_sph.npz"))
    out = {}
    name = "zeldovich16"
    pos, mass, vel, ent, box, h0 = (g[name + "/" + k] for k in ("pos", "mass", "vel", "entropy", "box", "h0"))
    for DI in (0, 1):
        d = r.sph_density(pos, mass, float(box), h0, vel=vel, entropy=ent, kerneltype=4, init_hsml=False, DoEgyDensity=DI)
        h = r.sph_hydro(atime=0.5, hubble=0.2, dloga_bin=0.01, DensityIndependentSphOn=DI)
        key = "%s/k4_di%d/" % (name, DI)
        for k, v in d.items():
            out[key + k] = v
        for k, v in h.items():
            out[key + "hydro_" + k] = v
    path = os.path.join(HERE, "ref_sph_quartic.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
