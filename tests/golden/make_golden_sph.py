#!/usr/bin/env python3
"""Generate tests/golden/ref_sph.npz with the reference's OWN density.c, hydra.c,
densitykernel.c, treewalk.c, forcetree.c (oracle/_ref/libref_tree.so, compiled
unmodified from /root/reference, single thread) on small gas-only fixtures set
up as tests/test_density.c:55-107 does (time bin 0, synchronised).
Run in the build container:  make -C oracle ref && python tests/golden/make_golden_sph.py"""
import importlib
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref as R          # noqa: E402
ics = importlib.import_module("mp-gadget_b200.ics")


def cases():
    box = 8.0
    yield "lattice16", ics.lattice(16, box), box, np.full(4096, 1.5 * box / 16)
    bg = np.random.MT19937(); bg._legacy_seeding(4357)
    yield "clustered16", ics.clustered_mix_from(bg, 4096, box), box, np.full(4096, box / 16)
    pos, _ = ics.zeldovich_lattice(16, box, seed=11, rms=0.3)
    yield "zeldovich16", pos, box, np.full(4096, 1.2 * box / 16)


def main():
    r = R.load(nthreads=1)
    out = {}
    for name, pos, box, h0 in cases():
        n = len(pos)
        rng = np.random.default_rng(n + len(name))
        mass = (1 + 0.2 * rng.random(n)).astype(np.float32)
        vel = rng.standard_normal((n, 3)) * 0.3
        ent = 1 + 0.5 * rng.random(n)
        out[name + "/pos"] = pos; out[name + "/mass"] = mass; out[name + "/vel"] = vel
        out[name + "/entropy"] = ent; out[name + "/box"] = np.float64(box); out[name + "/h0"] = h0
        for kt in (1, 2):
            for DI in (0, 1):
                d = r.sph_density(pos, mass, box, h0, vel=vel, entropy=ent, kerneltype=kt, init_hsml=False, DoEgyDensity=DI)
                h = r.sph_hydro(atime=0.5, hubble=0.2, dloga_bin=0.01, DensityIndependentSphOn=DI)
                key = "%s/k%d_di%d/" % (name, kt, DI)
                for k, v in d.items():
                    out[key + k] = v
                for k, v in h.items():
                    out[key + "hydro_" + k] = v
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_sph.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
