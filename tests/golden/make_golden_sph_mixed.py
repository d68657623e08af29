#!/usr/bin/env python3
"""Generate tests/golden/ref_sph_mixed.npz: a MIXED-TIME-BIN density + hydro step computed by
the reference's own density.c / hydra.c (oracle/_ref/libref_tree.so, single thread).

Starting from the synchronised state of the ref_sph.npz fixtures (quintic kernel,
pressure-entropy SPH), every particle is put on one of the time bins 2..5; at Ti_Current = 1000
bins 2 and 3 are active (is_timebin_active, timestep.c:143-150) and form the ActiveParticles
list, bins 4 and 5 are not: their Density / EgyWtDensity / DivVel / CurlVel / Hsml stay stale and
are drifted inside hydro (SPH_DensityPred, hydra.c:300-312).  The kick/drift factors per bin
(timefac.c integrals in the reference) are numbers chosen here and handed to the driver's
stand-ins (oracle/ref_driver.c); density.c and hydra.c themselves run unmodified.
Run in the build container:  make -C oracle ref && python tests/golden/make_golden_sph_mixed.py"""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref as R          # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
TI = 1000
HYDRO = dict(atime=0.5, hubble=0.2)


def tables():
    b = np.arange(48, dtype=np.float64)
    t = dict(gravkick=0.011 * b, hydrokick=0.017 * b, drift=0.023 * b, dloga_pred=0.0041 * b, dloga_bin=0.0013 * 2.0 ** (b - 2))
    t["gravkick"][47] = 0.05          # PM kick factor (FgravkickB)
    # The reference's hydro walk stores each finished particle's HydroAccel at once (hydro_reduce,
    # hydra.c:279-293) while later particles still read it through SPH_VelPred of their neighbours
    # (hydra.c:376-377): with a non-zero hydro kick factor on an ACTIVE bin its result depends on the
    # processing order (and races between threads).  The fixture keeps that factor zero on the active
    # bins so that the reference itself is deterministic; the inactive bins keep theirs.
    t["hydrokick"][2] = t["hydrokick"][3] = 0.0
    return t


def main():
    G = np.load(os.path.join(HERE, "ref_sph.npz"))
    r = R.load(nthreads=1)
    out = {}
    tb = tables()
    for k, v in tb.items():
        out["tables/" + k] = v
    out["Ti_Current"] = np.int64(TI)
    for name in ("clustered16", "zeldovich16"):
        g = lambda k: G[name + "/" + k]
        pos, mass, vel, ent, box, h0 = g("pos"), g("mass"), g("vel"), g("entropy"), float(g("box")), g("h0")
        n = len(mass)
        d0 = r.sph_density(pos, mass, box, h0, vel=vel, entropy=ent, kerneltype=2, init_hsml=False, DoEgyDensity=1)
        h0r = r.sph_hydro(atime=HYDRO["atime"], hubble=HYDRO["hubble"], dloga_bin=0.01, DensityIndependentSphOn=1)
        rng = np.random.default_rng(77 + n + len(name))
        bins = rng.integers(2, 6, n).astype(np.uint8)
        vel_new = vel + 0.05 * rng.standard_normal((n, 3))
        fullacc = 0.5 * rng.standard_normal((n, 3))
        act, m = r.sph_mixed(bins, TI, tb, HYDRO["atime"], HYDRO["hubble"], 1, vel=vel_new, fullacc=fullacc,
                             hydroacc=h0r["acc"], dtentropy=h0r["dtentropy"])
        key = name + "/"
        out[key + "bins"] = bins; out[key + "vel_new"] = vel_new; out[key + "fullacc"] = fullacc
        out[key + "active"] = act
        for k, v in d0.items():
            out[key + "sync_" + k] = v
        for k, v in h0r.items():
            out[key + "sync_hydro_" + k] = v
        for k, v in m.items():
            out[key + "mixed_" + k] = v
        print(name, "active", len(act), "of", n, "mean hsml", m["hsml"].mean())
    path = os.path.join(HERE, "ref_sph_mixed.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
