#!/usr/bin/env python3
"""Generate tests/golden/ref_pm.npz with the reference's OWN petapm.c, gravpm.c and
powerspectrum.c (oracle/_ref/libref_pm.so, compiled unmodified from /root/reference): the whole
gravpm_force of run.c:522 -- region selection from the reference's tree, CIC deposit, pencil
layout, potential_transfer, force_*_transfer, readout, power spectrum file -- on one rank.
The one third-party piece, PFFT (not available offline), is replaced by plain DFTs in PFFT's
single-rank layout (oracle/pfft_standin.c); every other line that runs is the reference's.
Run in the build container:  make -C oracle ref && python tests/golden/make_golden_pm.py"""
import importlib
import os
import sys
import tempfile
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref as R          # noqa: E402
ics = importlib.import_module("mp-gadget_b200.ics")
G = 43.0071


def cases():
    box = 16.0
    pos, mass = ics.zeldovich_lattice(16, box, seed=4, rms=0.4)
    yield "zeldovich16_n32", pos, mass, box, 32, 1.5
    bg = np.random.MT19937(); bg._legacy_seeding(4357)
    pos = ics.clustered_mix_from(bg, 4096, 8.0)
    rng = np.random.default_rng(9)
    yield "clustered16_n48", pos, (1 + rng.random(4096)).astype(np.float32), 8.0, 48, 1.25
    # a few particles exactly on cell boundaries, at the origin and just inside the far faces (CIC wrap-around)
    pos = np.array([[7.9999999, 7.9999999, 7.9999999], [0.0, 0.0, 0.0], [7.875, 0.0, 4.0], [3.999, 4.0, 4.001], [0.5, 7.75, 2.25],
                    [2.0, 2.0, 2.0], [6.0 + 1.0 / 3, 1.0 / 3, 5.0], [1.1, 6.9, 0.05], [4.4, 4.6, 7.7]])
    pos = np.concatenate([pos, np.random.default_rng(12).random((503, 3)) * 8.0])      # the reference's tree needs a few hundred
    yield "edges_n24", pos, np.ones(len(pos), np.float32), 8.0, 24, 1.5


def main():
    r = R.Ref(arena_gib=2.0, nthreads=1, so=R.SO_PM)
    out = {}
    for name, pos, mass, box, nmesh, asmth in cases():
        d = tempfile.mkdtemp()
        g, p = r.gravpm_force(pos, mass, box, nmesh, asmth, G, d, time=1.0)
        ps = np.loadtxt(os.path.join(d, "powerspectrum-1.0000.txt"))
        for k, v in (("pos", pos), ("mass", mass), ("box", np.float64(box)), ("nmesh", np.int64(nmesh)), ("asmth", np.float64(asmth)),
                     ("gravpm", g), ("potential", p), ("ps_k", ps[:, 0]), ("ps_P", ps[:, 1]), ("ps_N", ps[:, 2].astype(np.int64))):
            out[name + "/" + k] = v
        print(name, "max |GravPM|", np.abs(g).max(), "P(k) bins", len(ps))
    out["UnitLength_in_cm"] = np.float64(3.085678e21)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_pm.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
