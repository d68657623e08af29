#!/usr/bin/env python3
"""Generate tests/golden/ref_pm_c2r.npz with the reference's OWN petapm.c (oracle/_ref/libref_pm.so, compiled unmodified
from /root/reference): petapm_force_init + petapm_force_c2r + petapm_force_finish driven as MP-GenIC's displacement_fields
drives them (libgenic/zeldovich.c:150-229) -- a source spectrum handed in, transfer functions of the density_transfer /
disp_transfer forms (:276-313) whose k-dependent factor is read from a table by the integer k2, CIC read-outs -- on one rank,
with PFFT replaced by plain DFTs in its single-rank layout (oracle/pfft_standin.c).
Run in the build container:  make -C oracle ref && python tests/golden/make_golden_pm_c2r.py"""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import ref as R          # noqa: E402
import pm_c2r_scenarios as SC        # noqa: E402


def main():
    r = R.Ref(arena_gib=2.0, nthreads=1, so=R.SO_PM)
    out = {}
    for name, pos, box, nmesh, rho_k, functions in SC.cases():
        res = r.petapm_c2r(pos, box, nmesh, rho_k, functions)
        for j, a in enumerate(res):
            out["%s/out%d" % (name, j)] = a
        print(name, [float(np.abs(a).max()) for a in res])
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_pm_c2r.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
