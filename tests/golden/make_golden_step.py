#!/usr/bin/env python3
"""Generate tests/golden/ref_step.npz with the reference's OWN drift.c, timestep.c and timebinmgr.c
(oracle/_ref/libref_step.so, compiled unmodified from /root/reference on top of its own tree
gravity): the integer timeline, drift_all_particles, build_active_particles / build_active_sublist,
apply_half_kick / apply_hydro_half_kick / apply_PM_half_kick / update_kick_times on a mixed
DM + gas set with garbage, and eight passes of the hierarchical KDK loop of run.c:355-800
(hierarchical_gravity_accelerations + hierarchical_gravity_and_timesteps) from the initial step, and six passes of
the SplitGravityTimestepsOn = 0 loop (force_tree_full + grav_short_tree for the active particles, find_timesteps), and ten
passes of the hierarchical loop with gas taking part (hydro half kicks, find_hydro_timesteps, Hsml prediction; hydro
accelerations and signal velocities held fixed), including hydro-only sub-steps and one where 5 of 174 active particles are
gravitationally active.
Stand-ins only for what needs GSL (flat matter + Lambda H(a), Gauss-Legendre kick integrals;
oracle/ref_driver.c).  Run in the build container:
    make -C oracle ref && python tests/golden/make_golden_step.py"""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import ref as R          # noqa: E402
import step_scenarios as SC          # noqa: E402


def main():
    S = R.RefStep(nthreads=2, arena_gib=1.0, **SC.TIMELINE)
    out = {"sync_loga": S.sync_loga}
    for k, v in S.cosmo.items():
        out["cosmo/" + k] = v
    for k, v in S.tspar.items():
        out["tspar/" + k] = v
    ti, dloga, span = SC.timeline_samples()
    out["tl/loga"] = np.array([S.loga_from_ti(int(t)) for t in ti])
    out["tl/dti"] = np.array([S.dti_from_dloga(float(d), int(t)) for t, d in zip(ti, dloga)], np.int64)
    out["tl/dloga"] = np.array([S.dloga_from_dti(12345, int(t)) for t in ti])
    out["tl/factor"] = np.array([[S.factor(k, int(t), int(t + s)) for k in range(3)] for t, s in zip(ti, span)])
    prim = SC.run_primitives(S, SC.primitives_inputs())
    for k, v in prim.items():
        out["prim/" + k] = v
    rec = SC.run_hierarchy(S, SC.hierarchy_inputs())
    for s, r in enumerate(rec):
        for k in ("bad", "info", "scal", "kick", "last", "bin_grav"):
            out["hier/%d/%s" % (s, k)] = r[k]
        if s in SC.HIER_KEEP:
            for k in ("pos", "vel", "fullacc"):
                out["hier/%d/%s" % (s, k)] = r[k]
    rec2 = SC.run_nonsplit(S, SC.hierarchy_inputs(seed=15, n=1536))
    for s, r in enumerate(rec2):
        for k in ("bad", "info", "scal", "kick", "last", "bin_grav"):
            out["nonsplit/%d/%s" % (s, k)] = r[k]
        if s in SC.NONSPLIT_KEEP:
            for k in ("pos", "vel", "fullacc"):
                out["nonsplit/%d/%s" % (s, k)] = r[k]
    rec3 = SC.run_gas_hierarchy(S, SC.gas_hierarchy_inputs())
    for s, r in enumerate(rec3):
        for k in ("bad", "info", "scal", "kick", "last", "bin_grav"):
            out["gas/%d/%s" % (s, k)] = r[k]
        if s in SC.GAS_KEEP:
            for k in ("pos", "vel", "fullacc", "hsml", "entropy"):
                out["gas/%d/%s" % (s, k)] = r[k]
    path = os.path.join(ROOT, "tests", "golden", "ref_step.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", "bins per step:", [np.bincount(r["bin_grav"])[24:].tolist() for r in rec][-1],
          "active:", [int(r["info"][0]) for r in rec], "nonsplit active:", [int(r["info"][0]) for r in rec2])


if __name__ == "__main__":
    main()
