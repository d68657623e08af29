import importlib, os, sys, numpy as np
sys.path.insert(0, ".")
b200 = importlib.import_module("mp-gadget_b200")
GOLD = np.load("tests/golden/ref_sph.npz"); MIXED = np.load("tests/golden/ref_sph_mixed.npz")
name = "clustered16"
g = lambda k: GOLD[name + "/" + k]; m = lambda k: MIXED[name + "/" + k]
pos, mass, vel, ent, box = g("pos"), g("mass"), g("vel"), g("entropy"), float(g("box"))
n = len(mass)
tb = {k: MIXED["tables/" + k] for k in ("gravkick", "hydrokick", "drift", "dloga_pred", "dloga_bin")}
bins, act = m("bins"), m("active"); Ti = int(MIXED["Ti_Current"])
active_bin = np.array([b <= 0 or Ti % (1 << b) == 0 for b in range(47)])
for variant in ("asis", "asis2", "asis3"):
    tabs = dict(gravkick=tb["gravkick"][:47], hydrokick=tb["hydrokick"][:47], dloga_pred=tb["dloga_pred"][:47],
                drift=np.where(active_bin, 0.0, tb["drift"][:47]), dloga_bin=tb["dloga_bin"][:47])
    if variant == "nodrift": tabs["drift"] = np.zeros(47)
    e = b200.Engine(0)
    e.set_particles(pos, mass, type=np.zeros(n, np.uint8)); e.force_tree_build(box, mask=1)
    e.sph_set_gas(m("sync_hsml"), vel=m("vel_new"), entropy=ent, dtentropy=m("sync_hydro_dtentropy"), fullacc=m("fullacc"), hydroacc=m("sync_hydro_acc"))
    e.sph_set_timebins(bins, bins, tabs); e.sph_set_active(act)
    e.sph_set_state(density=m("sync_density"), egywtdensity=m("sync_egywtdensity"), dhsmlfac=m("sync_dhsmlfac"), divvel=m("sync_divvel"), curlvel=m("sync_curlvel"))
    sp = b200.sph_params(KernelType=2, MinGasHsml=0.006, DensityIndependentSphOn=1, atime=0.5, hubble=0.2, pmkick=float(tb["gravkick"][47]))
    d = e.density(sp, update_hsml=1, DoEgyDensity=1); h = e.hydro_force(sp)
    ra, ga = m("mixed_acc")[act], h["acc"][act]
    rel = np.abs(ga - ra).max(1) / (np.abs(ra).max(1) + 1e-300)
    bad = rel > 1e-9
    print(variant, "bad", bad.sum(), "of", len(act), "median rel of bad", np.median(rel[bad]) if bad.any() else 0,
          "maxsig bad", (np.abs(h["maxsignalvel"][act] - m("mixed_maxsignalvel")[act]) > 1e-9 * np.abs(m("mixed_maxsignalvel")[act])).sum(),
          "dte bad", (np.abs(h["dtentropy"][act] - m("mixed_dtentropy")[act]) > 1e-9 * np.abs(m("mixed_dtentropy")[act]) + 1e-300).sum())
    print("  bins of bad:", np.bincount(bins[act][bad], minlength=6), " of good:", np.bincount(bins[act][~bad], minlength=6))
    print("  ncand gpu mean", h["ninteract"][act].mean())
    e.close()
