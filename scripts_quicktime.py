import importlib, sys, time
import numpy as np
sys.path.insert(0, ".")
pkg = importlib.import_module("mp-gadget_b200"); ics = importlib.import_module("mp-gadget_b200.ics")
G = 43.0071
e = pkg.Engine(0)
for ng in (64, 128, 256):
    box = float(ng); nmesh = ics.default_nmesh(ng)
    t0 = time.time(); pos, mass = ics.zeldovich_lattice(ng, box); n = len(mass)
    print("ng", ng, "ics s", time.time() - t0, flush=True)
    par = ics.tree_params(box, n, treeusebh=1)
    e.set_particles(pos, mass)
    e.gravpm_init_periodic(box, 1.5, nmesh, G)
    for it in range(3):
        t0 = time.time(); gpm, _ = e.gravpm_force(want_potential=False); t1 = time.time()
        info = e.force_tree_full(box); t2 = time.time()
        par["TreeUseBH"] = 1 if it == 0 else 0
        acc, pot, cnt = e.grav_short_tree(par, want_counts=(it == 2)); t3 = time.time()
        e.oldacc_from_last_step()
        tm = e.timings()
        print("it", it, "wall pm %.3f tree %.3f walk %.3f" % (t1 - t0, t2 - t1, t3 - t2), "nodes", info.numnodes, "depth", info.maxdepth, flush=True)
        print("   ", {k: round(v, 3) for k, v in tm.items()}, flush=True)
    print("   counts mean: acc %.1f open %.1f disc %.1f part %.1f" % tuple(cnt[f].mean() for f in ("nodes_accepted", "nodes_opened", "nodes_discarded", "particles")))
    print("   |acc| mean", np.abs(acc).mean(), "|gpm| mean", np.abs(gpm).mean(), "sum m*a / sum|m a|", np.abs((acc+gpm).sum(0)).max() / np.abs(acc+gpm).sum())
