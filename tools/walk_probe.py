"""Timing + result probe of the device-resident force step on a bench box (experiment builds: B200_LIB=...).
usage: walk_probe.py [ng] [state] [iters].  Prints per-phase ms, and sums of the walk counts / accelerations of
the last step so that two builds can be compared line by line."""
import sys, time, importlib, numpy as np, torch
sys.path.insert(0, ".")
pkg = importlib.import_module("mp-gadget_b200"); ics = importlib.import_module("mp-gadget_b200.ics")
ng = int(sys.argv[1]) if len(sys.argv) > 1 else 128
state = sys.argv[2] if len(sys.argv) > 2 else "displaced"
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 4
box = 1000.0 * ng; nmesh = ics.default_nmesh(ng)
d_pos, d_mass = ics.bench_ics(state, ng, box, device="cuda")
n = d_mass.shape[0]
e = pkg.Engine(0)
e.gravpm_init_periodic(box, 1.5, nmesh, 43.0071)
e.set_particles_dev(d_pos.data_ptr(), d_mass.data_ptr(), n)
par = ics.tree_params(box, n, treeusebh=1)
g = torch.empty((n, 3), dtype=torch.float64, device="cuda"); a = torch.empty_like(g); p = torch.empty(n, dtype=torch.float64, device="cuda")
for it in range(iters):
    t0 = time.perf_counter()
    e.force_step_dev(par, g.data_ptr(), a.data_ptr(), p.data_ptr()); e.oldacc_from_last_step()
    torch.cuda.synchronize()
    tm = e.timings()
    print("%s %d^3 it %d: step %.2f ms  pm %.2f tree %.2f walk %.2f pairs %.2f" % (state, ng, it, 1e3 * (time.perf_counter() - t0),
          tm.get("pm_total", 0), tm.get("tree_total", 0), tm.get("walk", 0), tm.get("walk_post", 0)), flush=True)
    par["TreeUseBH"] = 0
print("acc abs-sum %.15e pot sum %.15e gpm abs-sum %.15e" % (a.abs().sum().item(), p.sum().item(), g.abs().sum().item()))
if ng <= 128:
    e.force_tree_full(box)
    acc, pot, cnt = e.grav_short_tree(par, want_counts=True)
    print("counts acc %d open %d disc %d part %d" % tuple(int(cnt[f].astype(np.int64).sum()) for f in ("nodes_accepted", "nodes_opened", "nodes_discarded", "particles")))
e.close()
