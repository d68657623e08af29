"""Timing probe: the clustered state through the device-resident force step."""
import sys, time, importlib, numpy as np, torch
sys.path.insert(0, ".")
pkg = importlib.import_module("mp-gadget_b200"); ics = importlib.import_module("mp-gadget_b200.ics")
for ng in [int(a) for a in sys.argv[1:]] or [64, 128]:
    box = float(ng); nmesh = ics.default_nmesh(ng)
    d_pos, d_mass = ics.bench_ics("clustered", ng, box, device="cuda")
    n = d_mass.shape[0]
    e = pkg.Engine(0)
    e.gravpm_init_periodic(box, 1.5, nmesh, 43.0071)
    e.set_particles_dev(d_pos.data_ptr(), d_mass.data_ptr(), n)
    par = ics.tree_params(box, n, treeusebh=1)
    g = torch.empty((n, 3), dtype=torch.float64, device="cuda"); a = torch.empty_like(g); p = torch.empty(n, dtype=torch.float64, device="cuda")
    for it in range(4):
        t0 = time.perf_counter()
        e.force_step_dev(par, g.data_ptr(), a.data_ptr(), p.data_ptr()); e.oldacc_from_last_step()
        torch.cuda.synchronize()
        print(ng, it, "ms", 1e3 * (time.perf_counter() - t0), {k: round(v, 2) for k, v in e.timings().items() if v > 0.05}, flush=True)
        par["TreeUseBH"] = 0
    e.close()
