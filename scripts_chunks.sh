#!/bin/bash
for c in $@; do
B200_E2E_CHUNKS=$c timeout 600 python - <<PY
import importlib, numpy as np, torch, time
pkg = importlib.import_module("mp-gadget_b200"); ics = importlib.import_module("mp-gadget_b200.ics")
ng=256; box=float(ng); nmesh=ics.default_nmesh(ng)
d_pos,d_mass = ics.planewave_lattice(ng, box, device="cuda")
pos=d_pos.cpu().numpy(); mass=d_mass.cpu().numpy(); n=len(mass)
par = ics.tree_params(box, n, treeusebh=0)
e = pkg.Engine(0); e.gravpm_init_periodic(box, 1.5, nmesh, 43.0071)
P = np.zeros(n, dtype=pkg.PARTICLE_DTYPE); P["Pos"]=pos; P["Mass"]=mass; P["Type"]=1
P["FullTreeGravAccel"]=np.random.default_rng(1).standard_normal((n,3))*100
pinned = torch.empty(n*160, dtype=torch.uint8).pin_memory(); pinned.numpy()[:] = P.view(np.uint8).reshape(-1)
for it in range(4):
    torch.cuda.synchronize(); t0=time.perf_counter()
    e.force_step_aos(None, par, ptr=pinned.data_ptr(), n=n)
    t1=time.perf_counter(); t=e.timings()
print("chunks $c nocopy=$B200_E2E_NOCOPY wall %.1f ms h2d %.1f last-chunk walk %.2f pairs %.2f pm %.1f tree %.1f tail %.1f" % (1e3*(t1-t0), t["h2d"], t["walk"], t["walk_post"], t["pm_total"], t["tree_total"], t["d2h"]))
PY
done
