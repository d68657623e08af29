"""PM step timing at one mesh size with the engine's own transform passes and with cuFFT (tools, not a bench line):
python tools/pm_probe.py [ng] [nmesh] ; B200_FFT_THREADS selects the block size of the passes."""
import importlib
import os
import sys
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("mp-gadget_b200")
ics = importlib.import_module("mp-gadget_b200.ics")
ng = int(sys.argv[1]) if len(sys.argv) > 1 else 256
nmesh = int(sys.argv[2]) if len(sys.argv) > 2 else 3 * ng
box = float(ng)
pos, mass = ics.planewave_lattice(ng, box, rms=1.0, device="cuda")
n = len(mass)
res = {}
only = os.environ.get("PM_PROBE_ONLY")
for sel in ((only,) if only else ("own", "cufft")):
    os.environ["B200_PM_FFT"] = sel
    e = pkg.Engine(0)
    e.set_particles_dev(pos.data_ptr(), mass.data_ptr(), n) if hasattr(e, "set_particles_dev") else e.set_particles(pos.cpu().numpy(), mass.cpu().numpy())
    e.gravpm_init_periodic(box, 1.5, nmesh, 43.0071)
    g = torch.zeros((n, 3), dtype=torch.float64, device="cuda")
    p = torch.zeros(n, dtype=torch.float64, device="cuda")
    for it in range(int(os.environ.get("PM_PROBE_ITERS", "4"))):
        e.gravpm_force_dev(g.data_ptr(), p.data_ptr())
        torch.cuda.synchronize()
    t = e.timings()
    print(sel, "kind", e.pm_transform_kind(), {k: round(v, 3) for k, v in t.items() if k.startswith("pm_")}, flush=True)
    res[sel] = (g.cpu().numpy(), p.cpu().numpy())
    e.close()
    del e
    torch.cuda.empty_cache()
if only:
    sys.exit(0)
g1, p1 = res["own"]
g0, p0 = res["cufft"]
print("own vs cufft: GravPM max diff / max %.3g   Potential %.3g" % (np.abs(g1 - g0).max() / np.abs(g0).max(), np.abs(p1 - p0).max() / np.abs(p0).max()))
