"""b200_pm_c2r_readout at bench size (tools; bench.py embeds its line under `extras`): MP-GenIC's seven functions (Density,
DispX/Y/Z, VelX/Y/Z; libgenic/zeldovich.c:181-190) on a 768^3 mesh read out at 256^3 particles.  Wall time of the call -- the
3.6 GB source spectrum and the tables in from pageable host memory, seven read-out arrays back -- and a linearity check (the
velocity table is 0.7 x the displacement table, so must the read-outs be)."""
import importlib
import json
import os
import sys
import time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
pkg = importlib.import_module("mp-gadget_b200")
ics = importlib.import_module("mp-gadget_b200.ics")
import pm_c2r_scenarios as SC        # noqa: E402

ng = int(sys.argv[1]) if len(sys.argv) > 1 else 256
nmesh = int(sys.argv[2]) if len(sys.argv) > 2 else 3 * ng
box = float(ng)
pos_t, mass_t = ics.bench_ics("displaced", ng, box, device="cuda")
pos, mass = pos_t.cpu().numpy(), mass_t.cpu().numpy()
del pos_t, mass_t
n = len(mass)
e = pkg.Engine(0)
e.set_particles(pos, mass)
e.gravpm_init_periodic(box, 1.5, nmesh, 43.0071)
assert e.pm_transform_kind() == 1
rk = np.empty((nmesh, nmesh, nmesh // 2 + 1), np.complex128)
rk[:] = 1.0                                   # a point source at the origin: every mode present, cheap to make
rk[1::2, :, :] *= -1.0                        # ... moved half a cell along x
dens, disp, vel = SC.genic_tables(nmesh, box, growth=0.7)
fn = [(0, dens), (1, disp), (2, disp), (3, disp), (1, vel), (2, vel), (3, vel)]
ms = []
for rep in range(2):
    t0 = time.perf_counter(); res = e.pm_c2r_readout(rk, fn); ms.append(1e3 * (time.perf_counter() - t0))
tm = e.timings()
assert all(np.isfinite(a).all() for a in res)
lin = max(float(np.abs(res[4 + k] - 0.7 * res[1 + k]).max() / (np.abs(res[1 + k]).max() + 1e-300)) for k in range(3))
assert lin < 1e-12, lin
# the CPU restatement (oracle.pm_c2r_readout: numpy transfer, scipy irfftn, C read-out -- pinned to the reference's own petapm.c
# by tests/golden/ref_pm_c2r.npz) on a 1/8 sample of the same workload, all host threads: its time, and the GPU against it
cpu = None
try:
    import oracle
    ngs, nms = ng // 2, nmesh // 2
    sbox = float(ngs)
    sp_t, sm_t = ics.bench_ics("displaced", ngs, sbox, device="cpu")
    spos, smass = sp_t.numpy(), sm_t.numpy()
    srk = np.empty((nms, nms, nms // 2 + 1), np.complex128)
    srk[:] = 1.0
    srk[1::2, :, :] *= -1.0
    sd, sdi, sv = SC.genic_tables(nms, sbox, growth=0.7)
    sfn = [(0, sd), (1, sdi), (2, sdi), (3, sdi), (1, sv), (2, sv), (3, sv)]
    t0 = time.perf_counter(); want = oracle.pm_c2r_readout(spos, sbox, nms, srk, sfn); dt = time.perf_counter() - t0
    e.set_particles(spos, smass)
    e.gravpm_init_periodic(sbox, 1.5, nms, 43.0071)
    got = e.pm_c2r_readout(srk, sfn)
    err = max(float(np.abs(a - b).max() / np.abs(b).max()) for a, b in zip(got, want))
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    cpu = {"value": 7 * float(nms) ** 3 / dt, "unit": "mesh cells x functions / s", "cores": cores,
           "kind": "port (oracle.pm_c2r_readout: scipy irfftn + C read-out)", "sample": "Nmesh %d, %d^3 particles, 7 functions: %.1f s" % (nms, ngs, dt),
           "gpu_vs_port_max_rel_on_sample": err, "tolerance": 1e-10, "ok": bool(err < 1e-10)}
except Exception as ex:
    cpu = {"failed": repr(ex)}
print(json.dumps({"what": "b200_pm_c2r_readout: 7 functions (MP-GenIC displacement_fields set), Nmesh %d, %d^3 particles" % (nmesh, ng),
                  "wall_ms": [round(x, 1) for x in ms], "wall_ms_per_function": round(min(ms) / 7, 1),
                  "device_ms_last_function": {"transfer": round(tm["pm_transfer"], 3), "inverse_passes": round(tm["pm_fft_inverse"], 3), "readout": round(tm["pm_readout"], 3)},
                  "host_bytes_in": int(rk.nbytes + 7 * dens.nbytes), "host_bytes_out": int(7 * n * 8),
                  "cells_x_functions_per_s": 7 * float(nmesh) ** 3 / (min(ms) * 1e-3), "cpu_baseline": cpu,
                  "kernel_launches": e.kernel_launches(), "linearity_max_rel": lin,
                  "checks": "read-outs finite; velocity read-outs = 0.7 x displacement read-outs (tables differ by that factor)"}))
