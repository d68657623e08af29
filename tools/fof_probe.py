"""b200_fof_primary at bench size (tools, not a bench line): 256^3 displaced particles, linking length 0.2 mean spacings
(the reference's FOFHaloLinkingLength); wall time of the call (host ids in, labels out) and size-independent checks."""
import importlib
import json
import os
import sys
import time
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("mp-gadget_b200")
ics = importlib.import_module("mp-gadget_b200.ics")
ng = int(sys.argv[1]) if len(sys.argv) > 1 else 256
state = sys.argv[2] if len(sys.argv) > 2 else "clustered"
box = float(ng)
pos_t, mass_t = ics.bench_ics(state, ng, box, device="cuda")
pos, mass = pos_t.cpu().numpy(), mass_t.cpu().numpy()
n = len(mass)
ids = np.random.default_rng(1).permutation(n).astype(np.int64)
e = pkg.Engine(0)
e.set_particles(pos, mass)
ll = 0.2 * box / ng
out = []
for rep in range(2):
    t0 = time.perf_counter(); lab, ngrp = e.fof_primary(ids, box, ll); out.append(1e3 * (time.perf_counter() - t0))
    if rep == 0:
        first = lab
assert np.array_equal(first, lab)                                    # the hooks race, the labels do not
assert (lab <= ids).all()
where = np.empty(n, np.int64); where[ids] = np.arange(n)
assert np.array_equal(lab[where[lab]], lab)                          # the particle that owns a label carries it
uniq, cnt = np.unique(lab, return_counts=True)
assert len(uniq) == ngrp
print(json.dumps({"what": "b200_fof_primary, %d^3 %s particles, ll = 0.2 spacings, search: %s" % (ng, state, os.environ.get("B200_FOF", "default")), "n": n, "wall_ms": [round(x, 1) for x in out],
                  "particles_per_s": n / (min(out) * 1e-3), "groups": int(ngrp), "largest_group": int(cnt.max()),
                  "groups_of_32_or_more": int((cnt >= 32).sum()), "kernel_launches": e.kernel_launches(),
                  "checks": "labels repeatable, label <= own ID, label owner carries its label, group count = distinct labels"}))
