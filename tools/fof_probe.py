"""b200_fof_primary at bench size (tools; bench.py embeds its line under `extras`): 256^3 particles, linking length 0.2
mean spacings (the reference's FOFHaloLinkingLength); wall time of the call (host IDs in, labels out), size-independent
checks, and the reference's own fof.c (oracle/_ref/libref_domain.so, compiled unmodified) on a 64^3 sample of the same state
on the host cores: its labels against the GPU's on that sample, and its time."""
import importlib
import json
import os
import sys
import time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def sample(ics, state, ngs):
    box = float(ngs)
    pos_t, mass_t = ics.bench_ics(state, ngs, box, device="cpu")
    pos, mass = pos_t.numpy(), mass_t.numpy()
    ids = np.random.default_rng(1).permutation(len(mass)).astype(np.int64)
    return pos, mass, ids, box, 0.2 * box / ngs


def cpu_reference(ics, state, ngs):
    """(labels, record) of the reference's fof_label_primary on the sample, all host threads; None where _ref is absent"""
    from oracle import ref as R
    if not R.domain_available():
        return None, {"unavailable": "oracle/_ref/libref_domain.so not built"}
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    D = R.RefDomain(arena_gib=2.0, nthreads=cores)
    if not hasattr(D.L, "ref_fof_primary"):
        return None, {"unavailable": "prebuilt libref_domain.so predates ref_fof_primary"}
    pos, mass, ids, box, ll = sample(ics, state, ngs)
    t0 = time.perf_counter(); lab = D.fof_primary(pos, ids, np.ones(len(ids), np.uint8), box, ll); dt = time.perf_counter() - t0
    return lab, {"value": len(ids) / dt, "unit": "particles/s", "cores": cores, "kind": "reference (fof.c compiled unmodified)",
                 "sample": "%d^3 %s particles, ll = 0.2 spacings: %.1f s" % (ngs, state, dt)}


def main():
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
    rec = {"what": "b200_fof_primary, %d^3 %s particles, ll = 0.2 spacings, search: %s" % (ng, state, os.environ.get("B200_FOF", "default")),
           "n": n, "wall_ms": [round(x, 1) for x in out], "particles_per_s": n / (min(out) * 1e-3), "groups": int(ngrp),
           "largest_group": int(cnt.max()), "groups_of_32_or_more": int((cnt >= 32).sum()), "kernel_launches": e.kernel_launches(),
           "checks": "labels repeatable, label <= own ID, label owner carries its label, group count = distinct labels"}
    try:
        ngs = int(os.environ.get("FOF_PROBE_SAMPLE", "64"))
        ref_lab, cpu = cpu_reference(ics, state, ngs)
        if ref_lab is not None:
            spos, smass, sids, sbox, sll = sample(ics, state, ngs)
            e.set_particles(spos, smass)
            got, _ = e.fof_primary(sids, sbox, sll)
            cpu["labels_equal_gpu_on_sample"] = bool(np.array_equal(got, ref_lab))
        rec["cpu_baseline"] = cpu
    except Exception as ex:
        rec["cpu_baseline"] = {"failed": repr(ex)}
    print(json.dumps(rec))


if __name__ == "__main__":
    main()
