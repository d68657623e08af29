"""Primary friends-of-friends linking: oracle/oracle_fof.c against the reference's OWN fof.c
(tests/golden/ref_fof.npz, generator make_golden_fof.py; live where oracle/_ref/libref_domain.so exists), and the CUDA path
(csrc/fof.cu behind b200_fof_primary) against both -- under the CPU emulation of tests/emul here, on the GPU under the gpu
marker.  Labels are integers: every comparison is exact."""
import subprocess
import os
import sys
import numpy as np
import pytest

import oracle
from oracle import ref as R

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import domain_scenarios as DS        # noqa: E402

GOLD = np.load(os.path.join(HERE, "golden", "ref_fof.npz"))


def test_oracle_fof_primary_equals_reference():
    for k, (pos, ids, typ, box, ll) in enumerate(DS.fof_cases()):
        got = oracle.fof_primary(pos, ids, typ, box, ll)
        want = GOLD["%d/minid" % k]
        assert np.array_equal(got, want), k
        prim = typ == 1
        assert np.array_equal(got[~prim], ids[~prim])                       # other types are not linked
        assert got[-1] == got[-2] == min(ids[-1], ids[-2]) or got[-1] == got[-2]          # the pair across the periodic face
        # the label is the smallest ID of the group and is carried by one of its members
        for label in np.unique(got[prim])[:50]:
            m = prim & (got == label)
            assert ids[m].min() == label
    sizes = np.unique(GOLD["0/minid"], return_counts=True)[1]
    assert sizes.max() > 500 and (sizes == 1).sum() > 500                   # one big clump, many singles


@pytest.mark.skipif(not R.domain_available(), reason="oracle/_ref/libref_domain.so not built")
def test_oracle_fof_equals_reference_live():
    D = R.RefDomain(arena_gib=1.0, nthreads=2)
    if not hasattr(D.L, "ref_fof_primary"):
        pytest.skip("prebuilt libref_domain.so predates ref_fof_primary")
    rng = np.random.default_rng(77)
    n, box = 5000, 64.0
    pos = np.mod(32.0 + rng.standard_normal((n, 3)) * np.array([12.0, 3.0, 1.0]), box)
    ids = rng.permutation(n).astype(np.int64)
    typ = np.ones(n, np.uint8)
    for ll in (0.1, 0.4, 1.5):
        assert np.array_equal(oracle.fof_primary(pos, ids, typ, box, ll), D.fof_primary(pos, ids, typ, box, ll)), ll


def test_fof_source_under_emulation():
    """csrc/fof.cu, kernels and host driver unchanged, on the CPU stand-in: the reference's golden labels and the oracle on
    edge cases (one-cell grid, two link types, garbage, everything / nothing joins, groups across faces and corners)."""
    env = dict(os.environ, OMP_WAIT_POLICY="passive")
    r = subprocess.run([sys.executable, os.path.join(HERE, "emul", "run_fof_emul.py")], env=env, capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0 and "fof ok" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_fof_emulation_under_address_sanitizer():
    libs = [subprocess.run(["gcc", "-print-file-name=" + l], capture_output=True, text=True).stdout.strip() for l in ("libasan.so", "libubsan.so")]
    if not all(os.path.isabs(l) and os.path.exists(l) for l in libs):
        pytest.skip("libasan / libubsan not available")
    env = dict(os.environ, OMP_WAIT_POLICY="passive", EMUL_ASAN="1", LD_PRELOAD=" ".join(libs), ASAN_OPTIONS="detect_leaks=0")
    r = subprocess.run([sys.executable, os.path.join(HERE, "emul", "run_fof_emul.py")], env=env, capture_output=True, text=True, timeout=1800)
    assert r.returncode == 0 and "fof ok" in r.stdout, r.stdout[-2000:] + r.stderr[-6000:]


@pytest.mark.gpu
def test_gpu_fof_primary_equals_reference(engine, monkeypatch):
    """b200_fof_primary on the device: the reference's golden labels, the oracle on the edge cases, and a 200 000-particle
    clustered box (long chains of hooks, many threads per group) against the oracle."""
    for mode in ("clique", "cells"):      # the clique-cell search (default where the grid allows it) and the plain cell list
        monkeypatch.setenv("B200_FOF", mode)
        for k, (pos, ids, typ, box, ll) in enumerate(DS.fof_cases()):
            engine.set_particles(pos, np.ones(len(ids), np.float32), type=typ)
            got, ng = engine.fof_primary(ids, box, ll)
            assert np.array_equal(got, GOLD["%d/minid" % k]), (mode, k)
            assert ng == len(np.unique(got[typ == 1]))
        for pos, ids, typ, box, ll, mask, flags in DS.fof_edge_cases():
            if flags is not None:
                continue                       # the SoA entry point takes no garbage flags; covered under emulation
            engine.set_particles(pos, np.ones(len(ids), np.float32), type=typ)
            got, _ = engine.fof_primary(ids, box, ll, mask=mask)
            assert np.array_equal(got, oracle.fof_primary(pos, ids, typ, box, ll, mask=mask)), (mode, box, ll, mask)
    monkeypatch.delenv("B200_FOF")
    rng = np.random.default_rng(5)
    n, box = 200000, 100.0
    centres = rng.random((40, 3)) * box
    pos = np.mod(centres[rng.integers(0, 40, n)] + rng.standard_normal((n, 3)) * rng.choice([0.3, 1.0, 3.0], (n, 1)), box)
    pos[: n // 4] = rng.random((n // 4, 3)) * box
    ids = rng.permutation(n).astype(np.int64)
    typ = np.ones(n, np.uint8)
    ll = 0.2 * box / n ** (1 / 3)
    engine.set_particles(pos, np.ones(n, np.float32), type=typ)
    for rep in range(3):                   # the hooks race differently every time; the labels may not
        got, ng = engine.fof_primary(ids, box, ll)
        want = oracle.fof_primary(pos, ids, typ, box, ll)
        assert np.array_equal(got, want)
        assert ng == len(np.unique(want))
    sizes = np.unique(want, return_counts=True)[1]
    assert sizes.max() > 2000
