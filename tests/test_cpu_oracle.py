"""CPU tests (no GPU): the oracle against the reference's own acceptance
bounds and structural invariants, and the C-ABI library's symbol table."""
import ctypes
import os
import re
import importlib
import numpy as np
import pytest

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = 43.0071


def _treepm(ics, pos, box, usebh=1):
    n = len(pos)
    mass = np.ones(n, dtype=np.float32)
    soft = (1 / 30.) * box / np.cbrt(n)
    gpm, _, _ = oracle.pm_force(pos, mass, box, 48, 1.5, G)
    tr = oracle.OracleTree(pos, mass, box)
    par = dict(ErrTolForceAcc=0.002, BHOpeningAngle=0.175, MaxBHOpeningAngle=0.0, TreeUseBH=usebh, Rcut=7.0,
               GravitySoftening=soft, rho0=1.0)
    acc, _, _ = tr.grav_short_tree(par, G, 48, 1.5)
    acc, _, _ = tr.grav_short_tree(par, G, 48, 1.5, oldacc=acc + gpm)     # "twice so the opening angle is consistent"
    return acc + gpm, mass, soft


def test_oracle_flat_lattice(ics):
    """tests/test_gravity.c:222-260: homogeneous lattice -> force ~ 0."""
    tot, _, _ = _treepm(ics, ics.lattice(16, 8.0), 8.0)
    assert np.abs(tot).max() < 0.015
    assert np.abs(tot).mean() < 0.005


@pytest.mark.parametrize("case", ["close", "random0", "random1"])
def test_oracle_vs_direct_sum(ics, case):
    """tests/test_gravity.c:146-160,262-318: TreePM vs periodic direct sum,
    max rel err < 3*ErrTolForceAcc, mean < 0.8*ErrTolForceAcc."""
    box = 8.0
    if case == "close":
        pos = ics.close_cluster(16)
    else:
        bg = np.random.MT19937()
        bg._legacy_seeding(4357)            # gsl_rng_mt19937, gsl_rng_set(r, 0)
        pos = ics.clustered_mix_from(bg, 16 ** 3, box)
        if case == "random1":
            pos = ics.clustered_mix_from(bg, 16 ** 3, box)
    tot, mass, soft = _treepm(ics, pos, box)
    ds = oracle.direct_sum(pos, mass, box, G, 2.8 * soft, repeat=1)
    err = np.abs(ds - tot) / np.abs(ds).mean()
    assert err.max() < 3 * 0.002
    assert err.mean() < 0.8 * 0.002


@pytest.mark.parametrize("kind", ["lattice", "close", "random"])
def test_oracle_tree_invariants(ics, kind):
    """tests/test_forcetree.c:33-114,119-171: every particle in exactly one
    leaf, child side = half the parent's, child centre on the right side,
    sibling threading, root mass = N, moments close."""
    box = 8.0
    if kind == "lattice":
        pos = ics.lattice(32, box)
    elif kind == "close":
        pos = ics.close_cluster(16)
    else:
        pos = np.random.default_rng(0).random((40000, 3)) * box
    n = len(pos)
    mass = np.ones(n, dtype=np.float32)
    t = oracle.OracleTree(pos, mass, box)
    nd = t.nodes
    leaves = nd["nocc"] >= 0
    parts = nd["part"][leaves].ravel()
    parts = parts[parts >= 0]
    assert len(parts) == n and len(np.unique(parts)) == n
    assert nd["nocc"][leaves].max() <= 8
    assert nd["mass"][0] == n
    kids = np.nonzero(nd["father"] >= 0)[0]
    fa = nd["father"][kids]
    assert np.array_equal(nd["len"][kids], 0.5 * nd["len"][fa])
    # child centre = parent +- len/4, rounded once (forcetree.c:305-320)
    off = np.abs(nd["center"][kids] - nd["center"][fa])
    assert np.allclose(off, 0.25 * nd["len"][fa][:, None], rtol=1e-9, atol=0)
    assert np.all(nd["nocc"][fa] == -1)
    # every particle lies inside its leaf (inside_node forcetree.c:289-298)
    li = np.nonzero(leaves)[0]
    for k in range(8):
        p = nd["part"][li, k]
        ok = p >= 0
        d = np.abs(2 * (pos[p[ok]] - nd["center"][li[ok]]))
        assert np.all(d <= nd["len"][li[ok]][:, None])
    # depth-first walk via sibling/firstchild visits every node once
    seen = 0
    no = 0
    while no >= 0:
        seen += 1
        no = nd["firstchild"][no] if nd["nocc"][no] < 0 else nd["sibling"][no]
    assert seen == len(nd)
    assert np.all((nd["cofm"] >= 0) & (nd["cofm"] <= box))


def test_oracle_pm_mass_conservation_and_cells(ics):
    box, nmesh = 8.0, 48
    pos = np.random.default_rng(1).random((5000, 3)) * box
    pos[0] = [box, box, box]
    mass = np.ones(len(pos), np.float32)
    g, p, ic, dens, pot = oracle.pm_force(pos, mass, box, nmesh, 1.5, G, return_mesh=True)
    assert abs(dens.sum() - len(pos)) < 1e-9
    assert np.array_equal(ic, np.floor(pos / (box / nmesh)).astype(np.int32))
    assert ic.max() == nmesh                     # Pos == BoxSize wraps (petapm.c:903-906)
    assert abs(pot.mean()) < 1e-9 * np.abs(pot).max()      # k=0 mode removed (gravpm.c:441-449)


def test_pm_point_mass_matches_shortrange_table(ics):
    """the long-range PM force of a point mass + the tabulated
    short-range window must add up to Newton (the table was calibrated on
    exactly this pipeline, tools/generate-force-kernels.py:93-127)."""
    box, nmesh = 48.0, 48            # cell = 1
    src = np.array([[24.3, 24.1, 23.8]])
    seps = np.array([0.7, 1.5, 2.5, 4.0, 6.0])
    rng = np.random.default_rng(3)
    errs = []
    for r in seps:
        dirs = rng.standard_normal((40, 3))
        dirs /= np.linalg.norm(dirs, axis=1)[:, None]
        pos = np.vstack([src, src + r * dirs])
        mass = np.zeros(len(pos), np.float32)
        mass[0] = 1.0
        mass[1:] = 1e-30
        g, _, _ = oracle.pm_force(pos, mass, box, nmesh, 1.5, G)
        radial = -(g[1:] * dirs).sum(1)                          # attraction towards src
        par = dict(ErrTolForceAcc=0.002, BHOpeningAngle=0.175, MaxBHOpeningAngle=0.9, TreeUseBH=1, Rcut=9.0,
                   GravitySoftening=1e-3, rho0=1.0)
        tr = oracle.OracleTree(pos, mass, box)
        acc, _, _ = tr.grav_short_tree(par, G, nmesh, 1.5)
        short = -(acc[1:] * dirs).sum(1)
        newton = G / r ** 2
        errs.append(np.abs((radial + short).mean() / newton - 1))
    assert max(errs) < 0.01, errs


def test_cabi_exports_every_declared_symbol(b200):
    hdr = open(os.path.join(ROOT, "include", "b200force.h")).read()
    declared = sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 20
    if not os.path.exists(b200.LIB_PATH):
        b200.build()
    L = ctypes.CDLL(b200.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), name
    assert sorted(b200.EXPORTED) == declared
    assert L.b200_abi_version() == 1
    lay = b200.ParticleLayout()
    L.b200_default_particle_layout(ctypes.byref(lay))
    dt = b200.PARTICLE_DTYPE
    assert lay.stride == dt.itemsize == 160
    assert lay.off_mass == dt.fields["Mass"][1] and lay.off_gravpm == dt.fields["GravPM"][1]
    assert lay.off_fulltreeacc == dt.fields["FullTreeGravAccel"][1] and lay.off_potential == dt.fields["Potential"][1]
    assert lay.off_type == dt.fields["Type"][1] and lay.off_hsml == dt.fields["Hsml"][1]


def test_no_cpu_fallback(b200):
    """Without a CUDA device the engine must refuse to start, loudly."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(b200.B200Error):
        b200.Engine(0)


def test_oracle_power_spectrum_sums(ics):
    """powerspectrum_add_mode restated: every non-zero mode of the half spectrum lands in a bin
    with weight 1 (kz = 0, N/2 planes) or 2, so sum(Nmodes) = N^3 - 1; Norm = (total mass)^2;
    a single plane wave puts its power in the bin of its wavenumber."""
    nmesh, box = 24, 24.0
    pos, mass = ics.zeldovich_lattice(12, box, seed=3)
    pw, kk, nm, norm = oracle.pm_power(pos, mass, box, nmesh)
    assert nm.sum() == nmesh ** 3 - 1
    assert abs(norm - float(mass.sum(dtype=np.float64)) ** 2) <= 1e-9 * norm
    assert np.all(pw[nm > 0] >= 0) and np.all(kk[nm > 0] > 0)
    # bin of a mode: floor(binsperunit * log(k2) / 2), binsperunit = (N-1)/log(sqrt(3) N/2)  (gravpm.c:341-342)
    bpu = (nmesh - 1) / np.log(np.sqrt(3) * nmesh / 2.0)
    assert nm[0] == 6          # k2 = 1: the six axis modes (+-x, +-y with weight 1 each... kz = +-1 folded with weight 2)
    assert int(np.floor(bpu * np.log(3.0) / 2)) < nmesh


def test_every_device_buffer_is_released():
    """Engine's grow-only device buffers have no destructor: each must be released by a destroy / release function."""
    src_dir = os.path.join(ROOT, "mp-gadget_b200", "csrc")
    hdr = open(os.path.join(src_dir, "engine.h")).read()
    names = [n.strip() for m in re.finditer(r"DevBuf<[^>]+>\s+([^;]+);", hdr) for n in m.group(1).split(",")]
    src = "".join(open(os.path.join(src_dir, f)).read() for f in os.listdir(src_dir) if f.endswith(".cu"))
    missing = [n for n in names if not re.search(r"\b%s\.release\(\)" % re.escape(n), src)]
    assert len(names) > 100 and not missing, missing
