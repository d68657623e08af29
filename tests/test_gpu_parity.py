"""GPU parity tests: CUDA path (through the C-ABI) vs the CPU oracle.

Bars (BASELINE.json north_star): PM cell indices, tree topology, tree opening
and interaction counts bit-exact; fp64 accelerations within 1e-6 relative.
"""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

G = 43.0071
ACC_RTOL = 1e-6          # north_star tolerance for fp64 accelerations


def _distributions(ics):
    box = 8.0
    rng = np.random.default_rng(42)
    out = {
        "lattice16": (ics.lattice(16, box), box),
        "close16": (ics.close_cluster(16), box),
        "gslrandom16": (ics.clustered_mix(16 ** 3, box, seed=0), box),
        "uniform20k": (rng.random((20000, 3)) * box, box),
    }
    p, _ = ics.zeldovich_lattice(32, 32.0)
    out["zeldovich32"] = (p, 32.0)
    return out


def _relerr(a, b):
    scale = np.sqrt((b ** 2).sum(axis=1)).mean() + 1e-300
    return np.abs(a - b).max() / scale


@pytest.mark.parametrize("name", ["lattice16", "close16", "gslrandom16", "uniform20k", "zeldovich32"])
def test_pm_parity(engine, ics, name):
    pos, box = _distributions(ics)[name]
    n = len(pos)
    mass = np.ones(n, dtype=np.float32)
    nmesh, asmth = 48, 1.5
    engine.set_particles(pos, mass)
    engine.gravpm_init_periodic(box, asmth, nmesh, G)
    g, p = engine.gravpm_force()
    og, op, oic, odens, opot = oracle.pm_force(pos, mass, box, nmesh, asmth, G, return_mesh=True)
    # PM indexing bit-exact (petapm.c:976-980)
    assert np.array_equal(engine.pm_cell_index(), oic)
    # potential mesh and the readouts
    potmesh = engine.pm_copy_mesh(1)
    assert np.abs(potmesh - opot).max() <= 1e-11 * np.abs(opot).max()
    dens = engine.pm_copy_mesh(0)
    assert np.abs(dens - odens).max() <= 1e-13 * odens.max()
    assert abs(dens.sum() - mass.sum(dtype=np.float64)) <= 1e-9 * n      # verify_density_field petapm.c:1059-1089
    assert np.abs(p - op).max() <= 1e-10 * np.abs(op).max()
    # forces: real-space 4-point difference vs the reference's k-space filter
    # (scale floor: on the exact lattice the true force is zero and both sides are rounding noise)
    scale = max(np.abs(og).max(), 1e-3 * G * float(mass.max()) / (box / nmesh) ** 2)
    assert np.abs(g - og).max() <= 1e-9 * scale, (np.abs(g - og).max(), scale)


@pytest.mark.parametrize("name", ["lattice16", "close16", "gslrandom16", "uniform20k", "zeldovich32"])
@pytest.mark.parametrize("topdepth", [0, 2])
def test_tree_parity(engine, ics, name, topdepth):
    pos, box = _distributions(ics)[name]
    n = len(pos)
    rng = np.random.default_rng(7)
    mass = (1.0 + rng.random(n)).astype(np.float32)
    engine.set_particles(pos, mass)
    info = engine.force_tree_build(box, toplevel_depth=topdepth)
    ot = oracle.OracleTree(pos, mass, box, toplevel_depth=topdepth)
    assert info.numnodes == ot.t.numnodes
    assert info.numparticles == n
    assert info.overfull_leaves == 0
    t = engine.tree_export()
    on = ot.nodes
    assert np.array_equal(t["len"], on["len"])
    assert np.array_equal(t["center"], on["center"])
    assert np.array_equal(t["sibling"], on["sibling"])
    assert np.array_equal(t["firstchild"], on["firstchild"])
    assert np.array_equal(t["nocc"], on["nocc"])
    assert np.array_equal(t["part"], on["part"])
    assert np.array_equal(t["mass"], on["mass"])
    assert np.array_equal(t["cofm"], on["cofm"])
    assert info.root_mass == on["mass"][0]


def _walk_case(engine, ics, pos, box, nmesh, par, oldacc=None, active=None):
    n = len(pos)
    mass = np.ones(n, dtype=np.float32)
    engine.set_particles(pos, mass, oldacc=oldacc)
    engine.gravpm_init_periodic(box, 1.5, nmesh, G)
    engine.force_tree_build(box, active=active)
    acc, pot, cnt = engine.grav_short_tree(par, active=active, want_counts=True)
    ot = oracle.OracleTree(pos, mass, box, active=active)
    oacc, opot, ocnt = ot.grav_short_tree(par, G, nmesh, 1.5, oldacc=oldacc, active=active, full=active is None)
    return acc, pot, cnt, oacc, opot, ocnt


@pytest.mark.parametrize("name", ["lattice16", "close16", "gslrandom16", "uniform20k", "zeldovich32"])
@pytest.mark.parametrize("usebh", [1, 0])
def test_walk_parity(engine, ics, name, usebh):
    pos, box = _distributions(ics)[name]
    n = len(pos)
    par = ics.tree_params(box, n, treeusebh=usebh, rcut=7.0)
    oldacc = None
    if usebh == 0:
        rng = np.random.default_rng(3)
        oldacc = rng.standard_normal((n, 3)) * 500.0
    acc, pot, cnt, oacc, opot, ocnt = _walk_case(engine, ics, pos, box, 48, par, oldacc=oldacc)
    for f in ("nodes_accepted", "nodes_opened", "nodes_discarded", "particles"):
        assert np.array_equal(cnt[f], ocnt[f]), f
    assert _relerr(acc, oacc) < ACC_RTOL
    assert np.abs(pot - opot).max() <= ACC_RTOL * np.abs(opot).max()


def test_walk_active_subset(engine, ics):
    """Tree of active particles only, walked for the same subset
    (force_tree_active_moments + grav_short_tree, timestep.c:282-290)."""
    pos, box = _distributions(ics)["gslrandom16"]
    n = len(pos)
    rng = np.random.default_rng(11)
    active = np.sort(rng.choice(n, size=n // 3, replace=False)).astype(np.int32)
    par = ics.tree_params(box, n, treeusebh=1, rcut=7.0)
    acc, pot, cnt, oacc, opot, ocnt = _walk_case(engine, ics, pos, box, 48, par, active=active)
    assert np.array_equal(cnt["particles"][active], ocnt["particles"][active])
    assert np.array_equal(cnt["nodes_accepted"][active], ocnt["nodes_accepted"][active])
    assert _relerr(acc[active], oacc[active]) < ACC_RTOL
    inactive = np.setdiff1d(np.arange(n), active)
    assert np.all(acc[inactive] == 0)


@pytest.mark.parametrize("name,direct", [("lattice16", False), ("close16", True), ("gslrandom16", True)])
def test_reference_gravity_bounds(engine, ics, name, direct):
    """The reference's own acceptance test for TreePM (tests/test_gravity.c:146-160,
    222-318) run through the CUDA path: Nmesh 48, Asmth 1.5, BH angle 0.175,
    Rcut 7, softening 1/30, two tree passes."""
    pos, box = _distributions(ics)[name]
    n = len(pos)
    mass = np.ones(n, dtype=np.float32)
    errtol = 0.002
    par = dict(ErrTolForceAcc=errtol, BHOpeningAngle=0.175, MaxBHOpeningAngle=0.0, TreeUseBH=1, Rcut=7.0,
               GravitySoftening=(1 / 30.) * box / np.cbrt(n), rho0=1.0)
    engine.set_particles(pos, mass)
    engine.gravpm_init_periodic(box, 1.5, 48, G)
    gpm, _ = engine.gravpm_force()
    engine.force_tree_full(box)
    acc, _, _ = engine.grav_short_tree(par)
    engine.oldacc_from_last_step()
    acc, _, _ = engine.grav_short_tree(par)
    tot = acc + gpm
    if not direct:
        assert np.abs(tot).max() < 0.015
        assert np.abs(tot).mean() < 0.005
        return
    ds = oracle.direct_sum(pos, mass, box, G, 2.8 * par["GravitySoftening"], repeat=1)
    meanacc = np.abs(ds).mean()
    err = np.abs(ds - tot) / meanacc
    assert err.max() < 3 * errtol
    assert err.mean() < 0.8 * errtol


@pytest.mark.parametrize("bulk", [0, 1])
@pytest.mark.parametrize("chunks", [1, 3, 8])
def test_force_step_aos(engine, b200, ics, chunks, bulk, monkeypatch):
    """b200_force_step_aos on the reference's 160-byte particle records equals
    the separate PM + tree calls and writes GravPM / FullTreeGravAccel / Potential in place.
    chunks > 1: the walk is issued in index-range groups whose records travel back
    while the next group is walked (the path large inputs take).  bulk = 0: only the 40 + 48 input bytes and 48 + 8
    output bytes of a record cross PCIe, as strided copies; bulk = 1: whole records (the path of an unusual layout)."""
    monkeypatch.setenv("B200_E2E_CHUNKS", str(chunks))
    if bulk:
        monkeypatch.setenv("B200_E2E_BULK", "1")
    else:
        monkeypatch.delenv("B200_E2E_BULK", raising=False)
    assert engine.force_step_aos_bytes() == ((160, 160) if bulk else (88, 56))
    pos, box = _distributions(ics)["gslrandom16"]
    n = len(pos)
    P = np.zeros(n, dtype=b200.PARTICLE_DTYPE)
    P["Pos"] = pos
    P["Mass"] = 1.0
    P["Type"] = 1
    P["ID"] = np.arange(n)
    rng = np.random.default_rng(5)
    P["FullTreeGravAccel"] = rng.standard_normal((n, 3)) * 300
    P["GravPM"] = rng.standard_normal((n, 3)) * 30
    P["Vel"] = rng.standard_normal((n, 3)); P["Hsml"] = rng.random(n); P["Potential"] = -7.0
    before = P.copy()
    old = P["FullTreeGravAccel"] + P["GravPM"]
    par = ics.tree_params(box, n, treeusebh=0, rcut=7.0)
    engine.gravpm_init_periodic(box, 1.5, 48, G)
    engine.force_step_aos(P, par)
    engine.set_particles(pos, np.ones(n, np.float32), oldacc=old)
    g, _ = engine.gravpm_force()
    engine.force_tree_full(box)
    acc, pot, _ = engine.grav_short_tree(par)
    # the CIC deposit uses fp64 atomics, so the mesh (hence GravPM) is reproducible only to rounding
    assert np.abs(P["GravPM"] - g).max() <= 1e-11 * np.abs(g).max()
    if chunks == 1:
        assert np.array_equal(P["FullTreeGravAccel"], acc)
        assert np.array_equal(P["Potential"], pot)
    else:       # same pairs and nodes per particle, summed in a different order
        assert np.abs(P["FullTreeGravAccel"] - acc).max() <= 1e-12 * np.abs(acc).max()
        assert np.abs(P["Potential"] - pot).max() <= 1e-12 * np.abs(pot).max()
    assert np.array_equal(P["Pos"], pos) and np.all(P["ID"] == np.arange(n))
    for f in P.dtype.names:         # every other byte of the records is the caller's
        if f not in ("GravPM", "FullTreeGravAccel", "Potential"):
            assert np.array_equal(P[f], before[f]), f


def test_force_step_dev(engine, ics):
    """b200_force_step_dev (PM on a second stream, concurrent with tree build + walk)
    equals the three separate calls."""
    import torch
    pos, box = _distributions(ics)["gslrandom16"]
    n = len(pos)
    mass = np.ones(n, np.float32)
    rng = np.random.default_rng(6)
    old = rng.standard_normal((n, 3)) * 300
    par = ics.tree_params(box, n, treeusebh=0, rcut=7.0)
    engine.gravpm_init_periodic(box, 1.5, 48, G)
    engine.set_particles(pos, mass, oldacc=old)
    g, _ = engine.gravpm_force()
    engine.force_tree_full(box)
    acc, pot, _ = engine.grav_short_tree(par)
    d = [torch.zeros(s, dtype=torch.float64, device="cuda") for s in ((n, 3), (n, 3), (n,))]
    for _ in range(3):
        engine.set_particles(pos, mass, oldacc=old)
        engine.force_step_dev(par, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr())
        assert np.abs(d[0].cpu().numpy() - g).max() <= 1e-11 * np.abs(g).max()
        assert np.array_equal(d[1].cpu().numpy(), acc)
        assert np.array_equal(d[2].cpu().numpy(), pot)


@pytest.mark.parametrize("nmesh", [40, 48, 96, 192])
def test_pm_own_transforms_equal_cufft(b200, ics, nmesh, monkeypatch):
    """The engine's shared-memory transform passes with the Green's function inside (csrc/pm_fft.cu) against cuFFT D2Z ->
    k_pm_potential_transfer -> cuFFT Z2D on the same deposit: potential mesh, readouts and the power-spectrum sums.
    Mesh sizes: 40 = 2^3.5, 48 = 2^4.3, 96 = 2^5.3, 192 = 2^6.3 (configs[0]); 768 is covered by test_config_parity.py."""
    pos, _ = ics.zeldovich_lattice(32, 32.0)
    mass = (1 + np.arange(len(pos)) % 3).astype(np.float32)
    res = []
    for sel in ("own", "cufft"):
        monkeypatch.setenv("B200_PM_FFT", sel)
        e = b200.Engine(0)
        try:
            e.set_particles(pos, mass)
            e.gravpm_init_periodic(32.0, 1.5, nmesh, G)
            assert e.pm_transform_kind() == (1 if sel == "own" else 0)
            e.pm_set_power(True)
            g, p = e.gravpm_force()
            res.append((g, p, e.pm_copy_mesh(1), e.pm_power()))
            e.pm_set_power(False)
            g2, _ = e.gravpm_force()           # the pass without the power-spectrum sums
            assert np.abs(g2 - g).max() <= 1e-12 * np.abs(g).max()
        finally:
            e.close()
    (g, p, m, ps), (g0, p0, m0, ps0) = res
    assert np.abs(m - m0).max() <= 1e-13 * np.abs(m0).max()
    assert np.abs(g - g0).max() <= 1e-11 * np.abs(g0).max()
    assert np.abs(p - p0).max() <= 1e-12 * np.abs(p0).max()
    assert np.array_equal(ps[2], ps0[2])
    assert abs(ps[3] - ps0[3]) <= 1e-12 * ps0[3]
    assert np.abs(ps[1] - ps0[1]).max() <= 1e-12 * ps0[1].max()
    assert np.abs(ps[0] - ps0[0]).max() <= 1e-11 * ps0[0].max()


def test_pm_transform_fallback_sizes(b200, ics):
    """Mesh sizes the shared-memory passes do not take (a prime factor other than 2, 3, 5) run on cuFFT."""
    pos, _ = ics.zeldovich_lattice(16, 16.0)
    mass = np.ones(len(pos), np.float32)
    e = b200.Engine(0)
    try:
        assert e.pm_transform_kind() == -1
        e.set_particles(pos, mass)
        e.gravpm_init_periodic(16.0, 1.5, 56, G)
        assert e.pm_transform_kind() == 0
        g, _ = e.gravpm_force()
        og, _, _ = oracle.pm_force(pos, mass, 16.0, 56, 1.5, G)
        assert np.abs(g - og).max() <= 1e-9 * np.abs(og).max()
        e.gravpm_init_periodic(16.0, 1.5, 60, G)
        assert e.pm_transform_kind() == 1
        g, _ = e.gravpm_force()
        og, _, _ = oracle.pm_force(pos, mass, 16.0, 60, 1.5, G)
        assert np.abs(g - og).max() <= 1e-9 * np.abs(og).max()
    finally:
        e.close()


def test_pm_power_spectrum_side_effect(engine, ics):
    """gravpm_force's power spectrum (powerspectrum_add_mode inside potential_transfer,
    gravpm.c:330-361,440): mode counts bit-exact, sums to rounding; forces unchanged."""
    pos, box = _distributions(ics)["gslrandom16"]
    n = len(pos)
    mass = np.ones(n, np.float32)
    engine.gravpm_init_periodic(box, 1.5, 48, G)
    engine.set_particles(pos, mass)
    g0, _ = engine.gravpm_force()
    engine.pm_set_power(True)
    g1, _ = engine.gravpm_force()
    pw, kk, nm, norm = engine.pm_power()
    engine.pm_set_power(False)
    opw, okk, onm, onorm = oracle.pm_power(pos, mass, box, 48)
    assert np.array_equal(nm, onm)
    assert abs(norm - onorm) <= 1e-12 * onorm
    assert np.abs(kk - okk).max() <= 1e-11 * okk.max()
    assert np.abs(pw - opw).max() <= 1e-9 * opw.max()
    assert np.abs(g1 - g0).max() <= 1e-11 * np.abs(g0).max()


def test_empty_and_tiny_inputs(engine, ics):
    """Edge cases: no particles, one particle, particles exactly on the box edge
    (Pos == BoxSize is legal, drift.c:77-78 -> iCell == Nmesh wraps, petapm.c:903-906)."""
    box = 8.0
    engine.gravpm_init_periodic(box, 1.5, 48, G)
    par = ics.tree_params(box, 8, treeusebh=1)
    engine.set_particles(np.zeros((0, 3)), np.zeros(0, np.float32))
    g, p = engine.gravpm_force()
    assert g.shape == (0, 3)
    info = engine.force_tree_full(box)
    assert info.numnodes == 1 and info.numparticles == 0
    pos = np.array([[box, box, box], [0.0, 0.0, 0.0], [box, 0.0, 4.0], [3.999, 4.0, 4.001]])
    mass = np.ones(len(pos), np.float32)
    engine.set_particles(pos, mass)
    g, p = engine.gravpm_force()
    og, op, oic = oracle.pm_force(pos, mass, box, 48, 1.5, G)
    assert np.array_equal(engine.pm_cell_index(), oic)
    assert np.abs(g - og).max() <= 1e-9 * np.abs(og).max()
    engine.force_tree_full(box)
    acc, pot, cnt = engine.grav_short_tree(par, want_counts=True)
    ot = oracle.OracleTree(pos, mass, box)
    oacc, opot, ocnt = ot.grav_short_tree(par, G, 48, 1.5)
    assert np.array_equal(cnt["particles"], ocnt["particles"])
    assert np.abs(acc - oacc).max() <= ACC_RTOL * max(np.abs(oacc).max(), 1e-300)
