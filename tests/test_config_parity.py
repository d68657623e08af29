"""Parity at the sizes BASELINE.json names ("no GPU test at 64^3/Nmesh 192, nothing checks the 256^3
result against the CPU path").

  * configs[0]  examples/dm-small: 64^3 dark matter, Nmesh 192, Box 64000 kpc/h -- the full force step (PM + tree build +
    walk), a Barnes-Hut pass then the relative criterion fed by it, as run.c:519-548 does.  PM cell indices, node count and
    the four interaction counters of every particle bit-exact against the oracle; accelerations <= 1e-6 against the
    oracle AND against the reference's own compiled tree C (oracle/_ref) where that library travelled with the repo.
  * configs[1]  256^3 / Nmesh 768: the same step on the displaced-lattice state the bench times.  The oracle builds the
    same tree and walks 10^4 sampled targets (gravshort-tree.c:253-379); counters bit-exact, accelerations <= 1e-6; the
    oracle's full 768^3 PM (pocketfft) gives GravPM for every particle, cell indices bit-exact.  At this size the walk
    takes the paths small fixtures never reach: the piece pool is regrown and the walk repeated, the 8-group end-to-end walk
    of b200_force_step_aos runs for real, and its results must equal the single-call walk to summation order (1e-12).
"""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu
G = 43.0071
ACC_RTOL = 1e-6
CNT = ("nodes_accepted", "nodes_opened", "nodes_discarded", "particles")


def _rel(a, b):
    return np.abs(a - b).max() / (np.sqrt((b ** 2).sum(axis=1)).mean() + 1e-300)


def test_config0_dm_small_64(engine, ics):
    ng, box, nmesh = 64, 64000.0, 192
    pos, mass = ics.zeldovich_lattice(ng, box)
    n = len(mass)
    par = ics.tree_params(box, n, treeusebh=1)
    engine.gravpm_init_periodic(box, 1.5, nmesh, G)
    og, opot, oic = oracle.pm_force(pos, mass, box, nmesh, 1.5, G)
    ot = oracle.OracleTree(pos, mass, box)
    ref = None
    try:
        from oracle import ref as R
        ref = R.load()
    except Exception:
        ref = None
    oldacc = None
    for usebh in (1, 0):                  # TreeUseBH = 2 semantics: one Barnes-Hut pass, then the relative criterion
        par["TreeUseBH"] = usebh
        engine.set_particles(pos, mass, oldacc=oldacc)
        gpm, ppm = engine.gravpm_force()
        assert np.array_equal(engine.pm_cell_index(), oic)
        assert np.abs(gpm - og).max() <= 1e-9 * np.abs(og).max()
        info = engine.force_tree_full(box)
        assert info.numnodes == ot.t.numnodes and info.numparticles == n
        acc, pot, cnt = engine.grav_short_tree(par, want_counts=True)
        oacc, opot_t, ocnt = ot.grav_short_tree(par, G, nmesh, 1.5, oldacc=oldacc)
        for f in CNT:
            assert np.array_equal(cnt[f], ocnt[f]), (usebh, f)
        assert _rel(acc, oacc) < ACC_RTOL
        assert _rel(acc + gpm, oacc + og) < ACC_RTOL
        if ref is not None:
            racc = ref.tree_gravity(pos, mass, box, nmesh, 1.5, G, par, oldacc, topdepth=0)
            assert _rel(acc, racc) < ACC_RTOL
        oldacc = acc + gpm


def test_config1_256_sampled(engine, b200, ics):
    import torch
    ng = 256
    box, nmesh = float(ng), 768
    d_pos, d_mass = ics.bench_ics("displaced", ng, box, device="cuda")
    pos, mass = d_pos.cpu().numpy(), d_mass.cpu().numpy()
    del d_pos, d_mass
    n = len(mass)
    par = ics.tree_params(box, n, treeusebh=1)
    engine.gravpm_init_periodic(box, 1.5, nmesh, G)
    engine.set_particles(pos, mass)
    gpm, _ = engine.gravpm_force()
    og, _, oic = oracle.pm_force(pos, mass, box, nmesh, 1.5, G)
    assert np.array_equal(engine.pm_cell_index(), oic)
    assert np.abs(gpm - og).max() <= 1e-9 * np.abs(og).max()
    info = engine.force_tree_full(box)
    ot = oracle.OracleTree(pos, mass, box)
    assert info.numnodes == ot.t.numnodes
    rng = np.random.default_rng(5)
    sample = np.sort(rng.choice(n, size=12000, replace=False)).astype(np.int32)
    oldacc = None
    for usebh in (1, 0):
        par["TreeUseBH"] = usebh
        if oldacc is not None:
            engine.set_particles(pos, mass, oldacc=oldacc)
            engine.force_tree_full(box)
        acc, pot, cnt = engine.grav_short_tree(par, want_counts=True)
        oacc, opot, ocnt = ot.grav_short_tree(par, G, nmesh, 1.5, oldacc=oldacc, active=sample, full=True)
        for f in CNT:
            assert np.array_equal(cnt[f][sample], ocnt[f][sample]), (usebh, f)
        assert _rel(acc[sample], oacc[sample]) < ACC_RTOL
        assert _rel((acc + gpm)[sample], (oacc + og)[sample]) < ACC_RTOL
        oldacc = acc + gpm
    # the end-to-end entry on the reference's 160-byte records: 8 walk groups, write-back pipelined.
    # Same inputs (FullTreeGravAccel + GravPM = the old acceleration of the last pass) as the single-call walk.
    P = np.zeros(n, dtype=b200.PARTICLE_DTYPE)
    P["Pos"] = pos; P["Mass"] = mass; P["Type"] = 1; P["ID"] = np.arange(n)
    P["FullTreeGravAccel"] = acc; P["GravPM"] = gpm
    engine.set_particles(pos, mass, oldacc=acc + gpm)
    engine.force_tree_full(box)
    acc1, pot1, _ = engine.grav_short_tree(par, want_counts=False)
    pinned = torch.empty(n * 160, dtype=torch.uint8).pin_memory()
    pinned.numpy()[:] = P.view(np.uint8).reshape(-1)
    engine.force_step_aos(None, par, ptr=pinned.data_ptr(), n=n)
    out = pinned.numpy().view(b200.PARTICLE_DTYPE)
    # (the groups cut the curve into different warps of 32 targets, so a target's terms are summed in another order)
    assert _rel(out["FullTreeGravAccel"], acc1) < 1e-12
    assert np.abs(out["GravPM"] - og).max() <= 1e-9 * np.abs(og).max()
