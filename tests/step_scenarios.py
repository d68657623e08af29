"""The step-loop scenarios shared by tests/golden/make_golden_step.py (run on the reference's own
drift.c / timestep.c / timebinmgr.c) and tests/test_step.py (run on the oracle and on the GPU).
A 'stepper' is anything with the RefStep / StepOracle interface (oracle/ref.py, oracle/step.py)."""
import importlib
import numpy as np

TIMELINE = dict(TimeIC=0.1, TimeMax=1.0, outtimes=[0.2, 0.5])
G = 43.0071
TB = 46
NB = 47


def timeline_samples(seed=1, m=400):
    rng = np.random.default_rng(seed)
    ti = rng.integers(0, 3 << TB, m).astype(np.int64)
    ti[:4] = [0, 1 << TB, (2 << TB) - 1, (3 << TB) - 2]
    dloga = 10 ** rng.uniform(-7, -1, m)
    span = rng.integers(0, 1 << 40, m).astype(np.int64)
    span[(ti + span) >> TB >= 3] = 0
    return ti, dloga, span


def primitives_inputs(seed=7, n=1536, box=4000.0):
    """Mixed DM / gas set with garbage and swallowed particles, random time bins, particles next to the
    box faces, a few gas particles above the velocity cap."""
    rng = np.random.default_rng(seed)
    pos = rng.random((n, 3)) * box
    pos[:8] = [[box, box, box], [1e-9, box, 5.0], [box - 1e-7, 1e-7, 2000.0], [0.5, 0.5, 0.5],
               [box - 0.5, box - 0.5, box - 0.5], [12.5, 7.25, box - 3.0], [box - 12.5, 7.25, 3.0], [2000.0, 2000.0, 2000.0]]
    vel = 40.0 * rng.standard_normal((n, 3))
    typ = np.ones(n, np.uint8); typ[n // 3: 2 * n // 3] = 0; typ[-3:] = 4
    typ[-40:-3] = 2                                  # fast (neutrino) particles: excluded from the PM step criterion, timestep.c:1272
    flags = np.zeros(n, np.uint8); flags[[5, n // 3 + 7, n - 20]] = 1; flags[[6, n // 3 + 11]] = 2
    gas = typ == 0
    vel[np.flatnonzero(gas)[:5]] *= 3e4
    vel[typ == 2] *= 1e6
    mass = (1.0 + rng.random(n)).astype(np.float32)
    bin_grav = rng.integers(36, 42, n).astype(np.uint8)
    bin_hydro = np.zeros(n, np.uint8); bin_hydro[gas] = np.minimum(bin_grav[gas], rng.integers(35, 41, gas.sum())).astype(np.uint8)
    hsml = np.where(gas, 30.0 + 20.0 * rng.random(n), 0.0); hsml[np.flatnonzero(gas)[5]] = 0.4999 * box
    dthsml = np.where(gas, 4e2 * rng.standard_normal(n), 0.0); dthsml[np.flatnonzero(gas)[5]] = 4e2      # runs into the Box/2 cap
    return dict(pos=pos, vel=vel, type=typ, flags=flags, mass=mass, bin_grav=bin_grav, bin_hydro=bin_hydro, hsml=hsml, dthsml=dthsml,
                fullacc=3.0 * rng.standard_normal((n, 3)), gravpm=2.0 * rng.standard_normal((n, 3)),
                hydroacc=np.where(gas[:, None], 5.0 * rng.standard_normal((n, 3)), 0.0),
                entropy=np.where(gas, 1.0 + rng.random(n), 0.0), dtentropy=np.where(gas, 0.3 * rng.standard_normal(n), 0.0), box=box,
                maxsig=np.where(gas, 50.0 * 2.0 ** rng.uniform(-3, 3, n), 0.0))


def primitives_times(pm=False):
    """Kick times half a step behind an (odd multiple of 2^38) current time in the second sync interval."""
    Ti = (1 << TB) + 5 * (1 << 38)
    PM_start, PM_length = 1 << TB, 1 << 42
    if pm:
        Ti = PM_start + PM_length
    kick = np.zeros(NB, np.int64); last = np.zeros(NB, np.int64)
    for b in range(NB):
        d = (1 << b) if b > 0 else 0
        if d == 0:
            kick[b] = last[b] = Ti
        elif Ti % d == 0:
            kick[b] = Ti - d // 2; last[b] = Ti - d
        else:
            kick[b] = (Ti // d) * d + d // 2; last[b] = (Ti // d) * d
    scal = np.array([35, 41, 36, Ti, PM_length, PM_start, PM_start + PM_length // 2], np.int64)
    return scal, kick, last


def _load(S, d, **over):
    a = dict(d); a.update(over)
    S.set_particles(a["pos"], a["mass"], a["type"], a["box"], vel=a["vel"], flags=a["flags"], fullacc=a["fullacc"], gravpm=a["gravpm"],
                    bin_grav=a["bin_grav"], bin_hydro=a["bin_hydro"], hsml=a["hsml"], dthsml=a["dthsml"], hydroacc=a["hydroacc"],
                    entropy=a["entropy"], dtentropy=a["dtentropy"])


def run_primitives(S, d):
    """-> dict of outputs of each primitive (drift, active lists, the three kicks, kick-time update)."""
    out = {}
    scal, kick, last = primitives_times()
    Ti = int(scal[3])
    _load(S, d); S.set_times(scal, kick, last)
    out["ddrift"] = S.drift(Ti - (1 << 35), Ti, (12.5, -7.25, 3.0))
    g = S.get(); out["drift_pos"] = g["pos"].copy(); out["drift_hsml"] = g["hsml"].copy()
    act, counts = S.build_active()
    out["active"] = act.copy(); out["active_counts"] = counts.copy()
    out["last_drift"] = S.get_times()[2].copy()
    for mb in (36, 37, 41):
        out["sublist%d" % mb] = S.sublist(mb)
    atime = float(np.exp(S.loga_from_ti(Ti)))
    S.kick(0, atime)
    g = S.get(); out["halfkick_vel"] = g["vel"].copy(); out["halfkick_entropy"] = g["entropy"].copy()
    _load(S, d); S.set_times(scal, kick, last); S.build_active()
    S.kick(1, atime)
    g = S.get(); out["hydrokick_vel"] = g["vel"].copy(); out["hydrokick_entropy"] = g["entropy"].copy()
    S.kick(2)
    out["pmkick_vel"] = S.get()["vel"].copy(); out["pmkick_times"] = S.get_times()[0].copy()
    S.kick(3)
    out["kick_times"] = S.get_times()[1].copy()
    # hydro time bins of the active gas (find_hydro_timesteps)
    scal, kick, last = primitives_times()
    _load(S, d); S.set_times(scal, kick, last); S.build_active()
    bad, bh = S.hydro_timesteps(d["maxsig"], atime)
    out["hydro_bad"] = np.int64(bad); out["hydro_bins"] = np.array(bh, np.uint8); out["hydro_times"] = S.get_times()[0].copy()
    # the SplitGravityTimestepsOn = 0 assignment (find_timesteps), off and on a PM step
    S.set_gravity(importlib.import_module("mp-gadget_b200.ics").tree_params(d["box"], len(d["mass"])), G, 64, 1.5)
    for tag, pm in (("find", False), ("findpm", True)):
        scal, kick, last = primitives_times(pm=pm)
        _load(S, d); S.set_times(scal, kick, last); S.build_active()
        at = float(np.exp(S.loga_from_ti(int(scal[3]))))
        bad, bg, bh = S.find_timesteps(d["maxsig"], at, 1.5 * d["box"] / 64)
        out[tag + "_bad"] = np.int64(bad); out[tag + "_bin_grav"] = np.array(bg, np.uint8); out[tag + "_bin_hydro"] = np.array(bh, np.uint8)
        out[tag + "_times"] = S.get_times()[0].copy()
    # PM step: implicit list
    scal, kick, last = primitives_times(pm=True)
    _load(S, d); S.set_times(scal, kick, last)
    act, counts = S.build_active()
    assert act is None
    out["pm_active_counts"] = counts.copy()
    out["pm_sublist38"] = S.sublist(38)
    return out


def hierarchy_inputs(seed=5, n=2048, box=12000.0):
    ics = importlib.import_module("mp-gadget_b200.ics")
    rng = np.random.default_rng(seed)
    pos = rng.random((n, 3)) * box
    pos[: n // 2] = box / 2 + box / 16 * rng.standard_normal((n // 2, 3))
    pos = np.mod(pos, box)
    mass = np.full(n, ics.OMEGA0 * ics.rho_crit() * box ** 3 / n, np.float32)
    par = ics.tree_params(box, n, treeusebh=2)
    return dict(pos=pos, mass=mass, type=np.ones(n, np.uint8), vel=30.0 * rng.standard_normal((n, 3)), gravpm=5.0 * rng.standard_normal((n, 3)),
                box=box, par=par, nmesh=36, asmth=1.5)


HIER_STEPS = 8
HIER_KEEP = (0, 3, 7)


def run_hierarchy(S, d, steps=HIER_STEPS):
    """The hierarchical KDK loop from the initial (PM) step: -> per-step records."""
    S.set_particles(d["pos"], d["mass"], d["type"], d["box"], vel=d["vel"], gravpm=d["gravpm"])
    S.set_gravity(d["par"], G, d["nmesh"], d["asmth"])
    S.set_times(np.zeros(7, np.int64), np.zeros(NB, np.int64), np.zeros(NB, np.int64))
    rec = []
    for s in range(steps):
        bad, info = S.advance(first=(s == 0))
        g = S.get(); t = S.get_times()
        rec.append(dict(bad=bad, info=info.copy(), scal=t[0].copy(), kick=t[1].copy(), last=t[2].copy(), bin_grav=g["bin_grav"].copy(),
                        pos=g["pos"].copy(), vel=g["vel"].copy(), fullacc=g["fullacc"].copy()))
    return rec


NONSPLIT_STEPS = 6
NONSPLIT_KEEP = (0, 5)


def run_nonsplit(S, d, steps=NONSPLIT_STEPS):
    """The same loop with SplitGravityTimestepsOn = 0 (one tree over all particles per step, find_timesteps)."""
    S.set_particles(d["pos"], d["mass"], d["type"], d["box"], vel=d["vel"], gravpm=d["gravpm"])
    S.set_gravity(d["par"], G, d["nmesh"], d["asmth"])
    S.set_times(np.zeros(7, np.int64), np.zeros(NB, np.int64), np.zeros(NB, np.int64))
    rec = []
    for s in range(steps):
        bad, info = S.advance_nonsplit(d["asmth"] * d["box"] / d["nmesh"], first=(s == 0))
        g = S.get(); t = S.get_times()
        rec.append(dict(bad=bad, info=info.copy(), scal=t[0].copy(), kick=t[1].copy(), last=t[2].copy(), bin_grav=g["bin_grav"].copy(),
                        pos=g["pos"].copy(), vel=g["vel"].copy(), fullacc=g["fullacc"].copy()))
    return rec


GAS_STEPS = 10
GAS_KEEP = (0, 9)


def gas_hierarchy_inputs(seed=25, n=1536, box=9000.0):
    """DM + gas (every third particle) for the hierarchical loop with the hydro accelerations, DtEntropy, DtHsml and signal
    velocities held fixed: exercises the hydro kicks / hydro time bins inside the loop and, because gas can be hydro-active
    while gravitationally inactive, the sub-list branches at the top of both hierarchical drivers."""
    d = hierarchy_inputs(seed=seed, n=n, box=box)
    rng = np.random.default_rng(seed + 1)
    typ = np.ones(n, np.uint8); typ[::3] = 0
    gas = typ == 0
    d.update(type=typ, hsml=np.where(gas, 0.03 * box * (0.5 + rng.random(n)), 0.0), dthsml=np.where(gas, 0.5 * rng.standard_normal(n), 0.0),
             hydroacc=np.where(gas[:, None], 2.0 * rng.standard_normal((n, 3)), 0.0), entropy=np.where(gas, 1.0 + rng.random(n), 0.0),
             dtentropy=np.where(gas, 0.05 * rng.standard_normal(n), 0.0), maxsig=np.where(gas, 400.0 * 2.0 ** rng.uniform(4, 12, n), 0.0))
    return d


def run_gas_hierarchy(S, d, steps=GAS_STEPS):
    S.set_particles(d["pos"], d["mass"], d["type"], d["box"], vel=d["vel"], gravpm=d["gravpm"], hsml=d["hsml"], dthsml=d["dthsml"],
                    hydroacc=d["hydroacc"], entropy=d["entropy"], dtentropy=d["dtentropy"])
    S.set_gravity(d["par"], G, d["nmesh"], d["asmth"])
    S.set_times(np.zeros(7, np.int64), np.zeros(NB, np.int64), np.zeros(NB, np.int64))
    rec = []
    for s in range(steps):
        bad, info = S.advance(first=(s == 0), maxsig=d["maxsig"])
        g = S.get(); t = S.get_times()
        rec.append(dict(bad=bad, info=info.copy(), scal=t[0].copy(), kick=t[1].copy(), last=t[2].copy(), bin_grav=g["bin_grav"].copy(),
                        pos=g["pos"].copy(), vel=g["vel"].copy(), fullacc=g["fullacc"].copy(), hsml=g["hsml"].copy(), entropy=g["entropy"].copy()))
    return rec
