"""ctypes front-end of oracle/_ref/libref_tree.so: the reference's OWN tree,
treewalk and short-range gravity C compiled unmodified from /root/reference
(oracle/Makefile.ref).  TEST INFRASTRUCTURE ONLY -- pins the oracle and serves
as the CPU baseline.  load() returns None when the library was never built
(e.g. on a box without /root/reference and without a prebuilt oracle/_ref)."""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libref_tree.so")
# same driver + reference files, but gravshort-tree.c replaced by the B200 shim
# (mp-gadget_b200/host/libgadget_shims.c): grav_short_tree() runs on the GPU.
SO_DROPIN = os.path.join(_HERE, "_ref", "libref_dropin.so")
# same driver, but density.c + hydra.c replaced by libgadget_sph_shims.c: density() / hydro_force() run on the GPU
SO_DROPIN_SPH = os.path.join(_HERE, "_ref", "libref_dropin_sph.so")
# forcetree.c, gravshort-tree.c, density.c, hydra.c all replaced: no host octree at all
SO_DROPIN_ALL = os.path.join(_HERE, "_ref", "libref_dropin_all.so")
# the reference's own petapm.c + gravpm.c + powerspectrum.c with the single-rank PFFT stand-in
SO_PM = os.path.join(_HERE, "_ref", "libref_pm.so")
# gravpm.c + petapm.c replaced by the shim (gravpm_force on the GPU, P(k) through the reference's powerspectrum.c)
SO_DROPIN_PM = os.path.join(_HERE, "_ref", "libref_dropin_pm.so")
_inst = None


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


class Ref:
    def __init__(self, arena_gib=None, nthreads=0, so=SO):
        self.L = C.CDLL(so)
        self.L.ref_numnodes.restype = C.c_int64
        self.L.ref_tree_export.restype = C.c_int64
        if arena_gib is None:
            try:
                avail = os.sysconf("SC_AVPHYS_PAGES") * os.sysconf("SC_PAGE_SIZE") / 2.0 ** 30
            except Exception:
                avail = 16.0
            arena_gib = max(2.0, min(8.0, 0.4 * avail))      # the arena is touched on creation: 24 GiB cost ~25 s of page faults per process
        self.L.ref_init(C.c_double(arena_gib), C.c_int(nthreads))
        self.n = 0

    def tree_build(self, pos, mass, box, type=None, oldacc=None, topdepth=0):
        pos = np.ascontiguousarray(pos, np.float64)
        mass = np.ascontiguousarray(mass, np.float32)
        type = None if type is None else np.ascontiguousarray(type, np.uint8)
        oldacc = None if oldacc is None else np.ascontiguousarray(oldacc, np.float64)
        self.n = len(mass)
        self.L.ref_tree_build(C.c_int64(self.n), _p(pos), _p(mass), _p(type), _p(oldacc), C.c_double(box), C.c_int(topdepth))
        return int(self.L.ref_numnodes())

    def tree_build_top(self, pos, mass, box, top, type=None, oldacc=None):
        """force_tree_full below an arbitrary domain top tree: top = (Daughter, StartKey, Shift, Leaf) of DomainDecomp::TopNodes"""
        pos = np.ascontiguousarray(pos, np.float64)
        mass = np.ascontiguousarray(mass, np.float32)
        type = None if type is None else np.ascontiguousarray(type, np.uint8)
        oldacc = None if oldacc is None else np.ascontiguousarray(oldacc, np.float64)
        d, sk, sh, lf = (np.ascontiguousarray(top[0], np.int32), np.ascontiguousarray(top[1], np.uint64),
                         np.ascontiguousarray(top[2], np.int32), np.ascontiguousarray(top[3], np.int32))
        self.n = len(mass)
        self.L.ref_tree_build_top(C.c_int64(self.n), _p(pos), _p(mass), _p(type), _p(oldacc), C.c_double(box),
                                  C.c_int(len(d)), _p(d), _p(sk), _p(sh), _p(lf))
        return int(self.L.ref_numnodes())

    def grav_short_tree(self, par, G, nmesh, asmth):
        acc = np.zeros((self.n, 3))
        pot = np.zeros(self.n)
        self.L.ref_grav_short_tree(C.c_double(G), C.c_int(nmesh), C.c_double(asmth), C.c_double(par["ErrTolForceAcc"]),
                                   C.c_double(par["BHOpeningAngle"]), C.c_double(par["MaxBHOpeningAngle"]),
                                   C.c_int(par["TreeUseBH"]), C.c_double(par["Rcut"]), C.c_double(par["GravitySoftening"]),
                                   C.c_double(par["rho0"]), _p(acc), _p(pot))
        return acc, pot

    def tree_export(self):
        nn = int(self.L.ref_numnodes())
        out = dict(center=np.zeros((nn, 3)), len=np.zeros(nn), cofm=np.zeros((nn, 3)), mass=np.zeros(nn),
                   nocc=np.zeros(nn, np.int32), part=np.zeros((nn, 8), np.int32), toplevel=np.zeros(nn, np.int32))
        k = self.L.ref_tree_export(_p(out["center"]), _p(out["len"]), _p(out["cofm"]), _p(out["mass"]),
                                   _p(out["nocc"]), _p(out["part"]), _p(out["toplevel"]))
        return {a: v[:k] for a, v in out.items()}

    def sph_density(self, pos, mass, box, hsml, vel=None, entropy=None, kerneltype=1, eta=1.0, maxdev=2.0,
                    mingashsml_frac=0.006, softening=1.0, init_hsml=False, meansep=None, update_hsml=1, DoEgyDensity=0):
        """density() as tests/test_density.c:55-107 drives it (gas only, time bin 0)."""
        pos = np.ascontiguousarray(pos, np.float64); mass = np.ascontiguousarray(mass, np.float32)
        n = len(mass)
        vel = None if vel is None else np.ascontiguousarray(vel, np.float64)
        entropy = None if entropy is None else np.ascontiguousarray(entropy, np.float64)
        out = dict(hsml=np.array(hsml, dtype=np.float64, copy=True), density=np.zeros(n), egywtdensity=np.zeros(n),
                   dhsmlfac=np.zeros(n), divvel=np.zeros(n), curlvel=np.zeros(n), dthsml=np.zeros(n))
        self.n = n
        self.L.ref_sph_density(C.c_int64(n), _p(pos), _p(mass), _p(vel), _p(entropy), C.c_double(box), C.c_int(kerneltype),
                               C.c_double(eta), C.c_double(maxdev), C.c_double(mingashsml_frac), C.c_double(softening),
                               C.c_int(1 if init_hsml else 0), C.c_double(box if meansep is None else meansep),
                               C.c_int(update_hsml), C.c_int(DoEgyDensity), _p(out["hsml"]), _p(out["density"]),
                               _p(out["egywtdensity"]), _p(out["dhsmlfac"]), _p(out["divvel"]), _p(out["curlvel"]), _p(out["dthsml"]))
        return out

    def sph_hydro(self, atime=1.0, hubble=0.1, dloga_bin=0.0, DensityIndependentSphOn=0, ArtBulkViscConst=0.75,
                  DensityContrastLimit=100.0):
        """force_tree_calc_moments + hydro_force right after sph_density (run.c:472-489)."""
        n = self.n
        out = dict(acc=np.zeros((n, 3)), dtentropy=np.zeros(n), maxsignalvel=np.zeros(n))
        self.L.ref_sph_hydro(C.c_double(atime), C.c_double(hubble), C.c_double(dloga_bin), C.c_int(DensityIndependentSphOn),
                             C.c_double(ArtBulkViscConst), C.c_double(DensityContrastLimit), _p(out["acc"]),
                             _p(out["dtentropy"]), _p(out["maxsignalvel"]))
        return out

    def sph_mixed(self, bins, Ti_Current, tables, atime, hubble, DoEgyDensity, vel=None, fullacc=None, hydroacc=None, dtentropy=None):
        """A mixed-time-bin density + hydro step on the state left by sph_density + sph_hydro.
        tables: dict of per-bin arrays (TIMEBINS + 2 = 48 entries) gravkick, hydrokick, drift,
        dloga_pred, dloga_bin.  Returns (active indices, outputs of all particles)."""
        n = self.n
        NT = 48
        tab = np.zeros((5, NT))
        for r, k in enumerate(("gravkick", "hydrokick", "drift", "dloga_pred", "dloga_bin")):
            tab[r, :len(tables[k])] = tables[k]
        bins = np.ascontiguousarray(bins, np.uint8)
        opt = lambda a: None if a is None else np.ascontiguousarray(a, np.float64)
        vel, fullacc, hydroacc, dtentropy = opt(vel), opt(fullacc), opt(hydroacc), opt(dtentropy)
        out = dict(hsml=np.zeros(n), density=np.zeros(n), egywtdensity=np.zeros(n), dhsmlfac=np.zeros(n), divvel=np.zeros(n),
                   curlvel=np.zeros(n), dthsml=np.zeros(n), acc=np.zeros((n, 3)), dtentropy=np.zeros(n), maxsignalvel=np.zeros(n))
        act = np.zeros(n, np.int32); na = C.c_int64()
        self.L.ref_sph_mixed(_p(bins), C.c_int64(Ti_Current), _p(tab), _p(vel), _p(fullacc), _p(hydroacc), _p(dtentropy),
                             C.c_double(atime), C.c_double(hubble), C.c_int(DoEgyDensity), _p(act), C.byref(na),
                             _p(out["hsml"]), _p(out["density"]), _p(out["egywtdensity"]), _p(out["dhsmlfac"]), _p(out["divvel"]),
                             _p(out["curlvel"]), _p(out["dthsml"]), _p(out["acc"]), _p(out["dtentropy"]), _p(out["maxsignalvel"]))
        return act[:na.value].copy(), out

    def gravpm_force(self, pos, mass, box, nmesh, asmth, G, outdir, time=1.0):
        """gravpm_init_periodic + gravpm_force of the reference (needs so=SO_PM).  Returns
        (GravPM[n,3], Potential[n]); the reference writes outdir/powerspectrum-<time>.txt."""
        pos = np.ascontiguousarray(pos, np.float64); mass = np.ascontiguousarray(mass, np.float32)
        n = len(mass)
        g = np.zeros((n, 3)); p = np.zeros(n)
        self.L.ref_gravpm_force(C.c_int64(n), _p(pos), _p(mass), C.c_double(box), C.c_int(nmesh), C.c_double(asmth), C.c_double(G),
                                C.c_char_p(outdir.encode()), C.c_double(time), _p(g), _p(p))
        return g, p

    def petapm_c2r(self, pos, box, nmesh, rho_k, functions):
        """petapm_force_init + petapm_force_c2r + petapm_force_finish of the reference (petapm.c:263-362) driven as
        libgenic/zeldovich.c:150-229 drives them, on a given spectrum rho_k[nmesh, nmesh, nmesh//2+1] (x slowest) with
        table-driven transfers [(kind, table[k2]), ...] -> one read-out array per function (needs so=SO_PM)."""
        pos = np.ascontiguousarray(pos, np.float64)
        n = len(pos)
        rk = np.ascontiguousarray(np.asarray(rho_k, np.complex128))
        assert rk.shape == (nmesh, nmesh, nmesh // 2 + 1)
        nk2 = 3 * (nmesh // 2) ** 2 + 1
        kinds = np.array([k for k, _ in functions], np.int32)
        tabs = np.ascontiguousarray(np.stack([np.asarray(t, np.float64) for _, t in functions]))
        assert tabs.shape == (len(functions), nk2)
        out = np.zeros((len(functions), n))
        rc = self.L.ref_petapm_c2r(C.c_int64(n), _p(pos), C.c_double(box), C.c_int(nmesh), C.c_void_p(rk.ctypes.data), C.c_int(len(functions)),
                                   _p(kinds), _p(tabs), C.c_int64(nk2), _p(out))
        assert rc == 0
        return [out[j] for j in range(len(functions))]

    def peano_keys(self, pos, box):
        pos = np.ascontiguousarray(pos, np.float64); keys = np.zeros(len(pos), np.uint64)
        self.L.ref_peano_keys(C.c_int64(len(pos)), _p(pos), C.c_double(box), _p(keys))
        return keys

    def topleaf(self, keys, daughter, startkey, shift, leaf):
        keys = np.ascontiguousarray(keys, np.uint64); out = np.zeros(len(keys), np.int32)
        a = [np.ascontiguousarray(daughter, np.int32), np.ascontiguousarray(startkey, np.uint64), np.ascontiguousarray(shift, np.int32),
             np.ascontiguousarray(leaf, np.int32)]
        self.L.ref_topleaf(C.c_int64(len(keys)), _p(keys), C.c_int(len(a[0])), _p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]), _p(out))
        return out

    def timings(self):
        b, w = C.c_double(), C.c_double()
        self.L.ref_timings(C.byref(b), C.byref(w))
        return b.value, w.value

    def tree_gravity(self, pos, mass, box, nmesh, asmth, G, par, oldacc=None, topdepth=None):
        """force_tree_full + grav_short_tree as run.c:546-548."""
        if topdepth is None:
            topdepth = 2 if len(mass) >= 100000 else 0      # >= nthreads top leaves so the merge is parallel
        self.tree_build(pos, mass, box, oldacc=oldacc, topdepth=topdepth)
        acc, _ = self.grav_short_tree(par, G, nmesh, asmth)
        return acc


def available():
    return os.path.exists(SO)


def load(**kw):
    global _inst
    if not available():
        return None
    if _inst is None:
        _inst = Ref(**kw)
    return _inst


# the reference's own drift.c + timestep.c + timebinmgr.c on top of its tree gravity (step loop)
SO_STEP = os.path.join(_HERE, "_ref", "libref_step.so")
# the same, with the step-loop functions redirected (ld --wrap) to mp-gadget_b200/host/libgadget_step_shims.c -> GPU
SO_DROPIN_STEP = os.path.join(_HERE, "_ref", "libref_dropin_step.so")
NBINS = 47      # TIMEBINS + 1, timebinmgr.h:13


class RefStep(Ref):
    """Step loop of the reference (run.c:355-800 for collisionless particles with HierarchicalGravity):
    drift, active lists, half kicks, hierarchical gravity + time-bin assignment.  One instance per
    process (the sync-point table is set once)."""

    def __init__(self, TimeIC, TimeMax, outtimes=(), Omega0=0.288, OmegaBaryon=0.0472, Hubble=0.1, G=43.0071,
                 ErrTolIntAccuracy=0.02, MaxGasVel=3e5, MaxSizeTimestep=0.1, MinSizeTimestep=0.0,
                 MaxRMSDisplacementFac=0.2, CourantFac=0.15, so=SO_STEP, **kw):
        super().__init__(so=so, **kw)
        L = self.L
        L.ref_step_factor.restype = C.c_double
        L.ref_loga_from_ti.restype = C.c_double
        L.ref_dloga_from_dti.restype = C.c_double
        L.ref_ti_from_loga.restype = C.c_int64
        L.ref_dti_from_dloga.restype = C.c_int64
        L.ref_step_drift.restype = C.c_double
        L.ref_step_build_active.restype = C.c_int64
        L.ref_step_sublist.restype = C.c_int64
        L.ref_step_softening.restype = C.c_double
        out = np.ascontiguousarray(outtimes, np.float64)
        ts = np.array([ErrTolIntAccuracy, MaxGasVel, MaxSizeTimestep, MinSizeTimestep, MaxRMSDisplacementFac, CourantFac])
        self.cosmo = dict(Omega0=Omega0, OmegaBaryon=OmegaBaryon, Hubble=Hubble, G=G)
        self.tspar = dict(ErrTolIntAccuracy=ErrTolIntAccuracy, MaxGasVel=MaxGasVel, MaxSizeTimestep=MaxSizeTimestep,
                          MinSizeTimestep=MinSizeTimestep, MaxRMSDisplacementFac=MaxRMSDisplacementFac, CourantFac=CourantFac)
        self.sync_loga = np.log(np.unique(np.concatenate([[TimeIC, TimeMax], out[(out >= TimeIC) & (out <= TimeMax)]])))
        rc = L.ref_step_init(C.c_double(TimeIC), C.c_double(TimeMax), C.c_int(len(out)), _p(out), C.c_double(Omega0),
                             C.c_double(OmegaBaryon), C.c_double(Hubble), C.c_double(G), _p(ts))
        if rc:
            raise RuntimeError("RefStep: the reference timeline can be set only once per process")

    # --- the integer timeline (timebinmgr.c:380-447) and the kick / drift integrals
    def loga_from_ti(self, ti):
        return self.L.ref_loga_from_ti(C.c_int64(ti))

    def dti_from_dloga(self, dloga, ti):
        return int(self.L.ref_dti_from_dloga(C.c_double(dloga), C.c_int64(ti)))

    def dloga_from_dti(self, dti, ti):
        return self.L.ref_dloga_from_dti(C.c_int64(dti), C.c_int64(ti))

    def factor(self, kind, t0, t1):
        """kind 0 drift, 1 gravkick, 2 hydrokick"""
        return self.L.ref_step_factor(C.c_int(kind), C.c_int64(t0), C.c_int64(t1))

    def set_times(self, scal, ti_kick, ti_last):
        self.L.ref_step_set_times(_p(np.ascontiguousarray(scal, np.int64)), _p(np.ascontiguousarray(ti_kick, np.int64)),
                                  _p(np.ascontiguousarray(ti_last, np.int64)))

    def get_times(self):
        scal = np.zeros(7, np.int64); kick = np.zeros(NBINS, np.int64); last = np.zeros(NBINS, np.int64)
        self.L.ref_step_get_times(_p(scal), _p(kick), _p(last))
        return scal, kick, last

    def set_particles(self, pos, mass, type, box, vel=None, flags=None, fullacc=None, gravpm=None, bin_grav=None, bin_hydro=None,
                      hsml=None, dthsml=None, hydroacc=None, entropy=None, dtentropy=None, topdepth=0, ti_drift=0):
        f = lambda a: None if a is None else np.ascontiguousarray(a, np.float64)
        b = lambda a: None if a is None else np.ascontiguousarray(a, np.uint8)
        pos = f(pos); mass = np.ascontiguousarray(mass, np.float32); type = b(type)
        self.n = len(mass)
        keep = [f(vel), b(flags), f(fullacc), f(gravpm), b(bin_grav), b(bin_hydro), f(hsml), f(dthsml), f(hydroacc), f(entropy), f(dtentropy)]
        self.L.ref_step_set_particles(C.c_int64(self.n), _p(pos), _p(keep[0]), _p(mass), _p(type), _p(keep[1]), _p(keep[2]), _p(keep[3]),
                                      _p(keep[4]), _p(keep[5]), _p(keep[6]), _p(keep[7]), _p(keep[8]), _p(keep[9]), _p(keep[10]),
                                      C.c_double(box), C.c_int(topdepth), C.c_int64(ti_drift))

    def get(self):
        n = self.n
        out = dict(pos=np.zeros((n, 3)), vel=np.zeros((n, 3)), hsml=np.zeros(n), entropy=np.zeros(n), bin_grav=np.zeros(n, np.uint8),
                   fullacc=np.zeros((n, 3)), ti_drift=np.zeros(n, np.int64))
        self.L.ref_step_get(_p(out["pos"]), _p(out["vel"]), _p(out["hsml"]), _p(out["entropy"]), _p(out["bin_grav"]),
                            _p(out["fullacc"]), _p(out["ti_drift"]))
        return out

    def drift(self, ti0, ti1, shift=(0.0, 0.0, 0.0)):
        return self.L.ref_step_drift(C.c_int64(ti0), C.c_int64(ti1), _p(np.ascontiguousarray(shift, np.float64)))

    def build_active(self):
        """-> (list or None when implicit, [NumActiveParticle, NumActiveGravity, NumActiveHydro])"""
        lst = np.zeros(self.n + 1, np.int32); counts = np.zeros(3, np.int64)
        na = int(self.L.ref_step_build_active(_p(lst), _p(counts)))
        return (None if na < 0 else lst[:na].copy()), counts

    def sublist(self, maxtimebin):
        lst = np.zeros(self.n + 1, np.int32)
        na = int(self.L.ref_step_sublist(C.c_int(maxtimebin), _p(lst)))
        return lst[:na].copy()

    def kick(self, kind, atime=1.0):
        """0 apply_half_kick, 1 apply_hydro_half_kick, 2 apply_PM_half_kick, 3 update_kick_times"""
        self.L.ref_step_kick(C.c_int(kind), C.c_double(atime))

    def hydro_timesteps(self, maxsig, atime, first=False):
        """find_hydro_timesteps on the current active list -> (bad count, TimeBinHydro[n])"""
        out = np.zeros(self.n, np.uint8)
        bad = int(self.L.ref_step_hydro_timesteps(_p(np.ascontiguousarray(maxsig, np.float64)), C.c_double(atime), C.c_int(1 if first else 0), _p(out)))
        return bad, out

    def find_timesteps(self, maxsig, atime, asmth, first=False):
        """find_timesteps (SplitGravityTimestepsOn = 0) on the current active list -> (bad, TimeBinGravity[n], TimeBinHydro[n])"""
        bg = np.zeros(self.n, np.uint8); bh = np.zeros(self.n, np.uint8)
        bad = int(self.L.ref_step_find_timesteps(_p(np.ascontiguousarray(maxsig, np.float64)), C.c_double(atime), C.c_double(asmth),
                                                 C.c_int(1 if first else 0), _p(bg), _p(bh)))
        return bad, bg, bh

    def set_gravity(self, par, G, nmesh, asmth):
        self.L.ref_step_set_gravity(C.c_double(G), C.c_int(nmesh), C.c_double(asmth), C.c_double(par["ErrTolForceAcc"]),
                                    C.c_double(par["BHOpeningAngle"]), C.c_double(par["MaxBHOpeningAngle"]), C.c_int(par["TreeUseBH"]),
                                    C.c_double(par["Rcut"]), C.c_double(par["GravitySoftening"]))
        return self.L.ref_step_softening()

    def advance(self, first=False, maxsig=None):
        """One pass of the run.c loop -> (bad-timestep count, [NumActiveParticle, NumActiveGravity, is_PM]).  maxsig
        (SphP[].MaxSignalVel by particle index): gas takes part -- hydro half kicks and find_hydro_timesteps with the hydro
        accelerations held fixed."""
        self._maxsig = None if maxsig is None else np.ascontiguousarray(maxsig, np.float64)
        self.L.ref_step_set_maxsig(_p(self._maxsig))
        info = np.zeros(3, np.int64)
        bad = int(self.L.ref_step_advance(C.c_int(1 if first else 0), _p(info)))
        return bad, info


    def advance_nonsplit(self, asmth, first=False):
        """The same pass with SplitGravityTimestepsOn = 0 -> (bad, [NumActiveParticle, NumActiveGravity, is_PM])"""
        info = np.zeros(3, np.int64)
        bad = int(self.L.ref_step_advance_nonsplit(C.c_int(1 if first else 0), C.c_double(asmth), _p(info)))
        return bad, info


def step_available():
    return os.path.exists(SO_STEP)


# file-static routines of the reference's domain.c (oracle/ref_domain_driver.c includes the file where it lies)
SO_DOMAIN = os.path.join(_HERE, "_ref", "libref_domain.so")


class RefDomain(Ref):
    def __init__(self, **kw):
        super().__init__(so=SO_DOMAIN, **kw)

    # method order: leaf_counts is defined below assign_balanced
    def assign_balanced(self, ntask, startkey, cost, nseg_per_task=1):
        """domain_assign_topleaves_balanced for `ntask` tasks -> (task per leaf in the final order, input leaf at each position)"""
        sk = np.ascontiguousarray(startkey, np.uint64); cost = np.ascontiguousarray(cost, np.int64)
        task = np.zeros(len(cost), np.int32); order = np.zeros(len(cost), np.int32)
        self.L.ref_domain_assign(C.c_int(ntask), C.c_int(len(cost)), _p(sk), _p(cost), C.c_int(nseg_per_task), _p(task), _p(order))
        return task, order


    def leaf_counts(self, pos, box, top, nleaf, flags=None):
        """domain_compute_costs -> TopLeafCount[nleaf]; top = (daughter, startkey, shift, leaf)"""
        pos = np.ascontiguousarray(pos, np.float64)
        flags = None if flags is None else np.ascontiguousarray(flags, np.uint8)
        a = [np.ascontiguousarray(top[0], np.int32), np.ascontiguousarray(top[1], np.uint64), np.ascontiguousarray(top[2], np.int32),
             np.ascontiguousarray(top[3], np.int32)]
        out = np.zeros(nleaf, np.int64)
        self.L.ref_domain_counts(C.c_int64(len(pos)), _p(pos), _p(flags), C.c_double(box), C.c_int(len(a[0])), _p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]),
                                 C.c_int(nleaf), _p(out))
        return out


    # --- the top tree, stage by stage (ref_domain_driver.c); trees are numpy arrays of oracle.TOPNODE_DTYPE
    def toptree_local(self, pos, box, subsample, maxnodes, flags=None, presort=0):
        from . import TOPNODE_DTYPE
        assert self.L.ref_toptree_node_size() == TOPNODE_DTYPE.itemsize
        pos = np.ascontiguousarray(pos, np.float64)
        flags = None if flags is None else np.ascontiguousarray(flags, np.uint8)
        tree = np.zeros(maxnodes, TOPNODE_DTYPE); size = C.c_int(0)
        rc = self.L.ref_toptree_local(C.c_int64(len(pos)), _p(pos), _p(flags), C.c_double(box), C.c_int(subsample), C.c_int(presort),
                                      C.c_int(maxnodes), _p(tree), C.byref(size))
        return rc, tree, size.value

    def toptree_truncate(self, tree, size, countlimit, costlimit):
        s = C.c_int(size)
        self.L.ref_toptree_truncate(_p(tree), C.byref(s), C.c_int64(countlimit), C.c_int64(costlimit))
        return s.value

    def toptree_merge(self, treeA, sizeA, treeB):
        s = C.c_int(sizeA)
        self.L.ref_toptree_merge(_p(treeA), C.byref(s), _p(treeB), C.c_int(len(treeA)))
        return s.value

    def toptree_global_refine(self, tree, size, countlimit, costlimit):
        s = C.c_int(size)
        rc = self.L.ref_toptree_global_refine(_p(tree), C.byref(s), C.c_int(len(tree)), C.c_int64(countlimit), C.c_int64(costlimit))
        return rc, s.value

    def toptree_leaves(self, tree, size):
        leaf = np.zeros(size, np.int32)
        nl = self.L.ref_toptree_leaves(_p(tree), C.c_int(size), _p(leaf))
        return nl, leaf


    def exchange_plan(self, type, flags, topleaf, task_of_leaf, ntask, thistask):
        """domain_build_exchange_list + domain_build_plan (exchange.c, static) -> (list, togo[ntask][7], ngarbage)"""
        type = np.ascontiguousarray(type, np.uint8); flags = np.ascontiguousarray(flags, np.uint8)
        tl = np.ascontiguousarray(topleaf, np.int32); tk = np.ascontiguousarray(task_of_leaf, np.int32)
        lst = np.zeros(len(tl) + 1, np.int32); togo = np.zeros((ntask, 7), np.int64); ng = C.c_int64()
        self.L.ref_exchange_plan.restype = C.c_int64
        nex = self.L.ref_exchange_plan(C.c_int64(len(tl)), _p(type), _p(flags), _p(tl), C.c_int(len(tk)), _p(tk), C.c_int(ntask), C.c_int(thistask),
                                       _p(lst), _p(togo), C.byref(ng))
        return lst[:nex].copy(), togo, int(ng.value)


    def fof_primary(self, pos, ids, type, box, ll):
        """fof_label_primary of the reference's own fof.c (FOFPrimaryLinkTypes = 2: dark matter) -> MinID[n]"""
        pos = np.ascontiguousarray(pos, np.float64); ids = np.ascontiguousarray(ids, np.int64); type = np.ascontiguousarray(type, np.uint8)
        out = np.zeros(len(ids), np.int64)
        self.L.ref_fof_primary(C.c_int64(len(ids)), _p(pos), _p(ids), _p(type), C.c_double(box), C.c_double(ll), _p(out))
        return out


def domain_available():
    return os.path.exists(SO_DOMAIN)
