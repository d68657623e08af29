"""ctypes front-end of oracle/oracle_step.c (the step loop around the force computation) --
TEST INFRASTRUCTURE ONLY.  StepOracle mirrors oracle.ref.RefStep call for call so that the parity
tests read the same on both sides."""
import ctypes as C
import numpy as np
from . import lib, GravShortParams, OracleTree

NBINS = 47      # TIMEBINS + 1


class Timeline(C.Structure):
    _fields_ = [("nsync", C.c_int64), ("loga", C.POINTER(C.c_double))]


class Cosmo(C.Structure):
    _fields_ = [("Omega0", C.c_double), ("OmegaBaryon", C.c_double), ("Hubble", C.c_double), ("G", C.c_double)]


class Times(C.Structure):
    _fields_ = [("mintimebin", C.c_int32), ("maxtimebin", C.c_int32), ("mingravtimebin", C.c_int32), ("pad_", C.c_int32),
                ("Ti_kick", C.c_int64 * NBINS), ("Ti_lastactivedrift", C.c_int64 * NBINS),
                ("Ti_Current", C.c_int64), ("PM_length", C.c_int64), ("PM_start", C.c_int64), ("PM_kick", C.c_int64)]


class StepParams(C.Structure):
    _fields_ = [("ErrTolIntAccuracy", C.c_double), ("MaxGasVel", C.c_double), ("MaxSizeTimestep", C.c_double),
                ("MinSizeTimestep", C.c_double), ("MaxRMSDisplacementFac", C.c_double), ("softening", C.c_double),
                ("CourantFac", C.c_double)]


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _lib():
    L = lib()
    if not getattr(L, "_step_ready", False):
        L.oracle_loga_from_ti.restype = C.c_double
        L.oracle_dloga_from_dti.restype = C.c_double
        L.oracle_step_factor.restype = C.c_double
        L.oracle_gravity_dloga.restype = C.c_double
        for f in ("oracle_ti_from_loga", "oracle_dti_from_dloga", "oracle_drift", "oracle_build_active", "oracle_active_sublist",
                  "oracle_convert_timestep", "oracle_pm_timestep_ti"):
            getattr(L, f).restype = C.c_int64
        L._step_ready = True
    return L


def is_timebin_active(b, ti):
    return bool(_lib().oracle_is_timebin_active(C.c_int(int(b)), C.c_int64(int(ti))))


def dti_from_timebin(b):
    return (1 << int(b)) if b > 0 else 0


class StepOracle:
    def __init__(self, sync_loga, Omega0=0.288, OmegaBaryon=0.0472, Hubble=0.1, G=43.0071, ErrTolIntAccuracy=0.02, MaxGasVel=3e5,
                 MaxSizeTimestep=0.1, MinSizeTimestep=0.0, MaxRMSDisplacementFac=0.2, CourantFac=0.15, **_):
        self.L = _lib()
        self.sync = np.ascontiguousarray(sync_loga, np.float64)
        self.tl = Timeline(len(self.sync), self.sync.ctypes.data_as(C.POINTER(C.c_double)))
        self.cosmo = Cosmo(Omega0, OmegaBaryon, Hubble, G)
        self.sp = StepParams(ErrTolIntAccuracy, MaxGasVel, MaxSizeTimestep, MinSizeTimestep, MaxRMSDisplacementFac, 0.0, CourantFac)
        self.t = Times()
        self.act = None
        self.counts = None

    # --- timeline
    def loga_from_ti(self, ti):
        return self.L.oracle_loga_from_ti(C.byref(self.tl), C.c_int64(ti))

    def dti_from_dloga(self, dloga, ti):
        return int(self.L.oracle_dti_from_dloga(C.byref(self.tl), C.c_double(dloga), C.c_int64(ti)))

    def dloga_from_dti(self, dti, ti):
        return self.L.oracle_dloga_from_dti(C.byref(self.tl), C.c_int64(dti), C.c_int64(ti))

    def factor(self, kind, t0, t1):
        return self.L.oracle_step_factor(C.byref(self.cosmo), C.byref(self.tl), C.c_int(kind), C.c_int64(t0), C.c_int64(t1))

    def hubble(self, a):
        return self.cosmo.Hubble * np.sqrt(self.cosmo.Omega0 / a ** 3 + 1 - self.cosmo.Omega0)

    def atime(self, ti=None):
        return float(np.exp(self.loga_from_ti(self.t.Ti_Current if ti is None else ti)))

    def set_times(self, scal, ti_kick, ti_last):
        t = self.t
        t.mintimebin, t.maxtimebin, t.mingravtimebin = int(scal[0]), int(scal[1]), int(scal[2])
        t.Ti_Current, t.PM_length, t.PM_start, t.PM_kick = int(scal[3]), int(scal[4]), int(scal[5]), int(scal[6])
        for b in range(NBINS):
            t.Ti_kick[b] = int(ti_kick[b]); t.Ti_lastactivedrift[b] = int(ti_last[b])

    def get_times(self):
        t = self.t
        scal = np.array([t.mintimebin, t.maxtimebin, t.mingravtimebin, t.Ti_Current, t.PM_length, t.PM_start, t.PM_kick], np.int64)
        return scal, np.array(list(t.Ti_kick), np.int64), np.array(list(t.Ti_lastactivedrift), np.int64)

    def is_pm(self):
        return self.t.Ti_Current == self.t.PM_start + self.t.PM_length        # is_PM_timestep timestep.c:153-159

    # --- particles
    def set_particles(self, pos, mass, type, box, vel=None, flags=None, fullacc=None, gravpm=None, bin_grav=None, bin_hydro=None,
                      hsml=None, dthsml=None, hydroacc=None, entropy=None, dtentropy=None, **_):
        n = len(mass)
        z3 = lambda a: np.zeros((n, 3)) if a is None else np.array(a, np.float64, copy=True)
        z1 = lambda a: np.zeros(n) if a is None else np.array(a, np.float64, copy=True)
        zb = lambda a: np.zeros(n, np.uint8) if a is None else np.array(a, np.uint8, copy=True)
        self.n, self.box = n, float(box)
        self.pos, self.mass, self.type = z3(pos), np.array(mass, np.float32, copy=True), zb(type)
        self.vel, self.flags, self.fullacc, self.gravpm = z3(vel), zb(flags), z3(fullacc), z3(gravpm)
        self.bin_grav, self.bin_hydro = zb(bin_grav), zb(bin_hydro)
        self.hsml, self.dthsml, self.hydroacc, self.entropy, self.dtentropy = z1(hsml), z1(dthsml), z3(hydroacc), z1(entropy), z1(dtentropy)
        self.act = None

    def get(self):
        return dict(pos=self.pos, vel=self.vel, hsml=self.hsml, entropy=self.entropy, bin_grav=self.bin_grav, fullacc=self.fullacc)

    def drift(self, ti0, ti1, shift=(0.0, 0.0, 0.0)):
        dd = self.factor(0, ti0, ti1)
        sh = np.ascontiguousarray(shift, np.float64)
        bad = self.L.oracle_drift(C.c_int64(self.n), _p(self.pos), _p(self.vel), _p(self.type), _p(self.flags), _p(self.hsml),
                                  _p(self.dthsml), C.c_double(dd), _p(sh), C.c_double(self.box))
        if bad:
            raise RuntimeError("drift: %d particles with bad Hsml / position" % bad)
        return dd

    def build_active(self):
        self.L.oracle_update_lastactive_drift(C.byref(self.t))
        lst = np.zeros(self.n + 1, np.int32); counts = np.zeros(3, np.int64); self.bincounts = np.zeros((6, NBINS), np.int64)
        nslots = int(((self.type == 0) | (self.type == 5)).sum())
        na = int(self.L.oracle_build_active(C.c_int64(self.n), _p(self.type), _p(self.flags), _p(self.bin_grav), _p(self.bin_hydro),
                                            C.c_int64(self.t.Ti_Current), C.c_int(1 if self.is_pm() else 0), C.c_int64(nslots),
                                            _p(lst), _p(counts), _p(self.bincounts)))
        self.act = None if na < 0 else lst[:na].copy()
        self.counts = counts
        return self.act, counts

    def sublist(self, maxtimebin):
        out = np.zeros(self.n + 1, np.int32)
        nl = self.n if self.act is None else len(self.act)
        na = int(self.L.oracle_active_sublist(_p(self.act), C.c_int64(nl), _p(self.flags), _p(self.bin_grav), C.c_int(maxtimebin),
                                              C.c_int64(self.t.Ti_Current), _p(out)))
        return out[:na].copy()

    def kick_tables(self):
        """gravkick / hydrokick / dt_entr by bin as apply_half_kick builds them (timestep.c:879-891,905-907)."""
        g = np.zeros(NBINS); h = np.zeros(NBINS); e = np.zeros(NBINS)
        t = self.t
        for b in range(NBINS):
            e[b] = self.dloga_from_dti(dti_from_timebin(b) // 2, t.Ti_Current)
            if b < t.mintimebin or not is_timebin_active(b, t.Ti_Current):
                continue
            new = t.Ti_kick[b] + dti_from_timebin(b) // 2
            g[b] = self.factor(1, t.Ti_kick[b], new); h[b] = self.factor(2, t.Ti_kick[b], new)
        return g, h, e

    def kick(self, kind, atime=1.0):
        t = self.t
        if kind in (0, 1):
            g, h, e = self.kick_tables()
            nl = self.n if self.act is None else len(self.act)
            self.L.oracle_half_kick(_p(self.act), C.c_int64(nl), _p(self.type), _p(self.flags), _p(self.bin_grav), _p(self.bin_hydro),
                                    _p(self.vel), _p(self.fullacc), _p(self.hydroacc), _p(self.entropy), _p(self.dtentropy),
                                    _p(g), _p(h), _p(e), C.c_int64(t.Ti_Current), C.c_double(atime), C.c_double(self.sp.MaxGasVel), C.c_int(kind))
        elif kind == 2:                     # apply_PM_half_kick timestep.c:972-993
            tiend = t.PM_kick + t.PM_length // 2
            F = self.factor(1, t.PM_kick, tiend)
            self.L.oracle_pm_kick(C.c_int64(self.n), _p(self.flags), _p(self.vel), _p(self.gravpm), C.c_double(F))
            t.PM_kick = tiend
        else:
            self.L.oracle_update_kick_times(C.byref(t))

    def hydro_timesteps(self, maxsig, atime, first=False):
        ms = np.ascontiguousarray(maxsig, np.float64)
        nl = self.n if self.act is None else len(self.act)
        bad = self.L.oracle_hydro_timebins(C.byref(self.tl), C.byref(self.sp), C.byref(self.t), _p(self.act), C.c_int64(nl), _p(self.type),
                                           _p(self.flags), _p(self.hsml), _p(self.dthsml), _p(ms), _p(self.bin_grav), _p(self.bin_hydro),
                                           C.c_double(atime), C.c_double(float(self.hubble(atime))))
        return int(bad), self.bin_hydro.copy()

    def find_timesteps(self, maxsig, atime, asmth, first=False):
        ms = np.ascontiguousarray(maxsig, np.float64)
        nl = self.n if self.act is None else len(self.act)
        bad = self.L.oracle_find_timesteps(C.byref(self.tl), C.byref(self.cosmo), C.byref(self.sp), C.byref(self.t), C.c_int64(self.n), _p(self.act),
                                           C.c_int64(nl), _p(self.type), _p(self.flags), _p(self.mass), _p(self.vel), _p(self.fullacc), _p(self.gravpm),
                                           _p(self.hsml), _p(self.dthsml), _p(ms), _p(self.bin_grav), _p(self.bin_hydro),
                                           C.c_int(1 if self.is_pm() else 0), C.c_double(atime), C.c_int(2), C.c_double(asmth))
        return int(bad), self.bin_grav.copy(), self.bin_hydro.copy()

    # --- hierarchy
    def set_gravity(self, par, G, nmesh, asmth):
        self.gp = GravShortParams(**par); self.G, self.nmesh, self.asmth = G, nmesh, asmth
        self.sp.softening = 2.8 * par["GravitySoftening"]       # FORCE_SOFTENING gravshort-tree.c:37-41
        return self.sp.softening

    def advance(self, first=False, maxsig=None):
        """One pass of run.c:355-800 (HierarchicalGravity, PM force and hydro accelerations held fixed): see RefStep.advance."""
        L, t = self.L, self.t
        last = t.Ti_Current
        if not first:
            t.Ti_Current = t.Ti_Current + dti_from_timebin(t.mintimebin)      # find_next_kick timestep.c:1324-1328
        atime = self.atime()
        is_pm = self.is_pm()
        if not first:
            self.drift(last, t.Ti_Current)
        act, counts = self.build_active()
        if maxsig is not None:
            self.kick(1, atime)                     # run.c:498-499
        nact = self.n if act is None else len(act)
        store = np.zeros((self.n, 3))
        common = (C.byref(self.tl), C.byref(self.cosmo), C.byref(self.sp), C.byref(self.gp), C.byref(t), C.c_int64(self.n),
                  _p(self.pos), _p(self.mass), _p(self.type), _p(self.flags), _p(self.vel), _p(self.fullacc), _p(self.gravpm), _p(self.bin_grav),
                  _p(act), C.c_int64(nact), C.c_int64(int(counts[1])))
        rc = 0
        if counts[1] > 0:                           # run.c:533
            rc = L.oracle_hier_accelerations(*common, C.c_double(self.G), C.c_int(self.nmesh), C.c_double(self.asmth), C.c_double(self.box), _p(store))
        if rc:
            raise RuntimeError("oracle_hier_accelerations failed")
        self.kick(3)
        if is_pm:
            self.kick(2)
        info = np.zeros(2, np.int64)
        bad = 0
        if counts[1] > 0:
            bad = L.oracle_hier_timesteps(*common, C.c_int(1 if is_pm else 0), C.c_double(self.G), C.c_int(self.nmesh), C.c_double(self.asmth),
                                          C.c_double(self.box), C.c_double(atime), C.c_int(2), _p(store), _p(info))
        if bad < 0:
            raise RuntimeError("oracle_hier_timesteps failed (%d)" % bad)
        if maxsig is not None:                      # run.c:767-773
            b2, _ = self.hydro_timesteps(maxsig, atime, first)
            bad += b2
            self.kick(1, atime)
        self.kick(3)
        if is_pm:
            self.kick(2)
        self.act = None
        return int(bad), np.array([counts[0], counts[1], 1 if is_pm else 0], np.int64)

    def grav_short_tree_active(self):
        """force_tree_full + grav_short_tree for the current active list (run.c:541-548): FullTreeGravAccel of the walked particles"""
        old = self.fullacc + self.gravpm
        t = OracleTree(self.pos, self.mass, self.box, type=self.type, mask=63)
        par = {k: getattr(self.gp, k) for k in ("ErrTolForceAcc", "BHOpeningAngle", "MaxBHOpeningAngle", "TreeUseBH", "Rcut", "GravitySoftening", "rho0")}
        res = t.grav_short_tree(par, self.G, self.nmesh, self.asmth, oldacc=old, active=self.act, full=True)
        acc = res[0]
        idx = np.arange(self.n) if self.act is None else self.act
        self.fullacc[idx] = acc[idx]
        if self.gp.TreeUseBH > 1:
            self.gp.TreeUseBH = 0                                           # gravshort-tree.c:150-151

    def advance_nonsplit(self, asmth, first=False):
        """run.c:355-800 with SplitGravityTimestepsOn = 0, PM force held fixed: see RefStep.advance_nonsplit"""
        t = self.t
        last = t.Ti_Current
        if not first:
            t.Ti_Current = t.Ti_Current + dti_from_timebin(t.mintimebin)
        atime = self.atime()
        is_pm = self.is_pm()
        if not first:
            self.drift(last, t.Ti_Current)
        act, counts = self.build_active()
        self.grav_short_tree_active()
        self.kick(0, atime); self.kick(3)
        if is_pm:
            self.kick(2)
        bad, _, _ = self.find_timesteps(np.zeros(self.n), atime, asmth, first)
        self.kick(0, atime); self.kick(3)
        if is_pm:
            self.kick(2)
        self.act = None
        return bad, np.array([counts[0], counts[1], 1 if is_pm else 0], np.int64)
