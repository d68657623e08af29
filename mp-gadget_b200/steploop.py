"""Host side of the device-resident step loop (include/b200force.h, b200_step_*): the harness mirror
of the reference's run.c:355-800 loop for HierarchicalGravity.  The host keeps what the reference's
host keeps -- DriftKickTimes, the sync-point table and the cosmology integrals -- and the particle
state (positions, velocities, time bins) stays on the GPU between force computations.

The cosmology enters through two callables, exactly the two things the reference's timestep.c asks
of cosmology.c / timefac.c: factor(kind, ti0, ti1) (kind 0 drift, 1 gravkick, 2 hydrokick:
get_exact_*_factor, timefac.c:58-73) and hubble(a) (hubble_function).  There is no CPU path here:
every particle loop runs in csrc/steploop.cu."""
import ctypes as C
import numpy as np
from . import B200Error, GravShortParams, _p, _c

TIMEBINS = 46
NBINS = TIMEBINS + 1
TIMEBASE = 1 << TIMEBINS


class StepState(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("vel", "fullacc", "gravpm", "bin_grav", "bin_hydro", "flags", "hsml", "dthsml",
                                          "hydroacc", "entropy", "dtentropy")] + [("BoxSize", C.c_double)]


class StepStateOut(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("pos", "vel", "fullacc", "hsml", "entropy", "bin_grav", "bin_hydro", "hydroacc", "dtentropy",
                                          "maxsignalvel")]


class StepTimes(C.Structure):
    """DriftKickTimes, libgadget/timestep.h:10-26"""
    _fields_ = [("mintimebin", C.c_int32), ("maxtimebin", C.c_int32), ("mingravtimebin", C.c_int32), ("pad_", C.c_int32),
                ("Ti_kick", C.c_int64 * NBINS), ("Ti_lastactivedrift", C.c_int64 * NBINS),
                ("Ti_Current", C.c_int64), ("PM_length", C.c_int64), ("PM_start", C.c_int64), ("PM_kick", C.c_int64)]


KICKFN = C.CFUNCTYPE(C.c_double, C.c_void_p, C.c_int64, C.c_int64)


class StepParams(C.Structure):
    _fields_ = [("ErrTolIntAccuracy", C.c_double), ("MaxSizeTimestep", C.c_double), ("MinSizeTimestep", C.c_double),
                ("MaxRMSDisplacementFac", C.c_double), ("CourantFac", C.c_double), ("softening", C.c_double), ("omega_type", C.c_double * 6), ("RhoCrit", C.c_double),
                ("FastParticleType", C.c_int32), ("pad_", C.c_int32), ("sync_loga", C.c_void_p), ("nsync", C.c_int64),
                ("gravkick_factor", KICKFN), ("user", C.c_void_p)]


def dti_from_timebin(b):
    return (1 << int(b)) if b > 0 else 0


def is_timebin_active(b, ti):          # timestep.c:143-150
    return b <= 0 or ti <= 0 or ti % dti_from_timebin(b) == 0


class StepEngine:
    """Same interface as oracle.ref.RefStep / oracle.step.StepOracle, running on a b200 Engine."""

    def __init__(self, engine, sync_loga, factor, hubble, Omega0=0.288, OmegaBaryon=0.0472, Hubble=0.1, G=43.0071,
                 ErrTolIntAccuracy=0.02, MaxGasVel=3e5, MaxSizeTimestep=0.1, MinSizeTimestep=0.0, MaxRMSDisplacementFac=0.2, CourantFac=0.15, **_):
        self.e, self.L, self.ctx = engine, engine.L, engine.ctx
        self.sync = np.ascontiguousarray(sync_loga, np.float64)
        self.factor, self.hubble = factor, hubble
        self.MaxGasVel = MaxGasVel
        self._kick_cb = KICKFN(lambda user, t0, t1: float(self.factor(1, int(t0), int(t1))))
        sp = StepParams()
        sp.ErrTolIntAccuracy, sp.MaxSizeTimestep, sp.MinSizeTimestep, sp.MaxRMSDisplacementFac = ErrTolIntAccuracy, MaxSizeTimestep, MinSizeTimestep, MaxRMSDisplacementFac
        om = [OmegaBaryon, Omega0 - OmegaBaryon, 0.0, Omega0 - OmegaBaryon, OmegaBaryon, OmegaBaryon]       # timestep.c:1251-1263
        for k in range(6):
            sp.omega_type[k] = om[k]
        sp.CourantFac = CourantFac
        sp.RhoCrit = 3 * Hubble * Hubble / (8 * np.pi * G)
        sp.FastParticleType = 2
        sp.sync_loga = self.sync.ctypes.data; sp.nsync = len(self.sync)
        sp.gravkick_factor = self._kick_cb
        self.sp = sp
        self.t = StepTimes()
        self.counts = np.zeros(3, np.int64)
        self.gas_slots = 0

    def _ck(self, rc):
        if rc != 0:
            raise B200Error(self.L.b200_last_error(self.ctx).decode())

    # --- the integer timeline on the host (timebinmgr.c:380-447), scalars only
    def _interval(self, ti):
        s = ti >> TIMEBINS
        return 0.0 if s >= len(self.sync) - 1 else (self.sync[s + 1] - self.sync[s]) / TIMEBASE

    def loga_from_ti(self, ti):
        return float(self.sync[ti >> TIMEBINS] + (ti & (TIMEBASE - 1)) * self._interval(ti))

    def dloga_from_dti(self, dti, ti):
        return self._interval(ti) * dti

    def atime(self):
        return float(np.exp(self.loga_from_ti(self.t.Ti_Current)))

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

    def is_pm(self):                    # is_PM_timestep timestep.c:153-159
        return self.t.Ti_Current == self.t.PM_start + self.t.PM_length

    # --- particles
    def set_particles(self, pos, mass, type, box, vel=None, flags=None, fullacc=None, gravpm=None, bin_grav=None, bin_hydro=None,
                      hsml=None, dthsml=None, hydroacc=None, entropy=None, dtentropy=None, **_):
        self.n, self.box = len(mass), float(box)
        self.e.set_particles(pos, mass, type=type)
        f = lambda a: _c(a, np.float64)
        b = lambda a: _c(a, np.uint8)
        keep = dict(vel=f(vel), fullacc=f(fullacc), gravpm=f(gravpm), bin_grav=b(bin_grav), bin_hydro=b(bin_hydro), flags=b(flags),
                    hsml=f(hsml), dthsml=f(dthsml), hydroacc=f(hydroacc), entropy=f(entropy), dtentropy=f(dtentropy))
        st = StepState(**{k: (None if v is None else v.ctypes.data) for k, v in keep.items()})
        st.BoxSize = self.box
        self._ck(self.L.b200_step_set_state(self.ctx, C.byref(st)))
        ty = np.asarray(type)
        self.gas_slots = int(((ty == 0) | (ty == 5)).sum())

    def get(self):
        n = self.n
        out = dict(pos=np.zeros((n, 3)), vel=np.zeros((n, 3)), fullacc=np.zeros((n, 3)), hsml=np.zeros(n), entropy=np.zeros(n),
                   bin_grav=np.zeros(n, np.uint8), bin_hydro=np.zeros(n, np.uint8))
        so = StepStateOut(**{k: v.ctypes.data for k, v in out.items()})
        self._ck(self.L.b200_step_get_state(self.ctx, C.byref(so)))
        return out

    def drift(self, ti0, ti1, shift=(0.0, 0.0, 0.0)):
        dd = float(self.factor(0, ti0, ti1))
        sh = np.ascontiguousarray(shift, np.float64)
        nbad = C.c_int64()
        self._ck(self.L.b200_step_drift(self.ctx, C.c_double(dd), _p(sh), C.byref(nbad)))
        return dd

    def build_active(self):
        t = self.t
        for b in range(NBINS):          # update_lastactive_drift timestep.c:860-871
            if is_timebin_active(b, t.Ti_Current):
                t.Ti_lastactivedrift[b] = t.Ti_Current
        counts = np.zeros(3, np.int64); self.bincounts = np.zeros((6, NBINS), np.int64)
        self._ck(self.L.b200_step_build_active(self.ctx, C.c_int64(t.Ti_Current), C.c_int(1 if self.is_pm() else 0), C.c_int64(self.gas_slots),
                                               _p(counts), _p(self.bincounts)))
        self.counts = counts
        return self.active_list(0), counts

    def active_list(self, which):
        n = C.c_int64()
        self._ck(self.L.b200_step_get_active(self.ctx, C.c_int(which), None, C.byref(n)))
        if n.value < 0:
            return None
        out = np.zeros(max(n.value, 1), np.int32)
        self._ck(self.L.b200_step_get_active(self.ctx, C.c_int(which), _p(out), C.byref(n)))
        return out[:n.value].copy()

    def sublist(self, maxtimebin):
        ns = C.c_int64()
        self._ck(self.L.b200_step_active_sublist(self.ctx, C.c_int(maxtimebin), C.c_int64(self.t.Ti_Current), C.byref(ns)))
        return self.active_list(1)

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
            self._ck(self.L.b200_step_half_kick(self.ctx, _p(g), _p(h), _p(e), C.c_int64(t.Ti_Current), C.c_double(atime),
                                                C.c_double(self.MaxGasVel), C.c_int(kind)))
        elif kind == 2:                 # apply_PM_half_kick timestep.c:972-993
            tiend = t.PM_kick + t.PM_length // 2
            self._ck(self.L.b200_step_pm_kick(self.ctx, C.c_double(float(self.factor(1, t.PM_kick, tiend)))))
            t.PM_kick = tiend
        else:                           # update_kick_times timestep.c:215-235
            if t.mintimebin == 0 and t.maxtimebin == 0:
                return
            for b in range(t.mintimebin, NBINS):
                if is_timebin_active(b, t.Ti_Current):
                    t.Ti_kick[b] += dti_from_timebin(b) // 2
            for b in range(1, t.mintimebin):
                t.Ti_kick[b] += dti_from_timebin(t.mintimebin) // 2

    # --- gas sub-step on the device (run.c:466-495): active list + tables to the SPH module, hydro results back
    def sph_prepare(self, tables):
        """tables: dict of per-bin arrays gravkick, hydrokick, dloga_pred, drift, dloga_bin (b200_sph_bins)"""
        t = np.zeros((5, NBINS))
        for r, k in enumerate(("gravkick", "hydrokick", "dloga_pred", "drift", "dloga_bin")):
            v = np.asarray(tables[k], dtype=np.float64)[:NBINS]
            t[r, :len(v)] = v
        self._ck(self.L.b200_step_sph_prepare(self.ctx, _p(t)))

    def adopt_hydro(self):
        """-> dict(hydroacc, dtentropy, maxsignalvel) as now held in the step state"""
        self._ck(self.L.b200_step_adopt_hydro(self.ctx))
        n = self.n
        out = dict(hydroacc=np.zeros((n, 3)), dtentropy=np.zeros(n), maxsignalvel=np.zeros(n))
        so = StepStateOut(**{k: v.ctypes.data for k, v in out.items()})
        self._ck(self.L.b200_step_get_state(self.ctx, C.byref(so)))
        return out

    def hydro_timesteps(self, maxsig, atime, first=False, fetch=True):
        """find_hydro_timesteps on the current active list -> (bad count, TimeBinHydro[n]).  maxsig = None: the
        SphP[].MaxSignalVel already in the step state (adopt_hydro); fetch = False: only the count crosses PCIe (the loop
        drivers), the bins stay on the device."""
        ms = _c(maxsig, np.float64)
        nbad = C.c_int64()
        self._ck(self.L.b200_step_hydro_timesteps(self.ctx, C.byref(self.sp), C.byref(self.t), _p(ms), C.c_double(atime),
                                                  C.c_double(float(self.hubble(atime))), C.byref(nbad)))
        return int(nbad.value), (self.get()["bin_hydro"] if fetch else None)

    def find_timesteps(self, maxsig, atime, asmth=None, first=False, fetch=True):
        """find_timesteps (SplitGravityTimestepsOn = 0) on the current active list -> (bad, TimeBinGravity[n], TimeBinHydro[n]);
        the PM smoothing scale comes from gravpm_init_periodic (set_gravity).  maxsig = None and fetch = False as in
        hydro_timesteps: nothing of size n crosses PCIe."""
        ms = _c(maxsig, np.float64)
        nbad = C.c_int64()
        self._ck(self.L.b200_step_find_timesteps(self.ctx, C.byref(self.sp), C.byref(self.t), _p(ms), C.c_int(1 if self.is_pm() else 0),
                                                 C.c_double(atime), C.c_double(float(self.hubble(atime))), C.byref(nbad)))
        if not fetch:
            return int(nbad.value), None, None
        g = self.get()
        return int(nbad.value), g["bin_grav"], g["bin_hydro"]

    # --- hierarchy
    def set_gravity(self, par, G, nmesh, asmth):
        self.e.gravpm_init_periodic(self.box, asmth, nmesh, G)
        self.gp = GravShortParams(**par)
        self.sp.softening = 2.8 * par["GravitySoftening"]           # FORCE_SOFTENING gravshort-tree.c:37-41
        return self.sp.softening

    def pm_force(self):
        """gravpm_force on a PM step (run.c:519-523) without leaving the device: b200_pm_force_dev, then P[].GravPM of
        the step state <- its result."""
        self.e.gravpm_force_dev()
        self._ck(self.L.b200_step_adopt_forces(self.ctx, C.c_int(0), C.c_int(1)))

    def advance(self, first=False, pm=False, maxsig=None):
        """One pass of run.c:355-800 (collisionless, HierarchicalGravity).  pm = False holds the PM force fixed (the
        parity scenarios); pm = True recomputes it on PM steps."""
        t = self.t
        last = t.Ti_Current
        if not first:
            t.Ti_Current = t.Ti_Current + dti_from_timebin(t.mintimebin)          # find_next_kick timestep.c:1324-1328
        atime = self.atime()
        is_pm = self.is_pm()
        prof = getattr(self, "prof", None)          # optional {stage: wall ms} of this pass (every stage ends with a scalar read-back)
        import time as _time
        def stage(name, t0):
            if prof is not None:
                prof[name] = prof.get(name, 0.0) + 1e3 * (_time.perf_counter() - t0)
            return _time.perf_counter()
        tt = _time.perf_counter()
        if not first:
            self.drift(last, t.Ti_Current)
        _, counts = self.build_active()
        tt = stage("drift+active", tt)
        if maxsig is not None:          # gas takes part with its hydro accelerations held fixed: closing hydro kick, run.c:498-499
            self.kick(1, atime)
        if pm and is_pm:
            self.pm_force()
        tt = stage("pm_force", tt)
        if counts[1] > 0:               # run.c:533
            self._ck(self.L.b200_step_hier_accelerations(self.ctx, C.byref(self.sp), C.byref(self.gp), C.byref(t), C.c_int64(int(counts[1]))))
        tt = stage("hier_accelerations", tt)
        self.kick(3)
        if is_pm:
            self.kick(2)
        tt = stage("kicks", tt)
        info = np.zeros(3, np.int64)
        if counts[1] > 0:
            self._ck(self.L.b200_step_hier_timesteps(self.ctx, C.byref(self.sp), C.byref(self.gp), C.byref(t), C.c_int64(int(counts[1])),
                                                     C.c_int(1 if is_pm else 0), C.c_double(atime), C.c_double(float(self.hubble(atime))), _p(info)))
        tt = stage("hier_timesteps", tt)
        if maxsig is not None:          # run.c:767-773
            b2, _ = self.hydro_timesteps(maxsig, atime, first, fetch=False)
            info[2] += b2
            self.kick(1, atime)
        self.kick(3)
        if is_pm:
            self.kick(2)
        stage("kicks", tt)
        return int(info[2]), np.array([counts[0], counts[1], 1 if is_pm else 0], np.int64)

    def advance_nonsplit(self, asmth=None, first=False, pm=False):
        """One pass of run.c:355-800 with SplitGravityTimestepsOn = 0: one full tree, the walk for the active particles,
        closing half kick, find_timesteps, opening half kick."""
        t = self.t
        last = t.Ti_Current
        if not first:
            t.Ti_Current = t.Ti_Current + dti_from_timebin(t.mintimebin)
        atime = self.atime()
        is_pm = self.is_pm()
        if not first:
            self.drift(last, t.Ti_Current)
        _, counts = self.build_active()
        if pm and is_pm:
            self.pm_force()
        self._ck(self.L.b200_step_grav_short_tree(self.ctx, C.byref(self.gp)))
        self.kick(0, atime); self.kick(3)
        if is_pm:
            self.kick(2)
        bad, _, _ = self.find_timesteps(None, atime, first=first, fetch=False)      # no gas criterion input: the state's own MaxSignalVel
        self.kick(0, atime); self.kick(3)
        if is_pm:
            self.kick(2)
        return bad, np.array([counts[0], counts[1], 1 if is_pm else 0], np.int64)
