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
            arena_gib = max(2.0, min(24.0, 0.4 * avail))
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
