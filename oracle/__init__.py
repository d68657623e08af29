"""ctypes front-end of the CPU oracle (oracle/*.c) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module.  See oracle/oracle.h.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class Node(C.Structure):
    _fields_ = [("sibling", C.c_int32), ("father", C.c_int32), ("firstchild", C.c_int32),
                ("nocc", C.c_int32), ("part", C.c_int32 * 8), ("toplevel", C.c_int32),
                ("level", C.c_int32), ("len", C.c_double), ("center", C.c_double * 3),
                ("cofm", C.c_double * 3), ("mass", C.c_double), ("hmax", C.c_double)]


NODE_DTYPE = np.dtype([("sibling", "i4"), ("father", "i4"), ("firstchild", "i4"), ("nocc", "i4"),
                       ("part", "i4", (8,)), ("toplevel", "i4"), ("level", "i4"), ("len", "f8"),
                       ("center", "f8", (3,)), ("cofm", "f8", (3,)), ("mass", "f8"), ("hmax", "f8")],
                      align=True)
assert NODE_DTYPE.itemsize == C.sizeof(Node)


class Tree(C.Structure):
    _fields_ = [("nodes", C.POINTER(Node)), ("numnodes", C.c_int64),
                ("numparticles", C.c_int64), ("BoxSize", C.c_double)]


class GravShortParams(C.Structure):
    _fields_ = [("ErrTolForceAcc", C.c_double), ("BHOpeningAngle", C.c_double),
                ("MaxBHOpeningAngle", C.c_double), ("TreeUseBH", C.c_int32), ("pad_", C.c_int32),
                ("Rcut", C.c_double), ("GravitySoftening", C.c_double), ("rho0", C.c_double)]


COUNTS_DTYPE = np.dtype([("nodes_accepted", "i4"), ("nodes_opened", "i4"),
                         ("nodes_discarded", "i4"), ("particles", "i4")])


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("oracle_tree.c", "oracle_pm.c", "oracle_sph.c", "oracle_step.c", "oracle_domain.c", "oracle_fof.c", "oracle.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.oracle_tree_build.restype = C.c_int
        _LIB.oracle_grav_short_tree.restype = C.c_int
    return _LIB


def _p(a, ty=C.c_void_p):
    return None if a is None else a.ctypes.data_as(ty)


def _c(a, dt):
    return None if a is None else np.ascontiguousarray(a, dtype=dt)


class OracleTree:
    """oracle_tree_build wrapper; .nodes is a structured numpy view in DFS order."""

    def __init__(self, pos, mass, box, type=None, hsml=None, mask=63, active=None, toplevel_depth=0, top_daughter=None):
        self.pos = _c(pos, np.float64)
        self.mass = _c(mass, np.float32)
        self.type = _c(type, np.uint8)
        self.hsml = _c(hsml, np.float64)
        self.active = _c(active, np.int32)
        self.n = len(self.mass)
        self.t = Tree()
        na = 0 if self.active is None else len(self.active)
        # top_daughter: the Daughter column of an arbitrary domain top tree (forcetree.c:654-687) instead of a uniform depth
        self.top_daughter = _c(top_daughter, np.int32)
        rc = lib().oracle_tree_build_top(C.byref(self.t), _p(self.pos), _p(self.mass), _p(self.type),
                                         _p(self.hsml), C.c_int64(self.n), C.c_double(box), C.c_int(mask),
                                         _p(self.active), C.c_int64(na), C.c_int(toplevel_depth),
                                         _p(self.top_daughter), C.c_int32(0 if self.top_daughter is None else len(self.top_daughter)))
        if rc:
            raise RuntimeError("oracle_tree_build failed (coincident particles?)")
        buf = (Node * self.t.numnodes).from_address(C.addressof(self.t.nodes.contents))
        self.nodes = np.frombuffer(buf, dtype=NODE_DTYPE)

    def __del__(self):
        try:
            self.nodes = None
            lib().oracle_tree_free(C.byref(self.t))
        except Exception:
            pass

    def grav_short_tree(self, par, G, Nmesh, Asmth, oldacc=None, active=None, full=True):
        n = self.n
        p = GravShortParams(**par) if isinstance(par, dict) else par
        acc = np.zeros((n, 3))
        pot = np.zeros(n)
        cnt = np.zeros(n, dtype=COUNTS_DTYPE)
        oldacc = _c(oldacc, np.float64)
        active = _c(active, np.int32)
        na = 0 if active is None else len(active)
        rc = lib().oracle_grav_short_tree(C.byref(self.t), _p(self.pos), _p(self.mass), C.c_int64(n),
                                          C.byref(p), C.c_double(G), C.c_int(Nmesh), C.c_double(Asmth),
                                          _p(oldacc), _p(active), C.c_int64(na), C.c_int(1 if full else 0),
                                          _p(acc), _p(pot), _p(cnt))
        if rc:
            raise MemoryError
        return acc, pot, cnt


class SphParams(C.Structure):
    _fields_ = [("KernelType", C.c_int32), ("DensityIndependentSphOn", C.c_int32),
                ("DensityResolutionEta", C.c_double), ("MaxNumNgbDeviation", C.c_double), ("MinGasHsml", C.c_double),
                ("ArtBulkViscConst", C.c_double), ("DensityContrastLimit", C.c_double),
                ("gravkick", C.c_double), ("hydrokick", C.c_double), ("pmkick", C.c_double),
                ("dloga_pred", C.c_double), ("drift", C.c_double), ("dloga_bin", C.c_double),
                ("atime", C.c_double), ("hubble", C.c_double)]


SPH_DEFAULTS = dict(KernelType=2, DensityIndependentSphOn=1, DensityResolutionEta=1.0, MaxNumNgbDeviation=2.0,
                    MinGasHsml=0.0, ArtBulkViscConst=0.75, DensityContrastLimit=100.0, gravkick=0.0, hydrokick=0.0,
                    pmkick=0.0, dloga_pred=0.0, drift=0.0, dloga_bin=0.0, atime=1.0, hubble=0.1)


def sph_params(**kw):
    d = dict(SPH_DEFAULTS)
    d.update(kw)
    return SphParams(**d)


def set_init_hsml(tree, kerneltype, eta, mean_gas_separation):
    h = np.zeros(tree.n)
    lib().oracle_set_init_hsml(C.byref(tree.t), _p(tree.mass), _p(tree.type), C.c_int64(tree.n), C.c_int(kerneltype),
                               C.c_double(eta), C.c_double(mean_gas_separation), _p(h))
    return h


_mixed_keep = []


def sph_set_mixed(bin_grav=None, bin_hydro=None, tables=None, active=None, n=None):
    """Install (or with no arguments clear) the mixed-time-bin context of the following density() /
    hydro() calls: per-particle bins, per-bin tables gravkick, hydrokick, dloga_pred, drift, dloga_bin
    (<= 47 entries each) and the active particle indices."""
    del _mixed_keep[:]
    if bin_grav is None:
        lib().oracle_sph_set_mixed(None, None, None, None)
        return
    bg = _c(bin_grav, np.uint8); bh = _c(bin_hydro, np.uint8)
    tab = np.zeros((5, 47))
    for r, k in enumerate(("gravkick", "hydrokick", "dloga_pred", "drift", "dloga_bin")):
        v = np.asarray(tables[k], dtype=np.float64)[:47]
        tab[r, :len(v)] = v
    flags = None
    if active is not None:
        flags = np.zeros(len(bg) if n is None else n, np.uint8)
        flags[np.asarray(active)] = 1
    _mixed_keep.extend([bg, bh, tab, flags])        # the C side keeps the pointers
    lib().oracle_sph_set_mixed(_p(bg), _p(bh), _p(tab), _p(flags))


def density(tree, sp, hsml, update_hsml=1, DoEgyDensity=0, vel=None, entropy=None, dtentropy=None,
            fullacc=None, gravpm=None, hydroacc=None, state=None):
    """oracle_density on the particles of an OracleTree (gas tree). Returns a dict.
    state: dict of arrays the outputs start from (the values particles outside the active set keep)."""
    n = tree.n
    f8 = lambda a: _c(a, np.float64)
    out = dict(hsml=np.array(hsml, dtype=np.float64, copy=True), density=np.zeros(n), egywtdensity=np.zeros(n),
               dhsmlfac=np.zeros(n), divvel=np.zeros(n), curlvel=np.zeros(n), dthsml=np.zeros(n), numngb=np.zeros(n),
               ninteract=np.zeros(n, np.int32), niter=np.zeros(n, np.int32), entvarpred=np.zeros(n))
    if state is not None:
        for k in ("density", "egywtdensity", "dhsmlfac", "divvel", "curlvel", "dthsml"):
            if k in state:
                out[k] = np.array(state[k], dtype=np.float64, copy=True)
    vel, entropy, dtentropy, fullacc, gravpm, hydroacc = map(f8, (vel, entropy, dtentropy, fullacc, gravpm, hydroacc))
    lib().oracle_density.restype = C.c_int
    rc = lib().oracle_density(C.byref(tree.t), _p(tree.pos), _p(tree.mass), _p(tree.type), C.c_int64(n), C.byref(sp),
                              C.c_int(update_hsml), C.c_int(DoEgyDensity), _p(vel), _p(fullacc), _p(gravpm), _p(hydroacc),
                              _p(entropy), _p(dtentropy), _p(out["hsml"]), _p(out["density"]), _p(out["egywtdensity"]),
                              _p(out["dhsmlfac"]), _p(out["divvel"]), _p(out["curlvel"]), _p(out["dthsml"]), _p(out["numngb"]),
                              _p(out["ninteract"]), _p(out["niter"]), _p(out["entvarpred"]))
    out["rc"] = rc
    return out


def hydro(tree, sp, dens, vel=None, entropy=None, dtentropy=None, fullacc=None, gravpm=None, hydroacc=None):
    """oracle_hydro after density(); dens is the dict density() returned."""
    n = tree.n
    f8 = lambda a: _c(a, np.float64)
    vel, entropy, dtentropy, fullacc, gravpm, hydroacc = map(f8, (vel, entropy, dtentropy, fullacc, gravpm, hydroacc))
    out = dict(acc=np.zeros((n, 3)), dtentropy=np.zeros(n), maxsignalvel=np.zeros(n), ninteract=np.zeros(n, np.int32))
    lib().oracle_hydro(C.byref(tree.t), _p(tree.pos), _p(tree.mass), _p(tree.type), C.c_int64(n), C.byref(sp),
                       _p(vel), _p(fullacc), _p(gravpm), _p(hydroacc), _p(entropy), _p(dtentropy),
                       _p(dens["hsml"]), _p(dens["density"]), _p(dens["egywtdensity"]), _p(dens["dhsmlfac"]),
                       _p(dens["divvel"]), _p(dens["curlvel"]),
                       _p(out["acc"]), _p(out["dtentropy"]), _p(out["maxsignalvel"]), _p(out["ninteract"]))
    return out


def pm_power(pos, mass, box, nmesh, workers=-1):
    """Raw power-spectrum sums of gravpm_force's side effect (powerspectrum_add_mode,
    gravpm.c:330-361) = (Power[nmesh], kk[nmesh], Nmodes[nmesh], Norm)."""
    import scipy.fft as sfft
    pos = _c(pos, np.float64)
    mass = _c(mass, np.float32)
    n = len(mass)
    L = lib()
    mesh = np.zeros((nmesh, nmesh, nmesh))
    icell = np.zeros((n, 3), dtype=np.int32)
    L.oracle_pm_deposit(_p(pos), _p(mass), C.c_int64(n), C.c_double(box), C.c_int(nmesh), _p(mesh), _p(icell))
    rhok = np.ascontiguousarray(sfft.rfftn(mesh, workers=workers))
    pw, kk, nm, norm = np.zeros(nmesh), np.zeros(nmesh), np.zeros(nmesh, np.int64), C.c_double()
    L.oracle_pm_power(_p(rhok), C.c_int(nmesh), _p(pw), _p(kk), _p(nm), C.byref(norm))
    return pw, kk, nm, norm.value


def pm_force(pos, mass, box, nmesh, asmth, G, workers=-1, return_mesh=False):
    """gravpm_force restated: deposit -> rfftn -> potential_transfer ->
    {copy, force_transfer_d} -> irfftn (unnormalised) -> readout.
    Returns (gravpm[n,3], potential[n], icell[n,3])."""
    import scipy.fft as sfft
    pos = _c(pos, np.float64)
    mass = _c(mass, np.float32)
    n = len(mass)
    L = lib()
    mesh = np.zeros((nmesh, nmesh, nmesh))
    icell = np.zeros((n, 3), dtype=np.int32)
    L.oracle_pm_deposit(_p(pos), _p(mass), C.c_int64(n), C.c_double(box), C.c_int(nmesh), _p(mesh), _p(icell))
    dens = mesh.copy() if return_mesh else None
    rhok = sfft.rfftn(mesh, workers=workers)                      # unnormalised forward
    rhok = np.ascontiguousarray(rhok)
    L.oracle_pm_potential_transfer(_p(rhok), C.c_int(nmesh), C.c_double(box), C.c_double(asmth), C.c_double(G))
    ntot = float(nmesh) ** 3
    out = np.zeros((n, 4))
    potmesh = None
    for f in range(4):
        if f == 0:
            fk = rhok
        else:
            fk = np.empty_like(rhok)
            L.oracle_pm_force_transfer(_p(rhok), _p(fk), C.c_int(nmesh), C.c_double(box), C.c_int(f - 1))
        real = sfft.irfftn(fk, s=(nmesh,) * 3, workers=workers, norm="forward")  # unnormalised inverse
        real = np.ascontiguousarray(real)
        if f == 0 and return_mesh:
            potmesh = real.copy()
        L.oracle_pm_readout(_p(real), _p(pos), C.c_int64(n), C.c_double(box), C.c_int(nmesh),
                            out[:, f].ctypes.data_as(C.c_void_p), C.c_int64(4))
    del ntot
    res = (np.ascontiguousarray(out[:, 1:4]), np.ascontiguousarray(out[:, 0]), icell)
    if return_mesh:
        return res + (dens, potmesh)
    return res


def pm_c2r_readout(pos, box, nmesh, rho_k, functions, workers=-1):
    """petapm_force_c2r (petapm.c:326-362) restated for a caller's spectrum: for every (kind, table) the transfer of
    pm_apply_transfer_function (petapm.c:1092-1132) in the forms of libgenic/zeldovich.c:276-313 -- kind 0: value *= table[k2];
    kind 1..3: fac = table[k2] * k_axis, (re, im) <- (-im fac, re fac); k2 = 0 unchanged --, the unnormalised inverse
    transform and the CIC read-out of oracle_pm_readout.  rho_k: complex [nmesh, nmesh, nmesh//2+1], x slowest."""
    import scipy.fft as sfft
    pos = _c(pos, np.float64)
    n = len(pos)
    N = nmesh
    k = np.arange(N)
    k = np.where(k <= N // 2, k, k - N)
    kx, ky, kz = k[:, None, None], k[None, :, None], np.arange(N // 2 + 1)[None, None, :]
    k2 = kx * kx + ky * ky + kz * kz
    L = lib()
    outs = []
    for kind, table in functions:
        table = np.asarray(table, np.float64)
        fac = table[k2]
        if kind == 0:
            fk = np.where(k2 > 0, rho_k * fac, rho_k)
        else:
            fac = fac * (kx, ky, kz)[kind - 1]
            fk = np.where(k2 > 0, (-rho_k.imag * fac) + 1j * (rho_k.real * fac), rho_k)
        real = np.ascontiguousarray(sfft.irfftn(np.ascontiguousarray(fk), s=(N,) * 3, workers=workers, norm="forward"))
        out = np.zeros(n)
        L.oracle_pm_readout(_p(real), _p(pos), C.c_int64(n), C.c_double(box), C.c_int(N), out.ctypes.data_as(C.c_void_p), C.c_int64(1))
        outs.append(out)
    return outs


def direct_sum(pos, mass, box, G, softening_h, repeat=1):
    pos = _c(pos, np.float64)
    mass = _c(mass, np.float32)
    acc = np.zeros((len(mass), 3))
    lib().oracle_direct_sum(_p(pos), _p(mass), C.c_int64(len(mass)), C.c_double(box), C.c_double(G),
                            C.c_double(softening_h), C.c_int(repeat), _p(acc))
    return acc


def peano_keys(pos, box):
    """PEANO(Pos, BoxSize) of utils/peano.h:15-21 for every row of pos -> uint64 keys."""
    pos = _c(pos, np.float64)
    keys = np.zeros(len(pos), np.uint64)
    lib().oracle_peano_keys(_p(pos), C.c_int64(len(pos)), C.c_double(box), _p(keys))
    return keys


def peano_key(x, y, z, bits=21):
    L = lib(); L.oracle_peano_key.restype = C.c_uint64
    return int(L.oracle_peano_key(C.c_int(x), C.c_int(y), C.c_int(z), C.c_int(bits)))


def topleaf(keys, daughter, startkey, shift, leaf):
    """domain_get_topleaf (domain.h:71-78) over TopNodes given as arrays."""
    keys = _c(keys, np.uint64); out = np.zeros(len(keys), np.int32)
    lib().oracle_topleaf(_p(keys), C.c_int64(len(keys)), _p(_c(daughter, np.int32)), _p(_c(startkey, np.uint64)), _p(_c(shift, np.int32)),
                         _p(_c(leaf, np.int32)), _p(out))
    return out


def leaf_counts(topleaf, nleaf, flags=None):
    """TopLeafCount of domain_compute_costs (domain.c:1398-1470)"""
    tl = _c(topleaf, np.int32); out = np.zeros(nleaf, np.int64)
    lib().oracle_leaf_counts(_p(tl), _p(_c(flags, np.uint8)), C.c_int64(len(tl)), C.c_int32(nleaf), _p(out))
    return out


def domain_assign_balanced(ntask, cost, nseg_per_task=1):
    """domain_assign_topleaves_balanced (domain.c:610-755) over leaves in key order -> task per leaf (None where the
    reference would stop with an error)."""
    cost = _c(cost, np.int64); task = np.zeros(len(cost), np.int32)
    rc = lib().oracle_domain_assign_balanced(C.c_int(ntask), C.c_int32(len(cost)), _p(cost), C.c_int(nseg_per_task), _p(task))
    return None if rc < 0 else task


TOPNODE_DTYPE = np.dtype([("StartKey", "u8"), ("Shift", "i4"), ("Daughter", "i4"), ("Parent", "i4"), ("pad_", "i4"), ("Count", "i8"), ("Cost", "i8")])


class TopTree:
    """The top tree of the domain decomposition stage by stage (domain.c:826-1395); .nodes[:size] is the tree."""

    def __init__(self, maxnodes):
        self.nodes = np.zeros(maxnodes, TOPNODE_DTYPE); self.size = C.c_int32(0); self.maxnodes = maxnodes

    @property
    def tree(self):
        return self.nodes[:self.size.value]

    def local(self, sample_keys):
        k = np.sort(_c(sample_keys, np.uint64))
        return lib().oracle_toptree_local(_p(k), None, C.c_int64(len(k)), _p(self.nodes), C.byref(self.size), C.c_int32(self.maxnodes))

    def truncate(self, countlimit, costlimit):
        lib().oracle_toptree_truncate(_p(self.nodes), C.byref(self.size), C.c_int64(countlimit), C.c_int64(costlimit))

    def merge(self, other):
        return lib().oracle_toptree_merge(_p(self.nodes), C.byref(self.size), _p(other.nodes), C.c_int32(self.maxnodes))

    def global_refine(self, countlimit, costlimit):
        return lib().oracle_toptree_global_refine(_p(self.nodes), C.byref(self.size), C.c_int32(self.maxnodes), C.c_int64(countlimit), C.c_int64(costlimit))

    def leaves(self):
        leaf = np.zeros(self.size.value, np.int32)
        nl = lib().oracle_toptree_leaves(_p(self.nodes), self.size, _p(leaf))
        return nl, leaf


def exchange_plan(type, flags, topleaf, task_of_leaf, ntask, thistask):
    """exchange.c:408-444,505-530 -> (exchange list, togo[ntask][7], ngarbage)"""
    type = _c(type, np.uint8); flags = _c(flags, np.uint8); tl = _c(topleaf, np.int32); tk = _c(task_of_leaf, np.int32)
    lst = np.zeros(len(tl) + 1, np.int32); togo = np.zeros((ntask, 7), np.int64); ng = C.c_int64()
    L = lib(); L.oracle_exchange_plan.restype = C.c_int64
    nex = L.oracle_exchange_plan(C.c_int64(len(tl)), _p(type), _p(flags), _p(tl), C.c_int32(len(tk)), _p(tk), C.c_int32(ntask),
                                 C.c_int32(thistask), _p(lst), _p(togo), C.byref(ng))
    if nex < 0:
        raise ValueError("exchange_plan: leaf or task out of range")
    return lst[:nex].copy(), togo, int(ng.value)


def fof_primary(pos, ids, type, box, ll, mask=2):
    """fof_label_primary (fof.c:366-470): MinID of every particle for linking length ll over the types in mask"""
    pos = _c(pos, np.float64); ids = _c(ids, np.int64); type = _c(type, np.uint8)
    out = np.zeros(len(ids), np.int64)
    rc = lib().oracle_fof_primary(C.c_int64(len(ids)), _p(pos), _p(ids), _p(type), C.c_int(mask), C.c_double(box), C.c_double(ll), _p(out))
    if rc:
        raise MemoryError
    return out
