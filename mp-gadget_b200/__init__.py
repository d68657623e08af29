"""ctypes harness over libb200force.so (include/b200force.h).

This is the Python-side test/bench harness of the B200 force engine.  The
product is the C-ABI shared library built from csrc/ (see DESIGN.md); the
reference's host code is C, and the reference-signature shims live in
host/libgadget_shims.c.  Function names here follow the reference entry points
they drive (gravpm_init_periodic, gravpm_force, force_tree_full,
grav_short_tree: libgadget/gravity.h:40-58, forcetree.h:127).

There is NO CPU fallback: loading fails loudly if the library is missing and
Engine() fails if no CUDA device is present.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200_LIB", os.path.join(_HERE, "libb200force.so"))   # B200_LIB: experiment builds only
_lib = None

ALLMASK = 63   # libgadget/forcetree.h:22


class ParticleLayout(C.Structure):
    _fields_ = [("stride", C.c_int64), ("off_pos", C.c_int32), ("off_mass", C.c_int32),
                ("off_flags", C.c_int32), ("off_type", C.c_int32), ("off_vel", C.c_int32),
                ("off_fulltreeacc", C.c_int32), ("off_gravpm", C.c_int32), ("off_hsml", C.c_int32),
                ("off_potential", C.c_int32), ("off_pi", C.c_int32), ("off_timebin_hydro", C.c_int32),
                ("off_timebin_gravity", C.c_int32)]


class TreeInfo(C.Structure):
    _fields_ = [("numnodes", C.c_int64), ("numparticles", C.c_int64), ("maxdepth", C.c_int32),
                ("overfull_leaves", C.c_int32), ("root_mass", C.c_double)]


class GravShortParams(C.Structure):
    """struct gravshort_tree_params (libgadget/gravity.h:9-22) + GravShortPriv scalars."""
    _fields_ = [("ErrTolForceAcc", C.c_double), ("BHOpeningAngle", C.c_double),
                ("MaxBHOpeningAngle", C.c_double), ("TreeUseBH", C.c_int32), ("pad_", C.c_int32),
                ("Rcut", C.c_double), ("GravitySoftening", C.c_double), ("rho0", C.c_double)]


class SphParams(C.Structure):
    """struct density_params (density.h:10-28) + struct hydro_params (hydra.c:26-35) + time factors."""
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


class Timings(C.Structure):
    _fields_ = [(k, C.c_double) for k in (
        "pm_deposit", "pm_fft_forward", "pm_transfer", "pm_fft_inverse", "pm_gradient", "pm_readout", "pm_total",
        "tree_keys", "tree_sort", "tree_nodes", "tree_moments", "tree_total", "walk", "walk_post", "h2d", "d2h",
        "sph_density", "sph_hydro", "walk_pieces", "walk_list_bytes")]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class ShardedInfo(C.Structure):
    """b200_sharded_info"""
    _fields_ = [(k, C.c_int64) for k in ("n_own", "n_from_left", "n_from_right", "n_to_left", "n_to_right")] + \
               [(k, C.c_double) for k in ("ms_ghost", "ms_pm_total", "ms_pm_deposit", "ms_pm_halo_add", "ms_pm_fft2d", "ms_pm_pack",
                                          "ms_pm_a2a_forward", "ms_pm_fft1d_transfer", "ms_pm_a2a_backward", "ms_pm_unpack",
                                          "ms_pm_ifft2d", "ms_pm_halo_fill", "ms_pm_readout", "ms_top_allreduce")]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def comm_unique_id():
    """ncclGetUniqueId through the engine library (rank 0 calls it, the host broadcasts the 128 bytes)."""
    buf = C.create_string_buffer(128)
    if lib().b200_comm_unique_id(buf) != 0:
        raise B200Error("b200_comm_unique_id failed (NCCL not available in this process)")
    return buf.raw


COUNTS_DTYPE = np.dtype([("nodes_accepted", "i4"), ("nodes_opened", "i4"),
                         ("nodes_discarded", "i4"), ("particles", "i4")])

# struct particle_data of the reference (libgadget/partmanager.h:9-71), 160 bytes.
PARTICLE_DTYPE = np.dtype({
    "names": ["Pos", "TopLeaf", "Mass", "PI", "flags", "TimeBinHydro", "TimeBinGravity", "Type",
              "Vel", "FullTreeGravAccel", "GravPM", "Ti_drift", "Hsml", "DtHsml", "ID", "GrNr", "Potential"],
    "formats": [("f8", 3), "i4", "f4", "i4", "u1", "u1", "u1", "u1",
                ("f8", 3), ("f8", 3), ("f8", 3), "i8", "f8", "f8", "u8", "i8", "f8"],
    "offsets": [0, 24, 28, 32, 36, 37, 38, 39, 40, 64, 88, 112, 120, 128, 136, 144, 152],
    "itemsize": 160})

EXPORTED = [
    "b200_default_particle_layout", "b200_ctx_create", "b200_ctx_destroy", "b200_last_error",
    "b200_abi_version", "b200_kernel_launches", "b200_set_particles_aos", "b200_set_particles_soa",
    "b200_set_particles_soa_dev", "b200_oldacc_from_last_step", "b200_pm_init", "b200_walk_set_mesh", "b200_pm_force",
    "b200_pm_force_dev", "b200_pm_c2r_readout", "b200_pm_transform_kind", "b200_pm_set_power", "b200_pm_get_power", "b200_pm_cell_index", "b200_pm_copy_mesh", "b200_tree_build", "b200_tree_free",
    "b200_tree_export", "b200_grav_short_tree", "b200_grav_short_tree_dev", "b200_force_step_aos", "b200_force_step_aos_bytes", "b200_force_step_dev",
    "b200_get_timings", "b200_stream",
    "b200_tree_top_get_dev", "b200_tree_top_set_dev",
    "b200_comm_unique_id", "b200_comm_init", "b200_sharded_init", "b200_sharded_force_step",
    "b200_sph_set_gas", "b200_sph_set_timebins", "b200_sph_set_active", "b200_sph_set_hsml_range", "b200_sph_set_state", "b200_density", "b200_density_gradrho", "b200_hydro_force",
    "b200_step_set_state", "b200_step_get_state", "b200_step_adopt_forces", "b200_step_drift", "b200_step_build_active",
    "b200_step_active_sublist", "b200_step_get_active", "b200_step_half_kick", "b200_step_pm_kick",
    "b200_step_hier_accelerations", "b200_step_hier_timesteps", "b200_step_hydro_timesteps", "b200_step_find_timesteps", "b200_step_grav_short_tree", "b200_step_set_active", "b200_step_get_store", "b200_step_set_store", "b200_step_sph_prepare", "b200_step_adopt_hydro",
    "b200_domain_peano_keys", "b200_domain_set_topnodes", "b200_domain_topleaf", "b200_domain_leaf_counts", "b200_domain_assign_balanced", "b200_domain_exchange_plan",
    "b200_domain_sample_keys", "b200_domain_toptree_local", "b200_domain_toptree_truncate", "b200_domain_toptree_merge",
    "b200_domain_toptree_global_refine", "b200_domain_toptree_leaves",
    "b200_fof_primary",
]


def build(verbose=False):
    """Compile csrc/ for sm_100a into libb200force.so (in-tree)."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-C", os.path.join(_HERE, "csrc"), "-j4"], stdout=out)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libb200force.so is not built (run __graft_entry__.build()); "
                               "there is no fallback path")
        L = C.CDLL(LIB_PATH)
        L.b200_last_error.restype = C.c_char_p
        L.b200_kernel_launches.restype = C.c_int64
        L.b200_stream.restype = C.c_void_p
        L.b200_ctx_destroy.restype = None
        L.b200_tree_free.restype = None
        L.b200_default_particle_layout.restype = None
        _lib = L
    return _lib


def _is_dev(a):
    return hasattr(a, "data_ptr")          # a torch tensor (device-resident callers: sharded.py)


def _p(a):
    if a is None:
        return None
    return C.c_void_p(a.data_ptr()) if _is_dev(a) else C.c_void_p(a.ctypes.data)


def _c(a, dt):
    if a is None:
        return None
    if _is_dev(a):
        assert str(a.dtype).endswith(np.dtype(dt).name) and a.is_contiguous(), "device arrays must come as contiguous %s" % np.dtype(dt).name
        return a
    return np.ascontiguousarray(a, dtype=dt)


class B200Error(RuntimeError):
    pass


class Engine:
    """One engine = one GPU (b200_ctx)."""

    def __init__(self, device=0):
        self.L = lib()
        self.ctx = C.c_void_p()
        rc = self.L.b200_ctx_create(C.byref(self.ctx), C.c_int(device))
        if rc != 0:
            self.ctx = None
            raise B200Error("b200_ctx_create failed (rc=%d): no CUDA device / no fallback" % rc)
        self.n = 0
        self.nmesh = 0

    def close(self):
        if self.ctx is not None:
            self.L.b200_ctx_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise B200Error(self.L.b200_last_error(self.ctx).decode())

    # -- particles ---------------------------------------------------------
    def set_particles(self, pos, mass, type=None, oldacc=None):
        pos = _c(pos, np.float64); mass = _c(mass, np.float32)
        type = _c(type, np.uint8); oldacc = _c(oldacc, np.float64)
        self.n = len(mass)
        self._keep = (pos, mass, type, oldacc)
        self._ck(self.L.b200_set_particles_soa(self.ctx, _p(pos), _p(mass), _p(type), _p(oldacc), C.c_int64(self.n)))

    def set_particles_dev(self, pos_ptr, mass_ptr, n, type_ptr=None, oldacc_ptr=None):
        self.n = int(n)
        self._ck(self.L.b200_set_particles_soa_dev(self.ctx, C.c_void_p(pos_ptr), C.c_void_p(mass_ptr),
                                                   C.c_void_p(type_ptr) if type_ptr else None,
                                                   C.c_void_p(oldacc_ptr) if oldacc_ptr else None, C.c_int64(self.n)))

    def set_particles_aos(self, P):
        assert P.dtype.itemsize == 160
        self.n = len(P)
        self._ck(self.L.b200_set_particles_aos(self.ctx, _p(P), C.c_int64(self.n), None))

    def oldacc_from_last_step(self):
        self._ck(self.L.b200_oldacc_from_last_step(self.ctx))

    # -- PM: gravpm_init_periodic / gravpm_force -----------------------------
    def gravpm_init_periodic(self, BoxSize, Asmth, Nmesh, G):
        self.nmesh = int(Nmesh)
        self._ck(self.L.b200_pm_init(self.ctx, C.c_double(BoxSize), C.c_double(Asmth), C.c_int(Nmesh), C.c_double(G)))

    def walk_set_mesh(self, BoxSize, Asmth, Nmesh, G):
        """The scalars grav_short_tree takes from the PetaPM struct (cell size, Asmth, G) without allocating a mesh."""
        self.nmesh = int(Nmesh)
        self._ck(self.L.b200_walk_set_mesh(self.ctx, C.c_double(BoxSize), C.c_double(Asmth), C.c_int(Nmesh), C.c_double(G)))

    def gravpm_force(self, want_potential=True):
        g = np.empty((self.n, 3))
        p = np.empty(self.n) if want_potential else None
        self._ck(self.L.b200_pm_force(self.ctx, _p(g), _p(p)))
        return g, p

    def gravpm_force_dev(self, gravpm_ptr=None, pot_ptr=None):
        self._ck(self.L.b200_pm_force_dev(self.ctx, C.c_void_p(gravpm_ptr) if gravpm_ptr else None,
                                          C.c_void_p(pot_ptr) if pot_ptr else None))

    def fof_primary(self, ids, box, ll, mask=2):
        """fof_label_primary (fof.c:366-470) for the particles set: (MinID of every particle, number of groups)."""
        ids = _c(ids, np.int64)
        assert len(ids) == self.n
        out = np.empty(self.n, np.int64)
        ng = C.c_int64(0)
        self._ck(self.L.b200_fof_primary(self.ctx, _p(ids), C.c_int(mask), C.c_double(box), C.c_double(ll), _p(out), C.byref(ng)))
        return out, int(ng.value)

    def pm_c2r_readout(self, rho_k, functions):
        """petapm_force_c2r (petapm.c:326-362) with a caller's spectrum: rho_k complex [Nmesh, Nmesh, Nmesh/2+1]; functions =
        [(kind, table[k2]), ...] -> one array of n read-outs per function."""
        N = self.nmesh
        rk = _c(np.asarray(rho_k, np.complex128), np.complex128)
        assert rk.shape == (N, N, N // 2 + 1)

        class F(C.Structure):
            _fields_ = [("kind", C.c_int), ("table", C.c_void_p), ("out", C.c_void_p)]
        nk2 = 3 * (N // 2) ** 2 + 1
        tabs = [_c(t, np.float64) for _, t in functions]
        assert all(len(t) == nk2 for t in tabs)
        outs = [np.zeros(max(self.n, 1)) for _ in functions]
        arr = (F * max(len(functions), 1))()
        for j, (kind, _) in enumerate(functions):
            arr[j].kind = int(kind); arr[j].table = tabs[j].ctypes.data; arr[j].out = outs[j].ctypes.data
        self._ck(self.L.b200_pm_c2r_readout(self.ctx, C.c_void_p(rk.ctypes.data), C.c_int(len(functions)), arr))
        return [o[:self.n] for o in outs]

    def pm_transform_kind(self):
        """1: the engine's own shared-memory transform passes (csrc/pm_fft.cu); 0: cuFFT; -1: no mesh."""
        return int(self.L.b200_pm_transform_kind(self.ctx))

    def pm_power(self):
        """Raw power-spectrum sums of the last gravpm_force (after pm_set_power(True))."""
        nb = self.nmesh
        pw, kk, nm, norm = np.zeros(nb), np.zeros(nb), np.zeros(nb, np.int64), C.c_double()
        self._ck(self.L.b200_pm_get_power(self.ctx, C.c_int(nb), _p(pw), _p(kk), _p(nm), C.byref(norm)))
        return pw, kk, nm, norm.value

    def pm_set_power(self, on=True):
        self._ck(self.L.b200_pm_set_power(self.ctx, C.c_int(1 if on else 0)))

    def pm_cell_index(self):
        ic = np.empty((self.n, 3), dtype=np.int32)
        self._ck(self.L.b200_pm_cell_index(self.ctx, _p(ic)))
        return ic

    def pm_copy_mesh(self, which):
        m = np.empty((self.nmesh,) * 3)
        self._ck(self.L.b200_pm_copy_mesh(self.ctx, C.c_int(which), _p(m)))
        return m

    # -- tree: force_tree_full / force_tree_active_moments --------------------
    def force_tree_build(self, BoxSize, mask=ALLMASK, active=None, toplevel_depth=0):
        info = TreeInfo()
        active = _c(active, np.int32)
        na = 0 if active is None else len(active)
        self._ck(self.L.b200_tree_build(self.ctx, C.c_double(BoxSize), C.c_int(mask), _p(active), C.c_int64(na),
                                        C.c_int(toplevel_depth), C.byref(info)))
        self.tree_info = info
        return info

    def force_tree_full(self, BoxSize):
        return self.force_tree_build(BoxSize, ALLMASK, None, 0)

    def tree_export(self):
        nn = self.tree_info.numnodes
        out = dict(center=np.empty((nn, 3)), len=np.empty(nn), cofm=np.empty((nn, 3)), mass=np.empty(nn),
                   hmax=np.empty(nn), sibling=np.empty(nn, np.int32), firstchild=np.empty(nn, np.int32),
                   nocc=np.empty(nn, np.int32), part=np.empty((nn, 8), np.int32))
        self._ck(self.L.b200_tree_export(self.ctx, _p(out["center"]), _p(out["len"]), _p(out["cofm"]), _p(out["mass"]),
                                         _p(out["hmax"]), _p(out["sibling"]), _p(out["firstchild"]), _p(out["nocc"]),
                                         _p(out["part"])))
        return out

    # -- grav_short_tree -------------------------------------------------------
    def grav_short_tree(self, par, active=None, want_counts=False, want_potential=True):
        p = GravShortParams(**par) if isinstance(par, dict) else par
        acc = np.zeros((self.n, 3))
        pot = np.zeros(self.n) if want_potential else None
        cnt = np.zeros(self.n, dtype=COUNTS_DTYPE) if want_counts else None
        active = _c(active, np.int32)
        na = 0 if active is None else len(active)
        self._ck(self.L.b200_grav_short_tree(self.ctx, C.byref(p), _p(active), C.c_int64(na), _p(acc), _p(pot), _p(cnt)))
        return acc, pot, cnt

    def grav_short_tree_dev(self, par, acc_ptr=None, pot_ptr=None, active_ptr=None, nactive=0):
        p = GravShortParams(**par) if isinstance(par, dict) else par
        self._ck(self.L.b200_grav_short_tree_dev(self.ctx, C.byref(p), C.c_void_p(active_ptr) if active_ptr else None, C.c_int64(nactive),
                                                 C.c_void_p(acc_ptr) if acc_ptr else None,
                                                 C.c_void_p(pot_ptr) if pot_ptr else None, None))

    # -- SPH: density / hydro_force ------------------------------------------------
    def sph_set_gas(self, hsml, vel=None, entropy=None, dtentropy=None, fullacc=None, gravpm=None, hydroacc=None):
        f8 = lambda a: _c(a, np.float64)
        arrs = [f8(vel), f8(hsml), f8(entropy), f8(dtentropy), f8(fullacc), f8(gravpm), f8(hydroacc)]
        self._keep_sph = arrs
        self._ck(self.L.b200_sph_set_gas(self.ctx, *[_p(a) for a in arrs]))

    def sph_set_timebins(self, bin_gravity, bin_hydro, tables):
        """tables: dict of per-bin arrays gravkick, hydrokick, dloga_pred, drift, dloga_bin (<= 47 entries)."""
        t = np.zeros((5, 47))
        for r, k in enumerate(("gravkick", "hydrokick", "dloga_pred", "drift", "dloga_bin")):
            v = np.asarray(tables[k], dtype=np.float64)[:47]
            t[r, :len(v)] = v
        bg = _c(bin_gravity, np.uint8); bh = _c(bin_hydro, np.uint8)
        self._ck(self.L.b200_sph_set_timebins(self.ctx, _p(bg), _p(bh), _p(t)))

    def sph_set_active(self, active):
        a = _c(active, np.int32)
        self._ck(self.L.b200_sph_set_active(self.ctx, _p(a), C.c_int64(0 if a is None else len(a))))

    def sph_set_hsml_range(self, hsml, first):
        h = _c(hsml, np.float64)
        self._ck(self.L.b200_sph_set_hsml_range(self.ctx, _p(h), C.c_int64(first), C.c_int64(len(h))))

    def sph_set_state(self, density=None, egywtdensity=None, dhsmlfac=None, divvel=None, curlvel=None):
        arr = [_c(x, np.float64) for x in (density, egywtdensity, dhsmlfac, divvel, curlvel)]
        self._ck(self.L.b200_sph_set_state(self.ctx, *[_p(x) for x in arr]))

    def _zeros(self, shape, dt, device):
        if device is None:
            return np.zeros(shape, dt)
        import torch
        return torch.zeros(shape, dtype={np.float64: torch.float64, np.int32: torch.int32}[dt], device=device)

    def density(self, sp, update_hsml=1, DoEgyDensity=0, device=None):
        """device: a torch device -> the outputs are tensors in HBM (no host copy)"""
        n = self.n
        z = lambda dt: self._zeros(n, dt, device)
        out = dict(hsml=z(np.float64), density=z(np.float64), egywtdensity=z(np.float64), dhsmlfac=z(np.float64), divvel=z(np.float64),
                   curlvel=z(np.float64), dthsml=z(np.float64), numngb=z(np.float64), ninteract=z(np.int32), niter=z(np.int32))
        self._ck(self.L.b200_density(self.ctx, C.byref(sp), C.c_int(update_hsml), C.c_int(DoEgyDensity),
                                     *[_p(out[k]) for k in ("hsml", "density", "egywtdensity", "dhsmlfac", "divvel", "curlvel",
                                                            "dthsml", "numngb", "ninteract", "niter")]))
        return out

    def hydro_force(self, sp, device=None):
        n = self.n
        out = dict(acc=self._zeros((n, 3), np.float64, device), dtentropy=self._zeros(n, np.float64, device),
                   maxsignalvel=self._zeros(n, np.float64, device), ninteract=self._zeros(n, np.int32, device))
        self._ck(self.L.b200_hydro_force(self.ctx, C.byref(sp), _p(out["acc"]), _p(out["dtentropy"]), _p(out["maxsignalvel"]),
                                         _p(out["ninteract"])))
        return out

    # -- multi-GPU building blocks ------------------------------------------------
    def tree_top_get_dev(self, level, ptr):
        self._ck(self.L.b200_tree_top_get_dev(self.ctx, C.c_int(level), C.c_void_p(ptr)))

    def tree_top_set_dev(self, level, ptr):
        self._ck(self.L.b200_tree_top_set_dev(self.ctx, C.c_int(level), C.c_void_p(ptr)))

    # -- the sharded TreePM force step (sharded.cu; NCCL issued from C) ---------------
    def comm_init(self, rank, world, id_tree=None, id_pm=None):
        """id_*: 128-byte buffers from comm_unique_id() of rank 0, broadcast by the host."""
        a = (C.c_char * 128).from_buffer_copy(bytes(id_tree)) if id_tree is not None else None
        b = (C.c_char * 128).from_buffer_copy(bytes(id_pm)) if id_pm is not None else None
        self._ck(self.L.b200_comm_init(self.ctx, C.c_int(rank), C.c_int(world), a, b))

    def sharded_init(self, BoxSize, Asmth, Nmesh, G, topdepth, halo=6, rcut_cells=0.0):
        self._ck(self.L.b200_sharded_init(self.ctx, C.c_double(BoxSize), C.c_double(Asmth), C.c_int(Nmesh), C.c_double(G),
                                          C.c_int(topdepth), C.c_int(halo), C.c_double(rcut_cells)))
        self.nmesh = int(Nmesh)

    def sharded_force_step(self, par, pos_ptr, mass_ptr, oldacc_ptr, n_own, gravpm_ptr, acc_ptr, pot_ptr):
        p = GravShortParams(**par) if isinstance(par, dict) else par
        info = ShardedInfo()
        v = lambda x: C.c_void_p(x) if x else None
        self._ck(self.L.b200_sharded_force_step(self.ctx, v(pos_ptr), v(mass_ptr), v(oldacc_ptr), C.c_int64(n_own), C.byref(p),
                                                v(gravpm_ptr), v(acc_ptr), v(pot_ptr), C.byref(info)))
        return info

    # -- whole step on the reference's AoS -------------------------------------
    def force_step_aos_bytes(self):
        """(host->device, device->host) bytes per particle force_step_aos moves for the default record layout"""
        a, b = C.c_int64(), C.c_int64()
        self.L.b200_force_step_aos_bytes(None, C.byref(a), C.byref(b))
        return int(a.value), int(b.value)

    def force_step_aos(self, P, par, ptr=None, n=None):
        p = GravShortParams(**par) if isinstance(par, dict) else par
        if ptr is None:
            ptr, n = P.ctypes.data, len(P)
        self.n = int(n)
        self._ck(self.L.b200_force_step_aos(self.ctx, C.c_void_p(ptr), C.c_int64(n), None, C.byref(p)))

    def force_step_dev(self, par, gravpm_ptr=None, acc_ptr=None, pot_ptr=None, pmpot_ptr=None):
        p = GravShortParams(**par) if isinstance(par, dict) else par
        v = lambda x: C.c_void_p(x) if x else None
        self._ck(self.L.b200_force_step_dev(self.ctx, C.byref(p), v(gravpm_ptr), v(pmpot_ptr), v(acc_ptr), v(pot_ptr)))

    def timings(self):
        t = Timings()
        self.L.b200_get_timings(self.ctx, C.byref(t))
        return t.asdict()

    # -- domain keys ---------------------------------------------------------
    def peano_keys(self, box):
        """PEANO(Pos, BoxSize) of every particle (utils/peano.h:15-21) -> uint64[n]"""
        keys = np.zeros(max(self.n, 1), np.uint64)
        self._ck(self.L.b200_domain_peano_keys(self.ctx, C.c_double(box), _p(keys)))
        return keys[:self.n]

    def domain_set_topnodes(self, top):
        """DomainDecomp::TopNodes as (Daughter, StartKey, Shift, Leaf): the forced top tree of force_tree_build(toplevel_depth=-1)"""
        a = [_c(top[0], np.int32), _c(top[1], np.uint64), _c(top[2], np.int32), _c(top[3], np.int32)]
        self._ck(self.L.b200_domain_set_topnodes(self.ctx, C.c_int32(len(a[0])), _p(a[0]), _p(a[1]), _p(a[2]), _p(a[3])))

    def topleaf(self, daughter, startkey, shift, leaf):
        """domain_get_topleaf (domain.h:71-78) of every particle over TopNodes given as arrays (after peano_keys)"""
        a = [_c(daughter, np.int32), _c(startkey, np.uint64), _c(shift, np.int32), _c(leaf, np.int32)]
        self._ck(self.L.b200_domain_set_topnodes(self.ctx, C.c_int32(len(a[0])), _p(a[0]), _p(a[1]), _p(a[2]), _p(a[3])))
        out = np.zeros(max(self.n, 1), np.int32)
        self._ck(self.L.b200_domain_topleaf(self.ctx, _p(out)))
        return out[:self.n]

    def sample_keys(self, box, subsample):
        """keys of every subsample-th particle (domain.c:1066-1074) -> uint64[n // subsample]"""
        keys = np.zeros(max(self.n // subsample, 1), np.uint64); ns = C.c_int64()
        self._ck(self.L.b200_domain_sample_keys(self.ctx, C.c_double(box), C.c_int32(subsample), _p(keys), C.byref(ns)))
        return keys[:ns.value]

    def exchange_plan(self, task_of_leaf, ntask, thistask):
        """exchange.c:408-444,505-530 after topleaf() -> (exchange list, togo[ntask][7], ngarbage)"""
        tk = _c(task_of_leaf, np.int32)
        lst = np.zeros(self.n + 1, np.int32); togo = np.zeros((ntask, 7), np.int64); nex = C.c_int64(); ng = C.c_int64()
        self._ck(self.L.b200_domain_exchange_plan(self.ctx, _p(tk), C.c_int32(len(tk)), C.c_int32(ntask), C.c_int32(thistask),
                                                  C.byref(nex), C.byref(ng), _p(togo), _p(lst)))
        return lst[:nex.value].copy(), togo, int(ng.value)

    def leaf_counts(self, nleaf):
        """TopLeafCount (domain.c:1396-1451) of the last topleaf() -> int64[nleaf]"""
        out = np.zeros(nleaf, np.int64)
        self._ck(self.L.b200_domain_leaf_counts(self.ctx, C.c_int32(nleaf), _p(out)))
        return out

    def kernel_launches(self):
        return int(self.L.b200_kernel_launches(self.ctx))

    def stream(self):
        return self.L.b200_stream(self.ctx)


def domain_assign_balanced(ntask, cost, nseg_per_task=1):
    """domain_assign_topleaves_balanced (domain.c:610-755) over leaves in key order -> task per leaf"""
    cost = _c(cost, np.int64); task = np.zeros(len(cost), np.int32)
    if lib().b200_domain_assign_balanced(C.c_int32(ntask), C.c_int32(len(cost)), _p(cost), C.c_int32(nseg_per_task), _p(task)) != 0:
        raise B200Error("b200_domain_assign_balanced: the leaves cannot be dealt out to %d tasks" % ntask)
    return task


TOPNODE_DTYPE = np.dtype([("StartKey", "u8"), ("Shift", "i4"), ("Daughter", "i4"), ("Parent", "i4"), ("pad_", "i4"), ("Count", "i8"), ("Cost", "i8")])


class TopTree:
    """b200_domain_toptree_* (host side of the domain decomposition, domain.c:826-1395); .tree is the node array."""

    def __init__(self, maxnodes):
        self.nodes = np.zeros(maxnodes, TOPNODE_DTYPE); self.size = C.c_int32(0); self.maxnodes = maxnodes

    @property
    def tree(self):
        return self.nodes[:self.size.value]

    def local(self, sample_keys):
        k = np.array(sample_keys, np.uint64, copy=True)
        return lib().b200_domain_toptree_local(_p(k), C.c_int64(len(k)), _p(self.nodes), C.byref(self.size), C.c_int32(self.maxnodes))

    def truncate(self, countlimit, costlimit):
        return lib().b200_domain_toptree_truncate(_p(self.nodes), C.byref(self.size), C.c_int64(countlimit), C.c_int64(costlimit))

    def merge(self, other):
        return lib().b200_domain_toptree_merge(_p(self.nodes), C.byref(self.size), _p(other.nodes), C.c_int32(self.maxnodes))

    def global_refine(self, countlimit, costlimit):
        return lib().b200_domain_toptree_global_refine(_p(self.nodes), C.byref(self.size), C.c_int32(self.maxnodes), C.c_int64(countlimit), C.c_int64(costlimit))

    def leaves(self):
        leaf = np.zeros(self.size.value, np.int32); nl = C.c_int32()
        lib().b200_domain_toptree_leaves(_p(self.nodes), self.size, _p(leaf), C.byref(nl))
        return nl.value, leaf
