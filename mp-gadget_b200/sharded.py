"""Domain-sharded TreePM force step: one process + one engine per GPU.

Host-side orchestration of the multi-GPU path.  The
compute is the C-ABI engine; torch.distributed (NCCL on GPUs, gloo in the CPU
tests) is the plumbing.  What the reference does with MPI:

  * domain cut -> rank owns a contiguous set of top-level tree cells
    (domain.c:154-256, domain.h:71-78).  Here: the cells of a uniform top tree
    of depth `topdepth` on the reference's own lattice (root 1.001*Box,
    peano.h:15-21), split into x-layers; rank r owns layers
    [r*2^d/W, (r+1)*2^d/W).
  * short-range tree: the reference exports *queries* to the ranks whose
    top-leaves must be opened (treewalk.c:325-371,399-793).  Here (SURVEY 8e
    option 2, "ghost import"): each rank imports the particles of the cell
    layer adjacent to its domain from both neighbours, builds the same forced
    top tree + complete subtrees for own and imported cells, and walks its own
    particles only.  A cell that was not imported lies more than one cell width
    (> Rcut) away from every own particle and is discarded by
    shall_we_discard_node (gravshort-tree.c:198-215) whatever its content.
    Top-cell moments are summed over ranks and the upper levels re-summed
    (force_exchange_pseudodata + force_treeupdate_pseudos,
    forcetree.c:1156-1284), so every node a target can see carries the same
    moments as in a single-rank run with the same top tree.
  * PM: x-slab decomposition of the mesh (the reference uses 2-D pencils,
    petapm.c:127-150): halo planes of the deposit are added into the
    neighbours (replaces layout_build_and_exchange_cells_to_pfft,
    petapm.c:792-840), 2-D FFT per plane, all-to-all transpose, 1-D FFT along
    x, Green's function, and back; halo planes of the potential are fetched
    from the neighbours (replaces ..._to_local, petapm.c:848-885).

The same class runs with world size 1 (self-neighbours), which is how the
single-GPU tests check it against the unsharded engine path.
"""
import numpy as np
import torch


class _DevView:
    """Expose a raw device pointer to torch through __cuda_array_interface__."""

    def __init__(self, ptr, shape, typestr="<f8"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def dev_tensor(ptr, shape, device):
    return torch.as_tensor(_DevView(ptr, shape), device=device)


class Comm:
    """Neighbour / all-to-all exchanges over torch.distributed (or none for world 1)."""

    def __init__(self, dist=None):
        self.dist = dist
        self.rank = dist.get_rank() if dist is not None else 0
        self.world = dist.get_world_size() if dist is not None else 1
        self.left = (self.rank - 1) % self.world
        self.right = (self.rank + 1) % self.world

    def neighbour_exchange(self, send_left, send_right, recv_from_right, recv_from_left):
        """send_left goes to the left neighbour (who receives it as 'from its right')."""
        if self.world == 1:
            recv_from_right.copy_(send_left)
            recv_from_left.copy_(send_right)
            return
        d = self.dist
        ops = [d.P2POp(d.isend, send_left, self.left), d.P2POp(d.isend, send_right, self.right),
               d.P2POp(d.irecv, recv_from_right, self.right), d.P2POp(d.irecv, recv_from_left, self.left)]
        for r in d.batch_isend_irecv(ops):
            r.wait()

    def neighbour_exchange_var(self, send_left, send_right):
        """Variable-length [k, C] tensors; returns (from_right, from_left)."""
        if self.world == 1:
            return send_left.clone(), send_right.clone()
        dev, dt = send_left.device, send_left.dtype
        cl = torch.tensor([send_left.shape[0]], dtype=torch.int64, device=dev)
        cr = torch.tensor([send_right.shape[0]], dtype=torch.int64, device=dev)
        nr, nl = torch.zeros_like(cl), torch.zeros_like(cl)
        self.neighbour_exchange(cl, cr, nr, nl)
        cols = send_left.shape[1]
        fr = torch.empty((int(nr.item()), cols), dtype=dt, device=dev)
        fl = torch.empty((int(nl.item()), cols), dtype=dt, device=dev)
        self.neighbour_exchange(send_left.contiguous(), send_right.contiguous(), fr, fl)
        return fr, fl

    def all_to_all_equal(self, send, recv):
        """send/recv: [world, ...] contiguous blocks of equal size."""
        if self.world == 1:
            recv.copy_(send)
            return
        d = self.dist
        if d.get_backend() == "nccl":
            d.all_to_all_single(recv.view(-1), send.view(-1))
            return
        ops = []
        for p in range(self.world):        # gloo has no all_to_all: pairwise sends
            if p == self.rank:
                recv[p].copy_(send[p])
            else:
                ops.append(d.P2POp(d.isend, send[p], p))
                ops.append(d.P2POp(d.irecv, recv[p], p))
        if ops:
            for r in d.batch_isend_irecv(ops):
                r.wait()

    def all_reduce_sum(self, t):
        if self.world > 1:
            self.dist.all_reduce(t)
        return t


class Domain:
    """x-layers of the top-tree cells on the reference's Peano lattice."""

    def __init__(self, box, topdepth, rank, world):
        self.box, self.d, self.rank, self.world = float(box), int(topdepth), rank, world
        self.ncell = 1 << self.d
        if self.ncell % world:
            raise ValueError("2^topdepth must be a multiple of the number of ranks")
        self.per = self.ncell // world
        self.lo, self.hi = rank * self.per, (rank + 1) * self.per
        self.cellwidth = 1.001 * self.box / self.ncell
        # width over which the last layer holds particles across the periodic seam (the root cell is 1.001 Box wide):
        # this, not cellwidth, has to exceed Rcut / the largest smoothing length
        self.seamwidth = self.cellwidth - 0.001 * self.box
        self.domainfac = 1.0 / (self.box * 1.001) * float(1 << 21)       # PEANO(), utils/peano.h:15-21

    def layer_of(self, x):
        """Top-cell x-index of positions x (torch f64)."""
        ix = ((x + self.box / 2000) * self.domainfac).to(torch.int64)
        return ix >> (21 - self.d)

    def owner_of(self, x):
        return torch.div(self.layer_of(x), self.per, rounding_mode="floor")

    def own_cell_mask(self, device):
        """[8^d] bool: Morton-indexed top cells owned by this rank (x bit = bit 0 of each octal digit)."""
        m = torch.arange(8 ** self.d, dtype=torch.int64, device=device)
        ix = torch.zeros_like(m)
        for level in range(self.d):            # digit `level` counted from the least significant
            ix |= ((m >> (3 * level)) & 1) << level
        return (ix >= self.lo) & (ix < self.hi)

    def ghost_sets(self, x):
        """Boolean masks of own particles to send to the left / right neighbour."""
        lay = self.layer_of(x)
        to_left = lay == self.lo
        to_right = lay == self.hi - 1
        if self.world == 2:
            # both neighbours are the same rank: it must receive each particle once
            both = to_left | to_right
            return both, torch.zeros_like(both)
        if self.world == 1:
            z = torch.zeros_like(to_left)
            return z, z
        return to_left, to_right


def slab_transpose_forward(comm, cplx, cplxT, sendbuf, recvbuf, sync=None):
    """[ix(local)][iy][kz] -> [jy(local)][ix][kz]: the all-to-all that PFFT does
    between its 1-D stages (petapm.c:305, PFFT_TRANSPOSED_OUT)."""
    W = comm.world
    nx, nz = cplx.shape[0], cplx.shape[2]
    if W == 1:
        cplxT.copy_(cplx.permute(1, 0, 2, 3))
        return
    sendbuf.copy_(cplx.view(nx, W, nx, nz, 2).permute(1, 0, 2, 3, 4))      # [dest][ix][jy][kz]
    if sync:
        sync()
    comm.all_to_all_equal(sendbuf, recvbuf)                                  # [src][ix][jy][kz]
    cplxT.view(nx, W, nx, nz, 2).copy_(recvbuf.permute(2, 0, 1, 3, 4))      # [jy][src][ix][kz]


def slab_transpose_backward(comm, cplx, cplxT, sendbuf, recvbuf, sync=None):
    W = comm.world
    nx, nz = cplx.shape[0], cplx.shape[2]
    if W == 1:
        cplx.copy_(cplxT.permute(1, 0, 2, 3))
        return
    sendbuf.copy_(cplxT.view(nx, W, nx, nz, 2).permute(1, 2, 0, 3, 4))     # [dest][ix][jy][kz]
    if sync:
        sync()
    comm.all_to_all_equal(sendbuf, recvbuf)                                  # [src][ix][jy(src)][kz]
    cplx.view(nx, W, nx, nz, 2).copy_(recvbuf.permute(1, 0, 2, 3, 4))       # [ix][src][jy][kz]


def halo_add(comm, real, nx, h, buf_a, buf_b):
    """Add the deposit that landed in my halo planes into the neighbours' edge planes."""
    comm.neighbour_exchange(real[0:h].contiguous(), real[h + nx:].contiguous(), buf_a, buf_b)
    real[nx:nx + h] += buf_a        # from my right neighbour: its left halo = my last owned planes
    real[h:2 * h] += buf_b          # from my left neighbour: its right halo = my first owned planes


def halo_fill(comm, real, nx, h, buf_a, buf_b):
    """Fetch the neighbours' edge planes of the potential into my halo planes."""
    comm.neighbour_exchange(real[h:2 * h].contiguous(), real[nx:nx + h].contiguous(), buf_a, buf_b)
    real[h + nx:] = buf_a           # my right halo = right neighbour's first owned planes
    real[0:h] = buf_b               # my left halo = left neighbour's last owned planes


class ShardedTreePM:
    """One rank of the sharded TreePM force step.  The step itself is b200_sharded_force_step (csrc/sharded.cu): ghost
    import, slab PM with its halo exchanges and all-to-all transposes, tree build, top-moment all-reduce and walk are
    issued from C on the engine's streams with NCCL.  This class only bootstraps the communicators (the two NCCL ids
    travel over torch.distributed, as they would over MPI_Bcast in an MP-Gadget host) and owns the output tensors.
    world size 1 needs no NCCL (self-neighbour copies): that is how the single-GPU tests compare it with the unsharded path."""

    def __init__(self, engine, box, nmesh, asmth, G, topdepth, dist=None, halo=6, device="cuda", rcut_cells=None):
        import importlib
        pkg = importlib.import_module(__package__)
        self.e = engine
        self.dist = dist
        self.rank = dist.get_rank() if dist is not None else 0
        self.world = dist.get_world_size() if dist is not None else 1
        self.box, self.nmesh, self.asmth, self.G = float(box), int(nmesh), float(asmth), float(G)
        self.dom = Domain(box, topdepth, self.rank, self.world)
        self.device = torch.device(device)
        self.topdepth = int(topdepth)
        if self.world > 1:
            if rcut_cells is None:
                raise ValueError("rcut_cells (TreeRcut) is required: the ghost layer must be validated against the cut-off radius")
            ids = torch.zeros(256, dtype=torch.uint8, device=self.device)
            if self.rank == 0:
                raw = pkg.comm_unique_id() + pkg.comm_unique_id()
                ids.copy_(torch.frombuffer(bytearray(raw), dtype=torch.uint8))
            dist.broadcast(ids, 0)
            raw = bytes(ids.cpu().numpy().tobytes())
            engine.comm_init(self.rank, self.world, raw[:128], raw[128:])
        else:
            engine.comm_init(0, 1)
        engine.sharded_init(box, asmth, nmesh, G, topdepth, halo=halo, rcut_cells=float(rcut_cells or 0.0))
        self.n_own = self.n_tot = 0
        self._out_n = -1
        self.info = None
        self.last_oldacc = None

    def _outputs(self, n):
        if n != self._out_n:
            self.gpm = torch.empty((n, 3), dtype=torch.float64, device=self.device)
            self.acc = torch.empty((n, 3), dtype=torch.float64, device=self.device)
            self.pot = torch.empty(n, dtype=torch.float64, device=self.device)
            self._out_n = n

    def force_step(self, pos, mass, oldacc, par):
        """pos [n,3] f64, mass [n] f32, oldacc [n,3] f64 or None: device tensors of the rank's OWN particles (every x
        inside the rank's layers), already visible to the engine's stream.  Returns (GravPM, FullTreeGravAccel,
        Potential) of those particles: tensors owned by this object, overwritten by the next call."""
        n = int(pos.shape[0])
        self._outputs(n)
        info = self.e.sharded_force_step(par, pos.data_ptr(), mass.data_ptr(), oldacc.data_ptr() if oldacc is not None else None, n,
                                         self.gpm.data_ptr(), self.acc.data_ptr(), self.pot.data_ptr())
        self.info = info
        self.last_oldacc = oldacc
        self.n_own = n
        self.n_tot = n + int(info.n_from_left) + int(info.n_from_right)
        return self.gpm, self.acc, self.pot

    def phase_ms(self):
        return {} if self.info is None else {k: v for k, v in self.info.asdict().items() if k.startswith("ms_")}

    def parity_check(self, pos, mass, acc, gpm, par, nsample=4096, check_pm=None):
        """The sharded result (acc, gpm of the last force_step on pos, mass) of sampled own particles against ONE engine
        holding the whole box: every rank gathers all particles, builds the same forced top tree, walks its sample with
        the old accelerations the sharded step used, and (while the unsharded mesh fits beside this rank's state) runs
        the unsharded PM.  Returns the errors relative to the mean |acceleration| of the sample."""
        import importlib
        pkg = importlib.import_module(__package__)
        dist, dev = self.dist, self.device
        n = int(pos.shape[0])
        counts = torch.zeros(self.world, dtype=torch.int64, device=dev)
        counts[self.rank] = n
        if dist is not None:
            dist.all_reduce(counts)
        counts = [int(c) for c in counts.cpu()]
        ntot = sum(counts)
        off = sum(counts[:self.rank])
        allpos = torch.zeros((ntot, 3), dtype=torch.float64, device=dev)
        allmass = torch.zeros(ntot, dtype=torch.float32, device=dev)
        allpos[off:off + n] = pos; allmass[off:off + n] = mass
        if dist is not None:
            dist.all_reduce(allpos); dist.all_reduce(allmass)
        old = None
        if self.last_oldacc is not None:
            old = torch.zeros((ntot, 3), dtype=torch.float64, device=dev)
            old[off:off + n] = self.last_oldacc
        g = torch.Generator(device="cpu"); g.manual_seed(1234 + self.rank)
        sample = torch.sort(torch.randperm(n, generator=g)[:min(nsample, n)])[0].to(dev)
        targets = (sample + off).to(torch.int32).contiguous()
        if check_pm is None:
            check_pm = self.nmesh <= 1280             # the unsharded mesh + spectrum + cuFFT work area must fit beside this rank's state
        torch.cuda.synchronize()
        e2 = pkg.Engine(self.device.index or 0)
        out = {"nsample": int(sample.shape[0]), "pm_checked": bool(check_pm)}
        try:
            e2.set_particles_dev(allpos.data_ptr(), allmass.data_ptr(), ntot, oldacc_ptr=old.data_ptr() if old is not None else None)
            a2 = torch.zeros((ntot, 3), dtype=torch.float64, device=dev); p2 = torch.zeros(ntot, dtype=torch.float64, device=dev)
            if check_pm:
                # the reference PM of this check runs on cuFFT: independent of the slab PM under test and of the engine's
                # own transform passes (which the single-GPU bench and tests check against cuFFT and the CPU arm)
                import os
                prev = os.environ.get("B200_PM_FFT")
                os.environ["B200_PM_FFT"] = "cufft"
                try:
                    e2.gravpm_init_periodic(self.box, self.asmth, self.nmesh, self.G)
                finally:
                    if prev is None:
                        os.environ.pop("B200_PM_FFT", None)
                    else:
                        os.environ["B200_PM_FFT"] = prev
                g2 = torch.zeros((ntot, 3), dtype=torch.float64, device=dev)
                e2.gravpm_force_dev(g2.data_ptr(), None)
                torch.cuda.synchronize()
                gref = g2[targets.long()]
                out["pm_max_err_over_max"] = float((gref - gpm[sample]).abs().max() / g2.abs().max())
                del g2
            else:
                e2.walk_set_mesh(self.box, self.asmth, self.nmesh, self.G)
                gref = gpm[sample]
            e2.force_tree_build(self.box, toplevel_depth=self.topdepth)
            e2.grav_short_tree_dev(par, a2.data_ptr(), p2.data_ptr(), active_ptr=targets.data_ptr(), nactive=int(targets.shape[0]))
            torch.cuda.synchronize()
            ref = a2[targets.long()]
            mean = torch.sqrt(((ref + gref) ** 2).sum(1)).mean()
            out["acc_max_err_over_mean"] = float((ref - acc[sample]).abs().max() / mean)
            out["compared"] = "sampled own targets of every rank: sharded step vs one engine holding the whole box (same forced top tree)"
            out["tolerance"] = 1e-6
            out["ok"] = bool(out["acc_max_err_over_mean"] < 1e-6 and out.get("pm_max_err_over_max", 0.0) < 1e-6)
        finally:
            e2.close()
        return out


class ShardedSPH:
    """SPH density + hydro force on a domain-sharded gas distribution ("SPH hydro is
    symmetric => halo width max(h) via hmax").  Same domain cut and ghost import as ShardedTreePM:

      1. every rank imports the gas of the cell layer adjacent to its domain from both neighbours
         (positions, masses and everything the velocity / entropy predictors read), builds the gas
         tree over own + ghost particles and runs the density pass for its OWN particles only
         (b200_sph_set_active): density is a gather over r < h_i, so it needs no neighbour
         smoothing length and is exact as long as the ghost layer is wider than h_i;
      2. the owners send the converged Hsml, Density, EgyWtDensity, DhsmlEgyDensityFactor, DivVel and
         CurlVel of exactly those particles back along the same route (the reference gets them through
         its result exchange, treewalk.c:560-793); the tree's hmax is refreshed;
      3. the symmetric hydro pass runs for the own particles with the ghosts as sources.

    A pair (i, j) interacts when r < max(h_i, h_j), so the ghost layer must be wider than the largest
    smoothing length anywhere: checked after the density pass (all-reduce of max h), ValueError otherwise.
    """

    GHOST_COLS = 3 + 1 + 3 + 1 + 1 + 1 + 3 + 3 + 3      # pos, mass, vel, hsml, entropy, dtentropy, fullacc, gravpm, hydroacc
    STATE_KEYS = ("hsml", "density", "egywtdensity", "dhsmlfac", "divvel", "curlvel")

    def __init__(self, engine, box, topdepth, dist=None, device="cuda"):
        self.e = engine
        self.comm = Comm(dist)
        self.rank, self.world = self.comm.rank, self.comm.world
        self.box = float(box)
        self.dom = Domain(box, topdepth, self.rank, self.world)
        self.device = torch.device(device)
        # every tensor operation of this class is issued on the ENGINE's stream: the engine reads and writes these tensors
        # through raw device pointers, which torch's own stream would not order
        self.stream = torch.cuda.ExternalStream(engine.stream(), device=self.device)

    def _enter(self):
        self._caller = torch.cuda.current_stream(self.device)
        self.stream.wait_stream(self._caller)           # inputs made on the caller's stream
        return torch.cuda.stream(self.stream)

    def _leave(self):
        self._caller.wait_stream(self.stream)           # results are used on the caller's stream

    def load(self, pos, mass, hsml, vel=None, entropy=None, dtentropy=None, fullacc=None, gravpm=None, hydroacc=None):
        """Own gas particles (device tensors, every x inside the rank's layers).  The state stays in HBM: the ghosts arrive
        by neighbour exchange of device tensors and the engine takes device pointers (include/b200force.h, SPH section)."""
        with self._enter():
            n = pos.shape[0]
            z3 = lambda a: torch.zeros((n, 3), dtype=torch.float64, device=self.device) if a is None else a
            one = torch.ones(n, dtype=torch.float64, device=self.device)
            cols = [pos, mass.to(torch.float64)[:, None], z3(vel), hsml[:, None], (one if entropy is None else entropy)[:, None],
                    (0 * one if dtentropy is None else dtentropy)[:, None], z3(fullacc), z3(gravpm), z3(hydroacc)]
            own = torch.cat(cols, dim=1).contiguous()
            self.to_l, self.to_r = self.dom.ghost_sets(pos[:, 0])
            fr, fl = self.comm.neighbour_exchange_var(own[self.to_l], own[self.to_r])
            self.n_from_left, self.n_from_right = fl.shape[0], fr.shape[0]
            allp = torch.cat([own, fl, fr], dim=0)
            self.n_own, self.n_tot = n, allp.shape[0]
            col = lambda a, b: allp[:, a:b].contiguous() if b - a > 1 else allp[:, a].contiguous()
            st = self.dev = dict(pos=col(0, 3), mass=allp[:, 3].to(torch.float32).contiguous(), vel=col(4, 7), hsml=col(7, 8), entropy=col(8, 9),
                                 dtentropy=col(9, 10), fullacc=col(10, 13), gravpm=col(13, 16), hydroacc=col(16, 19),
                                 type=torch.zeros(self.n_tot, dtype=torch.uint8, device=self.device),
                                 own=torch.arange(n, dtype=torch.int32, device=self.device))
            del allp, own
            self.e.set_particles_dev(st["pos"].data_ptr(), st["mass"].data_ptr(), self.n_tot, type_ptr=st["type"].data_ptr())
            self.e.force_tree_build(self.box, mask=1)
            self.e.sph_set_gas(st["hsml"], vel=st["vel"], entropy=st["entropy"], dtentropy=st["dtentropy"], fullacc=st["fullacc"],
                               gravpm=st["gravpm"], hydroacc=st["hydroacc"])
            self.e.sph_set_active(st["own"])
            self._leave()
            return self.n_tot - n

    def set_active(self, active_own):
        """Restrict the targets of the next density / hydro pass to these own particles (int32 device tensor of indices
        < n_own): the active particles of a sub-step (timestep.c:1334-1431); everything else stays a source."""
        with self._enter():
            self.e.sph_set_active(active_own.to(torch.int32).contiguous())

    def set_mixed(self, bins_own, tables, active_own, ghost_bin=12):
        """A sub-step with mixed time bins after load(): per-particle time bins (uint8 device tensor for the own particles;
        the ghosts, sources only, get ghost_bin) with the per-bin factor tables of Engine.sph_set_timebins, the state of the
        last density pass for the particles that stay inactive, and the active own particles as targets."""
        with self._enter():
            bh = torch.cat([bins_own.to(torch.uint8), torch.full((self.n_tot - self.n_own,), ghost_bin, dtype=torch.uint8, device=self.device)]).contiguous()
            self.e.sph_set_timebins(bh, bh, tables)
            self.e.sph_set_state(**{k: self.state[k] for k in ("density", "egywtdensity", "dhsmlfac", "divvel", "curlvel")})
            self.e.sph_set_active(active_own.to(torch.int32).contiguous())

    def density(self, sp, DoEgyDensity=0):
        with self._enter():

            d = self.e.density(sp, update_hsml=1, DoEgyDensity=DoEgyDensity, device=self.device)
            n = self.n_own
            # converged state of the particles I exported, back to the ranks that hold them as ghosts
            st = torch.stack([d[k][:n] for k in self.STATE_KEYS], dim=1)
            fr, fl = self.comm.neighbour_exchange_var(st[self.to_l].contiguous(), st[self.to_r].contiguous())
            assert fl.shape[0] == self.n_from_left and fr.shape[0] == self.n_from_right
            g = torch.cat([fl, fr], dim=0)
            full = {}
            for c, k in enumerate(self.STATE_KEYS):
                full[k] = torch.cat([d[k][:n], g[:, c]]).contiguous()
            hmax = full["hsml"][:n].max().reshape(1) if n else torch.zeros(1, dtype=torch.float64, device=self.device)
            if self.world > 1:
                self.comm.dist.all_reduce(hmax, op=self.comm.dist.ReduceOp.MAX)
            if self.world > 1 and not self.dom.seamwidth > float(hmax.item()):
                raise ValueError("top-tree cells (%.4g) must be wider than the largest smoothing length (%.4g): lower topdepth"
                                 % (self.dom.cellwidth, float(hmax.item())))
            self.e.sph_set_state(density=full["density"], egywtdensity=full["egywtdensity"], dhsmlfac=full["dhsmlfac"],
                                 divvel=full["divvel"], curlvel=full["curlvel"])
            if self.n_tot > n:
                self.e.sph_set_hsml_range(full["hsml"][n:].contiguous(), n)
            self.state = full
            self._leave()
            return {k: v[:n] for k, v in d.items()}

    def hydro_force(self, sp):
        with self._enter():

            h = self.e.hydro_force(sp, device=self.device)
            self._leave()
            return {k: v[:self.n_own] for k, v in h.items()}
