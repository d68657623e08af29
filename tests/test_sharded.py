"""Multi-GPU host logic.  CPU: world_size-2 gloo processes exercise the domain
cut, ghost import, halo-plane exchange and the slab-FFT transposes (the message
pattern csrc/sharded.cu issues over NCCL, stated in torch) against
single-process numpy.  GPU: b200_sharded_force_step with world size 1 against the
unsharded engine path, and (when >= 2 GPUs are visible) 2 ranks under torchrun."""
import importlib
import os
import subprocess
import sys
import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = 43.0071

WORKER = r'''
import os, sys, importlib
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %(root)r)
sh = importlib.import_module("mp-gadget_b200.sharded")
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
comm = sh.Comm(dist)
rng = np.random.default_rng(5)                 # same stream on every rank
box, d, N, h = 16.0, 3, 16, 4
nx, nz = N // world, N // 2 + 1
# ---- slab FFT: 2-D rfft per owned plane, transpose, 1-D fft along x == rfftn
field = rng.standard_normal((N, N, N))
mine = torch.from_numpy(field[rank * nx:(rank + 1) * nx].copy())
c2 = torch.view_as_real(torch.fft.rfft2(mine)).contiguous()                 # [nx][N][nz][2]
cT = torch.empty_like(c2)
sb = torch.empty((world, nx, nx, nz, 2), dtype=torch.float64); rb = torch.empty_like(sb)
sh.slab_transpose_forward(comm, c2, cT, sb, rb)
full = np.fft.rfftn(field)                                                   # [x][y][kz]
got = torch.fft.fft(torch.view_as_complex(cT), dim=1).numpy()                # [jy][kx][kz]
want = np.transpose(full[:, rank * nx:(rank + 1) * nx, :], (1, 0, 2))
assert np.abs(got - want).max() < 1e-10, "forward transpose"
back = torch.empty_like(c2)
sh.slab_transpose_backward(comm, back, cT, sb, rb)
assert torch.equal(back, c2), "backward transpose"
# ---- halo add / fill on planes labelled by their global index
real = torch.zeros((nx + 2 * h, N, N), dtype=torch.float64)
for l in range(nx + 2 * h):
    real[l] = float((rank * nx - h + l) %% N) + 1.0                          # deposit "1 + global plane" everywhere
a = torch.empty((h, N, N), dtype=torch.float64); b = torch.empty_like(a)
sh.halo_add(comm, real, nx, h, a, b)
for l in range(h, h + nx):                                                   # owned planes: own + what neighbours spilled
    g = (rank * nx - h + l) %% N
    extra = (1.0 + g) * ((1 if l < 2 * h else 0) + (1 if l >= nx else 0))
    assert torch.all(real[l] == 1.0 + g + extra), ("halo_add", l)
for l in range(h, h + nx):
    real[l] = float((rank * nx - h + l) %% N)
sh.halo_fill(comm, real, nx, h, a, b)
for l in range(nx + 2 * h):
    assert torch.all(real[l] == float((rank * nx - h + l) %% N)), ("halo_fill", l)
# ---- domain + ghosts: every particle within rcut of an own particle is own or imported
dom = sh.Domain(box, d, rank, world)
pos = torch.from_numpy(rng.random((4000, 3)) * box)
owner = dom.owner_of(pos[:, 0])
own = pos[owner == rank]
to_l, to_r = dom.ghost_sets(own[:, 0])
pm = torch.cat([own, torch.ones(len(own), 1, dtype=torch.float64)], 1)
fr, fl = comm.neighbour_exchange_var(pm[to_l], pm[to_r])
have = torch.cat([own, fr[:, :3], fl[:, :3]], 0).numpy()
rcut = 0.95 * dom.cellwidth
allp = pos.numpy(); o = own.numpy()
for p in o[::7]:
    dd = allp - p; dd -= box * np.round(dd / box)
    need = allp[(np.abs(dd) < rcut).all(1)]
    for q in need:
        assert (np.abs(have - q).sum(1) == 0).any(), "missing ghost"
assert len(have) == len(np.unique(have, axis=0)), "duplicate ghosts"
# ---- own-cell mask agrees with the layer of the cell centres
mask = dom.own_cell_mask("cpu").numpy()
m = np.arange(8 ** d); ix = np.zeros_like(m)
for l in range(d): ix |= ((m >> (3 * l)) & 1) << l
assert np.array_equal(mask, (ix // dom.per) == rank)
cnt = torch.tensor([float(mask.sum())]); comm.all_reduce_sum(cnt)
assert cnt.item() == 8 ** d
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def _torchrun(code, nproc, extra_env=None, timeout=600):
    path = os.path.join(ROOT, "tests", "_worker_tmp_%d.py" % os.getpid())
    with open(path, "w") as f:
        f.write(code)
    env = dict(os.environ)
    env.update(extra_env or {})
    try:
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
                            "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 2000), path],
                           capture_output=True, text=True, timeout=timeout, env=env)
    finally:
        os.unlink(path)
    return r


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_host_logic_gloo(world):
    r = _torchrun(WORKER % {"root": ROOT}, world)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("ok") == world


def _setup(ics, ng=32, seed=3):
    pos, mass = ics.zeldovich_lattice(ng, float(ng), seed=seed)
    return pos, mass, float(ng), ics.default_nmesh(ng)


@pytest.mark.gpu
def test_sharded_world1_equals_unsharded(b200, ics):
    """Slab PM (2-D FFT + transpose + 1-D FFT + fused difference/readout) and the
    forced-top-tree walk over a target subset reproduce the single-GPU path."""
    sh = importlib.import_module("mp-gadget_b200.sharded")
    pos, mass, box, nmesh = _setup(ics)
    n = len(mass)
    par = ics.tree_params(box, n, treeusebh=1)
    e0 = b200.Engine(0)
    e0.set_particles(pos, mass)
    e0.gravpm_init_periodic(box, 1.5, nmesh, G)
    g0, _ = e0.gravpm_force()
    e0.force_tree_build(box, toplevel_depth=2)
    a0, p0, _ = e0.grav_short_tree(par)
    e0.close()
    e1 = b200.Engine(0)
    s = sh.ShardedTreePM(e1, box, nmesh, 1.5, G, topdepth=2, dist=None, rcut_cells=par["Rcut"])
    tp, tm = torch.from_numpy(pos).cuda(), torch.from_numpy(mass).cuda()
    torch.cuda.synchronize()
    g1, a1, p1 = s.force_step(tp, tm, None, par)
    g1, a1, p1 = g1.cpu().numpy(), a1.cpu().numpy(), p1.cpu().numpy()
    assert s.n_tot == s.n_own == n
    e1.close()
    assert np.abs(g1 - g0).max() <= 1e-10 * np.abs(g0).max()
    assert np.abs(a1 - a0).max() <= 1e-12 * np.sqrt((a0 ** 2).sum(1)).mean()
    assert np.abs(p1 - p0).max() <= 1e-12 * np.abs(p0).max()


GPU_WORKER = r'''
import os, sys, importlib
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %(root)r)
pkg = importlib.import_module("mp-gadget_b200"); ics = importlib.import_module("mp-gadget_b200.ics")
sh = importlib.import_module("mp-gadget_b200.sharded")
local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
G = 43.0071
pos, mass = ics.zeldovich_lattice(32, 32.0, seed=3); box, nmesh = 32.0, 96
par = ics.tree_params(box, len(mass), treeusebh=1)
e = pkg.Engine(local)
s = sh.ShardedTreePM(e, box, nmesh, 1.5, G, topdepth=3, dist=dist, device="cuda:%%d" %% local, rcut_cells=par["Rcut"])
tp = torch.from_numpy(pos).cuda(); tm = torch.from_numpy(mass).cuda()
sel = s.dom.owner_of(tp[:, 0]) == rank
op, om = tp[sel].contiguous(), tm[sel].contiguous()
torch.cuda.synchronize()
g, a, p = s.force_step(op, om, None, par)
nghost = s.n_tot - s.n_own
par2 = dict(par); par2["TreeUseBH"] = 0
old = (a + g).clone()
g, a, p = s.force_step(op, om, old, par2)          # second pass: relative criterion fed by the first
chk = s.parity_check(op, om, a, g, par2, nsample=2000)
assert chk["ok"] and chk["pm_checked"], chk
g, a, p = s.force_step(op, om, None, par)
idx = torch.nonzero(sel)[:, 0].cpu().numpy()
np.savez(os.path.join(%(out)r, "shard_%%d.npz" %% rank), idx=idx, g=g.cpu().numpy(), a=a.cpu().numpy(), p=p.cpu().numpy(), nghost=nghost)
dist.barrier(); dist.destroy_process_group()
'''


@pytest.mark.gpu
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_two_gpus_equal_one(b200, ics, tmp_path):
    pos, mass, box, nmesh = _setup(ics)
    par = ics.tree_params(box, len(mass), treeusebh=1)
    e0 = b200.Engine(0)
    e0.set_particles(pos, mass)
    e0.gravpm_init_periodic(box, 1.5, nmesh, G)
    g0, _ = e0.gravpm_force()
    e0.force_tree_build(box, toplevel_depth=3)
    a0, p0, _ = e0.grav_short_tree(par)
    e0.close()
    r = _torchrun(GPU_WORKER % {"root": ROOT, "out": str(tmp_path)}, 2)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    seen = np.zeros(len(mass), bool)
    for rank in range(2):
        d = np.load(os.path.join(str(tmp_path), "shard_%d.npz" % rank))
        idx = d["idx"]
        assert not seen[idx].any()
        seen[idx] = True
        assert np.abs(d["g"] - g0[idx]).max() <= 1e-10 * np.abs(g0).max()
        assert np.abs(d["a"] - a0[idx]).max() <= 1e-11 * np.sqrt((a0 ** 2).sum(1)).mean()
        assert np.abs(d["p"] - p0[idx]).max() <= 1e-11 * np.abs(p0).max()
    assert seen.all()


def _gas_setup(ics, ng=24, seed=5):
    pos, mass = ics.zeldovich_lattice(ng, float(ng), seed=seed, rms=0.3)
    rng = np.random.default_rng(2)
    n = len(mass)
    return dict(pos=pos, mass=mass, box=float(ng), vel=rng.standard_normal((n, 3)) * 0.1, entropy=1 + 0.5 * rng.random(n),
                h0=np.full(n, 1.3))


def _sph_par(b200):
    return b200.sph_params(KernelType=1, DensityIndependentSphOn=1, MinGasHsml=1e-4, atime=0.5, hubble=0.2, dloga_bin=0.01)


def _sph_single(b200, gas):
    e = b200.Engine(0)
    n = len(gas["mass"])
    e.set_particles(gas["pos"], gas["mass"], type=np.zeros(n, np.uint8))
    e.force_tree_build(gas["box"], mask=1)
    e.sph_set_gas(gas["h0"], vel=gas["vel"], entropy=gas["entropy"])
    sp = _sph_par(b200)
    d = e.density(sp, update_hsml=1, DoEgyDensity=1)
    h = e.hydro_force(sp)
    e.close()
    return d, h


@pytest.mark.gpu
def test_sharded_sph_world1_equals_unsharded(b200, ics):
    """ShardedSPH with one rank (no ghosts, active set = all) reproduces the plain engine calls."""
    sh = importlib.import_module("mp-gadget_b200.sharded")
    gas = _gas_setup(ics)
    d0, h0 = _sph_single(b200, gas)
    e = b200.Engine(0)
    s = sh.ShardedSPH(e, gas["box"], topdepth=2, dist=None)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    s.load(t(gas["pos"]), t(gas["mass"]), t(gas["h0"]), vel=t(gas["vel"]), entropy=t(gas["entropy"]))
    sp = _sph_par(b200)
    d1 = s.density(sp, DoEgyDensity=1)
    h1 = s.hydro_force(sp)
    e.close()
    for k in ("hsml", "density", "egywtdensity", "dhsmlfac", "divvel", "curlvel"):       # the state never left the device
        assert d1[k].is_cuda and np.array_equal(d1[k].cpu().numpy(), d0[k]), k
    for k in ("acc", "dtentropy", "maxsignalvel"):
        assert np.array_equal(h1[k].cpu().numpy(), h0[k]), k


SPH_GPU_WORKER = r'''
import os, sys, importlib
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %(root)r)
pkg = importlib.import_module("mp-gadget_b200"); ics = importlib.import_module("mp-gadget_b200.ics")
sh = importlib.import_module("mp-gadget_b200.sharded")
local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
ng = 24
pos, mass = ics.zeldovich_lattice(ng, float(ng), seed=5, rms=0.3)
rng = np.random.default_rng(2); n = len(mass)
vel = rng.standard_normal((n, 3)) * 0.1; ent = 1 + 0.5 * rng.random(n); h0 = np.full(n, 1.3)
e = pkg.Engine(local)
s = sh.ShardedSPH(e, float(ng), topdepth=2, dist=dist, device="cuda:%%d" %% local)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
sel = (s.dom.owner_of(t(pos)[:, 0]) == rank).cpu().numpy()
nghost = s.load(t(pos[sel]), t(mass[sel]), t(h0[sel]), vel=t(vel[sel]), entropy=t(ent[sel]))
sp = pkg.sph_params(KernelType=1, DensityIndependentSphOn=1, MinGasHsml=1e-4, atime=0.5, hubble=0.2, dloga_bin=0.01)
d = s.density(sp, DoEgyDensity=1)
h = s.hydro_force(sp)
np.savez(os.path.join(%(out)r, "sph_%%d.npz" %% rank), idx=np.nonzero(sel)[0], nghost=nghost,
         **{"d_" + k: d[k].cpu().numpy() for k in ("hsml", "density", "egywtdensity", "dhsmlfac", "divvel", "curlvel")},
         **{"h_" + k: h[k].cpu().numpy() for k in ("acc", "dtentropy", "maxsignalvel")})
dist.barrier(); dist.destroy_process_group()
'''


@pytest.mark.gpu
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_sph_two_gpus_equal_one(b200, ics, tmp_path):
    """Density + hydro on 2 GPUs (ghost import, converged ghost state sent back, hmax refresh) equal
    the single-GPU result for every particle."""
    gas = _gas_setup(ics)
    d0, h0 = _sph_single(b200, gas)
    r = _torchrun(SPH_GPU_WORKER % {"root": ROOT, "out": str(tmp_path)}, 2)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    seen = np.zeros(len(gas["mass"]), bool)
    close = lambda a, b, tol: np.abs(a - b).max() <= tol * (np.abs(b).max() + 1e-300)
    for rank in range(2):
        z = np.load(os.path.join(str(tmp_path), "sph_%d.npz" % rank))
        idx = z["idx"]
        assert not seen[idx].any() and int(z["nghost"]) > 0
        seen[idx] = True
        for k in ("hsml", "density", "egywtdensity", "dhsmlfac", "divvel", "curlvel"):
            assert close(z["d_" + k], d0[k][idx], 1e-11), k
        for k in ("acc", "dtentropy", "maxsignalvel"):
            assert close(z["h_" + k], h0[k][idx], 1e-10), k
    assert seen.all()


DOMAIN_WORKER = r'''
import os, sys, importlib
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
import oracle                                   # test side: keys / leaf lookup of a box without a GPU
import domain_scenarios as DS
dom = importlib.import_module("mp-gadget_b200.domain")
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
GOLD = np.load(os.path.join(%(root)r, "tests", "golden", "ref_peano.npz"))
for k, case in enumerate(DS.TOPTREE_CASES):
    # rank r holds the particles the reference fixture gave to rank r (world 2); with more ranks the extra ones hold few
    if rank < 2:
        pos = DS.clustered(case["n"][rank], case["box"], case["seeds"][rank])
    else:
        pos = DS.clustered(300 * rank, case["box"], 50 + rank)
    sub = case["subsample"]
    keys = oracle.peano_keys(pos[::sub][: len(pos) // sub], case["box"])
    T, leaf, nleaf = dom.global_toptree(keys, case["ntopleaves"], dist)
    if world == 2:                              # the reference's own two-rank tree, node for node
        for f in DS.TOPTREE_FIELDS:
            assert np.array_equal(T.tree[f], GOLD["toptree/%%d/%%s" %% (k, f)]), (k, f)
        assert nleaf == int(GOLD["toptree/%%d/nleaf" %% k]) and np.array_equal(leaf, GOLD["toptree/%%d/leaf" %% k])
    # every rank ends with the same tree
    sig = torch.tensor([int(T.tree["StartKey"].sum() %% (1 << 62)), int(T.tree["Count"].sum()), nleaf], dtype=torch.int64)
    lo, hi = sig.clone(), sig.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert torch.equal(lo, hi), "trees differ between ranks"
    # balance: counts per leaf of all particles -> tasks
    tl = oracle.topleaf(oracle.peano_keys(pos, case["box"]), *dom.topnode_arrays(T, leaf))
    task, counts = dom.balance(np.bincount(tl, minlength=nleaf), dist)
    assert counts.sum() >= sum(case["n"]) and np.array_equal(task, oracle.domain_assign_balanced(world, counts))
    load = np.bincount(task, weights=counts, minlength=world)
    assert set(task) == set(range(world)) and load.max() <= 1.6 * load.mean(), load
# a node budget only the LAST rank's merge exceeds: the failure must reach every rank (MPIU_Any, domain.c:1266-1272)
# instead of leaving the others in the next collective
case = DS.TOPTREE_CASES[0]
pos = DS.clustered(case["n"][min(rank, 1)], case["box"], 7 + rank)
keys = oracle.peano_keys(pos[::4], case["box"])
for maxnodes in (40, 200):
    try:
        dom.global_toptree(keys, case["ntopleaves"], dist, maxnodes=maxnodes)
        raised = 0
    except dom.TopTreeOverflow:
        raised = 1
    t = torch.tensor([raised], dtype=torch.int64); lo, hi = t.clone(), t.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert int(lo) == int(hi), "ranks disagree about the failure"
assert raised == 1 or maxnodes > 40
print("ok", flush=True)
dist.destroy_process_group()
'''


@pytest.mark.parametrize("world", [2, 3])
def test_domain_toptree_and_balance_gloo(world):
    """domain_determine_global_toptree + domain_balance over torch.distributed: world 2 reproduces the reference's own
    two-rank top tree node for node; 3 ranks exercise the pairwise merge schedule with an odd rank count (rank 2 merges
    into rank 0 at separation 2)."""
    r = _torchrun(DOMAIN_WORKER % {"root": ROOT}, world)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("ok") == world


EXCHANGE_WORKER = r'''
import os, sys, importlib
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
import oracle
import domain_scenarios as DS
dom = importlib.import_module("mp-gadget_b200.domain")
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
box = 1000.0
n = 4000 + 1500 * rank
pos = DS.clustered(n, box, 20 + rank)
ids = (np.arange(n, dtype=np.int64) + (rank << 32))
vel = np.random.default_rng(rank).standard_normal((n, 3))
# decomposition: keys -> global top tree -> top leaf -> counts -> balanced tasks -> plan (device kernels stood in by the oracle here)
keys = oracle.peano_keys(pos, box)
T, leaf, nleaf = dom.global_toptree(keys[::8][: n // 8], 8 * world, dist)
top = dom.topnode_arrays(T, leaf)
tl = oracle.topleaf(keys, *top)
task, counts = dom.balance(np.bincount(tl, minlength=nleaf), dist)
lst, togo, ng = oracle.exchange_plan(np.ones(n, np.uint8), np.zeros(n, np.uint8), tl, task, world, rank)
new = dom.exchange(dict(pos=torch.from_numpy(pos), vel=torch.from_numpy(vel), id=torch.from_numpy(ids)), lst, task[tl[lst]], dist)
# every particle now lives on the task that owns its top leaf, with its own payload
npos = new["pos"].numpy(); nid = new["id"].numpy()
assert len(npos) == len(nid) == len(new["vel"]) == counts[task == rank].sum(), (len(npos), counts[task == rank].sum())
assert (task[oracle.topleaf(oracle.peano_keys(npos, box), *top)] == rank).all()
src = nid >> 32; loc = nid & 0xffffffff
for r in range(world):
    m = src == r
    assert np.array_equal(npos[m], DS.clustered(4000 + 1500 * r, box, 20 + r)[loc[m]])
    assert np.array_equal(new["vel"].numpy()[m], np.random.default_rng(r).standard_normal((4000 + 1500 * r, 3))[loc[m]])
# nothing lost, nothing doubled
tot = torch.tensor([len(nid), int(nid.sum() %% (1 << 40))], dtype=torch.int64); dist.all_reduce(tot)
want_n = sum(4000 + 1500 * r for r in range(world))
want_s = sum(int((np.arange(4000 + 1500 * r, dtype=np.int64) + (r << 32)).sum() %% (1 << 40)) for r in range(world))
assert int(tot[0]) == want_n and len(np.unique(nid)) == len(nid)
# kept particles stay in order in front, arrivals follow by source rank
kept = src == rank
assert kept[: kept.sum()].all() and (np.diff(loc[kept]) > 0).all() and (np.diff(src[~kept]) >= 0).all()
# ---- the reference's own exchange tests (libgadget/tests/test_exchange.c): layout ID %% NTask
def by_id(nlocal, garbage_every=0):
    ids = torch.arange(nlocal, dtype=torch.int64) + nlocal * rank if nlocal else torch.zeros(0, dtype=torch.int64)
    live = torch.ones(nlocal, dtype=torch.bool)
    if garbage_every:
        live[::garbage_every] = False                              # garbage is never on the exchange list (exchange.c:427-430)
    tgt = ids %% world
    leaving = torch.nonzero(live & (tgt != rank)).flatten()
    new = dom.exchange(dict(id=ids, live=live), leaving.numpy(), tgt[leaving].numpy(), dist)
    nid, nlive = new["id"], new["live"]
    assert bool(((nid[nlive] %% world) == rank).all())            # test_exchange.c:78
    tot = torch.tensor([int(nlive.sum())]); dist.all_reduce(tot)
    return int(tot), nid[nlive]
tot, mine = by_id(48)                                              # test_exchange
assert tot == 48 * world and len(torch.unique(mine)) == len(mine)
live_before = torch.tensor([48 - len(range(0, 48, 5))]); dist.all_reduce(live_before)
tot, mine = by_id(48, garbage_every=5)                             # test_exchange_with_garbage
assert tot == int(live_before)
tot, mine = by_id(48 * world if rank == 0 else 0)                 # test_exchange_uneven: everything starts on task 0
assert tot == 48 * world and len(mine) == 48
print("ok", flush=True)
dist.destroy_process_group()
'''


@pytest.mark.parametrize("world", [2, 3])
def test_domain_exchange_gloo(world):
    """The whole decomposition on CPU ranks: keys, global top tree, balanced leaf assignment, exchange plan, and the
    variable-size all-to-all of the particle state; afterwards every particle sits on the task owning its top leaf.
    Then the reference's own exchange tests (tests/test_exchange.c: layout ID mod NTask, with garbage, everything on task 0)."""
    r = _torchrun(EXCHANGE_WORKER % {"root": ROOT}, world)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("ok") == world
