"""ShardedSPH (world 1) cold / warm / mixed-bin sequence of bench.sharded_hydro_entry with diagnostics."""
import sys, importlib, numpy as np, torch
sys.path.insert(0, ".")
pkg = importlib.import_module("mp-gadget_b200"); ics = importlib.import_module("mp-gadget_b200.ics")
sh = importlib.import_module("mp-gadget_b200.sharded")
import os
ng = int(sys.argv[1]) if len(sys.argv) > 1 else 64
topd = int(sys.argv[2]) if len(sys.argv) > 2 else 3
world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
dist = None
if world > 1:
    import datetime, torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=120))
box = float(ng); dev = torch.device("cuda", local)
pos, mass = ics.bench_ics("displaced", ng, box, device="cuda:%d" % local)
n = pos.shape[0]
sp = pkg.sph_params(KernelType=2, DensityIndependentSphOn=1, MinGasHsml=1e-4, atime=0.1, hubble=3.0, dloga_bin=0.01)
e = pkg.Engine(local)
s = sh.ShardedSPH(e, box, topd, dist=dist, device="cuda:%d" % local)
if world > 1:
    keep = s.dom.owner_of(pos[:, 0]) == dist.get_rank()
    pos, mass = pos[keep].contiguous(), mass[keep].contiguous()
n = pos.shape[0]
gen = torch.Generator(device=dev); gen.manual_seed(1000)
vel = 0.05 * torch.randn((n, 3), dtype=torch.float64, device=dev, generator=gen)
ent = torch.ones(n, dtype=torch.float64, device=dev)
h0 = torch.full((n,), 3.0 * 0.8, dtype=torch.float64, device=dev)
import time
def step(hs, active=None, bins=None, tag=""):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    nghost = s.load(pos, mass, hs, vel=vel, entropy=ent)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    if bins is not None:
        tab = {k: np.zeros(47) for k in ("gravkick", "hydrokick", "dloga_pred", "drift")}; tab["dloga_bin"] = np.full(47, 0.01)
        s.set_mixed(bins, tab, active)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    try:
        d = s.density(sp, DoEgyDensity=1)
    except Exception as ex:
        hh = torch.zeros(s.n_tot, dtype=torch.float64, device=dev)
        print(tag, "FAILED", ex); raise
    torch.cuda.synchronize(); t3 = time.perf_counter()
    h = s.hydro_force(sp)
    torch.cuda.synchronize(); t4 = time.perf_counter()
    print("r%d" % local, tag, "wall ms: load %.1f set %.1f density %.1f hydro %.1f" % (1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), 1e3 * (t4 - t3)), flush=True)
    sel = slice(None) if active is None else active.long()
    print("r%d" % local, tag, "ghosts", nghost, "niter mean %.2f max %d" % (d["niter"][sel].double().mean().item(), d["niter"][sel].max().item()),
          "hsml min %.3f max %.3f" % (d["hsml"].min().item(), d["hsml"].max().item()), "dens min %.3g" % d["density"].min().item(),
          "t_dens %.1f t_hydro %.1f" % (e.timings()["sph_density"], e.timings()["sph_hydro"]), flush=True)
    return d
d = step(h0, tag="cold")
hw = (d["hsml"] * (1.0 + 0.01 * torch.randn(n, dtype=torch.float64, device=dev, generator=gen))).contiguous()
for k in range(2): step(hw, tag="warm%d" % k)
idx = torch.arange(n, device=dev)
act = idx[(idx % 4) == 0].to(torch.int32).contiguous()
bins = torch.where((idx % 4) == 0, 10, 12).to(torch.uint8)
for k in range(5): step(hw, active=act, bins=bins, tag="mixed%d" % k)
