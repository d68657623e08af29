"""Run the step-loop scenarios of tests/step_scenarios.py through mp-gadget_b200/steploop.py's
StepEngine on the CPU emulation build of csrc/steploop.cu (tests/emul/build.py) and check them
against the reference's golden vectors.  TEST INFRASTRUCTURE ONLY.  Started by
tests/test_step_emul.py in a subprocess with OMP_WAIT_POLICY=passive (256 OS threads per block)."""
import ctypes as C
import importlib
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, HERE)
import build as EB                   # noqa: E402
import step_scenarios as SC          # noqa: E402
import test_step as TS               # noqa: E402


PKG = importlib.import_module("mp-gadget_b200")


class EmulEngine(PKG.Engine):
    """The harness Engine with its methods bound to the emulation library instead of libb200force.so: every method
    whose entry point the emulation exports (particles, pm_init, step loop, domain keys) works unchanged."""

    def __init__(self):
        self.L = C.CDLL(EB.build())
        self.L.b200_last_error.restype = C.c_char_p
        self.L.b200_ctx_destroy.restype = None
        self.ctx = C.c_void_p()
        assert self.L.b200_ctx_create(C.byref(self.ctx), C.c_int(0)) == 0
        self.n = 0
        self.nmesh = 0


def dropin(which):
    """The reference's own loop (oracle/ref_driver.c) with its step-loop calls redirected by ld --wrap to
    mp-gadget_b200/host/libgadget_step_shims.c, which forwards to the emulated b200_step_*."""
    from oracle import ref as R
    so = EB.build_dropin()
    if so is None:
        print("skip: no /root/reference and no prebuilt drop-in library")
        return
    S = R.RefStep(nthreads=2, arena_gib=1.0, so=so, **SC.TIMELINE)
    if which == "dropin_primitives":
        TS.check_primitives(SC.run_primitives(S, SC.primitives_inputs()))
    elif which == "dropin_gas":
        TS.check_gas(SC.run_gas_hierarchy(S, SC.gas_hierarchy_inputs()))
    else:
        TS.check_hierarchy(SC.run_hierarchy(S, SC.hierarchy_inputs()))
    print(which + " ok")


def domain():
    """csrc/domain_keys.cu under emulation against the reference's known-answer keys and compiled peano.c / domain.h."""
    import domain_scenarios as DS
    G = np.load(os.path.join(ROOT, "tests", "golden", "ref_peano.npz"))
    e = EmulEngine()
    L, ctx = e.L, e.ctx
    p = lambda a: C.c_void_p(a.ctypes.data)

    def keys_of(pos, box):
        pos = np.ascontiguousarray(pos, np.float64)
        e.set_particles(pos, np.ones(len(pos), np.float32))
        k = np.zeros(len(pos), np.uint64)
        assert L.b200_domain_peano_keys(ctx, C.c_double(box), p(k)) == 0, L.b200_last_error(ctx)
        return k
    pos4, box4 = DS.peano_test_positions()
    assert np.array_equal(keys_of(pos4, box4), G["known_keys"])
    pos, box = DS.random_positions()
    assert np.array_equal(keys_of(pos, box), G["random_keys"])
    top = [np.ascontiguousarray(a) for a in DS.refined_toptree()]
    assert L.b200_domain_set_topnodes(ctx, C.c_int32(len(top[0])), p(top[0]), p(top[1]), p(top[2]), p(top[3])) == 0
    leaf = np.zeros(len(pos), np.int32)
    assert L.b200_domain_topleaf(ctx, p(leaf)) == 0, L.b200_last_error(ctx)
    assert np.array_equal(leaf, G["topleaf"])
    nleaf = int(top[3].max()) + 1
    counts = np.zeros(nleaf, np.int64)
    assert L.b200_domain_leaf_counts(ctx, C.c_int32(nleaf), p(counts)) == 0, L.b200_last_error(ctx)
    assert np.array_equal(counts, np.bincount(G["topleaf"], minlength=nleaf))
    for nt, cost in DS.assign_cases()[:30]:
        task = np.zeros(len(cost), np.int32)
        assert L.b200_domain_assign_balanced(C.c_int32(nt), C.c_int32(len(cost)), p(cost), C.c_int32(1), p(task)) == 0
    import oracle
    tasks = np.zeros(nleaf, np.int32)
    assert L.b200_domain_assign_balanced(C.c_int32(4), C.c_int32(nleaf), p(counts), C.c_int32(1), p(tasks)) == 0
    for r in range(4):                       # the exchange plan of every rank against the oracle (itself pinned by exchange.c)
        lst = np.zeros(len(pos) + 1, np.int32); togo = np.zeros((4, 7), np.int64); nex = C.c_int64(); ng = C.c_int64()
        assert L.b200_domain_exchange_plan(ctx, p(tasks), C.c_int32(nleaf), C.c_int32(4), C.c_int32(r), C.byref(nex), C.byref(ng), p(togo), p(lst)) == 0, L.b200_last_error(ctx)
        olst, otogo, ong = oracle.exchange_plan(np.ones(len(pos), np.uint8), np.zeros(len(pos), np.uint8), G["topleaf"], tasks, 4, r)
        assert nex.value == len(olst) and np.array_equal(lst[:nex.value], olst) and np.array_equal(togo, otogo) and ng.value == ong
    for sub in (1, 7, 256):                  # the strided subsample keys of the top-tree build
        ks = np.zeros(max(len(pos) // sub, 1), np.uint64); ns = C.c_int64()
        assert L.b200_domain_sample_keys(ctx, C.c_double(box), C.c_int32(sub), p(ks), C.byref(ns)) == 0, L.b200_last_error(ctx)
        assert ns.value == len(pos) // sub and np.array_equal(ks[:ns.value], oracle.peano_keys(pos[::sub][: len(pos) // sub], box))
    bad = top[0].copy(); bad[0] = 0          # a daughter pointing at its parent must be refused, not loop
    assert L.b200_domain_set_topnodes(ctx, C.c_int32(len(bad)), p(bad), p(top[1]), p(top[2]), p(top[3])) != 0
    print("domain ok")


def decompose():
    """mp-gadget_b200/domain.py::decompose -- device keys / lookup / counts / plan (emulated kernels) around the host top
    tree, over torch.distributed when started under torchrun -- against the oracle; with several ranks the exchange follows
    and every particle must end on the task owning its top leaf."""
    import oracle
    import torch
    import domain_scenarios as DS
    dom = importlib.import_module("mp-gadget_b200.domain")
    dist = None
    if "RANK" in os.environ:
        import torch.distributed as dist
        dist.init_process_group("gloo")
    rank = dist.get_rank() if dist else 0
    world = dist.get_world_size() if dist else 1
    box = 1000.0
    n = 20000 + 7000 * rank
    pos = DS.clustered(n, box, 31 + rank)
    e = EmulEngine()
    e.set_particles(pos, np.ones(n, np.float32))
    d = dom.decompose(e, box, dist, overdecomposition=8, subsample=16)
    keys = oracle.peano_keys(pos, box)
    tl = oracle.topleaf(keys, *d["topnodes"])
    assert np.array_equal(tl, d["topleaf"])
    lst, togo, ng = oracle.exchange_plan(np.ones(n, np.uint8), np.zeros(n, np.uint8), tl, d["task_of_leaf"], world, rank)
    assert np.array_equal(lst, d["leaving"]) and np.array_equal(togo, d["togo"]) and ng == d["ngarbage"] == 0
    if world == 1:
        O = oracle.TopTree(d["tree"].maxnodes)
        assert O.local(keys[::16][: n // 16]) == 0
        lim = int(O.tree["Count"][0]) // 8
        O.truncate(lim, lim); O.global_refine(lim, lim)
        for f in DS.TOPTREE_FIELDS:
            assert np.array_equal(O.tree[f], d["tree"].tree[f]), f
        assert (d["task_of_leaf"] == 0).all() and len(d["leaving"]) == 0
    else:
        new = dom.exchange(dict(pos=torch.from_numpy(pos)), d["leaving"], d["target"], dist)
        npos = new["pos"].numpy()
        assert (d["task_of_leaf"][oracle.topleaf(oracle.peano_keys(npos, box), *d["topnodes"])] == rank).all()
        tot = torch.tensor([len(npos)], dtype=torch.int64); dist.all_reduce(tot)
        assert int(tot) == sum(20000 + 7000 * r for r in range(world)) and len(npos) == d["counts"][d["task_of_leaf"] == rank].sum()
        # domain_maintain: drift the new local set on the (emulated) device, look the leaves up again, exchange the movers
        SL = importlib.import_module("mp-gadget_b200.steploop")
        n2 = len(npos)
        vel = np.random.default_rng(100 + rank).standard_normal((n2, 3)) * 40.0
        e.set_particles(npos, np.ones(n2, np.float32))
        S = SL.StepEngine(e, np.log([0.1, 1.0]), lambda k, a, b: 0.5, lambda a: 0.1)
        S.n, S.box = n2, box
        st = SL.StepState(vel=vel.ctypes.data, BoxSize=box)
        assert e.L.b200_step_set_state(e.ctx, C.byref(st)) == 0
        S.drift(0, 1 << 40)                                             # ddrift = 0.5: displacements of ~20 in a box of 1000
        moved = S.get()["pos"]
        want = np.mod(npos + 0.5 * vel, box); want[want <= 0] += box
        assert np.abs(moved - want).max() < 1e-9
        m = dom.maintain(e, box, d["topnodes"], d["task_of_leaf"], dist)
        assert len(m["leaving"]) > 0 and len(m["leaving"]) < n2 // 2
        new2 = dom.exchange(dict(pos=torch.from_numpy(moved)), m["leaving"], m["target"], dist)
        p2 = new2["pos"].numpy()
        assert (d["task_of_leaf"][oracle.topleaf(oracle.peano_keys(p2, box), *d["topnodes"])] == rank).all()
        tot = torch.tensor([len(p2)], dtype=torch.int64); dist.all_reduce(tot)
        assert int(tot) == sum(20000 + 7000 * r for r in range(world))
        dist.destroy_process_group()
    print("decompose ok", flush=True)


def main(which):
    if which == "domain":
        return domain()
    if which == "decompose":
        return decompose()
    if which.startswith("dropin"):
        return dropin(which)
    SL = importlib.import_module("mp-gadget_b200.steploop")
    O = TS.make_oracle()
    cosmo = {k: float(TS.GOLD["cosmo/" + k]) for k in ("Omega0", "OmegaBaryon", "Hubble", "G")}
    ts = {k: float(TS.GOLD["tspar/" + k]) for k in TS.TSKEYS}
    S = SL.StepEngine(EmulEngine(), TS.GOLD["sync_loga"], O.factor, O.hubble, **cosmo, **ts)
    if which in ("primitives", "all"):
        # error behaviour: every entry point refuses to run without its prerequisites and says which one is missing
        S.n, S.box = 0, 1.0
        for call, word in ((lambda: S.drift(0, 1 << 30), "b200_step_set_state"), (lambda: S.build_active(), "b200_step_set_state"),
                           (lambda: S.kick(2), "b200_step_set_state"), (lambda: S.adopt_hydro(), "b200_step_set_state")):
            try:
                call()
                raise AssertionError("call without state accepted")
            except PKG.B200Error as ex:
                assert word in str(ex), str(ex)
        out = SC.run_primitives(S, SC.primitives_inputs())
        TS.check_primitives(out)
        d = SC.primitives_inputs(); d["type"] = d["type"].copy(); d["type"][100] = 5       # a live black hole is refused, loudly
        try:
            SC._load(S, d)
            raise AssertionError("black-hole particle accepted")
        except PKG.B200Error as ex:
            assert "black-hole" in str(ex)
        print("primitives ok")
    if which in ("hierarchy", "all"):
        rec = SC.run_hierarchy(S, SC.hierarchy_inputs())
        TS.check_hierarchy(rec)
        print("hierarchy ok")
    if which in ("gas", "all"):
        rec = SC.run_gas_hierarchy(S, SC.gas_hierarchy_inputs())
        TS.check_gas(rec)
        print("gas ok")
    if which in ("nonsplit", "all"):
        rec = SC.run_nonsplit(S, SC.hierarchy_inputs(seed=15, n=1536))
        TS.check_nonsplit(rec)
        print("nonsplit ok")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "all")
