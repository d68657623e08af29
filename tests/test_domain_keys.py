"""Domain keys: Peano-Hilbert keys (utils/peano.c:108-129, peano.h:15-21) and the top-leaf lookup (domain.h:71-78).
Golden tests/golden/ref_peano.npz (generator make_golden_peano.py): the 64 known-answer keys of the reference's own
tests/test_peano.c:107, its compiled peano.c on random positions, domain_get_topleaf over a refined top tree.
CPU: the oracle (a generated state machine, no stored tables).  The CUDA kernels run under emulation in
tests/test_step_emul.py[domain]; the hardware tests below (marker `gpu`) first ran green in round 2."""
import os
import sys
import numpy as np
import pytest

import oracle
from oracle import ref as R

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import domain_scenarios as DS        # noqa: E402

GOLD = np.load(os.path.join(HERE, "golden", "ref_peano.npz"))


def test_oracle_reproduces_test_peano_known_answers():
    pos, box = DS.peano_test_positions()
    assert np.array_equal(oracle.peano_keys(pos, box), GOLD["known_keys"])
    # test_peano.c:120-128 compares the table walk with the older bit-twiddling routine on single-bit coordinates;
    # here: keys of a complete 8^3 grid are a permutation of 0..511 and consecutive keys are face neighbours
    g = np.arange(8)
    x, y, z = np.meshgrid(g, g, g, indexing="ij")
    k = np.array([oracle.peano_key(int(a), int(b), int(c), 3) for a, b, c in zip(x.ravel(), y.ravel(), z.ravel())])
    assert sorted(k) == list(range(512))
    order = np.argsort(k)
    steps = np.abs(np.diff(np.stack([x.ravel()[order], y.ravel()[order], z.ravel()[order]], 1), axis=0)).sum(1)
    assert (steps == 1).all()


def test_oracle_keys_and_topleaf_equal_reference():
    pos, box = DS.random_positions()
    keys = oracle.peano_keys(pos, box)
    assert np.array_equal(keys, GOLD["random_keys"])
    top = DS.refined_toptree()
    assert len(top[0]) == int(GOLD["ntop"])
    assert np.array_equal(oracle.topleaf(keys, *top), GOLD["topleaf"])


def test_oracle_leaf_counts_and_balanced_assignment_equal_reference():
    pos, box = DS.random_positions()
    top = DS.refined_toptree()
    nleaf = int(top[3].max()) + 1
    leaf = oracle.topleaf(oracle.peano_keys(pos, box), *top)
    counts = oracle.leaf_counts(leaf, nleaf, flags=DS.garbage_flags(len(pos)))
    assert np.array_equal(counts, GOLD["leaf_counts"]) and counts.sum() == len(pos) - DS.garbage_flags(len(pos)).sum()
    got = np.concatenate([oracle.domain_assign_balanced(nt, cost) for nt, cost in DS.assign_cases()])
    assert np.array_equal(got, GOLD["assign_tasks"])
    # properties: tasks non-decreasing along the curve within a round, every task used, loads near the mean for flat costs
    for nt, cost in DS.assign_cases()[:40]:
        t = oracle.domain_assign_balanced(nt, cost)
        assert set(t) == set(range(nt))


def _check_toptree(make_tree, keys_of):
    for k, case in enumerate(DS.TOPTREE_CASES):
        fields, leaf, nl, sizes = DS.toptree_pipeline(make_tree, keys_of, case)
        assert np.array_equal(sizes, GOLD["toptree/%d/sizes" % k]), (k, sizes)
        for f in DS.TOPTREE_FIELDS:
            assert np.array_equal(fields[f], GOLD["toptree/%d/%s" % (k, f)]), (k, f)
        assert nl == int(GOLD["toptree/%d/nleaf" % k]) and np.array_equal(leaf, GOLD["toptree/%d/leaf" % k])
        assert nl >= case["ntopleaves"] // 2                                # a real refinement


def _subsample_keys(pos, box, sub):
    return oracle.peano_keys(pos[::sub][: len(pos) // sub], box)           # domain.c:1066-1074


def test_oracle_toptree_equals_reference():
    """Local refinement, truncation, two-rank merge, global refinement and leaf numbering, node for node."""
    _check_toptree(oracle.TopTree, _subsample_keys)


def test_product_host_toptree_equals_reference(b200):
    """b200_domain_toptree_* are host functions of libb200force.so (the tree has ~10^2-10^3 nodes): checked without a GPU;
    the subsample keys come from the oracle here and from k_domain_keys on the device (emulation / hardware tests)."""
    _check_toptree(b200.TopTree, _subsample_keys)


def test_oracle_exchange_plan_equals_reference():
    typ, fl, tl, tk, ntask = DS.exchange_case()
    total = np.zeros((ntask, 7), np.int64)
    for r in range(ntask):
        lst, togo, ng = oracle.exchange_plan(typ, fl, tl, tk, ntask, r)
        assert np.array_equal(lst, GOLD["xplan/%d/list" % r]) and np.array_equal(togo, GOLD["xplan/%d/togo" % r])
        assert ng == int(GOLD["xplan/%d/ngarbage" % r]) == int(fl.sum())
        assert togo[r].sum() == 0 and togo[:, 0].sum() == len(lst) and (togo[:, 1:].sum(1) == togo[:, 0]).all()
        total += togo
    # what all ranks send to task t is what t's own particles elsewhere would be: every live particle is kept or sent once
    live = fl == 0
    assert total[:, 0].sum() == sum(int((live & (tk[tl] != r)).sum()) for r in range(ntask))


def test_product_host_assignment_equals_reference(b200):
    """b200_domain_assign_balanced is host arithmetic inside libb200force.so: callable without a GPU."""
    got = np.concatenate([b200.domain_assign_balanced(nt, cost) for nt, cost in DS.assign_cases()])
    assert np.array_equal(got, GOLD["assign_tasks"])
    with pytest.raises(b200.B200Error):
        b200.domain_assign_balanced(8, np.ones(3, np.int64))            # fewer leaves than tasks


@pytest.mark.skipif(not R.available(), reason="oracle/_ref/libref_tree.so not built")
def test_oracle_equals_reference_live():
    r = R.load()
    if not hasattr(r.L, "ref_peano_keys"):
        pytest.skip("prebuilt libref_tree.so predates ref_peano_keys")
    pos, box = DS.random_positions(seed=17, n=5000, box=731.5)
    keys = oracle.peano_keys(pos, box)
    assert np.array_equal(keys, r.peano_keys(pos, box))
    top = DS.refined_toptree(seed=9, nrefine=25)
    assert np.array_equal(oracle.topleaf(keys, *top), r.topleaf(keys, *top))


@pytest.mark.gpu
def test_gpu_domain_keys(b200):
    try:
        e = b200.Engine(0)
    except Exception as ex:
        pytest.skip("no CUDA device (%s)" % ex)
    pos, box = DS.peano_test_positions()
    e.set_particles(pos, np.ones(len(pos), np.float32))
    assert np.array_equal(e.peano_keys(box), GOLD["known_keys"])
    pos, box = DS.random_positions()
    e.set_particles(pos, np.ones(len(pos), np.float32))
    assert np.array_equal(e.peano_keys(box), GOLD["random_keys"])
    top = DS.refined_toptree()
    assert np.array_equal(e.topleaf(*top), GOLD["topleaf"])
    counts = e.leaf_counts(int(top[3].max()) + 1)
    assert counts.sum() == len(pos)
    leaf = GOLD["topleaf"]
    assert np.array_equal(counts, np.bincount(leaf, minlength=len(counts)))
    # exchange plan: leaves dealt to 4 tasks by the balanced assignment
    tasks = b200.domain_assign_balanced(4, counts)
    for r in range(4):
        lst, togo, ng = e.exchange_plan(tasks, 4, r)
        olst, otogo, ong = oracle.exchange_plan(np.ones(len(pos), np.uint8), np.zeros(len(pos), np.uint8), leaf, tasks, 4, r)
        assert np.array_equal(lst, olst) and np.array_equal(togo, otogo) and ng == ong == 0
    e.close()


@pytest.mark.gpu
def test_gpu_domain_decompose_chain(b200):
    """domain.decompose on one GPU: device keys / lookup / counts / plan around the host top tree; checked against the
    oracle for the same particles (the tree from the oracle's stages, fed with the oracle's subsample keys)."""
    import importlib
    try:
        e = b200.Engine(0)
    except Exception as ex:
        pytest.skip("no CUDA device (%s)" % ex)
    dom = importlib.import_module("mp-gadget_b200.domain")
    box = 1000.0
    pos = DS.clustered(60000, box, 31)
    e.set_particles(pos, np.ones(len(pos), np.float32))
    d = dom.decompose(e, box, None, overdecomposition=16, subsample=16)
    keys = oracle.peano_keys(pos, box)
    O = oracle.TopTree(d["tree"].maxnodes)
    assert O.local(keys[::16][: len(pos) // 16]) == 0
    lim = int(O.tree["Count"][0]) // 16
    O.truncate(lim, lim); O.global_refine(lim, lim)
    for f in DS.TOPTREE_FIELDS:
        assert np.array_equal(O.tree[f], d["tree"].tree[f]), f
    nl, leaf = O.leaves()
    assert nl == d["nleaf"] and np.array_equal(leaf, d["leaf"])
    tl = oracle.topleaf(keys, *d["topnodes"])
    assert np.array_equal(tl, d["topleaf"]) and np.array_equal(d["counts"], np.bincount(tl, minlength=nl))
    assert (d["task_of_leaf"] == 0).all() and len(d["leaving"]) == 0 and d["ngarbage"] == 0
    e.close()
