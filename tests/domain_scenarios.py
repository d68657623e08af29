"""Inputs of the domain-key tests (Peano-Hilbert keys, top-leaf lookup), shared by the golden generator and the tests."""
import numpy as np

BITS = 21


def peano_test_positions():
    """tests/test_peano.c:107-118: the 4^3 integer lattice in a box of 4."""
    B = 4
    i = np.arange(B ** 3)
    return np.stack([i % B, (i // B) % B, (i // B // B) % B], 1).astype(np.float64), float(B)


def random_positions(seed=3, n=20000, box=25000.0):
    rng = np.random.default_rng(seed)
    pos = rng.random((n, 3)) * box
    pos[:6] = [[0, 0, 0], [box, box, box], [box * (1 - 1e-12), 0, box / 2], [box / 2, box / 2, box / 2], [1e-9, box - 1e-9, 3.0], [box / 3, box / 7, box / 11]]
    return pos, box


def refined_toptree(seed=4, nrefine=40):
    """A top tree as domain_decompose_full leaves it (domain.c): the root over the whole key range, nodes split into 8
    daughters stored consecutively, leaves numbered along the curve; which nodes are split is random here."""
    rng = np.random.default_rng(seed)
    daughter = [-1]; startkey = [0]; shift = [3 * BITS]
    leaves = [0]
    for _ in range(nrefine):
        t = leaves.pop(int(rng.integers(len(leaves))))
        if shift[t] < 6:
            leaves.append(t); continue
        daughter[t] = len(daughter)
        for j in range(8):
            daughter.append(-1); shift.append(shift[t] - 3); startkey.append(startkey[t] + (j << (shift[t] - 3)))
            leaves.append(len(daughter) - 1)
    order = sorted(leaves, key=lambda t: startkey[t])
    leaf = [-1] * len(daughter)
    for k, t in enumerate(order):
        leaf[t] = k
    return (np.array(daughter, np.int32), np.array(startkey, np.uint64), np.array(shift, np.int32), np.array(leaf, np.int32))


def garbage_flags(n, seed=6):
    rng = np.random.default_rng(seed)
    return (rng.random(n) < 0.02).astype(np.uint8)


def assign_cases(seed=3, ncase=120):
    """(ntask, cost[nleaf]) pairs for the balanced top-leaf assignment: flat, heavy-tailed, many empty leaves, equal costs,
    as few leaves as tasks."""
    rng = np.random.default_rng(seed)
    out = []
    for trial in range(ncase):
        ntask = int(rng.choice([1, 2, 3, 4, 8, 16]))
        nleaf = ntask if trial % 15 == 14 else int(rng.integers(ntask, 40 * ntask))
        kind = trial % 4
        if kind == 0:
            cost = rng.integers(0, 1000, nleaf)
        elif kind == 1:
            cost = (rng.pareto(1.2, nleaf) * 100).astype(np.int64)
        elif kind == 2:
            cost = np.where(rng.random(nleaf) < 0.3, 0, rng.integers(1, 50, nleaf))
        else:
            cost = np.full(nleaf, 17)
        cost = cost.astype(np.int64)
        if cost.sum() == 0:
            cost[0] = 1
        out.append((ntask, cost))
    return out


def clustered(n, box, seed):
    r = np.random.default_rng(seed)
    pos = r.random((n, 3)) * box
    pos[: n // 2] = box * r.random(3) + box / 20 * r.standard_normal((n // 2, 3))
    return np.mod(pos, box)


TOPTREE_CASES = [dict(n=(40000, 17000), box=1000.0, subsample=16, ntopleaves=32, seeds=(1, 2)),
                 dict(n=(9000, 30000), box=250.0, subsample=4, ntopleaves=64, seeds=(3, 4)),
                 dict(n=(25000, 25000), box=5000.0, subsample=64, ntopleaves=8, seeds=(5, 6))]
TOPTREE_FIELDS = ("StartKey", "Shift", "Daughter", "Parent", "Count", "Cost")


def toptree_pipeline(make_tree, sample_keys_of, case):
    """The two-rank top tree of domain_determine_global_toptree: each rank's local tree from its subsample, truncation
    with the limits from the summed root counts, merge of rank 1 into rank 0, global refinement, leaf numbering.
    make_tree(maxnodes) -> object with local/truncate/merge/global_refine/leaves/.tree; sample_keys_of(pos, box, sub) -> keys.
    Returns (tree fields dict, leaf array, nleaf, sizes after each stage)."""
    trees, sizes = [], []
    maxn = 8 * max(case["n"]) // case["subsample"] + 64
    for n, seed in zip(case["n"], case["seeds"]):
        T = make_tree(maxn)
        assert T.local(sample_keys_of(clustered(n, case["box"], seed), case["box"], case["subsample"])) == 0
        trees.append(T); sizes.append(len(T.tree))
    lim = int(trees[0].tree["Count"][0] + trees[1].tree["Count"][0]) // case["ntopleaves"]
    for T in trees:
        T.truncate(lim, lim); sizes.append(len(T.tree))
    A, B = trees
    assert A.merge(B) == 0; sizes.append(len(A.tree))
    assert A.global_refine(lim, lim) == 0; sizes.append(len(A.tree))
    nl, leaf = A.leaves()
    return {f: A.tree[f].copy() for f in TOPTREE_FIELDS}, leaf, nl, np.array(sizes, np.int64)


def exchange_case(seed=8, n=30000, nleaf=97, ntask=5):
    """Particles of all types with garbage, random top leaves, a leaf -> task table in curve order (contiguous runs)."""
    rng = np.random.default_rng(seed)
    typ = rng.choice([0, 1, 1, 1, 4, 5], n).astype(np.uint8)
    flags = (rng.random(n) < 0.03).astype(np.uint8)
    topleaf = rng.integers(0, nleaf, n).astype(np.int32)
    cuts = np.sort(rng.choice(np.arange(1, nleaf), ntask - 1, replace=False))
    task_of_leaf = np.searchsorted(cuts, np.arange(nleaf), side="right").astype(np.int32)
    return typ, flags, topleaf, task_of_leaf, ntask


def fof_cases():
    """(pos, ids, type, box, ll): clustered dark matter with gas mixed in (not a primary link type), shuffled IDs, a close
    pair across the periodic faces; a uniform box where most groups are single particles."""
    out = []
    for seed, n, box in ((2, 6000, 100.0), (3, 4000, 40.0)):
        rng = np.random.default_rng(seed)
        pos = rng.random((n, 3)) * box
        if seed == 2:
            pos[: n // 3] = 0.5 * box + 0.02 * box * rng.standard_normal((n // 3, 3))
            pos[n // 3: n // 2] = 0.98 * box + 0.015 * box * rng.standard_normal((n // 2 - n // 3, 3))     # a clump over the corner
        pos = np.mod(pos, box)
        ll = 0.2 * box / n ** (1 / 3)
        pos[-1] = [box - 0.2 * ll, 1.0, 1.0]; pos[-2] = [0.2 * ll, 1.0, 1.0]                                # linked through the face
        ids = rng.permutation(n).astype(np.int64) + 10
        typ = np.ones(n, np.uint8); typ[::7] = 0; typ[-2:] = 1
        out.append((pos, ids, typ, box, ll))
    return out


def fof_edge_cases():
    """(pos, ids, type, box, ll, mask, flags): a box so small that the grid has one cell (ll above a third of the box); two
    link types; garbage and swallowed particles among the primaries; a linking length that joins everything; one that joins
    nothing; groups along every face and corner of the box; a single particle."""
    out = []
    rng = np.random.default_rng(11)
    n = 300
    pos = rng.random((n, 3)) * 1.0
    ids = rng.permutation(n).astype(np.int64) + 1
    out.append((pos, ids, np.ones(n, np.uint8), 1.0, 0.4, 2, None))                       # nc < 3: every pair examined
    n = 3000; box = 50.0
    pos = rng.random((n, 3)) * box
    typ = rng.integers(0, 6, n).astype(np.uint8)
    ids = rng.permutation(n).astype(np.int64) + 5
    ll = 0.6 * box / n ** (1 / 3)
    out.append((pos, ids, typ, box, ll, (1 << 1) | (1 << 4), None))                       # dark matter and stars link
    flags = np.zeros(n, np.uint8); flags[::5] = 1; flags[3::11] = 2
    out.append((pos, ids, np.ones(n, np.uint8), box, ll, 2, flags))                       # garbage / swallowed take no part
    out.append((pos[:500], ids[:500], np.ones(500, np.uint8), box, 0.34 * box, 2, None))  # one cell again, everything joins
    out.append((pos, ids, np.ones(n, np.uint8), box, 1e-6, 2, None))                      # nothing joins (1024-cell cap)
    m = 12
    g = (np.arange(m) + 0.5) / m * box
    face = np.array([[x, y, z] for x in (0.01, box - 0.01) for y in g for z in g] + [[x, y, z] for y in (0.02, box - 0.02) for x in g for z in g]
                    + [[x, y, z] for z in (0.0, box - 0.03) for x in g for y in g])
    face = face + 0.3 * rng.standard_normal(face.shape)
    face = np.mod(face, box)
    out.append((face, rng.permutation(len(face)).astype(np.int64), np.ones(len(face), np.uint8), box, 1.2, 2, None))   # groups across faces, edges, corners
    out.append((np.array([[1.0, 2.0, 3.0]]), np.array([42], np.int64), np.ones(1, np.uint8), box, 1.0, 2, None))
    # a clump far denser than the linking length (hundreds of particles per grid cell) over a corner of the box, a thin
    # filament leaving it, and background particles: the clique cells' early exits and whole-cell hooks
    nd = 4000
    clump = np.mod(0.4 * rng.standard_normal((nd, 3)), box)
    fil = np.stack([np.linspace(1.0, 20.0, 400), np.full(400, 0.2), np.full(400, box - 0.2)], axis=1) + 0.02 * rng.standard_normal((400, 3))
    bg = rng.random((1500, 3)) * box
    pts = np.mod(np.concatenate([clump, fil, bg]), box)
    out.append((pts, rng.permutation(len(pts)).astype(np.int64) + 3, np.ones(len(pts), np.uint8), box, 0.3, 2, None))
    return out
