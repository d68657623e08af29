"""Host orchestration of the domain decomposition over torch.distributed (one process per GPU):
domain_determine_global_toptree (libgadget/domain.c:1281-1340) and domain_balance (domain.c:482-502) on top of the
C-ABI stages (b200_domain_sample_keys on the device, b200_domain_toptree_* / b200_domain_assign_balanced on the host).
The exchanges are the reference's own -- two all-reduces, the pairwise tree merge of
domain_nonrecursively_combine_topTree (domain.c:1190-1278), one broadcast -- carried by torch.distributed (gloo on CPU
tests, NCCL on GPUs; the messages are a few kilobytes).  exchange() / decompose() / maintain() below restate
domain_exchange_once, domain_decompose_full (first policy) and domain_maintain for particle state held in torch tensors.
One departure from the reference: maintain() moves every particle whose top leaf changed task at once, where
domain.c:315-317 leaves gravitationally inactive dark matter in place until its next active step."""
import numpy as np
import torch

from . import TopTree, TOPNODE_DTYPE, B200Error, domain_assign_balanced


def _as_tensor(a):
    return torch.from_numpy(a.view(np.uint8).reshape(-1))


class TopTreeOverflow(B200Error):
    """The top tree ran out of nodes at some stage on some rank.  Raised on EVERY rank (the error flags are reduced, as the
    reference does with MPIU_Any, domain.c:1266-1272,1292,1334), so the caller can retry collectively with more nodes."""


def _any(flag, dist):
    """MPIU_Any: true on every rank when `flag` is set on one."""
    if dist is None:
        return bool(flag)
    t = torch.tensor([1 if flag else 0], dtype=torch.int64)
    dist.all_reduce(t)
    return int(t) > 0


def global_toptree(sample_keys, ntopleaves, dist=None, maxnodes=None):
    """sample_keys: this rank's subsample keys (Engine.sample_keys).  -> (TopTree, leaf number per node, nleaf), identical on
    every rank.  ntopleaves = policy->NTopLeaves (DomainOverDecompositionFactor * NTask * (attempt + 1), domain.c:369-371).
    A failure of any stage on any rank raises TopTreeOverflow on every rank: no rank is left waiting in a collective."""
    rank = dist.get_rank() if dist is not None else 0
    world = dist.get_world_size() if dist is not None else 1
    if maxnodes is None:
        maxnodes = 8 * max(len(sample_keys), 1) * world + 64 * ntopleaves
    T = TopTree(maxnodes)
    err = bool(T.local(sample_keys))                                                     # domain.c:1286-1296
    tot = torch.tensor([0 if err else int(T.tree["Count"][0]), 0 if err else int(T.tree["Cost"][0]), 1 if err else 0], dtype=torch.int64)
    if dist is not None:
        dist.all_reduce(tot)
    if int(tot[2]) > 0:
        raise TopTreeOverflow("top tree: local refinement failed on %d rank(s)" % int(tot[2]))
    countlimit, costlimit = int(tot[0]) // ntopleaves, int(tot[1]) // ntopleaves          # :1301-1302
    T.truncate(countlimit, costlimit)
    # domain_nonrecursively_combine_topTree: at separation sep the leaders of odd groups hand their tree to the leader of
    # the even group on their left and drop out, until rank 0 holds the merge of all.  A rank whose merge fails keeps
    # taking part (its partners are already in their send / recv); the flag is reduced after the loop (:1266-1272).
    alive = True
    sep = 1
    while sep < world:
        if alive and rank % sep == 0:
            if (rank // sep) % 2 == 0:
                src = rank + sep
                if src < world:
                    n = torch.zeros(1, dtype=torch.int64)
                    dist.recv(n, src=src)
                    other = TopTree(max(int(n), T.size.value, 1))
                    dist.recv(_as_tensor(other.nodes)[: int(n) * TOPNODE_DTYPE.itemsize], src=src)
                    other.size.value = int(n)
                    if not err and (T.size.value + int(n) > maxnodes or (int(n) > 0 and T.merge(other))):
                        err = True
            else:
                dist.send(torch.tensor([T.size.value], dtype=torch.int64), dst=rank - sep)
                dist.send(_as_tensor(T.nodes)[: T.size.value * TOPNODE_DTYPE.itemsize].clone(), dst=rank - sep)
                alive = False
        sep *= 2
    if _any(err, dist):
        raise TopTreeOverflow("top tree: out of nodes while merging")
    if dist is not None and world > 1:
        n = torch.tensor([T.size.value if rank == 0 else 0], dtype=torch.int64)
        dist.broadcast(n, src=0)
        if int(n) < 0 or int(n) >= maxnodes:                                              # :1318-1321, the same value on every rank
            raise TopTreeOverflow("top tree: merged size %d does not fit %d nodes" % (int(n), maxnodes))
        T.size.value = int(n)
        buf = _as_tensor(T.nodes)[: int(n) * TOPNODE_DTYPE.itemsize]
        dist.broadcast(buf, src=0)
    if _any(bool(T.global_refine(countlimit, costlimit)), dist):                          # :1334
        raise TopTreeOverflow("top tree: out of nodes in the global refinement")
    nleaf, leaf = T.leaves()
    return T, leaf, nleaf


def topnode_arrays(T, leaf):
    """(Daughter, StartKey, Shift, Leaf) of the tree: the arguments of b200_domain_set_topnodes / Engine.topleaf"""
    t = T.tree
    return t["Daughter"].astype(np.int32), t["StartKey"].astype(np.uint64), t["Shift"].astype(np.int32), leaf.astype(np.int32)


def balance(local_leaf_counts, dist=None, nseg_per_task=1):
    """domain_balance (domain.c:482-502): summed per-leaf particle counts -> task of every top leaf (leaves in key order)."""
    world = dist.get_world_size() if dist is not None else 1
    c = torch.from_numpy(np.ascontiguousarray(local_leaf_counts, np.int64).copy())
    if dist is not None:
        dist.all_reduce(c)
    return domain_assign_balanced(world, c.numpy(), nseg_per_task), c.numpy()


def exchange(tensors, leaving, target, dist=None):
    """domain_exchange_once (exchange.c:211-405) for particle state held as torch tensors (first dimension = particle):
    the particles `leaving` (indices, ascending: the exchange list of b200_domain_exchange_plan) go to the tasks `target`
    (one per leaving particle); everything else stays, in order, and what arrives is appended by source rank.  The
    counts go round first (MPI_Alltoall of toGo, exchange.c:530), then one variable-size all-to-all per tensor
    (NCCL all_to_all_single; pairwise sends on gloo).  Returns the new tensors (same keys)."""
    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    first = next(iter(tensors.values()))
    dev = first.device
    n = first.shape[0]
    leaving = torch.as_tensor(leaving, dtype=torch.int64, device=dev)
    target = torch.as_tensor(target, dtype=torch.int64, device=dev)
    if world == 1 or dist is None:
        if leaving.numel():
            raise B200Error("exchange: particles leave the only task")
        return dict(tensors)
    if target.numel() and (int(target.min()) < 0 or int(target.max()) >= world or bool((target == rank).any())):
        raise B200Error("exchange: bad target task")
    order = torch.argsort(target, stable=True)                 # by destination, ascending particle index inside one
    send_idx = leaving[order]
    togo = torch.bincount(target, minlength=world).to(torch.int64)
    toget = torch.zeros_like(togo)
    nccl = dist.get_backend() == "nccl"

    def pairwise(send_parts, recv_parts):
        ops = []
        for p in range(world):
            if p == rank:
                continue
            if send_parts[p].numel():
                ops.append(dist.P2POp(dist.isend, send_parts[p], p))
            if recv_parts[p].numel():
                ops.append(dist.P2POp(dist.irecv, recv_parts[p], p))
        if ops:
            for r in dist.batch_isend_irecv(ops):
                r.wait()

    if nccl:
        dist.all_to_all_single(toget, togo)
    else:
        sp = [togo[p:p + 1].clone() for p in range(world)]
        rp = [toget[p:p + 1] for p in range(world)]
        pairwise(sp, rp)
    send_counts = [int(c) for c in togo.tolist()]
    recv_counts = [int(c) for c in toget.tolist()]
    keep = torch.ones(n, dtype=torch.bool, device=dev)
    keep[leaving] = False
    out = {}
    for name, t in tensors.items():
        send = t[send_idx].contiguous()
        recv = torch.empty((sum(recv_counts),) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
        if nccl:
            dist.all_to_all_single(recv, send, output_split_sizes=recv_counts, input_split_sizes=send_counts)
        else:
            pairwise(list(torch.split(send, send_counts)), list(torch.split(recv, recv_counts)))
        out[name] = torch.cat([t[keep], recv])
    return out


def decompose(engine, box, dist=None, overdecomposition=4, subsample=256, attempt=0):
    """domain_decompose_full (domain.c:154-256) for the particles held by `engine`, first policy: the subsample keys,
    per-leaf counts, top-leaf lookup and exchange plan run on the device, the top tree and the leaf assignment on the
    host, the reductions over `dist`.  -> dict(tree, leaf, nleaf, topnodes, topleaf (int32[n]), task_of_leaf, counts,
    leaving (indices), target (task per leaving particle), togo, ngarbage)."""
    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    keys = engine.sample_keys(box, subsample)
    while True:                 # the retry loop of domain_decompose_full over its policies (domain.c:170-200): every rank
        try:                    # sees the same TopTreeOverflow, so every rank moves on to the next policy together
            ntopleaves = overdecomposition * world * (attempt + 1)         # domain_policies_init, domain.c:369-371
            T, leaf, nleaf = global_toptree(keys, ntopleaves, dist)
            break
        except TopTreeOverflow:
            attempt += 1
            if attempt >= 8:
                raise
    top = topnode_arrays(T, leaf)
    engine.peano_keys(box)
    topleaf = engine.topleaf(*top)
    task, counts = balance(engine.leaf_counts(nleaf), dist)
    leaving, togo, ngarbage = engine.exchange_plan(task, world, rank)
    return dict(tree=T, leaf=leaf, nleaf=nleaf, topnodes=top, topleaf=topleaf, task_of_leaf=task, counts=counts, leaving=leaving,
                target=task[topleaf[leaving]], togo=togo, ngarbage=ngarbage)


def maintain(engine, box, topnodes, task_of_leaf, dist=None):
    """domain_maintain (domain.c:282-347) after a drift on the device (b200_step_drift): the top tree and the leaf -> task
    table stay, every particle's top leaf is looked up again and those whose task changed form the exchange list.
    The reference additionally leaves gravitationally inactive dark matter where it is until its next active step
    (domain.c:315-317, an optimisation tied to its host tree); here every particle follows its leaf at once.
    -> dict(topleaf, leaving, target, togo, ngarbage)"""
    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    engine.peano_keys(box)
    topleaf = engine.topleaf(*topnodes)
    leaving, togo, ngarbage = engine.exchange_plan(task_of_leaf, world, rank)
    return dict(topleaf=topleaf, leaving=leaving, target=np.asarray(task_of_leaf)[topleaf[leaving]], togo=togo, ngarbage=ngarbage)
