/* ref_domain_driver.c -- drives file-static routines of the reference's domain.c by including that file where it
 * lies (nothing is copied): the balanced assignment of top leaves to tasks (domain_assign_topleaves_balanced,
 * domain.c:610-755) for any task count.  TEST INFRASTRUCTURE ONLY; pins oracle_domain_assign_balanced.
 * The routine is pure computation; the one MPI call it makes, MPI_Comm_size, is answered by the stand-in header with
 * the task count set here. */
#include <libgadget/domain.c>

extern int ref_stub_ntask;

/* entry points of files that are not built (exchange.c, utils/mpsort.c need a real MPI; blackhole.c is sub-grid
 * physics): domain.c links to them, the routines driven here never reach them */
static void unreachable(const char *what) { endrun(1, "ref_domain_driver: %s reached\n", what); }
/* domain_exchange itself comes from exchange.c, included by ref_exchange_driver.c for its static planning routines */
void mpsort_mpi_impl(void *base, size_t nmemb, size_t size, void (*radix)(const void *ptr, void *radix, void *arg), size_t rsize, void *arg,
                     MPI_Comm comm, const int line, const char *file) { unreachable("mpsort_mpi"); }
int blackhole_dynfric_treemask(void) { return 0; }

/* leaves in key order: startkey[nleaf], cost[nleaf] -> task[nleaf] (TopLeaves[].Task in the final leaf order) and
 * order[nleaf] (which input leaf ended up at each position).  Returns 0. */
int ref_domain_assign(int ntask, int nleaf, const uint64_t *startkey, const int64_t *cost, int nseg_per_task, int *task_out, int *order_out)
{
    DomainDecomp d;
    memset(&d, 0, sizeof(d));
    d.DomainComm = MPI_COMM_WORLD;
    d.NTopLeaves = nleaf; d.NTopNodes = nleaf;
    d.TopNodes = (struct topnode_data *) mymalloc("TopNodes", sizeof(struct topnode_data) * nleaf);
    d.TopLeaves = (struct topleaf_data *) mymalloc("TopLeaves", sizeof(struct topleaf_data) * (nleaf + 1));
    for(int i = 0; i < nleaf; i++) {
        memset(&d.TopNodes[i], 0, sizeof(d.TopNodes[i]));
        d.TopNodes[i].StartKey = startkey[i]; d.TopNodes[i].Daughter = -1; d.TopNodes[i].Leaf = i;
        d.TopLeaves[i].topnode = i; d.TopLeaves[i].Task = -1;
    }
    int64_t *c = (int64_t *) mymalloc("cost", sizeof(int64_t) * nleaf);
    memcpy(c, cost, sizeof(int64_t) * nleaf);
    ref_stub_ntask = ntask;
    domain_assign_topleaves_balanced(&d, c, nseg_per_task);
    ref_stub_ntask = 1;
    for(int i = 0; i < nleaf; i++) { task_out[i] = d.TopLeaves[i].Task; order_out[i] = d.TopLeaves[i].topnode; }
    myfree(c); myfree(d.TopLeaves); myfree(d.TopNodes);
    return 0;
}

/* domain_compute_costs (domain.c:1396-1451) over a top tree given as arrays: TopLeafCount[nleaf] of the particles
 * (garbage skipped), leaf = domain_get_topleaf(PEANO(Pos)).  flags bit 0 = IsGarbage. */
int ref_domain_counts(int64_t n, const double *pos, const unsigned char *flags, double BoxSize, int ntop, const int *daughter,
                      const uint64_t *startkey, const int *shift, const int *leaf, int nleaf, int64_t *counts_out)
{
    particle_alloc_memory(PartManager, BoxSize, n);
    PartManager->NumPart = n;
    for(int64_t i = 0; i < n; i++) {
        memset(&P[i], 0, sizeof(P[i]));
        for(int k = 0; k < 3; k++) P[i].Pos[k] = pos[3 * i + k];
        P[i].IsGarbage = flags ? (flags[i] & 1) : 0;
    }
    DomainDecomp d;
    memset(&d, 0, sizeof(d));
    d.DomainComm = MPI_COMM_WORLD;
    d.NTopLeaves = nleaf; d.NTopNodes = ntop;
    d.TopNodes = (struct topnode_data *) mymalloc("TopNodes", sizeof(struct topnode_data) * ntop);
    for(int t = 0; t < ntop; t++) {
        memset(&d.TopNodes[t], 0, sizeof(d.TopNodes[t]));
        d.TopNodes[t].Daughter = daughter[t]; d.TopNodes[t].StartKey = startkey[t]; d.TopNodes[t].Shift = shift[t]; d.TopNodes[t].Leaf = leaf[t];
    }
    domain_compute_costs(&d, NULL, counts_out);
    myfree(d.TopNodes);
    myfree(P);
    return 0;
}

/* ---- the top tree (domain.c:826-1395), stage by stage on struct local_topnode_data arrays (domain.c:60-70: StartKey,
 * Shift, Daughter, Parent, Count, Cost; 40 bytes) handed in and out as raw memory ---- */
int ref_toptree_node_size(void) { return (int) sizeof(struct local_topnode_data); }

/* domain_check_for_local_refine_subsample (domain.c:1006-1187): skeleton from every `subsample`-th particle, then counts */
int ref_toptree_local(int64_t n, const double *pos, const unsigned char *flags, double BoxSize, int subsample, int presort, int maxnodes,
                      void *tree, int *size)
{
    particle_alloc_memory(PartManager, BoxSize, n);
    PartManager->NumPart = n;
    for(int64_t i = 0; i < n; i++) {
        memset(&P[i], 0, sizeof(P[i]));
        for(int k = 0; k < 3; k++) P[i].Pos[k] = pos[3 * i + k];
        P[i].IsGarbage = flags ? (flags[i] & 1) : 0;
    }
    DomainDecompositionPolicy pol;
    pol.SubSampleDistance = subsample; pol.PreSort = presort; pol.NTopLeaves = 0;
    domain_params.DomainUseGlobalSorting = 0;
    memset(tree, 0, sizeof(struct local_topnode_data) * maxnodes);
    const int rc = domain_check_for_local_refine_subsample(&pol, (struct local_topnode_data *) tree, size, maxnodes, MPI_COMM_WORLD);
    myfree(P);
    return rc;
}
void ref_toptree_truncate(void *tree, int *size, int64_t countlimit, int64_t costlimit)
{
    domain_toptree_truncate((struct local_topnode_data *) tree, size, countlimit, costlimit);
}
void ref_toptree_merge(void *treeA, int *sizeA, void *treeB, int maxnodes)
{
    domain_toptree_merge((struct local_topnode_data *) treeA, (struct local_topnode_data *) treeB, 0, 0, sizeA, maxnodes);
}
int ref_toptree_global_refine(void *tree, int *size, int maxnodes, int64_t countlimit, int64_t costlimit)
{
    return domain_global_refine((struct local_topnode_data *) tree, size, maxnodes, countlimit, costlimit);
}
/* domain_attempt_decompose's copy + domain_create_topleaves (domain.c:445-459,810-824): Leaf of every node (-1 inside) */
int ref_toptree_leaves(const void *tree, int size, int *leaf_out)
{
    const struct local_topnode_data *t = (const struct local_topnode_data *) tree;
    DomainDecomp d;
    memset(&d, 0, sizeof(d));
    d.TopNodes = (struct topnode_data *) mymalloc("TopNodes", sizeof(struct topnode_data) * size);
    d.TopLeaves = (struct topleaf_data *) mymalloc("TopLeaves", sizeof(struct topleaf_data) * (size + 1));
    for(int i = 0; i < size; i++) {
        d.TopNodes[i].StartKey = t[i].StartKey; d.TopNodes[i].Shift = t[i].Shift; d.TopNodes[i].Daughter = t[i].Daughter; d.TopNodes[i].Leaf = -1;
    }
    d.NTopNodes = size; d.NTopLeaves = 0;
    domain_create_topleaves(&d, 0, &d.NTopLeaves);
    for(int i = 0; i < size; i++) leaf_out[i] = d.TopNodes[i].Leaf;
    const int nl = d.NTopLeaves;
    myfree(d.TopLeaves); myfree(d.TopNodes);
    return nl;
}
