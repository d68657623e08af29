/* ref_exchange_driver.c -- drives the file-static planning routines of the reference's exchange.c by including that
 * file where it lies (nothing is copied): domain_build_exchange_list (exchange.c:408-444) and domain_build_plan
 * (exchange.c:505-566) for any task count and rank.  TEST INFRASTRUCTURE ONLY; pins oracle_exchange_plan.
 * Both routines are pure computation up to an MPI_Alltoall of the counts, which the stand-in header turns into a copy. */
#include <libgadget/exchange.c>

extern int ref_stub_ntask, ref_stub_thistask;

static const int *plan_task_of_leaf;
static int plan_nleaf;
static int plan_layout(int n, const void *userdata)          /* domain_layoutfunc, domain.c:794-803 */
{
    const int topleaf = PartManager->Base[n].TopLeaf;
    if(topleaf < 0 || topleaf >= plan_nleaf) endrun(6, "Invalid topleaf %d\n", topleaf);
    return plan_task_of_leaf[topleaf];
}

/* type[n], flags[n] (bit 0 garbage), topleaf[n]; task_of_leaf[nleaf].  Out: the exchange list (particle indices leaving
 * this rank, ascending), togo[ntask][7] = {base, slots[0..5]} per target task.  Returns the list length; *ngarbage. */
int64_t ref_exchange_plan(int64_t n, const unsigned char *type, const unsigned char *flags, const int *topleaf, int nleaf,
                          const int *task_of_leaf, int ntask, int thistask, int *list_out, int64_t *togo, int64_t *ngarbage)
{
    particle_alloc_memory(PartManager, 1.0, n);
    PartManager->NumPart = n;
    for(int64_t i = 0; i < n; i++) {
        memset(&P[i], 0, sizeof(P[i]));
        P[i].Type = type[i]; P[i].IsGarbage = flags ? (flags[i] & 1) : 0; P[i].TopLeaf = topleaf[i];
    }
    plan_task_of_leaf = task_of_leaf; plan_nleaf = nleaf;
    ref_stub_ntask = ntask; ref_stub_thistask = thistask;
    ExchangePlan plan = domain_init_exchangeplan(MPI_COMM_WORLD);
    domain_build_exchange_list(plan_layout, NULL, &plan, PartManager, SlotsManager, MPI_COMM_WORLD);
    plan.last = plan.nexchange;
    domain_build_plan(0, plan_layout, NULL, &plan, PartManager, MPI_COMM_WORLD);
    for(size_t q = 0; q < plan.nexchange; q++) list_out[q] = plan.ExchangeList[q];
    for(int t = 0; t < ntask; t++) {
        togo[7 * t] = plan.toGo[t].base;
        for(int k = 0; k < 6; k++) togo[7 * t + 1 + k] = plan.toGo[t].slots[k];
    }
    *ngarbage = plan.ngarbage;
    const int64_t nex = plan.nexchange;
    myfree(plan.layouts);
    myfree(plan.ExchangeList);
    domain_free_exchangeplan(&plan);
    myfree(P);
    ref_stub_ntask = 1; ref_stub_thistask = 0;
    return nex;
}
