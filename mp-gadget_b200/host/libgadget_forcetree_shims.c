/* libgadget_forcetree_shims.c -- GPU-only tree mode.
 *
 * Drop into libgadget/ IN PLACE OF forcetree.c when every consumer of the ForceTree is
 * one of the b200 shims (grav_short_tree, gravpm_force, density, hydro_force): the
 * octree then only ever exists in HBM (b200_tree_build, a few ms) and the host never
 * spends time in force_tree_build / force_tree_calc_moments (forcetree.c:196-270,
 * 727-1143, the reference's second-largest cost after the walk itself).  The entry
 * points keep the reference signatures (libgadget/forcetree.h:117-169) and record in the
 * caller's ForceTree what the consumers read: mask, BoxSize, the particle set and the
 * flags.  The device tree itself is (re)built by the consumer from exactly that
 * description, so it can never be stale.
 *
 * Entry points that hand out HOST node data (force_get_father, update_tree_hmax_father,
 * force_tree_top_build, ...; used by domain.c, blackhole.c, winds.c, fof.c and
 * set_init_hsml) cannot be served from a device-only tree: they end in endrun().  Keep
 * the reference's forcetree.c if such consumers are active.
 */
#include <mpi.h>
#include <string.h>
#include <libgadget/utils/endrun.h>
#include <libgadget/partmanager.h>
#include <libgadget/domain.h>
#include <libgadget/forcetree.h>
#include <libgadget/timestep.h>
#include <libgadget/walltime.h>

void b200_shim_topnodes_from_domain(const DomainDecomp *ddecomp);
static double TreeAllocFactor;
void init_forcetree_params(const double treeallocfactor) { TreeAllocFactor = treeallocfactor; }     /* forcetree.c:30-35 */

int force_tree_allocated(const ForceTree *tree) { return tree->tree_allocated_flag; }

static ForceTree describe_tree(int mask, DomainDecomp *ddecomp, const ActiveParticles *act, int moments)
{
    ForceTree tree;
    memset(&tree, 0, sizeof(tree));
    int64_t np = 0;
    const int64_t nq = act->ActiveParticle ? act->NumActiveParticle : PartManager->NumPart;
    #pragma omp parallel for reduction(+: np)
    for(int64_t q = 0; q < nq; q++) {
        const int64_t i = act->ActiveParticle ? act->ActiveParticle[q] : q;
        if(P[i].IsGarbage || P[i].Swallowed) continue;
        if(!((1 << P[i].Type) & mask)) continue;
        np++;
    }
    tree.tree_allocated_flag = 1;
    tree.mask = mask;
    tree.BoxSize = PartManager->BoxSize;
    tree.NumParticles = np;
    tree.firstnode = PartManager->MaxPart;
    tree.lastnode = tree.firstnode;
    tree.numnodes = 0;                       /* no host nodes */
    b200_shim_topnodes_from_domain(ddecomp);             /* the device tree is built below these (b200_tree_build, toplevel_depth = -1) */
    tree.TopLeaves = ddecomp->TopLeaves;
    tree.NTopLeaves = ddecomp->NTopLeaves;
    MPI_Comm_rank(MPI_COMM_WORLD, &tree.ThisTask);
    tree.moments_computed_flag = moments;
    tree.hmax_computed_flag = 0;
    return tree;
}

void force_tree_full(ForceTree *tree, DomainDecomp *ddecomp, const int HybridNuTracer, const char *EmergencyOutputDir)
{
    ActiveParticles act = init_empty_active_particles(PartManager);
    const int mask = HybridNuTracer ? GASMASK + DMMASK + STARMASK + BHMASK : ALLMASK;      /* forcetree.c:118-122 */
    *tree = describe_tree(mask, ddecomp, &act, 1);
    tree->full_particle_tree_flag = 1;
}

void force_tree_active_moments(ForceTree *tree, DomainDecomp *ddecomp, const ActiveParticles *act, const int HybridNuTracer,
                               const int alloc_father, const char *EmergencyOutputDir)
{
    const int mask = HybridNuTracer ? GASMASK + DMMASK + STARMASK + BHMASK : ALLMASK;
    *tree = describe_tree(mask, ddecomp, act, 1);
    if(!act->ActiveParticle) tree->full_particle_tree_flag = 1;                              /* forcetree.c:146-148 */
}

void force_tree_rebuild_mask(ForceTree *tree, DomainDecomp *ddecomp, int mask, const char *EmergencyOutputDir)
{
    ActiveParticles act = init_empty_active_particles(PartManager);
    *tree = describe_tree(mask, ddecomp, &act, 0);
    if(mask == ALLMASK) tree->full_particle_tree_flag = 1;                                   /* forcetree.c:164-165 */
}

void force_tree_calc_moments(ForceTree *tree, DomainDecomp *ddecomp)                        /* forcetree.c:170-183 */
{
    tree->moments_computed_flag = 1;
    tree->hmax_computed_flag = 1;
}

void force_tree_free(ForceTree *tree) { tree->tree_allocated_flag = 0; }

/* ---- host-node accessors: not available without a host tree --------------------- */
#define NO_HOST_TREE(what) endrun(1, "b200: " what " needs the host octree; link the reference's forcetree.c instead of libgadget_forcetree_shims.c\n")
int force_get_father(int no, const ForceTree *tt) { NO_HOST_TREE("force_get_father"); return -1; }
void update_tree_hmax_father(const ForceTree *const tree, const int p_i, const double Pos[3], const double Hsml) { NO_HOST_TREE("update_tree_hmax_father"); }
void force_update_hmax(ActiveParticles *act, ForceTree *tt, DomainDecomp *ddecomp) { NO_HOST_TREE("force_update_hmax"); }
int force_tree_find_topnode(const double *const pos, const ForceTree *const tree) { NO_HOST_TREE("force_tree_find_topnode"); return -1; }
ForceTree force_tree_top_build(DomainDecomp *ddecomp, const int alloc_high) { ForceTree t; memset(&t, 0, sizeof(t)); NO_HOST_TREE("force_tree_top_build"); return t; }
