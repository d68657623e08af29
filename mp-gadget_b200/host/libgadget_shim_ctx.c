/* libgadget_shim_ctx.c -- the one engine context all reference-signature shims
 * (libgadget_shims.c, libgadget_sph_shims.c, libgadget_forcetree_shims.c) share:
 * one process per GPU, device chosen by B200_DEVICE (the local MPI rank). */
#include <mpi.h>
#include <stdlib.h>
#include <string.h>
#include <libgadget/utils/endrun.h>
#include <libgadget/utils/peano.h>
#include <libgadget/domain.h>
#include <libgadget/forcetree.h>
#include "../../include/b200force.h"

static b200_ctx *ShimCtx;

b200_ctx *b200_shim_context(void)
{
    if(!ShimCtx) {
        /* The reference's calls are MPI-collective: grav_short_tree / density / hydro_force export particles to the ranks
         * whose top leaves they must open (treewalk.c:325-371,399-793), force_tree_build hangs pseudo particles below remote
         * top leaves (forcetree.c:905-910), gravpm_force exchanges mesh regions (petapm.c:584-885).  These shims hold one
         * rank's particles only and would silently drop every cross-domain interaction, so they refuse to run in a
         * multi-rank job; the multi-GPU path is b200_sharded_force_step (include/b200force.h). */
        int NTask = 1;
        MPI_Comm_size(MPI_COMM_WORLD, &NTask);
        if(NTask > 1)
            endrun(0, "b200: the drop-in shims are single-rank (NTask = %d): cross-domain exchanges are not wired behind the "
                      "reference entry points; use b200_sharded_force_step for multi-GPU runs\n", NTask);
        int dev = 0;
        const char *e = getenv("B200_DEVICE");
        if(e) dev = atoi(e);
        if(b200_ctx_create(&ShimCtx, dev))
            endrun(1, "b200: cannot create a CUDA context on device %d (no CPU fallback in this build)\n", dev);
    }
    return ShimCtx;
}

/* ---- the domain's top tree for b200_tree_build(toplevel_depth = -1) --------------------------------------------------
 * force_tree_build hangs the particle tree below the domain's top nodes (force_tree_create_topnodes, forcetree.c:654-687:
 * every top node with daughters becomes a node with all eight children, empty or not), so the engine needs their shape to
 * build the same node set. */
static int ShimTopNodes;         /* 1 once a top tree has been handed to the engine */

/* From the DomainDecomp itself (the GPU-only tree mode: force_tree_full(tree, ddecomp, ...) is a shim). */
void b200_shim_topnodes_from_domain(const DomainDecomp *ddecomp)
{
    const int n = ddecomp->NTopNodes;
    int32_t *dau = (int32_t *) malloc(sizeof(int32_t) * 3 * (size_t) n);
    uint64_t *key = (uint64_t *) malloc(sizeof(uint64_t) * (size_t) n);
    int32_t *shift = dau + n, *leaf = dau + 2 * n;
    for(int t = 0; t < n; t++) {
        dau[t] = ddecomp->TopNodes[t].Daughter; key[t] = ddecomp->TopNodes[t].StartKey;
        shift[t] = ddecomp->TopNodes[t].Shift; leaf[t] = ddecomp->TopNodes[t].Daughter < 0 ? ddecomp->TopNodes[t].Leaf : -1;
    }
    if(b200_domain_set_topnodes(b200_shim_context(), n, dau, key, shift, leaf))
        endrun(1, "b200: %s\n", b200_last_error(b200_shim_context()));
    free(dau); free(key);
    ShimTopNodes = 1;
}

/* From a host ForceTree built by the reference's own forcetree.c (only grav_short_tree / density / hydro_force swapped):
 * grav_short_tree gets no DomainDecomp, but the tree's top-level nodes spell out the same shape.  The table is
 * numbered afresh (breadth first, daughters in curve order), which is all the engine needs. */
void b200_shim_topnodes_from_tree(const ForceTree *tree)
{
    if(!tree->Nodes_base || tree->numnodes <= 0) {
        if(!ShimTopNodes) endrun(1, "b200: no host tree and no domain top tree to build the device tree from\n");
        return;                 /* GPU-only tree mode: force_tree_full's shim has handed over ddecomp->TopNodes */
    }
    int cap = 1024, n = 1, head = 0;
    int32_t *dau = (int32_t *) malloc(sizeof(int32_t) * cap), *shift = (int32_t *) malloc(sizeof(int32_t) * cap), *leaf = (int32_t *) malloc(sizeof(int32_t) * cap);
    uint64_t *key = (uint64_t *) malloc(sizeof(uint64_t) * cap);
    int *node = (int *) malloc(sizeof(int) * cap), (*cell)[4] = (int (*)[4]) malloc(sizeof(int) * 4 * cap);     /* x, y, z, level */
    dau[0] = -1; shift[0] = 3 * BITS_PER_DIMENSION; key[0] = 0; node[0] = (int) tree->firstnode;
    cell[0][0] = cell[0][1] = cell[0][2] = cell[0][3] = 0;
    int nleaf = 0;
    for(head = 0; head < n; head++) {
        const struct NODE *nd = &tree->Nodes[node[head]];
        if(!nd->f.InternalTopLevel) { leaf[head] = nleaf++; continue; }
        leaf[head] = -1;
        if(n + 8 > cap) {
            cap *= 2;
            dau = (int32_t *) realloc(dau, sizeof(int32_t) * cap); shift = (int32_t *) realloc(shift, sizeof(int32_t) * cap);
            leaf = (int32_t *) realloc(leaf, sizeof(int32_t) * cap); key = (uint64_t *) realloc(key, sizeof(uint64_t) * cap);
            node = (int *) realloc(node, sizeof(int) * cap); cell = (int (*)[4]) realloc(cell, sizeof(int) * 4 * cap);
        }
        dau[head] = n;
        const int bits = cell[head][3] + 1;
        for(int count = 0; count < 8; count++) {        /* forcetree.c:882-886: count = i + 2 j + 4 k */
            const int x = 2 * cell[head][0] + (count & 1), y = 2 * cell[head][1] + ((count >> 1) & 1), z = 2 * cell[head][2] + (count >> 2);
            const int sub = (int) (7 & peano_hilbert_key(x, y, z, bits));
            const int t = n + sub;
            dau[t] = -1; shift[t] = shift[head] - 3; key[t] = key[head] + ((uint64_t) sub << shift[t]);
            node[t] = nd->s.suns[count];
            cell[t][0] = x; cell[t][1] = y; cell[t][2] = z; cell[t][3] = bits;
        }
        n += 8;
    }
    if(b200_domain_set_topnodes(b200_shim_context(), n, dau, key, shift, leaf))
        endrun(1, "b200: %s\n", b200_last_error(b200_shim_context()));
    free(dau); free(shift); free(leaf); free(key); free(node); free(cell);
    ShimTopNodes = 1;
}
