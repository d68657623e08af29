/* libgadget_shim_ctx.c -- the one engine context all reference-signature shims
 * (libgadget_shims.c, libgadget_sph_shims.c, libgadget_forcetree_shims.c) share:
 * one process per GPU, device chosen by B200_DEVICE (the local MPI rank). */
#include <mpi.h>
#include <stdlib.h>
#include <libgadget/utils/endrun.h>
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
