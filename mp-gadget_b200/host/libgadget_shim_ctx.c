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
        int dev = 0;
        const char *e = getenv("B200_DEVICE");
        if(e) dev = atoi(e);
        if(b200_ctx_create(&ShimCtx, dev))
            endrun(1, "b200: cannot create a CUDA context on device %d (no CPU fallback in this build)\n", dev);
    }
    return ShimCtx;
}
