/* libgadget_shims.c -- the reference-signature side of the drop-in boundary.
 *
 * MP-Gadget has no plugin/FFI layer: libgadget modules call each other through
 * C headers (libgadget/Makefile:46-77).  A maintainer drops this file into
 * libgadget/ IN PLACE OF gravshort-tree.c (and, with -DB200_SHIM_GRAVPM, of
 * gravpm.c), adds -lb200force to the link line, and the rest of the code
 * (run.c:519-548, timestep.c:282-290, runtests.c) keeps calling
 *     grav_short_tree(), gravpm_force(), set_gravshort_tree_params(), ...
 * with unchanged signatures.  It compiles against the reference's own headers
 * (never copied here) and forwards to the C-ABI of include/b200force.h.
 *
 * Error convention: a non-zero b200 status becomes endrun(1, ...) = MPI_Abort
 * (libgadget/utils/endrun.c:138-153), as everywhere in the reference.
 *
 * This file is compiled and exercised by tests/test_dropin.py through
 * oracle/Makefile.ref (target libref_dropin.so): the reference's own driver
 * code + forcetree.c + treewalk.c call grav_short_tree() below, which runs on
 * the GPU, and the result is compared with the stock CPU path.
 */
#include <mpi.h>
#include <math.h>
#include <string.h>
#include <stdlib.h>

#include <libgadget/utils/endrun.h>
#include <libgadget/utils/mymalloc.h>
#include <libgadget/partmanager.h>
#include <libgadget/forcetree.h>
#include <libgadget/treewalk.h>
#include <libgadget/timestep.h>
#include <libgadget/gravity.h>
#include <libgadget/petapm.h>
#include <libgadget/walltime.h>
#include <libgadget/powerspectrum.h>
#include <libgadget/cosmology.h>

#include "../../include/b200force.h"

/* the context shared by all shim files (libgadget_shim_ctx.c) */
b200_ctx *b200_shim_context(void);
void b200_shim_topnodes_from_tree(const ForceTree *tree);
#define b200_shim_ctx b200_shim_context
#define B200_CK(call) do { if((call) != 0) endrun(1, "b200: %s\n", b200_last_error(b200_shim_context())); } while(0)

/* ---- replaces libgadget/gravshort-tree.c -------------------------------- */

static struct gravshort_tree_params TreeParams;
static double GravitySoftening;

double FORCE_SOFTENING(void) { return 2.8 * GravitySoftening; }          /* gravshort-tree.c:37-41 */

void gravshort_set_softenings(double MeanSeparation)                        /* :44-50 */
{
    GravitySoftening = TreeParams.FractionalGravitySoftening * MeanSeparation;
    message(0, "GravitySoftening = %g\n", GravitySoftening);
}
void set_gravshort_treepar(struct gravshort_tree_params tree_params) { TreeParams = tree_params; }   /* :53-56 */
struct gravshort_tree_params get_gravshort_treepar(void) { return TreeParams; }                      /* :58-61 */

void set_gravshort_tree_params(ParameterSet *ps)                            /* :64-78 */
{
    int ThisTask;
    MPI_Comm_rank(MPI_COMM_WORLD, &ThisTask);
    if(ThisTask == 0) {
        TreeParams.BHOpeningAngle = param_get_double(ps, "BHOpeningAngle");
        TreeParams.ErrTolForceAcc = param_get_double(ps, "ErrTolForceAcc");
        TreeParams.TreeUseBH = param_get_int(ps, "TreeUseBH");
        TreeParams.Rcut = param_get_double(ps, "TreeRcut");
        TreeParams.FractionalGravitySoftening = param_get_double(ps, "GravitySoftening");
        TreeParams.MaxBHOpeningAngle = param_get_double(ps, "MaxBHOpeningAngle");
    }
    MPI_Bcast(&TreeParams, sizeof(struct gravshort_tree_params), MPI_BYTE, 0, MPI_COMM_WORLD);
}

/* grav_short_tree (gravity.h:58, gravshort-tree.c:96-154).  The octree lives on
 * the device: it is rebuilt there from the same particle set the caller's
 * ForceTree was built from (tree->mask, act), which costs a few ms. */
void grav_short_tree(const ActiveParticles *act, PetaPM *pm, ForceTree *tree, MyFloat (*AccelStore)[3], double rho0, inttime_t Ti_Current)
{
    b200_ctx *ctx = b200_shim_ctx();
    if(!tree->moments_computed_flag)
        endrun(2, "Gravtree called before tree moments computed!\n");
    const int64_t n = PartManager->NumPart;
    b200_particle_layout lay;
    b200_default_particle_layout(&lay);
    if(sizeof(struct particle_data) != (size_t) lay.stride)
        endrun(2, "b200: struct particle_data is %lu bytes, the shim was built for %ld\n", sizeof(struct particle_data), (long) lay.stride);

    B200_CK(b200_pm_init(ctx, tree->BoxSize, pm->Asmth, pm->Nmesh, pm->G));
    B200_CK(b200_set_particles_aos(ctx, P, n, &lay));
    /* the same node set as force_tree_build: below the domain's top nodes (forcetree.c:654-687) */
    b200_shim_topnodes_from_tree(tree);
    B200_CK(b200_tree_build(ctx, tree->BoxSize, tree->mask, act->ActiveParticle, act->NumActiveParticle, -1, NULL));

    b200_gravshort_params par;
    memset(&par, 0, sizeof(par));
    par.ErrTolForceAcc = TreeParams.ErrTolForceAcc;
    par.BHOpeningAngle = TreeParams.BHOpeningAngle;
    par.MaxBHOpeningAngle = TreeParams.MaxBHOpeningAngle;
    par.TreeUseBH = TreeParams.TreeUseBH;
    par.Rcut = TreeParams.Rcut;
    par.GravitySoftening = GravitySoftening;
    par.rho0 = rho0;

    double *acc = (double *) mymalloc2("B200Accel", sizeof(double) * 3 * (n > 0 ? n : 1));
    double *pot = (double *) mymalloc2("B200Pot", sizeof(double) * (n > 0 ? n : 1));
    B200_CK(b200_grav_short_tree(ctx, &par, act->ActiveParticle, act->NumActiveParticle, acc, pot, NULL));

    const int64_t nq = act->ActiveParticle ? act->NumActiveParticle : n;
    #pragma omp parallel for
    for(int64_t q = 0; q < nq; q++) {
        const int64_t i = act->ActiveParticle ? act->ActiveParticle[q] : q;
        if(P[i].IsGarbage || P[i].Swallowed) continue;                 /* treewalk.c:234 */
        int k;
        if(AccelStore) for(k = 0; k < 3; k++) AccelStore[i][k] = acc[3 * i + k];      /* gravshort.h:51-53 */
        if(tree->full_particle_tree_flag) {                                           /* gravshort.h:55-66 */
            for(k = 0; k < 3; k++) P[i].FullTreeGravAccel[k] = acc[3 * i + k];
            P[i].Potential = pot[i];
        }
    }
    myfree(pot);
    myfree(acc);
    b200_timings t;
    if(b200_get_timings(ctx, &t) == 0) {
        walltime_add("/Tree/Build/Nodes", 1e-3 * t.tree_total);
        walltime_add("/Tree/WalkPrim", 1e-3 * t.walk);
    }
    /* TreeUseBH > 1: Barnes-Hut on the first step only (gravshort-tree.c:148-151) */
    if(TreeParams.TreeUseBH > 1)
        TreeParams.TreeUseBH = 0;
}

#ifdef B200_SHIM_GRAVPM
/* ---- replaces libgadget/gravpm.c; the power-spectrum sums come from the GPU pass and go
 * through the reference's own powerspectrum_sum / powerspectrum_save (powerspectrum.c). The
 * massive-neutrino linear-response branch (gravpm.c:304-326,415-437) is not supported. */
void gravpm_init_periodic(PetaPM *pm, double BoxSize, double Asmth, int Nmesh, double G)   /* gravpm.c:51-54 */
{
    pm->BoxSize = BoxSize; pm->Asmth = Asmth; pm->Nmesh = Nmesh; pm->G = G;
    pm->CellSize = BoxSize / Nmesh;
    B200_CK(b200_pm_init(b200_shim_ctx(), BoxSize, Asmth, Nmesh, G));
}

void gravpm_force(PetaPM *pm, DomainDecomp *ddecomp, Cosmology *CP, double Time, double UnitLength_in_cm, const char *PowerOutputDir, double TimeIC)
{
    b200_ctx *ctx = b200_shim_ctx();
    const int64_t n = PartManager->NumPart;
    b200_particle_layout lay;
    b200_default_particle_layout(&lay);
    B200_CK(b200_pm_init(ctx, pm->BoxSize, pm->Asmth, pm->Nmesh, pm->G));
    B200_CK(b200_set_particles_aos(ctx, P, n, &lay));
    double *g = (double *) mymalloc2("B200GravPM", sizeof(double) * 3 * (n > 0 ? n : 1));
    double *pot = (double *) mymalloc2("B200PMPot", sizeof(double) * (n > 0 ? n : 1));
    if(CP && CP->MassiveNuLinRespOn)
        endrun(1, "b200 gravpm_force(): the linear-response neutrino branch is not supported by this build\n");
    B200_CK(b200_pm_set_power(ctx, 1));
    B200_CK(b200_pm_force(ctx, g, pot));
    #pragma omp parallel for
    for(int64_t i = 0; i < n; i++) {
        int k;
        for(k = 0; k < 3; k++) P[i].GravPM[k] = g[3 * i + k];      /* gravpm.c:88-92,502-510 */
        P[i].Potential += pot[i];                                   /* gravpm.c:499-501 */
    }
    myfree(pot);
    myfree(g);
    /* gravpm.c:207 (allocation inside _prepare), :110-118 */
    powerspectrum_alloc(pm->ps, pm->Nmesh, 1, 0, pm->BoxSize * UnitLength_in_cm);
    B200_CK(b200_pm_get_power(ctx, pm->Nmesh, pm->ps->Power, pm->ps->kk, pm->ps->Nmodes, &pm->ps->Norm));
    powerspectrum_sum(pm->ps);
    powerspectrum_save(pm->ps, PowerOutputDir, "powerspectrum", Time, GrowthFactor(CP, Time, 1.0));
    powerspectrum_free(pm->ps);
    walltime_measure("/PMgrav/B200");
}
#endif
