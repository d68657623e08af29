/* libgadget_sph_shims.c -- reference-signature side of the drop-in boundary for
 * the SPH passes.
 *
 * A maintainer drops this file into libgadget/ IN PLACE OF density.c and hydra.c
 * (libgadget/Makefile:46-77), adds -lb200force, and run.c:466-495 / init.c:431,511
 * keep calling
 *     density(), hydro_force(), set_density_params(), SPH_EntVarPred(), ...
 * with unchanged signatures (libgadget/density.h:40-71, hydra.h:10-17).  It is
 * compiled against the reference's own headers (never copied here) and forwards to
 * the C-ABI of include/b200force.h.  densitykernel.c stays the reference's.
 *
 * Scope of this build: gas targets on any mix of time bins and any
 * active set; no black-hole density targets and no decoupled wind particles -- those end in
 * endrun(): there is no CPU fallback.
 *
 * Exercised by tests/test_dropin.py through oracle/Makefile.ref (target
 * libref_dropin_sph.so): the reference's fixture code of ref_driver.c calls
 * density() / hydro_force() below, which run on the GPU.
 */
#include <mpi.h>
#include <math.h>
#include <string.h>
#include <stdlib.h>

#include <libgadget/utils/endrun.h>
#include <libgadget/utils/mymalloc.h>
#include <libgadget/physconst.h>
#include <libgadget/partmanager.h>
#include <libgadget/slotsmanager.h>
#include <libgadget/forcetree.h>
#include <libgadget/timestep.h>
#include <libgadget/timefac.h>
#include <libgadget/timebinmgr.h>
#include <libgadget/cosmology.h>
#include <libgadget/density.h>
#include <libgadget/hydra.h>
#include <libgadget/gravity.h>
#include <libgadget/walltime.h>
#include <libgadget/winds.h>

#include "../../include/b200force.h"

/* the context shared by all shim files (libgadget_shim_ctx.c) */
b200_ctx *b200_shim_context(void);
void b200_shim_topnodes_from_tree(const ForceTree *tree);
#define sph_ctx b200_shim_context
#define B200_CK(call) do { if((call) != 0) endrun(1, "b200: %s\n", b200_last_error(b200_shim_context())); } while(0)

/* ---- module parameters: the setters of density.c:21-66 and hydra.c:26-54 ---- */
static struct density_params DensityParams;
static struct { int DensityIndependentSphOn; double DensityContrastLimit, ArtBulkViscConst; } HydroParams;

void set_densitypar(struct density_params dp) { DensityParams = dp; }

void set_density_params(ParameterSet *ps)
{
    int ThisTask;
    MPI_Comm_rank(MPI_COMM_WORLD, &ThisTask);
    if(ThisTask == 0) {
        DensityParams.DensityKernelType = (enum DensityKernelType) param_get_enum(ps, "DensityKernelType");
        DensityParams.MaxNumNgbDeviation = param_get_double(ps, "MaxNumNgbDeviation");
        DensityParams.DensityResolutionEta = param_get_double(ps, "DensityResolutionEta");
        DensityParams.MinGasHsmlFractional = param_get_double(ps, "MinGasHsmlFractional");
        DensityParams.BlackHoleNgbFactor = param_get_double(ps, "BlackHoleNgbFactor");
        DensityParams.BlackHoleMaxAccretionRadius = param_get_double(ps, "BlackHoleMaxAccretionRadius");
        message(1, "b200 SPH: kernel type %d, eta %g = %g neighbours\n", (int) DensityParams.DensityKernelType,
                DensityParams.DensityResolutionEta, GetNumNgb(DensityParams.DensityKernelType));
    }
    MPI_Bcast(&DensityParams, sizeof(struct density_params), MPI_BYTE, 0, MPI_COMM_WORLD);
}

double GetNumNgb(enum DensityKernelType KernelType)
{
    DensityKernel kernel;
    density_kernel_init(&kernel, 1.0, KernelType);
    return density_kernel_desnumngb(&kernel, DensityParams.DensityResolutionEta);
}
enum DensityKernelType GetDensityKernelType(void) { return DensityParams.DensityKernelType; }

void set_hydro_params(ParameterSet *ps)
{
    int ThisTask;
    MPI_Comm_rank(MPI_COMM_WORLD, &ThisTask);
    if(ThisTask == 0) {
        HydroParams.ArtBulkViscConst = param_get_double(ps, "ArtBulkViscConst");
        HydroParams.DensityContrastLimit = param_get_double(ps, "DensityContrastLimit");
        HydroParams.DensityIndependentSphOn = param_get_int(ps, "DensityIndependentSphOn");
    }
    MPI_Bcast(&HydroParams, sizeof(HydroParams), MPI_BYTE, 0, MPI_COMM_WORLD);
}
int DensityIndependentSphOn(void) { return HydroParams.DensityIndependentSphOn; }

/* ---- predictors other modules call (density.h:55-66) ------------------------ */
MyFloat SPH_EntVarPred(const int p_i, const DriftKickTimes *times)             /* density.c:69-85 */
{
    const int bin = P[p_i].TimeBinHydro;
    const struct sph_particle_data *s = &SphP[P[p_i].PI];
    const double dloga = dloga_from_dti(times->Ti_Current - times->Ti_kick[bin], times->Ti_Current);
    double e = s->Entropy + s->DtEntropy * dloga;
    if(e < 0.05 * s->Entropy) e = 0.05 * s->Entropy;
    if(e <= 0) return 0;
    return exp(1. / GAMMA * log(e));
}

void SPH_VelPred(int i, MyFloat *VelPred, const struct kick_factor_data *kf)   /* density.c:91-100 */
{
    for(int j = 0; j < 3; j++)
        VelPred[j] = P[i].Vel[j] + kf->gravkicks[P[i].TimeBinGravity] * P[i].FullTreeGravAccel[j]
                   + P[i].GravPM[j] * kf->FgravkickB + kf->hydrokicks[P[i].TimeBinHydro] * SPHP(i).HydroAccel[j];
}
void DM_VelPred(int i, MyFloat *VelPred, const struct kick_factor_data *kf)    /* density.c:106-111 */
{
    for(int j = 0; j < 3; j++)
        VelPred[j] = P[i].Vel[j] + kf->gravkicks[P[i].TimeBinGravity] * P[i].FullTreeGravAccel[j] + P[i].GravPM[j] * kf->FgravkickB;
}
void init_kick_factor_data(struct kick_factor_data *kf, const DriftKickTimes *const times, Cosmology *CP)   /* density.c:114-132 */
{
    kf->FgravkickB = get_exact_gravkick_factor(CP, times->PM_kick, times->Ti_Current);
    for(int i = 0; i <= TIMEBINS; i++) kf->gravkicks[i] = kf->hydrokicks[i] = 0;
    for(int i = times->mintimebin; i <= TIMEBINS; i++) {
        kf->gravkicks[i] = get_exact_gravkick_factor(CP, times->Ti_kick[i], times->Ti_Current);
        kf->hydrokicks[i] = get_exact_hydrokick_factor(CP, times->Ti_kick[i], times->Ti_Current);
    }
}
void slots_free_sph_pred_data(struct sph_pred_data *sph_scratch)               /* density.c:692-697 */
{
    if(sph_scratch->EntVarPred) myfree(sph_scratch->EntVarPred);
    sph_scratch->EntVarPred = NULL;
}

/* First guess of the smoothing length from the mass of the enclosing tree nodes
 * (density.c:700-749); reads the caller's host ForceTree, which needs Father[]. */
void set_init_hsml(ForceTree *tree, DomainDecomp *ddecomp, const double MeanGasSeparation)
{
    force_tree_calc_moments(tree, ddecomp);
    if(!tree->Father) endrun(5, "tree Father array not allocated at initial hsml!\n");
    const double DesNumNgb = GetNumNgb(GetDensityKernelType());
    #pragma omp parallel for
    for(int i = 0; i < PartManager->NumPart; i++) {
        if((P[i].Type != 0 && P[i].Type != 5) || P[i].IsGarbage) continue;
        int no = i;
        for(;;) {                /* climb until the node holds 10 x the desired neighbour mass */
            const int up = force_get_father(no, tree);
            if(up < tree->firstnode) break;
            if(up > tree->numnodes + tree->firstnode) endrun(5, "Bad init father of particle %d: %d\n", i, up);
            no = up;
            if(!(10 * DesNumNgb * P[i].Mass > tree->Nodes[no].mom.mass)) break;
        }
        double h = MeanGasSeparation;
        if(no >= tree->firstnode) {
            const struct NODE *nd = &tree->Nodes[no];
            if(nd->len > tree->BoxSize || nd->mom.mass < P[i].Mass)
                endrun(5, "Bad tree moments at particle %d: node %d len %g mass %g\n", i, no, nd->len, nd->mom.mass);
            const double guess = nd->len * pow(3.0 / (4 * M_PI) * DesNumNgb * P[i].Mass / nd->mom.mass, 1.0 / 3);
            if(guess < 500. * MeanGasSeparation) h = guess;
        }
        if(h <= 0) endrun(5, "Bad hsml guess for particle %d: %g\n", i, h);
        P[i].Hsml = h;
    }
}

/* ---- the two passes -------------------------------------------------------- */
static int64_t ShimNumPart = -1;         /* particle count the device gas tree was built for */
static int ShimDoEgy;

/* Per-bin factors of this step, exactly as the reference derives them from DriftKickTimes:
 * init_kick_factor_data (density.c:114-132), SPH_EntVarPred (density.c:74), drifts[] of
 * hydro_force (hydra.c:178-186), get_dloga_for_bin (hydra.c:271,463). */
static void fill_bins(b200_sph_bins *B, double *pmkick, const DriftKickTimes *times, Cosmology *CP)
{
    struct kick_factor_data kf;
    init_kick_factor_data(&kf, times, CP);
    memset(B, 0, sizeof(*B));
    *pmkick = kf.FgravkickB;
    for(int b = 0; b <= TIMEBINS; b++) {
        B->gravkick[b] = kf.gravkicks[b]; B->hydrokick[b] = kf.hydrokicks[b];
        if(b < times->mintimebin) continue;
        B->dloga_pred[b] = dloga_from_dti(times->Ti_Current - times->Ti_kick[b], times->Ti_Current);
        if(!is_timebin_active(b, times->Ti_Current))
            B->drift[b] = get_exact_drift_factor(CP, times->Ti_lastactivedrift[b], times->Ti_Current);
        B->dloga_bin[b] = get_dloga_for_bin(b, times->Ti_Current);
    }
}

static void fill_params(b200_sph_params *sp, double pmkick)
{
    memset(sp, 0, sizeof(*sp));
    sp->KernelType = (int) DensityParams.DensityKernelType;
    sp->DensityIndependentSphOn = HydroParams.DensityIndependentSphOn;
    sp->DensityResolutionEta = DensityParams.DensityResolutionEta;
    sp->MaxNumNgbDeviation = DensityParams.MaxNumNgbDeviation;
    sp->MinGasHsml = DensityParams.MinGasHsmlFractional * (FORCE_SOFTENING() / 2.8);      /* density.c:265 */
    sp->ArtBulkViscConst = HydroParams.ArtBulkViscConst;
    sp->DensityContrastLimit = HydroParams.DensityContrastLimit;
    sp->pmkick = pmkick;
}

/* time bins of every particle + the factor tables -> engine */
static void push_bins(b200_ctx *ctx, const b200_sph_bins *B)
{
    const int64_t n = PartManager->NumPart;
    const size_t m = (size_t) (n > 0 ? n : 1);
    uint8_t *bg = (uint8_t *) mymalloc("B200Bins", 2 * m), *bh = bg + m;
    #pragma omp parallel for
    for(int64_t i = 0; i < n; i++) { bg[i] = P[i].TimeBinGravity; bh[i] = P[i].TimeBinHydro; }
    B200_CK(b200_sph_set_timebins(ctx, bg, bh, B));
    myfree(bg);
}

static int is_target(const int64_t i) { return P[i].Type == 0 && !P[i].IsGarbage && !P[i].Swallowed; }

void density(const ActiveParticles *act, int update_hsml, int DoEgyDensity, int BlackHoleOn, const DriftKickTimes times,
             Cosmology *CP, struct sph_pred_data *SPH_predicted, MyFloat *GradRho_mag, const ForceTree *const tree)
{
    b200_ctx *ctx = sph_ctx();
    const int64_t n = PartManager->NumPart;
    if(BlackHoleOn && SlotsManager->info[5].size > 0)
        endrun(1, "b200 density(): black-hole density targets are not supported by this build\n");
    if(!(tree->mask & GASMASK)) endrun(1, "b200 density(): the tree holds no gas\n");

    b200_particle_layout lay;
    b200_default_particle_layout(&lay);
    if(sizeof(struct particle_data) != (size_t) lay.stride)
        endrun(2, "b200: struct particle_data is %lu bytes, the shim was built for %ld\n", sizeof(struct particle_data), (long) lay.stride);
    B200_CK(b200_set_particles_aos(ctx, P, n, &lay));
    b200_shim_topnodes_from_tree(tree);
    B200_CK(b200_tree_build(ctx, tree->BoxSize, GASMASK, NULL, 0, -1, NULL));      /* force_tree_rebuild_mask(GASMASK), run.c:466 */
    ShimNumPart = n; ShimDoEgy = DoEgyDensity;

    const size_t m = (size_t) (n > 0 ? n : 1);
    double *buf = (double *) mymalloc("B200SphIn", sizeof(double) * 15 * m);
    double *vel = buf, *hsml = vel + 3 * m, *ent = hsml + m, *dte = ent + m, *fa = dte + m, *gp = fa + 3 * m, *ha = gp + 3 * m;
    #pragma omp parallel for
    for(int64_t i = 0; i < n; i++) {
        const int gas = P[i].Type == 0 && !P[i].IsGarbage;
        for(int k = 0; k < 3; k++) {
            vel[3 * i + k] = P[i].Vel[k]; fa[3 * i + k] = P[i].FullTreeGravAccel[k]; gp[3 * i + k] = P[i].GravPM[k];
            ha[3 * i + k] = gas ? SPHP(i).HydroAccel[k] : 0;
        }
        hsml[i] = P[i].Hsml;
        ent[i] = gas ? SPHP(i).Entropy : 1; dte[i] = gas ? SPHP(i).DtEntropy : 0;
    }
    B200_CK(b200_sph_set_gas(ctx, vel, hsml, ent, dte, fa, gp, ha));
    /* SphP state of the gas that is not a target of this call (stale neighbours in hydro) */
    double *st_rho = buf, *st_egy = st_rho + m, *st_fac = st_egy + m, *st_div = st_fac + m, *st_curl = st_div + m;
    #pragma omp parallel for
    for(int64_t i = 0; i < n; i++) {
        const int gas = P[i].Type == 0 && !P[i].IsGarbage;
        st_rho[i] = gas ? SPHP(i).Density : 0; st_egy[i] = gas ? SPHP(i).EgyWtDensity : 0;
        st_fac[i] = gas ? SPHP(i).DhsmlEgyDensityFactor : 0; st_div[i] = gas ? SPHP(i).DivVel : 0; st_curl[i] = gas ? SPHP(i).CurlVel : 0;
    }
    B200_CK(b200_sph_set_state(ctx, st_rho, st_egy, st_fac, st_div, st_curl));
    myfree(buf);

    b200_sph_bins bins;
    double pmkick;
    fill_bins(&bins, &pmkick, &times, CP);
    push_bins(ctx, &bins);
    B200_CK(b200_sph_set_active(ctx, act->ActiveParticle, act->ActiveParticle ? act->NumActiveParticle : 0));
    b200_sph_params sp;
    fill_params(&sp, pmkick);
    double *out = (double *) mymalloc("B200SphOut", sizeof(double) * 10 * m);
    double *o_h = out, *o_rho = o_h + m, *o_egy = o_rho + m, *o_fac = o_egy + m, *o_div = o_fac + m, *o_curl = o_div + m,
           *o_dth = o_curl + m, *o_grad = o_dth + m;
    B200_CK(b200_density(ctx, &sp, update_hsml, DoEgyDensity, o_h, o_rho, o_egy, o_fac, o_div, o_curl, o_dth, NULL, NULL, NULL));
    if(GradRho_mag) B200_CK(b200_density_gradrho(ctx, o_grad));

    /* EntVarPred for every gas slot (density.c:291-299; freed by the caller through slots_free_sph_pred_data) */
    SPH_predicted->EntVarPred = (MyFloat *) mymalloc2("EntVarPred", sizeof(MyFloat) * (SlotsManager->info[0].size > 0 ? SlotsManager->info[0].size : 1));
    #pragma omp parallel for
    for(int64_t i = 0; i < n; i++)
        if(P[i].Type == 0 && !P[i].IsGarbage) SPH_predicted->EntVarPred[P[i].PI] = SPH_EntVarPred(i, &times);
    const int64_t nq = act->ActiveParticle ? act->NumActiveParticle : n;
    #pragma omp parallel for
    for(int64_t q = 0; q < nq; q++) {
        const int64_t i = act->ActiveParticle ? act->ActiveParticle[q] : q;
        if(!is_target(i)) continue;                                           /* density_haswork density.c:521-530 */
        if(update_hsml) P[i].Hsml = o_h[i];
        P[i].DtHsml = o_dth[i];                                               /* density.c:574-577 */
        struct sph_particle_data *s = &SPHP(i);
        s->Density = o_rho[i];
        if(DoEgyDensity) s->EgyWtDensity = o_egy[i];
        s->DhsmlEgyDensityFactor = o_fac[i];
        s->DivVel = o_div[i]; s->CurlVel = o_curl[i];
        if(GradRho_mag) {
            const double *g = o_grad + 3 * i;
            GradRho_mag[P[i].PI] = sqrt(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
        }
    }
    myfree(out);
    b200_timings t;
    if(b200_get_timings(ctx, &t) == 0) walltime_add("/SPH/Density/WalkPrim", 1e-3 * t.sph_density);
}

void hydro_force(const ActiveParticles *act, const double atime, struct sph_pred_data *SPH_predicted, const DriftKickTimes times,
                 Cosmology *CP, const ForceTree *const tree)
{
    b200_ctx *ctx = sph_ctx();
    const int64_t n = PartManager->NumPart;
    if(ShimNumPart != n) endrun(5, "Hydro called before hmax computed\n");            /* hydra.c:174-175 */
    if(HydroParams.DensityIndependentSphOn && !ShimDoEgy)
        endrun(1, "b200 hydro_force(): pressure-entropy SPH needs density() with DoEgyDensity\n");
    for(int64_t i = 0; i < n; i++)
        if(P[i].Type == 0 && !P[i].IsGarbage && winds_is_particle_decoupled(i))              /* hydra.c:362,523 */
            endrun(1, "b200 hydro_force(): decoupled wind particles are not supported by this build\n");

    b200_sph_bins bins;
    double pmkick;
    fill_bins(&bins, &pmkick, &times, CP);
    push_bins(ctx, &bins);
    B200_CK(b200_sph_set_active(ctx, act->ActiveParticle, act->ActiveParticle ? act->NumActiveParticle : 0));
    b200_sph_params sp;
    fill_params(&sp, pmkick);
    sp.atime = atime; sp.hubble = hubble_function(CP, atime);

    const size_t m = (size_t) (n > 0 ? n : 1);
    double *out = (double *) mymalloc("B200HydroOut", sizeof(double) * 5 * m);
    double *acc = out, *dte = acc + 3 * m, *sig = dte + m;
    B200_CK(b200_hydro_force(ctx, &sp, acc, dte, sig, NULL));
    const int64_t nq = act->ActiveParticle ? act->NumActiveParticle : n;
    #pragma omp parallel for
    for(int64_t q = 0; q < nq; q++) {
        const int64_t i = act->ActiveParticle ? act->ActiveParticle[q] : q;
        if(!is_target(i)) continue;                                                  /* hydro_haswork, treewalk.c:234 */
        struct sph_particle_data *s = &SPHP(i);
        for(int k = 0; k < 3; k++) s->HydroAccel[k] = acc[3 * i + k];                /* hydro_reduce hydra.c:279-293 */
        s->DtEntropy = dte[i];
        s->MaxSignalVel = sig[i];
    }
    myfree(out);
    b200_timings t;
    if(b200_get_timings(ctx, &t) == 0) walltime_add("/SPH/Hydro/WalkPrim", 1e-3 * t.sph_hydro);
}
