/* ref_driver.c -- drives the reference's OWN tree / treewalk / short-range
 * gravity code (compiled unmodified from /root/reference by Makefile.ref) on
 * caller-supplied particles, as one MPI "rank".  TEST INFRASTRUCTURE ONLY:
 * used to pin the oracle restatement and as the CPU baseline
 * (cpu_baseline.kind = "reference").
 *
 * It plays the role of the reference's test fixtures
 * (libgadget/tests/test_gravity.c:162-220 do_force_test,
 * tests/test_forcetree.c:410-430 trivial_domain, tests/stub.c:19-41).
 */
#include <mpi.h>
#include <omp.h>
#include <math.h>
#include <string.h>
#include <stdlib.h>
#include <stdio.h>

#include <libgadget/utils/endrun.h>
#include <libgadget/utils/mymalloc.h>
#include <libgadget/utils/system.h>
#include <libgadget/utils/peano.h>
#include <libgadget/partmanager.h>
#include <libgadget/slotsmanager.h>
#include <libgadget/domain.h>
#include <libgadget/forcetree.h>
#include <libgadget/treewalk.h>
#include <libgadget/gravity.h>
#include <libgadget/petapm.h>
#include <libgadget/timestep.h>
#include <libgadget/walltime.h>
#include <libgadget/density.h>
#include <libgadget/hydra.h>
#include <libgadget/cosmology.h>

/* ---- stand-ins for functions that live in reference files we do not build
 * (timestep.c needs cosmology/GSL; petaio needs bigfile) ------------------- */
#ifndef REF_WITH_STEP      /* libref_step.so links the reference's own timestep.c / timebinmgr.c */
int is_timebin_active(int i, inttime_t current)          /* timestep.c:143-150 */
{
    if(i <= 0 || current <= 0) return 1;
    if(current % dti_from_timebin(i) == 0) return 1;
    return 0;
}
ActiveParticles init_empty_active_particles(struct part_manager_type *PartManager)   /* timestep.c */
{
    ActiveParticles act = {0};
    act.ActiveParticle = NULL;
    act.NumActiveParticle = PartManager->NumPart;
    act.MaxActiveParticle = PartManager->NumPart;
    act.Particles = PartManager->Base;
    return act;
}
#endif
void dump_snapshot(const char *dump, const double Time, void *CP, const char *OutputDir) {}

/* Time-integration helpers (timebinmgr.c, timefac.c need GSL; cosmology.c needs GSL;
 * winds.c is sub-grid physics).  The SPH fixtures are synchronised: every particle
 * sits on time bin 0 at Ti_Current = 0 with all kick times equal, for which the
 * reference's own functions return exactly these values (timefac.c:44-45:
 * t0 == t1 -> 0; timebinmgr.c:420-447 with dti = 0 / bin 0 -> 0). */
static double sph_dloga_bin, sph_hubble;
/* Mixed-bin fixtures (ref_sph_mixed below): the kick times are laid out so that the
 * integer-time difference handed to each function IS the time bin (Ti_kick[b] =
 * Ti_lastactivedrift[b] = Ti_Current - b, PM_kick = Ti_Current - MIXED_PM), and the value is read
 * from a per-bin table supplied by the test -- the reference's integrals over the scale
 * factor (timefac.c:12-73, GSL) are replaced by numbers, its SPH code is not touched. */
#define MIXED_PM (TIMEBINS + 1)
static int mixed_on;
static double mx_gravkick[TIMEBINS + 2], mx_hydrokick[TIMEBINS + 2], mx_drift[TIMEBINS + 2], mx_dloga_pred[TIMEBINS + 2], mx_dloga_bin[TIMEBINS + 2];
static int mixed_key(inttime_t t0, inttime_t t1)
{
    const int64_t d = (int64_t) t1 - (int64_t) t0;
    if(d < 0 || d > TIMEBINS + 1) endrun(1, "ref_driver: unexpected kick-time difference %ld\n", (long) d);
    return (int) d;
}
#ifndef REF_WITH_STEP
double dloga_from_dti(inttime_t dti, const inttime_t Ti_Current)
{
    if(mixed_on) return mx_dloga_pred[mixed_key(0, dti)];
    if(dti != 0) endrun(1, "ref_driver: unsynchronised fixture\n");
    return 0;
}
double get_dloga_for_bin(int timebin, const inttime_t Ti_Current) { return mixed_on ? mx_dloga_bin[timebin] : sph_dloga_bin; }
static double exact_factor(inttime_t t0, inttime_t t1) { if(t0 != t1) endrun(1, "ref_driver: unsynchronised fixture\n"); return 0; }
double get_exact_drift_factor(Cosmology *CP, inttime_t t0, inttime_t t1) { return mixed_on ? mx_drift[mixed_key(t0, t1)] : exact_factor(t0, t1); }
double get_exact_gravkick_factor(Cosmology *CP, inttime_t t0, inttime_t t1) { return mixed_on ? mx_gravkick[mixed_key(t0, t1)] : exact_factor(t0, t1); }
double get_exact_hydrokick_factor(Cosmology *CP, inttime_t t0, inttime_t t1) { return mixed_on ? mx_hydrokick[mixed_key(t0, t1)] : exact_factor(t0, t1); }
double hubble_function(const Cosmology *CP, double a) { return sph_hubble; }
#endif
int winds_is_particle_decoupled(int i) { return 0; }
void winds_decoupled_hydro(int i, double atime) {}
/* set_hydro_params (hydra.c:37-48) reads its three keys through param_get_*; the
 * driver supplies them from a table instead of a parsed parameter file. */
static double hp_visc, hp_contrast; static int hp_di;
#ifndef REF_WITH_STEP
double param_get_double(ParameterSet *ps, const char *name)
{
    if(!strcmp(name, "ArtBulkViscConst")) return hp_visc;
    if(!strcmp(name, "DensityContrastLimit")) return hp_contrast;
    endrun(1, "ref_driver: unexpected parameter %s\n", name); return 0;
}
int param_get_int(ParameterSet *ps, const char *name)
{
    if(!strcmp(name, "DensityIndependentSphOn")) return hp_di;
    endrun(1, "ref_driver: unexpected parameter %s\n", name); return 0;
}
#endif
int param_get_enum(ParameterSet *ps, const char *name) { endrun(1, "ref_driver: unexpected enum %s\n", name); return 0; }

int ref_stub_ntask = 1;          /* what the stand-in MPI_Comm_size reports (oracle/stubs/mpi.h) */
int ref_stub_thistask = 0;       /* ... and MPI_Comm_rank */
static struct ClockTable CT;
static int initialised = 0;
static DomainDecomp dd;
static ForceTree Tree;
static double t_build, t_walk;

static void build_uniform_domain(DomainDecomp *d, int depth)
{
    /* A complete top tree of the given depth in Peano-Hilbert order: the shape
     * domain_decompose_full produces for a uniform load (domain.c:154-256);
     * depth 0 is trivial_domain of tests/test_forcetree.c:410-430. */
    int ntop = 0, nleaf = 1;
    for(int l = 0, c = 1; l <= depth; l++, c *= 8) { ntop += c; if(l == depth) nleaf = c; }
    d->domain_allocated_flag = 1;
    d->NTopNodes = ntop;
    d->NTopLeaves = nleaf;
    d->TopNodes = (struct topnode_data *) mymalloc("TopNodes", sizeof(struct topnode_data) * ntop);
    d->TopLeaves = (struct topleaf_data *) mymalloc("TopLeaves", sizeof(struct topleaf_data) * nleaf);
    d->Tasks = (struct task_data *) mymalloc("Tasks", sizeof(struct task_data));
    d->Tasks[0].StartLeaf = 0;
    d->Tasks[0].EndLeaf = nleaf;
    d->DomainComm = MPI_COMM_WORLD;
    /* breadth-first numbering: node t at level l has daughters at first(l+1) + 8*(t - first(l)) */
    int first = 0, count = 1, nextleaf = 0;
    d->TopNodes[0].StartKey = 0;
    d->TopNodes[0].Shift = 3 * BITS_PER_DIMENSION;
    for(int l = 0; l <= depth; l++) {
        const int firstnext = first + count;
        for(int t = first; t < first + count; t++) {
            if(l == depth) {
                d->TopNodes[t].Daughter = -1;
                d->TopNodes[t].Leaf = -1;      /* assigned below in key order */
            } else {
                const int dau = firstnext + 8 * (t - first);
                d->TopNodes[t].Daughter = dau;
                d->TopNodes[t].Leaf = -1;
                for(int j = 0; j < 8; j++) {
                    d->TopNodes[dau + j].Shift = d->TopNodes[t].Shift - 3;
                    d->TopNodes[dau + j].StartKey = d->TopNodes[t].StartKey + ((peano_t) j << d->TopNodes[dau + j].Shift);
                }
            }
        }
        first = firstnext; count *= 8;
    }
    /* leaves numbered along the curve: BFS order inside the last level is key order */
    {
        int lfirst = ntop - nleaf;
        for(int t = lfirst; t < ntop; t++) {
            d->TopNodes[t].Leaf = nextleaf;
            d->TopLeaves[nextleaf].Task = 0;
            d->TopLeaves[nextleaf].topnode = t;
            d->TopLeaves[nextleaf].treenode = -1;
            nextleaf++;
        }
    }
}

void ref_build_uniform_domain(DomainDecomp *d, int depth) { build_uniform_domain(d, depth); }
/* An arbitrary top tree handed in as the columns of DomainDecomp::TopNodes (domain.h:20-33); leaves numbered by `leaf`. */
static void build_domain_from_arrays(DomainDecomp *d, int ntop, const int *daughter, const uint64_t *startkey, const int *shift, const int *leaf)
{
    int nleaf = 0;
    for(int t = 0; t < ntop; t++) if(daughter[t] < 0) nleaf++;
    d->domain_allocated_flag = 1;
    d->NTopNodes = ntop;
    d->NTopLeaves = nleaf;
    d->TopNodes = (struct topnode_data *) mymalloc("TopNodes", sizeof(struct topnode_data) * ntop);
    d->TopLeaves = (struct topleaf_data *) mymalloc("TopLeaves", sizeof(struct topleaf_data) * nleaf);
    d->Tasks = (struct task_data *) mymalloc("Tasks", sizeof(struct task_data));
    d->Tasks[0].StartLeaf = 0;
    d->Tasks[0].EndLeaf = nleaf;
    d->DomainComm = MPI_COMM_WORLD;
    for(int t = 0; t < ntop; t++) {
        memset(&d->TopNodes[t], 0, sizeof(d->TopNodes[t]));
        d->TopNodes[t].Daughter = daughter[t]; d->TopNodes[t].StartKey = startkey[t]; d->TopNodes[t].Shift = shift[t];
        d->TopNodes[t].Leaf = daughter[t] < 0 ? leaf[t] : -1;
        if(daughter[t] < 0) {
            d->TopLeaves[leaf[t]].Task = 0;
            d->TopLeaves[leaf[t]].topnode = t;
            d->TopLeaves[leaf[t]].treenode = -1;
        }
    }
}
     /* for the other fixture files */

int ref_init(double arena_gib, int nthreads)
{
    if(initialised) return 0;
    if(nthreads > 0) omp_set_num_threads(nthreads);
    init_endrun(0);
    tamalloc_init();
    mymalloc_init(arena_gib * 1024.);       /* MiB */
    walltime_init(&CT);
    init_forcetree_params(0.7);            /* TreeAllocFactor default, gadget/params.c */
    initialised = 1;
    return 0;
}

static int have_particles = 0;
static int slots_ready = 0;
static struct sph_pred_data sph_pred;
static void free_all(void)
{
    /* the arena is a stack: release in reverse order of allocation (utils/mymalloc.h:15-19) */
    if(sph_pred.EntVarPred) slots_free_sph_pred_data(&sph_pred);
    if(force_tree_allocated(&Tree)) force_tree_free(&Tree);
    if(dd.domain_allocated_flag) {
        myfree(dd.Tasks); myfree(dd.TopLeaves); myfree(dd.TopNodes);
        memset(&dd, 0, sizeof(dd));
    }
    if(slots_ready) { slots_free(SlotsManager); slots_ready = 0; }
    if(have_particles) { myfree(P); have_particles = 0; }
}

/* Load particles (DM type 1 unless type given), build the domain and the full tree.
 * oldacc[n][3] is stored in FullTreeGravAccel (GravPM = 0) for the relative criterion. */
static int ref_tree_build_dd(int64_t n, const double *pos, const float *mass, const unsigned char *type,
                             const double *oldacc, double BoxSize, int topdepth,
                             int ntop, const int *daughter, const uint64_t *startkey, const int *shift, const int *leaf);

int ref_tree_build(int64_t n, const double *pos, const float *mass, const unsigned char *type,
                   const double *oldacc, double BoxSize, int topdepth)
{
    return ref_tree_build_dd(n, pos, mass, type, oldacc, BoxSize, topdepth, 0, NULL, NULL, NULL, NULL);
}

/* The same below an arbitrary domain top tree (force_tree_create_topnodes, forcetree.c:654-687,869-934). */
int ref_tree_build_top(int64_t n, const double *pos, const float *mass, const unsigned char *type,
                       const double *oldacc, double BoxSize,
                       int ntop, const int *daughter, const uint64_t *startkey, const int *shift, const int *leaf)
{
    return ref_tree_build_dd(n, pos, mass, type, oldacc, BoxSize, 0, ntop, daughter, startkey, shift, leaf);
}

static int ref_tree_build_dd(int64_t n, const double *pos, const float *mass, const unsigned char *type,
                             const double *oldacc, double BoxSize, int topdepth,
                             int ntop, const int *daughter, const uint64_t *startkey, const int *shift, const int *leaf)
{
    free_all();
    particle_alloc_memory(PartManager, BoxSize, n);
    have_particles = 1;
    PartManager->NumPart = n;
    if(ntop > 0) build_domain_from_arrays(&dd, ntop, daughter, startkey, shift, leaf);
    else build_uniform_domain(&dd, topdepth);
    #pragma omp parallel for
    for(int64_t i = 0; i < n; i++) {
        memset(&P[i], 0, sizeof(P[i]));
        for(int k = 0; k < 3; k++) {
            P[i].Pos[k] = pos[3 * i + k];
            P[i].FullTreeGravAccel[k] = oldacc ? oldacc[3 * i + k] : 0;
        }
        P[i].Mass = mass[i];
        P[i].Type = type ? type[i] : 1;
        P[i].ID = i;
        P[i].TopLeaf = domain_get_topleaf(PEANO(P[i].Pos, BoxSize), &dd);
    }
    const double t0 = omp_get_wtime();
    force_tree_full(&Tree, &dd, 0, NULL);
    t_build = omp_get_wtime() - t0;
    return 0;
}

int ref_grav_short_tree(double G, int Nmesh, double Asmth, double ErrTolForceAcc, double BHOpeningAngle,
                        double MaxBHOpeningAngle, int TreeUseBH, double Rcut, double GravitySoftening, double rho0,
                        double *acc_out, double *pot_out)
{
    PetaPM pm;
    memset(&pm, 0, sizeof(pm));
    pm.BoxSize = PartManager->BoxSize; pm.Asmth = Asmth; pm.Nmesh = Nmesh; pm.G = G;
    pm.CellSize = pm.BoxSize / Nmesh;
    gravshort_fill_ntab(SHORTRANGE_FORCE_WINDOW_TYPE_EXACT, Asmth);
    struct gravshort_tree_params tp = {0};
    tp.ErrTolForceAcc = ErrTolForceAcc; tp.BHOpeningAngle = BHOpeningAngle; tp.MaxBHOpeningAngle = MaxBHOpeningAngle;
    tp.TreeUseBH = TreeUseBH; tp.Rcut = Rcut; tp.FractionalGravitySoftening = GravitySoftening;
    set_gravshort_treepar(tp);
    gravshort_set_softenings(1.0);
    ActiveParticles act = init_empty_active_particles(PartManager);
    const double t0 = omp_get_wtime();
    grav_short_tree(&act, &pm, &Tree, NULL, rho0, 0);
    t_walk = omp_get_wtime() - t0;
    const int64_t n = PartManager->NumPart;
    #pragma omp parallel for
    for(int64_t i = 0; i < n; i++) {
        if(acc_out) for(int k = 0; k < 3; k++) acc_out[3 * i + k] = P[i].FullTreeGravAccel[k];
        if(pot_out) pot_out[i] = P[i].Potential;
    }
    return 0;
}

/* Domain keys: PEANO() (utils/peano.h:15-21 over peano.c:108-129) and domain_get_topleaf (domain.h:71-78) over a top
 * tree handed in as arrays (TopNodes[].Daughter / StartKey / Shift / Leaf). */
void ref_peano_keys(int64_t n, const double *pos, double BoxSize, uint64_t *keys)
{
    for(int64_t i = 0; i < n; i++) keys[i] = PEANO(&pos[3 * i], BoxSize);
}
void ref_topleaf(int64_t n, const uint64_t *keys, int ntop, const int *daughter, const uint64_t *startkey, const int *shift, const int *leaf, int *out)
{
    DomainDecomp d;
    memset(&d, 0, sizeof(d));
    d.TopNodes = (struct topnode_data *) malloc(sizeof(struct topnode_data) * ntop);
    for(int t = 0; t < ntop; t++) {
        memset(&d.TopNodes[t], 0, sizeof(d.TopNodes[t]));
        d.TopNodes[t].Daughter = daughter[t]; d.TopNodes[t].StartKey = startkey[t]; d.TopNodes[t].Shift = shift[t]; d.TopNodes[t].Leaf = leaf[t];
    }
    d.NTopNodes = ntop;
    for(int64_t i = 0; i < n; i++) out[i] = domain_get_topleaf(keys[i], &d);
    free(d.TopNodes);
}

void ref_timings(double *build_s, double *walk_s) { *build_s = t_build; *walk_s = t_walk; }
int64_t ref_numnodes(void) { return Tree.numnodes; }

/* Export the tree in walk order (sibling / suns[0]) for comparison with the oracle.
 * Returns the number of nodes visited (all arrays sized ref_numnodes()). */
int64_t ref_tree_export(double *center, double *len, double *cofm, double *mass, int *nocc, int *part, int *toplevel)
{
    int64_t k = 0;
    int no = Tree.firstnode;
    while(no >= 0) {
        struct NODE *nop = &Tree.Nodes[no];
        for(int j = 0; j < 3; j++) { center[3 * k + j] = nop->center[j]; cofm[3 * k + j] = nop->mom.cofm[j]; }
        len[k] = nop->len; mass[k] = nop->mom.mass;
        toplevel[k] = nop->f.TopLevel;
        for(int j = 0; j < 8; j++) part[8 * k + j] = -1;
        if(nop->f.ChildType == PARTICLE_NODE_TYPE) {
            nocc[k] = nop->s.noccupied;
            for(int j = 0; j < nop->s.noccupied; j++) part[8 * k + j] = nop->s.suns[j];
            no = nop->sibling;
        } else if(nop->f.ChildType == NODE_NODE_TYPE) {
            nocc[k] = -1;
            no = nop->s.suns[0];
        } else {
            nocc[k] = -2;
            no = nop->sibling;
        }
        k++;
    }
    return k;
}

/* ---- SPH: the reference's density() and hydro_force() on a gas-only fixture, as
 * tests/test_density.c:55-152 sets it up (all particles type 0, time bin 0). ---- */
static double t_density, t_hydro;

int ref_sph_density(int64_t n, const double *pos, const float *mass, const double *vel, const double *entropy,
                    double BoxSize, int kerneltype, double eta, double maxdev, double mingashsml_frac, double softening,
                    int init_hsml, double meansep, int update_hsml, int DoEgyDensity,
                    double *hsml /*in/out*/, double *density_out, double *egy_out, double *dhsmlfac_out,
                    double *divvel_out, double *curlvel_out, double *dthsml_out)
{
    free_all();
    particle_alloc_memory(PartManager, BoxSize, n);
    have_particles = 1;
    PartManager->NumPart = n;
    slots_init(0.01 * n, SlotsManager);
    slots_set_enabled(0, sizeof(struct sph_particle_data), SlotsManager);
    int64_t atleast[6] = {0};
    atleast[0] = n;
    slots_reserve(1, atleast, SlotsManager);
    slots_ready = 1;
    SlotsManager->info[0].size = n;
    build_uniform_domain(&dd, 0);
    for(int64_t i = 0; i < n; i++) {
        memset(&P[i], 0, sizeof(P[i]));
        for(int k = 0; k < 3; k++) { P[i].Pos[k] = pos[3 * i + k]; P[i].Vel[k] = vel ? vel[3 * i + k] : 0; }
        P[i].Mass = mass[i]; P[i].Type = 0; P[i].PI = i; P[i].ID = i; P[i].Hsml = hsml[i];
        P[i].TopLeaf = 0;
        memset(&SPHP(i), 0, sizeof(struct sph_particle_data));
        SPHP(i).Entropy = entropy ? entropy[i] : 1; SPHP(i).DtEntropy = 0; SPHP(i).Density = 1;
    }
    struct density_params dp = {0};
    dp.DensityResolutionEta = eta; dp.MaxNumNgbDeviation = maxdev; dp.BlackHoleNgbFactor = 2;
    dp.BlackHoleMaxAccretionRadius = 99999.; dp.DensityKernelType = (enum DensityKernelType) kerneltype;
    dp.MinGasHsmlFractional = mingashsml_frac;
    set_densitypar(dp);
    struct gravshort_tree_params tp = {0};
    tp.FractionalGravitySoftening = softening;
    set_gravshort_treepar(tp);
    gravshort_set_softenings(1);
    ActiveParticles act = init_empty_active_particles(PartManager);
    if(init_hsml) {      /* tests/test_density.c:84-86 */
        force_tree_rebuild_mask(&Tree, &dd, GASMASK + BHMASK, NULL);
        set_init_hsml(&Tree, &dd, meansep);
    }
    force_tree_rebuild_mask(&Tree, &dd, GASMASK, NULL);
    DriftKickTimes kick = {0};
    Cosmology CP = {0};
    sph_pred.EntVarPred = NULL;
    const double t0 = omp_get_wtime();
    density(&act, update_hsml, DoEgyDensity, 0, kick, &CP, &sph_pred, NULL, &Tree);
    t_density = omp_get_wtime() - t0;
    for(int64_t i = 0; i < n; i++) {
        hsml[i] = P[i].Hsml; density_out[i] = SPHP(i).Density; egy_out[i] = SPHP(i).EgyWtDensity;
        dhsmlfac_out[i] = SPHP(i).DhsmlEgyDensityFactor; divvel_out[i] = SPHP(i).DivVel; curlvel_out[i] = SPHP(i).CurlVel;
        dthsml_out[i] = P[i].DtHsml;
    }
    return 0;
}

/* hydro_force right after ref_sph_density (run.c:472-489: density, force_tree_calc_moments, hydro_force). */
int ref_sph_hydro(double atime, double hubble, double dloga_bin, int DensityIndependentSphOn, double ArtBulkViscConst,
                  double DensityContrastLimit, double *acc_out, double *dtentropy_out, double *maxsig_out)
{
    const int64_t n = PartManager->NumPart;
    sph_hubble = hubble; sph_dloga_bin = dloga_bin;
    hp_visc = ArtBulkViscConst; hp_contrast = DensityContrastLimit; hp_di = DensityIndependentSphOn;
    set_hydro_params(NULL);
    force_tree_calc_moments(&Tree, &dd);
    ActiveParticles act = init_empty_active_particles(PartManager);
    DriftKickTimes kick = {0};
    Cosmology CP = {0};
    const double t0 = omp_get_wtime();
    hydro_force(&act, atime, &sph_pred, kick, &CP, &Tree);
    t_hydro = omp_get_wtime() - t0;
    for(int64_t i = 0; i < n; i++) {
        for(int k = 0; k < 3; k++) acc_out[3 * i + k] = SPHP(i).HydroAccel[k];
        dtentropy_out[i] = SPHP(i).DtEntropy; maxsig_out[i] = SPHP(i).MaxSignalVel;
    }
    slots_free_sph_pred_data(&sph_pred);
    return 0;
}
/* A mixed-time-bin step on the state left by ref_sph_density + ref_sph_hydro (which this call
 * needs first; pass keep_pred = 1 to ref_sph_hydro... the predictor array is re-made here):
 * particle i sits on time bin bins[i] (hydro and gravity); the bins that divide Ti_Current are
 * active (is_timebin_active, timestep.c:143-150) and form the ActiveParticles list, as
 * build_active_particles does; density() and hydro_force() then run for the active particles
 * only, with inactive neighbours contributing their stale SphP state.  tables[5][TIMEBINS+2]:
 * gravkick, hydrokick, drift, dloga_pred, dloga_bin by bin (index TIMEBINS+1 of gravkick = the PM kick factor). */
int ref_sph_mixed(const unsigned char *bins, int64_t Ti_Current, const double *tables, const double *vel_new,
                  const double *fullacc, const double *hydroacc_in, const double *dtentropy_in,
                  double atime, double hubble, int DoEgyDensity, int *active_out, int64_t *nactive_out,
                  double *hsml, double *density_out, double *egy_out, double *dhsmlfac_out, double *divvel_out, double *curlvel_out,
                  double *dthsml_out, double *acc_out, double *dtentropy_out, double *maxsig_out)
{
    const int64_t n = PartManager->NumPart;
    const int NT = TIMEBINS + 2;
    memcpy(mx_gravkick, tables, sizeof(double) * NT); memcpy(mx_hydrokick, tables + NT, sizeof(double) * NT);
    memcpy(mx_drift, tables + 2 * NT, sizeof(double) * NT); memcpy(mx_dloga_pred, tables + 3 * NT, sizeof(double) * NT);
    memcpy(mx_dloga_bin, tables + 4 * NT, sizeof(double) * NT);
    mixed_on = 1; sph_hubble = hubble;
    DriftKickTimes times;
    memset(&times, 0, sizeof(times));
    times.Ti_Current = Ti_Current;
    times.mintimebin = 1; times.maxtimebin = TIMEBINS; times.mingravtimebin = 1;
    for(int b = 0; b <= TIMEBINS; b++) { times.Ti_kick[b] = Ti_Current - b; times.Ti_lastactivedrift[b] = Ti_Current - b; }
    times.PM_kick = Ti_Current - MIXED_PM;
    for(int64_t i = 0; i < n; i++) {
        P[i].TimeBinHydro = bins[i]; P[i].TimeBinGravity = bins[i];
        P[i].Ti_drift = Ti_Current;             /* all particles are drifted to the current time (drift.c:17-102) */
        for(int k = 0; k < 3; k++) {
            if(vel_new) P[i].Vel[k] = vel_new[3 * i + k];
            if(fullacc) P[i].FullTreeGravAccel[k] = fullacc[3 * i + k];
            if(hydroacc_in) SPHP(i).HydroAccel[k] = hydroacc_in[3 * i + k];
        }
        if(dtentropy_in) SPHP(i).DtEntropy = dtentropy_in[i];
    }
    force_tree_rebuild_mask(&Tree, &dd, GASMASK, NULL);
    /* the arena is a stack: the list goes on top of the rebuilt tree */
    int *list = (int *) mymalloc("ActiveList", sizeof(int) * (n > 0 ? n : 1));
    int64_t na = 0;
    for(int64_t i = 0; i < n; i++)
        if(is_timebin_active(bins[i], Ti_Current)) list[na++] = i;
    ActiveParticles act = {0};
    act.ActiveParticle = list; act.NumActiveParticle = na; act.MaxActiveParticle = n; act.NumActiveHydro = na; act.NumActiveGravity = na;
    act.Particles = PartManager->Base;
    Cosmology CP = {0};
    sph_pred.EntVarPred = NULL;
    density(&act, 1, DoEgyDensity, 0, times, &CP, &sph_pred, NULL, &Tree);
    force_tree_calc_moments(&Tree, &dd);
    hydro_force(&act, atime, &sph_pred, times, &CP, &Tree);
    for(int64_t i = 0; i < n; i++) {
        hsml[i] = P[i].Hsml; density_out[i] = SPHP(i).Density; egy_out[i] = SPHP(i).EgyWtDensity;
        dhsmlfac_out[i] = SPHP(i).DhsmlEgyDensityFactor; divvel_out[i] = SPHP(i).DivVel; curlvel_out[i] = SPHP(i).CurlVel;
        dthsml_out[i] = P[i].DtHsml;
        for(int k = 0; k < 3; k++) acc_out[3 * i + k] = SPHP(i).HydroAccel[k];
        dtentropy_out[i] = SPHP(i).DtEntropy; maxsig_out[i] = SPHP(i).MaxSignalVel;
    }
    for(int64_t q = 0; q < na; q++) active_out[q] = list[q];
    *nactive_out = na;
    slots_free_sph_pred_data(&sph_pred);
    myfree(list);
    mixed_on = 0;
    return 0;
}

void ref_sph_timings(double *dens_s, double *hydro_s) { *dens_s = t_density; *hydro_s = t_hydro; }

#ifdef REF_WITH_PM
/* ---- PM: the reference's own petapm.c + gravpm.c + powerspectrum.c on one rank, with the PFFT
 * stand-in of oracle/pfft_standin.c (plain DFTs in PFFT's layout).  Stand-ins for the entry points
 * of files that need GSL (cosmology.c, neutrinos_lra.c, omega_nu_single.c); none of them influences
 * the forces with MassiveNuLinRespOn = HybridNeutrinosOn = 0. */
#include <libgadget/powerspectrum.h>
#include <libgadget/neutrinos_lra.h>
double GrowthFactor(Cosmology *CP, double astart, double aend) { return 1.0; }       /* only scales a column of the P(k) file */
int hybrid_nu_tracer(const Cosmology *CP, double atime) { return 0; }
double get_omega_nu_nopart(const _omega_nu *const omnu, const double a) { return 0; }
void delta_nu_from_power(struct _powerspectrum *PowerSpectrum, Cosmology *CP, const double Time, const double TimeIC) { endrun(1, "ref_driver: neutrino branch\n"); }
void powerspectrum_nu_save(struct _powerspectrum *PowerSpectrum, const char *OutputDir, const char *filename, const double Time) {}

/* gravpm_init_periodic + gravpm_force (gravpm.c:51-119) as run.c:330,522 call them: fills P[],
 * builds the domain, runs the reference PM (region selection from its own tree, CIC, transfer
 * functions, readout) and returns P[i].GravPM, P[i].Potential.  The power spectrum is written by
 * the reference itself to outdir/powerspectrum-<Time>.txt. */
int ref_gravpm_force(int64_t n, const double *pos, const float *mass, double BoxSize, int Nmesh, double Asmth, double G,
                     const char *outdir, double Time, double *gravpm_out, double *pot_out)
{
    free_all();
    particle_alloc_memory(PartManager, BoxSize, n);
    have_particles = 1;
    PartManager->NumPart = n;
    build_uniform_domain(&dd, 0);
    for(int64_t i = 0; i < n; i++) {
        memset(&P[i], 0, sizeof(P[i]));
        for(int k = 0; k < 3; k++) P[i].Pos[k] = pos[3 * i + k];
        P[i].Mass = mass[i]; P[i].Type = 1; P[i].ID = i; P[i].TopLeaf = 0;
    }
    static PetaPM pm;
#ifndef REF_PM_SHIM        /* with the gravpm shim (libref_dropin_pm.so) petapm.c is not linked at all */
    static int pm_module_ready = 0;
    if(!pm_module_ready) { petapm_module_init(omp_get_max_threads()); pm_module_ready = 1; }     /* run.c / main.c start-up */
#endif
    memset(&pm, 0, sizeof(pm));
    Cosmology CP;
    memset(&CP, 0, sizeof(CP));
    CP.Omega0 = 0.3;
    gravpm_init_periodic(&pm, BoxSize, Asmth, Nmesh, G);
    gravpm_force(&pm, &dd, &CP, Time, 3.085678e21, outdir, 0.01);
    for(int64_t i = 0; i < n; i++) {
        for(int k = 0; k < 3; k++) gravpm_out[3 * i + k] = P[i].GravPM[k];
        pot_out[i] = P[i].Potential;
    }
#ifndef REF_PM_SHIM
    petapm_destroy(&pm);
#endif
    return 0;
}

#ifndef REF_PM_SHIM
/* petapm_force_init + petapm_force_c2r + petapm_force_finish (petapm.c:263-362) driven the way MP-GenIC's
 * displacement_fields drives them (libgenic/zeldovich.c:150-229): one region round the particles (makeregion, :113-143), a
 * source spectrum handed in instead of gaussian_fill, up to 8 functions whose transfers have the forms of density_transfer /
 * disp_transfer (:276-313) with the k-dependent factor read from a table by the integer k2, and readouts that add
 * weight * mesh into out[f][i] (:338-359).  rho_k comes in natural [x][y][kz] order and is placed into the reference's
 * transposed Fourier region through its own offsets and strides (petapm.c:1092-1125: region axes are y, z, x). */
#define C2R_MAXF 8
static const double *c2r_table[C2R_MAXF];
static int c2r_kind[C2R_MAXF];
static double *c2r_out[C2R_MAXF];
static void c2r_transfer(int f, int64_t k2, int kpos[3], pfft_complex *value)
{
    if(k2) {
        if(c2r_kind[f] == 0) {
            const double fac = c2r_table[f][k2];
            value[0][0] *= fac;
            value[0][1] *= fac;
        } else {
            const double fac = c2r_table[f][k2] * kpos[c2r_kind[f] - 1];
            const double tmp = value[0][0];
            value[0][0] = -value[0][1] * fac;
            value[0][1] = tmp * fac;
        }
    }
}
#define C2R_FN(f) static void c2r_transfer_##f(PetaPM *pm, int64_t k2, int kpos[3], pfft_complex *value) { c2r_transfer(f, k2, kpos, value); } \
                  static void c2r_readout_##f(PetaPM *pm, int i, double *mesh, double weight) { c2r_out[f][i] += weight * mesh[0]; }
C2R_FN(0) C2R_FN(1) C2R_FN(2) C2R_FN(3) C2R_FN(4) C2R_FN(5) C2R_FN(6) C2R_FN(7)
static PetaPMRegion *c2r_makeregion(PetaPM *pm, PetaPMParticleStruct *pstruct, void *userdata, int *Nregions)
{
    PetaPMRegion *regions = (PetaPMRegion *) mymalloc2("Regions", sizeof(PetaPMRegion));
    double min[3] = {pm->BoxSize, pm->BoxSize, pm->BoxSize}, max[3] = {0, 0, 0};
    for(int64_t i = 0; i < PartManager->NumPart; i++)
        for(int k = 0; k < 3; k++) {
            if(min[k] > P[i].Pos[k]) min[k] = P[i].Pos[k];
            if(max[k] < P[i].Pos[k]) max[k] = P[i].Pos[k];
        }
    for(int k = 0; k < 3; k++) {
        regions[0].offset[k] = floor(min[k] / pm->BoxSize * pm->Nmesh - 1);
        regions[0].size[k] = ceil(max[k] / pm->BoxSize * pm->Nmesh + 2);
        regions[0].size[k] -= regions[0].offset[k];
    }
    petapm_region_init_strides(&regions[0]);
    *Nregions = 1;
    return regions;
}
int ref_petapm_c2r(int64_t n, const double *pos, double BoxSize, int Nmesh, const double *rho_k, int nfunc, const int *kind,
                   const double *tables, int64_t nk2, double *out)
{
    if(nfunc > C2R_MAXF) return 1;
    free_all();
    particle_alloc_memory(PartManager, BoxSize, n);
    have_particles = 1;
    PartManager->NumPart = n;
    for(int64_t i = 0; i < n; i++) {
        memset(&P[i], 0, sizeof(P[i]));
        for(int k = 0; k < 3; k++) P[i].Pos[k] = pos[3 * i + k];
        P[i].Mass = 1; P[i].Type = 1; P[i].ID = i;
    }
    static PetaPM pm;
    static int pm_module_ready = 0;
    if(!pm_module_ready) { petapm_module_init(omp_get_max_threads()); pm_module_ready = 1; }
    memset(&pm, 0, sizeof(pm));
    petapm_init(&pm, BoxSize, 1.25, Nmesh, 1.0, MPI_COMM_WORLD);          /* Asmth and G play no part in the c2r pass */
    PetaPMParticleStruct pstruct = {P, sizeof(P[0]), (char *) &P[0].Pos[0] - (char *) P, (char *) &P[0].Mass - (char *) P, NULL, NULL, (int) n};
    static void (*tf[C2R_MAXF])(PetaPM *, int64_t, int *, pfft_complex *) = {c2r_transfer_0, c2r_transfer_1, c2r_transfer_2, c2r_transfer_3,
                                                                               c2r_transfer_4, c2r_transfer_5, c2r_transfer_6, c2r_transfer_7};
    static void (*rf[C2R_MAXF])(PetaPM *, int, double *, double) = {c2r_readout_0, c2r_readout_1, c2r_readout_2, c2r_readout_3,
                                                                     c2r_readout_4, c2r_readout_5, c2r_readout_6, c2r_readout_7};
    static char names[C2R_MAXF][8];
    PetaPMFunctions functions[C2R_MAXF + 1];
    memset(functions, 0, sizeof(functions));
    for(int f = 0; f < nfunc; f++) {
        c2r_kind[f] = kind[f]; c2r_table[f] = tables + (size_t) f * nk2; c2r_out[f] = out + (size_t) f * n;
        for(int64_t i = 0; i < n; i++) c2r_out[f][i] = 0;
        snprintf(names[f], sizeof(names[f]), "F%d", f);
        functions[f].name = names[f]; functions[f].transfer = tf[f]; functions[f].readout = rf[f];
    }
    int Nregions = 0;
    PetaPMRegion *regions = petapm_force_init(&pm, c2r_makeregion, &pstruct, &Nregions, NULL);
    pfft_complex *rk = petapm_alloc_rhok(&pm);
    PetaPMRegion *fr = petapm_get_fourier_region(&pm);
    const int Nz = Nmesh / 2 + 1;
    for(int ix = 0; ix < Nmesh; ix++) for(int iy = 0; iy < Nmesh; iy++) for(int iz = 0; iz < Nz; iz++) {
        /* region axes: y, z, x */
        const int a = iy - fr->offset[0], b = iz - fr->offset[1], c = ix - fr->offset[2];
        if(a < 0 || b < 0 || c < 0 || a >= fr->size[0] || b >= fr->size[1] || c >= fr->size[2]) continue;
        const size_t ip = (size_t) a * fr->strides[0] + (size_t) b * fr->strides[1] + (size_t) c * fr->strides[2];
        const size_t in = ((size_t) ix * Nmesh + iy) * Nz + iz;
        rk[ip][0] = rho_k[2 * in]; rk[ip][1] = rho_k[2 * in + 1];
    }
    petapm_force_c2r(&pm, rk, regions, Nregions, functions);
    myfree(rk);
    myfree(regions);
    petapm_force_finish(&pm);
    petapm_destroy(&pm);
    return 0;
}
#endif
#endif

#ifdef REF_WITH_STEP
/* ---- Step loop: the reference's own drift.c, timestep.c and timebinmgr.c (compiled unmodified) on
 * one rank -- drift_all_particles, update_lastactive_drift, build_active_particles,
 * build_active_sublist, apply_half_kick / apply_hydro_half_kick / apply_PM_half_kick,
 * update_kick_times, hierarchical_gravity_accelerations and hierarchical_gravity_and_timesteps
 * (with the reference's own force_tree_active_moments + grav_short_tree underneath), driven the way
 * run.c:355-800 drives them.  Stand-ins only for what needs GSL: the background cosmology
 * (cosmology.c:64-86 reduces to the flat matter + Lambda form below when radiation, curvature and
 * neutrinos are off) and the kick/drift integrals of timefac.c:12-73, whose integrands are
 * restated and integrated by Gauss-Legendre panels instead of gsl_integration_qag. */
#include <libgadget/drift.h>
#include <libgadget/timebinmgr.h>
#include <libgadget/timefac.h>
#include <libgadget/physconst.h>

static Cosmology stepCP;
static double ts_par[6];
double hubble_function(const Cosmology *CP, double a)
{
    return CP->Hubble * sqrt(CP->Omega0 / (a * a * a) + CP->OmegaLambda);
}
static double step_integrand(int kind, double a)
{
    const double h = hubble_function(&stepCP, a);
    if(kind == 0) return 1 / (h * a * a * a);                      /* drift_integ     timefac.c:12-17 */
    if(kind == 1) return 1 / (h * a * a);                          /* gravkick_integ  timefac.c:20-26 */
    return 1 / (h * pow(a, 3 * GAMMA_MINUS1) * a);                 /* hydrokick_integ timefac.c:30-38 */
}
static double step_factor(int kind, inttime_t t0, inttime_t t1)    /* get_exact_factor timefac.c:41-56 */
{
    if(t0 == t1) return 0;
    static const double gx[4] = {0.1834346424956498, 0.5255324099163290, 0.7966664774136267, 0.9602898564975363};
    static const double gw[4] = {0.3626837833783620, 0.3137066458778873, 0.2223810344533745, 0.1012285362903763};
    const double a0 = exp(loga_from_ti(t0)), a1 = exp(loga_from_ti(t1));
    const int NP = 64;
    double sum = 0;
    for(int p = 0; p < NP; p++) {
        const double lo = a0 + (a1 - a0) * p / NP, hi = a0 + (a1 - a0) * (p + 1) / NP;
        const double c = 0.5 * (lo + hi), hw = 0.5 * (hi - lo);
        for(int k = 0; k < 4; k++) sum += gw[k] * hw * (step_integrand(kind, c - hw * gx[k]) + step_integrand(kind, c + hw * gx[k]));
    }
    return sum;
}
double get_exact_drift_factor(Cosmology *CP, inttime_t t0, inttime_t t1) { return step_factor(0, t0, t1); }
double get_exact_gravkick_factor(Cosmology *CP, inttime_t t0, inttime_t t1) { return step_factor(1, t0, t1); }
double get_exact_hydrokick_factor(Cosmology *CP, inttime_t t0, inttime_t t1) { return step_factor(2, t0, t1); }
double ref_step_factor(int kind, int64_t t0, int64_t t1) { return step_factor(kind, t0, t1); }
int BHGetRepositionEnabled(void) { return 0; }
#ifndef REF_WITH_PM
double get_omega_nu(const _omega_nu *const omnu, const double a) { return 0; }
#endif
double param_get_double(ParameterSet *ps, const char *name)
{
    static const char *keys[6] = {"ErrTolIntAccuracy", "MaxGasVel", "MaxSizeTimestep", "MinSizeTimestep", "MaxRMSDisplacementFac", "CourantFac"};
    if(!strcmp(name, "ArtBulkViscConst")) return hp_visc;
    if(!strcmp(name, "DensityContrastLimit")) return hp_contrast;
    for(int k = 0; k < 6; k++) if(!strcmp(name, keys[k])) return ts_par[k];
    endrun(1, "ref_driver: unexpected parameter %s\n", name); return 0;
}
int param_get_int(ParameterSet *ps, const char *name)
{
    if(!strcmp(name, "DensityIndependentSphOn")) return hp_di;
    if(!strcmp(name, "ForceEqualTimesteps")) return 0;
    endrun(1, "ref_driver: unexpected parameter %s\n", name); return 0;
}
char *param_get_string(ParameterSet *ps, const char *name) { return NULL; }

/* Timeline and parameters: set_sync_params_test + setup_sync_points (timebinmgr.c:151-330) and
 * set_timestep_params (timestep.c:51-67).  tspar = {ErrTolIntAccuracy, MaxGasVel, MaxSizeTimestep,
 * MinSizeTimestep, MaxRMSDisplacementFac, CourantFac}.  Call once, before any particles. */
int ref_step_init(double TimeIC, double TimeMax, int nout, double *outtimes, double Omega0, double OmegaBaryon,
                  double Hubble, double G, const double *tspar)
{
    static int done = 0;
    if(done) return 1;
    free_all();
    memset(&stepCP, 0, sizeof(stepCP));
    stepCP.Omega0 = Omega0; stepCP.OmegaLambda = 1 - Omega0; stepCP.OmegaBaryon = OmegaBaryon; stepCP.OmegaCDM = Omega0 - OmegaBaryon;
    stepCP.Hubble = Hubble; stepCP.GravInternal = G; stepCP.HubbleParam = 0.7;
    stepCP.RhoCrit = 3 * Hubble * Hubble / (8 * M_PI * G);
    memcpy(ts_par, tspar, sizeof(ts_par));
    set_timestep_params(NULL);
    set_sync_params_test(nout, outtimes);
    setup_sync_points(&stepCP, TimeIC, TimeMax, 0.0, 0);
    done = 1;
    return 0;
}
/* the integer timeline as the reference computes it (timebinmgr.c:380-447) */
double ref_loga_from_ti(int64_t ti) { return loga_from_ti(ti); }
int64_t ref_ti_from_loga(double loga) { return ti_from_loga(loga); }
int64_t ref_dti_from_dloga(double dloga, int64_t Ti_Current) { return dti_from_dloga(dloga, Ti_Current); }
double ref_dloga_from_dti(int64_t dti, int64_t Ti_Current) { return dloga_from_dti(dti, Ti_Current); }

static DriftKickTimes stepT;
static ActiveParticles stepAct;
static int step_act_built = 0;
static PetaPM step_pm;
static double step_rho0;

/* scal = {mintimebin, maxtimebin, mingravtimebin, Ti_Current, PM_length, PM_start, PM_kick} */
void ref_step_set_times(const int64_t *scal, const int64_t *ti_kick, const int64_t *ti_last)
{
    stepT.mintimebin = scal[0]; stepT.maxtimebin = scal[1]; stepT.mingravtimebin = scal[2];
    stepT.Ti_Current = scal[3]; stepT.PM_length = scal[4]; stepT.PM_start = scal[5]; stepT.PM_kick = scal[6];
    for(int b = 0; b <= TIMEBINS; b++) { stepT.Ti_kick[b] = ti_kick[b]; stepT.Ti_lastactivedrift[b] = ti_last[b]; }
}
void ref_step_get_times(int64_t *scal, int64_t *ti_kick, int64_t *ti_last)
{
    scal[0] = stepT.mintimebin; scal[1] = stepT.maxtimebin; scal[2] = stepT.mingravtimebin;
    scal[3] = stepT.Ti_Current; scal[4] = stepT.PM_length; scal[5] = stepT.PM_start; scal[6] = stepT.PM_kick;
    for(int b = 0; b <= TIMEBINS; b++) { ti_kick[b] = stepT.Ti_kick[b]; ti_last[b] = stepT.Ti_lastactivedrift[b]; }
}

/* Particles of any type; gas (type 0) gets an SPH slot in index order.  flags: bit 0 IsGarbage,
 * bit 1 Swallowed.  Any of the optional arrays may be NULL (zeros). */
int ref_step_set_particles(int64_t n, const double *pos, const double *vel, const float *mass, const unsigned char *type,
                           const unsigned char *flags, const double *fullacc, const double *gravpm,
                           const unsigned char *bin_grav, const unsigned char *bin_hydro, const double *hsml, const double *dthsml,
                           const double *hydroacc, const double *entropy, const double *dtentropy,
                           double BoxSize, int topdepth, int64_t ti_drift)
{
    if(step_act_built) { free_active_particles(&stepAct); step_act_built = 0; }
    free_all();
    particle_alloc_memory(PartManager, BoxSize, n);
    have_particles = 1;
    PartManager->NumPart = n;
    int64_t ngas = 0;
    for(int64_t i = 0; i < n; i++) ngas += (type[i] == 0);
    slots_init(0.01 * n, SlotsManager);
    slots_set_enabled(0, sizeof(struct sph_particle_data), SlotsManager);
    int64_t atleast[6] = {0};
    atleast[0] = ngas > 0 ? ngas : 1;
    slots_reserve(1, atleast, SlotsManager);
    slots_ready = 1;
    SlotsManager->info[0].size = ngas;
    build_uniform_domain(&dd, topdepth);
    int64_t pi = 0;
    for(int64_t i = 0; i < n; i++) {
        memset(&P[i], 0, sizeof(P[i]));
        for(int k = 0; k < 3; k++) {
            P[i].Pos[k] = pos[3 * i + k];
            P[i].Vel[k] = vel ? vel[3 * i + k] : 0;
            P[i].FullTreeGravAccel[k] = fullacc ? fullacc[3 * i + k] : 0;
            P[i].GravPM[k] = gravpm ? gravpm[3 * i + k] : 0;
        }
        P[i].Mass = mass[i]; P[i].Type = type[i]; P[i].ID = i;
        P[i].IsGarbage = flags ? (flags[i] & 1) : 0; P[i].Swallowed = flags ? ((flags[i] >> 1) & 1) : 0;
        P[i].TimeBinGravity = bin_grav ? bin_grav[i] : 0; P[i].TimeBinHydro = bin_hydro ? bin_hydro[i] : 0;
        P[i].Ti_drift = ti_drift;
        P[i].Hsml = hsml ? hsml[i] : 0; P[i].DtHsml = dthsml ? dthsml[i] : 0;
        P[i].TopLeaf = domain_get_topleaf(PEANO(P[i].Pos, BoxSize), &dd);
        if(type[i] == 0) {
            P[i].PI = pi++;
            memset(&SPHP(i), 0, sizeof(struct sph_particle_data));
            for(int k = 0; k < 3; k++) SPHP(i).HydroAccel[k] = hydroacc ? hydroacc[3 * i + k] : 0;
            SPHP(i).Entropy = entropy ? entropy[i] : 0; SPHP(i).DtEntropy = dtentropy ? dtentropy[i] : 0;
        }
    }
    return 0;
}
void ref_step_get(double *pos, double *vel, double *hsml, double *entropy, unsigned char *bin_grav, double *fullacc, int64_t *ti_drift)
{
    const int64_t n = PartManager->NumPart;
    for(int64_t i = 0; i < n; i++) {
        for(int k = 0; k < 3; k++) {
            if(pos) pos[3 * i + k] = P[i].Pos[k];
            if(vel) vel[3 * i + k] = P[i].Vel[k];
            if(fullacc) fullacc[3 * i + k] = P[i].FullTreeGravAccel[k];
        }
        if(hsml) hsml[i] = P[i].Hsml;
        if(entropy) entropy[i] = P[i].Type == 0 ? SPHP(i).Entropy : 0;
        if(bin_grav) bin_grav[i] = P[i].TimeBinGravity;
        if(ti_drift) ti_drift[i] = P[i].Ti_drift;
    }
}
/* drift_all_particles (drift.c:84-102): returns the drift factor it used */
double ref_step_drift(int64_t ti0, int64_t ti1, const double *shift)
{
    drift_all_particles(ti0, ti1, &stepCP, shift);
    /* the tree code needs the top leaf of the new position (domain_maintain would do this) */
    for(int64_t i = 0; i < PartManager->NumPart; i++) P[i].TopLeaf = domain_get_topleaf(PEANO(P[i].Pos, PartManager->BoxSize), &dd);
    return get_exact_drift_factor(&stepCP, ti0, ti1);
}
/* update_lastactive_drift + build_active_particles (timestep.c:860-871,1334-1431).  list_out (size
 * NumPart) receives the list; returns NumActiveParticle, or -1 - NumPart when the list is implicit
 * (PM step: ActiveParticle == NULL).  counts = {NumActiveParticle, NumActiveGravity, NumActiveHydro}. */
int64_t ref_step_build_active(int *list_out, int64_t *counts)
{
    if(step_act_built) { free_active_particles(&stepAct); step_act_built = 0; }
    update_lastactive_drift(&stepT);
    stepAct = init_empty_active_particles(PartManager);
    build_active_particles(&stepAct, &stepT, 0, get_atime(stepT.Ti_Current), PartManager);
    step_act_built = 1;
    counts[0] = stepAct.NumActiveParticle; counts[1] = stepAct.NumActiveGravity; counts[2] = stepAct.NumActiveHydro;
    if(!stepAct.ActiveParticle) return -1 - stepAct.NumActiveParticle;
    for(int64_t q = 0; q < stepAct.NumActiveParticle; q++) list_out[q] = stepAct.ActiveParticle[q];
    return stepAct.NumActiveParticle;
}
/* build_active_sublist (timestep.c:1435-1478) of the current list */
ActiveParticles build_active_sublist(const ActiveParticles *act, const int maxtimebin, const inttime_t Ti_Current);
int64_t ref_step_sublist(int maxtimebin, int *list_out)
{
    ActiveParticles sub = build_active_sublist(&stepAct, maxtimebin, stepT.Ti_Current);
    const int64_t na = sub.NumActiveParticle;
    for(int64_t q = 0; q < na; q++) list_out[q] = sub.ActiveParticle[q];
    myfree(sub.ActiveParticle);
    return na;
}
/* kind 0 apply_half_kick, 1 apply_hydro_half_kick, 2 apply_PM_half_kick, 3 update_kick_times
 * (timestep.c:874-994,215-235) on the current active list */
void ref_step_kick(int kind, double atime)
{
    if(kind == 0) apply_half_kick(&stepAct, &stepCP, &stepT, atime);
    else if(kind == 1) apply_hydro_half_kick(&stepAct, &stepCP, &stepT, atime);
    else if(kind == 2) apply_PM_half_kick(&stepCP, &stepT);
    else update_kick_times(&stepT);
}
/* find_hydro_timesteps (timestep.c:617-738) on the current active list; maxsig[n] = SphP[].MaxSignalVel
 * (entries of non-gas particles ignored).  Returns the bad-step count; bin_hydro_out[n] = P[].TimeBinHydro. */
int ref_step_hydro_timesteps(const double *maxsig, double atime, int first, unsigned char *bin_hydro_out)
{
    const int64_t n = PartManager->NumPart;
    for(int64_t i = 0; i < n; i++) if(P[i].Type == 0) SPHP(i).MaxSignalVel = maxsig[i];
    const int bad = find_hydro_timesteps(&stepAct, &stepT, atime, &stepCP, first);
    for(int64_t i = 0; i < n; i++) bin_hydro_out[i] = P[i].TimeBinHydro;
    return bad;
}
/* find_timesteps (timestep.c:739-853), the SplitGravityTimestepsOn = 0 path: one bin per particle from the gravity
 * and (gas) hydro criteria.  Returns the bad-step count; both bin arrays [n] out. */
int ref_step_find_timesteps(const double *maxsig, double atime, double asmth, int first, unsigned char *bin_grav_out, unsigned char *bin_hydro_out)
{
    const int64_t n = PartManager->NumPart;
    for(int64_t i = 0; i < n; i++) if(P[i].Type == 0) SPHP(i).MaxSignalVel = maxsig[i];
    const int bad = find_timesteps(&stepAct, &stepT, atime, 2, &stepCP, asmth, first);
    for(int64_t i = 0; i < n; i++) { bin_grav_out[i] = P[i].TimeBinGravity; bin_hydro_out[i] = P[i].TimeBinHydro; }
    return bad;
}
/* Short-range gravity parameters of the hierarchy (as ref_grav_short_tree above) */
void ref_step_set_gravity(double G, int Nmesh, double Asmth, double ErrTolForceAcc, double BHOpeningAngle,
                          double MaxBHOpeningAngle, int TreeUseBH, double Rcut, double GravitySoftening)
{
    memset(&step_pm, 0, sizeof(step_pm));
    step_pm.BoxSize = PartManager->BoxSize; step_pm.Asmth = Asmth; step_pm.Nmesh = Nmesh; step_pm.G = G;
    step_pm.CellSize = step_pm.BoxSize / Nmesh;
    gravshort_fill_ntab(SHORTRANGE_FORCE_WINDOW_TYPE_EXACT, Asmth);
    struct gravshort_tree_params tp = {0};
    tp.ErrTolForceAcc = ErrTolForceAcc; tp.BHOpeningAngle = BHOpeningAngle; tp.MaxBHOpeningAngle = MaxBHOpeningAngle;
    tp.TreeUseBH = TreeUseBH; tp.Rcut = Rcut; tp.FractionalGravitySoftening = GravitySoftening;
    set_gravshort_treepar(tp);
    gravshort_set_softenings(1.0);
    step_rho0 = stepCP.Omega0 * 3 * stepCP.Hubble * stepCP.Hubble / (8 * M_PI * stepCP.GravInternal);      /* run.c:544 */
}
double ref_step_softening(void) { return FORCE_SOFTENING(); }
/* One pass of run.c:441-795 for collisionless particles with HierarchicalGravity on, PM force held
 * fixed (P[].GravPM as loaded): advance Ti_Current to the next kick, drift, active list,
 * hierarchical_gravity_accelerations, update_kick_times, PM half kick, hierarchical_gravity_and_timesteps,
 * update_kick_times, PM half kick.  first != 0 starts from the state as loaded (no advance, no drift),
 * like NumCurrentTiStep == 0.  Returns the reference's bad-timestep count. */
static const double *step_maxsig;      /* non-NULL: gas takes part, hydro accelerations and signal velocities held fixed */
void ref_step_set_maxsig(const double *maxsig) { step_maxsig = maxsig; }
int ref_step_advance(int first, int64_t *nactive_out)
{
    const inttime_t Ti_Last = stepT.Ti_Current;
    if(!first) stepT.Ti_Current = find_next_kick(stepT.Ti_Current, stepT.mintimebin);
    const double atime = get_atime(stepT.Ti_Current);
    const int is_PM = is_PM_timestep(&stepT);
    const double zero[3] = {0, 0, 0};
    if(!first) ref_step_drift(Ti_Last, stepT.Ti_Current, zero);
    int64_t counts[3];
    int *tmp = (int *) malloc(sizeof(int) * (PartManager->NumPart + 1));
    ref_step_build_active(tmp, counts);
    free(tmp);
    if(nactive_out) { nactive_out[0] = counts[0]; nactive_out[1] = counts[1]; nactive_out[2] = is_PM; }
    if(step_maxsig) apply_hydro_half_kick(&stepAct, &stepCP, &stepT, atime);       /* run.c:498-499, after density + hydro_force */
    struct grav_accel_store GravAccel = {0};
    GravAccel.nstore = PartManager->NumPart;
    GravAccel.GravAccel = (MyFloat (*)[3]) mymalloc2("GravAccel", GravAccel.nstore * sizeof(GravAccel.GravAccel[0]));
    if(counts[1] > 0) hierarchical_gravity_accelerations(&stepAct, &step_pm, &dd, GravAccel, &stepT, 0, &stepCP, NULL);     /* run.c:533 */
    update_kick_times(&stepT);
    if(is_PM) apply_PM_half_kick(&stepCP, &stepT);
    int bad = 0;
    if(counts[1] > 0) bad = hierarchical_gravity_and_timesteps(&stepAct, &step_pm, &dd, GravAccel, &stepT, atime, 0, 2, &stepCP, NULL);
    else if(GravAccel.GravAccel) myfree(GravAccel.GravAccel);
    if(step_maxsig) {                                                               /* run.c:767-773 */
        for(int64_t i = 0; i < PartManager->NumPart; i++) if(P[i].Type == 0) SPHP(i).MaxSignalVel = step_maxsig[i];
        bad += find_hydro_timesteps(&stepAct, &stepT, atime, &stepCP, first);
        apply_hydro_half_kick(&stepAct, &stepCP, &stepT, atime);
    }
    update_kick_times(&stepT);
    if(is_PM) apply_PM_half_kick(&stepCP, &stepT);
    free_active_particles(&stepAct);
    step_act_built = 0;
    return bad;
}
/* The same pass with SplitGravityTimestepsOn = 0 (run.c:541-560,749-793): one full tree, grav_short_tree for the active
 * particles, closing half kick, find_timesteps, opening half kick. */
int ref_step_advance_nonsplit(int first, double asmth, int64_t *nactive_out)
{
    const inttime_t Ti_Last = stepT.Ti_Current;
    if(!first) stepT.Ti_Current = find_next_kick(stepT.Ti_Current, stepT.mintimebin);
    const double atime = get_atime(stepT.Ti_Current);
    const int is_PM = is_PM_timestep(&stepT);
    const double zero[3] = {0, 0, 0};
    if(!first) ref_step_drift(Ti_Last, stepT.Ti_Current, zero);
    int64_t counts[3];
    int *tmp = (int *) malloc(sizeof(int) * (PartManager->NumPart + 1));
    ref_step_build_active(tmp, counts);
    free(tmp);
    if(nactive_out) { nactive_out[0] = counts[0]; nactive_out[1] = counts[1]; nactive_out[2] = is_PM; }
    ForceTree T = {0};
    force_tree_full(&T, &dd, 0, NULL);
    grav_short_tree(&stepAct, &step_pm, &T, NULL, step_rho0, stepT.Ti_Current);
    force_tree_free(&T);
    apply_half_kick(&stepAct, &stepCP, &stepT, atime);
    update_kick_times(&stepT);
    if(is_PM) apply_PM_half_kick(&stepCP, &stepT);
    const int bad = find_timesteps(&stepAct, &stepT, atime, 2, &stepCP, asmth, first);
    apply_half_kick(&stepAct, &stepCP, &stepT, atime);
    update_kick_times(&stepT);
    if(is_PM) apply_PM_half_kick(&stepCP, &stepT);
    free_active_particles(&stepAct);
    step_act_built = 0;
    return bad;
}
#endif

void ref_shutdown(void) { free_all(); }
