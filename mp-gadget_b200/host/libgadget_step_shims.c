/* libgadget_step_shims.c -- reference-signature binding of the device-resident step loop
 * (include/b200force.h, b200_step_*).
 *
 * The functions replaced here live in libgadget/timestep.c and drift.c next to code that is NOT
 * replaced (hydro / black-hole time steps, statistics), so this is not a whole-object swap like the
 * other shim files: the maintainer links MP-Gadget with
 *
 *   -Wl,--wrap=set_timestep_params,--wrap=drift_all_particles,--wrap=build_active_particles
 *   -Wl,--wrap=apply_half_kick,--wrap=apply_hydro_half_kick,--wrap=apply_PM_half_kick
 *   -Wl,--wrap=hierarchical_gravity_accelerations,--wrap=hierarchical_gravity_and_timesteps
 *
 * and adds this file; every call from run.c (run.c:355-800) then lands in the __wrap_ functions
 * below with unchanged signatures, and timestep.c / drift.c stay in the link for everything else.
 *
 * Coherence.  Between calls the host arrays (P[], SphP[]) stay the master copy: each function
 * uploads what it reads and writes back what it changed (positions after a drift, velocities after a
 * kick, time bins after the time-step assignment).  That keeps every non-replaced consumer of P[]
 * correct; a host that has moved all its consumers to the device calls b200_step_* directly and
 * skips the copies (mp-gadget_b200/steploop.py is that loop).
 *
 * Not supported (endrun): black-hole particles (drag terms in the kicks, repositioning in the drift).
 *
 * Compiled and exercised by tests/test_step_emul.py (against the CPU emulation build of
 * csrc/steploop.cu) and tests/test_step_gpu.py (against libb200force.so) through
 * oracle/_ref/libref_dropin_step*.so: the reference's own ref_step_advance loop runs with these
 * functions in place of its own.
 */
#include <mpi.h>
#include <math.h>
#include <string.h>
#include <stdlib.h>

#include <libgadget/utils/endrun.h>
#include <libgadget/utils/mymalloc.h>
#include <libgadget/utils/system.h>
#include <libgadget/partmanager.h>
#include <libgadget/slotsmanager.h>
#include <libgadget/timestep.h>
#include <libgadget/timebinmgr.h>
#include <libgadget/timefac.h>
#include <libgadget/drift.h>
#include <libgadget/gravity.h>
#include <libgadget/petapm.h>
#include <libgadget/cosmology.h>
#include <libgadget/walltime.h>

#include "../../include/b200force.h"

b200_ctx *b200_shim_context(void);
#define B200_CK(call) do { if((call) != 0) endrun(1, "b200: %s\n", b200_last_error(b200_shim_context())); } while(0)

/* ---- parameters: timestep.c keeps them in a file-static struct, so read them alongside ---- */
static struct { double ErrTolIntAccuracy, MaxGasVel, MaxSizeTimestep, MinSizeTimestep, MaxRMSDisplacementFac, CourantFac; } StepPar;
void __real_set_timestep_params(ParameterSet *ps);
void __wrap_set_timestep_params(ParameterSet *ps)                          /* timestep.c:51-67 */
{
    __real_set_timestep_params(ps);
    int ThisTask;
    MPI_Comm_rank(MPI_COMM_WORLD, &ThisTask);
    if(ThisTask == 0) {
        StepPar.ErrTolIntAccuracy = param_get_double(ps, "ErrTolIntAccuracy");
        StepPar.MaxGasVel = param_get_double(ps, "MaxGasVel");
        StepPar.MaxSizeTimestep = param_get_double(ps, "MaxSizeTimestep");
        StepPar.MinSizeTimestep = param_get_double(ps, "MinSizeTimestep");
        StepPar.MaxRMSDisplacementFac = param_get_double(ps, "MaxRMSDisplacementFac");
        StepPar.CourantFac = param_get_double(ps, "CourantFac");
    }
    MPI_Bcast(&StepPar, sizeof(StepPar), MPI_BYTE, 0, MPI_COMM_WORLD);
}

/* ---- state transfer ---- */
static int64_t NumGasSlots(void) { return SlotsManager->info[0].size + SlotsManager->info[5].size; }

/* P[] / SphP[] -> device.  The per-particle arrays are staged in one mymalloc block, freed before return. */
static void step_upload(void)
{
    b200_ctx *ctx = b200_shim_context();
    const int64_t n = PartManager->NumPart;
    B200_CK(b200_set_particles_aos(ctx, P, n, NULL));
    const size_t m = (size_t) (n > 0 ? n : 1);
    double *buf = (double *) mymalloc("b200_step_stage", (3 * 4 + 4) * m * sizeof(double) + 2 * m);
    double *vel = buf, *fullacc = vel + 3 * m, *gravpm = fullacc + 3 * m, *hydroacc = gravpm + 3 * m;
    double *hsml = hydroacc + 3 * m, *dthsml = hsml + m, *entropy = dthsml + m, *dtentropy = entropy + m;
    uint8_t *bg = (uint8_t *) (dtentropy + m), *bh = bg + m;
    int have_gas = 0;
    #pragma omp parallel for reduction(|: have_gas)
    for(int64_t i = 0; i < n; i++) {
        for(int k = 0; k < 3; k++) {
            vel[3 * i + k] = P[i].Vel[k]; fullacc[3 * i + k] = P[i].FullTreeGravAccel[k]; gravpm[3 * i + k] = P[i].GravPM[k];
            hydroacc[3 * i + k] = 0;
        }
        hsml[i] = P[i].Hsml; dthsml[i] = P[i].DtHsml; entropy[i] = 0; dtentropy[i] = 0;
        bg[i] = P[i].TimeBinGravity; bh[i] = P[i].TimeBinHydro;
        if(P[i].IsGarbage || P[i].Swallowed) continue;
        if(P[i].Type == 5) endrun(1, "b200 step shims: black-hole particles are not supported\n");
        if(P[i].Type == 0) {
            have_gas = 1;
            for(int k = 0; k < 3; k++) hydroacc[3 * i + k] = SPHP(i).HydroAccel[k];
            entropy[i] = SPHP(i).Entropy; dtentropy[i] = SPHP(i).DtEntropy;
        }
    }
    b200_step_state st;
    memset(&st, 0, sizeof(st));
    st.vel = vel; st.fullacc = fullacc; st.gravpm = gravpm; st.bin_grav = bg; st.bin_hydro = bh;
    if(have_gas) { st.hsml = hsml; st.dthsml = dthsml; st.hydroacc = hydroacc; st.entropy = entropy; st.dtentropy = dtentropy; }
    st.BoxSize = PartManager->BoxSize;
    B200_CK(b200_step_set_state(ctx, &st));
    myfree(buf);
}
/* device -> P[] / SphP[]: what = bit 0 positions + Hsml, bit 1 velocities + entropy, bit 2 gravity bins, bit 3 FullTreeGravAccel */
static void step_download(int what)
{
    b200_ctx *ctx = b200_shim_context();
    const int64_t n = PartManager->NumPart;
    const size_t m = (size_t) (n > 0 ? n : 1);
    double *buf = (double *) mymalloc("b200_step_stage", (3 * 3 + 2) * m * sizeof(double) + m);
    double *pos = buf, *vel = pos + 3 * m, *fullacc = vel + 3 * m, *hsml = fullacc + 3 * m, *entropy = hsml + m;
    uint8_t *bg = (uint8_t *) (entropy + m);
    b200_step_state_out o;
    memset(&o, 0, sizeof(o));
    if(what & 1) { o.pos = pos; o.hsml = hsml; }
    if(what & 2) { o.vel = vel; o.entropy = entropy; }
    if(what & 4) o.bin_grav = bg;
    if(what & 8) o.fullacc = fullacc;
    B200_CK(b200_step_get_state(ctx, &o));
    #pragma omp parallel for
    for(int64_t i = 0; i < n; i++) {
        const int gas = P[i].Type == 0 && !P[i].IsGarbage && !P[i].Swallowed;
        if(what & 1) { for(int k = 0; k < 3; k++) P[i].Pos[k] = pos[3 * i + k]; if(gas) P[i].Hsml = hsml[i]; }
        if(what & 2) { for(int k = 0; k < 3; k++) P[i].Vel[k] = vel[3 * i + k]; if(gas) SPHP(i).Entropy = entropy[i]; }
        if(what & 4) P[i].TimeBinGravity = bg[i];
        if(what & 8) for(int k = 0; k < 3; k++) P[i].FullTreeGravAccel[k] = fullacc[3 * i + k];
    }
    myfree(buf);
}

/* ---- drift.c:84-102 ---- */
void __wrap_drift_all_particles(inttime_t ti0, inttime_t ti1, Cosmology *CP, const double random_shift[3])
{
    if(ti1 < ti0) endrun(12, "Trying to reverse time: ti0=%ld ti1=%ld\n", (long) ti0, (long) ti1);
    const double ddrift = get_exact_drift_factor(CP, ti0, ti1);
    step_upload();
    int64_t nbad = 0;
    if(b200_step_drift(b200_shim_context(), ddrift, random_shift, &nbad))
        endrun(5, "b200: %s\n", b200_last_error(b200_shim_context()));
    step_download(1);
    #pragma omp parallel for
    for(int64_t i = 0; i < PartManager->NumPart; i++) P[i].Ti_drift = ti1;
    walltime_measure("/Drift");
}

/* The per-bin occupancy log of build_active_particles (print_timebin_statistics, timestep.c:1491-1585, file-static there):
 * the counts come from the device histogram (TimeBinCountType of b200_step_build_active), summed over the ranks as the
 * reference does, and the root prints the step header and one line per occupied bin. */
static void shim_timebin_statistics(const DriftKickTimes *times, int NumCurrentTiStep, const int64_t *loc, double Time)
{
    static int64_t tot[6 * (TIMEBINS + 1)];
    MPI_Reduce((void *) loc, tot, 6 * (TIMEBINS + 1), MPI_INT64, MPI_SUM, 0, MPI_COMM_WORLD);
    int ThisTask;
    MPI_Comm_rank(MPI_COMM_WORLD, &ThisTask);
    if(ThisTask != 0) return;
    int64_t nforce = 0, npart = 0, ntype[6] = {0};
    for(int bin = 0; bin <= TIMEBINS; bin++) {
        int64_t inbin = 0;
        for(int ty = 0; ty < 6; ty++) { inbin += tot[(TIMEBINS + 1) * ty + bin]; ntype[ty] += tot[(TIMEBINS + 1) * ty + bin]; }
        npart += inbin;
        if(is_timebin_active(bin, times->Ti_Current)) nforce += inbin;
    }
    const double dloga = get_dloga_for_bin(times->mintimebin, times->Ti_Current);
    message(0, "Begin Step %d, Time: %g (%lx), Redshift: %g, Nf = %014ld, Systemstep: %g, Dloga: %g, status: %s\n",
            NumCurrentTiStep, Time, times->Ti_Current, 1.0 / Time - 1, nforce, dloga * Time, dloga, is_PM_timestep(times) ? "PM-Step" : "");
    message(0, "TotNumPart: %013ld SPH %013ld BH %010ld STAR %013ld \n", npart, ntype[0], ntype[5], ntype[4]);
    message(0, "Occupied: % 12ld % 12ld % 12ld % 12ld % 12ld % 12ld dt\n", 0L, 1L, 2L, 3L, 4L, 5L);
    int64_t act[6] = {0}, acttot = 0;
    for(int bin = TIMEBINS; bin >= 0; bin--) {
        const int64_t *c = tot + bin;
        const int64_t inbin = c[0] + c[TIMEBINS + 1] + c[2 * (TIMEBINS + 1)] + c[3 * (TIMEBINS + 1)] + c[4 * (TIMEBINS + 1)] + c[5 * (TIMEBINS + 1)];
        if(inbin == 0) continue;
        const int on = is_timebin_active(bin, times->Ti_Current);
        message(0, " %c bin=%2d % 12ld % 12ld % 12ld % 12ld % 12ld % 12ld %6g\n", on ? 'X' : ' ', bin, c[0], c[TIMEBINS + 1], c[2 * (TIMEBINS + 1)],
                c[3 * (TIMEBINS + 1)], c[4 * (TIMEBINS + 1)], c[5 * (TIMEBINS + 1)], get_dloga_for_bin(bin, times->Ti_Current));
        if(on) { acttot += inbin; for(int ty = 0; ty < 6; ty++) act[ty] += c[ty * (TIMEBINS + 1)]; }
    }
    message(0, "               -----------------------------------\n");
    message(0, "Total:    % 12ld % 12ld % 12ld % 12ld % 12ld % 12ld  Sum:% 14ld\n", act[0], act[1], act[2], act[3], act[4], act[5], acttot);
}

/* ---- timestep.c:1334-1431 ---- */
void __wrap_build_active_particles(ActiveParticles *act, const DriftKickTimes *const times, const int NumCurrentTiStep, const double Time,
                                   const struct part_manager_type *const PartManager)
{
    b200_ctx *ctx = b200_shim_context();
    step_upload();
    int64_t counts[3];
    static int64_t bincounts[6 * (TIMEBINS + 1)];
    const int is_pm = is_PM_timestep(times);
    B200_CK(b200_step_build_active(ctx, times->Ti_Current, is_pm, NumGasSlots(), counts, bincounts));
    act->NumActiveParticle = counts[0]; act->NumActiveGravity = counts[1]; act->NumActiveHydro = counts[2];
    act->Particles = PartManager->Base;
    act->ActiveParticle = NULL;
    if(!is_pm) {                           /* the host list other modules walk: same storage rule as timestep.c:1419-1421 */
        act->MaxActiveParticle = act->NumActiveParticle + PartManager->MaxPart - PartManager->NumPart;
        act->ActiveParticle = (int *) mymalloc("ActiveParticle", sizeof(int) * (act->MaxActiveParticle > 0 ? act->MaxActiveParticle : 1));
        int64_t nout = 0;
        B200_CK(b200_step_get_active(ctx, 0, act->ActiveParticle, &nout));
        if(nout != act->NumActiveParticle) endrun(1, "b200 step shims: active list length %ld != %ld\n", (long) nout, (long) act->NumActiveParticle);
    }
    walltime_measure("/Timeline/Active");
    shim_timebin_statistics(times, NumCurrentTiStep, bincounts, Time);
}

/* ---- timestep.c:874-994 ---- */
static void kick_tables(const DriftKickTimes *times, Cosmology *CP, double *gravkick, double *hydrokick, double *dt_entr)
{
    for(int bin = 0; bin <= TIMEBINS; bin++) {
        gravkick[bin] = hydrokick[bin] = 0;
        dt_entr[bin] = dloga_from_dti(dti_from_timebin(bin) / 2, times->Ti_Current);
        if(bin < times->mintimebin || !is_timebin_active(bin, times->Ti_Current)) continue;
        const inttime_t newkick = times->Ti_kick[bin] + dti_from_timebin(bin) / 2;
        gravkick[bin] = get_exact_gravkick_factor(CP, times->Ti_kick[bin], newkick);
        hydrokick[bin] = get_exact_hydrokick_factor(CP, times->Ti_kick[bin], newkick);
    }
}
/* Other host modules (density, hydro, gravpm_force, sub-grid physics) may have changed P[] / SphP[] since the last
 * call: start every function from the host's state and the host's list. */
static void fresh_state(const ActiveParticles *act)
{
    step_upload();
    B200_CK(b200_step_set_active(b200_shim_context(), act ? act->ActiveParticle : NULL, act ? act->NumActiveParticle : 0));
}
static void half_kick(const ActiveParticles *act, Cosmology *CP, DriftKickTimes *times, const double atime, int hydro_only)
{
    fresh_state(act);
    double gravkick[TIMEBINS + 1], hydrokick[TIMEBINS + 1], dt_entr[TIMEBINS + 1];
    kick_tables(times, CP, gravkick, hydrokick, dt_entr);
    B200_CK(b200_step_half_kick(b200_shim_context(), gravkick, hydrokick, dt_entr, times->Ti_Current, atime, StepPar.MaxGasVel, hydro_only));
    step_download(2);
    walltime_measure("/Timeline/HalfKick/Short");
}
void __wrap_apply_half_kick(const ActiveParticles *act, Cosmology *CP, DriftKickTimes *times, const double atime) { half_kick(act, CP, times, atime, 0); }
void __wrap_apply_hydro_half_kick(const ActiveParticles *act, Cosmology *CP, DriftKickTimes *times, const double atime) { half_kick(act, CP, times, atime, 1); }
void __wrap_apply_PM_half_kick(Cosmology *CP, DriftKickTimes *times)
{
    const inttime_t tistart = times->PM_kick, tiend = tistart + times->PM_length / 2;
    fresh_state(NULL);
    B200_CK(b200_step_pm_kick(b200_shim_context(), get_exact_gravkick_factor(CP, tistart, tiend)));
    step_download(2);
    times->PM_kick = tiend;
    walltime_measure("/Timeline/HalfKick/Long");
}

/* ---- timestep.c:296-598 ---- */
static Cosmology *KickCP;
static double kick_cb(void *user, int64_t t0, int64_t t1) { return get_exact_gravkick_factor(KickCP, t0, t1); }
static double *SyncLoga;
static int64_t NSync;
static void sync_table(void)              /* SyncPoints[] is file-static in timebinmgr.c: walk it through its accessors */
{
    int64_t cap = 64, k = 0;
    SyncLoga = (double *) realloc(SyncLoga, cap * sizeof(double));
    SyncLoga[k++] = loga_from_ti(0);
    inttime_t ti = 0;
    SyncPoint *sp;
    while((sp = find_next_sync_point(ti)) != NULL) {
        if(k == cap) { cap *= 2; SyncLoga = (double *) realloc(SyncLoga, cap * sizeof(double)); }
        SyncLoga[k++] = sp->loga;
        ti = sp->ti;
    }
    NSync = k;
}
static void times_to_b200(b200_step_times *b, const DriftKickTimes *t)
{
    memset(b, 0, sizeof(*b));
    b->mintimebin = t->mintimebin; b->maxtimebin = t->maxtimebin; b->mingravtimebin = t->mingravtimebin;
    for(int i = 0; i <= TIMEBINS; i++) { b->Ti_kick[i] = t->Ti_kick[i]; b->Ti_lastactivedrift[i] = t->Ti_lastactivedrift[i]; }
    b->Ti_Current = t->Ti_Current; b->PM_length = t->PM_length; b->PM_start = t->PM_start; b->PM_kick = t->PM_kick;
}
static void times_from_b200(DriftKickTimes *t, const b200_step_times *b)
{
    t->mintimebin = b->mintimebin; t->maxtimebin = b->maxtimebin; t->mingravtimebin = b->mingravtimebin;
    t->PM_length = b->PM_length; t->PM_start = b->PM_start; t->PM_kick = b->PM_kick;
}
static struct gravshort_tree_params HierTree;   /* carries TreeUseBH > 1 -> 0 between the calls like the reference's static */
static void hier_setup(b200_step_params *sp, b200_gravshort_params *gp, PetaPM *pm, Cosmology *CP, int FastParticleType)
{
    sync_table();
    memset(sp, 0, sizeof(*sp));
    sp->ErrTolIntAccuracy = StepPar.ErrTolIntAccuracy; sp->MaxSizeTimestep = StepPar.MaxSizeTimestep; sp->MinSizeTimestep = StepPar.MinSizeTimestep;
    sp->MaxRMSDisplacementFac = StepPar.MaxRMSDisplacementFac; sp->CourantFac = StepPar.CourantFac;
    sp->softening = FORCE_SOFTENING();
    sp->omega_type[0] = sp->omega_type[4] = sp->omega_type[5] = CP->OmegaBaryon;       /* timestep.c:1251-1263 */
    sp->omega_type[1] = sp->omega_type[3] = CP->OmegaCDM;
    sp->omega_type[2] = get_omega_nu(&CP->ONu, 1);
    sp->RhoCrit = CP->RhoCrit;
    sp->FastParticleType = FastParticleType;
    sp->sync_loga = SyncLoga; sp->nsync = NSync;
    KickCP = CP; sp->gravkick_factor = kick_cb;
    HierTree = get_gravshort_treepar();
    memset(gp, 0, sizeof(*gp));
    gp->ErrTolForceAcc = HierTree.ErrTolForceAcc; gp->BHOpeningAngle = HierTree.BHOpeningAngle; gp->MaxBHOpeningAngle = HierTree.MaxBHOpeningAngle;
    gp->TreeUseBH = HierTree.TreeUseBH; gp->Rcut = HierTree.Rcut; gp->GravitySoftening = FORCE_SOFTENING() / 2.8;
    gp->rho0 = CP->Omega0 * 3 * CP->Hubble * CP->Hubble / (8 * M_PI * CP->GravInternal);
    B200_CK(b200_pm_init(b200_shim_context(), PartManager->BoxSize, pm->Asmth, pm->Nmesh, pm->G));
}
static void hier_finish(const b200_gravshort_params *gp)
{
    if(gp->TreeUseBH != HierTree.TreeUseBH) { HierTree.TreeUseBH = gp->TreeUseBH; set_gravshort_treepar(HierTree); }   /* gravshort-tree.c:150-151 */
}

int __wrap_hierarchical_gravity_accelerations(const ActiveParticles *act, PetaPM *pm, DomainDecomp *ddecomp, struct grav_accel_store StoredGravAccel,
                                              DriftKickTimes *times, int HybridNuGrav, Cosmology *CP, const char *EmergencyOutputDir)
{
    fresh_state(act);
    if(HybridNuGrav) endrun(1, "b200 step shims: hybrid neutrino tracers are not supported\n");
    b200_ctx *ctx = b200_shim_context();
    b200_step_params sp; b200_gravshort_params gp; b200_step_times bt;
    hier_setup(&sp, &gp, pm, CP, 2);
    times_to_b200(&bt, times);
    B200_CK(b200_step_hier_accelerations(ctx, &sp, &gp, &bt, act->NumActiveGravity));
    hier_finish(&gp);
    if(StoredGravAccel.GravAccel)          /* the host copy other modules read (cooling_and_starformation, run.c:667) */
        B200_CK(b200_step_get_store(ctx, (double *) StoredGravAccel.GravAccel));
    step_download(2 | 8);
    return 0;
}
int __wrap_hierarchical_gravity_and_timesteps(const ActiveParticles *act, PetaPM *pm, DomainDecomp *ddecomp, struct grav_accel_store StoredGravAccel,
                                              DriftKickTimes *times, const double atime, int HybridNuGrav, int FastParticleType, Cosmology *CP,
                                              const char *EmergencyOutputDir)
{
    fresh_state(act);
    if(HybridNuGrav) endrun(1, "b200 step shims: hybrid neutrino tracers are not supported\n");
    b200_ctx *ctx = b200_shim_context();
    b200_step_params sp; b200_gravshort_params gp; b200_step_times bt;
    hier_setup(&sp, &gp, pm, CP, FastParticleType);
    times_to_b200(&bt, times);
    B200_CK(b200_step_set_store(ctx, (const double *) StoredGravAccel.GravAccel));   /* NULL: FullTreeGravAccel, timestep.c:355-358 */
    int64_t info[3] = {0, 0, 0};
    B200_CK(b200_step_hier_timesteps(ctx, &sp, &gp, &bt, act->NumActiveGravity, is_PM_timestep(times), atime, hubble_function(CP, atime), info));
    hier_finish(&gp);
    times_from_b200(times, &bt);
    if(StoredGravAccel.GravAccel) myfree(StoredGravAccel.GravAccel);       /* timestep.c:419-420: this call owns and frees it */
    step_download(2 | 4);
    walltime_measure("/Timeline/HierGrav/Kick");
    return (int) info[2];
}
