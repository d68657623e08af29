/* oracle.h -- CPU restatement of the MP-Gadget force-step algorithms.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (mp-gadget_b200/,
 * include/) may include, link or call this.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * use it, and only as the checker / CPU baseline.
 *
 * Every function cites the reference file:line whose arithmetic it restates
 * (paths relative to the MP-Gadget tree).  Parity status:
 *   - tree + short-range walk: PINNED against the reference's own C compiled
 *     from /root/reference (oracle/_ref, tests/test_golden.py) and
 *     against committed outputs of it (tests/golden/);
 *   - SPH density / hydro (synchronised and mixed time bins): PINNED against the
 *     reference's own density.c / hydra.c (tests/golden/ref_sph*.npz);
 *   - PM: PINNED against the reference's own petapm.c / gravpm.c / powerspectrum.c
 *     run on one rank (oracle/_ref/libref_pm.so, tests/golden/ref_pm.npz).  PFFT,
 *     the one third-party piece (fetched at build time, absent offline), is
 *     replaced there by plain DFTs in its single-rank layout
 *     (oracle/pfft_standin.c); every other line that runs is the reference's.
 */
#ifndef ORACLE_H
#define ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct oracle_node {
    int32_t sibling;     /* DFS index of next node when not descending, -1 at end */
    int32_t father;      /* DFS index of parent, -1 for root */
    int32_t firstchild;  /* DFS index of first child (internal) else -1 */
    int32_t nocc;        /* particle leaf: number of particles; internal: -1 */
    int32_t part[8];     /* particle indices of a leaf (ascending insertion order) */
    int32_t toplevel;    /* forced top-tree node */
    int32_t level;
    double len;
    double center[3];
    double cofm[3];
    double mass;
    double hmax;
} oracle_node;

typedef struct oracle_tree {
    oracle_node *nodes;
    int64_t numnodes;
    int64_t numparticles;
    double BoxSize;
} oracle_tree;

/* forcetree.c:196-270 (force_tree_build), :727-860 (create_nodes),
 * :1017-1104 (moments).  hsml may be NULL (hmax = 0). */
int oracle_tree_build(oracle_tree *t, const double *pos, const float *mass,
                      const uint8_t *type, const double *hsml, int64_t n, double BoxSize,
                      int mask, const int32_t *active, int64_t nactive, int toplevel_depth);
int oracle_tree_build_top(oracle_tree *t, const double *pos, const float *mass,
                          const uint8_t *type, const double *hsml, int64_t n, double BoxSize,
                          int mask, const int32_t *active, int64_t nactive, int toplevel_depth,
                          const int32_t *top_daughter, int32_t ntop);
void oracle_tree_free(oracle_tree *t);

typedef struct oracle_gravshort_params {
    double ErrTolForceAcc, BHOpeningAngle, MaxBHOpeningAngle;
    int32_t TreeUseBH, pad_;
    double Rcut, GravitySoftening, rho0;
} oracle_gravshort_params;

typedef struct oracle_walk_counts {
    int32_t nodes_accepted, nodes_opened, nodes_discarded, particles;
} oracle_walk_counts;

/* gravshort-tree.c:96-154,253-379; gravshort.h:47-96; gravity.c:54-66.
 * oldacc[n][3] = FullTreeGravAccel + GravPM (may be NULL = 0). */
int oracle_grav_short_tree(const oracle_tree *t, const double *pos, const float *mass, int64_t n,
                           const oracle_gravshort_params *par, double G, int Nmesh, double Asmth,
                           const double *oldacc, const int32_t *active, int64_t nactive,
                           int full_particle_tree,
                           double *accel_out, double *pot_out, oracle_walk_counts *counts_out);

/* petapm.c:955-1006 + :1138-1144.  mesh[Nmesh^3] (x slowest) must be zeroed by
 * the caller; icell_out[n][3] optional. */
void oracle_pm_deposit(const double *pos, const float *mass, int64_t n, double BoxSize, int Nmesh,
                       double *mesh, int32_t *icell_out);
/* petapm.c:1092-1132 + gravpm.c:383-454.  rhok: complex [Nmesh][Nmesh][Nmesh/2+1]
 * interleaved re,im, index order (x, y, z). In place. */
void oracle_pm_potential_transfer(double *rhok, int Nmesh, double BoxSize, double Asmth, double G);
void oracle_pm_power(const double *rhok, int Nmesh, double *power, double *kk, int64_t *nmodes, double *norm);
/* gravpm.c:458-489. dim 0/1/2 = x/y/z. out may alias nothing (copy then scale). */
void oracle_pm_force_transfer(const double *potk, double *out, int Nmesh, double BoxSize, int dim);
/* petapm.c:955-1006 + gravpm.c:499-510: out[i*ostride] += sum_c w_c mesh[c]. */
void oracle_pm_readout(const double *mesh, const double *pos, int64_t n, double BoxSize, int Nmesh,
                       double *out, int64_t ostride);
/* Direct periodic-image summation of tests/test_gravity.c:38-150 (grav_force,
 * force_direct) -- the reference's own ground truth for TreePM accuracy. */
void oracle_direct_sum(const double *pos, const float *mass, int64_t n, double BoxSize, double G,
                       double softening_h, int repeat, double *accel_out);

/* ---- SPH (synchronised step by default; mixed time bins / active sets via oracle_sph_set_mixed) ---- */
typedef struct oracle_sph_params {
    int32_t KernelType;               /* 1 cubic, 2 quintic, 4 quartic: densitykernel.h:20-24 */
    int32_t DensityIndependentSphOn;  /* hydra.c:37-53 */
    double DensityResolutionEta, MaxNumNgbDeviation, MinGasHsml;     /* density.c:30-51,264-265 */
    double ArtBulkViscConst, DensityContrastLimit;                   /* hydra.c:37-48 */
    double gravkick, hydrokick, pmkick;   /* kick_factor_data of the common time bin, density.c:114-132 */
    double dloga_pred;                    /* dloga of SPH_EntVarPred, density.c:74 */
    double drift;                         /* drifts[bin], hydra.c:178-186 */
    double dloga_bin;                     /* get_dloga_for_bin, hydra.c:271,463 */
    double atime, hubble;                 /* hydra.c:219-223 */
} oracle_sph_params;

double oracle_sph_desnumngb(int kerneltype, double eta);
#define ORACLE_NBINS 47          /* TIMEBINS + 1, timebinmgr.h:13 */
void oracle_sph_set_mixed(const uint8_t *bin_grav, const uint8_t *bin_hydro, const double *tab, const uint8_t *active);
void oracle_set_init_hsml(const oracle_tree *t, const float *mass, const uint8_t *type, int64_t n,
                          int kerneltype, double eta, double MeanGasSeparation, double *hsml);
int oracle_density(oracle_tree *t, const double *pos, const float *mass, const uint8_t *type, int64_t n,
                   const oracle_sph_params *sp, int update_hsml, int DoEgyDensity,
                   const double *vel, const double *fullacc, const double *gravpm, const double *hydroacc,
                   const double *entropy, const double *dtentropy,
                   double *hsml, double *density, double *egywtdensity, double *dhsmlfac,
                   double *divvel, double *curlvel, double *dthsml, double *numngb_out, int32_t *ninteract, int32_t *niter_out,
                   double *entvarpred_out);
int oracle_hydro(const oracle_tree *t, const double *pos, const float *mass, const uint8_t *type, int64_t n,
                 const oracle_sph_params *sp,
                 const double *vel, const double *fullacc, const double *gravpm, const double *hydroacc_in,
                 const double *entropy, const double *dtentropy_in,
                 const double *hsml, const double *density, const double *egywtdensity, const double *dhsmlfac,
                 const double *divvel, const double *curlvel,
                 double *acc_out, double *dtentropy_out, double *maxsignalvel_out, int32_t *ninteract);

/* ---- step loop (oracle_step.c): integer timeline, drift, active lists, kicks, hierarchical gravity ---- */
#define ORACLE_TIMEBINS 46       /* timebinmgr.h:13 */
typedef struct oracle_timeline { int64_t nsync; const double *loga; } oracle_timeline;     /* SyncPoints[].loga, timebinmgr.c:18 */
typedef struct oracle_cosmo { double Omega0, OmegaBaryon, Hubble, G; } oracle_cosmo;        /* flat matter + Lambda */
typedef struct oracle_times {                                                               /* DriftKickTimes timestep.h:10-26 */
    int32_t mintimebin, maxtimebin, mingravtimebin, pad_;
    int64_t Ti_kick[ORACLE_TIMEBINS + 1], Ti_lastactivedrift[ORACLE_TIMEBINS + 1];
    int64_t Ti_Current, PM_length, PM_start, PM_kick;
} oracle_times;
typedef struct oracle_step_params {                                                         /* TimestepParams timestep.c:21-47 */
    double ErrTolIntAccuracy, MaxGasVel, MaxSizeTimestep, MinSizeTimestep, MaxRMSDisplacementFac;
    double softening;                                                                       /* FORCE_SOFTENING() */
    double CourantFac;
} oracle_step_params;
double oracle_loga_from_ti(const oracle_timeline *tl, int64_t ti);
int64_t oracle_ti_from_loga(const oracle_timeline *tl, double loga);
int64_t oracle_dti_from_dloga(const oracle_timeline *tl, double dloga, int64_t Ti_Current);
double oracle_dloga_from_dti(const oracle_timeline *tl, int64_t dti, int64_t Ti_Current);
int oracle_is_timebin_active(int bin, int64_t ti);
double oracle_step_factor(const oracle_cosmo *c, const oracle_timeline *tl, int kind, int64_t t0, int64_t t1);
int64_t oracle_drift(int64_t n, double *pos, const double *vel, const uint8_t *type, const uint8_t *flags,
                     double *hsml, const double *dthsml, double ddrift, const double *shift, double BoxSize);
int64_t oracle_build_active(int64_t n, const uint8_t *type, const uint8_t *flags, const uint8_t *bin_grav, const uint8_t *bin_hydro,
                            int64_t Ti_Current, int is_pm, int64_t nhydro_slots, int32_t *list_out, int64_t *counts, int64_t *bincounts);
int64_t oracle_active_sublist(const int32_t *list, int64_t nlist, const uint8_t *flags, const uint8_t *bin_grav,
                              int maxtimebin, int64_t Ti_Current, int32_t *out);
void oracle_half_kick(const int32_t *list, int64_t nlist, const uint8_t *type, const uint8_t *flags, const uint8_t *bin_grav,
                      const uint8_t *bin_hydro, double *vel, const double *fullacc, const double *hydroacc, double *entropy,
                      const double *dtentropy, const double *gravkick, const double *hydrokick, const double *dt_entr,
                      int64_t Ti_Current, double atime, double MaxGasVel, int hydro_only);
void oracle_pm_kick(int64_t n, const uint8_t *flags, double *vel, const double *gravpm, double Fgravkick);
void oracle_grav_kick(const int32_t *list, int64_t nlist, const uint8_t *flags, double *vel, const double *acc, double gravkick);
void oracle_update_kick_times(oracle_times *t);
void oracle_update_lastactive_drift(oracle_times *t);
double oracle_gravity_dloga(const double *acc, const double *gravpm, double atime, double hubble, double ErrTolIntAccuracy, double softening);
int64_t oracle_convert_timestep(const oracle_timeline *tl, double dloga, int64_t dti_max, int64_t Ti_Current, double MinSizeTimestep);
int oracle_gravity_timebin(const oracle_timeline *tl, const double *acc, const double *gravpm, const oracle_step_params *sp,
                           double atime, double hubble, int64_t dti_max, int64_t Ti_Current, int largest_active);
int64_t oracle_pm_timestep_ti(const oracle_timeline *tl, const oracle_cosmo *c, const oracle_step_params *sp, const oracle_times *t,
                              int64_t n, const double *vel, const float *mass, const uint8_t *type, const uint8_t *flags,
                              double atime, int FastParticleType, double asmth);
int oracle_hydro_timebins(const oracle_timeline *tl, const oracle_step_params *sp, oracle_times *t, const int32_t *list, int64_t nlist,
                          const uint8_t *type, const uint8_t *flags, const double *hsml, const double *dthsml, const double *maxsig,
                          const uint8_t *bin_grav, uint8_t *bin_hydro, double atime, double hubble);
int oracle_find_timesteps(const oracle_timeline *tl, const oracle_cosmo *c, const oracle_step_params *sp, oracle_times *t, int64_t n,
                          const int32_t *list, int64_t nlist, const uint8_t *type, const uint8_t *flags, const float *mass, const double *vel,
                          const double *fullacc, const double *gravpm, const double *hsml, const double *dthsml, const double *maxsig,
                          uint8_t *bin_grav, uint8_t *bin_hydro, int is_pm, double atime, int FastParticleType, double asmth);
int oracle_hier_accelerations(const oracle_timeline *tl, const oracle_cosmo *c, const oracle_step_params *sp, oracle_gravshort_params *gp,
                              oracle_times *t, int64_t n, const double *pos, const float *mass, const uint8_t *type, const uint8_t *flags,
                              double *vel, double *fullacc, const double *gravpm, const uint8_t *bin_grav,
                              const int32_t *act, int64_t nact, int64_t ngrav,
                              double G, int Nmesh, double Asmth, double BoxSize, double *store);
int oracle_hier_timesteps(const oracle_timeline *tl, const oracle_cosmo *c, const oracle_step_params *sp, oracle_gravshort_params *gp,
                          oracle_times *t, int64_t n, const double *pos, const float *mass, const uint8_t *type, const uint8_t *flags,
                          double *vel, double *fullacc, const double *gravpm, uint8_t *bin_grav,
                          const int32_t *act, int64_t nact, int64_t ngrav, int is_pm,
                          double G, int Nmesh, double Asmth, double BoxSize, double atime, int FastParticleType,
                          const double *store, int64_t *info);

/* ---- domain keys (oracle_domain.c): utils/peano.c:108-129, peano.h:15-21, domain.h:71-78 ---- */
int oracle_peano_tables(uint8_t rank[48][8], uint8_t next[48][8]);
uint64_t oracle_peano_key(int x, int y, int z, int bits);
void oracle_peano_keys(const double *pos, int64_t n, double BoxSize, uint64_t *keys);
/* the top tree, domain.c:826-1395; node layout = struct local_topnode_data (domain.c:60-70) */
typedef struct oracle_topnode { uint64_t StartKey; int32_t Shift, Daughter, Parent, pad_; int64_t Count, Cost; } oracle_topnode;
int oracle_toptree_local(const uint64_t *sorted_keys, const int64_t *cost, int64_t nsample, oracle_topnode *tree, int32_t *size, int32_t maxnodes);
void oracle_toptree_truncate(oracle_topnode *tree, int32_t *size, int64_t countlimit, int64_t costlimit);
int oracle_toptree_merge(oracle_topnode *treeA, int32_t *sizeA, const oracle_topnode *treeB, int32_t maxnodes);
int oracle_toptree_global_refine(oracle_topnode *tree, int32_t *size, int32_t maxnodes, int64_t countlimit, int64_t costlimit);
int32_t oracle_toptree_leaves(const oracle_topnode *tree, int32_t size, int32_t *leaf_out);
int64_t oracle_exchange_plan(int64_t n, const uint8_t *type, const uint8_t *flags, const int32_t *topleaf, int32_t nleaf,
                             const int32_t *task_of_leaf, int32_t ntask, int32_t thistask, int32_t *list_out, int64_t *togo, int64_t *ngarbage);
void oracle_leaf_counts(const int32_t *topleaf, const uint8_t *flags, int64_t n, int32_t nleaf, int64_t *counts);
int oracle_domain_assign_balanced(int ntask, int32_t nleaf, const int64_t *cost, int nseg_per_task, int32_t *task);
void oracle_topleaf(const uint64_t *keys, int64_t n, const int32_t *daughter, const uint64_t *startkey, const int32_t *shift,
                    const int32_t *leaf, int32_t *out);

/* ---- friends-of-friends, primary linking (oracle_fof.c): fof.c:366-470,540-579 ---- */
int oracle_fof_primary(int64_t n, const double *pos, const int64_t *ids, const uint8_t *type, int mask, double BoxSize, double ll,
                       int64_t *minid);

#ifdef __cplusplus
}
#endif
#endif
