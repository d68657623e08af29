/* b200force.h -- C-ABI of the B200-native TreePM + SPH force engine.
 *
 * Drop-in boundary for the MP-Gadget force step.  The
 * reference has no FFI layer: its modules call each other through C headers in
 * libgadget.a, so the replacement is link-time: a maintainer swaps
 * gravpm.o / gravshort-tree.o / forcetree.o for thin shims with the reference
 * signatures (mp-gadget_b200/host/libgadget_shims.c) that call
 * the entry points below.  Each entry point cites the reference interface it
 * replaces (paths relative to the MP-Gadget tree).
 *
 * Conventions
 *  - plain C types only; every pointer is a HOST pointer unless the function
 *    name ends in _dev (then it is a device pointer on the context's GPU);
 *  - every function returns 0 on success, non-zero on failure; the message is
 *    available from b200_last_error().  The reference's convention is
 *    endrun()/MPI_Abort (libgadget/utils/endrun.c:138-153); the shims map a
 *    non-zero status to endrun(1, "%s", b200_last_error(ctx));
 *  - particle arrays are indexed by the caller's particle index (the index into
 *    the reference's global P[] array, libgadget/partmanager.h:73-85);
 *  - there is NO CPU fallback: without a CUDA device b200_ctx_create fails.  The b200_domain_toptree_* and
 *    b200_domain_assign_balanced functions take no context: they are the host part of the domain decomposition
 *    (sequential integer work on a tree of a few hundred nodes, as in the reference) and run anywhere.
 *
 * Sections: particles and context; PM (gravpm.c, petapm.c) incl. the generic inverse pass (petapm_force_c2r, libgenic/zeldovich.c);
 * tree build and short-range walk (forcetree.c, gravshort-tree.c); whole force step; SPH density and hydro (density.c, hydra.c);
 * multi-GPU building blocks; step loop (drift.c, timestep.c); domain decomposition (peano.c, domain.c, exchange.c);
 * friends-of-friends primary linking (fof.c); timings.
 */
#ifndef B200FORCE_H
#define B200FORCE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_ABI_VERSION 1

typedef struct b200_ctx b200_ctx;

/* Byte layout of the caller's array-of-structs particle record.  Defaults
 * (b200_default_particle_layout) are the offsets of `struct particle_data`
 * (libgadget/partmanager.h:9-71, 160 bytes, no -DDEBUG). */
typedef struct b200_particle_layout {
    int64_t stride;          /* sizeof(struct particle_data) = 160 */
    int32_t off_pos;         /* double Pos[3]                    0 */
    int32_t off_mass;        /* float  Mass                     28 */
    int32_t off_flags;       /* bitfield byte (IsGarbage bit0, Swallowed bit1) 36 */
    int32_t off_type;        /* unsigned char Type              39 */
    int32_t off_vel;         /* double Vel[3]                   40 */
    int32_t off_fulltreeacc; /* double FullTreeGravAccel[3]     64 */
    int32_t off_gravpm;      /* double GravPM[3]                88 */
    int32_t off_hsml;        /* double Hsml                    120 */
    int32_t off_potential;   /* double Potential               152 */
    int32_t off_pi;          /* int PI (slot index)             32 */
    int32_t off_timebin_hydro;   /* unsigned char TimeBinHydro   37 */
    int32_t off_timebin_gravity; /* unsigned char TimeBinGravity 38 */
} b200_particle_layout;

void b200_default_particle_layout(b200_particle_layout *layout);

/* ---- context ------------------------------------------------------------ */

/* Create an engine bound to CUDA device `device`.  Fails (non-zero, *out=NULL)
 * when no CUDA device is present. */
int b200_ctx_create(b200_ctx **out, int device);
void b200_ctx_destroy(b200_ctx *ctx);
const char *b200_last_error(const b200_ctx *ctx);
int b200_abi_version(void);
/* Number of kernel launches issued by this context since creation. */
int64_t b200_kernel_launches(const b200_ctx *ctx);

/* ---- particle ingest ------------------------------------------------------
 * Replaces the reference's direct reads of the global P[] array
 * (PartManager->Base, libgadget/partmanager.h:73-85).  The engine keeps a
 * struct-of-arrays copy in HBM until the next call. */

/* From the caller's AoS records (host memory). */
int b200_set_particles_aos(b200_ctx *ctx, const void *particles, int64_t n,
                           const b200_particle_layout *layout);
/* From separate arrays (host memory): pos[n][3] f64, mass[n] f32,
 * type[n] u8 (NULL = all type 1), oldacc[n][3] f64 = FullTreeGravAccel+GravPM
 * of the previous step (NULL = zeros). */
int b200_set_particles_soa(b200_ctx *ctx, const double *pos, const float *mass,
                           const uint8_t *type, const double *oldacc, int64_t n);
/* Same, device-resident inputs (no host traffic). */
int b200_set_particles_soa_dev(b200_ctx *ctx, const double *pos, const float *mass,
                               const uint8_t *type, const double *oldacc, int64_t n);
/* Set |FullTreeGravAccel + GravPM| inputs of the relative opening criterion
 * from the accelerations the engine itself computed in the last
 * b200_pm_force + b200_grav_short_tree pair (grav_get_abs_accel,
 * libgadget/gravshort.h:69-86) -- keeps the step device-resident. */
int b200_oldacc_from_last_step(b200_ctx *ctx);

/* ---- PM long-range force ---------------------------------------------------
 * b200_pm_init  replaces gravpm_init_periodic (libgadget/gravity.h:40,
 *               gravpm.c:51-54) = petapm_init (petapm.c:104-223).
 * b200_pm_force replaces gravpm_force (gravity.h:55, gravpm.c:60-119) =
 *               petapm_force (petapm.h:133, petapm.c:364-379) with gravpm's
 *               callback set {Potential, ForceX, ForceY, ForceZ}
 *               (gravpm.c:32-39): CIC deposit, forward FFT, potential_transfer
 *               (gravpm.c:383-454), gradient (gravpm.c:458-489), readout
 *               (gravpm.c:499-510). */
int b200_pm_init(b200_ctx *ctx, double BoxSize, double Asmth, int Nmesh, double G);
/* The three scalars grav_short_tree reads from its PetaPM argument (CellSize = BoxSize/Nmesh, Asmth, G;
 * gravshort-tree.c:101-102, gravshort.h:47-67) for a host that keeps the PM step elsewhere: no mesh is allocated. */
int b200_walk_set_mesh(b200_ctx *ctx, double BoxSize, double Asmth, int Nmesh, double G);
/* gravpm_out[n][3] (P[i].GravPM) and potential_out[n] (the PM contribution
 * added to P[i].Potential) for every particle; either may be NULL. */
int b200_pm_force(b200_ctx *ctx, double *gravpm_out, double *potential_out);
int b200_pm_force_dev(b200_ctx *ctx, double *gravpm_out, double *potential_out);
/* How pfft_execute_dft_r2c -> potential_transfer -> pfft_execute_dft_c2r (petapm.c:305,326-344) run for the
 * mesh b200_pm_init set up: 1 = the engine's five shared-memory passes with the Green's function inside the
 * x pass (mesh sizes 2^a 3^b 5^c up to 1614), 0 = cuFFT D2Z / Z2D around a separate Green's-function kernel
 * (other sizes, or B200_PM_FFT=cufft in the environment at b200_pm_init), -1 = no mesh. */
int b200_pm_transform_kind(b200_ctx *ctx);
/* The generic inverse pass of the PM engine: petapm_force_c2r (libgadget/petapm.c:326-362) for
 * callers that bring their own source spectrum and transfer functions -- MP-GenIC's displacement_fields
 * (libgenic/zeldovich.c:150-253: Density, DispX/Y/Z, VelX/Y/Z) and the radius filters of petapm_reion_c2r
 * (petapm.c:416-577).  For every function j: spectrum_j = transfer_j(rho_k) (pm_apply_transfer_function,
 * petapm.c:1092-1132), C2R (unnormalised, like pfft's backward plan), out_j[i] = sum over the 8 CIC cells of particle i
 * of weight * mesh (the readout_* callbacks, zeldovich.c:338-359; particles = the set given to b200_set_particles_*).
 * rho_k: host, complex [Nmesh][Nmesh][Nmesh/2+1], x slowest (the reference keeps the same modes in PFFT's transposed
 * [y][z][x] order).  A transfer function is radial in the INTEGER wave number: table[k2], k2 = kx^2 + ky^2 + kz^2 in
 * [0, 3 (Nmesh/2)^2], filled by the caller with the reference's own formula -- density_transfer:
 * exp(-k2 / Nmesh^2) DeltaSpec(k) / sqrt(L^3) (zeldovich.c:276-289); disp_transfer: DeltaSpec(k) or dlogGrowth(k) times
 * 1 / (2 pi sqrt(L) k2) (:291-313) -- so the device applies exactly the factor the host function would.
 *   kind 0:        value *= table[k2]
 *   kind 1, 2, 3:  fac = table[k2] * k_axis (x, y, z);  (re, im) <- (-im fac, re fac)        (disp_transfer's swap)
 * The k2 = 0 mode passes unchanged (`if(k2)`).  Needs the engine's own transform passes (b200_pm_transform_kind == 1). */
typedef struct b200_pm_function {
    int kind;
    const double *table;      /* host, 3 (Nmesh/2)^2 + 1 entries */
    double *out;              /* host, n entries */
} b200_pm_function;
int b200_pm_c2r_readout(b200_ctx *ctx, const double *rho_k, int nfunc, const b200_pm_function *functions);
/* Matter power spectrum side effect of gravpm_force: potential_transfer calls
 * powerspectrum_add_mode on every density mode before scaling it
 * (gravpm.c:330-361,440).  b200_pm_set_power(ctx, 1) makes the next b200_pm_force
 * accumulate the same sums in the Green's-function pass; b200_pm_get_power
 * returns the RAW sums of the nbins = Nmesh logarithmic bins
 * (Power[] = sum w |rho_k|^2 / window^2, kk[] = sum w |k|, Nmodes[] = sum w,
 * *norm = |rho_0|^2), i.e. the state of struct _powerspectrum
 * (libgadget/powerspectrum.h:8-26) just before powerspectrum_sum
 * (powerspectrum.c:56-92), which the caller applies. */
int b200_pm_set_power(b200_ctx *ctx, int on);
int b200_pm_get_power(b200_ctx *ctx, int nbins, double *power, double *kk, int64_t *nmodes, double *norm);
/* Parity hooks: integer CIC cell of every particle (iCell of pm_iterate_one,
 * petapm.c:976-980) icell_out[n][3]; and a copy of the real-space mesh
 * (density after deposit if which==0, potential after the inverse FFT if
 * which==1), mesh_out[Nmesh^3], x slowest. */
int b200_pm_cell_index(b200_ctx *ctx, int32_t *icell_out);
int b200_pm_copy_mesh(b200_ctx *ctx, int which, double *mesh_out);

/* ---- octree ------------------------------------------------------------------
 * Replaces force_tree_full / force_tree_active_moments /
 * force_tree_rebuild_mask / force_tree_calc_moments / force_tree_free
 * (libgadget/forcetree.h:117-148, forcetree.c:110-183,196-270). */
typedef struct b200_tree_info {
    int64_t numnodes;      /* ForceTree.numnodes */
    int64_t numparticles;  /* ForceTree.NumParticles */
    int32_t maxdepth;      /* deepest level (root = 0) */
    int32_t overfull_leaves; /* leaves at the key-depth limit holding > 8 particles */
    double  root_mass;
} b200_tree_info;

/* mask: bit t set = include particle type t (ALLMASK = 63, forcetree.h:22-27).
 * active: particle indices to insert (NULL = all n particles), as
 *         ActiveParticles.ActiveParticle (libgadget/timestep.h:29-38).
 * toplevel_depth: all nodes down to this level are forced to exist and are
 *         never pruned (the replicated domain top-tree,
 *         forcetree.c:654-687,869-934); 0 = the one-leaf domain. */
int b200_tree_build(b200_ctx *ctx, double BoxSize, int mask, const int32_t *active,
                    int64_t nactive, int toplevel_depth, b200_tree_info *info);
void b200_tree_free(b200_ctx *ctx);

/* Parity hook: export the tree in depth-first order (the order of the
 * reference's sibling/first-child walk).  Arrays have info.numnodes entries;
 * any may be NULL.  sibling/firstchild are DFS positions (-1 = none);
 * nocc = particle count of a particle leaf, -1 for an internal node;
 * leafpart[numnodes][8] original particle indices (unused = -1). */
int b200_tree_export(b200_ctx *ctx, double *center /*[.][3]*/, double *len,
                     double *cofm /*[.][3]*/, double *mass, double *hmax,
                     int32_t *sibling, int32_t *firstchild, int32_t *nocc,
                     int32_t *leafpart);

/* ---- short-range tree gravity ----------------------------------------------
 * Replaces grav_short_tree (libgadget/gravity.h:58, gravshort-tree.c:96-154)
 * = treewalk_run (treewalk.c:801-902) with visit force_treeev_shortrange
 * (gravshort-tree.c:253-379), fill grav_short_copy, postprocess
 * grav_short_postprocess (gravshort.h:47-96).  Parameters mirror
 * struct gravshort_tree_params (gravity.h:9-22) + GravShortPriv
 * (gravshort.h:25-43). */
typedef struct b200_gravshort_params {
    double ErrTolForceAcc;
    double BHOpeningAngle;
    double MaxBHOpeningAngle;
    int32_t TreeUseBH;      /* !=0: Barnes-Hut angle only; 0: relative criterion */
    int32_t pad_;
    double Rcut;            /* TreeRcut, in units of Asmth*cellsize */
    double GravitySoftening;/* absolute Plummer-equivalent length (gravshort_set_softenings) */
    double rho0;            /* mean matter density, for the potential self-term */
} b200_gravshort_params;

/* Per-particle walk counters (parity on tree opening). */
typedef struct b200_walk_counts {
    int32_t nodes_accepted;   /* monopoles applied           (gravshort-tree.c:314-323) */
    int32_t nodes_opened;     /* internal nodes opened       (gravshort-tree.c:358-360) */
    int32_t nodes_discarded;  /* nodes discarded beyond Rcut (gravshort-tree.c:304-309) */
    int32_t particles;        /* particles of opened leaves = the reference's ninteractions (gravshort-tree.c:344-352,375) */
} b200_walk_counts;

/* active/nactive as above (NULL = all particles).  accel_out[n][3]: entry i is
 * written for every active particle i (G applied), untouched otherwise --
 * GravShortPriv.Accel (gravshort.h:41).  potential_out[n]: short-range
 * potential after grav_short_postprocess (gravshort.h:59-64), only meaningful
 * for full-particle trees; counts_out[n] optional. */
int b200_grav_short_tree(b200_ctx *ctx, const b200_gravshort_params *par,
                         const int32_t *active, int64_t nactive,
                         double *accel_out, double *potential_out,
                         b200_walk_counts *counts_out);
int b200_grav_short_tree_dev(b200_ctx *ctx, const b200_gravshort_params *par,
                             const int32_t *active, int64_t nactive,
                             double *accel_out, double *potential_out,
                             b200_walk_counts *counts_out);

/* ---- whole DM force step on the caller's AoS (the e2e path) ----------------
 * = gravpm_force + force_tree_full + grav_short_tree as run.c:519-548 does on a
 * PM step with SplitGravityTimestepsOn=0: reads P[], writes P[i].GravPM,
 * P[i].FullTreeGravAccel and P[i].Potential in place.
 * Only the bytes that matter cross PCIe, as strided copies: Pos..Type (40 B of the 160-byte record) before the PM step
 * and the tree build start, FullTreeGravAccel + GravPM (48 B) underneath them for the walk's opening criterion; the same
 * 48 B and Potential travel back group by group behind the walk.  A layout that does not fit (unaligned fields, spans
 * wider than 0.7 of the record) or B200_E2E_BULK=1 moves whole records instead. */
int b200_force_step_aos(b200_ctx *ctx, void *particles, int64_t n,
                        const b200_particle_layout *layout,
                        const b200_gravshort_params *par);
/* Bytes per particle that call moves host->device / device->host for `layout` (NULL = the default layout). */
void b200_force_step_aos_bytes(const b200_particle_layout *layout, int64_t *h2d, int64_t *d2h);
/* The same three calls on particles already resident in HBM
 * (b200_set_particles_*): outputs are DEVICE pointers [n][3] / [n], any may be
 * NULL.  gravpm_force runs on a second stream concurrently with
 * force_tree_full + grav_short_tree (the walk reads no PM result: its opening
 * criterion uses the previous step's accelerations, gravshort.h:69-86). */
int b200_force_step_dev(b200_ctx *ctx, const b200_gravshort_params *par,
                        double *gravpm_out, double *pm_potential_out,
                        double *accel_out, double *potential_out);

/* ---- SPH density and hydro force -----------------------------------------------
 * b200_density     replaces density()     (libgadget/density.h:52, density.c:234-355)
 * b200_hydro_force replaces hydro_force() (libgadget/hydra.h:10,  hydra.c:153-245)
 * Call order as run.c:466-489: b200_tree_build(mask = 1: gas) -> b200_sph_set_gas
 * [-> b200_sph_set_timebins / _set_active / _set_state] -> b200_density ->
 * b200_hydro_force.  By default the step is synchronised (every gas particle on one
 * time bin, all of them targets) and the kick/drift factors are the scalars below.  Parameters mirror struct density_params (density.h:10-28),
 * struct hydro_params (hydra.c:26-35). */
typedef struct b200_sph_params {
    int32_t KernelType;               /* DensityKernelType: 1 cubic, 2 quintic, 4 quartic (densitykernel.h:20-24) */
    int32_t DensityIndependentSphOn;
    double DensityResolutionEta, MaxNumNgbDeviation;
    double MinGasHsml;                /* absolute: MinGasHsmlFractional * softening (density.c:265) */
    double ArtBulkViscConst, DensityContrastLimit;
    double gravkick, hydrokick, pmkick;   /* gravkicks[bin], hydrokicks[bin], FgravkickB */
    double dloga_pred;                    /* dloga in SPH_EntVarPred (density.c:74) */
    double drift;                         /* drifts[bin] (hydra.c:185) */
    double dloga_bin;                     /* get_dloga_for_bin (hydra.c:271,463) */
    double atime, hubble;                 /* scale factor and hubble_function(CP, atime) (hydra.c:219-223) */
} b200_sph_params;

/* ---- mixed time bins and active sets (optional; call after b200_sph_set_gas) -------
 * Without these calls every gas particle is a target and sits on one time bin whose
 * factors are the scalar fields of b200_sph_params.
 *
 * b200_sph_set_timebins: P[].TimeBinGravity / P[].TimeBinHydro per particle (NULL = bin 0)
 *   and, per bin, the factors the reference derives from DriftKickTimes:
 *   gravkick/hydrokick = kick_factor_data.gravkicks/hydrokicks (init_kick_factor_data,
 *   density.c:114-132), dloga_pred = dloga_from_dti(Ti_Current - Ti_kick[bin]) of
 *   SPH_EntVarPred (density.c:69-85), drift = drifts[bin] of hydro_force (hydra.c:178-186,
 *   0 for active bins), dloga_bin = get_dloga_for_bin(bin) (hydra.c:271,463).
 * b200_sph_set_active: ActiveParticles.ActiveParticle (timestep.h:29-38): only these
 *   particles are density / hydro targets; every gas particle of the tree stays a source.
 * b200_sph_set_state: SphP[].{Density, EgyWtDensity, DhsmlEgyDensityFactor, DivVel,
 *   CurlVel} of the particles that are NOT targets (the reference reads these stale values
 *   of inactive neighbours in hydro_ngbiter, hydra.c:395-462).  NULL arrays are left as is. */
#define B200_TIMEBINS 46        /* libgadget/timebinmgr.h:13 */
typedef struct b200_sph_bins {
    double gravkick[B200_TIMEBINS + 1], hydrokick[B200_TIMEBINS + 1], dloga_pred[B200_TIMEBINS + 1],
           drift[B200_TIMEBINS + 1], dloga_bin[B200_TIMEBINS + 1];
} b200_sph_bins;
int b200_sph_set_timebins(b200_ctx *ctx, const uint8_t *timebin_gravity, const uint8_t *timebin_hydro, const b200_sph_bins *bins);
int b200_sph_set_active(b200_ctx *ctx, const int32_t *active, int64_t nactive);
/* Multi-GPU: overwrite P[].Hsml of the particles [first, first+count) -- imported ghosts whose
 * owner rank has converged them -- and recompute the tree's hmax (update_tree_hmax_father,
 * forcetree.c:1287-1315) before b200_hydro_force. */
int b200_sph_set_hsml_range(b200_ctx *ctx, const double *hsml, int64_t first, int64_t count);
int b200_sph_set_state(b200_ctx *ctx, const double *density, const double *egywtdensity, const double *dhsmlfac,
                       const double *divvel, const double *curlvel);

/* The array arguments of the b200_sph_* setters and the outputs of b200_density / b200_hydro_force may live in host OR
 * device memory (the copies are issued with cudaMemcpyDefault): a multi-GPU driver keeps the gas state in HBM and hands
 * over device pointers (mp-gadget_b200/sharded.py, ShardedSPH).
 *
 * Per-particle gas state, arrays indexed by particle index (entries of
 * non-gas particles are ignored): Vel[n][3], Hsml[n] (required), Entropy[n],
 * DtEntropy[n], FullTreeGravAccel[n][3], GravPM[n][3], HydroAccel[n][3]; NULL = 0
 * (Entropy NULL = 1).  P[].Vel/Hsml, SphP[].Entropy/DtEntropy/HydroAccel of the
 * reference (partmanager.h:40-120, slotsmanager.h:93-129). */
int b200_sph_set_gas(b200_ctx *ctx, const double *vel, const double *hsml, const double *entropy,
                     const double *dtentropy, const double *fulltreeacc, const double *gravpm,
                     const double *hydroaccel);
/* Outputs (host, [n], any may be NULL): Hsml, Density, EgyWtDensity,
 * DhsmlEgyDensityFactor, DivVel, CurlVel, DtHsml, NumNgb (kernel-weighted),
 * ninteract = particles with r^2 <= h^2 in the last pass (treewalk.c:1233-1240),
 * niter = passes used. */
int b200_density(b200_ctx *ctx, const b200_sph_params *par, int update_hsml, int DoEgyDensity,
                 double *hsml, double *density, double *egywtdensity, double *dhsmlfac,
                 double *divvel, double *curlvel, double *dthsml, double *numngb,
                 int32_t *ninteract, int32_t *niter);
/* GradRho[n][3] of the last b200_density pass (density.c:512-515; the reference
 * returns its magnitude per gas slot in GradRho_mag, density.c:309-317). */
int b200_density_gradrho(b200_ctx *ctx, double *gradrho);
/* HydroAccel[n][3], DtEntropy[n], MaxSignalVel[n], ninteract = candidates from
 * opened leaves (treewalk.c:1056-1143). */
int b200_hydro_force(b200_ctx *ctx, const b200_sph_params *par, double *hydroaccel, double *dtentropy,
                     double *maxsignalvel, int32_t *ninteract);

/* ---- multi-GPU building blocks ------------------------------------------------
 *
 * Top-tree moments: the analogue of force_exchange_pseudodata (MPI_Allgatherv of
 * struct topleaf_momentsdata, libgadget/forcetree.c:1145-1208) and
 * force_treeupdate_pseudos (forcetree.c:1214-1284).  cells[8^level][4] =
 * {cofm.x, cofm.y, cofm.z, mass} of the level-`level` cells of the forced top
 * tree in Morton order (x bit 0, y bit 1, z bit 2 of each octal digit, root digit
 * first); device pointers.  _set overwrites them and re-sums all higher levels. */
int b200_tree_top_get_dev(b200_ctx *ctx, int level, double *cells_out);
int b200_tree_top_set_dev(b200_ctx *ctx, int level, const double *cells_in);

/* ---- the sharded TreePM force step (one process + one context per GPU, NCCL issued from C) ----------------
 * Replaces, for one rank of an N-rank run, what the reference does collectively inside gravpm_force and
 * grav_short_tree: the domain-boundary exchange of the short-range walk (treewalk.c:325-371,399-793; here a ghost
 * import of the adjacent top-cell layer), force_exchange_pseudodata / force_treeupdate_pseudos
 * (forcetree.c:1156-1284; an all-reduce of the level-d cell moments), the pencil exchange of petapm_force
 * (petapm.c:584-885; halo planes to the neighbours) and PFFT's transposes (petapm.c:305,344; an all-to-all between
 * the 2-D and the 1-D transforms of an x-slab mesh).  Domain: rank r owns the x-layers [r*2^d/W, (r+1)*2^d/W) of the
 * cells of the uniform forced top tree of depth d (d such that a cell is wider than Rcut: checked).
 *
 * b200_comm_unique_id: 128 bytes from ncclGetUniqueId, made on rank 0 and broadcast by the host (MPI_Bcast in an
 *   MP-Gadget host, torch.distributed in the harness); two ids -- the tree and the PM chains run on separate streams
 *   and need separate communicators.  NCCL is taken from the process at run time (dlopen of libnccl.so.2).
 * b200_comm_init: ncclCommInitRank x 2 on the context's device.  world == 1 needs no ids (self-neighbour copies).
 * b200_sharded_init: slab mesh, FFT plans, exchange buffers.  halo >= 4 mesh planes; rcut_cells = TreeRcut.
 * b200_sharded_force_step: DEVICE pointers.  pos_own[n][3], mass_own[n], oldacc3_own[n][3] (FullTreeGravAccel +
 *   GravPM of the last step, NULL on the first) of the particles inside this rank's layers; writes GravPM[n][3],
 *   FullTreeGravAccel[n][3], Potential[n] of those particles.  Collective: every rank of the communicator calls it. */
typedef struct b200_sharded_info {
    int64_t n_own, n_from_left, n_from_right, n_to_left, n_to_right;
    /* device time of the phases, ms (CUDA events): ghost import; the PM chain on its stream (runs beside the tree
     * build and walk); the top-moment all-reduce */
    double ms_ghost, ms_pm_total, ms_pm_deposit, ms_pm_halo_add, ms_pm_fft2d, ms_pm_pack, ms_pm_a2a_forward,
           ms_pm_fft1d_transfer, ms_pm_a2a_backward, ms_pm_unpack, ms_pm_ifft2d, ms_pm_halo_fill, ms_pm_readout,
           ms_top_allreduce;
} b200_sharded_info;
int b200_comm_unique_id(void *id_out_128_bytes);
int b200_comm_init(b200_ctx *ctx, int rank, int world, const void *id_tree, const void *id_pm);
int b200_sharded_init(b200_ctx *ctx, double BoxSize, double Asmth, int Nmesh, double G, int topdepth, int halo, double rcut_cells);
int b200_sharded_force_step(b200_ctx *ctx, const double *pos_own, const float *mass_own, const double *oldacc3_own, int64_t n_own,
                            const b200_gravshort_params *par, double *gravpm_out, double *accel_out, double *potential_out,
                            b200_sharded_info *info);

/* ---- Step loop around the force computation (device-resident particle state) -------------------
 * Replaces, for collisionless and gas particles, drift_all_particles (libgadget/drift.c:84-102),
 * build_active_particles / build_active_sublist (timestep.c:1334-1478), apply_half_kick /
 * apply_hydro_half_kick / apply_PM_half_kick (timestep.c:874-994) and the hierarchical gravity drivers
 * hierarchical_gravity_accelerations / hierarchical_gravity_and_timesteps (timestep.c:296-598), so
 * that P[].Pos / Vel / TimeBinGravity live in HBM between force computations (run.c:355-800).  The
 * host keeps DriftKickTimes, the cosmology and its integrals (get_exact_gravkick_factor & co.) and
 * hands factors in; black-hole particles (drag terms, repositioning) are not supported.
 * Status of this block in round 1: compiled for sm_100a and checked on the CPU side against the
 * oracle restatement; not yet run on hardware (tests/test_step_gpu.py). */
typedef struct b200_step_state {      /* host arrays in particle-index order, any may be NULL (= 0) */
    const double *vel;                /* [n][3] P[].Vel */
    const double *fullacc;            /* [n][3] P[].FullTreeGravAccel */
    const double *gravpm;             /* [n][3] P[].GravPM */
    const uint8_t *bin_grav;          /* [n] P[].TimeBinGravity */
    const uint8_t *bin_hydro;         /* [n] P[].TimeBinHydro */
    const uint8_t *flags;             /* [n] bit 0 IsGarbage, bit 1 Swallowed; NULL keeps what b200_set_particles_* set */
    const double *hsml, *dthsml;      /* [n] P[].Hsml, P[].DtHsml (gas) */
    const double *hydroacc;           /* [n][3] SphP[].HydroAccel */
    const double *entropy, *dtentropy;/* [n] SphP[].Entropy, SphP[].DtEntropy */
    double BoxSize;                   /* PartManager->BoxSize; 0 = the box b200_pm_init / b200_tree_build was given */
} b200_step_state;
typedef struct b200_step_state_out {  /* host outputs, any may be NULL */
    double *pos, *vel, *fullacc, *hsml, *entropy;
    uint8_t *bin_grav, *bin_hydro;
    double *hydroacc, *dtentropy, *maxsignalvel;   /* SphP[].HydroAccel [n][3], DtEntropy [n], MaxSignalVel [n] */
} b200_step_state_out;
typedef struct b200_step_times {      /* DriftKickTimes, timestep.h:10-26 */
    int32_t mintimebin, maxtimebin, mingravtimebin, pad_;
    int64_t Ti_kick[B200_TIMEBINS + 1], Ti_lastactivedrift[B200_TIMEBINS + 1];
    int64_t Ti_Current, PM_length, PM_start, PM_kick;
} b200_step_times;
typedef struct b200_step_params {
    double ErrTolIntAccuracy, MaxSizeTimestep, MinSizeTimestep, MaxRMSDisplacementFac, CourantFac;   /* TimestepParams, timestep.c:21-47 */
    double softening;                 /* FORCE_SOFTENING(), gravshort-tree.c:37-41 */
    double omega_type[6];             /* density parameter the mean spacing of each particle type is taken from
                                       * (OmegaBaryon / OmegaCDM / get_omega_nu, timestep.c:1251-1263) */
    double RhoCrit;                   /* CP->RhoCrit */
    int32_t FastParticleType, pad_;
    const double *sync_loga;          /* SyncPoints[].loga, timebinmgr.c:18 */
    int64_t nsync;
    double (*gravkick_factor)(void *user, int64_t ti0, int64_t ti1);   /* get_exact_gravkick_factor, timefac.c:65-68 */
    void *user;
} b200_step_params;

/* Upload the state the step loop owns (after b200_set_particles_*). */
int b200_step_set_state(b200_ctx *ctx, const b200_step_state *state);
int b200_step_get_state(b200_ctx *ctx, b200_step_state_out *out);
/* FullTreeGravAccel / GravPM of the state <- the device results of the last full-tree
 * b200_grav_short_tree / b200_pm_force (no host round trip). */
int b200_step_adopt_forces(b200_ctx *ctx, int tree, int pm);
/* drift_all_particles with ddrift = get_exact_drift_factor(ti0, ti1); shift[3] may be NULL.
 * *nbad = particles the reference would endrun on (Hsml <= 0, non-finite position); returns nonzero then. */
int b200_step_drift(b200_ctx *ctx, double ddrift, const double *shift, int64_t *nbad);
/* build_active_particles: the list stays on the device.  counts = {NumActiveParticle, NumActiveGravity,
 * NumActiveHydro}; bincounts[6][B200_TIMEBINS + 1] = TimeBinCountType (may be NULL); nhydro_slots =
 * SlotsManager->info[0].size + info[5].size (what the PM branch reports as NumActiveHydro). */
int b200_step_build_active(b200_ctx *ctx, int64_t Ti_Current, int is_pm, int64_t nhydro_slots, int64_t *counts, int64_t *bincounts);
/* The host's own ActiveParticles list (ascending particle indices) as the current list; NULL = every particle. */
int b200_step_set_active(b200_ctx *ctx, const int32_t *list, int64_t nlist);
/* build_active_sublist of the current active list */
int b200_step_active_sublist(b200_ctx *ctx, int maxtimebin, int64_t Ti_Current, int64_t *nsub);
/* which = 0: the active list (*nout = -1 - n when it is implicit), 1: the last sub-list */
int b200_step_get_active(b200_ctx *ctx, int which, int32_t *out, int64_t *nout);
/* apply_half_kick (hydro_only = 0) / apply_hydro_half_kick (1) on the current active list;
 * gravkick / hydrokick / dt_entr [B200_TIMEBINS + 1] by bin as timestep.c:879-891,905-907 compute them */
int b200_step_half_kick(b200_ctx *ctx, const double *gravkick, const double *hydrokick, const double *dt_entr,
                        int64_t Ti_Current, double atime, double MaxGasVel, int hydro_only);
/* apply_PM_half_kick: Vel += GravPM * Fgravkick (the caller advances PM_kick) */
int b200_step_pm_kick(b200_ctx *ctx, double Fgravkick);
/* Device-resident gas sub-step (run.c:466-495 between the active list and the kicks): b200_step_sph_prepare hands
 * the current active list and the per-bin factor tables to the SPH module (the time bins are the ones already on the
 * device), then b200_density / b200_hydro_force run as usual (output pointers may be NULL), then
 * b200_step_adopt_hydro takes SphP[].HydroAccel / DtEntropy / MaxSignalVel of the listed gas from the device
 * results, so that b200_step_half_kick and b200_step_hydro_timesteps (maxsignalvel = NULL) can follow. */
int b200_step_sph_prepare(b200_ctx *ctx, const b200_sph_bins *tables);
int b200_step_adopt_hydro(b200_ctx *ctx);
/* find_hydro_timesteps (timestep.c:617-738) for the gas on the current active list: new P[].TimeBinHydro on
 * the device, times->mintimebin updated.  maxsignalvel[n] = SphP[].MaxSignalVel by particle index (host). */
int b200_step_hydro_timesteps(b200_ctx *ctx, const b200_step_params *sp, b200_step_times *times, const double *maxsignalvel,
                              double atime, double hubble, int64_t *nbad);
/* force_tree_full + grav_short_tree for the current active list (run.c:541-548, SplitGravityTimestepsOn = 0): the
 * FullTreeGravAccel of the listed particles in the step state is replaced; gp->TreeUseBH > 1 -> 0 afterwards */
int b200_step_grav_short_tree(b200_ctx *ctx, b200_gravshort_params *gp);
/* find_timesteps (timestep.c:739-853), the SplitGravityTimestepsOn = 0 loop: one bin for TimeBinGravity and
 * TimeBinHydro from the gravity and (gas) hydro criteria, PM step length on PM steps, times->mintimebin / maxtimebin. */
int b200_step_find_timesteps(b200_ctx *ctx, const b200_step_params *sp, b200_step_times *times, const double *maxsignalvel,
                             int is_pm, double atime, double hubble, int64_t *nbad);
/* hierarchical_gravity_accelerations on the current active list (ngrav = NumActiveGravity);
 * gp->TreeUseBH > 1 is reset to 0 after the first walk like TreeParams.TreeUseBH */
int b200_step_hier_accelerations(b200_ctx *ctx, const b200_step_params *sp, b200_gravshort_params *gp,
                                 b200_step_times *times, int64_t ngrav);
/* StoredGravAccel of the last b200_step_hier_accelerations -> host [n][3] (0 for particles it did not walk) */
int b200_step_get_store(b200_ctx *ctx, double *out);
/* ... and the way back: the host's StoredGravAccel.GravAccel (timestep.h:92-96) for b200_step_hier_timesteps
 * after a new b200_step_set_state; NULL = none, the drivers then use FullTreeGravAccel as the reference does */
int b200_step_set_store(b200_ctx *ctx, const double *in);
/* hierarchical_gravity_and_timesteps; hubble = hubble_function(CP, atime);
 * info = {largest active bin, PM_length, bad-step count} */
int b200_step_hier_timesteps(b200_ctx *ctx, const b200_step_params *sp, b200_gravshort_params *gp,
                             b200_step_times *times, int64_t ngrav, int is_pm, double atime, double hubble, int64_t *info);

/* ---- Domain keys (first brick of the device-side domain decomposition) --------------------------------
 * Status as the step loop above: compiled, emulation-checked against the reference's known-answer keys, hardware run pending. */
/* PEANO(P[i].Pos, BoxSize) (libgadget/utils/peano.h:15-21, peano.c:108-129) of every particle set with
 * b200_set_particles_*; keys stay on the device, keys_out (host, [n]) may be NULL. */
int b200_domain_peano_keys(b200_ctx *ctx, double BoxSize, uint64_t *keys_out);
/* DomainDecomp::TopNodes (libgadget/domain.h:20-33) as arrays: Daughter, StartKey, Shift, Leaf */
int b200_domain_set_topnodes(b200_ctx *ctx, int32_t ntop, const int32_t *daughter, const uint64_t *startkey,
                             const int32_t *shift, const int32_t *leaf);
/* P[i].TopLeaf = domain_get_topleaf(key_i) (domain.h:71-78) from the keys of the last b200_domain_peano_keys */
int b200_domain_topleaf(b200_ctx *ctx, int32_t *topleaf_out);
/* The top tree (domain_determine_global_toptree, domain.c:1281-1340) stage by stage.  The device supplies the keys of
 * the subsample (every `subsample`-th particle, domain.c:1066-1074: 8 N / subsample bytes cross PCIe); the tree itself --
 * hundreds to thousands of nodes -- is sequential integer work on the host, as in the reference:
 *   local (domain_check_for_local_refine_subsample, sorts the keys in place) -> truncate with the limits from the
 *   summed root counts -> pairwise merge of the ranks' trees (domain_nonrecursively_combine_topTree) ->
 *   global_refine -> leaves (domain_create_topleaves; TopNodes[] = StartKey / Shift / Daughter of the nodes + leaf_out).
 * Status: 0 ok, 1 out of nodes (retry with a larger maxnodes, domain.c:186-193), 2 the samples are too clustered,
 * 3 bad arguments. */
typedef struct b200_topnode {         /* struct local_topnode_data, domain.c:60-70 */
    uint64_t StartKey; int32_t Shift, Daughter, Parent, pad_; int64_t Count, Cost;
} b200_topnode;
int b200_domain_sample_keys(b200_ctx *ctx, double BoxSize, int32_t subsample, uint64_t *keys_out, int64_t *nsample);
int b200_domain_toptree_local(uint64_t *sample_keys, int64_t nsample, b200_topnode *tree, int32_t *size, int32_t maxnodes);
int b200_domain_toptree_truncate(b200_topnode *tree, int32_t *size, int64_t countlimit, int64_t costlimit);
int b200_domain_toptree_merge(b200_topnode *treeA, int32_t *sizeA, const b200_topnode *treeB, int32_t maxnodes);
int b200_domain_toptree_global_refine(b200_topnode *tree, int32_t *size, int32_t maxnodes, int64_t countlimit, int64_t costlimit);
int b200_domain_toptree_leaves(const b200_topnode *tree, int32_t size, int32_t *leaf_out, int32_t *nleaf);
/* TopLeafCount of domain_compute_costs (domain.c:1396-1451): particles per top leaf from the last b200_domain_topleaf,
 * garbage skipped */
int b200_domain_leaf_counts(b200_ctx *ctx, int32_t nleaf, int64_t *counts_out);
/* The exchange plan of this rank (domain_build_exchange_list + domain_build_plan, exchange.c:408-444,505-530, with
 * domain_layoutfunc, domain.c:794-803) from the last b200_domain_topleaf and the leaf -> task table: the particles that
 * leave (ascending index; the list stays on the device, list_out may be NULL), the garbage count, and
 * togo[ntask][7] = {base, slots[0..5]}: how many go to every task in total and per particle type.  Moving the records
 * (exchange.c:211-405) is not part of this library yet. */
int b200_domain_exchange_plan(b200_ctx *ctx, const int32_t *task_of_leaf, int32_t nleaf, int32_t ntask, int32_t thistask,
                              int64_t *nexchange, int64_t *ngarbage, int64_t *togo, int32_t *list_out);
/* domain_assign_topleaves_balanced (domain.c:610-755): task of every top leaf (leaves in key order) from the per-leaf
 * cost; host arithmetic, no context.  Non-zero where the reference would endrun. */
int b200_domain_assign_balanced(int32_t ntask, int32_t nleaf, const int64_t *cost, int32_t nseg_per_task, int32_t *task_out);

/* ---- friends-of-friends, primary linking -------------------------------------------------------
 * Replaces fof_label_primary + fof_primary_ngbiter (libgadget/fof.c:366-470,540-579), the repeated treewalk_run of the
 * FOF group finder, for the particles set with b200_set_particles_*: particles whose type is in primary_mask
 * (FOFPrimaryLinkTypes, bit t = type t) and that lie within linking_length of each other (periodic, NEAREST) form one
 * group; minid_out[i] = the smallest ids[] of i's group (HaloLabel[i].MinID at the fixed point), the particle's own ID
 * for other types, garbage and swallowed particles.  ids and minid_out are host arrays of n entries; *ngroups_out
 * (may be NULL) = number of groups holding a primary particle.  BoxSize / linking_length up to 1024 grid cells a side
 * are used; a linking length above a third of the box is examined pair by pair (at most 65536 primary particles). */
int b200_fof_primary(b200_ctx *ctx, const int64_t *ids, int primary_mask, double BoxSize, double linking_length,
                     int64_t *minid_out, int64_t *ngroups_out);

/* Device-side timing of the phases of the last call, milliseconds (CUDA
 * events on the engine's stream).  Names follow the reference's walltime
 * categories (libgadget/walltime.c, gravshort-tree.c:134-144, petapm.c:280-355). */
typedef struct b200_timings {
    double pm_deposit, pm_fft_forward, pm_transfer, pm_fft_inverse, pm_gradient, pm_readout, pm_total;
    double tree_keys, tree_sort, tree_nodes, tree_moments, tree_total;
    double walk, walk_post;        /* k_grav_walk (traversal, node terms), k_grav_pairs (pair sums + postprocess) */
    double h2d, d2h;
    double sph_density, sph_hydro;
    /* statistics of the last tree walk: leaf pieces (<= 8 source particles each) queued for the
     * pair kernel = ninteractions / 8 rounded up per leaf, and the bytes of list storage used */
    double walk_pieces, walk_list_bytes;
} b200_timings;
int b200_get_timings(const b200_ctx *ctx, b200_timings *t);

/* Stream the engine launches on (cudaStream_t as void*), so harness code can
 * record events on it. */
void *b200_stream(const b200_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* B200FORCE_H */
