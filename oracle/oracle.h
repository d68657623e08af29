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
 *     from /root/reference (oracle/_ref, tests/test_oracle_vs_ref.py) and
 *     against committed outputs of it (tests/golden/);
 *   - PM: "parity unpinned" at the mesh level (the reference holds no golden
 *     PM value and PFFT is not buildable offline); pinned only through the
 *     reference's own TreePM-vs-direct-sum bounds (tests/test_gravity.c:146-160).
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
/* gravpm.c:458-489. dim 0/1/2 = x/y/z. out may alias nothing (copy then scale). */
void oracle_pm_force_transfer(const double *potk, double *out, int Nmesh, double BoxSize, int dim);
/* petapm.c:955-1006 + gravpm.c:499-510: out[i*ostride] += sum_c w_c mesh[c]. */
void oracle_pm_readout(const double *mesh, const double *pos, int64_t n, double BoxSize, int Nmesh,
                       double *out, int64_t ostride);
/* Direct periodic-image summation of tests/test_gravity.c:38-150 (grav_force,
 * force_direct) -- the reference's own ground truth for TreePM accuracy. */
void oracle_direct_sum(const double *pos, const float *mass, int64_t n, double BoxSize, double G,
                       double softening_h, int repeat, double *accel_out);

#ifdef __cplusplus
}
#endif
#endif
