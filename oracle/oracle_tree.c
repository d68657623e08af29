/* oracle_tree.c -- CPU restatement of the MP-Gadget octree and short-range
 * tree gravity.  TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * The reference builds its tree by sequential insertion (forcetree.c:481-520)
 * with a split whenever a 9th particle lands in a leaf (forcetree.c:370-477).
 * The resulting topology does not depend on insertion order: a cell is internal
 * iff it holds more than NMAXCHILD=8 particles, a child exists iff it is
 * non-empty (empty children are pruned at forcetree.c:1028-1049) or is part of
 * the forced top-tree.  Inside a leaf the particles sit in insertion order,
 * which for the single-threaded reference is ascending particle index.  This
 * file therefore builds the same tree top-down with stable partitions, which
 * is a different program with the same result.
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <stdio.h>
#include "oracle.h"
#include "../mp-gadget_b200/data/shortrange_table.h"

#define LEAFCAP 8      /* NMAXCHILD, forcetree.h:13 */

typedef struct {
    oracle_node *nodes;
    int64_t used, cap;
    const double *pos;
    const float *mass;
    const uint8_t *type;
    const double *hsml;
    int32_t *scratch;
    int toplevel_depth;
    /* an arbitrary domain top tree (DomainDecomp::TopNodes, domain.h:20-33) instead of a uniform depth: Daughter[] */
    const int32_t *top_daughter;
    int32_t ntop;
    int failed;
    double BoxSize;
} builder;

/* Octant of a particle below the cell at `level`.  Inside the forced top tree
 * the reference does not compare positions: it drops the particle below
 * P[i].TopLeaf (forcetree.c:819-823), which the domain code derived from the
 * integer Peano-Hilbert lattice, PEANO() utils/peano.h:15-21.  The lattice is
 * the tree's own (same 1.001*Box root, offset Box/2000), so the two agree
 * except for positions within rounding of a cell boundary; to follow the
 * reference there too, top-tree levels use the lattice bits.  Below the top
 * leaves it is get_subnode, forcetree.c:278-284 (strict >). */
static inline int octant_of(const builder *b, const double *x, const double c[3], int level, int in_top)
{
    if(in_top) {
        const double DomainFac = 1.0 / (b->BoxSize * 1.001) * (((uint64_t) 1) << 21);
        int s = 0;
        for(int j = 0; j < 3; j++) {
            const int ix = (int) ((x[j] + b->BoxSize / 2000) * DomainFac);
            s |= ((ix >> (21 - (level + 1))) & 1) << j;
        }
        return s;
    }
    return (x[0] > c[0]) + ((x[1] > c[1]) << 1) + ((x[2] > c[2]) << 2);
}

static int64_t new_node(builder *b)
{
    if(b->used == b->cap) {
        b->cap = b->cap * 2 + 1024;
        b->nodes = (oracle_node *) realloc(b->nodes, b->cap * sizeof(oracle_node));
    }
    memset(&b->nodes[b->used], 0, sizeof(oracle_node));
    return b->used++;
}

/* Leaf moments: add_particle_moment_to_node forcetree.c:947-966 in insertion
 * order, then force_update_particle_node forcetree.c:985-1004. */
static void leaf_moments(builder *b, oracle_node *nd)
{
    double m = 0, s[3] = {0, 0, 0}, hmax = 0;
    for(int k = 0; k < nd->nocc; k++) {
        const int32_t p = nd->part[k];
        const double pm = (double) b->mass[p];
        m += pm;
        for(int j = 0; j < 3; j++)
            s[j] += pm * b->pos[3 * p + j];
        if(b->hsml && b->type && (b->type[p] == 0 || b->type[p] == 5)) {
            for(int j = 0; j < 3; j++) {
                double h = fabs(b->pos[3 * p + j] - nd->center[j]) + b->hsml[p] - nd->len / 2.;
                if(h > hmax) hmax = h;
            }
        }
    }
    nd->mass = m;
    nd->hmax = hmax;
    for(int j = 0; j < 3; j++)
        nd->cofm[j] = m > 0 ? s[j] / m : nd->center[j];
}

/* Build the subtree holding idx[0..cnt) rooted at a cell (center c, side len).
 * Child geometry: init_internal_node forcetree.c:302-320; octant choice:
 * get_subnode forcetree.c:278-284 (strict >). Returns the node's DFS index. */
uint64_t oracle_peano_key(int x, int y, int z, int bits);      /* oracle_domain.c */

/* topnode: index of the cell's node in the domain top tree, -1 below the top leaves (and always in uniform-depth
 * mode); cx[3]: the cell's integer coordinates at its level (force_create_node_for_topnode forcetree.c:869-934). */
static int64_t build_cell(builder *b, int32_t *idx, int64_t cnt, const double c[3], double len,
                          int level, int64_t father, int32_t topnode, const int cx[3])
{
    const int64_t me = new_node(b);
    {
        oracle_node *nd = &b->nodes[me];
        nd->father = (int32_t) father;
        nd->sibling = -1;
        nd->firstchild = -1;
        nd->len = len;
        nd->level = level;
        nd->toplevel = b->top_daughter ? topnode >= 0 : level <= b->toplevel_depth;
        for(int j = 0; j < 3; j++) nd->center[j] = c[j];
        for(int k = 0; k < 8; k++) nd->part[k] = -1;
    }
    const int forced_internal = b->top_daughter ? (topnode >= 0 && b->top_daughter[topnode] >= 0) : level < b->toplevel_depth;
    if(cnt <= LEAFCAP && !forced_internal) {
        oracle_node *nd = &b->nodes[me];
        nd->nocc = (int32_t) cnt;
        for(int64_t k = 0; k < cnt; k++) nd->part[k] = idx[k];
        leaf_moments(b, nd);
        return me;
    }
    if(level > 200) { b->failed = 1; b->nodes[me].nocc = 0; return me; }

    /* stable 8-way partition */
    int64_t count[8] = {0}, start[9];
    for(int64_t k = 0; k < cnt; k++) {
        const double *x = &b->pos[3 * (int64_t) idx[k]];
        const int s = octant_of(b, x, c, level, forced_internal);
        count[s]++;
    }
    start[0] = 0;
    for(int s = 0; s < 8; s++) start[s + 1] = start[s] + count[s];
    {
        int64_t fill[8];
        for(int s = 0; s < 8; s++) fill[s] = start[s];
        for(int64_t k = 0; k < cnt; k++) {
            const double *x = &b->pos[3 * (int64_t) idx[k]];
            const int s = octant_of(b, x, c, level, forced_internal);
            b->scratch[fill[s]++] = idx[k];
        }
        memcpy(idx, b->scratch, cnt * sizeof(int32_t));
    }
    b->nodes[me].nocc = -1;
    const double lenhalf = 0.25 * len;
    int64_t kids[8]; int nk = 0;
    for(int s = 0; s < 8; s++) {
        /* pruning of empty non-top-level children: forcetree.c:1028-1049 */
        if(count[s] == 0 && !forced_internal) continue;
        double cc[3];
        for(int j = 0; j < 3; j++) {
            const int sign = (s & (1 << j)) ? 1 : -1;
            cc[j] = c[j] + sign * lenhalf;
        }
        int32_t ctop = -1;
        const int ccx[3] = {2 * cx[0] + (s & 1), 2 * cx[1] + ((s >> 1) & 1), 2 * cx[2] + ((s >> 2) & 1)};
        if(b->top_daughter && forced_internal) {
            /* forcetree.c:886,904: the daughter that covers this octant is found through the curve */
            const int sub = (int) (7 & oracle_peano_key(ccx[0], ccx[1], ccx[2], level + 1));
            ctop = b->top_daughter[topnode] + sub;
            if(ctop >= b->ntop) { b->failed = 1; ctop = -1; }
        }
        kids[nk++] = build_cell(b, idx + start[s], count[s], cc, 0.5 * len, level + 1, me, ctop, ccx);
    }
    /* moments of an internal node: forcetree.c:1081-1101 (children in octant order) */
    oracle_node *nd = &b->nodes[me];
    nd->firstchild = nk ? (int32_t) kids[0] : -1;
    double m = 0, s3[3] = {0, 0, 0}, hmax = 0;
    for(int k = 0; k < nk; k++) {
        const oracle_node *ch = &b->nodes[kids[k]];
        m += ch->mass;
        for(int j = 0; j < 3; j++) s3[j] += ch->mass * ch->cofm[j];
        if(ch->hmax > hmax) hmax = ch->hmax;
    }
    nd->mass = m;
    nd->hmax = hmax;
    /* zero-mass internal nodes only occur in the forced top tree:
     * force_treeupdate_pseudos forcetree.c:1272-1283 */
    for(int j = 0; j < 3; j++) nd->cofm[j] = m > 0 ? s3[j] / m : nd->center[j];
    /* sibling threading: force_get_sibling forcetree.c:969-982; the last child
     * inherits the parent's sibling, fixed up below by thread_siblings */
    for(int k = 0; k + 1 < nk; k++) b->nodes[kids[k]].sibling = (int32_t) kids[k + 1];
    return me;
}

/* The last child of every node points at the parent's sibling
 * (forcetree.c:1058-1060).  DFS preorder: parents precede children. */
static void thread_siblings(oracle_node *nodes, int64_t n)
{
    for(int64_t i = 1; i < n; i++)
        if(nodes[i].sibling == -1)
            nodes[i].sibling = nodes[nodes[i].father].sibling;
}

int oracle_tree_build(oracle_tree *t, const double *pos, const float *mass,
                      const uint8_t *type, const double *hsml, int64_t n, double BoxSize,
                      int mask, const int32_t *active, int64_t nactive, int toplevel_depth)
{
    return oracle_tree_build_top(t, pos, mass, type, hsml, n, BoxSize, mask, active, nactive, toplevel_depth, NULL, 0);
}

/* The same with the forced top tree given as the Daughter[] column of DomainDecomp::TopNodes (force_tree_create_topnodes,
 * forcetree.c:654-687): a cell is forced to be internal, with all eight children, iff its top node has daughters. */
int oracle_tree_build_top(oracle_tree *t, const double *pos, const float *mass,
                          const uint8_t *type, const double *hsml, int64_t n, double BoxSize,
                          int mask, const int32_t *active, int64_t nactive, int toplevel_depth,
                          const int32_t *top_daughter, int32_t ntop)
{
    builder b;
    memset(&b, 0, sizeof(b));
    b.pos = pos; b.mass = mass; b.type = type; b.hsml = hsml;
    b.toplevel_depth = toplevel_depth;
    b.BoxSize = BoxSize;
    const int64_t nin = active ? nactive : n;
    int32_t *idx = (int32_t *) malloc((nin + 1) * sizeof(int32_t));
    b.scratch = (int32_t *) malloc((nin + 1) * sizeof(int32_t));
    int64_t cnt = 0;
    /* particle filter: forcetree.c:799-807 (type mask) */
    for(int64_t j = 0; j < nin; j++) {
        const int32_t i = active ? active[j] : (int32_t) j;
        const int ty = type ? type[i] : 1;
        if(!((1 << ty) & mask)) continue;
        idx[cnt++] = i;
    }
    /* root: force_tree_create_topnodes forcetree.c:662-664 */
    const double c[3] = {BoxSize / 2., BoxSize / 2., BoxSize / 2.};
    const int cx0[3] = {0, 0, 0};
    b.top_daughter = top_daughter; b.ntop = ntop;
    build_cell(&b, idx, cnt, c, BoxSize * 1.001, 0, -1, top_daughter ? 0 : -1, cx0);
    thread_siblings(b.nodes, b.used);
    free(idx); free(b.scratch);
    t->nodes = b.nodes;
    t->numnodes = b.used;
    t->numparticles = cnt;
    t->BoxSize = BoxSize;
    return b.failed;
}

void oracle_tree_free(oracle_tree *t)
{
    free(t->nodes);
    memset(t, 0, sizeof(*t));
}

/* ---------------- short-range gravity ---------------- */

static inline double nearest(double x, double box)   /* NEAREST partmanager.h:99 */
{
    return (x > 0.5 * box) ? (x - box) : ((x < -0.5 * box) ? (x + box) : x);
}

typedef struct { double acc[3]; double pot; } gresult;

/* apply_accn_to_output gravshort-tree.c:158-193 + grav_apply_short_range_window
 * gravity.c:54-66 */
static void add_monopole(gresult *out, const double dx[3], double r2, double mass, double cellsize, double h)
{
    const double r = sqrt(r2);
    double fac = mass / (r2 * r);
    double facpot = -mass / r;
    if(r2 < h * h) {
        double wp;
        const double h3_inv = 1.0 / h / h / h;
        const double u = r / h;
        if(u < 0.5) {
            fac = mass * h3_inv * (10.666666666667 + u * u * (32.0 * u - 38.4));
            wp = -2.8 + u * u * (5.333333333333 + u * u * (6.4 * u - 9.6));
        } else {
            fac = mass * h3_inv * (21.333333333333 - 48.0 * u + 38.4 * u * u
                                   - 10.666666666667 * u * u * u - 0.066666666667 / (u * u * u));
            wp = -3.2 + 0.066666666667 / u + u * u * (10.666666666667 + u * (-16.0 + u * (9.6 - 2.133333333333 * u)));
        }
        facpot = mass / h * wp;
    }
    const double dxtab = B200_SR_DX;
    const double i = r / cellsize / dxtab;
    const size_t t = (size_t) floor(i);
    if(t >= B200_SR_NTAB - 1)
        return;
    fac *= (t + 1 - i) * b200_sr_force[t] + (i - t) * b200_sr_force[t + 1];
    facpot *= (t + 1 - i) * b200_sr_pot[t] + (i - t) * b200_sr_pot[t + 1];
    for(int j = 0; j < 3; j++) out->acc[j] += dx[j] * fac;
    out->pot += facpot;
}

int oracle_grav_short_tree(const oracle_tree *t, const double *pos, const float *mass, int64_t n,
                           const oracle_gravshort_params *par, double G, int Nmesh, double Asmth,
                           const double *oldacc, const int32_t *active, int64_t nactive,
                           int full_particle_tree,
                           double *accel_out, double *pot_out, oracle_walk_counts *counts_out)
{
    const double Box = t->BoxSize;
    const double cellsize = Box / Nmesh;                       /* gravshort-tree.c:101 */
    const double rcut = par->Rcut * Asmth * cellsize;          /* :102 */
    const double rcut2 = rcut * rcut;
    const double h = 2.8 * par->GravitySoftening;              /* FORCE_SOFTENING :37-41 */
    const int usebh = par->TreeUseBH;
    double theta2 = par->BHOpeningAngle * par->BHOpeningAngle; /* :266-270 */
    if(usebh == 0) theta2 = par->MaxBHOpeningAngle * par->MaxBHOpeningAngle;
    const double cbrtrho0 = pow(par->rho0, 1.0 / 3);
    const int64_t nq = active ? nactive : n;
    const oracle_node *N = t->nodes;
    int nomem = 0;

#pragma omp parallel
    {
        int32_t *cand = (int32_t *) malloc(sizeof(int32_t) * (t->numparticles + 8));
        if(!cand) nomem = 1;
#pragma omp for schedule(dynamic, 64)
        for(int64_t q = 0; q < nq; q++) {
            if(!cand) continue;
            const int64_t i = active ? active[q] : q;
            const double *inpos = &pos[3 * i];
            double aold = 0;                                    /* gravshort.h:69-86 */
            if(oldacc) {
                double s = 0;
                for(int j = 0; j < 3; j++) s += oldacc[3 * i + j] * oldacc[3 * i + j];
                aold = sqrt(s) / G;
            }
            aold *= par->ErrTolForceAcc;                        /* gravshort-tree.c:264 */
            gresult out = {{0, 0, 0}, 0};
            oracle_walk_counts cnt = {0, 0, 0, 0};
            int64_t numcand = 0;
            int32_t no = t->numnodes > 0 ? 0 : -1;
            while(no >= 0) {
                const oracle_node *nop = &N[no];
                double dx[3];
                for(int j = 0; j < 3; j++) dx[j] = nearest(nop->cofm[j] - inpos[j], Box);
                const double r2 = dx[0] * dx[0] + dx[1] * dx[1] + dx[2] * dx[2];
                /* shall_we_discard_node :198-215 */
                int discard = 0;
                if(r2 > rcut2) {
                    const double eff = rcut + 0.5 * nop->len;
                    for(int j = 0; j < 3; j++)
                        if(fabs(nearest(nop->center[j] - inpos[j], Box)) > eff) discard = 1;
                }
                if(discard) { cnt.nodes_discarded++; no = nop->sibling; continue; }
                /* shall_we_open_node :220-241 */
                int open = 0;
                if(usebh == 0 && (nop->mass * nop->len * nop->len > r2 * r2 * aold)) open = 1;
                else if(nop->len * nop->len / r2 > theta2) open = 1;
                else {
                    const double inside = 0.6 * nop->len;
                    if(fabs(nearest(nop->center[0] - inpos[0], Box)) < inside &&
                       fabs(nearest(nop->center[1] - inpos[1], Box)) < inside &&
                       fabs(nearest(nop->center[2] - inpos[2], Box)) < inside) open = 1;
                }
                if(!open) {
                    add_monopole(&out, dx, r2, nop->mass, cellsize, h);
                    cnt.nodes_accepted++;
                    no = nop->sibling;
                    continue;
                }
                if(nop->nocc >= 0) {                            /* particle leaf :344-352 */
                    for(int k = 0; k < nop->nocc; k++) cand[numcand++] = nop->part[k];
                    no = nop->sibling;
                } else {
                    cnt.nodes_opened++;
                    no = nop->firstchild;
                }
            }
            for(int64_t k = 0; k < numcand; k++) {              /* :364-374 */
                const int64_t pp = cand[k];
                double dx[3];
                for(int j = 0; j < 3; j++) dx[j] = nearest(pos[3 * pp + j] - inpos[j], Box);
                const double r2 = dx[0] * dx[0] + dx[1] * dx[1] + dx[2] * dx[2];
                add_monopole(&out, dx, r2, (double) mass[pp], cellsize, h);
            }
            cnt.particles = (int32_t) numcand;
            /* grav_short_postprocess gravshort.h:47-67 */
            if(accel_out)
                for(int j = 0; j < 3; j++) accel_out[3 * i + j] = out.acc[j] * G;
            if(pot_out) {
                double p = out.pot;
                if(full_particle_tree) {
                    p += mass[i] / (h / 2.8);
                    p -= 2.8372975 * pow(mass[i], 2.0 / 3) * cbrtrho0;
                    p *= G;
                }
                pot_out[i] = p;
            }
            if(counts_out) counts_out[i] = cnt;
        }
        free(cand);
    }
    return nomem;
}

/* tests/test_gravity.c:38-150: pairwise Newtonian + spline softening over the
 * box and its (2*repeat+1)^3 - 1 images. */
void oracle_direct_sum(const double *pos, const float *mass, int64_t n, double BoxSize, double G,
                       double h, int repeat, double *accel_out)
{
#pragma omp parallel for schedule(dynamic, 16)
    for(int64_t i = 0; i < n; i++) {
        double a[3] = {0, 0, 0};
        for(int xx = -repeat; xx <= repeat; xx++)
        for(int yy = -repeat; yy <= repeat; yy++)
        for(int zz = -repeat; zz <= repeat; zz++) {
            const double off[3] = {BoxSize * xx, BoxSize * yy, BoxSize * zz};
            for(int64_t j = 0; j < n; j++) {
                if(j == i) continue;
                double d[3], r2 = 0;
                for(int k = 0; k < 3; k++) {
                    d[k] = off[k] + pos[3 * i + k] - pos[3 * j + k];
                    r2 += d[k] * d[k];
                }
                const double r = sqrt(r2);
                double fac = 1 / (r2 * r);
                if(r < h) {
                    const double h_inv = 1.0 / h, h3_inv = h_inv * h_inv * h_inv, u = r * h_inv;
                    if(u < 0.5)
                        fac = h3_inv * (10.666666666667 + u * u * (32.0 * u - 38.4));
                    else
                        fac = h3_inv * (21.333333333333 - 48.0 * u + 38.4 * u * u
                                        - 10.666666666667 * u * u * u - 0.066666666667 / (u * u * u));
                }
                for(int k = 0; k < 3; k++) a[k] += -d[k] * fac * G * mass[j];
            }
        }
        for(int k = 0; k < 3; k++) accel_out[3 * i + k] = a[k];
    }
}
