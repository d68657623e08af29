// tree_walk.cu -- short-range Barnes-Hut tree gravity (replaces grav_short_tree,
// libgadget/gravshort-tree.c:96-154, i.e. treewalk_run + ev_primary
// (treewalk.c:266-310,801-902) with the visitor force_treeev_shortrange
// (gravshort-tree.c:253-379)).
//
// One warp walks the tree for 32 targets that are adjacent on the space-filling
// curve.  The warp follows the reference's depth-first sibling/first-child
// order; every lane takes ITS OWN discard / accept / open decision with the
// reference's criteria (gravshort-tree.c:198-241), so per-particle results and
// interaction counts are those of the per-particle CPU walk.  A subtree is
// entered when any lane opens the node (warp ballot); lanes that accepted or
// discarded the node park until the walk reaches that node's sibling.  Node
// rows are therefore fetched once per warp (uniform 128-bit loads), not once
// per particle.
//
// Decisions are evaluated with un-fused IEEE fp64 mul/add so that they agree
// bit-for-bit with a CPU evaluation; only the accepted-force arithmetic uses
// FMA / rsqrt (accelerations are compared to 1e-6 relative, far above that).
#include "engine.h"
#include "../data/shortrange_table.h"
#include <math.h>

namespace b200 {

struct WalkPar {
    double box, halfbox;
    double rcut, rcut2;
    double theta2;
    double errtol_over_G_inv;   // unused placeholder to keep layout explicit
    double ErrTol, G;
    double h, h2, hinv, h3inv;
    double inv_cell_dx;         // 1 / (cellsize * table spacing)
    int usebh;
    int ntargets;
};

__device__ __forceinline__ double nearest(double x, double box, double halfbox)   // NEAREST partmanager.h:99
{
    return (x > halfbox) ? (x - box) : ((x < -halfbox) ? (x + box) : x);
}

// apply_accn_to_output (gravshort-tree.c:158-193) with the tabulated window of
// grav_apply_short_range_window (gravity.c:54-66).
__device__ __forceinline__ void monopole(double dx, double dy, double dz, double r2, double m,
                                         const WalkPar &P, const float *__restrict__ tab,
                                         double &ax, double &ay, double &az, double &pot)
{
    double r, fac, facpot;
    if(r2 >= P.h2) {
        const double rinv = rsqrt(r2);
        r = r2 * rinv;
        const double rinv2 = rinv * rinv;
        fac = m * rinv * rinv2;
        facpot = -m * rinv;
    } else {
        r = sqrt(r2);
        const double u = r * P.hinv;
        double wp;
        if(u < 0.5) {
            fac = m * P.h3inv * (10.666666666667 + u * u * (32.0 * u - 38.4));
            wp = -2.8 + u * u * (5.333333333333 + u * u * (6.4 * u - 9.6));
        } else {
            const double u3 = u * u * u;
            fac = m * P.h3inv * (21.333333333333 - 48.0 * u + 38.4 * u * u - 10.666666666667 * u3 - 0.066666666667 / u3);
            wp = -3.2 + 0.066666666667 / u + u * u * (10.666666666667 + u * (-16.0 + u * (9.6 - 2.133333333333 * u)));
        }
        facpot = m * P.hinv * wp;
    }
    const double ti = r * P.inv_cell_dx;
    const int t = (int) ti;                 // ti >= 0: truncation == floor
    if(t >= B200_SR_NTAB - 1) return;       // gravity.c:60-61: contribution dropped
    const double w1 = ti - (double) t, w0 = 1.0 - w1;
    fac *= w0 * (double) tab[t] + w1 * (double) tab[t + 1];
    facpot *= w0 * (double) tab[B200_SR_NTAB + t] + w1 * (double) tab[B200_SR_NTAB + t + 1];
    ax += dx * fac; ay += dy * fac; az += dz * fac;
    pot += facpot;
}

template <bool COUNT>
__global__ void __launch_bounds__(128)
k_grav_walk(const double4 *__restrict__ nodeA, const double4 *__restrict__ nodeB,
            const int4 *__restrict__ nodeC, const double4 *__restrict__ spart,
            const int *__restrict__ targets,     // original indices of the walk targets
            const double *__restrict__ pos, const float *__restrict__ mass,
            const double *__restrict__ oldacc, const float *__restrict__ gtab,
            WalkPar P, int full_tree, double cbrtrho0,
            double *__restrict__ acc_out, double *__restrict__ pot_out, int4 *__restrict__ counts_out)
{
    __shared__ float tab[2 * B200_SR_NTAB];
    for(int k = threadIdx.x; k < 2 * B200_SR_NTAB; k += blockDim.x) tab[k] = gtab[k];
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int group = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int tslot = group * 32 + lane;
    const bool valid = tslot < P.ntargets;
    if(group * 32 >= P.ntargets) return;    // warp-uniform

    int me = -1;
    double px = 0, py = 0, pz = 0, aold = 0;
    if(valid) {
        me = targets[tslot];
        px = pos[3 * (int64_t) me]; py = pos[3 * (int64_t) me + 1]; pz = pos[3 * (int64_t) me + 2];
        // grav_get_abs_accel gravshort.h:69-86 (sqrt(..)/G), then * ErrTolForceAcc (gravshort-tree.c:264)
        aold = __dmul_rn(P.ErrTol, __ddiv_rn(oldacc[me], P.G));
    }
    double ax = 0, ay = 0, az = 0, pot = 0;
    int n_acc = 0, n_open = 0, n_disc = 0, n_part = 0;

    const int NONE = -2;
    int resume = NONE;        // node at which this lane wakes up again
    int cur = 0;
    while(cur >= 0) {
        if(resume == cur) resume = NONE;
        const bool awake = valid && (resume == NONE);
        const double4 A = nodeA[cur];      // cofm, mass
        const double4 B = nodeB[cur];      // center, len
        const int4 C = nodeC[cur];         // sibling, pstart, count, leaf
        int decision = 0;                   // 0 discard, 1 accept, 2 open
        double dx = 0, dy = 0, dz = 0, r2 = 0;
        if(awake) {
            dx = nearest(A.x - px, P.box, P.halfbox);
            dy = nearest(A.y - py, P.box, P.halfbox);
            dz = nearest(A.z - pz, P.box, P.halfbox);
            r2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
            const double cxd = fabs(nearest(B.x - px, P.box, P.halfbox));
            const double cyd = fabs(nearest(B.y - py, P.box, P.halfbox));
            const double czd = fabs(nearest(B.z - pz, P.box, P.halfbox));
            const double len = B.w;
            bool discard = false;
            if(r2 > P.rcut2) {              // shall_we_discard_node gravshort-tree.c:198-215
                const double eff = __dadd_rn(P.rcut, __dmul_rn(0.5, len));
                discard = (cxd > eff) || (cyd > eff) || (czd > eff);
            }
            if(discard) {
                decision = 0;
            } else {                        // shall_we_open_node gravshort-tree.c:220-241
                bool open = false;
                const double l2 = __dmul_rn(len, len);
                if(P.usebh == 0) {
                    const double lhs = __dmul_rn(__dmul_rn(A.w, len), len);
                    const double rhs = __dmul_rn(__dmul_rn(r2, r2), aold);
                    open = lhs > rhs;
                }
                if(!open) {
                    // len*len/r2 > theta2, evaluated without the division unless within rounding of the threshold
                    const double rhs = __dmul_rn(P.theta2, r2);
                    if(l2 > rhs * (1.0 + 1e-14)) open = true;
                    else if(l2 >= rhs * (1.0 - 1e-14)) open = __ddiv_rn(l2, r2) > P.theta2;
                }
                if(!open) {
                    const double inside = __dmul_rn(0.6, len);
                    open = (cxd < inside) && (cyd < inside) && (czd < inside);
                }
                decision = open ? 2 : 1;
            }
        }
        const bool wantopen = awake && decision == 2;
        const unsigned openmask = __ballot_sync(0xffffffffu, wantopen);
        if(awake && decision == 1) {
            monopole(dx, dy, dz, r2, A.w, P, tab, ax, ay, az, pot);
            if(COUNT) n_acc++;
        }
        if(COUNT && awake && decision == 0) n_disc++;
        if(openmask == 0) { cur = C.x; continue; }
        if(C.w) {
            // particle leaf: opened lanes sum its particles directly (gravshort-tree.c:344-352,364-374)
            const int ps = C.y, cnt = C.z;
            for(int k = 0; k < cnt; k++) {
                const double4 q = spart[ps + k];
                if(wantopen) {
                    const double qx = nearest(q.x - px, P.box, P.halfbox);
                    const double qy = nearest(q.y - py, P.box, P.halfbox);
                    const double qz = nearest(q.z - pz, P.box, P.halfbox);
                    const double q2 = __dadd_rn(__dadd_rn(__dmul_rn(qx, qx), __dmul_rn(qy, qy)), __dmul_rn(qz, qz));
                    monopole(qx, qy, qz, q2, q.w, P, tab, ax, ay, az, pot);
                }
            }
            if(COUNT && wantopen) n_part += cnt;
            cur = C.x;
        } else {
            if(awake && !wantopen) resume = C.x;
            if(COUNT && wantopen) n_open++;
            cur = cur + 1;
        }
    }
    if(valid) {
        // grav_short_postprocess gravshort.h:47-67
        if(acc_out) {
            acc_out[3 * (int64_t) me] = ax * P.G;
            acc_out[3 * (int64_t) me + 1] = ay * P.G;
            acc_out[3 * (int64_t) me + 2] = az * P.G;
        }
        if(pot_out) {
            double p = pot;
            if(full_tree) {
                const double m = (double) mass[me];
                p += m / (P.h / 2.8);
                p -= 2.8372975 * pow(m, 2.0 / 3) * cbrtrho0;
                p *= P.G;
            }
            pot_out[me] = p;
        }
        if(COUNT) counts_out[me] = make_int4(n_acc, n_open, n_disc, n_part);
    }
}

__global__ void k_iota(int *p, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i < n) p[i] = i;
}

int walk_init_tables(Engine *E)
{
    CK(E->srtab.ensure(2 * B200_SR_NTAB));
    std::vector<float> t(2 * B200_SR_NTAB);
    for(int i = 0; i < B200_SR_NTAB; i++) { t[i] = b200_sr_force[i]; t[B200_SR_NTAB + i] = b200_sr_pot[i]; }
    CK(cudaMemcpyAsync(E->srtab.p, t.data(), t.size() * sizeof(float), cudaMemcpyHostToDevice, E->stream));
    CK(cudaStreamSynchronize(E->stream));
    return 0;
}

int grav_short_tree(Engine *E, const b200_gravshort_params *par, const int32_t *d_active,
                    int64_t nactive, double *d_acc, double *d_pot, b200_walk_counts *d_counts)
{
    if(!E->tree_valid) return failmsg(E, "b200_grav_short_tree: tree moments not computed (call b200_tree_build)");   // gravshort-tree.c:113-114
    if(E->Nmesh == 0) return failmsg(E, "b200_grav_short_tree: call b200_pm_init first (needs Nmesh, Asmth, G)");
    if(!E->srtab.p) if(int rc = walk_init_tables(E)) return rc;
    WalkPar P;
    P.box = E->tree_box; P.halfbox = 0.5 * E->tree_box;
    const double cellsize = E->tree_box / E->Nmesh;                 // gravshort-tree.c:101
    P.rcut = par->Rcut * E->Asmth * cellsize;                       // :102
    P.rcut2 = P.rcut * P.rcut;
    P.usebh = par->TreeUseBH;
    P.theta2 = par->BHOpeningAngle * par->BHOpeningAngle;           // :266-270
    if(P.usebh == 0) P.theta2 = par->MaxBHOpeningAngle * par->MaxBHOpeningAngle;
    P.ErrTol = par->ErrTolForceAcc; P.G = E->G; P.errtol_over_G_inv = 0;
    P.h = 2.8 * par->GravitySoftening;                              // FORCE_SOFTENING :37-41
    P.h2 = P.h * P.h; P.hinv = 1.0 / P.h; P.h3inv = 1.0 / P.h / P.h / P.h;
    P.inv_cell_dx = 1.0 / (cellsize * (double) B200_SR_DX);
    const double cbrtrho0 = pow(par->rho0, 1.0 / 3);

    // Targets: the tree's own particles in curve order when the walk set is the
    // tree set (the usual case); otherwise the caller's list as given.
    const int *tg = nullptr;
    int64_t nt = 0;
    if(d_active == nullptr && E->tree_full) { tg = E->sidx.p; nt = E->tree_np; }
    else if(d_active == nullptr) {
        CK(E->targets.ensure(E->n > 0 ? E->n : 1));
        if(E->n > 0) { k_iota<<<(unsigned) ((E->n + 255) / 256), 256, 0, E->stream>>>(E->targets.p, (int) E->n); CKL(E); }
        tg = E->targets.p; nt = E->n;
    } else { tg = d_active; nt = nactive; }
    P.ntargets = (int) nt;
    if(nt == 0) return 0;

    timer_start(E, T_WALK);
    const int bs = 128;
    const int64_t nwarps = (nt + 31) / 32;
    const unsigned nb = (unsigned) ((nwarps * 32 + bs - 1) / bs);
    if(d_counts)
        k_grav_walk<true><<<nb, bs, 0, E->stream>>>((const double4 *) E->nodeA.p, (const double4 *) E->nodeB.p, (const int4 *) E->nodeC.p,
                                                   (const double4 *) E->spart.p, tg, E->pos.p, E->mass.p, E->oldacc.p, E->srtab.p,
                                                   P, E->tree_full ? 1 : 0, cbrtrho0, d_acc, d_pot, (int4 *) d_counts);
    else
        k_grav_walk<false><<<nb, bs, 0, E->stream>>>((const double4 *) E->nodeA.p, (const double4 *) E->nodeB.p, (const int4 *) E->nodeC.p,
                                                    (const double4 *) E->spart.p, tg, E->pos.p, E->mass.p, E->oldacc.p, E->srtab.p,
                                                    P, E->tree_full ? 1 : 0, cbrtrho0, d_acc, d_pot, nullptr);
    CKL(E);
    timer_stop(E, T_WALK);
    return 0;
}

} // namespace b200
