// tree_walk.cu -- short-range Barnes-Hut tree gravity (replaces grav_short_tree,
// libgadget/gravshort-tree.c:96-154, i.e. treewalk_run + ev_primary
// (treewalk.c:266-310,801-902) with the visitor force_treeev_shortrange
// (gravshort-tree.c:253-379)).
//
// Two kernels.
//
// k_grav_walk: one warp walks the tree for 32 targets that are adjacent on the
// space-filling curve.  The warp follows the reference's depth-first
// sibling/first-child order through a batched frontier; every lane takes ITS OWN
// discard / accept / open decision with the reference's criteria
// (gravshort-tree.c:198-241), so per-particle results and interaction counts are
// those of the per-particle CPU walk.  A subtree is entered when any lane opens
// the node (warp ballot).  Node rows are fetched once per warp, not once per
// particle.  Accepted nodes are applied at once; an opened particle leaf is NOT
// evaluated here: every lane that opened it appends the leaf piece
// (first particle, count <= 8) to its own list in a global chunk pool, laid out
// [slot][lane] so that lanes with equal list lengths write one 128-byte line.
//
// k_grav_pairs: the same warp/target assignment.  For each of its 32 targets the
// warp reads that target's piece list and sums the pairs with the lanes spread
// over SOURCE particles (4 pieces x 8 particle slots, two sets interleaved and
// the next two already in flight), then one warp reduction hands the sums to
// the target's lane (gravshort-tree.c:364-374 sums the same pairs).  The pair
// loop therefore runs at the number of pairs each target needs, not at the union
// over the warp (2.8x larger on a 256^3 box), and it owns the whole register
// file instead of sharing it with the walk state.
//
// Decisions are evaluated with un-fused IEEE fp64 mul/add so that they agree
// bit-for-bit with a CPU evaluation; only the accepted-force arithmetic uses
// FMA / a fix-up-free rsqrt (accelerations are compared to 1e-6 relative, far above that).
#include "piece_list.cuh"
#include "../data/shortrange_table.h"
#include <math.h>
#include <stdlib.h>
#include <string>
#include <string.h>
#include <cub/device/device_select.cuh>
#include <cub/device/device_radix_sort.cuh>

namespace b200 {

struct WalkPar {
    double box, halfbox;
    double rcut, rcut2;
    double theta2;
    double wrap_lo, wrap_hi;    // targets inside [wrap_lo, wrap_hi]^3 never need the periodic wrap
    double ErrTol, G;
    double h, h2, hinv, h3inv;
    double h2s;                 // max(h2, smallest normal): pairs with r2 < h2s take the general path
    int sentinel;               // first of the >= 16 far-away massless rows behind spart[np)
    double inv_cell_dx;         // 1 / (cellsize * table spacing)
    int usebh;
    int ntargets;
    int f32ok;                  // the fp32 pre-classification of the walk may be used (box within float range)
    float rcut2f, th2lo, th2hi; // fp32 thresholds of that pre-classification (theta^2 widened by 2e-6 either way)
};

__device__ __forceinline__ double nearest(double x, double box, double halfbox)   // NEAREST partmanager.h:99
{
    return (x > halfbox) ? (x - box) : ((x < -halfbox) ? (x + box) : x);
}

// Window-table accessors.  row(t) = {T[t], T[t+1]-T[t], Tpot[t], Tpot[t+1]-Tpot[t]} so that the
// linear interpolation of grav_apply_short_range_window (gravity.c:54-66) is one FMA.
// TabD4: one double4 row per entry (walk kernel, few lookups).
// TabF4x8: the pair kernel's layout.  The table values ARE floats (gravity.c:20), so a row
// is stored as float4 {T[t], T[t+1], Tpot[t], Tpot[t+1]} and widened in registers (the
// differences of two floats are exact in double: same values as TabD4).  Eight copies,
// copy j in 16-byte bank group j: the eight lanes of a quarter warp (j = lane & 7) never
// conflict however their rows differ.  One LDS.128 per pair instead of two conflicting ones.
struct TabD4 {
    const double4 *p;
    __device__ __forceinline__ double4 row(int t) const { return p[t]; }
};
struct TabF2 {            // walk kernel: {T, Tpot} floats, two rows per lookup
    const float2 *p;
    __device__ __forceinline__ double4 row(int t) const
    {
        const float2 a = p[t], b = p[t + 1 < B200_SR_NTAB ? t + 1 : t];
        const double f0 = (double) a.x, p0 = (double) a.y;
        return make_double4(f0, (double) b.x - f0, p0, (double) b.y - p0);
    }
};
struct TabF4x8 {
    const float4 *p;        // already offset by the lane's copy: p = base + (lane & 7)
    __device__ __forceinline__ double4 row(int t) const
    {
        const float4 f = p[t * 8];
        const double f0 = (double) f.x, p0 = (double) f.z;
        return make_double4(f0, (double) f.y - f0, p0, (double) f.w - p0);
    }
};

// apply_accn_to_output (gravshort-tree.c:158-193) with the tabulated window of
// grav_apply_short_range_window (gravity.c:54-66).  tab[t] = {T[t], T[t+1]-T[t],
// Tpot[t], Tpot[t+1]-Tpot[t]} so that the linear interpolation is one FMA.
template <class TAB>
__device__ __forceinline__ void monopole(double dx, double dy, double dz, double r2, double m,
                                         const WalkPar &P, const TAB &tab,
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
    const double w1 = ti - (double) t;
    const double4 e = tab.row(t);
    fac *= fma(w1, e.y, e.x);
    facpot *= fma(w1, e.w, e.z);
    ax = fma(dx, fac, ax); ay = fma(dy, fac, ay); az = fma(dz, fac, az);
    pot += facpot;
}

#ifndef WALK_MINB
#define WALK_MINB 6
#endif
#ifndef PAIR_MINB
#define PAIR_MINB 2
#endif
#define PAIR_WARPS 8

// Per-lane decision on one node: 0 discard, 1 accept, 2 open
// (shall_we_discard_node / shall_we_open_node, gravshort-tree.c:198-241).
// WRAP=false is used when every target of the warp is farther than the table
// reach from all box faces: then NEAREST cannot change any decision (a node
// whose wrapped and unwrapped distances differ is discarded either way) nor any
// accepted distance, so the selects are skipped.
// {rcut + len/2, len^2, (mass*len)*len, 0.6*len} (gravshort-tree.c:204,226,231,236)
__device__ __forceinline__ double4 node_consts(const double4 &A, const double4 &B, const WalkPar &P)
{
    const double len = B.w;
    return make_double4(__dadd_rn(P.rcut, __dmul_rn(0.5, len)), __dmul_rn(len, len), __dmul_rn(__dmul_rn(A.w, len), len),
                        __dmul_rn(0.6, len));
}

template <bool WRAP>
__device__ __forceinline__ int classify(const double4 &A, const double4 &B, const double4 &D, double px, double py, double pz,
                                        double aold, const WalkPar &P, double &dx, double &dy, double &dz, double &r2)
{
    // D = node_consts(A, B): the per-node products of the criteria, formed once by the lane that
    // fetched the node instead of by every lane (same un-fused operations, same values)
    // Straight-line (two of these are interleaved by the caller); the same comparisons in the
    // same arithmetic as the early-exit form of the reference.
    dx = A.x - px; dy = A.y - py; dz = A.z - pz;
    double cxd = B.x - px, cyd = B.y - py, czd = B.z - pz;
    if(WRAP) {
        dx = nearest(dx, P.box, P.halfbox); dy = nearest(dy, P.box, P.halfbox); dz = nearest(dz, P.box, P.halfbox);
        cxd = nearest(cxd, P.box, P.halfbox); cyd = nearest(cyd, P.box, P.halfbox); czd = nearest(czd, P.box, P.halfbox);
    }
    cxd = fabs(cxd); cyd = fabs(cyd); czd = fabs(czd);
    r2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    const double eff = D.x, l2 = D.y, inside = D.w;
    const bool disc = (r2 > P.rcut2) & ((cxd > eff) | (cyd > eff) | (czd > eff));          // shall_we_discard_node
    const bool orel = (P.usebh == 0) & (D.z > __dmul_rn(__dmul_rn(r2, r2), aold));
    // len*len/r2 > theta2, evaluated without the division unless within rounding of the threshold
    const double rhs = __dmul_rn(P.theta2, r2);
    const bool obh = l2 > rhs * (1.0 + 1e-14);
    const bool nearbh = (!obh) & (l2 >= rhs * (1.0 - 1e-14));
    const bool oin = (cxd < inside) & (cyd < inside) & (czd < inside);
    bool open = orel | obh | oin;
    if(nearbh & !open & !disc) open = __ddiv_rn(l2, r2) > P.theta2;                          // practically never
    return disc ? 0 : (open ? 2 : 1);
}

// 1/sqrt(x) for normal positive x: the hardware seed (MUFU.RSQ64H, ~2^-22) followed by one
// third-order correction y(1 + e/2 + 3e^2/8), e = 1 - x y^2; relative error < 2^-52.  Without
// the zero/denormal/infinity fix-up branch of rsqrt(): the callers never pass those.
__device__ __forceinline__ double rsqrt_pos(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-(x * y), y, 1.0);
    return fma(y * e, fma(e, 0.375, 0.5), y);
}

// Newtonian-regime pair with the window applied, branch-free (the caller
// guarantees r2 >= h^2 > 0 or passes m = 0, r2 = 1).  Row NTAB-1 of `tab` is all
// zero, so a pair beyond the table (gravity.c:60-61: contribution dropped) needs
// no select, only the index clamp.
__device__ __forceinline__ void pair_fast(double dx, double dy, double dz, double r2, double m,
                                          const WalkPar &P, const TabF4x8 &tab,
                                          double &ax, double &ay, double &az, double &pot)
{
    const double rinv = rsqrt_pos(r2);
    const double r = r2 * rinv;
    const double mr = m * rinv;
    const double ti = r * P.inv_cell_dx;
    const int t = min((int) ti, B200_SR_NTAB - 1);
    const double w1 = ti - (double) t;
    const double4 e = tab.row(t);
    const double wf = fma(w1, e.y, e.x), wp = fma(w1, e.w, e.z);
    const double fac = mr * rinv * rinv * wf;
    ax = fma(dx, fac, ax); ay = fma(dy, fac, ay); az = fma(dz, fac, az);
    pot = fma(-mr, wp, pot);
}

// An accepted node in the walk kernel: apply_accn_to_output with the {T, Tpot} float table.  Newtonian regime
// straight-line; inside the softening radius (practically never for an accepted node) the general path.
__device__ __forceinline__ void node_term(double dx, double dy, double dz, double r2, double m,
                                          const WalkPar &P, const float2 *tab,
                                          double &ax, double &ay, double &az, double &pot)
{
    if(!(r2 >= P.h2s)) { monopole(dx, dy, dz, r2, m, P, TabF2{tab}, ax, ay, az, pot); return; }
    const double rinv = rsqrt_pos(r2);
    const double r = r2 * rinv;
    const double mr = m * rinv;
    const double ti = r * P.inv_cell_dx;
    const int t = min((int) ti, B200_SR_NTAB - 2);
    const double w1 = ti - (double) t;
    const float2 a = tab[t], b = tab[t + 1];
    const double f0 = (double) a.x, p0 = (double) a.y;
    const double wf = fma(w1, (double) b.x - f0, f0), wp = fma(w1, (double) b.y - p0, p0);
    const double fac = ti < (double) (B200_SR_NTAB - 1) ? mr * rinv * rinv * wf : 0.0;     // gravity.c:60-61: dropped beyond the table
    ax = fma(dx, fac, ax); ay = fma(dy, fac, ay); az = fma(dz, fac, az);
    pot = ti < (double) (B200_SR_NTAB - 1) ? fma(-mr, wp, pot) : pot;
}

// One source row of a leaf piece for this lane's slot.  Slots past the count read
// the rows that follow (spart is padded with far-away massless rows) and are
// neutralised through the mass.
struct SrcRow { double4 q; int cnt; };

// The source rows are read from two 16-byte streams {x, y} and {z, m}: the eight lanes of a
// group then read 128 contiguous bytes per load (whole sectors) instead of half of 8 x 32.
struct SrcArrays { const double2 *__restrict__ xy; const double2 *__restrict__ zm; };

// The pair loop advances 16 pieces (= one chunk row block) per step (piece_list.cuh).  Three
// stages are in flight: the 4 entries of the next step, the 2 source rows of the next half
// step, and the arithmetic of the current half step.
struct Rows2 { SrcRow r[2]; };

__device__ __forceinline__ Rows2 fetch_rows(const Ent4 &E, int first, int slot, const SrcArrays &S)
{
    Rows2 R;
#pragma unroll
    for(int k = 0; k < 2; k++) {
        const unsigned e = E.e[first + k];
        const unsigned j = (e >> 4) + slot;
        const double2 a = S.xy[j], b = S.zm[j];
        R.r[k].q = make_double4(a.x, a.y, b.x, b.y);
        R.r[k].cnt = (int) (e & 15u);
    }
    return R;
}

template <bool WRAP>
__device__ __forceinline__ void row_delta(const SrcRow &s, int slot, const WalkPar &P, double tx, double ty, double tz,
                                          double &qx, double &qy, double &qz, double &q2, double &qm)
{
    qx = s.q.x - tx; qy = s.q.y - ty; qz = s.q.z - tz;
    if(WRAP) {
        qx = nearest(qx, P.box, P.halfbox); qy = nearest(qy, P.box, P.halfbox); qz = nearest(qz, P.box, P.halfbox);
    }
    qm = slot < s.cnt ? s.q.w : 0.0;
    q2 = fma(qz, qz, fma(qy, qy, qx * qx));
}

template <bool WRAP>
__device__ __forceinline__ void pair_step(const SrcRow &ca, const SrcRow &cb, int slot, const WalkPar &P,
                                          const TabF4x8 &tab, double tx, double ty, double tz,
                                          double &sx, double &sy, double &sz, double &sp)
{
    double ax_, ay_, az_, a2, am, bx_, by_, bz_, b2, bm;
    row_delta<WRAP>(ca, slot, P, tx, ty, tz, ax_, ay_, az_, a2, am);
    row_delta<WRAP>(cb, slot, P, tx, ty, tz, bx_, by_, bz_, b2, bm);
    const bool softa = !(a2 >= P.h2s), softb = !(b2 >= P.h2s);
    if(__any_sync(0xffffffffu, softa || softb)) {
        if(softa) { monopole(ax_, ay_, az_, a2, am, P, tab, sx, sy, sz, sp); am = 0.0; a2 = 1.0; }
        if(softb) { monopole(bx_, by_, bz_, b2, bm, P, tab, sx, sy, sz, sp); bm = 0.0; b2 = 1.0; }
    }
    pair_fast(ax_, ay_, az_, a2, am, P, tab, sx, sy, sz, sp);
    pair_fast(bx_, by_, bz_, b2, bm, P, tab, sx, sy, sz, sp);
}

// Pairs inside the softening radius (the target itself among them) are rare and take the
// general path; everything else is branch-free, two pieces interleaved.
template <bool WRAP>
__device__ __forceinline__ void pair_sum(const PieceList &L, int g, int slot,
                                         const SrcArrays &spart, const WalkPar &P,
                                         const TabF4x8 &tab, double tx, double ty, double tz,
                                         double &sx, double &sy, double &sz, double &sp)
{
    Ent4 eC = fetch_ent(L, 0, g), eN = fetch_ent(L, 16, g);
    Rows2 r01 = fetch_rows(eC, 0, slot, spart);
    for(int base = 0; base < L.nt; base += 16) {
        const Rows2 r23 = fetch_rows(eC, 2, slot, spart);
        pair_step<WRAP>(r01.r[0], r01.r[1], slot, P, tab, tx, ty, tz, sx, sy, sz, sp);
        eC = eN;
        eN = fetch_ent(L, base + 32, g);
        r01 = fetch_rows(eC, 0, slot, spart);
        if(base + 8 < L.nt) pair_step<WRAP>(r23.r[0], r23.r[1], slot, P, tab, tx, ty, tz, sx, sy, sz, sp);
    }
}

// Staged node rows of the current batch (one entry per lane).  The fp64 rows are the reference's
// operands; the fp32 rows are the same vectors taken relative to the centre of the warp's bounding box
// (after NEAREST), which is what the per-lane pre-classification reads.
struct BatchEntry {
    double4 A[32];    // cofm, mass
    float4 C[32];     // centre - bbox centre, len
    float4 F[32];     // cofm - bbox centre, (mass*len)*len
    float4 T[32];     // error bounds {tau2 on r2, mu on |centre - p|}, rcut + len/2, 0.6 len
    int4 M[32];       // pstart, count, mask of lanes (targets) that opened every ancestor,
                      // flags (bit0: leaf, bit1: rejected for all lanes by the bounding-box test,
                      //        bit2: fp32 rows unusable (box seam, underflow): exact classification only)
    int N[32];        // node index
};

// Per-lane constants of the fp32 pre-classification.
struct Lane32 {
    float px, py, pz;       // position - bbox centre
    float alo, ahi;         // bounds of ErrTolForceAcc * |a_old| / G
};
struct Walk32 {
    float rcut2, th2lo, th2hi;
    bool rel;
};

// Per-warp staging of the pair-wise phase 1: the targets' fp32 constants (read by whichever lane works on a pair
// of that target) and the decisions as bit sets: acc/opn/dis[l] = batch slots target l accepted / opened /
// discarded, openers[k] = targets that opened slot k, first[k] = index of slot k's first pair.
struct PairStage {
    float4 t32[32];     // position - bbox centre, alo
    float ahi[32];
    unsigned acc[32], opn[32], dis[32], openers[32];
    int first[32];
};

// position of the r-th (0-based) set bit of m (r < popc(m)): the byte by two population counts, the bit inside it
// from a table [byte value][r] in shared memory
__device__ __forceinline__ int nth_set_bit(unsigned m, int r, const unsigned char *lut)
{
    int pos = 0;
    int c = __popc(m & 0xffffu); if(r >= c) { r -= c; pos = 16; m >>= 16; }
    c = __popc(m & 0xffu); if(r >= c) { r -= c; pos += 8; m >>= 8; }
    return pos + lut[(m & 0xffu) * 8 + r];
}

// Interval version of classify(): every comparison of shall_we_discard_node / shall_we_open_node
// (gravshort-tree.c:198-241) is evaluated at both ends of the rounding-error interval of its fp32
// operands (r2 in [r2 - tau2, r2 + tau2], max |centre - p| in [c - mu, c + mu], the thresholds widened by
// 1e-6 relative).  `discard` increases with both, `open` decreases with both, so when the two ends agree the
// fp64 decision is known; otherwise 3 = undecided and the caller repeats the node in fp64.  The bounds:
// operands are fp32 roundings (u = 2^-24) of fp64 vectors relative to the bbox centre, |p'| <= hb, hence
// |err(dx)| <= u (2|dx| + 2 hb), |err(r2)| <= u (10.5 r2 + 3.5 hb^2) <= tau2 := 1e-6 ((|m'| + hb)^2 + hb^2) + ...,
// |err(c)| <= u (2 |c'| + 4 hb) <= mu (hb = half diagonal of the box).
__device__ __forceinline__ int classify32(const float4 C, const float4 F, const float4 T, const Lane32 &L, const Walk32 &W)
{
    const float dx = F.x - L.px, dy = F.y - L.py, dz = F.z - L.pz;
    const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
    const float cm = fmaxf(fmaxf(fabsf(C.x - L.px), fabsf(C.y - L.py)), fabsf(C.z - L.pz));
    const float r2lo = fmaxf(r2 - T.x, 0.f), r2hi = r2 + T.x;
    const float clo = cm - T.y, chi = cm + T.y;
    const bool dlo = (r2lo > W.rcut2) & (clo > T.z), dhi = (r2hi > W.rcut2) & (chi > T.z);
    const float l2 = C.w * C.w;
    const bool oup = (W.rel & (F.w > r2lo * r2lo * L.alo)) | (l2 > W.th2lo * r2lo) | (clo < T.w);   // open => oup
    const bool odn = (W.rel & (F.w > r2hi * r2hi * L.ahi)) | (l2 > W.th2hi * r2hi) | (chi < T.w);   // odn => open
    const bool undecided = (dlo != dhi) | ((!dhi) & (oup != odn));
    return undecided ? 3 : (dhi ? 0 : (oup ? 2 : 1));
}

// The fp64 decision for one node, out of line: taken only when the fp32 intervals straddle a threshold.
__device__ __noinline__ int classify_exact(const double4 A, const double4 B, double px, double py, double pz, double aold,
                                           const WalkPar &P, bool wrap)
{
    double dx, dy, dz, r2;
    const double4 D = node_consts(A, B, P);
    return wrap ? classify<true>(A, B, D, px, py, pz, aold, P, dx, dy, dz, r2) : classify<false>(A, B, D, px, py, pz, aold, P, dx, dy, dz, r2);
}

template <bool COUNT, bool MERGE>
__global__ void __launch_bounds__(128, WALK_MINB)
k_grav_walk(const double4 *__restrict__ nodeA, const double4 *__restrict__ nodeB,
            const int4 *__restrict__ nodeC, const int4 *__restrict__ nodeK, const double4 *__restrict__ spart,
            const int *__restrict__ targets,     // original indices of the walk targets
            const double *__restrict__ pos, const float *__restrict__ mass,
            const double *__restrict__ oldacc, const float *__restrict__ gtab,
            WalkPar P,
            PiecePool Q,                                    // piece lists (piece_list.cuh)
            double4 *__restrict__ partial,                  // [target slot] sums over accepted nodes {ax, ay, az, pot}
            int4 *__restrict__ counts_out)
{
    __shared__ float2 tab[B200_SR_NTAB];                // {T, Tpot} (gravity.c:20: the table IS float)
    extern __shared__ int s_ctab_dyn[];                 // [WALK_WARPS][Q.maxch]
    __shared__ int s_stk_node_all[WALK_WARPS][WALK_STACK];
    __shared__ unsigned s_stk_mask_all[WALK_WARPS][WALK_STACK];
    __shared__ BatchEntry s_ent_all[WALK_WARPS];
    __shared__ double s_bbox_all[WALK_WARPS][8];        // bbox centre, half extents, half diagonal of the warp's targets
    __shared__ PairStage s_res_all[WALK_WARPS];
    __shared__ unsigned char s_nthbit[256 * 8];         // nth_set_bit
    for(int k = threadIdx.x; k < B200_SR_NTAB; k += blockDim.x)
        tab[k] = make_float2(gtab[k], gtab[B200_SR_NTAB + k]);       // monopole() never takes t >= NTAB-1 as the base row
    for(int k = threadIdx.x; k < 256 * 8; k += blockDim.x) {
        int b = k >> 3, r = k & 7, pos = 0;
        for(int i = 0; i < 8; i++) if((b >> i) & 1) { if(r == 0) { pos = i; break; } r--; }
        s_nthbit[k] = (unsigned char) pos;
    }
    __syncthreads();
    const int wib = threadIdx.x >> 5;
    int *s_ctab = s_ctab_dyn + wib * Q.maxch;
    int *s_stk_node = s_stk_node_all[wib];
    unsigned *s_stk_mask = s_stk_mask_all[wib];
    BatchEntry &s_ent = s_ent_all[wib];
    double *s_bbox = s_bbox_all[wib];
    PairStage &s_res = s_res_all[wib];
    const unsigned ltmask = (1u << (threadIdx.x & 31)) - 1u;
    int mycnt = 0;          // pieces in this lane's list
    unsigned mylast = 0;    // its last entry (merged with the next piece when contiguous; stored when the next one starts)
    int nch_alloc = 0;      // chunks this warp owns (warp-uniform)

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
    Lane32 L32;
    bool warp_central;
    const unsigned validmask = __ballot_sync(0xffffffffu, valid);
    {
        // bounding box (centre, half extent) of the warp's targets, kept in shared memory
        const double big = 1e300;
        const double lox = warp_min(valid ? px : big), hix = warp_max(valid ? px : -big);
        const double loy = warp_min(valid ? py : big), hiy = warp_max(valid ? py : -big);
        const double loz = warp_min(valid ? pz : big), hiz = warp_max(valid ? pz : -big);
        const double bcx = 0.5 * (lox + hix), bcy = 0.5 * (loy + hiy), bcz = 0.5 * (loz + hiz);
        const double bhx = 0.5 * (hix - lox), bhy = 0.5 * (hiy - loy), bhz = 0.5 * (hiz - loz);
        warp_central = lox >= P.wrap_lo && hix <= P.wrap_hi && loy >= P.wrap_lo && hiy <= P.wrap_hi &&
                       loz >= P.wrap_lo && hiz <= P.wrap_hi;
        if(lane == 0) {
            s_bbox[0] = bcx; s_bbox[1] = bcy; s_bbox[2] = bcz; s_bbox[3] = bhx; s_bbox[4] = bhy; s_bbox[5] = bhz;
            // bounds |p - bc| of every lane (plus a rounding allowance)
            s_bbox[6] = 1.0000001 * sqrt(bhx * bhx + bhy * bhy + bhz * bhz) + 1e-30;
        }
        // fp32 side: everything relative to the bbox centre
        L32.px = valid ? (float) (px - bcx) : 0.f; L32.py = valid ? (float) (py - bcy) : 0.f; L32.pz = valid ? (float) (pz - bcz) : 0.f;
        const float a32 = (float) aold;
        const bool tiny = !(a32 >= 1e-30f);
        L32.alo = tiny ? 0.f : a32 * (1.f - 2e-6f);
        L32.ahi = (tiny ? 1e-30f : a32) * (1.f + 2e-6f);
        s_res.t32[lane] = make_float4(L32.px, L32.py, L32.pz, L32.alo);
        s_res.ahi[lane] = L32.ahi;
    }
    Walk32 W32;
    W32.rcut2 = P.rcut2f; W32.th2lo = P.th2lo; W32.th2hi = P.th2hi;
    W32.rel = P.usebh == 0;

    double ax = 0, ay = 0, az = 0, pot = 0;
    int n_acc = 0, n_open = 0, n_disc = 0, n_part = 0;

    // The walk visits the same (target, node) pairs as the reference's depth-first
    // walk: a target reaches a node iff it opened every ancestor (the mask).  The per-warp
    // stack holds OPENED INTERNAL NODES (node, mask of the lanes that opened it), one entry
    // whatever the number of children; a batch is formed by expanding the topmost entries into
    // up to 32 children whose rows are then fetched in parallel, one node per lane.  (A stack
    // of children, 8 entries per opened node, filled up and throttled the batches to ~10 nodes.)
    int sp = 0, nb = 1;
    if(lane == 0) { s_ent.N[0] = 0; s_ent.M[0].z = (int) validmask; }
    __syncwarp();
    for(bool first = true;; first = false) {
        if(!first) {
            if(sp == 0) break;
            // ---- expansion: lane j looks at entry sp-1-j (the top 16); children are counted from nodeK
            int myc = 0, pnode = 0;
            unsigned pmask = 0;
            int4 k0 = make_int4(-1, -1, -1, -1), k1 = k0;
            if(lane < 16 && lane < sp) {
                pnode = s_stk_node[sp - 1 - lane]; pmask = s_stk_mask[sp - 1 - lane];
                k0 = nodeK[2 * (size_t) pnode]; k1 = nodeK[2 * (size_t) pnode + 1];
                myc = (k0.x >= 0) + (k0.y >= 0) + (k0.z >= 0) + (k0.w >= 0) + (k1.x >= 0) + (k1.y >= 0) + (k1.z >= 0) + (k1.w >= 0);
            }
            int S = myc;            // inclusive scan from the top of the stack downwards
#pragma unroll
            for(int o = 1; o < 16; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, S, o); if(lane >= o) S += v; }
            // deep in a narrow descent the batch is limited to one node's children so that the stack cannot overflow
            const int cap = sp > WALK_STACK - WALK_RESERVE ? 8 : 32;
            const int J = __popc(__ballot_sync(0xffffffffu, lane < 16 && lane < sp && S <= cap));      // >= 1 (a node has <= 8 children)
            nb = __shfl_sync(0xffffffffu, S, J - 1);
            if(lane < J) {
                int w = nb - S;     // entries deeper in the stack first: batch slots stay in curve order
                const int kids[8] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
#pragma unroll
                for(int c = 0; c < 8; c++) if(kids[c] >= 0) { s_ent.N[w] = kids[c]; s_ent.M[w].z = (int) pmask; w++; }
            }
            sp -= J;
            __syncwarp();
        }
        // ---- lane-parallel: lane l fetches batch entry l and tests it against the warp's bounding box; the
        // survivors are staged in consecutive slots 0..nl-1 (batch order kept)
        bool dead = true;
        unsigned mydisc = 0;            // lanes (targets) for which the box test discards the node this lane fetched
        unsigned emask = 0;
        int mynode = -1, eflags0 = 0;
        double4 eA = make_double4(0, 0, 0, 0), eB = eA;
        int4 C = make_int4(0, 0, 0, 0);
        double ex = 0, ey = 0, ez = 0, eff = 0;
        if(lane < nb) {
            mynode = s_ent.N[lane];
            const unsigned emask0 = (unsigned) s_ent.M[lane].z;
            eB = nodeB[mynode]; C = nodeC[mynode]; eA = nodeA[mynode];
            eflags0 = C.w ? 1 : 0;
            // Early discard for all lanes (gravshort-tree.c:198-215): along some axis the
            // node centre is farther than rcut + len/2 (+ rounding margin) from the whole
            // bounding box; the centre of mass lies inside the cell, so r2 > rcut^2 follows.
            eff = P.rcut + 0.5 * eB.w;
            const double lim = eff + 1e-9 * (eff + eB.w);
            ex = eB.x - s_bbox[0]; ey = eB.y - s_bbox[1]; ez = eB.z - s_bbox[2];
            if(!warp_central) {
                ex = nearest(ex, P.box, P.halfbox); ey = nearest(ey, P.box, P.halfbox); ez = nearest(ez, P.box, P.halfbox);
            }
            if(fabs(ex) - s_bbox[3] > lim || fabs(ey) - s_bbox[4] > lim || fabs(ez) - s_bbox[5] > lim) mydisc = emask0;
            emask = emask0 & ~mydisc;
            dead = emask == 0;
        }
        const unsigned livemask = __ballot_sync(0xffffffffu, !dead);      // every lane has read its entry by now
        const int nl = __popc(livemask);
        if(!dead) {
            const int ck = __popc(livemask & ltmask);
            const double bcx = s_bbox[0], bcy = s_bbox[1], bcz = s_bbox[2], hbd = s_bbox[6];
            s_ent.A[ck] = eA;
            double mx = eA.x - bcx, my = eA.y - bcy, mz = eA.z - bcz;
            if(!warp_central) {
                mx = nearest(mx, P.box, P.halfbox); my = nearest(my, P.box, P.halfbox); mz = nearest(mz, P.box, P.halfbox);
            }
            const double ml2 = __dmul_rn(__dmul_rn(eA.w, eB.w), eB.w);
            const double cinf = fmax(fmax(fabs(ex), fabs(ey)), fabs(ez));
            const double minf = fmax(fmax(fabs(mx), fabs(my)), fabs(mz));
            const double rmax = sqrt(mx * mx + my * my + mz * mz) + hbd;
            // a lane's NEAREST(node - p) equals (node - bc wrapped) - (p - bc) only away from the +-box/2 seam
            const bool seam = !warp_central && (fmax(cinf, minf) + hbd >= 0.999999 * P.halfbox);
            if(!P.f32ok || seam || (ml2 > 0.0 && ml2 < 1e-30)) eflags0 |= 4;
            s_ent.C[ck] = make_float4((float) ex, (float) ey, (float) ez, (float) eB.w);
            s_ent.F[ck] = make_float4((float) mx, (float) my, (float) mz, (float) ml2);
            s_ent.T[ck] = make_float4((float) (1e-6 * (rmax * rmax + hbd * hbd) + 4e-7 * P.rcut2),
                                      (float) (4e-7 * (cinf + hbd + eff + eB.w)),
                                      (float) eff, (float) __dmul_rn(0.6, eB.w));
            s_ent.M[ck] = make_int4(C.y, C.z, (int) emask, eflags0);
            s_ent.N[ck] = mynode;
        }
        if(COUNT) {
            for(unsigned m = __ballot_sync(0xffffffffu, mydisc != 0); m; m &= m - 1) {
                const unsigned e = __shfl_sync(0xffffffffu, mydisc, __ffs(m) - 1);
                if((e >> lane) & 1u) n_disc++;
            }
        }
        s_res.acc[lane] = 0; s_res.opn[lane] = 0; s_res.openers[lane] = 0;
        if(COUNT) s_res.dis[lane] = 0;
        __syncwarp();
        // ---- phase 1, the per-target decisions, over the (node, target) pairs that exist: node k of the batch is
        // tested only for the lanes of its mask (about half of them on average), 32 pairs at a time whatever
        // nodes they belong to.  Pair w of the batch -> node k by the prefix sums of the mask populations,
        // target l = the (w - first pair of k)-th set bit of the mask.  Decisions go to bit sets in shared memory.
        const int4 Mc = lane < nl ? s_ent.M[lane] : make_int4(0, 0, 0, 0);
        int T;
        int excl;
        {
            int incl = __popc((unsigned) Mc.z);
#pragma unroll
            for(int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if(lane >= o) incl += v; }
            T = __shfl_sync(0xffffffffu, incl, 31);
            excl = incl - __popc((unsigned) Mc.z);
            s_res.first[lane] = excl;
        }
        const unsigned leafmask = __ballot_sync(0xffffffffu, (Mc.w & 1) != 0);
        const bool bigleaf = (Mc.w & 1) && Mc.y > 8;
        __syncwarp();
        for(int base = 0; base < T; base += 32) {
            const int sj = excl - base;
            const unsigned startmask = __reduce_or_sync(0xffffffffu, (lane < nl && sj >= 1 && sj <= 31) ? 1u << sj : 0u);
            const int kfirst = __popc(__ballot_sync(0xffffffffu, lane < nl && excl <= base)) - 1;
            const bool act = base + lane < T;
            int dec = -1, k = 0, l = 0;
            if(act) {
                k = kfirst + __popc(startmask & (0xfffffffeu & (0xffffffffu >> (31 - lane))));
                const int4 M = s_ent.M[k];
                l = nth_set_bit((unsigned) M.z, base + lane - s_res.first[k], s_nthbit);
                const float4 tp = s_res.t32[l];
                Lane32 L; L.px = tp.x; L.py = tp.y; L.pz = tp.z; L.alo = tp.w; L.ahi = s_res.ahi[l];
                dec = classify32(s_ent.C[k], s_ent.F[k], s_ent.T[k], L, W32);
                if(M.w & 4) dec = 3;
                if(M.w & 1) dec |= 8;           // leaf: nobody needs the list of its openers
            }
            if(__any_sync(0xffffffffu, (dec & 7) == 3)) {
                // rare: the pair is within rounding of a threshold (or the node sits on the box seam):
                // the reference's own fp64 comparisons decide
                const double qx = __shfl_sync(0xffffffffu, px, l), qy = __shfl_sync(0xffffffffu, py, l), qz = __shfl_sync(0xffffffffu, pz, l);
                const double qa = __shfl_sync(0xffffffffu, aold, l);
                if((dec & 7) == 3) dec = (dec & 8) | classify_exact(s_ent.A[k], nodeB[s_ent.N[k]], qx, qy, qz, qa, P, !warp_central);
            }
            if(act) {
                if((dec & 7) == 1) atomicOr(&s_res.acc[l], 1u << k);
                else if((dec & 7) == 2) { atomicOr(&s_res.opn[l], 1u << k); if(!(dec & 8)) atomicOr(&s_res.openers[k], 1u << l); }
                else if(COUNT) atomicOr(&s_res.dis[l], 1u << k);
            }
        }
        __syncwarp();
        unsigned accbits = s_res.acc[lane], openbits = s_res.opn[lane];     // batch slots this lane accepted / opened
        const unsigned myopeners = s_res.openers[lane];                     // lane k: the lanes that opened slot k (internal nodes)
        mynode = lane < nl ? s_ent.N[lane] : -1;
        if(COUNT) n_disc += __popc(s_res.dis[lane]);
        // ---- phase 2, accepted nodes: monopole x tabulated window in fp64 (apply_accn_to_output), every lane
        // working through ITS OWN accepted slots (lanes accept different nodes; this keeps them all busy), two
        // slots per turn (independent chains)
        if(COUNT) n_acc += __popc(accbits);
        while(__any_sync(0xffffffffu, accbits != 0)) {
            if(accbits) {
                const int k0 = __ffs(accbits) - 1; accbits &= accbits - 1;
                const bool two = accbits != 0;
                const int k1 = two ? __ffs(accbits) - 1 : k0; accbits &= accbits - 1;
                const double4 A0 = s_ent.A[k0], A1 = s_ent.A[k1];
                double dx0 = A0.x - px, dy0 = A0.y - py, dz0 = A0.z - pz, dx1 = A1.x - px, dy1 = A1.y - py, dz1 = A1.z - pz;
                if(!warp_central) {
                    dx0 = nearest(dx0, P.box, P.halfbox); dy0 = nearest(dy0, P.box, P.halfbox); dz0 = nearest(dz0, P.box, P.halfbox);
                    dx1 = nearest(dx1, P.box, P.halfbox); dy1 = nearest(dy1, P.box, P.halfbox); dz1 = nearest(dz1, P.box, P.halfbox);
                }
                const double r20 = __dadd_rn(__dadd_rn(__dmul_rn(dx0, dx0), __dmul_rn(dy0, dy0)), __dmul_rn(dz0, dz0));
                const double r21 = __dadd_rn(__dadd_rn(__dmul_rn(dx1, dx1), __dmul_rn(dy1, dy1)), __dmul_rn(dz1, dz1));
                node_term(dx0, dy0, dz0, r20, A0.w, P, tab, ax, ay, az, pot);
                node_term(dx1, dy1, dz1, r21, two ? A1.w : 0.0, P, tab, ax, ay, az, pot);
            }
        }
        // ---- phase 3, opened particle leaves (gravshort-tree.c:344-352): every lane appends its own leaves, in
        // slot order, to its piece list; pieces of <= 8 particles (only leaves at the key-depth limit hold more)
        if(COUNT) { n_open += __popc(openbits & ~leafmask); }
        openbits &= leafmask;
        if(__ballot_sync(0xffffffffu, bigleaf) == 0) {
            // chunks for the longest list this batch can produce are reserved by one vote; the appends need none
            piece_reserve(__popc(openbits), mycnt, nch_alloc, s_ctab, Q, group, lane);
            while(openbits) {
                const int k = __ffs(openbits) - 1; openbits &= openbits - 1;
                const int4 M = s_ent.M[k];
                if(COUNT) n_part += M.y;
                piece_append<MERGE>(PIECE(M.x, M.y), mycnt, mylast, nch_alloc, s_ctab, Q, lane);
            }
            __syncwarp();
        } else
        while(__any_sync(0xffffffffu, openbits != 0)) {
            const bool want = openbits != 0;
            int pstart = 0, cnt = 0;
            if(want) {
                const int k = __ffs(openbits) - 1; openbits &= openbits - 1;
                const int4 M = s_ent.M[k];
                pstart = M.x; cnt = M.y;
                if(COUNT) n_part += cnt;
            }
            if(__any_sync(0xffffffffu, cnt > 8)) {
                const int maxc = (int) __reduce_max_sync(0xffffffffu, (unsigned) cnt);
                for(int o = 0; o < maxc; o += 8) {
                    const int c = cnt - o < 8 ? cnt - o : 8;
                    piece_push<MERGE>(want && o < cnt, PIECE(pstart + o, c > 0 ? c : 0), mycnt, mylast, nch_alloc, s_ctab, Q, group, lane);
                }
            } else
                piece_push<MERGE>(want, PIECE(pstart, cnt), mycnt, mylast, nch_alloc, s_ctab, Q, group, lane);
        }
        // ---- lane-parallel: push the opened internal nodes, in slot order
        {
            const unsigned pushers = __ballot_sync(0xffffffffu, myopeners != 0);
            const int np_ = __popc(pushers);
            if(sp + np_ > WALK_STACK) { if(lane == 0) atomicOr(Q.ctl + 1, 4); break; }      // cannot happen (WALK_RESERVE); reported, not ignored
            if(myopeners) {
                const int w = sp + __popc(pushers & ((1u << lane) - 1u));
                s_stk_node[w] = mynode; s_stk_mask[w] = myopeners;
            }
            sp += np_;
        }
        __syncwarp();
    }
    // hand over to k_grav_pairs (target-slot order: coalesced)
    piece_finish<MERGE>(valid, tslot, mycnt, mylast, nch_alloc, s_ctab, Q, lane);
    if(valid) {
        partial[tslot] = make_double4(ax, ay, az, pot);
        if(COUNT) counts_out[me] = make_int4(n_acc, n_open, n_disc, n_part);
    }
}

// Pair sums over the piece lists written by k_grav_walk, then grav_short_postprocess
// (gravshort.h:47-67).
__global__ void __launch_bounds__(PAIR_WARPS * 32, PAIR_MINB)
k_grav_pairs(const double2 *__restrict__ spart_xy, const double2 *__restrict__ spart_zm, const int *__restrict__ targets,
             const double *__restrict__ pos, const float *__restrict__ mass, const float *__restrict__ gtab,
             WalkPar P, int full_tree, double cbrtrho0,
             const unsigned *__restrict__ pool, const int *__restrict__ chunk_tab, int maxch, const int *__restrict__ piece_cnt,
             const double4 *__restrict__ partial, double *__restrict__ acc_out, double *__restrict__ pot_out)
{
    extern __shared__ float4 s_tabx8[];                 // [B200_SR_NTAB][8] (64 KB), then int [PAIR_WARPS][maxch]
    for(int k = threadIdx.x; k < B200_SR_NTAB * 8; k += blockDim.x) {
        const int t = k >> 3, t1 = t + 1 < B200_SR_NTAB ? t + 1 : t;
        // row NTAB-1 all zero: pairs beyond the table (gravity.c:60-61) contribute nothing
        s_tabx8[k] = t + 1 < B200_SR_NTAB ? make_float4(gtab[t], gtab[t1], gtab[B200_SR_NTAB + t], gtab[B200_SR_NTAB + t1])
                                          : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    int *s_ctab = (int *) (s_tabx8 + B200_SR_NTAB * 8) + (threadIdx.x >> 5) * maxch;
    const int lane = threadIdx.x & 31;
    const int group = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int tslot = group * 32 + lane;
    const bool valid = tslot < P.ntargets;
    if(group * 32 >= P.ntargets) return;    // warp-uniform
    int me = -1, mycnt = 0;
    double px = 0, py = 0, pz = 0;
    double4 acc = make_double4(0, 0, 0, 0);
    if(valid) {
        me = targets[tslot];
        px = pos[3 * (int64_t) me]; py = pos[3 * (int64_t) me + 1]; pz = pos[3 * (int64_t) me + 2];
        mycnt = piece_cnt[tslot];
        acc = partial[tslot];
    }
    piece_load_ctab(s_ctab, chunk_tab, maxch, group, mycnt, lane);

    PieceList L;
    L.pool = pool; L.ctab = s_ctab; L.empty = PIECE(P.sentinel, 0);
    const int g = lane >> 3, slot = lane & 7;
    const TabF4x8 tab{s_tabx8 + slot};
    const SrcArrays spart{spart_xy, spart_zm};
    for(int t = 0; t < 32; t++) {
        const int nt = __shfl_sync(0xffffffffu, mycnt, t);
        if(nt == 0) continue;                           // warp-uniform
        const double tx = __shfl_sync(0xffffffffu, px, t);
        const double ty = __shfl_sync(0xffffffffu, py, t);
        const double tz = __shfl_sync(0xffffffffu, pz, t);
        L.t = t; L.nt = nt;
        double sx = 0, sy = 0, sz = 0, sp = 0;
        // No periodic wrap is needed when the target is farther than the reach of
        // the window table from every face: a pair that NEAREST would wrap is then
        // beyond the table on both sides of the wrap and contributes nothing.
        const bool central = tx >= P.wrap_lo && tx <= P.wrap_hi && ty >= P.wrap_lo && ty <= P.wrap_hi &&
                             tz >= P.wrap_lo && tz <= P.wrap_hi;
        if(central) pair_sum<false>(L, g, slot, spart, P, tab, tx, ty, tz, sx, sy, sz, sp);
        else pair_sum<true>(L, g, slot, spart, P, tab, tx, ty, tz, sx, sy, sz, sp);
        sx = warp_sum(sx); sy = warp_sum(sy); sz = warp_sum(sz); sp = warp_sum(sp);
        if(lane == t) { acc.x += sx; acc.y += sy; acc.z += sz; acc.w += sp; }
    }
    if(valid) {
        // grav_short_postprocess gravshort.h:47-67
        if(acc_out) {
            acc_out[3 * (int64_t) me] = acc.x * P.G;
            acc_out[3 * (int64_t) me + 1] = acc.y * P.G;
            acc_out[3 * (int64_t) me + 2] = acc.z * P.G;
        }
        if(pot_out) {
            double p = acc.w;
            if(full_tree) {
                const double m = (double) mass[me];
                p += m / (P.h / 2.8);
                p -= 2.8372975 * pow(m, 2.0 / 3) * cbrtrho0;
                p *= P.G;
            }
            pot_out[me] = p;
        }
    }
}

__global__ void k_iota(int *p, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i < n) p[i] = i;
}

__global__ void k_mark(const int *__restrict__ list, int64_t n, uint8_t *__restrict__ flags)
{
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if(i < n) flags[list[i]] = 1;
}
__global__ void k_gather_flags(const int *__restrict__ sidx, int np, const uint8_t *__restrict__ flags, uint8_t *__restrict__ out)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if(j < np) out[j] = flags[sidx[j]];
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

// Piece-list storage.  The pool is sized from the last walk's use (first call: 6 chunks per
// warp) and grown when a walk reports that it ran out; the walk is then repeated.
void piece_pool_reset(Engine *E) { E->walk_maxch = WALK_MAXCH0; }

int piece_pool_begin(Engine *E, int64_t nwarps, PiecePool *Q)
{
    CK(E->walk_chunktab.ensure((size_t) nwarps * E->walk_maxch));
    Q->maxch = E->walk_maxch;
    CK(E->walk_cnt.ensure((size_t) nwarps * 32));
    CK(E->scratch_i.ensure(256));
    if(E->walk_want == 0) E->walk_want = (size_t) (E->walk_chunks_per_warp * (double) nwarps) + 1024;
    CK(E->walk_pool.ensure(E->walk_want * CH_WORDS));
    E->walk_want = 0;
    const size_t capz = E->walk_pool.cap / CH_WORDS;
    Q->pool = E->walk_pool.p;
    Q->cap = (int) (capz < (size_t) 0x7fffffff ? capz : (size_t) 0x7fffffff);
    Q->ctl = E->scratch_i.p + 64;
    Q->chunk_tab = E->walk_chunktab.p;
    Q->piece_cnt = E->walk_cnt.p;
    CK(cudaMemsetAsync(Q->ctl, 0, 4 * sizeof(int), E->stream));
    return 0;
}

int piece_pool_check(Engine *E, int64_t nwarps, bool *retry, int attempt)
{
    int h[4] = {0, 0, 0, 0};
    CK(cudaMemcpyAsync(h, E->scratch_i.p + 64, 4 * sizeof(int), cudaMemcpyDeviceToHost, E->stream));
    CK(cudaStreamSynchronize(E->stream));
    *retry = false;
    if(h[1] & 4) return failmsg(E, "tree walk: node stack overflow (tree deeper than the walk kernels allow)");
    if(h[1] & 2) {      // a list outgrew the warp's chunk table: repeat with a table 8x as long
        if(E->walk_maxch >= WALK_MAXCH_LIMIT)
            return failmsg(E, "tree walk: a particle opened more than " + std::to_string(CH_SLOTS * WALK_MAXCH_LIMIT) + " leaf pieces");
        E->walk_maxch *= 8;
        E->walk_want = E->walk_pool.cap / CH_WORDS;
        *retry = true;
        return 0;
    }
    unsigned long long pieces; memcpy(&pieces, h + 2, sizeof(pieces));
    E->walk_pieces = (double) pieces; E->walk_chunks = h[0];
    E->walk_chunks_per_warp = 1.15 * (double) h[0] / (double) (nwarps > 0 ? nwarps : 1) + 0.05;
    const size_t capz = E->walk_pool.cap / CH_WORDS;
    if((size_t) h[0] <= capz) return 0;
    if(attempt >= 2) return failmsg(E, "tree walk: piece pool kept overflowing");
    E->walk_want = (size_t) h[0] + (size_t) h[0] / 8 + 1024;
    *retry = true;
    return 0;
}

// Split the tree's particles (curve order, E->sidx) into `nchunks` groups of equal
// original-index ranges, keeping the curve order inside each group: a stable
// 1-pass radix sort on the chunk number.  The e2e path walks one group at a time
// so that the finished index range can travel back to the host while the next
// group is walked.  offsets[nchunks+1] (host) delimit the groups in E->targets_sorted.
__global__ void k_chunk_keys(const int *__restrict__ sidx, int np, int chunk, unsigned char *__restrict__ key)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if(j < np) key[j] = (unsigned char) (sidx[j] / chunk);
}
__global__ void k_chunk_offsets(const unsigned char *__restrict__ key, int np, int nchunks, int *__restrict__ off)
{
    const int c = threadIdx.x;
    if(c > nchunks) return;
    int lo = 0, hi = np;            // first position with key >= c
    while(lo < hi) { const int mid = (lo + hi) >> 1; if((int) key[mid] < c) lo = mid + 1; else hi = mid; }
    off[c] = lo;
}

int walk_chunk_targets(Engine *E, int nchunks, int64_t chunk, int *offsets)
{
    const int np = (int) E->tree_np;
    if(nchunks < 1 || nchunks > 64) return failmsg(E, "walk_chunk_targets: bad chunk count");
    CK(E->targets_sorted.ensure((size_t) np + 1));
    CK(E->walk_flags.ensure(2 * (size_t) np + 64));
    CK(E->scratch_i.ensure(256));
    unsigned char *k0 = E->walk_flags.p, *k1 = k0 + np;
    if(np > 0) {
        k_chunk_keys<<<(np + 255) / 256, 256, 0, E->stream>>>(E->sidx.p, np, (int) chunk, k0); CKL(E);
        int bits = 1; while((1 << bits) < nchunks) bits++;
        size_t tb = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tb, k0, k1, E->sidx.p, E->targets_sorted.p, np, 0, bits, E->stream);
        CK(E->cubtemp.ensure(tb + 16));
        CK(cub::DeviceRadixSort::SortPairs(E->cubtemp.p, tb, k0, k1, E->sidx.p, E->targets_sorted.p, np, 0, bits, E->stream));
        E->launches += 2;
    }
    k_chunk_offsets<<<1, 128, 0, E->stream>>>(k1, np, nchunks, E->scratch_i.p + 32); CKL(E);
    CK(cudaMemcpyAsync(offsets, E->scratch_i.p + 32, (nchunks + 1) * sizeof(int), cudaMemcpyDeviceToHost, E->stream));
    CK(cudaStreamSynchronize(E->stream));
    return 0;
}

int grav_short_tree(Engine *E, const b200_gravshort_params *par, const int32_t *d_active,
                    int64_t nactive, double *d_acc, double *d_pot, b200_walk_counts *d_counts, bool presorted)
{
    if(!E->tree_valid) return failmsg(E, "b200_grav_short_tree: tree moments not computed (call b200_tree_build)");   // gravshort-tree.c:113-114
    if(E->NmeshWalk == 0) return failmsg(E, "b200_grav_short_tree: call b200_pm_init first (needs Nmesh, Asmth, G)");
    if(!E->srtab.p) if(int rc = walk_init_tables(E)) return rc;
    WalkPar P;
    P.box = E->tree_box; P.halfbox = 0.5 * E->tree_box;
    const double cellsize = E->tree_box / E->NmeshWalk;                 // gravshort-tree.c:101
    P.rcut = par->Rcut * E->Asmth * cellsize;                       // :102
    P.rcut2 = P.rcut * P.rcut;
    P.usebh = par->TreeUseBH;
    P.theta2 = par->BHOpeningAngle * par->BHOpeningAngle;           // :266-270
    if(P.usebh == 0) P.theta2 = par->MaxBHOpeningAngle * par->MaxBHOpeningAngle;
    P.ErrTol = par->ErrTolForceAcc; P.G = E->G;
    P.h = 2.8 * par->GravitySoftening;                              // FORCE_SOFTENING :37-41
    P.h2 = P.h * P.h; P.hinv = 1.0 / P.h; P.h3inv = 1.0 / P.h / P.h / P.h;
    P.h2s = P.h2 > 2.3e-308 ? P.h2 : 2.3e-308;
    P.sentinel = (int) E->tree_np;
    P.inv_cell_dx = 1.0 / (cellsize * (double) B200_SR_DX);
    {   // fp32 pre-classification: r2^2 * aold and mass*len^2 must stay far from the ends of the float range
        static const bool off = getenv("B200_WALK_F64") != nullptr;
        P.f32ok = (!off && E->tree_box > 1e-6 && E->tree_box < 1e7) ? 1 : 0;
        P.rcut2f = (float) P.rcut2; P.th2lo = (float) P.theta2 * (1.f - 2e-6f); P.th2hi = (float) P.theta2 * (1.f + 2e-6f);
    }
    {   // reach of the window table: index >= NTAB-1 <=> r >= (NTAB-1)*dx cells = 15 cells (gravity.c:57-61)
        const double reach = 1.001 * (B200_SR_NTAB - 1) * (double) B200_SR_DX * cellsize;
        P.wrap_lo = reach; P.wrap_hi = E->tree_box - reach;
        if(!(reach < 0.5 * E->tree_box)) { P.wrap_lo = 1.0; P.wrap_hi = -1.0; }
    }
    const double cbrtrho0 = pow(par->rho0, 1.0 / 3);

    // Targets: the tree's own particles in curve order when the walk set is the
    // tree set (the usual case); otherwise the caller's list as given.
    const int *tg = nullptr;
    int64_t nt = 0;
    if(presorted) { tg = d_active; nt = nactive; }
    else if(d_active == nullptr && E->tree_full) { tg = E->sidx.p; nt = E->tree_np; }
    else if(d_active == nullptr) {
        CK(E->targets.ensure(E->n > 0 ? E->n : 1));
        if(E->n > 0) { k_iota<<<(unsigned) ((E->n + 255) / 256), 256, 0, E->stream>>>(E->targets.p, (int) E->n); CKL(E); }
        tg = E->targets.p; nt = E->n;
    } else {
        // Walk the caller's targets in the tree's curve order (a warp then holds 32
        // neighbouring targets); falls back to the given order if some target is not
        // in the tree.
        tg = d_active; nt = nactive;
        if(E->tree_np > 0 && nactive > 0) {
            const size_t np = (size_t) E->tree_np;
            CK(E->walk_flags.ensure((size_t) E->n + np + 64));
            uint8_t *fl = E->walk_flags.p, *fs = fl + E->n;
            CK(cudaMemsetAsync(fl, 0, (size_t) E->n, E->stream));
            k_mark<<<(unsigned) ((nactive + 255) / 256), 256, 0, E->stream>>>(d_active, nactive, fl); CKL(E);
            k_gather_flags<<<(unsigned) ((np + 255) / 256), 256, 0, E->stream>>>(E->sidx.p, (int) np, fl, fs); CKL(E);
            CK(E->targets_sorted.ensure(np + 1));
            CK(E->scratch_i.ensure(256));
            size_t tb = 0;
            cub::DeviceSelect::Flagged(nullptr, tb, E->sidx.p, fs, E->targets_sorted.p, E->scratch_i.p + 8, (int) np, E->stream);
            CK(E->cubtemp.ensure(tb + 16));
            CK(cub::DeviceSelect::Flagged(E->cubtemp.p, tb, E->sidx.p, fs, E->targets_sorted.p, E->scratch_i.p + 8, (int) np, E->stream));
            E->launches += 2;
            int got = 0;
            CK(cudaMemcpyAsync(&got, E->scratch_i.p + 8, sizeof(int), cudaMemcpyDeviceToHost, E->stream));
            CK(cudaStreamSynchronize(E->stream));
            if(got == nactive) tg = E->targets_sorted.p;
        }
    }
    P.ntargets = (int) nt;
    if(nt == 0) return 0;

    if(E->tree_np + B200_SPART_PAD >= (1 << 28))
        return failmsg(E, "b200_grav_short_tree: more than 2^28 particles in one tree (piece entries hold a 28-bit particle offset)");
    const int bs = 128;
    const int64_t nwarps = (nt + 31) / 32;
    const unsigned nb = (unsigned) ((nwarps * 32 + bs - 1) / bs);
    CK(E->walk_partial.ensure((size_t) nwarps * 32 * 4));
    timer_start(E, T_WALK);
    piece_pool_reset(E);
    for(int attempt = 0;; attempt++) {
        PiecePool Q;
        if(int rc = piece_pool_begin(E, nwarps, &Q)) return rc;
        const size_t wsm = piece_ctab_bytes(E, WALK_WARPS);
        // merging of contiguous leaf pieces pays when the leaves are mostly not full (8 particles)
        const bool merge = (double) E->tree_np < 5.0 * (double) E->tree_nn;
        CK(piece_set_smem(k_grav_walk<true, true>, wsm)); CK(piece_set_smem(k_grav_walk<false, true>, wsm));
        CK(piece_set_smem(k_grav_walk<false, false>, wsm));
#define WALK_ARGS (const double4 *) E->nodeA.p, (const double4 *) E->nodeB.p, (const int4 *) E->nodeC.p, \
                  (const int4 *) E->nodeK.p, (const double4 *) E->spart.p, tg, E->pos.p, E->mass.p, E->oldacc.p, E->srtab.p, \
                  P, Q, (double4 *) E->walk_partial.p
        if(d_counts) k_grav_walk<true, true><<<nb, bs, wsm, E->stream>>>(WALK_ARGS, (int4 *) d_counts);
        else if(merge) k_grav_walk<false, true><<<nb, bs, wsm, E->stream>>>(WALK_ARGS, nullptr);
        else k_grav_walk<false, false><<<nb, bs, wsm, E->stream>>>(WALK_ARGS, nullptr);
#undef WALK_ARGS
        CKL(E);
        bool retry = false;
        if(int rc = piece_pool_check(E, nwarps, &retry, attempt)) return rc;
        if(!retry) break;
    }
    timer_stop(E, T_WALK);
    timer_start(E, T_WALK_POST);
    const size_t pair_smem = (size_t) B200_SR_NTAB * 8 * sizeof(float4) + piece_ctab_bytes(E, PAIR_WARPS);
    CK(piece_set_smem(k_grav_pairs, pair_smem));
    const unsigned nbp = (unsigned) ((nwarps + PAIR_WARPS - 1) / PAIR_WARPS);
    k_grav_pairs<<<nbp, PAIR_WARPS * 32, pair_smem, E->stream>>>((const double2 *) E->spart_xy.p, (const double2 *) E->spart_zm.p, tg, E->pos.p, E->mass.p, E->srtab.p, P, E->tree_full ? 1 : 0, cbrtrho0,
                                           E->walk_pool.p, E->walk_chunktab.p, E->walk_maxch, E->walk_cnt.p, (const double4 *) E->walk_partial.p, d_acc, d_pot);
    CKL(E);
    timer_stop(E, T_WALK_POST);
    return 0;
}


} // namespace b200
