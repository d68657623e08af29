// sph.cu -- SPH density (with the smoothing-length iteration) and hydro force on
// the device octree, for a synchronised step (all gas on one time bin).
//
// Replaces density() (libgadget/density.c:234-355) = treewalk_do_hsml_loop
// (treewalk.c:1269-1367) over treewalk_visit_nolist_ngbiter (:1152-1265) with
// density_ngbiter / density_postprocess / density_check_neighbours
// (density.c:424-689), and hydro_force() (hydra.c:153-245) =
// treewalk_visit_ngbiter (treewalk.c:930-1007) + ngb_treefind_threads
// (:1056-1143) with hydro_ngbiter / hydro_postprocess (hydra.c:318-528).
//
// Both passes use the machinery of the tree-gravity kernels (tree_walk.cu, piece_list.cuh):
//
//  k_sph_walk: one warp searches the tree for 32 gas particles adjacent on the curve.  Every
//    lane applies the reference's own node test (cull_node, treewalk.c:1015-1042) with ITS
//    smoothing length; a node is entered when any lane keeps it, and every lane that keeps a
//    particle leaf appends the leaf piece to its own list in the global chunk pool.  The set
//    of leaves a particle sees is therefore exactly that of the reference visitor
//    (treewalk_visit_nolist_ngbiter :1152-1265 / ngb_treefind_threads :1056-1143).
//  k_sph_density_pairs / k_sph_hydro_pairs: for each target the lanes of the warp are spread
//    over the SOURCE particles of its pieces (4 pieces x 8 slots per step), evaluate
//    density_ngbiter (density.c:424-519) / hydro_ngbiter (hydra.c:318-506) and combine the
//    partial sums with warp reductions; the target's lane then runs the post-processing
//    (density_postprocess / density_check_neighbours, hydro_postprocess).
//
// The reference re-queues unconverged particles and re-runs the walk (treewalk_do_hsml_loop,
// treewalk.c:1269-1367); so does sph_density(): each pass walks the still-unconverged
// particles (compacted, still in curve order) with their updated smoothing lengths.
// Sums are accumulated in a different order than on the CPU (lane-parallel, then a butterfly):
// the integer neighbour counts are identical, the floating-point sums agree to rounding.
// This file is compiled with -fmad=false: the per-term arithmetic is the CPU's.
#include "piece_list.cuh"
#include <math.h>
#include <stdlib.h>
#include <cub/device/device_select.cuh>

namespace b200 {

#define GAMMA (5.0 / 3.0)
#define GAMMA_MINUS1 (GAMMA - 1)
#define NORM_COEFF 4.188790204786
#define FACT1 0.366025403785      // treewalk.c:19
#define SPH_MAXITER 400
// per-warp stack of kept internal nodes in k_sph_walk: (node, mask) entries; above STACK - RESERVE the batches shrink to
// one node's children (net growth <= 7 per tree level, depth <= 22; a batch pushes <= 32)
#define SPH_WALK_STACK 320
#define SPH_WALK_RESERVE 192

struct SphDev {
    b200_sph_params p;
    double box, halfbox, desnumngb;
    double fac_mu, fac_vsic_fix, hubble_a2;
    int ktype;          // 0 cubic, 1 quintic, 2 quartic
    double support, sigma;
    // per-time-bin factors [5][B200_TIMEBINS + 1]: gravkick, hydrokick, dloga_pred, drift, dloga_bin
    // (kick_factor_data density.c:114-132, SPH_EntVarPred :74, drifts[] hydra.c:178-186,
    // get_dloga_for_bin hydra.c:271,463) and the particles' bins (TimeBinGravity, TimeBinHydro)
    const double *bins;
    const uint8_t *bg, *bh;
};
#define NB (B200_TIMEBINS + 1)

__device__ __forceinline__ double nearest_s(double x, double box, double halfbox)
{
    return (x > halfbox) ? (x - box) : ((x < -halfbox) ? (x + box) : x);
}

struct Kern { double H, HH, Hinv, Wknorm, dWknorm; };

__device__ __forceinline__ double pw2(double x) { return x * x; }
__device__ __forceinline__ double pw3(double x) { return x * x * x; }
__device__ __forceinline__ double pw4(double x) { const double y = x * x; return y * y; }
__device__ __forceinline__ double pw5(double x) { const double y = x * x; return y * y * x; }

__device__ __forceinline__ void kern_init(Kern &k, double H, const SphDev &S)     // densitykernel.c:140-155
{
    k.H = H; k.HH = H * H; k.Hinv = 1. / H;
    const double hinv = k.Hinv * S.support;
    k.Wknorm = S.sigma * pw3(hinv);
    k.dWknorm = k.Wknorm * hinv;
}
__device__ __forceinline__ double kern_w(const Kern &k, double u, const SphDev &S)   // densitykernel.c:24-90
{
    const double q = u * S.support;
    double w = 0;
    if(S.ktype == 0) {
        if(q < 1.0) w = 0.25 * pw3(2 - q) - pw3(1 - q);
        else if(q < 2.0) w = 0.25 * pw3(2 - q);
    } else if(S.ktype == 1) {
        if(q < 1.0) w = pw5(3 - q) - 6 * pw5(2 - q) + 15 * pw5(1 - q);
        else if(q < 2.0) w = pw5(3 - q) - 6 * pw5(2 - q);
        else if(q < 3.0) w = pw5(3 - q);
    } else {
        if(q < 0.5) w = pw4(2.5 - q) - 5 * pw4(1.5 - q) + 10 * pw4(0.5 - q);
        else if(q < 1.5) w = pw4(2.5 - q) - 5 * pw4(1.5 - q);
        else if(q < 2.5) w = pw4(2.5 - q);
    }
    return k.Wknorm * w;
}
__device__ __forceinline__ double kern_dw(const Kern &k, double u, const SphDev &S)
{
    const double q = u * S.support;
    double w = 0;
    if(S.ktype == 0) {
        if(q < 1.0) w = -0.25 * 3 * pw2(2 - q) + 3 * pw2(1 - q);
        else if(q < 2.0) w = -0.25 * 3 * pw2(2 - q);
    } else if(S.ktype == 1) {
        if(q < 1.0) w = -5 * pw4(3 - q) + 30 * pw4(2 - q) - 75 * pw4(1 - q);
        else if(q < 2.0) w = -5 * pw4(3 - q) + 30 * pw4(2 - q);
        else if(q < 3.0) w = -5 * pw4(3 - q);
    } else {
        if(q < 0.5) w = -4 * pw3(2.5 - q) + 20 * pw3(1.5 - q) - 40 * pw3(0.5 - q);
        else if(q < 1.5) w = -4 * pw3(2.5 - q) + 20 * pw3(1.5 - q);
        else if(q < 2.5) w = -4 * pw3(2.5 - q);
    }
    return k.dWknorm * w;
}

// SPH_VelPred (density.c:91-100) and SPH_EntVarPred (density.c:69-85) for every particle
__global__ void __launch_bounds__(256)
k_sph_predict(int64_t n, const double *__restrict__ vel, const double *__restrict__ fullacc, const double *__restrict__ gravpm,
              const double *__restrict__ hydroacc, const double *__restrict__ entropy, const double *__restrict__ dtentropy,
              SphDev S, double *__restrict__ velpred, double *__restrict__ evp)
{
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    const double gk = S.bins[S.bg[i]], hk = S.bins[NB + S.bh[i]], dlp = S.bins[2 * NB + S.bh[i]];
    for(int j = 0; j < 3; j++)
        velpred[3 * i + j] = (vel ? vel[3 * i + j] : 0) + gk * (fullacc ? fullacc[3 * i + j] : 0)
                           + (gravpm ? gravpm[3 * i + j] : 0) * S.p.pmkick + hk * (hydroacc ? hydroacc[3 * i + j] : 0);
    const double E = entropy ? entropy[i] : 1.0;
    double e = E + (dtentropy ? dtentropy[i] : 0.0) * dlp;
    if(e < 0.05 * E) e = 0.05 * E;
    evp[i] = e <= 0 ? 0 : exp(1. / GAMMA * log(e));
}

__global__ void __launch_bounds__(256)
k_sph_gather_vel(int np, const int *__restrict__ sidx, const double *__restrict__ velpred, const double *__restrict__ evp,
                 double4 *__restrict__ svel)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if(j >= np) return;
    const int64_t i = sidx[j];
    svel[j] = make_double4(velpred[3 * i], velpred[3 * i + 1], velpred[3 * i + 2], evp[i]);
}

// cull_node treewalk.c:1015-1042
__device__ __forceinline__ bool cull_keep(const double4 B, double hmaxnode, double px, double py, double pz, double hsml,
                                          bool symmetric, const SphDev &S)
{
    double dist = (symmetric ? (hmaxnode > hsml ? hmaxnode : hsml) : hsml) + 0.5 * B.w;
    double r2 = 0;
    double dx = nearest_s(B.x - px, S.box, S.halfbox);
    if(dx > dist || dx < -dist) return false;
    r2 += dx * dx;
    dx = nearest_s(B.y - py, S.box, S.halfbox);
    if(dx > dist || dx < -dist) return false;
    r2 += dx * dx;
    dx = nearest_s(B.z - pz, S.box, S.halfbox);
    if(dx > dist || dx < -dist) return false;
    r2 += dx * dx;
    dist += FACT1 * B.w;
    return !(r2 > dist * dist);
}

// Staged node rows of the current batch (one entry per lane).
struct SphBatch {
    double4 B[32];    // center, len
    int4 M[32];       // pstart, count, mask of lanes that kept every ancestor, flags (bit0 leaf, bit1 culled for all lanes)
    double H[32];     // hmax of the node (symmetric search)
    int N[32];        // node index (batch input)
};

// Neighbour search for 32 targets.  targets: sorted (curve-order) particle indices, NULL = 0..nt-1.
// SYM: symmetric search max(h_i, hmax(node)) of the hydro pass (NGB_TREEFIND_SYMMETRIC).
template <bool SYM>
__global__ void __launch_bounds__(128, 5)
k_sph_walk(const double4 *__restrict__ nodeB, const int4 *__restrict__ nodeC, const int4 *__restrict__ nodeK,
           const double *__restrict__ nodeH, const double4 *__restrict__ spart, const int *__restrict__ sidx,
           const int *__restrict__ targets, int nt, int tpw, const double *__restrict__ hsml, SphDev S, PiecePool Q,
           double *__restrict__ reach)     // [target slot] bound on the distance to any candidate of the kept leaves
{
    // tpw = targets per warp (lanes 0..tpw-1): 32 for a pass over (nearly) all particles; the later passes of the
    // smoothing-length iteration are sparse -- 32 consecutive unconverged targets lie far apart and the union of
    // their searches approaches the whole tree -- and use 8 or 1.
    __shared__ double4 s_tgt_all[WALK_WARPS][32];       // position, search radius of the warp's targets
    extern __shared__ int s_ctab_dyn[];                 // [WALK_WARPS][Q.maxch]
    __shared__ int s_stk_node_all[WALK_WARPS][SPH_WALK_STACK];
    __shared__ unsigned s_stk_mask_all[WALK_WARPS][SPH_WALK_STACK];
    __shared__ SphBatch s_ent_all[WALK_WARPS];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int *s_ctab = s_ctab_dyn + wib * Q.maxch;
    int *s_stk_node = s_stk_node_all[wib];
    unsigned *s_stk_mask = s_stk_mask_all[wib];
    SphBatch &s_ent = s_ent_all[wib];
    double4 *s_tgt = s_tgt_all[wib];
    const int group = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int tslot = group * tpw + lane;
    const bool valid = lane < tpw && tslot < nt;
    if((int64_t) group * tpw >= nt) return; // warp-uniform
    double px = 0, py = 0, pz = 0, h = 0;
    if(valid) {
        const int j = targets ? targets[tslot] : tslot;
        const double4 pm = spart[j];
        px = pm.x; py = pm.y; pz = pm.z;
        h = hsml[sidx[j]];
    }
    s_tgt[lane] = make_double4(px, py, pz, h);
    const double big = 1e300;
    const double lox = warp_min(valid ? px : big), hix = warp_max(valid ? px : -big);
    const double loy = warp_min(valid ? py : big), hiy = warp_max(valid ? py : -big);
    const double loz = warp_min(valid ? pz : big), hiz = warp_max(valid ? pz : -big);
    const double bcx = 0.5 * (lox + hix), bcy = 0.5 * (loy + hiy), bcz = 0.5 * (loz + hiz);
    const double bhx = 0.5 * (hix - lox), bhy = 0.5 * (hiy - loy), bhz = 0.5 * (hiz - loz);
    const double hw = warp_max(h);
    const unsigned validmask = __ballot_sync(0xffffffffu, valid);
    int mycnt = 0, nch_alloc = 0;
    unsigned mylast = 0;
    double myreach = 0;

    // The stack holds KEPT INTERNAL NODES (node, mask of the lanes that kept it); a batch is formed by expanding the
    // topmost entries into up to 32 children (see k_grav_walk, tree_walk.cu).
    int sp = 0, nb = 1;
    if(lane == 0) { s_ent.N[0] = 0; s_ent.M[0].z = (int) validmask; }
    __syncwarp();
    for(bool first = true;; first = false) {
        if(!first) {
            if(sp == 0) break;
            int myc = 0;
            unsigned pmask = 0;
            int4 k0 = make_int4(-1, -1, -1, -1), k1 = k0;
            if(lane < 16 && lane < sp) {
                const int pnode = s_stk_node[sp - 1 - lane]; pmask = s_stk_mask[sp - 1 - lane];
                k0 = nodeK[2 * (size_t) pnode]; k1 = nodeK[2 * (size_t) pnode + 1];
                myc = (k0.x >= 0) + (k0.y >= 0) + (k0.z >= 0) + (k0.w >= 0) + (k1.x >= 0) + (k1.y >= 0) + (k1.z >= 0) + (k1.w >= 0);
            }
            int Sc = myc;           // inclusive scan from the top of the stack downwards
#pragma unroll
            for(int o = 1; o < 16; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, Sc, o); if(lane >= o) Sc += v; }
            const int cap = sp > SPH_WALK_STACK - SPH_WALK_RESERVE ? 8 : 32;
            const int J = __popc(__ballot_sync(0xffffffffu, lane < 16 && lane < sp && Sc <= cap));
            nb = __shfl_sync(0xffffffffu, Sc, J - 1);
            if(lane < J) {
                int w = nb - Sc;
                const int kids[8] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
#pragma unroll
                for(int c = 0; c < 8; c++) if(kids[c] >= 0) { s_ent.N[w] = kids[c]; s_ent.M[w].z = (int) pmask; w++; }
            }
            sp -= J;
            __syncwarp();
        }
        // ---- lane-parallel: lane l fetches batch entry l and tests it against the warp's bounding box
        int mynode = -1;
        bool dead = true, bigleaf = false, isleaf = false;
        unsigned emask0 = 0;
        double4 eB = make_double4(0, 0, 0, 0);
        double eH = 0;
        if(lane < nb) {
            mynode = s_ent.N[lane];
            emask0 = (unsigned) s_ent.M[lane].z;
            eB = nodeB[mynode];
            const int4 C = nodeC[mynode];
            eH = SYM ? nodeH[mynode] : 0.0;
            isleaf = C.w != 0;
            bigleaf = isleaf && C.z > 8;
            // cull_node's per-axis test (treewalk.c:1023-1036) for the whole warp: the largest search
            // radius of the warp against the distance from the node centre to the bounding box
            const double dist = ((SYM && eH > hw) ? eH : hw) + 0.5 * eB.w;
            const double lim = dist + 1e-9 * (dist + eB.w);
            const double ex = nearest_s(eB.x - bcx, S.box, S.halfbox), ey = nearest_s(eB.y - bcy, S.box, S.halfbox),
                         ez = nearest_s(eB.z - bcz, S.box, S.halfbox);
            dead = fabs(ex) - bhx > lim || fabs(ey) - bhy > lim || fabs(ez) - bhz > lim;
            s_ent.B[lane] = eB;
            s_ent.H[lane] = eH;
            s_ent.M[lane] = make_int4(C.y, C.z, (int) emask0, isleaf ? 1 : 0);
        }
        __syncwarp();
        // ---- exact decisions, one TARGET per turn: every lane tests the node it holds (in registers) against the
        // target's position and radius (broadcast from shared memory), so a turn is 32 (node, target) tests whatever
        // the number of targets in the warp.  The vote gives the target its kept leaves; a lane collects who kept its node.
        unsigned myopeners = 0, keepbits = 0;
        const unsigned leafmask = __ballot_sync(0xffffffffu, isleaf);
        const unsigned anybig = __ballot_sync(0xffffffffu, bigleaf);
        for(unsigned m = __reduce_or_sync(0xffffffffu, dead ? 0u : emask0); m; m &= m - 1) {
            const int t = __ffs(m) - 1;
            const double4 T = s_tgt[t];
            const bool keep = !dead && ((emask0 >> t) & 1u) && cull_keep(eB, eH, T.x, T.y, T.z, T.w, SYM, S);
            const unsigned bal = __ballot_sync(0xffffffffu, keep);
            if(lane == t) keepbits = bal & leafmask;
            if(keep && !isleaf) myopeners |= 1u << t;
        }
        // ---- kept leaves: every lane appends its own pieces in slot order; chunks reserved by one vote
        if(anybig == 0) {
            piece_reserve(__popc(keepbits), mycnt, nch_alloc, s_ctab, Q, group, lane);
            while(keepbits) {
                const int k = __ffs(keepbits) - 1; keepbits &= keepbits - 1;
                const int4 M = s_ent.M[k];
                // every particle of a kept leaf is within search distance + len of the target (cull_node)
                const double hn = s_ent.H[k], len = s_ent.B[k].w;
                const double r = ((SYM && hn > h) ? hn : h) + len;
                myreach = r > myreach ? r : myreach;
                piece_append<true>(PIECE(M.x, M.y), mycnt, mylast, nch_alloc, s_ctab, Q, lane);
            }
            __syncwarp();
        } else
        while(__any_sync(0xffffffffu, keepbits != 0)) {         // a leaf at the key-depth limit holds more than 8 particles
            const bool want = keepbits != 0;
            int4 M = make_int4(0, 0, 0, 0);
            if(want) {
                const int k = __ffs(keepbits) - 1; keepbits &= keepbits - 1;
                M = s_ent.M[k];
                const double hn = s_ent.H[k], len = s_ent.B[k].w;
                const double r = ((SYM && hn > h) ? hn : h) + len;
                myreach = r > myreach ? r : myreach;
            }
            const int maxc = (int) __reduce_max_sync(0xffffffffu, (unsigned) (want ? M.y : 0));
            for(int o = 0; o < maxc; o += 8) {
                const int c = M.y - o < 8 ? M.y - o : 8;
                piece_push<true>(want && o < M.y, PIECE(M.x + o, c > 0 ? c : 0), mycnt, mylast, nch_alloc, s_ctab, Q, group, lane);
            }
        }
        // ---- lane-parallel: push the kept internal nodes, in slot order
        {
            const unsigned pushers = __ballot_sync(0xffffffffu, myopeners != 0);
            const int np_ = __popc(pushers);
            if(sp + np_ > SPH_WALK_STACK) { if(lane == 0) atomicOr(Q.ctl + 1, 4); break; }
            if(myopeners) {
                const int w = sp + __popc(pushers & ((1u << lane) - 1u));
                s_stk_node[w] = mynode; s_stk_mask[w] = myopeners;
            }
            sp += np_;
        }
        __syncwarp();
    }
    piece_finish<true>(valid, tslot, mycnt, mylast, nch_alloc, s_ctab, Q, lane);
    if(valid) reach[tslot] = myreach;
}

#define NSUM_DENS 12

// density_ngbiter over the piece lists + density_postprocess + density_check_neighbours for the
// targets of this pass.  State (Hsml, Left, Right, niter) is indexed by particle index.
#ifndef SPH_DENS_MINB
#define SPH_DENS_MINB 6
#endif
__global__ void __launch_bounds__(128, SPH_DENS_MINB)
k_sph_density_pairs(const int *__restrict__ targets, int nt, const int *__restrict__ sidx,
                    const double4 *__restrict__ spart, const double2 *__restrict__ spart_xy, const double2 *__restrict__ spart_zm,
                    const double *__restrict__ reach,
                    const double4 *__restrict__ svel, SphDev S, int update_hsml, int DoEgy, int tpw,
                    const unsigned *__restrict__ pool, const int *__restrict__ chunk_tab, int maxch, const int *__restrict__ piece_cnt, int sentinel,
                    double *__restrict__ hsml, double *__restrict__ left, double *__restrict__ right,
                    double *__restrict__ density, double *__restrict__ egy, double *__restrict__ dhsmlfac,
                    double *__restrict__ divvel, double *__restrict__ curlvel, double *__restrict__ dthsml, double *__restrict__ numngb,
                    double *__restrict__ gradrho, int *__restrict__ ninteract, int *__restrict__ niter,
                    uint8_t *__restrict__ notdone, int *__restrict__ err)
{
    extern __shared__ int s_ctab_dyn[];                 // [WALK_WARPS][maxch]
    int *s_ctab = s_ctab_dyn + (threadIdx.x >> 5) * maxch;
    const int lane = threadIdx.x & 31;
    const int group = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int tslot = group * tpw + lane;   // the walk's assignment: tpw targets per warp
    const bool valid = lane < tpw && tslot < nt;
    if((int64_t) group * tpw >= nt) return; // warp-uniform
    int me = -1, mycnt = 0;
    double4 pm = make_double4(0, 0, 0, 0), vm = pm;
    double h = 1, myreach = 0;
    if(valid) {
        const int j = targets ? targets[tslot] : tslot;
        me = sidx[j];
        pm = spart[j]; vm = svel[j];
        h = hsml[me];
        mycnt = piece_cnt[tslot];
        myreach = reach[tslot];
    }
    piece_load_ctab(s_ctab, chunk_tab, maxch, group, mycnt, lane);
    PieceList L;
    L.pool = pool; L.ctab = s_ctab; L.empty = PIECE(sentinel, 0);
    const int g = lane >> 3, slot = lane & 7;
    const unsigned ltmask = (1u << lane) - 1u;
    __shared__ int s_q_all[WALK_WARPS][64];
    int *s_q = s_q_all[threadIdx.x >> 5];
    double acc[NSUM_DENS];
#pragma unroll
    for(int q = 0; q < NSUM_DENS; q++) acc[q] = 0;
    int nint = 0;
    for(int t = 0; t < tpw; t++) {
        const int ntp = __shfl_sync(0xffffffffu, mycnt, t);
        if(ntp == 0) continue;                          // warp-uniform
        const double tx = __shfl_sync(0xffffffffu, pm.x, t), ty = __shfl_sync(0xffffffffu, pm.y, t), tz = __shfl_sync(0xffffffffu, pm.z, t);
        const double tvx = __shfl_sync(0xffffffffu, vm.x, t), tvy = __shfl_sync(0xffffffffu, vm.y, t), tvz = __shfl_sync(0xffffffffu, vm.z, t);
        const double th = __shfl_sync(0xffffffffu, h, t);
        Kern k; kern_init(k, th, S);
        const double vol = NORM_COEFF * pw3(k.H);
        const double h2 = th * th;
        L.t = t; L.nt = ntp;
        double s[NSUM_DENS];
#pragma unroll
        for(int q = 0; q < NSUM_DENS; q++) s[q] = 0;
        int ni = 0;
        // A target farther than its search reach from every face needs no periodic wrap: every
        // candidate of its kept leaves is its own nearest image (the reach bounds their distance).
        const double tr = __shfl_sync(0xffffffffu, myreach, t);
        const bool central = tx >= tr && tx + tr <= S.box && ty >= tr && ty + tr <= S.box && tz >= tr && tz + tr <= S.box;
        // Candidates are screened 32 at a time (r^2 <= h^2, treewalk.c:1223-1233); the neighbours (about 1/4 of them) are
        // queued in shared memory and density_ngbiter runs on full warps, in the same loop body.
        int qn = 0;
        Ent4 eC = fetch_ent(L, 0, g), eN = fetch_ent(L, 16, g);
#pragma unroll 1
        for(int it = 0;; it++) {
            const int base = (it >> 2) << 4, kk = it & 3;
            const bool more = base < ntp;               // warp-uniform
            if(!more && qn == 0) break;
            if(more) {
                const unsigned e = kk == 0 ? eC.e[0] : (kk == 1 ? eC.e[1] : (kk == 2 ? eC.e[2] : eC.e[3]));
                const int o = (int) (e >> 4) + slot;
                bool pass = false;
                if(slot < (int) (e & 15u)) {
                    const double2 qa = spart_xy[o], qb = spart_zm[o];       // whole sectors per 8-lane group
                    double d0 = tx - qa.x, d1 = ty - qa.y, d2 = tz - qb.x;
                    if(!central) { d0 = nearest_s(d0, S.box, S.halfbox); d1 = nearest_s(d1, S.box, S.halfbox); d2 = nearest_s(d2, S.box, S.halfbox); }
                    double r2 = d0 * d0; r2 += d1 * d1; r2 += d2 * d2;
                    pass = !(r2 > h2);
                }
                const unsigned bal = __ballot_sync(0xffffffffu, pass);
                if(pass) s_q[qn + __popc(bal & ltmask)] = o;
                qn += __popc(bal);
                ni += pass ? 1 : 0;
                if(kk == 3) { eC = eN; eN = fetch_ent(L, base + 32, g); }
                __syncwarp();
            }
            if(qn >= 32 || (!more && qn > 0)) {
                if(lane < qn) {
                    const int o = s_q[lane];
                    const double2 qa = spart_xy[o], qb = spart_zm[o];
                    double d0 = tx - qa.x, d1 = ty - qa.y, d2 = tz - qb.x;
                    if(!central) { d0 = nearest_s(d0, S.box, S.halfbox); d1 = nearest_s(d1, S.box, S.halfbox); d2 = nearest_s(d2, S.box, S.halfbox); }
                    double r2 = d0 * d0; r2 += d1 * d1; r2 += d2 * d2;
                    if(r2 < k.HH) {                       // density_ngbiter density.c:451-518
                        const double r = sqrt(r2);
                        const double u = r * k.Hinv;
                        const double wk = kern_w(k, u, S);
                        s[0] += wk * vol;
                        const double dwk = kern_dw(k, u, S);
                        const double mj = qb.y;
                        s[1] += mj * wk;
                        const double dW = -(3 * k.Hinv * wk + u * dwk);
                        s[2] += mj * dW;
                        const double4 vo = svel[o];
                        if(DoEgy) { s[3] += mj * vo.w * wk; s[4] += mj * vo.w * dW; }
                        if(r > 0) {
                            const double fac = mj * dwk / r;
                            const double v0 = tvx - vo.x, v1 = tvy - vo.y, v2 = tvz - vo.z;
                            s[5] += -fac * (d0 * v0 + d1 * v1 + d2 * v2);
                            s[6] += fac * (v1 * d2 - d1 * v2);
                            s[7] += fac * (v2 * d0 - d2 * v0);
                            s[8] += fac * (v0 * d1 - d0 * v1);
                            s[9] += fac * d0; s[10] += fac * d1; s[11] += fac * d2;         // density.c:512-515
                        }
                    }
                }
                const int mv = lane + 32 < qn ? s_q[32 + lane] : 0;
                __syncwarp();
                if(lane + 32 < qn) s_q[lane] = mv;
                qn = qn > 32 ? qn - 32 : 0;
                __syncwarp();
            }
        }
#pragma unroll
        for(int q = 0; q < NSUM_DENS; q++) s[q] = warp_sum(s[q]);
        ni = (int) __reduce_add_sync(0xffffffffu, (unsigned) ni);
        if(lane == t) {
#pragma unroll
            for(int q = 0; q < NSUM_DENS; q++) acc[q] = s[q];
            nint = ni;
        }
    }
    if(!valid) return;
    const double Ngb = acc[0], Rho = acc[1], Dh = acc[2], EgyRho = acc[3], DhEgy = acc[4], Div = acc[5];
    // density_postprocess density.c:532-586
    if(Rho <= 0 && Ngb > 0) atomicAdd(err, 1);
    double DhsmlDens = Dh * h / (3 * Rho);
    DhsmlDens = 1 / (1 + DhsmlDens);
    bool done = true;
    if(update_hsml) {
        // density_check_neighbours density.c:589-689
        double Left = left[me], Right = right[me];
        const double des = S.desnumngb, maxdev = S.p.MaxNumNgbDeviation;
        if(Ngb < (des - maxdev) || Ngb > (des + maxdev)) {
            if((Right - Left) < 1.0e-5 * Left) { h = Right; done = true; }
            else {
                if(Ngb < des) Left = h; else Right = h;
                if((Right < S.box && Left > 0) || (h * 1.26 > 0.99 * S.box))
                    h = cbrt(0.5 * (pw3(Left) + pw3(Right)));
                else {
                    double fac = 1.26;
                    if(Ngb > 0) fac = 1 - (Ngb - des) / (3 * Ngb) * DhsmlDens;
                    if(Right > 0.99 * S.box && Left > 0)
                        if(DhsmlDens <= 0 || fabs(Ngb - des) >= 0.5 * des || fac > 1.26) fac = 1.26;
                    if(Right < 0.99 * S.box && Left == 0)
                        if(DhsmlDens <= 0 || fac < 1. / 3) fac = 1. / 3;
                    h *= fac;
                }
                if(Right < S.p.MinGasHsml) { h = S.p.MinGasHsml; done = true; }
                else done = false;
            }
        } else {
            if(h < S.p.MinGasHsml) h = S.p.MinGasHsml;
            done = true;
        }
        left[me] = Left; right[me] = Right;
        hsml[me] = h;
        const int it = niter[me] + 1;
        niter[me] = it;
        if(!done && it > SPH_MAXITER + 1) { atomicAdd(err, 1); done = true; }
    } else niter[me] = 1;
    notdone[tslot] = done ? 0 : 1;
    if(!done) return;
    density[me] = Rho;
    if(DoEgy) {
        double f = DhEgy * h / (3 * EgyRho);
        f *= -DhsmlDens;
        dhsmlfac[me] = f;
        egy[me] = EgyRho / vm.w;
    } else {
        dhsmlfac[me] = DhsmlDens;
        egy[me] = 0;
    }
    curlvel[me] = sqrt(acc[6] * acc[6] + acc[7] * acc[7] + acc[8] * acc[8]) / Rho;
    const double dv = Div / Rho;
    divvel[me] = dv;
    dthsml[me] = (1.0 / 3) * dv * h;
    numngb[me] = Ngb;
    gradrho[3 * (int64_t) me] = acc[9]; gradrho[3 * (int64_t) me + 1] = acc[10]; gradrho[3 * (int64_t) me + 2] = acc[11];
    ninteract[me] = nint;
}

__global__ void k_sph_iota(int *p, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i < n) p[i] = i;
}

__global__ void k_sph_init_state(int64_t n, double box, double *__restrict__ left, double *__restrict__ right, int *__restrict__ niter,
                                 int *__restrict__ nint)
{
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if(i < n) { left[i] = 0; right[i] = box; niter[i] = 0; nint[i] = 0; }
}

// update_tree_hmax_father forcetree.c:1287-1315 for every leaf, then bottom-up max (forcetree.c:1090-1091)
__global__ void __launch_bounds__(256)
k_sph_hmax_leaf(int nn, const double4 *__restrict__ nodeB, const int4 *__restrict__ nodeC, const double4 *__restrict__ spart,
                const int *__restrict__ sidx, const double *__restrict__ hsml, const uint8_t *__restrict__ type,
                double *__restrict__ nodeH)
{
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if(d >= nn) return;
    const int4 C = nodeC[d];
    double hm = 0;
    if(C.w) {
        const double4 B = nodeB[d];
        for(int c = 0; c < C.z; c++) {
            const int o = sidx[C.y + c];
            if(type[o] != 0) continue;
            const double4 q = spart[C.y + c];
            const double hs = hsml[o];
            hm = fmax(hm, fabs(q.x - B.x) + hs - B.w / 2.);
            hm = fmax(hm, fabs(q.y - B.y) + hs - B.w / 2.);
            hm = fmax(hm, fabs(q.z - B.z) + hs - B.w / 2.);
        }
    }
    nodeH[d] = hm;
}
__global__ void __launch_bounds__(128)
k_sph_hmax_up(int first, int last, const int *__restrict__ b_firstchild, const int *__restrict__ b_nchild,
              const int *__restrict__ b_dfs, double *__restrict__ nodeH)
{
    const int b = first + blockIdx.x * blockDim.x + threadIdx.x;
    if(b >= last) return;
    const int nch = b_nchild[b];
    if(nch == 0) return;
    double hm = 0;
    const int fc = b_firstchild[b];
    for(int k = 0; k < nch; k++) hm = fmax(hm, nodeH[b_dfs[fc + k]]);
    nodeH[b_dfs[b]] = hm;
}

// Per-particle hydro inputs in curve order.  Everything hydro_ngbiter (hydra.c:350-505) derives from the OTHER particle
// alone -- kernel normalisation of h_j, P_j / rho_j^2, sound speed, Balsara factor, the density-contrast ratio -- is formed
// here once per particle with the reference's own expressions instead of once per pair (11 of the 14 divisions and 2 of
// the 3 square roots of a pair).
//   hA = {h, 1/h, dWknorm(h), predicted density}      hB = {P/eom^2, c_s, f2 (Balsara), rr2}
//   hC = {DhsmlEgyDensityFactor, dloga of the hydro bin, 1/EntVarPred, eom}     (eom = predicted EgyWtDensity or density)
__global__ void __launch_bounds__(256)
k_sph_gather_hydro(int np, const int *__restrict__ sidx, SphDev S, const double *__restrict__ hsml, const double *__restrict__ density,
                   const double *__restrict__ egy, const double *__restrict__ dhsmlfac, const double *__restrict__ divvel,
                   const double *__restrict__ curlvel, const double4 *__restrict__ svel,
                   double4 *__restrict__ hA, double4 *__restrict__ hB, double4 *__restrict__ hC, double4 *__restrict__ hT)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if(j >= np) return;
    const int64_t i = sidx[j];
    const int DI = S.p.DensityIndependentSphOn;
    const double dens = density[i], dvv = divvel[i], h = hsml[i], curl = curlvel[i];
    const double drift = S.bins[3 * NB + S.bh[i]];
    // SPH_DensityPred hydra.c:300-312
    double dj = dens - dvv * dens * drift; if(!(dj >= 1e-6 * dens)) dj = 1e-6 * dens;
    const double eom0 = DI ? egy[i] : dens;
    double eom = eom0 - dvv * eom0 * drift; if(!(eom >= 1e-6 * eom0)) eom = 1e-6 * eom0;
    const double ev = svel[j].w;
    double P = 0;                                   // PressurePred hydra.c:67-77, cache hydra.c:195-214
    if(ev != 0 && ev * eom > 0) P = exp(GAMMA * log(ev * eom));
    Kern k; kern_init(k, h, S);
    const double cs = sqrt(GAMMA * P / eom);                                           // hydra.c:388
    const double f2 = fabs(dvv) / (fabs(dvv) + curl + 0.0001 * cs / S.fac_mu / h);     // hydra.c:444-445
    double rr2 = 1;                                                                   // hydra.c:486-497
    if(DI) {
        rr2 = 0;
        if(S.p.DensityContrastLimit >= 0) {
            rr2 = eom / dj;
            if(S.p.DensityContrastLimit > 0 && S.p.DensityContrastLimit < rr2) rr2 = S.p.DensityContrastLimit;
        }
    }
    hA[j] = make_double4(h, k.Hinv, k.dWknorm, dj);
    hB[j] = make_double4(P / (eom * eom), cs, f2, rr2);
    hC[j] = make_double4(dhsmlfac[i], S.bins[4 * NB + S.bh[i]], 1.0 / ev, eom);
    // the same particle as a TARGET (hydro_copy hydra.c:247-277 uses the un-predicted density of the slot): {P, eom0, divv, curl}
    hT[j] = make_double4(P, eom0, dvv, curl);
}

// hydro_ngbiter (hydra.c:318-506) over the piece lists of the symmetric search, then
// hydro_postprocess (hydra.c:514-528).  Candidates = particles of the kept leaves
// (treewalk.c:962-999 filters them by r^2 <= max(h_i, h_j)^2).
// Candidates are screened 32 at a time; the survivors (about 1/4 of the particles of the opened leaves) are queued
// in shared memory and evaluated 32 at a time, so the long pair arithmetic runs on full warps.  Screening and
// evaluation share ONE loop body (a second inlined copy of the pair arithmetic doubled the kernel and made it
// instruction-fetch bound: ncu "no instruction" stalls, profiles/r01_sph_hydro_pairs_ncu_summary.txt).
#ifndef SPH_HYDRO_MINB
#define SPH_HYDRO_MINB 5
#endif
__global__ void __launch_bounds__(128, SPH_HYDRO_MINB)
k_sph_hydro_pairs(const int *__restrict__ targets, int nt, const int *__restrict__ sidx,
                  const double4 *__restrict__ spart, const double2 *__restrict__ spart_xy, const double2 *__restrict__ spart_zm,
                  const double *__restrict__ reach, const double4 *__restrict__ svel,
                  const double4 *__restrict__ hA, const double4 *__restrict__ hB, const double4 *__restrict__ hC, const double4 *__restrict__ hT,
                  const double *__restrict__ density, SphDev S,
                  const unsigned *__restrict__ pool, const int *__restrict__ chunk_tab, int maxch, const int *__restrict__ piece_cnt,
                  double *__restrict__ acc_out, double *__restrict__ dte_out, double *__restrict__ maxsig_out, int *__restrict__ ninteract)
{
    extern __shared__ int s_ctab_dyn[];                 // [WALK_WARPS][maxch]
    int *s_ctab = s_ctab_dyn + (threadIdx.x >> 5) * maxch;
    const int lane = threadIdx.x & 31;
    const int group = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int tslot = group * 32 + lane;
    const bool valid = tslot < nt;
    if(group * 32 >= nt) return;            // warp-uniform
    int me = -1, mycnt = 0, myj = 0;
    double myreach = 0;
    if(valid) {
        myj = targets ? targets[tslot] : tslot;
        me = sidx[myj];
        mycnt = piece_cnt[tslot];
        myreach = reach[tslot];
    }
    piece_load_ctab(s_ctab, chunk_tab, maxch, group, mycnt, lane);
    PieceList L;
    L.pool = pool; L.ctab = s_ctab; L.empty = PIECE(0, 0);
    const int g = lane >> 3, slot = lane & 7;
    const unsigned ltmask = (1u << lane) - 1u;
    __shared__ int s_q_all[WALK_WARPS][64];
    int *s_q = s_q_all[threadIdx.x >> 5];
    const int DI = S.p.DensityIndependentSphOn;
    double rA0 = 0, rA1 = 0, rA2 = 0, rDtE = 0, rMaxSig = 0, rdens = 1;
    int rncand = 0;
    for(int t = 0; t < 32; t++) {
        const int ntp = __shfl_sync(0xffffffffu, mycnt, t);
        if(ntp == 0) continue;                          // warp-uniform
        // the target's rows: one broadcast load each
        const int jt = __shfl_sync(0xffffffffu, myj, t), met = __shfl_sync(0xffffffffu, me, t);
        const double4 pm = spart[jt], vm = svel[jt], a_i = hA[jt], c_i = hC[jt], t_i = hT[jt];
        // no periodic wrap for a target farther than its search reach from every face (see k_sph_density_pairs)
        const double tr = __shfl_sync(0xffffffffu, myreach, t);
        const bool wrap = !(pm.x >= tr && pm.x + tr <= S.box && pm.y >= tr && pm.y + tr <= S.box && pm.z >= tr && pm.z + tr <= S.box);
        const double h_i = a_i.x, P_i = t_i.x, eom_i = t_i.y, dens_i = density[met], dloga_i = c_i.y;
        // hydro_copy hydra.c:247-277
        const double cs_i = sqrt(GAMMA * P_i / eom_i);
        const double F1 = fabs(t_i.z) / (fabs(t_i.z) + t_i.w + 0.0001 * cs_i / h_i / S.fac_mu);
        const double p_over_rho2_i = P_i / (eom_i * eom_i);
        const double inv_evp_i = 1.0 / vm.w;
        double rr1 = 1;
        if(DI) {
            rr1 = 0;
            if(S.p.DensityContrastLimit >= 0) {
                rr1 = eom_i / dens_i;
                if(S.p.DensityContrastLimit > 0 && S.p.DensityContrastLimit < rr1) rr1 = S.p.DensityContrastLimit;
            }
        }
        const double pd_i = p_over_rho2_i * c_i.x * rr1;         // hydra.c:499-500, target side
        Kern ki; kern_init(ki, h_i, S);
        double A0 = 0, A1 = 0, A2 = 0, DtE = 0, MaxSig = cs_i;
        int ncand = 0;
        L.t = t; L.nt = ntp;
        int qn = 0;
        Ent4 eC = fetch_ent(L, 0, g), eN = fetch_ent(L, 16, g);
#pragma unroll 1
        for(int it = 0;; it++) {
            const int base = (it >> 2) << 4, kk = it & 3;
            const bool more = base < ntp;               // warp-uniform
            if(!more && qn == 0) break;
            if(more) {
                // ---- screening of one entry per 8-lane group (treewalk.c:962-999)
                const unsigned e = kk == 0 ? eC.e[0] : (kk == 1 ? eC.e[1] : (kk == 2 ? eC.e[2] : eC.e[3]));
                if(slot == 0) ncand += (int) (e & 15u);
                const int o = (int) (e >> 4) + slot;
                bool pass = false;
                if(slot < (int) (e & 15u)) {
                    const double2 qa = spart_xy[o], qb = spart_zm[o];
                    const double hj = hA[o].x;
                    const double hm = hj > h_i ? hj : h_i, h2 = hm * hm;
                    double d0 = pm.x - qa.x, d1 = pm.y - qa.y, d2 = pm.z - qb.x;
                    if(wrap) { d0 = nearest_s(d0, S.box, S.halfbox); d1 = nearest_s(d1, S.box, S.halfbox); d2 = nearest_s(d2, S.box, S.halfbox); }
                    double rsq = d0 * d0; rsq += d1 * d1; rsq += d2 * d2;
                    pass = !(rsq > h2);
                }
                const unsigned bal = __ballot_sync(0xffffffffu, pass);
                if(pass) s_q[qn + __popc(bal & ltmask)] = o;
                qn += __popc(bal);
                if(kk == 3) { eC = eN; eN = fetch_ent(L, base + 32, g); }
                __syncwarp();
            }
            if(qn >= 32 || (!more && qn > 0)) {
                // ---- hydro_ngbiter (hydra.c:350-505) for up to 32 queued neighbours, one per lane
                if(lane < qn) {
                    const int o = s_q[lane];
                    const double2 qa = spart_xy[o], qb = spart_zm[o];
                    const double4 a_j = hA[o];
                    double d0 = pm.x - qa.x, d1 = pm.y - qa.y, d2 = pm.z - qb.x;
                    if(wrap) { d0 = nearest_s(d0, S.box, S.halfbox); d1 = nearest_s(d1, S.box, S.halfbox); d2 = nearest_s(d2, S.box, S.halfbox); }
                    double rsq = d0 * d0; rsq += d1 * d1; rsq += d2 * d2;
                    if(rsq > 0 && (rsq < ki.HH || rsq < a_j.x * a_j.x)) {
                        const double r = sqrt(rsq), rinv = 1.0 / r, mj = qb.y;
                        const double4 vo = svel[o], b_j = hB[o], c_j = hC[o];
                        const double density_j = a_j.w, p_over_rho2_j = b_j.x, cs_j = b_j.y;
                        Kern kj; kj.H = a_j.x; kj.HH = 0; kj.Hinv = a_j.y; kj.Wknorm = 0; kj.dWknorm = a_j.z;
                        const double vsig = cs_i + cs_j;
                        if(vsig > MaxSig) MaxSig = vsig;
                        const double v0 = vm.x - vo.x, v1 = vm.y - vo.y, v2 = vm.z - vo.z;
                        const double vdotr = d0 * v0 + d1 * v1 + d2 * v2;
                        const double vdotr2 = vdotr + S.hubble_a2 * rsq;
                        const double dwk_i = kern_dw(ki, r * ki.Hinv, S);
                        const double dwk_j = kern_dw(kj, r * kj.Hinv, S);
                        double visc = 0;
                        if(vdotr2 < 0) {
                            const double mu_ij = S.fac_mu * vdotr2 * rinv;
                            const double rho_ij = 0.5 * (dens_i + density_j);
                            double vs = cs_i + cs_j;
                            vs -= 3 * mu_ij;
                            if(vs > MaxSig) MaxSig = vs;
                            visc = 0.25 * S.p.ArtBulkViscConst * vs * (-mu_ij) / rho_ij * (F1 + b_j.z);
                            const double dloga = 2 * (dloga_i > c_j.y ? dloga_i : c_j.y);          // hydra.c:463
                            if(dloga > 0 && (dwk_i + dwk_j) < 0) {
                                const double msum = pm.w + mj;
                                if(msum > 0) {
                                    const double lim = 0.5 * S.fac_vsic_fix * vdotr2 / (0.5 * msum * (dwk_i + dwk_j) * r * dloga);
                                    if(lim < visc) visc = lim;
                                }
                            }
                        }
                        const double hfc_visc = 0.5 * mj * visc * (dwk_i + dwk_j) * rinv;
                        double hfc = hfc_visc;
                        if(DI) hfc += mj * (dwk_i * p_over_rho2_i * (vo.w * inv_evp_i) + dwk_j * p_over_rho2_j * (vm.w * c_j.z)) * rinv;
                        hfc += mj * (pd_i * dwk_i + p_over_rho2_j * c_j.x * dwk_j * b_j.w) * rinv;
                        A0 += (-hfc * d0); A1 += (-hfc * d1); A2 += (-hfc * d2);
                        DtE += (0.5 * hfc_visc * vdotr2);
                    }
                }
                const int mv = lane + 32 < qn ? s_q[32 + lane] : 0;
                __syncwarp();
                if(lane + 32 < qn) s_q[lane] = mv;
                qn = qn > 32 ? qn - 32 : 0;
                __syncwarp();
            }
        }
        A0 = warp_sum(A0); A1 = warp_sum(A1); A2 = warp_sum(A2); DtE = warp_sum(DtE); MaxSig = warp_max(MaxSig);
        ncand = (int) __reduce_add_sync(0xffffffffu, (unsigned) ncand);
        if(lane == t) { rA0 = A0; rA1 = A1; rA2 = A2; rDtE = DtE; rMaxSig = MaxSig; rncand = ncand; rdens = dens_i; }
    }
    if(!valid) return;
    // hydro_postprocess hydra.c:514-528
    rDtE *= GAMMA_MINUS1 / (S.hubble_a2 * pow(rdens, GAMMA_MINUS1));
    acc_out[3 * (int64_t) me] = rA0; acc_out[3 * (int64_t) me + 1] = rA1; acc_out[3 * (int64_t) me + 2] = rA2;
    dte_out[me] = rDtE;
    maxsig_out[me] = rMaxSig;
    if(ninteract) ninteract[me] = rncand;
}

static int make_dev(Engine *E, const b200_sph_params *p, SphDev &S)
{
    if(!p) return failmsg(E, "b200 sph: null params");
    S.p = *p;
    S.box = E->tree_box; S.halfbox = 0.5 * E->tree_box;
    switch(p->KernelType) {                         // densitykernel.c:92-105
        case 1: S.ktype = 0; S.support = 2.; S.sigma = 1 / M_PI; break;
        case 2: S.ktype = 1; S.support = 3.; S.sigma = 1 / (120 * M_PI); break;
        case 4: S.ktype = 2; S.support = 2.5; S.sigma = 1 / (20 * M_PI); break;
        default: return failmsg(E, "b200 sph: unknown DensityKernelType (1 cubic, 2 quintic, 4 quartic)");
    }
    S.desnumngb = NORM_COEFF * pow(S.support * p->DensityResolutionEta, 3);      // densitykernel.c:124-131
    S.fac_mu = pow(p->atime, 3 * (GAMMA - 1) / 2) / p->atime;                    // hydra.c:220-223
    S.fac_vsic_fix = p->hubble * pow(p->atime, 3 * GAMMA_MINUS1);
    S.hubble_a2 = p->hubble * p->atime * p->atime;
    // Time bins: per-particle bins + per-bin tables when b200_sph_set_timebins was called, otherwise
    // every particle on bin 0 with the scalar factors of b200_sph_params.
    const size_t n = (size_t) (E->n > 0 ? E->n : 1);
    if(!E->s_bins_set) {
        CK(E->s_bins.ensure(5 * NB)); CK(E->s_bin_grav.ensure(n)); CK(E->s_bin_hydro.ensure(n));
        double t[5 * NB];
        for(int b = 0; b < NB; b++) {
            t[b] = p->gravkick; t[NB + b] = p->hydrokick; t[2 * NB + b] = p->dloga_pred; t[3 * NB + b] = p->drift; t[4 * NB + b] = p->dloga_bin;
        }
        CK(cudaMemcpyAsync(E->s_bins.p, t, sizeof(t), cudaMemcpyHostToDevice, E->stream));
        CK(cudaStreamSynchronize(E->stream));       // t is a stack buffer
        CK(cudaMemsetAsync(E->s_bin_grav.p, 0, n, E->stream)); CK(cudaMemsetAsync(E->s_bin_hydro.p, 0, n, E->stream));
    }
    S.bins = E->s_bins.p; S.bg = E->s_bin_grav.p; S.bh = E->s_bin_hydro.p;
    return 0;
}

// Targets of a pass over the gas tree: the tree particles (curve order) that are in the caller's
// active set, or all of them.  Returns the list (NULL = identity) and its length.
__global__ void k_sph_active_flags(int np, const int *__restrict__ sidx, const uint8_t *__restrict__ active, uint8_t *__restrict__ out)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if(j < np) out[j] = active[sidx[j]];
}
__global__ void k_sph_mark_active(int64_t na, const int *__restrict__ list, uint8_t *__restrict__ active)
{
    const int64_t q = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if(q < na) active[list[q]] = 1;
}

static int sph_first_targets(Engine *E, int np, int *list_buf, int *iota_buf, const int **tg, int *nt)
{
    *tg = nullptr; *nt = np;
    if(!E->s_active_set || np == 0) return 0;
    CK(E->walk_flags.ensure((size_t) np + 64));
    k_sph_active_flags<<<(np + 255) / 256, 256, 0, E->stream>>>(np, E->sidx.p, E->s_active.p, E->walk_flags.p); CKL(E);
    k_sph_iota<<<(np + 255) / 256, 256, 0, E->stream>>>(iota_buf, np); CKL(E);
    CK(E->scratch_i.ensure(256));
    int *d_num = E->scratch_i.p + 13;
    size_t tb = 0;
    cub::DeviceSelect::Flagged(nullptr, tb, iota_buf, E->walk_flags.p, list_buf, d_num, np, E->stream);
    CK(E->cubtemp.ensure(tb + 16));
    CK(cub::DeviceSelect::Flagged(E->cubtemp.p, tb, iota_buf, E->walk_flags.p, list_buf, d_num, np, E->stream));
    E->launches += 1;
    int cnt = 0;
    CK(cudaMemcpyAsync(&cnt, d_num, sizeof(int), cudaMemcpyDeviceToHost, E->stream));
    CK(cudaStreamSynchronize(E->stream));
    *tg = list_buf; *nt = cnt;
    return 0;
}

int sph_set_timebins(Engine *E, const uint8_t *bin_grav, const uint8_t *bin_hydro, const b200_sph_bins *bins)
{
    if(!bins) return failmsg(E, "b200_sph_set_timebins: null factor tables");
    const size_t n = (size_t) (E->n > 0 ? E->n : 1);
    CK(E->s_bins.ensure(5 * NB)); CK(E->s_bin_grav.ensure(n)); CK(E->s_bin_hydro.ensure(n));
    CK(cudaMemcpyAsync(E->s_bins.p, bins, 5 * NB * sizeof(double), cudaMemcpyHostToDevice, E->stream));
    if(bin_grav) CK(cudaMemcpyAsync(E->s_bin_grav.p, bin_grav, E->n, cudaMemcpyDefault, E->stream));
    else CK(cudaMemsetAsync(E->s_bin_grav.p, 0, n, E->stream));
    if(bin_hydro) CK(cudaMemcpyAsync(E->s_bin_hydro.p, bin_hydro, E->n, cudaMemcpyDefault, E->stream));
    else CK(cudaMemsetAsync(E->s_bin_hydro.p, 0, n, E->stream));
    CK(cudaStreamSynchronize(E->stream));
    E->s_bins_set = true;
    return 0;
}

int sph_set_active(Engine *E, const int32_t *active, int64_t nactive)
{
    E->s_active_set = false;
    if(!active) return 0;
    const size_t n = (size_t) (E->n > 0 ? E->n : 1);
    CK(E->s_active.ensure(n)); CK(E->targets.ensure((size_t) (nactive > 0 ? nactive : 1)));
    CK(cudaMemsetAsync(E->s_active.p, 0, n, E->stream));
    if(nactive > 0) {
        CK(cudaMemcpyAsync(E->targets.p, active, nactive * sizeof(int32_t), cudaMemcpyDefault, E->stream));
        k_sph_mark_active<<<(unsigned) ((nactive + 255) / 256), 256, 0, E->stream>>>(nactive, E->targets.p, E->s_active.p); CKL(E);
    }
    CK(cudaStreamSynchronize(E->stream));
    E->s_active_set = true;
    return 0;
}

int sph_set_state(Engine *E, const double *density, const double *egy, const double *dhsmlfac, const double *divvel, const double *curlvel)
{
    const size_t n = (size_t) (E->n > 0 ? E->n : 1);
    struct { const double *src; DevBuf<double> *dst; } items[] = {
        {density, &E->s_density}, {egy, &E->s_egy}, {dhsmlfac, &E->s_dhsmlfac}, {divvel, &E->s_divvel}, {curlvel, &E->s_curlvel}};
    for(auto &it : items) {
        CK(it.dst->ensure(n));
        if(it.src) CK(cudaMemcpyAsync(it.dst->p, it.src, E->n * sizeof(double), cudaMemcpyDefault, E->stream));
    }
    CK(cudaStreamSynchronize(E->stream));
    return 0;
}

int sph_set_gas(Engine *E, const double *vel, const double *hsml, const double *entropy, const double *dtentropy,
                const double *fullacc, const double *gravpm, const double *hydroacc)
{
    const size_t n = (size_t) (E->n > 0 ? E->n : 1);
    struct { const double *src; DevBuf<double> *dst; size_t k; bool *have; } items[] = {
        {vel, &E->s_vel, 3, &E->s_have[0]}, {hsml, &E->s_hsml, 1, &E->s_have[1]}, {entropy, &E->s_entropy, 1, &E->s_have[2]},
        {dtentropy, &E->s_dtentropy, 1, &E->s_have[3]}, {fullacc, &E->s_fullacc, 3, &E->s_have[4]},
        {gravpm, &E->s_gravpm, 3, &E->s_have[5]}, {hydroacc, &E->s_hydroacc, 3, &E->s_have[6]}};
    for(auto &it : items) {
        *it.have = it.src != nullptr;
        if(!it.src) continue;
        CK(it.dst->ensure(it.k * n));
        CK(cudaMemcpyAsync(it.dst->p, it.src, it.k * E->n * sizeof(double), cudaMemcpyDefault, E->stream));
    }
    if(!hsml) return failmsg(E, "b200_sph_set_gas: hsml is required");
    CK(cudaStreamSynchronize(E->stream));
    E->sph_density_done = false;
    E->s_bins_set = false;
    E->s_active_set = false;
    return 0;
}

// hmax of the tree from the current smoothing lengths of all its gas particles
// (update_tree_hmax_father forcetree.c:1287-1315, force_tree_calc_moments at run.c:477)
int sph_update_hmax(Engine *E)
{
    if(!E->tree_valid) return failmsg(E, "b200_sph_update_hmax: no tree");
    const int nn = (int) E->tree_nn;
    if(E->tree_np == 0) return 0;
    k_sph_hmax_leaf<<<(nn + 255) / 256, 256, 0, E->stream>>>(nn, (const double4 *) E->nodeB.p, (const int4 *) E->nodeC.p,
        (const double4 *) E->spart.p, E->sidx.p, E->s_hsml.p, E->type.p, E->nodeH.p);
    CKL(E);
    for(int level = (int) E->tree_lvl.size() - 2; level >= 0; level--) {
        const int first = E->tree_lvl[level], last = E->tree_lvl[level + 1];
        k_sph_hmax_up<<<(last - first + 127) / 128, 128, 0, E->stream>>>(first, last, E->b_firstchild.p, E->b_nchild.p, E->b_dfs.p, E->nodeH.p);
        CKL(E);
    }
    return 0;
}

// Overwrite the smoothing lengths of the particles [first, first + count) (imported ghosts whose
// owner has converged them) and refresh hmax.
int sph_set_hsml_range(Engine *E, const double *hsml, int64_t first, int64_t count)
{
    if(!E->s_have[1]) return failmsg(E, "b200_sph_set_hsml_range: call b200_sph_set_gas first");
    if(first < 0 || count < 0 || first + count > E->n) return failmsg(E, "b200_sph_set_hsml_range: bad range");
    if(count > 0) CK(cudaMemcpyAsync(E->s_hsml.p + first, hsml, count * sizeof(double), cudaMemcpyDefault, E->stream));
    CK(cudaStreamSynchronize(E->stream));
    return sph_update_hmax(E);
}

int sph_density(Engine *E, const b200_sph_params *p, int update_hsml, int DoEgy, int *d_ninteract, int *d_niter)
{
    if(!E->tree_valid) return failmsg(E, "b200_density: build the gas tree first (b200_tree_build with the gas mask)");
    if(!E->s_have[1]) return failmsg(E, "b200_density: call b200_sph_set_gas first");
    SphDev S;
    if(int rc = make_dev(E, p, S)) return rc;
    const size_t n = (size_t) (E->n > 0 ? E->n : 1);
    const int np = (int) E->tree_np;
    CK(E->s_velpred.ensure(3 * n)); CK(E->s_evp.ensure(n));
    CK(E->s_density.ensure(n)); CK(E->s_egy.ensure(n)); CK(E->s_dhsmlfac.ensure(n)); CK(E->s_divvel.ensure(n));
    CK(E->s_curlvel.ensure(n)); CK(E->s_dthsml.ensure(n)); CK(E->s_numngb.ensure(n)); CK(E->s_gradrho.ensure(3 * n));
    CK(E->s_svel.ensure(4 * (size_t) (np > 0 ? np : 1)));
    CK(E->scratch_i.ensure(256));
    CK(cudaMemsetAsync(E->scratch_i.p + 12, 0, sizeof(int), E->stream));
    CK(E->s_left.ensure(n)); CK(E->s_right.ensure(n)); CK(E->s_niter.ensure(n)); CK(E->s_nint.ensure(n));
    CK(E->walk_flags.ensure((size_t) np + 64));
    timer_start(E, T_SPH_DENSITY);
    if(E->n > 0) {
        k_sph_predict<<<(unsigned) ((E->n + 255) / 256), 256, 0, E->stream>>>(E->n, E->s_have[0] ? E->s_vel.p : nullptr,
            E->s_have[4] ? E->s_fullacc.p : nullptr, E->s_have[5] ? E->s_gravpm.p : nullptr, E->s_have[6] ? E->s_hydroacc.p : nullptr,
            E->s_have[2] ? E->s_entropy.p : nullptr, E->s_have[3] ? E->s_dtentropy.p : nullptr, S, E->s_velpred.p, E->s_evp.p);
        CKL(E);
        k_sph_init_state<<<(unsigned) ((E->n + 255) / 256), 256, 0, E->stream>>>(E->n, S.box, E->s_left.p, E->s_right.p, E->s_niter.p, E->s_nint.p);
        CKL(E);
    }
    if(np > 0) {
        k_sph_gather_vel<<<(np + 255) / 256, 256, 0, E->stream>>>(np, E->sidx.p, E->s_velpred.p, E->s_evp.p, (double4 *) E->s_svel.p);
        CKL(E);
        // treewalk_do_hsml_loop (treewalk.c:1269-1367): walk, evaluate, re-queue the unconverged
        const double keep_estimate = E->walk_chunks_per_warp;
        E->walk_chunks_per_warp = E->sph_chunks_per_warp;
        // pass 0: every tree particle (or the active ones), in curve order
        CK(E->sph_list_a.ensure((size_t) np + 1)); CK(E->sph_list_b.ensure((size_t) np + 1));
        int *tg_next = E->sph_list_a.p, *tg_other = E->sph_list_b.p;
        const int *tg = nullptr;
        int nt = np;
        if(int rc = sph_first_targets(E, np, tg_other, tg_next, &tg, &nt)) return rc;
        int nt0 = nt;
        for(int pass = 0; nt > 0; pass++) {
            // targets per warp: the unconverged targets of the later passes are sparse in space (k_sph_walk)
            static const int force_tpw = getenv("B200_SPH_TPW") ? atoi(getenv("B200_SPH_TPW")) : 0;
            int tpw = (double) nt >= 0.5 * nt0 ? 32 : ((double) nt >= 0.02 * nt0 ? 8 : 1);
            if(force_tpw == 1 || force_tpw == 8 || force_tpw == 32) tpw = force_tpw;
            const int64_t nwarps = (nt + tpw - 1) / tpw;
            const unsigned nb = (unsigned) ((nwarps * 32 + 127) / 128);
            CK(E->walk_partial.ensure((size_t) nwarps * 32 * 4));
            piece_pool_reset(E);
            for(int attempt = 0;; attempt++) {
                PiecePool Q;
                if(int rc = piece_pool_begin(E, nwarps, &Q)) return rc;
                CK(piece_set_smem(k_sph_walk<false>, piece_ctab_bytes(E, WALK_WARPS)));
                k_sph_walk<false><<<nb, 128, piece_ctab_bytes(E, WALK_WARPS), E->stream>>>((const double4 *) E->nodeB.p, (const int4 *) E->nodeC.p, (const int4 *) E->nodeK.p,
                    E->nodeH.p, (const double4 *) E->spart.p, E->sidx.p, tg, nt, tpw, E->s_hsml.p, S, Q, E->walk_partial.p);
                CKL(E);
                bool retry = false;
                if(int rc = piece_pool_check(E, nwarps, &retry, attempt)) return rc;
                if(!retry) break;
            }
            CK(piece_set_smem(k_sph_density_pairs, piece_ctab_bytes(E, WALK_WARPS)));
            k_sph_density_pairs<<<nb, 128, piece_ctab_bytes(E, WALK_WARPS), E->stream>>>(tg, nt, E->sidx.p, (const double4 *) E->spart.p,
                (const double2 *) E->spart_xy.p, (const double2 *) E->spart_zm.p, E->walk_partial.p, (const double4 *) E->s_svel.p,
                S, update_hsml, DoEgy, tpw, E->walk_pool.p, E->walk_chunktab.p, E->walk_maxch, E->walk_cnt.p, 0,
                E->s_hsml.p, E->s_left.p, E->s_right.p, E->s_density.p, E->s_egy.p, E->s_dhsmlfac.p, E->s_divvel.p, E->s_curlvel.p,
                E->s_dthsml.p, E->s_numngb.p, E->s_gradrho.p, E->s_nint.p, E->s_niter.p, E->walk_flags.p, E->scratch_i.p + 12);
            CKL(E);
            if(!update_hsml) break;
            // the unconverged targets of this pass, still in curve order
            size_t tb = 0;
            int *d_num = E->scratch_i.p + 13;
            if(!tg) {       // pass 0 walked the identity list
                k_sph_iota<<<(nt + 255) / 256, 256, 0, E->stream>>>(tg_other, nt); CKL(E);
                tg = tg_other;
            }
            cub::DeviceSelect::Flagged(nullptr, tb, tg, E->walk_flags.p, tg_next, d_num, nt, E->stream);
            CK(E->cubtemp.ensure(tb + 16));
            CK(cub::DeviceSelect::Flagged(E->cubtemp.p, tb, tg, E->walk_flags.p, tg_next, d_num, nt, E->stream));
            E->launches += 1;
            int left_over = 0;
            CK(cudaMemcpyAsync(&left_over, d_num, sizeof(int), cudaMemcpyDeviceToHost, E->stream));
            CK(cudaStreamSynchronize(E->stream));
            E->sph_passes = pass + 1;
            nt = left_over;
            { int *sw = tg_next; tg_next = tg_other; tg_other = sw; }      // tg_other now holds the new list
            tg = tg_other;
        }
        E->sph_chunks_per_warp = E->walk_chunks_per_warp;
        E->walk_chunks_per_warp = keep_estimate;
        if(d_ninteract) CK(cudaMemcpyAsync(d_ninteract, E->s_nint.p, (size_t) E->n * sizeof(int), cudaMemcpyDeviceToDevice, E->stream));
        if(d_niter) CK(cudaMemcpyAsync(d_niter, E->s_niter.p, (size_t) E->n * sizeof(int), cudaMemcpyDeviceToDevice, E->stream));
        if(int rc = sph_update_hmax(E)) return rc;
    }
    timer_stop(E, T_SPH_DENSITY);
    int herr = 0;
    CK(cudaMemcpyAsync(&herr, E->scratch_i.p + 12, sizeof(int), cudaMemcpyDeviceToHost, E->stream));
    CK(cudaStreamSynchronize(E->stream));
    if(herr) return failmsg(E, "b200_density: bad density or smoothing length did not converge for some particles (density.c:541-543, treewalk.c:1361-1363)");
    E->sph_density_done = true;
    E->sph_DoEgy = DoEgy;
    return 0;
}

int sph_hydro(Engine *E, const b200_sph_params *p, double *d_acc, double *d_dte, double *d_maxsig, int *d_ninteract)
{
    if(!E->tree_valid || !E->sph_density_done) return failmsg(E, "b200_hydro_force: hmax not computed, call b200_density first (hydra.c:174-175)");
    SphDev S;
    if(int rc = make_dev(E, p, S)) return rc;
    if(S.p.DensityIndependentSphOn && !E->sph_DoEgy) return failmsg(E, "b200_hydro_force: pressure-entropy SPH needs b200_density with DoEgyDensity=1");
    const int np = (int) E->tree_np;
    CK(E->s_hA.ensure(4 * (size_t) (np > 0 ? np : 1))); CK(E->s_hB.ensure(4 * (size_t) (np > 0 ? np : 1))); CK(E->s_hD.ensure(8 * (size_t) (np > 0 ? np : 1)));
    timer_start(E, T_SPH_HYDRO);
    if(np > 0) {
        k_sph_gather_hydro<<<(np + 255) / 256, 256, 0, E->stream>>>(np, E->sidx.p, S, E->s_hsml.p, E->s_density.p, E->s_egy.p, E->s_dhsmlfac.p,
            E->s_divvel.p, E->s_curlvel.p, (const double4 *) E->s_svel.p, (double4 *) E->s_hA.p, (double4 *) E->s_hB.p, (double4 *) E->s_hD.p,
            (double4 *) E->s_hD.p + (size_t) (np > 0 ? np : 1));
        CKL(E);
        CK(E->sph_list_a.ensure((size_t) np + 1)); CK(E->sph_list_b.ensure((size_t) np + 1));
        const int *tg = nullptr;
        int nt = np;
        if(int rc = sph_first_targets(E, np, E->sph_list_a.p, E->sph_list_b.p, &tg, &nt)) return rc;
        const double keep_estimate = E->walk_chunks_per_warp;
        E->walk_chunks_per_warp = E->sph_chunks_per_warp;
        const int64_t nwarps = (nt + 31) / 32;
        const unsigned nb = (unsigned) ((nwarps * 32 + 127) / 128);
        CK(E->walk_partial.ensure((size_t) (nwarps > 0 ? nwarps : 1) * 32 * 4));
        piece_pool_reset(E);
        for(int attempt = 0; nt > 0; attempt++) {
            PiecePool Q;
            if(int rc = piece_pool_begin(E, nwarps, &Q)) return rc;
            CK(piece_set_smem(k_sph_walk<true>, piece_ctab_bytes(E, WALK_WARPS)));
            k_sph_walk<true><<<nb, 128, piece_ctab_bytes(E, WALK_WARPS), E->stream>>>((const double4 *) E->nodeB.p, (const int4 *) E->nodeC.p, (const int4 *) E->nodeK.p,
                E->nodeH.p, (const double4 *) E->spart.p, E->sidx.p, tg, nt, 32, E->s_hsml.p, S, Q, E->walk_partial.p);
            CKL(E);
            bool retry = false;
            if(int rc = piece_pool_check(E, nwarps, &retry, attempt)) return rc;
            if(!retry) break;
        }
        E->sph_chunks_per_warp = E->walk_chunks_per_warp;
        E->walk_chunks_per_warp = keep_estimate;
        CK(piece_set_smem(k_sph_hydro_pairs, piece_ctab_bytes(E, WALK_WARPS)));
        if(nt > 0)
        k_sph_hydro_pairs<<<nb, 128, piece_ctab_bytes(E, WALK_WARPS), E->stream>>>(tg, nt, E->sidx.p, (const double4 *) E->spart.p,
            (const double2 *) E->spart_xy.p, (const double2 *) E->spart_zm.p, E->walk_partial.p, (const double4 *) E->s_svel.p,
            (const double4 *) E->s_hA.p, (const double4 *) E->s_hB.p, (const double4 *) E->s_hD.p, (const double4 *) E->s_hD.p + (size_t) (np > 0 ? np : 1), E->s_density.p, S,
            E->walk_pool.p, E->walk_chunktab.p, E->walk_maxch, E->walk_cnt.p, d_acc, d_dte, d_maxsig, d_ninteract);
        CKL(E);
    }
    timer_stop(E, T_SPH_HYDRO);
    return 0;
}

} // namespace b200
