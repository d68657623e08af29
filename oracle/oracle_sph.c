/* oracle_sph.c -- CPU restatement of the MP-Gadget SPH density (with the
 * smoothing-length iteration) and hydro-force loops for a synchronised step
 * (all gas particles on one time bin).  TEST INFRASTRUCTURE ONLY (oracle.h).
 *
 * Follows: density.c:424-519 (density_ngbiter), :532-586 (postprocess),
 * :589-689 (density_check_neighbours), :69-100 (EntVarPred, VelPred);
 * treewalk.c:1015-1042 (cull_node), :1152-1265 (radius search without list),
 * :930-1007,1056-1143 (symmetric search); densitykernel.c:24-172;
 * hydra.c:67-77,247-277,300-312,318-506,514-528; forcetree.c:1287-1315 (hmax).
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "oracle.h"

#define GAMMA (5.0 / 3.0)
#define GAMMA_MINUS1 (GAMMA - 1)
#define NORM_COEFF 4.188790204786
#define FACT1 0.366025403785      /* treewalk.c:19 */
#define MAXITER 400

static inline double nearest(double x, double box)
{
    return (x > 0.5 * box) ? (x - box) : ((x < -0.5 * box) ? (x + box) : x);
}

typedef struct { double H, HH, Hinv, support, Wknorm, dWknorm; int type; } kern;

static void kern_init(kern *k, double H, int type)       /* densitykernel.c:140-172 */
{
    static const double supp[3] = {2., 3., 2.5};
    static const double sig[3] = {1 / M_PI, 1 / (120 * M_PI), 1 / (20 * M_PI)};
    const int t = type == 1 ? 0 : (type == 2 ? 1 : 2);
    k->H = H; k->HH = H * H; k->Hinv = 1. / H; k->type = t; k->support = supp[t];
    const double hinv = k->Hinv * k->support;
    k->Wknorm = sig[t] * pow(hinv, 3);
    k->dWknorm = k->Wknorm * hinv;
}
static double kern_w(const kern *k, double u)             /* densitykernel.c:24-90,116-122 */
{
    const double q = u * k->support;
    double w = 0;
    if(k->type == 0) {
        if(q < 1.0) w = 0.25 * pow(2 - q, 3) - pow(1 - q, 3);
        else if(q < 2.0) w = 0.25 * pow(2 - q, 3);
    } else if(k->type == 1) {
        if(q < 1.0) w = pow(3 - q, 5) - 6 * pow(2 - q, 5) + 15 * pow(1 - q, 5);
        else if(q < 2.0) w = pow(3 - q, 5) - 6 * pow(2 - q, 5);
        else if(q < 3.0) w = pow(3 - q, 5);
    } else {
        if(q < 0.5) w = pow(2.5 - q, 4) - 5 * pow(1.5 - q, 4) + 10 * pow(0.5 - q, 4);
        else if(q < 1.5) w = pow(2.5 - q, 4) - 5 * pow(1.5 - q, 4);
        else if(q < 2.5) w = pow(2.5 - q, 4);
    }
    return k->Wknorm * w;
}
static double kern_dw(const kern *k, double u)            /* densitykernel.c:32-90,108-114 */
{
    const double q = u * k->support;
    double w = 0;
    if(k->type == 0) {
        if(q < 1.0) w = -0.25 * 3 * pow(2 - q, 2) + 3 * pow(1 - q, 2);
        else if(q < 2.0) w = -0.25 * 3 * pow(2 - q, 2);
    } else if(k->type == 1) {
        if(q < 1.0) w = -5 * pow(3 - q, 4) + 30 * pow(2 - q, 4) - 75 * pow(1 - q, 4);
        else if(q < 2.0) w = -5 * pow(3 - q, 4) + 30 * pow(2 - q, 4);
        else if(q < 3.0) w = -5 * pow(3 - q, 4);
    } else {
        if(q < 0.5) w = -4 * pow(2.5 - q, 3) + 20 * pow(1.5 - q, 3) - 40 * pow(0.5 - q, 3);
        else if(q < 1.5) w = -4 * pow(2.5 - q, 3) + 20 * pow(1.5 - q, 3);
        else if(q < 2.5) w = -4 * pow(2.5 - q, 3);
    }
    return k->dWknorm * w;
}

double oracle_sph_desnumngb(int kerneltype, double eta)   /* densitykernel.c:124-131 */
{
    kern k; kern_init(&k, 1.0, kerneltype);
    return NORM_COEFF * pow(k.support * eta, 3);
}

static double entvarpred(double Entropy, double DtEntropy, double dloga)    /* density.c:69-85 */
{
    double e = Entropy + DtEntropy * dloga;
    if(e < 0.05 * Entropy) e = 0.05 * Entropy;
    if(e <= 0) return 0;
    return exp(1. / GAMMA * log(e));
}

/* cull_node treewalk.c:1015-1042 */
static int cull(const oracle_node *nd, const double *pos, double hsml, int symmetric, double box)
{
    double dist = (symmetric ? (nd->hmax > hsml ? nd->hmax : hsml) : hsml) + 0.5 * nd->len;
    double r2 = 0;
    for(int d = 0; d < 3; d++) {
        const double dx = nearest(nd->center[d] - pos[d], box);
        if(dx > dist) return 0;
        if(dx < -dist) return 0;
        r2 += dx * dx;
    }
    dist += FACT1 * nd->len;
    if(r2 > dist * dist) return 0;
    return 1;
}

/* Smoothing-length update: density_check_neighbours density.c:589-689 (gas only).
 * Returns 1 when done. */
static int check_neighbours(double *Hsml, double *Left, double *Right, double NumNgb, double DensFac,
                            double desnumngb, double maxdev, double box, double MinGasHsml)
{
    if(NumNgb < (desnumngb - maxdev) || NumNgb > (desnumngb + maxdev)) {
        if((*Right - *Left) < 1.0e-5 * *Left) { *Hsml = *Right; return 1; }
        if(NumNgb < desnumngb) *Left = *Hsml; else *Right = *Hsml;
        if((*Right < box && *Left > 0) || (*Hsml * 1.26 > 0.99 * box))
            *Hsml = cbrt(0.5 * (pow(*Left, 3) + pow(*Right, 3)));
        else {
            double fac = 1.26;
            if(NumNgb > 0) fac = 1 - (NumNgb - desnumngb) / (3 * NumNgb) * DensFac;
            if(*Right > 0.99 * box && *Left > 0)
                if(DensFac <= 0 || fabs(NumNgb - desnumngb) >= 0.5 * desnumngb || fac > 1.26) fac = 1.26;
            if(*Right < 0.99 * box && *Left == 0)
                if(DensFac <= 0 || fac < 1. / 3) fac = 1. / 3;
            *Hsml *= fac;
        }
        if(*Right < MinGasHsml) { *Hsml = MinGasHsml; return 1; }
        return 0;
    }
    if(*Hsml < MinGasHsml) *Hsml = MinGasHsml;
    return 1;
}

/* density() for gas targets; tree must hold the gas particles (GASMASK).
 * Arrays are indexed by particle index.  vel/acc inputs may be NULL (= 0). */
/* ---- mixed time bins / active sets ------------------------------------------------------
 * oracle_sph_set_mixed installs, for the following oracle_density / oracle_hydro calls, the
 * particles' time bins (TimeBinGravity, TimeBinHydro), the per-bin factor tables
 * tab[5][ORACLE_NBINS] = gravkick, hydrokick, dloga_pred, drift, dloga_bin (what the reference
 * derives from DriftKickTimes: density.c:74,114-132, hydra.c:178-186,271,463) and the active
 * flags (ActiveParticles, timestep.h:29-38): only flagged particles are targets, the others keep
 * the state passed in and act as neighbours.  NULL bins resets to the synchronised mode. */
static const uint8_t *mx_bg, *mx_bh, *mx_active;
static const double *mx_tab;
void oracle_sph_set_mixed(const uint8_t *bin_grav, const uint8_t *bin_hydro, const double *tab, const uint8_t *active)
{
    mx_bg = bin_grav; mx_bh = bin_hydro; mx_tab = tab; mx_active = active;
}
static inline double f_gravkick(const oracle_sph_params *sp, int64_t i) { return mx_tab ? mx_tab[mx_bg[i]] : sp->gravkick; }
static inline double f_hydrokick(const oracle_sph_params *sp, int64_t i) { return mx_tab ? mx_tab[ORACLE_NBINS + mx_bh[i]] : sp->hydrokick; }
static inline double f_dloga_pred(const oracle_sph_params *sp, int64_t i) { return mx_tab ? mx_tab[2 * ORACLE_NBINS + mx_bh[i]] : sp->dloga_pred; }
static inline double f_drift(const oracle_sph_params *sp, int64_t i) { return mx_tab ? mx_tab[3 * ORACLE_NBINS + mx_bh[i]] : sp->drift; }
static inline double f_dloga_bin(const oracle_sph_params *sp, int64_t i) { return mx_tab ? mx_tab[4 * ORACLE_NBINS + mx_bh[i]] : sp->dloga_bin; }

int oracle_density(oracle_tree *t, const double *pos, const float *mass, const uint8_t *type, int64_t n,
                   const oracle_sph_params *sp, int update_hsml, int DoEgyDensity,
                   const double *vel, const double *fullacc, const double *gravpm, const double *hydroacc,
                   const double *entropy, const double *dtentropy,
                   double *hsml /*in/out*/, double *density, double *egywtdensity, double *dhsmlfac,
                   double *divvel, double *curlvel, double *dthsml, double *numngb_out, int32_t *ninteract, int32_t *niter_out,
                   double *entvarpred_out)
{
    const double box = t->BoxSize;
    const double desnumngb = oracle_sph_desnumngb(sp->KernelType, sp->DensityResolutionEta);
    double *velpred = (double *) malloc(sizeof(double) * 3 * n);
    double *evp = (double *) malloc(sizeof(double) * n);
    for(int64_t i = 0; i < n; i++) {
        for(int j = 0; j < 3; j++)                    /* SPH_VelPred density.c:91-100 */
            velpred[3 * i + j] = (vel ? vel[3 * i + j] : 0) + f_gravkick(sp, i) * (fullacc ? fullacc[3 * i + j] : 0)
                               + (gravpm ? gravpm[3 * i + j] : 0) * sp->pmkick + f_hydrokick(sp, i) * (hydroacc ? hydroacc[3 * i + j] : 0);
        evp[i] = entvarpred(entropy ? entropy[i] : 1.0, dtentropy ? dtentropy[i] : 0.0, f_dloga_pred(sp, i));
        if(entvarpred_out) entvarpred_out[i] = evp[i];
    }
    const oracle_node *N = t->nodes;
    int bad = 0;
#pragma omp parallel for schedule(dynamic, 64)
    for(int64_t i = 0; i < n; i++) {
        if(type && type[i] != 0) continue;             /* density_haswork: gas (BH not modelled here) */
        if(mx_active && !mx_active[i]) continue;       /* not in the active set: state untouched */
        double Left = 0, Right = box, h = hsml[i];
        double Ngb = 0, Rho = 0, Dh = 0, EgyRho = 0, DhEgy = 0, Div = 0, Rot[3] = {0, 0, 0}, DhsmlDens = 0;
        int nint = 0, it = 0;
        for(it = 0; it < MAXITER + 2; it++) {
            kern k; kern_init(&k, h, sp->KernelType);
            const double vol = NORM_COEFF * pow(k.H, 3);
            Ngb = Rho = Dh = EgyRho = DhEgy = Div = 0; Rot[0] = Rot[1] = Rot[2] = 0; nint = 0;
            int no = t->numnodes > 0 ? 0 : -1;
            while(no >= 0) {                              /* treewalk.c:1170-1259 */
                const oracle_node *nd = &N[no];
                if(!cull(nd, &pos[3 * i], h, 0, box)) { no = nd->sibling; continue; }
                if(nd->nocc < 0) { no = nd->firstchild; continue; }
                for(int c = 0; c < nd->nocc; c++) {
                    const int64_t o = nd->part[c];
                    double dist[3], r2 = 0; const double h2 = h * h;
                    int d;
                    for(d = 0; d < 3; d++) {
                        dist[d] = nearest(pos[3 * i + d] - pos[3 * o + d], box);
                        r2 += dist[d] * dist[d];
                        if(r2 > h2) break;
                    }
                    if(r2 > h2) continue;
                    nint++;
                    const double r = sqrt(r2);
                    if(r2 < k.HH) {                       /* density_ngbiter density.c:451-518 */
                        const double u = r * k.Hinv;
                        const double wk = kern_w(&k, u);
                        Ngb += wk * vol;
                        const double dwk = kern_dw(&k, u);
                        const double mj = mass[o];
                        Rho += mj * wk;
                        const double dW = -(3 * k.Hinv * wk + u * dwk);
                        Dh += mj * dW;
                        if(DoEgyDensity) { EgyRho += mj * evp[o] * wk; DhEgy += mj * evp[o] * dW; }
                        if(r > 0) {
                            const double fac = mj * dwk / r;
                            double dv[3];
                            for(d = 0; d < 3; d++) dv[d] = velpred[3 * i + d] - velpred[3 * o + d];
                            Div += -fac * (dist[0] * dv[0] + dist[1] * dv[1] + dist[2] * dv[2]);
                            Rot[0] += fac * (dv[1] * dist[2] - dist[1] * dv[2]);
                            Rot[1] += fac * (dv[2] * dist[0] - dist[2] * dv[0]);
                            Rot[2] += fac * (dv[0] * dist[1] - dist[0] * dv[1]);
                        }
                    }
                }
                no = nd->sibling;
            }
            /* density_postprocess density.c:532-586 */
            if(Rho <= 0 && Ngb > 0) bad = 1;
            DhsmlDens = Dh * h / (3 * Rho);
            DhsmlDens = 1 / (1 + DhsmlDens);
            if(!update_hsml) break;
            if(check_neighbours(&h, &Left, &Right, Ngb, DhsmlDens, desnumngb, sp->MaxNumNgbDeviation, box, sp->MinGasHsml)) break;
            if(it > MAXITER) { bad = 1; break; }
        }
        /* the remaining post-processing uses the Hsml of the evaluation (density.c:556-580 runs in the same
         * postprocess call that accepted or replaced Hsml; P[i].Hsml may have been clamped there) */
        hsml[i] = h;
        density[i] = Rho;
        if(DoEgyDensity) {
            double f = DhEgy * h / (3 * EgyRho);
            f *= -DhsmlDens;
            dhsmlfac[i] = f;
            egywtdensity[i] = EgyRho / evp[i];
        } else {
            dhsmlfac[i] = DhsmlDens;
            egywtdensity[i] = 0;
        }
        curlvel[i] = sqrt(Rot[0] * Rot[0] + Rot[1] * Rot[1] + Rot[2] * Rot[2]) / Rho;
        divvel[i] = Div / Rho;
        dthsml[i] = (1.0 / 3) * divvel[i] * h;
        if(numngb_out) numngb_out[i] = Ngb;
        if(ninteract) ninteract[i] = nint;
        if(niter_out) niter_out[i] = it + 1;
    }
    /* leaf hmax from the converged smoothing lengths (update_tree_hmax_father forcetree.c:1287-1315),
     * then bottom-up max (forcetree.c:1090-1091); DFS order => children follow parents */
    for(int64_t k = 0; k < t->numnodes; k++) t->nodes[k].hmax = 0;
    for(int64_t k = t->numnodes - 1; k >= 0; k--) {
        oracle_node *nd = &t->nodes[k];
        if(nd->nocc >= 0) {
            for(int c = 0; c < nd->nocc; c++) {
                const int64_t o = nd->part[c];
                for(int j = 0; j < 3; j++) {
                    const double v = fabs(pos[3 * o + j] - nd->center[j]) + hsml[o] - nd->len / 2.;
                    if(v > nd->hmax) nd->hmax = v;
                }
            }
        }
        if(nd->father >= 0 && nd->hmax > t->nodes[nd->father].hmax) t->nodes[nd->father].hmax = nd->hmax;
    }
    free(velpred); free(evp);
    return bad;
}

/* set_init_hsml density.c:700-749: climb from the particle's leaf until the node
 * holds 10*DesNumNgb particle masses, then scale the node size. */
void oracle_set_init_hsml(const oracle_tree *t, const float *mass, const uint8_t *type, int64_t n,
                          int kerneltype, double eta, double MeanGasSeparation, double *hsml)
{
    const double DesNumNgb = oracle_sph_desnumngb(kerneltype, eta);
    int32_t *father = (int32_t *) malloc(sizeof(int32_t) * n);
    for(int64_t i = 0; i < n; i++) father[i] = -1;
    for(int64_t k = 0; k < t->numnodes; k++)
        for(int c = 0; c < t->nodes[k].nocc; c++) father[t->nodes[k].part[c]] = (int32_t) k;
    for(int64_t i = 0; i < n; i++) {
        if(type && type[i] != 0 && type[i] != 5) continue;
        int no = father[i];
        if(no < 0) continue;
        while(10 * DesNumNgb * mass[i] > t->nodes[no].mass) {
            const int p = t->nodes[no].father;
            if(p < 0) break;
            no = p;
        }
        hsml[i] = MeanGasSeparation;
        const double test = t->nodes[no].len * pow(3.0 / (4 * M_PI) * DesNumNgb * mass[i] / t->nodes[no].mass, 1.0 / 3);
        if(test < 500. * MeanGasSeparation) hsml[i] = test;
    }
    free(father);
}

static double density_pred(double Density, double DivVel, double dtdrift)     /* hydra.c:300-312 */
{
    const double p = Density - DivVel * Density * dtdrift;
    return p >= 1e-6 * Density ? p : 1e-6 * Density;
}
static double pressure_pred(double eom, double evp)                            /* hydra.c:67-77 */
{
    if(evp * eom <= 0) return 0;
    return exp(GAMMA * log(evp * eom));
}

/* hydro_force hydra.c:153-528; tree must carry hmax (oracle_density). */
int oracle_hydro(const oracle_tree *t, const double *pos, const float *mass, const uint8_t *type, int64_t n,
                 const oracle_sph_params *sp,
                 const double *vel, const double *fullacc, const double *gravpm, const double *hydroacc_in,
                 const double *entropy, const double *dtentropy_in,
                 const double *hsml, const double *density, const double *egywtdensity, const double *dhsmlfac,
                 const double *divvel, const double *curlvel,
                 double *acc_out, double *dtentropy_out, double *maxsignalvel_out, int32_t *ninteract)
{
    const double box = t->BoxSize;
    const int DI = sp->DensityIndependentSphOn;
    const double fac_mu = pow(sp->atime, 3 * (GAMMA - 1) / 2) / sp->atime;     /* hydra.c:220-223 */
    const double fac_vsic_fix = sp->hubble * pow(sp->atime, 3 * GAMMA_MINUS1);
    const double hubble_a2 = sp->hubble * sp->atime * sp->atime;
    double *velpred = (double *) malloc(sizeof(double) * 3 * n);
    double *evp = (double *) malloc(sizeof(double) * n);
    double *press = (double *) malloc(sizeof(double) * n);
    for(int64_t i = 0; i < n; i++) {
        for(int j = 0; j < 3; j++)
            velpred[3 * i + j] = (vel ? vel[3 * i + j] : 0) + f_gravkick(sp, i) * (fullacc ? fullacc[3 * i + j] : 0)
                               + (gravpm ? gravpm[3 * i + j] : 0) * sp->pmkick + f_hydrokick(sp, i) * (hydroacc_in ? hydroacc_in[3 * i + j] : 0);
        evp[i] = entvarpred(entropy ? entropy[i] : 1.0, dtentropy_in ? dtentropy_in[i] : 0.0, f_dloga_pred(sp, i));
        const double eom = density_pred(DI ? egywtdensity[i] : density[i], divvel[i], f_drift(sp, i));   /* hydra.c:205-212 */
        press[i] = evp[i] == 0 ? 0 : pressure_pred(eom, evp[i]);
    }
    const oracle_node *N = t->nodes;
#pragma omp parallel
    {
        int32_t *cand = (int32_t *) malloc(sizeof(int32_t) * (t->numparticles + 8));
#pragma omp for schedule(dynamic, 64)
        for(int64_t i = 0; i < n; i++) {
            if(type && type[i] != 0) continue;
            if(mx_active && !mx_active[i]) continue;
            /* hydro_copy hydra.c:247-277 */
            const double eom_i = DI ? egywtdensity[i] : density[i];
            const double P_i = press[i];
            const double cs_i = sqrt(GAMMA * P_i / eom_i);
            const double F1 = fabs(divvel[i]) / (fabs(divvel[i]) + curlvel[i] + 0.0001 * cs_i / hsml[i] / fac_mu);
            const double p_over_rho2_i = P_i / (eom_i * eom_i);
            kern ki; kern_init(&ki, hsml[i], sp->KernelType);
            double Acc[3] = {0, 0, 0}, DtE = 0, MaxSig = cs_i;
            /* ngb_treefind_threads treewalk.c:1056-1143 (symmetric) */
            int64_t numcand = 0;
            int no = t->numnodes > 0 ? 0 : -1;
            while(no >= 0) {
                const oracle_node *nd = &N[no];
                if(!cull(nd, &pos[3 * i], hsml[i], 1, box)) { no = nd->sibling; continue; }
                if(nd->nocc < 0) { no = nd->firstchild; continue; }
                for(int c = 0; c < nd->nocc; c++) cand[numcand++] = nd->part[c];
                no = nd->sibling;
            }
            if(ninteract) ninteract[i] = (int32_t) numcand;
            for(int64_t q = 0; q < numcand; q++) {        /* treewalk.c:962-999 */
                const int64_t o = cand[q];
                const double hmaxij = hsml[o] > hsml[i] ? hsml[o] : hsml[i];
                double dist[3], rsq = 0; const double h2 = hmaxij * hmaxij;
                int d;
                for(d = 0; d < 3; d++) {
                    dist[d] = nearest(pos[3 * i + d] - pos[3 * o + d], box);
                    rsq += dist[d] * dist[d];
                    if(rsq > h2) break;
                }
                if(rsq > h2) continue;
                const double r = sqrt(rsq);
                /* hydro_ngbiter hydra.c:350-505 */
                kern kj; kern_init(&kj, hsml[o], sp->KernelType);
                if(rsq <= 0 || !(rsq < ki.HH || rsq < kj.HH)) continue;
                const double density_j = density_pred(density[o], divvel[o], f_drift(sp, o));
                const double eom_j = density_pred(DI ? egywtdensity[o] : density[o], divvel[o], f_drift(sp, o));
                const double P_j = press[o];
                const double p_over_rho2_j = P_j / (eom_j * eom_j);
                const double cs_j = sqrt(GAMMA * P_j / eom_j);
                double vsig = cs_i + cs_j;
                if(vsig > MaxSig) MaxSig = vsig;
                double dv[3];
                for(d = 0; d < 3; d++) dv[d] = velpred[3 * i + d] - velpred[3 * o + d];
                const double vdotr = dist[0] * dv[0] + dist[1] * dv[1] + dist[2] * dv[2];
                const double vdotr2 = vdotr + hubble_a2 * rsq;
                const double dwk_i = kern_dw(&ki, r * ki.Hinv);
                const double dwk_j = kern_dw(&kj, r * kj.Hinv);
                double visc = 0;
                if(vdotr2 < 0) {
                    const double mu_ij = fac_mu * vdotr2 / r;
                    const double rho_ij = 0.5 * (density[i] + density_j);
                    double vs = cs_i + cs_j;
                    vs -= 3 * mu_ij;
                    if(vs > MaxSig) MaxSig = vs;
                    const double f2 = fabs(divvel[o]) / (fabs(divvel[o]) + curlvel[o] + 0.0001 * cs_j / fac_mu / hsml[o]);
                    visc = 0.25 * sp->ArtBulkViscConst * vs * (-mu_ij) / rho_ij * (F1 + f2);
                    const double dl_i = f_dloga_bin(sp, i), dl_o = f_dloga_bin(sp, o);
                    const double dloga = 2 * (dl_i > dl_o ? dl_i : dl_o);      /* hydra.c:463 */
                    if(dloga > 0 && (dwk_i + dwk_j) < 0) {
                        const double msum = (double) mass[i] + (double) mass[o];      /* I->Mass is MyFloat = double */
                        if(msum > 0) {
                            const double lim = 0.5 * fac_vsic_fix * vdotr2 / (0.5 * msum * (dwk_i + dwk_j) * r * dloga);
                            if(lim < visc) visc = lim;
                        }
                    }
                }
                const double hfc_visc = 0.5 * mass[o] * visc * (dwk_i + dwk_j) / r;
                double hfc = hfc_visc, rr1 = 1, rr2 = 1;
                if(DI) {
                    rr1 = 0; rr2 = 0;
                    hfc += mass[o] * (dwk_i * p_over_rho2_i * evp[o] / evp[i] + dwk_j * p_over_rho2_j * evp[i] / evp[o]) / r;
                    if(sp->DensityContrastLimit >= 0) {
                        rr1 = egywtdensity[i] / density[i];
                        rr2 = eom_j / density_j;
                        if(sp->DensityContrastLimit > 0) {
                            if(sp->DensityContrastLimit < rr1) rr1 = sp->DensityContrastLimit;
                            if(sp->DensityContrastLimit < rr2) rr2 = sp->DensityContrastLimit;
                        }
                    }
                }
                hfc += mass[o] * (p_over_rho2_i * dhsmlfac[i] * dwk_i * rr1 + p_over_rho2_j * dhsmlfac[o] * dwk_j * rr2) / r;
                for(d = 0; d < 3; d++) Acc[d] += (-hfc * dist[d]);
                DtE += (0.5 * hfc_visc * vdotr2);
            }
            /* hydro_postprocess hydra.c:514-528 */
            DtE *= GAMMA_MINUS1 / (hubble_a2 * pow(density[i], GAMMA_MINUS1));
            for(int d = 0; d < 3; d++) acc_out[3 * i + d] = Acc[d];
            dtentropy_out[i] = DtE;
            maxsignalvel_out[i] = MaxSig;
        }
        free(cand);
    }
    free(velpred); free(evp); free(press);
    return 0;
}
