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
// One thread owns one gas particle and walks the tree depth-first exactly like
// the reference visitor (cull_node treewalk.c:1015-1042, leaf particles in
// insertion order), so every sum is accumulated in the reference's order.  The
// reference re-queues unconverged particles and re-runs the whole walk
// (NPRedo, treewalk.c:1292-1364); here the bracket/Newton update of
// density_check_neighbours is iterated inside the thread -- the update only
// reads the particle's own sums, so the result is the same.  Targets are taken
// in curve order, so the threads of a warp walk nearly the same nodes.
// This file is compiled with -fmad=false: the arithmetic is the CPU's.
#include "engine.h"
#include <math.h>

namespace b200 {

#define GAMMA (5.0 / 3.0)
#define GAMMA_MINUS1 (GAMMA - 1)
#define NORM_COEFF 4.188790204786
#define FACT1 0.366025403785      // treewalk.c:19
#define SPH_MAXITER 400

struct SphDev {
    b200_sph_params p;
    double box, halfbox, desnumngb;
    double fac_mu, fac_vsic_fix, hubble_a2;
    int ktype;          // 0 cubic, 1 quintic, 2 quartic
    double support, sigma;
};

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
    for(int j = 0; j < 3; j++)
        velpred[3 * i + j] = (vel ? vel[3 * i + j] : 0) + S.p.gravkick * (fullacc ? fullacc[3 * i + j] : 0)
                           + (gravpm ? gravpm[3 * i + j] : 0) * S.p.pmkick + S.p.hydrokick * (hydroacc ? hydroacc[3 * i + j] : 0);
    const double E = entropy ? entropy[i] : 1.0;
    double e = E + (dtentropy ? dtentropy[i] : 0.0) * S.p.dloga_pred;
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

__global__ void __launch_bounds__(128)
k_sph_density(int np, const int *__restrict__ sidx, const double4 *__restrict__ nodeB, const int4 *__restrict__ nodeC,
              const double4 *__restrict__ spart, const double4 *__restrict__ svel, const uint8_t *__restrict__ type,
              SphDev S, int update_hsml, int DoEgy,
              double *__restrict__ hsml, double *__restrict__ density, double *__restrict__ egy, double *__restrict__ dhsmlfac,
              double *__restrict__ divvel, double *__restrict__ curlvel, double *__restrict__ dthsml, double *__restrict__ numngb,
              double *__restrict__ gradrho, int *__restrict__ ninteract, int *__restrict__ niter, int *__restrict__ err)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if(j >= np) return;
    const int me = sidx[j];
    if(type[me] != 0) return;                      // density_haswork density.c:521-530 (gas; black holes not modelled)
    const double4 pm = spart[j];
    const double4 vm = svel[j];
    double Left = 0, Right = S.box, h = hsml[me];
    double Ngb = 0, Rho = 0, Dh = 0, EgyRho = 0, DhEgy = 0, Div = 0, R0 = 0, R1 = 0, R2 = 0, DhsmlDens = 0;
    double G0 = 0, G1 = 0, G2 = 0;
    int nint = 0, it = 0;
    for(it = 0; it < SPH_MAXITER + 2; it++) {
        Kern k; kern_init(k, h, S);
        const double vol = NORM_COEFF * pw3(k.H);
        const double h2 = h * h;
        Ngb = Rho = Dh = EgyRho = DhEgy = Div = R0 = R1 = R2 = G0 = G1 = G2 = 0; nint = 0;
        int no = 0;
        while(no >= 0) {
            const double4 B = nodeB[no];
            const int4 C = nodeC[no];
            if(!cull_keep(B, 0.0, pm.x, pm.y, pm.z, h, false, S)) { no = C.x; continue; }
            if(!C.w) { no = no + 1; continue; }
            for(int c = 0; c < C.z; c++) {
                const double4 q = spart[C.y + c];
                // treewalk.c:1223-1233
                const double d0 = nearest_s(pm.x - q.x, S.box, S.halfbox);
                double r2 = d0 * d0;
                if(r2 > h2) continue;
                const double d1 = nearest_s(pm.y - q.y, S.box, S.halfbox);
                r2 += d1 * d1;
                if(r2 > h2) continue;
                const double d2 = nearest_s(pm.z - q.z, S.box, S.halfbox);
                r2 += d2 * d2;
                if(r2 > h2) continue;
                nint++;
                if(r2 < k.HH) {                       // density_ngbiter density.c:451-518
                    const double r = sqrt(r2);
                    const double u = r * k.Hinv;
                    const double wk = kern_w(k, u, S);
                    Ngb += wk * vol;
                    const double dwk = kern_dw(k, u, S);
                    const double mj = q.w;
                    Rho += mj * wk;
                    const double dW = -(3 * k.Hinv * wk + u * dwk);
                    Dh += mj * dW;
                    const double4 vo = svel[C.y + c];
                    if(DoEgy) { EgyRho += mj * vo.w * wk; DhEgy += mj * vo.w * dW; }
                    if(r > 0) {
                        const double fac = mj * dwk / r;
                        const double v0 = vm.x - vo.x, v1 = vm.y - vo.y, v2 = vm.z - vo.z;
                        Div += -fac * (d0 * v0 + d1 * v1 + d2 * v2);
                        R0 += fac * (v1 * d2 - d1 * v2);
                        R1 += fac * (v2 * d0 - d2 * v0);
                        R2 += fac * (v0 * d1 - d0 * v1);
                        G0 += fac * d0; G1 += fac * d1; G2 += fac * d2;         // density.c:512-515
                    }
                }
            }
            no = C.x;
        }
        // density_postprocess density.c:532-586
        if(Rho <= 0 && Ngb > 0) atomicAdd(err, 1);
        DhsmlDens = Dh * h / (3 * Rho);
        DhsmlDens = 1 / (1 + DhsmlDens);
        if(!update_hsml) break;
        // density_check_neighbours density.c:589-689
        const double des = S.desnumngb, maxdev = S.p.MaxNumNgbDeviation;
        bool done;
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
        if(done) break;
        if(it > SPH_MAXITER) { atomicAdd(err, 1); break; }
    }
    hsml[me] = h;
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
    curlvel[me] = sqrt(R0 * R0 + R1 * R1 + R2 * R2) / Rho;
    const double dv = Div / Rho;
    divvel[me] = dv;
    dthsml[me] = (1.0 / 3) * dv * h;
    if(numngb) numngb[me] = Ngb;
    gradrho[3 * (int64_t) me] = G0; gradrho[3 * (int64_t) me + 1] = G1; gradrho[3 * (int64_t) me + 2] = G2;
    if(ninteract) ninteract[me] = nint;
    if(niter) niter[me] = it + 1;
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

// per-particle hydro inputs in curve order
__global__ void __launch_bounds__(256)
k_sph_gather_hydro(int np, const int *__restrict__ sidx, SphDev S, const double *__restrict__ hsml, const double *__restrict__ density,
                   const double *__restrict__ egy, const double *__restrict__ dhsmlfac, const double *__restrict__ divvel,
                   const double *__restrict__ curlvel, const double4 *__restrict__ svel,
                   double4 *__restrict__ hA, double4 *__restrict__ hB)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if(j >= np) return;
    const int64_t i = sidx[j];
    const int DI = S.p.DensityIndependentSphOn;
    const double dens = density[i], dvv = divvel[i];
    // SPH_DensityPred hydra.c:300-312
    double dj = dens - dvv * dens * S.p.drift; if(!(dj >= 1e-6 * dens)) dj = 1e-6 * dens;
    const double eom0 = DI ? egy[i] : dens;
    double eom = eom0 - dvv * eom0 * S.p.drift; if(!(eom >= 1e-6 * eom0)) eom = 1e-6 * eom0;
    const double ev = svel[j].w;
    double P = 0;                                   // PressurePred hydra.c:67-77, cache hydra.c:195-214
    if(ev != 0 && ev * eom > 0) P = exp(GAMMA * log(ev * eom));
    hA[j] = make_double4(hsml[i], dj, eom, P);
    hB[j] = make_double4(dvv, curlvel[i], dhsmlfac[i], eom0);
}

__global__ void __launch_bounds__(128)
k_sph_hydro(int np, const int *__restrict__ sidx, const double4 *__restrict__ nodeB, const int4 *__restrict__ nodeC,
            const double *__restrict__ nodeH, const double4 *__restrict__ spart, const double4 *__restrict__ svel,
            const double4 *__restrict__ hA, const double4 *__restrict__ hB, const double *__restrict__ density,
            const uint8_t *__restrict__ type, SphDev S,
            double *__restrict__ acc_out, double *__restrict__ dte_out, double *__restrict__ maxsig_out, int *__restrict__ ninteract)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if(j >= np) return;
    const int me = sidx[j];
    if(type[me] != 0) return;                      // hydro_haswork hydra.c:508-512
    const int DI = S.p.DensityIndependentSphOn;
    const double4 pm = spart[j], vm = svel[j], a_i = hA[j], b_i = hB[j];
    const double h_i = a_i.x, P_i = a_i.w, eom_i = b_i.w, dens_i = density[me];
    // hydro_copy hydra.c:247-277
    const double cs_i = sqrt(GAMMA * P_i / eom_i);
    const double F1 = fabs(b_i.x) / (fabs(b_i.x) + b_i.y + 0.0001 * cs_i / h_i / S.fac_mu);
    const double p_over_rho2_i = P_i / (eom_i * eom_i);
    Kern ki; kern_init(ki, h_i, S);
    double A0 = 0, A1 = 0, A2 = 0, DtE = 0, MaxSig = cs_i;
    int ncand = 0;
    int no = 0;
    while(no >= 0) {                                // ngb_treefind_threads treewalk.c:1056-1143 (symmetric)
        const double4 B = nodeB[no];
        const int4 C = nodeC[no];
        if(!cull_keep(B, nodeH[no], pm.x, pm.y, pm.z, h_i, true, S)) { no = C.x; continue; }
        if(!C.w) { no = no + 1; continue; }
        ncand += C.z;
        for(int c = 0; c < C.z; c++) {              // treewalk.c:962-999, hydro_ngbiter hydra.c:350-505
            const int o = C.y + c;
            const double4 q = spart[o], a_j = hA[o];
            const double hm = a_j.x > h_i ? a_j.x : h_i, h2 = hm * hm;
            const double d0 = nearest_s(pm.x - q.x, S.box, S.halfbox);
            double rsq = d0 * d0;
            if(rsq > h2) continue;
            const double d1 = nearest_s(pm.y - q.y, S.box, S.halfbox);
            rsq += d1 * d1;
            if(rsq > h2) continue;
            const double d2 = nearest_s(pm.z - q.z, S.box, S.halfbox);
            rsq += d2 * d2;
            if(rsq > h2) continue;
            Kern kj; kern_init(kj, a_j.x, S);
            if(rsq <= 0 || !(rsq < ki.HH || rsq < kj.HH)) continue;
            const double r = sqrt(rsq);
            const double4 vo = svel[o], b_j = hB[o];
            const double density_j = a_j.y, eom_j = a_j.z, P_j = a_j.w;
            const double p_over_rho2_j = P_j / (eom_j * eom_j);
            const double cs_j = sqrt(GAMMA * P_j / eom_j);
            double vsig = cs_i + cs_j;
            if(vsig > MaxSig) MaxSig = vsig;
            const double v0 = vm.x - vo.x, v1 = vm.y - vo.y, v2 = vm.z - vo.z;
            const double vdotr = d0 * v0 + d1 * v1 + d2 * v2;
            const double vdotr2 = vdotr + S.hubble_a2 * rsq;
            const double dwk_i = kern_dw(ki, r * ki.Hinv, S);
            const double dwk_j = kern_dw(kj, r * kj.Hinv, S);
            double visc = 0;
            if(vdotr2 < 0) {
                const double mu_ij = S.fac_mu * vdotr2 / r;
                const double rho_ij = 0.5 * (dens_i + density_j);
                double vs = cs_i + cs_j;
                vs -= 3 * mu_ij;
                if(vs > MaxSig) MaxSig = vs;
                const double f2 = fabs(b_j.x) / (fabs(b_j.x) + b_j.y + 0.0001 * cs_j / S.fac_mu / a_j.x);
                visc = 0.25 * S.p.ArtBulkViscConst * vs * (-mu_ij) / rho_ij * (F1 + f2);
                const double dloga = 2 * S.p.dloga_bin;
                if(dloga > 0 && (dwk_i + dwk_j) < 0) {
                    const double msum = pm.w + q.w;
                    if(msum > 0) {
                        const double lim = 0.5 * S.fac_vsic_fix * vdotr2 / (0.5 * msum * (dwk_i + dwk_j) * r * dloga);
                        if(lim < visc) visc = lim;
                    }
                }
            }
            const double hfc_visc = 0.5 * q.w * visc * (dwk_i + dwk_j) / r;
            double hfc = hfc_visc, rr1 = 1, rr2 = 1;
            if(DI) {
                rr1 = 0; rr2 = 0;
                hfc += q.w * (dwk_i * p_over_rho2_i * vo.w / vm.w + dwk_j * p_over_rho2_j * vm.w / vo.w) / r;
                if(S.p.DensityContrastLimit >= 0) {
                    rr1 = eom_i / dens_i;
                    rr2 = eom_j / density_j;
                    if(S.p.DensityContrastLimit > 0) {
                        if(S.p.DensityContrastLimit < rr1) rr1 = S.p.DensityContrastLimit;
                        if(S.p.DensityContrastLimit < rr2) rr2 = S.p.DensityContrastLimit;
                    }
                }
            }
            hfc += q.w * (p_over_rho2_i * b_i.z * dwk_i * rr1 + p_over_rho2_j * b_j.z * dwk_j * rr2) / r;
            A0 += (-hfc * d0); A1 += (-hfc * d1); A2 += (-hfc * d2);
            DtE += (0.5 * hfc_visc * vdotr2);
        }
        no = C.x;
    }
    // hydro_postprocess hydra.c:514-528
    DtE *= GAMMA_MINUS1 / (S.hubble_a2 * pow(dens_i, GAMMA_MINUS1));
    acc_out[3 * (int64_t) me] = A0; acc_out[3 * (int64_t) me + 1] = A1; acc_out[3 * (int64_t) me + 2] = A2;
    dte_out[me] = DtE;
    maxsig_out[me] = MaxSig;
    if(ninteract) ninteract[me] = ncand;
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
        CK(cudaMemcpyAsync(it.dst->p, it.src, it.k * E->n * sizeof(double), cudaMemcpyHostToDevice, E->stream));
    }
    if(!hsml) return failmsg(E, "b200_sph_set_gas: hsml is required");
    CK(cudaStreamSynchronize(E->stream));
    E->sph_density_done = false;
    return 0;
}

int sph_density(Engine *E, const b200_sph_params *p, int update_hsml, int DoEgy, int *d_ninteract, int *d_niter)
{
    if(!E->tree_valid) return failmsg(E, "b200_density: build the gas tree first (b200_tree_build with the gas mask)");
    if(!E->s_have[1]) return failmsg(E, "b200_density: call b200_sph_set_gas first");
    SphDev S;
    if(int rc = make_dev(E, p, S)) return rc;
    const size_t n = (size_t) (E->n > 0 ? E->n : 1);
    const int np = (int) E->tree_np, nn = (int) E->tree_nn;
    CK(E->s_velpred.ensure(3 * n)); CK(E->s_evp.ensure(n));
    CK(E->s_density.ensure(n)); CK(E->s_egy.ensure(n)); CK(E->s_dhsmlfac.ensure(n)); CK(E->s_divvel.ensure(n));
    CK(E->s_curlvel.ensure(n)); CK(E->s_dthsml.ensure(n)); CK(E->s_numngb.ensure(n)); CK(E->s_gradrho.ensure(3 * n));
    CK(E->s_svel.ensure(4 * (size_t) (np > 0 ? np : 1)));
    CK(E->scratch_i.ensure(16));
    CK(cudaMemsetAsync(E->scratch_i.p + 12, 0, sizeof(int), E->stream));
    timer_start(E, T_SPH_DENSITY);
    if(E->n > 0) {
        k_sph_predict<<<(unsigned) ((E->n + 255) / 256), 256, 0, E->stream>>>(E->n, E->s_have[0] ? E->s_vel.p : nullptr,
            E->s_have[4] ? E->s_fullacc.p : nullptr, E->s_have[5] ? E->s_gravpm.p : nullptr, E->s_have[6] ? E->s_hydroacc.p : nullptr,
            E->s_have[2] ? E->s_entropy.p : nullptr, E->s_have[3] ? E->s_dtentropy.p : nullptr, S, E->s_velpred.p, E->s_evp.p);
        CKL(E);
    }
    if(np > 0) {
        k_sph_gather_vel<<<(np + 255) / 256, 256, 0, E->stream>>>(np, E->sidx.p, E->s_velpred.p, E->s_evp.p, (double4 *) E->s_svel.p);
        CKL(E);
        k_sph_density<<<(np + 127) / 128, 128, 0, E->stream>>>(np, E->sidx.p, (const double4 *) E->nodeB.p, (const int4 *) E->nodeC.p,
            (const double4 *) E->spart.p, (const double4 *) E->s_svel.p, E->type.p, S, update_hsml, DoEgy,
            E->s_hsml.p, E->s_density.p, E->s_egy.p, E->s_dhsmlfac.p, E->s_divvel.p, E->s_curlvel.p, E->s_dthsml.p, E->s_numngb.p,
            E->s_gradrho.p, d_ninteract, d_niter, E->scratch_i.p + 12);
        CKL(E);
        // hmax of the tree from the converged smoothing lengths (run.c:477 force_tree_calc_moments)
        k_sph_hmax_leaf<<<(nn + 255) / 256, 256, 0, E->stream>>>(nn, (const double4 *) E->nodeB.p, (const int4 *) E->nodeC.p,
            (const double4 *) E->spart.p, E->sidx.p, E->s_hsml.p, E->type.p, E->nodeH.p);
        CKL(E);
        for(int level = (int) E->tree_lvl.size() - 2; level >= 0; level--) {
            const int first = E->tree_lvl[level], last = E->tree_lvl[level + 1];
            k_sph_hmax_up<<<(last - first + 127) / 128, 128, 0, E->stream>>>(first, last, E->b_firstchild.p, E->b_nchild.p, E->b_dfs.p, E->nodeH.p);
            CKL(E);
        }
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
    CK(E->s_hA.ensure(4 * (size_t) (np > 0 ? np : 1))); CK(E->s_hB.ensure(4 * (size_t) (np > 0 ? np : 1)));
    timer_start(E, T_SPH_HYDRO);
    if(np > 0) {
        k_sph_gather_hydro<<<(np + 255) / 256, 256, 0, E->stream>>>(np, E->sidx.p, S, E->s_hsml.p, E->s_density.p, E->s_egy.p, E->s_dhsmlfac.p,
            E->s_divvel.p, E->s_curlvel.p, (const double4 *) E->s_svel.p, (double4 *) E->s_hA.p, (double4 *) E->s_hB.p);
        CKL(E);
        k_sph_hydro<<<(np + 127) / 128, 128, 0, E->stream>>>(np, E->sidx.p, (const double4 *) E->nodeB.p, (const int4 *) E->nodeC.p, E->nodeH.p,
            (const double4 *) E->spart.p, (const double4 *) E->s_svel.p, (const double4 *) E->s_hA.p, (const double4 *) E->s_hB.p,
            E->s_density.p, E->type.p, S, d_acc, d_dte, d_maxsig, d_ninteract);
        CKL(E);
    }
    timer_stop(E, T_SPH_HYDRO);
    return 0;
}

} // namespace b200
