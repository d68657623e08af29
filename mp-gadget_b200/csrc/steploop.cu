// steploop.cu -- the step loop around the force computation, device-resident:
// drift, active-particle lists, half kicks, gravity time-bin assignment and the hierarchical gravity
// drivers of the reference's timestep.c, so that positions, velocities and time bins stay in HBM
// between force computations instead of making the 160-byte-per-particle round trip every sub-step.
//
// Reference: libgadget/drift.c:17-102, libgadget/timestep.c:143-150 (is_timebin_active), :215-279
// (kick times, hierarchical kick), :296-598 (hierarchical drivers), :860-994 (half kicks),
// :1046-1073 (gravity time step), :1155-1173, :1201-1315 (PM step, bins), :1334-1478 (active lists),
// libgadget/timebinmgr.c:360-462 (integer timeline).
//
// State: Engine::pos / type / flags / mass plus the per-particle arrays the SPH module already keeps
// in original index order (s_vel, s_fullacc, s_gravpm, s_hydroacc, s_hsml, s_dthsml, s_entropy,
// s_dtentropy, s_bin_grav, s_bin_hydro).  All kernels are one thread per particle (or list entry)
// and HBM-bound; they are compiled without FMA contraction so that the arithmetic matches the CPU
// statement by statement.
#include <cub/cub.cuh>
#include <math.h>
#include <string.h>
#include "engine.h"

namespace b200 {

#define TB B200_TIMEBINS
#define NBIN (B200_TIMEBINS + 1)
#ifndef STEP_BLOCKS
#define STEP_BLOCKS 592          // 4 x 148 SMs: fixed grid of the per-type reductions (deterministic partial sums)
#endif

__device__ __forceinline__ long long dti_of_bin(int bin) { return bin > 0 ? (1ll << bin) : 0ll; }
__device__ __forceinline__ bool bin_active(int bin, long long ti)          // timestep.c:143-150
{
    if(bin <= 0 || ti <= 0) return true;
    return (ti & (dti_of_bin(bin) - 1)) == 0;                              // ti % 2^bin
}

// ---- drift.c:17-102 ----
__global__ void __launch_bounds__(256)
k_step_drift(int64_t n, double *__restrict__ pos, const double *__restrict__ vel, const uint8_t *__restrict__ type,
             const uint8_t *__restrict__ flags, double *__restrict__ hsml, const double *__restrict__ dthsml,
             double ddrift, double sx, double sy, double sz, double Box, unsigned long long *__restrict__ nbad)
{
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    const double sh[3] = {sx, sy, sz};
    double x[3] = {pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]};
    const bool dead = (flags[i] & 3) != 0;
    if(!dead) {
        if(type[i] == 0 && hsml) {                                         // :54-67
            double h = hsml[i] + dthsml[i] * ddrift;
            if(h <= 0) atomicAdd(nbad, 1ull);
            if(h > Box / 2.) h = Box / 2.;
            hsml[i] = h;
        }
#pragma unroll
        for(int j = 0; j < 3; j++) {                                       // :68-74
            x[j] += vel[3 * i + j] * ddrift + sh[j];
            if(!isfinite(x[j])) { atomicAdd(nbad, 1ull); x[j] = Box; }
        }
    } else {
#pragma unroll
        for(int j = 0; j < 3; j++) x[j] += sh[j];                          // :20-29
    }
#pragma unroll
    for(int j = 0; j < 3; j++) {                                           // :75-78
        while(x[j] > Box) x[j] -= Box;
        while(x[j] <= 0) x[j] += Box;
        pos[3 * i + j] = x[j];
    }
}

// black holes carry drag / dynamic-friction kicks and repositioning (timestep.c:1006-1012, drift.c:32-53) that this module
// does not have: count them so that b200_step_set_state can refuse instead of integrating them wrongly
__global__ void __launch_bounds__(256)
k_step_count_bh(int64_t n, const uint8_t *__restrict__ type, const uint8_t *__restrict__ flags, unsigned long long *__restrict__ cnt)
{
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if(i < n && type[i] == 5 && !(flags[i] & 3)) atomicAdd(cnt, 1ull);
}

// ---- active lists: timestep.c:1334-1478 ----
__global__ void k_step_iota(int *p, int64_t n)
{
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if(i < n) p[i] = (int) i;
}
// flag[i] = on the active list; cnt[0] gravitationally active, cnt[1 + type * NBIN + bin] = TimeBinCountType
__global__ void __launch_bounds__(256)
k_step_active_flags(int64_t n, const uint8_t *__restrict__ type, const uint8_t *__restrict__ flags, const uint8_t *__restrict__ bin_grav,
                    const uint8_t *__restrict__ bin_hydro, long long Ti, int is_pm, uint8_t *__restrict__ flag, unsigned long long *__restrict__ cnt)
{
    __shared__ unsigned int s_cnt[1 + 6 * NBIN];
    for(int k = threadIdx.x; k < 1 + 6 * NBIN; k += blockDim.x) s_cnt[k] = 0;
    __syncthreads();
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if(i < n) {
        uint8_t f = 0;
        if(!(flags[i] & 3)) {
            const int ty = type[i] < 6 ? type[i] : 5;
            // gas and black holes carry a hydro bin.  Taken from a bit table, not from `ty == 0 || ty == 5`: for that form
            // ptxas 12.9 (sm_100a) emits VIMNMX.U16x2 R, P4, P4, type, 0x5 -- one predicate for both 16-bit halves -- and the
            // hardware then reports every type as a hydro particle (first hardware run of this kernel, round 2).
            const bool hydro_particle = (0x21u >> ty) & 1u;
            const int bg = bin_grav[i], bh = bin_hydro[i];
            const int b = hydro_particle ? bh : bg;
            atomicAdd(&s_cnt[1 + ty * NBIN + (b < NBIN ? b : TB)], 1u);
            if(!is_pm) {
                const bool ga = bin_active(bg, Ti);
                const bool ha = hydro_particle && bin_active(bh, Ti);
                if(ga) atomicAdd(&s_cnt[0], 1u);
                f = (ga || ha) ? 1 : 0;
            }
        }
        flag[i] = f;
    }
    __syncthreads();
    for(int k = threadIdx.x; k < 1 + 6 * NBIN; k += blockDim.x)
        if(s_cnt[k]) atomicAdd(&cnt[k], (unsigned long long) s_cnt[k]);
}
// build_active_sublist :1453-1467; list == nullptr: the identity list
__global__ void __launch_bounds__(256)
k_step_sublist_flags(int64_t nlist, const int *__restrict__ list, const uint8_t *__restrict__ flags, const uint8_t *__restrict__ bin_grav,
                     int maxtimebin, long long Ti, uint8_t *__restrict__ flag)
{
    const int64_t q = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if(q >= nlist) return;
    const int64_t i = list ? list[q] : q;
    const int b = bin_grav[i];
    flag[q] = (!(flags[i] & 3) && b <= maxtimebin && bin_active(b, Ti)) ? 1 : 0;
}

// ---- kicks ----
// apply_half_kick / apply_hydro_half_kick timestep.c:874-970 with do_hydro_kick :1003-1043 for gas.
// tab = [3][NBIN]: gravkick, hydrokick, dt_entr by bin.
__global__ void __launch_bounds__(256)
k_step_half_kick(int64_t nlist, const int *__restrict__ list, const uint8_t *__restrict__ type, const uint8_t *__restrict__ flags,
                 const uint8_t *__restrict__ bin_grav, const uint8_t *__restrict__ bin_hydro, double *__restrict__ vel,
                 const double *__restrict__ fullacc, const double *__restrict__ hydroacc, double *__restrict__ entropy,
                 const double *__restrict__ dtentropy, const double *__restrict__ tab, long long Ti, double atime, double MaxGasVel, int hydro_only)
{
    __shared__ double s_tab[3 * NBIN];
    for(int k = threadIdx.x; k < 3 * NBIN; k += blockDim.x) s_tab[k] = tab[k];
    __syncthreads();
    const int64_t q = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if(q >= nlist) return;
    const int64_t i = list ? list[q] : q;
    if(flags[i] & 3) return;
    double v[3] = {vel[3 * i], vel[3 * i + 1], vel[3 * i + 2]};
    bool touched = false;
    const int bg = bin_grav[i];
    if(!hydro_only && bin_active(bg, Ti)) {                                // :904-906
        const double F = s_tab[bg];
#pragma unroll
        for(int j = 0; j < 3; j++) v[j] += fullacc[3 * i + j] * F;
        touched = true;
    }
    if(type[i] == 0) {                                                     // :915-919, do_hydro_kick
        const int bh = bin_hydro[i];
        const double Fh = s_tab[NBIN + bh], dt_entr = s_tab[2 * NBIN + bh];
#pragma unroll
        for(int j = 0; j < 3; j++) v[j] += hydroacc[3 * i + j] * Fh;
        double vv = 0;
#pragma unroll
        for(int j = 0; j < 3; j++) vv += v[j] * v[j];
        vv = sqrt(vv);
        if(vv > 0 && vv / atime > MaxGasVel) {
#pragma unroll
            for(int j = 0; j < 3; j++) v[j] *= MaxGasVel * atime / vv;
        }
        entropy[i] += dtentropy[i] * dt_entr;
        touched = true;
    }
    if(touched) { vel[3 * i] = v[0]; vel[3 * i + 1] = v[1]; vel[3 * i + 2] = v[2]; }
}
// apply_PM_half_kick :980-990, and the particle loop of apply_hierarchical_grav_kick :258-276
__global__ void __launch_bounds__(256)
k_step_acc_kick(int64_t nlist, const int *__restrict__ list, const uint8_t *__restrict__ flags, double *__restrict__ vel,
                const double *__restrict__ acc, double F)
{
    const int64_t q = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if(q >= nlist) return;
    const int64_t i = list ? list[q] : q;
    if(flags[i] & 3) return;
#pragma unroll
    for(int j = 0; j < 3; j++) vel[3 * i + j] += acc[3 * i + j] * F;
}

// ---- gravity time bins ----
struct StepTimeline {          // what dti_from_dloga needs, timebinmgr.c:360-378,435-440
    const double *__restrict__ sync;   // SyncPoints[].loga
    int nsync;
    double now;                // loga_from_ti(Ti_Current)
    long long ti_now;          // ti_from_loga(now)
};
__device__ __forceinline__ long long dev_ti_from_loga(const StepTimeline &T, double loga)
{
    int i;
    for(i = 1; i < T.nsync - 1; i++)
        if(T.sync[i] > loga) break;
    const double step = (T.sync[i] - T.sync[i - 1]) / (double) (1ll << TB);
    const long long ti = ((long long) (i - 1)) << TB;
    return (long long) ((double) ti + (loga - T.sync[i - 1]) / step);
}
// get_timestep_gravity_dloga + grav_acceleration2 :1046-1073, convert_timestep_to_ti :1155-1173
__device__ __forceinline__ long long dev_gravity_dti(const StepTimeline &T, const double *acc, const double *gpm, double atime, double hubble,
                                                     double ErrTolIntAccuracy, double softening, double MinSizeTimestep, long long dti_max)
{
    const double a2inv = 1 / (atime * atime);
    double ax = a2inv * acc[0], ay = a2inv * acc[1], az = a2inv * acc[2];
    ay += a2inv * gpm[1]; ax += a2inv * gpm[0]; az += a2inv * gpm[2];
    double ac2 = ax * ax + ay * ay + az * az;
    if(ac2 == 0) ac2 = 1.0e-60;
    const double ac = sqrt(ac2);
    const double dt = sqrt(2 * ErrTolIntAccuracy * atime * (softening / 2.8) / ac);
    double dloga = dt * hubble;
    if(dti_max == 0) return 0;
    if(dloga < MinSizeTimestep) dloga = MinSizeTimestep;
    long long dti = dev_ti_from_loga(T, dloga + T.now) - T.ti_now;
    if(dti > dti_max || dti < 0) dti = dti_max;
    return dti;
}
// top of the hierarchy, timestep.c:346-372: bin from the stored acceleration, clamped to largest_active
__global__ void __launch_bounds__(256)
k_step_assign_bins(int64_t nlist, const int *__restrict__ list, const uint8_t *__restrict__ flags, const double *__restrict__ acc,
                   const double *__restrict__ gravpm, StepTimeline T, double atime, double hubble, double ErrTolIntAccuracy, double softening,
                   double MinSizeTimestep, long long dti_max, int largest_active, uint8_t *__restrict__ bin_grav, unsigned long long *__restrict__ cnt)
{
    __shared__ unsigned int s_cnt[NBIN];
    for(int k = threadIdx.x; k < NBIN; k += blockDim.x) s_cnt[k] = 0;
    __syncthreads();
    const int64_t q = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if(q < nlist) {
        const int64_t i = list ? list[q] : q;
        if(!(flags[i] & 3)) {
            long long dti = dev_gravity_dti(T, acc + 3 * i, gravpm + 3 * i, atime, hubble, ErrTolIntAccuracy, softening, MinSizeTimestep, dti_max);
            // round_down_power_of_two + get_timestep_bin: the position of the highest set bit
            int bin = 0;
            if(dti > 1) {
                if(dti > (1ll << TB)) dti = 1ll << TB;
                bin = 63 - __clzll(dti);
            }
            if(bin > largest_active) bin = largest_active;
            bin_grav[i] = (uint8_t) bin;
            atomicAdd(&s_cnt[bin], 1u);
        }
    }
    __syncthreads();
    for(int k = threadIdx.x; k < NBIN; k += blockDim.x)
        if(s_cnt[k]) atomicAdd(&cnt[k], (unsigned long long) s_cnt[k]);
}
// push-down :405-412
__global__ void __launch_bounds__(256)
k_step_push_down(int64_t nlist, const int *__restrict__ list, int push_down, uint8_t *__restrict__ bin_grav)
{
    const int64_t q = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if(q >= nlist) return;
    const int64_t i = list ? list[q] : q;
    if(bin_grav[i] > push_down) bin_grav[i] = (uint8_t) push_down;
}
// lower levels :457-472: one bin down when this level's acceleration asks for it
__global__ void __launch_bounds__(256)
k_step_reduce_bins(int64_t nlist, const int *__restrict__ list, const uint8_t *__restrict__ flags, const double *__restrict__ acc,
                   const double *__restrict__ gravpm, StepTimeline T, double atime, double hubble, double ErrTolIntAccuracy, double softening,
                   double MinSizeTimestep, long long dti_max, int ti, uint8_t *__restrict__ bin_grav, unsigned long long *__restrict__ nbad)
{
    const int64_t q = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if(q >= nlist) return;
    const int64_t i = list ? list[q] : q;
    if(flags[i] & 3) return;
    const long long dti = dev_gravity_dti(T, acc + 3 * i, gravpm + 3 * i, atime, hubble, ErrTolIntAccuracy, softening, MinSizeTimestep, dti_max);
    if(dti < dti_of_bin(ti)) {
        bin_grav[i] = (uint8_t) (ti - 1);
        if(ti == 1) atomicAdd(nbad, 1ull);
    }
}
// find_hydro_timesteps timestep.c:617-693 for gas: Courant and smoothing-length criteria
// (get_timestep_hydro_dloga :1075-1117), get_timebin_from_dti :166-182, clamp to the gravity bin.
// cnt[0] bad steps, cnt[1] smallest new bin (atomicMin).
__global__ void __launch_bounds__(256)
k_step_hydro_bins(int64_t nlist, const int *__restrict__ list, const uint8_t *__restrict__ type, const uint8_t *__restrict__ flags,
                  const double *__restrict__ hsml, const double *__restrict__ dthsml, const double *__restrict__ maxsig,
                  const uint8_t *__restrict__ bin_grav, uint8_t *__restrict__ bin_hydro, StepTimeline T, double atime, double hubble, double fac3,
                  double CourantFac, double MinSizeTimestep, long long dti_max, long long Ti, unsigned long long *__restrict__ cnt)
{
    const int64_t q = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if(q >= nlist) return;
    const int64_t i = list ? list[q] : q;
    if((flags[i] & 3) || type[i] != 0) return;
    double dt = 2 * CourantFac * atime * hsml[i] / (fac3 * maxsig[i]);
    const double dt_hsml = CourantFac * atime * atime * fabs(hsml[i] / (dthsml[i] + 1e-20));
    if(dt_hsml < dt) dt = dt_hsml;
    double dloga = dt * hubble;
    long long dti = 0;                                                     // convert_timestep_to_ti :1155-1173
    if(dti_max != 0) {
        if(dloga < MinSizeTimestep) dloga = MinSizeTimestep;
        dti = dev_ti_from_loga(T, dloga + T.now) - T.ti_now;
        if(dti > dti_max || dti < 0) dti = dti_max;
    }
    int bin = 0;
    if(dti > 1) {
        if(dti > (1ll << TB)) dti = 1ll << TB;
        bin = 63 - __clzll(dti);
    }
    const int binold = bin_hydro[i];
    if(bin > binold)
        while(!bin_active(bin, Ti) && bin > binold && bin > 1) bin--;
    if(bin > bin_grav[i]) bin = bin_grav[i];
    if(bin < 1) atomicAdd(&cnt[0], 1ull);
    if(bin_active(binold, Ti) && bin_active(bin, Ti)) bin_hydro[i] = (uint8_t) bin;
    atomicMin(&cnt[1], (unsigned long long) bin);
}
// find_timesteps timestep.c:739-853 (SplitGravityTimestepsOn = 0): the smaller of the gravity and (gas) hydro steps -> one
// bin for TimeBinGravity and TimeBinHydro.  cnt[0] bad steps, cnt[1] smallest, cnt[2] largest new bin.
__global__ void __launch_bounds__(256)
k_step_find_bins(int64_t nlist, const int *__restrict__ list, const uint8_t *__restrict__ type, const uint8_t *__restrict__ flags,
                 const double *__restrict__ fullacc, const double *__restrict__ gravpm, const double *__restrict__ hsml,
                 const double *__restrict__ dthsml, const double *__restrict__ maxsig, uint8_t *__restrict__ bin_grav, uint8_t *__restrict__ bin_hydro,
                 StepTimeline T, double atime, double hubble, double fac3, double ErrTolIntAccuracy, double softening, double CourantFac,
                 double MinSizeTimestep, long long dti_max, long long Ti, unsigned long long *__restrict__ cnt)
{
    const int64_t q = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if(q >= nlist) return;
    const int64_t i = list ? list[q] : q;
    if(flags[i] & 3) return;
    long long dti = dev_gravity_dti(T, fullacc + 3 * i, gravpm + 3 * i, atime, hubble, ErrTolIntAccuracy, softening, MinSizeTimestep, dti_max);
    if(type[i] == 0) {                                                     // :783-791
        double dt = 2 * CourantFac * atime * hsml[i] / (fac3 * maxsig[i]);
        const double dt_hsml = CourantFac * atime * atime * fabs(hsml[i] / (dthsml[i] + 1e-20));
        if(dt_hsml < dt) dt = dt_hsml;
        double dloga = dt * hubble;
        long long dti_hydro = 0;
        if(dti_max != 0) {
            if(dloga < MinSizeTimestep) dloga = MinSizeTimestep;
            dti_hydro = dev_ti_from_loga(T, dloga + T.now) - T.ti_now;
            if(dti_hydro > dti_max || dti_hydro < 0) dti_hydro = dti_max;
        }
        if(dti_hydro < dti) dti = dti_hydro;
    }
    int bin = 0;                                                           // get_timebin_from_dti :166-182
    if(dti > 1) {
        if(dti > (1ll << TB)) dti = 1ll << TB;
        bin = 63 - __clzll(dti);
    }
    const int binold = bin_hydro[i];
    if(bin > binold)
        while(!bin_active(bin, Ti) && bin > binold && bin > 1) bin--;
    if(bin < 1) atomicAdd(&cnt[0], 1ull);
    if(bin_active(binold, Ti) && bin_active(bin, Ti)) { bin_hydro[i] = (uint8_t) bin; bin_grav[i] = (uint8_t) bin; }
    atomicMin(&cnt[1], (unsigned long long) bin);
    atomicMax(&cnt[2], (unsigned long long) bin);
}
// |FullTreeGravAccel + GravPM| for the relative opening criterion, gravshort.h:69-86
__global__ void __launch_bounds__(256)
k_step_oldacc(int64_t n, const double *__restrict__ fullacc, const double *__restrict__ gravpm, double *__restrict__ oldacc)
{
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    double s = 0;
#pragma unroll
    for(int j = 0; j < 3; j++) { const double a = fullacc[3 * i + j] + gravpm[3 * i + j]; s += a * a; }
    oldacc[i] = sqrt(s);
}
// get_long_range_timestep_dloga :1215-1229: per type sum of v^2, particle count and smallest mass.
// Fixed grid, block partials in block order -> the host finishes the sum in a fixed order.
__global__ void __launch_bounds__(256)
k_step_vel_stats(int64_t n, const double *__restrict__ vel, const float *__restrict__ mass, const uint8_t *__restrict__ type,
                 const uint8_t *__restrict__ flags, double *__restrict__ part)      // part[block][18]: v2[6], count[6], minmass[6]
{
    __shared__ double s_v[6][8], s_c[6][8], s_m[6][8];
    double v[6] = {0, 0, 0, 0, 0, 0}, c[6] = {0, 0, 0, 0, 0, 0}, m[6] = {1e30, 1e30, 1e30, 1e30, 1e30, 1e30};
    for(int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t) gridDim.x * blockDim.x) {
        if(flags[i] & 3) continue;
        const int ty = type[i] < 6 ? type[i] : 5;
        const double v2 = vel[3 * i] * vel[3 * i] + vel[3 * i + 1] * vel[3 * i + 1] + vel[3 * i + 2] * vel[3 * i + 2];
        const double mm = mass[i];
#pragma unroll
        for(int k = 0; k < 6; k++)
            if(k == ty) { v[k] += v2; c[k] += 1; if(mm > 0 && m[k] > mm) m[k] = mm; }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for(int k = 0; k < 6; k++) {
#pragma unroll
        for(int o = 16; o > 0; o >>= 1) {
            v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
            c[k] += __shfl_xor_sync(0xffffffffu, c[k], o);
            m[k] = fmin(m[k], __shfl_xor_sync(0xffffffffu, m[k], o));
        }
        if(lane == 0) { s_v[k][warp] = v[k]; s_c[k][warp] = c[k]; s_m[k][warp] = m[k]; }
    }
    __syncthreads();
    if(threadIdx.x < 6) {
        const int k = threadIdx.x;
        double sv = 0, sc = 0, sm = 1e30;
        for(int w = 0; w < 8; w++) { sv += s_v[k][w]; sc += s_c[k][w]; sm = fmin(sm, s_m[k][w]); }
        part[blockIdx.x * 18 + k] = sv; part[blockIdx.x * 18 + 6 + k] = sc; part[blockIdx.x * 18 + 12 + k] = sm;
    }
}

// ---- hand-over to / from the SPH module (sph.cu) without a host round trip ----
__global__ void __launch_bounds__(256)
k_step_mark_active(int64_t nlist, const int *__restrict__ list, uint8_t *__restrict__ active)
{
    const int64_t q = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if(q < nlist) active[list[q]] = 1;
}
// hydro_reduce / hydro_postprocess results (hydra.c:279-293,495-528) of the listed gas particles into the step state
__global__ void __launch_bounds__(256)
k_step_adopt_hydro(int64_t nlist, const int *__restrict__ list, const uint8_t *__restrict__ type, const uint8_t *__restrict__ flags,
                   const double *__restrict__ acc, const double *__restrict__ dte, const double *__restrict__ maxsig,
                   double *__restrict__ hydroacc, double *__restrict__ dtentropy, double *__restrict__ maxsig_out)
{
    const int64_t q = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if(q >= nlist) return;
    const int64_t i = list ? list[q] : q;
    if((flags[i] & 3) || type[i] != 0) return;
#pragma unroll
    for(int j = 0; j < 3; j++) hydroacc[3 * i + j] = acc[3 * i + j];
    dtentropy[i] = dte[i];
    maxsig_out[i] = maxsig[i];
}

// ---------------------------------------------------------------------------------------------
static inline unsigned grid_for(int64_t n) { return (unsigned) ((n + 255) / 256); }
static inline int64_t host_dti_of_bin(int bin) { return bin > 0 ? ((int64_t) 1 << bin) : 0; }
static inline bool host_bin_active(int bin, int64_t ti) { return bin <= 0 || ti <= 0 || ti % host_dti_of_bin(bin) == 0; }

static inline double step_box(const Engine *E) { return E->st_box > 0 ? E->st_box : (E->Box > 0 ? E->Box : E->tree_box); }

static int step_need_state(Engine *E, const char *who)
{
    if(!E->st_state) return failmsg(E, std::string(who) + ": call b200_step_set_state first");
    return 0;
}

// order-preserving compaction of `in` (nullptr: identity) by E->st_flag into out; returns the count
static int step_compact(Engine *E, const int *in, int64_t nin, int *out, int64_t *nout)
{
    *nout = 0;
    if(nin == 0) return 0;
    if(!in) {
        CK(E->st_iota.ensure((size_t) nin));
        k_step_iota<<<grid_for(nin), 256, 0, E->stream>>>(E->st_iota.p, nin); CKL(E);
        in = E->st_iota.p;
    }
    CK(E->scratch_i.ensure(256));
    int *d_num = E->scratch_i.p + 20;
    size_t tb = 0;
    cub::DeviceSelect::Flagged(nullptr, tb, in, E->st_flag.p, out, d_num, (int) nin, E->stream);
    CK(E->cubtemp.ensure(tb + 16));
    CK(cub::DeviceSelect::Flagged(E->cubtemp.p, tb, in, E->st_flag.p, out, d_num, (int) nin, E->stream));
    E->launches += 1;
    int cnt = 0;
    CK(cudaMemcpyAsync(&cnt, d_num, sizeof(int), cudaMemcpyDeviceToHost, E->stream));
    CK(cudaStreamSynchronize(E->stream));
    *nout = cnt;
    return 0;
}

int step_set_state(Engine *E, const b200_step_state *s)
{
    if(!s) return failmsg(E, "b200_step_set_state: null state");
    const size_t n = (size_t) (E->n > 0 ? E->n : 1);
    struct { const double *src; DevBuf<double> *dst; size_t k; int have; } items[] = {
        {s->vel, &E->s_vel, 3, 0}, {s->hsml, &E->s_hsml, 1, 1}, {s->entropy, &E->s_entropy, 1, 2}, {s->dtentropy, &E->s_dtentropy, 1, 3},
        {s->fullacc, &E->s_fullacc, 3, 4}, {s->gravpm, &E->s_gravpm, 3, 5}, {s->hydroacc, &E->s_hydroacc, 3, 6}, {s->dthsml, &E->s_dthsml, 1, -1}};
    for(auto &it : items) {
        CK(it.dst->ensure(it.k * n));
        if(it.src) CK(cudaMemcpyAsync(it.dst->p, it.src, it.k * E->n * sizeof(double), cudaMemcpyHostToDevice, E->stream));
        else CK(cudaMemsetAsync(it.dst->p, 0, it.k * n * sizeof(double), E->stream));
        if(it.have >= 0) E->s_have[it.have] = it.src != nullptr;      // the SPH module's "NULL = default" rules stay in force
    }
    CK(E->s_bin_grav.ensure(n)); CK(E->s_bin_hydro.ensure(n));
    if(s->bin_grav) CK(cudaMemcpyAsync(E->s_bin_grav.p, s->bin_grav, E->n, cudaMemcpyHostToDevice, E->stream));
    else CK(cudaMemsetAsync(E->s_bin_grav.p, 0, n, E->stream));
    if(s->bin_hydro) CK(cudaMemcpyAsync(E->s_bin_hydro.p, s->bin_hydro, E->n, cudaMemcpyHostToDevice, E->stream));
    else CK(cudaMemsetAsync(E->s_bin_hydro.p, 0, n, E->stream));
    if(s->flags) CK(cudaMemcpyAsync(E->flags.p, s->flags, E->n, cudaMemcpyHostToDevice, E->stream));
    CK(E->st_cnt.ensure(1 + 6 * NBIN));
    CK(cudaMemsetAsync(E->st_cnt.p, 0, sizeof(unsigned long long), E->stream));
    if(E->n > 0) { k_step_count_bh<<<(unsigned) ((E->n + 255) / 256), 256, 0, E->stream>>>(E->n, E->type.p, E->flags.p, E->st_cnt.p); CKL(E); }
    unsigned long long nbh = 0;
    CK(cudaMemcpyAsync(&nbh, E->st_cnt.p, sizeof(nbh), cudaMemcpyDeviceToHost, E->stream));
    CK(cudaStreamSynchronize(E->stream));
    if(nbh) { E->st_state = false; return failmsg(E, "b200_step_set_state: " + std::to_string(nbh) + " black-hole particles; their kicks and repositioning are not supported by the step loop"); }
    E->st_state = true;
    E->st_store_valid = false; E->st_maxsig_valid = false;
    if(s->BoxSize > 0) E->st_box = s->BoxSize;
    E->st_have_gas = s->hsml != nullptr;
    E->st_nact = E->n; E->st_act_implicit = true; E->st_nsub = 0;      // the active list does not survive a new state
    E->sph_density_done = false;
    return 0;
}

int step_get_state(Engine *E, b200_step_state_out *o)
{
    if(int rc = step_need_state(E, "b200_step_get_state")) return rc;
    if(!o) return failmsg(E, "b200_step_get_state: null output");
    const size_t n = (size_t) E->n;
    if(n == 0) return 0;
    if(o->pos) CK(cudaMemcpyAsync(o->pos, E->pos.p, 3 * n * sizeof(double), cudaMemcpyDeviceToHost, E->stream));
    if(o->vel) CK(cudaMemcpyAsync(o->vel, E->s_vel.p, 3 * n * sizeof(double), cudaMemcpyDeviceToHost, E->stream));
    if(o->fullacc) CK(cudaMemcpyAsync(o->fullacc, E->s_fullacc.p, 3 * n * sizeof(double), cudaMemcpyDeviceToHost, E->stream));
    if(o->hsml) CK(cudaMemcpyAsync(o->hsml, E->s_hsml.p, n * sizeof(double), cudaMemcpyDeviceToHost, E->stream));
    if(o->entropy) CK(cudaMemcpyAsync(o->entropy, E->s_entropy.p, n * sizeof(double), cudaMemcpyDeviceToHost, E->stream));
    if(o->bin_grav) CK(cudaMemcpyAsync(o->bin_grav, E->s_bin_grav.p, n, cudaMemcpyDeviceToHost, E->stream));
    if(o->bin_hydro) CK(cudaMemcpyAsync(o->bin_hydro, E->s_bin_hydro.p, n, cudaMemcpyDeviceToHost, E->stream));
    if(o->hydroacc) CK(cudaMemcpyAsync(o->hydroacc, E->s_hydroacc.p, 3 * n * sizeof(double), cudaMemcpyDeviceToHost, E->stream));
    if(o->dtentropy) CK(cudaMemcpyAsync(o->dtentropy, E->s_dtentropy.p, n * sizeof(double), cudaMemcpyDeviceToHost, E->stream));
    if(o->maxsignalvel) {
        if(!E->st_maxsig_valid) return failmsg(E, "b200_step_get_state: no MaxSignalVel on the device");
        CK(cudaMemcpyAsync(o->maxsignalvel, E->st_maxsig.p, n * sizeof(double), cudaMemcpyDeviceToHost, E->stream));
    }
    CK(cudaStreamSynchronize(E->stream));
    return 0;
}

int step_drift(Engine *E, double ddrift, const double *shift, int64_t *nbad)
{
    if(int rc = step_need_state(E, "b200_step_drift")) return rc;
    const double Box = step_box(E);
    if(Box <= 0) return failmsg(E, "b200_step_drift: box size unknown (b200_step_state.BoxSize, b200_pm_init or b200_tree_build)");
    if(nbad) *nbad = 0;
    if(E->n == 0) return 0;
    CK(E->st_cnt.ensure(1 + 6 * NBIN));
    CK(cudaMemsetAsync(E->st_cnt.p, 0, sizeof(unsigned long long), E->stream));
    k_step_drift<<<grid_for(E->n), 256, 0, E->stream>>>(E->n, E->pos.p, E->s_vel.p, E->type.p, E->flags.p,
        E->st_have_gas ? E->s_hsml.p : nullptr, E->s_dthsml.p, ddrift, shift ? shift[0] : 0.0, shift ? shift[1] : 0.0, shift ? shift[2] : 0.0,
        Box, E->st_cnt.p);
    CKL(E);
    unsigned long long bad = 0;
    CK(cudaMemcpyAsync(&bad, E->st_cnt.p, sizeof(bad), cudaMemcpyDeviceToHost, E->stream));
    CK(cudaStreamSynchronize(E->stream));
    if(nbad) *nbad = (int64_t) bad;
    E->tree_valid = false;               // positions moved
    E->potential_valid = false;
    E->sph_density_done = false;
    if(bad) return failmsg(E, "b200_step_drift: " + std::to_string(bad) + " particles with Hsml <= 0 or a non-finite position (drift.c:57-73)");
    return 0;
}

int step_build_active(Engine *E, int64_t Ti_Current, int is_pm, int64_t nhydro_slots, int64_t *counts, int64_t *bincounts)
{
    if(int rc = step_need_state(E, "b200_step_build_active")) return rc;
    const int64_t n = E->n;
    CK(E->st_cnt.ensure(1 + 6 * NBIN)); CK(E->st_flag.ensure((size_t) n + 64)); CK(E->st_act.ensure((size_t) n + 1));
    CK(cudaMemsetAsync(E->st_cnt.p, 0, (1 + 6 * NBIN) * sizeof(unsigned long long), E->stream));
    if(n > 0) {
        k_step_active_flags<<<grid_for(n), 256, 0, E->stream>>>(n, E->type.p, E->flags.p, E->s_bin_grav.p, E->s_bin_hydro.p,
            (long long) Ti_Current, is_pm, E->st_flag.p, E->st_cnt.p);
        CKL(E);
    }
    unsigned long long h[1 + 6 * NBIN];
    CK(cudaMemcpyAsync(h, E->st_cnt.p, sizeof(h), cudaMemcpyDeviceToHost, E->stream));
    CK(cudaStreamSynchronize(E->stream));
    if(bincounts) for(int k = 0; k < 6 * NBIN; k++) bincounts[k] = (int64_t) h[1 + k];
    E->st_nsub = 0;
    if(is_pm) {                                                            // :1341-1358: every particle, implicit list
        E->st_act_implicit = true; E->st_nact = n;
        if(counts) { counts[0] = n; counts[1] = n; counts[2] = nhydro_slots; }
        return 0;
    }
    int64_t na = 0;
    if(int rc = step_compact(E, nullptr, n, E->st_act.p, &na)) return rc;
    E->st_act_implicit = false; E->st_nact = na;
    if(counts) { counts[0] = na; counts[1] = (int64_t) h[0]; counts[2] = na; }     // nactivehydro counts every listed particle, :1392
    return 0;
}

// sublist of (list, nlist) into dst
static int step_sublist_of(Engine *E, const int *list, int64_t nlist, int maxtimebin, int64_t Ti_Current, int *dst, int64_t *nout)
{
    *nout = 0;
    if(nlist == 0) return 0;
    CK(E->st_flag.ensure((size_t) nlist + 64));
    k_step_sublist_flags<<<grid_for(nlist), 256, 0, E->stream>>>(nlist, list, E->flags.p, E->s_bin_grav.p, maxtimebin, (long long) Ti_Current, E->st_flag.p);
    CKL(E);
    return step_compact(E, list, nlist, dst, nout);
}

int step_active_sublist(Engine *E, int maxtimebin, int64_t Ti_Current, int64_t *nsub)
{
    if(int rc = step_need_state(E, "b200_step_active_sublist")) return rc;
    CK(E->st_listA.ensure((size_t) E->n + 1));
    int64_t ns = 0;
    if(int rc = step_sublist_of(E, E->st_act_implicit ? nullptr : E->st_act.p, E->st_nact, maxtimebin, Ti_Current, E->st_listA.p, &ns)) return rc;
    E->st_nsub = ns;
    if(nsub) *nsub = ns;
    return 0;
}

// The host's ActiveParticles list (ascending particle indices) -> device; list == nullptr: every particle.
int step_set_active(Engine *E, const int32_t *list, int64_t nlist)
{
    if(int rc = step_need_state(E, "b200_step_set_active")) return rc;
    E->st_nsub = 0;
    if(!list) { E->st_act_implicit = true; E->st_nact = E->n; return 0; }
    if(nlist < 0 || nlist > E->n) return failmsg(E, "b200_step_set_active: bad list length");
    CK(E->st_act.ensure((size_t) nlist + 1));
    if(nlist > 0) CK(cudaMemcpyAsync(E->st_act.p, list, nlist * sizeof(int32_t), cudaMemcpyHostToDevice, E->stream));
    CK(cudaStreamSynchronize(E->stream));
    E->st_act_implicit = false; E->st_nact = nlist;
    return 0;
}

// StoredGravAccel of the last b200_step_hier_accelerations, [n][3] (entries of particles it did not walk are 0)
int step_get_store(Engine *E, double *out)
{
    if(int rc = step_need_state(E, "b200_step_get_store")) return rc;
    if(!E->st_store_valid) return failmsg(E, "b200_step_get_store: no stored accelerations (run b200_step_hier_accelerations)");
    if(!out || E->n == 0) return 0;
    CK(cudaMemcpyAsync(out, E->st_store.p, 3 * (size_t) E->n * sizeof(double), cudaMemcpyDeviceToHost, E->stream));
    CK(cudaStreamSynchronize(E->stream));
    return 0;
}

// StoredGravAccel from the host (struct grav_accel_store, timestep.h:92-96) for b200_step_hier_timesteps
int step_set_store(Engine *E, const double *in)
{
    if(int rc = step_need_state(E, "b200_step_set_store")) return rc;
    if(!in) { E->st_store_valid = false; return 0; }
    const size_t n = (size_t) (E->n > 0 ? E->n : 1);
    CK(E->st_store.ensure(3 * n));
    CK(cudaMemcpyAsync(E->st_store.p, in, 3 * (size_t) E->n * sizeof(double), cudaMemcpyHostToDevice, E->stream));
    CK(cudaStreamSynchronize(E->stream));
    E->st_store_valid = true;
    return 0;
}

int step_get_active(Engine *E, int which, int32_t *out, int64_t *nout)
{
    if(int rc = step_need_state(E, "b200_step_get_active")) return rc;
    const int64_t n = which == 0 ? E->st_nact : E->st_nsub;
    if(nout) *nout = (which == 0 && E->st_act_implicit) ? -1 - n : n;
    if(!out || n == 0 || (which == 0 && E->st_act_implicit)) return 0;
    CK(cudaMemcpyAsync(out, which == 0 ? E->st_act.p : E->st_listA.p, n * sizeof(int32_t), cudaMemcpyDeviceToHost, E->stream));
    CK(cudaStreamSynchronize(E->stream));
    return 0;
}

int step_half_kick(Engine *E, const double *gravkick, const double *hydrokick, const double *dt_entr, int64_t Ti_Current, double atime,
                   double MaxGasVel, int hydro_only)
{
    if(int rc = step_need_state(E, "b200_step_half_kick")) return rc;
    if(!gravkick || !hydrokick || !dt_entr) return failmsg(E, "b200_step_half_kick: null factor table");
    const int64_t nl = E->st_nact;
    if(nl == 0) return 0;
    double h[3 * NBIN];
    memcpy(h, gravkick, NBIN * sizeof(double)); memcpy(h + NBIN, hydrokick, NBIN * sizeof(double)); memcpy(h + 2 * NBIN, dt_entr, NBIN * sizeof(double));
    CK(E->st_tab.ensure(3 * NBIN));
    CK(cudaMemcpyAsync(E->st_tab.p, h, sizeof(h), cudaMemcpyHostToDevice, E->stream));
    CK(cudaStreamSynchronize(E->stream));          // h is a stack array
    k_step_half_kick<<<grid_for(nl), 256, 0, E->stream>>>(nl, E->st_act_implicit ? nullptr : E->st_act.p, E->type.p, E->flags.p, E->s_bin_grav.p,
        E->s_bin_hydro.p, E->s_vel.p, E->s_fullacc.p, E->s_hydroacc.p, E->s_entropy.p, E->s_dtentropy.p, E->st_tab.p, (long long) Ti_Current,
        atime, MaxGasVel, hydro_only);
    CKL(E);
    CK(cudaStreamSynchronize(E->stream));
    return 0;
}

int step_pm_kick(Engine *E, double Fgravkick)
{
    if(int rc = step_need_state(E, "b200_step_pm_kick")) return rc;
    if(E->n == 0) return 0;
    k_step_acc_kick<<<grid_for(E->n), 256, 0, E->stream>>>(E->n, nullptr, E->flags.p, E->s_vel.p, E->s_gravpm.p, Fgravkick);
    CKL(E);
    CK(cudaStreamSynchronize(E->stream));
    return 0;
}

// FullTreeGravAccel / GravPM of the state <- the device results of the last full b200_grav_short_tree / b200_pm_force
int step_adopt_forces(Engine *E, int tree, int pm)
{
    if(int rc = step_need_state(E, "b200_step_adopt_forces")) return rc;
    const size_t bytes = 3 * (size_t) E->n * sizeof(double);
    if(tree) {
        if(!E->have_last_tree) return failmsg(E, "b200_step_adopt_forces: no full-tree accelerations on the device");
        CK(cudaMemcpyAsync(E->s_fullacc.p, E->last_tree_acc.p, bytes, cudaMemcpyDeviceToDevice, E->stream));
    }
    if(pm) {
        if(!E->have_last_pm) return failmsg(E, "b200_step_adopt_forces: no PM accelerations on the device");
        CK(cudaMemcpyAsync(E->s_gravpm.p, E->last_pm_acc.p, bytes, cudaMemcpyDeviceToDevice, E->stream));
    }
    CK(cudaStreamSynchronize(E->stream));
    return 0;
}

// ---- the integer timeline on the host (timebinmgr.c:360-462), for the scalars the drivers need ----
static double tl_interval(const b200_step_params *sp, int64_t ti)
{
    const int64_t lastsnap = ti >> TB;
    if(lastsnap >= sp->nsync - 1) return 0;
    return (sp->sync_loga[lastsnap + 1] - sp->sync_loga[lastsnap]) / (double) ((int64_t) 1 << TB);
}
static double tl_loga_from_ti(const b200_step_params *sp, int64_t ti)
{
    return sp->sync_loga[ti >> TB] + (ti & (((int64_t) 1 << TB) - 1)) * tl_interval(sp, ti);
}
static int64_t tl_ti_from_loga(const b200_step_params *sp, double loga)
{
    int64_t i;
    for(i = 1; i < sp->nsync - 1; i++)
        if(sp->sync_loga[i] > loga) break;
    const double step = (sp->sync_loga[i] - sp->sync_loga[i - 1]) / (double) ((int64_t) 1 << TB);
    const int64_t ti = (i - 1) << TB;
    return (int64_t) ((double) ti + (loga - sp->sync_loga[i - 1]) / step);
}
static int64_t tl_round_down_pow2(int64_t dti)
{
    int64_t t = (int64_t) 1 << TB;
    int sign = 1;
    if(dti < 0) { dti = -dti; sign = -1; }
    while(t > dti) t >>= 1;
    return t * sign;
}

struct Hier {
    Engine *E; const b200_step_params *sp; b200_gravshort_params *gp; b200_step_times *t;
    StepTimeline T;
};

static int hier_begin(Hier &H, Engine *E, const b200_step_params *sp, b200_gravshort_params *gp, b200_step_times *t, const char *who, bool fresh_store)
{
    if(int rc = step_need_state(E, who)) return rc;
    if(!sp || !gp || !t) return failmsg(E, std::string(who) + ": null argument");
    if(!sp->sync_loga || sp->nsync < 2 || !sp->gravkick_factor) return failmsg(E, std::string(who) + ": timeline / kick-factor callback missing");
    if(E->Nmesh == 0 && E->NmeshWalk == 0) return failmsg(E, std::string(who) + ": call b200_pm_init first");
    CK(E->st_sync.ensure((size_t) sp->nsync));
    CK(cudaMemcpyAsync(E->st_sync.p, sp->sync_loga, sp->nsync * sizeof(double), cudaMemcpyHostToDevice, E->stream));
    CK(cudaStreamSynchronize(E->stream));
    H.E = E; H.sp = sp; H.gp = gp; H.t = t;
    H.T.sync = E->st_sync.p; H.T.nsync = (int) sp->nsync;
    H.T.now = tl_loga_from_ti(sp, t->Ti_Current);
    H.T.ti_now = tl_ti_from_loga(sp, H.T.now);
    const size_t n = (size_t) (E->n > 0 ? E->n : 1);
    CK(E->st_listA.ensure(n + 1)); CK(E->st_listB.ensure(n + 1)); CK(E->st_store.ensure(3 * n)); CK(E->st_lower.ensure(3 * n));
    CK(E->st_cnt.ensure(1 + 6 * NBIN));
    // entries of particles outside the walked lists (garbage, inactive) stay zero rather than undefined
    if(fresh_store) { CK(cudaMemsetAsync(E->st_store.p, 0, 3 * n * sizeof(double), E->stream)); E->st_store_valid = false; }
    CK(cudaMemsetAsync(E->st_lower.p, 0, 3 * n * sizeof(double), E->stream));
    return 0;
}
// grav_short_tree_build_tree timestep.c:281-291
static int hier_gravity(Hier &H, const int *list, int64_t nlist, double *d_store)
{
    Engine *E = H.E;
    if(E->n == 0 || (list && nlist == 0)) return 0;
    k_step_oldacc<<<grid_for(E->n), 256, 0, E->stream>>>(E->n, E->s_fullacc.p, E->s_gravpm.p, E->oldacc.p); CKL(E);
    const double Box = step_box(E);
    if(int rc = tree_build(E, Box, 63, list, list ? nlist : 0, 0, nullptr)) return rc;
    // the walk set is the tree set: its particles in curve order
    if(E->tree_np > 0)
        if(int rc = grav_short_tree(E, H.gp, E->sidx.p, E->tree_np, d_store, nullptr, nullptr, true)) return rc;
    if(H.gp->TreeUseBH > 1) H.gp->TreeUseBH = 0;                           // gravshort-tree.c:150-151
    if(!list)                                                              // full tree: FullTreeGravAccel refreshed, gravshort.h:55-60
        CK(cudaMemcpyAsync(E->s_fullacc.p, d_store, 3 * (size_t) E->n * sizeof(double), cudaMemcpyDeviceToDevice, E->stream));
    return 0;
}
// apply_hierarchical_grav_kick timestep.c:238-279
static int hier_kick(Hier &H, const int *list, int64_t nlist, const double *d_acc, int ti, int largest_active)
{
    Engine *E = H.E;
    const b200_step_times *t = H.t;
    const int64_t dti = host_dti_of_bin(ti);
    double gravkick = H.sp->gravkick_factor(H.sp->user, t->Ti_kick[ti], t->Ti_kick[ti] + dti / 2);
    if(ti < largest_active) {
        const int64_t upper = host_dti_of_bin(ti + 1);
        gravkick -= H.sp->gravkick_factor(H.sp->user, t->Ti_kick[ti + 1], t->Ti_kick[ti + 1] + upper / 2);
    }
    if(nlist == 0) return 0;
    k_step_acc_kick<<<grid_for(nlist), 256, 0, E->stream>>>(nlist, list, E->flags.p, E->s_vel.p, d_acc, gravkick);
    CKL(E);
    return 0;
}
static int hier_largest_active(const b200_step_times *t)                   // timestep.c:311-318,506-513
{
    for(int ti = TB; ti >= 0; ti--)
        if(host_bin_active(ti, t->Ti_Current) && host_dti_of_bin(ti) <= t->PM_length) return ti;
    return TB;
}

// hierarchical_gravity_accelerations timestep.c:503-598 on the current active list (b200_step_build_active).
// The accelerations of the largest active bin stay in Engine::st_store for b200_step_hier_timesteps.
int step_hier_accelerations(Engine *E, const b200_step_params *sp, b200_gravshort_params *gp, b200_step_times *t, int64_t ngrav)
{
    Hier H;
    if(int rc = hier_begin(H, E, sp, gp, t, "b200_step_hier_accelerations", true)) return rc;
    const int largest_active = hier_largest_active(t);
    const int *last = E->st_act_implicit ? nullptr : E->st_act.p;
    int64_t nlast = E->st_nact;
    if(ngrav != E->st_nact) {                                              // :524-528
        if(int rc = step_sublist_of(E, last, nlast, largest_active, t->Ti_Current, E->st_listA.p, &nlast)) return rc;
        last = E->st_listA.p;
    }
    if(int rc = hier_gravity(H, last, nlast, E->st_store.p)) return rc;    // :533
    if(int rc = hier_kick(H, last, nlast, E->st_store.p, largest_active, largest_active)) return rc;   // :537
    bool have_lower = false;
    for(int ti = largest_active - 1; ti >= t->mingravtimebin; ti--) {      // :550-590
        int *dst = (last == E->st_listA.p) ? E->st_listB.p : E->st_listA.p;
        int64_t nsub = 0;
        if(int rc = step_sublist_of(E, last, nlast, ti, t->Ti_Current, dst, &nsub)) return rc;
        if(nsub != nlast) {                                                // :564-571: same set, same accelerations
            if(int rc = hier_gravity(H, dst, nsub, E->st_lower.p)) return rc;
            have_lower = true;
        }
        if(int rc = hier_kick(H, dst, nsub, have_lower ? E->st_lower.p : E->st_store.p, ti, largest_active)) return rc;   // :578-583
        last = dst; nlast = nsub;
    }
    CK(cudaStreamSynchronize(E->stream));
    E->st_store_valid = true;
    return 0;
}

// get_long_range_timestep_dloga + get_PM_timestep_ti timestep.c:1201-1298
static int hier_pm_timestep(Hier &H, double atime, double hubble, int64_t *dti_out)
{
    Engine *E = H.E;
    const b200_step_params *sp = H.sp;
    CK(E->st_part.ensure(STEP_BLOCKS * 18));
    k_step_vel_stats<<<STEP_BLOCKS, 256, 0, E->stream>>>(E->n, E->s_vel.p, E->mass.p, E->type.p, E->flags.p, E->st_part.p);
    CKL(E);
    std::vector<double> part(STEP_BLOCKS * 18);
    CK(cudaMemcpyAsync(part.data(), E->st_part.p, part.size() * sizeof(double), cudaMemcpyDeviceToHost, E->stream));
    CK(cudaStreamSynchronize(E->stream));
    double v[6] = {0}, c[6] = {0}, m[6];
    for(int k = 0; k < 6; k++) m[k] = 1e30;
    for(int b = 0; b < STEP_BLOCKS; b++)
        for(int k = 0; k < 6; k++) { v[k] += part[b * 18 + k]; c[k] += part[b * 18 + 6 + k]; if(part[b * 18 + 12 + k] < m[k]) m[k] = part[b * 18 + 12 + k]; }
    v[0] += v[4]; c[0] += c[4]; v[4] = v[0]; c[4] = c[0];                   // :1235-1244
    v[0] += v[5]; c[0] += c[5]; v[5] = v[0]; c[5] = c[0];
    m[5] = m[0];
    const double Box = step_box(E);
    const double asmth = E->Asmth * Box / (E->Nmesh > 0 ? E->Nmesh : E->NmeshWalk);
    double dloga = sp->MaxSizeTimestep;
    for(int ty = 0; ty < 6; ty++) {
        if(c[ty] == 0) continue;
        const double dmean = pow(m[ty] / (sp->omega_type[ty] * sp->RhoCrit), 1.0 / 3);
        const double dloga1 = sp->MaxRMSDisplacementFac * hubble * atime * atime * (asmth < dmean ? asmth : dmean) / sqrt(v[ty] / c[ty]);
        if(ty != sp->FastParticleType && dloga1 < dloga) dloga = dloga1;
    }
    if(dloga < sp->MinSizeTimestep) dloga = sp->MinSizeTimestep;
    const b200_step_times *t = H.t;
    int64_t dti = tl_round_down_pow2(tl_ti_from_loga(sp, dloga + H.T.now) - H.T.ti_now);
    const int64_t next = ((t->Ti_Current >> TB) + 1) << TB;                // find_next_sync_point: sync point i sits at i << TIMEBINS
    if((t->Ti_Current >> TB) + 1 >= sp->nsync) return failmsg(E, "b200_step_hier_timesteps: beyond the last sync point");
    const int64_t dti_max = next - t->PM_kick;
    if(dti > dti_max) dti = dti_max;
    *dti_out = dti;
    return 0;
}

// hierarchical_gravity_and_timesteps timestep.c:296-499.  Uses the accelerations left by
// b200_step_hier_accelerations (StoredGravAccel) or, without them, FullTreeGravAccel.
int step_hier_timesteps(Engine *E, const b200_step_params *sp, b200_gravshort_params *gp, b200_step_times *t, int64_t ngrav, int is_pm,
                        double atime, double hubble, int64_t *info)
{
    Hier H;
    if(int rc = hier_begin(H, E, sp, gp, t, "b200_step_hier_timesteps", false)) return rc;
    int64_t dti_max = t->PM_length;
    if(is_pm) {                                                            // :303-309
        if(int rc = hier_pm_timestep(H, atime, hubble, &dti_max)) return rc;
        t->PM_length = dti_max;
        t->PM_start = t->PM_kick;
    }
    int largest_active = hier_largest_active(t);
    const int *sub = E->st_act_implicit ? nullptr : E->st_act.p;
    int64_t nsub = E->st_nact;
    if(!(ngrav == E->st_nact || is_pm)) {                                  // :324-328
        if(int rc = step_sublist_of(E, sub, nsub, largest_active, t->Ti_Current, E->st_listA.p, &nsub)) return rc;
        sub = E->st_listA.p;
    }
    const double *d_top = E->st_store_valid ? E->st_store.p : E->s_fullacc.p;
    CK(cudaMemsetAsync(E->st_cnt.p, 0, (1 + 6 * NBIN) * sizeof(unsigned long long), E->stream));
    if(nsub > 0) {                                                         // :346-372
        k_step_assign_bins<<<grid_for(nsub), 256, 0, E->stream>>>(nsub, sub, E->flags.p, d_top, E->s_gravpm.p, H.T, atime, hubble,
            sp->ErrTolIntAccuracy, sp->softening, sp->MinSizeTimestep, (long long) dti_max, largest_active, E->s_bin_grav.p, E->st_cnt.p);
        CKL(E);
    }
    unsigned long long hc[NBIN];
    CK(cudaMemcpyAsync(hc, E->st_cnt.p, sizeof(hc), cudaMemcpyDeviceToHost, E->stream));
    CK(cudaStreamSynchronize(E->stream));
    int64_t counts[NBIN];
    for(int k = 0; k < NBIN; k++) counts[k] = (int64_t) hc[k];
    for(int ti = largest_active; ti >= 1; ti--)                            // :383-387
        if(counts[ti] > 0) { largest_active = ti; break; }
    int push_down = largest_active;                                        // :394-413
    if(is_pm)
        for(int ti = largest_active; ti >= 1; ti--) {
            if(counts[ti] / 3 > counts[ti - 1]) break;
            push_down = ti - 1;
            counts[ti - 1] += counts[ti];
        }
    if(push_down == 0) return failmsg(E, "b200_step_hier_timesteps: bad timestep, every particle pushed to bin 0 (timestep.c:402-403)");
    if(push_down != largest_active) {
        if(nsub > 0) { k_step_push_down<<<grid_for(nsub), 256, 0, E->stream>>>(nsub, sub, push_down, E->s_bin_grav.p); CKL(E); }
        largest_active = push_down;
    }
    t->maxtimebin = largest_active;                                        // :415
    if(int rc = hier_kick(H, sub, nsub, d_top, largest_active, largest_active)) return rc;   // :418
    CK(cudaMemsetAsync(E->st_cnt.p, 0, sizeof(unsigned long long), E->stream));
    const int *last = sub;
    int64_t nlast = nsub;
    for(int ti = largest_active - 1; ti > 0; ti--) {                       // :435-493
        int *dst = (last == E->st_listA.p) ? E->st_listB.p : E->st_listA.p;
        int64_t nnew = 0;
        if(int rc = step_sublist_of(E, last, nlast, ti, t->Ti_Current, dst, &nnew)) return rc;
        if(nnew == 0) { t->mingravtimebin = ti + 1; break; }               // :443-447
        if(int rc = hier_gravity(H, dst, nnew, E->st_lower.p)) return rc;
        k_step_reduce_bins<<<grid_for(nnew), 256, 0, E->stream>>>(nnew, dst, E->flags.p, E->st_lower.p, E->s_gravpm.p, H.T, atime, hubble,
            sp->ErrTolIntAccuracy, sp->softening, sp->MinSizeTimestep, (long long) dti_max, ti, E->s_bin_grav.p, E->st_cnt.p);
        CKL(E);
        if(int rc = hier_kick(H, dst, nnew, E->st_lower.p, ti, largest_active)) return rc;   // :474
        last = dst; nlast = nnew;
    }
    unsigned long long bad = 0;
    CK(cudaMemcpyAsync(&bad, E->st_cnt.p, sizeof(bad), cudaMemcpyDeviceToHost, E->stream));
    CK(cudaStreamSynchronize(E->stream));
    t->mintimebin = t->mingravtimebin;                                     // :496
    E->st_store_valid = false;                                             // the reference frees StoredGravAccel here (:419-420)
    if(info) { info[0] = largest_active; info[1] = t->PM_length; info[2] = (int64_t) bad; }
    return 0;
}

// find_hydro_timesteps timestep.c:617-738 on the current active list (gas; black holes not supported).
// maxsig: host array SphP[].MaxSignalVel by particle index, or NULL to use the one b200_hydro_force left on the device.
int step_hydro_timesteps(Engine *E, const b200_step_params *sp, b200_step_times *t, const double *maxsig, double atime, double hubble, int64_t *nbad)
{
    if(int rc = step_need_state(E, "b200_step_hydro_timesteps")) return rc;
    if(!sp || !t || !sp->sync_loga || sp->nsync < 2) return failmsg(E, "b200_step_hydro_timesteps: null argument / timeline missing");
    const size_t n = (size_t) (E->n > 0 ? E->n : 1);
    CK(E->st_sync.ensure((size_t) sp->nsync)); CK(E->st_cnt.ensure(1 + 6 * NBIN)); CK(E->st_maxsig.ensure(n));
    CK(cudaMemcpyAsync(E->st_sync.p, sp->sync_loga, sp->nsync * sizeof(double), cudaMemcpyHostToDevice, E->stream));
    if(maxsig) { CK(cudaMemcpyAsync(E->st_maxsig.p, maxsig, E->n * sizeof(double), cudaMemcpyHostToDevice, E->stream)); E->st_maxsig_valid = true; }
    if(!E->st_maxsig_valid) return failmsg(E, "b200_step_hydro_timesteps: no MaxSignalVel (pass it, or run b200_hydro_force with the step state set)");
    const unsigned long long init[2] = {0ull, (unsigned long long) TB};
    CK(cudaMemcpyAsync(E->st_cnt.p, init, sizeof(init), cudaMemcpyHostToDevice, E->stream));
    CK(cudaStreamSynchronize(E->stream));           // init is a stack array
    StepTimeline T;
    T.sync = E->st_sync.p; T.nsync = (int) sp->nsync;
    T.now = tl_loga_from_ti(sp, t->Ti_Current); T.ti_now = tl_ti_from_loga(sp, T.now);
    const double fac3 = pow(atime, 3 * (1 - 5.0 / 3) / 2.0);              // GAMMA = 5/3, physconst.h
    const int64_t nl = E->st_nact;
    if(nl > 0) {
        k_step_hydro_bins<<<grid_for(nl), 256, 0, E->stream>>>(nl, E->st_act_implicit ? nullptr : E->st_act.p, E->type.p, E->flags.p, E->s_hsml.p,
            E->s_dthsml.p, E->st_maxsig.p, E->s_bin_grav.p, E->s_bin_hydro.p, T, atime, hubble, fac3, sp->CourantFac, sp->MinSizeTimestep,
            (long long) t->PM_length, (long long) t->Ti_Current, E->st_cnt.p);
        CKL(E);
    }
    unsigned long long h[2];
    CK(cudaMemcpyAsync(h, E->st_cnt.p, sizeof(h), cudaMemcpyDeviceToHost, E->stream));
    CK(cudaStreamSynchronize(E->stream));
    int mTimeBin = (int) h[1];
    if(!host_bin_active(mTimeBin, t->Ti_Current)) {                        // :713-719
        mTimeBin = t->mintimebin;
        if(host_bin_active(mTimeBin + 1, t->Ti_Current)) mTimeBin++;
    }
    t->mintimebin = mTimeBin;                                              // :727-731
    if(t->mintimebin > t->mingravtimebin && t->mingravtimebin > 0) t->mintimebin = t->mingravtimebin;
    if(nbad) *nbad = (int64_t) h[0];
    return 0;
}

// The current active list and the per-bin factor tables for b200_density / b200_hydro_force, the time bins being
// the ones already on the device (what b200_sph_set_timebins + b200_sph_set_active do from host arrays).
int step_sph_prepare(Engine *E, const b200_sph_bins *tables)
{
    if(int rc = step_need_state(E, "b200_step_sph_prepare")) return rc;
    if(!tables) return failmsg(E, "b200_step_sph_prepare: null factor tables");
    const size_t n = (size_t) (E->n > 0 ? E->n : 1);
    CK(E->s_bins.ensure(5 * NBIN)); CK(E->s_active.ensure(n));
    CK(cudaMemcpyAsync(E->s_bins.p, tables, 5 * NBIN * sizeof(double), cudaMemcpyHostToDevice, E->stream));
    E->s_bins_set = true;
    E->s_active_set = false;
    if(!E->st_act_implicit) {
        CK(cudaMemsetAsync(E->s_active.p, 0, n, E->stream));
        if(E->st_nact > 0) { k_step_mark_active<<<grid_for(E->st_nact), 256, 0, E->stream>>>(E->st_nact, E->st_act.p, E->s_active.p); CKL(E); }
        E->s_active_set = true;
    }
    CK(cudaStreamSynchronize(E->stream));      // `tables` may be a stack object of the caller
    return 0;
}
// SphP[].HydroAccel / DtEntropy / MaxSignalVel of the listed gas <- the device results of the last b200_hydro_force
int step_adopt_hydro(Engine *E)
{
    if(int rc = step_need_state(E, "b200_step_adopt_hydro")) return rc;
    const size_t n = (size_t) (E->n > 0 ? E->n : 1);
    if(E->s_out3.cap < 3 * n || E->s_out1a.cap < n || E->s_out1b.cap < n) return failmsg(E, "b200_step_adopt_hydro: no hydro results on the device (run b200_hydro_force)");
    CK(E->st_maxsig.ensure(n));
    if(!E->st_maxsig_valid) CK(cudaMemsetAsync(E->st_maxsig.p, 0, n * sizeof(double), E->stream));
    const int64_t nl = E->st_nact;
    if(nl > 0) {
        k_step_adopt_hydro<<<grid_for(nl), 256, 0, E->stream>>>(nl, E->st_act_implicit ? nullptr : E->st_act.p, E->type.p, E->flags.p,
            E->s_out3.p, E->s_out1a.p, E->s_out1b.p, E->s_hydroacc.p, E->s_dtentropy.p, E->st_maxsig.p);
        CKL(E);
    }
    CK(cudaStreamSynchronize(E->stream));
    E->st_maxsig_valid = true;
    return 0;
}

// force_tree_full + grav_short_tree for the current active list (run.c:541-548, the SplitGravityTimestepsOn = 0 loop): a tree
// over every particle, the walk for the listed ones, their FullTreeGravAccel in the step state replaced (the walk reads the
// old modulus from Engine::oldacc, so writing into s_fullacc in place is safe).
int step_grav_short_tree(Engine *E, b200_gravshort_params *gp)
{
    if(int rc = step_need_state(E, "b200_step_grav_short_tree")) return rc;
    if(!gp) return failmsg(E, "b200_step_grav_short_tree: null parameters");
    if(E->Nmesh == 0 && E->NmeshWalk == 0) return failmsg(E, "b200_step_grav_short_tree: call b200_pm_init first");
    if(E->n == 0) return 0;
    k_step_oldacc<<<grid_for(E->n), 256, 0, E->stream>>>(E->n, E->s_fullacc.p, E->s_gravpm.p, E->oldacc.p); CKL(E);
    if(int rc = tree_build(E, step_box(E), 63, nullptr, 0, 0, nullptr)) return rc;
    if(E->st_act_implicit) { if(int rc = grav_short_tree(E, gp, nullptr, 0, E->s_fullacc.p, nullptr, nullptr)) return rc; }
    else if(E->st_nact > 0) { if(int rc = grav_short_tree(E, gp, E->st_act.p, E->st_nact, E->s_fullacc.p, nullptr, nullptr)) return rc; }
    if(gp->TreeUseBH > 1) gp->TreeUseBH = 0;                               // gravshort-tree.c:150-151
    CK(cudaStreamSynchronize(E->stream));
    return 0;
}

// find_timesteps timestep.c:739-853 on the current active list (the SplitGravityTimestepsOn = 0 loop: followed by
// b200_step_half_kick with hydro_only = 0).  maxsig as in b200_step_hydro_timesteps (ignored without gas).
int step_find_timesteps(Engine *E, const b200_step_params *sp, b200_step_times *t, const double *maxsig, int is_pm, double atime, double hubble,
                        int64_t *nbad)
{
    if(int rc = step_need_state(E, "b200_step_find_timesteps")) return rc;
    if(!sp || !t || !sp->sync_loga || sp->nsync < 2) return failmsg(E, "b200_step_find_timesteps: null argument / timeline missing");
    if(E->Nmesh == 0 && E->NmeshWalk == 0) return failmsg(E, "b200_step_find_timesteps: call b200_pm_init first (PM smoothing scale)");
    const size_t n = (size_t) (E->n > 0 ? E->n : 1);
    CK(E->st_sync.ensure((size_t) sp->nsync)); CK(E->st_cnt.ensure(1 + 6 * NBIN)); CK(E->st_maxsig.ensure(n));
    CK(cudaMemcpyAsync(E->st_sync.p, sp->sync_loga, sp->nsync * sizeof(double), cudaMemcpyHostToDevice, E->stream));
    if(maxsig) { CK(cudaMemcpyAsync(E->st_maxsig.p, maxsig, E->n * sizeof(double), cudaMemcpyHostToDevice, E->stream)); E->st_maxsig_valid = true; }
    if(E->st_have_gas && !E->st_maxsig_valid) return failmsg(E, "b200_step_find_timesteps: no MaxSignalVel for the gas");
    Hier H;
    H.E = E; H.sp = sp; H.gp = nullptr; H.t = t;
    H.T.sync = E->st_sync.p; H.T.nsync = (int) sp->nsync;
    H.T.now = tl_loga_from_ti(sp, t->Ti_Current); H.T.ti_now = tl_ti_from_loga(sp, H.T.now);
    int64_t dti_max = t->PM_length;
    if(is_pm) {                                                            // :751-755
        if(int rc = hier_pm_timestep(H, atime, hubble, &dti_max)) return rc;
        t->PM_length = dti_max;
        t->PM_start = t->PM_kick;
    }
    const unsigned long long init[3] = {0ull, (unsigned long long) TB, 0ull};
    CK(cudaMemcpyAsync(E->st_cnt.p, init, sizeof(init), cudaMemcpyHostToDevice, E->stream));
    CK(cudaStreamSynchronize(E->stream));           // init is a stack array
    const double fac3 = pow(atime, 3 * (1 - 5.0 / 3) / 2.0);
    const int64_t nl = E->st_nact;
    if(nl > 0) {
        k_step_find_bins<<<grid_for(nl), 256, 0, E->stream>>>(nl, E->st_act_implicit ? nullptr : E->st_act.p, E->type.p, E->flags.p, E->s_fullacc.p,
            E->s_gravpm.p, E->s_hsml.p, E->s_dthsml.p, E->st_maxsig.p, E->s_bin_grav.p, E->s_bin_hydro.p, H.T, atime, hubble, fac3,
            sp->ErrTolIntAccuracy, sp->softening, sp->CourantFac, sp->MinSizeTimestep, (long long) dti_max, (long long) t->Ti_Current, E->st_cnt.p);
        CKL(E);
    }
    unsigned long long h[3];
    CK(cudaMemcpyAsync(h, E->st_cnt.p, sizeof(h), cudaMemcpyDeviceToHost, E->stream));
    CK(cudaStreamSynchronize(E->stream));
    const int maxTimeBin = (int) h[2];
    if(is_pm && t->PM_length > host_dti_of_bin(maxTimeBin)) t->PM_length = host_dti_of_bin(maxTimeBin);       // :835-836
    t->mintimebin = (int) h[1]; t->maxtimebin = maxTimeBin;
    if(nbad) *nbad = (int64_t) h[0];
    return 0;
}

void step_release(Engine *E)
{
    E->st_iota.release(); E->st_listA.release(); E->st_listB.release(); E->st_act.release(); E->st_flag.release();
    E->st_store.release(); E->st_lower.release(); E->st_sync.release(); E->st_part.release(); E->st_tab.release(); E->st_cnt.release(); E->st_maxsig.release();
}

} // namespace b200

// ---- C-ABI ----------------------------------------------------------------------------------
using namespace b200;
#define STEP_ENTER(ctx) if(!(ctx)) return 1; Engine *E = &(ctx)->e; if(cudaSetDevice(E->device) != cudaSuccess) return failmsg(E, "cudaSetDevice failed")

extern "C" {
int b200_step_set_state(b200_ctx *ctx, const b200_step_state *state) { STEP_ENTER(ctx); return step_set_state(E, state); }
int b200_step_get_state(b200_ctx *ctx, b200_step_state_out *out) { STEP_ENTER(ctx); return step_get_state(E, out); }
int b200_step_adopt_forces(b200_ctx *ctx, int tree, int pm) { STEP_ENTER(ctx); return step_adopt_forces(E, tree, pm); }
int b200_step_drift(b200_ctx *ctx, double ddrift, const double *shift, int64_t *nbad) { STEP_ENTER(ctx); return step_drift(E, ddrift, shift, nbad); }
int b200_step_build_active(b200_ctx *ctx, int64_t Ti_Current, int is_pm, int64_t nhydro_slots, int64_t *counts, int64_t *bincounts)
{
    STEP_ENTER(ctx);
    return step_build_active(E, Ti_Current, is_pm, nhydro_slots, counts, bincounts);
}
int b200_step_active_sublist(b200_ctx *ctx, int maxtimebin, int64_t Ti_Current, int64_t *nsub)
{
    STEP_ENTER(ctx);
    return step_active_sublist(E, maxtimebin, Ti_Current, nsub);
}
int b200_step_set_active(b200_ctx *ctx, const int32_t *list, int64_t nlist) { STEP_ENTER(ctx); return step_set_active(E, list, nlist); }
int b200_step_set_store(b200_ctx *ctx, const double *in) { STEP_ENTER(ctx); return step_set_store(E, in); }
int b200_step_get_store(b200_ctx *ctx, double *out) { STEP_ENTER(ctx); return step_get_store(E, out); }
int b200_step_get_active(b200_ctx *ctx, int which, int32_t *out, int64_t *nout) { STEP_ENTER(ctx); return step_get_active(E, which, out, nout); }
int b200_step_half_kick(b200_ctx *ctx, const double *gravkick, const double *hydrokick, const double *dt_entr, int64_t Ti_Current,
                        double atime, double MaxGasVel, int hydro_only)
{
    STEP_ENTER(ctx);
    return step_half_kick(E, gravkick, hydrokick, dt_entr, Ti_Current, atime, MaxGasVel, hydro_only);
}
int b200_step_pm_kick(b200_ctx *ctx, double Fgravkick) { STEP_ENTER(ctx); return step_pm_kick(E, Fgravkick); }
int b200_step_sph_prepare(b200_ctx *ctx, const b200_sph_bins *tables) { STEP_ENTER(ctx); return step_sph_prepare(E, tables); }
int b200_step_adopt_hydro(b200_ctx *ctx) { STEP_ENTER(ctx); return step_adopt_hydro(E); }
int b200_step_grav_short_tree(b200_ctx *ctx, b200_gravshort_params *gp) { STEP_ENTER(ctx); return step_grav_short_tree(E, gp); }
int b200_step_find_timesteps(b200_ctx *ctx, const b200_step_params *sp, b200_step_times *times, const double *maxsignalvel, int is_pm,
                             double atime, double hubble, int64_t *nbad)
{
    STEP_ENTER(ctx);
    return step_find_timesteps(E, sp, times, maxsignalvel, is_pm, atime, hubble, nbad);
}
int b200_step_hydro_timesteps(b200_ctx *ctx, const b200_step_params *sp, b200_step_times *times, const double *maxsignalvel, double atime,
                              double hubble, int64_t *nbad)
{
    STEP_ENTER(ctx);
    return step_hydro_timesteps(E, sp, times, maxsignalvel, atime, hubble, nbad);
}
int b200_step_hier_accelerations(b200_ctx *ctx, const b200_step_params *sp, b200_gravshort_params *gp, b200_step_times *times, int64_t ngrav)
{
    STEP_ENTER(ctx);
    return step_hier_accelerations(E, sp, gp, times, ngrav);
}
int b200_step_hier_timesteps(b200_ctx *ctx, const b200_step_params *sp, b200_gravshort_params *gp, b200_step_times *times, int64_t ngrav,
                             int is_pm, double atime, double hubble, int64_t *info)
{
    STEP_ENTER(ctx);
    return step_hier_timesteps(E, sp, gp, times, ngrav, is_pm, atime, hubble, info);
}
}
