/* oracle_step.c -- CPU restatement of the reference's step loop around the force computation:
 * the integer timeline, drift, active lists, half kicks, and the hierarchical gravity driver
 * that assigns gravity time bins.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  Every function cites the reference lines it follows
 * (paths relative to libgadget/).  PINNED against the reference's own drift.c / timestep.c /
 * timebinmgr.c compiled unmodified (oracle/_ref/libref_step.so, tests/golden/ref_step.npz).
 *
 * Not restated (GSL): the background cosmology and the kick / drift integrals.  The fixtures use
 * a flat matter + Lambda background, H(a) = H0 sqrt(Om / a^3 + 1 - Om), and integrate the
 * integrands of timefac.c:12-38 with 64 eight-point Gauss-Legendre panels -- the same stand-in
 * the reference build of the fixture uses (oracle/ref_driver.c), so both sides see equal factors. */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

#define TB ORACLE_TIMEBINS
#define TIMEBASE_ (((int64_t) 1) << TB)

/* ---- integer timeline: timebinmgr.c:380-462, timebinmgr.h:47-50 ---- */
static int64_t dti_of_bin(int bin) { return bin > 0 ? (((int64_t) 1) << bin) : 0; }

static double interval_of(const oracle_timeline *tl, int64_t ti)          /* Dloga_interval_ti :400-414 */
{
    const int64_t lastsnap = ti >> TB;
    if(lastsnap >= tl->nsync - 1) return 0;
    return (tl->loga[lastsnap + 1] - tl->loga[lastsnap]) / TIMEBASE_;
}
double oracle_loga_from_ti(const oracle_timeline *tl, int64_t ti)         /* :380-391 */
{
    const int64_t lastsnap = ti >> TB;
    const int64_t dti = ti & (TIMEBASE_ - 1);
    return tl->loga[lastsnap] + dti * interval_of(tl, ti);
}
int64_t oracle_ti_from_loga(const oracle_timeline *tl, double loga)       /* :360-378 */
{
    int64_t i;
    for(i = 1; i < tl->nsync - 1; i++)
        if(tl->loga[i] > loga) break;
    const double step = (tl->loga[i] - tl->loga[i - 1]) / TIMEBASE_;
    int64_t ti = (i - 1) << TB;
    ti = (int64_t) ((double) ti + (loga - tl->loga[i - 1]) / step);       /* `ti += double` in the reference */
    return ti;
}
int64_t oracle_dti_from_dloga(const oracle_timeline *tl, double dloga, int64_t Ti_Current)   /* :435-440 */
{
    const double now = oracle_loga_from_ti(tl, Ti_Current);
    return oracle_ti_from_loga(tl, dloga + now) - oracle_ti_from_loga(tl, now);
}
double oracle_dloga_from_dti(const oracle_timeline *tl, int64_t dti, int64_t Ti_Current)     /* :420-432 */
{
    return interval_of(tl, Ti_Current) * dti;
}
int oracle_is_timebin_active(int bin, int64_t ti)                         /* timestep.c:143-150 */
{
    if(bin <= 0 || ti <= 0) return 1;
    return ti % dti_of_bin(bin) == 0;
}
static int64_t round_down_pow2(int64_t dti)                               /* timebinmgr.c:449-462 */
{
    int64_t t = TIMEBASE_;
    int sign = 1;
    if(dti < 0) { dti = -dti; sign = -1; }
    while(t > dti) t >>= 1;
    return t * sign;
}
static int bin_of_dti(int64_t dti)                                        /* get_timestep_bin timestep.c:1301-1315 */
{
    if(dti <= 1) return 0;
    int bin = -1;
    while(dti) { bin++; dti >>= 1; }
    return bin;
}

/* ---- background + integrals (stand-in shared with oracle/ref_driver.c, see header) ---- */
static double hubble_of(const oracle_cosmo *c, double a) { return c->Hubble * sqrt(c->Omega0 / (a * a * a) + (1 - c->Omega0)); }
static double integrand_of(const oracle_cosmo *c, int kind, double a)
{
    const double h = hubble_of(c, a);
    if(kind == 0) return 1 / (h * a * a * a);                             /* drift     timefac.c:12-17 */
    if(kind == 1) return 1 / (h * a * a);                                 /* gravkick  timefac.c:20-26 */
    return 1 / (h * pow(a, 3 * (5.0 / 3 - 1)) * a);                       /* hydrokick timefac.c:30-38, GAMMA = 5/3 */
}
double oracle_step_factor(const oracle_cosmo *c, const oracle_timeline *tl, int kind, int64_t t0, int64_t t1)
{
    if(t0 == t1) return 0;                                                /* timefac.c:44-45 */
    static const double gx[4] = {0.1834346424956498, 0.5255324099163290, 0.7966664774136267, 0.9602898564975363};
    static const double gw[4] = {0.3626837833783620, 0.3137066458778873, 0.2223810344533745, 0.1012285362903763};
    const double a0 = exp(oracle_loga_from_ti(tl, t0)), a1 = exp(oracle_loga_from_ti(tl, t1));
    const int NP = 64;
    double sum = 0;
    for(int p = 0; p < NP; p++) {
        const double lo = a0 + (a1 - a0) * p / NP, hi = a0 + (a1 - a0) * (p + 1) / NP;
        const double mid = 0.5 * (lo + hi), hw = 0.5 * (hi - lo);
        for(int k = 0; k < 4; k++) sum += gw[k] * hw * (integrand_of(c, kind, mid - hw * gx[k]) + integrand_of(c, kind, mid + hw * gx[k]));
    }
    return sum;
}

/* ---- drift.c:17-102 (real_drift_particle for types other than black holes) ----
 * flags bit 0 IsGarbage, bit 1 Swallowed.  Returns the number of particles the reference would
 * stop on (Hsml <= 0 or a non-finite position). */
int64_t oracle_drift(int64_t n, double *pos, const double *vel, const uint8_t *type, const uint8_t *flags,
                     double *hsml, const double *dthsml, double ddrift, const double *shift, double BoxSize)
{
    int64_t bad = 0;
    for(int64_t i = 0; i < n; i++) {
        double *x = pos + 3 * i;
        if(flags && (flags[i] & 3)) {                                     /* :20-29 */
            for(int j = 0; j < 3; j++) {
                x[j] += shift[j];
                while(x[j] > BoxSize) x[j] -= BoxSize;
                while(x[j] <= 0) x[j] += BoxSize;
            }
            continue;
        }
        if(type[i] == 0 && hsml) {                                        /* :54-67 */
            hsml[i] += dthsml[i] * ddrift;
            if(hsml[i] <= 0) bad++;
            if(hsml[i] > BoxSize / 2.) hsml[i] = BoxSize / 2.;
        }
        for(int j = 0; j < 3; j++) {                                      /* :68-79 */
            x[j] += vel[3 * i + j] * ddrift + shift[j];
            if(!isfinite(x[j])) { bad++; x[j] = BoxSize; }
        }
        for(int j = 0; j < 3; j++) {
            while(x[j] > BoxSize) x[j] -= BoxSize;
            while(x[j] <= 0) x[j] += BoxSize;
        }
    }
    return bad;
}

/* ---- build_active_particles timestep.c:1334-1431 ----
 * is_pm: the list is implicit (returns -1 - n); counts = {NumActiveParticle, NumActiveGravity,
 * NumActiveHydro}; bincounts[6][TIMEBINS+1] = TimeBinCountType.  nhydro_slots = gas + BH slots
 * in use (SlotsManager->info[0].size + info[5].size), which the PM branch reports. */
int64_t oracle_build_active(int64_t n, const uint8_t *type, const uint8_t *flags, const uint8_t *bin_grav, const uint8_t *bin_hydro,
                            int64_t Ti_Current, int is_pm, int64_t nhydro_slots, int32_t *list_out, int64_t *counts, int64_t *bincounts)
{
    memset(bincounts, 0, sizeof(int64_t) * 6 * (TB + 1));
    int64_t na = 0, ngrav = 0;
    for(int64_t i = 0; i < n; i++) {
        if(flags && (flags[i] & 3)) continue;
        const int hydro_particle = type[i] == 0 || type[i] == 5;
        bincounts[(TB + 1) * type[i] + (hydro_particle ? bin_hydro[i] : bin_grav[i])]++;
        if(is_pm) continue;
        const int hydro_active = hydro_particle && oracle_is_timebin_active(bin_hydro[i], Ti_Current);
        const int gravity_active = oracle_is_timebin_active(bin_grav[i], Ti_Current);
        ngrav += gravity_active;
        if(hydro_active || gravity_active) list_out[na++] = (int32_t) i;
    }
    if(is_pm) { counts[0] = n; counts[1] = n; counts[2] = nhydro_slots; return -1 - n; }
    counts[0] = na; counts[1] = ngrav; counts[2] = na;                    /* nactivehydro counts every listed particle, :1392 */
    return na;
}
/* build_active_sublist timestep.c:1435-1478; list == NULL: all n particles */
int64_t oracle_active_sublist(const int32_t *list, int64_t nlist, const uint8_t *flags, const uint8_t *bin_grav,
                              int maxtimebin, int64_t Ti_Current, int32_t *out)
{
    int64_t na = 0;
    for(int64_t q = 0; q < nlist; q++) {
        const int64_t i = list ? list[q] : q;
        if(flags && (flags[i] & 3)) continue;
        if(bin_grav[i] > maxtimebin) continue;
        if(!oracle_is_timebin_active(bin_grav[i], Ti_Current)) continue;
        out[na++] = (int32_t) i;
    }
    return na;
}

/* ---- kicks ---- */
/* do_hydro_kick timestep.c:1003-1043 for gas */
static void hydro_kick_one(double *v, const double *hydroacc, double *entropy, double dtentropy, double dt_entr, double Fhydrokick,
                           double atime, double MaxGasVel)
{
    for(int j = 0; j < 3; j++) v[j] += hydroacc[j] * Fhydrokick;
    double vv = 0;
    for(int j = 0; j < 3; j++) vv += v[j] * v[j];
    vv = sqrt(vv);
    if(vv > 0 && vv / atime > MaxGasVel)
        for(int j = 0; j < 3; j++) v[j] *= MaxGasVel * atime / vv;
    *entropy += dtentropy * dt_entr;
}
/* apply_half_kick (hydro_only = 0) timestep.c:874-928 / apply_hydro_half_kick (1) :931-970.
 * gravkick, hydrokick, dt_entr: [TIMEBINS+1] by bin (dt_entr[b] = dloga_from_dti(dti(b)/2)); bins
 * that are not active carry 0 in gravkick / hydrokick exactly as the reference's zero-initialised
 * tables do.  list == NULL: all n particles. */
void oracle_half_kick(const int32_t *list, int64_t nlist, const uint8_t *type, const uint8_t *flags, const uint8_t *bin_grav,
                      const uint8_t *bin_hydro, double *vel, const double *fullacc, const double *hydroacc, double *entropy,
                      const double *dtentropy, const double *gravkick, const double *hydrokick, const double *dt_entr,
                      int64_t Ti_Current, double atime, double MaxGasVel, int hydro_only)
{
    for(int64_t q = 0; q < nlist; q++) {
        const int64_t i = list ? list[q] : q;
        if(flags && (flags[i] & 3)) continue;
        if(!hydro_only && oracle_is_timebin_active(bin_grav[i], Ti_Current))
            for(int j = 0; j < 3; j++) vel[3 * i + j] += fullacc[3 * i + j] * gravkick[bin_grav[i]];
        if(type[i] == 0)         /* type 5 adds the black-hole drag terms (not supported here) */
            hydro_kick_one(vel + 3 * i, hydroacc + 3 * i, entropy + i, dtentropy[i], dt_entr[bin_hydro[i]], hydrokick[bin_hydro[i]], atime, MaxGasVel);
    }
}
/* apply_PM_half_kick timestep.c:972-993 */
void oracle_pm_kick(int64_t n, const uint8_t *flags, double *vel, const double *gravpm, double Fgravkick)
{
    for(int64_t i = 0; i < n; i++) {
        if(flags && (flags[i] & 3)) continue;
        for(int j = 0; j < 3; j++) vel[3 * i + j] += gravpm[3 * i + j] * Fgravkick;
    }
}
/* the particle loop of apply_hierarchical_grav_kick timestep.c:258-276 */
void oracle_grav_kick(const int32_t *list, int64_t nlist, const uint8_t *flags, double *vel, const double *acc, double gravkick)
{
    for(int64_t q = 0; q < nlist; q++) {
        const int64_t i = list ? list[q] : q;
        if(flags && (flags[i] & 3)) continue;
        for(int j = 0; j < 3; j++) vel[3 * i + j] += acc[3 * i + j] * gravkick;
    }
}
/* update_kick_times timestep.c:215-235 */
void oracle_update_kick_times(oracle_times *t)
{
    if(t->mintimebin == 0 && t->maxtimebin == 0) return;
    for(int bin = t->mintimebin; bin <= TB; bin++)
        if(oracle_is_timebin_active(bin, t->Ti_Current)) t->Ti_kick[bin] += dti_of_bin(bin) / 2;
    for(int bin = 1; bin < t->mintimebin; bin++) t->Ti_kick[bin] += dti_of_bin(t->mintimebin) / 2;
}
/* update_lastactive_drift timestep.c:860-871 */
void oracle_update_lastactive_drift(oracle_times *t)
{
    for(int bin = 0; bin <= TB; bin++)
        if(oracle_is_timebin_active(bin, t->Ti_Current)) t->Ti_lastactivedrift[bin] = t->Ti_Current;
}

/* ---- gravity time step ---- */
/* get_timestep_gravity_dloga + grav_acceleration2 timestep.c:1046-1073 */
double oracle_gravity_dloga(const double *acc, const double *gravpm, double atime, double hubble, double ErrTolIntAccuracy, double softening)
{
    const double a2inv = 1 / (atime * atime);
    double ax = a2inv * acc[0], ay = a2inv * acc[1], az = a2inv * acc[2];
    ay += a2inv * gravpm[1]; ax += a2inv * gravpm[0]; az += a2inv * gravpm[2];
    double ac2 = ax * ax + ay * ay + az * az;
    if(ac2 == 0) ac2 = 1.0e-60;
    const double ac = sqrt(ac2);
    const double dt = sqrt(2 * ErrTolIntAccuracy * atime * (softening / 2.8) / ac);
    return dt * hubble;
}
/* convert_timestep_to_ti timestep.c:1155-1173 */
int64_t oracle_convert_timestep(const oracle_timeline *tl, double dloga, int64_t dti_max, int64_t Ti_Current, double MinSizeTimestep)
{
    if(dti_max == 0) return 0;
    if(dloga < MinSizeTimestep) dloga = MinSizeTimestep;
    int64_t dti = oracle_dti_from_dloga(tl, dloga, Ti_Current);
    if(dti > dti_max || dti < 0) dti = dti_max;
    return dti;
}
/* The bin a gravitationally active particle is given at the top of the hierarchy,
 * timestep.c:349-371: power-of-two round down, bin, clamp to largest_active. */
int oracle_gravity_timebin(const oracle_timeline *tl, const double *acc, const double *gravpm, const oracle_step_params *sp,
                           double atime, double hubble, int64_t dti_max, int64_t Ti_Current, int largest_active)
{
    const double dloga = oracle_gravity_dloga(acc, gravpm, atime, hubble, sp->ErrTolIntAccuracy, sp->softening);
    int64_t dti = oracle_convert_timestep(tl, dloga, dti_max, Ti_Current, sp->MinSizeTimestep);
    dti = round_down_pow2(dti);
    int bin = bin_of_dti(dti);
    if(bin > largest_active) bin = largest_active;
    return bin;
}

/* find_hydro_timesteps timestep.c:617-738 for gas (with get_timestep_hydro_dloga :1075-1117 and
 * get_timebin_from_dti :166-182): Courant and smoothing-length criteria -> TimeBinHydro of the listed gas
 * particles, then the new minimum time bin.  Returns the bad-step count. */
int oracle_hydro_timebins(const oracle_timeline *tl, const oracle_step_params *sp, oracle_times *t, const int32_t *list, int64_t nlist,
                          const uint8_t *type, const uint8_t *flags, const double *hsml, const double *dthsml, const double *maxsig,
                          const uint8_t *bin_grav, uint8_t *bin_hydro, double atime, double hubble)
{
    const int64_t dti_max = t->PM_length;
    const double fac3 = pow(atime, 3 * (1 - 5.0 / 3) / 2.0);
    int bad = 0, mTimeBin = TB;
    for(int64_t q = 0; q < nlist; q++) {
        const int64_t i = list ? list[q] : q;
        if(flags && (flags[i] & 3)) continue;
        if(type[i] != 0) continue;                                         /* type 5 needs the black-hole slots: not supported */
        double dt = 2 * sp->CourantFac * atime * hsml[i] / (fac3 * maxsig[i]);
        const double dt_hsml = sp->CourantFac * atime * atime * fabs(hsml[i] / (dthsml[i] + 1e-20));
        if(dt_hsml < dt) dt = dt_hsml;
        const int64_t dti = round_down_pow2(oracle_convert_timestep(tl, dt * hubble, dti_max, t->Ti_Current, sp->MinSizeTimestep));
        int bin = bin_of_dti(dti);
        const int binold = bin_hydro[i];
        if(bin > binold)
            while(!oracle_is_timebin_active(bin, t->Ti_Current) && bin > binold && bin > 1) bin--;
        if(bin > bin_grav[i]) bin = bin_grav[i];
        if(bin < 1) bad++;
        if(oracle_is_timebin_active(binold, t->Ti_Current) && oracle_is_timebin_active(bin, t->Ti_Current)) bin_hydro[i] = (uint8_t) bin;
        if(bin < mTimeBin) mTimeBin = bin;
    }
    if(!oracle_is_timebin_active(mTimeBin, t->Ti_Current)) {               /* :713-719 */
        mTimeBin = t->mintimebin;
        if(oracle_is_timebin_active(mTimeBin + 1, t->Ti_Current)) mTimeBin++;
    }
    t->mintimebin = mTimeBin;                                              /* :727-731 */
    if(t->mintimebin > t->mingravtimebin && t->mingravtimebin > 0) t->mintimebin = t->mingravtimebin;
    return bad;
}

/* get_long_range_timestep_dloga + get_PM_timestep_ti timestep.c:1201-1298 (no neutrinos) */
int64_t oracle_pm_timestep_ti(const oracle_timeline *tl, const oracle_cosmo *c, const oracle_step_params *sp, const oracle_times *t,
                              int64_t n, const double *vel, const float *mass, const uint8_t *type, const uint8_t *flags,
                              double atime, int FastParticleType, double asmth)
{
    int64_t count[6] = {0};
    double v[6] = {0}, mim[6];
    for(int k = 0; k < 6; k++) mim[k] = 1.0e30;
    for(int64_t i = 0; i < n; i++) {
        if(flags && (flags[i] & 3)) continue;
        const int ty = type[i];
        v[ty] += vel[3 * i] * vel[3 * i] + vel[3 * i + 1] * vel[3 * i + 1] + vel[3 * i + 2] * vel[3 * i + 2];
        if(mass[i] > 0 && mim[ty] > mass[i]) mim[ty] = mass[i];
        count[ty]++;
    }
    v[0] += v[4]; count[0] += count[4]; v[4] = v[0]; count[4] = count[0];
    v[0] += v[5]; count[0] += count[5]; v[5] = v[0]; count[5] = count[0];
    mim[5] = mim[0];
    const double hubble = hubble_of(c, atime);
    const double RhoCrit = 3 * c->Hubble * c->Hubble / (8 * M_PI * c->G);
    double dloga = sp->MaxSizeTimestep;
    for(int ty = 0; ty < 6; ty++) {
        if(count[ty] == 0) continue;
        double omega = c->Omega0 - c->OmegaBaryon;
        if(ty == 0 || ty == 4 || ty == 5) omega = c->OmegaBaryon;
        else if(ty == 2) omega = 0;                                        /* get_omega_nu: no neutrinos in the fixtures */
        const double dmean = pow(mim[ty] / (omega * RhoCrit), 1.0 / 3);
        const double dloga1 = sp->MaxRMSDisplacementFac * hubble * atime * atime * (asmth < dmean ? asmth : dmean) / sqrt(v[ty] / count[ty]);
        if(ty != FastParticleType && dloga1 < dloga) dloga = dloga1;
    }
    if(dloga < sp->MinSizeTimestep) dloga = sp->MinSizeTimestep;
    int64_t dti = round_down_pow2(oracle_dti_from_dloga(tl, dloga, t->Ti_Current));
    /* find_next_sync_point timebinmgr.c:335-345: sync point i sits at i << TIMEBINS */
    const int64_t next = ((t->Ti_Current >> TB) + 1) << TB;
    const int64_t dti_max = next - t->PM_kick;
    if(dti > dti_max) dti = dti_max;
    return dti;
}

/* find_timesteps timestep.c:739-853 (ForceEqualTimesteps = 0, no black holes): the smaller of the gravity and (gas)
 * hydro steps -> one bin for TimeBinGravity and TimeBinHydro; PM step length, min / max bin.  Returns the bad-step count. */
int oracle_find_timesteps(const oracle_timeline *tl, const oracle_cosmo *c, const oracle_step_params *sp, oracle_times *t, int64_t n,
                          const int32_t *list, int64_t nlist, const uint8_t *type, const uint8_t *flags, const float *mass, const double *vel,
                          const double *fullacc, const double *gravpm, const double *hsml, const double *dthsml, const double *maxsig,
                          uint8_t *bin_grav, uint8_t *bin_hydro, int is_pm, double atime, int FastParticleType, double asmth)
{
    int64_t dti_max = t->PM_length;
    if(is_pm) {                                                            /* :751-755 */
        dti_max = oracle_pm_timestep_ti(tl, c, sp, t, n, vel, mass, type, flags, atime, FastParticleType, asmth);
        t->PM_length = dti_max;
        t->PM_start = t->PM_kick;
    }
    const double hubble = hubble_of(c, atime);
    const double fac3 = pow(atime, 3 * (1 - 5.0 / 3) / 2.0);
    int bad = 0, mTimeBin = TB, maxTimeBin = 0;
    for(int64_t q = 0; q < nlist; q++) {
        const int64_t i = list ? list[q] : q;
        if(flags && (flags[i] & 3)) continue;
        const double dloga = oracle_gravity_dloga(fullacc + 3 * i, gravpm + 3 * i, atime, hubble, sp->ErrTolIntAccuracy, sp->softening);
        int64_t dti = oracle_convert_timestep(tl, dloga, dti_max, t->Ti_Current, sp->MinSizeTimestep);
        if(type[i] == 0) {                                                 /* :783-791 */
            double dt = 2 * sp->CourantFac * atime * hsml[i] / (fac3 * maxsig[i]);
            const double dt_hsml = sp->CourantFac * atime * atime * fabs(hsml[i] / (dthsml[i] + 1e-20));
            if(dt_hsml < dt) dt = dt_hsml;
            const int64_t dti_hydro = oracle_convert_timestep(tl, dt * hubble, dti_max, t->Ti_Current, sp->MinSizeTimestep);
            if(dti_hydro < dti) dti = dti_hydro;
        }
        int bin = bin_of_dti(round_down_pow2(dti));                        /* get_timebin_from_dti :166-182 */
        const int binold = bin_hydro[i];
        if(bin > binold)
            while(!oracle_is_timebin_active(bin, t->Ti_Current) && bin > binold && bin > 1) bin--;
        if(bin < 1) bad++;
        if(oracle_is_timebin_active(binold, t->Ti_Current) && oracle_is_timebin_active(bin, t->Ti_Current)) { bin_hydro[i] = (uint8_t) bin; bin_grav[i] = (uint8_t) bin; }
        if(bin < mTimeBin) mTimeBin = bin;
        if(bin > maxTimeBin) maxTimeBin = bin;
    }
    if(is_pm && t->PM_length > dti_of_bin(maxTimeBin)) t->PM_length = dti_of_bin(maxTimeBin);      /* :835-836 */
    t->mintimebin = mTimeBin; t->maxtimebin = maxTimeBin;
    return bad;
}

/* ---- the hierarchical gravity drivers (collisionless particles, one rank) ---- */
typedef struct {
    const oracle_timeline *tl; const oracle_cosmo *c; const oracle_step_params *sp; oracle_gravshort_params *gp;
    int64_t n; const double *pos; const float *mass; const uint8_t *type; const uint8_t *flags;
    double *vel, *fullacc; const double *gravpm;
    double G, Asmth, BoxSize; int Nmesh;
} hier_ctx;

/* grav_short_tree_build_tree timestep.c:281-291: tree over the sub-list, walk for the sub-list;
 * store[n][3] receives the accelerations of the listed particles.  Implicit list = full tree:
 * FullTreeGravAccel is refreshed (gravshort.h:55-60). */
static int hier_gravity(hier_ctx *h, const int32_t *list, int64_t nlist, double *store)
{
    oracle_tree T;
    double *old = (double *) malloc(sizeof(double) * 3 * (h->n > 0 ? h->n : 1));
    for(int64_t k = 0; k < 3 * h->n; k++) old[k] = h->fullacc[k] + h->gravpm[k];
    int rc = oracle_tree_build(&T, h->pos, h->mass, h->type, NULL, h->n, h->BoxSize, 63, list, list ? nlist : 0, 0);
    if(!rc) rc = oracle_grav_short_tree(&T, h->pos, h->mass, h->n, h->gp, h->G, h->Nmesh, h->Asmth, old, list, nlist, list == NULL, store, NULL, NULL);
    if(!rc) oracle_tree_free(&T);
    free(old);
    if(h->gp->TreeUseBH > 1) h->gp->TreeUseBH = 0;                        /* gravshort-tree.c:150-151 */
    if(!rc && !list) memcpy(h->fullacc, store, sizeof(double) * 3 * h->n);
    return rc;
}
/* apply_hierarchical_grav_kick timestep.c:238-279 */
static void hier_kick(hier_ctx *h, const oracle_times *t, const int32_t *list, int64_t nlist, const double *acc, int ti, int largest_active)
{
    const int64_t dti = dti_of_bin(ti);
    double gravkick = oracle_step_factor(h->c, h->tl, 1, t->Ti_kick[ti], t->Ti_kick[ti] + dti / 2);
    if(ti < largest_active) {
        const int64_t upper = dti_of_bin(ti + 1);
        gravkick -= oracle_step_factor(h->c, h->tl, 1, t->Ti_kick[ti + 1], t->Ti_kick[ti + 1] + upper / 2);
    }
    oracle_grav_kick(list, nlist, h->flags, h->vel, acc, gravkick);
}
static int largest_active_bin(const oracle_times *t)                      /* timestep.c:311-318,506-513 */
{
    for(int ti = TB; ti >= 0; ti--)
        if(oracle_is_timebin_active(ti, t->Ti_Current) && dti_of_bin(ti) <= t->PM_length) return ti;
    return TB;
}

/* hierarchical_gravity_accelerations timestep.c:503-598.  act: the active list (NULL = all, PM step)
 * with nact entries of which ngrav are gravitationally active.  store[n][3]: StoredGravAccel. */
int oracle_hier_accelerations(const oracle_timeline *tl, const oracle_cosmo *c, const oracle_step_params *sp, oracle_gravshort_params *gp,
                              oracle_times *t, int64_t n, const double *pos, const float *mass, const uint8_t *type, const uint8_t *flags,
                              double *vel, double *fullacc, const double *gravpm, const uint8_t *bin_grav,
                              const int32_t *act, int64_t nact, int64_t ngrav,
                              double G, int Nmesh, double Asmth, double BoxSize, double *store)
{
    hier_ctx h = {tl, c, sp, gp, n, pos, mass, type, flags, vel, fullacc, gravpm, G, Asmth, BoxSize, Nmesh};
    const int largest_active = largest_active_bin(t);
    int32_t *bufA = (int32_t *) malloc(sizeof(int32_t) * (n + 1)), *bufB = (int32_t *) malloc(sizeof(int32_t) * (n + 1));
    const int32_t *last = act;
    int64_t nlast = nact;
    if(ngrav != nact) {                                                    /* :524-528 */
        nlast = oracle_active_sublist(act, nact, flags, bin_grav, largest_active, t->Ti_Current, bufA);
        last = bufA;
    }
    int rc = hier_gravity(&h, last, nlast, store);                         /* :533 */
    if(!rc) hier_kick(&h, t, last, nlast, store, largest_active, largest_active);   /* :537 */
    double *lower = NULL;
    for(int ti = largest_active - 1; !rc && ti >= t->mingravtimebin; ti--) {        /* :550-590 */
        int32_t *dst = (last == bufA) ? bufB : bufA;
        const int64_t nsub = oracle_active_sublist(last, nlast, flags, bin_grav, ti, t->Ti_Current, dst);
        if(nsub != nlast) {                                                /* :564-571: same set, same accelerations */
            if(!lower) lower = (double *) malloc(sizeof(double) * 3 * (n > 0 ? n : 1));
            rc = hier_gravity(&h, dst, nsub, lower);
            if(rc) break;
        }
        hier_kick(&h, t, dst, nsub, lower ? lower : store, ti, largest_active);     /* :578-583 */
        last = dst; nlast = nsub;
    }
    free(lower); free(bufA); free(bufB);
    return rc;
}

/* hierarchical_gravity_and_timesteps timestep.c:296-499.  store = StoredGravAccel of the call above
 * (accelerations of the largest active bin).  bin_grav is updated; returns the bad-step count
 * (< 0: oracle failure).  info = {largest_active after push-down, PM_length}. */
int oracle_hier_timesteps(const oracle_timeline *tl, const oracle_cosmo *c, const oracle_step_params *sp, oracle_gravshort_params *gp,
                          oracle_times *t, int64_t n, const double *pos, const float *mass, const uint8_t *type, const uint8_t *flags,
                          double *vel, double *fullacc, const double *gravpm, uint8_t *bin_grav,
                          const int32_t *act, int64_t nact, int64_t ngrav, int is_pm,
                          double G, int Nmesh, double Asmth, double BoxSize, double atime, int FastParticleType,
                          const double *store, int64_t *info)
{
    hier_ctx h = {tl, c, sp, gp, n, pos, mass, type, flags, vel, fullacc, gravpm, G, Asmth, BoxSize, Nmesh};
    int64_t dti_max = t->PM_length;
    if(is_pm) {                                                            /* :303-309 */
        const double asmth = Asmth * BoxSize / Nmesh;
        dti_max = oracle_pm_timestep_ti(tl, c, sp, t, n, vel, mass, type, flags, atime, FastParticleType, asmth);
        t->PM_length = dti_max;
        t->PM_start = t->PM_kick;
    }
    const double hubble = hubble_of(c, atime);
    int largest_active = largest_active_bin(t);
    int32_t *bufA = (int32_t *) malloc(sizeof(int32_t) * (n + 1)), *bufB = (int32_t *) malloc(sizeof(int32_t) * (n + 1));
    const int32_t *sub = act;
    int64_t nsub = nact;
    if(!(ngrav == nact || is_pm)) {                                        /* :324-328 */
        nsub = oracle_active_sublist(act, nact, flags, bin_grav, largest_active, t->Ti_Current, bufA);
        sub = bufA;
    }
    int64_t counts[TB + 1];
    memset(counts, 0, sizeof(counts));
    for(int64_t q = 0; q < nsub; q++) {                                    /* :346-372 */
        const int64_t i = sub ? sub[q] : q;
        if(flags && (flags[i] & 3)) continue;
        const double *acc = store ? store + 3 * i : fullacc + 3 * i;
        const int bin = oracle_gravity_timebin(tl, acc, gravpm + 3 * i, sp, atime, hubble, dti_max, t->Ti_Current, largest_active);
        counts[bin]++;
        bin_grav[i] = (uint8_t) bin;
    }
    for(int ti = largest_active; ti >= 1; ti--)                            /* :383-387 */
        if(counts[ti] > 0) { largest_active = ti; break; }
    int push_down = largest_active;                                        /* :394-413 */
    if(is_pm)
        for(int ti = largest_active; ti >= 1; ti--) {
            if(counts[ti] / 3 > counts[ti - 1]) break;
            push_down = ti - 1;
            counts[ti - 1] += counts[ti];
        }
    if(push_down == 0) { free(bufA); free(bufB); return -77; }
    if(push_down != largest_active) {
        for(int64_t q = 0; q < nsub; q++) {
            const int64_t i = sub ? sub[q] : q;
            if(bin_grav[i] > push_down) bin_grav[i] = (uint8_t) push_down;
        }
        largest_active = push_down;
    }
    t->maxtimebin = largest_active;                                        /* :415 */
    hier_kick(&h, t, sub, nsub, store ? store : fullacc, largest_active, largest_active);   /* :418 */
    int bad = 0;
    const int32_t *last = sub;
    int64_t nlast = nsub;
    double *lower = (double *) malloc(sizeof(double) * 3 * (n > 0 ? n : 1));
    for(int ti = largest_active - 1; ti > 0; ti--) {                       /* :435-493 */
        int32_t *dst = (last == bufA) ? bufB : bufA;
        const int64_t nnew = oracle_active_sublist(last, nlast, flags, bin_grav, ti, t->Ti_Current, dst);
        if(nnew == 0) { t->mingravtimebin = ti + 1; break; }               /* :443-447 */
        if(hier_gravity(&h, dst, nnew, lower)) { bad = -1; break; }
        for(int64_t q = 0; q < nnew; q++) {                                /* :457-472 */
            const int64_t i = dst[q];
            if(flags && (flags[i] & 3)) continue;
            const double dloga = oracle_gravity_dloga(lower + 3 * i, gravpm + 3 * i, atime, hubble, sp->ErrTolIntAccuracy, sp->softening);
            const int64_t dti = oracle_convert_timestep(tl, dloga, dti_max, t->Ti_Current, sp->MinSizeTimestep);
            if(dti < dti_of_bin(ti)) {
                bin_grav[i] = (uint8_t) (ti - 1);
                if(ti == 1) bad++;
            }
        }
        hier_kick(&h, t, dst, nnew, lower, ti, largest_active);            /* :474 */
        last = dst; nlast = nnew;
    }
    free(lower); free(bufA); free(bufB);
    t->mintimebin = t->mingravtimebin;                                     /* :496 */
    if(info) { info[0] = largest_active; info[1] = t->PM_length; }
    return bad;
}
