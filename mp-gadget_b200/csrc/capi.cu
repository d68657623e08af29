// capi.cu -- the extern "C" surface declared in include/b200force.h.
#include "engine.h"
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <new>
#include <vector>
#include <math.h>

namespace b200 {

int fail(Engine *e, const char *what, cudaError_t err, const char *file, int line)
{
    char buf[512];
    snprintf(buf, sizeof(buf), "%s failed: %s (%s:%d)", what, cudaGetErrorString(err), file, line);
    e->err = buf;
    return 1;
}
int failmsg(Engine *e, const std::string &msg) { e->err = msg; return 1; }

void timer_start(Engine *E, int id)
{
    Timer &t = E->timers[id];
    if(!t.a) { cudaEventCreate(&t.a); cudaEventCreate(&t.b); }
    cudaEventRecord(t.a, E->stream);
    t.used = false;
}
void timer_stop(Engine *E, int id)
{
    Timer &t = E->timers[id];
    cudaEventRecord(t.b, E->stream);
    t.used = true;
}
double timer_ms(Engine *E, int id)
{
    Timer &t = E->timers[id];
    if(!t.used) return 0;
    float ms = 0;
    if(cudaEventSynchronize(t.b) != cudaSuccess) return 0;
    if(cudaEventElapsedTime(&ms, t.a, t.b) != cudaSuccess) return 0;
    return ms;
}

// ---- ingest kernels -------------------------------------------------------
__global__ void __launch_bounds__(256)
k_unpack_aos(const uint8_t *__restrict__ aos, int64_t n, b200_particle_layout L,
             double *__restrict__ pos, float *__restrict__ mass, uint8_t *__restrict__ type,
             uint8_t *__restrict__ flags, double *__restrict__ oldacc)
{
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    const uint8_t *r = aos + i * L.stride;
    const double *p = (const double *) (r + L.off_pos);
    pos[3 * i] = p[0]; pos[3 * i + 1] = p[1]; pos[3 * i + 2] = p[2];
    mass[i] = *(const float *) (r + L.off_mass);
    type[i] = r[L.off_type];
    flags[i] = r[L.off_flags] & 3;
    const double *ft = (const double *) (r + L.off_fulltreeacc);
    const double *pm = (const double *) (r + L.off_gravpm);
    double s = 0;                       // grav_get_abs_accel gravshort.h:69-86
#pragma unroll
    for(int j = 0; j < 3; j++) {
        const double a = __dadd_rn(ft[j], pm[j]);
        s = __dadd_rn(s, __dmul_rn(a, a));
    }
    oldacc[i] = sqrt(s);
}

__global__ void __launch_bounds__(256)
k_set_soa(int64_t n, const uint8_t *__restrict__ type_in, const double *__restrict__ oldacc3,
          uint8_t *__restrict__ type, uint8_t *__restrict__ flags, double *__restrict__ oldacc)
{
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    type[i] = type_in ? type_in[i] : 1;
    flags[i] = 0;
    double s = 0;
    if(oldacc3) {
#pragma unroll
        for(int j = 0; j < 3; j++) { const double a = oldacc3[3 * i + j]; s = __dadd_rn(s, __dmul_rn(a, a)); }
    }
    oldacc[i] = sqrt(s);
}

__global__ void __launch_bounds__(256)
k_oldacc_from_last(int64_t n, const double *__restrict__ tr, const double *__restrict__ pm, double *__restrict__ oldacc)
{
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    double s = 0;
#pragma unroll
    for(int j = 0; j < 3; j++) {
        const double a = __dadd_rn(tr ? tr[3 * i + j] : 0.0, pm ? pm[3 * i + j] : 0.0);
        s = __dadd_rn(s, __dmul_rn(a, a));
    }
    oldacc[i] = sqrt(s);
}

// Write GravPM, FullTreeGravAccel, Potential back into the AoS staging copy.
__global__ void __launch_bounds__(256)
k_pack_aos(uint8_t *__restrict__ aos, int64_t n, b200_particle_layout L,
           const double *__restrict__ gravpm, const double *__restrict__ treeacc,
           const double *__restrict__ pot, int full_tree, int64_t first)
{
    const int64_t i = first + (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= first + n) return;
    uint8_t *r = aos + i * L.stride;
    double *pm = (double *) (r + L.off_gravpm);
    pm[0] = gravpm[3 * i]; pm[1] = gravpm[3 * i + 1]; pm[2] = gravpm[3 * i + 2];
    if(full_tree) {
        double *ft = (double *) (r + L.off_fulltreeacc);
        ft[0] = treeacc[3 * i]; ft[1] = treeacc[3 * i + 1]; ft[2] = treeacc[3 * i + 2];
        // TREEWALK_REDUCE assigns in primary mode (treewalk.h:202), so the tree
        // potential replaces the PM potential accumulated earlier in the step.
        *(double *) (r + L.off_potential) = pot[i];
    }
}

// The strided ingest of b200_force_step_aos: the caller's records are never copied whole.  Span A = the bytes that hold
// Pos, Mass, the flag byte and Type (what the PM step and the tree build read), span B = FullTreeGravAccel..GravPM
// (the walk's opening criterion reads |sum|, and both are outputs).  o_* are offsets inside the span.
__global__ void __launch_bounds__(256)
k_unpack_span_a(const uint8_t *__restrict__ a, int64_t n, int w, int o_pos, int o_mass, int o_flags, int o_type,
                double *__restrict__ pos, float *__restrict__ mass, uint8_t *__restrict__ type, uint8_t *__restrict__ flags)
{
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    const uint8_t *r = a + i * w;
    const double *p = (const double *) (r + o_pos);
    pos[3 * i] = p[0]; pos[3 * i + 1] = p[1]; pos[3 * i + 2] = p[2];
    mass[i] = *(const float *) (r + o_mass);
    type[i] = r[o_type];
    flags[i] = r[o_flags] & 3;
}
__global__ void __launch_bounds__(256)
k_unpack_span_b(const uint8_t *__restrict__ b, int64_t n, int w, int o_ft, int o_pm, double *__restrict__ oldacc)
{
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    const double *ft = (const double *) (b + i * w + o_ft), *pm = (const double *) (b + i * w + o_pm);
    double s = 0;                       // grav_get_abs_accel gravshort.h:69-86
#pragma unroll
    for(int j = 0; j < 3; j++) {
        const double v = __dadd_rn(ft[j], pm[j]);
        s = __dadd_rn(s, __dmul_rn(v, v));
    }
    oldacc[i] = sqrt(s);
}
__global__ void __launch_bounds__(256)
k_pack_span_b(uint8_t *__restrict__ b, int64_t n, int w, int o_ft, int o_pm, const double *__restrict__ gravpm,
              const double *__restrict__ treeacc, int64_t first)
{
    const int64_t i = first + (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= first + n) return;
    double *pm = (double *) (b + i * w + o_pm), *ft = (double *) (b + i * w + o_ft);
    pm[0] = gravpm[3 * i]; pm[1] = gravpm[3 * i + 1]; pm[2] = gravpm[3 * i + 2];
    ft[0] = treeacc[3 * i]; ft[1] = treeacc[3 * i + 1]; ft[2] = treeacc[3 * i + 2];
}

static int ensure_particles(Engine *E, int64_t n)
{
    const size_t m = (size_t) (n > 0 ? n : 1);
    CK(E->pos.ensure(3 * m)); CK(E->mass.ensure(m)); CK(E->type.ensure(m));
    CK(E->flags.ensure(m)); CK(E->oldacc.ensure(m));
    E->n = n;
    E->tree_valid = false;
    E->potential_valid = false;
    E->have_last_tree = E->have_last_pm = false;
    return 0;
}

int collect_timings(Engine *E)
{
    b200_timings &t = E->last;
    t.pm_deposit = timer_ms(E, T_PM_DEPOSIT); t.pm_fft_forward = timer_ms(E, T_PM_FFT_FWD);
    t.pm_transfer = timer_ms(E, T_PM_TRANSFER); t.pm_fft_inverse = timer_ms(E, T_PM_FFT_INV);
    t.pm_gradient = timer_ms(E, T_PM_GRADIENT); t.pm_readout = timer_ms(E, T_PM_READOUT);
    t.pm_total = t.pm_deposit + t.pm_fft_forward + t.pm_transfer + t.pm_fft_inverse + t.pm_gradient + t.pm_readout;
    t.tree_keys = timer_ms(E, T_TREE_KEYS); t.tree_sort = timer_ms(E, T_TREE_SORT);
    t.tree_nodes = timer_ms(E, T_TREE_NODES); t.tree_moments = timer_ms(E, T_TREE_MOMENTS);
    t.tree_total = t.tree_keys + t.tree_sort + t.tree_nodes + t.tree_moments;
    t.walk = timer_ms(E, T_WALK); t.walk_post = timer_ms(E, T_WALK_POST);
    t.h2d = timer_ms(E, T_H2D); t.d2h = timer_ms(E, T_D2H);
    t.sph_density = timer_ms(E, T_SPH_DENSITY); t.sph_hydro = timer_ms(E, T_SPH_HYDRO);
    t.walk_pieces = E->walk_pieces; t.walk_list_bytes = 4.0 * 32 * (1 << 4) * (double) E->walk_chunks;
    return 0;
}

} // namespace b200

using namespace b200;

extern "C" {

void b200_default_particle_layout(b200_particle_layout *L)
{
    // struct particle_data, libgadget/partmanager.h:9-71 (no -DDEBUG): 160 bytes
    L->stride = 160; L->off_pos = 0; L->off_mass = 28; L->off_pi = 32; L->off_flags = 36;
    L->off_timebin_hydro = 37; L->off_timebin_gravity = 38; L->off_type = 39;
    L->off_vel = 40; L->off_fulltreeacc = 64; L->off_gravpm = 88; L->off_hsml = 120; L->off_potential = 152;
}

int b200_abi_version(void) { return B200_ABI_VERSION; }

int b200_ctx_create(b200_ctx **out, int device)
{
    if(!out) return 1;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if(e != cudaSuccess || ndev == 0) {
        fprintf(stderr, "b200_ctx_create: no CUDA device (%s); this engine has no CPU fallback\n",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
        return 2;
    }
    if(device < 0 || device >= ndev) { fprintf(stderr, "b200_ctx_create: bad device %d of %d\n", device, ndev); return 3; }
    if(cudaSetDevice(device) != cudaSuccess) return 4;
    b200_ctx *c = new (std::nothrow) b200_ctx();
    if(!c) return 5;
    c->e.device = device;
    if(cudaStreamCreateWithFlags(&c->e.stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return 6; }
    if(cudaStreamCreateWithFlags(&c->e.copy_stream, cudaStreamNonBlocking) != cudaSuccess) { cudaStreamDestroy(c->e.stream); delete c; return 6; }
    {   // the side stream (PM step running beside the tree walk) outranks the main stream: its HBM-bound kernels are
        // placed as soon as walk blocks retire instead of queueing behind the walk's whole grid
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        if(cudaStreamCreateWithPriority(&c->e.side_stream, cudaStreamNonBlocking, hi) != cudaSuccess) return 6;
    }
    if(cudaEventCreateWithFlags(&c->e.fork_ev, cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&c->e.join_ev, cudaEventDisableTiming) != cudaSuccess) return 6;
    *out = c;
    return 0;
}

void b200_ctx_destroy(b200_ctx *ctx)
{
    if(!ctx) return;
    Engine *E = &ctx->e;
    cudaSetDevice(E->device);
    cudaStreamSynchronize(E->stream);
    pm_destroy(E);
    pmslab_destroy(E);
    E->pos.release(); E->mass.release(); E->type.release(); E->flags.release(); E->oldacc.release();
    E->last_tree_acc.release(); E->last_pm_acc.release(); E->aos.release(); E->pm_ps.release();
    E->s_hD.release(); E->s_bins.release(); E->s_bin_grav.release(); E->s_bin_hydro.release(); E->s_active.release();
    E->sph_list_a.release(); E->sph_list_b.release(); E->s_left.release(); E->s_right.release();
    E->keys.release(); E->keys_alt.release(); E->sidx.release(); E->sidx_alt.release(); E->cubtemp.release();
    E->spart.release(); E->spart_xy.release(); E->spart_zm.release(); E->b_start.release(); E->b_count.release(); E->b_father.release(); E->b_sibling.release();
    E->b_firstchild.release(); E->b_nchild.release(); E->b_level.release(); E->b_size.release(); E->b_dfs.release();
    E->b_scan.release(); E->b_center.release(); E->nodeA.release(); E->nodeB.release(); E->nodeC.release();
    E->nodeF.release(); E->nodeH.release(); E->nodeK.release(); E->scratch_i.release(); E->targets.release(); E->targets_sorted.release(); E->walk_flags.release();
    E->d_acc.release(); E->d_pot.release(); E->d_counts.release(); E->srtab.release();
    E->walk_pool.release(); E->walk_chunktab.release(); E->walk_cnt.release(); E->walk_partial.release();
    step_release(E); domain_release(E); fof_release(E);
    // the SPH state and scratch (grow-only buffers have no destructor: everything is released here)
    E->s_vel.release(); E->s_hsml.release(); E->s_entropy.release(); E->s_dtentropy.release(); E->s_fullacc.release(); E->s_gravpm.release();
    E->s_hydroacc.release(); E->s_velpred.release(); E->s_evp.release(); E->s_density.release(); E->s_egy.release(); E->s_dhsmlfac.release();
    E->s_divvel.release(); E->s_curlvel.release(); E->s_dthsml.release(); E->s_numngb.release(); E->s_gradrho.release(); E->s_svel.release();
    E->s_hA.release(); E->s_hB.release(); E->s_out3.release(); E->s_out1a.release(); E->s_out1b.release(); E->s_outi.release(); E->s_outi2.release();
    E->s_niter.release(); E->s_nint.release();
    for(int i = 0; i < T_COUNT; i++) if(E->timers[i].a) { cudaEventDestroy(E->timers[i].a); cudaEventDestroy(E->timers[i].b); }
    for(int i = 0; i < 66; i++) if(E->chunk_ev[i]) cudaEventDestroy(E->chunk_ev[i]);
    cudaStreamDestroy(E->copy_stream);
    sharded_destroy(E);
    cudaStreamDestroy(E->side_stream);
    cudaEventDestroy(E->fork_ev); cudaEventDestroy(E->join_ev);
    cudaStreamDestroy(E->stream);
    delete ctx;
}

const char *b200_last_error(const b200_ctx *ctx) { return ctx ? ctx->e.err.c_str() : "null context"; }
int64_t b200_kernel_launches(const b200_ctx *ctx) { return ctx ? ctx->e.launches : 0; }
void *b200_stream(const b200_ctx *ctx) { return ctx ? (void *) ctx->e.stream : nullptr; }

#define ENTER(ctx) if(!(ctx)) return 1; Engine *E = &(ctx)->e; if(cudaSetDevice(E->device) != cudaSuccess) return failmsg(E, "cudaSetDevice failed");

int b200_set_particles_aos(b200_ctx *ctx, const void *P, int64_t n, const b200_particle_layout *layout)
{
    ENTER(ctx);
    if(n < 0 || (n > 0 && !P)) return failmsg(E, "b200_set_particles_aos: bad arguments");
    b200_particle_layout L; if(layout) L = *layout; else b200_default_particle_layout(&L);
    if(int rc = ensure_particles(E, n)) return rc;
    if(n == 0) return 0;
    CK(E->aos.ensure((size_t) n * L.stride));
    timer_start(E, T_H2D);
    CK(cudaMemcpyAsync(E->aos.p, P, (size_t) n * L.stride, cudaMemcpyHostToDevice, E->stream));
    timer_stop(E, T_H2D);
    k_unpack_aos<<<(unsigned) ((n + 255) / 256), 256, 0, E->stream>>>(E->aos.p, n, L, E->pos.p, E->mass.p, E->type.p, E->flags.p, E->oldacc.p);
    CKL(E);
    return 0;
}

static int set_soa_common(Engine *E, const double *pos, const float *mass, const uint8_t *type,
                          const double *oldacc, int64_t n, cudaMemcpyKind kind)
{
    if(n < 0 || (n > 0 && (!pos || !mass))) return failmsg(E, "b200_set_particles_soa: bad arguments");
    if(int rc = ensure_particles(E, n)) return rc;
    if(n == 0) return 0;
    timer_start(E, T_H2D);
    CK(cudaMemcpyAsync(E->pos.p, pos, 3 * n * sizeof(double), kind, E->stream));
    CK(cudaMemcpyAsync(E->mass.p, mass, n * sizeof(float), kind, E->stream));
    const uint8_t *d_type = nullptr; const double *d_old = nullptr;
    if(type) {
        if(kind == cudaMemcpyHostToDevice) { CK(E->aos.ensure((size_t) n)); CK(cudaMemcpyAsync(E->aos.p, type, n, kind, E->stream)); d_type = E->aos.p; }
        else d_type = type;
    }
    if(oldacc) {
        if(kind == cudaMemcpyHostToDevice) {
            CK(E->last_tree_acc.ensure(3 * (size_t) n));
            CK(cudaMemcpyAsync(E->last_tree_acc.p, oldacc, 3 * n * sizeof(double), kind, E->stream));
            d_old = E->last_tree_acc.p;
        } else d_old = oldacc;
    }
    timer_stop(E, T_H2D);
    k_set_soa<<<(unsigned) ((n + 255) / 256), 256, 0, E->stream>>>(n, d_type, d_old, E->type.p, E->flags.p, E->oldacc.p);
    CKL(E);
    return 0;
}

int b200_set_particles_soa(b200_ctx *ctx, const double *pos, const float *mass, const uint8_t *type, const double *oldacc, int64_t n)
{
    ENTER(ctx);
    return set_soa_common(E, pos, mass, type, oldacc, n, cudaMemcpyHostToDevice);
}

int b200_set_particles_soa_dev(b200_ctx *ctx, const double *pos, const float *mass, const uint8_t *type, const double *oldacc, int64_t n)
{
    ENTER(ctx);
    return set_soa_common(E, pos, mass, type, oldacc, n, cudaMemcpyDeviceToDevice);
}

int b200_oldacc_from_last_step(b200_ctx *ctx)
{
    ENTER(ctx);
    if(!E->have_last_tree && !E->have_last_pm) return failmsg(E, "b200_oldacc_from_last_step: no previous accelerations");
    if(E->n == 0) return 0;
    k_oldacc_from_last<<<(unsigned) ((E->n + 255) / 256), 256, 0, E->stream>>>(E->n, E->have_last_tree ? E->last_tree_acc.p : nullptr,
                                                                          E->have_last_pm ? E->last_pm_acc.p : nullptr, E->oldacc.p);
    CKL(E);
    return 0;
}

int b200_walk_set_mesh(b200_ctx *ctx, double BoxSize, double Asmth, int Nmesh, double G)
{
    ENTER(ctx);
    if(!(BoxSize > 0) || !(Asmth > 0) || Nmesh < 2) return failmsg(E, "b200_walk_set_mesh: bad arguments");
    E->Box = BoxSize; E->Asmth = Asmth; E->G = G; E->NmeshWalk = Nmesh;
    return 0;
}

int b200_pm_init(b200_ctx *ctx, double BoxSize, double Asmth, int Nmesh, double G)
{
    ENTER(ctx);
    return pm_init(E, BoxSize, Asmth, Nmesh, G);
}

static int pm_force_common(Engine *E, double *gravpm_out, double *potential_out, bool host)
{
    const size_t n = (size_t) (E->n > 0 ? E->n : 1);
    CK(E->last_pm_acc.ensure(3 * n));
    double *d_g = E->last_pm_acc.p, *d_p = nullptr;
    if(potential_out) {
        if(host) { CK(E->d_pot.ensure(n)); d_p = E->d_pot.p; } else d_p = potential_out;
    }
    if(int rc = pm_force(E, d_g, d_p)) return rc;
    E->have_last_pm = true;
    if(E->n > 0) {
        timer_start(E, T_D2H);
        if(gravpm_out) CK(cudaMemcpyAsync(gravpm_out, d_g, 3 * E->n * sizeof(double), host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, E->stream));
        if(potential_out && host) CK(cudaMemcpyAsync(potential_out, d_p, E->n * sizeof(double), cudaMemcpyDeviceToHost, E->stream));
        timer_stop(E, T_D2H);
    }
    CK(cudaStreamSynchronize(E->stream));
    return collect_timings(E);
}

int b200_pm_force(b200_ctx *ctx, double *gravpm_out, double *potential_out) { ENTER(ctx); return pm_force_common(E, gravpm_out, potential_out, true); }
int b200_pm_force_dev(b200_ctx *ctx, double *gravpm_out, double *potential_out) { ENTER(ctx); return pm_force_common(E, gravpm_out, potential_out, false); }

int b200_fof_primary(b200_ctx *ctx, const int64_t *ids, int primary_mask, double BoxSize, double linking_length,
                     int64_t *minid_out, int64_t *ngroups_out)
{
    ENTER(ctx);
    return fof_primary(E, ids, primary_mask, BoxSize, linking_length, minid_out, ngroups_out);
}

int b200_pm_c2r_readout(b200_ctx *ctx, const double *rho_k, int nfunc, const b200_pm_function *functions)
{
    ENTER(ctx);
    if(E->Nmesh == 0) return failmsg(E, "b200_pm_c2r_readout: call b200_pm_init first");
    if(int rc = pmfft_c2r_readout(E, rho_k, nfunc, functions)) return rc;
    return collect_timings(E);
}

int b200_pm_transform_kind(b200_ctx *ctx)
{
    if(!ctx) return -1;
    Engine *E = &ctx->e;
    return E->Nmesh == 0 || (!E->ownfft && !E->plans) ? -1 : (E->ownfft ? 1 : 0);
}

int b200_pm_set_power(b200_ctx *ctx, int on)
{
    ENTER(ctx);
    E->pm_power = on != 0;
    if(!on) E->pm_ps_valid = false;
    return 0;
}

int b200_pm_get_power(b200_ctx *ctx, int nbins, double *power, double *kk, int64_t *nmodes, double *norm)
{
    ENTER(ctx);
    if(!E->pm_ps_valid) return failmsg(E, "b200_pm_get_power: no spectrum (b200_pm_set_power(1), then b200_pm_force)");
    if(nbins != E->Nmesh) return failmsg(E, "b200_pm_get_power: nbins must equal Nmesh (powerspectrum_alloc, gravpm.c:207)");
    std::vector<double> h(3 * (size_t) nbins + 1);
    CK(cudaMemcpyAsync(h.data(), E->pm_ps.p, h.size() * sizeof(double), cudaMemcpyDeviceToHost, E->stream));
    CK(cudaStreamSynchronize(E->stream));
    for(int b = 0; b < nbins; b++) {
        if(power) power[b] = h[b];
        if(kk) kk[b] = h[nbins + b];
        if(nmodes) nmodes[b] = (int64_t) llround(h[2 * (size_t) nbins + b]);
    }
    if(norm) *norm = h[3 * (size_t) nbins];
    return 0;
}

int b200_pm_cell_index(b200_ctx *ctx, int32_t *icell_out)
{
    ENTER(ctx);
    if(E->n == 0) return 0;
    CK(E->d_counts.ensure(3 * (size_t) E->n));
    if(int rc = pm_cell_index(E, E->d_counts.p)) return rc;
    CK(cudaMemcpyAsync(icell_out, E->d_counts.p, 3 * E->n * sizeof(int32_t), cudaMemcpyDeviceToHost, E->stream));
    CK(cudaStreamSynchronize(E->stream));
    return 0;
}

int b200_pm_copy_mesh(b200_ctx *ctx, int which, double *mesh_out)
{
    ENTER(ctx);
    if(E->Nmesh == 0) return failmsg(E, "b200_pm_copy_mesh: call b200_pm_init first");
    const size_t N3 = (size_t) E->Nmesh * E->Nmesh * E->Nmesh;
    const double *src = nullptr;
    if(which == 0) { if(int rc = pm_deposit(E)) return rc; src = E->mesh.p; }
    else if(which == 1) { if(!E->potential_valid) return failmsg(E, "b200_pm_copy_mesh: no potential (call b200_pm_force)"); src = E->mesh.p; }
    else if(which >= 2 && which <= 4) {
        if(!E->potential_valid) return failmsg(E, "b200_pm_copy_mesh: no force mesh");
        if(!E->fmesh_valid) if(int rc = pm_force_meshes(E)) return rc;
        src = E->fmesh.p + (which - 2) * N3;
    }
    else return failmsg(E, "b200_pm_copy_mesh: which must be 0..4");
    CK(cudaMemcpyAsync(mesh_out, src, N3 * sizeof(double), cudaMemcpyDeviceToHost, E->stream));
    CK(cudaStreamSynchronize(E->stream));
    return 0;
}

int b200_tree_build(b200_ctx *ctx, double BoxSize, int mask, const int32_t *active, int64_t nactive,
                    int toplevel_depth, b200_tree_info *info)
{
    ENTER(ctx);
    const int32_t *d_active = nullptr;
    if(active) {
        CK(E->targets.ensure((size_t) (nactive > 0 ? nactive : 1)));
        CK(cudaMemcpyAsync(E->targets.p, active, nactive * sizeof(int32_t), cudaMemcpyHostToDevice, E->stream));
        d_active = E->targets.p;
    }
    int rc = tree_build(E, BoxSize, mask, d_active, nactive, toplevel_depth, info);
    if(rc) return rc;
    CK(cudaStreamSynchronize(E->stream));
    return collect_timings(E);
}

void b200_tree_free(b200_ctx *ctx) { if(ctx) ctx->e.tree_valid = false; }

int b200_tree_export(b200_ctx *ctx, double *center, double *len, double *cofm, double *mass, double *hmax,
                     int32_t *sibling, int32_t *firstchild, int32_t *nocc, int32_t *leafpart)
{
    ENTER(ctx);
    return tree_export(E, center, len, cofm, mass, hmax, sibling, firstchild, nocc, leafpart);
}

static int grav_common(Engine *E, const b200_gravshort_params *par, const int32_t *active, int64_t nactive,
                       double *accel_out, double *potential_out, b200_walk_counts *counts_out, bool host)
{
    if(!par) return failmsg(E, "b200_grav_short_tree: null params");
    const size_t n = (size_t) (E->n > 0 ? E->n : 1);
    const int32_t *d_active = nullptr;
    if(active) {
        if(host) {
            CK(E->targets.ensure((size_t) (nactive > 0 ? nactive : 1)));
            CK(cudaMemcpyAsync(E->targets.p, active, nactive * sizeof(int32_t), cudaMemcpyHostToDevice, E->stream));
            d_active = E->targets.p;
        } else d_active = active;
    }
    CK(E->last_tree_acc.ensure(3 * n));
    double *d_a = E->last_tree_acc.p, *d_p = nullptr;
    b200_walk_counts *d_c = nullptr;
    if(potential_out) { if(host) { CK(E->d_pot.ensure(n)); d_p = E->d_pot.p; } else d_p = potential_out; }
    if(counts_out) {
        if(host) { CK(E->d_counts.ensure(4 * n)); CK(cudaMemsetAsync(E->d_counts.p, 0, 4 * n * sizeof(int), E->stream)); d_c = (b200_walk_counts *) E->d_counts.p; }
        else d_c = counts_out;
    }
    if(int rc = grav_short_tree(E, par, d_active, nactive, d_a, d_p, d_c)) return rc;
    E->have_last_tree = (active == nullptr);
    if(E->n > 0) {
        timer_start(E, T_D2H);
        const cudaMemcpyKind k = host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
        if(accel_out) {
            if(active == nullptr) CK(cudaMemcpyAsync(accel_out, d_a, 3 * E->n * sizeof(double), k, E->stream));
            else if(host) {
                // only the active entries are defined (GravShortPriv.Accel semantics, gravshort.h:41)
                std::vector<double> tmp(3 * (size_t) E->n);
                CK(cudaMemcpyAsync(tmp.data(), d_a, 3 * E->n * sizeof(double), k, E->stream));
                CK(cudaStreamSynchronize(E->stream));
                for(int64_t q = 0; q < nactive; q++) { const int64_t i = active[q]; for(int j = 0; j < 3; j++) accel_out[3 * i + j] = tmp[3 * i + j]; }
            } else CK(cudaMemcpyAsync(accel_out, d_a, 3 * E->n * sizeof(double), k, E->stream));
        }
        if(potential_out && host) CK(cudaMemcpyAsync(potential_out, d_p, E->n * sizeof(double), k, E->stream));
        if(counts_out && host) CK(cudaMemcpyAsync(counts_out, d_c, E->n * sizeof(b200_walk_counts), k, E->stream));
        timer_stop(E, T_D2H);
    }
    CK(cudaStreamSynchronize(E->stream));
    return collect_timings(E);
}

int b200_grav_short_tree(b200_ctx *ctx, const b200_gravshort_params *par, const int32_t *active, int64_t nactive,
                         double *accel_out, double *potential_out, b200_walk_counts *counts_out)
{
    ENTER(ctx);
    return grav_common(E, par, active, nactive, accel_out, potential_out, counts_out, true);
}

int b200_grav_short_tree_dev(b200_ctx *ctx, const b200_gravshort_params *par, const int32_t *active, int64_t nactive,
                             double *accel_out, double *potential_out, b200_walk_counts *counts_out)
{
    ENTER(ctx);
    return grav_common(E, par, active, nactive, accel_out, potential_out, counts_out, false);
}

// gravpm_force issued on the side stream, concurrent with whatever follows on the
// main stream until pm_join(): the PM kernels are HBM-bound, the tree walk is
// latency/issue-bound and reads none of the PM results, so the two overlap on the SMs.
static int pm_force_forked(Engine *E, double *d_gravpm, double *d_pot)
{
    CK(cudaEventRecord(E->fork_ev, E->stream));
    CK(cudaStreamWaitEvent(E->side_stream, E->fork_ev, 0));
    cudaStream_t main_stream = E->stream;
    E->stream = E->side_stream;
    if(E->plans) { cufftSetStream(E->plan_fwd, E->stream); cufftSetStream(E->plan_inv, E->stream); }
    const int rc = pm_force(E, d_gravpm, d_pot);
    cudaEventRecord(E->join_ev, E->stream);
    E->stream = main_stream;
    if(E->plans) { cufftSetStream(E->plan_fwd, E->stream); cufftSetStream(E->plan_inv, E->stream); }
    return rc;
}
static int pm_join(Engine *E)
{
    CK(cudaStreamWaitEvent(E->stream, E->join_ev, 0));
    return 0;
}

/* Device-resident form of the same step: particles already set
 * (b200_set_particles_*), outputs are device pointers (any may be NULL). */
int b200_force_step_dev(b200_ctx *ctx, const b200_gravshort_params *par, double *gravpm_out, double *pm_potential_out,
                        double *accel_out, double *potential_out)
{
    ENTER(ctx);
    if(E->Nmesh == 0) return failmsg(E, "b200_force_step_dev: call b200_pm_init first");
    if(!par) return failmsg(E, "b200_force_step_dev: null params");
    const size_t m = (size_t) (E->n > 0 ? E->n : 1);
    CK(E->last_pm_acc.ensure(3 * m)); CK(E->last_tree_acc.ensure(3 * m)); CK(E->d_pot.ensure(m));
    // The cuFFT-based PM step runs on the side stream next to the tree build and walk (no gain, no loss: 132.8 against
    // 133.0 ms at 256^3).  The engine's own transform passes hold 2 x 111 KB of shared memory per SM and slow the walk
    // down by more than they hide (130.2 ms forked, 126.4 ms one after the other), so they run in stream order.
    const char *fk = getenv("B200_PM_FORK");
    const bool fork = fk ? atoi(fk) != 0 : !E->ownfft;
    if(fork) { if(int rc = pm_force_forked(E, E->last_pm_acc.p, pm_potential_out)) return rc; }
    else if(int rc = pm_force(E, E->last_pm_acc.p, pm_potential_out)) return rc;
    if(int rc = tree_build(E, E->Box, 63, nullptr, 0, 0, nullptr)) return rc;
    if(int rc = grav_short_tree(E, par, nullptr, 0, E->last_tree_acc.p, potential_out ? potential_out : E->d_pot.p, nullptr)) return rc;
    if(fork) if(int rc = pm_join(E)) return rc;
    E->have_last_pm = E->have_last_tree = true;
    if(E->n > 0) {
        if(gravpm_out) CK(cudaMemcpyAsync(gravpm_out, E->last_pm_acc.p, 3 * E->n * sizeof(double), cudaMemcpyDeviceToDevice, E->stream));
        if(accel_out) CK(cudaMemcpyAsync(accel_out, E->last_tree_acc.p, 3 * E->n * sizeof(double), cudaMemcpyDeviceToDevice, E->stream));
    }
    CK(cudaStreamSynchronize(E->stream));
    return collect_timings(E);
}

// Spans of the strided ingest / write-back (see k_unpack_span_a).  Returns false when the layout does not fit the
// scheme (fields not 8-byte aligned inside the spans, spans too wide to pay): the caller then moves whole records.
static bool aos_spans(const b200_particle_layout &L, int &a0, int &aw, int &b0, int &bw)
{
    auto lo = [](int x, int y) { return x < y ? x : y; };
    auto hi = [](int x, int y) { return x > y ? x : y; };
    a0 = lo(lo(L.off_pos, L.off_mass), lo(L.off_flags, L.off_type));
    const int a1 = hi(hi(L.off_pos + 24, L.off_mass + 4), hi(L.off_flags + 1, L.off_type + 1));
    a0 &= ~7; aw = ((a1 + 7) & ~7) - a0;
    b0 = lo(L.off_fulltreeacc, L.off_gravpm);
    const int b1 = hi(L.off_fulltreeacc, L.off_gravpm) + 24;
    bw = b1 - b0;
    if((b0 & 7) || (bw & 7) || (L.off_pos & 7) || (L.off_potential & 7) || (L.stride & 7)) return false;
    return aw + bw <= (int) (0.7 * L.stride);
}

/* Bytes per particle b200_force_step_aos moves over PCIe in each direction for this layout. */
void b200_force_step_aos_bytes(const b200_particle_layout *layout, int64_t *h2d, int64_t *d2h)
{
    b200_particle_layout L; if(layout) L = *layout; else b200_default_particle_layout(&L);
    int a0, aw, b0, bw;
    const bool strided = aos_spans(L, a0, aw, b0, bw) && !getenv("B200_E2E_BULK");
    if(h2d) *h2d = strided ? aw + bw : L.stride;
    if(d2h) *d2h = strided ? bw + 8 : L.stride;
}

int b200_force_step_aos(b200_ctx *ctx, void *P, int64_t n, const b200_particle_layout *layout,
                        const b200_gravshort_params *par)
{
    ENTER(ctx);
    if(E->Nmesh == 0) return failmsg(E, "b200_force_step_aos: call b200_pm_init first");
    if(!par) return failmsg(E, "b200_force_step_aos: null params");
    b200_particle_layout L; if(layout) L = *layout; else b200_default_particle_layout(&L);
    int a0 = 0, aw = 0, b0 = 0, bw = 0;
    const bool strided = aos_spans(L, a0, aw, b0, bw) && !getenv("B200_E2E_BULK");
    uint8_t *spanB = nullptr;
    if(!strided) { if(int rc = b200_set_particles_aos(ctx, P, n, &L)) return rc; }
    else {
        // Two strided copies instead of the whole records (160 B/particle): span A (Pos .. Type, 40 B) first -- the PM step
        // and the tree build start as soon as it is unpacked -- and span B (FullTreeGravAccel, GravPM, 48 B), which only
        // the walk needs, on the copy stream underneath them.
        if(n < 0 || (n > 0 && !P)) return failmsg(E, "b200_force_step_aos: bad arguments");
        if(int rc = ensure_particles(E, n)) return rc;
        if(n > 0) {
            CK(E->aos.ensure((size_t) n * (aw + bw)));
            uint8_t *spanA = E->aos.p; spanB = E->aos.p + (size_t) n * aw;
            timer_start(E, T_H2D);
            CK(cudaMemcpy2DAsync(spanA, aw, (const uint8_t *) P + a0, L.stride, aw, n, cudaMemcpyHostToDevice, E->stream));
            timer_stop(E, T_H2D);
            k_unpack_span_a<<<(unsigned) ((n + 255) / 256), 256, 0, E->stream>>>(spanA, n, aw, L.off_pos - a0, L.off_mass - a0, L.off_flags - a0,
                                                                               L.off_type - a0, E->pos.p, E->mass.p, E->type.p, E->flags.p);
            CKL(E);
            if(!E->chunk_ev[65]) CK(cudaEventCreateWithFlags(&E->chunk_ev[65], cudaEventDisableTiming));
            CK(cudaEventRecord(E->chunk_ev[65], E->stream));                // span A is in: the link is free for span B
            CK(cudaStreamWaitEvent(E->copy_stream, E->chunk_ev[65], 0));
            CK(cudaMemcpy2DAsync(spanB, bw, (const uint8_t *) P + b0, L.stride, bw, n, cudaMemcpyHostToDevice, E->copy_stream));
            CK(cudaEventRecord(E->chunk_ev[65], E->copy_stream));
        }
    }
    if(n == 0) return 0;
    const size_t m = (size_t) n;
    CK(E->last_pm_acc.ensure(3 * m)); CK(E->last_tree_acc.ensure(3 * m)); CK(E->d_pot.ensure(m));
    // gravpm_force (run.c:519-523), concurrent with the tree build and walk
    if(int rc = pm_force_forked(E, E->last_pm_acc.p, nullptr)) return rc;
    // force_tree_full + grav_short_tree (run.c:546-548)
    if(int rc = tree_build(E, E->Box, 63, nullptr, 0, 0, nullptr)) return rc;
    E->have_last_pm = E->have_last_tree = true;
    if(strided) {
        CK(cudaStreamWaitEvent(E->stream, E->chunk_ev[65], 0));
        k_unpack_span_b<<<(unsigned) ((n + 255) / 256), 256, 0, E->stream>>>(spanB, n, bw, L.off_fulltreeacc - b0, L.off_gravpm - b0, E->oldacc.p);
        CKL(E);
    }
    // Results of a group of particles: into the caller's records.  Strided path: span B (packed in place, so bytes of
    // the span that are not outputs keep the caller's values) and the potential column; otherwise the whole records.
    auto pack = [&](int64_t lo, int64_t hi) -> int {
        if(strided) k_pack_span_b<<<(unsigned) ((hi - lo + 255) / 256), 256, 0, E->stream>>>(spanB, hi - lo, bw, L.off_fulltreeacc - b0, L.off_gravpm - b0,
                                                                                           E->last_pm_acc.p, E->last_tree_acc.p, lo);
        else k_pack_aos<<<(unsigned) ((hi - lo + 255) / 256), 256, 0, E->stream>>>(E->aos.p, hi - lo, L, E->last_pm_acc.p, E->last_tree_acc.p, E->d_pot.p, 1, lo);
        CKL(E);
        return 0;
    };
    auto copy_back = [&](int64_t lo, int64_t hi, cudaStream_t st) -> int {
        if(strided) {
            CK(cudaMemcpy2DAsync((uint8_t *) P + lo * L.stride + b0, L.stride, spanB + lo * bw, bw, bw, hi - lo, cudaMemcpyDeviceToHost, st));
            // TREEWALK_REDUCE assigns in primary mode (treewalk.h:202): the tree potential replaces the PM potential
            CK(cudaMemcpy2DAsync((uint8_t *) P + lo * L.stride + L.off_potential, L.stride, E->d_pot.p + lo, sizeof(double), sizeof(double), hi - lo,
                                 cudaMemcpyDeviceToHost, st));
        } else CK(cudaMemcpyAsync((uint8_t *) P + lo * L.stride, E->aos.p + lo * L.stride, (size_t) (hi - lo) * L.stride, cudaMemcpyDeviceToHost, st));
        return 0;
    };
    // The walk is issued in groups of equal particle-index ranges (targets in curve
    // order inside each group): as soon as a group is done its results are packed and
    // copied back on the copy stream while the next group is walked, so the write-back
    // hides behind the walk.  B200_E2E_CHUNKS overrides.
    int nchunks = n >= (1 << 20) ? 8 : 1;
    if(const char *ev = getenv("B200_E2E_CHUNKS")) { nchunks = atoi(ev); if(nchunks < 1) nchunks = 1; if(nchunks > 64) nchunks = 64; }
    if(nchunks == 1) {
        if(int rc = grav_short_tree(E, par, nullptr, 0, E->last_tree_acc.p, E->d_pot.p, nullptr)) return rc;
        if(int rc = pm_join(E)) return rc;
        if(int rc = pack(0, n)) return rc;
        timer_start(E, T_D2H);
        if(int rc = copy_back(0, n, E->stream)) return rc;
        timer_stop(E, T_D2H);
    } else {
        const int64_t chunk = (n + nchunks - 1) / nchunks;
        int off[65];
        if(int rc = walk_chunk_targets(E, nchunks, chunk, off)) return rc;
        // particles outside the tree (garbage) keep their input record: pack only touches tree members' ranges
        for(int c = 0; c < nchunks; c++) {
            const int64_t lo = c * chunk, hi = (lo + chunk < n) ? lo + chunk : n;
            if(hi <= lo) break;
            const int64_t nc = off[c + 1] - off[c];
            if(nc > 0)
                if(int rc = grav_short_tree(E, par, E->targets_sorted.p + off[c], nc, E->last_tree_acc.p, E->d_pot.p, nullptr, true)) return rc;
            if(c == 0) if(int rc = pm_join(E)) return rc;       // GravPM is packed with the first group
            if(int rc = pack(lo, hi)) return rc;
            if(c == nchunks - 1 || hi == n) timer_start(E, T_D2H);      // the exposed tail of the write-back
            if(!E->chunk_ev[c]) CK(cudaEventCreateWithFlags(&E->chunk_ev[c], cudaEventDisableTiming));
            CK(cudaEventRecord(E->chunk_ev[c], E->stream));
            CK(cudaStreamWaitEvent(E->copy_stream, E->chunk_ev[c], 0));
            if(int rc = copy_back(lo, hi, E->copy_stream)) return rc;
        }
        if(!E->chunk_ev[nchunks]) CK(cudaEventCreateWithFlags(&E->chunk_ev[nchunks], cudaEventDisableTiming));
        CK(cudaEventRecord(E->chunk_ev[nchunks], E->copy_stream));
        CK(cudaStreamWaitEvent(E->stream, E->chunk_ev[nchunks], 0));
        timer_stop(E, T_D2H);
    }
    CK(cudaStreamSynchronize(E->stream));
    return collect_timings(E);
}

int b200_sph_set_gas(b200_ctx *ctx, const double *vel, const double *hsml, const double *entropy, const double *dtentropy,
                     const double *fulltreeacc, const double *gravpm, const double *hydroaccel)
{
    ENTER(ctx);
    return sph_set_gas(E, vel, hsml, entropy, dtentropy, fulltreeacc, gravpm, hydroaccel);
}

int b200_sph_set_timebins(b200_ctx *ctx, const uint8_t *timebin_gravity, const uint8_t *timebin_hydro, const b200_sph_bins *bins)
{
    ENTER(ctx);
    return sph_set_timebins(E, timebin_gravity, timebin_hydro, bins);
}
int b200_sph_set_hsml_range(b200_ctx *ctx, const double *hsml, int64_t first, int64_t count)
{
    ENTER(ctx);
    if(int rc = sph_set_hsml_range(E, hsml, first, count)) return rc;
    CK(cudaStreamSynchronize(E->stream));
    return 0;
}
int b200_sph_set_active(b200_ctx *ctx, const int32_t *active, int64_t nactive) { ENTER(ctx); return sph_set_active(E, active, nactive); }
int b200_sph_set_state(b200_ctx *ctx, const double *density, const double *egywtdensity, const double *dhsmlfac,
                       const double *divvel, const double *curlvel)
{
    ENTER(ctx);
    return sph_set_state(E, density, egywtdensity, dhsmlfac, divvel, curlvel);
}

static int d2h_opt(Engine *E, void *dst, const void *src, size_t bytes)       // dst: host or device memory (UVA)
{
    if(dst && bytes) CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, E->stream));
    return 0;
}

int b200_density(b200_ctx *ctx, const b200_sph_params *par, int update_hsml, int DoEgyDensity,
                 double *hsml, double *density, double *egywtdensity, double *dhsmlfac, double *divvel, double *curlvel,
                 double *dthsml, double *numngb, int32_t *ninteract, int32_t *niter)
{
    ENTER(ctx);
    const size_t n = (size_t) (E->n > 0 ? E->n : 1);
    CK(E->s_outi.ensure(n)); CK(E->s_outi2.ensure(n));
    CK(cudaMemsetAsync(E->s_outi.p, 0, n * sizeof(int), E->stream));
    CK(cudaMemsetAsync(E->s_outi2.p, 0, n * sizeof(int), E->stream));
    if(int rc = sph_density(E, par, update_hsml, DoEgyDensity, E->s_outi.p, E->s_outi2.p)) return rc;
    const size_t b = (size_t) E->n * sizeof(double);
    if(d2h_opt(E, hsml, E->s_hsml.p, b) || d2h_opt(E, density, E->s_density.p, b) || d2h_opt(E, egywtdensity, E->s_egy.p, b) ||
       d2h_opt(E, dhsmlfac, E->s_dhsmlfac.p, b) || d2h_opt(E, divvel, E->s_divvel.p, b) || d2h_opt(E, curlvel, E->s_curlvel.p, b) ||
       d2h_opt(E, dthsml, E->s_dthsml.p, b) || d2h_opt(E, numngb, E->s_numngb.p, b) ||
       d2h_opt(E, ninteract, E->s_outi.p, (size_t) E->n * sizeof(int)) || d2h_opt(E, niter, E->s_outi2.p, (size_t) E->n * sizeof(int)))
        return 1;
    CK(cudaStreamSynchronize(E->stream));
    return collect_timings(E);
}

int b200_density_gradrho(b200_ctx *ctx, double *gradrho)
{
    ENTER(ctx);
    if(!E->sph_density_done) return failmsg(E, "b200_density_gradrho: call b200_density first");
    if(d2h_opt(E, gradrho, E->s_gradrho.p, 3 * (size_t) E->n * sizeof(double))) return 1;
    CK(cudaStreamSynchronize(E->stream));
    return 0;
}

int b200_hydro_force(b200_ctx *ctx, const b200_sph_params *par, double *hydroaccel, double *dtentropy, double *maxsignalvel,
                     int32_t *ninteract)
{
    ENTER(ctx);
    const size_t n = (size_t) (E->n > 0 ? E->n : 1);
    CK(E->s_out3.ensure(3 * n)); CK(E->s_out1a.ensure(n)); CK(E->s_out1b.ensure(n)); CK(E->s_outi.ensure(n));
    CK(cudaMemsetAsync(E->s_out3.p, 0, 3 * n * sizeof(double), E->stream));
    CK(cudaMemsetAsync(E->s_out1a.p, 0, n * sizeof(double), E->stream));
    CK(cudaMemsetAsync(E->s_out1b.p, 0, n * sizeof(double), E->stream));
    CK(cudaMemsetAsync(E->s_outi.p, 0, n * sizeof(int), E->stream));
    if(int rc = sph_hydro(E, par, E->s_out3.p, E->s_out1a.p, E->s_out1b.p, E->s_outi.p)) return rc;
    const size_t b = (size_t) E->n * sizeof(double);
    if(d2h_opt(E, hydroaccel, E->s_out3.p, 3 * b) || d2h_opt(E, dtentropy, E->s_out1a.p, b) || d2h_opt(E, maxsignalvel, E->s_out1b.p, b) ||
       d2h_opt(E, ninteract, E->s_outi.p, (size_t) E->n * sizeof(int)))
        return 1;
    CK(cudaStreamSynchronize(E->stream));
    return collect_timings(E);
}

int b200_tree_top_get_dev(b200_ctx *ctx, int level, double *cells_out)
{
    ENTER(ctx);
    if(int rc = tree_top_get(E, level, cells_out)) return rc;
    CK(cudaStreamSynchronize(E->stream));
    return 0;
}

int b200_tree_top_set_dev(b200_ctx *ctx, int level, const double *cells_in)
{
    ENTER(ctx);
    if(int rc = tree_top_set(E, level, cells_in)) return rc;
    CK(cudaStreamSynchronize(E->stream));
    return 0;
}

int b200_get_timings(const b200_ctx *ctx, b200_timings *t)
{
    if(!ctx || !t) return 1;
    *t = ctx->e.last;
    return 0;
}

} // extern "C"
