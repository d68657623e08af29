// fft_driver.cpp -- host entry that runs mp-gadget_b200/csrc/pm_fft.cu's passes under the CPU emulation
// (tests/emul, TEST INFRASTRUCTURE ONLY; see include/cuda_runtime.h): density mesh in, potential mesh out.
#include <new>
#include "engine.h"

emul_dim emul_blockIdx, emul_blockDim, emul_gridDim;
double emul_xchg[1024];
unsigned char emul_dyn_smem[256 * 1024] __attribute__((aligned(64)));

namespace b200 {
int fail(Engine *e, const char *what, cudaError_t, const char *file, int line)
{
    char buf[512];
    snprintf(buf, sizeof(buf), "%s failed (%s:%d)", what, file, line);
    e->err = buf;
    return 1;
}
int failmsg(Engine *e, const std::string &msg) { e->err = msg; return 1; }
void timer_start(Engine *, int) {}
void timer_stop(Engine *, int) {}
} // namespace b200

using namespace b200;
extern "C" int emul_pmfft_supported(int N) { return pmfft_supported(N) ? 1 : 0; }
extern "C" int emul_pmfft_potential(int N, double *mesh, const double *ktab, double asmth2, double pot_factor, double binsperunit, double *ps)
{
    Engine *E = new Engine();
    int rc = pmfft_init(E, N);
    if(!rc) {
        const size_t N3 = (size_t) N * N * N;
        if(E->mesh.ensure(N3) || E->cplx.ensure(pmfft_cplx_doubles(E)) || E->ktab.ensure(N)) rc = 2;
        else {
            memcpy(E->mesh.p, mesh, N3 * sizeof(double));
            memcpy(E->ktab.p, ktab, N * sizeof(double));
            rc = pmfft_potential(E, asmth2, pot_factor, binsperunit, ps);
            memcpy(mesh, E->mesh.p, N3 * sizeof(double));
        }
    }
    if(rc) fprintf(stderr, "emul_pmfft_potential: %s\n", E->err.c_str());
    pmfft_destroy(E);
    E->mesh.release(); E->cplx.release(); E->ktab.release();
    delete E;
    return rc;
}

extern "C" int emul_pm_c2r_readout(int N, double box, int64_t n, const double *pos, const double *rho_k, int nfunc, const int *kind,
                                   const double *tables, int64_t nk2, double *out)
{
    Engine *E = new Engine();
    int rc = pmfft_init(E, N);
    const size_t m = (size_t) (n > 0 ? n : 1);
    if(!rc) {
        if(E->mesh.ensure((size_t) N * N * N) || E->cplx.ensure(pmfft_cplx_doubles(E)) || E->ktab.ensure(N) || E->pos.ensure(3 * m) || E->flags.ensure(m)) rc = 2;
        else {
            E->n = n; E->Box = box; E->Nmesh = N;
            memcpy(E->pos.p, pos, 3 * n * sizeof(double));
            memset(E->flags.p, 0, m);
            b200_pm_function f[16];
            for(int j = 0; j < nfunc; j++) { f[j].kind = kind[j]; f[j].table = tables + (size_t) j * nk2; f[j].out = out + (size_t) j * n; }
            rc = pmfft_c2r_readout(E, rho_k, nfunc, f);
        }
    }
    if(rc) fprintf(stderr, "emul_pm_c2r_readout: %s\n", E->err.c_str());
    pmfft_destroy(E);
    E->mesh.release(); E->cplx.release(); E->ktab.release(); E->pos.release(); E->flags.release(); E->pm_rhok.release(); E->pm_table.release(); E->d_pot.release();
    delete E;
    return rc;
}
