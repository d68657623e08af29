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
