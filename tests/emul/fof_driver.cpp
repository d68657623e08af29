// fof_driver.cpp -- host entry that runs mp-gadget_b200/csrc/fof.cu under the CPU emulation (tests/emul, TEST
// INFRASTRUCTURE ONLY; see include/cuda_runtime.h).
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
} // namespace b200

using namespace b200;
extern "C" int emul_fof_primary(int64_t n, const double *pos, const uint8_t *type, const uint8_t *flags, const int64_t *ids, int mask,
                                double box, double ll, int64_t *minid, int64_t *ngroups)
{
    Engine *E = new Engine();
    const size_t m = (size_t) (n > 0 ? n : 1);
    int rc = 2;
    if(!E->pos.ensure(3 * m) && !E->type.ensure(m) && !E->flags.ensure(m)) {
        E->n = n;
        memcpy(E->pos.p, pos, 3 * n * sizeof(double));
        memcpy(E->type.p, type, n);
        for(int64_t i = 0; i < n; i++) E->flags.p[i] = flags ? flags[i] : 0;
        rc = fof_primary(E, ids, mask, box, ll, minid, ngroups);
    }
    if(rc) fprintf(stderr, "emul_fof_primary: %s\n", E->err.c_str());
    fof_release(E);
    E->pos.release(); E->type.release(); E->flags.release(); E->scratch_i.release(); E->cubtemp.release();
    delete E;
    return rc;
}
