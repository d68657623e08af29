// emul_mocks.cpp -- what steploop.cu needs from the rest of the engine, for the CPU emulation build
// (tests/emul, TEST INFRASTRUCTURE ONLY; see include/cuda_runtime.h).  The tree build and the
// short-range walk -- CUDA kernels verified on hardware elsewhere -- are replaced by the oracle here,
// so the emulation exercises exactly the unverified part: steploop.cu's kernels and host drivers.
#include <new>
#include <vector>
#include "engine.h"
#include "../../oracle/oracle.h"

emul_dim emul_blockIdx, emul_blockDim, emul_gridDim;
double emul_xchg[1024];

namespace b200 {
int fail(Engine *e, const char *what, cudaError_t, const char *file, int line)
{
    char buf[512];
    snprintf(buf, sizeof(buf), "%s failed (%s:%d)", what, file, line);
    e->err = buf;
    return 1;
}
int failmsg(Engine *e, const std::string &msg) { e->err = msg; return 1; }

static std::vector<int32_t> tree_list;
static bool tree_is_full;
int tree_build(Engine *E, double Box, int mask, const int32_t *d_active, int64_t nactive, int, b200_tree_info *)
{
    tree_list.clear();
    tree_is_full = d_active == nullptr;
    const int64_t nin = d_active ? nactive : E->n;
    for(int64_t q = 0; q < nin; q++) {
        const int64_t i = d_active ? d_active[q] : q;
        if((E->flags.p[i] & 3) || !((mask >> E->type.p[i]) & 1)) continue;
        tree_list.push_back((int32_t) i);
    }
    if(E->sidx.ensure(tree_list.size() + 1) != cudaSuccess) return failmsg(E, "emul: out of memory");
    memcpy(E->sidx.p, tree_list.data(), tree_list.size() * sizeof(int32_t));
    E->tree_np = (int64_t) tree_list.size(); E->tree_box = Box; E->tree_valid = true; E->tree_full = tree_is_full;
    return 0;
}
int grav_short_tree(Engine *E, const b200_gravshort_params *par, const int32_t *d_active, int64_t nactive, double *d_acc, double *, b200_walk_counts *, bool)
{
    oracle_tree T;
    if(oracle_tree_build(&T, E->pos.p, E->mass.p, E->type.p, NULL, E->n, E->tree_box, 63, tree_is_full ? NULL : tree_list.data(),
                         tree_is_full ? 0 : (int64_t) tree_list.size(), 0)) return failmsg(E, "emul: oracle_tree_build failed");
    std::vector<double> old(3 * (size_t) E->n, 0.0);
    for(int64_t i = 0; i < E->n; i++) old[3 * i] = E->oldacc.p[i];          // the walk only uses the modulus
    oracle_gravshort_params gp;
    memcpy(&gp, par, sizeof(gp));
    static_assert(sizeof(oracle_gravshort_params) == sizeof(b200_gravshort_params), "parameter structs differ");
    const int rc = oracle_grav_short_tree(&T, E->pos.p, E->mass.p, E->n, &gp, E->G, E->NmeshWalk, E->Asmth, old.data(), d_active, nactive,
                                          tree_is_full, d_acc, NULL, NULL);
    oracle_tree_free(&T);
    return rc ? failmsg(E, "emul: oracle_grav_short_tree failed") : 0;
}
} // namespace b200

using namespace b200;
extern "C" {
int b200_ctx_create(b200_ctx **out, int device) { *out = new (std::nothrow) b200_ctx(); return *out ? 0 : 5; }
void b200_ctx_destroy(b200_ctx *ctx) { if(ctx) { step_release(&ctx->e); domain_release(&ctx->e); delete ctx; } }
const char *b200_last_error(const b200_ctx *ctx) { return ctx ? ctx->e.err.c_str() : "null context"; }
int b200_set_particles_soa(b200_ctx *ctx, const double *pos, const float *mass, const uint8_t *type, const double *, int64_t n)
{
    Engine *E = &ctx->e;
    const size_t m = (size_t) (n > 0 ? n : 1);
    E->n = n;
    if(E->pos.ensure(3 * m) || E->mass.ensure(m) || E->type.ensure(m) || E->flags.ensure(m) || E->oldacc.ensure(m)) return failmsg(E, "emul: out of memory");
    memcpy(E->pos.p, pos, 3 * n * sizeof(double)); memcpy(E->mass.p, mass, n * sizeof(float));
    for(int64_t i = 0; i < n; i++) { E->type.p[i] = type ? type[i] : 1; E->flags.p[i] = 0; E->oldacc.p[i] = 0; }
    return 0;
}
int b200_set_particles_aos(b200_ctx *ctx, const void *P, int64_t n, const b200_particle_layout *)
{
    // struct particle_data with the default layout (b200_default_particle_layout in csrc/capi.cu): stride 160,
    // Pos 0, Mass 28, flag bits 36, Type 39
    Engine *E = &ctx->e;
    const size_t m = (size_t) (n > 0 ? n : 1);
    E->n = n;
    if(E->pos.ensure(3 * m) || E->mass.ensure(m) || E->type.ensure(m) || E->flags.ensure(m) || E->oldacc.ensure(m)) return failmsg(E, "emul: out of memory");
    for(int64_t i = 0; i < n; i++) {
        const uint8_t *r = (const uint8_t *) P + 160 * i;
        memcpy(E->pos.p + 3 * i, r, 24); memcpy(E->mass.p + i, r + 28, 4);
        E->flags.p[i] = r[36] & 3; E->type.p[i] = r[39]; E->oldacc.p[i] = 0;
    }
    return 0;
}
int b200_pm_init(b200_ctx *ctx, double BoxSize, double Asmth, int Nmesh, double G)
{
    Engine *E = &ctx->e;
    E->Box = BoxSize; E->Asmth = Asmth; E->Nmesh = Nmesh; E->NmeshWalk = Nmesh; E->G = G;
    return 0;
}
}
