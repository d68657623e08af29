// sharded.cu -- the TreePM force step of one rank of a multi-GPU run, issued from C on the engine's streams with
// NCCL.  One process and one engine per GPU.
//
// What the reference does with MPI, and what happens here:
//   * domain cut (domain.c:154-256, domain.h:71-78): rank r owns the x-layers [r*2^d/W, (r+1)*2^d/W) of the cells of a
//     uniform forced top tree of depth d on the reference's Peano lattice (utils/peano.h:15-21).
//   * short-range tree (treewalk.c:325-371,399-793 export queries; here SURVEY 8e option 2, ghost import): the
//     particles of the cell layer adjacent to the domain are imported from both neighbours (ncclSend/ncclRecv), the
//     same forced top tree + complete subtrees are built over own + ghost particles, the level-d cell moments are
//     summed over ranks (ncclAllReduce = force_exchange_pseudodata, forcetree.c:1156-1208) and the upper levels re-summed
//     (force_treeupdate_pseudos :1214-1284); the rank then walks its own particles.  A cell that was not imported is more
//     than one cell width (> Rcut) from every own particle and is discarded by shall_we_discard_node whatever it holds.
//   * PM (petapm.c:263-379 with its pencil exchange :584-885 and PFFT's transposes): x-slab mesh (pm_slab.cu).
//     deposit -> halo planes added into the neighbours -> batched 2-D cuFFT -> block swap -> all-to-all -> strided 1-D
//     cuFFT in place on the receive buffer -> Green's function -> inverse 1-D -> all-to-all straight from that buffer ->
//     block swap -> inverse 2-D -> halo planes fetched -> fused difference + readout.
// The PM chain runs on its own stream (own communicator) while the main stream builds the tree and walks: the
// collectives of the PM step (NVLink-bound) overlap the walk (issue/latency-bound).  One host synchronisation per step
// (the ghost counts) besides those inside tree_build / grav_short_tree; none inside the PM chain.
#include "pm_slab.h"
#include <nccl.h>
#include <dlfcn.h>
#include <math.h>
#include <string.h>
#include <stdlib.h>
#include <cub/device/device_select.cuh>
#include <thrust/iterator/counting_iterator.h>

namespace b200 {

// NCCL is taken from the process at run time (the library torch already loaded, else the system one), so that
// libb200force.so has no link-time dependency on it and single-GPU hosts never touch it.
struct NcclApi {
    void *h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    std::string err;
};

static NcclApi *nccl_api()
{
    static NcclApi A;
    static bool tried = false;
    if(tried) return A.h ? &A : nullptr;
    tried = true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for(const char *nm : names) { A.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if(A.h) break; }
    if(!A.h) { A.err = "libnccl.so.2 not found"; return nullptr; }
#define LD(field, sym) do { *(void **) (&A.field) = dlsym(A.h, sym); if(!A.field) { A.err = std::string("NCCL symbol missing: ") + sym; A.h = nullptr; return nullptr; } } while(0)
    LD(GetUniqueId, "ncclGetUniqueId"); LD(CommInitRank, "ncclCommInitRank"); LD(CommDestroy, "ncclCommDestroy");
    LD(Send, "ncclSend"); LD(Recv, "ncclRecv"); LD(AllReduce, "ncclAllReduce"); LD(GroupStart, "ncclGroupStart");
    LD(GroupEnd, "ncclGroupEnd"); LD(GetErrorString, "ncclGetErrorString");
#undef LD
    return &A;
}

#define NCK(call) do { ncclResult_t _r = (call); if(_r != ncclSuccess) return failmsg(E, std::string(#call) + ": " + NC->GetErrorString(_r)); } while(0)

enum { SH_EV_GHOST0, SH_EV_GHOST1, SH_EV_PM0, SH_EV_DEP, SH_EV_HADD, SH_EV_F2D, SH_EV_PACK, SH_EV_A2A1, SH_EV_F1D, SH_EV_A2A2,
       SH_EV_UNPACK, SH_EV_I2D, SH_EV_HFILL, SH_EV_READ, SH_EV_TOP0, SH_EV_TOP1, SH_EV_COUNT };

struct Sharded {
    ncclComm_t comm_tree = nullptr, comm_pm = nullptr;
    int rank = 0, world = 1, left = 0, right = 0;
    bool configured = false;
    double box = 0, asmth = 0, G = 0, rcut_cells = 0;
    int nmesh = 0, d = 0, halo = 0;
    int ncell = 1, per = 1, lo = 0, hi = 1;
    double domainfac = 0, shift = 0;
    cudaStream_t pm_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaEvent_t ev[SH_EV_COUNT] = {};
    DevBuf<uint8_t> flagL, flagR;
    DevBuf<int> idxL, idxR, iota, err;
    DevBuf<long long> cnt;                     // [0] to left [1] to right [2] from right [3] from left
    DevBuf<double> sendL, sendR, recvL, recvR;  // double4 rows {x, y, z, m}
    DevBuf<double> xbuf;                        // block-swapped spectrum (send side forward, receive side backward)
    DevBuf<double> haloA, haloB;
    DevBuf<double> top;
    int64_t n_own = 0, n_tot = 0;
};

void sharded_destroy(Engine *E)
{
    Sharded *S = E->sh;
    if(!S) return;
    NcclApi *NC = nccl_api();
    if(NC) { if(S->comm_tree) NC->CommDestroy(S->comm_tree); if(S->comm_pm) NC->CommDestroy(S->comm_pm); }
    if(S->pm_stream) cudaStreamDestroy(S->pm_stream);
    if(S->ev_fork) cudaEventDestroy(S->ev_fork);
    if(S->ev_join) cudaEventDestroy(S->ev_join);
    for(auto &e : S->ev) if(e) cudaEventDestroy(e);
    S->flagL.release(); S->flagR.release(); S->idxL.release(); S->idxR.release(); S->iota.release(); S->err.release(); S->cnt.release();
    S->sendL.release(); S->sendR.release(); S->recvL.release(); S->recvR.release(); S->xbuf.release(); S->haloA.release(); S->haloB.release();
    S->top.release();
    delete S;
    E->sh = nullptr;
}

static Sharded *sharded_get(Engine *E)
{
    if(!E->sh) E->sh = new Sharded();
    return E->sh;
}

// ---- kernels ----------------------------------------------------------------------------------------------------

// Layer (top-cell x index) of every own particle on the reference's Peano lattice, exactly as tree_build places
// particles below the forced top tree; flags of the particles the left / right neighbour must import.
__global__ void __launch_bounds__(256)
k_sh_ghost_flags(const double *__restrict__ pos, int64_t n, double shift, double domainfac, int sh, int lo, int hi, int world,
                 uint8_t *__restrict__ fl, uint8_t *__restrict__ fr, int *__restrict__ err)
{
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    const long long ix = ((long long) __dmul_rn(__dadd_rn(pos[3 * i], shift), domainfac)) >> sh;
    if(ix < lo || ix >= hi) atomicAdd(err, 1);
    bool l = ix == lo, r = ix == hi - 1;
    if(world == 2) { l = l || r; r = false; }       // both neighbours are the same rank: it must receive each particle once
    if(world == 1) { l = false; r = false; }
    fl[i] = l ? 1 : 0; fr[i] = r ? 1 : 0;
}

__global__ void __launch_bounds__(256)
k_sh_pack(const double *__restrict__ pos, const float *__restrict__ mass, const int *__restrict__ idx, const long long *__restrict__ cnt,
          double4 *__restrict__ out)
{
    const long long k = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    if(k >= *cnt) return;
    const int j = idx[k];
    out[k] = make_double4(pos[3 * (size_t) j], pos[3 * (size_t) j + 1], pos[3 * (size_t) j + 2], (double) mass[j]);
}

__global__ void __launch_bounds__(256)
k_sh_unpack(const double4 *__restrict__ in, int64_t n, int64_t first, double *__restrict__ pos, float *__restrict__ mass)
{
    const int64_t k = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if(k >= n) return;
    const double4 v = in[k];
    const size_t i = (size_t) (first + k);
    pos[3 * i] = v.x; pos[3 * i + 1] = v.y; pos[3 * i + 2] = v.z; mass[i] = (float) v.w;
}

// type / flags / |old acceleration| of own + ghost particles (ghosts are never walk targets)
__global__ void __launch_bounds__(256)
k_sh_state(int64_t n_tot, int64_t n_own, const double *__restrict__ old3, uint8_t *__restrict__ type, uint8_t *__restrict__ flags,
           double *__restrict__ oldacc, int *__restrict__ iota)
{
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n_tot) return;
    type[i] = 1; flags[i] = 0;
    double s = 0;
    if(old3 && i < n_own)
        for(int j = 0; j < 3; j++) { const double a = old3[3 * i + j]; s = __dadd_rn(s, __dmul_rn(a, a)); }
    oldacc[i] = sqrt(s);
    if(i < n_own) iota[i] = (int) i;
}

// level-d cells outside my x-layers carry no moments of mine (Morton index: x bit 0 of each octal digit)
__global__ void __launch_bounds__(256)
k_sh_zero_foreign(double4 *__restrict__ top, int ncell3, int d, int lo, int hi)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if(m >= ncell3) return;
    int ix = 0;
    for(int l = 0; l < d; l++) ix |= ((m >> (3 * l)) & 1) << l;
    if(ix < lo || ix >= hi) top[m] = make_double4(0, 0, 0, 0);
}

// in [A][B][chunk] -> out [B][A][chunk] (chunk contiguous complex values): the only data movement the slab
// transposes need besides the all-to-all itself.
__global__ void __launch_bounds__(256)
k_sh_swap_blocks(const double2 *__restrict__ in, double2 *__restrict__ out, int A, int B, size_t chunk)
{
    const size_t total = (size_t) A * B * chunk;
    for(size_t idx = (size_t) blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t) gridDim.x * blockDim.x) {
        const size_t c = idx % chunk, blk = idx / chunk;
        const size_t b = blk % B, a = blk / B;
        out[(b * A + a) * chunk + c] = in[idx];
    }
}

__global__ void __launch_bounds__(256)
k_sh_add(double *__restrict__ dst, const double *__restrict__ src, size_t n)
{
    for(size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) dst[i] += src[i];
}

// ---- host ---------------------------------------------------------------------------------------------------------

static int sh_neighbour_exchange(Engine *E, Sharded *S, ncclComm_t comm, cudaStream_t st, const double *send_left, size_t nl,
                                 const double *send_right, size_t nr, double *recv_from_right, size_t nfr, double *recv_from_left, size_t nfl)
{
    if(S->world == 1) {      // my own neighbour on both sides
        if(nl) CK(cudaMemcpyAsync(recv_from_right, send_left, nl * sizeof(double), cudaMemcpyDeviceToDevice, st));
        if(nr) CK(cudaMemcpyAsync(recv_from_left, send_right, nr * sizeof(double), cudaMemcpyDeviceToDevice, st));
        return 0;
    }
    NcclApi *NC = nccl_api();
    NCK(NC->GroupStart());
    if(nl) NCK(NC->Send(send_left, nl, ncclDouble, S->left, comm, st));
    if(nr) NCK(NC->Send(send_right, nr, ncclDouble, S->right, comm, st));
    if(nfr) NCK(NC->Recv(recv_from_right, nfr, ncclDouble, S->right, comm, st));
    if(nfl) NCK(NC->Recv(recv_from_left, nfl, ncclDouble, S->left, comm, st));
    NCK(NC->GroupEnd());
    return 0;
}

// blocks of `blk` doubles: block p of send goes to rank p, block p of recv comes from rank p
static int sh_all_to_all(Engine *E, Sharded *S, cudaStream_t st, const double *send, double *recv, size_t blk)
{
    if(S->world == 1) { CK(cudaMemcpyAsync(recv, send, blk * sizeof(double), cudaMemcpyDeviceToDevice, st)); return 0; }
    NcclApi *NC = nccl_api();
    CK(cudaMemcpyAsync(recv + (size_t) S->rank * blk, send + (size_t) S->rank * blk, blk * sizeof(double), cudaMemcpyDeviceToDevice, st));
    NCK(NC->GroupStart());
    for(int k = 1; k < S->world; k++) {
        const int to = (S->rank + k) % S->world, from = (S->rank - k + S->world) % S->world;
        NCK(NC->Send(send + (size_t) to * blk, blk, ncclDouble, to, S->comm_pm, st));
        NCK(NC->Recv(recv + (size_t) from * blk, blk, ncclDouble, from, S->comm_pm, st));
    }
    NCK(NC->GroupEnd());
    return 0;
}

int sharded_comm_init(Engine *E, int rank, int world, const void *id_tree, const void *id_pm)
{
    if(world < 1 || rank < 0 || rank >= world) return failmsg(E, "b200_comm_init: bad rank / world");
    Sharded *S = sharded_get(E);
    S->rank = rank; S->world = world; S->left = (rank - 1 + world) % world; S->right = (rank + 1) % world;
    if(world == 1) return 0;
    NcclApi *NC = nccl_api();
    if(!NC) return failmsg(E, "b200_comm_init: NCCL is not available in this process");
    if(!id_tree || !id_pm) return failmsg(E, "b200_comm_init: two unique ids are required (b200_comm_unique_id on rank 0, broadcast by the host)");
    ncclUniqueId a, b;
    memcpy(&a, id_tree, sizeof(a)); memcpy(&b, id_pm, sizeof(b));
    NCK(NC->CommInitRank(&S->comm_tree, world, a, rank));
    NCK(NC->CommInitRank(&S->comm_pm, world, b, rank));
    return 0;
}

int sharded_init(Engine *E, double Box, double Asmth, int Nmesh, double G, int topdepth, int halo, double rcut_cells)
{
    Sharded *S = sharded_get(E);
    if(S->world > 1 && (!S->comm_tree || !S->comm_pm)) return failmsg(E, "b200_sharded_init: call b200_comm_init first");
    if(topdepth < 1 || topdepth > 8) return failmsg(E, "b200_sharded_init: top-tree depth must be 1..8");
    const int ncell = 1 << topdepth;
    if(ncell % S->world) return failmsg(E, "b200_sharded_init: 2^topdepth must be a multiple of the number of ranks");
    const double cellwidth = 1.001 * Box / ncell;
    if(S->world > 1 && rcut_cells > 0) {
        // The imported layer must cover everything within Rcut of an own particle.  The tree's root cell is 1.001 Box wide
        // (forcetree.c:662-664), so across the periodic seam the last layer holds particles over only
        // cellwidth - 0.001 Box: that is the width that has to exceed Rcut.
        const double rcut = rcut_cells * Asmth * Box / Nmesh;
        if(!(cellwidth - 0.001 * Box > rcut * 1.0001))
            return failmsg(E, "b200_sharded_init: top-tree cells (" + std::to_string(cellwidth) + " wide, " + std::to_string(cellwidth - 0.001 * Box) +
                              " across the periodic seam) must be wider than Rcut (" + std::to_string(rcut) + "): lower topdepth");
    }
    if(S->world > 1 && !(rcut_cells > 0)) return failmsg(E, "b200_sharded_init: rcut_cells (TreeRcut) is required to validate the ghost layer");
    S->box = Box; S->asmth = Asmth; S->G = G; S->nmesh = Nmesh; S->d = topdepth; S->halo = halo; S->rcut_cells = rcut_cells;
    S->ncell = ncell; S->per = ncell / S->world; S->lo = S->rank * S->per; S->hi = S->lo + S->per;
    S->domainfac = 1.0 / (Box * 1.001) * 2097152.0;           // PEANO(), utils/peano.h:15-21
    S->shift = Box / 2000;
    if(int rc = pmslab_init(E, Box, Asmth, Nmesh, G, S->rank, S->world, halo, nullptr, nullptr, nullptr)) return rc;
    SlabPM *P = E->slab;
    const size_t N = Nmesh;
    CK(S->xbuf.ensure(2 * (size_t) P->nx * N * P->Nz));
    CK(S->haloA.ensure((size_t) halo * N * N)); CK(S->haloB.ensure((size_t) halo * N * N));
    size_t ncell3 = 1; for(int l = 0; l < topdepth; l++) ncell3 *= 8;
    CK(S->top.ensure(4 * ncell3));
    CK(S->cnt.ensure(8)); CK(S->err.ensure(4));
    CK(cudaMemsetAsync(S->err.p, 0, 4 * sizeof(int), E->stream));
    if(!S->pm_stream) {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        CK(cudaStreamCreateWithPriority(&S->pm_stream, cudaStreamNonBlocking, hi));
        CK(cudaEventCreateWithFlags(&S->ev_fork, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&S->ev_join, cudaEventDisableTiming));
        for(auto &e : S->ev) CK(cudaEventCreate(&e));
    }
    CK(cudaStreamSynchronize(E->stream));
    S->configured = true;
    return 0;
}

static int sh_select(Engine *E, const uint8_t *flags, int64_t n, int *out, long long *d_cnt)
{
    thrust::counting_iterator<int> it(0);
    size_t tb = 0;
    cub::DeviceSelect::Flagged(nullptr, tb, it, flags, out, d_cnt, (int) n, E->stream);
    CK(E->cubtemp.ensure(tb + 16));
    CK(cub::DeviceSelect::Flagged(E->cubtemp.p, tb, it, flags, out, d_cnt, (int) n, E->stream));
    E->launches += 1;
    return 0;
}

int sharded_force_step(Engine *E, const double *d_pos, const float *d_mass, const double *d_old3, int64_t n_own,
                       const b200_gravshort_params *par, double *d_gpm, double *d_acc, double *d_pot, b200_sharded_info *info)
{
    Sharded *S = E->sh;
    if(!S || !S->configured) return failmsg(E, "b200_sharded_force_step: call b200_sharded_init first");
    if(!par || n_own < 0 || (n_own > 0 && (!d_pos || !d_mass))) return failmsg(E, "b200_sharded_force_step: bad arguments");
    if(n_own >= (1ll << 31) - 64) return failmsg(E, "b200_sharded_force_step: too many particles on one rank");
    SlabPM *P = E->slab;
    cudaStream_t ms = E->stream;
    const size_t m = (size_t) (n_own > 0 ? n_own : 1);
    const unsigned gb = (unsigned) ((m + 255) / 256);

    // ---- 1. ghost import (the cell layer next to the domain, from both neighbours) ----
    CK(cudaEventRecord(S->ev[SH_EV_GHOST0], ms));
    CK(S->flagL.ensure(m)); CK(S->flagR.ensure(m)); CK(S->idxL.ensure(m)); CK(S->idxR.ensure(m));
    CK(cudaMemsetAsync(S->cnt.p, 0, 8 * sizeof(long long), ms));
    long long h_cnt[4] = {0, 0, 0, 0};
    int h_err[2] = {0, 0};
    if(n_own > 0) {
        k_sh_ghost_flags<<<gb, 256, 0, ms>>>(d_pos, n_own, S->shift, S->domainfac, 21 - S->d, S->lo, S->hi, S->world, S->flagL.p, S->flagR.p, S->err.p);
        CKL(E);
        if(S->world > 1) {
            if(int rc = sh_select(E, S->flagL.p, n_own, S->idxL.p, S->cnt.p + 0)) return rc;
            if(int rc = sh_select(E, S->flagR.p, n_own, S->idxR.p, S->cnt.p + 1)) return rc;
        }
    }
    if(S->world > 1) {       // the counts travel from device memory: one host synchronisation for all four
        NcclApi *NC = nccl_api();
        NCK(NC->GroupStart());
        NCK(NC->Send(S->cnt.p + 0, 1, ncclInt64, S->left, S->comm_tree, ms));
        NCK(NC->Send(S->cnt.p + 1, 1, ncclInt64, S->right, S->comm_tree, ms));
        NCK(NC->Recv(S->cnt.p + 2, 1, ncclInt64, S->right, S->comm_tree, ms));
        NCK(NC->Recv(S->cnt.p + 3, 1, ncclInt64, S->left, S->comm_tree, ms));
        NCK(NC->GroupEnd());
    }
    CK(cudaMemcpyAsync(h_cnt, S->cnt.p, 4 * sizeof(long long), cudaMemcpyDeviceToHost, ms));
    CK(cudaMemcpyAsync(h_err, S->err.p, sizeof(int), cudaMemcpyDeviceToHost, ms));
    CK(cudaStreamSynchronize(ms));
    if(h_err[0]) {
        cudaMemsetAsync(S->err.p, 0, sizeof(int), ms);
        return failmsg(E, "b200_sharded_force_step: " + std::to_string(h_err[0]) + " particles lie outside this rank's x-layers (they must be handed to their owner first)");
    }
    const int64_t nL = h_cnt[0], nR = h_cnt[1], nFR = h_cnt[2], nFL = h_cnt[3];
    const int64_t n_tot = n_own + nFL + nFR;
    CK(S->sendL.ensure(4 * (size_t) (nL + 1))); CK(S->sendR.ensure(4 * (size_t) (nR + 1)));
    CK(S->recvL.ensure(4 * (size_t) (nFL + 1))); CK(S->recvR.ensure(4 * (size_t) (nFR + 1)));
    if(nL > 0) { k_sh_pack<<<(unsigned) ((nL + 255) / 256), 256, 0, ms>>>(d_pos, d_mass, S->idxL.p, S->cnt.p + 0, (double4 *) S->sendL.p); CKL(E); }
    if(nR > 0) { k_sh_pack<<<(unsigned) ((nR + 255) / 256), 256, 0, ms>>>(d_pos, d_mass, S->idxR.p, S->cnt.p + 1, (double4 *) S->sendR.p); CKL(E); }
    if(int rc = sh_neighbour_exchange(E, S, S->comm_tree, ms, S->sendL.p, 4 * (size_t) nL, S->sendR.p, 4 * (size_t) nR,
                                      S->recvR.p, 4 * (size_t) nFR, S->recvL.p, 4 * (size_t) nFL)) return rc;
    // own + ghosts as the engine's particle set: [own | from left | from right]
    const size_t mt = (size_t) (n_tot > 0 ? n_tot : 1);
    CK(E->pos.ensure(3 * mt)); CK(E->mass.ensure(mt)); CK(E->type.ensure(mt)); CK(E->flags.ensure(mt)); CK(E->oldacc.ensure(mt));
    CK(S->iota.ensure(m));
    E->n = n_tot; E->tree_valid = false; E->potential_valid = false; E->have_last_tree = E->have_last_pm = false;
    S->n_own = n_own; S->n_tot = n_tot;
    if(n_own > 0) {
        CK(cudaMemcpyAsync(E->pos.p, d_pos, 3 * (size_t) n_own * sizeof(double), cudaMemcpyDeviceToDevice, ms));
        CK(cudaMemcpyAsync(E->mass.p, d_mass, (size_t) n_own * sizeof(float), cudaMemcpyDeviceToDevice, ms));
    }
    if(nFL > 0) { k_sh_unpack<<<(unsigned) ((nFL + 255) / 256), 256, 0, ms>>>((const double4 *) S->recvL.p, nFL, n_own, E->pos.p, E->mass.p); CKL(E); }
    if(nFR > 0) { k_sh_unpack<<<(unsigned) ((nFR + 255) / 256), 256, 0, ms>>>((const double4 *) S->recvR.p, nFR, n_own + nFL, E->pos.p, E->mass.p); CKL(E); }
    if(n_tot > 0) { k_sh_state<<<(unsigned) ((n_tot + 255) / 256), 256, 0, ms>>>(n_tot, n_own, d_old3, E->type.p, E->flags.p, E->oldacc.p, S->iota.p); CKL(E); }
    CK(cudaEventRecord(S->ev[SH_EV_GHOST1], ms));

    // ---- PM long-range step on its own stream ----
    CK(cudaEventRecord(S->ev_fork, ms));         // the chain depends on the assembled particle set only
    auto pm_chain = [&]() -> int {
    CK(cudaStreamWaitEvent(S->pm_stream, S->ev_fork, 0));
    {
        cudaStream_t ps = S->pm_stream;
        E->stream = ps;
        int rc = pmslab_set_stream(E, ps);
        const size_t N = P->N, NN = N * N, h = P->halo, nx = P->nx;
        const size_t chunk = (size_t) P->ny * P->Nz, cplx_doubles = 2 * nx * N * (size_t) P->Nz;
        double *real = P->real.p;
        auto mark = [&](int id) { if(!rc && cudaEventRecord(S->ev[id], ps) != cudaSuccess) rc = failmsg(E, "cudaEventRecord failed"); };
        mark(SH_EV_PM0);
        if(!rc) rc = pmslab_deposit(E, n_own, false);
        mark(SH_EV_DEP);
        // density halo planes -> added into the neighbours' edge planes (petapm.c:787-790)
        if(!rc) rc = sh_neighbour_exchange(E, S, S->comm_pm, ps, real, h * NN, real + (h + nx) * NN, h * NN, S->haloA.p, h * NN, S->haloB.p, h * NN);
        if(!rc) {
            k_sh_add<<<148 * 4, 256, 0, ps>>>(real + nx * NN, S->haloA.p, h * NN);        // from my right neighbour: its left halo = my last owned planes
            k_sh_add<<<148 * 4, 256, 0, ps>>>(real + h * NN, S->haloB.p, h * NN);         // from my left neighbour: its right halo = my first owned planes
            E->launches += 2;
        }
        mark(SH_EV_HADD);
        if(!rc) rc = pmslab_fft2d(E, 0);
        mark(SH_EV_F2D);
        // [ix][dest][jy][kz] -> [dest][ix][jy][kz], all-to-all, and the receive buffer is the x-major y-slab
        if(!rc) { k_sh_swap_blocks<<<148 * 8, 256, 0, ps>>>((const double2 *) P->cplx.p, (double2 *) S->xbuf.p, (int) nx, S->world, chunk); E->launches++; }
        mark(SH_EV_PACK);
        if(!rc) rc = sh_all_to_all(E, S, ps, S->xbuf.p, P->cplxT.p, cplx_doubles / S->world);
        mark(SH_EV_A2A1);
        if(!rc) rc = pmslab_fft1d(E, 0);
        if(!rc) rc = pmslab_transfer(E);
        if(!rc) rc = pmslab_fft1d(E, 1);
        mark(SH_EV_F1D);
        if(!rc) rc = sh_all_to_all(E, S, ps, P->cplxT.p, S->xbuf.p, cplx_doubles / S->world);
        mark(SH_EV_A2A2);
        // [src][ix][jy][kz] -> [ix][src][jy][kz]
        if(!rc) { k_sh_swap_blocks<<<148 * 8, 256, 0, ps>>>((const double2 *) S->xbuf.p, (double2 *) P->cplx.p, S->world, (int) nx, chunk); E->launches++; }
        mark(SH_EV_UNPACK);
        if(!rc) rc = pmslab_fft2d(E, 1);
        mark(SH_EV_I2D);
        // potential halo planes <- the neighbours' edge planes (petapm.c:848-885), received in place
        if(!rc) rc = sh_neighbour_exchange(E, S, S->comm_pm, ps, real + h * NN, h * NN, real + nx * NN, h * NN, real + (h + nx) * NN, h * NN, real, h * NN);
        mark(SH_EV_HFILL);
        if(!rc) rc = pmslab_readout(E, n_own, d_gpm, nullptr, false);
        mark(SH_EV_READ);
        if(!rc && cudaEventRecord(S->ev_join, ps) != cudaSuccess) rc = failmsg(E, "cudaEventRecord failed");
        E->stream = ms;
        pmslab_set_stream(E, ms);
        if(rc) { cudaStreamSynchronize(ps); return rc; }
    }
    return 0;
    };

    // The walk kernel is sensitive to memory latency: with the PM chain beside it the two simply add up (measured).  The
    // pair kernel prefetches its operands, so the chain is issued when the walk has finished and runs beside the pair
    // sums (grav_short_tree returns to the host once the walk is complete and the pair kernel is launched).
    // B200_SHARDED_PM_FIRST=1 issues it before the tree build instead.
    static const bool pm_first = getenv("B200_SHARDED_PM_FIRST") != nullptr && atoi(getenv("B200_SHARDED_PM_FIRST")) != 0;
    if(pm_first) if(int rc = pm_chain()) return rc;
    // ---- 2. tree build over own + ghosts, top-cell moments summed over ranks, walk of the own particles ----
    if(int rc = tree_build(E, S->box, 63, nullptr, 0, S->d, nullptr)) { cudaStreamSynchronize(S->pm_stream); return rc; }
    CK(cudaEventRecord(S->ev[SH_EV_TOP0], ms));
    if(S->world > 1) {
        size_t ncell3 = 1; for(int l = 0; l < S->d; l++) ncell3 *= 8;
        if(int rc = tree_top_get(E, S->d, S->top.p)) return rc;
        k_sh_zero_foreign<<<(unsigned) ((ncell3 + 255) / 256), 256, 0, ms>>>((double4 *) S->top.p, (int) ncell3, S->d, S->lo, S->hi); CKL(E);
        NcclApi *NC = nccl_api();
        NCK(NC->AllReduce(S->top.p, S->top.p, 4 * ncell3, ncclDouble, ncclSum, S->comm_tree, ms));
        if(int rc = tree_top_set(E, S->d, S->top.p)) return rc;
    }
    CK(cudaEventRecord(S->ev[SH_EV_TOP1], ms));
    if(n_own > 0)
        if(int rc = grav_short_tree(E, par, S->iota.p, n_own, d_acc, d_pot, nullptr)) { cudaStreamSynchronize(S->pm_stream); return rc; }
    if(!pm_first) if(int rc = pm_chain()) return rc;
    CK(cudaStreamWaitEvent(ms, S->ev_join, 0));
    CK(cudaMemcpyAsync(h_err, P->err.p, sizeof(int), cudaMemcpyDeviceToHost, ms));
    CK(cudaStreamSynchronize(ms));
    if(h_err[0]) {
        cudaMemsetAsync(P->err.p, 0, sizeof(int), ms);
        return failmsg(E, "b200_sharded_force_step: particles outside this rank's mesh slab + halo planes");
    }
    collect_timings(E);
    if(info) {
        memset(info, 0, sizeof(*info));
        info->n_own = n_own; info->n_from_left = nFL; info->n_from_right = nFR; info->n_to_left = nL; info->n_to_right = nR;
        auto el = [&](int a, int b) { float t = 0; cudaEventElapsedTime(&t, S->ev[a], S->ev[b]); return (double) t; };
        info->ms_ghost = el(SH_EV_GHOST0, SH_EV_GHOST1);
        info->ms_pm_total = el(SH_EV_PM0, SH_EV_READ);
        info->ms_pm_deposit = el(SH_EV_PM0, SH_EV_DEP); info->ms_pm_halo_add = el(SH_EV_DEP, SH_EV_HADD);
        info->ms_pm_fft2d = el(SH_EV_HADD, SH_EV_F2D); info->ms_pm_pack = el(SH_EV_F2D, SH_EV_PACK);
        info->ms_pm_a2a_forward = el(SH_EV_PACK, SH_EV_A2A1); info->ms_pm_fft1d_transfer = el(SH_EV_A2A1, SH_EV_F1D);
        info->ms_pm_a2a_backward = el(SH_EV_F1D, SH_EV_A2A2); info->ms_pm_unpack = el(SH_EV_A2A2, SH_EV_UNPACK);
        info->ms_pm_ifft2d = el(SH_EV_UNPACK, SH_EV_I2D); info->ms_pm_halo_fill = el(SH_EV_I2D, SH_EV_HFILL);
        info->ms_pm_readout = el(SH_EV_HFILL, SH_EV_READ);
        info->ms_top_allreduce = el(SH_EV_TOP0, SH_EV_TOP1);
    }
    return 0;
}

} // namespace b200

using namespace b200;
#define ENTER(ctx) if(!(ctx)) return 1; Engine *E = &(ctx)->e; if(cudaSetDevice(E->device) != cudaSuccess) return failmsg(E, "cudaSetDevice failed");

extern "C" {

int b200_comm_unique_id(void *id_out)
{
    NcclApi *NC = nccl_api();
    if(!NC || !id_out) return 1;
    ncclUniqueId id;
    if(NC->GetUniqueId(&id) != ncclSuccess) return 2;
    memcpy(id_out, &id, sizeof(id));
    return 0;
}

int b200_comm_init(b200_ctx *ctx, int rank, int world, const void *id_tree, const void *id_pm)
{
    ENTER(ctx);
    return sharded_comm_init(E, rank, world, id_tree, id_pm);
}

int b200_sharded_init(b200_ctx *ctx, double BoxSize, double Asmth, int Nmesh, double G, int topdepth, int halo, double rcut_cells)
{
    ENTER(ctx);
    return sharded_init(E, BoxSize, Asmth, Nmesh, G, topdepth, halo, rcut_cells);
}

int b200_sharded_force_step(b200_ctx *ctx, const double *pos_own, const float *mass_own, const double *oldacc3_own, int64_t n_own,
                            const b200_gravshort_params *par, double *gravpm_out, double *accel_out, double *potential_out,
                            b200_sharded_info *info)
{
    ENTER(ctx);
    return sharded_force_step(E, pos_own, mass_own, oldacc3_own, n_own, par, gravpm_out, accel_out, potential_out, info);
}

} // extern "C"
