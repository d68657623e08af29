// fof.cu -- primary friends-of-friends linking on the device (another consumer of the neighbour search).
//
// Replaces fof_label_primary + fof_primary_ngbiter (libgadget/fof.c:366-470,540-579): particles of the primary link
// types closer than the linking length (periodic, NEAREST) belong to one group and every member carries the smallest
// particle ID of its group (HaloLabel[].MinID); particles of other types keep their own ID.  The reference reaches that
// fixed point by repeated tree walks that merge heads and propagate MinID (one treewalk_run per iteration until no
// link changes); the result is the set of connected components of the distance graph, which is computed here in one
// pass: particles sorted by the cell of a grid of spacing >= the linking length, every particle examines the 27 cells
// round it (9 contiguous runs of the sorted cell keys found by binary search, split where the run wraps), and a link
// hooks the larger of the two roots to the smaller with atomicCAS (lock-free union-find, parents only ever decrease).
// Distances are the CPU's statement by statement (the file is compiled with -fmad=false), so membership -- hence every
// label -- equals the reference's bit for bit.
#include "engine.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <cub/cub.cuh>

namespace b200 {

typedef unsigned long long fofkey;          // grid cell (cx nc + cy) nc + cz, up to 2^20 cells a side

// Root of i with intermediate pointer jumping: every parent visited is re-pointed at its grandparent.  parent[x] <= x
// always, racing writers only ever store a smaller ancestor, so a reader can only be sent further up the same tree.
__device__ __forceinline__ int fof_find(int *parent, int i)
{
    int cur = __ldcg(parent + i);          // L2 reads: another SM's hook or jump is seen as soon as it lands
    if(cur != i) {
        int prev = i, next;
        while(cur > (next = __ldcg(parent + cur))) {
            parent[prev] = next;
            prev = cur; cur = next;
        }
    }
    return cur;
}

__device__ __forceinline__ void fof_union(int *parent, int a, int b)
{
    a = fof_find(parent, a); b = fof_find(parent, b);
    while(a != b) {
        if(a < b) { const int t = a; a = b; b = t; }          // a is the larger root
        const int old = atomicCAS(&parent[a], a, b);
        if(old == a) break;                                    // hooked
        a = fof_find(parent, old);                             // somebody hooked a first: start again from its new root
        b = fof_find(parent, b);
    }
}

// cell of every primary particle (key) and its index (val); others get the key past the grid so that they sort to the end
__global__ void __launch_bounds__(256)
k_fof_cells(const double *__restrict__ pos, const uint8_t *__restrict__ type, const uint8_t *__restrict__ flags, int64_t n, int mask,
            double cs, int nc, fofkey past, fofkey *__restrict__ key, int *__restrict__ val, int *__restrict__ parent)
{
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    parent[i] = (int) i;
    val[i] = (int) i;
    fofkey k = past;
    if(!(flags[i] & 3) && ((mask >> type[i]) & 1)) {
        int c[3];
#pragma unroll
        for(int d = 0; d < 3; d++) {
            int q = (int) floor(pos[3 * i + d] / cs);
            q %= nc; if(q < 0) q += nc;
            c[d] = q;
        }
        k = ((fofkey) c[0] * nc + c[1]) * nc + c[2];
    }
    key[i] = k;
}

__global__ void __launch_bounds__(256)
k_fof_gather(const double *__restrict__ pos, const int *__restrict__ val, int np, double *__restrict__ spos)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if(k >= np) return;
    const int i = val[k];
    spos[3 * k] = pos[3 * (size_t) i]; spos[3 * k + 1] = pos[3 * (size_t) i + 1]; spos[3 * k + 2] = pos[3 * (size_t) i + 2];
}

// first position in [lo, hi) whose key is >= want
__device__ __forceinline__ int fof_lower_bound(const fofkey *__restrict__ key, int lo, int hi, fofkey want)
{
    while(lo < hi) {
        const int mid = (lo + hi) >> 1;
        if(key[mid] < want) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__global__ void k_fof_count_primary(const fofkey *__restrict__ key, int n, fofkey past, int *__restrict__ out)
{
    *out = fof_lower_bound(key, 0, n, past);
}

// One thread per primary particle k (sorted order): candidates c > k of the 27 cells round it; r2 <= ll^2 (an asymmetric
// search, treewalk.c:989-993) links the two.  nc < 3: the grid has one cell and every pair is examined.
__global__ void __launch_bounds__(128)
k_fof_link(const double *__restrict__ spos, const fofkey *__restrict__ key, const int *__restrict__ val, int np, int nc, double Box,
           double ll2, int *__restrict__ parent)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if(k >= np) return;
    const double px = spos[3 * k], py = spos[3 * k + 1], pz = spos[3 * k + 2];
    const int me = val[k];
    const fofkey kk = key[k];
    const int cz = (int) (kk % nc), cy = (int) ((kk / nc) % nc), cx = (int) (kk / ((fofkey) nc * nc));
    const double half = 0.5 * Box;
    const int reach = nc >= 3 ? 1 : 0;
    for(int dx = -reach; dx <= reach; dx++) for(int dy = -reach; dy <= reach; dy++) {
        const int x = (cx + dx + nc) % nc, y = (cy + dy + nc) % nc;
        const fofkey row = ((fofkey) x * nc + y) * nc;
        // the z cells cz-1 .. cz+1 as runs of consecutive keys: one run, or two where the range wraps
        int z0[2], z1[2], nrun = 1;
        if(!reach) { z0[0] = 0; z1[0] = nc - 1; }
        else if(cz == 0) { z0[0] = 0; z1[0] = 1; z0[1] = nc - 1; z1[1] = nc - 1; nrun = 2; }
        else if(cz == nc - 1) { z0[0] = nc - 2; z1[0] = nc - 1; z0[1] = 0; z1[1] = 0; nrun = 2; }
        else { z0[0] = cz - 1; z1[0] = cz + 1; }
        for(int r = 0; r < nrun; r++) {
            const fofkey last = row + (fofkey) z1[r];
            for(int c = fof_lower_bound(key, 0, np, row + (fofkey) z0[r]); c < np && key[c] <= last; c++) {
                if(c <= k) continue;
                double ddx = px - spos[3 * c], ddy = py - spos[3 * c + 1], ddz = pz - spos[3 * c + 2];
                if(ddx > half) ddx -= Box;             // NEAREST
                if(ddx < -half) ddx += Box;
                if(ddy > half) ddy -= Box;
                if(ddy < -half) ddy += Box;
                if(ddz > half) ddz -= Box;
                if(ddz < -half) ddz += Box;
                double r2 = 0;
                r2 += ddx * ddx; r2 += ddy * ddy; r2 += ddz * ddz;
                if(r2 <= ll2) fof_union(parent, me, val[c]);
            }
        }
    }
}

__device__ __forceinline__ double fof_r2(const double *__restrict__ spos, double px, double py, double pz, int c, double Box, double half)
{
    double ddx = px - spos[3 * c], ddy = py - spos[3 * c + 1], ddz = pz - spos[3 * c + 2];
    if(ddx > half) ddx -= Box;             // NEAREST
    if(ddx < -half) ddx += Box;
    if(ddy > half) ddy -= Box;
    if(ddy < -half) ddy += Box;
    if(ddz > half) ddz -= Box;
    if(ddz < -half) ddz += Box;
    double r2 = 0;
    r2 += ddx * ddx; r2 += ddy * ddy; r2 += ddz * ddz;
    return r2;
}

// The same search on a grid of spacing just below ll / sqrt(3): the particles of one cell are within the linking length
// of each other whatever their positions, so a cell is a clique -- every member hooks to the cell's first particle
// without a distance test, and two cells need ONE pair inside the linking length to be joined for good.  A particle scans
// a neighbouring cell (up to `reach` = 2 cells away, only cells with a larger key: the pair of cells is examined from one
// side) until its first hit, and not at all when the two cells already share a root.  Dense clumps, where the plain cell
// list counts pairs (k^2 per cell), cost k per cell; the set of links found differs, the components do not.
__global__ void __launch_bounds__(128)
k_fof_link_clique(const double *__restrict__ spos, const fofkey *__restrict__ key, const int *__restrict__ val, int np, int nc, int reach,
                  double Box, double ll2, int *__restrict__ parent)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if(k >= np) return;
    const double px = spos[3 * k], py = spos[3 * k + 1], pz = spos[3 * k + 2];
    const int me = val[k];
    const fofkey kk = key[k];
    const int cz = (int) (kk % nc), cy = (int) ((kk / nc) % nc), cx = (int) (kk / ((fofkey) nc * nc));
    const double half = 0.5 * Box;
    const int f = fof_lower_bound(key, 0, k + 1, kk);          // first particle of my cell
    if(f != k) fof_union(parent, me, val[f]);
    for(int dx = -reach; dx <= reach; dx++) for(int dy = -reach; dy <= reach; dy++) {
        const int x = (cx + dx + nc) % nc, y = (cy + dy + nc) % nc;
        const fofkey row = ((fofkey) x * nc + y) * nc;
        int z0[2], z1[2], nrun = 1;
        const int zlo = cz - reach, zhi = cz + reach;           // nc >= 2 reach + 1: the two runs of a wrapped range do not overlap
        if(zlo < 0) { z0[0] = 0; z1[0] = zhi; z0[1] = zlo + nc; z1[1] = nc - 1; nrun = 2; }
        else if(zhi >= nc) { z0[0] = zlo; z1[0] = nc - 1; z0[1] = 0; z1[1] = zhi - nc; nrun = 2; }
        else { z0[0] = zlo; z1[0] = zhi; }
        for(int r = 0; r < nrun; r++) {
            const fofkey last = row + (fofkey) z1[r];
            int c = fof_lower_bound(key, 0, np, row + (fofkey) z0[r]);
            while(c < np && key[c] <= last) {
                const fofkey cell = key[c];
                const int cend = fof_lower_bound(key, c, np, cell + 1);
                if(cell > kk && fof_find(parent, me) != fof_find(parent, val[c])) {
                    for(int c2 = c; c2 < cend; c2++) {
                        if(((c2 - c) & 31) == 31 && fof_find(parent, me) == fof_find(parent, val[c])) break;      // joined meanwhile
                        if(fof_r2(spos, px, py, pz, c2, Box, half) <= ll2) { fof_union(parent, me, val[c2]); break; }
                    }
                }
                c = cend;
            }
        }
    }
}

// root of every particle; the smallest ID of the group collects at the root
__global__ void __launch_bounds__(256)
k_fof_min(int *__restrict__ parent, const long long *__restrict__ ids, int64_t n, int *__restrict__ root, unsigned long long *__restrict__ rootmin)
{
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    const int r = fof_find(parent, (int) i);
    root[i] = r;
    atomicMin(&rootmin[r], (unsigned long long) ids[i]);
}
__global__ void __launch_bounds__(256)
k_fof_label(const int *__restrict__ root, const unsigned long long *__restrict__ rootmin, int64_t n, long long *__restrict__ minid)
{
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    minid[i] = (long long) rootmin[root[i]];
}
__global__ void __launch_bounds__(256)
k_fof_init_min(const long long *__restrict__ ids, int64_t n, unsigned long long *__restrict__ rootmin)
{
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if(i < n) rootmin[i] = (unsigned long long) ids[i];
}

void fof_release(Engine *E)
{
    E->fof_key.release(); E->fof_key_alt.release(); E->fof_val.release(); E->fof_val_alt.release(); E->fof_parent.release();
    E->fof_root.release(); E->fof_spos.release(); E->fof_ids.release(); E->fof_min.release(); E->fof_out.release();
}

// ids, minid_out: host arrays of E->n entries.  *ngroups_out (optional): groups that hold a primary particle.
int fof_primary(Engine *E, const int64_t *ids, int mask, double Box, double ll, int64_t *minid_out, int64_t *ngroups_out)
{
    const int64_t n = E->n;
    if(!(Box > 0) || !(ll > 0)) return failmsg(E, "b200_fof_primary: BoxSize and the linking length must be positive");
    if(n > 0 && (!ids || !minid_out)) return failmsg(E, "b200_fof_primary: null ids / output");
    if(n >= ((int64_t) 1 << 31)) return failmsg(E, "b200_fof_primary: more than 2^31 particles");
    if(ngroups_out) *ngroups_out = 0;
    if(n == 0) return 0;
    // Grid.  Clique cells (spacing just below ll / sqrt(3), neighbours up to 2 cells away) when the box holds at least 5
    // and at most 2^20 of them a side; else the plain cell list: spacing >= ll, at most 1024 cells a side, neighbours 1
    // cell away -- or, with fewer than 3 such cells, one cell and every pair examined.  B200_FOF=cells forces the latter.
    const char *sel = getenv("B200_FOF");
    const double want = Box * sqrt(3.0) * (1.0 + 1e-6) / ll;
    bool clique = !(sel && !strcmp(sel, "cells")) && want >= 5.0 && want < 1048575.0;
    int nc, reach = 1;
    if(clique) {
        nc = (int) ceil(want);
        reach = (int) ceil(ll / (Box / nc));
        if(nc < 2 * reach + 1) clique = false;
    }
    if(!clique) {
        nc = (int) floor(Box / ll);
        if(nc > 1024) nc = 1024;
        if(nc < 3) nc = 1;
    }
    const double cs = Box / nc;
    int bits = 1;                               // key bits + 1: the key past the grid, taken by everything that is not a primary particle
    while(((fofkey) 1 << bits) < (fofkey) nc * nc * nc) bits++;
    const fofkey past = (fofkey) 1 << bits;
    bits++;
    const size_t m = (size_t) n;
    CK(E->fof_key.ensure(m)); CK(E->fof_key_alt.ensure(m)); CK(E->fof_val.ensure(m)); CK(E->fof_val_alt.ensure(m));
    CK(E->fof_parent.ensure(m)); CK(E->fof_root.ensure(m)); CK(E->fof_ids.ensure(m)); CK(E->fof_min.ensure(m)); CK(E->fof_out.ensure(m));
    CK(E->scratch_i.ensure(256));
    CK(cudaMemcpyAsync(E->fof_ids.p, ids, m * sizeof(long long), cudaMemcpyHostToDevice, E->stream));
    const unsigned nb = (unsigned) ((n + 255) / 256);
    k_fof_cells<<<nb, 256, 0, E->stream>>>(E->pos.p, E->type.p, E->flags.p, n, mask, cs, nc, past, E->fof_key_alt.p, E->fof_val_alt.p, E->fof_parent.p);
    CKL(E);
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, E->fof_key_alt.p, E->fof_key.p, E->fof_val_alt.p, E->fof_val.p, (int) n, 0, bits, E->stream);
    CK(E->cubtemp.ensure(tb + 16));
    CK(cub::DeviceRadixSort::SortPairs(E->cubtemp.p, tb, E->fof_key_alt.p, E->fof_key.p, E->fof_val_alt.p, E->fof_val.p, (int) n, 0, bits, E->stream));
    E->launches += 1;
    // primaries sort to the front: their number = first position of the past-the-grid key
    int *d_np = E->scratch_i.p + 28;
    k_fof_count_primary<<<1, 1, 0, E->stream>>>(E->fof_key.p, (int) n, past, d_np);
    CKL(E);
    int np = 0;
    CK(cudaMemcpyAsync(&np, d_np, sizeof(int), cudaMemcpyDeviceToHost, E->stream));
    CK(cudaStreamSynchronize(E->stream));
    if(np > 0) {
        if(nc == 1 && np > 65536) return failmsg(E, "b200_fof_primary: linking length above a third of the box with more than 65536 primary particles");
        CK(E->fof_spos.ensure(3 * (size_t) np));
        k_fof_gather<<<(unsigned) ((np + 255) / 256), 256, 0, E->stream>>>(E->pos.p, E->fof_val.p, np, E->fof_spos.p);
        CKL(E);
        if(clique) k_fof_link_clique<<<(unsigned) ((np + 127) / 128), 128, 0, E->stream>>>(E->fof_spos.p, E->fof_key.p, E->fof_val.p, np, nc, reach, Box, ll * ll, E->fof_parent.p);
        else k_fof_link<<<(unsigned) ((np + 127) / 128), 128, 0, E->stream>>>(E->fof_spos.p, E->fof_key.p, E->fof_val.p, np, nc, Box, ll * ll, E->fof_parent.p);
        CKL(E);
    }
    k_fof_init_min<<<nb, 256, 0, E->stream>>>(E->fof_ids.p, n, E->fof_min.p);
    CKL(E);
    k_fof_min<<<nb, 256, 0, E->stream>>>(E->fof_parent.p, E->fof_ids.p, n, E->fof_root.p, E->fof_min.p);
    CKL(E);
    k_fof_label<<<nb, 256, 0, E->stream>>>(E->fof_root.p, E->fof_min.p, n, E->fof_out.p);
    CKL(E);
    CK(cudaMemcpyAsync(minid_out, E->fof_out.p, m * sizeof(long long), cudaMemcpyDeviceToHost, E->stream));
    std::vector<int> hroot;
    if(ngroups_out) { hroot.resize(m); CK(cudaMemcpyAsync(hroot.data(), E->fof_root.p, m * sizeof(int), cudaMemcpyDeviceToHost, E->stream)); }
    std::vector<int> hval;
    if(ngroups_out && np > 0) { hval.resize(np); CK(cudaMemcpyAsync(hval.data(), E->fof_val.p, (size_t) np * sizeof(int), cudaMemcpyDeviceToHost, E->stream)); }
    CK(cudaStreamSynchronize(E->stream));
    if(ngroups_out) {
        int64_t g = 0;
        for(int k = 0; k < np; k++) if(hroot[hval[k]] == hval[k]) g++;
        *ngroups_out = g;
    }
    return 0;
}

} // namespace b200
