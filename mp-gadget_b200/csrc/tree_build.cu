// tree_build.cu -- on-device octree build (replaces force_tree_build,
// libgadget/forcetree.c:196-270,727-860 and the moment pass :1017-1143).
//
// The reference inserts particles one at a time and splits a leaf when a 9th
// particle arrives (forcetree.c:370-520).  The resulting tree is the unique
// octree in which a cell is internal iff it holds more than NMAXCHILD = 8
// particles and a child exists iff it is non-empty (forcetree.c:1028-1049) or
// belongs to the forced top tree (forcetree.c:869-934).  We build exactly that
// tree without insertion:
//   1. every particle descends the reference's own chain of cell centres
//      (root centre Box/2, side 1.001*Box, forcetree.c:662-664; child centre =
//      parent +- side/4, forcetree.c:302-320; octant by strict `>`,
//      forcetree.c:278-284) in fp64 and records 21 octant digits -> 63-bit key,
//      so cell membership is decided by the same comparisons as the reference;
//   2. CUB radix sort of (key, index);
//   3. level-by-level split of key ranges (one thread per cell, binary search
//      for the 8 digit boundaries);
//   4. subtree sizes bottom-up, depth-first positions top-down, scatter into
//      depth-first order (first child = self + 1, sibling = self + subtree size):
//      the order of the reference's sibling/suns[0] walk;
//   5. inside each leaf, particles are put in ascending original index =
//      the reference's insertion order, then moments are summed bottom-up in
//      that order (forcetree.c:947-1004,1081-1101) with un-fused fp64 mul/add.
#include "engine.h"
#include <cub/device/device_radix_sort.cuh>
#include <stdio.h>

namespace b200 {

#define KEY_LEVELS 21
#define LEAFCAP 8

__global__ void __launch_bounds__(256)
k_tree_keys(const double *__restrict__ pos, const uint8_t *__restrict__ type,
            const uint8_t *__restrict__ flags, const int *__restrict__ active, int64_t nin,
            double c0, double len0, double box, int topdepth, int mask, unsigned long long *__restrict__ keys,
            int *__restrict__ idx, int *__restrict__ nvalid,
            // an arbitrary domain top tree instead of a uniform depth (NULL = uniform): TopNodes[].Daughter / StartKey / Shift
            // and the state machine of the Peano-Hilbert curve (domain_keys.cu)
            const int *__restrict__ top_daughter, const unsigned long long *__restrict__ top_startkey,
            const int *__restrict__ top_shift, const uint8_t *__restrict__ ph_tab)
{
    __shared__ uint8_t s_tab[768];
    if(top_daughter) {
        for(int k = threadIdx.x; k < 768; k += blockDim.x) s_tab[k] = ph_tab[k];
        __syncthreads();
    }
    const int64_t j = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    bool ok = false;
    if(j < nin) {
        const int i = active ? active[j] : (int) j;
        const int ty = type[i];
        const int fl = flags[i];
        // forcetree.c:801-807: type mask, garbage, swallowed black holes
        ok = ((1 << ty) & mask) && !(fl & 1) && !((fl & 2) && ty == 5);
        unsigned long long key = ~0ull;
        if(ok) {
            const double x = pos[3 * (int64_t) i], y = pos[3 * (int64_t) i + 1], z = pos[3 * (int64_t) i + 2];
            double cx = c0, cy = c0, cz = c0, len = len0;
            key = 0;
            // Inside the forced top tree the reference places a particle below
            // P[i].TopLeaf, which comes from the integer Peano-Hilbert lattice
            // (PEANO, utils/peano.h:15-21; forcetree.c:819-823), not from position
            // compares; the lattice is the tree's own, so this only matters within
            // rounding of a cell boundary.
            const double DomainFac = 1.0 / (box * 1.001) * (double) (1ull << 21);
            const int ix = (int) ((x + box / 2000) * DomainFac);
            const int iy = (int) ((y + box / 2000) * DomainFac);
            const int iz = (int) ((z + box / 2000) * DomainFac);
            if(top_daughter) {
                // P[i].TopLeaf = domain_get_topleaf(PEANO(Pos)) (domain.h:71-78, peano.h:15-21): the particle follows the
                // lattice down to the level of its top leaf (forcetree.c:819-823)
                unsigned long long pk = 0;
                int st = 0;
#pragma unroll 1
                for(int bit = 20; bit >= 0; bit--) {
                    const int pix = (((ix >> bit) & 1) << 2) | (((iy >> bit) & 1) << 1) | ((iz >> bit) & 1);
                    pk = (pk << 3) | s_tab[st * 8 + pix];
                    st = s_tab[384 + st * 8 + pix];
                }
                int no = 0;
                while(top_daughter[no] >= 0) no = top_daughter[no] + (int) ((pk - top_startkey[no]) >> (top_shift[no] - 3));
                topdepth = (63 - top_shift[no]) / 3;
            }
#pragma unroll 1
            for(int l = 0; l < KEY_LEVELS; l++) {
                const double lenhalf = 0.25 * len;
                int bx = x > cx, by = y > cy, bz = z > cz;
                if(l < topdepth) {
                    bx = (ix >> (20 - l)) & 1; by = (iy >> (20 - l)) & 1; bz = (iz >> (20 - l)) & 1;
                }
                key = (key << 3) | (unsigned long long) (bx | (by << 1) | (bz << 2));
                cx = bx ? cx + lenhalf : cx - lenhalf;
                cy = by ? cy + lenhalf : cy - lenhalf;
                cz = bz ? cz + lenhalf : cz - lenhalf;
                len = 0.5 * len;
            }
        }
        keys[j] = key;
        idx[j] = i;
    }
    const unsigned m = __ballot_sync(0xffffffffu, ok);
    if((threadIdx.x & 31) == 0 && m) atomicAdd(nvalid, __popc(m));
}

// One thread per cell of the current level: split into children.
// counters[0] = total nodes allocated, counters[1] = overflow flag,
// counters[2] = overfull leaves at the key-depth limit.
__global__ void __launch_bounds__(128)
k_tree_split(const unsigned long long *__restrict__ keys, int first, int last, int level,
             int topdepth, int cap, int *__restrict__ counters,
             int *__restrict__ b_start, int *__restrict__ b_count, int *__restrict__ b_father,
             int *__restrict__ b_firstchild, int *__restrict__ b_nchild, double *__restrict__ b_center,
             // arbitrary top tree (NULL = uniform depth): b_top[node] = top node | curve state << 24, or -1 below the top leaves
             int *__restrict__ b_top, const int *__restrict__ top_daughter, const uint8_t *__restrict__ ph_tab)
{
    const int node = first + blockIdx.x * blockDim.x + threadIdx.x;
    if(node >= last) return;
    const int s = b_start[node], cnt = b_count[node];
    const int mytop = top_daughter ? b_top[node] : -1;
    const int mydau = mytop >= 0 ? top_daughter[mytop & 0xffffff] : -1;
    // force_create_node_for_topnode forcetree.c:869-934: a cell whose top node has daughters gets all eight children
    const bool forced = top_daughter ? mydau >= 0 : level < topdepth;
    bool internal = forced || cnt > LEAFCAP;
    if(internal && level >= KEY_LEVELS) {
        internal = false;
        atomicAdd(&counters[2], 1);
    }
    if(!internal) { b_firstchild[node] = -1; b_nchild[node] = 0; return; }
    const int shift = 3 * (KEY_LEVELS - (level + 1));
    int bound[9];
    bound[0] = s; bound[8] = s + cnt;
    for(int d = 1; d < 8; d++) {
        int lo = bound[d - 1], hi = s + cnt;      // first index with digit >= d
        while(lo < hi) {
            const int mid = (lo + hi) >> 1;
            const int dig = (int) ((keys[mid] >> shift) & 7ull);
            if(dig < d) lo = mid + 1; else hi = mid;
        }
        bound[d] = lo;
    }
    const bool keep_empty = top_daughter ? forced : (level + 1) <= topdepth;
    int nch = 0;
    for(int d = 0; d < 8; d++) nch += (keep_empty || bound[d + 1] > bound[d]) ? 1 : 0;
    // Forced top-tree levels are complete (every cell has 8 children), so their
    // children go to fixed slots: cell (level l, Morton index m) sits at
    // (8^l - 1)/7 + m.  b200_tree_top_get/set rely on this.
    int base;
    if(forced && !top_daughter) {
        base = last + 8 * (node - first);
        atomicMax(&counters[0], last + 8 * (last - first));
    } else {
        base = atomicAdd(&counters[0], nch);
    }
    if(base + nch > cap) { counters[1] = 1; b_firstchild[node] = -1; b_nchild[node] = 0; return; }
    b_firstchild[node] = base;
    b_nchild[node] = nch;
    const double cx = b_center[4 * (size_t) node], cy = b_center[4 * (size_t) node + 1],
                 cz = b_center[4 * (size_t) node + 2], len = b_center[4 * (size_t) node + 3];
    const double lenhalf = 0.25 * len;      // init_internal_node forcetree.c:305-320
    int k = 0;
    for(int d = 0; d < 8; d++) {
        const int c = bound[d + 1] - bound[d];
        if(!(keep_empty || c > 0)) continue;
        const int ch = base + k++;
        b_start[ch] = bound[d];
        b_count[ch] = c;
        b_father[ch] = node;
        b_center[4 * (size_t) ch] = (d & 1) ? cx + lenhalf : cx - lenhalf;
        b_center[4 * (size_t) ch + 1] = (d & 2) ? cy + lenhalf : cy - lenhalf;
        b_center[4 * (size_t) ch + 2] = (d & 4) ? cz + lenhalf : cz - lenhalf;
        b_center[4 * (size_t) ch + 3] = 0.5 * len;
        if(top_daughter) {
            int ctop = -1;
            if(forced) {        // the daughter covering octant d is found through the curve (forcetree.c:886,904)
                const int st = mytop >> 24, pix = ((d & 1) << 2) | (((d >> 1) & 1) << 1) | ((d >> 2) & 1);
                ctop = (mydau + ph_tab[st * 8 + pix]) | ((int) ph_tab[384 + st * 8 + pix] << 24);
            }
            b_top[ch] = ctop;
        }
    }
}

__global__ void k_tree_root(int np, double c0, double len0, int *b_start, int *b_count, int *b_father,
                            double *b_center, int *counters, int *b_top)
{
    if(b_top) b_top[0] = 0;
    b_start[0] = 0; b_count[0] = np; b_father[0] = -1;
    b_center[0] = c0; b_center[1] = c0; b_center[2] = c0; b_center[3] = len0;
    counters[0] = 1; counters[1] = 0; counters[2] = 0;
}

// Ascending original index inside every leaf = the reference's insertion order.
__global__ void __launch_bounds__(128)
k_tree_leaf_order(int nn, const int *__restrict__ b_start, const int *__restrict__ b_count,
                  const int *__restrict__ b_nchild, int *__restrict__ sidx)
{
    const int node = blockIdx.x * blockDim.x + threadIdx.x;
    if(node >= nn || b_nchild[node] != 0) return;
    const int s = b_start[node], c = b_count[node];
    if(c < 2 || c > 4096) return;
    for(int a = 1; a < c; a++) {
        const int v = sidx[s + a];
        int b = a - 1;
        while(b >= 0 && sidx[s + b] > v) { sidx[s + b + 1] = sidx[s + b]; b--; }
        sidx[s + b + 1] = v;
    }
}

__global__ void __launch_bounds__(256)
k_tree_gather(const double *__restrict__ pos, const float *__restrict__ mass,
              const int *__restrict__ sidx, int np, double far, double4 *__restrict__ spart,
              double2 *__restrict__ spart_xy, double2 *__restrict__ spart_zm)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if(j >= np + B200_SPART_PAD) return;
    double4 q = make_double4(far, far, far, 0.0);
    if(j < np) {
        const int64_t i = sidx[j];
        q = make_double4(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], (double) mass[i]);
    }
    spart[j] = q;
    spart_xy[j] = make_double2(q.x, q.y);       // the pair kernel's two 16-byte streams (tree_walk.cu)
    spart_zm[j] = make_double2(q.z, q.w);
}

// Moments + subtree sizes for the cells of one level (children already done).
__global__ void __launch_bounds__(128)
k_tree_moments(int first, int last, const int *__restrict__ b_start, const int *__restrict__ b_count,
               const int *__restrict__ b_firstchild, const int *__restrict__ b_nchild,
               const double *__restrict__ b_center, const double4 *__restrict__ spart,
               double4 *__restrict__ b_mom, int *__restrict__ b_size)
{
    const int node = first + blockIdx.x * blockDim.x + threadIdx.x;
    if(node >= last) return;
    double m = 0, sx = 0, sy = 0, sz = 0;
    int size = 1;
    const int nch = b_nchild[node];
    if(nch == 0) {
        const int s = b_start[node], c = b_count[node];
        for(int k = 0; k < c; k++) {         // add_particle_moment_to_node forcetree.c:947-954
            const double4 p = spart[s + k];
            m = __dadd_rn(m, p.w);
            sx = __dadd_rn(sx, __dmul_rn(p.w, p.x));
            sy = __dadd_rn(sy, __dmul_rn(p.w, p.y));
            sz = __dadd_rn(sz, __dmul_rn(p.w, p.z));
        }
    } else {
        const int fc = b_firstchild[node];
        for(int k = 0; k < nch; k++) {       // forcetree.c:1081-1092
            const double4 cm = b_mom[fc + k];
            m = __dadd_rn(m, cm.w);
            sx = __dadd_rn(sx, __dmul_rn(cm.w, cm.x));
            sy = __dadd_rn(sy, __dmul_rn(cm.w, cm.y));
            sz = __dadd_rn(sz, __dmul_rn(cm.w, cm.z));
            size += b_size[fc + k];
        }
    }
    double4 out;
    if(m > 0) {                              // forcetree.c:994-1003,1095-1101
        out.x = __ddiv_rn(sx, m); out.y = __ddiv_rn(sy, m); out.z = __ddiv_rn(sz, m);
    } else {
        out.x = b_center[4 * (size_t) node]; out.y = b_center[4 * (size_t) node + 1]; out.z = b_center[4 * (size_t) node + 2];
    }
    out.w = m;
    b_mom[node] = out;
    b_size[node] = size;
}

// Depth-first positions, top-down.
__global__ void __launch_bounds__(128)
k_tree_dfs(int first, int last, const int *__restrict__ b_firstchild, const int *__restrict__ b_nchild,
           const int *__restrict__ b_size, int *__restrict__ b_dfs)
{
    const int node = first + blockIdx.x * blockDim.x + threadIdx.x;
    if(node >= last) return;
    const int nch = b_nchild[node];
    if(nch == 0) return;
    int run = b_dfs[node] + 1;
    const int fc = b_firstchild[node];
    for(int k = 0; k < nch; k++) { b_dfs[fc + k] = run; run += b_size[fc + k]; }
}

__global__ void __launch_bounds__(256)
k_tree_scatter(int nn, const int *__restrict__ b_dfs, const int *__restrict__ b_size,
               const int *__restrict__ b_start, const int *__restrict__ b_count,
               const int *__restrict__ b_nchild, const int *__restrict__ b_firstchild, const int *__restrict__ b_father,
               const double4 *__restrict__ b_center, const double4 *__restrict__ b_mom,
               double4 *__restrict__ nodeA, double4 *__restrict__ nodeB, int4 *__restrict__ nodeC,
               int *__restrict__ nodeF, double *__restrict__ nodeH, int4 *__restrict__ nodeK)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if(b >= nn) return;
    const int d = b_dfs[b];
    nodeA[d] = b_mom[b];
    nodeB[d] = b_center[b];
    const int sib = d + b_size[b];
    nodeC[d] = make_int4(sib < nn ? sib : -1, b_start[b], b_count[b], b_nchild[b] == 0 ? 1 : 0);
    const int f = b_father[b];
    nodeF[d] = f >= 0 ? b_dfs[f] : -1;
    nodeH[d] = 0.0;
    // depth-first positions of the (up to 8) children, so a batched walk can
    // push them without chasing sibling pointers
    int kid[8];
    const int nch = b_nchild[b], fc = b_firstchild[b];
#pragma unroll
    for(int k = 0; k < 8; k++) kid[k] = k < nch ? b_dfs[fc + k] : -1;
    nodeK[2 * (size_t) d] = make_int4(kid[0], kid[1], kid[2], kid[3]);
    nodeK[2 * (size_t) d + 1] = make_int4(kid[4], kid[5], kid[6], kid[7]);
}

int tree_build(Engine *E, double Box, int mask, const int32_t *d_active, int64_t nactive,
               int toplevel_depth, b200_tree_info *info)
{
    E->tree_valid = false;
    if(!(Box > 0)) return failmsg(E, "b200_tree_build: BoxSize must be positive");
    // toplevel_depth = -1: the forced top tree is the domain's (b200_domain_set_topnodes), of any shape
    const bool topn = toplevel_depth == -1;
    if(topn) {
        if(E->dk_ntop == 0) return failmsg(E, "b200_tree_build: toplevel_depth = -1 needs b200_domain_set_topnodes first");
        if(int rc = domain_need_tables(E)) return rc;
    }
    else if(toplevel_depth < 0 || toplevel_depth > 8) return failmsg(E, "b200_tree_build: toplevel_depth out of range [0,8] (or -1: the domain's top nodes)");
    const int *t_dau = topn ? E->dk_daughter.p : nullptr;
    const unsigned long long *t_key = topn ? E->dk_startkey.p : nullptr;
    const int *t_shift = topn ? E->dk_shift.p : nullptr;
    const uint8_t *t_tab = topn ? E->dk_tab.p : nullptr;
    const int64_t nin = d_active ? nactive : E->n;
    if(nin >= (1ll << 30)) return failmsg(E, "b200_tree_build: too many particles for 32-bit node indices");
    const double c0 = Box / 2., len0 = Box * 1.001;       // forcetree.c:662-664
    CK(E->scratch_i.ensure(256));
    int *d_cnt = E->scratch_i.p;            // [0..2] split counters, [4] nvalid
    CK(cudaMemsetAsync(d_cnt, 0, 16 * sizeof(int), E->stream));

    const size_t nalloc = (size_t) (nin > 0 ? nin : 1);
    CK(E->keys.ensure(nalloc)); CK(E->keys_alt.ensure(nalloc));
    CK(E->sidx.ensure(nalloc)); CK(E->sidx_alt.ensure(nalloc));

    timer_start(E, T_TREE_KEYS);
    if(nin > 0) {
        k_tree_keys<<<(unsigned) ((nin + 255) / 256), 256, 0, E->stream>>>(E->pos.p, E->type.p, E->flags.p, d_active, nin,
                                                                         c0, len0, Box, toplevel_depth, mask, E->keys_alt.p, E->sidx_alt.p, d_cnt + 4,
                                                                         t_dau, t_key, t_shift, t_tab);
        CKL(E);
    }
    timer_stop(E, T_TREE_KEYS);

    timer_start(E, T_TREE_SORT);
    if(nin > 0) {
        size_t tb = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tb, E->keys_alt.p, E->keys.p, E->sidx_alt.p, E->sidx.p, (int) nin, 0, 64, E->stream);
        CK(E->cubtemp.ensure(tb + 16));
        CK(cub::DeviceRadixSort::SortPairs(E->cubtemp.p, tb, E->keys_alt.p, E->keys.p, E->sidx_alt.p, E->sidx.p, (int) nin, 0, 64, E->stream));
        E->launches += 8;    // histogram + onesweep passes
    }
    timer_stop(E, T_TREE_SORT);

    int h_np = 0;
    CK(cudaMemcpyAsync(&h_np, d_cnt + 4, sizeof(int), cudaMemcpyDeviceToHost, E->stream));
    CK(cudaStreamSynchronize(E->stream));
    const int np = h_np;

    timer_start(E, T_TREE_NODES);
    int cap = (int) (np * 0.75) + 4096;
    { int64_t top = 1; for(int l = 0; l < toplevel_depth; l++) top *= 8; cap += (int) (top * 2.5); }
    if(topn) cap += 2 * E->dk_ntop;
    std::vector<int> lvl;       // level offsets in BFS numbering
    int nn = 0, overfull = 0;
    for(int attempt = 0; attempt < 8; attempt++) {
        CK(E->b_start.ensure(cap)); CK(E->b_count.ensure(cap)); CK(E->b_father.ensure(cap));
        CK(E->b_firstchild.ensure(cap)); CK(E->b_nchild.ensure(cap));
        CK(E->b_size.ensure(cap)); CK(E->b_dfs.ensure(cap));
        CK(E->b_center.ensure(4 * (size_t) cap));
        if(topn) CK(E->b_top.ensure(cap));
        k_tree_root<<<1, 1, 0, E->stream>>>(np, c0, len0, E->b_start.p, E->b_count.p, E->b_father.p, E->b_center.p, d_cnt, topn ? E->b_top.p : nullptr);
        CKL(E);
        lvl.clear(); lvl.push_back(0); lvl.push_back(1);
        int h[3] = {1, 0, 0};
        for(int level = 0; level <= KEY_LEVELS; level++) {
            const int first = lvl[level], last = lvl[level + 1];
            if(last == first) break;
            k_tree_split<<<(last - first + 127) / 128, 128, 0, E->stream>>>(E->keys.p, first, last, level, toplevel_depth, cap, d_cnt,
                                                                        E->b_start.p, E->b_count.p, E->b_father.p,
                                                                        E->b_firstchild.p, E->b_nchild.p, E->b_center.p,
                                                                        topn ? E->b_top.p : nullptr, t_dau, t_tab);
            CKL(E);
            CK(cudaMemcpyAsync(h, d_cnt, 3 * sizeof(int), cudaMemcpyDeviceToHost, E->stream));
            CK(cudaStreamSynchronize(E->stream));
            if(h[1]) break;
            lvl.push_back(h[0]);
        }
        if(!h[1]) { nn = h[0]; overfull = h[2]; break; }
        cap = (int) (cap * 1.6) + 4096;     // the reference grows TreeAllocFactor by 1.15 and retries, forcetree.c:215-231
        nn = 0;
    }
    if(nn == 0) return failmsg(E, "b200_tree_build: could not allocate enough tree nodes");
    while(lvl.size() >= 2 && lvl[lvl.size() - 1] == lvl[lvl.size() - 2]) lvl.pop_back();
    const int nlevels = (int) lvl.size() - 1;
    timer_stop(E, T_TREE_NODES);

    timer_start(E, T_TREE_MOMENTS);
    k_tree_leaf_order<<<(nn + 127) / 128, 128, 0, E->stream>>>(nn, E->b_start.p, E->b_count.p, E->b_nchild.p, E->sidx.p);
    CKL(E);
    // 16 far-away massless rows follow the particles: the pair loop of the walk reads
    // whole 8-row pieces and empty padding pieces without bounds checks.
    CK(E->spart.ensure(4 * (size_t) (np + B200_SPART_PAD)));
    CK(E->spart_xy.ensure(2 * (size_t) (np + B200_SPART_PAD))); CK(E->spart_zm.ensure(2 * (size_t) (np + B200_SPART_PAD)));
    k_tree_gather<<<(np + B200_SPART_PAD + 255) / 256, 256, 0, E->stream>>>(E->pos.p, E->mass.p, E->sidx.p, np, 1e3 * Box, (double4 *) E->spart.p,
                                                                            (double2 *) E->spart_xy.p, (double2 *) E->spart_zm.p);
    CKL(E);
    // b_mom lives in nodeH's neighbour buffer: reuse keys_alt (np*8 bytes is too small) -> own buffer
    CK(E->nodeA.ensure(4 * (size_t) nn)); CK(E->nodeB.ensure(4 * (size_t) nn));
    CK(E->nodeC.ensure(4 * (size_t) nn)); CK(E->nodeF.ensure(nn)); CK(E->nodeH.ensure(nn));
    CK(E->nodeK.ensure(8 * (size_t) nn));
    CK(E->b_scan.ensure(8 * (size_t) nn));       // used as double4 b_mom storage (32 B per node)
    double4 *b_mom = (double4 *) E->b_scan.p;
    for(int level = nlevels - 1; level >= 0; level--) {
        const int first = lvl[level], last = lvl[level + 1];
        k_tree_moments<<<(last - first + 127) / 128, 128, 0, E->stream>>>(first, last, E->b_start.p, E->b_count.p, E->b_firstchild.p,
                                                                      E->b_nchild.p, E->b_center.p, (const double4 *) E->spart.p,
                                                                      b_mom, E->b_size.p);
        CKL(E);
    }
    CK(cudaMemsetAsync(E->b_dfs.p, 0, sizeof(int), E->stream));
    for(int level = 0; level < nlevels; level++) {
        const int first = lvl[level], last = lvl[level + 1];
        k_tree_dfs<<<(last - first + 127) / 128, 128, 0, E->stream>>>(first, last, E->b_firstchild.p, E->b_nchild.p, E->b_size.p, E->b_dfs.p);
        CKL(E);
    }
    k_tree_scatter<<<(nn + 255) / 256, 256, 0, E->stream>>>(nn, E->b_dfs.p, E->b_size.p, E->b_start.p, E->b_count.p, E->b_nchild.p,
                                                        E->b_firstchild.p, E->b_father.p, (const double4 *) E->b_center.p, b_mom,
                                                        (double4 *) E->nodeA.p, (double4 *) E->nodeB.p, (int4 *) E->nodeC.p,
                                                        E->nodeF.p, E->nodeH.p, (int4 *) E->nodeK.p);
    CKL(E);
    timer_stop(E, T_TREE_MOMENTS);

    E->tree_valid = true;
    E->tree_box = Box;
    E->tree_np = np;
    E->tree_nn = nn;
    E->tree_maxdepth = nlevels - 1;
    E->tree_overfull = overfull;
    E->tree_full = (d_active == nullptr);
    E->tree_topdepth = toplevel_depth;
    E->tree_lvl = lvl;
    E->sph_density_done = false;
    if(info) {
        double root[4];
        CK(cudaMemcpyAsync(root, E->nodeA.p, 4 * sizeof(double), cudaMemcpyDeviceToHost, E->stream));
        CK(cudaStreamSynchronize(E->stream));
        info->numnodes = nn;
        info->numparticles = np;
        info->maxdepth = nlevels - 1;
        info->overfull_leaves = overfull;
        info->root_mass = root[3];
    }
    return 0;
}

// ---- top-tree moments (multi-GPU): the analogue of force_exchange_pseudodata +
// force_treeupdate_pseudos (forcetree.c:1156-1284).  Level-`level` cells of the
// forced top tree are read out / overwritten as {cofm.xyz, mass}, indexed by
// Morton cell index; the levels above are then re-summed from their 8 children in
// octant order with the reference's arithmetic.
__global__ void __launch_bounds__(256)
k_top_get(int off, int ncell, const int *__restrict__ b_dfs, const double4 *__restrict__ nodeA, double4 *__restrict__ out)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if(m < ncell) out[m] = nodeA[b_dfs[off + m]];
}
__global__ void __launch_bounds__(256)
k_top_set(int off, int ncell, const int *__restrict__ b_dfs, double4 *__restrict__ nodeA, const double4 *__restrict__ in)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if(m < ncell) nodeA[b_dfs[off + m]] = in[m];
}
__global__ void __launch_bounds__(256)
k_top_resum(int off, int offchild, int ncell, const int *__restrict__ b_dfs, double4 *__restrict__ nodeA,
            const double4 *__restrict__ nodeB)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if(m >= ncell) return;
    double mass = 0, sx = 0, sy = 0, sz = 0;
    for(int k = 0; k < 8; k++) {
        const double4 c = nodeA[b_dfs[offchild + 8 * m + k]];
        mass = __dadd_rn(mass, c.w);
        sx = __dadd_rn(sx, __dmul_rn(c.w, c.x));
        sy = __dadd_rn(sy, __dmul_rn(c.w, c.y));
        sz = __dadd_rn(sz, __dmul_rn(c.w, c.z));
    }
    const int d = b_dfs[off + m];
    double4 o;
    if(mass > 0) { o.x = __ddiv_rn(sx, mass); o.y = __ddiv_rn(sy, mass); o.z = __ddiv_rn(sz, mass); }
    else { const double4 b = nodeB[d]; o.x = b.x; o.y = b.y; o.z = b.z; }
    o.w = mass;
    nodeA[d] = o;
}

static int top_offset(int level) { int64_t o = 0, c = 1; for(int l = 0; l < level; l++) { o += c; c *= 8; } return (int) o; }

int tree_top_get(Engine *E, int level, double *d_out)
{
    if(!E->tree_valid || level > E->tree_topdepth) return failmsg(E, "b200_tree_top_get: level is not inside the forced top tree");
    int ncell = 1; for(int l = 0; l < level; l++) ncell *= 8;
    k_top_get<<<(ncell + 255) / 256, 256, 0, E->stream>>>(top_offset(level), ncell, E->b_dfs.p, (const double4 *) E->nodeA.p, (double4 *) d_out);
    CKL(E);
    return 0;
}

int tree_top_set(Engine *E, int level, const double *d_in)
{
    if(!E->tree_valid || level > E->tree_topdepth) return failmsg(E, "b200_tree_top_set: level is not inside the forced top tree");
    int ncell = 1; for(int l = 0; l < level; l++) ncell *= 8;
    k_top_set<<<(ncell + 255) / 256, 256, 0, E->stream>>>(top_offset(level), ncell, E->b_dfs.p, (double4 *) E->nodeA.p, (const double4 *) d_in);
    CKL(E);
    for(int l = level - 1; l >= 0; l--) {
        ncell /= 8;
        k_top_resum<<<(ncell + 255) / 256, 256, 0, E->stream>>>(top_offset(l), top_offset(l + 1), ncell, E->b_dfs.p,
                                                             (double4 *) E->nodeA.p, (const double4 *) E->nodeB.p);
        CKL(E);
    }
    return 0;
}

int tree_export(Engine *E, double *center, double *len, double *cofm, double *mass, double *hmax,
                int32_t *sibling, int32_t *firstchild, int32_t *nocc, int32_t *leafpart)
{
    if(!E->tree_valid) return failmsg(E, "b200_tree_export: no tree");
    const int64_t nn = E->tree_nn, np = E->tree_np;
    std::vector<double> A(4 * nn), B(4 * nn), H(nn);
    std::vector<int> Cc(4 * nn), S(np > 0 ? np : 1);
    CK(cudaMemcpyAsync(A.data(), E->nodeA.p, 4 * nn * sizeof(double), cudaMemcpyDeviceToHost, E->stream));
    CK(cudaMemcpyAsync(B.data(), E->nodeB.p, 4 * nn * sizeof(double), cudaMemcpyDeviceToHost, E->stream));
    CK(cudaMemcpyAsync(H.data(), E->nodeH.p, nn * sizeof(double), cudaMemcpyDeviceToHost, E->stream));
    CK(cudaMemcpyAsync(Cc.data(), E->nodeC.p, 4 * nn * sizeof(int), cudaMemcpyDeviceToHost, E->stream));
    if(np > 0) CK(cudaMemcpyAsync(S.data(), E->sidx.p, np * sizeof(int), cudaMemcpyDeviceToHost, E->stream));
    CK(cudaStreamSynchronize(E->stream));
    for(int64_t d = 0; d < nn; d++) {
        if(center) for(int j = 0; j < 3; j++) center[3 * d + j] = B[4 * d + j];
        if(len) len[d] = B[4 * d + 3];
        if(cofm) for(int j = 0; j < 3; j++) cofm[3 * d + j] = A[4 * d + j];
        if(mass) mass[d] = A[4 * d + 3];
        if(hmax) hmax[d] = H[d];
        const int sib = Cc[4 * d], ps = Cc[4 * d + 1], cnt = Cc[4 * d + 2], leaf = Cc[4 * d + 3];
        if(sibling) sibling[d] = sib;
        if(firstchild) firstchild[d] = leaf ? -1 : (int) d + 1;
        if(nocc) nocc[d] = leaf ? cnt : -1;
        if(leafpart) {
            for(int k = 0; k < 8; k++) leafpart[8 * d + k] = -1;
            if(leaf) for(int k = 0; k < cnt && k < 8; k++) leafpart[8 * d + k] = S[ps + k];
        }
    }
    return 0;
}

} // namespace b200
