// domain_keys.cu -- the keys of the reference's domain decomposition on the device (SURVEY.md 8f rank 2,
// first brick): the Peano-Hilbert key of every particle (PEANO(), libgadget/utils/peano.h:15-21 over
// peano_hilbert_key, utils/peano.c:108-129) and P[].TopLeaf = domain_get_topleaf(key) (libgadget/domain.h:71-78)
// for a top tree handed in as arrays.  Integer work, bit-exact against the reference's own test vector
// (tests/test_peano.c:107) and its compiled peano.c.
//
// The curve is stated geometrically and its state machine generated at first use (no stored tables): a state
// is a symmetry of the cube applied to one base pattern -- octants in the Gray-code order
// 000,010,110,100,101,111,011,001 (bits x,y,z) -- and the sub-cube visited r-th continues in the state composed
// with the r-th of eight child symmetries (swap y,z | swap x,z twice | flip x,y twice | swap x,z + flip x,z twice |
// swap y,z + flip y,z).  Both kernels are one thread per particle and HBM-bound (24 B in, 8 B out; 8 B in, 4 B out).
#include <cub/cub.cuh>
#include <string.h>
#include <algorithm>
#include "engine.h"

namespace b200 {

#define PH_BITS 21               // BITS_PER_DIMENSION, peano.h:11

struct CubeSym { int perm[3], flip[3]; };      // new bit k = old bit perm[k] ^ flip[k]
static const int ph_base_order[8] = {0, 2, 6, 4, 5, 7, 3, 1};
static const CubeSym ph_child[8] = {
    {{0, 2, 1}, {0, 0, 0}}, {{2, 1, 0}, {0, 0, 0}}, {{2, 1, 0}, {0, 0, 0}}, {{0, 1, 2}, {1, 1, 0}},
    {{0, 1, 2}, {1, 1, 0}}, {{2, 1, 0}, {1, 0, 1}}, {{2, 1, 0}, {1, 0, 1}}, {{0, 2, 1}, {0, 1, 1}}};

static int ph_apply(const CubeSym &s, int pix)
{
    const int b[3] = {(pix >> 2) & 1, (pix >> 1) & 1, pix & 1};
    return ((b[s.perm[0]] ^ s.flip[0]) << 2) | ((b[s.perm[1]] ^ s.flip[1]) << 1) | (b[s.perm[2]] ^ s.flip[2]);
}
// tab[0..383] = rank[state][octant], tab[384..767] = next[state][octant]; returns the number of states: 24, the
// rotations reachable from the identity (the reference's tables also list their 24 mirror images, which the walk
// from state 0 never enters)
static int ph_tables(uint8_t *tab)
{
    CubeSym st[48];
    int ns = 1;
    st[0] = CubeSym{{0, 1, 2}, {0, 0, 0}};
    for(int s = 0; s < ns; s++)
        for(int r = 0; r < 8; r++) {
            const int pix = ph_apply(st[s], ph_base_order[r]);
            CubeSym c;
            for(int k = 0; k < 3; k++) { c.perm[k] = ph_child[r].perm[st[s].perm[k]]; c.flip[k] = ph_child[r].flip[st[s].perm[k]] ^ st[s].flip[k]; }
            int f = -1;
            for(int q = 0; q < ns; q++) if(!memcmp(&st[q], &c, sizeof(c))) f = q;
            if(f < 0) { if(ns == 48) return -1; st[ns] = c; f = ns++; }
            tab[s * 8 + pix] = (uint8_t) r; tab[384 + s * 8 + pix] = (uint8_t) f;
        }
    return ns;
}

__global__ void __launch_bounds__(256)
k_domain_keys(int64_t n, int64_t stride, const double *__restrict__ pos, double Box, double fac, const uint8_t *__restrict__ tab,
              unsigned long long *__restrict__ keys)      // keys[i] = key of particle i * stride
{
    __shared__ uint8_t s_tab[768];
    for(int k = threadIdx.x; k < 768; k += blockDim.x) s_tab[k] = tab[k];
    __syncthreads();
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    const double off = Box / 2000;                                         // peano.h:19
    const double *p = pos + 3 * i * stride;
    const int x = (int) ((p[0] + off) * fac), y = (int) ((p[1] + off) * fac), z = (int) ((p[2] + off) * fac);
    unsigned long long key = 0;
    int s = 0;
#pragma unroll 1
    for(int bit = PH_BITS - 1; bit >= 0; bit--) {                          // peano.c:114-124
        const int pix = (((x >> bit) & 1) << 2) | (((y >> bit) & 1) << 1) | ((z >> bit) & 1);
        key = (key << 3) | s_tab[s * 8 + pix];
        s = s_tab[384 + s * 8 + pix];
    }
    keys[i] = key;
}

__global__ void __launch_bounds__(256)
k_domain_topleaf(int64_t n, const unsigned long long *__restrict__ keys, const int *__restrict__ daughter,
                 const unsigned long long *__restrict__ startkey, const int *__restrict__ shift, const int *__restrict__ leaf, int *__restrict__ out)
{
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    const unsigned long long key = keys[i];
    int no = 0;
    while(daughter[no] >= 0) no = daughter[no] + (int) ((key - startkey[no]) >> (shift[no] - 3));     // domain.h:74-76
    out[i] = leaf[no];
}

// TopLeafCount of domain_compute_costs (domain.c:1396-1451): particles per top leaf, garbage skipped
__global__ void __launch_bounds__(256)
k_domain_leaf_counts(int64_t n, const int *__restrict__ topleaf, const uint8_t *__restrict__ flags, int nleaf, unsigned long long *__restrict__ counts)
{
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n || (flags[i] & 1)) return;
    const int l = topleaf[i];
    if(l >= 0 && l < nleaf) atomicAdd(&counts[l], 1ull);
}

int domain_need_tables(Engine *E)
{
    if(E->dk_tab.p) return 0;
    uint8_t tab[768];
    if(ph_tables(tab) != 24) return failmsg(E, "b200_domain_peano_keys: state machine generation failed");
    CK(E->dk_tab.ensure(768));
    CK(cudaMemcpyAsync(E->dk_tab.p, tab, 768, cudaMemcpyHostToDevice, E->stream));
    CK(cudaStreamSynchronize(E->stream));      // tab is a stack array
    return 0;
}

// The subsample domain_check_for_local_refine_subsample builds its skeleton from (PreSort = 0, domain.c:1066-1074):
// the key of every `subsample`-th particle in memory order; only these 8 N / subsample bytes go to the host.
int domain_sample_keys(Engine *E, double BoxSize, int32_t subsample, uint64_t *keys_out, int64_t *nsample_out)
{
    if(!(BoxSize > 0) || subsample < 1 || !keys_out || !nsample_out) return failmsg(E, "b200_domain_sample_keys: bad arguments");
    int64_t ns = E->n / subsample;
    if(ns == 0 && E->n != 0) ns = 1;                                       // :1033-1035
    *nsample_out = ns;
    if(ns == 0) return 0;
    if(int rc = domain_need_tables(E)) return rc;
    CK(E->dk_sample.ensure((size_t) ns));
    const double fac = 1.0 / (BoxSize * 1.001) * (double) (1ull << PH_BITS);
    k_domain_keys<<<(unsigned) ((ns + 255) / 256), 256, 0, E->stream>>>(ns, subsample, E->pos.p, BoxSize, fac, E->dk_tab.p, E->dk_sample.p);
    CKL(E);
    CK(cudaMemcpyAsync(keys_out, E->dk_sample.p, (size_t) ns * sizeof(uint64_t), cudaMemcpyDeviceToHost, E->stream));
    CK(cudaStreamSynchronize(E->stream));
    return 0;
}

// domain_build_exchange_list + domain_build_plan (exchange.c:408-444,505-530) with domain_layoutfunc (domain.c:794-803):
// flag[i] = particle i leaves this task; cnt[0] garbage, cnt[1] bad leaf / task, cnt[2 + 7 * target + {0, 1 + type}] = toGo
__global__ void __launch_bounds__(256)
k_domain_exchange_flags(int64_t n, const uint8_t *__restrict__ type, const uint8_t *__restrict__ flags, const int *__restrict__ topleaf,
                        const int *__restrict__ task_of_leaf, int nleaf, int ntask, int thistask, uint8_t *__restrict__ flag,
                        unsigned long long *__restrict__ cnt)
{
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    uint8_t f = 0;
    if(flags[i] & 1) atomicAdd(&cnt[0], 1ull);
    else {
        const int l = topleaf[i];
        const int target = (l >= 0 && l < nleaf) ? task_of_leaf[l] : -1;
        if(target < 0 || target >= ntask) atomicAdd(&cnt[1], 1ull);
        else if(target != thistask) {
            f = 1;
            atomicAdd(&cnt[2 + 7 * target], 1ull);
            atomicAdd(&cnt[2 + 7 * target + 1 + (type[i] < 6 ? type[i] : 5)], 1ull);
        }
    }
    flag[i] = f;
}
__global__ void k_domain_iota(int *p, int64_t n)
{
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if(i < n) p[i] = (int) i;
}

int domain_peano_keys(Engine *E, double BoxSize, uint64_t *keys_out)
{
    if(!(BoxSize > 0)) return failmsg(E, "b200_domain_peano_keys: bad box size");
    const size_t n = (size_t) (E->n > 0 ? E->n : 1);
    CK(E->dk_keys.ensure(n));
    if(int rc = domain_need_tables(E)) return rc;
    const double fac = 1.0 / (BoxSize * 1.001) * (double) (1ull << PH_BITS);      // peano.h:18
    if(E->n > 0) {
        k_domain_keys<<<(unsigned) ((E->n + 255) / 256), 256, 0, E->stream>>>(E->n, 1, E->pos.p, BoxSize, fac, E->dk_tab.p, E->dk_keys.p);
        CKL(E);
        if(keys_out) CK(cudaMemcpyAsync(keys_out, E->dk_keys.p, (size_t) E->n * sizeof(uint64_t), cudaMemcpyDeviceToHost, E->stream));
    }
    CK(cudaStreamSynchronize(E->stream));
    E->dk_keys_n = E->n;
    return 0;
}

int domain_set_topnodes(Engine *E, int32_t ntop, const int32_t *daughter, const uint64_t *startkey, const int32_t *shift, const int32_t *leaf)
{
    if(ntop < 1 || !daughter || !startkey || !shift || !leaf) return failmsg(E, "b200_domain_set_topnodes: bad arguments");
    for(int t = 0; t < ntop; t++)      // the lookup must terminate: daughters lie behind their parent and inside the table
        if(daughter[t] >= 0 && (daughter[t] <= t || daughter[t] + 8 > ntop || shift[t] < 3)) return failmsg(E, "b200_domain_set_topnodes: malformed top tree");
    CK(E->dk_daughter.ensure(ntop)); CK(E->dk_startkey.ensure(ntop)); CK(E->dk_shift.ensure(ntop)); CK(E->dk_leaf.ensure(ntop));
    CK(cudaMemcpyAsync(E->dk_daughter.p, daughter, ntop * sizeof(int32_t), cudaMemcpyHostToDevice, E->stream));
    CK(cudaMemcpyAsync(E->dk_startkey.p, startkey, ntop * sizeof(uint64_t), cudaMemcpyHostToDevice, E->stream));
    CK(cudaMemcpyAsync(E->dk_shift.p, shift, ntop * sizeof(int32_t), cudaMemcpyHostToDevice, E->stream));
    CK(cudaMemcpyAsync(E->dk_leaf.p, leaf, ntop * sizeof(int32_t), cudaMemcpyHostToDevice, E->stream));
    CK(cudaStreamSynchronize(E->stream));
    E->dk_ntop = ntop;
    return 0;
}

int domain_topleaf(Engine *E, int32_t *topleaf_out)
{
    if(E->dk_ntop == 0) return failmsg(E, "b200_domain_topleaf: call b200_domain_set_topnodes first");
    if(E->dk_keys_n != E->n) return failmsg(E, "b200_domain_topleaf: call b200_domain_peano_keys first");
    const size_t n = (size_t) (E->n > 0 ? E->n : 1);
    CK(E->dk_topleaf.ensure(n));
    if(E->n > 0) {
        k_domain_topleaf<<<(unsigned) ((E->n + 255) / 256), 256, 0, E->stream>>>(E->n, E->dk_keys.p, E->dk_daughter.p, E->dk_startkey.p, E->dk_shift.p,
                                                                               E->dk_leaf.p, E->dk_topleaf.p);
        CKL(E);
        if(topleaf_out) CK(cudaMemcpyAsync(topleaf_out, E->dk_topleaf.p, (size_t) E->n * sizeof(int32_t), cudaMemcpyDeviceToHost, E->stream));
    }
    CK(cudaStreamSynchronize(E->stream));
    E->dk_topleaf_n = E->n;
    return 0;
}

int domain_leaf_counts(Engine *E, int32_t nleaf, int64_t *counts_out)
{
    if(nleaf < 1 || !counts_out) return failmsg(E, "b200_domain_leaf_counts: bad arguments");
    if(E->dk_topleaf_n != E->n) return failmsg(E, "b200_domain_leaf_counts: call b200_domain_topleaf first");
    CK(E->dk_counts.ensure((size_t) nleaf));
    CK(cudaMemsetAsync(E->dk_counts.p, 0, (size_t) nleaf * sizeof(unsigned long long), E->stream));
    if(E->n > 0) {
        k_domain_leaf_counts<<<(unsigned) ((E->n + 255) / 256), 256, 0, E->stream>>>(E->n, E->dk_topleaf.p, E->flags.p, nleaf, E->dk_counts.p);
        CKL(E);
    }
    CK(cudaMemcpyAsync(counts_out, E->dk_counts.p, (size_t) nleaf * sizeof(int64_t), cudaMemcpyDeviceToHost, E->stream));
    CK(cudaStreamSynchronize(E->stream));
    return 0;
}

// The exchange plan of this rank from the last b200_domain_topleaf: which particles leave (the list stays on the device in
// dk_xlist, ascending index) and how many go to every task in total and per type.
int domain_exchange_plan(Engine *E, const int32_t *task_of_leaf, int32_t nleaf, int32_t ntask, int32_t thistask, int64_t *nexchange,
                         int64_t *ngarbage, int64_t *togo, int32_t *list_out)
{
    if(!task_of_leaf || nleaf < 1 || ntask < 1 || thistask < 0 || thistask >= ntask || !nexchange || !togo)
        return failmsg(E, "b200_domain_exchange_plan: bad arguments");
    if(E->dk_topleaf_n != E->n) return failmsg(E, "b200_domain_exchange_plan: call b200_domain_topleaf first");
    const size_t n = (size_t) (E->n > 0 ? E->n : 1), nc = 2 + 7 * (size_t) ntask;
    CK(E->dk_task.ensure((size_t) nleaf)); CK(E->dk_counts.ensure(nc)); CK(E->dk_xflag.ensure(n + 64)); CK(E->dk_xlist.ensure(n + 1));
    CK(E->dk_iota.ensure(n));
    CK(cudaMemcpyAsync(E->dk_task.p, task_of_leaf, (size_t) nleaf * sizeof(int32_t), cudaMemcpyHostToDevice, E->stream));
    CK(cudaMemsetAsync(E->dk_counts.p, 0, nc * sizeof(unsigned long long), E->stream));
    *nexchange = 0;
    if(ngarbage) *ngarbage = 0;
    for(size_t k = 0; k < 7 * (size_t) ntask; k++) togo[k] = 0;
    if(E->n == 0) { CK(cudaStreamSynchronize(E->stream)); return 0; }
    const unsigned grid = (unsigned) ((E->n + 255) / 256);
    k_domain_exchange_flags<<<grid, 256, 0, E->stream>>>(E->n, E->type.p, E->flags.p, E->dk_topleaf.p, E->dk_task.p, nleaf, ntask, thistask,
                                                        E->dk_xflag.p, E->dk_counts.p);
    CKL(E);
    k_domain_iota<<<grid, 256, 0, E->stream>>>(E->dk_iota.p, E->n); CKL(E);
    CK(E->scratch_i.ensure(256));
    int *d_num = E->scratch_i.p + 24;
    size_t tb = 0;
    cub::DeviceSelect::Flagged(nullptr, tb, E->dk_iota.p, E->dk_xflag.p, E->dk_xlist.p, d_num, (int) E->n, E->stream);
    CK(E->cubtemp.ensure(tb + 16));
    CK(cub::DeviceSelect::Flagged(E->cubtemp.p, tb, E->dk_iota.p, E->dk_xflag.p, E->dk_xlist.p, d_num, (int) E->n, E->stream));
    E->launches += 1;
    std::vector<unsigned long long> h(nc);
    int cnt = 0;
    CK(cudaMemcpyAsync(h.data(), E->dk_counts.p, nc * sizeof(unsigned long long), cudaMemcpyDeviceToHost, E->stream));
    CK(cudaMemcpyAsync(&cnt, d_num, sizeof(int), cudaMemcpyDeviceToHost, E->stream));
    CK(cudaStreamSynchronize(E->stream));
    if(h[1]) return failmsg(E, "b200_domain_exchange_plan: " + std::to_string(h[1]) + " particles with a top leaf or task out of range (domain.c:797-799)");
    *nexchange = cnt;
    if(ngarbage) *ngarbage = (int64_t) h[0];
    for(size_t k = 0; k < 7 * (size_t) ntask; k++) togo[k] = (int64_t) h[2 + k];
    if(list_out && cnt > 0) {
        CK(cudaMemcpyAsync(list_out, E->dk_xlist.p, (size_t) cnt * sizeof(int32_t), cudaMemcpyDeviceToHost, E->stream));
        CK(cudaStreamSynchronize(E->stream));
    }
    return 0;
}

// domain_assign_topleaves_balanced (domain.c:610-755) over leaves in key order: host arithmetic on NTopLeaves numbers.
// Returns 0, or 1 where the reference would endrun (fewer segments than asked for, cost not fully assigned).
int domain_assign_balanced(int ntask, int32_t nleaf, const int64_t *cost, int nseg_per_task, int32_t *task)
{
    if(ntask < 1 || nseg_per_task < 1 || nleaf < ntask * nseg_per_task || !cost || !task) return 1;
    const int nsegment = ntask * nseg_per_task;
    int64_t total = 0;
    for(int32_t i = 0; i < nleaf; i++) { total += cost[i]; task[i] = -1; }
    int64_t left = total;
    double mean_expected = 1.0 * total / nsegment, mean_task = 1.0 * total / ntask;
    int curleaf = 0, curseg = 0, curtask = 0, nrounds = 0;
    int64_t curload = 0, curtaskload = 0;
    while(nrounds < nleaf) {
        bool append = false, advance = false;
        if(curleaf == nleaf) advance = true;
        else if(nleaf - curleaf == nsegment - curseg) append = advance = true;       // one leaf per remaining segment
        else {
            const int64_t assigned = (total - left) + curload;
            if(mean_expected * (curseg + 1) - assigned > 0.5 * cost[curleaf] || curload == 0) append = true;
            else advance = true;
        }
        if(append) { curload += cost[curleaf]; task[curleaf] = curtask; curleaf++; }
        if(advance) {
            curtaskload += curload;
            if(mean_task - curtaskload < 0.5 * mean_expected || nsegment - curseg <= ntask - curtask) { curtaskload = 0; curtask++; }
            left -= curload;
            curload = 0;
            curseg++;
            if(curtask == ntask) {
                curtask = 0;
                mean_expected = 1.0 * left / nsegment;
                mean_task = 1.0 * left / ntask;
                nrounds++;
            }
            if(curleaf == nleaf) break;
        }
    }
    return (curseg < nsegment || left != 0) ? 1 : 0;
}

// ---- the top tree (domain.c:826-1395) on the host: it has a few hundred to a few thousand nodes and is built from the
// N / 256 subsample keys, so the device's part is the keys; the stages below are sequential integer work. ----
namespace toptree {
using Node = b200_topnode;

static int locate(const Node *t, uint64_t key)                            // domain_toptree_get_subnode :826-835
{
    int no = 0;
    while(t[no].Daughter >= 0) no = t[no].Daughter + (int) ((key - t[no].StartKey) >> (t[no].Shift - 3));
    return no;
}
static void make_children(Node *t, int parent, int first, int64_t count8, int64_t cost8, bool divide)
{
    for(int j = 0; j < 8; j++) {
        Node &c = t[first + j];
        c.Shift = t[parent].Shift - 3; c.pad_ = 0; c.Daughter = -1; c.Parent = parent;
        c.StartKey = t[parent].StartKey + (uint64_t) j * (1ull << c.Shift);
        c.Count = divide ? (j + 1) * count8 / 8 - j * count8 / 8 : count8;
        c.Cost = divide ? (j + 1) * cost8 / 8 - j * cost8 / 8 : cost8;
    }
}
static void accumulate(Node *t, int no)                                   // domain_toptree_update_cost :885-897
{
    if(t[no].Daughter < 0) return;
    for(int j = 0; j < 8; j++) {
        const int c = t[no].Daughter + j;
        accumulate(t, c);
        t[no].Count += t[c].Count; t[no].Cost += t[c].Cost;
    }
}
// domain_check_for_local_refine_subsample :1084-1187 from the sorted keys on.  0 ok, 1 out of nodes, 2 too clustered / unsorted
static int local(const uint64_t *keys, int64_t ns, Node *t, int32_t *size, int32_t maxnodes)
{
    *size = 1;
    memset(&t[0], 0, sizeof(Node));
    t[0].Daughter = -1; t[0].Parent = -1; t[0].Shift = 3 * PH_BITS;
    uint64_t prev_key = ~0ull;
    int prev_leaf = -1;
    for(int64_t i = 0; i < ns;) {
        const int leaf = locate(t, keys[i]);
        if(leaf == prev_leaf && t[leaf].Shift >= 3) {                     // two samples in one leaf: refine it, re-seat the previous one
            if(*size + 8 > maxnodes) return 1;
            t[leaf].Daughter = *size;
            make_children(t, leaf, *size, 0, 0, false);
            *size += 8;
            t[leaf].Count = 0;
            prev_leaf = locate(t, prev_key);
            t[prev_leaf].Count++;
            continue;
        }
        if(t[leaf].Count != 0 && leaf != prev_leaf) return 2;
        prev_key = keys[i]; prev_leaf = leaf;
        t[leaf].Count++;
        i++;
    }
    for(int k = 0; k < *size; k++) t[k].Count = 0;
    for(int64_t i = 0; i < ns; i++) { Node &l = t[locate(t, keys[i])]; l.Count++; l.Cost++; }
    accumulate(t, 0);
    return 0;
}
static void prune(Node *t, int no, int64_t countlimit, int64_t costlimit)  // domain_toptree_truncate_r :899-916
{
    if(t[no].Daughter < 0) return;
    if(t[no].Count < countlimit && t[no].Cost < costlimit) { t[no].Daughter = -1; return; }
    for(int j = 0; j < 8; j++) prune(t, t[no].Daughter + j, countlimit, costlimit);
}
static void repack(Node *t, int no, int32_t *next)                         // domain_toptree_garbage_collection :927-951
{
    if(t[no].Daughter < 0) return;
    const int from = t[no].Daughter, to = *next;
    t[no].Daughter = to;
    *next += 8;
    for(int j = 0; j < 8; j++) { t[to + j] = t[from + j]; t[to + j].Parent = no; }
    for(int j = 0; j < 8; j++) repack(t, to + j, next);
}
static int merge(Node *A, const Node *B, int a, int b, int32_t *sizeA, int32_t maxnodes)       // domain_toptree_merge :1473-1577
{
    if(B[b].Shift < A[a].Shift) {
        if(A[a].Daughter < 0) {
            if(*sizeA + 8 >= maxnodes) return 1;
            A[a].Daughter = *sizeA;
            make_children(A, a, *sizeA, A[a].Count - B[B[b].Parent].Count, A[a].Cost - B[B[b].Parent].Cost, true);
            *sizeA += 8;
        }
        return merge(A, B, A[a].Daughter + (int) ((B[b].StartKey - A[a].StartKey) >> (A[a].Shift - 3)), b, sizeA, maxnodes);
    }
    if(B[b].Shift == A[a].Shift) {
        A[a].Count += B[b].Count; A[a].Cost += B[b].Cost;
        if(B[b].Daughter >= 0) { for(int j = 0; j < 8; j++) if(merge(A, B, a, B[b].Daughter + j, sizeA, maxnodes)) return 1; }
        else if(A[a].Daughter >= 0) { for(int j = 0; j < 8; j++) if(merge(A, B, A[a].Daughter + j, b, sizeA, maxnodes)) return 1; }
        return 0;
    }
    const int up = B[b].Shift - A[a].Shift;                                // B's node covers 2^up cells of A's size
    if(up > 60) return 0;
    const uint64_t cells = 1ull << up;                                     // the reference divides in unsigned arithmetic (peano_t n)
    A[a].Count += (int64_t) ((uint64_t) B[b].Count / cells); A[a].Cost += (int64_t) ((uint64_t) B[b].Cost / cells);
    if(A[a].Daughter >= 0) for(int j = 0; j < 8; j++) if(merge(A, B, A[a].Daughter + j, b, sizeA, maxnodes)) return 1;
    return 0;
}
static void number_leaves(const Node *t, int no, int32_t *next, int32_t *leaf)   // domain_create_topleaves :810-824
{
    if(t[no].Daughter < 0) { leaf[no] = (*next)++; return; }
    for(int j = 0; j < 8; j++) number_leaves(t, t[no].Daughter + j, next, leaf);
}
} // namespace toptree

void domain_release(Engine *E)
{
    E->b_top.release(); E->dk_keys.release(); E->dk_tab.release(); E->dk_daughter.release(); E->dk_startkey.release(); E->dk_shift.release(); E->dk_leaf.release(); E->dk_topleaf.release(); E->dk_counts.release(); E->dk_sample.release(); E->dk_xflag.release(); E->dk_xlist.release(); E->dk_iota.release(); E->dk_task.release();
}

} // namespace b200

using namespace b200;
#define DK_ENTER(ctx) if(!(ctx)) return 1; Engine *E = &(ctx)->e; if(cudaSetDevice(E->device) != cudaSuccess) return failmsg(E, "cudaSetDevice failed")
extern "C" {
int b200_domain_peano_keys(b200_ctx *ctx, double BoxSize, uint64_t *keys_out) { DK_ENTER(ctx); return domain_peano_keys(E, BoxSize, keys_out); }
int b200_domain_set_topnodes(b200_ctx *ctx, int32_t ntop, const int32_t *daughter, const uint64_t *startkey, const int32_t *shift, const int32_t *leaf)
{
    DK_ENTER(ctx);
    return domain_set_topnodes(E, ntop, daughter, startkey, shift, leaf);
}
int b200_domain_topleaf(b200_ctx *ctx, int32_t *topleaf_out) { DK_ENTER(ctx); return domain_topleaf(E, topleaf_out); }
int b200_domain_leaf_counts(b200_ctx *ctx, int32_t nleaf, int64_t *counts_out) { DK_ENTER(ctx); return domain_leaf_counts(E, nleaf, counts_out); }
int b200_domain_exchange_plan(b200_ctx *ctx, const int32_t *task_of_leaf, int32_t nleaf, int32_t ntask, int32_t thistask, int64_t *nexchange,
                              int64_t *ngarbage, int64_t *togo, int32_t *list_out)
{
    DK_ENTER(ctx);
    return domain_exchange_plan(E, task_of_leaf, nleaf, ntask, thistask, nexchange, ngarbage, togo, list_out);
}
int b200_domain_sample_keys(b200_ctx *ctx, double BoxSize, int32_t subsample, uint64_t *keys_out, int64_t *nsample)
{
    DK_ENTER(ctx);
    return domain_sample_keys(E, BoxSize, subsample, keys_out, nsample);
}
int b200_domain_toptree_local(uint64_t *sample_keys, int64_t nsample, b200_topnode *tree, int32_t *size, int32_t maxnodes)
{
    if(!tree || !size || maxnodes < 1 || nsample < 0 || (nsample > 0 && !sample_keys)) return 3;
    std::sort(sample_keys, sample_keys + nsample);                         // qsort_openmp(LP, ..., order_by_key), domain.c:1079
    return toptree::local(sample_keys, nsample, tree, size, maxnodes);
}
int b200_domain_toptree_truncate(b200_topnode *tree, int32_t *size, int64_t countlimit, int64_t costlimit)   // :953-966
{
    if(!tree || !size || *size < 1) return 3;
    toptree::prune(tree, 0, countlimit, costlimit);
    *size = 1;
    toptree::repack(tree, 0, size);
    return 0;
}
int b200_domain_toptree_merge(b200_topnode *treeA, int32_t *sizeA, const b200_topnode *treeB, int32_t maxnodes)
{
    if(!treeA || !sizeA || !treeB) return 3;
    return toptree::merge(treeA, treeB, 0, 0, sizeA, maxnodes);
}
int b200_domain_toptree_global_refine(b200_topnode *t, int32_t *size, int32_t maxnodes, int64_t countlimit, int64_t costlimit)   // :1343-1393
{
    if(!t || !size) return 3;
    for(int i = 0; i < *size; i++) {
        if(t[i].Daughter >= 0 || t[i].Shift <= 0) continue;
        if(t[i].Count < countlimit && t[i].Cost < costlimit) continue;
        if(*size + 8 > maxnodes) return 1;
        t[i].Daughter = *size;
        toptree::make_children(t, i, *size, t[i].Count / 8, t[i].Cost / 8, false);
        *size += 8;
    }
    return 0;
}
int b200_domain_toptree_leaves(const b200_topnode *tree, int32_t size, int32_t *leaf_out, int32_t *nleaf)
{
    if(!tree || size < 1 || !leaf_out || !nleaf) return 3;
    for(int i = 0; i < size; i++) leaf_out[i] = -1;
    *nleaf = 0;
    toptree::number_leaves(tree, 0, nleaf, leaf_out);
    return 0;
}
int b200_domain_assign_balanced(int32_t ntask, int32_t nleaf, const int64_t *cost, int32_t nseg_per_task, int32_t *task_out)
{
    return domain_assign_balanced(ntask, nleaf, cost, nseg_per_task, task_out);
}
}
