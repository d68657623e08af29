/* oracle_domain.c -- CPU restatement of the domain keys of the reference: the Peano-Hilbert key of a
 * position (libgadget/utils/peano.c:108-129, peano.h:15-21) and the top-leaf lookup
 * (libgadget/domain.h:71-78).  TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * PINNED against the 64 known-answer keys of the reference's own tests/test_peano.c:107-118
 * (tests/golden/ref_peano.npz) and against the reference's compiled peano.c / domain.h on random input.
 *
 * The reference walks a state machine stored as two 48x8 tables.  Here the same curve is stated
 * geometrically and the machine is generated: a state is a symmetry of the cube (axis permutation +
 * flips) applied to one base pattern -- octants visited in the Gray-code order
 * 000,010,110,100,101,111,011,001 (bits x,y,z) -- and the sub-cube visited r-th continues in the state
 * composed with the r-th of eight fixed child symmetries:
 *   r = 0: swap y,z | 1,2: swap x,z | 3,4: flip x,y | 5,6: swap x,z, flip x,z | 7: swap y,z, flip y,z.
 * Closing the identity under these gives 24 states (the rotations); the reference's 48-row tables also hold their
 * mirror images, which a walk from state 0 never enters. */
#include <string.h>
#include "oracle.h"

typedef struct { int perm[3], flip[3]; } sym_t;                 /* new bit k = old bit perm[k] ^ flip[k] */
static const int base_order[8] = {0, 2, 6, 4, 5, 7, 3, 1};      /* octant (4x+2y+z) visited at rank r */
static const sym_t child[8] = {
    {{0, 2, 1}, {0, 0, 0}}, {{2, 1, 0}, {0, 0, 0}}, {{2, 1, 0}, {0, 0, 0}}, {{0, 1, 2}, {1, 1, 0}},
    {{0, 1, 2}, {1, 1, 0}}, {{2, 1, 0}, {1, 0, 1}}, {{2, 1, 0}, {1, 0, 1}}, {{0, 2, 1}, {0, 1, 1}}};

static int apply(const sym_t *s, int pix)
{
    const int b[3] = {(pix >> 2) & 1, (pix >> 1) & 1, pix & 1};
    return ((b[s->perm[0]] ^ s->flip[0]) << 2) | ((b[s->perm[1]] ^ s->flip[1]) << 1) | (b[s->perm[2]] ^ s->flip[2]);
}
static sym_t compose(const sym_t *a, const sym_t *b)            /* b first, then a */
{
    sym_t c;
    for(int k = 0; k < 3; k++) { c.perm[k] = b->perm[a->perm[k]]; c.flip[k] = b->flip[a->perm[k]] ^ a->flip[k]; }
    return c;
}
/* rank[state][octant] and next[state][octant], states numbered in order of discovery from the identity */
int oracle_peano_tables(uint8_t rank[48][8], uint8_t next[48][8])
{
    sym_t st[48];
    int ns = 1;
    const sym_t id = {{0, 1, 2}, {0, 0, 0}};
    st[0] = id;
    for(int s = 0; s < ns; s++)
        for(int r = 0; r < 8; r++) {
            const int pix = apply(&st[s], base_order[r]);
            const sym_t c = compose(&st[s], &child[r]);
            int f = -1;
            for(int q = 0; q < ns; q++) if(!memcmp(&st[q], &c, sizeof(c))) f = q;
            if(f < 0) { if(ns == 48) return -1; st[ns] = c; f = ns++; }
            rank[s][pix] = (uint8_t) r; next[s][pix] = (uint8_t) f;
        }
    return ns;
}
uint64_t oracle_peano_key(int x, int y, int z, int bits)        /* peano.c:108-129 */
{
    static uint8_t rank[48][8], next[48][8];
    static int ready;
    if(!ready) { oracle_peano_tables(rank, next); ready = 1; }
    uint64_t key = 0;
    int s = 0;
    for(int bit = bits - 1; bit >= 0; bit--) {
        const int pix = (((x >> bit) & 1) << 2) | (((y >> bit) & 1) << 1) | ((z >> bit) & 1);
        key = (key << 3) | rank[s][pix];
        s = next[s][pix];
    }
    return key;
}
/* PEANO(Pos, BoxSize) peano.h:15-21, BITS_PER_DIMENSION = 21 */
void oracle_peano_keys(const double *pos, int64_t n, double BoxSize, uint64_t *keys)
{
    const double fac = 1.0 / (BoxSize * 1.001) * (double) (((uint64_t) 1) << 21);
    for(int64_t i = 0; i < n; i++) {
        const double sx = pos[3 * i] + BoxSize / 2000, sy = pos[3 * i + 1] + BoxSize / 2000, sz = pos[3 * i + 2] + BoxSize / 2000;
        keys[i] = oracle_peano_key((int) (sx * fac), (int) (sy * fac), (int) (sz * fac), 21);
    }
}
/* domain_get_topleaf domain.h:71-78 over TopNodes given as arrays */
void oracle_topleaf(const uint64_t *keys, int64_t n, const int32_t *daughter, const uint64_t *startkey, const int32_t *shift,
                    const int32_t *leaf, int32_t *out)
{
    for(int64_t i = 0; i < n; i++) {
        int no = 0;
        while(daughter[no] >= 0) no = daughter[no] + (int) ((keys[i] - startkey[no]) >> (shift[no] - 3));
        out[i] = leaf[no];
    }
}

/* domain_compute_costs domain.c:1398-1470, the count part: particles per top leaf (garbage skipped) */
void oracle_leaf_counts(const int32_t *topleaf, const uint8_t *flags, int64_t n, int32_t nleaf, int64_t *counts)
{
    for(int32_t l = 0; l < nleaf; l++) counts[l] = 0;
    for(int64_t i = 0; i < n; i++) {
        if(flags && (flags[i] & 1)) continue;                   /* IsGarbage, domain.c:1426-1428 */
        counts[topleaf[i]]++;
    }
}

/* domain_assign_topleaves_balanced domain.c:610-755 for leaves already in key order: contiguous runs of leaves
 * (segments) of about the mean cost, handed to the tasks in order so that neighbours on the curve share a task;
 * when every task has its share and leaves remain, a further round deals out the rest.  task[nleaf] out.
 * Returns the number of segments made, or -1 where the reference would stop (fewer segments than tasks x
 * nseg_per_task, cost not fully assigned). */
int oracle_domain_assign_balanced(int ntask, int32_t nleaf, const int64_t *cost, int nseg_per_task, int32_t *task)
{
    const int nsegment = ntask * nseg_per_task;
    int64_t total = 0;
    for(int32_t i = 0; i < nleaf; i++) { total += cost[i]; task[i] = -1; }
    int64_t left = total;
    double mean_expected = 1.0 * total / nsegment, mean_task = 1.0 * total / ntask;
    int curleaf = 0, curseg = 0, curtask = 0, nrounds = 0;
    int64_t curload = 0, curtaskload = 0;
    while(nrounds < nleaf) {
        int append = 0, advance = 0;
        if(curleaf == nleaf) advance = 1;
        else if(nleaf - curleaf == nsegment - curseg) { append = 1; advance = 1; }        /* one leaf per remaining segment */
        else {
            const int64_t assigned = (total - left) + curload;
            if(mean_expected * (curseg + 1) - assigned > 0.5 * cost[curleaf] || curload == 0) append = 1;
            else advance = 1;
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
    if(curseg < nsegment || left != 0) return -1;
    return curseg;
}

/* ---- the top tree: domain.c:826-1395 ---- */
static int32_t tt_find(const oracle_topnode *t, uint64_t key)              /* domain_toptree_get_subnode :826-835 */
{
    int32_t no = 0;
    while(t[no].Daughter >= 0) no = t[no].Daughter + (int32_t) ((key - t[no].StartKey) >> (t[no].Shift - 3));
    return no;
}
static int tt_split(oracle_topnode *t, int32_t *size, int32_t maxnodes, int32_t i)      /* domain_toptree_split :849-883 */
{
    if(*size + 8 > maxnodes) return 1;
    if(t[i].Shift < 3) return -1;                                           /* the reference stops: particles overly clustered */
    t[i].Daughter = *size;
    *size += 8;
    for(int j = 0; j < 8; j++) {
        oracle_topnode *s = &t[t[i].Daughter + j];
        s->Daughter = -1; s->Parent = i; s->Shift = t[i].Shift - 3; s->pad_ = 0;
        s->StartKey = t[i].StartKey + (uint64_t) j * (((uint64_t) 1) << s->Shift);
        s->Count = 0; s->Cost = 0;
    }
    return 0;
}
static void tt_sum(oracle_topnode *t, int32_t no)                           /* domain_toptree_update_cost :885-897 */
{
    if(t[no].Daughter == -1) return;
    for(int j = 0; j < 8; j++) {
        const int32_t sub = t[no].Daughter + j;
        tt_sum(t, sub);
        t[no].Count += t[sub].Count; t[no].Cost += t[sub].Cost;
    }
}
/* domain_check_for_local_refine_subsample :1084-1187 from the sorted subsample keys on: the skeleton in which no two
 * samples share a leaf (scanning sorted keys, either the leaf of the previous sample is refined or a fresh leaf is
 * entered), then counts and costs.  cost == NULL: 1 per sample.  Returns 0, 1 out of nodes, -1 where the reference stops. */
int oracle_toptree_local(const uint64_t *keys, const int64_t *cost, int64_t nsample, oracle_topnode *t, int32_t *size, int32_t maxnodes)
{
    *size = 1;
    memset(&t[0], 0, sizeof(t[0]));
    t[0].Daughter = -1; t[0].Parent = -1; t[0].Shift = 21 * 3;
    uint64_t last_key = UINT64_MAX;
    int32_t last_leaf = -1;
    int64_t i = 0;
    while(i < nsample) {
        const int32_t leaf = tt_find(t, keys[i]);
        if(leaf == last_leaf && t[leaf].Shift >= 3) {
            const int rc = tt_split(t, size, maxnodes, leaf);
            if(rc) return rc;
            t[leaf].Count = 0;
            last_leaf = tt_find(t, last_key);
            t[last_leaf].Count++;
            continue;
        }
        if(t[leaf].Count != 0 && leaf != last_leaf) return -1;
        last_key = keys[i];
        last_leaf = leaf;
        t[leaf].Count++;
        i++;
    }
    for(int32_t k = 0; k < *size; k++) t[k].Count = 0;
    for(i = 0; i < nsample; i++) {
        const int32_t leaf = tt_find(t, keys[i]);
        t[leaf].Count++; t[leaf].Cost += cost ? cost[i] : 1;
    }
    tt_sum(t, 0);
    return 0;
}
static void tt_cut(oracle_topnode *t, int32_t no, int64_t countlimit, int64_t costlimit)   /* domain_toptree_truncate_r :899-916 */
{
    if(t[no].Daughter == -1) return;
    if(t[no].Count < countlimit && t[no].Cost < costlimit) { t[no].Daughter = -1; return; }
    for(int j = 0; j < 8; j++) tt_cut(t, t[no].Daughter + j, countlimit, costlimit);
}
static void tt_compact(oracle_topnode *t, int32_t no, int32_t *next)        /* domain_toptree_garbage_collection :927-951 */
{
    if(t[no].Daughter == -1) return;
    const int32_t from = t[no].Daughter, to = *next;
    t[no].Daughter = to;
    *next += 8;
    for(int j = 0; j < 8; j++) { t[to + j] = t[from + j]; t[to + j].Parent = no; }
    for(int j = 0; j < 8; j++) tt_compact(t, to + j, next);
}
void oracle_toptree_truncate(oracle_topnode *t, int32_t *size, int64_t countlimit, int64_t costlimit)      /* :953-966 */
{
    tt_cut(t, 0, countlimit, costlimit);
    *size = 1;
    tt_compact(t, 0, size);
}
/* domain_toptree_merge :1473-1577: B's counts and refinement into A */
static int tt_merge(oracle_topnode *A, const oracle_topnode *B, int32_t a, int32_t b, int32_t *sizeA, int32_t maxnodes)
{
    if(B[b].Shift < A[a].Shift) {
        if(A[a].Daughter < 0) {
            if(*sizeA + 8 >= maxnodes) return 1;
            const int64_t count = A[a].Count - B[B[b].Parent].Count, cost = A[a].Cost - B[B[b].Parent].Cost;
            A[a].Daughter = *sizeA;
            for(int j = 0; j < 8; j++) {
                oracle_topnode *s = &A[A[a].Daughter + j];
                s->Shift = A[a].Shift - 3; s->pad_ = 0;
                s->Count = (j + 1) * count / 8 - j * count / 8;
                s->Cost = (j + 1) * cost / 8 - j * cost / 8;
                s->Daughter = -1; s->Parent = a;
                s->StartKey = A[a].StartKey + (uint64_t) j * (((uint64_t) 1) << s->Shift);
            }
            *sizeA += 8;
        }
        const int32_t sub = A[a].Daughter + (int32_t) ((B[b].StartKey - A[a].StartKey) >> (A[a].Shift - 3));
        return tt_merge(A, B, sub, b, sizeA, maxnodes);
    }
    if(B[b].Shift == A[a].Shift) {
        A[a].Count += B[b].Count; A[a].Cost += B[b].Cost;
        if(B[b].Daughter >= 0) {
            for(int j = 0; j < 8; j++) if(tt_merge(A, B, a, B[b].Daughter + j, sizeA, maxnodes)) return 1;
        } else if(A[a].Daughter >= 0) {
            for(int j = 0; j < 8; j++) if(tt_merge(A, B, A[a].Daughter + j, b, sizeA, maxnodes)) return 1;
        }
        return 0;
    }
    /* B's node is the larger one: spread its counts evenly over the A cells it covers */
    uint64_t nn = ((uint64_t) 1) << (B[b].Shift - A[a].Shift);
    if(B[b].Shift - A[a].Shift > 60) nn = 0;
    if(nn > 0) {
        A[a].Count += (int64_t) ((uint64_t) B[b].Count / nn); A[a].Cost += (int64_t) ((uint64_t) B[b].Cost / nn);   /* unsigned, as peano_t n */
        if(A[a].Daughter >= 0)
            for(int j = 0; j < 8; j++) if(tt_merge(A, B, A[a].Daughter + j, b, sizeA, maxnodes)) return 1;
    }
    return 0;
}
int oracle_toptree_merge(oracle_topnode *A, int32_t *sizeA, const oracle_topnode *B, int32_t maxnodes) { return tt_merge(A, B, 0, 0, sizeA, maxnodes); }
/* domain_global_refine :1343-1393 */
int oracle_toptree_global_refine(oracle_topnode *t, int32_t *size, int32_t maxnodes, int64_t countlimit, int64_t costlimit)
{
    for(int32_t i = 0; i < *size; i++) {
        if(t[i].Daughter >= 0 || t[i].Shift <= 0) continue;
        if(t[i].Count < countlimit && t[i].Cost < costlimit) continue;
        if(*size + 8 > maxnodes) return 1;
        t[i].Daughter = *size;
        for(int j = 0; j < 8; j++) {
            oracle_topnode *s = &t[t[i].Daughter + j];
            s->Shift = t[i].Shift - 3; s->pad_ = 0; s->Count = t[i].Count / 8; s->Cost = t[i].Cost / 8; s->Daughter = -1; s->Parent = i;
            s->StartKey = t[i].StartKey + (uint64_t) j * (((uint64_t) 1) << s->Shift);
        }
        *size += 8;
    }
    return 0;
}
/* domain_create_topleaves :810-824: leaves numbered along the curve */
static void tt_leaves(const oracle_topnode *t, int32_t no, int32_t *next, int32_t *leaf)
{
    if(t[no].Daughter == -1) { leaf[no] = (*next)++; return; }
    for(int j = 0; j < 8; j++) tt_leaves(t, t[no].Daughter + j, next, leaf);
}
int32_t oracle_toptree_leaves(const oracle_topnode *t, int32_t size, int32_t *leaf)
{
    for(int32_t i = 0; i < size; i++) leaf[i] = -1;
    int32_t next = 0;
    tt_leaves(t, 0, &next, leaf);
    return next;
}

/* domain_build_exchange_list + domain_build_plan, exchange.c:408-444,505-530 with domain_layoutfunc (domain.c:794-803):
 * the particles that leave this task (ascending index, garbage skipped and counted) and, per target task, how many go
 * there in total and per particle type: togo[ntask][7] = {base, slots[0..5]}.  Returns the list length, -1 on a leaf or
 * task out of range (where the reference stops). */
int64_t oracle_exchange_plan(int64_t n, const uint8_t *type, const uint8_t *flags, const int32_t *topleaf, int32_t nleaf,
                             const int32_t *task_of_leaf, int32_t ntask, int32_t thistask, int32_t *list_out, int64_t *togo, int64_t *ngarbage)
{
    memset(togo, 0, sizeof(int64_t) * 7 * ntask);
    int64_t nex = 0, ng = 0;
    for(int64_t i = 0; i < n; i++) {
        if(flags && (flags[i] & 1)) { ng++; continue; }
        if(topleaf[i] < 0 || topleaf[i] >= nleaf) return -1;
        const int32_t target = task_of_leaf[topleaf[i]];
        if(target < 0 || target >= ntask) return -1;
        if(target == thistask) continue;
        list_out[nex++] = (int32_t) i;
        togo[7 * target]++;
        togo[7 * target + 1 + type[i]]++;
    }
    *ngarbage = ng;
    return nex;
}
