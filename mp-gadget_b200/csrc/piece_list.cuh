// piece_list.cuh -- per-target lists of leaf pieces in a global chunk pool, shared by the
// tree-gravity kernels (tree_walk.cu) and the SPH kernels (sph.cu).
//
// A walk kernel (one warp = 32 targets adjacent on the curve) appends, for every lane that
// opened a particle leaf, the piece {first particle, count <= 8} to that lane's own list.  A
// chunk holds CH_SLOTS list slots x 32 lanes x 4 bytes laid out [slot][lane], so lanes with
// equal list lengths write one 128-byte line; a warp owns up to `maxch` chunks handed out
// by an atomic counter, their ids kept in a (dynamic) shared-memory table.  The pair kernels (same warp/target assignment) read the lists back
// with the lanes spread over SOURCE particles: group g = lane / 8 takes the pieces
// base + g + {0, 4, 8, 12} of each 16-piece step, slot = lane % 8 one particle of the piece.
#pragma once
#include "engine.h"

namespace b200 {

#define WALK_WARPS 4
#define CH_SHIFT 4
#define CH_SLOTS (1 << CH_SHIFT)
#define CH_WORDS (CH_SLOTS * 32)
// A warp owns up to PiecePool::maxch chunks (CH_SLOTS * maxch pieces = 8x that many source
// particles per target); the host starts every call with WALK_MAXCH0 and repeats a walk with
// a larger table when a list outgrows it (dense clumps inside an SPH search radius).
#define WALK_MAXCH0 128
#define WALK_MAXCH_LIMIT 8192
// entry = first particle << 4 | count (0..8)
#define PIECE(pstart, cnt) (((unsigned) (pstart) << 4) | (unsigned) (cnt))

// Per-warp stack of the gravity walk: opened internal nodes (node, mask).  A batch pops >= 1 entries and pushes <= 32;
// above WALK_STACK - WALK_RESERVE the batches shrink to one node's children (net growth <= 7 per tree level, depth <= 22).
#ifndef WALK_STACK
#define WALK_STACK 320
#endif
#define WALK_RESERVE 192

struct PiecePool {
    unsigned *__restrict__ pool;    // chunk pool
    int cap;                        // capacity in chunks
    int *__restrict__ ctl;          // [0] next free chunk, [1] error bits, [2..3] pieces written (u64)
    int *__restrict__ chunk_tab;    // [warp][maxch] chunk ids
    int maxch;                      // chunk-table entries per warp
    int *__restrict__ piece_cnt;    // [target slot] pieces in the list
};

// Append `entry` to the list of every lane with want == true.  Must be called by all 32 lanes.
// mycnt: this lane's list length; last: its last entry; nch_alloc: chunks the warp owns
// (warp-uniform); s_ctab: the warp's chunk ids in shared memory.
// A piece that continues the lane's last piece in particle order (the next leaf on the curve)
// is merged into it while the sum stays <= 8: trees whose leaves hold 2-4 particles (anything
// but a power-of-two lattice) would otherwise leave most of the 8 source slots of a piece idle.
// MERGE is a compile-time switch: the host turns it off for trees whose leaves are mostly full.
// The lane's newest entry stays in the register `last` while later pieces can still be merged into it and is
// stored when the next entry starts (or by piece_finish): one store per entry, none per merge.
template <bool MERGE>
__device__ __forceinline__ void piece_push(bool want, unsigned entry, int &mycnt, unsigned &last, int &nch_alloc, int *s_ctab,
                                           const PiecePool &Q, int group, int lane)
{
    if(MERGE) {
        const bool merge = want && mycnt > 0 && (entry >> 4) == (last >> 4) + (last & 15u) && (last & 15u) + (entry & 15u) <= 8u;
        if(merge) { last += entry & 15u; want = false; }
    }
    if(__any_sync(0xffffffffu, want && (mycnt >> CH_SHIFT) >= nch_alloc)) {   // once per CH_SLOTS pieces of the longest list
        const int needch = (int) __reduce_max_sync(0xffffffffu, want ? (unsigned) (mycnt >> CH_SHIFT) : 0u);
        if(needch >= Q.maxch) { if(lane == 0) atomicOr(Q.ctl + 1, 2); }
        else {
            if(lane == 0)
                for(int ch = nch_alloc; ch <= needch; ch++) {
                    const int id = atomicAdd(Q.ctl, 1);
                    s_ctab[ch] = id;
                    Q.chunk_tab[(size_t) group * Q.maxch + ch] = id;
                }
            nch_alloc = needch + 1;
        }
        __syncwarp();
    }
    if(want) {
        if(MERGE) {
            if(mycnt > 0) {                                     // the previous entry is final now
                const int at = mycnt - 1, ch = at >> CH_SHIFT;
                if(ch < nch_alloc) {
                    const int id = s_ctab[ch];
                    if(id < Q.cap) Q.pool[(size_t) id * CH_WORDS + (at & (CH_SLOTS - 1)) * 32 + lane] = last;
                }
            }
        } else {
            const int ch = mycnt >> CH_SHIFT;
            if(ch < nch_alloc) {
                const int id = s_ctab[ch];
                if(id < Q.cap) Q.pool[(size_t) id * CH_WORDS + (mycnt & (CH_SLOTS - 1)) * 32 + lane] = entry;
            }
        }
        mycnt++;
        last = entry;
    }
}

// The same in two steps, for a batch of appends per lane: piece_reserve (all 32 lanes; nadd = entries this lane is
// about to append) makes sure the warp owns the chunks of the longest resulting list, after which piece_append
// (any subset of lanes, no votes) stores.  With MERGE a list may end shorter than reserved; the chunk stays the warp's.
__device__ __forceinline__ void piece_reserve(int nadd, int mycnt, int &nch_alloc, int *s_ctab, const PiecePool &Q, int group, int lane)
{
    const int needch = (int) __reduce_max_sync(0xffffffffu, nadd > 0 ? (unsigned) ((mycnt + nadd - 1) >> CH_SHIFT) + 1u : 0u);
    if(needch > nch_alloc) {                    // warp-uniform
        if(needch > Q.maxch) { if(lane == 0) atomicOr(Q.ctl + 1, 2); }
        else {
            if(lane == 0)
                for(int ch = nch_alloc; ch < needch; ch++) {
                    const int id = atomicAdd(Q.ctl, 1);
                    s_ctab[ch] = id;
                    Q.chunk_tab[(size_t) group * Q.maxch + ch] = id;
                }
            nch_alloc = needch;
        }
        __syncwarp();
    }
}

template <bool MERGE>
__device__ __forceinline__ void piece_append(unsigned entry, int &mycnt, unsigned &last, int nch_alloc, const int *s_ctab,
                                             const PiecePool &Q, int lane)
{
    if(MERGE) {
        if(mycnt > 0 && (entry >> 4) == (last >> 4) + (last & 15u) && (last & 15u) + (entry & 15u) <= 8u) { last += entry & 15u; return; }
        if(mycnt > 0) {                                     // the previous entry is final now
            const int at = mycnt - 1, ch = at >> CH_SHIFT;
            if(ch < nch_alloc) {
                const int id = s_ctab[ch];
                if(id < Q.cap) Q.pool[(size_t) id * CH_WORDS + (at & (CH_SLOTS - 1)) * 32 + lane] = last;
            }
        }
    } else {
        const int ch = mycnt >> CH_SHIFT;
        if(ch < nch_alloc) {
            const int id = s_ctab[ch];
            if(id < Q.cap) Q.pool[(size_t) id * CH_WORDS + (mycnt & (CH_SLOTS - 1)) * 32 + lane] = entry;
        }
    }
    mycnt++;
    last = entry;
}

// End of a walk: the pending entry (MERGE), list lengths in target-slot order + statistics.
template <bool MERGE = true>
__device__ __forceinline__ void piece_finish(bool valid, int tslot, int mycnt, unsigned last, int nch_alloc, const int *s_ctab,
                                             const PiecePool &Q, int lane)
{
    if(MERGE && mycnt > 0) {
        const int at = mycnt - 1, ch = at >> CH_SHIFT;
        if(ch < nch_alloc) {
            const int id = s_ctab[ch];
            if(id < Q.cap) Q.pool[(size_t) id * CH_WORDS + (at & (CH_SLOTS - 1)) * 32 + lane] = last;
        }
    }
    const unsigned wsum = __reduce_add_sync(0xffffffffu, (unsigned) mycnt);
    if(lane == 0) atomicAdd((unsigned long long *) (Q.ctl + 2), (unsigned long long) wsum);
    if(valid) Q.piece_cnt[tslot] = mycnt;
}

// Reader side.
struct PieceList {
    const unsigned *__restrict__ pool;     // chunk pool
    const int *__restrict__ ctab;          // this warp's chunk ids (shared memory)
    unsigned empty;                        // PIECE(sentinel, 0)
    int t;                                 // column = lane of the target in its warp
    int nt;                                // pieces in the list
};
struct Ent4 { unsigned e[4]; };

__device__ __forceinline__ Ent4 fetch_ent(const PieceList &L, int base, int g)
{
    Ent4 E;
#pragma unroll
    for(int k = 0; k < 4; k++) E.e[k] = L.empty;
    if(base < L.nt) {                       // warp-uniform; base is a multiple of CH_SLOTS = 16
        const unsigned *row = L.pool + (size_t) L.ctab[base >> CH_SHIFT] * CH_WORDS + g * 32 + L.t;
#pragma unroll
        for(int k = 0; k < 4; k++) if(base + g + 4 * k < L.nt) E.e[k] = row[k * 128];
    }
    return E;
}

// The warp's chunk ids into shared memory (reader side); returns nothing, syncs the warp.
__device__ __forceinline__ void piece_load_ctab(int *s_ctab, const int *__restrict__ chunk_tab, int maxch, int group, int mycnt, int lane)
{
    const int maxcnt = (int) __reduce_max_sync(0xffffffffu, (unsigned) mycnt);
    const int nch = (maxcnt + CH_SLOTS - 1) >> CH_SHIFT;
    for(int c = lane; c < nch; c += 32) s_ctab[c] = chunk_tab[(size_t) group * maxch + c];
    __syncwarp();
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for(int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_min(double v)
{
#pragma unroll
    for(int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_max(double v)
{
#pragma unroll
    for(int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Host side (tree_walk.cu): storage for `nwarps` target warps, and the check after a walk.
// piece_pool_check returns 0 and *retry = false when the lists are complete, *retry = true when
// the pool was too small (it has been grown; run the walk again).
// piece_pool_reset starts a top-level call (chunk table back to WALK_MAXCH0 entries);
// piece_ctab_bytes is the dynamic shared memory a kernel of `warps` warps needs for its tables.
void piece_pool_reset(Engine *E);
int piece_pool_begin(Engine *E, int64_t nwarps, PiecePool *Q);
int piece_pool_check(Engine *E, int64_t nwarps, bool *retry, int attempt);
inline size_t piece_ctab_bytes(const Engine *E, int warps) { return (size_t) warps * E->walk_maxch * sizeof(int); }
template <class K> inline cudaError_t piece_set_smem(K kernel, size_t bytes)
{
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) bytes);
}

} // namespace b200
