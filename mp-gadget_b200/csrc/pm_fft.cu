// pm_fft.cu -- the PM step's transforms as five shared-memory passes with the Green's function inside.
//
// Replaces, for one resident mesh, pfft_execute_dft_r2c -> potential_transfer -> pfft_execute_dft_c2r
// (petapm.c:305,326-344, gravpm.c:383-454): the density mesh goes in, the potential mesh comes out.
// cuFFT needs 5 passes over the data per 3-D transform plus the transfer kernel in between (11 passes,
// 20 ms at 768^3); here every pass moves each value through HBM exactly once in each direction:
//
//   z forward   real lines [x][y][0..N)  -> half spectrum [x][y][0..N/2]      (length-N/2 complex FFT of the
//                                                                               packed line + split)
//   y forward   columns over y, in place
//   x forward + potential_transfer (+ powerspectrum_add_mode) + x inverse, in place, one kernel
//   y inverse   in place
//   z inverse   half spectrum -> real lines
//
// A tile is L rows x 8 complex values (128 contiguous bytes of the half spectrum) in shared memory; 8
// neighbouring threads own the 8 columns of a row, so every shared-memory access of a quarter warp is
// one conflict-free 128-byte wavefront whatever the row, and every global access is a full 128-byte
// line.  The transform is an in-place decimation-in-frequency chain (radix 4/2/3/5, two radix-4 levels
// or a radix-4 and a radix-2 level fused in registers = radix 16 / 8); its digit-reversed output order is
// never undone in shared memory: stores and the Green's function look rows up through a small table,
// and the inverse is the exact mirror chain (decimation in time), which takes digit-reversed rows and
// returns natural order.  Conventions are cuFFT's / PFFT's: unnormalised in both directions.
#include "engine.h"
#include <math.h>
#include <stdlib.h>

#ifndef B200_DYN_SMEM
#define B200_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

namespace b200 {

#ifndef FFT_T
#define FFT_T 8        // complex values per tile row: 128 bytes (4 = 64-byte rows, experiment builds)
#endif
#define FFT_MAXST 12
#ifndef FFT_MINB
#define FFT_MINB 2      // resident blocks per SM the kernels are compiled for (register cap 128 at 256 threads)
#endif
enum { ST_44 = 44, ST_42 = 42 };

struct FftPlan {
    int L;                          // line length
    int nst;                        // stage groups
    int kind[FFT_MAXST];            // ST_44, ST_42 or the radix of a single stage (2, 3, 4, 5)
    int m[FFT_MAXST];               // sub-transform length the group starts from
    const double2 *tw;              // exp(-2 pi i k / L), k in [0, L)
    const unsigned short *pos;      // row that holds frequency k after the forward chain
    const unsigned short *freq;     // frequency held by row l
};

__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ double2 cmulc(double2 a, double2 b) { return make_double2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }   // a conj(b)

// u_q = sum_p v_p exp(-+ 2 pi i p q / R), in place
template <bool INV> __device__ __forceinline__ void bf2(double2 &a, double2 &b)
{
    const double2 s = cadd(a, b), d = csub(a, b);
    a = s; b = d;
}
template <bool INV> __device__ __forceinline__ void bf4(double2 &v0, double2 &v1, double2 &v2, double2 &v3)
{
    const double2 t0 = cadd(v0, v2), t1 = csub(v0, v2), t2 = cadd(v1, v3), t3 = csub(v1, v3);
    const double2 jt = INV ? make_double2(-t3.y, t3.x) : make_double2(t3.y, -t3.x);      // -+ i t3
    v0 = cadd(t0, t2); v2 = csub(t0, t2);
    v1 = cadd(t1, jt); v3 = csub(t1, jt);
}
template <bool INV> __device__ __forceinline__ void bf3(double2 &v0, double2 &v1, double2 &v2)
{
    const double s3 = 0.86602540378443864676;
    const double2 t1 = cadd(v1, v2);
    const double2 t2 = make_double2(v0.x - 0.5 * t1.x, v0.y - 0.5 * t1.y);
    const double2 d = csub(v1, v2);
    const double2 t3 = make_double2(s3 * d.x, s3 * d.y);
    const double2 jt = INV ? make_double2(-t3.y, t3.x) : make_double2(t3.y, -t3.x);
    v0 = cadd(v0, t1);
    v1 = cadd(t2, jt); v2 = csub(t2, jt);
}
template <bool INV> __device__ __forceinline__ void bf5(double2 &v0, double2 &v1, double2 &v2, double2 &v3, double2 &v4)
{
    const double c1 = 0.30901699437494742410, c2 = -0.80901699437494742410;
    const double s1 = 0.95105651629515357212, s2 = 0.58778525229247312917;
    const double2 a1 = cadd(v1, v4), a2 = cadd(v2, v3), b1 = csub(v1, v4), b2 = csub(v2, v3);
    const double2 r1 = make_double2(v0.x + c1 * a1.x + c2 * a2.x, v0.y + c1 * a1.y + c2 * a2.y);
    const double2 r2 = make_double2(v0.x + c2 * a1.x + c1 * a2.x, v0.y + c2 * a1.y + c1 * a2.y);
    const double2 i1 = make_double2(s1 * b1.x + s2 * b2.x, s1 * b1.y + s2 * b2.y);
    const double2 i2 = make_double2(s2 * b1.x - s1 * b2.x, s2 * b1.y - s1 * b2.y);
    const double2 j1 = INV ? make_double2(-i1.y, i1.x) : make_double2(i1.y, -i1.x);
    const double2 j2 = INV ? make_double2(-i2.y, i2.x) : make_double2(i2.y, -i2.x);
    v0 = cadd(v0, cadd(a1, a2));
    v1 = cadd(r1, j1); v4 = csub(r1, j1);
    v2 = cadd(r2, j2); v3 = csub(r2, j2);
}

// exp(-2 pi i k / 16)
__device__ __forceinline__ double2 c16(int k)
{
    switch(k) {
        case 1: return make_double2(0.92387953251128675613, -0.38268343236508977173);
        case 2: return make_double2(0.70710678118654752440, -0.70710678118654752440);
        case 3: return make_double2(0.38268343236508977173, -0.92387953251128675613);
        case 4: return make_double2(0.0, -1.0);
        case 6: return make_double2(-0.70710678118654752440, -0.70710678118654752440);
        case 9: return make_double2(-0.92387953251128675613, 0.38268343236508977173);
        default: return make_double2(1.0, 0.0);
    }
}

// ---- one stage group on the values of one butterfly, in registers ----
// v[p] is the value of row base + p s (s = m / R).  Forward: butterfly, then the twiddle exp(-2 pi i j q / m);
// inverse: the conjugate twiddle, then the conjugate butterfly (the adjoint of the forward group).  jt = j L / m
// indexes the table tw[k] = exp(-2 pi i k / L).
template <int KIND> struct GroupSize { static const int R = KIND == ST_44 ? 16 : (KIND == ST_42 ? 8 : KIND); };

template <int KIND, bool INV>
__device__ __forceinline__ void group_compute(double2 *v, const double2 *tw, int jt)
{
    if(KIND == ST_44) {
        // two radix-4 stages (sub-lengths m and m/4): the twiddle of the first, exp(-2 pi i (j + a m/16) q / m), is the
        // loaded exp(-2 pi i j q / m) times the constant 16th root a q
        if(!INV) {
            const double2 w1 = tw[jt], w2 = tw[2 * jt], w3 = tw[3 * jt];
#pragma unroll
            for(int a = 0; a < 4; a++) {
                bf4<false>(v[a], v[a + 4], v[a + 8], v[a + 12]);
                double2 u1 = cmul(v[a + 4], w1), u2 = cmul(v[a + 8], w2), u3 = cmul(v[a + 12], w3);
                if(a > 0) { u1 = cmul(u1, c16(a)); u2 = cmul(u2, c16(2 * a)); u3 = cmul(u3, c16(3 * a)); }
                v[a + 4] = u1; v[a + 8] = u2; v[a + 12] = u3;
            }
            const double2 x1 = tw[4 * jt], x2 = tw[8 * jt], x3 = tw[12 * jt];
#pragma unroll
            for(int q = 0; q < 4; q++) {
                bf4<false>(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                v[4 * q + 1] = cmul(v[4 * q + 1], x1); v[4 * q + 2] = cmul(v[4 * q + 2], x2); v[4 * q + 3] = cmul(v[4 * q + 3], x3);
            }
        } else {
            const double2 x1 = tw[4 * jt], x2 = tw[8 * jt], x3 = tw[12 * jt];
#pragma unroll
            for(int q = 0; q < 4; q++) {
                v[4 * q + 1] = cmulc(v[4 * q + 1], x1); v[4 * q + 2] = cmulc(v[4 * q + 2], x2); v[4 * q + 3] = cmulc(v[4 * q + 3], x3);
                bf4<true>(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            }
            const double2 w1 = tw[jt], w2 = tw[2 * jt], w3 = tw[3 * jt];
#pragma unroll
            for(int a = 0; a < 4; a++) {
                double2 u1 = cmulc(v[a + 4], w1), u2 = cmulc(v[a + 8], w2), u3 = cmulc(v[a + 12], w3);
                if(a > 0) { u1 = cmulc(u1, c16(a)); u2 = cmulc(u2, c16(2 * a)); u3 = cmulc(u3, c16(3 * a)); }
                v[a + 4] = u1; v[a + 8] = u2; v[a + 12] = u3;
                bf4<true>(v[a], v[a + 4], v[a + 8], v[a + 12]);
            }
        }
    } else if(KIND == ST_42) {
        // a radix-4 stage (sub-length m) and a radix-2 stage (sub-length m/4)
        const double2 w1 = tw[jt], w2 = tw[2 * jt], w3 = tw[3 * jt], x1 = tw[4 * jt];
        if(!INV) {
#pragma unroll
            for(int a = 0; a < 2; a++) {
                bf4<false>(v[a], v[a + 2], v[a + 4], v[a + 6]);
                double2 u1 = cmul(v[a + 2], w1), u2 = cmul(v[a + 4], w2), u3 = cmul(v[a + 6], w3);
                if(a > 0) { u1 = cmul(u1, c16(2)); u2 = cmul(u2, c16(4)); u3 = cmul(u3, c16(6)); }     // 8th roots
                v[a + 2] = u1; v[a + 4] = u2; v[a + 6] = u3;
            }
#pragma unroll
            for(int q = 0; q < 4; q++) {
                bf2<false>(v[2 * q], v[2 * q + 1]);
                v[2 * q + 1] = cmul(v[2 * q + 1], x1);
            }
        } else {
#pragma unroll
            for(int q = 0; q < 4; q++) {
                v[2 * q + 1] = cmulc(v[2 * q + 1], x1);
                bf2<true>(v[2 * q], v[2 * q + 1]);
            }
#pragma unroll
            for(int a = 0; a < 2; a++) {
                double2 u1 = cmulc(v[a + 2], w1), u2 = cmulc(v[a + 4], w2), u3 = cmulc(v[a + 6], w3);
                if(a > 0) { u1 = cmulc(u1, c16(2)); u2 = cmulc(u2, c16(4)); u3 = cmulc(u3, c16(6)); }
                v[a + 2] = u1; v[a + 4] = u2; v[a + 6] = u3;
                bf4<true>(v[a], v[a + 2], v[a + 4], v[a + 6]);
            }
        }
    } else {
        const int R = GroupSize<KIND>::R;
        if(INV) {
#pragma unroll
            for(int q = 1; q < R; q++) v[q] = cmulc(v[q], tw[jt * q]);
        }
        if(KIND == 2) bf2<INV>(v[0], v[1]);
        if(KIND == 3) bf3<INV>(v[0], v[1], v[2]);
        if(KIND == 4) bf4<INV>(v[0], v[1], v[2], v[3]);
        if(KIND == 5) bf5<INV>(v[0], v[1], v[2], v[3], v[4]);
        if(!INV) {
#pragma unroll
            for(int q = 1; q < R; q++) v[q] = cmul(v[q], tw[jt * q]);
        }
    }
}

// ---- where a group's rows come from and go to ----
// A group reads its rows either from the shared-memory tile or straight from global memory (the first group of a chain:
// 16 independent 128-byte lines in flight per quarter warp, and the tile is neither written nor read for it), and writes
// them to the tile or straight to global memory (the last group).  `map` translates a chain row into the global row
// (frequency order <-> the digit-reversed row order of the chain); null = identity.
// Tile element (row, column t) lives in slot (t + row) & 7 of its 128-byte row: 8 threads on the 8 columns of one row
// and 8 threads on 8 consecutive rows of one column (the z passes' global-memory side) are both conflict-free.
struct RowsTile {
    double2 *tile;
    int t;
    __device__ __forceinline__ double2 ld(int row) const { return tile[(size_t) row * FFT_T + ((t + row) & (FFT_T - 1))]; }
    __device__ __forceinline__ void st(int row, double2 v) const { tile[(size_t) row * FFT_T + ((t + row) & (FFT_T - 1))] = v; }
};
// The same through a row map (the z passes keep a second tile in frequency order for the split / merge of the packed line).
struct RowsTileMap {
    double2 *tile;
    int t;
    const unsigned short *map;
    __device__ __forceinline__ double2 ld(int row) const { const int r = map[row]; return tile[(size_t) r * FFT_T + ((t + r) & (FFT_T - 1))]; }
    __device__ __forceinline__ void st(int row, double2 v) const { const int r = map[row]; tile[(size_t) r * FFT_T + ((t + r) & (FFT_T - 1))] = v; }
};
// CONJ: values are conjugated on the way in and out, which turns the forward chain into the inverse transform.
template <bool CONJ>
struct RowsGlobal {
    double2 *p;             // first row + column
    size_t stride;
    const unsigned short *map;
    bool ok;                // stores enabled (ragged last tile of the z passes)
    __device__ __forceinline__ double2 ld(int row) const
    {
        const double2 v = p[(size_t) (map ? map[row] : row) * stride];
        return CONJ ? make_double2(v.x, -v.y) : v;
    }
    __device__ __forceinline__ void st(int row, double2 v) const
    {
        if(ok) p[(size_t) (map ? map[row] : row) * stride] = CONJ ? make_double2(v.x, -v.y) : v;
    }
};
struct NoMid { __device__ __forceinline__ double2 operator()(int, double2 v) const { return v; } };

// DIR 0: forward group; 1: inverse group; 2: forward group, mid(row, value) on every row, inverse group (the turn of the
// x pass: the last forward group, the Green's function and the first inverse group on the same registers).
template <int KIND, int DIR, class Src, class Dst, class Mid>
__device__ __forceinline__ void run_group(const Src &src, const Dst &dst, const Mid &mid, const double2 *tw, int L, int m, int lane, int NL)
{
    const int R = GroupSize<KIND>::R;
    const int U = R <= 3 ? 4 : (R <= 5 ? 2 : 1);      // butterflies per turn: 8 to 16 independent rows in flight per thread
    const int s = m / R, nb = L / R, tws = L / m;
    for(int b0 = lane; b0 < nb; b0 += NL * U) {
        double2 v[U * R];
        int base[U], jt[U];
        bool live[U];
#pragma unroll
        for(int u = 0; u < U; u++) {
            int b = b0 + u * NL;
            live[u] = b < nb;
            if(!live[u]) b = b0;                 // a spare slot of the last turn repeats the first butterfly and stores nothing
            const int g = b / s, j = b - g * s;
            base[u] = g * m + j; jt[u] = j * tws;
#pragma unroll
            for(int p = 0; p < R; p++) v[u * R + p] = src.ld(base[u] + p * s);
        }
#pragma unroll
        for(int u = 0; u < U; u++) {
            if(DIR == 0 || DIR == 2) group_compute<KIND, false>(v + u * R, tw, jt[u]);
            if(DIR == 2) {
                if(live[u]) {
#pragma unroll
                    for(int p = 0; p < R; p++) v[u * R + p] = mid(base[u] + p * s, v[u * R + p]);
                }
            }
            if(DIR == 1 || DIR == 2) group_compute<KIND, true>(v + u * R, tw, jt[u]);
        }
#pragma unroll
        for(int u = 0; u < U; u++) {
            if(live[u]) {
#pragma unroll
                for(int p = 0; p < R; p++) dst.st(base[u] + p * s, v[u * R + p]);
            }
        }
    }
}

template <int DIR, class Src, class Dst, class Mid>
__device__ __forceinline__ void run_kind(int kind, const Src &src, const Dst &dst, const Mid &mid, const double2 *tw, int L, int m, int lane, int NL)
{
    switch(kind) {
        case ST_44: run_group<ST_44, DIR>(src, dst, mid, tw, L, m, lane, NL); break;
        case ST_42: run_group<ST_42, DIR>(src, dst, mid, tw, L, m, lane, NL); break;
        case 4: run_group<4, DIR>(src, dst, mid, tw, L, m, lane, NL); break;
        case 2: run_group<2, DIR>(src, dst, mid, tw, L, m, lane, NL); break;
        case 3: run_group<3, DIR>(src, dst, mid, tw, L, m, lane, NL); break;
        default: run_group<5, DIR>(src, dst, mid, tw, L, m, lane, NL); break;
    }
}

// Which column and which butterflies a thread works on.  Standard: 8 neighbouring threads = the 8 columns of one row.
// Transposed (the groups of the z passes that touch global memory, where a column is a contiguous line): 8 neighbouring
// threads = 8 consecutive rows of one column.
struct ThreadMap { int t, lane; };
__device__ __forceinline__ ThreadMap map_standard() { ThreadMap M = {(int) (threadIdx.x & (FFT_T - 1)), (int) (threadIdx.x / FFT_T)}; return M; }
__device__ __forceinline__ ThreadMap map_transposed()
{
    ThreadMap M = {(int) ((threadIdx.x / FFT_T) & (FFT_T - 1)), (int) ((threadIdx.x / (FFT_T * FFT_T)) * FFT_T + (threadIdx.x & (FFT_T - 1)))};
    return M;
}

// Groups [first, last] of the chain in forward (DIR 0: ascending) or inverse (DIR 1: descending) order; the first group
// executed reads `src` under thread map Ma, the last one writes `dst` under Mz, everything in between goes through the
// tile under the standard map.  A barrier follows every group that wrote the tile.
template <int DIR, class Src, class Dst>
__device__ __forceinline__ void run_chain(const FftPlan &P, int first, int last, const Src &src, const Dst &dst, double2 *tile,
                                          bool dst_is_tile, const double2 *tw, ThreadMap Ma, ThreadMap Mz)
{
    const int n = last - first + 1, NL = blockDim.x / FFT_T;
    const ThreadMap Ms = map_standard();
    const RowsTile Ta = {tile, Ma.t}, Tz = {tile, Mz.t}, Ts = {tile, Ms.t};
    for(int i = 0; i < n; i++) {
        const int g = DIR == 1 ? last - i : first + i;
        const bool a = i == 0, z = i == n - 1;
        if(a && z && !dst_is_tile) {
            // a single group: its global stores (through a row map) must not overtake the loads of other threads
            run_kind<DIR>(P.kind[g], src, Ta, NoMid(), tw, P.L, P.m[g], Ma.lane, NL);
            __syncthreads();
            for(int l = Ma.lane; l < P.L; l += NL) dst.st(l, Ta.ld(l));
        } else if(a && z) run_kind<DIR>(P.kind[g], src, dst, NoMid(), tw, P.L, P.m[g], Ma.lane, NL);
        else if(a) run_kind<DIR>(P.kind[g], src, Ta, NoMid(), tw, P.L, P.m[g], Ma.lane, NL);
        else if(z) run_kind<DIR>(P.kind[g], Tz, dst, NoMid(), tw, P.L, P.m[g], Mz.lane, NL);
        else run_kind<DIR>(P.kind[g], Ts, Ts, NoMid(), tw, P.L, P.m[g], Ms.lane, NL);
        if(!z || dst_is_tile) __syncthreads();
    }
}

// tables of the plan into shared memory behind the tile: tw[L] | pos[L] | freq[L]
struct SmemTables { double2 *tw; unsigned short *pos, *freq; };
__device__ __forceinline__ SmemTables load_tables(double2 *behind_tiles, const FftPlan &P)
{
    SmemTables S;
    S.tw = behind_tiles;
    S.pos = (unsigned short *) (S.tw + P.L);
    S.freq = S.pos + P.L;
    for(int i = threadIdx.x; i < P.L; i += blockDim.x) { S.tw[i] = P.tw[i]; S.pos[i] = P.pos[i]; S.freq[i] = P.freq[i]; }
    return S;
}

// z forward: 8 consecutive real lines of the mesh per block; line = N reals = L = N/2 packed complex values
// z[j] = x[2j] + i x[2j+1].  With Z = FFT_L(z):  X[k] = (Z[k] + conj Z[L-k])/2 - i exp(-2 pi i k/N) (Z[k] - conj Z[L-k])/2,
// k = 0..L (Z[L] = Z[0]).  Output rows have pitch Nzp >= L + 1 (a multiple of 8); the padding is zeroed.
__global__ void __launch_bounds__(256, FFT_MINB)
k_fft_z_forward(const double *__restrict__ mesh, double2 *__restrict__ out, long long nlines, int Nzp,
                FftPlan P, const double2 *__restrict__ wN)
{
    B200_DYN_SMEM(smem);
    double2 *tile = (double2 *) smem;
    const int L = P.L, NL = blockDim.x / FFT_T;
    const ThreadMap Mg = map_transposed(), Ms = map_standard();
    long long line = (long long) blockIdx.x * FFT_T + Mg.t;
    const bool ok = line < nlines;
    if(!ok) line = nlines - 1;             // ragged last tile: the spare columns recompute the last line and store nothing
    double2 *tile2 = tile + (size_t) L * FFT_T;          // the transform in frequency order
    const SmemTables S = load_tables(tile2 + (size_t) L * FFT_T, P);
    __syncthreads();
    const RowsGlobal<false> in = {(double2 *) mesh + line * L, 1, nullptr, false};
    if(P.nst == 1) { const RowsTileMap T1 = {tile2, Mg.t, S.freq}; run_chain<0>(P, 0, 0, in, T1, tile, true, S.tw, Mg, Mg); }
    else { const RowsTileMap T1 = {tile2, Ms.t, S.freq}; run_chain<0>(P, 0, P.nst - 1, in, T1, tile, true, S.tw, Mg, Ms); }
    if(!ok) return;
    const RowsTile T = {tile2, Mg.t};
    double2 *o = out + line * Nzp;
    for(int kk = Mg.lane; kk <= L / 2; kk += NL) {
        if(kk == 0) {
            const double2 Z0 = T.ld(0);
            o[0] = make_double2(Z0.x + Z0.y, 0.0);
            o[L] = make_double2(Z0.x - Z0.y, 0.0);
            continue;
        }
        const int k2 = L - kk;
        const double2 Z1 = T.ld(kk), Z2 = T.ld(k2);
        const double2 A = make_double2(Z1.x + Z2.x, Z1.y - Z2.y);          // Z1 + conj Z2
        const double2 B = make_double2(Z1.x - Z2.x, Z1.y + Z2.y);          // Z1 - conj Z2
        const double2 Q = cmul(wN[kk], B);
        o[kk] = make_double2(0.5 * (A.x + Q.y), 0.5 * (A.y - Q.x));
        if(k2 != kk) o[k2] = make_double2(0.5 * (A.x - Q.y), 0.5 * (-A.y - Q.x));
    }
    for(int k = L + 1 + Mg.lane; k < Nzp; k += NL) o[k] = make_double2(0.0, 0.0);
}

// z inverse: Z[k] = (X[k] + conj X[L-k]) + i exp(+2 pi i k/N) (X[k] - conj X[L-k]), inverse FFT_L, unpack: N x the real line.
__global__ void __launch_bounds__(256, FFT_MINB)
k_fft_z_inverse(const double2 *__restrict__ in, double *__restrict__ mesh, long long nlines, int Nzp,
                FftPlan P, const double2 *__restrict__ wN)
{
    B200_DYN_SMEM(smem);
    double2 *tile = (double2 *) smem;
    const int L = P.L, NL = blockDim.x / FFT_T;
    const ThreadMap Mg = map_transposed(), Ms = map_standard();
    long long line = (long long) blockIdx.x * FFT_T + Mg.t;
    const bool ok = line < nlines;
    if(!ok) line = nlines - 1;
    double2 *tile2 = tile + (size_t) L * FFT_T;          // the packed line's spectrum in frequency order
    const SmemTables S = load_tables(tile2 + (size_t) L * FFT_T, P);
    __syncthreads();
    const RowsTile T = {tile2, Mg.t};
    const double2 *x = in + line * Nzp;
    for(int kk = Mg.lane; kk <= L / 2; kk += NL) {
        const int k2 = L - kk;
        double2 X1 = x[kk], X2 = x[k2];
        if(kk == 0) { X1.y = 0.0; X2.y = 0.0; }          // the two real modes of a Hermitian line (cuFFT ignores their imaginary parts too)
        const double2 A = make_double2(X1.x + X2.x, X1.y - X2.y);
        const double2 B = make_double2(X1.x - X2.x, X1.y + X2.y);
        const double2 Q = cmulc(B, wN[kk]);
        T.st(kk, make_double2(A.x - Q.y, A.y + Q.x));
        if(kk != 0 && k2 != kk) T.st(k2, make_double2(A.x + Q.y, -A.y + Q.x));
    }
    __syncthreads();
    const RowsGlobal<false> o = {(double2 *) mesh + line * L, 1, nullptr, ok};
    if(P.nst == 1) { const RowsTileMap T1 = {tile2, Mg.t, S.freq}; run_chain<1>(P, 0, 0, T1, o, tile, false, S.tw, Mg, Mg); }
    else { const RowsTileMap T1 = {tile2, Ms.t, S.freq}; run_chain<1>(P, 0, P.nst - 1, T1, o, tile, false, S.tw, Ms, Mg); }
}

struct GreenArgs {
    int N, Nz;
    const double *ktab;
    double asmth2, pot_factor, binsperunit;
    double *ps;
};

// potential_transfer (gravpm.c:383-454) on the value of chain row `row` of column iz, line iy; POWER: also
// powerspectrum_add_mode (gravpm.c:330-361) on the untouched density mode, into the block's shared-memory bins.
template <bool POWER>
struct GreenMid {
    GreenArgs G;
    const unsigned short *freq;
    int iy, iz, ky;
    double fy, fz;
    double *s_ps;
    __device__ __forceinline__ double2 operator()(int row, double2 val) const
    {
        const int N = G.N;
        if(iz >= G.Nz) return val;                       // padding columns of the row pitch
        const int ix = freq[row];
        const int kx = ix <= N / 2 ? ix : ix - N;       // petapm_mesh_to_k petapm.c:81-84
        const long long k2 = (long long) kx * kx + (long long) ky * ky + (long long) iz * iz;
        if(k2 == 0) {
            if(POWER) G.ps[3 * N] = val.x * val.x + val.y * val.y;       // gravpm.c:332-336
            return make_double2(0.0, 0.0);                               // gravpm.c:441-449
        }
        const double smth = exp((double) (-k2) * G.asmth2) / (double) k2;
        const double f = (G.ktab[ix] * fy) * fz;
        if(POWER) {
            const int kint = (int) floor(G.binsperunit * log((double) k2) / 2.);
            if(kint < N) {
                const double w = (iz == 0 || iz == N / 2) ? 1.0 : 2.0;
                const double mm = val.x * val.x + val.y * val.y;
                atomicAdd(&s_ps[kint], w * mm * f * f);
                atomicAdd(&s_ps[N + kint], w * sqrt((double) k2));
                atomicAdd(&s_ps[2 * N + kint], w);
            }
        }
        const double fac = ((G.pot_factor * smth) * f) * f;
        return make_double2(val.x * fac, val.y * fac);
    }
};

// Column passes over the half spectrum, in place.  Block (outer, tk) owns rows base + l * stride, l in [0, L), of 8
// complex values each, base = outer * outer_stride + 8 tk.
//   MODE 0: forward transform, rows back in frequency order.            (y forward: outer = ix, stride = Nzp)
//   MODE 1: inverse transform of rows given in frequency order.         (y inverse)
//   MODE 2, 3: forward, potential_transfer (MODE 3 also the power-spectrum sums), inverse.
//                                                                       (x: outer = iy, stride = N Nzp)
template <int MODE>
__global__ void __launch_bounds__(256, FFT_MINB)
k_fft_columns(double2 *__restrict__ v, int ntile, size_t outer_stride, size_t stride, FftPlan P, GreenArgs G)
{
    B200_DYN_SMEM(smem);
    double2 *tile = (double2 *) smem;
    const int NL = blockDim.x / FFT_T;
    const ThreadMap M = map_standard();
    const int t = M.t, lane = M.lane;
    const int outer = blockIdx.x / ntile, tk = blockIdx.x - outer * ntile;
    double2 *g = v + (size_t) outer * outer_stride + (size_t) tk * FFT_T + t;
    const SmemTables S = load_tables(tile + (size_t) P.L * FFT_T, P);
    double *s_ps = (double *) (S.freq + P.L);          // MODE 3: [3][N]
    if(MODE == 3) for(int b = threadIdx.x; b < 3 * G.N; b += blockDim.x) s_ps[b] = 0;
    __syncthreads();
    const RowsTile T = {tile, t};
    const int last = P.nst - 1;
    if(MODE == 0) {
        const RowsGlobal<false> in = {g, stride, nullptr, true}, out = {g, stride, S.freq, true};
        run_chain<0>(P, 0, last, in, out, tile, false, S.tw, M, M);
    } else if(MODE == 1) {
        // the inverse as the conjugate of the forward chain on conjugated input: the 16-row group reads global memory
        const RowsGlobal<true> in = {g, stride, nullptr, true}, out = {g, stride, S.freq, true};
        run_chain<0>(P, 0, last, in, out, tile, false, S.tw, M, M);
    } else {
        const RowsGlobal<false> io = {g, stride, nullptr, true};
        const int iy = outer, iz = tk * FFT_T + t, N = G.N;
        const GreenMid<MODE == 3> mid = {G, S.freq, iy, iz, iy <= N / 2 ? iy : iy - N, G.ktab[iy], G.ktab[iz < G.Nz ? iz : 0], s_ps};
        const int kind = P.kind[last], m = P.m[last];
        if(last == 0) run_kind<2>(kind, io, io, mid, S.tw, P.L, m, lane, NL);
        else {
            run_chain<0>(P, 0, last - 1, io, T, tile, true, S.tw, M, M);
            run_kind<2>(kind, T, T, mid, S.tw, P.L, m, lane, NL);
            __syncthreads();
            run_chain<1>(P, 0, last - 1, T, io, tile, false, S.tw, M, M);
        }
        if(MODE == 3) {
            __syncthreads();
            for(int b = threadIdx.x; b < 3 * N; b += blockDim.x) if(s_ps[b] != 0) atomicAdd(&G.ps[b], s_ps[b]);
        }
    }
}

// ---- host side ----

static bool plan_stages(int L, std::vector<int> &kind, std::vector<int> &ms, std::vector<int> &radices)
{
    int n = L, m = L, p2 = 0;
    while(n % 2 == 0) { n /= 2; p2++; }
    int n3 = 0, n5 = 0;
    while(n % 3 == 0) { n /= 3; n3++; }
    while(n % 5 == 0) { n /= 5; n5++; }
    if(n != 1) return false;
    while(p2 >= 4) { kind.push_back(ST_44); ms.push_back(m); radices.push_back(4); radices.push_back(4); m /= 16; p2 -= 4; }
    if(p2 == 3) { kind.push_back(ST_42); ms.push_back(m); radices.push_back(4); radices.push_back(2); m /= 8; p2 = 0; }
    if(p2 == 2) { kind.push_back(4); ms.push_back(m); radices.push_back(4); m /= 4; p2 = 0; }
    if(p2 == 1) { kind.push_back(2); ms.push_back(m); radices.push_back(2); m /= 2; p2 = 0; }
    for(int i = 0; i < n3; i++) { kind.push_back(3); ms.push_back(m); radices.push_back(3); m /= 3; }
    for(int i = 0; i < n5; i++) { kind.push_back(5); ms.push_back(m); radices.push_back(5); m /= 5; }
    return (int) kind.size() <= FFT_MAXST;
}

// tables of one line length behind fft_tab + off (doubles): tw[L] double2 | pos[L], freq[L] unsigned short
static size_t smem_bytes(int L) { return (size_t) L * FFT_T * 16 + (size_t) L * 16 + (((size_t) L * 4 + 15) & ~(size_t) 15); }
static size_t plan_bytes(int L) { return (size_t) L * 16 + (((size_t) L * 4 + 15) & ~(size_t) 15); }

static bool plan_fill(int L, unsigned char *h, FftPlan *P, const unsigned char *dev)
{
    std::vector<int> kind, ms, rad;
    if(!plan_stages(L, kind, ms, rad)) return false;
    double2 *tw = (double2 *) h;
    unsigned short *pos = (unsigned short *) (h + (size_t) L * 16), *freq = pos + L;
    for(int k = 0; k < L; k++) {
        const long double a = -2.0L * 3.14159265358979323846264338327950288L * k / L;
        tw[k].x = (double) cosl(a); tw[k].y = (double) sinl(a);
    }
    for(int k = 0; k < L; k++) {
        int kk = k, p = 0, m = L;
        for(size_t i = 0; i < rad.size(); i++) { const int q = kk % rad[i]; kk /= rad[i]; m /= rad[i]; p += q * m; }
        pos[k] = (unsigned short) p; freq[p] = (unsigned short) k;
    }
    P->L = L; P->nst = (int) kind.size();
    for(int i = 0; i < P->nst; i++) { P->kind[i] = kind[i]; P->m[i] = ms[i]; }
    P->tw = (const double2 *) dev;
    P->pos = (const unsigned short *) (dev + (size_t) L * 16);
    P->freq = P->pos + L;
    return true;
}

struct OwnFFT {
    int N = 0, Nzp = 0, threads = 256;
    FftPlan pN, pH;               // line lengths N (columns) and N/2 (packed z lines)
    const double2 *wN = nullptr;  // exp(-2 pi i k / N), k in [0, N/2]
    size_t smemN = 0, smemH = 0;
};

void pmfft_destroy(Engine *E)
{
    delete E->ownfft; E->ownfft = nullptr;
    E->fft_tab.release();
}

bool pmfft_supported(int N)
{
    std::vector<int> a, b, c;
    if(N < 8 || (N & 1) || N > 65534) return false;
    if(!plan_stages(N, a, b, c)) return false;
    a.clear(); b.clear(); c.clear();
    if(!plan_stages(N / 2, a, b, c)) return false;
    return smem_bytes(N) + 3 * (size_t) N * 8 <= 227 * 1024 && smem_bytes(N / 2) + (size_t) (N / 2) * FFT_T * 16 <= 227 * 1024;
}

// Tables and kernel attributes for mesh size N.  The half spectrum lives in E->cplx with row pitch Nzp.
int pmfft_init(Engine *E, int N)
{
    pmfft_destroy(E);
    if(!pmfft_supported(N)) return failmsg(E, "pmfft_init: mesh size not supported by the shared-memory transform");
    OwnFFT *F = new OwnFFT();
    E->ownfft = F;
    F->N = N;
    F->Nzp = (N / 2 + 1 + FFT_T - 1) / FFT_T * FFT_T;
    if(const char *s = getenv("B200_FFT_THREADS")) { const int v = atoi(s); if(v >= 64 && v <= 256 && v % 64 == 0) F->threads = v; }       // map_transposed() needs whole 8 x 8 thread groups
    const int L = N / 2;
    const size_t bN = plan_bytes(N), bH = plan_bytes(L), bW = (size_t) (L + 1) * 16;
    std::vector<unsigned char> h(bN + bH + bW);
    CK(E->fft_tab.ensure((bN + bH + bW) / 8 + 2));
    const unsigned char *dev = (const unsigned char *) E->fft_tab.p;
    if(!plan_fill(N, h.data(), &F->pN, dev) || !plan_fill(L, h.data() + bN, &F->pH, dev + bN)) return failmsg(E, "pmfft_init: plan");
    double2 *w = (double2 *) (h.data() + bN + bH);
    for(int k = 0; k <= L; k++) {
        const long double a = -2.0L * 3.14159265358979323846264338327950288L * k / N;
        w[k].x = (double) cosl(a); w[k].y = (double) sinl(a);
    }
    F->wN = (const double2 *) (dev + bN + bH);
    CK(cudaMemcpyAsync(E->fft_tab.p, h.data(), h.size(), cudaMemcpyHostToDevice, E->stream));
    CK(cudaStreamSynchronize(E->stream));
    F->smemN = smem_bytes(N);
    F->smemH = smem_bytes(L) + (size_t) L * FFT_T * 16;          // two tiles
    // The limit is a property of the kernel, not of this engine: engines with different mesh sizes share it, so it only
    // ever grows (one device per process).
    static size_t limH = 0, limN = 0, limP = 0;
    const size_t smemP = F->smemN + 3 * (size_t) N * 8;
    if(F->smemH > limH) {
        CK(cudaFuncSetAttribute(k_fft_z_forward, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) F->smemH));
        CK(cudaFuncSetAttribute(k_fft_z_inverse, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) F->smemH));
        limH = F->smemH;
    }
    if(F->smemN > limN) {
        CK(cudaFuncSetAttribute(k_fft_columns<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) F->smemN));
        CK(cudaFuncSetAttribute(k_fft_columns<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) F->smemN));
        CK(cudaFuncSetAttribute(k_fft_columns<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) F->smemN));
        limN = F->smemN;
    }
    if(smemP > limP) {
        CK(cudaFuncSetAttribute(k_fft_columns<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smemP));
        limP = smemP;
    }
    return 0;
}

size_t pmfft_cplx_doubles(const Engine *E) { return 2 * (size_t) E->ownfft->N * E->ownfft->N * E->ownfft->Nzp; }

// density mesh (E->mesh) -> potential mesh (E->mesh); E->cplx is the work area.  ps != NULL: [3][N] + 1 power-spectrum sums.
int pmfft_potential(Engine *E, double asmth2, double pot_factor, double binsperunit, double *ps)
{
    OwnFFT *F = E->ownfft;
    const int N = F->N, Nzp = F->Nzp, ntile = Nzp / FFT_T, th = F->threads;
    const long long nlines = (long long) N * N;
    const unsigned zblocks = (unsigned) ((nlines + FFT_T - 1) / FFT_T), cblocks = (unsigned) N * ntile;
    double2 *c = (double2 *) E->cplx.p;
    GreenArgs G = {N, N / 2 + 1, E->ktab.p, asmth2, pot_factor, binsperunit, ps};

    timer_start(E, T_PM_FFT_FWD);
    k_fft_z_forward<<<zblocks, th, F->smemH, E->stream>>>(E->mesh.p, c, nlines, Nzp, F->pH, F->wN);
    CKL(E);
    k_fft_columns<0><<<cblocks, th, F->smemN, E->stream>>>(c, ntile, (size_t) N * Nzp, (size_t) Nzp, F->pN, G);
    CKL(E);
    timer_stop(E, T_PM_FFT_FWD);

    timer_start(E, T_PM_TRANSFER);
    if(ps) k_fft_columns<3><<<cblocks, th, F->smemN + 3 * (size_t) N * 8, E->stream>>>(c, ntile, (size_t) Nzp, (size_t) N * Nzp, F->pN, G);
    else k_fft_columns<2><<<cblocks, th, F->smemN, E->stream>>>(c, ntile, (size_t) Nzp, (size_t) N * Nzp, F->pN, G);
    CKL(E);
    timer_stop(E, T_PM_TRANSFER);

    timer_start(E, T_PM_FFT_INV);
    k_fft_columns<1><<<cblocks, th, F->smemN, E->stream>>>(c, ntile, (size_t) N * Nzp, (size_t) Nzp, F->pN, G);
    CKL(E);
    k_fft_z_inverse<<<zblocks, th, F->smemH, E->stream>>>(c, E->mesh.p, nlines, Nzp, F->pH, F->wN);
    CKL(E);
    timer_stop(E, T_PM_FFT_INV);
    return 0;
}

// ---- the generic inverse pass: petapm_force_c2r (petapm.c:326-362) for callers that bring their own source spectrum and
// transfer functions (MP-GenIC's displacement_fields, libgenic/zeldovich.c:150-253; the radius filters of petapm_reion_c2r,
// petapm.c:416-577).  A transfer function is radial in the integer wave number, table[k2], filled by the caller with the
// reference's own formula, so the device applies exactly the factor the host function would.
//   kind 0 (density_transfer, zeldovich.c:276-289):  value *= table[k2]
//   kind a = 1, 2, 3 (disp_transfer, :291-313, axis x, y, z): fac = table[k2] * k_axis;  (re, im) <- (-im fac, re fac)
// The k2 = 0 mode passes unchanged (the reference's `if(k2)`); pm_apply_transfer_function copies src to dst first
// (petapm.c:1126-1130).
__global__ void __launch_bounds__(256)
k_pm_apply_radial(const double2 *__restrict__ src, double2 *__restrict__ dst, int N, int Nz, int Nzp, int kind, const double *__restrict__ table)
{
    const size_t total = (size_t) N * N * Nzp;
    for(size_t idx = (size_t) blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t) gridDim.x * blockDim.x) {
        const int iz = (int) (idx % Nzp);
        const size_t row = idx / Nzp;
        const int iy = (int) (row % N), ix = (int) (row / N);
        double2 v = make_double2(0.0, 0.0);
        if(iz < Nz) {
            v = src[idx];
            const int kx = ix <= N / 2 ? ix : ix - N, ky = iy <= N / 2 ? iy : iy - N, kz = iz;       // petapm_mesh_to_k petapm.c:81-84
            const long long k2 = (long long) kx * kx + (long long) ky * ky + (long long) kz * kz;
            if(k2) {
                if(kind == 0) { const double fac = table[k2]; v.x *= fac; v.y *= fac; }
                else {
                    const double fac = table[k2] * (kind == 1 ? kx : (kind == 2 ? ky : kz));
                    const double tmp = v.x;
                    v.x = -v.y * fac; v.y = tmp * fac;
                }
            }
        }
        dst[idx] = v;
    }
}

// readout_* of the caller (zeldovich.c:338-359): out[i] = sum over the 8 CIC cells of weight * mesh, weights and cell
// order of pm_iterate_one (petapm.c:955-1006)
__global__ void __launch_bounds__(256)
k_pm_readout_field(const double *__restrict__ pos, const uint8_t *__restrict__ flags, int64_t n, double cellsize, int N,
                   const double *__restrict__ mesh, double *__restrict__ out)
{
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    double acc = 0;
    if(!(flags[i] & 3)) {
        int ic[3]; double res[3];
#pragma unroll
        for(int k = 0; k < 3; k++) {
            const double t = pos[3 * i + k] / cellsize;
            const double f = floor(t);
            ic[k] = (int) f; res[k] = t - f;
        }
        int c0[3], c1[3];
#pragma unroll
        for(int k = 0; k < 3; k++) { c0[k] = ((ic[k] % N) + N) % N; c1[k] = (((ic[k] + 1) % N) + N) % N; }
        const double wx[2] = {1 - res[0], res[0]}, wy[2] = {1 - res[1], res[1]}, wz[2] = {1 - res[2], res[2]};
#pragma unroll
        for(int c = 0; c < 8; c++) {
            const int ox = c & 1, oy = (c >> 1) & 1, oz = (c >> 2) & 1;
            const double w = (wx[ox] * wy[oy]) * wz[oz];
            const size_t lin = ((size_t) (ox ? c1[0] : c0[0]) * N + (oy ? c1[1] : c0[1])) * N + (oz ? c1[2] : c0[2]);
            acc += w * mesh[lin];
        }
    }
    out[i] = acc;
}

// rho_k: host, complex [N][N][N/2+1] (x slowest).  For every function: transfer, inverse transform (the mirror passes of
// pmfft_potential: x, y, z), readout into f[j].out (host, n doubles).  E->mesh and E->cplx are overwritten.
int pmfft_c2r_readout(Engine *E, const double *rho_k, int nfunc, const b200_pm_function *f)
{
    OwnFFT *F = E->ownfft;
    if(!F) return failmsg(E, "b200_pm_c2r_readout: needs the engine's own transform passes (mesh size 2^a 3^b 5^c, b200_pm_init without B200_PM_FFT=cufft)");
    if(!rho_k || nfunc < 0 || (nfunc > 0 && !f)) return failmsg(E, "b200_pm_c2r_readout: null argument");
    const int N = F->N, Nz = N / 2 + 1, Nzp = F->Nzp, ntile = Nzp / FFT_T, th = F->threads;
    const long long nlines = (long long) N * N;
    const unsigned zblocks = (unsigned) ((nlines + FFT_T - 1) / FFT_T), cblocks = (unsigned) N * ntile;
    const size_t nk2 = 3 * (size_t) (N / 2) * (N / 2) + 1;
    const size_t n = (size_t) (E->n > 0 ? E->n : 1);
    CK(E->pm_rhok.ensure(2 * (size_t) N * N * Nzp));
    CK(E->pm_table.ensure(nk2));
    CK(E->d_pot.ensure(n));
    CK(cudaMemcpy2DAsync(E->pm_rhok.p, (size_t) Nzp * 16, rho_k, (size_t) Nz * 16, (size_t) Nz * 16, (size_t) nlines, cudaMemcpyHostToDevice, E->stream));
    GreenArgs G = {N, Nz, E->ktab.p, 0.0, 0.0, 0.0, nullptr};
    double2 *c = (double2 *) E->cplx.p;
    for(int j = 0; j < nfunc; j++) {
        if(f[j].kind < 0 || f[j].kind > 3 || !f[j].table || (E->n > 0 && !f[j].out)) return failmsg(E, "b200_pm_c2r_readout: bad function entry");
        CK(cudaMemcpyAsync(E->pm_table.p, f[j].table, nk2 * sizeof(double), cudaMemcpyHostToDevice, E->stream));
        // the timers keep the phases of the last function (b200_get_timings: pm_transfer, pm_fft_inverse, pm_readout)
        timer_start(E, T_PM_TRANSFER);
        k_pm_apply_radial<<<148 * 16, 256, 0, E->stream>>>((const double2 *) E->pm_rhok.p, c, N, Nz, Nzp, f[j].kind, E->pm_table.p);
        CKL(E);
        timer_stop(E, T_PM_TRANSFER);
        timer_start(E, T_PM_FFT_INV);
        k_fft_columns<1><<<cblocks, th, F->smemN, E->stream>>>(c, ntile, (size_t) Nzp, (size_t) N * Nzp, F->pN, G);          // x
        CKL(E);
        k_fft_columns<1><<<cblocks, th, F->smemN, E->stream>>>(c, ntile, (size_t) N * Nzp, (size_t) Nzp, F->pN, G);          // y
        CKL(E);
        k_fft_z_inverse<<<zblocks, th, F->smemH, E->stream>>>(c, E->mesh.p, nlines, Nzp, F->pH, F->wN);
        CKL(E);
        timer_stop(E, T_PM_FFT_INV);
        if(E->n > 0) {
            timer_start(E, T_PM_READOUT);
            k_pm_readout_field<<<(unsigned) ((E->n + 255) / 256), 256, 0, E->stream>>>(E->pos.p, E->flags.p, E->n, E->Box / N, N, E->mesh.p, E->d_pot.p);
            CKL(E);
            timer_stop(E, T_PM_READOUT);
            CK(cudaMemcpyAsync(f[j].out, E->d_pot.p, (size_t) E->n * sizeof(double), cudaMemcpyDeviceToHost, E->stream));
        }
        CK(cudaStreamSynchronize(E->stream));          // the table buffer is reused by the next function
    }
    E->potential_valid = false;
    return 0;
}

} // namespace b200
