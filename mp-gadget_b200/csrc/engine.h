// engine.h -- internal state of the B200 force engine (not part of the C-ABI).
#pragma once
#include <cuda_runtime.h>
#include <cufft.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/b200force.h"

namespace b200 {

// Grow-only device buffer.
template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;   // elements
    cudaError_t ensure(size_t n) {
        if(n <= cap) return cudaSuccess;
        if(p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = n + n / 16 + 64;
        cudaError_t e = cudaMalloc((void **) &p, want * sizeof(T));
        if(e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if(p) cudaFree(p); p = nullptr; cap = 0; }
};

struct Timer {
    cudaEvent_t a = nullptr, b = nullptr;
    bool used = false;
};

enum TimerId {
    T_PM_DEPOSIT, T_PM_FFT_FWD, T_PM_TRANSFER, T_PM_FFT_INV, T_PM_GRADIENT, T_PM_READOUT,
    T_TREE_KEYS, T_TREE_SORT, T_TREE_NODES, T_TREE_MOMENTS,
    T_WALK, T_WALK_POST, T_H2D, T_D2H, T_SPH_DENSITY, T_SPH_HYDRO, T_COUNT
};

// Node record used by the walk, DFS order.  A,B are 32-byte rows so that one
// warp-uniform LDG.128 pair fetches each.
struct NodeAux {           // int4
    int sibling;           // DFS index of the next node when skipping the subtree, -1 at the end
    int pstart;            // first particle (sorted order) of the subtree
    int count;             // particles in the subtree
    int leaf;              // 1: particle leaf, 0: internal (first child = self + 1)
};

struct SlabPM;
struct Sharded;
struct OwnFFT;

#define B200_SPART_PAD 16     // massless far-away rows behind spart[np) (tree_walk.cu pair loop)

struct Engine {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;     // host<->device copies that overlap kernels on `stream`
    cudaEvent_t chunk_ev[66] = {};           // [0..64] walk groups of the AoS step, [65] its span-B ingest
    cudaStream_t side_stream = nullptr;     // PM long-range step when it runs concurrently with the tree walk
    cudaEvent_t fork_ev = nullptr, join_ev = nullptr;
    std::string err;
    int64_t launches = 0;

    // ---- particles (original index order) ----
    int64_t n = 0;
    DevBuf<double> pos;        // [n][3]
    DevBuf<float> mass;        // [n]
    DevBuf<uint8_t> type;      // [n]
    DevBuf<uint8_t> flags;     // [n] bit0 garbage bit1 swallowed
    DevBuf<double> oldacc;     // [n] |FullTreeGravAccel + GravPM| (not divided by G)
    DevBuf<double> last_tree_acc;  // [n][3]
    DevBuf<double> last_pm_acc;    // [n][3]
    bool have_last_tree = false, have_last_pm = false;
    DevBuf<uint8_t> aos;       // staging for AoS ingest / write-back

    // ---- PM ----
    double Box = 0, Asmth = 0, G = 0;
    int Nmesh = 0;
    int NmeshWalk = 0;         // mesh size seen by the short-range walk (set by pm_init / pmslab_init)
    cufftHandle plan_fwd = 0, plan_inv = 0;
    bool plans = false;
    DevBuf<double> mesh;       // real mesh Nmesh^3 (density, then potential)
    DevBuf<double> cplx;       // Nmesh^2 (Nmesh/2+1) complex, interleaved
    DevBuf<double> fmesh;      // 3 force meshes
    DevBuf<double> ktab;       // per-dimension deconvolution factor 1/sinc^2, [Nmesh]
    DevBuf<uint8_t> fftwork;
    OwnFFT *ownfft = nullptr;  // pm_fft.cu: the shared-memory transform passes (null: cuFFT plans above)
    DevBuf<double> fft_tab;    // its twiddle and row-order tables
    DevBuf<double> pm_rhok, pm_table;   // b200_pm_c2r_readout: the caller's source spectrum (padded row pitch), one transfer table
    bool potential_valid = false;
    bool pm_fused = true;      // difference + readout in one kernel, no force meshes (B200_PM_FUSED=0: separate passes)
    bool fmesh_valid = false;
    bool pm_power = false;     // accumulate the matter power spectrum inside the Green's-function pass
    bool pm_ps_valid = false;
    DevBuf<double> pm_ps;      // [3][Nmesh] Power, kk, Nmodes sums + Norm
    SlabPM *slab = nullptr;
    Sharded *sh = nullptr;      // multi-GPU driver state (sharded.cu)

    // ---- tree ----
    bool tree_valid = false;
    double tree_box = 0;
    int64_t tree_np = 0;        // particles in the tree
    int64_t tree_nn = 0;        // nodes
    int tree_maxdepth = 0;
    int tree_overfull = 0;
    int tree_topdepth = 0;
    std::vector<int> tree_lvl;  // BFS level offsets of the current tree
    bool tree_full = false;     // contains every particle (full_particle_tree_flag)
    DevBuf<unsigned long long> keys, keys_alt;
    DevBuf<int> sidx, sidx_alt;   // sorted -> original index
    DevBuf<uint8_t> cubtemp;
    DevBuf<double> spart;       // [np] double4 {x,y,z,m} in sorted order
    DevBuf<double> spart_xy, spart_zm;   // the same as two double2 streams (pair kernel)
    // build-time (BFS order) node fields
    DevBuf<int> b_start, b_count, b_father, b_sibling, b_firstchild, b_nchild, b_level, b_size, b_dfs, b_scan;
    DevBuf<double> b_center;    // [.][4] cx,cy,cz,len
    // final (DFS order)
    DevBuf<double> nodeA;       // [nn] double4 cofm.xyz, mass
    DevBuf<double> nodeB;       // [nn] double4 center.xyz, len
    DevBuf<int> nodeC;          // [nn] NodeAux
    DevBuf<int> nodeF;          // [nn] father (DFS)
    DevBuf<int> nodeK;          // [nn][8] DFS positions of the children (-1 = none)
    DevBuf<double> nodeH;       // [nn] hmax
    DevBuf<int> scratch_i;      // small device scalars; always ensure(256): a later, larger ensure() would
                                // reallocate and drop counters that are already in flight.  Slots: [0..15] tree build
                                // (cleared by every build), [8] walk target select, [12..13] SPH pass counters,
                                // [20] step-loop list select, [24] exchange-list select, [32..63] walk chunk offsets,
                                // [64..67] piece-pool control

    // ---- SPH (original index order unless noted) ----
    DevBuf<double> s_vel, s_hsml, s_entropy, s_dtentropy, s_fullacc, s_gravpm, s_hydroacc;
    bool s_have[7] = {false, false, false, false, false, false, false};
    DevBuf<double> s_velpred, s_evp, s_density, s_egy, s_dhsmlfac, s_divvel, s_curlvel, s_dthsml, s_numngb, s_gradrho;
    DevBuf<double> s_svel, s_hA, s_hB;      // curve order: double4 rows
    DevBuf<double> s_out3, s_out1a, s_out1b;
    DevBuf<int> s_outi, s_outi2, s_niter, s_nint;
    DevBuf<double> s_left, s_right;       // smoothing-length brackets of the density iteration
    DevBuf<double> s_hD;                  // curve order: two more double4 rows per particle (hC, hT of k_sph_gather_hydro)
    DevBuf<double> s_bins;                // [5][B200_TIMEBINS + 1] per-bin factors
    DevBuf<uint8_t> s_bin_grav, s_bin_hydro, s_active;
    DevBuf<int> sph_list_a, sph_list_b;   // target lists of the density passes
    bool s_bins_set = false, s_active_set = false;
    double sph_chunks_per_warp = 4.0;
    int sph_passes = 0;
    bool sph_density_done = false;
    int sph_DoEgy = 0;

    // ---- walk ----
    DevBuf<int> targets;        // walk target list (original indices)
    DevBuf<int> targets_sorted; // the same in tree (curve) order
    DevBuf<uint8_t> walk_flags;
    DevBuf<double> d_acc, d_pot;  // outputs [n][3], [n]
    DevBuf<int> d_counts;       // [n] b200_walk_counts
    DevBuf<float> srtab;        // 2*512 short-range window tables
    DevBuf<unsigned> walk_pool; // chunk pool of the per-target leaf-piece lists (tree_walk.cu)
    DevBuf<int> walk_chunktab, walk_cnt;
    DevBuf<double> walk_partial;
    double walk_chunks_per_warp = 6.0;
    size_t walk_want = 0;       // chunks to allocate for the next walk (0: estimate)
    int walk_maxch = 128;       // chunk-table entries per warp of the current call (piece_list.cuh)
    double walk_pieces = 0;     // leaf pieces (<= 8 particles each) queued by the last walk
    int walk_chunks = 0;        // pool chunks it used

    // ---- step loop (steploop.cu): lists and scratch of the device-resident drift / kick / time-bin code ----
    bool st_state = false, st_have_gas = false;
    double st_box = 0;                                 // PartManager->BoxSize when given with the state
    DevBuf<int> st_iota, st_act, st_listA, st_listB;   // identity, the active list, two sub-list buffers
    DevBuf<uint8_t> st_flag;
    int64_t st_nact = 0, st_nsub = 0;
    bool st_act_implicit = true;                       // PM step: every particle, no list (ActiveParticle == NULL)
    DevBuf<double> st_store, st_lower;                 // [n][3] StoredGravAccel / accelerations of the lower levels
    bool st_store_valid = false;
    DevBuf<double> st_sync, st_part, st_tab, st_maxsig;  // st_maxsig: SphP[].MaxSignalVel by particle index
    bool st_maxsig_valid = false;
    DevBuf<unsigned long long> st_cnt;

    // ---- domain keys (domain_keys.cu) ----
    DevBuf<unsigned long long> dk_keys, dk_startkey;   // Peano-Hilbert key per particle; TopNodes[].StartKey
    DevBuf<uint8_t> dk_tab;                            // generated state machine of the curve
    DevBuf<int> dk_daughter, dk_shift, dk_leaf, dk_topleaf, dk_xlist, dk_iota, dk_task;   // dk_xlist: particles leaving this task
    DevBuf<uint8_t> dk_xflag;
    DevBuf<unsigned long long> dk_counts, dk_sample;
    int dk_ntop = 0;
    DevBuf<int> b_top;                                 // tree build below the domain's top nodes: top node | curve state << 24 per cell
    int64_t dk_keys_n = -1, dk_topleaf_n = -1;

    // ---- friends-of-friends (fof.cu) ----
    DevBuf<unsigned long long> fof_key, fof_key_alt;   // grid cell of every primary particle, sorted / unsorted
    DevBuf<int> fof_val, fof_val_alt, fof_parent, fof_root;
    DevBuf<double> fof_spos;                           // positions in cell order
    DevBuf<long long> fof_ids, fof_out;
    DevBuf<unsigned long long> fof_min;

    Timer timers[T_COUNT];
    b200_timings last = {};
};

// error helpers -----------------------------------------------------------
int fail(Engine *e, const char *what, cudaError_t err, const char *file, int line);
int failmsg(Engine *e, const std::string &msg);
#define CK(call) do { cudaError_t _e = (call); if(_e != cudaSuccess) return b200::fail(E, #call, _e, __FILE__, __LINE__); } while(0)
#define CKL(E_) do { (E_)->launches++; cudaError_t _e = cudaGetLastError(); if(_e != cudaSuccess) return b200::fail((E_), "kernel launch", _e, __FILE__, __LINE__); } while(0)

void timer_start(Engine *E, int id);
void timer_stop(Engine *E, int id);
double timer_ms(Engine *E, int id);

// PM (pm.cu)
int pm_init(Engine *E, double Box, double Asmth, int Nmesh, double G);
void pm_destroy(Engine *E);
int pm_deposit(Engine *E);
int pm_force(Engine *E, double *d_gravpm, double *d_pot);   // device outputs, [n][3] / [n], may be null
int pm_force_meshes(Engine *E);
int pm_cell_index(Engine *E, int32_t *d_icell);

// PM transforms (pm_fft.cu)
bool pmfft_supported(int N);
int pmfft_init(Engine *E, int N);
void pmfft_destroy(Engine *E);
size_t pmfft_cplx_doubles(const Engine *E);
int pmfft_potential(Engine *E, double asmth2, double pot_factor, double binsperunit, double *ps);
int pmfft_c2r_readout(Engine *E, const double *rho_k, int nfunc, const b200_pm_function *f);

// tree (tree_build.cu)
int tree_build(Engine *E, double Box, int mask, const int32_t *d_active, int64_t nactive,
               int toplevel_depth, b200_tree_info *info);
int tree_export(Engine *E, double *center, double *len, double *cofm, double *mass, double *hmax,
                int32_t *sibling, int32_t *firstchild, int32_t *nocc, int32_t *leafpart);

int tree_top_get(Engine *E, int level, double *d_out);
int tree_top_set(Engine *E, int level, const double *d_in);

// slab-decomposed PM for multi-GPU (pm_slab.cu)
struct SlabPM;
int pmslab_init(Engine *E, double Box, double Asmth, int Nmesh, double G, int rank, int nranks, int halo,
                void **real_buf, void **cplx_buf, void **cplxT_buf);
void pmslab_destroy(Engine *E);
int pmslab_deposit(Engine *E, int64_t n_own, bool check = true);
int pmslab_fft2d(Engine *E, int inverse);
int pmslab_fft1d(Engine *E, int inverse);
int pmslab_transfer(Engine *E);
int pmslab_readout(Engine *E, int64_t n_own, double *d_gravpm, double *d_pot, bool check = true);

// multi-GPU driver (sharded.cu)
void sharded_destroy(Engine *E);
int collect_timings(Engine *E);       // capi.cu: Engine::last from the event timers

// SPH (sph.cu)
int sph_set_gas(Engine *E, const double *vel, const double *hsml, const double *entropy, const double *dtentropy,
                const double *fullacc, const double *gravpm, const double *hydroacc);
int sph_density(Engine *E, const b200_sph_params *p, int update_hsml, int DoEgy, int *d_ninteract, int *d_niter);
int sph_hydro(Engine *E, const b200_sph_params *p, double *d_acc, double *d_dte, double *d_maxsig, int *d_ninteract);
int sph_update_hmax(Engine *E);
int sph_set_hsml_range(Engine *E, const double *hsml, int64_t first, int64_t count);
int sph_set_timebins(Engine *E, const uint8_t *bin_grav, const uint8_t *bin_hydro, const b200_sph_bins *bins);
int sph_set_active(Engine *E, const int32_t *active, int64_t nactive);
int sph_set_state(Engine *E, const double *density, const double *egy, const double *dhsmlfac, const double *divvel, const double *curlvel);

// step loop (steploop.cu), domain keys (domain_keys.cu)
void step_release(Engine *E);
void domain_release(Engine *E);
int domain_need_tables(Engine *E);    // domain_keys.cu: the curve's state machine in dk_tab

// friends-of-friends (fof.cu)
int fof_primary(Engine *E, const int64_t *ids, int mask, double Box, double ll, int64_t *minid_out, int64_t *ngroups_out);
void fof_release(Engine *E);

// walk (tree_walk.cu)
int walk_init_tables(Engine *E);
// presorted: d_active is a device list already in tree (curve) order; it is walked as given
int grav_short_tree(Engine *E, const b200_gravshort_params *par, const int32_t *d_active,
                    int64_t nactive, double *d_acc, double *d_pot, b200_walk_counts *d_counts, bool presorted = false);
int walk_chunk_targets(Engine *E, int nchunks, int64_t chunk, int *offsets);

} // namespace b200

struct b200_ctx { b200::Engine e; };
