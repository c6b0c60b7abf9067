// psc_b200: the per-rank device context behind the C ABI (include/psc_b200.h).
//
// One context = one process = one GPU.  It owns
//   - the particle store: two float4 streams (xi4 = x,y,z,kind ; pxi4 = ux,uy,uz,q*w),
//     double-buffered, all patches back to back, patch p = [off[p], off[p+1]);
//     when `sorted` is set the store is ordered by (patch, cell) and cell_off holds
//     the per-cell offsets (what the tiled push kernel walks)
//   - the field arrays in PSC's layout float [slot][m][iz][iy][ix]; slots
//     0..n_patches-1 are this rank's patches, the following ones are proxies of
//     patches owned by neighbouring ranks (filled by the NCCL halo exchange)
//   - scratch memory, the stream, NCCL state and per-kernel timers.
#pragma once

#include "grid.hpp"

#include <cuda_runtime.h>

#include <cstdint>
#include <map>
#include <string>
#include <vector>

namespace psc_b200
{

void set_error(const std::string& msg);
int fail(const std::string& msg);

#define PSC_CUDA_TRY(expr)                                                               \
  do {                                                                                   \
    cudaError_t e__ = (expr);                                                            \
    if (e__ != cudaSuccess) {                                                            \
      return ::psc_b200::fail(std::string(#expr) + ": " + cudaGetErrorString(e__));      \
    }                                                                                    \
  } while (0)

#define PSC_TRY(expr)                                                                    \
  do {                                                                                   \
    int rc__ = (expr);                                                                   \
    if (rc__) {                                                                          \
      return rc__;                                                                       \
    }                                                                                    \
  } while (0)

// grid facts every kernel needs, passed by value
struct GridDev
{
  int ldims[3], im[3], ibn[3];
  int n_patches; // local
  int n_cells;   // per patch
  long fld_len;  // im0*im1*im2
  int dim;       // pm::DIM_*
  int deposit;   // pm::DEPOSIT_*
  pm::PushConst pc;
};

// grow-only device allocation
struct DevBuf
{
  void* p = nullptr;
  size_t bytes = 0;
  int reserve(size_t n);
  void release();
  template <typename T>
  T* as()
  {
    return static_cast<T*>(p);
  }
};

struct FieldArr
{
  float* d = nullptr;
  int n_comps = 0;
};

struct ProfEntry
{
  const char* name;
  float ms = 0.f;
  uint64_t launches = 0;
};

struct Comm; // nccl.cpp

struct Ctx
{
  GridHost g;
  GridDev gd;
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr; // step(): the field chain runs here next to the particle sort
  cudaStream_t stream_io = nullptr; // step_begin(): the caller's async host transfers
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  cudaEvent_t ev_j_ready = nullptr, ev_flds_done = nullptr; // field chain: J final / E, H final

  // ---- options (psc_b200_set_option)
  int opt_tiled = 1;       // use the tiled shared-memory push when the store is sorted
  int opt_fma = 0;         // 0: -fmad=false build of the push (bit-exact vs CPU), 1: FMA build
  int opt_tma = 1;         // stage the E/B tile with TMA (tensor map / cp.async.bulk) instead of LDG/STS
  int opt_pull = 0;         // the push of the next step completes the sort of this one (one pass less over the
                            // particles; byte-identical stores, but 40.6 ms against 33.1 ms per S3D step as it
                            // stands -- DESIGN.md 3.2d): opt-in
  int opt_cell_moments = 1; // moments of a cell-ordered store: summed per cell in registers, one add per cell and node
  int opt_pull_cap = 0;     // > 0: capacity of the pull push's mover list (tests: force the overflow route)
  int opt_push_collect = 1; // multi-rank: k_push_lean lists the remote leavers (else a pass over the boundary cells)
  int opt_lean = 1;        // k_push_lean (push_lean.cuh) whenever the tile geometry is compile-time
  int opt_vec_fields = 1;  // Yee update with 128-bit accesses where the rows allow it (3D, im0 % 4 == 0)
  int opt_threads = 256;   // CTA size of the tiled push
  int opt_min_blocks = 3;  // resident CTAs per SM the tiled push is compiled for
  int opt_tile[3] = {0, 0, 0}; // cells per tile edge, 0 = default
  int opt_profile = 0;
  int opt_fused_sort = 1;  // step(): fuse boundary exchange and sort when possible
  int opt_keep_sorted = 1; // keep the store cell-ordered on every step, whatever the deck's sort_interval:
                           // the push sorts an unordered store first, the exchange after a sorted push is
                           // the fused exchange + sort.  Same physics (only the particle order differs from
                           // a run that sorts every 10th step); the unsorted-store push is 6x slower.
  int opt_overlap = 0;     // step(): J ghosts + Yee on a second (high-priority) stream next to the particle
                           // sort; measured +0.2 % only (the field kernels fill the GPU while they run)
  int opt_gapped = 0;      // step(): gapped store (gap.cuh), no sort pass; needs fused_sort.  Off: its push
                           // variant is still slower than push + fused sort
  int opt_gap_slack = 0;   // free slots behind every cell's run, 0 = half the mean population

  // ---- particles
  float4* xi4[2] = {nullptr, nullptr};
  float4* pxi4[2] = {nullptr, nullptr};
  int cur = 0;
  size_t cap = 0;
  uint32_t n_prts = 0;
  std::vector<uint32_t> h_off; // n_patches + 1
  uint32_t* d_off = nullptr;
  bool sorted = false;           // store ordered by (patch, cell) and cell_off valid
  bool pushed_from_sorted = false; // store = cell-ordered store after exactly one push;
                                   // cell_off still describes the pre-push cell runs
  uint64_t n_fused = 0;          // steps that took the fused boundary+sort pass
  bool want_counts = false;      // step(): the next push should count per-cell destinations
  bool counts_valid = false;     // scr[9]/scr[11] hold the counts of the last push
  uint32_t* d_cell_off = nullptr; // n_patches * n_cells + 1
  uint32_t* d_cell_off_alt = nullptr; // written by the fused boundary+sort pass
  uint64_t n_fused_fallback = 0;
  // pull mode (push_lean.cuh PULL, fused_sort.cu): after a step the particles that changed cell
  // sit in the OTHER buffer (laid out by d_cell_off_alt, scr[13] = per cell {arrivals in front
  // of the stayers, stayers}), the stayers still in this one among the dead copies of the
  // leavers; the next push pulls them across.  pull_materialize() completes the sort instead
  // whenever something else wants the store.
  bool pull_pending = false;
  bool pulled = false;        // the last push was a pull push (its mover list is in scr[12])
  bool fs_pull = false;       // the fused pass in flight placed the movers only
  uint32_t mv_cap = 0, last_n_movers = 0;
  uint64_t n_pulled = 0, n_pull_materialized = 0, n_pull_overflow = 0;
  // step_begin / step_end (pipelined host I/O): the exchange + sort is in flight on `stream`,
  // the field chain and the host transfers on `stream2`
  bool step_pending = false;
  bool fs_deferred = false;      // ... and its flags / new offsets have not been read back yet
  uint32_t* fs_host = nullptr;   // pinned landing zone of that read-back
  size_t fs_host_bytes = 0;
  uint32_t fs_n_expected = 0;    // particle count the deferred commit must find (multi-rank: after the exchange)
  bool fs_multi = false;
  uint64_t n_lean = 0;           // pushes that ran k_push_lean (stat "lean_pushes")
  double* en_host = nullptr;     // pinned: DiagEnergies of the last step that asked for them
  bool en_valid = false;
  bool want_scatter_energies = false;  // ask the next fused scatter to reduce the particle energies ...
  bool scatter_energies_done = false;  // ... and whether it did (single rank, no fallback)
  uint64_t n_dropped = 0;        // absorbed at open/absorbing walls so far

  // ---- gapped store (gap.cuh): when `gapped` is set, xi4/pxi4[cur] hold one run per cell,
  // [g_start[t], g_start[t] + g_n[t]) inside the slab [g_v[t], g_v[t + 1]); h_off / n_prts
  // describe the contiguous patch-by-patch sequence the store stands for
  bool gapped = false;
  uint64_t n_gap_steps = 0;      // steps taken on the gapped path
  uint64_t n_gap_relayouts = 0;  // ... of which re-laid the store out (a slab overflowed)
  uint64_t n_gap_redone = 0;     // steps redone on the eager path
  uint32_t* g_v = nullptr;       // slab starts [nct + 1]
  uint32_t* g_v_alt = nullptr;
  uint32_t* g_start[2] = {nullptr, nullptr}; // [0]: current runs, [1]: written by the step
  uint32_t* g_n[2] = {nullptr, nullptr};
  uint32_t* g_nstay = nullptr;   // stayers per cell of the last push
  uint32_t* g_ctl = nullptr;     // GAP_CTL_WORDS control words
  uint32_t g_rl = 8;             // slots kept free in front of the stayers (adaptive)
  uint32_t g_slack = 0;          // slots kept free behind a run at layout time
  float4* mvx = nullptr;         // mover list
  float4* mvp = nullptr;
  uint4* mvtag = nullptr;
  size_t mov_cap = 0;
  uint32_t g_mov_used = 0;       // mover slots used by the last gapped push

  // ---- per-patch tables (device)
  pm::PatchBnd* d_patch_bnd = nullptr; // n_patches
  int* d_nei_slot = nullptr;   // [n_patches][27] field slot of the neighbour or -1
  int* d_nei_patch = nullptr;  // [n_patches][27] local patch index, -1 none, -2-r remote rank r
  int8_t* d_add_order = nullptr; // [n_patches][26] receiver-side dir idx in reference order
  std::vector<int> h_nei_patch, h_nei_slot;
  int n_slots = 0; // n_patches + proxies
  std::vector<int> proxy_gp; // global patch index of each proxy slot (ascending)

  // ---- fields
  std::vector<FieldArr> flds;

  // ---- scratch
  DevBuf scr[14];
  DevBuf stage; // host<->device staging of AoS records
  void* h_pinned = nullptr;
  size_t h_pinned_bytes = 0;

  // ---- checks
  int rho_m_id = -1, rho_p_id = -1, div_id = -1;
  double last_continuity = 0., last_gauss = 0.;

  // ---- timers
  cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
  std::vector<ProfEntry> prof;
  std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> prof_pending;
  uint64_t n_launches = 0;

  // ---- multi-GPU
  Comm* comm = nullptr;
  DevBuf rf_cells;      // cells from which particles can leave for another rank
  uint32_t n_rf_cells = 0;
  bool rf_built = false;
  // the lean push lists the remote leavers itself (push.cu launch_lean): capacity of the
  // list in scr[7] for the step in flight (0 = none), and last step's count to size it by
  uint32_t rem_cap = 0;
  uint32_t last_n_rem = 0;

  float4* xi() { return xi4[cur]; }
  float4* pxi() { return pxi4[cur]; }
  float4* xi_alt() { return xi4[cur ^ 1]; }
  float4* pxi_alt() { return pxi4[cur ^ 1]; }
  float* fld(int id) { return flds[id].d; }
  long fld_slot_len(int id) const { return gd.fld_len * flds[id].n_comps; }
};

// RAII-free kernel timing helper: KernelScope ks(ctx, "name"); launch...; (dtor records)
struct KernelScope
{
  Ctx* c;
  int idx = -1;
  cudaEvent_t a = nullptr, b = nullptr;
  KernelScope(Ctx* ctx, const char* name, int n_launches = 1);
  ~KernelScope();
};

// ---- particles.cu
int prts_reserve(Ctx* c, size_t n);
int prts_set(Ctx* c, const void* aos, const uint32_t* n_by_patch);
int prts_inject(Ctx* c, const void* aos, const uint32_t* n_by_patch);
int prts_get(Ctx* c, void* aos, uint32_t* off);
int prts_setup_thermal(Ctx* c, int ppc, const int* ppc_by_patch, const double* vth, uint64_t seed);
int prts_upload_off(Ctx* c);
// d_n: particle count read on the device (a deferred sort has not told the host yet); alt:
// the store the fused sort is writing
int prts_energies(Ctx* c, double out2[2], bool sync = true, const uint32_t* d_n = nullptr, bool alt = false);
int selftest_math(Ctx* c, uint64_t* n_bad);

// ---- push.cu (two builds of the same source: exact = -fmad=false, fast = FMA)
int push_mprts_exact(Ctx* c);
int push_mprts_fast(Ctx* c);
int deposit_paths_exact(Ctx* c, const psc_b200_jpath* d_paths, uint32_t n);
int deposit_paths_fast(Ctx* c, const psc_b200_jpath* d_paths, uint32_t n);

// ---- sort.cu
int sort_mprts(Ctx* c);
int sort_pairs(Ctx* c, uint32_t* keys, uint32_t* vals, uint32_t* keys_alt, uint32_t* vals_alt,
               size_t n, int key_bits, bool iota_vals, bool* result_in_alt);
// boundary exchange + sort of a pushed, previously sorted store; defer: see step_begin (capi.cu)
int fused_bnd_sort(Ctx* c, bool defer = false);
int fused_bnd_sort_finish(Ctx* c);
// gapped store (gap.cuh)
int gap_prepare(Ctx* c, bool* ok); // allocations, layout, per-step clears; *ok = false: not applicable
int gap_finish(Ctx* c, bool* redo); // offsets, mover placement, commit; *redo: take the eager path
int gap_compact(Ctx* c);           // gapped store -> contiguous cell-ordered store (sorted = true)
void gap_release(Ctx* c);
int push_gap_exact(Ctx* c);
int push_gap_fast(Ctx* c);
// every operator that reads the particle store directly calls this first
int pull_materialize(Ctx* c); // fused_sort.cu: pull mode -> contiguous cell-ordered store
int pull_enter(Ctx* c);       // ... and into it from a cell-ordered store (no arrivals yet)
bool pull_possible(const Ctx* c);
inline int store_ready(Ctx* c)
{
  if (c->pull_pending) {
    return pull_materialize(c);
  }
  return c->gapped ? gap_compact(c) : 0;
}

// ---- bndp.cu
int bnd_particles(Ctx* c);

// ---- collision.cu
int collide(Ctx* c, const psc_b200_collision_params* prm, uint64_t* n_collisions);
int heating_spot_foil(Ctx* c, const psc_b200_heating_params* prm, uint64_t* n_kicked);

// ---- fields.cu
int flds_create(Ctx* c, int n_comps, int* id);
int flds_add(Ctx* c, int y_id, int y_mb, int x_id, int x_mb, int n_comps);
int flds_scale(Ctx* c, int id, int mb, int me, double a);
int flds_download_interior(Ctx* c, int id, int mb, int me, float* host);
int flds_zero(Ctx* c, int id, int mb, int me);
int flds_fill(Ctx* c, int id, int m, float v);
int flds_upload(Ctx* c, int id, int mb, int me, const float* host, bool sync = true);
int flds_download(Ctx* c, int id, int mb, int me, float* host, bool sync = true);
int bnd_fill_ghosts(Ctx* c, int id, int mb, int me);
int bnd_add_ghosts(Ctx* c, int id, int mb, int me);
int bndf_fill_ghosts_E(Ctx* c);
int bndf_fill_ghosts_H(Ctx* c);
int bndf_add_ghosts_J(Ctx* c);
int push_E(Ctx* c, double dt_fac);
int push_H(Ctx* c, double dt_fac);
int moment_rho_1st_nc(Ctx* c, int id);
int moment_n_comps(const Ctx* c, int which);
int moment_1st(Ctx* c, int id, int which);
int marder(Ctx* c, double diffusion, int loop);
int check_continuity_begin(Ctx* c);
int check_continuity_end(Ctx* c, double* err);
int check_gauss(Ctx* c, double* err);
int field_energies(Ctx* c, double out6[6], bool sync = true);

// ---- comm.cpp (NCCL through dlopen; nothing here runs unless nccl_init was called)
int comm_unique_id(void* id128);
int comm_init(Ctx* c, const void* id128);
void comm_destroy(Ctx* c);
int comm_halo_exchange(Ctx* c, int id, int mb, int me, bool add);
// ships particles leaving for other ranks; `fixup`: apply the boundary fix-ups while
// packing (the fused pass leaves the source store untouched)
int comm_exchange_particles(Ctx* c, const float4* xi_src, const float4* pxi_src,
                            const uint32_t* d_src_idx, const uint32_t* d_keys, uint32_t n_remote,
                            uint32_t key_remote_base, bool fixup,
                            std::vector<uint32_t>& n_recv_by_patch, float4** xi_recv,
                            float4** pxi_recv);
int comm_allreduce_max(Ctx* c, double* v, int n);
int comm_allreduce_sum(Ctx* c, double* v, int n);

// ---- checkpoint.cu
int checkpoint_write(Ctx* c, const char* path, int64_t timestep);
int checkpoint_read(Ctx* c, const char* path, int64_t* timestep);

// ---- tables
int build_patch_tables(Ctx* c);

// a CUtensorMap (cuda.h) over field array `id`, seen as (x, y, z, component x slot) -- or
// (y, z, component x slot) for rank 3 -- with the given box; encoded through the driver entry
// point the runtime hands out (no link against libcuda)
struct alignas(64) TensorMap128
{
  unsigned char b[128];
};
int field_tile_tensor_map(Ctx* c, int id, int rank, const int* box, TensorMap128* out);

// ---- balance (comm.cu)
int balance(Ctx* c, double factor_fields, int* changed);

} // namespace psc_b200
