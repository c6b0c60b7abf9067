// psc_b200: the C ABI (include/psc_b200.h) -- context life cycle, per-patch tables,
// options/timers, the Psc::step sequence and the entry points that forward to the
// operator implementations.  No exception crosses this boundary and there is no CPU
// fallback: without a CUDA device psc_b200_create fails.
#include "dev_util.cuh"

#include <cuda.h> // CUtensorMap and its enums (types only)

#include <algorithm>
#include <cstring>
#include <exception>

namespace psc_b200
{

static thread_local std::string g_last_error;

void set_error(const std::string& msg)
{
  g_last_error = msg;
}

int fail(const std::string& msg)
{
  g_last_error = msg;
  return 1;
}

int check_launch(Ctx* c, const char* what)
{
  (void)c;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    return fail(std::string(what) + ": " + cudaGetErrorString(e));
  }
  return 0;
}

int DevBuf::reserve(size_t n)
{
  if (n <= bytes) {
    return 0;
  }
  if (p) {
    cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
  size_t want = (n + n / 8 + 255) & ~size_t(255);
  PSC_CUDA_TRY(cudaMalloc(&p, want));
  bytes = want;
  return 0;
}

void DevBuf::release()
{
  if (p) {
    cudaFree(p);
  }
  p = nullptr;
  bytes = 0;
}

KernelScope::KernelScope(Ctx* ctx, const char* name, int) : c(ctx)
{
  if (!c->opt_profile) {
    return;
  }
  for (size_t i = 0; i < c->prof.size(); i++) {
    if (c->prof[i].name == name || !strcmp(c->prof[i].name, name)) {
      idx = (int)i;
    }
  }
  if (idx < 0) {
    ProfEntry e;
    e.name = name;
    c->prof.push_back(e);
    idx = (int)c->prof.size() - 1;
  }
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  cudaEventRecord(a, c->stream);
}

KernelScope::~KernelScope()
{
  if (idx < 0) {
    return;
  }
  cudaEventRecord(b, c->stream);
  c->prof_pending.push_back({idx, {a, b}});
}

static void prof_collect(Ctx* c)
{
  if (c->prof_pending.empty()) {
    return;
  }
  cudaStreamSynchronize(c->stream);
  for (auto& e : c->prof_pending) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e.second.first, e.second.second);
    c->prof[e.first].ms += ms;
    c->prof[e.first].launches++;
    cudaEventDestroy(e.second.first);
    cudaEventDestroy(e.second.second);
  }
  c->prof_pending.clear();
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no libcuda link)
int field_tile_tensor_map(Ctx* c, int id, int rank, const int* box, TensorMap128* out)
{
  using Encode = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                              const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                              CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static Encode encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    PSC_CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (qres != cudaDriverEntryPointSuccess || !fn) {
      return fail("cuTensorMapEncodeTiled is not available from this driver");
    }
    encode = reinterpret_cast<Encode>(fn);
  }
  const GridDev& G = c->gd;
  const cuuint64_t comps = (cuuint64_t)c->flds[id].n_comps * (cuuint64_t)c->n_slots;
  cuuint64_t dims[4], strides[3];
  cuuint32_t bx[4], es[4] = {1, 1, 1, 1};
  if (rank == 4) {
    dims[0] = G.im[0], dims[1] = G.im[1], dims[2] = G.im[2], dims[3] = comps;
    strides[0] = (cuuint64_t)G.im[0] * 4, strides[1] = (cuuint64_t)G.im[0] * G.im[1] * 4;
    strides[2] = (cuuint64_t)G.fld_len * 4;
  } else if (rank == 3 && G.im[0] == 1) {
    dims[0] = G.im[1], dims[1] = G.im[2], dims[2] = comps;
    strides[0] = (cuuint64_t)G.im[1] * 4, strides[1] = (cuuint64_t)G.fld_len * 4;
  } else {
    return fail("field_tile_tensor_map: unsupported rank");
  }
  for (int d = 0; d < rank; d++) {
    bx[d] = (cuuint32_t)box[d];
  }
  static_assert(sizeof(CUtensorMap) == sizeof(TensorMap128), "");
  CUresult r = encode(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank,
                      c->flds[id].d, dims, strides, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    return fail("cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
  }
  return 0;
}

// neighbour / boundary tables of this rank's patches
int build_patch_tables(Ctx* c)
{
  const GridHost& g = c->g;
  int np = g.n_patches;
  std::vector<pm::PatchBnd> pb(np);
  c->h_nei_patch.assign((size_t)np * 27, -1);
  c->h_nei_slot.assign((size_t)np * 27, -1);
  std::vector<int8_t> add_order((size_t)np * 26, -1);
  // proxies: remote neighbour patches in ascending global index
  std::vector<int> remote;
  for (int p = 0; p < np; p++) {
    int gp = g.patch_begin + p;
    for (int di = 0; di < 27; di++) {
      if (di == 13) {
        continue;
      }
      int dir[3] = {di % 3 - 1, (di / 3) % 3 - 1, di / 9 - 1};
      int ngp = g.neighbor_patch(gp, dir);
      if (ngp >= 0 && g.rank_of_patch(ngp) != g.rank) {
        remote.push_back(ngp);
      }
    }
  }
  std::sort(remote.begin(), remote.end());
  remote.erase(std::unique(remote.begin(), remote.end()), remote.end());
  c->n_slots = np + (int)remote.size();
  c->proxy_gp = remote;
  for (int p = 0; p < np; p++) {
    int gp = g.patch_begin + p;
    pb[p] = make_patch_bnd(g, gp);
    std::vector<std::pair<std::pair<int, int>, int>> order; // ((sender gp, sender dir idx), di)
    for (int di = 0; di < 27; di++) {
      if (di == 13) {
        continue;
      }
      int dir[3] = {di % 3 - 1, (di / 3) % 3 - 1, di / 9 - 1};
      int ngp = g.neighbor_patch(gp, dir);
      if (ngp < 0) {
        continue;
      }
      int r = g.rank_of_patch(ngp);
      if (r == g.rank) {
        c->h_nei_patch[p * 27 + di] = ngp - g.patch_begin;
        c->h_nei_slot[p * 27 + di] = ngp - g.patch_begin;
      } else {
        c->h_nei_patch[p * 27 + di] = -2 - r;
        c->h_nei_slot[p * 27 + di] =
          np + (int)(std::lower_bound(remote.begin(), remote.end(), ngp) - remote.begin());
      }
      order.push_back({{ngp, 26 - di}, di});
    }
    // the reference adds contributions in the order its sequential loop produces them:
    // sender patch ascending, then the sender's direction index ascending
    // (mrc_ddc_multi.c:519-538)
    std::sort(order.begin(), order.end());
    for (size_t o = 0; o < order.size(); o++) {
      add_order[(size_t)p * 26 + o] = (int8_t)order[o].second;
    }
  }
  PSC_CUDA_TRY(cudaMalloc(&c->d_patch_bnd, np * sizeof(pm::PatchBnd)));
  PSC_CUDA_TRY(cudaMalloc(&c->d_nei_patch, (size_t)np * 27 * sizeof(int)));
  PSC_CUDA_TRY(cudaMalloc(&c->d_nei_slot, (size_t)np * 27 * sizeof(int)));
  PSC_CUDA_TRY(cudaMalloc(&c->d_add_order, (size_t)np * 26));
  PSC_CUDA_TRY(cudaMemcpy(c->d_patch_bnd, pb.data(), np * sizeof(pm::PatchBnd),
                          cudaMemcpyHostToDevice));
  PSC_CUDA_TRY(cudaMemcpy(c->d_nei_patch, c->h_nei_patch.data(), (size_t)np * 27 * sizeof(int),
                          cudaMemcpyHostToDevice));
  PSC_CUDA_TRY(cudaMemcpy(c->d_nei_slot, c->h_nei_slot.data(), (size_t)np * 27 * sizeof(int),
                          cudaMemcpyHostToDevice));
  PSC_CUDA_TRY(cudaMemcpy(c->d_add_order, add_order.data(), (size_t)np * 26,
                          cudaMemcpyHostToDevice));
  return 0;
}

static int ctx_create(const psc_b200_grid_desc* desc, Ctx** out)
{
  std::string err;
  Ctx* c = new Ctx;
  if (!grid_setup(*desc, c->g, err)) {
    delete c;
    return fail(err);
  }
  int n_dev = 0;
  cudaError_t e = cudaGetDeviceCount(&n_dev);
  if (e != cudaSuccess || n_dev == 0) {
    delete c;
    return fail(std::string("psc_b200 needs a CUDA device and there is no CPU fallback: ") +
                cudaGetErrorString(e));
  }
  if (desc->device >= 0) {
    PSC_CUDA_TRY(cudaSetDevice(desc->device));
  }
  PSC_CUDA_TRY(cudaGetDevice(&c->device));
  PSC_CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  {
    // the field chain is a string of small kernels: at high priority its CTAs are placed
    // as soon as slots free up instead of behind the pending CTAs of the particle scatter
    int prio_lo = 0, prio_hi = 0;
    PSC_CUDA_TRY(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    PSC_CUDA_TRY(cudaStreamCreateWithPriority(&c->stream2, cudaStreamNonBlocking, prio_hi));
  }
  {
    // (2-D copies may run as kernels: they must not queue behind the scatter's CTAs)
    int prio_lo = 0, prio_hi = 0;
    PSC_CUDA_TRY(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    PSC_CUDA_TRY(cudaStreamCreateWithPriority(&c->stream_io, cudaStreamNonBlocking, prio_hi));
  }
  PSC_CUDA_TRY(cudaEventCreateWithFlags(&c->ev_j_ready, cudaEventDisableTiming));
  PSC_CUDA_TRY(cudaEventCreateWithFlags(&c->ev_flds_done, cudaEventDisableTiming));
  PSC_CUDA_TRY(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
  PSC_CUDA_TRY(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
  PSC_CUDA_TRY(cudaEventCreate(&c->ev_start));
  PSC_CUDA_TRY(cudaEventCreate(&c->ev_stop));

  const GridHost& g = c->g;
  GridDev& G = c->gd;
  for (int d = 0; d < 3; d++) {
    G.ldims[d] = g.ldims[d];
    G.im[d] = g.im[d];
    G.ibn[d] = g.ibn[d];
  }
  G.n_patches = g.n_patches;
  G.n_cells = g.n_cells;
  G.fld_len = g.fld_len;
  G.dim = g.dim;
  G.deposit = g.deposit;
  G.pc = make_push_const(g);

  c->h_off.assign(g.n_patches + 1, 0);
  PSC_CUDA_TRY(cudaMalloc(&c->d_off, (g.n_patches + 1) * sizeof(uint32_t)));
  PSC_CUDA_TRY(cudaMemset(c->d_off, 0, (g.n_patches + 1) * sizeof(uint32_t)));
  size_t nct = (size_t)g.n_cells * g.n_patches;
  PSC_CUDA_TRY(cudaMalloc(&c->d_cell_off, (nct + 1) * sizeof(uint32_t)));
  PSC_CUDA_TRY(cudaMemset(c->d_cell_off, 0, (nct + 1) * sizeof(uint32_t)));
  PSC_CUDA_TRY(cudaMalloc(&c->d_cell_off_alt, (nct + 1) * sizeof(uint32_t)));
  PSC_TRY(build_patch_tables(c));
  int id;
  PSC_TRY(flds_create(c, PSC_B200_NR_FIELDS, &id)); // field 0 = MfieldsState
  if (desc->max_n_prts) {
    PSC_TRY(prts_reserve(c, desc->max_n_prts));
  }
  PSC_CUDA_TRY(cudaStreamSynchronize(c->stream));
  *out = c;
  return 0;
}

static void ctx_destroy(Ctx* c)
{
  if (!c) {
    return;
  }
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  comm_destroy(c);
  for (int b = 0; b < 2; b++) {
    cudaFree(c->xi4[b]);
    cudaFree(c->pxi4[b]);
  }
  gap_release(c);
  if (c->fs_host) {
    cudaFreeHost(c->fs_host);
  }
  if (c->en_host) {
    cudaFreeHost(c->en_host);
  }
  cudaFree(c->d_off);
  cudaFree(c->d_cell_off);
  cudaFree(c->d_cell_off_alt);
  cudaFree(c->d_patch_bnd);
  cudaFree(c->d_nei_patch);
  cudaFree(c->d_nei_slot);
  cudaFree(c->d_add_order);
  for (auto& f : c->flds) {
    cudaFree(f.d);
  }
  for (auto& s : c->scr) {
    s.release();
  }
  c->stage.release();
  for (auto& e : c->prof_pending) {
    cudaEventDestroy(e.second.first);
    cudaEventDestroy(e.second.second);
  }
  cudaEventDestroy(c->ev_start);
  cudaEventDestroy(c->ev_stop);
  cudaEventDestroy(c->ev_fork);
  cudaEventDestroy(c->ev_join);
  cudaEventDestroy(c->ev_j_ready);
  cudaEventDestroy(c->ev_flds_done);
  cudaStreamDestroy(c->stream_io);
  cudaStreamDestroy(c->stream2);
  cudaStreamDestroy(c->stream);
  delete c;
}

static int push_mprts(Ctx* c)
{
  return c->opt_fma ? push_mprts_fast(c) : push_mprts_exact(c);
}

// ---- the particle operators as a deck calls them (psc.hxx:359, :389, :412), with the
// keep_sorted policy: PSC's decks sort every 10th step because sorting pays for itself
// slowly on a CPU; here the tiled push NEEDS the cell order (S3D: 27 ms against 156 ms
// for the unordered store), and exchange + sort cost 14 ms as one fused pass.
static int op_sort(Ctx* c)
{
  if (c->sorted) {
    return 0; // a stable sort of a cell-ordered store is the identity
  }
  return sort_mprts(c);
}

static int op_push(Ctx* c)
{
  if (c->opt_keep_sorted && c->opt_tiled && !c->sorted && c->n_prts) {
    PSC_TRY(sort_mprts(c));
  }
  c->want_counts = c->opt_keep_sorted && c->opt_fused_sort && c->sorted;
  return push_mprts(c);
}

static int op_bnd_particles(Ctx* c)
{
  if (c->opt_keep_sorted && c->opt_fused_sort && c->pushed_from_sorted) {
    return fused_bnd_sort(c);
  }
  return bnd_particles(c);
}

// psc.hxx:417-467 without Marder: J ghosts, then the Yee leapfrog with its ghost fills
static int field_chain(Ctx* c, const psc_b200_step_params* prm)
{
  PSC_TRY(bndf_add_ghosts_J(c));                       // :417
  PSC_TRY(bnd_add_ghosts(c, 0, pm::JXI, pm::JXI + 3)); // :418
  PSC_TRY(bnd_fill_ghosts(c, 0, pm::JXI, pm::JXI + 3)); // :419
  PSC_CUDA_TRY(cudaEventRecord(c->ev_j_ready, c->stream)); // (a queued download of J need not wait for Yee)
  if (prm->push_fields) {
    PSC_TRY(push_H(c, .5)); // :426
    PSC_TRY(bndf_fill_ghosts_H(c));
    PSC_TRY(bnd_fill_ghosts(c, 0, pm::HX, pm::HX + 3));
    PSC_TRY(push_E(c, 1.)); // :439
    PSC_TRY(bndf_fill_ghosts_E(c));
    PSC_TRY(bnd_fill_ghosts(c, 0, pm::EX, pm::EX + 3));
    PSC_TRY(push_H(c, .5)); // :461
    PSC_TRY(bndf_fill_ghosts_H(c));
    PSC_TRY(bnd_fill_ghosts(c, 0, pm::HX, pm::HX + 3));
  }
  return 0;
}

static int en_host_ready(Ctx* c)
{
  if (!c->en_host) {
    PSC_CUDA_TRY(cudaMallocHost(&c->en_host, 8 * sizeof(double)));
  }
  return 0;
}

static int step_core(Ctx* c, const psc_b200_step_params* prm);

// Psc::step (src/include/psc.hxx:321-486) without collisions / injection / output
static int step(Ctx* c, const psc_b200_step_params* prm)
{
  PSC_TRY(step_core(c, prm));
  if (prm->energies) {
    PSC_TRY(en_host_ready(c));
    PSC_TRY(store_ready(c));
    PSC_TRY(field_energies(c, c->en_host));
    PSC_TRY(prts_energies(c, c->en_host + 6));
    c->en_valid = true;
  }
  return 0;
}

static int step_core(Ctx* c, const psc_b200_step_params* prm)
{
  // gapped store (gap.cuh): push + deposit + boundary exchange + sort are one pass over the
  // particles; the store stays gapped from step to step
  const bool do_sort = prm->sort || (c->opt_keep_sorted && c->opt_fused_sort && c->opt_tiled);
  bool gap_ok = do_sort && c->opt_fused_sort && c->opt_gapped && c->opt_tiled && !c->comm &&
                !prm->checks && prm->marder_loop <= 0 && c->n_prts > 0;
  // (pull mode: the push below completes the pending sort itself; the continuity check wants
  // the store before the push)
  if (!gap_ok && !(c->pull_pending && !prm->checks)) {
    PSC_TRY(store_ready(c));
  }
  if (do_sort && !c->sorted && !c->gapped) {
    PSC_TRY(sort_mprts(c)); // psc.hxx:356-361
  }
  if (gap_ok) {
    PSC_TRY(gap_prepare(c, &gap_ok));
  }
  if (gap_ok) {
    bool redo = false;
    PSC_TRY(c->opt_fma ? push_gap_fast(c) : push_gap_exact(c)); // :389 (+ :412, :356 of the next step)
    PSC_TRY(gap_finish(c, &redo));
    if (redo) {
      // nothing was committed (the push never writes the store it reads): eager path
      PSC_TRY(store_ready(c));
      gap_ok = false;
    }
  }
  if (gap_ok) {
    return field_chain(c, prm);
  }
  if (prm->checks) {
    PSC_TRY(check_continuity_begin(c)); // :379-384
  }
  c->want_counts = do_sort && c->opt_fused_sort && c->sorted;
  PSC_TRY(push_mprts(c)); // :389
  if (c->opt_overlap && do_sort && c->opt_fused_sort && c->pushed_from_sorted && !c->comm && !prm->checks &&
      prm->marder_loop <= 0) {
    // The field chain (:417-467) touches only the field arrays, the fused exchange + sort
    // (:412, :356 of the next step) only the particles, and both depend only on the push:
    // they run side by side, the fields on the second stream.  The sort is DRAM-bound and
    // the field kernels are small, so the fields come almost for free.
    PSC_CUDA_TRY(cudaEventRecord(c->ev_fork, c->stream));
    PSC_CUDA_TRY(cudaStreamWaitEvent(c->stream2, c->ev_fork, 0));
    std::swap(c->stream, c->stream2);
    int rc = field_chain(c, prm);
    cudaEventRecord(c->ev_join, c->stream);
    std::swap(c->stream, c->stream2);
    PSC_TRY(rc);
    rc = fused_bnd_sort(c);
    PSC_CUDA_TRY(cudaStreamWaitEvent(c->stream, c->ev_join, 0));
    return rc;
  }
  // :412 bndp_ -- when this step's store was cell-ordered, the exchange is fused with
  // the sort the next step would start with (same result, one pass over the particles)
  if (do_sort && c->opt_fused_sort && c->pushed_from_sorted) {
    PSC_TRY(fused_bnd_sort(c));
  } else {
    PSC_TRY(bnd_particles(c));
  }
  PSC_TRY(bndf_add_ghosts_J(c));                       // :417
  PSC_TRY(bnd_add_ghosts(c, 0, pm::JXI, pm::JXI + 3)); // :418
  PSC_TRY(bnd_fill_ghosts(c, 0, pm::JXI, pm::JXI + 3)); // :419
  if (prm->push_fields) {
    PSC_TRY(push_H(c, .5)); // :426
    PSC_TRY(bndf_fill_ghosts_H(c));
    PSC_TRY(bnd_fill_ghosts(c, 0, pm::HX, pm::HX + 3));
    PSC_TRY(push_E(c, 1.)); // :439
    PSC_TRY(bndf_fill_ghosts_E(c));
    PSC_TRY(bnd_fill_ghosts(c, 0, pm::EX, pm::EX + 3));
    if (prm->marder_loop > 0) {
      PSC_TRY(store_ready(c));
      PSC_TRY(marder(c, prm->marder_diffusion, prm->marder_loop)); // :448-455
    }
    PSC_TRY(push_H(c, .5)); // :461
    PSC_TRY(bndf_fill_ghosts_H(c));
    PSC_TRY(bnd_fill_ghosts(c, 0, pm::HX, pm::HX + 3));
  }
  if (prm->checks) {
    double e;
    PSC_TRY(store_ready(c));
    PSC_TRY(check_continuity_end(c, &e)); // :471-476
    PSC_TRY(check_gauss(c, &e));          // :479-483
  }
  return 0;
}

// ---- pipelined host I/O (psc_b200_step_begin / _step_end)
//
// A deck whose field solver or diagnostics live on the host needs J down and E/B up every
// step.  After the push nothing that touches the particles depends on the fields and vice
// versa, so step_begin() leaves two streams running side by side:
//   stream   the fused boundary exchange + sort (12 of the step's 33 ms at S3D)
//   stream2  J ghosts (+ the Yee half of the step), then whatever the caller queues with
//            mflds_download_async / mflds_upload_async
// and returns.  io_wait() blocks on stream2 only (J is on the host while the sort is still
// running), step_end() waits for the sort, commits it and joins the streams.  Every other
// entry point completes a pending step first.
static int step_pending_finish(Ctx* c)
{
  if (!c->step_pending) {
    return 0;
  }
  c->step_pending = false;
  const uint64_t fb0 = c->n_fused_fallback;
  int rc = fused_bnd_sort_finish(c);
  if (!rc && c->n_fused_fallback != fb0 && c->en_valid) {
    // the scatter's result was discarded (general path taken instead): so are its energies
    rc = prts_energies(c, c->en_host + 6);
  }
  cudaError_t e = cudaEventRecord(c->ev_join, c->stream2);
  if (e == cudaSuccess) {
    e = cudaStreamWaitEvent(c->stream, c->ev_join, 0);
  }
  if (e == cudaSuccess) {
    e = cudaStreamSynchronize(c->stream2);
  }
  if (e == cudaSuccess) {
    e = cudaStreamSynchronize(c->stream_io); // host buffers of the async transfers are free again
  }
  if (rc) {
    return rc;
  }
  if (e != cudaSuccess) {
    return fail(std::string("step_end: ") + cudaGetErrorString(e));
  }
  return 0;
}

static int step_begin(Ctx* c, const psc_b200_step_params* prm)
{
  const bool do_sort = prm->sort || (c->opt_keep_sorted && c->opt_fused_sort && c->opt_tiled);
  const bool split = do_sort && c->opt_fused_sort && c->opt_tiled && !c->opt_gapped && !prm->checks &&
                     prm->marder_loop <= 0 && c->n_prts > 0;
  if (!split) {
    return step(c, prm); // nothing left pending: the async transfers simply follow it
  }
  if (!c->pull_pending) {
    PSC_TRY(store_ready(c));
  }
  if (!c->sorted) {
    PSC_TRY(sort_mprts(c)); // psc.hxx:356-361
  }
  c->want_counts = true;
  PSC_TRY(push_mprts(c)); // :389
  if (!c->pushed_from_sorted) {
    // (the tiled push was not applicable: no counts, general exchange)
    PSC_TRY(bnd_particles(c));
    PSC_TRY(field_chain(c, prm));
    if (prm->energies) {
      PSC_TRY(en_host_ready(c));
      PSC_TRY(field_energies(c, c->en_host));
      PSC_TRY(prts_energies(c, c->en_host + 6));
      c->en_valid = true;
    }
    return 0;
  }
  // fields on stream2 ...
  PSC_CUDA_TRY(cudaEventRecord(c->ev_fork, c->stream));
  PSC_CUDA_TRY(cudaStreamWaitEvent(c->stream2, c->ev_fork, 0));
  if (prm->energies) {
    PSC_TRY(en_host_ready(c));
  }
  std::swap(c->stream, c->stream2);
  int rc = field_chain(c, prm); // :417-467
  if (!rc && prm->energies) {
    rc = field_energies(c, c->en_host, /*sync=*/false); // (before the caller's uploads touch E, B)
  }
  if (!rc && cudaEventRecord(c->ev_flds_done, c->stream) != cudaSuccess) {
    rc = fail("step_begin: cudaEventRecord");
  }
  std::swap(c->stream, c->stream2);
  PSC_TRY(rc);
  // ... particles on stream (:412 + :356 of the next step): enqueued only (multi-rank: up to
  // and including the exchange of the leavers on the host's clock, the scatter enqueued)
  c->step_pending = true;
  const uint32_t nct = (uint32_t)c->gd.n_cells * c->gd.n_patches;
  c->want_scatter_energies = prm->energies != 0;
  c->scatter_energies_done = false;
  rc = fused_bnd_sort(c, /*defer=*/true);
  c->want_scatter_energies = false;
  PSC_TRY(rc);
  if (prm->energies && c->scatter_energies_done) {
    c->en_valid = true; // (the scatter reduced them on its way)
  } else if (prm->energies && (c->fs_pull || c->pull_pending)) {
    // pull mode: no pass over the stayers to ride on
    PSC_TRY(fused_bnd_sort_finish(c));
    PSC_TRY(store_ready(c));
    PSC_TRY(prts_energies(c, c->en_host + 6, false));
    c->en_valid = true;
  } else if (prm->energies) {
    if (c->fs_deferred) {
      // the sorted store is still the "other" buffer and its size is only on the device
      PSC_TRY(prts_energies(c, c->en_host + 6, false, c->d_cell_off_alt + nct, true));
    } else {
      PSC_TRY(prts_energies(c, c->en_host + 6, false));
    }
    c->en_valid = true;
  }
  return 0;
}

// the caller's transfers run on their own stream: a download of J components waits for the J
// ghost sums only (not for the Yee update behind them), anything else for the end of the
// field chain; without a pending step they are ordered behind everything queued so far
static int io_stream_ready(Ctx* c, int id, int mb, int me, bool upload)
{
  if (c->step_pending) {
    const bool j_only = !upload && id == 0 && mb >= pm::JXI && me <= pm::JXI + 3;
    PSC_CUDA_TRY(cudaStreamWaitEvent(c->stream_io, j_only ? c->ev_j_ready : c->ev_flds_done, 0));
  } else {
    PSC_CUDA_TRY(cudaEventRecord(c->ev_fork, c->stream));
    PSC_CUDA_TRY(cudaStreamWaitEvent(c->stream_io, c->ev_fork, 0));
  }
  return 0;
}

} // namespace psc_b200

using namespace psc_b200;

#define CTX(ctx) reinterpret_cast<Ctx*>(ctx)
#define GUARD_(finish, body)                                                             \
  try {                                                                                  \
    if (!ctx) {                                                                          \
      return fail("null context");                                                       \
    }                                                                                    \
    Ctx* c = CTX(ctx);                                                                   \
    cudaSetDevice(c->device);                                                            \
    if (finish) {                                                                        \
      PSC_TRY(step_pending_finish(c));                                                   \
    }                                                                                    \
    body                                                                                 \
  } catch (const std::exception& e) {                                                    \
    return fail(std::string("exception: ") + e.what());                                  \
  } catch (...) {                                                                        \
    return fail("unknown exception");                                                    \
  }
#define GUARD(body) GUARD_(true, body)
#define GUARD_ASYNC(body) GUARD_(false, body)

extern "C" {

const char* psc_b200_last_error(void)
{
  return g_last_error.c_str();
}

const char* psc_b200_version(void)
{
  return "psc_b200 0.1 (sm_100a)";
}

int psc_b200_create(const psc_b200_grid_desc* desc, psc_b200_ctx** ctx)
{
  try {
    if (!desc || !ctx) {
      return fail("null argument");
    }
    Ctx* c = nullptr;
    int rc = ctx_create(desc, &c);
    *ctx = reinterpret_cast<psc_b200_ctx*>(c);
    return rc;
  } catch (const std::exception& e) {
    return fail(std::string("exception: ") + e.what());
  }
}

void psc_b200_destroy(psc_b200_ctx* ctx)
{
  ctx_destroy(CTX(ctx));
}

int psc_b200_sync(psc_b200_ctx* ctx)
{
  GUARD(PSC_CUDA_TRY(cudaStreamSynchronize(c->stream)); return check_launch(c, "sync");)
}

int psc_b200_n_patches(const psc_b200_ctx* ctx)
{
  return ctx ? reinterpret_cast<const Ctx*>(ctx)->g.n_patches : -1;
}

int psc_b200_patch_begin(const psc_b200_ctx* ctx)
{
  return ctx ? reinterpret_cast<const Ctx*>(ctx)->g.patch_begin : -1;
}

int psc_b200_get_ldims(const psc_b200_ctx* ctx, int ldims[3], int ibn[3])
{
  if (!ctx) {
    return fail("null context");
  }
  const Ctx* c = reinterpret_cast<const Ctx*>(ctx);
  for (int d = 0; d < 3; d++) {
    ldims[d] = c->g.ldims[d];
    ibn[d] = c->g.ibn[d];
  }
  return 0;
}

int psc_b200_mprts_set(psc_b200_ctx* ctx, const void* prts, const uint32_t* n_by_patch)
{
  GUARD(c->gapped = false; c->pull_pending = false; return prts_set(c, prts, n_by_patch);)
}

int psc_b200_mprts_inject(psc_b200_ctx* ctx, const void* prts, const uint32_t* n_by_patch)
{
  GUARD(PSC_TRY(store_ready(c)); return prts_inject(c, prts, n_by_patch);)
}

int psc_b200_mprts_size(psc_b200_ctx* ctx, uint64_t* n_total)
{
  GUARD(*n_total = c->n_prts; return 0;)
}

int psc_b200_mprts_size_by_patch(psc_b200_ctx* ctx, uint32_t* n_by_patch)
{
  GUARD(for (int p = 0; p < c->g.n_patches; p++) { n_by_patch[p] = c->h_off[p + 1] - c->h_off[p]; } return 0;)
}

int psc_b200_mprts_get(psc_b200_ctx* ctx, void* prts, uint32_t* off)
{
  GUARD(PSC_TRY(store_ready(c)); return prts_get(c, prts, off);)
}

int psc_b200_mprts_setup_thermal(psc_b200_ctx* ctx, int ppc, const double* vth, uint64_t seed)
{
  GUARD(c->gapped = false; c->pull_pending = false; return prts_setup_thermal(c, ppc, nullptr, vth, seed);)
}

int psc_b200_mprts_setup_thermal_by_patch(psc_b200_ctx* ctx, const int* ppc_by_patch, const double* vth,
                                          uint64_t seed)
{
  GUARD(c->gapped = false; c->pull_pending = false; if (!ppc_by_patch) { return fail("null ppc_by_patch"); }
        return prts_setup_thermal(c, 0, ppc_by_patch, vth, seed);)
}

int psc_b200_mflds_create(psc_b200_ctx* ctx, int n_comps, int* field_id)
{
  GUARD(return flds_create(c, n_comps, field_id);)
}

int psc_b200_mflds_upload(psc_b200_ctx* ctx, int id, int mb, int me, const float* host)
{
  GUARD(return flds_upload(c, id, mb, me, host);)
}

int psc_b200_mflds_download(psc_b200_ctx* ctx, int id, int mb, int me, float* host)
{
  GUARD(return flds_download(c, id, mb, me, host);)
}

int psc_b200_mflds_download_async(psc_b200_ctx* ctx, int id, int mb, int me, float* host)
{
  GUARD_ASYNC(PSC_TRY(io_stream_ready(c, id, mb, me, false)); std::swap(c->stream, c->stream_io);
              int rc = flds_download(c, id, mb, me, host, /*sync=*/false);
              std::swap(c->stream, c->stream_io); return rc;)
}

int psc_b200_mflds_upload_async(psc_b200_ctx* ctx, int id, int mb, int me, const float* host)
{
  GUARD_ASYNC(PSC_TRY(io_stream_ready(c, id, mb, me, true)); std::swap(c->stream, c->stream_io);
              int rc = flds_upload(c, id, mb, me, host, /*sync=*/false);
              std::swap(c->stream, c->stream_io); return rc;)
}

int psc_b200_io_wait(psc_b200_ctx* ctx)
{
  GUARD_ASYNC(PSC_CUDA_TRY(cudaStreamSynchronize(c->stream_io)); return check_launch(c, "io_wait");)
}

int psc_b200_step_begin(psc_b200_ctx* ctx, const psc_b200_step_params* prm)
{
  GUARD(if (!prm) { return fail("null step params"); } return step_begin(c, prm);)
}

int psc_b200_last_energies(psc_b200_ctx* ctx, double out[8])
{
  GUARD(if (!c->en_valid) { return fail("no step has reduced the energies yet (step_params.energies)"); }
        PSC_CUDA_TRY(cudaStreamSynchronize(c->stream)); PSC_CUDA_TRY(cudaStreamSynchronize(c->stream2));
        memcpy(out, c->en_host, 8 * sizeof(double)); return 0;)
}

int psc_b200_step_end(psc_b200_ctx* ctx)
{
  GUARD(return 0;) // (GUARD completes the pending step)
}

int psc_b200_mflds_zero(psc_b200_ctx* ctx, int id, int mb, int me)
{
  GUARD(return flds_zero(c, id, mb, me);)
}

int psc_b200_mflds_fill(psc_b200_ctx* ctx, int id, int m, float value)
{
  GUARD(return flds_fill(c, id, m, value);)
}

int psc_b200_mflds_add(psc_b200_ctx* ctx, int y_id, int y_mb, int x_id, int x_mb, int n_comps)
{
  GUARD(return flds_add(c, y_id, y_mb, x_id, x_mb, n_comps);)
}

int psc_b200_mflds_scale(psc_b200_ctx* ctx, int id, int mb, int me, double a)
{
  GUARD(return flds_scale(c, id, mb, me, a);)
}

int psc_b200_mflds_download_interior(psc_b200_ctx* ctx, int id, int mb, int me, float* host)
{
  GUARD(return flds_download_interior(c, id, mb, me, host);)
}

int psc_b200_push_mprts(psc_b200_ctx* ctx)
{
  GUARD(PSC_TRY(store_ready(c)); return op_push(c);)
}

int psc_b200_sort(psc_b200_ctx* ctx)
{
  GUARD(PSC_TRY(store_ready(c)); return op_sort(c);)
}

int psc_b200_bnd_particles(psc_b200_ctx* ctx)
{
  GUARD(PSC_TRY(store_ready(c)); return op_bnd_particles(c);)
}

int psc_b200_collide(psc_b200_ctx* ctx, const psc_b200_collision_params* prm, uint64_t* n_collisions)
{
  GUARD(PSC_TRY(store_ready(c)); return collide(c, prm, n_collisions);)
}

int psc_b200_heating_spot_foil(psc_b200_ctx* ctx, const psc_b200_heating_params* prm, uint64_t* n_kicked)
{
  GUARD(PSC_TRY(store_ready(c)); return heating_spot_foil(c, prm, n_kicked);)
}

int psc_b200_deposit_j(psc_b200_ctx* ctx, const psc_b200_jpath* paths, uint64_t n)
{
  GUARD(
    if (n == 0) { return 0; }
    if (!paths || n > 0xffffffffu) { return fail("deposit_j: bad trajectory list"); }
    for (uint64_t i = 0; i < n; i++) {
      if (paths[i].patch < 0 || paths[i].patch >= c->g.n_patches) {
        return fail("deposit_j: trajectory " + std::to_string(i) + " names patch " +
                    std::to_string(paths[i].patch) + " of " + std::to_string(c->g.n_patches));
      }
    }
    PSC_TRY(c->scr[0].reserve(n * sizeof(psc_b200_jpath)));
    psc_b200_jpath* d = c->scr[0].as<psc_b200_jpath>();
    // (pageable source: the copy is staged and the call returns with `paths` free again)
    PSC_CUDA_TRY(cudaMemcpyAsync(d, paths, n * sizeof(psc_b200_jpath), cudaMemcpyHostToDevice, c->stream));
    return c->opt_fma ? deposit_paths_fast(c, d, (uint32_t)n) : deposit_paths_exact(c, d, (uint32_t)n);)
}

int psc_b200_checkpoint_write(psc_b200_ctx* ctx, const char* path, int64_t timestep)
{
  GUARD(PSC_TRY(store_ready(c)); return checkpoint_write(c, path, timestep);)
}

int psc_b200_checkpoint_read(psc_b200_ctx* ctx, const char* path, int64_t* timestep)
{
  GUARD(c->pull_pending = false; return checkpoint_read(c, path, timestep);)
}

int psc_b200_bnd_add_ghosts(psc_b200_ctx* ctx, int id, int mb, int me)
{
  GUARD(return bnd_add_ghosts(c, id, mb, me);)
}

int psc_b200_bnd_fill_ghosts(psc_b200_ctx* ctx, int id, int mb, int me)
{
  GUARD(return bnd_fill_ghosts(c, id, mb, me);)
}

int psc_b200_bndf_fill_ghosts_E(psc_b200_ctx* ctx)
{
  GUARD(return bndf_fill_ghosts_E(c);)
}

int psc_b200_bndf_fill_ghosts_H(psc_b200_ctx* ctx)
{
  GUARD(return bndf_fill_ghosts_H(c);)
}

int psc_b200_bndf_add_ghosts_J(psc_b200_ctx* ctx)
{
  GUARD(return bndf_add_ghosts_J(c);)
}

int psc_b200_push_E(psc_b200_ctx* ctx, double dt_fac)
{
  GUARD(return push_E(c, dt_fac);)
}

int psc_b200_push_H(psc_b200_ctx* ctx, double dt_fac)
{
  GUARD(return push_H(c, dt_fac);)
}

int psc_b200_marder(psc_b200_ctx* ctx, double diffusion, int loop)
{
  GUARD(PSC_TRY(store_ready(c)); return marder(c, diffusion, loop);)
}

int psc_b200_moment_n_comps(psc_b200_ctx* ctx, int moment)
{
  return ctx ? moment_n_comps(CTX(ctx), moment) : -1;
}

int psc_b200_moment_1st(psc_b200_ctx* ctx, int field_id, int moment)
{
  GUARD(PSC_TRY(store_ready(c)); return moment_1st(c, field_id, moment);)
}

int psc_b200_moment_rho_1st_nc(psc_b200_ctx* ctx, int field_id)
{
  GUARD(PSC_TRY(store_ready(c)); return moment_rho_1st_nc(c, field_id);)
}

int psc_b200_check_continuity_begin(psc_b200_ctx* ctx)
{
  GUARD(PSC_TRY(store_ready(c)); return check_continuity_begin(c);)
}

int psc_b200_check_continuity_end(psc_b200_ctx* ctx, double* max_err)
{
  GUARD(PSC_TRY(store_ready(c)); return check_continuity_end(c, max_err);)
}

int psc_b200_check_gauss(psc_b200_ctx* ctx, double* max_err)
{
  GUARD(PSC_TRY(store_ready(c)); return check_gauss(c, max_err);)
}

int psc_b200_energies(psc_b200_ctx* ctx, double out[8])
{
  GUARD(PSC_TRY(store_ready(c)); PSC_TRY(field_energies(c, out)); PSC_TRY(prts_energies(c, out + 6));
        if (c->comm) { PSC_TRY(comm_allreduce_sum(c, out, 8)); } return 0;)
}

int psc_b200_step(psc_b200_ctx* ctx, const psc_b200_step_params* prm)
{
  GUARD(return step(c, prm);)
}

int psc_b200_last_checks(psc_b200_ctx* ctx, double* continuity, double* gauss)
{
  GUARD(*continuity = c->last_continuity; *gauss = c->last_gauss; return 0;)
}

int psc_b200_nccl_unique_id(void* id128)
{
  try {
    return comm_unique_id(id128);
  } catch (...) {
    return fail("exception in nccl_unique_id");
  }
}

int psc_b200_nccl_init(psc_b200_ctx* ctx, const void* id128)
{
  GUARD(return comm_init(c, id128);)
}

int psc_b200_balance(psc_b200_ctx* ctx, double factor_fields, int* changed)
{
  GUARD(PSC_TRY(store_ready(c)); return balance(c, factor_fields, changed);)
}

int psc_b200_best_mapping(int n_ranks, const double* capability, int n_patches,
                          const double* loads, int* n_patches_by_rank)
{
  try {
    std::vector<double> cap(capability, capability + n_ranks), ld(loads, loads + n_patches);
    std::vector<int> r = best_mapping(cap, ld);
    for (int i = 0; i < n_ranks; i++) {
      n_patches_by_rank[i] = r[i];
    }
    return 0;
  } catch (const std::exception& e) {
    return fail(std::string("exception: ") + e.what());
  }
}

int psc_b200_selftest_math(psc_b200_ctx* ctx, uint64_t* n_mismatch)
{
  GUARD(return selftest_math(c, n_mismatch);)
}

int psc_b200_set_option(psc_b200_ctx* ctx, const char* name, double value)
{
  GUARD(
    std::string n(name); int v = (int)value;
    if (n == "tiled") { c->opt_tiled = v; }
    else if (n == "warp_reduce") { (void)v; } // accepted for old scripts: the per-cell warp reduction is always on
    else if (n == "fma") { c->opt_fma = v; }
    else if (n == "tma") { c->opt_tma = v; }
    else if (n == "lean") { c->opt_lean = v; }
    else if (n == "push_collect") { c->opt_push_collect = v; }
    else if (n == "pull") { c->opt_pull = v; }
    else if (n == "pull_cap") { c->opt_pull_cap = v; }
    else if (n == "cell_moments") { c->opt_cell_moments = v; }
    else if (n == "vec_fields") { c->opt_vec_fields = v; }
    else if (n == "threads") { c->opt_threads = v; }
    else if (n == "min_blocks") { c->opt_min_blocks = v; }
    else if (n == "tile") { c->opt_tile[0] = c->opt_tile[1] = c->opt_tile[2] = v; }
    else if (n == "tile_x") { c->opt_tile[0] = v; }
    else if (n == "tile_y") { c->opt_tile[1] = v; }
    else if (n == "tile_z") { c->opt_tile[2] = v; }
    else if (n == "profile") { c->opt_profile = v; }
    else if (n == "fused_sort") { c->opt_fused_sort = v; }
    else if (n == "keep_sorted") { c->opt_keep_sorted = v; }
    else if (n == "overlap") { c->opt_overlap = v; }
    else if (n == "gapped") { c->opt_gapped = v; }
    else if (n == "gap_slack") { c->opt_gap_slack = v; }
    else { return fail("unknown option " + n); }
    return 0;)
}

int psc_b200_get_stat(psc_b200_ctx* ctx, const char* name, double* value)
{
  GUARD(
    std::string n(name);
    if (n == "n_launches") { *value = (double)c->n_launches; }
    else if (n == "n_dropped") { *value = (double)c->n_dropped; }
    else if (n == "sorted") { *value = c->sorted; }
    else if (n == "n_prts") { *value = c->n_prts; }
    else if (n == "capacity") { *value = (double)c->cap; }
    else if (n == "n_slots") { *value = c->n_slots; }
    else if (n == "fused_steps") { *value = (double)c->n_fused; }
    else if (n == "lean_pushes") { *value = (double)c->n_lean; }
    else if (n == "gap_steps") { *value = (double)c->n_gap_steps; }
    else if (n == "gap_relayouts") { *value = (double)c->n_gap_relayouts; }
    else if (n == "gap_redone") { *value = (double)c->n_gap_redone; }
    else if (n == "gap_movers") { *value = (double)c->g_mov_used; }
    else if (n == "fused_fallbacks") { *value = (double)c->n_fused_fallback; }
    else if (n == "pull_steps") { *value = (double)c->n_pulled; }
    else if (n == "pull_materialized") { *value = (double)c->n_pull_materialized; }
    else if (n == "pull_overflows") { *value = (double)c->n_pull_overflow; }
    else { return fail("unknown stat " + n); }
    return 0;)
}

int psc_b200_timer_start(psc_b200_ctx* ctx)
{
  GUARD(PSC_CUDA_TRY(cudaEventRecord(c->ev_start, c->stream)); return 0;)
}

int psc_b200_timer_stop(psc_b200_ctx* ctx, float* ms)
{
  GUARD(PSC_CUDA_TRY(cudaEventRecord(c->ev_stop, c->stream));
        PSC_CUDA_TRY(cudaEventSynchronize(c->ev_stop));
        PSC_CUDA_TRY(cudaEventElapsedTime(ms, c->ev_start, c->ev_stop)); return 0;)
}

int psc_b200_prof_get(psc_b200_ctx* ctx, int max, const char** names, float* ms,
                      uint64_t* launches)
{
  if (!ctx) {
    return 0;
  }
  Ctx* c = CTX(ctx);
  prof_collect(c);
  int n = std::min<int>(max, (int)c->prof.size());
  for (int i = 0; i < n; i++) {
    names[i] = c->prof[i].name;
    ms[i] = c->prof[i].ms;
    launches[i] = c->prof[i].launches;
  }
  return n;
}

int psc_b200_prof_reset(psc_b200_ctx* ctx)
{
  GUARD(prof_collect(c); c->prof.clear(); return 0;)
}

} // extern "C"
