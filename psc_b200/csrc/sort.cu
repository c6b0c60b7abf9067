// psc_b200: SortB200 -- stable counting sort of every patch's particles by cell index,
// i.e. exactly the permutation SortCountsort2 produces
// (libpsc/psc_sort/psc_sort_impl.hxx:65-124, cell index = ParticleIndexer::validCellIndex,
// include/particle_indexer.hxx:74-94, which uses float(dx_inv) -- SURVEY.md A.1).
//
// Two device algorithms give that same permutation:
//
//  sort_mprts      any input order.  Keys (patch * n_cells + cell) go through a stable
//                  LSD radix sort (8-bit digits; per-warp-segment digit histograms, one
//                  device-wide scan, rank inside a segment with match.any ballots), then
//                  one gather moves the float4 streams.  Also builds cell_off.
//
//  fused_bnd_sort  the store was cell-ordered before the push (so every particle moved
//                  by at most one cell per direction).  Boundary exchange
//                  (bnd_particles_impl.hxx:93-218, ddc_particles.hxx:421-468) and the
//                  sort of the next step are done in ONE pass over the particles with no
//                  key array at all: per source cell 27 direction counters, a per-target-
//                  cell gather of those counters in the reference's arrival order, one
//                  scan, and a scatter that recomputes each particle's direction.  The
//                  result equals bnd_particles() followed by sort_mprts() bit for bit.
#include "dev_util.cuh"

#include <algorithm>

namespace psc_b200
{

namespace
{

constexpr unsigned FULL = 0xffffffffu;
constexpr int RS_WARPS = 8;
constexpr int RS_ITERS = 64;
constexpr int RS_SEG = 32 * RS_ITERS; // items per warp segment

__global__ void k_cell_keys(GridDev G, uint32_t n, const uint32_t* __restrict__ off,
                            const float4* __restrict__ xi4, uint32_t* __restrict__ keys,
                            uint32_t* __restrict__ cell_cnt, int* __restrict__ err)
{
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) {
    return;
  }
  int p = patch_of(off, G.n_patches, i);
  float4 X = xi4[i];
  float x[3] = {X.x, X.y, X.z};
  int ci = pm::cell_index(G.pc, G.ldims, x);
  if (ci < 0) { // validCellIndex asserts (particle_indexer.hxx:96-101)
    atomicExch(err, 1);
    ci = 0;
  }
  uint32_t key = (uint32_t)p * G.n_cells + ci;
  keys[i] = key;
  atomicAdd(&cell_cnt[key], 1u);
}

__global__ void __launch_bounds__(RS_WARPS * 32)
  k_radix_hist(const uint32_t* __restrict__ keys, size_t n, int shift, uint32_t* __restrict__ hist,
               size_t nseg)
{
  __shared__ uint32_t cnt[RS_WARPS][256];
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int d = lane; d < 256; d += 32) {
    cnt[w][d] = 0;
  }
  __syncwarp();
  size_t seg = (size_t)blockIdx.x * RS_WARPS + w;
  size_t base = seg * RS_SEG;
  for (int it = 0; it < RS_ITERS; it++) {
    size_t i = base + (size_t)it * 32 + lane;
    if (i < n) {
      atomicAdd(&cnt[w][(keys[i] >> shift) & 255u], 1u);
    }
  }
  __syncwarp();
  if (seg < nseg) {
    for (int d = lane; d < 256; d += 32) {
      hist[(size_t)d * nseg + seg] = cnt[w][d];
    }
  }
}

__global__ void __launch_bounds__(RS_WARPS * 32)
  k_radix_scatter(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals, size_t n,
                  int shift, const uint32_t* __restrict__ hist, size_t nseg,
                  uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out)
{
  __shared__ uint32_t cnt[RS_WARPS][256];
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  size_t seg = (size_t)blockIdx.x * RS_WARPS + w;
  if (seg >= nseg) {
    return;
  }
  for (int d = lane; d < 256; d += 32) {
    cnt[w][d] = hist[(size_t)d * nseg + seg];
  }
  __syncwarp();
  size_t base = seg * RS_SEG;
  unsigned lt = (1u << lane) - 1u;
  for (int it = 0; it < RS_ITERS; it++) {
    size_t i = base + (size_t)it * 32 + lane;
    bool valid = i < n;
    unsigned act = __ballot_sync(FULL, valid);
    if (!act) {
      break;
    }
    if (valid) {
      uint32_t key = keys[i];
      uint32_t val = vals ? vals[i] : (uint32_t)i;
      uint32_t d = (key >> shift) & 255u;
      unsigned peers = __match_any_sync(act, d);
      uint32_t b = cnt[w][d];
      __syncwarp(act);
      if (lane == __ffs(peers) - 1) {
        cnt[w][d] = b + __popc(peers);
      }
      __syncwarp(act);
      uint32_t dst = b + __popc(peers & lt);
      keys_out[dst] = key;
      vals_out[dst] = val;
    }
  }
}

__global__ void k_gather(const float4* __restrict__ xi4, const float4* __restrict__ pxi4,
                         const uint32_t* __restrict__ idx, uint32_t n, float4* __restrict__ xo,
                         float4* __restrict__ po)
{
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) {
    return;
  }
  uint32_t i = idx[j];
  xo[j] = xi4[i];
  po[j] = pxi4[i];
}

} // namespace

// stable LSD radix sort of (key, value) pairs; iota_vals: the input values are 0..n-1 and
// `vals` need not be initialised (it is still used as a ping-pong buffer).
// The result is in (keys_alt, vals_alt) if *result_in_alt, else in (keys, vals).
int sort_pairs(Ctx* c, uint32_t* keys, uint32_t* vals, uint32_t* keys_alt, uint32_t* vals_alt,
               size_t n, int key_bits, bool iota_vals, bool* result_in_alt)
{
  *result_in_alt = false;
  if (n == 0) {
    return 0;
  }
  int passes = std::max(1, (key_bits + 7) / 8);
  size_t nseg = (n + RS_SEG - 1) / RS_SEG;
  unsigned nblk = div_up(nseg, RS_WARPS);
  PSC_TRY(c->scr[1].reserve((256 * nseg + 1) * sizeof(uint32_t)));
  uint32_t* hist = c->scr[1].as<uint32_t>();
  uint32_t *ki = keys, *vi = vals, *ko = keys_alt, *vo = vals_alt;
  bool first = true;
  for (int pass = 0; pass < passes; pass++) {
    int shift = 8 * pass;
    k_radix_hist<<<nblk, RS_WARPS * 32, 0, c->stream>>>(ki, n, shift, hist, nseg);
    PSC_TRY(scan_exclusive<uint32_t>(c, LoadArr<uint32_t>{hist}, 256 * nseg, hist, c->scr[2]));
    k_radix_scatter<<<nblk, RS_WARPS * 32, 0, c->stream>>>(ki, first && iota_vals ? nullptr : vi,
                                                          n, shift, hist, nseg, ko, vo);
    c->n_launches += 2;
    std::swap(ki, ko);
    std::swap(vi, vo);
    first = false;
    *result_in_alt = !*result_in_alt;
  }
  return check_launch(c, "sort_pairs");
}

int sort_mprts(Ctx* c)
{
  const GridDev& G = c->gd;
  uint32_t n = c->n_prts;
  size_t nct = (size_t)G.n_cells * G.n_patches;
  PSC_CUDA_TRY(cudaMemsetAsync(c->d_cell_off, 0, (nct + 1) * sizeof(uint32_t), c->stream));
  if (n == 0) {
    c->sorted = true;
  c->pushed_from_sorted = false;
    return 0;
  }
  // scratch: keys, vals and their alternates, error flag
  PSC_TRY(c->scr[3].reserve((4 * (size_t)n + 4) * sizeof(uint32_t)));
  uint32_t* keys = c->scr[3].as<uint32_t>();
  uint32_t* vals = keys + n;
  uint32_t* keys_alt = vals + n;
  uint32_t* vals_alt = keys_alt + n;
  int* err = (int*)(vals_alt + n);
  PSC_CUDA_TRY(cudaMemsetAsync(err, 0, sizeof(int), c->stream));
  {
    KernelScope ks(c, "sort_keys");
    k_cell_keys<<<div_up(n, 256), 256, 0, c->stream>>>(G, n, c->d_off, c->xi(), keys, c->d_cell_off,
                                                      err);
    c->n_launches++;
  }
  int bits = 1;
  while ((size_t(1) << bits) < nct) {
    bits++;
  }
  bool in_alt = false;
  {
    KernelScope ks(c, "sort_radix");
    PSC_TRY(sort_pairs(c, keys, vals, keys_alt, vals_alt, n, bits, true, &in_alt));
  }
  uint32_t* perm = in_alt ? vals_alt : vals;
  {
    KernelScope ks(c, "sort_gather");
    k_gather<<<div_up(n, 256), 256, 0, c->stream>>>(c->xi(), c->pxi(), perm, n, c->xi_alt(),
                                                   c->pxi_alt());
    c->n_launches++;
  }
  c->cur ^= 1;
  {
    KernelScope ks(c, "sort_cell_scan");
    PSC_TRY(scan_exclusive<uint32_t>(c, LoadArr<uint32_t>{c->d_cell_off}, nct, c->d_cell_off,
                                     c->scr[2]));
  }
  int h_err = 0;
  PSC_CUDA_TRY(cudaMemcpyAsync(&h_err, err, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  PSC_CUDA_TRY(cudaStreamSynchronize(c->stream));
  PSC_TRY(check_launch(c, "sort_mprts"));
  if (h_err) {
    c->sorted = false;
  c->pushed_from_sorted = false;
    return fail("sort: particle outside its patch (validCellIndex, particle_indexer.hxx:96-101); "
                "run bnd_particles first");
  }
  c->sorted = true;
  c->pushed_from_sorted = false;
  return 0;
}

} // namespace psc_b200
