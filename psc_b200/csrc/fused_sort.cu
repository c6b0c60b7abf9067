// psc_b200: fused particle boundary exchange + counting sort (the fast path of step()).
//
// Precondition: the store was ordered by (patch, cell) before push_mprts ran, and
// cell_off still describes those pre-push cell runs.  Every particle then sits in the
// run of its *source* cell s and has moved by at most one cell per direction, so its
// target cell differs from s by an offset delta in {-1,0,1}^3 (27 classes), possibly
// through a patch boundary (wrap / neighbour patch / reflection).
//
// The reference result we must reproduce is BndParticles (bnd_particles_impl.hxx:93-218,
// ddc_particles.hxx:421-468) followed by SortCountsort2 (psc_sort_impl.hxx:65-124):
// inside a target cell the particles are ordered
//     [stayers of the patch | arrivals by the receiver's direction loop], each block in
//     the sender's original order = (source cell ascending, index ascending).
// Three kernels and one scan produce exactly that order with no key array and a single
// read+write of the particle data:
//   k_fs_count    one warp per source cell: classify every particle (pm::bnd_classify
//                 arithmetic), count per delta class -> cnt[class][cell]
//   k_fs_offsets  one thread per target cell: walk its <= 27 (route, source cell)
//                 contributions in the reference order, turn cnt into the offset of
//                 each (source cell, delta) group inside the target cell, emit the
//                 target cell's population
//   scan          new cell offsets (these are next step's cell_off and patch offsets)
//   k_fs_scatter  one warp per source cell: re-classify, rank inside the (cell, delta)
//                 group by ballot, write the particle (with the boundary fix-ups) to
//                 new_cell_off[target] + group offset + rank
// If any particle breaks the precondition (moved more than one cell, leaves for another
// rank) a flag is raised and the caller falls back to bnd_particles() + sort_mprts();
// the source store is never modified here.
#include "gap.cuh"

#include <algorithm>

namespace psc_b200
{

namespace
{

constexpr unsigned FULL = 0xffffffffu;
constexpr int FS_WARPS = 8;
constexpr int FS_CELLS = 32; // source cells per CTA

// One CTA = FS_CELLS consecutive source cells, each warp takes FS_CELLS / FS_WARPS of
// them.  Counters are staged in shared memory and written as planes cnt[class][cell] so
// that k_fs_offsets (one thread per target cell) reads and rewrites them coalesced.
__global__ void __launch_bounds__(FS_WARPS * 32)
  k_fs_count(GridDev G, FsTables T, uint32_t nct, const uint32_t* __restrict__ cell_off,
             const float4* __restrict__ xi4, cnt_t* __restrict__ cnt, uint32_t* __restrict__ flags)
{
  __shared__ uint32_t cnt_s[FS_CELLS][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t g0 = blockIdx.x * FS_CELLS;
  for (int k = 0; k < FS_CELLS / FS_WARPS; k++) {
    const int cl = warp * (FS_CELLS / FS_WARPS) + k;
    const uint32_t g = g0 + cl;
    uint32_t mycount = 0;
    if (g < nct) {
      const int p = g / G.n_cells;
      const int s = g - p * G.n_cells;
      const int s0 = s % G.ldims[0], s1 = (s / G.ldims[0]) % G.ldims[1],
                s2 = s / (G.ldims[0] * G.ldims[1]);
      const uint32_t begin = __ldg(&cell_off[g]), end = __ldg(&cell_off[g + 1]);
      for (uint32_t base = begin; base < end; base += 32) {
        uint32_t i = base + lane;
        bool act = i < end;
        int cls = CLS_NONE;
        if (act) {
          float4 X = xi4[i];
          float x[3] = {X.x, X.y, X.z}, u[3] = {0.f, 0.f, 0.f};
          int q, c;
          cls = fs_classify(G, T, p, s0, s1, s2, x, u, q, c);
        }
        unsigned rem = __ballot_sync(FULL, act);
        unsigned grp = __ballot_sync(FULL, cls == CLS_CENTER);
        if (lane == CLS_CENTER) {
          mycount += __popc(grp);
        }
        rem &= ~grp;
        while (rem) {
          int v = __shfl_sync(FULL, cls, __ffs(rem) - 1);
          grp = __ballot_sync(FULL, cls == v);
          if (lane == v) {
            mycount += __popc(grp);
          }
          rem &= ~grp;
        }
      }
      if (mycount) {
        if (lane == CLS_BAD) {
          atomicExch(&flags[0], 1u);
        } else if (lane == CLS_DROP) {
          atomicAdd(&flags[1], mycount);
        } else if (lane == CLS_REMOTE) {
          atomicAdd(&flags[2], mycount);
        }
      }
    }
    cnt_s[cl][lane] = mycount;
  }
  __syncthreads();
  if (g0 + lane < nct) {
    for (int plane = warp; plane < 27; plane += FS_WARPS) {
      cnt[(size_t)plane * nct + g0 + lane] = (cnt_t)cnt_s[lane][plane];
      if (cnt_s[lane][plane] > CNT_MAX) {
        atomicExch(&flags[0], 1u);
      }
    }
  }
}

// contributions of one route (travel direction t, sender patch ps) to target cell c, in
// ascending source-cell order = descending delta.  All counters are loaded before the
// first one is rewritten, so the (up to 27) loads are in flight together instead of
// forming a load -> store -> load chain (the array is updated in place).
template <bool CENTER_INFO = false>
__device__ __forceinline__ void fs_route(const GridDev& G, uint32_t nct, cnt_t* __restrict__ cnt,
                                         int ps, int c0, int c1, int c2, int t0, int t1, int t2,
                                         uint32_t& total, uint32_t* n_lower = nullptr,
                                         uint32_t* n_center = nullptr)
{
  const int ld0 = G.ldims[0], ld1 = G.ldims[1], ld2 = G.ldims[2];
  const size_t pbase = (size_t)ps * G.n_cells;
  uint32_t n[27];
  unsigned ok = 0;
#pragma unroll
  for (int k = 0; k < 27; k++) {
    // k ascending = delta descending
    const int e2 = 1 - k / 9, e1 = 1 - (k / 3) % 3, e0 = 1 - k % 3;
    int z = c2 - e2 + t2 * ld2, y = c1 - e1 + t1 * ld1, x = c0 - e0 + t0 * ld0;
    bool v = !(t2 != 0 && e2 != t2) && (unsigned)z < (unsigned)ld2 && !(t1 != 0 && e1 != t1) &&
             (unsigned)y < (unsigned)ld1 && !(t0 != 0 && e0 != t0) && (unsigned)x < (unsigned)ld0;
    n[k] = 0;
    if (v) {
      ok |= 1u << k;
      n[k] = cnt[(size_t)(((e2 + 1) * 3 + e1 + 1) * 3 + e0 + 1) * nct + pbase +
                 (size_t)((z * ld1 + y) * ld0 + x)];
    }
  }
#pragma unroll
  for (int k = 0; k < 27; k++) {
    const int e2 = 1 - k / 9, e1 = 1 - (k / 3) % 3, e0 = 1 - k % 3;
    if ((ok >> k) & 1) {
      int z = c2 - e2 + t2 * ld2, y = c1 - e1 + t1 * ld1, x = c0 - e0 + t0 * ld0;
      cnt[(size_t)(((e2 + 1) * 3 + e1 + 1) * 3 + e0 + 1) * nct + pbase +
          (size_t)((z * ld1 + y) * ld0 + x)] = (cnt_t)total; // (a target cell beyond CNT_MAX is flagged by the caller)
      if (CENTER_INFO && k == 13) {
        *n_lower = total;
        *n_center = n[k];
      }
      total += n[k];
    }
  }
}

// fs_route for the same-patch route (t = 0) of k_fs_offsets_same: every thread runs it, so
// it is written for registers -- one base pointer, validity from six per-direction tests,
// and the 27 loaded counts kept two per register (they are 16-bit): the general routine
// above spilled 176 bytes per thread at the kernel's 64 registers
template <bool CENTER_INFO, typename IDX>
__device__ __forceinline__ void fs_route_same(const GridDev& G, uint32_t nct, cnt_t* __restrict__ cnt, int q,
                                              int c0, int c1, int c2, uint32_t& total, uint32_t* n_lower,
                                              uint32_t* n_center)
{
  static_assert(sizeof(cnt_t) == 2, "two counts per register");
  const int ld0 = G.ldims[0], ld1 = G.ldims[1], ld2 = G.ldims[2];
  // element index of the target cell + a warp-uniform offset per plane (IDX: 32 bits while
  // 27 nct < 2^32, chosen by the host), so that no 64-bit address is kept
  const IDX s1 = ld0, s2 = (IDX)ld0 * ld1;
  IDX b = (IDX)q * G.n_cells + (IDX)((c2 * ld1 + c1) * ld0 + c0);
  // source coordinate c_d - e_d inside the patch, for e_d = +1 (lo) and -1 (hi)
  const bool lo0 = c0 >= 1, hi0 = c0 + 1 < ld0, lo1 = c1 >= 1, hi1 = c1 + 1 < ld1, lo2 = c2 >= 1, hi2 = c2 + 1 < ld2;
  auto valid = [&](int e0, int e1, int e2) {
    return (e0 == 1 ? lo0 : (e0 == -1 ? hi0 : true)) && (e1 == 1 ? lo1 : (e1 == -1 ? hi1 : true)) &&
           (e2 == 1 ? lo2 : (e2 == -1 ? hi2 : true));
  };
  uint32_t np[14];
#pragma unroll
  for (int k = 0; k < 27; k++) {
    // k ascending = delta descending
    const int e2 = 1 - k / 9, e1 = 1 - (k / 3) % 3, e0 = 1 - k % 3;
    uint32_t v = 0;
    if (valid(e0, e1, e2)) {
      v = cnt[b + ((IDX)(((e2 + 1) * 3 + e1 + 1) * 3 + e0 + 1) * nct - (e2 * s2 + e1 * s1 + e0))];
    }
    if (k & 1) {
      np[k >> 1] |= v << 16;
    } else {
      np[k >> 1] = v;
    }
  }
  // (the store addresses are re-derived from b: kept from the loads they would be 54 registers)
  if constexpr (sizeof(IDX) == 4) {
    asm volatile("" : "+r"(b));
  } else {
    asm volatile("" : "+l"(b));
  }
#pragma unroll
  for (int k = 0; k < 27; k++) {
    const int e2 = 1 - k / 9, e1 = 1 - (k / 3) % 3, e0 = 1 - k % 3;
    if (valid(e0, e1, e2)) {
      const uint32_t n = (k & 1) ? np[k >> 1] >> 16 : np[k >> 1] & 0xffffu;
      // (a target cell beyond CNT_MAX is flagged by the caller)
      cnt[b + ((IDX)(((e2 + 1) * 3 + e1 + 1) * 3 + e0 + 1) * nct - (e2 * s2 + e1 * s1 + e0))] = (cnt_t)total;
      if (CENTER_INFO && k == 13) {
        *n_lower = total;
        *n_center = n;
      }
      total += n;
    }
  }
}

// per target cell: offsets of its contributions in the reference's order; cnt is
// rewritten in place (every (source cell, delta) entry has exactly one target).
// Class order: stayers first, then the receiver's direction loop dir' ascending; the
// sender sits in direction dir' from q and travelled in direction t = -dir'.  A particle
// that crossed the lower (upper) patch face lands in the first (last) cell, so only
// cells on a patch face have routes other than "stay": k_fs_offsets_same handles the
// same-patch contributions of every cell (regular, one thread per cell), and
// k_fs_offsets_face appends the neighbour-patch routes for the cells on the patch faces
// only (one thread per face cell; the face slabs are enumerated z, y, x and a cell that
// lies in several is taken by the first).
// STAY (pull mode): also emit, per target cell, {arrivals placed in front of its stayers, stayers}
#ifndef OFFS_MINB
#define OFFS_MINB 4 // CTAs per SM (56 registers, none spilled)
#endif
template <bool STAY, typename IDX>
__global__ void __launch_bounds__(256, OFFS_MINB)
  k_fs_offsets_same(GridDev G, uint32_t nct, cnt_t* __restrict__ cnt, uint32_t* __restrict__ new_cnt,
                    uint32_t* __restrict__ flags, uint2* __restrict__ stay)
{
  uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= nct) {
    return;
  }
  const int q = g / G.n_cells;
  const int c = g - q * G.n_cells;
  const int ld0 = G.ldims[0], ld1 = G.ldims[1];
  const int c0 = c % ld0, c1 = (c / ld0) % ld1, c2 = c / (ld0 * ld1);
  uint32_t total = 0;
  if constexpr (STAY) {
    uint32_t nl = 0, nc = 0;
    fs_route_same<true, IDX>(G, nct, cnt, q, c0, c1, c2, total, &nl, &nc);
    stay[g] = make_uint2(nl, nc);
  } else {
    fs_route_same<false, IDX>(G, nct, cnt, q, c0, c1, c2, total, nullptr, nullptr);
  }
  new_cnt[g] = total;
  if (total > CNT_MAX) {
    atomicExch(&flags[0], 1u); // offsets inside this cell do not fit the 16-bit planes
  }
}

__device__ __forceinline__ void fs_face_routes(const GridDev& G, const int* __restrict__ nei_patch, uint32_t nct,
                                               cnt_t* __restrict__ cnt, int q, int c0, int c1, int c2,
                                               uint32_t& total)
{
  const int ld0 = G.ldims[0], ld1 = G.ldims[1], ld2 = G.ldims[2];
  const bool lo0 = c0 == 0, hi0 = c0 == ld0 - 1, lo1 = c1 == 0, hi1 = c1 == ld1 - 1, lo2 = c2 == 0,
             hi2 = c2 == ld2 - 1;
  // dir' ascending <=> t descending, z slowest
  for (int t2 = 1; t2 >= -1; t2--) {
    if ((t2 == 1 && !lo2) || (t2 == -1 && !hi2)) {
      continue;
    }
    for (int t1 = 1; t1 >= -1; t1--) {
      if ((t1 == 1 && !lo1) || (t1 == -1 && !hi1)) {
        continue;
      }
      for (int t0 = 1; t0 >= -1; t0--) {
        if ((t0 == 1 && !lo0) || (t0 == -1 && !hi0) || (t0 == 0 && t1 == 0 && t2 == 0)) {
          continue;
        }
        int dip = ((-t2 + 1) * 3 + (-t1 + 1)) * 3 + (-t0 + 1);
        int ps = nei_patch[q * 27 + dip];
        if (ps >= 0) {
          fs_route(G, nct, cnt, ps, c0, c1, c2, t0, t1, t2, total);
        }
      }
    }
  }
}

// face slabs of one patch: areas (0 for a direction without neighbours to look at)
struct FaceGeom
{
  int a[3];  // cells of one slab normal to d
  int s[3];  // slabs normal to d (2, or 1 when the patch is one cell thick, or 0)
  int per_patch;
};

__global__ void __launch_bounds__(128)
  k_fs_offsets_face(GridDev G, FaceGeom FG, const int* __restrict__ nei_patch, uint32_t nct,
                    cnt_t* __restrict__ cnt, uint32_t* __restrict__ new_cnt, uint32_t* __restrict__ flags)
{
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const int q = t / FG.per_patch;
  if (q >= G.n_patches) {
    return;
  }
  int r = t - q * FG.per_patch;
  const int ld0 = G.ldims[0], ld1 = G.ldims[1], ld2 = G.ldims[2];
  const bool f0 = FG.s[0] > 0, f1 = FG.s[1] > 0; // directions whose faces are enumerated
  int c0, c1, c2;
  if (r < FG.s[2] * FG.a[2]) {
    const int side = r / FG.a[2], rr = r - side * FG.a[2];
    c2 = side ? ld2 - 1 : 0, c1 = rr / ld0, c0 = rr - c1 * ld0;
  } else if ((r -= FG.s[2] * FG.a[2]) < FG.s[1] * FG.a[1]) {
    const int side = r / FG.a[1], rr = r - side * FG.a[1];
    c1 = side ? ld1 - 1 : 0, c2 = rr / ld0, c0 = rr - c2 * ld0;
    if (FG.s[2] && (c2 == 0 || c2 == ld2 - 1)) {
      return; // taken by a z slab
    }
  } else {
    r -= FG.s[1] * FG.a[1];
    const int side = r / FG.a[0], rr = r - side * FG.a[0];
    c0 = side ? ld0 - 1 : 0, c2 = rr / ld1, c1 = rr - c2 * ld1;
    if ((FG.s[2] && (c2 == 0 || c2 == ld2 - 1)) || (f1 && (c1 == 0 || c1 == ld1 - 1))) {
      return;
    }
  }
  (void)f0;
  const uint32_t g = (uint32_t)q * G.n_cells + (uint32_t)((c2 * ld1 + c1) * ld0 + c0);
  uint32_t total = new_cnt[g];
  fs_face_routes(G, nei_patch, nct, cnt, q, c0, c1, c2, total);
  new_cnt[g] = total;
  if (total > CNT_MAX) {
    atomicExch(&flags[0], 1u);
  }
}

// Scatter: one warp walks SC_CPW consecutive source cells.  Like the push kernel it
// issues full 32-particle chunks regardless of cell boundaries (coalesced 128-bit loads,
// the next chunk in flight through cp.async while this one is ranked) and visits a chunk
// once per cell it touches; lane k carries the running offset of class k of the current
// cell inside its target cell.
constexpr int SC_WARPS = 8;
constexpr int SC_CPW = 8;                 // cells per warp
constexpr int SC_CELLS = SC_WARPS * SC_CPW; // source cells per CTA

__device__ __forceinline__ void fs_cp_async16(uint32_t dst, const void* src)
{
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

#ifndef SC_MINB
#define SC_MINB 4
#endif
#ifndef SC_DEPTH
#define SC_DEPTH 3 // stages of the per-warp chunk pipeline
#endif
// DiagEnergiesParticle.h:15-40 for the particles the scatter keeps (EN): the pass is
// DRAM-bound with issue slots to spare, a separate reduction pass costs 5.5 ms at S3D
struct ScatterEnergies
{
  float q[pm::MAX_KINDS], m[pm::MAX_KINDS];
  double fnqs_fac; // fnqs * dx * dy * dz
  double* out2;    // [electrons (q < 0), ions (q > 0)]
};

template <bool EN>
__global__ void __launch_bounds__(SC_WARPS * 32, SC_MINB)
  k_fs_scatter(GridDev G, FsTables T, uint32_t nct, const uint32_t* __restrict__ cell_off,
               const uint32_t* __restrict__ new_cell_off, const cnt_t* __restrict__ pre,
               const float4* __restrict__ xi4, const float4* __restrict__ pxi4,
               float4* __restrict__ xo, float4* __restrict__ po, ScatterEnergies E)
{
  [[maybe_unused]] double e_neg = 0., e_pos = 0.;
  __shared__ uint32_t pre_s[SC_CELLS][33];
  // SC_DEPTH - 1 chunks per warp are in flight while one is ranked: the pass is bound by
  // memory latency x bytes in flight (profiles/r01_v5_fs_scatter_ncu.txt: long-scoreboard
  // stalls, 46 % occupancy), not by issue slots
  __shared__ float4 stage_s[SC_WARPS][SC_DEPTH][2][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned lt = (1u << lane) - 1u;
  const uint32_t g0 = blockIdx.x * SC_CELLS;
  const uint32_t gw = g0 + warp * SC_CPW; // first cell of this warp
  const int n_cells_w = gw < nct ? (int)min((uint32_t)SC_CPW, nct - gw) : 0;
  // the warp's first chunk is requested before anything else waits on memory
  const uint32_t myoff = n_cells_w ? __ldg(&cell_off[gw + min(lane, n_cells_w)]) : 0u;
  const uint32_t begin = __shfl_sync(FULL, myoff, 0), end = __shfl_sync(FULL, myoff, n_cells_w);
  const uint32_t stg0 = (uint32_t)__cvta_generic_to_shared(&stage_s[warp][0][0][lane]);
  constexpr uint32_t STAGE_BYTES = 2 * 32 * sizeof(float4);
#pragma unroll
  for (int d = 0; d < SC_DEPTH - 1; d++) {
    const uint32_t i0 = begin + 32 * d + lane;
    if (i0 < end) {
      fs_cp_async16(stg0 + d * STAGE_BYTES, xi4 + i0);
      fs_cp_async16(stg0 + d * STAGE_BYTES + 32 * sizeof(float4), pxi4 + i0);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  for (int plane = warp; plane < 32; plane += SC_WARPS) {
    for (int c = lane; c < SC_CELLS; c += 32) {
      pre_s[c][plane] = (plane < 27 && g0 + c < nct) ? __ldg(&pre[(size_t)plane * nct + g0 + c]) : 0u;
    }
  }
  __syncthreads();
  if (n_cells_w == 0) {
    return;
  }
  // coordinates of the current cell, advanced incrementally
  int p = gw / G.n_cells;
  int s = gw - p * G.n_cells;
  int s0 = s % G.ldims[0], s1 = (s / G.ldims[0]) % G.ldims[1], s2 = s / (G.ldims[0] * G.ldims[1]);
  int cur = 0;
  uint32_t cb = begin, ce = __shfl_sync(FULL, myoff, 1);
  uint32_t run = pre_s[warp * SC_CPW][lane];
  int st = 0; // stage that holds the chunk at `base`
  for (uint32_t base = begin; base < end; base += 32) {
    const uint32_t i = base + lane;
    const bool act = i < end;
    asm volatile("cp.async.wait_group %0;" ::"n"(SC_DEPTH - 2) : "memory");
    float4 X = stage_s[warp][st][0][lane], U = stage_s[warp][st][1][lane];
    {
      // the stage read SC_DEPTH - 1 iterations from now (each lane refills only the slots
      // it reads itself: no warp barrier needed)
      const int sf = st == 0 ? SC_DEPTH - 1 : st - 1;
      const uint32_t inext = i + 32 * (SC_DEPTH - 1);
      if (inext < end) {
        fs_cp_async16(stg0 + sf * STAGE_BYTES, xi4 + inext);
        fs_cp_async16(stg0 + sf * STAGE_BYTES + 32 * sizeof(float4), pxi4 + inext);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    st = st + 1 == SC_DEPTH ? 0 : st + 1;
    for (;;) {
      const bool mine = act && i >= cb && i < ce;
      int cls = CLS_NONE;
      uint32_t tbase = 0;
      if (mine) {
        float x[3] = {X.x, X.y, X.z}, u[3] = {U.x, U.y, U.z};
        int q = 0, c = 0;
        cls = fs_classify(G, T, p, s0, s1, s2, x, u, q, c);
        if (cls < 27) {
          tbase = __ldg(&new_cell_off[(size_t)q * G.n_cells + c]);
          X = make_float4(x[0], x[1], x[2], X.w);
          U = make_float4(u[0], u[1], u[2], U.w);
        }
      }
      unsigned rem = __ballot_sync(FULL, mine);
      uint32_t dst = 0;
      int v = CLS_CENTER;
      while (rem) {
        const unsigned grp = __ballot_sync(FULL, cls == v);
        const uint32_t r0 = __shfl_sync(FULL, run, v);
        if (cls == v) {
          dst = tbase + r0 + __popc(grp & lt);
        }
        if (lane == v) {
          run += __popc(grp);
        }
        rem &= ~grp;
        if (rem) {
          v = __shfl_sync(FULL, cls, __ffs(rem) - 1);
        }
      }
      if (mine && cls < 27) {
        xo[dst] = X;
        po[dst] = U;
        if constexpr (EN) {
          const int kind = __float_as_int(X.w);
          const float qf = E.q[kind], mf = E.m[kind];
          const float w = U.w / qf;
          const double gamma = sqrtf(1.f + U.x * U.x + U.y * U.y + U.z * U.z);
          const double ekin = (gamma - 1.) * mf * w;
          if (qf < 0.f) {
            e_neg += ekin;
          } else if (qf > 0.f) {
            e_pos += ekin;
          }
        }
      }
      if (ce > base + 32) {
        break; // the cell continues in the next chunk
      }
      if (++cur == n_cells_w) {
        break;
      }
      cb = ce;
      ce = __shfl_sync(FULL, myoff, cur + 1);
      run = pre_s[warp * SC_CPW + cur][lane];
      if (++s0 == G.ldims[0]) {
        s0 = 0;
        if (++s1 == G.ldims[1]) {
          s1 = 0;
          if (++s2 == G.ldims[2]) {
            s2 = 0;
            p++;
          }
        }
      }
    }
  }
  if constexpr (EN) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      e_neg += __shfl_xor_sync(FULL, e_neg, o);
      e_pos += __shfl_xor_sync(FULL, e_pos, o);
    }
    if (lane == 0) {
      atomicAdd(&E.out2[0], e_neg * E.fnqs_fac);
      atomicAdd(&E.out2[1], e_pos * E.fnqs_fac);
    }
  }
}

// ---- pull mode (push_lean_pull.cuh): only the particles that change cell are moved here;
// the stayers cross over inside the next push.

// one thread per entry of the push's mover list: the particle at mv_idx goes to
// new_cell_off[target] + offset of its (source cell, class) group + its rank in the group,
// with the boundary fix-ups -- the same place k_fs_scatter would put it
__global__ void __launch_bounds__(256)
  k_fs_place_movers(GridDev G, FsTables T, uint32_t nct, const uint32_t* __restrict__ flags, uint32_t cap,
                    uint32_t n_host, const uint32_t* __restrict__ mv_idx, const uint2* __restrict__ mv_key,
                    const cnt_t* __restrict__ pre, const uint32_t* __restrict__ new_cell_off,
                    const float4* __restrict__ xi4, const float4* __restrict__ pxi4, float4* __restrict__ xo,
                    float4* __restrict__ po)
{
  const uint32_t n = n_host != 0xffffffffu ? n_host : min(flags[3], cap);
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    const uint32_t i = mv_idx[k];
    const uint2 key = mv_key[k];
    const uint32_t cls = key.x / nct, g = key.x - cls * nct;
    const int p = g / G.n_cells;
    const int s = g - p * G.n_cells;
    const int s0 = s % G.ldims[0], s1 = (s / G.ldims[0]) % G.ldims[1], s2 = s / (G.ldims[0] * G.ldims[1]);
    const float4 X = xi4[i], U = pxi4[i];
    float x[3] = {X.x, X.y, X.z}, u[3] = {U.x, U.y, U.z};
    int q = 0, c = 0;
    if (fs_classify(G, T, p, s0, s1, s2, x, u, q, c) != (int)cls) {
      continue; // (cannot happen: the push classified the same record with the same arithmetic)
    }
    const uint32_t dst = __ldg(&new_cell_off[(size_t)q * G.n_cells + c]) + pre[key.x] + key.y;
    xo[dst] = make_float4(x[0], x[1], x[2], X.w);
    po[dst] = make_float4(u[0], u[1], u[2], U.w);
  }
}

// the rest of the sort, when something other than the next push wants the store: one warp
// per source cell copies the records that still live in the cell to their place
__global__ void __launch_bounds__(FS_WARPS * 32)
  k_pull_stayers(GridDev G, uint32_t nct, const uint32_t* __restrict__ cell_off,
                 const uint32_t* __restrict__ new_cell_off, const uint2* __restrict__ stay,
                 const float4* __restrict__ xi4, const float4* __restrict__ pxi4, float4* __restrict__ xo,
                 float4* __restrict__ po)
{
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  const uint32_t g = blockIdx.x * FS_WARPS + (threadIdx.x >> 5);
  if (g >= nct) {
    return;
  }
  const int p = g / G.n_cells;
  const int s = g - p * G.n_cells;
  const int s0 = s % G.ldims[0], s1 = (s / G.ldims[0]) % G.ldims[1], s2 = s / (G.ldims[0] * G.ldims[1]);
  const uint32_t begin = __ldg(&cell_off[g]), end = __ldg(&cell_off[g + 1]);
  uint32_t dst = __ldg(&new_cell_off[g]) + stay[g].x;
  for (uint32_t base = begin; base < end; base += 32) {
    const uint32_t i = base + lane;
    bool live = false;
    float4 X, U;
    if (i < end) {
      X = xi4[i], U = pxi4[i];
      live = pm::cell_position(G.pc, X.x, 0) == s0 && pm::cell_position(G.pc, X.y, 1) == s1 &&
             pm::cell_position(G.pc, X.z, 2) == s2;
    }
    const unsigned lm = __ballot_sync(FULL, live);
    if (live) {
      xo[dst + __popc(lm & lt)] = X;
      po[dst + __popc(lm & lt)] = U;
    }
    dst += __popc(lm);
  }
}

// entering pull mode from a cell-ordered store: nothing in front of the stayers, everybody stays
__global__ void k_stay_init(uint32_t nct, const uint32_t* __restrict__ cell_off, uint2* __restrict__ stay)
{
  const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g < nct) {
    stay[g] = make_uint2(0u, cell_off[g + 1] - cell_off[g]);
  }
}

__global__ void k_patch_offsets(const uint32_t* __restrict__ cell_off, int n_patches, int n_cells,
                                uint32_t* __restrict__ off)
{
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p <= n_patches) {
    off[p] = cell_off[(size_t)p * n_cells];
  }
}

// ---- multi-rank: particles leaving for another rank (CLS_REMOTE) can only start in
// cells on a patch face that touches a remote patch.  One warp per such cell appends
// them to a list (group key = (rank * n_patches + patch) * 32 + direction, index); the
// list is then sorted by (group, index) = the reference's send order
// (ddc_particles.hxx:360-409).
__global__ void __launch_bounds__(FS_WARPS * 32)
  k_fs_collect_remote(GridDev G, FsTables T, uint32_t n_rf, const uint32_t* __restrict__ rf_cells,
                      const uint32_t* __restrict__ cell_off, const float4* __restrict__ xi4,
                      uint32_t* __restrict__ rkey, uint32_t* __restrict__ ridx,
                      uint32_t* __restrict__ counter, uint32_t cap)
{
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  const uint32_t w = blockIdx.x * FS_WARPS + (threadIdx.x >> 5);
  if (w >= n_rf) {
    return;
  }
  const uint32_t g = rf_cells[w];
  const int p = g / G.n_cells;
  const int s = g - p * G.n_cells;
  const int s0 = s % G.ldims[0], s1 = (s / G.ldims[0]) % G.ldims[1], s2 = s / (G.ldims[0] * G.ldims[1]);
  const uint32_t begin = __ldg(&cell_off[g]), end = __ldg(&cell_off[g + 1]);
  for (uint32_t base = begin; base < end; base += 32) {
    uint32_t i = base + lane;
    bool rem = false;
    uint32_t key = 0;
    if (i < end) {
      float4 X = xi4[i];
      float x[3] = {X.x, X.y, X.z}, u[3] = {0.f, 0.f, 0.f};
      int q, c;
      if (fs_classify(G, T, p, s0, s1, s2, x, u, q, c) == CLS_REMOTE) {
        rem = true;
        key = ((uint32_t)(-2 - q) * G.n_patches + p) * 32u + (uint32_t)c;
      }
    }
    unsigned m = __ballot_sync(FULL, rem);
    if (m) {
      uint32_t b = 0;
      if (lane == __ffs(m) - 1) {
        b = atomicAdd(counter, (uint32_t)__popc(m));
      }
      b = __shfl_sync(FULL, b, __ffs(m) - 1);
      if (rem) {
        uint32_t slot = b + __popc(m & lt);
        if (slot < cap) {
          rkey[slot] = key;
          ridx[slot] = i;
        }
      }
    }
  }
}

// received particles: target cell key, and one more particle for that cell
__global__ void k_fs_recv_keys(GridDev G, uint32_t n, const uint32_t* __restrict__ recv_off,
                               const float4* __restrict__ xr, uint32_t* __restrict__ keys,
                               uint32_t* __restrict__ new_cnt, uint32_t* __restrict__ flags)
{
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) {
    return;
  }
  int q = patch_of(recv_off, G.n_patches, j);
  float4 X = xr[j];
  float x[3] = {X.x, X.y, X.z};
  int ci = pm::cell_index(G.pc, G.ldims, x);
  if (ci < 0) {
    atomicExch(&flags[0], 1u);
    ci = 0;
  }
  uint32_t key = (uint32_t)q * G.n_cells + ci;
  keys[j] = key;
  atomicAdd(&new_cnt[key], 1u);
}

// received particles fill the tail of their target cell in list order
// (ddc_particles.hxx:456-468: behind every local arrival)
__global__ void k_fs_place_remote(uint32_t n, const uint32_t* __restrict__ skey,
                                  const uint32_t* __restrict__ sidx,
                                  const uint32_t* __restrict__ new_cell_off,
                                  const float4* __restrict__ xr, const float4* __restrict__ pr,
                                  float4* __restrict__ xo, float4* __restrict__ po)
{
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) {
    return;
  }
  uint32_t key = skey[j];
  // first index behind this key's run
  uint32_t lo = j, hi = n; // skey[lo] == key, skey[hi] > key (hi = n: none)
  while (hi - lo > 1) {
    uint32_t mid = (lo + hi) >> 1;
    if (skey[mid] == key) {
      lo = mid;
    } else {
      hi = mid;
    }
  }
  uint32_t dst = new_cell_off[key + 1] - (hi - j);
  uint32_t r = sidx[j];
  xo[dst] = xr[r];
  po[dst] = pr[r];
}

} // namespace

// cells from which a particle can leave for another rank
static int build_remote_cells(Ctx* c)
{
  const GridHost& g = c->g;
  std::vector<uint32_t> cells;
  for (int p = 0; p < g.n_patches; p++) {
    for (int di = 0; di < 27; di++) {
      if (di == 13 || c->h_nei_patch[p * 27 + di] >= -1) {
        continue;
      }
      int dir[3] = {di % 3 - 1, (di / 3) % 3 - 1, di / 9 - 1};
      int lo[3], hi[3];
      for (int d = 0; d < 3; d++) {
        lo[d] = dir[d] > 0 ? g.ldims[d] - 1 : 0;
        hi[d] = dir[d] < 0 ? 1 : g.ldims[d];
      }
      for (int k = lo[2]; k < hi[2]; k++) {
        for (int j = lo[1]; j < hi[1]; j++) {
          for (int i = lo[0]; i < hi[0]; i++) {
            cells.push_back((uint32_t)p * g.n_cells + (uint32_t)((k * g.ldims[1] + j) * g.ldims[0] + i));
          }
        }
      }
    }
  }
  std::sort(cells.begin(), cells.end());
  cells.erase(std::unique(cells.begin(), cells.end()), cells.end());
  c->n_rf_cells = (uint32_t)cells.size();
  PSC_TRY(c->rf_cells.reserve(std::max<size_t>(cells.size(), 4) * sizeof(uint32_t)));
  PSC_CUDA_TRY(cudaMemcpy(c->rf_cells.p, cells.data(), cells.size() * sizeof(uint32_t),
                          cudaMemcpyHostToDevice));
  c->rf_built = true;
  return 0;
}

static int fused_fallback(Ctx* c)
{
  // precondition broken for some particle: the source store is intact, take the general path
  c->n_fused_fallback++;
  c->counts_valid = false;
  c->fs_pull = false; // (a pull push leaves a complete pushed store like any other)
  PSC_TRY(bnd_particles(c));
  return sort_mprts(c);
}

// the scatter pass of fused_bnd_sort (count planes turned into offsets, new cell offsets scanned)
static int launch_scatter(Ctx* c, bool energies)
{
  const GridDev& G = c->gd;
  const uint32_t nct = (uint32_t)G.n_cells * G.n_patches;
  const cnt_t* cnt = c->scr[9].as<cnt_t>();
  FsTables T{c->d_patch_bnd, c->d_nei_patch};
  ScatterEnergies SE{};
  if (energies) {
    // step_begin asked for DiagEnergies: the particle part rides on this pass
    for (int k = 0; k < c->g.desc.n_kinds; k++) {
      SE.q[k] = (float)c->g.desc.q[k];
      SE.m[k] = (float)c->g.desc.m[k];
    }
    SE.fnqs_fac = c->g.desc.fnqs * c->g.dx[0] * c->g.dx[1] * c->g.dx[2];
    PSC_TRY(c->scr[0].reserve(2 * sizeof(double)));
    SE.out2 = c->scr[0].as<double>();
    PSC_CUDA_TRY(cudaMemsetAsync(SE.out2, 0, 2 * sizeof(double), c->stream));
    k_fs_scatter<true><<<div_up(nct, SC_CELLS), SC_WARPS * 32, 0, c->stream>>>(
      G, T, nct, c->d_cell_off, c->d_cell_off_alt, cnt, c->xi(), c->pxi(), c->xi_alt(), c->pxi_alt(), SE);
    PSC_CUDA_TRY(cudaMemcpyAsync(c->en_host + 6, SE.out2, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    c->scatter_energies_done = true;
  } else {
    k_fs_scatter<false><<<div_up(nct, SC_CELLS), SC_WARPS * 32, 0, c->stream>>>(
      G, T, nct, c->d_cell_off, c->d_cell_off_alt, cnt, c->xi(), c->pxi(), c->xi_alt(), c->pxi_alt(), SE);
  }
  c->n_launches++;
  return 0;
}

// second half of fused_bnd_sort: the flags and the new patch offsets are on the host
static int fused_commit(Ctx* c, const uint32_t* h, uint32_t n_expected, bool multi)
{
  const int np = c->gd.n_patches;
  if (h[0]) {
    if (multi) {
      return fail("fused boundary+sort: a received particle lies outside its patch");
    }
    return fused_fallback(c);
  }
  if (c->fs_pull && h[3] > c->mv_cap) {
    // the push's mover list overflowed: the full scatter instead (the counters, the new cell
    // offsets and the pushed store are all still in place)
    c->fs_pull = false;
    c->n_pull_overflow++;
    c->last_n_movers = h[3];
    PSC_TRY(launch_scatter(c, false));
    PSC_CUDA_TRY(cudaStreamSynchronize(c->stream));
    PSC_TRY(check_launch(c, "fused_bnd_sort (scatter after a mover-list overflow)"));
  }
  if (c->fs_pull) {
    // the stayers cross over inside the next push (or in pull_materialize): the store is
    // "this buffer's live records + the other buffer's arrivals" until then
    c->fs_pull = false;
    c->pull_pending = true;
    c->last_n_movers = h[3];
    c->n_pulled++;
  } else {
    c->cur ^= 1;
    std::swap(c->d_cell_off, c->d_cell_off_alt);
  }
  for (int p = 0; p <= np; p++) {
    c->h_off[p] = h[4 + p];
  }
  c->n_prts = c->h_off[np];
  if (multi && c->n_prts != n_expected) {
    return fail("fused boundary+sort: particle count mismatch after the exchange");
  }
  c->n_dropped += h[1];
  c->sorted = true;
  c->pushed_from_sorted = false;
  c->n_fused++;
  return prts_upload_off(c);
}

// a deferred fused_bnd_sort (step_begin): wait for the scatter, commit its result
int fused_bnd_sort_finish(Ctx* c)
{
  if (!c->fs_deferred) {
    return 0;
  }
  c->fs_deferred = false;
  PSC_CUDA_TRY(cudaStreamSynchronize(c->stream));
  PSC_TRY(check_launch(c, "fused_bnd_sort"));
  return fused_commit(c, c->fs_host, c->fs_n_expected, c->fs_multi);
}

// defer: enqueue everything, leave the read-back of the flags / new patch offsets (and with it
// the host's wait for the scatter) to fused_bnd_sort_finish().  Multi-rank: the host still waits
// for the push and the exchange of the leavers in here (their counts size the transfers); what is
// deferred is the scatter behind them
int fused_bnd_sort(Ctx* c, bool defer)
{
  const GridDev& G = c->gd;
  const uint32_t rem_cap = c->rem_cap; // remote leavers listed by the push of this step (0: not)
  c->rem_cap = 0;
  const bool pull = c->pulled && c->opt_pull; // the push listed the movers: place them, leave the stayers
  c->pulled = false;
  c->fs_pull = false;
  if (!c->pushed_from_sorted) {
    c->counts_valid = false;
    PSC_TRY(bnd_particles(c));
    return sort_mprts(c);
  }
  const uint32_t nct = (uint32_t)G.n_cells * G.n_patches;
  const int np = G.n_patches;
  const bool multi = c->comm != nullptr;
  PSC_TRY(c->scr[9].reserve((size_t)nct * FS_PLANES * sizeof(cnt_t)));
  PSC_TRY(c->scr[10].reserve(((size_t)nct + 1) * sizeof(uint32_t)));
  PSC_TRY(c->scr[11].reserve((np + 1 + 4) * sizeof(uint32_t)));
  cnt_t* cnt = c->scr[9].as<cnt_t>();
  uint32_t* new_cnt = c->scr[10].as<uint32_t>();
  uint32_t* flags = c->scr[11].as<uint32_t>(); // [bad, dropped, remote, counter, new patch offsets...]
  uint32_t* d_new_off = flags + 4;
  FsTables T{c->d_patch_bnd, c->d_nei_patch};
  const bool counted_by_push = c->counts_valid;
  if (!counted_by_push) { // else: the push kernel already filled cnt and flags
    PSC_CUDA_TRY(cudaMemsetAsync(flags, 0, 4 * sizeof(uint32_t), c->stream));
    KernelScope ks(c, "fsort_count");
    k_fs_count<<<div_up(nct, FS_CELLS), FS_WARPS * 32, 0, c->stream>>>(G, T, nct, c->d_cell_off,
                                                                      c->xi(), cnt, flags);
    c->n_launches++;
  }
  c->counts_valid = false;
  {
    KernelScope ks(c, "fsort_offsets");
    const bool idx32 = 27ull * nct < (1ull << 32);
    if (pull) {
      PSC_TRY(c->scr[13].reserve((size_t)nct * sizeof(uint2)));
      auto k = idx32 ? k_fs_offsets_same<true, uint32_t> : k_fs_offsets_same<true, unsigned long long>;
      k<<<div_up(nct, 256), 256, 0, c->stream>>>(G, nct, cnt, new_cnt, flags, c->scr[13].as<uint2>());
    } else {
      auto k = idx32 ? k_fs_offsets_same<false, uint32_t> : k_fs_offsets_same<false, unsigned long long>;
      k<<<div_up(nct, 256), 256, 0, c->stream>>>(G, nct, cnt, new_cnt, flags, nullptr);
    }
    // faces towards a direction in which some patch has a neighbour (none along an
    // invariant direction: Grid_ / MrcDomain)
    FaceGeom FG{};
    const int ld[3] = {G.ldims[0], G.ldims[1], G.ldims[2]};
    for (int d = 0; d < 3; d++) {
      bool any = false;
      for (int p = 0; p < np && !any; p++) {
        for (int di = 0; di < 27 && !any; di++) {
          const int dir[3] = {di % 3 - 1, (di / 3) % 3 - 1, di / 9 - 1};
          any = dir[d] != 0 && c->h_nei_patch[p * 27 + di] >= 0;
        }
      }
      FG.a[d] = G.n_cells / ld[d];
      FG.s[d] = any ? (ld[d] > 1 ? 2 : 1) : 0;
      FG.per_patch += FG.s[d] * FG.a[d];
    }
    if (FG.per_patch) {
      k_fs_offsets_face<<<div_up((size_t)FG.per_patch * np, 128), 128, 0, c->stream>>>(G, FG, c->d_nei_patch, nct,
                                                                                        cnt, new_cnt, flags);
    }
    c->n_launches += 2;
  }
  uint32_t n_expected = c->n_prts;
  uint32_t n_movers_multi = 0;
  // ---- multi-rank: ship the leavers, merge the arrivals into the target cells' tails
  uint32_t n_recv_tot = 0;
  float4 *xr = nullptr, *pr = nullptr;
  uint32_t *skey = nullptr, *sidx = nullptr;
  if (multi) {
    uint32_t h4[4];
    PSC_CUDA_TRY(cudaMemcpyAsync(h4, flags, sizeof(h4), cudaMemcpyDeviceToHost, c->stream));
    PSC_CUDA_TRY(cudaStreamSynchronize(c->stream));
    // A rank whose precondition broke takes the general path right here.  The ranks stay
    // in step because BOTH paths issue exactly one comm_exchange_particles per call
    // (bnd_particles: fixup = false, below: fixup = true) over the same neighbour tables:
    // the sends / receives pair up whichever path a peer is on.  Keep it that way (an early
    // return in either that skips the exchange would hang NCCL).
    if (h4[0] != 0) {
      return fused_fallback(c);
    }
    const uint32_t n_rem = h4[2];
    n_movers_multi = h4[3];
    if (!c->rf_built) {
      PSC_TRY(build_remote_cells(c));
    }
    uint32_t *gk = nullptr, *gi = nullptr;
    c->last_n_rem = n_rem;
    if (n_rem) {
      KernelScope ks(c, "fsort_remote_collect");
      uint32_t *k0, *v0, *k1, *v1;
      // pass 1 sorts by index (keys = index, values = group), pass 2 by group
      if (counted_by_push && n_rem <= rem_cap) { // the push listed them (push.cu launch_lean) ...
        uint32_t* a = c->scr[7].as<uint32_t>();
        k0 = a, v0 = a + rem_cap, k1 = a + 2 * (size_t)rem_cap, v1 = a + 3 * (size_t)rem_cap;
      } else { // ... or it did not, or its list overflowed: pass over the boundary cells
        PSC_TRY(c->scr[7].reserve(4 * (size_t)n_rem * sizeof(uint32_t)));
        uint32_t* a = c->scr[7].as<uint32_t>();
        k0 = a, v0 = a + n_rem, k1 = a + 2 * (size_t)n_rem, v1 = a + 3 * (size_t)n_rem;
        PSC_CUDA_TRY(cudaMemsetAsync(flags + 3, 0, sizeof(uint32_t), c->stream)); // (a pull push counted its movers there)
        k_fs_collect_remote<<<div_up(c->n_rf_cells, FS_WARPS), FS_WARPS * 32, 0, c->stream>>>(
          G, T, c->n_rf_cells, c->rf_cells.as<uint32_t>(), c->d_cell_off, c->xi(), v0, k0, flags + 3,
          n_rem);
        c->n_launches++;
      }
      bool in_alt = false;
      PSC_TRY(sort_pairs(c, k0, v0, k1, v1, n_rem, 32, false, &in_alt));
      uint32_t *ik = in_alt ? k1 : k0, *iv = in_alt ? v1 : v0; // (index, group) by index
      uint32_t *ok = in_alt ? k0 : k1, *ov = in_alt ? v0 : v1;
      size_t key_space = (size_t)c->g.n_ranks * np * 32;
      int bits = 1;
      while ((size_t(1) << bits) < key_space) {
        bits++;
      }
      // keys = group (iv), values = index (ik)
      PSC_TRY(sort_pairs(c, iv, ik, ov, ok, n_rem, bits, false, &in_alt));
      gk = in_alt ? ov : iv;
      gi = in_alt ? ok : ik;
    }
    std::vector<uint32_t> n_recv(np, 0);
    PSC_TRY(comm_exchange_particles(c, c->xi(), c->pxi(), gi, gk, n_rem, 0, true, n_recv, &xr, &pr));
    std::vector<uint32_t> roff(np + 1, 0);
    for (int p = 0; p < np; p++) {
      roff[p + 1] = roff[p] + n_recv[p];
    }
    n_recv_tot = roff[np];
    n_expected = c->n_prts - n_rem - h4[1] + n_recv_tot;
    if (n_recv_tot) {
      KernelScope ks(c, "fsort_remote_merge");
      PSC_TRY(c->scr[6].reserve((np + 1 + 4 * (size_t)n_recv_tot) * sizeof(uint32_t)));
      uint32_t* d_roff = c->scr[6].as<uint32_t>();
      uint32_t *k0 = d_roff + np + 1, *v0 = k0 + n_recv_tot, *k1 = v0 + n_recv_tot, *v1 = k1 + n_recv_tot;
      PSC_CUDA_TRY(cudaMemcpyAsync(d_roff, roff.data(), (np + 1) * sizeof(uint32_t),
                                   cudaMemcpyHostToDevice, c->stream));
      k_fs_recv_keys<<<div_up(n_recv_tot, 256), 256, 0, c->stream>>>(G, n_recv_tot, d_roff, xr, k0,
                                                                    new_cnt, flags);
      c->n_launches++;
      int bits = 1;
      while ((size_t(1) << bits) < nct) {
        bits++;
      }
      bool in_alt = false;
      PSC_TRY(sort_pairs(c, k0, v0, k1, v1, n_recv_tot, bits, true, &in_alt));
      skey = in_alt ? k1 : k0;
      sidx = in_alt ? v1 : v0;
      PSC_CUDA_TRY(cudaStreamSynchronize(c->stream)); // roff is a host temporary
    }
    PSC_TRY(prts_reserve(c, n_expected));
  }
  {
    KernelScope ks(c, "fsort_scan");
    PSC_TRY(scan_exclusive<uint32_t>(c, LoadArr<uint32_t>{new_cnt}, nct, c->d_cell_off_alt, c->scr[2]));
  }
  uint32_t n_movers_host = 0xffffffffu; // (multi-rank: known on the host already)
  if (multi) {
    n_movers_host = std::min(n_movers_multi, c->mv_cap);
  }
  {
    KernelScope ks(c, pull ? "fsort_place_movers" : "fsort_scatter");
    if (pull) {
      const uint2* mv_key = c->scr[12].as<uint2>();
      const uint32_t* mv_idx = reinterpret_cast<const uint32_t*>(mv_key + c->mv_cap);
      k_fs_place_movers<<<148 * 8, 256, 0, c->stream>>>(G, T, nct, flags, c->mv_cap, n_movers_host, mv_idx, mv_key,
                                                       cnt, c->d_cell_off_alt, c->xi(), c->pxi(), c->xi_alt(),
                                                       c->pxi_alt());
      c->n_launches++;
      c->fs_pull = true;
    } else {
      PSC_TRY(launch_scatter(c, c->want_scatter_energies && !multi));
    }
    if (n_recv_tot) {
      k_fs_place_remote<<<div_up(n_recv_tot, 256), 256, 0, c->stream>>>(
        n_recv_tot, skey, sidx, c->d_cell_off_alt, xr, pr, c->xi_alt(), c->pxi_alt());
      c->n_launches++;
    }
    k_patch_offsets<<<div_up(np + 1, 128), 128, 0, c->stream>>>(c->d_cell_off_alt, np, G.n_cells,
                                                               d_new_off);
  }
  c->n_launches += 1;
  if (defer) {
    const size_t need = (size_t)(np + 1 + 4) * sizeof(uint32_t);
    if (c->fs_host_bytes < need) {
      if (c->fs_host) {
        cudaFreeHost(c->fs_host);
        c->fs_host = nullptr;
      }
      PSC_CUDA_TRY(cudaMallocHost(&c->fs_host, need));
      c->fs_host_bytes = need;
    }
    PSC_CUDA_TRY(cudaMemcpyAsync(c->fs_host, flags, need, cudaMemcpyDeviceToHost, c->stream));
    c->fs_deferred = true;
    c->fs_n_expected = n_expected;
    c->fs_multi = multi;
    return 0;
  }
  std::vector<uint32_t> h(np + 1 + 4);
  PSC_CUDA_TRY(cudaMemcpyAsync(h.data(), flags, h.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                               c->stream));
  PSC_CUDA_TRY(cudaStreamSynchronize(c->stream));
  PSC_TRY(check_launch(c, "fused_bnd_sort"));
  return fused_commit(c, h.data(), n_expected, multi);
}


// ====================================================================== pull mode

// pull mode needs: the lean push with its class counts and SAME cell arithmetic (push.cu checks
// those when it launches), no reflecting particle walls (a reflected particle that stays in its
// cell would have to be re-ranked among the stayers), plane entries that fit the mover key
bool pull_possible(const Ctx* c)
{
  if (!c->opt_pull || !c->opt_fused_sort || c->opt_gapped || c->comm) {
    return false; // (multi-rank: the remote arrivals' placement has not been exercised in pull mode)
  }
  for (int d = 0; d < 3; d++) {
    if (!c->g.invar[d] && (c->g.desc.bc_prt_lo[d] == PSC_B200_BND_PRT_REFLECTING ||
                           c->g.desc.bc_prt_hi[d] == PSC_B200_BND_PRT_REFLECTING)) {
      return false;
    }
  }
  return (size_t)FS_PLANES * c->gd.n_cells * c->gd.n_patches < (size_t(1) << 32);
}

// from a cell-ordered store: the output layout is this one, nothing lies in front of the
// stayers, everybody stays
int pull_enter(Ctx* c)
{
  const uint32_t nct = (uint32_t)c->gd.n_cells * c->gd.n_patches;
  PSC_TRY(c->scr[13].reserve((size_t)nct * sizeof(uint2)));
  PSC_CUDA_TRY(cudaMemcpyAsync(c->d_cell_off_alt, c->d_cell_off, ((size_t)nct + 1) * sizeof(uint32_t),
                               cudaMemcpyDeviceToDevice, c->stream));
  k_stay_init<<<div_up(nct, 256), 256, 0, c->stream>>>(nct, c->d_cell_off, c->scr[13].as<uint2>());
  c->n_launches++;
  c->pull_pending = true;
  return check_launch(c, "pull_enter");
}

int pull_materialize(Ctx* c)
{
  if (!c->pull_pending) {
    return 0;
  }
  const GridDev& G = c->gd;
  const uint32_t nct = (uint32_t)G.n_cells * G.n_patches;
  {
    KernelScope ks(c, "pull_materialize");
    k_pull_stayers<<<div_up(nct, FS_WARPS), FS_WARPS * 32, 0, c->stream>>>(
      G, nct, c->d_cell_off, c->d_cell_off_alt, c->scr[13].as<uint2>(), c->xi(), c->pxi(), c->xi_alt(),
      c->pxi_alt());
    c->n_launches++;
  }
  c->cur ^= 1;
  std::swap(c->d_cell_off, c->d_cell_off_alt);
  c->pull_pending = false;
  c->n_pull_materialized++;
  return check_launch(c, "pull_materialize");
}

// ====================================================================== gapped store

namespace
{

// per target cell: like k_fs_offsets, but the positions are relative to the start of the
// cell's run, the run is anchored so that the stayers (already written by the push) sit at
// V + RL, and a slab that cannot take its arrivals is flagged
__global__ void k_gap_offsets(GridDev G, const int* __restrict__ nei_patch, uint32_t nct,
                              cnt_t* __restrict__ cnt, const uint32_t* __restrict__ v, uint32_t rl,
                              uint32_t* __restrict__ new_start, uint32_t* __restrict__ new_n,
                              uint32_t* __restrict__ nstay, uint32_t* __restrict__ ctl)
{
  uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= nct) {
    return;
  }
  const int q = g / G.n_cells;
  const int c = g - q * G.n_cells;
  const int ld0 = G.ldims[0], ld1 = G.ldims[1], ld2 = G.ldims[2];
  const int c0 = c % ld0, c1 = (c / ld0) % ld1, c2 = c / (ld0 * ld1);
  uint32_t total = 0, nl = 0, nc = 0;
  fs_route<true>(G, nct, cnt, q, c0, c1, c2, 0, 0, 0, total, &nl, &nc);
  if (c0 == 0 || c0 == ld0 - 1 || c1 == 0 || c1 == ld1 - 1 || c2 == 0 || c2 == ld2 - 1) {
    fs_face_routes(G, nei_patch, nct, cnt, q, c0, c1, c2, total);
  }
  const uint32_t v0 = v[g], v1 = v[g + 1];
  new_start[g] = v0 + rl - nl; // (meaningless when nl > rl: the store is re-laid out then)
  new_n[g] = total;
  nstay[g] = nc;
  if (nl > rl || rl - nl + total > v1 - v0) {
    atomicExch(&ctl[GAP_CTL_OVERFLOW], 1u);
  }
  if (nl > rl) {
    atomicMax(&ctl[GAP_CTL_MAX_NL], nl);
  }
  if (total > CNT_MAX) {
    atomicExch(&ctl[GAP_CTL_M_FULL], 1u); // 16-bit planes: the step is redone on the eager path
  }
}

// per patch: population of the new store (one CTA per patch)
__global__ void __launch_bounds__(256)
  k_gap_patch_sums(const uint32_t* __restrict__ n, int n_cells, uint32_t* __restrict__ out)
{
  __shared__ uint32_t ws[8];
  const uint32_t* np = n + (size_t)blockIdx.x * n_cells;
  uint32_t s = 0;
  for (int i = threadIdx.x; i < n_cells; i += 256) {
    s += np[i];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(FULL, s, o);
  }
  if ((threadIdx.x & 31) == 0) {
    ws[threadIdx.x >> 5] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int k = 0; k < 8; k++) {
      t += ws[k];
    }
    out[blockIdx.x] = t;
  }
}

// one thread per mover slot: the record goes to start[target] + offset of its (source
// cell, class) group inside the target's run + its rank inside the group
__global__ void k_gap_place(uint32_t n_slots, const uint4* __restrict__ mtag, const float4* __restrict__ mx,
                            const float4* __restrict__ mp, const uint32_t* __restrict__ start,
                            const cnt_t* __restrict__ rel, float4* __restrict__ xo,
                            float4* __restrict__ po)
{
  uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_slots) {
    return;
  }
  const uint4 t = mtag[s];
  if (!t.w) {
    return;
  }
  const uint32_t dst = __ldg(&start[t.x]) + __ldg(&rel[t.y]) + t.z;
  xo[dst] = mx[s];
  po[dst] = mp[s];
}

// copy one run per cell: [src_start[t] (+ src_bias), + n[t]) -> [dst_start[t] + dst_bias (+
// dst_shift[t]), ...).  One warp per GC_CPW consecutive cells.
constexpr int GC_WARPS = 8;
constexpr int GC_CPW = 8;

__global__ void __launch_bounds__(GC_WARPS * 32)
  k_gap_copy_runs(uint32_t nct, const uint32_t* __restrict__ src_start, uint32_t src_bias,
                  const uint32_t* __restrict__ n, const uint32_t* __restrict__ dst_start, uint32_t dst_bias,
                  const uint32_t* __restrict__ dst_shift_a, const uint32_t* __restrict__ dst_shift_b,
                  const float4* __restrict__ xi, const float4* __restrict__ pi, float4* __restrict__ xo,
                  float4* __restrict__ po, uint32_t* __restrict__ new_start)
{
  const int lane = threadIdx.x & 31;
  const uint32_t g0 = (blockIdx.x * GC_WARPS + (threadIdx.x >> 5)) * GC_CPW;
  for (int k = 0; k < GC_CPW; k++) {
    const uint32_t g = g0 + k;
    if (g >= nct) {
      return;
    }
    const uint32_t cnt = n[g], src = src_start[g] + src_bias;
    uint32_t dst = dst_start[g] + dst_bias;
    if (new_start && lane == 0) {
      new_start[g] = dst; // the run begins where the lower movers will be placed
    }
    if (dst_shift_a) {
      // + (a - b): the lower movers of the cell come first (a = V + RL, b = old run start)
      dst += dst_shift_a[g] + src_bias - dst_shift_b[g];
    }
    for (uint32_t i = lane; i < cnt; i += 32) {
      xo[dst + i] = xi[src + i];
      po[dst + i] = pi[src + i];
    }
  }
}

__global__ void k_gap_runs_of_sorted(uint32_t nct, const uint32_t* __restrict__ cell_off,
                                     uint32_t* __restrict__ start, uint32_t* __restrict__ n)
{
  uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g < nct) {
    const uint32_t a = cell_off[g];
    start[g] = a;
    n[g] = cell_off[g + 1] - a;
  }
}

// slab size of a cell = RL + population + slack
struct SlabSize
{
  const uint32_t* n;
  uint32_t extra;
  __device__ __forceinline__ uint32_t operator()(size_t i) const { return n[i] + extra; }
};

} // namespace

void gap_release(Ctx* c)
{
  cudaFree(c->g_v);
  cudaFree(c->g_v_alt);
  for (int b = 0; b < 2; b++) {
    cudaFree(c->g_start[b]);
    cudaFree(c->g_n[b]);
    c->g_start[b] = c->g_n[b] = nullptr;
  }
  cudaFree(c->g_nstay);
  cudaFree(c->g_ctl);
  cudaFree(c->mvx);
  cudaFree(c->mvp);
  cudaFree(c->mvtag);
  c->g_v = c->g_v_alt = c->g_nstay = c->g_ctl = nullptr;
  c->mvx = c->mvp = nullptr;
  c->mvtag = nullptr;
  c->mov_cap = 0;
  c->gapped = false;
}

static uint32_t gap_slack(const Ctx* c, size_t nct)
{
  if (c->opt_gap_slack > 0) {
    return (uint32_t)c->opt_gap_slack;
  }
  // half the mean population: room for the Poisson-level drift of a cell's population
  size_t mean = c->n_prts / std::max<size_t>(nct, 1);
  return (uint32_t)std::min<size_t>(64, std::max<size_t>(8, mean / 2));
}

// allocations (first use / growth), layout of the store the push writes, per-step clears.
// *ok = false (and nothing changed): this step cannot take the gapped path.
int gap_prepare(Ctx* c, bool* ok)
{
  const GridDev& G = c->gd;
  const size_t nct = (size_t)G.n_cells * G.n_patches;
  *ok = false;
  if (nct * FS_PLANES >= (size_t(1) << 32) || c->n_prts == 0) {
    return 0;
  }
  if (!c->gapped) {
    // entering from a cell-ordered contiguous store: lay out the slabs around its runs
    const uint32_t slack = gap_slack(c, nct);
    const size_t slots = (size_t)c->n_prts + nct * ((size_t)c->g_rl + slack) + c->g_rl + 1024;
    if (slots >= (size_t(1) << 32)) {
      return 0;
    }
    if (slots > c->cap) {
      // the store grows one pair of buffers at a time (prts_reserve): is there room for
      // the slack, the mover list and the counters?
      size_t free_b = 0, total_b = 0;
      PSC_CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
      const size_t have = free_b + c->cap * 4 * sizeof(float4);
      const size_t need = slots * 4 * sizeof(float4) + (size_t)c->n_prts / 8 * 48 + nct * (FS_PLANES + 8) * 4 +
                          (size_t(2) << 30);
      if (need > have) {
        c->opt_gapped = 0; // not enough memory for the slack: stay on the eager path
        return 0;
      }
      PSC_TRY(prts_reserve(c, slots));
    }
    c->g_slack = slack;
  }
  if (!c->g_ctl) {
    PSC_CUDA_TRY(cudaMalloc(&c->g_v, (nct + 1) * sizeof(uint32_t)));
    PSC_CUDA_TRY(cudaMalloc(&c->g_v_alt, (nct + 1) * sizeof(uint32_t)));
    for (int b = 0; b < 2; b++) {
      PSC_CUDA_TRY(cudaMalloc(&c->g_start[b], nct * sizeof(uint32_t)));
      PSC_CUDA_TRY(cudaMalloc(&c->g_n[b], nct * sizeof(uint32_t)));
    }
    PSC_CUDA_TRY(cudaMalloc(&c->g_nstay, nct * sizeof(uint32_t)));
    PSC_CUDA_TRY(cudaMalloc(&c->g_ctl, GAP_CTL_WORDS * sizeof(uint32_t)));
  }
  // mover list: every particle of a small store may move, an eighth of a big one; grown
  // when the last step used more than half
  size_t want = c->n_prts <= (size_t(1) << 24) ? (size_t)c->n_prts + GAP_BATCH * 4096 : (size_t)c->n_prts / 8;
  if (c->g_mov_used > c->mov_cap / 2) {
    want = std::max(want, 2 * c->mov_cap);
  }
  if (want > c->mov_cap) {
    cudaFree(c->mvx);
    cudaFree(c->mvp);
    cudaFree(c->mvtag);
    c->mvx = c->mvp = nullptr;
    c->mvtag = nullptr;
    c->mov_cap = 0;
    PSC_CUDA_TRY(cudaMalloc(&c->mvx, want * sizeof(float4)));
    PSC_CUDA_TRY(cudaMalloc(&c->mvp, want * sizeof(float4)));
    PSC_CUDA_TRY(cudaMalloc(&c->mvtag, want * sizeof(uint4)));
    c->mov_cap = want;
  }
  if (!c->gapped) {
    KernelScope ks(c, "gap_layout");
    k_gap_runs_of_sorted<<<div_up(nct, 256), 256, 0, c->stream>>>((uint32_t)nct, c->d_cell_off, c->g_start[0],
                                                                   c->g_n[0]);
    c->n_launches++;
    PSC_TRY(scan_exclusive<uint32_t>(c, SlabSize{c->g_n[0], c->g_rl + c->g_slack}, nct, c->g_v, c->scr[2]));
  }
  PSC_CUDA_TRY(cudaMemsetAsync(c->g_ctl, 0, GAP_CTL_WORDS * sizeof(uint32_t), c->stream));
  *ok = true;
  return 0;
}

// after the gapped push: group offsets, overflow check, (re-layout,) mover placement,
// commit.  *redo = true: nothing was committed, the step has to take the eager path (the
// store the push read is intact).
int gap_finish(Ctx* c, bool* redo)
{
  const GridDev& G = c->gd;
  const size_t nct = (size_t)G.n_cells * G.n_patches;
  const int np = G.n_patches;
  *redo = false;
  cnt_t* cnt = c->scr[9].as<cnt_t>();
  uint32_t* flags = c->scr[11].as<uint32_t>(); // [bad, dropped, remote, -, new patch sizes...]
  {
    KernelScope ks(c, "gap_offsets");
    k_gap_offsets<<<div_up(nct, 128), 128, 0, c->stream>>>(G, c->d_nei_patch, (uint32_t)nct, cnt, c->g_v,
                                                           c->g_rl, c->g_start[1], c->g_n[1], c->g_nstay,
                                                           c->g_ctl);
    k_gap_patch_sums<<<np, 256, 0, c->stream>>>(c->g_n[1], G.n_cells, flags + 4);
    c->n_launches += 2;
  }
  std::vector<uint32_t> h(np + 4);
  uint32_t ctl[GAP_CTL_WORDS];
  PSC_CUDA_TRY(cudaMemcpyAsync(h.data(), flags, h.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  PSC_CUDA_TRY(cudaMemcpyAsync(ctl, c->g_ctl, sizeof(ctl), cudaMemcpyDeviceToHost, c->stream));
  PSC_CUDA_TRY(cudaStreamSynchronize(c->stream));
  PSC_TRY(check_launch(c, "gap_offsets"));
  c->counts_valid = false;
  c->g_mov_used = ctl[GAP_CTL_MOVERS];
  if (h[0] || h[2] || ctl[GAP_CTL_M_FULL]) {
    // a particle moved further than one cell / left for another rank / the mover list is
    // full: nothing has been committed
    *redo = true;
    c->n_gap_redone++;
    if (ctl[GAP_CTL_M_FULL]) {
      c->g_mov_used = (uint32_t)std::min<size_t>(c->mov_cap, 0xffffffffu); // grows the list next time
    }
    return 0;
  }
  size_t n_new = 0;
  for (int p = 0; p < np; p++) {
    n_new += h[4 + p];
  }
  const uint32_t n_slots = std::min<uint32_t>(ctl[GAP_CTL_MOVERS], (uint32_t)std::min<size_t>(c->mov_cap, 0xffffffffu));
  float4 *xo = c->xi_alt(), *po = c->pxi_alt();
  bool relayout = ctl[GAP_CTL_OVERFLOW] != 0;
  if (relayout) {
    // fresh slabs around the new populations; the stayers move from the buffer the push
    // wrote back into the one it read (consumed), behind the room of their lower movers
    uint32_t rl_new = c->g_rl;
    if (ctl[GAP_CTL_MAX_NL] > c->g_rl) {
      rl_new = 2 * ctl[GAP_CTL_MAX_NL];
    }
    uint32_t slack = gap_slack(c, nct);
    size_t slots = n_new + nct * ((size_t)rl_new + slack) + rl_new + 1024;
    if (slots > c->cap) {
      // shrink the slack to what the buffers hold
      size_t fixed = n_new + nct * (size_t)rl_new + rl_new + 1024;
      if (fixed + nct * 4 > c->cap) {
        // no room at all: this step is redone on the eager path, which regrows the store
        *redo = true;
        c->n_gap_redone++;
        return 0;
      }
      slack = (uint32_t)((c->cap - fixed) / nct);
    }
    KernelScope ks(c, "gap_relayout");
    PSC_TRY(scan_exclusive<uint32_t>(c, SlabSize{c->g_n[1], rl_new + slack}, nct, c->g_v_alt, c->scr[2]));
    // stayers: [V + RL, + nstay) of the written buffer -> V' + RL' + n_L of the read one;
    // n_L = (V + RL) - start_new
    k_gap_copy_runs<<<div_up(nct, GC_WARPS * GC_CPW), GC_WARPS * 32, 0, c->stream>>>(
      (uint32_t)nct, c->g_v, c->g_rl, c->g_nstay, c->g_v_alt, rl_new, c->g_v, c->g_start[1], xo, po, c->xi(),
      c->pxi(), c->g_start[0]);
    c->n_launches++;
    std::swap(c->g_v, c->g_v_alt);
    std::swap(c->g_n[0], c->g_n[1]);
    // g_start[0] = V' + RL' was written by the copy; the new store is the buffer that was read
    xo = c->xi();
    po = c->pxi();
    c->g_rl = rl_new;
    c->g_slack = slack;
    c->n_gap_relayouts++;
  } else {
    std::swap(c->g_start[0], c->g_start[1]);
    std::swap(c->g_n[0], c->g_n[1]);
    c->cur ^= 1;
  }
  if (n_slots) {
    KernelScope ks(c, "gap_place");
    k_gap_place<<<div_up(n_slots, 256), 256, 0, c->stream>>>(n_slots, c->mvtag, c->mvx, c->mvp, c->g_start[0], cnt,
                                                             xo, po);
    c->n_launches++;
  }
  PSC_TRY(check_launch(c, "gap_place"));
  c->h_off[0] = 0;
  for (int p = 0; p < np; p++) {
    c->h_off[p + 1] = c->h_off[p] + h[4 + p];
  }
  c->n_prts = c->h_off[np];
  c->n_dropped += h[1];
  c->gapped = true;
  c->sorted = false;
  c->pushed_from_sorted = false;
  c->n_gap_steps++;
  return prts_upload_off(c);
}

// gapped store -> the reference's contiguous patch-by-patch array, ordered by cell
int gap_compact(Ctx* c)
{
  if (!c->gapped) {
    return 0;
  }
  const GridDev& G = c->gd;
  const size_t nct = (size_t)G.n_cells * G.n_patches;
  {
    KernelScope ks(c, "gap_compact");
    PSC_TRY(scan_exclusive<uint32_t>(c, LoadArr<uint32_t>{c->g_n[0]}, nct, c->d_cell_off, c->scr[2]));
    k_gap_copy_runs<<<div_up(nct, GC_WARPS * GC_CPW), GC_WARPS * 32, 0, c->stream>>>(
      (uint32_t)nct, c->g_start[0], 0u, c->g_n[0], c->d_cell_off, 0u, nullptr, nullptr, c->xi(), c->pxi(),
      c->xi_alt(), c->pxi_alt(), nullptr);
    c->n_launches++;
  }
  PSC_TRY(check_launch(c, "gap_compact"));
  c->cur ^= 1;
  c->gapped = false;
  c->sorted = true; // ordered by (patch, cell), d_cell_off describes it
  c->pushed_from_sorted = false;
  return 0;
}

} // namespace psc_b200
