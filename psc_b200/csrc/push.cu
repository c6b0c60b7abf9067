// psc_b200: PushParticlesB200::push_mprts -- 1st-order "ec" gather, Boris push, move and
// charge-conserving 1vb current deposit in one kernel.
//
// What: PushParticlesVb<C>::push_mprts (libpsc/psc_push_particles/push_particles_1vb.hxx:27-84)
// with Current1vbVar1 (yz) / Current1vbSplit (xyz, yz); the per-particle arithmetic is
// pic_math.cuh.  How (ours, not PSC's cuda_push_mprts_yz.cxx):
//
//  k_push_tiled   the store is ordered by (patch, cell).  One CTA owns a tile of cells of
//                 one patch (8x8x8 in 3D, 16x16 in yz; compile-time geometry whenever the
//                 patch is a multiple of it).  It stages the E/B tile (+halo) into shared
//                 memory with cp.async.bulk (TMA bulk copies, one per contiguous row,
//                 completion on an mbarrier) -- or LDG/STS for odd geometries --, keeps a
//                 J tile of the same shape in shared memory, and walks the tile's particle
//                 runs warp by warp with 128-bit loads/stores, the next chunk prefetched
//                 while the current one is computed.
//                   * particles that stay in their cell (one trajectory segment, ~95 %):
//                     lanes are grouped by target cell (sorted store => 1-2 groups per
//                     warp) and each group's 8/12 deposit values are summed with a
//                     transposing shuffle butterfly (9/16 SHFL instead of 40/60); a few
//                     lanes then issue one shared-memory atomic each
//                   * particles that cross a cell face are parked in a per-warp shared
//                     queue and split/deposited 32 at a time, so the divergent
//                     Villasenor-Buneman walk runs with full warps
//                 The J tile (halo included) is flushed with global red.add, so no
//                 checkerboard passes are needed.  Optionally the kernel also counts, per
//                 source cell, where its particles went (27 direction classes): the input
//                 of the fused boundary-exchange + sort pass (fused_sort.cu).
//  k_push_general any particle order: one thread per particle, fields through the
//                 read-only path, J with global red.add.  Used when the store is not
//                 sorted (same results, same particle order).
//
// This file is compiled twice: -fmad=false (namespace exact: particle update is
// bit-identical to PSC's x86-64 build, which has no FMA) and with FMA contraction
// (namespace fast, within a few ULP).
#include "gap.cuh"

#include <cuda.h> // CUtensorMap (the type only: the encoder is fetched through the runtime, ctx.hpp)

#include <algorithm>
#include <cstring>

#ifndef PUSH_VARIANT
#define PUSH_VARIANT exact
#endif
#define PUSH_CAT_(a, b) a##b
#define PUSH_CAT(a, b) PUSH_CAT_(a, b)

namespace psc_b200
{
namespace PUSH_VARIANT
{

constexpr unsigned FULL = 0xffffffffu;
constexpr int QCAP = 56;     // crossing-particle queue entries per warp
constexpr int QCAP_GAP = 48; // ... of the gapped-store variant (makes room for its count table)
constexpr int CT = 30;       // GAP: per-warp count table [cell of the row][class 0..26, DROP, BAD, REMOTE]

template <bool GAP>
__host__ __device__ constexpr int qcap()
{
  return GAP ? QCAP_GAP : QCAP;
}
// uint16 entries of the count table of one warp (a multiple of 8: 16-byte granules)
__host__ __device__ inline int ct_entries(int row_cells)
{
  return (row_cells * CT + 7) & ~7;
}

// ---------------------------------------------------------------- tile geometry
// f = shared tile extent in nodes, g = distance of tile node 0 below the tile origin.
// Deposits touch cells [o-1, o+t] and their +1 neighbours, the gather nodes [o-1, o+t+1]
// (the extra margin covers float(dx_inv) vs 1/float(dx) disagreeing by one ulp).

struct GeoDyn
{
  int t_[3], nt_[3], f_[3], g_[3];
  __host__ __device__ int t(int d) const { return t_[d]; }
  __host__ __device__ int nt(int d) const { return nt_[d]; }
  __host__ __device__ int f(int d) const { return f_[d]; }
  __host__ __device__ int g(int d) const { return g_[d]; }
  __host__ __device__ int sy() const { return f_[0]; }
  __host__ __device__ int sz() const { return f_[0] * f_[1]; }
  __host__ __device__ int sm() const { return f_[0] * f_[1] * f_[2]; }
};

// compile-time geometry; the contiguous direction (x in 3D, y in yz) starts two nodes
// below the tile origin and is four nodes longer than the tile so that every row is a
// 16-byte aligned, 16-byte multiple bulk copy
template <int DIM>
struct GeoStatic
{
  int nt_[3];
  static constexpr bool XYZ = DIM == pm::DIM_XYZ;
#ifndef PUSH_TILE_X
#define PUSH_TILE_X 16
#define PUSH_TILE_Y 4
#define PUSH_TILE_Z 4
#endif
  __host__ __device__ static constexpr int t(int d)
  {
    return XYZ ? (d == 0 ? PUSH_TILE_X : (d == 1 ? PUSH_TILE_Y : PUSH_TILE_Z)) : (d == 0 ? 1 : 16);
  }
  __host__ __device__ int nt(int d) const { return nt_[d]; }
  __host__ __device__ static constexpr int f(int d)
  {
    return XYZ ? (d == 0 ? PUSH_TILE_X + 4 : (d == 1 ? PUSH_TILE_Y + 3 : PUSH_TILE_Z + 3))
               : (d == 0 ? 1 : (d == 1 ? 20 : 19));
  }
  __host__ __device__ static constexpr int g(int d)
  {
    return XYZ ? (d == 0 ? 2 : 1) : (d == 0 ? 0 : (d == 1 ? 2 : 1));
  }
  __host__ __device__ static constexpr int sy() { return f(0); }
  __host__ __device__ static constexpr int sz() { return f(0) * f(1); }
  __host__ __device__ static constexpr int sm() { return f(0) * f(1) * f(2); }
};

// ---------------------------------------------------------------- field accessors

struct FldGlobal
{
  const float* __restrict__ F;
  const GridDev& G;
  __device__ __forceinline__ float operator()(int m, int i, int j, int k) const
  {
    return __ldg(F + fld_off(G, m, i, j, k));
  }
};

template <typename GEO>
struct FldTile
{
  const float* s; // EM tile, component-major; s points at node (n0, n1, n2)
  const GEO& geo;
  int n0, n1, n2;
  __device__ __forceinline__ float operator()(int m, int i, int j, int k) const
  {
    return s[(m - pm::EX) * geo.sm() + (k - n2) * geo.sz() + (j - n1) * geo.sy() + (i - n0)];
  }
};

// ---------------------------------------------------------------- leaves

template <int DIM, int DEPOSIT>
struct Walker;

template <int DIM>
struct Walker<DIM, pm::DEPOSIT_SPLIT>
{
  pm::SplitWalker<DIM> w;
  __device__ __forceinline__ bool first(const pm::PushConst& c, const pm::Trajectory& t,
                                        float qw, int ci[3], float* val)
  {
    w.begin(c, t);
    w.descend();
    pm::split_leaf<DIM>(c, qw, w.a, w.b, ci, val);
    if (DIM == pm::DIM_YZ) {
      ci[0] = 0; // Fields3d forces invariant indices to 0 (fields.hxx:50-57)
    }
    return w.pending != 0;
  }
  __device__ __forceinline__ bool next(const pm::PushConst& c, float qw, int ci[3], float* val)
  {
    w.pop();
    w.descend();
    pm::split_leaf<DIM>(c, qw, w.a, w.b, ci, val);
    if (DIM == pm::DIM_YZ) {
      ci[0] = 0;
    }
    return w.pending != 0;
  }
};

template <>
struct Walker<pm::DIM_YZ, pm::DEPOSIT_VAR1>
{
  pm::Var1Walker w;
  __device__ __forceinline__ bool first(const pm::PushConst& c, const pm::Trajectory& t,
                                        float qw, int ci[3], float* val)
  {
    w.begin(c, t);
    w.next(c, qw, ci, val);
    return w.n_left > 0;
  }
  __device__ __forceinline__ bool next(const pm::PushConst& c, float qw, int ci[3], float* val)
  {
    w.next(c, qw, ci, val);
    return w.n_left > 0;
  }
};

// leaf value n -> (component, offset) as a linear offset for strides (sy, sz, sm)
template <int DIM>
__device__ __forceinline__ constexpr int leaf_lin(int n, int sy, int sz, int sm)
{
  // value order of pm::split_leaf / Var1Walker::cell_values (pic_math.cuh)
  if (DIM == pm::DIM_XYZ) {
    return n < 4 ? ((n >> 1) * sz + (n & 1) * sy)
                 : (n < 8 ? (sm + ((n - 4) & 1) * sz + ((n - 4) >> 1))
                          : (2 * sm + (((n - 8) >> 1) * sy) + ((n - 8) & 1)));
  }
  return n < 4 ? ((n >> 1) * sz + (n & 1) * sy)
               : (n < 6 ? (sm + (n - 4) * sz) : (2 * sm + (n - 6) * sy));
}

template <int DIM>
__device__ __forceinline__ void leaf_to_global(const GridDev& G, float* F, const int ci[3],
                                               const float* val)
{
  constexpr int NV = pm::LeafShape<DIM>::NV;
  // guard against writes outside the patch array (cannot happen for particles that
  // index into their patch; the CPU code would scribble over the next component)
  bool ok = true;
#pragma unroll
  for (int d = 0; d < 3; d++) {
    if (DIM == pm::DIM_YZ && d == 0) {
      continue;
    }
    ok = ok && ci[d] >= -G.ibn[d] && ci[d] + 1 < G.ldims[d] + G.ibn[d];
  }
  if (!ok) {
    return;
  }
  float* base = F + fld_off(G, 0, ci[0], ci[1], ci[2]);
  int sy = G.im[0], sz = G.im[0] * G.im[1];
  int sm = (int)G.fld_len;
#pragma unroll
  for (int n = 0; n < NV; n++) {
    atomicAdd(base + leaf_lin<DIM>(n, sy, sz, sm), val[n]);
  }
}

// ---------------------------------------------------------------- general kernel

template <int DIM, int DEPOSIT>
__global__ void __launch_bounds__(256)
  k_push_general(GridDev G, uint32_t n, const uint32_t* __restrict__ off, float4* __restrict__ xi4,
                 float4* __restrict__ pxi4, float* __restrict__ flds, long slot_len)
{
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) {
    return;
  }
  int p = patch_of(off, G.n_patches, i);
  float* F = flds + p * slot_len;
  float4 X = xi4[i], U = pxi4[i];
  float x[3] = {X.x, X.y, X.z}, u[3] = {U.x, U.y, U.z};
  int kind = __float_as_int(X.w);
  FldGlobal EM{F, G};
  pm::Trajectory t;
  pm::advance<DIM>(G.pc, EM, x, u, kind, t);
  xi4[i] = make_float4(x[0], x[1], x[2], X.w);
  pxi4[i] = make_float4(u[0], u[1], u[2], U.w);

  Walker<DIM, DEPOSIT> w;
  float val[12];
  int ci[3];
  bool more = w.first(G.pc, t, U.w, ci, val);
  leaf_to_global<DIM>(G, F, ci, val);
  while (more) {
    more = w.next(G.pc, U.w, ci, val);
    leaf_to_global<DIM>(G, F, ci, val);
  }
}

// Current of trajectories handed over by the host (BoundaryInjector::inject,
// boundary_injector.hxx:136-147: current.calc_j(J, xm, xp, lf, lg, qni_wni, v) for every
// particle that entered through the wall): one thread per trajectory, the same walk and
// the same leaf values as the pusher's, added to J with global atomics.
template <int DIM, int DEPOSIT>
__global__ void __launch_bounds__(128)
  k_deposit_paths(GridDev G, uint32_t n, const psc_b200_jpath* __restrict__ paths, float* __restrict__ flds,
                  long slot_len)
{
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) {
    return;
  }
  const psc_b200_jpath P = paths[i];
  if ((unsigned)P.patch >= (unsigned)G.n_patches) {
    return; // (the host entry has checked the list: never taken)
  }
  float* F = flds + P.patch * slot_len;
  pm::Trajectory t;
#pragma unroll
  for (int d = 0; d < 3; d++) {
    t.xm[d] = P.xm[d];
    t.xp[d] = P.xp[d];
    t.v[d] = P.v[d];
    t.lg[d] = P.lg[d];
    t.lf[d] = pm::fint(P.xp[d]);
  }
  Walker<DIM, DEPOSIT> w;
  float val[12];
  int ci[3];
  bool more = w.first(G.pc, t, P.qni_wni, ci, val);
  leaf_to_global<DIM>(G, F, ci, val);
  while (more) {
    more = w.next(G.pc, P.qni_wni, ci, val);
    leaf_to_global<DIM>(G, F, ci, val);
  }
}

// ---------------------------------------------------------------- tiled kernel

// mbarrier / bulk-copy PTX (sm_90+): one thread arms the barrier with the byte count,
// one warp issues a cp.async.bulk per contiguous row and polls the phase.
__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned phase)
{
  asm volatile("{\n"
               ".reg .pred p;\n"
               "WAIT_%=:\n"
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
               "@p bra DONE_%=;\n"
               "bra WAIT_%=;\n"
               "DONE_%=:\n"
               "}" ::"r"(smem_u32(bar)),
               "r"(phase)
               : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
                 "r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// Ampere-style async copy: 16 bytes per lane, global -> shared, bypassing L1 and registers
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src)
{
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit()
{
  asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void cp_async_wait_all()
{
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr)
{
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(addr)
               : "memory");
  return v;
}

// sum N values across the warp with a transposing butterfly; afterwards every lane
// holds in v[0] the warp total of value slot_of_lane<N>(lane)
template <int N>
__device__ __forceinline__ void warp_transpose_reduce(float (&v)[N], int lane)
{
  static_assert(N == 16 || N == 8, "");
  if (N == 16) {
    bool up = lane & 16;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      float send = up ? v[j] : v[j + 8];
      float keep = up ? v[j + 8] : v[j];
      v[j] = keep + __shfl_xor_sync(FULL, send, 16);
    }
  }
  constexpr int S = (N == 16) ? 8 : 16; // xor mask of the 8 -> 4 step
  {
    bool up = lane & S;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      float send = up ? v[j] : v[j + 4];
      float keep = up ? v[j + 4] : v[j];
      v[j] = keep + __shfl_xor_sync(FULL, send, S);
    }
  }
  {
    bool up = lane & (S / 2);
#pragma unroll
    for (int j = 0; j < 2; j++) {
      float send = up ? v[j] : v[j + 2];
      float keep = up ? v[j + 2] : v[j];
      v[j] = keep + __shfl_xor_sync(FULL, send, S / 2);
    }
  }
  {
    bool up = lane & (S / 4);
    float send = up ? v[0] : v[1];
    float keep = up ? v[1] : v[0];
    v[0] = keep + __shfl_xor_sync(FULL, send, S / 4);
  }
#pragma unroll
  for (int M = S / 8; M >= 1; M >>= 1) {
    v[0] += __shfl_xor_sync(FULL, v[0], M);
  }
}

template <int N>
__device__ __forceinline__ int slot_of_lane(int lane)
{
  if (N == 16) {
    return ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
  }
  return ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
}

struct PushArgs
{
  const uint32_t* cell_off;
  float4* xi4;
  float4* pxi4;
  float* flds;
  long slot_len;
  // counting hook (COUNT): cnt[class][cell] planes (every entry is written exactly once,
  // no memset needed), flags[0] = precondition broken, flags[1] = dropped particles,
  // flags[2] = particles leaving for another rank
  cnt_t* cnt;
  uint32_t* flags;
  uint32_t nct;
  int same_dxi; // 1/float(dx) == float(dx_inv) bitwise: the pusher's cell = the indexer's cell
  FsTables tab;
  // lean kernel, SAME, multi-rank: the drain lists the particles that leave for another rank
  // (group key, index) while it classifies them; flags[2] is the list's fill count.  rem_cap
  // = 0: no list (fused_sort.cu collects them from the boundary cells instead).
  uint32_t* rem_key;
  uint32_t* rem_idx;
  uint32_t rem_cap;
  // lean kernel, PULL (push_lean.cuh): the store the stayers are pulled from (cell_off describes
  // it; xi4 / pxi4 are the output store then), the output store's cell offsets, per cell
  // {arrivals in front of the stayers, stayers}, and the list of the particles that change
  // cell in this push: index in the output store, (plane entry = class * nct + source cell, rank
  // inside that group); flags[3] is the list's fill count
  const float4* xin4;
  const float4* pin4;
  const uint32_t* out_off;
  const uint2* stay;
  uint32_t* mv_idx;
  uint2* mv_key;
  uint32_t mv_cap;
  GapPush gap; // GAP variant only (gap.cuh)
};

// first step of the lane-parallel search for the cell of a virtual particle index
template <typename GEO, int DIM>
struct RowSearch
{
  static constexpr int top = 16;
};
template <int DIM>
struct RowSearch<GeoStatic<DIM>, DIM>
{
  static constexpr int top = GeoStatic<DIM>::t(DIM == pm::DIM_XYZ ? 0 : 1) / 2;
};

// per-lane deposit of one leaf into the shared J tile (or global when outside it)
template <int DIM, typename GEO>
__device__ __forceinline__ void leaf_deposit(const GridDev& G, const GEO& geo, float* sJ, float* F,
                                             int n0, int n1, int n2, const int ci[3],
                                             const float* val)
{
  constexpr int NV = pm::LeafShape<DIM>::NV;
  int r0 = ci[0] - n0, r1 = ci[1] - n1, r2 = ci[2] - n2;
  bool in_tile = (DIM == pm::DIM_YZ || (unsigned)r0 < (unsigned)(geo.f(0) - 1)) &&
                 (unsigned)r1 < (unsigned)(geo.f(1) - 1) && (unsigned)r2 < (unsigned)(geo.f(2) - 1);
  if (in_tile) {
    float* b = sJ + r2 * geo.sz() + r1 * geo.sy() + r0;
#pragma unroll
    for (int n = 0; n < NV; n++) {
      atomicAdd(b + leaf_lin<DIM>(n, geo.sy(), geo.sz(), geo.sm()), val[n]);
    }
  } else {
    leaf_to_global<DIM>(G, F, ci, val);
  }
}

// The store is ordered by (patch, cell), so a warp that walks a row of cells sees the
// particles of one cell after the other.  Work is issued in chunks of 32 consecutive
// particles (full lanes, coalesced 128-bit loads/stores) regardless of where the cell
// boundaries fall; the deposit of the particles that stay in their cell (one trajectory
// segment, ~95 %) is accumulated per lane in registers and summed across the warp once
// per CELL (transposing shuffle butterfly, then one shared-memory atomic per value), not
// once per chunk.  A chunk that straddles a cell boundary is visited in one pass per
// cell; the cell a pass belongs to is warp-uniform, so the source cell of every particle
// is known without computing it, and so is its destination class (COUNT) in the common
// case.  Particles that leave their cell are parked in a per-warp shared queue and
// split/deposited 32 at a time, so the divergent Villasenor-Buneman walk runs with full
// warps.
//
// GAP (gap.cuh): the rows are read through per-cell runs (start, n) instead of contiguous
// offsets, nothing is written in place: stayers go to their final slot of the other
// buffer, movers (boundary fix-ups applied) to the tagged mover list.
template <int DIM, int DEPOSIT, typename GEO, bool TMA, bool COUNT, bool GAP, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) k_push_tiled(GridDev G, GEO geo, PushArgs A)
{
  static_assert(!GAP || COUNT, "the gapped store needs the destination counts");
  constexpr int NV = pm::LeafShape<DIM>::NV;
  constexpr int NVP = (DIM == pm::DIM_XYZ) ? 16 : 8;
  constexpr bool XYZ = DIM == pm::DIM_XYZ;
  extern __shared__ __align__(128) float smem[];
  __shared__ uint64_t bar;
  const int nodes = geo.sm();
  float* sEM = smem;             // [6][f2][f1][f0]
  float* sJ = smem + 6 * nodes;  // [3][f2][f1][f0]
  constexpr int QC = qcap<GAP>();
  float4* sQ = reinterpret_cast<float4*>(smem + ((9 * nodes + 3) & ~3)); // [warps][QC][2]
  float4* sP = sQ + (size_t)(blockDim.x >> 5) * QC * 2;                  // [warps][2][32] next chunk
  __shared__ int row_ctr; // rows are handed out dynamically (balances the warps)

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, n_warps = blockDim.x >> 5;
  const int tiles_per_patch = geo.nt(0) * geo.nt(1) * geo.nt(2);
  const int p = blockIdx.x / tiles_per_patch;
  int tt = blockIdx.x - p * tiles_per_patch;
  int o[3], e[3];
  o[0] = (tt % geo.nt(0)) * geo.t(0);
  o[1] = ((tt / geo.nt(0)) % geo.nt(1)) * geo.t(1);
  o[2] = (tt / (geo.nt(0) * geo.nt(1))) * geo.t(2);
#pragma unroll
  for (int d = 0; d < 3; d++) {
    e[d] = min(geo.t(d), G.ldims[d] - o[d]);
  }
  float* F = A.flds + p * A.slot_len;
  // global index of tile node 0
  const int n0 = o[0] - geo.g(0), n1 = o[1] - geo.g(1), n2 = o[2] - geo.g(2);

  if (tid == 0) {
    row_ctr = n_warps;
  }
  // ---- stage E/B, zero J
  if (TMA) {
    // rows are contiguous in the first non-invariant direction: one bulk copy per row.
    // The host selects this path only when every row start and size is a 16-byte multiple
    // and lies inside the patch array.
    const int row_len = XYZ ? geo.f(0) : geo.f(1);
    const int rows_per_comp = XYZ ? geo.f(1) * geo.f(2) : geo.f(2);
    const int rows = 6 * rows_per_comp;
    const unsigned row_bytes = (unsigned)row_len * 4u;
    if (tid == 0) {
      mbar_init(&bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp == 0) {
      if (lane == 0) {
        mbar_expect_tx(&bar, rows * row_bytes);
      }
      __syncwarp();
      for (int r = lane; r < rows; r += 32) {
        int m = r / rows_per_comp;
        int rem = r - m * rows_per_comp;
        int kz = XYZ ? rem / geo.f(1) : rem;
        int ky = XYZ ? rem - kz * geo.f(1) : 0;
        const float* src = F + fld_off(G, pm::EX + m, n0, n1 + ky, n2 + kz);
        bulk_g2s(sEM + m * nodes + kz * geo.sz() + ky * geo.sy(), src, row_bytes, &bar);
      }
    }
    for (int idx = tid; idx < 3 * nodes; idx += blockDim.x) {
      sJ[idx] = 0.f;
    }
    // one warp polls the mbarrier (512 spinning threads cost 10 % of the issue slots,
    // profiles/r01_v1_push_tiled_ncu.txt); the block barrier publishes the tile
    if (warp == 0) {
      mbar_wait(&bar, 0);
    }
    __syncthreads();
  } else {
    for (int idx = tid; idx < 6 * nodes; idx += blockDim.x) {
      int m = idx / nodes;
      int rem = idx - m * nodes;
      int kz = rem / geo.sz();
      rem -= kz * geo.sz();
      int ky = rem / geo.sy(), kx = rem - ky * geo.sy();
      int gi = n0 + kx, gj = n1 + ky, gk = n2 + kz;
      float v = 0.f;
      if (gi < G.ldims[0] + G.ibn[0] && gj < G.ldims[1] + G.ibn[1] && gk < G.ldims[2] + G.ibn[2]) {
        v = __ldg(F + fld_off(G, pm::EX + m, gi, gj, gk));
      }
      sEM[idx] = v;
    }
    for (int idx = tid; idx < 3 * nodes; idx += blockDim.x) {
      sJ[idx] = 0.f;
    }
    __syncthreads();
  }

  FldTile<GEO> EM{sEM, geo, n0, n1, n2};
  float4* myQ = sQ + (size_t)warp * QC * 2;
  // GAP: what this warp's row sent where, [cell][class] (written to the count planes per row)
  [[maybe_unused]] uint16_t* const myT =
    reinterpret_cast<uint16_t*>(sP + (size_t)n_warps * 64) + (size_t)warp * ct_entries(geo.t(XYZ ? 0 : 1));
  const uint32_t myP = smem_u32(sP + (size_t)warp * 64 + lane);
  int qn = 0; // queued trajectories of this warp (warp-uniform)
  uint32_t mb_next = 0, mb_end = 0; // GAP: this warp's batch of mover slots (warp-uniform)

  // split + deposit `cnt` queued trajectories (entries [qn - cnt, qn)), one per lane
  auto drain = [&](int cnt) {
    const bool a2 = lane < cnt;
    Walker<DIM, DEPOSIT> w;
    float val[NVP];
    int ci[3] = {0, 0, 0};
    bool more = false;
    float qw = 0.f;
    if (a2) {
      float4 A0 = myQ[2 * (qn - cnt + lane)], A1 = myQ[2 * (qn - cnt + lane) + 1];
      pm::Trajectory t;
      t.xm[0] = A0.x, t.xm[1] = A0.y, t.xm[2] = A0.z;
      t.xp[0] = A1.x, t.xp[1] = A1.y, t.xp[2] = A1.z;
      t.v[0] = A1.w, t.v[1] = 0.f, t.v[2] = 0.f;
      qw = A0.w;
#pragma unroll
      for (int d = 0; d < 3; d++) {
        t.lg[d] = pm::fint(t.xm[d]);
        t.lf[d] = pm::fint(t.xp[d]);
      }
      more = w.first(G.pc, t, qw, ci, val);
      leaf_deposit<DIM>(G, geo, sJ, F, n0, n1, n2, ci, val);
    }
    while (__any_sync(FULL, more)) {
      if (more) {
        more = w.next(G.pc, qw, ci, val);
        leaf_deposit<DIM>(G, geo, sJ, F, n0, n1, n2, ci, val);
      }
    }
    qn -= cnt;
    __syncwarp();
  };

  // ---- particle runs: rows of cells along the first non-invariant dim
  const int n_rows = XYZ ? e[1] * e[2] : e[2];
  const int run_cells = XYZ ? e[0] : e[1]; // <= 31 (host)
  const uint32_t* coff = A.cell_off + (size_t)p * G.n_cells;
  for (int row = warp; row < n_rows;) {
    int c0, rs1, rs2; // first cell of the row; row coordinates
    if (XYZ) {
      int ry = row % e[1], rz = row / e[1];
      rs1 = o[1] + ry, rs2 = o[2] + rz;
      c0 = (rs2 * G.ldims[1] + rs1) * G.ldims[0] + o[0];
    } else {
      rs1 = o[1], rs2 = o[2] + row;
      c0 = rs2 * G.ldims[1] + o[1];
    }
    // lane j holds the offset of the row's j-th cell boundary (GAP: inside the row's
    // virtual sequence = its cells' runs back to back; my_st = where run j really starts,
    // my_vo = where the stayers of cell j are written)
    uint32_t myoff, my_st = 0, my_vo = 0;
    if constexpr (GAP) {
      const size_t gc = (size_t)p * G.n_cells + (size_t)(c0 + min(lane, run_cells - 1));
      const uint32_t nj = lane < run_cells ? __ldg(&A.gap.in_n[gc]) : 0u;
      for (int k = lane; k < run_cells * CT; k += 32) {
        myT[k] = 0;
      }
      __syncwarp();
      my_st = __ldg(&A.gap.in_start[gc]);
      my_vo = __ldg(&A.gap.out_v[gc]) + A.gap.rl;
      uint32_t incl = nj;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) {
          incl += t;
        }
      }
      myoff = incl - nj;
    } else {
      myoff = __ldg(&coff[c0 + min(lane, run_cells)]);
    }
    // GAP: cell (of the row) that holds record iv of the row's virtual sequence
    auto cell_of = [&](uint32_t iv) -> int {
      int j = 0;
#pragma unroll
      for (int st = RowSearch<GEO, DIM>::top; st >= 1; st >>= 1) {
        const int jj = j + st;
        const uint32_t t = __shfl_sync(FULL, myoff, jj & 31);
        if (jj < run_cells && iv >= t) {
          j = jj;
        }
      }
      return j;
    };
    // ... and its index into the store being read
    auto src_of = [&](uint32_t iv, int j) -> uint32_t {
      return __shfl_sync(FULL, my_st, j) + (iv - __shfl_sync(FULL, myoff, j));
    };
    const float4* const src_x = GAP ? A.gap.in_x : A.xi4;
    const float4* const src_p = GAP ? A.gap.in_p : A.pxi4;
    const uint32_t begin = __shfl_sync(FULL, myoff, 0), end = __shfl_sync(FULL, myoff, run_cells);
    int cur = 0;                                     // cell of the row the passes are at
    uint32_t cb = begin, ce = __shfl_sync(FULL, myoff, 1); // its particle range
    float acc[NV];                                   // this lane's share of the cell's deposit
#pragma unroll
    for (int n = 0; n < NV; n++) {
      acc[n] = 0.f;
    }
    bool dirty = false;
    uint32_t mycount = 0; // lane k: particles of the current cell in class k

    // the next chunk travels global -> shared with cp.async while this one is computed
    // (no registers held across the chunk body)
    int jc = 0;               // GAP: cell of this lane's record in the next chunk to be computed
    uint32_t carry_leave = 0; // GAP: particles that left the cell which continues into the next chunk
    {
      if constexpr (GAP) {
        jc = cell_of(begin + lane);
      }
      const uint32_t a0 = GAP ? src_of(begin + lane, jc) : begin + lane;
      if (begin + lane < end) {
        cp_async16(myP, src_x + a0);
        cp_async16(myP + 32 * sizeof(float4), src_p + a0);
      }
    }
    cp_async_commit();
    uint32_t base = begin;
    do {
      const uint32_t i = base + lane;
      const bool act = i < end;
      cp_async_wait_all();
      float4 X = lds128(myP), U = lds128(myP + 32 * sizeof(float4));
      const int jcur = jc; // GAP: cell of this lane's record
      {
        if constexpr (GAP) {
          jc = cell_of(i + 32);
        }
        const uint32_t a1 = GAP ? src_of(i + 32, jc) : i + 32;
        if (i + 32 < end) {
          cp_async16(myP, src_x + a1);
          cp_async16(myP + 32 * sizeof(float4), src_p + a1);
        }
      }
      cp_async_commit();
      if (qn > QC - 32) {
        drain(min(qn, 32));
      }
      float val[NV];
      int ci[3] = {0, 0, 0};
      int pos[3] = {0, 0, 0};
      bool single = false; // one trajectory segment, leaf in (ci, val)
      [[maybe_unused]] float sxm[3] = {0.f, 0.f, 0.f}, sxp[3] = {0.f, 0.f, 0.f}, sv0 = 0.f, sqw = 0.f;
      {
        bool cross = false;
        pm::Trajectory t;
        if (act) {
          float x[3] = {X.x, X.y, X.z}, u[3] = {U.x, U.y, U.z};
          pm::advance<DIM>(G.pc, EM, x, u, __float_as_int(X.w), t);
          if constexpr (GAP) {
            X = make_float4(x[0], x[1], x[2], X.w);
            U = make_float4(u[0], u[1], u[2], U.w);
          } else {
            A.xi4[i] = make_float4(x[0], x[1], x[2], X.w);
            A.pxi4[i] = make_float4(u[0], u[1], u[2], U.w);
          }
          cross = (XYZ && t.lf[0] != t.lg[0]) || t.lf[1] != t.lg[1] || t.lf[2] != t.lg[2];
          single = !cross;
          if constexpr (!GAP) {
            if (single) {
              Walker<DIM, DEPOSIT> w;
              w.first(G.pc, t, U.w, ci, val);
            }
          } else {
            // the leaf is computed once the record has left (below): its values and the
            // record are never live together
#pragma unroll
            for (int d = 0; d < 3; d++) {
              sxm[d] = t.xm[d], sxp[d] = t.xp[d];
            }
            sv0 = t.v[0], sqw = U.w;
          }
          if (COUNT) {
#pragma unroll
            for (int d = 0; d < 3; d++) {
              pos[d] = A.same_dxi ? t.lf[d] : pm::cell_position(G.pc, x[d], d);
            }
          }
        }
        // park cell-crossing particles for the split/deposit walk
        const unsigned cm = __ballot_sync(FULL, cross);
        if (cm) {
          if (cross) {
            const int slot = qn + __popc(cm & ((1u << lane) - 1u));
            myQ[2 * slot] = make_float4(t.xm[0], t.xm[1], t.xm[2], U.w);
            myQ[2 * slot + 1] = make_float4(t.xp[0], t.xp[1], t.xp[2], t.v[0]);
          }
          qn += __popc(cm);
          __syncwarp();
        }
      }
      int gcls = CLS_NONE; // GAP: destination class, known ahead of the passes
      if constexpr (GAP) {
        // Every lane knows its own source cell, so the record leaves right here: a stayer
        // goes to its final slot = (index inside the cell) - (particles of the cell that left
        // before it), a mover to the tagged list.  X, U, pos are dead before the passes.
        const unsigned lt = (1u << lane) - 1u;
        const uint32_t cstart = __shfl_sync(FULL, myoff, jcur); // where my cell starts in the row
        const uint32_t vo = __shfl_sync(FULL, my_vo, jcur);
        uint32_t tcell = 0;
        if (act) {
          const int s0 = XYZ ? o[0] + jcur : 0, s1 = XYZ ? rs1 : rs1 + jcur, s2 = rs2;
          const int d0 = pos[0] - s0, d1 = pos[1] - s1, d2 = pos[2] - s2;
          const bool ok = (unsigned)pos[0] < (unsigned)G.ldims[0] && (unsigned)pos[1] < (unsigned)G.ldims[1] &&
                          (unsigned)pos[2] < (unsigned)G.ldims[2] && (unsigned)(d0 + 1) <= 2u &&
                          (unsigned)(d1 + 1) <= 2u && (unsigned)(d2 + 1) <= 2u;
          if (ok) {
            gcls = ((d2 + 1) * 3 + d1 + 1) * 3 + d0 + 1;
            tcell = (uint32_t)p * G.n_cells + (uint32_t)((pos[2] * G.ldims[1] + pos[1]) * G.ldims[0] + pos[0]);
          } else {
            // patch boundary: the record leaves with the boundary fix-ups applied
            float xx[3] = {X.x, X.y, X.z}, uu[3] = {U.x, U.y, U.z};
            int q = 0, c = 0;
            gcls = fs_classify(G, A.tab, p, s0, s1, s2, xx, uu, q, c);
            X = make_float4(xx[0], xx[1], xx[2], X.w);
            U = make_float4(uu[0], uu[1], uu[2], U.w);
            tcell = (uint32_t)q * G.n_cells + (uint32_t)c;
          }
        }
        // lanes of my cell below me; cont: my cell began in an earlier chunk
        const unsigned below = lt & ~((1u << min(31u, cstart > base ? cstart - base : 0u)) - 1u);
        const bool cont = cstart < base;
        const unsigned leave = __ballot_sync(FULL, act && gcls != CLS_CENTER);
        if (act && gcls == CLS_CENTER) {
          const uint32_t dst = vo + (i - cstart) - (__popc(leave & below) + (cont ? carry_leave : 0u));
          A.gap.out_x[dst] = X;
          A.gap.out_p[dst] = U;
        }
        {
          // the cell of the chunk's last record may continue
          const uint32_t cl = __shfl_sync(FULL, cstart, 31);
          const unsigned from = ~((1u << min(31u, cl > base ? cl - base : 0u)) - 1u);
          carry_leave = __popc(leave & from) + (cl < base ? carry_leave : 0u);
        }
        // whatever left its cell is counted per (cell, class) in the warp's table; the old
        // value is the rank of the group's first member (groups keep the old order)
        if (leave) {
          const bool lv = act && gcls != CLS_CENTER;
          const unsigned same = __match_any_sync(FULL, lv ? ((jcur << 5) | gcls) : 0xffff);
          uint32_t r0 = 0;
          if (lv) {
            r0 = myT[jcur * CT + gcls];
          }
          __syncwarp();
          if (lv && (same & lt) == 0) {
            myT[jcur * CT + gcls] = (uint16_t)(r0 + __popc(same));
          }
          __syncwarp();
          // movers: parked in the tagged list, slots handed out GAP_BATCH at a time
          const bool mover = lv && gcls < FS_PLANES;
          const unsigned mm = __ballot_sync(FULL, mover);
          if (mm) {
            const uint32_t nm = __popc(mm), room = mb_end - mb_next;
            uint32_t nb = 0;
            if (nm > room) {
              if (lane == 0) {
                nb = atomicAdd(&A.gap.ctl[GAP_CTL_MOVERS], (uint32_t)GAP_BATCH);
              }
              nb = __shfl_sync(FULL, nb, 0);
            }
            if (mover) {
              const uint32_t r = __popc(mm & lt);
              const uint32_t slot = r < room ? mb_next + r : nb + (r - room);
              if (slot < A.gap.m_cap) {
                A.gap.mx[slot] = X;
                A.gap.mp[slot] = U;
                A.gap.mtag[slot] =
                  make_uint4(tcell, (uint32_t)gcls * A.nct + (uint32_t)p * G.n_cells + (uint32_t)(c0 + jcur),
                             r0 + __popc(same & lt), 1u);
              } else {
                atomicExch(&A.gap.ctl[GAP_CTL_M_FULL], 1u);
              }
            }
            if (nm > room) {
              mb_next = nb + (nm - room);
              mb_end = nb + GAP_BATCH;
            } else {
              mb_next += nm;
            }
          }
        }
        if (single) {
          pm::Trajectory t;
#pragma unroll
          for (int d = 0; d < 3; d++) {
            t.xm[d] = sxm[d], t.xp[d] = sxp[d];
            t.lg[d] = pm::fint(sxm[d]), t.lf[d] = pm::fint(sxp[d]);
            t.v[d] = 0.f;
          }
          t.v[0] = sv0;
          Walker<DIM, DEPOSIT> w;
          w.first(G.pc, t, sqw, ci, val);
          const int s0 = XYZ ? o[0] + jcur : 0, s1 = XYZ ? rs1 : rs1 + jcur, s2 = rs2;
          if (ci[0] != s0 || ci[1] != s1 || ci[2] != s2) {
            // the leaf is not the run's cell (1/float(dx) and float(dx_inv) disagree at a
            // cell edge): deposit it on its own
            leaf_deposit<DIM>(G, geo, sJ, F, n0, n1, n2, ci, val);
            single = false;
          }
        }
      }
      // ---- one pass per cell that has particles in this chunk
      for (;;) {
        const bool mine = act && i >= cb && i < ce;
        const int s0 = XYZ ? o[0] + cur : 0, s1 = XYZ ? rs1 : rs1 + cur, s2 = rs2;
        if (mine && single) {
          if (GAP || (ci[0] == s0 && ci[1] == s1 && ci[2] == s2)) {
#pragma unroll
            for (int n = 0; n < NV; n++) {
              acc[n] += val[n];
            }
            dirty = true;
          } else {
            // the leaf is not the run's cell (1/float(dx) and float(dx_inv) disagree at a
            // cell edge): deposit it on its own
            leaf_deposit<DIM>(G, geo, sJ, F, n0, n1, n2, ci, val);
          }
        }
        if (COUNT && !GAP) {
          int cls = CLS_NONE;
          if (mine) {
            const int d0 = pos[0] - s0, d1 = pos[1] - s1, d2 = pos[2] - s2;
            const bool ok = (unsigned)pos[0] < (unsigned)G.ldims[0] && (unsigned)pos[1] < (unsigned)G.ldims[1] &&
                            (unsigned)pos[2] < (unsigned)G.ldims[2] && (unsigned)(d0 + 1) <= 2u &&
                            (unsigned)(d1 + 1) <= 2u && (unsigned)(d2 + 1) <= 2u;
            if (ok) {
              cls = ((d2 + 1) * 3 + d1 + 1) * 3 + d0 + 1;
            } else {
              // patch boundary: the pushed record is re-read (written by this thread above)
              const float4 Xr = A.xi4[i], Ur = A.pxi4[i];
              float xx[3] = {Xr.x, Xr.y, Xr.z}, uu[3] = {Ur.x, Ur.y, Ur.z};
              int q, c;
              cls = fs_classify(G, A.tab, p, s0, s1, s2, xx, uu, q, c);
            }
          }
          unsigned grp = __ballot_sync(FULL, cls == CLS_CENTER);
          if (lane == CLS_CENTER) {
            mycount += __popc(grp);
          }
          unsigned rem = __ballot_sync(FULL, mine) & ~grp;
          while (rem) {
            const int v = __shfl_sync(FULL, cls, __ffs(rem) - 1);
            grp = __ballot_sync(FULL, cls == v);
            if (lane == v) {
              mycount += __popc(grp);
            }
            rem &= ~grp;
          }
        }
        if (ce > base + 32) {
          break; // the cell continues in the next chunk
        }
        // ---- the cell is complete: flush its deposit and its class counts
        if (__any_sync(FULL, dirty)) {
          float v[NVP];
#pragma unroll
          for (int n = 0; n < NVP; n++) {
            v[n] = n < NV ? acc[n] : 0.f;
          }
          warp_transpose_reduce<NVP>(v, lane);
          const int my_slot = slot_of_lane<NVP>(lane);
          const int my_lin = (my_slot < NV) ? leaf_lin<DIM>(my_slot, geo.sy(), geo.sz(), geo.sm()) : 0;
          const bool writer = (my_slot < NV) && ((lane & (NVP == 16 ? 1 : 3)) == 0);
          if (writer) {
            atomicAdd(&sJ[(s2 - n2) * geo.sz() + (s1 - n1) * geo.sy() + (s0 - n0) + my_lin], v[0]);
          }
#pragma unroll
          for (int n = 0; n < NV; n++) {
            acc[n] = 0.f;
          }
          dirty = false;
        }
        if (COUNT && !GAP) {
          if (lane < FS_PLANES) {
            A.cnt[(size_t)lane * A.nct + (size_t)p * G.n_cells + (size_t)(c0 + cur)] = (cnt_t)mycount;
            if (mycount > CNT_MAX) {
              atomicExch(&A.flags[0], 1u);
            }
          } else if (mycount) {
            if (lane == CLS_BAD) {
              atomicExch(&A.flags[0], 1u);
            } else if (lane == CLS_DROP) {
              atomicAdd(&A.flags[1], mycount);
            } else if (lane == CLS_REMOTE) {
              atomicAdd(&A.flags[2], mycount);
            }
          }
          mycount = 0;
        }
        if (++cur == run_cells) {
          break;
        }
        cb = ce;
        ce = __shfl_sync(FULL, myoff, cur + 1);
      }
      base += 32;
    } while (base < end);
    if constexpr (GAP) {
      // the row's counts: lane k writes class k of every cell; stayers = population - leavers
      __syncwarp();
      uint32_t special = 0;
      for (int j = 0; j < run_cells; j++) {
        const uint32_t v = lane < CT ? myT[j * CT + lane] : 0u;
        const uint32_t left = __reduce_add_sync(FULL, lane == CLS_CENTER ? 0u : v);
        const uint32_t nj = __shfl_sync(FULL, myoff, j + 1) - __shfl_sync(FULL, myoff, j);
        if (lane < FS_PLANES) {
          const uint32_t cv = lane == CLS_CENTER ? nj - left : v;
          A.cnt[(size_t)lane * A.nct + (size_t)p * G.n_cells + (size_t)(c0 + j)] = (cnt_t)cv;
          if (cv > CNT_MAX) {
            atomicExch(&A.flags[0], 1u);
          }
        } else {
          special += v;
        }
      }
      if (special) {
        if (lane == CLS_BAD) {
          atomicExch(&A.flags[0], 1u);
        } else if (lane == CLS_DROP) {
          atomicAdd(&A.flags[1], special);
        } else if (lane == CLS_REMOTE) {
          atomicAdd(&A.flags[2], special);
        }
      }
      __syncwarp();
    }
    if (lane == 0) {
      row = atomicAdd(&row_ctr, 1);
    }
    row = __shfl_sync(FULL, row, 0);
  }
  while (qn > 0) {
    drain(min(qn, 32));
  }
  if constexpr (GAP) {
    // slots of the last batch that were not used
    for (uint32_t sl = mb_next + lane; sl < mb_end && sl < A.gap.m_cap; sl += 32) {
      A.gap.mtag[sl] = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  __syncthreads();

  // ---- flush the J tile (halo included) with global reductions
  for (int idx = tid; idx < 3 * nodes; idx += blockDim.x) {
    float v = sJ[idx];
    if (v != 0.f) {
      int m = idx / nodes;
      int rem = idx - m * nodes;
      int kz = rem / geo.sz();
      rem -= kz * geo.sz();
      int ky = rem / geo.sy(), kx = rem - ky * geo.sy();
      int gi = n0 + kx, gj = n1 + ky, gk = n2 + kz;
      if (gi < G.ldims[0] + G.ibn[0] && gj < G.ldims[1] + G.ibn[1] && gk < G.ldims[2] + G.ibn[2]) {
        atomicAdd(F + fld_off(G, m, gi, gj, gk), v);
      }
    }
  }
}

#include "push_lean.cuh"
#include "push_lean_pull.cuh"

// ---------------------------------------------------------------- host side

// k_push_lean: compile-time geometry, tensor-map TMA, no gapped store
template <int DIM, int DEPOSIT>
static int launch_lean(Ctx* c, const GeoStatic<DIM>& geo, bool count, PushArgs A)
{
  const GridDev& G = c->gd;
  constexpr bool XYZ = DIM == pm::DIM_XYZ;
  if (count && c->comm && A.same_dxi && c->opt_push_collect) {
    // room for half as many again as left last step; a step that overflows it falls back to
    // the boundary-cell collection
    const size_t cap = std::max<size_t>(1u << 16, (size_t)c->last_n_rem + c->last_n_rem / 2);
    PSC_TRY(c->scr[7].reserve(4 * cap * sizeof(uint32_t)));
    A.rem_idx = c->scr[7].as<uint32_t>();
    A.rem_key = A.rem_idx + cap;
    A.rem_cap = (uint32_t)cap;
    c->rem_cap = A.rem_cap;
  }
  const int tiles = geo.nt(0) * geo.nt(1) * geo.nt(2) * G.n_patches;
  // particles per lane.  2 = the update on the packed FP32 pipe (FADD2 / FFMA2): 18 % fewer
  // warp instructions, but 128 registers => 16 warps per SM instead of 24, and the kernel
  // loses more to latency than it gains in issue slots (S3D: 18.9 ms against 17.7 ms,
  // DESIGN.md 3.1): opt-in.  The FMA build always takes W = 1 (ptxas contracts the packed
  // products on its own terms, which leaves the 4-ULP contract).
#ifdef PM_FAST_MATH
  const int W = 1;
#else
  const int W = c->opt_lean >= 2 ? 2 : 1;
#endif
  // tiles + per-warp queue + staging (W = 1: the ring of n_stages chunks and its alignment slack)
  const size_t ring = (size_t)lean::n_stages<DIM>() * lean::STAGE_BYTES;
  const size_t smem_bytes =
    (size_t)((9 * geo.sm() + 3) & ~3) * sizeof(float) +
    (W == 1 ? (size_t)lean::n_warps<1>() * (lean::qcap<1>() * 2 * sizeof(float4) + ring) + ring
            : (size_t)lean::n_warps<2>() * (lean::qcap<2>() * 2 + 64 * W) * sizeof(float4));
  TensorMap128 tm128;
  const int box[4] = {XYZ ? geo.f(0) : geo.f(1), XYZ ? geo.f(1) : geo.f(2), XYZ ? geo.f(2) : 6, 6};
  PSC_TRY(field_tile_tensor_map(c, 0, XYZ ? 4 : 3, box, &tm128));
  CUtensorMap tm;
  static_assert(sizeof(tm) == sizeof(tm128), "");
  memcpy(&tm, &tm128, sizeof(tm));
  if (c->pull_pending) {
    // the mover list: room for a quarter of the store, and at least
    // twice what moved last step; a step that overflows it takes the full scatter instead
    size_t cap = std::min<size_t>(0xffffffffu, std::max<size_t>({size_t(1) << 16, (size_t)c->n_prts / 4,
                                                                 2 * (size_t)c->last_n_movers}));
    if (c->opt_pull_cap > 0) {
      cap = (size_t)c->opt_pull_cap;
    }
    PSC_TRY(c->scr[12].reserve(cap * (sizeof(uint32_t) + sizeof(uint2))));
    A.mv_key = c->scr[12].as<uint2>();
    A.mv_idx = reinterpret_cast<uint32_t*>(A.mv_key + cap);
    A.mv_cap = (uint32_t)cap;
    c->mv_cap = A.mv_cap;
  }
  if (c->pull_pending) {
    // (push_dim has checked that this launch is possible: count, same_dxi, W = 1)
    A.xin4 = c->xi();
    A.pin4 = c->pxi();
    A.xi4 = c->xi_alt();
    A.pxi4 = c->pxi_alt();
    A.out_off = c->d_cell_off_alt;
    A.stay = c->scr[13].as<uint2>();
    PSC_CUDA_TRY(cudaMemsetAsync(A.cnt, 0, (size_t)A.nct * FS_PLANES * sizeof(cnt_t), c->stream));
    auto kern = lean::k_push_lean_pull<DIM, DEPOSIT>;
    PSC_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    kern<<<tiles, lean::n_warps<1>() * 32, smem_bytes, c->stream>>>(tm, G, geo, A);
    // the output store is the store now
    c->cur ^= 1;
    std::swap(c->d_cell_off, c->d_cell_off_alt);
    c->pull_pending = false;
    c->pulled = true;
    return 0;
  }
#define PSC_LEAN_W(CN, SM, WW)                                                                    \
  do {                                                                                            \
    auto kern = lean::k_push_lean<DIM, DEPOSIT, CN, SM, WW>;                                      \
    PSC_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,          \
                                      (int)smem_bytes));                                          \
    kern<<<tiles, lean::n_warps<WW>() * 32, smem_bytes, c->stream>>>(tm, G, geo, A);              \
  } while (0)
#define PSC_LEAN(CN, SM)                                                                          \
  do {                                                                                            \
    if (W == 1) {                                                                                 \
      PSC_LEAN_W(CN, SM, 1);                                                                      \
    } else {                                                                                      \
      PSC_LEAN_W(CN, SM, 2);                                                                      \
    }                                                                                             \
  } while (0)
  if (count) {
    // the planes are added to (leavers one by one, stayers once per cell)
    PSC_CUDA_TRY(cudaMemsetAsync(A.cnt, 0, (size_t)A.nct * FS_PLANES * sizeof(cnt_t), c->stream));
    if (A.same_dxi) {
      PSC_LEAN(true, true);
    } else {
      PSC_LEAN(true, false);
    }
  } else {
    if (A.same_dxi) {
      PSC_LEAN(false, true);
    } else {
      PSC_LEAN(false, false);
    }
  }
#undef PSC_LEAN_W
#undef PSC_LEAN
  return 0;
}

template <int DIM, int DEPOSIT, typename GEO, bool TUNE>
static int launch_tiled(Ctx* c, const GEO& geo, bool tma, bool count, bool gap, const PushArgs& A)
{
  const GridDev& G = c->gd;
  int tiles = geo.nt(0) * geo.nt(1) * geo.nt(2) * G.n_patches;
  int threads = std::max(32, std::min(512, c->opt_threads)) & ~31;
  if (geo.t(DIM == pm::DIM_XYZ ? 0 : 1) > 31) {
    return -1; // a row's cell boundaries are held one per lane
  }
  // register budget variants (launch bounds); odd geometries get the roomy one only
  int lb = 1; // 0: (256, 3)  1: (256, 2)  2: (512, 1)  3: (384, 2)  4: (192, 3)  5: (192, 4)
  if (TUNE) {
    lb = threads > 384 ? 2
                       : (threads == 384 ? 3
                                         : (threads == 192 ? (c->opt_min_blocks == 4 ? 5 : 4)
                                                           : (c->opt_min_blocks == 2 ? 1 : 0)));
  } else {
    threads = std::min(threads, 256);
  }
  size_t smem_bytes = (size_t)((9 * geo.sm() + 3) & ~3) * sizeof(float) +
                      (size_t)(threads / 32) * ((gap ? QCAP_GAP : QCAP) * 2 + 64) * sizeof(float4) +
                      (gap ? (size_t)(threads / 32) * ct_entries(geo.t(DIM == pm::DIM_XYZ ? 0 : 1)) * sizeof(uint16_t) : 0);
  if (smem_bytes > 220 * 1024) {
    return -1;
  }
#define PSC_LAUNCH(TM, CN, GP, MT, MB)                                                            \
  do {                                                                                            \
    auto kern = k_push_tiled<DIM, DEPOSIT, GEO, TM, CN, GP, MT, MB>;                              \
    PSC_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,          \
                                      (int)smem_bytes));                                          \
    kern<<<tiles, threads, smem_bytes, c->stream>>>(G, geo, A);                                   \
  } while (0)
#ifdef PUSH_PROBE
  PSC_LAUNCH(true, true, PUSH_PROBE_G, PUSH_PROBE_T, PUSH_PROBE_B);
#else
#define PSC_LAUNCH_LB(TM, CN, GP)                                                                 \
  do {                                                                                            \
    if constexpr (TUNE) {                                                                         \
      if (lb == 0) {                                                                              \
        PSC_LAUNCH(TM, CN, GP, 256, 3);                                                           \
      } else if (lb == 2) {                                                                       \
        PSC_LAUNCH(TM, CN, GP, 512, 1);                                                           \
      } else if (lb == 3) {                                                                       \
        PSC_LAUNCH(TM, CN, GP, 384, 2);                                                           \
      } else if (lb == 4) {                                                                       \
        PSC_LAUNCH(TM, CN, GP, 192, 3);                                                           \
      } else if (lb == 5) {                                                                       \
        PSC_LAUNCH(TM, CN, GP, 192, 4);                                                           \
      } else {                                                                                    \
        PSC_LAUNCH(TM, CN, GP, 256, 2);                                                           \
      }                                                                                           \
    } else {                                                                                      \
      PSC_LAUNCH(TM, CN, GP, 256, 2);                                                             \
    }                                                                                             \
  } while (0)
  if (gap) {
    if (tma) {
      if constexpr (TUNE) {
        PSC_LAUNCH_LB(true, true, true);
      }
    } else {
      PSC_LAUNCH_LB(false, true, true);
    }
  } else if (count) {
    if (tma) {
      if constexpr (TUNE) {
        PSC_LAUNCH_LB(true, true, false);
      }
    } else {
      PSC_LAUNCH_LB(false, true, false);
    }
  } else {
    if (tma) {
      if constexpr (TUNE) {
        PSC_LAUNCH_LB(true, false, false);
      }
    } else {
      PSC_LAUNCH_LB(false, false, false);
    }
  }
#undef PSC_LAUNCH_LB
#endif
#undef PSC_LAUNCH
  return 0;
}

// gap: the gapped-store variant (gap.cuh); the caller (gap_prepare) has laid out the store
// that is written and cleared the control words
template <int DIM, int DEPOSIT>
static int push_dim(Ctx* c, bool gap)
{
  const GridDev& G = c->gd;
  c->counts_valid = false;
  if (c->pull_pending && !(c->sorted && c->opt_tiled && !gap)) {
    PSC_TRY(pull_materialize(c)); // (options changed under a pending pull)
  }
  if (c->n_prts == 0) {
    return gap ? fail("gapped push: empty store") : 0;
  }
  if (gap || (c->sorted && c->opt_tiled)) {
    PushArgs A{};
    A.cell_off = c->d_cell_off;
    A.xi4 = c->xi();
    A.pxi4 = c->pxi();
    A.flds = c->fld(0);
    A.slot_len = c->fld_slot_len(0);
    A.nct = (uint32_t)G.n_cells * G.n_patches;
    A.tab = FsTables{c->d_patch_bnd, c->d_nei_patch};
    A.same_dxi = 1;
    for (int d = 0; d < 3; d++) {
      A.same_dxi = A.same_dxi && G.pc.dxi[d] == G.pc.dxi_idx[d];
    }
    bool count = c->want_counts || gap;
    c->rem_cap = 0;
    if (count) {
      // the kernel writes every cnt[class][cell] entry exactly once: no memset
      PSC_TRY(c->scr[9].reserve((size_t)A.nct * FS_PLANES * sizeof(cnt_t)));
      PSC_TRY(c->scr[11].reserve((G.n_patches + 1 + 4) * sizeof(uint32_t)));
      A.cnt = c->scr[9].as<cnt_t>();
      A.flags = c->scr[11].as<uint32_t>();
      PSC_CUDA_TRY(cudaMemsetAsync(A.flags, 0, 4 * sizeof(uint32_t), c->stream));
    }
    if (gap) {
      GapPush& g = A.gap;
      g.in_start = c->g_start[0];
      g.in_n = c->g_n[0];
      g.in_x = c->xi();
      g.in_p = c->pxi();
      g.out_v = c->g_v;
      g.rl = c->g_rl;
      g.out_x = c->xi_alt();
      g.out_p = c->pxi_alt();
      g.mx = c->mvx;
      g.mp = c->mvp;
      g.mtag = c->mvtag;
      g.ctl = c->g_ctl;
      g.m_cap = (uint32_t)std::min<size_t>(c->mov_cap, 0xffffffffu);
    }
    const bool xyz = DIM == pm::DIM_XYZ;
    bool custom_tile = c->opt_tile[0] > 0 || c->opt_tile[1] > 0 || c->opt_tile[2] > 0;
    GeoStatic<DIM> gs{};
    bool stat = !custom_tile;
    for (int d = 0; d < 3; d++) {
      bool inv = (!xyz && d == 0);
      gs.nt_[d] = inv ? 1 : G.ldims[d] / gs.t(d);
      stat = stat && (inv || (G.ibn[d] == 2 && G.ldims[d] % gs.t(d) == 0));
    }
    int rc = -1;
    // (tensor-map strides are multiples of 16 bytes)
    const bool lean_ok = stat && !gap && c->opt_lean && c->opt_tma && G.im[xyz ? 0 : 1] % 4 == 0;
    // pull mode (push_lean_pull.cuh): this push completes the previous step's sort on its way
    bool pull_now = lean_ok && count && A.same_dxi && c->opt_lean < 2 && pull_possible(c);
#ifdef PM_FAST_MATH
    pull_now = lean_ok && count && A.same_dxi && pull_possible(c);
#endif
    if (c->pull_pending && !pull_now) {
      PSC_TRY(pull_materialize(c));
      A.cell_off = c->d_cell_off;
      A.xi4 = c->xi();
      A.pxi4 = c->pxi();
    } else if (pull_now && !c->pull_pending) {
      PSC_TRY(pull_enter(c));
    }
    if (lean_ok) {
      KernelScope ks(c, "push_lean");
      rc = launch_lean<DIM, DEPOSIT>(c, gs, count, A);
      c->n_lean += rc == 0;
    }
    if (rc == -1 && stat) {
      // bulk copies: contiguous rows (x in 3D, y in yz) must be 16-byte aligned
      int cd = xyz ? 0 : 1;
      bool tma = c->opt_tma && (G.im[cd] % 4 == 0) && (c->fld_slot_len(0) % 4 == 0) &&
                 (G.fld_len % 4 == 0);
      KernelScope ks(c, gap ? "push_gap" : (tma ? "push_tiled_tma" : "push_tiled"));
      rc = launch_tiled<DIM, DEPOSIT, GeoStatic<DIM>, true>(c, gs, tma, count, gap, A);
    }
    if (rc == -1) {
      GeoDyn gd{};
      int def[3] = {xyz ? 8 : 1, xyz ? 8 : 16, xyz ? 8 : 16};
      for (int d = 0; d < 3; d++) {
        bool inv = (!xyz && d == 0);
        int t = c->opt_tile[d] > 0 ? c->opt_tile[d] : def[d];
        gd.t_[d] = inv ? 1 : std::min(t, G.ldims[d]);
        gd.nt_[d] = (G.ldims[d] + gd.t_[d] - 1) / gd.t_[d];
        gd.f_[d] = inv ? 1 : gd.t_[d] + 3;
        gd.g_[d] = inv ? 0 : 1;
      }
      KernelScope ks(c, gap ? "push_gap_dyn" : "push_tiled_dyn");
      rc = launch_tiled<DIM, DEPOSIT, GeoDyn, false>(c, gd, false, count, gap, A);
    }
    if (rc == 0) {
      c->n_launches++;
      c->counts_valid = count;
      return check_launch(c, "push_tiled");
    }
    if (rc > 0) {
      return rc;
    }
    if (gap) {
      return fail("gapped push: no tile geometry fits this grid");
    }
  }
  {
    KernelScope ks(c, "push_general");
    k_push_general<DIM, DEPOSIT><<<div_up(c->n_prts, 256), 256, 0, c->stream>>>(
      G, c->n_prts, c->d_off, c->xi(), c->pxi(), c->fld(0), c->fld_slot_len(0));
    c->n_launches++;
  }
  return check_launch(c, "push_general");
}

template <int DIM, int DEPOSIT>
static int push_dim(Ctx* c)
{
  return push_dim<DIM, DEPOSIT>(c, false);
}

} // namespace PUSH_VARIANT

// gapped store: gap_prepare() has run; gap_finish() follows (capi.cu step)
int PUSH_CAT(push_gap_, PUSH_VARIANT)(Ctx* c)
{
  using namespace PUSH_VARIANT;
  PSC_TRY(flds_zero(c, 0, pm::JXI, pm::JXI + 3));
#ifdef PUSH_PROBE
  return push_dim<pm::DIM_XYZ, pm::DEPOSIT_SPLIT>(c, true);
#else
  if (c->gd.dim == pm::DIM_XYZ) {
    return push_dim<pm::DIM_XYZ, pm::DEPOSIT_SPLIT>(c, true);
  } else if (c->gd.deposit == pm::DEPOSIT_VAR1) {
    return push_dim<pm::DIM_YZ, pm::DEPOSIT_VAR1>(c, true);
  }
  return push_dim<pm::DIM_YZ, pm::DEPOSIT_SPLIT>(c, true);
#endif
}

template <int DIM, int DEPOSIT>
static int deposit_paths_dim(Ctx* c, const psc_b200_jpath* d_paths, uint32_t n)
{
  KernelScope ks(c, "deposit_paths");
  PUSH_VARIANT::k_deposit_paths<DIM, DEPOSIT><<<div_up(n, 128), 128, 0, c->stream>>>(c->gd, n, d_paths, c->fld(0),
                                                                     c->fld_slot_len(0));
  c->n_launches++;
  return check_launch(c, "deposit_paths");
}

// d_paths: n trajectories on the device (capi.cu psc_b200_deposit_j stages them)
int PUSH_CAT(deposit_paths_, PUSH_VARIANT)(Ctx* c, const psc_b200_jpath* d_paths, uint32_t n)
{
  using namespace PUSH_VARIANT;
#ifdef PUSH_PROBE
  return deposit_paths_dim<pm::DIM_XYZ, pm::DEPOSIT_SPLIT>(c, d_paths, n);
#else
  if (c->gd.dim == pm::DIM_XYZ) {
    return deposit_paths_dim<pm::DIM_XYZ, pm::DEPOSIT_SPLIT>(c, d_paths, n);
  } else if (c->gd.deposit == pm::DEPOSIT_VAR1) {
    return deposit_paths_dim<pm::DIM_YZ, pm::DEPOSIT_VAR1>(c, d_paths, n);
  }
  return deposit_paths_dim<pm::DIM_YZ, pm::DEPOSIT_SPLIT>(c, d_paths, n);
#endif
}

int PUSH_CAT(push_mprts_, PUSH_VARIANT)(Ctx* c)
{
  using namespace PUSH_VARIANT;
  // push_particles_1vb.hxx:48: J = 0 on every patch
  PSC_TRY(flds_zero(c, 0, pm::JXI, pm::JXI + 3));
  int rc;
#ifdef PUSH_PROBE
  rc = push_dim<pm::DIM_XYZ, pm::DEPOSIT_SPLIT>(c);
#else
  if (c->gd.dim == pm::DIM_XYZ) {
    rc = push_dim<pm::DIM_XYZ, pm::DEPOSIT_SPLIT>(c);
  } else if (c->gd.deposit == pm::DEPOSIT_VAR1) {
    rc = push_dim<pm::DIM_YZ, pm::DEPOSIT_VAR1>(c);
  } else {
    rc = push_dim<pm::DIM_YZ, pm::DEPOSIT_SPLIT>(c);
  }
#endif
  // particles have moved: cell order and cell offsets no longer describe the store
  // (the fused boundary+sort pass of step() picks the store up from here)
  c->pushed_from_sorted = c->sorted && rc == 0;
  c->sorted = false;
  c->want_counts = false;
  return rc;
}

} // namespace psc_b200
