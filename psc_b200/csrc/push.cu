// psc_b200: PushParticlesB200::push_mprts -- 1st-order "ec" gather, Boris push, move and
// charge-conserving 1vb current deposit in one kernel.
//
// What: PushParticlesVb<C>::push_mprts (libpsc/psc_push_particles/push_particles_1vb.hxx:27-84)
// with Current1vbVar1 (yz) / Current1vbSplit (xyz, yz); the per-particle arithmetic is
// pic_math.cuh.  How (ours, not PSC's cuda_push_mprts_yz.cxx):
//
//  k_push_tiled   the store is ordered by (patch, cell).  One CTA owns a tile of cells
//                 of one patch; it stages the E/B tile (+1 node halo, +1 guard) into
//                 shared memory -- with cp.async.bulk (TMA bulk copies, one per
//                 contiguous x-row, completion on an mbarrier) or plain LDG/STS --,
//                 keeps a J tile of the same shape in shared memory, and walks the
//                 tile's particle runs warp by warp with 128-bit loads/stores.
//                 Deposits of the first trajectory segment are pre-reduced across the
//                 warp: lanes are grouped by target cell (sorted particles => 1-2
//                 groups per warp) and each group's 8/12 values are summed with a
//                 transposing shuffle butterfly (9/16 SHFL instead of 40/60), then a
//                 few lanes issue one shared-memory atomic each.  Extra segments of
//                 cell-crossing particles (a few %) use per-lane shared atomics.  The
//                 tile (halo included) is flushed with global red.add, so no
//                 checkerboard passes are needed.
//  k_push_general any particle order: one thread per particle, fields through the
//                 read-only path, J with global red.add.  Used when the store is not
//                 sorted (same results, same particle order).
//
// This file is compiled twice: -fmad=false (namespace exact: particle update is
// bit-identical to PSC's x86-64 build, which has no FMA) and with FMA contraction
// (namespace fast, within a few ULP).
#include "dev_util.cuh"

#include <algorithm>

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

struct TileGeom
{
  int t[3];    // cells per tile edge
  int nt[3];   // tiles per patch edge
  int f[3];    // shared tile extent (nodes): t + 3 in non-invariant dims, 1 otherwise
  int g[3];    // 1 in non-invariant dims (origin shift), 0 otherwise
  int fx_pad;  // row pitch of the shared tile
  int n_tile_nodes; // f[2]*f[1]*fx_pad
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

struct FldTile
{
  const float* s; // EM tile, component-major
  int o0, o1, o2; // global index of tile node 0
  int sy, sz, sm;
  __device__ __forceinline__ float operator()(int m, int i, int j, int k) const
  {
    return s[(m - pm::EX) * sm + (k - o2) * sz + (j - o1) * sy + (i - o0)];
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
__device__ __forceinline__ int leaf_lin(int n, int sy, int sz, int sm)
{
  int m, ox, oy, oz;
  pm::leaf_slot<DIM>(n, m, ox, oy, oz);
  return m * sm + oz * sz + oy * sy + ox;
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
  long base = fld_off(G, 0, ci[0], ci[1], ci[2]);
  int sy = G.im[0], sz = G.im[0] * G.im[1];
  long sm = G.fld_len;
#pragma unroll
  for (int n = 0; n < NV; n++) {
    int m, ox, oy, oz;
    pm::leaf_slot<DIM>(n, m, ox, oy, oz);
    atomicAdd(F + base + m * sm + oz * sz + oy * sy + ox, val[n]);
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

// ---------------------------------------------------------------- tiled kernel

// mbarrier / bulk-copy PTX (sm_90+): one thread arms the barrier with the byte count,
// issues one cp.async.bulk per contiguous row, everybody waits on the phase.
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

template <int DIM, int DEPOSIT, bool WARP_REDUCE, bool TMA>
__global__ void __launch_bounds__(512)
  k_push_tiled(GridDev G, TileGeom T, const uint32_t* __restrict__ cell_off,
               float4* __restrict__ xi4, float4* __restrict__ pxi4, float* __restrict__ flds,
               long slot_len)
{
  constexpr int NV = pm::LeafShape<DIM>::NV;
  constexpr int NVP = (DIM == pm::DIM_XYZ) ? 16 : 8;
  extern __shared__ __align__(128) float smem[];
  __shared__ uint64_t bar;
  float* sEM = smem;                      // [6][f2][f1][fx_pad]
  float* sJ = smem + 6 * T.n_tile_nodes;  // [3][f2][f1][fx_pad]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, n_warps = blockDim.x >> 5;
  const int tiles_per_patch = T.nt[0] * T.nt[1] * T.nt[2];
  const int p = blockIdx.x / tiles_per_patch;
  int tt = blockIdx.x - p * tiles_per_patch;
  int o[3], e[3];
  o[0] = (tt % T.nt[0]) * T.t[0];
  o[1] = ((tt / T.nt[0]) % T.nt[1]) * T.t[1];
  o[2] = (tt / (T.nt[0] * T.nt[1])) * T.t[2];
#pragma unroll
  for (int d = 0; d < 3; d++) {
    e[d] = min(T.t[d], G.ldims[d] - o[d]);
  }
  float* F = flds + p * slot_len;
  const int sy = T.fx_pad, sz = T.fx_pad * T.f[1], sm = T.n_tile_nodes;
  // global index of tile node 0
  const int n0 = o[0] - T.g[0], n1 = o[1] - T.g[1], n2 = o[2] - T.g[2];

  // ---- stage E/B, zero J
  if (TMA) {
    // rows are contiguous in x: one bulk copy per (comp, z, y) row.  Source address and
    // size must be 16-byte multiples: the host only selects this path when they are.
    const int rows = 6 * T.f[2] * T.f[1];
    const unsigned row_bytes = (unsigned)T.fx_pad * 4u;
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
        int m = r / (T.f[2] * T.f[1]);
        int rem = r - m * (T.f[2] * T.f[1]);
        int kz = rem / T.f[1], ky = rem - kz * T.f[1];
        const float* src = F + fld_off(G, pm::EX + m, n0, n1 + ky, n2 + kz);
        bulk_g2s(sEM + m * sm + kz * sz + ky * sy, src, row_bytes, &bar);
      }
    }
    for (int idx = tid; idx < 3 * T.n_tile_nodes; idx += blockDim.x) {
      sJ[idx] = 0.f;
    }
    mbar_wait(&bar, 0);
    __syncthreads();
  } else {
    for (int idx = tid; idx < 6 * T.n_tile_nodes; idx += blockDim.x) {
      int m = idx / T.n_tile_nodes;
      int rem = idx - m * T.n_tile_nodes;
      int kz = rem / sz;
      rem -= kz * sz;
      int ky = rem / sy, kx = rem - ky * sy;
      int gi = n0 + kx, gj = n1 + ky, gk = n2 + kz;
      float v = 0.f;
      if (kx < T.f[0] && gi < G.ldims[0] + G.ibn[0] && gj < G.ldims[1] + G.ibn[1] &&
          gk < G.ldims[2] + G.ibn[2]) {
        v = __ldg(F + fld_off(G, pm::EX + m, gi, gj, gk));
      }
      sEM[idx] = v;
    }
    for (int idx = tid; idx < 3 * T.n_tile_nodes; idx += blockDim.x) {
      sJ[idx] = 0.f;
    }
    __syncthreads();
  }

  FldTile EM{sEM, n0, n1, n2, sy, sz, sm};
  const int my_slot = slot_of_lane<NVP>(lane);
  const int my_lin = (my_slot < NV) ? leaf_lin<DIM>(my_slot, sy, sz, sm) : 0;
  const bool writer = (my_slot < NV) && ((lane & (NVP == 16 ? 1 : 3)) == 0);

  // ---- particle runs: contiguous cells along the first non-invariant dim
  const int n_rows = (DIM == pm::DIM_XYZ) ? e[1] * e[2] : e[2];
  const int run_cells = (DIM == pm::DIM_XYZ) ? e[0] : e[1];
  const uint32_t* coff = cell_off + (size_t)p * G.n_cells;
  for (int row = warp; row < n_rows; row += n_warps) {
    int c0;
    if (DIM == pm::DIM_XYZ) {
      int ry = row % e[1], rz = row / e[1];
      c0 = ((o[2] + rz) * G.ldims[1] + (o[1] + ry)) * G.ldims[0] + o[0];
    } else {
      c0 = (o[2] + row) * G.ldims[1] + o[1];
    }
    const uint32_t begin = __ldg(&coff[c0]), end = __ldg(&coff[c0 + run_cells]);
    for (uint32_t base = begin; base < end; base += 32) {
      const uint32_t i = base + lane;
      const bool act = i < end;
      Walker<DIM, DEPOSIT> w;
      float val[NVP];
      int ci[3] = {0, 0, 0};
      bool more = false;
      float qw = 0.f;
      if (act) {
        float4 X = xi4[i], U = pxi4[i];
        float x[3] = {X.x, X.y, X.z}, u[3] = {U.x, U.y, U.z};
        qw = U.w;
        pm::Trajectory t;
        pm::advance<DIM>(G.pc, EM, x, u, __float_as_int(X.w), t);
        xi4[i] = make_float4(x[0], x[1], x[2], X.w);
        pxi4[i] = make_float4(u[0], u[1], u[2], U.w);
        more = w.first(G.pc, t, qw, ci, val);
      }
      // tile-relative leaf cell; in_tile: all NV targets lie inside the shared tile
      int r0 = ci[0] - n0, r1 = ci[1] - n1, r2 = ci[2] - n2;
      bool in_tile = (DIM == pm::DIM_YZ || (unsigned)r0 < (unsigned)(T.f[0] - 1)) &&
                     (unsigned)r1 < (unsigned)(T.f[1] - 1) && (unsigned)r2 < (unsigned)(T.f[2] - 1);
      int key = r2 * sz + r1 * sy + r0;
      if (act && !in_tile) {
        leaf_to_global<DIM>(G, F, ci, val);
      }
      if (WARP_REDUCE) {
        unsigned rem = __ballot_sync(FULL, act && in_tile);
        int iter = 0;
        while (rem) {
          if (iter == 2 || __popc(rem) < 4) {
            if ((rem >> lane) & 1) {
#pragma unroll
              for (int n = 0; n < NV; n++) {
                atomicAdd(&sJ[key + leaf_lin<DIM>(n, sy, sz, sm)], val[n]);
              }
            }
            break;
          }
          int leader = __ffs(rem) - 1;
          int kl = __shfl_sync(FULL, key, leader);
          bool mine = ((rem >> lane) & 1) && key == kl;
          unsigned grp = __ballot_sync(FULL, mine);
          float v[NVP];
#pragma unroll
          for (int n = 0; n < NVP; n++) {
            v[n] = (mine && n < NV) ? val[n] : 0.f;
          }
          warp_transpose_reduce<NVP>(v, lane);
          if (writer) {
            atomicAdd(&sJ[kl + my_lin], v[0]);
          }
          rem &= ~grp;
          iter++;
        }
      } else {
        if (act && in_tile) {
#pragma unroll
          for (int n = 0; n < NV; n++) {
            atomicAdd(&sJ[key + leaf_lin<DIM>(n, sy, sz, sm)], val[n]);
          }
        }
      }
      // further segments of cell-crossing particles
      while (__any_sync(FULL, more)) {
        if (more) {
          more = w.next(G.pc, qw, ci, val);
          r0 = ci[0] - n0;
          r1 = ci[1] - n1;
          r2 = ci[2] - n2;
          in_tile = (DIM == pm::DIM_YZ || (unsigned)r0 < (unsigned)(T.f[0] - 1)) &&
                    (unsigned)r1 < (unsigned)(T.f[1] - 1) && (unsigned)r2 < (unsigned)(T.f[2] - 1);
          if (in_tile) {
            key = r2 * sz + r1 * sy + r0;
#pragma unroll
            for (int n = 0; n < NV; n++) {
              atomicAdd(&sJ[key + leaf_lin<DIM>(n, sy, sz, sm)], val[n]);
            }
          } else {
            leaf_to_global<DIM>(G, F, ci, val);
          }
        }
      }
    }
  }
  __syncthreads();

  // ---- flush the J tile (halo included) with global reductions
  for (int idx = tid; idx < 3 * T.n_tile_nodes; idx += blockDim.x) {
    float v = sJ[idx];
    if (v != 0.f) {
      int m = idx / T.n_tile_nodes;
      int rem = idx - m * T.n_tile_nodes;
      int kz = rem / sz;
      rem -= kz * sz;
      int ky = rem / sy, kx = rem - ky * sy;
      int gi = n0 + kx, gj = n1 + ky, gk = n2 + kz;
      if (gi < G.ldims[0] + G.ibn[0] && gj < G.ldims[1] + G.ibn[1] && gk < G.ldims[2] + G.ibn[2]) {
        atomicAdd(F + fld_off(G, m, gi, gj, gk), v);
      }
    }
  }
}

// ---------------------------------------------------------------- host side

template <int DIM, int DEPOSIT>
static int launch_tiled(Ctx* c, const TileGeom& T, bool tma, size_t smem_bytes)
{
  const GridDev& G = c->gd;
  int tiles = T.nt[0] * T.nt[1] * T.nt[2] * G.n_patches;
  float* F = c->fld(0);
  long slot_len = c->fld_slot_len(0);
  int threads = std::max(32, std::min(512, c->opt_threads)) & ~31;
#define PSC_LAUNCH(WR, TM)                                                                        \
  do {                                                                                            \
    auto kern = k_push_tiled<DIM, DEPOSIT, WR, TM>;                                               \
    PSC_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,          \
                                      (int)smem_bytes));                                          \
    kern<<<tiles, threads, smem_bytes, c->stream>>>(G, T, c->d_cell_off, c->xi(), c->pxi(), F,    \
                                                    slot_len);                                    \
  } while (0)
  if (c->opt_warp_reduce) {
    if (tma) {
      PSC_LAUNCH(true, true);
    } else {
      PSC_LAUNCH(true, false);
    }
  } else {
    if (tma) {
      PSC_LAUNCH(false, true);
    } else {
      PSC_LAUNCH(false, false);
    }
  }
#undef PSC_LAUNCH
  return 0;
}

template <int DIM, int DEPOSIT>
static int push_dim(Ctx* c)
{
  const GridDev& G = c->gd;
  if (c->n_prts == 0) {
    return 0;
  }
  if (c->sorted && c->opt_tiled) {
    TileGeom T{};
    int def[3] = {DIM == pm::DIM_XYZ ? 8 : 1, DIM == pm::DIM_XYZ ? 8 : 16, DIM == pm::DIM_XYZ ? 8 : 16};
    for (int d = 0; d < 3; d++) {
      bool inv = (DIM == pm::DIM_YZ && d == 0);
      int t = c->opt_tile[d] > 0 ? c->opt_tile[d] : def[d];
      T.t[d] = inv ? 1 : std::min(t, G.ldims[d]);
      T.nt[d] = (G.ldims[d] + T.t[d] - 1) / T.t[d];
      T.f[d] = inv ? 1 : T.t[d] + 3;
      T.g[d] = inv ? 0 : 1;
    }
    // TMA rows: the contiguous dim is x (xyz) -- pad the pitch to 4 floats; the row
    // start o-1+ibn = o+1 has to be a multiple of 4 floats as well, which no tile
    // origin satisfies in general => the bulk-copy path is taken only when every
    // row start is 16-byte aligned (checked here), else LDG/STS staging.
    T.fx_pad = T.f[0];
    bool tma = false;
    if (c->opt_tma && DIM == pm::DIM_XYZ && G.ibn[0] == 2) {
      // bulk copies need 16-byte aligned row starts and sizes: start the shared tile
      // two nodes left of the tile origin (row start = o + ibn - 2 = o floats into the
      // row) and round the pitch up to 4 floats
      int pitch = (T.t[0] + 4 + 3) & ~3;
      bool ok = (G.im[0] % 4 == 0) && (T.t[0] % 4 == 0) && (c->fld_slot_len(0) % 4 == 0);
      // all rows of the padded tile must stay inside the patch array
      ok = ok && ((T.nt[0] - 1) * T.t[0] + pitch <= G.im[0]) &&
           ((T.nt[1] - 1) * T.t[1] - 1 + G.ibn[1] + T.f[1] <= G.im[1]) &&
           ((T.nt[2] - 1) * T.t[2] - 1 + G.ibn[2] + T.f[2] <= G.im[2]);
      if (ok) {
        tma = true;
        T.g[0] = 2;
        T.f[0] = pitch;
        T.fx_pad = pitch;
      }
    }
    T.n_tile_nodes = T.f[2] * T.f[1] * T.fx_pad;
    size_t smem_bytes = (size_t)9 * T.n_tile_nodes * sizeof(float);
    if (smem_bytes <= 200 * 1024) {
      KernelScope ks(c, tma ? "push_tiled_tma" : "push_tiled");
      PSC_TRY((launch_tiled<DIM, DEPOSIT>(c, T, tma, smem_bytes)));
      c->n_launches++;
      return check_launch(c, "push_tiled");
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

} // namespace PUSH_VARIANT

int PUSH_CAT(push_mprts_, PUSH_VARIANT)(Ctx* c)
{
  using namespace PUSH_VARIANT;
  // push_particles_1vb.hxx:48: J = 0 on every patch
  PSC_TRY(flds_zero(c, 0, pm::JXI, pm::JXI + 3));
  int rc;
  if (c->gd.dim == pm::DIM_XYZ) {
    rc = push_dim<pm::DIM_XYZ, pm::DEPOSIT_SPLIT>(c);
  } else if (c->gd.deposit == pm::DEPOSIT_VAR1) {
    rc = push_dim<pm::DIM_YZ, pm::DEPOSIT_VAR1>(c);
  } else {
    rc = push_dim<pm::DIM_YZ, pm::DEPOSIT_SPLIT>(c);
  }
  // particles have moved: cell order and cell offsets no longer describe the store
  // (the fused boundary+sort pass of step() picks the store up from here)
  c->pushed_from_sorted = c->sorted && rc == 0;
  c->sorted = false;
  return rc;
}

} // namespace psc_b200
