// psc_b200: k_push_lean -- the tiled push + deposit kernel for a cell-ordered store and a
// compile-time tile geometry (included by push.cu inside namespace PUSH_VARIANT).
//
// Same tiling as k_push_tiled (one CTA = one tile of cells of one patch, E/B tile + halo in
// shared memory, J tile in shared memory flushed with global red.add, warps walk rows of
// cells in 32-particle chunks, cell-crossing trajectories parked in a per-warp queue and
// walked 32 at a time).  What differs, all of it to cut issue slots per particle
// (profiles/r01_v10_push_tiled_ncu.txt: 609 warp instructions per chunk, 40 % of them FP):
//
//  * the E/B tile arrives with ONE tensor-map TMA (cp.async.bulk.tensor over
//    (x, y, z, component x patch)); out-of-range rows are zero-filled by the unit, so there
//    is no alignment fallback inside the kernel
//  * a particle that stays in its cell deposits through per-cell MOMENTS instead of leaf
//    values: with m_d = q dx_d and (a, b) the centred offsets in the two other directions,
//    the four values of component d are linear in  S = sum m_d, Sa = sum m_d a,
//    Sb = sum m_d b, Sab = sum (m_d a b + q h):
//        v(0,0) = S - Sa - Sb + Sab   v(1,0) = Sa - Sab   v(0,1) = Sb - Sab   v(1,1) = Sab
//    (calc_j2_one_cell + CurrentDeposition1vb, inc_curr_1vb_split.cxx:25-33,
//    psc/current_deposition.hxx:17-40; curr_3d_vb_cell, inc_curr_1vb_var1.cxx:62-89).
//    Per particle that is 6 FMA-class operations per component instead of 18; the
//    combination runs once per cell after the warp reduction.  J is compared at 1e-5 of
//    max|J| (summation order is free: the reference sums sequentially, we sum by warp), the
//    particle update itself keeps the reference's operation order bit for bit.
//  * destination classes (the input of the fused boundary exchange + sort) are counted only
//    for the particles that left their cell: a particle whose trajectory is one segment is
//    in class CENTER by construction.  The count planes are cleared before the launch; the
//    queue walk adds one per leaver (32-bit red on the 16-bit planes), the row walk adds
//    (population - leavers) to the CENTER plane once per cell.
//  * the pass loop over the cells a chunk touches carries warp-uniform lane masks only.
#pragma once

namespace lean
{

// warps per CTA: 8 x 3 CTAs per SM at 80 registers (W = 1), 8 x 2 CTAs at 128 (W = 2; with 10 warps at 96
// registers the xyz kernel spills and takes 20.8 ms instead of 18.9, with 12 at 80 it spills 264 bytes)
template <int W>
__host__ __device__ constexpr int n_warps()
{
#ifdef LEAN_NW2
  return W == 1 ? 8 : LEAN_NW2;
#else
  return W == 1 ? 8 : 8;
#endif
}
// chunks in flight per warp (W = 1): the records travel global -> shared with cp.async that many
// chunks ahead of the one being computed.  One ahead left every chunk waiting for its records
// (the cp.async wait was the largest stall site of the kernel: a chunk takes a warp ~3600 cycles,
// about the loaded memory latency); the yz kernel has fewer instructions per chunk and more
// shared memory to spare, so it runs further ahead.  S3D push 17.1 -> 16.4 ms, yz 24.6 -> 20.6 ms
template <int DIM>
__host__ __device__ constexpr int n_stages()
{
  return DIM == pm::DIM_XYZ ? 2 : 4;
}
// bytes of staging per warp and stage: 32 x (xi4, pxi4)
constexpr uint32_t STAGE_BYTES = 1024;

// queue entries per warp: a chunk (32 W particles) must always fit behind what a walk leaves
template <int W>
__host__ __device__ constexpr int qcap()
{
  return W == 1 ? 56 : 96;
}

// E/B tile accessor for two particles per lane
template <typename GEO>
struct FldTile2
{
  const float* s; // EM tile, component-major; s points at node (n0, n1, n2)
  const GEO& geo;
  int n0, n1, n2;
  __device__ __forceinline__ pm::f2 operator()(int m, pm::i2 i, pm::i2 j, pm::i2 k) const
  {
    return pm::mk2(s[(m - pm::EX) * geo.sm() + (k.x - n2) * geo.sz() + (j.x - n1) * geo.sy() + (i.x - n0)],
                   s[(m - pm::EX) * geo.sm() + (k.y - n2) * geo.sz() + (j.y - n1) * geo.sy() + (i.y - n0)]);
  }
};

// float add to shared memory through a 32-bit shared-space address: same CAS loop as
// atomicAdd(float*), but no generic -> shared conversion (an S2UR on the critical path of
// every cell flush: 9 % of the stall samples, profiles/README.md)
__device__ __forceinline__ void red_shared_f32(uint32_t addr, float v)
{
  asm volatile("red.shared.add.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

__device__ __forceinline__ void tma_load_tile(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3,
                                              uint64_t* bar, bool four_d)
{
  if (four_d) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
                 " [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(smem_u32(dst)),
                 "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
                 : "memory");
  } else {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
                 " [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(dst)),
                 "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
                 : "memory");
  }
}

// after warp_transpose_reduce every lane holds one moment of the cell; combine the moments
// of a component into the leaf value that lane deposits (header comment)
template <int DIM>
__device__ __forceinline__ float moments_to_leaf(float v, int lane)
{
  if (DIM == pm::DIM_XYZ) {
    // slot = lane >> 1 = 4 * component + k, k = 0: S, 1: Sa, 2: Sb, 3: Sab
    const int k = (lane >> 1) & 3;
    const int b = lane & ~6;
    const float sab = __shfl_sync(FULL, v, lane | 6);
    const float sa = __shfl_sync(FULL, v, b | 2);
    const float sb = __shfl_sync(FULL, v, b | 4);
    float r = v - (k == 0 ? sa : (k == 3 ? 0.f : sab));
    if (k == 0) {
      r = (r - sb) + sab;
    }
    return r;
  } else {
    // slot = lane >> 2: jx S, Sa, Sb, Sab | jy S, Sb | jz S, Sa
    const int s = lane >> 2;
    const float m1 = __shfl_sync(FULL, v, 4), m2 = __shfl_sync(FULL, v, 8), m3 = __shfl_sync(FULL, v, 12);
    const float m5 = __shfl_sync(FULL, v, 20), m7 = __shfl_sync(FULL, v, 28);
    float r = v;
    if (s == 0) {
      r = ((v - m1) - m2) + m3;
    } else if (s == 1 || s == 2) {
      r = v - m3;
    } else if (s == 4) {
      r = v - m5;
    } else if (s == 6) {
      r = v - m7;
    }
    return r;
  }
}

// W = particles per lane: 1, or 2 with the update on the packed FP32 pipe (pic_math.cuh f2)
template <int DIM, int DEPOSIT, bool COUNT, bool SAME, int W>
__global__ void __launch_bounds__(n_warps<W>() * 32, W == 1 ? 3 : 2)
  k_push_lean(const __grid_constant__ CUtensorMap tm, GridDev G, GeoStatic<DIM> geo, PushArgs A)
{
  constexpr int QC = qcap<W>();
  constexpr int NW = n_warps<W>();
  constexpr bool XYZ = DIM == pm::DIM_XYZ;
  constexpr int NM = XYZ ? 12 : 8;   // moments per cell = leaf values per cell
  constexpr int NVP = XYZ ? 16 : 8;  // padded to the butterfly width
  constexpr int NODES = GeoStatic<DIM>::sm();
  constexpr int SY = GeoStatic<DIM>::sy(), SZ = GeoStatic<DIM>::sz();
  constexpr int RD = XYZ ? 0 : 1;    // direction a row of cells runs along
  constexpr int ROW_STRIDE = XYZ ? 1 : SY;
  extern __shared__ __align__(128) float smem[];
  float* sEM = smem;             // [6][f2][f1][f0]
  float* sJ = smem + 6 * NODES;  // [3][f2][f1][f0]
  const uint32_t sJ32 = smem_u32(sJ);
  float4* sQ = reinterpret_cast<float4*>(smem + ((9 * NODES + 3) & ~3)); // [NW][QC][2]
  float4* sP = sQ + NW * QC * 2;                                         // [NW][2][32 W] next chunk
  __shared__ uint64_t bar;
  __shared__ int row_ctr; // rows are handed out dynamically (balances the warps)

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned lt = (1u << lane) - 1u;
  const int tiles_per_patch = geo.nt(0) * geo.nt(1) * geo.nt(2);
  const int p = blockIdx.x / tiles_per_patch;
  const int tt = blockIdx.x - p * tiles_per_patch;
  const int o0 = (tt % geo.nt(0)) * geo.t(0);
  const int o1 = ((tt / geo.nt(0)) % geo.nt(1)) * geo.t(1);
  const int o2 = (tt / (geo.nt(0) * geo.nt(1))) * geo.t(2);
  float* F = A.flds + p * A.slot_len;
  // global index of tile node 0
  const int n0 = o0 - geo.g(0), n1 = o1 - geo.g(1), n2 = o2 - geo.g(2);

  // ---- stage E/B (one TMA), zero J
  if (tid == 0) {
    row_ctr = NW;
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(&bar, 6u * NODES * sizeof(float));
    if (XYZ) {
      tma_load_tile(sEM, &tm, n0 + 2, n1 + 2, n2 + 2, p * 9 + pm::EX, &bar, true);
    } else {
      tma_load_tile(sEM, &tm, n1 + 2, n2 + 2, p * 9 + pm::EX, 0, &bar, false);
    }
  }
  for (int idx = tid; idx < 3 * NODES; idx += NW * 32) {
    sJ[idx] = 0.f;
  }
  if (warp == 0) {
    mbar_wait(&bar, 0);
  }
  __syncthreads();

  FldTile<GeoStatic<DIM>> EM{sEM, geo, n0, n1, n2};
  float4* const myQ = sQ + warp * QC * 2;
  // W = 1: NST stages of 1 KB per warp, the block aligned to its size so that the stage is a
  // bit field of the address (W = 2 keeps one stage of two chunks)
  constexpr int NST = n_stages<DIM>();
  constexpr uint32_t RING = NST * STAGE_BYTES;
  uint32_t myP = W == 1 ? ((smem_u32(sP) + (RING - 1)) & ~(RING - 1)) + (uint32_t)warp * RING + (uint32_t)lane * 16u
                        : smem_u32(sP + warp * 64 * W + lane);
  constexpr uint32_t PSTRIDE = 32 * sizeof(float4); // xi4 -> pxi4 inside a stage
  int qn = 0; // queued trajectories of this warp (warp-uniform)

  // what this lane deposits when a cell is flushed: its slot of the leaf, scaled
  const int my_slot = slot_of_lane<NVP>(lane);
  const bool writer = (my_slot < NM) && ((lane & (NVP == 16 ? 1 : 3)) == 0);
  const int my_comp = XYZ ? (my_slot >> 2) : (my_slot < 4 ? 0 : (my_slot < 6 ? 1 : 2));
  const float my_fnq = DEPOSIT == pm::DEPOSIT_SPLIT ? G.pc.fnqs_split[my_comp % 3] : G.pc.fnq_var1[my_comp % 3];
  const int myJ = my_slot < NM ? leaf_lin<DIM>(my_slot, SY, SZ, NODES) : 0;
  uint32_t* const cnt32 = reinterpret_cast<uint32_t*>(A.cnt);
  const size_t cen0 = (size_t)CLS_CENTER * A.nct + (size_t)p * G.n_cells; // CENTER plane, this patch

  // split + deposit `cnt` queued trajectories (entries [qn - cnt, qn)), one per lane; COUNT:
  // add each one to the plane of its destination class
  auto drain = [&](int cnt) {
    const bool a2 = lane < cnt;
    Walker<DIM, DEPOSIT> w;
    float val[NM];
    int ci[3] = {0, 0, 0};
    bool more = false;
    float qw = 0.f;
    if (a2) {
      const float4 A0 = myQ[2 * (qn - cnt + lane)], A1 = myQ[2 * (qn - cnt + lane) + 1];
      pm::Trajectory t;
      int sc[3], dc[3]; // the indexer's source and destination cells
      float xn[3];      // pushed position (SAME: read back when needed)
      if constexpr (SAME) {
        t.xm[0] = XYZ ? A0.x : 0.f, t.xm[1] = A0.y, t.xm[2] = A0.z;
        t.xp[0] = XYZ ? A1.x : 0.f, t.xp[1] = A1.y, t.xp[2] = A1.z;
      } else {
        xn[0] = A1.x, xn[1] = A1.y, xn[2] = A1.z;
        const float xo[3] = {A0.x, A0.y, A0.z};
#pragma unroll
        for (int d = 0; d < 3; d++) {
          t.xm[d] = xo[d] * G.pc.dxi[d];
          t.xp[d] = xn[d] * G.pc.dxi[d];
          sc[d] = pm::cell_position(G.pc, xo[d], d);
          dc[d] = pm::cell_position(G.pc, xn[d], d);
        }
      }
      t.v[0] = XYZ ? 0.f : A1.w, t.v[1] = 0.f, t.v[2] = 0.f;
      qw = A0.w;
#pragma unroll
      for (int d = 0; d < 3; d++) {
        t.lg[d] = pm::fint(t.xm[d]);
        t.lf[d] = pm::fint(t.xp[d]);
        if constexpr (SAME) {
          sc[d] = t.lg[d], dc[d] = t.lf[d];
        }
      }
      if constexpr (COUNT) {
        const int d0 = dc[0] - sc[0], d1 = dc[1] - sc[1], d2 = dc[2] - sc[2];
        const bool ok = (unsigned)dc[0] < (unsigned)G.ldims[0] && (unsigned)dc[1] < (unsigned)G.ldims[1] &&
                        (unsigned)dc[2] < (unsigned)G.ldims[2] && (unsigned)(d0 + 1) <= 2u &&
                        (unsigned)(d1 + 1) <= 2u && (unsigned)(d2 + 1) <= 2u;
        int cls, rq = 0, rc = 0; // (rq, rc): fs_classify's target rank / direction of a remote leaver
        if (ok) {
          cls = ((d2 + 1) * 3 + d1 + 1) * 3 + d0 + 1;
        } else {
          // patch boundary (or further than one cell): the pushed record decides
          if constexpr (SAME) {
            const uint32_t i = (uint32_t)__float_as_int(XYZ ? A1.w : A0.x);
            const float4 Xr = A.xi4[i];
            xn[0] = Xr.x, xn[1] = Xr.y, xn[2] = Xr.z;
          }
          float uu[3] = {0.f, 0.f, 0.f};
          cls = fs_classify(G, A.tab, p, sc[0], sc[1], sc[2], xn, uu, rq, rc);
        }
        if (cls < FS_PLANES) {
          const size_t e = (size_t)cls * A.nct + (size_t)p * G.n_cells +
                           (size_t)((sc[2] * G.ldims[1] + sc[1]) * G.ldims[0] + sc[0]);
          atomicAdd(cnt32 + (e >> 1), 1u << (16 * (e & 1)));
        } else if (cls == CLS_BAD) {
          atomicExch(&A.flags[0], 1u);
        } else if (cls == CLS_DROP) {
          atomicAdd(&A.flags[1], 1u);
        } else if (cls == CLS_REMOTE) {
          const uint32_t slot = atomicAdd(&A.flags[2], 1u);
          if constexpr (SAME) {
            if (slot < A.rem_cap) {
              A.rem_key[slot] = ((uint32_t)(-2 - rq) * G.n_patches + p) * 32u + (uint32_t)rc;
              A.rem_idx[slot] = (uint32_t)__float_as_int(XYZ ? A1.w : A0.x);
            }
          }
        }
      }
      // The leaves of a cell-crossing trajectory go to the global J directly (red.global.add,
      // fire and forget).  Into the shared tile they are float compare-and-swap loops
      // (ATOMS.CAST.SPIN), twelve per leaf one after the other, and the 32 trajectories of a walk
      // come from neighbouring cells, so the loops also retry: those loops were the largest
      // single source of stall samples in the kernel (profiles/r02_push_lean_ncu.txt) for 3 % of
      // the particles.  The J tile of the CTA is in L2 while the CTA runs; S3D push 17.58 -> 17.08 ms
      more = w.first(G.pc, t, qw, ci, val);
#ifdef LEAN_DRAIN_SHARED // (A/B: tools/build_variant.sh)
      leaf_deposit<DIM>(G, geo, sJ, F, n0, n1, n2, ci, val);
#else
      leaf_to_global<DIM>(G, F, ci, val);
#endif
    }
    while (__any_sync(FULL, more)) {
      if (more) {
        more = w.next(G.pc, qw, ci, val);
#ifdef LEAN_DRAIN_SHARED
        leaf_deposit<DIM>(G, geo, sJ, F, n0, n1, n2, ci, val);
#else
        leaf_to_global<DIM>(G, F, ci, val);
#endif
      }
    }
    qn -= cnt;
    __syncwarp();
  };

  // ---- particle runs: rows of cells along the first non-invariant dim
  constexpr int N_ROWS = XYZ ? GeoStatic<DIM>::t(1) * GeoStatic<DIM>::t(2) : GeoStatic<DIM>::t(2);
  constexpr int RUN = GeoStatic<DIM>::t(RD); // cells per row, <= 31
  const uint32_t* const coff = A.cell_off + (size_t)p * G.n_cells;
  for (int row = warp; row < N_ROWS;) {
    int c0, rs1, rs2; // first cell of the row; row coordinates
    if (XYZ) {
      const int ry = row % geo.t(1), rz = row / geo.t(1);
      rs1 = o1 + ry, rs2 = o2 + rz;
      c0 = (rs2 * G.ldims[1] + rs1) * G.ldims[0] + o0;
    } else {
      rs1 = o1, rs2 = o2 + row;
      c0 = rs2 * G.ldims[1] + o1;
    }
    // lane j holds the offset of the row's j-th cell boundary
    const uint32_t myoff = __ldg(&coff[c0 + min(lane, RUN)]);
    const uint32_t begin = __shfl_sync(FULL, myoff, 0), end = __shfl_sync(FULL, myoff, RUN);
    // shared J of the row's first cell, as seen by this lane's leaf slot
    const int jrow = myJ + (rs2 - n2) * SZ + (XYZ ? (rs1 - n1) * SY + (o0 - n0) : (o1 - n1) * SY);
    if constexpr (W == 1) {
    if (begin < end) {
      int cur = 0;                                           // cell of the row the passes are at
      uint32_t cb = begin, ce = __shfl_sync(FULL, myoff, 1); // its particle range
      uint32_t n_left = 0;                                   // ... and whether this lane's particles left it (summed at the flush)
      float acc[NM];                                         // this lane's share of the cell's moments
#pragma unroll
      for (int n = 0; n < NM; n++) {
        acc[n] = 0.f;
      }
      // warp-sum the moments, turn them into leaf values, add those to the shared J at `at`
      auto flush_moments = [&](int at) {
        float v[NVP];
#pragma unroll
        for (int n = 0; n < NVP; n++) {
          v[n] = n < NM ? acc[n] : 0.f;
        }
        warp_transpose_reduce<NVP>(v, lane);
        const float leaf = moments_to_leaf<DIM>(v[0], lane) * my_fnq;
        if (writer) {
          red_shared_f32(sJ32 + 4u * (uint32_t)at, leaf);
        }
#pragma unroll
        for (int n = 0; n < NM; n++) {
          acc[n] = 0.f;
        }
      };
      // the next chunk travels global -> shared with cp.async while this one is computed
#pragma unroll
      for (int k = 0; k < NST; k++) {
        // chunk k of the row into stage (current + k) of the ring, one commit group per chunk
        const uint32_t dst = (myP & ~(RING - 1)) | ((myP + k * STAGE_BYTES) & (RING - 1));
        if (begin + 32 * k + lane < end) {
          cp_async16(dst, A.xi4 + begin + 32 * k + lane);
          cp_async16(dst + PSTRIDE, A.pxi4 + begin + 32 * k + lane);
        }
        cp_async_commit();
      }
      uint32_t base = begin;
      do {
        const uint32_t i = base + lane;
        const bool act = i < end;
        if (qn > QC - 32) {
          // the queue may not take another chunk: walk it now.  The cell's moments so far are
          // flushed first (they are additive), so that nothing but the row state is live
          // across the walk
          flush_moments(jrow + cur * ROW_STRIDE);
          drain(min(qn, 32));
        }
        // all but the NST - 1 youngest groups have landed: this chunk's stage is complete
        asm volatile("cp.async.wait_group %0;" ::"n"(NST - 1) : "memory");
        const float4 X = lds128(myP), U = lds128(myP + PSTRIDE);
        if (i + 32 * NST < end) {
          cp_async16(myP, A.xi4 + i + 32 * NST);
          cp_async16(myP + PSTRIDE, A.pxi4 + i + 32 * NST);
        }
        cp_async_commit();
        myP = (myP & ~(RING - 1)) | ((myP + STAGE_BYTES) & (RING - 1));
        // ---- gather, Boris, move (the reference's arithmetic, pic_math.cuh)
        bool cross = false;
        float dx[3] = {0.f, 0.f, 0.f}, xa[3] = {0.f, 0.f, 0.f}; // displacement, centred offset
        float q = 0.f;                                         // q w of a particle that stayed in its cell
        pm::Trajectory t;
        float x[3] = {X.x, X.y, X.z};
        if (act) {
          float u[3] = {U.x, U.y, U.z};
          pm::advance<DIM>(G.pc, EM, x, u, __float_as_int(X.w), t);
          A.xi4[i] = make_float4(x[0], x[1], x[2], X.w);
          A.pxi4[i] = make_float4(u[0], u[1], u[2], U.w);
          cross = (XYZ && t.lf[0] != t.lg[0]) || t.lf[1] != t.lg[1] || t.lf[2] != t.lg[2];
          if constexpr (!SAME) {
            // 1/float(dx) and float(dx_inv) may disagree at a cell edge: such a particle is
            // walked like a crossing one (its leaf is not the run's cell)
            const float xo[3] = {X.x, X.y, X.z};
#pragma unroll
            for (int d = XYZ ? 0 : 1; d < 3; d++) {
              cross = cross || pm::cell_position(G.pc, xo[d], d) != t.lg[d];
            }
          }
#pragma unroll
          for (int d = 0; d < 3; d++) {
            dx[d] = t.xp[d] - t.xm[d];
            xa[d] = __fmaf_rn(.5f, t.xp[d] + t.xm[d], -(float)t.lg[d]);
          }
          if (!XYZ) {
            dx[0] = t.v[0] * G.pc.dt * G.pc.dxi_idx[0];
          }
          q = cross ? 0.f : U.w;
        }
        // park cell-crossing particles for the split/deposit walk
        const unsigned cm = __ballot_sync(FULL, cross);
        if (cm) {
          if (cross) {
            const int slot = qn + __popc(cm & lt);
            const float fi = __int_as_float((int)i);
            if constexpr (SAME) {
              // (xm | i, qw), (xp, i | vx)
              myQ[2 * slot] = make_float4(XYZ ? t.xm[0] : fi, t.xm[1], t.xm[2], U.w);
              myQ[2 * slot + 1] = make_float4(t.xp[0], t.xp[1], t.xp[2], XYZ ? fi : t.v[0]);
            } else {
              // (x_old, qw), (x_new, - | vx)
              myQ[2 * slot] = make_float4(X.x, X.y, X.z, U.w);
              myQ[2 * slot + 1] = make_float4(x[0], x[1], x[2], XYZ ? 0.f : t.v[0]);
            }
          }
          qn += __popc(cm);
          __syncwarp();
        }
        const float h12 = (1.f / 12.f) * dx[0] * dx[1] * dx[2];
        // ---- one pass per cell that has particles in this chunk
        for (;;) {
          // lanes [lo, hi) of this chunk belong to the cell (warp-uniform bounds)
          const uint32_t hi = min(ce - base, 32u), lo = cb > base ? cb - base : 0u;
          const bool mine = (uint32_t)lane - lo < hi - lo;
          const float qe = mine ? q : 0.f;
          if (COUNT) {
            n_left += mine && cross;
          }
          {
            const float qh = qe * h12;
            if (XYZ) {
#pragma unroll
              for (int d = 0; d < 3; d++) {
                const float m = qe * dx[d];
                const float a = xa[(d + 1) % 3], b = xa[(d + 2) % 3];
                const float ma = m * a;
                acc[4 * d + 0] += m;
                acc[4 * d + 1] += ma;
                acc[4 * d + 2] = __fmaf_rn(m, b, acc[4 * d + 2]);
                acc[4 * d + 3] = __fmaf_rn(ma, b, acc[4 * d + 3] + qh);
              }
            } else {
              const float m0 = qe * dx[0], m1 = qe * dx[1], m2 = qe * dx[2];
              const float ma = m0 * xa[1];
              acc[0] += m0;
              acc[1] += ma;
              acc[2] = __fmaf_rn(m0, xa[2], acc[2]);
              acc[3] = __fmaf_rn(ma, xa[2], acc[3] + qh);
              acc[4] += m1;
              acc[5] = __fmaf_rn(m1, xa[2], acc[5]);
              acc[6] += m2;
              acc[7] = __fmaf_rn(m2, xa[1], acc[7]);
            }
          }
          if (ce > base + 32) {
            break; // the cell continues in the next chunk
          }
          // ---- the cell is complete: flush its moments as leaf values, count its stayers
          if (ce > cb) {
            flush_moments(jrow + cur * ROW_STRIDE);
            if (COUNT) {
              const uint32_t pop = ce - cb, left = __reduce_add_sync(FULL, n_left);
              if (lane == 0) {
                const size_t e = cen0 + (size_t)(c0 + cur);
                atomicAdd(cnt32 + (e >> 1), (pop - left) << (16 * (e & 1)));
                if (pop > CNT_MAX) {
                  atomicExch(&A.flags[0], 1u);
                }
              }
              n_left = 0;
            }
          }
          if (++cur == RUN) {
            break;
          }
          cb = ce;
          ce = __shfl_sync(FULL, myoff, cur + 1);
          if (cb >= base + 32) {
            break; // the next cell starts in the next chunk
          }
        }
        base += 32;
      } while (base < end);
    }
    } else if (begin < end) {
      // ---- two particles per lane: records base + lane (A) and base + 32 + lane (B)
      using pm::f2;
      using pm::i2;
      using pm::mk2;
      FldTile2<GeoStatic<DIM>> EM2{sEM, geo, n0, n1, n2};
      int cur = 0;
      uint32_t cb = begin, ce = __shfl_sync(FULL, myoff, 1);
      uint32_t n_left = 0;
      float acc[NM]; // this lane's share of the cell's moments (both particles)
#pragma unroll
      for (int n = 0; n < NM; n++) {
        acc[n] = 0.f;
      }
      auto flush_moments = [&](int at) {
        float v[NVP];
#pragma unroll
        for (int n = 0; n < NVP; n++) {
          v[n] = n < NM ? acc[n] : 0.f;
        }
        warp_transpose_reduce<NVP>(v, lane);
        const float leaf = moments_to_leaf<DIM>(v[0], lane) * my_fnq;
        if (writer) {
          red_shared_f32(sJ32 + 4u * (uint32_t)at, leaf);
        }
#pragma unroll
        for (int n = 0; n < NM; n++) {
          acc[n] = 0.f;
        }
      };
      // staging: x[0..63], p[0..63]; every lane copies and reads its own four slots
      auto prefetch = [&](uint32_t ia) {
        if (ia < end) {
          cp_async16(myP, A.xi4 + ia);
          cp_async16(myP + 64 * sizeof(float4), A.pxi4 + ia);
        }
        if (ia + 32 < end) {
          cp_async16(myP + 32 * sizeof(float4), A.xi4 + ia + 32);
          cp_async16(myP + 96 * sizeof(float4), A.pxi4 + ia + 32);
        }
        cp_async_commit();
      };
      prefetch(begin + lane);
      uint32_t base = begin;
      do {
        const uint32_t ia = base + lane, ib = ia + 32;
        const bool act_a = ia < end, act_b = ib < end;
        while (qn > QC - 64) {
          flush_moments(jrow + cur * ROW_STRIDE);
          drain(min(qn, 32));
        }
        cp_async_wait_all();
        const float4 XA = lds128(myP), UA = lds128(myP + 64 * sizeof(float4));
        float4 XB = lds128(myP + 32 * sizeof(float4)), UB = lds128(myP + 96 * sizeof(float4));
        prefetch(ia + 64);
        bool cross_a = false, cross_b = false;
        f2 dx[3], xa[3];
#pragma unroll
        for (int d = 0; d < 3; d++) {
          dx[d] = xa[d] = mk2(0.f, 0.f);
        }
        f2 q = mk2(0.f, 0.f);
        pm::TrajectoryT<f2> t;
        f2 x[3];
        if (act_a) {
          if (!act_b) {
            XB = XA, UB = UA; // (computed, never stored or deposited)
          }
          x[0] = mk2(XA.x, XB.x), x[1] = mk2(XA.y, XB.y), x[2] = mk2(XA.z, XB.z);
          f2 u[3] = {mk2(UA.x, UB.x), mk2(UA.y, UB.y), mk2(UA.z, UB.z)};
          pm::advance<DIM>(G.pc, EM2, x, u, i2{__float_as_int(XA.w), __float_as_int(XB.w)}, t);
          A.xi4[ia] = make_float4(x[0].v.x, x[1].v.x, x[2].v.x, XA.w);
          A.pxi4[ia] = make_float4(u[0].v.x, u[1].v.x, u[2].v.x, UA.w);
          if (act_b) {
            A.xi4[ib] = make_float4(x[0].v.y, x[1].v.y, x[2].v.y, XB.w);
            A.pxi4[ib] = make_float4(u[0].v.y, u[1].v.y, u[2].v.y, UB.w);
          }
          cross_a = (XYZ && t.lf[0].x != t.lg[0].x) || t.lf[1].x != t.lg[1].x || t.lf[2].x != t.lg[2].x;
          cross_b = (XYZ && t.lf[0].y != t.lg[0].y) || t.lf[1].y != t.lg[1].y || t.lf[2].y != t.lg[2].y;
          if constexpr (!SAME) {
            const float xoa[3] = {XA.x, XA.y, XA.z}, xob[3] = {XB.x, XB.y, XB.z};
#pragma unroll
            for (int d = XYZ ? 0 : 1; d < 3; d++) {
              cross_a = cross_a || pm::cell_position(G.pc, xoa[d], d) != t.lg[d].x;
              cross_b = cross_b || pm::cell_position(G.pc, xob[d], d) != t.lg[d].y;
            }
          }
          cross_b = cross_b && act_b;
          // (J is compared at 1e-5: contraction is welcome from here on)
#pragma unroll
          for (int d = 0; d < 3; d++) {
            dx[d] = t.xp[d] - t.xm[d];
            xa[d] = pm::fma2(pm::bc2(.5f), t.xp[d] + t.xm[d], -pm::to_real(t.lg[d]));
          }
          if (!XYZ) {
            dx[0] = f2{__fmul2_rn(t.v[0].v, make_float2(G.pc.dt * G.pc.dxi_idx[0], G.pc.dt * G.pc.dxi_idx[0]))};
          }
          q = mk2(cross_a ? 0.f : UA.w, (cross_b || !act_b) ? 0.f : UB.w);
        }
        // park cell-crossing particles for the split/deposit walk
        const unsigned cma = __ballot_sync(FULL, cross_a), cmb = __ballot_sync(FULL, cross_b);
        if (cma | cmb) {
          if (cross_a) {
            const int slot = qn + __popc(cma & lt);
            const float fi = __int_as_float((int)ia);
            if constexpr (SAME) {
              myQ[2 * slot] = make_float4(XYZ ? t.xm[0].v.x : fi, t.xm[1].v.x, t.xm[2].v.x, UA.w);
              myQ[2 * slot + 1] = make_float4(t.xp[0].v.x, t.xp[1].v.x, t.xp[2].v.x, XYZ ? fi : t.v[0].v.x);
            } else {
              myQ[2 * slot] = make_float4(XA.x, XA.y, XA.z, UA.w);
              myQ[2 * slot + 1] = make_float4(x[0].v.x, x[1].v.x, x[2].v.x, XYZ ? 0.f : t.v[0].v.x);
            }
          }
          if (cross_b) {
            const int slot = qn + __popc(cma) + __popc(cmb & lt);
            const float fi = __int_as_float((int)ib);
            if constexpr (SAME) {
              myQ[2 * slot] = make_float4(XYZ ? t.xm[0].v.y : fi, t.xm[1].v.y, t.xm[2].v.y, UB.w);
              myQ[2 * slot + 1] = make_float4(t.xp[0].v.y, t.xp[1].v.y, t.xp[2].v.y, XYZ ? fi : t.v[0].v.y);
            } else {
              myQ[2 * slot] = make_float4(XB.x, XB.y, XB.z, UB.w);
              myQ[2 * slot + 1] = make_float4(x[0].v.y, x[1].v.y, x[2].v.y, XYZ ? 0.f : t.v[0].v.y);
            }
          }
          qn += __popc(cma) + __popc(cmb);
          __syncwarp();
        }
        const f2 h12 = f2{__fmul2_rn(__fmul2_rn(__fmul2_rn(dx[0].v, make_float2(1.f / 12.f, 1.f / 12.f)), dx[1].v), dx[2].v)};
        // ---- one pass per cell that has particles in this chunk
        for (;;) {
          // lanes [lo, hi) of each half belong to the cell (warp-uniform bounds)
          const uint32_t rb = cb > base ? cb - base : 0u, re = ce - base; // cell range relative to the chunk
          const uint32_t lo_a = min(rb, 32u), hi_a = min(re, 32u);
          const uint32_t lo_b = rb > 32u ? min(rb - 32u, 32u) : 0u, hi_b = re > 32u ? min(re - 32u, 32u) : 0u;
          const bool mine_a = (uint32_t)lane - lo_a < hi_a - lo_a, mine_b = (uint32_t)lane - lo_b < hi_b - lo_b;
          const f2 qe = mk2(mine_a ? q.v.x : 0.f, mine_b ? q.v.y : 0.f);
          if (COUNT) {
            n_left += (mine_a && cross_a) + (mine_b && cross_b);
          }
          {
            // packed products for both particles, summed into the lane's scalar moments (twelve
            // packed accumulators would cost twelve more registers, i.e. occupancy)
            const float2 qh = __fmul2_rn(qe.v, h12.v);
            auto add2 = [](float& a, float2 v) { a += v.x + v.y; };
            if (XYZ) {
#pragma unroll
              for (int d = 0; d < 3; d++) {
                const float2 m = __fmul2_rn(qe.v, dx[d].v);
                const float2 a = xa[(d + 1) % 3].v, b = xa[(d + 2) % 3].v;
                const float2 ma = __fmul2_rn(m, a);
                add2(acc[4 * d + 0], m);
                add2(acc[4 * d + 1], ma);
                add2(acc[4 * d + 2], __fmul2_rn(m, b));
                add2(acc[4 * d + 3], __ffma2_rn(ma, b, qh));
              }
            } else {
              const float2 m0 = __fmul2_rn(qe.v, dx[0].v), m1 = __fmul2_rn(qe.v, dx[1].v), m2 = __fmul2_rn(qe.v, dx[2].v);
              const float2 ma = __fmul2_rn(m0, xa[1].v);
              add2(acc[0], m0);
              add2(acc[1], ma);
              add2(acc[2], __fmul2_rn(m0, xa[2].v));
              add2(acc[3], __ffma2_rn(ma, xa[2].v, qh));
              add2(acc[4], m1);
              add2(acc[5], __fmul2_rn(m1, xa[2].v));
              add2(acc[6], m2);
              add2(acc[7], __fmul2_rn(m2, xa[1].v));
            }
          }
          if (ce > base + 64) {
            break; // the cell continues in the next chunk
          }
          if (ce > cb) {
            flush_moments(jrow + cur * ROW_STRIDE);
            if (COUNT) {
              const uint32_t pop = ce - cb, left = __reduce_add_sync(FULL, n_left);
              if (lane == 0) {
                const size_t e = cen0 + (size_t)(c0 + cur);
                atomicAdd(cnt32 + (e >> 1), (pop - left) << (16 * (e & 1)));
                if (pop > CNT_MAX) {
                  atomicExch(&A.flags[0], 1u);
                }
              }
              n_left = 0;
            }
          }
          if (++cur == RUN) {
            break;
          }
          cb = ce;
          ce = __shfl_sync(FULL, myoff, cur + 1);
          if (cb >= base + 64) {
            break; // the next cell starts in the next chunk
          }
        }
        base += 64;
      } while (base < end);
    }
    if (lane == 0) {
      row = atomicAdd(&row_ctr, 1);
    }
    row = __shfl_sync(FULL, row, 0);
  }
  while (qn > 0) {
    drain(min(qn, 32));
  }
  __syncthreads();

  // ---- flush the J tile (halo included) with global reductions
  for (int idx = tid; idx < 3 * NODES; idx += NW * 32) {
    const float v = sJ[idx];
    if (v != 0.f) {
      const int m = idx / NODES;
      int rem = idx - m * NODES;
      const int kz = rem / SZ;
      rem -= kz * SZ;
      const int ky = rem / SY, kx = rem - ky * SY;
      const int gi = n0 + kx, gj = n1 + ky, gk = n2 + kz;
      if (gi < G.ldims[0] + G.ibn[0] && gj < G.ldims[1] + G.ibn[1] && gk < G.ldims[2] + G.ibn[2]) {
        atomicAdd(F + fld_off(G, m, gi, gj, gk), v);
      }
    }
  }
}

} // namespace lean
