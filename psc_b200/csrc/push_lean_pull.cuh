// psc_b200: k_push_lean_pull -- k_push_lean (push_lean.cuh) with the previous step's sort folded
// into it ("pull mode", option `pull`, off by default).  A separate copy of the
// kernel on purpose: k_push_lean sits exactly at its register budget (80 registers, 3 CTAs per
// SM, no spills) and every line shared with this variant moved its allocation.
#pragma once

namespace lean
{

// Split + deposit `cnt` queued trajectories of this warp (entries [qn - cnt, qn); PULL: the
// first cnt), one per lane; COUNT: add each one to the plane of its destination class; PULL:
// rank the ones that change cell and list them.  Returns the new queue length.
template <int DIM, int DEPOSIT, bool COUNT, bool SAME, bool PULL>
__device__ __forceinline__ int lean_drain_body(const GridDev& G, const GeoStatic<DIM>& geo, const PushArgs& A,
                                               float* sJ, float* F, float4* myQ, int n0, int n1, int n2, int p,
                                               int qn, int cnt)
{
  constexpr bool XYZ = DIM == pm::DIM_XYZ;
  constexpr int NM = XYZ ? 12 : 8;
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  uint32_t* const cnt32 = reinterpret_cast<uint32_t*>(A.cnt);
  const bool a2 = lane < cnt;
  // PULL walks the queue from its head (first in, first out: ranks in index order)
  const int q0 = PULL ? 0 : qn - cnt;
  pm::Trajectory t;
  float qw = 0.f;
  [[maybe_unused]] bool mover = false;     // PULL: changes cell inside this rank's patches
  [[maybe_unused]] uint32_t mv_e = 0;      // ... its counter: plane entry (class, source cell)
  [[maybe_unused]] uint32_t mv_i = 0;      // ... its index in the store
  float4 A0, A1;
  if (a2) {
    A0 = myQ[2 * (q0 + lane)], A1 = myQ[2 * (q0 + lane) + 1];
  }
  if constexpr (PULL) {
    // close the gap at the head (at most QC - 32 < 32 entries remain)
    const int rem = qn - cnt;
    float4 m0, m1;
    __syncwarp();
    if (lane < rem) {
      m0 = myQ[2 * (cnt + lane)], m1 = myQ[2 * (cnt + lane) + 1];
    }
    __syncwarp();
    if (lane < rem) {
      myQ[2 * lane] = m0, myQ[2 * lane + 1] = m1;
    }
  }
  if (a2) {
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
        if constexpr (PULL) {
          if (cls == CLS_CENTER) {
            // wrapped around a periodic direction and landed on the far edge, which the
            // exchange folds back to 0 (bnd_particles_impl.hxx:131-140): it stays in its cell
            // with the folded position.  Nobody rewrites a stayer in pull mode but us.
            const uint32_t i = (uint32_t)__float_as_int(XYZ ? A1.w : A0.x);
            A.xi4[i] = make_float4(xn[0], xn[1], xn[2], A.xi4[i].w);
          }
        }
      }
      if (cls < FS_PLANES) {
        const size_t e = (size_t)cls * A.nct + (size_t)p * G.n_cells +
                         (size_t)((sc[2] * G.ldims[1] + sc[1]) * G.ldims[0] + sc[0]);
        if (PULL && cls != CLS_CENTER) {
          mover = true; // (counted below, together with the others of its group)
          mv_e = (uint32_t)e;
          mv_i = (uint32_t)__float_as_int(XYZ ? A1.w : A0.x);
        } else {
          atomicAdd(cnt32 + (e >> 1), 1u << (16 * (e & 1)));
        }
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
  }
  if constexpr (PULL) {
    // rank of every mover inside its (source cell, class) group -- the lanes are in index
    // order, so are the walks of a cell's entries -- and its record in the mover list
    const unsigned mm = __ballot_sync(FULL, mover);
    if (mm) {
      uint32_t sb = 0;
      if (lane == __ffs(mm) - 1) {
        sb = atomicAdd(&A.flags[3], (uint32_t)__popc(mm));
      }
      sb = __shfl_sync(FULL, sb, __ffs(mm) - 1);
      if (mover) {
        const unsigned grp = __match_any_sync(mm, mv_e);
        const int leader = __ffs(grp) - 1;
        const int sh = 16 * (mv_e & 1);
        uint32_t old = 0;
        if (lane == leader) {
          old = atomicAdd(cnt32 + (mv_e >> 1), (uint32_t)__popc(grp) << sh);
        }
        old = __shfl_sync(grp, old, leader);
        const uint32_t rank = ((old >> sh) & 0xffffu) + __popc(grp & lt);
        const uint32_t slot = sb + __popc(mm & lt);
        if (slot < A.mv_cap) {
          A.mv_idx[slot] = mv_i;
          A.mv_key[slot] = make_uint2(mv_e, rank);
        }
      }
    }
  }
  // ---- the walk (nothing but the trajectory is live here)
  Walker<DIM, DEPOSIT> w;
  float val[NM];
  int ci[3] = {0, 0, 0};
  bool more = false;
  if (a2) {
    more = w.first(G.pc, t, qw, ci, val);
    leaf_deposit<DIM>(G, geo, sJ, F, n0, n1, n2, ci, val);
  }
  while (__any_sync(FULL, more)) {
    if (more) {
      more = w.next(G.pc, qw, ci, val);
      leaf_deposit<DIM>(G, geo, sJ, F, n0, n1, n2, ci, val);
    }
  }
  __syncwarp();
  return qn - cnt;
}

// the same out of line: the pull kernel calls it from three places and has no registers to
// spare around them (inlined, its stack frame outgrows the L1 cache: profiles/README.md)
template <int DIM, int DEPOSIT, bool COUNT, bool SAME, bool PULL>
__device__ __noinline__ int lean_drain_call(const GridDev& G, const GeoStatic<DIM>& geo, const PushArgs& A,
                                            float* sJ, float* F, float4* myQ, int n0, int n1, int n2, int p,
                                            int qn, int cnt)
{
  return lean_drain_body<DIM, DEPOSIT, COUNT, SAME, PULL>(G, geo, A, sJ, F, myQ, n0, n1, n2, p, qn, cnt);
}

// PULL (COUNT, SAME, W = 1): the push of step n + 1 doubles as the sort of step n.  The input
// store (A.xin4 / A.pin4, cell runs A.cell_off) is the previous push's output: ordered by the
// cells the particles were in BEFORE that push, so a run holds the particles that stayed in
// the cell ("live": their cell index still is the run's) and, dead, those that left it.  The
// output store (A.xi4 / A.pxi4, cell runs A.out_off) already holds every particle that changed
// cell, at its place in the reference's order (k_fs_place_movers): inside a cell
// [arrivals from lower cells | stayers | arrivals from higher cells and other patches].  For
// each row the kernel pushes the arrivals in front of the stayers where they lie, pulls the
// live particles of the input run through the push to out_off + n_before + rank, then pushes
// the arrivals behind them -- in this order, and the queue is walked first-in first-out, so
// that the particles leaving a cell in one direction are ranked in index order (the rank goes
// into the mover list for the next k_fs_place_movers).  One read and one write of every
// particle per step instead of two.
#ifndef LEAN_PULL_MINB
#define LEAN_PULL_MINB 2
#endif
template <int DIM, int DEPOSIT, bool COUNT = true, bool SAME = true, int W = 1, bool PULL = true>
__global__ void __launch_bounds__(n_warps<W>() * 32, W == 1 ? (PULL ? LEAN_PULL_MINB : 3) : 2)
  k_push_lean_pull(const __grid_constant__ CUtensorMap tm, const __grid_constant__ GridDev G,
              const __grid_constant__ GeoStatic<DIM> geo, const __grid_constant__ PushArgs A)
{
  static_assert(!PULL || (COUNT && SAME && W == 1), "PULL needs the class counts and the index in the queue");
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
  const uint32_t myP = smem_u32(sP + warp * 64 * W + lane);
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
    if constexpr (PULL) {
      qn = lean_drain_call<DIM, DEPOSIT, COUNT, SAME, PULL>(G, geo, A, sJ, F, myQ, n0, n1, n2, p, qn, cnt);
    } else {
      qn = lean_drain_body<DIM, DEPOSIT, COUNT, SAME, PULL>(G, geo, A, sJ, F, myQ, n0, n1, n2, p, qn, cnt);
    }
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
    // lane j < RUN holds where the stayers of the row's j-th cell go in the output store
    uint32_t my_sbase = 0;
    // Arrivals of the row.  Segments: lane j < RUN = the arrivals in front of cell j's stayers,
    // lane RUN + j = those behind them; a virtual index runs over the 2 RUN segments.  They are
    // pushed where they lie.  One that stays in its cell deposits its single leaf directly (the
    // moment formulas of the chunk loop for one particle) and is counted in the CENTER plane;
    // one that moves on is parked for the ordered walk: in front of the row's stayers if it lies
    // in front of them.  One that lies BEHIND them must be parked after them: the early pass
    // (row start) leaves such a record untouched and remembers its lane (late_mask), the late
    // pass (row end) pushes it again and parks it.  Rows with more than 32 arrivals push the
    // ones behind the stayers in the late pass altogether.
    static_assert(2 * RUN <= 32, "one lane per arrival segment of a row");
    uint32_t late_mask = 0;
    bool one_shot = true;
    auto arrival_pass = [&](const bool late) {
      if (late && one_shot && late_mask == 0) {
        return;
      }
      const int jj = lane < RUN ? lane : lane - RUN; // (lanes >= 2 RUN: empty segments)
      uint32_t seg_first = 0, seg_n = 0;
      if (lane < 2 * RUN) {
        const uint32_t* const ooff = A.out_off + (size_t)p * G.n_cells + c0 + jj;
        const uint32_t o_lo = __ldg(ooff), o_hi = __ldg(ooff + 1);
        const uint2 st = __ldg(&A.stay[(size_t)p * G.n_cells + c0 + jj]); // {n_before, n_stay}
        if (lane < RUN) {
          my_sbase = o_lo + st.x;
          seg_first = o_lo, seg_n = st.x;
        } else {
          seg_first = o_lo + st.x + st.y, seg_n = o_hi - seg_first;
        }
      }
      uint32_t incl = seg_n; // inclusive scan over the segments
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) {
          incl += v;
        }
      }
      const uint32_t total = __shfl_sync(FULL, incl, 31);
      if (!late) {
        one_shot = total <= 32;
      }
      const uint32_t vb0 = (late && !one_shot) ? __shfl_sync(FULL, incl, RUN - 1) & ~31u : 0u;
      for (uint32_t vb = vb0; vb < total; vb += 32) {
        if (qn > QC - 32) {
          drain(min(qn, 32));
        }
        const uint32_t v = vb + lane;
        int sg = 0; // segment of this lane's virtual index
#pragma unroll
        for (int k = 0; k < 2 * RUN - 1; k++) {
          sg += v >= __shfl_sync(FULL, incl, k);
        }
        const uint32_t ex = __shfl_sync(FULL, incl - seg_n, sg), f0 = __shfl_sync(FULL, seg_first, sg);
        const uint32_t i = f0 + (v - ex);
        const bool behind = sg >= RUN;
        const int j = behind ? sg - RUN : sg; // cell of the row
        bool doit = v < total;
        if (late) {
          doit = doit && behind && (one_shot ? ((late_mask >> lane) & 1u) != 0u : true);
        } else {
          doit = doit && (!behind || one_shot);
        }
        pm::Trajectory t;
        float qw = 0.f;
        bool park = false, defer = false;
        if (doit) {
          const float4 X = A.xi4[i], U = A.pxi4[i];
          float x[3] = {X.x, X.y, X.z}, u[3] = {U.x, U.y, U.z};
          pm::advance<DIM>(G.pc, EM, x, u, __float_as_int(X.w), t);
          const bool cross = (XYZ && t.lf[0] != t.lg[0]) || t.lf[1] != t.lg[1] || t.lf[2] != t.lg[2];
          qw = U.w;
          defer = !late && behind && cross; // (left as it is: the late pass pushes it again)
          if (!defer) {
            A.xi4[i] = make_float4(x[0], x[1], x[2], X.w);
            A.pxi4[i] = make_float4(u[0], u[1], u[2], U.w);
            park = cross;
            if (!cross) {
              // its leaf, from its moments (header comment of push_lean.cuh)
              float dx[3], xa[3];
#pragma unroll
              for (int d = 0; d < 3; d++) {
                dx[d] = t.xp[d] - t.xm[d];
                xa[d] = __fmaf_rn(.5f, t.xp[d] + t.xm[d], -(float)t.lg[d]);
              }
              if (!XYZ) {
                dx[0] = t.v[0] * G.pc.dt * G.pc.dxi_idx[0];
              }
              const float qh = qw * ((1.f / 12.f) * dx[0] * dx[1] * dx[2]);
              float* const b = sJ + (rs2 - n2) * SZ + (XYZ ? (rs1 - n1) * SY + (o0 - n0) : (o1 - n1) * SY) +
                               j * ROW_STRIDE;
              const float* const fnq = DEPOSIT == pm::DEPOSIT_SPLIT ? G.pc.fnqs_split : G.pc.fnq_var1;
              if (XYZ) {
#pragma unroll
                for (int d = 0; d < 3; d++) {
                  const float m = qw * dx[d];
                  const float sa = m * xa[(d + 1) % 3], sb = m * xa[(d + 2) % 3];
                  const float sab = __fmaf_rn(sa, xa[(d + 2) % 3], qh);
                  atomicAdd(b + leaf_lin<DIM>(4 * d + 0, SY, SZ, NODES), (((m - sa) - sb) + sab) * fnq[d]);
                  atomicAdd(b + leaf_lin<DIM>(4 * d + 1, SY, SZ, NODES), (sa - sab) * fnq[d]);
                  atomicAdd(b + leaf_lin<DIM>(4 * d + 2, SY, SZ, NODES), (sb - sab) * fnq[d]);
                  atomicAdd(b + leaf_lin<DIM>(4 * d + 3, SY, SZ, NODES), sab * fnq[d]);
                }
              } else {
                const float m0 = qw * dx[0], m1 = qw * dx[1], m2 = qw * dx[2];
                const float sa = m0 * xa[1], sb = m0 * xa[2];
                const float sab = __fmaf_rn(sa, xa[2], qh);
                atomicAdd(b + leaf_lin<DIM>(0, SY, SZ, NODES), (((m0 - sa) - sb) + sab) * fnq[0]);
                atomicAdd(b + leaf_lin<DIM>(1, SY, SZ, NODES), (sa - sab) * fnq[0]);
                atomicAdd(b + leaf_lin<DIM>(2, SY, SZ, NODES), (sb - sab) * fnq[0]);
                atomicAdd(b + leaf_lin<DIM>(3, SY, SZ, NODES), sab * fnq[0]);
                const float s1b = m1 * xa[2], s2a = m2 * xa[1];
                atomicAdd(b + leaf_lin<DIM>(4, SY, SZ, NODES), (m1 - s1b) * fnq[1]);
                atomicAdd(b + leaf_lin<DIM>(5, SY, SZ, NODES), s1b * fnq[1]);
                atomicAdd(b + leaf_lin<DIM>(6, SY, SZ, NODES), (m2 - s2a) * fnq[2]);
                atomicAdd(b + leaf_lin<DIM>(7, SY, SZ, NODES), s2a * fnq[2]);
              }
              const size_t e = cen0 + (size_t)(c0 + j);
              atomicAdd(cnt32 + (e >> 1), 1u << (16 * (e & 1)));
            }
          }
        }
        late_mask |= __ballot_sync(FULL, defer); // (warp-uniform)
        const unsigned pm_ = __ballot_sync(FULL, park);
        if (pm_) {
          if (park) {
            const int slot = qn + __popc(pm_ & lt);
            const float fi = __int_as_float((int)i);
            myQ[2 * slot] = make_float4(XYZ ? t.xm[0] : fi, t.xm[1], t.xm[2], qw);
            myQ[2 * slot + 1] = make_float4(t.xp[0], t.xp[1], t.xp[2], XYZ ? fi : t.v[0]);
          }
          qn += __popc(pm_);
          __syncwarp();
        }
      }
    };
#pragma unroll 1
    for (int ph = 0; ph < 2; ph++) {
    // ph 0: the arrivals (early pass), then the stayers; ph 1: the late pass (one copy of the
    // arrival code: the kernel has to stay inside the instruction cache)
    arrival_pass(ph == 1);
    if (ph == 1) {
      break;
    }
    if constexpr (W == 1) {
    if (begin < end) {
      int cur = 0;                                           // cell of the row the passes are at
      uint32_t cb = begin, ce = __shfl_sync(FULL, myoff, 1); // its particle range
      uint32_t n_left = 0;                                   // ... and whether this lane's particles left it (summed at the flush)
      [[maybe_unused]] uint32_t srun = 0;                    // PULL: live particles of the cell so far (= their rank base)
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
          atomicAdd(&sJ[at], leaf);
        }
#pragma unroll
        for (int n = 0; n < NM; n++) {
          acc[n] = 0.f;
        }
      };
      // the next chunk travels global -> shared with cp.async while this one is computed
      if (begin + lane < end) {
        cp_async16(myP, (PULL ? A.xin4 : A.xi4) + begin + lane);
        cp_async16(myP + 32 * sizeof(float4), (PULL ? A.pin4 : A.pxi4) + begin + lane);
      }
      cp_async_commit();
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
        cp_async_wait_all();
        const float4 X = lds128(myP), U = lds128(myP + 32 * sizeof(float4));
        if (i + 32 < end) {
          cp_async16(myP, (PULL ? A.xin4 : A.xi4) + i + 32);
          cp_async16(myP + 32 * sizeof(float4), (PULL ? A.pin4 : A.pxi4) + i + 32);
        }
        cp_async_commit();
        // ---- gather, Boris, move (the reference's arithmetic, pic_math.cuh)
        bool cross = false;
        float dx[3] = {0.f, 0.f, 0.f}, xa[3] = {0.f, 0.f, 0.f}; // displacement, centred offset
        float q = 0.f;                                         // q w of a particle that stayed in its cell
        pm::Trajectory t;
        float x[3] = {X.x, X.y, X.z};
        uint32_t di = i;  // where the pushed record goes (PULL: its place in the output store)
        bool live = act;  // PULL: the record still belongs to the cell of its run
        if constexpr (PULL) {
          // destination of the live records: the stayers' base of the lane's cell + the rank among
          // the cell's live records.  The loop visits the cells this chunk touches exactly like the
          // pass loop below, on copies of its state.  (Before the push: little is live here.)
          const int l1 = pm::fint(X.y * G.pc.dxi[1]), l2 = pm::fint(X.z * G.pc.dxi[2]);
          const int lr = XYZ ? pm::fint(X.x * G.pc.dxi[0]) : l1;
          const bool in_row = act && (XYZ ? (l1 == rs1 && l2 == rs2) : l2 == rs2);
          live = false;
          int cu = cur;
          uint32_t b2 = cb, e2 = ce, sr = srun;
          for (;;) {
            const uint32_t hi = min(e2 - base, 32u), lo = b2 > base ? b2 - base : 0u;
            const bool lv = ((uint32_t)lane - lo < hi - lo) && in_row && lr == (XYZ ? o0 : o1) + cu;
            const unsigned lm = __ballot_sync(FULL, lv);
            const uint32_t sb = __shfl_sync(FULL, my_sbase, cu);
            if (lv) {
              live = true;
              di = sb + sr + __popc(lm & lt);
            }
            if (e2 > base + 32 || ++cu == RUN) {
              break;
            }
            sr = 0;
            b2 = e2;
            e2 = __shfl_sync(FULL, myoff, cu + 1);
            if (b2 >= base + 32) {
              break;
            }
          }
          // (a record that left its cell a step ago is skipped: its copy among the arrivals is
          // the particle)
        }
        if (live) {
          float u[3] = {U.x, U.y, U.z};
          pm::advance<DIM>(G.pc, EM, x, u, __float_as_int(X.w), t);
          A.xi4[di] = make_float4(x[0], x[1], x[2], X.w);
          A.pxi4[di] = make_float4(u[0], u[1], u[2], U.w);
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
            const float fi = __int_as_float((int)di);
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
          if constexpr (PULL) {
            srun += __popc(__ballot_sync(FULL, mine && live));
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
              const uint32_t pop = PULL ? srun : ce - cb, left = __reduce_add_sync(FULL, n_left);
              if (lane == 0) {
                const size_t e = cen0 + (size_t)(c0 + cur);
                atomicAdd(cnt32 + (e >> 1), (pop - left) << (16 * (e & 1)));
                if (ce - cb > CNT_MAX) {
                  atomicExch(&A.flags[0], 1u);
                }
              }
              n_left = 0;
            }
          }
          if constexpr (PULL) {
            srun = 0;
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
    }
    } // ph
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
