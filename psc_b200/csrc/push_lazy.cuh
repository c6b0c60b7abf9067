// psc_b200: k_push_lazy -- the tiled push of push.cu reading and writing the lazy store of
// lazy.cuh: it gathers the cell-ordered particle sequence from (at most 28) segments per
// cell while it loads, and writes its output so that the *next* push can do the same.
// The boundary exchange + sort of the reference (BndParticles + SortCountsort2) therefore
// costs no pass over the particles at all.  Included by push.cu inside
// namespace psc_b200::PUSH_VARIANT (after k_push_tiled and its helpers).
//
// Per work unit (LZ_UNIT consecutive cells of a row, handed out dynamically):
//   1. segment table: lz_cell_segments() for each cell -> shared table sorted by position
//      in the unit's virtual sequence
//   2. chunks of 32 consecutive records of that sequence (full lanes): table lookup ->
//      cp.async from B / M / R, push, deposit exactly as k_push_tiled does; every record is
//      written to the run of its cell in B': stayers compacted from the front in their old
//      order, movers from the back; per cell: stayer count, class counts, the populations
//      of the target cells (atomics)
//   3. movers of the unit (5 %): re-read from the back of their runs, boundary fix-ups
//      applied, grouped by (cell, class) into a block of M' -- the segments the next
//      step's table points at

struct LazyArgs
{
  const uint32_t* v;  // current virtual cell offsets [nct + 1] = run offsets of B'
  float* flds;
  long slot_len;
  uint32_t* flags;    // [0] precondition broken, [1] dropped, [2] leaving for another rank,
                      // [3] mover array overflow
  int same_dxi;
  FsTables tab;
  LzIn in;
  LzOut out;
};

template <int DIM, int DEPOSIT, typename GEO, bool TMA, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB)
  k_push_lazy(const __grid_constant__ GridDev G, const __grid_constant__ GEO geo,
              const __grid_constant__ LazyArgs A)
{
  constexpr int NV = pm::LeafShape<DIM>::NV;
  constexpr int NVP = (DIM == pm::DIM_XYZ) ? 16 : 8;
  constexpr bool XYZ = DIM == pm::DIM_XYZ;
  extern __shared__ __align__(128) float smem[];
  __shared__ uint64_t bar;
  __shared__ int unit_ctr;
  const int nodes = geo.sm();
  const int n_warps = blockDim.x >> 5;
  float* sEM = smem;            // [6][f2][f1][f0]
  float* sJ = smem + 6 * nodes; // [3][f2][f1][f0]
  float4* sQ = reinterpret_cast<float4*>(smem + ((9 * nodes + 3) & ~3)); // [warps][QCAP][2]
  float4* sP = sQ + (size_t)n_warps * QCAP * 2;                          // [warps][2][32]
  LzSeg* sT = reinterpret_cast<LzSeg*>(sP + (size_t)n_warps * 64);       // [warps][LZ_TAB]
  uint16_t* sC = reinterpret_cast<uint16_t*>(sT + (size_t)n_warps * LZ_TAB); // [warps][LZ_UNIT][32]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned lt = (1u << lane) - 1u;
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
  const int n0 = o[0] - geo.g(0), n1 = o[1] - geo.g(1), n2 = o[2] - geo.g(2);

  if (tid == 0) {
    unit_ctr = n_warps;
  }
  // ---- stage E/B, zero J (as k_push_tiled)
  if (TMA) {
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
  float4* myQ = sQ + (size_t)warp * QCAP * 2;
  const uint32_t myP = smem_u32(sP + (size_t)warp * 64 + lane);
  LzSeg* myT = sT + (size_t)warp * LZ_TAB;
  uint16_t* myC = sC + (size_t)warp * LZ_UNIT * 32;
  int qn = 0; // queued trajectories of this warp (warp-uniform)

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

  // request record i of the unit's sequence into this lane's staging slot
  auto request = [&](int n_ent, uint32_t i) {
    int src;
    uint32_t addr;
    lz_lookup(myT, n_ent, i, src, addr);
    const float4* sx = src == LZ_SRC_B ? A.in.bx : (src == LZ_SRC_M ? A.in.mx : A.in.rx);
    const float4* sp = src == LZ_SRC_B ? A.in.bp : (src == LZ_SRC_M ? A.in.mp : A.in.rp);
    cp_async16(myP, sx + addr);
    cp_async16(myP + 32 * sizeof(float4), sp + addr);
  };

  const int n_rows = XYZ ? e[1] * e[2] : e[2];
  const int row_cells = XYZ ? e[0] : e[1];
  const int units_per_row = (row_cells + LZ_UNIT - 1) / LZ_UNIT;
  const int n_units = n_rows * units_per_row;
  for (int unit = warp; unit < n_units;) {
    const int row = unit / units_per_row;
    const int cs = (unit - row * units_per_row) * LZ_UNIT;
    const int run_cells = min(LZ_UNIT, row_cells - cs);
    int c0, rs0, rs1, rs2; // first cell of the unit: index in the patch and coordinates
    if (XYZ) {
      int ry = row % e[1], rz = row / e[1];
      rs0 = o[0] + cs, rs1 = o[1] + ry, rs2 = o[2] + rz;
      c0 = (rs2 * G.ldims[1] + rs1) * G.ldims[0] + rs0;
    } else {
      rs0 = 0, rs1 = o[1] + cs, rs2 = o[2] + row;
      c0 = rs2 * G.ldims[1] + rs1;
    }
    const uint32_t g0 = (uint32_t)p * G.n_cells + (uint32_t)c0;

    // ---- 1. segment table of the unit; lane j keeps the j-th cell boundary (myoff) of the
    // unit's virtual sequence and the run offset V[g0 + j] it writes to (myv)
    const uint32_t myv = __ldg(&A.v[g0 + min(lane, run_cells)]);
    uint32_t myoff = 0;
    int n_ent = 0;
    {
      uint32_t vbase = 0;
      for (int j = 0; j < run_cells; j++) {
        uint32_t len, addr, vstart, total;
        int src, eidx, ne;
        lz_cell_segments(G, A.tab.nei_patch, A.in, p, XYZ ? rs0 + j : 0, XYZ ? rs1 : rs1 + j, rs2, lane,
                         len, addr, src, vstart, eidx, ne, total);
        if (len) {
          myT[n_ent + eidx] = LzSeg{(vbase + vstart) | ((uint32_t)src << 28), addr};
        }
        n_ent += ne;
        vbase += total;
        if (lane == j + 1) {
          myoff = vbase;
        }
      }
      // the population the previous step predicted for a cell must be what its segments
      // add up to
      const uint32_t vnext = __shfl_down_sync(FULL, myv, 1);
      const uint32_t onext = __shfl_down_sync(FULL, myoff, 1);
      if (lane < run_cells && vnext - myv != onext - myoff) {
        atomicExch(&A.flags[0], 1u);
      }
    }
    __syncwarp();
    const uint32_t end = __shfl_sync(FULL, myoff, run_cells); // records of the unit

    int cur = 0;                                          // cell of the unit the passes are at
    uint32_t cb = 0, ce = __shfl_sync(FULL, myoff, 1);    // its range in the unit's sequence
    uint32_t wlo = __shfl_sync(FULL, myv, 0), whi = __shfl_sync(FULL, myv, 1); // its run in B'
    uint32_t nst = 0, nmv = 0;                            // stayers / movers written so far
    float acc[NV];
#pragma unroll
    for (int n = 0; n < NV; n++) {
      acc[n] = 0.f;
    }
    bool dirty = false;
    uint32_t mycount = 0; // lane k: particles of the current cell in class k

    if (lane < end) {
      request(n_ent, lane);
    }
    cp_async_commit();
    uint32_t base = 0;
    do {
      const uint32_t i = base + lane;
      const bool act = i < end;
      cp_async_wait_all();
      const float4 X = lds128(myP), U = lds128(myP + 32 * sizeof(float4));
      if (i + 32 < end) {
        request(n_ent, i + 32);
      }
      cp_async_commit();
      if (qn > QCAP - 32) {
        drain(min(qn, 32));
      }
      float val[NV];
      int ci[3] = {0, 0, 0};
      int pos[3] = {0, 0, 0};
      bool single = false;
      float4 Xo = X, Uo = U; // the pushed record
      {
        bool cross = false;
        pm::Trajectory t;
        if (act) {
          float x[3] = {X.x, X.y, X.z}, u[3] = {U.x, U.y, U.z};
          pm::advance<DIM>(G.pc, EM, x, u, __float_as_int(X.w), t);
          Xo = make_float4(x[0], x[1], x[2], X.w);
          Uo = make_float4(u[0], u[1], u[2], U.w);
          cross = (XYZ && t.lf[0] != t.lg[0]) || t.lf[1] != t.lg[1] || t.lf[2] != t.lg[2];
          single = !cross;
          if (single) {
            Walker<DIM, DEPOSIT> w;
            w.first(G.pc, t, U.w, ci, val);
          }
#pragma unroll
          for (int d = 0; d < 3; d++) {
            pos[d] = A.same_dxi ? t.lf[d] : pm::cell_position(G.pc, x[d], d);
          }
        }
        const unsigned cm = __ballot_sync(FULL, cross);
        if (cm) {
          if (cross) {
            const int slot = qn + __popc(cm & lt);
            myQ[2 * slot] = make_float4(t.xm[0], t.xm[1], t.xm[2], U.w);
            myQ[2 * slot + 1] = make_float4(t.xp[0], t.xp[1], t.xp[2], t.v[0]);
          }
          qn += __popc(cm);
          __syncwarp();
        }
      }
      // ---- one pass per cell that has records in this chunk
      for (;;) {
        const bool mine = act && i >= cb && i < ce;
        const int s0 = XYZ ? rs0 + cur : 0, s1 = XYZ ? rs1 : rs1 + cur, s2 = rs2;
        if (mine && single) {
          if (ci[0] == s0 && ci[1] == s1 && ci[2] == s2) {
#pragma unroll
            for (int n = 0; n < NV; n++) {
              acc[n] += val[n];
            }
            dirty = true;
          } else {
            leaf_deposit<DIM>(G, geo, sJ, F, n0, n1, n2, ci, val);
          }
        }
        // destination class relative to the cell of this pass
        int cls = CLS_NONE;
        if (mine) {
          const int d0 = pos[0] - s0, d1 = pos[1] - s1, d2 = pos[2] - s2;
          const bool ok = (unsigned)pos[0] < (unsigned)G.ldims[0] && (unsigned)pos[1] < (unsigned)G.ldims[1] &&
                          (unsigned)pos[2] < (unsigned)G.ldims[2] && (unsigned)(d0 + 1) <= 2u &&
                          (unsigned)(d1 + 1) <= 2u && (unsigned)(d2 + 1) <= 2u;
          if (ok) {
            cls = ((d2 + 1) * 3 + d1 + 1) * 3 + d0 + 1;
          } else {
            float xx[3] = {Xo.x, Xo.y, Xo.z}, uu[3] = {Uo.x, Uo.y, Uo.z};
            int q, c;
            cls = fs_classify(G, A.tab, p, s0, s1, s2, xx, uu, q, c);
            if (cls == CLS_CENTER) {
              // reflected back into its own cell: a stayer, stored with the fix-up applied
              // (movers get theirs when they are copied to M')
              Xo = make_float4(xx[0], xx[1], xx[2], Xo.w);
              Uo = make_float4(uu[0], uu[1], uu[2], Uo.w);
            }
          }
        }
        // the record goes to the run of its cell in B': stayers from the front, movers
        // from the back (the fix-ups of a mover are applied when it is copied to M')
        {
          const unsigned sm = __ballot_sync(FULL, cls == CLS_CENTER);
          const unsigned mm = __ballot_sync(FULL, mine && cls != CLS_CENTER);
          if (mine) {
            const uint32_t dst = cls == CLS_CENTER ? wlo + nst + __popc(sm & lt)
                                                   : whi - 1u - (nmv + __popc(mm & lt));
            A.out.bx[dst] = Xo;
            A.out.bp[dst] = Uo;
          }
          nst += __popc(sm);
          nmv += __popc(mm);
          // class counts: lane k counts class k
          unsigned rem = mm;
          while (rem) {
            const int v = __shfl_sync(FULL, cls, __ffs(rem) - 1);
            const unsigned grp = __ballot_sync(FULL, cls == v);
            if (lane == v) {
              mycount += __popc(grp);
            }
            rem &= ~grp;
          }
        }
        if (ce > base + 32) {
          break; // the cell continues in the next chunk
        }
        // ---- the cell is complete
        if (__any_sync(FULL, dirty)) {
          float v[NVP];
#pragma unroll
          for (int n = 0; n < NVP; n++) {
            v[n] = n < NV ? acc[n] : 0.f;
          }
          warp_transpose_reduce<NVP>(v, lane);
          const int my_slot = slot_of_lane<NVP>(lane);
          const int my_lin = (my_slot < NV) ? leaf_lin<DIM>(my_slot, geo.sy(), geo.sz(), geo.sm()) : 0;
          if ((my_slot < NV) && ((lane & (NVP == 16 ? 1 : 3)) == 0)) {
            atomicAdd(&sJ[(s2 - n2) * geo.sz() + (s1 - n1) * geo.sy() + (s0 - n0) + my_lin], v[0]);
          }
#pragma unroll
          for (int n = 0; n < NV; n++) {
            acc[n] = 0.f;
          }
          dirty = false;
        }
        {
          // stayer count, mover class counts (kept for step 3) and the populations of the
          // target cells
          const uint32_t g = g0 + cur;
          if (lane == CLS_CENTER) {
            A.out.ncen[g] = nst;
            if (nst) {
              atomicAdd(&A.out.newpop[g], nst);
            }
          } else if (lane < 27) {
            if (mycount) {
              const uint32_t tg = lz_target_cell(G, A.tab.nei_patch, p, s0, s1, s2, lane);
              if (tg != 0xffffffffu) {
                atomicAdd(&A.out.newpop[tg], mycount);
              } else {
                atomicExch(&A.flags[0], 1u); // classified into a patch that is not here
              }
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
          // table of the cell for step 3: slot q < 27 class q, slot 27 leaving for another rank
          const uint32_t rc = __shfl_sync(FULL, mycount, CLS_REMOTE);
          myC[cur * 32 + lane] = (uint16_t)(lane < 27 ? mycount : (lane == LZ_Q_REMOTE ? rc : 0u));
          if ((lane < 27 && mycount > 0xffffu) || nst + nmv != whi - wlo) {
            atomicExch(&A.flags[0], 1u);
          }
          if (lane == 31) {
            myC[cur * 32 + 31] = (uint16_t)min(nmv, 0xffffu); // back-region length of the cell
          }
          mycount = 0;
        }
        if (++cur == run_cells) {
          break;
        }
        cb = ce;
        ce = __shfl_sync(FULL, myoff, cur + 1);
        wlo = whi;
        whi = __shfl_sync(FULL, myv, cur + 1);
        nst = 0, nmv = 0;
      }
      base += 32;
    } while (base < end);
    __syncwarp();

    // ---- 3. movers of the unit -> M', grouped by (cell, class); metadata of the cells
    {
      // per cell: exclusive prefix over the classes; lane q holds class q of each cell
      uint32_t cellbase[LZ_UNIT]; // start of the cell's movers inside the unit's block
      uint32_t unit_total = 0, back_total = 0;
      uint32_t mypre[LZ_UNIT];
#pragma unroll
      for (int j = 0; j < LZ_UNIT; j++) {
        uint32_t cnt = (j < run_cells && lane < 28) ? myC[j * 32 + lane] : 0u;
        const uint32_t incl = lz_warp_incl_scan(cnt, lane);
        mypre[j] = incl - cnt;
        cellbase[j] = unit_total;
        unit_total += __shfl_sync(FULL, incl, 27);
        back_total += (j < run_cells) ? (uint32_t)myC[j * 32 + 31] : 0u;
      }
      uint32_t mb = 0;
      if (lane == 0 && unit_total) {
        mb = atomicAdd(A.out.mov_counter, unit_total);
      }
      mb = __shfl_sync(FULL, mb, 0);
      const bool fits = mb + unit_total <= A.out.mov_cap;
      if (!fits && lane == 0) {
        atomicExch(&A.flags[3], 1u);
      }
      __syncwarp();
#pragma unroll
      for (int j = 0; j < LZ_UNIT; j++) {
        if (j < run_cells) {
          const uint32_t g = g0 + j;
          if (lane < LZ_PLANES) {
            // lane 28 holds the total (its own count is 0)
            A.out.pre[(size_t)lane * A.in.nct + g] = (uint16_t)mypre[j];
          }
          if (lane == 0) {
            A.out.mbase[g] = mb + cellbase[j];
          }
          // running fill position of every (cell, class) group
          if (lane < 28) {
            myC[j * 32 + lane] = (uint16_t)mypre[j];
          }
        }
      }
      __syncwarp();
      // the movers sit at the back of their runs, first mover last
      uint32_t backs[LZ_UNIT + 1];
      backs[0] = 0;
#pragma unroll
      for (int j = 0; j < LZ_UNIT; j++) {
        backs[j + 1] = backs[j] + ((j < run_cells) ? (uint32_t)myC[j * 32 + 31] : 0u);
      }
      for (uint32_t mbase_i = 0; mbase_i < back_total && fits; mbase_i += 32) {
        const uint32_t m = mbase_i + lane;
        const bool am = m < back_total;
        int j = 0;
#pragma unroll
        for (int k = 1; k < LZ_UNIT; k++) {
          j += (m >= backs[k]) ? 1 : 0;
        }
        int key = -1 - lane; // distinct for idle lanes
        float4 Xm, Um;
        uint32_t dstbase = 0;
        if (am) {
          const uint32_t r = m - backs[j];
          const uint32_t whi_j = __ldg(&A.v[g0 + j + 1]);
          Xm = A.out.bx[whi_j - 1u - r];
          Um = A.out.bp[whi_j - 1u - r];
          float xx[3] = {Xm.x, Xm.y, Xm.z}, uu[3] = {Um.x, Um.y, Um.z};
          int q, c;
          const int cls = fs_classify(G, A.tab, p, XYZ ? rs0 + j : 0, XYZ ? rs1 : rs1 + j, rs2, xx, uu, q, c);
          const int slot = cls < 27 ? cls : (cls == CLS_REMOTE ? LZ_Q_REMOTE : -1);
          if (slot >= 0) {
            key = j * 32 + slot;
            Xm = make_float4(xx[0], xx[1], xx[2], Xm.w);
            Um = make_float4(uu[0], uu[1], uu[2], Um.w);
            dstbase = mb + cellbase[j];
          }
        }
        const unsigned grp = __match_any_sync(FULL, key);
        uint32_t r0 = 0;
        if (key >= 0) {
          r0 = myC[key];
        }
        __syncwarp();
        if (key >= 0 && lane == __ffs(grp) - 1) {
          myC[key] = (uint16_t)(r0 + __popc(grp));
        }
        __syncwarp();
        if (key >= 0) {
          const uint32_t dst = dstbase + r0 + __popc(grp & lt);
          A.out.mx[dst] = Xm;
          A.out.mp[dst] = Um;
        }
      }
    }
    __syncwarp();
    if (lane == 0) {
      unit = atomicAdd(&unit_ctr, 1);
    }
    unit = __shfl_sync(FULL, unit, 0);
  }
  while (qn > 0) {
    drain(min(qn, 32));
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
