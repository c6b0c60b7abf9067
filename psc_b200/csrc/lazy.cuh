// psc_b200: the "lazy" particle store -- the sort that never runs as a pass of its own.
//
// The reference sorts by cell at the top of a step (SortCountsort2, psc_sort_impl.hxx:65-124)
// after exchanging particles at the end of the previous one (BndParticles,
// bnd_particles_impl.hxx:93-218).  Both only *reorder* records.  A particle moves at most
// one cell per step, so the cell-ordered sequence the next push must see is, for every
// target cell t, the concatenation of at most 27 (+1) groups, in the reference's order:
//     [stayers of the patch: source cells t - d in ascending cell order]
//     [arrivals from neighbour patches, receiver's direction loop ascending]
//     [arrivals from other ranks]
// where group (s, d) = the particles of source cell s that moved by offset d, in their
// old order.  So instead of materialising that sequence (one read + one write of every
// particle), the push kernel writes its output such that every group is a contiguous
// *segment*, and the next push gathers the segments while it loads:
//
//   B  run of source cell s = [V[s], V[s+1]):  stayers compacted at the front (old order);
//      the cell's movers behind them, reversed (scratch for the mover copy and nothing else)
//   M  mover array: the movers of s grouped by class d (old order inside a group), boundary
//      fix-ups applied; found through mbase[s] + pre[d][s] .. pre[d+1][s]
//   ncen[s]   number of stayers of s
//   V         run offsets = exclusive scan of the populations the cells had when the push
//             that wrote B ran;  Vnext = scan of the populations after it (accumulated with
//             atomics while pushing)
//
// lz_cell_segments() below enumerates the segments of one target cell in the reference's
// order; the push kernel (push.cu k_push_lazy) and the materialise kernel (fused_sort.cu)
// both build their gather tables from it.  The resulting particle order is bit-identical
// to BndParticles + SortCountsort2 (tests/test_gpu_sort_bnd.py, test_gpu_fields.py).
#pragma once

#include "fs_classify.cuh"

namespace psc_b200
{

// mover classes: 0..26 = cell offset class (13 = centre, unused), 27 = leaves for another
// rank; pre[q] for q = 0..28 (pre[28] = movers of the cell stored in M)
constexpr int LZ_Q_REMOTE = 27;
constexpr int LZ_PLANES = 29;
constexpr int LZ_UNIT = 4;                       // cells per work unit
constexpr int LZ_TAB = LZ_UNIT * 28;             // segment table entries per unit
constexpr unsigned LZ_FULL = 0xffffffffu;

enum
{
  LZ_SRC_B = 0, // run array
  LZ_SRC_M = 1, // mover array
  LZ_SRC_R = 2  // arrivals from other ranks
};

// the store being read
struct LzIn
{
  const uint32_t* vprev;  // run offsets of B [nct + 1]
  const uint32_t* ncen;   // stayers per run; nullptr: plain cell-ordered store (whole run)
  const uint32_t* mbase;  // [nct]
  const uint16_t* pre;    // [LZ_PLANES][nct]
  const uint32_t* rstart; // arrivals from other ranks per target cell; nullptr: none
  const uint32_t* rcount;
  const float4 *bx, *bp;  // B
  const float4 *mx, *mp;  // M
  const float4 *rx, *rp;  // R
  uint32_t nct;
};

// the store being written
struct LzOut
{
  float4 *bx, *bp; // B'
  float4 *mx, *mp; // M'
  uint32_t* ncen;
  uint32_t* mbase;
  uint16_t* pre;
  uint32_t* newpop;      // [nct] populations after this push (atomics)
  uint32_t* mov_counter; // M' allocation cursor
  uint32_t mov_cap;
};

struct LzSeg
{
  uint32_t vstart_tag; // start inside the unit's virtual sequence | source << 28
  uint32_t addr;       // first record in the source array
};

template <typename T>
__device__ __forceinline__ T lz_warp_incl_scan(T v, int lane)
{
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    T t = __shfl_up_sync(LZ_FULL, v, o);
    if (lane >= o) {
      v += t;
    }
  }
  return v;
}

// Segments of target cell (c0, c1, c2) of local patch p, one per lane (lanes 0..26 the
// offset classes, lane 27 the arrivals from other ranks).  On return a lane with len > 0
// holds its segment's source array, first record, start inside the cell's sequence
// (vstart) and rank among the cell's non-empty segments (eidx); n_ent / total are
// warp-uniform.  Warp-collective.
__device__ __forceinline__ void lz_cell_segments(const GridDev& G, const int* __restrict__ nei_patch,
                                                 const LzIn& in, int p, int c0, int c1, int c2, int lane,
                                                 uint32_t& len, uint32_t& addr, int& src, uint32_t& vstart,
                                                 int& eidx, int& n_ent, uint32_t& total)
{
  const int ld0 = G.ldims[0], ld1 = G.ldims[1], ld2 = G.ldims[2];
  len = 0, addr = 0, src = LZ_SRC_B;
  int key = 99;
  if (lane < 27) {
    // lane k: offset e = (e0, e1, e2), ascending k = ascending source cell (fused_sort.cu fs_route)
    const int e2 = 1 - lane / 9, e1 = 1 - (lane / 3) % 3, e0 = 1 - lane % 3;
    int x = c0 - e0, y = c1 - e1, z = c2 - e2;
    // travel direction t of a particle arriving through a patch face
    const int t0 = x < 0 ? 1 : (x >= ld0 ? -1 : 0);
    const int t1 = y < 0 ? 1 : (y >= ld1 ? -1 : 0);
    const int t2 = z < 0 ? 1 : (z >= ld2 ? -1 : 0);
    x += t0 * ld0, y += t1 * ld1, z += t2 * ld2;
    int ps = p;
    key = 0;
    if (t0 | t1 | t2) {
      // the sender sits in direction -t of the receiver; receiver's loop: direction ascending
      const int dip = ((-t2 + 1) * 3 + (-t1 + 1)) * 3 + (-t0 + 1);
      ps = nei_patch[p * 27 + dip];
      key = 1 + dip;
    }
    if (ps >= 0) {
      const uint32_t gs = (uint32_t)ps * G.n_cells + (uint32_t)((z * ld1 + y) * ld0 + x);
      const int q = ((e2 + 1) * 3 + e1 + 1) * 3 + e0 + 1; // class of the group
      if (q == CLS_CENTER) {
        const uint32_t b = __ldg(&in.vprev[gs]);
        addr = b;
        len = in.ncen ? __ldg(&in.ncen[gs]) : __ldg(&in.vprev[gs + 1]) - b;
      } else if (in.ncen) {
        const uint32_t a = __ldg(&in.pre[(size_t)q * in.nct + gs]);
        const uint32_t b = __ldg(&in.pre[(size_t)(q + 1) * in.nct + gs]);
        len = b - a;
        if (len) {
          addr = __ldg(&in.mbase[gs]) + a;
          src = LZ_SRC_M;
        }
      }
    }
  } else if (lane == 27 && in.rstart) {
    const uint32_t g = (uint32_t)p * G.n_cells + (uint32_t)((c2 * ld1 + c1) * ld0 + c0);
    len = __ldg(&in.rcount[g]);
    addr = __ldg(&in.rstart[g]);
    src = LZ_SRC_R;
    key = 98;
  }
  const unsigned lt = (1u << lane) - 1u;
  unsigned todo = __ballot_sync(LZ_FULL, len > 0);
  vstart = 0, eidx = 0, n_ent = 0, total = 0;
  while (todo) {
    const int mk = __reduce_min_sync(LZ_FULL, ((todo >> lane) & 1) ? key : 99);
    const unsigned grp = __ballot_sync(LZ_FULL, ((todo >> lane) & 1) && key == mk);
    const bool in_grp = (grp >> lane) & 1;
    const uint32_t mine = in_grp ? len : 0u;
    const uint32_t incl = lz_warp_incl_scan(mine, lane);
    if (in_grp) {
      vstart = total + incl - mine;
      eidx = n_ent + __popc(grp & lt);
    }
    total += __shfl_sync(LZ_FULL, incl, 31);
    n_ent += __popc(grp);
    todo &= ~grp;
  }
}

// record `i` of the unit's virtual sequence: binary search in the segment table
// (entries sorted by vstart, n_ent >= 1)
__device__ __forceinline__ void lz_lookup(const LzSeg* __restrict__ tab, int n_ent, uint32_t i, int& src,
                                          uint32_t& addr)
{
  int lo = 0, hi = n_ent; // invariant: vstart[lo] <= i < vstart[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if ((tab[mid].vstart_tag & 0x0fffffffu) <= i) {
      lo = mid;
    } else {
      hi = mid;
    }
  }
  const LzSeg s = tab[lo];
  src = (int)(s.vstart_tag >> 28);
  addr = s.addr + (i - (s.vstart_tag & 0x0fffffffu));
}

// target cell of class q particles that started in cell (s0, s1, s2) of patch p: global
// cell index, or 0xffffffff when it lies on another rank / outside the domain
__device__ __forceinline__ uint32_t lz_target_cell(const GridDev& G, const int* __restrict__ nei_patch,
                                                   int p, int s0, int s1, int s2, int q)
{
  const int ld0 = G.ldims[0], ld1 = G.ldims[1], ld2 = G.ldims[2];
  int x = s0 + q % 3 - 1, y = s1 + (q / 3) % 3 - 1, z = s2 + q / 9 - 1;
  const int d0 = x < 0 ? -1 : (x >= ld0 ? 1 : 0);
  const int d1 = y < 0 ? -1 : (y >= ld1 ? 1 : 0);
  const int d2 = z < 0 ? -1 : (z >= ld2 ? 1 : 0);
  x -= d0 * ld0, y -= d1 * ld1, z -= d2 * ld2;
  int pt = p;
  if (d0 | d1 | d2) {
    pt = nei_patch[p * 27 + ((d2 + 1) * 3 + d1 + 1) * 3 + d0 + 1];
    if (pt < 0) {
      return 0xffffffffu;
    }
  }
  return (uint32_t)pt * G.n_cells + (uint32_t)((z * ld1 + y) * ld0 + x);
}

} // namespace psc_b200
