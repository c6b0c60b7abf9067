// psc_b200: the "gapped" particle store -- BndParticles + SortCountsort2 without a pass of
// their own over the particles.
//
// The reference exchanges particles at the end of a step (BndParticles,
// bnd_particles_impl.hxx:93-218) and sorts them by cell at the top of the next one
// (SortCountsort2, psc_sort_impl.hxx:65-124).  Both only *reorder* records, and because a
// particle moves at most one cell per direction per step (|v| dt < dx) the cell-ordered
// sequence the next push must see is, for every target cell t,
//     [movers from the 13 lower neighbour cells of t's patch, ascending source cell]
//     [the particles that stayed in t, in their old order]
//     [movers from the 13 higher neighbour cells of the patch]
//     [arrivals through patch faces, receiver's direction loop ascending]
// (fused_sort.cu derives this order).  ~97 % of the particles are "stayers".  The gapped
// store therefore gives every cell a *slab* [V[t], V[t+1]) with some slack, the run of the
// cell being [start[t], start[t] + n[t]) inside it, and a step is
//   k_push_tiled<GAP>   reads the runs of one buffer; writes every stayer to its FINAL place
//                       in the other buffer, V[t] + RL + (rank among the stayers) -- RL slots
//                       are kept free in front for the lower movers; parks the movers (with
//                       the boundary fix-ups applied) in a tagged list M; counts the 27
//                       destination classes per source cell (as the eager fused path does)
//   k_gap_offsets       one thread per target cell: n_L, the new run start V + RL - n_L and
//                       length, the position of every (source cell, class) group inside
//                       the run; flags a slab that cannot take its arrivals
//   k_gap_place         one thread per mover: copy it to start[t] + group offset + rank
// so the particle data are read once and written once per step (64 B + 3 % instead of
// 136 B).  When a slab overflows, the same step re-lays the store out (fresh slack around
// the new populations, one extra copy).  The push never modifies the buffer it reads, so a
// step that cannot be finished on this path (a particle moved further than one cell, the
// mover list is full) is redone on the eager path.  Every operator that needs the
// reference's contiguous patch-by-patch array calls store_ready() -> gap_compact() first.
// The particle order is bit-identical to BndParticles + SortCountsort2
// (tests/test_gpu_gapped.py).
#pragma once

#include "fs_classify.cuh"

namespace psc_b200
{

constexpr int GAP_BATCH = 64; // mover slots a warp reserves at a time

// control words of one gapped step (device, cleared before the push)
enum
{
  GAP_CTL_MOVERS = 0,   // mover slots handed out
  GAP_CTL_OVERFLOW = 1, // a slab cannot take its arrivals
  GAP_CTL_MAX_NL = 2,   // max over cells of the lower-mover count
  GAP_CTL_M_FULL = 3,   // the mover list is full
  GAP_CTL_WORDS = 8
};

// what the gap variant of the tiled push kernel reads and writes
struct GapPush
{
  const uint32_t* in_start; // first record of every cell's run in the store being read [nct]
  const uint32_t* in_n;     // its length [nct]
  const float4 *in_x, *in_p;
  const uint32_t* out_v;    // slab starts of the store being written [nct + 1]
  uint32_t rl;              // slots kept free in front of the stayers
  float4 *out_x, *out_p;
  float4 *mx, *mp;          // mover list: records ...
  uint4* mtag;              // ... and {target cell, class * nct + source cell, rank, valid}
  uint32_t* ctl;
  uint32_t m_cap;
};

} // namespace psc_b200
