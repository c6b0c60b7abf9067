// psc_b200: classification of a pushed particle relative to its source cell -- shared by
// the fused boundary+sort pass (fused_sort.cu) and the counting hook of the tiled push
// kernel (push.cu).  Arithmetic: ParticleIndexer::cellPosition (particle_indexer.hxx:71)
// and BndParticlesCommon::process_patch (bnd_particles_impl.hxx:93-218) via pic_math.cuh.
#pragma once

#include "dev_util.cuh"

namespace psc_b200
{

constexpr int CLS_CENTER = 13, CLS_DROP = 27, CLS_BAD = 28, CLS_REMOTE = 29, CLS_NONE = 31;
constexpr int FS_PLANES = 27; // cnt[class][cell], class = ((dz+1)*3 + dy+1)*3 + dx+1
// Every entry of the planes is bounded by the population of one cell (a count of its
// particles, later their offset inside the target cell): 16 bits.  A cell with more than
// 65535 particles raises the "precondition broken" flag and the step takes the general path.
using cnt_t = uint16_t;
constexpr uint32_t CNT_MAX = 0xffffu;

struct FsTables
{
  const pm::PatchBnd* pbs;
  const int* nei_patch; // [n_patches][27]: local patch, -1 none, -2-r rank r
};

// class of a pushed particle that started in cell (s0,s1,s2) of patch p; on return x/u
// carry the boundary fix-ups, (q, c) is the target patch and cell (for CLS_REMOTE:
// q = -2 - rank, c = direction index)
__device__ __forceinline__ int fs_classify(const GridDev& G, const FsTables& T, int p, int s0, int s1,
                                           int s2, float x[3], float u[3], int& q, int& c)
{
  int pos[3];
  bool valid = true;
#pragma unroll
  for (int d = 0; d < 3; d++) {
    pos[d] = pm::cell_position(G.pc, x[d], d);
    valid = valid && (unsigned)pos[d] < (unsigned)G.ldims[d];
  }
  int dir[3] = {0, 0, 0};
  q = p;
  if (!valid) {
    bool drop;
    pm::PatchBnd pb = T.pbs[p];
    pm::bnd_classify(G.pc, pb, x, u, dir, drop);
    if (drop) {
      return CLS_DROP;
    }
    if (dir[0] | dir[1] | dir[2]) {
      int nq = T.nei_patch[p * 27 + pm::dir2idx(dir)];
      if (nq == -1) {
        return CLS_DROP;
      }
      if (nq < 0) {
        // leaves for rank -2 - nq: shipped by the NCCL exchange (q = nq, c = direction)
        q = nq;
        c = pm::dir2idx(dir);
        return CLS_REMOTE;
      }
      q = nq;
    }
#pragma unroll
    for (int d = 0; d < 3; d++) {
      pos[d] = pm::cell_position(G.pc, x[d], d);
      if ((unsigned)pos[d] >= (unsigned)G.ldims[d]) {
        return CLS_BAD;
      }
    }
  }
  c = (pos[2] * G.ldims[1] + pos[1]) * G.ldims[0] + pos[0];
  int d0 = pos[0] - s0 + dir[0] * G.ldims[0];
  int d1 = pos[1] - s1 + dir[1] * G.ldims[1];
  int d2 = pos[2] - s2 + dir[2] * G.ldims[2];
  if ((unsigned)(d0 + 1) > 2u || (unsigned)(d1 + 1) > 2u || (unsigned)(d2 + 1) > 2u) {
    return CLS_BAD;
  }
  return ((d2 + 1) * 3 + d1 + 1) * 3 + d0 + 1;
}

} // namespace psc_b200
