// TEST INFRASTRUCTURE ONLY.  Compiles the product's per-particle device math
// (psc_b200/csrc/pic_math.cuh) and host grid logic (grid.hpp) for the HOST so the
// `-m "not gpu"` tests can compare that arithmetic bit for bit with the oracle
// without a GPU.  Nothing in psc_b200/ links or loads this; the product path
// only ever runs pic_math.cuh inside CUDA kernels.
#include "../../psc_b200/csrc/grid.hpp"

#include <cstring>

using namespace psc_b200;

namespace
{
struct Prt
{
  float x[3], u[3];
  int kind;
  float qni_wni;
};

struct FldAcc
{
  float* f;
  int im[3], ib[3];
  float operator()(int m, int i, int j, int k) const
  {
    return f[(((long)m * im[2] + (k - ib[2])) * im[1] + (j - ib[1])) * im[0] + (i - ib[0])];
  }
  void add(int m, int i, int j, int k, float v)
  {
    f[(((long)m * im[2] + (k - ib[2])) * im[1] + (j - ib[1])) * im[0] + (i - ib[0])] += v;
  }
};

template <int DIM>
void deposit_leaf(FldAcc& J, const int ci[3], const float* val)
{
  for (int n = 0; n < pm::LeafShape<DIM>::NV; n++) {
    int m, ox, oy, oz;
    pm::leaf_slot<DIM>(n, m, ox, oy, oz);
    J.add(m, ci[0] + ox, ci[1] + oy, ci[2] + oz, val[n]);
  }
}

template <int DIM>
void push_patch(const pm::PushConst& c, int deposit, FldAcc F, Prt* prts, unsigned n)
{
  memset(F.f, 0, sizeof(float) * 3 * F.im[0] * F.im[1] * F.im[2]);
  for (unsigned k = 0; k < n; k++) {
    Prt& p = prts[k];
    pm::Trajectory t;
    pm::advance<DIM>(c, F, p.x, p.u, p.kind, t);
    float val[12];
    int ci[3];
    if (deposit == pm::DEPOSIT_SPLIT) {
      pm::SplitWalker<DIM> w;
      w.begin(c, t);
      do {
        w.descend();
        pm::split_leaf<DIM>(c, p.qni_wni, w.a, w.b, ci, val);
        deposit_leaf<DIM>(F, ci, val);
      } while (w.pop());
    } else {
      pm::Var1Walker w;
      w.begin(c, t);
      while (w.n_left > 0) {
        w.next(c, p.qni_wni, ci, val);
        deposit_leaf<pm::DIM_YZ>(F, ci, val);
      }
    }
  }
}
} // namespace

extern "C" {

int hc_push_mprts(const psc_b200_grid_desc* desc, float* flds, void* prts,
                  const unsigned* off)
{
  GridHost g;
  std::string err;
  if (!grid_setup(*desc, g, err)) {
    return -1;
  }
  pm::PushConst c = make_push_const(g);
  for (int p = 0; p < g.n_patches; p++) {
    FldAcc F{flds + (long)p * g.fld_len * 9, {g.im[0], g.im[1], g.im[2]},
             {-g.ibn[0], -g.ibn[1], -g.ibn[2]}};
    Prt* P = static_cast<Prt*>(prts) + off[p];
    unsigned n = off[p + 1] - off[p];
    if (g.dim == pm::DIM_XYZ) {
      push_patch<pm::DIM_XYZ>(c, g.deposit, F, P, n);
    } else {
      push_patch<pm::DIM_YZ>(c, g.deposit, F, P, n);
    }
  }
  return 0;
}

// per particle: out_dir[3n..], out_flag = 0 inside / 1 slow path / 2 dropped;
// particles are modified in place like process_patch does
int hc_bnd_classify(const psc_b200_grid_desc* desc, void* prts, const unsigned* off,
                    int* out_dir, int* out_flag)
{
  GridHost g;
  std::string err;
  if (!grid_setup(*desc, g, err)) {
    return -1;
  }
  pm::PushConst c = make_push_const(g);
  for (int p = 0; p < g.n_patches; p++) {
    pm::PatchBnd pb = make_patch_bnd(g, g.patch_begin + p);
    Prt* P = static_cast<Prt*>(prts);
    for (unsigned k = off[p]; k < off[p + 1]; k++) {
      int dir[3];
      bool drop;
      int slow = pm::bnd_classify(c, pb, P[k].x, P[k].u, dir, drop);
      out_flag[k] = drop ? 2 : slow;
      for (int d = 0; d < 3; d++) {
        out_dir[3 * k + d] = dir[d];
      }
    }
  }
  return 0;
}

int hc_grid_info(const psc_b200_grid_desc* desc, int* n_patches, int* patch_begin,
                 int* ldims, int* nei27_of_first)
{
  GridHost g;
  std::string err;
  if (!grid_setup(*desc, g, err)) {
    return -1;
  }
  *n_patches = g.n_patches;
  *patch_begin = g.patch_begin;
  for (int d = 0; d < 3; d++) {
    ldims[d] = g.ldims[d];
  }
  int dir[3];
  for (dir[2] = -1; dir[2] <= 1; dir[2]++)
    for (dir[1] = -1; dir[1] <= 1; dir[1]++)
      for (dir[0] = -1; dir[0] <= 1; dir[0]++)
        nei27_of_first[pm::dir2idx(dir)] = g.neighbor_patch(g.patch_begin, dir);
  return 0;
}
}

extern "C" {

// neighbour table of this rank: for every local patch and each of the 27 directions the
// global neighbour patch (-1 none) and its owner rank -- the host logic the NCCL halo and
// particle exchanges are planned from (GridHost::neighbor_patch / rank_of_patch)
int hc_neighbor_table(const psc_b200_grid_desc* desc, int* nei_gp, int* nei_rank)
{
  GridHost g;
  std::string err;
  if (!grid_setup(*desc, g, err)) {
    return -1;
  }
  for (int p = 0; p < g.n_patches; p++) {
    for (int di = 0; di < 27; di++) {
      int dir[3] = {di % 3 - 1, (di / 3) % 3 - 1, di / 9 - 1};
      int gp = g.neighbor_patch(g.patch_begin + p, dir);
      nei_gp[p * 27 + di] = gp;
      nei_rank[p * 27 + di] = gp >= 0 ? g.rank_of_patch(gp) : -1;
    }
  }
  return g.n_patches;
}
}
