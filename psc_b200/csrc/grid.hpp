// psc_b200: host-side grid description derived from psc_b200_grid_desc.
// Pure C++ (no CUDA) so the same code serves the CUDA context and the CPU-side
// unit tests of the host logic.
//
// Restates what PSC's Grid_t / Domain / MrcDomain derive from the same inputs:
//   grid/domain.hxx:25-47   ldims, dx, dx_inv
//   grid.hxx:68-101         patches (off, xb, xe), invariant dims forced periodic
//   mrc_domain.hxx:18-29    a dim is topologically periodic iff fld BC periodic && gdims>1
//   mrc_domain_lib.c:21-35  "bydim" patch order;  mrc_domain_multi.c:162-193 rank ranges
//   mrc_domain_multi.c:518-545 neighbour lookup
#pragma once

#include "../../include/psc_b200.h"
#include "pic_math.cuh"

#include <cassert>
#include <cmath>
#include <string>
#include <vector>

namespace psc_b200
{

struct GridHost
{
  psc_b200_grid_desc desc;
  int ldims[3], ibn[3], im[3], invar[3], periodic[3];
  double dx[3], dx_inv[3];
  int dim;     // pm::DIM_XYZ / pm::DIM_YZ
  int deposit; // pm::DEPOSIT_*
  int n_patches_global;
  int n_cells; // per patch
  long fld_len; // im0*im1*im2
  std::vector<int> patch_off_by_rank; // n_ranks + 1
  int rank, n_ranks;
  int patch_begin, n_patches; // local range

  int rank_of_patch(int gp) const
  {
    for (int r = 0; r < n_ranks; r++) {
      if (gp < patch_off_by_rank[r + 1]) {
        return r;
      }
    }
    return -1;
  }

  void patch_idx3(int gp, int idx3[3]) const
  {
    idx3[0] = gp % desc.np[0];
    idx3[1] = (gp / desc.np[0]) % desc.np[1];
    idx3[2] = gp / (desc.np[0] * desc.np[1]);
  }

  void patch_off(int gp, int off[3]) const
  {
    int idx3[3];
    patch_idx3(gp, idx3);
    for (int d = 0; d < 3; d++) {
      off[d] = idx3[d] * ldims[d];
    }
  }

  // mrc_domain_multi.c:518-545; returns global patch or -1
  int neighbor_patch(int gp, const int dir[3]) const
  {
    int idx3[3], nei[3];
    patch_idx3(gp, idx3);
    for (int d = 0; d < 3; d++) {
      nei[d] = idx3[d] + dir[d];
      if (periodic[d]) {
        if (nei[d] < 0) {
          nei[d] += desc.np[d];
        }
        if (nei[d] >= desc.np[d]) {
          nei[d] -= desc.np[d];
        }
      }
      if (nei[d] < 0 || nei[d] >= desc.np[d]) {
        return -1;
      }
    }
    return (nei[2] * desc.np[1] + nei[1]) * desc.np[0] + nei[0];
  }

  bool at_boundary_lo(int gp, int d) const
  { // grid.hxx:115
    int off[3];
    patch_off(gp, off);
    return off[d] == 0;
  }
  bool at_boundary_hi(int gp, int d) const
  { // grid.hxx:116-119
    int off[3];
    patch_off(gp, off);
    return off[d] + ldims[d] == desc.gdims[d];
  }
};

inline bool grid_setup(const psc_b200_grid_desc& desc, GridHost& g, std::string& err)
{
  g.desc = desc;
  g.n_patches_global = 1;
  for (int d = 0; d < 3; d++) {
    if (desc.gdims[d] <= 0 || desc.np[d] <= 0 || desc.gdims[d] % desc.np[d] != 0) {
      err = "gdims must be positive and divisible by np (grid/domain.hxx:29-39)";
      return false;
    }
    if (!(desc.length[d] > 0.)) {
      err = "non-positive domain length (grid/domain.hxx:41-43)";
      return false;
    }
    g.ldims[d] = desc.gdims[d] / desc.np[d];
    g.dx[d] = desc.length[d] / double(desc.gdims[d]);
    g.dx_inv[d] = double(desc.gdims[d]) / desc.length[d];
    g.invar[d] = desc.gdims[d] == 1;
    g.ibn[d] = g.invar[d] ? 0 : 2;
    g.im[d] = g.ldims[d] + 2 * g.ibn[d];
    if (g.invar[d]) { // grid.hxx:90-98
      g.desc.bc_fld_lo[d] = g.desc.bc_fld_hi[d] = PSC_B200_BND_FLD_PERIODIC;
      g.desc.bc_prt_lo[d] = g.desc.bc_prt_hi[d] = PSC_B200_BND_PRT_PERIODIC;
    }
    for (int bc : {g.desc.bc_fld_lo[d], g.desc.bc_fld_hi[d]}) {
      if (bc != PSC_B200_BND_FLD_PERIODIC && bc != PSC_B200_BND_FLD_CONDUCTING_WALL &&
          bc != PSC_B200_BND_FLD_OPEN) {
        // the reference itself asserts on BND_FLD_ABSORBING (psc_bnd_fields_impl.hxx:48-50);
        // running with untouched ghost cells would be wrong physics
        err = "field boundary condition BND_FLD_ABSORBING is not implemented (nor is it in PSC)";
        return false;
      }
    }
    g.periodic[d] =
      g.desc.bc_fld_lo[d] == PSC_B200_BND_FLD_PERIODIC && desc.gdims[d] > 1;
    g.n_patches_global *= desc.np[d];
  }
  if (!g.invar[0] && !g.invar[1] && !g.invar[2]) {
    g.dim = pm::DIM_XYZ;
  } else if (g.invar[0] && !g.invar[1] && !g.invar[2]) {
    g.dim = pm::DIM_YZ;
  } else {
    err = "only dim_xyz and dim_yz geometries are implemented";
    return false;
  }
  g.deposit = desc.deposit;
  if (g.deposit == PSC_B200_DEPOSIT_DEFAULT) { // psc_config.hxx:47-72
    g.deposit = g.dim == pm::DIM_YZ ? pm::DEPOSIT_VAR1 : pm::DEPOSIT_SPLIT;
  }
  if (g.deposit == pm::DEPOSIT_VAR1 && g.dim != pm::DIM_YZ) {
    err = "Current1vbVar1 exists for dim_yz only (inc_curr_1vb_var1.cxx:162-166)";
    return false;
  }
  if (desc.n_kinds < 1 || desc.n_kinds > PSC_B200_MAX_KINDS) {
    err = "n_kinds out of range (push_particles_1vb.hxx:11)";
    return false;
  }
  g.n_cells = g.ldims[0] * g.ldims[1] * g.ldims[2];
  g.fld_len = (long)g.im[0] * g.im[1] * g.im[2];

  g.n_ranks = desc.n_ranks > 0 ? desc.n_ranks : 1;
  g.rank = desc.rank;
  if (g.rank < 0 || g.rank >= g.n_ranks) {
    err = "rank out of range";
    return false;
  }
  g.patch_off_by_rank.assign(g.n_ranks + 1, 0);
  if (desc.n_patches_by_rank) {
    for (int r = 0; r < g.n_ranks; r++) {
      g.patch_off_by_rank[r + 1] = g.patch_off_by_rank[r] + desc.n_patches_by_rank[r];
    }
    if (g.patch_off_by_rank[g.n_ranks] != g.n_patches_global) {
      err = "n_patches_by_rank does not sum to the number of global patches";
      return false;
    }
  } else { // mrc_domain_multi.c:181-189
    int per = g.n_patches_global / g.n_ranks, rem = g.n_patches_global % g.n_ranks;
    for (int r = 0; r < g.n_ranks; r++) {
      g.patch_off_by_rank[r + 1] = g.patch_off_by_rank[r] + per + (r < rem);
    }
  }
  g.patch_begin = g.patch_off_by_rank[g.rank];
  g.n_patches = g.patch_off_by_rank[g.rank + 1] - g.patch_begin;
  if (g.n_patches < 1) { // psc_balance_impl.hxx:104-105
    err = "every rank needs at least one patch";
    return false;
  }
  g.desc.n_patches_by_rank = nullptr; // borrowed pointer, do not keep
  return true;
}

// every narrowing point of the reference
inline pm::PushConst make_push_const(const GridHost& g)
{
  pm::PushConst c{};
  const auto& D = g.desc;
  for (int d = 0; d < 3; d++) {
    c.dxi[d] = 1.f / float(g.dx[d]);
    c.dxi_idx[d] = float(g.dx_inv[d]);
    c.fnqs_split[d] = float(D.fnqs / D.dt) * float(g.dx[d]);
    c.fnq_var1[d] = float(g.dx[d] * D.fnqs / D.dt);
  }
  c.dt = float(D.dt);
  for (int k = 0; k < D.n_kinds; k++) {
    c.dq_kind[k] = float(.5f * D.eta * D.dt * D.q[k] / D.m[k]);
  }
  return c;
}

inline pm::PatchBnd make_patch_bnd(const GridHost& g, int gp)
{
  pm::PatchBnd pb{};
  int off[3];
  g.patch_off(gp, off);
  pb.at_lo = pb.at_hi = 0;
  for (int d = 0; d < 3; d++) {
    double xb = double(off[d]) * g.dx[d] + g.desc.corner[d];
    double xe = double(off[d] + g.ldims[d]) * g.dx[d] + g.desc.corner[d];
    pb.patch_size[d] = float(xe - xb);
    pb.ldims[d] = g.ldims[d];
    pb.at_lo |= g.at_boundary_lo(gp, d) << d;
    pb.at_hi |= g.at_boundary_hi(gp, d) << d;
    pb.bc_lo[d] = g.desc.bc_prt_lo[d];
    pb.bc_hi[d] = g.desc.bc_prt_hi[d];
  }
  return pb;
}

// Yee coefficients (psc_push_fields_impl.hxx:31-43)
struct YeeConst
{
  float dth, cnx, cny, cnz;
};

inline YeeConst make_yee_const(const GridHost& g, double dt_fac)
{
  YeeConst y;
  y.dth = float(dt_fac * g.desc.dt);
  y.cnx = g.invar[0] ? 0.f : float(double(y.dth) / g.dx[0]);
  y.cny = g.invar[1] ? 0.f : float(double(y.dth) / g.dx[1]);
  y.cnz = g.invar[2] ? 0.f : float(double(y.dth) / g.dx[2]);
  return y;
}

// Balance: best_mapping (psc_balance_impl.hxx:99-160)
std::vector<int> best_mapping(const std::vector<double>& capability,
                              const std::vector<double>& loads);

} // namespace psc_b200
