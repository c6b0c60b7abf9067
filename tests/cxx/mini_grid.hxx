// Stand-alone stand-in for PSC's Grid_t (src/include/grid.hxx:68-160) with the member
// names the psc_b200 wrapper types read.  Test scaffolding: lets
// include/psc_b200/psc_config_b200.hxx compile and run in this repository, where PSC
// itself cannot be built.
#pragma once

#include <array>
#include <vector>

namespace mini
{

using Int3 = std::array<int, 3>;
using Real3 = std::array<double, 3>;

// grid/domain.hxx:25-47
struct Domain
{
  Int3 gdims, np, ldims;
  Real3 length, corner, dx;
};

// grid/BC.h:26-50
struct BC
{
  Int3 fld_lo{1, 1, 1}, fld_hi{1, 1, 1}, prt_lo{1, 1, 1}, prt_hi{1, 1, 1}; // periodic
};

// grid.hxx:18-33
struct Kind
{
  double q, m;
  const char* name;
};

// grid.hxx:265-293: dimensionless normalisation, fnqs = 1 / nicell
struct Normalization
{
  double fnqs = 1., eta = 1., cori = 1., prts_per_unit_density = 1.;
};

// grid.hxx:35-61
struct Patch
{
  Int3 off;
  Real3 xb, xe;
};

struct Grid
{
  Grid(Int3 gdims, Real3 length, Int3 np, double dt_, std::vector<Kind> kinds_, int nicell,
       BC bc_ = BC{})
    : bc(bc_), dt(dt_), kinds(std::move(kinds_))
  {
    domain.gdims = gdims;
    domain.np = np;
    domain.length = length;
    domain.corner = {0., 0., 0.};
    for (int d = 0; d < 3; d++) {
      domain.ldims[d] = gdims[d] / np[d];
      domain.dx[d] = length[d] / gdims[d];
      ibn[d] = gdims[d] == 1 ? 0 : 2;
    }
    ldims = domain.ldims;
    norm.fnqs = 1. / nicell;
    norm.cori = 1. / nicell; // grid.hxx:288
    norm.prts_per_unit_density = nicell; // grid.hxx:287
    // "bydim" patch order (libmrc/src/mrc_domain_lib.c:21-35)
    for (int pz = 0; pz < np[2]; pz++) {
      for (int py = 0; py < np[1]; py++) {
        for (int px = 0; px < np[0]; px++) {
          Patch p;
          Int3 idx{px, py, pz};
          for (int d = 0; d < 3; d++) {
            p.off[d] = idx[d] * ldims[d];
            p.xb[d] = p.off[d] * domain.dx[d] + domain.corner[d];
            p.xe[d] = (p.off[d] + ldims[d]) * domain.dx[d] + domain.corner[d];
          }
          patches.push_back(p);
        }
      }
    }
  }

  int n_patches() const { return (int)patches.size(); }
  bool isInvar(int d) const { return domain.gdims[d] == 1; }
  bool atBoundaryLo(int p, int d) const { return patches[p].off[d] == 0; } // grid.hxx:114

  Int3 ldims;
  Domain domain;
  BC bc;
  Normalization norm;
  double dt;
  std::vector<Patch> patches;
  std::vector<Kind> kinds;
  Int3 ibn;
};

// psc::particle::Inject (src/include/particle.h:22-36)
struct Inject
{
  Real3 x, u;
  double w;
  int kind;
};

struct dim_xyz
{};
struct dim_yz
{};

} // namespace mini
