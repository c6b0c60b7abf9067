// The reference's BoundaryInjector integration tests
// (src/libpsc/tests/test_boundary_injector.cxx:106-283) written against the psc_b200 wrapper
// types the way the originals are written against PscConfig1vbecDouble: same grid
// (setupGrid, :45-74), same test generator (:76-104), Psc::step with add_injector, the
// continuity and Gauss checks after every step, the same final assertions.  The final state
// is dumped for tests/test_gpu_cxx.py to compare with the CPU oracle.
//
//   test_injector <out.bin> <one_particle|many_particles|many_species> <fused 0|1>
//
// Needs a GPU at run time; compiling and linking it is the CPU-side check of the header.
#include "mini_grid.hxx"

#include <psc_b200/psc_config_b200.hxx>

#include <cstdio>
#include <cstring>
#include <string>

using Grid = mini::Grid;
using Config = psc_b200::PscConfig<mini::dim_yz, Grid>;
using Mparticles = Config::Mparticles;
using MfieldsState = Config::MfieldsState;

// float state: the checks hold to rounding (the reference's thresholds are for its double config)
static const double CHECK_EPS = 1e-5;

// struct ParticleGenerator of the reference's test
struct ParticleGenerator
{
  ParticleGenerator(int max_n_injected, int kind_idx) : max_n_injected(max_n_injected), kind_idx(kind_idx) {}

  mini::Inject get(mini::Real3 min_pos, mini::Real3 pos_range)
  {
    double uy = 2.0;
    if (max_n_injected > 0 && n_injected++ >= max_n_injected) {
      // uy = 0: the particle does not enter the domain and is not injected
      uy = 0.0;
    }
    mini::Inject prt;
    prt.x = {min_pos[0], min_pos[1] + pos_range[1] * .999, min_pos[2]};
    prt.u = {0.0, uy, 0.0};
    prt.w = 1.0;
    prt.kind = kind_idx;
    return prt;
  }

  int n_injected = 0;
  int max_n_injected;
  int kind_idx;
};

using Injector = psc_b200::BoundaryInjectorB200<ParticleGenerator, Grid>;

int main(int argc, char** argv)
{
  if (argc < 4) {
    std::fprintf(stderr, "usage: %s out.bin one_particle|many_particles|many_species fused\n", argv[0]);
    return 1;
  }
  const std::string which = argv[2];
  const bool fused = std::atoi(argv[3]) != 0;

  mini::BC bc;
  bc.fld_lo = bc.fld_hi = {PSC_B200_BND_FLD_PERIODIC, PSC_B200_BND_FLD_OPEN, PSC_B200_BND_FLD_PERIODIC};
  bc.prt_lo = bc.prt_hi = {PSC_B200_BND_PRT_PERIODIC, PSC_B200_BND_PRT_OPEN, PSC_B200_BND_PRT_PERIODIC};
  Grid grid({1, 8, 2}, {1., 8., 2.}, {1, 1, 1}, 1., {{-1., 1., "e"}, {1., 1., "i"}}, 1, bc);

  MfieldsState mflds{grid};
  Mparticles mprts{grid};

  psc_b200::PscParamsB200 params;
  params.fused = fused;
  psc_b200::ChecksParamsB200 checks_params;
  checks_params.continuity_every_step = 1;
  checks_params.gauss_every_step = 1;
  psc_b200::Step<Config> psc(grid, mflds, mprts, params, checks_params);

  Injector inject_ions{ParticleGenerator(which == "one_particle" ? 1 : -1, 1), grid};
  Injector inject_electrons{ParticleGenerator(-1, 0), grid};
  psc.add_injector(&inject_ions);
  if (which == "many_species") {
    psc.add_injector(&inject_electrons);
  }

  if (mprts.size() != 0) {
    return 2;
  }
  psc.initialize();
  const int nmax = 2;
  while (psc.timestep() < nmax) {
    psc();
    const double cont = psc.checks().continuity.last_max_err, gauss = psc.checks().gauss.last_max_err;
    if (!(cont < CHECK_EPS) || !(gauss < CHECK_EPS)) {
      std::fprintf(stderr, "step %d: continuity %g gauss %g\n", psc.timestep(), cont, gauss);
      return 3;
    }
  }

  std::vector<psc_b200::Particle> prts;
  std::vector<uint32_t> off;
  mprts.get(prts, off);
  bool found_electrons = false, found_ions = false;
  for (const auto& prt : prts) {
    found_electrons |= prt.kind == 0;
    found_ions |= prt.kind == 1;
  }
  if (which == "one_particle" && prts.size() != 1) {
    return 4;
  }
  if (which != "one_particle" && !(prts.size() > 1)) {
    return 5;
  }
  if (which == "many_species" && !(found_electrons && found_ions)) {
    return 6;
  }

  auto flds = mflds.download(0, PSC_B200_NR_FIELDS);
  FILE* f = std::fopen(argv[1], "wb");
  if (!f) {
    return 7;
  }
  const int hdr[4] = {grid.n_patches(), (int)prts.size(), (int)flds.size(), 0};
  std::fwrite(hdr, sizeof(int), 4, f);
  std::fwrite(off.data(), sizeof(uint32_t), off.size(), f);
  std::fwrite(prts.data(), sizeof(psc_b200::Particle), prts.size(), f);
  std::fwrite(flds.data(), sizeof(float), flds.size(), f);
  std::fclose(f);
  std::printf("ok: %zu particles after %d steps (%d + %d entered in the last one)\n", prts.size(), nmax,
              inject_ions.n_injected(), which == "many_species" ? inject_electrons.n_injected() : 0);
  return 0;
}
