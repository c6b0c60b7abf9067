// Drives the PscConfig wrapper types (include/psc_b200/psc_config_b200.hxx) the way a
// PSC case deck + Psc<PscConfig> would (psc_bubble_yz.cxx:117-147,290-340; psc.hxx:321-486):
// build a grid, construct Mparticles / MfieldsState, set fields with a lambda, inject
// particles patch by patch, run N steps twice -- once operator by operator, once through
// the fused entry point -- and dump the final state for the Python test to compare with
// the CPU oracle (tests/test_gpu_cxx.py).
//
//   test_wrappers <out.bin> <xyz|yz> <n_steps> <fused 0|1>
//
// Needs a GPU at run time; compiling and linking it is the CPU-side check of the header.
#include "mini_grid.hxx"

#include <psc_b200/psc_config_b200.hxx>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>
#include <string>

using Grid = mini::Grid;

template <typename Dim>
static int run(const std::string& out, bool yz, int n_steps, bool fused)
{
  using Config = psc_b200::PscConfig<Dim, Grid>;
  using Mparticles = typename Config::Mparticles;
  using MfieldsState = typename Config::MfieldsState;

  const int ppc = 6;
  Grid grid(yz ? mini::Int3{1, 16, 32} : mini::Int3{16, 8, 16},
            yz ? mini::Real3{1., 20., 30.} : mini::Real3{16., 8., 16.},
            yz ? mini::Int3{1, 2, 2} : mini::Int3{2, 1, 2}, yz ? 0.3 : 0.4,
            {{-1., 1., "e"}, {1., 100., "i"}}, ppc);

  Mparticles mprts{grid};
  MfieldsState mflds{grid};

  // setupFields with a lambda (psc_bubble_yz.cxx:248-285)
  mflds.setup([&](int m, double x[3]) {
    switch (m) {
      case PSC_B200_EY: return 0.05 * std::sin(0.4 * x[2]);
      case PSC_B200_EZ: return 0.03 * std::cos(0.3 * x[1]);
      case PSC_B200_HX: return 0.1;
      case PSC_B200_HZ: return 0.02 * std::sin(0.2 * x[1] + 0.1 * x[0]);
      default: return 0.;
    }
  });

  // inject patch by patch (setup_particles.hxx:263-331 does the same through injector())
  {
    std::mt19937 rng(42);
    std::uniform_real_distribution<double> uni(0.02, 0.98);
    std::normal_distribution<double> nrm(0., 1.);
    auto inj = mprts.injector();
    for (int p = 0; p < grid.n_patches(); p++) {
      auto injp = inj[p];
      const auto& patch = grid.patches[p];
      for (int k = 0; k < grid.ldims[2]; k++) {
        for (int j = 0; j < grid.ldims[1]; j++) {
          for (int i = 0; i < grid.ldims[0]; i++) {
            for (int n = 0; n < 2 * ppc; n++) {
              int kind = n & 1;
              double vth = kind ? 0.02 : 0.2;
              mini::Inject prt{{patch.xb[0] + (i + (yz ? 0.5 : uni(rng))) * grid.domain.dx[0],
                                patch.xb[1] + (j + uni(rng)) * grid.domain.dx[1],
                                patch.xb[2] + (k + uni(rng)) * grid.domain.dx[2]},
                               {vth * nrm(rng), vth * nrm(rng), vth * nrm(rng)},
                               1.,
                               kind};
              injp(prt);
            }
          }
        }
      }
    }
  } // injector flushes here

  // initial state for the oracle
  std::vector<psc_b200::Particle> prts0;
  std::vector<uint32_t> off0;
  mprts.get(prts0, off0);
  std::vector<float> flds0 = mflds.download(0, PSC_B200_NR_FIELDS);

  psc_b200::PscParamsB200 prm;
  prm.sort_interval = 2;
  prm.marder_interval = 0;
  prm.fused = fused;
  psc_b200::ChecksParamsB200 cprm;
  cprm.continuity_every_step = 1;
  cprm.continuity_threshold = 1e-4;
  psc_b200::Step<Config> step(grid, mflds, mprts, prm, cprm);
  // OutputFields / OutputMoments as a deck sets them up (psc_bubble_yz.cxx:322-337): fields
  // every 2 steps, time averages over the last 3 steps written every 4
  psc_b200::OutputFieldsItemParamsB200 outf_prm;
  outf_prm.pfield.out_interval = 2;
  outf_prm.tfield.out_interval = 4;
  outf_prm.tfield.average_length = 3;
  psc_b200::OutputFieldsB200<Grid> outf{grid, outf_prm};
  psc_b200::OutputMomentsB200<Grid> outm{grid, outf_prm};
  step.add_diagnostic(&outf);
  step.add_diagnostic(&outm);
  step.initialize();
  step.perform_diagnostics(); // Psc::integrate: initial output (psc.hxx:252-254)
  double max_cont = 0.;
  for (int n = 0; n < n_steps; n++) {
    step();
    step.perform_diagnostics();
    max_cont = std::fmax(max_cont, step.checks().continuity.last_max_err);
  }

  // accessor vocabulary (const_accessor_simple.hxx:47-81)
  double sum_w = 0.;
  {
    auto acc = mprts.accessor();
    for (int p = 0; p < grid.n_patches(); p++) {
      for (auto prt : acc[p]) {
        sum_w += prt.w();
        auto pos = prt.position();
        if (!(pos[1] >= grid.patches[p].xb[1] - 1e-3 && pos[1] <= grid.patches[p].xe[1] + 1e-3)) {
          std::fprintf(stderr, "particle outside its patch\n");
          return 2;
        }
      }
    }
  }

  std::vector<psc_b200::Particle> prts1;
  std::vector<uint32_t> off1;
  mprts.get(prts1, off1);
  std::vector<float> flds1 = mflds.download(0, PSC_B200_NR_FIELDS);
  auto en = psc_b200::energies(mprts);

  FILE* f = std::fopen(out.c_str(), "wb");
  if (!f) {
    return 3;
  }
  auto wr = [&](const void* p, size_t n) { std::fwrite(p, 1, n, f); };
  int hdr[8] = {grid.n_patches(), (int)prts0.size(), (int)prts1.size(), (int)flds0.size(),
                grid.domain.gdims[0], grid.domain.gdims[1], grid.domain.gdims[2], n_steps};
  wr(hdr, sizeof(hdr));
  wr(off0.data(), off0.size() * 4);
  wr(prts0.data(), prts0.size() * 32);
  wr(flds0.data(), flds0.size() * 4);
  wr(off1.data(), off1.size() * 4);
  wr(prts1.data(), prts1.size() * 32);
  wr(flds1.data(), flds1.size() * 4);
  double tail[10] = {max_cont, sum_w, en[0], en[1], en[2], en[3], en[4], en[5], en[6], en[7]};
  wr(tail, sizeof(tail));
  // Collision as Psc<PscConfig> uses it (psc.hxx:363-366): off in this run; the host round
  // trip (PSC's CollisionCudaHost pattern) with an operator that leaves the momenta alone
  // must hand back exactly the same store
  {
    typename Config::Collision collision{grid, 0, 0.1};
    if (collision.interval() > 0) {
      return 5;
    }
    auto nop = [](std::vector<psc_b200::Particle>& prts, const std::vector<uint32_t>& off) {
      (void)prts;
      (void)off;
    };
    psc_b200::CollisionViaHostB200<Grid, decltype(nop)> coll_host{grid, 10, 0.1, nop};
    coll_host(mprts);
    std::vector<psc_b200::Particle> prts2;
    std::vector<uint32_t> off2;
    mprts.get(prts2, off2);
    if (off2 != off1 || std::memcmp(prts2.data(), prts1.data(), prts1.size() * sizeof(prts1[0])) != 0) {
      std::fprintf(stderr, "host round trip changed the store\n");
      return 6;
    }
  }
  // device-side moments through the ItemMoment-shaped wrappers
  // (fields_item_moments_1st.hxx:9-30): density per kind and the 13-component set
  {
    psc_b200::Moment_n_1st_B200<Grid> mom_n{grid};
    psc_b200::Moments_1st_B200<Grid> mom_all{grid};
    auto& mn = mom_n(mprts);
    auto& ma = mom_all(mprts);
    std::vector<float> hn = mn.download(0, mn.n_comps()), ha = ma.download(0, ma.n_comps());
    int nn[2] = {(int)hn.size(), (int)ha.size()};
    wr(nn, sizeof(nn));
    wr(hn.data(), hn.size() * 4);
    wr(ha.data(), ha.size() * 4);
    if (mom_n.name() != "n_1st_cc" || mom_all.name() != "all_1st_cc" || mn.n_comps() != 2 || ma.n_comps() != 26) {
      return 4;
    }
  }
  // what the writers were handed: pfd at steps 0, 2, 4, one tfd (mean over steps 2..4) at step 4
  {
    auto& pf = outf.io_pfd().steps;
    auto& tf = outf.io_tfd().steps;
    auto& tm = outm.io_tfd().steps;
    if (n_steps == 4) {
      if (pf.size() != 3 || pf[0].item.timestep != 0 || pf[1].item.timestep != 2 || pf[2].item.timestep != 4 ||
          tf.size() != 1 || tm.size() != 1 || tf[0].item.timestep != 4 || outf.io_pfd().pfx != "pfd" ||
          outm.io_tfd().pfx != "tfd_moments" || tf[0].name != "jeh" || tm[0].name != "all_1st_cc" ||
          tf[0].comp_names.size() != 9 || tf[0].comp_names[3] != "ex_ec" || tm[0].comp_names.size() != 26 ||
          tm[0].comp_names[0] != "rho_e" || tm[0].comp_names[14] != "jx_i" || tf[0].item.n_comps != 9 ||
          tf[0].item.ldims != grid.ldims) {
        std::fprintf(stderr, "OutputFields: the writers were not handed what the cadence says\n");
        return 12;
      }
      int no[3] = {(int)tf[0].item.data.size(), (int)tm[0].item.data.size(), (int)pf[1].item.data.size()};
      wr(no, sizeof(no));
      wr(tf[0].item.data.data(), tf[0].item.data.size() * 4);
      wr(tm[0].item.data.data(), tm[0].item.data.size() * 4);
      wr(pf[1].item.data.data(), pf[1].item.data.size() * 4);
    } else {
      int no[3] = {0, 0, 0};
      wr(no, sizeof(no));
    }
  }
  std::fclose(f);

  // one more step written with the exact call shapes of Psc::step (psc.hxx:340-486) and of a
  // deck's constructors (psc_bubble_yz.cxx:296-320) -- a compile-and-run check that the
  // operator types are source-compatible with it
  {
    struct PscCheckParams // the fields of CheckParams (include/checks_params.hxx:3-29)
    {
      int check_interval = 1;
      double err_threshold = 1e-4;
      bool print_max_err_always = false;
      bool dump_always = false;
      bool exit_on_failure = false;
    };
    struct PscChecksParams
    {
      PscCheckParams continuity, gauss;
    } checks_params;
    checks_params.gauss.err_threshold = 1e30; // the random start is not Gauss-consistent
    typename Config::Checks checks_{grid, 0 /* MPI_Comm */, checks_params};
    typename Config::Collision collision_{grid, 1, 0.1}; // device binary collisions, every step
    typename Config::Sort sort_;
    typename Config::PushParticles pushp_;
    typename Config::BndParticles bndp_{grid};
    typename Config::Bnd bnd_;
    typename Config::BndFields bndf;
    typename Config::PushFields pushf_;
    const int timestep = n_steps + 1;
    {
      // psc.hxx:346: balance_(grid_, mprts_) -- one rank: nothing moves, the grid stays
      typename Config::Balance balance_{1.};
      Grid* grid_ptr = &grid;
      balance_(grid_ptr, mprts);
      if (grid_ptr != &grid) {
        return 8;
      }
    }
    {
      // what Balance does after patches have moved (psc_balance_impl.hxx:893-1016): the host
      // Grid_t is REPLACED and every container is reset(new_grid).  The device context must
      // follow the new grid object with its particles and fields, not be re-created empty.
      Grid new_grid = grid;
      psc_b200_ctx* ctx_before = mprts.ctx();
      const int n_before = mprts.size();
      const auto e_before = mflds.download(PSC_B200_EX, PSC_B200_EX + 1);
      mprts.reset(new_grid);
      mflds.reset(new_grid);
      if (mprts.ctx() != ctx_before || mflds.ctx() != ctx_before || &mprts.grid() != &new_grid ||
          &mflds.grid() != &new_grid || mprts.size() != n_before ||
          mflds.download(PSC_B200_EX, PSC_B200_EX + 1) != e_before) {
        std::fprintf(stderr, "regrid: the containers did not follow the new grid with their data\n");
        return 9;
      }
      // a container constructed on the new grid attaches to the same context
      typename Config::Mparticles mprts2{new_grid};
      if (mprts2.ctx() != ctx_before || mprts2.size() != n_before) {
        return 10;
      }
      // ... and back (the rest of the test keeps using `grid`)
      mprts.reset(grid);
      mflds.reset(grid);
      if (mprts.ctx() != ctx_before || &mflds.grid() != &grid) {
        return 11;
      }
    }
    sort_(mprts);
    if (collision_.interval() > 0 && timestep % collision_.interval() == 0) {
      collision_(mprts);
    }
    if (checks_.continuity.should_do_check(timestep)) {
      checks_.continuity.before_particle_push(mprts);
    }
    pushp_.push_mprts(mprts, mflds);
    bndp_(mprts);
    bndf.add_ghosts_J(mflds);
    bnd_.add_ghosts(mflds, PSC_B200_JXI, PSC_B200_JXI + 3);
    bnd_.fill_ghosts(mflds, PSC_B200_JXI, PSC_B200_JXI + 3);
    pushf_.push_H(mflds, .5, Dim{});
    bndf.fill_ghosts_H(mflds);
    bnd_.fill_ghosts(mflds, PSC_B200_HX, PSC_B200_HX + 3);
    pushf_.push_E(mflds, 1., Dim{});
    bndf.fill_ghosts_E(mflds);
    bnd_.fill_ghosts(mflds, PSC_B200_EX, PSC_B200_EX + 3);
    pushf_.push_H(mflds, .5, Dim{});
    bndf.fill_ghosts_H(mflds);
    bnd_.fill_ghosts(mflds, PSC_B200_HX, PSC_B200_HX + 3);
    if (checks_.continuity.should_do_check(timestep)) {
      checks_.continuity.after_particle_push(mprts, mflds);
    }
    if (checks_.gauss.should_do_check(timestep)) {
      checks_.gauss(mprts, mflds);
    }
    if (!(checks_.continuity.last_max_err < 1e-4) || (size_t)mprts.size() != prts1.size()) {
      std::fprintf(stderr, "Psc::step-shaped step: continuity %g\n", checks_.continuity.last_max_err);
      return 7;
    }
  }
  std::printf("ok: %zu particles, %d steps, continuity %.3g, sum w %.1f\n", prts1.size(), n_steps,
              max_cont, sum_w);
  return 0;
}

// OutputFieldsParamsTest.DoOut / Tfield_DoAccum (src/libpsc/tests/test_mfields_io.cxx:232-259)
// on the parameter structs of the wrapper header; no device needed
static int selftest_output_params()
{
#define EXPECT(cond)                                                                               \
  if (!(cond)) {                                                                                   \
    std::fprintf(stderr, "selftest: %s (line %d)\n", #cond, __LINE__);                             \
    return 1;                                                                                      \
  }
  {
    psc_b200::BaseOutputFieldItemParamsB200 prm;
    EXPECT(!prm.do_out(0)); // should be disabled
    prm.out_interval = 10;  // now enabled
    EXPECT(prm.do_out(0));
  }
  {
    psc_b200::OutputTfieldItemParamsB200 prm; // default: use every step between outs
    EXPECT(!prm.do_accum(0));                 // should be disabled
    prm.out_interval = 100;                   // now enabled
    EXPECT(prm.do_accum(0));                  // accum on out step itself
    prm.average_length = 50;
    EXPECT(!prm.do_accum(0));
    EXPECT(!prm.do_accum(50));
    EXPECT(prm.do_accum(51));
    EXPECT(prm.do_accum(52));
    EXPECT(prm.do_accum(53));
    prm.sample_interval = 2;
    EXPECT(!prm.do_accum(51));
    EXPECT(prm.do_accum(52));
    EXPECT(!prm.do_accum(53));
    EXPECT(prm.do_accum(100));
  }
#undef EXPECT
  std::printf("selftest ok\n");
  return 0;
}

int main(int argc, char** argv)
{
  if (argc == 2 && std::strcmp(argv[1], "selftest") == 0) {
    return selftest_output_params();
  }
  if (argc < 5) {
    std::fprintf(stderr, "usage: %s out.bin xyz|yz n_steps fused\n", argv[0]);
    return 1;
  }
  bool yz = std::strcmp(argv[2], "yz") == 0;
  int n_steps = std::atoi(argv[3]);
  bool fused = std::atoi(argv[4]) != 0;
  return yz ? run<mini::dim_yz>(argv[1], true, n_steps, fused)
            : run<mini::dim_xyz>(argv[1], false, n_steps, fused);
}
