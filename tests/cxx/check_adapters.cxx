// Compile check (never linked, never run): include/psc_b200/psc_adapters_b200.hxx against the
// reference's REAL particles.hxx / particles_simple.hxx (tests/cxx/check_adapters.sh).
#include "psc_b200/psc_adapters_b200.hxx"

#include <type_traits>

using psc_b200::MparticlesB200Psc;
static_assert(std::is_base_of<MparticlesBase, MparticlesB200Psc>::value, "");
static_assert(!std::is_abstract<MparticlesB200Psc>::value, "every pure virtual of MparticlesBase is overridden");

// the calls a deck / Balance / the output writers make on the polymorphic base
int use(const Grid_t& grid, Grid_t* new_grid)
{
  MparticlesB200Psc mprts{grid};
  MparticlesBase& base = mprts;
  int n = base.size() + (int)base.sizeByPatch().size() + base.n_patches();
  auto& single = base.get_as<MparticlesSingle>();   // convert_to map
  base.put_as(single);                              // convert_from map
  auto& dbl = base.get_as<MparticlesDouble>();
  base.put_as(dbl, MP_DONT_COPY);
  base.reset(*new_grid);
  // the PscConfig operator types take the same object
  psc_b200::PushParticlesB200<Grid_t> pushp;
  psc_b200::MfieldsStateB200<Grid_t> mflds{grid};
  pushp.push_mprts(mprts, mflds);
  psc_b200::BalanceB200<Grid_t> balance{1.};
  balance(new_grid, mprts);
  return n;
}
