// psc_b200: the glue a PSC tree needs beyond the duck-typed PscConfig member types
// (psc_config_b200.hxx): PSC's polymorphic container bases.
//
//   MparticlesB200Psc   : MparticlesBase  (src/include/particles.hxx:19-110)
//       virtual size() / sizeByPatch() / reset(grid), and the convert_to() / convert_from()
//       maps keyed by type_index that MparticlesBase::get_as<MparticlesSingle>() / put_as()
//       walk (used by Balance psc_balance_impl.hxx:893-909, the particle output writers and
//       every deck that calls get_as).  Conversions are one bulk AoS transfer each way.
//   MfieldsStateB200Psc : MfieldsStateBase (src/include/fields3d.hxx:159-245)
//       registers in MfieldsStateBase::instances (so that Balance can reset() it,
//       psc_balance_impl.hxx:968-976), convert maps to MfieldsStateSingle / Double, and a
//       host-mirror storage() for the diagnostics that do
//       `gt::host_mirror(mflds.storage()); gt::copy(...)` (DiagEnergiesField.h:24-26).
//
// Unlike psc_config_b200.hxx this header INCLUDES PSC headers, so it only compiles inside a
// PSC tree.  What is verified in this repository (no gtensor, no MPI, no libmrc build here):
//   * the Mparticles half compiles against the reference's REAL particles.hxx,
//     particles_simple.hxx, psc_particles_single.h / _double.h (tests/cxx/check_adapters.sh:
//     declaration-only <mpi.h>, empty mrc_config.h / PscConfig.h, the 20-line gtensor shim
//     of oracle/shim)
//   * the Mfields half (PSC_B200_ADAPT_FIELDS) is NOT compile-checked here: fields3d.hxx
//     needs the real gtensor (first stop: src/include/fields3d.hxx:287 `gt::gtensor_span`,
//     then kg/SArrayView.h:46-61).  It is written against fields3d.hxx:159-245,415-464 and
//     psc_fields_cuda.h:78-113 and is the part a maintainer should expect to touch.
#pragma once

#include "psc_config_b200.hxx"

#include "particles.hxx"
#include "particles_simple.hxx"
#include "psc_particles_single.h"
#include "psc_particles_double.h"

#include <typeindex>

namespace psc_b200
{

// ----------------------------------------------------------------------
// MparticlesB200Psc

struct MparticlesB200Psc
  : MparticlesBase
  , MparticlesB200<Grid_t>
{
  using Impl = MparticlesB200<Grid_t>;
  using real_t = float;
  using Real3 = Vec3<real_t>;

  explicit MparticlesB200Psc(const Grid_t& grid) : MparticlesBase(grid), Impl(grid) {}

  // both bases know the grid: the device context is the authority after a rebalance
  const Grid_t& grid() const { return Impl::grid(); }
  int n_patches() const { return Impl::n_patches(); }

  int size() const override { return Impl::size(); }
  std::vector<uint> sizeByPatch() const override { return Impl::sizeByPatch(); }
  void reset(const Grid_t& grid) override
  {
    MparticlesBase::reset(grid);
    Impl::reset(grid);
  }

  static const Convert convert_to_, convert_from_;
  const Convert& convert_to() override { return convert_to_; }
  const Convert& convert_from() override { return convert_from_; }
};

namespace detail
{

// device store -> MparticlesSimple<P>: what MparticlesCuda's copy_to does one particle at a
// time through its accessor (psc_particles_cuda.cxx:41-63), as one bulk transfer
template <typename MP>
void copy_to(MparticlesBase& mprts_base, MparticlesBase& mprts_other_base)
{
  auto& mprts = dynamic_cast<MparticlesB200Psc&>(mprts_base);
  auto& other = dynamic_cast<MP&>(mprts_other_base);
  using oreal_t = typename MP::real_t;
  using OReal3 = typename MP::Real3;
  std::vector<Particle> prts;
  std::vector<uint32_t> off;
  mprts.get(prts, off);
  other.reserve_all(mprts.sizeByPatch());
  other.clear();
  for (int p = 0; p < mprts.n_patches(); p++) {
    for (uint32_t n = off[p]; n < off[p + 1]; n++) {
      const Particle& q = prts[n];
      other.push_back(p, {OReal3{oreal_t(q.x[0]), oreal_t(q.x[1]), oreal_t(q.x[2])},
                          OReal3{oreal_t(q.u[0]), oreal_t(q.u[1]), oreal_t(q.u[2])}, oreal_t(q.qni_wni),
                          q.kind, psc::particle::Id(), psc::particle::Tag()});
    }
  }
}

// MparticlesSimple<P> -> device store (psc_particles_cuda.cxx:15-39)
template <typename MP>
void copy_from(MparticlesBase& mprts_base, MparticlesBase& mprts_other_base)
{
  auto& mprts = dynamic_cast<MparticlesB200Psc&>(mprts_base);
  auto& other = dynamic_cast<MP&>(mprts_other_base);
  const auto n_by_patch = other.sizeByPatch();
  std::vector<Particle> prts;
  prts.reserve(other.size());
  auto accessor = other.accessor();
  for (int p = 0; p < mprts.n_patches(); p++) {
    for (auto prt : accessor[p]) {
      Particle q;
      const auto x = prt.x();
      const auto u = prt.u();
      for (int d = 0; d < 3; d++) {
        q.x[d] = float(x[d]);
        q.u[d] = float(u[d]);
      }
      q.kind = prt.kind();
      q.qni_wni = float(prt.qni_wni());
      prts.push_back(q);
    }
  }
  mprts.set(prts, std::vector<uint32_t>(n_by_patch.begin(), n_by_patch.end()));
}

} // namespace detail

inline const MparticlesB200Psc::Convert MparticlesB200Psc::convert_to_ = {
  {std::type_index(typeid(MparticlesSingle)), detail::copy_to<MparticlesSingle>},
  {std::type_index(typeid(MparticlesDouble)), detail::copy_to<MparticlesDouble>},
};
inline const MparticlesB200Psc::Convert MparticlesB200Psc::convert_from_ = {
  {std::type_index(typeid(MparticlesSingle)), detail::copy_from<MparticlesSingle>},
  {std::type_index(typeid(MparticlesDouble)), detail::copy_from<MparticlesDouble>},
};

} // namespace psc_b200

// ----------------------------------------------------------------------
// MfieldsStateB200Psc (needs the real gtensor: see the header comment)

#ifdef PSC_B200_ADAPT_FIELDS

#include "fields3d.hxx"
#include "psc_fields_single.h"
#include "psc_fields_c.h"

namespace psc_b200
{

struct MfieldsStateB200Psc
  : MfieldsStateBase
  , MfieldsStateB200<Grid_t>
{
  using Impl = MfieldsStateB200<Grid_t>;
  using real_t = float;
  using space = gt::space::host; // what storage() hands out lives on the host
  using Storage = gt::gtensor<float, 5>;

  explicit MfieldsStateB200Psc(const Grid_t& grid)
    : MfieldsStateBase(grid, PSC_B200_NR_FIELDS, grid.ibn), Impl(grid)
  {}

  const Grid_t& grid() const { return Impl::grid(); }
  int n_patches() const { return Impl::n_patches(); }
  int n_comps() const { return Impl::n_comps(); }
  Int3 ib() const { return {-Impl::ibn()[0], -Impl::ibn()[1], -Impl::ibn()[2]}; }
  Int3 ibn() const { return {Impl::ibn()[0], Impl::ibn()[1], Impl::ibn()[2]}; }

  void reset(const Grid_t& grid) override
  {
    MfieldsStateBase::reset(grid);
    Impl::reset(grid);
  }

  // A host snapshot in PSC's layout (x fastest, then y, z, component, patch:
  // fields3d.hxx:29-32,284-291), refreshed from the device on every call.  It is its own
  // gt::host_mirror, so `auto&& h = gt::host_mirror(mflds.storage()); gt::copy(mflds.storage(), h);`
  // (DiagEnergiesField.h:24-26, the output writers) reads the current device fields.
  // Writers go through hostMirror() / copy() below, as with MfieldsStateCuda.
  Storage& storage()
  {
    const auto im = Impl::im();
    if (h_.size() == 0) {
      h_ = gt::empty<float>({im[0], im[1], im[2], n_comps(), n_patches()});
    }
    PSC_B200_CHECK(psc_b200_mflds_download(ctx(), id(), 0, n_comps(), h_.data()));
    return h_;
  }
  Storage& gt() { return storage(); }
  // after the deck wrote into storage() (setupFields): push it to the device
  void upload_storage() { PSC_B200_CHECK(psc_b200_mflds_upload(ctx(), id(), 0, n_comps(), h_.data())); }

  static const Convert convert_to_, convert_from_;
  const Convert& convert_to() override { return convert_to_; }
  const Convert& convert_from() override { return convert_from_; }

private:
  Storage h_;
};

// setup_fields_cuda.hxx / psc_fields_cuda.h:78-113: hostMirror(mflds) + copy(from, to)
inline MfieldsStateSingle hostMirror(MfieldsStateB200Psc& mflds)
{
  return MfieldsStateSingle{mflds.grid()};
}
inline void copy(MfieldsStateB200Psc& mflds, MfieldsStateSingle& hmflds)
{
  gt::copy(mflds.storage(), hmflds.storage());
}
inline void copy(MfieldsStateSingle& hmflds, MfieldsStateB200Psc& mflds)
{
  auto& h = mflds.storage();
  gt::copy(hmflds.storage(), h);
  mflds.upload_storage();
}

namespace detail
{

template <typename MF>
void flds_copy_to(MfieldsStateBase& from_base, MfieldsStateBase& to_base, int mb, int me)
{
  auto& from = dynamic_cast<MfieldsStateB200Psc&>(from_base);
  auto& to = dynamic_cast<MF&>(to_base);
  auto& h = from.storage();
  to.storage().view(_all, _all, _all, _s(mb, me), _all) = h.view(_all, _all, _all, _s(mb, me), _all);
}

template <typename MF>
void flds_copy_from(MfieldsStateBase& to_base, MfieldsStateBase& from_base, int mb, int me)
{
  auto& to = dynamic_cast<MfieldsStateB200Psc&>(to_base);
  auto& from = dynamic_cast<MF&>(from_base);
  auto& h = to.storage();
  h.view(_all, _all, _all, _s(mb, me), _all) = from.storage().view(_all, _all, _all, _s(mb, me), _all);
  to.upload_storage();
}

} // namespace detail

inline const MfieldsStateB200Psc::Convert MfieldsStateB200Psc::convert_to_ = {
  {std::type_index(typeid(MfieldsStateSingle)), detail::flds_copy_to<MfieldsStateSingle>},
};
inline const MfieldsStateB200Psc::Convert MfieldsStateB200Psc::convert_from_ = {
  {std::type_index(typeid(MfieldsStateSingle)), detail::flds_copy_from<MfieldsStateSingle>},
};

} // namespace psc_b200

#endif // PSC_B200_ADAPT_FIELDS
