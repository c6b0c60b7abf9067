// psc_b200 -- the PscConfig plugin surface of psc-code/psc over the psc_b200 C ABI.
//
// PSC's plugin API is compile-time duck typing: `template <typename PscConfig> struct Psc`
// pulls its operator types out of PscConfig (src/include/psc.hxx:100-119) and a case deck
// picks a config next to PscConfig1vbecCuda (src/psc_config.hxx:99-146,170-180).  This
// header provides that set of types for the B200 backend:
//
//     template <typename Dim> using PscConfig1vbecB200 = psc_b200::PscConfig<Dim, Grid_t>;
//
//   Mparticles     MparticlesB200       ~ MparticlesSimple  (particles_simple.hxx:123-247)
//   MfieldsState   MfieldsStateB200     ~ MfieldsStateFromMfields (fields3d.hxx:415-464)
//   Mfields        MfieldsB200          ~ Mfields<float>    (fields3d.hxx:321-382)
//   PushParticles  PushParticlesB200    ~ PushParticlesVb   (push_particles_1vb.hxx:27-84)
//   Sort           SortB200             ~ SortCountsort2    (psc_sort_impl.hxx:65-124)
//   BndParticles   BndParticlesB200     ~ BndParticlesCommon (bnd_particles_impl.hxx:234-247)
//   Bnd            BndB200              ~ Bnd_              (psc_bnd_impl.hxx:105-158)
//   BndFields      BndFieldsB200        ~ BndFields_        (psc_bnd_fields_impl.hxx:27-188)
//   PushFields     PushFieldsB200       ~ PushFields        (psc_push_fields_impl.hxx:134-178)
//   Marder         MarderB200           ~ MarderCommon      (marder_impl.hxx:197-264)
//   Checks         ChecksB200           ~ Checks_           (checks_impl.hxx:33-215)
//   Balance        BalanceB200          ~ Balance_          (psc_balance_impl.hxx:770-1026)
//   Collision      CollisionB200        ~ Collision_ / CollisionHost (psc_collision_impl.hxx:20-274) on the
//                                        device; CollisionViaHostB200 = PSC's CollisionCudaHost round trip
//   moments        Moment_n_1st_B200 ...~ ItemMoment<moment_*> (fields_item_moments_1st.hxx:9-37)
//
// Every method is one call into libpsc_b200.so (include/psc_b200.h); nothing is computed
// on the host and there is no CPU fallback.  Error behaviour is PSC's: a failed call
// prints the library's error text and abort()s (libpsc/bits.hxx:35-40,
// cuda/cuda_bits.h:32-40) -- no exceptions, no error codes.
//
// The types are templates over the grid type so that this header compiles both inside a
// PSC tree (GridT = Grid_t, src/include/grid.hxx:68-160) and stand-alone in this
// repository's tests (tests/cxx/mini_grid.hxx, a struct with the same member names).
// What is read from GridT:
//   domain.gdims/np/length/corner/dx, ldims, ibn, dt, norm.fnqs/eta, kinds[k].q/.m,
//   bc.fld_lo/fld_hi/prt_lo/prt_hi, patches[p].xb, n_patches()
#pragma once

#include "../psc_b200.h"

#include <algorithm>
#include <array>
#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <random>
#include <string>
#include <vector>

namespace psc_b200
{

// LOG_ERROR + abort, like cudaCheck (cuda/cuda_bits.h:32-40)
inline void check(int rc, const char* what)
{
  if (rc != 0) {
    std::fprintf(stderr, "psc_b200: %s failed: %s\n", what, psc_b200_last_error());
    std::abort();
  }
}
#define PSC_B200_CHECK(call) ::psc_b200::check((call), #call)

// ParticleSimple<float> (particle_simple.hxx:10-42) as stored on the wire
struct Particle
{
  using real_t = float;
  float x[3];
  float u[3];
  int kind;
  float qni_wni;
};
static_assert(sizeof(Particle) == 32, "AoS record must be 32 bytes");

// ----------------------------------------------------------------------
// Context: one psc_b200_ctx per (rank, Grid_t).  PSC constructs Mparticles(grid) and
// MfieldsState(grid) separately (psc_bubble_yz.cxx:290-291); both attach to the context
// of their grid, which is created on first use and lives as long as one of them does.

template <typename GridT>
class Context
{
public:
  static std::shared_ptr<Context> get(const GridT& grid, int deposit = PSC_B200_DEPOSIT_DEFAULT)
  {
    auto& reg = registry();
    auto it = reg.find(&grid);
    if (it != reg.end()) {
      if (auto sp = it->second.lock()) {
        return sp;
      }
    }
    auto sp = std::shared_ptr<Context>(new Context(grid, deposit));
    reg[&grid] = sp;
    return sp;
  }

  ~Context()
  {
    psc_b200_destroy(ctx_);
    registry().erase(grid_);
  }

  psc_b200_ctx* ctx() const { return ctx_; }
  const GridT& grid() const { return *grid_; }

  static bool known(const GridT& grid)
  {
    auto it = registry().find(&grid);
    return it != registry().end() && !it->second.expired();
  }

  // Balance replaced the host Grid_t (psc_balance_impl.hxx:893-1016) after the patches moved
  // INSIDE this context: the context follows the new grid object, it is never re-created
  // (that would leave the rebalanced particles and fields behind in the old one).  Every
  // container that shares the context sees the new grid from here on.
  void rebind(const GridT& new_grid)
  {
    assert(psc_b200_n_patches(ctx_) == new_grid.n_patches());
    auto& reg = registry();
    auto it = reg.find(grid_);
    std::weak_ptr<Context> self;
    if (it != reg.end()) {
      self = it->second;
      reg.erase(it);
    }
    grid_ = &new_grid;
    reg[grid_] = self;
  }

  // multi-GPU: rank / n_ranks / patch counts of the decomposition, set before the first
  // container is constructed (defaults: one rank owning every patch)
  struct Decomposition
  {
    int rank = 0, n_ranks = 1, device = -1;
    std::vector<int> n_patches_by_rank; // empty = uniform
    unsigned long long max_n_prts = 0;
  };
  static Decomposition& decomposition()
  {
    static Decomposition d;
    return d;
  }

private:
  Context(const GridT& grid, int deposit) : grid_(&grid)
  {
    psc_b200_grid_desc d;
    std::memset(&d, 0, sizeof(d));
    for (int i = 0; i < 3; i++) {
      d.gdims[i] = grid.domain.gdims[i];
      d.np[i] = grid.domain.np[i];
      d.length[i] = grid.domain.length[i];
      d.corner[i] = grid.domain.corner[i];
      d.bc_fld_lo[i] = grid.bc.fld_lo[i];
      d.bc_fld_hi[i] = grid.bc.fld_hi[i];
      d.bc_prt_lo[i] = grid.bc.prt_lo[i];
      d.bc_prt_hi[i] = grid.bc.prt_hi[i];
    }
    d.dt = grid.dt;
    d.fnqs = grid.norm.fnqs;
    d.eta = grid.norm.eta;
    d.n_kinds = (int)grid.kinds.size();
    assert(d.n_kinds <= PSC_B200_MAX_KINDS);
    for (int k = 0; k < d.n_kinds; k++) {
      d.q[k] = grid.kinds[k].q;
      d.m[k] = grid.kinds[k].m;
    }
    d.deposit = deposit;
    const auto& dec = decomposition();
    d.rank = dec.rank;
    d.n_ranks = dec.n_ranks;
    d.device = dec.device;
    d.n_patches_by_rank = dec.n_patches_by_rank.empty() ? nullptr : dec.n_patches_by_rank.data();
    d.max_n_prts = dec.max_n_prts;
    PSC_B200_CHECK(psc_b200_create(&d, &ctx_));
    assert(psc_b200_n_patches(ctx_) == grid.n_patches());
  }

  static std::map<const GridT*, std::weak_ptr<Context>>& registry()
  {
    static std::map<const GridT*, std::weak_ptr<Context>> r;
    return r;
  }

  const GridT* grid_;
  psc_b200_ctx* ctx_ = nullptr;
};

// ----------------------------------------------------------------------
// InjectorB200: injector()[p](psc::particle::Inject) -- buffered, flushed when the
// injector goes out of scope (cuda/injector_buffered.hxx:62-66); patches must be
// visited in ascending order like there.  The conversion to the stored record is
// InjectorSimple::Patch::operator() (injector_simple.hxx:20-37):
//   x = Real3(new.x) - Real3(patch.xb)  (both narrowed to float first), qni_wni = w*q(kind)

template <typename Mparticles>
class InjectorB200
{
public:
  class Patch
  {
  public:
    Patch(InjectorB200& inj, int p) : inj_(inj), p_(p) {}

    template <typename Inject>
    void operator()(const Inject& new_prt)
    {
      const auto& grid = inj_.mprts_.grid();
      const auto& patch = grid.patches[p_];
      Particle prt;
      for (int d = 0; d < 3; d++) {
        prt.x[d] = float(new_prt.x[d]) - float(patch.xb[d]);
        prt.u[d] = float(new_prt.u[d]);
      }
      prt.kind = new_prt.kind;
      prt.qni_wni = float(new_prt.w * grid.kinds[new_prt.kind].q);
      inj_.push_back(p_, prt);
    }

    // already-converted record (MparticlesSimple::push_back)
    void raw(const Particle& prt) { inj_.push_back(p_, prt); }

  private:
    InjectorB200& inj_;
    int p_;
  };

  explicit InjectorB200(Mparticles& mprts) : mprts_(mprts), n_by_patch_(mprts.n_patches(), 0) {}
  InjectorB200(const InjectorB200&) = delete;
  InjectorB200(InjectorB200&& o)
    : mprts_(o.mprts_), buf_(std::move(o.buf_)), n_by_patch_(std::move(o.n_by_patch_)), last_(o.last_)
  {
    o.buf_.clear();
    o.n_by_patch_.assign(mprts_.n_patches(), 0);
  }

  ~InjectorB200() { flush(); }

  void reserve(int n_prts_total) { buf_.reserve(n_prts_total); }

  Patch operator[](int p) { return Patch(*this, p); }

  void flush()
  {
    if (buf_.empty()) {
      return;
    }
    PSC_B200_CHECK(psc_b200_mprts_inject(mprts_.ctx(), buf_.data(), n_by_patch_.data()));
    buf_.clear();
    n_by_patch_.assign(mprts_.n_patches(), 0);
    last_ = 0;
  }

private:
  void push_back(int p, const Particle& prt)
  {
    assert(p >= last_ && "InjectorB200: patches must be injected in ascending order");
    last_ = p;
    buf_.push_back(prt);
    n_by_patch_[p]++;
  }

  Mparticles& mprts_;
  std::vector<Particle> buf_;
  std::vector<uint32_t> n_by_patch_;
  int last_ = 0;
};

// ----------------------------------------------------------------------
// ConstAccessorB200: accessor()[p] -> range of particle proxies with PSC's accessor
// vocabulary (const_accessor_simple.hxx:47-81).  A device->host copy of the whole store
// (the reference's CUDA backend does the same, cuda_mparticles.cu get_particles).

template <typename Mparticles>
class ConstAccessorB200
{
public:
  using GridT = typename Mparticles::Grid;

  struct Proxy
  {
    const Particle& prt;
    const GridT& grid;
    int p;
    std::array<float, 3> x() const { return {prt.x[0], prt.x[1], prt.x[2]}; }
    std::array<float, 3> u() const { return {prt.u[0], prt.u[1], prt.u[2]}; }
    float qni_wni() const { return prt.qni_wni; }
    int kind() const { return prt.kind; }
    float q() const { return float(grid.kinds[prt.kind].q); }
    float m() const { return float(grid.kinds[prt.kind].m); }
    float w() const { return prt.qni_wni / q(); }
    // global position (const_accessor_simple.hxx:69-75)
    std::array<double, 3> position() const
    {
      const auto& patch = grid.patches[p];
      return {patch.xb[0] + prt.x[0], patch.xb[1] + prt.x[1], patch.xb[2] + prt.x[2]};
    }
  };

  struct Patch
  {
    const ConstAccessorB200& acc;
    int p;
    struct iterator
    {
      const Patch& patch;
      uint32_t n;
      Proxy operator*() const { return {patch.acc.data_[n], patch.acc.mprts_.grid(), patch.p}; }
      iterator& operator++()
      {
        ++n;
        return *this;
      }
      bool operator!=(const iterator& o) const { return n != o.n; }
    };
    iterator begin() const { return {*this, acc.off_[p]}; }
    iterator end() const { return {*this, acc.off_[p + 1]}; }
    uint32_t size() const { return acc.off_[p + 1] - acc.off_[p]; }
    Proxy operator[](uint32_t n) const { return {acc.data_[acc.off_[p] + n], acc.mprts_.grid(), p}; }
  };

  explicit ConstAccessorB200(Mparticles& mprts)
    : mprts_(mprts), data_(mprts.size()), off_(mprts.n_patches() + 1)
  {
    PSC_B200_CHECK(psc_b200_mprts_get(mprts.ctx(), data_.data(), off_.data()));
  }

  Patch operator[](int p) const { return {*this, p}; }
  const std::vector<Particle>& data() const { return data_; }
  const std::vector<uint32_t>& offsets() const { return off_; }

private:
  Mparticles& mprts_;
  std::vector<Particle> data_;
  std::vector<uint32_t> off_;
};

// ----------------------------------------------------------------------
// MparticlesB200

template <typename GridT>
class MparticlesB200
{
public:
  using Grid = GridT;
  using real_t = float;
  using Particle = psc_b200::Particle;
  using is_cuda = std::true_type; // device-resident: deck code takes its is_cuda branches

  explicit MparticlesB200(const GridT& grid) : cx_(Context<GridT>::get(grid)) {}

  const GridT& grid() const { return cx_->grid(); }
  psc_b200_ctx* ctx() const { return cx_->ctx(); }
  int n_patches() const { return psc_b200_n_patches(ctx()); }

  // MparticlesBase::size / sizeByPatch (particles.hxx:19-33)
  int size() const
  {
    uint64_t n;
    PSC_B200_CHECK(psc_b200_mprts_size(ctx(), &n));
    return (int)n;
  }
  std::vector<unsigned int> sizeByPatch() const
  {
    std::vector<unsigned int> n(n_patches());
    PSC_B200_CHECK(psc_b200_mprts_size_by_patch(ctx(), n.data()));
    return n;
  }

  // MparticlesBase::reset (Balance hands over the new grid, psc_balance_impl.hxx:893-909)
  // (a grid the registry does not know yet is the regridded one of THIS context)
  void reset(const GridT& grid)
  {
    if (!Context<GridT>::known(grid) && psc_b200_n_patches(ctx()) == grid.n_patches()) {
      cx_->rebind(grid);
    }
    cx_ = Context<GridT>::get(grid);
  }
  Context<GridT>& context() { return *cx_; }

  InjectorB200<MparticlesB200> injector() { return InjectorB200<MparticlesB200>(*this); }
  ConstAccessorB200<MparticlesB200> accessor() { return ConstAccessorB200<MparticlesB200>(*this); }

  // bulk host interface (get_as<MparticlesSingle> / put_as, particles.hxx:35-54)
  void set(const std::vector<Particle>& prts, const std::vector<uint32_t>& n_by_patch)
  {
    assert((int)n_by_patch.size() == n_patches());
    PSC_B200_CHECK(psc_b200_mprts_set(ctx(), prts.data(), n_by_patch.data()));
  }
  void get(std::vector<Particle>& prts, std::vector<uint32_t>& off) const
  {
    prts.resize(size());
    off.resize(n_patches() + 1);
    PSC_B200_CHECK(psc_b200_mprts_get(ctx(), prts.data(), off.data()));
  }

private:
  std::shared_ptr<Context<GridT>> cx_;
};

// ----------------------------------------------------------------------
// MfieldsB200 / MfieldsStateB200: PSC's layout float [p][m][iz][iy][ix]
// (fields3d.hxx:29-32,284-291); ib = -ibn, im = ldims + 2 ibn

template <typename GridT>
class MfieldsB200
{
public:
  using Grid = GridT;
  using real_t = float;

  // scratch container with n_comps components, ghosts = grid.ibn
  MfieldsB200(const GridT& grid, int n_comps) : cx_(Context<GridT>::get(grid)), n_comps_(n_comps)
  {
    PSC_B200_CHECK(psc_b200_mflds_create(ctx(), n_comps, &id_));
    init_dims();
  }

  const GridT& grid() const { return cx_->grid(); }
  psc_b200_ctx* ctx() const { return cx_->ctx(); }
  // MfieldsStateBase / MfieldsBase::reset (fields3d.hxx:159-245): after a rebalance the
  // container stays with its context (which follows the new grid, Context::rebind)
  void reset(const GridT& grid)
  {
    if (!Context<GridT>::known(grid) && psc_b200_n_patches(ctx()) == grid.n_patches()) {
      cx_->rebind(grid);
    }
    cx_ = Context<GridT>::get(grid);
  }
  int id() const { return id_; }
  int n_comps() const { return n_comps_; }
  int n_patches() const { return psc_b200_n_patches(ctx()); }
  std::array<int, 3> ibn() const { return ibn_; }
  std::array<int, 3> ib() const { return {-ibn_[0], -ibn_[1], -ibn_[2]}; }
  std::array<int, 3> im() const { return im_; }
  size_t patch_len() const { return (size_t)im_[0] * im_[1] * im_[2]; }

  void zero(int mb, int me) { PSC_B200_CHECK(psc_b200_mflds_zero(ctx(), id_, mb, me)); }
  void zero() { zero(0, n_comps_); }

  // hostMirror + copy of setup_fields_cuda.hxx: components [mb, me) of every patch
  std::vector<float> download(int mb, int me) const
  {
    std::vector<float> h((size_t)n_patches() * (me - mb) * patch_len());
    PSC_B200_CHECK(psc_b200_mflds_download(ctx(), id_, mb, me, h.data()));
    return h;
  }
  void upload(int mb, int me, const std::vector<float>& h)
  {
    assert(h.size() == (size_t)n_patches() * (me - mb) * patch_len());
    PSC_B200_CHECK(psc_b200_mflds_upload(ctx(), id_, mb, me, h.data()));
  }
  // what OutputFieldsItem does to an item (output_fields.hxx:179,203,221), on the device:
  // this[mb + m] += other[other_mb + m]
  void add(const MfieldsB200& other, int mb = 0, int other_mb = 0, int n_comps = -1)
  {
    if (n_comps < 0) {
      n_comps = std::min(n_comps_ - mb, other.n_comps_ - other_mb);
    }
    PSC_B200_CHECK(psc_b200_mflds_add(ctx(), id_, mb, other.id_, other_mb, n_comps));
  }
  // this[m] = float(a * double(this[m]))
  void scale(double a) { PSC_B200_CHECK(psc_b200_mflds_scale(ctx(), id_, 0, n_comps_, a)); }
  // psc::mflds::interior on the host: [p][m][k][j][i] over the patches' own cells
  std::vector<float> download_interior(int mb, int me) const
  {
    int ld[3], ibn[3];
    PSC_B200_CHECK(psc_b200_get_ldims(ctx(), ld, ibn));
    std::vector<float> h((size_t)n_patches() * (me - mb) * ld[0] * ld[1] * ld[2]);
    PSC_B200_CHECK(psc_b200_mflds_download_interior(ctx(), id_, mb, me, h.data()));
    return h;
  }
  std::vector<float> download_interior() const { return download_interior(0, n_comps_); }

  // offset of (m, i, j, k) of patch p inside a download(mb, me) buffer
  size_t index(int p, int m_rel, int n_m, int i, int j, int k) const
  {
    return ((((size_t)p * n_m + m_rel) * im_[2] + (k + ibn_[2])) * im_[1] + (j + ibn_[1])) * im_[0] +
           (i + ibn_[0]);
  }

protected:
  struct state_tag
  {};
  MfieldsB200(const GridT& grid, state_tag)
    : cx_(Context<GridT>::get(grid)), n_comps_(PSC_B200_NR_FIELDS), id_(0)
  {
    init_dims();
  }
  void init_dims()
  {
    int ld[3], ibn[3];
    PSC_B200_CHECK(psc_b200_get_ldims(ctx(), ld, ibn));
    for (int d = 0; d < 3; d++) {
      ibn_[d] = ibn[d];
      im_[d] = ld[d] + 2 * ibn[d];
    }
  }

  std::shared_ptr<Context<GridT>> cx_;
  int n_comps_;
  int id_ = -1;
  std::array<int, 3> ibn_, im_;
};

template <typename GridT>
class MfieldsStateB200 : public MfieldsB200<GridT>
{
public:
  using Base = MfieldsB200<GridT>;
  explicit MfieldsStateB200(const GridT& grid) : Base(grid, typename Base::state_tag{}) {}
  void reset(const GridT& grid) { *this = MfieldsStateB200(grid); }

  // setupFields (setup_fields.hxx:19-45): init(m, x[3]) evaluated at each component's
  // Yee position (centering.hxx; bits/discretization.txt:4-12), ghosts included.
  template <typename F>
  void setup(F&& init)
  {
    const auto& g = this->grid();
    // component m: staggered (+1/2) in dimension d?  E_d: own dim; H_d: the two others
    static const int stag[6][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {0, 1, 1}, {1, 0, 1}, {1, 1, 0}};
    auto ibn = this->ibn();
    auto im = this->im();
    std::vector<float> h((size_t)this->n_patches() * 6 * this->patch_len());
    for (int p = 0; p < this->n_patches(); p++) {
      const auto& patch = g.patches[p];
      for (int m = 0; m < 6; m++) {
        for (int k = -ibn[2]; k < im[2] - ibn[2]; k++) {
          for (int j = -ibn[1]; j < im[1] - ibn[1]; j++) {
            for (int i = -ibn[0]; i < im[0] - ibn[0]; i++) {
              int idx[3] = {i, j, k};
              double x[3];
              for (int d = 0; d < 3; d++) {
                x[d] = patch.xb[d] + (idx[d] + .5 * stag[m][d]) * g.domain.dx[d];
              }
              h[this->index(p, m, 6, i, j, k)] = float(init(PSC_B200_EX + m, x));
            }
          }
        }
      }
    }
    this->upload(PSC_B200_EX, PSC_B200_EX + 6, h);
  }
};

// ----------------------------------------------------------------------
// operators

template <typename GridT>
struct PushParticlesB200
{
  using Mparticles = MparticlesB200<GridT>;
  using MfieldsState = MfieldsStateB200<GridT>;
  // push_particles_1vb.hxx:27-84
  void push_mprts(Mparticles& mprts, MfieldsState& mflds)
  {
    assert(mprts.ctx() == mflds.ctx());
    PSC_B200_CHECK(psc_b200_push_mprts(mprts.ctx()));
  }
};

template <typename GridT>
struct SortB200
{
  // psc_sort_impl.hxx:65-124
  void operator()(MparticlesB200<GridT>& mprts) { PSC_B200_CHECK(psc_b200_sort(mprts.ctx())); }
};

template <typename GridT>
struct BndParticlesB200
{
  explicit BndParticlesB200(const GridT&) {}
  // bnd_particles_impl.hxx:234-247; rebuilds nothing here: the neighbour tables live in
  // the context and follow the grid (balance bumps them, psc_balance_impl.hxx:1018)
  void operator()(MparticlesB200<GridT>& mprts) { PSC_B200_CHECK(psc_b200_bnd_particles(mprts.ctx())); }
};

template <typename GridT>
struct BndB200
{
  // psc_bnd_impl.hxx:105-158
  void add_ghosts(MfieldsB200<GridT>& mflds, int mb, int me)
  {
    PSC_B200_CHECK(psc_b200_bnd_add_ghosts(mflds.ctx(), mflds.id(), mb, me));
  }
  void fill_ghosts(MfieldsB200<GridT>& mflds, int mb, int me)
  {
    PSC_B200_CHECK(psc_b200_bnd_fill_ghosts(mflds.ctx(), mflds.id(), mb, me));
  }
};

template <typename GridT>
struct BndFieldsB200
{
  using MfieldsState = MfieldsStateB200<GridT>;
  // psc_bnd_fields_impl.hxx:27-188
  void fill_ghosts_E(MfieldsState& mflds) { PSC_B200_CHECK(psc_b200_bndf_fill_ghosts_E(mflds.ctx())); }
  void fill_ghosts_H(MfieldsState& mflds) { PSC_B200_CHECK(psc_b200_bndf_fill_ghosts_H(mflds.ctx())); }
  void add_ghosts_J(MfieldsState& mflds) { PSC_B200_CHECK(psc_b200_bndf_add_ghosts_J(mflds.ctx())); }
};

template <typename GridT>
struct PushFieldsB200
{
  using MfieldsState = MfieldsStateB200<GridT>;
  // psc_push_fields_impl.hxx:134-178; the Dim tag is what the context was created for
  template <typename Dim>
  void push_E(MfieldsState& mflds, double dt_fac, Dim)
  {
    PSC_B200_CHECK(psc_b200_push_E(mflds.ctx(), dt_fac));
  }
  template <typename Dim>
  void push_H(MfieldsState& mflds, double dt_fac, Dim)
  {
    PSC_B200_CHECK(psc_b200_push_H(mflds.ctx(), dt_fac));
  }
  void push_E(MfieldsState& mflds, double dt_fac) { PSC_B200_CHECK(psc_b200_push_E(mflds.ctx(), dt_fac)); }
  void push_H(MfieldsState& mflds, double dt_fac) { PSC_B200_CHECK(psc_b200_push_H(mflds.ctx(), dt_fac)); }
};

template <typename GridT>
struct MarderB200
{
  using MfieldsState = MfieldsStateB200<GridT>;
  using Mparticles = MparticlesB200<GridT>;
  // marder_impl.hxx:150-264
  MarderB200(const GridT&, double diffusion, int loop, bool /*dump*/) : diffusion_(diffusion), loop_(loop)
  {}
  void correct_gauss(MfieldsState& mflds, Mparticles&)
  {
    PSC_B200_CHECK(psc_b200_marder(mflds.ctx(), diffusion_, loop_));
  }
  void operator()(MfieldsState& mflds, Mparticles& mprts) { correct_gauss(mflds, mprts); }

  double diffusion_;
  int loop_;
};

// ItemMoment (include/fields_item.hxx:97-134) over the device-side 1st-order moments of
// libpsc/psc_output_fields/fields_item_moments_1st.hxx:9-37: same call shape as the
// reference's items -- construct from the grid, call on the particles, get the result
// container (ghost add and reflecting folds done) -- so OutputFields / the flatfoil
// injector can take them without a device -> host particle copy.
template <typename GridT, int WHICH>
class MomentB200
{
public:
  using Mparticles = MparticlesB200<GridT>;
  using Mfields = MfieldsB200<GridT>;

  static std::string name()
  {
    static const char* names[] = {"n_1st_cc", "v_1st_cc", "p_1st_cc", "T_1st_cc", "all_1st_cc", "rho_1st_nc"};
    return names[WHICH];
  }
  explicit MomentB200(const GridT& grid)
    : mres_(grid, psc_b200_moment_n_comps(Context<GridT>::get(grid)->ctx(), WHICH))
  {}
  int n_comps() const { return mres_.n_comps(); }
  // addKindSuffix (fields_item.hxx:22-32) over the moment's stems (psc/moment.hxx:130-133,
  // 158-161, 185-188, 217-220, 246-249, 281-286): kinds outermost
  std::vector<std::string> comp_names() const
  {
    static const std::vector<std::string> stems[] = {
      {"n"},
      {"vx", "vy", "vz"},
      {"px", "py", "pz"},
      {"Txx", "Tyy", "Tzz", "Txy", "Txz", "Tyz"},
      {"rho", "jx", "jy", "jz", "px", "py", "pz", "txx", "tyy", "tzz", "txy", "tyz", "tzx"}};
    if (WHICH == PSC_B200_MOMENT_RHO_NC) {
      return {"rho"};
    }
    std::vector<std::string> result;
    const auto& kinds = mres_.grid().kinds;
    for (size_t k = 0; k < kinds.size(); k++) {
      for (const auto& stem : stems[WHICH]) {
        result.emplace_back(stem + "_" + kinds[k].name);
      }
    }
    return result;
  }
  Mfields& operator()(Mparticles& mprts)
  {
    PSC_B200_CHECK(psc_b200_moment_1st(mprts.ctx(), mres_.id(), WHICH));
    return mres_;
  }

private:
  Mfields mres_;
};

template <typename GridT>
using Moment_n_1st_B200 = MomentB200<GridT, PSC_B200_MOMENT_N>;
template <typename GridT>
using Moment_v_1st_B200 = MomentB200<GridT, PSC_B200_MOMENT_V>;
template <typename GridT>
using Moment_p_1st_B200 = MomentB200<GridT, PSC_B200_MOMENT_P>;
template <typename GridT>
using Moment_T_1st_B200 = MomentB200<GridT, PSC_B200_MOMENT_T>;
template <typename GridT>
using Moments_1st_B200 = MomentB200<GridT, PSC_B200_MOMENT_ALL>;
template <typename GridT>
using Moment_rho_1st_nc_B200 = MomentB200<GridT, PSC_B200_MOMENT_RHO_NC>;

// ----------------------------------------------------------------------
// OutputFields / OutputMoments (src/include/output_fields.hxx).  The parameter structs are
// the reference's, member for member; OutputFieldsItemB200 is its OutputFieldsItem with the
// item left on the device: evaluated there (Item_jeh = the state fields themselves, no copy;
// the moments by psc_b200_moment_1st), accumulated for the time average there, and only what
// a writer is handed crosses to the host -- the interior of the item (pfd) or of the mean
// (tfd).  Writer: `explicit operator bool`, `open(pfx, dir)`, `write_step(grid, rn, rx, data,
// name, comp_names)` as WriterMRC / WriterADIOS2 have them (writer_mrc.hxx:8-123), `data`
// being a HostItemB200 (interior, [p][m][k][j][i]) instead of a gtensor expression.

struct BaseOutputFieldItemParamsB200 // output_fields.hxx:63-78
{
  int out_interval = 0; // difference between output timesteps (0 = disable)
  std::string data_dir = ".";
  std::array<int, 3> rn = {};
  std::array<int, 3> rx = {10000000, 10000000, 10000000};

  bool enabled() const { return out_interval > 0; }
  bool do_out(int timestep) const { return enabled() && timestep % out_interval == 0; }
};

struct OutputPfieldItemParamsB200 : BaseOutputFieldItemParamsB200
{};

struct OutputTfieldItemParamsB200 : BaseOutputFieldItemParamsB200 // :83-101
{
  int average_length = 1000000; // max range of timesteps over which to average
  int sample_interval = 1;      // difference between timesteps used for average

  bool do_accum(int timestep) const
  {
    if (!enabled()) {
      return false;
    }
    int n_intervals_elapsed = (timestep - 1) / out_interval;
    int next_out = out_interval * (n_intervals_elapsed + 1); // could be this timestep
    bool in_averaging_range = next_out - timestep < average_length;
    bool on_averaging_step = (next_out - timestep) % sample_interval == 0;
    return in_averaging_range && on_averaging_step;
  }
};

struct OutputFieldsItemParamsB200 // :139-143
{
  OutputPfieldItemParamsB200 pfield;
  OutputTfieldItemParamsB200 tfield;
};

struct HostItemB200
{
  std::vector<float> data; // [p][m][k][j][i], the patches' own cells
  int n_patches, n_comps;
  std::array<int, 3> ldims;
  int timestep; // the step the item belongs to (inside PSC: grid.timestep())
};

// keeps what it was given (tests; a deck inside PSC wraps WriterMRC / WriterADIOS2 instead)
struct WriterMemoryB200
{
  struct Step
  {
    int timestep;
    std::string name;
    std::vector<std::string> comp_names;
    HostItemB200 item;
  };
  explicit operator bool() const { return !pfx.empty(); }
  void open(const std::string& pfx_, const std::string& dir_ = ".")
  {
    assert(pfx.empty());
    pfx = pfx_, dir = dir_;
  }
  template <typename GridT>
  void write_step(const GridT& grid, const std::array<int, 3>&, const std::array<int, 3>&, HostItemB200&& item,
                  const std::string& name, const std::vector<std::string>& comp_names)
  {
    const int timestep = item.timestep;
    steps.push_back(Step{timestep, name, comp_names, std::move(item)});
    (void)grid;
  }
  std::string pfx, dir;
  std::vector<Step> steps;
};

template <typename GridT>
struct Item_jeh_B200 // fields_item_fields.hxx:14-31
{
  static std::string name() { return "jeh"; }
  static int n_comps() { return PSC_B200_NR_FIELDS; }
  static std::vector<std::string> comp_names()
  {
    return {"jx_ec", "jy_ec", "jz_ec", "ex_ec", "ey_ec", "ez_ec", "hx_fc", "hy_fc", "hz_fc"};
  }
  MfieldsB200<GridT>& operator()(MfieldsStateB200<GridT>& mflds) const { return mflds; }
};

template <typename GridT>
struct GetItemJehB200 // :106-117
{
  static std::string suffix() { return ""; }
  explicit GetItemJehB200(const GridT&) {}
  MfieldsB200<GridT>& get_item(MparticlesB200<GridT>&, MfieldsStateB200<GridT>& mflds) { return item_(mflds); }
  std::string name() const { return item_.name(); }
  std::vector<std::string> comp_names() const { return item_.comp_names(); }
  Item_jeh_B200<GridT> item_;
};

template <typename GridT>
struct GetItemMomentsB200 // :119-131 (Item_Moments = Moments_1st, all 13 moments per kind)
{
  static std::string suffix() { return "_moments"; }
  explicit GetItemMomentsB200(const GridT& grid) : item_(grid) {}
  MfieldsB200<GridT>& get_item(MparticlesB200<GridT>& mprts, MfieldsStateB200<GridT>&) { return item_(mprts); }
  std::string name() const { return item_.name(); }
  std::vector<std::string> comp_names() const { return item_.comp_names(); }
  Moments_1st_B200<GridT> item_;
};

template <typename GridT, typename GetItem, typename Writer = WriterMemoryB200>
class OutputFieldsItemB200 : public OutputFieldsItemParamsB200 // :150-236
{
public:
  using Mparticles = MparticlesB200<GridT>;
  using MfieldsState = MfieldsStateB200<GridT>;
  using Mfields = MfieldsB200<GridT>;

  OutputFieldsItemB200(const GridT& grid, const OutputFieldsItemParamsB200& prm = {})
    : OutputFieldsItemParamsB200{prm}, get_item_(grid)
  {}
  virtual ~OutputFieldsItemB200() {}

  // DiagnosticBase::perform_diagnostic; `timestep` = grid.timestep() inside PSC
  virtual void perform_diagnostic(Mparticles& mprts, MfieldsState& mflds, int timestep)
  {
    const GridT& grid = mflds.grid();
    bool do_pfield = pfield.do_out(timestep);
    bool do_tfield = tfield.do_out(timestep);
    bool do_tfield_accum = tfield.do_accum(timestep);
    if (!(do_pfield || do_tfield_accum)) {
      return;
    }
    Mfields& item = get_item_.get_item(mprts, mflds);
    if (do_pfield) {
      if (!io_pfd_) {
        io_pfd_.open("pfd" + GetItem::suffix(), pfield.data_dir);
      }
      io_pfd_.write_step(grid, pfield.rn, pfield.rx, host_item(item, timestep), get_item_.name(),
                         get_item_.comp_names());
    }
    if (do_tfield_accum) {
      if (!tfd_) {
        tfd_.reset(new Mfields{grid, item.n_comps()});
      }
      tfd_->add(item);
      naccum_++;
    }
    if (do_tfield && naccum_ > 0) {
      // (naccum_ == 0 happens at the initial output when average_length < out_interval; the
      // reference dereferences its unallocated tfd_ there)
      if (!io_tfd_) {
        io_tfd_.open("tfd" + GetItem::suffix(), tfield.data_dir);
      }
      // convert accumulated values to correct temporal mean
      tfd_->scale(1. / naccum_);
      io_tfd_.write_step(grid, tfield.rn, tfield.rx, host_item(*tfd_, timestep), get_item_.name(),
                         get_item_.comp_names());
      naccum_ = 0;
      tfd_->zero();
    }
  }

  Writer& io_pfd() { return io_pfd_; }
  Writer& io_tfd() { return io_tfd_; }

private:
  static HostItemB200 host_item(const Mfields& m, int timestep)
  {
    HostItemB200 h;
    h.timestep = timestep;
    h.data = m.download_interior();
    h.n_patches = m.n_patches();
    h.n_comps = m.n_comps();
    auto im = m.im(), ibn = m.ibn();
    h.ldims = {im[0] - 2 * ibn[0], im[1] - 2 * ibn[1], im[2] - 2 * ibn[2]};
    return h;
  }

  GetItem get_item_;
  Writer io_pfd_;
  Writer io_tfd_;
  std::unique_ptr<Mfields> tfd_;
  int naccum_ = 0;
};

template <typename GridT, typename Writer = WriterMemoryB200>
using OutputFieldsB200 = OutputFieldsItemB200<GridT, GetItemJehB200<GridT>, Writer>; // :238-243
template <typename GridT, typename Writer = WriterMemoryB200>
using OutputMomentsB200 = OutputFieldsItemB200<GridT, GetItemMomentsB200<GridT>, Writer>; // :245-250

// Collision (psc.hxx:111,363-371; decks: `Collision collision{grid, interval, nu}`, e.g.
// psc_bubble_yz.cxx:303-306): binary Coulomb collisions inside every cell on the device
// (psc_b200_collide: CollisionHost's pairing, psc_collision_impl.hxx:56-252, around
// BinaryCollision, binary_collision.hxx:57-295).  Like the reference's CUDA operator it brings
// its own random streams (counter-based, keyed by seed / time step / cell / pair); Psc::step
// calls it right after the sort, which is the order it needs (the pairing walks cell runs; an
// unordered store is sorted first).  CollisionViaHostB200 below remains for a deck that wants
// PSC's host operator verbatim (the CollisionCudaHost round trip).
template <typename GridT>
struct CollisionB200
{
  using Mparticles = MparticlesB200<GridT>;
  CollisionB200(const GridT& grid, int interval, double nu, uint64_t seed = 0)
    : interval_(interval), nu_(nu), cori_(grid.norm.cori), seed_(seed)
  {}
  int interval() const { return interval_; }
  double nu() const { return nu_; }
  void operator()(Mparticles& mprts)
  {
    psc_b200_collision_params prm;
    prm.interval = interval_;
    prm.nu = nu_;
    prm.cori = cori_;
    prm.rng = rng_;
    prm.seed = seed_;
    prm.step = n_calls_++; // (a fresh stream per call; Psc::step does not hand the time step down)
    PSC_B200_CHECK(psc_b200_collide(mprts.ctx(), &prm, nullptr));
  }
  // 0 = RngFake (uniform() = .5, identity permutation): the reference's known-answer setting
  void set_rng(int rng) { rng_ = rng; }

private:
  int interval_;
  double nu_, cori_;
  uint64_t seed_;
  uint64_t n_calls_ = 0;
  int rng_ = 1;
};

// Heating (decks: `HeatingSelector<Mparticles>::Heating heating{grid, interval, HeatingSpotFoil<Dim>{grid, params}}`,
// psc_flatfoil_yz.cxx:556-570; called from the deck's inject/heat lambda, :649-656): the
// HeatingSpotFoil profile and kick_particle on the device (psc_b200_heating_spot_foil).
// `Spot` is anything with HeatingSpotFoilParams' members (heating_spot_foil.hxx:6-16).
template <typename GridT>
struct HeatingB200
{
  using Mparticles = MparticlesB200<GridT>;
  template <typename Spot>
  HeatingB200(const GridT&, int interval, const Spot& spot, uint64_t seed = 0)
  {
    std::memset(&prm_, 0, sizeof(prm_));
    prm_.zl = spot.zl, prm_.zh = spot.zh, prm_.xc = spot.xc, prm_.yc = spot.yc, prm_.rH = spot.rH;
    prm_.Mi = spot.Mi;
    prm_.n_kinds = spot.n_kinds;
    for (int k = 0; k < spot.n_kinds && k < PSC_B200_MAX_KINDS; k++) {
      prm_.T[k] = spot.T[k];
    }
    prm_.interval = interval;
    prm_.seed = seed;
  }
  void operator()(Mparticles& mprts)
  {
    prm_.step = n_calls_++;
    PSC_B200_CHECK(psc_b200_heating_spot_foil(mprts.ctx(), &prm_, nullptr));
  }

private:
  psc_b200_heating_params prm_;
  uint64_t n_calls_ = 0;
};

// Host round trip: the particles come back as PSC's 32-byte records (patch by patch, off[p]..off[p+1]),
// `collide(prts, off)` changes momenta in place (a lambda around Collision_<MparticlesSingle, ...>), the
// records go back.  Counts per patch must not change.
template <typename GridT, typename HostCollide>
struct CollisionViaHostB200
{
  using Mparticles = MparticlesB200<GridT>;
  CollisionViaHostB200(const GridT&, int interval, double nu, HostCollide collide)
    : interval_(interval), nu_(nu), collide_(collide)
  {}
  int interval() const { return interval_; }
  double nu() const { return nu_; }
  void operator()(Mparticles& mprts)
  {
    std::vector<Particle> prts;
    std::vector<uint32_t> off;
    mprts.get(prts, off);
    collide_(prts, off);
    std::vector<uint32_t> n_by_patch(off.size() - 1);
    for (size_t p = 0; p + 1 < off.size(); p++) {
      n_by_patch[p] = off[p + 1] - off[p];
    }
    mprts.set(prts, n_by_patch);
  }

private:
  int interval_;
  double nu_;
  HostCollide collide_;
};

// ChecksParams (checks_params.hxx): the cadence/threshold fields the step loop reads
struct ChecksParamsB200
{
  int continuity_every_step = 0;
  double continuity_threshold = 1e-13;
  bool continuity_verbose = false; // print_max_err_always
  bool continuity_exit_on_failure = false;
  int gauss_every_step = 0;
  double gauss_threshold = 1e-13;
  bool gauss_verbose = false;
  bool gauss_exit_on_failure = false;
};

// what psc::checks::continuity / gauss do with max_err (checks_impl.hxx:114-123, 198-207;
// CheckParams, checks_params.hxx:3-29): print it when asked to or when it exceeds the
// threshold, abort only if exit_on_failure is set (the default 1e-13 threshold is always
// exceeded by a single-precision run, and the reference simply carries on)
inline void checks_report(const char* what, double max_err, double threshold, bool print_always,
                          bool exit_on_failure)
{
  if (print_always || max_err > threshold) {
    std::printf("%s: max_err = %g (thres %g)\n", what, max_err, threshold);
  }
  if (exit_on_failure && max_err >= threshold) {
    std::fprintf(stderr, "psc_b200: %s check failed (exit_on_failure)\n", what);
    std::abort();
  }
}

template <typename GridT>
struct ChecksB200
{
  using MfieldsState = MfieldsStateB200<GridT>;
  using Mparticles = MparticlesB200<GridT>;

  // checks_impl.hxx:33-132
  struct Continuity
  {
    int every_step;
    double threshold;
    bool print_max_err_always = true;
    bool exit_on_failure = false;
    double last_max_err = 0.;
    bool armed = false;
    bool should_do_check(int timestep) const { return every_step > 0 && timestep % every_step == 0; }
    void before_particle_push(Mparticles& mprts, int timestep)
    {
      armed = should_do_check(timestep);
      if (armed) {
        PSC_B200_CHECK(psc_b200_check_continuity_begin(mprts.ctx()));
      }
    }
    // the shape Psc::step uses (psc.hxx:379-383): the caller has tested should_do_check(timestep)
    void before_particle_push(Mparticles& mprts)
    {
      armed = true;
      PSC_B200_CHECK(psc_b200_check_continuity_begin(mprts.ctx()));
    }
    void after_particle_push(Mparticles& mprts, MfieldsState&)
    {
      if (armed) {
        PSC_B200_CHECK(psc_b200_check_continuity_end(mprts.ctx(), &last_max_err));
        checks_report("continuity", last_max_err, threshold, print_max_err_always, exit_on_failure);
        armed = false;
      }
    }
  };
  // checks_impl.hxx:137-215
  struct Gauss
  {
    int every_step;
    double threshold;
    bool print_max_err_always = true;
    bool exit_on_failure = false;
    double last_max_err = 0.;
    bool should_do_check(int timestep) const { return every_step > 0 && timestep % every_step == 0; }
    void operator()(Mparticles& mprts, MfieldsState&, int timestep)
    {
      if (should_do_check(timestep)) {
        PSC_B200_CHECK(psc_b200_check_gauss(mprts.ctx(), &last_max_err));
        checks_report("gauss", last_max_err, threshold, print_max_err_always, exit_on_failure);
      }
    }
    // the shape Psc::step uses (psc.hxx:234-237,478-482): the caller has tested should_do_check
    void operator()(Mparticles& mprts, MfieldsState&)
    {
      PSC_B200_CHECK(psc_b200_check_gauss(mprts.ctx(), &last_max_err));
      checks_report("gauss", last_max_err, threshold, print_max_err_always, exit_on_failure);
    }
  };

  ChecksB200(const GridT&, const ChecksParamsB200& prm)
    : continuity{prm.continuity_every_step, prm.continuity_threshold, prm.continuity_verbose,
                 prm.continuity_exit_on_failure},
      gauss{prm.gauss_every_step, prm.gauss_threshold, prm.gauss_verbose, prm.gauss_exit_on_failure}
  {}
  // a deck's `Checks checks{grid, MPI_COMM_WORLD, checks_params}` with PSC's ChecksParams
  // (include/checks_params.hxx:3-42: continuity / gauss . check_interval, err_threshold)
  template <typename Comm, typename PscChecksParams>
  ChecksB200(const GridT&, Comm, const PscChecksParams& prm)
    : continuity{prm.continuity.check_interval, prm.continuity.err_threshold,
                 prm.continuity.print_max_err_always, prm.continuity.exit_on_failure},
      gauss{prm.gauss.check_interval, prm.gauss.err_threshold, prm.gauss.print_max_err_always,
            prm.gauss.exit_on_failure}
  {}

  Continuity continuity;
  Gauss gauss;
};

template <typename GridT>
struct BalanceB200
{
  // psc_balance_impl.hxx:770-1026: redistributes patches over the ranks by
  // load = n_prts + factor_fields * n_cells; returns whether anything moved
  explicit BalanceB200(double factor_fields = 1.) : factor_fields_(factor_fields) {}
  bool operator()(MparticlesB200<GridT>& mprts)
  {
    int changed = 0;
    PSC_B200_CHECK(psc_b200_balance(mprts.ctx(), factor_fields_, &changed));
    return changed != 0;
  }

  // The shape Psc::step uses (psc.hxx:346-350): balance_(grid_, mprts_).  The patches (particles
  // and every field container) move between the GPUs inside the library; PSC additionally expects
  // its host Grid_t to be REPLACED by one for the new decomposition (psc_balance_impl.hxx:893-1016).
  // That part is the deck's: `regrid(old_grid, first_local_patch, n_local_patches)` returns the new
  // grid (in a PSC tree: new Grid_t{domain, bc, kinds, norm, dt, n_patches, ibn}); it is only called
  // when something moved.
  using Regrid = std::function<GridT*(GridT*, int, int)>;
  void set_regrid(Regrid regrid) { regrid_ = std::move(regrid); }
  void operator()(GridT*& grid_ptr, MparticlesB200<GridT>& mprts)
  {
    if (!(*this)(mprts)) {
      return;
    }
    if (!regrid_) {
      std::fprintf(stderr, "psc_b200: patches were rebalanced but no regrid hook is installed "
                           "(BalanceB200::set_regrid): the host grid no longer matches the device\n");
      std::abort();
    }
    grid_ptr = regrid_(grid_ptr, psc_b200_patch_begin(mprts.ctx()), psc_b200_n_patches(mprts.ctx()));
    // the device context (particles, every field container, NCCL state) is the one that was just
    // rebalanced: it is re-keyed to the new grid, so mprts, the MfieldsState and every scratch
    // Mfields of this context keep pointing at the same device data
    mprts.context().rebind(*grid_ptr);
    if (psc_b200_n_patches(mprts.ctx()) != grid_ptr->n_patches()) {
      std::fprintf(stderr, "psc_b200: the regrid hook returned a grid with %d patches, the device holds %d\n",
                   grid_ptr->n_patches(), psc_b200_n_patches(mprts.ctx()));
      std::abort();
    }
  }

  double factor_fields_;
  Regrid regrid_;
};

// write_checkpoint / read_checkpoint / Checkpointing (src/include/checkpoint.hxx:14-133): same
// names, arguments and cadence rules; the file set is "checkpoint_<timestep>.b200.<rank>"
// (psc_b200_checkpoint_write: PSC's variable decomposition in a flat binary, no ADIOS2).
template <typename GridT>
inline void write_checkpoint(const GridT& grid, MparticlesB200<GridT>& mprts, MfieldsStateB200<GridT>&)
{
  const std::string filename = "checkpoint_" + std::to_string(grid.timestep()) + ".b200";
  PSC_B200_CHECK(psc_b200_checkpoint_write(mprts.ctx(), filename.c_str(), grid.timestep()));
}
// (the containers must have been constructed on `grid`, which must describe the run that wrote
// the checkpoint; returns the time step the checkpoint was written at)
template <typename GridT>
inline long read_checkpoint(const std::string& filename, GridT&, MparticlesB200<GridT>& mprts,
                            MfieldsStateB200<GridT>&)
{
  int64_t timestep = 0;
  PSC_B200_CHECK(psc_b200_checkpoint_read(mprts.ctx(), filename.c_str(), &timestep));
  return (long)timestep;
}
class CheckpointingB200
{
public:
  explicit CheckpointingB200(int interval) : interval_{interval} {}
  // called every step (checkpoint.hxx:96-113): not right after start-up / restart
  template <typename GridT>
  void operator()(const GridT& grid, MparticlesB200<GridT>& mprts, MfieldsStateB200<GridT>& mflds)
  {
    if (interval_ <= 0) {
      return;
    }
    if (first_time_) {
      first_time_ = false;
      return;
    }
    if (grid.timestep() % interval_ == 0) {
      write_checkpoint(grid, mprts, mflds);
    }
  }
  // after the time loop (checkpoint.hxx:117-126)
  template <typename GridT>
  void final(const GridT& grid, MparticlesB200<GridT>& mprts, MfieldsStateB200<GridT>& mflds)
  {
    if (interval_ > 0) {
      write_checkpoint(grid, mprts, mflds);
    }
  }

private:
  int interval_;
  bool first_time_ = true;
};

// DiagEnergies (DiagEnergiesField.h:19-42, DiagEnergiesParticle.h:15-40)
template <typename GridT>
inline std::array<double, 8> energies(MparticlesB200<GridT>& mprts)
{
  std::array<double, 8> out;
  PSC_B200_CHECK(psc_b200_energies(mprts.ctx(), out.data()));
  return out;
}

// ----------------------------------------------------------------------
// BoundaryInjectorB200: BoundaryInjector<PARTICLE_GENERATOR, PUSH_PARTICLES>
// (src/include/boundary_injector.hxx:66-167), same constructor and inject(mprts, mflds)
// (InjectorBase, injector_base.hxx:5-11), so Psc::add_injector takes it.  Every step, for
// every ghost cell just below the lower y wall, particles of an imaginary unit-density
// population are drawn from the generator, advanced one step in y, and those that enter the
// patch are injected with the current of their way in.  The generator, the cell loop and the
// advance are host code there and stay host code here; the device takes the accepted
// particles (injector(), one H2D append) and the deposit (psc_b200_deposit_j =
// Current::calc_j for each trajectory).  The reference instantiates its template for double
// configurations only (push_x binds Vec3<real_t>& to the generator's Double3); here the same
// statements run in this configuration's real_t = float.
//
// ParticleGenerator: get(min_pos, pos_range) -> record with x[3], u[3], w, kind
// (psc::particle::Inject; positions patch-local).  ParticleGeneratorMaxwellian of
// boundary_injector.hxx:16-57 is host code over PSC's rng.hxx and works unchanged.

template <typename PARTICLE_GENERATOR, typename GridT>
class BoundaryInjectorB200
{
  static const int INJECT_DIM_IDX_ = 1;

public:
  using ParticleGenerator = PARTICLE_GENERATOR;
  using Mparticles = MparticlesB200<GridT>;
  using MfieldsState = MfieldsStateB200<GridT>;
  using real_t = float;

  BoundaryInjectorB200(ParticleGenerator particle_generator, const GridT& grid)
    : particle_generator_(particle_generator),
      dt_(real_t(grid.dt)),
      prts_per_unit_density_(real_t(grid.norm.prts_per_unit_density))
  {}
  virtual ~BoundaryInjectorB200() {}

  // get_n_in_cell(1.0, prts_per_unit_density, true) (setup_particles.hxx:110-122): the density
  // plus a uniform draw, truncated.  A deck inside PSC passes ::get_n_in_cell's own generator
  // through this hook; the default draws from <random>.
  std::function<int()> n_in_cell;

  int n_injected() const { return n_injected_; }

  virtual void inject(Mparticles& mprts, MfieldsState& mflds)
  {
    const GridT& grid = mprts.grid();
    std::vector<psc_b200_jpath> paths;
    n_injected_ = 0;
    {
      auto injectors_by_patch = mprts.injector();
      real_t dxi[3];
      for (int d = 0; d < 3; d++) {
        dxi[d] = real_t(double(grid.domain.gdims[d]) / grid.domain.length[d]); // Real3 dxi = dx_inv
      }
      for (int p = 0; p < grid.n_patches(); p++) {
        if (grid.patches[p].off[INJECT_DIM_IDX_] != 0) { // !grid.atBoundaryLo(p, 1)
          continue;
        }
        int ilo[3] = {0, 0, 0}, ihi[3] = {grid.ldims[0], grid.ldims[1], grid.ldims[2]};
        ilo[INJECT_DIM_IDX_] = -1;
        ihi[INJECT_DIM_IDX_] = 0;
        auto injector = injectors_by_patch[p];
        // VecRange(ilo, ihi): row-major, the last index runs fastest (kg/VecRange.hxx:5-24)
        for (int i0 = ilo[0]; i0 < ihi[0]; i0++) {
          for (int i1 = ilo[1]; i1 < ihi[1]; i1++) {
            for (int i2 = ilo[2]; i2 < ihi[2]; i2++) {
              const int initial_idx[3] = {i0, i1, i2};
              auto cell_corner = grid.domain.dx; // (same vector type as dx; narrowed like Real3)
              for (int d = 0; d < 3; d++) {
                cell_corner[d] = real_t(double(initial_idx[d]) * grid.domain.dx[d]);
              }
              const int n_prts_to_try_inject = n_in_cell ? n_in_cell() : default_n_in_cell();
              for (int cnt = 0; cnt < n_prts_to_try_inject; cnt++) {
                auto prt = particle_generator_.get(cell_corner, grid.domain.dx);
                // calc_v (pushp.hxx:68-72)
                const real_t u[3] = {real_t(prt.u[0]), real_t(prt.u[1]), real_t(prt.u[2])};
                const real_t root = real_t(1.) / std::sqrt(real_t(1.) + u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
                const real_t v[3] = {u[0] * root, u[1] * root, u[2] * root};
                const real_t initial_x[3] = {real_t(prt.x[0]), real_t(prt.x[1]), real_t(prt.x[2])};
                real_t x[3] = {initial_x[0], initial_x[1], initial_x[2]};
                // push_x of AdvanceParticle<real_t, dim_y> (pushp.hxx:17-29): y only ("can't
                // move along x or z, or else might leave patch")
                x[INJECT_DIM_IDX_] += real_t(1.) * dt_ * v[INJECT_DIM_IDX_];
                if (x[INJECT_DIM_IDX_] < real_t(0.)) {
                  continue; // don't inject a particle that fails to enter the patch
                }
                // injectors expect global positions, the deposit patch-local ones (:131-135)
                auto prt_with_global_x = prt;
                for (int d = 0; d < 3; d++) {
                  prt_with_global_x.x[d] = double(x[d]) + grid.patches[p].xb[d];
                }
                injector(prt_with_global_x);
                n_injected_++;

                psc_b200_jpath path;
                path.patch = p;
                for (int d = 0; d < 3; d++) {
                  path.lg[d] = initial_idx[d];
                  path.xm[d] = initial_x[d] * dxi[d];
                  path.xp[d] = x[d] * dxi[d];
                  path.v[d] = v[d];
                }
                path.qni_wni = real_t(grid.kinds[prt.kind].q * prt.w);
                paths.push_back(path);
              }
            }
          }
        }
      }
    } // (the buffered injector appends on destruction)
    (void)mflds;
    PSC_B200_CHECK(psc_b200_deposit_j(mprts.ctx(), paths.data(), paths.size()));
  }

private:
  int default_n_in_cell()
  {
    static std::mt19937 gen(0);
    static std::uniform_real_distribution<float> dist(0.f, 1.f);
    return int(real_t(1.) * prts_per_unit_density_ + dist(gen));
  }

  ParticleGenerator particle_generator_;
  real_t dt_;
  real_t prts_per_unit_density_;
  int n_injected_ = 0;
};

// ----------------------------------------------------------------------
// the config bundle (src/psc_config.hxx:99-146)

template <typename _Dim, typename GridT>
struct PscConfig
{
  using Dim = _Dim;
  using Grid = GridT;
  using Mparticles = MparticlesB200<GridT>;
  using MfieldsState = MfieldsStateB200<GridT>;
  using Mfields = MfieldsB200<GridT>;
  using PushParticles = PushParticlesB200<GridT>;
  using Sort = SortB200<GridT>;
  using PushFields = PushFieldsB200<GridT>;
  using BndParticles = BndParticlesB200<GridT>;
  using Bnd = BndB200<GridT>;
  using BndFields = BndFieldsB200<GridT>;
  using Balance = BalanceB200<GridT>;
  using Checks = ChecksB200<GridT>;
  using Marder = MarderB200<GridT>;
  using Collision = CollisionB200<GridT>; // psc.hxx:111
#ifdef PSC_B200_WITH_PSC_GRID
  // PSC's own particle output works through accessor() (psc_config.hxx:94)
  using OutputParticles = OutputParticlesDefault<Mparticles>;
#endif
};

// ----------------------------------------------------------------------
// Step: the operator sequence of Psc::step (src/include/psc.hxx:321-486) over a config,
// with PscParams' cadence (psc.hxx:66-83).  `fused` routes the same sequence through the
// single entry point psc_b200_step, which lets the library fuse the particle boundary
// exchange with the following sort (same results).

struct PscParamsB200
{
  int sort_interval = 0;
  int marder_interval = 0;
  double marder_diffusion = 0.9;
  int marder_loop = 3;
  bool fused = true;
};

template <typename Config>
class Step
{
public:
  using GridT = typename Config::Grid;
  using Mparticles = typename Config::Mparticles;
  using MfieldsState = typename Config::MfieldsState;

  Step(const GridT& grid, MfieldsState& mflds, Mparticles& mprts, const PscParamsB200& p,
       const ChecksParamsB200& cp = {})
    : p_(p), mflds_(mflds), mprts_(mprts), bndp_(grid), marder_(grid, p.marder_diffusion, p.marder_loop, false),
      checks_(grid, cp)
  {}

  // psc.hxx:220-238 initialize(): ghost fills before the first step
  void initialize()
  {
    bndf_.fill_ghosts_H(mflds_);
    bnd_.fill_ghosts(mflds_, PSC_B200_HX, PSC_B200_HX + 3);
    bnd_.fill_ghosts(mflds_, PSC_B200_JXI, PSC_B200_JXI + 3);
    bndf_.fill_ghosts_E(mflds_);
    bnd_.fill_ghosts(mflds_, PSC_B200_EX, PSC_B200_EX + 3);
  }

  // Psc::add_injector (psc.hxx:172-176): anything with inject(mprts, mflds).  The injectors
  // act between the push and the particle exchange (psc.hxx:391-399), so a step with
  // injectors is issued operator by operator instead of as the single fused call.
  void add_injector(std::function<void(Mparticles&, MfieldsState&)> injector)
  {
    injectors_.push_back(std::move(injector));
  }
  template <typename Injector>
  void add_injector(Injector* injector)
  {
    assert(injector);
    injectors_.push_back([injector](Mparticles& mprts, MfieldsState& mflds) { injector->inject(mprts, mflds); });
  }

  // Psc::add_diagnostic / perform_diagnostics / integrate (psc.hxx:184-198, 514-528, 243-310)
  void add_diagnostic(std::function<void(Mparticles&, MfieldsState&, int)> diagnostic)
  {
    diagnostics_.push_back(std::move(diagnostic));
  }
  template <typename Diagnostic>
  void add_diagnostic(Diagnostic* diagnostic)
  {
    assert(diagnostic);
    diagnostics_.push_back([diagnostic](Mparticles& mprts, MfieldsState& mflds, int timestep) {
      diagnostic->perform_diagnostic(mprts, mflds, timestep);
    });
  }
  void perform_diagnostics()
  {
    for (auto& diagnostic : diagnostics_) {
      diagnostic(mprts_, mflds_, timestep_);
    }
  }
  void integrate(int nmax)
  {
    initialize();
    perform_diagnostics(); // initial output
    while (timestep_ < nmax) {
      (*this)();
      perform_diagnostics();
    }
  }

  void operator()()
  {
    const int t = ++timestep_;
    const bool do_sort = p_.sort_interval > 0 && t % p_.sort_interval == 0;
    const bool do_marder = p_.marder_interval > 0 && t % p_.marder_interval == 0;
    if (p_.fused && injectors_.empty()) {
      psc_b200_step_params sp;
      sp.sort = do_sort;
      sp.marder_loop = do_marder ? p_.marder_loop : 0;
      sp.marder_diffusion = p_.marder_diffusion;
      sp.push_fields = 1;
      sp.checks = checks_.continuity.should_do_check(t) || checks_.gauss.should_do_check(t);
      sp.energies = 0;
      PSC_B200_CHECK(psc_b200_step(mprts_.ctx(), &sp));
      if (sp.checks) {
        PSC_B200_CHECK(psc_b200_last_checks(mprts_.ctx(), &checks_.continuity.last_max_err,
                                            &checks_.gauss.last_max_err));
      }
      return;
    }
    typename Config::Dim dim{};
    if (do_sort) {
      sort_(mprts_); // psc.hxx:356-361
    }
    checks_.continuity.before_particle_push(mprts_, t); // :379-384
    pushp_.push_mprts(mprts_, mflds_);                  // :389
    for (auto& injector : injectors_) {                 // :391-399
      injector(mprts_, mflds_);
    }
    bndp_(mprts_);                                      // :412
    bndf_.add_ghosts_J(mflds_);                         // :417
    bnd_.add_ghosts(mflds_, PSC_B200_JXI, PSC_B200_JXI + 3);  // :418
    bnd_.fill_ghosts(mflds_, PSC_B200_JXI, PSC_B200_JXI + 3); // :419
    pushf_.push_H(mflds_, .5, dim);                     // :426
    bndf_.fill_ghosts_H(mflds_);                        // :428
    bnd_.fill_ghosts(mflds_, PSC_B200_HX, PSC_B200_HX + 3); // :432
    pushf_.push_E(mflds_, 1., dim);                     // :439
    bndf_.fill_ghosts_E(mflds_);                        // :441
    bnd_.fill_ghosts(mflds_, PSC_B200_EX, PSC_B200_EX + 3); // :445
    if (do_marder) {
      marder_(mflds_, mprts_); // :448-455
    }
    pushf_.push_H(mflds_, .5, dim);                     // :461
    bndf_.fill_ghosts_H(mflds_);                        // :463
    bnd_.fill_ghosts(mflds_, PSC_B200_HX, PSC_B200_HX + 3); // :467
    checks_.continuity.after_particle_push(mprts_, mflds_); // :471-476
    checks_.gauss(mprts_, mflds_, t);                   // :479-483
  }

  int timestep() const { return timestep_; }
  typename Config::Checks& checks() { return checks_; }

private:
  PscParamsB200 p_;
  MfieldsState& mflds_;
  Mparticles& mprts_;
  typename Config::Sort sort_;
  typename Config::PushParticles pushp_;
  typename Config::PushFields pushf_;
  typename Config::Bnd bnd_;
  typename Config::BndFields bndf_;
  typename Config::BndParticles bndp_;
  typename Config::Marder marder_;
  typename Config::Checks checks_;
  std::vector<std::function<void(Mparticles&, MfieldsState&)>> injectors_;
  std::vector<std::function<void(Mparticles&, MfieldsState&, int)>> diagnostics_;
  int timestep_ = 0;
};

} // namespace psc_b200

// inside a PSC tree (grid.hxx included first): the name a deck uses
#ifdef PSC_B200_WITH_PSC_GRID
template <typename Dim>
using PscConfig1vbecB200 = psc_b200::PscConfig<Dim, Grid_t>;
#endif
