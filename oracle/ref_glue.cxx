// TEST INFRASTRUCTURE ONLY -- never linked into, imported by or called from the
// product path (psc_b200/).  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load the library built from this.
//
// oracle/_ref/libpsc_ref.so: the reference's OWN arithmetic for the particle
// hot path, compiled unmodified from where it lies under /root/reference:
//
//   src/kg/include/kg/Vec3.h
//   src/include/{dim,pushp,interpolate,fields}.hxx, psc_bits.h, cuda_compat.h
//   src/include/psc/current_deposition.hxx
//   src/libpsc/psc_push_particles/{inc_defs.h,inc_curr_1vb_split.cxx,
//                                  inc_curr_1vb_var1.cxx,push_particles_1vb.hxx}
//
// i.e. PushParticlesVb<C>::push_mprts (push_particles_1vb.hxx:27-84) itself is
// what runs here, instantiated over small mock containers that satisfy its
// duck-typed interface (the real containers need gtensor/libmrc/MPI, which this
// image does not have -- SURVEY.md D3).  The only restated piece is
// curr_cache_t (push_config.hxx:17-35, ten lines), because push_config.hxx
// drags in the MPI-dependent container headers.
//
// Build: see oracle/Makefile (g++ -O3 -DNDEBUG, no -march: PSC's Release flags,
// hence no FMA contraction).

#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <type_traits>
#include <vector>

// psc.h:26-38 field component enum, needed by interpolate.hxx / deposit
enum
{
  JXI,
  JYI,
  JZI,
  EX,
  EY,
  EZ,
  HX,
  HY,
  HZ,
  NR_FIELDS,
};

#include <kg/Vec3.h>
#include <dim.hxx>

using namespace gt::placeholders;

// ----------------------------------------------------------------------
// Grid_t stand-in: exactly the members the included headers touch
// (grid.hxx:141-160, grid/domain.hxx:44-46, grid.hxx:265-293)

struct Grid_t
{
  struct Domain
  {
    Vec3<double> dx, dx_inv;
  } domain;
  struct Norm
  {
    double fnqs, eta;
  } norm;
  struct Kind
  {
    double q, m;
  };
  std::vector<Kind> kinds;
  double dt;
};

struct checks_order_1st
{};

#include <pushp.hxx>
#include <fields.hxx>
#include <interpolate.hxx>
#include <psc/current_deposition.hxx>
#include "../libpsc/psc_push_particles/inc_defs.h"
#include "../libpsc/psc_push_particles/inc_curr_1vb_split.cxx"
#include "../libpsc/psc_push_particles/inc_curr_1vb_var1.cxx"
#include "../libpsc/psc_push_particles/push_particles_1vb.hxx"

namespace
{

// ----------------------------------------------------------------------
// Storage4: non-owning (ix,iy,iz,m) view, x fastest -- the per-patch slice of
// PSC's Mfields storage (fields3d.hxx:29-32,284-291)

template <typename T>
struct Storage4
{
  using value_type = T;
  using shape_type = gt::sarray<int, 4>;

  T* d;
  int n[4];

  shape_type shape() const { return {n[0], n[1], n[2], n[3]}; }
  int shape(int i) const { return n[i]; }

  T& operator()(int i, int j, int k, int m) const
  {
#ifdef PSC_REF_BOUNDS_CHECK
    assert(i >= 0 && i < n[0] && j >= 0 && j < n[1] && k >= 0 && k < n[2] &&
           m >= 0 && m < n[3]);
#endif
    return d[((size_t(m) * n[2] + k) * n[1] + j) * n[0] + i];
  }

  struct CompView
  {
    Storage4 s;
    int mb, me;
    void operator=(T val)
    {
      size_t len = size_t(s.n[0]) * s.n[1] * s.n[2];
      for (size_t i = len * mb; i < len * me; i++) {
        s.d[i] = val;
      }
    }
  };

  CompView view(all_t, all_t, all_t, slice_t sl) const
  {
    return {*this, sl.b, sl.e};
  }
};

template <typename T>
struct FieldsView
{
  using real_t = T;
  using value_type = T;
  using Storage = Storage4<T>;

  Storage st;
  Int3 ib_;

  Storage storage() const { return st; }
  Int3 ib() const { return ib_; }
};

// curr_cache_t restated from push_config.hxx:17-35
template <typename fields_t>
class curr_cache
{
public:
  using real_t = typename fields_t::value_type;
  using value_type = typename fields_t::value_type;
  using storage_type = typename fields_t::Storage;

  curr_cache(fields_t& f) : storage_(f.storage()), ib_(f.ib()) {}
  curr_cache(const fields_t& f) : storage_(f.storage()), ib_(f.ib()) {}

  void add(int m, int i, int j, int k, real_t val)
  {
    storage_(i - ib_[0], j - ib_[1], k - ib_[2], JXI + m) += val;
  }

private:
  storage_type storage_;
  Int3 ib_;
};

// particle record = ParticleSimple<T> (particle_simple.hxx:10-42)
template <typename T>
struct Prt
{
  Vec3<T> x_;
  Vec3<T> u_;
  int kind_;
  T qni_wni_;
};
static_assert(sizeof(Prt<float>) == 32, "ParticleSimple<float> is 32 bytes");

template <typename T>
struct PrtProxy
{
  Prt<T>* p;
  Vec3<T>& x() { return p->x_; }
  Vec3<T>& u() { return p->u_; }
  int kind() const { return p->kind_; }
  T qni_wni() const { return p->qni_wni_; }
};

template <typename T>
struct PrtRange
{
  Prt<T>* b;
  Prt<T>* e;
  struct It
  {
    Prt<T>* p;
    PrtProxy<T> operator*() const { return {p}; }
    It& operator++()
    {
      ++p;
      return *this;
    }
    bool operator!=(const It& o) const { return p != o.p; }
  };
  It begin() const { return {b}; }
  It end() const { return {e}; }
};

template <typename T>
struct MockMprts
{
  using real_t = T;
  const Grid_t* grid_;
  std::vector<PrtRange<T>> patches;

  const Grid_t& grid() const { return *grid_; }
  struct Acc
  {
    MockMprts* m;
    PrtRange<T> operator[](int p) const { return m->patches[p]; }
  };
  Acc accessor_() { return {this}; }
};

template <typename T>
struct MockMflds
{
  using fields_view_t = FieldsView<T>;
  std::vector<FieldsView<T>> patches;
  int n_patches() const { return patches.size(); }
  FieldsView<T> operator[](int p) const { return patches[p]; }
};

template <typename T, typename DIM,
          template <typename, typename, typename> class CURRENT>
struct Cfg
{
  using Mparticles = MockMprts<T>;
  using MfieldsState = MockMflds<T>;
  using Dim = DIM;
  using InterpolateEM_t = InterpolateEM1vbec<Fields3d<Storage4<T>>, DIM>;
  using Current_t = CURRENT<opt_order_1st, DIM, curr_cache<FieldsView<T>>>;
  using AdvanceParticle_t = AdvanceParticle<T, DIM>;
};

Grid_t make_grid(const int gdims[3], const double length[3], double dt,
                 double fnqs, double eta, int n_kinds, const double* q,
                 const double* m)
{
  Grid_t g;
  for (int d = 0; d < 3; d++) {
    // grid/domain.hxx:44-46
    g.domain.dx[d] = length[d] / double(gdims[d]);
    g.domain.dx_inv[d] = double(gdims[d]) / length[d];
  }
  g.norm.fnqs = fnqs;
  g.norm.eta = eta;
  g.dt = dt;
  for (int k = 0; k < n_kinds; k++) {
    g.kinds.push_back({q[k], m[k]});
  }
  return g;
}

template <typename T, typename DIM,
          template <typename, typename, typename> class CURRENT>
void run_push(const Grid_t& grid, T* flds, const int im[3], const int ib[3],
              int n_patches, void* prts, const unsigned* off)
{
  using C = Cfg<T, DIM, CURRENT>;
  MockMprts<T> mprts;
  mprts.grid_ = &grid;
  MockMflds<T> mflds;
  size_t patch_len = size_t(im[0]) * im[1] * im[2] * NR_FIELDS;
  auto* p0 = static_cast<Prt<T>*>(prts);
  for (int p = 0; p < n_patches; p++) {
    mprts.patches.push_back({p0 + off[p], p0 + off[p + 1]});
    FieldsView<T> fv{{flds + p * patch_len, {im[0], im[1], im[2], NR_FIELDS}},
                     {ib[0], ib[1], ib[2]}};
    mflds.patches.push_back(fv);
  }
  PushParticlesVb<C>::push_mprts(mprts, mflds);
}

template <typename T, typename DIM,
          template <typename, typename, typename> class CURRENT>
void run_calc_j(const Grid_t& grid, T* flds, const int im[3], const int ib[3],
                const double xm_[3], const double xp_[3], const double vxi_[3],
                double qni_wni)
{
  using Current = CURRENT<opt_order_1st, DIM, curr_cache<FieldsView<T>>>;
  FieldsView<T> fv{{flds, {im[0], im[1], im[2], NR_FIELDS}},
                   {ib[0], ib[1], ib[2]}};
  curr_cache<FieldsView<T>> J(fv);
  Current curr(grid);
  Vec3<T> xm = {T(xm_[0]), T(xm_[1]), T(xm_[2])};
  Vec3<T> xp = {T(xp_[0]), T(xp_[1]), T(xp_[2])};
  Vec3<T> vxi = {T(vxi_[0]), T(vxi_[1]), T(vxi_[2])};
  Int3 lg = xm.fint(), lf = xp.fint();
  curr.calc_j(J, xm, xp, lf, lg, T(qni_wni), vxi);
}

} // namespace

// ======================================================================
// C entry points (loaded with ctypes from tests/ and bench.py only)

extern "C" {

// dim: 0 = xyz, 1 = yz.  deposit: 0 = Current1vbVar1 (yz only), 1 = Split.
// flds: n_patches x (im0*im1*im2*9) floats, PSC layout.
// prts: AoS 32-byte ParticleSimple<float> records; off[n_patches+1].
// returns 0 on success, -1 for an unsupported combination.
int psc_ref_push_mprts(int dim, int deposit, const int gdims[3],
                       const double length[3], double dt, double fnqs,
                       double eta, int n_kinds, const double* q,
                       const double* m, float* flds, const int im[3],
                       const int ib[3], int n_patches, void* prts,
                       const unsigned* off)
{
  Grid_t grid = make_grid(gdims, length, dt, fnqs, eta, n_kinds, q, m);
  if (dim == 0 && deposit == 1) {
    run_push<float, dim_xyz, Current1vbSplit>(grid, flds, im, ib, n_patches,
                                              prts, off);
  } else if (dim == 1 && deposit == 1) {
    run_push<float, dim_yz, Current1vbSplit>(grid, flds, im, ib, n_patches,
                                             prts, off);
  } else if (dim == 1 && deposit == 0) {
    run_push<float, dim_yz, Current1vbVar1>(grid, flds, im, ib, n_patches, prts,
                                            off);
  } else {
    return -1;
  }
  return 0;
}

// single-trajectory deposit (what test_current_deposition.cxx:80-101 drives).
// real: 0 = float, 1 = double.  flds: im0*im1*im2*9 of that type, zeroed by caller.
int psc_ref_calc_j(int real, int dim, int deposit, const int gdims[3],
                   const double length[3], double dt, double fnqs, void* flds,
                   const int im[3], const int ib[3], const double xm[3],
                   const double xp[3], const double vxi[3], double qni_wni)
{
  Grid_t grid = make_grid(gdims, length, dt, fnqs, 1., 0, nullptr, nullptr);
#define CASE(R, T, D, DIMT, DEP, CUR)                                          \
  if (real == R && dim == D && deposit == DEP) {                               \
    run_calc_j<T, DIMT, CUR>(grid, static_cast<T*>(flds), im, ib, xm, xp, vxi, \
                             qni_wni);                                         \
    return 0;                                                                  \
  }
  CASE(0, float, 0, dim_xyz, 1, Current1vbSplit)
  CASE(0, float, 1, dim_yz, 1, Current1vbSplit)
  CASE(0, float, 1, dim_yz, 0, Current1vbVar1)
  CASE(1, double, 0, dim_xyz, 1, Current1vbSplit)
  CASE(1, double, 1, dim_yz, 1, Current1vbSplit)
  CASE(1, double, 1, dim_yz, 0, Current1vbVar1)
#undef CASE
  return -1;
}

// AdvanceParticle<float,dim_xyz>::push_p (pushp.hxx:36-63) on one particle
void psc_ref_push_p(float u[3], const float E[3], const float H[3], float dq)
{
  AdvanceParticle<float, dim_xyz> adv(1.f);
  Vec3<float> p = {u[0], u[1], u[2]};
  adv.push_p(p, {E[0], E[1], E[2]}, {H[0], H[1], H[2]}, dq);
  u[0] = p[0];
  u[1] = p[1];
  u[2] = p[2];
}

const char* psc_ref_describe()
{
  return "psc-code/psc reference headers (push_particles_1vb.hxx, pushp.hxx, "
         "interpolate.hxx, inc_curr_1vb_{split,var1}.cxx, "
         "psc/current_deposition.hxx) compiled unmodified, g++ -O3 -DNDEBUG";
}
}
