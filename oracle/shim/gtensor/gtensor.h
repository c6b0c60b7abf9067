// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// Minimal stand-in for the un-vendored gtensor library (wdmapp/gtensor@main,
// fetched by CPM in /root/reference/CMakeLists.txt:61-66), just enough for the
// reference's own arithmetic headers (kg/Vec3.h, pushp.hxx, interpolate.hxx,
// fields.hxx, psc/current_deposition.hxx, inc_curr_1vb_{split,var1}.cxx,
// push_particles_1vb.hxx) to compile UNMODIFIED from where they lie under
// /root/reference.  Only gt::sarray, the GT_INLINE/GT_LAMBDA macros and the
// _all/_s slicing placeholders are provided; no expression templates.
#pragma once

#include <cstddef>
#include <initializer_list>

#define GT_INLINE inline
#define GT_LAMBDA

namespace gt
{

template <typename T, std::size_t N>
struct sarray
{
  using value_type = T;

  sarray() : v_{} {}
  sarray(std::initializer_list<T> il)
  {
    std::size_t i = 0;
    for (auto it = il.begin(); it != il.end() && i < N; ++it, ++i) {
      v_[i] = *it;
    }
    for (; i < N; i++) {
      v_[i] = T{};
    }
  }

  T& operator[](std::size_t i) { return v_[i]; }
  const T& operator[](std::size_t i) const { return v_[i]; }

  T* data() { return v_; }
  const T* data() const { return v_; }
  T* begin() { return v_; }
  T* end() { return v_ + N; }
  const T* begin() const { return v_; }
  const T* end() const { return v_ + N; }
  static constexpr std::size_t size() { return N; }

  T v_[N];
};

namespace placeholders
{
struct all_t
{};
static constexpr all_t _all{};
struct slice_t
{
  int b, e;
};
inline slice_t _s(int b, int e) { return {b, e}; }
} // namespace placeholders

} // namespace gt
