// psc_b200: device-side helpers shared by the kernels -- field indexing, patch
// lookup, a device-wide exclusive scan and launch plumbing.
#pragma once

#include "ctx.hpp"

namespace psc_b200
{

// Fields3d offset (fields3d.hxx:29-32): ix fastest, lower bound -ibn
__device__ __forceinline__ long fld_off(const GridDev& G, int m, int i, int j, int k)
{
  return (((long)m * G.im[2] + (k + G.ibn[2])) * G.im[1] + (j + G.ibn[1])) * G.im[0] +
         (i + G.ibn[0]);
}

// patch of particle i: last p with off[p] <= i
__device__ __forceinline__ int patch_of(const uint32_t* __restrict__ off, int n_patches,
                                        uint32_t i)
{
  int lo = 0, hi = n_patches; // invariant: off[lo] <= i < off[hi]
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (__ldg(&off[mid]) <= i) {
      lo = mid;
    } else {
      hi = mid;
    }
  }
  return lo;
}

inline unsigned div_up(size_t a, size_t b)
{
  return (unsigned)((a + b - 1) / b);
}

int check_launch(Ctx* c, const char* what);

// ----------------------------------------------------------------------
// exclusive scan over n values f(0..n-1); writes out[0..n] (out[n] = total)

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

template <typename T>
__device__ __forceinline__ T warp_incl_scan(T v, int lane)
{
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    T t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) {
      v += t;
    }
  }
  return v;
}

template <typename T, typename F>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_reduce(F f, size_t n, T* block_sums)
{
  __shared__ T ws[SCAN_THREADS / 32];
  size_t base = (size_t)blockIdx.x * SCAN_TILE;
  T s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    size_t i = base + (size_t)k * SCAN_THREADS + threadIdx.x;
    if (i < n) {
      s += f(i);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
  }
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) {
    ws[w] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    T t = 0;
    for (int k = 0; k < SCAN_THREADS / 32; k++) {
      t += ws[k];
    }
    block_sums[blockIdx.x] = t;
  }
}

template <typename T, typename F>
__global__ void __launch_bounds__(SCAN_THREADS)
  k_scan_apply(F f, size_t n, const T* block_off, T* out)
{
  __shared__ T ws[SCAN_THREADS / 32];
  size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
  T v[SCAN_ITEMS];
  T s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    size_t i = base + k;
    v[k] = i < n ? f(i) : T(0);
    s += v[k];
  }
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  T incl = warp_incl_scan(s, lane);
  if (lane == 31) {
    ws[w] = incl;
  }
  __syncthreads();
  T woff = 0;
  for (int k = 0; k < w; k++) {
    woff += ws[k];
  }
  T run = (block_off ? block_off[blockIdx.x] : T(0)) + woff + incl - s;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    size_t i = base + k;
    if (i < n) {
      out[i] = run;
      run += v[k];
      if (i == n - 1) {
        out[n] = run;
      }
    }
  }
}

template <typename T>
struct LoadArr
{
  const T* p;
  __device__ __forceinline__ T operator()(size_t i) const { return p[i]; }
};

// out may alias the array f reads from (each element is read, then written, by the
// same thread).  scratch is grown as needed.
template <typename T, typename F>
int scan_exclusive(Ctx* c, F f, size_t n, T* out, DevBuf& scratch)
{
  if (n == 0) {
    PSC_CUDA_TRY(cudaMemsetAsync(out, 0, sizeof(T), c->stream));
    return 0;
  }
  // level sizes
  std::vector<size_t> nb;
  for (size_t m = n; m > 1 || nb.empty();) {
    m = (m + SCAN_TILE - 1) / SCAN_TILE;
    nb.push_back(m);
    if (m == 1) {
      break;
    }
  }
  size_t tot = 0;
  for (size_t m : nb) {
    tot += m + 1;
  }
  PSC_TRY(scratch.reserve(tot * sizeof(T)));
  std::vector<T*> lvl(nb.size());
  {
    T* p = scratch.as<T>();
    for (size_t l = 0; l < nb.size(); l++) {
      lvl[l] = p;
      p += nb[l] + 1;
    }
  }
  // reduce up
  k_scan_reduce<T, F><<<(unsigned)nb[0], SCAN_THREADS, 0, c->stream>>>(f, n, lvl[0]);
  for (size_t l = 1; l < nb.size(); l++) {
    k_scan_reduce<T, LoadArr<T>><<<(unsigned)nb[l], SCAN_THREADS, 0, c->stream>>>(
      LoadArr<T>{lvl[l - 1]}, nb[l - 1], lvl[l]);
  }
  // scan down
  for (size_t l = nb.size(); l-- > 1;) {
    // level l-1 (nb[l-1] entries) scanned in place with block offsets from level l
    k_scan_apply<T, LoadArr<T>><<<(unsigned)nb[l], SCAN_THREADS, 0, c->stream>>>(
      LoadArr<T>{lvl[l - 1]}, nb[l - 1], (nb[l] > 1) ? lvl[l] : nullptr, lvl[l - 1]);
  }
  k_scan_apply<T, F><<<(unsigned)nb[0], SCAN_THREADS, 0, c->stream>>>(
    f, n, nb[0] > 1 ? lvl[0] : nullptr, out);
  c->n_launches += 2 * nb.size();
  return check_launch(c, "scan_exclusive");
}

} // namespace psc_b200
