// psc_b200: field-side operators of the hot path, every one a data-parallel kernel over
// all patches of the rank at once (PSC launches per patch / per expression):
//   PushFieldsB200   Yee E/H leapfrog        include/psc_push_fields_impl.hxx:24-178
//   BndB200          fill_ghosts/add_ghosts  libpsc/psc_bnd/psc_bnd_impl.hxx:105-158 with the
//                    box patterns of libmrc/src/mrc_ddc_multi.c:60-135 -- written as a
//                    *gather*: every ghost (fill) or near-boundary interior (add) point
//                    pulls from its neighbour patches, in the reference's summation order,
//                    so there are no atomics and the result is deterministic
//   BndFieldsB200    conducting wall         libpsc/psc_bnd_fields/psc_bnd_fields_impl.hxx:301-530
//   Moment_rho_1st_nc, div_nc, continuity / Gauss checks, Marder correction, field energies
//                    include/psc/moment.hxx:149-171, psc/deposit.hxx:24-65,
//                    psc_output_fields/fields_item_fields.hxx:65-104,
//                    libpsc/psc_checks/checks_impl.hxx:33-215,
//                    libpsc/psc_push_fields/marder_impl.hxx:26-61,197-264,
//                    include/DiagEnergiesField.h:19-42
// Fields keep PSC's layout float [slot][m][iz][iy][ix] (fields3d.hxx:29-32).
#include "dev_util.cuh"

#include <algorithm>
#include <cstring>

namespace psc_b200
{

namespace
{

struct Idx3
{
  int p, m, i, j, k;
};

// decode a linear index over [n_patches][n_m][im2][im1][im0] (array coordinates)
__device__ __forceinline__ Idx3 decode_full(const GridDev& G, size_t idx, int n_m)
{
  Idx3 r;
  r.i = (int)(idx % G.im[0]) - G.ibn[0];
  idx /= G.im[0];
  r.j = (int)(idx % G.im[1]) - G.ibn[1];
  idx /= G.im[1];
  r.k = (int)(idx % G.im[2]) - G.ibn[2];
  idx /= G.im[2];
  r.m = (int)(idx % n_m);
  r.p = (int)(idx / n_m);
  return r;
}

__global__ void k_fill_value(float* F, long slot_len, long fld_len, int n_slots, int m, float v)
{
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < (size_t)n_slots * fld_len) {
    size_t s = idx / fld_len, r = idx % fld_len;
    F[s * slot_len + (size_t)m * fld_len + r] = v;
  }
}

// ---------------------------------------------------------------- ghost exchange

// Enumeration of the points a ghost exchange touches, so that no thread is spent on the
// rest of the patch: ghost shell (fill) or the interior band within ibn of a face (add).
// Three groups: z in band | y in band | x in band (remaining directions not in band).
struct BandGeom
{
  int lo[3], hi[3]; // band widths below / above
  int e[3];         // extent the bands wrap around: im (ghost shell) or ldims (interior band)
  int n2, n1, n0;   // points per group
  int inner_lo[3];  // first non-band coordinate (array coordinates of the extent)
};

inline BandGeom make_band(const GridHost& g, bool ghost_shell)
{
  BandGeom B;
  for (int d = 0; d < 3; d++) {
    if (ghost_shell) {
      B.e[d] = g.im[d];
      B.lo[d] = B.hi[d] = g.ibn[d];
    } else {
      B.e[d] = g.ldims[d];
      B.lo[d] = std::min(g.ibn[d], g.ldims[d]);
      B.hi[d] = std::min(g.ibn[d], g.ldims[d] - B.lo[d]);
    }
    B.inner_lo[d] = B.lo[d];
  }
  int nb[3], in[3];
  for (int d = 0; d < 3; d++) {
    nb[d] = B.lo[d] + B.hi[d];
    in[d] = B.e[d] - nb[d];
  }
  B.n2 = nb[2] * B.e[1] * B.e[0];
  B.n1 = in[2] * nb[1] * B.e[0];
  B.n0 = in[2] * in[1] * nb[0];
  return B;
}

// band index -> coordinate in [0, e)
__device__ __forceinline__ int band_coord(const BandGeom& B, int d, int kk)
{
  return kk < B.lo[d] ? kk : B.e[d] - B.hi[d] + (kk - B.lo[d]);
}

// idx in [0, n2+n1+n0) -> coordinates in [0, e)^3
__device__ __forceinline__ void band_decode(const BandGeom& B, int idx, int& x, int& y, int& z)
{
  if (idx < B.n2) {
    x = idx % B.e[0];
    idx /= B.e[0];
    y = idx % B.e[1];
    z = band_coord(B, 2, idx / B.e[1]);
    return;
  }
  idx -= B.n2;
  if (idx < B.n1) {
    int nb1 = B.lo[1] + B.hi[1];
    x = idx % B.e[0];
    idx /= B.e[0];
    y = band_coord(B, 1, idx % nb1);
    z = B.inner_lo[2] + idx / nb1;
    return;
  }
  idx -= B.n1;
  int nb0 = B.lo[0] + B.hi[0];
  int in1 = B.e[1] - B.lo[1] - B.hi[1];
  x = band_coord(B, 0, idx % nb0);
  idx /= nb0;
  y = B.inner_lo[1] + idx % in1;
  z = B.inner_lo[2] + idx / in1;
}

__global__ void k_fill_ghosts(GridDev G, BandGeom B, float* __restrict__ F, long slot_len, int mb,
                              int me, const int* __restrict__ nei_slot)
{
  // grid.x walks the ghost points of one (patch, component), grid.y strides over the pairs:
  // the point is decoded once and everything stays in 32-bit arithmetic
  const int nb = B.n2 + B.n1 + B.n0;
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nb) {
    return;
  }
  const int n_m = me - mb;
  int x, y, z;
  band_decode(B, r, x, y, z);
  const int i = x - G.ibn[0], j = y - G.ibn[1], k = z - G.ibn[2];
  int dir[3];
  dir[0] = i < 0 ? -1 : (i >= G.ldims[0] ? 1 : 0);
  dir[1] = j < 0 ? -1 : (j >= G.ldims[1] ? 1 : 0);
  dir[2] = k < 0 ? -1 : (k >= G.ldims[2] ? 1 : 0);
  const int di = pm::dir2idx(dir);
  const long dst_off = fld_off(G, 0, i, j, k);
  const long src_off = fld_off(G, 0, i - dir[0] * G.ldims[0], j - dir[1] * G.ldims[1], k - dir[2] * G.ldims[2]);
  for (int pc = blockIdx.y; pc < G.n_patches * n_m; pc += gridDim.y) {
    const int p = pc / n_m, m = mb + (pc - p * n_m);
    const int slot = nei_slot[p * 27 + di];
    if (slot >= 0) {
      F[p * slot_len + m * G.fld_len + dst_off] = F[slot * slot_len + m * G.fld_len + src_off];
    }
  }
}

__global__ void k_add_ghosts(GridDev G, BandGeom B, float* __restrict__ F, long slot_len, int mb,
                             int me, const int* __restrict__ nei_slot,
                             const int8_t* __restrict__ add_order)
{
  // grid.x walks the boundary-band points of one (patch, component), grid.y strides over
  // the pairs (32-bit arithmetic, the point and its member set are decoded once)
  const int nb = B.n2 + B.n1 + B.n0;
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nb) {
    return;
  }
  const int n_m = me - mb;
  int i, j, k;
  band_decode(B, r, i, j, k);
  const int c[3] = {i, j, k};
  bool lo[3], hi[3];
#pragma unroll
  for (int d = 0; d < 3; d++) {
    lo[d] = c[d] < G.ibn[d];
    hi[d] = c[d] >= G.ldims[d] - G.ibn[d] && G.ibn[d] > 0;
  }
  // which of the 26 neighbour images fold onto this point (directions are compile-time
  // here: a few ANDs of the lo / hi flags each) ...
  unsigned members0 = 0;
#pragma unroll
  for (int di = 0; di < 27; di++) {
    if (di == 13) {
      continue;
    }
    const int d0 = di % 3 - 1, d1 = (di / 3) % 3 - 1, d2 = di / 9 - 1;
    const bool mem = (d0 == 0 || (d0 < 0 ? lo[0] : hi[0])) && (d1 == 0 || (d1 < 0 ? lo[1] : hi[1])) &&
                     (d2 == 0 || (d2 < 0 ? lo[2] : hi[2]));
    members0 |= (mem ? 1u : 0u) << di;
  }
  const long dst_off = fld_off(G, 0, i, j, k);
  for (int pc = blockIdx.y; pc < G.n_patches * n_m; pc += gridDim.y) {
    const int p = pc / n_m, m = mb + (pc - p * n_m);
    float* dst = F + p * slot_len + m * G.fld_len + dst_off;
    float acc = *dst;
    unsigned members = members0;
    // ... summed in the reference's order (mrc_ddc_multi.c:519-538); most points have one
    for (int o = 0; o < 26 && members; o++) {
      const int di = add_order[p * 26 + o];
      if (di < 0) {
        break;
      }
      if ((members >> di) & 1u) {
        members &= ~(1u << di);
        const int dir[3] = {di % 3 - 1, (di / 3) % 3 - 1, di / 9 - 1};
        const int slot = nei_slot[p * 27 + di];
        acc += F[slot * slot_len + fld_off(G, m, i - dir[0] * G.ldims[0], j - dir[1] * G.ldims[1],
                                            k - dir[2] * G.ldims[2])];
      }
    }
    *dst = acc;
  }
}

// ---------------------------------------------------------------- Yee

struct YeeDev
{
  float dth, cnx, cny, cnz;
  int inv[3];
};

#define FX(m, ii, jj, kk) Fp[fld_off(G, (m), Y.inv[0] ? 0 : (ii), (jj), (kk))]

template <bool IS_E>
__global__ void k_push_fields(GridDev G, YeeDev Y, float* __restrict__ F, long slot_len)
{
  // loop bounds grid.hxx:124-139 with (l, r) = (1, 2) for E, (2, 1) for H
  const int l = IS_E ? 1 : 2, r = IS_E ? 2 : 1;
  int e0 = Y.inv[0] ? 1 : G.ldims[0] + l + r;
  int e1 = G.ldims[1] + l + r, e2 = G.ldims[2] + l + r;
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)G.n_patches * e0 * e1 * e2) {
    return;
  }
  int i = (int)(idx % e0) - (Y.inv[0] ? 0 : l);
  idx /= e0;
  int j = (int)(idx % e1) - l;
  idx /= e1;
  int k = (int)(idx % e2) - l;
  int p = (int)(idx / e2);
  float* Fp = F + p * slot_len;
  if (IS_E) {
    FX(pm::EX, i, j, k) += (Y.cny * (FX(pm::HZ, i, j, k) - FX(pm::HZ, i, j - 1, k)) -
                            Y.cnz * (FX(pm::HY, i, j, k) - FX(pm::HY, i, j, k - 1)) -
                            Y.dth * FX(pm::JXI, i, j, k));
    FX(pm::EY, i, j, k) += (Y.cnz * (FX(pm::HX, i, j, k) - FX(pm::HX, i, j, k - 1)) -
                            Y.cnx * (FX(pm::HZ, i, j, k) - FX(pm::HZ, i - 1, j, k)) -
                            Y.dth * FX(pm::JYI, i, j, k));
    FX(pm::EZ, i, j, k) += (Y.cnx * (FX(pm::HY, i, j, k) - FX(pm::HY, i - 1, j, k)) -
                            Y.cny * (FX(pm::HX, i, j, k) - FX(pm::HX, i, j - 1, k)) -
                            Y.dth * FX(pm::JZI, i, j, k));
  } else {
    FX(pm::HX, i, j, k) -= (Y.cny * (FX(pm::EZ, i, j + 1, k) - FX(pm::EZ, i, j, k)) -
                            Y.cnz * (FX(pm::EY, i, j, k + 1) - FX(pm::EY, i, j, k)));
    FX(pm::HY, i, j, k) -= (Y.cnz * (FX(pm::EX, i, j, k + 1) - FX(pm::EX, i, j, k)) -
                            Y.cnx * (FX(pm::EZ, i + 1, j, k) - FX(pm::EZ, i, j, k)));
    FX(pm::HZ, i, j, k) -= (Y.cnx * (FX(pm::EY, i + 1, j, k) - FX(pm::EY, i, j, k)) -
                            Y.cny * (FX(pm::EX, i, j + 1, k) - FX(pm::EX, i, j, k)));
  }
}
#undef FX

// The same update, four x-neighbours per thread with 128-bit loads / stores (3D, rows a multiple
// of four floats: every row of PSC's layout then starts 16-byte aligned).  Per element the
// arithmetic is the expression above, so the result is bit-identical; the scalar kernel issues
// 16 loads per point and runs at 42 % of the HBM peak, this one 15 per four points.
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

template <bool IS_E>
__global__ void __launch_bounds__(256) k_push_fields_v4(GridDev G, YeeDev Y, float* __restrict__ F, long slot_len)
{
  const int l = IS_E ? 1 : 2, r = IS_E ? 2 : 1;
  const int nv = G.im[0] >> 2;
  const int e1 = G.ldims[1] + l + r, e2 = G.ldims[2] + l + r;
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)G.n_patches * nv * e1 * e2) {
    return;
  }
  const int v = (int)(idx % nv);
  idx /= nv;
  const int j = (int)(idx % e1) - l;
  idx /= e1;
  const int k = (int)(idx % e2) - l;
  const int p = (int)(idx / e2);
  float* Fp = F + p * slot_len;
  const int i0 = 4 * v - G.ibn[0]; // x index of the vector's first element
  const long sy = G.im[0], sz = (long)G.im[0] * G.im[1];
  auto at = [&](int m) { return Fp + fld_off(G, m, i0, j, k); };
  // element a of the row is updated for a in [1, im0) (E) / [0, im0 - 1) (H)  (grid.hxx:124-139)
  const int a0 = 4 * v;
  float4 o0, o1, o2;
  if (IS_E) {
    float* ex = at(pm::EX);
    float* ey = at(pm::EY);
    float* ez = at(pm::EZ);
    const float *hx = at(pm::HX), *hy = at(pm::HY), *hz = at(pm::HZ);
    const float4 HZc = ld4(hz), HZj = ld4(hz - sy), HYc = ld4(hy), HYk = ld4(hy - sz);
    const float4 HXc = ld4(hx), HXk = ld4(hx - sz), HXj = ld4(hx - sy);
    const float4 JX = ld4(at(pm::JXI)), JY = ld4(at(pm::JYI)), JZ = ld4(at(pm::JZI));
    const float hz_m = a0 > 0 ? hz[-1] : 0.f, hy_m = a0 > 0 ? hy[-1] : 0.f; // (a = 0 is not updated)
    const float4 HZi = make_float4(hz_m, HZc.x, HZc.y, HZc.z), HYi = make_float4(hy_m, HYc.x, HYc.y, HYc.z);
    o0 = ld4(ex), o1 = ld4(ey), o2 = ld4(ez);
#define UPD_E(c)                                                                                   \
  o0.c += (Y.cny * (HZc.c - HZj.c) - Y.cnz * (HYc.c - HYk.c) - Y.dth * JX.c);                       \
  o1.c += (Y.cnz * (HXc.c - HXk.c) - Y.cnx * (HZc.c - HZi.c) - Y.dth * JY.c);                       \
  o2.c += (Y.cnx * (HYc.c - HYi.c) - Y.cny * (HXc.c - HXj.c) - Y.dth * JZ.c);
    if (a0 > 0) {
      UPD_E(x)
    }
    UPD_E(y)
    UPD_E(z)
    UPD_E(w)
#undef UPD_E
    *reinterpret_cast<float4*>(ex) = o0;
    *reinterpret_cast<float4*>(ey) = o1;
    *reinterpret_cast<float4*>(ez) = o2;
  } else {
    float* hx = at(pm::HX);
    float* hy = at(pm::HY);
    float* hz = at(pm::HZ);
    const float *ex = at(pm::EX), *ey = at(pm::EY), *ez = at(pm::EZ);
    const float4 EZc = ld4(ez), EZj = ld4(ez + sy), EYc = ld4(ey), EYk = ld4(ey + sz);
    const float4 EXc = ld4(ex), EXk = ld4(ex + sz), EXj = ld4(ex + sy);
    const bool last = a0 + 4 >= G.im[0];
    const float ez_p = last ? 0.f : ez[4], ey_p = last ? 0.f : ey[4]; // (a = im0 - 1 is not updated)
    const float4 EZi = make_float4(EZc.y, EZc.z, EZc.w, ez_p), EYi = make_float4(EYc.y, EYc.z, EYc.w, ey_p);
    o0 = ld4(hx), o1 = ld4(hy), o2 = ld4(hz);
#define UPD_H(c)                                                                                   \
  o0.c -= (Y.cny * (EZj.c - EZc.c) - Y.cnz * (EYk.c - EYc.c));                                      \
  o1.c -= (Y.cnz * (EXk.c - EXc.c) - Y.cnx * (EZi.c - EZc.c));                                      \
  o2.c -= (Y.cnx * (EYi.c - EYc.c) - Y.cny * (EXj.c - EXc.c));
    UPD_H(x)
    UPD_H(y)
    UPD_H(z)
    if (!last) {
      UPD_H(w)
    }
#undef UPD_H
    *reinterpret_cast<float4*>(hx) = o0;
    *reinterpret_cast<float4*>(hy) = o1;
    *reinterpret_cast<float4*>(hz) = o2;
  }
}

// ---------------------------------------------------------------- conducting wall

enum
{
  CW_E,
  CW_H,
  CW_J
};

// one thread per (patch, transverse index t, x index) of the wall plane in dim d (1 or 2)
template <int OP>
__global__ void k_conducting_wall(GridDev G, float* __restrict__ F, long slot_len, int d, int hi,
                                  const pm::PatchBnd* __restrict__ pbs)
{
  const int x0 = G.ibn[0] ? -2 : 0, nx = G.ibn[0] ? G.ldims[0] + 4 : 1;
  const int dt = d == 1 ? 2 : 1; // transverse dim
  int t_lo = -2;
  if (OP == CW_H && d == 1 && !hi) {
    t_lo = -1; // psc_bnd_fields_impl.hxx:393 starts at -1
  }
  const int nt = G.ldims[dt] + 2 - t_lo;
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)G.n_patches * nt * nx) {
    return;
  }
  int ix = x0 + (int)(idx % nx);
  idx /= nx;
  int it = t_lo + (int)(idx % nt);
  int p = (int)(idx / nt);
  pm::PatchBnd pb = pbs[p];
  if (!(((hi ? pb.at_hi : pb.at_lo) >> d) & 1)) {
    return;
  }
  float* Fp = F + p * slot_len;
  const int M = G.ldims[d];
  // a(m, n): component m at wall-normal index n
#define A(m, n) Fp[d == 1 ? fld_off(G, (m), ix, (n), it) : fld_off(G, (m), ix, it, (n))]
  // tangential components of a wall in y are x,z; in z they are x,y
  const int T1 = 0, T2 = d == 1 ? 2 : 1, N = d == 1 ? 1 : 2;
  if (OP == CW_E) {
    const int E = pm::EX;
    if (!hi) {
      A(E + T1, 0) = 0.f;
      A(E + T1, -1) = A(E + T1, 1);
      A(E + T2, 0) = 0.f;
      A(E + T2, -1) = A(E + T2, 1);
      A(E + N, -1) = -A(E + N, 0);
    } else {
      A(E + T1, M) = 0.f;
      A(E + T1, M + 1) = A(E + T1, M - 1);
      A(E + T2, M) = 0.f;
      A(E + T2, M + 1) = A(E + T2, M - 1);
      A(E + N, M) = -A(E + N, M - 1);
    }
  } else if (OP == CW_H) {
    const int H = pm::HX;
    if (!hi) {
      A(H + T1, -1) = -A(H + T1, 0);
      A(H + T2, -1) = -A(H + T2, 0);
      A(H + N, -1) = A(H + N, 1);
    } else {
      A(H + T1, M) = -A(H + T1, M - 1);
      A(H + T2, M) = -A(H + T2, M - 1);
      A(H + N, M + 1) = A(H + N, M - 1);
    }
  } else {
    const int J = pm::JXI;
    if (!hi) {
      A(J + T1, 1) += A(J + T1, -1);
      A(J + T1, -1) = 0.f;
      A(J + T2, 1) += A(J + T2, -1);
      A(J + T2, -1) = 0.f;
      A(J + N, 0) -= A(J + N, -1);
      A(J + N, -1) = 0.f;
    } else {
      A(J + T1, M - 1) += A(J + T1, M + 1);
      A(J + T1, M + 1) = 0.f;
      A(J + T2, M - 1) += A(J + T2, M + 1);
      A(J + T2, M + 1) = 0.f;
      A(J + N, M - 1) -= A(J + N, M);
      A(J + N, M) = 0.f;
    }
  }
#undef A
}

// ---------------------------------------------------------------- moments, div, checks

__global__ void k_rho_1st_nc(GridDev G, uint32_t n, const uint32_t* __restrict__ off,
                             const float4* __restrict__ xi4, const float4* __restrict__ pxi4,
                             float* __restrict__ R, long slot_len, float fnqs,
                             const float* __restrict__ qk)
{
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) {
    return;
  }
  int p = patch_of(off, G.n_patches, i);
  float4 X = xi4[i];
  float qw = pxi4[i].w;
  float q = qk[__float_as_int(X.w)];
  // const_accessor_simple.hxx:60-63 w = qni_wni / q ; moment.hxx:77 val = w * q ; fnqs * val
  float w = qw / q;
  float value = fnqs * (w * q);
  float x[3] = {X.x, X.y, X.z};
  int l[3];
  float h[3];
#pragma unroll
  for (int d = 0; d < 3; d++) {
    float xn = x[d] * G.pc.dxi[d];
    l[d] = pm::fint(xn);
    h[d] = xn - (float)l[d];
  }
  float* Rp = R + p * slot_len;
  if (G.dim == pm::DIM_YZ) {
    atomicAdd(Rp + fld_off(G, 0, 0, l[1], l[2]), value * (1.f - h[1]) * (1.f - h[2]));
    atomicAdd(Rp + fld_off(G, 0, 0, l[1] + 1, l[2]), value * h[1] * (1.f - h[2]));
    atomicAdd(Rp + fld_off(G, 0, 0, l[1], l[2] + 1), value * (1.f - h[1]) * h[2]);
    atomicAdd(Rp + fld_off(G, 0, 0, l[1] + 1, l[2] + 1), value * h[1] * h[2]);
  } else {
#pragma unroll
    for (int c = 0; c < 8; c++) {
      int ox = c & 1, oy = (c >> 1) & 1, oz = c >> 2;
      float wgt = value * (ox ? h[0] : 1.f - h[0]) * (oy ? h[1] : 1.f - h[1]) *
                  (oz ? h[2] : 1.f - h[2]);
      atomicAdd(Rp + fld_off(G, 0, l[0] + ox, l[1] + oy, l[2] + oz), wgt);
    }
  }
}

// The same deposit for a cell-ordered store: one warp per cell sums its particles' eight
// (four) node weights in registers and adds them once per cell -- 64 x fewer global atomics at
// 64 particles per cell (S3D: 150 ms -> the time of reading the particles).  A particle whose
// node index is not the run's cell (1/float(dx) vs float(dx_inv) at a cell edge) adds on its own.
constexpr int RHO_WARPS = 8;
__global__ void __launch_bounds__(RHO_WARPS * 32)
  k_rho_1st_nc_cells(GridDev G, uint32_t nct, const uint32_t* __restrict__ cell_off,
                     const float4* __restrict__ xi4, const float4* __restrict__ pxi4, float* __restrict__ R,
                     long slot_len, float fnqs, const float* __restrict__ qk)
{
  const int lane = threadIdx.x & 31;
  const uint32_t g = blockIdx.x * RHO_WARPS + (threadIdx.x >> 5);
  if (g >= nct) {
    return;
  }
  const int p = g / G.n_cells;
  const int s = g - p * G.n_cells;
  const int c[3] = {s % G.ldims[0], (s / G.ldims[0]) % G.ldims[1], s / (G.ldims[0] * G.ldims[1])};
  const uint32_t begin = __ldg(&cell_off[g]), end = __ldg(&cell_off[g + 1]);
  if (begin == end) {
    return;
  }
  const bool yz = G.dim == pm::DIM_YZ;
  float* Rp = R + p * slot_len;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (uint32_t i = begin + lane; i < end; i += 32) {
    const float4 X = xi4[i];
    const float qw = pxi4[i].w;
    const float q = qk[__float_as_int(X.w)];
    const float value = fnqs * ((qw / q) * q); // (moment.hxx:77 val = w * q, w = qni_wni / q)
    const float x[3] = {X.x, X.y, X.z};
    int l[3];
    float h[3];
#pragma unroll
    for (int d = 0; d < 3; d++) {
      const float xn = x[d] * G.pc.dxi[d];
      l[d] = pm::fint(xn);
      h[d] = xn - (float)l[d];
    }
    const bool here = (yz || l[0] == c[0]) && l[1] == c[1] && l[2] == c[2];
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const int ox = k & 1, oy = (k >> 1) & 1, oz = k >> 2;
      if (yz && ox) {
        continue;
      }
      const float wgt = yz ? value * (oy ? h[1] : 1.f - h[1]) * (oz ? h[2] : 1.f - h[2])
                           : value * (ox ? h[0] : 1.f - h[0]) * (oy ? h[1] : 1.f - h[1]) * (oz ? h[2] : 1.f - h[2]);
      if (here) {
        acc[k] += wgt;
      } else {
        atomicAdd(Rp + fld_off(G, 0, yz ? 0 : l[0] + ox, l[1] + oy, l[2] + oz), wgt);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 8; k++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
    }
  }
  // lane k adds node k
  float mine = 0.f;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    mine = lane == k ? acc[k] : mine;
  }
  if (lane < 8 && !(yz && (lane & 1))) {
    const int ox = lane & 1, oy = (lane >> 1) & 1, oz = lane >> 2;
    atomicAdd(Rp + fld_off(G, 0, yz ? 0 : c[0] + ox, c[1] + oy, c[2] + oz), mine);
  }
}

// add_ghosts_reflecting.hxx:77-154, node-centred, one (d, hi) at a time
__global__ void k_reflect_nc(GridDev G, float* __restrict__ R, long slot_len, int d, int hi,
                             const pm::PatchBnd* __restrict__ pbs)
{
  int b[3], e[3];
#pragma unroll
  for (int a = 0; a < 3; a++) {
    int unused = G.ibn[a] ? 1 : 0;
    b[a] = -G.ibn[a] + unused;
    e[a] = G.ldims[a] + G.ibn[a];
  }
  {
    int unused = G.ibn[d] ? 1 : 0;
    if (!hi) {
      b[d] = 1;
      e[d] = 1 + G.ibn[d] - unused;
    } else {
      b[d] = G.ldims[d] - G.ibn[d] + unused;
      e[d] = G.ldims[d];
    }
  }
  int n0 = max(e[0] - b[0], 0), n1 = max(e[1] - b[1], 0), n2 = max(e[2] - b[2], 0);
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)G.n_patches * n0 * n1 * n2) {
    return;
  }
  int c[3];
  c[0] = b[0] + (int)(idx % n0);
  idx /= n0;
  c[1] = b[1] + (int)(idx % n1);
  idx /= n1;
  c[2] = b[2] + (int)(idx % n2);
  int p = (int)(idx / n2);
  pm::PatchBnd pb = pbs[p];
  bool at = ((hi ? pb.at_hi : pb.at_lo) >> d) & 1;
  int bc = hi ? pb.bc_hi[d] : pb.bc_lo[d];
  if (!at || bc != pm::BND_PRT_REFLECTING) {
    return;
  }
  int r[3] = {c[0], c[1], c[2]};
  r[d] = hi ? 2 * G.ldims[d] - c[d] : -c[d];
  float* Rp = R + p * slot_len;
  Rp[fld_off(G, 0, c[0], c[1], c[2])] += Rp[fld_off(G, 0, r[0], r[1], r[2])];
}

struct DxDev
{
  double dx[3];
  int inv[3];
};

// psc::item::div_nc, fields_item_fields.hxx:65-104 (interior points)
__device__ __forceinline__ float div_nc_at(const GridDev& G, const DxDev& D, const float* Fp, int m0,
                                           int i, int j, int k)
{
  float acc = 0.f;
#pragma unroll
  for (int a = 0; a < 3; a++) {
    if (D.inv[a]) {
      continue;
    }
    int c[3] = {i, j, k};
    c[a] -= 1;
    float diff = Fp[fld_off(G, m0 + a, i, j, k)] - Fp[fld_off(G, m0 + a, c[0], c[1], c[2])];
    acc = (float)((double)acc + (double)diff / D.dx[a]);
  }
  return acc;
}

__device__ __forceinline__ void decode_interior(const GridDev& G, size_t idx, int& p, int& i, int& j,
                                                int& k)
{
  i = (int)(idx % G.ldims[0]);
  idx /= G.ldims[0];
  j = (int)(idx % G.ldims[1]);
  idx /= G.ldims[1];
  k = (int)(idx % G.ldims[2]);
  p = (int)(idx / G.ldims[2]);
}

__device__ __forceinline__ void atomic_max_nonneg(double* addr, double v)
{
  atomicMax((unsigned long long*)addr, (unsigned long long)__double_as_longlong(v));
}

__global__ void k_div_nc(GridDev G, DxDev D, const float* __restrict__ F, long slot_len, int m0,
                         float* __restrict__ out, long out_slot_len)
{
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)G.n_patches * G.n_cells) {
    return;
  }
  int p, i, j, k;
  decode_interior(G, idx, p, i, j, k);
  out[p * out_slot_len + fld_off(G, 0, i, j, k)] = div_nc_at(G, D, F + p * slot_len, m0, i, j, k);
}

// checks_impl.hxx:60-97: max | rho_p - rho_m + dt * div J |
struct OpenLoDev
{
  int lo[3]; // BND_FLD_OPEN at the lower domain boundary in dim d
};

__global__ void k_continuity(GridDev G, DxDev D, OpenLoDev O, double dt, const float* __restrict__ F,
                             long slot_len, const float* __restrict__ rho_m,
                             const float* __restrict__ rho_p, const pm::PatchBnd* __restrict__ pbs,
                             double* __restrict__ err)
{
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  double v = 0.;
  if (idx < (size_t)G.n_patches * G.n_cells) {
    int p, i, j, k;
    decode_interior(G, idx, p, i, j, k);
    long o = p * G.fld_len + fld_off(G, 0, i, j, k);
    float d_rho = rho_p[o] - rho_m[o];
    float divj = div_nc_at(G, D, F + p * slot_len, pm::JXI, i, j, k);
    v = fabs((double)d_rho + dt * (double)divj);
    // checks_impl.hxx:78-90: particles enter / leave through a lower open boundary: div j is
    // DEFINED as -d rho / dt in the first cell layer there
    const int c3[3] = {i, j, k};
    const int at_lo = pbs[p].at_lo;
#pragma unroll
    for (int d = 0; d < 3; d++) {
      if (O.lo[d] && ((at_lo >> d) & 1) && c3[d] == 0) {
        v = 0.;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  }
  if ((threadIdx.x & 31) == 0 && v > 0.) {
    atomic_max_nonneg(err, v);
  }
}

struct WallDev
{
  int lo[3], hi[3]; // conducting wall / open at the lower / upper domain boundary in dim d
};

// checks_impl.hxx:157-184: max | div E - rho | (rho := div E on lower wall planes)
__global__ void k_gauss(GridDev G, DxDev D, WallDev W, const float* __restrict__ F, long slot_len,
                        const float* __restrict__ rho, const pm::PatchBnd* __restrict__ pbs,
                        double* __restrict__ err)
{
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  double v = 0.;
  if (idx < (size_t)G.n_patches * G.n_cells) {
    int p, i, j, k;
    decode_interior(G, idx, p, i, j, k);
    pm::PatchBnd pb = pbs[p];
    int c[3] = {i, j, k};
    bool skip = false;
#pragma unroll
    for (int d = 0; d < 3; d++) {
      skip = skip || (((pb.at_lo >> d) & 1) && c[d] == 0 && W.lo[d]);
    }
    if (!skip) {
      float dive = div_nc_at(G, D, F + p * slot_len, pm::EX, i, j, k);
      float r = dive - rho[p * G.fld_len + fld_off(G, 0, i, j, k)];
      v = fabs((double)r);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  }
  if ((threadIdx.x & 31) == 0 && v > 0.) {
    atomic_max_nonneg(err, v);
  }
}

// marder_impl.hxx:213-250: res = div E - rho on [0, ldims) -- and on the upper wall plane
// index ldims it is zeroed (it is a ghost here, the fill below overwrites or keeps 0)
__global__ void k_marder_res(GridDev G, DxDev D, WallDev W, const float* __restrict__ F,
                             long slot_len, const float* __restrict__ rho,
                             const pm::PatchBnd* __restrict__ pbs, float* __restrict__ res)
{
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)G.n_patches * G.n_cells) {
    return;
  }
  int p, i, j, k;
  decode_interior(G, idx, p, i, j, k);
  pm::PatchBnd pb = pbs[p];
  int c[3] = {i, j, k};
  bool zero = false;
#pragma unroll
  for (int d = 0; d < 3; d++) {
    zero = zero || (((pb.at_lo >> d) & 1) && c[d] == 0 && W.lo[d]);
  }
  long o = p * G.fld_len + fld_off(G, 0, i, j, k);
  res[o] = zero ? 0.f : div_nc_at(G, D, F + p * slot_len, pm::EX, i, j, k) - rho[o];
}

struct MarderDev
{
  float fac[3];
  int inv[3];
};

// psc::marder::correct, marder_impl.hxx:26-61
__global__ void k_marder_correct(GridDev G, MarderDev M, float* __restrict__ F, long slot_len,
                                 const float* __restrict__ res)
{
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)G.n_patches * G.n_cells) {
    return;
  }
  int p, i, j, k;
  decode_interior(G, idx, p, i, j, k);
  const float* R = res + p * G.fld_len;
  float r0 = R[fld_off(G, 0, i, j, k)];
  float* Fp = F + p * slot_len;
#pragma unroll
  for (int d = 0; d < 3; d++) {
    if (M.inv[d]) {
      continue;
    }
    int c[3] = {i, j, k};
    c[d] += 1;
    float* e = Fp + fld_off(G, pm::EX + d, i, j, k);
    *e = *e + (R[fld_off(G, 0, c[0], c[1], c[2])] - r0) * M.fac[d];
  }
}

// DiagEnergiesField.h:19-42
__global__ void k_field_energies(GridDev G, const float* __restrict__ F, long slot_len, double fac,
                                 double* __restrict__ out6)
{
  double s[6] = {0., 0., 0., 0., 0., 0.};
  size_t n = (size_t)G.n_patches * G.n_cells;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
       idx += (size_t)gridDim.x * blockDim.x) {
    int p, i, j, k;
    decode_interior(G, idx, p, i, j, k);
    const float* Fp = F + p * slot_len;
#pragma unroll
    for (int m = 0; m < 6; m++) {
      float v = Fp[fld_off(G, pm::EX + m, i, j, k)];
      s[m] += (double)(v * v) * fac;
    }
  }
#pragma unroll
  for (int m = 0; m < 6; m++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s[m] += __shfl_xor_sync(0xffffffffu, s[m], o);
    }
    if ((threadIdx.x & 31) == 0) {
      atomicAdd(&out6[m], s[m]);
    }
  }
}

DxDev make_dx(const Ctx* c)
{
  DxDev D;
  for (int d = 0; d < 3; d++) {
    D.dx[d] = c->g.dx[d];
    D.inv[d] = c->g.invar[d];
  }
  return D;
}

WallDev make_wall(const Ctx* c)
{
  WallDev W;
  for (int d = 0; d < 3; d++) {
    int lo = c->g.desc.bc_fld_lo[d], hi = c->g.desc.bc_fld_hi[d];
    W.lo[d] = lo == PSC_B200_BND_FLD_CONDUCTING_WALL || lo == PSC_B200_BND_FLD_OPEN;
    W.hi[d] = hi == PSC_B200_BND_FLD_CONDUCTING_WALL || hi == PSC_B200_BND_FLD_OPEN;
  }
  return W;
}

int check_field(Ctx* c, int id, int mb, int me)
{
  if (id < 0 || id >= (int)c->flds.size() || !c->flds[id].d) {
    return fail("invalid field id");
  }
  if (mb < 0 || me > c->flds[id].n_comps || mb > me) {
    return fail("component range out of bounds");
  }
  return 0;
}

} // namespace

// ---------------------------------------------------------------- container

int flds_create(Ctx* c, int n_comps, int* id)
{
  if (n_comps < 1) {
    return fail("n_comps must be positive");
  }
  FieldArr f;
  f.n_comps = n_comps;
  size_t bytes = (size_t)c->n_slots * n_comps * c->gd.fld_len * sizeof(float);
  PSC_CUDA_TRY(cudaMalloc(&f.d, bytes));
  PSC_CUDA_TRY(cudaMemsetAsync(f.d, 0, bytes, c->stream));
  c->flds.push_back(f);
  *id = (int)c->flds.size() - 1;
  return 0;
}

int flds_zero(Ctx* c, int id, int mb, int me)
{
  PSC_TRY(check_field(c, id, mb, me));
  if (me == mb) {
    return 0;
  }
  size_t pitch = (size_t)c->fld_slot_len(id) * sizeof(float);
  PSC_CUDA_TRY(cudaMemset2DAsync(c->fld(id) + (size_t)mb * c->gd.fld_len, pitch, 0,
                                 (size_t)(me - mb) * c->gd.fld_len * sizeof(float),
                                 c->gd.n_patches, c->stream));
  return 0;
}

int flds_fill(Ctx* c, int id, int m, float v)
{
  PSC_TRY(check_field(c, id, m, m + 1));
  size_t n = (size_t)c->n_slots * c->gd.fld_len;
  k_fill_value<<<div_up(n, 256), 256, 0, c->stream>>>(c->fld(id), c->fld_slot_len(id),
                                                     c->gd.fld_len, c->n_slots, m, v);
  c->n_launches++;
  return check_launch(c, "flds_fill");
}

int flds_upload(Ctx* c, int id, int mb, int me, const float* host, bool sync)
{
  PSC_TRY(check_field(c, id, mb, me));
  size_t w = (size_t)(me - mb) * c->gd.fld_len * sizeof(float);
  PSC_CUDA_TRY(cudaMemcpy2DAsync(c->fld(id) + (size_t)mb * c->gd.fld_len,
                                 (size_t)c->fld_slot_len(id) * sizeof(float), host, w, w,
                                 c->gd.n_patches, cudaMemcpyHostToDevice, c->stream));
  if (sync) {
    PSC_CUDA_TRY(cudaStreamSynchronize(c->stream));
  }
  return 0;
}

int flds_download(Ctx* c, int id, int mb, int me, float* host, bool sync)
{
  PSC_TRY(check_field(c, id, mb, me));
  size_t w = (size_t)(me - mb) * c->gd.fld_len * sizeof(float);
  PSC_CUDA_TRY(cudaMemcpy2DAsync(host, w, c->fld(id) + (size_t)mb * c->gd.fld_len,
                                 (size_t)c->fld_slot_len(id) * sizeof(float), w, c->gd.n_patches,
                                 cudaMemcpyDeviceToHost, c->stream));
  if (sync) {
    PSC_CUDA_TRY(cudaStreamSynchronize(c->stream));
  }
  return 0;
}

// ---------------------------------------------------------------- output hand-off
// What OutputFieldsItem does to its items between the operator that produces them and the
// writer (output_fields.hxx:170-231): take the interior, add it to the running sum, turn
// the sum into the mean.  All three stream once through HBM.

// y[p][ymb + m][r] += x[p][xmb + m][r]   (tfd_->gt() = tfd_->gt() + pfd, float + float)
template <typename V>
__global__ void __launch_bounds__(256)
  k_flds_add(V* __restrict__ y, const V* __restrict__ x, long y_slot, long x_slot, long row, int n_patches)
{
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)row * n_patches) {
    return;
  }
  const size_t p = i / row, r = i - p * row;
  V a = y[p * y_slot + r];
  const V b = x[p * x_slot + r];
  if constexpr (sizeof(V) == 16) {
    a.x += b.x, a.y += b.y, a.z += b.z, a.w += b.w;
  } else {
    a += b;
  }
  y[p * y_slot + r] = a;
}

// y = float(a * double(y))   ((1. / naccum_) * tfd_->gt(): a double scalar times float data)
template <typename V>
__global__ void __launch_bounds__(256) k_flds_scale(V* __restrict__ y, long y_slot, long row, int n_patches, double a)
{
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)row * n_patches) {
    return;
  }
  const size_t p = i / row, r = i - p * row;
  V v = y[p * y_slot + r];
  if constexpr (sizeof(V) == 16) {
    v.x = (float)(a * (double)v.x), v.y = (float)(a * (double)v.y);
    v.z = (float)(a * (double)v.z), v.w = (float)(a * (double)v.w);
  } else {
    v = (float)(a * (double)v);
  }
  y[p * y_slot + r] = v;
}

// psc::mflds::interior: out[p][m][k][j][i] over the patch's own cells
__global__ void __launch_bounds__(256)
  k_flds_pack_interior(GridDev G, const float* __restrict__ F, long slot_len, int mb, int n_m,
                       float* __restrict__ out, size_t n)
{
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) {
    return;
  }
  size_t r = idx;
  const int i = (int)(r % G.ldims[0]);
  r /= G.ldims[0];
  const int j = (int)(r % G.ldims[1]);
  r /= G.ldims[1];
  const int k = (int)(r % G.ldims[2]);
  r /= G.ldims[2];
  const int m = (int)(r % n_m);
  const size_t p = r / n_m;
  out[idx] = F[p * slot_len + fld_off(G, mb + m, i, j, k)];
}

int flds_add(Ctx* c, int y_id, int y_mb, int x_id, int x_mb, int n_comps)
{
  PSC_TRY(check_field(c, y_id, y_mb, y_mb + n_comps));
  PSC_TRY(check_field(c, x_id, x_mb, x_mb + n_comps));
  if (n_comps == 0) {
    return 0;
  }
  if (y_id == x_id && y_mb < x_mb + n_comps && x_mb < y_mb + n_comps) {
    return fail("mflds_add: source and destination components overlap");
  }
  KernelScope ks(c, "flds_add");
  const long fl = c->gd.fld_len, row = fl * n_comps;
  float* y = c->fld(y_id) + (size_t)y_mb * fl;
  const float* x = c->fld(x_id) + (size_t)x_mb * fl;
  const long ys = c->fld_slot_len(y_id), xs = c->fld_slot_len(x_id);
  if (fl % 4 == 0) { // every component of every slot starts on a 16-byte boundary
    k_flds_add<float4><<<div_up((size_t)(row / 4) * c->gd.n_patches, 256), 256, 0, c->stream>>>(
      reinterpret_cast<float4*>(y), reinterpret_cast<const float4*>(x), ys / 4, xs / 4, row / 4, c->gd.n_patches);
  } else {
    k_flds_add<float><<<div_up((size_t)row * c->gd.n_patches, 256), 256, 0, c->stream>>>(y, x, ys, xs, row,
                                                                                         c->gd.n_patches);
  }
  c->n_launches++;
  return check_launch(c, "flds_add");
}

int flds_scale(Ctx* c, int id, int mb, int me, double a)
{
  PSC_TRY(check_field(c, id, mb, me));
  if (me == mb) {
    return 0;
  }
  KernelScope ks(c, "flds_scale");
  const long fl = c->gd.fld_len, row = fl * (me - mb);
  float* y = c->fld(id) + (size_t)mb * fl;
  const long ys = c->fld_slot_len(id);
  if (fl % 4 == 0) {
    k_flds_scale<float4><<<div_up((size_t)(row / 4) * c->gd.n_patches, 256), 256, 0, c->stream>>>(
      reinterpret_cast<float4*>(y), ys / 4, row / 4, c->gd.n_patches, a);
  } else {
    k_flds_scale<float><<<div_up((size_t)row * c->gd.n_patches, 256), 256, 0, c->stream>>>(y, ys, row,
                                                                                           c->gd.n_patches, a);
  }
  c->n_launches++;
  return check_launch(c, "flds_scale");
}

int flds_download_interior(Ctx* c, int id, int mb, int me, float* host)
{
  PSC_TRY(check_field(c, id, mb, me));
  const size_t n = (size_t)c->gd.n_patches * (me - mb) * c->gd.n_cells;
  if (n == 0) {
    return 0;
  }
  PSC_TRY(c->scr[1].reserve(n * sizeof(float)));
  float* d = c->scr[1].as<float>();
  {
    KernelScope ks(c, "flds_pack_interior");
    k_flds_pack_interior<<<div_up(n, 256), 256, 0, c->stream>>>(c->gd, c->fld(id), c->fld_slot_len(id), mb,
                                                               me - mb, d, n);
    c->n_launches++;
  }
  PSC_TRY(check_launch(c, "flds_pack_interior"));
  PSC_CUDA_TRY(cudaMemcpyAsync(host, d, n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  PSC_CUDA_TRY(cudaStreamSynchronize(c->stream));
  return 0;
}

// ---------------------------------------------------------------- Bnd

int bnd_fill_ghosts(Ctx* c, int id, int mb, int me)
{
  PSC_TRY(check_field(c, id, mb, me));
  if (c->comm) {
    PSC_TRY(comm_halo_exchange(c, id, mb, me, false));
  }
  BandGeom B = make_band(c->g, true);
  size_t n = (size_t)c->gd.n_patches * (me - mb) * (B.n2 + B.n1 + B.n0);
  if (n == 0) {
    return 0;
  }
  KernelScope ks(c, "fill_ghosts");
  {
    const int nb = B.n2 + B.n1 + B.n0;
    dim3 grid(div_up(nb, 128), (unsigned)std::min(c->gd.n_patches * (me - mb), 32768));
    k_fill_ghosts<<<grid, 128, 0, c->stream>>>(c->gd, B, c->fld(id), c->fld_slot_len(id), mb, me, c->d_nei_slot);
  }
  c->n_launches++;
  return check_launch(c, "fill_ghosts");
}

int bnd_add_ghosts(Ctx* c, int id, int mb, int me)
{
  PSC_TRY(check_field(c, id, mb, me));
  if (c->comm) {
    PSC_TRY(comm_halo_exchange(c, id, mb, me, true));
  }
  BandGeom B = make_band(c->g, false);
  size_t n = (size_t)c->gd.n_patches * (me - mb) * (B.n2 + B.n1 + B.n0);
  if (n == 0) {
    return 0;
  }
  KernelScope ks(c, "add_ghosts");
  {
    const int nb = B.n2 + B.n1 + B.n0;
    dim3 grid(div_up(nb, 128), (unsigned)std::min(c->gd.n_patches * (me - mb), 32768));
    k_add_ghosts<<<grid, 128, 0, c->stream>>>(c->gd, B, c->fld(id), c->fld_slot_len(id), mb, me, c->d_nei_slot,
                                              c->d_add_order);
  }
  c->n_launches++;
  return check_launch(c, "add_ghosts");
}

// ---------------------------------------------------------------- BndFields

// ---- BND_FLD_OPEN (psc_bnd_fields_impl.hxx:210-300 set_lower/upper_ghosts with include_edge =
// false; :535-640 radiative_H_lo/hi without an incoming pulse; background_e = background_h = 0).
// One thread per point of the patch array; `d` is the wall's direction.
struct OpenDev
{
  float dt, dtdx[3];
};

// E: every ghost of the three E components behind the wall := background (0); at the upper wall
// the normal component is a ghost already in the edge plane
__global__ void k_open_E(GridDev G, float* __restrict__ F, long slot_len, int d, int hi,
                         const pm::PatchBnd* __restrict__ pbs)
{
  const size_t per = (size_t)G.fld_len;
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= per * G.n_patches) {
    return;
  }
  const int p = (int)(idx / per);
  size_t r = idx - (size_t)p * per;
  const int c3[3] = {(int)(r % G.im[0]) - G.ibn[0], (int)((r / G.im[0]) % G.im[1]) - G.ibn[1],
                     (int)(r / ((size_t)G.im[0] * G.im[1])) - G.ibn[2]};
  const pm::PatchBnd pb = pbs[p];
  if (!(((hi ? pb.at_hi : pb.at_lo) >> d) & 1)) {
    return;
  }
  float* Fp = F + p * slot_len;
  const bool ghost = hi ? c3[d] >= G.ldims[d] + 1 : c3[d] < 0;
  const bool edge = hi && c3[d] == G.ldims[d];
#pragma unroll
  for (int m = 0; m < 3; m++) {
    if (ghost || (edge && m == d)) {
      Fp[fld_off(G, pm::EX + m, c3[0], c3[1], c3[2])] = 0.f;
    }
  }
}

// H: first-order absorbing condition for the two tangential components in the first ghost plane
__global__ void k_open_H(GridDev G, OpenDev O, float* __restrict__ F, long slot_len, int d, int hi,
                         const pm::PatchBnd* __restrict__ pbs)
{
  const int d0 = d, d1 = (d + 1) % 3, d2 = (d + 2) % 3;
  const size_t plane = (size_t)G.im[d1] * G.im[d2];
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= plane * G.n_patches) {
    return;
  }
  const int p = (int)(idx / plane);
  size_t r = idx - (size_t)p * plane;
  const pm::PatchBnd pb = pbs[p];
  if (!(((hi ? pb.at_hi : pb.at_lo) >> d) & 1)) {
    return;
  }
  int i3[3], e[3];
  i3[d0] = hi ? G.ldims[d0] : -1;
  i3[d1] = (int)(r % G.im[d1]) - G.ibn[d1];
  i3[d2] = (int)(r / G.im[d1]) - G.ibn[d2];
  e[0] = i3[0], e[1] = i3[1], e[2] = i3[2];
  e[d0] += hi ? -1 : 1; // edge_idx
  float* Fp = F + p * slot_len;
  auto at = [&](int m, const int* q) -> float& { return Fp[fld_off(G, m, q[0], q[1], q[2])]; };
  const int H0 = pm::HX + d0, H1 = pm::HX + d1, H2 = pm::HX + d2;
  const int E1 = pm::EX + d1, E2 = pm::EX + d2, J1 = pm::JXI + d1, J2 = pm::JXI + d2;
  // the H0 differences are taken at edge_idx (lo) / i3 (hi); one index below the array (the
  // reference reads out of bounds there) the difference is taken as zero
  const int* q = hi ? i3 : e;
  int qm[3] = {q[0], q[1], q[2]};
  float dH0_2 = 0.f, dH0_1 = 0.f;
  if (G.im[d2] > 1 && q[d2] - 1 >= -G.ibn[d2]) {
    qm[d2] -= 1;
    dH0_2 = at(H0, q) - at(H0, qm);
    qm[d2] += 1;
  }
  if (G.im[d1] > 1 && q[d1] - 1 >= -G.ibn[d1]) {
    qm[d1] -= 1;
    dH0_1 = at(H0, q) - at(H0, qm);
  }
  const float dt = O.dt;
  if (!hi) {
    at(H2, i3) = (4.f * 0.f - 2.f * (at(E1, e) - 0.f) - O.dtdx[d2] * dH0_2 - (1.f - O.dtdx[d0]) * (at(H2, e) - 0.f) +
                  dt * at(J1, e)) /
                   (1.f + O.dtdx[d0]) +
                 0.f;
    at(H1, i3) = (-4.f * 0.f + 2.f * (at(E2, e) - 0.f) - O.dtdx[d1] * dH0_1 - (1.f - O.dtdx[d0]) * (at(H1, e) - 0.f) +
                  dt * at(J2, e)) /
                   (1.f + O.dtdx[d0]) +
                 0.f;
  } else {
    at(H2, i3) = (-4.f * 0.f + 2.f * (at(E1, i3) - 0.f) + O.dtdx[d2] * dH0_2 - (1.f - O.dtdx[d0]) * (at(H2, e) - 0.f) -
                  dt * at(J1, i3)) /
                   (1.f + O.dtdx[d0]) +
                 0.f;
    at(H1, i3) = (4.f * 0.f - 2.f * (at(E2, i3) - 0.f) + O.dtdx[d1] * dH0_1 - (1.f - O.dtdx[d0]) * (at(H1, e) - 0.f) -
                  dt * at(J2, i3)) /
                   (1.f + O.dtdx[d0]) +
                 0.f;
  }
}

template <int OP>
static int bndf_open(Ctx* c, int d, int hi, const char* name)
{
  const GridHost& g = c->g;
  if (g.invar[d] || OP == CW_J) {
    return 0; // (add_ghosts_J: BND_FLD_OPEN does nothing, psc_bnd_fields_impl.hxx:169-171)
  }
  KernelScope ks(c, name);
  if (OP == CW_E) {
    const size_t n = (size_t)c->gd.fld_len * c->gd.n_patches;
    k_open_E<<<div_up(n, 256), 256, 0, c->stream>>>(c->gd, c->fld(0), c->fld_slot_len(0), d, hi, c->d_patch_bnd);
  } else {
    OpenDev O;
    O.dt = (float)g.desc.dt;
    for (int a = 0; a < 3; a++) {
      O.dtdx[a] = O.dt * (float)g.dx_inv[a];
    }
    const size_t n = (size_t)c->gd.im[(d + 1) % 3] * c->gd.im[(d + 2) % 3] * c->gd.n_patches;
    k_open_H<<<div_up(n, 128), 128, 0, c->stream>>>(c->gd, O, c->fld(0), c->fld_slot_len(0), d, hi, c->d_patch_bnd);
  }
  c->n_launches++;
  return 0;
}

template <int OP>
static int bndf_apply(Ctx* c, const char* name)
{
  const GridHost& g = c->g;
  // psc_bnd_fields_impl.hxx:27-188: lower walls for d = 0..2, then upper walls
  for (int hi = 0; hi < 2; hi++) {
    for (int d = 0; d < 3; d++) {
      int bc = hi ? g.desc.bc_fld_hi[d] : g.desc.bc_fld_lo[d];
      if (bc == PSC_B200_BND_FLD_OPEN) {
        PSC_TRY(bndf_open<OP>(c, d, hi, name));
        continue;
      }
      if (bc != PSC_B200_BND_FLD_CONDUCTING_WALL) {
        continue;
      }
      if (d == 0) {
        return fail("conducting wall in x is not implemented (nor is it in PSC: "
                    "psc_bnd_fields_impl.hxx:301-530 covers y and z)");
      }
      int dt = d == 1 ? 2 : 1;
      size_t n = (size_t)c->gd.n_patches * (g.ldims[dt] + 4) * (g.ibn[0] ? g.ldims[0] + 4 : 1);
      KernelScope ks(c, name);
      k_conducting_wall<OP><<<div_up(n, 128), 128, 0, c->stream>>>(c->gd, c->fld(0),
                                                                  c->fld_slot_len(0), d, hi,
                                                                  c->d_patch_bnd);
      c->n_launches++;
    }
  }
  return check_launch(c, name);
}

int bndf_fill_ghosts_E(Ctx* c)
{
  return bndf_apply<CW_E>(c, "bndf_E");
}
int bndf_fill_ghosts_H(Ctx* c)
{
  return bndf_apply<CW_H>(c, "bndf_H");
}
int bndf_add_ghosts_J(Ctx* c)
{
  return bndf_apply<CW_J>(c, "bndf_J");
}

// ---------------------------------------------------------------- PushFields

static int push_fields(Ctx* c, double dt_fac, bool is_E)
{
  YeeConst y = make_yee_const(c->g, dt_fac);
  YeeDev Y{y.dth, y.cnx, y.cny, y.cnz, {c->g.invar[0], c->g.invar[1], c->g.invar[2]}};
  const GridDev& G = c->gd;
  size_t n = (size_t)G.n_patches * (G.ibn[0] ? G.ldims[0] + 3 : 1) * (G.ldims[1] + 3) *
             (G.ldims[2] + 3);
  KernelScope ks(c, is_E ? "push_E" : "push_H");
  if (G.dim == pm::DIM_XYZ && G.im[0] % 4 == 0 && G.ibn[0] == 2 && c->fld_slot_len(0) % 4 == 0 && c->opt_vec_fields) {
    const size_t nv = (size_t)G.n_patches * (G.im[0] / 4) * (G.ldims[1] + 3) * (G.ldims[2] + 3);
    if (is_E) {
      k_push_fields_v4<true><<<div_up(nv, 256), 256, 0, c->stream>>>(G, Y, c->fld(0), c->fld_slot_len(0));
    } else {
      k_push_fields_v4<false><<<div_up(nv, 256), 256, 0, c->stream>>>(G, Y, c->fld(0), c->fld_slot_len(0));
    }
    c->n_launches++;
    return check_launch(c, "push_fields");
  }
  if (is_E) {
    k_push_fields<true><<<div_up(n, 256), 256, 0, c->stream>>>(G, Y, c->fld(0), c->fld_slot_len(0));
  } else {
    k_push_fields<false><<<div_up(n, 256), 256, 0, c->stream>>>(G, Y, c->fld(0), c->fld_slot_len(0));
  }
  c->n_launches++;
  return check_launch(c, "push_fields");
}

int push_E(Ctx* c, double dt_fac)
{
  return push_fields(c, dt_fac, true);
}
int push_H(Ctx* c, double dt_fac)
{
  return push_fields(c, dt_fac, false);
}

// ---------------------------------------------------------------- moments / checks / Marder

namespace
{

// ---- the 1st-order moment family (psc/moment.hxx:119-311, psc/deposit.hxx:24-65,
// 172-212, 262-285): one thread per particle, up to 13 values per particle deposited
// with the 1st-order weights to the cell centres (n, v, p, T, all) or the nodes (rho)
struct MomentPrm
{
  int which; // PSC_B200_MOMENT_*
  float fnqs;
  float q[pm::MAX_KINDS], m[pm::MAX_KINDS];
};

__global__ void __launch_bounds__(256)
  k_moment_1st(GridDev G, MomentPrm M, uint32_t n, const uint32_t* __restrict__ off,
               const float4* __restrict__ xi4, const float4* __restrict__ pxi4, float* __restrict__ R,
               long slot_len)
{
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) {
    return;
  }
  const int p = patch_of(off, G.n_patches, i);
  const float4 X = xi4[i], U = pxi4[i];
  const int kind = __float_as_int(X.w);
  const float q = M.q[kind], ms = M.m[kind];
  const float w = U.w / q; // const_accessor_simple.hxx:60-63
  const float u[3] = {U.x, U.y, U.z};
  const float x[3] = {X.x, X.y, X.z};
  const bool cc = M.which != PSC_B200_MOMENT_RHO_NC;
  int l[3];
  float h[3];
#pragma unroll
  for (int d = 0; d < 3; d++) {
    const float xn = x[d] * G.pc.dxi[d];
    if (cc) {
      l[d] = pm::fint(xn - .5f);
      h[d] = xn - .5f - (float)l[d];
    } else {
      l[d] = pm::fint(xn);
      h[d] = xn - (float)l[d];
    }
  }
  float vxi[3];
  {
    const float root = pm::rsqrt_ref(1.f + u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
#pragma unroll
    for (int d = 0; d < 3; d++) {
      vxi[d] = u[d] * root;
    }
  }
  int m0 = 0, nv = 0;
  float val[13];
  switch (M.which) {
    case PSC_B200_MOMENT_N: m0 = kind, val[0] = w, nv = 1; break;
    case PSC_B200_MOMENT_RHO_NC: m0 = 0, val[0] = w * q, nv = 1; break;
    case PSC_B200_MOMENT_V:
      m0 = 3 * kind, nv = 3;
      for (int d = 0; d < 3; d++) {
        val[d] = w * vxi[d];
      }
      break;
    case PSC_B200_MOMENT_P:
      m0 = 3 * kind, nv = 3;
      for (int d = 0; d < 3; d++) {
        val[d] = w * ms * u[d];
      }
      break;
    case PSC_B200_MOMENT_T: {
      const int a[6] = {0, 1, 2, 0, 0, 1}, b[6] = {0, 1, 2, 1, 2, 2};
      m0 = 6 * kind, nv = 6;
      for (int k = 0; k < 6; k++) {
        val[k] = w * ms * u[a[k]] * vxi[b[k]];
      }
      break;
    }
    default: { // PSC_B200_MOMENT_ALL
      const int a[6] = {0, 1, 2, 0, 1, 2}, b[6] = {0, 1, 2, 1, 2, 0};
      m0 = 13 * kind, nv = 13;
      val[0] = w * q;
      for (int d = 0; d < 3; d++) {
        val[1 + d] = w * q * vxi[d];
        val[4 + d] = w * ms * u[d];
      }
      for (int k = 0; k < 6; k++) {
        val[7 + k] = w * ms * u[a[k]] * vxi[b[k]];
      }
    }
  }
  float* Rp = R + p * slot_len;
  const bool yz = G.dim == pm::DIM_YZ;
  // corner weights in the reference's association: value * wx * wy * wz
  for (int k = 0; k < nv; k++) {
    const float value = M.fnqs * val[k];
    for (int c = yz ? 0 : 0; c < 8; c++) {
      const int ox = c & 1, oy = (c >> 1) & 1, oz = c >> 2;
      if (yz && ox) {
        continue;
      }
      float wgt = value;
      if (!yz) {
        wgt = wgt * (ox ? h[0] : 1.f - h[0]);
      }
      wgt = wgt * (oy ? h[1] : 1.f - h[1]) * (oz ? h[2] : 1.f - h[2]);
      atomicAdd(Rp + fld_off(G, m0 + k, yz ? 0 : l[0] + ox, l[1] + oy, l[2] + oz), wgt);
    }
  }
}

// The cell-centred moments of a cell-ordered store: one warp per cell.  A particle of cell c
// deposits to the cells l .. l + 1 with l in {c - 1, c} per direction (which half of the cell
// it sits in), i.e. to the 27 cells around c with three weights per direction, one of them 0:
//   l = c - 1: {1 - h, h, 0}      l = c: {0, 1 - h, h}
// -- the reference's factors, multiplied in its order (value * wx * wy * wz).  For one component
// at a time every lane sums its particles' 27 contributions in registers, a transposing
// butterfly leaves the warp total of target t on lane t, and lane t adds it once: 27 global
// atomics per cell and component instead of 8 per particle and value (S3D, Moments_1st:
// 648 ms before).  A particle whose l is not c - 1 or c (1/float(dx) against float(dx_inv) at
// an edge) adds on its own.
__device__ __forceinline__ void warp_transpose_reduce32(float (&v)[32], int lane)
{
#pragma unroll
  for (int h = 16; h >= 1; h >>= 1) {
    const bool up = lane & h;
#pragma unroll
    for (int j = 0; j < h; j++) {
      const float send = up ? v[j] : v[j + h];
      const float keep = up ? v[j + h] : v[j];
      v[j] = keep + __shfl_xor_sync(0xffffffffu, send, h);
    }
  }
}

constexpr int MOM_WARPS = 4;
// (WHICH at compile time: Moment_n needs neither the momenta nor the 1/gamma of the others)
template <int WHICH>
__global__ void __launch_bounds__(MOM_WARPS * 32)
  k_moment_1st_cells(GridDev G, MomentPrm M, int n_kinds, uint32_t nct, const uint32_t* __restrict__ cell_off,
                     const float4* __restrict__ xi4, const float4* __restrict__ pxi4, float* __restrict__ R,
                     long slot_len)
{
  const int lane = threadIdx.x & 31;
  const uint32_t g = blockIdx.x * MOM_WARPS + (threadIdx.x >> 5);
  if (g >= nct) {
    return;
  }
  const int p = g / G.n_cells;
  const int s = g - p * G.n_cells;
  const int c[3] = {s % G.ldims[0], (s / G.ldims[0]) % G.ldims[1], s / (G.ldims[0] * G.ldims[1])};
  const uint32_t begin = __ldg(&cell_off[g]), end = __ldg(&cell_off[g + 1]);
  if (begin == end) {
    return;
  }
  const bool yz = G.dim == pm::DIM_YZ;
  float* Rp = R + p * slot_len;
  constexpr int nv = WHICH == PSC_B200_MOMENT_N ? 1 : (WHICH == PSC_B200_MOMENT_V || WHICH == PSC_B200_MOMENT_P) ? 3
                     : WHICH == PSC_B200_MOMENT_T ? 6 : 13;
  for (int comp = 0; comp < nv * n_kinds; comp++) {
    const int ck = comp / nv, kk = comp - ck * nv; // kind, value of that kind
    float acc[32];
#pragma unroll
    for (int t = 0; t < 32; t++) {
      acc[t] = 0.f;
    }
    for (uint32_t i = begin + lane; i < end; i += 32) {
      const float4 X = xi4[i];
      const int kind = __float_as_int(X.w);
      if (kind != ck) {
        continue;
      }
      const float4 U = pxi4[i];
      const float q = M.q[kind], ms = M.m[kind];
      const float w = U.w / q; // const_accessor_simple.hxx:60-63
      const float u[3] = {U.x, U.y, U.z};
      const float x[3] = {X.x, X.y, X.z};
      int l[3];
      float h[3];
      bool near = true;
#pragma unroll
      for (int d = 0; d < 3; d++) {
        const float xn = x[d] * G.pc.dxi[d];
        l[d] = pm::fint(xn - .5f);
        h[d] = xn - .5f - (float)l[d];
        near = near && ((yz && d == 0) || l[d] == c[d] - 1 || l[d] == c[d]);
      }
      float vxi[3];
      {
        const float root = pm::rsqrt_ref(1.f + u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
#pragma unroll
        for (int d = 0; d < 3; d++) {
          vxi[d] = u[d] * root;
        }
      }
      float val;
      switch (WHICH) {
        case PSC_B200_MOMENT_N: val = w; break;
        case PSC_B200_MOMENT_V: val = w * (kk == 0 ? vxi[0] : kk == 1 ? vxi[1] : vxi[2]); break;
        case PSC_B200_MOMENT_P: val = w * ms * (kk == 0 ? u[0] : kk == 1 ? u[1] : u[2]); break;
        case PSC_B200_MOMENT_T: {
          // (a, b) = (0,0) (1,1) (2,2) (0,1) (0,2) (1,2)
          const float ua = kk == 0 || kk == 3 || kk == 4 ? u[0] : (kk == 1 || kk == 5 ? u[1] : u[2]);
          const float vb = kk == 0 ? vxi[0] : (kk == 1 || kk == 3 ? vxi[1] : vxi[2]);
          val = w * ms * ua * vb;
          break;
        }
        default: { // PSC_B200_MOMENT_ALL: rho, j (3), p (3), t: (0,0) (1,1) (2,2) (0,1) (1,2) (2,0)
          if (kk == 0) {
            val = w * q;
          } else if (kk < 4) {
            val = w * q * (kk == 1 ? vxi[0] : kk == 2 ? vxi[1] : vxi[2]);
          } else if (kk < 7) {
            val = w * ms * (kk == 4 ? u[0] : kk == 5 ? u[1] : u[2]);
          } else {
            const int k6 = kk - 7;
            const float ua = k6 == 0 || k6 == 3 ? u[0] : (k6 == 1 || k6 == 4 ? u[1] : u[2]);
            const float vb = k6 == 0 || k6 == 5 ? vxi[0] : (k6 == 1 || k6 == 3 ? vxi[1] : vxi[2]);
            val = w * ms * ua * vb;
          }
        }
      }
      const float value = M.fnqs * val;
      if (!near) {
        for (int cn = 0; cn < 8; cn++) {
          const int ox = cn & 1, oy = (cn >> 1) & 1, oz = cn >> 2;
          if (yz && ox) {
            continue;
          }
          float wgt = value;
          if (!yz) {
            wgt = wgt * (ox ? h[0] : 1.f - h[0]);
          }
          wgt = wgt * (oy ? h[1] : 1.f - h[1]) * (oz ? h[2] : 1.f - h[2]);
          atomicAdd(Rp + fld_off(G, comp, yz ? 0 : l[0] + ox, l[1] + oy, l[2] + oz), wgt);
        }
        continue;
      }
      // three weights per direction (the factor of a target the particle does not reach is never used)
      float w3[3][3];
      bool hit[3][3];
#pragma unroll
      for (int d = 0; d < 3; d++) {
        const bool low = l[d] == c[d] - 1;
        w3[d][0] = 1.f - h[d];
        w3[d][1] = low ? h[d] : 1.f - h[d];
        w3[d][2] = h[d];
        hit[d][0] = low, hit[d][1] = true, hit[d][2] = !low;
      }
#pragma unroll
      for (int t = 0; t < 27; t++) {
        const int tx = t % 3, ty = (t / 3) % 3, tz = t / 9;
        if (yz && tx != 1) {
          continue;
        }
        const bool on = (yz || hit[0][tx]) && hit[1][ty] && hit[2][tz];
        float wgt = value;
        if (!yz) {
          wgt = wgt * w3[0][tx];
        }
        wgt = wgt * w3[1][ty] * w3[2][tz];
        acc[t] += on ? wgt : 0.f;
      }
    }
    warp_transpose_reduce32(acc, lane);
    if (lane < 27 && acc[0] != 0.f) {
      const int tx = lane % 3, ty = (lane / 3) % 3, tz = lane / 9;
      atomicAdd(Rp + fld_off(G, comp, yz ? 0 : c[0] + tx - 1, c[1] + ty - 1, c[2] + tz - 1), acc[0]);
    }
  }
}

// add_ghosts_reflecting.hxx:7-70, cell-centred, one (d, hi) at a time, all components
__global__ void k_reflect_cc(GridDev G, float* __restrict__ R, long slot_len, int n_comps, int d, int hi,
                             const pm::PatchBnd* __restrict__ pbs)
{
  int b[3], e[3];
#pragma unroll
  for (int a = 0; a < 3; a++) {
    b[a] = -G.ibn[a];
    e[a] = G.ldims[a] + G.ibn[a];
  }
  if (!hi) {
    b[d] = 0;
    e[d] = G.ibn[d];
  } else {
    b[d] = G.ldims[d] - G.ibn[d];
    e[d] = G.ldims[d];
  }
  const int n0 = max(e[0] - b[0], 0), n1 = max(e[1] - b[1], 0), n2 = max(e[2] - b[2], 0);
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)G.n_patches * n_comps * n0 * n1 * n2) {
    return;
  }
  int c[3];
  c[0] = b[0] + (int)(idx % n0);
  idx /= n0;
  c[1] = b[1] + (int)(idx % n1);
  idx /= n1;
  c[2] = b[2] + (int)(idx % n2);
  idx /= n2;
  const int m = (int)(idx % n_comps);
  const int p = (int)(idx / n_comps);
  const pm::PatchBnd pb = pbs[p];
  const bool at = ((hi ? pb.at_hi : pb.at_lo) >> d) & 1;
  const int bc = hi ? pb.bc_hi[d] : pb.bc_lo[d];
  if (!at || bc != pm::BND_PRT_REFLECTING) {
    return;
  }
  int r[3] = {c[0], c[1], c[2]};
  r[d] = hi ? 2 * G.ldims[d] - c[d] - 1 : -c[d] - 1;
  float* Rp = R + p * slot_len;
  Rp[fld_off(G, m, c[0], c[1], c[2])] += Rp[fld_off(G, m, r[0], r[1], r[2])];
}

} // namespace

int moment_n_comps(const Ctx* c, int which)
{
  const int nk = c->g.desc.n_kinds;
  switch (which) {
    case PSC_B200_MOMENT_N: return nk;
    case PSC_B200_MOMENT_V: return 3 * nk;
    case PSC_B200_MOMENT_P: return 3 * nk;
    case PSC_B200_MOMENT_T: return 6 * nk;
    case PSC_B200_MOMENT_ALL: return 13 * nk;
    case PSC_B200_MOMENT_RHO_NC: return 1;
  }
  return -1;
}

// ItemMoment::operator() (fields_item.hxx:117-129): zeros, moment, reflecting folds, add_ghosts
int moment_1st(Ctx* c, int id, int which)
{
  const int nc = moment_n_comps(c, which);
  if (nc < 0) {
    return fail("moment_1st: unknown moment");
  }
  PSC_TRY(check_field(c, id, 0, nc));
  if (c->flds[id].n_comps != nc) {
    return fail("moment_1st: the field must have exactly the moment's components");
  }
  const GridDev& G = c->gd;
  // local patches and proxies (they receive neighbours' ghost sums)
  PSC_CUDA_TRY(cudaMemsetAsync(c->fld(id), 0, (size_t)c->n_slots * c->fld_slot_len(id) * sizeof(float),
                               c->stream));
  MomentPrm M{};
  M.which = which;
  M.fnqs = (float)c->g.desc.fnqs;
  for (int k = 0; k < c->g.desc.n_kinds; k++) {
    M.q[k] = (float)c->g.desc.q[k];
    M.m[k] = (float)c->g.desc.m[k];
  }
  if (c->n_prts) {
    KernelScope ks(c, "moment_1st");
    if (c->sorted && c->opt_cell_moments && which != PSC_B200_MOMENT_RHO_NC) {
      const uint32_t nct = (uint32_t)G.n_cells * G.n_patches;
#define PSC_MOM_CELLS(W)                                                                           \
  k_moment_1st_cells<W><<<div_up(nct, MOM_WARPS), MOM_WARPS * 32, 0, c->stream>>>(                 \
    G, M, c->g.desc.n_kinds, nct, c->d_cell_off, c->xi(), c->pxi(), c->fld(id), c->fld_slot_len(id))
      switch (which) {
        case PSC_B200_MOMENT_N: PSC_MOM_CELLS(PSC_B200_MOMENT_N); break;
        case PSC_B200_MOMENT_V: PSC_MOM_CELLS(PSC_B200_MOMENT_V); break;
        case PSC_B200_MOMENT_P: PSC_MOM_CELLS(PSC_B200_MOMENT_P); break;
        case PSC_B200_MOMENT_T: PSC_MOM_CELLS(PSC_B200_MOMENT_T); break;
        default: PSC_MOM_CELLS(PSC_B200_MOMENT_ALL); break;
      }
#undef PSC_MOM_CELLS
    } else {
      k_moment_1st<<<div_up(c->n_prts, 256), 256, 0, c->stream>>>(G, M, c->n_prts, c->d_off, c->xi(), c->pxi(),
                                                                  c->fld(id), c->fld_slot_len(id));
    }
    c->n_launches++;
  }
  for (int hi = 0; hi < 2; hi++) {
    for (int d = 0; d < 3; d++) {
      const int bc = hi ? c->g.desc.bc_prt_hi[d] : c->g.desc.bc_prt_lo[d];
      if (bc != PSC_B200_BND_PRT_REFLECTING || c->g.invar[d]) {
        continue;
      }
      const size_t n = (size_t)G.n_patches * nc * G.fld_len;
      if (which == PSC_B200_MOMENT_RHO_NC) {
        k_reflect_nc<<<div_up(n, 256), 256, 0, c->stream>>>(G, c->fld(id), c->fld_slot_len(id), d, hi,
                                                           c->d_patch_bnd);
      } else {
        k_reflect_cc<<<div_up(n, 256), 256, 0, c->stream>>>(G, c->fld(id), c->fld_slot_len(id), nc, d, hi,
                                                           c->d_patch_bnd);
      }
      c->n_launches++;
    }
  }
  PSC_TRY(check_launch(c, "moment_1st"));
  return bnd_add_ghosts(c, id, 0, nc);
}

int moment_rho_1st_nc(Ctx* c, int id)
{
  PSC_TRY(check_field(c, id, 0, 1));
  const GridDev& G = c->gd;
  PSC_TRY(flds_zero(c, id, 0, 1));
  // zero proxies too (they receive neighbours' ghost sums)
  if (c->n_slots > G.n_patches) {
    PSC_CUDA_TRY(cudaMemsetAsync(c->fld(id) + (size_t)G.n_patches * c->fld_slot_len(id), 0,
                                 (size_t)(c->n_slots - G.n_patches) * c->fld_slot_len(id) *
                                   sizeof(float),
                                 c->stream));
  }
  PSC_TRY(c->scr[8].reserve(pm::MAX_KINDS * sizeof(float)));
  float qk[pm::MAX_KINDS] = {};
  for (int k = 0; k < c->g.desc.n_kinds; k++) {
    qk[k] = (float)c->g.desc.q[k];
  }
  PSC_CUDA_TRY(cudaMemcpyAsync(c->scr[8].p, qk, sizeof(qk), cudaMemcpyHostToDevice, c->stream));
  PSC_CUDA_TRY(cudaStreamSynchronize(c->stream));
  if (c->n_prts) {
    KernelScope ks(c, "rho_1st_nc");
    if (c->sorted && c->opt_cell_moments) {
      const uint32_t nct = (uint32_t)G.n_cells * G.n_patches;
      k_rho_1st_nc_cells<<<div_up(nct, RHO_WARPS), RHO_WARPS * 32, 0, c->stream>>>(
        G, nct, c->d_cell_off, c->xi(), c->pxi(), c->fld(id), c->fld_slot_len(id), (float)c->g.desc.fnqs,
        c->scr[8].as<float>());
    } else {
      k_rho_1st_nc<<<div_up(c->n_prts, 256), 256, 0, c->stream>>>(
        G, c->n_prts, c->d_off, c->xi(), c->pxi(), c->fld(id), c->fld_slot_len(id),
        (float)c->g.desc.fnqs, c->scr[8].as<float>());
    }
    c->n_launches++;
  }
  // ItemMomentBnd::add_ghosts, fields_item.hxx:36-90
  for (int hi = 0; hi < 2; hi++) {
    for (int d = 0; d < 3; d++) {
      int bc = hi ? c->g.desc.bc_prt_hi[d] : c->g.desc.bc_prt_lo[d];
      if (bc != PSC_B200_BND_PRT_REFLECTING || c->g.invar[d]) {
        continue;
      }
      size_t n = (size_t)G.n_patches * G.fld_len;
      k_reflect_nc<<<div_up(n, 256), 256, 0, c->stream>>>(G, c->fld(id), c->fld_slot_len(id), d, hi,
                                                         c->d_patch_bnd);
      c->n_launches++;
    }
  }
  PSC_TRY(check_launch(c, "rho_1st_nc"));
  return bnd_add_ghosts(c, id, 0, 1);
}

static int ensure_scalar(Ctx* c, int& id)
{
  if (id < 0) {
    PSC_TRY(flds_create(c, 1, &id));
  }
  return 0;
}

int check_continuity_begin(Ctx* c)
{
  PSC_TRY(ensure_scalar(c, c->rho_m_id));
  return moment_rho_1st_nc(c, c->rho_m_id);
}

static int read_max(Ctx* c, double* d_err, double* out)
{
  PSC_CUDA_TRY(cudaMemcpyAsync(out, d_err, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  PSC_CUDA_TRY(cudaStreamSynchronize(c->stream));
  if (c->comm) {
    PSC_TRY(comm_allreduce_max(c, out, 1));
  }
  return 0;
}

int check_continuity_end(Ctx* c, double* err)
{
  if (c->rho_m_id < 0) {
    return fail("check_continuity_end without check_continuity_begin");
  }
  PSC_TRY(ensure_scalar(c, c->rho_p_id));
  PSC_TRY(moment_rho_1st_nc(c, c->rho_p_id));
  PSC_TRY(c->scr[8].reserve(64));
  double* d_err = c->scr[8].as<double>() + 2;
  PSC_CUDA_TRY(cudaMemsetAsync(d_err, 0, sizeof(double), c->stream));
  size_t n = (size_t)c->gd.n_patches * c->gd.n_cells;
  OpenLoDev O;
  for (int d = 0; d < 3; d++) {
    O.lo[d] = c->g.desc.bc_fld_lo[d] == PSC_B200_BND_FLD_OPEN;
  }
  k_continuity<<<div_up(n, 256), 256, 0, c->stream>>>(c->gd, make_dx(c), O, c->g.desc.dt, c->fld(0),
                                                     c->fld_slot_len(0), c->fld(c->rho_m_id),
                                                     c->fld(c->rho_p_id), c->d_patch_bnd, d_err);
  c->n_launches++;
  PSC_TRY(check_launch(c, "continuity"));
  PSC_TRY(read_max(c, d_err, err));
  c->last_continuity = *err;
  return 0;
}

int check_gauss(Ctx* c, double* err)
{
  PSC_TRY(ensure_scalar(c, c->rho_p_id));
  PSC_TRY(moment_rho_1st_nc(c, c->rho_p_id));
  PSC_TRY(c->scr[8].reserve(64));
  double* d_err = c->scr[8].as<double>() + 2;
  PSC_CUDA_TRY(cudaMemsetAsync(d_err, 0, sizeof(double), c->stream));
  size_t n = (size_t)c->gd.n_patches * c->gd.n_cells;
  k_gauss<<<div_up(n, 256), 256, 0, c->stream>>>(c->gd, make_dx(c), make_wall(c), c->fld(0),
                                                c->fld_slot_len(0), c->fld(c->rho_p_id),
                                                c->d_patch_bnd, d_err);
  c->n_launches++;
  PSC_TRY(check_launch(c, "gauss"));
  PSC_TRY(read_max(c, d_err, err));
  c->last_gauss = *err;
  return 0;
}

int marder(Ctx* c, double diffusion_, int loop)
{
  const GridHost& g = c->g;
  const GridDev& G = c->gd;
  double inv_sum = 0.;
  for (int d = 0; d < 3; d++) {
    if (!g.invar[d]) {
      inv_sum += g.dx_inv[d] * g.dx_inv[d];
    }
  }
  // marder_impl.hxx:160-176
  double diffusion_max = 1. / 2. / (.5 * g.desc.dt) / inv_sum;
  double diffusion = diffusion_max * (double)(float)diffusion_;
  PSC_TRY(ensure_scalar(c, c->rho_p_id));
  PSC_TRY(ensure_scalar(c, c->div_id));
  int rho = c->rho_p_id, res = c->div_id;
  PSC_TRY(moment_rho_1st_nc(c, rho));
  MarderDev M;
  float s = .5f * (float)g.desc.dt * (float)diffusion;
  for (int d = 0; d < 3; d++) {
    M.fac[d] = s * (float)g.dx_inv[d];
    M.inv[d] = g.invar[d];
  }
  size_t n = (size_t)G.n_patches * G.n_cells;
  for (int it = 0; it < loop; it++) {
    PSC_TRY(bnd_fill_ghosts(c, 0, pm::EX, pm::EX + 3));
    // res = 0 everywhere (ghosts included), then div E - rho on the interior
    PSC_CUDA_TRY(cudaMemsetAsync(c->fld(res), 0,
                                 (size_t)c->n_slots * c->fld_slot_len(res) * sizeof(float),
                                 c->stream));
    {
      KernelScope ks(c, "marder_res");
      k_marder_res<<<div_up(n, 256), 256, 0, c->stream>>>(G, make_dx(c), make_wall(c), c->fld(0),
                                                         c->fld_slot_len(0), c->fld(rho),
                                                         c->d_patch_bnd, c->fld(res));
      c->n_launches++;
    }
    PSC_TRY(bnd_fill_ghosts(c, res, 0, 1));
    {
      KernelScope ks(c, "marder_correct");
      k_marder_correct<<<div_up(n, 256), 256, 0, c->stream>>>(G, M, c->fld(0), c->fld_slot_len(0),
                                                             c->fld(res));
      c->n_launches++;
    }
  }
  PSC_TRY(check_launch(c, "marder"));
  return bnd_fill_ghosts(c, 0, pm::EX, pm::EX + 3);
}

int field_energies(Ctx* c, double out6[6], bool sync)
{
  PSC_TRY(c->scr[8].reserve(128));
  double* d = c->scr[8].as<double>() + 4;
  PSC_CUDA_TRY(cudaMemsetAsync(d, 0, 6 * sizeof(double), c->stream));
  size_t n = (size_t)c->gd.n_patches * c->gd.n_cells;
  unsigned nb = std::min<unsigned>(div_up(n, 256), 148 * 8);
  k_field_energies<<<nb, 256, 0, c->stream>>>(c->gd, c->fld(0), c->fld_slot_len(0),
                                             c->g.dx[0] * c->g.dx[1] * c->g.dx[2], d);
  c->n_launches++;
  PSC_CUDA_TRY(cudaMemcpyAsync(out6, d, 6 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  if (sync) {
    PSC_CUDA_TRY(cudaStreamSynchronize(c->stream));
  }
  return check_launch(c, "field_energies");
}

} // namespace psc_b200
