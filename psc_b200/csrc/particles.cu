// psc_b200: the particle store (MparticlesB200).
//
// Mirrors what MparticlesSimple / InjectorSimple / ConstAccessorSimple provide to a
// deck (src/include/particles_simple.hxx:123-247, injector_simple.hxx:24-43,
// particles.hxx:35-54 get_as<MparticlesSingle>) but keeps the data on the device as
// two float4 streams so that every kernel moves a particle with two coalesced
// 128-bit accesses.  Host records are PSC's 32-byte ParticleSimple<float>
// (particle_simple.hxx:10-42) and are transposed on the device, chunk by chunk.
#include "dev_util.cuh"

#include <algorithm>
#include <cstring>

namespace psc_b200
{

namespace
{

constexpr size_t STAGE_PRTS = size_t(1) << 22; // records per host<->device chunk (128 MiB)

// AoS record r = {x0 x1 x2 u0 | u1 u2 kind qw}
__global__ void k_aos_to_soa(const float4* __restrict__ aos, uint32_t n_chunk, uint32_t i0,
                             const uint32_t* __restrict__ inj_off,
                             const uint32_t* __restrict__ dst_base, int n_patches,
                             float4* __restrict__ xi4, float4* __restrict__ pxi4)
{
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_chunk) {
    return;
  }
  uint32_t i = i0 + t; // index into the injected array
  int p = patch_of(inj_off, n_patches, i);
  uint32_t dst = dst_base[p] + (i - inj_off[p]);
  float4 a = aos[2 * (size_t)t], b = aos[2 * (size_t)t + 1];
  xi4[dst] = make_float4(a.x, a.y, a.z, b.z);
  pxi4[dst] = make_float4(a.w, b.x, b.y, b.w);
}

__global__ void k_soa_to_aos(const float4* __restrict__ xi4, const float4* __restrict__ pxi4,
                             uint32_t i0, uint32_t n_chunk, float4* __restrict__ aos)
{
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_chunk) {
    return;
  }
  float4 x = xi4[i0 + t], u = pxi4[i0 + t];
  aos[2 * (size_t)t] = make_float4(x.x, x.y, x.z, u.x);
  aos[2 * (size_t)t + 1] = make_float4(u.y, u.z, x.w, u.w);
}

// move patch segments to their new offsets (old store -> alt store)
__global__ void k_move_segments(const float4* __restrict__ xi4, const float4* __restrict__ pxi4,
                                uint32_t n, const uint32_t* __restrict__ off,
                                const uint32_t* __restrict__ new_off, int n_patches,
                                float4* __restrict__ xo, float4* __restrict__ po)
{
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) {
    return;
  }
  int p = patch_of(off, n_patches, i);
  uint32_t dst = new_off[p] + (i - off[p]);
  xo[dst] = xi4[i];
  po[dst] = pxi4[i];
}

// counter-based generator: splitmix64 finaliser over (seed, particle id, stream)
__device__ __forceinline__ uint64_t mix64(uint64_t z)
{
  z += 0x9e3779b97f4a7c15ull;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}

__device__ __forceinline__ float u01(uint64_t h)
{ // (0,1)
  return ((float)(h >> 40) + 0.5f) * (1.f / 16777216.f);
}

struct ThermalPrm
{
  int ppc, n_kinds;
  float vth[pm::MAX_KINDS], q[pm::MAX_KINDS];
  float dx[3];
  uint64_t seed;
  uint64_t id0; // global id of this rank's first particle
};

// ppc_by_patch == nullptr: T.ppc particles per cell and kind everywhere; else patch p holds
// ppc_by_patch[p] per cell and kind and starts at off[p] (a density profile by patch)
__global__ void k_setup_thermal(GridDev G, ThermalPrm T, uint32_t n, const int* __restrict__ ppc_by_patch,
                                const uint32_t* __restrict__ off, float4* __restrict__ xi4,
                                float4* __restrict__ pxi4)
{
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) {
    return;
  }
  uint32_t ppc = T.ppc, il = i;
  if (ppc_by_patch) {
    const int p = patch_of(off, G.n_patches, i);
    ppc = ppc_by_patch[p];
    il = i - off[p];
  }
  uint32_t per_cell = ppc * T.n_kinds;
  uint32_t cell_g = il / per_cell; // (patch * n_cells +) cell
  uint32_t r = il - cell_g * per_cell;
  int kind = r / ppc;
  uint32_t cell = cell_g % G.n_cells;
  int c[3];
  c[0] = cell % G.ldims[0];
  c[1] = (cell / G.ldims[0]) % G.ldims[1];
  c[2] = cell / (G.ldims[0] * G.ldims[1]);
  uint64_t id = T.id0 + i;
  uint64_t h0 = mix64(T.seed ^ mix64(id * 4 + 0));
  uint64_t h1 = mix64(T.seed ^ mix64(id * 4 + 1));
  uint64_t h2 = mix64(T.seed ^ mix64(id * 4 + 2));
  uint64_t h3 = mix64(T.seed ^ mix64(id * 4 + 3));
  float rr[3] = {u01(h0), u01(h0 << 24), u01(h1)};
  float x[3];
#pragma unroll
  for (int d = 0; d < 3; d++) {
    x[d] = ((float)c[d] + rr[d]) * T.dx[d];
    // the particle has to index into its own cell (particle_indexer.hxx:74-94)
    if (pm::fint(x[d] * G.pc.dxi_idx[d]) != c[d] || pm::fint(x[d] * G.pc.dxi[d]) != c[d]) {
      x[d] = ((float)c[d] + .5f) * T.dx[d];
    }
  }
  // Box-Muller
  float a0 = sqrtf(-2.f * logf(u01(h1 << 24))), ph0 = 6.2831853f * u01(h2);
  float a1 = sqrtf(-2.f * logf(u01(h2 << 24))), ph1 = 6.2831853f * u01(h3);
  float vth = T.vth[kind];
  float u0 = vth * a0 * cosf(ph0), u1 = vth * a0 * sinf(ph0), u2 = vth * a1 * cosf(ph1);
  xi4[i] = make_float4(x[0], x[1], x[2], __int_as_float(kind));
  pxi4[i] = make_float4(u0, u1, u2, T.q[kind]);
}

__global__ void k_iota_cell_off(uint32_t n_cells_total, uint32_t per_cell, uint32_t* cell_off)
{
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i <= n_cells_total) {
    cell_off[i] = i * per_cell;
  }
}

__global__ void k_profile_cell_off(uint32_t n_cells_total, uint32_t n_cells, int n_kinds,
                                   const int* __restrict__ ppc_by_patch, const uint32_t* __restrict__ off,
                                   uint32_t* cell_off)
{
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_cells_total) {
    const uint32_t p = i / n_cells;
    cell_off[i] = off[p] + (i - p * n_cells) * (uint32_t)(ppc_by_patch[p] * n_kinds);
  } else if (i == n_cells_total) {
    cell_off[i] = off[n_cells_total / n_cells];
  }
}

struct KindPrm
{
  float q[pm::MAX_KINDS], m[pm::MAX_KINDS];
  double fnqs, fac;
};

// DiagEnergiesParticle.h:15-40
__global__ void k_prt_energies(const float4* __restrict__ pxi4, const float4* __restrict__ xi4,
                               uint32_t n, const uint32_t* __restrict__ d_n, KindPrm K, double* out2)
{
  if (d_n) {
    n = *d_n;
  }
  double e_neg = 0., e_pos = 0.;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float4 u = pxi4[i];
    int kind = __float_as_int(xi4[i].w);
    float qf = K.q[kind], mf = K.m[kind];
    float w = u.w / qf;
    double gamma = sqrtf(1.f + u.x * u.x + u.y * u.y + u.z * u.z);
    double ekin = (gamma - 1.) * mf * w * K.fnqs;
    if (qf < 0.f) {
      e_neg += ekin * K.fac;
    } else if (qf > 0.f) {
      e_pos += ekin * K.fac;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    e_neg += __shfl_xor_sync(0xffffffffu, e_neg, o);
    e_pos += __shfl_xor_sync(0xffffffffu, e_pos, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&out2[0], e_neg);
    atomicAdd(&out2[1], e_pos);
  }
}

} // namespace

int prts_reserve(Ctx* c, size_t n)
{
  if (n <= c->cap) {
    return 0;
  }
  size_t ncap = std::max(n, c->cap + c->cap / 4);
  ncap = (ncap + 1023) & ~size_t(1023);
  if (ncap >= (size_t(1) << 32)) {
    return fail("more than 2^32 particles on one rank (PSC indexes particles with uint)");
  }
  // one pair of buffers at a time (the idle pair first), so that the peak is the old
  // current pair + the new buffers, not twice everything.  If an allocation fails the store
  // is left EMPTY and consistent (cap = 0, no particles) and the error is returned: a caller
  // that carries on gets "no particles", never kernels on null buffers.
  const int cur = c->cur, alt = c->cur ^ 1;
  auto lost = [&](const char* what, cudaError_t e) {
    for (int b = 0; b < 2; b++) {
      cudaFree(c->xi4[b]);
      cudaFree(c->pxi4[b]);
      c->xi4[b] = c->pxi4[b] = nullptr;
    }
    c->cap = 0;
    c->n_prts = 0;
    std::fill(c->h_off.begin(), c->h_off.end(), 0u);
    c->sorted = c->pushed_from_sorted = c->counts_valid = false;
    cudaMemsetAsync(c->d_off, 0, c->h_off.size() * sizeof(uint32_t), c->stream);
    cudaGetLastError();
    return fail(std::string("prts_reserve: ") + what + ": " + cudaGetErrorString(e) +
                " (the particle store was released)");
  };
  cudaError_t e;
  cudaFree(c->xi4[alt]);
  cudaFree(c->pxi4[alt]);
  c->xi4[alt] = c->pxi4[alt] = nullptr;
  if ((e = cudaMalloc(&c->xi4[alt], ncap * sizeof(float4))) != cudaSuccess ||
      (e = cudaMalloc(&c->pxi4[alt], ncap * sizeof(float4))) != cudaSuccess) {
    return lost("growing the idle buffers", e);
  }
  if (c->n_prts) {
    if ((e = cudaMemcpyAsync(c->xi4[alt], c->xi4[cur], c->n_prts * sizeof(float4), cudaMemcpyDeviceToDevice,
                             c->stream)) != cudaSuccess ||
        (e = cudaMemcpyAsync(c->pxi4[alt], c->pxi4[cur], c->n_prts * sizeof(float4), cudaMemcpyDeviceToDevice,
                             c->stream)) != cudaSuccess ||
        (e = cudaStreamSynchronize(c->stream)) != cudaSuccess) {
      return lost("copying the store", e);
    }
  }
  cudaFree(c->xi4[cur]);
  cudaFree(c->pxi4[cur]);
  c->xi4[cur] = c->pxi4[cur] = nullptr;
  c->cur = alt; // the data live in the pair that was grown first
  if ((e = cudaMalloc(&c->xi4[cur], ncap * sizeof(float4))) != cudaSuccess ||
      (e = cudaMalloc(&c->pxi4[cur], ncap * sizeof(float4))) != cudaSuccess) {
    return lost("growing the second pair of buffers", e);
  }
  c->cap = ncap;
  return 0;
}

int prts_upload_off(Ctx* c)
{
  PSC_CUDA_TRY(cudaMemcpyAsync(c->d_off, c->h_off.data(), c->h_off.size() * sizeof(uint32_t),
                               cudaMemcpyHostToDevice, c->stream));
  // h_off may be rewritten by the caller right away
  PSC_CUDA_TRY(cudaStreamSynchronize(c->stream));
  return 0;
}

int prts_inject(Ctx* c, const void* aos, const uint32_t* n_by_patch)
{
  int np = c->g.n_patches;
  std::vector<uint32_t> inj_off(np + 1, 0), new_off(np + 1, 0), dst_base(np);
  for (int p = 0; p < np; p++) {
    inj_off[p + 1] = inj_off[p] + n_by_patch[p];
  }
  size_t n_inj = inj_off[np];
  if (n_inj == 0) {
    return 0;
  }
  size_t n_new = (size_t)c->n_prts + n_inj;
  PSC_TRY(prts_reserve(c, n_new));
  for (int p = 0; p < np; p++) {
    uint32_t n_old = c->h_off[p + 1] - c->h_off[p];
    new_off[p + 1] = new_off[p] + n_old + n_by_patch[p];
    dst_base[p] = new_off[p] + n_old;
  }
  // tables: [inj_off (np+1) | new_off (np+1) | dst_base (np)]
  PSC_TRY(c->scr[0].reserve((3 * np + 2) * sizeof(uint32_t)));
  uint32_t* d_inj_off = c->scr[0].as<uint32_t>();
  uint32_t* d_new_off = d_inj_off + np + 1;
  uint32_t* d_dst_base = d_new_off + np + 1;
  PSC_CUDA_TRY(cudaMemcpyAsync(d_inj_off, inj_off.data(), (np + 1) * sizeof(uint32_t),
                               cudaMemcpyHostToDevice, c->stream));
  PSC_CUDA_TRY(cudaMemcpyAsync(d_new_off, new_off.data(), (np + 1) * sizeof(uint32_t),
                               cudaMemcpyHostToDevice, c->stream));
  PSC_CUDA_TRY(cudaMemcpyAsync(d_dst_base, dst_base.data(), np * sizeof(uint32_t),
                               cudaMemcpyHostToDevice, c->stream));
  float4 *xo = c->xi(), *po = c->pxi();
  if (c->n_prts) {
    // existing particles keep their order at the head of each patch
    xo = c->xi_alt();
    po = c->pxi_alt();
    k_move_segments<<<div_up(c->n_prts, 256), 256, 0, c->stream>>>(
      c->xi(), c->pxi(), c->n_prts, c->d_off, d_new_off, np, xo, po);
    c->n_launches++;
    c->cur ^= 1;
  }
  size_t chunk = std::min(n_inj, STAGE_PRTS);
  PSC_TRY(c->stage.reserve(chunk * 32));
  for (size_t i0 = 0; i0 < n_inj; i0 += chunk) {
    size_t nc = std::min(chunk, n_inj - i0);
    PSC_CUDA_TRY(cudaMemcpyAsync(c->stage.p, (const char*)aos + i0 * 32, nc * 32,
                                 cudaMemcpyHostToDevice, c->stream));
    k_aos_to_soa<<<div_up(nc, 256), 256, 0, c->stream>>>(c->stage.as<float4>(), (uint32_t)nc,
                                                        (uint32_t)i0, d_inj_off, d_dst_base, np,
                                                        xo, po);
    c->n_launches++;
    PSC_CUDA_TRY(cudaStreamSynchronize(c->stream)); // stage is reused
  }
  PSC_TRY(check_launch(c, "prts_inject"));
  c->h_off = new_off;
  c->n_prts = (uint32_t)n_new;
  c->sorted = false;
  c->pushed_from_sorted = false;
  return prts_upload_off(c);
}

int prts_set(Ctx* c, const void* aos, const uint32_t* n_by_patch)
{
  c->n_prts = 0;
  std::fill(c->h_off.begin(), c->h_off.end(), 0u);
  c->sorted = false;
  c->pushed_from_sorted = false;
  PSC_TRY(prts_upload_off(c));
  return prts_inject(c, aos, n_by_patch);
}

int prts_get(Ctx* c, void* aos, uint32_t* off)
{
  if (off) {
    memcpy(off, c->h_off.data(), c->h_off.size() * sizeof(uint32_t));
  }
  size_t n = c->n_prts;
  if (n == 0 || !aos) {
    return 0;
  }
  size_t chunk = std::min(n, STAGE_PRTS);
  PSC_TRY(c->stage.reserve(chunk * 32));
  for (size_t i0 = 0; i0 < n; i0 += chunk) {
    size_t nc = std::min(chunk, n - i0);
    k_soa_to_aos<<<div_up(nc, 256), 256, 0, c->stream>>>(c->xi(), c->pxi(), (uint32_t)i0,
                                                        (uint32_t)nc, c->stage.as<float4>());
    c->n_launches++;
    PSC_CUDA_TRY(cudaMemcpyAsync((char*)aos + i0 * 32, c->stage.p, nc * 32,
                                 cudaMemcpyDeviceToHost, c->stream));
    PSC_CUDA_TRY(cudaStreamSynchronize(c->stream));
  }
  return check_launch(c, "prts_get");
}

int prts_setup_thermal(Ctx* c, int ppc, const int* ppc_by_patch, const double* vth, uint64_t seed)
{
  const GridHost& g = c->g;
  std::vector<uint32_t> off(g.n_patches + 1, 0);
  size_t n = 0, id0 = 0;
  for (int p = 0; p < g.n_patches; p++) {
    const int pp = ppc_by_patch ? ppc_by_patch[p] : ppc;
    if (pp < 0) {
      return fail("setup_thermal: negative particle count");
    }
    n += (size_t)g.n_cells * pp * g.desc.n_kinds;
    if (n >= (size_t(1) << 32)) {
      return fail("setup_thermal: more than 2^32 particles on one rank");
    }
    off[p + 1] = (uint32_t)n;
  }
  if (!ppc_by_patch) {
    id0 = (size_t)g.patch_begin * g.n_cells * ppc * g.desc.n_kinds;
  } else {
    id0 = (size_t)g.patch_begin << 32; // (distinct streams per rank; ids need not be dense)
  }
  PSC_TRY(prts_reserve(c, n));
  ThermalPrm T{};
  T.ppc = ppc;
  T.n_kinds = g.desc.n_kinds;
  for (int k = 0; k < T.n_kinds; k++) {
    T.vth[k] = (float)vth[k];
    T.q[k] = (float)g.desc.q[k];
  }
  for (int d = 0; d < 3; d++) {
    T.dx[d] = (float)g.dx[d];
  }
  T.seed = seed;
  T.id0 = id0;
  c->h_off = off;
  c->n_prts = (uint32_t)n;
  PSC_TRY(prts_upload_off(c));
  const int* d_ppc = nullptr;
  if (ppc_by_patch) {
    PSC_TRY(c->scr[0].reserve(g.n_patches * sizeof(int)));
    PSC_CUDA_TRY(cudaMemcpyAsync(c->scr[0].p, ppc_by_patch, g.n_patches * sizeof(int), cudaMemcpyHostToDevice,
                                 c->stream));
    d_ppc = c->scr[0].as<int>();
  }
  if (n) {
    KernelScope ks(c, "setup_thermal");
    k_setup_thermal<<<div_up(n, 256), 256, 0, c->stream>>>(c->gd, T, (uint32_t)n, d_ppc, c->d_off, c->xi(),
                                                          c->pxi());
    c->n_launches++;
  }
  {
    uint32_t nct = (uint32_t)g.n_cells * g.n_patches;
    if (ppc_by_patch) {
      k_profile_cell_off<<<div_up(nct + 1, 256), 256, 0, c->stream>>>(nct, (uint32_t)g.n_cells, g.desc.n_kinds,
                                                                      d_ppc, c->d_off, c->d_cell_off);
    } else {
      k_iota_cell_off<<<div_up(nct + 1, 256), 256, 0, c->stream>>>(nct, (uint32_t)(ppc * g.desc.n_kinds),
                                                                   c->d_cell_off);
    }
    c->n_launches++;
  }
  c->sorted = true;
  c->pushed_from_sorted = false;
  PSC_CUDA_TRY(cudaStreamSynchronize(c->stream)); // (ppc_by_patch is the caller's)
  return check_launch(c, "setup_thermal");
}

// every float in [1, 2^80): the guard-free sequences of pic_math.cuh against the
// compiler's correctly rounded 1.f / sqrtf(x) and 1.f / x
__global__ void k_selftest_math(unsigned long long* n_bad)
{
  const uint32_t lo = 0x3f800000u, hi = 0x3f800000u + (80u << 23);
  unsigned long long bad = 0;
  for (uint64_t b = lo + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b < hi;
       b += (uint64_t)gridDim.x * blockDim.x) {
    const float x = __uint_as_float((uint32_t)b);
    const float a0 = 1.f / sqrtf(x), a1 = pm::rsqrt_ref(x);
    const float b0 = 1.f / x, b1 = pm::rcp_ref(x);
    bad += (__float_as_uint(a0) != __float_as_uint(a1)) + (__float_as_uint(b0) != __float_as_uint(b1));
  }
  if (bad) {
    atomicAdd(n_bad, bad);
  }
}

int selftest_math(Ctx* c, uint64_t* n_bad)
{
  PSC_TRY(c->scr[0].reserve(sizeof(unsigned long long)));
  unsigned long long* d = c->scr[0].as<unsigned long long>();
  PSC_CUDA_TRY(cudaMemsetAsync(d, 0, sizeof(unsigned long long), c->stream));
  k_selftest_math<<<148 * 16, 256, 0, c->stream>>>(d);
  c->n_launches++;
  unsigned long long h = 0;
  PSC_CUDA_TRY(cudaMemcpyAsync(&h, d, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
  PSC_CUDA_TRY(cudaStreamSynchronize(c->stream));
  *n_bad = h;
  return check_launch(c, "selftest_math");
}

int prts_energies(Ctx* c, double out2[2], bool sync, const uint32_t* d_n, bool alt)
{
  const GridHost& g = c->g;
  KindPrm K{};
  for (int k = 0; k < g.desc.n_kinds; k++) {
    K.q[k] = (float)g.desc.q[k];
    K.m[k] = (float)g.desc.m[k];
  }
  K.fnqs = g.desc.fnqs;
  K.fac = g.dx[0] * g.dx[1] * g.dx[2];
  PSC_TRY(c->scr[0].reserve(2 * sizeof(double)));
  double* d = c->scr[0].as<double>();
  PSC_CUDA_TRY(cudaMemsetAsync(d, 0, 2 * sizeof(double), c->stream));
  if (c->n_prts) {
    KernelScope ks(c, "prt_energies");
    unsigned nb = std::min<unsigned>(div_up(c->n_prts, 256), 148 * 8);
    k_prt_energies<<<nb, 256, 0, c->stream>>>(alt ? c->pxi_alt() : c->pxi(), alt ? c->xi_alt() : c->xi(), c->n_prts,
                                             d_n, K, d);
    c->n_launches++;
  }
  PSC_CUDA_TRY(cudaMemcpyAsync(out2, d, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  if (sync) {
    PSC_CUDA_TRY(cudaStreamSynchronize(c->stream));
  }
  return check_launch(c, "prt_energies");
}

} // namespace psc_b200
