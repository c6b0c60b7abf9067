// psc_b200: binary Coulomb collisions on the cell-ordered particle store.
//
// What: CollisionHost::operator() (libpsc/psc_collision/psc_collision_impl.hxx:56-100:
// per cell, a random permutation of its particles, then collide_in_cell :215-252 -- pairs
// (perm[n], perm[n+1]), the first three as a triangle at half the rate when the population
// is odd) around BinaryCollision::operator() (include/binary_collision.hxx:57-295), in the
// reference's operation order for real_t = float (this file is built with -fmad=false).
//
// How (ours; the reference's CUDA version, cuda_collision.hxx, walks one thread per cell
// through its pairs serially with a curand state per thread): one WARP per cell.  The pairs
// of a permutation are disjoint, so after the permutation is known every lane takes a pair
// of its own.  The permutation is a sort of counter-based random keys (one 64-bit composite
// key per particle, bitonic sort in shared memory), the scattering angles are counter-based
// too -- no generator state in memory, the same streams whatever the decomposition, and the
// CPU oracle (oracle/psc_oracle_collision.inc) can walk exactly the same pairs with exactly
// the same random numbers, so the test compares particle for particle.  The reference's own
// streams (std::mt19937 shuffle, libc random()) are not reproducible on a device.
//
// Not built: the per-cell statistics fields (nudt min / median / max, mflds_stats_) and the
// momentum-transfer field mflds_rei_ (:168-212); both are output-only diagnostics.
#include "dev_util.cuh"

namespace psc_b200
{

namespace
{

constexpr unsigned FULL = 0xffffffffu;
constexpr int COLL_CAP = 1024; // particles of a cell permuted together (oracle: PO_COLL_CAP)
constexpr int COLL_WARPS = 4;

struct CollPrm
{
  float q[pm::MAX_KINDS], m[pm::MAX_KINDS];
  double cori, nu, dt;
  int interval;
  int rng; // 0: RngFake (uniform() = .5, identity permutation), 1: counter-based streams
  uint64_t seed, step;
  uint64_t cell0; // global index of this rank's first cell
};

__device__ __forceinline__ uint64_t mix64(uint64_t z)
{
  z += 0x9e3779b97f4a7c15ull;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}

// counter-based streams: one key per (seed, step, cell), then one mix per draw (k, stream)
// (a splitmix64 sequence seeded by the cell key; 64-bit multiplies are 4 IMADs each, and a
// four-deep hash per draw was 17 % of the kernel's instructions)
__device__ __forceinline__ uint64_t coll_cell_key(uint64_t seed, uint64_t step, uint64_t cell)
{
  return mix64(seed ^ mix64(step ^ mix64(cell)));
}
__device__ __forceinline__ uint64_t coll_hash(uint64_t cell_key, uint64_t k, uint64_t stream)
{
  return mix64(cell_key ^ (2 * k + stream));
}

// ]0, 1] with 24 random bits (RngC::uniform's range, binary_collision.hxx:13-28)
__device__ __forceinline__ float coll_u01(uint64_t h)
{
  return ((float)(h >> 40) + 1.f) * (1.f / 16777216.f);
}

// BinaryCollision::operator() (binary_collision.hxx:57-295), real_t = float.  ran1, ran2 are
// the two rng.uniform() draws.  Returns nudt.
__device__ __noinline__ float binary_collision(float u1[3], float u2[3], float q1, float m1, float q2, float m2, float nudt1,
                                  float ran1, float ran2)
{
  float px1 = u1[0], py1 = u1[1], pz1 = u1[2];
  float px2 = u2[0], py2 = u2[1], pz2 = u2[2];
  if (q1 * q2 == 0.f) {
    return 0.f; // no Coulomb collisions with neutrals
  }
  px1 = m1 * px1, py1 = m1 * py1, pz1 = m1 * pz1;
  px2 = m2 * px2, py2 = m2 * py2, pz2 = m2 * pz2;

  // absolute value of pre-collision momentum in cm-frame
  const float p01 = sqrtf(m1 * m1 + px1 * px1 + py1 * py1 + pz1 * pz1);
  const float p02 = sqrtf(m2 * m2 + px2 * px2 + py2 * py2 + pz2 * pz2);
  float h1 = p01 * p02 - px1 * px2 - py1 * py2 - pz1 * pz2;
  const float ss = m1 * m1 + m2 * m2 + 2.f * h1;
  float h2 = ss - m1 * m1 - m2 * m2;
  float h3 = (h2 * h2 - 4.f * m1 * m1 * m2 * m2) / (4.f * ss);
  if (h3 < 0.f) {
    return 0.f;
  }
  const float ppc = sqrtf(h3);

  // cm-velocity
  const float vcx = (px1 + px2) / (p01 + p02);
  const float vcy = (py1 + py2) / (p01 + p02);
  const float vcz = (pz1 + pz2) / (p01 + p02);
  const float nnorm = sqrtf(vcx * vcx + vcy * vcy + vcz * vcz);
  float nx = 0.f, ny = 0.f, nz = 0.f;
  if (nnorm > 0.f) {
    nx = vcx / nnorm, ny = vcy / nnorm, nz = vcz / nnorm;
  }
  const float bet = nnorm;
  const float gam = 1.f / sqrtf(1.f - bet * bet);

  // pre-collision momenta in cm-frame
  const float pn1 = px1 * nx + py1 * ny + pz1 * nz;
  const float pn2 = px2 * nx + py2 * ny + pz2 * nz;
  const float pc01 = sqrtf(m1 * m1 + ppc * ppc);
  const float pcx1 = px1 + (gam - 1.f) * pn1 * nx - gam * vcx * p01;
  const float pcy1 = py1 + (gam - 1.f) * pn1 * ny - gam * vcy * p01;
  const float pcz1 = pz1 + (gam - 1.f) * pn1 * nz - gam * vcz * p01;
  const float pc02 = sqrtf(m2 * m2 + ppc * ppc);
  const float pcx2 = px2 + (gam - 1.f) * pn2 * nx - gam * vcx * p02;
  const float pcy2 = py2 + (gam - 1.f) * pn2 * ny - gam * vcy * p02;
  const float pcz2 = pz2 + (gam - 1.f) * pn2 * nz - gam * vcz * p02;

  // right-handed coordinate system
  const float nn1 = sqrtf(pcx1 * pcx1 + pcy1 * pcy1 + pcz1 * pcz1);
  const float nn2 = sqrtf(pcx1 * pcx1 + pcy1 * pcy1);
  const float nn3 = nn1 * nn2;
  float nx1, ny1, nz1, nx2, ny2, nz2, nx3, ny3, nz3;
  if (nn2 != 0.f) {
    nx1 = pcx1 / nn1, ny1 = pcy1 / nn1, nz1 = pcz1 / nn1;
    nx2 = pcy1 / nn2, ny2 = -pcx1 / nn2, nz2 = 0.f;
    nx3 = -pcx1 * pcz1 / nn3, ny3 = -pcy1 * pcz1 / nn3, nz3 = nn2 * nn2 / nn3;
  } else {
    nx1 = 0.f, ny1 = 0.f, nz1 = 1.f;
    nx2 = 0.f, ny2 = 1.f, nz2 = 0.f;
    nx3 = 1.f, ny3 = 0.f, nz3 = 0.f;
  }

  // relative particle velocity in cm-frame
  const float vcx1 = pcx1 / pc01, vcy1 = pcy1 / pc01, vcz1 = pcz1 / pc01;
  const float vcx2 = pcx2 / pc02, vcy2 = pcy2 / pc02, vcz2 = pcz2 / pc02;
  const float vcn = 1.f / (1.f - (vcx1 * vcx2 + vcy1 * vcy2 + vcz1 * vcz2));
  const float vcxr = vcn * (vcx1 - vcx2);
  const float vcyr = vcn * (vcy1 - vcy2);
  const float vczr = vcn * (vcz1 - vcz2);
  float vcr = sqrtf(vcxr * vcxr + vcyr * vcyr + vczr * vczr);
  if (vcr < 1.e-20f) {
    vcr = 1.e-20f;
  }
  const float m3 = m1, m4 = m2;

  // absolute value of post-collision momentum in cm-frame
  h2 = ss - m3 * m3 - m4 * m4;
  h3 = (h2 * h2 - 4.f * m3 * m3 * m4 * m4) / (4.f * ss);
  if (h3 < 0.f) {
    return 0.f;
  }
  const float qqc = sqrtf(h3);
  const float m12 = m1 * m2 / (m1 + m2);
  const float q12 = q1 * q2;
  const float nudt = nudt1 * q12 * q12 / (m12 * m12 * vcr * vcr * vcr);

  // event generator of angles for post collision vectors
  if (ran2 < 1e-20f) {
    ran2 = 1e-20f;
  }
  const float nu = (float)(2. * M_PI) * ran1;
  float psi;
  if (nudt < 1.f) { // small angle collision
    psi = 2.f * atanf(sqrtf(-.5f * nudt * logf(1.f - ran2)));
  } else {
    psi = acosf(1.f - 2.f * ran2); // isotropic angles
  }

  // post-collision momentum in cm-frame
  h1 = cosf(psi);
  h2 = sinf(psi);
  h3 = sinf(nu);
  const float h4 = cosf(nu);
  const float pc03 = sqrtf(m3 * m3 + qqc * qqc);
  const float pcx3 = qqc * (h1 * nx1 + h2 * h3 * nx2 + h2 * h4 * nx3);
  const float pcy3 = qqc * (h1 * ny1 + h2 * h3 * ny2 + h2 * h4 * ny3);
  const float pcz3 = qqc * (h1 * nz1 + h2 * h3 * nz2 + h2 * h4 * nz3);
  const float pc04 = sqrtf(m4 * m4 + qqc * qqc);
  const float pcx4 = -pcx3, pcy4 = -pcy3, pcz4 = -pcz3;

  // post-collision momentum in lab-frame
  const float pn3 = pcx3 * nx + pcy3 * ny + pcz3 * nz;
  const float pn4 = pcx4 * nx + pcy4 * ny + pcz4 * nz;
  const float px3 = pcx3 + (gam - 1.f) * pn3 * nx + gam * vcx * pc03;
  const float py3 = pcy3 + (gam - 1.f) * pn3 * ny + gam * vcy * pc03;
  const float pz3 = pcz3 + (gam - 1.f) * pn3 * nz + gam * vcz * pc03;
  const float px4 = pcx4 + (gam - 1.f) * pn4 * nx + gam * vcx * pc04;
  const float py4 = pcy4 + (gam - 1.f) * pn4 * ny + gam * vcy * pc04;
  const float pz4 = pcz4 + (gam - 1.f) * pn4 * nz + gam * vcz * pc04;

  u1[0] = px3 / m3, u1[1] = py3 / m3, u1[2] = pz3 / m3;
  u2[0] = px4 / m4, u2[1] = py4 / m4, u2[2] = pz4 / m4;
  return nudt;
}

// One warp takes 32 consecutive cells.  Cells with more than COLL_SMALL particles are walked
// one after the other by the whole warp (permutation by a bitonic sort in shared memory, then
// one pair per lane); the small ones -- a deck's background plasma has 2-3 particles in most
// cells -- are taken one cell per LANE (insertion sort of its few keys, its pairs in sequence),
// so that a warp of 32 two-particle cells issues one collision with 32 lanes instead of 32
// collisions with one lane each.  Which lane does what never changes the result: keys, pairs
// and random numbers are functions of (seed, step, cell, index) only.
constexpr int COLL_SMALL = 8;

__global__ void __launch_bounds__(COLL_WARPS * 32)
  k_collide(GridDev G, CollPrm P, uint32_t nct, const uint32_t* __restrict__ cell_off,
            const float4* __restrict__ xi4, float4* __restrict__ pxi4, unsigned long long* __restrict__ n_coll)
{
  __shared__ uint64_t sk_all[COLL_WARPS][COLL_CAP];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint64_t* const sk = sk_all[warp];
  const uint32_t g0 = (blockIdx.x * COLL_WARPS + warp) * 32u;
  if (g0 >= nct) {
    return;
  }
  const uint32_t my_g = min(g0 + (uint32_t)lane, nct - 1);
  const uint32_t my_cb = __ldg(&cell_off[my_g]);
  const int my_nn = g0 + lane < nct ? (int)(__ldg(&cell_off[my_g + 1]) - my_cb) : 0;
  unsigned long long my_coll = 0;

  // one binary collision between records a and b (indices into the store); gcell: the cell's stream key
  auto do_bc = [&](uint32_t a, uint32_t b, float nudt1, uint64_t gcell, uint64_t s0, uint64_t j) {
    float4 ua = pxi4[a], ub = pxi4[b];
    const int ka = __float_as_int(xi4[a].w), kb = __float_as_int(xi4[b].w);
    float u1[3] = {ua.x, ua.y, ua.z}, u2[3] = {ub.x, ub.y, ub.z};
    const float r1 = P.rng ? coll_u01(coll_hash(gcell, s0 + j, 1)) : .5f;
    const float r2 = P.rng ? coll_u01(coll_hash(gcell, s0 + j, 2)) : .5f;
    binary_collision(u1, u2, P.q[ka], P.m[ka], P.q[kb], P.m[kb], nudt1, r1, r2);
    pxi4[a] = make_float4(u1[0], u1[1], u1[2], ua.w);
    pxi4[b] = make_float4(u2[0], u2[1], u2[2], ub.w);
    my_coll++;
  };
  // all particles need to have same weight (:227-229): nudt1 from the first of the permutation
  auto nudt1_of = [&](uint32_t f, int nn) {
    const int kf = __float_as_int(xi4[f].w);
    const float wni = pxi4[f].w / P.q[kf];
    return (float)((double)wni * P.cori * nn * P.interval * P.dt * P.nu);
  };

  // ---- small cells: one cell per lane
  if (my_nn >= 2 && my_nn <= COLL_SMALL) {
    const uint64_t gcell = coll_cell_key(P.seed, P.step, P.cell0 + my_g);
    uint64_t* const k8 = sk + lane * COLL_SMALL; // this lane's keys
    for (int i = 0; i < my_nn; i++) {
      // insertion sort of the composite keys (random 40 bits, index)
      const uint64_t key = P.rng ? (((coll_hash(gcell, (uint64_t)i, 0) >> 24) << 24) | (uint64_t)i)
                                 : (uint64_t)i;
      int j = i;
      while (j > 0 && k8[j - 1] > key) {
        k8[j] = k8[j - 1];
        j--;
      }
      k8[j] = key;
    }
    const float nudt1 = nudt1_of(my_cb + (uint32_t)(k8[0] & 0xffffffu), my_nn);
    int n = 0, jp = 0;
    if (my_nn & 1) { // odd # of particles: do 3-collision (:235-240)
      const uint32_t p0 = my_cb + (uint32_t)(k8[0] & 0xffffffu), p1 = my_cb + (uint32_t)(k8[1] & 0xffffffu),
                     p2 = my_cb + (uint32_t)(k8[2] & 0xffffffu);
      const float half = (float)(.5 * (double)nudt1);
      // (one call site in a loop: the collision is a large function)
      for (int t = 0; t < 3; t++) {
        do_bc(t == 2 ? p1 : p0, t == 0 ? p1 : p2, half, gcell, 0, (uint64_t)t);
      }
      n = 3, jp = 3;
    }
    for (; n < my_nn; n += 2, jp++) {
      do_bc(my_cb + (uint32_t)(k8[n] & 0xffffffu), my_cb + (uint32_t)(k8[n + 1] & 0xffffffu), nudt1, gcell, 0,
            (uint64_t)jp);
    }
  }
  __syncwarp();

  // ---- the other cells, by the whole warp.  Two at a time when both fit half the key buffer:
  // the three collisions of an odd population's triangle depend on each other and hold their
  // lane for three rounds, and a cell alone has nothing for the other 31 lanes to do meanwhile
  // (profile: 11.8 of 32 lanes active in the collision proper); with two cells in flight every
  // round is "one binary collision per lane" -- triangle lanes do their next step, the free
  // lanes take the next pairs of either cell from one pool.
  auto sort_keys = [&](uint64_t* k_, int m, uint64_t gcell, int& m2) {
    m2 = 2;
    while (m2 < m) {
      m2 <<= 1;
    }
    for (int i = lane; i < m2; i += 32) {
      uint64_t key = ~0ull; // padding sorts to the end
      if (i < m) {
        key = P.rng ? (((coll_hash(gcell, (uint64_t)i, 0) >> 24) << 24) | (uint64_t)i) : (uint64_t)i;
      }
      k_[i] = key;
    }
    __syncwarp();
    if (P.rng) {
      for (int k = 2; k <= m2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
          for (int t = lane; t < (m2 >> 1); t += 32) {
            const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
            const int l = i | j;
            const bool up = (i & k) == 0;
            const uint64_t a = k_[i], b = k_[l];
            if ((a > b) == up) {
              k_[i] = b;
              k_[l] = a;
            }
          }
          __syncwarp();
        }
      }
    }
  };
  unsigned big = __ballot_sync(FULL, my_nn > COLL_SMALL);
  while (big) {
    const int src = __ffs(big) - 1;
    big &= big - 1;
    const uint32_t cb = __shfl_sync(FULL, my_cb, src);
    const int nn = __shfl_sync(FULL, my_nn, src);
    const uint64_t gcell = coll_cell_key(P.seed, P.step, P.cell0 + g0 + (uint32_t)src);
    if (big && nn <= COLL_CAP / 2 && __shfl_sync(FULL, my_nn, __ffs(big) - 1) <= COLL_CAP / 2) {
      const int srcB = __ffs(big) - 1;
      big &= big - 1;
      const uint32_t cbB = __shfl_sync(FULL, my_cb, srcB);
      const int nnB = __shfl_sync(FULL, my_nn, srcB);
      const uint64_t gcellB = coll_cell_key(P.seed, P.step, P.cell0 + g0 + (uint32_t)srcB);
      uint64_t* const skB = sk + COLL_CAP / 2;
      int m2;
      sort_keys(sk, nn, gcell, m2);
      sort_keys(skB, nnB, gcellB, m2);
      const float nudtA = nudt1_of(cb + (uint32_t)(sk[0] & 0xffffffu), nn);
      const float nudtB = nudt1_of(cbB + (uint32_t)(skB[0] & 0xffffffu), nnB);
      const int firstA = (nn & 1) ? 3 : 0, firstB = (nnB & 1) ? 3 : 0;
      const int pairsA = (nn - firstA) >> 1, pairsB = (nnB - firstB) >> 1;
      // lane 0 / lane 1 own the triangles (steps 0..2), everybody else (and they, afterwards) the pool
      int tri = (lane == 0 && firstA) || (lane == 1 && firstB) ? 0 : 3;
      int next = 0;
      const int total = pairsA + pairsB;
      for (;;) {
        const bool busy = tri < 3;
        const unsigned fm = __ballot_sync(FULL, !busy);
        if (!__any_sync(FULL, busy) && next >= total) {
          break;
        }
        // (one call site for both kinds of work: the lanes run the collision together)
        bool act = busy, isA = lane == 0;
        int n = 0, j = tri;
        if (!busy) {
          const int idx = next + __popc(fm & ((1u << lane) - 1u));
          act = idx < total;
          isA = idx < pairsA;
          const int tp = isA ? idx : idx - pairsA;
          const int first = isA ? firstA : firstB;
          n = first + 2 * tp;
          j = first + tp;
        }
        if (act) {
          const uint64_t* const k_ = isA ? sk : skB;
          const uint32_t b0 = isA ? cb : cbB;
          uint32_t pa, pb;
          float nudt = isA ? nudtA : nudtB;
          if (busy) {
            // triangle step tri: (p0, p1), (p0, p2), (p1, p2) at half the frequency (:235-240)
            const uint32_t p0 = b0 + (uint32_t)(k_[0] & 0xffffffu), p1 = b0 + (uint32_t)(k_[1] & 0xffffffu),
                           p2 = b0 + (uint32_t)(k_[2] & 0xffffffu);
            pa = tri == 2 ? p1 : p0, pb = tri == 0 ? p1 : p2;
            nudt = (float)(.5 * (double)nudt);
          } else {
            pa = b0 + (uint32_t)(k_[n] & 0xffffffu), pb = b0 + (uint32_t)(k_[n + 1] & 0xffffffu);
          }
          do_bc(pa, pb, nudt, isA ? gcell : gcellB, 0, (uint64_t)j);
        }
        if (busy) {
          tri++;
        }
        next += __popc(fm);
      }
      __syncwarp();
      continue;
    }
    for (int s0 = 0; s0 < nn; s0 += COLL_CAP) {
      const int m = min(COLL_CAP, nn - s0);
      if (m < 2) {
        break;
      }
      // ---- randomize_in_cell (:160-166): sort composite keys (random 40 bits, index)
      int m2 = 2;
      while (m2 < m) {
        m2 <<= 1;
      }
      for (int i = lane; i < m2; i += 32) {
        uint64_t key = ~0ull; // padding sorts to the end
        if (i < m) {
          key = P.rng ? (((coll_hash(gcell, (uint64_t)(s0 + i), 0) >> 24) << 24) | (uint64_t)i)
                      : (uint64_t)i;
        }
        sk[i] = key;
      }
      __syncwarp();
      if (P.rng) {
        for (int k = 2; k <= m2; k <<= 1) {
          for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = lane; t < (m2 >> 1); t += 32) {
              // t-th compare-exchange of this stage
              const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
              const int l = i | j;
              const bool up = (i & k) == 0;
              const uint64_t a = sk[i], b = sk[l];
              if ((a > b) == up) {
                sk[i] = b;
                sk[l] = a;
              }
            }
            __syncwarp();
          }
        }
      }
      const uint32_t base = cb + (uint32_t)s0;
      const float nudt1 = nudt1_of(base + (uint32_t)(sk[0] & 0xffffffu), nn);
      // the triangle of an odd population runs on lane 0 while the others start on the pairs:
      // work item t of lane 0 is the triangle, the pairs follow
      const int first = (m & 1) ? 3 : 0;
      const int n_pairs = (m - first) >> 1;
      const int n_items = n_pairs + (first ? 1 : 0);
      for (int t = lane; t < n_items; t += 32) {
        if (first && t == 0) {
          const uint32_t p0 = base + (uint32_t)(sk[0] & 0xffffffu), p1 = base + (uint32_t)(sk[1] & 0xffffffu),
                         p2 = base + (uint32_t)(sk[2] & 0xffffffu);
          const float half = (float)(.5 * (double)nudt1);
          for (int q = 0; q < 3; q++) {
            do_bc(q == 2 ? p1 : p0, q == 0 ? p1 : p2, half, gcell, (uint64_t)s0, (uint64_t)q);
          }
        } else {
          const int tp = first ? t - 1 : t;
          const int n = first + 2 * tp;
          do_bc(base + (uint32_t)(sk[n] & 0xffffffu), base + (uint32_t)(sk[n + 1] & 0xffffffu), nudt1, gcell,
                (uint64_t)s0, (uint64_t)(first + tp));
        }
      }
      __syncwarp();
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    my_coll += __shfl_xor_sync(FULL, my_coll, o);
  }
  if (lane == 0 && n_coll) {
    atomicAdd(n_coll, my_coll);
  }
}

} // namespace

int collide(Ctx* c, const psc_b200_collision_params* prm, uint64_t* n_collisions)
{
  if (!prm) {
    return fail("null collision params");
  }
  if (!(prm->nu > 0.) || prm->interval < 1) {
    return fail("collision: nu must be positive and interval >= 1 (psc_collision_impl.hxx:45-52)");
  }
  // the pairing walks cell runs: CollisionHost asserts the store is ordered by cell
  // (find_cell_offsets :139-156); Psc::step sorts before it collides (psc.hxx:356-371)
  if (!c->sorted && c->n_prts) {
    PSC_TRY(sort_mprts(c));
  }
  const GridHost& g = c->g;
  CollPrm P{};
  for (int k = 0; k < g.desc.n_kinds; k++) {
    P.q[k] = (float)g.desc.q[k];
    P.m[k] = (float)g.desc.m[k];
  }
  P.cori = prm->cori;
  P.nu = prm->nu;
  P.dt = g.desc.dt;
  P.interval = prm->interval;
  P.rng = prm->rng;
  P.seed = prm->seed;
  P.step = prm->step;
  P.cell0 = (uint64_t)g.patch_begin * (uint64_t)g.n_cells;
  PSC_TRY(c->scr[0].reserve(sizeof(unsigned long long)));
  unsigned long long* d_n = c->scr[0].as<unsigned long long>();
  PSC_CUDA_TRY(cudaMemsetAsync(d_n, 0, sizeof(unsigned long long), c->stream));
  const uint32_t nct = (uint32_t)g.n_cells * g.n_patches;
  if (c->n_prts) {
    KernelScope ks(c, "collide");
    k_collide<<<div_up(nct, COLL_WARPS * 32), COLL_WARPS * 32, 0, c->stream>>>(c->gd, P, nct, c->d_cell_off, c->xi(),
                                                                        c->pxi(), d_n);
    c->n_launches++;
  }
  if (n_collisions) {
    unsigned long long h = 0;
    PSC_CUDA_TRY(cudaMemcpyAsync(&h, d_n, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    PSC_CUDA_TRY(cudaStreamSynchronize(c->stream));
    *n_collisions = h;
  }
  // (momenta changed, positions did not: the store stays cell-ordered, counts of a previous
  // push are unaffected)
  return check_launch(c, "collide");
}

} // namespace psc_b200

// ====================================================================== heating
//
// Heating__::operator() / kick_particle (libpsc/psc_heating/psc_heating_impl.hxx:27-76) with the
// HeatingSpotFoil profile (include/heating_spot_foil.hxx:22-89): a particle inside the spot
// gets a Gaussian momentum kick of variance H(x, kind) * interval * dt per component.  One
// thread per particle; the six uniforms are counter-based (the reference draws from libc
// random()) and shared with the oracle.

namespace psc_b200
{

namespace
{

struct HeatPrm
{
  double zl, zh, xc, yc, rH, Lx, Ly;
  double fac[pm::MAX_KINDS];
  double xb_dx[3], corner[3]; // patch origin = off * dx + corner
  float heating_dt;
  int np[3], ldims[3];
  int patch_begin;
  int xyz;
  uint64_t seed, step;
};

__device__ __forceinline__ double sqr_d(double a) { return a * a; }

__device__ double heating_H(const HeatPrm& P, const double crd[3], int kind)
{
  const double fac = P.fac[kind];
  if (fac == 0.0) {
    return 0.;
  }
  if (crd[2] <= P.zl || crd[2] >= P.zh) {
    return 0.;
  }
  if (P.rH == 0) {
    return fac; // uniform heating, not a spot
  }
  const double x = crd[0], y = crd[1], xc = P.xc, yc = P.yc, r2 = P.rH * P.rH, Lx = P.Lx, Ly = P.Ly;
  if (P.xyz) {
    return fac * (exp(-(sqr_d(x - (xc)) + sqr_d(y - (yc))) / r2) + exp(-(sqr_d(x - (xc)) + sqr_d(y - (yc + Ly))) / r2) +
                  exp(-(sqr_d(x - (xc)) + sqr_d(y - (yc - Ly))) / r2) + exp(-(sqr_d(x - (xc + Lx)) + sqr_d(y - (yc))) / r2) +
                  exp(-(sqr_d(x - (xc + Lx)) + sqr_d(y - (yc + Ly))) / r2) +
                  exp(-(sqr_d(x - (xc + Lx)) + sqr_d(y - (yc - Ly))) / r2) + exp(-(sqr_d(x - (xc - Lx)) + sqr_d(y - (yc))) / r2) +
                  exp(-(sqr_d(x - (xc - Lx)) + sqr_d(y - (yc + Ly))) / r2) +
                  exp(-(sqr_d(x - (xc - Lx)) + sqr_d(y - (yc - Ly))) / r2));
  }
  return fac * (exp(-(sqr_d(y - (yc))) / r2) + exp(-(sqr_d(y - (yc + Ly))) / r2) + exp(-(sqr_d(y - (yc - Ly))) / r2));
}

__global__ void k_heating(HeatPrm P, int n_patches, uint32_t n, const uint32_t* __restrict__ off,
                          const float4* __restrict__ xi4, float4* __restrict__ pxi4,
                          unsigned long long* __restrict__ n_kicked)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  bool kicked = false;
  if (i < n) {
    const int p = patch_of(off, n_patches, i);
    const int gp = P.patch_begin + p;
    const int idx3[3] = {gp % P.np[0], (gp / P.np[0]) % P.np[1], gp / (P.np[0] * P.np[1])};
    const float4 X = xi4[i];
    const float xf[3] = {X.x, X.y, X.z};
    double xx[3];
#pragma unroll
    for (int d = 0; d < 3; d++) {
      xx[d] = xf[d] + ((idx3[d] * P.ldims[d]) * P.xb_dx[d] + P.corner[d]);
    }
    const int kind = __float_as_int(X.w);
    const double Hd = heating_H(P, xx, kind);
    if (Hd > 0.f) {
      const float H = (float)Hd;
      const uint64_t pkey = mix64(P.seed ^ mix64(P.step ^ mix64((uint64_t)gp))); // one key per (seed, step, patch)
      float ran[6];
#pragma unroll
      for (int k = 0; k < 6; k++) {
        const uint64_t h = mix64(pkey ^ (((uint64_t)(i - off[p]) << 3) | (uint64_t)k));
        ran[k] = (float)(h >> 40) * (1.f / 16777216.f);
      }
      const float ranx = sqrtf(-2.f * logf((float)(1.0 - ran[0]))) * cosf((float)(2.f * M_PI * ran[1]));
      const float rany = sqrtf(-2.f * logf((float)(1.0 - ran[2]))) * cosf((float)(2.f * M_PI * ran[3]));
      const float ranz = sqrtf(-2.f * logf((float)(1.0 - ran[4]))) * cosf((float)(2.f * M_PI * ran[5]));
      const float Dp = sqrtf(H * P.heating_dt);
      float4 U = pxi4[i];
      U.x += Dp * ranx;
      U.y += Dp * rany;
      U.z += Dp * ranz;
      pxi4[i] = U;
      kicked = true;
    }
  }
  const unsigned m = __ballot_sync(FULL, kicked);
  if (m && (threadIdx.x & 31) == 0 && n_kicked) {
    atomicAdd(n_kicked, (unsigned long long)__popc(m));
  }
}

} // namespace

int heating_spot_foil(Ctx* c, const psc_b200_heating_params* prm, uint64_t* n_kicked)
{
  if (!prm) {
    return fail("null heating params");
  }
  const GridHost& g = c->g;
  if (prm->n_kinds > g.desc.n_kinds || prm->n_kinds >= PSC_B200_MAX_KINDS) {
    return fail("heating: n_kinds out of range (heating_spot_foil.hxx:36)");
  }
  HeatPrm P{};
  P.zl = prm->zl, P.zh = prm->zh, P.xc = prm->xc, P.yc = prm->yc, P.rH = prm->rH;
  P.Lx = g.desc.length[0], P.Ly = g.desc.length[1];
  const double width = prm->zh - prm->zl;
  for (int k = 0; k < prm->n_kinds; k++) {
    P.fac[k] = (8.f * pow(prm->T[k], 1.5)) / (sqrt(prm->Mi) * width);
  }
  for (int d = 0; d < 3; d++) {
    P.xb_dx[d] = g.dx[d];
    P.corner[d] = g.desc.corner[d];
    P.np[d] = g.desc.np[d];
    P.ldims[d] = g.ldims[d];
  }
  P.heating_dt = (float)(prm->interval * g.desc.dt);
  P.patch_begin = g.patch_begin;
  P.xyz = g.dim == pm::DIM_XYZ;
  P.seed = prm->seed, P.step = prm->step;
  PSC_TRY(c->scr[0].reserve(sizeof(unsigned long long)));
  unsigned long long* d_n = c->scr[0].as<unsigned long long>();
  PSC_CUDA_TRY(cudaMemsetAsync(d_n, 0, sizeof(unsigned long long), c->stream));
  if (c->n_prts) {
    KernelScope ks(c, "heating");
    k_heating<<<div_up(c->n_prts, 256), 256, 0, c->stream>>>(P, g.n_patches, c->n_prts, c->d_off, c->xi(), c->pxi(), d_n);
    c->n_launches++;
  }
  if (n_kicked) {
    unsigned long long h = 0;
    PSC_CUDA_TRY(cudaMemcpyAsync(&h, d_n, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    PSC_CUDA_TRY(cudaStreamSynchronize(c->stream));
    *n_kicked = h;
  }
  return check_launch(c, "heating");
}

} // namespace psc_b200
