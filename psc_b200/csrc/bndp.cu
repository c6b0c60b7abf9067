// psc_b200: BndParticlesB200::operator() -- particle boundary exchange with the
// reference's exact result, including the order inside every patch:
//   [ stayers in their original order | arrivals from same-rank neighbours in the
//     receiver's direction-loop order | arrivals from other ranks by (rank, sender patch,
//     sender direction) ]
// (include/bnd_particles_impl.hxx:93-218 process_patch, include/ddc_particles.hxx:283-478).
//
// Device algorithm (any particle order):
//   1. classify every particle with the reference's arithmetic (pm::bnd_classify):
//      stay / leave in direction dir / drop; position and momentum fix-ups in place
//   2. one 64-bit scan gives each stayer its rank among the stayers and each leaver its
//      rank among the leavers (both in original order)
//   3. leavers (a few per cent) are compacted to an index list and stably sorted by
//      (destination patch, arrival class) with the radix sort of sort.cu
//   4. stayers are compacted per patch, arrivals appended behind them
// Remote destinations sort behind all local ones by (rank, sender patch, direction);
// comm.cpp ships those segments with ncclSend/ncclRecv.
#include "dev_util.cuh"

#include <algorithm>

namespace psc_b200
{

namespace
{

enum : uint8_t
{
  CODE_STAY = 0,
  CODE_DROP = 255
};

__global__ void k_bnd_classify(GridDev G, uint32_t n, const uint32_t* __restrict__ off,
                               const pm::PatchBnd* __restrict__ pbs,
                               const int* __restrict__ nei_patch, float4* __restrict__ xi4,
                               float4* __restrict__ pxi4, uint8_t* __restrict__ code)
{
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) {
    return;
  }
  int p = patch_of(off, G.n_patches, i);
  float4 X = xi4[i];
  float x[3] = {X.x, X.y, X.z};
  // fast path: inside
  int cidx = pm::cell_index(G.pc, G.ldims, x);
  if (cidx >= 0) {
    code[i] = CODE_STAY;
    return;
  }
  float4 U = pxi4[i];
  float u[3] = {U.x, U.y, U.z};
  int dir[3];
  bool drop;
  pm::PatchBnd pb = pbs[p];
  pm::bnd_classify(G.pc, pb, x, u, dir, drop);
  uint8_t cd;
  if (drop) {
    cd = CODE_DROP;
  } else if (dir[0] == 0 && dir[1] == 0 && dir[2] == 0) {
    cd = CODE_STAY;
  } else {
    int di = pm::dir2idx(dir);
    cd = (nei_patch[p * 27 + di] == -1) ? CODE_DROP : (uint8_t)(1 + di);
  }
  if (cd != CODE_DROP) {
    xi4[i] = make_float4(x[0], x[1], x[2], X.w);
    pxi4[i] = make_float4(u[0], u[1], u[2], U.w);
  }
  code[i] = cd;
}

struct CodeFlags
{
  const uint8_t* code;
  __device__ __forceinline__ uint64_t operator()(size_t i) const
  {
    uint8_t cd = code[i];
    return cd == CODE_STAY ? 1ull : (cd == CODE_DROP ? 0ull : (1ull << 32));
  }
};

__global__ void k_patch_stay_counts(const uint64_t* __restrict__ pos,
                                    const uint32_t* __restrict__ off, int n_patches,
                                    uint32_t* __restrict__ n_stay)
{
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n_patches) {
    n_stay[p] = (uint32_t)pos[off[p + 1]] - (uint32_t)pos[off[p]];
  }
}

// key of a leaver: local destination  -> dest_patch * 32 + arrival class (26 - dir idx:
//                  the receiver walks its neighbours in ascending direction order and
//                  the sender sits in direction -dir from it)
//                  remote destination -> key_remote_base + (rank * n_patches + sender
//                  patch) * 32 + sender dir idx
__global__ void k_leavers(uint32_t n, const uint32_t* __restrict__ off, int n_patches,
                          const uint8_t* __restrict__ code, const uint64_t* __restrict__ pos,
                          const int* __restrict__ nei_patch, uint32_t key_remote_base,
                          uint32_t* __restrict__ l_key, uint32_t* __restrict__ l_src,
                          uint32_t* __restrict__ n_arr)
{
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) {
    return;
  }
  uint8_t cd = code[i];
  if (cd == CODE_STAY || cd == CODE_DROP) {
    return;
  }
  int p = patch_of(off, n_patches, i);
  int di = cd - 1;
  int dest = nei_patch[p * 27 + di];
  uint32_t j = (uint32_t)(pos[i] >> 32);
  uint32_t key;
  if (dest >= 0) {
    key = (uint32_t)dest * 32u + (uint32_t)(26 - di);
    atomicAdd(&n_arr[dest], 1u);
  } else {
    int r = -2 - dest;
    key = key_remote_base + ((uint32_t)r * n_patches + p) * 32u + di;
    atomicAdd(&n_arr[n_patches], 1u);
  }
  l_key[j] = key;
  l_src[j] = i;
}

__global__ void k_place_stayers(uint32_t n, const uint32_t* __restrict__ off, int n_patches,
                                const uint8_t* __restrict__ code,
                                const uint64_t* __restrict__ pos,
                                const uint32_t* __restrict__ new_off,
                                const float4* __restrict__ xi4, const float4* __restrict__ pxi4,
                                float4* __restrict__ xo, float4* __restrict__ po)
{
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || code[i] != CODE_STAY) {
    return;
  }
  int p = patch_of(off, n_patches, i);
  uint32_t dst = new_off[p] + ((uint32_t)pos[i] - (uint32_t)pos[off[p]]);
  xo[dst] = xi4[i];
  po[dst] = pxi4[i];
}

// arr_base[q] = new_off[q] + n_stay[q] - (first index of q's arrivals in the sorted list)
__global__ void k_place_arrivals(uint32_t n_local, const uint32_t* __restrict__ l_key,
                                 const uint32_t* __restrict__ l_src,
                                 const uint32_t* __restrict__ arr_base,
                                 const float4* __restrict__ xi4, const float4* __restrict__ pxi4,
                                 float4* __restrict__ xo, float4* __restrict__ po)
{
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_local) {
    return;
  }
  uint32_t q = l_key[j] >> 5;
  uint32_t dst = arr_base[q] + j;
  uint32_t i = l_src[j];
  xo[dst] = xi4[i];
  po[dst] = pxi4[i];
}

__global__ void k_copy_prts(uint32_t n, const float4* __restrict__ xs, const float4* __restrict__ ps,
                            float4* __restrict__ xo, float4* __restrict__ po)
{
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    xo[i] = xs[i];
    po[i] = ps[i];
  }
}

} // namespace

int bnd_particles(Ctx* c)
{
  const GridDev& G = c->gd;
  const int np = G.n_patches;
  const uint32_t n = c->n_prts;
  const bool multi = c->comm != nullptr;
  if (n == 0 && !multi) {
    return 0;
  }
  // scratch[4]: code (n) ; scratch[5]: pos (n+1) u64 ; scratch[6]: per-patch tables
  PSC_TRY(c->scr[4].reserve((size_t)n + 16));
  PSC_TRY(c->scr[5].reserve(((size_t)n + 1) * sizeof(uint64_t)));
  PSC_TRY(c->scr[6].reserve((4 * (size_t)np + 8) * sizeof(uint32_t)));
  uint8_t* code = c->scr[4].as<uint8_t>();
  uint64_t* pos = c->scr[5].as<uint64_t>();
  uint32_t* d_n_stay = c->scr[6].as<uint32_t>();  // np
  uint32_t* d_n_arr = d_n_stay + np;              // np + 1 (last: remote total)
  uint32_t* d_new_off = d_n_arr + np + 1;         // np + 1
  uint32_t* d_arr_base = d_new_off + np + 1;      // np

  PSC_CUDA_TRY(cudaMemsetAsync(d_n_arr, 0, (np + 1) * sizeof(uint32_t), c->stream));
  if (n) {
    KernelScope ks(c, "bndp_classify");
    k_bnd_classify<<<div_up(n, 256), 256, 0, c->stream>>>(G, n, c->d_off, c->d_patch_bnd,
                                                         c->d_nei_patch, c->xi(), c->pxi(), code);
    c->n_launches++;
  }
  {
    KernelScope ks(c, "bndp_scan");
    PSC_TRY(scan_exclusive<uint64_t>(c, CodeFlags{code}, n, pos, c->scr[2]));
    k_patch_stay_counts<<<div_up(np, 128), 128, 0, c->stream>>>(pos, c->d_off, np, d_n_stay);
    c->n_launches++;
  }
  uint64_t tot = 0;
  PSC_CUDA_TRY(cudaMemcpyAsync(&tot, pos + n, sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
  PSC_CUDA_TRY(cudaStreamSynchronize(c->stream));
  const uint32_t n_stay_tot = (uint32_t)tot, n_leave = (uint32_t)(tot >> 32);
  c->n_dropped += n - n_stay_tot - n_leave;

  if (n_leave == 0 && n_stay_tot == n && !multi) {
    // nobody moved: the store (and its cell order, if any) is untouched
    return check_launch(c, "bnd_particles");
  }

  // leavers: compact, sort by (destination, class)
  const uint32_t key_remote_base = (uint32_t)np * 32u;
  uint32_t* l_key = nullptr;
  uint32_t* l_src = nullptr;
  if (n_leave) {
    PSC_TRY(c->scr[7].reserve(4 * (size_t)n_leave * sizeof(uint32_t)));
    uint32_t* a = c->scr[7].as<uint32_t>();
    uint32_t *k0 = a, *v0 = a + n_leave, *k1 = a + 2 * (size_t)n_leave, *v1 = a + 3 * (size_t)n_leave;
    {
      KernelScope ks(c, "bndp_leavers");
      k_leavers<<<div_up(n, 256), 256, 0, c->stream>>>(n, c->d_off, np, code, pos, c->d_nei_patch,
                                                      key_remote_base, k0, v0, d_n_arr);
      c->n_launches++;
    }
    size_t key_space = (size_t)key_remote_base + (size_t)c->g.n_ranks * np * 32;
    int bits = 1;
    while ((size_t(1) << bits) < key_space) {
      bits++;
    }
    bool in_alt = false;
    {
      KernelScope ks(c, "bndp_sort_leavers");
      PSC_TRY(sort_pairs(c, k0, v0, k1, v1, n_leave, bits, false, &in_alt));
    }
    l_key = in_alt ? k1 : k0;
    l_src = in_alt ? v1 : v0;
  }

  // counts -> new offsets (host; n_patches entries)
  std::vector<uint32_t> h_cnt(2 * np + 1);
  PSC_CUDA_TRY(cudaMemcpyAsync(h_cnt.data(), d_n_stay, (2 * np + 1) * sizeof(uint32_t),
                               cudaMemcpyDeviceToHost, c->stream));
  PSC_CUDA_TRY(cudaStreamSynchronize(c->stream));
  const uint32_t* h_n_stay = h_cnt.data();
  const uint32_t* h_n_arr = h_cnt.data() + np;
  const uint32_t n_remote = h_cnt[2 * np];
  const uint32_t n_local_arr = n_leave - n_remote;

  // remote exchange (multi-GPU): ships the tail of the sorted leaver list
  std::vector<uint32_t> n_recv(np, 0);
  float4 *xi_recv = nullptr, *pxi_recv = nullptr;
  if (multi) {
    PSC_TRY(comm_exchange_particles(c, c->xi(), c->pxi(), l_src ? l_src + n_local_arr : nullptr,
                                    l_key ? l_key + n_local_arr : nullptr, n_remote,
                                    key_remote_base, false, n_recv, &xi_recv, &pxi_recv));
  }

  std::vector<uint32_t> new_off(np + 1, 0), arr_base(np, 0);
  uint32_t first = 0;
  for (int p = 0; p < np; p++) {
    new_off[p + 1] = new_off[p] + h_n_stay[p] + h_n_arr[p] + n_recv[p];
    arr_base[p] = new_off[p] + h_n_stay[p] - first;
    first += h_n_arr[p];
  }
  const uint32_t n_new = new_off[np];
  PSC_TRY(prts_reserve(c, n_new));
  PSC_CUDA_TRY(cudaMemcpyAsync(d_new_off, new_off.data(), (np + 1) * sizeof(uint32_t),
                               cudaMemcpyHostToDevice, c->stream));
  PSC_CUDA_TRY(cudaMemcpyAsync(d_arr_base, arr_base.data(), np * sizeof(uint32_t),
                               cudaMemcpyHostToDevice, c->stream));
  if (n) {
    KernelScope ks(c, "bndp_place");
    k_place_stayers<<<div_up(n, 256), 256, 0, c->stream>>>(n, c->d_off, np, code, pos, d_new_off,
                                                          c->xi(), c->pxi(), c->xi_alt(),
                                                          c->pxi_alt());
    c->n_launches++;
    if (n_local_arr) {
      k_place_arrivals<<<div_up(n_local_arr, 256), 256, 0, c->stream>>>(
        n_local_arr, l_key, l_src, d_arr_base, c->xi(), c->pxi(), c->xi_alt(), c->pxi_alt());
      c->n_launches++;
    }
  }
  if (multi) {
    // received records are already ordered by (destination patch, sender rank, sender
    // patch, sender direction): append per patch
    uint32_t roff = 0;
    for (int p = 0; p < np; p++) {
      if (n_recv[p]) {
        uint32_t dst = new_off[p] + h_n_stay[p] + h_n_arr[p];
        k_copy_prts<<<div_up(n_recv[p], 256), 256, 0, c->stream>>>(
          n_recv[p], xi_recv + roff, pxi_recv + roff, c->xi_alt() + dst, c->pxi_alt() + dst);
        c->n_launches++;
        roff += n_recv[p];
      }
    }
  }
  PSC_CUDA_TRY(cudaStreamSynchronize(c->stream)); // new_off / arr_base are host temporaries
  c->cur ^= 1;
  c->h_off = new_off;
  c->n_prts = n_new;
  c->sorted = false;
  c->pushed_from_sorted = false;
  PSC_TRY(check_launch(c, "bnd_particles"));
  return prts_upload_off(c);
}

} // namespace psc_b200
