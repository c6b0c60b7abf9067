// psc_b200: multi-GPU plumbing -- one process per GPU, NCCL point-to-point over
// NVLink/NVSwitch in place of every MPI site of the reference's hot path:
//   field halos       mrc_ddc_multi ddc_run_begin/end (libmrc/src/mrc_ddc_multi.c:458-565)
//   particle exchange ddc_particles::comm (include/ddc_particles.hxx:283-478)
//   scalar reductions MPI_Allreduce in checks / energies
//   balance           Balance_::balance (libpsc/psc_balance/psc_balance_impl.hxx:770-1026)
//
// Halo design: every patch owned by a neighbouring rank that touches one of ours has a
// *proxy slot* behind our own patches in each field array.  An exchange packs the strips
// the neighbour's kernels will read (inside strips for fill, outside strips for add),
// ships them with one grouped ncclSend/ncclRecv per neighbour rank, and unpacks them
// into the proxies; the gather kernels of fields.cu then treat proxies like local
// patches.  Both sides enumerate (patch ascending, direction ascending), so no headers
// are sent.  NCCL is bound at run time with dlopen so that the library can be loaded
// in processes that already carry another NCCL (torch) and on hosts without it.
#include "dev_util.cuh"

#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstring>
#include <stdexcept>

namespace psc_b200
{

namespace
{

struct NcclApi
{
  void* h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi g_nccl;

int nccl_load()
{
  if (g_nccl.h) {
    return 0;
  }
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* n : names) {
    h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (h) {
      break;
    }
  }
  if (!h) {
    return fail(std::string("cannot load NCCL: ") + dlerror());
  }
#define SYM(field, name)                                                                 \
  g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(h, name));               \
  if (!g_nccl.field) {                                                                   \
    return fail("NCCL symbol missing: " name);                                           \
  }
  SYM(GetUniqueId, "ncclGetUniqueId")
  SYM(CommInitRank, "ncclCommInitRank")
  SYM(CommDestroy, "ncclCommDestroy")
  SYM(Send, "ncclSend")
  SYM(Recv, "ncclRecv")
  SYM(GroupStart, "ncclGroupStart")
  SYM(GroupEnd, "ncclGroupEnd")
  SYM(AllReduce, "ncclAllReduce")
  SYM(AllGather, "ncclAllGather")
  SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
  g_nccl.h = h;
  return 0;
}

#define PSC_NCCL_TRY(expr)                                                               \
  do {                                                                                   \
    ncclResult_t r__ = (expr);                                                           \
    if (r__ != ncclSuccess) {                                                            \
      return fail(std::string(#expr) + ": " + g_nccl.GetErrorString(r__));               \
    }                                                                                    \
  } while (0)

struct Box
{
  int slot;
  int lo[3], n[3];
  uint32_t off; // first element (per component) in the rank buffer
};

struct HaloPlan
{
  // per neighbour rank: boxes to pack (local patches) and to unpack (proxies)
  std::vector<int> ranks;
  std::vector<uint32_t> send_cells, recv_cells; // per rank, per component
  std::vector<uint32_t> send_first, recv_first; // first box of each rank
  Box* d_send = nullptr;
  Box* d_recv = nullptr;
  uint32_t* d_send_boff = nullptr; // cumulative cells over all boxes (n+1)
  uint32_t* d_recv_boff = nullptr;
  int n_send = 0, n_recv = 0;
  uint32_t send_total = 0, recv_total = 0;
  bool built = false;
};

} // namespace

struct Comm
{
  ncclComm_t comm = nullptr;
  HaloPlan plan[2]; // 0 = fill (inside strips), 1 = add (outside strips)
  DevBuf sendbuf, recvbuf, cnt, prt_send, prt_recv, prt_final, seg;
};

namespace
{

// strip of a patch in direction dir: inside (what a fill ships) or outside (what an add ships)
void strip_box(const GridHost& g, const int dir[3], bool inside, int lo[3], int n[3])
{
  for (int d = 0; d < 3; d++) {
    int L = g.ldims[d], bn = g.ibn[d];
    switch (dir[d]) {
      case -1:
        lo[d] = inside ? 0 : -bn;
        n[d] = bn;
        break;
      case 0:
        lo[d] = 0;
        n[d] = L;
        break;
      default:
        lo[d] = inside ? L - bn : L;
        n[d] = bn;
        break;
    }
  }
}

int build_plan(Ctx* c, HaloPlan& P, bool inside)
{
  const GridHost& g = c->g;
  std::vector<Box> sb, rb;
  std::vector<uint32_t> sboff{0}, rboff{0};
  P.ranks.clear();
  for (int r = 0; r < g.n_ranks; r++) {
    if (r == g.rank) {
      continue;
    }
    size_t s0 = sb.size(), r0 = rb.size();
    uint32_t scells = 0, rcells = 0;
    // what we send to r: (our patch p asc, dir asc) with the neighbour on r
    for (int p = 0; p < g.n_patches; p++) {
      for (int di = 0; di < 27; di++) {
        if (di == 13) {
          continue;
        }
        int dir[3] = {di % 3 - 1, (di / 3) % 3 - 1, di / 9 - 1};
        int ngp = g.neighbor_patch(g.patch_begin + p, dir);
        if (ngp < 0 || g.rank_of_patch(ngp) != r) {
          continue;
        }
        Box b;
        b.slot = p;
        strip_box(g, dir, inside, b.lo, b.n);
        uint32_t cells = (uint32_t)b.n[0] * b.n[1] * b.n[2];
        if (!cells) {
          continue;
        }
        b.off = scells;
        scells += cells;
        sb.push_back(b);
        sboff.push_back(sboff.back() + cells);
      }
    }
    // what r sends to us: (its patch gp asc, dir asc) with the neighbour on our rank
    for (int gp = g.patch_off_by_rank[r]; gp < g.patch_off_by_rank[r + 1]; gp++) {
      for (int di = 0; di < 27; di++) {
        if (di == 13) {
          continue;
        }
        int dir[3] = {di % 3 - 1, (di / 3) % 3 - 1, di / 9 - 1};
        int ngp = g.neighbor_patch(gp, dir);
        if (ngp < 0 || g.rank_of_patch(ngp) != g.rank) {
          continue;
        }
        Box b;
        auto it = std::lower_bound(c->proxy_gp.begin(), c->proxy_gp.end(), gp);
        if (it == c->proxy_gp.end() || *it != gp) {
          return fail("halo plan: missing proxy slot");
        }
        b.slot = g.n_patches + (int)(it - c->proxy_gp.begin());
        strip_box(g, dir, inside, b.lo, b.n);
        uint32_t cells = (uint32_t)b.n[0] * b.n[1] * b.n[2];
        if (!cells) {
          continue;
        }
        b.off = rcells;
        rcells += cells;
        rb.push_back(b);
        rboff.push_back(rboff.back() + cells);
      }
    }
    if (sb.size() > s0 || rb.size() > r0) {
      P.ranks.push_back(r);
      P.send_cells.push_back(scells);
      P.recv_cells.push_back(rcells);
      P.send_first.push_back((uint32_t)s0);
      P.recv_first.push_back((uint32_t)r0);
    }
  }
  P.n_send = (int)sb.size();
  P.n_recv = (int)rb.size();
  P.send_total = sboff.back();
  P.recv_total = rboff.back();
  // make box offsets global over the whole buffer (rank segments back to back)
  {
    uint32_t acc = 0;
    for (size_t k = 0; k < P.ranks.size(); k++) {
      size_t e = k + 1 < P.ranks.size() ? P.send_first[k + 1] : sb.size();
      for (size_t b = P.send_first[k]; b < e; b++) {
        sb[b].off += acc;
      }
      acc += P.send_cells[k];
    }
    acc = 0;
    for (size_t k = 0; k < P.ranks.size(); k++) {
      size_t e = k + 1 < P.ranks.size() ? P.recv_first[k + 1] : rb.size();
      for (size_t b = P.recv_first[k]; b < e; b++) {
        rb[b].off += acc;
      }
      acc += P.recv_cells[k];
    }
  }
  auto up = [&](const void* src, size_t bytes, void** dst) -> int {
    PSC_CUDA_TRY(cudaMalloc(dst, std::max<size_t>(bytes, 16)));
    PSC_CUDA_TRY(cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice));
    return 0;
  };
  PSC_TRY(up(sb.data(), sb.size() * sizeof(Box), (void**)&P.d_send));
  PSC_TRY(up(rb.data(), rb.size() * sizeof(Box), (void**)&P.d_recv));
  PSC_TRY(up(sboff.data(), sboff.size() * sizeof(uint32_t), (void**)&P.d_send_boff));
  PSC_TRY(up(rboff.data(), rboff.size() * sizeof(uint32_t), (void**)&P.d_recv_boff));
  P.built = true;
  return 0;
}

// buffer layout per rank segment: [m][cells of the rank's boxes back to back]
template <bool PACK>
__global__ void k_halo_copy(GridDev G, float* __restrict__ F, long slot_len, int mb, int n_m,
                            const Box* __restrict__ boxes, const uint32_t* __restrict__ boff,
                            int n_boxes, uint32_t total, const uint32_t* __restrict__ seg_first,
                            const uint32_t* __restrict__ seg_cells, int n_seg,
                            float* __restrict__ buf)
{
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)total * n_m) {
    return;
  }
  uint32_t cell = (uint32_t)(idx % total);
  int m = (int)(idx / total);
  // box containing `cell`
  int lo = 0, hi = n_boxes;
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (boff[mid] <= cell) {
      lo = mid;
    } else {
      hi = mid;
    }
  }
  Box b = boxes[lo];
  uint32_t r = cell - boff[lo];
  int x = b.lo[0] + (int)(r % b.n[0]);
  r /= b.n[0];
  int y = b.lo[1] + (int)(r % b.n[1]);
  int z = b.lo[2] + (int)(r / b.n[1]);
  // rank segment of this box: segments are [n_m][seg_cells]
  int s = 0;
  while (s + 1 < n_seg && seg_first[s + 1] <= (uint32_t)lo) {
    s++;
  }
  uint32_t seg_base = boff[seg_first[s]];
  size_t bpos = (size_t)seg_base * n_m + (size_t)m * seg_cells[s] + (cell - seg_base);
  float* f = F + b.slot * slot_len + fld_off(G, mb + m, x, y, z);
  if (PACK) {
    buf[bpos] = *f;
  } else {
    *f = buf[bpos];
  }
}

} // namespace

int comm_unique_id(void* id128)
{
  PSC_TRY(nccl_load());
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  PSC_NCCL_TRY(g_nccl.GetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
  return 0;
}

int comm_init(Ctx* c, const void* id128)
{
  PSC_TRY(nccl_load());
  if (c->comm) {
    return fail("nccl_init called twice");
  }
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  Comm* cm = new Comm;
  ncclResult_t r = g_nccl.CommInitRank(&cm->comm, c->g.n_ranks, id, c->g.rank);
  if (r != ncclSuccess) {
    delete cm;
    return fail(std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r));
  }
  c->comm = cm;
  // NCCL connects the collective channels lazily (seconds on the first all-reduce): do it
  // here, not inside the first energy / checks call of a run
  double warm = 0.;
  return comm_allreduce_sum(c, &warm, 1);
}

void comm_destroy(Ctx* c)
{
  if (!c->comm) {
    return;
  }
  Comm* cm = c->comm;
  for (auto& P : cm->plan) {
    cudaFree(P.d_send);
    cudaFree(P.d_recv);
    cudaFree(P.d_send_boff);
    cudaFree(P.d_recv_boff);
  }
  cm->sendbuf.release();
  cm->recvbuf.release();
  cm->cnt.release();
  cm->prt_send.release();
  cm->prt_recv.release();
  cm->prt_final.release();
  cm->seg.release();
  if (cm->comm) {
    g_nccl.CommDestroy(cm->comm);
  }
  delete cm;
  c->comm = nullptr;
}

int comm_halo_exchange(Ctx* c, int id, int mb, int me, bool add)
{
  Comm* cm = c->comm;
  HaloPlan& P = cm->plan[add ? 1 : 0];
  if (!P.built) {
    PSC_TRY(build_plan(c, P, !add));
  }
  if (P.ranks.empty()) {
    return 0;
  }
  int n_m = me - mb;
  int n_seg = (int)P.ranks.size();
  PSC_TRY(cm->sendbuf.reserve((size_t)P.send_total * n_m * sizeof(float) + 16));
  PSC_TRY(cm->recvbuf.reserve((size_t)P.recv_total * n_m * sizeof(float) + 16));
  // segment tables: [send_first | send_cells | recv_first | recv_cells]
  PSC_TRY(cm->seg.reserve(4 * n_seg * sizeof(uint32_t)));
  uint32_t* d_seg = cm->seg.as<uint32_t>();
  PSC_CUDA_TRY(cudaMemcpyAsync(d_seg, P.send_first.data(), n_seg * 4, cudaMemcpyHostToDevice, c->stream));
  PSC_CUDA_TRY(cudaMemcpyAsync(d_seg + n_seg, P.send_cells.data(), n_seg * 4, cudaMemcpyHostToDevice, c->stream));
  PSC_CUDA_TRY(cudaMemcpyAsync(d_seg + 2 * n_seg, P.recv_first.data(), n_seg * 4, cudaMemcpyHostToDevice, c->stream));
  PSC_CUDA_TRY(cudaMemcpyAsync(d_seg + 3 * n_seg, P.recv_cells.data(), n_seg * 4, cudaMemcpyHostToDevice, c->stream));
  KernelScope ks(c, add ? "halo_add_xchg" : "halo_fill_xchg");
  if (P.send_total) {
    size_t n = (size_t)P.send_total * n_m;
    k_halo_copy<true><<<div_up(n, 256), 256, 0, c->stream>>>(
      c->gd, c->fld(id), c->fld_slot_len(id), mb, n_m, P.d_send, P.d_send_boff, P.n_send,
      P.send_total, d_seg, d_seg + n_seg, n_seg, cm->sendbuf.as<float>());
    c->n_launches++;
  }
  PSC_NCCL_TRY(g_nccl.GroupStart());
  size_t so = 0, ro = 0;
  for (int k = 0; k < n_seg; k++) {
    size_t sn = (size_t)P.send_cells[k] * n_m, rn = (size_t)P.recv_cells[k] * n_m;
    if (sn) {
      PSC_NCCL_TRY(g_nccl.Send(cm->sendbuf.as<float>() + so, sn, ncclFloat, P.ranks[k], cm->comm, c->stream));
    }
    if (rn) {
      PSC_NCCL_TRY(g_nccl.Recv(cm->recvbuf.as<float>() + ro, rn, ncclFloat, P.ranks[k], cm->comm, c->stream));
    }
    so += sn;
    ro += rn;
  }
  PSC_NCCL_TRY(g_nccl.GroupEnd());
  if (P.recv_total) {
    size_t n = (size_t)P.recv_total * n_m;
    k_halo_copy<false><<<div_up(n, 256), 256, 0, c->stream>>>(
      c->gd, c->fld(id), c->fld_slot_len(id), mb, n_m, P.d_recv, P.d_recv_boff, P.n_recv,
      P.recv_total, d_seg + 2 * n_seg, d_seg + 3 * n_seg, n_seg, cm->recvbuf.as<float>());
    c->n_launches++;
  }
  // the host-side segment tables above are members of the plan: no sync needed
  return check_launch(c, "halo_exchange");
}

static int allreduce(Ctx* c, double* v, int n, ncclRedOp_t op)
{
  Comm* cm = c->comm;
  PSC_TRY(cm->cnt.reserve(std::max(n, 64) * sizeof(double)));
  double* d = cm->cnt.as<double>();
  PSC_CUDA_TRY(cudaMemcpyAsync(d, v, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  PSC_NCCL_TRY(g_nccl.AllReduce(d, d, n, ncclDouble, op, cm->comm, c->stream));
  PSC_CUDA_TRY(cudaMemcpyAsync(v, d, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  PSC_CUDA_TRY(cudaStreamSynchronize(c->stream));
  return 0;
}

int comm_allreduce_max(Ctx* c, double* v, int n)
{
  return allreduce(c, v, n, ncclMax);
}

int comm_allreduce_sum(Ctx* c, double* v, int n)
{
  return allreduce(c, v, n, ncclSum);
}

// ---------------------------------------------------------------- particles

namespace
{

// gather the leaving particles into {xi4, pxi4} records; with `fixup` the boundary
// arithmetic (bnd_particles_impl.hxx:93-218) is applied here because the source store
// still holds the raw pushed positions
__global__ void k_pack_prts(GridDev G, uint32_t n, const uint32_t* __restrict__ src_idx,
                            const uint32_t* __restrict__ keys, uint32_t key_remote_base,
                            const pm::PatchBnd* __restrict__ pbs, bool fixup,
                            const float4* __restrict__ xi4, const float4* __restrict__ pxi4,
                            float4* __restrict__ out)
{
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) {
    uint32_t i = src_idx[j];
    float4 X = xi4[i], U = pxi4[i];
    if (fixup) {
      int p = ((keys[j] - key_remote_base) >> 5) % G.n_patches;
      float x[3] = {X.x, X.y, X.z}, u[3] = {U.x, U.y, U.z};
      int dir[3];
      bool drop;
      pm::PatchBnd pb = pbs[p];
      pm::bnd_classify(G.pc, pb, x, u, dir, drop);
      X = make_float4(x[0], x[1], x[2], X.w);
      U = make_float4(u[0], u[1], u[2], U.w);
    }
    out[2 * (size_t)j] = X;
    out[2 * (size_t)j + 1] = U;
  }
}

struct Seg
{
  uint32_t src, dst, n;
};

__global__ void k_unpack_prts(const Seg* __restrict__ segs, int n_segs, const float4* __restrict__ in,
                              float4* __restrict__ xo, float4* __restrict__ po)
{
  int s = blockIdx.y;
  if (s >= n_segs) {
    return;
  }
  Seg sg = segs[s];
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < sg.n; k += gridDim.x * blockDim.x) {
    xo[sg.dst + k] = in[2 * (size_t)(sg.src + k)];
    po[sg.dst + k] = in[2 * (size_t)(sg.src + k) + 1];
  }
}

} // namespace

// Ships the remote tail of the sorted leaver list.  d_keys are
// key_remote_base + ((rank * n_patches + sender patch) * 32 + sender dir idx), ascending.
// Returns the received particles ordered by (destination patch, sender rank, sender
// patch, sender direction) -- ddc_particles.hxx:456-468 -- and the count per patch.
int comm_exchange_particles(Ctx* c, const float4* xi_src, const float4* pxi_src,
                            const uint32_t* d_src_idx, const uint32_t* d_keys, uint32_t n_remote,
                            uint32_t key_remote_base, bool fixup,
                            std::vector<uint32_t>& n_recv_by_patch, float4** xi_recv,
                            float4** pxi_recv)
{
  Comm* cm = c->comm;
  const GridHost& g = c->g;
  const int np = g.n_patches;
  *xi_recv = *pxi_recv = nullptr;
  // entry tables, both directions, enumerated (patch asc, dir asc)
  struct Entry
  {
    int rank, patch /* sender-local or sender-global */, di, dest /* local patch */;
    uint32_t n;
  };
  std::vector<Entry> se, re;
  std::vector<int> ranks;
  for (int r = 0; r < g.n_ranks; r++) {
    if (r == g.rank) {
      continue;
    }
    size_t s0 = se.size(), r0 = re.size();
    for (int p = 0; p < np; p++) {
      for (int di = 0; di < 27; di++) {
        if (di != 13 && c->h_nei_patch[p * 27 + di] == -2 - r) {
          se.push_back({r, p, di, -1, 0});
        }
      }
    }
    for (int gp = g.patch_off_by_rank[r]; gp < g.patch_off_by_rank[r + 1]; gp++) {
      for (int di = 0; di < 27; di++) {
        if (di == 13) {
          continue;
        }
        int dir[3] = {di % 3 - 1, (di / 3) % 3 - 1, di / 9 - 1};
        int ngp = g.neighbor_patch(gp, dir);
        if (ngp >= 0 && g.rank_of_patch(ngp) == g.rank) {
          re.push_back({r, gp, di, ngp - g.patch_begin, 0});
        }
      }
    }
    if (se.size() > s0 || re.size() > r0) {
      ranks.push_back(r);
    }
  }
  if (ranks.empty()) {
    return 0;
  }
  // counts of what we send
  std::vector<uint32_t> h_keys(n_remote);
  if (n_remote) {
    PSC_CUDA_TRY(cudaMemcpyAsync(h_keys.data(), d_keys, n_remote * sizeof(uint32_t),
                                 cudaMemcpyDeviceToHost, c->stream));
    PSC_CUDA_TRY(cudaStreamSynchronize(c->stream));
  }
  {
    size_t e = 0;
    for (uint32_t k : h_keys) {
      uint32_t v = k - key_remote_base;
      int di = v & 31, p = (v >> 5) % np, r = (v >> 5) / np;
      while (e < se.size() && !(se[e].rank == r && se[e].patch == p && se[e].di == di)) {
        e++;
      }
      if (e == se.size()) {
        return fail("particle exchange: leaver key without a send entry");
      }
      se[e].n++;
    }
  }
  // exchange counts
  size_t n_cnt = se.size() + re.size();
  PSC_TRY(cm->cnt.reserve(std::max<size_t>(n_cnt, 16) * sizeof(uint32_t)));
  uint32_t* d_cnt = cm->cnt.as<uint32_t>();
  std::vector<uint32_t> h_cnt(n_cnt, 0);
  for (size_t e = 0; e < se.size(); e++) {
    h_cnt[e] = se[e].n;
  }
  PSC_CUDA_TRY(cudaMemcpyAsync(d_cnt, h_cnt.data(), se.size() * sizeof(uint32_t),
                               cudaMemcpyHostToDevice, c->stream));
  PSC_NCCL_TRY(g_nccl.GroupStart());
  {
    size_t so = 0, ro = se.size();
    for (int r : ranks) {
      size_t sn = 0, rn = 0;
      for (auto& e : se) {
        sn += e.rank == r;
      }
      for (auto& e : re) {
        rn += e.rank == r;
      }
      if (sn) {
        PSC_NCCL_TRY(g_nccl.Send(d_cnt + so, sn, ncclUint32, r, cm->comm, c->stream));
      }
      if (rn) {
        PSC_NCCL_TRY(g_nccl.Recv(d_cnt + ro, rn, ncclUint32, r, cm->comm, c->stream));
      }
      so += sn;
      ro += rn;
    }
  }
  PSC_NCCL_TRY(g_nccl.GroupEnd());
  PSC_CUDA_TRY(cudaMemcpyAsync(h_cnt.data() + se.size(), d_cnt + se.size(),
                               re.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  PSC_CUDA_TRY(cudaStreamSynchronize(c->stream));
  uint32_t n_recv = 0;
  for (size_t e = 0; e < re.size(); e++) {
    re[e].n = h_cnt[se.size() + e];
    n_recv += re[e].n;
  }
  // payload: records are {xi4, pxi4} pairs in leaver-list order (= rank, patch, dir)
  PSC_TRY(cm->prt_send.reserve(std::max<size_t>(n_remote, 1) * 32));
  PSC_TRY(cm->prt_recv.reserve(std::max<size_t>(n_recv, 1) * 32));
  if (n_remote) {
    k_pack_prts<<<div_up(n_remote, 256), 256, 0, c->stream>>>(
      c->gd, n_remote, d_src_idx, d_keys, key_remote_base, c->d_patch_bnd, fixup, xi_src, pxi_src,
      cm->prt_send.as<float4>());
    c->n_launches++;
  }
  PSC_NCCL_TRY(g_nccl.GroupStart());
  {
    size_t so = 0, ro = 0;
    for (int r : ranks) {
      size_t sn = 0, rn = 0;
      for (auto& e : se) {
        sn += e.rank == r ? e.n : 0;
      }
      for (auto& e : re) {
        rn += e.rank == r ? e.n : 0;
      }
      if (sn) {
        PSC_NCCL_TRY(g_nccl.Send(cm->prt_send.as<float>() + so * 8, sn * 8, ncclFloat, r, cm->comm, c->stream));
      }
      if (rn) {
        PSC_NCCL_TRY(g_nccl.Recv(cm->prt_recv.as<float>() + ro * 8, rn * 8, ncclFloat, r, cm->comm, c->stream));
      }
      so += sn;
      ro += rn;
    }
  }
  PSC_NCCL_TRY(g_nccl.GroupEnd());
  // bucket by destination patch, keeping (rank, patch, dir) order inside a patch
  n_recv_by_patch.assign(np, 0);
  for (auto& e : re) {
    n_recv_by_patch[e.dest] += e.n;
  }
  if (n_recv == 0) {
    return check_launch(c, "exchange_particles");
  }
  std::vector<uint32_t> dst_off(np + 1, 0);
  for (int p = 0; p < np; p++) {
    dst_off[p + 1] = dst_off[p] + n_recv_by_patch[p];
  }
  std::vector<Seg> segs;
  {
    std::vector<uint32_t> fill(dst_off.begin(), dst_off.end() - 1);
    uint32_t src = 0;
    for (auto& e : re) {
      if (e.n) {
        segs.push_back({src, fill[e.dest], e.n});
        fill[e.dest] += e.n;
        src += e.n;
      }
    }
  }
  PSC_TRY(cm->prt_final.reserve((size_t)n_recv * 32));
  PSC_TRY(cm->seg.reserve(std::max<size_t>(segs.size(), 4) * sizeof(Seg) + 64));
  PSC_CUDA_TRY(cudaMemcpyAsync(cm->seg.p, segs.data(), segs.size() * sizeof(Seg),
                               cudaMemcpyHostToDevice, c->stream));
  float4* xo = cm->prt_final.as<float4>();
  float4* po = xo + n_recv;
  dim3 grid(16, (unsigned)segs.size());
  k_unpack_prts<<<grid, 256, 0, c->stream>>>(cm->seg.as<Seg>(), (int)segs.size(),
                                            cm->prt_recv.as<float4>(), xo, po);
  c->n_launches++;
  PSC_CUDA_TRY(cudaStreamSynchronize(c->stream)); // segs is a host temporary
  *xi_recv = xo;
  *pxi_recv = po;
  return check_launch(c, "exchange_particles");
}

// ---------------------------------------------------------------- balance

// best_mapping: recursive bisection of the patch list by load, at least one patch per
// rank (libpsc/psc_balance/psc_balance_impl.hxx:99-160)
static void bisect(const std::vector<double>& loads, std::vector<int>& out, int r0, int r1, int p0,
                   int p1)
{
  if (r1 == r0 + 1) {
    out[r0] = p1 - p0;
    return;
  }
  int rm = (r0 + r1) / 2;
  double total = 0.;
  for (int p = p0; p < p1; p++) {
    total += loads[p];
  }
  double target = (total * (rm - r0)) / (r1 - r0);
  double load = 0.;
  int pm_ = p0;
  while (pm_ < p1) {
    double prev = load;
    load += loads[pm_];
    pm_++;
    if (load > target) {
      // closer to the target one patch earlier?
      if (target - prev < load - target && pm_) {
        pm_--;
        load = prev;
      }
      break;
    }
  }
  if (pm_ - p0 < rm - r0) {
    pm_ = p0 + (rm - r0);
  }
  if (p1 - pm_ < r1 - rm) {
    pm_ = p1 - (r1 - rm);
  }
  bisect(loads, out, r0, rm, p0, pm_);
  bisect(loads, out, rm, r1, pm_, p1);
}

std::vector<int> best_mapping(const std::vector<double>& capability,
                              const std::vector<double>& loads)
{
  int n_ranks = (int)capability.size();
  std::vector<int> out(n_ranks, 0);
  if ((int)loads.size() < n_ranks) {
    throw std::runtime_error("best_mapping: fewer patches than ranks (psc_balance_impl.hxx:104-105)");
  }
  bisect(loads, out, 0, n_ranks, 0, (int)loads.size());
  return out;
}

// Balance_::operator() (psc_balance_impl.hxx:770-1026): loads -> best_mapping -> new
// contiguous patch ranges -> whole patches (particles + every field array) move to
// their new owner.  The reference round-trips everything through the host and MPI
// (communicate_ctx, :353-760); here every rank knows all per-patch particle counts
// after one all-reduce, so the moves are direct device-to-device ncclSend/ncclRecv with
// no headers: both sides walk the global patch list in ascending order.
int balance(Ctx* c, double factor_fields, int* changed)
{
  *changed = 0;
  const GridHost g = c->g; // the old decomposition (copy)
  if (g.n_ranks == 1) {
    return 0; // one rank owns every patch: nothing to redistribute
  }
  if (!c->comm) {
    return fail("balance: psc_b200_nccl_init has not been called");
  }
  Comm* cm = c->comm;
  const int NG = g.n_patches_global, me = g.rank;
  // get_loads (:223-269) and the particle count of every global patch
  std::vector<double> v(2 * (size_t)NG, 0.);
  for (int p = 0; p < g.n_patches; p++) {
    double n = (double)(c->h_off[p + 1] - c->h_off[p]);
    v[g.patch_begin + p] = n + factor_fields * g.n_cells;
    v[NG + g.patch_begin + p] = n;
  }
  PSC_TRY(comm_allreduce_sum(c, v.data(), 2 * NG));
  std::vector<double> loads(v.begin(), v.begin() + NG);
  std::vector<int> new_n = best_mapping(std::vector<double>(g.n_ranks, 1.), loads);
  std::vector<int> new_off(g.n_ranks + 1, 0);
  for (int r = 0; r < g.n_ranks; r++) {
    new_off[r + 1] = new_off[r] + new_n[r];
  }
  if (new_off == g.patch_off_by_rank) {
    return 0;
  }
  *changed = 1;
  auto owner = [](const std::vector<int>& off, int gp) {
    return (int)(std::upper_bound(off.begin(), off.end(), gp) - off.begin()) - 1;
  };
  const int nb = new_off[me], nnp = new_off[me + 1] - new_off[me];
  std::vector<uint32_t> noff(nnp + 1, 0);
  for (int k = 0; k < nnp; k++) {
    noff[k + 1] = noff[k] + (uint32_t)v[NG + nb + k];
  }
  const uint32_t n_new = noff[nnp];
  PSC_TRY(prts_reserve(c, std::max<size_t>(n_new, c->n_prts)));
  float4 *xs = c->xi(), *ps = c->pxi(), *xd = c->xi_alt(), *pd = c->pxi_alt();

  // new decomposition, tables and (zeroed) field arrays
  std::vector<FieldArr> old_flds = c->flds;
  c->g.patch_off_by_rank = new_off;
  c->g.patch_begin = nb;
  c->g.n_patches = nnp;
  c->gd.n_patches = nnp;
  cudaFree(c->d_patch_bnd);
  cudaFree(c->d_nei_patch);
  cudaFree(c->d_nei_slot);
  cudaFree(c->d_add_order);
  c->d_patch_bnd = nullptr, c->d_nei_patch = nullptr, c->d_nei_slot = nullptr, c->d_add_order = nullptr;
  PSC_TRY(build_patch_tables(c));
  c->flds.clear();
  for (const FieldArr& f : old_flds) {
    int id;
    PSC_TRY(flds_create(c, f.n_comps, &id));
  }

  // move the patches
  PSC_NCCL_TRY(g_nccl.GroupStart());
  for (int gp = 0; gp < NG; gp++) {
    const int ro = owner(g.patch_off_by_rank, gp), rn = owner(new_off, gp);
    if (ro != me && rn != me) {
      continue;
    }
    const size_t n = (size_t)v[NG + gp];
    const int po = gp - g.patch_begin, pn = gp - nb; // old / new local index
    if (ro == me && rn == me) {
      if (n) {
        PSC_CUDA_TRY(cudaMemcpyAsync(xd + noff[pn], xs + c->h_off[po], n * sizeof(float4),
                                     cudaMemcpyDeviceToDevice, c->stream));
        PSC_CUDA_TRY(cudaMemcpyAsync(pd + noff[pn], ps + c->h_off[po], n * sizeof(float4),
                                     cudaMemcpyDeviceToDevice, c->stream));
      }
      for (size_t f = 0; f < old_flds.size(); f++) {
        size_t len = (size_t)g.fld_len * old_flds[f].n_comps;
        PSC_CUDA_TRY(cudaMemcpyAsync(c->flds[f].d + pn * len, old_flds[f].d + po * len,
                                     len * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
      }
    } else if (ro == me) {
      if (n) {
        PSC_NCCL_TRY(g_nccl.Send(xs + c->h_off[po], n * 4, ncclFloat, rn, cm->comm, c->stream));
        PSC_NCCL_TRY(g_nccl.Send(ps + c->h_off[po], n * 4, ncclFloat, rn, cm->comm, c->stream));
      }
      for (size_t f = 0; f < old_flds.size(); f++) {
        size_t len = (size_t)g.fld_len * old_flds[f].n_comps;
        PSC_NCCL_TRY(g_nccl.Send(old_flds[f].d + po * len, len, ncclFloat, rn, cm->comm, c->stream));
      }
    } else {
      if (n) {
        PSC_NCCL_TRY(g_nccl.Recv(xd + noff[pn], n * 4, ncclFloat, ro, cm->comm, c->stream));
        PSC_NCCL_TRY(g_nccl.Recv(pd + noff[pn], n * 4, ncclFloat, ro, cm->comm, c->stream));
      }
      for (size_t f = 0; f < old_flds.size(); f++) {
        size_t len = (size_t)g.fld_len * old_flds[f].n_comps;
        PSC_NCCL_TRY(g_nccl.Recv(c->flds[f].d + pn * len, len, ncclFloat, ro, cm->comm, c->stream));
      }
    }
  }
  PSC_NCCL_TRY(g_nccl.GroupEnd());
  PSC_CUDA_TRY(cudaStreamSynchronize(c->stream));
  for (FieldArr& f : old_flds) {
    cudaFree(f.d);
  }

  // the store now lives in the other buffer, in the new patch order
  c->cur ^= 1;
  c->n_prts = n_new;
  c->h_off = noff;
  cudaFree(c->d_off);
  cudaFree(c->d_cell_off);
  cudaFree(c->d_cell_off_alt);
  const size_t nct = (size_t)g.n_cells * nnp;
  PSC_CUDA_TRY(cudaMalloc(&c->d_off, (nnp + 1) * sizeof(uint32_t)));
  PSC_CUDA_TRY(cudaMalloc(&c->d_cell_off, (nct + 1) * sizeof(uint32_t)));
  PSC_CUDA_TRY(cudaMalloc(&c->d_cell_off_alt, (nct + 1) * sizeof(uint32_t)));
  PSC_CUDA_TRY(cudaMemset(c->d_cell_off, 0, (nct + 1) * sizeof(uint32_t)));
  PSC_TRY(prts_upload_off(c));
  c->sorted = false;
  c->pushed_from_sorted = false;
  c->counts_valid = false;
  c->want_counts = false;
  c->rf_built = false;
  // halo plans follow the tables (psc_balance_generation_cnt, bnd_particles_impl.hxx:236-239)
  for (auto& P : cm->plan) {
    cudaFree(P.d_send);
    cudaFree(P.d_recv);
    cudaFree(P.d_send_boff);
    cudaFree(P.d_recv_boff);
    P = HaloPlan{};
  }
  // ghost cells of the moved patches are rebuilt by the next fill; proxies start empty
  return 0;
}

} // namespace psc_b200
