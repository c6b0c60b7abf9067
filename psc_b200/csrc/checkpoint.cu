// psc_b200: checkpoint / restart of the device state.
//
// What: write_checkpoint / read_checkpoint (src/include/checkpoint.hxx:14-82): "grid", "mprts"
// and "mflds" go to one file set per checkpoint, a restarted run continues as if it had never
// stopped.  The reference writes ADIOS2 BP through kg::io (not available here, and an
// external dependency there); the variable decomposition is kept so that a reader on the PSC
// side maps one to one -- particles as `size_by_patch` plus one array per record component
// (particles_simple.inl:113-131, ForComponents: x y z ux uy uz kind qni_wni), fields as
// ib / im plus the component-major patch arrays (fields3d.inl:18-57) -- in a flat
// self-describing binary, one file per rank:  <path>.<rank>
//
//   u64 magic "PSCB200C", u32 version, u32 header bytes
//   grid:   gdims[3] np[3] ldims[3] ibn[3] (i32), length[3] corner[3] dt fnqs eta (f64), n_kinds (i32),
//           q[10] m[10] (f64), bc (12 x i32), rank n_ranks patch_begin n_patches (i32), timestep (i64)
//   mprts:  size_by_patch[n_patches] (u32), then x, y, z, ux, uy, uz (f32[n]), kind (i32[n]), qni_wni (f32[n])
//   mflds:  n_comps (i32), ib[3] im[3] (i32), data f32[n_patches][n_comps][im2][im1][im0]
#include "dev_util.cuh"

#include <cstdio>
#include <cstring>
#include <memory>

namespace psc_b200
{

namespace
{

constexpr uint64_t CKPT_MAGIC = 0x43303032'42435350ull; // "PSCB200C" little endian
constexpr uint32_t CKPT_VERSION = 1;

struct CkptHeader
{
  uint64_t magic;
  uint32_t version, bytes;
  int32_t gdims[3], np[3], ldims[3], ibn[3];
  double length[3], corner[3], dt, fnqs, eta;
  int32_t n_kinds;
  int32_t pad0;
  double q[PSC_B200_MAX_KINDS], m[PSC_B200_MAX_KINDS];
  int32_t bc[12];
  int32_t rank, n_ranks, patch_begin, n_patches;
  int64_t timestep;
  uint64_t n_prts;
};

void fill_header(const Ctx* c, CkptHeader& h, int64_t timestep)
{
  const GridHost& g = c->g;
  std::memset(&h, 0, sizeof(h));
  h.magic = CKPT_MAGIC;
  h.version = CKPT_VERSION;
  h.bytes = sizeof(h);
  for (int d = 0; d < 3; d++) {
    h.gdims[d] = g.desc.gdims[d], h.np[d] = g.desc.np[d], h.ldims[d] = g.ldims[d], h.ibn[d] = g.ibn[d];
    h.length[d] = g.desc.length[d], h.corner[d] = g.desc.corner[d];
    h.bc[d] = g.desc.bc_fld_lo[d], h.bc[3 + d] = g.desc.bc_fld_hi[d];
    h.bc[6 + d] = g.desc.bc_prt_lo[d], h.bc[9 + d] = g.desc.bc_prt_hi[d];
  }
  h.dt = g.desc.dt, h.fnqs = g.desc.fnqs, h.eta = g.desc.eta;
  h.n_kinds = g.desc.n_kinds;
  for (int k = 0; k < g.desc.n_kinds; k++) {
    h.q[k] = g.desc.q[k], h.m[k] = g.desc.m[k];
  }
  h.rank = g.rank, h.n_ranks = g.n_ranks, h.patch_begin = g.patch_begin, h.n_patches = g.n_patches;
  h.timestep = timestep;
  h.n_prts = c->n_prts;
}

struct File
{
  FILE* f = nullptr;
  ~File()
  {
    if (f) {
      fclose(f);
    }
  }
};

} // namespace

int checkpoint_write(Ctx* c, const char* path, int64_t timestep)
{
  if (!path) {
    return fail("checkpoint_write: null path");
  }
  const std::string name = std::string(path) + "." + std::to_string(c->g.rank);
  File fp;
  fp.f = fopen(name.c_str(), "wb");
  if (!fp.f) {
    return fail("checkpoint_write: cannot open " + name);
  }
  CkptHeader h;
  fill_header(c, h, timestep);
  bool ok = fwrite(&h, sizeof(h), 1, fp.f) == 1;
  // ---- mprts
  const size_t n = c->n_prts;
  std::vector<uint32_t> sbp(c->g.n_patches);
  for (int p = 0; p < c->g.n_patches; p++) {
    sbp[p] = c->h_off[p + 1] - c->h_off[p];
  }
  ok = ok && fwrite(sbp.data(), sizeof(uint32_t), sbp.size(), fp.f) == sbp.size();
  {
    std::vector<float4> hx(n), hp(n);
    if (n) {
      PSC_CUDA_TRY(cudaMemcpyAsync(hx.data(), c->xi(), n * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
      PSC_CUDA_TRY(cudaMemcpyAsync(hp.data(), c->pxi(), n * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
      PSC_CUDA_TRY(cudaStreamSynchronize(c->stream));
    }
    std::vector<float> comp(n);
    auto put = [&](auto get) {
      for (size_t i = 0; i < n; i++) {
        comp[i] = get(i);
      }
      ok = ok && fwrite(comp.data(), sizeof(float), n, fp.f) == n;
    };
    put([&](size_t i) { return hx[i].x; });
    put([&](size_t i) { return hx[i].y; });
    put([&](size_t i) { return hx[i].z; });
    put([&](size_t i) { return hp[i].x; });
    put([&](size_t i) { return hp[i].y; });
    put([&](size_t i) { return hp[i].z; });
    put([&](size_t i) { return hx[i].w; }); // kind: the int's bits
    put([&](size_t i) { return hp[i].w; });
  }
  // ---- mflds (field 0: the 9-component state)
  {
    const int32_t n_comps = c->flds[0].n_comps;
    int32_t ibim[6];
    for (int d = 0; d < 3; d++) {
      ibim[d] = -c->g.ibn[d];
      ibim[3 + d] = c->g.im[d];
    }
    ok = ok && fwrite(&n_comps, sizeof(n_comps), 1, fp.f) == 1 && fwrite(ibim, sizeof(ibim), 1, fp.f) == 1;
    const size_t len = (size_t)c->g.n_patches * n_comps * c->gd.fld_len;
    std::vector<float> hf(len);
    PSC_TRY(flds_download(c, 0, 0, n_comps, hf.data()));
    ok = ok && fwrite(hf.data(), sizeof(float), len, fp.f) == len;
  }
  if (!ok) {
    return fail("checkpoint_write: short write to " + name);
  }
  return 0;
}

int checkpoint_read(Ctx* c, const char* path, int64_t* timestep)
{
  if (!path) {
    return fail("checkpoint_read: null path");
  }
  const std::string name = std::string(path) + "." + std::to_string(c->g.rank);
  File fp;
  fp.f = fopen(name.c_str(), "rb");
  if (!fp.f) {
    return fail("checkpoint_read: cannot open " + name);
  }
  CkptHeader h, want;
  if (fread(&h, sizeof(h), 1, fp.f) != 1 || h.magic != CKPT_MAGIC || h.version != CKPT_VERSION ||
      h.bytes != sizeof(h)) {
    return fail("checkpoint_read: " + name + " is not a psc_b200 checkpoint of this version");
  }
  fill_header(c, want, h.timestep);
  want.n_prts = h.n_prts;
  if (std::memcmp(&h, &want, sizeof(h)) != 0) {
    // (read_checkpoint constructs its containers from the grid it read, checkpoint.hxx:66-70;
    // here the caller constructs the context, so the grids must agree)
    return fail("checkpoint_read: the checkpoint was written for a different grid / decomposition");
  }
  const size_t n = h.n_prts;
  std::vector<uint32_t> sbp(c->g.n_patches);
  bool ok = fread(sbp.data(), sizeof(uint32_t), sbp.size(), fp.f) == sbp.size();
  size_t tot = 0;
  for (uint32_t v : sbp) {
    tot += v;
  }
  if (!ok || tot != n) {
    return fail("checkpoint_read: size_by_patch does not add up");
  }
  {
    std::vector<float4> hx(n), hp(n);
    std::vector<float> comp(n);
    auto get = [&](auto set) {
      ok = ok && fread(comp.data(), sizeof(float), n, fp.f) == n;
      for (size_t i = 0; i < n; i++) {
        set(i, comp[i]);
      }
    };
    get([&](size_t i, float v) { hx[i].x = v; });
    get([&](size_t i, float v) { hx[i].y = v; });
    get([&](size_t i, float v) { hx[i].z = v; });
    get([&](size_t i, float v) { hp[i].x = v; });
    get([&](size_t i, float v) { hp[i].y = v; });
    get([&](size_t i, float v) { hp[i].z = v; });
    get([&](size_t i, float v) { hx[i].w = v; });
    get([&](size_t i, float v) { hp[i].w = v; });
    if (!ok) {
      return fail("checkpoint_read: short read (particles)");
    }
    PSC_TRY(prts_reserve(c, n));
    if (n) {
      PSC_CUDA_TRY(cudaMemcpyAsync(c->xi(), hx.data(), n * sizeof(float4), cudaMemcpyHostToDevice, c->stream));
      PSC_CUDA_TRY(cudaMemcpyAsync(c->pxi(), hp.data(), n * sizeof(float4), cudaMemcpyHostToDevice, c->stream));
      PSC_CUDA_TRY(cudaStreamSynchronize(c->stream));
    }
    c->n_prts = (uint32_t)n;
    c->h_off[0] = 0;
    for (int p = 0; p < c->g.n_patches; p++) {
      c->h_off[p + 1] = c->h_off[p] + sbp[p];
    }
    // the order inside the patches is whatever was written; the next sort re-derives the
    // cell offsets (a stable sort of a cell-ordered store is the identity)
    c->sorted = c->pushed_from_sorted = c->counts_valid = false;
    c->gapped = false;
    PSC_TRY(prts_upload_off(c));
  }
  {
    int32_t n_comps = 0, ibim[6];
    ok = fread(&n_comps, sizeof(n_comps), 1, fp.f) == 1 && fread(ibim, sizeof(ibim), 1, fp.f) == 1;
    if (!ok || n_comps != c->flds[0].n_comps) {
      return fail("checkpoint_read: field container mismatch");
    }
    const size_t len = (size_t)c->g.n_patches * n_comps * c->gd.fld_len;
    std::vector<float> hf(len);
    if (fread(hf.data(), sizeof(float), len, fp.f) != len) {
      return fail("checkpoint_read: short read (fields)");
    }
    PSC_TRY(flds_upload(c, 0, 0, n_comps, hf.data()));
  }
  if (timestep) {
    *timestep = h.timestep;
  }
  return 0;
}

} // namespace psc_b200
