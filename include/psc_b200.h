/* psc_b200 -- C ABI of the B200-native (sm_100a) implementation of PSC's per-timestep
 * particle-in-cell hot path.  This is the drop-in boundary: the C++ wrapper types in
 * include/psc_b200/psc_config_b200.hxx (MparticlesB200, MfieldsStateB200,
 * PushParticlesB200, SortB200, BndParticlesB200, BndB200, BndFieldsB200,
 * PushFieldsB200, MarderB200, ChecksB200 -> PscConfig1vbecB200<Dim>) are thin
 * shells over these entry points, and so are the Python test/bench bindings.
 *
 * Conventions
 *   - one opaque context per process/GPU ("rank"); not thread-safe per context
 *   - every call returns 0 on success, non-zero on error; psc_b200_last_error()
 *     gives the text.  No exceptions cross the boundary.  (PSC itself aborts on
 *     error: libpsc/bits.hxx:35-40, cuda/cuda_bits.h:32-40.)
 *   - all device memory is owned by the context; host pointers are borrowed for
 *     the duration of the call only
 *   - calls are stream-ordered on the context's stream; calls that return data to
 *     the host synchronise, everything else is asynchronous until psc_b200_sync
 *   - there is NO CPU fallback: every operator launches CUDA kernels and fails
 *     with an error if no device is present
 *
 * Data layouts are PSC's (so a deck's data moves with plain copies):
 *   fields    float [p][m][iz][iy][ix], ix fastest; dims ldims + 2*ibn, lower bound
 *             -ibn (src/include/fields3d.hxx:29-32,284-291); m = JXI..HZ
 *             (src/include/psc.h:26-38)
 *   particles 32-byte records {float x[3]; float u[3]; int kind; float qni_wni}
 *             = ParticleSimple<float> (src/include/particle_simple.hxx:10-42),
 *             x patch-relative in physical units, patch p = [off[p], off[p+1])
 *   patches   "bydim" order p = (pz*npy + py)*npx + px
 *             (src/libmrc/src/mrc_domain_lib.c:21-35); a rank owns a contiguous
 *             range of that list (src/libmrc/src/mrc_domain_multi.c:162-193)
 */
#ifndef PSC_B200_H
#define PSC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PSC_B200_MAX_KINDS 10 /* push_particles_1vb.hxx:11 */
#define PSC_B200_NR_FIELDS 9

/* field components, src/include/psc.h:26-38 */
enum
{
  PSC_B200_JXI,
  PSC_B200_JYI,
  PSC_B200_JZI,
  PSC_B200_EX,
  PSC_B200_EY,
  PSC_B200_EZ,
  PSC_B200_HX,
  PSC_B200_HY,
  PSC_B200_HZ
};

/* src/include/grid/BC.h:8-24 */
enum
{
  PSC_B200_BND_FLD_OPEN,
  PSC_B200_BND_FLD_PERIODIC,
  PSC_B200_BND_FLD_CONDUCTING_WALL,
  PSC_B200_BND_FLD_ABSORBING
};
enum
{
  PSC_B200_BND_PRT_REFLECTING,
  PSC_B200_BND_PRT_PERIODIC,
  PSC_B200_BND_PRT_ABSORBING,
  PSC_B200_BND_PRT_OPEN
};

/* which 1vb deposit: src/psc_config.hxx:47-72 picks VAR1 for dim_yz and SPLIT for
 * dim_xyz in production; PSC's tests use SPLIT for yz too. */
enum
{
  PSC_B200_DEPOSIT_DEFAULT = -1,
  PSC_B200_DEPOSIT_VAR1 = 0,
  PSC_B200_DEPOSIT_SPLIT = 1
};

typedef struct psc_b200_ctx psc_b200_ctx;

/* Grid_t as far as the hot path reads it (src/include/grid.hxx:68-101,141-160;
 * grid/domain.hxx:25-47; grid.hxx:265-293 Normalization).  All reals are double
 * here and narrowed inside exactly where PSC narrows them. */
typedef struct
{
  int gdims[3];     /* global cells; gdims[d] == 1 marks an invariant direction */
  int np[3];        /* patches per direction */
  double length[3]; /* domain size */
  double corner[3]; /* lower corner */
  double dt;
  double fnqs; /* grid.norm.fnqs */
  double eta;  /* grid.norm.eta  */
  int n_kinds;
  double q[PSC_B200_MAX_KINDS];
  double m[PSC_B200_MAX_KINDS];
  int bc_fld_lo[3], bc_fld_hi[3];
  int bc_prt_lo[3], bc_prt_hi[3];
  int deposit; /* PSC_B200_DEPOSIT_* */
  /* decomposition: global patch range owned by each rank; n_patches_by_rank may be
   * NULL = uniform split like mrc_domain_multi.c:181-189 */
  int rank, n_ranks;
  const int* n_patches_by_rank;
  int device;           /* CUDA device ordinal, -1 = current device */
  uint64_t max_n_prts;  /* particle capacity of this rank, 0 = grow on demand */
} psc_b200_grid_desc;

const char* psc_b200_last_error(void);
const char* psc_b200_version(void);

/* Grid_t ctor + Mparticles(grid) + MfieldsState(grid)  (psc_bubble_yz.cxx:117-147,290-291) */
int psc_b200_create(const psc_b200_grid_desc* desc, psc_b200_ctx** ctx);
void psc_b200_destroy(psc_b200_ctx* ctx);
int psc_b200_sync(psc_b200_ctx* ctx);

int psc_b200_n_patches(const psc_b200_ctx* ctx);   /* local */
int psc_b200_patch_begin(const psc_b200_ctx* ctx); /* global index of local patch 0 */
int psc_b200_get_ldims(const psc_b200_ctx* ctx, int ldims[3], int ibn[3]);

/* ---- Mparticles (src/include/particles_simple.hxx:123-247, injector_simple.hxx) ---- */
/* replace all particles: n_by_patch[n_patches], records patch after patch */
int psc_b200_mprts_set(psc_b200_ctx* ctx, const void* prts_aos32,
                       const uint32_t* n_by_patch);
/* append per patch (InjectorSimple::Patch::operator(), injector_simple.hxx:24-43,
 * already converted to patch-relative float records) */
int psc_b200_mprts_inject(psc_b200_ctx* ctx, const void* prts_aos32,
                          const uint32_t* n_by_patch);
int psc_b200_mprts_size(psc_b200_ctx* ctx, uint64_t* n_total);
int psc_b200_mprts_size_by_patch(psc_b200_ctx* ctx, uint32_t* n_by_patch);
/* MparticlesBase::get_as<MparticlesSingle> (src/include/particles.hxx:35-54) */
int psc_b200_mprts_get(psc_b200_ctx* ctx, void* prts_aos32, uint32_t* off);
/* synthetic loader on the device (uniform-in-cell thermal plasma):
 * ppc particles per cell for each kind, u ~ N(0, vth[kind]), w = 1, sorted by cell */
int psc_b200_mprts_setup_thermal(psc_b200_ctx* ctx, int ppc, const double* vth,
                                 uint64_t seed);
/* ... with a density profile by patch: local patch p gets ppc_by_patch[p] particles per cell
 * and kind (the non-uniform state the load-balancing measurement of bench.py starts from) */
int psc_b200_mprts_setup_thermal_by_patch(psc_b200_ctx* ctx, const int* ppc_by_patch,
                                          const double* vth, uint64_t seed);

/* ---- MfieldsState / Mfields (src/include/fields3d.hxx:321-464) ---- */
/* field 0 is the 9-component state; further scratch fields via _create */
int psc_b200_mflds_create(psc_b200_ctx* ctx, int n_comps, int* field_id);
int psc_b200_mflds_upload(psc_b200_ctx* ctx, int field_id, int mb, int me,
                          const float* host);
int psc_b200_mflds_download(psc_b200_ctx* ctx, int field_id, int mb, int me,
                            float* host);
int psc_b200_mflds_zero(psc_b200_ctx* ctx, int field_id, int mb, int me);
/* uniform value everywhere incl. ghosts (setupFields with a constant lambda) */
int psc_b200_mflds_fill(psc_b200_ctx* ctx, int field_id, int m, float value);

/* ---- OutputFieldsItem's hand-off (src/include/output_fields.hxx:150-236): what happens to an
 * item (the state fields = Item_jeh, or a moment container) between the operator that produced
 * it and the writer stays on the device except for the one interior copy a write needs.
 *   mflds_add:   y[y_mb + m] += x[x_mb + m], m < n_comps   (:203  tfd = tfd + pfd, float + float)
 *   mflds_scale: y[m] = float(a * double(y[m]))            (:221  (1. / naccum) * tfd)
 *   mflds_download_interior: host[p][m - mb][k][j][i] over the patches' own cells
 *                            (psc::mflds::interior, :179) -- 1 / (1 + 2 ibn / ldims)^3 of the bytes
 *                            of mflds_download ---- */
int psc_b200_mflds_add(psc_b200_ctx* ctx, int y_field_id, int y_mb, int x_field_id, int x_mb, int n_comps);
int psc_b200_mflds_scale(psc_b200_ctx* ctx, int field_id, int mb, int me, double a);
int psc_b200_mflds_download_interior(psc_b200_ctx* ctx, int field_id, int mb, int me, float* host);

/* ---- PushParticles::push_mprts (push_particles_1vb.hxx:27-84) ---- */
int psc_b200_push_mprts(psc_b200_ctx* ctx);
/* ---- Sort::operator() = SortCountsort2 (psc_sort_impl.hxx:65-124) ---- */
int psc_b200_sort(psc_b200_ctx* ctx);
/* ---- BndParticles::operator() (bnd_particles_impl.hxx:234-247, ddc_particles.hxx:283-478) ---- */
int psc_b200_bnd_particles(psc_b200_ctx* ctx);
/* ---- Bnd::add_ghosts / fill_ghosts (psc_bnd_impl.hxx:105-158) ---- */
int psc_b200_bnd_add_ghosts(psc_b200_ctx* ctx, int field_id, int mb, int me);
int psc_b200_bnd_fill_ghosts(psc_b200_ctx* ctx, int field_id, int mb, int me);
/* ---- BndFields (psc_bnd_fields_impl.hxx:27-188) ---- */
int psc_b200_bndf_fill_ghosts_E(psc_b200_ctx* ctx);
int psc_b200_bndf_fill_ghosts_H(psc_b200_ctx* ctx);
int psc_b200_bndf_add_ghosts_J(psc_b200_ctx* ctx);
/* ---- PushFields::push_E / push_H (psc_push_fields_impl.hxx:134-178) ---- */
int psc_b200_push_E(psc_b200_ctx* ctx, double dt_fac);
int psc_b200_push_H(psc_b200_ctx* ctx, double dt_fac);
/* ---- Marder::correct_gauss (marder_impl.hxx:197-264) ---- */
int psc_b200_marder(psc_b200_ctx* ctx, double diffusion, int loop);
/* ---- Moment_rho_1st_nc incl. ghost add (psc/moment.hxx:149-171) into comp 0 of field_id ---- */
int psc_b200_moment_rho_1st_nc(psc_b200_ctx* ctx, int field_id);
/* ---- the 1st-order moments a deck's diagnostics and injectors use
 * (libpsc/psc_output_fields/fields_item_moments_1st.hxx:9-37 = ItemMoment<moment_*> of
 * include/psc/moment.hxx:119-311): Moment_n_1st, Moment_v_1st, Moment_p_1st, Moment_T_1st,
 * Moments_1st ("all": rho jx jy jz px py pz txx tyy tzz txy tyz tzx per kind) at cell
 * centres, Moment_rho_1st_nc at nodes.  The field must have psc_b200_moment_n_comps()
 * components; the result includes the reflecting-wall folds and the ghost add
 * (include/fields_item.hxx:36-134), like the reference's. ---- */
enum
{
  PSC_B200_MOMENT_N = 0,
  PSC_B200_MOMENT_V = 1,
  PSC_B200_MOMENT_P = 2,
  PSC_B200_MOMENT_T = 3,
  PSC_B200_MOMENT_ALL = 4,
  PSC_B200_MOMENT_RHO_NC = 5
};
int psc_b200_moment_n_comps(psc_b200_ctx* ctx, int moment);
int psc_b200_moment_1st(psc_b200_ctx* ctx, int field_id, int moment);
/* ---- Checks (checks_impl.hxx:33-215) ---- */
int psc_b200_check_continuity_begin(psc_b200_ctx* ctx);
int psc_b200_check_continuity_end(psc_b200_ctx* ctx, double* max_err);
int psc_b200_check_gauss(psc_b200_ctx* ctx, double* max_err);
/* ---- DiagEnergies (DiagEnergiesField.h:19-42, DiagEnergiesParticle.h:15-40):
 * out[0..5] = EX2 EY2 EZ2 HX2 HY2 HZ2 ; out[6] = E_electron (q<0), out[7] = E_ion ---- */
int psc_b200_energies(psc_b200_ctx* ctx, double out[8]);

/* ---- Collision (src/libpsc/psc_collision/psc_collision_impl.hxx:56-252 around
 * src/include/binary_collision.hxx:57-295): per cell a random permutation, then binary
 * Coulomb collisions of the pairs (a triangle at half rate when the population is odd).
 * Psc::step calls it every `interval` steps right after the sort (psc.hxx:363-371). ---- */
typedef struct
{
  int interval;  /* steps between calls (enters the collision frequency) */
  double nu;     /* collision frequency scale */
  double cori;   /* grid.norm.cori = 1 / nicell (grid.hxx:288) */
  int rng;       /* 0: RngFake -- uniform() = .5, identity permutation (binary_collision.hxx:36-41,
                    the reference's known-answer setting); 1: counter-based streams keyed by
                    (seed, step, global cell, pair) */
  uint64_t seed;
  uint64_t step; /* time step (keys the streams) */
} psc_b200_collision_params;
int psc_b200_collide(psc_b200_ctx* ctx, const psc_b200_collision_params* prm, uint64_t* n_collisions);

/* ---- Heating (libpsc/psc_heating/psc_heating_impl.hxx:27-76) with the HeatingSpotFoil
 * profile (include/heating_spot_foil.hxx:6-89): particles inside the spot get a Gaussian
 * momentum kick of variance H(x, kind) * interval * dt per component.  psc_flatfoil_yz
 * heats every 20 steps (psc_flatfoil_yz.cxx:464-467, 649-656). ---- */
typedef struct
{
  double zl, zh, xc, yc, rH; /* HeatingSpotFoilParams, internal units */
  double T[PSC_B200_MAX_KINDS];
  double Mi;
  int n_kinds;
  int interval;              /* heating_dt = interval * dt */
  uint64_t seed, step;       /* counter-based streams keyed by (seed, step, patch, index) */
} psc_b200_heating_params;
int psc_b200_heating_spot_foil(psc_b200_ctx* ctx, const psc_b200_heating_params* prm,
                               uint64_t* n_kicked);

/* ---- BoundaryInjector (src/include/boundary_injector.hxx:93-160; Psc::step calls the
 * injectors between the push and the particle exchange, psc.hxx:391-399).  The generator,
 * the one-step advance and the "did it enter the patch" test are host code there and stay
 * host code here (psc_config_b200.hxx BoundaryInjectorB200, psc_b200.api.BoundaryInjector);
 * the device takes the two halves that touch its data: the accepted particles go in through
 * psc_b200_mprts_inject, and the current of their way in is deposited by this call =
 * Current::calc_j(J, xm, xp, lf, lg, qni_wni, v) (boundary_injector.hxx:146-147;
 * inc_curr_1vb_split.cxx / inc_curr_1vb_var1.cxx as the grid's deposit selects) for every
 * trajectory, added to JXI..JZI of the state fields.
 * xm / xp: start and end, patch-local and normalised (x * dx_inv formed in float as the
 * reference does, :140-141); lg: the cell the trajectory starts in (the ghost cell the
 * particle was generated in; lf = fint(xp) is formed on the device); v: calc_v(u). ---- */
typedef struct
{
  int patch; /* local patch */
  int lg[3];
  float xm[3], xp[3], v[3];
  float qni_wni;
} psc_b200_jpath;
int psc_b200_deposit_j(psc_b200_ctx* ctx, const psc_b200_jpath* paths, uint64_t n);

/* ---- Psc::step (src/include/psc.hxx:321-486): the whole sequence on the stream ---- */
typedef struct
{
  int sort;             /* run Sort this step (PscParams::sort_interval hit) */
  int marder_loop;      /* > 0: Marder correction with this many relaxation loops */
  double marder_diffusion;
  int push_fields;      /* 0 = particle part only (push + exchange + sort) */
  int checks;           /* continuity + gauss checks (results via _last_checks) */
  int energies;         /* DiagEnergies reduced inside the step (field part behind the field
                           chain, particle part behind the sort); read with _last_energies */
} psc_b200_step_params;
int psc_b200_step(psc_b200_ctx* ctx, const psc_b200_step_params* prm);
int psc_b200_last_checks(psc_b200_ctx* ctx, double* continuity, double* gauss);
/* out[8] as psc_b200_energies, of the last step that asked for them */
int psc_b200_last_energies(psc_b200_ctx* ctx, double out[8]);

/* ---- pipelined host I/O: a deck whose field solver / diagnostics live on the host ----
 * The reference's CUDA backend stages fields through the host synchronously
 * (psc_fields_cuda.h:78-113 hostMirror/copy).  Here the transfers overlap the particle
 * re-sort of the same step:
 *   step_begin            sort (if needed), push + deposit, then side by side: the particle
 *                         boundary exchange + sort on one stream, J ghosts (+ Yee when
 *                         push_fields) on the field stream; returns without waiting
 *   mflds_download_async  / mflds_upload_async: queued on the field stream (PINNED host
 *                         memory, or they degrade to synchronous copies)
 *   io_wait               blocks until the field stream's transfers are done (the sort may
 *                         still be running)
 *   step_end              waits for the sort, commits it, joins the streams
 * Any other entry point completes a pending step first, so forgetting step_end is safe. */
int psc_b200_step_begin(psc_b200_ctx* ctx, const psc_b200_step_params* prm);
int psc_b200_step_end(psc_b200_ctx* ctx);
int psc_b200_mflds_download_async(psc_b200_ctx* ctx, int field_id, int mb, int me, float* host);
int psc_b200_mflds_upload_async(psc_b200_ctx* ctx, int field_id, int mb, int me,
                                const float* host);
int psc_b200_io_wait(psc_b200_ctx* ctx);

/* ---- checkpoint / restart (src/include/checkpoint.hxx:14-82: write_checkpoint puts "grid",
 * "mprts", "mflds"; read_checkpoint restores them).  One flat binary per rank, <path>.<rank>,
 * with PSC's variable decomposition (size_by_patch + one array per particle component,
 * particles_simple.inl:113-131; ib / im + component-major field patches, fields3d.inl:18-57).
 * The context that reads must have been created for the same grid and decomposition. ---- */
int psc_b200_checkpoint_write(psc_b200_ctx* ctx, const char* path, int64_t timestep);
int psc_b200_checkpoint_read(psc_b200_ctx* ctx, const char* path, int64_t* timestep);

/* ---- multi-GPU: NCCL over NVLink (replaces every MPI site of SURVEY.md 2.3) ---- */
int psc_b200_nccl_unique_id(void* id128);
int psc_b200_nccl_init(psc_b200_ctx* ctx, const void* id128);
/* Balance::operator() (psc_balance_impl.hxx:770-1026): redistribute patches by load */
int psc_b200_balance(psc_b200_ctx* ctx, double factor_fields, int* changed);
/* best_mapping (psc_balance_impl.hxx:99-160): pure function, exposed for tests */
int psc_b200_best_mapping(int n_ranks, const double* capability, int n_patches,
                          const double* loads, int* n_patches_by_rank);

/* ---- options, measurement ---- */
/* "keep_sorted" (0/1), "tile" (cells per tile edge), "warp_reduce" (0/1),
 * "threads" (CTA size of the push kernel), "profile" (0/1) */
/* Self-test of the exact build's device arithmetic: compares the guard-free 1/sqrt(x) and
 * 1/x sequences the push uses for arguments >= 1 with the compiler's correctly rounded
 * forms (= the reference's host arithmetic, cuda_compat.h:26-30) on every float in
 * [1, 2^80); *n_mismatch must come back 0. */
int psc_b200_selftest_math(psc_b200_ctx* ctx, uint64_t* n_mismatch);
int psc_b200_set_option(psc_b200_ctx* ctx, const char* name, double value);
int psc_b200_get_stat(psc_b200_ctx* ctx, const char* name, double* value);
/* CUDA-event timer on the context's stream */
int psc_b200_timer_start(psc_b200_ctx* ctx);
int psc_b200_timer_stop(psc_b200_ctx* ctx, float* ms);
/* per-kernel accumulators (when option "profile" = 1): writes up to max entries,
 * returns the number of kernels; names are static strings */
int psc_b200_prof_get(psc_b200_ctx* ctx, int max, const char** names, float* ms,
                      uint64_t* launches);
int psc_b200_prof_reset(psc_b200_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif
