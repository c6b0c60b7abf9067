#include <stdint.h>
/* TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of psc-code/psc's
 * per-timestep PIC hot path.  Never linked into, imported by or called from the
 * product path (psc_b200/): only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may use it, and only as the checker.
 *
 * Parity pinning (see tests/test_oracle_*.py):
 *   - push/interpolate/deposit: bit-exact against oracle/_ref/libpsc_ref.so
 *     (the reference's own headers compiled unmodified) on random inputs, and
 *     against the golden vectors of src/libpsc/tests/test_push_particles.cxx and
 *     test_current_deposition.cxx (tests/golden/ JSON fixtures).
 *   - sort: pinned by the 15-particle vector of test_collision_cuda.cxx:145-190
 *     plus the stable-counting-sort definition (no CPU test exists upstream).
 *   - ghost fill/add: the exact FillGhosts / AddGhosts patterns of test_bnd.cxx:104-303;
 *     Yee, Marder correct, div: Pushf1/2, MarderCorrect, ItemDivE/J of
 *     test_push_fields.cxx:26-260; the 1st-order moments: the 14 known-answer cases of
 *     test_moments.cxx:149-402; push + exchange + J ghosts + continuity over many steps:
 *     Accel / Cyclo of test_push_particles_2.cxx:23-170 (tests/golden_cases.py,
 *     tests/test_oracle_golden.py, tests/test_accel_cyclo.py).
 *   - particle migration order, Var1 vs Split on yz, the balancer mapping, energies over
 *     1000 steps: parity unpinned by upstream tests; pinned here only against _ref where
 *     _ref covers it.
 *
 * All arrays use PSC's layouts:
 *   fields   : float [p][m][iz][iy][ix], ix fastest, dims im = ldims + 2*ibn,
 *              lower bound ib = -ibn (fields3d.hxx:29-32,284-291); m = JXI..HZ
 *   particles: 32-byte AoS records {x[3], u[3], int kind, qni_wni}
 *              (particle_simple.hxx:10-42), patch p = [off[p], off[p+1])
 */
#ifndef PSC_ORACLE_H
#define PSC_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

enum
{
  PO_JXI,
  PO_JYI,
  PO_JZI,
  PO_EX,
  PO_EY,
  PO_EZ,
  PO_HX,
  PO_HY,
  PO_HZ,
  PO_NR_FIELDS
};

/* grid/BC.h */
enum
{
  PO_BND_FLD_OPEN,
  PO_BND_FLD_PERIODIC,
  PO_BND_FLD_CONDUCTING_WALL,
  PO_BND_FLD_ABSORBING
};
enum
{
  PO_BND_PRT_REFLECTING,
  PO_BND_PRT_PERIODIC,
  PO_BND_PRT_ABSORBING,
  PO_BND_PRT_OPEN
};

enum
{
  PO_DEPOSIT_VAR1 = 0, /* Current1vbVar1 (yz only; PSC production for dim_yz) */
  PO_DEPOSIT_SPLIT = 1 /* Current1vbSplit (PSC production for xyz; tests for yz) */
};

#define PO_MAX_KINDS 10

typedef struct
{
  float x[3];
  float u[3];
  int kind;
  float qni_wni;
} po_prt;

typedef struct
{
  /* inputs */
  int gdims[3], np[3];
  double length[3], corner[3];
  double dt;
  double fnqs, eta;
  int n_kinds;
  double q[PO_MAX_KINDS], m[PO_MAX_KINDS];
  int bc_fld_lo[3], bc_fld_hi[3], bc_prt_lo[3], bc_prt_hi[3];
  int deposit;
  /* derived (po_grid_setup) */
  int ldims[3], ibn[3], im[3], ib[3];
  int invar[3];
  double dx[3], dx_inv[3];
  int n_patches;
  int periodic[3]; /* mrc_domain bc: fld_lo periodic && gdims > 1 */
} po_grid;

/* fills the derived members; ibn = 2 in non-invariant dims (all decks) */
void po_grid_setup(po_grid* g);
void po_patch_idx3(const po_grid* g, int p, int idx3[3]);
void po_patch_off(const po_grid* g, int p, int off[3]);
/* neighbour patch in direction dir (mrc_domain_multi.c:518-545); -1 if none */
int po_neighbor_patch(const po_grid* g, int p, const int dir[3]);
long po_fld_patch_len(const po_grid* g); /* im0*im1*im2 */

/* push_particles_1vb.hxx:27-84 : zero J, gather, Boris, move, deposit */
void po_push_mprts(const po_grid* g, float* flds, po_prt* prts,
                   const unsigned* off);
/* same, restricted to patches [p0,p1) (for threaded CPU-baseline timing) */
void po_push_mprts_range(const po_grid* g, float* flds, po_prt* prts,
                         const unsigned* off, int p0, int p1);

/* single-trajectory deposit in float / double (test_current_deposition.cxx) */
void po_calc_j_f(const po_grid* g, float* flds_patch, const float xm[3],
                 const float xp[3], const float vxi[3], float qni_wni);
void po_calc_j_d(const po_grid* g, double* flds_patch, const double xm[3],
                 const double xp[3], const double vxi[3], double qni_wni);

/* psc_sort_impl.hxx:65-124 (SortCountsort2): stable counting sort by cell,
 * per patch, in place.  If perm != NULL it receives, per particle slot of the
 * OUTPUT, the index (within the patch) of the input particle placed there.
 * returns 0, or -1 if a particle has no valid cell (PSC asserts). */
int po_sort(const po_grid* g, po_prt* prts, const unsigned* off,
            unsigned* perm);
int po_sort_range(const po_grid* g, po_prt* prts, const unsigned* off,
                  unsigned* perm, int p0, int p1);
/* particle_indexer.hxx:74-94 : cell index or -1 */
int po_cell_index(const po_grid* g, const float x[3]);
/* per-cell counts, n_patches * ldims0*ldims1*ldims2 entries */
void po_count_by_cell(const po_grid* g, const po_prt* prts,
                      const unsigned* off, unsigned* cnt);

/* bnd_particles_impl.hxx:93-218 + ddc_particles.hxx:283-478.
 * rank_of_patch == NULL means everything on one rank.  Output per patch:
 * [stayers in order | same-rank arrivals in direction order | other-rank
 * arrivals by (rank, sender patch, sender direction)].
 * prts_out must hold off_in[n_patches] records. n_dropped counts absorbed. */
void po_bnd_particles(const po_grid* g, const po_prt* prts_in,
                      const unsigned* off_in, po_prt* prts_out,
                      unsigned* off_out, const int* rank_of_patch,
                      unsigned* n_dropped);

/* psc_bnd_impl.hxx:105-158 + mrc_ddc_multi.c:60-135,519-538, components [mb,me)
 * of an n_comps-component field array with the grid's im/ib */
void po_fill_ghosts(const po_grid* g, float* flds, int n_comps, int mb, int me);
void po_add_ghosts(const po_grid* g, float* flds, int n_comps, int mb, int me);

/* psc_push_fields_impl.hxx:50-178 */
void po_push_E(const po_grid* g, float* flds, double dt_fac);
void po_push_H(const po_grid* g, float* flds, double dt_fac);

/* psc_bnd_fields_impl.hxx:27-188,301-530 (conducting wall; periodic = no-op) */
void po_bndf_fill_ghosts_E(const po_grid* g, float* flds);
void po_bndf_fill_ghosts_H(const po_grid* g, float* flds);
void po_bndf_add_ghosts_J(const po_grid* g, float* flds);

/* Moment_rho_1st_nc (psc/moment.hxx:149-171, psc/deposit.hxx:24-65,172-191)
 * into a 1-component array with the grid's im/ib, INCLUDING the ghost add */
void po_moment_rho_1st_nc(const po_grid* g, const po_prt* prts,
                          const unsigned* off, float* rho);
/* the 1st-order moment family (fields_item_moments_1st.hxx:9-37): n, v, p, T, "all" (13 per
 * kind) at cell centres, rho at nodes; out has po_moment_n_comps() components with the
 * grid's im/ib; reflecting-wall folds and the ghost add included */
enum
{
  PO_MOM_N = 0,
  PO_MOM_V = 1,
  PO_MOM_P = 2,
  PO_MOM_T = 3,
  PO_MOM_ALL = 4,
  PO_MOM_RHO_NC = 5
};
int po_moment_n_comps(const po_grid* g, int which);
void po_moment_1st(const po_grid* g, const po_prt* prts, const unsigned* off, int which,
                   float* out);
/* psc::item::div_nc (fields_item_fields.hxx:65-104) of components m0..m0+2 of
 * flds into a 1-component array (interior points only, ghosts left 0) */
void po_div_nc(const po_grid* g, const float* flds, int n_comps, int m0,
               float* div);
/* checks_impl.hxx:33-132 : max |rho_p - rho_m + dt * div J| over interior */
double po_continuity(const po_grid* g, const float* rho_m, const float* rho_p,
                     const float* flds);
/* checks_impl.hxx:137-215 : max |div E - rho| */
double po_gauss(const po_grid* g, const float* rho, const float* flds);
/* marder_impl.hxx:26-61,197-264 */
/* psc::marder::correct alone (marder_impl.hxx:26-61): E += grad(res) * .5 dt diffusion */
void po_marder_apply(const po_grid* g, float* flds, const float* res, double diffusion);
void po_marder_correct(const po_grid* g, float* flds, const po_prt* prts,
                       const unsigned* off, double diffusion, int loop);

/* DiagEnergiesField.h:19-42, DiagEnergiesParticle.h:15-40:
 * out[0..5] = EX2 EY2 EZ2 HX2 HY2 HZ2, out[6..6+n_kinds) kinetic by kind */
void po_energies(const po_grid* g, const float* flds, const po_prt* prts,
                 const unsigned* off, double* out);

/* Balance: best_mapping / best_mapping_recursive (psc_balance_impl.hxx:99-160): 1-D
 * recursive bisection of the patch list by load, at least one patch per rank;
 * get_loads (psc_balance_impl.hxx:223-269): load = n_prts + factor_fields * n_cells */
/* ---- binary Coulomb collisions (psc_oracle_collision.inc) ---- */
#define PO_RNG_FAKE 0 /* RngFake: uniform() = .5, identity permutation (binary_collision.hxx:36-41) */
#define PO_RNG_HASH 1 /* the counter-based streams shared with the device kernel */
float po_binary_collision_f(float u1[3], float u2[3], float q1, float m1, float q2, float m2, float nudt1,
                            float ran1, float ran2);
double po_binary_collision_d(double u1[3], double u2[3], double q1, double m1, double q2, double m2,
                             double nudt1, double ran1, double ran2);
long po_collide(const po_grid* g, po_prt* prts, const unsigned* off, int interval, double nu, double cori,
                int rng, uint64_t seed, uint64_t step, int patch_begin);

/* ---- heating (psc_oracle_collision.inc): HeatingSpotFoilParams (heating_spot_foil.hxx:6-16) + cadence/streams */
typedef struct
{
  double zl, zh, xc, yc, rH;
  double T[PO_MAX_KINDS];
  double Mi;
  int n_kinds;
  int interval;
  uint64_t seed, step;
} po_heating_prm;
double po_heating_spot_foil_H(const po_grid* g, const po_heating_prm* hp, const double crd[3], int kind);
long po_heating(const po_grid* g, const po_heating_prm* hp, po_prt* prts, const unsigned* off, int patch_begin);

/* BoundaryInjector::inject (boundary_injector.hxx:93-160) on the generator's draws */
typedef struct
{
  int patch;
  int idx[3];  /* the ghost cell the particle was generated in (initial_idx) */
  double x[3]; /* patch-local (generator.get(cell_corner, dx)) */
  double u[3];
  double w;
  int kind;
} po_inject_cand;
long po_boundary_inject(const po_grid* g, float* flds, const po_inject_cand* cand, long n, po_prt* out,
                        int* out_patch);

void po_best_mapping(int n_ranks, const double* capability, int n_patches,
                     const double* loads, int* n_patches_by_rank);
void po_get_loads(const po_grid* g, const unsigned* off, double factor_fields,
                  double* loads);

const char* po_describe(void);

#ifdef __cplusplus
}
#endif
#endif
