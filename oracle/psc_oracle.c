/* TEST INFRASTRUCTURE ONLY -- see psc_oracle.h.
 *
 * Plain-C restatement of psc-code/psc's per-timestep PIC hot path.  Every
 * function cites the reference file:line it follows.  Build with
 * gcc -O3 -ffp-contract=off and no -march (PSC Release: no FMA), see Makefile.
 */
#include "psc_oracle.h"

#include <assert.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ====================================================================== */
/* grid */

/* grid/domain.hxx:25-47, grid.hxx:68-101, mrc_domain.hxx:18-29 */
void po_grid_setup(po_grid* g)
{
  g->n_patches = 1;
  for (int d = 0; d < 3; d++) {
    assert(g->gdims[d] > 0 && g->np[d] > 0 && g->gdims[d] % g->np[d] == 0);
    g->ldims[d] = g->gdims[d] / g->np[d];
    g->dx[d] = g->length[d] / (double)g->gdims[d];
    g->dx_inv[d] = (double)g->gdims[d] / g->length[d];
    g->invar[d] = g->gdims[d] == 1; /* domain.hxx:55 */
    g->ibn[d] = g->invar[d] ? 0 : 2; /* decks: psc_bubble_yz.cxx:143-144 */
    g->im[d] = g->ldims[d] + 2 * g->ibn[d];
    g->ib[d] = -g->ibn[d];
    if (g->invar[d]) { /* grid.hxx:90-98 */
      g->bc_fld_lo[d] = g->bc_fld_hi[d] = PO_BND_FLD_PERIODIC;
      g->bc_prt_lo[d] = g->bc_prt_hi[d] = PO_BND_PRT_PERIODIC;
    }
    g->periodic[d] = g->bc_fld_lo[d] == PO_BND_FLD_PERIODIC && g->gdims[d] > 1;
    g->n_patches *= g->np[d];
  }
}

/* "bydim" curve: p = (pz*npy + py)*npx + px (mrc_domain_lib.c:21-35) */
void po_patch_idx3(const po_grid* g, int p, int idx3[3])
{
  idx3[0] = p % g->np[0];
  idx3[1] = (p / g->np[0]) % g->np[1];
  idx3[2] = p / (g->np[0] * g->np[1]);
}

void po_patch_off(const po_grid* g, int p, int off[3])
{
  int idx3[3];
  po_patch_idx3(g, p, idx3);
  for (int d = 0; d < 3; d++) {
    off[d] = idx3[d] * g->ldims[d];
  }
}

/* mrc_domain_multi.c:518-545 */
int po_neighbor_patch(const po_grid* g, int p, const int dir[3])
{
  int idx3[3], nei[3];
  po_patch_idx3(g, p, idx3);
  for (int d = 0; d < 3; d++) {
    nei[d] = idx3[d] + dir[d];
    if (g->periodic[d]) {
      if (nei[d] < 0) {
        nei[d] += g->np[d];
      }
      if (nei[d] >= g->np[d]) {
        nei[d] -= g->np[d];
      }
    }
    if (nei[d] < 0 || nei[d] >= g->np[d]) {
      return -1;
    }
  }
  return (nei[2] * g->np[1] + nei[1]) * g->np[0] + nei[0];
}

long po_fld_patch_len(const po_grid* g)
{
  return (long)g->im[0] * g->im[1] * g->im[2];
}

static int at_boundary_lo(const po_grid* g, int p, int d)
{ /* grid.hxx:115 */
  int off[3];
  po_patch_off(g, p, off);
  return off[d] == 0;
}

static int at_boundary_hi(const po_grid* g, int p, int d)
{ /* grid.hxx:116-119 */
  int off[3];
  po_patch_off(g, p, off);
  return off[d] + g->ldims[d] == g->gdims[d];
}

/* Fields3d<F, dim_xyz>::operator() (fields.hxx:50-57): offset only */
#define FLD(f, g, m, i, j, k)                                                  \
  ((f)[(((long)(m) * (g)->im[2] + ((k) - (g)->ib[2])) * (g)->im[1] +           \
        ((j) - (g)->ib[1])) *                                                  \
         (g)->im[0] +                                                          \
       ((i) - (g)->ib[0])])

/* Fields3d<F, dim> with invariant directions forced to 0 (fields.hxx:50-57) */
#define FLDI(f, g, m, i, j, k)                                                 \
  FLD(f, g, m, (g)->invar[0] ? (g)->ib[0] : (i),                               \
      (g)->invar[1] ? (g)->ib[1] : (j), (g)->invar[2] ? (g)->ib[2] : (k))

/* ====================================================================== */
/* deposit (float and double instances) */

#define REAL float
#define SUFFIX _f
#include "psc_oracle_deposit.inc"
#undef REAL
#undef SUFFIX

#define REAL double
#define SUFFIX _d
#include "psc_oracle_deposit.inc"
#undef REAL
#undef SUFFIX

void po_calc_j_f(const po_grid* g, float* flds_patch, const float xm_[3],
                 const float xp_[3], const float vxi[3], float qni_wni)
{
  po_curr_f c;
  po_curr_setup_f(&c, g, flds_patch);
  float xm[3] = {xm_[0], xm_[1], xm_[2]}, xp[3] = {xp_[0], xp_[1], xp_[2]};
  int lg[3], lf[3];
  for (int d = 0; d < 3; d++) {
    lg[d] = po_fint_f(xm[d]);
    lf[d] = po_fint_f(xp[d]);
  }
  po_calc_j_impl_f(&c, g->deposit, xm, xp, lf, lg, qni_wni, vxi);
}

void po_calc_j_d(const po_grid* g, double* flds_patch, const double xm_[3],
                 const double xp_[3], const double vxi[3], double qni_wni)
{
  po_curr_d c;
  po_curr_setup_d(&c, g, flds_patch);
  double xm[3] = {xm_[0], xm_[1], xm_[2]}, xp[3] = {xp_[0], xp_[1], xp_[2]};
  int lg[3], lf[3];
  for (int d = 0; d < 3; d++) {
    lg[d] = po_fint_d(xm[d]);
    lf[d] = po_fint_d(xp[d]);
  }
  po_calc_j_impl_d(&c, g->deposit, xm, xp, lf, lg, qni_wni, vxi);
}

/* ====================================================================== */
/* push */

static inline float sqrf(float a) { return a * a; }
/* cuda_compat.h:26-30: host rsqrt = 1/sqrt */
static inline float rsqrt_host(float x) { return 1.f / sqrtf(x); }

/* pushp.hxx:36-63 */
static inline void push_p(float p[3], const float E[3], const float H[3],
                          float dq)
{
  float pxm = p[0] + dq * E[0];
  float pym = p[1] + dq * E[1];
  float pzm = p[2] + dq * E[2];

  float root = dq * rsqrt_host(1.f + sqrf(pxm) + sqrf(pym) + sqrf(pzm));
  float taux = H[0] * root, tauy = H[1] * root, tauz = H[2] * root;

  float tau = 1.f / (1.f + sqrf(taux) + sqrf(tauy) + sqrf(tauz));
  float pxp = ((1.f + sqrf(taux) - sqrf(tauy) - sqrf(tauz)) * pxm +
               (2.f * taux * tauy + 2.f * tauz) * pym +
               (2.f * taux * tauz - 2.f * tauy) * pzm) *
              tau;
  float pyp = ((2.f * taux * tauy - 2.f * tauz) * pxm +
               (1.f - sqrf(taux) + sqrf(tauy) - sqrf(tauz)) * pym +
               (2.f * tauy * tauz + 2.f * taux) * pzm) *
              tau;
  float pzp = ((2.f * taux * tauz + 2.f * tauy) * pxm +
               (2.f * tauy * tauz - 2.f * taux) * pym +
               (1.f - sqrf(taux) - sqrf(tauy) + sqrf(tauz)) * pzm) *
              tau;

  p[0] = pxp + dq * E[0];
  p[1] = pyp + dq * E[1];
  p[2] = pzp + dq * E[2];
}

/* push_particles_1vb.hxx:27-84 for one patch */
static void push_patch(const po_grid* g, float* F, po_prt* prts, unsigned n)
{
  int yz = g->invar[0] && !g->invar[1] && !g->invar[2];
  assert(yz || (!g->invar[0] && !g->invar[1] && !g->invar[2]));

  /* :30  Real3 dxi = Real3(grid.domain.dx).inv() */
  float dxi[3];
  for (int d = 0; d < 3; d++) {
    dxi[d] = 1.f / (float)g->dx[d];
  }
  /* :31-36  (double expression, narrowed on assignment) */
  float dq_kind[PO_MAX_KINDS];
  for (int k = 0; k < g->n_kinds; k++) {
    dq_kind[k] = (float)(.5f * g->eta * g->dt * g->q[k] / g->m[k]);
  }
  float dt = (float)g->dt; /* AdvanceParticle(grid.dt), pushp.hxx:12 */

  po_curr_f cur;
  po_curr_setup_f(&cur, g, F);

  /* :48 */
  memset(F, 0, sizeof(float) * 3 * po_fld_patch_len(g));

  for (unsigned n_ = 0; n_ < n; n_++) {
    po_prt* prt = &prts[n_];
    float* x = prt->x;
    float xn[3];
    for (int d = 0; d < 3; d++) {
      xn[d] = x[d] * dxi[d];
    }
    /* ip.set_coeffs: interpolate.hxx:44-58 (opt_ip_1st_ec uses only .g) */
    int l[3];
    float v0[3], v1[3];
    for (int d = 0; d < 3; d++) {
      l[d] = (int)floorf(xn[d]);
      float h = xn[d] - (float)l[d];
      v0[d] = 1.f - h;
      v1[d] = h;
    }

    float E[3], H[3];
    if (!yz) {
      /* interpolate.hxx:140-192 */
      int lx = l[0], ly = l[1], lz = l[2];
      E[0] = (v0[2] * (v0[1] * FLD(F, g, PO_EX, lx, ly, lz) +
                       v1[1] * FLD(F, g, PO_EX, lx, ly + 1, lz)) +
              v1[2] * (v0[1] * FLD(F, g, PO_EX, lx, ly, lz + 1) +
                       v1[1] * FLD(F, g, PO_EX, lx, ly + 1, lz + 1)));
      E[1] = (v0[0] * (v0[2] * FLD(F, g, PO_EY, lx, ly, lz) +
                       v1[2] * FLD(F, g, PO_EY, lx, ly, lz + 1)) +
              v1[0] * (v0[2] * FLD(F, g, PO_EY, lx + 1, ly, lz) +
                       v1[2] * FLD(F, g, PO_EY, lx + 1, ly, lz + 1)));
      E[2] = (v0[1] * (v0[0] * FLD(F, g, PO_EZ, lx, ly, lz) +
                       v1[0] * FLD(F, g, PO_EZ, lx + 1, ly, lz)) +
              v1[1] * (v0[0] * FLD(F, g, PO_EZ, lx, ly + 1, lz) +
                       v1[0] * FLD(F, g, PO_EZ, lx + 1, ly + 1, lz)));
      H[0] = (v0[0] * FLD(F, g, PO_HX, lx, ly, lz) +
              v1[0] * FLD(F, g, PO_HX, lx + 1, ly, lz));
      H[1] = (v0[1] * FLD(F, g, PO_HY, lx, ly, lz) +
              v1[1] * FLD(F, g, PO_HY, lx, ly + 1, lz));
      H[2] = (v0[2] * FLD(F, g, PO_HZ, lx, ly, lz) +
              v1[2] * FLD(F, g, PO_HZ, lx, ly, lz + 1));
    } else {
      /* interpolate.hxx:245-287 */
      int ly = l[1], lz = l[2];
      E[0] = (v0[2] * (v0[1] * FLD(F, g, PO_EX, 0, ly, lz) +
                       v1[1] * FLD(F, g, PO_EX, 0, ly + 1, lz)) +
              v1[2] * (v0[1] * FLD(F, g, PO_EX, 0, ly, lz + 1) +
                       v1[1] * FLD(F, g, PO_EX, 0, ly + 1, lz + 1)));
      E[1] = (v0[2] * FLD(F, g, PO_EY, 0, ly, lz) +
              v1[2] * FLD(F, g, PO_EY, 0, ly, lz + 1));
      E[2] = (v0[1] * FLD(F, g, PO_EZ, 0, ly, lz) +
              v1[1] * FLD(F, g, PO_EZ, 0, ly + 1, lz));
      H[0] = FLD(F, g, PO_HX, 0, ly, lz);
      H[1] = (v0[1] * FLD(F, g, PO_HY, 0, ly, lz) +
              v1[1] * FLD(F, g, PO_HY, 0, ly + 1, lz));
      H[2] = (v0[2] * FLD(F, g, PO_HZ, 0, ly, lz) +
              v1[2] * FLD(F, g, PO_HZ, 0, ly, lz + 1));
    }

    /* :59-60 */
    float dq = dq_kind[prt->kind];
    push_p(prt->u, E, H, dq);

    /* :63-64  calc_v (pushp.hxx:68-72), push_x (pushp.hxx:17-29) */
    float root = rsqrt_host(1.f + sqrf(prt->u[0]) + sqrf(prt->u[1]) +
                            sqrf(prt->u[2]));
    float v[3] = {prt->u[0] * root, prt->u[1] * root, prt->u[2] * root};
    for (int d = 0; d < 3; d++) {
      if (!g->invar[d]) {
        x[d] += 1.f * dt * v[d];
      }
    }

    /* :66-67 */
    float xp[3];
    int lf[3];
    for (int d = 0; d < 3; d++) {
      xp[d] = x[d] * dxi[d];
      lf[d] = (int)floorf(xp[d]);
    }
    /* :70-82 */
    po_calc_j_impl_f(&cur, g->deposit, xn, xp, lf, l, prt->qni_wni, v);
  }
}

void po_push_mprts_range(const po_grid* g, float* flds, po_prt* prts,
                         const unsigned* off, int p0, int p1)
{
  long plen = po_fld_patch_len(g) * PO_NR_FIELDS;
  for (int p = p0; p < p1; p++) {
    push_patch(g, flds + p * plen, prts + off[p], off[p + 1] - off[p]);
  }
}

void po_push_mprts(const po_grid* g, float* flds, po_prt* prts,
                   const unsigned* off)
{
  po_push_mprts_range(g, flds, prts, off, 0, g->n_patches);
}

/* ====================================================================== */
/* sort */

/* particle_indexer.hxx:68-94: dxi_ = real_t(grid.domain.dx_inv) */
static inline int cell_position(const po_grid* g, float x, int d)
{
  return (int)floorf(x * (float)g->dx_inv[d]);
}

int po_cell_index(const po_grid* g, const float x[3])
{
  int cpos[3];
  for (int d = 0; d < 3; d++) {
    cpos[d] = cell_position(g, x[d], d);
    if ((unsigned)cpos[d] >= (unsigned)g->ldims[d]) {
      return -1;
    }
  }
  return (cpos[2] * g->ldims[1] + cpos[1]) * g->ldims[0] + cpos[0];
}

/* psc_sort_impl.hxx:65-124 */
int po_sort_range(const po_grid* g, po_prt* prts, const unsigned* off,
                  unsigned* perm, int p0, int p1)
{
  unsigned n_cells = (unsigned)g->ldims[0] * g->ldims[1] * g->ldims[2];
  int rc = 0;
  for (int p = p0; p < p1; p++) {
    po_prt* P = prts + off[p];
    unsigned n_prts = off[p + 1] - off[p];
    unsigned* cnis = malloc(sizeof(unsigned) * (n_prts ? n_prts : 1));
    for (unsigned i = 0; i < n_prts; i++) {
      int ci = po_cell_index(g, P[i].x);
      if (ci < 0) { /* validCellIndex asserts */
        rc = -1;
        ci = 0;
      }
      cnis[i] = ci;
    }
    unsigned* cnts = calloc(n_cells, sizeof(unsigned));
    for (unsigned i = 0; i < n_prts; i++) {
      cnts[cnis[i]]++;
    }
    unsigned cur = 0;
    for (unsigned i = 0; i < n_cells; i++) {
      unsigned n = cnts[i];
      cnts[i] = cur;
      cur += n;
    }
    assert(cur == n_prts);
    po_prt* P2 = malloc(sizeof(po_prt) * (n_prts ? n_prts : 1));
    for (unsigned i = 0; i < n_prts; i++) {
      unsigned cni = cnis[i];
      unsigned n = 1;
      while (i + n < n_prts && cnis[i + n] == cni) {
        n++;
      }
      memcpy(&P2[cnts[cni]], &P[i], n * sizeof(po_prt));
      if (perm) {
        for (unsigned k = 0; k < n; k++) {
          perm[off[p] + cnts[cni] + k] = i + k;
        }
      }
      cnts[cni] += n;
      i += n - 1;
    }
    memcpy(P, P2, n_prts * sizeof(po_prt));
    free(P2);
    free(cnis);
    free(cnts);
  }
  return rc;
}

int po_sort(const po_grid* g, po_prt* prts, const unsigned* off, unsigned* perm)
{
  return po_sort_range(g, prts, off, perm, 0, g->n_patches);
}

void po_count_by_cell(const po_grid* g, const po_prt* prts,
                      const unsigned* off, unsigned* cnt)
{
  long n_cells = (long)g->ldims[0] * g->ldims[1] * g->ldims[2];
  memset(cnt, 0, sizeof(unsigned) * n_cells * g->n_patches);
  for (int p = 0; p < g->n_patches; p++) {
    for (unsigned i = off[p]; i < off[p + 1]; i++) {
      int ci = po_cell_index(g, prts[i].x);
      if (ci >= 0) {
        cnt[p * n_cells + ci]++;
      }
    }
  }
}

/* ====================================================================== */
/* particle boundary exchange */

typedef struct
{
  po_prt* v;
  unsigned n, cap;
} prt_vec;

static void pv_push(prt_vec* pv, const po_prt* prt)
{
  if (pv->n == pv->cap) {
    pv->cap = pv->cap ? 2 * pv->cap : 16;
    pv->v = realloc(pv->v, sizeof(po_prt) * pv->cap);
  }
  pv->v[pv->n++] = *prt;
}

static inline int dir2idx(const int dir[3])
{ /* mrc_ddc.h:64-67 */
  return ((dir[2] + 1) * 3 + dir[1] + 1) * 3 + dir[0] + 1;
}

void po_bnd_particles(const po_grid* g, const po_prt* prts_in,
                      const unsigned* off_in, po_prt* prts_out,
                      unsigned* off_out, const int* rank_of_patch,
                      unsigned* n_dropped)
{
  int np = g->n_patches;
  prt_vec* send = calloc((size_t)np * 27, sizeof(prt_vec));
  prt_vec* stay = calloc(np, sizeof(prt_vec));
  unsigned dropped = 0;

  /* process_patch: bnd_particles_impl.hxx:93-218.  Patches are independent (the
   * reference runs them on different MPI ranks): one OpenMP task per patch, same
   * results in any order. */
#pragma omp parallel for schedule(dynamic) reduction(+ : dropped)
  for (int p = 0; p < np; p++) {
    int poff[3];
    po_patch_off(g, p, poff);
    float patch_size[3];
    for (int d = 0; d < 3; d++) {
      /* grid.hxx:82-86: xb, xe double; :105 Real3(xe - xb) */
      double xb = (double)poff[d] * g->dx[d] + g->corner[d];
      double xe = (double)(poff[d] + g->ldims[d]) * g->dx[d] + g->corner[d];
      patch_size[d] = (float)(xe - xb);
    }
    for (unsigned n = off_in[p]; n < off_in[p + 1]; n++) {
      po_prt prt = prts_in[n];
      float* xi = prt.x;
      float* pxi = prt.u;
      int pos[3];
      int valid = 1;
      for (int d = 0; d < 3; d++) {
        pos[d] = cell_position(g, xi[d], d);
        if ((unsigned)pos[d] >= (unsigned)g->ldims[d]) {
          valid = 0;
        }
      }
      if (valid) {
        pv_push(&stay[p], &prt);
        continue;
      }
      int drop = 0;
      int dir[3] = {0, 0, 0};
      for (int d = 0; d < 3; d++) {
        if (pos[d] < 0) {
          if (!at_boundary_lo(g, p, d) ||
              g->bc_prt_lo[d] == PO_BND_PRT_PERIODIC) {
            xi[d] += patch_size[d];
            dir[d] = -1;
            int ci = cell_position(g, xi[d], d);
            if (ci >= g->ldims[d]) {
              xi[d] = 0.;
              dir[d] = 0;
            }
          } else {
            switch (g->bc_prt_lo[d]) {
              case PO_BND_PRT_REFLECTING:
                xi[d] = -xi[d];
                pxi[d] = -pxi[d];
                dir[d] = 0;
                break;
              case PO_BND_PRT_OPEN:
              case PO_BND_PRT_ABSORBING: drop = 1; break;
              default: assert(0);
            }
          }
        } else if (pos[d] >= g->ldims[d]) {
          if (!at_boundary_hi(g, p, d) ||
              g->bc_prt_hi[d] == PO_BND_PRT_PERIODIC) {
            xi[d] -= patch_size[d];
            dir[d] = +1;
            int ci = cell_position(g, xi[d], d);
            if (ci < 0) {
              xi[d] = 0.;
            }
          } else {
            switch (g->bc_prt_hi[d]) {
              case PO_BND_PRT_REFLECTING: {
                xi[d] = 2.f * patch_size[d] - xi[d];
                pxi[d] = -pxi[d];
                dir[d] = 0;
                int ci = cell_position(g, xi[d], d);
                if (ci >= g->ldims[d]) {
                  xi[d] = (float)(xi[d] * (1. - 1e-6));
                }
                break;
              }
              case PO_BND_PRT_OPEN:
              case PO_BND_PRT_ABSORBING: drop = 1; break;
              default: assert(0);
            }
          }
        } else {
          dir[d] = 0;
        }
        if (!drop) {
          if (xi[d] < 0.f && xi[d] > -1e-6f) {
            xi[d] = 0.f;
          }
        }
      }
      if (!drop) {
        if (dir[0] == 0 && dir[1] == 0 && dir[2] == 0) {
          pv_push(&stay[p], &prt);
        } else {
          pv_push(&send[p * 27 + dir2idx(dir)], &prt);
        }
      } else {
        dropped++;
      }
    }
  }

  /* ddc_particles::comm (ddc_particles.hxx:283-478): first the size of every
   * receiving patch (its stayers + what its neighbours send it), then the copies, one
   * patch per task */
  off_out[0] = 0;
  for (int p = 0; p < np; p++) {
    unsigned cnt = stay[p].n;
    int my_rank = rank_of_patch ? rank_of_patch[p] : 0;
    int dir[3];
    for (dir[2] = -1; dir[2] <= 1; dir[2]++) {
      for (dir[1] = -1; dir[1] <= 1; dir[1]++) {
        for (dir[0] = -1; dir[0] <= 1; dir[0]++) {
          if (dir[0] == 0 && dir[1] == 0 && dir[2] == 0) {
            continue;
          }
          int nei = po_neighbor_patch(g, p, dir);
          if (nei < 0 || (rank_of_patch ? rank_of_patch[nei] : 0) != my_rank) {
            continue;
          }
          int dirneg[3] = {-dir[0], -dir[1], -dir[2]};
          cnt += send[nei * 27 + dir2idx(dirneg)].n;
        }
      }
    }
    if (rank_of_patch) {
      for (int q = 0; q < np; q++) {
        if (rank_of_patch[q] == my_rank) {
          continue;
        }
        for (int di = 0; di < 27; di++) {
          int d3[3] = {di % 3 - 1, (di / 3) % 3 - 1, di / 9 - 1};
          if (di != 13 && po_neighbor_patch(g, q, d3) == p) {
            cnt += send[q * 27 + di].n;
          }
        }
      }
    }
    off_out[p + 1] = off_out[p] + cnt;
  }
#pragma omp parallel for schedule(dynamic)
  for (int p = 0; p < np; p++) {
    unsigned cur = off_out[p];
    memcpy(prts_out + cur, stay[p].v, sizeof(po_prt) * stay[p].n);
    cur += stay[p].n;
    int my_rank = rank_of_patch ? rank_of_patch[p] : 0;
    /* :433-454 local neighbours, direction loop order */
    int dir[3];
    for (dir[2] = -1; dir[2] <= 1; dir[2]++) {
      for (dir[1] = -1; dir[1] <= 1; dir[1]++) {
        for (dir[0] = -1; dir[0] <= 1; dir[0]++) {
          if (dir[0] == 0 && dir[1] == 0 && dir[2] == 0) {
            continue;
          }
          int nei = po_neighbor_patch(g, p, dir);
          if (nei < 0) {
            continue;
          }
          int nei_rank = rank_of_patch ? rank_of_patch[nei] : 0;
          if (nei_rank != my_rank) {
            continue;
          }
          int dirneg[3] = {-dir[0], -dir[1], -dir[2]};
          prt_vec* sb = &send[nei * 27 + dir2idx(dirneg)];
          memcpy(prts_out + cur, sb->v, sizeof(po_prt) * sb->n);
          cur += sb->n;
        }
      }
    }
    /* :456-468 remote: by sender rank, then the sender's send_entry order =
     * sender patch ascending, sender direction ascending (:213-232).  Patch
     * ranges per rank are contiguous, so ascending global sender patch index
     * already is (rank, local patch) order. */
    if (rank_of_patch) {
      for (int q = 0; q < np; q++) {
        if (rank_of_patch[q] == my_rank) {
          continue;
        }
        for (dir[2] = -1; dir[2] <= 1; dir[2]++) {
          for (dir[1] = -1; dir[1] <= 1; dir[1]++) {
            for (dir[0] = -1; dir[0] <= 1; dir[0]++) {
              if (dir[0] == 0 && dir[1] == 0 && dir[2] == 0) {
                continue;
              }
              if (po_neighbor_patch(g, q, dir) != p) {
                continue;
              }
              prt_vec* sb = &send[q * 27 + dir2idx(dir)];
              memcpy(prts_out + cur, sb->v, sizeof(po_prt) * sb->n);
              cur += sb->n;
            }
          }
        }
      }
    }
  }
  if (n_dropped) {
    *n_dropped = dropped;
  }
  for (int i = 0; i < np * 27; i++) {
    free(send[i].v);
  }
  for (int p = 0; p < np; p++) {
    free(stay[p].v);
  }
  free(send);
  free(stay);
}

/* ====================================================================== */
/* field ghost exchange */

typedef struct
{
  int patch, nei_patch;
  int ilo[3], ihi[3];
} sr_entry;

/* mrc_ddc_multi.c:60-96 (outside), 101-135 (inside) */
static int init_box(const po_grid* g, int p, const int dir[3], int inside,
                    sr_entry* e)
{
  if (dir[0] == 0 && dir[1] == 0 && dir[2] == 0) {
    return 0;
  }
  int nei = po_neighbor_patch(g, p, dir);
  if (nei < 0) {
    return 0;
  }
  e->patch = p;
  e->nei_patch = nei;
  for (int d = 0; d < 3; d++) {
    int ilo = 0, ihi = g->ldims[d], bn = g->ibn[d];
    switch (dir[d]) {
      case -1:
        e->ilo[d] = inside ? ilo : ilo - bn;
        e->ihi[d] = inside ? ilo + bn : ilo;
        break;
      case 0:
        e->ilo[d] = ilo;
        e->ihi[d] = ihi;
        break;
      case 1:
        e->ilo[d] = inside ? ihi - bn : ihi;
        e->ihi[d] = inside ? ihi : ihi + bn;
        break;
    }
  }
  return 1;
}

/* ddc_run_local (mrc_ddc_multi.c:519-538) with copy_to_buf / {copy,add}_from_buf
 * (psc_bnd_impl.hxx:20-78); single rank => every entry is local.  The i-th
 * send entry pairs with the i-th recv entry after the reordering at
 * mrc_ddc_multi.c:355-372, i.e. recv box of nei_patch in direction -dir. */
static void ddc_run(const po_grid* g, float* flds, int n_comps, int mb, int me,
                    int send_inside, int add)
{
  long plen = po_fld_patch_len(g) * n_comps;
  size_t maxbuf = (size_t)g->im[0] * g->im[1] * g->im[2] * (me - mb);
  float* buf = malloc(sizeof(float) * (maxbuf ? maxbuf : 1));
  for (int p = 0; p < g->n_patches; p++) {
    int dir[3];
    for (dir[2] = -1; dir[2] <= 1; dir[2]++) {
      for (dir[1] = -1; dir[1] <= 1; dir[1]++) {
        for (dir[0] = -1; dir[0] <= 1; dir[0]++) {
          sr_entry se = {0}, re = {0};
          if (!init_box(g, p, dir, send_inside, &se)) {
            continue;
          }
          int dirneg[3] = {-dir[0], -dir[1], -dir[2]};
          int ok = init_box(g, se.nei_patch, dirneg, !send_inside, &re);
          assert(ok && re.nei_patch == p);
          (void)ok;
          if (se.ilo[0] == se.ihi[0] || se.ilo[1] == se.ihi[1] ||
              se.ilo[2] == se.ihi[2]) {
            continue;
          }
          float* fs = flds + se.patch * plen;
          float* fr = flds + re.patch * plen;
          long n = 0;
          for (int m = mb; m < me; m++) {
            for (int iz = se.ilo[2]; iz < se.ihi[2]; iz++) {
              for (int iy = se.ilo[1]; iy < se.ihi[1]; iy++) {
                for (int ix = se.ilo[0]; ix < se.ihi[0]; ix++) {
                  buf[n++] = FLD(fs, g, m, ix, iy, iz);
                }
              }
            }
          }
          n = 0;
          for (int m = mb; m < me; m++) {
            for (int iz = re.ilo[2]; iz < re.ihi[2]; iz++) {
              for (int iy = re.ilo[1]; iy < re.ihi[1]; iy++) {
                for (int ix = re.ilo[0]; ix < re.ihi[0]; ix++) {
                  if (add) {
                    FLD(fr, g, m, ix, iy, iz) += buf[n++];
                  } else {
                    FLD(fr, g, m, ix, iy, iz) = buf[n++];
                  }
                }
              }
            }
          }
        }
      }
    }
  }
  free(buf);
}

/* mrc_ddc_multi.c:426-427: fill = send inside -> recv outside */
void po_fill_ghosts(const po_grid* g, float* flds, int n_comps, int mb, int me)
{
  ddc_run(g, flds, n_comps, mb, me, 1, 0);
}

/* mrc_ddc_multi.c:428-429: add = send outside -> += inside */
void po_add_ghosts(const po_grid* g, float* flds, int n_comps, int mb, int me)
{
  ddc_run(g, flds, n_comps, mb, me, 0, 1);
}

/* ====================================================================== */
/* Yee push */

/* psc_push_fields_impl.hxx:24-43,50-129; loop bounds grid.hxx:124-139 */
static void push_fields(const po_grid* g, float* flds, double dt_fac, int is_E)
{
  float dth = (float)(dt_fac * g->dt);
  float cnx = g->invar[0] ? 0 : (float)((double)dth / g->dx[0]);
  float cny = g->invar[1] ? 0 : (float)((double)dth / g->dx[1]);
  float cnz = g->invar[2] ? 0 : (float)((double)dth / g->dx[2]);
  int l = is_E ? 1 : 2, r = is_E ? 2 : 1; /* :156, :175 */
  int ilo[3], ihi[3];
  for (int d = 0; d < 3; d++) {
    ilo[d] = g->invar[d] ? 0 : -l;
    ihi[d] = g->ldims[d] + (g->invar[d] ? 0 : r);
  }
  long plen = po_fld_patch_len(g) * PO_NR_FIELDS;
  for (int p = 0; p < g->n_patches; p++) {
    float* F = flds + p * plen;
    for (int k = ilo[2]; k < ihi[2]; k++) {
      for (int j = ilo[1]; j < ihi[1]; j++) {
        for (int i = ilo[0]; i < ihi[0]; i++) {
          if (is_E) {
            FLDI(F, g, PO_EX, i, j, k) +=
              (cny * (FLDI(F, g, PO_HZ, i, j, k) - FLDI(F, g, PO_HZ, i, j - 1, k)) -
               cnz * (FLDI(F, g, PO_HY, i, j, k) - FLDI(F, g, PO_HY, i, j, k - 1)) -
               dth * FLDI(F, g, PO_JXI, i, j, k));
            FLDI(F, g, PO_EY, i, j, k) +=
              (cnz * (FLDI(F, g, PO_HX, i, j, k) - FLDI(F, g, PO_HX, i, j, k - 1)) -
               cnx * (FLDI(F, g, PO_HZ, i, j, k) - FLDI(F, g, PO_HZ, i - 1, j, k)) -
               dth * FLDI(F, g, PO_JYI, i, j, k));
            FLDI(F, g, PO_EZ, i, j, k) +=
              (cnx * (FLDI(F, g, PO_HY, i, j, k) - FLDI(F, g, PO_HY, i - 1, j, k)) -
               cny * (FLDI(F, g, PO_HX, i, j, k) - FLDI(F, g, PO_HX, i, j - 1, k)) -
               dth * FLDI(F, g, PO_JZI, i, j, k));
          } else {
            FLDI(F, g, PO_HX, i, j, k) -=
              (cny * (FLDI(F, g, PO_EZ, i, j + 1, k) - FLDI(F, g, PO_EZ, i, j, k)) -
               cnz * (FLDI(F, g, PO_EY, i, j, k + 1) - FLDI(F, g, PO_EY, i, j, k)));
            FLDI(F, g, PO_HY, i, j, k) -=
              (cnz * (FLDI(F, g, PO_EX, i, j, k + 1) - FLDI(F, g, PO_EX, i, j, k)) -
               cnx * (FLDI(F, g, PO_EZ, i + 1, j, k) - FLDI(F, g, PO_EZ, i, j, k)));
            FLDI(F, g, PO_HZ, i, j, k) -=
              (cnx * (FLDI(F, g, PO_EY, i + 1, j, k) - FLDI(F, g, PO_EY, i, j, k)) -
               cny * (FLDI(F, g, PO_EX, i, j + 1, k) - FLDI(F, g, PO_EX, i, j, k)));
          }
        }
      }
    }
  }
}

void po_push_E(const po_grid* g, float* flds, double dt_fac)
{
  push_fields(g, flds, dt_fac, 1);
}

void po_push_H(const po_grid* g, float* flds, double dt_fac)
{
  push_fields(g, flds, dt_fac, 0);
}

/* ====================================================================== */
/* conducting-wall field boundaries: psc_bnd_fields_impl.hxx:301-530 */

static void x_range(const po_grid* g, int* x0, int* x1)
{ /* for (ix = max(-2, ib[0]); ix < min(ldims[0] + 2, ib[0] + im[0]); ix++) */
  *x0 = -2 > g->ib[0] ? -2 : g->ib[0];
  int a = g->ldims[0] + 2, b = g->ib[0] + g->im[0];
  *x1 = a < b ? a : b;
}

static void cw_E(const po_grid* g, float* F, int d, int hi)
{
  int x0, x1;
  x_range(g, &x0, &x1);
  const int* ld = g->ldims;
  assert(d == 1 || d == 2);
  if (d == 1) {
    int my = ld[1];
    for (int iz = -2; iz < ld[2] + 2; iz++) {
      for (int ix = x0; ix < x1; ix++) {
        if (!hi) {
          FLDI(F, g, PO_EX, ix, 0, iz) = 0.;
          FLDI(F, g, PO_EX, ix, -1, iz) = FLDI(F, g, PO_EX, ix, 1, iz);
          FLDI(F, g, PO_EY, ix, -1, iz) = -FLDI(F, g, PO_EY, ix, 0, iz);
          FLDI(F, g, PO_EZ, ix, 0, iz) = 0.;
          FLDI(F, g, PO_EZ, ix, -1, iz) = FLDI(F, g, PO_EZ, ix, 1, iz);
        } else {
          FLDI(F, g, PO_EX, ix, my, iz) = 0.;
          FLDI(F, g, PO_EX, ix, my + 1, iz) = FLDI(F, g, PO_EX, ix, my - 1, iz);
          FLDI(F, g, PO_EY, ix, my, iz) = -FLDI(F, g, PO_EY, ix, my - 1, iz);
          FLDI(F, g, PO_EZ, ix, my, iz) = 0.;
          FLDI(F, g, PO_EZ, ix, my + 1, iz) = FLDI(F, g, PO_EZ, ix, my - 1, iz);
        }
      }
    }
  } else {
    int mz = ld[2];
    for (int iy = -2; iy < ld[1] + 2; iy++) {
      for (int ix = x0; ix < x1; ix++) {
        if (!hi) {
          FLDI(F, g, PO_EX, ix, iy, 0) = 0.;
          FLDI(F, g, PO_EX, ix, iy, -1) = FLDI(F, g, PO_EX, ix, iy, 1);
          FLDI(F, g, PO_EY, ix, iy, 0) = 0.;
          FLDI(F, g, PO_EY, ix, iy, -1) = FLDI(F, g, PO_EY, ix, iy, 1);
          FLDI(F, g, PO_EZ, ix, iy, -1) = -FLDI(F, g, PO_EZ, ix, iy, 0);
        } else {
          FLDI(F, g, PO_EX, ix, iy, mz) = 0.;
          FLDI(F, g, PO_EX, ix, iy, mz + 1) = FLDI(F, g, PO_EX, ix, iy, mz - 1);
          FLDI(F, g, PO_EY, ix, iy, mz) = 0.;
          FLDI(F, g, PO_EY, ix, iy, mz + 1) = FLDI(F, g, PO_EY, ix, iy, mz - 1);
          FLDI(F, g, PO_EZ, ix, iy, mz) = -FLDI(F, g, PO_EZ, ix, iy, mz - 1);
        }
      }
    }
  }
}

static void cw_H(const po_grid* g, float* F, int d, int hi)
{
  int x0, x1;
  x_range(g, &x0, &x1);
  const int* ld = g->ldims;
  assert(d == 1 || d == 2);
  if (d == 1) {
    int my = ld[1];
    /* NB: lo loop starts at iz = -1 upstream (:393), hi at -2 (:434) */
    for (int iz = hi ? -2 : -1; iz < ld[2] + 2; iz++) {
      for (int ix = x0; ix < x1; ix++) {
        if (!hi) {
          FLDI(F, g, PO_HX, ix, -1, iz) = -FLDI(F, g, PO_HX, ix, 0, iz);
          FLDI(F, g, PO_HY, ix, -1, iz) = FLDI(F, g, PO_HY, ix, 1, iz);
          FLDI(F, g, PO_HZ, ix, -1, iz) = -FLDI(F, g, PO_HZ, ix, 0, iz);
        } else {
          FLDI(F, g, PO_HX, ix, my, iz) = -FLDI(F, g, PO_HX, ix, my - 1, iz);
          FLDI(F, g, PO_HY, ix, my + 1, iz) = FLDI(F, g, PO_HY, ix, my - 1, iz);
          FLDI(F, g, PO_HZ, ix, my, iz) = -FLDI(F, g, PO_HZ, ix, my - 1, iz);
        }
      }
    }
  } else {
    int mz = ld[2];
    for (int iy = -2; iy < ld[1] + 2; iy++) {
      for (int ix = x0; ix < x1; ix++) {
        if (!hi) {
          FLDI(F, g, PO_HX, ix, iy, -1) = -FLDI(F, g, PO_HX, ix, iy, 0);
          FLDI(F, g, PO_HY, ix, iy, -1) = -FLDI(F, g, PO_HY, ix, iy, 0);
          FLDI(F, g, PO_HZ, ix, iy, -1) = FLDI(F, g, PO_HZ, ix, iy, 1);
        } else {
          FLDI(F, g, PO_HX, ix, iy, mz) = -FLDI(F, g, PO_HX, ix, iy, mz - 1);
          FLDI(F, g, PO_HY, ix, iy, mz) = -FLDI(F, g, PO_HY, ix, iy, mz - 1);
          FLDI(F, g, PO_HZ, ix, iy, mz + 1) = FLDI(F, g, PO_HZ, ix, iy, mz - 1);
        }
      }
    }
  }
}

static void cw_J(const po_grid* g, float* F, int d, int hi)
{
  int x0, x1;
  x_range(g, &x0, &x1);
  const int* ld = g->ldims;
  assert(d == 1 || d == 2);
  if (d == 1) {
    int my = ld[1];
    for (int iz = -2; iz < ld[2] + 2; iz++) {
      for (int ix = x0; ix < x1; ix++) {
        if (!hi) {
          FLDI(F, g, PO_JXI, ix, 1, iz) += FLDI(F, g, PO_JXI, ix, -1, iz);
          FLDI(F, g, PO_JXI, ix, -1, iz) = 0.;
          FLDI(F, g, PO_JYI, ix, 0, iz) -= FLDI(F, g, PO_JYI, ix, -1, iz);
          FLDI(F, g, PO_JYI, ix, -1, iz) = 0.;
          FLDI(F, g, PO_JZI, ix, 1, iz) += FLDI(F, g, PO_JZI, ix, -1, iz);
          FLDI(F, g, PO_JZI, ix, -1, iz) = 0.;
        } else {
          FLDI(F, g, PO_JXI, ix, my - 1, iz) += FLDI(F, g, PO_JXI, ix, my + 1, iz);
          FLDI(F, g, PO_JXI, ix, my + 1, iz) = 0.;
          FLDI(F, g, PO_JYI, ix, my - 1, iz) -= FLDI(F, g, PO_JYI, ix, my, iz);
          FLDI(F, g, PO_JYI, ix, my, iz) = 0.;
          FLDI(F, g, PO_JZI, ix, my - 1, iz) += FLDI(F, g, PO_JZI, ix, my + 1, iz);
          FLDI(F, g, PO_JZI, ix, my + 1, iz) = 0.;
        }
      }
    }
  } else {
    int mz = ld[2];
    for (int iy = -2; iy < ld[1] + 2; iy++) {
      for (int ix = x0; ix < x1; ix++) {
        if (!hi) {
          FLDI(F, g, PO_JXI, ix, iy, 1) += FLDI(F, g, PO_JXI, ix, iy, -1);
          FLDI(F, g, PO_JXI, ix, iy, -1) = 0.;
          FLDI(F, g, PO_JYI, ix, iy, 1) += FLDI(F, g, PO_JYI, ix, iy, -1);
          FLDI(F, g, PO_JYI, ix, iy, -1) = 0.;
          FLDI(F, g, PO_JZI, ix, iy, 0) -= FLDI(F, g, PO_JZI, ix, iy, -1);
          FLDI(F, g, PO_JZI, ix, iy, -1) = 0.;
        } else {
          FLDI(F, g, PO_JXI, ix, iy, mz - 1) += FLDI(F, g, PO_JXI, ix, iy, mz + 1);
          FLDI(F, g, PO_JXI, ix, iy, mz + 1) = 0.;
          FLDI(F, g, PO_JYI, ix, iy, mz - 1) += FLDI(F, g, PO_JYI, ix, iy, mz + 1);
          FLDI(F, g, PO_JYI, ix, iy, mz + 1) = 0.;
          FLDI(F, g, PO_JZI, ix, iy, mz - 1) -= FLDI(F, g, PO_JZI, ix, iy, mz);
          FLDI(F, g, PO_JZI, ix, iy, mz) = 0.;
        }
      }
    }
  }
}

/* ---- BND_FLD_OPEN (psc_bnd_fields_impl.hxx:210-300 set_lower/upper_ghosts with
 * include_edge = false, :535-640 radiative_H_lo/hi; no incoming pulse, background_e = background_h = 0
 * as the decks leave them).  E: every ghost of the three E components behind the wall is set to the
 * background.  H: first-order absorbing condition for the two tangential components in the first
 * ghost plane.  The reference's loop also evaluates H0(edge - 1) one index below the array at the
 * outermost transverse ghost line (an out-of-bounds read there); the restatement takes the
 * difference as zero on that line -- nothing reads those corner values. */
static void open_E(const po_grid* g, float* F, int d, int hi)
{
  int lo3[3], hi3[3];
  for (int a = 0; a < 3; a++) {
    lo3[a] = g->ib[a];
    hi3[a] = g->ib[a] + g->im[a];
  }
  if (g->invar[d]) {
    return;
  }
  if (!hi) {
    hi3[d] = 0; /* stop[d] = 0 (:223) */
  } else {
    lo3[d] = g->ldims[d] + 1; /* start[d] = ldims + 1 (:270) */
  }
  for (int m = PO_EX; m < PO_EX + 3; m++) {
    for (int k = lo3[2]; k < hi3[2]; k++) {
      for (int j = lo3[1]; j < hi3[1]; j++) {
        for (int i = lo3[0]; i < hi3[0]; i++) {
          FLD(F, g, m, i, j, k) = 0.f;
        }
      }
    }
  }
  if (hi) {
    /* upper edge plane (:281-296): the component normal to the wall is not on the edge */
    lo3[d] = g->ldims[d];
    hi3[d] = g->ldims[d] + 1;
    const int m = PO_EX + d;
    for (int k = lo3[2]; k < hi3[2]; k++) {
      for (int j = lo3[1]; j < hi3[1]; j++) {
        for (int i = lo3[0]; i < hi3[0]; i++) {
          FLD(F, g, m, i, j, k) = 0.f;
        }
      }
    }
  }
}

static void open_H(const po_grid* g, float* F, int d, int hi)
{
  if (g->invar[d]) {
    return;
  }
  const float dt = (float)g->dt;
  float dtdx[3];
  for (int a = 0; a < 3; a++) {
    dtdx[a] = dt * (float)g->dx_inv[a];
  }
  const int d0 = d, d1 = (d + 1) % 3, d2 = (d + 2) % 3;
  const int H0 = PO_HX + d0, H1 = PO_HX + d1, H2 = PO_HX + d2;
  const int E1 = PO_EX + d1, E2 = PO_EX + d2;
  const int J1 = PO_JXI + d1, J2 = PO_JXI + d2;
  int lo3[3], hi3[3];
  for (int a = 0; a < 3; a++) {
    lo3[a] = g->ib[a];
    hi3[a] = g->ib[a] + g->im[a];
  }
  lo3[d0] = hi ? g->ldims[d0] : -1;
  hi3[d0] = lo3[d0] + 1;
#define AT(m, q) FLD(F, g, m, (q)[0], (q)[1], (q)[2])
  for (int k = lo3[2]; k < hi3[2]; k++) {
    for (int j = lo3[1]; j < hi3[1]; j++) {
      for (int i = lo3[0]; i < hi3[0]; i++) {
        int i3[3] = {i, j, k}, e[3] = {i, j, k}, e_m2[3], e_m1[3];
        e[d0] += hi ? -1 : 1; /* edge_idx */
        /* the point the H0 differences are taken at: edge_idx (lo) / i3 (hi) */
        int* at = hi ? i3 : e;
        for (int a = 0; a < 3; a++) {
          e_m2[a] = e_m1[a] = at[a];
        }
        float dH0_2 = 0.f, dH0_1 = 0.f;
        if (!g->invar[d2] && at[d2] - 1 >= g->ib[d2]) {
          e_m2[d2] -= 1;
          dH0_2 = AT(H0, at) - AT(H0, e_m2);
        }
        if (!g->invar[d1] && at[d1] - 1 >= g->ib[d1]) {
          e_m1[d1] -= 1;
          dH0_1 = AT(H0, at) - AT(H0, e_m1);
        }
        if (!hi) {
          AT(H2, i3) = (4.f * 0.f - 2.f * (AT(E1, e) - 0.f) - dtdx[d2] * dH0_2 - (1.f - dtdx[d0]) * (AT(H2, e) - 0.f) +
                        dt * AT(J1, e)) /
                         (1.f + dtdx[d0]) +
                       0.f;
          AT(H1, i3) = (-4.f * 0.f + 2.f * (AT(E2, e) - 0.f) - dtdx[d1] * dH0_1 - (1.f - dtdx[d0]) * (AT(H1, e) - 0.f) +
                        dt * AT(J2, e)) /
                         (1.f + dtdx[d0]) +
                       0.f;
        } else {
          AT(H2, i3) = (-4.f * 0.f + 2.f * (AT(E1, i3) - 0.f) + dtdx[d2] * dH0_2 - (1.f - dtdx[d0]) * (AT(H2, e) - 0.f) -
                        dt * AT(J1, i3)) /
                         (1.f + dtdx[d0]) +
                       0.f;
          AT(H1, i3) = (4.f * 0.f - 2.f * (AT(E2, i3) - 0.f) + dtdx[d1] * dH0_1 - (1.f - dtdx[d0]) * (AT(H1, e) - 0.f) -
                        dt * AT(J2, i3)) /
                         (1.f + dtdx[d0]) +
                       0.f;
        }
      }
    }
  }
#undef AT
}

static void open_J(const po_grid* g, float* F, int d, int hi)
{ /* add_ghosts_J: BND_FLD_OPEN does nothing (:169-171) */
  (void)g, (void)F, (void)d, (void)hi;
}

/* psc_bnd_fields_impl.hxx:27-188: per patch, lo for d=0..2 then hi for d=0..2 */
static void bndf_apply(const po_grid* g, float* flds,
                       void (*cw)(const po_grid*, float*, int, int),
                       void (*open)(const po_grid*, float*, int, int))
{
  long plen = po_fld_patch_len(g) * PO_NR_FIELDS;
  for (int p = 0; p < g->n_patches; p++) {
    float* F = flds + p * plen;
    for (int d = 0; d < 3; d++) {
      if (at_boundary_lo(g, p, d)) {
        if (g->bc_fld_lo[d] == PO_BND_FLD_CONDUCTING_WALL) {
          cw(g, F, d, 0);
        } else if (g->bc_fld_lo[d] == PO_BND_FLD_OPEN) {
          open(g, F, d, 0);
        }
      }
    }
    for (int d = 0; d < 3; d++) {
      if (at_boundary_hi(g, p, d)) {
        if (g->bc_fld_hi[d] == PO_BND_FLD_CONDUCTING_WALL) {
          cw(g, F, d, 1);
        } else if (g->bc_fld_hi[d] == PO_BND_FLD_OPEN) {
          open(g, F, d, 1);
        }
      }
    }
  }
}

void po_bndf_fill_ghosts_E(const po_grid* g, float* flds)
{
  bndf_apply(g, flds, cw_E, open_E);
}
void po_bndf_fill_ghosts_H(const po_grid* g, float* flds)
{
  bndf_apply(g, flds, cw_H, open_H);
}
void po_bndf_add_ghosts_J(const po_grid* g, float* flds)
{
  bndf_apply(g, flds, cw_J, open_J);
}

/* ====================================================================== */
/* rho moment, div, checks, Marder, energies */

/* 1-component scalar array with the grid's im/ib */
#define SC(f, g, i, j, k) FLD(f, g, 0, i, j, k)

/* add_ghosts_reflecting.hxx:77-154 (node-centred) */
static void add_ghosts_reflecting_nc(const po_grid* g, float* R, int d, int hi)
{
  int unused[3], b[3], e[3];
  for (int a = 0; a < 3; a++) {
    unused[a] = !!g->ib[a];
    b[a] = g->ib[a] + unused[a];
    e[a] = g->ldims[a] - g->ib[a];
  }
  if (!hi) {
    b[d] = 1;
    e[d] = 1 - g->ib[d] - unused[d];
  } else {
    b[d] = g->ldims[d] + g->ib[d] + unused[d];
    e[d] = g->ldims[d];
  }
  int idx[3];
  for (idx[2] = b[2]; idx[2] < e[2]; idx[2]++) {
    for (idx[1] = b[1]; idx[1] < e[1]; idx[1]++) {
      for (idx[0] = b[0]; idx[0] < e[0]; idx[0]++) {
        int r[3] = {idx[0], idx[1], idx[2]};
        r[d] = hi ? 2 * g->ldims[d] - idx[d] : -idx[d];
        SC(R, g, idx[0], idx[1], idx[2]) += SC(R, g, r[0], r[1], r[2]);
      }
    }
  }
}

/* Moment_rho_1st_nc = ItemMoment<moment_rho<Deposit1stNc>> :
 * fields_item.hxx:97-134 (zeros, moment, add_ghosts),
 * psc/moment.hxx:72-81,149-171, psc/deposit.hxx:24-65,172-191,262-285,
 * const_accessor_simple.hxx:60-63 */
void po_moment_rho_1st_nc(const po_grid* g, const po_prt* prts,
                          const unsigned* off, float* rho)
{
  long plen = po_fld_patch_len(g);
  int yz = g->invar[0];
  memset(rho, 0, sizeof(float) * plen * g->n_patches);
  float dxi[3];
  for (int d = 0; d < 3; d++) {
    dxi[d] = 1.f / (float)g->dx[d]; /* deposit.hxx:272 real_t(1.)/dx */
  }
  float fnqs = (float)g->fnqs;
  for (int p = 0; p < g->n_patches; p++) {
    float* R = rho + p * plen;
    for (unsigned n = off[p]; n < off[p + 1]; n++) {
      const po_prt* prt = &prts[n];
      float q = (float)g->q[prt->kind];
      float w = prt->qni_wni / q;
      float val = w * q;
      float value = fnqs * val;
      int l[3];
      float h[3];
      for (int d = 0; d < 3; d++) {
        float x = prt->x[d] * dxi[d];
        l[d] = (int)floorf(x);
        h[d] = x - (float)l[d];
      }
      if (yz) {
        SC(R, g, 0, l[1] + 0, l[2] + 0) += value * (1.f - h[1]) * (1.f - h[2]);
        SC(R, g, 0, l[1] + 1, l[2] + 0) += value * h[1] * (1.f - h[2]);
        SC(R, g, 0, l[1] + 0, l[2] + 1) += value * (1.f - h[1]) * h[2];
        SC(R, g, 0, l[1] + 1, l[2] + 1) += value * h[1] * h[2];
      } else {
        /* clang-format off */
        SC(R, g, l[0] + 0, l[1] + 0, l[2] + 0) += value * (1.f - h[0]) * (1.f - h[1]) * (1.f - h[2]);
        SC(R, g, l[0] + 1, l[1] + 0, l[2] + 0) += value *        h[0]  * (1.f - h[1]) * (1.f - h[2]);
        SC(R, g, l[0] + 0, l[1] + 1, l[2] + 0) += value * (1.f - h[0]) *        h[1]  * (1.f - h[2]);
        SC(R, g, l[0] + 1, l[1] + 1, l[2] + 0) += value *        h[0]  *        h[1]  * (1.f - h[2]);
        SC(R, g, l[0] + 0, l[1] + 0, l[2] + 1) += value * (1.f - h[0]) * (1.f - h[1]) *        h[2];
        SC(R, g, l[0] + 1, l[1] + 0, l[2] + 1) += value *        h[0]  * (1.f - h[1]) *        h[2];
        SC(R, g, l[0] + 0, l[1] + 1, l[2] + 1) += value * (1.f - h[0]) *        h[1]  *        h[2];
        SC(R, g, l[0] + 1, l[1] + 1, l[2] + 1) += value *        h[0]  *        h[1]  *        h[2];
        /* clang-format on */
      }
    }
  }
  /* ItemMomentBnd::add_ghosts, fields_item.hxx:36-90 */
  for (int p = 0; p < g->n_patches; p++) {
    float* R = rho + p * plen;
    for (int d = 0; d < 3; d++) {
      if (at_boundary_lo(g, p, d) && g->bc_prt_lo[d] == PO_BND_PRT_REFLECTING) {
        add_ghosts_reflecting_nc(g, R, d, 0);
      }
    }
    for (int d = 0; d < 3; d++) {
      if (at_boundary_hi(g, p, d) && g->bc_prt_hi[d] == PO_BND_PRT_REFLECTING) {
        add_ghosts_reflecting_nc(g, R, d, 1);
      }
    }
  }
  po_add_ghosts(g, rho, 1, 0, 1);
}

/* ---------------------------------------------------------------------- */
/* the 1st-order moment family:
 * Moment_n_1st / Moment_v_1st / Moment_p_1st / Moment_T_1st / Moments_1st
 * (fields_item_moments_1st.hxx:9-30) = ItemMoment<moment_*<Deposit1stCc>>, and
 * Moment_rho_1st_nc (:35-37).  psc/moment.hxx:119-311 (what is deposited, in which
 * order), psc/deposit.hxx:24-65 (weights), :172-212 (cell / node centring), :262-285
 * (code units: x / dx, fnqs * val), fields_item.hxx:36-134 (zeros, moment, reflecting
 * folds, add_ghosts), add_ghosts_reflecting.hxx:7-70 (cc), :72-154 (nc). */

int po_moment_n_comps(const po_grid* g, int which)
{
  switch (which) {
    case PO_MOM_N: return g->n_kinds;
    case PO_MOM_V: return 3 * g->n_kinds;
    case PO_MOM_P: return 3 * g->n_kinds;
    case PO_MOM_T: return 6 * g->n_kinds;
    case PO_MOM_ALL: return 13 * g->n_kinds;
    case PO_MOM_RHO_NC: return 1;
  }
  return 0;
}

/* add_ghosts_reflecting.hxx:7-70 (cell-centred), component m of a patch */
static void add_ghosts_reflecting_cc(const po_grid* g, float* R, int m, int d, int hi)
{
  int b[3], e[3];
  for (int a = 0; a < 3; a++) {
    b[a] = g->ib[a];
    e[a] = g->ldims[a] - g->ib[a];
  }
  if (!hi) {
    b[d] = 0;
    e[d] = -g->ib[d];
  } else {
    b[d] = g->ldims[d] + g->ib[d];
    e[d] = g->ldims[d];
  }
  int idx[3];
  for (idx[2] = b[2]; idx[2] < e[2]; idx[2]++) {
    for (idx[1] = b[1]; idx[1] < e[1]; idx[1]++) {
      for (idx[0] = b[0]; idx[0] < e[0]; idx[0]++) {
        int r[3] = {idx[0], idx[1], idx[2]};
        r[d] = hi ? 2 * g->ldims[d] - idx[d] - 1 : -idx[d] - 1;
        FLD(R, g, m, idx[0], idx[1], idx[2]) += FLD(R, g, m, r[0], r[1], r[2]);
      }
    }
  }
}

static void deposit_1st(const po_grid* g, float* R, int m, const int l[3], const float h[3],
                        float value)
{
  if (g->invar[0]) {
    FLD(R, g, m, 0, l[1] + 0, l[2] + 0) += value * (1.f - h[1]) * (1.f - h[2]);
    FLD(R, g, m, 0, l[1] + 1, l[2] + 0) += value * h[1] * (1.f - h[2]);
    FLD(R, g, m, 0, l[1] + 0, l[2] + 1) += value * (1.f - h[1]) * h[2];
    FLD(R, g, m, 0, l[1] + 1, l[2] + 1) += value * h[1] * h[2];
  } else {
    /* clang-format off */
    FLD(R, g, m, l[0] + 0, l[1] + 0, l[2] + 0) += value * (1.f - h[0]) * (1.f - h[1]) * (1.f - h[2]);
    FLD(R, g, m, l[0] + 1, l[1] + 0, l[2] + 0) += value *        h[0]  * (1.f - h[1]) * (1.f - h[2]);
    FLD(R, g, m, l[0] + 0, l[1] + 1, l[2] + 0) += value * (1.f - h[0]) *        h[1]  * (1.f - h[2]);
    FLD(R, g, m, l[0] + 1, l[1] + 1, l[2] + 0) += value *        h[0]  *        h[1]  * (1.f - h[2]);
    FLD(R, g, m, l[0] + 0, l[1] + 0, l[2] + 1) += value * (1.f - h[0]) * (1.f - h[1]) *        h[2];
    FLD(R, g, m, l[0] + 1, l[1] + 0, l[2] + 1) += value *        h[0]  * (1.f - h[1]) *        h[2];
    FLD(R, g, m, l[0] + 0, l[1] + 1, l[2] + 1) += value * (1.f - h[0]) *        h[1]  *        h[2];
    FLD(R, g, m, l[0] + 1, l[1] + 1, l[2] + 1) += value *        h[0]  *        h[1]  *        h[2];
    /* clang-format on */
  }
}

void po_moment_1st(const po_grid* g, const po_prt* prts, const unsigned* off, int which,
                   float* out)
{
  const int nc = po_moment_n_comps(g, which);
  const long plen = po_fld_patch_len(g) * nc;
  const int cc = which != PO_MOM_RHO_NC;
  memset(out, 0, sizeof(float) * plen * g->n_patches);
  float dxi[3];
  for (int d = 0; d < 3; d++) {
    dxi[d] = 1.f / (float)g->dx[d]; /* deposit.hxx:272 real_t(1.)/dx */
  }
  const float fnqs = (float)g->fnqs;
  for (int p = 0; p < g->n_patches; p++) {
    float* R = out + p * plen;
    for (unsigned n = off[p]; n < off[p + 1]; n++) {
      const po_prt* prt = &prts[n];
      const int kind = prt->kind;
      const float q = (float)g->q[kind], m = (float)g->m[kind];
      const float w = prt->qni_wni / q; /* const_accessor_simple.hxx:60-63 */
      const float* u = prt->u;
      int l[3];
      float h[3];
      for (int d = 0; d < 3; d++) {
        float x = prt->x[d] * dxi[d];
        if (cc) {
          l[d] = (int)floorf(x - .5f);
          h[d] = x - .5f - (float)l[d];
        } else {
          l[d] = (int)floorf(x);
          h[d] = x - (float)l[d];
        }
      }
      float vxi[3];
      {
        float root = 1.f / sqrtf(1.f + u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
        for (int d = 0; d < 3; d++) {
          vxi[d] = u[d] * root;
        }
      }
      int mm[13];
      float val[13];
      int nv = 0;
      switch (which) {
        case PO_MOM_N: mm[0] = kind, val[0] = w, nv = 1; break;
        case PO_MOM_RHO_NC: mm[0] = 0, val[0] = w * q, nv = 1; break;
        case PO_MOM_V:
          for (int d = 0; d < 3; d++) {
            mm[d] = d + 3 * kind, val[d] = w * vxi[d];
          }
          nv = 3;
          break;
        case PO_MOM_P:
          for (int d = 0; d < 3; d++) {
            mm[d] = d + 3 * kind, val[d] = w * m * u[d];
          }
          nv = 3;
          break;
        case PO_MOM_T: {
          const int a[6] = {0, 1, 2, 0, 0, 1}, b[6] = {0, 1, 2, 1, 2, 2};
          for (int k = 0; k < 6; k++) {
            mm[k] = 6 * kind + k, val[k] = w * m * u[a[k]] * vxi[b[k]];
          }
          nv = 6;
          break;
        }
        case PO_MOM_ALL: {
          const int a[6] = {0, 1, 2, 0, 1, 2}, b[6] = {0, 1, 2, 1, 2, 0};
          mm[0] = 13 * kind, val[0] = w * q;
          for (int d = 0; d < 3; d++) {
            mm[1 + d] = 13 * kind + 1 + d, val[1 + d] = w * q * vxi[d];
            mm[4 + d] = 13 * kind + 4 + d, val[4 + d] = w * m * u[d];
          }
          for (int k = 0; k < 6; k++) {
            mm[7 + k] = 13 * kind + 7 + k, val[7 + k] = w * m * u[a[k]] * vxi[b[k]];
          }
          nv = 13;
          break;
        }
      }
      for (int k = 0; k < nv; k++) {
        deposit_1st(g, R, mm[k], l, h, fnqs * val[k]);
      }
    }
  }
  /* ItemMomentBnd::add_ghosts, fields_item.hxx:36-90 */
  for (int p = 0; p < g->n_patches; p++) {
    float* R = out + p * plen;
    for (int hi = 0; hi < 2; hi++) {
      for (int d = 0; d < 3; d++) {
        int at = hi ? at_boundary_hi(g, p, d) : at_boundary_lo(g, p, d);
        int bc = hi ? g->bc_prt_hi[d] : g->bc_prt_lo[d];
        if (!at || bc != PO_BND_PRT_REFLECTING) {
          continue;
        }
        if (cc) {
          for (int m = 0; m < nc; m++) {
            add_ghosts_reflecting_cc(g, R, m, d, hi);
          }
        } else {
          add_ghosts_reflecting_nc(g, R, d, hi);
        }
      }
    }
  }
  po_add_ghosts(g, out, nc, 0, nc);
}

/* fields_item_fields.hxx:65-104: float difference, divided by double dx,
 * accumulated through the float result array axis by axis */
void po_div_nc(const po_grid* g, const float* flds, int n_comps, int m0,
               float* div)
{
  long plen1 = po_fld_patch_len(g);
  long plen = plen1 * n_comps;
  memset(div, 0, sizeof(float) * plen1 * g->n_patches);
  for (int p = 0; p < g->n_patches; p++) {
    const float* F = flds + p * plen;
    float* D = div + p * plen1;
    for (int a = 0; a < 3; a++) {
      if (g->invar[a]) {
        continue;
      }
      for (int k = 0; k < g->ldims[2]; k++) {
        for (int j = 0; j < g->ldims[1]; j++) {
          for (int i = 0; i < g->ldims[0]; i++) {
            int im1[3] = {i, j, k};
            im1[a] -= 1;
            float diff =
              FLD(F, g, m0 + a, i, j, k) - FLD(F, g, m0 + a, im1[0], im1[1], im1[2]);
            SC(D, g, i, j, k) =
              (float)((double)SC(D, g, i, j, k) + (double)diff / g->dx[a]);
          }
        }
      }
    }
  }
}

/* checks_impl.hxx:60-97 (no open boundaries) */
double po_continuity(const po_grid* g, const float* rho_m, const float* rho_p,
                     const float* flds)
{
  long plen1 = po_fld_patch_len(g);
  float* divj = malloc(sizeof(float) * plen1 * g->n_patches);
  po_div_nc(g, flds, PO_NR_FIELDS, PO_JXI, divj);
  double err = 0.;
  for (int p = 0; p < g->n_patches; p++) {
    for (int k = 0; k < g->ldims[2]; k++) {
      for (int j = 0; j < g->ldims[1]; j++) {
        for (int i = 0; i < g->ldims[0]; i++) {
          float d_rho = SC(rho_p + p * plen1, g, i, j, k) -
                        SC(rho_m + p * plen1, g, i, j, k);
          double v =
            (double)d_rho + g->dt * (double)SC(divj + p * plen1, g, i, j, k);
          /* checks_impl.hxx:78-90: div j := -d rho / dt in the first cell layer at a lower open boundary */
          {
            int idx3[3] = {i, j, k};
            for (int d = 0; d < 3; d++) {
              if (at_boundary_lo(g, p, d) && idx3[d] == 0 && g->bc_fld_lo[d] == PO_BND_FLD_OPEN) {
                v = 0.;
              }
            }
          }
          if (fabs(v) > err) {
            err = fabs(v);
          }
        }
      }
    }
  }
  free(divj);
  return err;
}

/* checks_impl.hxx:157-184 */
double po_gauss(const po_grid* g, const float* rho, const float* flds)
{
  long plen1 = po_fld_patch_len(g);
  float* dive = malloc(sizeof(float) * plen1 * g->n_patches);
  po_div_nc(g, flds, PO_NR_FIELDS, PO_EX, dive);
  double err = 0.;
  for (int p = 0; p < g->n_patches; p++) {
    for (int k = 0; k < g->ldims[2]; k++) {
      for (int j = 0; j < g->ldims[1]; j++) {
        for (int i = 0; i < g->ldims[0]; i++) {
          int skip = 0;
          int idx[3] = {i, j, k};
          for (int d = 0; d < 3; d++) {
            if (at_boundary_lo(g, p, d) && idx[d] == 0 &&
                (g->bc_fld_lo[d] == PO_BND_FLD_CONDUCTING_WALL ||
                 g->bc_fld_lo[d] == PO_BND_FLD_OPEN)) {
              skip = 1; /* rho := dive there */
            }
          }
          if (skip) {
            continue;
          }
          float v = SC(dive + p * plen1, g, i, j, k) -
                    SC(rho + p * plen1, g, i, j, k);
          if (fabs((double)v) > err) {
            err = fabs((double)v);
          }
        }
      }
    }
  }
  free(dive);
  return err;
}

/* psc::marder::correct (libpsc/psc_push_fields/marder_impl.hxx:26-61):
 * E_d += (res[+1_d] - res) * fac_d over the patch interior,
 * fac = .5f * real_t(dt) * diffusion * Real3(dx_inv); res has 1 component */
void po_marder_apply(const po_grid* g, float* flds, const float* res, double diffusion)
{
  long plen1 = po_fld_patch_len(g);
  long plen = plen1 * PO_NR_FIELDS;
  float diff_f = (float)diffusion;
  float s = .5f * (float)g->dt * diff_f;
  for (int p = 0; p < g->n_patches; p++) {
    float* F = flds + p * plen;
    const float* R = res + p * plen1;
    for (int d = 0; d < 3; d++) {
      if (g->invar[d]) {
        continue;
      }
      float fac = s * (float)g->dx_inv[d];
      for (int k = 0; k < g->ldims[2]; k++) {
        for (int j = 0; j < g->ldims[1]; j++) {
          for (int i = 0; i < g->ldims[0]; i++) {
            int ip[3] = {i, j, k};
            ip[d] += 1;
            FLD(F, g, PO_EX + d, i, j, k) =
              FLD(F, g, PO_EX + d, i, j, k) +
              (SC(R, g, ip[0], ip[1], ip[2]) - SC(R, g, i, j, k)) * fac;
          }
        }
      }
    }
  }
}

/* marder_impl.hxx:197-264 + 26-61 */
void po_marder_correct(const po_grid* g, float* flds, const po_prt* prts,
                       const unsigned* off, double diffusion_, int loop)
{
  long plen1 = po_fld_patch_len(g);
  long ntot = plen1 * g->n_patches;

  double inv_sum = 0.;
  for (int d = 0; d < 3; d++) {
    if (!g->invar[d]) {
      inv_sum += g->dx_inv[d] * g->dx_inv[d];
    }
  }
  double diffusion_max = 1. / 2. / (.5 * g->dt) / inv_sum;
  double diffusion = diffusion_max * (double)(float)diffusion_;

  float* rho = malloc(sizeof(float) * ntot);
  float* dive = malloc(sizeof(float) * ntot);
  float* res = malloc(sizeof(float) * ntot);
  po_moment_rho_1st_nc(g, prts, off, rho);

  for (int it = 0; it < loop; it++) {
    po_fill_ghosts(g, flds, PO_NR_FIELDS, PO_EX, PO_EX + 3);
    po_div_nc(g, flds, PO_NR_FIELDS, PO_EX, dive);
    memset(res, 0, sizeof(float) * ntot);
    for (int p = 0; p < g->n_patches; p++) {
      for (int k = 0; k < g->ldims[2]; k++) {
        for (int j = 0; j < g->ldims[1]; j++) {
          for (int i = 0; i < g->ldims[0]; i++) {
            SC(res + p * plen1, g, i, j, k) =
              SC(dive + p * plen1, g, i, j, k) - SC(rho + p * plen1, g, i, j, k);
          }
        }
      }
      /* :223-250 zero the residual on wall planes */
      for (int d = 0; d < 3; d++) {
        int lo[3] = {0, 0, 0}, hi[3] = {g->ldims[0], g->ldims[1], g->ldims[2]};
        if ((g->bc_fld_lo[d] == PO_BND_FLD_CONDUCTING_WALL ||
             g->bc_fld_lo[d] == PO_BND_FLD_OPEN) &&
            at_boundary_lo(g, p, d)) {
          int l2[3] = {lo[0], lo[1], lo[2]}, h2[3] = {hi[0], hi[1], hi[2]};
          h2[d] = l2[d] + 1;
          for (int k = l2[2]; k < h2[2]; k++)
            for (int j = l2[1]; j < h2[1]; j++)
              for (int i = l2[0]; i < h2[0]; i++)
                SC(res + p * plen1, g, i, j, k) = 0.f;
        }
        if ((g->bc_fld_hi[d] == PO_BND_FLD_CONDUCTING_WALL ||
             g->bc_fld_hi[d] == PO_BND_FLD_OPEN) &&
            at_boundary_hi(g, p, d)) {
          int l2[3] = {lo[0], lo[1], lo[2]}, h2[3] = {hi[0], hi[1], hi[2]};
          l2[d] = hi[d];
          h2[d] = hi[d] + 1;
          for (int k = l2[2]; k < h2[2]; k++)
            for (int j = l2[1]; j < h2[1]; j++)
              for (int i = l2[0]; i < h2[0]; i++)
                SC(res + p * plen1, g, i, j, k) = 0.f;
        }
      }
    }
    po_fill_ghosts(g, res, 1, 0, 1);

    po_marder_apply(g, flds, res, diffusion);
  }
  po_fill_ghosts(g, flds, PO_NR_FIELDS, PO_EX, PO_EX + 3);
  free(rho);
  free(dive);
  free(res);
}

/* DiagEnergiesField.h:19-42, DiagEnergiesParticle.h:15-40
 * out[0..5] field energies, out[6] = E_electron (q<0), out[7] = E_ion (q>0) */
void po_energies(const po_grid* g, const float* flds, const po_prt* prts,
                 const unsigned* off, double* out)
{
  long plen = po_fld_patch_len(g) * PO_NR_FIELDS;
  double fac = g->dx[0] * g->dx[1] * g->dx[2];
  for (int m = 0; m < 8; m++) {
    out[m] = 0.;
  }
  for (int p = 0; p < g->n_patches; p++) {
    const float* F = flds + p * plen;
    for (int k = 0; k < g->ldims[2]; k++) {
      for (int j = 0; j < g->ldims[1]; j++) {
        for (int i = 0; i < g->ldims[0]; i++) {
          for (int m = 0; m < 6; m++) {
            float v = FLD(F, g, PO_EX + m, i, j, k);
            out[m] += (double)(v * v) * fac;
          }
        }
      }
    }
  }
  double fnqs = g->fnqs;
  for (int p = 0; p < g->n_patches; p++) {
    for (unsigned n = off[p]; n < off[p + 1]; n++) {
      const po_prt* prt = &prts[n];
      float qf = (float)g->q[prt->kind];
      float mf = (float)g->m[prt->kind];
      float w = prt->qni_wni / qf;
      double gamma = sqrtf(1.f + sqrf(prt->u[0]) + sqrf(prt->u[1]) + sqrf(prt->u[2]));
      double Ekin = (gamma - 1.) * mf * w * fnqs;
      double q = qf;
      if (q < 0.) {
        out[6] += Ekin * fac;
      } else if (q > 0.) {
        out[7] += Ekin * fac;
      }
    }
  }
}

/* ====================================================================== */
/* balance */

/* best_mapping_recursive (psc_balance_impl.hxx:99-151) */
static void po_map_rec(const double* load_by_patch, int* n_by_proc, int proc_begin,
                       int proc_end, int patch_begin, int patch_end)
{
  assert(patch_end - patch_begin >= proc_end - proc_begin);
  if (proc_end == proc_begin + 1) {
    n_by_proc[proc_begin] = patch_end - patch_begin;
    return;
  }
  int proc_middle = (proc_begin + proc_end) / 2;
  double load_total = 0.;
  for (int p = patch_begin; p < patch_end; p++) {
    load_total += load_by_patch[p];
  }
  double load_target =
    (load_total * (proc_middle - proc_begin)) / (proc_end - proc_begin);
  double load = 0.;
  int patch_middle = patch_begin;
  for (;;) {
    double prev_load = load;
    load += load_by_patch[patch_middle];
    patch_middle++;
    if (load > load_target) {
      double above = load - load_target;
      double below = load_target - prev_load;
      if (below < above && patch_middle) {
        patch_middle--;
        load = prev_load;
      }
      break;
    }
    if (patch_middle >= patch_end) { /* the reference would read past the end here */
      break;
    }
  }
  if (patch_middle - patch_begin < proc_middle - proc_begin) {
    patch_middle = patch_begin + (proc_middle - proc_begin);
  }
  if (patch_end - patch_middle < proc_end - proc_middle) {
    patch_middle = patch_end - (proc_end - proc_middle);
  }
  po_map_rec(load_by_patch, n_by_proc, proc_begin, proc_middle, patch_begin, patch_middle);
  po_map_rec(load_by_patch, n_by_proc, proc_middle, proc_end, patch_middle, patch_end);
}

/* best_mapping (psc_balance_impl.hxx:153-160); capabilities are all equal in PSC's
 * non-capability build (this branch, :95-160) */
void po_best_mapping(int n_ranks, const double* capability, int n_patches,
                     const double* loads, int* n_patches_by_rank)
{
  (void)capability;
  po_map_rec(loads, n_patches_by_rank, 0, n_ranks, 0, n_patches);
}

/* get_loads (psc_balance_impl.hxx:223-269) */
void po_get_loads(const po_grid* g, const unsigned* off, double factor_fields,
                  double* loads)
{
  int n_cells = g->ldims[0] * g->ldims[1] * g->ldims[2];
  for (int p = 0; p < g->n_patches; p++) {
    loads[p] = (double)(off[p + 1] - off[p]) + factor_fields * n_cells;
  }
}

/* ---- BoundaryInjector::inject (src/include/boundary_injector.hxx:93-160) ----
 * The particle generator and get_n_in_cell draw from the host's sequential generators;
 * their output is handed in (cand: what generator.get(cell_corner, dx) returned for ghost
 * cell idx of patch `patch`, positions patch-local), and this restates what inject() does
 * with each draw, statement by statement, in the configuration's real_t (float; the
 * reference itself instantiates the template for its double configurations only, because
 * push_x binds a Vec3<real_t>& to the generator's Double3):
 *   :122     v = calc_v(u)                       (pushp.hxx:68-72)
 *   :123-124 initial_x = x; push_x(x, v)         (AdvanceParticle<real_t, dim_y>: y only, pushp.hxx:17-29)
 *   :126-129 a particle that does not enter the patch (x_y < 0) is not injected
 *   :133-135 x + xb goes to the injector, which stores real_t(x_glob) - real_t(xb) and
 *            qni_wni = real_t(w * q)             (injector_simple.hxx:27-34)
 *   :140-147 calc_j(J, initial_x * dxi, x * dxi, fint(..), initial_idx, q * w, v) with
 *            dxi = real_t(grid.domain.dx_inv)
 * out / out_patch (room for n) receive the accepted records in candidate order.
 * Returns the number accepted. */
long po_boundary_inject(const po_grid* g, float* flds, const po_inject_cand* cand, long n, po_prt* out,
                        int* out_patch)
{
  const int DIM_INJ = 1; /* INJECT_DIM_IDX_ */
  long plen = po_fld_patch_len(g) * PO_NR_FIELDS;
  float dt = (float)g->dt; /* AdvanceParticle(grid.dt) */
  float dxi[3];
  for (int d = 0; d < 3; d++) {
    dxi[d] = (float)g->dx_inv[d];
  }
  long n_out = 0;
  for (long i = 0; i < n; i++) {
    const po_inject_cand* c = &cand[i];
    float u[3] = {(float)c->u[0], (float)c->u[1], (float)c->u[2]};
    float root = rsqrt_host(1.f + sqrf(u[0]) + sqrf(u[1]) + sqrf(u[2]));
    float v[3] = {u[0] * root, u[1] * root, u[2] * root};
    float x0[3] = {(float)c->x[0], (float)c->x[1], (float)c->x[2]};
    float x[3] = {x0[0], x0[1], x0[2]};
    x[DIM_INJ] += 1.f * dt * v[DIM_INJ];
    if (x[DIM_INJ] < 0.f) {
      continue;
    }
    int poff[3];
    po_patch_off(g, c->patch, poff);
    po_prt* o = &out[n_out];
    for (int d = 0; d < 3; d++) {
      double xb = (double)poff[d] * g->dx[d] + g->corner[d]; /* grid.hxx:82-86 */
      o->x[d] = (float)((double)x[d] + xb) - (float)xb;
      o->u[d] = (float)c->u[d];
    }
    o->kind = c->kind;
    o->qni_wni = (float)(c->w * g->q[c->kind]);
    out_patch[n_out++] = c->patch;

    float xm[3], xp[3];
    int lf[3], lg[3] = {c->idx[0], c->idx[1], c->idx[2]};
    for (int d = 0; d < 3; d++) {
      xm[d] = x0[d] * dxi[d];
      xp[d] = x[d] * dxi[d];
      lf[d] = (int)floorf(xp[d]);
    }
    po_curr_f cur;
    po_curr_setup_f(&cur, g, flds + c->patch * plen);
    po_calc_j_impl_f(&cur, g->deposit, xm, xp, lf, lg, (float)(g->q[c->kind] * c->w), v);
  }
  return n_out;
}

const char* po_describe(void)
{
  return "plain-C restatement of psc-code/psc 1vb hot path (oracle/psc_oracle.c), "
         "gcc -O3 -ffp-contract=off, no -march";
}

#include "psc_oracle_collision.inc"
