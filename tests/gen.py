"""Synthetic inputs for the parity tests."""
import numpy as np

from oracle_lib import PRT_DTYPE, NR_FIELDS, EX, HX, off_from_counts


def random_fields(grid, seed=0, amp_e=0.05, amp_b=0.1):
    """smooth-ish random E/B in every point incl. ghosts (ghost consistency is
    not required by push_mprts, it just reads them)."""
    rng = np.random.default_rng(seed)
    f = grid.zeros_fields()
    f[:, EX:EX + 3] = amp_e * rng.standard_normal(f[:, EX:EX + 3].shape).astype(np.float32)
    f[:, HX:HX + 3] = amp_b * rng.standard_normal(f[:, HX:HX + 3].shape).astype(np.float32)
    return f


def thermal_plasma(grid, ppc, seed=0, vth=(0.05, 0.005), in_cell_uniform=True,
                   shuffle=True, margin=0.0):
    """ppc particles per cell per kind, positions uniform in the cell (or at cell
    centres like setup_particles.hxx:314-322), u ~ N(0, vth[kind]), w = 1."""
    rng = np.random.default_rng(seed)
    ld = grid.ldims
    dx = grid.dx
    nk = len(grid.kinds)
    n_cells = grid.n_cells
    n_per_patch = n_cells * ppc * nk
    prts = np.zeros(n_per_patch * grid.n_patches, dtype=PRT_DTYPE)
    cz, cy, cx = np.meshgrid(np.arange(ld[2]), np.arange(ld[1]), np.arange(ld[0]), indexing="ij")
    cell = np.stack([cx.ravel(), cy.ravel(), cz.ravel()], axis=1).astype(np.float64)
    for p in range(grid.n_patches):
        sl = slice(p * n_per_patch, (p + 1) * n_per_patch)
        c = np.repeat(cell, ppc * nk, axis=0)
        if in_cell_uniform:
            r = margin + (1 - 2 * margin) * rng.random(c.shape)
        else:
            r = np.full(c.shape, 0.5)
        x = (c + r) * np.array(dx)
        kind = np.tile(np.repeat(np.arange(nk), ppc), n_cells).astype(np.int32)
        v = np.array([vth[k % len(vth)] for k in range(nk)])[kind]
        u = rng.standard_normal(c.shape) * v[:, None]
        q = np.array([k[0] for k in grid.kinds])[kind]
        a = prts[sl]
        a["x"] = x.astype(np.float32)
        for d in range(3):
            if grid.g.invar[d]:
                a["x"][:, d] = np.float32(0.5 * dx[d]) if not in_cell_uniform else a["x"][:, d]
        a["u"] = u.astype(np.float32)
        a["kind"] = kind
        a["qni_wni"] = q.astype(np.float32)
        # keep strictly inside the patch in float arithmetic
        for d in range(3):
            hi = np.float32(ld[d] * dx[d])
            a["x"][:, d] = np.minimum(a["x"][:, d], np.nextafter(hi, np.float32(0)))
        if shuffle:
            prts[sl] = a[rng.permutation(n_per_patch)]
    off = off_from_counts([n_per_patch] * grid.n_patches)
    return prts, off


def prts_equal(a, b):
    return (a.tobytes() == b.tobytes())


def ulp_diff(a, b):
    """max distance in float32 ULPs between two float32 arrays"""
    a = np.ascontiguousarray(a, dtype=np.float32).view(np.int32).astype(np.int64)
    b = np.ascontiguousarray(b, dtype=np.float32).view(np.int32).astype(np.int64)
    a = np.where(a < 0, -(a & 0x7FFFFFFF), a)
    b = np.where(b < 0, -(b & 0x7FFFFFFF), b)
    return int(np.max(np.abs(a - b))) if a.size else 0
