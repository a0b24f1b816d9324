/*
 * cmx_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU (fp64) restatement of the per-frame minimum-distance hot path of
 * ComplexMixtures.jl (reference at /root/reference, v2.18.3-DEV).  Only tests/,
 * __graft_entry__.smoke() and the cpu_baseline / --impl reference legs of bench.py
 * may load this library; the product path (complexmixtures.jl_b200/csrc) never does.
 *
 * Parity status: the deterministic (real-distribution) half is PINNED by the
 * reference's own known-answer tests (tests/test_oracle_kat.py): frame-1 protein-TMAO
 * coordination numbers 7 / 14 / 1171 (src/tools/coordination_number.jl:126-134), the toy
 * systems of src/mddf.jl:587-758, the unit KATs of update_md / atom_type / eulermat /
 * move! / shellradius.  The random half is pinned only statistically (the reference
 * itself uses rtol 0.1, src/mddf.jl:833-839): Julia's StableRNG/Xoroshiro streams cannot
 * be regenerated here, so the oracle and the device share a Philox4x32-10 counter stream
 * instead ("parity unpinned beyond rtol 0.1" against real Julia numbers).
 *
 * The pair search of the reference lives in CellListMap.jl 0.10.1 (Project.toml:36, not
 * vendored).  Its documented behaviour is restated here: every (i in x, j in y) pair with
 * minimum-image distance d <= cutoff is visited once; the output list is reset before each
 * pairwise! call; the unit cell must be wider than 2*cutoff in every perpendicular
 * direction.  Ties on equal d are traversal-order dependent in the reference
 * (update_md keeps the later pair, src/minimum_distances.jl:33-37); this restatement fixes
 * the order-independent rule "smallest (d, j, i) wins" and the device does the same.
 *
 * All floating-point here is IEEE fp64 without contraction (compile with
 * -ffp-contract=off); the device's exact path uses the same operation order.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------ */
/* Public structs (mirrored by oracle/cmx_oracle.py with ctypes)                          */
/* ------------------------------------------------------------------------------------ */
typedef struct {
    int32_t nmols_solute, napm_solute;   /* AtomSelection.nmols / natomspermol, src/AtomSelection.jl:48-59 */
    int32_t nmols_solvent, napm_solvent;
    int32_t autocorrelation;             /* src/results.jl:240-243 */
    int32_t irefatom;                    /* 0-based here; reference is 1-based (src/Trajectory.jl:206-213) */
    int32_t usecutoff;                   /* src/Options.jl:6-42 */
    int32_t nbins;                       /* setbin(cutoff, binstep), src/results.jl:131 */
    int32_t n_random_samples;
    int32_t coordination_number_only;    /* src/mddf.jl:438-471 */
    int32_t ngroups_solute, ngroups_solvent;
    int32_t custom_solute, custom_solvent; /* custom_groups flags */
    double cutoff, dbulk, binstep;
    uint64_t seed;
    /* CSR "selection position -> groups containing that atom" (only if custom_*). */
    const int32_t *solute_grp_off, *solute_grp_ids;
    const int32_t *solvent_grp_off, *solvent_grp_ids;
} orc_config;

typedef struct {
    double *md_count, *md_count_random;              /* [nbins] */
    double *rdf_count, *rdf_count_random;            /* [nbins] */
    double *solute_group, *solute_group_random;      /* [ngroups_solute][nbins] row-major */
    double *solvent_group, *solvent_group_random;    /* [ngroups_solvent][nbins] */
    double volume_total;
    int64_t pair_evals;                              /* pairs with d <= cutoff visited (P_alg) */
} orc_counters;

/* MinimumDistance, src/minimum_distances.jl:12-21.  i is local to the current solute
 * molecule, j is global in the solvent selection; both 0-based here, -1 when empty. */
typedef struct {
    int32_t within_cutoff, i, j, ref_atom_within_cutoff;
    double d, d_ref_atom;
} orc_md;

/* ------------------------------------------------------------------------------------ */
/* Philox4x32-10 (Salmon et al., SC'11) -- shared counter convention with the device      */
/*   key = (seed_lo, seed_hi); ctr = (slot, sample, frame, block)                         */
/* ------------------------------------------------------------------------------------ */
void orc_philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static void draw(const orc_config *c, uint32_t slot, uint32_t sample, uint32_t frame, uint32_t block,
                 uint32_t out[4]) {
    uint32_t ctr[4] = {slot, sample, frame, block};
    uint32_t key[2] = {(uint32_t)(c->seed & 0xffffffffu), (uint32_t)(c->seed >> 32)};
    orc_philox4x32(ctr, key, out);
}
/* uniform in (0,1) from 32 bits; integer pick in [0,n) without floating point */
static double u01(uint32_t r) { return ((double)r + 0.5) * (1.0 / 4294967296.0); }
static uint32_t pick(uint32_t r, uint32_t n) { return (uint32_t)(((uint64_t)r * (uint64_t)n) >> 32); }

/* ------------------------------------------------------------------------------------ */
/* Geometry                                                                              */
/* ------------------------------------------------------------------------------------ */
typedef struct {
    double m[9];    /* column-major: m[0..2] = a, m[3..5] = b, m[6..8] = c (src/Trajectory.jl:72-81) */
    double inv[9];  /* column-major inverse */
    int ortho;
    double w[3];    /* perpendicular widths */
    double invl[3]; /* 1/L per axis (orthorhombic minimum image) */
    double volume;
} orc_cell;

void orc_cell_init(orc_cell *c, const double cell[9]) {
    memcpy(c->m, cell, sizeof(double) * 9);
    const double *a = cell, *b = cell + 3, *cc = cell + 6;
    double mind = fmin(fabs(a[0]), fmin(fabs(b[1]), fabs(cc[2])));
    double tol = 1e-10 * mind; /* convert_unitcell, src/Trajectory.jl:72-77 */
    c->ortho = fabs(a[1]) < tol && fabs(a[2]) < tol && fabs(b[0]) < tol && fabs(b[2]) < tol &&
               fabs(cc[0]) < tol && fabs(cc[1]) < tol;
    /* cross products */
    double bxc[3] = {b[1] * cc[2] - b[2] * cc[1], b[2] * cc[0] - b[0] * cc[2], b[0] * cc[1] - b[1] * cc[0]};
    double cxa[3] = {cc[1] * a[2] - cc[2] * a[1], cc[2] * a[0] - cc[0] * a[2], cc[0] * a[1] - cc[1] * a[0]};
    double axb[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
    /* cell_volume, src/mddf.jl:350-352: dot(cross(a, b), c) */
    double det = axb[0] * cc[0] + axb[1] * cc[1] + axb[2] * cc[2];
    c->volume = det;
    /* inverse: rows are bxc/det, cxa/det, axb/det; stored column-major */
    for (int k = 0; k < 3; ++k) {
        c->inv[0 + 3 * k] = bxc[k] / det;
        c->inv[1 + 3 * k] = cxa[k] / det;
        c->inv[2 + 3 * k] = axb[k] / det;
    }
    c->invl[0] = 1.0 / a[0]; c->invl[1] = 1.0 / b[1]; c->invl[2] = 1.0 / cc[2];
    c->w[0] = fabs(det) / sqrt(bxc[0] * bxc[0] + bxc[1] * bxc[1] + bxc[2] * bxc[2]);
    c->w[1] = fabs(det) / sqrt(cxa[0] * cxa[0] + cxa[1] * cxa[1] + cxa[2] * cxa[2]);
    c->w[2] = fabs(det) / sqrt(axb[0] * axb[0] + axb[1] * axb[1] + axb[2] * axb[2]);
}

/* Minimum-image of a difference vector.  Orthorhombic: per-axis rounding.  Triclinic:
 * rounding of the fractional components; this is the true minimum image whenever the
 * result is shorter than half the smallest perpendicular width, which holds for every
 * pair within the cutoff because CellListMap requires width > 2*cutoff. */
static void min_image(const orc_cell *c, double dr[3]) {
    if (c->ortho) {
        /* the image count is rint(dr * (1/L)): identical to rint(dr / L) except within one ulp of a
         * half-integer, i.e. for |dr| = L/2 >= cutoff, where both images are equally far */
        dr[0] = dr[0] - c->m[0] * rint(dr[0] * c->invl[0]);
        dr[1] = dr[1] - c->m[4] * rint(dr[1] * c->invl[1]);
        dr[2] = dr[2] - c->m[8] * rint(dr[2] * c->invl[2]);
    } else {
        const double *v = c->inv, *m = c->m;
        double s0 = (v[0] * dr[0] + v[3] * dr[1]) + v[6] * dr[2];
        double s1 = (v[1] * dr[0] + v[4] * dr[1]) + v[7] * dr[2];
        double s2 = (v[2] * dr[0] + v[5] * dr[1]) + v[8] * dr[2];
        s0 = s0 - rint(s0); s1 = s1 - rint(s1); s2 = s2 - rint(s2);
        dr[0] = (m[0] * s0 + m[3] * s1) + m[6] * s2;
        dr[1] = (m[1] * s0 + m[4] * s1) + m[7] * s2;
        dr[2] = (m[2] * s0 + m[5] * s1) + m[8] * s2;
    }
}

static double dist_pbc(const orc_cell *c, const double *xi, const double *xj) {
    double dr[3] = {xj[0] - xi[0], xj[1] - xi[1], xj[2] - xi[2]};
    min_image(c, dr);
    return sqrt((dr[0] * dr[0] + dr[1] * dr[1]) + dr[2] * dr[2]);
}

double orc_dist_pbc(const double cell[9], const double *xi, const double *xj) {
    orc_cell c; orc_cell_init(&c, cell);
    return dist_pbc(&c, xi, xj);
}

/* setbin, src/results.jl:28 (1-based there; 0-based bin returned here) */
static int setbin0(double d, double step) {
    int ib = (int)ceil(d / step);
    if (ib < 1) ib = 1;
    return ib - 1;
}
int orc_setbin(double d, double step) { return setbin0(d, step) + 1; }

/* eulermat, src/rigid_body.jl:45-57 (row-major 3x3 out) */
void orc_eulermat(double beta, double gamma, double theta, double A[9]) {
    double c1 = cos(beta), s1 = sin(beta), c2 = cos(gamma), s2 = sin(gamma), c3 = cos(theta), s3 = sin(theta);
    A[0] = c2 * c3;                 A[1] = -c2 * s3;                A[2] = s2;
    A[3] = c1 * s3 + c3 * s1 * s2;  A[4] = c1 * c3 - s1 * s2 * s3;  A[5] = -c2 * s1;
    A[6] = s1 * s3 - c1 * c3 * s2;  A[7] = c1 * s2 * s3 + c3 * s1;  A[8] = c1 * c2;
}

/* move!, src/rigid_body.jl:73-80 */
void orc_move(double *x, int n, const double newcm[3], double beta, double gamma, double theta) {
    double cm[3] = {0, 0, 0};
    for (int k = 0; k < n; ++k) { cm[0] += x[3 * k]; cm[1] += x[3 * k + 1]; cm[2] += x[3 * k + 2]; }
    cm[0] /= n; cm[1] /= n; cm[2] /= n;
    double A[9]; orc_eulermat(beta, gamma, theta, A);
    for (int k = 0; k < n; ++k) {
        double p0 = x[3 * k] - cm[0], p1 = x[3 * k + 1] - cm[1], p2 = x[3 * k + 2] - cm[2];
        x[3 * k]     = ((A[0] * p0 + A[1] * p1) + A[2] * p2) + newcm[0];
        x[3 * k + 1] = ((A[3] * p0 + A[4] * p1) + A[5] * p2) + newcm[1];
        x[3 * k + 2] = ((A[6] * p0 + A[7] * p1) + A[8] * p2) + newcm[2];
    }
}

/* random_move!, src/rigid_body.jl:107-137.  The reference draws the new centre in a box
 * 10^4 times the computing box and lets CellListMap wrap it, i.e. uniformly over the
 * periodic cell; here the centre is M*(u1,u2,u3) with u uniform in (0,1)^3 (same
 * distribution, never forming 10^6-Angstrom coordinates).  Angles: three independent
 * uniform Euler angles as in the reference (:123-125). */
static void random_move(const orc_config *cfg, const orc_cell *c, double *x, int n, int iref,
                        const uint32_t r0[4], const uint32_t r1[4]) {
    (void)cfg;
    double u0 = u01(r0[1]), u1 = u01(r0[2]), u2 = u01(r0[3]);
    const double *m = c->m;
    double newcm[3] = {(m[0] * u0 + m[3] * u1) + m[6] * u2, (m[1] * u0 + m[4] * u1) + m[7] * u2,
                       (m[2] * u0 + m[5] * u1) + m[8] * u2};
    const double twopi = 6.283185307179586476925286766559;
    double beta = twopi * u01(r1[0]), gamma = twopi * u01(r1[1]), theta = twopi * u01(r1[2]);
    /* wrap_relative_to(x[iat], x[irefatom], uc), :129-131 */
    double ref[3] = {x[3 * iref], x[3 * iref + 1], x[3 * iref + 2]};
    for (int k = 0; k < n; ++k) {
        double dr[3] = {x[3 * k] - ref[0], x[3 * k + 1] - ref[1], x[3 * k + 2] - ref[2]};
        min_image(c, dr);
        x[3 * k] = ref[0] + dr[0]; x[3 * k + 1] = ref[1] + dr[1]; x[3 * k + 2] = ref[2] + dr[2];
    }
    orc_move(x, n, newcm, beta, gamma, theta);
}

/* exported single-molecule version for unit tests (rigidity, src/rigid_body.jl:139-190) */
void orc_random_move(const double cell[9], double *x, int n, int iref, uint64_t seed, uint32_t slot,
                     uint32_t sample, uint32_t frame) {
    orc_config cfg; memset(&cfg, 0, sizeof cfg); cfg.seed = seed;
    orc_cell c; orc_cell_init(&c, cell);
    uint32_t r0[4], r1[4];
    draw(&cfg, slot, sample, frame, 0, r0); draw(&cfg, slot, sample, frame, 1, r1);
    random_move(&cfg, &c, x, n, iref, r0, r1);
}

/* ------------------------------------------------------------------------------------ */
/* update_md / update_list!, src/minimum_distances.jl:30-39, 81-120                       */
/* ------------------------------------------------------------------------------------ */
static const orc_md MD_ZERO = {0, -1, -1, 0, INFINITY, INFINITY};

static void update_list(const orc_config *cfg, orc_md *list, int i, int j, double d, int isolute) {
    int jmol = j / cfg->napm_solvent;                 /* mol_index, :72 */
    if (cfg->autocorrelation && jmol == isolute) return; /* :90 */
    orc_md *m = &list[jmol];
    int is_ref = (j % cfg->napm_solvent) == cfg->irefatom; /* atom_type(j) == jref_atom */
    if (is_ref) {
        m->ref_atom_within_cutoff = 1;
        if (d < m->d_ref_atom) m->d_ref_atom = d;
    }
    /* smallest (d, j, i) wins */
    if (d < m->d || (d == m->d && (j < m->j || (j == m->j && i < m->i)))) {
        m->within_cutoff = 1; m->i = i; m->j = j; m->d = d;
    }
}

/* exported for the update_md KAT (src/minimum_distances.jl:41-46) */
void orc_update_md(const orc_md *a, const orc_md *b, orc_md *out) {
    int refw = a->ref_atom_within_cutoff || b->ref_atom_within_cutoff;
    double dref = refw ? fmin(a->d_ref_atom, b->d_ref_atom) : INFINITY;
    const orc_md *w = (a->d < b->d) ? a : b;
    *out = *w; out->ref_atom_within_cutoff = refw; out->d_ref_atom = dref;
}

/* ------------------------------------------------------------------------------------ */
/* minimum_distances!, src/minimum_distances.jl:129-148                                   */
/*   x = one solute molecule, y = all solvent atoms, list[nmols_solvent] reset first.     */
/* ------------------------------------------------------------------------------------ */
static double eff_cutoff(const orc_config *cfg) { return cfg->usecutoff ? cfg->cutoff : cfg->dbulk; } /* :168 */

static int64_t md_brute(const orc_config *cfg, const orc_cell *c, const double *x, const double *y,
                        int isolute, orc_md *list) {
    int ny = cfg->nmols_solvent * cfg->napm_solvent;
    double cut = eff_cutoff(cfg);
    int64_t npairs = 0;
    for (int m = 0; m < cfg->nmols_solvent; ++m) list[m] = MD_ZERO;
    for (int j = 0; j < ny; ++j)
        for (int i = 0; i < cfg->napm_solute; ++i) {
            double d = dist_pbc(c, x + 3 * i, y + 3 * j);
            if (d <= cut) { update_list(cfg, list, i, j, d, isolute); ++npairs; }
        }
    return npairs;
}

/* Cell list over the solvent atoms (fractional space, cells at least `cut` wide), the
 * CPU-side stand-in for CellListMap's structure; rebuilt when y changes. */
typedef struct {
    int n[3], ncells;
    int *start;    /* [ncells+1] */
    int *atoms;    /* [ny] atom indices sorted by cell */
    int *cellof;   /* scratch [ny] */
    double *wpos;  /* [3*ny] positions wrapped into the primary cell, in cell-sorted order (pre-test only) */
    int ny;
} orc_clist;

static void clist_alloc(orc_clist *cl, int ny) {
    memset(cl, 0, sizeof *cl);
    cl->ny = ny;
    cl->atoms = (int *)malloc(sizeof(int) * (size_t)(ny > 0 ? ny : 1));
    cl->cellof = (int *)malloc(sizeof(int) * (size_t)(ny > 0 ? ny : 1));
    cl->wpos = (double *)malloc(sizeof(double) * 3 * (size_t)(ny > 0 ? ny : 1));
}
static void clist_free(orc_clist *cl) { free(cl->start); free(cl->atoms); free(cl->cellof); free(cl->wpos); }

static void cart_of_frac(const orc_cell *c, const double s[3], double r[3]) {
    const double *m = c->m;
    r[0] = (m[0] * s[0] + m[3] * s[1]) + m[6] * s[2];
    r[1] = (m[1] * s[0] + m[4] * s[1]) + m[7] * s[2];
    r[2] = (m[2] * s[0] + m[5] * s[1]) + m[8] * s[2];
}

static void frac_of(const orc_cell *c, const double *r, double s[3]) {
    if (c->ortho) { s[0] = r[0] / c->m[0]; s[1] = r[1] / c->m[4]; s[2] = r[2] / c->m[8]; }
    else {
        const double *v = c->inv;
        s[0] = (v[0] * r[0] + v[3] * r[1]) + v[6] * r[2];
        s[1] = (v[1] * r[0] + v[4] * r[1]) + v[7] * r[2];
        s[2] = (v[2] * r[0] + v[5] * r[1]) + v[8] * r[2];
    }
    for (int k = 0; k < 3; ++k) { s[k] -= floor(s[k]); if (s[k] >= 1.0) s[k] = 0.0; }
}

/* cells are cut/ORC_LCELL wide (CellListMap's `lcell`, src/minimum_distances.jl:170): with 2 the 5x5x5 block of
 * neighbour cells holds ~58 % of the candidates of the 3x3x3 block of cut-wide cells */
#define ORC_LCELL 2
static void clist_build(orc_clist *cl, const orc_cell *c, const double *y, double cut) {
    int n[3];
    for (int k = 0; k < 3; ++k) { n[k] = (int)floor(c->w[k] * ORC_LCELL / cut); if (n[k] < 1) n[k] = 1; if (n[k] > 256) n[k] = 256; }
    int ncells = n[0] * n[1] * n[2];
    if (ncells != cl->ncells) { free(cl->start); cl->start = (int *)malloc(sizeof(int) * (size_t)(ncells + 1)); }
    cl->n[0] = n[0]; cl->n[1] = n[1]; cl->n[2] = n[2]; cl->ncells = ncells;
    memset(cl->start, 0, sizeof(int) * (size_t)(ncells + 1));
    for (int j = 0; j < cl->ny; ++j) {
        double s[3]; frac_of(c, y + 3 * j, s);
        int cx = (int)(s[0] * n[0]), cy = (int)(s[1] * n[1]), cz = (int)(s[2] * n[2]);
        if (cx >= n[0]) cx = n[0] - 1; if (cy >= n[1]) cy = n[1] - 1; if (cz >= n[2]) cz = n[2] - 1;
        int id = (cz * n[1] + cy) * n[0] + cx;
        cl->cellof[j] = id; cl->start[id + 1]++;
    }
    for (int k = 0; k < ncells; ++k) cl->start[k + 1] += cl->start[k];
    int *fill = (int *)calloc((size_t)ncells, sizeof(int));
    for (int j = 0; j < cl->ny; ++j) {
        int id = cl->cellof[j], slot = cl->start[id] + fill[id]++;
        cl->atoms[slot] = j;
        double sj[3]; frac_of(c, y + 3 * j, sj);
        cart_of_frac(c, sj, cl->wpos + 3 * (size_t)slot);
    }
    free(fill);
}

/* distinct cells within ORC_LCELL cells of cell `c0` along an axis with n cells (periodic) */
static int nb_cells(int n, int c0, int out[2 * ORC_LCELL + 1]) {
    if (n <= 2 * ORC_LCELL + 1) { for (int k = 0; k < n; ++k) out[k] = k; return n; }
    for (int d = -ORC_LCELL; d <= ORC_LCELL; ++d) out[d + ORC_LCELL] = (c0 + d + n) % n;
    return 2 * ORC_LCELL + 1;
}

/* Pair search of one solute molecule.  Candidates come from the (2*ORC_LCELL+1)^3 neighbour cells; each is first
 * tested on wrapped coordinates with the periodic shift of its cell (three subtractions, no rounding calls), and
 * only the pairs that pass (with a relative margin far above the rounding of that test) are evaluated with
 * dist_pbc on the ORIGINAL coordinates -- the arithmetic every counted distance goes through.  A pair within
 * the cutoff always lies, at its minimum image, in one of those cells (cell width >= cut/ORC_LCELL in every
 * perpendicular direction), so the pre-test never loses one. */
static int64_t md_clist(const orc_config *cfg, const orc_cell *c, const orc_clist *cl, const double *x,
                        const double *y, int isolute, orc_md *list) {
    const double cut = eff_cutoff(cfg);
    const double cut2m = cut * cut * (1.0 + 1e-9);
    int64_t npairs = 0;
    for (int m = 0; m < cfg->nmols_solvent; ++m) list[m] = MD_ZERO;
    const int *n = cl->n;
    const int L = ORC_LCELL;
    const int shifted = n[0] > 2 * L + 1 && n[1] > 2 * L + 1 && n[2] > 2 * L + 1;
    for (int i = 0; i < cfg->napm_solute; ++i) {
        const double *xi = x + 3 * i;
        double s[3]; frac_of(c, xi, s);
        int cx = (int)(s[0] * n[0]), cy = (int)(s[1] * n[1]), cz = (int)(s[2] * n[2]);
        if (cx >= n[0]) cx = n[0] - 1; if (cy >= n[1]) cy = n[1] - 1; if (cz >= n[2]) cz = n[2] - 1;
        if (!shifted) {   /* small grids: every cell along a short axis, exact arithmetic for every candidate */
            int ex[2 * ORC_LCELL + 1], ey[2 * ORC_LCELL + 1], ez[2 * ORC_LCELL + 1];
            const int nx = nb_cells(n[0], cx, ex), ny = nb_cells(n[1], cy, ey), nz = nb_cells(n[2], cz, ez);
            for (int kz = 0; kz < nz; ++kz)
                for (int ky = 0; ky < ny; ++ky)
                    for (int kx = 0; kx < nx; ++kx) {
                        int id = (ez[kz] * n[1] + ey[ky]) * n[0] + ex[kx];
                        for (int p = cl->start[id]; p < cl->start[id + 1]; ++p) {
                            int j = cl->atoms[p];
                            double d = dist_pbc(c, xi, y + 3 * j);
                            if (d <= cut) { update_list(cfg, list, i, j, d, isolute); ++npairs; }
                        }
                    }
            continue;
        }
        double xw[3]; cart_of_frac(c, s, xw);
        for (int dz = -L; dz <= L; ++dz)
            for (int dy = -L; dy <= L; ++dy)
                for (int dx = -L; dx <= L; ++dx) {
                    int rx = cx + dx, ry = cy + dy, rz = cz + dz;
                    int kx = rx < 0 ? -1 : (rx >= n[0] ? 1 : 0), ky = ry < 0 ? -1 : (ry >= n[1] ? 1 : 0), kz = rz < 0 ? -1 : (rz >= n[2] ? 1 : 0);
                    int id = ((rz - kz * n[2]) * n[1] + (ry - ky * n[1])) * n[0] + (rx - kx * n[0]);
                    /* the neighbour cell's atoms sit at wpos + kx*a + ky*b + kz*c: move the solute atom the other way */
                    const double *m = c->m;
                    const double px = xw[0] - (kx * m[0] + ky * m[3] + kz * m[6]);
                    const double py = xw[1] - (kx * m[1] + ky * m[4] + kz * m[7]);
                    const double pz = xw[2] - (kx * m[2] + ky * m[5] + kz * m[8]);
                    const double *w = cl->wpos + 3 * (size_t)cl->start[id];
                    const int cnt = cl->start[id + 1] - cl->start[id];
                    for (int q = 0; q < cnt; ++q) {
                        double ax = w[3 * q] - px, ay = w[3 * q + 1] - py, az = w[3 * q + 2] - pz;
                        if (ax * ax + ay * ay + az * az > cut2m) continue;
                        int j = cl->atoms[cl->start[id] + q];
                        double d = dist_pbc(c, xi, y + 3 * j);
                        if (d <= cut) { update_list(cfg, list, i, j, d, isolute); ++npairs; }
                    }
                }
    }
    return npairs;
}

/* exported: list of one solute molecule vs. a solvent configuration */
int64_t orc_minimum_distances(const orc_config *cfg, const double cell[9], const double *x, const double *y,
                              int isolute, int use_clist, orc_md *list) {
    orc_cell c; orc_cell_init(&c, cell);
    if (!use_clist) return md_brute(cfg, &c, x, y, isolute, list);
    orc_clist cl; clist_alloc(&cl, cfg->nmols_solvent * cfg->napm_solvent);
    clist_build(&cl, &c, y, eff_cutoff(cfg));
    int64_t np = md_clist(cfg, &c, &cl, x, y, isolute, list);
    clist_free(&cl);
    return np;
}

/* ------------------------------------------------------------------------------------ */
/* update_counters!, src/update_counters.jl:22-88                                         */
/* ------------------------------------------------------------------------------------ */
static void group_add(double *arr, int nbins, int ibin, int pos, int napm, int custom, const int32_t *off,
                      const int32_t *ids, double w) {
    if (!custom) arr[(size_t)(pos % napm) * nbins + ibin] += w;  /* atom_type, :9,24-25 */
    else for (int q = off[pos]; q < off[pos + 1]; ++q) arr[(size_t)ids[q] * nbins + ibin] += w; /* :27-33 */
}

static void update_counters(const orc_config *cfg, const orc_md *list, double w, int random, orc_counters *R) {
    int nb = cfg->nbins;
    double *md = random ? R->md_count_random : R->md_count;
    double *rdf = random ? R->rdf_count_random : R->rdf_count;
    double *gsol = random ? R->solute_group_random : R->solute_group;
    double *gsolv = random ? R->solvent_group_random : R->solvent_group;
    for (int m = 0; m < cfg->nmols_solvent; ++m) {
        const orc_md *e = &list[m];
        if (!e->within_cutoff) continue;
        int ib = setbin0(e->d, cfg->binstep);
        if (ib >= nb) ib = nb - 1; /* guard (never taken when d <= cutoff = nbins*binstep) */
        md[ib] += w;
        if (cfg->autocorrelation) {
            /* both atoms into the *solute* array at w/2; md.i is local to the current
             * molecule, md.j is global (:48-53) */
            group_add(gsol, nb, ib, e->i, cfg->napm_solute, cfg->custom_solute, cfg->solute_grp_off, cfg->solute_grp_ids, w / 2);
            group_add(gsol, nb, ib, e->j, cfg->napm_solute, cfg->custom_solute, cfg->solute_grp_off, cfg->solute_grp_ids, w / 2);
        } else {
            group_add(gsol, nb, ib, e->i, cfg->napm_solute, cfg->custom_solute, cfg->solute_grp_off, cfg->solute_grp_ids, w);
            group_add(gsolv, nb, ib, e->j, cfg->napm_solvent, cfg->custom_solvent, cfg->solvent_grp_off, cfg->solvent_grp_ids, w);
        }
        if (e->ref_atom_within_cutoff) {
            int ir = setbin0(e->d_ref_atom, cfg->binstep);
            if (ir >= nb) ir = nb - 1;
            rdf[ir] += w;
        }
    }
}

/* inbulk, src/mddf.jl:55-57 */
static int inbulk(const orc_config *cfg, const orc_md *e) {
    return cfg->usecutoff ? (e->within_cutoff && e->d > cfg->dbulk) : !e->within_cutoff;
}

/* which solute molecule is the reference of random sample s (src/mddf.jl:374-376) */
static int ref_solute_of_sample(const orc_config *cfg, uint32_t frame, uint32_t s) {
    uint32_t r[4]; draw(cfg, 0xffffffffu, s, frame, 2, r);
    return (int)pick(r[0], (uint32_t)cfg->nmols_solute);
}
int orc_ref_solute(const orc_config *cfg, uint32_t frame, uint32_t s) { return ref_solute_of_sample(cfg, frame, s); }

/* randomize_solvent!, src/mddf.jl:65-88 */
static void randomize_solvent(const orc_config *cfg, const orc_cell *c, const double *yread, double *y,
                              const int *bulk, int nbulk, uint32_t frame, uint32_t sample) {
    int na = cfg->napm_solvent;
    for (int m = 0; m < cfg->nmols_solvent; ++m) {
        uint32_t r0[4], r1[4];
        draw(cfg, (uint32_t)m, sample, frame, 0, r0);
        draw(cfg, (uint32_t)m, sample, frame, 1, r1);
        int jmol = nbulk > 0 ? bulk[pick(r0[0], (uint32_t)nbulk)] : (int)pick(r0[0], (uint32_t)cfg->nmols_solvent);
        memcpy(y + (size_t)3 * na * m, yread + (size_t)3 * na * jmol, sizeof(double) * 3 * (size_t)na);
        random_move(cfg, c, y + (size_t)3 * na * m, na, cfg->irefatom, r0, r1);
    }
}

void orc_randomize_solvent(const orc_config *cfg, const double cell[9], const double *yread, double *y,
                           const int *bulk, int nbulk, uint32_t frame, uint32_t sample) {
    orc_cell c; orc_cell_init(&c, cell);
    randomize_solvent(cfg, &c, yread, y, bulk, nbulk, frame, sample);
}

/* ------------------------------------------------------------------------------------ */
/* mddf_frame! / coordination_number_frame!, src/mddf.jl:361-471                          */
/*   list_out (optional): [nmols_solute][nmols_solvent] real-phase lists                  */
/*   rand_list_out (optional): [n_random_samples][nmols_solvent] random-phase lists       */
/* ------------------------------------------------------------------------------------ */
typedef struct { orc_md *list, *save; int *bulk; double *yrand; orc_clist cl; } orc_scratch;

static void scratch_alloc(const orc_config *cfg, orc_scratch *s) {
    int nm = cfg->nmols_solvent, ny = nm * cfg->napm_solvent;
    s->list = (orc_md *)malloc(sizeof(orc_md) * (size_t)(nm > 0 ? nm : 1));
    s->save = (orc_md *)malloc(sizeof(orc_md) * (size_t)(nm > 0 ? nm : 1));
    s->bulk = (int *)malloc(sizeof(int) * (size_t)(nm > 0 ? nm : 1));
    s->yrand = (double *)malloc(sizeof(double) * 3 * (size_t)(ny > 0 ? ny : 1));
    clist_alloc(&s->cl, ny);
}
static void scratch_free(orc_scratch *s) { free(s->list); free(s->save); free(s->bulk); free(s->yrand); clist_free(&s->cl); }

static void frame_impl(const orc_config *cfg, const double *xsolute, const double *xsolvent, const double cell[9],
                       double w, uint32_t frame, int use_clist, orc_counters *R, orc_scratch *S,
                       orc_md *list_out, orc_md *rand_list_out) {
    orc_cell c; orc_cell_init(&c, cell);
    R->volume_total += w * c.volume;  /* update_volume!, :351 */
    int nm = cfg->nmols_solvent, nrs = cfg->coordination_number_only ? 0 : cfg->n_random_samples;
    double cut = eff_cutoff(cfg);
    if (use_clist) clist_build(&S->cl, &c, xsolvent, cut);
    for (int isolute = 0; isolute < cfg->nmols_solute; ++isolute) {
        const double *x = xsolute + (size_t)3 * cfg->napm_solute * isolute;  /* viewmol, :385 */
        R->pair_evals += use_clist ? md_clist(cfg, &c, &S->cl, x, xsolvent, isolute, S->list)
                                   : md_brute(cfg, &c, x, xsolvent, isolute, S->list);
        if (list_out) memcpy(list_out + (size_t)isolute * nm, S->list, sizeof(orc_md) * (size_t)nm);
        update_counters(cfg, S->list, w, 0, R);
        int nrand = 0;
        for (int s = 0; s < nrs; ++s) nrand += (ref_solute_of_sample(cfg, frame, (uint32_t)s) == isolute);
        if (nrand == 0) continue;
        int nbulk = 0;
        for (int m = 0; m < nm; ++m) {
            if (cfg->autocorrelation && m == isolute) continue;  /* :409 */
            if (inbulk(cfg, &S->list[m])) S->bulk[nbulk++] = m;
        }
        for (int s = 0; s < nrs; ++s) {
            if (ref_solute_of_sample(cfg, frame, (uint32_t)s) != isolute) continue;
            randomize_solvent(cfg, &c, xsolvent, S->yrand, S->bulk, nbulk, frame, (uint32_t)s);
            if (use_clist) {
                orc_clist cl2; clist_alloc(&cl2, nm * cfg->napm_solvent);
                clist_build(&cl2, &c, S->yrand, cut);
                R->pair_evals += md_clist(cfg, &c, &cl2, x, S->yrand, isolute, S->save);
                clist_free(&cl2);
            } else R->pair_evals += md_brute(cfg, &c, x, S->yrand, isolute, S->save);
            if (rand_list_out) memcpy(rand_list_out + (size_t)s * nm, S->save, sizeof(orc_md) * (size_t)nm);
            update_counters(cfg, S->save, w, 1, R);
        }
    }
}

void orc_frame(const orc_config *cfg, const double *xsolute, const double *xsolvent, const double cell[9],
               double w, uint32_t frame, int use_clist, orc_counters *R, orc_md *list_out, orc_md *rand_list_out) {
    orc_scratch S; scratch_alloc(cfg, &S);
    frame_impl(cfg, xsolute, xsolvent, cell, w, frame, use_clist, R, &S, list_out, rand_list_out);
    scratch_free(&S);
}

/* ------------------------------------------------------------------------------------ */
/* Frame-parallel driver = the reference's chunk loop (src/mddf.jl:285-338,                */
/* src/parallel_setup.jl:26): one private counter set per thread, frames distributed over  */
/* threads, final sum! (src/results.jl:629-649).  Input coordinates are fp32 (as stored in */
/* DCD/XTC files, NamdDCD.jl:41-43) and promoted to fp64 per frame.                        */
/*   xs: [nframes][ns][3] float, xv: [nframes][nv][3] float (xv == xs rows if autocorr)    */
/*   cells: [nframes][9], weights: [nframes] (NULL = 1.0), frame_ids: [nframes] (NULL=0..) */
/* ------------------------------------------------------------------------------------ */
static size_t counters_len(const orc_config *cfg) {
    return (size_t)cfg->nbins * (4 + 2 * (size_t)cfg->ngroups_solute + 2 * (size_t)cfg->ngroups_solvent);
}
static void counters_bind(const orc_config *cfg, double *buf, orc_counters *R) {
    size_t nb = (size_t)cfg->nbins;
    R->md_count = buf; buf += nb; R->md_count_random = buf; buf += nb;
    R->rdf_count = buf; buf += nb; R->rdf_count_random = buf; buf += nb;
    R->solute_group = buf; buf += nb * cfg->ngroups_solute; R->solute_group_random = buf; buf += nb * cfg->ngroups_solute;
    R->solvent_group = buf; buf += nb * cfg->ngroups_solvent; R->solvent_group_random = buf;
    R->volume_total = 0; R->pair_evals = 0;
}

int orc_run_frames(const orc_config *cfg, int nframes, const float *xs, const float *xv, const double *cells,
                   const double *weights, const uint32_t *frame_ids, int use_clist, int nthreads, orc_counters *out) {
    size_t ns = (size_t)cfg->nmols_solute * cfg->napm_solute, nv = (size_t)cfg->nmols_solvent * cfg->napm_solvent;
    size_t len = counters_len(cfg);
    if (nthreads < 1) nthreads = 1;
#ifdef _OPENMP
    omp_set_num_threads(nthreads);
#else
    nthreads = 1;
#endif
    double *priv = (double *)calloc(len * (size_t)nthreads, sizeof(double));
    double *vol = (double *)calloc((size_t)nthreads, sizeof(double));
    int64_t *pe = (int64_t *)calloc((size_t)nthreads, sizeof(int64_t));
#pragma omp parallel
    {
#ifdef _OPENMP
        int t = omp_get_thread_num();
#else
        int t = 0;
#endif
        orc_counters R; counters_bind(cfg, priv + len * (size_t)t, &R);
        orc_scratch S; scratch_alloc(cfg, &S);
        double *x = (double *)malloc(sizeof(double) * 3 * (ns > 0 ? ns : 1));
        double *y = cfg->autocorrelation ? x : (double *)malloc(sizeof(double) * 3 * (nv > 0 ? nv : 1));
#pragma omp for schedule(dynamic, 1)
        for (int f = 0; f < nframes; ++f) {
            double w = weights ? weights[f] : 1.0;
            if (w == 0.0) continue; /* zero-weight frames are skipped, src/mddf.jl:102 */
            for (size_t k = 0; k < 3 * ns; ++k) x[k] = (double)xs[(size_t)f * 3 * ns + k];
            if (!cfg->autocorrelation) for (size_t k = 0; k < 3 * nv; ++k) y[k] = (double)xv[(size_t)f * 3 * nv + k];
            frame_impl(cfg, x, y, cells + (size_t)9 * f, w, frame_ids ? frame_ids[f] : (uint32_t)f, use_clist, &R, &S, NULL, NULL);
        }
        vol[t] = R.volume_total; pe[t] = R.pair_evals;
        free(x); if (!cfg->autocorrelation) free(y);
        scratch_free(&S);
    }
    /* sum!, src/results.jl:629-649 */
    orc_counters tmp; counters_bind(cfg, priv, &tmp);
    for (int t = 0; t < nthreads; ++t) {
        const double *p = priv + len * (size_t)t;
        size_t nb = (size_t)cfg->nbins, o = 0;
        for (size_t k = 0; k < nb; ++k) out->md_count[k] += p[o + k]; o += nb;
        for (size_t k = 0; k < nb; ++k) out->md_count_random[k] += p[o + k]; o += nb;
        for (size_t k = 0; k < nb; ++k) out->rdf_count[k] += p[o + k]; o += nb;
        for (size_t k = 0; k < nb; ++k) out->rdf_count_random[k] += p[o + k]; o += nb;
        size_t gs = nb * cfg->ngroups_solute, gv = nb * cfg->ngroups_solvent;
        for (size_t k = 0; k < gs; ++k) out->solute_group[k] += p[o + k]; o += gs;
        for (size_t k = 0; k < gs; ++k) out->solute_group_random[k] += p[o + k]; o += gs;
        for (size_t k = 0; k < gv; ++k) out->solvent_group[k] += p[o + k]; o += gv;
        for (size_t k = 0; k < gv; ++k) out->solvent_group_random[k] += p[o + k];
        out->volume_total += vol[t]; out->pair_evals += pe[t];
    }
    free(priv); free(vol); free(pe);
    return 0;
}
