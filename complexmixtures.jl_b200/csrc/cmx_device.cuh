// cmx_device.cuh -- device-side building blocks of libcmx_b200 (sm_100a).
//
// Exact (fp64, uncontracted) geometry used to finalise every counted distance, the Philox
// counter RNG, the rigid-body random placement and the histogram update.  Reference
// semantics restated: src/minimum_distances.jl:12-39,72-120 (MinimumDistance, update_md,
// update_list!), src/update_counters.jl:9-88, src/rigid_body.jl:45-57,73-80,107-137,
// src/results.jl:28 (setbin), src/mddf.jl:55-57 (inbulk).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace cmx {

typedef unsigned long long u64;

// ---- per-frame geometry (passed to kernels by value) --------------------------------------
struct Geom {
    double m[9], inv[9];      // unit cell, column-major (columns = lattice vectors) and inverse
    double invl[3];           // 1/L per axis (orthorhombic minimum image)
    double ctr[3];            // centre of the search grid; fp32 positions are stored relative to it
    double elo[3], ehi[3];    // extended AABB = AABB(primary cell) +- margin
    int ortho;
    // fine grid over the extended AABB (solute atoms + their periodic images)
    float gmin[3];            // elo - ctr
    float side, inv_side;     // cell size along y and z (a "row" is one (y,z) column of cells)
    float sidex, inv_sidex;   // cell size along x (finer: rows are scanned over x-spans)
    int nx, ny, nz;
    int rw;                   // 64-bit words per fine-grid row of the occupancy bitmask
    // query cells (cubic): tiles of 32 query atoms are consecutive runs in this order
    float qside, inv_qside;
    int nqx, nqy, nqz;
    // cull grid (same box, isotropic): occupancy bitmap -> lower bound of the distance to the solute
    float cside, inv_cside;
    int ncx, ncy, ncz, cw;    // cw = 64-bit words per cull-grid x-row
    int dwin;                 // window (cells) of the distance transform
    float rmax_bound;         // molecule radius bound the distance transform window was sized for
    float cut_hi2;            // (cut + tau)^2
    // cutoffs (effective cutoff = usecutoff ? cutoff : dbulk, src/minimum_distances.jl:168)
    float cut, tau;           // tau: fp32 distance uncertainty used to flag near-ties / edge cases
    float cut_lo, cut_hi;     // cut - tau, cut + tau
    float search2;            // (cut + tau)^2, initial search bound
    float ring;               // width of the distance rings in which the search consumes the solute-grid rows
    float tol_d2;             // 2*cut*tau + tau^2: d2 window that still may hide a near-tie
    double cutd;              // effective cutoff (fp64)
};

// ---- static problem description -----------------------------------------------------------
struct Prob {
    int ns_mols, ns_apm, nv_mols, nv_apm;
    int autocorr, iref, usecutoff, nbins, nrand, cn_only;
    int ng_sol, ng_solv, custom_sol, custom_solv;
    double cutoff, dbulk, binstep;
    uint32_t seed_lo, seed_hi;
    const int *sol_off, *sol_ids, *solv_off, *solv_ids;   // CSR position -> groups (device)
    u64 *md, *md_r, *rdf, *rdf_r, *gsol, *gsol_r, *gsolv, *gsolv_r;   // run accumulators (integer: hits at the first weight seen)
    const u64 *cnt_base;      // start of the contiguous accumulator block (= md)
    double *acc;              // fp64 twin of the block, same layout: sums of w for frames whose weight differs from the
                              // first one (nullptr while every frame has had the same weight)
    double w;                 // molecule-pair path: weight of the frame being processed
    int priv_solv, priv_sol;  // the solvent / solute group rows are small enough for the shared-memory histograms
};

// MinimumDistance record kept per solvent molecule (i local to the solute molecule, j global)
struct MdRec {
    double d, dref;
    int i, j;
    int flags;   // bit0 within_cutoff, bit1 ref_atom_within_cutoff
    int pad;
};

// ---- exact fp64 arithmetic (never contracted; same operation order as the CPU checker in tests) --
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }

__device__ __forceinline__ void min_image64(const Geom &g, double &x, double &y, double &z) {
    if (g.ortho) {
        // image count = rint(x * (1/L)); differs from rint(x / L) only within one ulp of a half-integer,
        // i.e. for |x| = L/2 >= cutoff where both images are equally far
        x = dsub(x, dmul(g.m[0], rint(dmul(x, g.invl[0]))));
        y = dsub(y, dmul(g.m[4], rint(dmul(y, g.invl[1]))));
        z = dsub(z, dmul(g.m[8], rint(dmul(z, g.invl[2]))));
    } else {
        const double *v = g.inv, *m = g.m;
        double s0 = dadd(dadd(dmul(v[0], x), dmul(v[3], y)), dmul(v[6], z));
        double s1 = dadd(dadd(dmul(v[1], x), dmul(v[4], y)), dmul(v[7], z));
        double s2 = dadd(dadd(dmul(v[2], x), dmul(v[5], y)), dmul(v[8], z));
        s0 = dsub(s0, rint(s0)); s1 = dsub(s1, rint(s1)); s2 = dsub(s2, rint(s2));
        x = dadd(dadd(dmul(m[0], s0), dmul(m[3], s1)), dmul(m[6], s2));
        y = dadd(dadd(dmul(m[1], s0), dmul(m[4], s1)), dmul(m[7], s2));
        z = dadd(dadd(dmul(m[2], s0), dmul(m[5], s1)), dmul(m[8], s2));
    }
}

// minimum-image distance between xi and xj (dr = xj - xi), fp64 exact path
__device__ __forceinline__ double dist_pbc64(const Geom &g, double xi, double yi, double zi, double xj,
                                             double yj, double zj) {
    double dx = dsub(xj, xi), dy = dsub(yj, yi), dz = dsub(zj, zi);
    min_image64(g, dx, dy, dz);
    return __dsqrt_rn(dadd(dadd(dmul(dx, dx), dmul(dy, dy)), dmul(dz, dz)));
}

// wrap a position into the primary cell (cartesian); not part of the exact path: the result feeds grid look-ups only,
// so the image count comes from a multiplication by 1/L (a point within an ulp of a face may land on either side of it;
// both are images of the same point and the grids extend a cutoff beyond the cell)
__device__ __forceinline__ void wrap_to_cell(const Geom &g, double x, double y, double z, double &wx, double &wy,
                                             double &wz) {
    if (g.ortho) {
        wx = x - g.m[0] * floor(x * g.invl[0]);
        wy = y - g.m[4] * floor(y * g.invl[1]);
        wz = z - g.m[8] * floor(z * g.invl[2]);
    } else {
        const double *v = g.inv, *m = g.m;
        double s0 = v[0] * x + v[3] * y + v[6] * z;
        double s1 = v[1] * x + v[4] * y + v[7] * z;
        double s2 = v[2] * x + v[5] * y + v[8] * z;
        s0 -= floor(s0); s1 -= floor(s1); s2 -= floor(s2);
        wx = m[0] * s0 + m[3] * s1 + m[6] * s2;
        wy = m[1] * s0 + m[4] * s1 + m[7] * s2;
        wz = m[2] * s0 + m[5] * s1 + m[8] * s2;
    }
}
// fp32 wrap into the primary cell, relative to the grid centre (culls only)
__device__ __forceinline__ void wrap_to_cell32(const Geom &g, float x, float y, float z, float &wx, float &wy, float &wz) {
    if (g.ortho) {
        wx = x - (float)g.m[0] * floorf(x * (float)g.invl[0]) - (float)g.ctr[0];
        wy = y - (float)g.m[4] * floorf(y * (float)g.invl[1]) - (float)g.ctr[1];
        wz = z - (float)g.m[8] * floorf(z * (float)g.invl[2]) - (float)g.ctr[2];
    } else {
        const double *v = g.inv, *m = g.m;
        float s0 = (float)v[0] * x + (float)v[3] * y + (float)v[6] * z;
        float s1 = (float)v[1] * x + (float)v[4] * y + (float)v[7] * z;
        float s2 = (float)v[2] * x + (float)v[5] * y + (float)v[8] * z;
        s0 -= floorf(s0); s1 -= floorf(s1); s2 -= floorf(s2);
        wx = (float)m[0] * s0 + (float)m[3] * s1 + (float)m[6] * s2 - (float)g.ctr[0];
        wy = (float)m[1] * s0 + (float)m[4] * s1 + (float)m[7] * s2 - (float)g.ctr[1];
        wz = (float)m[2] * s0 + (float)m[5] * s1 + (float)m[8] * s2 - (float)g.ctr[2];
    }
}
// fp32 minimum image of a difference of two fp32 coordinates (bounds only: radii, culls)
__device__ __forceinline__ void min_image32f(const Geom &g, float &x, float &y, float &z) {
    if (g.ortho) {
        x -= (float)g.m[0] * rintf(x * (float)g.invl[0]);
        y -= (float)g.m[4] * rintf(y * (float)g.invl[1]);
        z -= (float)g.m[8] * rintf(z * (float)g.invl[2]);
    } else {
        const double *v = g.inv, *m = g.m;
        float s0 = (float)v[0] * x + (float)v[3] * y + (float)v[6] * z;
        float s1 = (float)v[1] * x + (float)v[4] * y + (float)v[7] * z;
        float s2 = (float)v[2] * x + (float)v[5] * y + (float)v[8] * z;
        s0 -= rintf(s0); s1 -= rintf(s1); s2 -= rintf(s2);
        x = (float)m[0] * s0 + (float)m[3] * s1 + (float)m[6] * s2;
        y = (float)m[1] * s0 + (float)m[4] * s1 + (float)m[7] * s2;
        z = (float)m[2] * s0 + (float)m[5] * s1 + (float)m[8] * s2;
    }
}

// setbin (src/results.jl:28), 0-based
__device__ __forceinline__ int setbin0(double d, double step, int nbins) {
    int ib = (int)ceil(__ddiv_rn(d, step));
    ib = ib < 1 ? 1 : ib;
    ib -= 1;
    return ib >= nbins ? nbins - 1 : ib;
}

// ---- Philox4x32-10; ctr = (slot, sample, frame, block), key = seed -----------------------
__device__ __forceinline__ uint4 philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                            uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}
__device__ __forceinline__ double u01(uint32_t r) { return dmul(dadd((double)r, 0.5), 1.0 / 4294967296.0); }
__device__ __forceinline__ uint32_t pick(uint32_t r, uint32_t n) { return (uint32_t)(((u64)r * (u64)n) >> 32); }

// which solute molecule is the reference of random sample s (src/mddf.jl:374-376)
__device__ __forceinline__ int ref_solute_of_sample(const Prob &P, uint32_t frame, uint32_t s) {
    uint4 r = philox4x32(0xffffffffu, s, frame, 2u, P.seed_lo, P.seed_hi);
    return (int)pick(r.x, (uint32_t)P.ns_mols);
}

// ---- atom position providers -----------------------------------------------------------------
// Real phase: the coordinates of the frame as read (fp32, unwrapped).
struct RealMol {
    const float *x;   // first atom of the molecule
    __device__ __forceinline__ void get(int k, double &px, double &py, double &pz) const {
        px = (double)x[3 * k]; py = (double)x[3 * k + 1]; pz = (double)x[3 * k + 2];
    }
};

// Random phase: a bulk molecule copied from the frame and moved rigidly (randomize_solvent!,
// src/mddf.jl:65-88 + random_move!, src/rigid_body.jl:107-137), regenerated on the fly from
// the Philox counters (frame, sample, slot) -- the random box is never written to memory.
struct RandMol {
    const float *x;          // source molecule (jmol) in the frame
    double ref[3], cm[3], newcm[3], A[9];
    int napm;

    __device__ __forceinline__ void whole(const Geom &g, int k, double &qx, double &qy, double &qz) const {
        double dx = dsub((double)x[3 * k], ref[0]), dy = dsub((double)x[3 * k + 1], ref[1]),
               dz = dsub((double)x[3 * k + 2], ref[2]);
        min_image64(g, dx, dy, dz);   // wrap_relative_to(x[iat], x[irefatom], uc), :129-131
        qx = dadd(ref[0], dx); qy = dadd(ref[1], dy); qz = dadd(ref[2], dz);
    }
    // centre from draw block 0 (r0.y..w), Euler angles from block 1
    __device__ __forceinline__ void init(const Geom &g, const float *src, int napm_, int iref, uint4 r0, uint4 r1) {
        x = src; napm = napm_;
        ref[0] = (double)x[3 * iref]; ref[1] = (double)x[3 * iref + 1]; ref[2] = (double)x[3 * iref + 2];
        double s0 = 0, s1 = 0, s2 = 0;
        for (int k = 0; k < napm; ++k) {
            double qx, qy, qz; whole(g, k, qx, qy, qz);
            s0 = dadd(s0, qx); s1 = dadd(s1, qy); s2 = dadd(s2, qz);
        }
        cm[0] = __ddiv_rn(s0, (double)napm); cm[1] = __ddiv_rn(s1, (double)napm); cm[2] = __ddiv_rn(s2, (double)napm);
        double u0 = u01(r0.y), u1 = u01(r0.z), u2 = u01(r0.w);
        const double *m = g.m;
        newcm[0] = dadd(dadd(dmul(m[0], u0), dmul(m[3], u1)), dmul(m[6], u2));
        newcm[1] = dadd(dadd(dmul(m[1], u0), dmul(m[4], u1)), dmul(m[7], u2));
        newcm[2] = dadd(dadd(dmul(m[2], u0), dmul(m[5], u1)), dmul(m[8], u2));
        const double twopi = 6.283185307179586476925286766559;
        double c1, s1_, c2, s2_, c3, s3_;
#ifdef CMX_SINCOSPI     // sin/cos(2 pi u) as sincospi(2 u): no argument-reduction slow path (no stack frame, fewer registers)
        (void)twopi;
        sincospi(dmul(2.0, u01(r1.x)), &s1_, &c1);
        sincospi(dmul(2.0, u01(r1.y)), &s2_, &c2);
        sincospi(dmul(2.0, u01(r1.z)), &s3_, &c3);
#else
        sincos(dmul(twopi, u01(r1.x)), &s1_, &c1);
        sincos(dmul(twopi, u01(r1.y)), &s2_, &c2);
        sincos(dmul(twopi, u01(r1.z)), &s3_, &c3);
#endif
        // eulermat, src/rigid_body.jl:45-57 (row-major)
        A[0] = dmul(c2, c3);                                   A[1] = dmul(-c2, s3_);                                  A[2] = s2_;
        A[3] = dadd(dmul(c1, s3_), dmul(dmul(c3, s1_), s2_));  A[4] = dsub(dmul(c1, c3), dmul(dmul(s1_, s2_), s3_));   A[5] = dmul(-c2, s1_);
        A[6] = dsub(dmul(s1_, s3_), dmul(dmul(c1, c3), s2_));  A[7] = dadd(dmul(dmul(c1, s2_), s3_), dmul(c3, s1_));   A[8] = dmul(c1, c2);
    }
    __device__ __forceinline__ void get(const Geom &g, int k, double &px, double &py, double &pz) const {
        double qx, qy, qz; whole(g, k, qx, qy, qz);
        double p0 = dsub(qx, cm[0]), p1 = dsub(qy, cm[1]), p2 = dsub(qz, cm[2]);
        // move!, src/rigid_body.jl:73-80
        px = dadd(dadd(dadd(dmul(A[0], p0), dmul(A[1], p1)), dmul(A[2], p2)), newcm[0]);
        py = dadd(dadd(dadd(dmul(A[3], p0), dmul(A[4], p1)), dmul(A[5], p2)), newcm[1]);
        pz = dadd(dadd(dadd(dmul(A[6], p0), dmul(A[7], p1)), dmul(A[8], p2)), newcm[2]);
    }
};

// ---- counters: update_counters!, src/update_counters.jl:43-88 ----------------------------------
// The counters are sums of frame weights (:47,60).  While every frame has the same weight they are kept as exact
// integers (hits) and scaled once at the end; frames with another weight add w (w/2 for the group counts of an
// autocorrelation, :52-53) straight into the fp64 twin of the block -- no fold, no synchronisation between frames.
__device__ __forceinline__ void bump(const Prob &P, u64 *cell, u64 inc, double w) {
    if (P.acc == nullptr) atomicAdd(cell, inc);
    else atomicAdd(P.acc + (cell - P.cnt_base), w * (double)inc);
}
__device__ __forceinline__ void group_add(const Prob &P, u64 *arr, int nbins, int ibin, int pos, int apm, int custom,
                                          const int *off, const int *ids, u64 inc, double w) {
    if (!custom) bump(P, &arr[(size_t)(pos % apm) * nbins + ibin], inc, w);            // atom_type, :9
    else for (int q = off[pos]; q < off[pos + 1]; ++q) bump(P, &arr[(size_t)ids[q] * nbins + ibin], inc, w);
}

// One molecule within the cutoff.  `w` is the frame weight (only used in fp64 mode); `mult` = 2 when one evaluated
// molecule pair stands for both ordered pairs (symmetric autocorrelation pass).
// NOTE (reference quirk kept on purpose, src/update_counters.jl:27): `i` is the atom index WITHIN the current solute
// molecule, so with custom groups and several solute molecules every hit is credited through the group map of the
// FIRST molecule's atoms; `j` is global in the solvent selection.
__device__ __forceinline__ void count_hit(const Prob &P, double w, bool random, double d, int i, int j, u64 mult) {
    int ib = setbin0(d, P.binstep, P.nbins);
    bump(P, &(random ? P.md_r : P.md)[ib], mult, w);
    u64 *gs = random ? P.gsol_r : P.gsol;
    if (P.autocorr) {
        group_add(P, gs, P.nbins, ib, i, P.ns_apm, P.custom_sol, P.sol_off, P.sol_ids, mult, 0.5 * w);
        group_add(P, gs, P.nbins, ib, j, P.ns_apm, P.custom_sol, P.sol_off, P.sol_ids, mult, 0.5 * w);
    } else {
        group_add(P, gs, P.nbins, ib, i, P.ns_apm, P.custom_sol, P.sol_off, P.sol_ids, mult, w);
        group_add(P, random ? P.gsolv_r : P.gsolv, P.nbins, ib, j, P.nv_apm, P.custom_solv, P.solv_off, P.solv_ids, mult, w);
    }
}
__device__ __forceinline__ void count_ref(const Prob &P, double w, bool random, double dref) {
    bump(P, &(random ? P.rdf_r : P.rdf)[setbin0(dref, P.binstep, P.nbins)], 1ull, w);
}

// ---- shared-memory privatised histograms (the "fused histogram kernel" of the path) ----------------------
// md_count, rdf_count and the per-atom-type rows of the solvent receive EVERY hit of a frame in a few hundred bins:
// direct global atomics serialise on a few dozen L2 lines (measured on C4, random phase: 80 % of the finalisation
// kernel, 780 k hits of a batch on 47 lines).  A block therefore counts into a private u32 copy in shared memory
// and adds its non-zero bins to the global u64 counters once, at the end.  Rows: 0 md, 1 rdf, then the solvent
// group rows and the solute group rows when they are few (Prob::priv_solv / priv_sol); everything else -- the
// per-atom rows of a large solute -- stays with global atomics, which are spread over megabytes.
struct HistPriv {
    unsigned *sh;             // [rows][nbins]; nullptr = not privatised (fp64 weight mode)
    int nbins, rows, row_gsolv, row_gsol;
};
__host__ __device__ __forceinline__ int hist_rows(const Prob &P) {
    return 2 + (P.priv_sol ? P.ng_sol : 0) + (P.priv_solv ? P.ng_solv : 0);
}
// (block-wide; ends with a barrier)
__device__ __forceinline__ HistPriv hist_init(const Prob &P, unsigned *smem) {
    HistPriv H; H.nbins = P.nbins; H.sh = P.acc ? nullptr : smem;
    H.row_gsol = P.priv_sol ? 2 : -1;
    H.row_gsolv = P.priv_solv ? 2 + (P.priv_sol ? P.ng_sol : 0) : -1;     // (never set for an autocorrelation: one selection)
    H.rows = hist_rows(P);
    if (H.sh) for (int k = threadIdx.x; k < H.rows * H.nbins; k += blockDim.x) smem[k] = 0u;
    __syncthreads();
    return H;
}
__device__ __forceinline__ void hist_flush(const Prob &P, const HistPriv &H, bool random) {
    __syncthreads();
    if (!H.sh) return;
    u64 *md = random ? P.md_r : P.md, *rdf = random ? P.rdf_r : P.rdf, *gsol = random ? P.gsol_r : P.gsol, *gsolv = random ? P.gsolv_r : P.gsolv;
    for (int k = threadIdx.x; k < H.rows * H.nbins; k += blockDim.x) {
        const unsigned v = H.sh[k];
        if (!v) continue;
        const int row = k / H.nbins, bin = k - row * H.nbins;
        u64 *dst;
        if (row == 0) dst = md;
        else if (row == 1) dst = rdf;
        else if (H.row_gsol >= 0 && row < H.row_gsol + P.ng_sol) dst = gsol + (size_t)(row - H.row_gsol) * H.nbins;
        else dst = gsolv + (size_t)(row - H.row_gsolv) * H.nbins;
        atomicAdd(dst + bin, (u64)v);
    }
}
__device__ __forceinline__ void group_add_priv(const Prob &P, const HistPriv &H, int row0, u64 *arr, int ibin, int pos, int apm, int custom,
                                               const int *off, const int *ids, u64 inc, double w) {
    if (H.sh && row0 >= 0) {
        if (!custom) atomicAdd(&H.sh[(row0 + pos % apm) * H.nbins + ibin], (unsigned)inc);
        else for (int q = off[pos]; q < off[pos + 1]; ++q) atomicAdd(&H.sh[(row0 + ids[q]) * H.nbins + ibin], (unsigned)inc);
    } else group_add(P, arr, P.nbins, ibin, pos, apm, custom, off, ids, inc, w);
}
// count_hit / count_ref through the block's private histograms
__device__ __forceinline__ void count_hit_priv(const Prob &P, const HistPriv &H, double w, bool random, double d, int i, int j, u64 mult) {
    const int ib = setbin0(d, P.binstep, P.nbins);
    if (H.sh) atomicAdd(&H.sh[ib], (unsigned)mult); else bump(P, &(random ? P.md_r : P.md)[ib], mult, w);
    u64 *gs = random ? P.gsol_r : P.gsol;
    if (P.autocorr) {
        group_add_priv(P, H, H.row_gsol, gs, ib, i, P.ns_apm, P.custom_sol, P.sol_off, P.sol_ids, mult, 0.5 * w);
        group_add_priv(P, H, H.row_gsol, gs, ib, j, P.ns_apm, P.custom_sol, P.sol_off, P.sol_ids, mult, 0.5 * w);
    } else {
        group_add_priv(P, H, H.row_gsol, gs, ib, i, P.ns_apm, P.custom_sol, P.sol_off, P.sol_ids, mult, w);
        group_add_priv(P, H, H.row_gsolv, random ? P.gsolv_r : P.gsolv, ib, j, P.nv_apm, P.custom_solv, P.solv_off, P.solv_ids, mult, w);
    }
}
__device__ __forceinline__ void count_ref_priv(const Prob &P, const HistPriv &H, double w, bool random, double dref) {
    const int ib = setbin0(dref, P.binstep, P.nbins);
    if (H.sh) atomicAdd(&H.sh[H.nbins + ib], 1u); else bump(P, &(random ? P.rdf_r : P.rdf)[ib], 1ull, w);
}

// inbulk, src/mddf.jl:55-57
__device__ __forceinline__ bool inbulk(const Prob &P, const MdRec &e) {
    return P.usecutoff ? ((e.flags & 1) && e.d > P.dbulk) : !(e.flags & 1);
}

}  // namespace cmx
