// cmx_pairs.cuh -- molecule-pair path: many small solute molecules (cross- and auto-correlation).
//
// The reference runs minimum_distances! + update_counters! once per solute molecule, an
// O(nmols_solute * nmols_solvent) scan (the author's comment, src/mddf.jl:392-395).  Here one
// pass produces every (solute molecule, solvent molecule) minimum distance:
//   * every molecule is reduced to an anchor (wrapped reference atom, fp64) plus fp32 atom
//     offsets about it and a radius;
//   * solvent anchors are binned in a periodic fractional-space cell list;
//   * one warp owns a solute molecule, streams the neighbouring anchor cells, compacts the
//     candidates (anchor distance <= cutoff + radii) through a warp queue, and each lane then
//     evaluates all atom pairs of one candidate in fp32 (offsets -> ~1e-6 A accuracy);
//   * winners are finalised in fp64 (reference arithmetic) and histogrammed, ambiguous pairs
//     are deferred to an exact fp64 kernel;
//   * autocorrelation evaluates each unordered pair once and counts both ordered pairs
//     (src/minimum_distances.jl:81-99, src/update_counters.jl:48-53).
// Random phase: per sample, the list of the chosen reference solute molecule is computed exactly,
// its bulk molecules are compacted in ascending order (src/mddf.jl:406-415) and the random
// placements are generated and measured on the fly (src/mddf.jl:65-88).
#pragma once

#include "cmx_device.cuh"

namespace cmx {

struct PairGeom {
    double w[3];        // perpendicular widths of the unit cell
    int n[3];           // anchor cells per fractional axis
    float half_wmin;    // half of the smallest perpendicular width (single-image validity)
    float tau;          // fp32 uncertainty of an offset-based distance
    float cut, cut_lo, cut_hi;
};

struct MolData {       // per selection
    double *anchor;     // [nmols][3] wrapped anchor (cartesian)
    float *off;         // [nmols][napm][3] offsets about the anchor (minimum image)
    float *rad;         // [nmols] max |offset|
};

struct PairScratch {
    MolData sol{}, solv{};
    int *cell_count = nullptr, *cell_start = nullptr;   // anchor cells
    int *sorted_id = nullptr;                            // solvent molecule ids sorted by cell
    double *s_anchor = nullptr; float *s_rad = nullptr;  // gathered in sorted order
    float4 *s_anchor4 = nullptr;                         // fp32 {x, y, z (grid-relative), radius} for the candidate pre-test
    MdRec *ref_lists = nullptr;                          // [nrand][nv_mols]
    int *bulk_idx = nullptr, *n_bulk = nullptr;          // [nrand][nv_mols], [nrand]
    u64 *deferred = nullptr; int *def_count = nullptr;   // [cap], [2] (count, overflow)
    size_t def_cap = 0, ncells_cap = 0;
    int sample_chunk = 1;                                // samples of the random phase processed per pass
    float *h_radii = nullptr;                            // pinned: [0] ra_sol, [1] ra_solv, [2] rc_solv
    int *d_radii = nullptr;                              // device float bits, same layout
    float ra_sol_bound = 0, ra_solv_bound = 0;
    bool primed = false;
};

// ---------------------------------------------------------------------------------------------
__global__ void k_mol_prep(Geom g, const float *__restrict__ x, int nmols, int napm, int ianchor, MolData md,
                           int *__restrict__ ra_bits, int *__restrict__ rc_bits) {
    int m = blockIdx.x * blockDim.x + threadIdx.x;
    float ra = 0.f, rc = 0.f;
    if (m < nmols) {
        const float *xm = x + (size_t)3 * napm * m;
        double rx = xm[3 * ianchor], ry = xm[3 * ianchor + 1], rz = xm[3 * ianchor + 2];
        double wx, wy, wz; wrap_to_cell(g, rx, ry, rz, wx, wy, wz);
        md.anchor[3 * (size_t)m] = wx; md.anchor[3 * (size_t)m + 1] = wy; md.anchor[3 * (size_t)m + 2] = wz;
        double sx = 0, sy = 0, sz = 0;
        for (int k = 0; k < napm; ++k) {
            double dx = xm[3 * k] - rx, dy = xm[3 * k + 1] - ry, dz = xm[3 * k + 2] - rz;
            min_image64(g, dx, dy, dz);
            float *o = md.off + ((size_t)m * napm + k) * 3;
            o[0] = (float)dx; o[1] = (float)dy; o[2] = (float)dz;
            ra = fmaxf(ra, (float)sqrt(dx * dx + dy * dy + dz * dz));
            sx += dx; sy += dy; sz += dz;
        }
        sx /= napm; sy /= napm; sz /= napm;
        for (int k = 0; k < napm; ++k) {
            const float *o = md.off + ((size_t)m * napm + k) * 3;
            double dx = o[0] - sx, dy = o[1] - sy, dz = o[2] - sz;
            rc = fmaxf(rc, (float)sqrt(dx * dx + dy * dy + dz * dz));
        }
        ra = ra * 1.000001f + 1e-5f; rc = rc * 1.000001f + 1e-5f;
        md.rad[m] = ra;
    }
    for (int o = 16; o; o >>= 1) { ra = fmaxf(ra, __shfl_xor_sync(0xffffffffu, ra, o)); rc = fmaxf(rc, __shfl_xor_sync(0xffffffffu, rc, o)); }
    if ((threadIdx.x & 31) == 0) {
        if (ra > 0.f) atomicMax(ra_bits, __float_as_int(ra));
        if (rc_bits && rc > 0.f) atomicMax(rc_bits, __float_as_int(rc));
    }
}

// fp32 minimum image for conservative pre-tests (error ~1e-5 * L; callers add a 1e-2 A margin)
__device__ __forceinline__ void min_image32(const Geom &g, float &x, float &y, float &z) {
    if (g.ortho) {
        x -= (float)g.m[0] * rintf(x * (float)g.invl[0]);
        y -= (float)g.m[4] * rintf(y * (float)g.invl[1]);
        z -= (float)g.m[8] * rintf(z * (float)g.invl[2]);
    } else {
        float s0 = (float)g.inv[0] * x + (float)g.inv[3] * y + (float)g.inv[6] * z;
        float s1 = (float)g.inv[1] * x + (float)g.inv[4] * y + (float)g.inv[7] * z;
        float s2 = (float)g.inv[2] * x + (float)g.inv[5] * y + (float)g.inv[8] * z;
        s0 -= rintf(s0); s1 -= rintf(s1); s2 -= rintf(s2);
        x = (float)g.m[0] * s0 + (float)g.m[3] * s1 + (float)g.m[6] * s2;
        y = (float)g.m[1] * s0 + (float)g.m[4] * s1 + (float)g.m[7] * s2;
        z = (float)g.m[2] * s0 + (float)g.m[5] * s1 + (float)g.m[8] * s2;
    }
}

__device__ __forceinline__ void anchor_cell(const Geom &g, const PairGeom &pg, const double *a, int &cx, int &cy, int &cz) {
    double s0, s1, s2;
    if (g.ortho) { s0 = a[0] / g.m[0]; s1 = a[1] / g.m[4]; s2 = a[2] / g.m[8]; }
    else {
        const double *v = g.inv;
        s0 = v[0] * a[0] + v[3] * a[1] + v[6] * a[2];
        s1 = v[1] * a[0] + v[4] * a[1] + v[7] * a[2];
        s2 = v[2] * a[0] + v[5] * a[1] + v[8] * a[2];
    }
    s0 -= floor(s0); s1 -= floor(s1); s2 -= floor(s2);
    cx = min(max((int)(s0 * pg.n[0]), 0), pg.n[0] - 1);
    cy = min(max((int)(s1 * pg.n[1]), 0), pg.n[1] - 1);
    cz = min(max((int)(s2 * pg.n[2]), 0), pg.n[2] - 1);
}

template <bool SCATTER>
__global__ void k_anchor_bin(Geom g, PairGeom pg, MolData md, int nmols, int *__restrict__ cell_count,
                             const int *__restrict__ cell_start, int *__restrict__ sorted_id,
                             double *__restrict__ s_anchor, float *__restrict__ s_rad, float4 *__restrict__ s_anchor4) {
    int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= nmols) return;
    int cx, cy, cz; anchor_cell(g, pg, md.anchor + 3 * (size_t)m, cx, cy, cz);
    int c = (cz * pg.n[1] + cy) * pg.n[0] + cx;
    if (!SCATTER) atomicAdd(&cell_count[c], 1);
    else {
        int slot = cell_start[c] + atomicSub(&cell_count[c], 1) - 1;
        sorted_id[slot] = m;
        s_anchor[3 * (size_t)slot] = md.anchor[3 * (size_t)m];
        s_anchor[3 * (size_t)slot + 1] = md.anchor[3 * (size_t)m + 1];
        s_anchor[3 * (size_t)slot + 2] = md.anchor[3 * (size_t)m + 2];
        s_rad[slot] = md.rad[m];
        s_anchor4[slot] = make_float4((float)(md.anchor[3 * (size_t)m] - g.ctr[0]), (float)(md.anchor[3 * (size_t)m + 1] - g.ctr[1]),
                                      (float)(md.anchor[3 * (size_t)m + 2] - g.ctr[2]), md.rad[m]);
    }
}

// ---- fp32 evaluation of one molecule pair from offsets -------------------------------------------
struct PairFound {
    float b1, b2; int i, j;       // best / second best / winning atoms (i in a, j in b)
    float r1, r2; int ri;         // atoms of a -> reference atom of b
    float q1, q2; int qj;         // atoms of b -> reference atom of a (symmetric pass only)
};

__device__ __forceinline__ void upd(float d2, float &b1, float &b2) { if (d2 < b1) { b2 = b1; b1 = d2; } else b2 = fminf(b2, d2); }

// offa: float4 per atom of a (x, y, z, -) in shared memory; offb: xyz triplets of b in global memory.
// The main loop only tracks best / second best; the reference-atom rows (atoms of a -> ref atom of b and, in
// the symmetric pass, atoms of b -> ref atom of a) are re-evaluated afterwards (napm_a + napm_b extra pairs).
template <bool SYM>
__device__ __forceinline__ PairFound eval_pair(const float4 *__restrict__ offa, int napm_a,
                                               const float *__restrict__ offb, int napm_b, float Dx, float Dy,
                                               float Dz, int irefb, int irefa) {
    PairFound F; F.b1 = F.b2 = F.r1 = F.r2 = F.q1 = F.q2 = CUDART_INF_F; F.i = F.j = F.ri = F.qj = -1;
    int best_pair = 0, pair = 0;          // running index j * napm_a + i of the current / best atom pair
    for (int j = 0; j < napm_b; ++j) {
        float vx = Dx + offb[3 * j], vy = Dy + offb[3 * j + 1], vz = Dz + offb[3 * j + 2];
#pragma unroll 2
        for (int i = 0; i < napm_a; ++i, ++pair) {
            float4 a = offa[i];
            float dx = vx - a.x, dy = vy - a.y, dz = vz - a.z;
            float d2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
            F.b2 = fminf(F.b2, fmaxf(d2, F.b1));
            if (d2 < F.b1) { F.b1 = d2; best_pair = pair; }
        }
    }
    F.j = best_pair / napm_a; F.i = best_pair - F.j * napm_a;
    {   // atoms of a -> reference atom of b
        float vx = Dx + offb[3 * irefb], vy = Dy + offb[3 * irefb + 1], vz = Dz + offb[3 * irefb + 2];
        for (int i = 0; i < napm_a; ++i) {
            float4 a = offa[i];
            float dx = vx - a.x, dy = vy - a.y, dz = vz - a.z;
            float d2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
            F.r2 = fminf(F.r2, fmaxf(d2, F.r1));
            if (d2 < F.r1) { F.r1 = d2; F.ri = i; }
        }
    }
    if (SYM) {   // atoms of b -> reference atom of a
        float4 a = offa[irefa];
        for (int j = 0; j < napm_b; ++j) {
            float dx = Dx + offb[3 * j] - a.x, dy = Dy + offb[3 * j + 1] - a.y, dz = Dz + offb[3 * j + 2] - a.z;
            float d2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
            F.q2 = fminf(F.q2, fmaxf(d2, F.q1));
            if (d2 < F.q1) { F.q1 = d2; F.qj = j; }
        }
    }
    return F;
}

__device__ __forceinline__ int classify_p(const PairGeom &pg, float b1, float b2) {
    float d1 = sqrtf(b1);
    if (d1 > pg.cut_hi) return 0;
    if (d1 >= pg.cut_lo) return 2;
    if (sqrtf(b2) - d1 <= pg.tau) return 2;
    return 1;
}

__device__ __forceinline__ void defer_pair(u64 *deferred, int *def_count, size_t cap, int phase, int a, int b) {
    int slot = atomicAdd(def_count, 1);
    if ((size_t)slot < cap) deferred[slot] = ((u64)phase << 48) | ((u64)(uint32_t)a << 24) | (u64)(uint32_t)b;
    else atomicExch(def_count + 1, 1);   // overflow flag
}

__device__ __forceinline__ double exact_atoms(const Geom &g, const float *xa, int i, const float *xb, int j) {
    return dist_pbc64(g, (double)xa[3 * i], (double)xa[3 * i + 1], (double)xa[3 * i + 2], (double)xb[3 * j],
                      (double)xb[3 * j + 1], (double)xb[3 * j + 2]);
}

// finalise one evaluated pair: classify, exact recompute, count (or defer)
template <bool SYM>
__device__ __forceinline__ void finish_pair(const Geom &g, const PairGeom &pg, const Prob &P, const PairFound &F,
                                            bool single_image_ok, const float *__restrict__ xs,
                                            const float *__restrict__ xv, int a, int b, u64 *deferred, int *def_count,
                                            size_t cap) {
    int cls = classify_p(pg, F.b1, F.b2);
    if (cls == 0 && single_image_ok) return;
    int rcls = classify_p(pg, F.r1, F.r2), qcls = SYM ? classify_p(pg, F.q1, F.q2) : 0;
    if (!single_image_ok || cls == 2 || rcls == 2 || (SYM && qcls == 2)) { defer_pair(deferred, def_count, cap, 0, a, b); return; }
    const float *xa = xs + (size_t)3 * P.ns_apm * a, *xb = xv + (size_t)3 * P.nv_apm * b;
    double d = exact_atoms(g, xa, F.i, xb, F.j);
    count_hit(P, P.w, false, d, F.i, b * P.nv_apm + F.j, 1ull);
    if (rcls == 1) count_ref(P, P.w, false, exact_atoms(g, xa, F.ri, xb, P.iref));
    if (SYM) {   // the ordered pair (solute b, solvent a)
        count_hit(P, P.w, false, d, F.j, a * P.nv_apm + F.i, 1ull);
        if (qcls == 1) count_ref(P, P.w, false, exact_atoms(g, xb, F.qj, xa, P.iref));
    }
}

// ---------------------------------------------------------------------------------------------
// main pair kernel: one warp per solute molecule
// ---------------------------------------------------------------------------------------------
#define CMX_PAIR_WARPS 4
template <bool SYM>
__global__ void __launch_bounds__(CMX_PAIR_WARPS * 32, 6)
k_pairs(Geom g, PairGeom pg, Prob P, const float *__restrict__ xs, const float *__restrict__ xv, MolData sol,
        MolData solv, const int *__restrict__ cell_start, const int *__restrict__ sorted_id,
        const double *__restrict__ s_anchor, const float *__restrict__ s_rad, const float4 *__restrict__ s_anchor4,
        const int *__restrict__ ra_solv_bits,
        u64 *__restrict__ deferred, int *__restrict__ def_count, size_t def_cap, u64 *__restrict__ pair_evals, int nsplit) {
    extern __shared__ float4 smem4[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float4 *offa = smem4 + (size_t)warp * P.ns_apm;
    int *queue = (int *)(smem4 + (size_t)CMX_PAIR_WARPS * P.ns_apm) + warp * 64;
    const float ra_solv_max = __int_as_float(*ra_solv_bits);
    unsigned long long npairs = 0;
    // a solute molecule's neighbour cells are dealt to `nsplit` warps (more warps in flight for few molecules)
    for (int task = blockIdx.x * CMX_PAIR_WARPS + warp; task < P.ns_mols * nsplit; task += gridDim.x * CMX_PAIR_WARPS) {
        const int a = task / nsplit, part = task - a * nsplit;
        __syncwarp();
        for (int t = lane; t < P.ns_apm; t += 32) {
            const float *o = sol.off + ((size_t)a * P.ns_apm + t) * 3;
            offa[t] = make_float4(o[0], o[1], o[2], 0.f);
        }
        __syncwarp();
        const double ax = sol.anchor[3 * (size_t)a], ay = sol.anchor[3 * (size_t)a + 1], az = sol.anchor[3 * (size_t)a + 2];
        const float ra = sol.rad[a];
        const float axf = (float)(ax - g.ctr[0]), ayf = (float)(ay - g.ctr[1]), azf = (float)(az - g.ctr[2]);
        const float reach = pg.cut_hi + ra + ra_solv_max + 1e-3f;
        // single-image regime: every candidate's atom-pair vectors built from ONE image of the anchor
        // difference are shorter than half the smallest cell width, hence true minimum images
        const bool regime = pg.cut_hi + 2.f * (ra + ra_solv_max) + 3e-2f < pg.half_wmin;
        int c0[3]; { double aa[3] = {ax, ay, az}; anchor_cell(g, pg, aa, c0[0], c0[1], c0[2]); }
        int lo[3], cnt[3];
        for (int k = 0; k < 3; ++k) {
            int nr = (int)ceil((double)reach / (pg.w[k] / pg.n[k]));
            if (2 * nr + 1 >= pg.n[k]) { lo[k] = 0; cnt[k] = pg.n[k]; } else { lo[k] = c0[k] - nr; cnt[k] = 2 * nr + 1; }
        }
        int qn = 0;
        const int ncell = cnt[0] * cnt[1] * cnt[2];
        auto process = [&](int sidx) {
            int b = sorted_id[sidx];
            double dx = dsub(s_anchor[3 * (size_t)sidx], ax), dy = dsub(s_anchor[3 * (size_t)sidx + 1], ay),
                   dz = dsub(s_anchor[3 * (size_t)sidx + 2], az);
            min_image64(g, dx, dy, dz);
            float Dx = (float)dx, Dy = (float)dy, Dz = (float)dz;
            bool ok = regime;
            PairFound F = eval_pair<SYM>(offa, P.ns_apm, solv.off + (size_t)b * 3 * P.nv_apm, P.nv_apm, Dx, Dy, Dz, P.iref, P.iref);
            npairs += (unsigned long long)P.ns_apm * P.nv_apm;
            finish_pair<SYM>(g, pg, P, F, ok, xs, xv, a, b, deferred, def_count, def_cap);
        };
        for (int cc = part; cc < ncell; cc += nsplit) {
            int iz = cc / (cnt[0] * cnt[1]); int rem = cc - iz * (cnt[0] * cnt[1]);
            int iy = rem / cnt[0], ix = rem - iy * cnt[0];
            int cx = lo[0] + ix; cx += cx < 0 ? pg.n[0] : 0; cx -= cx >= pg.n[0] ? pg.n[0] : 0;
            int cy = lo[1] + iy; cy += cy < 0 ? pg.n[1] : 0; cy -= cy >= pg.n[1] ? pg.n[1] : 0;
            int cz = lo[2] + iz; cz += cz < 0 ? pg.n[2] : 0; cz -= cz >= pg.n[2] ? pg.n[2] : 0;
            int c = (cz * pg.n[1] + cy) * pg.n[0] + cx;
            int beg = cell_start[c], end = cell_start[c + 1];
            for (int base = beg; base < end; base += 32) {
                int sidx = base + lane;
                bool pass = false;
                if (sidx < end) {
                    int b = sorted_id[sidx];
                    if (!(SYM && b <= a)) {   // autocorrelation: each unordered pair once; never the molecule itself
                        float4 c4 = __ldg(&s_anchor4[sidx]);            // fp32 pre-test with a 1e-2 A margin; exact Delta later
                        float dx = c4.x - axf, dy = c4.y - ayf, dz = c4.z - azf;
                        min_image32(g, dx, dy, dz);
                        float lim = pg.cut_hi + ra + c4.w + 1e-2f;
                        pass = dx * dx + dy * dy + dz * dz <= lim * lim;
                    }
                }
                unsigned ball = __ballot_sync(0xffffffffu, pass);
                if (pass) queue[qn + __popc(ball & ((1u << lane) - 1))] = sidx;
                qn += __popc(ball);
                __syncwarp();
                if (qn >= 32) {
                    int mine = queue[lane];
                    __syncwarp();
                    int rest = qn - 32;
                    int moved = lane < rest ? queue[32 + lane] : 0;
                    __syncwarp();
                    if (lane < rest) queue[lane] = moved;
                    qn = rest;
                    __syncwarp();
                    process(mine);
                }
            }
        }
        __syncwarp();
        if (lane < qn) process(queue[lane]);
    }
    if (pair_evals) {
        for (int o = 16; o; o >>= 1) npairs += __shfl_xor_sync(0xffffffffu, npairs, o);
        if (lane == 0 && npairs) atomicAdd(pair_evals, npairs);
    }
}

// ---------------------------------------------------------------------------------------------
// exact list of one solute molecule against every solvent molecule (fp64 brute force with an
// anchor-distance early-out).  Used for the bulk lists of the random phase and for the parity hook.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ MdRec exact_list_entry(const Geom &g, const Prob &P, const float *xa, const float *xb, int b) {
    MdRec e; e.d = CUDART_INF; e.dref = CUDART_INF; e.i = -1; e.j = -1; e.flags = 0; e.pad = 0;
    for (int j = 0; j < P.nv_apm; ++j)
        for (int i = 0; i < P.ns_apm; ++i) {
            double d = exact_atoms(g, xa, i, xb, j);
            if (d <= g.cutd) {
                int jg = b * P.nv_apm + j;
                if (d < e.d || (d == e.d && (jg < e.j || (jg == e.j && i < e.i)))) { e.d = d; e.i = i; e.j = jg; e.flags |= 1; }
                if (j == P.iref && d < e.dref) { e.dref = d; e.flags |= 2; }
            }
        }
    return e;
}

// the same list entry computed by a whole warp: the lanes stride over the atom pairs and the partial results are merged
// with the rule "smallest (d, j, i)" (every lane returns the entry)
__device__ __forceinline__ MdRec exact_list_entry_warp(const Geom &g, const Prob &P, const float *xa, const float *xb, int b) {
    const int lane = threadIdx.x & 31;
    double bd = CUDART_INF, bref = CUDART_INF; int bi = 0x7fffffff, bj = 0x7fffffff;
    const int npair = P.ns_apm * P.nv_apm;
    for (int p = lane; p < npair; p += 32) {
        const int j = p / P.ns_apm, i = p - j * P.ns_apm;
        const double d = exact_atoms(g, xa, i, xb, j);
        if (d <= g.cutd) {
            const int jg = b * P.nv_apm + j;
            if (d < bd || (d == bd && (jg < bj || (jg == bj && i < bi)))) { bd = d; bi = i; bj = jg; }
            if (j == P.iref && d < bref) bref = d;
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const double od = __shfl_xor_sync(0xffffffffu, bd, o), oref = __shfl_xor_sync(0xffffffffu, bref, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o), oj = __shfl_xor_sync(0xffffffffu, bj, o);
        if (od < bd || (od == bd && (oj < bj || (oj == bj && oi < bi)))) { bd = od; bi = oi; bj = oj; }
        bref = fmin(bref, oref);
    }
    MdRec e; e.d = CUDART_INF; e.dref = CUDART_INF; e.i = -1; e.j = -1; e.flags = 0; e.pad = 0;
    if (bd <= g.cutd) { e.d = bd; e.i = bi; e.j = bj; e.flags = 1; if (bref <= g.cutd) { e.dref = bref; e.flags |= 2; } }
    return e;
}

// grid.y = list index within the current chunk of samples; solute molecule = fixed_a (>= 0) or the reference
// solute of sample s0 + blockIdx.y
__global__ void k_ref_lists(Geom g, PairGeom pg, Prob P, uint32_t frame, int fixed_a, int s0, const float *__restrict__ xs,
                            const float *__restrict__ xv, MolData sol, MolData solv, MdRec *__restrict__ lists) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31;
    const int s = blockIdx.y;
    const int a = fixed_a >= 0 ? fixed_a : ref_solute_of_sample(P, frame, (uint32_t)(s0 + s));
    MdRec e; e.d = CUDART_INF; e.dref = CUDART_INF; e.i = -1; e.j = -1; e.flags = 0; e.pad = 0;
    bool heavy = false;
    if (b < P.nv_mols && !(P.autocorr && b == a)) {
        double dx = solv.anchor[3 * (size_t)b] - sol.anchor[3 * (size_t)a], dy = solv.anchor[3 * (size_t)b + 1] - sol.anchor[3 * (size_t)a + 1],
               dz = solv.anchor[3 * (size_t)b + 2] - sol.anchor[3 * (size_t)a + 2];
        min_image64(g, dx, dy, dz);
        double lim = g.cutd + sol.rad[a] + solv.rad[b] + 1e-3;
        double dn = sqrt(dx * dx + dy * dy + dz * dz);
        // |v0| > lim implies the true anchor distance > lim whenever lim < half the smallest width;
        // otherwise the anchor test is not a valid bound: evaluate everything
        heavy = dn <= lim || lim >= pg.half_wmin;
    }
    // the few molecules that pass the anchor test are evaluated by the whole warp, one after the other
    unsigned todo = __ballot_sync(0xffffffffu, heavy);
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const int bb = __shfl_sync(0xffffffffu, b, src);
        const MdRec r = exact_list_entry_warp(g, P, xs + (size_t)3 * P.ns_apm * a, xv + (size_t)3 * P.nv_apm * bb, bb);
        if (lane == src) e = r;
    }
    if (b < P.nv_mols) lists[(size_t)s * P.nv_mols + b] = e;
}

// block-wide exclusive prefix sum of one int per thread (blockDim.x = THREADS, a multiple of 32): position of the thread's
// value and the block total.  warp_sums: THREADS / 32 ints of shared memory, free to reuse after the call returns.
template <int THREADS>
__device__ __forceinline__ void block_exclusive_sum(int v, int &pos, int &total, int *warp_sums) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    __syncthreads();                                   // (warp_sums of a previous call are no longer read)
    if (lane == 31) warp_sums[w] = incl;
    __syncthreads();
    int before = 0, all = 0;
#pragma unroll
    for (int k = 0; k < THREADS / 32; ++k) { const int sk = warp_sums[k]; before += k < w ? sk : 0; all += sk; }
    pos = before + incl - v; total = all;
}

// exclusive prefix sum of the anchor-cell counts (n = cells + 1, a few hundred to a few thousand entries): ONE block walks
// the array in chunks of 1024 -- the pair path has no library kernel either
__global__ void __launch_bounds__(1024)
k_scan_block(const int *__restrict__ in, int *__restrict__ out, int n) {
    __shared__ int warp_sums[32];
    int carry = 0;
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < n ? in[i] : 0;
        int pos, total;
        block_exclusive_sum<1024>(v, pos, total, warp_sums);
        if (i < n) out[i] = carry + pos;
        carry += total;
    }
}

// ordered compaction of the bulk molecules of each sample's list (one block per sample)
__global__ void __launch_bounds__(512)
k_bulk_compact(Prob P, uint32_t frame, int s0, const MdRec *__restrict__ lists, int *__restrict__ bulk_idx, int *__restrict__ n_bulk) {
    __shared__ int warp_sums[16];
    __shared__ int base_sh;
    int s = blockIdx.x;                      // sample within the chunk
    int a = ref_solute_of_sample(P, frame, (uint32_t)(s0 + s));
    if (threadIdx.x == 0) base_sh = 0;
    __syncthreads();
    for (int start = 0; start < P.nv_mols; start += 512) {
        int m = start + threadIdx.x;
        int f = 0;
        if (m < P.nv_mols && !(P.autocorr && m == a)) f = inbulk(P, lists[(size_t)s * P.nv_mols + m]) ? 1 : 0;
        int pos, total;
        block_exclusive_sum<512>(f, pos, total, warp_sums);
        int base = base_sh;
        if (f) bulk_idx[(size_t)s * P.nv_mols + base + pos] = m;
        __syncthreads();
        if (threadIdx.x == 0) base_sh = base + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) n_bulk[s] = base_sh;
}

// ---------------------------------------------------------------------------------------------
// random phase: one thread per (sample, slot)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_pair_random(Geom g, PairGeom pg, Prob P, uint32_t frame, int s0, int ns, const float *__restrict__ xs, const float *__restrict__ xv,
              MolData sol, const int *__restrict__ rc_solv_bits, const int *__restrict__ bulk_idx,
              const int *__restrict__ n_bulk, MdRec *__restrict__ rand_list, u64 *__restrict__ deferred,
              int *__restrict__ def_count, size_t def_cap) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)ns * P.nv_mols) return;
    const int sl = (int)(t / P.nv_mols), slot = (int)(t - (long long)sl * P.nv_mols);   // sl: sample within the chunk
    const int s = s0 + sl;
    int a = ref_solute_of_sample(P, frame, (uint32_t)s);
    if (P.autocorr && slot == a) return;   // src/minimum_distances.jl:90
    const float rc = __int_as_float(*rc_solv_bits);
    uint4 r0 = philox4x32((uint32_t)slot, (uint32_t)s, frame, 0u, P.seed_lo, P.seed_hi);
    const double ax = sol.anchor[3 * (size_t)a], ay = sol.anchor[3 * (size_t)a + 1], az = sol.anchor[3 * (size_t)a + 2];
    const float ra = sol.rad[a];
    {   // cull by the new centre
        double u0 = u01(r0.y), u1 = u01(r0.z), u2 = u01(r0.w);
        const double *m = g.m;
        double dx = m[0] * u0 + m[3] * u1 + m[6] * u2 - ax, dy = m[1] * u0 + m[4] * u1 + m[7] * u2 - ay,
               dz = m[2] * u0 + m[5] * u1 + m[8] * u2 - az;
        min_image64(g, dx, dy, dz);
        double lim = (double)pg.cut_hi + ra + rc + 1e-3;
        double dn2 = dx * dx + dy * dy + dz * dz;
        if (dn2 > lim * lim && lim < pg.half_wmin) return;   // true centre distance > lim as well
    }
    uint4 r1 = philox4x32((uint32_t)slot, (uint32_t)s, frame, 1u, P.seed_lo, P.seed_hi);
    int nb = n_bulk[sl];
    int jmol = nb > 0 ? bulk_idx[(size_t)sl * P.nv_mols + pick(r0.x, (uint32_t)nb)] : (int)pick(r0.x, (uint32_t)P.nv_mols);
    RandMol rm; rm.init(g, xv + (size_t)3 * P.nv_apm * jmol, P.nv_apm, P.iref, r0, r1);
    const float *offa = sol.off + (size_t)a * 3 * P.ns_apm;
    float b1 = CUDART_INF_F, b2 = CUDART_INF_F, r1_ = CUDART_INF_F, r2_ = CUDART_INF_F;
    int bi = -1, bk = -1, ri = -1;
    // single-image regime for atom-to-anchor vectors (see k_pairs)
    const bool ok = pg.cut_hi + 2.f * ra + 2e-3f < pg.half_wmin;
    const float skip2 = (pg.cut_hi + ra + 1e-3f) * (pg.cut_hi + ra + 1e-3f);
    for (int k = 0; k < P.nv_apm; ++k) {
        double ex, ey, ez; rm.get(g, k, ex, ey, ez);
        double dx = dsub(ex, ax), dy = dsub(ey, ay), dz = dsub(ez, az);
        min_image64(g, dx, dy, dz);
        float vx = (float)dx, vy = (float)dy, vz = (float)dz;
        if (ok && vx * vx + vy * vy + vz * vz > skip2) continue;   // this atom is beyond the cutoff of every solute atom
        for (int i = 0; i < P.ns_apm; ++i) {
            float qx = vx - offa[3 * i], qy = vy - offa[3 * i + 1], qz = vz - offa[3 * i + 2];
            float d2 = fmaf(qx, qx, fmaf(qy, qy, qz * qz));
            if (d2 < b1) { b2 = b1; b1 = d2; bi = i; bk = k; } else b2 = fminf(b2, d2);
            if (k == P.iref) { if (d2 < r1_) { r2_ = r1_; r1_ = d2; ri = i; } else r2_ = fminf(r2_, d2); }
        }
    }
    int cls = classify_p(pg, b1, b2);
    if (cls == 0 && ok) return;
    int rcls = classify_p(pg, r1_, r2_);
    if (!ok || cls == 2 || rcls == 2) { defer_pair(deferred, def_count, def_cap, 1 + sl, a, slot); return; }
    const float *xa = xs + (size_t)3 * P.ns_apm * a;
    double ex, ey, ez; rm.get(g, bk, ex, ey, ez);
    MdRec e; e.pad = 0; e.flags = 1; e.i = bi; e.j = slot * P.nv_apm + bk; e.dref = CUDART_INF;
    e.d = dist_pbc64(g, (double)xa[3 * bi], (double)xa[3 * bi + 1], (double)xa[3 * bi + 2], ex, ey, ez);
    count_hit(P, P.w, true, e.d, e.i, e.j, 1ull);
    if (rcls == 1) {
        rm.get(g, P.iref, ex, ey, ez);
        e.dref = dist_pbc64(g, (double)xa[3 * ri], (double)xa[3 * ri + 1], (double)xa[3 * ri + 2], ex, ey, ez);
        e.flags |= 2;
        count_ref(P, P.w, true, e.dref);
    }
    if (rand_list) rand_list[(size_t)s * P.nv_mols + slot] = e;
}

// ---------------------------------------------------------------------------------------------
// exact resolve of deferred pairs: one WARP per item (the lanes share the ns_apm x nv_apm exact fp64 distances; one
// thread per item made the kernel as long as one thread's ~400 serial fp64 distances)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_pair_resolve(Geom g, Prob P, uint32_t frame, int s0, const float *__restrict__ xs, const float *__restrict__ xv,
               const int *__restrict__ bulk_idx, const int *__restrict__ n_bulk,
               const u64 *__restrict__ deferred, const int *__restrict__ def_count, size_t def_cap,
               MdRec *__restrict__ rand_list) {
    const int count = (int)min((long long)*def_count, (long long)def_cap);
    const int lane = threadIdx.x & 31, nwarp = (gridDim.x * blockDim.x) >> 5;
    for (int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < count; w += nwarp) {
        u64 item = deferred[w];
        int phase = (int)(item >> 48), a = (int)((item >> 24) & 0xffffffull), b = (int)(item & 0xffffffull);
        const float *xa = xs + (size_t)3 * P.ns_apm * a;
        if (phase == 0) {
            const float *xb = xv + (size_t)3 * P.nv_apm * b;
            MdRec e = exact_list_entry_warp(g, P, xa, xb, b);
            if (lane == 0 && (e.flags & 1)) { count_hit(P, P.w, false, e.d, e.i, e.j, 1ull); if (e.flags & 2) count_ref(P, P.w, false, e.dref); }
            if (P.autocorr) {   // the other ordered pair: solute b, solvent a
                MdRec f = exact_list_entry_warp(g, P, xb, xa, a);
                if (lane == 0 && (f.flags & 1)) { count_hit(P, P.w, false, f.d, f.i, f.j, 1ull); if (f.flags & 2) count_ref(P, P.w, false, f.dref); }
            }
        } else {
            int sl = phase - 1, s = s0 + sl, slot = b;    // sl: sample within the chunk
            uint4 r0 = philox4x32((uint32_t)slot, (uint32_t)s, frame, 0u, P.seed_lo, P.seed_hi);
            uint4 r1 = philox4x32((uint32_t)slot, (uint32_t)s, frame, 1u, P.seed_lo, P.seed_hi);
            int nb = n_bulk[sl];
            int jmol = nb > 0 ? bulk_idx[(size_t)sl * P.nv_mols + pick(r0.x, (uint32_t)nb)] : (int)pick(r0.x, (uint32_t)P.nv_mols);
            RandMol rm; rm.init(g, xv + (size_t)3 * P.nv_apm * jmol, P.nv_apm, P.iref, r0, r1);
            double bd = CUDART_INF, bref = CUDART_INF; int bi = 0x7fffffff, bj = 0x7fffffff;
            const int npair = P.ns_apm * P.nv_apm;
            for (int p = lane; p < npair; p += 32) {
                const int k = p / P.ns_apm, i = p - k * P.ns_apm;
                double ex, ey, ez; rm.get(g, k, ex, ey, ez);
                const int jg = slot * P.nv_apm + k;
                const double d = dist_pbc64(g, (double)xa[3 * i], (double)xa[3 * i + 1], (double)xa[3 * i + 2], ex, ey, ez);
                if (d <= g.cutd) {
                    if (d < bd || (d == bd && (jg < bj || (jg == bj && i < bi)))) { bd = d; bi = i; bj = jg; }
                    if (k == P.iref && d < bref) bref = d;
                }
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                const double od = __shfl_xor_sync(0xffffffffu, bd, o), oref = __shfl_xor_sync(0xffffffffu, bref, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o), oj = __shfl_xor_sync(0xffffffffu, bj, o);
                if (od < bd || (od == bd && (oj < bj || (oj == bj && oi < bi)))) { bd = od; bi = oi; bj = oj; }
                bref = fmin(bref, oref);
            }
            if (lane == 0) {
                MdRec e; e.d = CUDART_INF; e.dref = CUDART_INF; e.i = -1; e.j = -1; e.flags = 0; e.pad = 0;
                if (bd <= g.cutd) { e.d = bd; e.i = bi; e.j = bj; e.flags = 1; if (bref <= g.cutd) { e.dref = bref; e.flags |= 2; } }
                if (e.flags & 1) { count_hit(P, P.w, true, e.d, e.i, e.j, 1ull); if (e.flags & 2) count_ref(P, P.w, true, e.dref); }
                if (rand_list) rand_list[(size_t)s * P.nv_mols + slot] = e;
            }
        }
    }
}

__global__ void k_accumulate_stats(const int *__restrict__ a, const int *__restrict__ b, u64 *__restrict__ stats) {
    if (threadIdx.x == 0 && blockIdx.x == 0) atomicAdd(&stats[1], (u64)(a ? *a : 0) + (u64)(b ? *b : 0));   // frames overlap on several streams
}

}  // namespace cmx
