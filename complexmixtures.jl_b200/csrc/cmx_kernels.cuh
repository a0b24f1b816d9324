// cmx_kernels.cuh -- the sm_100a kernels of the grid path (one large solute molecule against
// every solvent molecule; real phase and random ideal-gas phase).
//
// Replaces, per frame, the CellListMap.pairwise! traversal driven by minimum_distances!
// (src/minimum_distances.jl:129-148), update_counters! (src/update_counters.jl:43-88),
// randomize_solvent!/random_move! (src/mddf.jl:65-88, src/rigid_body.jl:107-137) and the
// orchestration of mddf_frame! (src/mddf.jl:361-429).
//
// Algorithm (not a translation of the reference's "all pairs within the cutoff" visit):
//   1. the solute atoms and their periodic images are binned into a fine cartesian grid
//      (cell-sorted float4 {x,y,z,index}); a coarse bitmap of occupied cells is turned into a
//      Chebyshev distance map;
//   2. solvent molecules are culled with the distance map; survivors go to a work list;
//   3. a group of G lanes owns one surviving molecule and runs, atom by atom, a pruned
//      nearest-neighbour search over grid rows ordered by distance (fp32, vector loads of the
//      cell-sorted solute), keeping best / second-best squared distances;
//   4. the winning pair is re-evaluated in fp64 with the reference's minimum-image arithmetic and
//      histogrammed; molecules whose fp32 result is ambiguous (near-tie, cutoff edge) are deferred to
//      an exact fp64 resolve kernel, so the counts equal the fp64 oracle's bit for bit.
//   The random phase regenerates every random molecule from Philox counters inside the same
//   search kernel; the random box is never materialised.
#pragma once
#include "cmx_device.cuh"

namespace cmx {

#define CMX_MAX_ROWTAB 289   // (2*8+1)^2
__constant__ short c_row_dy[CMX_MAX_ROWTAB];
__constant__ short c_row_dz[CMX_MAX_ROWTAB];
__constant__ float c_row_lb[CMX_MAX_ROWTAB];   // lower bound of the row distance^2, in units of side^2

// ---------------------------------------------------------------------------------------------
// K1/K3: bin the solute molecule (plus periodic images inside the extended box) into the grid
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int fine_cell_of(const Geom &g, float px, float py, float pz) {
    int cx = (int)floorf((px - g.gmin[0]) * g.inv_sidex);
    int cy = (int)floorf((py - g.gmin[1]) * g.inv_side);
    int cz = (int)floorf((pz - g.gmin[2]) * g.inv_side);
    cx = min(max(cx, 0), g.nx - 1); cy = min(max(cy, 0), g.ny - 1); cz = min(max(cz, 0), g.nz - 1);
    return (cz * g.ny + cy) * g.nx + cx;
}
__device__ __forceinline__ void coarse_cell_of(const Geom &g, float px, float py, float pz, int &cx, int &cy, int &cz) {
    cx = (int)floorf((px - g.gmin[0]) * g.inv_cside);
    cy = (int)floorf((py - g.gmin[1]) * g.inv_cside);
    cz = (int)floorf((pz - g.gmin[2]) * g.inv_cside);
    cx = min(max(cx, 0), g.ncx - 1); cy = min(max(cy, 0), g.ncy - 1); cz = min(max(cz, 0), g.ncz - 1);
}

// SCATTER=false: count images per cell and mark coarse occupancy; SCATTER=true: write the
// cell-sorted float4 array (cell_count is consumed as the per-cell fill counter).
template <bool SCATTER>
__global__ void k_solute_bin(Geom g, const float *__restrict__ xs, int natoms, int *__restrict__ cell_count,
                             const int *__restrict__ cell_start, u64 *__restrict__ occ_bits,
                             u64 *__restrict__ rowmask, float4 *__restrict__ sorted) {
    int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= natoms) return;
    double wx, wy, wz;
    wrap_to_cell(g, (double)xs[3 * a], (double)xs[3 * a + 1], (double)xs[3 * a + 2], wx, wy, wz);
    for (int n2 = -1; n2 <= 1; ++n2)
        for (int n1 = -1; n1 <= 1; ++n1)
            for (int n0 = -1; n0 <= 1; ++n0) {
                double rx = wx + g.m[0] * n0 + g.m[3] * n1 + g.m[6] * n2;
                double ry = wy + g.m[1] * n0 + g.m[4] * n1 + g.m[7] * n2;
                double rz = wz + g.m[2] * n0 + g.m[5] * n1 + g.m[8] * n2;
                if (rx < g.elo[0] || rx >= g.ehi[0] || ry < g.elo[1] || ry >= g.ehi[1] || rz < g.elo[2] || rz >= g.ehi[2])
                    continue;
                float px = (float)(rx - g.ctr[0]), py = (float)(ry - g.ctr[1]), pz = (float)(rz - g.ctr[2]);
                int c = fine_cell_of(g, px, py, pz);
                if (!SCATTER) {
                    atomicAdd(&cell_count[c], 1);
                    int fx = c % g.nx, frow = c / g.nx;
                    atomicOr(&rowmask[(size_t)frow * g.rw + (fx >> 6)], 1ull << (fx & 63));
                    int cx, cy, cz; coarse_cell_of(g, px, py, pz, cx, cy, cz);
                    atomicOr(&occ_bits[(size_t)(cz * g.ncy + cy) * g.cw + (cx >> 6)], 1ull << (cx & 63));
                } else {
                    int slot = cell_start[c] + atomicSub(&cell_count[c], 1) - 1;
                    sorted[slot] = make_float4(px, py, pz, __int_as_float(a));
                }
            }
}

// ---------------------------------------------------------------------------------------------
// K4: Chebyshev distance (in coarse cells, capped) from every coarse cell to the nearest occupied one
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t row_window(const u64 *row, int cw, int cx) {
    // 31-bit window of the row bitmap, bit 15 = column cx
    int lo = cx - 15;
    int wi = lo >= 0 ? (lo >> 6) : -1;
    int sh = lo - wi * 64;   // 0..63
    u64 w0 = (wi >= 0 && wi < cw) ? row[wi] : 0ull;
    u64 w1 = (wi + 1 >= 0 && wi + 1 < cw) ? row[wi + 1] : 0ull;
    u64 v = (w0 >> sh) | (sh ? (w1 << (64 - sh)) : 0ull);
    return (uint32_t)(v & 0x7fffffffull);
}

// Separable lower-bound distance transform on the cull grid.  For two points in cells that are
// (ax,ay,az) cells apart, every component of their separation is at least max(a-1,0) cells, so
//   lb^2 = cside^2 * min over occupied cells of  f(ax)+f(ay)+f(az),   f(a) = max(a-1,0)^2
// is a lower bound of the squared distance from ANY point of the cell to ANY solute atom.
// Pass X uses the row bitmaps (nearest set bit), passes Y and Z are min-plus scans over the window.
__device__ __forceinline__ int edt_f(int a) { int t = max(a - 1, 0); return t * t; }

__global__ void k_edt_x(Geom g, const u64 *__restrict__ occ_bits, unsigned char *__restrict__ ex) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    int ncc = g.ncx * g.ncy * g.ncz;
    if (c >= ncc) return;
    int cx = c % g.ncx, row = c / g.ncx;
    uint32_t w = row_window(occ_bits + (size_t)row * g.cw, g.cw, cx);
    uint32_t hi = w >> 15, lo = w & 0xffffu;
    int dr = hi ? (__ffs(hi) - 1) : 99;
    int dl = lo ? (__clz(lo) - 16) : 99;
    ex[c] = (unsigned char)min(min(dl, dr), g.dwin + 1);
}

__global__ void k_edt_y(Geom g, const unsigned char *__restrict__ ex, unsigned short *__restrict__ exy) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    int ncc = g.ncx * g.ncy * g.ncz;
    if (c >= ncc) return;
    int cx = c % g.ncx, cy = (c / g.ncx) % g.ncy, cz = c / (g.ncx * g.ncy);
    int D = g.dwin, best = 3 * D * D;
    for (int dy = -D; dy <= D; ++dy) {
        int ry = cy + dy;
        if (ry < 0 || ry >= g.ncy) continue;
        best = min(best, edt_f(ex[(cz * g.ncy + ry) * g.ncx + cx]) + edt_f(abs(dy)));
    }
    exy[c] = (unsigned short)best;
}

__global__ void k_edt_z(Geom g, const unsigned short *__restrict__ exy, float *__restrict__ lbd2) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    int ncc = g.ncx * g.ncy * g.ncz;
    if (c >= ncc) return;
    int cx = c % g.ncx, cy = (c / g.ncx) % g.ncy, cz = c / (g.ncx * g.ncy);
    int D = g.dwin, best = 3 * D * D;
    for (int dz = -D; dz <= D; ++dz) {
        int rz = cz + dz;
        if (rz < 0 || rz >= g.ncz) continue;
        best = min(best, (int)exy[(rz * g.ncy + cy) * g.ncx + cx] + edt_f(abs(dz)));
    }
    // anything at or beyond the window is "far": the window is sized so that D*cside exceeds every threshold
    lbd2[c] = best >= D * D ? CUDART_INF_F : (float)best * g.cside * g.cside;
}

__device__ __forceinline__ float cull_lb2(const Geom &g, const float *__restrict__ lbd2, float px, float py, float pz) {
    int cx, cy, cz; coarse_cell_of(g, px, py, pz, cx, cy, cz);
    return __ldg(&lbd2[(cz * g.ncy + cy) * g.ncx + cx]);
}

// ---------------------------------------------------------------------------------------------
// K5: cull the solvent molecules of the frame; survivors -> work list.  Also the largest
// centroid-to-atom distance of any molecule (bound used to cull random placements).
// ---------------------------------------------------------------------------------------------
__global__ void k_filter_real(Geom g, Prob P, const float *__restrict__ xv, int skip_mol,
                              const float *__restrict__ lbd2, MdRec *__restrict__ list,
                              int *__restrict__ worklist, int *__restrict__ work_count, int *__restrict__ rmax_bits) {
    int m = blockIdx.x * blockDim.x + threadIdx.x;
    bool near = false;
    float r2max = 0.f;
    if (m < P.nv_mols) {
        const float *x = xv + (size_t)3 * P.nv_apm * m;
        double rx = x[3 * P.iref], ry = x[3 * P.iref + 1], rz = x[3 * P.iref + 2];
        double sx = 0, sy = 0, sz = 0;
        for (int k = 0; k < P.nv_apm; ++k) {
            double px = x[3 * k], py = x[3 * k + 1], pz = x[3 * k + 2];
            double wx, wy, wz; wrap_to_cell(g, px, py, pz, wx, wy, wz);
            near |= cull_lb2(g, lbd2, (float)(wx - g.ctr[0]), (float)(wy - g.ctr[1]), (float)(wz - g.ctr[2])) <= g.cut_hi2;
            double dx = px - rx, dy = py - ry, dz = pz - rz;
            min_image64(g, dx, dy, dz);
            sx += dx; sy += dy; sz += dz;
        }
        sx /= P.nv_apm; sy /= P.nv_apm; sz /= P.nv_apm;
        for (int k = 0; k < P.nv_apm; ++k) {
            double dx = x[3 * k] - rx, dy = x[3 * k + 1] - ry, dz = x[3 * k + 2] - rz;
            min_image64(g, dx, dy, dz);
            dx -= sx; dy -= sy; dz -= sz;
            r2max = fmaxf(r2max, (float)(dx * dx + dy * dy + dz * dz));
        }
        if (m == skip_mol) near = false;   // autocorrelation: the solute molecule itself (minimum_distances.jl:90)
        MdRec e; e.d = CUDART_INF; e.dref = CUDART_INF; e.i = -1; e.j = -1; e.flags = 0; e.pad = 0;
        list[m] = e;
    }
    // warp-aggregated append
    unsigned ball = __ballot_sync(0xffffffffu, near);
    int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == 0 && ball) base = atomicAdd(work_count, __popc(ball));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (near) worklist[base + __popc(ball & ((1u << lane) - 1))] = m;
    // block max of the molecule radius (rounded up)
    float r = sqrtf(r2max) * 1.000001f + 1e-6f;
    for (int o = 16; o; o >>= 1) r = fmaxf(r, __shfl_xor_sync(0xffffffffu, r, o));
    if (lane == 0 && r > 0.f) atomicMax(rmax_bits, __float_as_int(r));
}

// ---------------------------------------------------------------------------------------------
// K6a: query positions of the work-list molecules.  Real phase: wrapped fp32 positions of the
// frame's atoms.  Random phase: the random molecules that survived the centre cull are generated
// (Philox + rigid move, fp64) and stored as exact fp64 + wrapped fp32 positions; culled placements
// are never materialised.
// ---------------------------------------------------------------------------------------------
// .w carries the lower bound (squared) of the distance from the atom to the solute
__device__ __forceinline__ float4 query_pos(const Geom &g, const float *__restrict__ lbd2, double ex, double ey, double ez) {
    double wx, wy, wz; wrap_to_cell(g, ex, ey, ez, wx, wy, wz);
    float px = (float)(wx - g.ctr[0]), py = (float)(wy - g.ctr[1]), pz = (float)(wz - g.ctr[2]);
    return make_float4(px, py, pz, cull_lb2(g, lbd2, px, py, pz));
}

__global__ void k_gen_real(Geom g, Prob P, const float *__restrict__ xv, const float *__restrict__ cdist,
                           const int *__restrict__ worklist, const int *__restrict__ work_count, float4 *__restrict__ qpos) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)(*work_count) * P.nv_apm;
    for (; t < total; t += (long long)gridDim.x * blockDim.x) {
        int w = (int)(t / P.nv_apm), k = (int)(t - (long long)w * P.nv_apm);
        const float *x = xv + ((size_t)worklist[w] * P.nv_apm + k) * 3;
        qpos[t] = query_pos(g, cdist, (double)x[0], (double)x[1], (double)x[2]);
    }
}

__global__ void __launch_bounds__(128)
k_gen_rand(Geom g, Prob P, uint32_t frame, const float *__restrict__ xv, const float *__restrict__ cdist,
           const int *__restrict__ worklist, const int *__restrict__ work_count, const int *__restrict__ bulk_idx,
           const int *__restrict__ n_bulk_ptr, float4 *__restrict__ qpos, double *__restrict__ xexact) {
    const int count = *work_count;
    for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < count; w += gridDim.x * blockDim.x) {
        int item = worklist[w];
        int sample = item / P.nv_mols, mol = item - sample * P.nv_mols;
        uint4 r0 = philox4x32((uint32_t)mol, (uint32_t)sample, frame, 0u, P.seed_lo, P.seed_hi);
        uint4 r1 = philox4x32((uint32_t)mol, (uint32_t)sample, frame, 1u, P.seed_lo, P.seed_hi);
        int nb = *n_bulk_ptr;
        int jmol = nb > 0 ? bulk_idx[pick(r0.x, (uint32_t)nb)] : (int)pick(r0.x, (uint32_t)P.nv_mols);
        RandMol rm; rm.init(g, xv + (size_t)3 * P.nv_apm * jmol, P.nv_apm, P.iref, r0, r1);
        for (int k = 0; k < P.nv_apm; ++k) {
            double ex, ey, ez; rm.get(g, k, ex, ey, ez);
            size_t o = (size_t)w * P.nv_apm + k;
            xexact[3 * o] = ex; xexact[3 * o + 1] = ey; xexact[3 * o + 2] = ez;
            qpos[o] = query_pos(g, cdist, ex, ey, ez);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K6: the search.  G lanes per molecule.
// ---------------------------------------------------------------------------------------------
struct LaneBest {
    float b1, b2;   // best and second-best squared distance seen by this lane
    int i, k;       // solute atom index and solvent-molecule atom of b1
};

template <int G>
__device__ __forceinline__ unsigned group_mask() {
    if constexpr (G == 32) return 0xffffffffu;
    else {
        int lane = threadIdx.x & 31;
        return ((1u << G) - 1u) << (lane & ~(G - 1));
    }
}
template <int G>
__device__ __forceinline__ float group_min(float v, unsigned mask) {
#pragma unroll
    for (int o = G / 2; o; o >>= 1) v = fminf(v, __shfl_xor_sync(mask, v, o));
    return v;
}

// Nearest solute atoms of one query atom.  The (dy,dz) rows of the fine grid are taken in batches
// of G in order of increasing distance: every lane probes ONE row of the batch -- occupancy
// bitmask (one 8-byte load tells which cells of the row hold atoms), then the cell_start pair of
// the occupied span -- so a batch costs two dependent load latencies instead of two per row.
// The occupied rows of the batch are then scanned by the whole group with coalesced 16-byte
// loads of the cell-sorted solute, nearest row first, shrinking the bound after every row.
template <int G, bool COUNT>
__device__ __forceinline__ void search_atom(const Geom &g, const u64 *__restrict__ rowmask,
                                            const int *__restrict__ cell_start, const float4 *__restrict__ sorted,
                                            float px, float py, float pz, int k, float &bound, LaneBest &lb,
                                            unsigned mask, int gl, unsigned long long &npairs) {
    const float slack = 2e-3f;
    int cy0 = (int)floorf((py - g.gmin[1]) * g.inv_side), cz0 = (int)floorf((pz - g.gmin[2]) * g.inv_side);
    cy0 = min(max(cy0, 0), g.ny - 1); cz0 = min(max(cz0, 0), g.nz - 1);
    const float side2 = g.side * g.side;
    const int gbase = (threadIdx.x & 31) & ~(G - 1);
    for (int t0 = 0; t0 < g.nrows_tab; t0 += G) {
        if (c_row_lb[t0] * side2 > bound) break;   // table sorted: every later row is farther
        int a = 0, b = 0; float rd2 = CUDART_INF_F;
        int t = t0 + gl;
        if (t < g.nrows_tab) {
            int ry = cy0 + c_row_dy[t], rz = cz0 + c_row_dz[t];
            if (ry >= 0 && ry < g.ny && rz >= 0 && rz < g.nz) {
                float y0 = g.gmin[1] + ry * g.side, z0 = g.gmin[2] + rz * g.side;
                float gy = fmaxf(fmaxf(y0 - py, py - (y0 + g.side)) - slack, 0.f);
                float gz = fmaxf(fmaxf(z0 - pz, pz - (z0 + g.side)) - slack, 0.f);
                float r2 = gy * gy + gz * gz;
                if (r2 <= bound) {
                    float hx = sqrtf(bound - r2) + slack;
                    int cxl = max((int)floorf((px - hx - g.gmin[0]) * g.inv_sidex), 0);
                    int cxh = min((int)floorf((px + hx - g.gmin[0]) * g.inv_sidex), g.nx - 1);
                    if (cxl <= cxh) {
                        int row = rz * g.ny + ry;
                        int w0 = cxl >> 6, w1 = cxh >> 6;
                        u64 m0 = __ldg(&rowmask[(size_t)row * g.rw + w0]) & (~0ull << (cxl & 63));
                        u64 m1 = 0;
                        if (w1 == w0) m0 &= (~0ull >> (63 - (cxh & 63)));
                        else m1 = __ldg(&rowmask[(size_t)row * g.rw + w1]) & (~0ull >> (63 - (cxh & 63)));
                        if (m0 | m1) {
                            int first = m0 ? (w0 << 6) + __ffsll((long long)m0) - 1 : (w1 << 6) + __ffsll((long long)m1) - 1;
                            int last = m1 ? (w1 << 6) + 63 - __clzll((long long)m1) : (w0 << 6) + 63 - __clzll((long long)m0);
                            a = __ldg(&cell_start[row * g.nx + first]);
                            b = __ldg(&cell_start[row * g.nx + last + 1]);
                            rd2 = r2;
                        }
                    }
                }
            }
        }
        unsigned ball = __ballot_sync(mask, b > a);
        unsigned sub = (G == 32) ? ball : ((ball >> gbase) & ((1u << (G & 31)) - 1u));
        while (sub) {
            int r = __ffs(sub) - 1;
            sub &= sub - 1;
            int ra = __shfl_sync(mask, a, gbase + r), rb = __shfl_sync(mask, b, gbase + r);
            float rr = __shfl_sync(mask, rd2, gbase + r);
            if (rr > bound) continue;
            for (int p = ra + gl; p < rb; p += G) {
                float4 s = __ldg(&sorted[p]);
                float dx = s.x - px, dy = s.y - py, dz = s.z - pz;
                float d2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
                if (d2 < lb.b1) { lb.b2 = lb.b1; lb.b1 = d2; lb.i = __float_as_int(s.w); lb.k = k; }
                else lb.b2 = fminf(lb.b2, d2);
            }
            if (COUNT) npairs += (unsigned long long)((rb - ra - gl + G - 1) / G);
            float gb = group_min<G>(lb.b1, mask);
            bound = fminf(bound, gb + g.tol_d2);
        }
    }
}

// Result of the fp32 search of one molecule, identical in all lanes of the group
struct Found {
    float b1, b2; int i, k;       // molecule: best, second best, winning pair
    float r1, r2; int ri;         // reference atom: best, second best, winning solute atom
};

template <int G>
__device__ __forceinline__ void group_combine(LaneBest lb, unsigned mask, int gl, float &b1, float &b2, int &bi, int &bk) {
    b1 = group_min<G>(lb.b1, mask);
    unsigned ball = __ballot_sync(mask, lb.b1 == b1);
    int lane = threadIdx.x & 31;
    int gbase = lane & ~(G - 1);
    unsigned sub = (G == 32) ? ball : ((ball >> gbase) & ((1u << (G & 31)) - 1u));
    int wl = __ffs(sub) - 1;   // winning lane within the group
    float cand = (gl == wl) ? lb.b2 : lb.b1;
    b2 = group_min<G>(cand, mask);
    bi = __shfl_sync(mask, lb.i, gbase + wl);
    bk = __shfl_sync(mask, lb.k, gbase + wl);
}

// classification of the fp32 result: 0 = certainly outside, 1 = certainly inside and unambiguous,
// 2 = ambiguous (needs the exact path)
__device__ __forceinline__ int classify(const Geom &g, float b1, float b2) {
    float d1 = sqrtf(b1);
    if (d1 > g.cut_hi) return 0;
    if (d1 >= g.cut_lo) return 2;
    if (sqrtf(b2) - d1 <= g.tau) return 2;
    return 1;
}

template <int G, bool RANDOM, bool COUNT>
__global__ void __launch_bounds__(256)
k_search(Geom g, Prob P, const float *__restrict__ xs /* solute molecule, fp32 as read */,
         const float *__restrict__ xv /* solvent of the frame, fp32 as read */,
         const int *__restrict__ cell_start, const float4 *__restrict__ sorted, const u64 *__restrict__ rowmask,
         const float4 *__restrict__ qpos, const double *__restrict__ xexact,
         const int *__restrict__ worklist, const int *__restrict__ work_count,
         MdRec *__restrict__ list /* real: [nv_mols]; random: debug [nrand][nv_mols] or NULL */,
         u64 *__restrict__ deferred, int *__restrict__ deferred_count, u64 *__restrict__ pair_evals) {
    const int gl = threadIdx.x & (G - 1);
    const unsigned mask = group_mask<G>();
    const int ngroups = (gridDim.x * blockDim.x) / G;
    const int count = *work_count;
    unsigned long long npairs = 0;
    for (int w = (blockIdx.x * blockDim.x + threadIdx.x) / G; w < count; w += ngroups) {
        const int item = worklist[w];
        const float4 *q = qpos + (size_t)w * P.nv_apm;
        LaneBest lb; lb.b1 = CUDART_INF_F; lb.b2 = CUDART_INF_F; lb.i = -1; lb.k = -1;
        Found F; F.r1 = CUDART_INF_F; F.r2 = CUDART_INF_F; F.ri = -1;
        float bound = g.search2;
        // reference atom first: its own nearest solute atom is needed exactly (rdf_count), and it
        // seeds the bound for the other atoms
        for (int kk = 0; kk < P.nv_apm; ++kk) {
            int k = kk == 0 ? P.iref : (kk <= P.iref ? kk - 1 : kk);
            float4 p = __ldg(&q[k]);
            if (p.w <= bound) search_atom<G, COUNT>(g, rowmask, cell_start, sorted, p.x, p.y, p.z, k, bound, lb, mask, gl, npairs);
            if (kk == 0) { int tk; group_combine<G>(lb, mask, gl, F.r1, F.r2, F.ri, tk); }
        }
        group_combine<G>(lb, mask, gl, F.b1, F.b2, F.i, F.k);
        int sample = 0, mol = item;
        if (RANDOM) { sample = item / P.nv_mols; mol = item - sample * P.nv_mols; }
        int cls = classify(g, F.b1, F.b2);
        if (cls == 0) continue;   // list entry stays "not within"
        int rcls = classify(g, F.r1, F.r2);
        // the reference atom only matters when the molecule is inside
        if (cls == 2 || rcls == 2) {
            if (gl == 0) {
                int slot = atomicAdd(deferred_count, 1);
                deferred[slot] = ((u64)(RANDOM ? 1 + sample : 0) << 32) | (u64)(uint32_t)mol;
            }
            continue;
        }
        // exact fp64 finalisation: lane 0 the winning pair, lane 1 (same instruction stream) the
        // reference-atom pair
        const bool second = (G > 1) && gl == 1;
        const int fi = second ? F.ri : F.i, fk = second ? P.iref : F.k;
        double dd = 0;
        if (gl == 0 || (second && rcls == 1)) {
            double ex, ey, ez;
            if (RANDOM) { const double *xe = xexact + ((size_t)P.nv_apm * w + fk) * 3; ex = xe[0]; ey = xe[1]; ez = xe[2]; }
            else { const float *xr = xv + ((size_t)P.nv_apm * mol + fk) * 3; ex = (double)xr[0]; ey = (double)xr[1]; ez = (double)xr[2]; }
            dd = dist_pbc64(g, (double)xs[3 * fi], (double)xs[3 * fi + 1], (double)xs[3 * fi + 2], ex, ey, ez);
        }
        double dref = CUDART_INF;
        if (G > 1) dref = __shfl_sync(mask, dd, ((threadIdx.x & 31) & ~(G - 1)) + 1);
        if (gl != 0) continue;
        if (G == 1 && rcls == 1) {
            double ex, ey, ez;
            if (RANDOM) { const double *xe = xexact + ((size_t)P.nv_apm * w + P.iref) * 3; ex = xe[0]; ey = xe[1]; ez = xe[2]; }
            else { const float *xr = xv + ((size_t)P.nv_apm * mol + P.iref) * 3; ex = (double)xr[0]; ey = (double)xr[1]; ez = (double)xr[2]; }
            dref = dist_pbc64(g, (double)xs[3 * F.ri], (double)xs[3 * F.ri + 1], (double)xs[3 * F.ri + 2], ex, ey, ez);
        }
        MdRec e;
        e.d = dd; e.i = F.i; e.j = mol * P.nv_apm + F.k; e.flags = 1; e.dref = CUDART_INF; e.pad = 0;
        if (rcls == 1) { e.dref = dref; e.flags |= 2; }
        count_hit(P, RANDOM, e.d, e.i, e.j, 1ull);
        if (e.flags & 2) count_ref(P, RANDOM, e.dref);
        if (list) list[RANDOM ? (size_t)sample * P.nv_mols + mol : (size_t)mol] = e;
    }
    if (COUNT && pair_evals) {
        for (int o = 16; o; o >>= 1) npairs += __shfl_xor_sync(0xffffffffu, npairs, o);
        if ((threadIdx.x & 31) == 0 && npairs) atomicAdd(pair_evals, npairs);
    }
}

// ---------------------------------------------------------------------------------------------
// K7: bulk flags (inbulk, src/mddf.jl:55-57,406-415); the ordered compaction is a cub::DeviceSelect
// ---------------------------------------------------------------------------------------------
__global__ void k_bulk_flags(Prob P, const MdRec *__restrict__ list, int skip_mol, unsigned char *__restrict__ flags) {
    int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= P.nv_mols) return;
    flags[m] = (m != skip_mol) && inbulk(P, list[m]);
}

// ---------------------------------------------------------------------------------------------
// K8: cull the random placements by the position of their centre
// ---------------------------------------------------------------------------------------------
__global__ void k_filter_rand(Geom g, Prob P, uint32_t frame, int isolute, int skip_mol,
                              const float *__restrict__ lbd2, const int *__restrict__ rmax_bits,
                              int *__restrict__ worklist, int *__restrict__ work_count) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)P.nrand * P.nv_mols;
    bool near = false;
    if (t < total) {
        int sample = (int)(t / P.nv_mols), mol = (int)(t - (long long)sample * P.nv_mols);
        bool mine = P.ns_mols == 1 || ref_solute_of_sample(P, frame, (uint32_t)sample) == isolute;
        if (mine && mol != skip_mol) {
            float rmax = __int_as_float(*rmax_bits);
            if (rmax > g.rmax_bound) near = true;   // transform window not valid for this radius: no culling
            else {
                uint4 r0 = philox4x32((uint32_t)mol, (uint32_t)sample, frame, 0u, P.seed_lo, P.seed_hi);
                double u0 = u01(r0.y), u1 = u01(r0.z), u2 = u01(r0.w);
                const double *m = g.m;
                double cx_ = m[0] * u0 + m[3] * u1 + m[6] * u2, cy_ = m[1] * u0 + m[4] * u1 + m[7] * u2,
                       cz_ = m[2] * u0 + m[5] * u1 + m[8] * u2;
                float lim = g.cut_hi + rmax + 1e-3f;
                near = cull_lb2(g, lbd2, (float)(cx_ - g.ctr[0]), (float)(cy_ - g.ctr[1]), (float)(cz_ - g.ctr[2])) <= lim * lim;
            }
        }
    }
    unsigned ball = __ballot_sync(0xffffffffu, near);
    int lane = threadIdx.x & 31, base = 0;
    if (lane == 0 && ball) base = atomicAdd(work_count, __popc(ball));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (near) worklist[base + __popc(ball & ((1u << lane) - 1))] = (int)t;
}

// ---------------------------------------------------------------------------------------------
// K9: exact resolve of the deferred molecules: fp64 brute force over the whole solute molecule
// with the oracle's rule "smallest (d, j, i) wins".  One block per deferred item.
// ---------------------------------------------------------------------------------------------
struct RealMolG {   // adapter: same interface as RandMol::get(g, k, ...)
    RealMol m;
    __device__ __forceinline__ void get(const Geom &, int k, double &x, double &y, double &z) const { m.get(k, x, y, z); }
};
struct ExactBest { double d, dref; int i, j; };
__device__ __forceinline__ bool better(double d, int j, int i, const ExactBest &b) {
    return d < b.d || (d == b.d && (j < b.j || (j == b.j && i < b.i)));
}

// fp32 bound first (block minimum over the cell-sorted solute incl. images), then the exact fp64
// evaluation only of the solute atoms that can matter: those within (min + 4 tau) of the molecule's
// nearest atom or of the reference atom's nearest atom.
template <class Mol>
__device__ __forceinline__ void resolve_one(const Geom &g, const Prob &P, const float *__restrict__ xs,
                                            const float4 *__restrict__ sorted, int nsorted, const Mol &mol,
                                            int molidx, bool random, MdRec *out, ExactBest *sh, float *shf) {
    float mb = CUDART_INF_F, rb = CUDART_INF_F;
    for (int k = 0; k < P.nv_apm; ++k) {
        double ex, ey, ez; mol.get(g, k, ex, ey, ez);
        double wx, wy, wz; wrap_to_cell(g, ex, ey, ez, wx, wy, wz);
        float px = (float)(wx - g.ctr[0]), py = (float)(wy - g.ctr[1]), pz = (float)(wz - g.ctr[2]);
#pragma unroll 4
        for (int p = threadIdx.x; p < nsorted; p += blockDim.x) {
            float4 s = __ldg(&sorted[p]);
            float dx = s.x - px, dy = s.y - py, dz = s.z - pz;
            float d2 = dx * dx + dy * dy + dz * dz;
            mb = fminf(mb, d2);
            if (k == P.iref) rb = fminf(rb, d2);
        }
    }
    shf[threadIdx.x] = mb; shf[blockDim.x + threadIdx.x] = rb;
    __syncthreads();
    for (int o = blockDim.x / 2; o; o >>= 1) {
        if (threadIdx.x < o) {
            shf[threadIdx.x] = fminf(shf[threadIdx.x], shf[threadIdx.x + o]);
            shf[blockDim.x + threadIdx.x] = fminf(shf[blockDim.x + threadIdx.x], shf[blockDim.x + threadIdx.x + o]);
        }
        __syncthreads();
    }
    float lim_m = sqrtf(shf[0]) + 4.f * g.tau, lim_r = sqrtf(shf[blockDim.x]) + 4.f * g.tau;
    lim_m *= lim_m; lim_r *= lim_r;
    __syncthreads();
    ExactBest b; b.d = CUDART_INF; b.dref = CUDART_INF; b.i = 0x7fffffff; b.j = 0x7fffffff;
    for (int k = 0; k < P.nv_apm; ++k) {
        double ex, ey, ez; mol.get(g, k, ex, ey, ez);
        double wx, wy, wz; wrap_to_cell(g, ex, ey, ez, wx, wy, wz);
        float px = (float)(wx - g.ctr[0]), py = (float)(wy - g.ctr[1]), pz = (float)(wz - g.ctr[2]);
        int j = molidx * P.nv_apm + k;
#pragma unroll 4
        for (int p = threadIdx.x; p < nsorted; p += blockDim.x) {
            float4 s = __ldg(&sorted[p]);
            float dx = s.x - px, dy = s.y - py, dz = s.z - pz;
            float d2 = dx * dx + dy * dy + dz * dz;
            if (!(d2 <= lim_m || (k == P.iref && d2 <= lim_r))) continue;
            int i = __float_as_int(s.w);
            double d = dist_pbc64(g, (double)xs[3 * i], (double)xs[3 * i + 1], (double)xs[3 * i + 2], ex, ey, ez);
            if (d <= g.cutd) {
                if (better(d, j, i, b)) { b.d = d; b.i = i; b.j = j; }
                if (k == P.iref && d < b.dref) b.dref = d;
            }
        }
    }
    sh[threadIdx.x] = b;
    __syncthreads();
    for (int o = blockDim.x / 2; o; o >>= 1) {
        if (threadIdx.x < o) {
            ExactBest a = sh[threadIdx.x], c = sh[threadIdx.x + o];
            double dref = fmin(a.dref, c.dref);
            if (better(c.d, c.j, c.i, a)) a = c;
            a.dref = dref;
            sh[threadIdx.x] = a;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        ExactBest r = sh[0];
        MdRec e; e.d = CUDART_INF; e.dref = CUDART_INF; e.i = -1; e.j = -1; e.flags = 0; e.pad = 0;
        if (r.d <= g.cutd) {
            e.d = r.d; e.i = r.i; e.j = r.j; e.flags = 1;
            if (r.dref <= g.cutd) { e.dref = r.dref; e.flags |= 2; }
            count_hit(P, random, e.d, e.i, e.j, 1ull);
            if (e.flags & 2) count_ref(P, random, e.dref);
        }
        if (out) *out = e;
    }
    __syncthreads();
}

#define CMX_RESOLVE_THREADS 256
__global__ void __launch_bounds__(CMX_RESOLVE_THREADS)
k_resolve(Geom g, Prob P, uint32_t frame, const float *__restrict__ xs, const float *__restrict__ xv,
          const float4 *__restrict__ sorted, const int *__restrict__ cell_start, int ncells, const int *__restrict__ bulk_idx, const int *__restrict__ n_bulk_ptr, const u64 *__restrict__ deferred,
          const int *__restrict__ deferred_count, MdRec *__restrict__ list, MdRec *__restrict__ rand_list) {
    __shared__ ExactBest sh[CMX_RESOLVE_THREADS];
    __shared__ float shf[2 * CMX_RESOLVE_THREADS];
    int count = *deferred_count;
    const int nsorted = cell_start[ncells];
    for (int w = blockIdx.x; w < count; w += gridDim.x) {
        u64 item = deferred[w];
        int phase = (int)(item >> 32), mol = (int)(item & 0xffffffffu);
        if (phase == 0) {
            RealMolG rl; rl.m.x = xv + (size_t)3 * P.nv_apm * mol;
            resolve_one(g, P, xs, sorted, nsorted, rl, mol, false, list ? &list[mol] : nullptr, sh, shf);
        } else {
            int sample = phase - 1;
            uint4 r0 = philox4x32((uint32_t)mol, (uint32_t)sample, frame, 0u, P.seed_lo, P.seed_hi);
            uint4 r1 = philox4x32((uint32_t)mol, (uint32_t)sample, frame, 1u, P.seed_lo, P.seed_hi);
            int nb = *n_bulk_ptr;
            int jmol = nb > 0 ? bulk_idx[pick(r0.x, (uint32_t)nb)] : (int)pick(r0.x, (uint32_t)P.nv_mols);
            RandMol rm; rm.init(g, xv + (size_t)3 * P.nv_apm * jmol, P.nv_apm, P.iref, r0, r1);
            resolve_one(g, P, xs, sorted, nsorted, rm, mol, true, rand_list ? &rand_list[(size_t)sample * P.nv_mols + mol] : nullptr, sh, shf);
        }
    }
}

}  // namespace cmx
