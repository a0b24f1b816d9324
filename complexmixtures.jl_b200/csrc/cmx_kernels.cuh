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
//   3. the atoms of the surviving molecules are binned into compact tiles of 32 query atoms; one
//      warp owns a tile, every lane one query, and all lanes walk the same solute atoms of the
//      occupied grid rows nearest-first (fp32, broadcast vector loads of the cell-sorted solute),
//      keeping best / second-best squared distances; a per-molecule kernel combines the atoms;
//   4. the winning pair is re-evaluated in fp64 with the reference's minimum-image arithmetic and
//      histogrammed; molecules whose fp32 result is ambiguous (near-tie, cutoff edge) are deferred to
//      an exact fp64 resolve kernel, so the counts equal the fp64 oracle's bit for bit.
//   The random phase generates a random molecule from its Philox counters only if its centre
//   survives the cull; the random box as a whole is never materialised.
//
// Launch structure: every kernel has a FRAME dimension (blockIdx.y = frame of the batch).  The host
// collects up to B frames, uploads B frame descriptors (GridFrame: geometry + the frame's scratch
// pointers) in one copy and launches each kernel ONCE for the whole batch, so the launch count per
// frame is (kernels per batch) / B.  Scans and the ordered bulk compaction are single-pass chained
// scans (decoupled look-back) written here: no library kernel is on the path.
#pragma once
#include "cmx_device.cuh"

namespace cmx {

// per-frame scalars (GridFrame::sc), zeroed by k_zero_frame
enum {
    SC_WORK = 0, SC_RWORK = 1, SC_DEF_REAL = 2, SC_DEF_RAND = 3, SC_NBULK = 4, SC_RMAX = 5, SC_TQ_REAL = 6, SC_TQ_RAND = 7,
    SC_SCAN_CELLS = 8, SC_SCAN_TILES_REAL = 9, SC_SCAN_TILES_RAND = 10, SC_SCAN_BULK = 11, SC_FQ_REAL = 12, SC_FQ_RAND = 13, SC_COUNT = 16
};

// One frame (x one solute molecule) in flight on the device: geometry, inputs and the scratch of its slot.
// An array of these (one per frame of the batch) lives in device memory; kernels pick theirs with blockIdx.y.
struct alignas(16) GridFrame {
    Geom g;
    const float *xs, *xv;          // the solute molecule, all solvent molecules (fp32 xyz as read)
    int *sc;                       // SC_COUNT scalars
    u64 *bits;                     // [cull bitmap][row bitmap], zeroed per frame
    u64 *occ, *rowmask;            // the two parts of `bits`
    int *cell_count, *cell_start;  // solute grid
    float4 *sorted;
    unsigned short *edt_xy;
    float *lbd2;
    float4 *qpos, *qsorted, *res;  // query atoms of the current phase: positions, tile order, results
    double *xexact;
    int *qcell_count, *qcell_start;
    unsigned char *tile_valid;     // valid lanes of every tile (a cell's last tile is partially filled)
    MdRec *list, *rand_list;       // real-phase list of this solute molecule; random lists only for the parity hooks
    int *worklist, *rand_worklist, *bulk_idx;
    u64 *def_real, *def_rand;
    float2 *def_real_info, *def_rand_info;
    u64 *scan_state;               // tile states of the chained scans
    long long bits_words;
    int ncells, nqcells, ncull;    // sizes of this frame's grids
    uint32_t frame;                // Philox frame key
    int isolute, skip_mol, nrand_k, pad0;
    double weight, pad1;           // frame weight (fp64 accumulation once the weights vary)
};
static_assert(sizeof(GridFrame) % 16 == 0, "GridFrame is copied with 16-byte loads");

// the block's frame descriptor -> shared memory (a few hundred bytes; every field is then a broadcast read)
#define CMX_FRAME(F)                                                                                         \
    __shared__ __align__(16) GridFrame F##_sh;                                                               \
    {                                                                                                        \
        const int4 *src_ = reinterpret_cast<const int4 *>(fds + blockIdx.y);                                 \
        int4 *dst_ = reinterpret_cast<int4 *>(&F##_sh);                                                      \
        for (int k_ = threadIdx.x; k_ < (int)(sizeof(GridFrame) / 16); k_ += blockDim.x) dst_[k_] = __ldg(src_ + k_); \
    }                                                                                                        \
    __syncthreads();                                                                                         \
    const GridFrame &F = F##_sh;                                                                             \
    const Geom &g = F.g;                                                                                     \
    (void)g;

// ---------------------------------------------------------------------------------------------
// per-frame reset: scalars and the two bitmaps (one launch per batch instead of a memset per frame)
// ---------------------------------------------------------------------------------------------
__global__ void k_zero_frame(const GridFrame *__restrict__ fds) {
    CMX_FRAME(F)
    const long long n = F.bits_words;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) F.bits[k] = 0ull;
    if (blockIdx.x == 0 && threadIdx.x < SC_COUNT) F.sc[threadIdx.x] = 0;
}

// between two chunks of random samples: the counters of the random phase
__global__ void k_reset_rand(const GridFrame *__restrict__ fds) {
    const GridFrame &F = fds[blockIdx.x];
    if (threadIdx.x == 0) { F.sc[SC_RWORK] = 0; F.sc[SC_DEF_RAND] = 0; F.sc[SC_TQ_RAND] = 0; F.sc[SC_SCAN_TILES_RAND] = 0; F.sc[SC_FQ_RAND] = 0; }
}

// ---------------------------------------------------------------------------------------------
// Solute grid: bin the solute molecule (plus periodic images inside the extended box) into the grid
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
__global__ void k_solute_bin(const GridFrame *__restrict__ fds, int natoms) {
    CMX_FRAME(F)
    int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= natoms) return;
    const float *xs = F.xs;
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
                    atomicAdd(&F.cell_count[c], 1);
                    int fx = c % g.nx, frow = c / g.nx;
                    atomicOr(&F.rowmask[(size_t)frow * g.rw + (fx >> 6)], 1ull << (fx & 63));
                    int cx, cy, cz; coarse_cell_of(g, px, py, pz, cx, cy, cz);
                    atomicOr(&F.occ[(size_t)(cz * g.ncy + cy) * g.cw + (cx >> 6)], 1ull << (cx & 63));
                } else {
                    int slot = F.cell_start[c] + atomicSub(&F.cell_count[c], 1) - 1;
                    F.sorted[slot] = make_float4(px, py, pz, __int_as_float(a));
                }
            }
}

// ---------------------------------------------------------------------------------------------
// Single-pass chained scan (decoupled look-back), one launch for arrays of any size and for every frame of the
// batch (blockIdx.y).  A tile is CMX_SCAN_TILE consecutive items; blocks take tiles in order from a per-frame
// counter (so a tile's predecessors are always running or done), publish their aggregate, look back over the
// predecessors' states for their exclusive prefix and publish the inclusive one.  State word = epoch (30 bits) |
// flag (2 bits: 1 aggregate, 2 inclusive prefix) | value (32 bits); the epoch is a per-launch number, so the state
// array is never reset.  The policy supplies the item values and consumes (item, value, exclusive prefix):
//   ScanCells      exclusive sum of the solute-grid cell counts -> cell_start
//   ScanTiles<R>   per query cell: tiles = ceil(count / 32) -> first tile of the cell, and the valid-lane count of
//                  each of its tiles (no sentinel fill of the tile array)
//   ScanBulk       inbulk (src/mddf.jl:55-57) of every solvent molecule -> bulk list in ascending molecule order
//                  (src/mddf.jl:406-415) + n_bulk: an ordered stream compaction
// ---------------------------------------------------------------------------------------------
#define CMX_SCAN_THREADS 256
#define CMX_SCAN_ITEMS 8
#define CMX_SCAN_TILE (CMX_SCAN_THREADS * CMX_SCAN_ITEMS)

struct ScanCells {
    static __device__ __forceinline__ int counter() { return SC_SCAN_CELLS; }
    static __device__ __forceinline__ int n(const GridFrame &F, const Prob &) { return F.ncells + 1; }
    static __device__ __forceinline__ int load(const GridFrame &F, const Prob &, int i) { return F.cell_count[i]; }
    static __device__ __forceinline__ void store(const GridFrame &F, const Prob &, int i, int, int excl) { F.cell_start[i] = excl; }
    static __device__ __forceinline__ void total(const GridFrame &, int) {}
};
template <bool RANDOM>
struct ScanTiles {
    static __device__ __forceinline__ int counter() { return RANDOM ? SC_SCAN_TILES_RAND : SC_SCAN_TILES_REAL; }
    static __device__ __forceinline__ int n(const GridFrame &F, const Prob &) { return F.nqcells + 1; }
    static __device__ __forceinline__ int load(const GridFrame &F, const Prob &, int i) { return (F.qcell_count[i] + 31) >> 5; }
    static __device__ __forceinline__ void store(const GridFrame &F, const Prob &, int i, int v, int excl) {
        F.qcell_start[i] = excl;
        if (v) {
            int cnt = F.qcell_count[i];
            for (int t = 0; t < v; ++t) F.tile_valid[excl + t] = (unsigned char)min(32, cnt - 32 * t);
        }
    }
    static __device__ __forceinline__ void total(const GridFrame &, int) {}
};
struct ScanBulk {
    static __device__ __forceinline__ int counter() { return SC_SCAN_BULK; }
    static __device__ __forceinline__ int n(const GridFrame &, const Prob &P) { return P.nv_mols; }
    static __device__ __forceinline__ int load(const GridFrame &F, const Prob &P, int m) {
        if (m == F.skip_mol) return 0;      // autocorrelation: the solute itself, src/mddf.jl:409
        return inbulk(P, F.list[m]) ? 1 : 0;
    }
    static __device__ __forceinline__ void store(const GridFrame &F, const Prob &, int m, int v, int excl) { if (v) F.bulk_idx[excl] = m; }
    static __device__ __forceinline__ void total(const GridFrame &F, int sum) { F.sc[SC_NBULK] = sum; }
};

__device__ __forceinline__ u64 scan_pack(uint32_t epoch, uint32_t flag, int value) {
    return ((u64)epoch << 34) | ((u64)flag << 32) | (u64)(uint32_t)value;
}

template <class Policy>
__global__ void __launch_bounds__(CMX_SCAN_THREADS)
k_chain_scan(const GridFrame *__restrict__ fds, Prob P, uint32_t epoch) {
    CMX_FRAME(F)
    __shared__ int s_tile, s_prefix, warp_sums[CMX_SCAN_THREADS / 32];
    const int n = Policy::n(F, P);
    const int ntiles = (n + CMX_SCAN_TILE - 1) / CMX_SCAN_TILE;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    volatile u64 *state = F.scan_state;
    while (true) {
        __syncthreads();     // s_tile / s_prefix / warp_sums of the previous tile are no longer read
        if (threadIdx.x == 0) s_tile = atomicAdd(&F.sc[Policy::counter()], 1);
        __syncthreads();
        const int tile = s_tile;
        if (tile >= ntiles) return;
        const int i0 = tile * CMX_SCAN_TILE + threadIdx.x * CMX_SCAN_ITEMS;
        int v[CMX_SCAN_ITEMS];
        int mine = 0;
#pragma unroll
        for (int k = 0; k < CMX_SCAN_ITEMS; ++k) { v[k] = (i0 + k < n) ? Policy::load(F, P, i0 + k) : 0; mine += v[k]; }
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) warp_sums[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            int ws = lane < CMX_SCAN_THREADS / 32 ? warp_sums[lane] : 0, wi = ws;
#pragma unroll
            for (int o = 1; o < CMX_SCAN_THREADS / 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += t; }
            if (lane < CMX_SCAN_THREADS / 32) warp_sums[lane] = wi - ws;      // exclusive prefix of the warp totals
            const int agg = __shfl_sync(0xffffffffu, wi, CMX_SCAN_THREADS / 32 - 1);
            int excl = 0;
            if (tile == 0) {
                if (lane == 0) state[0] = scan_pack(epoch, 2u, agg);
            } else {
                if (lane == 0) state[tile] = scan_pack(epoch, 1u, agg);
                int look = tile - 1;                                           // nearest predecessor handled by lane 0
                while (true) {
                    const int idx = look - lane;
                    u64 s = scan_pack(epoch, 2u, 0);                          // before the first tile: prefix 0
                    if (idx >= 0) {
                        do { s = state[idx]; } while ((uint32_t)(s >> 34) != epoch || ((s >> 32) & 3ull) == 0ull);
                    }
                    const unsigned full = __ballot_sync(0xffffffffu, ((s >> 32) & 3ull) == 2ull);
                    const int stop = full ? __ffs(full) - 1 : 32;            // first lane that holds an inclusive prefix
                    int val = (lane <= stop) ? (int)(uint32_t)s : 0;
#pragma unroll
                    for (int o = 16; o; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
                    excl += val;
                    if (full) break;
                    look -= 32;
                }
                if (lane == 0) state[tile] = scan_pack(epoch, 2u, excl + agg);
            }
            if (lane == 0) { s_prefix = excl; if (tile == ntiles - 1) Policy::total(F, excl + agg); }
        }
        __syncthreads();
        int run = s_prefix + warp_sums[wid] + (incl - mine);
#pragma unroll
        for (int k = 0; k < CMX_SCAN_ITEMS; ++k) { if (i0 + k < n) Policy::store(F, P, i0 + k, v[k], run); run += v[k]; }
    }
}

// ---------------------------------------------------------------------------------------------
// Cull grid: lower bound of the distance from every cull cell to the solute
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

__device__ __forceinline__ int edt_x_of(const Geom &g, const u64 *__restrict__ occ_bits, int row, int cx) {
    uint32_t w = row_window(occ_bits + (size_t)row * g.cw, g.cw, cx);
    uint32_t hi = w >> 15, lo = w & 0xffffu;
    int dr = hi ? (__ffs(hi) - 1) : 99;
    int dl = lo ? (__clz(lo) - 16) : 99;
    return min(min(dl, dr), g.dwin + 1);
}

// passes X and Y in one kernel.  A block owns a tile of 32 (x) by 8 (y) cells of one z layer: the X distance of the
// 8 + 2*dwin rows the tile's windows touch is computed ONCE from the row bitmaps (a handful of bit operations each)
// into shared memory, then every cell takes the minimum over its window of rows.
#define CMX_EDT_TX 32
#define CMX_EDT_TY 8
__global__ void __launch_bounds__(CMX_EDT_TX * CMX_EDT_TY)
k_edt_xy(const GridFrame *__restrict__ fds) {
    CMX_FRAME(F)
    __shared__ unsigned char ex[CMX_EDT_TY + 2 * 15][CMX_EDT_TX];      // dwin <= 15
    const int D = g.dwin;
    const int tx = threadIdx.x & (CMX_EDT_TX - 1), ty = threadIdx.x / CMX_EDT_TX;
    const int ntx = (g.ncx + CMX_EDT_TX - 1) / CMX_EDT_TX, nty = (g.ncy + CMX_EDT_TY - 1) / CMX_EDT_TY;
    const int ntiles = ntx * nty * g.ncz;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int cz = tile / (ntx * nty), rem = tile - cz * (ntx * nty);
        const int x0 = (rem % ntx) * CMX_EDT_TX, y0 = (rem / ntx) * CMX_EDT_TY;
        const int cx = x0 + tx;
        __syncthreads();
        for (int r = ty; r < CMX_EDT_TY + 2 * D; r += CMX_EDT_TY) {
            const int ry = y0 - D + r;
            int v = D + 1;                                                  // outside the grid: nothing there
            if (ry >= 0 && ry < g.ncy && cx < g.ncx) v = edt_x_of(g, F.occ, cz * g.ncy + ry, cx);
            ex[r][tx] = (unsigned char)v;
        }
        __syncthreads();
        const int cy = y0 + ty;
        if (cx < g.ncx && cy < g.ncy) {
            int best = 3 * D * D;
            for (int dy = -D; dy <= D; ++dy) best = min(best, edt_f((int)ex[ty + D + dy][tx]) + edt_f(abs(dy)));
            F.edt_xy[(cz * g.ncy + cy) * g.ncx + cx] = (unsigned short)best;
        }
    }
}

__global__ void k_edt_z(const GridFrame *__restrict__ fds) {
    CMX_FRAME(F)
    const int ncc = F.ncull;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < ncc; c += gridDim.x * blockDim.x) {
        int cx = c % g.ncx, cy = (c / g.ncx) % g.ncy, cz = c / (g.ncx * g.ncy);
        int D = g.dwin, best = 3 * D * D;
        for (int dz = -D; dz <= D; ++dz) {
            int rz = cz + dz;
            if (rz < 0 || rz >= g.ncz) continue;
            best = min(best, (int)F.edt_xy[(rz * g.ncy + cy) * g.ncx + cx] + edt_f(abs(dz)));
        }
        // anything at or beyond the window is "far": the window is sized so that D*cside exceeds every threshold
        F.lbd2[c] = best >= D * D ? CUDART_INF_F : (float)best * g.cside * g.cside;
    }
}

__device__ __forceinline__ float cull_lb2(const Geom &g, const float *__restrict__ lbd2, float px, float py, float pz) {
    int cx, cy, cz; coarse_cell_of(g, px, py, pz, cx, cy, cz);
    return __ldg(&lbd2[(cz * g.ncy + cy) * g.ncx + cx]);
}

// ---------------------------------------------------------------------------------------------
// Real-phase cull: the solvent molecules of the frame; survivors -> work list.  Also the largest
// centroid-to-atom distance of any molecule (bound used to cull random placements).
// ---------------------------------------------------------------------------------------------
__global__ void k_filter_real(const GridFrame *__restrict__ fds, Prob P) {
    CMX_FRAME(F)
    int m = blockIdx.x * blockDim.x + threadIdx.x;
    bool near = false;
    float r2max = 0.f;
    if (m < P.nv_mols) {
        const float *x = F.xv + (size_t)3 * P.nv_apm * m;
        const float rx = x[3 * P.iref], ry = x[3 * P.iref + 1], rz = x[3 * P.iref + 2];
        float sx = 0, sy = 0, sz = 0;
        for (int k = 0; k < P.nv_apm; ++k) {
            const float px = x[3 * k], py = x[3 * k + 1], pz = x[3 * k + 2];
            // cull by the distance map (fp32 wrap: its ~1e-4 A error is covered by the margin of the map's cells)
            // (an fp32 coordinate of magnitude X carries ~1e-7 X; the test is widened by that much: including a
            // molecule too many costs a search, never a count)
            float wx, wy, wz; wrap_to_cell32(g, px, py, pz, wx, wy, wz);
            const float lim = g.cut_hi + 1e-4f + 8e-7f * fmaxf(fmaxf(fabsf(px), fabsf(py)), fabsf(pz));
            near |= cull_lb2(g, F.lbd2, wx, wy, wz) <= lim * lim;
            // the molecule re-assembled about its reference atom (bounds only: fp32 differences of the fp32 coordinates)
            float dx = px - rx, dy = py - ry, dz = pz - rz;
            min_image32f(g, dx, dy, dz);
            sx += dx; sy += dy; sz += dz;
        }
        const float inv = 1.0f / (float)P.nv_apm;
        sx *= inv; sy *= inv; sz *= inv;
        for (int k = 0; k < P.nv_apm; ++k) {
            float dx = x[3 * k] - rx, dy = x[3 * k + 1] - ry, dz = x[3 * k + 2] - rz;
            min_image32f(g, dx, dy, dz);
            dx -= sx; dy -= sy; dz -= sz;
            r2max = fmaxf(r2max, dx * dx + dy * dy + dz * dz);
        }
        if (m == F.skip_mol) near = false;   // autocorrelation: the solute molecule itself (minimum_distances.jl:90)
        MdRec e; e.d = CUDART_INF; e.dref = CUDART_INF; e.i = -1; e.j = -1; e.flags = 0; e.pad = 0;
        F.list[m] = e;
    }
    // warp-aggregated append
    unsigned ball = __ballot_sync(0xffffffffu, near);
    int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == 0 && ball) base = atomicAdd(&F.sc[SC_WORK], __popc(ball));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (near) F.worklist[base + __popc(ball & ((1u << lane) - 1))] = m;
    // block max of the molecule radius (rounded up: fp32 coordinates of ~1e3 A carry ~1e-4 A)
    float r = sqrtf(r2max) * 1.00001f + 1e-3f;
    for (int o = 16; o; o >>= 1) r = fmaxf(r, __shfl_xor_sync(0xffffffffu, r, o));
    if (lane == 0 && r > 0.f) atomicMax(&F.sc[SC_RMAX], __float_as_int(r));
}

// ---------------------------------------------------------------------------------------------
// Query atoms of the work-list molecules.  Real phase: wrapped fp32 positions of the frame's
// atoms.  Random phase: the random molecules that survived the centre cull are generated (Philox +
// rigid move, fp64) and stored as exact fp64 + wrapped fp32 positions; culled placements are never
// materialised.  Every query atom that can be within the cutoff (distance-transform bound) is
// counted into a cubic "query cell" so that the search can work on spatially compact tiles.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int query_cell_of(const Geom &g, float px, float py, float pz) {
    int cx = min(max((int)floorf((px - g.gmin[0]) * g.inv_qside), 0), g.nqx - 1);
    int cy = min(max((int)floorf((py - g.gmin[1]) * g.inv_qside), 0), g.nqy - 1);
    int cz = min(max((int)floorf((pz - g.gmin[2]) * g.inv_qside), 0), g.nqz - 1);
    return (cz * g.nqy + cy) * g.nqx + cx;
}

// position (fp32, grid-relative) + lower bound of its squared distance to the solute; results reset
__device__ __forceinline__ void emit_query(const GridFrame &F, double ex, double ey, double ez, size_t qid) {
    const Geom &g = F.g;
    double wx, wy, wz; wrap_to_cell(g, ex, ey, ez, wx, wy, wz);
    float px = (float)(wx - g.ctr[0]), py = (float)(wy - g.ctr[1]), pz = (float)(wz - g.ctr[2]);
    float lb = cull_lb2(g, F.lbd2, px, py, pz);
    F.qpos[qid] = make_float4(px, py, pz, lb);
    F.res[qid] = make_float4(CUDART_INF_F, CUDART_INF_F, __int_as_float(-1), 0.f);
    if (lb <= g.cut_hi2) atomicAdd(&F.qcell_count[query_cell_of(g, px, py, pz)], 1);
}

__global__ void k_gen_real(const GridFrame *__restrict__ fds, Prob P) {
    CMX_FRAME(F)
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)F.sc[SC_WORK] * P.nv_apm;
    for (; t < total; t += (long long)gridDim.x * blockDim.x) {
        int w = (int)(t / P.nv_apm), k = (int)(t - (long long)w * P.nv_apm);
        const float *x = F.xv + ((size_t)F.worklist[w] * P.nv_apm + k) * 3;
        emit_query(F, (double)x[0], (double)x[1], (double)x[2], (size_t)t);
    }
}

#ifndef CMX_GEN_MINBLOCKS
#define CMX_GEN_MINBLOCKS 1
#endif
__global__ void __launch_bounds__(128, CMX_GEN_MINBLOCKS)
k_gen_rand(const GridFrame *__restrict__ fds, Prob P, int s0) {
    CMX_FRAME(F)
    const int count = F.sc[SC_RWORK];
    const int nb = F.sc[SC_NBULK];
    for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < count; w += gridDim.x * blockDim.x) {
        int item = F.rand_worklist[w];
        int sl = item / P.nv_mols, mol = item - sl * P.nv_mols;
        int sample = s0 + sl;
        uint4 r0 = philox4x32((uint32_t)mol, (uint32_t)sample, F.frame, 0u, P.seed_lo, P.seed_hi);
        uint4 r1 = philox4x32((uint32_t)mol, (uint32_t)sample, F.frame, 1u, P.seed_lo, P.seed_hi);
        int jmol = nb > 0 ? F.bulk_idx[pick(r0.x, (uint32_t)nb)] : (int)pick(r0.x, (uint32_t)P.nv_mols);
        RandMol rm; rm.init(g, F.xv + (size_t)3 * P.nv_apm * jmol, P.nv_apm, P.iref, r0, r1);
        for (int k = 0; k < P.nv_apm; ++k) {
            double ex, ey, ez; rm.get(g, k, ex, ey, ez);
            size_t o = (size_t)w * P.nv_apm + k;
            F.xexact[3 * o] = ex; F.xexact[3 * o + 1] = ey; F.xexact[3 * o + 2] = ez;
            emit_query(F, ex, ey, ez, o);
        }
    }
}

// scatter the counted query atoms into their cell's tiles: qsorted[tile*32 + rank] = {x, y, z, query id}; the
// ranks of a cell with n atoms are 0..n-1, so the valid lanes of its tiles are the ones k_chain_scan<ScanTiles>
// recorded in tile_valid -- the unused lanes of a cell's last tile are never written nor read
template <bool RANDOM>
__global__ void k_qscatter(const GridFrame *__restrict__ fds, Prob P) {
    CMX_FRAME(F)
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)F.sc[RANDOM ? SC_RWORK : SC_WORK] * P.nv_apm;
    for (; t < total; t += (long long)gridDim.x * blockDim.x) {
        float4 q = F.qpos[t];
        if (q.w <= g.cut_hi2) {
            int c = query_cell_of(g, q.x, q.y, q.z);
            size_t slot = (size_t)F.qcell_start[c] * 32 + (size_t)(atomicSub(&F.qcell_count[c], 1) - 1);
            F.qsorted[slot] = make_float4(q.x, q.y, q.z, __int_as_float((int)t));
        }
    }
}

// ---------------------------------------------------------------------------------------------
// The search.  One warp per tile of 32 query atoms that are neighbours in space (consecutive in
// query-cell order); every lane owns one query.  Rows of the solute grid ((y,z) columns of cells)
// inside the tile's reach are probed in lane-parallel (occupancy bitmask -> lower bound rd of the
// squared distance from the tile to the row).  The rows are then consumed in RINGS of increasing
// distance: all rows with rd <= (sqrt(nearest remaining rd) + ring)^2 are taken at once; each lane
// works out the cell range of ITS rows (no warp-serial per-row bookkeeping), a warp prefix sum lays the
// ranges out back to back, the lanes copy their ranges into the warp's staging buffer in shared
// memory, and ALL lanes then sweep the staged atoms (broadcast reads) in one uninterrupted loop.
// The loop handles two solute atoms per step with the packed fp32 instructions of sm_100
// (FADD2 / FMUL2 / FFMA2, staged atoms are stored pair-transposed so that both operands are register
// pairs) and 3-input min/max; each lane keeps best / second-best squared distance and the solute atom
// of the best: res[query] = {b1, b2, atom}.  After every ring the tile bound shrinks to the largest
// best-distance of its lanes, which ends the walk as soon as no remaining row can matter.
// The tiles of ALL frames of the batch form one queue (frames differ in how many tiles survive the
// cull): a warp takes the next (frame, tile) with one atomic -- issued one tile ahead, so its latency
// is hidden -- and the batch is balanced as a whole.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int fkey(float x) { int i = __float_as_int(x); return i ^ ((i >> 31) & 0x7fffffff); }
__device__ __forceinline__ float fkey_inv(int k) { return __int_as_float(k ^ ((k >> 31) & 0x7fffffff)); }
__device__ __forceinline__ float warp_minf(float x) { return fkey_inv(__reduce_min_sync(0xffffffffu, fkey(x))); }
__device__ __forceinline__ float warp_maxf(float x) { return fkey_inv(__reduce_max_sync(0xffffffffu, fkey(x))); }

// what the search needs of a frame, in shared memory for every frame of the batch
struct SearchFrame {
    const int *cell_start; const float4 *sorted; const u64 *rowmask; const float4 *qsorted; const unsigned char *tile_valid;
    float4 *res;
    float gmin[3], side, inv_side, sidex, inv_sidex, search2, tol_d2, ring;
    int nx, ny, nz, rw, tile_end;     // tile_end: end of this frame's range in the batch-wide tile numbering
};
#define CMX_MAX_BATCH 32
#define CMX_ROWS_PER_LANE 4
#define CMX_SEARCH_WARPS 8
#ifndef CMX_STAGE
#define CMX_STAGE 256                 // solute atoms staged per warp and sweep (pair-transposed: 16 B per atom)
// (measured, C4 / C2 frames/s: 4 blocks/SM x 256 staged x unroll 4: 3464 / 10452; unroll 2: 3395 / 10231; 3 blocks x 384: 3198 / 9685;
//  2 blocks x 512: 3119 / 9186 -- occupancy beats staging depth)
#endif
#ifndef CMX_SEARCH_MINBLOCKS
#define CMX_SEARCH_MINBLOCKS 4
#endif
#ifndef CMX_SWEEP_UNROLL
#define CMX_SWEEP_UNROLL 4
#endif
#ifndef CMX_XRING_REAL
#define CMX_XRING_REAL 0              // 1: the real-phase search uses the x-limited rings too (measured: C4 93 -> 105 us, C2 159 -> 188 us per batch)
#endif
#ifndef CMX_XRING
#define CMX_XRING 1                   // rings limited along x too (shells), rows taken span by span; 0 = rows swept over the full reach at once
#endif
constexpr int kSweepUnroll = CMX_SWEEP_UNROLL;
#ifndef CMX_TILE_TMA
#define CMX_TILE_TMA 0                // 1: the next tile's 32 queries arrive by a 1-D bulk async copy (cp.async.bulk + mbarrier: UBLKCP / SYNCS in the
                                      // SASS) while this tile is searched.  Measured, C4 / C2 frames/s: off 4018 / 10214, on 3759 / 9902 (kernel alone
                                      // 112.5 vs 116.1 us per C4 frame): the resident warps already hide the 512-byte load, the lane-0 issue path costs more
#endif
#define CMX_SEG_MAX (32 * CMX_ROWS_PER_LANE)
#define CMX_WARP_SMEM_BASE (CMX_STAGE * 16 + CMX_SEG_MAX * 4 + CMX_STAGE + 16)   // staged atoms, segment table, owner marks, carry
#if CMX_TILE_TMA
#define CMX_WARP_SMEM (CMX_WARP_SMEM_BASE + 2 * 512 + 16)                         // + two query tiles (double buffer) + two mbarriers
#else
#define CMX_WARP_SMEM CMX_WARP_SMEM_BASE
#endif
static_assert(CMX_WARP_SMEM % 16 == 0 && CMX_WARP_SMEM_BASE % 16 == 0, "per-warp shared memory is carved in 16-byte units (bulk copies)");

// ---- 1-D bulk async copy global -> shared, completion on an mbarrier (sm_90+: SASS UBLKCP / SYNCS) -------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void bulk_load(void *dst_smem, const void *src_gmem, uint32_t bytes, void *bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(void *bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
                 ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
#define CMX_SEARCH_SMEM (CMX_SEARCH_WARPS * CMX_WARP_SMEM)

template <bool COUNT, bool RANDOM>
__global__ void __launch_bounds__(CMX_SEARCH_WARPS * 32, CMX_SEARCH_MINBLOCKS)
k_tile_search(const GridFrame *__restrict__ fds, int nframes, u64 *__restrict__ pair_evals) {
    __shared__ SearchFrame sf[CMX_MAX_BATCH];
    // per warp: CMX_STAGE atoms as pairs {x0,x1,y0,y1},{z0,z1,w0,w1}
    extern __shared__ __align__(16) unsigned char search_smem[];      // per warp: CMX_WARP_SMEM bytes (dynamic)
    const int lane = threadIdx.x & 31;
    if (threadIdx.x < nframes) {
        const GridFrame &F = fds[threadIdx.x];
        SearchFrame s;
        s.cell_start = F.cell_start; s.sorted = F.sorted; s.rowmask = F.rowmask; s.qsorted = F.qsorted; s.tile_valid = F.tile_valid; s.res = F.res;
        s.gmin[0] = F.g.gmin[0]; s.gmin[1] = F.g.gmin[1]; s.gmin[2] = F.g.gmin[2];
        s.side = F.g.side; s.inv_side = F.g.inv_side; s.sidex = F.g.sidex; s.inv_sidex = F.g.inv_sidex; s.search2 = F.g.search2; s.tol_d2 = F.g.tol_d2; s.ring = F.g.ring;
        s.nx = F.g.nx; s.ny = F.g.ny; s.nz = F.g.nz; s.rw = F.g.rw;
        s.tile_end = F.qcell_start[F.nqcells];    // exclusive scan of the per-cell tile counts: total tiles of the frame
        sf[threadIdx.x] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) { int run = 0; for (int k = 0; k < nframes; ++k) { run += sf[k].tile_end; sf[k].tile_end = run; } }
    __syncthreads();
    const int ntiles_all = sf[nframes - 1].tile_end;
    int *tile_queue = fds[0].sc + (RANDOM ? SC_TQ_RAND : SC_TQ_REAL);
    const float slack = 2e-3f;
    unsigned long long npairs = 0;
    unsigned char *wsm = search_smem + (size_t)(threadIdx.x >> 5) * CMX_WARP_SMEM;
    float *st = reinterpret_cast<float *>(wsm);                                   // CMX_STAGE atoms, pair-transposed
    int *seg_src = reinterpret_cast<int *>(wsm + CMX_STAGE * 16);                 // per segment: (first cell-sorted atom) - (position in the ring's flattened order)
    unsigned char *owner = wsm + CMX_STAGE * 16 + CMX_SEG_MAX * 4;                // per staged position: 1 + segment that STARTS there, else 0
    int *carry = reinterpret_cast<int *>(wsm + CMX_STAGE * 16 + CMX_SEG_MAX * 4 + CMX_STAGE);
#if CMX_TILE_TMA
    // Tickets run TWO tiles ahead and the query tile ONE ahead: while tile i is searched, the 512 bytes of tile i+1 are
    // on their way into the warp's other buffer (bulk async copy, completion counted on an mbarrier) together with its
    // valid-lane count, and the atomic for ticket i+2 is in flight -- the warp never waits for a global round trip between tiles.
    float4 *qbuf = reinterpret_cast<float4 *>(wsm + CMX_WARP_SMEM_BASE);            // [2][32]
    unsigned long long *mbar = reinterpret_cast<unsigned long long *>(wsm + CMX_WARP_SMEM_BASE + 1024);
    if (lane == 0) { mbar_init(&mbar[0], 1); mbar_init(&mbar[1], 1); mbar_init_fence(); }
    __syncwarp();
    int t_cur = 0, t_nxt = 0, nv_cur = 0, fi_cur = 0, tile_cur = 0;
    auto locate = [&](int gt, int &fi, int &tile) { fi = 0; while (gt >= sf[fi].tile_end) ++fi; tile = gt - (fi ? sf[fi - 1].tile_end : 0); };
    if (lane == 0) {
        t_cur = atomicAdd(tile_queue, 1); t_nxt = atomicAdd(tile_queue, 1);
        if (t_cur < ntiles_all) {
            locate(t_cur, fi_cur, tile_cur);
            bulk_load(qbuf, sf[fi_cur].qsorted + (size_t)tile_cur * 32, 512u, &mbar[0]);
            nv_cur = sf[fi_cur].tile_valid[tile_cur];
        }
    }
    for (int it = 0;; ++it) {
        const int gt = __shfl_sync(0xffffffffu, t_cur, 0);
        if (gt >= ntiles_all) break;
        const int fi = __shfl_sync(0xffffffffu, fi_cur, 0), nvalid = __shfl_sync(0xffffffffu, nv_cur, 0);
        const int buf = it & 1;
        if (lane == 0) {        // (the buffer of tile it+1 was last read at the start of tile it-1: free)
            t_cur = t_nxt;
            t_nxt = atomicAdd(tile_queue, 1);
            if (t_cur < ntiles_all) {
                locate(t_cur, fi_cur, tile_cur);
                bulk_load(qbuf + 32 * (buf ^ 1), sf[fi_cur].qsorted + (size_t)tile_cur * 32, 512u, &mbar[buf ^ 1]);
                nv_cur = sf[fi_cur].tile_valid[tile_cur];
            }
        }
        const SearchFrame &S = sf[fi];
        const bool valid = lane < nvalid;
        mbar_wait(&mbar[buf], (uint32_t)((it >> 1) & 1));
        float4 q = qbuf[32 * buf + (valid ? lane : 0)];                           // unused lanes shadow the tile's first query
        __syncwarp();
#else
    int gt_next = 0;
    if (lane == 0) gt_next = atomicAdd(tile_queue, 1);
    while (true) {
        const int gt = __shfl_sync(0xffffffffu, gt_next, 0);
        if (gt >= ntiles_all) break;
        if (lane == 0) gt_next = atomicAdd(tile_queue, 1);      // the next tile's ticket travels while this tile is searched
        int fi = 0;
        while (gt >= sf[fi].tile_end) ++fi;
        const SearchFrame &S = sf[fi];
        const int tile = gt - (fi ? sf[fi - 1].tile_end : 0);
        const int nvalid = S.tile_valid[tile];
        const bool valid = lane < nvalid;
        float4 q = __ldg(&S.qsorted[(size_t)tile * 32 + (valid ? lane : 0)]);   // unused lanes shadow the tile's first query
#endif
        const float xmin = warp_minf(q.x), xmax = warp_maxf(q.x), ymin = warp_minf(q.y), ymax = warp_maxf(q.y),
                    zmin = warp_minf(q.z), zmax = warp_maxf(q.z);
        const float2 nqx = make_float2(-q.x, -q.x), nqy = make_float2(-q.y, -q.y), nqz = make_float2(-q.z, -q.z);
        float b1 = CUDART_INF_F, b2 = CUDART_INF_F; int bi = -1;
        float bound = S.search2;
        const float gmin0 = S.gmin[0], gmin1 = S.gmin[1], gmin2 = S.gmin[2], side = S.side, inv_side = S.inv_side, inv_sidex = S.inv_sidex;
#if CMX_XRING
        const float sidex = S.sidex;
#endif
        const int nx = S.nx, ny = S.ny, nz = S.nz, rw = S.rw;
        const float reach = sqrtf(bound) + slack;
        const int ry_lo = max((int)floorf((ymin - reach - gmin1) * inv_side), 0);
        const int ry_hi = min((int)floorf((ymax + reach - gmin1) * inv_side), ny - 1);
        const int rz_lo = max((int)floorf((zmin - reach - gmin2) * inv_side), 0);
        const int rz_hi = min((int)floorf((zmax + reach - gmin2) * inv_side), nz - 1);
        const int nry = ry_hi - ry_lo + 1, nrows = nry * (rz_hi - rz_lo + 1);
        const int *__restrict__ cell_start = S.cell_start;
        const float4 *__restrict__ sorted = S.sorted;
        for (int chunk = 0; chunk < nrows; chunk += 32 * CMX_ROWS_PER_LANE) {
            // ---- probe: lane handles rows chunk + u*32 + lane
            float rd[CMX_ROWS_PER_LANE];
            int rrow[CMX_ROWS_PER_LANE];          // (rz * ny + ry) of the probed row
#if CMX_XRING
            unsigned cons[CMX_ROWS_PER_LANE];     // cells of the row already swept: [lo 16 bits, hi 16 bits]; lo > hi = none yet
#endif
#pragma unroll
            for (int u = 0; u < CMX_ROWS_PER_LANE; ++u) {
                rd[u] = CUDART_INF_F; rrow[u] = 0;
#if CMX_XRING
                cons[u] = 0x0000ffffu;
#endif
                int r = chunk + u * 32 + lane;
                if (r < nrows) {
                    int rzq = r / nry;
                    int ry = ry_lo + (r - rzq * nry), rz = rz_lo + rzq;
                    rrow[u] = rz * ny + ry;
                    float y0 = gmin1 + ry * side, z0 = gmin2 + rz * side;
                    float gy = fmaxf(fmaxf(y0 - ymax, ymin - (y0 + side)) - slack, 0.f);
                    float gz = fmaxf(fmaxf(z0 - zmax, zmin - (z0 + side)) - slack, 0.f);
                    float r2 = gy * gy + gz * gz;
                    if (r2 <= bound) {
                        float hx = sqrtf(bound - r2) + slack;
                        int cxl = max((int)floorf((xmin - hx - gmin0) * inv_sidex), 0);
                        int cxh = min((int)floorf((xmax + hx - gmin0) * inv_sidex), nx - 1);
                        const u64 *mrow = S.rowmask + (size_t)(rz * ny + ry) * rw;
                        bool any = false;
                        for (int w = cxl >> 6; w <= (cxh >> 6) && cxl <= cxh; ++w) {
                            u64 m = __ldg(&mrow[w]);
                            if (w == (cxl >> 6)) m &= (~0ull << (cxl & 63));
                            if (w == (cxh >> 6)) m &= (~0ull >> (63 - (cxh & 63)));
                            any |= m != 0ull;
                        }
                        if (any) rd[u] = r2;
                    }
                }
            }
            // ---- consume the occupied rows in rings of increasing distance
#if CMX_XRING
            // A ring is a SHELL around the tile in all three directions: a row within reach of the ring is swept only over
            // the x-span the ring's radius allows (hx = sqrt(lim - rd), not sqrt(bound - rd)); what is left of it -- the
            // cells to the left and to the right of the swept span -- stays in play with the lower bound
            // key = rd + (swept half-width)^2 and is taken by a later ring, if the shrinking tile bound still reaches it.
            // (Before: the first ring swept its rows over the full reach of the INITIAL bound, +-cutoff along x, although
            // the bound of a tile inside or next to the solute drops to a few A^2 after that ring.)
            // (XR = false, the real phase: its tiles lie outside the solute, where the shells save few pair evaluations and
            // their bookkeeping costs more -- a row is swept over the full reach of the tile bound the first time it is taken)
            constexpr bool XR = RANDOM || CMX_XRING_REAL;
            while (true) {
                float key[CMX_ROWS_PER_LANE];
                float m = CUDART_INF_F;
#pragma unroll
                for (int u = 0; u < CMX_ROWS_PER_LANE; ++u) {
                    const int cl = (int)(cons[u] & 0xffffu), ch = (int)(cons[u] >> 16);
                    float k = rd[u];
                    if (XR && cl <= ch) {
                        const float l = cl == 0 ? CUDART_INF_F : xmin - (gmin0 + cl * sidex);
                        const float r = ch == nx - 1 ? CUDART_INF_F : (gmin0 + (ch + 1) * sidex) - xmax;
                        const float hw = fmaxf(fminf(l, r) - slack, 0.f);
                        k = rd[u] + hw * hw;              // (rd = inf stays inf; both sides at the grid edge -> inf)
                    }
                    key[u] = k;
                    m = fminf(m, k);
                }
                const float wm = __uint_as_float(__reduce_min_sync(0xffffffffu, __float_as_uint(m)));   // keys >= 0: bit order == value order
                if (!(wm <= bound)) break;
                const float rt = sqrtf(wm) + S.ring;
                const float lim = fminf(rt * rt, bound);
                const bool last = lim >= bound;           // this ring reaches as far as the tile bound: its rows are finished
              for (int sd = XR ? 0 : 1; sd < 2; ++sd) {   // sd 0: the spans to the right of what was swept; sd 1: to the left (or the whole span)
                bool mine = false;
#pragma unroll
                for (int u = 0; u < CMX_ROWS_PER_LANE; ++u) mine |= key[u] <= lim && (sd == 1 || (cons[u] & 0xffffu) <= (cons[u] >> 16));
                if (!__any_sync(0xffffffffu, mine)) continue;
                int aa[CMX_ROWS_PER_LANE], na[CMX_ROWS_PER_LANE];
                int mytotal = 0;
#pragma unroll
                for (int u = 0; u < CMX_ROWS_PER_LANE; ++u) {
                    aa[u] = 0; na[u] = 0;
                    if (key[u] <= lim) {
                        const float hx = sqrtf(fmaxf((XR ? lim : bound) - rd[u], 0.f)) + slack;
                        const int cxl = max((int)floorf((xmin - hx - gmin0) * inv_sidex), 0);
                        const int cxh = min((int)floorf((xmax + hx - gmin0) * inv_sidex), nx - 1);
                        int cl = (int)(cons[u] & 0xffffu), ch = (int)(cons[u] >> 16);
                        const bool none = cl > ch;
                        int s0 = 0, s1 = -1;              // cells [s0, s1] of this pass
                        if (sd == 0) { if (!none && cxh > ch) { s0 = ch + 1; s1 = cxh; ch = cxh; } }
                        else if (none) { s0 = cxl; s1 = cxh; if (cxl <= cxh) { cl = cxl; ch = cxh; } }
                        else if (cxl < cl) { s0 = cxl; s1 = cl - 1; cl = cxl; }
                        if (s0 <= s1) {
                            const int rowbase = rrow[u] * nx;
                            aa[u] = __ldg(&cell_start[rowbase + s0]);
                            na[u] = __ldg(&cell_start[rowbase + s1 + 1]) - aa[u];
                        }
                        cons[u] = (unsigned)cl | ((unsigned)ch << 16);
                        if (sd == 1 && (!XR || last || (cl == 0 && ch == nx - 1) || cxl > cxh)) rd[u] = CUDART_INF_F;
                        mytotal += na[u];
                    }
                }
#else
            while (true) {
                float m = rd[0];
#pragma unroll
                for (int u = 1; u < CMX_ROWS_PER_LANE; ++u) m = fminf(m, rd[u]);
                const float wm = __uint_as_float(__reduce_min_sync(0xffffffffu, __float_as_uint(m)));   // rd >= 0: bit order == value order
                if (!(wm <= bound)) break;
                const float rt = sqrtf(wm) + S.ring;
                const float lim = fminf(rt * rt, bound);
              {
                // my rows of this ring -> cell ranges [aa, aa + na)
                int aa[CMX_ROWS_PER_LANE], na[CMX_ROWS_PER_LANE];
                int mytotal = 0;
#pragma unroll
                for (int u = 0; u < CMX_ROWS_PER_LANE; ++u) {
                    aa[u] = 0; na[u] = 0;
                    if (rd[u] <= lim) {
                        float hx = sqrtf(bound - rd[u]) + slack;
                        int cxl = max((int)floorf((xmin - hx - gmin0) * inv_sidex), 0);
                        int cxh = min((int)floorf((xmax + hx - gmin0) * inv_sidex), nx - 1);
                        if (cxl <= cxh) {
                            int rowbase = rrow[u] * nx;
                            aa[u] = __ldg(&cell_start[rowbase + cxl]);
                            na[u] = __ldg(&cell_start[rowbase + cxh + 1]) - aa[u];
                        }
                        rd[u] = CUDART_INF_F;
                        mytotal += na[u];
                    }
                }
#endif
                int incl = mytotal;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
                const int total = __shfl_sync(0xffffffffu, incl, 31);
                const int myfirst = incl - mytotal;
                if (COUNT) npairs += (unsigned long long)total;
                // segment table: my ranges' positions in the flattened order of the ring
                int pu[CMX_ROWS_PER_LANE];
                {
                    int p = myfirst;
#pragma unroll
                    for (int u = 0; u < CMX_ROWS_PER_LANE; ++u) {
                        pu[u] = p;
                        if (na[u] > 0) seg_src[lane * CMX_ROWS_PER_LANE + u] = aa[u] - p;
                        p += na[u];
                    }
                }
                for (int base = 0; base < total; base += CMX_STAGE) {
                    const int cnt = min(CMX_STAGE, total - base);
                    __syncwarp();
                    // ---- stage [base, base + cnt) of the flattened order, ALL lanes copying (element base + c0 + lane):
                    // the lanes that own ranges mark where they start; a max-scan over the marks tells every element its range
                    for (int e = lane; e < cnt; e += 32) owner[e] = 0;
                    if (lane == 0) *carry = 0;
                    __syncwarp();
#pragma unroll
                    for (int u = 0; u < CMX_ROWS_PER_LANE; ++u)
                        if (na[u] > 0) {
                            if (pu[u] >= base && pu[u] < base + cnt) owner[pu[u] - base] = (unsigned char)(lane * CMX_ROWS_PER_LANE + u + 1);
                            else if (pu[u] < base && pu[u] + na[u] > base) *carry = lane * CMX_ROWS_PER_LANE + u + 1;   // straddles the window start
                        }
                    __syncwarp();
                    int run = *carry;
                    for (int c0 = 0; c0 < cnt; c0 += 32) {
                        const int e = c0 + lane;
                        int id = e < cnt ? (int)owner[e] : 0;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, id, o); if (lane >= o) id = max(id, t); }
                        id = max(id, run);
                        run = __shfl_sync(0xffffffffu, id, 31);
                        if (e < cnt) {
                            const float4 a = __ldg(&sorted[seg_src[id - 1] + base + e]);
                            float *d = st + (e >> 1) * 8 + (e & 1);
                            d[0] = a.x; d[2] = a.y; d[4] = a.z; d[6] = a.w;
                        }
                    }
                    if ((cnt & 1) && lane == 0) {          // odd count: the pair's second atom is infinitely far away
                        float *d = st + (cnt >> 1) * 8 + 1;
                        d[0] = 1e18f; d[2] = 1e18f; d[4] = 1e18f; d[6] = __int_as_float(-1);
                    }
                    __syncwarp();
                    // ---- sweep: two staged atoms per step; only the PAIR that holds the new minimum is recorded, the
                    // atom is identified after the sweep (b1 improves O(log n) times, the sweep visits n atoms)
                    const float4 *sp = reinterpret_cast<const float4 *>(st);
                    const int npair = (cnt + 1) >> 1;
                    int bj = -1;
#pragma unroll kSweepUnroll
                    for (int j = 0; j < npair; ++j) {
                        const float4 xy = sp[2 * j];
                        const float2 zz = *reinterpret_cast<const float2 *>(&sp[2 * j + 1]);
                        const float2 dx = __fadd2_rn(make_float2(xy.x, xy.y), nqx);
                        const float2 dy = __fadd2_rn(make_float2(xy.z, xy.w), nqy);
                        const float2 dz = __fadd2_rn(zz, nqz);
                        const float2 d2 = __ffma2_rn(dx, dx, __ffma2_rn(dy, dy, __fmul2_rn(dz, dz)));
                        const float lo = fminf(d2.x, d2.y), hi = fmaxf(d2.x, d2.y);
                        b2 = fminf(b2, fminf(fmaxf(b1, lo), hi));       // second smallest of {b1, b2, d2.x, d2.y}
                        bj = lo < b1 ? j : bj;
                        b1 = fminf(b1, lo);
                    }
                    if (bj >= 0) {      // which atom of pair bj: the same arithmetic gives the same two numbers
                        const float4 xy = sp[2 * bj], zw = sp[2 * bj + 1];
                        const float2 dx = __fadd2_rn(make_float2(xy.x, xy.y), nqx);
                        const float2 dy = __fadd2_rn(make_float2(xy.z, xy.w), nqy);
                        const float2 dz = __fadd2_rn(make_float2(zw.x, zw.y), nqz);
                        const float2 d2 = __ffma2_rn(dx, dx, __ffma2_rn(dy, dy, __fmul2_rn(dz, dz)));
                        bi = d2.x <= d2.y ? __float_as_int(zw.z) : __float_as_int(zw.w);
                    }
                }
              }   // pass (x-limited rings: two passes per ring)
                const float mine = valid ? fminf(b1 + S.tol_d2, S.search2) : 0.f;
                bound = fminf(bound, __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(mine))));
            }
        }
        if (valid) S.res[__float_as_int(q.w)] = make_float4(b1, b2, __int_as_float(bi), 0.f);
    }
    if (COUNT && pair_evals) {
        npairs *= 32ull;
        if (lane == 0 && npairs) atomicAdd(pair_evals, npairs);
    }
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

// ---------------------------------------------------------------------------------------------
// Finalisation: per molecule, combine its atoms' results (update_md, src/minimum_distances.jl:30-39), finalise
// the winning pair and the reference-atom pair in fp64 with the reference's arithmetic, histogram
// (update_counters!, src/update_counters.jl:43-88) -- or defer an ambiguous molecule to the exact kernel.
// The work items of ALL frames of the batch are one queue of 256-item chunks; a block keeps pulling chunks, counts
// into its shared-memory histograms (HistPriv) and flushes them once -- the hits of several frames per flush.
// ---------------------------------------------------------------------------------------------
#define CMX_FIN_THREADS 256
template <bool RANDOM>
__global__ void __launch_bounds__(CMX_FIN_THREADS)
k_finalise(const GridFrame *__restrict__ fds, int nframes, Prob P, int s0) {
    extern __shared__ unsigned hist_smem[];
    __shared__ __align__(16) GridFrame F_sh;
    __shared__ int s_end[CMX_MAX_BATCH];       // chunks of frames 0..k, cumulative
    __shared__ int s_chunk;
    HistPriv H = hist_init(P, hist_smem);
    if (threadIdx.x < nframes) s_end[threadIdx.x] = (fds[threadIdx.x].sc[RANDOM ? SC_RWORK : SC_WORK] + CMX_FIN_THREADS - 1) / CMX_FIN_THREADS;
    __syncthreads();
    if (threadIdx.x == 0) { int run = 0; for (int k = 0; k < nframes; ++k) { run += s_end[k]; s_end[k] = run; } }
    __syncthreads();
    const int nchunks = s_end[nframes - 1];
    int *queue = fds[0].sc + (RANDOM ? SC_FQ_RAND : SC_FQ_REAL);
    int cur = -1;
    const GridFrame &F = F_sh;
    const Geom &g = F.g;
    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) s_chunk = atomicAdd(queue, 1);
        __syncthreads();
        const int chunk = s_chunk;
        if (chunk >= nchunks) break;
        int fi = 0;
        while (chunk >= s_end[fi]) ++fi;
        if (fi != cur) {     // (uniform in the block) this chunk belongs to another frame: fetch its descriptor
            const int4 *src_ = reinterpret_cast<const int4 *>(fds + fi);
            int4 *dst_ = reinterpret_cast<int4 *>(&F_sh);
            for (int k_ = threadIdx.x; k_ < (int)(sizeof(GridFrame) / 16); k_ += blockDim.x) dst_[k_] = __ldg(src_ + k_);
            cur = fi;
            __syncthreads();
        }
        const int w = (chunk - (fi ? s_end[fi - 1] : 0)) * CMX_FIN_THREADS + threadIdx.x;
        if (w >= F.sc[RANDOM ? SC_RWORK : SC_WORK]) continue;
        const int item = (RANDOM ? F.rand_worklist : F.worklist)[w];
        int sample = 0, mol = item;
        if (RANDOM) { int sl = item / P.nv_mols; mol = item - sl * P.nv_mols; sample = s0 + sl; }
        const float4 *r = F.res + (size_t)w * P.nv_apm;
        float best = CUDART_INF_F, second = CUDART_INF_F; int bi = -1, bk = -1;
        for (int k = 0; k < P.nv_apm; ++k) {
            float4 a = r[k];
            if (a.x < best) { second = fminf(fminf(second, best), a.y); best = a.x; bi = __float_as_int(a.z); bk = k; }
            else second = fminf(second, a.x);
        }
        int cls = classify(g, best, second);
        if (cls == 0) continue;   // list entry stays "not within"
        float4 rr = r[P.iref];
        int rcls = classify(g, rr.x, rr.y);
        if (cls == 2 || rcls == 2) {
            int slot = atomicAdd(&F.sc[RANDOM ? SC_DEF_RAND : SC_DEF_REAL], 1);
            (RANDOM ? F.def_rand : F.def_real)[slot] = ((u64)(RANDOM ? 1 + sample : 0) << 32) | (u64)(uint32_t)mol;
            (RANDOM ? F.def_rand_info : F.def_real_info)[slot] = make_float2(best, rr.x);   // fp32 bounds: the exact kernel only looks at atoms that can matter
            continue;
        }
        const float *xs = F.xs;
        auto pos = [&](int k, double &ex, double &ey, double &ez) {
            if (RANDOM) { const double *xe = F.xexact + ((size_t)P.nv_apm * w + k) * 3; ex = xe[0]; ey = xe[1]; ez = xe[2]; }
            else { const float *xr = F.xv + ((size_t)P.nv_apm * mol + k) * 3; ex = (double)xr[0]; ey = (double)xr[1]; ez = (double)xr[2]; }
        };
        double ex, ey, ez;
        pos(bk, ex, ey, ez);
        MdRec e;
        e.d = dist_pbc64(g, (double)xs[3 * bi], (double)xs[3 * bi + 1], (double)xs[3 * bi + 2], ex, ey, ez);
        e.i = bi; e.j = mol * P.nv_apm + bk; e.flags = 1; e.dref = CUDART_INF; e.pad = 0;
        if (rcls == 1) {
            int ri = __float_as_int(rr.z);
            pos(P.iref, ex, ey, ez);
            e.dref = dist_pbc64(g, (double)xs[3 * ri], (double)xs[3 * ri + 1], (double)xs[3 * ri + 2], ex, ey, ez);
            e.flags |= 2;
        }
        count_hit_priv(P, H, F.weight, RANDOM, e.d, e.i, e.j, 1ull);
        if (e.flags & 2) count_ref_priv(P, H, F.weight, RANDOM, e.dref);
        MdRec *list = RANDOM ? F.rand_list : F.list;
        if (list) list[RANDOM ? (size_t)sample * P.nv_mols + mol : (size_t)mol] = e;
    }
    hist_flush(P, H, RANDOM);
}

// ---------------------------------------------------------------------------------------------
// Random-phase cull: the random placements, by the position of their centre
// ---------------------------------------------------------------------------------------------
// A thread owns one solvent slot and walks the samples of the chunk (up to 32 per pass: its survivors are a bit mask);
// the centre is evaluated in fp32 (its error, ~1e-5 A, is far below the 1e-3 A margin of the test).  The block's
// survivors are appended with ONE scan and ONE global atomic per pass.
#define CMX_FRAND_THREADS 256
__global__ void __launch_bounds__(CMX_FRAND_THREADS)
k_filter_rand(const GridFrame *__restrict__ fds, Prob P, int s0, int s1) {
    CMX_FRAME(F)
    __shared__ int warp_tot[CMX_FRAND_THREADS / 32];
    __shared__ int s_base;
    if (F.nrand_k == 0) return;                  // this solute molecule is the reference of no sample of the frame
    const int mol = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const float rmax = __int_as_float(F.sc[SC_RMAX]);
    const bool cull = rmax <= g.rmax_bound;      // else: transform window not valid for this radius, every placement survives
    const float lim = g.cut_hi + rmax + 2e-3f, lim2 = lim * lim;
    const float m0 = (float)g.m[0], m1 = (float)g.m[1], m2 = (float)g.m[2], m3 = (float)g.m[3], m4 = (float)g.m[4], m5 = (float)g.m[5],
                m6 = (float)g.m[6], m7 = (float)g.m[7], m8 = (float)g.m[8], c0 = (float)g.ctr[0], c1 = (float)g.ctr[1], c2 = (float)g.ctr[2];
    for (int sb = s0 + 32 * (int)blockIdx.z; sb < s1; sb += 32 * (int)gridDim.z) {     // (grid.z strides the passes: many samples, few molecules)
        const int ns = min(32, s1 - sb);
        unsigned mask = 0u;
        if (mol < P.nv_mols && mol != F.skip_mol) {
            for (int t = 0; t < ns; ++t) {
                const int sample = sb + t;
                if (P.ns_mols != 1 && ref_solute_of_sample(P, F.frame, (uint32_t)sample) != F.isolute) continue;
                bool near = true;
                if (cull) {
                    const uint4 r0 = philox4x32((uint32_t)mol, (uint32_t)sample, F.frame, 0u, P.seed_lo, P.seed_hi);
                    const float sc = 1.0f / 4294967296.0f;
                    const float u0 = ((float)r0.y + 0.5f) * sc, u1 = ((float)r0.z + 0.5f) * sc, u2 = ((float)r0.w + 0.5f) * sc;
                    const float cx_ = m0 * u0 + m3 * u1 + m6 * u2 - c0;
                    const float cy_ = m1 * u0 + m4 * u1 + m7 * u2 - c1;
                    const float cz_ = m2 * u0 + m5 * u1 + m8 * u2 - c2;
                    near = cull_lb2(g, F.lbd2, cx_, cy_, cz_) <= lim2;
                }
                mask |= near ? (1u << t) : 0u;
            }
        }
        // block-wide exclusive scan of the survivor counts
        const int mine = __popc(mask);
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        __syncthreads();                          // (warp_tot / s_base of the previous pass are no longer read)
        if (lane == 31) warp_tot[wid] = incl;
        __syncthreads();
        if (threadIdx.x == 0) {
            int run = 0;
            for (int k = 0; k < CMX_FRAND_THREADS / 32; ++k) { int v = warp_tot[k]; warp_tot[k] = run; run += v; }
            s_base = run ? atomicAdd(&F.sc[SC_RWORK], run) : 0;
        }
        __syncthreads();
        int pos = s_base + warp_tot[wid] + incl - mine;
        while (mask) {
            const int t = __ffs(mask) - 1;
            mask &= mask - 1;
            F.rand_worklist[pos++] = (sb + t - s0) * P.nv_mols + mol;      // item within the chunk
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Exact resolve of the deferred (ambiguous in fp32) molecules: fp64 with the reference's arithmetic over the
// solute atoms within reach of the fp32 result, rule "smallest (d, j, i) wins".  One block per deferred item.
// ---------------------------------------------------------------------------------------------
struct RealMolG {   // adapter: same interface as RandMol::get(g, k, ...)
    RealMol m;
    __device__ __forceinline__ void get(const Geom &, int k, double &x, double &y, double &z) const { m.get(k, x, y, z); }
};
struct ExactBest { double d, dref; int i, j; };
__device__ __forceinline__ bool better(double d, int j, int i, const ExactBest &b) {
    return d < b.d || (d == b.d && (j < b.j || (j == b.j && i < b.i)));
}

// Exact fp64 evaluation of one deferred molecule.  The fp32 search already bounded the answer: only
// solute atoms (incl. images) whose fp32 distance is within 4 tau of the molecule's best, or of the
// reference atom's best, can be the exact winner -- everything else is skipped after one fp32 test.
template <class Mol>
__device__ __forceinline__ void resolve_one(const Geom &g, const Prob &P, const float *__restrict__ xs,
                                            const float4 *__restrict__ sorted, const int *__restrict__ cell_start,
                                            const Mol &mol, int molidx, bool random, double weight, float2 info, MdRec *out, ExactBest *sh) {
    const float capd = g.cut_hi + 4.f * g.tau;
    float lim_m = fminf(sqrtf(info.x) + 4.f * g.tau, capd), lim_r = fminf(sqrtf(info.y) + 4.f * g.tau, capd);
    lim_m *= lim_m; lim_r *= lim_r;
    ExactBest b; b.d = CUDART_INF; b.dref = CUDART_INF; b.i = 0x7fffffff; b.j = 0x7fffffff;
    const int wid = threadIdx.x >> 5, nw = blockDim.x >> 5, lane = threadIdx.x & 31;
    for (int k = 0; k < P.nv_apm; ++k) {
        double ex, ey, ez; mol.get(g, k, ex, ey, ez);
        double wx, wy, wz; wrap_to_cell(g, ex, ey, ez, wx, wy, wz);
        float px = (float)(wx - g.ctr[0]), py = (float)(wy - g.ctr[1]), pz = (float)(wz - g.ctr[2]);
        int j = molidx * P.nv_apm + k;
        // only the grid rows / x-span within reach of this atom (warps take rows, lanes stride the span)
        const float lim = (k == P.iref) ? fmaxf(lim_m, lim_r) : lim_m;
        const float reach = sqrtf(lim) + 4e-3f;
        const int ry_lo = max((int)floorf((py - reach - g.gmin[1]) * g.inv_side), 0), ry_hi = min((int)floorf((py + reach - g.gmin[1]) * g.inv_side), g.ny - 1);
        const int rz_lo = max((int)floorf((pz - reach - g.gmin[2]) * g.inv_side), 0), rz_hi = min((int)floorf((pz + reach - g.gmin[2]) * g.inv_side), g.nz - 1);
        const int cxl = max((int)floorf((px - reach - g.gmin[0]) * g.inv_sidex), 0), cxh = min((int)floorf((px + reach - g.gmin[0]) * g.inv_sidex), g.nx - 1);
        const int nry = ry_hi - ry_lo + 1, nrows = nry * (rz_hi - rz_lo + 1);
        if (cxl > cxh) continue;
        for (int r = wid; r < nrows; r += nw) {
            int rzq = r / nry;
            int rowbase = ((rz_lo + rzq) * g.ny + ry_lo + (r - rzq * nry)) * g.nx;
            int pa = __ldg(&cell_start[rowbase + cxl]), pb = __ldg(&cell_start[rowbase + cxh + 1]);
            for (int p = pa + lane; p < pb; p += 32) {
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
            count_hit(P, weight, random, e.d, e.i, e.j, 1ull);
            if (e.flags & 2) count_ref(P, weight, random, e.dref);
        }
        if (out) *out = e;
    }
    __syncthreads();
}

#define CMX_RESOLVE_THREADS 256
// RANDOM=false: the real-phase deferred molecules (their list entries feed the bulk list, so this runs before the
// bulk compaction); RANDOM=true: the random-phase ones.  stats[1] accumulates the number of deferred molecules.
template <bool RANDOM>
__global__ void __launch_bounds__(CMX_RESOLVE_THREADS)
k_resolve(const GridFrame *__restrict__ fds, Prob P, u64 *__restrict__ stats) {
    CMX_FRAME(F)
    __shared__ ExactBest sh[CMX_RESOLVE_THREADS];
    const int count = F.sc[RANDOM ? SC_DEF_RAND : SC_DEF_REAL];
    if (stats && blockIdx.x == 0 && threadIdx.x == 0 && count) atomicAdd(&stats[1], (u64)count);
    const u64 *deferred = RANDOM ? F.def_rand : F.def_real;
    const float2 *deferred_info = RANDOM ? F.def_rand_info : F.def_real_info;
    for (int w = blockIdx.x; w < count; w += gridDim.x) {
        u64 item = deferred[w];
        const float2 info = deferred_info[w];
        int phase = (int)(item >> 32), mol = (int)(item & 0xffffffffu);
        if (!RANDOM) {
            RealMolG rl; rl.m.x = F.xv + (size_t)3 * P.nv_apm * mol;
            resolve_one(g, P, F.xs, F.sorted, F.cell_start, rl, mol, false, F.weight, info, &F.list[mol], sh);
        } else {
            int sample = phase - 1;
            uint4 r0 = philox4x32((uint32_t)mol, (uint32_t)sample, F.frame, 0u, P.seed_lo, P.seed_hi);
            uint4 r1 = philox4x32((uint32_t)mol, (uint32_t)sample, F.frame, 1u, P.seed_lo, P.seed_hi);
            int nb = F.sc[SC_NBULK];
            int jmol = nb > 0 ? F.bulk_idx[pick(r0.x, (uint32_t)nb)] : (int)pick(r0.x, (uint32_t)P.nv_mols);
            RandMol rm; rm.init(g, F.xv + (size_t)3 * P.nv_apm * jmol, P.nv_apm, P.iref, r0, r1);
            resolve_one(g, P, F.xs, F.sorted, F.cell_start, rm, mol, true, F.weight, info,
                        F.rand_list ? &F.rand_list[(size_t)sample * P.nv_mols + mol] : nullptr, sh);
        }
    }
}

}  // namespace cmx
