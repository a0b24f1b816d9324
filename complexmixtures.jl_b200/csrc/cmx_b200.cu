// cmx_b200.cu -- C-ABI (include/cmx_b200.h) and host-side frame pipeline of libcmx_b200.so.
//
// Host responsibilities restated from the reference's chunk task (src/mddf.jl:288-337):
// per-handle state (build_particle_system / Buffer / Result -> cmx_create), frame staging
// (pinned ring, async H2D), per-frame kernel sequence (mddf_frame!, src/mddf.jl:361-429),
// frame-weight handling and the final counters (sum!, src/results.jl:629-649).
#include <cub/cub.cuh>
#include <cuda_runtime.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <cctype>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/cmx_b200.h"
#include "cmx_kernels.cuh"
#include "cmx_pairs.cuh"

using namespace cmx;

namespace {

std::string g_create_error;

struct Slot {
    float *h_in = nullptr;   // pinned: solute xyz then solvent xyz (autocorrelation: solvent only)
    float *d_in = nullptr;
    cudaEvent_t h2d_done = nullptr, consumed = nullptr;
    bool in_flight = false;
};

// Pinned staging memory is read by the GPU's DMA engine for every frame: allocate it on the NUMA node the GPU hangs
// off (sysfs numa_node of its PCI function), unless the caller already runs under a memory policy (numactl etc.).
int gpu_numa_node(int device) {
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, (int)sizeof bus, device) != cudaSuccess) { (void)cudaGetLastError(); return -1; }
    for (char *p = bus; *p; ++p) *p = (char)std::tolower((unsigned char)*p);
    std::string path = std::string("/sys/bus/pci/devices/") + bus + "/numa_node";
    FILE *f = std::fopen(path.c_str(), "r");
    if (!f) return -1;
    int node = -1;
    if (std::fscanf(f, "%d", &node) != 1) node = -1;
    std::fclose(f);
    return node;
}
struct NumaPrefer {
    bool active = false;
    explicit NumaPrefer(int node) {
        if (node < 0 || node >= 1024) return;
        int mode = -1;
        if (syscall(SYS_get_mempolicy, &mode, nullptr, 0ul, nullptr, 0ul) != 0 || mode != 0 /*MPOL_DEFAULT*/) return;
        unsigned long mask[16] = {0};
        mask[node / 64] |= 1ul << (node % 64);
        active = syscall(SYS_set_mempolicy, 1 /*MPOL_PREFERRED*/, mask, 1025ul) == 0;
    }
    ~NumaPrefer() { if (active) syscall(SYS_set_mempolicy, 0 /*MPOL_DEFAULT*/, nullptr, 0ul); }
};

template <class T>
struct DevBuf {
    T *p = nullptr; size_t n = 0;
    cudaError_t ensure(size_t want, bool zero = false) {
        if (want <= n && p) return cudaSuccess;
        if (p) { cudaError_t e = cudaFree(p); if (e != cudaSuccess) return e; p = nullptr; n = 0; }
        size_t cap = want + want / 4 + 64;
        cudaError_t e = cudaMalloc(&p, cap * sizeof(T));
        if (e != cudaSuccess) return e;
        n = cap;
        if (zero) return cudaMemset(p, 0, cap * sizeof(T));
        return cudaSuccess;
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
};

}  // namespace

struct cmx_feed;   // native DCD feed ring + group-reduction scratch (cmx_feed.inl)

// Everything one in-flight frame needs: its compute stream and scratch buffers.
struct FrameCtx {
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_end = nullptr;
    DevBuf<int> d_cell_count, d_cell_start;
    DevBuf<float4> d_sorted;
    DevBuf<u64> d_occ;                 // [8 scalars][cull bitmap][row bitmap]
    DevBuf<float4> d_qpos, d_qsorted, d_res;   // query atoms of the current phase: positions, tile order, results
    DevBuf<int> d_qcell_count, d_qcell_start;
    DevBuf<double> d_xexact;
    DevBuf<unsigned char> d_edt_x;
    DevBuf<unsigned short> d_edt_xy;
    DevBuf<float> d_lbd2;
    DevBuf<MdRec> d_list;
    DevBuf<int> d_worklist, d_rand_worklist, d_bulk_idx;
    DevBuf<u64> d_def_real, d_def_rand;
    DevBuf<float2> d_def_real_info, d_def_rand_info;
    DevBuf<int> d_scalars;          // [0] work_count, [1] rand_work_count, [2] def_real, [3] def_rand, [4] n_bulk, [5] rmax bits, [8] sticky overflow
    DevBuf<unsigned char> d_cub_tmp;
    PairScratch pairs;
    int *h_scalars = nullptr;       // pinned mirror (rmax feedback)
    void release() {
        d_cell_count.release(); d_cell_start.release(); d_sorted.release(); d_occ.release();
        d_qpos.release(); d_qsorted.release(); d_res.release(); d_qcell_count.release(); d_qcell_start.release(); d_xexact.release(); d_edt_x.release();
        d_edt_xy.release(); d_lbd2.release(); d_list.release(); d_worklist.release(); d_rand_worklist.release();
        d_bulk_idx.release(); d_def_real.release(); d_def_rand.release(); d_def_real_info.release(); d_def_rand_info.release(); d_scalars.release(); d_cub_tmp.release();
        if (h_scalars) cudaFreeHost(h_scalars);
        if (ev_end) cudaEventDestroy(ev_end);
        if (stream) cudaStreamDestroy(stream);
    }
};

struct cmx_handle {
    cmx_config cfg{};
    std::string err;
    int device = 0;
    int path = 1;            // 1 grid path, 2 molecule-pair path
    int nbins = 0;
    size_t ns_atoms = 0, nv_atoms = 0, in_floats = 0;
    double cut_eff = 0;
    int Kdiv = 2;
    double side = 0, sidex = 0, cside = 0, qside = 0;
    cudaStream_t s_copy = nullptr;
    int64_t submitted = 0;
    std::vector<Slot> ring;
    int next_slot = 0, acquired = -1;
    // static device data
    Prob P{};
    DevBuf<int> d_sol_off, d_sol_ids, d_solv_off, d_solv_ids;
    DevBuf<u64> d_cnt;              // integer run accumulators (contiguous block)
    DevBuf<double> d_acc;           // fp64 accumulators, only once the frame weight changes
    DevBuf<double> d_emit;          // staging of cmx_finish
    size_t cnt_len = 0;
    bool acc_used = false;
    // per-frame scratch lives in FrameCtx (one per compute stream)
    DevBuf<MdRec> d_rand_list, d_list_all;   // parity hooks (keep_lists => one stream)
    DevBuf<u64> d_stats;            // [0] pair_evals, [1] deferred total
    std::vector<FrameCtx *> ctx;    // frames are dealt round-robin to the contexts; kernels of different frames overlap
    FrameCtx *cur = nullptr;
    int active_ctx = 0;             // contexts actually used (option "active_streams"; 0 = all)
    float rmax_bound = 0.f;
    int sample_chunk = 1;           // samples of the random phase per pass (grid path)
    // bookkeeping
    double cur_weight = 1.0; bool have_weight = false;
    double volume_total = 0, sum_weights = 0;
    cmx_stats stats{};
    bool count_pairs = false, profile = false;
    cudaEvent_t ev_first = nullptr, ev_last = nullptr; bool ev_first_set = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;
    std::vector<int> prof_tags;
    size_t prof_used = 0;
    int64_t last_frame = -1;
    Geom last_g{};
    const float *last_dsol = nullptr, *last_dsolv = nullptr;
    int num_sms = 148;
    int search_grid[2] = {148 * 5, 148 * 4};   // one resident wave of k_tile_search<false/true> (occupancy query at create)
    int search_blocks_env = 0;
    cmx_feed *feed = nullptr;
    int numa_node = -1;                        // NUMA node of the GPU (-1 unknown): pinned staging memory is placed there
};

namespace {

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            h->err = std::string(#call) + ": " + cudaGetErrorString(e_);                           \
            return CMX_ERR_CUDA;                                                                   \
        }                                                                                          \
    } while (0)

int fail(cmx_handle *h, int code, const std::string &msg) { h->err = msg; return code; }

// molecule-pair path (cmx_pairs_host.inl)
void feed_destroy(cmx_handle *h);
int frame_pair_path(cmx_handle *h, const float *d_solute, const float *d_solvent, uint32_t frame, const Geom &g);
int pairs_create(cmx_handle *h);
void pairs_release(cmx_handle *h);
PairGeom make_pair_geom(cmx_handle *h, const Geom &g);

// ---- host Philox (same counter convention as the device) for ref_solutes -------------------
void philox_host(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3; k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
int ref_solute_host(const cmx_handle *h, uint32_t frame, uint32_t s) {
    uint32_t o[4];
    philox_host(0xffffffffu, s, frame, 2u, h->P.seed_lo, h->P.seed_hi, o);
    return (int)(((uint64_t)o[0] * (uint64_t)h->cfg.solute_nmols) >> 32);
}

// ---- geometry ---------------------------------------------------------------------------------
int build_geom(cmx_handle *h, const double cell[9], Geom &g) {
    std::memset(&g, 0, sizeof g);
    std::memcpy(g.m, cell, sizeof(double) * 9);
    const double *a = cell, *b = cell + 3, *c = cell + 6;
    double mind = std::min(std::fabs(a[0]), std::min(std::fabs(b[1]), std::fabs(c[2])));
    double tol = 1e-10 * mind;   // convert_unitcell, src/Trajectory.jl:72-77
    g.ortho = std::fabs(a[1]) < tol && std::fabs(a[2]) < tol && std::fabs(b[0]) < tol && std::fabs(b[2]) < tol &&
              std::fabs(c[0]) < tol && std::fabs(c[1]) < tol;
    double bxc[3] = {b[1] * c[2] - b[2] * c[1], b[2] * c[0] - b[0] * c[2], b[0] * c[1] - b[1] * c[0]};
    double cxa[3] = {c[1] * a[2] - c[2] * a[1], c[2] * a[0] - c[0] * a[2], c[0] * a[1] - c[1] * a[0]};
    double axb[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
    double det = axb[0] * c[0] + axb[1] * c[1] + axb[2] * c[2];
    if (!(std::fabs(det) > 0)) return fail(h, CMX_ERR_CELL, "singular unit cell");
    for (int k = 0; k < 3; ++k) {
        g.inv[0 + 3 * k] = bxc[k] / det; g.inv[1 + 3 * k] = cxa[k] / det; g.inv[2 + 3 * k] = axb[k] / det;
    }
    g.invl[0] = 1.0 / a[0]; g.invl[1] = 1.0 / b[1]; g.invl[2] = 1.0 / c[2];
    double w[3] = {std::fabs(det) / std::sqrt(bxc[0] * bxc[0] + bxc[1] * bxc[1] + bxc[2] * bxc[2]),
                   std::fabs(det) / std::sqrt(cxa[0] * cxa[0] + cxa[1] * cxa[1] + cxa[2] * cxa[2]),
                   std::fabs(det) / std::sqrt(axb[0] * axb[0] + axb[1] * axb[1] + axb[2] * axb[2])};
    for (int k = 0; k < 3; ++k)
        if (w[k] < 2.0 * h->cut_eff) {
            char buf[200];
            std::snprintf(buf, sizeof buf, "unit cell too small for the cutoff: perpendicular width %.4f < 2*%.4f "
                          "(CellListMap requires sides > 2*cutoff)", w[k], h->cut_eff);
            return fail(h, CMX_ERR_CELL, buf);
        }
    // AABB of the primary cell
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int n = 0; n < 8; ++n)
        for (int k = 0; k < 3; ++k) {
            double v = ((n & 1) ? a[k] : 0) + ((n & 2) ? b[k] : 0) + ((n & 4) ? c[k] : 0);
            lo[k] = std::min(lo[k], v); hi[k] = std::max(hi[k], v);
        }
    double maxabs = 0;
    double margin0 = h->cut_eff + 0.1;
    for (int k = 0; k < 3; ++k) maxabs = std::max(maxabs, 0.5 * (hi[k] - lo[k]) + margin0);
    int ex; std::frexp(maxabs, &ex);                 // maxabs = f * 2^ex, f in [0.5,1)
    double ulp = std::ldexp(1.0, ex - 24);
    double tau = std::max(16.0 * ulp, 2e-5);
    g.tau = (float)tau; g.cut = (float)h->cut_eff;
    g.cut_lo = (float)(h->cut_eff - tau); g.cut_hi = (float)(h->cut_eff + tau);
    g.search2 = (float)((h->cut_eff + tau) * (h->cut_eff + tau) * (1.0 + 1e-6));
    g.tol_d2 = (float)(2.0 * h->cut_eff * tau + tau * tau);
    g.cutd = h->cut_eff;
    double margin = h->cut_eff + tau + 0.05;
    for (int k = 0; k < 3; ++k) {
        g.elo[k] = lo[k] - margin; g.ehi[k] = hi[k] + margin; g.ctr[k] = 0.5 * (lo[k] + hi[k]);
        g.gmin[k] = (float)(g.elo[k] - g.ctr[k]);
    }
    g.side = (float)h->side; g.inv_side = (float)(1.0 / h->side);
    g.sidex = (float)h->sidex; g.inv_sidex = (float)(1.0 / h->sidex);
    g.cut_hi2 = g.cut_hi * g.cut_hi * (1.0f + 1e-6f);
    g.nx = (int)std::ceil((g.ehi[0] - g.elo[0]) / h->sidex) + 1;
    g.ny = (int)std::ceil((g.ehi[1] - g.elo[1]) / h->side) + 1;
    g.nz = (int)std::ceil((g.ehi[2] - g.elo[2]) / h->side) + 1;
    g.cside = (float)h->cside; g.inv_cside = (float)(1.0 / h->cside);
    g.ncx = (int)std::ceil((g.ehi[0] - g.elo[0]) / h->cside) + 1;
    g.ncy = (int)std::ceil((g.ehi[1] - g.elo[1]) / h->cside) + 1;
    g.ncz = (int)std::ceil((g.ehi[2] - g.elo[2]) / h->cside) + 1;
    g.cw = (g.ncx + 63) / 64;
    g.rw = (g.nx + 63) / 64;
    g.qside = (float)h->qside; g.inv_qside = (float)(1.0 / h->qside);
    g.nqx = (int)std::ceil((g.ehi[0] - g.elo[0]) / h->qside) + 1;
    g.nqy = (int)std::ceil((g.ehi[1] - g.elo[1]) / h->qside) + 1;
    g.nqz = (int)std::ceil((g.ehi[2] - g.elo[2]) / h->qside) + 1;
    g.rmax_bound = h->rmax_bound;
    g.dwin = std::min(15, (int)std::ceil((h->cut_eff + tau + h->rmax_bound + 1e-3) / h->cside) + 2);
    if ((double)g.nx * g.ny * g.nz > 2.0e8) return fail(h, CMX_ERR_CELL, "search grid too large for this cell/cutoff");
    return CMX_OK;
}

template <class K, class... Args>
void launch(cmx_handle *h, K kernel, dim3 grid, dim3 block, Args... args) {
    kernel<<<grid, block, 0, h->cur->stream>>>(args...);
    h->stats.kernel_launches++;
}

cudaEvent_t prof_begin(cmx_handle *h, int tag = 0) {
    if (!h->profile) return nullptr;
    if (h->prof_used == h->prof_events.size()) {
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        h->prof_events.push_back({a, b}); h->prof_tags.push_back(0);
    }
    h->prof_tags[h->prof_used] = tag;
    cudaEventRecord(h->prof_events[h->prof_used].first, h->cur->stream);
    return h->prof_events[h->prof_used].second;
}
void prof_end(cmx_handle *h, cudaEvent_t e) {
    if (!e) return;
    cudaEventRecord(e, h->cur->stream);
    h->prof_used++;
}
void prof_collect(cmx_handle *h) {
    for (size_t k = 0; k < h->prof_used; ++k) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, h->prof_events[k].first, h->prof_events[k].second) == cudaSuccess) {
            h->stats.gpu_ms_main += ms;
            (h->prof_tags[k] ? h->stats.gpu_ms_search_random : h->stats.gpu_ms_search_real) += ms;
        }
    }
    h->prof_used = 0;
}

// one search phase (real or random) over the molecules of a work list: tile the query atoms, search, combine
template <bool RANDOM>
int search_phase(cmx_handle *h, const Geom &g, const float *xs, const float *xv, const int *worklist, const int *work_count,
                 size_t max_atoms, MdRec *list, u64 *deferred, float2 *def_info, int *def_count, int tag, int s0) {
    FrameCtx &x = *h->cur;
    size_t nqc = (size_t)g.nqx * g.nqy * g.nqz;
    // (query positions, res and the per-cell counts were produced by k_gen_*)
    if (nqc + 1 <= CMX_SCAN_SMALL_MAX)
        launch(h, k_scan_small<TileCountOp>, dim3(1), dim3(CMX_SCAN_THREADS), (const int *)x.d_qcell_count.p, x.d_qcell_start.p, (int)(nqc + 1), TileCountOp());
    else {
        size_t tmp_bytes = x.d_cub_tmp.n;
        cub::TransformInputIterator<int, TileCountOp, const int *> tiles_of((const int *)x.d_qcell_count.p, TileCountOp());
        CK(cub::DeviceScan::ExclusiveSum(x.d_cub_tmp.p, tmp_bytes, tiles_of, x.d_qcell_start.p, (int)(nqc + 1), x.stream));
        h->stats.kernel_launches += 2;
    }
    // worst case: every atom in a tile of its own cell's last, partially filled tile
    size_t slots = std::min(max_atoms + 32 * nqc, x.d_qsorted.n);
    CK(cudaMemsetAsync(x.d_qsorted.p, 0xff, sizeof(float4) * slots, x.stream));
    int nb = (int)std::min<size_t>((max_atoms + 255) / 256, (size_t)h->num_sms * 8);
    launch(h, k_qscatter, dim3(std::max(nb, 1)), dim3(256), g, h->P, work_count, (const float4 *)x.d_qpos.p, x.d_qcell_count.p,
           (const int *)x.d_qcell_start.p, x.d_qsorted.p);
    cudaEvent_t pe = prof_begin(h, tag);
    u64 *pev = h->count_pairs ? h->d_stats.p : nullptr;
    int *tile_queue = (int *)x.d_occ.p + (RANDOM ? 7 : 6);   // per-frame scalars, zeroed with the bitmaps
    // Each of the n frames in flight gets 1/n of the resident block slots of an SM: the searches of several frames
    // then share the SMs with the small kernels of the other frames instead of one search filling every register
    // file for its whole duration (measured on C2, 8 streams: 5 blocks/SM 8.2 k frames/s, 2 -> 8.9 k, 1 -> 9.4 k).
    const int nfly = h->active_ctx > 0 ? h->active_ctx : (int)h->ctx.size();
    auto grid_of = [&](int k) {
        if (h->search_blocks_env > 0) return h->num_sms * h->search_blocks_env;
        int per_sm = h->search_grid[k] / h->num_sms;
        return h->num_sms * std::max(1, (per_sm + nfly - 1) / nfly);
    };
    if (pev) launch(h, k_tile_search<true>, dim3(grid_of(1)), dim3(256), g, (const int *)x.d_cell_start.p, (const float4 *)x.d_sorted.p,
                    (const u64 *)(x.d_occ.p + 4 + (size_t)g.ncy * g.ncz * g.cw), (const float4 *)x.d_qsorted.p, (const int *)x.d_qcell_start.p, (int)nqc, x.d_res.p, pev, tile_queue);
    else launch(h, k_tile_search<false>, dim3(grid_of(0)), dim3(256), g, (const int *)x.d_cell_start.p, (const float4 *)x.d_sorted.p,
                (const u64 *)(x.d_occ.p + 4 + (size_t)g.ncy * g.ncz * g.cw), (const float4 *)x.d_qsorted.p, (const int *)x.d_qcell_start.p, (int)nqc, x.d_res.p, pev, tile_queue);
    prof_end(h, pe);
    launch(h, k_finalise<RANDOM>, dim3(h->num_sms * 8), dim3(128), g, h->P, xs, xv, (const float4 *)x.d_res.p,
           (const double *)(RANDOM ? x.d_xexact.p : nullptr), worklist, work_count, list, deferred, def_info, def_count, s0);
    return CMX_OK;
}

// ---- one frame on the grid path (mddf_frame!, src/mddf.jl:361-429) --------------------------------
int frame_grid_path(cmx_handle *h, const float *d_solute, const float *d_solvent, uint32_t frame, const Geom &g) {
    const cmx_config &c = h->cfg;
    const int ns_apm = c.solute_natomspermol, nv_mols = c.solvent_nmols;
    size_t ncells = (size_t)g.nx * g.ny * g.nz, ncc = (size_t)g.ncx * g.ncy * g.ncz;
    size_t occ_words = (size_t)g.ncy * g.ncz * g.cw;
    CK(h->cur->d_cell_count.ensure(ncells + 1, true));
    CK(h->cur->d_cell_start.ensure(ncells + 1));
    // [8 per-frame scalars][cull-grid bitmap][row bitmap] share one buffer: one memset per solute molecule
    CK(h->cur->d_occ.ensure(4 + occ_words + ((size_t)g.ny * g.nz * g.rw)));
    CK(h->cur->d_edt_xy.ensure(ncc)); CK(h->cur->d_lbd2.ensure(ncc));
    size_t rowmask_words = (size_t)g.ny * g.nz * g.rw;
    u64 *occ_p = h->cur->d_occ.p + 4, *rowmask_p = occ_p + occ_words;
    size_t nqc = (size_t)g.nqx * g.nqy * g.nqz;
    CK(h->cur->d_qcell_count.ensure(nqc + 1, true)); CK(h->cur->d_qcell_start.ensure(nqc + 1));
    {   // tile array: atoms + padding of each cell's last tile
        size_t maxq = std::max<size_t>(h->nv_atoms, (size_t)(c.coordination_number_only ? 0 : h->sample_chunk) * h->nv_atoms);
        CK(h->cur->d_qsorted.ensure(maxq + 32 * nqc));
    }
    const int nrand = c.coordination_number_only ? 0 : c.n_random_samples;
    int *sc = (int *)h->cur->d_occ.p;
    for (int isolute = 0; isolute < c.solute_nmols; ++isolute) {
        const float *xs = d_solute + (size_t)3 * ns_apm * isolute;
        const int skip = c.autocorrelation ? isolute : -1;
        int nrand_k = 0;
        if (c.solute_nmols == 1) nrand_k = nrand;
        else for (int s = 0; s < nrand; ++s) nrand_k += (ref_solute_host(h, frame, (uint32_t)s) == isolute);
        CK(cudaMemsetAsync(h->cur->d_occ.p, 0, (4 + occ_words + rowmask_words) * sizeof(u64), h->cur->stream));
        int tb = 128;
        launch(h, k_solute_bin<false>, dim3((ns_apm + tb - 1) / tb), dim3(tb), g, xs, ns_apm, h->cur->d_cell_count.p,
               (const int *)nullptr, occ_p, rowmask_p, (float4 *)nullptr);
        if (ncells + 1 <= CMX_SCAN_SMALL_MAX)
            launch(h, k_scan_small<IdentityOp>, dim3(1), dim3(CMX_SCAN_THREADS), (const int *)h->cur->d_cell_count.p, h->cur->d_cell_start.p, (int)(ncells + 1), IdentityOp());
        else {
            size_t tmp_bytes = h->cur->d_cub_tmp.n;
            CK(cub::DeviceScan::ExclusiveSum(h->cur->d_cub_tmp.p, tmp_bytes, h->cur->d_cell_count.p, h->cur->d_cell_start.p, (int)(ncells + 1), h->cur->stream));
            h->stats.kernel_launches += 2;
        }
        launch(h, k_solute_bin<true>, dim3((ns_apm + tb - 1) / tb), dim3(tb), g, xs, ns_apm, h->cur->d_cell_count.p,
               (const int *)h->cur->d_cell_start.p, occ_p, rowmask_p, h->cur->d_sorted.p);
        launch(h, k_edt_xy, dim3((unsigned)((ncc + 255) / 256)), dim3(256), g, (const u64 *)occ_p, h->cur->d_edt_xy.p);
        launch(h, k_edt_z, dim3((unsigned)((ncc + 255) / 256)), dim3(256), g, (const unsigned short *)h->cur->d_edt_xy.p, h->cur->d_lbd2.p);
        launch(h, k_filter_real, dim3((nv_mols + 127) / 128), dim3(128), g, h->P, d_solvent, skip,
               (const float *)h->cur->d_lbd2.p, h->cur->d_list.p, h->cur->d_worklist.p, sc + 0, sc + 5);
        if (h->stats.frames < 64 || (h->stats.frames & 15) == 0)   // host-side bound for the NEXT frames' cull window (monotone)
            CK(cudaMemcpyAsync(h->cur->h_scalars + 5, sc + 5, sizeof(int), cudaMemcpyDeviceToHost, h->cur->stream));
        launch(h, k_gen_real, dim3(h->num_sms * 4), dim3(256), g, h->P, d_solvent, (const float *)h->cur->d_lbd2.p,
               (const int *)h->cur->d_worklist.p, (const int *)(sc + 0), h->cur->d_qpos.p, h->cur->d_res.p, h->cur->d_qcell_count.p);
        { int rc = search_phase<false>(h, g, xs, d_solvent, h->cur->d_worklist.p, sc + 0, h->nv_atoms, h->cur->d_list.p,
                                       h->cur->d_def_real.p, h->cur->d_def_real_info.p, sc + 2, 0, 0); if (rc) return rc; }
        launch(h, k_resolve, dim3(h->num_sms * 2), dim3(CMX_RESOLVE_THREADS), g, h->P, frame, xs, d_solvent, (const float4 *)h->cur->d_sorted.p,
               (const int *)h->cur->d_cell_start.p, (int)ncells, (const int *)h->cur->d_bulk_idx.p, (const int *)(sc + 4),
               (const u64 *)h->cur->d_def_real.p, (const float2 *)h->cur->d_def_real_info.p, (const int *)(sc + 2), h->cur->d_list.p, (MdRec *)nullptr,
               nrand_k == 0 ? h->d_stats.p : (u64 *)nullptr, (const int *)(sc + 2), (const int *)nullptr);
        if (c.keep_lists)
            CK(cudaMemcpyAsync(h->d_list_all.p + (size_t)isolute * nv_mols, h->cur->d_list.p, sizeof(MdRec) * (size_t)nv_mols,
                               cudaMemcpyDeviceToDevice, h->cur->stream));
        if (nrand_k == 0) continue;
        // bulk list of this solute molecule, ascending molecule index (src/mddf.jl:406-415)
        size_t tmp_bytes2 = h->cur->d_cub_tmp.n;
        BulkPred pred{(const MdRec *)h->cur->d_list.p, skip, h->P.usecutoff, h->P.dbulk};
        CK(cub::DeviceSelect::If(h->cur->d_cub_tmp.p, tmp_bytes2, cub::CountingInputIterator<int>(0), h->cur->d_bulk_idx.p, sc + 4,
                                 nv_mols, pred, h->cur->stream));
        h->stats.kernel_launches += 2;
        // the random phase runs over chunks of samples so that the scratch (query atoms of the surviving
        // random molecules) stays bounded for any n_random_samples; normally one chunk
        for (int s0 = 0; s0 < nrand; s0 += h->sample_chunk) {
            const int s1 = std::min(nrand, s0 + h->sample_chunk);
            if (s0 > 0) { CK(cudaMemsetAsync(sc + 1, 0, sizeof(int), h->cur->stream)); CK(cudaMemsetAsync(sc + 3, 0, sizeof(int), h->cur->stream));
                          CK(cudaMemsetAsync(sc + 7, 0, sizeof(int), h->cur->stream)); }
            launch(h, k_filter_rand, dim3((unsigned)((nv_mols + 255) / 256), (unsigned)std::min(s1 - s0, 65535)), dim3(256), g, h->P, frame, isolute, skip,
                   s0, s1, (const float *)h->cur->d_lbd2.p, (const int *)(sc + 5), h->cur->d_rand_worklist.p, sc + 1);
            launch(h, k_gen_rand, dim3(h->num_sms * 8), dim3(128), g, h->P, frame, s0, d_solvent, (const float *)h->cur->d_lbd2.p,
                   (const int *)h->cur->d_rand_worklist.p, (const int *)(sc + 1), (const int *)h->cur->d_bulk_idx.p, (const int *)(sc + 4),
                   h->cur->d_qpos.p, h->cur->d_xexact.p, h->cur->d_res.p, h->cur->d_qcell_count.p);
            { int rc = search_phase<true>(h, g, xs, d_solvent, h->cur->d_rand_worklist.p, sc + 1, (size_t)(s1 - s0) * h->nv_atoms,
                                          c.keep_lists ? h->d_rand_list.p : (MdRec *)nullptr, h->cur->d_def_rand.p, h->cur->d_def_rand_info.p,
                                          sc + 3, 1, s0); if (rc) return rc; }
            launch(h, k_resolve, dim3(h->num_sms * 2), dim3(CMX_RESOLVE_THREADS), g, h->P, frame, xs, d_solvent, (const float4 *)h->cur->d_sorted.p,
                   (const int *)h->cur->d_cell_start.p, (int)ncells, (const int *)h->cur->d_bulk_idx.p, (const int *)(sc + 4),
                   (const u64 *)h->cur->d_def_rand.p, (const float2 *)h->cur->d_def_rand_info.p, (const int *)(sc + 3), (MdRec *)nullptr,
                   c.keep_lists ? h->d_rand_list.p : (MdRec *)nullptr, h->d_stats.p, (const int *)(s0 == 0 ? sc + 2 : nullptr), (const int *)(sc + 3));
        }
    }
    return CMX_OK;
}

__global__ void k_check_overflow(const int *flag, int *sticky) {
    if (threadIdx.x == 0 && *flag) *sticky = 1;
}

__global__ void k_emit(const u64 *cnt, const double *acc, double *out, size_t n, size_t half_lo, size_t half_hi, double w) {
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    double s = (k >= half_lo && k < half_hi) ? w / 2 : w;   // src/update_counters.jl:52-53
    out[k] = (acc ? acc[k] : 0.0) + s * (double)cnt[k];
}

__global__ void k_fold(const u64 *cnt, double *acc, size_t n, size_t half_lo, size_t half_hi, double w) {
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    double s = (k >= half_lo && k < half_hi) ? w / 2 : w;
    acc[k] += s * (double)cnt[k];
}

int sync_all(cmx_handle *h) {
    CK(cudaStreamSynchronize(h->s_copy));
    for (FrameCtx *x : h->ctx) CK(cudaStreamSynchronize(x->stream));
    return CMX_OK;
}

int fold_weight(cmx_handle *h) {
    { int rc = sync_all(h); if (rc) return rc; }   // every stream adds into the same integer block
    // acc += w * cnt ; cnt = 0  (the counters are sums of frame weights, src/update_counters.jl:47,60)
    if (!h->acc_used) { CK(h->d_acc.ensure(h->cnt_len, true)); CK(cudaMemsetAsync(h->d_acc.p, 0, sizeof(double) * h->cnt_len, h->cur->stream)); h->acc_used = true; }
    size_t nb = h->nbins, lo = 4 * nb, hi = 4 * nb + 2 * nb * h->cfg.n_groups_solute;
    if (!h->cfg.autocorrelation) lo = hi = 0;
    launch(h, k_fold, dim3((unsigned)((h->cnt_len + 255) / 256)), dim3(256), (const u64 *)h->d_cnt.p, h->d_acc.p, h->cnt_len, lo, hi, h->cur_weight);
    CK(cudaMemsetAsync(h->d_cnt.p, 0, sizeof(u64) * h->cnt_len, h->cur->stream));
    CK(cudaStreamSynchronize(h->cur->stream));
    return CMX_OK;
}

int submit_common(cmx_handle *h, const float *d_solute, const float *d_solvent, int64_t frame_index, double weight,
                  const double cell[9]) {
    if (!(weight > 0) && weight != 0) return fail(h, CMX_ERR_ARG, "frame weight must be finite and non-negative");
    if (weight == 0) return fail(h, CMX_ERR_ARG, "zero-weight frames must be skipped by the caller (src/mddf.jl:102)");
    Geom g;
    int rc = build_geom(h, cell, g);
    if (rc) return rc;
    h->cur = h->ctx[(size_t)(h->submitted++ % (int64_t)(h->active_ctx > 0 ? h->active_ctx : (int)h->ctx.size()))];
    if (h->have_weight && weight != h->cur_weight) { rc = fold_weight(h); if (rc) return rc; }
    h->cur_weight = weight; h->have_weight = true;
    // rmax feedback from earlier frames (pinned mirror, may lag)
    float seen = 0.f;
    for (FrameCtx *x : h->ctx) { float v = 0.f; std::memcpy(&v, x->h_scalars + 5, sizeof(float)); seen = std::max(seen, v); }
    if (seen > 0.f && seen * 1.25f + 0.1f > h->rmax_bound && seen > h->rmax_bound * 0.999f)
        h->rmax_bound = std::max(h->rmax_bound, seen * 1.25f + 0.1f);
    g.rmax_bound = h->rmax_bound;
    g.dwin = std::min(15, (int)std::ceil((h->cut_eff + g.tau + h->rmax_bound + 1e-3) / h->cside) + 2);
    // the transform marks everything at >= (dwin-1) cells as "far": only valid while that exceeds the thresholds
    if ((g.dwin - 1) * h->cside < h->cut_eff + g.tau + h->rmax_bound + 1e-3) g.rmax_bound = -1.f;   // random cull disabled
    if (!h->ev_first_set) { CK(cudaEventRecord(h->ev_first, h->cur->stream)); h->ev_first_set = true; }
    uint32_t frame = (uint32_t)(frame_index & 0xffffffffll);
    if (h->cfg.keep_lists && h->d_rand_list.p) CK(cudaMemsetAsync(h->d_rand_list.p, 0, sizeof(MdRec) * h->d_rand_list.n, h->cur->stream));
    h->last_g = g; h->last_dsol = d_solute; h->last_dsolv = d_solvent;
    rc = h->path == 1 ? frame_grid_path(h, d_solute, d_solvent, frame, g)
                      : frame_pair_path(h, d_solute, d_solvent, frame, g);
    if (rc) return rc;
    CK(cudaGetLastError());
    // update_volume!, src/mddf.jl:350-352
    const double *a = cell, *b = cell + 3, *c = cell + 6;
    double axb[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
    h->volume_total += weight * (axb[0] * c[0] + axb[1] * c[1] + axb[2] * c[2]);
    h->sum_weights += weight;
    h->stats.frames++;
    h->last_frame = frame_index;
    return CMX_OK;
}

}  // namespace

#include "cmx_pairs_host.inl"

// ==================================================================================================
// C ABI
// ==================================================================================================
extern "C" {

const char *cmx_version(void) { return "cmx_b200 0.1.0 (sm_100a)"; }

const char *cmx_last_error(cmx_handle *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int32_t cmx_destroy(cmx_handle *h) {
    if (!h) return CMX_OK;
    cudaSetDevice(h->device);
    for (FrameCtx *x : h->ctx) if (x->stream) cudaStreamSynchronize(x->stream);
    if (h->s_copy) cudaStreamSynchronize(h->s_copy);
    for (auto &s : h->ring) {
        if (s.h_in) cudaFreeHost(s.h_in);
        if (s.d_in) cudaFree(s.d_in);
        if (s.h2d_done) cudaEventDestroy(s.h2d_done);
        if (s.consumed) cudaEventDestroy(s.consumed);
    }
    h->d_sol_off.release(); h->d_sol_ids.release(); h->d_solv_off.release(); h->d_solv_ids.release();
    h->d_cnt.release(); h->d_acc.release(); h->d_emit.release(); h->d_rand_list.release(); h->d_list_all.release(); h->d_stats.release();
    feed_destroy(h);
    for (FrameCtx *x : h->ctx) { h->cur = x; pairs_release(h); x->release(); delete x; }
    h->ctx.clear(); h->cur = nullptr;
    for (auto &p : h->prof_events) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
    if (h->ev_first) cudaEventDestroy(h->ev_first);
    if (h->ev_last) cudaEventDestroy(h->ev_last);
    if (h->s_copy) cudaStreamDestroy(h->s_copy);
    delete h;
    return CMX_OK;
}

static int create_impl(cmx_handle *h, const cmx_config *cfg) {
    const cmx_config &c = *cfg;
    if (c.struct_size != (int32_t)sizeof(cmx_config)) return fail(h, CMX_ERR_ARG, "cmx_config.struct_size mismatch (ABI)");
    if (c.solute_nmols < 1 || c.solute_natomspermol < 1 || c.solvent_nmols < 1 || c.solvent_natomspermol < 1)
        return fail(h, CMX_ERR_ARG, "selections must have at least one molecule and one atom per molecule");
    if (c.irefatom < 1 || c.irefatom > c.solvent_natomspermol)
        return fail(h, CMX_ERR_ARG, "in MDDF options: Reference atom index is greater than number of atoms of the solvent molecule.");
    if (!(c.binstep > 0) || !(c.cutoff > 0) || !(c.dbulk > 0)) return fail(h, CMX_ERR_ARG, "binstep, cutoff and dbulk must be positive");
    if (c.usecutoff && c.dbulk >= c.cutoff) return fail(h, CMX_ERR_ARG, "in MDDF options: The bulk volume is zero (dbulk must be smaller than cutoff).");
    if (c.n_random_samples < 1 && !c.coordination_number_only) return fail(h, CMX_ERR_ARG, "in MDDF options: n_random_samples must be greater than 0.");
    if (c.autocorrelation && (c.solute_nmols != c.solvent_nmols || c.solute_natomspermol != c.solvent_natomspermol))
        return fail(h, CMX_ERR_ARG, "autocorrelation requires identical solute and solvent selections");
    if (c.n_groups_solute < 1 || c.n_groups_solvent < 1) return fail(h, CMX_ERR_ARG, "n_groups_* must be positive");
    if (!c.solute_group_offsets && c.n_groups_solute != c.solute_natomspermol) return fail(h, CMX_ERR_ARG, "n_groups_solute must equal solute_natomspermol without custom groups");
    if (!c.solvent_group_offsets && c.n_groups_solvent != c.solvent_natomspermol) return fail(h, CMX_ERR_ARG, "n_groups_solvent must equal solvent_natomspermol without custom groups");
    if (c.n_random_samples > 100000000) return fail(h, CMX_ERR_ARG, "n_random_samples is limited to 1e8");
    h->cfg = c;
    h->device = c.device;
    CK(cudaSetDevice(c.device));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, c.device));
    h->num_sms = prop.multiProcessorCount;
    {
        int b0 = 0, b1 = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b0, k_tile_search<false>, 256, 0));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b1, k_tile_search<true>, 256, 0));
        h->search_grid[0] = h->num_sms * std::max(1, b0); h->search_grid[1] = h->num_sms * std::max(1, b1);
        if (const char *e = std::getenv("CMX_SEARCH_BLOCKS_PER_SM")) h->search_blocks_env = std::max(1, atoi(e));   // experiments only
    }
    h->nbins = std::max(1, (int)std::ceil(c.cutoff / c.binstep));   // setbin(cutoff, binstep), src/results.jl:131
    h->cut_eff = c.usecutoff ? c.cutoff : c.dbulk;                   // src/minimum_distances.jl:168
    h->ns_atoms = (size_t)c.solute_nmols * c.solute_natomspermol;
    h->nv_atoms = (size_t)c.solvent_nmols * c.solvent_natomspermol;
    h->in_floats = 3 * (c.autocorrelation ? h->nv_atoms : h->ns_atoms + h->nv_atoms);
    h->path = c.path ? c.path : ((c.solute_nmols == 1 || c.solute_natomspermol > 64) ? 1 : 2);
    if (h->path != 1 && h->path != 2) return fail(h, CMX_ERR_ARG, "path must be 0 (auto), 1 (grid) or 2 (molecule pairs)");
    // (cmx_config.group_lanes is accepted for ABI stability and ignored: the search works on 32-query tiles)
    // search grid: a "row" is a (y,z) column of cells cut/3 wide; along x the cells are ~2.5 A so that
    // a row is scanned over a tight x-span.  Query atoms are tiled in cubic cells of ~5 A (a dozen
    // solvent atoms, one warp).  Values from a sweep on C2 (rows cut/2..cut/6, tiles 4.5..8 A): +10 % over
    // cut/4 rows with 6 A tiles; fewer, fatter rows cost pair evaluations but save per-row bookkeeping.
    h->Kdiv = 3;
    h->side = (h->cut_eff + 0.02) / h->Kdiv;
    h->sidex = (h->cut_eff + 0.02) / std::max(2, (int)std::lround(h->cut_eff / 2.5));
    h->qside = std::min(8.0, std::max(5.0, h->cut_eff / 3.0));
    // cull grid (distance transform): cut/5
    h->cside = (h->cut_eff + 0.02) / 5.0;
    // tuning overrides (experiments only)
    if (const char *e = std::getenv("CMX_ROWDIV")) { h->Kdiv = std::max(1, atoi(e)); h->side = (h->cut_eff + 0.02) / h->Kdiv; }
    if (const char *e = std::getenv("CMX_XSIDE")) h->sidex = std::max(0.5, atof(e));
    if (const char *e = std::getenv("CMX_QSIDE")) h->qside = std::max(1.0, atof(e));
    if (const char *e = std::getenv("CMX_CULLDIV")) h->cside = (h->cut_eff + 0.02) / std::max(1.0, atof(e));
    CK(cudaStreamCreateWithFlags(&h->s_copy, cudaStreamNonBlocking));
    int nctx = c.n_streams > 0 ? c.n_streams : (h->nv_atoms + h->ns_atoms > 2000000 ? 4 : 8);
    if (c.keep_lists) nctx = 1;                  // the parity hooks read the scratch of the last frame
    if (nctx > 16) return fail(h, CMX_ERR_ARG, "n_streams must be <= 16");
    for (int k = 0; k < nctx; ++k) {
        FrameCtx *x = new FrameCtx();
        h->ctx.push_back(x);
        CK(cudaStreamCreateWithFlags(&x->stream, cudaStreamNonBlocking));
        CK(cudaEventCreate(&x->ev_end));
        CK(cudaHostAlloc(&x->h_scalars, sizeof(int) * 8, cudaHostAllocDefault));
        std::memset(x->h_scalars, 0, sizeof(int) * 8);
    }
    h->cur = h->ctx[0];
    CK(cudaEventCreate(&h->ev_first)); CK(cudaEventCreate(&h->ev_last));
    int slots = c.ring_slots > 0 ? c.ring_slots : 3;
    h->ring.resize(slots);
    h->numa_node = gpu_numa_node(c.device);
    NumaPrefer numa_guard(h->numa_node);
    for (auto &s : h->ring) {
        CK(cudaHostAlloc(&s.h_in, sizeof(float) * h->in_floats, cudaHostAllocDefault));
        CK(cudaMalloc(&s.d_in, sizeof(float) * h->in_floats));
        CK(cudaEventCreateWithFlags(&s.h2d_done, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&s.consumed, cudaEventDisableTiming));
    }
    // problem description
    Prob &P = h->P;
    P.ns_mols = c.solute_nmols; P.ns_apm = c.solute_natomspermol; P.nv_mols = c.solvent_nmols; P.nv_apm = c.solvent_natomspermol;
    P.autocorr = c.autocorrelation; P.iref = c.irefatom - 1; P.usecutoff = c.usecutoff; P.nbins = h->nbins;
    P.nrand = c.coordination_number_only ? 0 : c.n_random_samples; P.cn_only = c.coordination_number_only;
    P.ng_sol = c.n_groups_solute; P.ng_solv = c.n_groups_solvent;
    P.custom_sol = c.solute_group_offsets != nullptr; P.custom_solv = c.solvent_group_offsets != nullptr;
    P.cutoff = c.cutoff; P.dbulk = c.dbulk; P.binstep = c.binstep;
    P.seed_lo = (uint32_t)(c.seed & 0xffffffffull); P.seed_hi = (uint32_t)(c.seed >> 32);
    auto upload_csr = [&](const int32_t *off, const int32_t *ids, size_t npos, DevBuf<int> &doff, DevBuf<int> &dids) -> cudaError_t {
        if (!off) return cudaSuccess;
        size_t nids = (size_t)off[npos];
        cudaError_t e = doff.ensure(npos + 1); if (e != cudaSuccess) return e;
        e = dids.ensure(std::max<size_t>(nids, 1)); if (e != cudaSuccess) return e;
        e = cudaMemcpy(doff.p, off, sizeof(int) * (npos + 1), cudaMemcpyHostToDevice); if (e != cudaSuccess) return e;
        if (nids) e = cudaMemcpy(dids.p, ids, sizeof(int) * nids, cudaMemcpyHostToDevice);
        return e;
    };
    if (P.custom_sol) {
        for (size_t k = 0; k < (size_t)c.solute_group_offsets[h->ns_atoms]; ++k)
            if (c.solute_group_ids[k] < 0 || c.solute_group_ids[k] >= c.n_groups_solute) return fail(h, CMX_ERR_ARG, "solute group id out of range");
        CK(upload_csr(c.solute_group_offsets, c.solute_group_ids, h->ns_atoms, h->d_sol_off, h->d_sol_ids));
    }
    if (P.custom_solv) {
        for (size_t k = 0; k < (size_t)c.solvent_group_offsets[h->nv_atoms]; ++k)
            if (c.solvent_group_ids[k] < 0 || c.solvent_group_ids[k] >= c.n_groups_solvent) return fail(h, CMX_ERR_ARG, "solvent group id out of range");
        CK(upload_csr(c.solvent_group_offsets, c.solvent_group_ids, h->nv_atoms, h->d_solv_off, h->d_solv_ids));
    }
    P.sol_off = h->d_sol_off.p; P.sol_ids = h->d_sol_ids.p; P.solv_off = h->d_solv_off.p; P.solv_ids = h->d_solv_ids.p;
    size_t nb = h->nbins;
    h->cnt_len = nb * (4 + 2 * (size_t)c.n_groups_solute + 2 * (size_t)c.n_groups_solvent);
    CK(h->d_cnt.ensure(h->cnt_len)); CK(cudaMemset(h->d_cnt.p, 0, sizeof(u64) * h->d_cnt.n));
    u64 *q = h->d_cnt.p;
    P.md = q; q += nb; P.md_r = q; q += nb; P.rdf = q; q += nb; P.rdf_r = q; q += nb;
    P.gsol = q; q += nb * c.n_groups_solute; P.gsol_r = q; q += nb * c.n_groups_solute;
    P.gsolv = q; q += nb * c.n_groups_solvent; P.gsolv_r = q;
    // scratch common to both paths
    size_t nvm = c.solvent_nmols, nrand = (size_t)P.nrand;
    // random phase in chunks of samples: at most ~48 M query atoms of scratch per frame context
    h->sample_chunk = (int)std::max<size_t>(1, std::min<size_t>(std::max<size_t>(nrand, 1), (size_t)(48.0e6 / (double)h->nv_atoms)));
    const size_t nchunk = (size_t)h->sample_chunk;
    CK(h->d_stats.ensure(8, true));
    for (FrameCtx *x_ : h->ctx) {
        h->cur = x_;
        CK(h->cur->d_scalars.ensure(16, true)); 
        CK(h->cur->d_list.ensure(nvm));
        CK(h->cur->d_bulk_idx.ensure(nvm));
        CK(h->cur->d_worklist.ensure(nvm)); CK(h->cur->d_rand_worklist.ensure(std::max<size_t>(nchunk * nvm, 1)));
        CK(h->cur->d_def_real.ensure(nvm)); CK(h->cur->d_def_rand.ensure(std::max<size_t>(nchunk * nvm, 1)));
    CK(h->cur->d_def_real_info.ensure(nvm)); CK(h->cur->d_def_rand_info.ensure(std::max<size_t>(nchunk * nvm, 1)));
        if (c.keep_lists) {
            if (h->path == 1) CK(h->d_list_all.ensure((size_t)c.solute_nmols * nvm));
            CK(h->d_rand_list.ensure(std::max<size_t>(nrand * nvm, 1)));
        }
        if (h->path == 1) {
            CK(h->cur->d_sorted.ensure(27 * (size_t)c.solute_natomspermol));
            size_t maxq = std::max<size_t>(h->nv_atoms, nchunk * h->nv_atoms);
            CK(h->cur->d_qpos.ensure(maxq)); CK(h->cur->d_qsorted.ensure(maxq)); CK(h->cur->d_res.ensure(maxq));
            CK(h->cur->d_xexact.ensure(std::max<size_t>(3 * nchunk * h->nv_atoms, 1)));
        } else {
            int rc = pairs_create(h); if (rc) return rc;
        }
    }
    h->cur = h->ctx[0];
    // cub temp storage sized for the largest call we make
    size_t t1 = 0, t2 = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, t1, (int *)nullptr, (int *)nullptr, 1 << 28);
    cub::DeviceSelect::If(nullptr, t2, cub::CountingInputIterator<int>(0), (int *)nullptr, (int *)nullptr, (int)std::max<size_t>(nvm, 1),
                          BulkPred{nullptr, -1, 0, 0.0});
    for (FrameCtx *x_ : h->ctx) CK(x_->d_cub_tmp.ensure(std::max(t1, t2) + 1024));
    CK(cudaDeviceSynchronize());
    return CMX_OK;
}

int32_t cmx_create(const cmx_config *cfg, cmx_handle **out) {
    if (!cfg || !out) { g_create_error = "cmx_create: null argument"; return CMX_ERR_ARG; }
    cmx_handle *h = new cmx_handle();
    int rc = create_impl(h, cfg);
    if (rc) { g_create_error = h->err; cmx_destroy(h); *out = nullptr; return rc; }
    *out = h;
    return CMX_OK;
}

int32_t cmx_acquire_frame_buffer(cmx_handle *h, float **solute_xyz, float **solvent_xyz) {
    if (!h) return CMX_ERR_ARG;
    CK(cudaSetDevice(h->device));
    if (h->acquired >= 0) return fail(h, CMX_ERR_STATE, "cmx_acquire_frame_buffer: previous slot not submitted");
    Slot &s = h->ring[h->next_slot];
    if (s.in_flight) { CK(cudaEventSynchronize(s.consumed)); s.in_flight = false; }
    h->acquired = h->next_slot;
    h->next_slot = (h->next_slot + 1) % (int)h->ring.size();
    if (h->cfg.autocorrelation) { if (solute_xyz) *solute_xyz = s.h_in; if (solvent_xyz) *solvent_xyz = s.h_in; }
    else { if (solute_xyz) *solute_xyz = s.h_in; if (solvent_xyz) *solvent_xyz = s.h_in + 3 * h->ns_atoms; }
    return CMX_OK;
}

int32_t cmx_submit_frame(cmx_handle *h, int64_t frame_index, double weight, const double cell[9]) {
    if (!h || !cell) return CMX_ERR_ARG;
    CK(cudaSetDevice(h->device));
    if (h->acquired < 0) return fail(h, CMX_ERR_STATE, "cmx_submit_frame: no frame buffer acquired");
    Slot &s = h->ring[h->acquired];
    h->acquired = -1;
    CK(cudaMemcpyAsync(s.d_in, s.h_in, sizeof(float) * h->in_floats, cudaMemcpyHostToDevice, h->s_copy));
    CK(cudaEventRecord(s.h2d_done, h->s_copy));
    FrameCtx *next = h->ctx[(size_t)(h->submitted % (int64_t)(h->active_ctx > 0 ? h->active_ctx : (int)h->ctx.size()))];
    CK(cudaStreamWaitEvent(next->stream, s.h2d_done, 0));
    h->stats.h2d_bytes += (int64_t)(sizeof(float) * h->in_floats);
    const float *dsol = s.d_in, *dsolv = h->cfg.autocorrelation ? s.d_in : s.d_in + 3 * h->ns_atoms;
    int rc = submit_common(h, dsol, dsolv, frame_index, weight, cell);
    // the slot is reusable once the frame's kernels are done (even on error, to keep the ring consistent)
    cudaEventRecord(s.consumed, next->stream);
    s.in_flight = true;
    return rc;
}

int32_t cmx_submit_frame_device(cmx_handle *h, const float *d_solute_xyz, const float *d_solvent_xyz, int64_t frame_index,
                                double weight, const double cell[9]) {
    if (!h || !cell || !d_solvent_xyz) return CMX_ERR_ARG;
    CK(cudaSetDevice(h->device));
    const float *dsol = h->cfg.autocorrelation ? d_solvent_xyz : d_solute_xyz;
    if (!dsol) return fail(h, CMX_ERR_ARG, "cmx_submit_frame_device: null solute pointer");
    return submit_common(h, dsol, d_solvent_xyz, frame_index, weight, cell);
}

int32_t cmx_sync(cmx_handle *h) {
    if (!h) return CMX_ERR_ARG;
    CK(cudaSetDevice(h->device));
    if (h->ev_first_set) for (FrameCtx *x : h->ctx) CK(cudaEventRecord(x->ev_end, x->stream));
    { int rc = sync_all(h); if (rc) return rc; }
    if (h->ev_first_set) {
        float best = 0;
        for (FrameCtx *x : h->ctx) { float ms = 0; if (cudaEventElapsedTime(&ms, h->ev_first, x->ev_end) == cudaSuccess) best = std::max(best, ms); }
        (void)cudaGetLastError();
        h->stats.gpu_ms_total += best; h->ev_first_set = false;
    }
    prof_collect(h);
    for (auto &s : h->ring) s.in_flight = false;
    int sticky = 0;
    for (FrameCtx *x : h->ctx) { int v = 0; CK(cudaMemcpy(&v, x->d_scalars.p + 8, sizeof(int), cudaMemcpyDeviceToHost)); sticky |= v; }
    if (sticky) return fail(h, CMX_ERR_STATE, "deferred-pair buffer overflow: too many exactly tied / cutoff-edge pairs in one frame");
    return CMX_OK;
}

int32_t cmx_counters_device(cmx_handle *h, void **device_ptr, int64_t *n_uint64) {
    if (!h || !device_ptr || !n_uint64) return CMX_ERR_ARG;
    if (h->acc_used) return fail(h, CMX_ERR_STATE, "cmx_counters_device: frame weights varied; integer counters were folded to fp64 (use cmx_finish per GPU and sum)");
    *device_ptr = h->d_cnt.p; *n_uint64 = (int64_t)h->cnt_len;
    return CMX_OK;
}

int32_t cmx_finish(cmx_handle *h, cmx_counters *out) {
    if (!h || !out) return CMX_ERR_ARG;
    int rc = cmx_sync(h); if (rc) return rc;
    size_t n = h->cnt_len, nb = h->nbins;
    const double w = h->have_weight ? h->cur_weight : 1.0;
    const size_t gs = nb * h->cfg.n_groups_solute, gv = nb * h->cfg.n_groups_solvent;
    size_t lo = 4 * nb, hi = 4 * nb + 2 * gs;
    if (!h->cfg.autocorrelation) lo = hi = 0;
    CK(h->d_emit.ensure(n));
    launch(h, k_emit, dim3((unsigned)((n + 255) / 256)), dim3(256), (const u64 *)h->d_cnt.p,
           (const double *)(h->acc_used ? h->d_acc.p : nullptr), h->d_emit.p, n, lo, hi, w);
    CK(cudaStreamSynchronize(h->cur->stream));
    auto emit = [&](double *dst, size_t off, size_t len) -> cudaError_t {
        if (!dst || !len) return cudaSuccess;
        return cudaMemcpy(dst, h->d_emit.p + off, sizeof(double) * len, cudaMemcpyDeviceToHost);
    };
    CK(emit(out->md_count, 0, nb)); CK(emit(out->md_count_random, nb, nb));
    CK(emit(out->rdf_count, 2 * nb, nb)); CK(emit(out->rdf_count_random, 3 * nb, nb));
    CK(emit(out->solute_group_count, 4 * nb, gs)); CK(emit(out->solute_group_count_random, 4 * nb + gs, gs));
    CK(emit(out->solvent_group_count, 4 * nb + 2 * gs, gv)); CK(emit(out->solvent_group_count_random, 4 * nb + 2 * gs + gv, gv));
    out->nbins = h->nbins; out->n_groups_solute = h->cfg.n_groups_solute; out->n_groups_solvent = h->cfg.n_groups_solvent;
    out->volume_total = h->volume_total; out->sum_weights = h->sum_weights;
    return CMX_OK;
}

static void md_to_abi(const MdRec &e, cmx_md &o) {
    o.within_cutoff = e.flags & 1; o.ref_atom_within_cutoff = (e.flags >> 1) & 1;
    o.i = (e.flags & 1) ? e.i + 1 : 0; o.j = (e.flags & 1) ? e.j + 1 : 0;
    o.d = (e.flags & 1) ? e.d : INFINITY; o.d_ref_atom = (e.flags & 2) ? e.dref : INFINITY;
}

int32_t cmx_read_minimum_distances(cmx_handle *h, int32_t isolute, cmx_md *out) {
    if (!h || !out) return CMX_ERR_ARG;
    if (!h->cfg.keep_lists) return fail(h, CMX_ERR_STATE, "cmx_read_minimum_distances needs cmx_config.keep_lists = 1");
    if (isolute < 0 || isolute >= h->cfg.solute_nmols) return fail(h, CMX_ERR_ARG, "isolute out of range");
    int rc = cmx_sync(h); if (rc) return rc;
    size_t nvm = h->cfg.solvent_nmols;
    std::vector<MdRec> tmp(nvm);
    if (h->path == 2) {
        // molecule-pair path keeps no per-solute lists: recompute the exact list of this solute
        // molecule from the last frame (still resident in its staging slot)
        if (!h->last_dsolv) return fail(h, CMX_ERR_STATE, "no frame submitted yet");
        PairGeom pg = make_pair_geom(h, h->last_g);
        launch(h, k_ref_lists, dim3((unsigned)((nvm + 127) / 128), 1), dim3(128), h->last_g, pg, h->P, (uint32_t)h->last_frame, (int)isolute, 0,
               h->last_dsol, h->last_dsolv, h->cur->pairs.sol, h->cur->pairs.solv, h->cur->d_list.p);
        CK(cudaStreamSynchronize(h->cur->stream));
        CK(cudaMemcpy(tmp.data(), h->cur->d_list.p, sizeof(MdRec) * nvm, cudaMemcpyDeviceToHost));
    } else
    CK(cudaMemcpy(tmp.data(), h->d_list_all.p + (size_t)isolute * nvm, sizeof(MdRec) * nvm, cudaMemcpyDeviceToHost));
    for (size_t m = 0; m < nvm; ++m) md_to_abi(tmp[m], out[m]);
    return CMX_OK;
}

int32_t cmx_read_random_minimum_distances(cmx_handle *h, int32_t sample, cmx_md *out) {
    if (!h || !out) return CMX_ERR_ARG;
    if (!h->cfg.keep_lists) return fail(h, CMX_ERR_STATE, "cmx_read_random_minimum_distances needs cmx_config.keep_lists = 1");
    if (sample < 0 || sample >= h->P.nrand) return fail(h, CMX_ERR_ARG, "sample out of range");
    int rc = cmx_sync(h); if (rc) return rc;
    size_t nvm = h->cfg.solvent_nmols;
    std::vector<MdRec> tmp(nvm);
    CK(cudaMemcpy(tmp.data(), h->d_rand_list.p + (size_t)sample * nvm, sizeof(MdRec) * nvm, cudaMemcpyDeviceToHost));
    for (size_t m = 0; m < nvm; ++m) md_to_abi(tmp[m], out[m]);
    return CMX_OK;
}

int32_t cmx_alloc_pinned(void **ptr, int64_t bytes) {
    if (!ptr || bytes <= 0) return CMX_ERR_ARG;
    return cudaHostAlloc(ptr, (size_t)bytes, cudaHostAllocDefault) == cudaSuccess ? CMX_OK : CMX_ERR_CUDA;
}
int32_t cmx_free_pinned(void *ptr) { return (!ptr || cudaFreeHost(ptr) == cudaSuccess) ? CMX_OK : CMX_ERR_CUDA; }

int32_t cmx_get_stats(cmx_handle *h, cmx_stats *out) {
    if (!h || !out) return CMX_ERR_ARG;
    int rc = cmx_sync(h); if (rc) return rc;
    u64 st[8];
    CK(cudaMemcpy(st, h->d_stats.p, sizeof st, cudaMemcpyDeviceToHost));
    h->stats.pair_evals = (int64_t)st[0]; h->stats.deferred = (int64_t)st[1];
    *out = h->stats;
    return CMX_OK;
}

int32_t cmx_reset(cmx_handle *h) {
    if (!h) return CMX_ERR_ARG;
    int rc = cmx_sync(h); if (rc) return rc;
    CK(cudaMemset(h->d_cnt.p, 0, sizeof(u64) * h->d_cnt.n));
    if (h->acc_used) CK(cudaMemset(h->d_acc.p, 0, sizeof(double) * h->d_acc.n));
    CK(cudaMemset(h->d_stats.p, 0, sizeof(u64) * 8));
    h->acc_used = false; h->have_weight = false; h->cur_weight = 1.0;
    h->volume_total = 0; h->sum_weights = 0;
    h->stats = cmx_stats{};
    return CMX_OK;
}

int32_t cmx_set_option(cmx_handle *h, const char *name, double value) {
    if (!h || !name) return CMX_ERR_ARG;
    std::string n(name);
    if (n == "count_pairs") h->count_pairs = value != 0;
    else if (n == "profile") h->profile = value != 0;
    else if (n == "sample_chunk") {   // testing knob: smaller chunks of random samples per pass (never larger than allocated)
        int rc = cmx_sync(h); if (rc) return rc;
        int v = (int)value;
        if (v < 1) return fail(h, CMX_ERR_ARG, "sample_chunk must be >= 1");
        h->sample_chunk = std::min(h->sample_chunk, v);
        for (FrameCtx *x : h->ctx) x->pairs.sample_chunk = std::max(1, std::min(x->pairs.sample_chunk, v));
    }
    else if (n == "active_streams") {
        int rc = cmx_sync(h); if (rc) return rc;
        int v = (int)value;
        if (v < 0 || v > (int)h->ctx.size()) return fail(h, CMX_ERR_ARG, "active_streams out of range");
        h->active_ctx = v;
    }
    else if (n == "group_lanes") { /* accepted, ignored */ }
    else return fail(h, CMX_ERR_ARG, "unknown option: " + n);
    return CMX_OK;
}

}  // extern "C"

#include "cmx_feed.inl"
#include "cmx_xtc.inl"
