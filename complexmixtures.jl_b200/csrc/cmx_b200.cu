// cmx_b200.cu -- C-ABI (include/cmx_b200.h) and host-side frame pipeline of libcmx_b200.so.
//
// Host responsibilities restated from the reference's chunk task (src/mddf.jl:288-337):
// per-handle state (build_particle_system / Buffer / Result -> cmx_create), frame staging
// (pinned ring, async H2D), per-frame kernel sequence (mddf_frame!, src/mddf.jl:361-429),
// frame-weight handling and the final counters (sum!, src/results.jl:629-649).
#include <cuda_runtime.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <cctype>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
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
    // grow-only; a fresh allocation can be zeroed ON `stream` (ordered before the kernels that will use it)
    cudaError_t ensure(size_t want, bool zero = false, cudaStream_t stream = nullptr) {
        if (want <= n && p) return cudaSuccess;
        if (p) { cudaError_t e = cudaFree(p); if (e != cudaSuccess) return e; p = nullptr; n = 0; }
        size_t cap = want + want / 4 + 64;
        cudaError_t e = cudaMalloc(&p, cap * sizeof(T));
        if (e != cudaSuccess) return e;
        n = cap;
        if (zero) return cudaMemsetAsync(p, 0, cap * sizeof(T), stream);
        return cudaSuccess;
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
};

}  // namespace

struct cmx_feed;   // native DCD feed ring + group-reduction scratch (cmx_feed.inl)

// Device scratch of ONE frame in flight on the grid path (the pointers of a GridFrame descriptor).  The buffers of all
// slots of a batch context are carved out of ONE device allocation (the context's arena): a handle makes a handful of
// cudaMalloc / cudaFree calls instead of a thousand -- those calls, not the kernels, were the cost of a short run.
struct GridSlot {
    int *sc = nullptr, *cell_count = nullptr, *cell_start = nullptr, *qcell_count = nullptr, *qcell_start = nullptr, *worklist = nullptr,
        *rand_worklist = nullptr, *bulk_idx = nullptr;
    float4 *sorted = nullptr, *qpos = nullptr, *qsorted = nullptr, *res = nullptr;
    u64 *bits = nullptr, *def_real = nullptr, *def_rand = nullptr, *scan_state = nullptr;
    float2 *def_real_info = nullptr, *def_rand_info = nullptr;
    double *xexact = nullptr;
    unsigned short *edt_xy = nullptr;
    float *lbd2 = nullptr;
    MdRec *list = nullptr;
    unsigned char *tile_valid = nullptr;
};

// sizes (elements) of the geometry-dependent buffers an arena was laid out for
struct ArenaDims { size_t ncells = 0, ncull = 0, nqc = 0, bits = 0; };

// A frame (x one solute molecule) that was submitted and waits for its batch to be launched.
struct PendingFrame {
    Geom g;
    const float *xs, *xv;
    uint32_t frame;
    int isolute, skip_mol, nrand_k;
    double weight;
};

// One batch of frames in flight: a compute stream, the frame slots of the batch and their descriptors.  Frames are
// collected until the batch is full (or a sync / weight change / ring wrap-around forces it out), then every kernel of
// the per-frame sequence is launched ONCE for the whole batch.  Several batch contexts are in flight on their own streams.
struct FrameCtx {
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_end = nullptr, ev_fd = nullptr;
    bool fd_busy = false;
    std::vector<GridSlot> slots;
    unsigned char *arena = nullptr; size_t arena_bytes = 0;   // ONE device allocation behind all slots of this context
    ArenaDims dims;
    GridFrame *h_fd = nullptr, *d_fd = nullptr;     // pinned / device descriptor arrays [batch]
    std::vector<PendingFrame> pending;
    std::vector<cudaEvent_t> release_events;        // recorded on `stream` once the pending frames' kernels are enqueued
    // molecule-pair path (one frame per context at a time) and shared odds and ends
    DevBuf<MdRec> d_list;
    DevBuf<int> d_scalars;                          // [8] sticky overflow flag of the pair path
    PairScratch pairs;
    int *h_scalars = nullptr;                       // pinned mirror (rmax feedback)
    void release() {
        slots.clear();
        if (arena) cudaFree(arena);
        arena = nullptr; arena_bytes = 0;
        d_list.release(); d_scalars.release();
        if (h_fd) cudaFreeHost(h_fd);
        if (d_fd) cudaFree(d_fd);
        if (h_scalars) cudaFreeHost(h_scalars);
        if (ev_end) cudaEventDestroy(ev_end);
        if (ev_fd) cudaEventDestroy(ev_fd);
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
    double side = 0, sidex = 0, cside = 0, qside = 0, ring_width = 4.5;   // (ring: measured with the x-limited rings in the random phase, C4 / C2 frames/s: 2.5 A 3813 / 9567, 3.5 A 4040 / 10618, 4.5 A 4133 / 10815, 5.5 A worse)
    cudaStream_t s_copy = nullptr;
    size_t hist_smem = 0;           // bytes of the shared-memory histograms (HistPriv) of the counting kernels
    int batch = 1;                  // frames per batch (grid path); 1 on the molecule-pair path
    int fill = 0;                   // index of the batch context being filled
    uint32_t scan_epoch = 0;        // launch number of the chained scans (tags their tile states)
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
    bool emit_valid = false;        // d_emit holds the (possibly all-reduced) f64 counters that cmx_finish writes out
    // per-frame scratch lives in FrameCtx (one per compute stream)
    DevBuf<MdRec> d_rand_list, d_list_all;   // parity hooks (keep_lists => one stream)
    DevBuf<u64> d_stats;            // [0] pair_evals, [1] deferred total
    std::vector<FrameCtx *> ctx;    // batches are dealt round-robin to the contexts; kernels of different batches overlap
    FrameCtx *cur = nullptr;        // context whose kernels are being enqueued (launch / prof_* use its stream)
    int active_ctx = 0;             // contexts actually used (option "active_streams"; 0 = all)
    float rmax_bound = 0.f;
    int sample_chunk = 1;           // samples of the random phase per pass (grid path)
    // bookkeeping
    double w0 = 1.0; bool have_weight = false;   // weight of the first frame: the integer counters are hits at this weight
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
    double grid_scale = 1.0;           // experiments only (CMX_GRID_SCALE): scales the block caps of the latency-bound grid-stride kernels
    cmx_feed *feed = nullptr;
    // group handle (cmx_group.inl): one child per device; a group owns no device state of its own
    std::vector<cmx_handle *> children;
    int next_child = 0;
    cmx_handle *acquired_child = nullptr;
    // CMX_TRACE=skip:count -- device timeline of `count` batches after `skip` flushes (events between the launches), printed at sync
    int trace_skip = -1, trace_count = 0; bool tracing = false;
    std::vector<std::pair<const char *, cudaEvent_t>> trace_events;
    bool xtc_host_decode = false;   // option "xtc_host_decode": decode XTC frames in the reader threads instead of on the device
    bool sync_destroy = false;      // option "sync_destroy": free everything on the caller's thread
    bool poll_stop_file = true, stopped_by_file = false;   // native feed: the reference's cooperative stop file
    int numa_node = -1;                        // NUMA node of the GPU (-1 unknown): pinned staging memory is placed there
};

namespace {

#define CK_G(call)                                                                                 \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            g->err = std::string(#call) + ": " + cudaGetErrorString(e_);                           \
            return CMX_ERR_CUDA;                                                                   \
        }                                                                                          \
    } while (0)

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            h->err = std::string(#call) + ": " + cudaGetErrorString(e_);                           \
            return CMX_ERR_CUDA;                                                                   \
        }                                                                                          \
    } while (0)

int fail(cmx_handle *h, int code, const std::string &msg) { h->err = msg; return code; }

double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
// cudaEventSynchronize with the blocked time booked as back-pressure (cmx_stats.host_wait_ms)
cudaError_t wait_event(cmx_handle *h, cudaEvent_t e) {
    if (cudaEventQuery(e) == cudaSuccess) return cudaSuccess;
    (void)cudaGetLastError();
    const double t0 = now_ms();
    cudaError_t r = cudaEventSynchronize(e);
    h->stats.host_wait_ms += now_ms() - t0;
    return r;
}
struct SubmitTimer {   // cmx_stats.host_submit_ms
    cmx_handle *h; double t0;
    explicit SubmitTimer(cmx_handle *h_) : h(h_), t0(now_ms()) {}
    ~SubmitTimer() { h->stats.host_submit_ms += now_ms() - t0; }
};

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
    g.ring = (float)std::max(h->ring_width, 1.01 * h->sidex + 0.01);   // a ring must reach at least one more x cell of a partly swept row (k_tile_search)
    double margin = h->cut_eff + tau + 0.05;
    for (int k = 0; k < 3; ++k) {
        g.elo[k] = lo[k] - margin; g.ehi[k] = hi[k] + margin; g.ctr[k] = 0.5 * (lo[k] + hi[k]);
        g.gmin[k] = (float)(g.elo[k] - g.ctr[k]);
    }
    g.side = (float)h->side; g.inv_side = (float)(1.0 / h->side);
    g.sidex = (float)h->sidex; g.inv_sidex = (float)(1.0 / h->sidex);
    g.cut_hi2 = g.cut_hi * g.cut_hi * (1.0f + 1e-6f);
    g.nx = (int)std::ceil((g.ehi[0] - g.elo[0]) / h->sidex) + 1;
    if (g.nx > 65535) return fail(h, CMX_ERR_CELL, "unit cell too long along x for the search grid (more than 65535 cells of the x spacing)");
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

void trace_mark(cmx_handle *h, const char *name) {
    if (!h->tracing) return;
    cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, h->cur->stream);
    h->trace_events.push_back({name, e});
}
void trace_report(cmx_handle *h) {
    if (h->trace_events.empty()) return;
    std::vector<std::pair<std::string, double>> agg;
    double total = 0;
    for (size_t k = 1; k < h->trace_events.size(); ++k) {
        float ms = 0;
        const char *name = h->trace_events[k].first;
        if (std::string(name) == "begin") continue;
        if (cudaEventElapsedTime(&ms, h->trace_events[k - 1].second, h->trace_events[k].second) != cudaSuccess) continue;
        total += ms;
        bool found = false;
        for (auto &a : agg) if (a.first == name) { a.second += ms; found = true; }
        if (!found) agg.push_back({name, (double)ms});
    }
    (void)cudaGetLastError();
    std::fprintf(stderr, "[cmx trace] %zu marks, %.1f us in total (interval = previous mark -> this mark, launch gaps included)\n", h->trace_events.size(), total * 1e3);
    for (auto &a : agg) std::fprintf(stderr, "[cmx trace] %10.1f us %5.1f%%  %s\n", a.second * 1e3, 100.0 * a.second / std::max(total, 1e-12), a.first.c_str());
    for (auto &e : h->trace_events) cudaEventDestroy(e.second);
    h->trace_events.clear();
}

// number of contexts frames are dealt to (option "active_streams" restricts it for single-stream measurements)
int nctx_active(const cmx_handle *h) { return h->active_ctx > 0 ? h->active_ctx : (int)h->ctx.size(); }

template <class K, class... Args>
void launch_y(cmx_handle *h, K kernel, unsigned gx, unsigned nb, unsigned threads, Args... args) {
    kernel<<<dim3(std::max(gx, 1u), nb), dim3(threads), 0, h->cur->stream>>>(args...);
    h->stats.kernel_launches++;
}

// tile the query atoms of the batch's work lists, search, combine (one phase: real or random)
template <bool RANDOM>
int search_phase(cmx_handle *h, const GridFrame *fd, unsigned nb, size_t max_atoms, size_t nqc_max, int s0) {
    const unsigned sms = (unsigned)h->num_sms;
    launch_y(h, k_chain_scan<ScanTiles<RANDOM>>, (unsigned)std::min<size_t>((nqc_max + CMX_SCAN_TILE) / CMX_SCAN_TILE, sms * 2), nb, CMX_SCAN_THREADS,
             fd, h->P, ++h->scan_epoch);
    trace_mark(h, RANDOM ? "scan_tiles<rand>" : "scan_tiles<real>");
    launch_y(h, k_qscatter<RANDOM>, (unsigned)std::min<size_t>((max_atoms + 255) / 256, (size_t)std::max(1.0, sms * 8 * h->grid_scale)), nb, 256u, fd, h->P);
    trace_mark(h, RANDOM ? "qscatter<rand>" : "qscatter<real>");
    cudaEvent_t pe = prof_begin(h, RANDOM ? 1 : 0);
    u64 *pev = h->count_pairs ? h->d_stats.p : nullptr;
    // The tiles of the whole batch are one queue.  Each of the n batches in flight gets 1/n of the resident block
    // slots of an SM: the searches of several batches then share the SMs with the small kernels of the others
    // (measured on C2 in round 1: 5 blocks/SM 8.2 k frames/s, 2 -> 8.9 k, 1 -> 9.4 k with 8 frames in flight).
    const int nfly = nctx_active(h);
    const int k = pev ? 1 : 0;
    int per_sm = h->search_blocks_env > 0 ? h->search_blocks_env : std::max(1, (h->search_grid[k] / h->num_sms + nfly - 1) / nfly);
    const unsigned grid = sms * (unsigned)per_sm;
    if (pev) k_tile_search<true, RANDOM><<<grid, CMX_SEARCH_WARPS * 32, CMX_SEARCH_SMEM, h->cur->stream>>>(fd, (int)nb, pev);
    else k_tile_search<false, RANDOM><<<grid, CMX_SEARCH_WARPS * 32, CMX_SEARCH_SMEM, h->cur->stream>>>(fd, (int)nb, pev);
    h->stats.kernel_launches++;
    prof_end(h, pe);
    trace_mark(h, RANDOM ? "tile_search<rand>" : "tile_search<real>");
    {   // persistent blocks (their private histograms are flushed once)
        unsigned gfin = (unsigned)std::max(1.0, sms * 4 * h->grid_scale);   // (measured on C4: 74 blocks 270 us per batch, 296 -> 90, 592 -> 70, 1184 -> 72; the flush is not the limit)
        if (const char *e = std::getenv("CMX_FIN_BLOCKS")) gfin = (unsigned)std::max(1, atoi(e));   // experiments only
        k_finalise<RANDOM><<<gfin, CMX_FIN_THREADS, h->hist_smem, h->cur->stream>>>(fd, (int)nb, h->P, s0);
        h->stats.kernel_launches++;
    }
    trace_mark(h, RANDOM ? "finalise<rand>" : "finalise<real>");
    return CMX_OK;
}

// ---- launch the pending frames of one batch context (mddf_frame!, src/mddf.jl:361-429, for every frame of the batch) ----
int flush_ctx(cmx_handle *h, FrameCtx *x) {
    if (x->pending.empty()) return CMX_OK;
    const cmx_config &c = h->cfg;
    h->cur = x;
    const unsigned nb = (unsigned)x->pending.size();
    const unsigned sms = (unsigned)h->num_sms;
    const int ns_apm = c.solute_natomspermol, nv_mols = c.solvent_nmols;
    const int nrand = c.coordination_number_only ? 0 : c.n_random_samples;
    if (x->fd_busy) { CK(wait_event(h, x->ev_fd)); x->fd_busy = false; }   // previous descriptor upload of this context
    size_t ncells_max = 0, ncull_max = 0, nqc_max = 0, bits_max = 0;
    bool any_random = false;
    for (unsigned k = 0; k < nb; ++k) {
        const Geom &g = x->pending[k].g;
        ncells_max = std::max(ncells_max, (size_t)g.nx * g.ny * g.nz); ncull_max = std::max(ncull_max, (size_t)g.ncx * g.ncy * g.ncz);
        nqc_max = std::max(nqc_max, (size_t)g.nqx * g.nqy * g.nqz);
        bits_max = std::max(bits_max, (size_t)g.ncy * g.ncz * g.cw + (size_t)g.ny * g.nz * g.rw);
    }
    {   // the context's arena: every buffer of every slot, laid out for the largest grids seen (+25 % when it has to grow)
        const size_t nvm = (size_t)nv_mols, nchunk = (size_t)(nrand ? h->sample_chunk : 0);
        const size_t maxq = std::max<size_t>(h->nv_atoms, nchunk * h->nv_atoms), nrw = std::max<size_t>(nchunk * nvm, 1);
        ArenaDims &D = x->dims;
        const bool grow = !x->arena || ncells_max > D.ncells || ncull_max > D.ncull || nqc_max > D.nqc || bits_max > D.bits;
        if (grow) {
            auto up = [](size_t have, size_t want) { return want > have ? want + want / 4 + 64 : have; };
            D.ncells = up(D.ncells, ncells_max); D.ncull = up(D.ncull, ncull_max); D.nqc = up(D.nqc, nqc_max); D.bits = up(D.bits, bits_max);
        }
        const size_t nscan = std::max(std::max(D.ncells, D.nqc) + 1, nvm) / CMX_SCAN_TILE + 2;
        size_t off = 0;
        auto carve = [&](size_t bytes) { const size_t at = off; off += (bytes + 255) & ~(size_t)255; return at; };
        struct Lay { size_t sc, cell_count, cell_start, qcell_count, qcell_start, worklist, rand_worklist, bulk_idx, sorted, qpos, qsorted, res, bits,
                            def_real, def_rand, scan_state, def_real_info, def_rand_info, xexact, edt_xy, lbd2, list, tile_valid; } L;
        // (the buffers that must start as zeros come first: one memset covers them)
        L.sc = carve(sizeof(int) * SC_COUNT); L.cell_count = carve(sizeof(int) * (D.ncells + 1)); L.qcell_count = carve(sizeof(int) * (D.nqc + 1));
        L.scan_state = carve(sizeof(u64) * nscan);
        const size_t zero_bytes = off;
        L.cell_start = carve(sizeof(int) * (D.ncells + 1)); L.qcell_start = carve(sizeof(int) * (D.nqc + 1));
        L.worklist = carve(sizeof(int) * nvm); L.rand_worklist = carve(sizeof(int) * nrw); L.bulk_idx = carve(sizeof(int) * nvm);
        L.sorted = carve(sizeof(float4) * 27 * (size_t)c.solute_natomspermol);
        L.qpos = carve(sizeof(float4) * maxq); L.qsorted = carve(sizeof(float4) * (maxq + 32 * D.nqc)); L.res = carve(sizeof(float4) * maxq);
        L.bits = carve(sizeof(u64) * D.bits);
        L.def_real = carve(sizeof(u64) * nvm); L.def_rand = carve(sizeof(u64) * nrw);
        L.def_real_info = carve(sizeof(float2) * nvm); L.def_rand_info = carve(sizeof(float2) * nrw);
        L.xexact = carve(sizeof(double) * std::max<size_t>(3 * nchunk * h->nv_atoms, 1));
        L.edt_xy = carve(sizeof(unsigned short) * D.ncull); L.lbd2 = carve(sizeof(float) * D.ncull);
        L.list = carve(sizeof(MdRec) * nvm); L.tile_valid = carve(maxq / 32 + D.nqc + 1);
        const size_t slot_bytes = off;
        if (grow) {
            CK(cudaStreamSynchronize(x->stream));             // the previous batch of this context still reads the old arena
            if (x->arena) { CK(cudaFree(x->arena)); x->arena = nullptr; }
            x->arena_bytes = slot_bytes * x->slots.size();
            CK(cudaMalloc(&x->arena, x->arena_bytes));
            for (size_t k = 0; k < x->slots.size(); ++k) {
                unsigned char *base = x->arena + k * slot_bytes;
                CK(cudaMemsetAsync(base, 0, zero_bytes, x->stream));
                GridSlot &S = x->slots[k];
                S.sc = (int *)(base + L.sc); S.cell_count = (int *)(base + L.cell_count); S.cell_start = (int *)(base + L.cell_start);
                S.qcell_count = (int *)(base + L.qcell_count); S.qcell_start = (int *)(base + L.qcell_start);
                S.worklist = (int *)(base + L.worklist); S.rand_worklist = (int *)(base + L.rand_worklist); S.bulk_idx = (int *)(base + L.bulk_idx);
                S.sorted = (float4 *)(base + L.sorted); S.qpos = (float4 *)(base + L.qpos); S.qsorted = (float4 *)(base + L.qsorted); S.res = (float4 *)(base + L.res);
                S.bits = (u64 *)(base + L.bits); S.def_real = (u64 *)(base + L.def_real); S.def_rand = (u64 *)(base + L.def_rand);
                S.scan_state = (u64 *)(base + L.scan_state);
                S.def_real_info = (float2 *)(base + L.def_real_info); S.def_rand_info = (float2 *)(base + L.def_rand_info);
                S.xexact = (double *)(base + L.xexact); S.edt_xy = (unsigned short *)(base + L.edt_xy); S.lbd2 = (float *)(base + L.lbd2);
                S.list = (MdRec *)(base + L.list); S.tile_valid = base + L.tile_valid;
            }
        }
    }
    for (unsigned k = 0; k < nb; ++k) {
        const PendingFrame &pf = x->pending[k];
        const Geom &g = pf.g;
        GridSlot &S = x->slots[k];
        const size_t ncells = (size_t)g.nx * g.ny * g.nz, ncc = (size_t)g.ncx * g.ncy * g.ncz, nqc = (size_t)g.nqx * g.nqy * g.nqz;
        const size_t occ_words = (size_t)g.ncy * g.ncz * g.cw, row_words = (size_t)g.ny * g.nz * g.rw;
        GridFrame &F = x->h_fd[k];
        F.g = g; F.xs = pf.xs; F.xv = pf.xv;
        F.sc = S.sc; F.bits = S.bits; F.occ = S.bits; F.rowmask = S.bits + occ_words;
        F.cell_count = S.cell_count; F.cell_start = S.cell_start; F.sorted = S.sorted;
        F.edt_xy = S.edt_xy; F.lbd2 = S.lbd2;
        F.qpos = S.qpos; F.qsorted = S.qsorted; F.res = S.res; F.xexact = S.xexact;
        F.qcell_count = S.qcell_count; F.qcell_start = S.qcell_start; F.tile_valid = S.tile_valid;
        F.list = S.list; F.rand_list = c.keep_lists ? h->d_rand_list.p : nullptr;
        F.worklist = S.worklist; F.rand_worklist = S.rand_worklist; F.bulk_idx = S.bulk_idx;
        F.def_real = S.def_real; F.def_rand = S.def_rand; F.def_real_info = S.def_real_info; F.def_rand_info = S.def_rand_info;
        F.scan_state = S.scan_state;
        F.bits_words = (long long)(occ_words + row_words);
        F.ncells = (int)ncells; F.nqcells = (int)nqc; F.ncull = (int)ncc;
        F.frame = pf.frame; F.isolute = pf.isolute; F.skip_mol = pf.skip_mol; F.nrand_k = pf.nrand_k; F.pad0 = 0; F.weight = pf.weight; F.pad1 = 0;
        (void)ncells; (void)ncc; (void)nqc;
        any_random |= pf.nrand_k > 0;
    }
    CK(cudaMemcpyAsync(x->d_fd, x->h_fd, sizeof(GridFrame) * nb, cudaMemcpyHostToDevice, x->stream));
    CK(cudaEventRecord(x->ev_fd, x->stream)); x->fd_busy = true;
    if (!h->ev_first_set) { CK(cudaEventRecord(h->ev_first, x->stream)); h->ev_first_set = true; }
    const GridFrame *fd = x->d_fd;
    if (h->trace_skip >= 0) { if (h->trace_skip-- == 0) h->tracing = true; }
    if (h->tracing) { if (h->trace_count-- <= 0) h->tracing = false; }
    trace_mark(h, "begin");
    if (c.keep_lists && h->d_rand_list.p && x->pending[0].isolute == 0)   // (keep_lists: one frame x one solute molecule per batch)
        CK(cudaMemsetAsync(h->d_rand_list.p, 0, sizeof(MdRec) * h->d_rand_list.n, x->stream));
    // ---- solute grid + cull grid
    launch_y(h, k_zero_frame, (unsigned)std::min<size_t>((bits_max + 255) / 256, 64), nb, 256u, fd);
    trace_mark(h, "zero_frame");
    launch_y(h, k_solute_bin<false>, (unsigned)((ns_apm + 127) / 128), nb, 128u, fd, ns_apm);
    trace_mark(h, "solute_bin<count>");
    launch_y(h, k_chain_scan<ScanCells>, (unsigned)std::min<size_t>((ncells_max + CMX_SCAN_TILE) / CMX_SCAN_TILE, sms * 2), nb, CMX_SCAN_THREADS,
             fd, h->P, ++h->scan_epoch);
    trace_mark(h, "scan_cells");
    launch_y(h, k_solute_bin<true>, (unsigned)((ns_apm + 127) / 128), nb, 128u, fd, ns_apm);
    trace_mark(h, "solute_bin<scatter>");
    const unsigned gcull = (unsigned)std::min<size_t>((ncull_max + 255) / 256, (size_t)std::max(1.0, sms * 16 * h->grid_scale));
    launch_y(h, k_edt_xy, gcull, nb, 256u, fd);
    trace_mark(h, "edt_xy");
    launch_y(h, k_edt_z, gcull, nb, 256u, fd);
    trace_mark(h, "edt_z");
    // ---- real phase
    launch_y(h, k_filter_real, (unsigned)((nv_mols + 127) / 128), nb, 128u, fd, h->P);
    trace_mark(h, "filter_real");
    if (h->stats.frames < 64 || (h->stats.frames & 15) < (int64_t)nb)   // host-side bound for the NEXT frames' cull window (monotone)
        CK(cudaMemcpyAsync(x->h_scalars + 5, x->slots[0].sc + SC_RMAX, sizeof(int), cudaMemcpyDeviceToHost, x->stream));
    launch_y(h, k_gen_real, (unsigned)std::min<size_t>((h->nv_atoms + 255) / 256, (size_t)std::max(1.0, sms * 4 * h->grid_scale)), nb, 256u, fd, h->P);
    trace_mark(h, "gen_real");
    { int rc = search_phase<false>(h, fd, nb, h->nv_atoms, nqc_max, 0); if (rc) return rc; }
    launch_y(h, k_resolve<false>, std::max(16u, sms * 2 / nb), nb, (unsigned)CMX_RESOLVE_THREADS, fd, h->P, h->d_stats.p);
    trace_mark(h, "resolve<real>");
    if (c.keep_lists)
        for (unsigned k = 0; k < nb; ++k)
            CK(cudaMemcpyAsync(h->d_list_all.p + (size_t)x->pending[k].isolute * nv_mols, x->slots[k].list, sizeof(MdRec) * (size_t)nv_mols,
                               cudaMemcpyDeviceToDevice, x->stream));
    if (any_random) {
        // bulk list of every frame, ascending molecule index (src/mddf.jl:406-415): an ordered compaction
        launch_y(h, k_chain_scan<ScanBulk>, (unsigned)std::min<size_t>(((size_t)nv_mols + CMX_SCAN_TILE - 1) / CMX_SCAN_TILE, sms * 2), nb, CMX_SCAN_THREADS,
                 fd, h->P, ++h->scan_epoch);
        trace_mark(h, "scan_bulk");
        // the random phase runs over chunks of samples so that the scratch (query atoms of the surviving
        // random molecules) stays bounded for any n_random_samples; normally one chunk
        for (int s0 = 0; s0 < nrand; s0 += h->sample_chunk) {
            const int s1 = std::min(nrand, s0 + h->sample_chunk);
            if (s0 > 0) { k_reset_rand<<<nb, 32, 0, x->stream>>>(fd); h->stats.kernel_launches++; }
            k_filter_rand<<<dim3((unsigned)((nv_mols + CMX_FRAND_THREADS - 1) / CMX_FRAND_THREADS), nb, (unsigned)std::min((s1 - s0 + 31) / 32, 4096)),
                            CMX_FRAND_THREADS, 0, x->stream>>>(fd, h->P, s0, s1);
            h->stats.kernel_launches++;
            trace_mark(h, "filter_rand");
            const size_t max_items = (size_t)(s1 - s0) * (size_t)nv_mols;
            launch_y(h, k_gen_rand, (unsigned)std::min<size_t>((max_items + 127) / 128, (size_t)std::max(1.0, sms * 8 * h->grid_scale)), nb, 128u, fd, h->P, s0);
    trace_mark(h, "gen_rand");
            { int rc = search_phase<true>(h, fd, nb, (size_t)(s1 - s0) * h->nv_atoms, nqc_max, s0); if (rc) return rc; }
            launch_y(h, k_resolve<true>, std::max(16u, sms * 2 / nb), nb, (unsigned)CMX_RESOLVE_THREADS, fd, h->P, h->d_stats.p);
    trace_mark(h, "resolve<rand>");
        }
    }
    for (cudaEvent_t e : x->release_events) CK(cudaEventRecord(e, x->stream));
    x->release_events.clear();
    x->pending.clear();
    h->stats.batches++;
    CK(cudaGetLastError());
    return CMX_OK;
}

int flush_all(cmx_handle *h) {
    for (FrameCtx *x : h->ctx) { int rc = flush_ctx(h, x); if (rc) return rc; }
    return CMX_OK;
}

// an event that is recorded when a pending frame's kernels are enqueued: make sure that has happened
int flush_if_pending(cmx_handle *h, cudaEvent_t e) {
    for (FrameCtx *x : h->ctx)
        for (cudaEvent_t p : x->release_events)
            if (p == e) return flush_ctx(h, x);
    return CMX_OK;
}

// ---- one frame on the grid path: queue it (x each solute molecule) into the batch being filled ----------------------
int queue_grid_frame(cmx_handle *h, const float *d_solute, const float *d_solvent, uint32_t frame, const Geom &g, double weight, cudaEvent_t release) {
    const cmx_config &c = h->cfg;
    const int nrand = c.coordination_number_only ? 0 : c.n_random_samples;
    for (int isolute = 0; isolute < c.solute_nmols; ++isolute) {
        FrameCtx *x = h->ctx[(size_t)h->fill];
        PendingFrame pf;
        pf.g = g; pf.xs = d_solute + (size_t)3 * c.solute_natomspermol * isolute; pf.xv = d_solvent; pf.frame = frame;
        pf.isolute = isolute; pf.skip_mol = c.autocorrelation ? isolute : -1; pf.weight = weight;
        pf.nrand_k = 0;
        if (c.solute_nmols == 1) pf.nrand_k = nrand;
        else for (int s = 0; s < nrand; ++s) pf.nrand_k += (ref_solute_host(h, frame, (uint32_t)s) == isolute);
        x->pending.push_back(pf);
        if (release && isolute == c.solute_nmols - 1) x->release_events.push_back(release);
        if ((int)x->pending.size() >= h->batch) {
            // a frame's solute molecules may straddle batches: its buffers are released with the LAST of them
            int rc = flush_ctx(h, x); if (rc) return rc;
            h->fill = (h->fill + 1) % nctx_active(h);
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

int sync_all(cmx_handle *h) {
    { int rc = flush_all(h); if (rc) return rc; }
    CK(cudaStreamSynchronize(h->s_copy));
    for (FrameCtx *x : h->ctx) CK(cudaStreamSynchronize(x->stream));
    return CMX_OK;
}

// A frame whose weight differs from the first weight of the run: from here on the kernels add w (fp64 atomics) into
// the fp64 twin of the accumulator block instead of counting integer hits.  The integer block keeps the hits of the
// frames before the switch (all of weight w0) and is scaled once, at the end.  Nothing is synchronised.
int enter_acc_mode(cmx_handle *h) {
    if (h->acc_used) return CMX_OK;
    { int rc = flush_all(h); if (rc) return rc; }     // the pending frames were queued for the integer counters
    CK(h->d_acc.ensure(h->cnt_len));
    // every compute stream may add into it: zero it before any of them continues
    CK(cudaMemsetAsync(h->d_acc.p, 0, sizeof(double) * h->d_acc.n, h->ctx[0]->stream));
    CK(cudaStreamSynchronize(h->ctx[0]->stream));
    h->P.acc = h->d_acc.p;
    h->acc_used = true;
    return CMX_OK;
}

int submit_common(cmx_handle *h, const float *d_solute, const float *d_solvent, int64_t frame_index, double weight,
                  const double cell[9], cudaEvent_t release) {
    if (!(weight > 0) && weight != 0) return fail(h, CMX_ERR_ARG, "frame weight must be finite and non-negative");
    if (weight == 0) return fail(h, CMX_ERR_ARG, "zero-weight frames must be skipped by the caller (src/mddf.jl:102)");
    Geom g;
    int rc = build_geom(h, cell, g);
    if (rc) return rc;
    h->emit_valid = false;
    if (!h->have_weight) { h->w0 = weight; h->have_weight = true; }
    if (weight != h->w0 && !h->acc_used) { rc = enter_acc_mode(h); if (rc) return rc; }
    // rmax feedback from earlier frames (pinned mirror, may lag)
    float seen = 0.f;
    for (FrameCtx *x : h->ctx) { float v = 0.f; std::memcpy(&v, x->h_scalars + 5, sizeof(float)); seen = std::max(seen, v); }
    if (seen > 0.f && seen * 1.25f + 0.1f > h->rmax_bound && seen > h->rmax_bound * 0.999f)
        h->rmax_bound = std::max(h->rmax_bound, seen * 1.25f + 0.1f);
    g.rmax_bound = h->rmax_bound;
    g.dwin = std::min(15, (int)std::ceil((h->cut_eff + g.tau + h->rmax_bound + 1e-3) / h->cside) + 2);
    // the transform marks everything at >= (dwin-1) cells as "far": only valid while that exceeds the thresholds
    if ((g.dwin - 1) * h->cside < h->cut_eff + g.tau + h->rmax_bound + 1e-3) g.rmax_bound = -1.f;   // random cull disabled
    uint32_t frame = (uint32_t)(frame_index & 0xffffffffll);
    h->last_g = g; h->last_dsol = d_solute; h->last_dsolv = d_solvent;
    if (h->path == 1) rc = queue_grid_frame(h, d_solute, d_solvent, frame, g, weight, release);
    else {
        h->cur = h->ctx[(size_t)h->fill];
        h->fill = (h->fill + 1) % nctx_active(h);
        if (!h->ev_first_set) { CK(cudaEventRecord(h->ev_first, h->cur->stream)); h->ev_first_set = true; }
        if (h->cfg.keep_lists && h->d_rand_list.p) CK(cudaMemsetAsync(h->d_rand_list.p, 0, sizeof(MdRec) * h->d_rand_list.n, h->cur->stream));
        h->P.w = weight;
        rc = frame_pair_path(h, d_solute, d_solvent, frame, g);
        if (rc == CMX_OK && release) CK(cudaEventRecord(release, h->cur->stream));
    }
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

extern "C" { static int create_impl(cmx_handle *h, const cmx_config *cfg); }
#include "cmx_group.inl"

// ==================================================================================================
// C ABI
// ==================================================================================================
extern "C" {

const char *cmx_version(void) { return "cmx_b200 0.1.0 (sm_100a)"; }

const char *cmx_last_error(cmx_handle *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

// Releasing a handle is dominated by driver calls that unpin / unmap memory (0.5 s for the C4 problem) and nobody waits
// for their result: the streams are drained on the caller's thread, the frees run on a detached reaper thread.  The next
// cmx_create waits for the reapers first, so that its allocations find the memory.
namespace {
std::mutex g_reaper_mu;
std::condition_variable g_reaper_cv;
int g_reapers_pending = 0;
void join_reapers() {
    std::unique_lock<std::mutex> lk(g_reaper_mu);
    g_reaper_cv.wait(lk, [] { return g_reapers_pending == 0; });
}
// a reaper must not outlive the CUDA runtime: the first one registers an exit handler (it runs BEFORE the runtime's own,
// which was registered earlier, at CUDA initialisation) that waits for the pending frees
void reap(std::function<void()> work) {
    static std::once_flag once;
    std::call_once(once, [] { std::atexit(join_reapers); });
    { std::lock_guard<std::mutex> lk(g_reaper_mu); ++g_reapers_pending; }
    std::thread([work] {
        work();
        std::lock_guard<std::mutex> lk(g_reaper_mu);
        --g_reapers_pending;
        g_reaper_cv.notify_all();
    }).detach();
}

void release_handle(cmx_handle *h) {
    cudaSetDevice(h->device);
    const double t_release = now_ms();
    const bool trace_release = std::getenv("CMX_TRACE") != nullptr;
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
    if (trace_release) std::fprintf(stderr, "[cmx trace] release: %8.1f ms (reaper thread)\n", now_ms() - t_release);
}
}  // namespace

int32_t cmx_destroy(cmx_handle *h) {
    if (!h) return CMX_OK;
    if (is_group(h)) {
        for (cmx_handle *c : h->children) cmx_destroy(c);
        delete h;
        return CMX_OK;
    }
    if (!h->s_copy && h->ctx.empty()) { delete h; return CMX_OK; }   // a group handle whose creation failed early
    cudaSetDevice(h->device);
    for (FrameCtx *x : h->ctx) if (x->stream) cudaStreamSynchronize(x->stream);
    if (h->s_copy) cudaStreamSynchronize(h->s_copy);
    if (h->sync_destroy) { release_handle(h); return CMX_OK; }
    reap([h] { release_handle(h); });      // nothing waits for it but the next cmx_create and the process exit
    return CMX_OK;
}

static int create_impl(cmx_handle *h, const cmx_config *cfg) {
    const cmx_config &c = *cfg;
    const bool trace_create = std::getenv("CMX_TRACE") != nullptr;
    double t_phase = now_ms();
    auto phase = [&](const char *what) {
        if (!trace_create) return;
        cudaDeviceSynchronize();
        const double t = now_ms();
        std::fprintf(stderr, "[cmx trace] create: %8.1f ms  %s\n", t - t_phase, what);
        t_phase = t;
    };
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
        CK(cudaFuncSetAttribute(k_tile_search<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, CMX_SEARCH_SMEM));
        CK(cudaFuncSetAttribute(k_tile_search<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, CMX_SEARCH_SMEM));
        CK(cudaFuncSetAttribute(k_tile_search<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, CMX_SEARCH_SMEM));
        CK(cudaFuncSetAttribute(k_tile_search<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, CMX_SEARCH_SMEM));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b0, k_tile_search<false, true>, CMX_SEARCH_WARPS * 32, CMX_SEARCH_SMEM));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b1, k_tile_search<true, true>, CMX_SEARCH_WARPS * 32, CMX_SEARCH_SMEM));
        h->search_grid[0] = h->num_sms * std::max(1, b0); h->search_grid[1] = h->num_sms * std::max(1, b1);
        if (const char *e = std::getenv("CMX_TRACE")) { int a = 0, b = 1; if (std::sscanf(e, "%d:%d", &a, &b) >= 1) { h->trace_skip = a; h->trace_count = b; } }
        if (const char *e = std::getenv("CMX_SEARCH_BLOCKS_PER_SM")) h->search_blocks_env = std::max(1, atoi(e));   // experiments only
        if (const char *e = std::getenv("CMX_GRID_SCALE")) h->grid_scale = std::max(0.01, atof(e));                  // experiments only
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
    if (const char *e = std::getenv("CMX_RING")) h->ring_width = std::max(0.0, atof(e));
    if (const char *e = std::getenv("CMX_QSIDE")) h->qside = std::max(1.0, atof(e));
    if (const char *e = std::getenv("CMX_CULLDIV")) h->cside = (h->cut_eff + 0.02) / std::max(1.0, atof(e));
    phase("device properties, kernel attributes");
    CK(cudaStreamCreateWithFlags(&h->s_copy, cudaStreamNonBlocking));
    // Frames per batch (grid path): enough frames per launch that the small kernels of the sequence fill the GPU and
    // the launch count per frame drops below 2; large systems fill the GPU on their own and their slots are big.
    const size_t natoms_in = h->nv_atoms + (c.autocorrelation ? 0 : h->ns_atoms);
    int batch = c.batch_frames > 0 ? c.batch_frames : (int)std::min<size_t>(16, std::max<size_t>(1, (size_t)4000000 / std::max<size_t>(natoms_in, 1)));   // (C4, 930 k atoms in: 4 frames per launch; 8 gave +3 % frames/s for twice the scratch and staging ring)
    if (const char *e = std::getenv("CMX_BATCH")) { int v = atoi(e); if (v > 0) batch = std::min(v, CMX_MAX_BATCH); }   // experiments only
    if (h->path == 2) batch = 1;
    int nctx = c.n_streams > 0 ? c.n_streams : (h->path == 2 ? (natoms_in > 2000000 ? 4 : 8) : (batch > 1 ? 3 : 4));
    if (c.keep_lists) { nctx = 1; batch = 1; }   // the parity hooks read the scratch of the last frame
    if (nctx > 16) return fail(h, CMX_ERR_ARG, "n_streams must be <= 16");
    if (batch > CMX_MAX_BATCH) return fail(h, CMX_ERR_ARG, "batch_frames must be <= 32");
    h->batch = batch;
    for (int k = 0; k < nctx; ++k) {
        FrameCtx *x = new FrameCtx();
        h->ctx.push_back(x);
        CK(cudaStreamCreateWithFlags(&x->stream, cudaStreamNonBlocking));
        CK(cudaEventCreate(&x->ev_end));
        CK(cudaEventCreateWithFlags(&x->ev_fd, cudaEventDisableTiming));
        CK(cudaHostAlloc(&x->h_scalars, sizeof(int) * 8, cudaHostAllocDefault));
        std::memset(x->h_scalars, 0, sizeof(int) * 8);
        CK(cudaHostAlloc(&x->h_fd, sizeof(GridFrame) * batch, cudaHostAllocDefault));
        CK(cudaMalloc(&x->d_fd, sizeof(GridFrame) * batch));
    }
    h->cur = h->ctx[0];
    CK(cudaEventCreate(&h->ev_first)); CK(cudaEventCreate(&h->ev_last));
    // staging ring of acquire/submit: by default the frames of every batch in flight plus two batches being staged
    // behind them (their H2D copies run while the older batches compute).  A slot's pinned and device buffers are
    // allocated the first time it is acquired: a short run never pays for the whole ring, a long one pins the later
    // slots while the first frames already compute.
    int slots = c.ring_slots > 0 ? c.ring_slots : std::max(4, batch * (nctx + 2) + 1);
    h->ring.resize(slots);
    h->numa_node = gpu_numa_node(c.device);
    phase("streams, descriptor arrays, staging ring (pinned + device)");
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
    {   // memory guard (the reference checks the Result copies against the RAM, src/parallel_setup.jl:29-54): ONE set of
        // counters per GPU here, plus its f64 image for cmx_finish, plus the scratch of the frames in flight
        size_t free_b = 0, total_b = 0;
        CK(cudaMemGetInfo(&free_b, &total_b));
        const double nrand_d = (double)(c.coordination_number_only ? 0 : c.n_random_samples);
        const double chunk = std::max(1.0, std::min(std::max(nrand_d, 1.0), 48.0e6 / (double)h->nv_atoms));
        const double maxq = std::max((double)h->nv_atoms, chunk * (double)h->nv_atoms);
        const double slot_b = h->path == 1 ? maxq * (16 + 16 + 16 + 24 + 1) + (double)c.solvent_nmols * (32 + 8 + 16 * chunk) + 27.0 * 16 * c.solute_natomspermol
                                           : 36.0 * (double)h->nv_atoms + 64.0 * (double)c.solvent_nmols * chunk;
        const double counters_b = 16.0 * (double)h->cnt_len;      // u64 block + the f64 image written by cmx_finish
        const double need = counters_b + slot_b * (double)h->ctx.size() * (double)h->batch + 8.0 * (double)h->in_floats * (double)h->ring.size();
        if (need > 0.95 * (double)free_b) {
            char buf[400];
            std::snprintf(buf, sizeof buf, "not enough device memory: counters %.1f GB (nbins x (4 + 2 n_groups_solute + 2 n_groups_solvent) x 16 B) + "
                          "frames in flight %.1f GB > %.1f GB free on device %d; use fewer groups (ResidueContributions-style custom groups instead of "
                          "per-atom rows), fewer batches in flight (n_streams) or frames per batch (batch_frames)",
                          counters_b / 1e9, (need - counters_b) / 1e9, (double)free_b / 1e9, c.device);
            return fail(h, CMX_ERR_MEMORY, buf);
        }
    }
    CK(h->d_cnt.ensure(h->cnt_len)); CK(cudaMemset(h->d_cnt.p, 0, sizeof(u64) * h->d_cnt.n));
    u64 *q = h->d_cnt.p;
    P.md = q; q += nb; P.md_r = q; q += nb; P.rdf = q; q += nb; P.rdf_r = q; q += nb;
    P.gsol = q; q += nb * c.n_groups_solute; P.gsol_r = q; q += nb * c.n_groups_solute;
    P.gsolv = q; q += nb * c.n_groups_solvent; P.gsolv_r = q;
    // scratch common to both paths
    size_t nvm = c.solvent_nmols, nrand = (size_t)P.nrand;
    // random phase in chunks of samples: at most ~48 M query atoms of scratch per frame context
    h->sample_chunk = (int)std::max<size_t>(1, std::min<size_t>(std::max<size_t>(nrand, 1), (size_t)(48.0e6 / (double)h->nv_atoms)));
    phase("group maps, counters");
    CK(h->d_stats.ensure(8, true));
    P.cnt_base = h->d_cnt.p; P.acc = nullptr; P.w = 1.0;
    // shared-memory histograms: md and rdf always; group rows while all rows fit in 64 KB next to them
    {
        const size_t row = sizeof(unsigned) * nb, budget = 64 * 1024;
        P.priv_sol = (c.autocorrelation || h->path == 2) && row * (2 + (size_t)c.n_groups_solute) <= budget;
        P.priv_solv = !c.autocorrelation && row * (2 + (size_t)c.n_groups_solvent + (P.priv_sol ? (size_t)c.n_groups_solute : 0)) <= budget;
        h->hist_smem = row * (size_t)hist_rows(P);
        if (h->hist_smem > 200 * 1024) return fail(h, CMX_ERR_ARG, "too many histogram bins for the shared-memory counters (cutoff / binstep > 25000)");
        CK(cudaFuncSetAttribute(k_finalise<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->hist_smem));
        CK(cudaFuncSetAttribute(k_finalise<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->hist_smem));
    }
    if (c.keep_lists) {
        if (h->path == 1) CK(h->d_list_all.ensure((size_t)c.solute_nmols * nvm));
        CK(h->d_rand_list.ensure(std::max<size_t>(nrand * nvm, 1)));
    }
    for (FrameCtx *x_ : h->ctx) {
        h->cur = x_;
        CK(x_->d_scalars.ensure(16, true));
        if (h->path == 1) {
            x_->slots.resize((size_t)h->batch);      // (their scratch is allocated at the first flush that uses them)
        } else {
            CK(x_->d_list.ensure(nvm));
            int rc = pairs_create(h); if (rc) return rc;
        }
    }
    h->cur = h->ctx[0];
    CK(cudaDeviceSynchronize());
    phase("frame contexts");
    return CMX_OK;
}

int32_t cmx_create(const cmx_config *cfg, cmx_handle **out) {
    if (!cfg || !out) { g_create_error = "cmx_create: null argument"; return CMX_ERR_ARG; }
    {
        const double t0 = now_ms();
        join_reapers();      // the memory of handles destroyed before must be back before this one allocates
        if (std::getenv("CMX_TRACE")) std::fprintf(stderr, "[cmx trace] create: %8.1f ms  waiting for the release of earlier handles\n", now_ms() - t0);
    }
    cmx_handle *h = new cmx_handle();
    if (cfg->struct_size != (int32_t)sizeof(cmx_config)) { g_create_error = "cmx_config.struct_size mismatch (ABI)"; delete h; *out = nullptr; return CMX_ERR_ARG; }
    int rc = cfg->n_devices > 1 ? group_create(h, cfg) : create_impl(h, cfg);
    if (rc) { g_create_error = h->err; cmx_destroy(h); *out = nullptr; return rc; }
    *out = h;
    return CMX_OK;
}

int32_t cmx_acquire_frame_buffer(cmx_handle *h, float **solute_xyz, float **solvent_xyz) {
    if (!h) return CMX_ERR_ARG;
    if (is_group(h)) {
        if (h->acquired_child) return fail(h, CMX_ERR_STATE, "cmx_acquire_frame_buffer: previous slot not submitted");
        cmx_handle *c = h->children[(size_t)h->next_child];
        int rc = cmx_acquire_frame_buffer(c, solute_xyz, solvent_xyz);
        if (rc) return group_fail(h, c, rc);
        h->acquired_child = c;
        return CMX_OK;
    }
    CK(cudaSetDevice(h->device));
    if (h->acquired >= 0) return fail(h, CMX_ERR_STATE, "cmx_acquire_frame_buffer: previous slot not submitted");
    SubmitTimer timer(h);
    Slot &s = h->ring[h->next_slot];
    if (!s.h_in) {      // first use of this slot: pinned memory on the GPU's NUMA node + its device twin
        NumaPrefer numa_guard(h->numa_node);
        CK(cudaHostAlloc(&s.h_in, sizeof(float) * h->in_floats, cudaHostAllocDefault));
        CK(cudaMalloc(&s.d_in, sizeof(float) * h->in_floats));
        CK(cudaEventCreateWithFlags(&s.h2d_done, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&s.consumed, cudaEventDisableTiming));
    }
    if (s.in_flight) {
        // the slot's previous frame may still wait in a batch that is not launched yet (ring shorter than the batches)
        { int rc = flush_if_pending(h, s.consumed); if (rc) return rc; }
        CK(wait_event(h, s.consumed)); s.in_flight = false;
    }
    h->acquired = h->next_slot;
    h->next_slot = (h->next_slot + 1) % (int)h->ring.size();
    if (h->cfg.autocorrelation) { if (solute_xyz) *solute_xyz = s.h_in; if (solvent_xyz) *solvent_xyz = s.h_in; }
    else { if (solute_xyz) *solute_xyz = s.h_in; if (solvent_xyz) *solvent_xyz = s.h_in + 3 * h->ns_atoms; }
    return CMX_OK;
}

int32_t cmx_submit_frame(cmx_handle *h, int64_t frame_index, double weight, const double cell[9]) {
    if (!h || !cell) return CMX_ERR_ARG;
    if (is_group(h)) {
        cmx_handle *c = h->acquired_child;
        if (!c) return fail(h, CMX_ERR_STATE, "cmx_submit_frame: no frame buffer acquired");
        h->acquired_child = nullptr;
        h->next_child = (h->next_child + 1) % (int)h->children.size();      // frame k of the run -> device k mod n
        if (!h->have_weight && weight > 0) {      // ONE reference weight for the integer counters of every device
            h->have_weight = true; h->w0 = weight;
            for (cmx_handle *k : h->children) { k->have_weight = true; k->w0 = weight; }
        }
        int rc = cmx_submit_frame(c, frame_index, weight, cell);
        return rc ? group_fail(h, c, rc) : CMX_OK;
    }
    CK(cudaSetDevice(h->device));
    if (h->acquired < 0) return fail(h, CMX_ERR_STATE, "cmx_submit_frame: no frame buffer acquired");
    SubmitTimer timer(h);
    Slot &s = h->ring[h->acquired];
    h->acquired = -1;
    CK(cudaMemcpyAsync(s.d_in, s.h_in, sizeof(float) * h->in_floats, cudaMemcpyHostToDevice, h->s_copy));
    CK(cudaEventRecord(s.h2d_done, h->s_copy));
    FrameCtx *next = h->ctx[(size_t)h->fill];     // the context whose batch this frame joins: its kernels come later on this stream
    CK(cudaStreamWaitEvent(next->stream, s.h2d_done, 0));
    h->stats.h2d_bytes += (int64_t)(sizeof(float) * h->in_floats);
    const float *dsol = s.d_in, *dsolv = h->cfg.autocorrelation ? s.d_in : s.d_in + 3 * h->ns_atoms;
    // the slot is reusable once the frame's kernels are done: `consumed` is recorded behind them
    int rc = submit_common(h, dsol, dsolv, frame_index, weight, cell, s.consumed);
    if (rc != CMX_OK) { (void)flush_all(h); cudaEventRecord(s.consumed, next->stream); }   // keep the ring consistent
    s.in_flight = true;
    return rc;
}

int32_t cmx_submit_frame_device(cmx_handle *h, const float *d_solute_xyz, const float *d_solvent_xyz, int64_t frame_index,
                                double weight, const double cell[9]) {
    if (!h || !cell || !d_solvent_xyz) return CMX_ERR_ARG;
    if (is_group(h)) return fail(h, CMX_ERR_STATE, "cmx_submit_frame_device: device-resident frames need a single-device handle (n_devices <= 1)");
    CK(cudaSetDevice(h->device));
    const float *dsol = h->cfg.autocorrelation ? d_solvent_xyz : d_solute_xyz;
    if (!dsol) return fail(h, CMX_ERR_ARG, "cmx_submit_frame_device: null solute pointer");
    SubmitTimer timer(h);
    return submit_common(h, dsol, d_solvent_xyz, frame_index, weight, cell, nullptr);
}

int32_t cmx_sync(cmx_handle *h) {
    if (!h) return CMX_ERR_ARG;
    if (is_group(h)) {
        for (cmx_handle *c : h->children) { int rc = cmx_sync(c); if (rc) return group_fail(h, c, rc); }
        return CMX_OK;
    }
    CK(cudaSetDevice(h->device));
    { int rc = flush_all(h); if (rc) return rc; }
    if (h->ev_first_set) for (FrameCtx *x : h->ctx) CK(cudaEventRecord(x->ev_end, x->stream));
    { int rc = sync_all(h); if (rc) return rc; }
    if (h->ev_first_set) {
        float best = 0;
        for (FrameCtx *x : h->ctx) { float ms = 0; if (cudaEventElapsedTime(&ms, h->ev_first, x->ev_end) == cudaSuccess) best = std::max(best, ms); }
        (void)cudaGetLastError();
        h->stats.gpu_ms_total += best; h->ev_first_set = false;
    }
    prof_collect(h);
    trace_report(h);
    for (auto &s : h->ring) s.in_flight = false;   // (everything was synchronised above)
    int sticky = 0;
    for (FrameCtx *x : h->ctx) { int v = 0; CK(cudaMemcpy(&v, x->d_scalars.p + 8, sizeof(int), cudaMemcpyDeviceToHost)); sticky |= v; }
    if (sticky) return fail(h, CMX_ERR_STATE, "deferred-pair buffer overflow: too many exactly tied / cutoff-edge pairs in one frame");
    return CMX_OK;
}

int32_t cmx_counters_device(cmx_handle *h, void **device_ptr, int64_t *n_uint64) {
    if (!h || !device_ptr || !n_uint64) return CMX_ERR_ARG;
    if (is_group(h)) {
        int rc = group_merge(h); if (rc) return rc;
        rc = cmx_counters_device(h->children[0], device_ptr, n_uint64);
        return rc ? group_fail(h, h->children[0], rc) : CMX_OK;
    }
    if (h->acc_used) return fail(h, CMX_ERR_STATE, "cmx_counters_device: frame weights varied; integer counters were folded to fp64 (use cmx_finish per GPU and sum)");
    *device_ptr = h->d_cnt.p; *n_uint64 = (int64_t)h->cnt_len;
    return CMX_OK;
}

// f64 image of the counters with the frame weights applied; head_only: just md / md_random / rdf / rdf_random (what
// a caller that leaves the group arrays on the device asks for -- C5 per-atom: 24 kB instead of 12 GB)
static int emit_counters(cmx_handle *h, bool head_only = false) {
    size_t n = head_only ? 4 * (size_t)h->nbins : h->cnt_len, nb = h->nbins;
    const double w = h->have_weight ? h->w0 : 1.0;
    const size_t gs = nb * h->cfg.n_groups_solute;
    size_t lo = 4 * nb, hi = 4 * nb + 2 * gs;
    if (!h->cfg.autocorrelation) lo = hi = 0;
    CK(h->d_emit.ensure(n));
    launch(h, k_emit, dim3((unsigned)((n + 255) / 256)), dim3(256), (const u64 *)h->d_cnt.p,
           (const double *)(h->acc_used ? h->d_acc.p : nullptr), h->d_emit.p, n, lo, hi, w);
    CK(cudaStreamSynchronize(h->cur->stream));
    return CMX_OK;
}

int32_t cmx_counters_device_f64(cmx_handle *h, double **device_ptr, int64_t *n_f64) {
    if (!h || !device_ptr || !n_f64) return CMX_ERR_ARG;
    if (is_group(h)) {
        int rc = group_merge(h); if (rc) return rc;
        rc = cmx_counters_device_f64(h->children[0], device_ptr, n_f64);
        return rc ? group_fail(h, h->children[0], rc) : CMX_OK;
    }
    int rc = cmx_sync(h); if (rc) return rc;
    rc = emit_counters(h); if (rc) return rc;
    h->emit_valid = true;
    *device_ptr = h->d_emit.p; *n_f64 = (int64_t)h->cnt_len;
    return CMX_OK;
}

int32_t cmx_finish(cmx_handle *h, cmx_counters *out) {
    if (!h || !out) return CMX_ERR_ARG;
    if (is_group(h)) {
        int rc = group_merge(h); if (rc) return rc;
        rc = cmx_finish(h->children[0], out);
        return rc ? group_fail(h, h->children[0], rc) : CMX_OK;
    }
    int rc = cmx_sync(h); if (rc) return rc;
    size_t nb = h->nbins;
    const size_t gs = nb * h->cfg.n_groups_solute, gv = nb * h->cfg.n_groups_solvent;
    const bool head_only = !out->solute_group_count && !out->solute_group_count_random && !out->solvent_group_count && !out->solvent_group_count_random;
    if (!h->emit_valid) { rc = emit_counters(h, head_only); if (rc) return rc; }
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
    if (is_group(h)) return fail(h, CMX_ERR_STATE, "cmx_read_minimum_distances needs a single-device handle");
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
    if (is_group(h)) return fail(h, CMX_ERR_STATE, "cmx_read_random_minimum_distances needs a single-device handle");
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
int32_t cmx_free_pinned(void *ptr) {
    if (ptr) reap([ptr] { cudaFreeHost(ptr); });      // unpinning hundreds of MB takes longer than a short run: off the caller's thread
    return CMX_OK;
}

int32_t cmx_get_stats(cmx_handle *h, cmx_stats *out) {
    if (!h || !out) return CMX_ERR_ARG;
    if (is_group(h)) { int rc = CMX_OK; *out = group_stats_sum(h, &rc); return rc; }
    int rc = cmx_sync(h); if (rc) return rc;
    u64 st[8];
    CK(cudaMemcpy(st, h->d_stats.p, sizeof st, cudaMemcpyDeviceToHost));
    h->stats.pair_evals = (int64_t)st[0]; h->stats.deferred = (int64_t)st[1];
    h->stats.volume_total = h->volume_total; h->stats.sum_weights = h->sum_weights;
    *out = h->stats;
    return CMX_OK;
}

int32_t cmx_reset(cmx_handle *h) {
    if (!h) return CMX_ERR_ARG;
    if (is_group(h)) {
        for (cmx_handle *c : h->children) { int rc = cmx_reset(c); if (rc) return group_fail(h, c, rc); }
        h->have_weight = false; h->w0 = 1.0; h->next_child = 0; h->stopped_by_file = false;
        return CMX_OK;
    }
    int rc = cmx_sync(h); if (rc) return rc;
    CK(cudaMemset(h->d_cnt.p, 0, sizeof(u64) * h->d_cnt.n));
    if (h->acc_used) CK(cudaMemset(h->d_acc.p, 0, sizeof(double) * h->d_acc.n));
    CK(cudaMemset(h->d_stats.p, 0, sizeof(u64) * 8));
    h->acc_used = false; h->P.acc = nullptr; h->have_weight = false; h->w0 = 1.0; h->emit_valid = false;
    h->volume_total = 0; h->sum_weights = 0;
    h->stats = cmx_stats{};
    return CMX_OK;
}

int32_t cmx_set_option(cmx_handle *h, const char *name, double value) {
    if (!h || !name) return CMX_ERR_ARG;
    if (is_group(h)) {
        for (cmx_handle *c : h->children) { int rc = cmx_set_option(c, name, value); if (rc) return group_fail(h, c, rc); }
        return CMX_OK;
    }
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
        h->active_ctx = v; h->fill = 0;
    }
    else if (n == "group_lanes") { /* accepted, ignored */ }
    else if (n == "poll_stop_file") h->poll_stop_file = value != 0;
    else if (n == "sync_destroy") h->sync_destroy = value != 0;
    else if (n == "xtc_host_decode") h->xtc_host_decode = value != 0;
    else return fail(h, CMX_ERR_ARG, "unknown option: " + n);
    return CMX_OK;
}

}  // extern "C"

#include "cmx_feed.inl"
#include "cmx_xtc.inl"
#include "cmx_final.inl"
