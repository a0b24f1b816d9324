// cmx_feed.inl -- native DCD frame feed and the device-side group reduction (included by cmx_b200.cu)
//
// Feed: the serial host work of the reference's frame loop -- FortranFiles record reads of the cell
// and the X, Y, Z blocks followed by a per-atom gather of the selected indices
// (src/trajectory_formats/NamdDCD.jl:141-169), under read_lock (src/mddf.jl:307) -- becomes:
// reader threads pread() whole raw frames into a pinned ring, ONE async H2D copy per frame, and a
// gather kernel that builds the xyz triplets of the selections on the device.
// Reduction: rows of a group-count array summed per group on the device (contributions /
// ResidueContributions count stage, src/tools/contributions.jl:70-248).

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>

struct cmx_dcd {
    int fd = -1;
    cmx_dcd_info info{};
    std::string path;
};

namespace {

std::string g_dcd_error;
std::mutex g_dcd_error_mu;

int dcd_fail(int code, const std::string &msg) {
    std::lock_guard<std::mutex> lk(g_dcd_error_mu);
    g_dcd_error = msg;
    return code;
}

bool pread_all(int fd, void *dst, size_t n, off_t off) {
    unsigned char *p = (unsigned char *)dst;
    while (n) {
        ssize_t r = pread(fd, p, n, off);
        if (r <= 0) return false;
        p += r; n -= (size_t)r; off += r;
    }
    return true;
}

// getunitcell(::NamdDCD), src/trajectory_formats/NamdDCD.jl:175-188: record = [A, gamma, B, beta, alpha, C];
// any zero angle -> all 90; lattice vectors as matrix columns, a along x, b in the xy plane (Chemfiles).
void dcd_cell(const double u[6], double cell[9]) {
    double A = u[0], gam = u[1], B = u[2], bet = u[3], alp = u[4], C = u[5];
    if (alp == 0.0 || bet == 0.0 || gam == 0.0) alp = bet = gam = 90.0;
    for (int k = 0; k < 9; ++k) cell[k] = 0.0;
    if (alp == 90.0 && bet == 90.0 && gam == 90.0) { cell[0] = A; cell[4] = B; cell[8] = C; return; }
    const double d2r = 3.141592653589793238462643383279502884 / 180.0;
    double ca = std::cos(alp * d2r), cb = std::cos(bet * d2r), cg = std::cos(gam * d2r), sg = std::sin(gam * d2r);
    cell[0] = A;
    cell[3] = B * cg; cell[4] = B * sg;
    double cx = C * cb, cy = C * (ca - cb * cg) / sg;
    cell[6] = cx; cell[7] = cy; cell[8] = std::sqrt(std::max(C * C - cx * cx - cy * cy, 0.0));
}

// raw frame block: [4][48 cell][4] then for X, Y, Z: [4][4*natoms][4]
inline size_t dcd_block_offset(int64_t natoms, int k) { return 56 + (size_t)k * (8 + 4 * (size_t)natoms) + 4; }

__global__ void k_gather_dcd(const unsigned char *__restrict__ raw, long long natoms_file, const int *__restrict__ idx,
                             int n, float *__restrict__ out) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const size_t stride = 8 + 4 * (size_t)natoms_file;
    const float *X = (const float *)(raw + 60), *Y = (const float *)(raw + 60 + stride), *Z = (const float *)(raw + 60 + 2 * stride);
    int a = idx[t];
    out[3 * (size_t)t] = __ldg(&X[a]); out[3 * (size_t)t + 1] = __ldg(&Y[a]); out[3 * (size_t)t + 2] = __ldg(&Z[a]);
}

// same gather for a raw frame that already is fp32 xyz triplets (decoded XTC)
__global__ void k_gather_xyz(const float *__restrict__ raw, const int *__restrict__ idx, int n, float *__restrict__ out) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const size_t a = (size_t)idx[t];
    out[3 * (size_t)t] = __ldg(&raw[3 * a]); out[3 * (size_t)t + 1] = __ldg(&raw[3 * a + 1]); out[3 * (size_t)t + 2] = __ldg(&raw[3 * a + 2]);
}

}  // namespace

// What the feed loop needs from a trajectory format: the size of one raw frame in the ring, its layout on the device
// (0 = DCD records, 1 = fp32 xyz triplets in Angstrom) and a thread-safe "produce raw frame `frame` in `dst`" that also
// yields the unit cell.
// Layouts: 0 = DCD records, 1 = fp32 xyz triplets, 2 = a compressed XTC frame (header + group records + bit stream)
// that a kernel decodes into a device buffer of `decoded_bytes` before the gather.  `fill` may restrict the H2D copy to
// byte ranges of the slot (a DCD read up to the last selected atom, a compressed frame shorter than the slot) and may
// pass one integer to the device stage (the number of XTC groups).
struct FeedFill {
    std::vector<std::pair<size_t, size_t>> ranges;   // (offset, bytes) of the slot to copy; empty = the whole slot
    int aux = 0;
};
struct FeedSource {
    const char *what = "";
    int64_t natoms = 0, nframes = 0;
    size_t slot_bytes = 0, decoded_bytes = 0;
    int layout = 0;
    std::function<bool(int64_t frame, unsigned char *dst, double cell[9], FeedFill &out, std::string &err)> fill;
};
void launch_xtc_decode(const unsigned char *d_raw, float *d_dec, int ngroups, cudaStream_t stream);   // cmx_xtc.inl

struct FeedSlot {
    unsigned char *h_raw = nullptr, *d_raw = nullptr;
    float *d_xyz = nullptr, *d_dec = nullptr;
    FeedFill fill;
    cudaEvent_t h2d_done = nullptr, gathered = nullptr, consumed = nullptr;
    int64_t filled = -1, h2d_issued = -1;
    double cell[9] = {0};
    bool used = false;
};

struct cmx_feed {
    std::vector<FeedSlot> slots;
    unsigned char *h_arena = nullptr, *d_arena = nullptr;      // ONE pinned and ONE device allocation behind the whole ring
    size_t frame_bytes = 0, decoded_bytes = 0;
    int *d_idx = nullptr; size_t n_idx = 0;
    std::vector<int32_t> idx_host;
    // group reduction scratch
    DevBuf<u64> red_cnt; DevBuf<double> red_acc, red_out; DevBuf<int> red_rows, red_items;
    bool red_has_acc = false;      // the last row reduction also produced an fp64 part (F.red_acc)
    DevBuf<double> fin_prof, fin_sc, fin_selr_acc; DevBuf<u64> fin_selr;   // cmx_final.inl: profile block, scalars, ideal-gas row sums
    void release_ring() {
        for (auto &s : slots) {
            if (s.h2d_done) cudaEventDestroy(s.h2d_done);
            if (s.gathered) cudaEventDestroy(s.gathered);
            if (s.consumed) cudaEventDestroy(s.consumed);
        }
        if (h_arena) cudaFreeHost(h_arena);
        if (d_arena) cudaFree(d_arena);
        h_arena = d_arena = nullptr;
    }
    void release() {
        release_ring();
        slots.clear();
        if (d_idx) cudaFree(d_idx);
        d_idx = nullptr; n_idx = 0; frame_bytes = 0; idx_host.clear();
        red_cnt.release(); red_acc.release(); red_out.release(); red_rows.release(); red_items.release();
        fin_prof.release(); fin_sc.release(); fin_selr_acc.release(); fin_selr.release();
    }
};

namespace {

void feed_destroy(cmx_handle *h) {
    if (h->feed) { h->feed->release(); delete h->feed; h->feed = nullptr; }
}

int feed_prepare(cmx_handle *h, int64_t natoms_file, size_t slot_bytes, size_t decoded_bytes, const int32_t *sol_idx, const int32_t *solv_idx, int nslots) {
    if (!h->feed) h->feed = new cmx_feed();
    cmx_feed &F = *h->feed;
    const size_t fb = slot_bytes;
    if (F.frame_bytes != fb || F.decoded_bytes != decoded_bytes || (int)F.slots.size() != nslots) {
        { int rc = sync_all(h); if (rc) return rc; }
        F.release_ring();
        F.slots.assign((size_t)nslots, FeedSlot());
        NumaPrefer numa_guard(h->numa_node);
        auto pad = [](size_t b) { return (b + 255) & ~(size_t)255; };
        const size_t hb = pad(fb), xb = pad(sizeof(float) * h->in_floats), db = pad(decoded_bytes);
        CK(cudaHostAlloc(&F.h_arena, hb * (size_t)nslots, cudaHostAllocDefault));
        CK(cudaMalloc(&F.d_arena, (hb + xb + db) * (size_t)nslots));
        for (size_t k = 0; k < F.slots.size(); ++k) {
            FeedSlot &s = F.slots[k];
            s.h_raw = F.h_arena + k * hb;
            unsigned char *d = F.d_arena + k * (hb + xb + db);
            s.d_raw = d; s.d_xyz = (float *)(d + hb); s.d_dec = decoded_bytes ? (float *)(d + hb + xb) : nullptr;
            CK(cudaEventCreateWithFlags(&s.h2d_done, cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&s.gathered, cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&s.consumed, cudaEventDisableTiming));
        }
        F.frame_bytes = fb; F.decoded_bytes = decoded_bytes;
    }
    CK(cudaStreamSynchronize(h->s_copy));   // copies of an earlier run still reading the pinned slots
    for (auto &s : F.slots) { s.filled = -1; s.h2d_issued = -1; }
    // selection indices (1-based file positions -> 0-based), solute first then solvent
    const size_t ns = h->cfg.autocorrelation ? 0 : h->ns_atoms, nv = h->nv_atoms;
    std::vector<int32_t> idx(ns + nv);
    for (size_t k = 0; k < ns; ++k) idx[k] = sol_idx[k] - 1;
    for (size_t k = 0; k < nv; ++k) idx[ns + k] = solv_idx[k] - 1;
    for (int32_t v : idx)
        if (v < 0 || (int64_t)v >= natoms_file) return fail(h, CMX_ERR_ARG, "selection index outside the atoms of the file");
    if (idx != F.idx_host) {
        { int rc = sync_all(h); if (rc) return rc; }
        if (F.d_idx) { CK(cudaFree(F.d_idx)); F.d_idx = nullptr; }
        CK(cudaMalloc(&F.d_idx, sizeof(int) * idx.size()));
        CK(cudaMemcpy(F.d_idx, idx.data(), sizeof(int) * idx.size(), cudaMemcpyHostToDevice));
        F.n_idx = idx.size();
        F.idx_host.swap(idx);
    }
    return CMX_OK;
}

// ---- group reduction kernels ------------------------------------------------------------------------
// one block per (item, bin chunk); an item is a run of <= CMX_RED_ROWS rows of one group: {group, q_begin, q_end}
#define CMX_RED_ROWS 256
template <bool HAS_CNT>      // false: the f64 block alone is authoritative (after cmx_counters_device_f64 + all-reduce)
__global__ void __launch_bounds__(256)
k_reduce_rows(const u64 *__restrict__ cnt, const double *__restrict__ acc, int nbins, const int *__restrict__ items,
              const int *__restrict__ rows, u64 *__restrict__ out_cnt, double *__restrict__ out_acc) {
    const int b = blockIdx.y * blockDim.x + threadIdx.x;
    if (b >= nbins) return;
    const int g = items[3 * blockIdx.x], q0 = items[3 * blockIdx.x + 1], q1 = items[3 * blockIdx.x + 2];
    u64 s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    double a = 0.0;
    int q = q0;
    for (; q + 4 <= q1; q += 4) {   // four independent 8-byte streams per thread, each row segment coalesced over the bins
        const size_t r0 = (size_t)__ldg(&rows[q]), r1 = (size_t)__ldg(&rows[q + 1]), r2 = (size_t)__ldg(&rows[q + 2]), r3 = (size_t)__ldg(&rows[q + 3]);
        if (HAS_CNT) {
            s0 += __ldcs(&cnt[r0 * nbins + b]); s1 += __ldcs(&cnt[r1 * nbins + b]);
            s2 += __ldcs(&cnt[r2 * nbins + b]); s3 += __ldcs(&cnt[r3 * nbins + b]);
        }
        if (acc) a += (__ldcs(&acc[r0 * nbins + b]) + __ldcs(&acc[r1 * nbins + b])) + (__ldcs(&acc[r2 * nbins + b]) + __ldcs(&acc[r3 * nbins + b]));
    }
    for (; q < q1; ++q) {
        const size_t r = (size_t)__ldg(&rows[q]);
        if (HAS_CNT) s0 += __ldcs(&cnt[r * nbins + b]);
        if (acc) a += __ldcs(&acc[r * nbins + b]);
    }
    const u64 s = (s0 + s1) + (s2 + s3);
    if (s) atomicAdd(&out_cnt[(size_t)g * nbins + b], s);
    if (acc && a != 0.0) atomicAdd(&out_acc[(size_t)g * nbins + b], a);
}

__global__ void k_reduce_emit(const u64 *__restrict__ cnt, const double *__restrict__ acc, size_t n, double scale, double *__restrict__ out) {
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) out[k] = (acc ? acc[k] : 0.0) + scale * (double)cnt[k];
}

// Row sums of one group-count array into F.red_cnt (+ F.red_acc when frame weights varied): [n_groups][nbins], on the
// handle's current stream (not synchronised).  Shared by cmx_reduce_groups and cmx_contributions.
int reduce_rows_device(cmx_handle *h, int32_t which, int32_t n_groups, const int32_t *offsets, const int32_t *rows) {
    if (which < 0 || which > 3 || n_groups < 1 || !offsets) return fail(h, CMX_ERR_ARG, "cmx_reduce_groups: invalid argument");
    const int nrows_arr = which < 2 ? h->cfg.n_groups_solute : h->cfg.n_groups_solvent;
    const size_t nb = h->nbins, gs = nb * h->cfg.n_groups_solute, gv = nb * h->cfg.n_groups_solvent;
    const size_t base = 4 * nb + (which == 0 ? 0 : which == 1 ? gs : which == 2 ? 2 * gs : 2 * gs + gv);
    const size_t nids = (size_t)offsets[n_groups];
    if (offsets[0] != 0) return fail(h, CMX_ERR_ARG, "cmx_reduce_groups: offsets[0] must be 0");
    if (nids && !rows) return fail(h, CMX_ERR_ARG, "cmx_reduce_groups: null rows");
    std::vector<int> items;
    for (int g = 0; g < n_groups; ++g) {
        if (offsets[g + 1] < offsets[g]) return fail(h, CMX_ERR_ARG, "cmx_reduce_groups: offsets must be non-decreasing");
        for (int q = offsets[g]; q < offsets[g + 1]; q += CMX_RED_ROWS) {
            items.push_back(g); items.push_back(q); items.push_back(std::min(q + CMX_RED_ROWS, offsets[g + 1]));
        }
    }
    for (size_t k = 0; k < nids; ++k)
        if (rows[k] < 0 || rows[k] >= nrows_arr) return fail(h, CMX_ERR_ARG, "cmx_reduce_groups: row id out of range");
    { int rc = cmx_sync(h); if (rc) return rc; }
    if (!h->feed) h->feed = new cmx_feed();
    cmx_feed &F = *h->feed;
    const size_t nout = (size_t)n_groups * nb;
    // emit_valid: the f64 block of cmx_counters_device_f64 holds the (all-reduced) counters with the weights applied
    const bool f64only = h->emit_valid;
    F.red_has_acc = h->acc_used || f64only;
    CK(F.red_cnt.ensure(nout)); CK(F.red_out.ensure(nout));
    if (F.red_has_acc) CK(F.red_acc.ensure(nout));
    CK(F.red_rows.ensure(std::max<size_t>(nids, 1))); CK(F.red_items.ensure(std::max<size_t>(items.size(), 3)));
    cudaStream_t st = h->cur->stream;
    CK(cudaMemsetAsync(F.red_cnt.p, 0, sizeof(u64) * nout, st));
    if (F.red_has_acc) CK(cudaMemsetAsync(F.red_acc.p, 0, sizeof(double) * nout, st));
    if (nids) CK(cudaMemcpyAsync(F.red_rows.p, rows, sizeof(int) * nids, cudaMemcpyHostToDevice, st));
    if (!items.empty()) {
        CK(cudaMemcpyAsync(F.red_items.p, items.data(), sizeof(int) * items.size(), cudaMemcpyHostToDevice, st));
        dim3 grid((unsigned)(items.size() / 3), (unsigned)((nb + 255) / 256));
        cudaEvent_t pe = prof_begin(h, 2);
        if (f64only)
            launch(h, k_reduce_rows<false>, grid, dim3(256), (const u64 *)nullptr, (const double *)(h->d_emit.p + base),
                   (int)nb, (const int *)F.red_items.p, (const int *)F.red_rows.p, F.red_cnt.p, F.red_acc.p);
        else
            launch(h, k_reduce_rows<true>, grid, dim3(256), (const u64 *)(h->d_cnt.p + base), (const double *)(h->acc_used ? h->d_acc.p + base : nullptr),
                   (int)nb, (const int *)F.red_items.p, (const int *)F.red_rows.p, F.red_cnt.p, h->acc_used ? F.red_acc.p : (double *)nullptr);
        prof_end(h, pe);
        if (h->profile && h->prof_used) {   // time of the reduction kernel alone (option "profile")
            CK(cudaStreamSynchronize(st));
            float ms = 0;
            if (cudaEventElapsedTime(&ms, h->prof_events[h->prof_used - 1].first, h->prof_events[h->prof_used - 1].second) == cudaSuccess)
                h->stats.gpu_ms_reduce = ms;
            h->prof_used--;
        }
    }
    return CMX_OK;
}

}  // namespace

extern "C" {

const char *cmx_dcd_last_error(void) {
    std::lock_guard<std::mutex> lk(g_dcd_error_mu);
    static thread_local std::string copy;
    copy = g_dcd_error;
    return copy.c_str();
}

int32_t cmx_dcd_open(const char *path, cmx_dcd **out, cmx_dcd_info *info) {
    if (!path || !out) return dcd_fail(CMX_ERR_ARG, "cmx_dcd_open: null argument");
    *out = nullptr;
    int fd = open(path, O_RDONLY);
    if (fd < 0) return dcd_fail(CMX_ERR_IO, std::string("cannot open ") + path);
    struct stat st;
    if (fstat(fd, &st) != 0) { close(fd); return dcd_fail(CMX_ERR_IO, std::string("cannot stat ") + path); }
    // three Fortran records: header ("CORD" + 20 int32), title, natoms (firstframe!, NamdDCD.jl:192-201)
    off_t pos = 0;
    int32_t len = 0, natoms = 0;
    char magic[4];
    bool ok = pread_all(fd, &len, 4, pos) && len >= 4 && pread_all(fd, magic, 4, pos + 4);
    if (!ok || std::memcmp(magic, "CORD", 4) != 0) { close(fd); return dcd_fail(CMX_ERR_IO, "not a little-endian DCD file"); }
    pos += 8 + len;
    ok = pread_all(fd, &len, 4, pos) && len >= 0;
    if (!ok) { close(fd); return dcd_fail(CMX_ERR_IO, "truncated DCD header (title record)"); }
    pos += 8 + len;
    ok = pread_all(fd, &len, 4, pos) && len == 4 && pread_all(fd, &natoms, 4, pos + 4) && natoms > 0;
    if (!ok) { close(fd); return dcd_fail(CMX_ERR_IO, "truncated DCD header (natoms record)"); }
    pos += 8 + len;
    // the first record of a frame must be the 48-byte unit cell (NamdDCD.jl:69-77)
    ok = pread_all(fd, &len, 4, pos);
    if (!ok || len != 48) { close(fd); return dcd_fail(CMX_ERR_IO, "DCD file does not contain unit cell information."); }
    cmx_dcd *d = new cmx_dcd();
    d->fd = fd; d->path = path;
    d->info.natoms = natoms;
    d->info.first_frame_offset = (int64_t)pos;
    d->info.frame_bytes = 56 + 3 * (8 + 4 * (int64_t)natoms);
    d->info.nframes = ((int64_t)st.st_size - (int64_t)pos) / d->info.frame_bytes;   // NamdDCD.jl:211-230: count, do not trust the header
    if (info) *info = d->info;
    *out = d;
    return CMX_OK;
}

int32_t cmx_dcd_close(cmx_dcd *d) {
    if (!d) return CMX_OK;
    if (d->fd >= 0) close(d->fd);
    delete d;
    return CMX_OK;
}

int32_t cmx_dcd_read_frame(cmx_dcd *d, int64_t iframe, float *x, float *y, float *z, double cell[9]) {
    if (!d) return dcd_fail(CMX_ERR_ARG, "cmx_dcd_read_frame: null handle");
    if (iframe < 0 || iframe >= d->info.nframes) return dcd_fail(CMX_ERR_ARG, "cmx_dcd_read_frame: frame out of range");
    const off_t base = (off_t)(d->info.first_frame_offset + iframe * d->info.frame_bytes);
    const int64_t n = d->info.natoms;
    if (cell) {
        double u[6];
        if (!pread_all(d->fd, u, 48, base + 4)) return dcd_fail(CMX_ERR_IO, "short read (unit cell record)");
        dcd_cell(u, cell);
    }
    float *dst[3] = {x, y, z};
    for (int k = 0; k < 3; ++k)
        if (dst[k] && !pread_all(d->fd, dst[k], 4 * (size_t)n, base + (off_t)dcd_block_offset(n, k)))
            return dcd_fail(CMX_ERR_IO, "short read (coordinate record)");
    return CMX_OK;
}

static int run_feed(cmx_handle *h, const FeedSource &src, const int32_t *solute_indices, const int32_t *solvent_indices,
                    const int64_t *frames, const double *weights, int64_t nframes, int32_t n_reader_threads) {
    if (!solvent_indices || (!frames && nframes > 0) || nframes < 0) return fail(h, CMX_ERR_ARG, std::string(src.what) + ": null argument");
    if (!h->cfg.autocorrelation && !solute_indices) return fail(h, CMX_ERR_ARG, std::string(src.what) + ": null solute indices");
    if (h->acquired >= 0) return fail(h, CMX_ERR_STATE, std::string(src.what) + ": a staging slot is acquired and not submitted");
    for (int64_t k = 0; k < nframes; ++k) {
        if (frames[k] < 0 || frames[k] >= src.nframes) return fail(h, CMX_ERR_ARG, std::string(src.what) + ": frame number outside the file");
        if (weights && !(weights[k] > 0)) return fail(h, CMX_ERR_ARG, std::string(src.what) + ": frame weights must be positive (skip zero-weight frames)");
    }
    if (nframes == 0) return CMX_OK;
    if (is_group(h)) {
        // one reader/consumer team per device: child c takes the frames k = c, c + n, c + 2n, ... of the list
        const size_t n = h->children.size();
        std::vector<std::vector<int64_t>> fr(n);
        std::vector<std::vector<double>> wt(n);
        for (int64_t k = 0; k < nframes; ++k) { fr[(size_t)k % n].push_back(frames[k]); if (weights) wt[(size_t)k % n].push_back(weights[k]); }
        if (!h->have_weight) {      // ONE reference weight for the integer counters of every device
            h->have_weight = true; h->w0 = weights ? weights[0] : 1.0;
            for (cmx_handle *c : h->children) if (!c->have_weight) { c->have_weight = true; c->w0 = h->w0; }
        }
        std::vector<int> rcs(n, CMX_OK);
        std::vector<std::thread> team;
        for (size_t c = 0; c < n; ++c)
            team.emplace_back([&, c] {
                if (fr[c].empty()) return;
                rcs[c] = run_feed(h->children[c], src, solute_indices, solvent_indices, fr[c].data(), weights ? wt[c].data() : nullptr,
                                  (int64_t)fr[c].size(), n_reader_threads);
            });
        for (auto &t : team) t.join();
        for (size_t c = 0; c < n; ++c) {
            h->stopped_by_file |= h->children[c]->stopped_by_file;
            if (rcs[c]) return group_fail(h, h->children[c], rcs[c]);
        }
        return CMX_OK;
    }
    CK(cudaSetDevice(h->device));
    SubmitTimer timer(h);
    const int T = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(n_reader_threads > 0 ? n_reader_threads : 2, 16), nframes));
    // ring slots: every compute stream needs a frame in flight and one being staged behind it (a slot is busy from the
    // read until its frame's kernels have finished), bounded to ~4 GB of pinned memory
    const int nctx = h->active_ctx > 0 ? h->active_ctx : (int)h->ctx.size();
    const int64_t cap = std::max<int64_t>(3, (int64_t)(4.0e9 / (double)src.slot_bytes));
    // (a frame also occupies its slot while it waits in a batch that is not launched yet: B frames per context)
    const int S = (int)std::min<int64_t>(std::max((nctx + 1) * h->batch + 2, T + 2), cap);
    { int rc = feed_prepare(h, src.natoms, src.slot_bytes, src.decoded_bytes, solute_indices, solvent_indices, S); if (rc) return rc; }
    cmx_feed &F = *h->feed;
    std::mutex mu;
    std::condition_variable cv;
    std::atomic<bool> abort_flag{false};
    std::string io_error;
    const int device = h->device;
    const size_t fb = src.slot_bytes;
    auto reader = [&](int t) {
        cudaSetDevice(device);
        for (int64_t k = t; k < nframes; k += T) {
            FeedSlot &s = F.slots[(size_t)(k % S)];
            {   // the slot's previous occupant must have had its H2D issued ...
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return abort_flag.load() || k < S || s.h2d_issued == k - S; });
                if (abort_flag.load()) return;
            }
            if (k >= S && cudaEventSynchronize(s.h2d_done) != cudaSuccess) {   // ... and finished, before the pinned buffer is overwritten
                std::lock_guard<std::mutex> lk(mu);
                io_error = "cudaEventSynchronize failed in the reader thread"; abort_flag = true; cv.notify_all();
                return;
            }
            std::string err;
            s.fill.ranges.clear(); s.fill.aux = 0;
            bool ok = src.fill(frames[k], s.h_raw, s.cell, s.fill, err);
            std::lock_guard<std::mutex> lk(mu);
            if (!ok) { if (!abort_flag.load()) io_error = err; abort_flag = true; }
            else s.filled = k;
            cv.notify_all();
            if (!ok) return;
        }
    };
    std::vector<std::thread> threads;
    for (int t = 0; t < T; ++t) threads.emplace_back(reader, t);
    int rc = CMX_OK;
    const size_t ns = h->cfg.autocorrelation ? 0 : h->ns_atoms;
    for (int64_t k = 0; k < nframes && rc == CMX_OK; ++k) {
        // cooperative interrupt of the reference (src/mddf.jl:301-304): a file named stop_complexmixtures in the working
        // directory ends the frame loop; the frames enqueued so far are finished and reported (cmx_stats.frames)
        if ((k & 7) == 0 && h->poll_stop_file && access("stop_complexmixtures", F_OK) == 0) { h->stopped_by_file = true; break; }
        FeedSlot &s = F.slots[(size_t)(k % S)];
        {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return abort_flag.load() || s.filled == k; });
            if (abort_flag.load()) { rc = fail(h, CMX_ERR_IO, io_error); break; }
        }
        double cell[9];
        std::memcpy(cell, s.cell, sizeof cell);
        FrameCtx *next = h->ctx[(size_t)h->fill];   // the context whose batch this frame joins
        cudaError_t e = cudaSuccess;
        auto step = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
        if (s.used) step(cudaStreamWaitEvent(h->s_copy, s.gathered, 0));            // d_raw free again
        size_t copied = 0;
        if (s.fill.ranges.empty()) { step(cudaMemcpyAsync(s.d_raw, s.h_raw, fb, cudaMemcpyHostToDevice, h->s_copy)); copied = fb; }
        else for (const auto &r : s.fill.ranges) {
            step(cudaMemcpyAsync(s.d_raw + r.first, s.h_raw + r.first, r.second, cudaMemcpyHostToDevice, h->s_copy));
            copied += r.second;
        }
        step(cudaEventRecord(s.h2d_done, h->s_copy));
        {
            std::lock_guard<std::mutex> lk(mu);
            s.h2d_issued = k;
            cv.notify_all();
        }
        step(cudaStreamWaitEvent(next->stream, s.h2d_done, 0));
        if (s.used) {                                                               // d_xyz free again
            if (flush_if_pending(h, s.consumed) != CMX_OK) { rc = CMX_ERR_CUDA; break; }
            step(cudaStreamWaitEvent(next->stream, s.consumed, 0));
        }
        if (e == cudaSuccess) {
            const unsigned nblk = (unsigned)((F.n_idx + 255) / 256);
            if (src.layout == 0) k_gather_dcd<<<nblk, 256, 0, next->stream>>>(s.d_raw, (long long)src.natoms, F.d_idx, (int)F.n_idx, s.d_xyz);
            else if (src.layout == 1) k_gather_xyz<<<nblk, 256, 0, next->stream>>>((const float *)s.d_raw, F.d_idx, (int)F.n_idx, s.d_xyz);
            else {      // compressed XTC frame: decode on the device (one thread per group of the bit stream), then gather
                launch_xtc_decode(s.d_raw, s.d_dec, s.fill.aux, next->stream);
                k_gather_xyz<<<nblk, 256, 0, next->stream>>>((const float *)s.d_dec, F.d_idx, (int)F.n_idx, s.d_xyz);
                h->stats.kernel_launches++;
            }
            h->stats.kernel_launches++;
            step(cudaEventRecord(s.gathered, next->stream));
        }
        if (e != cudaSuccess) { h->err = std::string(src.what) + ": " + cudaGetErrorString(e); rc = CMX_ERR_CUDA; break; }
        h->stats.h2d_bytes += (int64_t)copied;
        const float *dsol = s.d_xyz, *dsolv = s.d_xyz + 3 * ns;
        // `consumed` is recorded behind the frame's kernels, i.e. when its batch is launched
        rc = submit_common(h, dsol, dsolv, frames[k] + 1, weights ? weights[k] : 1.0, cell, s.consumed);
        if (rc != CMX_OK) { (void)flush_all(h); cudaEventRecord(s.consumed, next->stream); }
        s.used = true;
    }
    if (rc != CMX_OK || h->stopped_by_file) {
        std::lock_guard<std::mutex> lk(mu);
        abort_flag = true;
        cv.notify_all();
    }
    for (auto &t : threads) t.join();
    return rc;
}

int32_t cmx_run_dcd(cmx_handle *h, cmx_dcd *d, const int32_t *solute_indices, const int32_t *solvent_indices,
                    const int64_t *frames, const double *weights, int64_t nframes, int32_t n_reader_threads) {
    if (!h) return CMX_ERR_ARG;
    if (!d) return fail(h, CMX_ERR_ARG, "cmx_run_dcd: null argument");
    FeedSource src;
    src.what = "cmx_run_dcd"; src.natoms = d->info.natoms; src.nframes = d->info.nframes;
    src.slot_bytes = (size_t)d->info.frame_bytes; src.layout = 0;
    const int fd = d->fd;
    const int64_t first = d->info.first_frame_offset, fb = d->info.frame_bytes, natoms = d->info.natoms;
    // As the reference does with `lastatom` (src/trajectory_formats/NamdDCD.jl:86-118,151-166), each of the X, Y, Z
    // records is read -- and sent to the device -- only up to the last selected atom: solute + cosolvent selections
    // usually sit at the start of the file (C1: 4 012 of 62 026 atoms, 15 x fewer bytes per frame).
    int64_t last = 0;
    const cmx_handle *hh = is_group(h) ? h->children[0] : h;      // (a group handle keeps the problem sizes in its children)
    if (solvent_indices) for (size_t k = 0; k < hh->nv_atoms; ++k) last = std::max<int64_t>(last, solvent_indices[k]);
    if (solute_indices && !hh->cfg.autocorrelation) for (size_t k = 0; k < hh->ns_atoms; ++k) last = std::max<int64_t>(last, solute_indices[k]);
    last = std::min<int64_t>(std::max<int64_t>(last, 1), natoms);
    const bool partial = last * 4 < natoms * 3;       // worth three reads instead of one
    src.fill = [fd, first, fb, natoms, last, partial](int64_t frame, unsigned char *dst, double cell[9], FeedFill &out, std::string &err) {
        const off_t base = (off_t)(first + frame * fb);
        if (!partial) {
            if (!pread_all(fd, dst, (size_t)fb, base)) { err = "short read in DCD frame " + std::to_string((long long)frame); return false; }
        } else {
            if (!pread_all(fd, dst, 56, base)) { err = "short read in DCD frame " + std::to_string((long long)frame); return false; }
            for (int k = 0; k < 3; ++k) {
                const size_t off = 56 + (size_t)k * (8 + 4 * (size_t)natoms);      // [4-byte marker][4*natoms][4-byte marker]
                if (!pread_all(fd, dst + off, 4 + 4 * (size_t)last, base + (off_t)off)) { err = "short read in DCD frame " + std::to_string((long long)frame); return false; }
                out.ranges.push_back({off, 4 + 4 * (size_t)last});
            }
        }
        double u[6];
        std::memcpy(u, dst + 4, sizeof u);
        dcd_cell(u, cell);
        return true;
    };
    return run_feed(h, src, solute_indices, solvent_indices, frames, weights, nframes, n_reader_threads);
}

int32_t cmx_reduce_groups(cmx_handle *h, int32_t which, int32_t n_groups, const int32_t *offsets, const int32_t *rows, double *out) {
    if (!h) return CMX_ERR_ARG;
    if (is_group(h)) {
        int rc = group_merge(h); if (rc) return rc;
        rc = cmx_reduce_groups(h->children[0], which, n_groups, offsets, rows, out);
        return rc ? group_fail(h, h->children[0], rc) : CMX_OK;
    }
    if (!out) return fail(h, CMX_ERR_ARG, "cmx_reduce_groups: invalid argument");
    { int rc = reduce_rows_device(h, which, n_groups, offsets, rows); if (rc) return rc; }
    cmx_feed &F = *h->feed;
    const size_t nout = (size_t)n_groups * h->nbins;
    cudaStream_t st = h->cur->stream;
    // frame weight as in cmx_finish; the group counts of an autocorrelation carry w/2 (src/update_counters.jl:52-53)
    const double w = h->have_weight ? h->w0 : 1.0;
    const double scale = (h->cfg.autocorrelation && which < 2) ? w / 2 : w;
    launch(h, k_reduce_emit, dim3((unsigned)((nout + 255) / 256)), dim3(256), (const u64 *)F.red_cnt.p,
           (const double *)(F.red_has_acc ? F.red_acc.p : nullptr), nout, scale, F.red_out.p);
    CK(cudaStreamSynchronize(st));
    CK(cudaMemcpy(out, F.red_out.p, sizeof(double) * nout, cudaMemcpyDeviceToHost));
    return CMX_OK;
}

}  // extern "C"
