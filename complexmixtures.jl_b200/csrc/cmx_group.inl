// cmx_group.inl -- several GPUs behind ONE handle (included by cmx_b200.cu)
//
// The reference's single mddf() call uses the whole machine: parallel_setup (src/parallel_setup.jl:7-57) makes one
// chunk per thread, every chunk owns a private Result, frames are pulled under read_lock (src/mddf.jl:285-338) and the
// chunks are summed at the end (sum!, src/results.jl:629-649).  Here a "group" handle owns one complete device context
// (a child handle: streams, batches, staging ring, counters) per entry of cmx_config.device_ids; frames are dealt to the
// children in submission order, the native feeds run one reader/consumer team per child, and the children's
// accumulators are MOVED onto the first child over peer access (or a staged peer copy when the devices cannot address
// each other) whenever the caller looks at the counters.  Integer hits are summed exactly, so the result does not depend
// on the number of devices; the Philox stream is keyed by the frame index, not by the device.

namespace {

__global__ void k_merge_u64(u64 *__restrict__ dst, u64 *__restrict__ src, size_t n) {
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
        const u64 v = src[k];
        if (v) { dst[k] += v; src[k] = 0ull; }      // moved, not copied: a later merge must not count it again
    }
}
__global__ void k_merge_f64(double *__restrict__ dst, double *__restrict__ src, size_t n) {
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
        const double v = src[k];
        if (v != 0.0) { dst[k] += v; src[k] = 0.0; }
    }
}

bool is_group(const cmx_handle *h) { return !h->children.empty(); }

int group_fail(cmx_handle *g, cmx_handle *child, int rc) {
    g->err = child->err;
    return rc;
}

// dst (on child 0's device) += src (on child k's device); src is zeroed
template <class T, class K>
int merge_block(cmx_handle *g, cmx_handle *c0, cmx_handle *ck, T *dst, T *src, size_t n, K kernel) {
    cmx_handle *h = g;
    CK(cudaSetDevice(c0->device));
    const unsigned blocks = (unsigned)std::min<size_t>((n + 255) / 256, (size_t)c0->num_sms * 16);
    bool direct = ck->device == c0->device;
    if (!direct) {
        int can = 0;
        CK(cudaDeviceCanAccessPeer(&can, c0->device, ck->device));
        if (can) {
            cudaError_t e = cudaDeviceEnablePeerAccess(ck->device, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) { (void)cudaGetLastError(); e = cudaSuccess; }
            direct = e == cudaSuccess;
            if (!direct) (void)cudaGetLastError();
        }
    }
    if (direct) {
        // one kernel on the first device reads the peer's block through NVLink, adds and clears it
        kernel<<<blocks, 256, 0, c0->ctx[0]->stream>>>(dst, src, n);
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(c0->ctx[0]->stream));
    } else {
        // no peer addressing: stage through a scratch block on the first device
        DevBuf<T> tmp;
        CK(tmp.ensure(n));
        CK(cudaMemcpyPeer(tmp.p, c0->device, src, ck->device, sizeof(T) * n));
        kernel<<<blocks, 256, 0, c0->ctx[0]->stream>>>(dst, tmp.p, n);
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(c0->ctx[0]->stream));
        tmp.release();
        CK(cudaSetDevice(ck->device));
        CK(cudaMemset(src, 0, sizeof(T) * n));
    }
    return CMX_OK;
}

// sum!(R, r_chunk) for every child: everything the children accumulated moves onto the first child
int group_merge(cmx_handle *g) {
    cmx_handle *c0 = g->children[0];
    for (cmx_handle *c : g->children) { int rc = cmx_sync(c); if (rc) return group_fail(g, c, rc); }
    bool any_acc = false;
    for (cmx_handle *c : g->children) any_acc |= c->acc_used;
    if (any_acc && !c0->acc_used) { CK_G(cudaSetDevice(c0->device)); int rc = enter_acc_mode(c0); if (rc) return group_fail(g, c0, rc); }
    for (size_t k = 1; k < g->children.size(); ++k) {
        cmx_handle *c = g->children[k];
        int rc = merge_block(g, c0, c, c0->d_cnt.p, c->d_cnt.p, c0->cnt_len, k_merge_u64);
        if (rc) return rc;
        if (c->acc_used) { rc = merge_block(g, c0, c, c0->d_acc.p, c->d_acc.p, c0->cnt_len, k_merge_f64); if (rc) return rc; }
        c0->volume_total += c->volume_total; c0->sum_weights += c->sum_weights;
        c->volume_total = 0; c->sum_weights = 0;
        if (c->have_weight && !c0->have_weight) { c0->have_weight = true; c0->w0 = c->w0; }
        c0->emit_valid = false;
    }
    CK_G(cudaSetDevice(c0->device));
    return CMX_OK;
}

int group_create(cmx_handle *g, const cmx_config *cfg) {
    cmx_handle *h = g;
    if (cfg->n_devices > 64) return fail(h, CMX_ERR_ARG, "n_devices must be <= 64");
    if (!cfg->device_ids) return fail(h, CMX_ERR_ARG, "n_devices > 1 needs device_ids");
    if (cfg->keep_lists) return fail(h, CMX_ERR_ARG, "keep_lists (the parity hooks) needs a single device");
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    g->cfg = *cfg;
    for (int k = 0; k < cfg->n_devices; ++k) {
        if (cfg->device_ids[k] < 0 || cfg->device_ids[k] >= ndev) return fail(h, CMX_ERR_ARG, "device_ids: no such CUDA device");
        cmx_config cc = *cfg;
        cc.n_devices = 0; cc.device_ids = nullptr; cc.device = cfg->device_ids[k];
        cmx_handle *c = new cmx_handle();
        g->children.push_back(c);
        int rc = create_impl(c, &cc);
        if (rc) { g->err = c->err; return rc; }
    }
    g->device = g->children[0]->device;
    g->nbins = g->children[0]->nbins;
    return CMX_OK;
}

cmx_stats group_stats_sum(cmx_handle *g, int *rc_out) {
    cmx_stats s{};
    *rc_out = CMX_OK;
    for (cmx_handle *c : g->children) {
        cmx_stats t{};
        int rc = cmx_get_stats(c, &t);
        if (rc) { *rc_out = group_fail(g, c, rc); return s; }
        s.frames += t.frames; s.kernel_launches += t.kernel_launches; s.deferred += t.deferred; s.pair_evals += t.pair_evals;
        s.hits_real += t.hits_real; s.hits_random += t.hits_random; s.h2d_bytes += t.h2d_bytes; s.batches += t.batches;
        s.gpu_ms_total = std::max(s.gpu_ms_total, t.gpu_ms_total);         // the devices work side by side
        s.gpu_ms_main = std::max(s.gpu_ms_main, t.gpu_ms_main);
        s.gpu_ms_search_real = std::max(s.gpu_ms_search_real, t.gpu_ms_search_real);
        s.gpu_ms_search_random = std::max(s.gpu_ms_search_random, t.gpu_ms_search_random);
        s.gpu_ms_reduce = std::max(s.gpu_ms_reduce, t.gpu_ms_reduce);
        s.host_submit_ms += t.host_submit_ms; s.host_wait_ms += t.host_wait_ms;
        s.volume_total += t.volume_total; s.sum_weights += t.sum_weights;
    }
    return s;
}

}  // namespace
