// cmx_pairs_host.inl -- host side of the molecule-pair path (included by cmx_b200.cu)

namespace {

PairGeom make_pair_geom(cmx_handle *h, const Geom &g) {
    PairGeom pg{};
    const double *a = g.m, *b = g.m + 3, *c = g.m + 6;
    double bxc[3] = {b[1] * c[2] - b[2] * c[1], b[2] * c[0] - b[0] * c[2], b[0] * c[1] - b[1] * c[0]};
    double cxa[3] = {c[1] * a[2] - c[2] * a[1], c[2] * a[0] - c[0] * a[2], c[0] * a[1] - c[1] * a[0]};
    double axb[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
    double det = std::fabs(axb[0] * c[0] + axb[1] * c[1] + axb[2] * c[2]);
    pg.w[0] = det / std::sqrt(bxc[0] * bxc[0] + bxc[1] * bxc[1] + bxc[2] * bxc[2]);
    pg.w[1] = det / std::sqrt(cxa[0] * cxa[0] + cxa[1] * cxa[1] + cxa[2] * cxa[2]);
    pg.w[2] = det / std::sqrt(axb[0] * axb[0] + axb[1] * axb[1] + axb[2] * axb[2]);
    double wmin = std::min(pg.w[0], std::min(pg.w[1], pg.w[2]));
    pg.half_wmin = (float)(0.5 * wmin * (1.0 - 1e-6));
    // offsets are O(molecule size), the anchor difference is rounded once from fp64: error << 1e-5 A
    double tau = 2e-5 + 4e-7 * (h->cut_eff + 8.0);
    pg.tau = (float)tau; pg.cut = (float)h->cut_eff;
    pg.cut_lo = (float)(h->cut_eff - tau); pg.cut_hi = (float)(h->cut_eff + tau);
    // anchor cells: about one reach wide (big cells keep the 32 lanes of a warp busy)
    double reach = h->cut_eff + h->cur->pairs.ra_sol_bound + h->cur->pairs.ra_solv_bound + 0.05;
    for (int k = 0; k < 3; ++k) pg.n[k] = std::min(64, std::max(1, (int)std::floor(pg.w[k] / reach)));
    return pg;
}

int pairs_create(cmx_handle *h) {
    const cmx_config &c = h->cfg;
    PairScratch &S = h->cur->pairs;
    if (c.solute_natomspermol > 512) return fail(h, CMX_ERR_ARG, "molecule-pair path: solute_natomspermol > 512 (use path=1)");
    if (c.solute_nmols >= (1 << 24) || c.solvent_nmols >= (1 << 24))
        return fail(h, CMX_ERR_ARG, "molecule-pair path: too many molecules for the deferred-item encoding");
    size_t nsm = c.solute_nmols, nvm = c.solvent_nmols;
    // the random phase runs in chunks of samples: bounded scratch (256 MB of lists) and grid.y <= 32768
    S.sample_chunk = (int)std::max<size_t>(1, std::min<size_t>(std::min<size_t>((size_t)std::max(1, h->P.nrand), 32768), (size_t)(256.0e6 / (36.0 * nvm))));
    size_t nrand = (size_t)S.sample_chunk;
    CK(cudaMalloc(&S.solv.anchor, sizeof(double) * 3 * nvm));
    CK(cudaMalloc(&S.solv.off, sizeof(float) * 3 * h->nv_atoms));
    CK(cudaMalloc(&S.solv.rad, sizeof(float) * nvm));
    if (c.autocorrelation) S.sol = S.solv;
    else {
        CK(cudaMalloc(&S.sol.anchor, sizeof(double) * 3 * nsm));
        CK(cudaMalloc(&S.sol.off, sizeof(float) * 3 * h->ns_atoms));
        CK(cudaMalloc(&S.sol.rad, sizeof(float) * nsm));
    }
    S.ncells_cap = 64 * 64 * 64;
    CK(cudaMalloc(&S.cell_count, sizeof(int) * (S.ncells_cap + 1)));
    CK(cudaMemset(S.cell_count, 0, sizeof(int) * (S.ncells_cap + 1)));
    CK(cudaMalloc(&S.cell_start, sizeof(int) * (S.ncells_cap + 1)));
    CK(cudaMalloc(&S.sorted_id, sizeof(int) * nvm));
    CK(cudaMalloc(&S.s_anchor, sizeof(double) * 3 * nvm));
    CK(cudaMalloc(&S.s_rad, sizeof(float) * nvm));
    CK(cudaMalloc(&S.s_anchor4, sizeof(float4) * nvm));
    CK(cudaMalloc(&S.ref_lists, sizeof(MdRec) * nrand * nvm));
    CK(cudaMalloc(&S.bulk_idx, sizeof(int) * nrand * nvm));
    CK(cudaMalloc(&S.n_bulk, sizeof(int) * nrand));
    double pairs_total = (double)nsm * (double)nvm + (double)nrand * (double)nvm;
    S.def_cap = (size_t)std::min(pairs_total, std::max(1048576.0, 64.0 * (double)nvm));
    CK(cudaMalloc(&S.deferred, sizeof(u64) * S.def_cap));
    CK(cudaMalloc(&S.def_count, sizeof(int) * 4));
    CK(cudaMemset(S.def_count, 0, sizeof(int) * 4));
    CK(cudaMalloc(&S.d_radii, sizeof(int) * 4));
    CK(cudaMemset(S.d_radii, 0, sizeof(int) * 4));
    CK(cudaHostAlloc(&S.h_radii, sizeof(float) * 4, cudaHostAllocDefault));
    std::memset(S.h_radii, 0, sizeof(float) * 4);
    return CMX_OK;
}

void pairs_release(cmx_handle *h) {
    PairScratch &S = h->cur->pairs;
    bool shared = S.sol.anchor == S.solv.anchor;
    if (S.solv.anchor) cudaFree(S.solv.anchor);
    if (S.solv.off) cudaFree(S.solv.off);
    if (S.solv.rad) cudaFree(S.solv.rad);
    if (!shared) { if (S.sol.anchor) cudaFree(S.sol.anchor); if (S.sol.off) cudaFree(S.sol.off); if (S.sol.rad) cudaFree(S.sol.rad); }
    void *ptrs[] = {S.s_anchor4, S.cell_count, S.cell_start, S.sorted_id, S.s_anchor, S.s_rad, S.ref_lists, S.bulk_idx, S.n_bulk, S.deferred, S.def_count, S.d_radii};
    for (void *p : ptrs) if (p) cudaFree(p);
    if (S.h_radii) cudaFreeHost(S.h_radii);
    S = PairScratch{};
}

// molecule preparation of the frame (anchors, offsets, radii); returns after enqueueing
int pairs_prep(cmx_handle *h, const float *d_solute, const float *d_solvent, const Geom &g) {
    const cmx_config &c = h->cfg;
    PairScratch &S = h->cur->pairs;
    CK(cudaMemsetAsync(S.d_radii, 0, sizeof(int) * 4, h->cur->stream));
    launch(h, k_mol_prep, dim3((c.solvent_nmols + 127) / 128), dim3(128), g, d_solvent, c.solvent_nmols, c.solvent_natomspermol,
           c.irefatom - 1, S.solv, S.d_radii + 1, S.d_radii + 2);
    if (!c.autocorrelation)
        launch(h, k_mol_prep, dim3((c.solute_nmols + 127) / 128), dim3(128), g, d_solute, c.solute_nmols, c.solute_natomspermol, 0,
               S.sol, S.d_radii + 0, (int *)nullptr);
    CK(cudaMemcpyAsync(S.h_radii, S.d_radii, sizeof(float) * 4, cudaMemcpyDeviceToHost, h->cur->stream));
    return CMX_OK;
}

int frame_pair_path(cmx_handle *h, const float *d_solute, const float *d_solvent, uint32_t frame, const Geom &g) {
    const cmx_config &c = h->cfg;
    PairScratch &S = h->cur->pairs;
    int rc = pairs_prep(h, d_solute, d_solvent, g);
    if (rc) return rc;
    if (!S.primed) {
        // first frame: the molecule radii are needed to size the anchor cells (one-time sync)
        CK(cudaStreamSynchronize(h->cur->stream));
        S.primed = true;
    }
    // radii feedback (pinned mirror; may lag one frame, the kernels use the exact device values)
    float ra_solv = S.h_radii[1], ra_sol = c.autocorrelation ? S.h_radii[1] : S.h_radii[0];
    S.ra_sol_bound = std::max(S.ra_sol_bound, ra_sol); S.ra_solv_bound = std::max(S.ra_solv_bound, ra_solv);
    PairGeom pg = make_pair_geom(h, g);
    size_t ncells = (size_t)pg.n[0] * pg.n[1] * pg.n[2];
    CK(cudaMemsetAsync(S.def_count, 0, sizeof(int) * 1, h->cur->stream));
    int nvm = c.solvent_nmols;
    launch(h, k_anchor_bin<false>, dim3((nvm + 127) / 128), dim3(128), g, pg, S.solv, nvm, S.cell_count, (const int *)nullptr,
           (int *)nullptr, (double *)nullptr, (float *)nullptr, (float4 *)nullptr);
    launch(h, k_scan_block, dim3(1), dim3(1024), (const int *)S.cell_count, S.cell_start, (int)(ncells + 1));
    launch(h, k_anchor_bin<true>, dim3((nvm + 127) / 128), dim3(128), g, pg, S.solv, nvm, S.cell_count, (const int *)S.cell_start,
           S.sorted_id, S.s_anchor, S.s_rad, S.s_anchor4);
    size_t smem = sizeof(float4) * CMX_PAIR_WARPS * c.solute_natomspermol + sizeof(int) * CMX_PAIR_WARPS * 64;
    // (measured on C3: splitting a molecule's cells over several warps lowers the queue fill and is slower)
    int nsplit = c.solute_nmols >= h->num_sms * 8 ? 1 : std::max(1, std::min(8, (h->num_sms * 8 + c.solute_nmols - 1) / c.solute_nmols));
    long long ntask = (long long)c.solute_nmols * nsplit;
    int nblk = (int)std::min<long long>((ntask + CMX_PAIR_WARPS - 1) / CMX_PAIR_WARPS, (long long)h->num_sms * 32);
    u64 *pe = h->count_pairs ? h->d_stats.p : nullptr;
    cudaEvent_t ev = prof_begin(h);
    if (c.autocorrelation) {
        k_pairs<true><<<nblk, CMX_PAIR_WARPS * 32, smem, h->cur->stream>>>(g, pg, h->P, d_solute, d_solvent, S.sol, S.solv, S.cell_start,
            S.sorted_id, S.s_anchor, S.s_rad, S.s_anchor4, S.d_radii + 1, S.deferred, S.def_count, S.def_cap, pe, nsplit);
    } else {
        k_pairs<false><<<nblk, CMX_PAIR_WARPS * 32, smem, h->cur->stream>>>(g, pg, h->P, d_solute, d_solvent, S.sol, S.solv, S.cell_start,
            S.sorted_id, S.s_anchor, S.s_rad, S.s_anchor4, S.d_radii + 1, S.deferred, S.def_count, S.def_cap, pe, nsplit);
    }
    h->stats.kernel_launches++;
    prof_end(h, ev);
    // real-phase ambiguous pairs first, so that the deferred list can be reused by every chunk of samples
    launch(h, k_pair_resolve, dim3(h->num_sms * 2), dim3(128), g, h->P, frame, 0, d_solute, d_solvent, (const int *)S.bulk_idx,
           (const int *)S.n_bulk, (const u64 *)S.deferred, (const int *)S.def_count, S.def_cap, (MdRec *)nullptr);
    launch(h, k_accumulate_stats, dim3(1), dim3(32), (const int *)S.def_count, (const int *)nullptr, h->d_stats.p);
    const int nrand = h->P.nrand;
    for (int s0 = 0; s0 < nrand; s0 += S.sample_chunk) {
        const int ns = std::min(S.sample_chunk, nrand - s0);
        CK(cudaMemsetAsync(S.def_count, 0, sizeof(int), h->cur->stream));
        launch(h, k_ref_lists, dim3((nvm + 127) / 128, ns), dim3(128), g, pg, h->P, frame, -1, s0, d_solute, d_solvent, S.sol, S.solv, S.ref_lists);
        launch(h, k_bulk_compact, dim3(ns), dim3(512), h->P, frame, s0, (const MdRec *)S.ref_lists, S.bulk_idx, S.n_bulk);
        long long total = (long long)ns * nvm;
        ev = prof_begin(h, 1);
        launch(h, k_pair_random, dim3((unsigned)((total + 127) / 128)), dim3(128), g, pg, h->P, frame, s0, ns, d_solute, d_solvent, S.sol,
               (const int *)(S.d_radii + 2), (const int *)S.bulk_idx, (const int *)S.n_bulk,
               c.keep_lists ? h->d_rand_list.p : (MdRec *)nullptr, S.deferred, S.def_count, S.def_cap);
        prof_end(h, ev);
        launch(h, k_pair_resolve, dim3(h->num_sms * 2), dim3(128), g, h->P, frame, s0, d_solute, d_solvent, (const int *)S.bulk_idx,
               (const int *)S.n_bulk, (const u64 *)S.deferred, (const int *)S.def_count, S.def_cap,
               c.keep_lists ? h->d_rand_list.p : (MdRec *)nullptr);
        launch(h, k_accumulate_stats, dim3(1), dim3(32), (const int *)S.def_count, (const int *)nullptr, h->d_stats.p);
    }
    launch(h, k_check_overflow, dim3(1), dim3(32), (const int *)(S.def_count + 1), h->cur->d_scalars.p + 8);
    return CMX_OK;
}

}  // namespace
