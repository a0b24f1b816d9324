// cmx_final.inl -- the step immediately after the path, on the device (the remainder):
//   cmx_final_results   finalresults! = _mddf_final_results! / _coordination_number_final_results! + renormalize!
//                       (src/results.jl:311-469) from the accumulators as they lie in HBM
//   cmx_contributions   contributions(R, SoluteGroup|SolventGroup; type) (src/tools/contributions.jl:70-248) and the rows
//                       of ResidueContributions (src/tools/residue_contributions.jl:157-215) for MANY groups at once:
//                       row sums of the (possibly multi-GB) group-count array by k_reduce_rows, normalisation and type
//                       conversion (:mddf, :coordination_number, :md_count, :kbi) in one emit kernel; only
//                       [n_groups][nbins] f64 travel to the host.
// Included at the end of cmx_b200.cu (after cmx_feed.inl: it reuses the row reduction).  fp64 throughout, the arithmetic
// of the reference statement by statement (division by the product `nmols * Q`, serial cumulative sums).

namespace {

struct FinalScalars {
    double volume_total, volume_domain, volume_bulk, density_solute, density_solvent, density_solvent_bulk, density_fix, Q;
};
// rows of the profile block written by k_final_profile: [FP_ROWS][nbins] f64
enum { FP_D = 0, FP_MD, FP_MDR, FP_CN, FP_CNR, FP_MDDF, FP_KB, FP_RDFC, FP_RDFCR, FP_SRDF, FP_SRDFR, FP_RDF, FP_KBRDF, FP_SHELL, FP_ROWS };

struct FinalArgs {
    int nbins, usecutoff, cn_only, ibulk;          // ibulk: setbin(dbulk + binstep/2, binstep), 1-based
    double binstep, nmols_solute, nmols_solvent, nsolv_samples, nrand, Q, volume_sum;
};

__device__ __forceinline__ double block_sum(double v, double *sh) {     // fixed order: deterministic
    const int t = threadIdx.x;
    sh[t] = v;
    __syncthreads();
    for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
        if (t < o) sh[t] = __dadd_rn(sh[t], sh[t + o]);
        __syncthreads();
    }
    const double r = sh[0];
    __syncthreads();
    return r;
}

// ONE block.  cnt / acc = the head of the accumulator block: md, md_random, rdf, rdf_random [4][nbins] (integer hits at
// frame weight w, + the fp64 twin when the weights varied) -- read in place, nothing else of the block is touched.
__global__ void __launch_bounds__(512)
k_final_profile(const u64 *__restrict__ cnt, const double *__restrict__ acc, double w, FinalArgs A, double *__restrict__ o,
                FinalScalars *__restrict__ sc) {
    auto emit = [&](int k) { return __dadd_rn(acc ? acc[k] : 0.0, cnt ? __dmul_rn(w, (double)cnt[k]) : 0.0); };
    __shared__ double sh[512];
    __shared__ FinalScalars S;
    const int nb = A.nbins, t = threadIdx.x, nt = blockDim.x;
    const double den_real = __dmul_rn(A.nmols_solute, A.Q), den_rand = __dmul_rn(A.nrand, A.Q);
    const double vtot = __ddiv_rn(A.volume_sum, A.Q);
    double s_dom = 0, s_rdf = 0, s_rdf_bulk = 0, s_shell_bulk = 0;
    for (int b = t; b < nb; b += nt) {
        const double rmin = __dmul_rn((double)b, A.binstep), rmax = __dadd_rn(rmin, A.binstep);
        o[FP_D * nb + b] = pow(__dmul_rn(0.5, __dadd_rn(__dmul_rn(__dmul_rn(rmax, rmax), rmax), __dmul_rn(__dmul_rn(rmin, rmin), rmin))), 1.0 / 3.0);   // shellradius, src/results.jl:272-275
        const double md = __ddiv_rn(emit(b), den_real), rdf = __ddiv_rn(emit(2 * nb + b), den_real);
        o[FP_MD * nb + b] = md; o[FP_RDFC * nb + b] = rdf;
        if (!A.cn_only) {
            const double mdr = __ddiv_rn(emit(nb + b), den_rand), rdfr = __ddiv_rn(emit(3 * nb + b), den_rand);
            const double shell = __dmul_rn(vtot, __ddiv_rn(rdfr, A.nsolv_samples));
            o[FP_MDR * nb + b] = mdr; o[FP_RDFCR * nb + b] = rdfr; o[FP_SHELL * nb + b] = shell;
            if (b + 1 < A.ibulk) s_dom += shell; else { s_rdf_bulk += rdf; s_shell_bulk += shell; }
            s_rdf += rdf;
        }
    }
    if (!A.cn_only) {
        s_dom = block_sum(s_dom, sh); s_rdf = block_sum(s_rdf, sh);
        s_rdf_bulk = block_sum(s_rdf_bulk, sh); s_shell_bulk = block_sum(s_shell_bulk, sh);
    }
    if (t == 0) {
        S.Q = A.Q; S.volume_total = vtot; S.volume_domain = s_dom;
        double n_bulk;
        if (!A.usecutoff) { S.volume_bulk = __dsub_rn(vtot, s_dom); n_bulk = __dsub_rn(A.nsolv_samples, s_rdf); }
        else { S.volume_bulk = s_shell_bulk; n_bulk = s_rdf_bulk; }
        S.density_solvent = __ddiv_rn(A.nmols_solvent, vtot);
        S.density_solute = __ddiv_rn(A.nmols_solute, vtot);
        if (A.cn_only) { S.volume_domain = S.volume_bulk = 0.0; S.density_solvent_bulk = 0.0; S.density_fix = 1.0; }
        else { S.density_solvent_bulk = __ddiv_rn(n_bulk, S.volume_bulk); S.density_fix = __ddiv_rn(S.density_solvent_bulk, S.density_solvent); }
        *sc = S;
    }
    __syncthreads();
    if (!A.cn_only)
        for (int b = t; b < nb; b += nt) {      // renormalize!: the ideal-gas counts at the bulk density
            o[FP_MDR * nb + b] = __dmul_rn(o[FP_MDR * nb + b], S.density_fix);
            o[FP_RDFCR * nb + b] = __dmul_rn(o[FP_RDFCR * nb + b], S.density_fix);
        }
    __syncthreads();
    if (t < 4 && !(A.cn_only && (t & 1))) {     // the four cumulative sums, each serial like cumsum! (same rounding)
        const int src = t == 0 ? FP_MD : t == 1 ? FP_MDR : t == 2 ? FP_RDFC : FP_RDFCR;
        const int dst = t == 0 ? FP_CN : t == 1 ? FP_CNR : t == 2 ? FP_SRDF : FP_SRDFR;
        double run = 0.0;
        for (int b = 0; b < nb; ++b) { run = __dadd_rn(run, o[src * nb + b]); o[dst * nb + b] = run; }
    }
    __syncthreads();
    const double kfac = __dmul_rn(6.022140857e23 / 1e24, __ddiv_rn(1.0, S.density_solvent_bulk));
    for (int b = t; b < nb; b += nt) {
        if (A.cn_only) {
            o[FP_MDR * nb + b] = o[FP_CNR * nb + b] = o[FP_MDDF * nb + b] = o[FP_KB * nb + b] = 0.0;
            o[FP_RDFCR * nb + b] = o[FP_SRDFR * nb + b] = o[FP_RDF * nb + b] = o[FP_KBRDF * nb + b] = o[FP_SHELL * nb + b] = 0.0;
            continue;
        }
        const double mdr = o[FP_MDR * nb + b], rdfr = o[FP_RDFCR * nb + b];
        o[FP_MDDF * nb + b] = mdr > 0.0 ? __ddiv_rn(o[FP_MD * nb + b], mdr) : 0.0;
        o[FP_RDF * nb + b] = rdfr > 0.0 ? __ddiv_rn(o[FP_RDFC * nb + b], rdfr) : 0.0;
        o[FP_KB * nb + b] = __dmul_rn(kfac, __dsub_rn(o[FP_CN * nb + b], o[FP_CNR * nb + b]));
        o[FP_KBRDF * nb + b] = __dmul_rn(kfac, __dsub_rn(o[FP_SRDF * nb + b], o[FP_SRDFR * nb + b]));
    }
}

// contributions: one warp per group.  sel / sel_r = row sums of the group (integer hits + fp64 part), scale* = frame
// weight of the integer part (w or w/2), den* = nmols*Q | nrand*Q; prof = the profile block of k_final_profile.
// type: 0 :mddf, 1 :coordination_number, 2 :md_count, 3 :kbi
__global__ void __launch_bounds__(256)
k_contrib_emit(const u64 *__restrict__ sel, const double *__restrict__ sel_acc, const u64 *__restrict__ selr, const double *__restrict__ selr_acc,
               int n_groups, int nb, int type, double scale, double den_real, double den_rand, const double *__restrict__ prof,
               const FinalScalars *__restrict__ sc, double *__restrict__ out) {
    const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (g >= n_groups) return;
    const double fix = sc->density_fix;
    const double kfac = __dmul_rn(6.022140857e23 / 1e24, __ddiv_rn(1.0, sc->density_solvent_bulk));
    double run = 0.0, run_r = 0.0;
    for (int b0 = 0; b0 < nb; b0 += 32) {
        const int b = b0 + lane;
        const size_t k = (size_t)g * nb + b;
        double v = 0.0, vr = 0.0;
        if (b < nb) {
            v = __ddiv_rn(__dadd_rn(sel_acc ? sel_acc[k] : 0.0, __dmul_rn(scale, (double)sel[k])), den_real);
            if (type == 3) vr = __dmul_rn(__ddiv_rn(__dadd_rn(selr_acc ? selr_acc[k] : 0.0, __dmul_rn(scale, (double)selr[k])), den_rand), fix);
        }
        if (type == 0) {
            if (b < nb) { const double mdr = prof[FP_MDR * nb + b]; out[k] = mdr == 0.0 ? 0.0 : __ddiv_rn(v, mdr); }
        } else if (type == 2) {
            if (b < nb) out[k] = v;
        } else {                                  // cumulative sums over the bins: warp scan + carry
            double c = v, cr = vr;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double u = __shfl_up_sync(0xffffffffu, c, o), ur = __shfl_up_sync(0xffffffffu, cr, o);
                if (lane >= o) { c = __dadd_rn(c, u); cr = __dadd_rn(cr, ur); }
            }
            c = __dadd_rn(c, run); cr = __dadd_rn(cr, run_r);
            run = __shfl_sync(0xffffffffu, c, 31); run_r = __shfl_sync(0xffffffffu, cr, 31);
            if (b < nb) out[k] = type == 1 ? c : __dmul_rn(kfac, __dsub_rn(c, cr));
        }
    }
}

static_assert(sizeof(FinalScalars) == 8 * sizeof(double), "FinalScalars lives in a DevBuf<double> of 8");

// profile block + scalars on the device (left in fs), scalars also on the host
int final_profile_device(cmx_handle *h, double sum_weights, double volume_sum, FinalScalars &host_sc, FinalArgs &A) {
    int rc = cmx_sync(h); if (rc) return rc;
    if (!h->feed) h->feed = new cmx_feed();
    cmx_feed &F = *h->feed;
    const double Q = sum_weights > 0 ? sum_weights : h->sum_weights;
    if (!(Q > 0)) return fail(h, CMX_ERR_STATE, "cmx_final_results: no frames were accumulated (sum of the frame weights is zero)");
    A.nbins = h->nbins; A.usecutoff = h->cfg.usecutoff; A.cn_only = h->cfg.coordination_number_only; A.binstep = h->cfg.binstep;
    A.ibulk = std::max(1, (int)std::ceil((h->cfg.dbulk + 0.5 * h->cfg.binstep) / h->cfg.binstep));
    A.nmols_solute = h->cfg.solute_nmols; A.nmols_solvent = h->cfg.solvent_nmols;
    A.nsolv_samples = h->cfg.autocorrelation ? h->cfg.solvent_nmols - 1 : h->cfg.solvent_nmols;      // set_samples, src/results.jl:230-237
    A.nrand = h->cfg.n_random_samples; A.Q = Q; A.volume_sum = volume_sum > 0 ? volume_sum : h->volume_total;
    CK(F.fin_prof.ensure((size_t)FP_ROWS * h->nbins)); CK(F.fin_sc.ensure(8));
    // emit_valid: the f64 block (cmx_counters_device_f64, possibly all-reduced by the caller) is the authoritative one
    launch(h, k_final_profile, dim3(1), dim3(512), (const u64 *)(h->emit_valid ? nullptr : h->d_cnt.p),
           (const double *)(h->emit_valid ? h->d_emit.p : h->acc_used ? h->d_acc.p : nullptr), h->have_weight ? h->w0 : 1.0, A, F.fin_prof.p, reinterpret_cast<FinalScalars *>(F.fin_sc.p));
    CK(cudaStreamSynchronize(h->cur->stream));
    CK(cudaMemcpy(&host_sc, F.fin_sc.p, sizeof host_sc, cudaMemcpyDeviceToHost));
    return CMX_OK;
}

}  // namespace

extern "C" {

int32_t cmx_final_results(cmx_handle *h, double sum_weights, double volume_sum, cmx_final *out) {
    if (!h || !out) return CMX_ERR_ARG;
    if (is_group(h)) {
        int rc = group_merge(h); if (rc) return rc;
        rc = cmx_final_results(h->children[0], sum_weights > 0 ? sum_weights : h->children[0]->sum_weights, volume_sum, out);
        return rc ? group_fail(h, h->children[0], rc) : CMX_OK;
    }
    FinalScalars sc; FinalArgs A;
    int rc = final_profile_device(h, sum_weights, volume_sum, sc, A); if (rc) return rc;
    const size_t nb = h->nbins;
    std::vector<double> tmp((size_t)FP_ROWS * nb);
    CK(cudaMemcpy(tmp.data(), h->feed->fin_prof.p, sizeof(double) * tmp.size(), cudaMemcpyDeviceToHost));
    double *dst[FP_ROWS] = {out->d, out->md_count, out->md_count_random, out->coordination_number, out->coordination_number_random, out->mddf,
                            out->kb, out->rdf_count, out->rdf_count_random, out->sum_rdf_count, out->sum_rdf_count_random, out->rdf, out->kb_rdf,
                            out->volume_shell};
    for (int r = 0; r < FP_ROWS; ++r) if (dst[r]) std::memcpy(dst[r], tmp.data() + (size_t)r * nb, sizeof(double) * nb);
    out->nbins = h->nbins;
    out->volume_total = sc.volume_total; out->volume_domain = sc.volume_domain; out->volume_bulk = sc.volume_bulk;
    out->density_solute = sc.density_solute; out->density_solvent = sc.density_solvent; out->density_solvent_bulk = sc.density_solvent_bulk;
    out->density_fix = sc.density_fix; out->sum_weights = sc.Q;
    return CMX_OK;
}

int32_t cmx_contributions(cmx_handle *h, int32_t side, int32_t type, double sum_weights, double volume_sum, int32_t n_groups,
                          const int32_t *offsets, const int32_t *rows, double *out) {
    if (!h) return CMX_ERR_ARG;
    if (is_group(h)) {
        int rc = group_merge(h); if (rc) return rc;
        rc = cmx_contributions(h->children[0], side, type, sum_weights > 0 ? sum_weights : h->children[0]->sum_weights, volume_sum, n_groups, offsets, rows, out);
        return rc ? group_fail(h, h->children[0], rc) : CMX_OK;
    }
    if (side < 0 || side > 1 || type < 0 || type > 3 || n_groups < 1 || !offsets || !out) return fail(h, CMX_ERR_ARG, "cmx_contributions: invalid argument");
    if (h->cfg.coordination_number_only && (type == 0 || type == 3))
        return fail(h, CMX_ERR_STATE, "cmx_contributions: :mddf and :kbi need the ideal-gas counts (handle was created with coordination_number_only)");
    // an autocorrelation keeps ONE set of group counts: the solvent's are the solute's (src/results.jl:341-343)
    const int which = (h->cfg.autocorrelation || side == 0) ? 0 : 2;
    FinalScalars sc; FinalArgs A;
    int rc = final_profile_device(h, sum_weights, volume_sum, sc, A); if (rc) return rc;
    cmx_feed &F = *h->feed;
    const size_t nb = h->nbins, nout = (size_t)n_groups * nb;
    if (type == 3) {        // :kbi needs the group's ideal-gas counts too: reduce them first, keep them aside
        rc = reduce_rows_device(h, which + 1, n_groups, offsets, rows); if (rc) return rc;
        CK(F.fin_selr.ensure(nout)); if (F.red_has_acc) CK(F.fin_selr_acc.ensure(nout));
        CK(cudaMemcpyAsync(F.fin_selr.p, F.red_cnt.p, sizeof(u64) * nout, cudaMemcpyDeviceToDevice, h->cur->stream));
        if (F.red_has_acc) CK(cudaMemcpyAsync(F.fin_selr_acc.p, F.red_acc.p, sizeof(double) * nout, cudaMemcpyDeviceToDevice, h->cur->stream));
    }
    rc = reduce_rows_device(h, which, n_groups, offsets, rows); if (rc) return rc;
    const double w = h->have_weight ? h->w0 : 1.0;
    const double scale = h->cfg.autocorrelation ? w / 2 : w;       // src/update_counters.jl:52-53
    launch(h, k_contrib_emit, dim3((unsigned)((n_groups + 7) / 8)), dim3(256), (const u64 *)F.red_cnt.p,
           (const double *)(F.red_has_acc ? F.red_acc.p : nullptr), (const u64 *)(type == 3 ? F.fin_selr.p : nullptr),
           (const double *)(type == 3 && F.red_has_acc ? F.fin_selr_acc.p : nullptr), (int)n_groups, (int)nb, (int)type, scale,
           A.nmols_solute * A.Q, A.nrand * A.Q, (const double *)F.fin_prof.p, reinterpret_cast<const FinalScalars *>(F.fin_sc.p), F.red_out.p);
    CK(cudaStreamSynchronize(h->cur->stream));
    CK(cudaMemcpy(out, F.red_out.p, sizeof(double) * nout, cudaMemcpyDeviceToHost));
    return CMX_OK;
}

}  // extern "C"
