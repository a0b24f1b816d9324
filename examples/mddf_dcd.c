/*
 * mddf_dcd.c -- the C ABI of libcmx_b200.so from plain C, no Python and no Julia: counts of a minimum-distance
 * distribution for a solute (one molecule, atoms [s0, s0+ns)) and a solvent (atoms [v0, v0+nv), napm atoms per
 * molecule) over every frame of a NAMD/CHARMM DCD file, with the library's own reader threads feeding the GPU.
 *
 *   gcc -O2 -Iinclude examples/mddf_dcd.c -Lcomplexmixtures.jl_b200 -lcmx_b200 -Wl,-rpath,$PWD/complexmixtures.jl_b200 \
 *       -Wl,--allow-shlib-undefined -o mddf_dcd
 *   ./mddf_dcd trajectory.dcd 1 1463 1479 2534 14      (protein x TMAO of the reference's NAMD example, 1-based)
 *
 * This is the call sequence a ComplexMixtures.jl maintainer makes through ccall (julia/CMXB200.jl): cmx_create,
 * cmx_run_dcd, cmx_finish -- and, for the step after the path, cmx_final_results (finalresults!, src/results.jl:311-469)
 * and cmx_contributions (contributions(R, SoluteGroup(...); type = :coordination_number)) evaluated on the device.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "cmx_b200.h"

#define CHECK(call, h)                                                                    \
    do {                                                                                  \
        int rc_ = (call);                                                                 \
        if (rc_ != CMX_OK) { fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, cmx_last_error(h)); return 1; } \
    } while (0)

int main(int argc, char **argv) {
    if (argc < 7) { fprintf(stderr, "usage: %s traj.dcd solute_first n_solute solvent_first n_solvent atoms_per_solvent_molecule [cutoff dbulk]\n", argv[0]); return 2; }
    const int s0 = atoi(argv[2]), ns = atoi(argv[3]), v0 = atoi(argv[4]), nv = atoi(argv[5]), napm = atoi(argv[6]);
    const double cutoff = argc > 7 ? atof(argv[7]) : 10.0, dbulk = argc > 8 ? atof(argv[8]) : 8.0;
    if (ns < 1 || nv < 1 || napm < 1 || nv % napm) { fprintf(stderr, "bad selection sizes\n"); return 2; }

    cmx_dcd *dcd = NULL;
    cmx_dcd_info info;
    if (cmx_dcd_open(argv[1], &dcd, &info) != CMX_OK) { fprintf(stderr, "%s\n", cmx_dcd_last_error()); return 1; }
    printf("%s: %lld atoms, %lld frames\n", argv[1], (long long)info.natoms, (long long)info.nframes);

    cmx_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.struct_size = (int32_t)sizeof cfg;
    cfg.solute_nmols = 1; cfg.solute_natomspermol = ns;
    cfg.solvent_nmols = nv / napm; cfg.solvent_natomspermol = napm;
    cfg.irefatom = 1; cfg.usecutoff = 1; cfg.n_random_samples = 10;
    cfg.n_groups_solute = ns; cfg.n_groups_solvent = napm;          /* per-atom contributions (no custom groups) */
    cfg.cutoff = cutoff; cfg.dbulk = dbulk; cfg.binstep = 0.02; cfg.seed = 321;
    cmx_handle *h = NULL;
    if (cmx_create(&cfg, &h) != CMX_OK) { fprintf(stderr, "cmx_create: %s\n", cmx_last_error(NULL)); return 1; }

    int32_t *sol = malloc(sizeof(int32_t) * (size_t)ns), *solv = malloc(sizeof(int32_t) * (size_t)nv);
    int64_t *frames = malloc(sizeof(int64_t) * (size_t)info.nframes);
    for (int k = 0; k < ns; ++k) sol[k] = s0 + k;                   /* 1-based positions in the file, as AtomSelection.indices */
    for (int k = 0; k < nv; ++k) solv[k] = v0 + k;
    for (int64_t k = 0; k < info.nframes; ++k) frames[k] = k;
    CHECK(cmx_run_dcd(h, dcd, sol, solv, frames, NULL, info.nframes, 2), h);

    const int nbins = (int)(cutoff / 0.02 + 0.999999);
    double *md = calloc((size_t)nbins, sizeof(double)), *md_r = calloc((size_t)nbins, sizeof(double));
    cmx_counters out;
    memset(&out, 0, sizeof out);
    out.md_count = md; out.md_count_random = md_r;                  /* NULL members are skipped: no per-atom arrays here */
    CHECK(cmx_finish(h, &out), h);
    double hits = 0, hits_r = 0;
    for (int b = 0; b < out.nbins; ++b) { hits += md[b]; hits_r += md_r[b]; }
    cmx_stats st;
    CHECK(cmx_get_stats(h, &st), h);
    printf("nbins %d  solvent molecules within %.1f A per frame: %.3f  (ideal-gas reference: %.3f per sample)  volume %.1f A^3\n",
           out.nbins, cutoff, hits / out.sum_weights, hits_r / out.sum_weights / cfg.n_random_samples, out.volume_total / out.sum_weights);
    printf("%lld frames, %lld kernel launches, %.1f ms on the device, %lld molecules resolved in fp64\n",
           (long long)st.frames, (long long)st.kernel_launches, st.gpu_ms_total, (long long)st.deferred);
    /* finalresults! on the device: the mddf and the KB integral; contributions of the first half of the solute's atoms */
    double *mddf = calloc((size_t)nbins, sizeof(double)), *kb = calloc((size_t)nbins, sizeof(double)), *d = calloc((size_t)nbins, sizeof(double));
    cmx_final fin;
    memset(&fin, 0, sizeof fin);
    fin.mddf = mddf; fin.kb = kb; fin.d = d;
    CHECK(cmx_final_results(h, 0.0, 0.0, &fin), h);
    int bmax = 0;
    for (int b = 1; b < fin.nbins; ++b) if (mddf[b] > mddf[bmax]) bmax = b;
    printf("mddf peak %.3f at %.2f A, KB integral %.1f cm^3/mol, bulk density %.5f /A^3, domain volume %.1f A^3\n",
           mddf[bmax], d[bmax], kb[fin.nbins - 1], fin.density_solvent_bulk, fin.volume_domain);
    const int nhalf = ns / 2 > 0 ? ns / 2 : 1;
    int32_t offsets[2] = {0, nhalf}, *rows = malloc(sizeof(int32_t) * (size_t)nhalf);
    for (int k = 0; k < nhalf; ++k) rows[k] = k;                   /* 0-based rows of solute_group_count */
    double *cn = calloc((size_t)nbins, sizeof(double));
    CHECK(cmx_contributions(h, 0 /* solute */, 1 /* :coordination_number */, 0.0, 0.0, 1, offsets, rows, cn), h);
    printf("coordination number of the first %d solute atoms at the cutoff: %.3f\n", nhalf, cn[fin.nbins - 1]);
    cmx_destroy(h); cmx_dcd_close(dcd);
    free(sol); free(solv); free(frames); free(md); free(md_r); free(mddf); free(kb); free(d); free(rows); free(cn);
    return 0;
}
