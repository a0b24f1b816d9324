/*
 * cmx_b200.h -- C ABI of libcmx_b200.so, the B200 (sm_100a) replacement of the per-frame
 * minimum-distance hot path of ComplexMixtures.jl.
 *
 * The reference has no FFI/plugin interface: the path is reached by ordinary Julia calls.
 * The seam this ABI replaces is the body of the chunk task of mddf(), reference
 * src/mddf.jl:288-337:
 *
 *     build_particle_system(...)            src/minimum_distances.jl:157-176   -> cmx_create
 *     Buffer(...), Result(...)              src/mddf.jl:18-26, results.jl:124  -> cmx_create
 *     @. buff.solute_read = trajectory.x_solute ; unitcell      mddf.jl:307-321 -> cmx_acquire_frame_buffer
 *     update!(system; unitcell)             src/mddf.jl:326                    -> cmx_submit_frame
 *     mddf_frame!(r_chunk, system, buff, options, w, RNG)       mddf.jl:329,361-429 -> cmx_submit_frame
 *     coordination_number_frame!(...)       src/mddf.jl:331,438-471            -> cmx_submit_frame (coordination_number_only)
 *     sum!(R, r_chunk)                      src/mddf.jl:336, results.jl:629-649 -> cmx_finish (+ cmx_counters_device for the NCCL all-reduce)
 *
 * Conventions: every function returns int32 status (0 = CMX_OK); cmx_last_error() gives the
 * message; no exception crosses the boundary; all pointers are caller-owned unless stated.
 * Plain C types only -- no torch / CUDA types in any signature.
 */
#ifndef CMX_B200_H
#define CMX_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CMX_OK 0
#define CMX_ERR_ARG 1      /* invalid argument / configuration                         */
#define CMX_ERR_CUDA 2     /* CUDA runtime error (message has the CUDA error string)   */
#define CMX_ERR_CELL 3     /* unit cell narrower than 2*cutoff in some direction       */
#define CMX_ERR_STATE 4    /* call sequence error (e.g. submit without acquire)        */
#define CMX_ERR_IO 5       /* trajectory file error (open / short read / bad header)   */
#define CMX_ERR_MEMORY 6   /* the counters and scratch of this problem do not fit the device memory (cmx_create) */

typedef struct cmx_handle cmx_handle;

/* Flattened Options (src/Options.jl:6-42) + AtomSelection (src/AtomSelection.jl:48-59) +
 * TrajectoryMetaData (src/Trajectory.jl:180-187). */
typedef struct cmx_config {
    int32_t struct_size;              /* = sizeof(cmx_config), ABI check                        */
    int32_t device;                   /* CUDA device ordinal                                     */
    int32_t solute_nmols;             /* AtomSelection.nmols                                     */
    int32_t solute_natomspermol;      /* AtomSelection.natomspermol                              */
    int32_t solvent_nmols;
    int32_t solvent_natomspermol;
    int32_t autocorrelation;          /* solute and solvent are the same selection               */
    int32_t irefatom;                 /* 1-based, resolved (TrajectoryMetaData.irefatom)          */
    int32_t usecutoff;                /* Options.usecutoff                                       */
    int32_t n_random_samples;         /* Options.n_random_samples                                */
    int32_t coordination_number_only; /* mddf(...; coordination_number_only)                     */
    int32_t lcell;                    /* Options.lcell: accepted, a hint only                    */
    int32_t n_groups_solute;          /* rows of solute_group_count (natomspermol or #groups)     */
    int32_t n_groups_solvent;
    int32_t path;                     /* 0 auto, 1 grid path (large solute molecule), 2 molecule-pair path */
    int32_t ring_slots;               /* pinned staging slots of acquire/submit (0 -> one per frame of every batch in flight + 1) */
    int32_t keep_lists;               /* keep per-frame minimum-distance lists for cmx_read_*    */
    int32_t group_lanes;              /* reserved (ignored: the search works on 32-query tiles)  */
    int32_t n_streams;                /* batches in flight on separate compute streams (0 -> auto)             */
    int32_t batch_frames;             /* frames per kernel launch on the grid path (0 -> auto: 16 for small systems ... 1 above 4 M atoms per frame; <= 32) */
    double cutoff;                    /* Options.cutoff                                          */
    double dbulk;                     /* Options.dbulk                                           */
    double binstep;                   /* Options.binstep                                         */
    uint64_t seed;                    /* Options.seed (Philox key)                               */
    /* custom groups: CSR "position in the selection -> group ids" (NULL = per atom type,
     * src/update_counters.jl:22-36) */
    const int32_t *solute_group_offsets, *solute_group_ids;
    const int32_t *solvent_group_offsets, *solvent_group_ids;
    /* Several GPUs behind ONE handle (the reference's one mddf() call uses the whole machine: nchunks = nthreads,
     * src/parallel_setup.jl:7-57): n_devices > 1 and device_ids[n_devices] make cmx_create build one device context per
     * entry (`device` is then ignored); frames are dealt to them in submission order (k mod n_devices), the native feeds
     * run one reader/consumer team per device, and cmx_finish / cmx_counters_device / cmx_reduce_groups first sum the
     * per-device accumulators onto the first device over peer access (sum!, src/results.jl:629-649).  The same ordinal
     * may be listed more than once (two contexts on one GPU).  n_devices <= 1: one context on `device`. */
    int32_t n_devices, reserved1;
    const int32_t *device_ids;
} cmx_config;

/* Output of cmx_finish: f64 arrays laid out like Result (src/results.jl:73-104); the group
 * arrays are row-major [n_groups][nbins] (one row per group = one Vector{Float64}). */
typedef struct cmx_counters {
    int32_t nbins, n_groups_solute, n_groups_solvent, reserved;
    double *md_count, *md_count_random;
    double *rdf_count, *rdf_count_random;
    double *solute_group_count, *solute_group_count_random;
    double *solvent_group_count, *solvent_group_count_random;
    double volume_total;              /* sum_f w_f * det(cell_f), src/mddf.jl:350-352             */
    double sum_weights;               /* sum of the weights of the frames submitted               */
} cmx_counters;

/* MinimumDistance (src/minimum_distances.jl:12-19); i, j are 1-based as in the reference
 * (i within the current solute molecule, j in the solvent selection), 0 when empty. */
typedef struct cmx_md {
    int32_t within_cutoff, i, j, ref_atom_within_cutoff;
    double d, d_ref_atom;
} cmx_md;

typedef struct cmx_stats {
    int64_t frames;            /* frames submitted                                               */
    int64_t kernel_launches;   /* kernels of this library launched so far                        */
    int64_t deferred;          /* molecules re-evaluated by the exact fp64 resolve kernel
                                  (fp32 near-ties / cutoff-edge cases), all phases               */
    int64_t pair_evals;        /* atom-pair distances evaluated by the search kernels (0 unless
                                  cmx_set_option("count_pairs", 1))                              */
    int64_t hits_real, hits_random; /* molecules within the cutoff counted so far                 */
    int64_t h2d_bytes;         /* bytes copied host->device by cmx_submit_frame                  */
    double gpu_ms_total;       /* device time of all frames (CUDA events on the compute stream)  */
    double gpu_ms_main;        /* device time spent in the search kernels (needs option "profile") */
    double gpu_ms_search_real;   /* ... of which: real-phase search / pair kernel                  */
    double gpu_ms_search_random; /* ... of which: random-phase search kernel                       */
    double gpu_ms_reduce;        /* device time of the last cmx_reduce_groups row-sum kernel (option "profile") */
    double host_submit_ms;       /* wall time spent inside the submit / run_* calls ...                        */
    double host_wait_ms;         /* ... of which blocked on the device (back-pressure: ring slot or batch context busy) */
    int64_t batches;             /* kernel-sequence launches (one per batch of frames on the grid path)        */
    double volume_total;         /* running sum_f w_f * det(cell_f) and sum_f w_f of the frames submitted so far: */
    double sum_weights;          /* the two host-side scalars a multi-process driver adds to its all-reduce     */
} cmx_stats;

const char *cmx_version(void);
const char *cmx_last_error(cmx_handle *h);   /* h may be NULL: error of the last failed cmx_create */

int32_t cmx_create(const cmx_config *cfg, cmx_handle **out);
int32_t cmx_destroy(cmx_handle *h);

/* Pointers into the next free pinned staging slot: the host reader writes fp32 xyz triplets in
 * place (solute_xyz[3*n_solute_atoms], solvent_xyz[3*n_solvent_atoms]); for an autocorrelation
 * only solvent_xyz is used and *solute_xyz aliases it.  Blocks while the ring is full. */
int32_t cmx_acquire_frame_buffer(cmx_handle *h, float **solute_xyz, float **solvent_xyz);

/* Enqueue the frame in the acquired slot: async H2D + all kernels; returns immediately.
 * cell: column-major 3x3, columns = lattice vectors (convert_unitcell, src/Trajectory.jl:72-81).
 * frame_index keys the Philox stream (seed, frame_index, sample, molecule), so results do not
 * depend on which GPU processes the frame.  weight = 0 frames must not be submitted. */
int32_t cmx_submit_frame(cmx_handle *h, int64_t frame_index, double weight, const double cell[9]);

/* Same, for coordinates already resident in device memory (device pointers to fp32 xyz).  The arrays
 * must stay valid and unmodified until cmx_sync: up to n_streams batches of batch_frames frames are in flight. */
int32_t cmx_submit_frame_device(cmx_handle *h, const float *d_solute_xyz, const float *d_solvent_xyz,
                                int64_t frame_index, double weight, const double cell[9]);

int32_t cmx_sync(cmx_handle *h);

/* Device pointer to the contiguous block of integer (uint64) run accumulators, for the single
 * all-reduce(sum) over the GPUs that shared the frames (src/mddf.jl:336 -> sum!).  Valid after
 * cmx_sync.  Layout: md, md_random, rdf, rdf_random [nbins each], solute_group,
 * solute_group_random [n_groups_solute*nbins each], solvent_group, solvent_group_random. */
int32_t cmx_counters_device(cmx_handle *h, void **device_ptr, int64_t *n_uint64);

/* The same block as f64 with the frame weights applied (what cmx_finish will write), on the device: the payload of
 * the all-reduce when the frame weights of the ranks are not all one and the same number (then the integer blocks
 * cannot be summed: every rank scales them by its own weight).  After an in-place all-reduce of this array the next
 * cmx_finish writes it out as is.  Valid until the next submit / reset. */
int32_t cmx_counters_device_f64(cmx_handle *h, double **device_ptr, int64_t *n_f64);

/* Syncs and writes the f64 counters (buffers pre-allocated by the caller; NULL members skipped: with all four group arrays
 * NULL only md / rdf (+ random) are converted and copied -- the group arrays stay on the device for cmx_contributions). */
int32_t cmx_finish(cmx_handle *h, cmx_counters *out);

/* Parity hooks: system.list after minimum_distances! (src/minimum_distances.jl:147) of the LAST
 * submitted frame (needs keep_lists=1).  isolute is 0-based; sample is the random sample index. */
int32_t cmx_read_minimum_distances(cmx_handle *h, int32_t isolute, cmx_md *out /*[solvent_nmols]*/);
int32_t cmx_read_random_minimum_distances(cmx_handle *h, int32_t sample, cmx_md *out /*[solvent_nmols]*/);

/* Page-locked host memory for the caller's result arrays (cmx_finish then copies at full PCIe speed). */
int32_t cmx_alloc_pinned(void **ptr, int64_t bytes);
int32_t cmx_free_pinned(void *ptr);

/* ---- frame feed: native DCD reader ------------------------------------------------
 * Replaces the host-side record read + per-atom gather of nextframe!(::NamdDCD)
 * (src/trajectory_formats/NamdDCD.jl:141-169), getunitcell (:175-188) and the frame count from the
 * file size (:211-230).  The cmx_dcd_* functions are pure host code (no GPU needed). */
typedef struct cmx_dcd cmx_dcd;
typedef struct cmx_dcd_info {
    int64_t natoms;             /* atoms per frame in the file                                   */
    int64_t nframes;            /* counted from the file size, not taken from the header         */
    int64_t first_frame_offset; /* byte offset of the first frame                                */
    int64_t frame_bytes;        /* 56 + 3*(8 + 4*natoms): cell record + X, Y, Z records          */
} cmx_dcd_info;

const char *cmx_dcd_last_error(void);
int32_t cmx_dcd_open(const char *path, cmx_dcd **out, cmx_dcd_info *info);
int32_t cmx_dcd_close(cmx_dcd *d);
/* One frame (0-based) to host arrays: x, y, z [natoms] (any may be NULL) and the unit cell as the
 * column-major 3x3 matrix cmx_submit_frame takes.  Thread-safe (pread). */
int32_t cmx_dcd_read_frame(cmx_dcd *d, int64_t iframe, float *x, float *y, float *z, double cell[9]);

/* The whole frame loop of the chunk task (src/mddf.jl:296-334) for a DCD file, in the library:
 * reader threads pread the raw frame records straight into a pinned ring, the raw records go to the
 * device in one async copy, a kernel gathers the selected atoms (indices are 1-based positions in
 * the file, as AtomSelection.indices) and the frame is enqueued like cmx_submit_frame does.
 * frames[k] = 0-based frame number in the file; the Philox frame key is frames[k] + 1 (the
 * reference's 1-based iframe), so the result equals the acquire/submit path on the same frames.
 * weights may be NULL (all 1); zero-weight frames must be left out by the caller.
 * For an autocorrelation solute_indices is ignored.  Returns after the last frame is enqueued. */
int32_t cmx_run_dcd(cmx_handle *h, cmx_dcd *d, const int32_t *solute_indices, const int32_t *solvent_indices,
                    const int64_t *frames, const double *weights, int64_t nframes, int32_t n_reader_threads);

/* ---- frame feed: native XTC reader (host code; SURVEY 8 f1, second format) -------------------------
 * Replaces the Chemfiles read of GROMACS XTC frames (src/trajectory_formats/ChemFiles.jl:112-138): frame index
 * built at open time, compressed coordinate block decoded to fp32 xyz triplets in Angstrom (nm x 10), unit cell
 * as the column-major 3x3 matrix cmx_submit_frame takes.  cmx_xtc_* are pure host code (errors through
 * cmx_dcd_last_error()); cmx_run_xtc is the XTC twin of cmx_run_dcd. */
typedef struct cmx_xtc cmx_xtc;
typedef struct cmx_xtc_info {
    int64_t natoms;
    int64_t nframes;            /* complete frames found by walking the file */
} cmx_xtc_info;
int32_t cmx_xtc_open(const char *path, cmx_xtc **out, cmx_xtc_info *info);
int32_t cmx_xtc_close(cmx_xtc *x);
/* frame (0-based) -> xyz[3*natoms] fp32 Angstrom (may be NULL), cell[9] (may be NULL), MD step and time (may be NULL) */
int32_t cmx_xtc_read_frame(cmx_xtc *x, int64_t iframe, float *xyz, double cell[9], int32_t *step, float *time);
/* The same frame decoded on CUDA device `device` (the decoder cmx_run_xtc uses: the host walks only the run codes of the
 * compressed bit stream, one device thread decodes each group of atoms); bit-identical to cmx_xtc_read_frame. */
int32_t cmx_xtc_read_frame_device(cmx_xtc *x, int64_t iframe, int32_t device, float *xyz, double cell[9]);
/* The frame loop for an XTC file (arguments as cmx_run_dcd): the reader threads read the compressed frames into the pinned
 * ring and walk their run codes (n_reader_threads = 0: half of the available cores, 2..8), one H2D per COMPRESSED frame,
 * decode + selection gather on the device (option "xtc_host_decode" = 1: decode in the reader threads instead). */
int32_t cmx_run_xtc(cmx_handle *h, cmx_xtc *x, const int32_t *solute_indices, const int32_t *solvent_indices,
                    const int64_t *frames, const double *weights, int64_t nframes, int32_t n_reader_threads);

/* ---- group reduction on the device ------------------------------------------------
 * The count stage of contributions()/ResidueContributions (src/tools/contributions.jl:70-248,
 * src/tools/residue_contributions.jl:157-215): out[g][b] = sum of the rows of one group-count array
 * that belong to group g, with the frame weights applied as in cmx_finish -- without reading the
 * (possibly multi-GB) per-atom array back to the host.
 * which: 0 solute_group_count, 1 solute_group_count_random, 2 solvent_group_count,
 * 3 solvent_group_count_random.  CSR group -> rows (0-based rows of that array).
 * out: host array [n_groups][nbins] f64.  Integer counts are summed exactly (order independent). */
int32_t cmx_reduce_groups(cmx_handle *h, int32_t which, int32_t n_groups, const int32_t *offsets,
                          const int32_t *rows, double *out);

/* ---- finalresults! and contributions on the device -----------------------------------
 * cmx_final_results: _mddf_final_results! / _coordination_number_final_results! + renormalize!
 * (src/results.jl:311-469) evaluated on the device from the accumulators, in fp64 and in the reference's operation
 * order (cumulative sums serial).  Every array pointer is caller-allocated [nbins] f64 and may be NULL.
 * sum_weights: Q = sum of the weights of the frames of the run (sum_frame_weights, src/results.jl:288-297); <= 0 means
 * "the frames this handle was given" (a multi-process driver passes the global sum after its all-reduce).
 * volume_sum: sum_f w_f det(cell_f); <= 0 means the handle's own. */
typedef struct cmx_final {
    int32_t nbins, reserved;
    double *d, *md_count, *md_count_random, *coordination_number, *coordination_number_random, *mddf, *kb;
    double *rdf_count, *rdf_count_random, *sum_rdf_count, *sum_rdf_count_random, *rdf, *kb_rdf;
    double *volume_shell;                                   /* Volume.shell                                  */
    double volume_total, volume_bulk, volume_domain;        /* Volume (src/results.jl:35-42)                 */
    double density_solute, density_solvent, density_solvent_bulk;   /* Density (:50-56)                      */
    double density_fix;                                     /* density.solvent_bulk / density.solvent (:375) */
    double sum_weights;                                     /* the Q that was used                           */
} cmx_final;
int32_t cmx_final_results(cmx_handle *h, double sum_weights, double volume_sum, cmx_final *out);

/* contributions(R, SoluteGroup|SolventGroup; type) (src/tools/contributions.jl:70-248) for n_groups groups at once --
 * also the matrix of ResidueContributions (src/tools/residue_contributions.jl:157-215) when the groups are residues:
 * the rows of the group-count array that belong to each group are summed on the device (exact integer sums), scaled
 * like finalresults! and converted to `type`; out: host array [n_groups][nbins] f64.
 * side: 0 solute groups, 1 solvent groups (an autocorrelation has one set: src/results.jl:341-343).
 * type: 0 :mddf (count / md_count_random, 0 where that is 0), 1 :coordination_number (cumulative sum),
 *       2 :md_count, 3 :kbi (units.Angs3tocm3permol / density.solvent_bulk * (cumsum(count) - cumsum(count_random))).
 * CSR group -> rows as in cmx_reduce_groups; sum_weights / volume_sum as in cmx_final_results. */
int32_t cmx_contributions(cmx_handle *h, int32_t side, int32_t type, double sum_weights, double volume_sum, int32_t n_groups,
                          const int32_t *offsets, const int32_t *rows, double *out);

int32_t cmx_get_stats(cmx_handle *h, cmx_stats *out);
int32_t cmx_reset(cmx_handle *h);                       /* zero all accumulators and statistics */
int32_t cmx_set_option(cmx_handle *h, const char *name, double value);

#ifdef __cplusplus
}
#endif
#endif /* CMX_B200_H */
