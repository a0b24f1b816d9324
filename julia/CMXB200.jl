# CMXB200.jl -- reference-side binding of libcmx_b200.so (what a ComplexMixtures.jl maintainer adds).
#
# NOT exercised in this repository's CI: Julia is not available in the build image.  The Python
# ctypes binding (complexmixtures.jl_b200/engine.py) drives exactly the same C ABI and is what the
# tests run; this file shows the `ccall` stubs and the replacement of the chunk loop of
# `mddf(trajectory, options; ...)` (ComplexMixtures.jl src/mddf.jl:263-339).  Everything before the
# loop (TrajectoryMetaData, Result construction, frame selection) and after it (finalresults!,
# contributions, ...) is the unmodified reference code.
module CMXB200

using ComplexMixtures
import ComplexMixtures: Trajectory, Options, Result, TrajectoryMetaData, opentraj!, closetraj!, firstframe!,
    nextframe!, getunitcell, convert_unitcell, finalresults!, isautocorrelation

const libcmx = get(ENV, "CMX_B200_LIB", "libcmx_b200.so")

# struct cmx_config (include/cmx_b200.h) -- field order and types must match exactly
struct CmxConfig
    struct_size::Int32; device::Int32
    solute_nmols::Int32; solute_natomspermol::Int32; solvent_nmols::Int32; solvent_natomspermol::Int32
    autocorrelation::Int32; irefatom::Int32; usecutoff::Int32; n_random_samples::Int32
    coordination_number_only::Int32; lcell::Int32; n_groups_solute::Int32; n_groups_solvent::Int32
    path::Int32; ring_slots::Int32; keep_lists::Int32; group_lanes::Int32; n_streams::Int32; batch_frames::Int32
    cutoff::Float64; dbulk::Float64; binstep::Float64; seed::UInt64
    solute_group_offsets::Ptr{Int32}; solute_group_ids::Ptr{Int32}
    solvent_group_offsets::Ptr{Int32}; solvent_group_ids::Ptr{Int32}
    n_devices::Int32; reserved1::Int32; device_ids::Ptr{Int32}       # several GPUs behind ONE handle (src/parallel_setup.jl:7-57)
end

mutable struct CmxCounters
    nbins::Int32; n_groups_solute::Int32; n_groups_solvent::Int32; reserved::Int32
    md_count::Ptr{Float64}; md_count_random::Ptr{Float64}; rdf_count::Ptr{Float64}; rdf_count_random::Ptr{Float64}
    solute_group_count::Ptr{Float64}; solute_group_count_random::Ptr{Float64}
    solvent_group_count::Ptr{Float64}; solvent_group_count_random::Ptr{Float64}
    volume_total::Float64; sum_weights::Float64
end

check(h, rc) = rc == 0 || error("libcmx_b200: " * unsafe_string(ccall((:cmx_last_error, libcmx), Cstring, (Ptr{Cvoid},), h)))

# CSR "position in the selection -> groups" (restates the search in update_group_count!, src/update_counters.jl:27-33)
function group_csr(sel::AtomSelection)
    sel.custom_groups || return (Int32[], Int32[])
    pos = Dict(a => p for (p, a) in enumerate(sel.indices))
    per = [Int32[] for _ in sel.indices]
    for (g, inds) in enumerate(sel.group_atom_indices), a in inds
        push!(per[pos[a]], Int32(g - 1))
    end
    off = Int32[0; cumsum(length.(per))]
    return off, isempty(per) ? Int32[0] : reduce(vcat, per; init=Int32[])
end

"""
    mddf_b200(trajectory, options; frame_weights, coordination_number_only, device=0, devices=Int[])

Drop-in for `ComplexMixtures.mddf(trajectory, options; ...)`: same arguments, same `Result`.
`devices = [0, 1, ..., 7]` puts the whole box behind this ONE call (as the reference's single `mddf` uses every
thread): frames are dealt to the GPUs in order and their counters are summed on the first one inside `cmx_finish`.
"""
function mddf_b200(trajectory::Trajectory, options::Options=Options();
                   frame_weights=Float64[], coordination_number_only=false, low_memory=false, device::Integer=0,
                   devices::Vector{<:Integer}=Int[])
    tmeta = TrajectoryMetaData(trajectory, options)
    R = Result(trajectory, options; trajectory_data=tmeta, frame_weights)
    sol, solv = trajectory.solute, trajectory.solvent
    soff, sids = group_csr(sol); voff, vids = group_csr(solv)
    devs = Int32.(devices)
    GC.@preserve soff sids voff vids devs begin
        cfg = Ref(CmxConfig(sizeof(CmxConfig), device, sol.nmols, sol.natomspermol, solv.nmols, solv.natomspermol,
            R.autocorrelation, tmeta.irefatom, options.usecutoff, options.n_random_samples, coordination_number_only,
            options.lcell, tmeta.n_groups_solute, tmeta.n_groups_solvent, 0, 0, 0, 0, 0, 0,
            options.cutoff, options.dbulk, options.binstep, UInt64(max(options.seed, 0)),
            sol.custom_groups ? pointer(soff) : C_NULL, sol.custom_groups ? pointer(sids) : C_NULL,
            solv.custom_groups ? pointer(voff) : C_NULL, solv.custom_groups ? pointer(vids) : C_NULL,
            length(devs) > 1 ? length(devs) : 0, 0, length(devs) > 1 ? pointer(devs) : C_NULL))
        href = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:cmx_create, libcmx), Int32, (Ref{CmxConfig}, Ref{Ptr{Cvoid}}), cfg, href)
        rc == 0 || error(unsafe_string(ccall((:cmx_last_error, libcmx), Cstring, (Ptr{Cvoid},), C_NULL)))
    end
    h = href[]
    ns, nv = sol.nmols * sol.natomspermol, solv.nmols * solv.natomspermol
    opentraj!(trajectory); firstframe!(trajectory)
    try
        for iframe in 1:R.files[1].lastframe_read          # src/mddf.jl:285-338 without threads/locks
            isfile("stop_complexmixtures") && break        # src/mddf.jl:301-304
            nextframe!(trajectory)
            w = R.files[1].frame_weights[iframe]
            (iframe in options.firstframe:options.stride:R.files[1].lastframe_read && !iszero(w)) || continue
            ps = Ref{Ptr{Float32}}(); pv = Ref{Ptr{Float32}}()
            check(h, ccall((:cmx_acquire_frame_buffer, libcmx), Int32, (Ptr{Cvoid}, Ref{Ptr{Float32}}, Ref{Ptr{Float32}}), h, ps, pv))
            xv = unsafe_wrap(Array, pv[], (3, nv))         # pinned staging slot: fp32, written in place
            @inbounds for i in 1:nv, k in 1:3; xv[k, i] = trajectory.x_solvent[i][k]; end
            if !R.autocorrelation
                xs = unsafe_wrap(Array, ps[], (3, ns))
                @inbounds for i in 1:ns, k in 1:3; xs[k, i] = trajectory.x_solute[i][k]; end
            end
            uc = getunitcell(trajectory)                    # 3x3, columns = lattice vectors -> column-major double[9]
            cell = Float64[uc[i, j] for i in 1:3, j in 1:3]
            check(h, ccall((:cmx_submit_frame, libcmx), Int32, (Ptr{Cvoid}, Int64, Float64, Ptr{Float64}), h, iframe, w, cell))
        end
    finally
        closetraj!(trajectory)
    end
    # sum!(R, r_chunk) (src/mddf.jl:336): the library writes Float64 arrays laid out like Result; the
    # Vector{Vector{Float64}} group arrays are filled row by row from one contiguous buffer
    nb = R.nbins
    gs = zeros(nb, tmeta.n_groups_solute); gsr = zeros(nb, tmeta.n_groups_solute)
    gv = zeros(nb, tmeta.n_groups_solvent); gvr = zeros(nb, tmeta.n_groups_solvent)
    c = CmxCounters(0, 0, 0, 0, pointer(R.md_count), pointer(R.md_count_random), pointer(R.rdf_count), pointer(R.rdf_count_random),
                    pointer(gs), pointer(gsr), pointer(gv), pointer(gvr), 0.0, 0.0)
    GC.@preserve R gs gsr gv gvr check(h, ccall((:cmx_finish, libcmx), Int32, (Ptr{Cvoid}, Ref{CmxCounters}), h, c))
    for g in 1:tmeta.n_groups_solute
        R.solute_group_count[g] .= @view gs[:, g]; R.solute_group_count_random[g] .= @view gsr[:, g]
    end
    for g in 1:tmeta.n_groups_solvent
        R.solvent_group_count[g] .= @view gv[:, g]; R.solvent_group_count_random[g] .= @view gvr[:, g]
    end
    R.volume.total = c.volume_total
    ccall((:cmx_destroy, libcmx), Int32, (Ptr{Cvoid},), h)
    return finalresults!(R, options; coordination_number_only)   # unmodified reference normalisation
end

# ---- SURVEY 8(f1): the library's own DCD feed ------------------------------------------------------
struct CmxDcdInfo
    natoms::Int64; nframes::Int64; first_frame_offset::Int64; frame_bytes::Int64
end

"""
    run_dcd!(h, trajectory::ComplexMixtures.NamdDCD, frames, weights; reader_threads=2)

Replaces the whole `for iframe ...` loop above for DCD files: `cmx_run_dcd` preads the raw frame records
into a pinned ring on reader threads, copies each raw frame to the device once and gathers the selected
atoms there (the host-side gather of `nextframe!`, src/trajectory_formats/NamdDCD.jl:141-169, disappears).
`frames` are 1-based frame numbers as in the reference; the Philox key stays the 1-based frame number.
"""
function run_dcd!(h::Ptr{Cvoid}, trajectory, frames::Vector{Int}, weights::Vector{Float64}; reader_threads::Integer=2)
    dref = Ref{Ptr{Cvoid}}(C_NULL); info = Ref(CmxDcdInfo(0, 0, 0, 0))
    rc = ccall((:cmx_dcd_open, libcmx), Int32, (Cstring, Ref{Ptr{Cvoid}}, Ref{CmxDcdInfo}), trajectory.filename, dref, info)
    rc == 0 || error(unsafe_string(ccall((:cmx_dcd_last_error, libcmx), Cstring, ())))
    sol = Int32.(trajectory.solute.indices); solv = Int32.(trajectory.solvent.indices)
    fr0 = Int64.(frames .- 1)
    try
        GC.@preserve sol solv fr0 weights check(h, ccall((:cmx_run_dcd, libcmx), Int32,
            (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Int32}, Ptr{Int32}, Ptr{Int64}, Ptr{Float64}, Int64, Int32),
            h, dref[], sol, solv, fr0, weights, length(fr0), reader_threads))
    finally
        ccall((:cmx_dcd_close, libcmx), Int32, (Ptr{Cvoid},), dref[])
    end
end

# XTC twin: `cmx_xtc_open` / `cmx_run_xtc` / `cmx_xtc_close` take the same arguments (a Chemfiles trajectory's
# `trajectory.filename` ending in .xtc); errors of cmx_xtc_open come through `cmx_dcd_last_error` as well.
function run_xtc!(h::Ptr{Cvoid}, trajectory, frames::Vector{Int}, weights::Vector{Float64}; reader_threads::Integer=4)
    xref = Ref{Ptr{Cvoid}}(C_NULL); info = Ref((Int64(0), Int64(0)))
    rc = ccall((:cmx_xtc_open, libcmx), Int32, (Cstring, Ref{Ptr{Cvoid}}, Ptr{Cvoid}), trajectory.filename, xref, info)
    rc == 0 || error(unsafe_string(ccall((:cmx_dcd_last_error, libcmx), Cstring, ())))
    sol = Int32.(trajectory.solute.indices); solv = Int32.(trajectory.solvent.indices)
    fr0 = Int64.(frames .- 1)
    try
        GC.@preserve sol solv fr0 weights check(h, ccall((:cmx_run_xtc, libcmx), Int32,
            (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Int32}, Ptr{Int32}, Ptr{Int64}, Ptr{Float64}, Int64, Int32),
            h, xref[], sol, solv, fr0, weights, length(fr0), reader_threads))
    finally
        ccall((:cmx_xtc_close, libcmx), Int32, (Ptr{Cvoid},), xref[])
    end
end

# ---- SURVEY 8(f2): contributions / ResidueContributions count stage on the device ----------------------
"""
    reduce_groups(h, which, groups, nbins) -> Matrix{Float64}(nbins, length(groups))

`which`: 0 solute_group_count, 1 solute_group_count_random, 2 solvent_group_count, 3 solvent_group_count_random;
`groups[g]` = 1-based rows (atoms of the selection, or custom groups) summed into output column g -- the loop over
`group_count[igroup]` of `contributions` (src/tools/contributions.jl:206-236) without reading the per-atom array back.
"""
function reduce_groups(h::Ptr{Cvoid}, which::Integer, groups::Vector{Vector{Int}}, nbins::Integer)
    off = Int32[0; cumsum(length.(groups))]
    rows = Int32.(reduce(vcat, groups; init=Int[]) .- 1)
    out = zeros(nbins, length(groups))
    GC.@preserve off rows out check(h, ccall((:cmx_reduce_groups, libcmx), Int32,
        (Ptr{Cvoid}, Int32, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}), h, which, length(groups), off, rows, out))
    return out
end

"""
    final_results(h, nbins; sum_weights = 0.0, volume_sum = 0.0) -> NamedTuple

`finalresults!` (src/results.jl:311-469) on the device: the `nbins` vectors of `Result` and the `Volume` / `Density`
scalars, from the accumulators in HBM.  Field order = `cmx_final` (include/cmx_b200.h).
"""
mutable struct CmxFinal
    nbins::Int32; reserved::Int32
    d::Ptr{Float64}; md_count::Ptr{Float64}; md_count_random::Ptr{Float64}; coordination_number::Ptr{Float64}
    coordination_number_random::Ptr{Float64}; mddf::Ptr{Float64}; kb::Ptr{Float64}; rdf_count::Ptr{Float64}
    rdf_count_random::Ptr{Float64}; sum_rdf_count::Ptr{Float64}; sum_rdf_count_random::Ptr{Float64}; rdf::Ptr{Float64}
    kb_rdf::Ptr{Float64}; volume_shell::Ptr{Float64}
    volume_total::Float64; volume_bulk::Float64; volume_domain::Float64
    density_solute::Float64; density_solvent::Float64; density_solvent_bulk::Float64; density_fix::Float64; sum_weights::Float64
end
const FINAL_VECTORS = (:d, :md_count, :md_count_random, :coordination_number, :coordination_number_random, :mddf, :kb, :rdf_count,
                       :rdf_count_random, :sum_rdf_count, :sum_rdf_count_random, :rdf, :kb_rdf, :volume_shell)
function final_results(h::Ptr{Cvoid}, nbins::Integer; sum_weights::Float64 = 0.0, volume_sum::Float64 = 0.0)
    vecs = Dict(k => zeros(nbins) for k in FINAL_VECTORS)
    f = CmxFinal(nbins, 0, (pointer(vecs[k]) for k in FINAL_VECTORS)..., 0, 0, 0, 0, 0, 0, 0, 0)
    GC.@preserve vecs check(h, ccall((:cmx_final_results, libcmx), Int32, (Ptr{Cvoid}, Float64, Float64, Ref{CmxFinal}), h, sum_weights, volume_sum, f))
    return (; vecs..., volume_total = f.volume_total, volume_bulk = f.volume_bulk, volume_domain = f.volume_domain,
            density_solute = f.density_solute, density_solvent = f.density_solvent, density_solvent_bulk = f.density_solvent_bulk)
end

"""
    contributions(h, side, type, groups, nbins) -> Matrix{Float64}(nbins, length(groups))

`contributions(R, SoluteGroup|SolventGroup; type)` (src/tools/contributions.jl:70-248) for many groups at once on the
device; `side` = :solute | :solvent, `type` = :mddf | :coordination_number | :md_count | :kbi, `groups[g]` = 1-based rows.
"""
function contributions(h::Ptr{Cvoid}, side::Symbol, type::Symbol, groups::Vector{Vector{Int}}, nbins::Integer;
                       sum_weights::Float64 = 0.0, volume_sum::Float64 = 0.0)
    off = Int32[0; cumsum(length.(groups))]
    rows = Int32.(reduce(vcat, groups; init=Int[]) .- 1)
    out = zeros(nbins, length(groups))
    t = findfirst(==(type), (:mddf, :coordination_number, :md_count, :kbi)) - 1
    GC.@preserve off rows out check(h, ccall((:cmx_contributions, libcmx), Int32,
        (Ptr{Cvoid}, Int32, Int32, Float64, Float64, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}),
        h, side == :solute ? 0 : 1, t, sum_weights, volume_sum, length(groups), off, rows, out))
    return out
end

end # module
