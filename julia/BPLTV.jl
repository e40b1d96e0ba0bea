# BPLTV.jl — thin `ccall` wrappers over libbpltv.so (include/bpltv.h).
#
# Drop-in for the learning-function path of dvillacis/BPLDenoising: the methods below have the
# signatures of src/TVLearningFunctionVec.jl (`tv_op_learning_function` :14-27, `denoise` :45-70)
# and src/BPLDenoising.jl (`TVDenoise` :41-82), so `bilevel_learn` (src/TRBox.jl:192-273) and the
# experiment functions (src/BPLDenoising.jl:325-376) call them unchanged:
#
#     include("julia/BPLTV.jl"); using .BPLTV
#     x, u, st = bilevel_learn((b, b_noisy), BPLTV.tv_op_learning_function; xinit=0.1, iterate=iterate, params=params)
#
# NOTE: Julia is not installed in the build environment; this file is exercised by inspection and
# by the equivalent ctypes binding (bpldenoising_b200/_lib.py), which calls the same symbols.
module BPLTV

export tv_op_learning_function, sumregs_learning_function, sumregs_denoise, denoise, TVDenoise, generate_cost, bpltv_context, set_devices!,
       comm_unique_id, comm_init!, comm_destroy!, shard_range

const lib = get(ENV, "BPLTV_LIB", joinpath(@__DIR__, "..", "bpldenoising_b200", "libbpltv.so"))

# mirrors of the C structs (field order and types of include/bpltv.h)
struct PdpsOpts
    tau0::Cdouble; sigma0::Cdouble; rho::Cdouble; opnorm::Cdouble
    accel::Cint; maxiter::Cint; init_mode::Cint; arith::Cint; kernel::Cint; tblock::Cint
    reserved::NTuple{4,Cint}
end
struct EvalOpts
    pdps::PdpsOpts
    delta_t::Cdouble; gamma::Cdouble; act_tol::Cdouble; eps_act::Cdouble; solver_tol::Cdouble
    solver_maxit::Cint; solver::Cint; force_branch::Cint
    reserved0::Cint; gamma_patch::Cdouble
    reserved::NTuple{2,Cint}
end

lasterr() = unsafe_string(ccall((:bpltv_last_error, lib), Cstring, ()))
# the BPLTV_* developer switches are snapshotted when the first context is created; re-read them after changing ENV
reload_env() = ccall((:bpltv_reload_env, lib), Cvoid, ())
# arithmetic self-test of the strict kernels' projection scale (include/bpltv.h): (pairs taken, mismatches, a bits, α bits)
function selftest(ctx::Ptr{Cvoid}, mode::Integer, count::Integer; seed::Integer = 1)
    res = zeros(Culonglong, 4)
    rc = ccall((:bpltv_selftest, lib), Cint, (Ptr{Cvoid}, Cint, Cint, Culonglong, Culonglong, Ptr{Culonglong}),
               ctx, 0, mode, count, seed, res)
    rc == 0 || error(unsafe_string(ccall((:bpltv_last_error, lib), Cstring, ())))
    return res
end
check(rc) = rc == 0 ? nothing : (rc == -1 ? throw(ArgumentError(lasterr())) : error("libbpltv ($rc): " * lasterr()))

function default_pdps()
    r = Ref{PdpsOpts}()
    ccall((:bpltv_default_pdps_opts, lib), Cvoid, (Ref{PdpsOpts},), r)
    r[]
end
function default_eval()
    r = Ref{EvalOpts}()
    ccall((:bpltv_default_eval_opts, lib), Cvoid, (Ref{EvalOpts},), r)
    r[]
end
# `denoising_default_params ⬿ kwargs` (TVLearningFunctionVec.jl:50): override by keyword
function with(o::T; kw...) where T
    vals = [haskey(kw, f) ? convert(fieldtype(T, f), kw[f]) : getfield(o, f) for f in fieldnames(T)]
    T(vals...)
end
const kwalias = Dict(:τ₀ => :tau0, :σ₀ => :sigma0, :ρ => :rho)
pdps_from(kwargs) = with(default_pdps(); Dict(get(kwalias, k, k) => v for (k, v) in kwargs
                         if !(k in (:verbose_iter, :save_results, :save_iterations, :α, :op)))...)

# one context per process; devices default to GPU 0 (set_devices!(0:7) shards images over a box)
const ctx = Ref{Ptr{Cvoid}}(C_NULL)
const devices = Ref{Vector{Cint}}(Cint[0])
const resident = Ref{Any}(nothing)
function set_devices!(ids)
    ctx[] != C_NULL && (ccall((:bpltv_destroy, lib), Cint, (Ptr{Cvoid},), ctx[]); ctx[] = C_NULL)
    devices[] = collect(Cint, ids); resident[] = nothing
end
function bpltv_context()
    if ctx[] == C_NULL
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:bpltv_create, lib), Cint, (Ptr{Cint}, Cint, Cint, Ref{Ptr{Cvoid}}),
                    devices[], length(devices[]), 64, h))
        ctx[] = h[]
    end
    ctx[]
end

# ---- one process per GPU (MPI.jl, Distributed.jl): the job's single collective lives in the library (include/bpltv.h) ----
# rank 0: id = comm_unique_id(); ship the 128 bytes to every rank by your own transport; all ranks: comm_init!(n, r, id).
# Afterwards tv_op_learning_function / sumregs_learning_function called with THIS rank's block of the data
# (shard_range) return the loss and gradient of the whole job on every rank (one ncclAllReduce per evaluation).
function comm_unique_id()
    id = zeros(UInt8, 128)
    check(ccall((:bpltv_comm_unique_id, lib), Cint, (Ptr{UInt8},), id))
    id
end
function comm_init!(nranks::Integer, rank::Integer, id::Vector{UInt8})
    length(devices[]) == 1 || throw(ArgumentError("a communicator joins single-device contexts: set_devices!([gpu]) first"))
    check(ccall((:bpltv_comm_init, lib), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{UInt8}), bpltv_context(), nranks, rank, id))
end
comm_destroy!() = ctx[] == C_NULL ? nothing : check(ccall((:bpltv_comm_destroy, lib), Cint, (Ptr{Cvoid},), ctx[]))
"first image (1-based) and count of `rank`'s contiguous block of O images: blocks of ceil(O/nranks), as inside the library"
function shard_range(O::Integer, nranks::Integer, rank::Integer)
    per = cld(O, nranks); b = min(O, rank * per)
    (b + 1, min(O, b + per) - b)
end

lam(x::Real) = (Float64[x;;], 1, 1)
lam(x::AbstractMatrix) = (Matrix{Float64}(x), size(x, 1), size(x, 2))

"denoise(data, x, op; kwargs...) — src/TVLearningFunctionVec.jl:45-70"
function denoise(data::AbstractArray{<:Real,3}, x, op=nothing; kwargs...)
    f = Array{Float64,3}(data); M, N, O = size(f)
    l, lm, ln = lam(x); o = Ref(pdps_from(kwargs)); u = similar(f)
    check(ccall((:bpltv_denoise, lib), Cint,
                (Ptr{Cvoid}, Ptr{Cdouble}, Cint, Cint, Cint, Ptr{Cdouble}, Cint, Cint, Ref{PdpsOpts}, Ptr{Cdouble}),
                bpltv_context(), f, M, N, O, l, lm, ln, o, u))
    u
end

"TVDenoise(data, parameter) — src/BPLDenoising.jl:41-82 (maxiter = 10000)"
TVDenoise(data, parameter; visualize=false) = denoise(data, parameter; maxiter=10000)

"tv_op_learning_function(x, data, Δ; Δt=1e-6, kwargs...) → (u, cost, grad) — src/TVLearningFunctionVec.jl:14-27"
function tv_op_learning_function(x, data, Δ; Δt=1e-6, kwargs...)
    ū, f = data[1], data[2]
    M, N, O = size(f)
    h = bpltv_context()
    if resident[] !== data        # the dataset is constant over a learn run (src/TRBox.jl:210,227)
        check(ccall((:bpltv_set_dataset, lib), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Cint, Cint),
                    h, Array{Float64,3}(ū), Array{Float64,3}(f), M, N, O))
        resident[] = data
    end
    l, lm, ln = lam(x)
    eo = Ref(with(default_eval(); pdps=pdps_from(kwargs), delta_t=Δt))
    u = Array{Float64,3}(undef, M, N, O); cost = Ref{Cdouble}(0); grad = zeros(lm, ln)
    check(ccall((:bpltv_learn_eval, lib), Cint,
                (Ptr{Cvoid}, Ptr{Cdouble}, Cint, Cint, Cdouble, Ref{EvalOpts}, Ptr{Cdouble}, Ref{Cdouble}, Ptr{Cdouble}),
                h, l, lm, ln, Δ, eo, u, cost, grad))
    return u, cost[], (x isa Real ? grad[1] : grad)   # grad has the shape of x (src/TRBox.jl:63,237)
end

# ---- sum-of-regularisers interface (src/SumRegsLearningFunction.jl) -------------------------------
lam3(x::AbstractVector) = (Vector{Float64}(x), 1, 1)
lam3(x::AbstractArray{<:Real,3}) = (Array{Float64,3}(x), size(x, 1), size(x, 2))
function default_sumregs_eval()
    r = Ref{EvalOpts}()
    ccall((:bpltv_default_sumregs_eval_opts, lib), Cvoid, (Ref{EvalOpts},), r)
    r[]
end

"sumregs_denoise(data, x, op₁, op₂, op₃[, pOp]) — src/SumRegsLearningFunction.jl:38-85"
function sumregs_denoise(data::AbstractArray{<:Real,3}, x, ops...; kwargs...)
    f = Array{Float64,3}(data); M, N, O = size(f)
    l, lm, ln = lam3(x); o = Ref(with(default_sumregs_eval().pdps; kwargs...)); u = similar(f)
    check(ccall((:bpltv_sumregs_denoise, lib), Cint,
                (Ptr{Cvoid}, Ptr{Cdouble}, Cint, Cint, Cint, Ptr{Cdouble}, Cint, Cint, Ref{PdpsOpts}, Ptr{Cdouble}),
                bpltv_context(), f, M, N, O, l, lm, ln, o, u))
    u
end

"sumregs_learning_function(x, data, Δ; Δt=1e-3) → (u, cost, grad) — src/SumRegsLearningFunction.jl:8-36"
function sumregs_learning_function(x, data, Δ; Δt=1e-3)
    ū, f = data[1], data[2]
    M, N, O = size(f)
    h = bpltv_context()
    if resident[] !== data
        check(ccall((:bpltv_set_dataset, lib), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Cint, Cint),
                    h, Array{Float64,3}(ū), Array{Float64,3}(f), M, N, O))
        resident[] = data
    end
    l, lm, ln = lam3(x)
    eo = Ref(with(default_sumregs_eval(); delta_t=Δt))
    u = Array{Float64,3}(undef, M, N, O); cost = Ref{Cdouble}(0); grad = zeros(size(x))
    check(ccall((:bpltv_sumregs_learn_eval, lib), Cint,
                (Ptr{Cvoid}, Ptr{Cdouble}, Cint, Cint, Cdouble, Ref{EvalOpts}, Ptr{Cdouble}, Ref{Cdouble}, Ptr{Cdouble}),
                h, l, lm, ln, Δ, eo, u, cost, grad))
    return u, cost[], grad                       # grad has the shape of x (3-vector or m×n×3)
end

"""
generate_cost(true_, data, parameter_range) — the loop of src/BPLDenoising.jl:92-111 / :136-158
(`u = TVDenoise(data, parameter_range[i]); costs[i] = L2CostFunction(u, true_)`) as ONE batched call:
all parameter sets × images are independent solves.  `parameter_range`: reals or equally sized matrices.
"""
function generate_cost(true_::AbstractArray{<:Real,3}, data::AbstractArray{<:Real,3}, parameter_range; kwargs...)
    M, N, O = size(data); h = bpltv_context()
    check(ccall((:bpltv_set_dataset, lib), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Cint, Cint),
                h, Array{Float64,3}(true_), Array{Float64,3}(data), M, N, O))
    resident[] = nothing
    ps = [lam(p) for p in parameter_range]; lm, ln = ps[1][2], ps[1][3]
    lams = reduce(vcat, [vec(p[1]) for p in ps]); L = length(ps)
    o = Ref(pdps_from((; maxiter=10000, kwargs...))); costs = zeros(L)
    check(ccall((:bpltv_sweep, lib), Cint,
                (Ptr{Cvoid}, Ptr{Cdouble}, Cint, Cint, Cint, Ref{PdpsOpts}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                h, lams, L, lm, ln, o, costs, C_NULL, C_NULL))
    costs
end

end # module
