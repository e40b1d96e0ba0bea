# Dump golden (u, cost, grad) vectors of the REAL reference for the five datasets.
# Needs the full Julia stack of /root/reference/README.md (AlgTools, ImageTools,
# VariationalImaging, BPLDenoising) — it cannot run in the build container (no Julia).
# Output: tests/golden/julia/<dataset>_<case>.{u,meta}.f64  (raw little-endian Float64)
#
#   julia --project=/path/to/BPLDenoising tools/dump_reference_vectors.jl tests/golden/julia
using BPLDenoising
using BPLDenoising.Datasets: testdataset
using ColorTypes: Gray

outdir = length(ARGS) > 0 ? ARGS[1] : "tests/golden/julia"
mkpath(outdir)

function dump(name, x, Δ, tag)
    b, b_noisy = testdataset(name)
    b = Float64.(Gray{Float64}.(b)); b_noisy = Float64.(Gray{Float64}.(b_noisy))
    u, cost, grad = tv_op_learning_function(x, (b, b_noisy), Δ)
    open(joinpath(outdir, "$(name)_$(tag).u.f64"), "w") do io; write(io, vec(u)); end
    open(joinpath(outdir, "$(name)_$(tag).meta.f64"), "w") do io
        write(io, Float64[size(u)..., cost, length(grad), vec(collect(grad))...])
    end
    @info "dumped" name tag cost grad
end

for name in ("cameraman_128_5", "cameraman_128_10", "faces_train_128_10", "faces_val_128_10", "circle_128_10")
    dump(name, 0.1, 0.1, "scalar_nonreg")       # Δ > Δt  → gradient      (TVLearningFunctionVec.jl:21-22)
    dump(name, 0.1, 1e-7, "scalar_reg")         # Δ ≤ Δt  → gradient_reg  (:23-24)
    dump(name, 1e-4 * ones(2, 2), 1e-4, "patch_nonreg")
    dump(name, 1e-4 * ones(2, 2), 1e-7, "patch_reg")
end

# sum-of-regularisers interface (src/SumRegsLearningFunction.jl:8-36): scalar parameter, both branches
function dump_sumregs(name, x, Δ, tag)
    b, b_noisy = testdataset(name)
    b = Float64.(Gray{Float64}.(b)); b_noisy = Float64.(Gray{Float64}.(b_noisy))
    u, cost, grad = sumregs_learning_function(x, (b, b_noisy), Δ)
    open(joinpath(outdir, "$(name)_$(tag).u.f64"), "w") do io; write(io, vec(u)); end
    open(joinpath(outdir, "$(name)_$(tag).meta.f64"), "w") do io
        write(io, Float64[size(u)..., cost, length(grad), vec(collect(grad))...])
    end
    @info "dumped" name tag cost grad
end
for name in ("cameraman_128_5", "circle_128_10")
    dump_sumregs(name, [0.001; 0.001; 0.001], 0.01, "sumregs_nonreg")   # Δ > Δt = 1e-3 → sumregs_gradient
    dump_sumregs(name, [0.001; 0.001; 0.001], 1e-4, "sumregs_reg")      # Δ ≤ Δt        → sumregs_gradient_reg
end
