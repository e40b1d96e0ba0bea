# cost_curves_to_jld2.jl — turns the cost curves the Python mirror wrote (bpldenoising_b200/results.py: save_cost_curve,
# `<name>_cost.json` + raw Float64 files) into the `.jld2` files generate_cost_plot / generate_2d_cost_plot of the
# reference `@load` (/root/reference/src/BPLDenoising.jl:113-126, :160-174), with the reference's variable names.
#   julia julia/cost_curves_to_jld2.jl output/cameraman_128_5/cameraman_128_5_cost.json
# (Needs Julia with JLD2 and JSON; neither exists in the build environment of this repository, so this script is untested there.)
using JLD2, JSON

function convert_index(index_path::AbstractString)
    idx = JSON.parsefile(index_path)
    dir = dirname(index_path)
    vars = Dict{String,Any}()
    for v in idx["variables"]
        shape = Tuple(Int.(v["shape"]))
        a = Array{Float64}(undef, shape...)
        read!(joinpath(dir, v["file"]), a)          # column-major little-endian Float64
        vars[v["name"]] = a
    end
    out = joinpath(dir, idx["jld2"])
    if haskey(vars, "parameter_range")
        parameter_range, costs = vars["parameter_range"], vars["costs"]
        @save out parameter_range costs                                   # BPLDenoising.jl:110
    else
        parameter_range_1, parameter_range_2, costs = vars["parameter_range_1"], vars["parameter_range_2"], vars["costs"]
        @save out parameter_range_1 parameter_range_2 costs               # BPLDenoising.jl:157
    end
    out
end

for p in ARGS
    println(convert_index(p))
end
