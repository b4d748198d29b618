# dump_golden.jl — run on ANY machine with Julia + ExaModels 0.11.2 + InfiniteExaModels to turn
# "parity unpinned" into pinned goldens (SURVEY.md §8(c), §8(f) rank 2).  NOT executed here.
#
#   julia --project=/path/to/InfiniteExaModels.jl julia/dump_golden.jl tests/golden
#
# For each model it writes  <name>.golden  (little-endian):
#   int64 nvar, ncon, nnzj, nnzh ; float64 x[nvar], y[ncon], obj_weight ;
#   float64 obj ; float64 grad[nvar], cons[ncon] ;
#   int64 jrows[nnzj], jcols[nnzj] ; float64 jvals[nnzj] ;
#   int64 hrows[nnzh], hcols[nnzh] ; float64 hvals[nnzh]
# tests/test_golden_dumps.py loads every *.golden it finds, rebuilds the same model through the Python
# front end, IDENTIFIES the slot-order policy (iexa_set_option IEXA_OPT_SLOT_ORDER: the order in which ExaModels' symbolic
# passes meet the leaves is data in the engine and in the oracle, not code) under which the plan compiler reproduces the dump's COO
# structure bit for bit, and then compares values to 1e-12 relative / 1e-14 absolute.  If no known policy matches, the assertion
# message names the two places (gen.hpp: GenCompiler::kids / jr / hr, oracle.c: kids) where a further order is added.
using InfiniteExaModels, InfiniteOpt, ExaModels, NLPModels, Random

function dump(name, im::InfiniteModel, dir)
    em = ExaModels.ExaModel(im)
    nvar, ncon = NLPModels.get_nvar(em), NLPModels.get_ncon(em)
    nnzj, nnzh = NLPModels.get_nnzj(em), NLPModels.get_nnzh(em)
    rng = MersenneTwister(0)
    x = NLPModels.get_x0(em) .+ 0.1 .* (2 .* rand(rng, nvar) .- 1)
    x = clamp.(x, NLPModels.get_lvar(em), NLPModels.get_uvar(em))
    y = 2 .* rand(rng, ncon) .- 1
    σ = 0.7
    jr, jc = NLPModels.jac_structure(em); hr, hc = NLPModels.hess_structure(em)
    open(joinpath(dir, name * ".golden"), "w") do io
        write(io, Int64[nvar, ncon, nnzj, nnzh]); write(io, x); write(io, y); write(io, σ)
        write(io, NLPModels.obj(em, x)); write(io, NLPModels.grad(em, x)); write(io, NLPModels.cons(em, x))
        write(io, Int64.(jr)); write(io, Int64.(jc)); write(io, NLPModels.jac_coord(em, x))
        write(io, Int64.(hr)); write(io, Int64.(hc)); write(io, NLPModels.hess_coord(em, x, y; obj_weight = σ))
    end
    @info "wrote $name" nvar ncon nnzj nnzh
end

function ode_5x5()   # test/madnlp.jl:4-11
    m = InfiniteModel()
    @infinite_parameter(m, t in [0, 1], num_supports = 5)
    @infinite_parameter(m, x in [-1, 1], num_supports = 5)
    @variable(m, y >= 0, Infinite(t, x))
    @variable(m, z, start = 10)
    @objective(m, Min, ∫(∫(y^2, t) + 2z, x))
    @constraint(m, ∂(y, t) == sin(y) + z + 1.2)
    @constraint(m, y + z <= 42 + t)
    return m
end

dir = length(ARGS) >= 1 ? ARGS[1] : "tests/golden"
mkpath(dir)
dump("ode_5x5", ode_5x5(), dir)
include(joinpath(@__DIR__, "..", "..", "reference", "ESCAPE34", "quadrotor.jl"))   # adjust to the checkout
dump("quadrotor_oc_40", quad(num_supports = 40), dir)
include(joinpath(@__DIR__, "..", "..", "reference", "ESCAPE34", "pandemic.jl"))
pm = pandemic(num_supports = 50, num_scenarios = 4)
dump("pandemic_50x4", pm, dir)
# Julia's RNG stream cannot be reproduced in Python: hand the drawn scenario supports to the test next to the dump
open(joinpath(dir, "pandemic_50x4.xi"), "w") do io
    write(io, Float64.(vec(supports(pm[:ξ]))))
end
