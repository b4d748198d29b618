# planify.jl — twin of `_exafy` / `_map_variable` (src/transform.jl:290-389) that emits postfix tapes and
# affine index expressions for libiexa_b200.so instead of ExaModels node objects.
#
# STATUS: written, NOT executed (no Julia offline).  It deliberately does not depend on ExaModels'
# internal node types.  Include after B200ExaModels.jl inside the InfiniteExaModels module; the
# emitters of transform.jl then call `tape_con!(plan, expr, itr_id, int_names, fp_names, data; lcon, ucon)`
# where they call `ExaModels.add_con(core, em_expr, itr; lcon, ucon)` today (:458,:559,:597), and
# `tape_obj!` where they call `ExaModels.add_obj` (:614,:700,:741).

using .B200ExaModels: IexaNode, IexaIndex, OP, Plan, add_con!, add_obj!

"affine integer index `base + Σ coef*int_field` (what Variable[data_src[alias]...] reduces to)"
struct AffIdx
    base::Int64
    terms::Dict{Symbol,Int64}
end
AffIdx(c::Integer) = AffIdx(Int64(c), Dict{Symbol,Int64}())
AffIdx(s::Symbol) = AffIdx(0, Dict(s => Int64(1)))
Base.:+(a::AffIdx, b::AffIdx) = AffIdx(a.base + b.base, mergewith(+, a.terms, b.terms))
Base.:+(a::AffIdx, c::Integer) = AffIdx(a.base + c, copy(a.terms))
Base.:-(a::AffIdx, c::Integer) = a + (-c)
Base.:*(a::AffIdx, c::Integer) = AffIdx(a.base * c, Dict(k => v * c for (k, v) in a.terms))

"a variable / parameter block: ExaModels' `offset` and `size` (column-major)"
struct Block
    offset::Int64            # 0-based
    size::Tuple{Vararg{Int}}
end
function index(b::Block, subs)             # subs: Int or Symbol (iterator int field) or AffIdx per dimension
    ix = AffIdx(b.offset + 1); stride = 1
    for (s, n) in zip(subs, b.size)
        a = s isa AffIdx ? s : AffIdx(s)
        ix = ix + (a - 1) * stride
        stride *= n
    end
    return ix
end

mutable struct TapeBuilder
    nodes::Vector{IexaNode}
    idx::Vector{IexaIndex}
    icol::Dict{Symbol,Int32}     # iterator int field  -> column
    fcol::Dict{Symbol,Int32}     # iterator fp field   -> column
end
TapeBuilder(int_names, fp_names) = TapeBuilder(IexaNode[], IexaIndex[],
    Dict(n => Int32(i - 1) for (i, n) in enumerate(int_names)), Dict(n => Int32(i - 1) for (i, n) in enumerate(fp_names)))

push_node!(t::TapeBuilder, op, a = 0, b = 0, c = 0.0) = (push!(t.nodes, IexaNode(op, a, b, 0, c)); Int32(length(t.nodes) - 1))
function push_index!(t::TapeBuilder, ix::AffIdx)
    ks = collect(keys(ix.terms)); length(ks) <= 4 || error("index expression with more than 4 integer fields")
    col = ntuple(i -> i <= length(ks) ? t.icol[ks[i]] : Int32(0), 4)
    coef = ntuple(i -> i <= length(ks) ? ix.terms[ks[i]] : Int64(0), 4)
    push!(t.idx, IexaIndex(ix.base, length(ks), col, 0, coef)); return Int32(length(t.idx) - 1)
end
konst!(t, c::Real) = push_node!(t, OP[:CONST], 0, 0, Float64(c))
field!(t, s::Symbol) = push_node!(t, OP[:FIELD], t.fcol[s])
var!(t, ix::AffIdx) = push_node!(t, OP[:VAR], push_index!(t, ix))
par!(t, ix::AffIdx) = push_node!(t, OP[:PAR], push_index!(t, ix))
bin!(t, op::Symbol, a, b) = push_node!(t, OP[op], a, b)
un!(t, op::Symbol, a) = push_node!(t, OP[op], a)

# ---- _map_variable twin (transform.jl:290-334): returns a node id ------------------------------------------
function tape_variable!(t::TapeBuilder, vref, data)
    IT = vref.index_type
    if IT == InfiniteOpt.FiniteVariableIndex || IT == InfiniteOpt.PointVariableIndex
        return var!(t, AffIdx(data.finvar_index[vref]))                       # Var(const i)            :290-301
    elseif IT <: Union{InfiniteOpt.InfiniteVariableIndex,InfiniteOpt.DerivativeIndex}
        groups = InfiniteOpt.parameter_group_int_indices(vref)
        return var!(t, index(data.infvar_block[vref], [data.group_alias[g] for g in groups]))   # :302-311
    elseif IT == InfiniteOpt.SemiInfiniteVariableIndex
        blk, inds = data.semivar_info[vref]                                  # const ints and aliases   :312-319
        return var!(t, index(blk, inds))
    elseif IT <: InfiniteOpt.InfiniteParameterIndex
        return field!(t, data.param_alias[vref])                             # data_src[alias]          :320-322
    elseif IT == InfiniteOpt.FiniteParameterIndex
        return par!(t, AffIdx(data.param_block[vref].offset + 1))            # Parameter[1]             :323-325
    elseif IT == InfiniteOpt.ParameterFunctionIndex
        groups = InfiniteOpt.parameter_group_int_indices(vref)
        return par!(t, index(data.param_block[vref], [data.group_alias[g] for g in groups]))   # :326-330
    end
    error("Unable to add `$vref` to an ExaModel, it's index type `$IT` is not yet supported by InfiniteExaModels.")
end

# ---- _exafy twin (transform.jl:337-389) ----------------------------------------------------------------
tape_expr!(t::TapeBuilder, c::Number, data) = konst!(t, c)
tape_expr!(t::TapeBuilder, v::InfiniteOpt.GeneralVariableRef, data) = tape_variable!(t, v, data)
function tape_expr!(t::TapeBuilder, aff::JuMP.GenericAffExpr, data)
    c = JuMP.constant(aff)
    isempty(aff.terms) && return konst!(t, c)
    acc = nothing
    for (coef, v) in JuMP.linear_terms(aff)                                    # sum(c*v), coefficient on the left
        n = tape_variable!(t, v, data)
        isone(coef) || (n = bin!(t, :*, konst!(t, coef), n))                   # isone(c) elision :352
        acc = acc === nothing ? n : bin!(t, :+, acc, n)
    end
    return iszero(c) ? acc : bin!(t, :+, acc, konst!(t, c))                   # + constant last :355
end
function tape_expr!(t::TapeBuilder, quad::JuMP.GenericQuadExpr, data)
    isempty(quad.terms) && return tape_expr!(t, quad.aff, data)
    acc = nothing
    for (coef, v1, v2) in JuMP.quad_terms(quad)
        if v1 == v2
            n = un!(t, :abs2, tape_variable!(t, v1, data))                     # abs2(v) :370
            isone(coef) || (n = bin!(t, :*, konst!(t, coef), n))
        else
            a = tape_variable!(t, v1, data); b = tape_variable!(t, v2, data)
            n = isone(coef) ? bin!(t, :*, a, b) : bin!(t, :*, bin!(t, :*, konst!(t, coef), a), b)   # c*v1*v2 :374
        end
        acc = acc === nothing ? n : bin!(t, :+, acc, n)
    end
    return iszero(quad.aff) ? acc : bin!(t, :+, acc, tape_expr!(t, quad.aff, data))   # ex + aff :378
end
function tape_expr!(t::TapeBuilder, nl::JuMP.GenericNonlinearExpr, data)
    haskey(OP, nl.head) || error("`InfiniteExaModel`s does not support the nonlinear operator `$(nl.head)`. " *
                                 "If you need support for this operator, please open an issue.")   # operators.jl:50-53
    args = [tape_expr!(t, a, data) for a in nl.args]
    if length(args) == 1
        nl.head == :- && return un!(t, :neg, args[1])
        nl.head == :+ && return un!(t, :pos, args[1])
        return un!(t, nl.head, args[1])
    end
    acc = args[1]
    for a in args[2:end]                                                       # Base's left fold of n-ary + and *
        acc = bin!(t, nl.head, acc, a)
    end
    return acc
end

"add_con twin: `itr_id`, `int_names`, `fp_names` come from `B200ExaModels.add_iterator!`"
function tape_con!(plan::Plan, expr, itr_id, int_names, fp_names, data; lcon = 0.0, ucon = 0.0)
    t = TapeBuilder(int_names, fp_names)
    tape_expr!(t, expr, data)
    return add_con!(plan, t.nodes, t.idx, Int32(itr_id); lcon = lcon, ucon = ucon)
end
function tape_obj!(plan::Plan, expr, itr_id, int_names, fp_names, data)
    t = TapeBuilder(int_names, fp_names)
    tape_expr!(t, expr, data)
    return add_obj!(plan, t.nodes, t.idx, Int32(itr_id))
end
