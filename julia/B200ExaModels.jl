# B200ExaModels.jl — Julia binding of libiexa_b200.so (include/iexa.h).
#
# STATUS: written against the C ABI, NOT executed (there is no Julia in the build container or on the
# GPU boxes).  It is the reference-side binding a maintainer would add next to
# src/InfiniteExaModels.jl:6-8 so that `ExaTranscriptionBackend` hands MadNLP / Ipopt a
# `B200ExaModel <: NLPModels.AbstractNLPModel` instead of an `ExaModels.ExaModel`
# (src/infiniteopt_backend.jl:155-156).  The Python package infiniteexamodels.jl_b200 binds the same
# symbols with ctypes and is what the tests exercise.
module B200ExaModels

import NLPModels
using CUDA: CuArray, CuVector, CuPtr, stream

const LIB = get(ENV, "IEXA_B200_LIB", "libiexa_b200.so")

const MEM_HOST = Int32(0)
const MEM_DEVICE = Int32(1)
const MEM_HOST_SAME_X = Int32(2)   # host buffers, x unchanged since the previous host call (Ipopt's new_x == false)

# ---- mirrors of the C structs ------------------------------------------------------------------
struct IexaNode            # iexa_node (24 bytes)
    op::Int32
    a::Int32
    b::Int32
    pad::Int32
    c::Float64
end

struct IexaIndex           # iexa_index (64 bytes)
    base::Int64
    nterms::Int32
    col::NTuple{4,Int32}
    pad::Int32
    coef::NTuple{4,Int64}
end

struct IexaMeta
    nvar::Int64; ncon::Int64; npar::Int64; nobj_gen::Int64; ncon_gen::Int64
    nnzj::Int64; nnzh::Int64; nnzg::Int64
    loc_ncon::Int64; loc_nnzj::Int64; loc_nnzh::Int64
    minimize::Int32; rank::Int32; world::Int32; device::Int32
    n_kernels_specialised::Int32; pad::Int32
end

last_error() = unsafe_string(ccall((:iexa_last_error, LIB), Cstring, ()))
check(rc::Int32) = rc == 0 ? nothing : error("iexa error $rc: $(last_error())")

# ---- operator codes: the table of src/operators.jl:2-46 -----------------------------------------
const OP = Dict{Symbol,Int32}(
    :CONST => 0, :FIELD => 1, :VAR => 2, :PAR => 3,
    :+ => 10, :- => 11, :* => 12, :/ => 13, :^ => 14, :neg => 20, :pos => 21,
    :inv => 22, :sqrt => 23, :cbrt => 24, :abs => 25, :abs2 => 26, :exp => 27, :exp2 => 28, :log => 29,
    :log2 => 30, :log10 => 31, :log1p => 32, :sin => 33, :cos => 34, :tan => 35, :asin => 36, :acos => 37,
    :csc => 38, :sec => 39, :cot => 40, :atan => 41, :acot => 42, :sind => 43, :cosd => 44, :tand => 45,
    :cscd => 46, :secd => 47, :cotd => 48, :atand => 49, :acotd => 50, :sinh => 51, :cosh => 52, :tanh => 53,
    :csch => 38,   # the reference maps :csch to csc (src/operators.jl:41); the true csch is code 54
    :sech => 55, :coth => 56, :atanh => 57, :acoth => 58,
)

# ---- plan builder (twin of ExaModels.ExaCore as driven by src/transform.jl) -------------------------
mutable struct Plan
    h::Ptr{Cvoid}
    function Plan(; minimize::Bool = true)
        ref = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:iexa_plan_create, LIB), Int32, (Ref{Ptr{Cvoid}}, Int32), ref, minimize))
        p = new(ref[])
        finalizer(q -> ccall((:iexa_plan_destroy, LIB), Int32, (Ptr{Cvoid},), q.h), p)
        return p
    end
end

"`ExaModels.add_var` (transform.jl:113,154): returns the 0-based offset of the block"
function add_var!(p::Plan, x0::Vector{Float64}, lvar::Vector{Float64}, uvar::Vector{Float64})
    off = Ref{Int64}(0)
    check(ccall((:iexa_add_var, LIB), Int32, (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{Int64}),
                p.h, length(x0), x0, lvar, uvar, off))
    return off[]
end

"`ExaModels.add_par` (transform.jl:127,179); `vals` column-major like `vec(vals)`"
function add_par!(p::Plan, vals::AbstractArray{Float64})
    v = collect(vec(vals)); off = Ref{Int64}(0)
    check(ccall((:iexa_add_par, LIB), Int32, (Ptr{Cvoid}, Int64, Ptr{Float64}, Ref{Int64}), p.h, length(v), v, off))
    return off[]
end

"element-wise patch of x0/lvar/uvar (transform.jl:216-231); `i` 1-based"
patch_var!(p::Plan, which::Integer, i::Integer, v::Real) =
    check(ccall((:iexa_patch_var, LIB), Int32, (Ptr{Cvoid}, Int32, Int64, Float64), p.h, which, i, v))

"""
SoA iterator from the reference's `Vector{NamedTuple}` (transform.jl:31): integer fields become
int columns, floating fields fp columns.  Returns `(id, int_names, fp_names)`.
"""
function add_iterator!(p::Plan, itr::AbstractVector{<:NamedTuple})
    isempty(itr) && error("empty iterator")
    names = keys(first(itr))
    isempty(names) && return (Int32(0), Symbol[], Symbol[])          # [(;)] is iterator 0
    inames = [n for n in names if first(itr)[n] isa Integer]
    fnames = [n for n in names if !(first(itr)[n] isa Integer)]
    icols = [Int64[e[n] for e in itr] for n in inames]
    fcols = [Float64[e[n] for e in itr] for n in fnames]
    id = Ref{Int32}(0)
    GC.@preserve icols fcols begin
        ip = Ptr{Int64}[pointer(c) for c in icols]; fp = Ptr{Float64}[pointer(c) for c in fcols]
        check(ccall((:iexa_itr_base, LIB), Int32,
                    (Ptr{Cvoid}, Int64, Int32, Ptr{Ptr{Int64}}, Int32, Ptr{Ptr{Float64}}, Ref{Int32}),
                    p.h, length(itr), length(ip), ip, length(fp), fp, id))
    end
    return (id[], inames, fnames)
end

"product iterator, first factor fastest (transform.jl:445): nothing is materialised"
function product_iterator!(p::Plan, ids::Vector{Int32})
    id = Ref{Int32}(0)
    check(ccall((:iexa_itr_product, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Int32}, Ref{Int32}), p.h, length(ids), ids, id))
    return id[]
end

"`ExaModels.add_con` (transform.jl:458,559,597)"
function add_con!(p::Plan, tape::Vector{IexaNode}, idx::Vector{IexaIndex}, itr::Int32; lcon = 0.0, ucon = 0.0)
    off = Ref{Int64}(0)
    check(ccall((:iexa_add_con, LIB), Int32,
                (Ptr{Cvoid}, Ptr{IexaNode}, Int32, Ptr{IexaIndex}, Int32, Int32, Float64, Float64, Ref{Int64}),
                p.h, tape, length(tape), idx, length(idx), itr, lcon, ucon, off))
    return off[]
end

"`ExaModels.add_obj` (transform.jl:614,700,741)"
add_obj!(p::Plan, tape::Vector{IexaNode}, idx::Vector{IexaIndex}, itr::Int32) =
    check(ccall((:iexa_add_obj, LIB), Int32, (Ptr{Cvoid}, Ptr{IexaNode}, Int32, Ptr{IexaIndex}, Int32, Int32),
                p.h, tape, length(tape), idx, length(idx), itr))

# ---- the NLPModel --------------------------------------------------------------------------------
"""
    B200ExaModel(plan; device = 0, VT = CuVector{Float64})

`ExaModels.ExaModel(core)` replacement (src/infiniteopt_backend.jl:156).  `VT` decides where the
solver's vectors live, exactly like ExaModels' `backend` keyword: `CuVector{Float64}` for
MadNLP+cuDSS, `Vector{Float64}` for Ipopt (host buffers are copied inside each C call).
"""
mutable struct B200ExaModel{VT<:AbstractVector{Float64}} <: NLPModels.AbstractNLPModel{Float64,VT}
    meta::NLPModels.NLPModelMeta{Float64,VT}
    counters::NLPModels.Counters
    plan::Plan
    cmeta::IexaMeta
end

function B200ExaModel(plan::Plan; device::Integer = 0, rank::Integer = 0, world::Integer = 1,
                      VT::Type = CuVector{Float64})
    check(ccall((:iexa_finalize, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Int32, UInt32), plan.h, device, rank, world, 0))
    m = Ref{IexaMeta}()
    check(ccall((:iexa_get_meta, LIB), Int32, (Ptr{Cvoid}, Ref{IexaMeta}), plan.h, m))
    cm = m[]
    getv(which, n) = (v = zeros(n); n > 0 && check(ccall((:iexa_get_vector, LIB), Int32,
                      (Ptr{Cvoid}, Int32, Ptr{Float64}), plan.h, which, v)); VT(v))
    meta = NLPModels.NLPModelMeta{Float64,VT}(
        cm.nvar; ncon = cm.ncon, nnzj = cm.nnzj, nnzh = cm.nnzh,
        x0 = getv(0, cm.nvar), lvar = getv(1, cm.nvar), uvar = getv(2, cm.nvar),
        lcon = getv(3, cm.ncon), ucon = getv(4, cm.ncon), y0 = getv(5, cm.ncon), minimize = cm.minimize != 0)
    return B200ExaModel{VT}(meta, NLPModels.Counters(), plan, cm)
end

# memory space + raw pointer + stream of a solver vector
_ms(::CuArray) = MEM_DEVICE
_ms(::Array) = MEM_HOST
_ptr(v::CuArray) = reinterpret(Ptr{Cvoid}, pointer(v))
_ptr(v::Array) = Ptr{Cvoid}(pointer(v))
_ptr(::Nothing) = C_NULL
_st(v::CuArray) = reinterpret(Ptr{Cvoid}, stream().handle)     # order work on CUDA.jl's task-local stream
_st(::Array) = C_NULL

function NLPModels.obj(m::B200ExaModel, x::AbstractVector)
    f = Ref{Float64}(0.0)
    GC.@preserve x check(ccall((:iexa_obj, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ref{Float64}, Int32, Ptr{Cvoid}),
                               m.plan.h, _ptr(x), f, _ms(x), _st(x)))
    return f[]
end

for (jl, sym) in ((:grad!, :iexa_grad), (:cons_nln!, :iexa_cons), (:jac_coord!, :iexa_jac_coord))
    @eval function NLPModels.$jl(m::B200ExaModel, x::AbstractVector, out::AbstractVector)
        GC.@preserve x out check(ccall(($(QuoteNode(sym)), LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int32, Ptr{Cvoid}),
                                       m.plan.h, _ptr(x), _ptr(out), _ms(x), _st(x)))
        return out
    end
end

function NLPModels.hess_coord!(m::B200ExaModel, x::AbstractVector, y::AbstractVector, vals::AbstractVector;
                               obj_weight = 1.0)
    GC.@preserve x y vals check(ccall((:iexa_hess_coord, LIB), Int32,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Float64, Ptr{Cvoid}, Int32, Ptr{Cvoid}),
        m.plan.h, _ptr(x), _ptr(y), obj_weight, _ptr(vals), _ms(x), _st(x)))
    return vals
end
function NLPModels.hess_coord!(m::B200ExaModel, x::AbstractVector, vals::AbstractVector; obj_weight = 1.0)
    GC.@preserve x vals check(ccall((:iexa_hess_coord, LIB), Int32,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Float64, Ptr{Cvoid}, Int32, Ptr{Cvoid}),
        m.plan.h, _ptr(x), C_NULL, obj_weight, _ptr(vals), _ms(x), _st(x)))
    return vals
end

for (jl, sym) in ((:jac_structure!, :iexa_jac_structure), (:hess_structure!, :iexa_hess_structure))
    @eval function NLPModels.$jl(m::B200ExaModel, rows::AbstractVector{T}, cols::AbstractVector{T}) where {T<:Integer}
        GC.@preserve rows cols check(ccall(($(QuoteNode(sym)), LIB), Int32,
            (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int32, Int32, Ptr{Cvoid}),
            m.plan.h, _ptr(rows), _ptr(cols), sizeof(T), _ms(rows), _st(rows)))
        return rows, cols
    end
end

function NLPModels.jprod_nln!(m::B200ExaModel, x::AbstractVector, v::AbstractVector, Jv::AbstractVector)
    GC.@preserve x v Jv check(ccall((:iexa_jprod, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int32, Ptr{Cvoid}),
                                    m.plan.h, _ptr(x), _ptr(v), _ptr(Jv), _ms(x), _st(x)))
    return Jv
end
function NLPModels.jtprod_nln!(m::B200ExaModel, x::AbstractVector, v::AbstractVector, Jtv::AbstractVector)
    GC.@preserve x v Jtv check(ccall((:iexa_jtprod, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int32, Ptr{Cvoid}),
                                     m.plan.h, _ptr(x), _ptr(v), _ptr(Jtv), _ms(x), _st(x)))
    return Jtv
end
"cons! + jac_coord! + hess_coord! at the same (x, y) in ONE call (one fused kernel for CuArrays)"
function eval3!(m::B200ExaModel, x::AbstractVector, y::AbstractVector, c::AbstractVector, jvals::AbstractVector, hvals::AbstractVector;
                obj_weight = 1.0)
    GC.@preserve x y c jvals hvals check(ccall((:iexa_eval3, LIB), Int32,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Float64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int32, Ptr{Cvoid}),
        m.plan.h, _ptr(x), _ptr(y), obj_weight, _ptr(c), _ptr(jvals), _ptr(hvals), _ms(x), _st(x)))
    return c, jvals, hvals
end
function NLPModels.hprod!(m::B200ExaModel, x::AbstractVector, y::AbstractVector, v::AbstractVector, Hv::AbstractVector;
                          obj_weight = 1.0)
    GC.@preserve x y v Hv check(ccall((:iexa_hprod, LIB), Int32,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Float64, Ptr{Cvoid}, Int32, Ptr{Cvoid}),
        m.plan.h, _ptr(x), _ptr(y), _ptr(v), obj_weight, _ptr(Hv), _ms(x), _st(x)))
    return Hv
end

# ---- device-side transcription: columns described by a closed form, parameter functions as tapes ------------------------
struct IexaColGen
    kind::Int32
    src::Int32
    n::Int64
    a::Float64
    b::Float64
end
"iterator whose fp columns are generated ON the device (kind 1 linspace, 2 linspace + interval midpoints, 3 trapezoid weights of column `src`, 4 constant); `int_cols[j] == C_NULL` is 1:K"
function itr_generated(h::Ptr{Cvoid}, K::Integer, int_cols::Vector{Ptr{Int64}}, gens::Vector{IexaColGen}, fp_cols::Vector{Ptr{Float64}})
    out = Ref{Int32}(0)
    check(ccall((:iexa_itr_generated, LIB), Int32, (Ptr{Cvoid}, Int64, Int32, Ptr{Ptr{Int64}}, Int32, Ptr{IexaColGen}, Ptr{Ptr{Float64}}, Ref{Int32}),
                h, K, length(int_cols), int_cols, length(gens), gens, fp_cols, out))
    return out[]
end
"a parameter function (transform.jl:161-183) as a tape over iterator `itr`: its block of θ is evaluated on the device at finalize"
function add_par_function(h::Ptr{Cvoid}, nodes::Vector{IexaNode}, idx::Vector{IexaIndex}, itr::Integer)
    off = Ref{Int64}(0)
    check(ccall((:iexa_add_par_function, LIB), Int32, (Ptr{Cvoid}, Ptr{IexaNode}, Int32, Ptr{IexaIndex}, Int32, Int32, Ref{Int64}),
                h, nodes, length(nodes), idx, length(idx), itr, off))
    return off[]
end

"plan options (before the first generator): key 1 = slot-order policy (0 left to right, 1 right to left, 2 left to right with the first-order slots of every constraint row in column order: `jac_is_csr`), key 2 = strict IEEE"
set_option!(h::Ptr{Cvoid}, key::Integer, value::Integer) =
    check(ccall((:iexa_set_option, LIB), Int32, (Ptr{Cvoid}, Int32, Int64), h, key, value))

"`ExaModels.set_parameter!(core, param, vals)` (infiniteopt_backend.jl:522,546): in place, no rebuild"
function set_parameter!(m::B200ExaModel, offset0::Integer, vals::AbstractArray{Float64})
    v = collect(vec(vals))
    check(ccall((:iexa_set_par, LIB), Int32, (Ptr{Cvoid}, Int64, Int64, Ptr{Float64}), m.plan.h, offset0, length(v), v))
end

"the same update enqueued on the task-local CUDA stream (ordered like a callback: no device synchronisation)"
function set_parameter_async!(m::B200ExaModel, offset0::Integer, vals::AbstractArray{Float64})
    v = collect(vec(vals))
    check(ccall((:iexa_set_par_stream, LIB), Int32, (Ptr{Cvoid}, Int64, Int64, Ptr{Float64}, Ptr{Cvoid}),
                m.plan.h, offset0, length(v), v, CUDA.stream().handle))
end

"`model.θ` (infiniteopt_backend.jl:479)"
function Base.getproperty(m::B200ExaModel, s::Symbol)
    if s === :θ
        n = getfield(m, :cmeta).npar; v = zeros(n)
        n > 0 && check(ccall((:iexa_get_par, LIB), Int32, (Ptr{Cvoid}, Int64, Int64, Ptr{Float64}), getfield(m, :plan).h, 0, n, v))
        return v
    end
    return getfield(m, s)
end

"page-lock a host vector the solver reuses every iteration (Ipopt path): copies then run at PCIe speed"
function page_lock!(m::B200ExaModel, v::Array{Float64})
    GC.@preserve v check(ccall((:iexa_host_register, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int64), m.plan.h, _ptr(v), sizeof(v)))
    return v
end
function page_unlock!(m::B200ExaModel, v::Array{Float64})
    GC.@preserve v check(ccall((:iexa_host_unregister, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), m.plan.h, _ptr(v)))
    return v
end

"objective into a device scalar, no host synchronisation (GPU-resident solvers; the partial of a sharded model)"
function obj_device!(m::B200ExaModel, x::CuArray{Float64}, f::CuArray{Float64})
    GC.@preserve x f check(ccall((:iexa_obj_device, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
                                 m.plan.h, _ptr(x), _ptr(f), _st(x)))
    return f
end

# ---- multi-GPU queries (one Julia process per GPU; iexa_finalize(rank, world) shards the supports) ---------------------
struct IexaSegment
    global_start::Int64
    local_start::Int64
    length::Int64
end

"local ranges of rows (0) / Jacobian slots (1) / Hessian slots (2) and their global positions (0-based)"
function segments(m::B200ExaModel, which::Integer)
    n = ccall((:iexa_segments, LIB), Int64, (Ptr{Cvoid}, Int32, Ptr{IexaSegment}, Int64), m.plan.h, which, C_NULL, 0)
    out = Vector{IexaSegment}(undef, max(n, 0))
    n > 0 && ccall((:iexa_segments, LIB), Int64, (Ptr{Cvoid}, Int32, Ptr{IexaSegment}, Int64), m.plan.h, which, out, n)
    return out
end

"1-based variable indices whose gradient entries must be all-reduced after grad! (NCCL.jl on the same buffer)"
function shared_vars(m::B200ExaModel)
    n = ccall((:iexa_shared_vars, LIB), Int64, (Ptr{Cvoid}, Ptr{Int64}, Int64), m.plan.h, C_NULL, 0)
    out = Vector{Int64}(undef, max(n, 0))
    n > 0 && ccall((:iexa_shared_vars, LIB), Int64, (Ptr{Cvoid}, Ptr{Int64}, Int64), m.plan.h, out, n)
    return out
end

"the same set as merged 0-based ranges (shard-boundary entries of shifted references and product-iterator indices included)"
function shared_ranges(m::B200ExaModel)
    n = ccall((:iexa_shared_ranges, LIB), Int64, (Ptr{Cvoid}, Ptr{IexaSegment}, Int64), m.plan.h, C_NULL, 0)
    out = Vector{IexaSegment}(undef, max(n, 0))
    n > 0 && ccall((:iexa_shared_ranges, LIB), Int64, (Ptr{Cvoid}, Ptr{IexaSegment}, Int64), m.plan.h, out, n)
    return out
end

"bytes of x a host-memory callback uploads on this rank (the read ranges only when the model is sharded)"
host_x_bytes(m::B200ExaModel) = ccall((:iexa_host_x_bytes, LIB), Int64, (Ptr{Cvoid},), m.plan.h)

"the ranges of x this rank reads (own supports + shared variables + halos): what a distributed solver keeps current here"
function x_ranges(m::B200ExaModel)
    n = ccall((:iexa_x_ranges, LIB), Int64, (Ptr{Cvoid}, Ptr{IexaSegment}, Int64), m.plan.h, C_NULL, 0)
    out = Vector{IexaSegment}(undef, max(n, 0))
    n > 0 && ccall((:iexa_x_ranges, LIB), Int64, (Ptr{Cvoid}, Ptr{IexaSegment}, Int64), m.plan.h, out, n)
    return out
end

# ---- x halo exchange + small all-reduce over NVLink peer memory (csrc/halo.cu; one Julia process per GPU) ---------------
# Setup (collective): h = PeerHalo(device, rank, world); (hx, off, hf) = export_handles(h, x); ship them to every peer (MPI.jl /
# Distributed); connect!(h, peer, hx_peer, off_peer, hf_peer); set_sends!/set_recvs! from the ranks' x_ranges.
mutable struct PeerHalo
    h::Ptr{Cvoid}
end
function PeerHalo(device::Integer, rank::Integer, world::Integer)
    r = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:iexa_halo_create, LIB), Int32, (Ref{Ptr{Cvoid}}, Int32, Int32, Int32), r, device, rank, world))
    p = PeerHalo(r[])
    finalizer(p -> ccall((:iexa_halo_destroy, LIB), Int32, (Ptr{Cvoid},), p.h), p)
    return p
end
function export_handles(p::PeerHalo, x::CuArray{Float64})
    hx = zeros(UInt8, 64); hf = zeros(UInt8, 64); off = Ref{Int64}(0)
    GC.@preserve x check(ccall((:iexa_halo_export, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{UInt8}, Ref{Int64}, Ptr{UInt8}),
                               p.h, _ptr(x), hx, off, hf))
    return hx, off[], hf
end
connect!(p::PeerHalo, peer::Integer, hx::Vector{UInt8}, off::Integer, hf::Vector{UInt8}) =
    check(ccall((:iexa_halo_connect, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{UInt8}, Int64, Ptr{UInt8}), p.h, peer, hx, off, hf))
"lo_hi: 0-based [lo, hi) pairs of x this rank owns and `peer` reads"
set_sends!(p::PeerHalo, peer::Integer, lo_hi::Vector{Int64}) =
    check(ccall((:iexa_halo_set_sends, LIB), Int32, (Ptr{Cvoid}, Int32, Int64, Ptr{Int64}), p.h, peer, length(lo_hi) ÷ 2, lo_hi))
set_recvs!(p::PeerHalo, peers::Vector{Int32}) =
    check(ccall((:iexa_halo_set_recvs, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Int32}), p.h, length(peers), peers))
"collective: push the shared variables / shard-boundary halos this rank owns into its readers' x (one kernel, no host sync)"
function exchange!(p::PeerHalo, x::CuArray{Float64})
    GC.@preserve x check(ccall((:iexa_halo_exchange, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), p.h, _ptr(x), _st(x)))
    return x
end
"collective: buf .= sum over ranks (length <= 1024), summed in rank order — objective + shared gradient slice in one kernel"
function allreduce_small!(p::PeerHalo, buf::CuArray{Float64})
    GC.@preserve buf check(ccall((:iexa_halo_allreduce_small, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Ptr{Cvoid}),
                                 p.h, _ptr(buf), length(buf), _st(buf)))
    return buf
end
status(p::PeerHalo) = ccall((:iexa_halo_status, LIB), Int64, (Ptr{Cvoid},), p.h)

# ---- KKT assembly: COO -> CSR value map for cuDSS (replaces MadNLPGPU's transfer! kernel) ------------------------------
mutable struct CsrMap
    h::Ptr{Cvoid}
    nnz::Int64
end

"""
`jac_is_csr(m)`: under `slot_order = 2` (IEXA_SLOT_ORDER_JAC_ROW_SORTED) the array `jac_coord!` writes is already a CSR value
array; `jac_csr_rowptr!(m, rowptr)` fills the `ncon + 1` zero-based row pointers and `jac_structure!`'s cols are the column
indices, so a KKT assembly needs no COO->CSR pass for the Jacobian.
"""
"engine-held device bytes of this rank: (columns, columns_unsharded, theta, theta_unsharded, programs_tables, host_path_staging)"
function device_bytes(m::B200ExaModel)
    out = zeros(Int64, 6)
    GC.@preserve out check(ccall((:iexa_device_bytes, LIB), Int32, (Ptr{Cvoid}, Ptr{Int64}), m.plan.h, out))
    return NamedTuple{(:columns, :columns_unsharded, :theta, :theta_unsharded, :programs_tables, :host_path_staging)}(Tuple(out))
end
function jac_is_csr(m::B200ExaModel)
    out = Ref{Int32}(0)
    check(ccall((:iexa_jac_is_csr, LIB), Int32, (Ptr{Cvoid}, Ptr{Int32}), m.plan.h, out))
    return out[] == 1
end
function jac_csr_rowptr!(m::B200ExaModel, rowptr::AbstractVector{T}) where {T<:Union{Int32,Int64}}
    GC.@preserve rowptr check(ccall((:iexa_jac_csr_rowptr, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Int32, Ptr{Cvoid}),
                                    m.plan.h, _ptr(rowptr), sizeof(T), _ms(rowptr), _st(rowptr)))
    return rowptr
end

"locality keys of the Jacobian (0) / Hessian (1) COO slots, for `CsrMap(...; keys)` of patterns with duplicates"
function coo_locality!(m::B200ExaModel, which::Integer, keys::AbstractVector{Int32})
    GC.@preserve keys check(ccall((:iexa_coo_locality, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Cvoid}, Int32, Ptr{Cvoid}),
                                  m.plan.h, which, _ptr(keys), _ms(keys), _st(keys)))
    return keys
end

function CsrMap(nrows::Integer, ncols::Integer, rows::AbstractVector{T}, cols::AbstractVector{T};
                keys::Union{Nothing,AbstractVector{Int32}} = nothing, device::Integer = 0) where {T<:Integer}
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve rows cols keys check(ccall((:iexa_csr_create_keyed, LIB), Int32,
        (Ref{Ptr{Cvoid}}, Int64, Int64, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Int32, Ptr{Cvoid}, Int32, Int32),
        h, nrows, ncols, length(rows), _ptr(rows), _ptr(cols), sizeof(T), _ptr(keys), _ms(rows), device))
    c = CsrMap(h[], ccall((:iexa_csr_nnz, LIB), Int64, (Ptr{Cvoid},), h[]))
    finalizer(c -> ccall((:iexa_csr_destroy, LIB), Int32, (Ptr{Cvoid},), c.h), c)
    return c
end

"0-based CSR pattern (cuDSS convention)"
function pattern!(c::CsrMap, rowptr::AbstractVector{Int32}, colind::AbstractVector{Int32})
    GC.@preserve rowptr colind check(ccall((:iexa_csr_pattern, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int32),
                                           c.h, _ptr(rowptr), _ptr(colind), _ms(rowptr)))
    return rowptr, colind
end

"COO values -> CSR values, duplicates summed, no atomics; once per iteration"
function apply!(c::CsrMap, coo_vals::AbstractVector{Float64}, csr_vals::AbstractVector{Float64})
    GC.@preserve coo_vals csr_vals check(ccall((:iexa_csr_apply, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int32, Ptr{Cvoid}),
                                               c.h, _ptr(coo_vals), _ptr(csr_vals), _ms(coo_vals), _st(coo_vals)))
    return csr_vals
end

end # module
