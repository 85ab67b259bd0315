# More-specific methods of GPUArrays.mapreducedim! for dense CuArrays and built-in (f, op, T)
# — replaces CUDACore/src/mapreduce.jl:168-301 for that matrix; everything else (closures,
# Broadcasted inputs, wrapped outputs, non-neutral init, non-contiguous reduced dims)
# `invoke`s the reference's generic method.

op_code(::typeof(Base.add_sum)) = OP_ADD
op_code(::typeof(+)) = OP_ADD
op_code(::typeof(Base.mul_prod)) = OP_MUL
op_code(::typeof(*)) = OP_MUL
op_code(::typeof(max)) = OP_MAX
op_code(::typeof(min)) = OP_MIN
op_code(::typeof(&)) = OP_AND
op_code(::typeof(|)) = OP_OR
map_code(::typeof(identity)) = MAP_IDENTITY
map_code(::typeof(abs2)) = MAP_ABS2

const GraftOps = Union{typeof(Base.add_sum),typeof(+),typeof(Base.mul_prod),typeof(*),typeof(max),typeof(min),
                       typeof(&),typeof(|)}
const GraftMaps = Union{typeof(identity),typeof(abs2)}

# (pre, red, post) of the contiguous group of reduced dims, or nothing (mapreduce.jl:187-200)
function factor_dims(szA::Dims, szR::Dims)
    nd = length(szA)
    szR = ntuple(d -> d <= length(szR) ? szR[d] : 1, nd)
    reduced = [d for d in 1:nd if szR[d] == 1 && szA[d] != 1]
    isempty(reduced) && return (prod(szA), 1, 1)
    lo, hi = first(reduced), last(reduced)
    all(d -> szA[d] == 1 || d in reduced, lo:hi) || return nothing
    return (prod(szA[1:lo-1]), prod(szA[lo:hi]), prod(szA[hi+1:end]))
end

# The TRUE neutral element of `op` on `T`.  GPUArrays.neutral_element agrees with it except for integer `&`, where it
# returns one(T) (not all-ones): the reference seeds every thread with that value (mapreduce.jl:114,148), so its result
# is `reduce(&, A) & 1`-like and launch-shape dependent — that case must stay on the reference path (SURVEY §8 a1').
true_neutral(::Union{typeof(+),typeof(Base.add_sum)}, ::Type{T}) where {T} = zero(T)
true_neutral(::Union{typeof(*),typeof(Base.mul_prod)}, ::Type{T}) where {T} = one(T)
true_neutral(::typeof(max), ::Type{T}) where {T} = typemin(T)
true_neutral(::typeof(min), ::Type{T}) where {T} = typemax(T)
true_neutral(::typeof(|), ::Type{T}) where {T} = zero(T)
true_neutral(::typeof(&), ::Type{Bool}) = true
true_neutral(::typeof(&), ::Type{T}) where {T<:Integer} = ~zero(T)
true_neutral(::typeof(&), ::Type{T}) where {T} = nothing          # bitwise ops on floats: never grafted

reference_mapreducedim!(f, op, R, A; init) =
    invoke(GPUArrays.mapreducedim!, Tuple{Any,Any,CUDACore.AnyCuArray,Union{AbstractArray,Broadcast.Broadcasted}},
           f, op, R, A; init)

function GPUArrays.mapreducedim!(f::F, op::OP, R::CuArray{To}, A::CuArray{Ti};
                                 init=nothing) where {F<:GraftMaps,OP<:GraftOps,To<:ReduceTypes,Ti<:ReduceTypes}
    enabled() || return reference_mapreducedim!(f, op, R, A; init)
    Base.check_reducedims(R, A)
    length(A) == 0 && return R                                   # mapreduce.jl:175
    dims = factor_dims(size(A), size(R))
    # `init` must be neutral (what _mapreduce / sum! / count pass); see SURVEY §8 a1
    neutral_ok = init === nothing || init === true_neutral(op, To)
    (dims === nothing || !neutral_ok) && return reference_mapreducedim!(f, op, R, A; init)
    pre, red, post = dims
    din, dout, opc, mapc = dtype_code(Ti), dtype_code(To), op_code(op), map_code(f)
    bytes = b200_reduce_workspace_bytes(din, dout, opc, mapc, pre, red, post)
    bytes === nothing && return reference_mapreducedim!(f, op, R, A; init)   # (Ti, To, op) not compiled into the library
    with_ws(bytes) do ws, n
        b200_reduce(ws, n, A, R, din, dout, opc, mapc, pre, red, post,
                    init === nothing ? INIT_FROM_OUT : INIT_NEUTRAL, stream())
    end
    return R                                                     # mapreduce.jl:211,300
end

# extrema(A; dims) = mapreduce(Base.ExtremaMap(identity), Base._extrema_rf, A; dims, init=(typemax(T), typemin(T)))
# (Base, reached through GPUArrays._mapreduce -> mapreducedim!, CUDACore/src/mapreduce.jl:168).  The output eltype is
# Tuple{T,T}: an isbits pair laid out (min, max) — exactly B200_OP_EXTREMA's "2 x dtype per output" (b200arrays.h:90).
# Only the identity map and the neutral init (or no init with a pre-seeded R) are grafted.
function GPUArrays.mapreducedim!(f::Base.ExtremaMap{typeof(identity)}, op::typeof(Base._extrema_rf),
                                 R::CuArray{Tuple{T,T}}, A::CuArray{T}; init=nothing) where {T<:ReduceTypes}
    enabled() || return reference_mapreducedim!(f, op, R, A; init)
    Base.check_reducedims(R, A)
    length(A) == 0 && return R
    dims = factor_dims(size(A), size(R))
    neutral_ok = init === nothing || init === (typemax(T), typemin(T))
    (dims === nothing || !neutral_ok || T === Bool) && return reference_mapreducedim!(f, op, R, A; init)
    pre, red, post = dims
    din = dtype_code(T)
    bytes = b200_reduce_workspace_bytes(din, din, OP_EXTREMA, MAP_IDENTITY, pre, red, post)
    bytes === nothing && return reference_mapreducedim!(f, op, R, A; init)
    with_ws(bytes) do ws, n
        b200_reduce(ws, n, A, R, din, din, OP_EXTREMA, MAP_IDENTITY, pre, red, post,
                    init === nothing ? INIT_FROM_OUT : INIT_NEUTRAL, stream())
    end
    return R
end

# findmax / findmin (GPUArrays: Base.findmax(a::AnyGPUArray; dims=:) -> findminmax -> mapreducedim! with a
# tuple-valued op; tests test/core/array.jl:767-769,782-784).  One (pre, red, post) call to b200_findminmax;
# linear indices are turned into CartesianIndices on the host side of the API like GPUArrays does.
function b200_findminmax(ws, ws_bytes, A, vals, idx, din, is_max, pre, red, post, s::CuStream)
    initialize_context()
    check(@gcsafe_ccall libb200arrays[].b200_findminmax(ws::CuPtr{Cvoid}, ws_bytes::Csize_t, A::CuPtr{Cvoid},
              vals::CuPtr{Cvoid}, idx::CuPtr{Cvoid}, din::Int32, is_max::Int32, pre::Int64, red::Int64, post::Int64,
              s::CUDACore.CUstream)::Int32)
end

function graft_findminmax(A::CuArray{T}, dims, is_max::Bool) where {T<:ReduceTypes}
    isempty(A) && throw(ArgumentError("reducing over an empty collection is not allowed"))
    rsz = dims === Colon() ? ntuple(_ -> 1, ndims(A)) : ntuple(d -> d in dims ? 1 : size(A, d), ndims(A))
    f = factor_dims(size(A), rsz)
    f === nothing && return nothing
    pre, red, post = f
    bytes = Ref{Csize_t}(0)
    # B200_UNSUPPORTED (element types the kernels are not instantiated for) -> nothing -> reference method
    nbytes = query_bytes((@ccall libb200arrays[].b200_findminmax_workspace_bytes(dtype_code(T)::Int32, pre::Int64,
                              red::Int64, post::Int64, bytes::Ref{Csize_t})::Int32), bytes)
    nbytes === nothing && return nothing
    vals = similar(A, T, rsz); idx = similar(A, Int64, rsz)
    with_ws(nbytes) do ws, nb
        b200_findminmax(ws, nb, A, vals, idx, dtype_code(T), Int32(is_max), pre, red, post, stream())
    end
    if dims === Colon()
        v, i = CUDACore.@allowscalar(vals[1]), CUDACore.@allowscalar(idx[1])
        return (v, ndims(A) == 1 ? i : CartesianIndices(A)[i])
    end
    # keys(A) of a vector are Ints (Base.findmax(A; dims) returns linear indices there), CartesianIndices otherwise
    ndims(A) == 1 && return (vals, idx)
    ci = CartesianIndices(A)
    return (vals, map(i -> ci[i], idx))
end

for (fn, mx) in ((:findmax, true), (:findmin, false))
    @eval function Base.$fn(A::CuArray{T}; dims=:) where {T<:ReduceTypes}
        r = enabled() ? graft_findminmax(A, dims, $mx) : nothing
        r === nothing ? invoke(Base.$fn, Tuple{GPUArrays.AnyGPUArray}, A; dims) : r
    end
end
