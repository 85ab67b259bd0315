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
    neutral_ok = init === nothing || init === GPUArrays.neutral_element(op, To)
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
    vals = similar(A, T, rsz); idx = similar(A, Int64, rsz)
    bytes = Ref{Csize_t}(0)
    check(@ccall libb200arrays[].b200_findminmax_workspace_bytes(dtype_code(T)::Int32, pre::Int64, red::Int64,
              post::Int64, bytes::Ref{Csize_t})::Int32)
    with_ws(Int(bytes[])) do ws, nb
        b200_findminmax(ws, nb, A, vals, idx, dtype_code(T), Int32(is_max), pre, red, post, stream())
    end
    if dims === Colon()
        v, i = CUDACore.@allowscalar(vals[1]), CUDACore.@allowscalar(idx[1])
        return (v, ndims(A) == 1 ? i : CartesianIndices(A)[i])
    end
    ci = CartesianIndices(A)
    return (vals, map(i -> ci[i], idx))
end

for (fn, mx) in ((:findmax, true), (:findmin, false))
    @eval function Base.$fn(A::CuArray{T}; dims=:) where {T<:ReduceTypes}
        r = enabled() ? graft_findminmax(A, dims, $mx) : nothing
        r === nothing ? invoke(Base.$fn, Tuple{GPUArrays.AnyGPUArray}, A; dims) : r
    end
end
