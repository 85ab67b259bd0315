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
    with_ws(bytes) do ws, n
        b200_reduce(ws, n, A, R, din, dout, opc, mapc, pre, red, post,
                    init === nothing ? INIT_FROM_OUT : INIT_NEUTRAL, stream())
    end
    return R                                                     # mapreduce.jl:211,300
end
