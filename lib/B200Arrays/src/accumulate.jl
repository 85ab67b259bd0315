# CUDACore.scan! for dense CuArrays and built-in ops — replaces CUDACore/src/accumulate.jl:135-199.
# All four Base._accumulate! methods and accumulate_pairwise! (accumulate.jl:204-216) funnel
# into scan!, so hooking it avoids method ambiguities (SURVEY §8b).

function CUDACore.scan!(f::OP, output::CuArray{To}, input::CuArray{Ti};
                        dims::Integer, init=nothing, neutral=GPUArrays.neutral_element(f, To)) where
                        {OP<:GraftOps,To<:ReduceTypes,Ti<:ReduceTypes}
    reference() = invoke(CUDACore.scan!, Tuple{Function,CUDACore.AnyCuArray,CUDACore.AnyCuArray},
                         f, output, input; dims, init, neutral)
    enabled() || return reference()
    # the reference kernel loads every element as op(neutral, x) (accumulate.jl:39-43); GPUArrays' default for integer `&`
    # is one(T), which is not neutral — that (reference-defined) result is not the plain scan the library computes, so
    # the graft only answers for a neutral that IS neutral (true_neutral, mapreduce.jl)
    neutral === true_neutral(f, To) || return reference()
    dims > 0 || throw(ArgumentError("dims must be a positive integer"))          # accumulate.jl:137
    axes(output) == axes(input) || throw(DimensionMismatch("shape of B must match A"))
    dims > ndims(input) && return copyto!(output, input)
    isempty(axes(input)[dims]) && return output
    sz = size(input)
    pre, n, post = prod(sz[1:dims-1]), sz[dims], prod(sz[dims+1:end])
    din, dout, opc = dtype_code(Ti), dtype_code(To), op_code(f)
    bytes = b200_scan_workspace_bytes(din, dout, opc, pre, n, post)
    bytes === nothing && return reference()                     # (Ti, To, op) not compiled into the library
    initref = init === nothing ? nothing : Ref{To}(convert(To, something(init)))
    GC.@preserve initref with_ws(bytes) do ws, nb
        ptr = initref === nothing ? C_NULL : Base.unsafe_convert(Ptr{Cvoid}, Base.cconvert(Ptr{To}, initref))
        b200_scan(ws, nb, input, output, din, dout, opc, Int32(0), pre, n, post, ptr, stream())
    end
    return output
end
