# Raw bindings, one per symbol of include/b200arrays.h.  Every wrapper makes the task's
# context current first (initialize_context, CUDACore/lib/cudadrv/libcuda.jl:20-23) and lets
# the GC run during the call (@gcsafe_ccall), exactly like lib/curand/src/libcurand.jl:210-215.
# Passing a CuArray where CuPtr{Cvoid} is declared runs unsafe_convert -> take_ownership!
# on the task-local stream (CUDACore/src/array.jl:467, memory.jl:597-660).

# b200_dtype_t
dtype_code(::Type{Float16}) = Int32(0)
dtype_code(::Type{BFloat16}) = Int32(1)
dtype_code(::Type{Float32}) = Int32(2)
dtype_code(::Type{Float64}) = Int32(3)
dtype_code(::Type{Int32}) = Int32(4)
dtype_code(::Type{Int64}) = Int32(5)
dtype_code(::Type{UInt32}) = Int32(6)
dtype_code(::Type{UInt64}) = Int32(7)
dtype_code(::Type{Bool}) = Int32(8)
dtype_code(::Type{Int8}) = Int32(9)
dtype_code(::Type{UInt8}) = Int32(10)
dtype_code(::Type{Int16}) = Int32(11)
dtype_code(::Type{UInt16}) = Int32(12)

const ReduceTypes = Union{Float16,BFloat16,Float32,Float64,Int32,Int64,UInt32,UInt64,Bool,Int8,UInt8,Int16,UInt16}
const SortTypes = ReduceTypes

# b200_op_t / b200_map_t
const OP_ADD, OP_MUL, OP_MAX, OP_MIN, OP_AND, OP_OR, OP_EXTREMA = Int32.(0:6)
const MAP_IDENTITY, MAP_ABS2 = Int32(0), Int32(1)
const INIT_NEUTRAL, INIT_FROM_OUT = Int32(0), Int32(1)
const SORTPERM_INDEX = Int32(100)

# The *_workspace_bytes queries double as the "is this (eltype, op) combination compiled in?" probe:
# B200_UNSUPPORTED -> `nothing`, and the caller `invoke`s the reference method instead of throwing.
@inline function query_bytes(res::Int32, bytes)
    res == B200_UNSUPPORTED && return nothing
    check(res)
    return Int(bytes[])
end

function b200_reduce_workspace_bytes(din, dout, op, map, pre, red, post)
    bytes = Ref{Csize_t}(0)
    query_bytes((@ccall libb200arrays[].b200_reduce_workspace_bytes(din::Int32, dout::Int32, op::Int32, map::Int32,
                     pre::Int64, red::Int64, post::Int64, bytes::Ref{Csize_t})::Int32), bytes)
end

function b200_reduce(ws, ws_bytes, A, R, din, dout, op, map, pre, red, post, init_mode, s::CuStream)
    initialize_context()
    check(@gcsafe_ccall libb200arrays[].b200_reduce(ws::CuPtr{Cvoid}, ws_bytes::Csize_t, A::CuPtr{Cvoid}, R::CuPtr{Cvoid},
              din::Int32, dout::Int32, op::Int32, map::Int32, pre::Int64, red::Int64, post::Int64,
              init_mode::Int32, s::CUDACore.CUstream)::Int32)
end

function b200_scan_workspace_bytes(din, dout, op, pre, n, post)
    bytes = Ref{Csize_t}(0)
    query_bytes((@ccall libb200arrays[].b200_scan_workspace_bytes(din::Int32, dout::Int32, op::Int32,
                     pre::Int64, n::Int64, post::Int64, bytes::Ref{Csize_t})::Int32), bytes)
end

function b200_scan(ws, ws_bytes, A, B, din, dout, op, exclusive, pre, n, post, init::Ptr{Cvoid}, s::CuStream)
    initialize_context()
    check(@gcsafe_ccall libb200arrays[].b200_scan(ws::CuPtr{Cvoid}, ws_bytes::Csize_t, A::CuPtr{Cvoid}, B::CuPtr{Cvoid},
              din::Int32, dout::Int32, op::Int32, exclusive::Int32, pre::Int64, n::Int64, post::Int64,
              init::Ptr{Cvoid}, s::CUDACore.CUstream)::Int32)
end

function b200_sort_workspace_bytes(dkey, dpayload, n, nseg)
    bytes = Ref{Csize_t}(0)
    check(@ccall libb200arrays[].b200_sort_workspace_bytes(dkey::Int32, dpayload::Int32, n::Int64, nseg::Int64,
              bytes::Ref{Csize_t})::Int32)
    return Int(bytes[])
end

function b200_sort_keys(ws, ws_bytes, keys, dkey, n, nseg, descending, s::CuStream)
    initialize_context()
    check(@gcsafe_ccall libb200arrays[].b200_sort_keys(ws::CuPtr{Cvoid}, ws_bytes::Csize_t, keys::CuPtr{Cvoid},
              dkey::Int32, n::Int64, nseg::Int64, descending::Int32, s::CUDACore.CUstream)::Int32)
end

function b200_sortperm(ws, ws_bytes, keys, perm, dkey, dindex, n, nseg, descending, linear, s::CuStream)
    initialize_context()
    check(@gcsafe_ccall libb200arrays[].b200_sortperm(ws::CuPtr{Cvoid}, ws_bytes::Csize_t, keys::CuPtr{Cvoid},
              perm::CuPtr{Cvoid}, dkey::Int32, dindex::Int32, n::Int64, nseg::Int64, descending::Int32,
              linear::Int32, s::CUDACore.CUstream)::Int32)
end

function b200_select_kth_workspace_bytes(dkey, n)
    bytes = Ref{Csize_t}(0)
    check(@ccall libb200arrays[].b200_select_kth_workspace_bytes(dkey::Int32, n::Int64, bytes::Ref{Csize_t})::Int32)
    return Int(bytes[])
end

function b200_select_kth(ws, ws_bytes, keys, dkey, n, k, descending, out, s::CuStream)
    initialize_context()
    check(@gcsafe_ccall libb200arrays[].b200_select_kth(ws::CuPtr{Cvoid}, ws_bytes::Csize_t, keys::CuPtr{Cvoid},
              dkey::Int32, n::Int64, k::Int64, descending::Int32, out::CuPtr{Cvoid}, s::CUDACore.CUstream)::Int32)
end

function b200_sort_dims_workspace_bytes(dkey, sortperm, pre, n, post)
    bytes = Ref{Csize_t}(0)
    check(@ccall libb200arrays[].b200_sort_dims_workspace_bytes(dkey::Int32, sortperm::Int32, pre::Int64, n::Int64,
              post::Int64, bytes::Ref{Csize_t})::Int32)
    return Int(bytes[])
end

function b200_sort_keys_dims(ws, ws_bytes, keys, dkey, pre, n, post, descending, s::CuStream)
    initialize_context()
    check(@gcsafe_ccall libb200arrays[].b200_sort_keys_dims(ws::CuPtr{Cvoid}, ws_bytes::Csize_t, keys::CuPtr{Cvoid},
              dkey::Int32, pre::Int64, n::Int64, post::Int64, descending::Int32, s::CUDACore.CUstream)::Int32)
end

function b200_sortperm_dims(ws, ws_bytes, keys, perm, dkey, dindex, pre, n, post, descending, linear, s::CuStream)
    initialize_context()
    check(@gcsafe_ccall libb200arrays[].b200_sortperm_dims(ws::CuPtr{Cvoid}, ws_bytes::Csize_t, keys::CuPtr{Cvoid},
              perm::CuPtr{Cvoid}, dkey::Int32, dindex::Int32, pre::Int64, n::Int64, post::Int64, descending::Int32,
              linear::Int32, s::CUDACore.CUstream)::Int32)
end

# Temporaries come from CUDA.jl's stream-ordered pool, freed in stream order
# (with_workspace, CUDACore/src/utils/call.jl:23-103).
function with_ws(f, bytes::Int)
    if bytes == 0
        return f(CUDACore.CU_NULL, 0)
    end
    ws = CuVector{UInt8}(undef, bytes)
    try
        return f(ws, bytes)
    finally
        unsafe_free!(ws)
    end
end
