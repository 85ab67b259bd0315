# Base.findall(::CuArray{Bool}) — the in-tree caller of cumsum (CUDACore/src/indexing.jl:26-66): the reference runs
# cumsum over the flags, reads indices[end] back, then launches a scatter kernel.  b200_select_flagged fuses count,
# decoupled look-back and compaction into one pass; logical indexing `A[mask]` goes through the same method
# (Base.to_index -> findall, indexing.jl:15-21).

function b200_select_workspace_bytes(n)
    bytes = Ref{Csize_t}(0)
    check(@ccall libb200arrays[].b200_select_workspace_bytes(n::Int64, bytes::Ref{Csize_t})::Int32)
    return Int(bytes[])
end

function b200_select_flagged(ws, ws_bytes, flags, indices, dindex, count, n, s::CuStream)
    initialize_context()
    check(@gcsafe_ccall libb200arrays[].b200_select_flagged(ws::CuPtr{Cvoid}, ws_bytes::Csize_t, flags::CuPtr{Cvoid},
              indices::CuPtr{Cvoid}, dindex::Int32, count::CuPtr{Cvoid}, n::Int64, s::CUDACore.CUstream)::Int32)
end

reference_findall(bools) = invoke(Base.findall, Tuple{CUDACore.AnyCuArray{Bool}}, bools)

function Base.findall(bools::CuArray{Bool,N}) where {N}
    enabled() || return reference_findall(bools)
    n = length(bools)
    n == 0 && return CuArray{keytype(bools)}(undef, 0)
    linear = CuVector{Int64}(undef, n)
    count = CuVector{Int64}(undef, 1)
    with_ws(b200_select_workspace_bytes(n)) do ws, nb
        b200_select_flagged(ws, nb, bools, linear, dtype_code(Int64), count, n, stream())
    end
    m = Int(Array(count)[1])                  # same single read-back as the reference's indices[end] (indexing.jl:30)
    unsafe_free!(count)
    ys = if N == 1
        m == n ? linear : linear[1:m]         # keytype of a vector is Int: the linear indices are the answer
    else
        # N-d: keytype is CartesianIndex{N}; convert the m linear indices (a broadcast over m elements, not n)
        ci = CartesianIndices(bools)
        map(i -> @inbounds(ci[i]), view(linear, 1:m))
    end
    ys === linear || unsafe_free!(linear)
    return ys
end
