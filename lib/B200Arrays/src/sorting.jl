# Base.sort! / sortperm! / partialsort! for dense CuArrays along any `dims` with
# lt=isless, by=identity — replaces the bitonic path of CUDACore/src/sorting.jl:1003-1070.
# kwargs do not dispatch, so each hook guards at run time and `invoke`s the reference
# method otherwise (custom lt/by, alg=QuickSort, user-supplied initialized=true).

graftable(lt, by, alg) = lt === isless && by === identity && alg === CUDACore.BitonicSort

function Base.sort!(c::CuArray{T}; alg::CUDACore.SortingAlgorithm=CUDACore.BitonicSort, lt=isless, by=identity,
                    rev::Bool=false, dims::Integer=1) where {T<:SortTypes}
    if !enabled() || !graftable(lt, by, alg) || !(1 <= dims <= ndims(c))
        return invoke(sort!, Tuple{CUDACore.AnyCuArray}, c; alg, lt, by, rev, dims)
    end
    sz = size(c)
    pre, n, post = prod(sz[1:dims-1]), sz[dims], prod(sz[dims+1:end])
    (n <= 1 || pre * post == 0) && return c
    bytes = b200_sort_dims_workspace_bytes(dtype_code(T), Int32(0), pre, n, post)
    with_ws(bytes) do ws, nb
        b200_sort_keys_dims(ws, nb, c, dtype_code(T), pre, n, post, Int32(rev), stream())
    end
    return c
end

function Base.sortperm!(ix::CuArray{I}, A::CuArray{T}; initialized::Bool=false,
                        alg::CUDACore.SortingAlgorithm=CUDACore.BitonicSort, lt=isless, by=identity,
                        rev::Bool=false, dims::Integer=1) where {I<:Union{Int32,Int64,UInt32,UInt64},T<:SortTypes}
    if axes(ix) != axes(A)
        throw(ArgumentError("index array must have the same size/axes as the source array, $(axes(ix)) != $(axes(A))"))
    end
    isempty(A) && return ix
    if !enabled() || !graftable(lt, by, alg) || initialized || !(1 <= dims <= ndims(A))
        return invoke(sortperm!, Tuple{CUDACore.AnyCuArray,CUDACore.AnyCuArray}, ix, A; initialized, alg, lt, by, rev, dims)
    end
    sz = size(A)
    pre, n, post = prod(sz[1:dims-1]), sz[dims], prod(sz[dims+1:end])
    bytes = b200_sort_dims_workspace_bytes(dtype_code(T), Int32(1), pre, n, post)
    with_ws(bytes) do ws, nb
        b200_sortperm_dims(ws, nb, A, ix, dtype_code(T), dtype_code(I), pre, n, post, Int32(rev),
                           Int32(ndims(A) > 1), stream())
    end
    return ix
end

# sortperm(c) builds the identity permutation and passes initialized=true (sorting.jl:1063-1065);
# the library generates the index payload itself, so hook sortperm directly.
function Base.sortperm(c::CuVector{T}; kwargs...) where {T<:SortTypes}
    ix = CuVector{Int}(undef, length(c))
    get(kwargs, :initialized, false) && return invoke(sortperm, Tuple{CUDACore.AnyCuVector}, c; kwargs...)
    return sortperm!(ix, c; kwargs...)
end

# partialsort!(c, k, ::BitonicSortAlg) = sort! + copy(c[k]) (sorting.jl:1015-1020) reaches the
# sort! method above unchanged (it must leave `c` sorted); nothing to override.
#
# Base.partialsort(c::AnyCuArray, k; kw...) = partialsort!(copy(c), k; kw...) (sorting.jl:1046-1048): for a scalar k
# neither the copy nor the sort is needed — b200_select_kth (MSD radix select: sizeof(T) streaming reads of c, c
# untouched) returns the element a stable sort would put at position k.  Ranges, custom lt/by and alg=QuickSort keep the
# reference path.
function Base.partialsort(c::CuVector{T}, k::Integer; alg::CUDACore.SortingAlgorithm=CUDACore.BitonicSort,
                          lt=isless, by=identity, rev::Bool=false) where {T<:SortTypes}
    if !enabled() || !graftable(lt, by, alg)
        return invoke(partialsort, Tuple{CUDACore.AnyCuArray,Union{Integer,OrdinalRange}}, c, k; alg, lt, by, rev)
    end
    checkbounds(c, k)
    out = CuVector{T}(undef, 1)
    bytes = b200_select_kth_workspace_bytes(dtype_code(T), length(c))
    with_ws(bytes) do ws, nb
        b200_select_kth(ws, nb, c, dtype_code(T), length(c), Int64(k - firstindex(c)), Int32(rev), out, stream())
    end
    return CUDACore.@allowscalar out[1]
end
