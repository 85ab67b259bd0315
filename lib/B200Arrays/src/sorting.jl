# Base.sort! / sortperm! / partialsort! for dense CuVectors (and dims=1 of matrices) with
# lt=isless, by=identity — replaces the bitonic path of CUDACore/src/sorting.jl:1003-1070.
# kwargs do not dispatch, so each hook guards at run time and `invoke`s the reference
# method otherwise (custom lt/by, alg=QuickSort, user-supplied initialized=true, other dims).

graftable(lt, by, alg) = lt === isless && by === identity && alg === CUDACore.BitonicSort

function Base.sort!(c::CuArray{T}; alg::CUDACore.SortingAlgorithm=CUDACore.BitonicSort, lt=isless, by=identity,
                    rev::Bool=false, dims::Integer=1) where {T<:SortTypes}
    if !enabled() || !graftable(lt, by, alg) || dims != 1
        return invoke(sort!, Tuple{CUDACore.AnyCuArray}, c; alg, lt, by, rev, dims)
    end
    n = size(c, 1); nseg = n == 0 ? 0 : length(c) ÷ n
    (n <= 1 || nseg == 0) && return c
    bytes = b200_sort_workspace_bytes(dtype_code(T), Int32(-1), n, nseg)
    with_ws(bytes) do ws, nb
        b200_sort_keys(ws, nb, c, dtype_code(T), n, nseg, Int32(rev), stream())
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
    if !enabled() || !graftable(lt, by, alg) || initialized || dims != 1
        return invoke(sortperm!, Tuple{CUDACore.AnyCuArray,CUDACore.AnyCuArray}, ix, A; initialized, alg, lt, by, rev, dims)
    end
    n = size(A, 1); nseg = length(A) ÷ n
    bytes = b200_sort_workspace_bytes(dtype_code(T), SORTPERM_INDEX, n, nseg)
    with_ws(bytes) do ws, nb
        b200_sortperm(ws, nb, A, ix, dtype_code(T), dtype_code(I), n, nseg, Int32(rev), Int32(ndims(A) > 1), stream())
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
# sort! method above unchanged; nothing to override.
