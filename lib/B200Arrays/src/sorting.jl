# Base.sort! / sortperm! / sortperm / partialsort for dense CuArrays along any `dims` with lt=isless,
# by=identity — replaces the bitonic path of CUDACore/src/sorting.jl:1003-1070.
#
# How the reference dispatches (and therefore where the hooks sit):
#   sort!(c; alg=BitonicSort, kw...)            -> sort!(c, alg; kw...)                       (sorting.jl:1007-1009)
#   partialsort!(c, k, ::BitonicSortAlg; ...)   -> sort!(c, alg; lt, by, rev); copy(c[k])     (sorting.jl:1015-1020)
#   sort!(c, ::BitonicSortAlg; kw...)           -> bitonic_sort!(c; kw...)                    (sorting.jl:1003-1005)
# Both public entry points pass `alg` POSITIONALLY, so the graft is ONE more-specific positional method,
# `sort!(::CuArray{T}, ::BitonicSortAlg; kw...)`; `sort!(c)`, `sort(c)`, `partialsort!` and `partialsort(c, range)`
# all reach it unchanged.  QuickSort keeps the reference's `sort!(::AnyCuVector, ::QuickSortAlg; lt, by, rev)`
# (sorting.jl:993) — nothing here touches it.
#
# kwargs do not dispatch, so every hook inspects what the caller passed and otherwise `invoke`s the reference
# method WITH EXACTLY THE CALLER'S kwargs (the reference methods differ in which keywords they accept:
# `bitonic_sort!(c; by, lt, rev, dims)` has no `alg`, sorting.jl:909).

const SORT_KW = (:lt, :by, :rev, :dims)

# The caller's kwargs are graftable when only known keywords appear, lt/by are the defaults, `rev` is a Bool and
# `dims` names a dimension of `c`.
function graftable_sort(c, kwargs)
    enabled() || return false
    all(k -> k in SORT_KW, keys(kwargs)) || return false
    get(kwargs, :lt, isless) === isless || return false
    get(kwargs, :by, identity) === identity || return false
    get(kwargs, :rev, false) isa Bool || return false
    d = get(kwargs, :dims, 1)
    return d isa Integer && 1 <= d <= ndims(c)
end

factor_sort_dims(sz, dims) = (prod(sz[1:dims-1]), sz[dims], prod(sz[dims+1:end]))

function Base.sort!(c::CuArray{T}, alg::CUDACore.BitonicSortAlg; kwargs...) where {T<:SortTypes}
    if !graftable_sort(c, kwargs)
        return invoke(sort!, Tuple{CUDACore.AnyCuArray,CUDACore.BitonicSortAlg}, c, alg; kwargs...)
    end
    rev::Bool = get(kwargs, :rev, false)
    pre, n, post = factor_sort_dims(size(c), Int(get(kwargs, :dims, 1)))
    (n <= 1 || pre * post == 0) && return c
    bytes = b200_sort_dims_workspace_bytes(dtype_code(T), Int32(0), pre, n, post)
    with_ws(bytes) do ws, nb
        b200_sort_keys_dims(ws, nb, c, dtype_code(T), pre, n, post, Int32(rev), stream())
    end
    return c
end

# sortperm!(ix, A; initialized=false, kw...) (sorting.jl:1051-1061).  Grafted for initialized=false: the library
# generates the 1-based (linear, for N-d `A`) index payload itself.  A caller-supplied starting permutation
# (initialized=true) breaks ties by index VALUE (sorting.jl:642-643,732-733) and stays on the reference.
function Base.sortperm!(ix::CuArray{I}, A::CuArray{T}; initialized=false,
                        kwargs...) where {I<:Union{Int32,Int64,UInt32,UInt64},T<:SortTypes}
    if axes(ix) != axes(A)
        throw(ArgumentError("index array must have the same size/axes as the source array, $(axes(ix)) != $(axes(A))"))
    end
    isempty(A) && return ix
    if initialized !== false || !graftable_sort(A, kwargs)
        return invoke(sortperm!, Tuple{CUDACore.AnyCuArray,CUDACore.AnyCuArray}, ix, A; initialized, kwargs...)
    end
    return graft_sortperm!(ix, A, get(kwargs, :rev, false), Int(get(kwargs, :dims, 1)))
end

function graft_sortperm!(ix::CuArray{I}, A::CuArray{T}, rev::Bool, dims::Int) where {I,T}
    pre, n, post = factor_sort_dims(size(A), dims)
    bytes = b200_sort_dims_workspace_bytes(dtype_code(T), Int32(1), pre, n, post)
    with_ws(bytes) do ws, nb
        b200_sortperm_dims(ws, nb, A, ix, dtype_code(T), dtype_code(I), pre, n, post, Int32(rev),
                           Int32(ndims(A) > 1), stream())
    end
    return ix
end

# sortperm(c::AnyCuVector; kw...) = sortperm!(CuArray(1:length(c)), c; initialized=true, kw...) (sorting.jl:1063-1065)
# and sortperm(c::AnyCuArray; dims, kw...) with LinearIndices (sorting.jl:1067-1070): both build the identity
# permutation and pass initialized=true, which the sortperm! hook above must treat as a user permutation — so these
# two are hooked directly.  The result eltype is Int like the reference's `CuArray(1:length(c))`.
function Base.sortperm(c::CuVector{T}; kwargs...) where {T<:SortTypes}
    if !graftable_sort(c, kwargs)
        return invoke(sortperm, Tuple{CUDACore.AnyCuVector}, c; kwargs...)
    end
    isempty(c) && return CuVector{Int}(undef, 0)
    return graft_sortperm!(CuVector{Int}(undef, length(c)), c, get(kwargs, :rev, false), 1)
end

function Base.sortperm(c::CuArray{T}; dims, kwargs...) where {T<:SortTypes}
    if !(dims isa Integer && 1 <= dims <= ndims(c)) || !graftable_sort(c, kwargs)
        return invoke(sortperm, Tuple{CUDACore.AnyCuArray}, c; dims, kwargs...)
    end
    ix = CuArray{Int}(undef, size(c))
    isempty(c) && return ix
    return graft_sortperm!(ix, c, get(kwargs, :rev, false), Int(dims))
end

# partialsort!(c, k; alg=BitonicSort, kw...) -> partialsort!(c, k, alg; kw...) -> sort!(c, alg; lt, by, rev)
# (sorting.jl:1015-1020,1042-1044): reaches the positional sort! method above; nothing to override, and `c` is left
# fully sorted exactly like the reference leaves it.
#
# Base.partialsort(c::AnyCuArray, k; kw...) = partialsort!(copy(c), k; kw...) (sorting.jl:1046-1048): for a scalar k
# neither the copy nor the sort is needed — b200_select_kth (radix select, c untouched) returns the element a stable
# sort would put at position k.  Ranges, custom lt/by and alg=QuickSort keep the reference path.
function Base.partialsort(c::CuVector{T}, k::Integer; alg::CUDACore.SortingAlgorithm=CUDACore.BitonicSort,
                          kwargs...) where {T<:SortTypes}
    if alg !== CUDACore.BitonicSort || haskey(kwargs, :dims) || !graftable_sort(c, kwargs)
        return invoke(partialsort, Tuple{CUDACore.AnyCuArray,Union{Integer,OrdinalRange}}, c, k; alg, kwargs...)
    end
    checkbounds(c, k)
    rev::Bool = get(kwargs, :rev, false)
    out = CuVector{T}(undef, 1)
    bytes = b200_select_kth_workspace_bytes(dtype_code(T), length(c))
    bytes === nothing && return invoke(partialsort, Tuple{CUDACore.AnyCuArray,Union{Integer,OrdinalRange}}, c, k; alg, kwargs...)
    with_ws(bytes) do ws, nb
        b200_select_kth(ws, nb, c, dtype_code(T), length(c), Int64(k - firstindex(c)), Int32(rev), out, stream())
    end
    return CUDACore.@allowscalar out[1]
end
