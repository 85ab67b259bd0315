# Runs the reference's own test files for this path against the graft (SURVEY §8 f4):
# with B200Arrays loaded, test/core/sorting.jl and the accumulate / mapreduce testsets of
# test/core/array.jl exercise the new methods unchanged.
using Test, CUDA, B200Arrays
@assert B200Arrays.enabled() "graft disabled: set the libb200arrays preference or B200ARRAYS_LIB"
const CUDA_TEST = joinpath(dirname(dirname(pathof(CUDA))), "test")
include(joinpath(CUDA_TEST, "setup.jl"))
@testset "B200Arrays graft" begin
    include(joinpath(CUDA_TEST, "core", "sorting.jl"))
    include(joinpath(CUDA_TEST, "core", "array.jl"))
    # A/B: same inputs through the reference kernels
    x = CUDA.rand(Float32, 1 << 20)
    s_graft = sum(x); p_graft = Array(sortperm(x)); c_graft = Array(cumsum(x))
    k_graft = [partialsort(x, k) for k in (1, 1 << 19, 1 << 20)]           # radix select, x untouched
    @test k_graft == sort(Array(x))[[1, 1 << 19, 1 << 20]]
    @test partialsort(x, 7; rev=true) == sort(Array(x); rev=true)[7]
    B200Arrays.enable!(false)
    @test sum(x) ≈ s_graft
    @test Array(sortperm(x)) == p_graft
    @test Array(cumsum(x)) ≈ c_graft
    B200Arrays.enable!(true)
    # large keys-only sorts: the hybrid MSD path (n >= 2^27) on uniform and skewed keys, against Base on the host
    for gen in (() -> CUDA.rand(UInt32, 1 << 27), () -> CUDA.rand(Float32, 1 << 27), () -> CUDA.randn(Float32, 1 << 27))
        y = gen()
        @test Array(sort(y)) == sort(Array(y))
        @test partialsort(y, 1 << 26) == sort(Array(y))[1 << 26]          # sampled radix select (n >= 2^22)
    end
    # sharded sort over every visible device (multigpu.jl): concatenated slices == Base's sort of the concatenation
    if length(CUDA.devices()) >= 2
        shards = [CUDA.device!(d) do; CUDA.rand(UInt32, 1 << 22); end for d in CUDA.devices()]
        host = reduce(vcat, Array.(shards))
        out = B200Arrays.sharded_sort!(shards)
        @test reduce(vcat, Array.(out)) == sort(host)
    end
end
