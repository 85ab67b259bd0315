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
end
