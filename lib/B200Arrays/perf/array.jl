# The reduce / scan / sort / findall groups of the reference's perf/array.jl (same shapes, same group names:
# perf/array.jl:5-22,37-41,48-51,72-125,149-153), run twice — graft enabled and disabled — so that
# perf/runbenchmarks.jl-style JSON can be compared A/B on one machine.
#   julia --project lib/B200Arrays/perf/array.jl
# (SURVEY §8 f4.  Never executed in this repository: no Julia toolchain in the build image; the Python
# counterpart tools/perf_array.py times the same shapes through the same C ABI.)
using BenchmarkTools, CUDA, B200Arrays, StableRNGs
rng = StableRNG(123)

macro async_benchmarkable(ex...)
    quote
        @benchmarkable CUDA.@sync blocking=true $(ex...)
    end
end

const m, n = 512, 1000
const m_long, n_long = 3, 1_000_000
gpu_mat = CuArray{Float32}(rand(rng, Float32, m, n))
gpu_mat_long = CuArray{Float32}(rand(rng, Float32, m_long, n_long))
gpu_vec = reshape(gpu_mat, length(gpu_mat))
gpu_mat_ints = CuArray(rand(rng, -10:10, m, n))
gpu_mat_long_ints = CuArray(rand(rng, -10:10, m_long, n_long))
gpu_vec_ints = reshape(gpu_mat_ints, length(gpu_mat_ints))
gpu_vec_bools = CuArray(rand(rng, Bool, m * n))

function suite()
    s = BenchmarkGroup()
    s["findall"]["bool"] = @benchmarkable findall($gpu_vec_bools)
    s["logical"] = @benchmarkable $gpu_vec[$gpu_vec_bools]
    s["findmin"]["1d"] = @async_benchmarkable findmin($gpu_vec)
    s["findmin"]["2d"] = @async_benchmarkable findmin($gpu_mat; dims=1)
    for (name, vec, mat, long) in (("Float32", gpu_vec, gpu_mat, gpu_mat_long),
                                   ("Int64", gpu_vec_ints, gpu_mat_ints, gpu_mat_long_ints))
        a = s["accumulate"][name]
        a["1d"] = @async_benchmarkable accumulate(+, $vec)
        a["dims=1"] = @async_benchmarkable accumulate(+, $mat; dims=1)
        a["dims=2"] = @async_benchmarkable accumulate(+, $mat; dims=2)
        a["dims=1L"] = @async_benchmarkable accumulate(+, $long; dims=1)
        a["dims=2L"] = @async_benchmarkable accumulate(+, $long; dims=2)
        r = s["reductions"]["reduce"][name]
        r["1d"] = @async_benchmarkable reduce(+, $vec)
        r["dims=1"] = @async_benchmarkable reduce(+, $mat; dims=1)
        r["dims=2"] = @async_benchmarkable reduce(+, $mat; dims=2)
        r["dims=1L"] = @async_benchmarkable reduce(+, $long; dims=1)
        r["dims=2L"] = @async_benchmarkable reduce(+, $long; dims=2)
    end
    s["sorting"]["1d"] = @async_benchmarkable sort($gpu_vec)
    s["sorting"]["2d"] = @async_benchmarkable sort($gpu_mat; dims=1)
    s["sorting"]["sortperm"] = @async_benchmarkable sortperm($gpu_vec)
    return s
end

results = Dict{String,Any}()
for (label, flag) in (("graft", true), ("reference", false))
    B200Arrays.enable!(flag)
    SUITE = suite()
    warmup(SUITE; verbose=false)
    results[label] = minimum(run(SUITE; verbose=false))       # min-of-samples, like perf/runbenchmarks.jl:55-61
end
B200Arrays.enable!(true)
for (k, g) in leaves(results["graft"])
    r = results["reference"][k]
    println(rpad(join(k, "/"), 40), lpad(BenchmarkTools.prettytime(time(g)), 12), "  (reference ",
            BenchmarkTools.prettytime(time(r)), ", ", round(time(r) / time(g); digits=2), "x)")
end
