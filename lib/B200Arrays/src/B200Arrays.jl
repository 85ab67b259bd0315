# B200Arrays — Julia side of the graft: ccall layer over libb200arrays.so plus
# concrete-signature methods on CuArray for reduce / scan / sort.
#
# Layout follows the reference's own library packages (lib/curand/src/cuRAND.jl:1-119):
#   libb200arrays.jl  raw @ccall wrappers, one per symbol of include/b200arrays.h
#   error.jl          status -> exception (mirror of lib/curand/src/libcurand.jl:10-32)
#   mapreduce.jl      GPUArrays.mapreducedim! methods   (replaces CUDACore/src/mapreduce.jl:168)
#   accumulate.jl     CUDACore.scan! methods            (replaces CUDACore/src/accumulate.jl:135)
#   sorting.jl        Base.sort!/sortperm!/partialsort! (replaces CUDACore/src/sorting.jl:1003-1070)
#   indexing.jl       Base.findall(::CuArray{Bool})     (replaces CUDACore/src/indexing.jl:26-66)
#   multigpu.jl       sharded_sort! over the devices of one process (b200_mgpu_*; no reference counterpart)
#
# NOTE: no Julia toolchain exists in the build image, so this package has never been
# executed there; the Python mirror in cuda.jl_b200/ drives the same C ABI in the tests.
module B200Arrays

using CUDACore
using CUDACore: CuArray, CuVector, CuPtr, CuStream, stream, device, capability, initialize_context,
                unsafe_free!, @gcsafe_ccall
using GPUArrays
using BFloat16s: BFloat16
using Preferences

const libb200arrays = Ref{String}("")
const _enabled = Ref(false)

"Is the graft active for the current process? (`B200Arrays.enable!(false)` falls back to the reference kernels for A/B timing.)"
enabled() = _enabled[]
enable!(flag::Bool=true) = (_enabled[] = flag && !isempty(libb200arrays[]); _enabled[])

include("error.jl")
include("libb200arrays.jl")
include("mapreduce.jl")
include("accumulate.jl")
include("sorting.jl")
include("indexing.jl")
include("multigpu.jl")

function __init__()
    CUDACore.functional() || return
    path = @load_preference("libb200arrays", get(ENV, "B200ARRAYS_LIB", ""))
    if isempty(path) || !isfile(path)
        @debug "B200Arrays: libb200arrays.so not found (set the `libb200arrays` preference or B200ARRAYS_LIB); graft disabled"
        return
    end
    libb200arrays[] = path
    # the library is sm_100a-only: leave the reference methods in charge elsewhere
    if capability(device()) < v"10.0" || capability(device()) >= v"11.0"
        @debug "B200Arrays: device is not compute capability 10.x; graft disabled"
        return
    end
    _enabled[] = @load_preference("enabled", true)
    return
end

end # module
