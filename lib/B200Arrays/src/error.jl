# Status codes of include/b200arrays.h -> Julia exceptions
# (pattern: lib/curand/src/libcurand.jl:10-32, CUDACore/lib/cudadrv/libcuda.jl:25-40).

export B200ArraysError

struct B200ArraysError <: Exception
    code::Int32
    cuda::Int32
end

const B200_OK = Int32(0)
const B200_UNSUPPORTED = Int32(2)
const B200_LAUNCH_FAILURE = Int32(4)

function description(err::B200ArraysError)
    unsafe_string(@ccall libb200arrays[].b200_status_string(err.code::Int32)::Cstring)
end

function Base.showerror(io::IO, err::B200ArraysError)
    print(io, "B200ArraysError: ", description(err), " (code $(err.code))")
    err.cuda != 0 && print(io, " [cudaError $(err.cuda)]")
end

@noinline function throw_api_error(res::Int32)
    cuda = res == B200_LAUNCH_FAILURE ? (@ccall libb200arrays[].b200_last_cuda_error()::Int32) : Int32(0)
    # cudaErrorMemoryAllocation: let CUDA.jl's reclaim ladder have a go (CUDACore/src/memory.jl:505-545)
    cuda == 2 && throw(CUDACore.OutOfGPUMemoryError())
    throw(B200ArraysError(res, cuda))
end

@inline function check(res::Int32)
    res == B200_OK || throw_api_error(res)
    return
end
