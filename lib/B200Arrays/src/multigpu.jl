# multigpu.jl — sharded keys-only sort over the devices of ONE Julia process (SURVEY §8e).
#
# The reference has no multi-GPU reduce / scan / sort; its building blocks are `device!` per task
# (CUDACore/lib/cudadrv/state.jl:225-252) and peer access between contexts (lib/cudadrv/context.jl:372-405).
# In a single process every device pointer lives in one unified address space, so once peer access is enabled a
# `CuPtr` of device B can be handed to a kernel running on device A: the fused partition + exchange kernel
# (`b200_mgpu_scatter`) then stores bucket b straight into device b's receive buffer over NVLink.  (The Python mirror
# does the same across PROCESSES with torch's symmetric memory; `cuda.jl_b200/multigpu.py`.)
#
#   sharded_sort!(shards; rev=false) -> Vector{CuVector{T}}   (result slice r lives on device(shards[r]))
#
# Flow (include/b200arrays.h, "device-driven multi-GPU sort"):
#   b200_mgpu_hist16 on every device -> gather of the G x 65536 histograms onto every device (256 KB each)
#   -> b200_mgpu_plan (identical decisions everywhere) -> 1 read-back of recv_total / max_recv_total
#   -> b200_mgpu_scatter (one pass, peer stores) -> b200_sort_keys_from (local hybrid MSD sort on the plan's histogram)
#
# NOTE: like the rest of this package, never executed in the build image (no Julia toolchain).

const MgpuTypes = Union{UInt32,Int32,Float32}
const MGPU_MAX_RANKS = 16
const MSD_BINS = 65536

# mirror of b200_mgpu_plan_t (include/b200arrays.h)
struct MgpuPlan
    spl::NTuple{17,UInt32}
    ticket::UInt32
    max_shard_len::UInt32
    pad::UInt32
    dst_base::NTuple{16,UInt64}
    send_count::NTuple{16,UInt64}
    recv_count::NTuple{16,UInt64}
    recv_total::UInt64
    max_recv_total::UInt64
    total::UInt64
end

function b200_mgpu_plan_bytes()
    bytes = Ref{Csize_t}(0)
    check(@ccall libb200arrays[].b200_mgpu_plan_bytes(bytes::Ref{Csize_t})::Int32)
    return Int(bytes[])
end

function b200_mgpu_hist16(keys, dkey, n, rev, nan_canonical, hist16, s::CuStream)
    initialize_context()
    check(@gcsafe_ccall libb200arrays[].b200_mgpu_hist16(keys::CuPtr{Cvoid}, dkey::Int32, n::Int64, rev::Int32,
              nan_canonical::Int32, hist16::CuPtr{Cvoid}, s::CUDACore.CUstream)::Int32)
end

function b200_mgpu_plan(hist_all, nranks, rank, plan, hist_recv, s::CuStream)
    initialize_context()
    check(@gcsafe_ccall libb200arrays[].b200_mgpu_plan(hist_all::CuPtr{Cvoid}, nranks::Int32, rank::Int32,
              plan::CuPtr{Cvoid}, hist_recv::CuPtr{Cvoid}, s::CUDACore.CUstream)::Int32)
end

function b200_mgpu_scatter_workspace_bytes(n)
    bytes = Ref{Csize_t}(0)
    check(@ccall libb200arrays[].b200_mgpu_scatter_workspace_bytes(n::Int64, bytes::Ref{Csize_t})::Int32)
    return Int(bytes[])
end

function b200_mgpu_scatter(ws, ws_bytes, keys, dkey, n, rev, plan, dst::Vector{CuPtr{Cvoid}}, s::CuStream)
    initialize_context()
    check(@gcsafe_ccall libb200arrays[].b200_mgpu_scatter(ws::CuPtr{Cvoid}, ws_bytes::Csize_t, keys::CuPtr{Cvoid},
              dkey::Int32, n::Int64, rev::Int32, plan::CuPtr{Cvoid}, dst::Ptr{CuPtr{Cvoid}}, length(dst)::Int32,
              s::CUDACore.CUstream)::Int32)
end

function b200_sort_keys_from(ws, ws_bytes, src, dst, dkey, n, rev, hist16, s::CuStream)
    initialize_context()
    check(@gcsafe_ccall libb200arrays[].b200_sort_keys_from(ws::CuPtr{Cvoid}, ws_bytes::Csize_t, src::CuPtr{Cvoid},
              dst::CuPtr{Cvoid}, dkey::Int32, n::Int64, rev::Int32, hist16::CuPtr{Cvoid}, s::CUDACore.CUstream)::Int32)
end

"""
    sharded_sort!(shards::Vector{<:CuVector{T}}; rev=false) where T<:Union{UInt32,Int32,Float32}

Sort the concatenation of `shards` (shard `r` resident on its own device, at most 2^30 keys each, at most 16 shards).
Returns one `CuVector{T}` per device; their concatenation in shard order is the sorted sequence (`isless` order, equal
keys never split between two devices).  One task per device, like the reference's multi-device examples.
"""
function sharded_sort!(shards::Vector{<:CuVector{T}}; rev::Bool=false) where {T<:MgpuTypes}
    G = length(shards)
    1 <= G <= MGPU_MAX_RANKS || throw(ArgumentError("sharded_sort!: 1 to $MGPU_MAX_RANKS shards"))
    enabled() || error("B200Arrays is not enabled on this device")
    devs = [device(s) for s in shards]
    dk = dtype_code(T)
    # peer access in both directions (CUDACore/lib/cudadrv/context.jl:372-405); throws if the topology has no P2P path
    for a in devs, b in devs
        a == b && continue
        CUDACore.device!(a) do
            CUDACore.can_access_peer(a, b) || error("no peer access from $a to $b")
            CUDACore.enable_peer_access(CUDACore.context(b))
        end
    end
    # 1. every device: histogram of the top 16 key bits of its shard
    hists = Vector{CuVector{UInt32}}(undef, G)
    @sync for r in 1:G
        @async CUDACore.device!(devs[r]) do
            hists[r] = CuVector{UInt32}(undef, MSD_BINS)
            b200_mgpu_hist16(shards[r], dk, length(shards[r]), Int32(rev), Int32(0), hists[r], stream())
            CUDACore.synchronize()
        end
    end
    # 2. every device: the G histograms side by side (peer copies of 256 KB; NCCL.Allgather! does the same across processes)
    plans = Vector{MgpuPlan}(undef, G)
    plan_bufs = Vector{CuVector{UInt8}}(undef, G)
    hist_recv = Vector{CuVector{UInt32}}(undef, G)
    @sync for r in 1:G
        @async CUDACore.device!(devs[r]) do
            hist_all = CuVector{UInt32}(undef, G * MSD_BINS)
            for s in 1:G
                copyto!(hist_all, (s - 1) * MSD_BINS + 1, hists[s], 1, MSD_BINS)
            end
            # 3. the plan: splitters, count matrix, this rank's offsets inside every receive buffer
            plan_bufs[r] = CuVector{UInt8}(undef, b200_mgpu_plan_bytes())
            hist_recv[r] = CuVector{UInt32}(undef, MSD_BINS)
            b200_mgpu_plan(hist_all, Int32(G), Int32(r - 1), plan_bufs[r], hist_recv[r], stream())
            host = Array(view(plan_bufs[r], 1:sizeof(MgpuPlan)))          # the one read-back: slice lengths
            plans[r] = unsafe_load(Ptr{MgpuPlan}(pointer(host)))
            unsafe_free!(hist_all)
        end
    end
    cap = max(Int(plans[1].max_recv_total), 1)
    # 4. receive buffers, then the fused partition + exchange: bucket b goes straight to device b (NVLink stores)
    recv = Vector{CuVector{T}}(undef, G)
    @sync for r in 1:G
        @async CUDACore.device!(devs[r]) do
            recv[r] = CuVector{T}(undef, cap)
            CUDACore.synchronize()
        end
    end
    dst = CuPtr{Cvoid}[reinterpret(CuPtr{Cvoid}, pointer(recv[b])) for b in 1:G]
    @sync for r in 1:G
        @async CUDACore.device!(devs[r]) do
            n = length(shards[r])
            bytes = b200_mgpu_scatter_workspace_bytes(n)
            with_ws(bytes) do ws, _
                b200_mgpu_scatter(ws, bytes, shards[r], dk, n, Int32(rev), plan_bufs[r], dst, stream())
            end
            CUDACore.synchronize()                                   # every source has finished writing before anyone sorts
        end
    end
    # 5. local hybrid MSD sort of what arrived, on the histogram the plan already holds
    out = Vector{CuVector{T}}(undef, G)
    @sync for r in 1:G
        @async CUDACore.device!(devs[r]) do
            n = Int(plans[r].recv_total)
            out[r] = CuVector{T}(undef, n)
            if n > 0
                bytes = b200_sort_workspace_bytes(dk, Int32(-1), n, 1)
                with_ws(bytes) do ws, _
                    b200_sort_keys_from(ws, bytes, recv[r], out[r], dk, n, Int32(rev), hist_recv[r], stream())
                end
            end
            unsafe_free!(recv[r]); unsafe_free!(plan_bufs[r]); unsafe_free!(hist_recv[r]); unsafe_free!(hists[r])
        end
    end
    return out
end
