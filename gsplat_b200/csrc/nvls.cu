// nvls.cu -- gradient all-reduce over NVSwitch multicast memory (NVLS), the collective of the view-parallel
// train step (one SUM all-reduce of the 59 floats / gaussian).
//
// Every rank keeps its gradients in a SYMMETRIC buffer (same size and layout on every GPU, allocated and
// exchanged by the host plumbing, bound to one multicast address `mc`).  One launch per rank:
//   1. block barrier across the ranks (every rank's backward has finished writing its buffer);
//   2. rank r owns the r-th 1/world slice: `multimem.ld_reduce.add` on the multicast address makes the SWITCH
//      fetch that 16-byte chunk from all GPUs and return their sum; `multimem.st` broadcasts the sum back into
//      every GPU's buffer.  Per GPU: ~1x the payload up, ~1x down, no intermediate copies, no staging;
//   3. block barrier (all slices are written everywhere before anybody's next kernel reads or overwrites).
// Barriers are flag exchanges through per-rank signal pads in peer memory: block b of rank r raises
// pad[peer][b * world + r] on every peer and waits for pad[r][b * world + peer]; compare-and-swap in both
// directions leaves the pads zero again, so no epoch counter is needed.
#include "common.cuh"

namespace gsb
{
enum class Order
{
    Relaxed,
    AcqRel
};

template<Order O>
__device__ __forceinline__ uint32_t cas_sys(uint32_t *addr, uint32_t expect, uint32_t desired, bool acquire_side)
{
    uint32_t old;
    if constexpr(O == Order::Relaxed)
        asm volatile("atom.global.relaxed.sys.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(addr), "r"(expect), "r"(desired) : "memory");
    else if(acquire_side)
        asm volatile("atom.global.acquire.sys.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(addr), "r"(expect), "r"(desired) : "memory");
    else
        asm volatile("atom.global.release.sys.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(addr), "r"(expect), "r"(desired) : "memory");
    return old;
}

// All threads of the block call this; threads [0, world) each handle one peer.
template<Order O>
__device__ __forceinline__ void rank_barrier(uint32_t *const *pads, int rank, int world)
{
    __syncthreads();
    if((int)threadIdx.x < world)
    {
        const int peer   = (int)threadIdx.x;
        uint32_t *theirs = pads[peer] + (size_t)blockIdx.x * world + rank;
        uint32_t *mine   = pads[rank] + (size_t)blockIdx.x * world + peer;
        while(cas_sys<O>(theirs, 0u, 1u, false) != 0u) // raise my flag at the peer (waits until the last one was consumed)
            ;
        while(cas_sys<O>(mine, 1u, 0u, true) != 1u) // consume the peer's flag
            ;
    }
    __syncthreads();
}

__device__ __forceinline__ float4 multimem_ld_reduce_add(const float4 *mc)
{
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(mc)
                 : "memory");
    return v;
}

__device__ __forceinline__ void multimem_st(float4 *mc, const float4 v)
{
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

constexpr int kNvlsThreads = 512, kNvlsUnroll = 4;

__global__ void __launch_bounds__(kNvlsThreads) nvls_allreduce_kernel(
    float4 *__restrict__ mc, int64_t n_vec, int rank, int world, uint32_t *const *__restrict__ pads
)
{
    rank_barrier<Order::Relaxed>(pads, rank, world); // stream order already made this rank's own writes visible
    const int64_t per = (n_vec + world - 1) / world;
    const int64_t lo = (int64_t)rank * per, hi = (lo + per < n_vec) ? lo + per : n_vec;
    const int64_t stride = (int64_t)gridDim.x * kNvlsThreads;
    for(int64_t i = lo + (int64_t)blockIdx.x * kNvlsThreads + threadIdx.x; i < hi; i += stride * kNvlsUnroll)
    {
        float4 v[kNvlsUnroll];
#pragma unroll
        for(int u = 0; u < kNvlsUnroll; ++u)
            if(i + u * stride < hi)
                v[u] = multimem_ld_reduce_add(mc + i + u * stride);
#pragma unroll
        for(int u = 0; u < kNvlsUnroll; ++u)
            if(i + u * stride < hi)
                multimem_st(mc + i + u * stride, v[u]);
    }
    __threadfence_system();
    rank_barrier<Order::AcqRel>(pads, rank, world);
}
// Two-rank variant without the switch reduction: rank r pulls the peer copies of ITS slice with plain peer
// loads, adds, and pushes the sum into every buffer.  Per direction (world - 1) / world * 2 payloads -- less
// than the multicast scheme's (1 + 1 / world) only for world == 2, which is when the host picks it.
__global__ void __launch_bounds__(kNvlsThreads) p2p_allreduce_kernel(
    float4 *const *__restrict__ bufs, int64_t n_vec, int rank, int world, uint32_t *const *__restrict__ pads
)
{
    rank_barrier<Order::Relaxed>(pads, rank, world);
    const int64_t per = (n_vec + world - 1) / world;
    const int64_t lo = (int64_t)rank * per, hi = (lo + per < n_vec) ? lo + per : n_vec;
    const int64_t stride = (int64_t)gridDim.x * kNvlsThreads;
    float4 *mine         = bufs[rank];
    for(int64_t i = lo + (int64_t)blockIdx.x * kNvlsThreads + threadIdx.x; i < hi; i += stride * kNvlsUnroll)
    {
        float4 v[kNvlsUnroll];
#pragma unroll
        for(int u = 0; u < kNvlsUnroll; ++u)
            if(i + u * stride < hi)
                v[u] = mine[i + u * stride];
        for(int p = 1; p < world; ++p)
        {
            const float4 *peer = bufs[(rank + p) % world];
            float4 w[kNvlsUnroll];
#pragma unroll
            for(int u = 0; u < kNvlsUnroll; ++u)
                if(i + u * stride < hi)
                    w[u] = __ldcv(peer + i + u * stride); // peer memory: never from a stale cache line
#pragma unroll
            for(int u = 0; u < kNvlsUnroll; ++u)
                v[u].x += w[u].x, v[u].y += w[u].y, v[u].z += w[u].z, v[u].w += w[u].w;
        }
        for(int p = 0; p < world; ++p)
        {
            float4 *dst = bufs[(rank + p) % world];
#pragma unroll
            for(int u = 0; u < kNvlsUnroll; ++u)
                if(i + u * stride < hi)
                    dst[i + u * stride] = v[u];
        }
    }
    __threadfence_system();
    rank_barrier<Order::AcqRel>(pads, rank, world);
}
} // namespace gsb

extern "C" int gsb200_p2p_allreduce_f32(
    void *const *buffers_dev, int64_t n_floats, int rank, int world, void *const *signal_pads_dev, int64_t signal_pad_bytes,
    int blocks, void *stream
)
{
    if(n_floats < 0 || world < 1 || rank < 0 || rank >= world || blocks < 1 || world > gsb::kNvlsThreads)
        return GSB200_E_INVALID;
    if(n_floats == 0 || world == 1)
        return GSB200_OK;
    if(!buffers_dev || !signal_pads_dev || (n_floats & 3))
        return GSB200_E_INVALID;
    if((int64_t)blocks * world * (int64_t)sizeof(uint32_t) > signal_pad_bytes)
        return GSB200_E_WORKSPACE;
    gsb::p2p_allreduce_kernel<<<blocks, gsb::kNvlsThreads, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<float4 *const *>(buffers_dev), n_floats / 4, rank, world, reinterpret_cast<uint32_t *const *>(signal_pads_dev)
    );
    return gsb::check_launch();
}

extern "C" int gsb200_nvls_allreduce_f32(
    void *multicast_ptr, int64_t n_floats, int rank, int world, void *const *signal_pads_dev, int64_t signal_pad_bytes,
    int blocks, void *stream
)
{
    if(n_floats < 0 || world < 1 || rank < 0 || rank >= world || blocks < 1 || world > gsb::kNvlsThreads)
        return GSB200_E_INVALID;
    if(n_floats == 0 || world == 1)
        return GSB200_OK;
    if(!multicast_ptr || !signal_pads_dev || (n_floats & 3) || (reinterpret_cast<uintptr_t>(multicast_ptr) & 15))
        return GSB200_E_INVALID;
    if((int64_t)blocks * world * (int64_t)sizeof(uint32_t) > signal_pad_bytes)
        return GSB200_E_WORKSPACE;
    gsb::nvls_allreduce_kernel<<<blocks, gsb::kNvlsThreads, 0, (cudaStream_t)stream>>>(
        static_cast<float4 *>(multicast_ptr), n_floats / 4, rank, world, reinterpret_cast<uint32_t *const *>(signal_pads_dev)
    );
    return gsb::check_launch();
}
