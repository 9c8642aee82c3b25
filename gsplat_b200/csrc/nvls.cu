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

// ---------------------------------------------------------------------------------------------
// Row-sparse variants.  In the view-parallel step most gaussians are culled in a rank's view (71 % at BASELINE
// configs[3]): their gradient rows are zeros on that rank, and a row that NO rank touched needs no traffic at all.
// The fused backward (project_sh_bwd) publishes one bit per gaussian ("seen in some view of this rank") in the
// symmetric buffer; here a warp takes a group of 32 rows, ORs the ranks' bitmap words (in the switch:
// multimem.ld_reduce.or, or one peer load at two ranks) and moves only the rows whose bit is set.  The buffer is
// a structure of arrays: segment k holds N rows of w_k floats (means 3, quats 4, scales 3, opacities 1, SH 48);
// a group's slice of a segment is 8 w_k float4, always 16-byte aligned.  Exact: skipped rows are zero everywhere.
constexpr int kMaxSegs = 8;
struct RowSegs
{
    int n;
    int64_t vec_off[kMaxSegs]; // first float4 of the segment
    int32_t width[kMaxSegs];   // floats per row
    uint32_t magic_vpr[kMaxSegs]; // width % 4 == 0: floor(2^32 / (width / 4)) + 1 (0 for one vector per row): n / vpr by __umulhi
    int64_t n_vec[kMaxSegs];   // float4 in the segment (rows * width / 4, rounded up)
};

__device__ __forceinline__ uint32_t multimem_ld_reduce_or(const uint32_t *mc)
{
    uint32_t v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.or.b32 %0, [%1];" : "=r"(v) : "l"(mc) : "memory");
    return v;
}

constexpr int kRowsUnroll = 8; // 16-byte vectors in flight per lane and round

// One warp per group of 32 rows; groups are dealt round-robin over the ranks (visibility is spatially correlated with
// the row index, contiguous halves would leave one rank with most of the touched rows) and grid-stride over the warps.
// Inside a group the TOUCHED work is flattened -- for segments whose rows are whole vectors (quats: 1, SH: 12 per row)
// only the vectors of touched rows, found with __fns on the ORed bitmap word; the narrow segments (means, scales,
// opacities: 56 vectors per group together) as a whole -- so that every lane has a vector in every round: with a third of
// the rows touched, predicating a dense walk left two thirds of the lanes idle in a latency-bound loop (measured at 2
// ranks: 0.62 ms for 35 % of the payload, slower than the dense kernel's 0.38 ms for all of it).
template<bool NVLS>
__global__ void __launch_bounds__(kNvlsThreads) rows_allreduce_kernel(
    float4 *__restrict__ mc, float4 *const *__restrict__ bufs, const RowSegs segs, int64_t n_rows, int64_t bitmap_off_words,
    int rank, int world, uint32_t *const *__restrict__ pads, unsigned long long *__restrict__ stats
)
{
    rank_barrier<Order::Relaxed>(pads, rank, world);
    const int64_t groups = (n_rows + 31) / 32;
    const int lane       = threadIdx.x & 31;
    const int64_t warp0  = (int64_t)blockIdx.x * (kNvlsThreads / 32) + (threadIdx.x >> 5);
    const int64_t nwarp  = (int64_t)gridDim.x * (kNvlsThreads / 32);
    float4 *mine         = NVLS ? nullptr : bufs[rank];
    unsigned long long moved = 0;
    // the warp's groups, 32 at a time: lane l fetches the ORed bitmap word of the l-th of them (one round trip for all)
    const int64_t gstride = (int64_t)world * nwarp;
    for(int64_t gbase = (int64_t)rank + (int64_t)world * warp0; gbase < groups; gbase += 32 * gstride)
    {
        const int64_t myg = gbase + (int64_t)lane * gstride;
        uint32_t myword   = 0u;
        if(myg < groups)
        {
            if constexpr(NVLS)
                myword = multimem_ld_reduce_or(reinterpret_cast<const uint32_t *>(mc) + bitmap_off_words + myg);
            else
                for(int p = 0; p < world; ++p)
                    myword |= __ldcv(reinterpret_cast<const uint32_t *>(bufs[p]) + bitmap_off_words + myg);
        }
        uint32_t todo = __ballot_sync(0xffffffffu, myword != 0u);
        while(todo)
        {
        const int gi = __ffs(todo) - 1;
        todo &= todo - 1;
        const uint32_t word = __shfl_sync(0xffffffffu, myword, gi);
        const int64_t g     = gbase + (int64_t)gi * gstride;
        const int nt = __popc(word);
        // flattened work list of this group: prefix[k] = first item of segment k
        int prefix[kMaxSegs + 1];
        prefix[0] = 0;
#pragma unroll
        for(int k = 0; k < kMaxSegs; ++k)
        {
            int cnt = 0;
            if(k < segs.n)
                cnt = (segs.width[k] & 3) == 0 ? nt * (segs.width[k] >> 2) : 8 * segs.width[k];
            prefix[k + 1] = prefix[k] + cnt;
        }
        const int total = prefix[kMaxSegs];
        for(int j0 = 0; j0 < total; j0 += 32 * kRowsUnroll)
        {
            float4 x[kRowsUnroll];
            int64_t addr[kRowsUnroll];
#pragma unroll
            for(int u = 0; u < kRowsUnroll; ++u)
            {
                const int j = j0 + 32 * u + lane;
                addr[u]     = -1;
                if(j < total)
                {
                    int k = 0;
#pragma unroll
                    for(int q = 1; q < kMaxSegs; ++q)
                        if(q < segs.n && j >= prefix[q])
                            k = q;
                    const int jl = j - prefix[k];
                    const int w  = segs.width[k];
                    int64_t i;
                    if((w & 3) == 0)
                    { // whole-vector rows: the r-th touched row of the group
                        const int vpr = w >> 2;
                        const int r   = segs.magic_vpr[k] ? (int)__umulhi((uint32_t)jl, segs.magic_vpr[k]) : jl; // 0: one vector per row
                        const int row = (int)__fns(word, 0, r + 1);
                        i             = segs.vec_off[k] + (g * 32 + row) * vpr + (jl - r * vpr);
                    }
                    else
                        i = segs.vec_off[k] + g * 8 * (int64_t)w + jl;
                    if(i < segs.vec_off[k] + segs.n_vec[k])
                        addr[u] = i;
                }
                if(addr[u] >= 0)
                {
                    if constexpr(NVLS)
                        x[u] = multimem_ld_reduce_add(mc + addr[u]);
                    else
                    {
                        x[u] = mine[addr[u]];
                        for(int p = 1; p < world; ++p)
                        {
                            const float4 y = __ldcv(bufs[(rank + p) % world] + addr[u]);
                            x[u].x += y.x, x[u].y += y.y, x[u].z += y.z, x[u].w += y.w;
                        }
                    }
                }
            }
#pragma unroll
            for(int u = 0; u < kRowsUnroll; ++u)
                if(addr[u] >= 0)
                {
                    if constexpr(NVLS)
                        multimem_st(mc + addr[u], x[u]);
                    else
                        for(int p = 0; p < world; ++p)
                            bufs[(rank + p) % world][addr[u]] = x[u];
                    ++moved;
                }
        }
        } // groups of this chunk
    }
    if(stats != nullptr && moved != 0)
        atomicAdd(stats, moved); // float4 actually reduced by this rank (diagnostics)
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

// Row-sparse all-reduce of a structure-of-arrays gradient buffer (see rows_allreduce_kernel).  seg_vec_offsets /
// seg_widths describe n_segs segments of n_rows rows; bitmap_offset_floats locates the per-rank "row touched" bitmap
// (ceil(n_rows / 32) 32-bit words) inside the same symmetric buffer.  multicast_ptr != NULL selects the in-switch
// reduction, otherwise peer loads / stores through buffers_dev.  stats (optional, device u64) accumulates the number of
// 16-byte vectors this rank reduced.
extern "C" int gsb200_rows_allreduce_f32(
    void *multicast_ptr, void *const *buffers_dev, int n_segs, const int64_t *seg_offsets_floats, const int32_t *seg_widths,
    int64_t n_rows, int64_t bitmap_offset_floats, int rank, int world, void *const *signal_pads_dev, int64_t signal_pad_bytes,
    int blocks, void *stats, void *stream
)
{
    if(n_rows < 0 || world < 1 || rank < 0 || rank >= world || blocks < 1 || world > gsb::kNvlsThreads || n_segs < 1
       || n_segs > gsb::kMaxSegs || !seg_offsets_floats || !seg_widths)
        return GSB200_E_INVALID;
    if(n_rows == 0 || world == 1)
        return GSB200_OK;
    if((!multicast_ptr && !buffers_dev) || !signal_pads_dev || bitmap_offset_floats < 0)
        return GSB200_E_INVALID;
    if((int64_t)blocks * world * (int64_t)sizeof(uint32_t) > signal_pad_bytes)
        return GSB200_E_WORKSPACE;
    gsb::RowSegs segs;
    segs.n = n_segs;
    for(int k = 0; k < n_segs; ++k)
    {
        if(seg_widths[k] < 1 || seg_widths[k] > 4096 || (seg_offsets_floats[k] & 3))
            return GSB200_E_INVALID;
        segs.vec_off[k] = seg_offsets_floats[k] / 4;
        segs.width[k]   = seg_widths[k];
        const int vpr    = seg_widths[k] / 4;
        segs.magic_vpr[k] = ((seg_widths[k] & 3) != 0 || vpr <= 1) ? 0u : (uint32_t)(0x100000000ULL / (uint64_t)vpr) + 1u;
        segs.n_vec[k]   = (n_rows * seg_widths[k] + 3) / 4;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if(multicast_ptr)
        gsb::rows_allreduce_kernel<true><<<blocks, gsb::kNvlsThreads, 0, st>>>(
            static_cast<float4 *>(multicast_ptr), nullptr, segs, n_rows, bitmap_offset_floats, rank, world,
            reinterpret_cast<uint32_t *const *>(signal_pads_dev), static_cast<unsigned long long *>(stats)
        );
    else
        gsb::rows_allreduce_kernel<false><<<blocks, gsb::kNvlsThreads, 0, st>>>(
            nullptr, reinterpret_cast<float4 *const *>(buffers_dev), segs, n_rows, bitmap_offset_floats, rank, world,
            reinterpret_cast<uint32_t *const *>(signal_pads_dev), static_cast<unsigned long long *>(stats)
        );
    return gsb::check_launch();
}
