// loss.cu -- fused photometric L1 loss of the train step (caller-side fusion).
// The trainer's `(render - target).abs().mean()` is seven ATen launches forward + backward over the 25 MB image;
// here: one pass for the loss (deterministic two-level sum) and one pass that writes d loss / d render.
#include "common.cuh"

namespace gsb
{
constexpr int kLossThreads = 256, kLossBlocks = 592; // 4 CTAs per SM

__global__ void __launch_bounds__(kLossThreads) l1_partial_kernel(
    int64_t n, const float *__restrict__ a, const float *__restrict__ b, float *__restrict__ partial
)
{
    __shared__ float warp_sum[kLossThreads / 32];
    float acc            = 0.f;
    const int64_t n4     = n >> 2;
    const float4 *a4     = reinterpret_cast<const float4 *>(a);
    const float4 *b4     = reinterpret_cast<const float4 *>(b);
    const int64_t stride = (int64_t)gridDim.x * kLossThreads;
    for(int64_t i = (int64_t)blockIdx.x * kLossThreads + threadIdx.x; i < n4; i += stride)
    {
        const float4 x = a4[i], y = b4[i];
        acc += (fabsf(x.x - y.x) + fabsf(x.y - y.y)) + (fabsf(x.z - y.z) + fabsf(x.w - y.w));
    }
    if(blockIdx.x == 0 && threadIdx.x < (n & 3)) // tail
        acc += fabsf(a[(n4 << 2) + threadIdx.x] - b[(n4 << 2) + threadIdx.x]);
#pragma unroll
    for(int o = 16; o > 0; o >>= 1)
        acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if((threadIdx.x & 31) == 0)
        warp_sum[threadIdx.x >> 5] = acc;
    __syncthreads();
    if(threadIdx.x == 0)
    {
        float s = 0.f;
#pragma unroll
        for(int w = 0; w < kLossThreads / 32; ++w)
            s += warp_sum[w];
        partial[blockIdx.x] = s;
    }
}

// one warp: fixed-order sum of the block partials -> mean
__global__ void l1_final_kernel(int blocks, const float *__restrict__ partial, float inv_n, float *__restrict__ loss)
{
    float acc = 0.f;
    for(int i = threadIdx.x; i < blocks; i += 32)
        acc += partial[i];
#pragma unroll
    for(int o = 16; o > 0; o >>= 1)
        acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if(threadIdx.x == 0)
        *loss = acc * inv_n;
}

// v_a = (*v_loss / n) * sign(a - b)   (sign(0) = 0, the subgradient torch.abs uses)
__global__ void __launch_bounds__(kLossThreads) l1_bwd_kernel(
    int64_t n, const float *__restrict__ a, const float *__restrict__ b, const float *__restrict__ v_loss, float inv_n,
    float *__restrict__ v_a
)
{
    const float g        = *v_loss * inv_n;
    const int64_t n4     = n >> 2;
    const float4 *a4     = reinterpret_cast<const float4 *>(a);
    const float4 *b4     = reinterpret_cast<const float4 *>(b);
    float4 *o4           = reinterpret_cast<float4 *>(v_a);
    const int64_t stride = (int64_t)gridDim.x * kLossThreads;
    auto sg = [g](float d) { return d > 0.f ? g : (d < 0.f ? -g : 0.f); };
    for(int64_t i = (int64_t)blockIdx.x * kLossThreads + threadIdx.x; i < n4; i += stride)
    {
        const float4 x = a4[i], y = b4[i];
        o4[i] = make_float4(sg(x.x - y.x), sg(x.y - y.y), sg(x.z - y.z), sg(x.w - y.w));
    }
    if(blockIdx.x == 0 && threadIdx.x < (n & 3))
    {
        const int64_t t = (n4 << 2) + threadIdx.x;
        v_a[t]          = sg(a[t] - b[t]);
    }
}
} // namespace gsb

extern "C" size_t gsb200_l1_loss_workspace_bytes(void) { return sizeof(float) * gsb::kLossBlocks; }

// loss[0] = mean |a - b| over n elements (a, b 16-byte aligned).  workspace: gsb200_l1_loss_workspace_bytes().
extern "C" int gsb200_l1_loss_fwd(int64_t n, const float *a, const float *b, float *loss, void *workspace, void *stream)
{
    if(n <= 0 || !a || !b || !loss || !workspace || ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15))
        return GSB200_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    float *partial  = static_cast<float *>(workspace);
    gsb::l1_partial_kernel<<<gsb::kLossBlocks, gsb::kLossThreads, 0, st>>>(n, a, b, partial);
    if(int rc = gsb::check_launch())
        return rc;
    gsb::l1_final_kernel<<<1, 32, 0, st>>>(gsb::kLossBlocks, partial, 1.0f / (float)n, loss);
    return gsb::check_launch();
}

// v_a[i] = v_loss[0] / n * sign(a[i] - b[i]);  v_loss is a DEVICE scalar (the incoming gradient of the loss).
extern "C" int gsb200_l1_loss_bwd(int64_t n, const float *a, const float *b, const float *v_loss, float *v_a, void *stream)
{
    if(n <= 0 || !a || !b || !v_loss || !v_a
       || ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(v_a)) & 15))
        return GSB200_E_INVALID;
    gsb::l1_bwd_kernel<<<gsb::kLossBlocks, gsb::kLossThreads, 0, (cudaStream_t)stream>>>(n, a, b, v_loss, 1.0f / (float)n, v_a);
    return gsb::check_launch();
}
