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

// ---------------------------------------------------------------------------------------------
// Fused SSIM loss (the D-SSIM term of the reference trainer: gsplat/losses.py:110-201 `torch_ssim_loss` /
// `ssim_loss`, used at examples/simple_trainer.py:951-961).  Semantics of the torch path the reference runs
// without the third-party fused_ssim package: 11x11 Gaussian window (sigma 1.5), zero padding, per channel,
// loss = 1 - mean over B*C*H*W of the SSIM map.  The reference spends 5 depthwise conv2d + ~20 element-wise
// launches forward and as many backward on it; here one pass forward (which also stores the three partial
// derivative maps) and one pass backward.  Images are addressed through (batch, channel, row, col) element
// strides, so the trainer's NHWC render permuted to NCHW is read in place.
namespace gsb
{
constexpr int kSsimT  = 16;            // output tile edge
constexpr int kSsimR  = 5;             // window radius
constexpr int kSsimIn = kSsimT + 2 * kSsimR;
struct Strides4
{
    int64_t b, c, h, w;
};
struct SsimWin
{
    float w[11];
};

__device__ __forceinline__ float ssim_block_sum(float v, float *scratch)
{
#pragma unroll
    for(int o = 16; o > 0; o >>= 1)
        v += __shfl_xor_sync(0xffffffffu, v, o);
    const int t = threadIdx.y * kSsimT + threadIdx.x;
    if((t & 31) == 0)
        scratch[t >> 5] = v;
    __syncthreads();
    float s = 0.f;
    if(t == 0)
    {
#pragma unroll
        for(int w = 0; w < kSsimT * kSsimT / 32; ++w)
            s += scratch[w];
    }
    return s;
}

// grid: (ceil(W/16), ceil(H/16), B*C); block (16,16)
__global__ void __launch_bounds__(kSsimT *kSsimT) ssim_fwd_kernel(
    const int C, const int H, const int W, const float *__restrict__ x, const Strides4 xs, const float *__restrict__ y,
    const Strides4 ys, float *__restrict__ maps, const int64_t plane, float *__restrict__ partial, const SsimWin win
)
{
    __shared__ float sx[kSsimIn][kSsimIn + 1], sy[kSsimIn][kSsimIn + 1];
    __shared__ float hz[5][kSsimIn][kSsimT + 1];
    __shared__ float scratch[kSsimT * kSsimT / 32];
    const int bc = blockIdx.z, b = bc / C, c = bc % C;
    const int x0 = blockIdx.x * kSsimT - kSsimR, y0 = blockIdx.y * kSsimT - kSsimR;
    const int t  = threadIdx.y * kSsimT + threadIdx.x;
    const float *xb = x + b * xs.b + c * xs.c;
    const float *yb = y + b * ys.b + c * ys.c;
    for(int i = t; i < kSsimIn * kSsimIn; i += kSsimT * kSsimT)
    {
        const int r = i / kSsimIn, q = i % kSsimIn;
        const int gy = y0 + r, gx = x0 + q;
        const bool in = gy >= 0 && gy < H && gx >= 0 && gx < W;
        sx[r][q] = in ? xb[gy * xs.h + gx * xs.w] : 0.f;
        sy[r][q] = in ? yb[gy * ys.h + gx * ys.w] : 0.f;
    }
    __syncthreads();
    for(int i = t; i < kSsimIn * kSsimT; i += kSsimT * kSsimT)
    {
        const int r = i / kSsimT, q = i % kSsimT;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f;
#pragma unroll
        for(int k = 0; k < 11; ++k)
        {
            const float w = win.w[k], u = sx[r][q + k], v = sy[r][q + k];
            a0 += w * u, a1 += w * v, a2 += w * u * u, a3 += w * v * v, a4 += w * u * v;
        }
        hz[0][r][q] = a0, hz[1][r][q] = a1, hz[2][r][q] = a2, hz[3][r][q] = a3, hz[4][r][q] = a4;
    }
    __syncthreads();
    float mu1 = 0.f, mu2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
    for(int k = 0; k < 11; ++k)
    {
        const float w = win.w[k];
        mu1 += w * hz[0][threadIdx.y + k][threadIdx.x];
        mu2 += w * hz[1][threadIdx.y + k][threadIdx.x];
        e11 += w * hz[2][threadIdx.y + k][threadIdx.x];
        e22 += w * hz[3][threadIdx.y + k][threadIdx.x];
        e12 += w * hz[4][threadIdx.y + k][threadIdx.x];
    }
    const int gx = blockIdx.x * kSsimT + threadIdx.x, gy = blockIdx.y * kSsimT + threadIdx.y;
    const bool in = gx < W && gy < H;
    constexpr float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
    const float s11 = e11 - mu1 * mu1, s22 = e22 - mu2 * mu2, s12 = e12 - mu1 * mu2;
    const float a = 2.f * mu1 * mu2 + C1, bq = 2.f * s12 + C2, cq = mu1 * mu1 + mu2 * mu2 + C1, d = s11 + s22 + C2;
    const float inv_cd = 1.f / (cq * d);
    const float m      = a * bq * inv_cd;
    if(maps != nullptr && in)
    {
        // dm/d(sigma1^2), dm/d(sigma12), and dm/d(mu1) with the sigma terms' mu1-dependence folded in
        const float dm_s11 = -m / d;
        const float dm_s12 = 2.f * a * inv_cd;
        const float dm_mu1 = 2.f * mu2 * bq * inv_cd - 2.f * mu1 * m / cq - 2.f * mu1 * dm_s11 - mu2 * dm_s12;
        const int64_t o    = (int64_t)bc * H * W + (int64_t)gy * W + gx;
        maps[o] = dm_mu1, maps[plane + o] = dm_s11, maps[2 * plane + o] = dm_s12;
    }
    const float s = ssim_block_sum(in ? m : 0.f, scratch);
    if(t == 0)
        partial[((int64_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = s;
}

// fixed-order sum of the block partials: one CTA, 1024 threads
__global__ void __launch_bounds__(1024) ssim_final_kernel(int64_t n, const float *__restrict__ partial, float inv_count, float *__restrict__ loss)
{
    __shared__ float ws[32];
    float acc = 0.f;
    for(int64_t i = threadIdx.x; i < n; i += 1024)
        acc += partial[i];
#pragma unroll
    for(int o = 16; o > 0; o >>= 1)
        acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if((threadIdx.x & 31) == 0)
        ws[threadIdx.x >> 5] = acc;
    __syncthreads();
    if(threadIdx.x < 32)
    {
        float v = ws[threadIdx.x];
#pragma unroll
        for(int o = 16; o > 0; o >>= 1)
            v += __shfl_xor_sync(0xffffffffu, v, o);
        if(threadIdx.x == 0)
            *loss = 1.0f - v * inv_count;
    }
}

// v_x = -(v_loss / count) * (conv(D1) + 2 x conv(D2) + y conv(D3))
__global__ void __launch_bounds__(kSsimT *kSsimT) ssim_bwd_kernel(
    const int C, const int H, const int W, const float *__restrict__ x, const Strides4 xs, const float *__restrict__ y,
    const Strides4 ys, const float *__restrict__ maps, const int64_t plane, const float *__restrict__ v_loss,
    const float inv_count, float *__restrict__ v_x, const Strides4 vs, const SsimWin win
)
{
    __shared__ float sm[3][kSsimIn][kSsimIn + 1];
    __shared__ float hz[3][kSsimIn][kSsimT + 1];
    const int bc = blockIdx.z, b = bc / C, c = bc % C;
    const int x0 = blockIdx.x * kSsimT - kSsimR, y0 = blockIdx.y * kSsimT - kSsimR;
    const int t  = threadIdx.y * kSsimT + threadIdx.x;
    const float *mb = maps + (int64_t)bc * H * W;
    for(int i = t; i < kSsimIn * kSsimIn; i += kSsimT * kSsimT)
    {
        const int r = i / kSsimIn, q = i % kSsimIn;
        const int gy = y0 + r, gx = x0 + q;
        const bool in = gy >= 0 && gy < H && gx >= 0 && gx < W;
        const int64_t o = (int64_t)gy * W + gx;
        sm[0][r][q] = in ? mb[o] : 0.f;
        sm[1][r][q] = in ? mb[plane + o] : 0.f;
        sm[2][r][q] = in ? mb[2 * plane + o] : 0.f;
    }
    __syncthreads();
    for(int i = t; i < kSsimIn * kSsimT; i += kSsimT * kSsimT)
    {
        const int r = i / kSsimT, q = i % kSsimT;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
        for(int k = 0; k < 11; ++k)
        {
            const float w = win.w[k];
            a0 += w * sm[0][r][q + k], a1 += w * sm[1][r][q + k], a2 += w * sm[2][r][q + k];
        }
        hz[0][r][q] = a0, hz[1][r][q] = a1, hz[2][r][q] = a2;
    }
    __syncthreads();
    float c1 = 0.f, c2 = 0.f, c3 = 0.f;
#pragma unroll
    for(int k = 0; k < 11; ++k)
    {
        const float w = win.w[k];
        c1 += w * hz[0][threadIdx.y + k][threadIdx.x];
        c2 += w * hz[1][threadIdx.y + k][threadIdx.x];
        c3 += w * hz[2][threadIdx.y + k][threadIdx.x];
    }
    const int gx = blockIdx.x * kSsimT + threadIdx.x, gy = blockIdx.y * kSsimT + threadIdx.y;
    if(gx < W && gy < H)
    {
        const float xv = x[b * xs.b + c * xs.c + gy * xs.h + gx * xs.w];
        const float yv = y[b * ys.b + c * ys.c + gy * ys.h + gx * ys.w];
        const float g  = -(*v_loss) * inv_count;
        v_x[b * vs.b + c * vs.c + gy * vs.h + gx * vs.w] = g * (c1 + 2.f * xv * c2 + yv * c3);
    }
}

static SsimWin ssim_window()
{
    // gsplat/losses.py:82-87 `_gaussian_kernel_1d(11, 1.5)`, float32 as torch evaluates it
    SsimWin win;
    float s = 0.f;
    for(int i = 0; i < 11; ++i)
    {
        const float d = (float)(i - 5);
        win.w[i]      = expf(-(d * d) / (2.f * 1.5f * 1.5f));
        s += win.w[i];
    }
    for(int i = 0; i < 11; ++i)
        win.w[i] /= s;
    return win;
}
} // namespace gsb

extern "C" size_t gsb200_ssim_workspace_bytes(int64_t B, int64_t C, int64_t H, int64_t W)
{
    if(B <= 0 || C <= 0 || H <= 0 || W <= 0)
        return 0;
    return sizeof(float) * (size_t)(B * C * ((H + gsb::kSsimT - 1) / gsb::kSsimT) * ((W + gsb::kSsimT - 1) / gsb::kSsimT));
}

// loss[0] = 1 - mean(SSIM(x, y)) over B*C*H*W; x, y addressed as p[b*s[0] + c*s[1] + h*s[2] + w*s[3]] (element strides).
// maps: 3*B*C*H*W floats kept for the backward, or NULL when no gradient is needed.
extern "C" int gsb200_ssim_fwd(
    int64_t B, int64_t C, int64_t H, int64_t W, const float *x, const int64_t *x_strides, const float *y,
    const int64_t *y_strides, float *maps, void *workspace, float *loss, void *stream
)
{
    if(B <= 0 || C <= 0 || H <= 0 || W <= 0 || !x || !y || !x_strides || !y_strides || !workspace || !loss)
        return GSB200_E_INVALID;
    if(B * C > 65535 || H > (1 << 24) || W > (1 << 24))
        return GSB200_E_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    static const gsb::SsimWin win = gsb::ssim_window();
    const gsb::Strides4 xs{x_strides[0], x_strides[1], x_strides[2], x_strides[3]};
    const gsb::Strides4 ys{y_strides[0], y_strides[1], y_strides[2], y_strides[3]};
    dim3 grid((unsigned)((W + gsb::kSsimT - 1) / gsb::kSsimT), (unsigned)((H + gsb::kSsimT - 1) / gsb::kSsimT), (unsigned)(B * C));
    dim3 block(gsb::kSsimT, gsb::kSsimT);
    float *partial = static_cast<float *>(workspace);
    gsb::ssim_fwd_kernel<<<grid, block, 0, st>>>((int)C, (int)H, (int)W, x, xs, y, ys, maps, B * C * H * W, partial, win);
    if(int rc = gsb::check_launch())
        return rc;
    const int64_t nblocks = (int64_t)grid.x * grid.y * grid.z;
    gsb::ssim_final_kernel<<<1, 1024, 0, st>>>(nblocks, partial, 1.0f / (float)((double)B * C * H * W), loss);
    return gsb::check_launch();
}

// v_x (addressed through v_strides) = d loss / d x * v_loss[0]; v_loss is a DEVICE scalar.
extern "C" int gsb200_ssim_bwd(
    int64_t B, int64_t C, int64_t H, int64_t W, const float *x, const int64_t *x_strides, const float *y,
    const int64_t *y_strides, const float *maps, const float *v_loss, float *v_x, const int64_t *v_strides, void *stream
)
{
    if(B <= 0 || C <= 0 || H <= 0 || W <= 0 || !x || !y || !x_strides || !y_strides || !maps || !v_loss || !v_x || !v_strides)
        return GSB200_E_INVALID;
    if(B * C > 65535)
        return GSB200_E_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    static const gsb::SsimWin win = gsb::ssim_window();
    const gsb::Strides4 xs{x_strides[0], x_strides[1], x_strides[2], x_strides[3]};
    const gsb::Strides4 ys{y_strides[0], y_strides[1], y_strides[2], y_strides[3]};
    const gsb::Strides4 vs{v_strides[0], v_strides[1], v_strides[2], v_strides[3]};
    dim3 grid((unsigned)((W + gsb::kSsimT - 1) / gsb::kSsimT), (unsigned)((H + gsb::kSsimT - 1) / gsb::kSsimT), (unsigned)(B * C));
    dim3 block(gsb::kSsimT, gsb::kSsimT);
    gsb::ssim_bwd_kernel<<<grid, block, 0, st>>>(
        (int)C, (int)H, (int)W, x, xs, y, ys, maps, B * C * H * W, v_loss, 1.0f / (float)((double)B * C * H * W), v_x, vs, win
    );
    return gsb::check_launch();
}
