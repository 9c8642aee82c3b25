// sort.cu -- (isect_id, flatten_id) radix sort and library-level helpers.
//
// The sorts are cub::DeviceRadixSort::SortPairs (the library call the reference makes,
// csrc/IntersectTile.cu:1078-1121; its onesweep kernels are compiled here for sm_100a), used twice:
// once over the N projected rows on (image, depth) and once over the S intersections on the (image, tile)
// bits only -- see gsb200_isect_depth_order below.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace gsb
{
static thread_local cudaError_t g_last_error = cudaSuccess;
void set_last_cuda_error(cudaError_t e) { g_last_error = e; }
} // namespace gsb

extern "C" const char *gsb200_version(void) { return "gsplat_b200 0.1.0 (sm_100a)"; }

extern "C" const char *gsb200_error_string(int code)
{
    switch(code)
    {
    case GSB200_OK: return "ok";
    case GSB200_E_INVALID: return "invalid argument";
    case GSB200_E_UNSUPPORTED: return "unsupported configuration (camera model / channel count / tile size not built)";
    case GSB200_E_CUDA: return "CUDA error";
    case GSB200_E_WORKSPACE: return "workspace too small";
    case GSB200_E_KEYBITS: return "intersect_tile: (image, tile) id packing needs more than 32 bits";
    default: return "unknown error";
    }
}

extern "C" const char *gsb200_last_cuda_error(void) { return cudaGetErrorString(gsb::g_last_error); }

extern "C" uint32_t gsb200_bits_for_count(int64_t count) { return gsb::bits_for_count(count); }

extern "C" size_t gsb200_sort_workspace_bytes(int64_t n_isects, int begin_bit, int end_bit)
{
    if(n_isects <= 0)
        return 0;
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(
        (void *)nullptr, bytes, (const int64_t *)nullptr, (int64_t *)nullptr, (const int32_t *)nullptr, (int32_t *)nullptr,
        n_isects, begin_bit, end_bit
    );
    return bytes + 256;
}

// Stable LSD radix sort of the (key, value) pairs on key bits [begin_bit, end_bit).
extern "C" int gsb200_sort_pairs(
    int64_t n_isects, int begin_bit, int end_bit, const int64_t *keys_in, const int32_t *vals_in, int64_t *keys_out,
    int32_t *vals_out, void *workspace, size_t workspace_bytes, void *stream
)
{
    if(n_isects < 0 || begin_bit < 0 || end_bit < begin_bit || end_bit > 64)
        return GSB200_E_INVALID;
    if(n_isects == 0)
        return GSB200_OK;
    if(!keys_in || !vals_in || !keys_out || !vals_out || !workspace)
        return GSB200_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    size_t need     = 0;
    cub::DeviceRadixSort::SortPairs((void *)nullptr, need, keys_in, keys_out, vals_in, vals_out, n_isects, begin_bit, end_bit, st);
    if(need > workspace_bytes)
        return GSB200_E_WORKSPACE;
    GSB_CUDA_TRY(cub::DeviceRadixSort::SortPairs(workspace, need, keys_in, keys_out, vals_in, vals_out, n_isects, begin_bit, end_bit, st));
    return GSB200_OK;
}

// ---- depth order of the projected gaussians (pre-pass of the tile intersection)
// The reference sorts all S intersections on (image, tile, depth) = 45+ key bits.  Sorting the rows ONCE by
// (image, depth) and emitting the intersections in that order leaves only the (image, tile) bits for the
// S-sized sort: both sorts are stable, so inside a tile the result is (depth, then row index) -- exactly the
// order the reference's single stable sort produces (csrc/Intersect.cpp:283-326).
namespace gsb
{
__global__ void __launch_bounds__(256) depth_key_kernel(
    int64_t total, int64_t N, const int32_t *__restrict__ radii, const float *__restrict__ depths,
    const int64_t *__restrict__ image_ids, uint64_t *__restrict__ keys, int32_t *__restrict__ rows
)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= total)
        return;
    const int2 r = reinterpret_cast<const int2 *>(radii)[i];
    uint64_t k   = ~0ull; // culled rows go last
    if(r.x > 0 && r.y > 0)
        k = ((uint64_t)(image_ids ? image_ids[i] : i / N) << 32) | (uint64_t)__float_as_uint(depths[i]);
    keys[i] = k;
    rows[i] = (int32_t)i;
}

struct DepthOrderLayout
{
    size_t keys_in, keys_out, rows_in, cub, total;
};
static DepthOrderLayout depth_order_layout(int64_t total, int end_bit)
{
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    DepthOrderLayout L;
    size_t cub_bytes = 0;
    cub::DeviceRadixSort::SortPairs(
        (void *)nullptr, cub_bytes, (const uint64_t *)nullptr, (uint64_t *)nullptr, (const int32_t *)nullptr, (int32_t *)nullptr,
        total, 0, end_bit
    );
    L.keys_in  = 0;
    L.keys_out = L.keys_in + al(sizeof(uint64_t) * (size_t)total);
    L.rows_in  = L.keys_out + al(sizeof(uint64_t) * (size_t)total);
    L.cub      = L.rows_in + al(sizeof(int32_t) * (size_t)total);
    L.total    = L.cub + al(cub_bytes) + 256;
    return L;
}
} // namespace gsb

extern "C" size_t gsb200_isect_depth_order_workspace_bytes(int64_t I, int64_t total_rows)
{
    if(total_rows <= 0 || I <= 0)
        return 0;
    return gsb::depth_order_layout(total_rows, 32 + (int)gsb::bits_for_count(I)).total;
}

extern "C" int gsb200_isect_depth_order(
    int64_t I, int64_t N, const int32_t *radii, const float *depths, const int64_t *image_ids, int32_t *order,
    void *workspace, size_t workspace_bytes, void *stream
)
{
    if(I < 0 || N < 0)
        return GSB200_E_INVALID;
    const int64_t total = image_ids ? N : I * N; // packed: N rows in total
    if(total == 0)
        return GSB200_OK;
    if(!radii || !depths || !order || !workspace || total > 0x7fffffffLL)
        return GSB200_E_INVALID;
    const int end_bit = 32 + (int)gsb::bits_for_count(I);
    const auto L      = gsb::depth_order_layout(total, end_bit);
    if(L.total > workspace_bytes)
        return GSB200_E_WORKSPACE;
    char *ws         = static_cast<char *>(workspace);
    uint64_t *k_in   = reinterpret_cast<uint64_t *>(ws + L.keys_in), *k_out = reinterpret_cast<uint64_t *>(ws + L.keys_out);
    int32_t *rows_in = reinterpret_cast<int32_t *>(ws + L.rows_in);
    cudaStream_t st  = (cudaStream_t)stream;
    gsb::depth_key_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(total, N, radii, depths, image_ids, k_in, rows_in);
    if(int rc = gsb::check_launch())
        return rc;
    size_t cub_bytes = L.total - L.cub;
    GSB_CUDA_TRY(cub::DeviceRadixSort::SortPairs(ws + L.cub, cub_bytes, k_in, k_out, rows_in, order, total, 0, end_bit, st));
    return GSB200_OK;
}

