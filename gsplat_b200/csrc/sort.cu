// sort.cu -- (isect_id, flatten_id) radix sort and library-level helpers.
//
// The sort is cub::DeviceRadixSort::SortPairs on key bits [0, 32 + tile_bits + image_bits), the same
// library call the reference makes (csrc/IntersectTile.cu:1078-1121); the onesweep kernels it
// instantiates are compiled here for sm_100a.  Stable, so equal (tile, depth) keys keep emit order.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_segmented_sort.cuh>

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

extern "C" size_t gsb200_sort_workspace_bytes(int64_t n_isects, int end_bit)
{
    if(n_isects <= 0)
        return 0;
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(
        (void *)nullptr, bytes, (const int64_t *)nullptr, (int64_t *)nullptr, (const int32_t *)nullptr, (int32_t *)nullptr,
        n_isects, 0, end_bit
    );
    return bytes + 256;
}

extern "C" int gsb200_sort_pairs(
    int64_t n_isects, int end_bit, const int64_t *keys_in, const int32_t *vals_in, int64_t *keys_out, int32_t *vals_out,
    void *workspace, size_t workspace_bytes, void *stream
)
{
    if(n_isects < 0 || end_bit < 0 || end_bit > 64)
        return GSB200_E_INVALID;
    if(n_isects == 0)
        return GSB200_OK;
    if(!keys_in || !vals_in || !keys_out || !vals_out || !workspace)
        return GSB200_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    size_t need     = 0;
    cub::DeviceRadixSort::SortPairs((void *)nullptr, need, keys_in, keys_out, vals_in, vals_out, n_isects, 0, end_bit, st);
    if(need > workspace_bytes)
        return GSB200_E_WORKSPACE;
    GSB_CUDA_TRY(cub::DeviceRadixSort::SortPairs(workspace, need, keys_in, keys_out, vals_in, vals_out, n_isects, 0, end_bit, st));
    return GSB200_OK;
}

// Per-tile sort of the bucketed (depth << 32 | gaussian) keys: cub::DeviceSegmentedSort, segments =
// tiles (offsets has n_segments + 1 entries).  Ascending 64-bit keys == (depth bits, emit order), i.e.
// exactly the order the stable global radix sort of the reference produces inside a tile.
extern "C" size_t gsb200_segsort_workspace_bytes(int64_t n_items, int64_t n_segments)
{
    if(n_items <= 0 || n_segments <= 0)
        return 0;
    size_t bytes = 0;
    cub::DeviceSegmentedSort::SortKeys(
        (void *)nullptr, bytes, (const uint64_t *)nullptr, (uint64_t *)nullptr, (int)n_items, (int)n_segments,
        (const int32_t *)nullptr, (const int32_t *)nullptr
    );
    return bytes + 256;
}

extern "C" int gsb200_segsort_keys(
    int64_t n_items, int64_t n_segments, const int32_t *offsets, const uint64_t *keys_in, uint64_t *keys_out,
    void *workspace, size_t workspace_bytes, void *stream
)
{
    if(n_items < 0 || n_segments < 0 || n_items > 0x7fffffffLL || n_segments > 0x7fffffffLL)
        return GSB200_E_INVALID;
    if(n_items == 0 || n_segments == 0)
        return GSB200_OK;
    if(!offsets || !keys_in || !keys_out || !workspace)
        return GSB200_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    size_t need     = 0;
    cub::DeviceSegmentedSort::SortKeys((void *)nullptr, need, keys_in, keys_out, (int)n_items, (int)n_segments, offsets, offsets + 1, st);
    if(need > workspace_bytes)
        return GSB200_E_WORKSPACE;
    GSB_CUDA_TRY(cub::DeviceSegmentedSort::SortKeys(workspace, need, keys_in, keys_out, (int)n_items, (int)n_segments, offsets, offsets + 1, st));
    return GSB200_OK;
}
