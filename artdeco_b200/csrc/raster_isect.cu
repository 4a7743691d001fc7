// Tile intersection: scan of per-Gaussian tile counts, (key,value) emission, radix sort, tile ranges.
// Integer / byte work with a BIT-EXACT contract against the oracle:
//   key = (cam << (32+tile_bits)) | (tile << 32) | float_bits(depth),  val = cam*N + gaussian,
//   emission order: Gaussian-major, then ty outer / tx inner; stable ascending sort over the live bits only.
// Replaces gsplat's isect_tiles / radix sort / isect_offset_encode that the reference reaches through
// Reconstruct/scene/scene_models/h3dgsv3.py:664-680.
//
// Compiled with -fmad=false like raster_project.cu would not matter here (only /16, floor, ceil) but the tile
// bound arithmetic below must stay textually identical to project_fwd_kernel's count.
#include "raster_common.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

namespace {

struct ToI64 {
    __host__ __device__ __forceinline__ int64_t operator()(const int32_t& v) const { return (int64_t)v; }
};

__global__ void __launch_bounds__(256)
isect_emit_kernel(int N, const int32_t* __restrict__ radii, const float* __restrict__ splats,
                  const int64_t* __restrict__ cum_tiles, int W, int H, int cam_id, int tile_bits, int convention,
                  int64_t* __restrict__ keys, int32_t* __restrict__ vals) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int2 r = reinterpret_cast<const int2*>(radii)[i];
    if (r.x <= 0 && r.y <= 0) return;
    const float4 rec0 = reinterpret_cast<const float4*>(splats + (size_t)i * ADB_SPLAT_STRIDE)[0];
    const float depth = splats[(size_t)i * ADB_SPLAT_STRIDE + 11];
    const int tw = (W + ADB_TILE - 1) / ADB_TILE;
    int x0, x1, y0, y1;
    adb_tile_rect(rec0.x, rec0.y, r.x, r.y, W, H, convention, x0, x1, y0, y1);
    int64_t o = (i == 0) ? 0 : cum_tiles[i - 1];
    if (convention != ADB_CONV_GSPLAT && cum_tiles[i] == o) return;  // kept its radius but was given no tiles
    const uint64_t hi = (uint64_t)cam_id << (32 + tile_bits);
    const uint64_t dbits = (uint64_t)__float_as_uint(depth);
    const int32_t val = cam_id * N + i;
    for (int ty = y0; ty < y1; ++ty)
        for (int tx = x0; tx < x1; ++tx) {
            uint64_t tile = (uint64_t)(ty * tw + tx);
            keys[o] = (int64_t)(hi | (tile << 32) | dbits);
            vals[o] = val;
            ++o;
        }
}

// offsets[t] = first sorted index whose tile id >= t; offsets[T] = n.
__global__ void __launch_bounds__(256)
tile_offsets_kernel(int64_t n, const int64_t* __restrict__ keys, int tile_bits, int T, int32_t* __restrict__ offsets) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t mask = ((uint64_t)1 << tile_bits) - 1;
    if (n == 0) {
        if (i <= T) offsets[i] = 0;
        return;
    }
    if (i >= n) return;
    const int cur = (int)((((uint64_t)keys[i]) >> 32) & mask);
    if (i == 0) {
        for (int t = 0; t <= cur; ++t) offsets[t] = 0;
    } else {
        const int prev = (int)((((uint64_t)keys[i - 1]) >> 32) & mask);
        for (int t = prev + 1; t <= cur; ++t) offsets[t] = (int32_t)i;
    }
    if (i == n - 1)
        for (int t = cur + 1; t <= T; ++t) offsets[t] = (int32_t)n;
}

}  // namespace

ADB_API int adb_raster_scan_workspace_bytes(int N, size_t* bytes) {
    ADB_REQUIRE(bytes && N >= 0, "adb_raster_scan_workspace_bytes: bad args");
    size_t b = 0;
    cub::TransformInputIterator<int64_t, ToI64, const int32_t*> it((const int32_t*)nullptr, ToI64());
    ADB_CUDA(cub::DeviceScan::InclusiveSum(nullptr, b, it, (int64_t*)nullptr, N));
    *bytes = b + 256;
    return ADB_OK;
}

// cum_tiles[i] = sum_{j<=i} tiles_per_gauss[j]  (int64).  The caller reads cum_tiles[N-1] to size keys/vals.
ADB_API int adb_raster_isect_scan(int N, const int32_t* tiles_per_gauss, int64_t* cum_tiles, void* ws,
                                  size_t ws_bytes, cudaStream_t stream) {
    ADB_REQUIRE(N >= 0, "adb_raster_isect_scan: bad N");
    if (N == 0) return ADB_OK;
    ADB_REQUIRE(tiles_per_gauss && cum_tiles && ws, "adb_raster_isect_scan: null pointer");
    cub::TransformInputIterator<int64_t, ToI64, const int32_t*> it(tiles_per_gauss, ToI64());
    size_t need = 0;
    ADB_CUDA(cub::DeviceScan::InclusiveSum(nullptr, need, it, cum_tiles, N));
    if (need > ws_bytes) { adb_set_error_msg("adb_raster_isect_scan: workspace too small"); return ADB_ERR_WORKSPACE; }
    ADB_CUDA(cub::DeviceScan::InclusiveSum(ws, need, it, cum_tiles, N, stream));
    return ADB_OK;
}

static int isect_emit_impl(int convention, int N, const int32_t* radii, const float* splats, const int64_t* cum_tiles,
                           int W, int H, int cam_id, int n_cams, int64_t* keys, int32_t* vals, cudaStream_t stream) {
    ADB_REQUIRE(N >= 0 && W > 0 && H > 0 && cam_id >= 0 && n_cams > cam_id, "adb_raster_isect_emit: bad sizes");
    if (N == 0) return ADB_OK;
    ADB_REQUIRE(radii && splats && cum_tiles && keys && vals, "adb_raster_isect_emit: null pointer");
    isect_emit_kernel<<<adb_cdiv(N, 256), 256, 0, stream>>>(N, radii, splats, cum_tiles, W, H, cam_id,
                                                           adb_tile_bits(W, H), convention, keys, vals);
    ADB_CHECK_LAUNCH("isect_emit_kernel");
    return ADB_OK;
}

ADB_API int adb_raster_isect_emit(int N, const int32_t* radii, const float* splats, const int64_t* cum_tiles,
                                  int W, int H, int cam_id, int n_cams, int64_t* keys, int32_t* vals,
                                  cudaStream_t stream) {
    return isect_emit_impl(ADB_CONV_GSPLAT, N, radii, splats, cum_tiles, W, H, cam_id, n_cams, keys, vals, stream);
}

// Legacy (Inria) tile rectangle; everything else (key layout, emission order) as above.
ADB_API int adb_raster_isect_emit_legacy(int N, const int32_t* radii, const float* splats, const int64_t* cum_tiles,
                                         int W, int H, int cam_id, int n_cams, int64_t* keys, int32_t* vals,
                                         cudaStream_t stream) {
    return isect_emit_impl(ADB_CONV_INRIA, N, radii, splats, cum_tiles, W, H, cam_id, n_cams, keys, vals, stream);
}

ADB_API int adb_raster_sort_workspace_bytes(long long n_isect, size_t* bytes) {
    ADB_REQUIRE(bytes && n_isect >= 0, "adb_raster_sort_workspace_bytes: bad args");
    size_t b = 0;
    cub::DoubleBuffer<int64_t> dk(nullptr, nullptr);
    cub::DoubleBuffer<int32_t> dv(nullptr, nullptr);
    ADB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, b, dk, dv, (int64_t)n_isect, 0, 64));
    *bytes = b + 256;
    return ADB_OK;
}

// Stable ascending sort over the live key bits.  keys_a/vals_a hold the input; the sorted result lands in
// whichever buffer pair *sorted_in_b reports (0 -> a, 1 -> b); both pairs must hold n_isect elements.
ADB_API int adb_raster_sort(long long n_isect, int W, int H, int n_cams, int64_t* keys_a, int32_t* vals_a,
                            int64_t* keys_b, int32_t* vals_b, void* ws, size_t ws_bytes, int* sorted_in_b,
                            cudaStream_t stream) {
    ADB_REQUIRE(n_isect >= 0 && sorted_in_b, "adb_raster_sort: bad args");
    *sorted_in_b = 0;
    if (n_isect == 0) return ADB_OK;
    ADB_REQUIRE(keys_a && vals_a && keys_b && vals_b && ws, "adb_raster_sort: null pointer");
    const int end_bit = 32 + adb_tile_bits(W, H) + adb_cam_bits(n_cams);
    cub::DoubleBuffer<int64_t> dk(keys_a, keys_b);
    cub::DoubleBuffer<int32_t> dv(vals_a, vals_b);
    size_t need = 0;
    ADB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, need, dk, dv, (int64_t)n_isect, 0, end_bit));
    if (need > ws_bytes) { adb_set_error_msg("adb_raster_sort: workspace too small"); return ADB_ERR_WORKSPACE; }
    ADB_CUDA(cub::DeviceRadixSort::SortPairs(ws, need, dk, dv, (int64_t)n_isect, 0, end_bit, stream));
    *sorted_in_b = (dk.Current() == keys_b) ? 1 : 0;
    return ADB_OK;
}

// tile_offsets has T+1 entries (T = ceil(W/16)*ceil(H/16)); entry T is n_isect.
ADB_API int adb_raster_tile_offsets(long long n_isect, const int64_t* keys_sorted, int W, int H,
                                    int32_t* tile_offsets, cudaStream_t stream) {
    ADB_REQUIRE(n_isect >= 0 && W > 0 && H > 0 && tile_offsets, "adb_raster_tile_offsets: bad args");
    ADB_REQUIRE(n_isect < 2147483647LL, "adb_raster_tile_offsets: more than 2^31 intersections");
    const int T = adb_cdiv(W, ADB_TILE) * adb_cdiv(H, ADB_TILE);
    long long threads = n_isect > 0 ? n_isect : (long long)T + 1;
    tile_offsets_kernel<<<adb_cdiv(threads, 256), 256, 0, stream>>>(n_isect, keys_sorted, adb_tile_bits(W, H), T,
                                                                   tile_offsets);
    ADB_CHECK_LAUNCH("tile_offsets_kernel");
    return ADB_OK;
}

