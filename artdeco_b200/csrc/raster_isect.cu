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


// =====================================================================================================================
// Tile-bucketed intersection (round 2): no library sort, no host sync, bit-identical output.
//
//   count    per Gaussian: one RED.ADD per touched tile into tile_counts[T]                (integer, 16 B read / Gaussian)
//   scan     one CTA: exclusive scan of tile_counts -> tile_offsets[T+1] (clamped to the caller's capacity; total and
//            overflow flag stay on the device)
//   scatter  per Gaussian: slot = offsets[tile] + atomicSub(counts[tile]) - 1 -> packed[slot] = depth_bits<<32 | gaussian
//            (arbitrary order inside a tile; the counters return to zero, so the buffer needs no memset between calls)
//   sort     one CTA per tile: bitonic sort of the tile's 64-bit words in shared memory (in place in global memory for
//            tiles beyond 4096 entries), then keys = cam|tile|depth_bits and vals = cam*N + gaussian are written out.
//
// Why the order equals the reference's stable radix sort of (tile, depth): the reference emits keys Gaussian-major, so
// inside one (tile, depth) tie the stable sort leaves ascending Gaussian ids — exactly the ascending order of the
// composite word (depth_bits, gaussian).  A Gaussian touches a tile at most once, so the words of a tile are distinct
// and the result does not depend on the scatter order.  (oracle: adbo_isect_sort.)
// The global 45-bit onesweep radix sort this replaces moved 6 x 24 B per intersection and was 12.6 % of the round-1 step.
namespace {

__global__ void __launch_bounds__(256)
tile_count_kernel(int N, const int32_t* __restrict__ radii, const float* __restrict__ splats,
                  const int32_t* __restrict__ tiles_per_gauss, int W, int H, int convention,
                  int32_t* __restrict__ tile_counts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    if (tiles_per_gauss[i] <= 0) return;
    const int2 r = reinterpret_cast<const int2*>(radii)[i];
    const float2 m = reinterpret_cast<const float2*>(splats + (size_t)i * ADB_SPLAT_STRIDE)[0];
    const int tw = (W + ADB_TILE - 1) / ADB_TILE;
    int x0, x1, y0, y1;
    adb_tile_rect(m.x, m.y, r.x, r.y, W, H, convention, x0, x1, y0, y1);
    for (int ty = y0; ty < y1; ++ty)
        for (int tx = x0; tx < x1; ++tx) atomicAdd(tile_counts + ty * tw + tx, 1);
}

// Single CTA, 1024 threads.  offsets[t] = min(capacity, sum_{u<t} counts[u]); offsets[T] likewise; *total = unclamped sum.
__global__ void __launch_bounds__(1024)
tile_scan_kernel(int T, const int32_t* __restrict__ counts, long long capacity, int32_t* __restrict__ offsets,
                 long long* __restrict__ total, int32_t* __restrict__ overflow) {
    __shared__ long long s_warp[32];
    __shared__ long long s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < T; base += 1024) {
        const int t = base + tid;
        const long long c = t < T ? (long long)counts[t] : 0;
        long long x = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) s_warp[warp] = x;
        __syncthreads();
        if (warp == 0) {
            long long w = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const long long y = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += y;
            }
            s_warp[lane] = w;   // inclusive over warps
        }
        __syncthreads();
        const long long carry = s_carry;
        const long long excl = carry + (warp ? s_warp[warp - 1] : 0) + x - c;
        if (t < T) offsets[t] = (int32_t)(excl < capacity ? excl : capacity);
        __syncthreads();
        if (tid == 1023) s_carry = carry + s_warp[31];
        __syncthreads();
    }
    if (tid == 0) {
        const long long tot = s_carry;
        offsets[T] = (int32_t)(tot < capacity ? tot : capacity);
        if (total) *total = tot;
        if (overflow && tot > capacity) *overflow = 1;
    }
}

__global__ void __launch_bounds__(256)
tile_scatter_kernel(int N, const int32_t* __restrict__ radii, const float* __restrict__ splats,
                    const int32_t* __restrict__ tiles_per_gauss, int W, int H, int convention,
                    const int32_t* __restrict__ offsets, int32_t* __restrict__ tile_counts, long long capacity,
                    unsigned long long* __restrict__ packed) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    if (tiles_per_gauss[i] <= 0) return;
    const int2 r = reinterpret_cast<const int2*>(radii)[i];
    const float2 m = reinterpret_cast<const float2*>(splats + (size_t)i * ADB_SPLAT_STRIDE)[0];
    const float depth = splats[(size_t)i * ADB_SPLAT_STRIDE + 11];
    const int tw = (W + ADB_TILE - 1) / ADB_TILE;
    int x0, x1, y0, y1;
    adb_tile_rect(m.x, m.y, r.x, r.y, W, H, convention, x0, x1, y0, y1);
    const unsigned long long word = ((unsigned long long)__float_as_uint(depth) << 32) | (unsigned)i;
    for (int ty = y0; ty < y1; ++ty)
        for (int tx = x0; tx < x1; ++tx) {
            const int t = ty * tw + tx;
            const long long slot = (long long)offsets[t] + (atomicSub(tile_counts + t, 1) - 1);
            if (slot < capacity) packed[slot] = word;
        }
}

constexpr int SORT_THREADS = 256;
constexpr int SORT_SMEM_CAP = 4096;    // words sorted in shared memory (32 KB); larger tiles sort in place in global memory

// All compare-exchanges ascending ("flip" then "disperse" steps), so indices >= n behave as +inf and are simply skipped:
// any n works without padding.
template <class Ptr>
__device__ __forceinline__ void bitonic_sort_words(Ptr a, int n, int tid, int nthreads) {
    int np2 = 1;
    while (np2 < n) np2 <<= 1;
    const int pairs = np2 >> 1;
    for (int k = 2; k <= np2; k <<= 1) {
        const int h = k >> 1;
        for (int q = tid; q < pairs; q += nthreads) {                  // flip: i <-> block_end - 1 - (i - block_start)
            const int blk = q / h, r = q - blk * h;
            const int i = blk * k + r, p = blk * k + (k - 1 - r);
            if (p < n) {
                const unsigned long long x = a[i], y = a[p];
                if (x > y) { a[i] = y; a[p] = x; }
            }
        }
        __syncthreads();
        for (int j = h >> 1; j > 0; j >>= 1) {                          // disperse: i <-> i + j
            for (int q = tid; q < pairs; q += nthreads) {
                const int blk = q / j, r = q - blk * j;
                const int i = blk * 2 * j + r, p = i + j;
                if (p < n) {
                    const unsigned long long x = a[i], y = a[p];
                    if (x > y) { a[i] = y; a[p] = x; }
                }
            }
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(SORT_THREADS)
tile_sort_kernel(int T, const int32_t* __restrict__ offsets, unsigned long long* __restrict__ packed, int cam_id,
                 int n_per_cam, int tile_bits, int64_t* __restrict__ keys, int32_t* __restrict__ vals) {
    __shared__ unsigned long long s_words[SORT_SMEM_CAP];
    const int tile = blockIdx.x;
    const int start = offsets[tile], n = offsets[tile + 1] - start;
    if (n <= 0) return;
    const int tid = threadIdx.x;
    unsigned long long* g = packed + start;
    const unsigned long long hi = ((unsigned long long)cam_id << (32 + tile_bits)) | ((unsigned long long)tile << 32);
    const int vbase = cam_id * n_per_cam;
    if (n <= SORT_SMEM_CAP) {
        for (int i = tid; i < n; i += SORT_THREADS) s_words[i] = g[i];
        __syncthreads();
        if (n > 1) bitonic_sort_words(s_words, n, tid, SORT_THREADS);
        for (int i = tid; i < n; i += SORT_THREADS) {
            const unsigned long long w = s_words[i];
            keys[start + i] = (int64_t)(hi | (w >> 32));
            vals[start + i] = vbase + (int32_t)(unsigned)(w & 0xffffffffu);
        }
    } else {
        // rare: a tile with more than SORT_SMEM_CAP splats.  Same network, operands in global memory (L2-resident).
        __syncthreads();
        bitonic_sort_words((volatile unsigned long long*)g, n, tid, SORT_THREADS);
        __threadfence_block();
        for (int i = tid; i < n; i += SORT_THREADS) {
            const unsigned long long w = g[i];
            keys[start + i] = (int64_t)(hi | (w >> 32));
            vals[start + i] = vbase + (int32_t)(unsigned)(w & 0xffffffffu);
        }
    }
}

}  // namespace

// tile_counts [T] int32 must be all-zero on entry (it is zero again after adb_raster_tile_scatter: allocate + zero once).
// offsets [T+1]; total (int64, device) and overflow (int32, device, only ever SET) may be null.  `capacity` = number of
// elements `packed`, `keys` and `vals` can hold; intersections beyond it are dropped (memory-safe) and flagged.
static int tile_bucket_impl(int convention, int N, const int32_t* radii, const float* splats,
                            const int32_t* tiles_per_gauss, int W, int H, long long capacity, int32_t* tile_counts,
                            int32_t* tile_offsets, long long* total, int32_t* overflow, cudaStream_t stream) {
    ADB_REQUIRE(N >= 0 && W > 0 && H > 0 && capacity >= 0 && capacity < 2147483647LL, "adb_raster_tile_count_scan: bad sizes");
    ADB_REQUIRE(tile_counts && tile_offsets, "adb_raster_tile_count_scan: null pointer");
    const int T = adb_cdiv(W, ADB_TILE) * adb_cdiv(H, ADB_TILE);
    if (N > 0) {
        ADB_REQUIRE(radii && splats && tiles_per_gauss, "adb_raster_tile_count_scan: null pointer");
        tile_count_kernel<<<adb_cdiv(N, 256), 256, 0, stream>>>(N, radii, splats, tiles_per_gauss, W, H, convention,
                                                               tile_counts);
        ADB_CHECK_LAUNCH("tile_count_kernel");
    }
    tile_scan_kernel<<<1, 1024, 0, stream>>>(T, tile_counts, capacity, tile_offsets, total, overflow);
    ADB_CHECK_LAUNCH("tile_scan_kernel");
    return ADB_OK;
}

ADB_API int adb_raster_tile_count_scan(int N, const int32_t* radii, const float* splats, const int32_t* tiles_per_gauss,
                                       int W, int H, int legacy, long long capacity, int32_t* tile_counts,
                                       int32_t* tile_offsets, long long* total, int32_t* overflow, cudaStream_t stream) {
    return tile_bucket_impl(legacy ? ADB_CONV_INRIA : ADB_CONV_GSPLAT, N, radii, splats, tiles_per_gauss, W, H, capacity,
                            tile_counts, tile_offsets, total, overflow, stream);
}

// Scatter + per-tile sort.  keys [capacity] int64, vals [capacity] int32, packed [capacity] uint64 scratch.
ADB_API int adb_raster_tile_scatter_sort(int N, const int32_t* radii, const float* splats, const int32_t* tiles_per_gauss,
                                         int W, int H, int legacy, int cam_id, int n_cams, long long capacity,
                                         int32_t* tile_counts, const int32_t* tile_offsets, void* packed, int64_t* keys,
                                         int32_t* vals, cudaStream_t stream) {
    ADB_REQUIRE(N >= 0 && W > 0 && H > 0 && cam_id >= 0 && n_cams > cam_id && capacity >= 0,
                "adb_raster_tile_scatter_sort: bad sizes");
    if (N == 0 || capacity == 0) return ADB_OK;
    ADB_REQUIRE(radii && splats && tiles_per_gauss && tile_counts && tile_offsets && packed && keys && vals,
                "adb_raster_tile_scatter_sort: null pointer");
    const int T = adb_cdiv(W, ADB_TILE) * adb_cdiv(H, ADB_TILE);
    tile_scatter_kernel<<<adb_cdiv(N, 256), 256, 0, stream>>>(N, radii, splats, tiles_per_gauss, W, H,
                                                             legacy ? ADB_CONV_INRIA : ADB_CONV_GSPLAT, tile_offsets,
                                                             tile_counts, capacity, (unsigned long long*)packed);
    ADB_CHECK_LAUNCH("tile_scatter_kernel");
    tile_sort_kernel<<<T, SORT_THREADS, 0, stream>>>(T, tile_offsets, (unsigned long long*)packed, cam_id, N,
                                                    adb_tile_bits(W, H), keys, vals);
    ADB_CHECK_LAUNCH("tile_sort_kernel");
    return ADB_OK;
}
