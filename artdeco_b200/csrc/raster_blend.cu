// Per-tile alpha blending, forward and backward (RGB + depth = 4 channels).
//
// Replaces gsplat's rasterize_to_pixels fwd/bwd that the reference reaches through
// Reconstruct/scene/scene_models/h3dgsv3.py:664-680:
//   pixel centre (j+0.5, i+0.5); sigma = .5(a dx^2 + c dy^2) + b dx dy; alpha = min(.999, o exp(-sigma));
//   skip if sigma<0 or alpha<1/255; stop when T(1-alpha) <= 1e-4; out += feat*alpha*T; alpha_out = 1-T.
//
// One CTA per 16x16 tile, one thread per pixel.  The tile's slice of the depth-sorted list is walked in
// batches of 256: each thread gathers one 48 B splat record (3 x LDG.128) into shared memory, then every
// pixel consumes the batch through conflict-free broadcast LDS.128 reads.
// Backward walks the same list back-to-front; per-splat gradients are reduced with warp shuffles, combined
// across the CTA's 8 warps in shared memory, and flushed with ONE global atomic per (tile, splat, component)
// (gsplat issues one per warp).
#include "raster_common.cuh"

namespace {

constexpr int BLOCK = ADB_TILE * ADB_TILE;  // 256

__device__ __forceinline__ void gather_splat(const float* __restrict__ splats, int g, float4& A, float4& B, float2& C) {
    const float4* p = reinterpret_cast<const float4*>(splats + (size_t)g * ADB_SPLAT_STRIDE);
    A = __ldg(p);
    B = __ldg(p + 1);
    float4 c4 = __ldg(p + 2);
    C = make_float2(c4.x, c4.y);
}

__global__ void __launch_bounds__(BLOCK)
blend_fwd_kernel(int W, int H, const float* __restrict__ splats, const int32_t* __restrict__ vals,
                 const int32_t* __restrict__ tile_offsets, int n_per_cam,
                 float* __restrict__ colors, float* __restrict__ alphas, int32_t* __restrict__ last_ids) {
    __shared__ float4 sA[BLOCK];
    __shared__ float4 sB[BLOCK];
    __shared__ float2 sC[BLOCK];

    const int tw = (W + ADB_TILE - 1) / ADB_TILE;
    const int tile = blockIdx.y * tw + blockIdx.x;
    const int tid = threadIdx.x;
    const int j = blockIdx.x * ADB_TILE + (tid & 15), i = blockIdx.y * ADB_TILE + (tid >> 4);
    const bool inside = (i < H && j < W);
    const float px = (float)j + 0.5f, py = (float)i + 0.5f;
    const int start = tile_offsets[tile], end = tile_offsets[tile + 1];

    float T = 1.0f;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int cur = 0;
    bool done = !inside;
    const int nb = (end - start + BLOCK - 1) / BLOCK;
    for (int b = 0; b < nb; ++b) {
        if (__syncthreads_count(done) >= BLOCK) break;
        const int bstart = start + b * BLOCK;
        const int idx = bstart + tid;
        if (idx < end) {
            int g = vals[idx] % n_per_cam;
            gather_splat(splats, g, sA[tid], sB[tid], sC[tid]);
        }
        __syncthreads();
        const int bsize = min(BLOCK, end - bstart);
        for (int t = 0; t < bsize && !done; ++t) {
            const float4 A = sA[t];
            const float4 B = sB[t];
            const float dx = A.x - px, dy = A.y - py;
            const float sigma = 0.5f * (A.z * dx * dx + B.x * dy * dy) + A.w * dx * dy;
            const float alpha = fminf(ADB_MAX_ALPHA, B.y * __expf(-sigma));
            if (sigma < 0.f || alpha < ADB_ALPHA_THRESHOLD) continue;
            const float nT = T * (1.0f - alpha);
            if (nT <= ADB_T_EPS) { done = true; break; }
            const float w = alpha * T;
            const float2 C = sC[t];
            acc.x += B.z * w; acc.y += B.w * w; acc.z += C.x * w; acc.w += C.y * w;
            cur = bstart + t;
            T = nT;
        }
    }
    if (inside) {
        const size_t pix = (size_t)i * W + j;
        reinterpret_cast<float4*>(colors)[pix] = acc;
        alphas[pix] = 1.0f - T;
        last_ids[pix] = cur;
    }
}

__global__ void __launch_bounds__(BLOCK)
blend_bwd_kernel(int W, int H, const float* __restrict__ splats, const int32_t* __restrict__ vals,
                 const int32_t* __restrict__ tile_offsets, int n_per_cam,
                 const float* __restrict__ alphas, const int32_t* __restrict__ last_ids,
                 const float* __restrict__ v_colors, const float* __restrict__ v_alphas,
                 float* __restrict__ v_splats) {
    __shared__ float4 sA[BLOCK];
    __shared__ float4 sB[BLOCK];
    __shared__ float2 sC[BLOCK];
    __shared__ int sG[BLOCK];
    __shared__ float sAcc[BLOCK][10];

    const int tw = (W + ADB_TILE - 1) / ADB_TILE;
    const int tile = blockIdx.y * tw + blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31;
    const int j = blockIdx.x * ADB_TILE + (tid & 15), i = blockIdx.y * ADB_TILE + (tid >> 4);
    const bool inside = (i < H && j < W);
    const float px = (float)j + 0.5f, py = (float)i + 0.5f;
    const int start = tile_offsets[tile], end = tile_offsets[tile + 1];
    if (end <= start) return;

    const size_t pix = (size_t)(inside ? i : 0) * W + (inside ? j : 0);
    const float T_final = inside ? 1.0f - alphas[pix] : 1.0f;
    const int bin_final = inside ? last_ids[pix] : -1;
    float4 vo = inside ? reinterpret_cast<const float4*>(v_colors)[pix] : make_float4(0.f, 0.f, 0.f, 0.f);
    const float va = inside ? v_alphas[pix] : 0.f;
    float T = T_final;
    float4 buf = make_float4(0.f, 0.f, 0.f, 0.f);

    // highest list index any pixel of this warp still needs
    int warp_bin_final = bin_final;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) warp_bin_final = max(warp_bin_final, __shfl_xor_sync(0xffffffffu, warp_bin_final, o));

    const int nb = (end - start + BLOCK - 1) / BLOCK;
    for (int b = 0; b < nb; ++b) {
        __syncthreads();  // previous batch fully consumed / flushed
        const int batch_end = end - 1 - b * BLOCK;
        const int bsize = min(BLOCK, batch_end + 1 - start);
        const int idx = batch_end - tid;
        if (idx >= start) {
            int g = vals[idx] % n_per_cam;
            sG[tid] = g;
            gather_splat(splats, g, sA[tid], sB[tid], sC[tid]);
        }
#pragma unroll
        for (int k = 0; k < 10; ++k) sAcc[tid][k] = 0.f;
        __syncthreads();
        for (int t = max(0, batch_end - warp_bin_final); t < bsize; ++t) {
            const bool active = inside && (batch_end - t <= bin_final);
            const float4 A = sA[t];
            const float4 B = sB[t];
            const float dx = A.x - px, dy = A.y - py;
            const float sigma = 0.5f * (A.z * dx * dx + B.x * dy * dy) + A.w * dx * dy;
            const float vis = __expf(-sigma);
            const float alpha = fminf(ADB_MAX_ALPHA, B.y * vis);
            const bool valid = active && !(sigma < 0.f || alpha < ADB_ALPHA_THRESHOLD);
            if (!__any_sync(0xffffffffu, valid)) continue;
            float g0 = 0.f, g1 = 0.f, g2 = 0.f, g3 = 0.f, g4 = 0.f, g5 = 0.f, g6 = 0.f, g7 = 0.f, g8 = 0.f, g9 = 0.f;
            if (valid) {
                const float2 C = sC[t];
                const float ra = 1.0f / (1.0f - alpha);
                T *= ra;
                const float fac = alpha * T;
                g6 = fac * vo.x; g7 = fac * vo.y; g8 = fac * vo.z; g9 = fac * vo.w;
                float v_alpha = (B.z * T - buf.x * ra) * vo.x + (B.w * T - buf.y * ra) * vo.y +
                                (C.x * T - buf.z * ra) * vo.z + (C.y * T - buf.w * ra) * vo.w;
                v_alpha += T_final * ra * va;
                buf.x += B.z * fac; buf.y += B.w * fac; buf.z += C.x * fac; buf.w += C.y * fac;
                if (B.y * vis <= ADB_MAX_ALPHA) {
                    const float v_sigma = -B.y * vis * v_alpha;
                    g0 = v_sigma * (A.z * dx + A.w * dy);
                    g1 = v_sigma * (A.w * dx + B.x * dy);
                    g2 = 0.5f * v_sigma * dx * dx;
                    g3 = v_sigma * dx * dy;
                    g4 = 0.5f * v_sigma * dy * dy;
                    g5 = vis * v_alpha;
                }
            }
            g0 = adb_warp_sum(g0); g1 = adb_warp_sum(g1); g2 = adb_warp_sum(g2); g3 = adb_warp_sum(g3);
            g4 = adb_warp_sum(g4); g5 = adb_warp_sum(g5); g6 = adb_warp_sum(g6); g7 = adb_warp_sum(g7);
            g8 = adb_warp_sum(g8); g9 = adb_warp_sum(g9);
            if (lane == 0) {
                float* a = sAcc[t];
                atomicAdd(a + 0, g0); atomicAdd(a + 1, g1); atomicAdd(a + 2, g2); atomicAdd(a + 3, g3);
                atomicAdd(a + 4, g4); atomicAdd(a + 5, g5); atomicAdd(a + 6, g6); atomicAdd(a + 7, g7);
                atomicAdd(a + 8, g8); atomicAdd(a + 9, g9);
            }
        }
        __syncthreads();
        if (tid < bsize) {
            const float* a = sAcc[tid];
            float* dst = v_splats + (size_t)sG[tid] * ADB_SPLAT_STRIDE;
#pragma unroll
            for (int k = 0; k < 10; ++k)
                if (a[k] != 0.f) atomicAdd(dst + k, a[k]);
        }
    }
}

}  // namespace

// `vals` are the sorted values (cam*N + gaussian); n_per_cam = N.  colors [H,W,4], alphas [H,W], last_ids [H,W].
ADB_API int adb_raster_blend_fwd(int W, int H, int n_per_cam, const float* splats, const int32_t* vals_sorted,
                                 const int32_t* tile_offsets, float* colors, float* alphas, int32_t* last_ids,
                                 cudaStream_t stream) {
    ADB_REQUIRE(W > 0 && H > 0 && n_per_cam >= 0, "adb_raster_blend_fwd: bad sizes");
    ADB_REQUIRE(tile_offsets && colors && alphas && last_ids, "adb_raster_blend_fwd: null pointer");
    dim3 grid(adb_cdiv(W, ADB_TILE), adb_cdiv(H, ADB_TILE));
    blend_fwd_kernel<<<grid, BLOCK, 0, stream>>>(W, H, splats, vals_sorted, tile_offsets, max(n_per_cam, 1), colors,
                                                alphas, last_ids);
    ADB_CHECK_LAUNCH("blend_fwd_kernel");
    return ADB_OK;
}

// v_splats [N,12] must be zeroed by the caller; gradients are accumulated into it.
ADB_API int adb_raster_blend_bwd(int W, int H, int n_per_cam, const float* splats, const int32_t* vals_sorted,
                                 const int32_t* tile_offsets, const float* alphas, const int32_t* last_ids,
                                 const float* v_colors, const float* v_alphas, float* v_splats,
                                 cudaStream_t stream) {
    ADB_REQUIRE(W > 0 && H > 0 && n_per_cam >= 0, "adb_raster_blend_bwd: bad sizes");
    ADB_REQUIRE(tile_offsets && alphas && last_ids && v_colors && v_alphas, "adb_raster_blend_bwd: null pointer");
    if (n_per_cam == 0) return ADB_OK;
    ADB_REQUIRE(v_splats, "adb_raster_blend_bwd: null v_splats");
    dim3 grid(adb_cdiv(W, ADB_TILE), adb_cdiv(H, ADB_TILE));
    blend_bwd_kernel<<<grid, BLOCK, 0, stream>>>(W, H, splats, vals_sorted, tile_offsets, n_per_cam, alphas, last_ids,
                                                v_colors, v_alphas, v_splats);
    ADB_CHECK_LAUNCH("blend_bwd_kernel");
    return ADB_OK;
}
