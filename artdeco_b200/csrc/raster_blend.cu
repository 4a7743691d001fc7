// Per-tile alpha blending, forward and backward (RGB + depth = 4 channels).
//
// Replaces gsplat's rasterize_to_pixels fwd/bwd that the reference reaches through
// Reconstruct/scene/scene_models/h3dgsv3.py:664-680:
//   pixel centre (j+0.5, i+0.5); sigma = .5(a dx^2 + c dy^2) + b dx dy; alpha = min(.999, o exp(-sigma));
//   skip if sigma<0 or alpha<1/255; stop when T(1-alpha) <= 1e-4; out += feat*alpha*T; alpha_out = 1-T.
//
// B200-first structure (one CTA per 16x16 tile, 8 warps, each warp owns an 8x4 pixel block):
//  * the tile's slice of the depth-sorted list is staged in shared memory 256 splats at a time, each thread
//    gathering one 48 B record with 3 coalesced-in-record LDG.128;
//  * WARP-COOPERATIVE CULLING: lane l tests splat l of a 32-chunk against the warp's 8x4 block with the
//    splat's own integer radii (outside them alpha < 1/255 by construction of the radius), a ballot gives
//    the warp its private hit list, and only those splats are evaluated (LDS.128 broadcast reads);
//  * a lane evaluates exp() only when sigma <= ln(255*opacity)+margin (stored in the record), i.e. only when
//    alpha can reach 1/255 — the exact alpha test still runs inside, so results equal the plain algorithm;
//  * backward: per-splat gradients (10 values) are reduced over the warp with a packed 12-shuffle
//    butterfly (instead of 10 x 5), combined across the 8 warps with one shared-memory RED per value, and
//    flushed with three 128-bit vector REDs per (tile, splat) (red.global.add.v4.f32; gsplat issues 10
//    scalar atomics per WARP).
#include "raster_common.cuh"
#include <stdlib.h>

namespace {

constexpr int BLOCK = ADB_TILE * ADB_TILE;  // 256
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ void gather_splat(const float* __restrict__ splats, int g, float4& A, float4& B, float4& C) {
    const float4* p = reinterpret_cast<const float4*>(splats + (size_t)g * ADB_SPLAT_STRIDE);
    A = __ldg(p);
    B = __ldg(p + 1);
    C = __ldg(p + 2);
}

// warp block geometry: warp w covers pixels x in [bx*16 + (w&1)*8, +8), y in [by*16 + (w>>1)*4, +4)
struct WarpRect {
    float xlo, xhi, ylo, yhi;  // pixel-centre range
};

// Can splat (A,B) reach alpha >= 1/255 anywhere in the warp's block?  Two conservative tests:
//  1. the integer-radius box written by the projection (outside it alpha < 1/255 by construction of the radius);
//  2. the exact minimum of sigma(d) = .5(a dx^2 + c dy^2) + b dx dy over the block's pixel-centre rectangle against the
//     record's sigma_max = ln(255 o) + margin — the same bound the per-pixel pre-test uses, so a culled splat would have
//     failed that pre-test on every pixel of the block and the image is unchanged.  For anisotropic splats the box is
//     several times larger than the ellipse; this test costs ~1 warp-instruction per (warp, splat) because 32 splats
//     are tested per instruction, against ~35 (forward) / ~110 (backward) for an evaluation it avoids.
__device__ __forceinline__ bool splat_hits(const float4& A, const float4& B, const WarpRect& r) {
    const unsigned pr = __float_as_uint(B.w);
    // 65535 is the saturation value written by the projection: treat it as unbounded
    const float rx = (pr & 0xffffu) == 0xffffu ? 3.0e38f : (float)(pr & 0xffffu);
    const float ry = (pr >> 16) == 0xffffu ? 3.0e38f : (float)(pr >> 16);
    if (!((A.x + rx >= r.xlo) && (A.x - rx <= r.xhi) && (A.y + ry >= r.ylo) && (A.y - ry <= r.yhi))) return false;
#ifndef ADB_NO_TIGHT_CULL
    const float x0 = r.xlo - A.x, x1 = r.xhi - A.x, y0 = r.ylo - A.y, y1 = r.yhi - A.y;
    if (x0 <= 0.f && x1 >= 0.f && y0 <= 0.f && y1 >= 0.f) return true;   // centre inside the block
    const float a = A.z, b = A.w, c = B.x;
    const float nb_c = -__fdividef(b, c), nb_a = -__fdividef(b, a);
    // minimum over each edge (1-D quadratic, minimiser clamped to the edge); the rectangle's minimum is on its boundary
    float m;
    {
        const float dy0 = fminf(y1, fmaxf(y0, nb_c * x0)), dy1 = fminf(y1, fmaxf(y0, nb_c * x1));
        const float q0 = 0.5f * (a * x0 * x0 + c * dy0 * dy0) + b * x0 * dy0;
        const float q1 = 0.5f * (a * x1 * x1 + c * dy1 * dy1) + b * x1 * dy1;
        m = fminf(q0, q1);
    }
    {
        const float dx0 = fminf(x1, fmaxf(x0, nb_a * y0)), dx1 = fminf(x1, fmaxf(x0, nb_a * y1));
        const float q0 = 0.5f * (a * dx0 * dx0 + c * y0 * y0) + b * dx0 * y0;
        const float q1 = 0.5f * (a * dx1 * dx1 + c * y1 * y1) + b * dx1 * y1;
        m = fminf(m, fminf(q0, q1));
    }
    // rounding slack: the per-pixel test is sigma <= sigma_max in the same fp32 arithmetic
    return m <= B.z * 1.0001f + 1e-4f;
#else
    return true;
#endif
}

// LEGACY = Inria conventions (ADB_CONV_INRIA): alpha <= 0.99, stop when T(1-alpha) < 1e-4 (strict), 4th channel
// accumulates 1/z, and main_ids gets the Gaussian with the largest blending weight alpha*T per pixel (-1: none).
template <bool LEGACY>
__global__ void __launch_bounds__(BLOCK)
blend_fwd_kernel(int W, int H, const float* __restrict__ splats, const int32_t* __restrict__ vals,
                 const int32_t* __restrict__ tile_offsets, int n_per_cam,
                 float* __restrict__ colors, float* __restrict__ alphas, int32_t* __restrict__ last_ids,
                 int32_t* __restrict__ main_ids) {
    constexpr float MAXA = LEGACY ? ADB_MAX_ALPHA_INRIA : ADB_MAX_ALPHA;
    __shared__ float4 sA[BLOCK];
    __shared__ float4 sB[BLOCK];
    __shared__ float4 sC[BLOCK];
    __shared__ unsigned char sList[BLOCK / 32][BLOCK];

    const int tw = (W + ADB_TILE - 1) / ADB_TILE;
    const int tile = blockIdx.y * tw + blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int x0 = blockIdx.x * ADB_TILE + (warp & 1) * 8, y0 = blockIdx.y * ADB_TILE + (warp >> 1) * 4;
    const int j = x0 + (lane & 7), i = y0 + (lane >> 3);
    const bool inside = (i < H && j < W);
    const float px = (float)j + 0.5f, py = (float)i + 0.5f;
    const WarpRect rect{(float)x0 + 0.5f, (float)x0 + 7.5f, (float)y0 + 0.5f, (float)y0 + 3.5f};
    const int start = tile_offsets[tile], end = tile_offsets[tile + 1];

    float T = 1.0f;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int cur = 0;
    float best_w = 0.f;
    int best_k = -1;
    bool done = !inside;
    const int nb = (end - start + BLOCK - 1) / BLOCK;
    for (int b = 0; b < nb; ++b) {
        if (__syncthreads_count(done) >= BLOCK) break;
        const int bstart = start + b * BLOCK;
        const int idx = bstart + tid;
        if (idx < end) {
            gather_splat(splats, vals[idx] % n_per_cam, sA[tid], sB[tid], sC[tid]);
            if (LEGACY) sC[tid].w = 1.0f / sC[tid].w;
        }
        __syncthreads();
        const int bsize = min(BLOCK, end - bstart);
        if (__all_sync(FULL, done)) continue;
        // the warp's private hit list for this batch (ascending list order == front to back)
        int nhit = 0;
        for (int c0 = 0; c0 < bsize; c0 += 32) {
            const int s = c0 + lane;
            bool hit = false;
            if (s < bsize) hit = splat_hits(sA[s], sB[s], rect);
            const unsigned mask = __ballot_sync(FULL, hit);
            if (hit) sList[warp][nhit + __popc(mask & lt_mask)] = (unsigned char)s;
            nhit += __popc(mask);
        }
        __syncwarp();
        for (int k = 0; k < nhit; ++k) {
            const int t = sList[warp][k];
            const float4 A = sA[t];
            const float4 B = sB[t];
            const float dx = A.x - px, dy = A.y - py;
            const float sigma = 0.5f * (A.z * dx * dx + B.x * dy * dy) + A.w * dx * dy;
            if (!done && sigma >= 0.f && sigma <= B.z) {
                const float alpha = fminf(MAXA, B.y * __expf(-sigma));
                if (alpha >= ADB_ALPHA_THRESHOLD) {
                    const float nT = T * (1.0f - alpha);
                    if (LEGACY ? (nT < ADB_T_EPS) : (nT <= ADB_T_EPS)) {
                        done = true;
                    } else {
                        const float w = alpha * T;
                        const float4 C = sC[t];
                        acc.x += C.x * w; acc.y += C.y * w; acc.z += C.z * w; acc.w += C.w * w;
                        cur = bstart + t;
                        if (LEGACY && w > best_w) { best_w = w; best_k = cur; }
                        T = nT;
                    }
                }
            }
            if ((k & 7) == 7 && __all_sync(FULL, done)) break;
        }
        __syncwarp();
    }
    if (inside) {
        const size_t pix = (size_t)i * W + j;
        reinterpret_cast<float4*>(colors)[pix] = acc;
        alphas[pix] = 1.0f - T;
        last_ids[pix] = cur;
        if (LEGACY && main_ids) main_ids[pix] = best_k >= 0 ? vals[best_k] % n_per_cam : -1;
    }
}

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <bool DIRECT, bool LEGACY>
__global__ void __launch_bounds__(BLOCK)
blend_bwd_kernel(int W, int H, const float* __restrict__ splats, const int32_t* __restrict__ vals,
                 const int32_t* __restrict__ tile_offsets, int n_per_cam,
                 const float* __restrict__ alphas, const int32_t* __restrict__ last_ids,
                 const float* __restrict__ v_colors, const float* __restrict__ v_alphas,
                 float* __restrict__ v_splats) {
    constexpr float MAXA = LEGACY ? ADB_MAX_ALPHA_INRIA : ADB_MAX_ALPHA;
    __shared__ float4 sA[BLOCK];
    __shared__ float4 sB[BLOCK];
    __shared__ float4 sC[BLOCK];
    __shared__ int sG[BLOCK];
    __shared__ unsigned char sList[BLOCK / 32][BLOCK];
    __shared__ __align__(16) float sAcc[BLOCK][12];

    const int tw = (W + ADB_TILE - 1) / ADB_TILE;
    const int tile = blockIdx.y * tw + blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int x0 = blockIdx.x * ADB_TILE + (warp & 1) * 8, y0 = blockIdx.y * ADB_TILE + (warp >> 1) * 4;
    const int j = x0 + (lane & 7), i = y0 + (lane >> 3);
    const bool inside = (i < H && j < W);
    const float px = (float)j + 0.5f, py = (float)i + 0.5f;
    const WarpRect rect{(float)x0 + 0.5f, (float)x0 + 7.5f, (float)y0 + 0.5f, (float)y0 + 3.5f};
    const int start = tile_offsets[tile], end = tile_offsets[tile + 1];
    if (end <= start) return;

    const size_t pix = (size_t)(inside ? i : 0) * W + (inside ? j : 0);
    const float T_final = inside ? 1.0f - alphas[pix] : 1.0f;
    const int bin_final = inside ? last_ids[pix] : -1;
    const float4 vo = inside ? reinterpret_cast<const float4*>(v_colors)[pix] : make_float4(0.f, 0.f, 0.f, 0.f);
    const float va = inside ? v_alphas[pix] : 0.f;
    float T = T_final;
    float bv = 0.f;
    const float Tva = T_final * va;

    int warp_bin_final = bin_final;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) warp_bin_final = max(warp_bin_final, __shfl_xor_sync(FULL, warp_bin_final, o));

    // which of the 10 reduced components this lane ends up owning after the packed butterfly (-1: none)
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
    int my_comp = (b4 ? 5 : 0) + (b3 ? 3 : 0) + (b2 ? 2 : 0) + (b1 ? 1 : 0);
    if ((lane & 1) || (b2 && b1) || (b3 && b2)) my_comp = -1;

    const int nb = (end - start + BLOCK - 1) / BLOCK;
    for (int b = 0; b < nb; ++b) {
        __syncthreads();  // previous batch fully consumed / flushed
        const int batch_end = end - 1 - b * BLOCK;
        const int bsize = min(BLOCK, batch_end + 1 - start);
        const int idx = batch_end - tid;
        if (idx >= start) {
            const int g = vals[idx] % n_per_cam;
            sG[tid] = g;
            gather_splat(splats, g, sA[tid], sB[tid], sC[tid]);
            if (LEGACY) sC[tid].w = 1.0f / sC[tid].w;   // v_splats slot 9 is then dL/d(1/z)
        }
        if (!DIRECT) {
            float4* z = reinterpret_cast<float4*>(sAcc[tid]);
            z[0] = z[1] = z[2] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        __syncthreads();
        // (a warp whose last contributing splat lies before this whole batch has every `hit` false below)
        int nhit = 0;
        for (int c0 = 0; c0 < bsize; c0 += 32) {
            if (batch_end - (c0 + 31) > warp_bin_final) continue;  // chunk entirely behind this warp's last splat
            const int s = c0 + lane;
            bool hit = false;
            if (s < bsize && batch_end - s <= warp_bin_final) hit = splat_hits(sA[s], sB[s], rect);
            const unsigned mask = __ballot_sync(FULL, hit);
            if (hit) sList[warp][nhit + __popc(mask & lt_mask)] = (unsigned char)s;
            nhit += __popc(mask);
        }
        __syncwarp();
        {
            for (int k = 0; k < nhit; ++k) {
                const int t = sList[warp][k];
                const float4 A = sA[t];
                const float4 B = sB[t];
                const float dx = A.x - px, dy = A.y - py;
                const float sigma = 0.5f * (A.z * dx * dx + B.x * dy * dy) + A.w * dx * dy;
                bool valid = inside && (batch_end - t <= bin_final) && sigma >= 0.f && sigma <= B.z;
                float vis = 0.f, alpha = 0.f;
                if (valid) {
                    vis = __expf(-sigma);
                    alpha = fminf(MAXA, B.y * vis);
                    valid = alpha >= ADB_ALPHA_THRESHOLD;
                }
                if (!__any_sync(FULL, valid)) continue;
                // Branch-free from here: an invalid lane runs with alpha = vis = 0, which leaves T and bv unchanged
                // (ra = 1, fac = 0) and contributes exact zeros to every sum.
                alpha = valid ? alpha : 0.f;
                vis = valid ? vis : 0.f;
                float v[10];
                {
                    const float4 C = sC[t];
                    const float ra = __fdividef(1.0f, 1.0f - alpha);
                    T *= ra;
                    const float fac = alpha * T;
                    // cv = <feat, v_out>;  bv = sum over later splats of fac*cv  (the scalar the 4-channel
                    // "buffer" of the textbook backward collapses to once it is dotted with v_out)
                    const float cv = C.x * vo.x + C.y * vo.y + C.z * vo.z + C.w * vo.w;
                    const float v_alpha = cv * T + (Tva - bv) * ra;
                    bv += fac * cv;
                    v[6] = fac * vo.x; v[7] = fac * vo.y; v[8] = fac * vo.z; v[9] = fac * vo.w;
                    // d(alpha)/d(sigma) path is cut where alpha was clamped to 0.999
                    const float ov = B.y * vis;
                    const float v_sigma = ov <= MAXA ? -ov * v_alpha : 0.f;
                    // raw moments of v_sigma about the splat centre; project_bwd turns them into
                    // v_mean2d / v_conic / v_opacity once per Gaussian instead of once per pair
                    v[5] = v_sigma;
                    v[0] = v_sigma * dx;
                    v[1] = v_sigma * dy;
                    v[2] = v[0] * dx;
                    v[3] = v[0] * dy;
                    v[4] = v[1] * dy;
                }
                // packed butterfly: 10 values -> 5 -> 3 -> 2 -> 1 -> 1  (12 shuffles)
                float w0, w1, w2, w3, w4;
                {
                    float s0 = b4 ? v[0] : v[5], k0 = b4 ? v[5] : v[0];
                    float s1 = b4 ? v[1] : v[6], k1 = b4 ? v[6] : v[1];
                    float s2 = b4 ? v[2] : v[7], k2 = b4 ? v[7] : v[2];
                    float s3 = b4 ? v[3] : v[8], k3 = b4 ? v[8] : v[3];
                    float s4 = b4 ? v[4] : v[9], k4 = b4 ? v[9] : v[4];
                    w0 = k0 + __shfl_xor_sync(FULL, s0, 16);
                    w1 = k1 + __shfl_xor_sync(FULL, s1, 16);
                    w2 = k2 + __shfl_xor_sync(FULL, s2, 16);
                    w3 = k3 + __shfl_xor_sync(FULL, s3, 16);
                    w4 = k4 + __shfl_xor_sync(FULL, s4, 16);
                }
                float u0, u1, u2;
                {
                    float s0 = b3 ? w0 : w3, k0 = b3 ? w3 : w0;
                    float s1 = b3 ? w1 : w4, k1 = b3 ? w4 : w1;
                    float s2 = b3 ? w2 : 0.f, k2 = b3 ? 0.f : w2;
                    u0 = k0 + __shfl_xor_sync(FULL, s0, 8);
                    u1 = k1 + __shfl_xor_sync(FULL, s1, 8);
                    u2 = k2 + __shfl_xor_sync(FULL, s2, 8);
                }
                float t0, t1;
                {
                    float s0 = b2 ? u0 : u2, k0 = b2 ? u2 : u0;
                    float s1 = b2 ? u1 : 0.f, k1 = b2 ? 0.f : u1;
                    t0 = k0 + __shfl_xor_sync(FULL, s0, 4);
                    t1 = k1 + __shfl_xor_sync(FULL, s1, 4);
                }
                float r = (b1 ? t1 : t0) + __shfl_xor_sync(FULL, b1 ? t0 : t1, 2);
                r += __shfl_xor_sync(FULL, r, 1);
                if (my_comp >= 0) {
                    if (DIRECT) atomicAdd(v_splats + (size_t)sG[t] * ADB_SPLAT_STRIDE + my_comp, r);
                    else atomicAdd(&sAcc[t][my_comp], r);
                }
            }
        }
        if (DIRECT) continue;
        __syncthreads();
        if (tid < bsize) {
            const float4* a = reinterpret_cast<const float4*>(sAcc[tid]);
            const float4 q0 = a[0], q1 = a[1], q2 = a[2];
            float* dst = v_splats + (size_t)sG[tid] * ADB_SPLAT_STRIDE;
            if (q0.x != 0.f || q0.y != 0.f || q0.z != 0.f || q0.w != 0.f) red_add_v4(dst, q0.x, q0.y, q0.z, q0.w);
            if (q1.x != 0.f || q1.y != 0.f || q1.z != 0.f || q1.w != 0.f) red_add_v4(dst + 4, q1.x, q1.y, q1.z, q1.w);
            if (q2.x != 0.f || q2.y != 0.f) red_add_v4(dst + 8, q2.x, q2.y, 0.f, 0.f);
        }
    }
}

}  // namespace

static int blend_fwd_impl(bool legacy, int W, int H, int n_per_cam, const float* splats, const int32_t* vals_sorted,
                          const int32_t* tile_offsets, float* colors, float* alphas, int32_t* last_ids,
                          int32_t* main_ids, cudaStream_t stream) {
    ADB_REQUIRE(W > 0 && H > 0 && n_per_cam >= 0, "adb_raster_blend_fwd: bad sizes");
    ADB_REQUIRE(tile_offsets && colors && alphas && last_ids, "adb_raster_blend_fwd: null pointer");
    dim3 grid(adb_cdiv(W, ADB_TILE), adb_cdiv(H, ADB_TILE));
    if (legacy)
        blend_fwd_kernel<true><<<grid, BLOCK, 0, stream>>>(W, H, splats, vals_sorted, tile_offsets, max(n_per_cam, 1),
                                                          colors, alphas, last_ids, main_ids);
    else
        blend_fwd_kernel<false><<<grid, BLOCK, 0, stream>>>(W, H, splats, vals_sorted, tile_offsets, max(n_per_cam, 1),
                                                           colors, alphas, last_ids, nullptr);
    ADB_CHECK_LAUNCH("blend_fwd_kernel");
    return ADB_OK;
}

// `vals` are the sorted values (cam*N + gaussian); n_per_cam = N.  colors [H,W,4], alphas [H,W], last_ids [H,W].
ADB_API int adb_raster_blend_fwd(int W, int H, int n_per_cam, const float* splats, const int32_t* vals_sorted,
                                 const int32_t* tile_offsets, float* colors, float* alphas, int32_t* last_ids,
                                 cudaStream_t stream) {
    return blend_fwd_impl(false, W, H, n_per_cam, splats, vals_sorted, tile_offsets, colors, alphas, last_ids, nullptr,
                          stream);
}

// Legacy (Inria) blending: colors[...,3] = sum alpha*T/z (inverse depth); main_ids [H,W] may be NULL.
ADB_API int adb_raster_blend_fwd_legacy(int W, int H, int n_per_cam, const float* splats, const int32_t* vals_sorted,
                                        const int32_t* tile_offsets, float* colors, float* alphas, int32_t* last_ids,
                                        int32_t* main_ids, cudaStream_t stream) {
    return blend_fwd_impl(true, W, H, n_per_cam, splats, vals_sorted, tile_offsets, colors, alphas, last_ids, main_ids,
                          stream);
}

static int blend_bwd_impl(bool legacy, int W, int H, int n_per_cam, const float* splats, const int32_t* vals_sorted,
                          const int32_t* tile_offsets, const float* alphas, const int32_t* last_ids,
                          const float* v_colors, const float* v_alphas, float* v_splats, cudaStream_t stream) {
    ADB_REQUIRE(W > 0 && H > 0 && n_per_cam >= 0, "adb_raster_blend_bwd: bad sizes");
    ADB_REQUIRE(tile_offsets && alphas && last_ids && v_colors && v_alphas, "adb_raster_blend_bwd: null pointer");
    if (n_per_cam == 0) return ADB_OK;
    ADB_REQUIRE(v_splats, "adb_raster_blend_bwd: null v_splats");
    dim3 grid(adb_cdiv(W, ADB_TILE), adb_cdiv(H, ADB_TILE));
    static const int mode = getenv("ADB_BWD_MODE") ? atoi(getenv("ADB_BWD_MODE")) : 0;
    if (legacy)
        blend_bwd_kernel<false, true><<<grid, BLOCK, 0, stream>>>(W, H, splats, vals_sorted, tile_offsets, n_per_cam,
                                                                 alphas, last_ids, v_colors, v_alphas, v_splats);
    else if (mode == 1)
        blend_bwd_kernel<true, false><<<grid, BLOCK, 0, stream>>>(W, H, splats, vals_sorted, tile_offsets, n_per_cam,
                                                                 alphas, last_ids, v_colors, v_alphas, v_splats);
    else
        blend_bwd_kernel<false, false><<<grid, BLOCK, 0, stream>>>(W, H, splats, vals_sorted, tile_offsets, n_per_cam,
                                                                  alphas, last_ids, v_colors, v_alphas, v_splats);
    ADB_CHECK_LAUNCH("blend_bwd_kernel");
    return ADB_OK;
}

// v_splats [N,12] must be zeroed by the caller; gradients are accumulated into it.
ADB_API int adb_raster_blend_bwd(int W, int H, int n_per_cam, const float* splats, const int32_t* vals_sorted,
                                 const int32_t* tile_offsets, const float* alphas, const int32_t* last_ids,
                                 const float* v_colors, const float* v_alphas, float* v_splats,
                                 cudaStream_t stream) {
    return blend_bwd_impl(false, W, H, n_per_cam, splats, vals_sorted, tile_offsets, alphas, last_ids, v_colors,
                          v_alphas, v_splats, stream);
}

// Legacy (Inria) conventions; v_colors[...,3] is the upstream gradient of the inverse-depth channel and v_splats slot 9
// receives dL/d(1/z) per Gaussian (the caller multiplies by -1/z^2 before adb_raster_project_bwd).
ADB_API int adb_raster_blend_bwd_legacy(int W, int H, int n_per_cam, const float* splats, const int32_t* vals_sorted,
                                        const int32_t* tile_offsets, const float* alphas, const int32_t* last_ids,
                                        const float* v_colors, const float* v_alphas, float* v_splats,
                                        cudaStream_t stream) {
    return blend_bwd_impl(true, W, H, n_per_cam, splats, vals_sorted, tile_offsets, alphas, last_ids, v_colors,
                          v_alphas, v_splats, stream);
}
