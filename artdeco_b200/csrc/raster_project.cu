// Per-Gaussian projection + SH colour + tile count, fused into one streaming pass (compiled with
// -fmad=false: the values that feed the integer tile keys must be bit-identical to the C oracle).
//
// Replaces what the reference reaches through gsplat.rendering.rasterization at
// Reconstruct/scene/scene_models/h3dgsv3.py:664-680: the projection, the degree-3 SH
// evaluation with the +0.5 / clamp (App. B.2) and the tile count of isect_tiles (App. B.3).  gsplat runs
// these as three kernels with [N]-sized round trips through HBM between them; here one thread per
// Gaussian reads the 236 B of parameters once and writes one 48 B splat record + radii + count.
#include "raster_common.cuh"
#include "adb_detmath.h"

namespace {

__device__ __forceinline__ void mat3_mul(const float* A, const float* B, float* C) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            C[i * 3 + j] = A[i * 3 + 0] * B[0 * 3 + j] + A[i * 3 + 1] * B[1 * 3 + j] + A[i * 3 + 2] * B[2 * 3 + j];
}
__device__ __forceinline__ void mat3_mul_bt(const float* A, const float* B, float* C) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            C[i * 3 + j] = A[i * 3 + 0] * B[j * 3 + 0] + A[i * 3 + 1] * B[j * 3 + 1] + A[i * 3 + 2] * B[j * 3 + 2];
}

__device__ __forceinline__ void sh_basis(int deg, float x, float y, float z, float* B) {
#pragma unroll
    for (int k = 0; k < 16; ++k) B[k] = 0.f;
    B[0] = 0.2820947917738781f;
    if (deg < 1) return;
    B[1] = -0.48860251190292f * y; B[2] = 0.48860251190292f * z; B[3] = -0.48860251190292f * x;
    if (deg < 2) return;
    float z2 = z * z, fT0B = -1.092548430592079f * z, fC1 = x * x - y * y, fS1 = 2.f * x * y;
    B[4] = 0.5462742152960395f * fS1; B[5] = fT0B * y; B[6] = 0.9461746957575601f * z2 - 0.3153915652525201f;
    B[7] = fT0B * x; B[8] = 0.5462742152960395f * fC1;
    if (deg < 3) return;
    float fT0C = -2.285228997322329f * z2 + 0.4570457994644658f, fT1B = 1.445305721320277f * z;
    float fC2 = x * fC1 - y * fS1, fS2 = x * fS1 + y * fC1;
    B[9] = -0.5900435899266435f * fS2; B[10] = fT1B * fS1; B[11] = fT0C * y;
    B[12] = z * (1.865881662950577f * z2 - 1.119528997770346f); B[13] = fT0C * x; B[14] = fT1B * fC1;
    B[15] = -0.5900435899266435f * fC2;
}

__global__ void __launch_bounds__(256)
project_fwd_kernel(int N, const float* __restrict__ means, const float* __restrict__ quats,
                   const float* __restrict__ scales, const float* __restrict__ opacities,
                   const float* __restrict__ sh, int sh_degree, AdbCam cam,
                   int32_t* __restrict__ radii, float* __restrict__ splats, int32_t* __restrict__ tiles_per_gauss,
                   int32_t* __restrict__ tile_counts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    int tr_x0 = 0, tr_x1 = 0, tr_y0 = 0, tr_y1 = 0;    // tile rectangle (for the fused per-tile counting)
    const float* V = cam.viewmat;
    const float R[9] = {V[0], V[1], V[2], V[4], V[5], V[6], V[8], V[9], V[10]};
    const float t[3] = {V[3], V[7], V[11]};
    const float fx = cam.K[0], fy = cam.K[4], cx = cam.K[2], cy = cam.K[5];
    const float m0 = means[3 * i], m1 = means[3 * i + 1], m2 = means[3 * i + 2];

    int rx_i = 0, ry_i = 0, count = 0, cull_rx = 0, cull_ry = 0;
    float u = 0.f, v = 0.f, ca = 0.f, cb = 0.f, cc = 0.f, lg = 0.f;
    const float x = R[0] * m0 + R[1] * m1 + R[2] * m2 + t[0];
    const float y = R[3] * m0 + R[4] * m1 + R[5] * m2 + t[1];
    const float z = R[6] * m0 + R[7] * m1 + R[8] * m2 + t[2];
    const float opacity = opacities[i];
    bool ok = !(z < cam.near_plane || z > cam.far_plane);
    if (ok) {
        float Rq[9];
        {
            float4 q = reinterpret_cast<const float4*>(quats)[i];
            float w = q.x, qx = q.y, qy = q.z, qz = q.w;
            float n2 = w * w + qx * qx + qy * qy + qz * qz;
            float inv = 1.0f / sqrtf(n2);
            w = w * inv; qx = qx * inv; qy = qy * inv; qz = qz * inv;
            float x2 = qx * qx, y2 = qy * qy, z2 = qz * qz;
            float xy = qx * qy, xz = qx * qz, yz = qy * qz, wx = w * qx, wy = w * qy, wz = w * qz;
            Rq[0] = 1.0f - 2.0f * (y2 + z2); Rq[1] = 2.0f * (xy - wz);        Rq[2] = 2.0f * (xz + wy);
            Rq[3] = 2.0f * (xy + wz);        Rq[4] = 1.0f - 2.0f * (x2 + z2); Rq[5] = 2.0f * (yz - wx);
            Rq[6] = 2.0f * (xz - wy);        Rq[7] = 2.0f * (yz + wx);        Rq[8] = 1.0f - 2.0f * (x2 + y2);
        }
        const float s[3] = {scales[3 * i], scales[3 * i + 1], scales[3 * i + 2]};
        float M[9], Sigma[9], RS[9], Sc[9];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) M[a * 3 + b] = Rq[a * 3 + b] * s[b];
        mat3_mul_bt(M, M, Sigma);
        mat3_mul(R, Sigma, RS);
        mat3_mul_bt(RS, R, Sc);

        const float W = (float)cam.W, H = (float)cam.H;
        float tanx = 0.5f * W / fx, tany = 0.5f * H / fy;
        float lxp = (W - cx) / fx + 0.3f * tanx, lxn = cx / fx + 0.3f * tanx;
        float lyp = (H - cy) / fy + 0.3f * tany, lyn = cy / fy + 0.3f * tany;
        float rz = 1.0f / z;
        float rz2 = rz * rz;
        float xr = x * rz, yr = y * rz;
        float tx = z * fminf(lxp, fmaxf(-lxn, xr));
        float ty = z * fminf(lyp, fmaxf(-lyn, yr));
        float J0 = fx * rz, J2 = -fx * tx * rz2, J4 = fy * rz, J5 = -fy * ty * rz2;
        float k00 = J0 * Sc[0] + J2 * Sc[6], k01 = J0 * Sc[1] + J2 * Sc[7], k02 = J0 * Sc[2] + J2 * Sc[8];
        float k11 = J4 * Sc[4] + J5 * Sc[7], k12 = J4 * Sc[5] + J5 * Sc[8];
        float a = k00 * J0 + k02 * J2;
        float b = k01 * J4 + k02 * J5;
        float c = k11 * J4 + k12 * J5;
        a = a + cam.eps2d;
        c = c + cam.eps2d;
        float det = a * c - b * b;
        if (cam.convention == ADB_CONV_INRIA) {
            ok = det > 0.f;
            if (ok) {
                u = fx * x * rz + cx;
                v = fy * y * rz + cy;
                const float mid = 0.5f * (a + c);
                const float lam1 = mid + sqrtf(fmaxf(0.1f, mid * mid - det));
                const int rI = (int)fminf(ceilf(3.0f * sqrtf(lam1)), 1.0e9f);
                int x0, x1, y0, y1;
                adb_tile_rect(u, v, rI, rI, cam.W, cam.H, ADB_CONV_INRIA, x0, x1, y0, y1);
                count = (x1 - x0) * (y1 - y0);
                tr_x0 = x0; tr_x1 = x1; tr_y0 = y0; tr_y1 = y1;
                ok = count > 0;          // Inria leaves radii = 0 when no tile is touched
                if (ok) {
                    rx_i = ry_i = rI;
                    ca = c / det; cb = -b / det; cc = a / det;
                    if (opacity < ADB_ALPHA_THRESHOLD) {
                        // alpha can never reach 1/255: keeps its radius (visibility filter) and a record (so the
                        // backward reads defined values) but emits no keys
                        count = 0;
                    } else {
                        // conservative extent for the blend kernels' block-level culling (alpha < 1/255 outside it)
                        lg = adb_det_logf(opacity / ADB_ALPHA_THRESHOLD);
                        const float ext = fminf(3.33f, sqrtf(2.0f * lg));
                        const float r1 = ext * sqrtf(lam1);
                        cull_rx = (int)fminf(ceilf(fminf(ext * sqrtf(a), r1)), 65535.f);
                        cull_ry = (int)fminf(ceilf(fminf(ext * sqrtf(c), r1)), 65535.f);
                    }
                }
            }
        } else {
        ok = det > 0.f && !(opacity < ADB_ALPHA_THRESHOLD);
        if (ok) {
            u = fx * x * rz + cx;
            v = fy * y * rz + cy;
            lg = adb_det_logf(opacity / ADB_ALPHA_THRESHOLD);
            float ext = sqrtf(2.0f * lg);
            ext = fminf(3.33f, ext);
            float bb = 0.5f * (a + c);
            float lam = bb + sqrtf(fmaxf(0.01f, bb * bb - det));
            float r1 = ext * sqrtf(lam);
            float rx = ceilf(fminf(ext * sqrtf(a), r1));
            float ry = ceilf(fminf(ext * sqrtf(c), r1));
            ok = !(rx <= cam.radius_clip && ry <= cam.radius_clip) &&
                 !(u + rx <= 0.f || u - rx >= W || v + ry <= 0.f || v - ry >= H);
            if (ok) {
                rx_i = (int)rx; ry_i = (int)ry;
                cull_rx = rx_i; cull_ry = ry_i;
                ca = c / det; cb = -b / det; cc = a / det;
                // tile count (same arithmetic as the emit kernel and the oracle's tile_bounds)
                int x0, x1, y0, y1;
                adb_tile_rect(u, v, rx_i, ry_i, cam.W, cam.H, ADB_CONV_GSPLAT, x0, x1, y0, y1);
                count = (x1 - x0) * (y1 - y0);
                tr_x0 = x0; tr_x1 = x1; tr_y0 = y0; tr_y1 = y1;
            }
        }
        }
    }
    reinterpret_cast<int2*>(radii)[i] = make_int2(rx_i, ry_i);
    tiles_per_gauss[i] = count;
    if (tile_counts && count > 0) {
        // fused first stage of the tile-bucketed intersection (raster_isect.cu): fire-and-forget REDs into the replicated
        // per-tile counters, overlapped with the SH loads below instead of a separate pass over radii / means2d
        const int tw = (cam.W + ADB_TILE - 1) / ADB_TILE;
        const int cp = blockIdx.x & (ADB_TILE_COUNTER_COPIES - 1);
        for (int ty = tr_y0; ty < tr_y1; ++ty)
            for (int tx = tr_x0; tx < tr_x1; ++tx) atomicAdd(tile_counts + (ty * tw + tx) * ADB_TILE_COUNTER_COPIES + cp, 1);
    }
    if (!ok) return;

    float r = 0.f, g = 0.f, bl = 0.f;
    if (sh) {
        float dx = m0 - cam.campos[0], dy = m1 - cam.campos[1], dz = m2 - cam.campos[2];
        float inv = 1.0f / sqrtf(dx * dx + dy * dy + dz * dz);
        float B[16];
        sh_basis(sh_degree, dx * inv, dy * inv, dz * inv, B);
        const float4* sp = reinterpret_cast<const float4*>(sh + (size_t)i * 48);
        float c48[48];
#pragma unroll
        for (int k = 0; k < 12; ++k) {
            float4 q = adb_ldg_stream4(sp + k);
            c48[4 * k] = q.x; c48[4 * k + 1] = q.y; c48[4 * k + 2] = q.z; c48[4 * k + 3] = q.w;
        }
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            r += B[k] * c48[3 * k]; g += B[k] * c48[3 * k + 1]; bl += B[k] * c48[3 * k + 2];
        }
        r = fmaxf(r + 0.5f, 0.f); g = fmaxf(g + 0.5f, 0.f); bl = fmaxf(bl + 0.5f, 0.f);
    }
    float4* out = reinterpret_cast<float4*>(splats + (size_t)i * ADB_SPLAT_STRIDE);
    // sigma <= ln(255*opacity) <=> alpha >= 1/255; the margin keeps the pre-test a strict superset of the exact test
    const float sigma_max = lg + 0.02f;
    const unsigned packed_radii = (unsigned)min(cull_rx, 65535) | ((unsigned)min(cull_ry, 65535) << 16);
    out[0] = make_float4(u, v, ca, cb);
    out[1] = make_float4(cc, opacity, sigma_max, __uint_as_float(packed_radii));
    out[2] = make_float4(r, g, bl, z);
}

}  // namespace

// One camera.  `sh` may be NULL (then rgb = 0 and campos is ignored).  viewmat/K/campos are DEVICE pointers.
static int project_fwd_impl(int convention, int N, const float* means, const float* quats, const float* scales,
                            const float* opacities, const float* sh, int sh_degree,
                            const float* viewmat, const float* K, const float* campos, int W, int H,
                            float eps2d, float near_plane, float far_plane, float radius_clip,
                            int32_t* radii, float* splats, int32_t* tiles_per_gauss, int32_t* tile_counts,
                            cudaStream_t stream) {
    ADB_REQUIRE(N >= 0 && W > 0 && H > 0, "adb_raster_project_fwd: bad sizes");
    if (N == 0) return ADB_OK;
    ADB_REQUIRE(means && quats && scales && opacities && viewmat && K && radii && splats && tiles_per_gauss,
                "adb_raster_project_fwd: null pointer");
    ADB_REQUIRE(!sh || campos, "adb_raster_project_fwd: sh needs campos");
    ADB_REQUIRE(sh_degree >= 0 && sh_degree <= 3, "adb_raster_project_fwd: sh_degree must be 0..3");
    AdbCam cam{viewmat, K, campos, W, H, eps2d, near_plane, far_plane, radius_clip, convention};
    project_fwd_kernel<<<adb_cdiv(N, 256), 256, 0, stream>>>(N, means, quats, scales, opacities, sh, sh_degree, cam,
                                                            radii, splats, tiles_per_gauss, tile_counts);
    ADB_CHECK_LAUNCH("project_fwd_kernel");
    return ADB_OK;
}

ADB_API int adb_raster_project_fwd(int N, const float* means, const float* quats, const float* scales,
                                   const float* opacities, const float* sh, int sh_degree,
                                   const float* viewmat, const float* K, const float* campos, int W, int H,
                                   float eps2d, float near_plane, float far_plane, float radius_clip,
                                   int32_t* radii, float* splats, int32_t* tiles_per_gauss, cudaStream_t stream) {
    return project_fwd_impl(ADB_CONV_GSPLAT, N, means, quats, scales, opacities, sh, sh_degree, viewmat, K, campos, W, H,
                            eps2d, near_plane, far_plane, radius_clip, radii, splats, tiles_per_gauss, nullptr, stream);
}

// Same, plus the first stage of the tile-bucketed intersection fused in: tile_counts (int32 [2*4*T], first half zero on entry,
// see adb_raster_tile_count_scan) receives one RED per (Gaussian, touched tile); pass counts_ready = 1 to the scan afterwards.
ADB_API int adb_raster_project_fwd_counts(int N, const float* means, const float* quats, const float* scales,
                                          const float* opacities, const float* sh, int sh_degree,
                                          const float* viewmat, const float* K, const float* campos, int W, int H,
                                          float eps2d, float near_plane, float far_plane, float radius_clip,
                                          int32_t* radii, float* splats, int32_t* tiles_per_gauss, int32_t* tile_counts,
                                          cudaStream_t stream) {
    ADB_REQUIRE(tile_counts, "adb_raster_project_fwd_counts: null tile_counts");
    return project_fwd_impl(ADB_CONV_GSPLAT, N, means, quats, scales, opacities, sh, sh_degree, viewmat, K, campos, W, H,
                            eps2d, near_plane, far_plane, radius_clip, radii, splats, tiles_per_gauss, tile_counts, stream);
}

// Legacy (Inria / diff_gaussian_rasterization) conventions: radii[:,0] == radii[:,1] == ceil(3 sqrt(lambda_max)); the
// caller passes eps2d = 0.3 and near_plane = 0.2.  Same record layout (slot 11 stays the camera-space depth z).
ADB_API int adb_raster_project_fwd_legacy(int N, const float* means, const float* quats, const float* scales,
                                          const float* opacities, const float* sh, int sh_degree,
                                          const float* viewmat, const float* K, const float* campos, int W, int H,
                                          float eps2d, float near_plane, float far_plane,
                                          int32_t* radii, float* splats, int32_t* tiles_per_gauss, cudaStream_t stream) {
    return project_fwd_impl(ADB_CONV_INRIA, N, means, quats, scales, opacities, sh, sh_degree, viewmat, K, campos, W, H,
                            eps2d, near_plane, far_plane, 0.f, radii, splats, tiles_per_gauss, nullptr, stream);
}
