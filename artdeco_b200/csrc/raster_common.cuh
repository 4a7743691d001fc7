// Shared definitions for the Gaussian-splat rasterizer kernels.
//
// HBM layout (all fp32 unless noted), one camera per call:
//   inputs   means[N,3] quats[N,4] (wxyz) scales[N,3] opacities[N] sh[N,16,3]
//            viewmat[16] (row-major world->camera), K[9], campos[3]   -- DEVICE pointers (no host sync)
//   splats   [N,12] packed per-Gaussian record written by the projection and gathered (3x LDG.128) by the
//            blend kernels:  0 mx | 1 my | 2 conic_a | 3 conic_b |
//                            4 conic_c | 5 opacity | 6 sigma_max = ln(255*opacity)+0.02 | 7 bits: rx | ry<<16 |
//                            8 r  | 9 g  | 10 b | 11 depth
//   radii    int32[N,2]   (0,0) == culled; culled Gaussians have no splat record and emit no keys
//   keys     int64[I]  (cam << (32+tile_bits)) | (tile << 32) | float_bits(depth);   vals int32[I] = cam*N + gaussian
//   tile_offsets int32[T+1]  first sorted index of each tile, [T] = I
//   colors   [H,W,4] = (r,g,b, sum z*alpha*T);  alphas [H,W];  last_ids int32[H,W]
//   v_splats [N,12] per-Gaussian accumulators written by blend_bwd (zeroed by the caller first):
//            0 M1x | 1 M1y | 2 M2xx | 3 M2xy | 4 M2yy | 5 M0   raw moments sum_p v_sigma * {dx,dy,dx^2,dxdy,dy^2,1}
//            (project_bwd converts them to v_mean2d / v_conic / v_opacity) | 6..8 v_rgb | 9 v_depth | 10,11 unused.
#pragma once
#include "common.cuh"

#define ADB_TILE 16
#define ADB_SPLAT_STRIDE 12
#define ADB_ALPHA_THRESHOLD (1.0f / 255.0f)
#define ADB_MAX_ALPHA 0.999f
#define ADB_T_EPS 1e-4f
#define ADB_TILE_COUNTER_COPIES 4   // replicas of every tile counter in the bucketed intersection (raster_isect.cu)
#define ADB_SIGMA_MARGIN 0.02f   // splat record slot 6 = ln(255*opacity) + this (raster_project.cu)
// Rendering conventions.  GSPLAT: what the live path uses (h3dgsv3.py:664-680).  INRIA: the legacy
// diff_gaussian_rasterization contract of the web viewer (Reconstruct/webviewer/scene_models.py:559-605):
// radius ceil(3 sqrt(lambda_max)) with a 0.1 floor under the root, tile rectangle on pixel-index coordinates
// ((int)((p-r)/16) .. (int)((p+r+15)/16)), alpha <= 0.99, stop when T(1-alpha) < 1e-4, 4th channel = sum alpha T / z.
#define ADB_CONV_GSPLAT 0
#define ADB_CONV_INRIA 1
#define ADB_MAX_ALPHA_INRIA 0.99f

struct AdbCam {
    const float* viewmat;  // device, 16
    const float* K;        // device, 9
    const float* campos;   // device, 3 (may be null when no SH)
    int W, H;
    float eps2d, near_plane, far_plane, radius_clip;
    int convention;  // ADB_CONV_*
};

// Tile rectangle [x0,x1) x [y0,y1) a splat touches.  The GSPLAT arithmetic must stay exactly as written: it feeds the
// bit-exact tile keys (oracle/raster_oracle.c tile_bounds).
__device__ __forceinline__ void adb_tile_rect(float u, float v, int rx_i, int ry_i, int W, int H, int convention,
                                              int& x0, int& x1, int& y0, int& y1) {
    const int tw = (W + ADB_TILE - 1) / ADB_TILE, th = (H + ADB_TILE - 1) / ADB_TILE;
    if (convention == ADB_CONV_INRIA) {
        const float px = u - 0.5f, py = v - 0.5f;  // Inria pixel-index coordinates (centres at integers)
        const float rx = (float)rx_i, ry = (float)ry_i, t = (float)ADB_TILE;
        x0 = min(tw, max(0, (int)((px - rx) / t)));
        x1 = min(tw, max(0, (int)((px + rx + (t - 1.f)) / t)));
        y0 = min(th, max(0, (int)((py - ry) / t)));
        y1 = min(th, max(0, (int)((py + ry + (t - 1.f)) / t)));
        return;
    }
    float mx = u / (float)ADB_TILE, my = v / (float)ADB_TILE;
    float trx = (float)rx_i / (float)ADB_TILE, try_ = (float)ry_i / (float)ADB_TILE;
    x0 = (int)fminf(fmaxf(0.f, floorf(mx - trx)), (float)tw);
    x1 = (int)fminf(fmaxf(0.f, ceilf(mx + trx)), (float)tw);
    y0 = (int)fminf(fmaxf(0.f, floorf(my - try_)), (float)th);
    y1 = (int)fminf(fmaxf(0.f, ceilf(my + try_)), (float)th);
}

__host__ __device__ inline int adb_tile_bits(int W, int H) {
    int tw = (W + ADB_TILE - 1) / ADB_TILE, th = (H + ADB_TILE - 1) / ADB_TILE;
    int n = tw * th, bits = 0;
    while (n > 0) { bits++; n >>= 1; }
    return bits;
}
__host__ __device__ inline int adb_cam_bits(int n_cams) {
    int bits = 0;
    while (n_cams > 0) { bits++; n_cams >>= 1; }
    return bits;
}
