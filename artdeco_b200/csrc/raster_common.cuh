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

struct AdbCam {
    const float* viewmat;  // device, 16
    const float* K;        // device, 9
    const float* campos;   // device, 3 (may be null when no SH)
    int W, H;
    float eps2d, near_plane, far_plane, radius_clip;
};

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
