/* TEST INFRASTRUCTURE ONLY.  CPU restatement (plain C + OpenMP) of the Gaussian-splat renderer on the
 * reference's hot path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this; the product path (artdeco_b200/) never does.
 *
 * What it restates: the call at /root/reference Reconstruct/scene/scene_models/h3dgsv3.py:664-680
 *   gsplat.rendering.rasterization(means, quats, scales, opacities, colors=SH[N,16,3], viewmats, Ks,
 *       width, height, render_mode="RGB+D", rasterize_mode="classic", absgrad=False, packed=False,
 *       sh_degree=3, eps2d=0.01)
 * and its autograd backward.  gsplat is an un-vendored, unpinned pip dependency (README.md:82; >=1.5
 * inferred from meta['radii'][C,N,2] at h3dgsv3.py:689) that cannot be installed offline, so this file
 * restates its published algorithm.
 *
 * PARITY UNPINNED: the reference has no test, golden vector or fixture for the renderer (SURVEY.md §4,
 * §8c).  What pins this file: (1) oracle/raster_torch.py + torch.autograd agree with every forward
 * output and every analytic gradient here (tests/test_oracle_raster.py); (2) finite differences.
 *
 * Floating point: compiled with -ffp-contract=off so the projection (which feeds the bit-exact
 * integer tile keys) evaluates exactly the expression trees written here; the CUDA projection kernel
 * is compiled with -fmad=false and written with the same trees.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "../include/adb_detmath.h"

#define ALPHA_THRESHOLD (1.0f / 255.0f)
#define MAX_ALPHA 0.999f
#define T_EPS 1e-4f
#define TILE 16

typedef struct {
    float fx, fy, cx, cy;
    int W, H;
    float eps2d, near_plane, far_plane, radius_clip;
} adbo_cam;

int adbo_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* bench.py --impl reference under torchrun inherits OMP_NUM_THREADS=1: the CPU arm sets its own thread count. */
void adbo_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

static void quat_to_rot(const float* q, float* R /*9*/, float* inv_norm) {
    float w = q[0], x = q[1], y = q[2], z = q[3];
    float n2 = w * w + x * x + y * y + z * z;
    float inv = 1.0f / sqrtf(n2);
    *inv_norm = inv;
    w = w * inv; x = x * inv; y = y * inv; z = z * inv;
    float x2 = x * x, y2 = y * y, z2 = z * z;
    float xy = x * y, xz = x * z, yz = y * z, wx = w * x, wy = w * y, wz = w * z;
    R[0] = 1.0f - 2.0f * (y2 + z2); R[1] = 2.0f * (xy - wz);        R[2] = 2.0f * (xz + wy);
    R[3] = 2.0f * (xy + wz);        R[4] = 1.0f - 2.0f * (x2 + z2); R[5] = 2.0f * (yz - wx);
    R[6] = 2.0f * (xz - wy);        R[7] = 2.0f * (yz + wx);        R[8] = 1.0f - 2.0f * (x2 + y2);
}

/* 3x3 row-major helpers */
static void mat3_mul(const float* A, const float* B, float* C) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            C[i * 3 + j] = A[i * 3 + 0] * B[0 * 3 + j] + A[i * 3 + 1] * B[1 * 3 + j] + A[i * 3 + 2] * B[2 * 3 + j];
}
static void mat3_mul_bt(const float* A, const float* B, float* C) { /* A * B^T */
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            C[i * 3 + j] = A[i * 3 + 0] * B[j * 3 + 0] + A[i * 3 + 1] * B[j * 3 + 1] + A[i * 3 + 2] * B[j * 3 + 2];
}
static void mat3_mul_at(const float* A, const float* B, float* C) { /* A^T * B */
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            C[i * 3 + j] = A[0 * 3 + i] * B[0 * 3 + j] + A[1 * 3 + i] * B[1 * 3 + j] + A[2 * 3 + i] * B[2 * 3 + j];
}

typedef struct {
    float p[3];      /* camera-space mean */
    float Rq[9];     /* rotation from the normalised quaternion */
    float qinv;      /* 1/|q| */
    float Sigma[9];  /* world covariance */
    float Sc[9];     /* camera covariance */
    float J[6];      /* 2x3 */
    float tx, ty;
    int clamp_x, clamp_y; /* 1 when x/z (y/z) was clamped */
    float a, b, c, det;   /* blurred 2D covariance */
} proj_state;

/* Forward projection of one Gaussian. Returns 1 if it survives culling. */
static int project_one(const float* mean, const float* quat, const float* scale, float opacity,
                       const float* V /*16 row-major world->cam*/, const adbo_cam* cam, proj_state* st,
                       int* radii, float* mean2d, float* depth, float* conic) {
    const float R[9] = {V[0], V[1], V[2], V[4], V[5], V[6], V[8], V[9], V[10]};
    const float t[3] = {V[3], V[7], V[11]};
    float x = R[0] * mean[0] + R[1] * mean[1] + R[2] * mean[2] + t[0];
    float y = R[3] * mean[0] + R[4] * mean[1] + R[5] * mean[2] + t[1];
    float z = R[6] * mean[0] + R[7] * mean[1] + R[8] * mean[2] + t[2];
    st->p[0] = x; st->p[1] = y; st->p[2] = z;
    radii[0] = radii[1] = 0;
    mean2d[0] = mean2d[1] = 0.f; *depth = 0.f; conic[0] = conic[1] = conic[2] = 0.f;
    if (z < cam->near_plane || z > cam->far_plane) return 0;

    quat_to_rot(quat, st->Rq, &st->qinv);
    float M[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) M[i * 3 + j] = st->Rq[i * 3 + j] * scale[j];
    mat3_mul_bt(M, M, st->Sigma);
    float RS[9];
    mat3_mul(R, st->Sigma, RS);
    mat3_mul_bt(RS, R, st->Sc);

    float fx = cam->fx, fy = cam->fy, cx = cam->cx, cy = cam->cy;
    float W = (float)cam->W, H = (float)cam->H;
    float tanx = 0.5f * W / fx, tany = 0.5f * H / fy;
    float lxp = (W - cx) / fx + 0.3f * tanx, lxn = cx / fx + 0.3f * tanx;
    float lyp = (H - cy) / fy + 0.3f * tany, lyn = cy / fy + 0.3f * tany;
    float rz = 1.0f / z;
    float rz2 = rz * rz;
    float xr = x * rz, yr = y * rz;
    float cxr = fminf(lxp, fmaxf(-lxn, xr));
    float cyr = fminf(lyp, fmaxf(-lyn, yr));
    st->clamp_x = !(xr <= lxp && xr >= -lxn);
    st->clamp_y = !(yr <= lyp && yr >= -lyn);
    float tx = z * cxr, ty = z * cyr;
    st->tx = tx; st->ty = ty;
    float* J = st->J;
    J[0] = fx * rz; J[1] = 0.f; J[2] = -fx * tx * rz2;
    J[3] = 0.f; J[4] = fy * rz; J[5] = -fy * ty * rz2;
    /* Sigma2 = J Sc J^T, J has zeros at [1] and [3] */
    const float* S = st->Sc;
    float k00 = J[0] * S[0] + J[2] * S[6], k01 = J[0] * S[1] + J[2] * S[7], k02 = J[0] * S[2] + J[2] * S[8];
    float k10 = J[4] * S[3] + J[5] * S[6], k11 = J[4] * S[4] + J[5] * S[7], k12 = J[4] * S[5] + J[5] * S[8];
    float a = k00 * J[0] + k02 * J[2];
    float b = k01 * J[4] + k02 * J[5];
    float c = k11 * J[4] + k12 * J[5];
    (void)k10;
    a = a + cam->eps2d;
    c = c + cam->eps2d;
    float det = a * c - b * b;
    st->a = a; st->b = b; st->c = c; st->det = det;
    if (!(det > 0.f)) return 0;
    float u = fx * x * rz + cx, v = fy * y * rz + cy;

    if (opacity < ALPHA_THRESHOLD) return 0;
    float ext = sqrtf(2.0f * adb_det_logf(opacity / ALPHA_THRESHOLD));
    ext = fminf(3.33f, ext);
    float bb = 0.5f * (a + c);
    float lam = bb + sqrtf(fmaxf(0.01f, bb * bb - det));
    float r1 = ext * sqrtf(lam);
    float rx = ceilf(fminf(ext * sqrtf(a), r1));
    float ry = ceilf(fminf(ext * sqrtf(c), r1));
    if (rx <= cam->radius_clip && ry <= cam->radius_clip) return 0;
    if (u + rx <= 0.f || u - rx >= W || v + ry <= 0.f || v - ry >= H) return 0;
    radii[0] = (int)rx; radii[1] = (int)ry;
    mean2d[0] = u; mean2d[1] = v;
    *depth = z;
    conic[0] = c / det; conic[1] = -b / det; conic[2] = a / det;
    return 1;
}

void adbo_project(int N, const float* means, const float* quats, const float* scales, const float* opac,
                  const float* viewmat, const adbo_cam* cam, int32_t* radii, float* means2d, float* depths,
                  float* conics) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < N; ++i) {
        proj_state st;
        int r[2];
        project_one(means + 3 * i, quats + 4 * i, scales + 3 * i, opac[i], viewmat, cam, &st, r, means2d + 2 * i,
                    depths + i, conics + 3 * i);
        radii[2 * i] = r[0]; radii[2 * i + 1] = r[1];
    }
}

/* ---- spherical harmonics (C0 at Reconstruct/utils.py:119) ---- */
static void sh_basis(int deg, float x, float y, float z, float* B /*16*/) {
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

/* d(basis_k)/d(x,y,z) for the unit direction */
static void sh_basis_grad(int deg, float x, float y, float z, float* Bx, float* By, float* Bz) {
    for (int k = 0; k < 16; ++k) Bx[k] = By[k] = Bz[k] = 0.f;
    if (deg < 1) return;
    By[1] = -0.48860251190292f; Bz[2] = 0.48860251190292f; Bx[3] = -0.48860251190292f;
    if (deg < 2) return;
    float z2 = z * z, fT0B = -1.092548430592079f * z, fT0B_z = -1.092548430592079f;
    float fC1 = x * x - y * y, fS1 = 2.f * x * y;
    float fC1_x = 2.f * x, fC1_y = -2.f * y, fS1_x = 2.f * y, fS1_y = 2.f * x;
    Bx[4] = 0.5462742152960395f * fS1_x; By[4] = 0.5462742152960395f * fS1_y;
    By[5] = fT0B; Bz[5] = fT0B_z * y;
    Bz[6] = 2.f * 0.9461746957575601f * z;
    Bx[7] = fT0B; Bz[7] = fT0B_z * x;
    Bx[8] = 0.5462742152960395f * fC1_x; By[8] = 0.5462742152960395f * fC1_y;
    if (deg < 3) return;
    float fT0C = -2.285228997322329f * z2 + 0.4570457994644658f, fT0C_z = -2.285228997322329f * 2.f * z;
    float fT1B = 1.445305721320277f * z, fT1B_z = 1.445305721320277f;
    float fC2_x = fC1 + x * fC1_x - y * fS1_x, fC2_y = x * fC1_y - fS1 - y * fS1_y;
    float fS2_x = fS1 + x * fS1_x + y * fC1_x, fS2_y = x * fS1_y + fC1 + y * fC1_y;
    Bx[9] = -0.5900435899266435f * fS2_x; By[9] = -0.5900435899266435f * fS2_y;
    Bx[10] = fT1B * fS1_x; By[10] = fT1B * fS1_y; Bz[10] = fT1B_z * fS1;
    By[11] = fT0C; Bz[11] = fT0C_z * y;
    Bz[12] = 3.f * 1.865881662950577f * z2 - 1.119528997770346f;
    Bx[13] = fT0C; Bz[13] = fT0C_z * x;
    Bx[14] = fT1B * fC1_x; By[14] = fT1B * fC1_y; Bz[14] = fT1B_z * fC1;
    Bx[15] = -0.5900435899266435f * fC2_x; By[15] = -0.5900435899266435f * fC2_y;
}

/* rgb = max(sum_k B_k(dir) sh_k + 0.5, 0); only for Gaussians with radii > 0 (others: 0) */
void adbo_sh_fwd(int N, int deg, const float* means, const float* campos, const float* sh /*[N,16,3]*/,
                 const int32_t* radii, float* rgb) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < N; ++i) {
        rgb[3 * i] = rgb[3 * i + 1] = rgb[3 * i + 2] = 0.f;
        if (radii[2 * i] <= 0 && radii[2 * i + 1] <= 0) continue;
        float dx = means[3 * i] - campos[0], dy = means[3 * i + 1] - campos[1], dz = means[3 * i + 2] - campos[2];
        float inv = 1.0f / sqrtf(dx * dx + dy * dy + dz * dz);
        float B[16];
        sh_basis(deg, dx * inv, dy * inv, dz * inv, B);
        for (int c = 0; c < 3; ++c) {
            float acc = 0.f;
            for (int k = 0; k < 16; ++k) acc += B[k] * sh[(size_t)i * 48 + k * 3 + c];
            rgb[3 * i + c] = fmaxf(acc + 0.5f, 0.f);
        }
    }
}

/* v_rgb -> v_sh[N,16,3], v_means[N,3] (+=), v_campos[3] (+=, double accumulation) */
void adbo_sh_bwd(int N, int deg, const float* means, const float* campos, const float* sh, const int32_t* radii,
                 const float* rgb, const float* v_rgb, float* v_sh, float* v_means, float* v_campos) {
    double vc[3] = {0, 0, 0};
#pragma omp parallel for schedule(static) reduction(+ : vc[:3])
    for (int i = 0; i < N; ++i) {
        for (int k = 0; k < 48; ++k) v_sh[(size_t)i * 48 + k] = 0.f;
        if (radii[2 * i] <= 0 && radii[2 * i + 1] <= 0) continue;
        float dx = means[3 * i] - campos[0], dy = means[3 * i + 1] - campos[1], dz = means[3 * i + 2] - campos[2];
        float inv = 1.0f / sqrtf(dx * dx + dy * dy + dz * dz);
        float nx = dx * inv, ny = dy * inv, nz = dz * inv;
        float B[16], Bx[16], By[16], Bz[16];
        sh_basis(deg, nx, ny, nz, B);
        sh_basis_grad(deg, nx, ny, nz, Bx, By, Bz);
        float vnx = 0.f, vny = 0.f, vnz = 0.f;
        for (int c = 0; c < 3; ++c) {
            float g = rgb[3 * i + c] > 0.f ? v_rgb[3 * i + c] : 0.f; /* clamp_min(.,0) has zero grad when clamped */
            for (int k = 0; k < 16; ++k) {
                float s = sh[(size_t)i * 48 + k * 3 + c];
                v_sh[(size_t)i * 48 + k * 3 + c] = B[k] * g;
                vnx += Bx[k] * s * g; vny += By[k] * s * g; vnz += Bz[k] * s * g;
            }
        }
        float dot = vnx * nx + vny * ny + vnz * nz;
        float gx = (vnx - dot * nx) * inv, gy = (vny - dot * ny) * inv, gz = (vnz - dot * nz) * inv;
        v_means[3 * i] += gx; v_means[3 * i + 1] += gy; v_means[3 * i + 2] += gz;
        vc[0] -= gx; vc[1] -= gy; vc[2] -= gz;
    }
    for (int k = 0; k < 3; ++k) v_campos[k] += (float)vc[k];
}

/* ---- tile intersection, keys, stable sort, tile ranges (bit-exact contract) ---- */
static void tile_bounds(const float* mean2d, const int32_t* radii, int tw, int th, int* x0, int* x1, int* y0, int* y1) {
    float mx = mean2d[0] / (float)TILE, my = mean2d[1] / (float)TILE;
    float rx = (float)radii[0] / (float)TILE, ry = (float)radii[1] / (float)TILE;
    float fx0 = floorf(mx - rx), fx1 = ceilf(mx + rx), fy0 = floorf(my - ry), fy1 = ceilf(my + ry);
    *x0 = (int)fminf(fmaxf(0.f, fx0), (float)tw);
    *x1 = (int)fminf(fmaxf(0.f, fx1), (float)tw);
    *y0 = (int)fminf(fmaxf(0.f, fy0), (float)th);
    *y1 = (int)fminf(fmaxf(0.f, fy1), (float)th);
}

int adbo_tile_bits(int W, int H) {
    int tw = (W + TILE - 1) / TILE, th = (H + TILE - 1) / TILE;
    int n = tw * th, bits = 0;
    while (n > 0) { bits++; n >>= 1; }
    return bits; /* == floor(log2(n_tiles)) + 1 == Python int.bit_length() */
}

/* tiles_per_gauss[N]; returns total */
int64_t adbo_isect_count(int N, const int32_t* radii, const float* means2d, int W, int H, int32_t* tiles_per_gauss) {
    int tw = (W + TILE - 1) / TILE, th = (H + TILE - 1) / TILE;
    int64_t total = 0;
#pragma omp parallel for schedule(static) reduction(+ : total)
    for (int i = 0; i < N; ++i) {
        int cnt = 0;
        if (radii[2 * i] > 0 || radii[2 * i + 1] > 0) {
            int x0, x1, y0, y1;
            tile_bounds(means2d + 2 * i, radii + 2 * i, tw, th, &x0, &x1, &y0, &y1);
            cnt = (x1 - x0) * (y1 - y0);
        }
        tiles_per_gauss[i] = cnt;
        total += cnt;
    }
    return total;
}

static void radix_sort_pairs(int64_t n, uint64_t* keys, int32_t* vals, int bits) {
    uint64_t* k2 = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(n > 0 ? n : 1));
    int32_t* v2 = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    size_t* hist = (size_t*)malloc(sizeof(size_t) * 65537);
    uint64_t *ka = keys, *kb = k2;
    int32_t *va = vals, *vb = v2;
    for (int shift = 0; shift < bits; shift += 16) {
        memset(hist, 0, sizeof(size_t) * 65537);
        for (int64_t i = 0; i < n; ++i) hist[((ka[i] >> shift) & 0xFFFF) + 1]++;
        for (int d = 0; d < 65536; ++d) hist[d + 1] += hist[d];
        for (int64_t i = 0; i < n; ++i) {
            size_t pos = hist[(ka[i] >> shift) & 0xFFFF]++;
            kb[pos] = ka[i]; vb[pos] = va[i];
        }
        uint64_t* tk = ka; ka = kb; kb = tk;
        int32_t* tv = va; va = vb; vb = tv;
    }
    if (ka != keys) { memcpy(keys, ka, sizeof(uint64_t) * (size_t)n); memcpy(vals, va, sizeof(int32_t) * (size_t)n); }
    free(k2); free(v2); free(hist);
}

/* Emits (key,val) in Gaussian order (ty outer, tx inner), stable-sorts ascending, and writes the first
 * sorted index of every tile into tile_offsets[T] (T = tw*th).  keys/vals have room for `total`. */
void adbo_isect_sort(int N, const int32_t* radii, const float* means2d, const float* depths, int W, int H,
                     int cam_id, int n_cams, const int32_t* tiles_per_gauss, int64_t total, int64_t* keys,
                     int32_t* vals, int32_t* tile_offsets, int sort) {
    int tw = (W + TILE - 1) / TILE, th = (H + TILE - 1) / TILE;
    int tile_bits = adbo_tile_bits(W, H);
    int cam_bits = 0;
    { int c = n_cams; while (c > 0) { cam_bits++; c >>= 1; } } /* gsplat: floor(log2(C)) + 1 */
    int64_t* offs = (int64_t*)malloc(sizeof(int64_t) * (size_t)(N + 1));
    offs[0] = 0;
    for (int i = 0; i < N; ++i) offs[i + 1] = offs[i] + tiles_per_gauss[i];
#pragma omp parallel for schedule(dynamic, 1024)
    for (int i = 0; i < N; ++i) {
        if (tiles_per_gauss[i] == 0) continue;
        int x0, x1, y0, y1;
        tile_bounds(means2d + 2 * i, radii + 2 * i, tw, th, &x0, &x1, &y0, &y1);
        uint32_t dbits;
        memcpy(&dbits, depths + i, 4);
        int64_t o = offs[i];
        for (int ty = y0; ty < y1; ++ty)
            for (int tx = x0; tx < x1; ++tx) {
                uint64_t tile = (uint64_t)(ty * tw + tx);
                keys[o] = (int64_t)(((uint64_t)cam_id << (32 + tile_bits)) | (tile << 32) | (uint64_t)dbits);
                vals[o] = cam_id * N + i;
                ++o;
            }
    }
    free(offs);
    if (sort) radix_sort_pairs(total, (uint64_t*)keys, vals, 32 + tile_bits + cam_bits);
    if (tile_offsets) {
        int T = tw * th;
        uint64_t mask = ((uint64_t)1 << tile_bits) - 1;
        int64_t k = 0;
        for (int tl = 0; tl < T; ++tl) {
            while (k < total && ((((uint64_t)keys[k]) >> 32) & mask) < (uint64_t)tl) ++k;
            tile_offsets[tl] = (int32_t)k;
        }
    }
}

/* ---- alpha blending forward.  feats [N,CH] with CH<=8. ---- */
void adbo_blend_fwd(int W, int H, int CH, const float* means2d, const float* conics, const float* opac,
                    const float* feats, const int32_t* vals, int64_t n_isect, const int32_t* tile_offsets,
                    float* out /*[H,W,CH]*/, float* alphas /*[H,W]*/, int32_t* last_ids /*[H,W]*/) {
    int tw = (W + TILE - 1) / TILE, th = (H + TILE - 1) / TILE;
#pragma omp parallel for schedule(dynamic, 1)
    for (int tl = 0; tl < tw * th; ++tl) {
        int tyi = tl / tw, txi = tl % tw;
        int64_t start = tile_offsets[tl];
        int64_t end = (tl + 1 < tw * th) ? tile_offsets[tl + 1] : n_isect;
        for (int i = tyi * TILE; i < tyi * TILE + TILE && i < H; ++i)
            for (int j = txi * TILE; j < txi * TILE + TILE && j < W; ++j) {
                float px = (float)j + 0.5f, py = (float)i + 0.5f;
                float T = 1.0f;
                float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                int32_t cur = 0;
                for (int64_t k = start; k < end; ++k) {
                    int g = vals[k];
                    float dx = means2d[2 * g] - px, dy = means2d[2 * g + 1] - py;
                    float a = conics[3 * g], b = conics[3 * g + 1], c = conics[3 * g + 2];
                    float sigma = 0.5f * (a * dx * dx + c * dy * dy) + b * dx * dy;
                    float alpha = fminf(MAX_ALPHA, opac[g] * expf(-sigma));
                    if (sigma < 0.f || alpha < ALPHA_THRESHOLD) continue;
                    float nT = T * (1.0f - alpha);
                    if (nT <= T_EPS) break;
                    float vis = alpha * T;
                    for (int ch = 0; ch < CH; ++ch) acc[ch] += feats[(size_t)g * CH + ch] * vis;
                    cur = (int32_t)k;
                    T = nT;
                }
                size_t pix = (size_t)i * W + j;
                for (int ch = 0; ch < CH; ++ch) out[pix * CH + ch] = acc[ch];
                alphas[pix] = 1.0f - T;
                last_ids[pix] = cur;
            }
    }
}

/* ---- alpha blending backward.  Outputs are ACCUMULATED in double then stored. ---- */
void adbo_blend_bwd(int W, int H, int CH, int N, const float* means2d, const float* conics, const float* opac,
                    const float* feats, const int32_t* vals, int64_t n_isect, const int32_t* tile_offsets,
                    const float* alphas, const int32_t* last_ids, const float* v_out /*[H,W,CH]*/,
                    const float* v_alphas /*[H,W]*/, float* v_means2d /*[N,2]*/, float* v_conics /*[N,3]*/,
                    float* v_opac /*[N]*/, float* v_feats /*[N,CH]*/) {
    int tw = (W + TILE - 1) / TILE, th = (H + TILE - 1) / TILE;
    int stride = 6 + CH;
    double* acc = (double*)calloc((size_t)N * stride, sizeof(double));
#pragma omp parallel for schedule(dynamic, 1)
    for (int tl = 0; tl < tw * th; ++tl) {
        int tyi = tl / tw, txi = tl % tw;
        int64_t start = tile_offsets[tl];
        int64_t end = (tl + 1 < tw * th) ? tile_offsets[tl + 1] : n_isect;
        if (end <= start) continue;
        for (int i = tyi * TILE; i < tyi * TILE + TILE && i < H; ++i)
            for (int j = txi * TILE; j < txi * TILE + TILE && j < W; ++j) {
                size_t pix = (size_t)i * W + j;
                float px = (float)j + 0.5f, py = (float)i + 0.5f;
                float T_final = 1.0f - alphas[pix];
                float T = T_final;
                float buffer[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                const float* vo = v_out + pix * CH;
                float va = v_alphas[pix];
                for (int64_t k = last_ids[pix]; k >= start; --k) {
                    int g = vals[k];
                    float dx = means2d[2 * g] - px, dy = means2d[2 * g + 1] - py;
                    float a = conics[3 * g], b = conics[3 * g + 1], c = conics[3 * g + 2];
                    float sigma = 0.5f * (a * dx * dx + c * dy * dy) + b * dx * dy;
                    float vis = expf(-sigma);
                    float alpha = fminf(MAX_ALPHA, opac[g] * vis);
                    if (sigma < 0.f || alpha < ALPHA_THRESHOLD) continue;
                    float ra = 1.0f / (1.0f - alpha);
                    T *= ra;
                    float fac = alpha * T;
                    float v_alpha = 0.f;
                    double* ag = acc + (size_t)g * stride;
                    for (int ch = 0; ch < CH; ++ch) {
                        float f = feats[(size_t)g * CH + ch];
                        v_alpha += (f * T - buffer[ch] * ra) * vo[ch];
#pragma omp atomic
                        ag[6 + ch] += (double)(fac * vo[ch]);
                        buffer[ch] += f * fac;
                    }
                    v_alpha += T_final * ra * va;
                    if (opac[g] * vis <= MAX_ALPHA) {
                        float v_sigma = -opac[g] * vis * v_alpha;
                        float gx = v_sigma * (a * dx + b * dy), gy = v_sigma * (b * dx + c * dy);
                        float ca = 0.5f * v_sigma * dx * dx, cb = v_sigma * dx * dy, cc = 0.5f * v_sigma * dy * dy;
                        float go = vis * v_alpha;
#pragma omp atomic
                        ag[0] += (double)gx;
#pragma omp atomic
                        ag[1] += (double)gy;
#pragma omp atomic
                        ag[2] += (double)ca;
#pragma omp atomic
                        ag[3] += (double)cb;
#pragma omp atomic
                        ag[4] += (double)cc;
#pragma omp atomic
                        ag[5] += (double)go;
                    }
                }
            }
    }
#pragma omp parallel for schedule(static)
    for (int g = 0; g < N; ++g) {
        const double* ag = acc + (size_t)g * stride;
        v_means2d[2 * g] = (float)ag[0]; v_means2d[2 * g + 1] = (float)ag[1];
        v_conics[3 * g] = (float)ag[2]; v_conics[3 * g + 1] = (float)ag[3]; v_conics[3 * g + 2] = (float)ag[4];
        v_opac[g] = (float)ag[5];
        for (int ch = 0; ch < CH; ++ch) v_feats[(size_t)g * CH + ch] = (float)ag[6 + ch];
    }
    free(acc);
}

/* ---- projection backward.  v_viewmat[16] accumulated in double. ---- */
void adbo_project_bwd(int N, const float* means, const float* quats, const float* scales, const float* opac,
                      const float* viewmat, const adbo_cam* cam, const int32_t* radii, const float* v_means2d,
                      const float* v_depths, const float* v_conics, float* v_means /*[N,3] =*/,
                      float* v_quats /*[N,4] =*/, float* v_scales /*[N,3] =*/, float* v_viewmat /*[16] +=*/) {
    const float* V = viewmat;
    const float R[9] = {V[0], V[1], V[2], V[4], V[5], V[6], V[8], V[9], V[10]};
    double vR[9] = {0}, vt[3] = {0};
#pragma omp parallel for schedule(static) reduction(+ : vR[:9], vt[:3])
    for (int i = 0; i < N; ++i) {
        for (int k = 0; k < 3; ++k) v_means[3 * i + k] = 0.f, v_scales[3 * i + k] = 0.f;
        for (int k = 0; k < 4; ++k) v_quats[4 * i + k] = 0.f;
        if (radii[2 * i] <= 0 && radii[2 * i + 1] <= 0) continue;
        proj_state st;
        int r[2];
        float m2[2], dep, con[3];
        if (!project_one(means + 3 * i, quats + 4 * i, scales + 3 * i, opac[i], viewmat, cam, &st, r, m2, &dep, con))
            continue;
        float fx = cam->fx, fy = cam->fy;
        float x = st.p[0], y = st.p[1], z = st.p[2];
        float rz = 1.0f / z, rz2 = rz * rz, rz3 = rz2 * rz;
        /* conic -> blurred covariance: V2 = -Q G Q, G = [[vA, vB/2],[vB/2, vC]] */
        float A = con[0], B = con[1], C = con[2];
        float gA = v_conics[3 * i], gB = 0.5f * v_conics[3 * i + 1], gC = v_conics[3 * i + 2];
        /* QG */
        float q00 = A * gA + B * gB, q01 = A * gB + B * gC, q10 = B * gA + C * gB, q11 = B * gB + C * gC;
        float V00 = -(q00 * A + q01 * B), V01 = -(q00 * B + q01 * C), V10 = -(q10 * A + q11 * B), V11 = -(q10 * B + q11 * C);
        /* v_Sc = J^T V2 J  (3x3) ; v_J = 2 * V2 J Sc  (V2, Sc symmetric) */
        const float* J = st.J;
        float VJ[6] = {V00 * J[0] + V01 * J[3], V00 * J[1] + V01 * J[4], V00 * J[2] + V01 * J[5],
                       V10 * J[0] + V11 * J[3], V10 * J[1] + V11 * J[4], V10 * J[2] + V11 * J[5]};
        float vSc[9];
        for (int a_ = 0; a_ < 3; ++a_)
            for (int b_ = 0; b_ < 3; ++b_) vSc[a_ * 3 + b_] = J[a_] * VJ[b_] + J[3 + a_] * VJ[3 + b_];
        float vJ[6];
        for (int a_ = 0; a_ < 2; ++a_)
            for (int b_ = 0; b_ < 3; ++b_)
                vJ[a_ * 3 + b_] = 2.0f * (VJ[a_ * 3 + 0] * st.Sc[0 * 3 + b_] + VJ[a_ * 3 + 1] * st.Sc[1 * 3 + b_] +
                                          VJ[a_ * 3 + 2] * st.Sc[2 * 3 + b_]);
        float vu = v_means2d[2 * i], vv = v_means2d[2 * i + 1];
        float vp[3];
        vp[0] = fx * rz * vu;
        vp[1] = fy * rz * vv;
        vp[2] = -(fx * x * vu + fy * y * vv) * rz2 + v_depths[i];
        vp[2] += -fx * rz2 * vJ[0] - fy * rz2 * vJ[4];
        if (!st.clamp_x) { vp[0] += -fx * rz2 * vJ[2]; vp[2] += 2.0f * fx * st.tx * rz3 * vJ[2]; }
        else             { vp[2] += fx * st.tx * rz3 * vJ[2]; }
        if (!st.clamp_y) { vp[1] += -fy * rz2 * vJ[5]; vp[2] += 2.0f * fy * st.ty * rz3 * vJ[5]; }
        else             { vp[2] += fy * st.ty * rz3 * vJ[5]; }
        /* p = R mu + t */
        const float* mu = means + 3 * i;
        for (int a_ = 0; a_ < 3; ++a_) {
            v_means[3 * i + a_] = R[0 * 3 + a_] * vp[0] + R[1 * 3 + a_] * vp[1] + R[2 * 3 + a_] * vp[2];
            vt[a_] += vp[a_];
            for (int b_ = 0; b_ < 3; ++b_) vR[a_ * 3 + b_] += (double)(vp[a_] * mu[b_]);
        }
        /* Sc = R Sigma R^T : v_Sigma = R^T vSc R ; v_R += 2 * sym(vSc) R Sigma  (vSc symmetric here) */
        float tmp[9], vSigma[9], RSig[9], add[9];
        mat3_mul_at(R, vSc, tmp);
        mat3_mul(tmp, R, vSigma);
        mat3_mul(R, st.Sigma, RSig);
        mat3_mul(vSc, RSig, add);
        for (int k = 0; k < 9; ++k) vR[k] += (double)(2.0f * add[k]);
        /* Sigma = M M^T : v_M = 2 vSigma M ; M = Rq diag(s) */
        float M[9], vM[9];
        const float* s = scales + 3 * i;
        for (int a_ = 0; a_ < 3; ++a_)
            for (int b_ = 0; b_ < 3; ++b_) M[a_ * 3 + b_] = st.Rq[a_ * 3 + b_] * s[b_];
        mat3_mul(vSigma, M, vM);
        float vRq[9];
        for (int a_ = 0; a_ < 3; ++a_)
            for (int b_ = 0; b_ < 3; ++b_) {
                vM[a_ * 3 + b_] *= 2.0f;
                vRq[a_ * 3 + b_] = vM[a_ * 3 + b_] * s[b_];
            }
        for (int b_ = 0; b_ < 3; ++b_)
            v_scales[3 * i + b_] = vM[0 * 3 + b_] * st.Rq[0 * 3 + b_] + vM[1 * 3 + b_] * st.Rq[1 * 3 + b_] +
                                   vM[2 * 3 + b_] * st.Rq[2 * 3 + b_];
        /* rotation matrix -> normalised quaternion -> raw quaternion */
        const float* q = quats + 4 * i;
        float inv = st.qinv;
        float w = q[0] * inv, qx = q[1] * inv, qy = q[2] * inv, qz = q[3] * inv;
        float vn[4];
        vn[0] = 2.0f * (qx * (vRq[7] - vRq[5]) + qy * (vRq[2] - vRq[6]) + qz * (vRq[3] - vRq[1]));
        vn[1] = 2.0f * (-2.0f * qx * (vRq[4] + vRq[8]) + qy * (vRq[1] + vRq[3]) + qz * (vRq[2] + vRq[6]) + w * (vRq[7] - vRq[5]));
        vn[2] = 2.0f * (qx * (vRq[1] + vRq[3]) - 2.0f * qy * (vRq[0] + vRq[8]) + qz * (vRq[5] + vRq[7]) + w * (vRq[2] - vRq[6]));
        vn[3] = 2.0f * (qx * (vRq[2] + vRq[6]) + qy * (vRq[5] + vRq[7]) - 2.0f * qz * (vRq[0] + vRq[4]) + w * (vRq[3] - vRq[1]));
        float dot = vn[0] * w + vn[1] * qx + vn[2] * qy + vn[3] * qz;
        v_quats[4 * i + 0] = (vn[0] - dot * w) * inv;
        v_quats[4 * i + 1] = (vn[1] - dot * qx) * inv;
        v_quats[4 * i + 2] = (vn[2] - dot * qy) * inv;
        v_quats[4 * i + 3] = (vn[3] - dot * qz) * inv;
    }
    for (int a_ = 0; a_ < 3; ++a_) {
        for (int b_ = 0; b_ < 3; ++b_) v_viewmat[a_ * 4 + b_] += (float)vR[a_ * 3 + b_];
        v_viewmat[a_ * 4 + 3] += (float)vt[a_];
    }
}

/* ---- sparse Adam restatement (call sites Reconstruct/scene/optimizers.py:48-57,
 * 90-99,116-128,144-156).  Source of adamUpdate is un-vendored (on-the-fly-nvs); PARITY UNPINNED. ---- */
void adbo_adam(int64_t N, int64_t M, float* param, const float* grad, float* m1, float* m2,
               const uint8_t* visible /*nullable*/, const float* lr, int64_t lr_numel, float b1, float b2, float eps) {
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < N; ++r) {
        if (visible && !visible[r]) continue;
        for (int64_t c = 0; c < M; ++c) {
            int64_t i = r * M + c;
            float l = lr_numel == 1 ? lr[0] : (lr_numel == N ? lr[r] : lr[i]);
            float g = grad[i];
            float a = b1 * m1[i] + (1.0f - b1) * g;
            float v = b2 * m2[i] + (1.0f - b2) * g * g;
            m1[i] = a; m2[i] = v;
            param[i] -= l * a / (sqrtf(v) + eps);
        }
    }
}
