// Per-Gaussian covariance modulation MLP, forward and backward, fused.
//
// Replaces the torch block of SceneModel.render at Reconstruct/scene/scene_models/h3dgsv3.py:656-662:
//   x = cat(global_feat[cls_id], local_feat)            [N, D]   D = Fg + Fl  (32 with run.sh, 64 by default)
//   o = Linear(D,7)(ReLU(Linear(D,D)(x)))               (mlp_cov, h3dgsv3.py:173-177)
//   scaling_out  = scaling * sigmoid(o[:, :3])
//   rotation_out = F.normalize(rotation * o[:, 3:])     (eps 1e-12)
// which the reference runs as a gather, a cat, two skinny cuBLAS GEMMs and ~8 elementwise kernels over N rows.
// Here one thread owns one Gaussian: weights live in shared memory, the D-vector in registers; the backward
// recomputes the hidden layer instead of storing it, reduces the weight gradients per CTA as a [D x 256] x [256 x D]
// product in shared memory, and scatters v_global_feat with atomics.
#include "common.cuh"

namespace {

constexpr int DMAX = 64;
constexpr int TPB = 128;

struct CovMlp {
    int D, Fg, Fl;
    const float *W1, *b1, *W2, *b2;   // W1 [D,D] row-major (out,in); W2 [7,D]
};

__device__ __forceinline__ void load_weights(const CovMlp& m, float* sW1, float* sb1, float* sW2, float* sb2) {
    for (int i = threadIdx.x; i < m.D * m.D; i += blockDim.x) sW1[i] = m.W1[i];
    for (int i = threadIdx.x; i < 7 * m.D; i += blockDim.x) sW2[i] = m.W2[i];
    if (threadIdx.x < m.D) sb1[threadIdx.x] = m.b1[threadIdx.x];
    if (threadIdx.x < 7) sb2[threadIdx.x] = m.b2[threadIdx.x];
}

template <int D>
__device__ __forceinline__ void gather_x(const CovMlp& m, const float* __restrict__ gfeat, const float* __restrict__ lfeat,
                                         long long cls, long long i, float* x) {
#pragma unroll
    for (int k = 0; k < D; ++k) x[k] = k < m.Fg ? __ldg(gfeat + cls * m.Fg + k) : __ldg(lfeat + i * m.Fl + (k - m.Fg));
}

template <int D>
__global__ void __launch_bounds__(TPB)
cov_mlp_fwd_kernel(long long N, CovMlp m, const float* __restrict__ gfeat, const float* __restrict__ lfeat,
                   const long long* __restrict__ cls_id, const float* __restrict__ scaling,
                   const float* __restrict__ rotation, float* __restrict__ scale_out, float* __restrict__ rot_out) {
    __shared__ float sW1[D * D], sb1[D], sW2[7 * D], sb2[8];
    load_weights(m, sW1, sb1, sW2, sb2);
    __syncthreads();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float x[D];
    gather_x<D>(m, gfeat, lfeat, cls_id[i], i, x);
    float o[7];
#pragma unroll
    for (int c = 0; c < 7; ++c) o[c] = sb2[c];
    for (int j = 0; j < D; ++j) {
        float a = sb1[j];
#pragma unroll
        for (int k = 0; k < D; ++k) a += sW1[j * D + k] * x[k];
        const float h = fmaxf(a, 0.f);
#pragma unroll
        for (int c = 0; c < 7; ++c) o[c] += sW2[c * D + j] * h;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) scale_out[3 * i + c] = scaling[3 * i + c] * (1.0f / (1.0f + expf(-o[c])));
    float r[4], n2 = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) { r[c] = rotation[4 * i + c] * o[3 + c]; n2 += r[c] * r[c]; }
    const float inv = 1.0f / fmaxf(sqrtf(n2), 1e-12f);
    reinterpret_cast<float4*>(rot_out)[i] = make_float4(r[0] * inv, r[1] * inv, r[2] * inv, r[3] * inv);
}

template <int D>
__global__ void __launch_bounds__(TPB)
cov_mlp_bwd_kernel(long long N, CovMlp m, const float* __restrict__ gfeat, const float* __restrict__ lfeat,
                   const long long* __restrict__ cls_id, const float* __restrict__ scaling,
                   const float* __restrict__ rotation, const float* __restrict__ v_scale_out,
                   const float* __restrict__ v_rot_out, float* __restrict__ v_scaling, float* __restrict__ v_rotation,
                   float* __restrict__ v_lfeat, float* __restrict__ v_gfeat, float* __restrict__ v_W1,
                   float* __restrict__ v_b1, float* __restrict__ v_W2, float* __restrict__ v_b2) {
    extern __shared__ float dyn[];
    float* sW1 = dyn;                         // [D*D]
    float* sb1 = sW1 + D * D;                 // [D]
    float* sW2 = sb1 + D;                     // [7*D]
    float* sb2 = sW2 + 7 * D;                 // [8]
    float (*sX)[D + 1] = reinterpret_cast<float (*)[D + 1]>(sb2 + 8);     // x of every Gaussian of the CTA (padded)
    float (*sA)[D + 1] = sX + TPB;                                        // v_a
    float (*sH)[D + 1] = sA + TPB;                                        // h
    float (*sO)[8] = reinterpret_cast<float (*)[8]>(sH + TPB);            // v_o
    load_weights(m, sW1, sb1, sW2, sb2);
    __syncthreads();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int t = threadIdx.x;
    float x[D], va[D];
    float vo[7];
#pragma unroll
    for (int k = 0; k < D; ++k) { x[k] = 0.f; va[k] = 0.f; }
#pragma unroll
    for (int c = 0; c < 7; ++c) vo[c] = 0.f;
    long long cls = 0;
    if (i < N) {
        cls = cls_id[i];
        gather_x<D>(m, gfeat, lfeat, cls, i, x);
        float o[7], a[D];
#pragma unroll
        for (int c = 0; c < 7; ++c) o[c] = sb2[c];
#pragma unroll
        for (int j = 0; j < D; ++j) {
            float acc = sb1[j];
#pragma unroll
            for (int k = 0; k < D; ++k) acc += sW1[j * D + k] * x[k];
            a[j] = acc;
            const float h = fmaxf(acc, 0.f);
            sH[t][j] = h;
#pragma unroll
            for (int c = 0; c < 7; ++c) o[c] += sW2[c * D + j] * h;
        }
        // scale path
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float sg = 1.0f / (1.0f + expf(-o[c]));
            const float g = v_scale_out[3 * i + c];
            v_scaling[3 * i + c] = g * sg;
            vo[c] = g * scaling[3 * i + c] * sg * (1.0f - sg);
        }
        // normalize path
        float rot[4], r[4], vy[4], n2 = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) { rot[c] = rotation[4 * i + c]; r[c] = rot[c] * o[3 + c]; n2 += r[c] * r[c]; vy[c] = v_rot_out[4 * i + c]; }
        const float n = sqrtf(n2);
        float vr[4];
        if (n > 1e-12f) {
            const float inv = 1.0f / n;
            float dot = 0.f;
#pragma unroll
            for (int c = 0; c < 4; ++c) dot += vy[c] * r[c] * inv;
#pragma unroll
            for (int c = 0; c < 4; ++c) vr[c] = (vy[c] - dot * r[c] * inv) * inv;
        } else {
#pragma unroll
            for (int c = 0; c < 4; ++c) vr[c] = vy[c] * 1e12f;     // F.normalize clamps the norm at eps
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) { v_rotation[4 * i + c] = vr[c] * o[3 + c]; vo[3 + c] = vr[c] * rot[c]; }
        // hidden layer
#pragma unroll
        for (int j = 0; j < D; ++j) {
            float vh = 0.f;
#pragma unroll
            for (int c = 0; c < 7; ++c) vh += sW2[c * D + j] * vo[c];
            va[j] = a[j] > 0.f ? vh : 0.f;
        }
        // input gradient: v_x = W1^T v_a
#pragma unroll
        for (int k = 0; k < D; ++k) {
            float vx = 0.f;
#pragma unroll
            for (int j = 0; j < D; ++j) vx += sW1[j * D + k] * va[j];
            if (k < m.Fg) atomicAdd(v_gfeat + cls * m.Fg + k, vx);
            else v_lfeat[i * m.Fl + (k - m.Fg)] = vx;
        }
    } else {
#pragma unroll
        for (int j = 0; j < D; ++j) sH[t][j] = 0.f;
    }
#pragma unroll
    for (int k = 0; k < D; ++k) { sX[t][k] = x[k]; sA[t][k] = va[k]; }
#pragma unroll
    for (int c = 0; c < 7; ++c) sO[t][c] = vo[c];
    __syncthreads();
    // weight gradients of this CTA: v_W1[j][k] = sum_t va[t][j] x[t][k];  v_W2[c][j] = sum_t vo[t][c] h[t][j]
    for (int e = t; e < D * D; e += TPB) {
        const int j = e / D, k = e % D;
        float acc = 0.f;
        for (int r = 0; r < TPB; ++r) acc += sA[r][j] * sX[r][k];
        if (acc != 0.f) atomicAdd(v_W1 + e, acc);
    }
    for (int e = t; e < 7 * D; e += TPB) {
        const int c = e / D, j = e % D;
        float acc = 0.f;
        for (int r = 0; r < TPB; ++r) acc += sO[r][c] * sH[r][j];
        if (acc != 0.f) atomicAdd(v_W2 + e, acc);
    }
    if (t < D) {
        float acc = 0.f;
        for (int r = 0; r < TPB; ++r) acc += sA[r][t];
        if (acc != 0.f) atomicAdd(v_b1 + t, acc);
    }
    if (t < 7) {
        float acc = 0.f;
        for (int r = 0; r < TPB; ++r) acc += sO[r][t];
        if (acc != 0.f) atomicAdd(v_b2 + t, acc);
    }
}

}  // namespace

// global_feat [G,Fg], local_feat [N,Fl], cls_id int64 [N], W1 [D,D], b1 [D], W2 [7,D], b2 [7] with D = Fg+Fl in {32, 64}.
ADB_API int adb_cov_mlp_forward(long long N, int Fg, int Fl, const float* global_feat, const float* local_feat,
                                const long long* cls_id, const float* W1, const float* b1, const float* W2,
                                const float* b2, const float* scaling, const float* rotation, float* scale_out,
                                float* rot_out, cudaStream_t stream) {
    ADB_REQUIRE(N >= 0 && Fg >= 0 && Fl >= 0, "adb_cov_mlp_forward: bad sizes");
    const int D = Fg + Fl;
    ADB_REQUIRE(D == 32 || D == 64, "adb_cov_mlp_forward: global_feat_dim + local_feat_dim must be 32 or 64");
    if (N == 0) return ADB_OK;
    ADB_REQUIRE(global_feat && local_feat && cls_id && W1 && b1 && W2 && b2 && scaling && rotation && scale_out && rot_out,
                "adb_cov_mlp_forward: null pointer");
    CovMlp m{D, Fg, Fl, W1, b1, W2, b2};
    const int blocks = adb_cdiv(N, TPB);
    if (D == 32) cov_mlp_fwd_kernel<32><<<blocks, TPB, 0, stream>>>(N, m, global_feat, local_feat, cls_id, scaling, rotation, scale_out, rot_out);
    else cov_mlp_fwd_kernel<64><<<blocks, TPB, 0, stream>>>(N, m, global_feat, local_feat, cls_id, scaling, rotation, scale_out, rot_out);
    ADB_CHECK_LAUNCH("cov_mlp_fwd_kernel");
    return ADB_OK;
}

// v_global_feat [G,Fg], v_W1, v_b1, v_W2, v_b2 are ACCUMULATED (caller zeroes them); the per-Gaussian outputs are overwritten.
ADB_API int adb_cov_mlp_backward(long long N, int Fg, int Fl, const float* global_feat, const float* local_feat,
                                 const long long* cls_id, const float* W1, const float* b1, const float* W2,
                                 const float* b2, const float* scaling, const float* rotation,
                                 const float* v_scale_out, const float* v_rot_out, float* v_scaling, float* v_rotation,
                                 float* v_local_feat, float* v_global_feat, float* v_W1, float* v_b1, float* v_W2,
                                 float* v_b2, cudaStream_t stream) {
    ADB_REQUIRE(N >= 0 && Fg >= 0 && Fl >= 0, "adb_cov_mlp_backward: bad sizes");
    const int D = Fg + Fl;
    ADB_REQUIRE(D == 32 || D == 64, "adb_cov_mlp_backward: global_feat_dim + local_feat_dim must be 32 or 64");
    if (N == 0) return ADB_OK;
    ADB_REQUIRE(global_feat && local_feat && cls_id && W1 && b1 && W2 && b2 && scaling && rotation && v_scale_out &&
                    v_rot_out && v_scaling && v_rotation && v_local_feat && v_global_feat && v_W1 && v_b1 && v_W2 && v_b2,
                "adb_cov_mlp_backward: null pointer");
    CovMlp m{D, Fg, Fl, W1, b1, W2, b2};
    const int blocks = adb_cdiv(N, TPB);
    const size_t smem = sizeof(float) * ((size_t)D * D + D + 7 * D + 8 + 3 * (size_t)TPB * (D + 1) + (size_t)TPB * 8);
    if (D == 32) {
        static AdbDeviceOnce once;
        const int rc = once.ensure([smem]() -> int {
            ADB_CUDA(cudaFuncSetAttribute(cov_mlp_bwd_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            return ADB_OK;
        });
        if (rc != ADB_OK) return rc;
        cov_mlp_bwd_kernel<32><<<blocks, TPB, smem, stream>>>(N, m, global_feat, local_feat, cls_id, scaling, rotation, v_scale_out,
                                                             v_rot_out, v_scaling, v_rotation, v_local_feat, v_global_feat, v_W1, v_b1, v_W2, v_b2);
    } else {
        static AdbDeviceOnce once;
        const int rc = once.ensure([smem]() -> int {
            ADB_CUDA(cudaFuncSetAttribute(cov_mlp_bwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            return ADB_OK;
        });
        if (rc != ADB_OK) return rc;
        cov_mlp_bwd_kernel<64><<<blocks, TPB, smem, stream>>>(N, m, global_feat, local_feat, cls_id, scaling, rotation, v_scale_out,
                                                             v_rot_out, v_scaling, v_rotation, v_local_feat, v_global_feat, v_W1, v_b1, v_W2, v_b2);
    }
    ADB_CHECK_LAUNCH("cov_mlp_bwd_kernel");
    return ADB_OK;
}
