// Batched GEMM on the 5th-generation tensor cores: D = act(alpha * A * B^T + bias) + residual.
//
// This is the one dense-contraction engine of the MASt3R path (every nn.Linear / q@k^T / attn@v of
// VSLAM/thirdparty/mast3r/dust3r/croco/models/blocks.py:58-112,140-169, the patch embedding
// dust3r/dust3r/patch_embed.py:19-29 and the local-feature MLP mast3r/catmlp_dpt_head.py:67-69).  The
// reference runs them as fp32 (TF32) cuBLAS calls.
//
// sm_100a structure (hand-written PTX, no CUTLASS):
//   * operands are bf16, K-major, staged by TMA (cp.async.bulk.tensor.3d, SWIZZLE_128B) into a multi-stage
//     shared-memory ring guarded by mbarriers (full/empty);
//   * one elected thread issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=128, K=16) with the fp32
//     accumulator in TENSOR MEMORY (128 lanes x 128 columns), tcgen05.commit releases ring slots;
//   * four epilogue warps read the accumulator back with tcgen05.ld.32x32b.x32, apply alpha / bias / exact-erf
//     GELU / residual and write fp32 and/or a bf16 (hi, lo) split for the next GEMM;
//   * "bf16x3": when the lo operands are given the kernel accumulates A_hi*B_hi + A_hi*B_lo + A_lo*B_hi into the
//     same TMEM accumulator, which carries ~16 mantissa bits per operand (error ~1e-5) — the north-star's 1e-4
//     pointmap tolerance cannot be met by single-pass bf16 or TF32 (the reference's own GPU mode) over 36 layers.
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator, warps 4..7 = epilogue.
#include "common.cuh"
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdio.h>

namespace {

constexpr int BM = 128, BN = 128, BK = 64;    // BK * 2 B = 128 B = one swizzle row
constexpr int TILE_BYTES = BM * BK * 2;       // 16 KB (A and B tiles have the same shape)
constexpr int UMMA_K = 16;
constexpr int NTHREADS = 256;
constexpr int TMEM_COLS = 128;

struct GemmParams {
    int M, N, K;
    int nterms;  // 1 or 3
    int stages;
    float* D; long long ldd, sD;
    __nv_bfloat16* Dhi; __nv_bfloat16* Dlo; long long ldo, sO;
    const float* bias;
    const float* residual; long long ldr, sR;
    float alpha;
    int act;
    int zdiv;  // blockIdx.z = z_outer * zdiv + z_inner; outputs use (z_outer * s?2 + z_inner * s?)
    long long sD2, sO2, sR2;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// K-major, SWIZZLE_128B operand tile: rows of 128 B, 8-row groups 1024 B apart (SBO), LBO unused.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fff);        // start address
    d |= (uint64_t)1 << 16;                            // leading byte offset (ignored for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                  // stride byte offset
    d |= (uint64_t)1 << 46;                            // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                            // SWIZZLE_128B
    return d;
}

// kind::f16, A = B = BF16 (K-major), D = F32, M = 128, N = 128
__device__ __forceinline__ uint32_t make_idesc() {
    uint32_t d = 0;
    d |= 1u << 4;                 // c_format = F32
    d |= 1u << 7;                 // a_format = BF16
    d |= 1u << 10;                // b_format = BF16
    d |= (uint32_t)(BN >> 3) << 17;
    d |= (uint32_t)(BM >> 4) << 24;
    return d;
}

__device__ __forceinline__ float gelu_erf(float v) { return 0.5f * v * (1.0f + erff(v * 0.70710678118654752f)); }

__global__ void __launch_bounds__(NTHREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mapAhi, const __grid_constant__ CUtensorMap mapAlo,
               const __grid_constant__ CUtensorMap mapBhi, const __grid_constant__ CUtensorMap mapBlo,
               const GemmParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // carve: [stages][4 tiles] | barriers | tmem ptr
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int tiles_per_stage = p.nterms == 3 ? 4 : 2;
    const uint32_t stage_bytes = tiles_per_stage * TILE_BYTES;
    uint64_t* full_bar = (uint64_t*)(smem + (size_t)p.stages * stage_bytes);
    uint64_t* empty_bar = full_bar + p.stages;
    uint64_t* tmem_full_bar = empty_bar + p.stages;
    uint32_t* tmem_ptr = (uint32_t*)(tmem_full_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN, bz = blockIdx.z;
    const int num_kb = (p.K + BK - 1) / BK;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapAhi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapBhi) : "memory");
        if (p.nterms == 3) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&mapAlo) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&mapBlo) : "memory");
        }
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 1); }
        mbar_init(tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ===== TMA producer =====
        if (elect_one()) {
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % p.stages;
                const uint32_t ph = (kb / p.stages) & 1;
                mbar_wait(empty_bar + s, ph ^ 1);
                uint8_t* st = smem + (size_t)s * stage_bytes;
                mbar_expect_tx(full_bar + s, stage_bytes);
                tma_load_3d(st, &mapAhi, full_bar + s, kb * BK, m0, bz);
                tma_load_3d(st + TILE_BYTES, &mapBhi, full_bar + s, kb * BK, n0, bz);
                if (p.nterms == 3) {
                    tma_load_3d(st + 2 * TILE_BYTES, &mapAlo, full_bar + s, kb * BK, m0, bz);
                    tma_load_3d(st + 3 * TILE_BYTES, &mapBlo, full_bar + s, kb * BK, n0, bz);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        const uint32_t idesc = make_idesc();
        for (int kb = 0; kb < num_kb; ++kb) {
            const int s = kb % p.stages;
            const uint32_t ph = (kb / p.stages) & 1;
            mbar_wait(full_bar + s, ph);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t st = smem_u32(smem + (size_t)s * stage_bytes);
                const uint64_t dAhi = make_smem_desc(st), dBhi = make_smem_desc(st + TILE_BYTES);
                const uint64_t dAlo = make_smem_desc(st + 2 * TILE_BYTES), dBlo = make_smem_desc(st + 3 * TILE_BYTES);
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k) {
                    const uint64_t adv = (uint64_t)((k * UMMA_K * 2) >> 4);  // +32 B per K step, in 16 B units
                    tc_mma(tmem_base, dAhi + adv, dBhi + adv, idesc, (kb | k) ? 1u : 0u);
                    if (p.nterms == 3) {
                        tc_mma(tmem_base, dAhi + adv, dBlo + adv, idesc, 1u);
                        tc_mma(tmem_base, dAlo + adv, dBhi + adv, idesc, 1u);
                    }
                }
                tc_commit(empty_bar + s);                          // slot free once these MMAs retire
                if (kb == num_kb - 1) tc_commit(tmem_full_bar);    // accumulator complete
            }
            __syncwarp();
        }
    } else if (warp >= 4) {
        // ===== epilogue: TMEM -> registers -> global =====
        const int q = warp & 3;                 // TMEM lane quarter this warp may access
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
        const int row = m0 + q * 32 + lane;
        const bool row_ok = row < p.M;
        const int zo = bz / p.zdiv, zi = bz % p.zdiv;
        const size_t drow = (size_t)zo * p.sD2 + (size_t)zi * p.sD + (size_t)row * p.ldd;
        const size_t orow = (size_t)zo * p.sO2 + (size_t)zi * p.sO + (size_t)row * p.ldo;
        const size_t rrow = (size_t)zo * p.sR2 + (size_t)zi * p.sR + (size_t)row * p.ldr;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t r[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                  "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                  "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                  "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(taddr) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (row_ok) {
                const int ncol = min(32, p.N - (n0 + c0));
                if (ncol > 0) {
                    float v[32];
#pragma unroll
                    for (int jj = 0; jj < 32; ++jj) {
                        float x = __uint_as_float(r[jj]) * p.alpha;
                        const int col = n0 + c0 + jj;
                        if (jj < ncol) {
                            if (p.bias) x += __ldg(p.bias + col);
                            if (p.act == 1) x = gelu_erf(x);
                            if (p.residual) x += __ldg(p.residual + rrow + col);
                        }
                        v[jj] = x;
                    }
                    if (ncol == 32 && ((p.ldd | (n0 + c0)) % 4 == 0) && ((p.ldo | (n0 + c0)) % 8 == 0)) {
                        if (p.D) {
                            float4* dp = reinterpret_cast<float4*>(p.D + drow + n0 + c0);
#pragma unroll
                            for (int jj = 0; jj < 8; ++jj) dp[jj] = make_float4(v[4 * jj], v[4 * jj + 1], v[4 * jj + 2], v[4 * jj + 3]);
                        }
                        if (p.Dhi) {
                            uint4* hp = reinterpret_cast<uint4*>(p.Dhi + orow + n0 + c0);
                            uint4* lp = p.Dlo ? reinterpret_cast<uint4*>(p.Dlo + orow + n0 + c0) : nullptr;
#pragma unroll
                            for (int jj = 0; jj < 4; ++jj) {
                                uint32_t hw[4], lw[4];
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const float a = v[8 * jj + 2 * e], b = v[8 * jj + 2 * e + 1];
                                    const __nv_bfloat16 ah = __float2bfloat16_rn(a), bh = __float2bfloat16_rn(b);
                                    const __nv_bfloat16 al = __float2bfloat16_rn(a - __bfloat162float(ah));
                                    const __nv_bfloat16 bl = __float2bfloat16_rn(b - __bfloat162float(bh));
                                    hw[e] = (uint32_t)__bfloat16_as_ushort(ah) | ((uint32_t)__bfloat16_as_ushort(bh) << 16);
                                    lw[e] = (uint32_t)__bfloat16_as_ushort(al) | ((uint32_t)__bfloat16_as_ushort(bl) << 16);
                                }
                                hp[jj] = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                                if (lp) lp[jj] = make_uint4(lw[0], lw[1], lw[2], lw[3]);
                            }
                        }
                    } else {
                        for (int jj = 0; jj < ncol; ++jj) {
                            const int col = n0 + c0 + jj;
                            if (p.D) p.D[drow + col] = v[jj];
                            if (p.Dhi) {
                                const __nv_bfloat16 h = __float2bfloat16_rn(v[jj]);
                                p.Dhi[orow + col] = h;
                                if (p.Dlo) p.Dlo[orow + col] = __float2bfloat16_rn(v[jj] - __bfloat162float(h));
                            }
                        }
                    }
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

// ---- host side: TMA descriptors through the driver entry point (no libcuda link dependency) ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// bf16 tensor [batch][rows][K] with element strides (ld, batch_stride); box = 64 (K) x 128 (rows) x 1
int make_map(CUtensorMap* map, const void* ptr, int rows, int K, long long ld, long long bstride, int batch) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) { adb_set_error_msg("cuTensorMapEncodeTiled driver entry point unavailable"); return ADB_ERR_CUDA; }
    cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)batch};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)(batch > 1 ? bstride : (long long)rows * ld) * 2};
    cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)BM, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        static char msg[160];
        snprintf(msg, sizeof(msg), "cuTensorMapEncodeTiled failed (%d) rows=%d K=%d ld=%lld", (int)r, rows, K, ld);
        adb_set_error_msg(msg);
        return ADB_ERR_INVALID;
    }
    return ADB_OK;
}

}  // namespace

// D[b] (fp32, optional) / Dhi,Dlo[b] (bf16 split, optional) = act(alpha * A[b] * B[b]^T + bias) + residual[b]
//   A: bf16 [batch][M][K] (ld = lda elements, batch stride sA), B: bf16 [batch][N][K] (ldb, sB); both K-contiguous,
//   16-byte aligned, lda/ldb multiples of 8.  A_lo and B_lo both non-NULL selects the 3-term bf16x3 product.
//   act: 0 none, 1 GELU (erf).  Output batch addressing: z = z_outer * zdiv + z_inner and the output/residual
//   offset is z_outer * s?2 + z_inner * s? (zdiv <= 0: plain z * s?), which lets a [B*h] attention batch write
//   straight into a [B, N, h*64] activation.
ADB_API int adb_gemm_bf16(int batch, int M, int N, int K, const void* A_hi, const void* A_lo, long long lda,
                          long long sA, const void* B_hi, const void* B_lo, long long ldb, long long sB, float* D,
                          long long ldd, long long sD, void* D_hi, void* D_lo, long long ldo, long long sO,
                          const float* bias, const float* residual, long long ldr, long long sR, float alpha, int act,
                          int zdiv, long long sD2, long long sO2, long long sR2, cudaStream_t stream) {
    ADB_REQUIRE(batch >= 1 && M >= 1 && N >= 1 && K >= 1, "adb_gemm_bf16: bad sizes");
    ADB_REQUIRE(A_hi && B_hi && (D || D_hi), "adb_gemm_bf16: null pointer");
    ADB_REQUIRE((A_lo == nullptr) == (B_lo == nullptr), "adb_gemm_bf16: A_lo and B_lo must be given together");
    ADB_REQUIRE(lda % 8 == 0 && ldb % 8 == 0 && sA % 8 == 0 && sB % 8 == 0, "adb_gemm_bf16: strides must be multiples of 8 elements");
    ADB_REQUIRE(((uintptr_t)A_hi % 16 == 0) && ((uintptr_t)B_hi % 16 == 0), "adb_gemm_bf16: operands must be 16-byte aligned");
    ADB_REQUIRE(act == 0 || act == 1, "adb_gemm_bf16: unknown activation");
    const int nterms = A_lo ? 3 : 1;
    CUtensorMap mAhi, mAlo, mBhi, mBlo;
    int rc;
    // a broadcast weight (sB == 0) is described as a batch of size 1 and always read at batch coordinate 0 ...
    if ((rc = make_map(&mAhi, A_hi, M, K, lda, sA, batch))) return rc;
    if ((rc = make_map(&mBhi, B_hi, N, K, ldb, sB ? sB : (long long)N * ldb, sB ? batch : 1))) return rc;
    mAlo = mAhi; mBlo = mBhi;
    if (nterms == 3) {
        if ((rc = make_map(&mAlo, A_lo, M, K, lda, sA, batch))) return rc;
        if ((rc = make_map(&mBlo, B_lo, N, K, ldb, sB ? sB : (long long)N * ldb, sB ? batch : 1))) return rc;
    }
    ADB_REQUIRE(sB != 0 || batch == 1, "adb_gemm_bf16: weight broadcast over batch > 1: fold the batch into M instead");
    GemmParams p;
    p.M = M; p.N = N; p.K = K; p.nterms = nterms;
    p.stages = nterms == 3 ? 3 : 6;
    p.D = D; p.ldd = ldd; p.sD = sD;
    p.Dhi = (__nv_bfloat16*)D_hi; p.Dlo = (__nv_bfloat16*)D_lo; p.ldo = ldo; p.sO = sO;
    p.bias = bias; p.residual = residual; p.ldr = ldr; p.sR = sR;
    p.alpha = alpha; p.act = act;
    p.zdiv = zdiv > 0 ? zdiv : (1 << 30); p.sD2 = sD2; p.sO2 = sO2; p.sR2 = sR2;
    const size_t smem = (size_t)p.stages * (nterms == 3 ? 4 : 2) * TILE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
    static bool attr_set = false;
    if (!attr_set) {
        ADB_CUDA(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set = true;
    }
    dim3 grid(adb_cdiv(N, BN), adb_cdiv(M, BM), batch);
    gemm_tc_kernel<<<grid, NTHREADS, smem, stream>>>(mAhi, mAlo, mBhi, mBlo, p);
    ADB_CHECK_LAUNCH("gemm_tc_kernel");
    return ADB_OK;
}
