// Batched GEMM / implicit-GEMM 3x3 convolution on the 5th-generation tensor cores:
//     D = act(alpha * A * B^T + bias) + residual
//
// This is the one dense-contraction engine of the MASt3R path: every nn.Linear / q@k^T / attn@v of
// VSLAM/thirdparty/mast3r/dust3r/croco/models/blocks.py:58-112,140-169, the patch embedding
// dust3r/dust3r/patch_embed.py:19-29, the local-feature MLP mast3r/catmlp_dpt_head.py:67-69 and the 3x3 / 1x1
// convolutions of the DPT head croco/models/dpt_block.py:79-142,356-410.  The reference runs them as fp32 (TF32)
// cuBLAS / cuDNN calls.
//
// sm_100a structure (hand-written PTX, no CUTLASS):
//   * PERSISTENT: one CTA per SM walks a static round-robin list of 128x128 output tiles;
//   * operands are bf16, K-major, staged by TMA (cp.async.bulk.tensor, SWIZZLE_128B) into a multi-stage
//     shared-memory ring guarded by full/empty mbarriers; the producer runs ahead across tile boundaries;
//   * one elected thread issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=128, K=16); the fp32 accumulator
//     lives in TENSOR MEMORY, DOUBLE-BUFFERED (2 x 128 columns) so the epilogue of tile i overlaps the main
//     loop of tile i+1; tcgen05.commit releases ring slots and publishes finished accumulators;
//   * eight epilogue warps (two per TMEM lane quarter, 64 columns each) read the accumulator with
//     tcgen05.ld.32x32b.x32, PROMOTE partial sums to round-to-nearest fp32 registers every 16 k-blocks (the tensor
//     core's accumulator truncates: its error grows linearly with the chain length), then apply alpha / bias /
//     GELU(erf) or ReLU / residual and write fp32 and/or a bf16 (hi, lo) split, one full 128 B line per thread;
//   * "bf16x3": with the lo operands the kernel accumulates A_hi*B_hi + A_hi*B_lo + A_lo*B_hi into the same
//     TMEM accumulator (~16 mantissa bits per operand, error ~1e-5): single-pass bf16 or TF32 (the reference's
//     own GPU mode) cannot hold the north-star's 1e-4 pointmap tolerance over 36 layers;
//   * CONV MODE: A is an NHWC activation described by a 4-D tensor map; a tile is a Wb x Hb pixel rectangle
//     and K-block (tap, channel chunk) is ONE shifted TMA box load — out-of-image taps are zero-filled by the
//     TMA unit, which is exactly the conv's zero padding; no im2col buffer exists.
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator, warps 4..11 = epilogue.
#include "common.cuh"
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdio.h>

namespace {

constexpr int BM = 128, BN = 128, BK = 64;    // BK * 2 B = 128 B = one swizzle row
constexpr int TILE_BYTES = BM * BK * 2;       // 16 KB (A and B tiles have the same shape)
constexpr int UMMA_K = 16;
constexpr int NTHREADS = 384;                 // 4 control warps + 8 epilogue warps
constexpr int TMEM_COLS = 256;                // two 128-column accumulators
constexpr int MAX_STAGES = 8;
constexpr int KCHUNK_KB = 8;                 // k-blocks (of 64) accumulated in TMEM before promotion to registers

struct GemmParams {
    int M, N, K;
    int batch;
    int nterms;  // 1 or 3
    int stages;
    float* D; long long ldd, sD;
    __nv_bfloat16* Dhi; __nv_bfloat16* Dlo; long long ldo, sO;
    const float* bias;
    const float* residual; long long ldr, sR;
    float alpha;
    int act;          // 0 none, 1 GELU(erf), 2 ReLU
    int split_relu;   // the bf16 split output stores relu(value) (fp32 D keeps the raw value)
    int zdiv;         // z = z_outer * zdiv + z_inner; outputs use (z_outer * s?2 + z_inner * s?)
    long long sD2, sO2, sR2;
    // conv mode (conv_wb > 0): M tile = conv_hb x conv_wb pixel rectangle of an H x W image, K block = (tap, chunk)
    int conv_wb, conv_hb, conv_H, conv_W, conv_chunks, conv_tiles_x, conv_tiles_y;
    // RoPE epilogue (rope_ndst > 0): output columns [0, rope_ndst * rope_C) are attention heads (64 wide); each is rotated
    // with RoPE2D for its token's (y, x) position and written as a bf16 split in head-major layout [B*heads, ntok, 64] to
    // rope_hi/lo[col / rope_C]; the remaining columns take the ordinary D path (D is pre-shifted by the caller).
    int rope_ndst, rope_C, rope_heads, rope_ntok, rope_npos;
    const long long* rope_pos;      // int64 [M, 2] (y, x)
    const float* rope_table;        // [n_pos][16][2] (cos, sin)
    __nv_bfloat16* rope_hi[2]; __nv_bfloat16* rope_lo[2];
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
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

// 256-bit global accesses (sm_100: LDG/STG.E.ENL2.256): one full 32 B sector per lane per instruction
__device__ __forceinline__ void st_v8(float* p, const float* v) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
                 "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
}
__device__ __forceinline__ void st_v8_b32(void* p, const uint32_t* w) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]),
                 "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
}
__device__ __forceinline__ void ld_v8_nc(const float* p, float* v) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]),
                 "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(p));
}

struct TileCoord {
    int bz, mt, nt;
};
__device__ __forceinline__ TileCoord decode_tile(const GemmParams& p, int tile, int tiles_m, int tiles_n) {
    TileCoord c;
    c.nt = tile % tiles_n;
    const int r = tile / tiles_n;
    c.mt = r % tiles_m;
    c.bz = r / tiles_m;
    return c;
}

// ROPE: the epilogue variant of adb_gemm_bf16_rope (own instantiation so that its extra live values do not cost the plain
// GEMM registers); it has no residual / activation / split-output paths.
template <bool ROPE>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mapAhi, const __grid_constant__ CUtensorMap mapAlo,
               const __grid_constant__ CUtensorMap mapBhi, const __grid_constant__ CUtensorMap mapBlo,
               const GemmParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int tiles_per_stage = p.nterms == 3 ? 4 : 2;
    const uint32_t stage_bytes = tiles_per_stage * TILE_BYTES;
    uint64_t* full_bar = (uint64_t*)(smem + (size_t)p.stages * stage_bytes);
    uint64_t* empty_bar = full_bar + MAX_STAGES;
    uint64_t* tmem_full_bar = empty_bar + MAX_STAGES;   // [2]
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;       // [2]
    uint32_t* tmem_ptr = (uint32_t*)(tmem_empty_bar + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool conv = p.conv_wb > 0;
    const int tiles_m = conv ? p.conv_tiles_x * p.conv_tiles_y : (p.M + BM - 1) / BM;
    const int tiles_n = (p.N + BN - 1) / BN;
    const int num_tiles = tiles_m * tiles_n * p.batch;
    const int num_kb = conv ? 9 * p.conv_chunks : (p.K + BK - 1) / BK;

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
        for (int b = 0; b < 2; ++b) { mbar_init(tmem_full_bar + b, 1); mbar_init(tmem_empty_bar + b, 8); }
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
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const TileCoord tc = decode_tile(p, tile, tiles_m, tiles_n);
                const int n0 = tc.nt * BN;
                int m0 = tc.mt * BM, x0 = 0, y0 = 0;
                if (conv) { x0 = (tc.mt % p.conv_tiles_x) * p.conv_wb; y0 = (tc.mt / p.conv_tiles_x) * p.conv_hb; }
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % p.stages;
                    const uint32_t ph = (it / p.stages) & 1;
                    mbar_wait(empty_bar + s, ph ^ 1);
                    uint8_t* st = smem + (size_t)s * stage_bytes;
                    mbar_expect_tx(full_bar + s, stage_bytes);
                    if (conv) {
                        const int tap = kb / p.conv_chunks, ch = kb - tap * p.conv_chunks;
                        const int cx = x0 + (tap % 3) - 1, cy = y0 + (tap / 3) - 1;
                        tma_load_4d(st, &mapAhi, full_bar + s, ch * BK, cx, cy, tc.bz);
                        if (p.nterms == 3) tma_load_4d(st + 2 * TILE_BYTES, &mapAlo, full_bar + s, ch * BK, cx, cy, tc.bz);
                        tma_load_3d(st + TILE_BYTES, &mapBhi, full_bar + s, kb * BK, n0, 0);
                        if (p.nterms == 3) tma_load_3d(st + 3 * TILE_BYTES, &mapBlo, full_bar + s, kb * BK, n0, 0);
                    } else {
                        tma_load_3d(st, &mapAhi, full_bar + s, kb * BK, m0, tc.bz);
                        tma_load_3d(st + TILE_BYTES, &mapBhi, full_bar + s, kb * BK, n0, tc.bz);
                        if (p.nterms == 3) {
                            tma_load_3d(st + 2 * TILE_BYTES, &mapAlo, full_bar + s, kb * BK, m0, tc.bz);
                            tma_load_3d(st + 3 * TILE_BYTES, &mapBlo, full_bar + s, kb * BK, n0, tc.bz);
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        const uint32_t idesc = make_idesc();
        uint32_t it = 0, lc = 0;   // lc counts (tile, K-chunk) units: each owns one TMEM accumulator buffer
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            for (int kb0 = 0; kb0 < num_kb; kb0 += KCHUNK_KB, ++lc) {
                const uint32_t buf = lc & 1;
                mbar_wait(tmem_empty_bar + buf, ((lc >> 1) & 1) ^ 1);   // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + buf * BN;
                const int kb1 = min(kb0 + KCHUNK_KB, num_kb);
                for (int kb = kb0; kb < kb1; ++kb, ++it) {
                    const int s = it % p.stages;
                    const uint32_t ph = (it / p.stages) & 1;
                    mbar_wait(full_bar + s, ph);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t st = smem_u32(smem + (size_t)s * stage_bytes);
                        const uint64_t dAhi = make_smem_desc(st), dBhi = make_smem_desc(st + TILE_BYTES);
                        const uint64_t dAlo = make_smem_desc(st + 2 * TILE_BYTES), dBlo = make_smem_desc(st + 3 * TILE_BYTES);
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k) {
                            const uint64_t adv = (uint64_t)((k * UMMA_K * 2) >> 4);  // +32 B per K step, in 16 B units
                            tc_mma(tmem_d, dAhi + adv, dBhi + adv, idesc, ((kb - kb0) | k) ? 1u : 0u);
                            if (p.nterms == 3) {
                                tc_mma(tmem_d, dAhi + adv, dBlo + adv, idesc, 1u);
                                tc_mma(tmem_d, dAlo + adv, dBhi + adv, idesc, 1u);
                            }
                        }
                        tc_commit(empty_bar + s);                            // slot free once these MMAs retire
                        if (kb == kb1 - 1) tc_commit(tmem_full_bar + buf);   // this chunk's partial sum is complete
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue (8 warps): TMEM -> fp32 registers (K-chunk promotion) -> global =====
        // warp e = warp-4: TMEM lane quarter q = e % 4 (hardware restriction), column half = e / 4 (64 columns each).
        const int e = warp - 4;
        const int q = e & 3, half = e >> 2;
        uint32_t lc = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const TileCoord tc = decode_tile(p, tile, tiles_m, tiles_n);
            const int n0 = tc.nt * BN + half * 64, m0 = tc.mt * BM, bz = tc.bz;
            int x0 = 0, y0 = 0;
            if (conv) { x0 = (tc.mt % p.conv_tiles_x) * p.conv_wb; y0 = (tc.mt / p.conv_tiles_x) * p.conv_hb; }
            const int zo = bz / p.zdiv, zi = bz % p.zdiv;
            const size_t dbase = (size_t)zo * p.sD2 + (size_t)zi * p.sD;
            const size_t obase = (size_t)zo * p.sO2 + (size_t)zi * p.sO;
            const size_t rbase = (size_t)zo * p.sR2 + (size_t)zi * p.sR;
            // 256-bit accesses need 32 B alignment of every row segment
            const bool vec_ok = (p.ldd % 8 == 0) && (p.ldr % 8 == 0) && (p.ldo % 16 == 0) &&
                                (((uintptr_t)p.D | (uintptr_t)p.residual | (uintptr_t)p.Dhi | (uintptr_t)p.Dlo) % 32 == 0) &&
                                ((dbase | rbase) % 8 == 0) && (obase % 16 == 0);
            // The tensor core's fp32 accumulator truncates, so its error grows linearly with the length of the
            // accumulation chain (measured: 1.2e-6 at K=1024, 2e-5 at K=16384).  Partial sums are therefore taken
            // out of TMEM every KCHUNK_KB k-blocks and added here in round-to-nearest fp32 registers.
            float acc[2][32];
            for (int kb0 = 0; kb0 < num_kb; kb0 += KCHUNK_KB, ++lc) {
                const uint32_t buf = lc & 1;
                mbar_wait(tmem_full_bar + buf, (lc >> 1) & 1);
                tc_fence_after();
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    uint32_t r[32];
                    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + half * 64 + g * 32);
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
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc[g][j] = kb0 == 0 ? __uint_as_float(r[j]) : acc[g][j] + __uint_as_float(r[j]);
                }
                tc_fence_before();
                if (lane == 0) mbar_arrive(tmem_empty_bar + buf);   // hand the buffer back to the MMA warp
            }
            // lane == accumulator row: each thread owns 2 x 32 consecutive columns of one output row (two full 128 B
            // lines), so every 128-bit load/store below moves whole sectors.
            const int trow = q * 32 + lane;
            long long row;
            if (conv) {
                const int yy = y0 + trow / p.conv_wb, xx = x0 + trow % p.conv_wb;
                row = (yy < p.conv_H && xx < p.conv_W) ? (long long)yy * p.conv_W + xx : -1;
            } else {
                row = m0 + trow < p.M ? m0 + trow : -1;
            }
            if (row < 0) continue;
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                const int c0 = n0 + g * 32;
                if (c0 >= p.N) continue;
                float (&v)[32] = acc[g];      // finished accumulators are transformed in place
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    float t = acc[g][j] * p.alpha;
                    if (p.bias && c0 + j < p.N) t += __ldg(p.bias + c0 + j);
                    if (!ROPE) {
                        if (p.act == 1) t = gelu_erf(t);
                        else if (p.act == 2) t = fmaxf(t, 0.f);
                    }
                    v[j] = t;
                }
                if (ROPE && c0 < p.rope_ndst * p.rope_C) {
                    // this thread holds dims [d0, d0+32) of one head of token `row`: one RoPE axis (y: d0 = 0, x: d0 = 32),
                    // pairs (i, i+16) -- blocks.py:101-103 / pos_embed.py:129-159, fused with the head split
                    const int sec = c0 / p.rope_C, cc = c0 - sec * p.rope_C;
                    const int hh = cc >> 6, axis = (cc >> 5) & 1;
                    long long pz = __ldg(p.rope_pos + row * 2 + axis);
                    pz = pz < 0 ? 0 : (pz >= p.rope_npos ? p.rope_npos - 1 : pz);
                    const float4* tb = reinterpret_cast<const float4*>(p.rope_table + pz * 32);
#pragma unroll
                    for (int i4 = 0; i4 < 8; ++i4) {
                        const float4 cs = __ldg(tb + i4);          // (cos, sin) of frequencies 2*i4 and 2*i4+1
                        const int i = 2 * i4;
                        const float u0 = v[i], w0 = v[i + 16], u1 = v[i + 1], w1 = v[i + 17];
                        v[i] = u0 * cs.x - w0 * cs.y;      v[i + 16] = w0 * cs.x + u0 * cs.y;
                        v[i + 1] = u1 * cs.z - w1 * cs.w;  v[i + 17] = w1 * cs.z + u1 * cs.w;
                    }
                    const long long bq = row / p.rope_ntok, nq = row - bq * p.rope_ntok;
                    const size_t off = (((size_t)bq * p.rope_heads + hh) * p.rope_ntok + nq) * 64 + axis * 32;
                    __nv_bfloat16* hp = p.rope_hi[sec] + off;
                    __nv_bfloat16* lp = p.rope_lo[sec] ? p.rope_lo[sec] + off : nullptr;
#pragma unroll
                    for (int h8 = 0; h8 < 2; ++h8) {               // 16 values -> one 32-byte store of hi and of lo
                        uint32_t hw[8], lw[8];
#pragma unroll
                        for (int k2 = 0; k2 < 8; ++k2) {
                            const float a = v[16 * h8 + 2 * k2], b = v[16 * h8 + 2 * k2 + 1];
                            const __nv_bfloat162 h2 = __floats2bfloat162_rn(a, b);
                            const uint32_t hb = *reinterpret_cast<const uint32_t*>(&h2);
                            const __nv_bfloat162 l2 = __floats2bfloat162_rn(a - __uint_as_float(hb << 16),
                                                                            b - __uint_as_float(hb & 0xffff0000u));
                            hw[k2] = hb;
                            lw[k2] = *reinterpret_cast<const uint32_t*>(&l2);
                        }
                        st_v8_b32(hp + 16 * h8, hw);
                        if (lp) st_v8_b32(lp + 16 * h8, lw);
                    }
                    continue;
                }
                if (vec_ok && c0 + 32 <= p.N) {
                    if (!ROPE && p.residual) {
                        const float* rp = p.residual + rbase + (size_t)row * p.ldr + c0;
#pragma unroll
                        for (int j8 = 0; j8 < 4; ++j8) {
                            float rs[8];
                            ld_v8_nc(rp + 8 * j8, rs);
#pragma unroll
                            for (int k2 = 0; k2 < 8; ++k2) v[8 * j8 + k2] += rs[k2];
                        }
                    }
                    if (p.D) {
                        float* dp = p.D + dbase + (size_t)row * p.ldd + c0;
#pragma unroll
                        for (int j8 = 0; j8 < 4; ++j8) st_v8(dp + 8 * j8, v + 8 * j8);
                    }
                    if (!ROPE && p.Dhi) {
                        __nv_bfloat16* hp = p.Dhi + obase + (size_t)row * p.ldo + c0;
                        __nv_bfloat16* lp = p.Dlo ? p.Dlo + obase + (size_t)row * p.ldo + c0 : nullptr;
#pragma unroll
                        for (int h8 = 0; h8 < 2; ++h8) {
                            uint32_t hw[8], lw[8];
#pragma unroll
                            for (int k2 = 0; k2 < 8; ++k2) {
                                float a = v[16 * h8 + 2 * k2], b = v[16 * h8 + 2 * k2 + 1];
                                if (p.split_relu) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
                                const __nv_bfloat162 h2 = __floats2bfloat162_rn(a, b);
                                const uint32_t hb = *reinterpret_cast<const uint32_t*>(&h2);
                                const __nv_bfloat162 l2 = __floats2bfloat162_rn(a - __uint_as_float(hb << 16),
                                                                                b - __uint_as_float(hb & 0xffff0000u));
                                hw[k2] = hb;
                                lw[k2] = *reinterpret_cast<const uint32_t*>(&l2);
                            }
                            st_v8_b32(hp + 16 * h8, hw);
                            if (lp) st_v8_b32(lp + 16 * h8, lw);
                        }
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        if (c0 + j < p.N) {
                            float t = v[j];
                            if (!ROPE && p.residual) t += __ldg(p.residual + rbase + (size_t)row * p.ldr + c0 + j);
                            if (p.D) p.D[dbase + (size_t)row * p.ldd + c0 + j] = t;
                            if (!ROPE && p.Dhi) {
                                const float ts = p.split_relu ? fmaxf(t, 0.f) : t;
                                const __nv_bfloat16 h = __float2bfloat16_rn(ts);
                                p.Dhi[obase + (size_t)row * p.ldo + c0 + j] = h;
                                if (p.Dlo) p.Dlo[obase + (size_t)row * p.ldo + c0 + j] = __float2bfloat16_rn(ts - __bfloat162float(h));
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

int encode(CUtensorMap* map, const void* ptr, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
           const cuuint32_t* box, const char* what) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) { adb_set_error_msg("cuTensorMapEncodeTiled driver entry point unavailable"); return ADB_ERR_CUDA; }
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        static char msg[200];
        snprintf(msg, sizeof(msg), "cuTensorMapEncodeTiled(%s) failed (%d): dims %llu %llu %llu stride0 %llu", what, (int)r,
                 (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2],
                 (unsigned long long)strides[0]);
        adb_set_error_msg(msg);
        return ADB_ERR_INVALID;
    }
    return ADB_OK;
}

// bf16 tensor [batch][rows][K] with element strides (ld, batch_stride); box = 64 (K) x 128 (rows) x 1
int make_map(CUtensorMap* map, const void* ptr, int rows, int K, long long ld, long long bstride, int batch) {
    cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)batch};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)(batch > 1 ? bstride : (long long)rows * ld) * 2};
    cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)BM, 1};
    return encode(map, ptr, 3, dims, strides, box, "matrix");
}

// bf16 NHWC activation [B][H][W][C]; box = 64 channels x Wb x Hb x 1
int make_map_nhwc(CUtensorMap* map, const void* ptr, int B, int H, int W, int C, int Wb, int Hb) {
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)Wb, (cuuint32_t)Hb, 1};
    return encode(map, ptr, 4, dims, strides, box, "nhwc");
}

int launch(const CUtensorMap& mAhi, const CUtensorMap& mAlo, const CUtensorMap& mBhi, const CUtensorMap& mBlo,
           GemmParams& p, long long num_tiles, cudaStream_t stream) {
    p.stages = p.nterms == 3 ? 3 : 6;
    const size_t smem = (size_t)p.stages * (p.nterms == 3 ? 4 : 2) * TILE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
    static AdbDeviceOnce once;
    int num_sms = 0;
    {
        const int rc = once.ensure([]() -> int {
            ADB_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
            ADB_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
            return ADB_OK;
        }, &num_sms);
        if (rc != ADB_OK) return rc;
    }
    const int grid = (int)(num_tiles < num_sms ? num_tiles : num_sms);
    if (p.rope_ndst > 0) gemm_tc_kernel<true><<<grid, NTHREADS, smem, stream>>>(mAhi, mAlo, mBhi, mBlo, p);
    else gemm_tc_kernel<false><<<grid, NTHREADS, smem, stream>>>(mAhi, mAlo, mBhi, mBlo, p);
    ADB_CHECK_LAUNCH("gemm_tc_kernel");
    return ADB_OK;
}

}  // namespace

// D[z] (fp32, optional) / Dhi,Dlo[z] (bf16 split, optional) = act(alpha * A[z] * B[z]^T + bias) + residual[z]
//   A: bf16 [batch][M][K] (ld = lda elements, batch stride sA), B: bf16 [batch][N][K] (ldb, sB); both K-contiguous,
//   16-byte aligned, lda/ldb multiples of 8.  A_lo and B_lo both non-NULL selects the 3-term bf16x3 product.
//   act: 0 none, 1 GELU (erf), 2 ReLU.  Output batch addressing: z = z_outer * zdiv + z_inner and the
//   output/residual offset is z_outer * s?2 + z_inner * s? (zdiv <= 0: plain z * s?), which lets a [B*h] attention
//   batch write straight into a [B, N, h*64] activation.
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
    ADB_REQUIRE(act >= 0 && act <= 2, "adb_gemm_bf16: unknown activation");
    ADB_REQUIRE(sB != 0 || batch == 1, "adb_gemm_bf16: weight broadcast over batch > 1: fold the batch into M instead");
    GemmParams p{};
    p.nterms = A_lo ? 3 : 1;
    CUtensorMap mAhi, mAlo, mBhi, mBlo;
    int rc;
    if ((rc = make_map(&mAhi, A_hi, M, K, lda, sA, batch))) return rc;
    if ((rc = make_map(&mBhi, B_hi, N, K, ldb, sB, batch))) return rc;
    mAlo = mAhi; mBlo = mBhi;
    if (p.nterms == 3) {
        if ((rc = make_map(&mAlo, A_lo, M, K, lda, sA, batch))) return rc;
        if ((rc = make_map(&mBlo, B_lo, N, K, ldb, sB, batch))) return rc;
    }
    p.M = M; p.N = N; p.K = K; p.batch = batch;
    p.D = D; p.ldd = ldd; p.sD = sD;
    p.Dhi = (__nv_bfloat16*)D_hi; p.Dlo = (__nv_bfloat16*)D_lo; p.ldo = ldo; p.sO = sO;
    p.bias = bias; p.residual = residual; p.ldr = ldr; p.sR = sR;
    p.alpha = alpha; p.act = act; p.split_relu = 0;
    p.zdiv = zdiv > 0 ? zdiv : (1 << 30); p.sD2 = sD2; p.sO2 = sO2; p.sR2 = sR2;
    p.conv_wb = 0;
    const long long tiles = (long long)adb_cdiv(M, BM) * adb_cdiv(N, BN) * batch;
    return launch(mAhi, mAlo, mBhi, mBlo, p, tiles, stream);
}

// Linear layer whose leading output columns are attention heads: y = A W^T + bias with
//   columns [0, n_rope_dst*C)   (C = heads*64)  rotated by RoPE2D (table lookup by the token's (y,x) position), split into
//                               bf16 (hi, lo) and written head-major [B*heads, ntok, 64] to q (section 0) / k (section 1);
//   columns [n_rope_dst*C, N)   written as plain fp32 to D[row * ldd + (col - n_rope_dst*C)]  (the V projection).
// Replaces  qkv = Linear(x); q, k = rope(q), rope(k)  (croco/models/blocks.py:94-103, 150-160) + the head transposes:
// the [rows, 3C] fp32 tensor is never written.  M = B * ntok rows.
ADB_API int adb_gemm_bf16_rope(int M, int N, int K, const void* A_hi, const void* A_lo, long long lda, const void* B_hi,
                               const void* B_lo, long long ldb, const float* bias, int ntok, int heads, int n_rope_dst,
                               const long long* pos, const float* table, int n_pos, void* q_hi, void* q_lo, void* k_hi,
                               void* k_lo, float* D_tail, long long ldd, cudaStream_t stream) {
    ADB_REQUIRE(M >= 1 && N >= 1 && K >= 1 && ntok >= 1 && heads >= 1 && M % ntok == 0, "adb_gemm_bf16_rope: bad sizes");
    ADB_REQUIRE(n_rope_dst == 1 || n_rope_dst == 2, "adb_gemm_bf16_rope: n_rope_dst must be 1 or 2");
    const int C = heads * 64;
    ADB_REQUIRE(N >= n_rope_dst * C, "adb_gemm_bf16_rope: N smaller than the rotated sections");
    ADB_REQUIRE(A_hi && B_hi && pos && table && n_pos >= 1 && q_hi && (n_rope_dst == 1 || k_hi), "adb_gemm_bf16_rope: null pointer");
    ADB_REQUIRE((N == n_rope_dst * C) || D_tail, "adb_gemm_bf16_rope: tail columns need D_tail");
    ADB_REQUIRE((A_lo == nullptr) == (B_lo == nullptr), "adb_gemm_bf16_rope: A_lo and B_lo must be given together");
    ADB_REQUIRE(lda % 8 == 0 && ldb % 8 == 0, "adb_gemm_bf16_rope: strides must be multiples of 8 elements");
    ADB_REQUIRE(((uintptr_t)A_hi % 16 == 0) && ((uintptr_t)B_hi % 16 == 0) && ((uintptr_t)table % 16 == 0) &&
                (((uintptr_t)q_hi | (uintptr_t)q_lo | (uintptr_t)k_hi | (uintptr_t)k_lo) % 32 == 0),
                "adb_gemm_bf16_rope: operands must be 16-byte (outputs 32-byte) aligned");
    GemmParams p{};
    p.nterms = A_lo ? 3 : 1;
    CUtensorMap mAhi, mAlo, mBhi, mBlo;
    int rc;
    if ((rc = make_map(&mAhi, A_hi, M, K, lda, 0, 1))) return rc;
    if ((rc = make_map(&mBhi, B_hi, N, K, ldb, 0, 1))) return rc;
    mAlo = mAhi; mBlo = mBhi;
    if (p.nterms == 3) {
        if ((rc = make_map(&mAlo, A_lo, M, K, lda, 0, 1))) return rc;
        if ((rc = make_map(&mBlo, B_lo, N, K, ldb, 0, 1))) return rc;
    }
    p.M = M; p.N = N; p.K = K; p.batch = 1;
    p.D = D_tail ? D_tail - (long long)n_rope_dst * C : nullptr;   // column c of the GEMM lands at D_tail[c - n_rope_dst*C]
    p.ldd = ldd; p.sD = 0;
    p.Dhi = nullptr; p.Dlo = nullptr; p.ldo = N; p.sO = 0;
    p.bias = bias; p.residual = nullptr; p.ldr = N; p.sR = 0;
    p.alpha = 1.0f; p.act = 0; p.split_relu = 0;
    p.zdiv = 1 << 30; p.sD2 = p.sO2 = p.sR2 = 0;
    p.conv_wb = 0;
    p.rope_ndst = n_rope_dst; p.rope_C = C; p.rope_heads = heads; p.rope_ntok = ntok; p.rope_npos = n_pos;
    p.rope_pos = pos; p.rope_table = table;
    p.rope_hi[0] = (__nv_bfloat16*)q_hi; p.rope_lo[0] = (__nv_bfloat16*)q_lo;
    p.rope_hi[1] = (__nv_bfloat16*)k_hi; p.rope_lo[1] = (__nv_bfloat16*)k_lo;
    const long long tiles = (long long)adb_cdiv(M, BM) * adb_cdiv(N, BN);
    return launch(mAhi, mAlo, mBhi, mBlo, p, tiles, stream);
}

// 3x3 / stride 1 / zero-pad 1 convolution as an implicit GEMM (replaces the nn.Conv2d(k=3) calls of
// croco/models/dpt_block.py:20-77,93-112,368-372):
//   x: bf16 split NHWC [B, H, W, Cin] (Cin % 8 == 0);  w: bf16 split [Cout][9 * Cin_pad], tap-major (ky, kx, ci),
//   Cin_pad = Cin rounded up to 64 with zero fill;  outputs are NHWC [B*H*W, Cout] (fp32 D and/or bf16 split).
//   residual: fp32 NHWC [B*H*W, Cout] or NULL.  act: 0 none, 2 ReLU.  split_relu: split output stores relu(value).
ADB_API int adb_conv3x3_bf16(int B, int H, int W, int Cin, int Cout, const void* x_hi, const void* x_lo,
                             const void* w_hi, const void* w_lo, const float* bias, const float* residual, float* D,
                             void* D_hi, void* D_lo, int act, int split_relu, cudaStream_t stream) {
    ADB_REQUIRE(B >= 1 && H >= 1 && W >= 1 && Cin >= 8 && Cout >= 1, "adb_conv3x3_bf16: bad sizes");
    ADB_REQUIRE(Cin % 8 == 0, "adb_conv3x3_bf16: Cin must be a multiple of 8");
    ADB_REQUIRE(x_hi && w_hi && (D || D_hi), "adb_conv3x3_bf16: null pointer");
    ADB_REQUIRE((x_lo == nullptr) == (w_lo == nullptr), "adb_conv3x3_bf16: x_lo and w_lo must be given together");
    ADB_REQUIRE(act == 0 || act == 2, "adb_conv3x3_bf16: act must be 0 or 2 (ReLU)");
    GemmParams p{};
    p.nterms = x_lo ? 3 : 1;
    int wb = 128;
    while (wb > W && wb > 8) wb >>= 1;          // widest power-of-two row segment that fits the image row
    const int hb = BM / wb;
    const int chunks = (Cin + BK - 1) / BK;
    const int Kpad = 9 * chunks * BK;
    CUtensorMap mAhi, mAlo, mBhi, mBlo;
    int rc;
    if ((rc = make_map_nhwc(&mAhi, x_hi, B, H, W, Cin, wb, hb))) return rc;
    if ((rc = make_map(&mBhi, w_hi, Cout, Kpad, Kpad, 0, 1))) return rc;
    mAlo = mAhi; mBlo = mBhi;
    if (p.nterms == 3) {
        if ((rc = make_map_nhwc(&mAlo, x_lo, B, H, W, Cin, wb, hb))) return rc;
        if ((rc = make_map(&mBlo, w_lo, Cout, Kpad, Kpad, 0, 1))) return rc;
    }
    p.M = H * W; p.N = Cout; p.K = Kpad; p.batch = B;
    p.D = D; p.ldd = Cout; p.sD = (long long)H * W * Cout;
    p.Dhi = (__nv_bfloat16*)D_hi; p.Dlo = (__nv_bfloat16*)D_lo; p.ldo = Cout; p.sO = (long long)H * W * Cout;
    p.bias = bias; p.residual = residual; p.ldr = Cout; p.sR = (long long)H * W * Cout;
    p.alpha = 1.0f; p.act = act; p.split_relu = split_relu;
    p.zdiv = 1 << 30; p.sD2 = p.sO2 = p.sR2 = 0;
    p.conv_wb = wb; p.conv_hb = hb; p.conv_H = H; p.conv_W = W; p.conv_chunks = chunks;
    p.conv_tiles_x = adb_cdiv(W, wb); p.conv_tiles_y = adb_cdiv(H, hb);
    const long long tiles = (long long)p.conv_tiles_x * p.conv_tiles_y * adb_cdiv(Cout, BN) * B;
    return launch(mAhi, mAlo, mBhi, mBlo, p, tiles, stream);
}
