// Shared device/host helpers for the artdeco_b200 C-ABI library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define ADB_OK 0
#define ADB_ERR_INVALID 1   // bad argument (null pointer, negative size, unsupported shape)
#define ADB_ERR_CUDA 2      // a CUDA runtime call or launch failed
#define ADB_ERR_WORKSPACE 3 // caller-provided workspace too small

#define ADB_API extern "C" __attribute__((visibility("default")))

// Records the last CUDA error string for adb_last_error().
void adb_set_error(const char* where, cudaError_t e);
void adb_set_error_msg(const char* msg);

#define ADB_CHECK_LAUNCH(where)                                  \
    do {                                                         \
        cudaError_t _e = cudaGetLastError();                     \
        if (_e != cudaSuccess) { adb_set_error(where, _e); return ADB_ERR_CUDA; } \
    } while (0)

#define ADB_CUDA(call)                                           \
    do {                                                         \
        cudaError_t _e = (call);                                 \
        if (_e != cudaSuccess) { adb_set_error(#call, _e); return ADB_ERR_CUDA; } \
    } while (0)

#define ADB_REQUIRE(cond, msg)                                   \
    do { if (!(cond)) { adb_set_error_msg(msg); return ADB_ERR_INVALID; } } while (0)

static inline int adb_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

__device__ __forceinline__ float adb_warp_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 16);
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}

// Streaming (read-once) 128-bit load that does not pollute L1.
__device__ __forceinline__ float4 adb_ldg_stream4(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
