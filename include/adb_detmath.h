/* Deterministic elementary functions shared by the CUDA projection kernel and the C oracle.
 *
 * The tile-key contract is bit-exact, and the tile bounds depend on
 * ceil(extent * sqrt(.)) with extent = sqrt(2 ln(255 * opacity)).  libm's logf and CUDA's logf differ in
 * the last ulp, which would flip a ceil() roughly once per million Gaussians.  This header defines
 * ln() from IEEE-exact operations only (bit masks, +, -, *, /), so any compiler that does not contract
 * or reassociate floating point (gcc -ffp-contract=off, nvcc -fmad=false) returns identical bits.
 * Max relative error ~2e-7 on [1e-3, 1e3].
 */
#ifndef ADB_DETMATH_H
#define ADB_DETMATH_H
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define ADB_HD __host__ __device__ __forceinline__
#else
#define ADB_HD static inline
#endif

/* x must be a positive normal float. */
ADB_HD float adb_det_logf(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    int e = (int)(u >> 23) - 127;
    u = (u & 0x007fffffu) | 0x3f800000u;
    float m;
    memcpy(&m, &u, 4);
    if (m > 1.41421356f) { m = m * 0.5f; e += 1; }
    float f = m - 1.0f;
    float s = f / (2.0f + f);
    float z = s * s;
    float p = 0.0909090909f;          /* 1/11 */
    p = p * z + 0.111111111f;         /* 1/9  */
    p = p * z + 0.142857143f;         /* 1/7  */
    p = p * z + 0.2f;                 /* 1/5  */
    p = p * z + 0.333333333f;         /* 1/3  */
    p = p * z + 1.0f;
    float r = (2.0f * s) * p;
    return r + (float)e * 0.693147181f;
}

#endif
