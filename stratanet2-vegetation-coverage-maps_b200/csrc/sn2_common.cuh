// Shared device helpers for libsn2_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/sn2.h"

#define SN2_FULL 0xffffffffu

namespace sn2 {

// d2 = ((dx*dx + dy*dy) + dz*dz), every operation rounded separately:
// the _rn intrinsics are never contracted into FMA by nvcc.
__device__ __forceinline__ float dist2(float ax, float ay, float az, float bx, float by, float bz)
{
    float dx = __fsub_rn(ax, bx);
    float dy = __fsub_rn(ay, by);
    float dz = __fsub_rn(az, bz);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// order-preserving float -> uint32 key (monotone for all non-NaN floats)
__device__ __forceinline__ unsigned fkey(float f)
{
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float fkey_inv(unsigned k)
{
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

__device__ __forceinline__ float warp_max(float v)
{
    return fkey_inv(__reduce_max_sync(SN2_FULL, fkey(v)));
}

__device__ __forceinline__ float4 ldg4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }

}  // namespace sn2

// host-side helpers -------------------------------------------------------------------------
void sn2_set_cuda_error(cudaError_t e, const char *where);
#define SN2_LAUNCH_CHECK(where)                      \
    do {                                             \
        cudaError_t e__ = cudaGetLastError();        \
        if (e__ != cudaSuccess) {                    \
            sn2_set_cuda_error(e__, where);          \
            return SN2_ECUDA;                        \
        }                                            \
    } while (0)
#define SN2_CUDA_TRY(call, where)                    \
    do {                                             \
        cudaError_t e__ = (call);                    \
        if (e__ != cudaSuccess) {                    \
            sn2_set_cuda_error(e__, where);          \
            return SN2_ECUDA;                        \
        }                                            \
    } while (0)
