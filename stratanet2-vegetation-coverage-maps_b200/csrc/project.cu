// K6 2D max-projection (Appendix A6/A7).
//
// project_plotwise : one CTA per plot.  Pixel ids with the reference's data-dependent min/max
//   normalisation (/root/reference/model/project_to_2d.py:15-22) evaluated op by op in fp32
//   (sub, add, div, mul -- never contracted, never a reciprocal), per-(pixel, stratum) max through
//   64-bit shared-memory atomicMax on (order-preserving value key << 32 | ~point index), which gives
//   the max AND the first-index arg-max deterministically, then the mean over OCCUPIED pixels.
// project_rasters  : one CTA per plot, fixed affine pixel map + clip (:68-78), float64 NaN-filled
//   rasters with the row flip of :108-110.
#include "sn2_common.cuh"

namespace sn2 {

constexpr int PJ_THREADS = 1024;

__device__ __forceinline__ float block_reduce_sum(float v, float *red, int tid)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(SN2_FULL, v, o);
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    float t = (tid < PJ_THREADS / 32) ? red[tid] : 0.f;
    if (tid < 32) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(SN2_FULL, t, o);
        if (tid == 0) red[0] = t;
    }
    __syncthreads();
    return red[0];
}

__global__ void __launch_bounds__(PJ_THREADS, 1)
project_plotwise_kernel(const float *__restrict__ cloud, const float4 *__restrict__ pred, int N, int F, int D,
                        float *__restrict__ out, int *__restrict__ pix, float *__restrict__ pmax,
                        int *__restrict__ parg)
{
    extern __shared__ unsigned long long keys[];  // [3][(D+1)*(D+1)]
    __shared__ float red[4][32];
    __shared__ float sred[32];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int D1 = D + 1, P = D1 * D1;
    const float *xs = cloud + (size_t)b * F * N;
    const float *ys = xs + N;
    for (int i = tid; i < 3 * P; i += PJ_THREADS) keys[i] = 0ull;

    float mnx = INFINITY, mny = INFINITY, mxx = -INFINITY, mxy = -INFINITY;
    for (int i = tid; i < N; i += PJ_THREADS) {
        float x = __ldg(xs + i), y = __ldg(ys + i);
        mnx = fminf(mnx, x);
        mxx = fmaxf(mxx, x);
        mny = fminf(mny, y);
        mxy = fmaxf(mxy, y);
    }
    mnx = -warp_max(-mnx);
    mny = -warp_max(-mny);
    mxx = warp_max(mxx);
    mxy = warp_max(mxy);
    if (lane == 0) {
        red[0][warp] = mnx;
        red[1][warp] = mny;
        red[2][warp] = mxx;
        red[3][warp] = mxy;
    }
    __syncthreads();
    mnx = -warp_max(-red[0][lane]);
    mny = -warp_max(-red[1][lane]);
    mxx = warp_max(red[2][lane]);
    mxy = warp_max(red[3][lane]);
    // (max - min + 0.0001): fp32 sub then fp32 add
    const float denx = __fadd_rn(__fsub_rn(mxx, mnx), 0.0001f);
    const float deny = __fadd_rn(__fsub_rn(mxy, mny), 0.0001f);
    const float fD = (float)D;

    const float4 *pr = pred + (size_t)b * N;
    for (int i = tid; i < N; i += PJ_THREADS) {
        const float x = __ldg(xs + i), y = __ldg(ys + i);
        int px = (int)floorf(__fmul_rn(__fdiv_rn(__fsub_rn(x, mnx), denx), fD));
        int py = (int)floorf(__fmul_rn(__fdiv_rn(__fsub_rn(y, mny), deny), fD));
        px = min(max(px, 0), D);
        py = min(max(py, 0), D);
        const int p = px * D1 + py;
        if (pix) pix[(size_t)b * N + i] = p;  // px * (D + 1) + py
        const float4 v = __ldg(pr + i);
        const unsigned long long lo = (unsigned long long)(0xffffffffu - (unsigned)i);
        atomicMax(&keys[p], ((unsigned long long)fkey(v.x) << 32) | lo);
        atomicMax(&keys[P + p], ((unsigned long long)fkey(v.z) << 32) | lo);
        atomicMax(&keys[2 * P + p], ((unsigned long long)fkey(v.w) << 32) | lo);
    }
    __syncthreads();

    float s_low = 0.f, s_bare = 0.f, s_med = 0.f, s_high = 0.f, cnt = 0.f;
    for (int p = tid; p < P; p += PJ_THREADS) {
        const unsigned long long k0 = keys[p], k1 = keys[P + p], k2 = keys[2 * P + p];
        const bool occ = k0 != 0ull;
        const float v0 = occ ? fkey_inv((unsigned)(k0 >> 32)) : 0.f;
        const float v1 = occ ? fkey_inv((unsigned)(k1 >> 32)) : 0.f;
        const float v2 = occ ? fkey_inv((unsigned)(k2 >> 32)) : 0.f;
        if (occ) {
            s_low += v0;
            s_bare += __fsub_rn(1.0f, v0);
            s_med += v1;
            s_high += v2;
            cnt += 1.f;
        }
        // aux arrays use the kernel's own (D+1) x (D+1) pixel frame (index = px*(D+1)+py, as `pix`): a point that
        // lands in row / column D is counted by the forward mean, so the backward must see it too
        const size_t o = (size_t)b * 3 * P + p;
        if (pmax) {
            pmax[o] = v0;
            pmax[o + P] = v1;
            pmax[o + 2 * (size_t)P] = v2;
        }
        if (parg) {
            parg[o] = occ ? b * N + (int)(0xffffffffu - (unsigned)(k0 & 0xffffffffull)) : -1;
            parg[o + P] = occ ? b * N + (int)(0xffffffffu - (unsigned)(k1 & 0xffffffffull)) : -1;
            parg[o + 2 * (size_t)P] = occ ? b * N + (int)(0xffffffffu - (unsigned)(k2 & 0xffffffffull)) : -1;
        }
    }
    s_low = block_reduce_sum(s_low, sred, tid);
    s_bare = block_reduce_sum(s_bare, sred, tid);
    s_med = block_reduce_sum(s_med, sred, tid);
    s_high = block_reduce_sum(s_high, sred, tid);
    cnt = block_reduce_sum(cnt, sred, tid);
    if (tid == 0) {
        const float c = fmaxf(cnt, 1.f);
        out[b * 4 + 0] = s_low / c;
        out[b * 4 + 1] = s_bare / c;
        out[b * 4 + 2] = s_med / c;
        out[b * 4 + 3] = s_high / c;
    }
}

constexpr int RS_THREADS = 512;
__global__ void __launch_bounds__(RS_THREADS)
project_rasters_kernel(const float *__restrict__ cloud, const float *__restrict__ cov, long long cov_sb,
                       long long cov_sn, long long cov_sc, int N, int F, int D, float scale, float shift,
                       double *__restrict__ rasters, int *__restrict__ pix)
{
    extern __shared__ unsigned rkeys[];  // [3][D*D]
    const int b = blockIdx.x, tid = threadIdx.x;
    const int P = D * D;
    for (int i = tid; i < 3 * P; i += RS_THREADS) rkeys[i] = 0u;
    __syncthreads();
    const float *xs = cloud + (size_t)b * F * N;
    const float *ys = xs + N;
    const float *cv = cov + b * cov_sb;
    for (int i = tid; i < N; i += RS_THREADS) {
        const float x = __ldg(xs + i), y = __ldg(ys + i);
        // floor((xy + 0.0001) * scale + shift), then clip to [0, D-1]
        int px = (int)floorf(__fadd_rn(__fmul_rn(__fadd_rn(x, 0.0001f), scale), shift));
        int py = (int)floorf(__fadd_rn(__fmul_rn(__fadd_rn(y, 0.0001f), scale), shift));
        px = min(max(px, 0), D - 1);
        py = min(max(py, 0), D - 1);
        const int p = py * D + px;  // image[m = y_pix, k = x_pix]
        if (pix) pix[(size_t)b * N + i] = p;
        const float *c = cv + i * cov_sn;
        atomicMax(&rkeys[p], fkey(__ldg(c)));
        atomicMax(&rkeys[P + p], fkey(__ldg(c + 2 * cov_sc)));
        atomicMax(&rkeys[2 * P + p], fkey(__ldg(c + 3 * cov_sc)));
    }
    __syncthreads();
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    for (int i = tid; i < 3 * P; i += RS_THREADS) {
        const int band = i / P, p = i - band * P;
        const int y = p / D, x = p - y * D;
        const unsigned k = rkeys[i];
        rasters[(size_t)b * 3 * P + (size_t)band * P + (size_t)(D - 1 - y) * D + x] = k ? (double)fkey_inv(k) : nan;
    }
}

}  // namespace sn2

extern "C" int sn2_project_plotwise(const float *cloud, const float *pred, int B, int N, int F, int D, float *out,
                                    int *pix, float *pmax, int *parg, void *stream)
{
    if (!cloud || !pred || !out || B <= 0 || N <= 0 || F < 2 || D <= 0) return SN2_EINVAL;
    size_t smem = (size_t)3 * (D + 1) * (D + 1) * sizeof(unsigned long long);
    if (smem > 200 * 1024) return SN2_EUNSUPPORTED;
    auto kern = sn2::project_plotwise_kernel;
    SN2_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "plotwise attr");
    kern<<<B, sn2::PJ_THREADS, smem, (cudaStream_t)stream>>>(cloud, reinterpret_cast<const float4 *>(pred), N, F, D, out,
                                                             pix, pmax, parg);
    SN2_LAUNCH_CHECK("project_plotwise_kernel");
    return SN2_OK;
}

extern "C" int sn2_project_rasters(const float *cloud, const float *cov, long long cov_sb, long long cov_sn,
                                   long long cov_sc, int B, int N, int F, int D, float scale, float shift,
                                   double *rasters, int *pix, void *stream)
{
    if (!cloud || !cov || !rasters || B <= 0 || N <= 0 || F < 2 || D <= 0) return SN2_EINVAL;
    size_t smem = (size_t)3 * D * D * sizeof(unsigned);
    if (smem > 200 * 1024) return SN2_EUNSUPPORTED;
    auto kern = sn2::project_rasters_kernel;
    SN2_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "rasters attr");
    kern<<<B, sn2::RS_THREADS, smem, (cudaStream_t)stream>>>(cloud, cov, cov_sb, cov_sn, cov_sc, N, F, D, scale, shift,
                                                             rasters, pix);
    SN2_LAUNCH_CHECK("project_rasters_kernel");
    return SN2_OK;
}
