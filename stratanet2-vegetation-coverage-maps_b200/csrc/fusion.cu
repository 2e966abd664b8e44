// Local-map fusion (BASELINE config 4): weighted-average mosaic of per-plot rasters
// into the parcel grid.  Replaces the per-plot GeoTIFF round trip + rasterio.merge with the custom
// `_weighted_average_of_rasters` rule of /root/reference/inference/geotiff_raster.py:103-118, 199-235, 294-347:
// weight w = 1.5 - r (r = normalised distance to the plot centre, NaN beyond r > 0.5), score = sum(s*w) / sum(w)
// over the plots that have a value at the pixel, weight band = sum of the weights of all plots covering it.
// The accumulation is order independent (fp64 atomics); the reference's file-order-dependent pairwise update
// gives the same result whenever every in-disk pixel of a plot has a value.
#include "sn2_common.cuh"

namespace sn2 {

__global__ void __launch_bounds__(256)
fuse_accumulate_kernel(const double *__restrict__ rasters, const int *__restrict__ offsets, int P, int D, int H, int W,
                       double *__restrict__ num, double *__restrict__ den, double *__restrict__ wsum)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int DD = D * D;
    if (t >= (long long)P * DD) return;
    const int pl = (int)(t / DD), px = (int)(t - (long long)pl * DD);
    const int i = px / D, j = px - i * D;
    // normalised meshgrid of data_loader/loader.py:108-124: (arange(-D//2, D//2) + 0.5) / D
    const double xx = ((double)(j - D / 2) + 0.5) / D, yy = ((double)(i - D / 2) + 0.5) / D;
    const double r = sqrt(xx * xx + yy * yy);
    if (r > 0.5) return;  // weight NaN outside the disk: the plot does not cover this pixel
    const double w = 1.5 - r;
    const int R = offsets[2 * pl] + i, C = offsets[2 * pl + 1] + j;
    if (R < 0 || R >= H || C < 0 || C >= W) return;
    const size_t g = (size_t)R * W + C;
    atomicAdd(wsum + g, w);
#pragma unroll
    for (int b = 0; b < 3; ++b) {
        const double s = rasters[((size_t)pl * 3 + b) * DD + px];
        if (s == s) {  // not NaN
            atomicAdd(num + (size_t)b * H * W + g, s * w);
            atomicAdd(den + (size_t)b * H * W + g, w);
        }
    }
}

__global__ void __launch_bounds__(256)
fuse_finalize_kernel(const double *__restrict__ num, const double *__restrict__ den, const double *__restrict__ wsum,
                     long long HW, double *__restrict__ out)
{
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= HW) return;
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
#pragma unroll
    for (int b = 0; b < 3; ++b) {
        const double d = den[(size_t)b * HW + g];
        out[(size_t)b * HW + g] = d > 0.0 ? num[(size_t)b * HW + g] / d : nan;
    }
    const double w = wsum[g];
    out[(size_t)3 * HW + g] = w > 0.0 ? w : nan;
}

}  // namespace sn2

extern "C" int sn2_fuse_accumulate(const double *rasters, const int *offsets, int P, int D, int H, int W, double *num,
                                   double *den, double *wsum, void *stream)
{
    if (!rasters || !offsets || !num || !den || !wsum || P <= 0 || D <= 0 || H <= 0 || W <= 0) return SN2_EINVAL;
    const long long total = (long long)P * D * D;
    sn2::fuse_accumulate_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rasters, offsets, P, D, H, W,
                                                                                                 num, den, wsum);
    SN2_LAUNCH_CHECK("fuse_accumulate_kernel");
    return SN2_OK;
}

extern "C" int sn2_fuse_finalize(const double *num, const double *den, const double *wsum, int H, int W, double *out,
                                 void *stream)
{
    if (!num || !den || !wsum || !out || H <= 0 || W <= 0) return SN2_EINVAL;
    const long long HW = (long long)H * W;
    sn2::fuse_finalize_kernel<<<(unsigned)((HW + 255) / 256), 256, 0, (cudaStream_t)stream>>>(num, den, wsum, HW, out);
    SN2_LAUNCH_CHECK("fuse_finalize_kernel");
    return SN2_OK;
}
