// Device-side loader transforms: the train-time half of load_cloud that follows centering and the
// fake ground points -- augment (data_loader/loader.py:161-214) and rescale_cloud (:135-158) -- for a whole batch in one
// launch instead of numpy per plot in the DataLoader's main process.
//
// The random draws are INPUTS (one angle and two flip flags per plot, standard-normal float64 draws per point for the
// xy and colour noise), so the kernel is a pure function that can be checked against the numpy restatement draw for
// draw; the caller produces them with torch's device generator.  dtypes follow the reference: the rotation is a
// float64 product cast back to float32 (np.dot of a float32 array with a float64 matrix), the noise is
// clip(sigma * randn, +-clip) in float64 cast to float32 and added in float32; the colour noise uses the xy sigma
// (0.1), as the reference does (:199-207 read `sigma`, not `sigm`).
#include "sn2_common.cuh"

namespace sn2 {

__global__ void __launch_bounds__(256)
augment_rescale_kernel(const float *__restrict__ in, int B, int N, const double *__restrict__ angle, const unsigned char *__restrict__ flip,
                       const double *__restrict__ noise, float z_max, float *__restrict__ xyz, float *__restrict__ cloud)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)B * N) return;
    const int b = (int)(t / N), n = (int)(t - (long long)b * N);
    const float *src = in + (size_t)b * 10 * N + n;
    float v[10];
#pragma unroll
    for (int r = 0; r < 10; ++r) v[r] = __ldg(src + (size_t)r * N);
    float x = v[0], y = v[1];
    float px = x, py = y;  // positions (xyz) get the rotation and the flips, not the noise
    if (angle) {
        const double c = cos(angle[b]), s = sin(angle[b]);
        // [x y] . [[c, -s], [s, c]]  (rotate_around_z, :210-214), float64 then cast
        px = (float)((double)x * c + (double)y * s);
        py = (float)((double)x * -s + (double)y * c);
        if (flip[2 * b]) px = -px;
        if (flip[2 * b + 1]) py = -py;
        x = px;
        y = py;
    }
    if (noise) {
        const double *nz = noise + (size_t)b * 6 * N + n;
        const double sigma = 0.01 * 10, clip_xy = 0.03 * 10, clip_c = 0.03 * 65536;
        x = __fadd_rn(x, (float)fmin(fmax(sigma * nz[0], -clip_xy), clip_xy));
        y = __fadd_rn(y, (float)fmin(fmax(sigma * nz[(size_t)N], -clip_xy), clip_xy));
#pragma unroll
        for (int q = 0; q < 4; ++q) v[3 + q] = __fadd_rn(v[3 + q], (float)fmin(fmax(sigma * nz[(size_t)(2 + q) * N], -clip_c), clip_c));
    }
    float *ox = xyz + (size_t)b * 3 * N + n, *oc = cloud + (size_t)b * 10 * N + n;
    ox[0] = px; ox[(size_t)N] = py; ox[2 * (size_t)N] = v[2];
    oc[0] = __fdiv_rn(x, 10.f);
    oc[(size_t)N] = __fdiv_rn(y, 10.f);
    oc[2 * (size_t)N] = __fdiv_rn(v[2], z_max);
#pragma unroll
    for (int q = 0; q < 4; ++q) oc[(size_t)(3 + q) * N] = __fdiv_rn(v[3 + q], 65536.f);
    oc[7 * (size_t)N] = __fdiv_rn(v[7], 32768.f);
    oc[8 * (size_t)N] = __fdiv_rn(__fsub_rn(v[8], 1.f), 6.f);
    oc[9 * (size_t)N] = __fdiv_rn(__fsub_rn(v[9], 1.f), 6.f);
}

}  // namespace sn2

extern "C" int sn2_augment_rescale(const float *in, int B, int N, const double *angle, const unsigned char *flip, const double *noise,
                                   float z_max, float *xyz, float *cloud, void *stream)
{
    if (!in || !xyz || !cloud || B <= 0 || N <= 0 || (angle && !flip)) return SN2_EINVAL;
    const long long total = (long long)B * N;
    sn2::augment_rescale_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(in, B, N, angle, flip, noise, z_max,
                                                                                                 xyz, cloud);
    SN2_LAUNCH_CHECK("augment_rescale_kernel");
    return SN2_OK;
}
