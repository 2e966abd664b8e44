// Train-mode head and point-wise losses (§8f rank 3).
//
//   head_fwd / head_bwd   relu(lin1) -> lin2 -> softmax(4) x sigmoid(1)   (reference model/point_net2.py:141-153) in one
//                         kernel each way.  torch ran it as addmm, relu, addmm, slice, softmax, sigmoid, mul and the
//                         mirrored backward nodes (two cuBLAS SIMT GEMMs each way over B*N rows); here the forward
//                         reads the 34-wide FP1 row once and writes coverages + probabilities, and the backward
//                         RECOMPUTES the head from the same row (nothing of the head is saved), writes the gradient of
//                         the row and accumulates dW1, db1, dW2, db2 in registers (deterministic two-stage reduction).
//   pointwise_loss        the two point-wise terms of the reference loss (learning/loss_functions.py:19-57): negative
//                         log-likelihood of the strata probabilities under the KDE pdf of z (fp64, as the reference
//                         computes it) and the binary entropy of the medium / high columns; one pass forward, one
//                         backward that writes d(loss)/d(proba) directly.
//   kde_lut               the pdf itself: scipy interp1d(kind="linear") over the KDE grid (learning/kde_mixture.py:64-75)
//                         as a device look-up, replacing numpy interpolation on the CPU + an H2D copy every step.
#include "sn2_common.cuh"

namespace sn2 {

constexpr int HD_CI = 34, HD_CH = 16, HD_CO = 5;
constexpr int HD_NP = HD_CH * (HD_CI + 1) + HD_CO * (HD_CH + 1);  // 560 + 85 floats of one partial [dW1|db1|dW2|db2]

struct HeadOut {
    float hp[HD_CH];  // lin1 pre-activation
    float proba[4];
    float dens;
};

// x: the row (BatchNorm of the producing block already applied); W1t [34][16] k-major, W2 [5][16] in shared memory
__device__ __forceinline__ void head_row(const float (&x)[HD_CI], const float *W1t, const float *b1, const float *W2,
                                         const float *b2, HeadOut &o)
{
#pragma unroll
    for (int j = 0; j < HD_CH; ++j) o.hp[j] = b1[j];
#pragma unroll
    for (int k = 0; k < HD_CI; ++k) {
#pragma unroll
        for (int j4 = 0; j4 < HD_CH / 4; ++j4) {
            const float4 w = *reinterpret_cast<const float4 *>(W1t + k * HD_CH + 4 * j4);
            o.hp[4 * j4] = fmaf(x[k], w.x, o.hp[4 * j4]);
            o.hp[4 * j4 + 1] = fmaf(x[k], w.y, o.hp[4 * j4 + 1]);
            o.hp[4 * j4 + 2] = fmaf(x[k], w.z, o.hp[4 * j4 + 2]);
            o.hp[4 * j4 + 3] = fmaf(x[k], w.w, o.hp[4 * j4 + 3]);
        }
    }
    float s[HD_CO];
#pragma unroll
    for (int c = 0; c < HD_CO; ++c) {
        float a = b2[c];
#pragma unroll
        for (int j = 0; j < HD_CH; ++j) a = fmaf(fmaxf(o.hp[j], 0.f), W2[c * HD_CH + j], a);
        s[c] = a;
    }
    const float mx = fmaxf(fmaxf(s[0], s[1]), fmaxf(s[2], s[3]));
    float e[4], sum = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        e[c] = expf(s[c] - mx);
        sum += e[c];
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) o.proba[c] = e[c] / sum;
    o.dens = 1.f / (1.f + expf(-s[4]));
}

__device__ __forceinline__ void head_load_weights(const float *W1, const float *b1, const float *W2, const float *b2,
                                                  const float *in_ss, float *W1t, float *b1S, float *W2S, float *b2S, float *inS)
{
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int e = tid; e < HD_CH * HD_CI; e += nt) {
        const int j = e / HD_CI, k = e - j * HD_CI;
        W1t[k * HD_CH + j] = __ldg(W1 + e);
    }
    for (int e = tid; e < HD_CO * HD_CH; e += nt) W2S[e] = __ldg(W2 + e);
    for (int e = tid; e < HD_CH; e += nt) b1S[e] = __ldg(b1 + e);
    for (int e = tid; e < HD_CO; e += nt) b2S[e] = __ldg(b2 + e);
    for (int e = tid; e < HD_CI; e += nt) {
        inS[e] = in_ss ? __ldg(in_ss + e) : 1.f;
        inS[HD_CI + e] = in_ss ? __ldg(in_ss + HD_CI + e) : 0.f;
    }
}

__device__ __forceinline__ void head_load_row(const float *f1, long long row, const float *inS, float (&x)[HD_CI])
{
    const float2 *p = reinterpret_cast<const float2 *>(f1 + row * HD_CI);  // 136-byte rows: 8-byte aligned
#pragma unroll
    for (int i = 0; i < HD_CI / 2; ++i) {
        const float2 t = __ldg(p + i);
        x[2 * i] = fmaf(t.x, inS[2 * i], inS[HD_CI + 2 * i]);
        x[2 * i + 1] = fmaf(t.y, inS[2 * i + 1], inS[HD_CI + 2 * i + 1]);
    }
}

__global__ void __launch_bounds__(256)
head_fwd_kernel(const float *__restrict__ f1, const float *__restrict__ in_ss, const float *__restrict__ W1,
                const float *__restrict__ b1, const float *__restrict__ W2, const float *__restrict__ b2, long long R,
                float4 *__restrict__ cov, float4 *__restrict__ proba)
{
    __shared__ __align__(16) float W1t[HD_CI * HD_CH];
    __shared__ float b1S[HD_CH], W2S[HD_CO * HD_CH], b2S[HD_CO], inS[2 * HD_CI];
    head_load_weights(W1, b1, W2, b2, in_ss, W1t, b1S, W2S, b2S, inS);
    __syncthreads();
    for (long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x; row < R; row += (long long)gridDim.x * blockDim.x) {
        float x[HD_CI];
        head_load_row(f1, row, inS, x);
        HeadOut o;
        head_row(x, W1t, b1S, W2S, b2S, o);
        proba[row] = make_float4(o.proba[0], o.proba[1], o.proba[2], o.proba[3]);
        cov[row] = make_float4(o.proba[0] * o.dens, o.proba[1] * o.dens, o.proba[2] * o.dens, o.proba[3] * o.dens);
    }
}

// Backward.  128-row tiles, thread = row in phase 1 (recompute, ds, dh, row gradient; rows staged in shared memory),
// thread = (row group, input column) in phase 2 (weight gradients in registers over all tiles of the CTA).
constexpr int HB_T = 128;
constexpr int HB_XS = HD_CI + 1, HB_HS = HD_CH + 1;  // odd row strides: conflict-free for thread = row
constexpr int HB_G1 = HB_T / (HD_CI + 1), HB_G2 = HB_T / (HD_CH + 1);

__global__ void __launch_bounds__(HB_T)
head_bwd_kernel(const float *__restrict__ f1, const float *__restrict__ in_ss, const float *__restrict__ W1,
                const float *__restrict__ b1, const float *__restrict__ W2, const float *__restrict__ b2,
                const float4 *__restrict__ dcov, const float4 *__restrict__ dproba, long long R, float *__restrict__ df1,
                float *__restrict__ partial)
{
    __shared__ __align__(16) float W1t[HD_CI * HD_CH];
    __shared__ float b1S[HD_CH], W2S[HD_CO * HD_CH], b2S[HD_CO], inS[2 * HD_CI];
    __shared__ float W1S[HD_CH * HD_CI];              // [j][k] for the row gradient
    __shared__ float xS[HB_T * HB_XS], dhS[HB_T * HB_HS], hS[HB_T * HB_HS], dsS[HB_T * HD_CO];
    const int tid = threadIdx.x;
    head_load_weights(W1, b1, W2, b2, in_ss, W1t, b1S, W2S, b2S, inS);
    for (int e = tid; e < HD_CH * HD_CI; e += HB_T) W1S[e] = __ldg(W1 + e);
    float acc1[HD_CH], acc2[HD_CO];
#pragma unroll
    for (int j = 0; j < HD_CH; ++j) acc1[j] = 0.f;
#pragma unroll
    for (int c = 0; c < HD_CO; ++c) acc2[c] = 0.f;
    const int g1 = tid / (HD_CI + 1), i1 = tid - g1 * (HD_CI + 1);
    const int g2 = tid / (HD_CH + 1), j2 = tid - g2 * (HD_CH + 1);
    __syncthreads();
    const long long ntiles = (R + HB_T - 1) / HB_T;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long row = tile * HB_T + tid;
        const int rows = (int)min((long long)HB_T, R - tile * HB_T);
        if (tid < rows) {
            float x[HD_CI];
            head_load_row(f1, row, inS, x);
            HeadOut o;
            head_row(x, W1t, b1S, W2S, b2S, o);
            const float4 gc = dcov ? __ldg(dcov + row) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 gp = dproba ? __ldg(dproba + row) : make_float4(0.f, 0.f, 0.f, 0.f);
            // cov_c = proba_c * dens
            const float gpt[4] = {fmaf(gc.x, o.dens, gp.x), fmaf(gc.y, o.dens, gp.y), fmaf(gc.z, o.dens, gp.z), fmaf(gc.w, o.dens, gp.w)};
            const float gd = gc.x * o.proba[0] + gc.y * o.proba[1] + gc.z * o.proba[2] + gc.w * o.proba[3];
            const float dot = gpt[0] * o.proba[0] + gpt[1] * o.proba[1] + gpt[2] * o.proba[2] + gpt[3] * o.proba[3];
            float ds[HD_CO];
#pragma unroll
            for (int c = 0; c < 4; ++c) ds[c] = o.proba[c] * (gpt[c] - dot);
            ds[4] = gd * o.dens * (1.f - o.dens);
#pragma unroll
            for (int c = 0; c < HD_CO; ++c) dsS[tid * HD_CO + c] = ds[c];
            float dhp[HD_CH];
#pragma unroll
            for (int j = 0; j < HD_CH; ++j) {
                float a = 0.f;
#pragma unroll
                for (int c = 0; c < HD_CO; ++c) a = fmaf(ds[c], W2S[c * HD_CH + j], a);
                dhp[j] = o.hp[j] > 0.f ? a : 0.f;
                dhS[tid * HB_HS + j] = dhp[j];
                hS[tid * HB_HS + j] = fmaxf(o.hp[j], 0.f);
            }
#pragma unroll
            for (int k = 0; k < HD_CI; ++k) xS[tid * HB_XS + k] = x[k];
            if (df1) {
                float2 *out = reinterpret_cast<float2 *>(df1 + row * HD_CI);
#pragma unroll
                for (int k2 = 0; k2 < HD_CI / 2; ++k2) {
                    float a = 0.f, b = 0.f;
#pragma unroll
                    for (int j = 0; j < HD_CH; ++j) {
                        a = fmaf(dhp[j], W1S[j * HD_CI + 2 * k2], a);
                        b = fmaf(dhp[j], W1S[j * HD_CI + 2 * k2 + 1], b);
                    }
                    out[k2] = make_float2(a, b);
                }
            }
        }
        __syncthreads();
        if (g1 < HB_G1) {
            for (int r = g1; r < rows; r += HB_G1) {
                const float xv = i1 < HD_CI ? xS[r * HB_XS + i1] : 1.f;
#pragma unroll
                for (int j = 0; j < HD_CH; ++j) acc1[j] = fmaf(dhS[r * HB_HS + j], xv, acc1[j]);
            }
        }
        if (g2 < HB_G2) {
            for (int r = g2; r < rows; r += HB_G2) {
                const float hv = j2 < HD_CH ? hS[r * HB_HS + j2] : 1.f;
#pragma unroll
                for (int c = 0; c < HD_CO; ++c) acc2[c] = fmaf(dsS[r * HD_CO + c], hv, acc2[c]);
            }
        }
        __syncthreads();
    }
    // group partials -> one partial per CTA (xS is free now: HB_T * 35 floats >= G1 * 560 / G2 * 85)
    float *red = xS;
    static_assert(HB_G1 * HD_CH * (HD_CI + 1) <= HB_T * HB_XS && HB_G2 * HD_CO * (HD_CH + 1) <= HB_T * HB_HS, "reduction scratch");
    if (g1 < HB_G1) {
#pragma unroll
        for (int j = 0; j < HD_CH; ++j) red[(g1 * HD_CH + j) * (HD_CI + 1) + i1] = acc1[j];
    }
    if (g2 < HB_G2) {
#pragma unroll
        for (int c = 0; c < HD_CO; ++c) dhS[(g2 * HD_CO + c) * (HD_CH + 1) + j2] = acc2[c];
    }
    __syncthreads();
    float *out = partial + (size_t)blockIdx.x * HD_NP;
    for (int t = tid; t < HD_CH * (HD_CI + 1); t += HB_T) {
        float s = 0.f;
        for (int g = 0; g < HB_G1; ++g) s += red[g * HD_CH * (HD_CI + 1) + t];
        out[t] = s;
    }
    for (int t = tid; t < HD_CO * (HD_CH + 1); t += HB_T) {
        float s = 0.f;
        for (int g = 0; g < HB_G2; ++g) s += dhS[g * HD_CO * (HD_CH + 1) + t];
        out[HD_CH * (HD_CI + 1) + t] = s;
    }
}

// fixed-order reduction of the CTA partials: one warp per element
__global__ void __launch_bounds__(256)
head_wgrad_reduce_kernel(const float *__restrict__ partial, int nblk, float *__restrict__ dW1, float *__restrict__ db1,
                         float *__restrict__ dW2, float *__restrict__ db2)
{
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (t >= HD_NP) return;
    float s = 0.f;
    for (int b = lane; b < nblk; b += 32) s += __ldg(partial + (size_t)b * HD_NP + t);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(SN2_FULL, s, d);
    if (lane) return;
    if (t < HD_CH * (HD_CI + 1)) {
        const int j = t / (HD_CI + 1), i = t - j * (HD_CI + 1);
        if (i < HD_CI) dW1[j * HD_CI + i] = s;
        else db1[j] = s;
    } else {
        const int u = t - HD_CH * (HD_CI + 1), c = u / (HD_CH + 1), j = u - c * (HD_CH + 1);
        if (j < HD_CH) dW2[c * HD_CH + j] = s;
        else db2[c] = s;
    }
}

// ---- point-wise losses --------------------------------------------------------------------------------------
// nll_i = -log((p0 + p1) * pdf0 + p2 * pdf1 + p3 * pdf2)   in fp64 (pdf is float64 in the reference, :42, :53-57)
// ent_i = -sum_{c in 2,3} [p_c log(p_c + EPS) + (1 - p_c) log(1 - p_c + EPS)]   in fp32 (:19-24), EPS = 1e-4
constexpr float LOSS_EPS = 0.0001f;

__global__ void __launch_bounds__(256)
pointwise_loss_fwd_kernel(const float4 *__restrict__ proba, const double *__restrict__ pdf, long long R, double *__restrict__ sums)
{
    __shared__ double red[2][8];
    double nll = 0.0, ent = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < R; i += (long long)gridDim.x * blockDim.x) {
        const float4 p = __ldg(proba + i);
        const double lik = (double)(p.x + p.y) * pdf[3 * i] + (double)p.z * pdf[3 * i + 1] + (double)p.w * pdf[3 * i + 2];
        nll -= log(lik);
        const float e = p.z * logf(p.z + LOSS_EPS) + (1.f - p.z) * logf(1.f - p.z + LOSS_EPS) + p.w * logf(p.w + LOSS_EPS) +
                        (1.f - p.w) * logf(1.f - p.w + LOSS_EPS);
        ent -= (double)e;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        nll += __shfl_xor_sync(SN2_FULL, nll, d);
        ent += __shfl_xor_sync(SN2_FULL, ent, d);
    }
    if ((threadIdx.x & 31) == 0) {
        red[0][threadIdx.x >> 5] = nll;
        red[1][threadIdx.x >> 5] = ent;
    }
    __syncthreads();
    if (threadIdx.x < 2) {
        double s = 0.0;
        for (int w = 0; w < 8; ++w) s += red[threadIdx.x][w];
        atomicAdd(sums + threadIdx.x, s);
    }
}

// out[0] = mean nll (over R points), out[1] = mean entropy (over 2R values)
__global__ void pointwise_loss_finalize_kernel(const double *__restrict__ sums, long long R, double *__restrict__ out)
{
    out[0] = sums[0] / (double)R;
    out[1] = sums[1] / (double)(2 * R);
}

// dproba = g[0] * d(mean nll)/dproba + g[1] * d(mean ent)/dproba, g = upstream gradient of the two means (device, fp64)
__global__ void __launch_bounds__(256)
pointwise_loss_bwd_kernel(const float4 *__restrict__ proba, const double *__restrict__ pdf, const double *__restrict__ g,
                          long long R, float4 *__restrict__ dproba)
{
    const double gn = g[0] / (double)R;
    const float ge = (float)(g[1] / (double)(2 * R));
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < R; i += (long long)gridDim.x * blockDim.x) {
        const float4 p = __ldg(proba + i);
        const double f0 = pdf[3 * i], f1 = pdf[3 * i + 1], f2 = pdf[3 * i + 2];
        const double s = -gn / ((double)(p.x + p.y) * f0 + (double)p.z * f1 + (double)p.w * f2);
        auto dent = [&](float q) {  // d/dq of -[q log(q+e) + (1-q) log(1-q+e)]
            return -(logf(q + LOSS_EPS) + q / (q + LOSS_EPS) - logf(1.f - q + LOSS_EPS) - (1.f - q) / (1.f - q + LOSS_EPS));
        };
        dproba[i] = make_float4((float)(s * f0), (float)(s * f0), (float)(s * f1) + ge * dent(p.z), (float)(s * f2) + ge * dent(p.w));
    }
}

// pdf[i][c] = linear interpolation of Y[c][.] at z_i = fp32(zrow[i] * z_max) over the sorted knots X (fp64), clamped to
// the grid; zrow = row 2 of the plot's normalised cloud ((B,F,N) layout: plot stride F*N).
__global__ void __launch_bounds__(256)
kde_lut_kernel(const float *__restrict__ cloud, int F, int N, long long R, float z_max, const double *__restrict__ X,
               const double *__restrict__ Y, int K, double *__restrict__ pdf)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < R; i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / N, n = i - b * N;
        const double z = (double)__fmul_rn(__ldg(cloud + (b * F + 2) * N + n), z_max);
        int lo = 0, hi = K - 1;  // largest lo with X[lo] <= z (clamped to [0, K-2])
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (X[mid] <= z) lo = mid;  // np.interp / interp1d bin: X[lo] <= z < X[lo+1]
            else hi = mid;
        }
        const double x0 = X[lo], x1 = X[lo + 1];
        const double zc = fmin(fmax(z, X[0]), X[K - 1]);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const double y0 = Y[(size_t)c * K + lo], y1 = Y[(size_t)c * K + lo + 1];
            pdf[3 * i + c] = (y1 - y0) / (x1 - x0) * (zc - x0) + y0;  // scipy interp1d: slope * (x_new - x_lo) + y_lo
        }
    }
}

}  // namespace sn2

using namespace sn2;

extern "C" int sn2_head_fwd(const float *f1, const float *in_ss, const float *W1, const float *b1, const float *W2, const float *b2,
                            long long R, float *cov, float *proba, void *stream)
{
    if (!f1 || !W1 || !b1 || !W2 || !b2 || !cov || !proba || R <= 0) return SN2_EINVAL;
    if ((reinterpret_cast<uintptr_t>(f1) & 7) || ((reinterpret_cast<uintptr_t>(cov) | reinterpret_cast<uintptr_t>(proba)) & 15)) return SN2_EINVAL;
    const int grid = (int)min((R + 255) / 256, (long long)148 * 8);
    head_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(f1, in_ss, W1, b1, W2, b2, R, reinterpret_cast<float4 *>(cov),
                                                          reinterpret_cast<float4 *>(proba));
    SN2_LAUNCH_CHECK("head_fwd_kernel");
    return SN2_OK;
}

extern "C" int sn2_head_bwd_partials(void) { return HD_NP; }

extern "C" int sn2_head_bwd(const float *f1, const float *in_ss, const float *W1, const float *b1, const float *W2, const float *b2,
                            const float *dcov, const float *dproba, long long R, float *df1, float *partial, int nblk, float *dW1,
                            float *db1, float *dW2, float *db2, void *stream)
{
    if (!f1 || !W1 || !b1 || !W2 || !b2 || !partial || !dW1 || !db1 || !dW2 || !db2 || R <= 0 || nblk <= 0) return SN2_EINVAL;
    if ((reinterpret_cast<uintptr_t>(f1) | reinterpret_cast<uintptr_t>(df1)) & 7) return SN2_EINVAL;
    if ((reinterpret_cast<uintptr_t>(dcov) | reinterpret_cast<uintptr_t>(dproba)) & 15) return SN2_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    const long long ntiles = (R + HB_T - 1) / HB_T;
    const int grid = (int)min(ntiles, (long long)nblk);
    head_bwd_kernel<<<grid, HB_T, 0, st>>>(f1, in_ss, W1, b1, W2, b2, reinterpret_cast<const float4 *>(dcov),
                                           reinterpret_cast<const float4 *>(dproba), R, df1, partial);
    SN2_LAUNCH_CHECK("head_bwd_kernel");
    head_wgrad_reduce_kernel<<<(HD_NP * 32 + 255) / 256, 256, 0, st>>>(partial, grid, dW1, db1, dW2, db2);
    SN2_LAUNCH_CHECK("head_wgrad_reduce_kernel");
    return SN2_OK;
}

extern "C" int sn2_pointwise_loss_fwd(const float *proba, const double *pdf, long long R, double *sums, double *out, void *stream)
{
    if (!proba || !pdf || !sums || !out || R <= 0) return SN2_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    SN2_CUDA_TRY(cudaMemsetAsync(sums, 0, 2 * sizeof(double), st), "pointwise_loss memset");
    const int grid = (int)min((R + 255) / 256, (long long)148 * 4);
    pointwise_loss_fwd_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const float4 *>(proba), pdf, R, sums);
    SN2_LAUNCH_CHECK("pointwise_loss_fwd_kernel");
    pointwise_loss_finalize_kernel<<<1, 1, 0, st>>>(sums, R, out);
    SN2_LAUNCH_CHECK("pointwise_loss_finalize_kernel");
    return SN2_OK;
}

extern "C" int sn2_pointwise_loss_bwd(const float *proba, const double *pdf, const double *g, long long R, float *dproba, void *stream)
{
    if (!proba || !pdf || !g || !dproba || R <= 0) return SN2_EINVAL;
    const int grid = (int)min((R + 255) / 256, (long long)148 * 8);
    pointwise_loss_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4 *>(proba), pdf, g, R,
                                                                    reinterpret_cast<float4 *>(dproba));
    SN2_LAUNCH_CHECK("pointwise_loss_bwd_kernel");
    return SN2_OK;
}

extern "C" int sn2_kde_lut(const float *cloud, int B, int F, int N, float z_max, const double *X, const double *Y, int K, double *pdf,
                           void *stream)
{
    if (!cloud || !X || !Y || !pdf || B <= 0 || F < 3 || N <= 0 || K < 2) return SN2_EINVAL;
    const long long R = (long long)B * N;
    const int grid = (int)min((R + 255) / 256, (long long)148 * 8);
    kde_lut_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(cloud, F, N, R, z_max, X, Y, K, pdf);
    SN2_LAUNCH_CHECK("kde_lut_kernel");
    return SN2_OK;
}
