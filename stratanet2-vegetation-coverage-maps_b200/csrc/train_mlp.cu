// Train-mode MLP block  Linear -> ReLU -> BatchNorm1d(batch statistics)  over R ~ 10^5..10^7 rows with <= 80 input and
// <= 34 output channels: the per-edge message MLPs of SA1 / SA2 and the per-point MLPs of FP2 / FP1
// (reference model/point_net2.py:45-53 `MLP`, applied at :23-29 (PointConv local_nn) and :62-67 (FPModule.nn)).
//
// torch runs this block as GEMM, ReLU, BN statistics, BN transform forward and BN reduce, BN elementwise, ReLU
// backward, two GEMMs backward: ~19 passes over [R, C] arrays, all HBM bound.  Here it is 4 + 6 array touches:
//   lrb_fwd        x -> y = relu(x W^T + b), per-channel sum(y), sum(y^2)               (read x, write y)
//   bn_finalize    statistics -> scale / shift, running-stat update                       ([C] work)
//   bn_apply       z = y * scale + shift                                                  (read y, write z)
//   lrb_bwd_reduce sum(dz), sum(dz * y) per channel                                       (read dz, y)
//   lrb_bwd        dy = relu'(y) * BN'(dz);  dx = dy W;  dW += dy^T x;  db += dy          (read dz, y, x; write dx)
// The batch statistics travel as raw fp64 sums (+ the row count) so that SyncBatchNorm is one all-reduce of that
// small vector between the kernels, forward and backward (host side: sn2/autograd_ops.py::LinReluBN).
// Row-per-thread kernels, 256-row tiles staged through shared memory so every global access is coalesced;
// weights live in shared memory and are read as broadcast float4.
#include "sn2_common.cuh"

namespace sn2 {

constexpr int LRB_T = 256;  // threads per CTA = rows per tile

__device__ __forceinline__ double warp_sum_d(double v)
{
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(SN2_FULL, v, s);
    return v;
}

template <int CO>
__device__ __forceinline__ void load_row(const float *__restrict__ p, float (&v)[CO])
{
    if constexpr (CO % 4 == 0) {
#pragma unroll
        for (int i = 0; i < CO / 4; ++i) {
            const float4 t = __ldg(reinterpret_cast<const float4 *>(p) + i);
            v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < CO / 2; ++i) {
            const float2 t = __ldg(reinterpret_cast<const float2 *>(p) + i);
            v[2 * i] = t.x; v[2 * i + 1] = t.y;
        }
    }
}
template <int CO>
__device__ __forceinline__ void store_row(float *__restrict__ p, const float (&v)[CO])
{
    if constexpr (CO % 4 == 0) {
#pragma unroll
        for (int i = 0; i < CO / 4; ++i)
            reinterpret_cast<float4 *>(p)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    } else {
#pragma unroll
        for (int i = 0; i < CO / 2; ++i) reinterpret_cast<float2 *>(p)[i] = make_float2(v[2 * i], v[2 * i + 1]);
    }
}

// Per-thread channel sums -> CTA sums in shared memory (fp64) -> one fp64 atomic per channel and CTA.
template <int CO>
__device__ __forceinline__ void cta_channel_sums(const float (&a)[CO], const float (&b)[CO], double *red, double *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 0; o < CO; ++o) {
        const double s = warp_sum_d((double)a[o]), q = warp_sum_d((double)b[o]);
        if (lane == 0) {
            atomicAdd(red + o, s);
            atomicAdd(red + CO + o, q);
        }
    }
    __syncthreads();
    for (int o = threadIdx.x; o < 2 * CO; o += blockDim.x) atomicAdd(out + o, red[o]);
}

// ---------------------------------------------------------------------------------------------------------------
// forward: y = relu(x W^T + b), stats[0..CO) += sum y, stats[CO..2CO) += sum y^2
// ---------------------------------------------------------------------------------------------------------------
template <int CI, int CO>
struct LrbFwd {
    static constexpr int WARPS = LRB_T / 32;
    static constexpr int COP = (CO + 3) & ~3;
    static constexpr int XS = (CI & 1) ? CI : CI + 1;  // odd row stride: conflict-free row-per-thread reads
    static constexpr size_t SMEM = sizeof(double) * 2 * CO + sizeof(float) * ((size_t)CI * COP + COP + (size_t)LRB_T * XS);
};

// Every warp owns 32-row tiles: it copies the tile's 32 * CI contiguous floats into its private shared-memory slab
// (coalesced), then each lane runs the layer on its row.  No CTA barrier inside the loop, so the warps of an SM
// drift apart and the loads of some overlap the FMAs of others.
template <int CI, int CO>
__global__ void __launch_bounds__(LRB_T, (CO > 16 ? 1 : 2))
lrb_fwd_kernel(const float *__restrict__ x, const float *__restrict__ W, const float *__restrict__ b, long long R,
               const int *__restrict__ rows_dev, float *__restrict__ y, double *__restrict__ stats)
{
    if (rows_dev) R = min(R, (long long)__ldg(rows_dev));
    using L = LrbFwd<CI, CO>;
    constexpr int COP = L::COP, XS = L::XS;
    extern __shared__ __align__(16) unsigned char lrb_smem[];
    double *red = reinterpret_cast<double *>(lrb_smem);
    float *Wt = reinterpret_cast<float *>(red + 2 * CO);  // [CI][COP]: Wt[k][o] = W[o][k]
    float *bS = Wt + CI * COP;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float *xS = bS + COP + warp * (32 * XS);              // this warp's [32][XS] slab
    for (int e = tid; e < CI * COP; e += LRB_T) {
        const int k = e / COP, o = e - k * COP;
        Wt[e] = o < CO ? __ldg(W + o * CI + k) : 0.f;
    }
    for (int o = tid; o < COP; o += LRB_T) bS[o] = o < CO ? __ldg(b + o) : 0.f;
    for (int o = tid; o < 2 * CO; o += LRB_T) red[o] = 0.0;
    __syncthreads();
    float s1[CO], s2[CO];
#pragma unroll
    for (int o = 0; o < CO; ++o) s1[o] = s2[o] = 0.f;
    const long long ntiles = (R + 31) / 32;
    for (long long tile = (long long)blockIdx.x * L::WARPS + warp; tile < ntiles; tile += (long long)gridDim.x * L::WARPS) {
        const long long base = tile * 32;
        const int rows = (int)min(32LL, R - base);
        const float *xt = x + base * CI;
        __syncwarp();
#pragma unroll
        for (int j = 0; j < CI; ++j) {
            const int e = j * 32 + lane;  // element e of the tile = (row e / CI, column e % CI)
            if (e < rows * CI) {
                const int r = e / CI, k = e - r * CI;
                xS[r * XS + k] = __ldg(xt + e);
            }
        }
        __syncwarp();
        if (lane < rows) {
            float acc[COP];
#pragma unroll
            for (int o = 0; o < COP; ++o) acc[o] = bS[o];
#pragma unroll
            for (int k = 0; k < CI; ++k) {
                const float xk = xS[lane * XS + k];
#pragma unroll
                for (int o4 = 0; o4 < COP / 4; ++o4) {
                    const float4 w = *reinterpret_cast<const float4 *>(Wt + k * COP + 4 * o4);
                    acc[4 * o4] = fmaf(xk, w.x, acc[4 * o4]);
                    acc[4 * o4 + 1] = fmaf(xk, w.y, acc[4 * o4 + 1]);
                    acc[4 * o4 + 2] = fmaf(xk, w.z, acc[4 * o4 + 2]);
                    acc[4 * o4 + 3] = fmaf(xk, w.w, acc[4 * o4 + 3]);
                }
            }
            float out[CO];
#pragma unroll
            for (int o = 0; o < CO; ++o) {
                out[o] = fmaxf(acc[o], 0.f);
                s1[o] += out[o];
                s2[o] = fmaf(out[o], out[o], s2[o]);
            }
            store_row<CO>(y + (base + lane) * CO, out);
        }
    }
    __syncwarp();
    cta_channel_sums<CO>(s1, s2, red, stats);
}

// statistics (+ count at stats[2*Co]) -> ss = [scale, shift, mean, invstd]; running statistics as torch BatchNorm.
__global__ void bn_finalize_kernel(const double *__restrict__ stats, const float *__restrict__ gamma,
                                   const float *__restrict__ beta, float eps, float momentum, float *running_mean,
                                   float *running_var, long long *num_batches_tracked, float *__restrict__ ss, int Co)
{
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= Co) return;
    if (o == 0 && num_batches_tracked) *num_batches_tracked += 1;
    const double n = stats[2 * Co];
    const double mean = stats[o] / n;
    const double var = fmax(stats[Co + o] / n - mean * mean, 0.0);
    const float inv = (float)(1.0 / sqrt(var + (double)eps));
    const float s = gamma[o] * inv;
    ss[o] = s;
    ss[Co + o] = fmaf(-(float)mean, s, beta[o]);
    ss[2 * Co + o] = (float)mean;
    ss[3 * Co + o] = inv;
    if (running_mean) running_mean[o] = (1.f - momentum) * running_mean[o] + momentum * (float)mean;
    if (running_var) running_var[o] = (1.f - momentum) * running_var[o] + momentum * (float)(n > 1.0 ? var * n / (n - 1.0) : var);
}

__global__ void set_count_kernel(double *stats, int Co, long long R, const int *__restrict__ rows_dev)
{
    stats[2 * Co] = (double)(rows_dev ? min(R, (long long)*rows_dev) : R);
}

// BatchNorm affine gradients from one rank's raw sums: dbeta = sum dz, dgamma = sum dz * yhat = invstd * (S2 - mean * S1)
__global__ void bn_param_grad_kernel(const double *__restrict__ sums, const float *__restrict__ ss, int Co,
                                     float *__restrict__ dgamma, float *__restrict__ dbeta)
{
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= Co) return;
    const double S1 = sums[o], S2 = sums[Co + o];
    dbeta[o] = (float)S1;
    dgamma[o] = (float)((double)ss[3 * Co + o] * (S2 - (double)ss[2 * Co + o] * S1));
}

// z = y * scale[c] + shift[c]; two channels per thread (every supported Co is even).
__global__ void __launch_bounds__(256)
bn_apply_kernel(const float2 *__restrict__ y, const float *__restrict__ ss, long long R, const int *__restrict__ rows_dev,
                int Co, float2 *__restrict__ z)
{
    if (rows_dev) R = min(R, (long long)__ldg(rows_dev));
    const long long n2 = R * (Co >> 1);
    const int half = Co >> 1;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const int step = (int)(stride % half);
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int c2 = (int)(i % half);
    for (; i < n2; i += stride) {
        const int c = 2 * c2;
        const float2 v = __ldg(y + i);
        z[i] = make_float2(fmaf(v.x, __ldg(ss + c), __ldg(ss + Co + c)), fmaf(v.y, __ldg(ss + c + 1), __ldg(ss + Co + c + 1)));
        c2 += step;
        if (c2 >= half) c2 -= half;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// backward, pass 1: sums[0..CO) += sum dz, sums[CO..2CO) += sum dz * y
// ---------------------------------------------------------------------------------------------------------------
template <int CO>
__global__ void __launch_bounds__(LRB_T, (CO > 16 ? 1 : 2))
lrb_bwd_reduce_kernel(const float *__restrict__ dz, const float *__restrict__ y, long long R,
                      const int *__restrict__ rows_dev, double *__restrict__ sums)
{
    if (rows_dev) R = min(R, (long long)__ldg(rows_dev));
    __shared__ double red[2 * CO];
    for (int o = threadIdx.x; o < 2 * CO; o += LRB_T) red[o] = 0.0;
    float a1[CO], a2[CO];
#pragma unroll
    for (int o = 0; o < CO; ++o) a1[o] = a2[o] = 0.f;
    for (long long row = (long long)blockIdx.x * LRB_T + threadIdx.x; row < R; row += (long long)gridDim.x * LRB_T) {
        float d[CO], v[CO];
        load_row<CO>(dz + row * CO, d);
        load_row<CO>(y + row * CO, v);
#pragma unroll
        for (int o = 0; o < CO; ++o) {
            a1[o] += d[o];
            a2[o] = fmaf(d[o], v[o], a2[o]);
        }
    }
    __syncthreads();
    cta_channel_sums<CO>(a1, a2, red, sums);
}

// ---------------------------------------------------------------------------------------------------------------
// backward, pass 2.  Per row: yhat = (y - mean) * invstd,
//   dy = [y > 0] * gamma * invstd * (dz - mean(dz) - yhat * mean(dz * yhat))      (BatchNorm, then ReLU)
//   dx = dy W;   dW[o][:] += dy[o] * x;   db[o] += dy[o]
// sums are the (all-reduced) raw sums of pass 1, count the (global) row count in stats[2*CO].
// ---------------------------------------------------------------------------------------------------------------
template <int CI, int CO>
struct LrbBwd {
    static constexpr int TR = CI > 48 ? 128 : 256;  // rows per tile = threads per CTA
    static constexpr int CIP = (CI + 3) & ~3;
    static constexpr int G = TR / CO;               // row groups of the weight-gradient phase
    static constexpr int NP = CO * (CI + 1);        // floats of one partial [dW | db]
    static constexpr int XS_FLOATS = (TR * CIP > G * NP) ? TR * CIP : G * NP;
    static constexpr size_t SMEM = sizeof(float) * ((size_t)CO * CIP + 4 * CO + XS_FLOATS + (size_t)TR * CO + (size_t)TR * CI);
};

template <int CI, int CO>
__global__ void __launch_bounds__(LrbBwd<CI, CO>::TR, (CI > 48 ? 1 : 2))
lrb_bwd_kernel(const float *__restrict__ dz, const float *__restrict__ y, const float *__restrict__ x,
               const float *__restrict__ W, const float *__restrict__ ss, const double *__restrict__ sums,
               const double *__restrict__ stats, long long R, const int *__restrict__ rows_dev, float *__restrict__ dx,
               float *__restrict__ partial)
{
    if (rows_dev) R = min(R, (long long)__ldg(rows_dev));
    using L = LrbBwd<CI, CO>;
    constexpr int TR = L::TR, CIP = L::CIP, G = L::G, NP = L::NP;
    extern __shared__ __align__(16) unsigned char lrb_smem[];
    float *Ws = reinterpret_cast<float *>(lrb_smem);  // [CO][CIP] = W[o][k], zero padded
    float *cA = Ws + CO * CIP;                        // dy = mask * (cA*dz + cB*y + cC)
    float *cB = cA + CO;
    float *cC = cB + CO;
    float *xS = cC + 2 * CO;                          // [TR][CIP] (later: group partials)
    float *dyS = xS + L::XS_FLOATS;                   // [TR][CO]
    float *dxS = dyS + TR * CO;                       // [TR][CI]
    const int tid = threadIdx.x;
    for (int e = tid; e < CO * CIP; e += TR) {
        const int o = e / CIP, k = e - o * CIP;
        Ws[e] = k < CI ? __ldg(W + o * CI + k) : 0.f;
    }
    for (int e = tid; e < TR * CIP; e += TR) xS[e] = 0.f;
    if (tid < CO) {
        // dz - m1 - (y - mean) * inv * m2, m1 = S1/n, m2 = mean(dz * yhat) = inv * (S2 - mean * S1) / n, times gamma * inv
        const double n = stats[2 * CO];
        const float sc = ss[tid], mean = ss[2 * CO + tid], inv = ss[3 * CO + tid];
        const double S1 = sums[tid], S2 = sums[CO + tid];
        const float m1 = (float)(S1 / n);
        const float m2 = (float)((double)inv * (S2 - (double)mean * S1) / n);
        const float kb = -inv * m2;
        cA[tid] = sc;
        cB[tid] = sc * kb;
        cC[tid] = sc * (-m1 - mean * kb);
    }
    float acc[CI + 1];
#pragma unroll
    for (int i = 0; i <= CI; ++i) acc[i] = 0.f;
    const int g = tid / CO, o_w = tid - g * CO;
    const long long ntiles = (R + TR - 1) / TR;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        __syncthreads();
        const long long base = tile * TR;
        const int rows = (int)min((long long)TR, R - base);
        const float *xt = x + base * CI;
        for (int e = tid; e < rows * CI; e += TR) {
            const int r = e / CI, k = e - r * CI;
            xS[r * CIP + k] = __ldg(xt + e);
        }
        if (tid < rows) {
            float d[CO], v[CO];
            load_row<CO>(dz + (base + tid) * CO, d);
            load_row<CO>(y + (base + tid) * CO, v);
#pragma unroll
            for (int o = 0; o < CO; ++o) d[o] = v[o] > 0.f ? fmaf(cA[o], d[o], fmaf(cB[o], v[o], cC[o])) : 0.f;
            store_row<CO>(dyS + tid * CO, d);
            if (dx) {
                float dxr[CIP];
#pragma unroll
                for (int k = 0; k < CIP; ++k) dxr[k] = 0.f;
#pragma unroll
                for (int o = 0; o < CO; ++o) {
#pragma unroll
                    for (int k4 = 0; k4 < CIP / 4; ++k4) {
                        const float4 w = *reinterpret_cast<const float4 *>(Ws + o * CIP + 4 * k4);
                        dxr[4 * k4] = fmaf(d[o], w.x, dxr[4 * k4]);
                        dxr[4 * k4 + 1] = fmaf(d[o], w.y, dxr[4 * k4 + 1]);
                        dxr[4 * k4 + 2] = fmaf(d[o], w.z, dxr[4 * k4 + 2]);
                        dxr[4 * k4 + 3] = fmaf(d[o], w.w, dxr[4 * k4 + 3]);
                    }
                }
#pragma unroll
                for (int k = 0; k < CI; ++k) dxS[tid * CI + k] = dxr[k];
            }
        }
        __syncthreads();
        if (dx) {
            float *dxt = dx + base * CI;
            for (int e = tid; e < rows * CI; e += TR) dxt[e] = dxS[e];
        }
        if (g < G) {
            // rows r = g, g + G, ...: dyS address (r * CO + o) = (G * it) * CO + tid -> contiguous over the CTA
            for (int r = g; r < rows; r += G) {
                const float dv = dyS[r * CO + o_w];
#pragma unroll
                for (int k4 = 0; k4 < CIP / 4; ++k4) {
                    const float4 xv = *reinterpret_cast<const float4 *>(xS + r * CIP + 4 * k4);
                    if (4 * k4 < CI) acc[4 * k4] = fmaf(dv, xv.x, acc[4 * k4]);
                    if (4 * k4 + 1 < CI) acc[4 * k4 + 1] = fmaf(dv, xv.y, acc[4 * k4 + 1]);
                    if (4 * k4 + 2 < CI) acc[4 * k4 + 2] = fmaf(dv, xv.z, acc[4 * k4 + 2]);
                    if (4 * k4 + 3 < CI) acc[4 * k4 + 3] = fmaf(dv, xv.w, acc[4 * k4 + 3]);
                }
                acc[CI] += dv;
            }
        }
    }
    __syncthreads();
    if (g < G) {
#pragma unroll
        for (int i = 0; i <= CI; ++i) xS[(g * CO + o_w) * (CI + 1) + i] = acc[i];
    }
    __syncthreads();
    for (int t = tid; t < NP; t += TR) {
        float s = 0.f;
        for (int gg = 0; gg < G; ++gg) s += xS[gg * NP + t];
        partial[(size_t)blockIdx.x * NP + t] = s;
    }
}

// fixed-order reduction of the CTA partials (layout [Co][Ci + 1], last column = db): one warp per output, lanes
// stride over the CTAs, xor-shuffle tree (deterministic)
__global__ void __launch_bounds__(256)
lrb_wgrad_reduce_kernel(const float *__restrict__ partial, int nblk, int Co, int CI, float *__restrict__ dW, float *__restrict__ db)
{
    const int n = Co * (CI + 1);
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (t >= n) return;
    float s = 0.f;
    for (int b = lane; b < nblk; b += 32) s += __ldg(partial + (size_t)b * n + t);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(SN2_FULL, s, d);
    if (lane == 0) {
        const int o = t / (CI + 1), i = t - o * (CI + 1);
        if (i < CI) dW[o * CI + i] = s;
        else db[o] = s;
    }
}

template <int CI, int CO>
static int launch_lrb_fwd(const float *x, const float *W, const float *b, long long R, const int *rows_dev, float *y,
                          double *stats, cudaStream_t st)
{
    using L = LrbFwd<CI, CO>;
    auto kern = lrb_fwd_kernel<CI, CO>;
    SN2_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::SMEM), "lrb_fwd attr");
    SN2_CUDA_TRY(cudaMemsetAsync(stats, 0, sizeof(double) * 2 * CO, st), "lrb_fwd memset");
    set_count_kernel<<<1, 1, 0, st>>>(stats, CO, R, rows_dev);
    const long long nblocks = ((R + 31) / 32 + L::WARPS - 1) / L::WARPS;
    const int grid = (int)min(nblocks, (long long)148 * 2);
    kern<<<grid, LRB_T, L::SMEM, st>>>(x, W, b, R, rows_dev, y, stats);
    SN2_LAUNCH_CHECK("lrb_fwd_kernel");
    return SN2_OK;
}

template <int CO>
static int launch_lrb_bwd_reduce(const float *dz, const float *y, long long R, const int *rows_dev, double *sums, cudaStream_t st)
{
    SN2_CUDA_TRY(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * CO, st), "lrb_bwd_reduce memset");
    const long long ntiles = (R + LRB_T - 1) / LRB_T;
    const int grid = (int)min(ntiles, (long long)148 * 2);
    lrb_bwd_reduce_kernel<CO><<<grid, LRB_T, 0, st>>>(dz, y, R, rows_dev, sums);
    SN2_LAUNCH_CHECK("lrb_bwd_reduce_kernel");
    return SN2_OK;
}

template <int CI, int CO>
static int launch_lrb_bwd(const float *dz, const float *y, const float *x, const float *W, const float *ss, const double *sums,
                          const double *stats, long long R, const int *rows_dev, float *dx, float *partial, int nblk, float *dW,
                          float *db, cudaStream_t st)
{
    using L = LrbBwd<CI, CO>;
    auto kern = lrb_bwd_kernel<CI, CO>;
    SN2_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::SMEM), "lrb_bwd attr");
    kern<<<nblk, L::TR, L::SMEM, st>>>(dz, y, x, W, ss, sums, stats, R, rows_dev, dx, partial);
    SN2_LAUNCH_CHECK("lrb_bwd_kernel");
    lrb_wgrad_reduce_kernel<<<(L::NP * 32 + 255) / 256, 256, 0, st>>>(partial, nblk, CO, CI, dW, db);
    SN2_LAUNCH_CHECK("lrb_wgrad_reduce_kernel");
    return SN2_OK;
}

}  // namespace sn2

// (Ci, Co) pairs of the train-mode blocks that see >= 65 536 rows: SA1 [11,16,16], SA2 [19,32], FP2 [80,34], FP1 [42,34]
#define SN2_LRB_SHAPES(X) X(11, 16) X(16, 16) X(19, 32) X(80, 34) X(42, 34)

extern "C" int sn2_lrb_supported(int Co, int Ci)
{
#define X(ci, co) if (Ci == ci && Co == co) return 1;
    SN2_LRB_SHAPES(X)
#undef X
    return 0;
}

extern "C" int sn2_lrb_fwd(const float *x, const float *W, const float *b, long long R, const int *rows_dev, int Co, int Ci,
                           float *y, double *stats, void *stream)
{
    if (!x || !W || !b || !y || !stats || R <= 0) return SN2_EINVAL;
#define X(ci, co) if (Ci == ci && Co == co) return sn2::launch_lrb_fwd<ci, co>(x, W, b, R, rows_dev, y, stats, (cudaStream_t)stream);
    SN2_LRB_SHAPES(X)
#undef X
    return SN2_EUNSUPPORTED;
}

extern "C" int sn2_bn_finalize(const double *stats, const float *gamma, const float *beta, float eps, float momentum,
                               float *running_mean, float *running_var, long long *num_batches_tracked, float *ss, int Co,
                               void *stream)
{
    if (!stats || !gamma || !beta || !ss || Co <= 0) return SN2_EINVAL;
    sn2::bn_finalize_kernel<<<(Co + 63) / 64, 64, 0, (cudaStream_t)stream>>>(stats, gamma, beta, eps, momentum, running_mean,
                                                                           running_var, num_batches_tracked, ss, Co);
    SN2_LAUNCH_CHECK("bn_finalize_kernel");
    return SN2_OK;
}

extern "C" int sn2_bn_apply(const float *y, const float *ss, long long R, const int *rows_dev, int Co, float *z, void *stream)
{
    if (!y || !ss || !z || R <= 0 || Co <= 0 || (Co & 1)) return SN2_EINVAL;
    const long long n2 = R * Co / 2;
    const int grid = (int)((n2 + 255) / 256 < 148 * 8 ? (n2 + 255) / 256 : 148 * 8);
    sn2::bn_apply_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2 *>(y), ss, R, rows_dev, Co,
                                                               reinterpret_cast<float2 *>(z));
    SN2_LAUNCH_CHECK("bn_apply_kernel");
    return SN2_OK;
}

extern "C" int sn2_lrb_bwd_reduce(const float *dz, const float *y, long long R, const int *rows_dev, int Co, double *sums,
                                  void *stream)
{
    if (!dz || !y || !sums || R <= 0) return SN2_EINVAL;
    switch (Co) {
    case 16: return sn2::launch_lrb_bwd_reduce<16>(dz, y, R, rows_dev, sums, (cudaStream_t)stream);
    case 32: return sn2::launch_lrb_bwd_reduce<32>(dz, y, R, rows_dev, sums, (cudaStream_t)stream);
    case 34: return sn2::launch_lrb_bwd_reduce<34>(dz, y, R, rows_dev, sums, (cudaStream_t)stream);
    default: return SN2_EUNSUPPORTED;
    }
}

extern "C" int sn2_bn_param_grad(const double *sums, const float *ss, int Co, float *dgamma, float *dbeta, void *stream)
{
    if (!sums || !ss || !dgamma || !dbeta || Co <= 0) return SN2_EINVAL;
    sn2::bn_param_grad_kernel<<<(Co + 63) / 64, 64, 0, (cudaStream_t)stream>>>(sums, ss, Co, dgamma, dbeta);
    SN2_LAUNCH_CHECK("bn_param_grad_kernel");
    return SN2_OK;
}

extern "C" int sn2_lrb_bwd(const float *dz, const float *y, const float *x, const float *W, const float *ss,
                           const double *sums, const double *stats, long long R, const int *rows_dev, int Co, int Ci, float *dx,
                           float *partial, int nblk, float *dW, float *db, void *stream)
{
    if (!dz || !y || !x || !W || !ss || !sums || !stats || !partial || !dW || !db || R <= 0 || nblk <= 0) return SN2_EINVAL;
#define X(ci, co)                  \
    if (Ci == ci && Co == co)      \
        return sn2::launch_lrb_bwd<ci, co>(dz, y, x, W, ss, sums, stats, R, rows_dev, dx, partial, nblk, dW, db, (cudaStream_t)stream);
    SN2_LRB_SHAPES(X)
#undef X
    return SN2_EUNSUPPORTED;
}

// Single-process block in one call each way (no all-reduce between the kernels): what LinReluBN uses when the
// block's BatchNorm is not a SyncBatchNorm.
extern "C" int sn2_lrb_block_fwd(const float *x, const float *W, const float *b, const float *gamma, const float *beta,
                                 float eps, float momentum, float *running_mean, float *running_var,
                                 long long *num_batches_tracked, long long R, const int *rows_dev, int Co, int Ci, float *y,
                                 double *stats, float *ss, float *z, void *stream)
{
    if (int rc = sn2_lrb_fwd(x, W, b, R, rows_dev, Co, Ci, y, stats, stream)) return rc;
    if (int rc = sn2_bn_finalize(stats, gamma, beta, eps, momentum, running_mean, running_var, num_batches_tracked, ss, Co, stream))
        return rc;
    return sn2_bn_apply(y, ss, R, rows_dev, Co, z, stream);
}

extern "C" int sn2_lrb_block_bwd(const float *dz, const float *y, const float *x, const float *W, const float *ss,
                                 const double *stats, long long R, const int *rows_dev, int Co, int Ci, double *sums,
                                 float *dgamma, float *dbeta, float *dx, float *partial, int nblk, float *dW, float *db,
                                 void *stream)
{
    if (int rc = sn2_lrb_bwd_reduce(dz, y, R, rows_dev, Co, sums, stream)) return rc;
    if (int rc = sn2_bn_param_grad(sums, ss, Co, dgamma, dbeta, stream)) return rc;
    return sn2_lrb_bwd(dz, y, x, W, ss, sums, stats, R, rows_dev, Co, Ci, dx, partial, nblk, dW, db, stream);
}
