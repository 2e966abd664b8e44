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
__device__ __forceinline__ void load_row_s(const float *p, float (&v)[CO])  // from shared memory
{
    if constexpr (CO % 4 == 0) {
#pragma unroll
        for (int i = 0; i < CO / 4; ++i) {
            const float4 t = reinterpret_cast<const float4 *>(p)[i];
            v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < CO / 2; ++i) {
            const float2 t = reinterpret_cast<const float2 *>(p)[i];
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
// ---- TMA bulk copy + mbarrier helpers (1-D cp.async.bulk: contiguous global bytes -> shared memory) -------------
__device__ __forceinline__ unsigned smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_addr(bar)), "r"(count));
}
// one arrival that also arms the barrier phase with the bytes the following bulk copies will deliver
__device__ __forceinline__ void mbar_expect(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
// bulk copy global -> shared; its bytes count towards the phase of `bar`
__device__ __forceinline__ void tma_copy_1d(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_addr(dst)),
                 "l"(src), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
    mbar_expect(bar, bytes);
    tma_copy_1d(dst, src, bytes, bar);
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    unsigned done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(smem_addr(bar)), "r"(parity) : "memory");
}

template <int CI, int CO>
struct LrbFwd {
    static constexpr int WARPS = LRB_T / 32;
    static constexpr int COP = (CO + 3) & ~3;
    static constexpr int STAGES = CI > 48 ? 2 : 3;      // 32-row slabs in flight per warp
    static constexpr int SLAB = 32 * CI;                 // floats; 128 * CI bytes: a multiple of 16 for any CI
    static constexpr int CIP = (CI + 3) & ~3;
    static constexpr size_t SMEM = sizeof(double) * 2 * CO + sizeof(float) * ((size_t)CI * COP + COP + 2 * CIP + (size_t)WARPS * STAGES * SLAB) +
                                   sizeof(unsigned long long) * WARPS * STAGES;
};

// Every warp owns 32-row tiles.  A tile's 32 * CI floats are contiguous in global memory: lane 0 fetches them with
// one TMA bulk copy into a slab of the warp's private ring, STAGES tiles ahead of the one being computed, and an
// mbarrier per slab tells the warp when the bytes have landed.  No CTA barrier inside the loop and no register
// staging: the loads of the next tiles are in flight while each lane runs the layer on its row.
template <int CI, int CO>
__global__ void __launch_bounds__(LRB_T, (CO > 16 ? 1 : 2))
lrb_fwd_kernel(const float *__restrict__ x, const float *__restrict__ in_ss, const float *__restrict__ W,
               const float *__restrict__ b, long long R, const int *__restrict__ rows_dev, float *__restrict__ y,
               double *__restrict__ stats)
{
    if (rows_dev) R = min(R, (long long)__ldg(rows_dev));
    using L = LrbFwd<CI, CO>;
    constexpr int COP = L::COP, STAGES = L::STAGES, SLAB = L::SLAB, CIP = L::CIP;
    extern __shared__ __align__(16) unsigned char lrb_smem[];
    double *red = reinterpret_cast<double *>(lrb_smem);
    float *Wt = reinterpret_cast<float *>(red + 2 * CO);  // [CI][COP]: Wt[k][o] = W[o][k]
    float *bS = Wt + CI * COP;
    float *inS = bS + COP;                                // input affine: scale [CIP] | shift [CIP] (identity if none)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float *ring = inS + 2 * CIP + warp * (STAGES * SLAB); // this warp's STAGES slabs of [32][CI]
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(inS + 2 * CIP + L::WARPS * STAGES * SLAB) + warp * STAGES;
    for (int k = tid; k < CI; k += LRB_T) {
        inS[k] = in_ss ? __ldg(in_ss + k) : 1.f;
        inS[CIP + k] = in_ss ? __ldg(in_ss + CI + k) : 0.f;
    }
    for (int e = tid; e < CI * COP; e += LRB_T) {
        const int k = e / COP, o = e - k * COP;
        Wt[e] = o < CO ? __ldg(W + o * CI + k) : 0.f;
    }
    for (int o = tid; o < COP; o += LRB_T) bS[o] = o < CO ? __ldg(b + o) : 0.f;
    for (int o = tid; o < 2 * CO; o += LRB_T) red[o] = 0.0;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) mbar_init(bars + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    float s1[CO], s2[CO];
#pragma unroll
    for (int o = 0; o < CO; ++o) s1[o] = s2[o] = 0.f;
    const long long ntiles = (R + 31) / 32, nfull = R / 32;
    const long long tile0 = (long long)blockIdx.x * L::WARPS + warp, stride = (long long)gridDim.x * L::WARPS;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            const long long t = tile0 + s * stride;
            if (t < nfull) tma_load_1d(ring + s * SLAB, x + t * SLAB, SLAB * 4, bars + s);
        }
    }
    int stage = 0;
    unsigned parity = 0;
    for (long long tile = tile0; tile < ntiles; tile += stride) {
        const long long base = tile * 32;
        const float *xS = ring + stage * SLAB;
        int rows = 32;
        if (tile < nfull) {
            mbar_wait(bars + stage, parity);
        } else {  // the ragged last tile of the array: its byte count need not be a multiple of 16 -> plain loads
            rows = (int)(R - base);
            float *dst = ring + stage * SLAB;
            for (int e = lane; e < rows * CI; e += 32) dst[e] = __ldg(x + base * CI + e);
            __syncwarp();
        }
        if (lane < rows) {
            float acc[COP];
#pragma unroll
            for (int o = 0; o < COP; ++o) acc[o] = bS[o];
            // dense [32][CI] slab: rows of even width are read as float4 / float2 (a scalar read at stride CI = 16
            // would be a 16-way bank conflict), odd widths are conflict-free as they are
            constexpr int V = (CI % 4 == 0) ? 4 : ((CI % 2 == 0) ? 2 : 1);
#pragma unroll
            for (int kv = 0; kv < CI / V; ++kv) {
                float xv[V];
                if constexpr (V == 4) {
                    const float4 t = *reinterpret_cast<const float4 *>(xS + lane * CI + 4 * kv);
                    xv[0] = t.x; xv[1] = t.y; xv[2] = t.z; xv[3] = t.w;
                } else if constexpr (V == 2) {
                    const float2 t = *reinterpret_cast<const float2 *>(xS + lane * CI + 2 * kv);
                    xv[0] = t.x; xv[1] = t.y;
                } else {
                    xv[0] = xS[lane * CI + kv];
                }
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    const int k = kv * V + j;
                    const float xk = fmaf(xv[j], inS[k], inS[CIP + k]);  // the producer block's BatchNorm, applied on load
#pragma unroll
                    for (int o4 = 0; o4 < COP / 4; ++o4) {
                        const float4 w = *reinterpret_cast<const float4 *>(Wt + k * COP + 4 * o4);
                        acc[4 * o4] = fmaf(xk, w.x, acc[4 * o4]);
                        acc[4 * o4 + 1] = fmaf(xk, w.y, acc[4 * o4 + 1]);
                        acc[4 * o4 + 2] = fmaf(xk, w.z, acc[4 * o4 + 2]);
                        acc[4 * o4 + 3] = fmaf(xk, w.w, acc[4 * o4 + 3]);
                    }
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
        __syncwarp();  // every lane has read its row: the slab can be refilled
        const long long nt = tile + STAGES * stride;
        if (lane == 0 && nt < nfull) tma_load_1d(ring + stage * SLAB, x + nt * SLAB, SLAB * 4, bars + stage);
        if (++stage == STAGES) {
            stage = 0;
            parity ^= 1u;
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
//
// 128-row tiles.  The three input tiles (x, dz, y: contiguous row blocks) arrive by TMA bulk copies into a two-stage
// ring, one tile ahead of the compute; dx leaves through shared memory with a TMA bulk store.  Phase 1: thread =
// row (dy to shared memory, dx).  Phase 2: thread = (row group g, input column i): dW[:, i] += dy[r][:] * x[r][i]
// over the rows r = g, g + G, ... of the tile, with column i = CI standing for the bias (x = 1); x is read dense
// (consecutive i: conflict-free), dy rows as broadcast float4.
// ---------------------------------------------------------------------------------------------------------------
template <int CI, int CO>
struct LrbBwd {
    static constexpr int TR = 128;                  // rows per tile = threads per CTA
    static constexpr int COP = (CO + 3) & ~3;
    static constexpr int DYS = COP + 4;             // dy row stride: float4 rows, conflict-free for a quarter warp
    static constexpr int CIP = (CI + 3) & ~3;
    static constexpr int G = TR / (CI + 1);         // row groups of the weight-gradient phase
    static constexpr int NP = CO * (CI + 1);        // floats of one partial [dW | db]
    static constexpr int STAGE = TR * (CI + 2 * CO); // floats of one stage: x | dz | y tiles
    static constexpr int RED = (G * NP > TR * CI) ? G * NP : TR * CI;  // dx tile, reused for the group partials at the end
    static constexpr int MINB = (2 * STAGE + TR * DYS + RED + CO * CIP) * 4 <= 72 * 1024 ? 3 : ((2 * STAGE + TR * DYS + RED + CO * CIP) * 4 <= 110 * 1024 ? 2 : 1);
    static constexpr size_t SMEM = sizeof(float) * ((size_t)2 * STAGE + (size_t)TR * DYS + RED + (size_t)CO * CIP + 4 * CO) + 2 * sizeof(unsigned long long);
    static_assert(RED % 4 == 0 && SMEM <= 227 * 1024, "shared-memory layout");
};

template <int CI, int CO>
__global__ void __launch_bounds__(LrbBwd<CI, CO>::TR, LrbBwd<CI, CO>::MINB)
lrb_bwd_kernel(const float *__restrict__ dz, const float *__restrict__ y, const float *__restrict__ x,
               const float *__restrict__ in_ss, const float *__restrict__ W, const float *__restrict__ ss, const double *__restrict__ sums,
               const double *__restrict__ stats, long long R, const int *__restrict__ rows_dev, float *__restrict__ dx,
               float *__restrict__ partial)
{
    if (rows_dev) R = min(R, (long long)__ldg(rows_dev));
    using L = LrbBwd<CI, CO>;
    constexpr int TR = L::TR, COP = L::COP, DYS = L::DYS, CIP = L::CIP, G = L::G, NP = L::NP, STAGE = L::STAGE;
    extern __shared__ __align__(16) unsigned char lrb_smem[];
    float *ring = reinterpret_cast<float *>(lrb_smem);  // 2 x { x [TR][CI] | dz [TR][CO] | y [TR][CO] }
    float *dyS = ring + 2 * STAGE;                        // [TR][DYS]
    float *dxS = dyS + TR * DYS;                          // [TR][CI]  (group partials at the very end)
    float *Ws = dxS + L::RED;                             // [CO][CIP] = W[o][k], zero padded
    float *cA = Ws + CO * CIP;                            // dy = mask * (cA*dz + cB*y + cC)
    float *cB = cA + CO;
    float *cC = cB + CO;
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(cC + 2 * CO);
    const int tid = threadIdx.x;
    for (int e = tid; e < CO * CIP; e += TR) {
        const int o = e / CIP, k = e - o * CIP;
        Ws[e] = k < CI ? __ldg(W + o * CI + k) : 0.f;
    }
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
    if (tid == 0) {
        mbar_init(bars, 1);
        mbar_init(bars + 1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    const long long ntiles = (R + TR - 1) / TR, nfull = R / TR;
    // one barrier phase per tile: three bulk copies, their byte counts add up on the same mbarrier
    auto fetch = [&](long long tile, int st) {
        float *dst = ring + st * STAGE;
        mbar_expect(bars + st, STAGE * 4);
        tma_copy_1d(dst, x + tile * TR * CI, TR * CI * 4, bars + st);
        tma_copy_1d(dst + TR * CI, dz + tile * TR * CO, TR * CO * 4, bars + st);
        tma_copy_1d(dst + TR * (CI + CO), y + tile * TR * CO, TR * CO * 4, bars + st);
    };
    if (tid == 0 && blockIdx.x < nfull) fetch(blockIdx.x, 0);
    float acc[CO];
#pragma unroll
    for (int o = 0; o < CO; ++o) acc[o] = 0.f;
    const int g = tid / (CI + 1), i_w = tid - g * (CI + 1);
    // x as the forward saw it: the producer block's BatchNorm (scale, shift of column i_w) applied on load
    const float xs = (in_ss && i_w < CI) ? __ldg(in_ss + i_w) : 1.f, xt = (in_ss && i_w < CI) ? __ldg(in_ss + CI + i_w) : 0.f;
    int stage = 0;
    unsigned it = 0;  // stages alternate strictly, so the phase parity of a stage is bit 1 of the iteration count
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const long long base = tile * TR;
        float *xS = ring + stage * STAGE, *dzS = xS + TR * CI, *yS = dzS + TR * CO;
        const long long nxt = tile + gridDim.x;
        if (tid == 0 && nxt < nfull) fetch(nxt, stage ^ 1);  // its previous readers finished before the last barrier
        int rows = TR;
        if (tile < nfull) {
            mbar_wait(bars + stage, (it >> 1) & 1u);
        } else {  // ragged last tile: byte counts need not be multiples of 16 -> plain loads
            rows = (int)(R - base);
            for (int e = tid; e < rows * CI; e += TR) xS[e] = __ldg(x + base * CI + e);
            for (int e = tid; e < rows * CO; e += TR) {
                dzS[e] = __ldg(dz + base * CO + e);
                yS[e] = __ldg(y + base * CO + e);
            }
            __syncthreads();
        }
        if (tid == 0 && dx) asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");  // previous dx tile has left dxS
        __syncthreads();
        if (tid < rows) {
            float d[CO], v[CO];
            load_row_s<CO>(dzS + tid * CO, d);
            load_row_s<CO>(yS + tid * CO, v);
#pragma unroll
            for (int o = 0; o < CO; ++o) d[o] = v[o] > 0.f ? fmaf(cA[o], d[o], fmaf(cB[o], v[o], cC[o])) : 0.f;
#pragma unroll
            for (int o4 = 0; o4 < COP / 4; ++o4)
                *reinterpret_cast<float4 *>(dyS + tid * DYS + 4 * o4) =
                    make_float4(d[4 * o4], 4 * o4 + 1 < CO ? d[4 * o4 + 1] : 0.f, 4 * o4 + 2 < CO ? d[4 * o4 + 2] : 0.f,
                                4 * o4 + 3 < CO ? d[4 * o4 + 3] : 0.f);
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
                if constexpr (CI % 4 == 0) {
#pragma unroll
                    for (int k4 = 0; k4 < CI / 4; ++k4)
                        *reinterpret_cast<float4 *>(dxS + tid * CI + 4 * k4) = make_float4(dxr[4 * k4], dxr[4 * k4 + 1], dxr[4 * k4 + 2], dxr[4 * k4 + 3]);
                } else if constexpr (CI % 2 == 0) {
#pragma unroll
                    for (int k2 = 0; k2 < CI / 2; ++k2)
                        *reinterpret_cast<float2 *>(dxS + tid * CI + 2 * k2) = make_float2(dxr[2 * k2], dxr[2 * k2 + 1]);
                } else {
#pragma unroll
                    for (int k = 0; k < CI; ++k) dxS[tid * CI + k] = dxr[k];
                }
            }
        }
        if (dx) asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");  // dxS writes -> visible to the TMA store
        __syncthreads();
        if (dx) {
            if (rows == TR) {
                if (tid == 0) {
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(dx + base * CI),
                                 "r"(smem_addr(dxS)), "r"(TR * CI * 4)
                                 : "memory");
                    asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
                }
            } else {
                for (int e = tid; e < rows * CI; e += TR) dx[base * CI + e] = dxS[e];
            }
        }
        if (g < G) {
            for (int r = g; r < rows; r += G) {
                const float xv = i_w < CI ? fmaf(xS[r * CI + i_w], xs, xt) : 1.f;
#pragma unroll
                for (int o4 = 0; o4 < COP / 4; ++o4) {
                    const float4 dv = *reinterpret_cast<const float4 *>(dyS + r * DYS + 4 * o4);
                    if (4 * o4 < CO) acc[4 * o4] = fmaf(dv.x, xv, acc[4 * o4]);
                    if (4 * o4 + 1 < CO) acc[4 * o4 + 1] = fmaf(dv.y, xv, acc[4 * o4 + 1]);
                    if (4 * o4 + 2 < CO) acc[4 * o4 + 2] = fmaf(dv.z, xv, acc[4 * o4 + 2]);
                    if (4 * o4 + 3 < CO) acc[4 * o4 + 3] = fmaf(dv.w, xv, acc[4 * o4 + 3]);
                }
            }
        }
        __syncthreads();  // xS / dzS / yS of this stage and dyS are free again
        stage ^= 1;
    }
    if (tid == 0 && dx) asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");
    __syncthreads();
    if (g < G) {
#pragma unroll
        for (int o = 0; o < CO; ++o) dxS[g * NP + o * (CI + 1) + i_w] = acc[o];
    }
    __syncthreads();
    for (int t = tid; t < NP; t += TR) {
        float s = 0.f;
        for (int gg = 0; gg < G; ++gg) s += dxS[gg * NP + t];
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
static int launch_lrb_fwd(const float *x, const float *in_ss, const float *W, const float *b, long long R, const int *rows_dev,
                          float *y, double *stats, cudaStream_t st)
{
    using L = LrbFwd<CI, CO>;
    auto kern = lrb_fwd_kernel<CI, CO>;
    SN2_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::SMEM), "lrb_fwd attr");
    SN2_CUDA_TRY(cudaMemsetAsync(stats, 0, sizeof(double) * 2 * CO, st), "lrb_fwd memset");
    set_count_kernel<<<1, 1, 0, st>>>(stats, CO, R, rows_dev);
    const long long nblocks = ((R + 31) / 32 + L::WARPS - 1) / L::WARPS;
    const int grid = (int)min(nblocks, (long long)148 * 2);
    kern<<<grid, LRB_T, L::SMEM, st>>>(x, in_ss, W, b, R, rows_dev, y, stats);
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
static int launch_lrb_bwd(const float *dz, const float *y, const float *x, const float *in_ss, const float *W, const float *ss,
                          const double *sums,
                          const double *stats, long long R, const int *rows_dev, float *dx, float *partial, int nblk, float *dW,
                          float *db, cudaStream_t st)
{
    using L = LrbBwd<CI, CO>;
    auto kern = lrb_bwd_kernel<CI, CO>;
    SN2_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::SMEM), "lrb_bwd attr");
    kern<<<nblk, L::TR, L::SMEM, st>>>(dz, y, x, in_ss, W, ss, sums, stats, R, rows_dev, dx, partial);
    SN2_LAUNCH_CHECK("lrb_bwd_kernel");
    lrb_wgrad_reduce_kernel<<<(L::NP * 32 + 255) / 256, 256, 0, st>>>(partial, nblk, CO, CI, dW, db);
    SN2_LAUNCH_CHECK("lrb_wgrad_reduce_kernel");
    return SN2_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// The two wide, short blocks -- SA3 [35 -> 64] and FP3 [96 -> 64] over the B * M2 sa2 points (20 000 rows at config 3)
// -- do not fit the row-per-thread kernels above (64 accumulators + 64 + 64 statistics per thread).  Same interface,
// different mapping: 32-row tiles in shared memory, thread = (row group, output channel) forward and in the
// reductions, thread = (row, input column) for dx, thread = set of (output, input) pairs for the weight gradient.
// A few tens of microseconds each; they exist so that EVERY block of the network runs the same fused
// Linear-ReLU-BatchNorm path (statistics as raw fp64 sums: one code path for SyncBatchNorm).
// ---------------------------------------------------------------------------------------------------------------
constexpr int LS_T = 256, LS_TR = 32;

template <int CI, int CO>
struct LrbSmall {
    static constexpr int G = LS_T / CO;                 // row groups
    static constexpr int RG = (LS_TR + G - 1) / G;      // rows per group and tile
    static constexpr int XS = CI + 1;                   // odd / padded row strides
    static constexpr int DS = CO + 1;
    static constexpr int NP = CO * (CI + 1);
    static constexpr int NPT = (NP + LS_T - 1) / LS_T;  // (output, input) pairs per thread
    static constexpr size_t SMEM_FWD = sizeof(float) * ((size_t)CI * CO + CO + LS_TR * XS) + sizeof(double) * 2 * CO;
    static constexpr size_t SMEM_BWD = sizeof(float) * ((size_t)CI * CO + LS_TR * XS + LS_TR * DS + 3 * CO);
    static_assert(LS_T % CO == 0, "CO must divide the CTA size");
};

template <int CI>
__device__ __forceinline__ void ls_load_x(const float *__restrict__ x, const float *__restrict__ in_ss, long long base, int rows, float *xS)
{
    for (int e = threadIdx.x; e < rows * CI; e += LS_T) {
        const int r = e / CI, k = e - r * CI;
        float v = __ldg(x + base * CI + e);
        if (in_ss) v = fmaf(v, __ldg(in_ss + k), __ldg(in_ss + CI + k));
        xS[r * (CI + 1) + k] = v;
    }
}

template <int CI, int CO>
__global__ void __launch_bounds__(LS_T)
lrb_small_fwd_kernel(const float *__restrict__ x, const float *__restrict__ in_ss, const float *__restrict__ W,
                     const float *__restrict__ b, long long R, const int *__restrict__ rows_dev, float *__restrict__ y,
                     double *__restrict__ stats)
{
    if (rows_dev) R = min(R, (long long)__ldg(rows_dev));
    using L = LrbSmall<CI, CO>;
    extern __shared__ __align__(16) unsigned char lrb_smem[];
    double *red = reinterpret_cast<double *>(lrb_smem);
    float *Wt = reinterpret_cast<float *>(red + 2 * CO);  // [CI][CO]
    float *bS = Wt + CI * CO;
    float *xS = bS + CO;
    const int tid = threadIdx.x, g = tid / CO, o = tid - g * CO;
    for (int e = tid; e < CI * CO; e += LS_T) {
        const int k = e / CO, oo = e - k * CO;
        Wt[e] = __ldg(W + oo * CI + k);
    }
    for (int e = tid; e < CO; e += LS_T) bS[e] = __ldg(b + e);
    for (int e = tid; e < 2 * CO; e += LS_T) red[e] = 0.0;
    float s1 = 0.f, s2 = 0.f;
    const long long ntiles = (R + LS_TR - 1) / LS_TR;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long base = tile * LS_TR;
        const int rows = (int)min((long long)LS_TR, R - base);
        __syncthreads();
        ls_load_x<CI>(x, in_ss, base, rows, xS);
        __syncthreads();
        float acc[L::RG];
#pragma unroll
        for (int i = 0; i < L::RG; ++i) acc[i] = bS[o];
        for (int k = 0; k < CI; ++k) {
            const float w = Wt[k * CO + o];
#pragma unroll
            for (int i = 0; i < L::RG; ++i) acc[i] = fmaf(xS[(g * L::RG + i) * L::XS + k], w, acc[i]);  // rows past `rows`: stale, unused
        }
#pragma unroll
        for (int i = 0; i < L::RG; ++i) {
            const int r = g * L::RG + i;
            if (r < rows) {
                const float v = fmaxf(acc[i], 0.f);
                y[(base + r) * CO + o] = v;
                s1 += v;
                s2 = fmaf(v, v, s2);
            }
        }
    }
    __syncthreads();
    atomicAdd(red + o, (double)s1);
    atomicAdd(red + CO + o, (double)s2);
    __syncthreads();
    for (int e = tid; e < 2 * CO; e += LS_T) atomicAdd(stats + e, red[e]);
}

template <int CO>
__global__ void __launch_bounds__(LS_T)
lrb_small_bwd_reduce_kernel(const float *__restrict__ dz, const float *__restrict__ y, long long R,
                            const int *__restrict__ rows_dev, double *__restrict__ sums)
{
    if (rows_dev) R = min(R, (long long)__ldg(rows_dev));
    __shared__ double red[2 * CO];
    const int tid = threadIdx.x, G = LS_T / CO, g = tid / CO, o = tid - g * CO;
    for (int e = tid; e < 2 * CO; e += LS_T) red[e] = 0.0;
    float a1 = 0.f, a2 = 0.f;
    for (long long r = (long long)blockIdx.x * G + g; r < R; r += (long long)gridDim.x * G) {
        const float d = __ldg(dz + r * CO + o);
        a1 += d;
        a2 = fmaf(d, __ldg(y + r * CO + o), a2);
    }
    __syncthreads();
    atomicAdd(red + o, (double)a1);
    atomicAdd(red + CO + o, (double)a2);
    __syncthreads();
    for (int e = tid; e < 2 * CO; e += LS_T) atomicAdd(sums + e, red[e]);
}

template <int CI, int CO>
__global__ void __launch_bounds__(LS_T)
lrb_small_bwd_kernel(const float *__restrict__ dz, const float *__restrict__ y, const float *__restrict__ x,
                     const float *__restrict__ in_ss, const float *__restrict__ W, const float *__restrict__ ss,
                     const double *__restrict__ sums, const double *__restrict__ stats, long long R,
                     const int *__restrict__ rows_dev, float *__restrict__ dx, float *__restrict__ partial)
{
    if (rows_dev) R = min(R, (long long)__ldg(rows_dev));
    using L = LrbSmall<CI, CO>;
    extern __shared__ __align__(16) unsigned char lrb_smem[];
    float *Ws = reinterpret_cast<float *>(lrb_smem);  // [CO][CI]
    float *xS = Ws + CO * CI;                          // [32][CI+1], column CI = 1 (bias)
    float *dyS = xS + LS_TR * L::XS;                   // [32][CO+1]
    float *cA = dyS + LS_TR * L::DS, *cB = cA + CO, *cC = cB + CO;
    const int tid = threadIdx.x, g = tid / CO, o = tid - g * CO;
    for (int e = tid; e < CO * CI; e += LS_T) Ws[e] = __ldg(W + e);
    if (tid < CO) {
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
    float acc[L::NPT];
#pragma unroll
    for (int j = 0; j < L::NPT; ++j) acc[j] = 0.f;
    const long long ntiles = (R + LS_TR - 1) / LS_TR;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long base = tile * LS_TR;
        const int rows = (int)min((long long)LS_TR, R - base);
        __syncthreads();
        ls_load_x<CI>(x, in_ss, base, rows, xS);
        if (tid < LS_TR) xS[tid * L::XS + CI] = 1.f;
#pragma unroll
        for (int i = 0; i < L::RG; ++i) {
            const int r = g * L::RG + i;
            float d = 0.f;
            if (r < rows) {
                const float v = __ldg(y + (base + r) * CO + o);
                d = v > 0.f ? fmaf(cA[o], __ldg(dz + (base + r) * CO + o), fmaf(cB[o], v, cC[o])) : 0.f;
            }
            if (r < LS_TR) dyS[r * L::DS + o] = d;
        }
        __syncthreads();
        if (dx) {
            for (int e = tid; e < rows * CI; e += LS_T) {
                const int r = e / CI, k = e - r * CI;
                float a = 0.f;
#pragma unroll 8
                for (int oo = 0; oo < CO; ++oo) a = fmaf(dyS[r * L::DS + oo], Ws[oo * CI + k], a);
                dx[base * CI + e] = a;
            }
        }
#pragma unroll
        for (int j = 0; j < L::NPT; ++j) {
            const int pr = tid + j * LS_T;
            if (pr < L::NP) {
                const int oo = pr / (CI + 1), k = pr - oo * (CI + 1);
                float a = acc[j];
                for (int r = 0; r < rows; ++r) a = fmaf(dyS[r * L::DS + oo], xS[r * L::XS + k], a);
                acc[j] = a;
            }
        }
    }
#pragma unroll
    for (int j = 0; j < L::NPT; ++j) {
        const int pr = tid + j * LS_T;
        if (pr < L::NP) partial[(size_t)blockIdx.x * L::NP + pr] = acc[j];
    }
}

template <int CI, int CO>
static int launch_lrb_small_fwd(const float *x, const float *in_ss, const float *W, const float *b, long long R, const int *rows_dev,
                                float *y, double *stats, cudaStream_t st)
{
    using L = LrbSmall<CI, CO>;
    auto kern = lrb_small_fwd_kernel<CI, CO>;
    SN2_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::SMEM_FWD), "lrb_small_fwd attr");
    SN2_CUDA_TRY(cudaMemsetAsync(stats, 0, sizeof(double) * 2 * CO, st), "lrb_small_fwd memset");
    set_count_kernel<<<1, 1, 0, st>>>(stats, CO, R, rows_dev);
    const int grid = (int)min((R + LS_TR - 1) / LS_TR, (long long)148 * 2);
    kern<<<grid, LS_T, L::SMEM_FWD, st>>>(x, in_ss, W, b, R, rows_dev, y, stats);
    SN2_LAUNCH_CHECK("lrb_small_fwd_kernel");
    return SN2_OK;
}

template <int CO>
static int launch_lrb_small_bwd_reduce(const float *dz, const float *y, long long R, const int *rows_dev, double *sums, cudaStream_t st)
{
    SN2_CUDA_TRY(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * CO, st), "lrb_small_bwd_reduce memset");
    const int G = LS_T / CO;
    const int grid = (int)min((R + G - 1) / G, (long long)148 * 2);
    lrb_small_bwd_reduce_kernel<CO><<<grid, LS_T, 0, st>>>(dz, y, R, rows_dev, sums);
    SN2_LAUNCH_CHECK("lrb_small_bwd_reduce_kernel");
    return SN2_OK;
}

template <int CI, int CO>
static int launch_lrb_small_bwd(const float *dz, const float *y, const float *x, const float *in_ss, const float *W, const float *ss,
                                const double *sums, const double *stats, long long R, const int *rows_dev, float *dx, float *partial,
                                int nblk, float *dW, float *db, cudaStream_t st)
{
    using L = LrbSmall<CI, CO>;
    auto kern = lrb_small_bwd_kernel<CI, CO>;
    SN2_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::SMEM_BWD), "lrb_small_bwd attr");
    const int grid = (int)min((R + LS_TR - 1) / LS_TR, (long long)min(nblk, 148));
    kern<<<grid, LS_T, L::SMEM_BWD, st>>>(dz, y, x, in_ss, W, ss, sums, stats, R, rows_dev, dx, partial);
    SN2_LAUNCH_CHECK("lrb_small_bwd_kernel");
    lrb_wgrad_reduce_kernel<<<(L::NP * 32 + 255) / 256, 256, 0, st>>>(partial, grid, CO, CI, dW, db);
    SN2_LAUNCH_CHECK("lrb_wgrad_reduce_kernel");
    return SN2_OK;
}

}  // namespace sn2

// (Ci, Co) pairs of the train-mode blocks: SA1 [11,16,16], SA2 [19,32], FP2 [80,34], FP1 [42,34] (row-per-thread, TMA fed)
#define SN2_LRB_SHAPES(X) X(11, 16) X(16, 16) X(19, 32) X(80, 34) X(42, 34)
// SA3 [35,64], FP3 [96,64]: the tiled kernels
#define SN2_LRB_SMALL_SHAPES(X) X(35, 64) X(96, 64)

extern "C" int sn2_lrb_supported(int Co, int Ci)
{
#define X(ci, co) if (Ci == ci && Co == co) return 1;
    SN2_LRB_SHAPES(X)
    SN2_LRB_SMALL_SHAPES(X)
#undef X
    return 0;
}

extern "C" int sn2_lrb_fwd(const float *x, const float *in_ss, const float *W, const float *b, long long R, const int *rows_dev,
                           int Co, int Ci, float *y, double *stats, void *stream)
{
    if (!x || !W || !b || !y || !stats || R <= 0) return SN2_EINVAL;
#define X(ci, co) if (Ci == ci && Co == co) return sn2::launch_lrb_small_fwd<ci, co>(x, in_ss, W, b, R, rows_dev, y, stats, (cudaStream_t)stream);
    SN2_LRB_SMALL_SHAPES(X)
#undef X
    if ((reinterpret_cast<uintptr_t>(x) & 15) != 0) return SN2_EINVAL;  // TMA bulk copies read 16-byte aligned tiles
#define X(ci, co) if (Ci == ci && Co == co) return sn2::launch_lrb_fwd<ci, co>(x, in_ss, W, b, R, rows_dev, y, stats, (cudaStream_t)stream);
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
    case 64: return sn2::launch_lrb_small_bwd_reduce<64>(dz, y, R, rows_dev, sums, (cudaStream_t)stream);
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

extern "C" int sn2_lrb_bwd(const float *dz, const float *y, const float *x, const float *in_ss, const float *W, const float *ss,
                           const double *sums, const double *stats, long long R, const int *rows_dev, int Co, int Ci, float *dx,
                           float *partial, int nblk, float *dW, float *db, void *stream)
{
    if (!dz || !y || !x || !W || !ss || !sums || !stats || !partial || !dW || !db || R <= 0 || nblk <= 0) return SN2_EINVAL;
#define X(ci, co)                  \
    if (Ci == ci && Co == co)      \
        return sn2::launch_lrb_small_bwd<ci, co>(dz, y, x, in_ss, W, ss, sums, stats, R, rows_dev, dx, partial, nblk, dW, db, (cudaStream_t)stream);
    SN2_LRB_SMALL_SHAPES(X)
#undef X
    if (((reinterpret_cast<uintptr_t>(dz) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(x) |
          reinterpret_cast<uintptr_t>(dx)) & 15) != 0)
        return SN2_EINVAL;  // TMA bulk copies move 16-byte aligned tiles
#define X(ci, co)                  \
    if (Ci == ci && Co == co)      \
        return sn2::launch_lrb_bwd<ci, co>(dz, y, x, in_ss, W, ss, sums, stats, R, rows_dev, dx, partial, nblk, dW, db, (cudaStream_t)stream);
    SN2_LRB_SHAPES(X)
#undef X
    return SN2_EUNSUPPORTED;
}

// Single-process block in one call each way (no all-reduce between the kernels): what LinReluBN uses when the
// block's BatchNorm is not a SyncBatchNorm.
extern "C" int sn2_lrb_block_fwd(const float *x, const float *in_ss, const float *W, const float *b, const float *gamma, const float *beta,
                                 float eps, float momentum, float *running_mean, float *running_var,
                                 long long *num_batches_tracked, long long R, const int *rows_dev, int Co, int Ci, float *y,
                                 double *stats, float *ss, float *z, void *stream)
{
    if (int rc = sn2_lrb_fwd(x, in_ss, W, b, R, rows_dev, Co, Ci, y, stats, stream)) return rc;
    if (int rc = sn2_bn_finalize(stats, gamma, beta, eps, momentum, running_mean, running_var, num_batches_tracked, ss, Co, stream))
        return rc;
    return z ? sn2_bn_apply(y, ss, R, rows_dev, Co, z, stream) : SN2_OK;  // z == NULL: the consumer applies ss on load
}

extern "C" int sn2_lrb_block_bwd(const float *dz, const float *y, const float *x, const float *in_ss, const float *W, const float *ss,
                                 const double *stats, long long R, const int *rows_dev, int Co, int Ci, double *sums,
                                 float *dgamma, float *dbeta, float *dx, float *partial, int nblk, float *dW, float *db,
                                 void *stream)
{
    if (int rc = sn2_lrb_bwd_reduce(dz, y, R, rows_dev, Co, sums, stream)) return rc;
    if (int rc = sn2_bn_param_grad(sums, ss, Co, dgamma, dbeta, stream)) return rc;
    return sn2_lrb_bwd(dz, y, x, in_ss, W, ss, sums, stats, R, rows_dev, Co, Ci, dx, partial, nblk, dW, db, stream);
}
