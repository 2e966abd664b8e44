// K3-K5 forward (eval-mode BatchNorm): PointConv (gather + rel-pos concat + shared MLP + max),
// global SA, the three feature-propagation MLPs and the per-point head.  SURVEY.md §8a a3-a10.
//
// fp32 SIMT formulation.  Every MLP layer is Linear -> ReLU -> BatchNorm (BN LAST,
// /root/reference/model/point_net2.py:45-53); in eval mode BN is the per-channel affine s*x + t
// with s = gamma / sqrt(running_var + eps), t = beta - running_mean * s, folded on the host.
// Weights travel as a __grid_constant__ kernel parameter: all loops over weights are fully
// unrolled, so every FFMA takes its weight as a constant-bank operand (no load instruction, no
// register, no shared-memory traffic for weights).  Activations stay in registers from the gather
// to the max / store.
#include "mlp_common.cuh"

namespace sn2 {

// ---------------------------------------------------------------------------------------------
// PointConv over a materialised neighbour list: one warp per query (centroid), one lane per edge.
// ---------------------------------------------------------------------------------------------
template <int LEVEL>
__global__ void __launch_bounds__(256)
pointconv_kernel(const float4 *__restrict__ pos, const float *__restrict__ feat, const float4 *__restrict__ qpos,
                 const int *__restrict__ rowptr, const int *__restrict__ col, int Q,
                 const __grid_constant__ typename SAEdge<LEVEL>::W W, float *__restrict__ out)
{
    using E = SAEdge<LEVEL>;
    const int lane = threadIdx.x & 31;
    const long long q = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (q >= Q) return;
    const int s = __ldg(rowptr + q), e = __ldg(rowptr + q + 1);
    const float4 qp = __ldg(qpos + q);
    float mx[E::COUT];
#pragma unroll
    for (int o = 0; o < E::COUT; ++o) mx[o] = -INFINITY;
    for (int base = s; base < e; base += 32) {
        const int j = base + lane;
        if (j < e) {
            const int p = __ldg(col + j);
            const float4 pp = __ldg(pos + p);
            float h[E::COUT];
            E::run(W, feat + (size_t)p * E::CIN, pp.x - qp.x, pp.y - qp.y, pp.z - qp.z, h);
#pragma unroll
            for (int o = 0; o < E::COUT; ++o) mx[o] = fmaxf(mx[o], h[o]);
        }
    }
#pragma unroll
    for (int o = 0; o < E::COUT; ++o) mx[o] = (e > s) ? warp_max(mx[o]) : 0.f;  // torch_scatter: empty row -> 0
    if (lane == 0) {
        float4 *o4 = reinterpret_cast<float4 *>(out + (size_t)q * E::COUT);
#pragma unroll
        for (int v = 0; v < E::COUT / 4; ++v) o4[v] = make_float4(mx[4 * v], mx[4 * v + 1], mx[4 * v + 2], mx[4 * v + 3]);
    }
}

// ---------------------------------------------------------------------------------------------
// Global SA: MLP[35,64] on [x2, pos2] then per-plot max.  One CTA per plot.
// ---------------------------------------------------------------------------------------------
constexpr int GSA_THREADS = 256;
__global__ void __launch_bounds__(GSA_THREADS)
global_sa_kernel(const float *__restrict__ x2, const float4 *__restrict__ pos, int M,
                 const __grid_constant__ W_SA3 W, float *__restrict__ g)
{
    __shared__ float red[GSA_THREADS / 32][SN2_C3];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float mx[SN2_C3];
#pragma unroll
    for (int o = 0; o < SN2_C3; ++o) mx[o] = -INFINITY;
    for (int i = tid; i < M; i += GSA_THREADS) {
        const size_t row = (size_t)b * M + i;
        float h[SN2_C3];
        acc_init(W.l1, h);
#pragma unroll
        for (int v = 0; v < SN2_C2 / 4; ++v) {
            const float4 f = ldg4(x2 + row * SN2_C2 + 4 * v);
            acc_step<0>(W.l1, f.x, h, 4 * v);
            acc_step<0>(W.l1, f.y, h, 4 * v + 1);
            acc_step<0>(W.l1, f.z, h, 4 * v + 2);
            acc_step<0>(W.l1, f.w, h, 4 * v + 3);
        }
        const float4 pp = __ldg(pos + row);
        acc_step<SN2_C2>(W.l1, pp.x, h, 0);
        acc_step<SN2_C2>(W.l1, pp.y, h, 1);
        acc_step<SN2_C2>(W.l1, pp.z, h, 2);
        relu_bn(W.l1, h);
#pragma unroll
        for (int o = 0; o < SN2_C3; ++o) mx[o] = fmaxf(mx[o], h[o]);
    }
#pragma unroll
    for (int o = 0; o < SN2_C3; ++o) {
        float m = warp_max(mx[o]);
        if (lane == 0) red[warp][o] = m;
    }
    __syncthreads();
    if (tid < SN2_C3) {
        float m = red[0][tid];
#pragma unroll
        for (int w = 1; w < GSA_THREADS / 32; ++w) m = fmaxf(m, red[w][tid]);
        g[(size_t)b * SN2_C3 + tid] = m;
    }
}

// ---------------------------------------------------------------------------------------------
// FP3: k=1 interpolation of the plot vector (y = (g*w)/w with w = 1/max(|pos|^2, 1e-16), the exact
// arithmetic of knn_interpolate with a single source at the origin) ++ x2 -> MLP[96,64].
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
fp3_kernel(const float *__restrict__ g, const float *__restrict__ x2, const float4 *__restrict__ pos, int B, int M,
           const __grid_constant__ W_FP3 W, float *__restrict__ out)
{
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= (long long)B * M) return;
    const int b = (int)(row / M);
    const float4 pp = __ldg(pos + row);
    const float d2 = dist2(0.f, 0.f, 0.f, pp.x, pp.y, pp.z);
    const float w = __fdiv_rn(1.0f, fmaxf(d2, 1e-16f));
    float h[SN2_C3];
    acc_init(W.l1, h);
    const float *gb = g + (size_t)b * SN2_C3;
#pragma unroll
    for (int v = 0; v < SN2_C3 / 4; ++v) {
        const float4 f = ldg4(gb + 4 * v);
        acc_step<0>(W.l1, __fdiv_rn(__fmul_rn(f.x, w), w), h, 4 * v);
        acc_step<0>(W.l1, __fdiv_rn(__fmul_rn(f.y, w), w), h, 4 * v + 1);
        acc_step<0>(W.l1, __fdiv_rn(__fmul_rn(f.z, w), w), h, 4 * v + 2);
        acc_step<0>(W.l1, __fdiv_rn(__fmul_rn(f.w, w), w), h, 4 * v + 3);
    }
#pragma unroll
    for (int v = 0; v < SN2_C2 / 4; ++v) {
        const float4 f = ldg4(x2 + (size_t)row * SN2_C2 + 4 * v);
        acc_step<SN2_C3>(W.l1, f.x, h, 4 * v);
        acc_step<SN2_C3>(W.l1, f.y, h, 4 * v + 1);
        acc_step<SN2_C3>(W.l1, f.z, h, 4 * v + 2);
        acc_step<SN2_C3>(W.l1, f.w, h, 4 * v + 3);
    }
    relu_bn(W.l1, h);
    float4 *o4 = reinterpret_cast<float4 *>(out + (size_t)row * SN2_C3);
#pragma unroll
    for (int v = 0; v < SN2_C3 / 4; ++v) o4[v] = make_float4(h[4 * v], h[4 * v + 1], h[4 * v + 2], h[4 * v + 3]);
}

// interpolation of one float4 column group: ((w0*a0 + w1*a1) + w2*a2) / ((w0 + w1) + w2), each
// operation rounded separately, in the oracle's accumulation order.
__device__ __forceinline__ float interp1(float a0, float a1, float a2, float w0, float w1, float w2, float den)
{
    float num = __fadd_rn(__fadd_rn(__fmul_rn(a0, w0), __fmul_rn(a1, w1)), __fmul_rn(a2, w2));
    return __fdiv_rn(num, den);
}

// ---------------------------------------------------------------------------------------------
// FP2: interpolate f3 (64 ch) ++ x1 (16 ch) -> MLP[80,34].  One thread per sa1 point.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
fp2_kernel(const float *__restrict__ f3, const int *__restrict__ nbr, const float *__restrict__ wgt,
           const float *__restrict__ x1, int Q, const __grid_constant__ W_FP2 W, float *__restrict__ out)
{
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= Q) return;
    const int i0 = __ldg(nbr + 3 * row), i1 = __ldg(nbr + 3 * row + 1), i2 = __ldg(nbr + 3 * row + 2);
    const float w0 = __ldg(wgt + 3 * row), w1 = __ldg(wgt + 3 * row + 1), w2 = __ldg(wgt + 3 * row + 2);
    const float den = __fadd_rn(__fadd_rn(w0, w1), w2);
    const float *a0 = f3 + (size_t)i0 * SN2_C3, *a1 = f3 + (size_t)i1 * SN2_C3, *a2 = f3 + (size_t)i2 * SN2_C3;
    float h[SN2_CF];
    acc_init(W.l1, h);
#pragma unroll
    for (int v = 0; v < SN2_C3 / 4; ++v) {
        const float4 p0 = ldg4(a0 + 4 * v), p1 = ldg4(a1 + 4 * v), p2 = ldg4(a2 + 4 * v);
        acc_step<0>(W.l1, interp1(p0.x, p1.x, p2.x, w0, w1, w2, den), h, 4 * v);
        acc_step<0>(W.l1, interp1(p0.y, p1.y, p2.y, w0, w1, w2, den), h, 4 * v + 1);
        acc_step<0>(W.l1, interp1(p0.z, p1.z, p2.z, w0, w1, w2, den), h, 4 * v + 2);
        acc_step<0>(W.l1, interp1(p0.w, p1.w, p2.w, w0, w1, w2, den), h, 4 * v + 3);
    }
#pragma unroll
    for (int v = 0; v < SN2_C1 / 4; ++v) {
        const float4 f = ldg4(x1 + (size_t)row * SN2_C1 + 4 * v);
        acc_step<SN2_C3>(W.l1, f.x, h, 4 * v);
        acc_step<SN2_C3>(W.l1, f.y, h, 4 * v + 1);
        acc_step<SN2_C3>(W.l1, f.z, h, 4 * v + 2);
        acc_step<SN2_C3>(W.l1, f.w, h, 4 * v + 3);
    }
    relu_bn(W.l1, h);
    float4 *o4 = reinterpret_cast<float4 *>(out + (size_t)row * SN2_CF_LD);
#pragma unroll
    for (int v = 0; v < 8; ++v) o4[v] = make_float4(h[4 * v], h[4 * v + 1], h[4 * v + 2], h[4 * v + 3]);
    o4[8] = make_float4(h[32], h[33], 0.f, 0.f);
}

// ---------------------------------------------------------------------------------------------
// FP1 + head: interpolate f2 (34 ch) ++ raw features (8) -> MLP[42,34] -> lin1+ReLU -> lin2 ->
// softmax over 4 class logits, sigmoid density, coverages = proba * density.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
fp1_head_kernel(const float *__restrict__ f2, const int *__restrict__ nbr, const float *__restrict__ wgt,
                const float *__restrict__ feat, int Q, const __grid_constant__ W_FP1 W, float4 *__restrict__ cov,
                float4 *__restrict__ proba)
{
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= Q) return;
    const int i0 = __ldg(nbr + 3 * row), i1 = __ldg(nbr + 3 * row + 1), i2 = __ldg(nbr + 3 * row + 2);
    const float w0 = __ldg(wgt + 3 * row), w1 = __ldg(wgt + 3 * row + 1), w2 = __ldg(wgt + 3 * row + 2);
    const float den = __fadd_rn(__fadd_rn(w0, w1), w2);
    const float *a0 = f2 + (size_t)i0 * SN2_CF_LD, *a1 = f2 + (size_t)i1 * SN2_CF_LD, *a2 = f2 + (size_t)i2 * SN2_CF_LD;
    float h[SN2_CF];
    acc_init(W.l1, h);
#pragma unroll
    for (int v = 0; v < 8; ++v) {
        const float4 p0 = ldg4(a0 + 4 * v), p1 = ldg4(a1 + 4 * v), p2 = ldg4(a2 + 4 * v);
        acc_step<0>(W.l1, interp1(p0.x, p1.x, p2.x, w0, w1, w2, den), h, 4 * v);
        acc_step<0>(W.l1, interp1(p0.y, p1.y, p2.y, w0, w1, w2, den), h, 4 * v + 1);
        acc_step<0>(W.l1, interp1(p0.z, p1.z, p2.z, w0, w1, w2, den), h, 4 * v + 2);
        acc_step<0>(W.l1, interp1(p0.w, p1.w, p2.w, w0, w1, w2, den), h, 4 * v + 3);
    }
    {
        const float4 p0 = ldg4(a0 + 32), p1 = ldg4(a1 + 32), p2 = ldg4(a2 + 32);
        acc_step<0>(W.l1, interp1(p0.x, p1.x, p2.x, w0, w1, w2, den), h, 32);
        acc_step<0>(W.l1, interp1(p0.y, p1.y, p2.y, w0, w1, w2, den), h, 33);
    }
#pragma unroll
    for (int v = 0; v < SN2_F0 / 4; ++v) {
        const float4 f = ldg4(feat + (size_t)row * SN2_F0 + 4 * v);
        acc_step<SN2_CF>(W.l1, f.x, h, 4 * v);
        acc_step<SN2_CF>(W.l1, f.y, h, 4 * v + 1);
        acc_step<SN2_CF>(W.l1, f.z, h, 4 * v + 2);
        acc_step<SN2_CF>(W.l1, f.w, h, 4 * v + 3);
    }
    relu_bn(W.l1, h);
    float u[16];
    acc_init(W.lin1, u);
#pragma unroll
    for (int k = 0; k < SN2_CF; ++k) acc_step<0>(W.lin1, h[k], u, k);
#pragma unroll
    for (int o = 0; o < 16; ++o) u[o] = fmaxf(u[o], 0.f);
    float sc[5];
    acc_init(W.lin2, sc);
#pragma unroll
    for (int k = 0; k < 16; ++k) acc_step<0>(W.lin2, u[k], sc, k);
    const float m = fmaxf(fmaxf(sc[0], sc[1]), fmaxf(sc[2], sc[3]));
    const float e0 = expf(sc[0] - m), e1 = expf(sc[1] - m), e2 = expf(sc[2] - m), e3 = expf(sc[3] - m);
    const float sum = (e0 + e1) + (e2 + e3);
    const float4 pr = make_float4(e0 / sum, e1 / sum, e2 / sum, e3 / sum);
    const float dens = 1.0f / (1.0f + expf(-sc[4]));
    proba[row] = pr;
    cov[row] = make_float4(pr.x * dens, pr.y * dens, pr.z * dens, pr.w * dens);
}

// ---------------------------------------------------------------------------------------------
// knn3: brute-force 3 nearest sources per query within the plot, sources staged in shared memory.
// Candidates are visited in ascending index with strict '<' insertion => ties keep the lower index.
// ---------------------------------------------------------------------------------------------
constexpr int KNN_THREADS = 256;
constexpr int KNN_TILE = 2048;
__global__ void __launch_bounds__(KNN_THREADS)
knn3_kernel(const float4 *__restrict__ spos, const float4 *__restrict__ qpos, int Ms, int Nq, int *__restrict__ nbr,
            float *__restrict__ wgt)
{
    __shared__ float4 tile[KNN_TILE];
    const int b = blockIdx.y;
    const int qi = blockIdx.x * KNN_THREADS + threadIdx.x;
    const bool active = qi < Nq;
    const float4 q = active ? __ldg(qpos + (size_t)b * Nq + qi) : make_float4(0.f, 0.f, 0.f, 0.f);
    float d0 = INFINITY, d1 = INFINITY, d2 = INFINITY;
    int i0 = -1, i1 = -1, i2 = -1;
    const float4 *sp = spos + (size_t)b * Ms;
    for (int base = 0; base < Ms; base += KNN_TILE) {
        const int n = min(KNN_TILE, Ms - base);
        __syncthreads();
        for (int t = threadIdx.x; t < n; t += KNN_THREADS) tile[t] = __ldg(sp + base + t);
        __syncthreads();
        if (active) {
#pragma unroll 4
            for (int t = 0; t < n; ++t) {
                const float4 s = tile[t];
                // diff = source - query (PyG knn_interpolate forms pos_x[x_idx] - pos_y[y_idx])
                const float d = dist2(s.x, s.y, s.z, q.x, q.y, q.z);
                if (d < d2 || i2 < 0) {
                    const int id = base + t;
                    if (d < d1 || i1 < 0) {
                        d2 = d1; i2 = i1;
                        if (d < d0 || i0 < 0) { d1 = d0; i1 = i0; d0 = d; i0 = id; }
                        else { d1 = d; i1 = id; }
                    } else { d2 = d; i2 = id; }
                }
            }
        }
    }
    if (active) {
        const size_t o = ((size_t)b * Nq + qi) * 3;
        const int gb = b * Ms;
        nbr[o] = gb + i0; nbr[o + 1] = gb + i1; nbr[o + 2] = gb + i2;
        wgt[o] = __fdiv_rn(1.0f, fmaxf(d0, 1e-16f));
        wgt[o + 1] = __fdiv_rn(1.0f, fmaxf(d1, 1e-16f));
        wgt[o + 2] = __fdiv_rn(1.0f, fmaxf(d2, 1e-16f));
    }
}

}  // namespace sn2

using namespace sn2;

extern "C" int sn2_pointconv_fwd(int level, const float *pos4, const float *feat, const float *qpos4,
                                 const int *rowptr, const int *col, int Q, const float *w_host, int nw, float *out,
                                 void *stream)
{
    if (!pos4 || !feat || !qpos4 || !rowptr || !col || !out || Q <= 0) return SN2_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    long long blocks = ((long long)Q * 32 + 255) / 256;
    if (level == 1) {
        W_SA1 w;
        if (int rc = load_weights(w, w_host, nw)) return rc;
        pointconv_kernel<1><<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const float4 *>(pos4), feat,
                                                              reinterpret_cast<const float4 *>(qpos4), rowptr, col, Q, w, out);
    } else if (level == 2) {
        W_SA2 w;
        if (int rc = load_weights(w, w_host, nw)) return rc;
        pointconv_kernel<2><<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const float4 *>(pos4), feat,
                                                              reinterpret_cast<const float4 *>(qpos4), rowptr, col, Q, w, out);
    } else {
        return SN2_EINVAL;
    }
    SN2_LAUNCH_CHECK("pointconv_kernel");
    return SN2_OK;
}

extern "C" int sn2_global_sa_fwd(const float *x2, const float *pos4, int B, int M, const float *w_host, int nw,
                                 float *g, void *stream)
{
    if (!x2 || !pos4 || !g || B <= 0 || M <= 0) return SN2_EINVAL;
    W_SA3 w;
    if (int rc = load_weights(w, w_host, nw)) return rc;
    global_sa_kernel<<<B, GSA_THREADS, 0, (cudaStream_t)stream>>>(x2, reinterpret_cast<const float4 *>(pos4), M, w, g);
    SN2_LAUNCH_CHECK("global_sa_kernel");
    return SN2_OK;
}

extern "C" int sn2_fp3_fwd(const float *g, const float *x2, const float *pos4, int B, int M, const float *w_host,
                           int nw, float *out, void *stream)
{
    if (!g || !x2 || !pos4 || !out || B <= 0 || M <= 0) return SN2_EINVAL;
    W_FP3 w;
    if (int rc = load_weights(w, w_host, nw)) return rc;
    long long rows = (long long)B * M;
    fp3_kernel<<<(unsigned)((rows + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        g, x2, reinterpret_cast<const float4 *>(pos4), B, M, w, out);
    SN2_LAUNCH_CHECK("fp3_kernel");
    return SN2_OK;
}

extern "C" int sn2_knn3(const float *spos4, const float *qpos4, int B, int Ms, int Nq, int *nbr, float *w,
                        void *stream)
{
    if (!spos4 || !qpos4 || !nbr || !w || B <= 0 || Ms < 3 || Nq <= 0) return SN2_EINVAL;
    dim3 grid((Nq + KNN_THREADS - 1) / KNN_THREADS, B);
    knn3_kernel<<<grid, KNN_THREADS, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4 *>(spos4),
                                                                reinterpret_cast<const float4 *>(qpos4), Ms, Nq, nbr, w);
    SN2_LAUNCH_CHECK("knn3_kernel");
    return SN2_OK;
}

extern "C" int sn2_fp2_fwd(const float *f3, const int *nbr, const float *w, const float *x1, int Q,
                           const float *w_host, int nw, float *out, void *stream)
{
    if (!f3 || !nbr || !w || !x1 || !out || Q <= 0) return SN2_EINVAL;
    W_FP2 ws;
    if (int rc = load_weights(ws, w_host, nw)) return rc;
    fp2_kernel<<<(unsigned)(((long long)Q + 127) / 128), 128, 0, (cudaStream_t)stream>>>(f3, nbr, w, x1, Q, ws, out);
    SN2_LAUNCH_CHECK("fp2_kernel");
    return SN2_OK;
}

extern "C" int sn2_fp1_head_fwd(const float *f2, const int *nbr, const float *w, const float *feat, int Q,
                                const float *w_host, int nw, float *cov, float *proba, void *stream)
{
    if (!f2 || !nbr || !w || !feat || !cov || !proba || Q <= 0) return SN2_EINVAL;
    W_FP1 ws;
    if (int rc = load_weights(ws, w_host, nw)) return rc;
    fp1_head_kernel<<<(unsigned)(((long long)Q + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        f2, nbr, w, feat, Q, ws, reinterpret_cast<float4 *>(cov), reinterpret_cast<float4 *>(proba));
    SN2_LAUNCH_CHECK("fp1_head_kernel");
    return SN2_OK;
}
