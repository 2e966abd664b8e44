// Training-path graph operators with their backward passes.
//
// In training mode the MLP blocks run with batch statistics over ALL edge messages of the batch
// (reference model/point_net2.py:45-53 inside PointConv), so the message matrix is
// materialised like the reference does and the Linear/BatchNorm arithmetic stays in torch (cuBLAS);
// what this file replaces is every torch_geometric / torch_scatter operator around it:
//   edge_msg       x_j ++ (pos_j - pos_i) per edge of the CSR neighbour list        (PointConv.message)
//   segment_max    per-row max over CSR rows + first-edge arg-max, arg-routed bwd   (scatter max, global_max_pool)
//   interp3        k=3 inverse-squared-distance interpolation, scatter-add bwd      (knn_interpolate)
//   interp_plot    k=1 broadcast of the plot vector (w*x)/w, per-plot sum bwd        (knn_interpolate, fp3)
//   project_plotwise_bwd   arg-routed gradient of the occupied-pixel mean            (scatter_max + scatter_mean)
#include "sn2_common.cuh"

namespace sn2 {

// ---- edge messages ---------------------------------------------------------------------------
// One warp per query row; the lanes walk the row's message block msg[s*(C+3) .. e*(C+3)) element by element, so
// every store instruction writes 128 contiguous bytes (a lane-per-edge mapping writes 32 rows 4 bytes at a time).
// Element (edge j, column c): c < C -> x[col[j]][c], else pos[col[j]] - qpos[q] component c - C.
template <int C>
__global__ void __launch_bounds__(256)
edge_msg_fwd_kernel(const float *__restrict__ x, const float4 *__restrict__ pos, const float4 *__restrict__ qpos,
                    const int *__restrict__ rowptr, const int *__restrict__ col, int Q, float *__restrict__ msg)
{
    constexpr int LD = C + 3;
    const int lane = threadIdx.x & 31;
    const long long q = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (q >= Q) return;
    const int s = __ldg(rowptr + q), e = __ldg(rowptr + q + 1);
    const float4 qp = __ldg(qpos + q);
    const int n = (e - s) * LD;
    float *m = msg + (size_t)s * LD;
    for (int t = lane; t < n; t += 32) {
        const int jl = t / LD, c = t - jl * LD;
        const int p = __ldg(col + s + jl);
        float v;
        if (c < C) {
            v = __ldg(x + (size_t)p * C + c);
        } else {
            const float4 pp = __ldg(pos + p);
            v = c == C ? pp.x - qp.x : (c == C + 1 ? pp.y - qp.y : pp.z - qp.z);
        }
        m[t] = v;
    }
}

// any other width: lane per edge
__global__ void __launch_bounds__(256)
edge_msg_fwd_generic_kernel(const float *__restrict__ x, const float4 *__restrict__ pos, const float4 *__restrict__ qpos,
                            const int *__restrict__ rowptr, const int *__restrict__ col, int Q, int C, float *__restrict__ msg)
{
    const int lane = threadIdx.x & 31;
    const long long q = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (q >= Q) return;
    const int s = __ldg(rowptr + q), e = __ldg(rowptr + q + 1);
    const float4 qp = __ldg(qpos + q);
    const int ld = C + 3;
    for (int j = s + lane; j < e; j += 32) {
        const int p = __ldg(col + j);
        float *m = msg + (size_t)j * ld;
        const float *xr = x + (size_t)p * C;
        for (int c = 0; c < C; ++c) m[c] = __ldg(xr + c);
        const float4 pp = __ldg(pos + p);
        m[C] = pp.x - qp.x;
        m[C + 1] = pp.y - qp.y;
        m[C + 2] = pp.z - qp.z;
    }
}

// dx[col[e]] += dmsg[e][0:C]   (positions are inputs: no gradient)
__global__ void __launch_bounds__(256)
edge_msg_bwd_kernel(const float *__restrict__ dmsg, const int *__restrict__ col, long long E,
                    const int *__restrict__ rows_dev, int C, float *__restrict__ dx)
{
    if (rows_dev) E = min(E, (long long)__ldg(rows_dev));
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= E * C) return;
    const long long e = t / C;
    const int c = (int)(t - e * C);
    atomicAdd(dx + (size_t)__ldg(col + e) * C + c, __ldg(dmsg + (size_t)e * (C + 3) + c));
}

// ---- segment max with first-edge arg-max -------------------------------------------------------
// Lane = (edge slot, channel quad): C/4 lanes cover one edge row with float4 loads (coalesced: a warp reads 32/(C/4)
// consecutive rows = 512 contiguous bytes per instruction); every lane keeps the best (value, edge) of its four
// channels over the edges of its slot, visited in ascending order with strict '>' (first edge wins inside a lane);
// slots are merged with xor shuffles on (value, edge): larger value, then lower edge index.  WARPS = 1: one warp per
// row (set-abstraction neighbourhoods, tens of edges); WARPS = 8: one CTA per row (global_max_pool: 32 rows of
// hundreds of points), warps merged through shared memory.
// NaN is the largest value (torch.max / torch_scatter propagate it); equal values -> lower edge index.  `e` always
// stays a valid edge of the row, also when every value is -inf or NaN (a diverged step must give a NaN loss, not an
// out-of-bounds write in the backward).
__device__ __forceinline__ bool argmax_better(float ov, int oe, float v, int e)
{
    const bool on = ov != ov, vn = v != v;
    if (on != vn) return on;
    return ov > v || ((ov == v || on) && oe < e);
}
__device__ __forceinline__ void argmax_merge(float &v, int &e, float ov, int oe)
{
    if (argmax_better(ov, oe, v, e)) { v = ov; e = oe; }
}

template <int C, int WARPS>
__global__ void __launch_bounds__(WARPS == 1 ? 256 : 32 * WARPS)
segment_max_fwd_kernel(const float *__restrict__ vals, const float *__restrict__ ss, const int *__restrict__ rowptr, int Q,
                       float *__restrict__ out, int *__restrict__ arg)
{
    constexpr int CQ = C / 4, SLOTS = 32 / CQ;  // lanes per edge row, edge rows per warp and step
    static_assert(C % 4 == 0 && 32 % CQ == 0, "C must be 16, 32, 64 or 128");
    __shared__ float sv[WARPS > 1 ? WARPS * C : 1];
    __shared__ int se[WARPS > 1 ? WARPS * C : 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long q = WARPS == 1 ? (((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5) : (long long)blockIdx.x;
    if (q >= Q) return;
    const int s = __ldg(rowptr + q), e = __ldg(rowptr + q + 1);
    const int slot = lane / CQ, cq = lane - slot * CQ;
    float bv[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    int be[4] = {0x7fffffff, 0x7fffffff, 0x7fffffff, 0x7fffffff};  // no edge seen yet (loses every merge)
    // optional per-channel affine (the producer block's BatchNorm scale | shift) applied on load: with a negative
    // scale the max of the transformed values is not the transform of the max, so it cannot wait until afterwards
    float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ss) {
        sc = __ldg(reinterpret_cast<const float4 *>(ss) + cq);
        sh = __ldg(reinterpret_cast<const float4 *>(ss + C) + cq);
    }
    const int first = s + (WARPS == 1 ? 0 : warp * SLOTS) + slot, step = SLOTS * WARPS;
    for (int j = first; j < e; j += step) {
        float4 t = __ldg(reinterpret_cast<const float4 *>(vals + (size_t)j * C) + cq);
        t = make_float4(fmaf(t.x, sc.x, sh.x), fmaf(t.y, sc.y, sh.y), fmaf(t.z, sc.z, sh.z), fmaf(t.w, sc.w, sh.w));
        // ascending j inside a lane: strict '>' keeps the first edge; the first edge of the slot always enters
        if (be[0] == 0x7fffffff || t.x > bv[0] || (t.x != t.x && bv[0] == bv[0])) { bv[0] = t.x; be[0] = j; }
        if (be[1] == 0x7fffffff || t.y > bv[1] || (t.y != t.y && bv[1] == bv[1])) { bv[1] = t.y; be[1] = j; }
        if (be[2] == 0x7fffffff || t.z > bv[2] || (t.z != t.z && bv[2] == bv[2])) { bv[2] = t.z; be[2] = j; }
        if (be[3] == 0x7fffffff || t.w > bv[3] || (t.w != t.w && bv[3] == bv[3])) { bv[3] = t.w; be[3] = j; }
    }
#pragma unroll
    for (int d = CQ; d < 32; d <<= 1) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float ov = __shfl_xor_sync(SN2_FULL, bv[c], d);
            const int oe = __shfl_xor_sync(SN2_FULL, be[c], d);
            argmax_merge(bv[c], be[c], ov, oe);
        }
    }
    if constexpr (WARPS > 1) {
        if (slot == 0) {
#pragma unroll
            for (int c = 0; c < 4; ++c) { sv[warp * C + 4 * cq + c] = bv[c]; se[warp * C + 4 * cq + c] = be[c]; }
        }
        __syncthreads();
        if (threadIdx.x < C) {
            float v = sv[threadIdx.x];
            int ee = se[threadIdx.x];
            for (int w = 1; w < WARPS; ++w) argmax_merge(v, ee, sv[w * C + threadIdx.x], se[w * C + threadIdx.x]);
            // torch_scatter: empty row -> value 0, arg = number of edges (we store -1)
            out[(size_t)q * C + threadIdx.x] = e > s ? v : 0.f;
            arg[(size_t)q * C + threadIdx.x] = e > s ? ee : -1;
        }
    } else if (slot == 0) {
        const bool any = e > s;
        *reinterpret_cast<float4 *>(out + (size_t)q * C + 4 * cq) =
            any ? make_float4(bv[0], bv[1], bv[2], bv[3]) : make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<int4 *>(arg + (size_t)q * C + 4 * cq) = any ? make_int4(be[0], be[1], be[2], be[3]) : make_int4(-1, -1, -1, -1);
    }
}

__global__ void __launch_bounds__(256)
segment_max_bwd_kernel(const float *__restrict__ dout, const int *__restrict__ arg, long long QC, int C, long long E,
                       float *__restrict__ dvals)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= QC) return;
    const int a = __ldg(arg + t);
    if (a >= 0 && a < E) dvals[(size_t)a * C + (int)(t % C)] = __ldg(dout + t);  // rows are disjoint: no conflict
}

// ---- k=3 interpolation ---------------------------------------------------------------------------
__device__ __forceinline__ float interp_rn(float a0, float a1, float a2, float w0, float w1, float w2, float den)
{
    return __fdiv_rn(__fadd_rn(__fadd_rn(__fmul_rn(a0, w0), __fmul_rn(a1, w1)), __fmul_rn(a2, w2)), den);
}

__global__ void __launch_bounds__(256)
interp3_fwd_kernel(const float *__restrict__ x, int ldx, const int *__restrict__ nbr, const float *__restrict__ w,
                   long long Q, int C, float *__restrict__ y)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= Q * C) return;
    const long long q = t / C;
    const int c = (int)(t - q * C);
    const int i0 = __ldg(nbr + 3 * q), i1 = __ldg(nbr + 3 * q + 1), i2 = __ldg(nbr + 3 * q + 2);
    const float w0 = __ldg(w + 3 * q), w1 = __ldg(w + 3 * q + 1), w2 = __ldg(w + 3 * q + 2);
    const float den = __fadd_rn(__fadd_rn(w0, w1), w2);
    y[t] = interp_rn(__ldg(x + (size_t)i0 * ldx + c), __ldg(x + (size_t)i1 * ldx + c), __ldg(x + (size_t)i2 * ldx + c), w0,
                     w1, w2, den);
}

// dx[nbr_k] += (dy / den) * w_k   (the oracle's autograd: y = num / den, num = sum w*x)
__global__ void __launch_bounds__(256)
interp3_bwd_kernel(const float *__restrict__ dy, const int *__restrict__ nbr, const float *__restrict__ w,
                   long long Q, int C, float *__restrict__ dx)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= Q * C) return;
    const long long q = t / C;
    const int c = (int)(t - q * C);
    const float w0 = __ldg(w + 3 * q), w1 = __ldg(w + 3 * q + 1), w2 = __ldg(w + 3 * q + 2);
    const float g = __fdiv_rn(__ldg(dy + t), __fadd_rn(__fadd_rn(w0, w1), w2));
    atomicAdd(dx + (size_t)__ldg(nbr + 3 * q) * C + c, g * w0);
    atomicAdd(dx + (size_t)__ldg(nbr + 3 * q + 1) * C + c, g * w1);
    atomicAdd(dx + (size_t)__ldg(nbr + 3 * q + 2) * C + c, g * w2);
}

// ---- k=1 interpolation of the plot vector (fp3) ----------------------------------------------------
__global__ void __launch_bounds__(256)
interp_plot_fwd_kernel(const float *__restrict__ g, const float4 *__restrict__ pos, long long Q, int M, int C,
                       float *__restrict__ y)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= Q * C) return;
    const long long q = t / C;
    const int c = (int)(t - q * C);
    const float4 p = __ldg(pos + q);
    const float w = __fdiv_rn(1.0f, fmaxf(dist2(0.f, 0.f, 0.f, p.x, p.y, p.z), 1e-16f));
    y[t] = __fdiv_rn(__fmul_rn(__ldg(g + (size_t)(q / M) * C + c), w), w);
}

// dg[b][c] = sum over the plot's points of (dy / w) * w ; one CTA per (plot, channel), fixed-order tree
__global__ void __launch_bounds__(256)
interp_plot_bwd_kernel(const float *__restrict__ dy, const float4 *__restrict__ pos, int M, int C,
                       float *__restrict__ dg)
{
    __shared__ float red[8];
    const int b = blockIdx.x, c = blockIdx.y, tid = threadIdx.x;
    float s = 0.f;
    for (int i = tid; i < M; i += 256) {
        const size_t q = (size_t)b * M + i;
        const float4 p = __ldg(pos + q);
        const float w = __fdiv_rn(1.0f, fmaxf(dist2(0.f, 0.f, 0.f, p.x, p.y, p.z), 1e-16f));
        s += __fmul_rn(__fdiv_rn(__ldg(dy + q * C + c), w), w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(SN2_FULL, s, o);
    if ((tid & 31) == 0) red[tid >> 5] = s;
    __syncthreads();
    if (tid == 0) {
        float t = 0.f;
        for (int k = 0; k < 8; ++k) t += red[k];
        dg[(size_t)b * C + c] = t;
    }
}

// ---- plot-wise coverage backward --------------------------------------------------------------------
// out[b] = [mean(low), mean(1-low), mean(med), mean(high)] over occupied pixels; grads go to the per-pixel
// arg-max point of channels 0, 2, 3 (parg from the forward).  dpred must be zero-initialised.
__global__ void __launch_bounds__(256)
project_plotwise_bwd_kernel(const float *__restrict__ dout, const int *__restrict__ parg, int D,
                            float *__restrict__ dpred)
{
    __shared__ int s_cnt;
    const int b = blockIdx.x, tid = threadIdx.x, P = (D + 1) * (D + 1);  // the forward's pixel frame
    const int *pa = parg + (size_t)b * 3 * P;
    if (tid == 0) s_cnt = 0;
    __syncthreads();
    int c = 0;
    for (int p = tid; p < P; p += 256) c += pa[p] >= 0;
    atomicAdd(&s_cnt, c);
    __syncthreads();
    const float inv = 1.0f / fmaxf((float)s_cnt, 1.f);
    const float g_low = (dout[b * 4 + 0] - dout[b * 4 + 1]) * inv;  // bare = 1 - low
    const float g_med = dout[b * 4 + 2] * inv, g_high = dout[b * 4 + 3] * inv;
    for (int p = tid; p < P; p += 256) {
        const int a0 = pa[p];
        if (a0 < 0) continue;
        // distinct pixels can share an arg-max point only across bands, which hit different columns
        atomicAdd(dpred + (size_t)a0 * 4 + 0, g_low);
        atomicAdd(dpred + (size_t)pa[P + p] * 4 + 2, g_med);
        atomicAdd(dpred + (size_t)pa[2 * P + p] * 4 + 3, g_high);
    }
}

static inline unsigned blocks_for(long long n, int t) { return (unsigned)((n + t - 1) / t); }

}  // namespace sn2

using namespace sn2;

extern "C" int sn2_edge_msg_fwd(const float *x, const float *pos4, const float *qpos4, const int *rowptr,
                                const int *col, int Q, int C, float *msg, void *stream)
{
    if (!x || !pos4 || !qpos4 || !rowptr || !col || !msg || Q <= 0 || C <= 0) return SN2_EINVAL;
    const unsigned blocks = blocks_for((long long)Q * 32, 256);
    const float4 *p4 = reinterpret_cast<const float4 *>(pos4), *q4 = reinterpret_cast<const float4 *>(qpos4);
    cudaStream_t st = (cudaStream_t)stream;
    if (C == 8) edge_msg_fwd_kernel<8><<<blocks, 256, 0, st>>>(x, p4, q4, rowptr, col, Q, msg);
    else if (C == 16) edge_msg_fwd_kernel<16><<<blocks, 256, 0, st>>>(x, p4, q4, rowptr, col, Q, msg);
    else edge_msg_fwd_generic_kernel<<<blocks, 256, 0, st>>>(x, p4, q4, rowptr, col, Q, C, msg);
    SN2_LAUNCH_CHECK("edge_msg_fwd_kernel");
    return SN2_OK;
}

extern "C" int sn2_edge_msg_bwd(const float *dmsg, const int *col, long long E, const int *rows_dev, int C, float *dx,
                                void *stream)
{
    if (!dmsg || !col || !dx || E < 0 || C <= 0) return SN2_EINVAL;
    if (E == 0) return SN2_OK;
    edge_msg_bwd_kernel<<<blocks_for(E * C, 256), 256, 0, (cudaStream_t)stream>>>(dmsg, col, E, rows_dev, C, dx);
    SN2_LAUNCH_CHECK("edge_msg_bwd_kernel");
    return SN2_OK;
}

extern "C" int sn2_segment_max_fwd(const float *vals, const float *ss, const int *rowptr, int Q, int C, float *out, int *arg,
                                   void *stream)
{
    if (!vals || !rowptr || !out || !arg || Q <= 0) return SN2_EINVAL;
    if (reinterpret_cast<uintptr_t>(ss) & 15) return SN2_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    if ((reinterpret_cast<uintptr_t>(vals) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(arg)) & 15) return SN2_EINVAL;
    // few long rows (global_max_pool: one row per plot) -> one CTA per row; many short rows -> one warp per row
    const bool cta_rows = (long long)Q * 32 < 148LL * 256 * 2;
    const unsigned blocks = blocks_for((long long)Q * 32, 256);
    switch (C) {
    case 16:
        if (cta_rows) segment_max_fwd_kernel<16, 8><<<Q, 256, 0, st>>>(vals, ss, rowptr, Q, out, arg);
        else segment_max_fwd_kernel<16, 1><<<blocks, 256, 0, st>>>(vals, ss, rowptr, Q, out, arg);
        break;
    case 32:
        if (cta_rows) segment_max_fwd_kernel<32, 8><<<Q, 256, 0, st>>>(vals, ss, rowptr, Q, out, arg);
        else segment_max_fwd_kernel<32, 1><<<blocks, 256, 0, st>>>(vals, ss, rowptr, Q, out, arg);
        break;
    case 64:
        if (cta_rows) segment_max_fwd_kernel<64, 8><<<Q, 256, 0, st>>>(vals, ss, rowptr, Q, out, arg);
        else segment_max_fwd_kernel<64, 1><<<blocks, 256, 0, st>>>(vals, ss, rowptr, Q, out, arg);
        break;
    default: return SN2_EUNSUPPORTED;
    }
    SN2_LAUNCH_CHECK("segment_max_fwd_kernel");
    return SN2_OK;
}

extern "C" int sn2_segment_max_bwd(const float *dout, const int *arg, long long Q, int C, long long E, float *dvals, void *stream)
{
    if (!dout || !arg || !dvals || Q <= 0 || C <= 0 || E < 0) return SN2_EINVAL;
    segment_max_bwd_kernel<<<blocks_for(Q * C, 256), 256, 0, (cudaStream_t)stream>>>(dout, arg, Q * C, C, E, dvals);
    SN2_LAUNCH_CHECK("segment_max_bwd_kernel");
    return SN2_OK;
}

extern "C" int sn2_interp3_fwd(const float *x, int ldx, const int *nbr, const float *w, long long Q, int C, float *y,
                               void *stream)
{
    if (!x || !nbr || !w || !y || Q <= 0 || C <= 0 || ldx < C) return SN2_EINVAL;
    interp3_fwd_kernel<<<blocks_for(Q * C, 256), 256, 0, (cudaStream_t)stream>>>(x, ldx, nbr, w, Q, C, y);
    SN2_LAUNCH_CHECK("interp3_fwd_kernel");
    return SN2_OK;
}

extern "C" int sn2_interp3_bwd(const float *dy, const int *nbr, const float *w, long long Q, int C, float *dx,
                               void *stream)
{
    if (!dy || !nbr || !w || !dx || Q <= 0 || C <= 0) return SN2_EINVAL;
    interp3_bwd_kernel<<<blocks_for(Q * C, 256), 256, 0, (cudaStream_t)stream>>>(dy, nbr, w, Q, C, dx);
    SN2_LAUNCH_CHECK("interp3_bwd_kernel");
    return SN2_OK;
}

extern "C" int sn2_interp_plot_fwd(const float *g, const float *pos4, int B, int M, int C, float *y, void *stream)
{
    if (!g || !pos4 || !y || B <= 0 || M <= 0 || C <= 0) return SN2_EINVAL;
    const long long Q = (long long)B * M;
    interp_plot_fwd_kernel<<<blocks_for(Q * C, 256), 256, 0, (cudaStream_t)stream>>>(
        g, reinterpret_cast<const float4 *>(pos4), Q, M, C, y);
    SN2_LAUNCH_CHECK("interp_plot_fwd_kernel");
    return SN2_OK;
}

extern "C" int sn2_interp_plot_bwd(const float *dy, const float *pos4, int B, int M, int C, float *dg, void *stream)
{
    if (!dy || !pos4 || !dg || B <= 0 || M <= 0 || C <= 0) return SN2_EINVAL;
    interp_plot_bwd_kernel<<<dim3(B, C), 256, 0, (cudaStream_t)stream>>>(dy, reinterpret_cast<const float4 *>(pos4), M, C,
                                                                        dg);
    SN2_LAUNCH_CHECK("interp_plot_bwd_kernel");
    return SN2_OK;
}

extern "C" int sn2_project_plotwise_bwd(const float *dout, const int *parg, int B, int D, float *dpred, void *stream)
{
    if (!dout || !parg || !dpred || B <= 0 || D <= 0) return SN2_EINVAL;
    project_plotwise_bwd_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(dout, parg, D, dpred);
    SN2_LAUNCH_CHECK("project_plotwise_bwd_kernel");
    return SN2_OK;
}

// ---- skinny weight gradient -------------------------------------------------------------------------
// dW [Co,Ci] = dy^T x and db [Co] = sum_rows dy for a Linear applied to E ~ 5 M edge rows with Co, Ci <= 64:
// a reduction over E with a tiny output, for which cuBLAS picks a large-K GEMM that runs at ~2 ms per
// layer (torch.profiler, config 3).  Here a group of Co threads shares a row: thread o keeps dW[o][0..Ci)
// and db[o] in registers, reads dy[r][o] and the (broadcast) x row; rows are strided over groups and CTAs;
// CTA partials are reduced in a fixed order by a second kernel (deterministic).
namespace sn2 {

constexpr int WG_THREADS = 256;

template <int CI>
__global__ void __launch_bounds__(WG_THREADS)
linear_wgrad_kernel(const float *__restrict__ dy, const float *__restrict__ x, long long E, int Co,
                    float *__restrict__ partial)
{
    extern __shared__ float wg_smem[];  // [groups][Co][CI+1]
    const int G = WG_THREADS / Co;
    const int g = threadIdx.x / Co, o = threadIdx.x - g * Co;
    float acc[CI + 1];
#pragma unroll
    for (int i = 0; i <= CI; ++i) acc[i] = 0.f;
    if (g < G) {
        for (long long r = (long long)blockIdx.x * G + g; r < E; r += (long long)gridDim.x * G) {
            const float d = __ldg(dy + r * Co + o);
            const float *xr = x + r * CI;
#pragma unroll
            for (int i = 0; i < CI; ++i) acc[i] = fmaf(d, __ldg(xr + i), acc[i]);
            acc[CI] += d;
        }
#pragma unroll
        for (int i = 0; i <= CI; ++i) wg_smem[((size_t)g * Co + o) * (CI + 1) + i] = acc[i];
    }
    __syncthreads();
    const int n = Co * (CI + 1);
    for (int t = threadIdx.x; t < n; t += WG_THREADS) {
        float s = 0.f;
        for (int gg = 0; gg < G; ++gg) s += wg_smem[(size_t)gg * n + t];
        partial[(size_t)blockIdx.x * n + t] = s;
    }
}

// one warp per output element, lanes stride over the CTA partials, xor-shuffle tree (fixed order)
__global__ void __launch_bounds__(256)
linear_wgrad_reduce_kernel(const float *__restrict__ partial, int nblk, int Co, int CI, float *__restrict__ dW,
                           float *__restrict__ db)
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

template <int CI>
static int launch_wgrad(const float *dy, const float *x, long long E, int Co, float *partial, int nblk, float *dW, float *db,
                        cudaStream_t st)
{
    const int G = WG_THREADS / Co;
    const size_t smem = (size_t)G * Co * (CI + 1) * sizeof(float);
    auto kern = linear_wgrad_kernel<CI>;
    SN2_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "wgrad attr");
    kern<<<nblk, WG_THREADS, smem, st>>>(dy, x, E, Co, partial);
    SN2_LAUNCH_CHECK("linear_wgrad_kernel");
    const int n = Co * (CI + 1);
    linear_wgrad_reduce_kernel<<<(n * 32 + 255) / 256, 256, 0, st>>>(partial, nblk, Co, CI, dW, db);
    SN2_LAUNCH_CHECK("linear_wgrad_reduce_kernel");
    return SN2_OK;
}

}  // namespace sn2

extern "C" int sn2_linear_wgrad_supported(int Co, int Ci)
{
    return (Co >= 1 && Co <= 64) && (Ci == 11 || Ci == 16 || Ci == 19 || Ci == 34 || Ci == 35 || Ci == 42);
}

extern "C" int sn2_linear_wgrad(const float *dy, const float *x, long long E, int Co, int Ci, float *partial, int nblk,
                                float *dW, float *db, void *stream)
{
    if (!dy || !x || !partial || !dW || !db || E <= 0 || nblk <= 0) return SN2_EINVAL;
    if (!sn2_linear_wgrad_supported(Co, Ci)) return SN2_EUNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    switch (Ci) {
    case 11: return sn2::launch_wgrad<11>(dy, x, E, Co, partial, nblk, dW, db, st);
    case 16: return sn2::launch_wgrad<16>(dy, x, E, Co, partial, nblk, dW, db, st);
    case 19: return sn2::launch_wgrad<19>(dy, x, E, Co, partial, nblk, dW, db, st);
    case 34: return sn2::launch_wgrad<34>(dy, x, E, Co, partial, nblk, dW, db, st);
    case 35: return sn2::launch_wgrad<35>(dy, x, E, Co, partial, nblk, dW, db, st);
    default: return sn2::launch_wgrad<42>(dy, x, E, Co, partial, nblk, dW, db, st);
    }
}
