// Fused set abstraction (eval): radius ball query + PointConv in ONE kernel, no neighbour list.
// SURVEY.md §8a a3/a4 (reference model/point_net2.py:21-29), Appendix A2 + A3.
//
// One warp per centroid, centroids taken in CELL order (the queries are binned on the same xy grid as
// the points), so the 8 warps of a CTA search the same few cells and gather the same feature rows: the
// candidate and feature traffic is served by L1 instead of L2.
//   * search: the 3 cell-row ranges around the centroid, 32 candidates per step; hits (d2 < r2) are
//     compacted through a 64-entry per-warp ring in shared memory (ballot + prefix popc);
//   * whenever 32 hits are queued every lane runs the message MLP on one of them (full lanes), keeping
//     a running per-lane max; the tail is processed masked; one REDUX max per channel at the end.
//   * cap: max aggregation does not depend on edge order, so as long as the hit count stays <= K the
//     streamed result IS the reference result.  A centroid with more than K hits (only with small caps,
//     e.g. config 5's K = 64) is appended to an overflow list and redone exactly by a second, persistent
//     launch of the same kernel (REDO = true): hits -> bitmap over the plot's point indices, first K
//     set bits in ascending index order (the canonical rule of A2), MLP over those.  Keeping the bitmap
//     out of the streaming kernel keeps its shared memory and registers (occupancy) small.
// First layer factorisation: W1 [x_j ; p_j - p_i] + b1 = (W1x x_j + W1p p_j) + (b1 - W1p p_i) = u_j + c_i.
// u_j is computed ONCE per point by sa_pre_kernel (instead of once per edge, ~25 edges per point), c_i once
// per centroid; an edge then costs a 64/128-byte gather of u_j plus C adds for layer 1.
//   * level 1 (MLP [11,16,16]): layer 1 = relu(u_j + c_i)*s + t elementwise, layer 2 = 256 FFMA per edge.
//   * level 2 (MLP [19,32], single layer): out = max_j f(u_j) with f(u) = fma(max(u + c, 0), s, t), which is
//     monotone in u (fp32 rounding keeps weak monotonicity), so out = f(max_j u_j) for s >= 0 and
//     f(min_j u_j) for s < 0 EXACTLY: the per-edge work is a running min/max of the gathered u_j.
// Results agree with the list path (sn2_ball_* + sn2_pointconv_fwd) to fp32 rounding (different
// association of the first layer), and exactly in which edges participate, including when K binds.
#include "mlp_common.cuh"

namespace sn2 {

constexpr int SF_WARPS = 8;

__device__ __forceinline__ int cell_coord_c(float v, float mn, float inv, int g)
{
    int c = (int)floorf(__fmul_rn(__fsub_rn(v, mn), inv));
    return min(max(c, 0), g - 1);
}

// u[p] = W1x x_p + W1p pos_p (no bias): the per-point half of the first layer.
template <int LEVEL>
__global__ void __launch_bounds__(128)
sa_pre_kernel(const float4 *__restrict__ pos, const float *__restrict__ feat, long long P,
              const __grid_constant__ typename SAEdge<LEVEL>::W W, float *__restrict__ u)
{
    using E = SAEdge<LEVEL>;
    constexpr int C = LEVEL == 1 ? SN2_C1 : SN2_C2;  // width of the first layer
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    float acc[C];
#pragma unroll
    for (int o = 0; o < C; ++o) acc[o] = 0.f;
#pragma unroll
    for (int v = 0; v < E::CIN / 4; ++v) {
        const float4 f = ldg4(feat + (size_t)p * E::CIN + 4 * v);
        acc_step<0>(W.l1, f.x, acc, 4 * v);
        acc_step<0>(W.l1, f.y, acc, 4 * v + 1);
        acc_step<0>(W.l1, f.z, acc, 4 * v + 2);
        acc_step<0>(W.l1, f.w, acc, 4 * v + 3);
    }
    const float4 pp = __ldg(pos + p);
    acc_step<E::CIN>(W.l1, pp.x, acc, 0);
    acc_step<E::CIN>(W.l1, pp.y, acc, 1);
    acc_step<E::CIN>(W.l1, pp.z, acc, 2);
    float4 *o4 = reinterpret_cast<float4 *>(u + (size_t)p * C);
#pragma unroll
    for (int v = 0; v < C / 4; ++v) o4[v] = make_float4(acc[4 * v], acc[4 * v + 1], acc[4 * v + 2], acc[4 * v + 3]);
}

template <int LEVEL, bool REDO>
__global__ void __launch_bounds__(SF_WARPS * 32, REDO ? 1 : (LEVEL == 1 ? 3 : 2))
sa_fused_kernel(const float *__restrict__ grid_hdr, const int *__restrict__ cell_start,
                const float4 *__restrict__ sorted, const float4 *__restrict__ qsorted,
                const float *__restrict__ u, int N, int M, float r2, int K, int words,
                const __grid_constant__ typename SAEdge<LEVEL>::W W, float *__restrict__ out,
                int *__restrict__ cnt_out, int *__restrict__ ovf)
{
    using E = SAEdge<LEVEL>;
    constexpr int C = LEVEL == 1 ? SN2_C1 : SN2_C2;  // width of the first layer (= output width)
    constexpr int UNR = 4;      // candidate loads in flight per lane: the search is latency bound otherwise
    constexpr int RING = 256;   // >= 31 leftover + UNR*32 new hits
    extern __shared__ __align__(16) unsigned char sf_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int *ring = reinterpret_cast<int *>(sf_smem) + warp * RING;
    unsigned *bm = reinterpret_cast<unsigned *>(sf_smem + SF_WARPS * RING * sizeof(int)) + (size_t)warp * 2 * words;
    unsigned *pre = bm + words;  // inclusive popc prefix (REDO only)
    const unsigned lt = (1u << lane) - 1u;

    // work items: streaming = one (plot, cell-ordered query) per warp; redo = entries of the overflow list
    long long item = REDO ? (long long)blockIdx.x * SF_WARPS + warp : 0;
    const long long n_items = REDO ? (long long)ovf[0] : 1;
    for (; item < n_items; item += REDO ? (long long)gridDim.x * SF_WARPS : 1) {
        int b, j;
        if (REDO) {
            const int code = ovf[1 + item];
            b = code / M;
            j = code - b * M;
        } else {
            b = blockIdx.y;
            j = blockIdx.x * SF_WARPS + warp;
            if (j >= M) return;
        }
        const float *hdr = grid_hdr + (size_t)b * SN2_GRID_HDR;
        const int *cs = cell_start + (size_t)b * (SN2_GRID_CELLS + 1);
        const float4 *so = sorted + (size_t)b * N;
        const float4 q = __ldg(qsorted + (size_t)b * M + j);
        const int qloc = __float_as_int(q.w);
        const float *ub = u + (size_t)b * N * C;

        const float ox = hdr[0], oy = hdr[1], inv = hdr[2];
        const int gx = __float_as_int(hdr[4]), gy = __float_as_int(hdr[5]);
        const int ix = cell_coord_c(q.x, ox, inv, gx), iy = cell_coord_c(q.y, oy, inv, gy);
        const int x0 = max(ix - 1, 0), x1 = min(ix + 1, gx - 1);
        const int y0 = max(iy - 1, 0), y1 = min(iy + 1, gy - 1);

        // c_i = b1 - W1p q  (the per-centroid half of the first layer)
        float c[C];
#pragma unroll
        for (int o = 0; o < C; ++o)
            c[o] = W.l1.b[o] - (W.l1.w[E::CIN][o] * q.x + W.l1.w[E::CIN + 1][o] * q.y + W.l1.w[E::CIN + 2][o] * q.z);
        // level 1: mx = running max of the layer-2 output; level 2: mx / mn = running max / min of u_j
        float mx[C], mn[LEVEL == 2 ? C : 1];
#pragma unroll
        for (int o = 0; o < C; ++o) mx[o] = -INFINITY;
        if (LEVEL == 2) {
#pragma unroll
            for (int o = 0; o < C; ++o) mn[o] = INFINITY;
        }
        auto edge = [&](const int id) {
            const float *ur = ub + (size_t)id * C;
            if constexpr (LEVEL == 1) {
                float h1[C];
#pragma unroll
                for (int g = 0; g < C / 4; ++g) {
                    const float4 t4 = ldg4(ur + 4 * g);
                    h1[4 * g] = t4.x + c[4 * g];
                    h1[4 * g + 1] = t4.y + c[4 * g + 1];
                    h1[4 * g + 2] = t4.z + c[4 * g + 2];
                    h1[4 * g + 3] = t4.w + c[4 * g + 3];
                }
                relu_bn(W.l1, h1);
                float h2[SN2_C1];
                acc_init(W.l2, h2);
#pragma unroll
                for (int k = 0; k < SN2_C1; ++k) acc_step<0>(W.l2, h1[k], h2, k);
                relu_bn(W.l2, h2);
#pragma unroll
                for (int o = 0; o < C; ++o) mx[o] = fmaxf(mx[o], h2[o]);
            } else {
#pragma unroll
                for (int g = 0; g < C / 4; ++g) {
                    const float4 t4 = ldg4(ur + 4 * g);
                    mx[4 * g] = fmaxf(mx[4 * g], t4.x);         mn[4 * g] = fminf(mn[4 * g], t4.x);
                    mx[4 * g + 1] = fmaxf(mx[4 * g + 1], t4.y); mn[4 * g + 1] = fminf(mn[4 * g + 1], t4.y);
                    mx[4 * g + 2] = fmaxf(mx[4 * g + 2], t4.z); mn[4 * g + 2] = fminf(mn[4 * g + 2], t4.z);
                    mx[4 * g + 3] = fmaxf(mx[4 * g + 3], t4.w); mn[4 * g + 3] = fminf(mn[4 * g + 3], t4.w);
                }
            }
        };

        if (REDO) {  // hit set as a bitmap over the plot's point indices + inclusive popcount prefix
            for (int w = lane; w < words; w += 32) bm[w] = 0u;
            __syncwarp();
            for (int yy = y0; yy <= y1; ++yy) {
                const int s2 = __ldg(cs + yy * gx + x0), e2 = __ldg(cs + yy * gx + x1 + 1);
                for (int i = s2 + lane; i < e2; i += 32) {
                    const float4 v = __ldg(so + i);
                    if (dist2(v.x, v.y, v.z, q.x, q.y, q.z) < r2) {
                        const int id = __float_as_int(v.w);
                        atomicOr(&bm[id >> 5], 1u << (id & 31));
                    }
                }
            }
            __syncwarp();
            int run = 0;
            for (int w0 = 0; w0 < words; w0 += 32) {
                const int w = w0 + lane;
                const int c1 = w < words ? __popc(bm[w]) : 0;
                int inc = c1;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(SN2_FULL, inc, o);
                    if (lane >= o) inc += t;
                }
                if (w < words) pre[w] = run + inc;
                run += __shfl_sync(SN2_FULL, inc, 31);
            }
            __syncwarp();
        }

        // ---- producer / consumer loop with ONE call site of the message MLP (keeps the loop in I-cache) ----
        int cnt = 0;
        int head = 0, tail = 0;          // warp-uniform ring cursors
        int y = y0, base = 0, e = 0;     // streaming iterator: current cell row and candidate range
        bool open_row = false;
        int rank = 0;                    // redo iterator: next rank of the set bits to extract
        bool more = true;
        while (true) {
            while (more && tail - head < 32) {
                if (!REDO) {
                    if (!open_row) {
                        if (y > y1) { more = false; break; }
                        base = __ldg(cs + y * gx + x0);
                        e = __ldg(cs + y * gx + x1 + 1);
                        open_row = true;
                    }
                    float4 vv[UNR];
#pragma unroll
                    for (int t = 0; t < UNR; ++t) {
                        const int i = base + t * 32 + lane;
                        vv[t] = i < e ? __ldg(so + i) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
#pragma unroll
                    for (int t = 0; t < UNR; ++t) {
                        const float4 v = vv[t];
                        const bool hit = (base + t * 32 + lane < e) && dist2(v.x, v.y, v.z, q.x, q.y, q.z) < r2;
                        const unsigned bal = __ballot_sync(SN2_FULL, hit);
                        if (hit) ring[(tail + __popc(bal & lt)) & (RING - 1)] = __float_as_int(v.w);
                        tail += __popc(bal);
                    }
                    base += 32 * UNR;
                    if (base >= e) { open_row = false; ++y; }
                } else {
                    const int rk = rank + lane;  // rank of the set bit this lane extracts (K < #set bits here)
                    if (rk < K) {
                        int lo_w = 0, hi_w = words - 1;  // first word with pre[w] > rk
                        while (lo_w < hi_w) {
                            const int mid = (lo_w + hi_w) >> 1;
                            if ((int)pre[mid] > rk) hi_w = mid; else lo_w = mid + 1;
                        }
                        const int before = lo_w ? (int)pre[lo_w - 1] : 0;
                        ring[(tail + lane) & (RING - 1)] = (lo_w << 5) + __fns(bm[lo_w], 0, rk - before + 1);
                    }
                    const int n = min(32, K - rank);
                    tail += n;
                    rank += n;
                    if (rank >= K) more = false;
                }
                __syncwarp();
            }
            const int avail = tail - head;
            if (avail <= 0) break;
            const int take = min(avail, 32);  // < 32 only once the producer is exhausted
            const int id = ring[(head + lane) & (RING - 1)];
            head += take;
            cnt += take;
            __syncwarp();
            if (lane < take) edge(id);
        }

        const size_t row = (size_t)b * M + qloc;
        if (!REDO && cnt > K) {  // the cap binds: leave this centroid to the exact redo launch
            if (lane == 0) ovf[1 + atomicAdd(ovf, 1)] = b * M + j;
            continue;
        }
#pragma unroll
        for (int o = 0; o < C; ++o) {
            if constexpr (LEVEL == 1) {
                mx[o] = cnt > 0 ? warp_max(mx[o]) : 0.f;
            } else {
                // out = max_j f(u_j) = f(max u) if s >= 0 else f(min u): f is monotone in u
                const float hi = warp_max(mx[o]), lo = -warp_max(-mn[o]);
                const float sel = W.l1.s[o] >= 0.f ? hi : lo;
                mx[o] = cnt > 0 ? fmaf(fmaxf(sel + c[o], 0.f), W.l1.s[o], W.l1.t[o]) : 0.f;
            }
        }
        if (lane == 0) {
            float4 *o4 = reinterpret_cast<float4 *>(out + row * C);
#pragma unroll
            for (int v = 0; v < C / 4; ++v) o4[v] = make_float4(mx[4 * v], mx[4 * v + 1], mx[4 * v + 2], mx[4 * v + 3]);
            if (cnt_out) cnt_out[row] = cnt;
        }
    }
}

__global__ void zero_int_kernel(int *p) { *p = 0; }

template <int LEVEL>
static int launch_sa_fused(const float *grid_hdr, const int *cell_start, const float *sorted4, const float *qsorted4,
                           const float *pos4, const float *feat, float *u_scratch, int *ovf, int B, int N, int M,
                           float r2, int K, const float *w_host, int nw, float *out, int *cnt_out, cudaStream_t st)
{
    typename SAEdge<LEVEL>::W w;
    if (int rc = load_weights(w, w_host, nw)) return rc;
    const long long P = (long long)B * N;
    zero_int_kernel<<<1, 1, 0, st>>>(ovf);
    sa_pre_kernel<LEVEL><<<(unsigned)((P + 127) / 128), 128, 0, st>>>(reinterpret_cast<const float4 *>(pos4), feat, P, w,
                                                                     u_scratch);
    SN2_LAUNCH_CHECK("sa_pre_kernel");
    const int words = (N + 31) / 32;
    const size_t smem_ring = (size_t)SF_WARPS * 256 * sizeof(int);
    {
        auto kern = sa_fused_kernel<LEVEL, false>;
        dim3 grid((M + SF_WARPS - 1) / SF_WARPS, B);
        kern<<<grid, SF_WARPS * 32, smem_ring, st>>>(grid_hdr, cell_start, reinterpret_cast<const float4 *>(sorted4),
                                                     reinterpret_cast<const float4 *>(qsorted4), u_scratch, N, M, r2, K, 0, w,
                                                     out, cnt_out, ovf);
        SN2_LAUNCH_CHECK("sa_fused_kernel");
    }
    if (K < N) {  // the cap can bind: exact redo of the overflow list (exits at once when the list is empty)
        const size_t smem = smem_ring + (size_t)SF_WARPS * 2 * words * sizeof(unsigned);
        if (smem > 200 * 1024) return SN2_EUNSUPPORTED;
        auto kern = sa_fused_kernel<LEVEL, true>;
        SN2_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "sa_fused attr");
        kern<<<148 * 2, SF_WARPS * 32, smem, st>>>(grid_hdr, cell_start, reinterpret_cast<const float4 *>(sorted4),
                                                   reinterpret_cast<const float4 *>(qsorted4), u_scratch, N, M, r2, K, words,
                                                   w, out, cnt_out, ovf);
        SN2_LAUNCH_CHECK("sa_fused_kernel<redo>");
    }
    return SN2_OK;
}

}  // namespace sn2

extern "C" int sn2_sa_fused_fwd(int level, const float *grid_hdr, const int *cell_start, const float *sorted4,
                                const float *qsorted4, const float *pos4, const float *feat, float *u_scratch,
                                int *ovf_scratch, int B, int N, int M, float r2, int K, const float *w_host, int nw,
                                float *out, int *cnt_out, void *stream)
{
    if (!grid_hdr || !cell_start || !sorted4 || !qsorted4 || !pos4 || !feat || !u_scratch || !ovf_scratch || !out ||
        B <= 0 || N <= 0 || M <= 0 || K <= 0)
        return SN2_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    if (level == 1)
        return sn2::launch_sa_fused<1>(grid_hdr, cell_start, sorted4, qsorted4, pos4, feat, u_scratch, ovf_scratch, B, N, M,
                                       r2, K, w_host, nw, out, cnt_out, st);
    if (level == 2)
        return sn2::launch_sa_fused<2>(grid_hdr, cell_start, sorted4, qsorted4, pos4, feat, u_scratch, ovf_scratch, B, N, M,
                                       r2, K, w_host, nw, out, cnt_out, st);
    return SN2_EINVAL;
}
