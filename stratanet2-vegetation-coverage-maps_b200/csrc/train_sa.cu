// Train-mode SA1 block WITHOUT materialised messages (reference
// model/point_net2.py:21-29 PointConv(local_nn = MLP([11, 16, 16])) + max aggregation under model.train()).
//
// The materialising path (edge_msg -> lrb_fwd x2 -> segment_max; backward segment_max_bwd -> lrb_bwd x2 -> edge_msg_bwd)
// writes and re-reads three [E, 11..16] arrays each way (E = 4.2 M edges at config 3: ~1.6 GB forward, ~3.4 GB backward).
// Here every sweep walks the CSR neighbour lists (one warp per centroid, lane = edge) and RECOMPUTES the message MLP
// from a per-point table, exactly as the eval kernel does (sa_fused.cu):
//     z1 = W1 [x_j ; p_j - q_i] + b1 = (W1x x_j + W1p p_j) + (b1 - W1p q_i) = u_j + c_i,     a1 = relu(z1)
//     h1 = BN1(a1) = a1 s1 + t1,   z2 = W2 h1 + b2 = (W2 diag(s1)) a1 + (b2 + W2 t1),         a2 = relu(z2)
//     x1[i] = max_e BN2(a2) = max_e (a2 s2 + t2)
// BatchNorm needs the statistics of a1 before anything of layer 2 can be evaluated, and those of a2 before the max
// can be finished, so the forward is two sweeps and a per-centroid finish:
//   pre    u_j for every point                                                   (P x 16 floats, L2 resident)
//   F1     sum a1, sum a1^2                                                      -> bn_finalize -> s1, t1
//   F2     sum a2, sum a2^2 and, per (centroid, channel), the arg-max edge of sign(gamma2) * a2 -- the sign of the
//          BatchNorm scale is known before its statistics are (invstd > 0)       -> bn_finalize -> s2, t2
//   finish x1 = a2[arg] s2 + t2
// Backward, given dx1 (sparse in the edges: one arg-max edge per centroid and channel):
//   B0     S1 = sum dx1, S2 = sum dx1 * a2[arg]  (the raw sums of BatchNorm 2's backward; dgamma2, dbeta2)
//   B1     sweep: dz2 = relu'(a2) BN2'(dz) for EVERY edge (the mean terms of BatchNorm's backward are dense);
//          G[o][k] = sum_e dz2[e][o] a1[e][k], db2 = sum_e dz2.  From those alone: dW2 = G diag(s1) + db2 t1^T, and
//          the raw sums of BatchNorm 1's backward T1 = W2^T db2, T2[k] = sum_o W2[o][k] G[o][k] -- no second reduction
//          sweep, and dh1 is not even evaluated in this pass.
//   B2     sweep: dh1 = W2^T dz2, dz1 = relu'(a1) BN1'(dh1); du[col] += dz1 (vector atomics), dc_i = sum_row dz1
//   W1     dW1x = sum_j du_j x_j^T, dW1p = sum_j du_j p_j^T - sum_i dc_i q_i^T, db1 = sum_j du_j     (per point)
// Statistics travel as raw fp64 sums + the edge count exactly like the lrb_* kernels (train_mlp.cu), so the host side
// reuses sn2_bn_finalize[_sync] / sn2_bn_param_grad / sn2_bn_bwd_sync and SyncBatchNorm is unchanged.  Weights are read
// from the live parameter tensors (device memory): nothing is baked in, the step stays graph-capturable.
#include "sn2_common.cuh"

namespace sn2 {
namespace {

constexpr int TS_WARPS = 8;
constexpr int TS_T = TS_WARPS * 32;
constexpr int TC = 16;              // width of both layers
constexpr int TF = 8;               // point features
constexpr int TK1 = TF + 3;         // inputs of layer 1
constexpr int TS_NP2 = TC * (TC + 1);  // [o][k] partial of G, column TC = db2
constexpr int TS_LD = 20;           // row stride of the per-warp tiles (conflict-free float4 rows)
constexpr int TS_CHUNK = 8;         // centroids a warp pulls from the work queue at a time

struct __align__(16) TsW {                         // shared-memory image of the parameters and BatchNorm coefficients
    float w1p[3][TC];                // [k][o] = W1[o][TF + k]
    float b1[TC];
    float w2f[TC][TC];               // [k][o] = W2[o][k] * s1[k]   (BatchNorm 1 folded into layer 2)
    float w2t[TC][TC];               // [o][k] = W2[o][k]
    float b2f[TC];                   // b2 + W2 t1
    float sgn2[TC];                  // sign of gamma2: the arg-max key is sgn2 * a2
    float cA2[TC], cB2[TC], cC2[TC];  // da2 = cA2 dz + cB2 a2 + cC2   (BatchNorm 2 backward)
    float cA1[TC], cB1[TC], cC1[TC];  // da1 = cA1 dh1 + cB1 a1 + cC1  (BatchNorm 1 backward)
};

struct TsArgs {
    const float *u;
    const float4 *qpos;
    const int *rowptr, *col;
    int M;
    const float *W1, *b1, *W2, *b2;       // live parameters: [16][11], [16], [16][16], [16]
    const float *ss1, *ss2, *gamma2;      // [4*16] scale | shift | mean | invstd
    const double *stats1, *stats2;        // forward sums, count at [32]
    const double *sums1, *sums2;          // backward raw sums [32]
    const float *dout;                    // [M][16]
    const int *arg;                       // [M][16]
    double *stats_out;                    // F1 / F2
    float *key;                           // F2: [M][16]
    int *arg_out;                         // F2: [M][16]
    float *partial;                       // B1: [gridDim.x][TS_NP2]
    float *du, *dc;                       // B2: [P][16], [M][16]
    int *queue;                           // work queue cursor (zero before the launch)
};

// dy = mask * (cA * dz + cB * y + cC): BatchNorm backward from the raw sums S1 = sum dz, S2 = sum dz * y (train_mlp.cu)
template <int C = TC>
__device__ __forceinline__ void bn_bwd_coeff(const float *ss, const double *sums, double n, int o, float &cA, float &cB, float &cC)
{
    const float sc = ss[o], mean = ss[2 * C + o], inv = ss[3 * C + o];
    const double S1 = sums[o], S2 = sums[C + o];
    const float m1 = (float)(S1 / n);
    const float m2 = (float)((double)inv * (S2 - (double)mean * S1) / n);
    const float kb = -inv * m2;
    cA = sc;
    cB = sc * kb;
    cC = sc * (-m1 - mean * kb);
}

struct KeyEdge {
    float k;
    int e;
};
// larger key wins, equal keys: lower edge index (first edge, as segment_max)
__device__ __forceinline__ KeyEdge ke_merge(KeyEdge a, KeyEdge b) { return (b.k > a.k || (b.k == a.k && b.e < a.e)) ? b : a; }
__device__ __forceinline__ KeyEdge ke_shfl(KeyEdge v, int d)
{
    KeyEdge r;
    r.k = __shfl_xor_sync(SN2_FULL, v.k, d);
    r.e = __shfl_xor_sync(SN2_FULL, v.e, d);
    return r;
}
__device__ __forceinline__ float f_shfl(float v, int d) { return __shfl_xor_sync(SN2_FULL, v, d); }
__device__ __forceinline__ float f_add(float a, float b) { return a + b; }

// Reduce 16 per-lane values over the 32 lanes with a transposing butterfly (8 + 4 + 2 + 1 + 1 exchanges instead of
// 16 x 5): lanes 2c and 2c + 1 return channel c.
template <typename T, typename Shfl, typename Op>
__device__ __forceinline__ T butterfly16(const T (&v)[TC], int lane, Shfl shfl, Op op)
{
    T v8[8], v4[4], v2[2];
    {
        const bool hi = lane & 16;
#pragma unroll
        for (int i = 0; i < 8; ++i) v8[i] = op(hi ? v[i + 8] : v[i], shfl(hi ? v[i] : v[i + 8], 16));
    }
    {
        const bool hi = lane & 8;
#pragma unroll
        for (int i = 0; i < 4; ++i) v4[i] = op(hi ? v8[i + 4] : v8[i], shfl(hi ? v8[i] : v8[i + 4], 8));
    }
    {
        const bool hi = lane & 4;
#pragma unroll
        for (int i = 0; i < 2; ++i) v2[i] = op(hi ? v4[i + 2] : v4[i], shfl(hi ? v4[i] : v4[i + 2], 4));
    }
    const bool hi = lane & 2;
    T v1 = op(hi ? v2[1] : v2[0], shfl(hi ? v2[0] : v2[1], 2));
    return op(v1, shfl(v1, 1));
}

__device__ __forceinline__ double warp_sum_f64(double v)
{
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(SN2_FULL, v, s);
    return v;
}

__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];\n" ::"l"(p)); }

__device__ __forceinline__ void red_add_v4(float *p, float a, float b, float c, float d)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// u_j = W1x x_j + W1p p_j   (CF point features -> CO channels; W [CO][CF + 3])
template <int CF, int CO>
__global__ void __launch_bounds__(256)
sa_t_pre_kernel(const float *__restrict__ feat, const float4 *__restrict__ pos, long long P, const float *__restrict__ W1,
                float *__restrict__ u)
{
    constexpr int CK = CF + 3;
    __shared__ __align__(16) float w[CK][CO];  // [k][o]
    for (int i = threadIdx.x; i < CK * CO; i += 256) w[i / CO][i % CO] = __ldg(W1 + (i % CO) * CK + i / CO);
    __syncthreads();
    const long long p = (long long)blockIdx.x * 256 + threadIdx.x;
    if (p >= P) return;
    float in[CK];
#pragma unroll
    for (int g = 0; g < CF / 4; ++g) {
        const float4 f = ldg4(feat + p * CF + 4 * g);
        in[4 * g] = f.x; in[4 * g + 1] = f.y; in[4 * g + 2] = f.z; in[4 * g + 3] = f.w;
    }
    const float4 pp = __ldg(pos + p);
    in[CF] = pp.x; in[CF + 1] = pp.y; in[CF + 2] = pp.z;
    float acc[CO];
#pragma unroll
    for (int o = 0; o < CO; ++o) acc[o] = 0.f;
#pragma unroll
    for (int k = 0; k < CK; ++k) {
#pragma unroll
        for (int g = 0; g < CO / 4; ++g) {
            const float4 wv = *reinterpret_cast<const float4 *>(&w[k][4 * g]);
            acc[4 * g] = fmaf(in[k], wv.x, acc[4 * g]);
            acc[4 * g + 1] = fmaf(in[k], wv.y, acc[4 * g + 1]);
            acc[4 * g + 2] = fmaf(in[k], wv.z, acc[4 * g + 2]);
            acc[4 * g + 3] = fmaf(in[k], wv.w, acc[4 * g + 3]);
        }
    }
    float4 *o4 = reinterpret_cast<float4 *>(u + p * CO);
#pragma unroll
    for (int g = 0; g < CO / 4; ++g) o4[g] = make_float4(acc[4 * g], acc[4 * g + 1], acc[4 * g + 2], acc[4 * g + 3]);
}

// zero the sums, count = live edge count = rowptr[M]
__global__ void sa1t_stats_init_kernel(double *stats, const int *__restrict__ rowptr, int M, int *queue, int C)
{
    const int t = threadIdx.x;
    if (t == 2 * C + 1) *queue = 0;
    if (t < 2 * C) stats[t] = 0.0;
    if (t == 2 * C) stats[2 * C] = (double)__ldg(rowptr + M);
}

// The sweeps read the parameters as constant-bank operands (uniform loads / direct FFMA operands), like the eval kernel:
// a first version kept them in shared memory and stalled on the MIO queue (64 broadcast LDS.128 per layer and pass; ncu:
// short_scoreboard + mio_throttle > half of the issue slots lost).  The parameters are LIVE device tensors, so a one-CTA
// kernel builds the image (folded weights, BatchNorm coefficients) in device memory and a device-to-device copy on the
// same stream moves it into the constant bank before the sweep -- both are ordinary graph nodes.  One training stream
// per process (as everywhere in the train path).
__constant__ TsW c_ts;
__device__ TsW g_ts_stage;

template <int MODE>
__global__ void __launch_bounds__(TS_T)
sa1t_prep_kernel(const TsArgs a)
{
    TsW &S = g_ts_stage;
    const int tid = threadIdx.x;
    for (int i = tid; i < 3 * TC; i += TS_T) S.w1p[i / TC][i % TC] = __ldg(a.W1 + (i % TC) * TK1 + TF + i / TC);
    if (tid < TC) S.b1[tid] = __ldg(a.b1 + tid);
    if constexpr (MODE >= 1) {
        for (int i = tid; i < TC * TC; i += TS_T) {
            const int k = i / TC, o = i % TC;
            const float wv = __ldg(a.W2 + o * TC + k);
            S.w2f[k][o] = wv * __ldg(a.ss1 + k);
            S.w2t[o][k] = wv;
        }
        if (tid < TC) {
            float bb = __ldg(a.b2 + tid);
            for (int k = 0; k < TC; ++k) bb = fmaf(__ldg(a.ss1 + TC + k), __ldg(a.W2 + tid * TC + k), bb);
            S.b2f[tid] = bb;
            S.sgn2[tid] = __ldg(a.gamma2 + tid) < 0.f ? -1.f : 1.f;
        }
    }
    if (MODE >= 3 && tid < TC) bn_bwd_coeff(a.ss2, a.sums2, a.stats2[2 * TC], tid, S.cA2[tid], S.cB2[tid], S.cC2[tid]);
    if (MODE == 4 && tid < TC) bn_bwd_coeff(a.ss1, a.sums1, a.stats1[2 * TC], tid, S.cA1[tid], S.cB1[tid], S.cC1[tid]);
}

// MODE 0: F1, 1: F2, 3: B1, 4: B2 (see the file header)
template <int MODE>
__global__ void __launch_bounds__(TS_T, MODE == 0 ? 3 : 2)
sa1t_sweep_kernel(const TsArgs a)
{
    const TsW &S = c_ts;
    __shared__ double red[2 * TC];
    __shared__ __align__(16) float tile[MODE == 3 ? TS_WARPS * 2 * 32 * TS_LD : 4];  // B1, per warp: a1 [32][LD] | dz2 [32][LD]
    static_assert(MODE != 3 || TS_WARPS * 2 * 32 * TS_LD >= TS_WARPS * TS_NP2, "tile doubles as the CTA reduction buffer");
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 2 * TC) red[tid] = 0.0;
    __syncthreads();

    // per-lane accumulators that live across all the centroids of this warp
    float sa[(MODE <= 1) ? TC : 1], sq[(MODE <= 1) ? TC : 1];  // F1 / F2: sum, sum of squares
    float gacc[(MODE == 3) ? 8 : 1], dbacc[(MODE == 3) ? TC : 1];  // B1: G[ob..ob+8)[k] of this lane, db2
    if constexpr (MODE <= 1) {
#pragma unroll
        for (int o = 0; o < TC; ++o) sa[o] = sq[o] = 0.f;
    }
    if constexpr (MODE == 3) {
#pragma unroll
        for (int o = 0; o < 8; ++o) gacc[o] = 0.f;
#pragma unroll
        for (int o = 0; o < TC; ++o) dbacc[o] = 0.f;
    }
    float *tA = tile + (MODE == 3 ? warp * (2 * 32 * TS_LD) : 0), *tD = tA + (MODE == 3 ? 32 * TS_LD : 0);

    // Work queue: warps pull chunks of TS_CHUNK consecutive centroids.  A static partition over a persistent grid breaks
    // down in the training loop, where the FPS CTAs of the NEXT batch (196 KB of shared memory each) hold 32 SMs: the
    // CTAs that do not fit beside them would run as a second wave.
    for (;;) {
        int base = 0;
        if (lane == 0) base = atomicAdd(a.queue, TS_CHUNK);
        base = __shfl_sync(SN2_FULL, base, 0);
        if (base >= a.M) break;
        const int rp_l = (lane <= TS_CHUNK && base + lane <= a.M) ? __ldg(a.rowptr + base + lane) : 0;
        const float4 q_l = (lane < TS_CHUNK && base + lane < a.M) ? __ldg(a.qpos + base + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
        // The passes of a chunk walk one contiguous range of the edge list.  A warp has nothing to overlap a pass's gather
        // latency with (16 warps per SM at 128 registers), so the u rows are pulled into L1 ahead of time: `pf` is the
        // first edge not yet prefetched, `pcol` the column indices of [pf, pf + 32), loaded a pass earlier.
        const int e_chunk = __shfl_sync(SN2_FULL, rp_l, min(TS_CHUNK, a.M - base));
        int pf = __shfl_sync(SN2_FULL, rp_l, 0) + 32;
        int pcol = pf + lane < e_chunk ? __ldg(a.col + pf + lane) : -1;
      for (int ci = 0; ci < TS_CHUNK && base + ci < a.M; ++ci) {
        const int i = base + ci;
        const int s = __shfl_sync(SN2_FULL, rp_l, ci), e = __shfl_sync(SN2_FULL, rp_l, ci + 1);
        float4 q;
        q.x = __shfl_sync(SN2_FULL, q_l.x, ci);
        q.y = __shfl_sync(SN2_FULL, q_l.y, ci);
        q.z = __shfl_sync(SN2_FULL, q_l.z, ci);
        float c[TC];
#pragma unroll
        for (int o = 0; o < TC; ++o) c[o] = S.b1[o] - (S.w1p[0][o] * q.x + S.w1p[1][o] * q.y + S.w1p[2][o] * q.z);
        KeyEdge best[(MODE == 1) ? TC : 1];
        float dcs[(MODE == 4) ? TC : 1];
        if constexpr (MODE == 1) {
#pragma unroll
            for (int o = 0; o < TC; ++o) { best[o].k = -INFINITY; best[o].e = 0x7fffffff; }
        }
        if constexpr (MODE == 4) {
#pragma unroll
            for (int o = 0; o < TC; ++o) dcs[o] = 0.f;
        }
        for (int j0 = s; j0 < e; j0 += 32) {
            while (pf < j0 + 64 && pf < e_chunk) {  // warp-uniform; one round per full pass
                if (pcol >= 0) prefetch_l1(a.u + (size_t)pcol * TC);
                pf += 32;
                pcol = pf + lane < e_chunk ? __ldg(a.col + pf + lane) : -1;
            }
            const int j = j0 + lane;
            const bool valid = j < e;
            const int p = __ldg(a.col + (valid ? j : s));
            const float *ur = a.u + (size_t)p * TC;
            float a1[TC];
#pragma unroll
            for (int g = 0; g < TC / 4; ++g) {
                const float4 t4 = ldg4(ur + 4 * g);
                a1[4 * g] = fmaxf(t4.x + c[4 * g], 0.f);
                a1[4 * g + 1] = fmaxf(t4.y + c[4 * g + 1], 0.f);
                a1[4 * g + 2] = fmaxf(t4.z + c[4 * g + 2], 0.f);
                a1[4 * g + 3] = fmaxf(t4.w + c[4 * g + 3], 0.f);
            }
            if constexpr (MODE == 0) {
                if (valid) {
#pragma unroll
                    for (int o = 0; o < TC; ++o) {
                        sa[o] += a1[o];
                        sq[o] = fmaf(a1[o], a1[o], sq[o]);
                    }
                }
                continue;
            }
            float a2[TC];
#pragma unroll
            for (int o = 0; o < TC; ++o) a2[o] = S.b2f[o];
#pragma unroll
            for (int k = 0; k < TC; ++k) {
#pragma unroll
                for (int o = 0; o < TC; ++o) a2[o] = fmaf(a1[k], S.w2f[k][o], a2[o]);
            }
#pragma unroll
            for (int o = 0; o < TC; ++o) a2[o] = fmaxf(a2[o], 0.f);
            if constexpr (MODE == 1) {
                if (valid) {
#pragma unroll
                    for (int o = 0; o < TC; ++o) {
                        sa[o] += a2[o];
                        sq[o] = fmaf(a2[o], a2[o], sq[o]);
                        const float kv = S.sgn2[o] * a2[o];
                        if (kv > best[o].k) { best[o].k = kv; best[o].e = j; }  // ascending j inside a lane: first edge stays
                    }
                }
                continue;
            }
            // ---- backward sweeps: dz2 = relu'(a2) * BatchNorm2'(dz), dz = dout[i][c] on the arg-max edge of channel c ----
            float dz2[TC];
#pragma unroll
            for (int g = 0; g < TC / 4; ++g) {
                const int4 ag = __ldg(reinterpret_cast<const int4 *>(a.arg + (size_t)i * TC) + g);
                const float4 dg = ldg4(a.dout + (size_t)i * TC + 4 * g);
                const float d0 = ag.x == j ? dg.x : 0.f, d1 = ag.y == j ? dg.y : 0.f, d2 = ag.z == j ? dg.z : 0.f, d3 = ag.w == j ? dg.w : 0.f;
                const int o = 4 * g;
                dz2[o] = (valid && a2[o] > 0.f) ? fmaf(S.cA2[o], d0, fmaf(S.cB2[o], a2[o], S.cC2[o])) : 0.f;
                dz2[o + 1] = (valid && a2[o + 1] > 0.f) ? fmaf(S.cA2[o + 1], d1, fmaf(S.cB2[o + 1], a2[o + 1], S.cC2[o + 1])) : 0.f;
                dz2[o + 2] = (valid && a2[o + 2] > 0.f) ? fmaf(S.cA2[o + 2], d2, fmaf(S.cB2[o + 2], a2[o + 2], S.cC2[o + 2])) : 0.f;
                dz2[o + 3] = (valid && a2[o + 3] > 0.f) ? fmaf(S.cA2[o + 3], d3, fmaf(S.cB2[o + 3], a2[o + 3], S.cC2[o + 3])) : 0.f;
            }
            if constexpr (MODE == 3) {
#pragma unroll
                for (int o = 0; o < TC; ++o) dbacc[o] += dz2[o];
                __syncwarp();  // the previous pass has been read
#pragma unroll
                for (int g = 0; g < TC / 4; ++g) {
                    *reinterpret_cast<float4 *>(tA + lane * TS_LD + 4 * g) = make_float4(a1[4 * g], a1[4 * g + 1], a1[4 * g + 2], a1[4 * g + 3]);
                    *reinterpret_cast<float4 *>(tD + lane * TS_LD + 4 * g) = make_float4(dz2[4 * g], dz2[4 * g + 1], dz2[4 * g + 2], dz2[4 * g + 3]);
                }
                __syncwarp();
                // G[o][k] += dz2[r][o] * a1[r][k] over the edges r of this pass: lane = (k, half of the o range)
                const int take = min(32, e - j0), k = lane >> 1, ob = (lane & 1) * 8;
#pragma unroll 4
                for (int r = 0; r < take; ++r) {
                    const float av = tA[r * TS_LD + k];
                    const float4 d0 = *reinterpret_cast<const float4 *>(tD + r * TS_LD + ob);
                    const float4 d1 = *reinterpret_cast<const float4 *>(tD + r * TS_LD + ob + 4);
                    gacc[0] = fmaf(av, d0.x, gacc[0]);
                    gacc[1] = fmaf(av, d0.y, gacc[1]);
                    gacc[2] = fmaf(av, d0.z, gacc[2]);
                    gacc[3] = fmaf(av, d0.w, gacc[3]);
                    gacc[4] = fmaf(av, d1.x, gacc[4]);
                    gacc[5] = fmaf(av, d1.y, gacc[5]);
                    gacc[6] = fmaf(av, d1.z, gacc[6]);
                    gacc[7] = fmaf(av, d1.w, gacc[7]);
                }
            }
            if constexpr (MODE == 4) {
                float dh1[TC];
#pragma unroll
                for (int k = 0; k < TC; ++k) dh1[k] = 0.f;
#pragma unroll
                for (int o = 0; o < TC; ++o) {
#pragma unroll
                    for (int k = 0; k < TC; ++k) dh1[k] = fmaf(dz2[o], S.w2t[o][k], dh1[k]);
                }
#pragma unroll
                for (int k = 0; k < TC; ++k) {
                    dh1[k] = (valid && a1[k] > 0.f) ? fmaf(S.cA1[k], dh1[k], fmaf(S.cB1[k], a1[k], S.cC1[k])) : 0.f;  // = dz1
                    dcs[k] += dh1[k];
                }
                if (valid) {
                    float *dp = a.du + (size_t)p * TC;
#pragma unroll
                    for (int g = 0; g < TC / 4; ++g) red_add_v4(dp + 4 * g, dh1[4 * g], dh1[4 * g + 1], dh1[4 * g + 2], dh1[4 * g + 3]);
                }
            }
        }
        if constexpr (MODE == 1) {
            const KeyEdge r = butterfly16(best, lane, ke_shfl, ke_merge);
            if (!(lane & 1)) {
                a.key[(size_t)i * TC + (lane >> 1)] = r.k;
                a.arg_out[(size_t)i * TC + (lane >> 1)] = e > s ? r.e : -1;
            }
        }
        if constexpr (MODE == 4) {
            const float r = butterfly16(dcs, lane, f_shfl, f_add);
            if (!(lane & 1)) a.dc[(size_t)i * TC + (lane >> 1)] = r;
        }
      }
    }

    if constexpr (MODE <= 1) {  // per-lane fp32 partial sums -> fp64 CTA sums -> one fp64 atomic per channel and CTA
#pragma unroll
        for (int o = 0; o < TC; ++o) {
            const double s1 = warp_sum_f64((double)sa[o]), s2 = warp_sum_f64((double)sq[o]);
            if (lane == 0) {
                atomicAdd(red + o, s1);
                atomicAdd(red + TC + o, s2);
            }
        }
        __syncthreads();
        if (tid < 2 * TC) atomicAdd(a.stats_out + tid, red[tid]);
    }
    if constexpr (MODE == 3) {  // warp partials -> CTA partial [o][k | db]
        const float dbr = butterfly16(dbacc, lane, f_shfl, f_add);
        __syncthreads();  // every warp is done with its tiles
        float *buf = tile + warp * TS_NP2;
        const int k = lane >> 1, ob = (lane & 1) * 8;
#pragma unroll
        for (int o = 0; o < 8; ++o) buf[(ob + o) * (TC + 1) + k] = gacc[o];
        if (!(lane & 1)) buf[(lane >> 1) * (TC + 1) + TC] = dbr;
        __syncthreads();
        for (int t = tid; t < TS_NP2; t += TS_T) {
            float v = 0.f;
#pragma unroll
            for (int w = 0; w < TS_WARPS; ++w) v += tile[w * TS_NP2 + t];
            a.partial[(size_t)blockIdx.x * TS_NP2 + t] = v;
        }
    }
}

// x = BN(a[arg]) = key * sgn * s + t (0 for a centroid without neighbours); amax = a[arg]
template <int C>
__global__ void __launch_bounds__(256)
sa_t_finish_kernel(const float *__restrict__ key, const int *__restrict__ arg, const float *__restrict__ gamma,
                   const float *__restrict__ ss, long long n, float *__restrict__ x, float *__restrict__ amax)
{
    const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
    if (t >= n) return;
    const int o = (int)(t % C);
    const bool any = __ldg(arg + t) >= 0;
    const float av = any ? (__ldg(gamma + o) < 0.f ? -__ldg(key + t) : __ldg(key + t)) : 0.f;
    amax[t] = av;
    x[t] = any ? fmaf(av, __ldg(ss + o), __ldg(ss + C + o)) : 0.f;
}

// B0: sums[o] = sum_i dout[i][o], sums[C + o] = sum_i dout[i][o] * amax[i][o] over the centroids that have an arg-max edge
template <int C>
__global__ void __launch_bounds__(256)
sa_t_bwd_sums_kernel(const float *__restrict__ dout, const float *__restrict__ amax, const int *__restrict__ arg, long long M,
                     double *__restrict__ sums)
{
    static_assert(C == 16 || C == 32, "channel <-> lane mapping below");
    __shared__ double red[2 * C];
    if (threadIdx.x < 2 * C) red[threadIdx.x] = 0.0;
    __syncthreads();
    // thread = (row slot, channel): 256 / C rows per step
    const int o = threadIdx.x % C;
    double s1 = 0.0, s2 = 0.0;
    for (long long i = (long long)blockIdx.x * (256 / C) + threadIdx.x / C; i < M; i += (long long)gridDim.x * (256 / C)) {
        const long long t = i * C + o;
        if (__ldg(arg + t) >= 0) {
            const float d = __ldg(dout + t);
            s1 += (double)d;
            s2 += (double)d * (double)__ldg(amax + t);
        }
    }
    if constexpr (C == 16) {  // lanes l and l + 16 hold the same channel
        s1 += __shfl_xor_sync(SN2_FULL, s1, 16);
        s2 += __shfl_xor_sync(SN2_FULL, s2, 16);
    }
    if ((threadIdx.x & 31) < C) {
        atomicAdd(red + o, s1);
        atomicAdd(red + C + o, s2);
    }
    __syncthreads();
    if (threadIdx.x < 2 * C) atomicAdd(sums + threadIdx.x, red[threadIdx.x]);
}

// CTA partials of B1 -> dW2 = G diag(s1) + db2 t1^T, db2, and BatchNorm 1's raw backward sums T1 = W2^T db2,
// T2[k] = sum_o W2[o][k] G[o][k].  One CTA; fixed summation order.
__global__ void __launch_bounds__(TS_NP2)
sa1t_w2_finish_kernel(const float *__restrict__ partial, int nblk, const float *__restrict__ W2, const float *__restrict__ ss1,
                      float *__restrict__ dW2, float *__restrict__ db2, double *__restrict__ sums1)
{
    __shared__ double G[TS_NP2];
    const int t = threadIdx.x;
    double v = 0.0;
    for (int b = 0; b < nblk; ++b) v += (double)__ldg(partial + (size_t)b * TS_NP2 + t);
    G[t] = v;
    __syncthreads();
    const int o = t / (TC + 1), k = t - o * (TC + 1);
    if (k < TC) dW2[o * TC + k] = (float)(G[t] * (double)__ldg(ss1 + k) + G[o * (TC + 1) + TC] * (double)__ldg(ss1 + TC + k));
    else db2[o] = (float)G[t];
    if (t < TC) {  // t = k
        double t1 = 0.0, t2 = 0.0;
        for (int oo = 0; oo < TC; ++oo) {
            const double w = (double)__ldg(W2 + oo * TC + t);
            t1 += w * G[oo * (TC + 1) + TC];
            t2 += w * G[oo * (TC + 1) + t];
        }
        sums1[t] = t1;
        sums1[TC + t] = t2;
    }
}

// W1: dW1[o][k] = sum_j du[j][o] in_j[k] - sum_i dc[i][o] q_i[k - CF] (k >= CF), db1[o] = sum_j du[j][o]; CTA partials
// in the [o][k | db] layout of lrb_wgrad_reduce.  thread = (o, k), k = CF + 3 stands for the bias (in = 1).
constexpr int W1_TILE = 64;
template <int CO, int CF>
__global__ void __launch_bounds__(CO * (CF + 4))
sa_t_w1_kernel(const float *__restrict__ du, const float *__restrict__ dc, const float *__restrict__ feat,
               const float4 *__restrict__ pos, const float4 *__restrict__ qpos, long long P, long long M, float *__restrict__ partial)
{
    constexpr int CK = CF + 3, T = CO * (CK + 1);
    __shared__ float sd[W1_TILE][CO + 1];
    __shared__ float sx[W1_TILE][CK + 2];
    const int t = threadIdx.x, o = t / (CK + 1), k = t - o * (CK + 1);
    float acc = 0.f;
    const long long ntp = (P + W1_TILE - 1) / W1_TILE, ntq = (M + W1_TILE - 1) / W1_TILE;
    for (long long tl = blockIdx.x; tl < ntp + ntq; tl += gridDim.x) {
        const bool pts = tl < ntp;
        const long long base = (pts ? tl : tl - ntp) * W1_TILE, lim = pts ? P : M;
        const float *d = pts ? du : dc;
        __syncthreads();
        for (int x = t; x < W1_TILE * CO; x += T) {
            const int r = x / CO, cc = x - r * CO;
            sd[r][cc] = base + r < lim ? __ldg(d + (base + r) * CO + cc) : 0.f;
        }
        for (int x = t; x < W1_TILE * (CK + 1); x += T) {
            const int r = x / (CK + 1), cc = x - r * (CK + 1);
            float v = 0.f;
            if (base + r < lim) {
                if (pts) {
                    if (cc < CF) v = __ldg(feat + (base + r) * CF + cc);
                    else if (cc < CK) { const float4 pp = __ldg(pos + base + r); v = cc == CF ? pp.x : (cc == CF + 1 ? pp.y : pp.z); }
                    else v = 1.f;
                } else if (cc >= CF && cc < CK) {
                    const float4 qq = __ldg(qpos + base + r);
                    v = -(cc == CF ? qq.x : (cc == CF + 1 ? qq.y : qq.z));
                }
            }
            sx[r][cc] = v;
        }
        __syncthreads();
#pragma unroll 8
        for (int r = 0; r < W1_TILE; ++r) acc = fmaf(sd[r][o], sx[r][k], acc);
    }
    partial[(size_t)blockIdx.x * T + t] = acc;
}

template <int CO, int CF>
__global__ void __launch_bounds__(256)
sa_t_w1_reduce_kernel(const float *__restrict__ partial, int nblk, float *__restrict__ dW1, float *__restrict__ db1)
{
    constexpr int CK = CF + 3, T = CO * (CK + 1);
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (t >= T) return;
    float s = 0.f;
    for (int b = lane; b < nblk; b += 32) s += __ldg(partial + (size_t)b * T + t);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(SN2_FULL, s, d);
    if (lane == 0) {
        const int o = t / (CK + 1), k = t - o * (CK + 1);
        if (k < CK) dW1[o * CK + k] = s;
        else db1[o] = s;
    }
}

template <int CO, int CF>
int launch_w1(const float *du, const float *dc, const float *feat, const float *pos4, const float *qpos4, long long P, long long M,
              float *partial, float *dW, float *db, cudaStream_t st)
{
    constexpr int T = CO * (CF + 4);
    const int grid = 148 * 6;
    sa_t_w1_kernel<CO, CF><<<grid, T, 0, st>>>(du, dc, feat, reinterpret_cast<const float4 *>(pos4), reinterpret_cast<const float4 *>(qpos4), P, M,
                                              partial);
    SN2_LAUNCH_CHECK("sa_t_w1_kernel");
    sa_t_w1_reduce_kernel<CO, CF><<<(T * 32 + 255) / 256, 256, 0, st>>>(partial, grid, dW, db);
    SN2_LAUNCH_CHECK("sa_t_w1_reduce_kernel");
    return SN2_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// sa2: PointConv(local_nn = MLP([19, 32])), ONE block: a = relu(u_j + c_i), x2 = max_e BN(a).  lane = channel (32): a
// neighbour's u row is one coalesced 128-byte read, the running statistics / arg-max / row sums are per-lane scalars and
// nothing is reduced across lanes.  MODE 0: forward sweep (sum a, sum a^2, arg-max edge of sign(gamma) * a);
// MODE 1: backward sweep (dz1 = relu'(a) BN'(dz): du[col] += dz1, dc_i = row sum).
// ---------------------------------------------------------------------------------------------------------------
constexpr int T2C = 32, T2F = 16, T2K = T2F + 3;
constexpr int T2_CHUNK = 2;  // few, long rows (20 000 centroids x ~100 edges at config 3): small chunks balance the warps
struct Ts2Args {
    const float *u;
    const float4 *qpos;
    const int *rowptr, *col;
    int M;
    const float *W, *b, *gamma, *ss;   // live parameters [32][19], [32], [32]; ss [4*32]
    const double *stats, *sums;
    const float *dout;
    const int *arg;
    double *stats_out;
    float *key;
    int *arg_out;
    float *du, *dc;
    int *queue;
};

template <int MODE>
__global__ void __launch_bounds__(TS_T)
sa2t_sweep_kernel(const Ts2Args a)
{
    __shared__ double red[2 * T2C];
    const int tid = threadIdx.x, lane = tid & 31;
    if (tid < 2 * T2C) red[tid] = 0.0;
    __syncthreads();
    const float w0 = __ldg(a.W + lane * T2K + T2F), w1 = __ldg(a.W + lane * T2K + T2F + 1), w2 = __ldg(a.W + lane * T2K + T2F + 2);
    const float bb = __ldg(a.b + lane);
    const float sgn = __ldg(a.gamma + lane) < 0.f ? -1.f : 1.f;
    float cA = 0.f, cB = 0.f, cC = 0.f;
    if constexpr (MODE == 1) bn_bwd_coeff<T2C>(a.ss, a.sums, a.stats[2 * T2C], lane, cA, cB, cC);
    double sad = 0.0, sqd = 0.0;
    for (;;) {
        int base = 0;
        if (lane == 0) base = atomicAdd(a.queue, T2_CHUNK);
        base = __shfl_sync(SN2_FULL, base, 0);
        if (base >= a.M) break;
        const int rp_l = (lane <= T2_CHUNK && base + lane <= a.M) ? __ldg(a.rowptr + base + lane) : 0;
        const float4 q_l = (lane < T2_CHUNK && base + lane < a.M) ? __ldg(a.qpos + base + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
        float sa = 0.f, sq = 0.f;
        for (int ci = 0; ci < T2_CHUNK && base + ci < a.M; ++ci) {
            const int i = base + ci;
            const int s = __shfl_sync(SN2_FULL, rp_l, ci), e = __shfl_sync(SN2_FULL, rp_l, ci + 1);
            const float qx = __shfl_sync(SN2_FULL, q_l.x, ci), qy = __shfl_sync(SN2_FULL, q_l.y, ci), qz = __shfl_sync(SN2_FULL, q_l.z, ci);
            const float c = bb - (w0 * qx + w1 * qy + w2 * qz);
            float best = -INFINITY, dcs = 0.f, dl = 0.f;
            int beste = 0x7fffffff, argl = -1;
            if constexpr (MODE == 1) {
                argl = __ldg(a.arg + (size_t)i * T2C + lane);
                dl = __ldg(a.dout + (size_t)i * T2C + lane);
            }
            auto edge = [&](int p, float v, int j) {
                const float av = fmaxf(v + c, 0.f);
                if constexpr (MODE == 0) {
                    sa += av;
                    sq = fmaf(av, av, sq);
                    const float kv = sgn * av;
                    if (kv > best) { best = kv; beste = j; }  // ascending j: the first edge stays on ties
                } else {
                    const float dz = argl == j ? dl : 0.f;
                    const float da = av > 0.f ? fmaf(cA, dz, fmaf(cB, av, cC)) : 0.f;
                    atomicAdd(a.du + (size_t)p * T2C + lane, da);
                    dcs += da;
                }
            };
            for (int j0 = s; j0 < e; j0 += 32) {
                const int idl = j0 + lane < e ? __ldg(a.col + j0 + lane) : 0;
                const int take = min(32, e - j0);
                // 16 neighbour rows in flight per warp: the loop is a chain of L2 round trips otherwise
                constexpr int UN = 16;
                for (int t = 0; t < take; t += UN) {
                    int pp[UN];
                    float vv[UN];
#pragma unroll
                    for (int k = 0; k < UN; ++k) {
                        pp[k] = __shfl_sync(SN2_FULL, idl, (t + k) & 31);
                        vv[k] = t + k < take ? __ldg(a.u + (size_t)pp[k] * T2C + lane) : 0.f;  // warp-uniform predicate
                    }
#pragma unroll
                    for (int k = 0; k < UN; ++k)
                        if (t + k < take) edge(pp[k], vv[k], j0 + t + k);
                }
            }
            if constexpr (MODE == 0) {
                a.key[(size_t)i * T2C + lane] = best;
                a.arg_out[(size_t)i * T2C + lane] = e > s ? beste : -1;
            } else {
                a.dc[(size_t)i * T2C + lane] = dcs;
            }
        }
        sad += (double)sa;
        sqd += (double)sq;
    }
    if constexpr (MODE == 0) {
        atomicAdd(red + lane, sad);
        atomicAdd(red + T2C + lane, sqd);
        __syncthreads();
        if (tid < 2 * T2C) atomicAdd(a.stats_out + tid, red[tid]);
    }
}

inline int sweep_grid(int M, int per_sm = 2) { return (int)min((long long)148 * per_sm, ((long long)M + TS_WARPS - 1) / TS_WARPS); }

template <int MODE>
int launch_sweep(const TsArgs &a, int grid, cudaStream_t st)
{
    sa1t_prep_kernel<MODE><<<1, TS_T, 0, st>>>(a);
    SN2_LAUNCH_CHECK("sa1t_prep_kernel");
    void *stage = nullptr;
    SN2_CUDA_TRY(cudaGetSymbolAddress(&stage, g_ts_stage), "sa1t stage address");
    SN2_CUDA_TRY(cudaMemcpyToSymbolAsync(c_ts, stage, sizeof(TsW), 0, cudaMemcpyDeviceToDevice, st), "sa1t constant image");
    sa1t_sweep_kernel<MODE><<<grid, TS_T, 0, st>>>(a);
    SN2_LAUNCH_CHECK("sa1t_sweep_kernel");
    return SN2_OK;
}

}  // namespace
}  // namespace sn2

using namespace sn2;

extern "C" int sn2_sa1t_partials(void) { return T2C * (T2K + 1); }  // the largest CTA partial: sa2's [32][19 | db]
extern "C" int sn2_sa1t_blocks(void) { return 148 * 6; }

extern "C" int sn2_sa1t_pre(const float *feat, const float *pos4, long long P, const float *W1, float *u, void *stream)
{
    if (!feat || !pos4 || !W1 || !u || P <= 0) return SN2_EINVAL;
    sa_t_pre_kernel<TF, TC><<<(unsigned)((P + 255) / 256), 256, 0, (cudaStream_t)stream>>>(feat, reinterpret_cast<const float4 *>(pos4), P, W1, u);
    SN2_LAUNCH_CHECK("sa_t_pre_kernel");
    return SN2_OK;
}

extern "C" int sn2_sa1t_stats1(const float *u, const float *qpos4, const int *rowptr, const int *col, int M, const float *W1,
                               const float *b1, double *stats1, int *queue, void *stream)
{
    if (!u || !qpos4 || !rowptr || !col || !W1 || !b1 || !stats1 || !queue || M <= 0) return SN2_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    sa1t_stats_init_kernel<<<1, 128, 0, st>>>(stats1, rowptr, M, queue, TC);
    TsArgs a = {};
    a.u = u; a.qpos = reinterpret_cast<const float4 *>(qpos4); a.rowptr = rowptr; a.col = col; a.M = M;
    a.W1 = W1; a.b1 = b1; a.stats_out = stats1; a.queue = queue;
    return launch_sweep<0>(a, sweep_grid(M, 3), st);
}

extern "C" int sn2_sa1t_stats2(const float *u, const float *qpos4, const int *rowptr, const int *col, int M, const float *W1,
                               const float *b1, const float *W2, const float *b2, const float *ss1, const float *gamma2,
                               double *stats2, float *key, int *arg, int *queue, void *stream)
{
    if (!u || !qpos4 || !rowptr || !col || !W1 || !b1 || !W2 || !b2 || !ss1 || !gamma2 || !stats2 || !key || !arg || !queue || M <= 0)
        return SN2_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    sa1t_stats_init_kernel<<<1, 128, 0, st>>>(stats2, rowptr, M, queue, TC);
    TsArgs a = {};
    a.u = u; a.qpos = reinterpret_cast<const float4 *>(qpos4); a.rowptr = rowptr; a.col = col; a.M = M;
    a.W1 = W1; a.b1 = b1; a.W2 = W2; a.b2 = b2; a.ss1 = ss1; a.gamma2 = gamma2; a.stats_out = stats2; a.key = key; a.arg_out = arg; a.queue = queue;
    return launch_sweep<1>(a, sweep_grid(M), st);
}

extern "C" int sn2_sa1t_finish(const float *key, const int *arg, const float *gamma2, const float *ss2, int M, float *x1,
                               float *amax, void *stream)
{
    if (!key || !arg || !gamma2 || !ss2 || !x1 || !amax || M <= 0) return SN2_EINVAL;
    const long long n = (long long)M * TC;
    sa_t_finish_kernel<TC><<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(key, arg, gamma2, ss2, n, x1, amax);
    SN2_LAUNCH_CHECK("sa1t_finish_kernel");
    return SN2_OK;
}

extern "C" int sn2_sa1t_bwd_sums(const float *dout, const float *amax, const int *arg, int M, double *sums2, void *stream)
{
    if (!dout || !amax || !arg || !sums2 || M <= 0) return SN2_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    SN2_CUDA_TRY(cudaMemsetAsync(sums2, 0, 2 * TC * sizeof(double), st), "sa1t sums2 memset");
    sa_t_bwd_sums_kernel<TC><<<148, 256, 0, st>>>(dout, amax, arg, M, sums2);
    SN2_LAUNCH_CHECK("sa1t_bwd_sums_kernel");
    return SN2_OK;
}

extern "C" int sn2_sa1t_bwd_w2(const float *u, const float *qpos4, const int *rowptr, const int *col, int M, const float *W1,
                               const float *b1, const float *W2, const float *b2, const float *gamma2, const float *ss1,
                               const float *ss2, const double *stats2, const double *sums2, const float *dout, const int *arg,
                               float *partial, float *dW2, float *db2, double *sums1, int *queue, void *stream)
{
    if (!u || !qpos4 || !rowptr || !col || !W1 || !b1 || !W2 || !b2 || !gamma2 || !ss1 || !ss2 || !stats2 || !sums2 || !dout ||
        !arg || !partial || !dW2 || !db2 || !sums1 || !queue || M <= 0)
        return SN2_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    TsArgs a = {};
    a.u = u; a.qpos = reinterpret_cast<const float4 *>(qpos4); a.rowptr = rowptr; a.col = col; a.M = M;
    a.W1 = W1; a.b1 = b1; a.W2 = W2; a.b2 = b2; a.gamma2 = gamma2; a.ss1 = ss1; a.ss2 = ss2; a.stats2 = stats2; a.sums2 = sums2;
    a.dout = dout; a.arg = arg; a.partial = partial; a.queue = queue;
    SN2_CUDA_TRY(cudaMemsetAsync(queue, 0, sizeof(int), st), "sa1t queue memset");
    const int grid = sweep_grid(M);
    if (int rc = launch_sweep<3>(a, grid, st)) return rc;
    sa1t_w2_finish_kernel<<<1, TS_NP2, 0, st>>>(partial, grid, W2, ss1, dW2, db2, sums1);
    SN2_LAUNCH_CHECK("sa1t_w2_finish_kernel");
    return SN2_OK;
}

extern "C" int sn2_sa1t_bwd_in(const float *u, const float *qpos4, const int *rowptr, const int *col, long long P, int M,
                               const float *W1, const float *b1, const float *W2, const float *b2, const float *gamma2,
                               const float *ss1, const float *ss2, const double *stats1, const double *stats2,
                               const double *sums1, const double *sums2, const float *dout, const int *arg, float *du, float *dc,
                               int *queue, void *stream)
{
    if (!u || !qpos4 || !rowptr || !col || !W1 || !b1 || !W2 || !b2 || !gamma2 || !ss1 || !ss2 || !stats1 || !stats2 || !sums1 ||
        !sums2 || !dout || !arg || !du || !dc || !queue || M <= 0 || P <= 0)
        return SN2_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    SN2_CUDA_TRY(cudaMemsetAsync(du, 0, (size_t)P * TC * sizeof(float), st), "sa1t du memset");
    TsArgs a = {};
    a.u = u; a.qpos = reinterpret_cast<const float4 *>(qpos4); a.rowptr = rowptr; a.col = col; a.M = M;
    a.W1 = W1; a.b1 = b1; a.W2 = W2; a.b2 = b2; a.gamma2 = gamma2; a.ss1 = ss1; a.ss2 = ss2; a.stats1 = stats1; a.stats2 = stats2;
    a.sums1 = sums1; a.sums2 = sums2; a.dout = dout; a.arg = arg; a.du = du; a.dc = dc; a.queue = queue;
    SN2_CUDA_TRY(cudaMemsetAsync(queue, 0, sizeof(int), st), "sa1t queue memset");
    return launch_sweep<4>(a, sweep_grid(M), st);
}

extern "C" int sn2_sa1t_bwd_w1(const float *du, const float *dc, const float *feat, const float *pos4, const float *qpos4,
                               long long P, int M, float *partial, float *dW1, float *db1, void *stream)
{
    if (!du || !dc || !feat || !pos4 || !qpos4 || !partial || !dW1 || !db1 || P <= 0 || M <= 0) return SN2_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    return launch_w1<TC, TF>(du, dc, feat, pos4, qpos4, P, (long long)M, partial, dW1, db1, st);
}

// ---- sa2 -------------------------------------------------------------------------------------------------------
extern "C" int sn2_sa2t_pre(const float *x, const float *pos4, long long P, const float *W, float *u, void *stream)
{
    if (!x || !pos4 || !W || !u || P <= 0) return SN2_EINVAL;
    sa_t_pre_kernel<T2F, T2C><<<(unsigned)((P + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, reinterpret_cast<const float4 *>(pos4), P, W, u);
    SN2_LAUNCH_CHECK("sa_t_pre_kernel<2>");
    return SN2_OK;
}

extern "C" int sn2_sa2t_fwd(const float *u, const float *qpos4, const int *rowptr, const int *col, int M, const float *W,
                            const float *b, const float *gamma, double *stats, float *key, int *arg, int *queue, void *stream)
{
    if (!u || !qpos4 || !rowptr || !col || !W || !b || !gamma || !stats || !key || !arg || !queue || M <= 0) return SN2_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    sa1t_stats_init_kernel<<<1, 128, 0, st>>>(stats, rowptr, M, queue, T2C);
    Ts2Args a = {};
    a.u = u; a.qpos = reinterpret_cast<const float4 *>(qpos4); a.rowptr = rowptr; a.col = col; a.M = M;
    a.W = W; a.b = b; a.gamma = gamma; a.stats_out = stats; a.key = key; a.arg_out = arg; a.queue = queue;
    sa2t_sweep_kernel<0><<<sweep_grid(M, 6), TS_T, 0, st>>>(a);
    SN2_LAUNCH_CHECK("sa2t_sweep_kernel<fwd>");
    return SN2_OK;
}

extern "C" int sn2_sa2t_finish(const float *key, const int *arg, const float *gamma, const float *ss, int M, float *x2, float *amax,
                               void *stream)
{
    if (!key || !arg || !gamma || !ss || !x2 || !amax || M <= 0) return SN2_EINVAL;
    const long long n = (long long)M * T2C;
    sa_t_finish_kernel<T2C><<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(key, arg, gamma, ss, n, x2, amax);
    SN2_LAUNCH_CHECK("sa_t_finish_kernel<2>");
    return SN2_OK;
}

extern "C" int sn2_sa2t_bwd_sums(const float *dout, const float *amax, const int *arg, int M, double *sums, void *stream)
{
    if (!dout || !amax || !arg || !sums || M <= 0) return SN2_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    SN2_CUDA_TRY(cudaMemsetAsync(sums, 0, 2 * T2C * sizeof(double), st), "sa2t sums memset");
    sa_t_bwd_sums_kernel<T2C><<<148, 256, 0, st>>>(dout, amax, arg, M, sums);
    SN2_LAUNCH_CHECK("sa_t_bwd_sums_kernel<2>");
    return SN2_OK;
}

extern "C" int sn2_sa2t_bwd(const float *u, const float *qpos4, const int *rowptr, const int *col, long long P, int M,
                            const float *W, const float *b, const float *gamma, const float *ss, const double *stats,
                            const double *sums, const float *dout, const int *arg, float *du, float *dc, int *queue, void *stream)
{
    if (!u || !qpos4 || !rowptr || !col || !W || !b || !gamma || !ss || !stats || !sums || !dout || !arg || !du || !dc || !queue ||
        M <= 0 || P <= 0)
        return SN2_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    SN2_CUDA_TRY(cudaMemsetAsync(du, 0, (size_t)P * T2C * sizeof(float), st), "sa2t du memset");
    SN2_CUDA_TRY(cudaMemsetAsync(queue, 0, sizeof(int), st), "sa2t queue memset");
    Ts2Args a = {};
    a.u = u; a.qpos = reinterpret_cast<const float4 *>(qpos4); a.rowptr = rowptr; a.col = col; a.M = M;
    a.W = W; a.b = b; a.gamma = gamma; a.ss = ss; a.stats = stats; a.sums = sums; a.dout = dout; a.arg = arg; a.du = du; a.dc = dc;
    a.queue = queue;
    sa2t_sweep_kernel<1><<<sweep_grid(M, 6), TS_T, 0, st>>>(a);
    SN2_LAUNCH_CHECK("sa2t_sweep_kernel<bwd>");
    return SN2_OK;
}

extern "C" int sn2_sa2t_bwd_w(const float *du, const float *dc, const float *x, const float *pos4, const float *qpos4, long long P,
                              int M, float *partial, float *dW, float *db, void *stream)
{
    if (!du || !dc || !x || !pos4 || !qpos4 || !partial || !dW || !db || P <= 0 || M <= 0) return SN2_EINVAL;
    return launch_w1<T2C, T2F>(du, dc, x, pos4, qpos4, P, (long long)M, partial, dW, db, (cudaStream_t)stream);
}
