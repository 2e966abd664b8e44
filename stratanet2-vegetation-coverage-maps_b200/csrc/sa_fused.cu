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
#include <stdlib.h>

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

// ---- tensor-core (tcgen05) helpers for the level-1 second layer ---------------------------------------
// K-major, no-swizzle canonical operand layout: core matrix = 8 rows x 16 bytes (128 contiguous bytes);
// LBO = bytes between core matrices adjacent in K, SBO = between core matrices adjacent in M/N.
// Element (row r, k) of a [rows x 16] fp32 operand sits at (r/8)*512 + (k/4)*128 + (r%8)*16 + (k%4)*4.
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned long long umma_desc(unsigned saddr)
{
    return (unsigned long long)((saddr >> 4) & 0x3fff) | ((unsigned long long)(128 >> 4) << 16) |
           ((unsigned long long)(512 >> 4) << 32) | (1ull << 46);  // LBO 128 B, SBO 512 B, version 1, SWIZZLE_NONE
}
// kind::tf32, D = F32, A/B = TF32 K-major, N = 16, M = 128
constexpr unsigned UMMA_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
__device__ __forceinline__ void umma_tf32(unsigned tmem_d, unsigned long long da, unsigned long long db, unsigned acc)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}\n" ::"r"(tmem_d), "l"(da),
                 "l"(db), "r"(UMMA_IDESC), "r"(acc), "r"(0), "r"(0), "r"(0), "r"(0));
}
// Packed fp32 FMA (Blackwell FFMA2): two IEEE fmas per issue slot, bit-identical to two FFMAs.  The level-1 kernel is
// bound by instruction issue (ncu: issue active 79 %, FMA pipe 37 %), so its 16 x 16 second layer was tried as 128 FFMA2
// with the weight pairs read from shared memory (64 LDS.128 per 32 edges) instead of 256 FFMA + 64 uniform constant
// loads.  Measured SLOWER (SA1 at config 2: 1.13 ms vs 0.92 ms, results identical): FFMA2 takes register pairs only, and
// the broadcast LDS.128 stream plus the pair moves cost more than the issue slots they free.  Kept as a compile-time
// switch for re-measurement; the product build uses 0.
#ifndef SN2_SA1_FFMA2
#define SN2_SA1_FFMA2 0
#endif
typedef unsigned long long u64_t;
__device__ __forceinline__ u64_t pack_f32x2(float a, float b)
{
    u64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ u64_t fma_f32x2(u64_t a, u64_t b, u64_t c)
{
    u64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ void unpack_f32x2(u64_t v, float &a, float &b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }

template <int LEVEL, bool REDO, int DBG = 0>
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
    unsigned *bm = reinterpret_cast<unsigned *>(sf_smem + SF_WARPS * RING * sizeof(int)) + (size_t)warp * words;  // REDO only
    // streaming only: per-warp table of the <= 9 candidate rows of the current centroid: [k] = inclusive prefix of the row
    // lengths (flat candidate index space), [16 + k] = offset from a flat index inside row k to its position in `sorted`
    int *rtab = reinterpret_cast<int *>(sf_smem + SF_WARPS * RING * sizeof(int)) + warp * 32;
    const unsigned lt = (1u << lane) - 1u;
    // level 1: second-layer weights [k][o] + bias as fp32 pairs for the packed FMA
    __shared__ __align__(16) float s_w2[(LEVEL == 1 && SN2_SA1_FFMA2) ? SN2_C1 * SN2_C1 + SN2_C1 : 4];
    if constexpr (LEVEL == 1 && SN2_SA1_FFMA2) {
        for (int i = threadIdx.x; i < SN2_C1 * SN2_C1 + SN2_C1; i += SF_WARPS * 32)
            s_w2[i] = i < SN2_C1 * SN2_C1 ? W.l2.w[i / SN2_C1][i % SN2_C1] : W.l2.b[i - SN2_C1 * SN2_C1];
        __syncthreads();
    }

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
            if (j >= M) break;
        }
        const float *hdr = grid_hdr + (size_t)b * SN2_GRID_HDR;
        const int *cs = cell_start + (size_t)b * (SN2_GRID_CELLS + 1);
        const float4 *so = sorted + (size_t)b * N;
        const float4 q = __ldg(qsorted + (size_t)b * M + j);
        const int qloc = __float_as_int(q.w);
        const float *ub = u + (size_t)b * N * C;

        const float ox = hdr[0], oy = hdr[1], inv = hdr[2], oz = hdr[6], invz = hdr[7];
        const int gx = __float_as_int(hdr[4]), gy = __float_as_int(hdr[5]), gz = __float_as_int(hdr[8]);
        const int ix = cell_coord_c(q.x, ox, inv, gx), iy = cell_coord_c(q.y, oy, inv, gy);
        const int iz = cell_coord_c(q.z, oz, invz, gz);
        const int x0 = max(ix - 1, 0), x1 = min(ix + 1, gx - 1);
        const int y0 = max(iy - 1, 0), y1 = min(iy + 1, gy - 1);
        const int z0 = max(iz - 1, 0), z1 = min(iz + 1, gz - 1);
        const int ny = y1 - y0 + 1, nrows = ny * (z1 - z0 + 1);  // (layer, cell row) pairs to scan
        auto row_of = [&](int t) { return ((z0 + t / ny) * gy + (y0 + t % ny)) * gx; };

        // c_i = b1 - W1p q  (the per-centroid half of the first layer)
        // level 1: every lane keeps all C channels (lane = edge); level 2: lane = channel (C == 32), so the
        // running max / min, c_i and the BN affine are per-lane scalars and u_j rows are read fully coalesced
        constexpr int CL_ = LEVEL == 1 ? C : 1;
        float c[CL_];
        float mx[CL_], mn[1];
        if constexpr (LEVEL == 1) {
#pragma unroll
            for (int o = 0; o < C; ++o)
                c[o] = W.l1.b[o] - (W.l1.w[E::CIN][o] * q.x + W.l1.w[E::CIN + 1][o] * q.y + W.l1.w[E::CIN + 2][o] * q.z);
#pragma unroll
            for (int o = 0; o < C; ++o) mx[o] = -INFINITY;
            mn[0] = 0.f;
        } else {
            static_assert(LEVEL == 1 || C == 32, "level 2 maps one channel to each lane");
            c[0] = W.l1.b[lane] - (W.l1.w[E::CIN][lane] * q.x + W.l1.w[E::CIN + 1][lane] * q.y + W.l1.w[E::CIN + 2][lane] * q.z);
            mx[0] = -INFINITY;
            mn[0] = INFINITY;
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
                // layer 1's BatchNorm affine is folded into layer 2 by the launcher: W2' = diag(s1) W2, b2' = b2 + t1 W2
#pragma unroll
                for (int k = 0; k < C; ++k) h1[k] = fmaxf(h1[k], 0.f);
                float h2[SN2_C1];
                if constexpr (SN2_SA1_FFMA2) {
                    u64_t a2[SN2_C1 / 2];
                    const ulonglong2 *bw = reinterpret_cast<const ulonglong2 *>(s_w2 + SN2_C1 * SN2_C1);
#pragma unroll
                    for (int g = 0; g < SN2_C1 / 4; ++g) {
                        const ulonglong2 b2 = bw[g];
                        a2[2 * g] = b2.x;
                        a2[2 * g + 1] = b2.y;
                    }
#pragma unroll
                    for (int k = 0; k < SN2_C1; ++k) {
                        const u64_t hk = pack_f32x2(h1[k], h1[k]);
                        const ulonglong2 *wr = reinterpret_cast<const ulonglong2 *>(s_w2 + SN2_C1 * k);
#pragma unroll
                        for (int g = 0; g < SN2_C1 / 4; ++g) {
                            const ulonglong2 w2 = wr[g];
                            a2[2 * g] = fma_f32x2(hk, w2.x, a2[2 * g]);
                            a2[2 * g + 1] = fma_f32x2(hk, w2.y, a2[2 * g + 1]);
                        }
                    }
#pragma unroll
                    for (int g = 0; g < SN2_C1 / 2; ++g) unpack_f32x2(a2[g], h2[2 * g], h2[2 * g + 1]);
                } else {
                    acc_init(W.l2, h2);
#pragma unroll
                    for (int k = 0; k < SN2_C1; ++k) acc_step<0>(W.l2, h1[k], h2, k);
                }
                // No ReLU / BatchNorm per edge: f(a) = fma(max(a, 0), s, t) is monotone in the pre-activation a, so
                // max_j f(a_j) = f(max_j a_j) for s >= 0 and f(min_j a_j) for s < 0.  The launcher negates the
                // second-layer weights and bias of the s < 0 channels (fma is odd: the accumulators come out exactly
                // negated), so a running max serves both signs; f is applied once per centroid at the end.
#pragma unroll
                for (int o = 0; o < C; ++o) mx[o] = fmaxf(mx[o], h2[o]);
            }
        };

        if (REDO) {  // hit set as a bitmap over the plot's point indices
            for (int w = lane; w < words; w += 32) bm[w] = 0u;
            __syncwarp();
            for (int t = 0; t < nrows; ++t) {
                const int s2 = __ldg(cs + row_of(t) + x0), e2 = __ldg(cs + row_of(t) + x1 + 1);
                for (int i = s2 + lane; i < e2; i += 32) {
                    const float4 v = __ldg(so + i);
                    if (dist2(v.x, v.y, v.z, q.x, q.y, q.z) < r2) {
                        const int id = __float_as_int(v.w);
                        atomicOr(&bm[id >> 5], 1u << (id & 31));
                    }
                }
            }
            __syncwarp();
        }

        // ---- producer / consumer loop with ONE call site of the message MLP (keeps the loop in I-cache) ----
        int cnt = 0;
        int head = 0, tail = 0;          // warp-uniform ring cursors
        // streaming iterator over the FLATTENED candidate space: the rows of a 3 x 3 x 3 cell block are short (18 candidates
        // on average, six of nine nearly empty), so walking them one by one left the 32-lane tests 44 % full and paid two
        // dependent cell_start loads, an integer division and ~100 instructions of loop control per row (ncu source page:
        // 37 % of the kernel's instructions).  Here lanes 0..8 fetch their row's range in parallel once, a prefix sum
        // turns the rows into one index space, and every 32-lane step maps flat index -> (row, position) with a
        // 9-entry table in shared memory.
        int f0 = 0, ftotal = 0;
        int mypre = 0x7fffffff, myoff = 0;  // lane k < (non-empty rows): inclusive end of row k in flat space, flat -> `sorted` offset
        if (!REDO) {
            int rs_ = 0, len_ = 0;
            if (lane < nrows) {
                // Row (cy, cz) trimmed to the cells the search sphere can reach: lower bounds dy, dz on the distance from the
                // centroid to the row's slab, chord half-width w = sqrt(r2 - dy^2 - dz^2) along x.  Conservative (margins
                // cover the rounding of the cell arithmetic), so the hit set is unchanged; one lane per row, in parallel.
                const int tz = lane / ny, ty = lane - tz * ny;
                const int cy = y0 + ty, cz = z0 + tz;
                const float csx = hdr[3], csz = hdr[9];
                const float ey = 1e-4f * csx + 1e-6f * (fabsf(q.y) + fabsf(oy) + csx * gy);
                const float ez = 1e-4f * csz + 1e-6f * (fabsf(q.z) + fabsf(oz) + csz * gz);
                float dy = cy > iy ? (oy + cy * csx) - q.y : (cy < iy ? q.y - (oy + (cy + 1) * csx) : 0.f);
                float dz = cz > iz ? (oz + cz * csz) - q.z : (cz < iz ? q.z - (oz + (cz + 1) * csz) : 0.f);
                dy = fmaxf(dy - ey, 0.f);
                dz = fmaxf(dz - ez, 0.f);
                const float rem = r2 * 1.00001f - dy * dy - dz * dz;
                if (rem > 0.f) {
                    const float w = sqrtf(rem) * 1.00001f + 1e-4f * csx + 1e-6f * (fabsf(q.x) + fabsf(ox) + csx * gx);
                    const int xa = max(x0, cell_coord_c(q.x - w, ox, inv, gx)), xb = min(x1, cell_coord_c(q.x + w, ox, inv, gx));
                    const int rb = (cz * gy + cy) * gx;
                    rs_ = __ldg(cs + rb + xa);
                    len_ = __ldg(cs + rb + xb + 1) - rs_;
                }
            }
            int pre_ = len_;
#pragma unroll
            for (int d = 1; d < 16; d <<= 1) {
                const int t_ = __shfl_up_sync(SN2_FULL, pre_, d);
                if (lane >= d) pre_ += t_;
            }
            ftotal = __shfl_sync(SN2_FULL, pre_, 15);  // lanes >= nrows carry the total (their rows are empty)
            // drop the empty rows: the ends of the remaining ones are strictly increasing, so a 32-candidate step can mark
            // them as distinct bits (below)
            const unsigned ne = __ballot_sync(SN2_FULL, len_ > 0);
            if (len_ > 0) {
                const int slot = __popc(ne & lt);
                rtab[slot] = pre_;
                rtab[16 + slot] = rs_ - (pre_ - len_);
            }
            __syncwarp();
            if (lane < __popc(ne)) {
                mypre = rtab[lane];
                myoff = rtab[16 + lane];
            }
            __syncwarp();
        }
        int emitted = 0, wblk = 0;       // redo iterator: set bits emitted so far, next 32-word block of the bitmap
        unsigned wmask = 0u, wword = 0u; //   non-empty words left in the current block, this lane's word of the block
        bool more = true;
        while (true) {
            while (more && tail - head < 32) {
                if (!REDO) {
                    if (f0 >= ftotal) { more = false; break; }
                    float4 vv[UNR];
#pragma unroll
                    for (int t = 0; t < UNR; ++t) {
                        // flat index -> row: r = number of rows that end at or before f = (rows ended before this step) +
                        // (row ends inside the step at or before this lane); lane k holds the end of row k
                        const int fb = f0 + t * 32, f = fb + lane;
                        vv[t] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (fb >= ftotal) continue;  // warp-uniform
                        const int pe = mypre - fb;
                        const unsigned done = __ballot_sync(SN2_FULL, pe <= 0);
                        const unsigned ends = __reduce_or_sync(SN2_FULL, (unsigned)(pe - 1) < 31u ? 1u << pe : 0u);
                        const int r = __popc(done) + __popc(ends & ((2u << lane) - 1u));
                        const int off = __shfl_sync(SN2_FULL, myoff, r);
                        if (f < ftotal) vv[t] = __ldg(so + f + off);
                    }
#pragma unroll
                    for (int t = 0; t < UNR; ++t) {
                        if (f0 + t * 32 < ftotal) {  // warp-uniform
                            const float4 v = vv[t];
                            const bool hit = (f0 + t * 32 + lane < ftotal) && dist2(v.x, v.y, v.z, q.x, q.y, q.z) < r2;
                            const unsigned bal = __ballot_sync(SN2_FULL, hit);
                            if (hit) ring[(tail + __popc(bal & lt)) & (RING - 1)] = __float_as_int(v.w);
                            tail += __popc(bal);
                        }
                    }
                    f0 += 32 * UNR;
                } else {
                    // first K set bits of the hit bitmap in ascending point index: 32-word blocks, empty words
                    // skipped by ballot, one word (<= 32 ids) per step, every lane places its own bit by popc rank
                    if (!wmask) {
                        if (emitted >= K || wblk >= words) { more = false; break; }
                        wword = wblk + lane < words ? bm[wblk + lane] : 0u;
                        wmask = __ballot_sync(SN2_FULL, wword != 0u);
                        if (!wmask) { wblk += 32; continue; }
                    }
                    const int wl = __ffs(wmask) - 1;
                    wmask &= wmask - 1;
                    const unsigned bits = __shfl_sync(SN2_FULL, wword, wl);
                    const int rk = __popc(bits & lt);
                    if (((bits >> lane) & 1u) && emitted + rk < K) ring[(tail + rk) & (RING - 1)] = ((wblk + wl) << 5) + lane;
                    const int n = min(__popc(bits), K - emitted);
                    tail += n;
                    emitted += n;
                    if (!wmask) wblk += 32;
                    if (emitted >= K) more = false;
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
            if (LEVEL == 2) {
                // lane = channel: one coalesced 128-byte read of u_j per neighbour, ids broadcast by shuffle
                int t = 0;
                for (; t + 4 <= take; t += 4) {
                    const float v0 = __ldg(ub + (size_t)__shfl_sync(SN2_FULL, id, t) * C + lane);
                    const float v1 = __ldg(ub + (size_t)__shfl_sync(SN2_FULL, id, t + 1) * C + lane);
                    const float v2 = __ldg(ub + (size_t)__shfl_sync(SN2_FULL, id, t + 2) * C + lane);
                    const float v3 = __ldg(ub + (size_t)__shfl_sync(SN2_FULL, id, t + 3) * C + lane);
                    mx[0] = fmaxf(fmaxf(mx[0], fmaxf(v0, v1)), fmaxf(v2, v3));
                    mn[0] = fminf(fminf(mn[0], fminf(v0, v1)), fminf(v2, v3));
                }
                for (; t < take; ++t) {
                    const float v0 = __ldg(ub + (size_t)__shfl_sync(SN2_FULL, id, t) * C + lane);
                    mx[0] = fmaxf(mx[0], v0);
                    mn[0] = fminf(mn[0], v0);
                }
            } else if (DBG != 1 && lane < take) {
                edge(id);  // DBG 1: search only (profiling)
            }
        }

        const size_t row = (size_t)b * M + qloc;
        if (!REDO && cnt > K) {  // the cap binds: leave this centroid to the exact redo launch
            if (lane == 0) ovf[1 + atomicAdd(ovf, 1)] = b * M + j;
            continue;
        }
        if constexpr (LEVEL == 1) {
            // max over the 32 lanes of all 16 channels by a transposing butterfly: every exchange halves the channels a
            // lane still carries (8 + 4 + 2 + 1 + 1 shuffles instead of 16 x 5); lanes 2c and 2c + 1 end up with channel c
            static_assert(C == 16, "butterfly below is written for 16 channels");
            float v8[8], v4[4], v2[2];
            {
                const bool hi = lane & 16;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float snd = hi ? mx[i] : mx[i + 8], keep = hi ? mx[i + 8] : mx[i];
                    v8[i] = fmaxf(keep, __shfl_xor_sync(SN2_FULL, snd, 16));
                }
            }
            {
                const bool hi = lane & 8;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float snd = hi ? v8[i] : v8[i + 4], keep = hi ? v8[i + 4] : v8[i];
                    v4[i] = fmaxf(keep, __shfl_xor_sync(SN2_FULL, snd, 8));
                }
            }
            {
                const bool hi = lane & 4;
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const float snd = hi ? v4[i] : v4[i + 2], keep = hi ? v4[i + 2] : v4[i];
                    v2[i] = fmaxf(keep, __shfl_xor_sync(SN2_FULL, snd, 4));
                }
            }
            float v1;
            {
                const bool hi = lane & 2;
                const float snd = hi ? v2[0] : v2[1], keep = hi ? v2[1] : v2[0];
                v1 = fmaxf(keep, __shfl_xor_sync(SN2_FULL, snd, 2));
            }
            v1 = fmaxf(v1, __shfl_xor_sync(SN2_FULL, v1, 1));
            // the kernel tracks the (sign-adjusted) pre-activation: finish it here (lane pair = channel)
            const int ch = lane >> 1;
            const float sc = W.l2.s[ch], a = sc < 0.f ? -v1 : v1;
            if (!(lane & 1)) out[row * C + ch] = cnt > 0 ? fmaf(fmaxf(a, 0.f), sc, W.l2.t[ch]) : 0.f;
        } else {
            // out = max_j f(u_j) = f(max u) if s >= 0 else f(min u): f is monotone in u (lane = channel)
            const float sc = W.l1.s[lane], sel = sc >= 0.f ? mx[0] : mn[0];
            out[row * C + lane] = cnt > 0 ? fmaf(fmaxf(sel + c[0], 0.f), sc, W.l1.t[lane]) : 0.f;
        }
        if (lane == 0) {
            if (cnt_out) cnt_out[row] = cnt;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Level 1 with the second layer (the one dense per-edge contraction left: [E,16] x [16,16]) on the tensor cores.
//
// Four warps (one warpgroup; a CTA holds two) share ONE M = 128 accumulator tile: warp r of the group owns rows
// 32r .. 32r+31 -- the TMEM lane quarter it is allowed to touch -- so every row of the tile is a real edge (the round-1
// variant gave each warp a private M = 128 tile: 3/4 of the rows were padding and the tensor core spent its time
// reading them from shared memory).  The A operand never passes through shared memory: each lane computes layer 1 of
// its edge in registers (u_j + c_i, ReLU, BN) and writes the 16 values of its row straight into tensor memory
// (tcgen05.st 32x32b.x16: lane = row, column = k); tcgen05.mma takes A from TMEM, B = W2^T (16 x 16, K-major,
// no swizzle) from shared memory, D into TMEM.  The warps of a group work on different centroids but advance in
// lock-step ROUNDS: produce a batch of <= 32 edges (search + ring as in the SIMT kernel), store the rows, meet at a
// named barrier (bar.red.or also tells the group whether anybody has edges left), one thread issues the MMAs and
// commits them to the stage's mbarrier, and the result is consumed one round later (two stages), so the MMA latency
// (~300 cycles) hides under the next batch's search.  The per-edge epilogue is 16 fmax: bias, ReLU and BatchNorm are
// applied once per centroid (monotone in the pre-activation; W2 columns with a negative BN scale are negated by the
// launcher, as in the SIMT kernel).
// MODE 1: 3xTF32 (A_hi B_hi + A_lo B_hi + A_hi B_lo, operands split into tf32 hi + lo): fp32-accurate, 6 MMAs / round.
// MODE 2: TF32, operands rounded to nearest (cvt.rna): what torch 1.8 + cuBLAS did on the reference's Ampere GPUs.
// MODE 3: operands rounded to BF16 precision (RNE) and multiplied on the same pipe -- bit-identical to a bf16 x bf16 ->
//         fp32 MMA (products of bf16 values are exact in the tf32 datapath).
// ---------------------------------------------------------------------------------------------------------------
constexpr int TC_RING = 256;
template <int MODE> struct TcLayout {
    static constexpr int A_COLS = MODE == 1 ? 32 : 16;         // A_hi (+ A_lo)
    static constexpr int STAGE = A_COLS + 16;                  // + D
    static constexpr int GROUP = 2 * STAGE;                    // two stages
    static constexpr int COLS = MODE == 1 ? 256 : 128;         // allocation (power of two) for the CTA's two groups
};

__device__ __forceinline__ void umma_tf32_ts(unsigned tmem_d, unsigned tmem_a, unsigned long long db, unsigned acc)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t}\n" ::"r"(tmem_d), "r"(tmem_a),
                 "l"(db), "r"(UMMA_IDESC), "r"(acc), "r"(0), "r"(0), "r"(0), "r"(0));
}
__device__ __forceinline__ void tmem_st16(unsigned addr, const float (&v)[16])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n" ::"r"(addr),
                 "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
                 "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
                 "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
                 "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
                 : "memory");
}
__device__ __forceinline__ float round_tf32(float x)
{
    unsigned r;
    asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ float round_bf16(float x)
{
    unsigned u = __float_as_uint(x);
    u += 0x7fffu + ((u >> 16) & 1u);  // round to nearest even on the upper 16 bits (finite inputs)
    return __uint_as_float(u & 0xffff0000u);
}
template <int MODE>
__device__ __forceinline__ float tc_operand(float x)
{
    if (MODE == 2) return round_tf32(x);
    if (MODE == 3) return round_bf16(x);
    return x;
}

// REDO = false: grid (M / 8, B), one centroid per warp (the streaming launch).  REDO = true: persistent launch over the
// overflow list (centroids whose hit count exceeds the cap K): hits -> bitmap over the plot's point indices -> the first
// K set bits in ascending index order (Appendix A2), exactly as the SIMT redo kernel, feeding the same rounds; a warp
// walks through its items back to back (the last batch of item A is consumed in the round that stores the first batch
// of item B, then A is finalised and the running max starts over).
template <int MODE, bool REDO, int MINB = 2>
__global__ void __launch_bounds__(SF_WARPS * 32, REDO ? 1 : MINB)
sa1_tc_kernel(const float *__restrict__ grid_hdr, const int *__restrict__ cell_start, const float4 *__restrict__ sorted,
              const float4 *__restrict__ qsorted, const float *__restrict__ u, int N, int M, float r2, int K, int words,
              const __grid_constant__ W_SA1 W, float *__restrict__ out, int *__restrict__ cnt_out, int *__restrict__ ovf, int B)
{
    using L = TcLayout<MODE>;
    constexpr int C = SN2_C1, UNR = 4;
    extern __shared__ __align__(16) unsigned char tc_dyn[];  // REDO: hit bitmaps [SF_WARPS][words]
    __shared__ int ring_all[SF_WARPS * TC_RING];
    __shared__ int rtab_all[SF_WARPS * 32];                 // per warp: flattened candidate rows (see sa_fused_kernel)
    __shared__ __align__(1024) float b_tiles[2 * 256];      // W2^T hi | lo, canonical K-major no-swizzle layout
    __shared__ __align__(8) unsigned long long bars[2 * 2];  // [group][stage]
    __shared__ unsigned s_tmem;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, grp = warp >> 2, rq = warp & 3;
    int *ring = ring_all + warp * TC_RING;
    int *rtab = rtab_all + warp * 32;
    unsigned *bm = reinterpret_cast<unsigned *>(tc_dyn) + (size_t)warp * words;
    const unsigned lt = (1u << lane) - 1u;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&s_tmem)), "r"(L::COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    if (threadIdx.x < 4) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bars[threadIdx.x])));
    {   // B operand [N = 16 (o) x K = 16 (k)]: element (o, k) at (o/8)*128 + (k/4)*32 + (o%8)*4 + k%4 floats
        const int o = threadIdx.x >> 4, k = threadIdx.x & 15;
        const float wv = tc_operand<MODE>(W.l2.w[k][o]);
        const float hi = __uint_as_float(__float_as_uint(wv) & 0xffffe000u);
        const int off = (o >> 3) * 128 + (k >> 2) * 32 + (o & 7) * 4 + (k & 3);
        b_tiles[off] = MODE == 1 ? hi : wv;
        b_tiles[256 + off] = MODE == 1 ? wv - hi : 0.f;
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n");
    const unsigned tbase = s_tmem + (unsigned)(grp * L::GROUP);             // this group's columns, lane 0
    const unsigned tmine = tbase + ((unsigned)(32 * rq) << 16);             // ... seen from this warp's lane quarter
    const unsigned long long d_bhi = umma_desc(smem_u32(b_tiles)), d_blo = umma_desc(smem_u32(b_tiles + 256));
    unsigned parity = 0;  // bit s: phase parity of this group's stage-s mbarrier

    // ---- work items of this warp ----
    // REDO: entries of the overflow list, strided over all warps of the grid.  Streaming: the CTA owns a contiguous chunk of
    // the (plot, cell-ordered centroid) sequence and its warps pull the next centroid from a shared counter, so a warp that
    // finishes a small neighbourhood (median 15 edges, p90 260) keeps feeding rows into its group's rounds instead of idling
    // until the largest neighbourhood of the group is done.
    __shared__ int s_next;
    const long long total = (long long)B * M;
    const long long chunk = (total + gridDim.x - 1) / gridDim.x;
    const long long q_end = REDO ? 0 : min(total, (long long)(blockIdx.x + 1) * chunk);
    if (!REDO && threadIdx.x == 0) s_next = (int)min(total, (long long)blockIdx.x * chunk);
    long long item = REDO ? (long long)blockIdx.x * SF_WARPS + warp : 0;
    const long long item_step = REDO ? (long long)gridDim.x * SF_WARPS : 0;
    const long long n_items = REDO ? (long long)ovf[0] : 0;
    __syncthreads();

    // current item
    bool have = false, more = false;
    int b = 0, j = 0, qloc = 0, x0 = 0, x1 = 0, y0 = 0, z0 = 0, ny = 1, nrows = 0, gx = 1, gy = 1;
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
    const int *cs = cell_start;
    const float4 *so = sorted;
    const float *ub = u;
    int cnt = 0, head = 0, tail = 0, f0 = 0, ftotal = 0;
    int emitted = 0, wblk = 0;
    unsigned wmask = 0u, wword = 0u;
    float c[C], mx[C];
#pragma unroll
    for (int o = 0; o < C; ++o) { c[o] = 0.f; mx[o] = -INFINITY; }
    auto row_of = [&](int t) { return ((z0 + t / ny) * gy + (y0 + t % ny)) * gx; };

    // batch in flight (stored last round, consumed this round)
    bool pending = false, p_last = false;
    int p_take = 0, p_cnt = 0, p_code = 0;
    long long p_row = 0;
    int stage = 0;

    auto consume = [&](const int take, const int st) {
        unsigned done = 0;
        for (int spin = 0; spin < (1 << 24) && !done; ++spin)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                         : "=r"(done) : "r"(smem_u32(&bars[2 * grp + st])), "r"((parity >> st) & 1u) : "memory");
        if (!done) __trap();  // the MMA never completed: fail loudly instead of hanging the GPU
        parity ^= 1u << st;
        asm volatile("tcgen05.fence::after_thread_sync;\n");
        unsigned d[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                     : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7]),
                       "=r"(d[8]), "=r"(d[9]), "=r"(d[10]), "=r"(d[11]), "=r"(d[12]), "=r"(d[13]), "=r"(d[14]), "=r"(d[15])
                     : "r"(tmine + (unsigned)(st * L::STAGE + L::A_COLS)));
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;\n");
        if (lane < take) {
#pragma unroll
            for (int o = 0; o < C; ++o) mx[o] = fmaxf(mx[o], __uint_as_float(d[o]));  // raw (sign-adjusted) pre-activation, bias later
        }
    };
    // the item whose last batch has just been consumed: bias / ReLU / BatchNorm once, write the row, start the max over
    auto finalize = [&]() {
        if (!REDO && p_cnt > K) {  // the cap binds: leave this centroid to the exact redo launch
            if (lane == 0) ovf[1 + atomicAdd(ovf, 1)] = p_code;
        } else {
            float r[C];
#pragma unroll
            for (int o = 0; o < C; ++o) {
                const float m = warp_max(mx[o]) + W.l2.b[o];  // bias of the (sign-adjusted) second layer, once per centroid
                const float sc = W.l2.s[o], a = sc < 0.f ? -m : m;
                r[o] = p_cnt > 0 ? fmaf(fmaxf(a, 0.f), sc, W.l2.t[o]) : 0.f;
            }
            if (lane == 0) {
                float4 *o4 = reinterpret_cast<float4 *>(out + p_row * C);
#pragma unroll
                for (int v = 0; v < C / 4; ++v) o4[v] = make_float4(r[4 * v], r[4 * v + 1], r[4 * v + 2], r[4 * v + 3]);
                if (cnt_out) cnt_out[p_row] = p_cnt;
            }
        }
#pragma unroll
        for (int o = 0; o < C; ++o) mx[o] = -INFINITY;
    };

    while (true) {
        // ---- next item of this warp ----
        long long code_next = -1;
        if (!have) {
            if (REDO) {
                if (item < n_items) code_next = ovf[1 + item];
                item += item_step;
            } else if (*reinterpret_cast<volatile int *>(&s_next) < q_end) {
                int t = 0;
                if (lane == 0) t = atomicAdd(&s_next, 1);
                t = __shfl_sync(SN2_FULL, t, 0);
                if (t < q_end) code_next = t;
            }
        }
        if (code_next >= 0) {
            b = (int)(code_next / M);
            j = (int)(code_next - (long long)b * M);
            const float *hdr = grid_hdr + (size_t)b * SN2_GRID_HDR;
            cs = cell_start + (size_t)b * (SN2_GRID_CELLS + 1);
            so = sorted + (size_t)b * N;
            q = __ldg(qsorted + (size_t)b * M + j);
            qloc = __float_as_int(q.w);
            ub = u + (size_t)b * N * C;
            const float ox = hdr[0], oy = hdr[1], inv = hdr[2], oz = hdr[6], invz = hdr[7];
            gx = __float_as_int(hdr[4]);
            gy = __float_as_int(hdr[5]);
            const int gz = __float_as_int(hdr[8]);
            const int ix = cell_coord_c(q.x, ox, inv, gx), iy = cell_coord_c(q.y, oy, inv, gy), iz = cell_coord_c(q.z, oz, invz, gz);
            x0 = max(ix - 1, 0); x1 = min(ix + 1, gx - 1);
            y0 = max(iy - 1, 0);
            const int y1 = min(iy + 1, gy - 1);
            z0 = max(iz - 1, 0);
            const int z1 = min(iz + 1, gz - 1);
            ny = y1 - y0 + 1;
            nrows = ny * (z1 - z0 + 1);
#pragma unroll
            for (int o = 0; o < C; ++o)
                c[o] = W.l1.b[o] - (W.l1.w[SN2_F0][o] * q.x + W.l1.w[SN2_F0 + 1][o] * q.y + W.l1.w[SN2_F0 + 2][o] * q.z);
            have = true; more = true;
            cnt = 0; head = 0; tail = 0; f0 = 0; emitted = 0; wblk = 0; wmask = 0u;
            if (!REDO) {  // rows of the 3 x 3 x 3 block -> one flat candidate index space
                int rs_ = 0, len_ = 0;
                if (lane < nrows) {
                    const int rb = row_of(lane);
                    rs_ = __ldg(cs + rb + x0);
                    len_ = __ldg(cs + rb + x1 + 1) - rs_;
                }
                int pre_ = len_;
#pragma unroll
                for (int d = 1; d < 16; d <<= 1) {
                    const int t_ = __shfl_up_sync(SN2_FULL, pre_, d);
                    if (lane >= d) pre_ += t_;
                }
                __syncwarp();
                if (lane < 16) {
                    rtab[lane] = pre_;
                    rtab[16 + lane] = rs_ - (pre_ - len_);
                }
                ftotal = __shfl_sync(SN2_FULL, pre_, 15);
                __syncwarp();
            }
            if (REDO) {  // hit set as a bitmap over the plot's point indices
                for (int w = lane; w < words; w += 32) bm[w] = 0u;
                __syncwarp();
                for (int t = 0; t < nrows; ++t) {
                    const int s2 = __ldg(cs + row_of(t) + x0), e2 = __ldg(cs + row_of(t) + x1 + 1);
                    for (int i = s2 + lane; i < e2; i += 32) {
                        const float4 v = __ldg(so + i);
                        if (dist2(v.x, v.y, v.z, q.x, q.y, q.z) < r2) {
                            const int id = __float_as_int(v.w);
                            atomicOr(&bm[id >> 5], 1u << (id & 31));
                        }
                    }
                }
                __syncwarp();
            }
        }
        // ---- produce: up to 32 edges of the current item ----
        while (have && more && tail - head < 32) {
            if (!REDO) {
                if (f0 >= ftotal) { more = false; break; }
                float4 vv[UNR];
#pragma unroll
                for (int t = 0; t < UNR; ++t) {
                    const int f = f0 + t * 32 + lane;
                    vv[t] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (f < ftotal) {
                        int r = 0;
#pragma unroll
                        for (int k = 0; k < 8; ++k) r += rtab[k] <= f;
                        vv[t] = __ldg(so + f + rtab[16 + r]);
                    }
                }
#pragma unroll
                for (int t = 0; t < UNR; ++t) {
                    if (f0 + t * 32 < ftotal) {
                        const float4 v = vv[t];
                        const bool hit = (f0 + t * 32 + lane < ftotal) && dist2(v.x, v.y, v.z, q.x, q.y, q.z) < r2;
                        const unsigned bal = __ballot_sync(SN2_FULL, hit);
                        if (hit) ring[(tail + __popc(bal & lt)) & (TC_RING - 1)] = __float_as_int(v.w);
                        tail += __popc(bal);
                    }
                }
                f0 += 32 * UNR;
            } else {
                // first K set bits of the hit bitmap in ascending point index (same walk as the SIMT redo kernel)
                if (!wmask) {
                    if (emitted >= K || wblk >= words) { more = false; break; }
                    wword = wblk + lane < words ? bm[wblk + lane] : 0u;
                    wmask = __ballot_sync(SN2_FULL, wword != 0u);
                    if (!wmask) { wblk += 32; continue; }
                }
                const int wl = __ffs(wmask) - 1;
                wmask &= wmask - 1;
                const unsigned bits = __shfl_sync(SN2_FULL, wword, wl);
                const int rk = __popc(bits & lt);
                if (((bits >> lane) & 1u) && emitted + rk < K) ring[(tail + rk) & (TC_RING - 1)] = ((wblk + wl) << 5) + lane;
                const int n = min(__popc(bits), K - emitted);
                tail += n;
                emitted += n;
                if (!wmask) wblk += 32;
                if (emitted >= K) more = false;
            }
            __syncwarp();
        }
        const int avail = tail - head, take = have ? min(avail, 32) : 0;
        const int id = take > 0 ? ring[(head + lane) & (TC_RING - 1)] : 0;
        head += take;
        cnt += take;
        __syncwarp();
        const bool last = have && !more && tail == head;              // this batch ends the current item
        const bool more_after = (have && !last) || (REDO ? item < n_items : *reinterpret_cast<volatile int *>(&s_next) < q_end);
        // ---- layer 1 of my edge -> my row of the A tile, straight into tensor memory ----
        float h1[C];
#pragma unroll
        for (int o = 0; o < C; ++o) h1[o] = 0.f;
        if (lane < take) {
            const float *ur = ub + (size_t)id * C;
#pragma unroll
            for (int g = 0; g < C / 4; ++g) {
                const float4 t4 = ldg4(ur + 4 * g);
                h1[4 * g] = t4.x + c[4 * g];
                h1[4 * g + 1] = t4.y + c[4 * g + 1];
                h1[4 * g + 2] = t4.z + c[4 * g + 2];
                h1[4 * g + 3] = t4.w + c[4 * g + 3];
            }
            relu_bn(W.l1, h1);
        }
        const unsigned a_mine = tmine + (unsigned)(stage * L::STAGE);
        if constexpr (MODE == 1) {
            float lo[C];
#pragma unroll
            for (int o = 0; o < C; ++o) {
                const float hi = __uint_as_float(__float_as_uint(h1[o]) & 0xffffe000u);
                lo[o] = h1[o] - hi;
                h1[o] = hi;
            }
            tmem_st16(a_mine, h1);
            tmem_st16(a_mine + 16, lo);
        } else {
#pragma unroll
            for (int o = 0; o < C; ++o) h1[o] = tc_operand<MODE>(h1[o]);
            tmem_st16(a_mine, h1);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;\n");
        // ---- rendezvous of the group: all 128 rows of this stage are in TMEM; does anybody have edges left? ----
        unsigned any_more;
        asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.u32 q, %1, 0;\n\tbar.red.or.pred p, %2, 128, q;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(any_more) : "r"((unsigned)more_after), "r"(1 + grp) : "memory");
        if (rq == 0 && lane == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;\n");
            const unsigned a_t = tbase + (unsigned)(stage * L::STAGE), d_t = a_t + (unsigned)L::A_COLS;
            // K = 16 = two K = 8 steps: +8 columns of A, +256 bytes (= +16 in the descriptor's address field) of B
            umma_tf32_ts(d_t, a_t, d_bhi, 0u);
            umma_tf32_ts(d_t, a_t + 8, d_bhi + 16, 1u);
            if constexpr (MODE == 1) {
                umma_tf32_ts(d_t, a_t + 16, d_bhi, 1u);
                umma_tf32_ts(d_t, a_t + 24, d_bhi + 16, 1u);
                umma_tf32_ts(d_t, a_t, d_blo, 1u);
                umma_tf32_ts(d_t, a_t + 8, d_blo + 16, 1u);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&bars[2 * grp + stage]))
                         : "memory");
        }
        __syncwarp();
        // ---- consume the previous round while this one is in the tensor core ----
        if (pending) {
            consume(p_take, stage ^ 1);
            if (p_last) finalize();
        }
        pending = true;
        p_take = take;
        p_last = last;
        if (last) {
            p_cnt = cnt;
            p_code = b * M + j;
            p_row = (long long)b * M + qloc;
            have = false;
        }
        stage ^= 1;
        if (!any_more) break;
    }
    if (pending) {
        consume(p_take, stage ^ 1);
        if (p_last) finalize();
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(s_tmem), "r"(L::COLS));
}

__global__ void zero_int_kernel(int *p) { *p = 0; }
// every centroid goes straight to the exact path (small caps bind for most centroids: streaming first is wasted work)
__global__ void all_overflow_kernel(int *ovf, int total)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) ovf[0] = total;
    if (i < total) ovf[1 + i] = i;
}

template <int LEVEL>
static int launch_sa_fused(const float *grid_hdr, const int *cell_start, const float *sorted4, const float *qsorted4,
                           const float *pos4, const float *feat, float *u_scratch, int *ovf, int B, int N, int M,
                           float r2, int K, const float *w_host, int nw, float *out, int *cnt_out, int tc, cudaStream_t st)
{
    typename SAEdge<LEVEL>::W w;
    if (int rc = load_weights(w, w_host, nw)) return rc;
    typename SAEdge<LEVEL>::W ws = w;  // kernels of level 1: second layer negated where the BatchNorm scale is negative
    typename SAEdge<LEVEL>::W wf = w;  // SIMT kernel of level 1: additionally layer 1's BatchNorm affine folded into layer 2
    if constexpr (LEVEL == 1) {
        for (int o = 0; o < SN2_C1; ++o) {
            if (ws.l2.s[o] < 0.f) {
                ws.l2.b[o] = -ws.l2.b[o];
                for (int k = 0; k < SN2_C1; ++k) ws.l2.w[k][o] = -ws.l2.w[k][o];
            }
        }
        wf = ws;
        for (int o = 0; o < SN2_C1; ++o) {
            double bacc = ws.l2.b[o];
            for (int k = 0; k < SN2_C1; ++k) {
                bacc += (double)ws.l1.t[k] * (double)ws.l2.w[k][o];
                wf.l2.w[k][o] = ws.l1.s[k] * ws.l2.w[k][o];
            }
            wf.l2.b[o] = (float)bacc;
        }
    }
    const long long P = (long long)B * N;
    zero_int_kernel<<<1, 1, 0, st>>>(ovf);
    sa_pre_kernel<LEVEL><<<(unsigned)((P + 127) / 128), 128, 0, st>>>(reinterpret_cast<const float4 *>(pos4), feat, P, w,
                                                                     u_scratch);
    SN2_LAUNCH_CHECK("sa_pre_kernel");
    const int words = (N + 31) / 32;
    const size_t smem_ring = (size_t)SF_WARPS * 256 * sizeof(int);
    const size_t smem_stream = smem_ring + (size_t)SF_WARPS * 32 * sizeof(int);  // + the per-warp row tables
    dim3 grid((M + SF_WARPS - 1) / SF_WARPS, B);
    const bool exact_only = K < 256 && K < N;
    if (exact_only) all_overflow_kernel<<<(B * M + 255) / 256, 256, 0, st>>>(ovf, B * M);
    const size_t smem_bm = (size_t)SF_WARPS * words * sizeof(unsigned);
    if constexpr (LEVEL == 1) {
        if (tc) {  // second layer on the tensor cores (tcgen05): 1 = 3xTF32, 2 = TF32, 3 = BF16-precision operands
            if (K < N && smem_bm + 12 * 1024 > 200 * 1024) return SN2_EUNSUPPORTED;
            auto run = [&](auto stream_kern, auto redo_kern) -> int {
                if (!exact_only) {
                    const long long ncta = ((long long)B * M + SF_WARPS * 4 - 1) / (SF_WARPS * 4);  // >= 4 centroids per warp
                    const unsigned qgrid = (unsigned)min(ncta, (long long)148 * 8);  // 1-D work queue over all (plot, centroid) pairs
                    stream_kern<<<qgrid, SF_WARPS * 32, 0, st>>>(grid_hdr, cell_start, reinterpret_cast<const float4 *>(sorted4),
                                                                 reinterpret_cast<const float4 *>(qsorted4), u_scratch, N, M, r2, K, 0, ws, out,
                                                                 cnt_out, ovf, B);
                    SN2_LAUNCH_CHECK("sa1_tc_kernel");
                }
                if (K < N) {  // the cap can bind: exact redo of the overflow list on the same tensor-core path
                    SN2_CUDA_TRY(cudaFuncSetAttribute(redo_kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bm), "sa1_tc redo attr");
                    redo_kern<<<148, SF_WARPS * 32, smem_bm, st>>>(grid_hdr, cell_start, reinterpret_cast<const float4 *>(sorted4),
                                                                   reinterpret_cast<const float4 *>(qsorted4), u_scratch, N, M, r2, K, words, ws,
                                                                   out, cnt_out, ovf, B);
                    SN2_LAUNCH_CHECK("sa1_tc_kernel<redo>");
                }
                return SN2_OK;
            };
            // CTAs per SM of the streaming launch: 2 (<= 128 registers, no spills; default) or 3 (80 registers, TF32 / BF16 modes
            // only: their 128 TMEM columns allow it) -- a tuning switch, read once
            static const int minb = [] { const char *e = getenv("SN2_TC_MINB"); return e && atoi(e) == 3 ? 3 : 2; }();
            if (tc == 1) return run(sa1_tc_kernel<1, false, 2>, sa1_tc_kernel<1, true>);
            if (tc == 2) return minb == 3 ? run(sa1_tc_kernel<2, false, 3>, sa1_tc_kernel<2, true>) : run(sa1_tc_kernel<2, false, 2>, sa1_tc_kernel<2, true>);
            return minb == 3 ? run(sa1_tc_kernel<3, false, 3>, sa1_tc_kernel<3, true>) : run(sa1_tc_kernel<3, false, 2>, sa1_tc_kernel<3, true>);
        }
    }
    if (!exact_only) {
        auto kern = sa_fused_kernel<LEVEL, false>;
        kern<<<grid, SF_WARPS * 32, smem_stream, st>>>(grid_hdr, cell_start, reinterpret_cast<const float4 *>(sorted4),
                                                       reinterpret_cast<const float4 *>(qsorted4), u_scratch, N, M, r2, K, 0, wf,
                                                       out, cnt_out, ovf);
        SN2_LAUNCH_CHECK("sa_fused_kernel");
    }
    if (K < N) {  // the cap can bind: exact redo of the overflow list (exits at once when the list is empty)
        const size_t smem = smem_ring + smem_bm;
        if (smem > 200 * 1024) return SN2_EUNSUPPORTED;
        auto kern = sa_fused_kernel<LEVEL, true>;
        SN2_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "sa_fused attr");
        kern<<<148 * 2, SF_WARPS * 32, smem, st>>>(grid_hdr, cell_start, reinterpret_cast<const float4 *>(sorted4),
                                                   reinterpret_cast<const float4 *>(qsorted4), u_scratch, N, M, r2, K, words,
                                                   wf, out, cnt_out, ovf);
        SN2_LAUNCH_CHECK("sa_fused_kernel<redo>");
    }
    return SN2_OK;
}

}  // namespace sn2

// Debug/profiling entry (not part of the product ABI): level-1 streaming kernel with the message MLP disabled.
extern "C" int sn2_debug_sa_search_only(const float *grid_hdr, const int *cell_start, const float *sorted4,
                                        const float *qsorted4, const float *u, int *ovf, int B, int N, int M, float r2,
                                        int K, const float *w_host, int nw, float *out, int *cnt_out, void *stream)
{
    using namespace sn2;
    W_SA1 w;
    if (int rc = load_weights(w, w_host, nw)) return rc;
    dim3 grid((M + SF_WARPS - 1) / SF_WARPS, B);
    sa_fused_kernel<1, false, 1><<<grid, SF_WARPS * 32, SF_WARPS * (256 + 32) * sizeof(int), (cudaStream_t)stream>>>(
        grid_hdr, cell_start, reinterpret_cast<const float4 *>(sorted4), reinterpret_cast<const float4 *>(qsorted4), u, N, M,
        r2, K, 0, w, out, cnt_out, ovf);
    SN2_LAUNCH_CHECK("sa_fused_kernel<dbg>");
    return SN2_OK;
}

extern "C" int sn2_sa_fused_fwd(int level, const float *grid_hdr, const int *cell_start, const float *sorted4,
                                const float *qsorted4, const float *pos4, const float *feat, float *u_scratch,
                                int *ovf_scratch, int B, int N, int M, float r2, int K, const float *w_host, int nw,
                                float *out, int *cnt_out, int tensor_core, void *stream)
{
    if (!grid_hdr || !cell_start || !sorted4 || !qsorted4 || !pos4 || !feat || !u_scratch || !ovf_scratch || !out ||
        B <= 0 || N <= 0 || M <= 0 || K <= 0)
        return SN2_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    if (level == 1)
        return sn2::launch_sa_fused<1>(grid_hdr, cell_start, sorted4, qsorted4, pos4, feat, u_scratch, ovf_scratch, B, N, M,
                                       r2, K, w_host, nw, out, cnt_out, (tensor_core >= 1 && tensor_core <= 3) ? tensor_core : 0, st);
    if (level == 2)
        return sn2::launch_sa_fused<2>(grid_hdr, cell_start, sorted4, qsorted4, pos4, feat, u_scratch, ovf_scratch, B, N, M,
                                       r2, K, w_host, nw, out, cnt_out, 0, st);
    return SN2_EINVAL;
}
