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
constexpr int TC_PAD = 6144;                                  // an M=128 operand window around a 32-row region
constexpr int TC_A_BYTES = TC_PAD + SF_WARPS * 8192 + 8192;   // per warp: 2 stages x (2 KB hi rows + 2 KB lo rows)
constexpr int TC_SMEM = TC_A_BYTES + 2 * 1024 + 256;          // + W2 hi / lo tiles + mbarriers
constexpr int TC_TMEM_COLS = SF_WARPS * 32;                   // 2 stages x 16 accumulator columns per warp

template <int LEVEL, bool REDO, int DBG = 0, int TC = 0>  // TC: 0 SIMT, 1 tcgen05 3xTF32 (fp32-accurate), 2 tcgen05 plain TF32
__global__ void __launch_bounds__(SF_WARPS * 32, REDO ? 1 : (LEVEL == 1 ? (TC ? 2 : 3) : 2))
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
    const unsigned lt = (1u << lane) - 1u;

    // ---- tensor-core state (TC): per-warp operand rows, W2 tiles, per-warp mbarrier, TMEM accumulators ----
    static_assert(!TC || (LEVEL == 1 && !REDO), "the tcgen05 path covers the level-1 streaming kernel");
    unsigned char *tc_base = sf_smem + SF_WARPS * RING * sizeof(int);  // 1024-byte aligned (ring = 8 KB)
    unsigned char *a_st = tc_base + TC_PAD + warp * 8192;  // stage s: hi rows at a_st + s*4096, lo rows 2 KB after
    float *b_hi = reinterpret_cast<float *>(tc_base + TC_A_BYTES), *b_lo = b_hi + 256;
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(tc_base + TC_A_BYTES + 2048);
    __shared__ unsigned s_tmem;
    unsigned tmem_d = 0, tc_parity = 0;  // tc_parity: bit s = phase parity of stage s's mbarrier
    unsigned long long d_a0 = 0, d_bhi = 0, d_blo = 0;
    if constexpr (TC) {
        if (warp == 0) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&s_tmem)),
                         "r"(TC_TMEM_COLS));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
        }
        if (lane < 2) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bars[2 * warp + lane])));
        {   // W2 as the B operand [N = 16 (o) x K = 16 (k)], split into tf32 hi + lo (3xTF32: fp32-accurate product)
            const int o = threadIdx.x >> 4, k = threadIdx.x & 15;
            const float wv = W.l2.w[k][o];
            const float hi = __uint_as_float(__float_as_uint(wv) & 0xffffe000u);
            const int off = (o >> 3) * 128 + (k >> 2) * 32 + (o & 7) * 4 + (k & 3);  // in floats
            b_hi[off] = TC == 1 ? hi : wv;
            b_lo[off] = wv - hi;
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n");
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;\n");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;\n");
        const int qd = warp & 3;  // TMEM lane quarter this warp can read = rows 32*qd.. of its M=128 tile
        tmem_d = s_tmem + ((unsigned)(32 * qd) << 16) + (unsigned)(32 * warp);
        d_a0 = umma_desc(smem_u32(a_st) - qd * 2048);  // stage s: +4096 B (= +256 in the address field); lo: +2048 B
        d_bhi = umma_desc(smem_u32(b_hi));
        d_blo = umma_desc(smem_u32(b_lo));
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
            if (j >= M) break;  // (no early return: the TC variant frees tensor memory after the loop)
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
                relu_bn(W.l1, h1);
                float h2[SN2_C1];
                acc_init(W.l2, h2);
#pragma unroll
                for (int k = 0; k < SN2_C1; ++k) acc_step<0>(W.l2, h1[k], h2, k);
                // No ReLU / BatchNorm per edge: f(a) = fma(max(a, 0), s, t) is monotone in the pre-activation a, so
                // max_j f(a_j) = f(max_j a_j) for s >= 0 and f(min_j a_j) for s < 0.  The launcher negates the
                // second-layer weights and bias of the s < 0 channels (fma is odd: the accumulators come out exactly
                // negated), so a running max serves both signs; f is applied once per centroid at the end.
#pragma unroll
                for (int o = 0; o < C; ++o) mx[o] = fmaxf(mx[o], h2[o]);
            }
        };

        // Tensor-core form of one 32-edge batch, split in two halves so the MMA latency hides under the next
        // batch's search + gather (two operand / accumulator stages per warp):
        //   tc_issue  : layer 1 in registers, rows -> this warp's private operand region, 6 x tcgen05.mma
        //               (128x16x8 TF32; 3xTF32 split A_hi B_hi + A_lo B_hi + A_hi B_lo keeps fp32 accuracy;
        //               the rows of the other three lane quarters of the M=128 tile are don't-care), commit.
        //   tc_consume: wait for that stage's mbarrier, tcgen05.ld the 32 rows x 16 columns, ReLU/BN/max.
        auto tc_issue = [&](const int id, const int take, const int st) {
            if constexpr (TC) {
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
                unsigned char *ah = a_st + st * 4096 + (lane >> 3) * 512 + (lane & 7) * 16;
#pragma unroll
                for (int kc = 0; kc < 4; ++kc) {
                    float hi[4], lo[4];
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        hi[t] = __uint_as_float(__float_as_uint(h1[4 * kc + t]) & 0xffffe000u);
                        lo[t] = h1[4 * kc + t] - hi[t];
                    }
                    if (TC == 1) {
                        *reinterpret_cast<float4 *>(ah + kc * 128) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                        *reinterpret_cast<float4 *>(ah + 2048 + kc * 128) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                    } else {  // plain TF32: the tensor core truncates the low 13 mantissa bits itself
                        *reinterpret_cast<float4 *>(ah + kc * 128) =
                            make_float4(h1[4 * kc], h1[4 * kc + 1], h1[4 * kc + 2], h1[4 * kc + 3]);
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    asm volatile("tcgen05.fence::after_thread_sync;\n");
                    const unsigned dcol = s_tmem + (unsigned)(32 * warp + 16 * st);  // all 128 lanes, this stage's 16 columns
                    const unsigned long long ahi = d_a0 + (unsigned long long)(st * 256), alo = ahi + 128;
                    // K = 16 = two K=8 steps: +256 bytes (2 core matrices, +16 in the address field) on both operands
                    umma_tf32(dcol, ahi, d_bhi, 0u);
                    umma_tf32(dcol, ahi + 16, d_bhi + 16, 1u);
                    if (TC == 1) {
                        umma_tf32(dcol, alo, d_bhi, 1u);
                        umma_tf32(dcol, alo + 16, d_bhi + 16, 1u);
                        umma_tf32(dcol, ahi, d_blo, 1u);
                        umma_tf32(dcol, ahi + 16, d_blo + 16, 1u);
                    }
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(
                                     smem_u32(&bars[2 * warp + st]))
                                 : "memory");
                }
                __syncwarp();
            }
        };
        auto tc_consume = [&](const int take, const int st) {
            if constexpr (TC) {
                unsigned done = 0;
                for (int spin = 0; spin < (1 << 24) && !done; ++spin)
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                                 "selp.u32 %0, 1, 0, p;\n\t}\n"
                                 : "=r"(done)
                                 : "r"(smem_u32(&bars[2 * warp + st])), "r"((tc_parity >> st) & 1u)
                                 : "memory");
                if (!done) __trap();  // the MMA never completed: fail loudly instead of hanging the GPU
                tc_parity ^= 1u << st;
                asm volatile("tcgen05.fence::after_thread_sync;\n");
                unsigned d[16];
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                             : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7]),
                               "=r"(d[8]), "=r"(d[9]), "=r"(d[10]), "=r"(d[11]), "=r"(d[12]), "=r"(d[13]), "=r"(d[14]), "=r"(d[15])
                             : "r"(tmem_d + (unsigned)(16 * st)));
                asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;\n");
                if (lane < take) {
#pragma unroll
                    for (int o = 0; o < C; ++o) {
                        const float h2 = fmaf(fmaxf(__uint_as_float(d[o]) + W.l2.b[o], 0.f), W.l2.s[o], W.l2.t[o]);
                        mx[o] = fmaxf(mx[o], h2);
                    }
                }
                __syncwarp();
            }
        };
        int tc_stage = 0, tc_pending = 0;  // tc_pending = rows of the batch in flight on stage tc_stage ^ 1 (0: none)

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
        int y = 0, base = 0, e = 0;      // streaming iterator: current (layer, cell row) and candidate range
        bool open_row = false;
        int emitted = 0, wblk = 0;       // redo iterator: set bits emitted so far, next 32-word block of the bitmap
        unsigned wmask = 0u, wword = 0u; //   non-empty words left in the current block, this lane's word of the block
        bool more = true;
        while (true) {
            while (more && tail - head < 32) {
                if (!REDO) {
                    if (!open_row) {
                        if (y >= nrows) { more = false; break; }
                        base = __ldg(cs + row_of(y) + x0);
                        e = __ldg(cs + row_of(y) + x1 + 1);
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
                        if (base + t * 32 < e) {  // warp-uniform: short runs (3-D cells) cost one sub-batch, not UNR
                            const float4 v = vv[t];
                            const bool hit = (base + t * 32 + lane < e) && dist2(v.x, v.y, v.z, q.x, q.y, q.z) < r2;
                            const unsigned bal = __ballot_sync(SN2_FULL, hit);
                            if (hit) ring[(tail + __popc(bal & lt)) & (RING - 1)] = __float_as_int(v.w);
                            tail += __popc(bal);
                        }
                    }
                    base += 32 * UNR;
                    if (base >= e) { open_row = false; ++y; }
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
            if (TC) {
                tc_issue(id, take, tc_stage);
                if (tc_pending) tc_consume(tc_pending, tc_stage ^ 1);
                tc_pending = take;
                tc_stage ^= 1;
            } else if (LEVEL == 2) {
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

        if (TC && tc_pending) tc_consume(tc_pending, tc_stage ^ 1);
        const size_t row = (size_t)b * M + qloc;
        if (!REDO && cnt > K) {  // the cap binds: leave this centroid to the exact redo launch
            if (lane == 0) ovf[1 + atomicAdd(ovf, 1)] = b * M + j;
            continue;
        }
        if constexpr (LEVEL == 1) {
#pragma unroll
            for (int o = 0; o < C; ++o) mx[o] = cnt > 0 ? warp_max(mx[o]) : 0.f;
            if constexpr (!TC) {  // streaming / redo kernels track the (sign-adjusted) pre-activation: finish it here
#pragma unroll
                for (int o = 0; o < C; ++o) {
                    const float sc = W.l2.s[o], a = sc < 0.f ? -mx[o] : mx[o];
                    mx[o] = cnt > 0 ? fmaf(fmaxf(a, 0.f), sc, W.l2.t[o]) : 0.f;
                }
            }
            if (lane == 0) {
                float4 *o4 = reinterpret_cast<float4 *>(out + row * C);
#pragma unroll
                for (int v = 0; v < C / 4; ++v) o4[v] = make_float4(mx[4 * v], mx[4 * v + 1], mx[4 * v + 2], mx[4 * v + 3]);
            }
        } else {
            // out = max_j f(u_j) = f(max u) if s >= 0 else f(min u): f is monotone in u (lane = channel)
            const float sc = W.l1.s[lane], sel = sc >= 0.f ? mx[0] : mn[0];
            out[row * C + lane] = cnt > 0 ? fmaf(fmaxf(sel + c[0], 0.f), sc, W.l1.t[lane]) : 0.f;
        }
        if (lane == 0) {
            if (cnt_out) cnt_out[row] = cnt;
        }
    }
    if constexpr (TC) {  // release tensor memory once every warp is done with it
        asm volatile("tcgen05.fence::before_thread_sync;\n");
        __syncthreads();
        if (warp == 0)
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(s_tmem), "r"(TC_TMEM_COLS));
    }
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
    typename SAEdge<LEVEL>::W ws = w;  // SIMT kernels of level 1: second layer negated where the BatchNorm scale is negative
    if constexpr (LEVEL == 1) {
        for (int o = 0; o < SN2_C1; ++o) {
            if (ws.l2.s[o] < 0.f) {
                ws.l2.b[o] = -ws.l2.b[o];
                for (int k = 0; k < SN2_C1; ++k) ws.l2.w[k][o] = -ws.l2.w[k][o];
            }
        }
    }
    const long long P = (long long)B * N;
    zero_int_kernel<<<1, 1, 0, st>>>(ovf);
    sa_pre_kernel<LEVEL><<<(unsigned)((P + 127) / 128), 128, 0, st>>>(reinterpret_cast<const float4 *>(pos4), feat, P, w,
                                                                     u_scratch);
    SN2_LAUNCH_CHECK("sa_pre_kernel");
    const int words = (N + 31) / 32;
    const size_t smem_ring = (size_t)SF_WARPS * 256 * sizeof(int);
    dim3 grid((M + SF_WARPS - 1) / SF_WARPS, B);
    const bool exact_only = K < 256 && K < N;
    if (exact_only) {
        all_overflow_kernel<<<(B * M + 255) / 256, 256, 0, st>>>(ovf, B * M);
        tc = 0;
    }
    if constexpr (LEVEL == 1) {
        if (tc && !exact_only) {  // second layer on the tensor cores (tcgen05): 1 = 3xTF32, 2 = plain TF32
            auto kern = tc == 1 ? sa_fused_kernel<1, false, 0, 1> : sa_fused_kernel<1, false, 0, 2>;
            const size_t smem_tc = smem_ring + TC_SMEM;
            SN2_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tc), "sa_fused_tc attr");
            kern<<<grid, SF_WARPS * 32, smem_tc, st>>>(grid_hdr, cell_start, reinterpret_cast<const float4 *>(sorted4),
                                                       reinterpret_cast<const float4 *>(qsorted4), u_scratch, N, M, r2, K, 0, w,
                                                       out, cnt_out, ovf);
            SN2_LAUNCH_CHECK("sa_fused_kernel<tc>");
        }
    }
    if (!(LEVEL == 1 && tc) && !exact_only) {
        auto kern = sa_fused_kernel<LEVEL, false>;
        kern<<<grid, SF_WARPS * 32, smem_ring, st>>>(grid_hdr, cell_start, reinterpret_cast<const float4 *>(sorted4),
                                                     reinterpret_cast<const float4 *>(qsorted4), u_scratch, N, M, r2, K, 0, ws,
                                                     out, cnt_out, ovf);
        SN2_LAUNCH_CHECK("sa_fused_kernel");
    }
    if (K < N) {  // the cap can bind: exact redo of the overflow list (exits at once when the list is empty)
        const size_t smem = smem_ring + (size_t)SF_WARPS * words * sizeof(unsigned);
        if (smem > 200 * 1024) return SN2_EUNSUPPORTED;
        auto kern = sa_fused_kernel<LEVEL, true>;
        SN2_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "sa_fused attr");
        kern<<<148 * 2, SF_WARPS * 32, smem, st>>>(grid_hdr, cell_start, reinterpret_cast<const float4 *>(sorted4),
                                                   reinterpret_cast<const float4 *>(qsorted4), u_scratch, N, M, r2, K, words,
                                                   ws, out, cnt_out, ovf);
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
    sa_fused_kernel<1, false, 1><<<grid, SF_WARPS * 32, SF_WARPS * 256 * sizeof(int), (cudaStream_t)stream>>>(
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
                                       r2, K, w_host, nw, out, cnt_out, tensor_core == 2 ? 2 : (tensor_core ? 1 : 0), st);
    if (level == 2)
        return sn2::launch_sa_fused<2>(grid_hdr, cell_start, sorted4, qsorted4, pos4, feat, u_scratch, ovf_scratch, B, N, M,
                                       r2, K, w_host, nw, out, cnt_out, 0, st);
    return SN2_EINVAL;
}
