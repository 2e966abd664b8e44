// K2 radius ball query: grid-binned search, canonical
// index-ordered neighbour lists capped at K, CSR output (no [M, cap] scratch).
//
//   grid_build : one CTA per plot: bounding box -> cell histogram -> scan -> scatter of
//                (x, y, z, local index) into cell order.  Cell edge = max(r*(1+1e-4), extent/64), so
//                any two points closer than r lie in the same or adjacent cells (3x3x3 search) even
//                after fp32 rounding of the cell coordinate.  The grid is 3-D (z layers of the same
//                edge, as many as fit in 4096 cells): FPS centroids are mostly sparse canopy points
//                whose xy column is full of ground points -- the z layers cut their candidates ~20x.
//   ball_count : one warp per query: scans the 3 cell-row ranges, counts d2 < r2 -> min(count, K).
//   rowptr_scan: per-plot block scan + plot bases -> CSR row pointer.
//   ball_fill  : one warp per query: same search, hits set bits in a per-warp bitmap over the plot's
//                point indices (shared memory); enumerating the bitmap yields ascending index order
//                for free, and "first K in ascending index" is a prefix of that enumeration.
#include "sn2_common.cuh"

namespace sn2 {

constexpr int GB_THREADS = 1024;

__device__ __forceinline__ int cell_coord(float v, float mn, float inv, int g)
{
    int c = (int)floorf(__fmul_rn(__fsub_rn(v, mn), inv));
    return min(max(c, 0), g - 1);
}

__global__ void __launch_bounds__(GB_THREADS, 1)
grid_build_kernel(const float4 *__restrict__ pos, int N, float r, float *__restrict__ grid_hdr,
                  int *__restrict__ cell_start, float4 *__restrict__ sorted)
{
    __shared__ float red[6][32];
    __shared__ int hist[SN2_GRID_CELLS];
    __shared__ int wsum[32];
    __shared__ float s_hdr[6];
    __shared__ int s_g[3];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float4 *p = pos + (size_t)b * N;

    // 1. bounding box
    float mnx = INFINITY, mny = INFINITY, mxx = -INFINITY, mxy = -INFINITY, mnz = INFINITY, mxz = -INFINITY;
    for (int i = tid; i < N; i += GB_THREADS) {
        float4 v = __ldg(p + i);
        mnx = fminf(mnx, v.x);
        mxx = fmaxf(mxx, v.x);
        mny = fminf(mny, v.y);
        mxy = fmaxf(mxy, v.y);
        mnz = fminf(mnz, v.z);
        mxz = fmaxf(mxz, v.z);
    }
    mnx = -warp_max(-mnx);
    mny = -warp_max(-mny);
    mnz = -warp_max(-mnz);
    mxx = warp_max(mxx);
    mxy = warp_max(mxy);
    mxz = warp_max(mxz);
    if (lane == 0) {
        red[0][warp] = mnx;
        red[1][warp] = mny;
        red[2][warp] = mxx;
        red[3][warp] = mxy;
        red[4][warp] = mnz;
        red[5][warp] = mxz;
    }
    for (int c = tid; c < SN2_GRID_CELLS; c += GB_THREADS) hist[c] = 0;
    __syncthreads();
    if (warp == 0) {
        mnx = -warp_max(-red[0][lane]);
        mny = -warp_max(-red[1][lane]);
        mxx = warp_max(red[2][lane]);
        mxy = warp_max(red[3][lane]);
        mnz = -warp_max(-red[4][lane]);
        mxz = warp_max(red[5][lane]);
        if (lane == 0) {
            float ext = fmaxf(mxx - mnx, mxy - mny);
            // r > 0: cell edge just above r (3x3 search is exhaustive for radius r);
            // r < 0: auto edge for ~(-r) points per cell (k-NN ring search, csrc/knn.cu)
            float want = r > 0.f ? r * 1.0001f : sqrtf(fmaxf((mxx - mnx) * (mxy - mny) * (-r) / (float)N, 0.f));
            float cs = fmaxf(want, ext / (float)(SN2_GRID_MAX - 1));
            cs = fmaxf(cs, 1e-20f);
            float inv = 1.0f / cs;
            int gx = min(SN2_GRID_MAX, (int)floorf((mxx - mnx) * inv) + 1);
            int gy = min(SN2_GRID_MAX, (int)floorf((mxy - mny) * inv) + 1);
            // z layers (ball-query grids only): same edge as xy when they fit in the cell budget, else thicker
            int gz = 1;
            float invz = 0.f, csz = INFINITY;
            const float zext = mxz - mnz;
            const int gzmax = SN2_GRID_CELLS / (gx * gy);
            if (r > 0.f && gzmax > 1 && zext > 0.f) {
                const int need = (int)floorf(zext / cs) + 1;
                if (need <= gzmax) { gz = need; csz = cs; }
                else { gz = gzmax; csz = zext / ((float)gz - 0.5f); }  // floor(zext / csz) = gz - 1, csz > cs
                invz = 1.0f / csz;
                gz = min(gz, (int)floorf(zext * invz) + 1);
            }
            s_hdr[0] = mnx;
            s_hdr[1] = mny;
            s_hdr[2] = inv;
            s_hdr[3] = mnz;
            s_hdr[4] = invz;
            s_g[0] = gx;
            s_g[1] = gy;
            s_g[2] = gz;
            float *h = grid_hdr + (size_t)b * SN2_GRID_HDR;
            h[0] = mnx;
            h[1] = mny;
            h[2] = inv;
            h[3] = cs;
            h[4] = __int_as_float(gx);
            h[5] = __int_as_float(gy);
            h[6] = mnz;
            h[7] = invz;
            h[8] = __int_as_float(gz);
            h[9] = csz;
            h[10] = 0.f;
            h[11] = 0.f;
        }
    }
    __syncthreads();
    const float ox = s_hdr[0], oy = s_hdr[1], inv = s_hdr[2], oz = s_hdr[3], invz = s_hdr[4];
    const int gx = s_g[0], gy = s_g[1], gz = s_g[2];

    // 2. histogram
    for (int i = tid; i < N; i += GB_THREADS) {
        float4 v = __ldg(p + i);
        int c = (cell_coord(v.z, oz, invz, gz) * gy + cell_coord(v.y, oy, inv, gy)) * gx + cell_coord(v.x, ox, inv, gx);
        atomicAdd(&hist[c], 1);
    }
    __syncthreads();

    // 3. exclusive scan of SN2_GRID_CELLS counters (4 per thread)
    constexpr int PER = SN2_GRID_CELLS / GB_THREADS;
    int loc[PER];
    int s = 0;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        loc[k] = hist[tid * PER + k];
        s += loc[k];
    }
    int incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(SN2_FULL, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = wsum[lane];
        int wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(SN2_FULL, wi, o);
            if (lane >= o) wi += t;
        }
        wsum[lane] = wi - w;  // exclusive warp base
    }
    __syncthreads();
    int base = wsum[warp] + incl - s;
    int *cs_out = cell_start + (size_t)b * (SN2_GRID_CELLS + 1);
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        hist[tid * PER + k] = base;  // becomes the scatter cursor
        cs_out[tid * PER + k] = base;
        base += loc[k];
    }
    if (tid == GB_THREADS - 1) cs_out[SN2_GRID_CELLS] = base;
    __syncthreads();

    // 4. scatter (order inside a cell is arbitrary; ball_fill canonicalises through its bitmap)
    float4 *so = sorted + (size_t)b * N;
    for (int i = tid; i < N; i += GB_THREADS) {
        float4 v = __ldg(p + i);
        int c = (cell_coord(v.z, oz, invz, gz) * gy + cell_coord(v.y, oy, inv, gy)) * gx + cell_coord(v.x, ox, inv, gx);
        int dst = atomicAdd(&hist[c], 1);
        so[dst] = make_float4(v.x, v.y, v.z, __int_as_float(i));
    }
}

// Shared search loop.  f(local_index) is called by the lane that found an in-radius point.
template <typename F>
__device__ __forceinline__ void ball_search(const float *__restrict__ hdr, const int *__restrict__ cs,
                                            const float4 *__restrict__ so, float4 q, float r2, int lane, F f)
{
    const float ox = hdr[0], oy = hdr[1], inv = hdr[2], oz = hdr[6], invz = hdr[7];
    const int gx = __float_as_int(hdr[4]), gy = __float_as_int(hdr[5]), gz = __float_as_int(hdr[8]);
    const int ix = cell_coord(q.x, ox, inv, gx), iy = cell_coord(q.y, oy, inv, gy), iz = cell_coord(q.z, oz, invz, gz);
    const int x0 = max(ix - 1, 0), x1 = min(ix + 1, gx - 1);
    for (int z = max(iz - 1, 0); z <= min(iz + 1, gz - 1); ++z)
        for (int y = max(iy - 1, 0); y <= min(iy + 1, gy - 1); ++y) {
            const int row = (z * gy + y) * gx;
            const int s = __ldg(cs + row + x0), e = __ldg(cs + row + x1 + 1);
            for (int j = s + lane; j < e; j += 32) {
                float4 v = __ldg(so + j);
                if (dist2(v.x, v.y, v.z, q.x, q.y, q.z) < r2) f(__float_as_int(v.w));
            }
        }
}

__global__ void __launch_bounds__(256)
ball_count_kernel(const float *__restrict__ grid_hdr, const int *__restrict__ cell_start,
                  const float4 *__restrict__ sorted, const float4 *__restrict__ qpos, int B, int N, int M,
                  float r2, int K, int *__restrict__ cnt)
{
    const int lane = threadIdx.x & 31;
    const long long q = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (q >= (long long)B * M) return;
    const int b = (int)(q / M);
    int c = 0;
    ball_search(grid_hdr + (size_t)b * SN2_GRID_HDR, cell_start + (size_t)b * (SN2_GRID_CELLS + 1),
                sorted + (size_t)b * N, __ldg(qpos + q), r2, lane, [&](int) { ++c; });
    c = __reduce_add_sync(SN2_FULL, c);
    if (lane == 0) cnt[q] = min(c, K);
}

constexpr int SCAN_THREADS = 1024;
__global__ void __launch_bounds__(SCAN_THREADS)
rowptr_local_kernel(const int *__restrict__ cnt, int M, int *__restrict__ rowptr, int *__restrict__ totals)
{
    __shared__ int wsum[32];
    __shared__ int carry_s;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int *c = cnt + (size_t)b * M;
    int *o = rowptr + (size_t)b * M;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < M; base += SCAN_THREADS) {
        const int carry = carry_s;  // stable here: last written before the previous trailing barrier
        int i = base + tid;
        int v = i < M ? c[i] : 0;
        int incl = v;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
            int t = __shfl_up_sync(SN2_FULL, incl, s);
            if (lane >= s) incl += t;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = wsum[lane], wi = w;
#pragma unroll
            for (int s = 1; s < 32; s <<= 1) {
                int t = __shfl_up_sync(SN2_FULL, wi, s);
                if (lane >= s) wi += t;
            }
            wsum[lane] = wi - w;
            if (lane == 31) carry_s = carry + wi;
        }
        __syncthreads();
        if (i < M) o[i] = carry + wsum[warp] + incl - v;
        __syncthreads();
    }
    if (tid == 0) totals[b] = carry_s;
}

__global__ void __launch_bounds__(256)
rowptr_final_kernel(int B, int M, int *__restrict__ rowptr, const int *__restrict__ totals)
{
    __shared__ int red[8];
    __shared__ int base_s;
    const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int s = 0;
    for (int i = tid; i < b; i += 256) s += totals[i];
    s = __reduce_add_sync(SN2_FULL, s);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    if (tid == 0) {
        int t = 0;
        for (int w = 0; w < 8; ++w) t += red[w];
        base_s = t;
        if (b == B - 1 && blockIdx.x == 0) rowptr[(size_t)B * M] = t + totals[b];
    }
    __syncthreads();
    const int base = base_s;
    int i = blockIdx.x * 256 + tid;
    if (i < M && base) rowptr[(size_t)b * M + i] += base;
}

constexpr int FILL_WARPS = 8;
__global__ void __launch_bounds__(FILL_WARPS * 32)
ball_fill_kernel(const float *__restrict__ grid_hdr, const int *__restrict__ cell_start,
                 const float4 *__restrict__ sorted, const float4 *__restrict__ qpos, int B, int N, int M,
                 float r2, int K, const int *__restrict__ rowptr, int *__restrict__ col, int words_per_lane)
{
    extern __shared__ unsigned bm_all[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int W = words_per_lane * 32;
    unsigned *bm = bm_all + (size_t)warp * W;
    const long long nq = (long long)B * M;
    for (long long q = (long long)blockIdx.x * FILL_WARPS + warp; q < nq; q += (long long)gridDim.x * FILL_WARPS) {
        const int b = (int)(q / M);
        for (int w = lane; w < W; w += 32) bm[w] = 0u;
        __syncwarp();
        ball_search(grid_hdr + (size_t)b * SN2_GRID_HDR, cell_start + (size_t)b * (SN2_GRID_CELLS + 1),
                    sorted + (size_t)b * N, __ldg(qpos + q), r2, lane,
                    [&](int i) { atomicOr(&bm[i >> 5], 1u << (i & 31)); });
        __syncwarp();
        // lane owns words [lane*wpl, (lane+1)*wpl): ascending lanes = ascending indices
        int c = 0;
        for (int k = 0; k < words_per_lane; ++k) c += __popc(bm[lane * words_per_lane + k]);
        int incl = c;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
            int t = __shfl_up_sync(SN2_FULL, incl, s);
            if (lane >= s) incl += t;
        }
        int off = incl - c;
        const int row = __ldg(rowptr + q);
        const int gbase = b * N;
        for (int k = 0; k < words_per_lane && off < K; ++k) {
            unsigned w = bm[lane * words_per_lane + k];
            const int ibase = (lane * words_per_lane + k) << 5;
            while (w && off < K) {
                int bit = __ffs(w) - 1;
                w &= w - 1;
                col[row + off] = gbase + ibase + bit;
                ++off;
            }
        }
        __syncwarp();
    }
}

}  // namespace sn2

extern "C" int sn2_grid_build(const float *pos4, int B, int N, float r, float *grid_hdr, int *cell_start,
                              float *sorted4, void *stream)
{
    if (!pos4 || !grid_hdr || !cell_start || !sorted4 || B <= 0 || N <= 0 || !(r != 0.f)) return SN2_EINVAL;
    sn2::grid_build_kernel<<<B, sn2::GB_THREADS, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4 *>(pos4), N, r, grid_hdr, cell_start, reinterpret_cast<float4 *>(sorted4));
    SN2_LAUNCH_CHECK("grid_build_kernel");
    return SN2_OK;
}

extern "C" int sn2_ball_count(const float *grid_hdr, const int *cell_start, const float *sorted4,
                              const float *qpos4, int B, int N, int M, float r2, int K, int *cnt, void *stream)
{
    if (!grid_hdr || !cell_start || !sorted4 || !qpos4 || !cnt || B <= 0 || N <= 0 || M <= 0 || K <= 0)
        return SN2_EINVAL;
    long long warps = (long long)B * M;
    long long blocks = (warps * 32 + 255) / 256;
    sn2::ball_count_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        grid_hdr, cell_start, reinterpret_cast<const float4 *>(sorted4), reinterpret_cast<const float4 *>(qpos4), B, N,
        M, r2, K, cnt);
    SN2_LAUNCH_CHECK("ball_count_kernel");
    return SN2_OK;
}

extern "C" int sn2_rowptr_scan(const int *cnt, int B, int M, int *rowptr, int *scratch, void *stream)
{
    if (!cnt || !rowptr || !scratch || B <= 0 || M <= 0) return SN2_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    sn2::rowptr_local_kernel<<<B, sn2::SCAN_THREADS, 0, st>>>(cnt, M, rowptr, scratch);
    SN2_LAUNCH_CHECK("rowptr_local_kernel");
    dim3 grid((M + 255) / 256, B);
    sn2::rowptr_final_kernel<<<grid, 256, 0, st>>>(B, M, rowptr, scratch);
    SN2_LAUNCH_CHECK("rowptr_final_kernel");
    return SN2_OK;
}

extern "C" int sn2_ball_fill(const float *grid_hdr, const int *cell_start, const float *sorted4,
                             const float *qpos4, int B, int N, int M, float r2, int K, const int *rowptr, int *col,
                             void *stream)
{
    if (!grid_hdr || !cell_start || !sorted4 || !qpos4 || !rowptr || !col || B <= 0 || N <= 0 || M <= 0 || K <= 0)
        return SN2_EINVAL;
    int words = (N + 31) / 32;
    int wpl = (words + 31) / 32;
    size_t smem = (size_t)sn2::FILL_WARPS * wpl * 32 * sizeof(unsigned);
    if (smem > 200 * 1024) return SN2_EUNSUPPORTED;
    auto kern = sn2::ball_fill_kernel;
    SN2_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "fill attr");
    long long nq = (long long)B * M;
    long long blocks = (nq + sn2::FILL_WARPS - 1) / sn2::FILL_WARPS;
    if (blocks > 148 * 16) blocks = 148 * 16;
    kern<<<(unsigned)blocks, sn2::FILL_WARPS * 32, smem, (cudaStream_t)stream>>>(
        grid_hdr, cell_start, reinterpret_cast<const float4 *>(sorted4), reinterpret_cast<const float4 *>(qpos4), B, N,
        M, r2, K, rowptr, col, wpl);
    SN2_LAUNCH_CHECK("ball_fill_kernel");
    return SN2_OK;
}
