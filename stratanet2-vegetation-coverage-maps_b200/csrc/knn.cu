// K4 search part: exact 3-nearest-neighbour search on the per-plot xy grid (grid_build with r < 0: one z layer) (SURVEY.md §8a a7/a8,
// Appendix A5).  One thread per query; the sources of a plot are binned by grid_build (auto cell edge,
// ~3 sources per cell).  The search visits the 3x3 block around the query's cell, then square rings,
// and stops as soon as the 3rd best squared distance is strictly below the squared xy-distance to the
// unvisited region (minus a safety margin covering the fp32 rounding of the cell assignment), so the
// result equals the brute-force scan.  Candidates arrive out of index order: ties are broken
// explicitly, lower source index first, as the oracle's ascending scan does.
#include "sn2_common.cuh"

namespace sn2 {

struct Top3 {
    float d0, d1, d2;
    int i0, i1, i2;
    __device__ __forceinline__ void init()
    {
        d0 = d1 = d2 = INFINITY;
        i0 = i1 = i2 = 0x7fffffff;
    }
    // strict lexicographic (d, index) order
    __device__ __forceinline__ void push(float d, int i)
    {
        if (d < d2 || (d == d2 && i < i2)) {
            if (d < d1 || (d == d1 && i < i1)) {
                d2 = d1; i2 = i1;
                if (d < d0 || (d == d0 && i < i0)) { d1 = d0; i1 = i0; d0 = d; i0 = i; }
                else { d1 = d; i1 = i; }
            } else { d2 = d; i2 = i; }
        }
    }
};

__device__ __forceinline__ void scan_range(const float4 *__restrict__ so, int s, int e, const float4 q, Top3 &t)
{
    for (int j = s; j < e; ++j) {
        const float4 v = __ldg(so + j);
        // diff = source - query (PyG knn_interpolate forms pos_x[x_idx] - pos_y[y_idx])
        t.push(dist2(v.x, v.y, v.z, q.x, q.y, q.z), __float_as_int(v.w));
    }
}

constexpr int KG_THREADS = 128;
__global__ void __launch_bounds__(KG_THREADS)
knn3_grid_kernel(const float *__restrict__ grid_hdr, const int *__restrict__ cell_start,
                 const float4 *__restrict__ sorted, const float4 *__restrict__ qpos, int Ms, int Nq,
                 int *__restrict__ nbr, float *__restrict__ wgt, int q_sorted)
{
    const int b = blockIdx.y;
    int qi = blockIdx.x * KG_THREADS + threadIdx.x;
    if (qi >= Nq) return;
    const float *hdr = grid_hdr + (size_t)b * SN2_GRID_HDR;
    const int *cs = cell_start + (size_t)b * (SN2_GRID_CELLS + 1);
    const float4 *so = sorted + (size_t)b * Ms;
    const float ox = hdr[0], oy = hdr[1], inv = hdr[2], cell = hdr[3];
    const int gx = __float_as_int(hdr[4]), gy = __float_as_int(hdr[5]);
    // q_sorted: qpos is a cell-ordered copy of the queries (x, y, z, original local index): neighbouring lanes
    // then walk the same cells (coherent loads, similar trip counts); results go to the original rows.
    const float4 q = __ldg(qpos + (size_t)b * Nq + qi);
    if (q_sorted) qi = __float_as_int(q.w);

    // unclamped cell of the query and its distance to the nearest edge of that cell
    const float ux = (q.x - ox) * inv, uy = (q.y - oy) * inv;
    const int ix = (int)floorf(ux), iy = (int)floorf(uy);
    const float fx = ux - floorf(ux), fy = uy - floorf(uy);  // in [0,1)
    const float edge = fminf(fminf(fx, 1.f - fx), fminf(fy, 1.f - fy)) * cell;
    const float eps = 1e-3f * cell;  // >> fp32 rounding of (x - ox) * inv for |cell index| <= 64

    Top3 t;
    t.init();
    // R = 1 block first (three contiguous row spans), then rings
    int R = 1;
    {
        const int x0 = max(ix - 1, 0), x1 = min(ix + 1, gx - 1);
        if (x0 <= x1)
            for (int y = max(iy - 1, 0); y <= min(iy + 1, gy - 1); ++y)
                scan_range(so, __ldg(cs + y * gx + x0), __ldg(cs + y * gx + x1 + 1), q, t);
    }
    while (true) {
        // everything within Chebyshev distance R (in cells) has been visited
        const float margin = fmaxf((float)R * cell + edge - eps, 0.f);
        if (t.i2 != 0x7fffffff && t.d2 < margin * margin) break;
        if (ix - R <= 0 && iy - R <= 0 && ix + R >= gx - 1 && iy + R >= gy - 1) break;  // whole grid visited
        ++R;
        const int x0 = max(ix - R, 0), x1 = min(ix + R, gx - 1);
        if (x0 <= x1) {
            const int yt = iy - R, yb = iy + R;
            if (yt >= 0 && yt < gy) scan_range(so, __ldg(cs + yt * gx + x0), __ldg(cs + yt * gx + x1 + 1), q, t);
            if (yb >= 0 && yb < gy) scan_range(so, __ldg(cs + yb * gx + x0), __ldg(cs + yb * gx + x1 + 1), q, t);
        }
        const int xl = ix - R, xr = ix + R;
        for (int y = max(iy - R + 1, 0); y <= min(iy + R - 1, gy - 1); ++y) {
            if (xl >= 0 && xl < gx) scan_range(so, __ldg(cs + y * gx + xl), __ldg(cs + y * gx + xl + 1), q, t);
            if (xr >= 0 && xr < gx) scan_range(so, __ldg(cs + y * gx + xr), __ldg(cs + y * gx + xr + 1), q, t);
        }
    }
    const size_t o = ((size_t)b * Nq + qi) * 3;
    const int gb = b * Ms;
    nbr[o] = gb + t.i0;
    nbr[o + 1] = gb + t.i1;
    nbr[o + 2] = gb + t.i2;
    wgt[o] = __fdiv_rn(1.0f, fmaxf(t.d0, 1e-16f));
    wgt[o + 1] = __fdiv_rn(1.0f, fmaxf(t.d1, 1e-16f));
    wgt[o + 2] = __fdiv_rn(1.0f, fmaxf(t.d2, 1e-16f));
}

}  // namespace sn2

extern "C" int sn2_knn3_grid(const float *grid_hdr, const int *cell_start, const float *sorted4, const float *qpos4,
                             int B, int Ms, int Nq, int *nbr, float *w, int q_sorted, void *stream)
{
    if (!grid_hdr || !cell_start || !sorted4 || !qpos4 || !nbr || !w || B <= 0 || Ms < 3 || Nq <= 0) return SN2_EINVAL;
    dim3 grid((Nq + sn2::KG_THREADS - 1) / sn2::KG_THREADS, B);
    sn2::knn3_grid_kernel<<<grid, sn2::KG_THREADS, 0, (cudaStream_t)stream>>>(
        grid_hdr, cell_start, reinterpret_cast<const float4 *>(sorted4), reinterpret_cast<const float4 *>(qpos4), Ms, Nq,
        nbr, w, q_sorted);
    SN2_LAUNCH_CHECK("knn3_grid_kernel");
    return SN2_OK;
}
