// Parcel-scale preparation and finalisation around the hot path (BASELINE config 4).
//
// parcel_grid / extract_plots  replace, for one parcel cloud resident in HBM, the reference's per-plot CPU chain
//     cKDTree ball query r = 10 m           inference/prepare_utils.py:47-53, prepare.py:75-76
//     drop plots with <= 50 points           inference/prepare_utils.py:67-69, prepare.py:91-94
//     z -= min z within 1.5 m (xy radius)    utils/load_data.py:228-249  (sklearn radius_neighbors, Python loop per point)
//     centre on the plot, add the 316 fake ground points, keep xyz, rescale, sample to subsample_size
//                                            data_loader/loader.py:73-105, 127-158, 233-255
//   and the pickle round trip between prepare.py and predict.py: the output IS the model input, (B,3,S) xyz and
//   (B,10,S) cloud, fp32, on the device.
//   Arithmetic follows the reference's dtypes: the LAS cloud is float32 (utils/load_data.py:163-178), the two radius
//   tests are evaluated in float64 on those fp32 values (scipy / sklearn convert to double), everything else is fp32
//   op by op (NumPy 1.x: a float32 array minus a float64 scalar is a float32 operation on the rounded scalar).
//   Canonical choices where the reference is order-free or random: points of a plot in ascending parcel index;
//   sub-sampling by a counter-based hash instead of np.random (S smallest hash values, order kept; up-sampling picks
//   hash % n) -- oracle/parcel_port.py restates the same rule.
//
// finalize_mosaic  the band finalisation after local-map fusion (inference/geotiff_raster.py:121-146, 273-291): the
//   hard medium-vegetation band (the threshold among linspace(0,1,10001) whose hard coverage is closest to the soft
//   one: a histogram + suffix sum instead of 10 001 passes over the raster) and the NaN rules.
#include "sn2_common.cuh"

namespace sn2 {

// ---------------------------------------------------------------------------------------------------------------
// uniform xy grid over the parcel: counting sort of point indices by cell (order inside a cell arbitrary)
// ---------------------------------------------------------------------------------------------------------------
struct PGrid {
    float x0, y0, inv_cell;
    int nx, ny;
};

__device__ __forceinline__ int pgrid_cell(const PGrid &g, float x, float y)
{
    int cx = (int)floorf((x - g.x0) * g.inv_cell), cy = (int)floorf((y - g.y0) * g.inv_cell);
    cx = min(max(cx, 0), g.nx - 1);
    cy = min(max(cy, 0), g.ny - 1);
    return cy * g.nx + cx;
}

__global__ void __launch_bounds__(256)
pgrid_count_kernel(const float *__restrict__ x, const float *__restrict__ y, long long P, PGrid g, int *__restrict__ cell_of,
                   int *__restrict__ count)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (long long)gridDim.x * blockDim.x) {
        const int c = pgrid_cell(g, __ldg(x + i), __ldg(y + i));
        cell_of[i] = c;
        atomicAdd(count + c, 1);
    }
}

// exclusive scan of `count` [n] -> start [n+1]; one CTA (n ~ 10^4..10^5 cells); cursor = copy of start for the scatter
__global__ void __launch_bounds__(1024)
pgrid_scan_kernel(const int *__restrict__ count, int n, int *__restrict__ start, int *__restrict__ cursor)
{
    __shared__ int warp_tot[32];
    __shared__ int carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + tid;
        const int v = i < n ? count[i] : 0;
        int s = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(SN2_FULL, s, d);
            if (lane >= d) s += t;
        }
        if (lane == 31) warp_tot[warp] = s;
        __syncthreads();
        if (warp == 0) {
            int w = warp_tot[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(SN2_FULL, w, d);
                if (lane >= d) w += t;
            }
            warp_tot[lane] = w;
        }
        __syncthreads();
        const int excl = carry + (warp ? warp_tot[warp - 1] : 0) + s - v;
        if (i < n) {
            start[i] = excl;
            cursor[i] = excl;
        }
        __syncthreads();
        if (tid == 1023) carry = excl + v;
        __syncthreads();
    }
    if (tid == 0) start[n] = carry;
}

// cell-grouped copy of the cloud: sorted4[k] = (x, y, z, parcel index as int bits); pos_of[parcel index] = k
__global__ void __launch_bounds__(256)
pgrid_scatter_kernel(const float *__restrict__ x, const float *__restrict__ y, const float *__restrict__ z,
                     const int *__restrict__ cell_of, long long P, int *__restrict__ cursor, float4 *__restrict__ sorted4,
                     int *__restrict__ pos_of)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (long long)gridDim.x * blockDim.x) {
        const int k = atomicAdd(cursor + __ldg(cell_of + i), 1);
        sorted4[k] = make_float4(__ldg(x + i), __ldg(y + i), __ldg(z + i), __int_as_float((int)i));
        pos_of[i] = k;
    }
}

// d2(p, q) <= r2 evaluated as the reference does (float64 on the fp32 coordinates), with an fp32 pre-test that settles
// every pair that is not within 1e-5 (relative) of the boundary
__device__ __forceinline__ bool within_f64(float px, float py, float qx, float qy, float r2f, double r2)
{
    const float dx = qx - px, dy = qy - py;
    const float d2 = fmaf(dx, dx, dy * dy);
    if (d2 < r2f * 0.99999f) return true;
    if (d2 > r2f * 1.00001f + 1e-30f) return false;
    const double ex = (double)qx - (double)px, ey = (double)qy - (double)py;
    return ex * ex + ey * ey <= r2;
}
__device__ __forceinline__ bool within_center_f64(double cx, double cy, float qx, float qy, float r2f, double r2)
{
    const float dx = qx - (float)cx, dy = qy - (float)cy;  // (float)cx is off by <= half an ulp of cx: covered by the margin below
    const float d2 = fmaf(dx, dx, dy * dy);
    const float slack = 1e-5f * r2f + 4.f * sqrtf(r2f) * (fabsf((float)cx) + fabsf((float)cy)) * 1.2e-7f;
    if (d2 < r2f - slack) return true;
    if (d2 > r2f + slack) return false;
    const double ex = (double)qx - cx, ey = (double)qy - cy;
    return ex * ex + ey * ey <= r2;
}

// ---------------------------------------------------------------------------------------------------------------
// plot extraction: one CTA per plot centre
// ---------------------------------------------------------------------------------------------------------------
constexpr int EX_T = 1024;
constexpr int EX_BITS = 65536;       // points under one disk's bounding square (with a one-cell margin)
constexpr int EX_ROWS = 32;          // grid rows under it
constexpr int EX_CAP = 24576;        // points of the parcel inside one plot disk (32 pts/m^2 -> ~10 000; 8 B of smem each)

// counter-based hash (two rounds of a 32-bit mix): the canonical replacement for np.random in sample_cloud
__host__ __device__ __forceinline__ unsigned sample_hash(unsigned plot_seed, unsigned j)
{
    unsigned h = plot_seed * 0x9E3779B1u + j * 0x85EBCA77u + 0x165667B1u;
    h ^= h >> 15; h *= 0x2C1B3C6Du;
    h ^= h >> 12; h *= 0x297A2D39u;
    h ^= h >> 15;
    return h;
}

// in-place ascending bitonic sort of n2 (power of two) keys in shared memory
__device__ void bitonic_sort_u32(unsigned *keys, int n2)
{
    for (int k = 2; k <= n2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n2; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned a = keys[i], b = keys[ixj];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) {
                        keys[i] = b;
                        keys[ixj] = a;
                    }
                }
            }
            __syncthreads();
        }
    }
}

__device__ __forceinline__ int block_excl_scan(int v, int *warp_tot, int &total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int s = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(SN2_FULL, s, d);
        if (lane >= d) s += t;
    }
    __syncthreads();
    if (lane == 31) warp_tot[warp] = s;
    __syncthreads();
    if (warp == 0) {
        int w = warp_tot[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(SN2_FULL, w, d);
            if (lane >= d) w += t;
        }
        warp_tot[lane] = w;
    }
    __syncthreads();
    total = warp_tot[31];
    return (warp ? warp_tot[warp - 1] : 0) + s - v;
}

struct ExtractArgs {
    const float *x, *y, *z;      // parcel cloud rows 0..2 (fp32, absolute coordinates, metres)
    const float *feat;           // rows 3..9: [7][P] (R, G, B, NIR, intensity, return_num, num_returns)
    long long P;
    PGrid grid;                  // cell side >= 1.05 * znorm_radius
    const int *cell_start, *pos_of;
    const float4 *sorted4;
    const double *centers;       // [C][2] float64
    const unsigned *seeds;       // [C] per-plot seed of the sampling hash
    int C, S;
    float radius, znorm_radius, z_max;
    int diam_meters, min_points;
    float *out_xyz;              // [C][3][S]
    float *out_cloud;            // [C][10][S]
    int *out_n;                  // [C] points of the parcel inside the disk (before fake points); < 0: over capacity
    int *out_src;                // nullable [C][S]: parcel index of every output point (-1 - k for fake ground point k)
};

__global__ void __launch_bounds__(EX_T, 1)
extract_plots_kernel(ExtractArgs a)
{
    extern __shared__ __align__(16) unsigned char ex_smem[];
    unsigned *idx = reinterpret_cast<unsigned *>(ex_smem);                  // [EX_CAP] parcel indices, sorted ascending
    float *zmin = reinterpret_cast<float *>(idx + EX_CAP);                  // [EX_CAP] local minimum of z per point
    __shared__ int s_n, warp_tot[32];
    __shared__ unsigned hist[256];
    __shared__ unsigned s_prefix, s_need;
    // in-disk flag of every point under the disk's bounding square: one bit per position of the cell-grouped copy, row by row
    __shared__ unsigned inbits[EX_BITS / 32];
    __shared__ int row_start[EX_ROWS], row_bit0[EX_ROWS + 1];

    const int c = blockIdx.x, tid = threadIdx.x;
    const double cx = a.centers[2 * c], cy = a.centers[2 * c + 1];
    const double r2 = (double)a.radius * (double)a.radius;
    const float r2f = a.radius * a.radius;
    const PGrid g = a.grid;
    if (tid == 0) s_n = 0;
    __syncthreads();

    // ---- 1. ball query: the grid rows under the disk's bounding square; the cells of a row are one contiguous range of
    //         the cell-grouped copy, read as float4 (coalesced).  The in-disk flags are kept as a bit set for step 3. ------
    int cx0 = (int)floorf(((float)(cx - a.radius) - g.x0) * g.inv_cell) - 1, cx1 = (int)floorf(((float)(cx + a.radius) - g.x0) * g.inv_cell) + 1;
    int cy0 = (int)floorf(((float)(cy - a.radius) - g.y0) * g.inv_cell) - 1, cy1 = (int)floorf(((float)(cy + a.radius) - g.y0) * g.inv_cell) + 1;
    cx0 = max(cx0, 0); cy0 = max(cy0, 0); cx1 = min(cx1, g.nx - 1); cy1 = min(cy1, g.ny - 1);
    const int nrows = cy1 - cy0 + 1;
    if (nrows > EX_ROWS || cx1 < cx0 || nrows <= 0) {  // a disk far outside the grid, or cells much smaller than the kernel is laid out for
        if (tid == 0) a.out_n[c] = nrows > EX_ROWS ? -1 : 0;
        return;
    }
    if (tid == 0) {
        int bit = 0;
        for (int j = 0; j < nrows; ++j) {
            const int s = __ldg(a.cell_start + (cy0 + j) * g.nx + cx0), e = __ldg(a.cell_start + (cy0 + j) * g.nx + cx1 + 1);
            row_start[j] = s;
            row_bit0[j] = bit;
            bit += (e - s + 31) & ~31;
        }
        row_bit0[nrows] = bit;
    }
    for (int i = tid; i < EX_BITS / 32; i += EX_T) inbits[i] = 0u;
    __syncthreads();
    const int nbits = row_bit0[nrows];
    if (nbits > EX_BITS) {
        if (tid == 0) a.out_n[c] = -nbits;
        return;
    }
    for (int j = 0; j < nrows; ++j) {
        const int s = row_start[j], e = __ldg(a.cell_start + (cy0 + j) * g.nx + cx1 + 1);
        for (int k0 = s; k0 < e; k0 += EX_T) {
            const int k = k0 + tid;
            bool in = false;
            float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k < e) {
                q = __ldg(a.sorted4 + k);
                in = within_center_f64(cx, cy, q.x, q.y, r2f, r2);
            }
            const unsigned word = __ballot_sync(SN2_FULL, in);  // k - s is a multiple of 32 at lane 0: one word per warp
            if ((tid & 31) == 0 && word) inbits[(row_bit0[j] + (k - s)) >> 5] = word;
            if (in) {
                const int pos = atomicAdd(&s_n, 1);
                if (pos < EX_CAP) idx[pos] = (unsigned)__float_as_int(q.w);
            }
        }
    }
    __syncthreads();
    const int n = s_n;
    if (n > EX_CAP) {
        if (tid == 0) a.out_n[c] = -n;
        return;
    }
    if (tid == 0) a.out_n[c] = n;
    if (n <= a.min_points) return;  // prepare_utils.py:67-69 (< 50 -> None) and prepare.py:91-94 (keeps > 50 only)

    // ---- 2. canonical order: ascending parcel index -------------------------------------------------------------
    int n2 = 1;
    while (n2 < n) n2 <<= 1;  // <= 32768: the padding may run into the zmin area behind idx, which is not in use yet
    for (int i = n + tid; i < n2; i += EX_T) idx[i] = 0xffffffffu;
    __syncthreads();
    bitonic_sort_u32(idx, n2);

    // ---- 3. z normalisation: z - min z over the PLOT's points within znorm_radius in xy (load_data.py:237-249) ---------
    // One warp per parcel cell under the disk: lanes = the cell's points, and the candidates -- the 3 x 3 cells around it
    // (cell side >= 1.05 * znorm_radius), three contiguous ranges of the cell-grouped copy -- are walked by the whole warp
    // together (every lane reads the same float4: one broadcast load, no divergence).  A candidate counts only if it lies in
    // this plot's disk too (the reference searches inside the extracted cloud); that test is lane independent.
    {
        const double zr2 = (double)a.znorm_radius * (double)a.znorm_radius;
        const float zr2f = a.znorm_radius * a.znorm_radius;
        const int lane = tid & 31, warp = tid >> 5;
        const int ncx = cx1 - cx0 + 1, ncells = ncx * nrows;
        for (int cell = warp; cell < ncells; cell += EX_T / 32) {
            const int ry = cell / ncx, iy = cy0 + ry, ix = cx0 + cell % ncx;
            const int s = __ldg(a.cell_start + iy * g.nx + ix), e = __ldg(a.cell_start + iy * g.nx + ix + 1);
            const int bit_s = row_bit0[ry] + (s - row_start[ry]);
            for (int base = s; base < e; base += 32) {
                const int k = base + lane, b = bit_s + (k - s);
                const bool mine = k < e && ((inbits[b >> 5] >> (b & 31)) & 1u);
                if (!__any_sync(SN2_FULL, mine)) continue;
                float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
                if (mine) p = __ldg(a.sorted4 + k);
                float m = p.z;
                for (int jr = max(ry - 1, 0); jr <= min(ry + 1, nrows - 1); ++jr) {
                    const int jy = cy0 + jr;
                    const int rs = __ldg(a.cell_start + jy * g.nx + max(ix - 1, cx0)), re = __ldg(a.cell_start + jy * g.nx + min(ix + 1, cx1) + 1);
                    const int bit_r = row_bit0[jr] + (rs - row_start[jr]);
                    // 32 candidates per step: one coalesced float4 load per lane, then the in-disk ones are handed round by
                    // shuffle (a per-candidate broadcast load was latency bound: the 196 KB of shared memory leave ~30 KB of L1)
                    for (int kb = rs; kb < re; kb += 32) {
                        const int kk = kb + lane, bb = bit_r + (kk - rs);
                        const bool cin = kk < re && ((inbits[bb >> 5] >> (bb & 31)) & 1u);
                        float4 q = make_float4(0.f, 0.f, INFINITY, 0.f);
                        if (cin) q = __ldg(a.sorted4 + kk);
                        // a candidate can only lower somebody's minimum if it lies below the LARGEST current minimum of the
                        // warp's points: once every lane has met a ground return most candidates drop out here
                        const float mtop = warp_max(mine ? m : -INFINITY);
                        unsigned cm = __ballot_sync(SN2_FULL, cin && q.z < mtop);
                        while (cm) {  // warp-uniform
                            const int t = __ffs(cm) - 1;
                            cm &= cm - 1;
                            const float qx = __shfl_sync(SN2_FULL, q.x, t), qy = __shfl_sync(SN2_FULL, q.y, t), qz = __shfl_sync(SN2_FULL, q.z, t);
                            if (mine && qz < m && within_f64(p.x, p.y, qx, qy, zr2f, zr2)) m = qz;
                        }
                    }
                }
                if (mine) {  // position of this point in the plot's ascending index list
                    const unsigned key = (unsigned)__float_as_int(p.w);
                    int lo = 0, hi = n - 1;
                    while (lo < hi) {
                        const int mid = (lo + hi) >> 1;
                        if (idx[mid] < key) lo = mid + 1;
                        else hi = mid;
                    }
                    zmin[lo] = m;
                }
            }
        }
    }
    __syncthreads();

    // ---- 4. the 316 fake ground points + sub-sampling to S ---------------------------------------------------------
    // combined list: j < n -> plot point j; j >= n -> fake point (j - n), enumerated like np.meshgrid(x, y).flatten()
    const int D = a.diam_meters, half = D / 2;
    // count the fake points (r < D // 2) -- 316 for D = 20
    int nfake = 0;
    for (int t = 0; t < D * D; ++t) {
        const float fx = (float)(t % D - half) + 0.5f, fy = (float)(t / D - half) + 0.5f;
        nfake += sqrtf(fx * fx + fy * fy) < (float)half;
    }
    const int ntot = n + nfake, S = a.S;
    const unsigned seed = a.seeds[c];
    float *oxyz = a.out_xyz + (size_t)c * 3 * S, *ocl = a.out_cloud + (size_t)c * 10 * S;
    const float cxf = (float)cx, cyf = (float)cy;  // NumPy 1.x: float32 array - float64 scalar -> float32 arithmetic

    auto emit = [&](int j, int slot) {
        float x, y, z, f[7];
        int src;
        if (j < n) {
            const int p = idx[j];
            x = __fsub_rn(__ldg(a.x + p), cxf);
            y = __fsub_rn(__ldg(a.y + p), cyf);
            z = __fsub_rn(__ldg(a.z + p), zmin[j]);
#pragma unroll
            for (int q = 0; q < 7; ++q) f[q] = __ldg(a.feat + (size_t)q * a.P + p);
            src = p;
        } else {
            int k = j - n, t = 0, seen = 0;  // k-th fake point in flatten order
            for (; t < D * D; ++t) {
                const float fx = (float)(t % D - half) + 0.5f, fy = (float)(t / D - half) + 0.5f;
                if (sqrtf(fx * fx + fy * fy) < (float)half) {
                    if (seen == k) break;
                    ++seen;
                }
            }
            x = (float)(t % D - half) + 0.5f;
            y = (float)(t / D - half) + 0.5f;
            z = 0.f;
#pragma unroll
            for (int q = 0; q < 7; ++q) f[q] = 0.f;
            src = -1 - k;
        }
        oxyz[slot] = x; oxyz[S + slot] = y; oxyz[2 * S + slot] = z;
        // rescale_cloud (loader.py:135-158), fp32 op by op
        ocl[slot] = __fdiv_rn(x, 10.f);
        ocl[S + slot] = __fdiv_rn(y, 10.f);
        ocl[2 * S + slot] = __fdiv_rn(z, a.z_max);
#pragma unroll
        for (int q = 0; q < 4; ++q) ocl[(3 + q) * S + slot] = __fdiv_rn(f[q], 65536.f);
        ocl[7 * S + slot] = __fdiv_rn(f[4], 32768.f);
        ocl[8 * S + slot] = __fdiv_rn(__fsub_rn(f[5], 1.f), 6.f);
        ocl[9 * S + slot] = __fdiv_rn(__fsub_rn(f[6], 1.f), 6.f);
        if (a.out_src) a.out_src[(size_t)c * S + slot] = src;
    };

    if (ntot <= S) {  // keep everything, then (S - ntot) picks with replacement: j = hash % ntot (loader.py:239-244)
        for (int j = tid; j < ntot; j += EX_T) emit(j, j);
        for (int k = tid; k < S - ntot; k += EX_T) emit((int)(sample_hash(seed ^ 0x5bd1e995u, (unsigned)k) % (unsigned)ntot), ntot + k);
        return;
    }
    // S smallest hash values of j in [0, ntot), ties by lower j: radix select of the S-th smallest key, 8 bits per pass
    if (tid == 0) { s_prefix = 0u; s_need = (unsigned)S; }
    __syncthreads();
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        for (int i = tid; i < 256; i += EX_T) hist[i] = 0u;
        __syncthreads();
        const unsigned prefix = s_prefix, mask = pass ? (0xffffffffu << (shift + 8)) : 0u;
        for (int j = tid; j < ntot; j += EX_T) {
            const unsigned h = sample_hash(seed, (unsigned)j);
            if ((h & mask) == prefix) atomicAdd(&hist[(h >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            unsigned need = s_need, b = 0;
            for (; b < 256; ++b) {
                if (hist[b] >= need) break;
                need -= hist[b];
            }
            s_prefix = prefix | (b << shift);
            s_need = need;
        }
        __syncthreads();
    }
    const unsigned T = s_prefix, take_eq = s_need;  // keys < T all taken; of the keys == T the first take_eq by j
    int base_lt = 0, base_eq = 0;
    for (int j0 = 0; j0 < ntot; j0 += EX_T) {
        const int j = j0 + tid;
        const unsigned h = j < ntot ? sample_hash(seed, (unsigned)j) : 0xffffffffu;
        const int lt = j < ntot && h < T, eq = j < ntot && h == T;
        int tot_lt, tot_eq;
        const int r_lt = block_excl_scan(lt, warp_tot, tot_lt);
        const int r_eq = block_excl_scan(eq, warp_tot, tot_eq);
        const bool take = lt || (eq && (unsigned)(base_eq + r_eq) < take_eq);
        // output slot = number of taken elements before j = (#lt before) + min(#eq before, take_eq)
        if (take) emit(j, base_lt + r_lt + (int)min((unsigned)(base_eq + r_eq), take_eq));
        base_lt += tot_lt;
        base_eq += tot_eq;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// band finalisation of the fused mosaic (geotiff_raster.py:121-146, 273-291)
// ---------------------------------------------------------------------------------------------------------------
constexpr int HV_BINS = 10001;

// lin[i] of np.linspace(0, 1, 10001): i * step with step = 1/10000 (float64), last element exactly 1
__device__ __forceinline__ double hv_lin(int i) { return i == HV_BINS - 1 ? 1.0 : (double)i * (1.0 / 10000.0); }

// k(v) = number of thresholds lin[i] with lin[i] < v  (so that v > lin[i]  <=>  i < k(v))
__device__ __forceinline__ int hv_rank(double v)
{
    if (!(v > 0.0)) return 0;
    int k = (int)fmin(floor(v * 10000.0), (double)(HV_BINS - 1));
    while (k > 0 && !(hv_lin(k - 1) < v)) --k;
    while (k < HV_BINS && hv_lin(k) < v) ++k;
    return k;
}

__global__ void __launch_bounds__(256)
hardveg_hist_kernel(const double *__restrict__ vm, long long HW, unsigned *__restrict__ hist, double *__restrict__ sum,
                    unsigned long long *__restrict__ nvalid)
{
    double s = 0.0;
    unsigned long long c = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += (long long)gridDim.x * blockDim.x) {
        const double v = vm[i];
        if (v == v) {
            s += v;
            ++c;
            atomicAdd(hist + hv_rank(v), 1u);
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        s += __shfl_xor_sync(SN2_FULL, s, d);
        c += __shfl_xor_sync(SN2_FULL, c, d);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(sum, s);
        atomicAdd(nvalid, c);
    }
}

// one CTA: count(v > lin[i]) = sum_{k > i} hist[k]; delta[i] = |target - count / nvalid|; first arg-min -> threshold
__global__ void __launch_bounds__(1024)
hardveg_threshold_kernel(const unsigned *__restrict__ hist, const double *__restrict__ sum, const unsigned long long *__restrict__ nvalid,
                         double *__restrict__ threshold_out)
{
    __shared__ unsigned suffix[HV_BINS + 1];
    __shared__ double best_d[32];
    __shared__ int best_i[32];
    const int tid = threadIdx.x;
    {   // suffix[i] = sum_{k > i} hist[k], i in [0, HV_BINS): per-thread chunks of 10, block scan of the chunk totals
        constexpr int CH = (HV_BINS + 1 + 1023) / 1024;
        __shared__ unsigned tot[1024];
        const int k0 = tid * CH;
        unsigned mine = 0;
        for (int k = k0; k < min(k0 + CH, HV_BINS + 1); ++k) mine += hist[k];
        tot[tid] = mine;
        __syncthreads();
        for (int d = 1; d < 1024; d <<= 1) {  // inclusive suffix scan of tot (sum over threads >= tid)
            const unsigned v = tid + d < 1024 ? tot[tid + d] : 0u;
            __syncthreads();
            tot[tid] += v;
            __syncthreads();
        }
        unsigned acc = tid + 1 < 1024 ? tot[tid + 1] : 0u;  // everything in later chunks
        for (int k = min(k0 + CH, HV_BINS + 1) - 1; k >= k0; --k) {
            acc += hist[k];                    // acc = sum_{k' >= k} hist[k']
            if (k >= 1) suffix[k - 1] = acc;
        }
    }
    __syncthreads();
    const double n = (double)*nvalid, target = *sum / n;
    double bd = INFINITY;
    int bi = HV_BINS;
    for (int i = tid; i < HV_BINS; i += 1024) {
        const double d = fabs(target - (double)suffix[i] / n);
        if (d < bd || (d == bd && i < bi)) { bd = d; bi = i; }
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        const double od = __shfl_xor_sync(SN2_FULL, bd, s);
        const int oi = __shfl_xor_sync(SN2_FULL, bi, s);
        if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
    }
    if ((tid & 31) == 0) { best_d[tid >> 5] = bd; best_i[tid >> 5] = bi; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < 32; ++w)
            if (best_d[w] < bd || (best_d[w] == bd && best_i[w] < bi)) { bd = best_d[w]; bi = best_i[w]; }
        threshold_out[0] = hv_lin(bi);
        threshold_out[1] = target;
    }
}

// in: fused [4][HW] = (Vb, Vm, Vh, weights), NaN = no value.  out [5][HW] = (Vb, Vm_soft, Vh, Vm_hard, weights):
// NaN -> 0 where at least one of the three coverages has a value, NaN in every band where none has (:273-291).
__global__ void __launch_bounds__(256)
finalize_bands_kernel(const double *__restrict__ in, long long HW, const double *__restrict__ threshold, double *__restrict__ out)
{
    const double thr = threshold[0];
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += (long long)gridDim.x * blockDim.x) {
        const double vb = in[i], vm = in[HW + i], vh = in[2 * HW + i], w = in[3 * HW + i];
        const bool none = vb != vb && vm != vm && vh != vh;
        const double hard = vm == vm ? (vm > thr ? 1.0 : 0.0) : nan;
        auto fix = [&](double v) { return none ? nan : (v == v ? v : 0.0); };
        out[i] = fix(vb);
        out[HW + i] = fix(vm);
        out[2 * HW + i] = fix(vh);
        out[3 * HW + i] = fix(hard);
        out[4 * HW + i] = fix(w);
    }
}

}  // namespace sn2

using namespace sn2;

extern "C" int sn2_parcel_grid_build(const float *x, const float *y, const float *z, long long P, float x0, float y0, float cell, int nx,
                                     int ny, int *cell_of, int *count, int *cell_start, int *cursor, float *sorted4, int *pos_of,
                                     void *stream)
{
    if (!x || !y || !z || !cell_of || !count || !cell_start || !cursor || !sorted4 || !pos_of || (reinterpret_cast<uintptr_t>(sorted4) & 15) || P <= 0 || P > 0x7fffffffll || nx <= 0 || ny <= 0 || !(cell > 0.f))
        return SN2_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    PGrid g{x0, y0, 1.f / cell, nx, ny};
    SN2_CUDA_TRY(cudaMemsetAsync(count, 0, sizeof(int) * (size_t)nx * ny, st), "parcel grid memset");
    const int grid = (int)min((P + 255) / 256, (long long)148 * 16);
    pgrid_count_kernel<<<grid, 256, 0, st>>>(x, y, P, g, cell_of, count);
    SN2_LAUNCH_CHECK("pgrid_count_kernel");
    pgrid_scan_kernel<<<1, 1024, 0, st>>>(count, nx * ny, cell_start, cursor);
    SN2_LAUNCH_CHECK("pgrid_scan_kernel");
    pgrid_scatter_kernel<<<grid, 256, 0, st>>>(x, y, z, cell_of, P, cursor, reinterpret_cast<float4 *>(sorted4), pos_of);
    SN2_LAUNCH_CHECK("pgrid_scatter_kernel");
    return SN2_OK;
}

extern "C" int sn2_plot_capacity(void) { return EX_CAP; }

extern "C" int sn2_extract_plots(const float *xyz, const float *feat, long long P, float x0, float y0, float cell, int nx, int ny,
                                 const int *cell_start, const float *sorted4, const int *pos_of, const double *centers, const unsigned *seeds, int C,
                                 int S, float radius, float znorm_radius, float z_max, int diam_meters, int min_points,
                                 float *out_xyz, float *out_cloud, int *out_n, int *out_src, void *stream)
{
    if (cell < 1.05f * znorm_radius) return SN2_EINVAL;  // the 3 x 3-cell neighbourhood must cover the z-normalisation radius
    if (!xyz || !feat || !cell_start || !sorted4 || !pos_of || !centers || !seeds || !out_xyz || !out_cloud || !out_n || C <= 0 || S <= 0 ||
        P <= 0 || diam_meters <= 0 || diam_meters > 64 || !(radius > 0.f) || !(znorm_radius > 0.f))
        return SN2_EINVAL;
    ExtractArgs a;
    a.x = xyz; a.y = xyz + P; a.z = xyz + 2 * P; a.feat = feat; a.P = P;
    a.grid = PGrid{x0, y0, 1.f / cell, nx, ny};
    a.cell_start = cell_start; a.sorted4 = reinterpret_cast<const float4 *>(sorted4); a.pos_of = pos_of; a.centers = centers; a.seeds = seeds; a.C = C; a.S = S;
    a.radius = radius; a.znorm_radius = znorm_radius; a.z_max = z_max; a.diam_meters = diam_meters; a.min_points = min_points;
    a.out_xyz = out_xyz; a.out_cloud = out_cloud; a.out_n = out_n; a.out_src = out_src;
    const size_t smem = (size_t)EX_CAP * (sizeof(unsigned) + sizeof(float));
    auto kern = extract_plots_kernel;
    SN2_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "extract_plots attr");
    kern<<<C, EX_T, smem, (cudaStream_t)stream>>>(a);
    SN2_LAUNCH_CHECK("extract_plots_kernel");
    return SN2_OK;
}

extern "C" unsigned sn2_sample_hash(unsigned seed, unsigned j) { return sample_hash(seed, j); }

extern "C" int sn2_finalize_mosaic(const double *fused, int H, int W, unsigned *hist, double *scratch, double *out, void *stream)
{
    if (!fused || !hist || !scratch || !out || H <= 0 || W <= 0) return SN2_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    const long long HW = (long long)H * W;
    SN2_CUDA_TRY(cudaMemsetAsync(hist, 0, sizeof(unsigned) * (HV_BINS + 1), st), "finalize_mosaic memset hist");
    SN2_CUDA_TRY(cudaMemsetAsync(scratch, 0, sizeof(double) * 4, st), "finalize_mosaic memset scratch");
    const int grid = (int)min((HW + 255) / 256, (long long)148 * 8);
    hardveg_hist_kernel<<<grid, 256, 0, st>>>(fused + HW, HW, hist, scratch, reinterpret_cast<unsigned long long *>(scratch + 1));
    SN2_LAUNCH_CHECK("hardveg_hist_kernel");
    hardveg_threshold_kernel<<<1, 1024, 0, st>>>(hist, scratch, reinterpret_cast<const unsigned long long *>(scratch + 1), scratch + 2);
    SN2_LAUNCH_CHECK("hardveg_threshold_kernel");
    finalize_bands_kernel<<<grid, 256, 0, st>>>(fused, HW, scratch + 2, out);
    SN2_LAUNCH_CHECK("finalize_bands_kernel");
    return SN2_OK;
}
