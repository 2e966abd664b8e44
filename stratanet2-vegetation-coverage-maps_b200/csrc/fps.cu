// K0 ingest + K1 farthest point sampling.
//
// FPS: one CTA per plot.  The running min-distance of every point lives in REGISTERS (PPT points per
// thread, point i = k*THREADS + tid so that ascending k is ascending index), coordinates live in
// registers too when they fit (REGXYZ) or are re-read from a SoA copy in shared memory.  Each
// iteration: update dist, per-thread arg-max with strict '>' (lowest index wins), warp arg-max with
// two REDUX (max of the dist bits -- dist >= 0 so the uint order is the float order -- then min of
// the candidate indices), one __syncthreads over double-buffered per-warp slots, and every warp
// re-reduces the slots itself, so there is exactly one barrier per sample.
#include "sn2_common.cuh"

namespace sn2 {

__global__ void ingest_kernel(const float *__restrict__ xyz, const float *__restrict__ cloud, int B, int N,
                              int F, float4 *__restrict__ pos4, float4 *__restrict__ feat)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)B * N;
    if (i >= total) return;
    int b = (int)(i / N);
    int n = (int)(i - (long long)b * N);
    const float *px = xyz + (size_t)b * 3 * N + n;
    pos4[i] = make_float4(__ldg(px), __ldg(px + N), __ldg(px + 2 * (size_t)N), 0.f);
    const float *pc = cloud + (size_t)b * F * N + n;
    float v[SN2_F0];
#pragma unroll
    for (int c = 0; c < SN2_F0; ++c) v[c] = __ldg(pc + (size_t)(c + 2) * N);
    feat[2 * i] = make_float4(v[0], v[1], v[2], v[3]);
    feat[2 * i + 1] = make_float4(v[4], v[5], v[6], v[7]);
}

template <int THREADS, int PPT, bool REGXYZ>
__global__ void __launch_bounds__(THREADS, 1)
fps_kernel(const float4 *__restrict__ pos, int N, int M, const int *__restrict__ start,
           int *__restrict__ idx_out, float4 *__restrict__ pos_out)
{
    constexpr int NW = THREADS / 32;
    extern __shared__ float fps_smem[];
    float *sx = fps_smem;
    float *sy = sx + THREADS * PPT;
    float *sz = sy + THREADS * PPT;
    __shared__ unsigned slot_v[2][32];
    __shared__ unsigned slot_i[2][32];

    const int b = blockIdx.x;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const float4 *p = pos + (size_t)b * N;

    float px[REGXYZ ? PPT : 1], py[REGXYZ ? PPT : 1], pz[REGXYZ ? PPT : 1];
    float dist[PPT];
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
        int i = k * THREADS + tid;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < N) v = __ldg(p + i);
        // padding points carry dist = 0 forever and an index >= N, so they can only tie with real
        // points at distance 0 and then lose to the lower (real) index.
        dist[k] = (i < N) ? __int_as_float(0x7f800000) : 0.f;
        sx[i] = v.x;
        sy[i] = v.y;
        sz[i] = v.z;
        if (REGXYZ) {
            px[k] = v.x;
            py[k] = v.y;
            pz[k] = v.z;
        }
    }
    int last = start ? start[b] : 0;
    if (tid == 0) {
        idx_out[(size_t)b * M] = b * N + last;
        if (pos_out) pos_out[(size_t)b * M] = __ldg(p + last);
    }
    __syncthreads();

    int buf = 0;
    for (int it = 1; it < M; ++it) {
        const float lx = sx[last], ly = sy[last], lz = sz[last];
        float best = -1.f;
        int bk = 0;
#pragma unroll
        for (int k = 0; k < PPT; ++k) {
            float d;
            if (REGXYZ) {
                d = dist2(px[k], py[k], pz[k], lx, ly, lz);
            } else {
                int i = k * THREADS + tid;
                d = dist2(sx[i], sy[i], sz[i], lx, ly, lz);
            }
            float dk = fminf(dist[k], d);
            dist[k] = dk;
            if (dk > best) {
                best = dk;
                bk = k;
            }
        }
        const unsigned vb = __float_as_uint(best);
        const unsigned gi = (unsigned)(bk * THREADS + tid);
        const unsigned wm = __reduce_max_sync(SN2_FULL, vb);
        const unsigned wi = __reduce_min_sync(SN2_FULL, vb == wm ? gi : 0xffffffffu);
        if (lane == 0) {
            slot_v[buf][warp] = wm;
            slot_i[buf][warp] = wi;
        }
        __syncthreads();
        const unsigned v2 = lane < NW ? slot_v[buf][lane] : 0u;
        const unsigned i2 = lane < NW ? slot_i[buf][lane] : 0xffffffffu;
        const unsigned m2 = __reduce_max_sync(SN2_FULL, v2);
        const unsigned w2 = __reduce_min_sync(SN2_FULL, (lane < NW && v2 == m2) ? i2 : 0xffffffffu);
        last = (int)w2;
        if (tid == 0) {
            idx_out[(size_t)b * M + it] = b * N + last;
            if (pos_out) pos_out[(size_t)b * M + it] = make_float4(sx[last], sy[last], sz[last], 0.f);
        }
        buf ^= 1;
    }
}

template <int THREADS, int PPT, bool REGXYZ>
static int launch_fps(const float4 *pos, int B, int N, int M, const int *start, int *idx, float4 *pos_out,
                      cudaStream_t st)
{
    size_t smem = (size_t)3 * THREADS * PPT * sizeof(float);
    auto kern = fps_kernel<THREADS, PPT, REGXYZ>;
    SN2_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "fps attr");
    kern<<<B, THREADS, smem, st>>>(pos, N, M, start, idx, pos_out);
    SN2_LAUNCH_CHECK("fps_kernel");
    return SN2_OK;
}

}  // namespace sn2

extern "C" int sn2_fps_max_points(void) { return 16384; }

extern "C" int sn2_ingest(const float *xyz, const float *cloud, int B, int N, int F, float *pos4, float *feat,
                          void *stream)
{
    if (!xyz || !cloud || !pos4 || !feat || B <= 0 || N <= 0) return SN2_EINVAL;
    if (F != SN2_F0 + 2) return SN2_EUNSUPPORTED;
    long long total = (long long)B * N;
    int threads = 256;
    long long blocks = (total + threads - 1) / threads;
    sn2::ingest_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(
        xyz, cloud, B, N, F, reinterpret_cast<float4 *>(pos4), reinterpret_cast<float4 *>(feat));
    SN2_LAUNCH_CHECK("ingest_kernel");
    return SN2_OK;
}

extern "C" int sn2_fps(const float *pos4, int B, int N, int M, const int *start, int *idx_out, float *pos4_out,
                       void *stream)
{
    if (!pos4 || !idx_out || B <= 0 || N <= 0 || M <= 0 || M > N) return SN2_EINVAL;
    if (N > sn2_fps_max_points()) return SN2_EUNSUPPORTED;
    const float4 *p = reinterpret_cast<const float4 *>(pos4);
    float4 *po = reinterpret_cast<float4 *>(pos4_out);
    cudaStream_t st = (cudaStream_t)stream;
    using namespace sn2;
    if (N <= 256) return launch_fps<128, 2, true>(p, B, N, M, start, idx_out, po, st);
    if (N <= 1024) return launch_fps<256, 4, true>(p, B, N, M, start, idx_out, po, st);
    if (N <= 2560) return launch_fps<512, 5, true>(p, B, N, M, start, idx_out, po, st);
    if (N <= 4096) return launch_fps<512, 8, true>(p, B, N, M, start, idx_out, po, st);
    if (N <= 8192) return launch_fps<1024, 8, true>(p, B, N, M, start, idx_out, po, st);
    if (N <= 10240) return launch_fps<1024, 10, false>(p, B, N, M, start, idx_out, po, st);
    return launch_fps<1024, 16, false>(p, B, N, M, start, idx_out, po, st);
}
