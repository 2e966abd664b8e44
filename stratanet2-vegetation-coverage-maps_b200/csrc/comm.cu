// Peer-memory collectives for data-parallel training (SURVEY.md §8e, §5 "distributed communication backend").
//
// The training step couples the ranks in three places only, all of them tiny and latency bound: the BatchNorm batch
// statistics of every Linear -> ReLU -> BatchNorm block (a [2C+1] fp64 vector forward, a [2C] one backward, 7 blocks)
// and the flat 14 997-float gradient.  As NCCL calls that is ~15 collectives of a few hundred bytes per step, each a
// kernel launch with its own protocol set-up, sequentially dependent on the kernels around them.  Here the reduction
// is a device function that runs INSIDE the kernel that produces or consumes the values (BatchNorm finalize, BatchNorm
// backward coefficients, the Adam update): every rank pushes its vector straight into every peer's memory over NVLink
// (plain stores through the NVSwitch fabric; the buffers are cudaMalloc'd here and mapped into the peers with CUDA IPC),
// raises a flag there, waits for the flags of its peers in its OWN memory and sums the copies in rank order -- one
// one-shot all-reduce, no extra launch, bitwise identical on every rank (same summation order), capturable in a CUDA
// graph (the sequence number lives in device memory).
//
// Protocol.  Every rank owns one region:  header {seq, error} | flags [SLOTS][MAXW] u64 | data [SLOTS][MAXW][SLOT_BYTES].
// Collective number s (1, 2, ...) uses slot s % SLOTS:  rank r stores its payload into data[slot][r] of EVERY rank,
// __threadfence_system, then st.release.sys flags[slot][r] = s on every rank; it then spins (ld.acquire.sys) on the
// flags of its own region until all equal s and sums data[slot][0..world) in rank order.  A rank can run at most one
// collective ahead of the slowest rank (it needs that rank's flag to finish), so a slot is never overwritten while it
// is still being read as long as SLOTS >= 2.  All collectives of a rank must be issued on ONE stream, in the same
// order on every rank.  A wait that exceeds ~20 s sets the sticky `error` word (later collectives stop waiting) so a
// crashed peer can never hang the GPU; the host polls it with sn2_comm_status.
#include "sn2_common.cuh"

namespace sn2 {

constexpr int COMM_MAXW = 8;
constexpr int COMM_SLOTS = 4;
constexpr int COMM_SLOT_BYTES = 64 * 1024;
constexpr size_t COMM_HDR = 4096;                     // header page: seq @0, error @8
constexpr size_t COMM_FLAGS = 4096;                   // flags page
constexpr size_t COMM_DATA_OFF = COMM_HDR + COMM_FLAGS;
constexpr size_t COMM_BYTES = COMM_DATA_OFF + (size_t)COMM_SLOTS * COMM_MAXW * COMM_SLOT_BYTES;
constexpr long long COMM_TIMEOUT_CYCLES = 40000000000ll;  // ~20 s at 2 GHz

struct CommDev {
    int rank, world;
    unsigned char *peer[COMM_MAXW];  // region base of every rank in THIS process' address space (peer[rank] = own)
};

struct CommHost {
    CommDev dev;
};

__device__ __forceinline__ unsigned long long *comm_flag(unsigned char *base, int slot, int src)
{
    return reinterpret_cast<unsigned long long *>(base + COMM_HDR) + slot * COMM_MAXW + src;
}
__device__ __forceinline__ unsigned char *comm_data(unsigned char *base, int slot, int src)
{
    return base + COMM_DATA_OFF + ((size_t)slot * COMM_MAXW + src) * COMM_SLOT_BYTES;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;\n" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
    return v;
}
template <typename T>
__device__ __forceinline__ T ld_volatile(const T *p)
{
    return *reinterpret_cast<const volatile T *>(p);
}

// In-place all-reduce (sum over ranks, rank order) of vals[0..n) -- shared or global memory of the calling CTA --
// scaled by `scale` on the way out.  Called by ALL threads of ONE CTA; n * sizeof(T) <= COMM_SLOT_BYTES.
template <typename T>
__device__ void peer_allreduce(const CommDev &c, T *vals, int n, T scale)
{
    const int tid = threadIdx.x, nt = blockDim.x;
    unsigned char *mine = c.peer[c.rank];
    unsigned long long *hdr = reinterpret_cast<unsigned long long *>(mine);
    const unsigned long long s = ld_volatile(hdr) + 1;  // every thread reads it before thread 0 advances it (barriers below)
    const int slot = (int)(s % COMM_SLOTS);
    for (int p = 0; p < c.world; ++p) {
        T *dst = reinterpret_cast<T *>(comm_data(c.peer[p], slot, c.rank));
        for (int k = tid; k < n; k += nt) dst[k] = vals[k] * scale;
    }
    __threadfence_system();
    __syncthreads();
    if (tid < c.world) st_release_sys(comm_flag(c.peer[tid], slot, c.rank), s);
    if (tid < c.world) {
        const unsigned long long *f = comm_flag(mine, slot, tid);
        if (ld_volatile(hdr + 1) == 0ull) {  // sticky error: a peer was lost before, do not wait again
            const long long t0 = clock64();
            while (ld_acquire_sys(f) != s) {
                if (clock64() - t0 > COMM_TIMEOUT_CYCLES) {
                    hdr[1] = s;
                    break;
                }
            }
        }
    }
    __syncthreads();
    for (int k = tid; k < n; k += nt) {
        T acc = 0;
        for (int r = 0; r < c.world; ++r) acc += ld_volatile(reinterpret_cast<const T *>(comm_data(mine, slot, r)) + k);
        vals[k] = acc;
    }
    __syncthreads();
    if (tid == 0) hdr[0] = s;
}

template <typename T>
__global__ void __launch_bounds__(1024)
allreduce_kernel(CommDev c, T *buf, int n, T scale)
{
    peer_allreduce<T>(c, buf, n, scale);
}

// BatchNorm finalize with the statistics summed over the ranks first (SyncBatchNorm forward): stats [2Co+1] fp64 =
// {sum y, sum y^2, rows} of this rank -> global sums (written back: the backward needs the global row count), then
// scale / shift / mean / invstd and the running statistics exactly as bn_finalize_kernel (csrc/train_mlp.cu).
__global__ void __launch_bounds__(256)
bn_finalize_sync_kernel(CommDev c, double *__restrict__ stats, const float *__restrict__ gamma, const float *__restrict__ beta,
                        float eps, float momentum, float *running_mean, float *running_var, long long *num_batches_tracked,
                        float *__restrict__ ss, int Co)
{
    __shared__ double st[2 * 128 + 1];
    const int n = 2 * Co + 1, tid = threadIdx.x;
    for (int k = tid; k < n; k += blockDim.x) st[k] = stats[k];
    __syncthreads();
    if (c.world > 1) peer_allreduce<double>(c, st, n, 1.0);
    for (int k = tid; k < n; k += blockDim.x) stats[k] = st[k];
    if (tid == 0 && num_batches_tracked) *num_batches_tracked += 1;
    for (int o = tid; o < Co; o += blockDim.x) {
        const double cnt = st[2 * Co];
        const double mean = st[o] / cnt;
        const double var = fmax(st[Co + o] / cnt - mean * mean, 0.0);
        const float inv = (float)(1.0 / sqrt(var + (double)eps));
        const float s = gamma[o] * inv;
        ss[o] = s;
        ss[Co + o] = fmaf(-(float)mean, s, beta[o]);
        ss[2 * Co + o] = (float)mean;
        ss[3 * Co + o] = inv;
        if (running_mean) running_mean[o] = (1.f - momentum) * running_mean[o] + momentum * (float)mean;
        if (running_var)
            running_var[o] = (1.f - momentum) * running_var[o] + momentum * (float)(cnt > 1.0 ? var * cnt / (cnt - 1.0) : var);
    }
}

// SyncBatchNorm backward: this rank's dgamma / dbeta from its OWN sums (they are summed with the other gradients
// later), then sums [2Co] = {sum dz, sum dz*y} all-reduced in place for the dx formula.
__global__ void __launch_bounds__(256)
bn_bwd_sync_kernel(CommDev c, double *__restrict__ sums, const float *__restrict__ ss, int Co, float *__restrict__ dgamma,
                   float *__restrict__ dbeta)
{
    __shared__ double st[2 * 128];
    const int n = 2 * Co, tid = threadIdx.x;
    for (int k = tid; k < n; k += blockDim.x) st[k] = sums[k];
    __syncthreads();
    for (int o = tid; o < Co; o += blockDim.x) {
        const double S1 = st[o], S2 = st[Co + o];
        dbeta[o] = (float)S1;
        dgamma[o] = (float)((double)ss[3 * Co + o] * (S2 - (double)ss[2 * Co + o] * S1));
    }
    __syncthreads();
    if (c.world > 1) peer_allreduce<double>(c, st, n, 1.0);
    for (int k = tid; k < n; k += blockDim.x) sums[k] = st[k];
}

// Gradient all-reduce + Adam in one kernel over the flat parameter / gradient buckets (sn2/optim.py::FusedAdam):
// grad <- sum over ranks of gscale_r * grad_r (rank order: every rank computes the same bits, parameters stay in
// sync without a broadcast), then torch.optim.Adam's update (L2 weight decay added to the gradient, bias
// correction, eps outside the square root) with the learning rate and the step count read from device memory, so
// a captured CUDA graph follows a scheduler without re-capture (learning/train.py:180-185: Adam + StepLR).
__global__ void __launch_bounds__(1024)
adam_sync_kernel(CommDev c, float *__restrict__ grad, float gscale, float *__restrict__ param, float *__restrict__ m,
                 float *__restrict__ v, int n, const float *__restrict__ lr_dev, float b1, float b2, float eps, float wd,
                 long long *step_dev)
{
    if (c.world > 1) peer_allreduce<float>(c, grad, n, gscale);
    __syncthreads();
    const long long t = *step_dev + 1;
    const float lr = *lr_dev;
    const double bc1 = 1.0 - pow((double)b1, (double)t), bc2 = 1.0 - pow((double)b2, (double)t);
    const float step_size = (float)((double)lr / bc1), sqrt_bc2 = (float)sqrt(bc2);
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        const float p = param[k];
        const float g = fmaf(wd, p, c.world > 1 ? grad[k] : grad[k] * gscale);
        const float mk = fmaf(1.f - b1, g - m[k], m[k]);              // lerp(m, g, 1 - b1)
        const float vk = fmaf(1.f - b2, g * g, b2 * v[k]);
        m[k] = mk;
        v[k] = vk;
        param[k] = p - step_size * (mk / (sqrtf(vk) / sqrt_bc2 + eps));
    }
    __syncthreads();
    if (threadIdx.x == 0) *step_dev = t;
}

}  // namespace sn2

using namespace sn2;

static CommDev single_rank()
{
    CommDev c;
    c.rank = 0;
    c.world = 1;
    for (int i = 0; i < COMM_MAXW; ++i) c.peer[i] = nullptr;
    return c;
}
static CommDev dev_of(void *comm) { return comm ? static_cast<CommHost *>(comm)->dev : single_rank(); }

extern "C" size_t sn2_comm_region_bytes(void) { return COMM_BYTES; }
extern "C" int sn2_comm_max_world(void) { return COMM_MAXW; }
extern "C" int sn2_comm_max_bytes(void) { return COMM_SLOT_BYTES; }

extern "C" int sn2_comm_region_alloc(void **region)
{
    if (!region) return SN2_EINVAL;
    SN2_CUDA_TRY(cudaMalloc(region, COMM_BYTES), "comm region cudaMalloc");
    SN2_CUDA_TRY(cudaMemset(*region, 0, COMM_BYTES), "comm region memset");
    SN2_CUDA_TRY(cudaDeviceSynchronize(), "comm region sync");
    return SN2_OK;
}

extern "C" int sn2_comm_region_free(void *region)
{
    if (region) SN2_CUDA_TRY(cudaFree(region), "comm region cudaFree");
    return SN2_OK;
}

extern "C" int sn2_comm_ipc_export(void *region, void *handle64_host)
{
    if (!region || !handle64_host) return SN2_EINVAL;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    SN2_CUDA_TRY(cudaIpcGetMemHandle(static_cast<cudaIpcMemHandle_t *>(handle64_host), region), "cudaIpcGetMemHandle");
    return SN2_OK;
}

extern "C" int sn2_comm_ipc_import(const void *handle64_host, void **peer_region)
{
    if (!handle64_host || !peer_region) return SN2_EINVAL;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64_host, sizeof(h));
    SN2_CUDA_TRY(cudaIpcOpenMemHandle(peer_region, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
    return SN2_OK;
}

extern "C" int sn2_comm_ipc_release(void *peer_region)
{
    if (peer_region) SN2_CUDA_TRY(cudaIpcCloseMemHandle(peer_region), "cudaIpcCloseMemHandle");
    return SN2_OK;
}

extern "C" int sn2_comm_create(int rank, int world, void *const *regions_host, void **comm_out)
{
    if (!regions_host || !comm_out || world < 1 || world > COMM_MAXW || rank < 0 || rank >= world) return SN2_EINVAL;
    CommHost *h = new CommHost;
    h->dev = single_rank();
    h->dev.rank = rank;
    h->dev.world = world;
    for (int i = 0; i < world; ++i) {
        if (!regions_host[i]) {
            delete h;
            return SN2_EINVAL;
        }
        h->dev.peer[i] = static_cast<unsigned char *>(regions_host[i]);
    }
    *comm_out = h;
    return SN2_OK;
}

extern "C" int sn2_comm_destroy(void *comm)
{
    delete static_cast<CommHost *>(comm);
    return SN2_OK;
}

// synchronous (debug / health check): number of collectives completed and the sticky error word (0 = healthy)
extern "C" int sn2_comm_status(void *comm, long long *seq_out, long long *err_out)
{
    if (!comm) return SN2_EINVAL;
    unsigned long long hdr[2];
    SN2_CUDA_TRY(cudaMemcpy(hdr, static_cast<CommHost *>(comm)->dev.peer[static_cast<CommHost *>(comm)->dev.rank], sizeof(hdr),
                            cudaMemcpyDeviceToHost),
                 "comm status");
    if (seq_out) *seq_out = (long long)hdr[0];
    if (err_out) *err_out = (long long)hdr[1];
    return SN2_OK;
}

extern "C" int sn2_comm_allreduce_f64(void *comm, double *buf, int n, void *stream)
{
    if (!comm || !buf || n <= 0 || (size_t)n * sizeof(double) > COMM_SLOT_BYTES) return SN2_EINVAL;
    allreduce_kernel<double><<<1, n >= 512 ? 1024 : 256, 0, (cudaStream_t)stream>>>(dev_of(comm), buf, n, 1.0);
    SN2_LAUNCH_CHECK("allreduce_kernel<double>");
    return SN2_OK;
}

extern "C" int sn2_comm_allreduce_f32(void *comm, float *buf, int n, float scale, void *stream)
{
    if (!comm || !buf || n <= 0 || (size_t)n * sizeof(float) > COMM_SLOT_BYTES) return SN2_EINVAL;
    allreduce_kernel<float><<<1, n >= 512 ? 1024 : 256, 0, (cudaStream_t)stream>>>(dev_of(comm), buf, n, scale);
    SN2_LAUNCH_CHECK("allreduce_kernel<float>");
    return SN2_OK;
}

extern "C" int sn2_bn_finalize_sync(void *comm, double *stats, const float *gamma, const float *beta, float eps, float momentum,
                                    float *running_mean, float *running_var, long long *num_batches_tracked, float *ss, int Co,
                                    void *stream)
{
    if (!stats || !gamma || !beta || !ss || Co <= 0 || Co > 128) return SN2_EINVAL;
    bn_finalize_sync_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(dev_of(comm), stats, gamma, beta, eps, momentum, running_mean,
                                                              running_var, num_batches_tracked, ss, Co);
    SN2_LAUNCH_CHECK("bn_finalize_sync_kernel");
    return SN2_OK;
}

extern "C" int sn2_bn_bwd_sync(void *comm, double *sums, const float *ss, int Co, float *dgamma, float *dbeta, void *stream)
{
    if (!sums || !ss || !dgamma || !dbeta || Co <= 0 || Co > 128) return SN2_EINVAL;
    bn_bwd_sync_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(dev_of(comm), sums, ss, Co, dgamma, dbeta);
    SN2_LAUNCH_CHECK("bn_bwd_sync_kernel");
    return SN2_OK;
}

extern "C" int sn2_adam_step(void *comm, float *grad, float gscale, float *param, float *m, float *v, long long n,
                             const float *lr_dev, float b1, float b2, float eps, float wd, long long *step_dev, void *stream)
{
    if (!grad || !param || !m || !v || !lr_dev || !step_dev || n <= 0) return SN2_EINVAL;
    if (n > (1 << 20) || (comm && (size_t)n * sizeof(float) > COMM_SLOT_BYTES)) return SN2_EUNSUPPORTED;
    adam_sync_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(dev_of(comm), grad, gscale, param, m, v, (int)n, lr_dev, b1, b2, eps, wd,
                                                          step_dev);
    SN2_LAUNCH_CHECK("adam_sync_kernel");
    return SN2_OK;
}
