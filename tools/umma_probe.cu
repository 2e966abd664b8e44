// Probe: one 128x16x16 TF32 tcgen05.mma (A, B from shared memory, K-major, no swizzle; D in TMEM), checked
// against a CPU product.  Validates the descriptor encodings used by csrc/sa_fused_tc.cu.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_probe umma_probe.cu && timeout 60 ./umma_probe
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

// K-major, SWIZZLE_NONE canonical layout: core matrix = 8 rows x 16 bytes (128 contiguous bytes);
// LBO = byte distance between core matrices adjacent in K, SBO = between core matrices adjacent in M/N.
__device__ __forceinline__ unsigned long long make_desc(unsigned saddr, unsigned lbo_bytes, unsigned sbo_bytes)
{
    unsigned long long d = 0;
    d |= (unsigned long long)((saddr >> 4) & 0x3fff);
    d |= (unsigned long long)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (unsigned long long)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= 1ull << 46;  // descriptor version 1 (Blackwell)
    return d;         // base_offset 0, lbo_mode 0, layout_type 0 = SWIZZLE_NONE
}

constexpr unsigned IDESC_TF32_M128_N16 =
    (1u << 4) |   // D format F32
    (2u << 7) |   // A format TF32
    (2u << 10) |  // B format TF32
    (0u << 15) | (0u << 16) |  // A, B K-major
    ((16u >> 3) << 17) |       // N = 16
    ((128u >> 4) << 24);       // M = 128

__global__ void __launch_bounds__(128, 1) probe(const float *A, const float *W, float *D, long long *cyc, int *err)
{
    __shared__ __align__(128) float sA[128 * 16];  // 16 row groups x (4 k-chunks x 128 B)
    __shared__ __align__(128) float sB[16 * 16];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ unsigned tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)), "r"(32));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;\n");
    }
    // row r = tid: element (r, k) -> (r/8)*512 + (k/4)*128 + (r%8)*16 + (k%4)*4 bytes
    {
        const int r = tid;
        char *base = reinterpret_cast<char *>(sA) + (r >> 3) * 512 + (r & 7) * 16;
        for (int kc = 0; kc < 4; ++kc)
            *reinterpret_cast<float4 *>(base + kc * 128) = *reinterpret_cast<const float4 *>(A + r * 16 + kc * 4);
        if (r < 16) {
            char *bb = reinterpret_cast<char *>(sB) + (r >> 3) * 512 + (r & 7) * 16;
            for (int kc = 0; kc < 4; ++kc)
                *reinterpret_cast<float4 *>(bb + kc * 128) = *reinterpret_cast<const float4 *>(W + r * 16 + kc * 4);
        }
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n");
    const unsigned tmem = tmem_base_s;
    long long t0 = clock64();
    if (tid == 0) {
        const unsigned long long da = make_desc(smem_u32(sA), 128, 512), db = make_desc(smem_u32(sB), 128, 512);
        for (int k = 0; k < 2; ++k) {  // K = 16 = 2 x (K = 8 per tf32 MMA): advance both operands by 2 core matrices
            const unsigned long long a = da + (unsigned long long)((k * 256) >> 4), b = db + (unsigned long long)((k * 256) >> 4);
            const unsigned acc = k > 0;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}\n" ::"r"(tmem),
                         "l"(a), "l"(b), "r"(IDESC_TF32_M128_N16), "r"(acc), "r"(0), "r"(0), "r"(0), "r"(0));
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&bar)) : "memory");
    }
    // bounded wait on phase 0
    unsigned ok = 0;
    for (int spin = 0; spin < 2000000 && !ok; ++spin) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0));
    }
    long long t1 = clock64();
    if (!ok) { if (tid == 0) *err = 1; }
    asm volatile("tcgen05.fence::after_thread_sync;\n");
    if (ok) {
        unsigned v[16];
        const unsigned taddr = tmem + ((unsigned)(32 * warp) << 16);
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                       "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
        for (int j = 0; j < 16; ++j) D[tid * 16 + j] = __uint_as_float(v[j]);
    }
    if (tid == 0) cyc[0] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(32));
}

int main()
{
    float hA[128 * 16], hW[16 * 16], hD[128 * 16];
    srand(1);
    for (float &v : hA) v = (rand() % 2001 - 1000) / 500.0f;
    for (float &v : hW) v = (rand() % 2001 - 1000) / 1000.0f;
    float *dA, *dW, *dD; long long *cyc; int *err;
    cudaMalloc(&dA, sizeof(hA)); cudaMalloc(&dW, sizeof(hW)); cudaMalloc(&dD, sizeof(hD)); cudaMalloc(&cyc, 8); cudaMalloc(&err, 4);
    cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice); cudaMemcpy(dW, hW, sizeof(hW), cudaMemcpyHostToDevice);
    cudaMemset(dD, 0, sizeof(hD)); cudaMemset(err, 0, 4);
    probe<<<1, 128>>>(dA, dW, dD, cyc, err);
    cudaError_t e = cudaDeviceSynchronize();
    printf("sync: %s\n", cudaGetErrorString(e));
    int herr; long long hc;
    cudaMemcpy(&herr, err, 4, cudaMemcpyDeviceToHost); cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(hD, dD, sizeof(hD), cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0;
    for (int r = 0; r < 128; ++r)
        for (int o = 0; o < 16; ++o) {
            double s = 0;
            for (int k = 0; k < 16; ++k) s += (double)hA[r * 16 + k] * hW[o * 16 + k];  // D = A * W^T
            maxerr = fmax(maxerr, fabs(s - hD[r * 16 + o]));
            maxref = fmax(maxref, fabs(s));
        }
    printf("timeout flag %d, mma issue->complete %lld cycles, max |err| %.3e (max |ref| %.3f)\n", herr, hc, maxerr, maxref);
    printf("D[0][0..3] = %f %f %f %f ; D[127][15] = %f\n", hD[0], hD[1], hD[2], hD[3], hD[127 * 16 + 15]);
    return 0;
}
