// Probe: TMEM (tcgen05.ld/st 32x32b) as dynamically indexed per-warp scratch on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_probe tmem_probe.cu && timeout 60 ./tmem_probe
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void tmem_ld2(unsigned addr, unsigned &a, unsigned &b)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];\n" : "=r"(a), "=r"(b) : "r"(addr));
}
__device__ __forceinline__ void tmem_st2(unsigned addr, unsigned a, unsigned b)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};\n" ::"r"(addr), "r"(a), "r"(b));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }

constexpr int NW = 8, COLS = 256, PER_WARP = 128;

__global__ void __launch_bounds__(NW * 32, 1) probe(unsigned *out, long long *cyc, int iters)
{
    __shared__ unsigned tbase_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        unsigned dst = (unsigned)__cvta_generic_to_shared(&tbase_s);
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(dst), "r"(COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n");
    const unsigned tbase = tbase_s;
    const unsigned wbase = tbase + ((unsigned)(32 * (warp & 3)) << 16) + (unsigned)((warp >> 2) * PER_WARP);
    // fill: slot k (0..63) holds (warp*1000000 + k*1000 + lane, ~that)
    for (int k = 0; k < PER_WARP / 2; ++k) {
        unsigned v = warp * 1000000u + k * 1000u + lane;
        tmem_st2(wbase + 2 * k, v, ~v);
    }
    tmem_wait_st();
    // dynamic read-back in a data-dependent order + verify
    unsigned errs = 0, k = (warp * 7 + 3) & 63;
    for (int i = 0; i < 64; ++i) {
        unsigned a, b;
        tmem_ld2(wbase + 2 * k, a, b);
        tmem_wait_ld();
        unsigned v = warp * 1000000u + k * 1000u + lane;
        if (a != v || b != ~v) ++errs;
        k = (k * 5 + 1) & 63;  // full-period LCG over 64 slots
    }
    // latency of a dependent ld -> modify -> st -> ld chain on one slot
    long long t0 = clock64();
    unsigned acc = 0;
    for (int i = 0; i < iters; ++i) {
        unsigned a, b;
        tmem_ld2(wbase + 2 * ((acc + i) & 63), a, b);
        tmem_wait_ld();
        acc += a & 1;
        tmem_st2(wbase + 2 * ((acc + i) & 63), a + 2, b);
        tmem_wait_st();
    }
    long long t1 = clock64();
    // ld-only dependent chain
    for (int i = 0; i < iters; ++i) {
        unsigned a, b;
        tmem_ld2(wbase + 2 * ((acc + i) & 63), a, b);
        tmem_wait_ld();
        acc += a & 1;
    }
    long long t2 = clock64();
    out[threadIdx.x] = errs;
    out[256 + threadIdx.x] = acc;
    if (lane == 0) { cyc[warp * 2] = t1 - t0; cyc[warp * 2 + 1] = t2 - t1; }
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tbase), "r"(COLS));
}

int main()
{
    unsigned *out; long long *cyc;
    cudaMalloc(&out, 512 * 4); cudaMalloc(&cyc, 16 * 8);
    int iters = 1000;
    probe<<<4, NW * 32>>>(out, cyc, iters);
    cudaError_t e = cudaDeviceSynchronize();
    printf("sync: %s\n", cudaGetErrorString(e));
    unsigned h[512]; long long c[16];
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost); cudaMemcpy(c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
    unsigned errs = 0; for (int i = 0; i < 256; ++i) errs += h[i];
    printf("readback errors: %u\n", errs);
    for (int w = 0; w < NW; ++w) printf("warp %d: ld+st chain %.1f cyc/iter, ld chain %.1f cyc/iter\n", w, (double)c[2*w]/iters, (double)c[2*w+1]/iters);
    return 0;
}
