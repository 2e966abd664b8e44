// Micro-probe of warp-collective / barrier latencies on sm_100a (used to design the FPS inner loop).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o latency_probe latency_probe.cu && ./latency_probe
#include <cstdio>
#include <cuda_runtime.h>
#define FULL 0xffffffffu

template <int OP>
__global__ void chain(unsigned *out, long long *cyc, int iters, int nthreads_sync)
{
    __shared__ unsigned sm[1024];
    unsigned v = threadIdx.x * 2654435761u + 12345u;
    sm[threadIdx.x] = v;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        if (OP == 0) v = __reduce_max_sync(FULL, v) + threadIdx.x;                 // REDUX
        if (OP == 1) v = __shfl_xor_sync(FULL, v, 1) + threadIdx.x;                // SHFL
        if (OP == 2) v = __ballot_sync(FULL, v & 1) + threadIdx.x;                 // VOTE
        if (OP == 3) { __syncthreads(); v += 1; }                                   // BAR
        if (OP == 4) { v = sm[(v + i) & 1023] + 1; }                                // LDS dependent
        if (OP == 5) { sm[threadIdx.x] = v; __syncthreads(); v = sm[(threadIdx.x + 32) & (blockDim.x - 1)] + 1; }  // STS+BAR+LDS
        if (OP == 6) { v = __float_as_uint(fminf(__uint_as_float(v), 1.5f) * 1.0001f); }  // FMNMX+FMUL dependent
        if (OP == 7) { unsigned m = __reduce_max_sync(FULL, v); unsigned mi = __reduce_min_sync(FULL, v == m ? threadIdx.x : 0xffffffffu);
                       unsigned who = __ballot_sync(FULL, v == m && threadIdx.x == mi); v = __shfl_sync(FULL, v, __ffs(who) - 1) + threadIdx.x; }  // full argmax
        if (OP == 8) { __syncwarp(); v += 1; }
        if (OP == 9) { sm[threadIdx.x] = v; __syncwarp(); v = sm[threadIdx.x ^ 1] + 1; }  // STS+syncwarp+LDS
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = v;
}

int main()
{
    unsigned *out; long long *cyc;
    cudaMalloc(&out, 1024 * 4 * 4); cudaMalloc(&cyc, 64);
    const char *names[] = {"REDUX.max", "SHFL", "VOTE.ballot", "BAR.SYNC", "LDS dep", "STS+BAR+LDS", "FMNMX+FMUL", "argmax(2 REDUX+ballot+shfl)", "syncwarp", "STS+syncwarp+LDS"};
    int iters = 2000;
    for (int threads : {32, 256, 1024}) {
        printf("--- %d threads/CTA, 1 CTA ---\n", threads);
#define RUN(OP) { chain<OP><<<1, threads>>>(out, cyc, iters, threads); chain<OP><<<1, threads>>>(out, cyc, iters, threads); long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost); printf("%-32s %7.1f cycles/iter\n", names[OP], (double)c / iters); }
        RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9)
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(e));
    return 0;
}
