// FP32 FMA issue-rate probe on sm_100a: FFMA (register / constant-bank operand) against the packed FFMA2
// (fma.rn.f32x2) that Blackwell adds.  Decides the inner-loop form of the per-edge / per-point MLP kernels.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma_probe ffma_probe.cu && ./ffma_probe
#include <cstdio>
#include <cuda_runtime.h>

struct Wt { float w[16][16]; };
typedef unsigned long long u64;

__device__ __forceinline__ u64 pack2(float a, float b)
{
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c)
{
    u64 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ float lo(u64 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a + b; }

// MODE 0: FFMA, weights in registers.  1: FFMA, weights from the constant bank (kernel parameter).
// 2: FFMA2, weight pairs in registers.  3: FFMA2, weight pairs from the constant bank.
template <int MODE>
__global__ void __launch_bounds__(256) probe(float *out, long long *cyc, int iters, const __grid_constant__ Wt W)
{
    float in[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) in[k] = threadIdx.x * 0.001f + k;
    float acc[16];
    u64 acc2[8];
#pragma unroll
    for (int o = 0; o < 16; ++o) acc[o] = 0.f;
#pragma unroll
    for (int o = 0; o < 8; ++o) acc2[o] = 0ull;
    float wr[16];
    u64 wr2[8];
#pragma unroll
    for (int o = 0; o < 16; ++o) wr[o] = out[o];
#pragma unroll
    for (int o = 0; o < 8; ++o) wr2[o] = pack2(out[2 * o], out[2 * o + 1]);
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            if (MODE == 0) {
#pragma unroll
                for (int o = 0; o < 16; ++o) acc[o] = fmaf(in[k], wr[o], acc[o]);
            } else if (MODE == 1) {
#pragma unroll
                for (int o = 0; o < 16; ++o) acc[o] = fmaf(in[k], W.w[k][o], acc[o]);
            } else if (MODE == 2) {
                const u64 i2 = pack2(in[k], in[k]);
#pragma unroll
                for (int o = 0; o < 8; ++o) acc2[o] = fma2(i2, wr2[o], acc2[o]);
            } else {
                const u64 i2 = pack2(in[k], in[k]);
#pragma unroll
                for (int o = 0; o < 8; ++o) acc2[o] = fma2(i2, pack2(W.w[k][2 * o], W.w[k][2 * o + 1]), acc2[o]);
            }
        }
#pragma unroll
        for (int k = 0; k < 16; ++k) in[k] += 1e-7f;
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int o = 0; o < 16; ++o) s += acc[o];
#pragma unroll
    for (int o = 0; o < 8; ++o) s += lo(acc2[o]);
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[64 + blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main()
{
    float *out; long long *cyc;
    cudaMalloc(&out, (64 + 148 * 4 * 256) * 4); cudaMalloc(&cyc, 148 * 4 * 8);
    cudaMemset(out, 0, (64 + 148 * 4 * 256) * 4);
    Wt w;
    for (int k = 0; k < 16; ++k) for (int o = 0; o < 16; ++o) w.w[k][o] = 0.001f * (k + o);
    const char *names[] = {"FFMA  reg weights", "FFMA  const-bank weights", "FFMA2 reg weights", "FFMA2 const-bank weights"};
    const int iters = 4000;
    for (int ctas : {1, 2, 4}) {
        printf("--- %d CTA(s) of 256 threads per SM ---\n", ctas);
#define RUN(M) { probe<M><<<148 * ctas, 256>>>(out, cyc, iters, w); probe<M><<<148 * ctas, 256>>>(out, cyc, iters, w); long long c; \
                 cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost); \
                 printf("%-26s %8.1f cycles/iter  -> %6.1f FMA/clk/SM\n", names[M], (double)c / iters, 256.0 * 256 * ctas / ((double)c / iters)); }
        RUN(0) RUN(1) RUN(2) RUN(3)
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
