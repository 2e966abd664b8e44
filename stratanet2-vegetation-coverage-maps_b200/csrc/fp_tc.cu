// FP1 + head on the tensor cores (tcgen05): the per-point MLP[42,34] of the last feature-propagation level as
// 128-point tiles.  SURVEY.md §8a a8 + a10 (reference model/point_net2.py:62-67, 141-151).
//
// The FP levels are the GEMM-shaped part of the network: every point is one dense row, no raggedness, so a CTA of
// 128 threads (one point per thread, one TMEM lane per point) builds a [128 x 48] operand tile (34 interpolated
// channels ++ 8 raw features, zero padded to K = 48), one thread issues tcgen05.mma.kind::tf32 against the
// [48 x 48] weight tile (N = 34 padded to 48) into 48 TMEM columns, and every thread reads its own accumulator row
// back with tcgen05.ld for the epilogue (bias, ReLU, BN, lin1 + ReLU, lin2, softmax / sigmoid) in registers.
// Operands are split tf32 hi + lo (3xTF32: A_hi B_hi + A_lo B_hi + A_hi B_lo), which keeps fp32 accuracy, so this
// kernel obeys the same 1e-3 parity bound as the SIMT one.  Tensor time per tile: 18 MMAs x 24 cycles = 432 cycles
// for 128 x 1428 FMAs that cost the FP32 pipes ~1430 cycles.
#include "mlp_common.cuh"

namespace sn2 {

__device__ __forceinline__ unsigned tc_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

// K-major, no-swizzle canonical layout: core matrix = 8 rows x 16 bytes; element (r, k) of a [rows x KP] fp32
// operand at (r/8)*SBO + (k/4)*128 + (r%8)*16 + (k%4)*4 with SBO = (KP/4)*128 bytes; LBO = 128 bytes.
template <int KP>
__device__ __forceinline__ unsigned long long tc_desc(unsigned saddr)
{
    constexpr unsigned SBO = (KP / 4) * 128;
    return (unsigned long long)((saddr >> 4) & 0x3fff) | ((unsigned long long)(128 >> 4) << 16) |
           ((unsigned long long)(SBO >> 4) << 32) | (1ull << 46);
}
template <int NP>
__device__ __forceinline__ void tc_mma_tf32(unsigned tmem_d, unsigned long long da, unsigned long long db, unsigned acc)
{
    // kind::tf32, D = F32, A / B = TF32 K-major, N = NP, M = 128
    constexpr unsigned IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(NP >> 3) << 17) | ((128u >> 4) << 24);
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}\n" ::"r"(tmem_d), "l"(da),
                 "l"(db), "r"(IDESC), "r"(acc), "r"(0), "r"(0), "r"(0), "r"(0));
}
__device__ __forceinline__ void tc_ld16(unsigned taddr, unsigned (&d)[16])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                 : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7]), "=r"(d[8]),
                   "=r"(d[9]), "=r"(d[10]), "=r"(d[11]), "=r"(d[12]), "=r"(d[13]), "=r"(d[14]), "=r"(d[15])
                 : "r"(taddr));
}

constexpr int FT_KP = 48, FT_NP = 48, FT_TCOLS = 64;
constexpr int FT_A_BYTES = 128 * FT_KP * 4, FT_B_BYTES = FT_NP * FT_KP * 4;
constexpr int FT_SMEM = 2 * FT_A_BYTES + 2 * FT_B_BYTES + 64;

__device__ __forceinline__ float interp_tc(float a0, float a1, float a2, float w0, float w1, float w2, float den)
{
    return __fdiv_rn(__fadd_rn(__fadd_rn(__fmul_rn(a0, w0), __fmul_rn(a1, w1)), __fmul_rn(a2, w2)), den);
}

__global__ void __launch_bounds__(128, 3)
fp1_head_tc_kernel(const float *__restrict__ f2, const int *__restrict__ nbr, const float *__restrict__ wgt,
                   const float *__restrict__ feat, int Q, int ntiles, const __grid_constant__ W_FP1 W,
                   float4 *__restrict__ cov, float4 *__restrict__ proba)
{
    extern __shared__ __align__(1024) unsigned char ft_smem[];
    unsigned char *a_hi = ft_smem, *a_lo = ft_smem + FT_A_BYTES;
    float *b_hi = reinterpret_cast<float *>(ft_smem + 2 * FT_A_BYTES);
    float *b_lo = reinterpret_cast<float *>(ft_smem + 2 * FT_A_BYTES + FT_B_BYTES);
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(ft_smem + 2 * FT_A_BYTES + 2 * FT_B_BYTES);
    __shared__ unsigned s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5;
    constexpr int SBO_F = (FT_KP / 4) * 32;  // SBO in floats

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(tc_smem_u32(&s_tmem)),
                     "r"(FT_TCOLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    if (tid == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(tc_smem_u32(bar)));
    // B operand [N = 48 (o) x K = 48 (k)]: B(o, k) = l1.w[k][o] inside 34 x 42, zero elsewhere; tf32 hi / lo split
    for (int e = tid; e < FT_NP * FT_KP; e += 128) {
        const int o = e / FT_KP, k = e - o * FT_KP;
        const float wv = (o < SN2_CF && k < SN2_CF + SN2_F0) ? W.l1.w[k][o] : 0.f;
        const float hi = __uint_as_float(__float_as_uint(wv) & 0xffffe000u);
        const int off = (o >> 3) * SBO_F + (k >> 2) * 32 + (o & 7) * 4 + (k & 3);
        b_hi[off] = hi;
        b_lo[off] = wv - hi;
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n");
    const unsigned tmem = s_tmem;
    const unsigned tmem_row = tmem + ((unsigned)(32 * warp) << 16);  // this thread's accumulator row = TMEM lane tid
    const unsigned long long d_ahi = tc_desc<FT_KP>(tc_smem_u32(a_hi)), d_alo = tc_desc<FT_KP>(tc_smem_u32(a_lo));
    const unsigned long long d_bhi = tc_desc<FT_KP>(tc_smem_u32(b_hi)), d_blo = tc_desc<FT_KP>(tc_smem_u32(b_lo));
    unsigned parity = 0;
    const int roff = (tid >> 3) * (SBO_F * 4) + (tid & 7) * 16;  // byte offset of this thread's operand row

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long row = (long long)tile * 128 + tid;
        const bool on = row < Q;
        // ---- operand row: 34 interpolated channels ++ 8 raw features ++ 6 zeros, written as tf32 hi / lo ----
        {
            int i0 = 0, i1 = 0, i2 = 0;
            float w0 = 0.f, w1 = 0.f, w2 = 0.f, den = 1.f;
            if (on) {
                i0 = __ldg(nbr + 3 * row); i1 = __ldg(nbr + 3 * row + 1); i2 = __ldg(nbr + 3 * row + 2);
                w0 = __ldg(wgt + 3 * row); w1 = __ldg(wgt + 3 * row + 1); w2 = __ldg(wgt + 3 * row + 2);
                den = __fadd_rn(__fadd_rn(w0, w1), w2);
            }
            const float *a0 = f2 + (size_t)i0 * SN2_CF_LD, *a1 = f2 + (size_t)i1 * SN2_CF_LD, *a2 = f2 + (size_t)i2 * SN2_CF_LD;
#pragma unroll
            for (int kc = 0; kc < FT_KP / 4; ++kc) {
                float v[4] = {0.f, 0.f, 0.f, 0.f};
                if (on) {
                    if (kc < 9) {  // interpolated f2 channels 4*kc .. 4*kc+3 (row stride 36: channels 34, 35 are padding)
                        const float4 p0 = ldg4(a0 + 4 * kc), p1 = ldg4(a1 + 4 * kc), p2 = ldg4(a2 + 4 * kc);
                        v[0] = interp_tc(p0.x, p1.x, p2.x, w0, w1, w2, den);
                        v[1] = interp_tc(p0.y, p1.y, p2.y, w0, w1, w2, den);
                        if (kc < 8) {
                            v[2] = interp_tc(p0.z, p1.z, p2.z, w0, w1, w2, den);
                            v[3] = interp_tc(p0.w, p1.w, p2.w, w0, w1, w2, den);
                        }
                    }
                    // raw features occupy k = 34 .. 41
                    if (kc == 8) { const float4 f = ldg4(feat + (size_t)row * SN2_F0); v[2] = f.x; v[3] = f.y; }
                    if (kc == 9) { const float4 f = ldg4(feat + (size_t)row * SN2_F0), g = ldg4(feat + (size_t)row * SN2_F0 + 4);
                                   v[0] = f.z; v[1] = f.w; v[2] = g.x; v[3] = g.y; }
                    if (kc == 10) { const float4 g = ldg4(feat + (size_t)row * SN2_F0 + 4); v[0] = g.z; v[1] = g.w; }
                }
                float hi[4], lo[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    hi[t] = __uint_as_float(__float_as_uint(v[t]) & 0xffffe000u);
                    lo[t] = v[t] - hi[t];
                }
                *reinterpret_cast<float4 *>(a_hi + roff + kc * 128) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<float4 *>(a_lo + roff + kc * 128) = make_float4(lo[0], lo[1], lo[2], lo[3]);
            }
        }
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;\n");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;\n");
#pragma unroll
            for (int ks = 0; ks < FT_KP / 8; ++ks) {  // K = 48 = six K=8 steps: +256 bytes (+16 in the address field)
                const unsigned long long o = (unsigned long long)(ks * 16);
                tc_mma_tf32<FT_NP>(tmem, d_ahi + o, d_bhi + o, ks > 0 ? 1u : 0u);
                tc_mma_tf32<FT_NP>(tmem, d_alo + o, d_bhi + o, 1u);
                tc_mma_tf32<FT_NP>(tmem, d_ahi + o, d_blo + o, 1u);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(tc_smem_u32(bar))
                         : "memory");
        }
        unsigned done = 0;
        for (int spin = 0; spin < (1 << 24) && !done; ++spin)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                         : "=r"(done) : "r"(tc_smem_u32(bar)), "r"(parity) : "memory");
        if (!done) __trap();  // the MMA never completed: fail loudly instead of hanging the GPU
        parity ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;\n");
        unsigned d0[16], d1[16], d2[16];
        tc_ld16(tmem_row, d0);
        tc_ld16(tmem_row + 16, d1);
        tc_ld16(tmem_row + 32, d2);
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;\n");
        // ---- epilogue in registers: bias, ReLU, BN, lin1 + ReLU, lin2, softmax / sigmoid ----------------------
        float h[SN2_CF];
#pragma unroll
        for (int o = 0; o < 16; ++o) h[o] = __uint_as_float(d0[o]);
#pragma unroll
        for (int o = 0; o < 16; ++o) h[16 + o] = __uint_as_float(d1[o]);
        h[32] = __uint_as_float(d2[0]);
        h[33] = __uint_as_float(d2[1]);
#pragma unroll
        for (int o = 0; o < SN2_CF; ++o) h[o] = fmaf(fmaxf(h[o] + W.l1.b[o], 0.f), W.l1.s[o], W.l1.t[o]);
        float u[16];
        acc_init(W.lin1, u);
#pragma unroll
        for (int k = 0; k < SN2_CF; ++k) acc_step<0>(W.lin1, h[k], u, k);
#pragma unroll
        for (int o = 0; o < 16; ++o) u[o] = fmaxf(u[o], 0.f);
        float sc[5];
        acc_init(W.lin2, sc);
#pragma unroll
        for (int k = 0; k < 16; ++k) acc_step<0>(W.lin2, u[k], sc, k);
        const float m = fmaxf(fmaxf(sc[0], sc[1]), fmaxf(sc[2], sc[3]));
        const float e0 = expf(sc[0] - m), e1 = expf(sc[1] - m), e2 = expf(sc[2] - m), e3 = expf(sc[3] - m);
        const float sum = (e0 + e1) + (e2 + e3);
        const float4 pr = make_float4(e0 / sum, e1 / sum, e2 / sum, e3 / sum);
        const float dens = 1.0f / (1.0f + expf(-sc[4]));
        if (on) {
            proba[row] = pr;
            cov[row] = make_float4(pr.x * dens, pr.y * dens, pr.z * dens, pr.w * dens);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(FT_TCOLS));
}

}  // namespace sn2

extern "C" int sn2_fp1_head_fwd_tc(const float *f2, const int *nbr, const float *w, const float *feat, int Q,
                                   const float *w_host, int nw, float *cov, float *proba, void *stream)
{
    using namespace sn2;
    if (!f2 || !nbr || !w || !feat || !cov || !proba || Q <= 0) return SN2_EINVAL;
    W_FP1 ws;
    if (int rc = load_weights(ws, w_host, nw)) return rc;
    const int ntiles = (Q + 127) / 128;
    const int grid = ntiles < 148 * 3 ? ntiles : 148 * 3;
    auto kern = fp1_head_tc_kernel;
    SN2_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, FT_SMEM), "fp1_head_tc attr");
    kern<<<grid, 128, FT_SMEM, (cudaStream_t)stream>>>(f2, nbr, w, feat, Q, ntiles, ws, reinterpret_cast<float4 *>(cov),
                                                      reinterpret_cast<float4 *>(proba));
    SN2_LAUNCH_CHECK("fp1_head_tc_kernel");
    return SN2_OK;
}
