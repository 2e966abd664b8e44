// Weight structs and per-thread MLP building blocks shared by the forward kernels.
#pragma once
#include "sn2_common.cuh"
#include <string.h>

namespace sn2 {

template <int CIN, int COUT>
struct Layer {          // flat float layout: w (k-major), b, s, t
    float w[CIN][COUT];  // transposed Linear weight: w[k][o] = W[o][k]
    float b[COUT];
    float s[COUT];       // eval BN scale
    float t[COUT];       // eval BN shift
};
template <int CIN, int COUT>
struct Lin {
    float w[CIN][COUT];
    float b[COUT];
};

struct W_SA1 { Layer<SN2_F0 + 3, SN2_C1> l1; Layer<SN2_C1, SN2_C1> l2; };
struct W_SA2 { Layer<SN2_C1 + 3, SN2_C2> l1; };
struct W_SA3 { Layer<SN2_C2 + 3, SN2_C3> l1; };
struct W_FP3 { Layer<SN2_C3 + SN2_C2, SN2_C3> l1; };
struct W_FP2 { Layer<SN2_C3 + SN2_C1, SN2_CF> l1; };
struct W_FP1 { Layer<SN2_CF + SN2_F0, SN2_CF> l1; Lin<SN2_CF, 16> lin1; Lin<16, 5> lin2; };

template <int COUT, typename L>
__device__ __forceinline__ void acc_init(const L &l, float (&acc)[COUT])
{
#pragma unroll
    for (int o = 0; o < COUT; ++o) acc[o] = l.b[o];
}
// acc += in * w[K0 + k][:]
template <int K0, int COUT, typename L>
__device__ __forceinline__ void acc_step(const L &l, float in, float (&acc)[COUT], int k)
{
#pragma unroll
    for (int o = 0; o < COUT; ++o) acc[o] = fmaf(in, l.w[K0 + k][o], acc[o]);
}
template <int COUT, typename L>
__device__ __forceinline__ void relu_bn(const L &l, float (&acc)[COUT])
{
#pragma unroll
    for (int o = 0; o < COUT; ++o) acc[o] = fmaf(fmaxf(acc[o], 0.f), l.s[o], l.t[o]);
}


// Per-edge message MLP of a set-abstraction level: in = [x_j (CIN), pos_j - pos_i (3)] -> COUT channels.
// features first, then relative position; every layer Linear -> ReLU -> BN(eval affine).
template <int LEVEL> struct SAEdge;
template <> struct SAEdge<1> {
    using W = W_SA1;
    static constexpr int CIN = SN2_F0, COUT = SN2_C1;
    __device__ static __forceinline__ void run(const W &w, const float *__restrict__ frow, float rx, float ry, float rz,
                                               float (&out)[COUT])
    {
        const float4 f0 = ldg4(frow), f1 = ldg4(frow + 4);
        const float in[CIN + 3] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w, rx, ry, rz};
        float h1[SN2_C1];
        acc_init(w.l1, h1);
#pragma unroll
        for (int k = 0; k < CIN + 3; ++k) acc_step<0>(w.l1, in[k], h1, k);
        relu_bn(w.l1, h1);
        acc_init(w.l2, out);
#pragma unroll
        for (int k = 0; k < SN2_C1; ++k) acc_step<0>(w.l2, h1[k], out, k);
        relu_bn(w.l2, out);
    }
};
template <> struct SAEdge<2> {
    using W = W_SA2;
    static constexpr int CIN = SN2_C1, COUT = SN2_C2;
    __device__ static __forceinline__ void run(const W &w, const float *__restrict__ frow, float rx, float ry, float rz,
                                               float (&out)[COUT])
    {
        acc_init(w.l1, out);
#pragma unroll
        for (int v = 0; v < CIN / 4; ++v) {
            const float4 f = ldg4(frow + 4 * v);
            acc_step<0>(w.l1, f.x, out, 4 * v);
            acc_step<0>(w.l1, f.y, out, 4 * v + 1);
            acc_step<0>(w.l1, f.z, out, 4 * v + 2);
            acc_step<0>(w.l1, f.w, out, 4 * v + 3);
        }
        acc_step<CIN>(w.l1, rx, out, 0);
        acc_step<CIN>(w.l1, ry, out, 1);
        acc_step<CIN>(w.l1, rz, out, 2);
        relu_bn(w.l1, out);
    }
};

template <typename WS>
static inline int load_weights(WS &w, const float *w_host, int nw)
{
    if (!w_host || (size_t)nw * sizeof(float) != sizeof(WS)) return SN2_EINVAL;
    memcpy(&w, w_host, sizeof(WS));
    return SN2_OK;
}

}  // namespace sn2
