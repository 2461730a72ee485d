// Shared device helpers and kernel-parameter structs for the WiFlow sm_100a kernels.
//
// Internal activation layout (DESIGN.md "Data layout in HBM"): every intermediate is
//   [channel][position][n]   with n = b*20 + t   (b = CSI window, t = time step),
// n contiguous.  TCN tensors have one position, the conv stack uses position = feature-axis
// index w, the attention/decoder tensors use position = keypoint slot h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define WF_T 20            // time steps per CSI window (models/pose_model.py:72)
#define WF_MAX_TAPS 9

// prologue modes: how an operand element is produced from what is stored in HBM
enum { PRO_NONE = 0,       // x
       PRO_BNSILU = 1,     // mask * silu(a[c]*x + b[c])           (BatchNorm + SiLU (+Dropout))
       PRO_AFFINE = 2,     // a[c]*x + b[c]                         (BatchNorm only)
       PRO_BNBWD = 3 };    // a[c]*dy + b[c]*raw + c[c]             (BatchNorm backward, two tensors)
// epilogue modes of the conv GEMM
enum { EPI_STORE = 0,      // out = acc + bias
       EPI_STATS = 1,      // ... and accumulate sum / sum-of-squares per output channel
       EPI_DSILU = 2,      // dy = acc * mask * silu'(s*raw+t); stats: sum dy, sum dy*raw
       EPI_DAFF = 3 };     // dy = acc;                          stats: sum dy, sum dy*raw

struct ConvP {
    // B operand ("input" of the conv)
    const float* in;
    const float* in2;                 // PRO_BNBWD: raw tensor, same addressing
    long long in_sc, in_sp, in_sb;    // element (c,p,b,t) at c*in_sc + p*in_sp + b*in_sb + t
    int pro_mode;
    const float *pro_a, *pro_b, *pro_c;
    const float* mask;                // multiplicative dropout mask or nullptr
    long long m_sb, m_sc; int m_st;   // mask[b*m_sb + c*m_sc + t*m_st]
    // A operand (packed weights [groups][ntaps][Kpad][Mpad], m contiguous, zero padded)
    const float* w;
    int Kpad, Mpad;
    // geometry
    int Cin, Cout, groups, Pin, Pout, N, ntaps;
    int pmul, pdiv;                   // ipos = (opos*pmul + dp[tap]) / pdiv, valid iff divisible and in range
    int dp[WF_MAX_TAPS], dn[WF_MAX_TAPS];
    // output
    float* out;
    long long out_sc, out_sp, out_sb;
    const float* bias;
    int epi_mode, accumulate;
    const float* eraw;                // EPI_DSILU/DAFF: raw tensor at the output location (addressing of out)
    const float *e_scale, *e_shift;
    const float* emask; long long em_sb, em_sc; int em_st;
    double *stat0, *stat1;            // per output channel
};

struct WgradP {
    const float* g;  const float* g2;          // dy / raw of the conv output, [groups*Cout][Pout][N]
    int g_pro;  const float *g_a, *g_b, *g_c;  // PRO_NONE or PRO_BNBWD
    const float* in; const float* in2;
    long long in_sc, in_sp, in_sb;
    int pro_mode; const float *pro_a, *pro_b, *pro_c;
    const float* mask; long long m_sb, m_sc; int m_st;
    int Cin, Cout, groups, Pin, Pout, N, ntaps, pmul;
    int dp[WF_MAX_TAPS], dn[WF_MAX_TAPS];
    float* dw;                                  // reference layout [groups*Cout][Cin][ntaps], atomically accumulated
    int kchunks;                                // split of the (p, n) reduction across blockIdx.x
};

__device__ __forceinline__ float wf_sigmoid(float x) { return __fdividef(1.f, 1.f + expf(-x)); }
__device__ __forceinline__ float wf_silu(float x) { return x * wf_sigmoid(x); }
__device__ __forceinline__ float wf_dsilu(float x) { float s = wf_sigmoid(x); return s * (1.f + x * (1.f - s)); }

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 f4zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide sum of two floats, result accumulated into two global doubles (one atomic pair per block)
template <int NT>
__device__ __forceinline__ void block_accum2(float a, float b, double* d0, double* d1) {
    __shared__ double red[2][NT / 32];
    double da = warp_sum_d((double)a), db = warp_sum_d((double)b);
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { red[0][w] = da; red[1][w] = db; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s0 = 0, s1 = 0;
#pragma unroll
        for (int i = 0; i < NT / 32; ++i) { s0 += red[0][i]; s1 += red[1][i]; }
        atomicAdd(d0, s0);
        atomicAdd(d1, s1);
    }
    __syncthreads();
}
