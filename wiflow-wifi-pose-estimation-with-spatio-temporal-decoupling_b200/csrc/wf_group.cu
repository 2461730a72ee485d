// TCN grouped causal dilated convolutions (models/tcn.py:20-22,33-35 + Chomp1d tcn.py:11-12), forward and backward-data:
// 20 groups of 12..27 channels, 3 taps that are COLUMN shifts (t - (2-k)*d, zero outside the 20-step window).
// These layers carry 6% of the FLOPs but were 10x above their HBM time as a generic implicit GEMM, because every tap re-staged
// (loaded + normalised) the same activation tile.  Here a CTA stages the group's [channels][256 columns + 16-column halos] tile
// ONCE (BatchNorm+SiLU+Dropout / BatchNorm-backward applied on load, values pre-split into tf32 hi/lo halves), keeps the
// group's 3 x K x M weights in shared memory, and runs the three taps as shifted fragment reads out of that tile:
//   D[m][c] = sum_tap sum_k W[tap][k][m] * X[k][c + dn(tap)]   (skipped where t(c) + dn leaves the window)
// on the warp-level tensor-core path (mma.sync m16n8k8 tf32, fp32 accumulate, 3xTF32 split: a_lo*b_hi + a_hi*b_lo in a
// correction accumulator, a_hi*b_hi in the main one).  No barrier inside the K loop.  The epilogue pairs lanes so every
// thread owns 4 consecutive columns of one output row and reuses wf_epilogue_quad (BatchNorm sums, SiLU', dropout mask).
#include "wf_common.cuh"
#include "wf_elem.h"

namespace {

constexpr int G_NT = 256, G_BN = 256, G_HALO = 16;
constexpr int G_XW = G_HALO + G_BN + G_HALO;       // staged columns
constexpr int G_XS = G_XW + 8;                     // row stride (words): conflict-free fragment loads
constexpr int G_KMAX = 32;                         // channels per group, padded

__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u); }
__device__ __forceinline__ void split4(float4 v, float4& h, float4& l)
{
    h.x = tf32_hi(v.x); h.y = tf32_hi(v.y); h.z = tf32_hi(v.z); h.w = tf32_hi(v.w);
    l.x = tf32_hi(v.x - h.x); l.y = tf32_hi(v.y - h.y); l.z = tf32_hi(v.z - h.z); l.w = tf32_hi(v.w - h.w);
}
// D(16x8) += A(16x8, row) * B(8x8, col), tf32 in / fp32 accumulate
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2])
{
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int BM>       // 16 or 32 output channels (the whole group)
__global__ void __launch_bounds__(G_NT, 2) group_conv_kernel(const ConvP p)
{
    constexpr int MI = BM / 16, WS = BM + 8;
    extern __shared__ __align__(16) float smem[];
    float* Xs = smem;                                         // [hi|lo][G_KMAX][G_XS]
    float* Ws = smem + 2 * G_KMAX * G_XS;                     // [hi|lo][3][G_KMAX][WS]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = blockIdx.z;
    const int n0 = blockIdx.x * G_BN;
    const int K8 = (p.Cin + 7) / 8 * 8;                       // channels rounded up to the mma K step (rows >= Cin are zero)

    // ---- stage the weights of this group: [tap][k][m] ----
    for (int idx = tid; idx < 3 * K8 * (BM / 4); idx += G_NT) {
        const int mq = idx % (BM / 4), k = (idx / (BM / 4)) % K8, tap = idx / ((BM / 4) * K8);
        float4 v = f4zero();
        if (k < p.Kpad && tap < p.ntaps) v = ld4(p.w + ((size_t)(g * p.ntaps + tap) * p.Kpad + k) * p.Mpad + mq * 4);
        float4 h, l;
        split4(v, h, l);
        st4(Ws + (tap * G_KMAX + k) * WS + mq * 4, h);
        st4(Ws + (3 * G_KMAX + tap * G_KMAX + k) * WS + mq * 4, l);
    }
    // ---- stage the activation tile once, prologue applied ----
    for (int idx = tid; idx < K8 * (G_XW / 4); idx += G_NT) {
        const int j = idx % (G_XW / 4), k = idx / (G_XW / 4);
        const int nn = n0 - G_HALO + 4 * j;
        float4 v = f4zero();
        if (k < p.Cin && nn >= 0 && nn < p.N) {
            const int b = nn / WF_T, t = nn - b * WF_T;
            const int c = g * p.Cin + k;
            const long long off = (long long)c * p.in_sc + (long long)b * p.in_sb + t;
            v = ld4(p.in + off);
            if (p.pro_mode == PRO_BNSILU) {
                const float a = p.pro_a[c], bb = p.pro_b[c], mu = p.pro_d[c];
                v.x = wf_silu(fmaf(a, v.x - mu, bb)); v.y = wf_silu(fmaf(a, v.y - mu, bb));
                v.z = wf_silu(fmaf(a, v.z - mu, bb)); v.w = wf_silu(fmaf(a, v.w - mu, bb));
                if (p.mask) {
                    const float* mp = p.mask + (long long)b * p.m_sb + (long long)c * p.m_sc + (long long)t * p.m_st;
                    if (p.m_st == 1) { const float4 m = ld4(mp); v.x *= m.x; v.y *= m.y; v.z *= m.z; v.w *= m.w; }
                    else { const float m = *mp; v.x *= m; v.y *= m; v.z *= m; v.w *= m; }
                }
            } else if (p.pro_mode == PRO_AFFINE) {
                const float a = p.pro_a[c], bb = p.pro_b[c], mu = p.pro_d[c];
                v.x = fmaf(a, v.x - mu, bb); v.y = fmaf(a, v.y - mu, bb); v.z = fmaf(a, v.z - mu, bb); v.w = fmaf(a, v.w - mu, bb);
            } else if (p.pro_mode == PRO_BNBWD) {
                const float4 r = ld4(p.in2 + off);
                const float a = p.pro_a[c], bb = p.pro_b[c], cc = p.pro_c[c], mu = p.pro_d[c];
                v.x = fmaf(a, v.x, fmaf(bb, r.x - mu, cc)); v.y = fmaf(a, v.y, fmaf(bb, r.y - mu, cc));
                v.z = fmaf(a, v.z, fmaf(bb, r.z - mu, cc)); v.w = fmaf(a, v.w, fmaf(bb, r.w - mu, cc));
            }
        }
        float4 h, l;
        split4(v, h, l);
        st4(Xs + k * G_XS + 4 * j, h);
        st4(Xs + (G_KMAX + k) * G_XS + 4 * j, l);
    }
    __syncthreads();

    // ---- three taps as shifted fragment reads ----
    const int fr = lane >> 2, fc = lane & 3;
    const uint32_t* xh = reinterpret_cast<const uint32_t*>(Xs);
    const uint32_t* xl = xh + G_KMAX * G_XS;
    const uint32_t* wh = reinterpret_cast<const uint32_t*>(Ws);
    const uint32_t* wl = wh + 3 * G_KMAX * WS;
    float acc[MI][4][4], cor[MI][4][4];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) { acc[i][j][e] = 0.f; cor[i][j][e] = 0.f; }
    int tcol[4];
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) tcol[ni] = (n0 + warp * 32 + ni * 8 + fr) % WF_T;

    for (int tap = 0; tap < p.ntaps; ++tap) {
        const int dn = p.dn[tap];
        bool ok[4];
        int xoff[4];
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
            const int ts = tcol[ni] + dn;
            ok[ni] = ts >= 0 && ts < WF_T;
            xoff[ni] = G_HALO + warp * 32 + ni * 8 + fr + dn;
        }
        for (int k8 = 0; k8 < K8; k8 += 8) {
            uint32_t fah[MI][4], fal[MI][4], fbh[4][2], fbl[4][2];
#pragma unroll
            for (int mi = 0; mi < MI; ++mi) {
                const int o0 = (tap * G_KMAX + k8 + fc) * WS + mi * 16 + fr, o1 = o0 + 4 * WS;
                fah[mi][0] = wh[o0]; fah[mi][1] = wh[o0 + 8]; fah[mi][2] = wh[o1]; fah[mi][3] = wh[o1 + 8];
                fal[mi][0] = wl[o0]; fal[mi][1] = wl[o0 + 8]; fal[mi][2] = wl[o1]; fal[mi][3] = wl[o1 + 8];
            }
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
                const int o0 = (k8 + fc) * G_XS + xoff[ni], o1 = o0 + 4 * G_XS;
                fbh[ni][0] = ok[ni] ? xh[o0] : 0u; fbh[ni][1] = ok[ni] ? xh[o1] : 0u;
                fbl[ni][0] = ok[ni] ? xl[o0] : 0u; fbl[ni][1] = ok[ni] ? xl[o1] : 0u;
            }
#pragma unroll
            for (int mi = 0; mi < MI; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) {
                    mma_tf32(cor[mi][ni], fal[mi], fbh[ni]);
                    mma_tf32(cor[mi][ni], fah[mi], fbl[ni]);
                    mma_tf32(acc[mi][ni], fah[mi], fbh[ni]);
                }
        }
    }

    // ---- epilogue ----
    // accumulator fragment: lane holds (row fr, cols 2fc, 2fc+1) and (row fr+8, same cols) of every 16x8 tile.  Lanes fc and
    // fc^1 swap halves so that the even lane owns 4 consecutive columns of row fr and the odd lane those of row fr+8.
    const bool want_stats = (p.epi_mode != EPI_STORE) && (p.stat0 != nullptr);
    const bool odd = fc & 1;
#pragma unroll
    for (int mi = 0; mi < MI; ++mi) {
        const int m = mi * 16 + fr + (odd ? 8 : 0);
        const bool mv = m < p.Cout;
        const int co = g * p.Cout + m;
        float s0 = 0.f, s1 = 0.f;
        float bias = 0.f, es = 0.f, et = 0.f, em = 0.f;
        if (mv) {
            if (p.bias) bias = p.bias[co];
            if (p.epi_mode == EPI_DSILU) { es = p.e_scale[co]; et = p.e_shift[co]; }
            if (p.epi_mode == EPI_DSILU || p.epi_mode == EPI_DAFF) em = p.e_mean[co];
        }
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
            const float d0 = acc[mi][ni][0] + cor[mi][ni][0], d1 = acc[mi][ni][1] + cor[mi][ni][1];
            const float d2 = acc[mi][ni][2] + cor[mi][ni][2], d3 = acc[mi][ni][3] + cor[mi][ni][3];
            const float sx = odd ? d0 : d2, sy = odd ? d1 : d3;      // even lane sends its row fr+8 half, odd lane its row fr half
            const float rx = __shfl_xor_sync(0xffffffffu, sx, 1), ry = __shfl_xor_sync(0xffffffffu, sy, 1);
            float v[4];
            if (odd) { v[0] = rx; v[1] = ry; v[2] = d2; v[3] = d3; }
            else { v[0] = d0; v[1] = d1; v[2] = rx; v[3] = ry; }
            const int n = n0 + warp * 32 + ni * 8 + (fc >> 1) * 4;
            if (mv && n < p.N) {
                v[0] += bias; v[1] += bias; v[2] += bias; v[3] += bias;
                wf_epilogue_quad(p, co, 0, n, es, et, em, v, s0, s1);
            }
        }
        if (want_stats) {
            s0 += __shfl_xor_sync(0xffffffffu, s0, 2);               // lanes fc and fc^2 share the row
            s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
            if (fc < 2 && mv) {
                atomicAdd(p.stat0 + co, (double)s0);
                atomicAdd(p.stat1 + co, (double)s1);
            }
        }
    }
}

template <int BM>
cudaError_t launch_group(const ConvP& p, cudaStream_t st)
{
    constexpr int smem = (2 * G_KMAX * G_XS + 2 * 3 * G_KMAX * (BM + 8)) * 4;
    static bool cfg = false;
    if (!cfg) {
        cudaError_t e = cudaFuncSetAttribute(group_conv_kernel<BM>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        cfg = true;
    }
    dim3 grid((p.N + G_BN - 1) / G_BN, 1, p.groups);
    group_conv_kernel<BM><<<grid, G_NT, smem, st>>>(p);
    return cudaGetLastError();
}

}  // namespace

// time-tap convs with whole groups of <= 32 channels: every tap is a column shift of at most the halo
bool wf_group_conv_ok(const ConvP& p)
{
    if (p.Pin != 1 || p.Pout != 1 || p.pmul != 1 || p.pdiv != 1 || p.ntaps != 3) return false;
    if (p.Cin > G_KMAX || p.Cout > 32 || (p.Mpad != 16 && p.Mpad != 32) || p.Mpad < p.Cout) return false;
    for (int t = 0; t < p.ntaps; ++t)
        if (p.dp[t] != 0 || p.dn[t] < -G_HALO || p.dn[t] > G_HALO) return false;
    return true;
}

cudaError_t wf_launch_group_conv(const ConvP& p, cudaStream_t st)
{
    return p.Mpad == 16 ? launch_group<16>(p, st) : launch_group<32>(p, st);
}
