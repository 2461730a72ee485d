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
#include <cstdlib>
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
    wf_pdl_enter();
    constexpr int MI = BM / 16, WS = BM + 8;
    extern __shared__ __align__(16) float smem[];
    float* Xs = smem;                                         // [hi|lo][G_KMAX][G_XS]
    float* Ws = smem + 2 * G_KMAX * G_XS;                     // [hi|lo][3][G_KMAX][WS]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = blockIdx.z;
    const int n0 = blockIdx.x * G_BN;
    const int K8 = (p.Cin + 7) / 8 * 8;                       // channels rounded up to the mma K step (rows >= Cin are zero)
    __shared__ double red[2][32];                             // BatchNorm sums of the CTA's rows: the 8 warps meet here first
    if (tid < 64) red[tid >> 5][tid & 31] = 0.0;

    // ---- stage the weights of this group: [tap][k][m]; all loads of a thread are in flight before the first is used ----
    {
        constexpr int WU = (3 * G_KMAX * (BM / 4) + G_NT - 1) / G_NT;
        float4 wv[WU];
#pragma unroll
        for (int u = 0; u < WU; ++u) {
            const int idx = tid + u * G_NT;
            const int mq = idx % (BM / 4), k = (idx / (BM / 4)) % K8, tap = idx / ((BM / 4) * K8);
            wv[u] = f4zero();
            if (idx < 3 * K8 * (BM / 4) && k < p.Kpad && tap < p.ntaps) wv[u] = ld4(p.w + ((size_t)(g * p.ntaps + tap) * p.Kpad + k) * p.Mpad + mq * 4);
        }
        // ---- stage the activation tile once, prologue applied.  The tile is K8 x 72 quads = at most 9 per thread: the value
        //      (and mask / second-tensor) loads of all of them are issued back to back, then transformed (the kernel used to pay
        //      one DRAM round trip per quad: 40 % of its stall samples sat on these loads) ----
        constexpr int XU = (G_KMAX * (G_XW / 4) + G_NT - 1) / G_NT;
        float4 xv[XU], xw[XU];
        int xc[XU];
#pragma unroll
        for (int u = 0; u < XU; ++u) {
            const int idx = tid + u * G_NT;
            const int j = idx % (G_XW / 4), k = idx / (G_XW / 4);
            const int nn = n0 - G_HALO + 4 * j;
            xv[u] = f4zero(); xw[u] = make_float4(1.f, 1.f, 1.f, 1.f); xc[u] = -1;
            if (idx < K8 * (G_XW / 4) && k < p.Cin && nn >= 0 && nn < p.N) {
                const int b = nn / WF_T, t = nn - b * WF_T;
                const int c = g * p.Cin + k;
                const long long off = (long long)c * p.in_sc + (long long)b * p.in_sb + t;
                xc[u] = c;
                xv[u] = ld4(p.in + off);
                if (p.pro_mode == PRO_BNBWD) xw[u] = ld4(p.in2 + off);
                else if (p.pro_mode == PRO_BNSILU && p.mask) {
                    const float* mp = p.mask + (long long)b * p.m_sb + (long long)c * p.m_sc + (long long)t * p.m_st;
                    if (p.m_st == 1) xw[u] = ld4(mp); else { const float m = *mp; xw[u] = make_float4(m, m, m, m); }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < WU; ++u) {
            const int idx = tid + u * G_NT;
            if (idx < 3 * K8 * (BM / 4)) {
                const int mq = idx % (BM / 4), k = (idx / (BM / 4)) % K8, tap = idx / ((BM / 4) * K8);
                float4 h, l;
                split4(wv[u], h, l);
                st4(Ws + (tap * G_KMAX + k) * WS + mq * 4, h);
                st4(Ws + (3 * G_KMAX + tap * G_KMAX + k) * WS + mq * 4, l);
            }
        }
#pragma unroll
        for (int u = 0; u < XU; ++u) {
            const int idx = tid + u * G_NT;
            if (idx < K8 * (G_XW / 4)) {
                const int j = idx % (G_XW / 4), k = idx / (G_XW / 4);
                float4 v = xv[u];
                if (xc[u] >= 0) {
                    const int c = xc[u];
                    if (p.pro_mode == PRO_BNSILU) {
                        const float a = p.pro_a[c], bb = p.pro_b[c], mu = p.pro_d[c];
                        v.x = wf_silu(fmaf(a, v.x - mu, bb)) * xw[u].x; v.y = wf_silu(fmaf(a, v.y - mu, bb)) * xw[u].y;
                        v.z = wf_silu(fmaf(a, v.z - mu, bb)) * xw[u].z; v.w = wf_silu(fmaf(a, v.w - mu, bb)) * xw[u].w;
                    } else if (p.pro_mode == PRO_AFFINE) {
                        const float a = p.pro_a[c], bb = p.pro_b[c], mu = p.pro_d[c];
                        v.x = fmaf(a, v.x - mu, bb); v.y = fmaf(a, v.y - mu, bb); v.z = fmaf(a, v.z - mu, bb); v.w = fmaf(a, v.w - mu, bb);
                    } else if (p.pro_mode == PRO_BNBWD) {
                        const float a = p.pro_a[c], bb = p.pro_b[c], cc = p.pro_c[c], mu = p.pro_d[c];
                        v.x = fmaf(a, v.x, fmaf(bb, xw[u].x - mu, cc)); v.y = fmaf(a, v.y, fmaf(bb, xw[u].y - mu, cc));
                        v.z = fmaf(a, v.z, fmaf(bb, xw[u].z - mu, cc)); v.w = fmaf(a, v.w, fmaf(bb, xw[u].w - mu, cc));
                    }
                }
                float4 h, l;
                split4(v, h, l);
                st4(Xs + k * G_XS + 4 * j, h);
                st4(Xs + (G_KMAX + k) * G_XS + 4 * j, l);
            }
        }
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
                atomicAdd(&red[0][m], (double)s0);
                atomicAdd(&red[1][m], (double)s1);
            }
        }
    }
    // one global reduction pair per row per CTA (was one per warp: 8x the traffic onto the same 2 x Cout addresses of the group)
    if (want_stats) {
        __syncthreads();
        if (tid < p.Cout) {
            atomicAdd(p.stat0 + g * p.Cout + tid, red[0][tid]);
            atomicAdd(p.stat1 + g * p.Cout + tid, red[1][tid]);
        }
    }
    wf_bn_tail(p.tail);
}

template <int BM>
cudaError_t launch_group(const ConvP& p, cudaStream_t st)
{
    constexpr int smem = (2 * G_KMAX * G_XS + 2 * 3 * G_KMAX * (BM + 8)) * 4;
    static WfSmemOptIn optin;
    if (cudaError_t e = wf_smem_optin(optin, group_conv_kernel<BM>, smem)) return e;
    dim3 grid((p.N + G_BN - 1) / G_BN, 1, p.groups);
    wf_launch_pdl(group_conv_kernel<BM>, dim3(grid), dim3(G_NT), smem, st, p);
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

// =========================================================================================================
// Backward-weights of the grouped causal convs:
//   dW[co][ci][tap] = sum_n G[co][n] * X'[ci][n + dn(tap)]      (terms whose source column leaves the 20-step window dropped)
// One CTA owns one group and a range of columns; per 64-column chunk it stages G (BatchNorm-backward applied) and X'
// (BatchNorm+SiLU+Dropout applied, with a 16-column left halo) ONCE and produces all three taps from them: the tap shift is
// a shifted fragment read of X', the window rule a per-column zeroing of the G fragment.  Each of the 8 warps takes one
// 8-column slice of the chunk for the whole 32 x 32 x 3 output; the warps' partial sums meet in shared memory and leave as
// one fp32 atomic per weight per CTA.
// =========================================================================================================
namespace {

constexpr int GW_KCH = 64;                         // columns per chunk (8 warps x one k8 step)
constexpr int GW_GS = GW_KCH + 4;                  // row strides (words): conflict-free fragment loads
constexpr int GW_XS = G_HALO + GW_KCH + 4;
constexpr int GW_STAGE = 2 * 32 * GW_GS + 2 * 32 * GW_XS;      // words per stage: G hi/lo + X hi/lo

__global__ void __launch_bounds__(G_NT, 1) group_wgrad_kernel(const WgradP p, long long cols_per_split)
{
    wf_pdl_enter();
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = blockIdx.z;
    const long long kbegin = (long long)blockIdx.x * cols_per_split;
    long long kend = kbegin + cols_per_split;
    if (kend > p.N) kend = p.N;
    const int nchunks = kend > kbegin ? (int)((kend - kbegin + GW_KCH - 1) / GW_KCH) : 0;
    auto Gs = [&](int buf, int hl) { return smem + buf * GW_STAGE + hl * 32 * GW_GS; };
    auto Xs = [&](int buf, int hl) { return smem + buf * GW_STAGE + 2 * 32 * GW_GS + hl * 32 * GW_XS; };

    // loader assignment: G: 32 rows x 16 quads = 512 items (2 per thread); X: 32 rows x 20 quads = 640 items (3 passes)
    float4 rg[2], rg2[2], rx[3], rm[3];
    auto load_chunk = [&](int ch) {
        const long long c0 = kbegin + (long long)ch * GW_KCH;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int idx = tid + i * G_NT, row = idx >> 4, q = idx & 15;
            const long long col = c0 + q * 4;
            rg[i] = f4zero(); rg2[i] = f4zero();
            if (row < p.Cout && col < kend) {
                const long long off = (long long)(g * p.Cout + row) * p.N + col;
                rg[i] = ld4(p.g + off);
                if (p.g_pro == PRO_BNBWD) rg2[i] = ld4(p.g2 + off);
            }
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int idx = tid + i * G_NT, row = idx / 20, q = idx % 20;
            const long long col = c0 - G_HALO + q * 4;
            rx[i] = f4zero(); rm[i] = make_float4(1.f, 1.f, 1.f, 1.f);
            if (idx < 640 && row < p.Cin && col >= 0 && col < p.N) {
                const long long b = col / WF_T; const int t = (int)(col - b * WF_T);
                const int c = g * p.Cin + row;
                rx[i] = ld4(p.in + (long long)c * p.in_sc + b * p.in_sb + t);
                if (p.pro_mode == PRO_BNSILU && p.mask) {
                    const float* mp = p.mask + b * p.m_sb + (long long)c * p.m_sc + (long long)t * p.m_st;
                    if (p.m_st == 1) rm[i] = ld4(mp); else { const float mm = *mp; rm[i] = make_float4(mm, mm, mm, mm); }
                }
            }
        }
    };
    auto store_chunk = [&](int buf, int ch) {
        const long long c0 = kbegin + (long long)ch * GW_KCH;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int idx = tid + i * G_NT, row = idx >> 4, q = idx & 15;
            float4 v = rg[i];
            if (p.g_pro == PRO_BNBWD && row < p.Cout && c0 + q * 4 < kend) {
                const int c = g * p.Cout + row;
                const float a = p.g_a[c], b = p.g_b[c], d = p.g_c[c], mu = p.g_d[c];
                v.x = fmaf(a, v.x, fmaf(b, rg2[i].x - mu, d)); v.y = fmaf(a, v.y, fmaf(b, rg2[i].y - mu, d));
                v.z = fmaf(a, v.z, fmaf(b, rg2[i].z - mu, d)); v.w = fmaf(a, v.w, fmaf(b, rg2[i].w - mu, d));
            }
            float4 h, l;
            split4(v, h, l);
            st4(Gs(buf, 0) + row * GW_GS + q * 4, h);
            st4(Gs(buf, 1) + row * GW_GS + q * 4, l);
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int idx = tid + i * G_NT, row = idx / 20, q = idx % 20;
            if (idx < 640) {
                float4 v = rx[i];
                const long long col = c0 - G_HALO + q * 4;
                if (row < p.Cin && col >= 0 && col < p.N && p.pro_mode != PRO_NONE) {
                    const int c = g * p.Cin + row;
                    const float a = p.pro_a[c], bb = p.pro_b[c], mu = p.pro_d[c];
                    if (p.pro_mode == PRO_BNSILU) {
                        v.x = wf_silu(fmaf(a, v.x - mu, bb)) * rm[i].x; v.y = wf_silu(fmaf(a, v.y - mu, bb)) * rm[i].y;
                        v.z = wf_silu(fmaf(a, v.z - mu, bb)) * rm[i].z; v.w = wf_silu(fmaf(a, v.w - mu, bb)) * rm[i].w;
                    } else {
                        v.x = fmaf(a, v.x - mu, bb); v.y = fmaf(a, v.y - mu, bb); v.z = fmaf(a, v.z - mu, bb); v.w = fmaf(a, v.w - mu, bb);
                    }
                }
                float4 h, l;
                split4(v, h, l);
                st4(Xs(buf, 0) + row * GW_XS + q * 4, h);
                st4(Xs(buf, 1) + row * GW_XS + q * 4, l);
            }
        }
    };

    float acc[3][2][4][4];
#pragma unroll
    for (int t = 0; t < 3; ++t)
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[t][i][j][e] = 0.f;

    const int fr = lane >> 2, fc = lane & 3;
    const int kcol = warp * 8 + fc;                                  // this lane's fragment columns inside a chunk: kcol, kcol + 4
    int t0 = (int)((kbegin + kcol) % WF_T);
    if (nchunks > 0) { load_chunk(0); store_chunk(0, 0); }
    __syncthreads();
    int buf = 0;
    for (int ch = 0; ch < nchunks; ++ch) {
        const bool more = ch + 1 < nchunks;
        if (more) load_chunk(ch + 1);
        const uint32_t* gh = reinterpret_cast<const uint32_t*>(Gs(buf, 0));
        const uint32_t* gl = reinterpret_cast<const uint32_t*>(Gs(buf, 1));
        const uint32_t* xh = reinterpret_cast<const uint32_t*>(Xs(buf, 0));
        const uint32_t* xl = reinterpret_cast<const uint32_t*>(Xs(buf, 1));
        const int t1 = t0 + 4 >= WF_T ? t0 + 4 - WF_T : t0 + 4;
        uint32_t fah[2][4], fal[2][4];
#pragma unroll
        for (int mi = 0; mi < 2; ++mi) {
            const int o0 = (mi * 16 + fr) * GW_GS + kcol, o1 = o0 + 8 * GW_GS;
            fah[mi][0] = gh[o0]; fah[mi][1] = gh[o1]; fah[mi][2] = gh[o0 + 4]; fah[mi][3] = gh[o1 + 4];
            fal[mi][0] = gl[o0]; fal[mi][1] = gl[o1]; fal[mi][2] = gl[o0 + 4]; fal[mi][3] = gl[o1 + 4];
        }
#pragma unroll
        for (int tap = 0; tap < 3; ++tap) {
            const int dn = p.dn[tap];
            const bool v0 = (t0 + dn >= 0) && (t0 + dn < WF_T), v1 = (t1 + dn >= 0) && (t1 + dn < WF_T);
            uint32_t ah[2][4], al[2][4];
#pragma unroll
            for (int mi = 0; mi < 2; ++mi) {
                ah[mi][0] = v0 ? fah[mi][0] : 0u; ah[mi][1] = v0 ? fah[mi][1] : 0u; ah[mi][2] = v1 ? fah[mi][2] : 0u; ah[mi][3] = v1 ? fah[mi][3] : 0u;
                al[mi][0] = v0 ? fal[mi][0] : 0u; al[mi][1] = v0 ? fal[mi][1] : 0u; al[mi][2] = v1 ? fal[mi][2] : 0u; al[mi][3] = v1 ? fal[mi][3] : 0u;
            }
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
                const int o0 = (ni * 8 + fr) * GW_XS + G_HALO + kcol + dn;
                uint32_t bh[2] = {xh[o0], xh[o0 + 4]}, bl[2] = {xl[o0], xl[o0 + 4]};
#pragma unroll
                for (int mi = 0; mi < 2; ++mi) {
                    mma_tf32(acc[tap][mi][ni], al[mi], bh);
                    mma_tf32(acc[tap][mi][ni], ah[mi], bl);
                    mma_tf32(acc[tap][mi][ni], ah[mi], bh);
                }
            }
        }
        if (more) store_chunk(buf ^ 1, ch + 1);
        __syncthreads();
        buf ^= 1;
        t0 = t0 + (GW_KCH % WF_T) >= WF_T ? t0 + (GW_KCH % WF_T) - WF_T : t0 + (GW_KCH % WF_T);
    }

    // the 8 warps' partial sums meet in shared memory (stage buffers are free now) through shared-memory reductions -- all warps at
    // once instead of eight barrier-separated turns --, then one global reduction per weight per CTA
    float* red = smem;                                               // [3][32][32]
    for (int idx = tid; idx < 3 * 32 * 32; idx += G_NT) red[idx] = 0.f;
    __syncthreads();
#pragma unroll
    for (int tap = 0; tap < 3; ++tap)
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int co = mi * 16 + fr + (e >= 2 ? 8 : 0), ci = ni * 8 + fc * 2 + (e & 1);
                    if (co < p.Cout && ci < p.Cin) atomicAdd(red + (tap * 32 + co) * 32 + ci, acc[tap][mi][ni][e]);
                }
    __syncthreads();
    // the group's weights are one contiguous block of dW ([co][ci][tap]): consecutive threads reduce into consecutive addresses
    {
        const int per_co = p.Cin * p.ntaps, total = p.Cout * per_co;
        float* dwg = p.dw + (size_t)g * total;
        for (int idx = tid; idx < total; idx += G_NT) {
            const int co = idx / per_co, rem = idx - co * per_co, ci = rem / p.ntaps, tap = rem - ci * p.ntaps;
            atomicAdd(dwg + idx, red[(tap * 32 + co) * 32 + ci]);
        }
    }
}

}  // namespace

bool wf_group_wgrad_ok(const WgradP& p)
{
    if (p.Pin != 1 || p.Pout != 1 || p.pmul != 1 || p.ntaps != 3 || p.Cin > 32 || p.Cout > 32) return false;
    for (int t = 0; t < p.ntaps; ++t)
        if (p.dp[t] != 0 || p.dn[t] < -G_HALO || p.dn[t] > 0) return false;
    return true;
}

cudaError_t wf_launch_group_wgrad(const WgradP& p, int num_sms, cudaStream_t st)
{
    constexpr int smem = 2 * GW_STAGE * 4;
    static_assert(2 * GW_STAGE >= 3 * 32 * 32, "reduction buffer must fit the stage buffers");
    static WfSmemOptIn optin;
    if (cudaError_t e = wf_smem_optin(optin, group_wgrad_kernel, smem)) return e;
    static const int rounds = [] { const char* e = std::getenv("WF_GROUP_WGRAD_ROUNDS"); const int v = e ? std::atoi(e) : 0; return v > 0 ? v : 1; }();
    long long splits = ((long long)rounds * num_sms) / p.groups;         // one round of one-CTA-per-SM (measured: every extra round costs 14 us per launch in prime + reduction)
    const long long max_splits = (p.N + 8LL * GW_KCH - 1) / (8LL * GW_KCH);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    long long per = (p.N + splits - 1) / splits;
    per = (per + GW_KCH - 1) / GW_KCH * GW_KCH;
    splits = (p.N + per - 1) / per;
    dim3 grid((unsigned)splits, 1, p.groups);
    wf_launch_pdl(group_wgrad_kernel, dim3(grid), dim3(G_NT), smem, st, p, per);
    return cudaGetLastError();
}
