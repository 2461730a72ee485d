// Implicit-GEMM convolution kernels (forward / backward-data share one kernel, backward-weights is a
// second one) for every conv-like layer of WiFlow:
//   * TCN grouped causal dilated convs   (models/tcn.py:20-22,33-35 + Chomp1d tcn.py:11-12)
//   * TCN / attention / decoder 1x1 convs (tcn.py:27,40,47; attention.py:22-24; pose_model.py:48)
//   * (1x3) strided convs + 1x1 shortcuts (convnet.py:11-12,17,22,27,48,53,58,63)
//   * decoder 3x3 conv                    (pose_model.py:45)
// The GEMM is  D[m][col] = sum_{tap,k} W[tap][k][m] * act(X[k][ipos(tap)][col + dn(tap)])  with col = n = b*20+t
// contiguous, so a conv tap along the position axis is a row-pointer shift and a tap along time is a column
// shift masked at window borders.  BatchNorm(+SiLU+Dropout) of the *previous* layer is applied while the operand
// tile is staged into shared memory (prologue), and the BatchNorm statistics of *this* layer are reduced in the
// epilogue (warp shuffles + one fp64 atomic per row per warp), so activations cross HBM once per layer.
#include "wf_common.cuh"
#include "wf_elem.h"

namespace {

template <int TM> struct RowMap {
    __device__ static __forceinline__ int row(int ty, int i, int BM) { return ty * 4 + i; }
};
template <> struct RowMap<8> {
    __device__ static __forceinline__ int row(int ty, int i, int BM) { return (i < 4) ? ty * 4 + i : BM / 2 + ty * 4 + (i - 4); }
};

template <int BM, int BN, int BK, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
conv_gemm_kernel(const ConvP p)
{
    wf_pdl_enter();
    constexpr int NTX = BN / TN, NTY = BM / TM, NT = NTX * NTY;
    constexpr int QPR = BN / 4;                        // column quads per B-tile row
    static_assert(NT % QPR == 0, "thread count must be a multiple of the quads per row");
    constexpr int KSTEP = NT / QPR;                    // B-tile rows covered per loader iteration
    static_assert(BK % KSTEP == 0, "BK must be a multiple of KSTEP");
    constexpr int B_ITERS = BK / KSTEP;
    constexpr int A_ITEMS = BK * BM / 4;
    constexpr int A_ITERS = (A_ITEMS + NT - 1) / NT;
    static_assert(TM == 4 || TM == 8, "TM");
    static_assert(TN == 4 || TN == 8, "TN");
    static_assert(NTX == 16 || NTX % 32 == 0, "NTX");

    __shared__ __align__(16) float As[2][BK][BM];
    __shared__ __align__(16) float Bs[2][BK][BN];

    const int tid = threadIdx.x;
    const int tx = tid % NTX, ty = tid / NTX;
    const int mtiles = p.Mpad / BM;
    const int opos = blockIdx.y / mtiles;
    const int m0 = (blockIdx.y % mtiles) * BM;
    const int g = blockIdx.z;
    const int n0 = blockIdx.x * BN;
    const int KT = p.Kpad / BK;
    const int S = p.ntaps * KT;

    // loader column (fixed per thread)
    const int lq = tid % QPR;
    const int lkk0 = tid / QPR;
    const int ln = n0 + lq * 4;
    const bool lnvalid = ln < p.N;
    const int lb = ln / WF_T, lt = ln % WF_T;
    const long long lcol_off = (long long)lb * p.in_sb + lt;
    const long long lmask_off = (long long)lb * p.m_sb + (long long)lt * p.m_st;

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    float4 ra[A_ITERS];
    float4 rb[B_ITERS], rb2[B_ITERS];
    float ca[B_ITERS], cb[B_ITERS], cc[B_ITERS], cd[B_ITERS];
    unsigned okb[B_ITERS];

    auto tap_ipos = [&](int tap, int& ipos) -> bool {
        int num = opos * p.pmul + p.dp[tap];
        if (num < 0) return false;
        if (p.pdiv > 1) { if (num % p.pdiv) return false; num /= p.pdiv; }
        ipos = num;
        return num < p.Pin;
    };
    auto next_stage = [&](int s) -> int {          // first valid stage index >= s, or S
        while (s < S) {
            int ipos;
            if (tap_ipos(s / KT, ipos)) return s;
            s = (s / KT + 1) * KT;
        }
        return S;
    };

    auto load_stage = [&](int s) {
        const int tap = s / KT, k0 = (s % KT) * BK;
        int ipos = 0;
        tap_ipos(tap, ipos);
        const int dn = p.dn[tap];
        const float* wbase = p.w + ((size_t)(g * p.ntaps + tap) * p.Kpad + k0) * p.Mpad + m0;
#pragma unroll
        for (int i = 0; i < A_ITERS; ++i) {
            int idx = tid + i * NT;
            if (A_ITEMS % NT == 0 || idx < A_ITEMS) {
                int kk = idx / (BM / 4), mq = idx % (BM / 4);
                ra[i] = ld4(wbase + (size_t)kk * p.Mpad + mq * 4);
            }
        }
#pragma unroll
        for (int i = 0; i < B_ITERS; ++i) {
            const int k = k0 + lkk0 + i * KSTEP;
            const bool kv = lnvalid && (k < p.Cin);
            const int c = g * p.Cin + k;
            const int ts = lt + dn;                                  // source time index of element 0
            unsigned ok = 0;
            float4 v = f4zero(), v2 = f4zero();
            if (kv) {
                const float* src = p.in + (long long)c * p.in_sc + (long long)ipos * p.in_sp + lcol_off + dn;
                if ((dn & 3) == 0) {
                    if (ts >= 0 && ts < WF_T) {
                        ok = 0xF;
                        v = ld4(src);
                        if (p.pro_mode == PRO_BNBWD) v2 = ld4(p.in2 + (src - p.in));
                        else if (p.pro_mode == PRO_BNSILU && p.mask) {
                            const float* mp = p.mask + lmask_off + (long long)c * p.m_sc + (long long)dn * p.m_st;
                            if (p.m_st == 1) v2 = ld4(mp);
                            else { float m = *mp; v2 = make_float4(m, m, m, m); }
                        }
                    }
                } else {
                    float e[4] = {0.f, 0.f, 0.f, 0.f}, e2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (ts + j >= 0 && ts + j < WF_T) {
                            ok |= 1u << j;
                            e[j] = src[j];
                            if (p.pro_mode == PRO_BNBWD) e2[j] = p.in2[(src - p.in) + j];
                            else if (p.pro_mode == PRO_BNSILU && p.mask)
                                e2[j] = p.mask[lmask_off + (long long)c * p.m_sc + (long long)(dn + j) * p.m_st];
                        }
                    }
                    v = make_float4(e[0], e[1], e[2], e[3]);
                    v2 = make_float4(e2[0], e2[1], e2[2], e2[3]);
                }
                if (p.pro_mode != PRO_NONE) {
                    ca[i] = p.pro_a[c];
                    cb[i] = p.pro_b[c];
                    cd[i] = p.pro_d[c];
                    if (p.pro_mode == PRO_BNBWD) cc[i] = p.pro_c[c];
                }
            }
            rb[i] = v; rb2[i] = v2; okb[i] = ok;
        }
    };

    auto store_stage = [&](int buf) {
#pragma unroll
        for (int i = 0; i < A_ITERS; ++i) {
            int idx = tid + i * NT;
            if (A_ITEMS % NT == 0 || idx < A_ITEMS) {
                int kk = idx / (BM / 4), mq = idx % (BM / 4);
                st4(&As[buf][kk][mq * 4], ra[i]);
            }
        }
#pragma unroll
        for (int i = 0; i < B_ITERS; ++i) {
            float e[4] = {rb[i].x, rb[i].y, rb[i].z, rb[i].w};
            const float e2[4] = {rb2[i].x, rb2[i].y, rb2[i].z, rb2[i].w};
            const unsigned ok = okb[i];
            if (ok) {
                if (p.pro_mode == PRO_BNSILU) {
                    const bool hm = p.mask != nullptr;
#pragma unroll
                    for (int j = 0; j < 4; ++j) { float y = wf_silu(fmaf(ca[i], e[j] - cd[i], cb[i])); e[j] = hm ? y * e2[j] : y; }
                } else if (p.pro_mode == PRO_AFFINE) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) e[j] = fmaf(ca[i], e[j] - cd[i], cb[i]);
                } else if (p.pro_mode == PRO_BNBWD) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) e[j] = fmaf(ca[i], e[j], fmaf(cb[i], e2[j] - cd[i], cc[i]));
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) if (!((ok >> j) & 1u)) e[j] = 0.f;
            }
            st4(&Bs[buf][lkk0 + i * KSTEP][lq * 4], make_float4(e[0], e[1], e[2], e[3]));
        }
    };

    int s = next_stage(0);
    int buf = 0;
    if (s < S) {
        load_stage(s);
        store_stage(0);
    }
    __syncthreads();
    while (s < S) {
        const int sn = next_stage(s + 1);
        if (sn < S) load_stage(sn);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[TM], b[TN];
            {
                float4 t0 = ld4(&As[buf][kk][RowMap<TM>::row(ty, 0, BM)]);
                a[0] = t0.x; a[1] = t0.y; a[2] = t0.z; a[3] = t0.w;
                if (TM == 8) {
                    float4 t1 = ld4(&As[buf][kk][RowMap<TM>::row(ty, 4, BM)]);
                    a[TM - 4] = t1.x; a[TM - 3] = t1.y; a[TM - 2] = t1.z; a[TM - 1] = t1.w;
                }
                float4 u0 = ld4(&Bs[buf][kk][tx * 4]);
                b[0] = u0.x; b[1] = u0.y; b[2] = u0.z; b[3] = u0.w;
                if (TN == 8) {
                    float4 u1 = ld4(&Bs[buf][kk][BN / 2 + tx * 4]);
                    b[TN - 4] = u1.x; b[TN - 3] = u1.y; b[TN - 2] = u1.z; b[TN - 1] = u1.w;
                }
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (sn < S) store_stage(buf ^ 1);
        __syncthreads();
        buf ^= 1;
        s = sn;
    }

    // ------------------------------ epilogue ------------------------------
    const bool want_stats = (p.epi_mode != EPI_STORE) && (p.stat0 != nullptr);
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int m = m0 + RowMap<TM>::row(ty, i, BM);
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
        for (int h = 0; h < TN / 4; ++h) {
            const int n = n0 + (h == 0 ? tx * 4 : BN / 2 + tx * 4);
            if (mv && n < p.N) {
                float v[4] = {acc[i][h * 4 + 0] + bias, acc[i][h * 4 + 1] + bias, acc[i][h * 4 + 2] + bias, acc[i][h * 4 + 3] + bias};
                wf_epilogue_quad(p, co, opos, n, es, et, em, v, s0, s1);
            }
        }
        if (want_stats) {
            // reduce over the threads that share this row: NTX==16 -> 16-lane halves, else whole warps
            constexpr int RED = (NTX == 16) ? 8 : 16;
#pragma unroll
            for (int o = RED; o > 0; o >>= 1) {
                s0 += __shfl_xor_sync(0xffffffffu, s0, o);
                s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            }
            const bool leader = (NTX == 16) ? ((tid & 15) == 0) : ((tid & 31) == 0);
            if (leader && mv) {
                atomicAdd(p.stat0 + co, (double)s0);
                atomicAdd(p.stat1 + co, (double)s1);
            }
        }
    }
    wf_bn_tail(p.tail);
}

// ---------------------------------------------------------------------------------------------------------
// Backward-weights:  dW[co][ci][tap] = sum_{p,n} G[co][p][n] * act(X[ci][ipos(p,tap)][n + dn(tap)])
// G = BatchNorm-backward of (dy, raw) applied on load.  The (p, n) reduction is split across blockIdx.x and
// combined with fp32 atomics into the (pre-zeroed) gradient buffer in the reference's weight layout.
// ---------------------------------------------------------------------------------------------------------
template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
conv_wgrad_kernel(const WgradP p)
{
    wf_pdl_enter();
    constexpr int BK = 8;
    constexpr int NTX = BN / TN, NTY = BM / TM, NT = NTX * NTY;
    constexpr int A_ITEMS = BM * 2, B_ITEMS = BN * 2;           // float4 items per stage (2 quads per row)
    constexpr int A_ITERS = (A_ITEMS + NT - 1) / NT, B_ITERS = (B_ITEMS + NT - 1) / NT;
    __shared__ __align__(16) float As[2][BK][BM + 4];
    __shared__ __align__(16) float Bs[2][BK][BN + 4];

    const int tid = threadIdx.x;
    const int tx = tid % NTX, ty = tid / NTX;
    const int ntile_n = (p.Cin + BN - 1) / BN;
    const int m0 = (blockIdx.y / ntile_n) * BM;
    const int c0 = (blockIdx.y % ntile_n) * BN;
    const int g = blockIdx.z / p.ntaps, tap = blockIdx.z % p.ntaps;
    const int dn = p.dn[tap], dpos = p.dp[tap];

    // this block's slice of the flattened (p, n) index space, in units of 8 columns
    const long long total8 = ((long long)p.Pout * p.N + 7) / 8;
    const long long per = (total8 + p.kchunks - 1) / p.kchunks;
    const long long q_begin = (long long)blockIdx.x * per;
    long long q_end = q_begin + per;
    if (q_end > total8) q_end = total8;

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    float4 ra[A_ITERS], rb[B_ITERS];

    // loads one float4 (4 consecutive n at one position) of G-row `co` / X-row `ci`, prologue applied
    auto load_g = [&](int co, long long col) -> float4 {
        if (co >= p.Cout || col >= (long long)p.Pout * p.N) return f4zero();
        const int c = g * p.Cout + co;
        const float* src = p.g + (long long)c * p.Pout * p.N + col;
        float4 v = ld4(src);
        if (p.g_pro == PRO_BNBWD) {
            const float4 r = ld4(p.g2 + (src - p.g));
            const float a = p.g_a[c], b = p.g_b[c], d = p.g_c[c], mu = p.g_d[c];
            v.x = fmaf(a, v.x, fmaf(b, r.x - mu, d)); v.y = fmaf(a, v.y, fmaf(b, r.y - mu, d));
            v.z = fmaf(a, v.z, fmaf(b, r.z - mu, d)); v.w = fmaf(a, v.w, fmaf(b, r.w - mu, d));
        }
        return v;
    };
    auto load_x = [&](int ci, long long col) -> float4 {
        if (ci >= p.Cin || col >= (long long)p.Pout * p.N) return f4zero();
        const int opos = (int)(col / p.N);
        const int n = (int)(col % p.N);
        const int ipos = opos * p.pmul + dpos;
        if (ipos < 0 || ipos >= p.Pin) return f4zero();
        const int b = n / WF_T, t = n % WF_T;
        const int c = g * p.Cin + ci;
        const float* src = p.in + (long long)c * p.in_sc + (long long)ipos * p.in_sp + (long long)b * p.in_sb + t + dn;
        float e[4] = {0.f, 0.f, 0.f, 0.f}, e2[4] = {1.f, 1.f, 1.f, 1.f};
        unsigned ok = 0;
        const int ts = t + dn;
        if ((dn & 3) == 0) {
            if (ts >= 0 && ts < WF_T) {
                ok = 0xF;
                float4 v = ld4(src);
                e[0] = v.x; e[1] = v.y; e[2] = v.z; e[3] = v.w;
                if (p.pro_mode == PRO_BNSILU && p.mask) {
                    const float* mp = p.mask + (long long)b * p.m_sb + (long long)c * p.m_sc + (long long)ts * p.m_st;
                    if (p.m_st == 1) { float4 m4 = ld4(mp); e2[0] = m4.x; e2[1] = m4.y; e2[2] = m4.z; e2[3] = m4.w; }
                    else { float mm = *mp; e2[0] = e2[1] = e2[2] = e2[3] = mm; }
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (ts + j >= 0 && ts + j < WF_T) {
                    ok |= 1u << j;
                    e[j] = src[j];
                    if (p.pro_mode == PRO_BNSILU && p.mask)
                        e2[j] = p.mask[(long long)b * p.m_sb + (long long)c * p.m_sc + (long long)(ts + j) * p.m_st];
                }
            }
        }
        if (!ok) return f4zero();
        if (p.pro_mode == PRO_BNSILU) {
            const float a = p.pro_a[c], bb = p.pro_b[c], mu = p.pro_d[c];
#pragma unroll
            for (int j = 0; j < 4; ++j) e[j] = wf_silu(fmaf(a, e[j] - mu, bb)) * e2[j];
        } else if (p.pro_mode == PRO_AFFINE) {
            const float a = p.pro_a[c], bb = p.pro_b[c], mu = p.pro_d[c];
#pragma unroll
            for (int j = 0; j < 4; ++j) e[j] = fmaf(a, e[j] - mu, bb);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) if (!((ok >> j) & 1u)) e[j] = 0.f;
        return make_float4(e[0], e[1], e[2], e[3]);
    };

    auto load_stage = [&](long long q8) {
#pragma unroll
        for (int i = 0; i < A_ITERS; ++i) {
            int idx = tid + i * NT;
            if (A_ITEMS % NT == 0 || idx < A_ITEMS) ra[i] = load_g(m0 + idx % BM, q8 * 8 + (idx / BM) * 4);
        }
#pragma unroll
        for (int i = 0; i < B_ITERS; ++i) {
            int idx = tid + i * NT;
            if (B_ITEMS % NT == 0 || idx < B_ITEMS) rb[i] = load_x(c0 + idx % BN, q8 * 8 + (idx / BN) * 4);
        }
    };
    auto store_stage = [&](int buf) {
#pragma unroll
        for (int i = 0; i < A_ITERS; ++i) {
            int idx = tid + i * NT;
            if (A_ITEMS % NT == 0 || idx < A_ITEMS) {
                int r = idx % BM, kq = (idx / BM) * 4;
                As[buf][kq + 0][r] = ra[i].x; As[buf][kq + 1][r] = ra[i].y; As[buf][kq + 2][r] = ra[i].z; As[buf][kq + 3][r] = ra[i].w;
            }
        }
#pragma unroll
        for (int i = 0; i < B_ITERS; ++i) {
            int idx = tid + i * NT;
            if (B_ITEMS % NT == 0 || idx < B_ITEMS) {
                int r = idx % BN, kq = (idx / BN) * 4;
                Bs[buf][kq + 0][r] = rb[i].x; Bs[buf][kq + 1][r] = rb[i].y; Bs[buf][kq + 2][r] = rb[i].z; Bs[buf][kq + 3][r] = rb[i].w;
            }
        }
    };

    int buf = 0;
    if (q_begin < q_end) { load_stage(q_begin); store_stage(0); }
    __syncthreads();
    for (long long q8 = q_begin; q8 < q_end; ++q8) {
        const bool more = (q8 + 1) < q_end;
        if (more) load_stage(q8 + 1);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; i += 4) {
                float4 t4 = ld4(&As[buf][kk][ty * TM + i]);
                a[i] = t4.x; a[i + 1] = t4.y; a[i + 2] = t4.z; a[i + 3] = t4.w;
            }
#pragma unroll
            for (int j = 0; j < TN; j += 4) {
                float4 t4 = ld4(&Bs[buf][kk][tx * TN + j]);
                b[j] = t4.x; b[j + 1] = t4.y; b[j + 2] = t4.z; b[j + 3] = t4.w;
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (more) store_stage(buf ^ 1);
        __syncthreads();
        buf ^= 1;
    }

#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int co = m0 + ty * TM + i;
        if (co >= p.Cout) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int ci = c0 + tx * TN + j;
            if (ci < p.Cin)
                atomicAdd(p.dw + ((size_t)(g * p.Cout + co) * p.Cin + ci) * p.ntaps + tap, acc[i][j]);
        }
    }
}

}  // namespace

// ------------------------------------------- launchers -------------------------------------------
// tile configurations of the forward / backward-data GEMM
enum { CFG_BIG = 0, CFG_MID = 1, CFG_THIN16 = 2, CFG_THIN8 = 3 };

static int conv_cfg_for(int M)
{
    if (M > 32) return CFG_BIG;
    if (M > 16) return CFG_MID;
    if (M > 8) return CFG_THIN16;
    return CFG_THIN8;
}
int wf_conv_bm_for(int M)
{
    switch (conv_cfg_for(M)) { case CFG_BIG: return 64; case CFG_MID: return 32; case CFG_THIN16: return 16; default: return 8; }
}
int wf_conv_bk_for(int M)
{
    switch (conv_cfg_for(M)) { case CFG_BIG: case CFG_MID: return 8; default: return 4; }
}

template <int BM, int BN, int BK, int TM, int TN>
static cudaError_t launch_conv_t(const ConvP& p, cudaStream_t st)
{
    dim3 grid((p.N + BN - 1) / BN, p.Pout * (p.Mpad / BM), p.groups);
    wf_launch_pdl(conv_gemm_kernel<BM, BN, BK, TM, TN>, dim3(grid), dim3((BM / TM) * (BN / TN)), 0, st, p);
    return cudaGetLastError();
}

cudaError_t wf_launch_conv(const ConvP& p, cudaStream_t st)
{
    if (wf_slabtc_conv_ok(p)) return wf_launch_slabtc_conv(p, st);
    if (wf_slide_conv_ok(p)) return wf_launch_slide_conv(p, st);
    if (wf_thin_conv_ok(p)) return wf_launch_thin_conv(p, st);
    if (wf_group_conv_ok(p)) return wf_launch_group_conv(p, st);
    switch (conv_cfg_for(p.Cout)) {
        case CFG_BIG: return launch_conv_t<64, 128, 8, 8, 8>(p, st);
        case CFG_MID: return launch_conv_t<32, 128, 8, 4, 8>(p, st);
        case CFG_THIN16: return launch_conv_t<16, 256, 4, 8, 8>(p, st);
        default: return launch_conv_t<8, 256, 4, 8, 4>(p, st);
    }
}

template <int BM, int BN, int TM, int TN>
static cudaError_t launch_wgrad_t(WgradP p, int target_ctas, cudaStream_t st)
{
    const int tiles = ((p.Cout + BM - 1) / BM) * ((p.Cin + BN - 1) / BN);
    const int z = p.groups * p.ntaps;
    const long long total8 = ((long long)p.Pout * p.N + 7) / 8;
    long long kc = target_ctas / ((long long)tiles * z);
    if (kc < 1) kc = 1;
    if (kc > total8 / 16) kc = total8 / 16;          // at least 16 K-steps per block
    if (kc < 1) kc = 1;
    p.kchunks = (int)kc;
    dim3 grid((unsigned)kc, tiles, z);
    wf_launch_pdl(conv_wgrad_kernel<BM, BN, TM, TN>, dim3(grid), dim3((BM / TM) * (BN / TN)), 0, st, p);
    return cudaGetLastError();
}

cudaError_t wf_launch_wgrad(const WgradP& p, int num_sms, cudaStream_t st)
{
    if (wf_slabtc_wgrad_ok(p)) return wf_launch_slabtc_wgrad(p, st);
    if (wf_slide_wgrad_ok(p)) return wf_launch_slide_wgrad(p, num_sms, st);
    if (wf_thin_wgrad_ok(p)) return wf_launch_thin_wgrad(p, num_sms, st);
    if (wf_group_wgrad_ok(p)) return wf_launch_group_wgrad(p, num_sms, st);
    const int target = num_sms * 6;
    if (p.Cout >= 48 && p.Cin >= 48) return launch_wgrad_t<64, 64, 8, 4>(p, target, st);
    if (p.Cout >= 48) return launch_wgrad_t<64, 32, 4, 4>(p, target, st);
    if (p.Cin >= 48) return launch_wgrad_t<32, 64, 4, 4>(p, target, st);
    return launch_wgrad_t<32, 32, 4, 4>(p, target, st);
}
