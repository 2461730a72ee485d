// Sliding-window tensor-core kernels for the position-tap convolutions of the conv stack and the decoder:
//   * (1x3) convs, stride (1,1) / (1,2), and their 1x1 strided shortcuts (models/convnet.py:11-12,17,22,27,48,53,58,63)
//   * decoder 3x3 conv (models/pose_model.py:45)
// forward, backward-data (one kernel) and backward-weights (second kernel).
//
// In the internal layout [channel][position][n] a tap along the feature axis is a whole-row ("slab") shift: output position
// `opos` reads the input slabs ipos = (opos*pmul + dp[tap]) / pdiv.  A CTA owns a tile of columns and walks a range of
// output positions; the [Cin][columns] slab of every input position it needs is staged in shared memory exactly ONCE
// (BatchNorm+SiLU+Dropout2d / BatchNorm-backward applied on the way in) and stays in a small ring while the window slides,
// so each activation is read from HBM and transformed once per layer instead of once per tap.  The contraction over
// channels runs on the warp-level tensor-core path (mma.sync m16n8k8 tf32, fp32 accumulate) with the 3xTF32 split
// a*b ~ a_hi*b_hi + (a_lo*b_hi + a_hi*b_lo) done in registers when a fragment is loaded, which keeps fp32 parity
// (1e-4 on outputs) while the FP32 pipe is left to the prologue / epilogue arithmetic.
//
// Everything that does not depend on the position is computed once per thread before the walk: which (channel, column quad)
// of a slab a thread loads, where it lands in the ring, its dropout-mask value, where its epilogue quads go.  A step of the
// walk covers PS output positions: the loads of the next step's slabs are issued first, then the MMAs + epilogues of this
// step's positions run out of the ring, then the loaded values are transformed and stored, then ONE barrier.
#include <climits>
#include <cstdlib>
#include "wf_common.cuh"
#include "wf_elem.h"

namespace {

constexpr int SL_H = 4;                              // halo columns on each side of a staged slab (time taps of the 3x3 conv)

struct SlideGeo {
    int R;                 // ring slots
    int PC, PS;            // output positions per CTA / per step (PC is a multiple of PS)
    int two_sync;          // ring too small to land the next step while this one is read: extra barrier before the stores
    int dpmin, dpmax;      // range of the position taps
    int cap;               // slabs a step may load (register capacity of the loader)
    // backward-weights only
    int MWs, NWs, KWs, nitems, ncol;
};

__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo)
{
    hi = (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u;          // round to nearest tf32
    lo = __float_as_uint(x - __uint_as_float(hi));              // the tensor core drops the low 13 bits of the remainder
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2])
{
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// input slabs touched by output positions [a, b): a contiguous superset [lo, hi] (hi < lo: none); pdiv is 1 or 2
__host__ __device__ __forceinline__ void step_range(int pmul, int pdiv, int Pin, int dpmin, int dpmax, int a, int b, int& lo, int& hi)
{
    lo = a * pmul + dpmin;
    hi = (b - 1) * pmul + dpmax;
    if (pdiv == 2) { lo = (lo + 1) >> 1; hi = hi >> 1; }       // ceil / floor of a signed value (arithmetic shift)
    if (lo < 0) lo = 0;
    if (hi > Pin - 1) hi = Pin - 1;
}
// ring slot of input slab `ipos` given that slab `base_ipos` sits in slot `base_slot` (ipos - base_ipos < R)
__device__ __forceinline__ int ring_slot(int ipos, int base_ipos, int base_slot, int R)
{
    int s = base_slot + (ipos - base_ipos);
    return s >= R ? s - R : s;
}

// prologue of one float4 of an operand (DESIGN.md "Why split at BatchNorm"); c4 = (a, b, c, d) of the channel
template <int PRO>
__device__ __forceinline__ float4 apply_pro(float4 v, float4 v2, float mk, float4 c4)
{
    if (PRO == PRO_BNSILU) {
        v.x = wf_silu(fmaf(c4.x, v.x - c4.w, c4.y)) * mk; v.y = wf_silu(fmaf(c4.x, v.y - c4.w, c4.y)) * mk;
        v.z = wf_silu(fmaf(c4.x, v.z - c4.w, c4.y)) * mk; v.w = wf_silu(fmaf(c4.x, v.w - c4.w, c4.y)) * mk;
    } else if (PRO == PRO_AFFINE) {
        v.x = fmaf(c4.x, v.x - c4.w, c4.y); v.y = fmaf(c4.x, v.y - c4.w, c4.y);
        v.z = fmaf(c4.x, v.z - c4.w, c4.y); v.w = fmaf(c4.x, v.w - c4.w, c4.y);
    } else if (PRO == PRO_BNBWD) {
        v.x = fmaf(c4.x, v.x, fmaf(c4.y, v2.x - c4.w, c4.z)); v.y = fmaf(c4.x, v.y, fmaf(c4.y, v2.y - c4.w, c4.z));
        v.z = fmaf(c4.x, v.z, fmaf(c4.y, v2.z - c4.w, c4.z)); v.w = fmaf(c4.x, v.w, fmaf(c4.y, v2.w - c4.w, c4.z));
    }
    return v;
}

// loader descriptor of one float4 item: bits 0..15 word offset inside a slab, 16..19 slab index inside the step, 20..27 channel
__device__ __forceinline__ int desc_off(int d) { return d & 0xFFFF; }
__device__ __forceinline__ int desc_slab(int d) { return (d >> 16) & 15; }
__device__ __forceinline__ int desc_chan(int d) { return (d >> 20) & 255; }

// =========================================================================================================
// forward / backward-data
//   D[m][opos][n] = sum_tap sum_k W[tap][k][m] * X'[k][ipos(opos,tap)][n + dn(tap)]
// warp grid MW x NW, warp tile (16*MI) x (8*NI); BM = 16*MI*MW covers all output channels, BN = 8*NI*NW columns.
// LD = float4 registers per thread (per tensor) for the slabs in flight; PRO = prologue mode; HASDN = some tap shifts time.
// =========================================================================================================
template <int MI, int NI, int MW, int NW, int LD, int MINB, int PRO, bool HASDN>
__global__ void __launch_bounds__(32 * MW * NW, MINB) slide_conv_kernel(const ConvP p, const SlideGeo g)
{
    wf_pdl_enter();
    constexpr int NT = 32 * MW * NW, BM = 16 * MI * MW, BN = 8 * NI * NW, XS = BN + 2 * SL_H, Q = XS / 4, WS = BM + 8;
    extern __shared__ __align__(16) float smem[];
    __shared__ double red[2][BM];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp % MW, wn = warp / MW;
    const int fr = lane >> 2, fc = lane & 3;
    const int K8 = (p.Cin + 7) & ~7;
    const int R = g.R;
    float* Ws = smem;                                   // [ntaps][K8][WS]
    float* ring = smem + p.ntaps * K8 * WS;             // R x [K8][XS]
    float* coef = ring + R * K8 * XS;                   // [K8] x (a, b, c, d)
    const int slab_words = K8 * XS;
    const int n0 = blockIdx.x * BN;
    const int op0 = blockIdx.y * g.PC, op1 = min(op0 + g.PC, p.Pout);
    const int per = p.Cin * Q;                          // float4 items per slab

    // ---- weights of all taps (rows beyond Cin / columns beyond Cout zero), prologue coefficients, ring padding rows ----
    for (int idx = tid; idx < p.ntaps * K8 * (BM / 4); idx += NT) {
        const int mq = idx % (BM / 4), k = (idx / (BM / 4)) % K8, tap = idx / ((BM / 4) * K8);
        float4 v = f4zero();
        if (k < p.Kpad && mq * 4 < p.Mpad) v = ld4(p.w + ((size_t)tap * p.Kpad + k) * p.Mpad + mq * 4);
        st4(Ws + (tap * K8 + k) * WS + mq * 4, v);
    }
    if (PRO != PRO_NONE)
        for (int k = tid; k < p.Cin; k += NT)
            st4(coef + 4 * k, make_float4(p.pro_a[k], p.pro_b[k], PRO == PRO_BNBWD ? p.pro_c[k] : 0.f, p.pro_d[k]));
    {
        const int padw = (K8 - p.Cin) * XS;
        for (int idx = tid; idx < R * padw; idx += NT) ring[(idx / padw) * slab_words + p.Cin * XS + idx % padw] = 0.f;
    }
    if (tid < 2 * BM) red[tid / BM][tid % BM] = 0.0;
    __syncthreads();                                    // the coefficient table is read by the stores that prime the ring

    // ---- loader descriptors (position independent) ----
    int goff[LD], sd[LD];
    float mk[LD];
    float4 rv[LD], rv2[PRO == PRO_BNBWD ? LD : 1];
#pragma unroll
    for (int i = 0; i < LD; ++i) {
        const int idx = tid + i * NT;
        goff[i] = -1; sd[i] = -1; mk[i] = 1.f;
        if (idx < g.cap * per) {
            const int s = idx / per, rem = idx - s * per, k = rem / Q, q = rem - k * Q;
            sd[i] = (k * XS + 4 * q) | (s << 16) | (k << 20);
            const int nn = n0 - SL_H + 4 * q;
            if (nn >= 0 && nn < p.N) {
                const int b = nn / WF_T, t = nn - b * WF_T;
                goff[i] = (int)((long long)k * p.in_sc + (long long)s * p.in_sp + (long long)b * p.in_sb + t);
                if (PRO == PRO_BNSILU && p.mask) mk[i] = p.mask[(long long)b * p.m_sb + (long long)k * p.m_sc];
            }
        }
    }
    // global -> registers for slabs [a, a+cnt)
    const int in_sp = (int)p.in_sp;
    int base_ipos = 0, base_slot = 0;                   // ring bookkeeping: slab base_ipos sits in slot base_slot
    auto load_slabs = [&](int a, int cnt) {
        const int aoff = a * in_sp;
#pragma unroll
        for (int i = 0; i < LD; ++i) {
            rv[i] = f4zero();
            if (PRO == PRO_BNBWD) rv2[PRO == PRO_BNBWD ? i : 0] = f4zero();
            if (goff[i] >= 0 && desc_slab(sd[i]) < cnt) {
                rv[i] = ld4(p.in + (goff[i] + aoff));
                if (PRO == PRO_BNBWD) rv2[PRO == PRO_BNBWD ? i : 0] = ld4(p.in2 + (goff[i] + aoff));
            }
        }
    };
    // registers -> prologue -> ring
    auto store_slabs = [&](int a, int cnt) {
        const int aslot = ring_slot(a, base_ipos, base_slot, R);
#pragma unroll
        for (int i = 0; i < LD; ++i) {
            if (sd[i] >= 0 && desc_slab(sd[i]) < cnt) {
                float4 v = rv[i];
                if (PRO != PRO_NONE && goff[i] >= 0) v = apply_pro<PRO>(v, rv2[PRO == PRO_BNBWD ? i : 0], mk[i], ld4(coef + 4 * desc_chan(sd[i])));
                int slot = aslot + desc_slab(sd[i]);
                if (slot >= R) slot -= R;
                st4(ring + slot * slab_words + desc_off(sd[i]), v);
            }
        }
    };

    // ---- epilogue descriptors ----
    const bool odd = fc & 1;
    int eoff[NI], tcol[NI];
    int rowoff[MI];
    bool mv[MI];
    float emk[MI][NI];
    double sd0[MI], sd1[MI];
    float bias[MI], es[MI], et[MI], em[MI];
#pragma unroll
    for (int mi = 0; mi < MI; ++mi) {
        sd0[mi] = 0.0; sd1[mi] = 0.0;
        const int m = (wm * MI + mi) * 16 + fr + (odd ? 8 : 0);
        mv[mi] = m < p.Cout;
        rowoff[mi] = (int)((long long)m * p.out_sc);
        bias[mi] = 0.f; es[mi] = 0.f; et[mi] = 0.f; em[mi] = 0.f;
        if (mv[mi]) {
            if (p.bias) bias[mi] = p.bias[m];
            if (p.epi_mode == EPI_DSILU) { es[mi] = p.e_scale[m]; et[mi] = p.e_shift[m]; }
            if (p.epi_mode == EPI_DSILU || p.epi_mode == EPI_DAFF) em[mi] = p.e_mean[m];
        }
    }
#pragma unroll
    for (int ni = 0; ni < NI; ++ni) {
        tcol[ni] = (n0 + (wn * NI + ni) * 8 + fr) % WF_T;
        const int n = n0 + (wn * NI + ni) * 8 + (fc >> 1) * 4;
        const int b = n / WF_T, t = n - b * WF_T;
        eoff[ni] = n < p.N ? (int)((long long)b * p.out_sb + t) : -1;
#pragma unroll
        for (int mi = 0; mi < MI; ++mi) {
            emk[mi][ni] = 1.f;
            if (p.epi_mode == EPI_DSILU && p.emask && mv[mi] && n < p.N)
                emk[mi][ni] = p.emask[(long long)b * p.em_sb + (long long)((wm * MI + mi) * 16 + fr + (odd ? 8 : 0)) * p.em_sc];
        }
    }
    const int epi = p.epi_mode;
    const bool accum = p.accumulate != 0;
    const uint32_t* wsu = reinterpret_cast<const uint32_t*>(Ws);

    // ---- one output position: taps out of the ring, then the epilogue ----
    auto do_pos = [&](int opos) {
        const int pbase = (int)((long long)opos * p.out_sp);
        const bool need_raw = epi == EPI_DSILU || epi == EPI_DAFF;
        float acc[MI][NI][4], cor[MI][NI][4];
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int j = 0; j < NI; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e) { acc[i][j][e] = 0.f; cor[i][j][e] = 0.f; }
        for (int tap = 0; tap < p.ntaps; ++tap) {
            int ipos = opos * p.pmul + p.dp[tap];
            if (ipos < 0) continue;
            if (p.pdiv == 2) { if (ipos & 1) continue; ipos >>= 1; }
            if (ipos >= p.Pin) continue;
            const int dn = HASDN ? p.dn[tap] : 0;
            bool ok[NI];
#pragma unroll
            for (int ni = 0; ni < NI; ++ni) ok[ni] = !HASDN || (tcol[ni] + dn >= 0 && tcol[ni] + dn < WF_T);
            const float* xs = ring + ring_slot(ipos, base_ipos, base_slot, R) * slab_words + fc * XS + SL_H + wn * NI * 8 + fr + dn;
            const uint32_t* wt = wsu + (tap * K8 + fc) * WS + wm * MI * 16 + fr;
            for (int k8 = 0; k8 < K8; k8 += 8, xs += 8 * XS, wt += 8 * WS) {
                uint32_t ah[MI][4], al[MI][4], bh[NI][2], bl[NI][2];
#pragma unroll
                for (int mi = 0; mi < MI; ++mi) {
                    split_tf32(__uint_as_float(wt[mi * 16]), ah[mi][0], al[mi][0]);
                    split_tf32(__uint_as_float(wt[mi * 16 + 8]), ah[mi][1], al[mi][1]);
                    split_tf32(__uint_as_float(wt[mi * 16 + 4 * WS]), ah[mi][2], al[mi][2]);
                    split_tf32(__uint_as_float(wt[mi * 16 + 4 * WS + 8]), ah[mi][3], al[mi][3]);
                }
#pragma unroll
                for (int ni = 0; ni < NI; ++ni) {
                    float v0 = xs[ni * 8], v1 = xs[ni * 8 + 4 * XS];
                    if (HASDN && !ok[ni]) { v0 = 0.f; v1 = 0.f; }
                    split_tf32(v0, bh[ni][0], bl[ni][0]);
                    split_tf32(v1, bh[ni][1], bl[ni][1]);
                }
                // three passes so that no MMA waits on the one issued just before it (the two correction products share an accumulator)
#pragma unroll
                for (int mi = 0; mi < MI; ++mi)
#pragma unroll
                    for (int ni = 0; ni < NI; ++ni) mma_tf32(cor[mi][ni], al[mi], bh[ni]);
#pragma unroll
                for (int mi = 0; mi < MI; ++mi)
#pragma unroll
                    for (int ni = 0; ni < NI; ++ni) mma_tf32(acc[mi][ni], ah[mi], bh[ni]);
#pragma unroll
                for (int mi = 0; mi < MI; ++mi)
#pragma unroll
                    for (int ni = 0; ni < NI; ++ni) mma_tf32(cor[mi][ni], ah[mi], bl[ni]);
            }
        }
        // lanes fc and fc^1 swap halves so each owns 4 consecutive columns of one output row
#pragma unroll
        for (int mi = 0; mi < MI; ++mi) {
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int ni = 0; ni < NI; ++ni) {
                const float d0 = acc[mi][ni][0] + cor[mi][ni][0], d1 = acc[mi][ni][1] + cor[mi][ni][1];
                const float d2 = acc[mi][ni][2] + cor[mi][ni][2], d3 = acc[mi][ni][3] + cor[mi][ni][3];
                const float sx = odd ? d0 : d2, sy = odd ? d1 : d3;
                const float rx = __shfl_xor_sync(0xffffffffu, sx, 1), ry = __shfl_xor_sync(0xffffffffu, sy, 1);
                float v[4];
                if (odd) { v[0] = rx; v[1] = ry; v[2] = d2; v[3] = d3; }
                else { v[0] = d0; v[1] = d1; v[2] = rx; v[3] = ry; }
                if (mv[mi] && eoff[ni] >= 0) {
                    const int off = rowoff[mi] + pbase + eoff[ni];
#pragma unroll
                    for (int j = 0; j < 4; ++j) v[j] += bias[mi];
                    if (accum) { const float4 o = ld4(p.out + off); v[0] += o.x; v[1] += o.y; v[2] += o.z; v[3] += o.w; }
                    if (epi == EPI_STATS) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) { s0 += v[j]; s1 = fmaf(v[j], v[j], s1); }
                    } else if (need_raw) {
                        const float4 r4 = ld4(p.eraw + off);
                        const float r[4] = {r4.x - em[mi], r4.y - em[mi], r4.z - em[mi], r4.w - em[mi]};
                        if (epi == EPI_DSILU) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) v[j] = v[j] * emk[mi][ni] * wf_dsilu(fmaf(es[mi], r[j], et[mi]));
                        }
#pragma unroll
                        for (int j = 0; j < 4; ++j) { s0 += v[j]; s1 = fmaf(v[j], r[j], s1); }
                    }
                    st4(p.out + off, make_float4(v[0], v[1], v[2], v[3]));
                }
            }
            sd0[mi] += (double)s0; sd1[mi] += (double)s1;
        }
    };

    // ---- prime the ring with the slabs of the first step ----
    int staged_hi = -1;
    {
        int lo, hi;
        step_range(p.pmul, p.pdiv, p.Pin, g.dpmin, g.dpmax, op0, min(op0 + g.PS, op1), lo, hi);
        base_ipos = lo;
        for (int a = lo; a <= hi; a += g.cap) {
            const int cnt = min(g.cap, hi - a + 1);
            load_slabs(a, cnt);
            store_slabs(a, cnt);
        }
        if (hi >= lo) staged_hi = hi;
    }
    __syncthreads();

    for (int op = op0; op < op1; op += g.PS) {
        const int oe = min(op + g.PS, op1);
        {       // slide the ring origin to the first slab this step reads (slabs below it are dead)
            int lo, hi;
            step_range(p.pmul, p.pdiv, p.Pin, g.dpmin, g.dpmax, op, oe, lo, hi);
            if (hi >= lo && lo > base_ipos) { base_slot = ring_slot(lo, base_ipos, base_slot, R); base_ipos = lo; }
        }
        // ---- issue the loads of the slabs the next step adds ----
        int na = 0, ncnt = 0;
        if (oe < op1) {
            int nlo, nhi;
            step_range(p.pmul, p.pdiv, p.Pin, g.dpmin, g.dpmax, oe, min(oe + g.PS, op1), nlo, nhi);
            na = max(staged_hi + 1, nlo);
            ncnt = max(nhi - na + 1, 0);
        }
        if (ncnt) load_slabs(na, ncnt);
        for (int opos = op; opos < oe; ++opos) do_pos(opos);
        if (g.two_sync) __syncthreads();
        if (ncnt) { store_slabs(na, ncnt); staged_hi = na + ncnt - 1; }
        __syncthreads();
    }

    // ---- BatchNorm sums: lanes fc, fc^2 share a row; warps along N meet in shared memory; one atomic pair per row per CTA ----
    if (epi != EPI_STORE && p.stat0 != nullptr) {
#pragma unroll
        for (int mi = 0; mi < MI; ++mi) {
            double a = sd0[mi], b = sd1[mi];
            a += __shfl_xor_sync(0xffffffffu, a, 2);
            b += __shfl_xor_sync(0xffffffffu, b, 2);
            const int ml = (wm * MI + mi) * 16 + fr + (odd ? 8 : 0);
            if (fc < 2) { atomicAdd(&red[0][ml], a); atomicAdd(&red[1][ml], b); }
        }
        __syncthreads();
        if (tid < BM && tid < p.Cout) {
            atomicAdd(p.stat0 + tid, red[0][tid]);
            atomicAdd(p.stat1 + tid, red[1][tid]);
        }
    }
    wf_bn_tail(p.tail);
}

// =========================================================================================================
// forward / backward-data of the thin layers (<= 8 output channels, <= 16 input channels: ConvBlock1 `up`, residual block 0,
// the input gradients of residual block 1).  Same sliding window, but the MMA operand roles are swapped: the COLUMNS are the
// M dimension (16 per tile) and the <= 8 output channels the N dimension, so no half of a 16-row tile is padding and the
// 3 x Cin x 8 weights live in registers (split once per CTA) instead of shared memory:
//   D[col][co] = sum_tap sum_k X'[k][ipos(opos,tap)][col] * W[tap][k][co]
// A warp owns MI tiles of 16 columns; accumulator lane layout: columns fr / fr+8, channels 2fc / 2fc+1.
// =========================================================================================================
template <int MI, int NTAPS, int K8S, int LD, int MINB, int PRO>
__global__ void __launch_bounds__(256, MINB) slide_thin_kernel(const ConvP p, const SlideGeo g)
{
    wf_pdl_enter();
    constexpr int NT = 256, NW = 8, BN = 16 * MI * NW, XS = BN + 2 * SL_H, Q = XS / 4, K8 = 8 * K8S;
    extern __shared__ __align__(16) float smem[];
    __shared__ double red[2][8];
    const int tid = threadIdx.x, lane = tid & 31, wn = tid >> 5;
    const int fr = lane >> 2, fc = lane & 3;
    const int R = g.R;
    float* ring = smem;                                 // R x [K8][XS]
    float* coef = ring + R * K8 * XS;                   // [K8] x (a, b, c, d)
    constexpr int slab_words = K8 * XS;
    const int n0 = blockIdx.x * BN;
    const int op0 = blockIdx.y * g.PC, op1 = min(op0 + g.PC, p.Pout);
    const int per = p.Cin * Q;

    // ---- weights: B fragments (k = input channel, n = output channel) of every tap, split once ----
    uint32_t bwh[NTAPS][K8S][2], bwl[NTAPS][K8S][2];
#pragma unroll
    for (int tap = 0; tap < NTAPS; ++tap)
#pragma unroll
        for (int ks = 0; ks < K8S; ++ks)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int k = ks * 8 + fc + 4 * h;
                const float w = (k < p.Kpad && fr < p.Mpad) ? p.w[((size_t)tap * p.Kpad + k) * p.Mpad + fr] : 0.f;
                split_tf32(w, bwh[tap][ks][h], bwl[tap][ks][h]);
            }
    if (PRO != PRO_NONE)
        for (int k = tid; k < p.Cin; k += NT)
            st4(coef + 4 * k, make_float4(p.pro_a[k], p.pro_b[k], PRO == PRO_BNBWD ? p.pro_c[k] : 0.f, p.pro_d[k]));
    {
        const int padw = (K8 - p.Cin) * XS;
        for (int idx = tid; idx < R * padw; idx += NT) ring[(idx / padw) * slab_words + p.Cin * XS + idx % padw] = 0.f;
    }
    if (tid < 16) red[tid / 8][tid % 8] = 0.0;
    __syncthreads();

    // ---- loader (same scheme as slide_conv_kernel) ----
    int goff[LD], sd[LD];
    float mk[LD];
    float4 rv[LD], rv2[PRO == PRO_BNBWD ? LD : 1];
#pragma unroll
    for (int i = 0; i < LD; ++i) {
        const int idx = tid + i * NT;
        goff[i] = -1; sd[i] = -1; mk[i] = 1.f;
        if (idx < g.cap * per) {
            const int s = idx / per, rem = idx - s * per, k = rem / Q, q = rem - k * Q;
            sd[i] = (k * XS + 4 * q) | (s << 16) | (k << 20);
            const int nn = n0 - SL_H + 4 * q;
            if (nn >= 0 && nn < p.N) {
                const int b = nn / WF_T, t = nn - b * WF_T;
                goff[i] = (int)((long long)k * p.in_sc + (long long)s * p.in_sp + (long long)b * p.in_sb + t);
                if (PRO == PRO_BNSILU && p.mask) mk[i] = p.mask[(long long)b * p.m_sb + (long long)k * p.m_sc];
            }
        }
    }
    const int in_sp = (int)p.in_sp;
    int base_ipos = 0, base_slot = 0;
    auto load_slabs = [&](int a, int cnt) {
        const int aoff = a * in_sp;
#pragma unroll
        for (int i = 0; i < LD; ++i) {
            rv[i] = f4zero();
            if (PRO == PRO_BNBWD) rv2[PRO == PRO_BNBWD ? i : 0] = f4zero();
            if (goff[i] >= 0 && desc_slab(sd[i]) < cnt) {
                rv[i] = ld4(p.in + (goff[i] + aoff));
                if (PRO == PRO_BNBWD) rv2[PRO == PRO_BNBWD ? i : 0] = ld4(p.in2 + (goff[i] + aoff));
            }
        }
    };
    auto store_slabs = [&](int a, int cnt) {
        const int aslot = ring_slot(a, base_ipos, base_slot, R);
#pragma unroll
        for (int i = 0; i < LD; ++i) {
            if (sd[i] >= 0 && desc_slab(sd[i]) < cnt) {
                float4 v = rv[i];
                if (PRO != PRO_NONE && goff[i] >= 0) v = apply_pro<PRO>(v, rv2[PRO == PRO_BNBWD ? i : 0], mk[i], ld4(coef + 4 * desc_chan(sd[i])));
                int slot = aslot + desc_slab(sd[i]);
                if (slot >= R) slot -= R;
                st4(ring + slot * slab_words + desc_off(sd[i]), v);
            }
        }
    };

    // ---- epilogue: the accumulator layout (columns fr / fr+8, channels 2fc / 2fc+1) is turned into 4-column quads through a
    //      per-warp scratch tile [8 channels][32 columns]; lane -> channels lane/8 and lane/8 + 4, column quad lane%8 ----
    static_assert(MI == 2, "the epilogue scratch assumes 32 columns per warp");
    constexpr int SS = 36;                               // scratch row stride: conflict-free scalar stores and 128-bit loads
    float* scr = coef + 4 * K8 + wn * 8 * SS;
    int rowoff[2];
    bool chv[2];
    float emk[2], bias[2], es[2], et[2], em[2];
    double sd0[2], sd1[2];
    const int qn = n0 + wn * 32 + (lane & 7) * 4;
    const int qb = qn / WF_T, qt = qn - qb * WF_T;
    const int ecol = qn < p.N ? (int)((long long)qb * p.out_sb + qt) : -1;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int m = (lane >> 3) + 4 * j;
        chv[j] = m < p.Cout;
        rowoff[j] = (int)((long long)m * p.out_sc);
        sd0[j] = 0.0; sd1[j] = 0.0; bias[j] = 0.f; es[j] = 0.f; et[j] = 0.f; em[j] = 0.f; emk[j] = 1.f;
        if (chv[j]) {
            if (p.bias) bias[j] = p.bias[m];
            if (p.epi_mode == EPI_DSILU) { es[j] = p.e_scale[m]; et[j] = p.e_shift[m]; }
            if (p.epi_mode == EPI_DSILU || p.epi_mode == EPI_DAFF) em[j] = p.e_mean[m];
            if (p.epi_mode == EPI_DSILU && p.emask && qn < p.N) emk[j] = p.emask[(long long)qb * p.em_sb + (long long)m * p.em_sc];
        }
    }
    const int epi = p.epi_mode;
    const bool accum = p.accumulate != 0;

    auto do_pos = [&](int opos) {
        // operands of the epilogue (raw tensor of the BatchNorm being differentiated, running input gradient) are requested
        // before the MMAs so that their latency is covered
        const int pbase = (int)((long long)opos * p.out_sp);
        const bool need_raw = epi == EPI_DSILU || epi == EPI_DAFF;
        constexpr bool PREF = PRO == PRO_BNBWD;          // backward-data launches
        float4 praw[PREF ? 2 : 1], pacc[PREF ? 2 : 1];
        if (PREF) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                praw[PREF ? j : 0] = f4zero(); pacc[PREF ? j : 0] = f4zero();
                if (chv[j] && ecol >= 0) {
                    if (need_raw) praw[PREF ? j : 0] = ld4(p.eraw + (rowoff[j] + pbase + ecol));
                    if (accum) pacc[PREF ? j : 0] = ld4(p.out + (rowoff[j] + pbase + ecol));
                }
            }
        }
        float acc[MI][4], cor[MI][4];
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int e = 0; e < 4; ++e) { acc[i][e] = 0.f; cor[i][e] = 0.f; }
#pragma unroll
        for (int tap = 0; tap < NTAPS; ++tap) {
            int ipos = opos * p.pmul + p.dp[tap];
            if (ipos < 0) continue;
            if (p.pdiv == 2) { if (ipos & 1) continue; ipos >>= 1; }
            if (ipos >= p.Pin) continue;
            const float* xs = ring + ring_slot(ipos, base_ipos, base_slot, R) * slab_words + fc * XS + SL_H + wn * MI * 16 + fr;
#pragma unroll
            for (int ks = 0; ks < K8S; ++ks) {
                uint32_t ah[MI][4], al[MI][4];
#pragma unroll
                for (int mi = 0; mi < MI; ++mi) {
                    const float* x0 = xs + ks * 8 * XS + mi * 16;
                    split_tf32(x0[0], ah[mi][0], al[mi][0]);
                    split_tf32(x0[8], ah[mi][1], al[mi][1]);
                    split_tf32(x0[4 * XS], ah[mi][2], al[mi][2]);
                    split_tf32(x0[4 * XS + 8], ah[mi][3], al[mi][3]);
                }
#pragma unroll
                for (int mi = 0; mi < MI; ++mi) mma_tf32(cor[mi], al[mi], bwh[tap][ks]);
#pragma unroll
                for (int mi = 0; mi < MI; ++mi) mma_tf32(acc[mi], ah[mi], bwh[tap][ks]);
#pragma unroll
                for (int mi = 0; mi < MI; ++mi) mma_tf32(cor[mi], ah[mi], bwl[tap][ks]);
            }
        }
        // accumulators -> scratch (element e of tile mi: column mi*16 + fr + 8*(e>>1), channel 2fc + (e&1))
        __syncwarp();
#pragma unroll
        for (int mi = 0; mi < MI; ++mi)
#pragma unroll
            for (int e = 0; e < 4; ++e) scr[(2 * fc + (e & 1)) * SS + mi * 16 + fr + 8 * (e >> 1)] = acc[mi][e] + cor[mi][e];
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            float s0 = 0.f, s1 = 0.f;
            if (chv[j] && ecol >= 0) {
                const float4 q4 = ld4(scr + ((lane >> 3) + 4 * j) * SS + (lane & 7) * 4);
                float v[4] = {q4.x + bias[j], q4.y + bias[j], q4.z + bias[j], q4.w + bias[j]};
                const int off = rowoff[j] + pbase + ecol;
                if (accum) { const float4 o = PREF ? pacc[PREF ? j : 0] : ld4(p.out + off); v[0] += o.x; v[1] += o.y; v[2] += o.z; v[3] += o.w; }
                if (epi == EPI_STATS) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) { s0 += v[e]; s1 = fmaf(v[e], v[e], s1); }
                } else if (need_raw) {
                    const float4 r4 = PREF ? praw[PREF ? j : 0] : ld4(p.eraw + off);
                    const float r[4] = {r4.x - em[j], r4.y - em[j], r4.z - em[j], r4.w - em[j]};
                    if (epi == EPI_DSILU) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) v[e] = v[e] * emk[j] * wf_dsilu(fmaf(es[j], r[e], et[j]));
                    }
#pragma unroll
                    for (int e = 0; e < 4; ++e) { s0 += v[e]; s1 = fmaf(v[e], r[e], s1); }
                }
                st4(p.out + off, make_float4(v[0], v[1], v[2], v[3]));
            }
            sd0[j] += (double)s0; sd1[j] += (double)s1;
        }
    };

    // ---- the walk (identical to slide_conv_kernel) ----
    int staged_hi = -1;
    {
        int lo, hi;
        step_range(p.pmul, p.pdiv, p.Pin, g.dpmin, g.dpmax, op0, min(op0 + g.PS, op1), lo, hi);
        base_ipos = lo;
        for (int a = lo; a <= hi; a += g.cap) {
            const int cnt = min(g.cap, hi - a + 1);
            load_slabs(a, cnt);
            store_slabs(a, cnt);
        }
        if (hi >= lo) staged_hi = hi;
    }
    __syncthreads();
    for (int op = op0; op < op1; op += g.PS) {
        const int oe = min(op + g.PS, op1);
        {
            int lo, hi;
            step_range(p.pmul, p.pdiv, p.Pin, g.dpmin, g.dpmax, op, oe, lo, hi);
            if (hi >= lo && lo > base_ipos) { base_slot = ring_slot(lo, base_ipos, base_slot, R); base_ipos = lo; }
        }
        int na = 0, ncnt = 0;
        if (oe < op1) {
            int nlo, nhi;
            step_range(p.pmul, p.pdiv, p.Pin, g.dpmin, g.dpmax, oe, min(oe + g.PS, op1), nlo, nhi);
            na = max(staged_hi + 1, nlo);
            ncnt = max(nhi - na + 1, 0);
        }
        if (ncnt) load_slabs(na, ncnt);
        for (int opos = op; opos < oe; ++opos) do_pos(opos);
        if (g.two_sync) __syncthreads();
        if (ncnt) { store_slabs(na, ncnt); staged_hi = na + ncnt - 1; }
        __syncthreads();
    }

    // ---- BatchNorm sums: the 8 lanes with equal lane/8 hold the same two channels ----
    if (epi != EPI_STORE && p.stat0 != nullptr) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            double a = sd0[j], b = sd1[j];
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
            if ((lane & 7) == 0) { atomicAdd(&red[0][(lane >> 3) + 4 * j], a); atomicAdd(&red[1][(lane >> 3) + 4 * j], b); }
        }
        __syncthreads();
        if (tid < 8 && tid < p.Cout) {
            atomicAdd(p.stat0 + tid, red[0][tid]);
            atomicAdd(p.stat1 + tid, red[1][tid]);
        }
    }
    wf_bn_tail(p.tail);
}

// =========================================================================================================
// backward-weights
//   dW[co][ci][tap] = sum_{opos,n} G'[co][opos][n] * X'[ci][ipos(opos,tap)][n + dn(tap)]
// G' = BatchNorm-backward of (dy, raw), X' = prologue of the layer input.  Persistent CTAs walk (column tile, position range)
// work items; per step the G' slabs of PS output positions (double buffered) and the new X' slabs are staged once and all taps
// are contracted over the slabs' 64 columns.  Warp grid (runtime) MWs x NWs x KWs: a warp owns MTW x NTW x NTAPS accumulator
// tiles and every KWs-th k8 step; the sums leave as fp32 atomics once per CTA.
// =========================================================================================================
constexpr int SW_BN = 64, SW_GS = SW_BN + 4, SW_XS = SW_BN + 2 * SL_H + 4, SW_GQ = SW_BN / 4, SW_XQ = (SW_BN + 2 * SL_H) / 4;
constexpr int SW_NT = 256;
constexpr int SW_UNROLL = 2;          // k8 steps of one position unrolled together in slide_wgrad_kernel

template <int MTW, int NTW, int NTAPS, int LDG, int LDX, int MINB, int PRO>
__global__ void __launch_bounds__(SW_NT, MINB) slide_wgrad_kernel(const WgradP p, const SlideGeo g)
{
    wf_pdl_enter();
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int fr = lane >> 2, fc = lane & 3;
    const int MWs = g.MWs, NWs = g.NWs, KWs = g.KWs, R = g.R, PS = g.PS;
    const int wm = warp % MWs, wn = (warp / MWs) % NWs, wk = warp / (MWs * NWs);
    const int BM = 16 * MTW * MWs, BCI = 8 * NTW * NWs;          // padded Cout / Cin
    const int gslab = BM * SW_GS, xslab = BCI * SW_XS;
    float* Gs = smem;                                            // 2 x PS x [BM][SW_GS]
    float* ring = Gs + 2 * PS * gslab;                           // R x [BCI][SW_XS]
    float* gcoef = ring + (R + 1) * xslab;                       // slab R of the ring is never written: all zeros; then [BM] x (a, b, c, d)
    float* xcoef = gcoef + 4 * BM;                               // [BCI] x (a, b, 0, d)
    const int gper = p.Cout * SW_GQ, xper = p.Cin * SW_XQ;
    constexpr bool HASDN = NTAPS == 9;
    constexpr int TG = NTAPS < 3 ? NTAPS : 3;                    // taps whose MMAs are interleaved

    // zero everything once: padding rows / columns stay zero
    for (int idx = tid; idx < 2 * PS * gslab + (R + 1) * xslab + 4 * BM + 4 * BCI; idx += SW_NT) smem[idx] = 0.f;
    __syncthreads();
    if (p.g_pro == PRO_BNBWD)
        for (int c = tid; c < p.Cout; c += SW_NT) st4(gcoef + 4 * c, make_float4(p.g_a[c], p.g_b[c], p.g_c[c], p.g_d[c]));
    if (PRO != PRO_NONE)
        for (int c = tid; c < p.Cin; c += SW_NT) st4(xcoef + 4 * c, make_float4(p.pro_a[c], p.pro_b[c], 0.f, p.pro_d[c]));
    __syncthreads();                                             // coefficient tables are read by the stores that prime the buffers

    float acc[NTAPS][MTW][NTW][4];
#pragma unroll
    for (int t = 0; t < NTAPS; ++t)
#pragma unroll
        for (int i = 0; i < MTW; ++i)
#pragma unroll
            for (int j = 0; j < NTW; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[t][i][j][e] = 0.f;

    float4 rg[LDG], rg2[LDG], rx[LDX];
    int ggo[LDG], gsd[LDG], xgo[LDX], xsd[LDX];
    float xmk[LDX];
    const bool gbn = p.g_pro == PRO_BNBWD;

    // descriptors depend on the column tile only
    auto make_desc = [&](int n0) {
#pragma unroll
        for (int i = 0; i < LDG; ++i) {
            const int idx = tid + i * SW_NT;
            ggo[i] = -1; gsd[i] = -1;
            if (idx < PS * gper) {
                const int s = idx / gper, rem = idx - s * gper, co = rem / SW_GQ, q = rem - co * SW_GQ, nn = n0 + 4 * q;
                gsd[i] = (co * SW_GS + 4 * q) | (s << 16) | (co << 20);
                if (nn < p.N) ggo[i] = (int)(((long long)co * p.Pout + s) * p.N + nn);
            }
        }
#pragma unroll
        for (int i = 0; i < LDX; ++i) {
            const int idx = tid + i * SW_NT;
            xgo[i] = -1; xsd[i] = -1; xmk[i] = 1.f;
            if (idx < g.cap * xper) {
                const int s = idx / xper, rem = idx - s * xper, k = rem / SW_XQ, q = rem - k * SW_XQ, nn = n0 - SL_H + 4 * q;
                xsd[i] = (k * SW_XS + 4 * q) | (s << 16) | (k << 20);
                if (nn >= 0 && nn < p.N) {
                    const int b = nn / WF_T, t = nn - b * WF_T;
                    xgo[i] = (int)((long long)k * p.in_sc + (long long)s * p.in_sp + (long long)b * p.in_sb + t);
                    if (PRO == PRO_BNSILU && p.mask) xmk[i] = p.mask[(long long)b * p.m_sb + (long long)k * p.m_sc];
                }
            }
        }
    };
    auto load_g = [&](int opos, int cnt) {
        const int ooff = opos * p.N;
#pragma unroll
        for (int i = 0; i < LDG; ++i) {
            rg[i] = f4zero(); rg2[i] = f4zero();
            if (ggo[i] >= 0 && desc_slab(gsd[i]) < cnt) {
                rg[i] = ld4(p.g + (ggo[i] + ooff));
                if (gbn) rg2[i] = ld4(p.g2 + (ggo[i] + ooff));
            }
        }
    };
    auto store_g = [&](int buf, int cnt) {
#pragma unroll
        for (int i = 0; i < LDG; ++i) {
            if (gsd[i] >= 0 && desc_slab(gsd[i]) < cnt) {
                float4 v = rg[i];
                if (gbn && ggo[i] >= 0) v = apply_pro<PRO_BNBWD>(v, rg2[i], 1.f, ld4(gcoef + 4 * desc_chan(gsd[i])));
                st4(Gs + (buf * PS + desc_slab(gsd[i])) * gslab + desc_off(gsd[i]), v);
            }
        }
    };
    const int in_sp = (int)p.in_sp;
    int base_ipos = 0, base_slot = 0;
    auto load_x = [&](int a, int cnt) {
        const int aoff = a * in_sp;
#pragma unroll
        for (int i = 0; i < LDX; ++i) {
            rx[i] = f4zero();
            if (xgo[i] >= 0 && desc_slab(xsd[i]) < cnt) rx[i] = ld4(p.in + (xgo[i] + aoff));
        }
    };
    auto store_x = [&](int a, int cnt) {
        const int aslot = ring_slot(a, base_ipos, base_slot, R);
#pragma unroll
        for (int i = 0; i < LDX; ++i) {
            if (xsd[i] >= 0 && desc_slab(xsd[i]) < cnt) {
                float4 v = rx[i];
                if (PRO != PRO_NONE && xgo[i] >= 0) v = apply_pro<PRO>(v, v, xmk[i], ld4(xcoef + 4 * desc_chan(xsd[i])));
                int slot = aslot + desc_slab(xsd[i]);
                if (slot >= R) slot -= R;
                st4(ring + slot * xslab + desc_off(xsd[i]), v);
            }
        }
    };

    int cur_ct = -1;
    for (int item = blockIdx.x; item < g.nitems; item += gridDim.x) {
        const int ct = item % g.ncol, pr = item / g.ncol;
        const int n0 = ct * SW_BN;
        if (ct != cur_ct) { make_desc(n0); cur_ct = ct; }
        const int op0 = pr * g.PC, op1 = min(op0 + g.PC, p.Pout);
        const int tb = (n0 + fc) % WF_T;

        // ---- prime: G' of the first step, X' slabs it reads ----
        int staged_hi = -1;
        {
            const int oe = min(op0 + PS, op1);
            load_g(op0, oe - op0); store_g(0, oe - op0);
            int lo, hi;
            step_range(p.pmul, 1, p.Pin, g.dpmin, g.dpmax, op0, oe, lo, hi);
            base_ipos = lo; base_slot = 0;
            for (int a = lo; a <= hi; a += g.cap) {
                const int cnt = min(g.cap, hi - a + 1);
                load_x(a, cnt);
                store_x(a, cnt);
            }
            if (hi >= lo) staged_hi = hi;
        }
        __syncthreads();

        int cur = 0;
        for (int op = op0; op < op1; op += PS) {
            const int oe = min(op + PS, op1);
            {
                int lo, hi;
                step_range(p.pmul, 1, p.Pin, g.dpmin, g.dpmax, op, oe, lo, hi);
                if (hi >= lo && lo > base_ipos) { base_slot = ring_slot(lo, base_ipos, base_slot, R); base_ipos = lo; }
            }
            int na = 0, ncnt = 0, gcnt = 0;
            if (oe < op1) {
                const int ne = min(oe + PS, op1);
                gcnt = ne - oe;
                load_g(oe, gcnt);
                int nlo, nhi;
                step_range(p.pmul, 1, p.Pin, g.dpmin, g.dpmax, oe, ne, nlo, nhi);
                na = max(staged_hi + 1, nlo);
                ncnt = max(nhi - na + 1, 0);
                if (ncnt) load_x(na, ncnt);
            }

            // K work of this step = (position, k8 step) units, dealt to the KWs k-slices in contiguous runs so that the per-position
            // tap setup is shared by as many k8 steps as possible; taps that leave the tensor read the all-zero slab instead of
            // being skipped, which keeps the MMA sequence free of divergence bookkeeping
            const int nunits = (oe - op) * (SW_BN / 8);
            const int upw = (nunits + KWs - 1) / KWs;
            int xb[NTAPS];
            const float* gs = nullptr;
            const int u0 = wk * upw, u1 = min((wk + 1) * upw, nunits);
            for (int pi = u0 >> 3; u0 < u1 && pi <= ((u1 - 1) >> 3); ++pi) {
                const int k_lo = pi == (u0 >> 3) ? (u0 & 7) : 0, k_hi = pi == ((u1 - 1) >> 3) ? ((u1 - 1) & 7) + 1 : 8;
                {
                    const int opos = op + pi;
                    gs = Gs + (cur * PS + pi) * gslab + (wm * MTW * 16 + fr) * SW_GS + fc;
#pragma unroll
                    for (int tap = 0; tap < NTAPS; ++tap) {
                        const int ipos = opos * p.pmul + p.dp[tap];
                        const int slot = (ipos >= 0 && ipos < p.Pin) ? ring_slot(ipos, base_ipos, base_slot, R) : R;      // slot R: zeros
                        xb[tap] = slot * xslab + (wn * NTW * 8 + fr) * SW_XS + SL_H + fc + (HASDN ? p.dn[tap] : 0);
                    }
                }
                // the k8 steps of one position share all addressing: a plain counted loop the compiler can software-pipeline
                // (fragment loads of step k+1 under the MMAs of step k)
#pragma unroll(SW_UNROLL)
                for (int k8i = k_lo; k8i < k_hi; ++k8i) {
                const int kc = k8i * 8;
                uint32_t ah[MTW][4], al[MTW][4];
#pragma unroll
                for (int mi = 0; mi < MTW; ++mi) {
                    const float* g0 = gs + mi * 16 * SW_GS + kc;
                    split_tf32(g0[0], ah[mi][0], al[mi][0]);
                    split_tf32(g0[8 * SW_GS], ah[mi][1], al[mi][1]);
                    split_tf32(g0[4], ah[mi][2], al[mi][2]);
                    split_tf32(g0[8 * SW_GS + 4], ah[mi][3], al[mi][3]);
                }
                int t0 = 0, t1 = 0;
                if (HASDN) { t0 = (tb + kc) % WF_T; t1 = t0 + 4 >= WF_T ? t0 + 4 - WF_T : t0 + 4; }
#pragma unroll
                for (int tg = 0; tg < NTAPS; tg += TG) {
                    uint32_t bh[TG][NTW][2], bl[TG][NTW][2];
#pragma unroll
                    for (int tt = 0; tt < TG; ++tt) {
                        const int tap = tg + tt;
                        const int dn = HASDN ? p.dn[tap] : 0;
                        // the window rule of a time tap zeroes columns whose source leaves the 20-step window
                        const bool v0 = !HASDN || ((t0 + dn >= 0) && (t0 + dn < WF_T)), v1 = !HASDN || ((t1 + dn >= 0) && (t1 + dn < WF_T));
                        const float* xs = ring + xb[tap] + kc;
#pragma unroll
                        for (int ni = 0; ni < NTW; ++ni) {
                            float x0 = xs[ni * 8 * SW_XS], x1 = xs[ni * 8 * SW_XS + 4];
                            if (HASDN) { x0 = v0 ? x0 : 0.f; x1 = v1 ? x1 : 0.f; }
                            split_tf32(x0, bh[tt][ni][0], bl[tt][ni][0]);
                            split_tf32(x1, bh[tt][ni][1], bl[tt][ni][1]);
                        }
                    }
#pragma unroll
                    for (int pass = 0; pass < 3; ++pass)
#pragma unroll
                        for (int tt = 0; tt < TG; ++tt)
#pragma unroll
                            for (int mi = 0; mi < MTW; ++mi)
#pragma unroll
                                for (int ni = 0; ni < NTW; ++ni) {
                                    if (pass == 0) mma_tf32(acc[tg + tt][mi][ni], al[mi], bh[tt][ni]);
                                    else if (pass == 1) mma_tf32(acc[tg + tt][mi][ni], ah[mi], bl[tt][ni]);
                                    else mma_tf32(acc[tg + tt][mi][ni], ah[mi], bh[tt][ni]);
                                }
                }
                }
            }

            if (oe < op1) {
                store_g(cur ^ 1, gcnt);
                if (ncnt) { store_x(na, ncnt); staged_hi = na + ncnt - 1; }
            }
            __syncthreads();
            cur ^= 1;
        }
    }

    // ---- leave: the CTA's sums are assembled in shared memory as an image of dW ([co][ci][tap], the reference layout) and leave
    //      as 128-bit reductions, one per four weights per CTA ----
    __syncthreads();                                             // ring / G' buffers are dead: reuse them
    const int ntot = p.Cout * p.Cin * p.ntaps;
    float* img = smem;
    for (int idx = tid; idx < ntot; idx += SW_NT) img[idx] = 0.f;
    __syncthreads();
#pragma unroll
    for (int tap = 0; tap < NTAPS; ++tap)
#pragma unroll
        for (int mi = 0; mi < MTW; ++mi)
#pragma unroll
            for (int ni = 0; ni < NTW; ++ni)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int co = (wm * MTW + mi) * 16 + fr + (e >= 2 ? 8 : 0), ci = (wn * NTW + ni) * 8 + fc * 2 + (e & 1);
                    if (tap < p.ntaps && co < p.Cout && ci < p.Cin) {
                        float* dst = img + ((size_t)co * p.Cin + ci) * p.ntaps + tap;
                        if (KWs > 1) atomicAdd(dst, acc[tap][mi][ni][e]); else *dst = acc[tap][mi][ni][e];
                    }
                }
    __syncthreads();
    if ((ntot & 3) == 0 && (reinterpret_cast<uintptr_t>(p.dw) & 15) == 0) {
        for (int q = tid; q < ntot / 4; q += SW_NT) atomicAdd(reinterpret_cast<float4*>(p.dw) + q, ld4(img + 4 * q));
    } else {
        for (int idx = tid; idx < ntot; idx += SW_NT) atomicAdd(p.dw + idx, img[idx]);
    }
}

// ------------------------------------------- host side -------------------------------------------
const bool g_use_slide = [] { const char* e = std::getenv("WF_DISABLE_SLIDE"); return !(e && e[0] == '1'); }();
// WF_SLIDE_MIN_CH: layers whose Cin and Cout are both below it stay on wf_thin.cu (A/B measurements)
const int g_slide_min_ch = [] { const char* e = std::getenv("WF_SLIDE_MIN_CH"); return e ? std::atoi(e) : 16; }();
// WF_DISABLE_SLIDE_THIN=1: <= 8-channel forward / backward-data layers stay on wf_thin.cu (A/B measurements)
const bool g_slide_thin = [] { const char* e = std::getenv("WF_DISABLE_SLIDE_THIN"); return !(e && e[0] == '1'); }();
const int g_slide_min_cin_wgrad = [] { const char* e = std::getenv("WF_SLIDE_MIN_CIN_WGRAD"); return e ? std::atoi(e) : 8; }();   // 1-channel layers: wf_thin.cu is faster (measured)
const int g_slide_min_ch_wgrad = [] { const char* e = std::getenv("WF_SLIDE_MIN_CH_WGRAD"); return e ? std::atoi(e) : 8; }();
constexpr int SMEM_MAX = 227 * 1024;
constexpr int PS_MAX = 4;

int device_sms() { return wf_device_sms(); }

// window geometry for steps of PS positions: widest step (slabs one step reads), the ring that lets the next step land while
// the current one is read, and the most slabs a step adds
void window_geometry(int pmul, int pdiv, int Pin, int Pout, int dpmin, int dpmax, int PS, int& span, int& ring, int& newmax)
{
    span = 1; ring = 1; newmax = 0;
    int staged = -1, prev_lo = 0;
    bool have_prev = false;
    for (int a = 0; a < Pout; a += PS) {
        const int b = a + PS < Pout ? a + PS : Pout;
        int lo, hi;
        step_range(pmul, pdiv, Pin, dpmin, dpmax, a, b, lo, hi);
        if (hi < lo) continue;
        if (hi - lo + 1 > span) span = hi - lo + 1;
        if (have_prev) {
            if (hi - prev_lo + 1 > ring) ring = hi - prev_lo + 1;
            const int na = staged + 1 > lo ? staged + 1 : lo;
            if (hi - na + 1 > newmax) newmax = hi - na + 1;
        }
        if (hi > staged) staged = hi;
        prev_lo = lo; have_prev = true;
    }
    if (ring < span) ring = span;
}

// positions per CTA (a multiple of PS): trade wave quantisation against re-staging the window head and the per-CTA setup
int pick_pc(int P, int PS, int coltiles, int slots, int span, double c_slab, double c_pos, double c_fix)
{
    int best = P; double best_cost = 1e300;
    for (int pc = PS; pc < P + PS; pc += PS) {
        const int pce = pc < P ? pc : P;
        const int ranges = (P + pce - 1) / pce;
        const long long ctas = (long long)coltiles * ranges;
        const double waves = (double)((ctas + slots - 1) / slots);
        const double cost = waves * (c_fix + (pce + span - 1) * c_slab + pce * c_pos);
        if (cost < best_cost * 0.999) { best_cost = cost; best = pce; }
    }
    return best;
}

void tap_range(const int* dp, int ntaps, int& dpmin, int& dpmax)
{
    dpmin = dp[0]; dpmax = dp[0];
    for (int t = 1; t < ntaps; ++t) { dpmin = dp[t] < dpmin ? dp[t] : dpmin; dpmax = dp[t] > dpmax ? dp[t] : dpmax; }
}

template <int MI, int NI, int MW, int NW, int LD, int MINB, int PRO, bool HASDN>
cudaError_t launch_slide_conv(const ConvP& p, int num_sms, cudaStream_t st, bool dry)
{
    constexpr int NT = 32 * MW * NW, BM = 16 * MI * MW, BN = 8 * NI * NW, XS = BN + 2 * SL_H, Q = XS / 4, WS = BM + 8;
    const int K8 = (p.Cin + 7) & ~7;
    SlideGeo g{};
    tap_range(p.dp, p.ntaps, g.dpmin, g.dpmax);
    g.cap = (LD * NT) / (p.Cin * Q);
    if (g.cap > 15) g.cap = 15;
    if (g.cap < 1) return cudaErrorInvalidConfiguration;
    const size_t wbytes = ((size_t)p.ntaps * K8 * WS + 4 * K8) * 4, slab = (size_t)K8 * XS * 4;
    int span = 0;
    g.PS = 0;
    for (int ps = PS_MAX; ps >= 1 && !g.PS; --ps) {
        int ring, newmax;
        window_geometry(p.pmul, p.pdiv, p.Pin, p.Pout, g.dpmin, g.dpmax, ps, span, ring, newmax);
        if (newmax > g.cap) continue;
        if (wbytes + ring * slab <= (size_t)SMEM_MAX - 2048 && (ps == 1 || ring * slab <= 64 * 1024)) { g.PS = ps; g.R = ring; g.two_sync = 0; }
        else if (ps == 1 && wbytes + span * slab <= (size_t)SMEM_MAX - 2048) { g.PS = 1; g.R = span; g.two_sync = 1; }
    }
    if (!g.PS) return cudaErrorInvalidConfiguration;
    const size_t smem = wbytes + g.R * slab;
    if (dry) return cudaSuccess;
    static WfSmemOptIn optin;
    if (cudaError_t e = wf_smem_optin(optin, slide_conv_kernel<MI, NI, MW, NW, LD, MINB, PRO, HASDN>, smem)) return e;
    int occ = (int)((size_t)SMEM_MAX / (smem + 2048));
    if (occ > MINB) occ = MINB;
    if (occ < 1) occ = 1;
    const int coltiles = (p.N + BN - 1) / BN;
    const double c_slab = (double)p.Cin * XS * 3.0 / 128.0 * 4.0;
    const double c_pos = (double)(BM / 16) * (BN / 8) * (K8 / 8) * p.ntaps * 3 * 2.0 * 1.5 + 150.0;
    const double c_fix = (double)p.ntaps * K8 * BM / 128.0 * 2.0 + 600.0;
    g.PC = pick_pc(p.Pout, g.PS, coltiles, num_sms * occ, span, c_slab, c_pos, c_fix);
    dim3 grid(coltiles, (p.Pout + g.PC - 1) / g.PC, 1);
    wf_launch_pdl(slide_conv_kernel<MI, NI, MW, NW, LD, MINB, PRO, HASDN>, dim3(grid), dim3(NT), smem, st, p, g);
    return cudaGetLastError();
}

template <int MI, int NI, int MW, int NW, int LD, int MINB>
cudaError_t slide_conv_pro(const ConvP& p, int num_sms, cudaStream_t st, bool dry)
{
    bool hasdn = false;
    for (int t = 0; t < p.ntaps; ++t) hasdn |= p.dn[t] != 0;
    if (hasdn) {        // only the decoder 3x3 conv has time taps: BatchNorm-only prologue forward, BatchNorm-backward prologue backward
        if (p.pro_mode == PRO_AFFINE) return launch_slide_conv<MI, NI, MW, NW, LD, MINB, PRO_AFFINE, true>(p, num_sms, st, dry);
        if (p.pro_mode == PRO_BNBWD) return launch_slide_conv<MI, NI, MW, NW, LD, MINB, PRO_BNBWD, true>(p, num_sms, st, dry);
        return cudaErrorInvalidConfiguration;
    }
    switch (p.pro_mode) {
        case PRO_NONE: return launch_slide_conv<MI, NI, MW, NW, LD, MINB, PRO_NONE, false>(p, num_sms, st, dry);
        case PRO_BNSILU: return launch_slide_conv<MI, NI, MW, NW, LD, MINB, PRO_BNSILU, false>(p, num_sms, st, dry);
        case PRO_AFFINE: return launch_slide_conv<MI, NI, MW, NW, LD, MINB, PRO_AFFINE, false>(p, num_sms, st, dry);
        default: return launch_slide_conv<MI, NI, MW, NW, LD, MINB, PRO_BNBWD, false>(p, num_sms, st, dry);
    }
}

template <int MI, int NTAPS, int K8S, int LD, int MINB, int PRO>
cudaError_t launch_slide_thin(const ConvP& p, int num_sms, cudaStream_t st, bool dry)
{
    constexpr int NT = 256, BN = 16 * MI * 8, XS = BN + 2 * SL_H, Q = XS / 4, K8 = 8 * K8S;
    SlideGeo g{};
    tap_range(p.dp, p.ntaps, g.dpmin, g.dpmax);
    g.cap = (LD * NT) / (p.Cin * Q);
    if (g.cap > 15) g.cap = 15;
    if (g.cap < 1) return cudaErrorInvalidConfiguration;
    const size_t fixed = ((size_t)4 * K8 + 8 * 8 * 36) * 4, slab = (size_t)K8 * XS * 4;       // coefficients + per-warp epilogue scratch
    int span = 0;
    g.PS = 0;
    for (int ps = PS_MAX; ps >= 1 && !g.PS; --ps) {
        int ring, newmax;
        window_geometry(p.pmul, p.pdiv, p.Pin, p.Pout, g.dpmin, g.dpmax, ps, span, ring, newmax);
        if (newmax > g.cap) continue;
        if (ps == 1 || ring * slab <= 56 * 1024) { g.PS = ps; g.R = ring; }
    }
    if (!g.PS) return cudaErrorInvalidConfiguration;
    const size_t smem = fixed + g.R * slab;
    if (dry) return cudaSuccess;
    static WfSmemOptIn optin;
    if (cudaError_t e = wf_smem_optin(optin, slide_thin_kernel<MI, NTAPS, K8S, LD, MINB, PRO>, smem)) return e;
    int occ = (int)((size_t)SMEM_MAX / (smem + 2048));
    if (occ > MINB) occ = MINB;
    if (occ < 1) occ = 1;
    const int coltiles = (p.N + BN - 1) / BN;
    const double c_slab = (double)p.Cin * XS * 3.0 / 128.0 * 4.0;
    const double c_pos = (double)(BN / 16) * K8S * p.ntaps * 3 * 2.0 * 1.5 + (double)BN * 8 / 128.0 * 8.0 + 100.0;
    g.PC = pick_pc(p.Pout, g.PS, coltiles, num_sms * occ, span, c_slab, c_pos, 400.0);
    dim3 grid(coltiles, (p.Pout + g.PC - 1) / g.PC, 1);
    wf_launch_pdl(slide_thin_kernel<MI, NTAPS, K8S, LD, MINB, PRO>, dim3(grid), dim3(NT), smem, st, p, g);
    return cudaGetLastError();
}

template <int NTAPS, int K8S>
cudaError_t slide_thin_pro(const ConvP& p, int num_sms, cudaStream_t st, bool dry)
{
    switch (p.pro_mode) {
        case PRO_NONE: return launch_slide_thin<2, NTAPS, K8S, 5, 2, PRO_NONE>(p, num_sms, st, dry);
        case PRO_BNSILU: return launch_slide_thin<2, NTAPS, K8S, 5, 2, PRO_BNSILU>(p, num_sms, st, dry);
        case PRO_AFFINE: return launch_slide_thin<2, NTAPS, K8S, 5, 2, PRO_AFFINE>(p, num_sms, st, dry);
        default: return launch_slide_thin<2, NTAPS, K8S, 5, 2, PRO_BNBWD>(p, num_sms, st, dry);
    }
}

// <= 8 output channels, 8..16 input channels, position taps only
bool slide_thin_shape(const ConvP& p)
{
    if (p.Cout > 8 || p.Cin > 16 || (p.ntaps != 1 && p.ntaps != 3) || p.Mpad > 8) return false;
    for (int t = 0; t < p.ntaps; ++t) if (p.dn[t] != 0) return false;
    return true;
}
cudaError_t slide_thin_dispatch(const ConvP& p, int num_sms, cudaStream_t st, bool dry)
{
    if (p.ntaps == 3) return p.Cin > 8 ? slide_thin_pro<3, 2>(p, num_sms, st, dry) : slide_thin_pro<3, 1>(p, num_sms, st, dry);
    return p.Cin > 8 ? slide_thin_pro<1, 2>(p, num_sms, st, dry) : slide_thin_pro<1, 1>(p, num_sms, st, dry);
}

cudaError_t slide_conv_dispatch(const ConvP& p, int num_sms, cudaStream_t st, bool dry)
{
    if (g_slide_thin && slide_thin_shape(p)) return slide_thin_dispatch(p, num_sms, st, dry);
    if (p.Cout > 32) return slide_conv_pro<2, 2, 2, 8, 5, 1>(p, num_sms, st, dry);        // 64 x 128 tile, 16 warps of 32 x 16
    if (p.Cout > 16 && p.Cin <= 32 && p.ntaps <= 3)
        return slide_conv_pro<2, 2, 1, 8, 5, 2>(p, num_sms, st, dry);                     // same tile, <= 32 input channels: two CTAs per SM
    if (p.Cout > 16) return slide_conv_pro<2, 2, 1, 8, 9, 1>(p, num_sms, st, dry);        // 32 x 128 tile, 8 warps of 32 x 16
    return slide_conv_pro<1, 2, 1, 8, 5, 2>(p, num_sms, st, dry);                         // 16 x 128 tile, 8 warps of 16 x 16
}

struct WgCfg { int mtw, ntw, mws, nws, kws; };
bool wgrad_cfg(const WgradP& p, WgCfg& c)
{
    const int MT = (p.Cout + 15) / 16, NT8 = (p.Cin + 7) / 8;
    if ((MT != 1 && MT != 2 && MT != 4) || (NT8 != 1 && NT8 != 2 && NT8 != 4 && NT8 != 8)) return false;
    c.mtw = MT == 4 ? 2 : (MT == 2 && NT8 == 8 ? 2 : 1);
    c.mws = MT / c.mtw;
    c.ntw = (MT == 4 && NT8 == 8) ? 2 : 1;
    c.nws = NT8 / c.ntw;
    if (c.mws * c.nws > 8) return false;
    c.kws = 8 / (c.mws * c.nws);
    return true;
}

template <int MTW, int NTW, int NTAPS, int LDG, int LDX, int MINB, int PRO>
cudaError_t launch_slide_wgrad(const WgradP& p, const WgCfg& c, int num_sms, cudaStream_t st, bool dry)
{
    const int BM = 16 * MTW * c.mws, BCI = 8 * NTW * c.nws;
    SlideGeo g{};
    tap_range(p.dp, p.ntaps, g.dpmin, g.dpmax);
    g.MWs = c.mws; g.NWs = c.nws; g.KWs = c.kws;
    g.cap = (LDX * SW_NT) / (p.Cin * SW_XQ);
    if (g.cap > 15) g.cap = 15;
    const int gcap = (LDG * SW_NT) / (p.Cout * SW_GQ);
    if (g.cap < 1 || gcap < 1) return cudaErrorInvalidConfiguration;
    const size_t fixed = ((size_t)4 * BM + 4 * BCI) * 4;
    const size_t gsl = (size_t)BM * SW_GS * 4, xsl = (size_t)BCI * SW_XS * 4;
    int span = 0;
    g.PS = 0;
    for (int ps = PS_MAX; ps >= 1 && !g.PS; --ps) {
        int ring, newmax;
        window_geometry(p.pmul, 1, p.Pin, p.Pout, g.dpmin, g.dpmax, ps, span, ring, newmax);
        if (newmax > g.cap || ps > gcap) continue;
        const size_t need = fixed + 2 * ps * gsl + (ring + 1) * xsl;
        if (need <= (size_t)SMEM_MAX - 1024 && (ps == 1 || need <= 64 * 1024)) { g.PS = ps; g.R = ring; }
    }
    if (!g.PS) return cudaErrorInvalidConfiguration;
    size_t smem = fixed + 2 * g.PS * gsl + (g.R + 1) * xsl;
    const size_t image = (size_t)p.Cout * p.Cin * p.ntaps * 4;          // the dW image assembled at the end reuses the staging buffers
    if (smem < image) smem = image;
    if (smem > (size_t)SMEM_MAX - 1024) return cudaErrorInvalidConfiguration;
    if (dry) return cudaSuccess;
    static WfSmemOptIn optin;
    if (cudaError_t e = wf_smem_optin(optin, slide_wgrad_kernel<MTW, NTW, NTAPS, LDG, LDX, MINB, PRO>, smem)) return e;
    int occ = (int)((size_t)SMEM_MAX / (smem + 1024));
    if (occ > MINB) occ = MINB;
    if (occ < 1) occ = 1;
    g.ncol = (p.N + SW_BN - 1) / SW_BN;
    const int slots = num_sms * occ;
    const double c_slab = ((double)p.Cin * (SW_BN + 8) + (double)p.Cout * SW_BN) * 3.0 / 128.0 * 4.0;
    const double c_pos = (double)(BM / 16) * (BCI / 8) * (SW_BN / 8) * NTAPS * 3 * 2.0 * 1.5 + 150.0;
    g.PC = pick_pc(p.Pout, g.PS, g.ncol, slots, span, c_slab, c_pos, 300.0);
    g.nitems = g.ncol * ((p.Pout + g.PC - 1) / g.PC);
    const int grid = g.nitems < slots ? g.nitems : slots;
    wf_launch_pdl(slide_wgrad_kernel<MTW, NTW, NTAPS, LDG, LDX, MINB, PRO>, dim3(grid), dim3(SW_NT), smem, st, p, g);
    return cudaGetLastError();
}

template <int MTW, int NTW, int NTAPS, int LDG, int LDX, int MINB>
cudaError_t slide_wgrad_pro(const WgradP& p, const WgCfg& c, int num_sms, cudaStream_t st, bool dry)
{
    switch (p.pro_mode) {
        case PRO_NONE: return launch_slide_wgrad<MTW, NTW, NTAPS, LDG, LDX, MINB, PRO_NONE>(p, c, num_sms, st, dry);
        case PRO_BNSILU: return launch_slide_wgrad<MTW, NTW, NTAPS, LDG, LDX, MINB, PRO_BNSILU>(p, c, num_sms, st, dry);
        case PRO_AFFINE: return launch_slide_wgrad<MTW, NTW, NTAPS, LDG, LDX, MINB, PRO_AFFINE>(p, c, num_sms, st, dry);
        default: return cudaErrorInvalidConfiguration;
    }
}

cudaError_t slide_wgrad_dispatch(const WgradP& p, int num_sms, cudaStream_t st, bool dry)
{
    WgCfg c;
    if (!wgrad_cfg(p, c)) return cudaErrorInvalidConfiguration;
    const int nt = p.ntaps;
    if (nt != 9)
        for (int t = 0; t < nt; ++t) if (p.dn[t] != 0) return cudaErrorInvalidConfiguration;
    // LDG / LDX: float4 registers per thread for the G' slabs (Cout*16 quads each) and the X' slabs (Cin*18 quads each) in flight
    if (c.mtw == 2 && c.ntw == 2 && nt == 3) return slide_wgrad_pro<2, 2, 3, 4, 5, 1>(p, c, num_sms, st, dry);
    if (c.mtw == 2 && c.ntw == 1 && nt == 3) return slide_wgrad_pro<2, 1, 3, 4, 5, 1>(p, c, num_sms, st, dry);
    if (c.mtw == 2 && c.ntw == 1 && nt == 1) return slide_wgrad_pro<2, 1, 1, 4, 5, 1>(p, c, num_sms, st, dry);
    if (c.mtw == 2 && c.ntw == 1 && nt == 9) return slide_wgrad_pro<2, 1, 9, 4, 5, 1>(p, c, num_sms, st, dry);
    if (c.mtw == 1 && c.ntw == 1 && nt == 3) return slide_wgrad_pro<1, 1, 3, 3, 4, 2>(p, c, num_sms, st, dry);
    if (c.mtw == 1 && c.ntw == 1 && nt == 1) return slide_wgrad_pro<1, 1, 1, 3, 4, 2>(p, c, num_sms, st, dry);
    return cudaErrorInvalidConfiguration;
}

bool fits_int32(long long v) { return v >= 0 && v < (1LL << 31); }

}  // namespace

// position-tap convs (any |dn| <= halo), one group, enough input channels to feed a k8 step, Dropout2d-style masks only
bool wf_slide_conv_ok(const ConvP& p)
{
    if (!g_use_slide || p.groups != 1 || p.Cout > 64 || p.Cin > 64 || p.Cin < 8 || p.Mpad < p.Cout || (p.Mpad & 3)) return false;
    if (p.Cin < g_slide_min_ch && p.Cout < g_slide_min_ch && !(g_slide_thin && slide_thin_shape(p))) return false;
    if (p.pdiv != 1 && p.pdiv != 2) return false;
    if (p.Pin == 1 && p.Pout == 1) return false;              // pure time-tap layers belong to wf_group.cu / wf_tc.cu
    if ((p.mask && p.m_st != 0) || (p.emask && p.em_st != 0)) return false;
    if (!fits_int32((long long)p.Cin * p.in_sc + (long long)p.Pin * p.in_sp) || !fits_int32((long long)p.Cout * p.out_sc + (long long)p.Pout * p.out_sp)) return false;
    for (int t = 0; t < p.ntaps; ++t)
        if (p.dn[t] < -SL_H || p.dn[t] > SL_H) return false;
    return slide_conv_dispatch(p, 148, nullptr, true) == cudaSuccess;
}
cudaError_t wf_launch_slide_conv(const ConvP& p, cudaStream_t st) { return slide_conv_dispatch(p, device_sms(), st, false); }
// true when wf_launch_slide_conv runs slide_thin_kernel for this layer (profiling labels)
bool wf_slide_conv_is_thin(const ConvP& p) { return g_slide_thin && slide_thin_shape(p); }

bool wf_slide_wgrad_ok(const WgradP& p)
{
    if (!g_use_slide || p.groups != 1 || p.Cout > 64 || p.Cin > 64 || p.Cin < g_slide_min_cin_wgrad) return false;
    if (p.Cin < g_slide_min_ch_wgrad && p.Cout < g_slide_min_ch_wgrad) return false;
    if (p.Pin == 1 && p.Pout == 1) return false;
    if (p.mask && p.m_st != 0) return false;
    if (!fits_int32((long long)p.Cin * p.in_sc + (long long)p.Pin * p.in_sp) || !fits_int32((long long)p.Cout * p.Pout * p.N)) return false;
    for (int t = 0; t < p.ntaps; ++t)
        if (p.dn[t] < -SL_H || p.dn[t] > SL_H) return false;
    return slide_wgrad_dispatch(p, 148, nullptr, true) == cudaSuccess;
}
cudaError_t wf_launch_slide_wgrad(const WgradP& p, int num_sms, cudaStream_t st) { return slide_wgrad_dispatch(p, num_sms, st, false); }
