// tcgen05 / TMEM kernels for the pointwise (1x1) convolutions of WiFlow -- the TCN channel-mixing GEMMs
// (models/tcn.py:27,40,47: 48.8% of all FLOPs) and the attention QKV projection (models/attention.py:22-24,50).
//
//   forward / backward-data :  D[m][col] = sum_k W[m][k] * act(X[k][col])         (pw_tc_kernel)
//   backward-weights        :  dW[m][c] += sum_col G[m][col] * act(X[c][col])     (pw_wgrad_tc_kernel)
//
// fp32 parity (1e-4 on outputs) rules out plain TF32, so every product is the 3xTF32 split
//   a*b ~= a_lo*b_hi + a_hi*b_lo + a_hi*b_hi     (hi = rna_tf32(x), lo = rna_tf32(x - hi); error ~2^-21)
// accumulated in fp32 in TMEM.  Operand staging (all operands K-major, UMMA no-swizzle core-matrix order: an operand tile is a
// grid of 8-row x 16-byte core matrices, 128 contiguous bytes each):
//   * weights are split and laid out ONCE per step by tc_pack_kernel as ready-to-use shared-memory images
//     ([M tile][K chunk][hi|lo]) and arrive by one 32 KB bulk async copy (cp.async.bulk + mbarrier complete_tx) per stage;
//   * activations go global -> registers -> (BatchNorm + SiLU + Dropout | BatchNorm-backward) -> hi/lo split -> shared
//     memory as an MN-major SWIZZLE_128B_BASE32B image (forward / backward-data: columns contiguous, as in HBM, no transpose;
//     a quarter warp fills one 128-byte row), so the previous layer's normalisation never costs an HBM round trip.
// One elected thread issues tcgen05.mma (M=128, N<=256, K=8 per instruction); tcgen05.commit releases the stage and
// finally hands the accumulators to the same 16 warps, which run the epilogue (bias / SiLU' / BatchNorm statistics) out of
// TMEM with one thread per output channel: the per-channel sums need no shuffles at all.
// The kernels are issue bound on the producer warps (profiles/), so the prologue mode is a template parameter and all
// addressing is strength-reduced out of the K loop.
#include <cstdlib>
#include "wf_tc.cuh"
#include "wf_common.cuh"
#include "wf_elem.h"

namespace {

using namespace tc;

constexpr int KC = TC_KC;                 // K elements per pipeline stage
constexpr int BM = 128;                   // UMMA M
constexpr int NPROD = 512;                // producer / epilogue threads (warps 0-15)
constexpr int NPW = NPROD / 32;
constexpr int NTHREADS = NPROD + 64;      // + one warp issuing the MMAs + one warp for TMEM allocation and weight copies
constexpr int A_HALF = BM * KC * 4;       // one of hi/lo of a 128 x KC K-major operand tile
constexpr int A_LBO = 128, A_SBO = (KC / 4) * 128;      // K-major image: KC/4 core matrices along K, then the next 8 rows

// prologue of one float4 (4 consecutive columns of one channel); MODE is a compile-time PRO_* value
template <int MODE, bool MASK>
__device__ __forceinline__ float4 pro4(float4 v, float4 v2, float a, float b, float c, float d)
{
    if (MODE == PRO_BNSILU) {
        v.x = wf_silu(fmaf(a, v.x - d, b)); v.y = wf_silu(fmaf(a, v.y - d, b)); v.z = wf_silu(fmaf(a, v.z - d, b)); v.w = wf_silu(fmaf(a, v.w - d, b));
        if (MASK) { v.x *= v2.x; v.y *= v2.y; v.z *= v2.z; v.w *= v2.w; }
    } else if (MODE == PRO_AFFINE) {
        v.x = fmaf(a, v.x - d, b); v.y = fmaf(a, v.y - d, b); v.z = fmaf(a, v.z - d, b); v.w = fmaf(a, v.w - d, b);
    } else if (MODE == PRO_BNBWD) {
        v.x = fmaf(a, v.x, fmaf(b, v2.x - d, c)); v.y = fmaf(a, v.y, fmaf(b, v2.y - d, c));
        v.z = fmaf(a, v.z, fmaf(b, v2.z - d, c)); v.w = fmaf(a, v.w, fmaf(b, v2.w - d, c));
    }
    return v;
}

__device__ __forceinline__ void split_store(uint8_t* hi_base, uint8_t* lo_base, uint32_t off, float4 v)
{
    float4 h, l;
    tf32_split(v.x, h.x, l.x); tf32_split(v.y, h.y, l.y); tf32_split(v.z, h.z, l.z); tf32_split(v.w, h.w, l.w);
    *reinterpret_cast<float4*>(hi_base + off) = h;
    *reinterpret_cast<float4*>(lo_base + off) = l;
}

__device__ __forceinline__ float4 sel4(bool c, float4 a, float4 b)
{
    return make_float4(c ? a.x : b.x, c ? a.y : b.y, c ? a.z : b.z, c ? a.w : b.w);
}

// =========================================================================================================
// forward / backward-data
// =========================================================================================================
// TMEM holds two accumulators per tile: the a_hi*b_hi products go to columns [0, bn) and the two small correction products
// to columns [bnp, bnp + bn).  The tensor core truncates when it adds into the fp32 accumulator, an error that grows with the
// number of accumulations; keeping the 2^-11-sized corrections out of the main accumulator cuts that count by three and
// makes their own truncation irrelevant.  The epilogue adds the two in fp32.
struct TcGeom { int bn, bnp, nst, nsa, tmem_cols, mp, mtiles, dbg, mgroups, total; };      // nst / nsa: stages of the activation / weight ring      // mp: 128-row M tiles per CTA (they share the staged activation tile)

// PERSIST: a CTA walks tiles blockIdx.x, + gridDim.x, ...; otherwise exactly one tile (straight-line code: the loop-carried state of
// the persistent form costs the BatchNorm+SiLU producers 13 % at the 96-register budget of an 18-warp CTA -- 5 warps per scheduler
// x 32 x R <= 16 K registers -- so it is used only where a CTA gets many tiles)
template <int PRO, bool MASK, bool PERSIST>
__global__ void __launch_bounds__(NTHREADS, 1) pw_tc_kernel(const ConvP p, const TcGeom g)
{
    wf_pdl_enter();
    const int BN = g.bn, STAGES = g.nst;
    const int B_HALF = KC * BN * 4;                       // one of hi/lo of the BN x KC activation tile
    const int MP = g.mp;
    const int A_BYTES = MP * 2 * A_HALF;                  // weight images (hi, lo) of this CTA's M tiles
    const int B_STAGE = 2 * B_HALF;
    const int NSA = g.nsa;
    const int NQ = BN / 4;                                // column quads per tile
    // Two rings: the activation tiles (64 KB a stage at bn = 256, so two stages) and the weight images (32 KB a stage, three or
    // four).  With ONE ring of two stages the bulk copy of a stage's weights could only be requested when the MMAs two K steps back
    // had retired and its ~1 us latency sat in every K step (measured: 1.3 us per step with the producers switched off, against
    // 0.8 us of MMA time); the deeper weight ring asks for the image two or three steps ahead.
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* smem_a = smem + STAGES * B_STAGE;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_a + NSA * A_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 2 * NSA + 1);
    const uint32_t bar0 = smem_u32(bars);
    auto full_b = [&](int s) { return bar0 + 8u * s; };
    auto empty = [&](int s) { return bar0 + 8u * (STAGES + s); };
    auto full_a = [&](int s) { return bar0 + 8u * (2 * STAGES + s); };
    auto empty_a = [&](int s) { return bar0 + 8u * (2 * STAGES + NSA + s); };
    const uint32_t accum_bar = bar0 + 8u * (2 * STAGES + 2 * NSA);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int KT = p.tc_kt;
    const long long NC = (long long)p.Pout * p.N;
    // Persistent CTAs: tile = blockIdx.x, + gridDim.x, ... over (column tile, M-tile group), M fastest so that the CTAs working at the
    // same time share the activation columns in L2.  Launch, TMEM allocation and barrier set-up are paid once per CTA instead of once
    // per tile (~5 of the ~32 us of a 128 x 256 tile); the rings and their phases simply run on across tiles.
#define WF_TC_TILE_VARS(tile)                                          \
    const int mt0 = ((tile) % g.mgroups) * MP;                         \
    const int ntile = min(MP, g.mtiles - mt0);                         \
    const int m0 = mt0 * BM;                                           \
    const long long col0 = (long long)((tile) / g.mgroups) * BN;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full_b(s), NPW); mbar_init(empty(s), 1); }
        for (int s = 0; s < NSA; ++s) { mbar_init(full_a(s), 1); mbar_init(empty_a(s), 1); }
        mbar_init(accum_bar, 1);
        fence_mbar_init();
    }
    if (warp == NPW + 1) { tmem_alloc(smem_u32(tmem_slot), g.tmem_cols); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < NPW) {
        // ------------------------------ activation producers ------------------------------
        // HBM holds the activations column-contiguous, i.e. as an MN-major operand.  A thread owns 4 channels x 4 columns: four
        // coalesced 128-bit loads (one per channel), prologue, split, and each float4 goes to shared memory as it is (round 1
        // transposed 4x4 micro blocks in registers into a K-major image: 32 selects per chunk and thread).
        const bool active = tid < (KC / 4) * NQ;
        const int q = tid % NQ, kq = tid / NQ;
        int s = 0; uint32_t ph = 0;
        int tile = blockIdx.x, it = 0;
        do {
        WF_TC_TILE_VARS(tile)
        const long long col = col0 + q * 4;
        const bool cval = active && col < NC;
        const float *pin = p.in, *pin2 = p.in2, *pm = p.mask;
        {
            long long pos = 0, n = col;
            if (p.Pout > 1) { pos = col / p.N; n = col - pos * p.N; }
            const long long b = n / WF_T; const int t = (int)(n - b * WF_T);
            const long long off = pos * p.in_sp + b * p.in_sb + t + (long long)(kq * 4) * p.in_sc;
            pin += off;
            if (PRO == PRO_BNBWD) pin2 += off;
            if (MASK) pm += b * p.m_sb + (long long)t * p.m_st + (long long)(kq * 4) * p.m_sc;
        }
        const long long in_sc = p.in_sc, m_sc = p.m_sc;
        const int m_st = p.m_st, Cin = p.Cin;
        // MN-major operand image (SWIZZLE_128B_BASE32B): row = channel (K) of the stage, 128 bytes = 32 columns, 32-byte chunks XORed
        // with the row index mod 4; 32-column blocks KC*128 bytes apart.  A thread's float4 (4 columns of one channel) is one 16-byte
        // chunk of that image: no transpose, and a quarter warp (8 lanes = 8 consecutive column quads) fills one 128-byte row.
        uint32_t soff[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int row = kq * 4 + j;
            soff[j] = (uint32_t)((q >> 3) * (KC * 128) + row * 128 + (((q & 7) ^ ((row & 3) << 1)) << 4));
        }
        for (int kc = 0; kc < KT; ++kc) {
            const int c0 = kc * KC + kq * 4;
            float4 v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = f4zero();
            if (cval && c0 < Cin && !(g.dbg & 1)) {
                float4 v2[4];
                float4 A4 = f4zero(), B4 = f4zero(), C4 = f4zero(), D4 = f4zero();
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    v2[j] = f4zero();
                    if (c0 + j < Cin) {
                        v[j] = ld4(pin + j * in_sc);
                        if (PRO == PRO_BNBWD) v2[j] = ld4(pin2 + j * in_sc);
                        if (MASK) {
                            const float* mp = pm + j * m_sc;
                            if (m_st == 1) v2[j] = ld4(mp); else { const float mm = *mp; v2[j] = make_float4(mm, mm, mm, mm); }
                        }
                    }
                }
                if (PRO != PRO_NONE) {       // per-channel coefficients of the 4 channels: one 128-bit load per array
                    A4 = ld4(p.pro_a + c0); B4 = ld4(p.pro_b + c0); D4 = ld4(p.pro_d + c0);      // (bases from the constant bank: no live pointers)
                    if (PRO == PRO_BNBWD) C4 = ld4(p.pro_c + c0);
                    v[0] = pro4<PRO, MASK>(v[0], v2[0], A4.x, B4.x, C4.x, D4.x);
                    v[1] = pro4<PRO, MASK>(v[1], v2[1], A4.y, B4.y, C4.y, D4.y);
                    v[2] = pro4<PRO, MASK>(v[2], v2[2], A4.z, B4.z, C4.z, D4.z);
                    v[3] = pro4<PRO, MASK>(v[3], v2[3], A4.w, B4.w, C4.w, D4.w);
#pragma unroll
                    for (int j = 0; j < 4; ++j) if (c0 + j >= Cin) v[j] = f4zero();
                }
            }
            pin += KC * in_sc;
            if (PRO == PRO_BNBWD) pin2 += KC * in_sc;
            if (MASK) pm += KC * m_sc;
            mbar_wait(empty(s), ph ^ 1u);
            if (active && !(g.dbg & 8)) {
                uint8_t* bh = smem + s * B_STAGE;
                uint8_t* bl = bh + B_HALF;
                split_store(bh, bl, soff[0], v[0]);
                split_store(bh, bl, soff[1], v[1]);
                split_store(bh, bl, soff[2], v[2]);
                split_store(bh, bl, soff[3], v[3]);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(full_b(s));
            if (++s == STAGES) { s = 0; ph ^= 1u; }
        }

        // ------------------------------ epilogue: one thread per output channel ------------------------------
        mbar_wait(accum_bar, (uint32_t)it & 1u);
        tc_fence_after();
        const int quarter = warp & 3, cgrp = warp >> 2;
        for (int tj = 0; tj < ntile; ++tj) {
        const int m = m0 + tj * BM + quarter * 32 + lane;
        const bool mv = m < p.Cout;
        const int co = m;
        float bias = 0.f, es = 0.f, et = 0.f, em = 0.f;
        if (mv) {
            if (p.bias) bias = p.bias[co];
            if (p.epi_mode == EPI_DSILU) { es = p.e_scale[co]; et = p.e_shift[co]; }
            if (p.epi_mode == EPI_DSILU || p.epi_mode == EPI_DAFF) em = p.e_mean[co];
        }
        float s0 = 0.f, s1 = 0.f;
        const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(tj * 2 * g.bnp);
        // Thread = output channel is the TMEM access pattern, but as a GLOBAL access pattern it is 32 scattered 16-byte pieces per
        // instruction (measured: 7 of the 37 us of a 128 x 256 tile).  So every 32 x 32 block goes through a per-warp staging tile
        // in the (by now idle) activation ring: raw tensor of the BatchNorm-backward epilogues in, results out, both as whole
        // 128-byte rows (8 lanes x float4 per row, 4 rows per instruction).
        const bool coal = p.out_sb == WF_T && (p.N & 3) == 0 && !(p.accumulate && p.epi_mode != EPI_STORE);
        const bool need_raw = p.epi_mode == EPI_DSILU || p.epi_mode == EPI_DAFF;
        float* stg = reinterpret_cast<float*>(smem) + warp * (32 * 36);          // [32 channels][32 columns + 4 pad]
        const int lr = lane >> 3, lq = lane & 7;
        for (int cb0 = 0; cb0 < BN; cb0 += 32) {
            if (((cb0 >> 5) % (NPW / 4)) != cgrp) continue;
            if (g.dbg & 64) continue;
            // the lane's column quad in the row-wise passes
            const long long colq = col0 + cb0 + lq * 4;
            const bool qv = cb0 + lq * 4 < BN && colq < NC;
            long long coff = 0;
            if (coal && qv) {
                long long pos = 0, n = colq;
                if (p.Pout > 1) { pos = colq / p.N; n = colq - pos * p.N; }
                const long long bb = n / WF_T;
                coff = pos * p.out_sp + bb * p.out_sb + (n - bb * WF_T);
            }
            const int mrow0 = m0 + tj * BM + quarter * 32;
            if (coal && need_raw) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int row = 4 * i + lr;
                    float4 e = f4zero();
                    if (qv && mrow0 + row < p.Cout) e = ld4(p.eraw + (long long)(mrow0 + row) * p.out_sc + coff);
                    *reinterpret_cast<float4*>(stg + row * 36 + lq * 4) = e;
                }
                __syncwarp();
            }
            float acc[32], cor[32];
            tmem_ld32(trow + (uint32_t)cb0, acc);
            tmem_ld32(trow + (uint32_t)(g.bnp + cb0), cor);
            tmem_ld_wait();
            if (g.dbg & 32) continue;
            if (!coal) {
                if (mv) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const long long colj = col0 + cb0 + j * 4;
                        if (cb0 + j * 4 < BN && colj < NC) {
                            int pos = 0; long long n = colj;
                            if (p.Pout > 1) { pos = (int)(colj / p.N); n = colj - (long long)pos * p.N; }
                            float q4[4] = {acc[j * 4 + 0] + cor[j * 4 + 0] + bias, acc[j * 4 + 1] + cor[j * 4 + 1] + bias,
                                           acc[j * 4 + 2] + cor[j * 4 + 2] + bias, acc[j * 4 + 3] + cor[j * 4 + 3] + bias};
                            wf_epilogue_quad(p, co, pos, (int)n, es, et, em, q4, s0, s1);
                        }
                    }
                }
                continue;
            }
            // channel-wise pass: epilogue math and statistics, result into the staging tile
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const long long colj = col0 + cb0 + j * 4;
                float v[4] = {acc[j * 4 + 0] + cor[j * 4 + 0] + bias, acc[j * 4 + 1] + cor[j * 4 + 1] + bias,
                              acc[j * 4 + 2] + cor[j * 4 + 2] + bias, acc[j * 4 + 3] + cor[j * 4 + 3] + bias};
                float* sp = stg + lane * 36 + j * 4;
                if (mv && cb0 + j * 4 < BN && colj < NC) {
                    if (p.epi_mode == EPI_STATS) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) { s0 += v[e]; s1 = fmaf(v[e], v[e], s1); }
                    } else if (need_raw) {
                        const float4 r4 = *reinterpret_cast<const float4*>(sp);
                        const float r[4] = {r4.x, r4.y, r4.z, r4.w};
                        if (p.epi_mode == EPI_DSILU) {
                            float mk[4] = {1.f, 1.f, 1.f, 1.f};
                            if (p.emask) {
                                long long n = colj;
                                if (p.Pout > 1) n = colj % p.N;
                                const long long bb = n / WF_T; const int t = (int)(n - bb * WF_T);
                                const float* mp = p.emask + bb * p.em_sb + (long long)co * p.em_sc + (long long)t * p.em_st;
                                if (p.em_st == 1) { const float4 m4 = ld4(mp); mk[0] = m4.x; mk[1] = m4.y; mk[2] = m4.z; mk[3] = m4.w; }
                                else { const float mm = *mp; mk[0] = mk[1] = mk[2] = mk[3] = mm; }
                            }
#pragma unroll
                            for (int e = 0; e < 4; ++e) v[e] = v[e] * mk[e] * wf_dsilu(fmaf(es, r[e] - em, et));
                        }
#pragma unroll
                        for (int e = 0; e < 4; ++e) { s0 += v[e]; s1 = fmaf(v[e], r[e] - em, s1); }
                    }
                }
                *reinterpret_cast<float4*>(sp) = make_float4(v[0], v[1], v[2], v[3]);
            }
            __syncwarp();
            // row-wise pass: whole 128-byte rows to global memory
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int row = 4 * i + lr;
                if (qv && mrow0 + row < p.Cout) {
                    float4 o = *reinterpret_cast<const float4*>(stg + row * 36 + lq * 4);
                    float* dst = p.out + (long long)(mrow0 + row) * p.out_sc + coff;
                    if (p.accumulate) { const float4 old = ld4(dst); o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w; }
                    st4(dst, o);
                }
            }
            __syncwarp();
        }
        if (mv && p.epi_mode != EPI_STORE && p.stat0 != nullptr && !(g.dbg & 128)) {
            atomicAdd(p.stat0 + co, (double)s0);
            atomicAdd(p.stat1 + co, (double)s1);
        }
        }
        // next tile: its operand stores reuse the staging tiles, its first MMA overwrites the accumulators just read
        if (!PERSIST) break;
        tc_fence_before();
        asm volatile("bar.sync 1, %0;" ::"r"(NPROD) : "memory");
        tile += gridDim.x; ++it;
        } while (tile < g.total);
    } else if (warp == NPW) {
        // ------------------------------ MMA issue (one thread) ------------------------------
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_tf32(BM, BN, 0, 1);             // A (weights) K-major, B (activations) MN-major
            int s = 0, sa = 0; uint32_t ph = 0, pha = 0;
            int tile = blockIdx.x;
            do {
            WF_TC_TILE_VARS(tile)
            (void)m0; (void)col0;
            for (int kc = 0; kc < KT; ++kc) {
                mbar_wait(full_a(sa), pha);
                mbar_wait(full_b(s), ph);
                tc_fence_after();
                const uint32_t a0 = smem_u32(smem_a + sa * A_BYTES);
                const uint32_t b_hi = smem_u32(smem + s * B_STAGE), b_lo = b_hi + B_HALF;
#pragma unroll
                for (int kk = 0; kk < KC / 8; ++kk) {
                    // 8 channels = 8 rows of 128 bytes per K step; 32-column blocks KC*128 bytes apart; groups of 4 rows 512 bytes apart
                    const uint64_t dbh = umma_desc_l(b_hi + kk * 1024, KC * 128, 512, 1), dbl = umma_desc_l(b_lo + kk * 1024, KC * 128, 512, 1);
                    const uint32_t accf = (kc | kk) != 0 ? 1u : 0u;
                    if (g.dbg & 2) continue;
                    for (int tj = 0; tj < ntile; ++tj) {          // the M tiles of this CTA share the activation descriptors
                        const uint32_t a_hi = a0 + tj * 2 * A_HALF, a_lo = a_hi + A_HALF;
                        const uint64_t dah = umma_desc(a_hi + kk * 2 * A_LBO, A_LBO, A_SBO), dal = umma_desc(a_lo + kk * 2 * A_LBO, A_LBO, A_SBO);
                        const uint32_t t_main = tmem_base + (uint32_t)(tj * 2 * g.bnp), t_cor = t_main + (uint32_t)g.bnp;
                        if (!(g.dbg & 16)) {
                            umma_tf32(t_cor, dal, dbh, idesc, accf);
                            umma_tf32(t_cor, dah, dbl, idesc, 1u);
                        }
                        umma_tf32(t_main, dah, dbh, idesc, accf);
                    }
                }
                umma_commit(empty(s));
                umma_commit(empty_a(sa));
                if (++s == STAGES) { s = 0; ph ^= 1u; }
                if (++sa == NSA) { sa = 0; pha ^= 1u; }
            }
            umma_commit(accum_bar);
            tile += gridDim.x;
            } while (PERSIST && tile < g.total);
        }
    } else {
        // ------------------------------ weight tiles: one 32 KB bulk copy per stage ------------------------------
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            int tile = blockIdx.x;
            do {
            WF_TC_TILE_VARS(tile)
            (void)m0; (void)col0;
            for (int kc = 0; kc < KT; ++kc) {
                mbar_wait(empty_a(s), ph ^ 1u);
                if (g.dbg & 4) { mbar_arrive(full_a(s)); if (++s == NSA) { s = 0; ph ^= 1u; } continue; }       // no weight copies
                mbar_arrive_expect_tx(full_a(s), ntile * 2 * A_HALF);
                for (int tj = 0; tj < ntile; ++tj) {
                    const float* wsrc = p.wtc + ((size_t)(mt0 + tj) * KT + kc) * (2 * A_HALF / 4);
                    bulk_g2s(smem_u32(smem_a + s * A_BYTES + tj * 2 * A_HALF), wsrc, 2 * A_HALF, full_a(s));
                }
                if (++s == NSA) { s = 0; ph ^= 1u; }
            }
            tile += gridDim.x;
            } while (PERSIST && tile < g.total);
        }
    }
#undef WF_TC_TILE_VARS
    tc_fence_before();
    __syncthreads();
    if (warp == NPW + 1) tmem_dealloc(tmem_base, g.tmem_cols);
    wf_bn_tail(p.tail);
}

// =========================================================================================================
// backward-weights:  dW[co][ci] += sum over this CTA's column range of G[co][col] * X'[ci][col]
// Both operands are K-major (K = columns) and are produced through registers; the column range is split across blockIdx.x.
// GPRO: prologue of G (PRO_NONE or PRO_BNBWD), XPRO/MASK: prologue of X.
// =========================================================================================================
constexpr int WG_BN_MAX = 256;
constexpr int WG_STAGES = 2;
constexpr int WG_STAGE_BYTES = 2 * A_HALF + 2 * WG_BN_MAX * KC * 4;

template <int GPRO, int XPRO, bool MASK>
__global__ void __launch_bounds__(NTHREADS, 1) pw_wgrad_tc_kernel(const WgradP p, int bn, int ntile_n, long long cols_per_split)
{
    wf_pdl_enter();
    extern __shared__ __align__(1024) uint8_t smem[];
    constexpr int B_HALF = WG_BN_MAX * KC * 4;
    constexpr int TCOLS = 2 * WG_BN_MAX;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + WG_STAGES * WG_STAGE_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * WG_STAGES + 1);
    const uint32_t bar0 = smem_u32(bars);
    auto full = [&](int s) { return bar0 + 8u * s; };
    auto empty = [&](int s) { return bar0 + 8u * (WG_STAGES + s); };
    const uint32_t accum_bar = bar0 + 8u * (2 * WG_STAGES);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = (blockIdx.y / ntile_n) * BM;
    const int c0 = (blockIdx.y % ntile_n) * bn;
    const long long NC = (long long)p.Pout * p.N;
    const long long kbegin = (long long)blockIdx.x * cols_per_split;
    long long kend = kbegin + cols_per_split;
    if (kend > NC) kend = NC;
    const int KT = kend > kbegin ? (int)((kend - kbegin + KC - 1) / KC) : 0;

    if (tid == 0) {
        for (int s = 0; s < WG_STAGES; ++s) { mbar_init(full(s), NPW); mbar_init(empty(s), 1); }
        mbar_init(accum_bar, 1);
        fence_mbar_init();
    }
    if (warp == NPW + 1) { tmem_alloc(smem_u32(tmem_slot), TCOLS); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < NPW) {
        // producers: a unit is 4 rows x 32 columns -- one warp-level request covers 4 rows x 128 contiguous bytes (the first version
        // staged 8 rows x 64 bytes because a no-swizzle K-major core matrix wants 8 rows per quarter warp: twice the tag look-ups
        // per byte, and the loads are what this kernel waits for, profiles/r2_pw_wgrad_tc_ablation.txt).  The images are
        // SWIZZLE_128B K-major: row r of a tile at r * 128 bytes, its 16-byte chunk c (4 columns) at ((c ^ (r & 7)) * 16), so the 8
        // lanes of a quarter warp (one row, chunks 0..7) hit 8 different bank groups.  Unit u = warp + 16 j: u < 32 are G' rows.
        static_assert(KC == 32 && BM / 4 == 2 * NPW, "a stage row is one 128-byte swizzle row; units 0, 1 of every warp are G' rows");
        constexpr int MAXG = (BM / 4 + WG_BN_MAX / 4 + NPW - 1) / NPW;
        const int r8 = lane >> 3;                              // row of this thread inside the unit (0..3)
        const int q = lane & 7;                                // column quad of this thread inside the 32-column stage
        const int ngroups = BM / 4 + (bn + 3) / 4;
        // per-unit row: element offset of the row inside its tensor, coefficients, validity
        unsigned rowoff[MAXG], moff[MAXG];
        float ca[MAXG], cb[MAXG], cc[MAXG], cd[MAXG];
        unsigned valid = 0;
#pragma unroll
        for (int j = 0; j < MAXG; ++j) {
            const int grp = warp + j * NPW;
            rowoff[j] = 0; moff[j] = 0; ca[j] = cb[j] = cc[j] = cd[j] = 0.f;
            if (grp < BM / 4) {
                const int co = m0 + grp * 4 + r8;
                if (co < p.Cout) {
                    valid |= 1u << j;
                    rowoff[j] = (unsigned)((long long)co * NC);
                    if (GPRO == PRO_BNBWD) { ca[j] = p.g_a[co]; cb[j] = p.g_b[co]; cc[j] = p.g_c[co]; cd[j] = p.g_d[co]; }
                }
            } else if (grp < ngroups) {
                const int ci = c0 + (grp - BM / 4) * 4 + r8;
                if (ci < p.Cin && (grp - BM / 4) * 4 + r8 < bn) {
                    valid |= 1u << j;
                    rowoff[j] = (unsigned)((long long)ci * p.in_sc);
                    if (MASK) moff[j] = (unsigned)((long long)ci * p.m_sc);
                    if (XPRO != PRO_NONE) { ca[j] = p.pro_a[ci]; cb[j] = p.pro_b[ci]; cd[j] = p.pro_d[ci]; }
                }
            }
        }
        // column state of this thread, advanced by KC columns per stage
        long long col = kbegin + q * 4;
        long long pos = 0, n = col;
        if (p.Pout > 1) { pos = col / p.N; n = col - pos * p.N; }
        long long b = n / WF_T; int t = (int)(n - b * WF_T);
        // byte offset inside a unit's 4 rows: (row & 7) = (warp & 1) * 4 + r8 for every unit of this warp (u = warp + 16 j)
        const uint32_t soff = (uint32_t)(r8 * 128 + ((q ^ ((warp & 1) * 4 + r8)) << 4));
        const int m_st = p.m_st;
        for (int kc = 0; kc < KT; ++kc) {
            const int s = kc % WG_STAGES;
            const uint32_t ph = (uint32_t)(kc / WG_STAGES) & 1u;
            const bool colv = col < kend;
            const float* gp = p.g + col;
            const float* gp2 = p.g2 + col;
            const float* xp = p.in + pos * p.in_sp + b * p.in_sb + t;
            const float* mp = p.mask + b * p.m_sb + (long long)t * m_st;
            // Unit j of a warp is row group (warp>>1) + 8j: j < GUNITS are G' rows, the rest X' rows -- a compile-time split.  All loads
            // of the chunk are issued first, as bare predicated loads with no arithmetic inside the guards, so that ptxas keeps up to
            // 2*MAXG 128-bit requests per thread in flight; the prologues run afterwards.  (With the math inside the guards every unit
            // waited for its own loads in turn: long-scoreboard stalls were 11.6 per issued instruction and the tensor pipe 9 % busy.)
            constexpr int GUNITS = (BM / 4) / NPW;
            float4 v[MAXG], w[MAXG];
#pragma unroll
            for (int j = 0; j < MAXG; ++j) {
                const bool ok = colv && ((valid >> j) & 1u);
                v[j] = f4zero(); w[j] = f4zero();
                if (j < GUNITS) {
                    if (ok) v[j] = ld4(gp + rowoff[j]);
                    if (GPRO == PRO_BNBWD) { if (ok) w[j] = ld4(gp2 + rowoff[j]); }
                } else {
                    if (ok) v[j] = ld4(xp + rowoff[j]);
                    if (MASK) {
                        if (m_st == 1) { if (ok) w[j] = ld4(mp + moff[j]); }
                        else { float mm = 0.f; if (ok) mm = mp[moff[j]]; w[j] = make_float4(mm, mm, mm, mm); }
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < MAXG; ++j) {
                const bool ok = colv && ((valid >> j) & 1u);
                float4 r;
                if (j < GUNITS) r = (GPRO == PRO_BNBWD) ? pro4<PRO_BNBWD, false>(v[j], w[j], ca[j], cb[j], cc[j], cd[j]) : v[j];
                else r = pro4<XPRO, MASK>(v[j], w[j], ca[j], cb[j], 0.f, cd[j]);
                v[j] = sel4(ok, r, f4zero());
            }
            mbar_wait(empty(s), ph ^ 1u);
            uint8_t* ah = smem + s * WG_STAGE_BYTES;
            uint8_t* bh = ah + 2 * A_HALF;
#pragma unroll
            for (int j = 0; j < MAXG; ++j) {
                const int grp = warp + j * NPW;
                if (grp < BM / 4) split_store(ah, ah + A_HALF, (uint32_t)(grp * 512) + soff, v[j]);
                else if (grp < ngroups) split_store(bh, bh + B_HALF, (uint32_t)((grp - BM / 4) * 512) + soff, v[j]);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(full(s));
            // advance KC = 32 columns: t += 12, b += 1 (mod 20), wrapping to the next position at n == N
            col += KC; n += KC; t += KC - WF_T; b += 1;
            if (t >= WF_T) { t -= WF_T; b += 1; }
            if (n >= p.N) { n -= p.N; pos += 1; b = n / WF_T; t = (int)(n - b * WF_T); }
        }

        // epilogue: thread = output channel co, columns = input channels ci; fp32 reductions into the gradient buffer
        if (KT > 0) {
            mbar_wait(accum_bar, 0);
            tc_fence_after();
            const int quarter = warp & 3, cgrp = warp >> 2;
            const int co = m0 + quarter * 32 + lane;
            const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16);
            for (int cb0 = 0; cb0 < bn; cb0 += 32) {
                if (((cb0 >> 5) % (NPW / 4)) != cgrp) continue;
                float acc[32], cor[32];
                tmem_ld32(trow + (uint32_t)cb0, acc);
                tmem_ld32(trow + (uint32_t)(WG_BN_MAX + cb0), cor);
                tmem_ld_wait();
                if (co < p.Cout) {
                    float* drow = p.dw + (size_t)co * p.Cin;
                    // 128-bit reductions (sm_90+ float4 atomicAdd) where the row pitch allows it: a thread owns 32 consecutive
                    // input channels of one output channel, i.e. 8 aligned quads instead of 32 scalar reductions
                    if ((p.Cin & 3) == 0 && (reinterpret_cast<uintptr_t>(p.dw) & 15) == 0) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const int ci = c0 + cb0 + j;
                            if (cb0 + j < bn && ci < p.Cin)
                                atomicAdd(reinterpret_cast<float4*>(drow + ci),
                                          make_float4(acc[j] + cor[j], acc[j + 1] + cor[j + 1], acc[j + 2] + cor[j + 2], acc[j + 3] + cor[j + 3]));
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int ci = c0 + cb0 + j;
                            if (cb0 + j < bn && ci < p.Cin) atomicAdd(drow + ci, acc[j] + cor[j]);
                        }
                    }
                }
            }
        }
    } else if (warp == NPW) {
        if (lane == 0 && KT > 0) {
            const uint32_t idesc = umma_idesc_tf32(BM, bn, 0, 0);
            const uint32_t tmem_cor = tmem_base + WG_BN_MAX;
            for (int kc = 0; kc < KT; ++kc) {
                const int s = kc % WG_STAGES;
                const uint32_t ph = (uint32_t)(kc / WG_STAGES) & 1u;
                mbar_wait(full(s), ph);
                tc_fence_after();
                const uint32_t a_hi = smem_u32(smem + s * WG_STAGE_BYTES), a_lo = a_hi + A_HALF;
                const uint32_t b_hi = a_hi + 2 * A_HALF, b_lo = b_hi + B_HALF;
#pragma unroll
                for (int kk = 0; kk < KC / 8; ++kk) {
                    // SWIZZLE_128B K-major: 8-row groups 1024 bytes apart, a K step of 8 columns = 32 bytes inside the 128-byte row
                    const uint64_t dah = umma_desc_l(a_hi + kk * 32, 16, 1024, 2), dal = umma_desc_l(a_lo + kk * 32, 16, 1024, 2);
                    const uint64_t dbh = umma_desc_l(b_hi + kk * 32, 16, 1024, 2), dbl = umma_desc_l(b_lo + kk * 32, 16, 1024, 2);
                    const uint32_t accf = (kc | kk) != 0 ? 1u : 0u;
                    umma_tf32(tmem_cor, dal, dbh, idesc, accf);
                    umma_tf32(tmem_cor, dah, dbl, idesc, 1u);
                    umma_tf32(tmem_base, dah, dbh, idesc, accf);
                }
                umma_commit(empty(s));
            }
            umma_commit(accum_bar);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == NPW + 1) tmem_dealloc(tmem_base, TCOLS);
}

// =========================================================================================================
// weight packing: reference [Cout][Cin] -> per (M tile, K chunk) shared-memory images, hi then lo, zero padded
// =========================================================================================================
__global__ void tc_pack_kernel(TcPackTable tab, const float* params, float* packed)
{
    wf_pdl_enter();
    const TcPackEntry e = tab.e[blockIdx.y];
    const int f_mt = (e.cout + BM - 1) / BM, f_kt = (e.cin + KC - 1) / KC;
    const int b_mt = (e.cin + BM - 1) / BM, b_kt = (e.cout + KC - 1) / KC;
    const long long nf = (long long)f_mt * f_kt * (BM * KC), nb = (long long)b_mt * b_kt * (BM * KC);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nf + nb; i += (long long)gridDim.x * blockDim.x) {
        const bool bwd = i >= nf;
        const long long k = bwd ? i - nf : i;
        const int kt = bwd ? b_kt : f_kt;
        const int within = (int)(k % (BM * KC));
        const long long blk = k / (BM * KC);
        const int mt = (int)(blk / kt), kc = (int)(blk % kt);
        // image order: [row group (16)][k quad (KC/4)][row in group (8)][k in quad (4)]
        const int c4 = within & 3, r8 = (within >> 2) & 7, kq = (within >> 5) % (KC / 4), rg = within / (8 * KC);
        const int row = mt * BM + rg * 8 + r8, kk = kc * KC + kq * 4 + c4;
        float v = 0.f;
        if (!bwd) { if (row < e.cout && kk < e.cin) v = params[e.param_off + (long long)row * e.cin + kk]; }     // M = cout, K = cin
        else { if (row < e.cin && kk < e.cout) v = params[e.param_off + (long long)kk * e.cin + row]; }         // M = cin,  K = cout
        float hi, lo;
        tf32_split(v, hi, lo);
        float* dst = packed + (bwd ? e.bwd_off : e.fwd_off) + blk * (2 * BM * KC);
        dst[within] = hi;
        dst[BM * KC + within] = lo;
    }
}

constexpr int SMEM_MAX = 227 * 1024;

template <int PRO, bool MASK, bool PERSIST>
cudaError_t launch_conv_p(const ConvP& p, const TcGeom& g, dim3 grid, int smem, cudaStream_t st)
{
    static WfSmemOptIn optin;
    if (cudaError_t e = wf_smem_optin(optin, pw_tc_kernel<PRO, MASK, PERSIST>, SMEM_MAX)) return e;
    wf_launch_pdl(pw_tc_kernel<PRO, MASK, PERSIST>, dim3(grid), dim3(NTHREADS), smem, st, p, g);
    return cudaGetLastError();
}
template <int PRO, bool MASK>
cudaError_t launch_conv_t(const ConvP& p, const TcGeom& g, dim3 grid, int smem, cudaStream_t st)
{
    return (int)grid.x < g.total ? launch_conv_p<PRO, MASK, true>(p, g, grid, smem, st) : launch_conv_p<PRO, MASK, false>(p, g, grid, smem, st);
}

template <int GPRO, int XPRO, bool MASK>
cudaError_t launch_wgrad_t(const WgradP& p, int bn, int nt, long long per, dim3 grid, cudaStream_t st)
{
    constexpr int smem = WG_STAGES * WG_STAGE_BYTES + (2 * WG_STAGES + 1) * 8 + 16;
    static WfSmemOptIn optin;
    if (cudaError_t e = wf_smem_optin(optin, pw_wgrad_tc_kernel<GPRO, XPRO, MASK>, smem)) return e;
    wf_launch_pdl(pw_wgrad_tc_kernel<GPRO, XPRO, MASK>, dim3(grid), dim3(NTHREADS), smem, st, p, bn, nt, per);
    return cudaGetLastError();
}

}  // namespace

long long wf_tc_pack_floats(int m, int k) { return (long long)((m + BM - 1) / BM) * ((k + KC - 1) / KC) * (2 * BM * KC); }

cudaError_t wf_launch_tc_pack(const TcPackTable& tab, const float* params, float* packed, cudaStream_t st)
{
    if (tab.n == 0) return cudaSuccess;
    dim3 grid(64, tab.n);
    wf_launch_pdl(tc_pack_kernel, dim3(grid), dim3(256), 0, st, tab, params, packed);
    return cudaGetLastError();
}

cudaError_t wf_launch_tc_conv(const ConvP& p, int num_sms, cudaStream_t st)
{
    const long long NC = (long long)p.Pout * p.N;
    const long long mt = (p.Cout + BM - 1) / BM;
    // M tiles per CTA.  WF_TC_MP=2 lets two 128-row tiles share one staged activation tile (2 x 2 accumulators x bn <= 512 TMEM
    // columns then limit the column tile to 128 and the ring to two stages).  Measured at B = 1024: forward 1.54 ms vs 1.31 ms,
    // backward-data 1.67 vs 1.47 ms with one tile per CTA -- halving the re-staged activation work does not pay for the narrower
    // tiles and the shallower ring, i.e. the one-shot CTA (fill, drain, serial epilogue), not the producers, is what bounds the kernel.
    static const int mp_env = [] { const char* e = std::getenv("WF_TC_MP"); const int v = e ? std::atoi(e) : 0; return v; }();
    const int mp = (mp_env == 2 && mt >= 2) ? 2 : 1;
    const int bn_max = mp == 2 ? 128 : 256;
    // column-tile width: the multiple of 32 in [64, bn_max] that minimises (rounds of tiles per SM) x (cost of one tile)
    int best_bn = bn_max; long long best_cost = -1;
    for (int bn = bn_max; bn >= 64; bn -= 32) {       // whole 32-column blocks of the MN-major activation image
        const long long tiles = ((mt + mp - 1) / mp) * ((NC + bn - 1) / bn);
        const long long cost = ((tiles + num_sms - 1) / num_sms) * (bn + 64);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_bn = bn; }
    }
    static const int bn_env = [] { const char* e = std::getenv("WF_TC_BN"); return e ? std::atoi(e) : 0; }();       // experiments only
    if (bn_env >= 64 && bn_env <= bn_max && bn_env % 32 == 0) best_bn = bn_env;
    TcGeom g{};
    static const int dbg_env = [] { const char* e = std::getenv("WF_TC_DBG"); return e ? std::atoi(e) : 0; }();       // measurements: 1 no activation loads, 2 no MMAs, 4 no weight copies, 8 no operand stores, 16 main product only, 32 no epilogue memory traffic, 64 no epilogue, 128 no statistics reductions
    g.dbg = dbg_env;
    g.mp = mp; g.mtiles = (int)mt;
    g.bn = best_bn;
    g.bnp = (best_bn + 31) / 32 * 32;
    g.tmem_cols = 32;
    while (g.tmem_cols < mp * 2 * g.bnp) g.tmem_cols *= 2;
    const int a_bytes = mp * 2 * A_HALF, b_stage = 2 * KC * g.bn * 4;
    const int budget = (g.tmem_cols <= 256 && mp == 1) ? 100 * 1024 : SMEM_MAX - 512;       // narrow tiles: leave room for a second CTA per SM
    g.nst = budget / (a_bytes + b_stage);
    if (g.nst > 4) g.nst = 4;
    if (g.nst < 2) g.nst = 2;
    g.nsa = g.nst + (budget - g.nst * (a_bytes + b_stage)) / a_bytes;      // what is left deepens the weight ring
    if (g.nsa > 4) g.nsa = 4;
    if (g.nsa < g.nst) g.nsa = g.nst;
    const int smem = g.nst * b_stage + g.nsa * a_bytes + (2 * g.nst + 2 * g.nsa + 1) * 8 + 16;
    g.mgroups = (int)((mt + mp - 1) / mp);
    const long long total = ((NC + g.bn - 1) / g.bn) * g.mgroups;
    if (total >= (1LL << 31)) return cudaErrorInvalidValue;
    g.total = (int)total;
    // persistent only when the epilogue's staging tiles (16 warps x 4.5 KB) stay inside the activation ring: the weight copies of the
    // next tile start while the epilogue of this one still runs.  WF_TC_PERSIST=0: one tile per CTA (measurements)
    static const bool persist_env = [] { const char* e = std::getenv("WF_TC_PERSIST"); return !(e && e[0] == '0'); }();
    // ... and only where a CTA would get at least four tiles (the attention projections: 15 positions x 20 480 columns)
    const bool persist = persist_env && (long long)g.nst * b_stage >= (long long)NPW * 32 * 36 * 4 && total >= 4LL * num_sms;
    dim3 grid((unsigned)(persist ? num_sms : total));
    const bool mask = p.pro_mode == PRO_BNSILU && p.mask != nullptr;
    switch (p.pro_mode) {
        case PRO_NONE: return launch_conv_t<PRO_NONE, false>(p, g, grid, smem, st);
        case PRO_BNSILU: return mask ? launch_conv_t<PRO_BNSILU, true>(p, g, grid, smem, st) : launch_conv_t<PRO_BNSILU, false>(p, g, grid, smem, st);
        case PRO_AFFINE: return launch_conv_t<PRO_AFFINE, false>(p, g, grid, smem, st);
        default: return launch_conv_t<PRO_BNBWD, false>(p, g, grid, smem, st);
    }
}

cudaError_t wf_launch_tc_wgrad(const WgradP& p, int num_sms, cudaStream_t st)
{
    const long long NC = (long long)p.Pout * p.N;
    // 32-bit element offsets inside the kernel
    if ((long long)p.Cout * NC >= (1LL << 31) || (long long)p.Cin * p.in_sc >= (1LL << 31) || (long long)p.Cin * p.m_sc >= (1LL << 31))
        return cudaErrorInvalidValue;
    // input-channel tile: an even split of Cin into <= 256-wide pieces, rounded up to the UMMA N granularity (16)
    const int nt = (p.Cin + WG_BN_MAX - 1) / WG_BN_MAX;
    int bn = ((p.Cin + nt - 1) / nt + 15) / 16 * 16;
    if (bn < 16) bn = 16;
    const int mtiles = (p.Cout + BM - 1) / BM;
    const int tiles = mtiles * nt;
    static const int rounds = [] { const char* e = std::getenv("WF_TC_WGRAD_ROUNDS"); const int v = e ? std::atoi(e) : 0; return v > 0 ? v : 1; }();       // measured: 1 round 2.16 ms, 2 rounds 2.27 ms, 3 rounds 2.39 ms per step
    long long splits = ((long long)rounds * num_sms) / tiles;            // full rounds of CTAs, never a ragged extra one
    const long long max_splits = (NC + 8 * KC - 1) / (8 * KC);           // at least 8 stages per CTA
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    long long per = (NC + splits - 1) / splits;
    per = (per + KC - 1) / KC * KC;
    splits = (NC + per - 1) / per;
    dim3 grid((unsigned)splits, (unsigned)tiles);
    const bool mask = p.pro_mode == PRO_BNSILU && p.mask != nullptr;
    if (p.g_pro == PRO_NONE) {
        if (p.pro_mode == PRO_NONE) return launch_wgrad_t<PRO_NONE, PRO_NONE, false>(p, bn, nt, per, grid, st);
        return cudaErrorInvalidValue;                                       // only the self-test uses a raw G
    }
    switch (p.pro_mode) {
        case PRO_NONE: return launch_wgrad_t<PRO_BNBWD, PRO_NONE, false>(p, bn, nt, per, grid, st);
        case PRO_BNSILU: return mask ? launch_wgrad_t<PRO_BNBWD, PRO_BNSILU, true>(p, bn, nt, per, grid, st)
                                     : launch_wgrad_t<PRO_BNBWD, PRO_BNSILU, false>(p, bn, nt, per, grid, st);
        case PRO_AFFINE: return launch_wgrad_t<PRO_BNBWD, PRO_AFFINE, false>(p, bn, nt, per, grid, st);
        default: return cudaErrorInvalidValue;
    }
}
