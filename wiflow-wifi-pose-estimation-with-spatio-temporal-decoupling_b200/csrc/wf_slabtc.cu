// TMA-fed tcgen05 / TMEM kernels for the position-tap convolutions of the conv stack
//   (1x3) convs, stride (1,1) / (1,2), and their strided 1x1 shortcuts (models/convnet.py:11-12,17,22,27,48,53,58,63)
// forward and backward-data (slab_tc_kernel).  They replace the warp-level mma.sync kernels of wf_slide.cu for every layer
// with Cin in {8,16,32,64} and 8 <= Cout <= 64.
//
// Roles of the operands are swapped with respect to the pointwise kernels of wf_tc.cu: the 128 COLUMNS (n = b*20 + t) of a
// tile are the UMMA M dimension, the output channels (8..64, padded to a multiple of 16) are N, and every tap contributes
// Cin/8 K-steps.  In the internal layout [channel][position][n] the activation tile of one input position ("slab") is a
// [Cin][128] matrix whose contiguous dimension is M, i.e. an MN-major operand.  For 32-bit (tf32) MN-major operands the only
// layout the tensor core accepts is SWIZZLE_128B_BASE32B (rows of 128 bytes = 32 columns of one channel, 32-byte chunks XORed
// with the row index mod 4, K groups of 4 rows): four TMA boxes {32 n, Cin, PBI positions} with
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B land the slab in exactly that layout -- no register transpose.  (Plain SWIZZLE_128B and the
// no-swizzle MN-major forms make the MMA read zeros: that is what round 1 ran into.)
//
// A conv with taps is evaluated as a SCATTER over input slabs: when slab q is in shared memory, tap t adds W[t] * X'[q] to the
// TMEM accumulator of output position p(q, t); an output position is complete after its last contributing slab.  A slab is
// therefore loaded from HBM once, transformed once, used by all taps and then dropped: the ring holds only the slabs in
// flight (2-3 stages), not the tap window, which is what lets 64-channel layers keep 128-column tiles plus all tap weights
// (hi and lo, 96 KB) in the 227 KB of shared memory.
//
// Warp roles (640 threads): warp 16 lane 0 issues the TMA loads, warps 17-19 issue tcgen05.mma (warp-uniform, one elected lane), warps 0-15 are workers:
//   * transform: the raw slab (TMA) is rewritten IN PLACE as the activated operand -- BatchNorm+SiLU+Dropout2d, BatchNorm only,
//     or BatchNorm-backward of (dy, raw) -- split into tf32 hi / lo images (3xTF32, fp32 parity: a*b ~ a_lo*b_hi + a_hi*b_lo +
//     a_hi*b_hi).  Addresses are linear (the swizzle permutes 32-byte chunks inside a row, a row is one channel), so the
//     128-bit shared-memory accesses are conflict free and no index arithmetic is left in the loop;
//   * epilogue: tcgen05.ld (thread = column n, registers = channels) -> bias / SiLU'*mask / BatchNorm sums -> coalesced stores
//     (a warp writes 128 contiguous bytes per channel).  Per-channel sums stay in registers for the whole walk of the CTA and
//     are reduced across lanes once, at the end.
// Work is the flattened list of (column tile, output position) units cut into one contiguous run per CTA (persistent, 1 CTA/SM).
#include <cuda.h>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include "wf_tc.cuh"
#include "wf_common.cuh"
#include "wf_elem.h"

namespace {

using namespace tc;

constexpr int TILE = 128;                 // columns per tile = UMMA M
constexpr int NWORK = 512;                // worker threads (warps 0-15)
constexpr int NWW = NWORK / 32;
constexpr int NI = 3;                     // MMA-issuing warps
constexpr int NTHR = NWORK + 32 + 32 * NI; // + TMA warp + MMA-issuing warps
constexpr int MAXACC = 16;                // accumulator slots (barriers) at most
constexpr int NGD = 4;                    // ring of "group done" barriers
constexpr int MAXNS = 4;
constexpr int MAXJ = 4;                   // float4 chunks per worker thread and stage (R = 64 rows)

struct SlabGeom {
    int R, PBI, NS, NPAD, NACC, tmem_cols, ntiles;
    int half_bytes;                       // one of hi/lo of a stage: 4 * R * 128
    int w_img_bytes;                      // one weight image: ntaps * 2*NPAD rows (hi rows, lo rows per tap) * Cin * 4
    int w_bytes;                          // images resident in shared memory: 1 (per-tap MMAs) or 2 (stacked-tap MMAs need the zero-padded image)
    int W;                                // accumulator slots released per mbarrier test of the MMA thread
    int stack;                            // all taps of a slab in one MMA pair (N = ntaps * 2*NPAD <= 256)
    int dpmin, dpmax;
    int lbo_mn, sbo_mn;                   // MN-major SW128_BASE32B descriptor strides: 32-column blocks / 4-row K groups of the activation operand (bytes)
    long long units;                      // ntiles * Pout
    int dbg;                              // measurement switches (WF_SLABTC_DBG): 1 no transform, 2 no MMA, 4 no epilogue memory traffic, 8 hi*hi only, 16 no TMA, 32 no TMEM load
};

struct SlabGeom;
// timeline probe of CTA 0 (WF_SLABTC_DBG bit 512): globaltimer stamps of the pipeline's milestones, read by the self-test
__device__ unsigned long long g_ts[32];
// per-CTA stamps of two consecutive launches: [0/3] before the programmatic-dependency wait, [1/4] kernel body entry, [2/5] final sync
__device__ unsigned long long g_cta[6][160];
__device__ __forceinline__ void stamp(const SlabGeom& g, int slot)
{
    if (g.dbg & 512) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (blockIdx.x == 0 && slot < 14 && g_ts[slot] == 0) g_ts[slot] = t;
        const int k = slot == 14 ? 0 : slot == 0 ? 1 : slot == 13 ? 2 : -1;
        if (k >= 0 && blockIdx.x < 160) { if (g_cta[k][blockIdx.x] == 0) g_cta[k][blockIdx.x] = t; else g_cta[k + 3][blockIdx.x] = t; }
    }
}

// ---- TMA tensor load (3-D tile), completion on an mbarrier ----
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}

// shared-memory matrix descriptor with a layout type (0 = no swizzle, 1 = SWIZZLE_128B_BASE32B, 2 = SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_desc_l(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(layout & 7u) << 61;
    return d;
}

// one lane polls the mbarrier, the warp reconverges behind it: 32 lanes spinning on try_wait only load the barrier unit
__device__ __forceinline__ void warp_wait(uint32_t bar, uint32_t parity, int lane)
{
    if (lane == 0) mbar_wait(bar, parity);
    __syncwarp();
}

// D[tmem] += A[smem] * B[smem], issued by the lanes whose `lead` is non-zero (one per warp); operands are warp-uniform
__device__ __forceinline__ void umma_tf32_pred(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t lead)
{
    asm volatile("{\n\t.reg .pred pl, pa;\n\tsetp.ne.b32 pl, %4, 0;\n\tsetp.eq.b32 pa, 0, 0;\n\t@pl tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, pa;\n\t}"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(lead) : "memory");
}

// TMEM -> registers, 32 lanes x NV consecutive columns (thread = lane)
template <int NV> __device__ __forceinline__ void tmem_ldn(uint32_t taddr, float* v);
template <> __device__ __forceinline__ void tmem_ldn<2>(uint32_t taddr, float* v)
{
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(taddr) : "memory");
}
template <> __device__ __forceinline__ void tmem_ldn<4>(uint32_t taddr, float* v)
{
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
}
template <> __device__ __forceinline__ void tmem_ldn<8>(uint32_t taddr, float* v)
{
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
}
template <> __device__ __forceinline__ void tmem_ldn<16>(uint32_t taddr, float* v)
{
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr) : "memory");
}

// zero NV consecutive TMEM columns of this warp's 32 lanes
template <int NV> __device__ __forceinline__ void tmem_zero(uint32_t taddr);
template <> __device__ __forceinline__ void tmem_zero<2>(uint32_t taddr)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %1};" ::"r"(taddr), "r"(0u) : "memory");
}
template <> __device__ __forceinline__ void tmem_zero<4>(uint32_t taddr)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %1, %1, %1};" ::"r"(taddr), "r"(0u) : "memory");
}
template <> __device__ __forceinline__ void tmem_zero<8>(uint32_t taddr)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr), "r"(0u) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float4 lds4(uint32_t a)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts4(uint32_t a, float4 v)
{
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// ---- position bookkeeping shared by the three roles (all deterministic functions of the launch geometry) ----
struct Seg { int ct, oa, ob, qa, qb; };
__device__ __forceinline__ void seg_make(const ConvP& p, const SlabGeom& g, long long u, long long u1, Seg& s)
{
    s.ct = (int)(u / p.Pout);
    s.oa = (int)(u - (long long)s.ct * p.Pout);
    const long long left = u1 - u;
    s.ob = (long long)(p.Pout - s.oa) < left ? p.Pout : s.oa + (int)left;
    int lo = s.oa * p.pmul + g.dpmin, hi = (s.ob - 1) * p.pmul + g.dpmax;
    if (p.pdiv == 2) { lo = (lo + 1) >> 1; hi = hi >> 1; }
    s.qa = lo < 0 ? 0 : lo;
    s.qb = hi > p.Pin - 1 ? p.Pin - 1 : hi;
}
// last input slab contributing to output position pp (clipped to the segment's slab range)
__device__ __forceinline__ int q_last(int pmul, int pdiv, int dpmax, int qb, int pp)
{
    int q = pp * pmul + dpmax;
    if (pdiv == 2) q >>= 1;
    return q > qb ? qb : q;
}
// output position that slab q feeds through the tap with offset dp, or -1
__device__ __forceinline__ int out_pos(int pmul, int pdiv, int oa, int ob, int q, int dp)
{
    int num = (pdiv == 2 ? 2 * q : q) - dp;
    if (num < 0) return -1;
    if (pmul == 2) { if (num & 1) return -1; num >>= 1; }
    return (num >= oa && num < ob) ? num : -1;
}
__device__ __forceinline__ bool has_contrib(const ConvP& p, int pp)
{
    if (p.pdiv == 1) return true;                 // the centre tap (dp = 0) always lands inside the input
    for (int t = 0; t < p.ntaps; ++t) {
        const int num = pp * p.pmul + p.dp[t];
        if (num < 0 || (num & 1)) continue;
        if ((num >> 1) < p.Pin) return true;
    }
    return false;
}

template <int PRO>
__device__ __forceinline__ float pro1(float x, float x2, float mk, float4 c4)
{
    if (PRO == PRO_BNSILU) return wf_silu(fmaf(c4.x, x - c4.w, c4.y)) * mk;
    if (PRO == PRO_AFFINE) return fmaf(c4.x, x - c4.w, c4.y);
    if (PRO == PRO_BNBWD) return fmaf(c4.x, x, fmaf(c4.y, x2 - c4.w, c4.z));
    return x;
}

// per-thread state of the epilogue that lives for the whole kernel
template <int CH>
struct EpiThread {
    uint32_t t_lane;          // TMEM address of this warp's lane quarter
    int ec0;                  // first output channel of this warp
    int lane;
    float s0[CH], s1[CH];     // running per-channel sums (BatchNorm statistics / BatchNorm-backward sums)
};

// epilogue of output positions [p_from, p_to) of one column tile; slot0 = accumulator slot of p_from.
// Every global load of a position (raw tensor of the layer below, previous gradient) is issued before the accumulator is
// read, so the tcgen05.ld latency and the memory latency overlap; offsets are 32-bit (the host declines larger tensors).
// Segment-invariant values of a worker thread's epilogue (a segment = one column tile)
struct EpiSeg { int obase_n; bool nv; const float* mkp; };

template <int EPI, bool ACC, int CH>
__device__ __forceinline__ void epi_run(const ConvP& p, const SlabGeom& g, EpiThread<CH>& e, const float* tab, uint32_t acc_empty0,
                                        int r_from, int p_from, int cnt, const EpiSeg& sg, bool use_tmem)
{
    // A pass handles PB positions x SUB channels per thread (8 values): all its global loads and TMEM loads are issued before the
    // first wait, so thin layers (2-4 channels per warp, hundreds of positions) pay one TMEM round trip per 2-4 positions.
    // r_from = running index of position p_from in this CTA's walk: slot = NACC-1 - (r mod NACC); only positions with
    // r mod W == W-1 signal their barrier (the MMA issuers wait on exactly those: in-order draining covers the others).
    constexpr int SUB = CH > 8 ? 8 : CH;
    constexpr int PB = 8 / SUB;
    const int out_sc = (int)p.out_sc, out_sp = (int)p.out_sp;
    const int NACCm = g.NACC - 1, Wm = g.W - 1, slotw = 2 * g.NPAD;
    const bool mem = sg.nv && !(g.dbg & 4);
    const bool every = p.pdiv == 1;               // every position receives contributions (centre tap)
    const bool tm = use_tmem && !(g.dbg & 32);
    for (int i0 = 0; i0 < cnt; i0 += PB) {
        const int pp = p_from + i0, r0 = r_from + i0;
        const int nb = cnt - i0 < PB ? cnt - i0 : PB;
        bool contrib[PB];
#pragma unroll
        for (int j = 0; j < PB; ++j) contrib[j] = j < nb && tm && (every || has_contrib(p, pp + j));
#pragma unroll
        for (int c0 = 0; c0 < CH; c0 += SUB) {
            const int ch0 = e.ec0 + c0;
            float rw[PB][SUB], old[PB][SUB], mk[SUB], v[PB][SUB], cr[PB][SUB];
            const int off00 = ch0 * out_sc + pp * out_sp + sg.obase_n;
#pragma unroll
            for (int c = 0; c < SUB; ++c) mk[c] = (EPI == EPI_DSILU && sg.mkp && mem) ? __ldg(sg.mkp + (ch0 + c) * (int)p.em_sc) : 1.f;
#pragma unroll
            for (int j = 0; j < PB; ++j)
#pragma unroll
                for (int c = 0; c < SUB; ++c) {
                    rw[j][c] = 0.f; old[j][c] = 0.f;
                    if (mem && j < nb) {
                        if (EPI == EPI_DSILU || EPI == EPI_DAFF) rw[j][c] = __ldg(p.eraw + off00 + j * out_sp + c * out_sc);
                        if (ACC) old[j][c] = p.out[off00 + j * out_sp + c * out_sc];
                    }
                }
#pragma unroll
            for (int j = 0; j < PB; ++j) {
                if (contrib[j]) {
                    const uint32_t ta = e.t_lane + (uint32_t)((NACCm - ((r0 + j) & NACCm)) * slotw + ch0);
                    tmem_ldn<SUB>(ta, v[j]); tmem_ldn<SUB>(ta + g.NPAD, cr[j]);
                } else {
#pragma unroll
                    for (int c = 0; c < SUB; ++c) { v[j][c] = 0.f; cr[j][c] = 0.f; }
                }
            }
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < PB; ++j)
                if (contrib[j]) {                                         // accumulators are zero-initialised: every MMA accumulates
                    const uint32_t ta = e.t_lane + (uint32_t)((NACCm - ((r0 + j) & NACCm)) * slotw + ch0);
                    tmem_zero<SUB>(ta); tmem_zero<SUB>(ta + g.NPAD);
                }
            if (c0 + SUB >= CH) {                 // last TMEM access of these positions: hand the accumulator slots back
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (e.lane == 0) {
#pragma unroll
                    for (int j = 0; j < PB; ++j)
                        if (j < nb && ((r0 + j) & Wm) == Wm) mbar_arrive(acc_empty0 + 8u * (NACCm - ((r0 + j) & NACCm)));
                }
            }
            if (mem) {
#pragma unroll
                for (int j = 0; j < PB; ++j) {
                    if (j < nb) {
#pragma unroll
                        for (int c = 0; c < SUB; ++c) {
                            float x = v[j][c] + cr[j][c] + tab[ch0 + c];
                            if (ACC) x += old[j][c];
                            if (EPI == EPI_STATS) { e.s0[c0 + c] += x; e.s1[c0 + c] = fmaf(x, x, e.s1[c0 + c]); }
                            else if (EPI == EPI_DSILU || EPI == EPI_DAFF) {
                                const float d = rw[j][c] - tab[192 + ch0 + c];
                                if (EPI == EPI_DSILU) x = x * mk[c] * wf_dsilu(fmaf(tab[64 + ch0 + c], d, tab[128 + ch0 + c]));
                                e.s0[c0 + c] += x; e.s1[c0 + c] = fmaf(x, d, e.s1[c0 + c]);
                            }
                            p.out[off00 + j * out_sp + c * out_sc] = x;
                        }
                    }
                }
            }
        }
    }
}

// =========================================================================================================
// forward / backward-data
// =========================================================================================================
template <int PRO, bool MASK, int CH>
__global__ void __launch_bounds__(NTHR, 1) slab_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                          const ConvP p, const SlabGeom g)
{
    if (threadIdx.x == 0) stamp(g, 14);
    wf_pdl_enter();
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) stamp(g, 0);
    const int stage_bytes = 2 * g.half_bytes;
    uint8_t* wsm = smem + g.NS * stage_bytes;                                   // weights: hi images of all taps, then lo
    float* tab = reinterpret_cast<float*>(wsm + g.w_bytes);            // epilogue tables: bias, e_scale, e_shift, e_mean [64] each
    uint64_t* bars = reinterpret_cast<uint64_t*>(tab + 4 * 64);
    const uint32_t bar0 = smem_u32(bars);
    // barrier map: raw_full[NS] | op_full[NS] | slab_empty[NS] | grp_done[NGD] | acc_empty[MAXACC] | w_full
    auto raw_full = [&](int s) { return bar0 + 8u * s; };
    auto op_full = [&](int s) { return bar0 + 8u * (MAXNS + s); };
    auto slab_empty = [&](int s) { return bar0 + 8u * (2 * MAXNS + s); };
    auto grp_done = [&](int s) { return bar0 + 8u * (3 * MAXNS + s); };
    const uint32_t acc_empty0 = bar0 + 8u * (3 * MAXNS + NGD);
    const uint32_t w_full = bar0 + 8u * (3 * MAXNS + NGD + MAXACC);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * MAXNS + NGD + MAXACC + 1);

    if (tid == 0) {
        for (int s = 0; s < g.NS; ++s) { mbar_init(raw_full(s), 1); mbar_init(op_full(s), NWW); mbar_init(slab_empty(s), NI); }
        for (int s = 0; s < NGD; ++s) mbar_init(grp_done(s), NI);
        for (int s = 0; s < g.NACC; ++s) mbar_init(acc_empty0 + 8u * s, NWW);
        mbar_init(w_full, 1);
        fence_mbar_init();
    }
    if (warp == NWW) { tmem_alloc(smem_u32(tmem_slot), g.tmem_cols); tmem_relinquish(); }
    for (int i = tid; i < 4 * 64; i += NTHR) {
        const int which = i >> 6, c = i & 63;
        float v = 0.f;
        if (c < p.Cout) {
            if (which == 0) v = p.bias ? p.bias[c] : 0.f;
            else if (which == 1) v = (p.epi_mode == EPI_DSILU) ? p.e_scale[c] : 0.f;
            else if (which == 2) v = (p.epi_mode == EPI_DSILU) ? p.e_shift[c] : 0.f;
            else v = (p.epi_mode == EPI_DSILU || p.epi_mode == EPI_DAFF) ? p.e_mean[c] : 0.f;
        }
        tab[i] = v;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (tid == 0) stamp(g, 1);
    if (warp < NWW) {                      // zero every accumulator slot: all MMAs accumulate, a slot is re-zeroed by the epilogue that drains it
        constexpr int SUB = CH > 8 ? 8 : CH;
        const uint32_t t0 = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * CH);
        for (int sl = 0; sl < g.NACC; ++sl)
#pragma unroll
            for (int c0 = 0; c0 < CH; c0 += SUB) { tmem_zero<SUB>(t0 + sl * 2 * g.NPAD + c0); tmem_zero<SUB>(t0 + sl * 2 * g.NPAD + g.NPAD + c0); }
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (tid == 0) stamp(g, 2);
    // this CTA's run of (column tile, output position) units
    const long long u0 = g.units * blockIdx.x / gridDim.x, u1 = g.units * (blockIdx.x + 1) / gridDim.x;
    const int R = g.R, PBI = g.PBI, NS = g.NS, NACCm = g.NACC - 1;
    const int pmul = p.pmul, pdiv = p.pdiv;
    const uint32_t smem0 = smem_u32(smem);

    if (warp < NWW) {
        // =============================== workers: transform + epilogue ===============================
        const int NJ = (32 * R + NWORK - 1) / NWORK;              // float4 chunks per thread and stage (R = 64: 4)
        const int ch16 = tid & 7;
        const int row = (tid >> 3) & (R - 1);                     // the same row (channel, position in group) for every chunk of a thread
        const int cin_c = row & (p.Cin - 1);
        float4 co4 = make_float4(1.f, 0.f, 0.f, 0.f);
        if (PRO != PRO_NONE) co4 = make_float4(p.pro_a[cin_c], p.pro_b[cin_c], PRO == PRO_BNBWD ? p.pro_c[cin_c] : 0.f, p.pro_d[cin_c]);
        const int nq = (ch16 ^ ((row & 3) << 1)) << 2;            // column of this thread's chunk inside its 32-column block
        // epilogue mapping: TMEM lane quarter = warp mod 4, channels [ec0, ec0 + CH)
        EpiThread<CH> et;
        et.lane = lane;
        et.ec0 = (warp >> 2) * CH;
        et.t_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
#pragma unroll
        for (int c = 0; c < CH; ++c) { et.s0[c] = 0.f; et.s1[c] = 0.f; }
        const bool ch_ok = et.ec0 < p.Cout;                       // Cout = 8 with 16 worker warps: CH = 2 covers it exactly
        const int epi = p.epi_mode;
        const bool acc = p.accumulate != 0;

        // running state of the walk (kept incrementally: a worker warp retires one dependent instruction every ~5 cycles, so
        // every division / modulo / loop in the per-group path is paid 16 times per group)
        int st = 0; uint32_t ph = 0;                  // operand stage of the current group and its barrier phase
        int gcount = 0;                               // groups done (timeline probe)
        int gd = 0; uint32_t gdph = 0;                // "group done" barrier of the current group and its phase
        int rbase = 0;                                // output positions of earlier segments (accumulator slot numbering)
        struct Pend { int cnt, p_from, r_from; uint32_t bar, par; EpiSeg sg; } pend;
        pend.cnt = 0;

        auto run_epilogue = [&](const Pend& e, bool ut) {
            if (ut) { warp_wait(e.bar, e.par, lane); tc_fence_after(); if (tid == 0 && gcount == 7) stamp(g, 9); }
            switch (epi) {
                case EPI_STORE:
                    if (acc) epi_run<EPI_STORE, true, CH>(p, g, et, tab, acc_empty0, e.r_from, e.p_from, e.cnt, e.sg, ut);
                    else epi_run<EPI_STORE, false, CH>(p, g, et, tab, acc_empty0, e.r_from, e.p_from, e.cnt, e.sg, ut);
                    break;
                case EPI_STATS: epi_run<EPI_STATS, false, CH>(p, g, et, tab, acc_empty0, e.r_from, e.p_from, e.cnt, e.sg, ut); break;
                case EPI_DSILU: epi_run<EPI_DSILU, false, CH>(p, g, et, tab, acc_empty0, e.r_from, e.p_from, e.cnt, e.sg, ut); break;
                default: epi_run<EPI_DAFF, false, CH>(p, g, et, tab, acc_empty0, e.r_from, e.p_from, e.cnt, e.sg, ut); break;
            }
        };

        for (long long u = u0; u < u1;) {
            Seg s; seg_make(p, g, u, u1, s);
            const int n0 = s.ct * TILE;
            // per-segment values: dropout-mask values of this thread's chunks (Dropout2d: one value per (window, channel)) ...
            float mk[MAXJ];
#pragma unroll
            for (int j = 0; j < MAXJ; ++j) {
                mk[j] = 1.f;
                if (MASK && j < NJ) {
                    const int id = tid + NWORK * j;
                    const int n = n0 + (id / (8 * R)) * 32 + nq;
                    if (n < p.N && id < 32 * R) mk[j] = p.mask[(long long)(n / WF_T) * p.m_sb + (long long)cin_c * p.m_sc];
                }
            }
            // ... and the epilogue's column: thread = TMEM lane = column n of the tile
            EpiSeg sg;
            {
                const int n = n0 + (warp & 3) * 32 + lane;
                const int b = n / WF_T, t = n - b * WF_T;
                sg.nv = n < p.N && ch_ok;
                sg.obase_n = b * (int)p.out_sb + t;
                sg.mkp = (epi == EPI_DSILU && p.emask && sg.nv) ? p.emask + (long long)b * p.em_sb + (long long)t * p.em_st : nullptr;
            }
            int p_done = s.oa;
            const int ngroups = s.qb >= s.qa ? (s.qb - s.qa + PBI) / PBI : 0;
            int ql = s.qa - 1;
            for (int gi = 0; gi < ngroups; ++gi) {
                // ---- transform stage st in place ----
                warp_wait(raw_full(st), ph, lane);
                if (tid == 0 && gcount == 6) stamp(g, 4);
                const uint32_t hi_base = smem0 + (uint32_t)(st * stage_bytes), lo_base = hi_base + (uint32_t)g.half_bytes;
                if (!(g.dbg & 1)) {
                    float4 x[MAXJ], x2[MAXJ];
#pragma unroll
                    for (int j = 0; j < MAXJ; ++j) {
                        const uint32_t off = (uint32_t)(tid + NWORK * j) * 16u;
                        x[j] = f4zero(); x2[j] = f4zero();
                        if (j < NJ && tid + NWORK * j < 32 * R) {
                            x[j] = lds4(hi_base + off);
                            if (PRO == PRO_BNBWD) x2[j] = lds4(lo_base + off);
                        }
                    }
#pragma unroll
                    for (int j = 0; j < MAXJ; ++j) {
                        if (j < NJ && tid + NWORK * j < 32 * R) {
                            const uint32_t off = (uint32_t)(tid + NWORK * j) * 16u;
                            float4 y;
                            y.x = pro1<PRO>(x[j].x, x2[j].x, mk[j], co4); y.y = pro1<PRO>(x[j].y, x2[j].y, mk[j], co4);
                            y.z = pro1<PRO>(x[j].z, x2[j].z, mk[j], co4); y.w = pro1<PRO>(x[j].w, x2[j].w, mk[j], co4);
                            float4 h, l;
                            tf32_split(y.x, h.x, l.x); tf32_split(y.y, h.y, l.y); tf32_split(y.z, h.z, l.z); tf32_split(y.w, h.w, l.w);
                            sts4(hi_base + off, h);
                            sts4(lo_base + off, l);
                        }
                    }
                }
                if (!(g.dbg & 256)) fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(op_full(st));
                if (tid == 0 && gcount == 6) stamp(g, 5);
                // ---- epilogue of the previous group's completed positions ----
                if (pend.cnt > 0) { run_epilogue(pend, true); pend.cnt = 0; if (tid == 0 && gcount == 7) stamp(g, 10); }
                // positions completed by this group: q_last(pp) = floor((pp*pmul + dpmax) / pdiv) <= ql  (all of them after the last slab)
                ql = min(s.qb, ql + PBI);
                int p_to = s.ob;
                if (ql < s.qb) {
                    const int lim = ql * pdiv + pdiv - 1 - g.dpmax;
                    const int pl = lim < 0 ? -1 : (pmul == 2 ? lim >> 1 : lim);
                    p_to = min(s.ob, pl + 1);
                    if (p_to < p_done) p_to = p_done;
                }
                pend.cnt = p_to - p_done; pend.p_from = p_done; pend.r_from = rbase + (p_done - s.oa);
                pend.bar = grp_done(gd); pend.par = gdph; pend.sg = sg;
                p_done = p_to;
                ++gcount;
                if (++st == NS) { st = 0; ph ^= 1u; }
                if (++gd == NGD) { gd = 0; gdph ^= 1u; }
            }
            if (ngroups == 0) {
                // a run of output positions that no input slab feeds (e.g. a single odd position of a stride-2 shortcut's
                // backward-data): nothing to wait for, but the positions are still written and their slots still cycle
                if (pend.cnt > 0) { run_epilogue(pend, true); pend.cnt = 0; }
                Pend e; e.cnt = s.ob - s.oa; e.p_from = s.oa; e.r_from = rbase; e.bar = 0; e.par = 0; e.sg = sg;
                run_epilogue(e, false);
            }
            rbase += s.ob - s.oa;
            u += s.ob - s.oa;
        }
        if (pend.cnt > 0) run_epilogue(pend, true);
        if (tid == 0) stamp(g, 11);

        // ---- per-channel sums: lanes -> warp (shuffles), the four lane-quarter warps of a channel group -> CTA (shared memory),
        //      then ONE fp64 atomic pair per channel and CTA (148 x 4 contending atomics per address made the tail 8 us long) ----
        if (epi != EPI_STORE && p.stat0 != nullptr) {
            float* sred = reinterpret_cast<float*>(smem);            // [64 channels][4 lane quarters][2]: the operand stages are idle by now
            asm volatile("bar.sync 1, %0;" ::"r"(NWORK) : "memory");
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                const float a = warp_sum(et.s0[c]), bsum = warp_sum(et.s1[c]);
                const int co = et.ec0 + c;
                if (lane == 0 && co < 64) { sred[(co * 4 + (warp & 3)) * 2] = a; sred[(co * 4 + (warp & 3)) * 2 + 1] = bsum; }
            }
            asm volatile("bar.sync 1, %0;" ::"r"(NWORK) : "memory");
            if (tid < 2 * p.Cout) {
                const int co = tid >> 1, k = tid & 1;
                const double v = (double)sred[(co * 4 + 0) * 2 + k] + (double)sred[(co * 4 + 1) * 2 + k] + (double)sred[(co * 4 + 2) * 2 + k] + (double)sred[(co * 4 + 3) * 2 + k];
                atomicAdd((k == 0 ? p.stat0 : p.stat1) + co, v);
            }
        }
        if (tid == 0) stamp(g, 12);
    } else if (warp == NWW) {
        // =============================== TMA producer ===============================
        if (lane == 0) {
            tma_prefetch_desc(&tmA);
            if (PRO == PRO_BNBWD) tma_prefetch_desc(&tmB);
            // all tap weights (hi + lo images) once
            mbar_arrive_expect_tx(w_full, g.w_bytes);
            bulk_g2s(smem_u32(wsm), p.wtc, g.w_bytes, w_full);
            int gg = 0;
            const uint32_t tx = (uint32_t)g.half_bytes * (PRO == PRO_BNBWD ? 2u : 1u);
            for (long long u = u0; u < u1;) {
                Seg s; seg_make(p, g, u, u1, s);
                const int n0 = s.ct * TILE;
                const int ngroups = s.qb >= s.qa ? (s.qb - s.qa + PBI) / PBI : 0;
                for (int gi = 0; gi < ngroups; ++gi, ++gg) {
                    const int st = gg % NS;
                    const uint32_t ph = (uint32_t)((gg / NS) & 1);
                    mbar_wait(slab_empty(st), ph ^ 1u);
                    if (g.dbg & 16) { mbar_arrive(raw_full(st)); continue; }
                    mbar_arrive_expect_tx(raw_full(st), tx);
                    const uint32_t hi_base = smem0 + (uint32_t)(st * stage_bytes);
                    const int q0 = s.qa + gi * PBI;
#pragma unroll
                    for (int blk = 0; blk < 4; ++blk) {
                        tma_load_3d(hi_base + (uint32_t)(blk * R * 128), &tmA, n0 + blk * 32, 0, q0, raw_full(st));
                        if (gg == 6) stamp(g, 3);
                        if (PRO == PRO_BNBWD) tma_load_3d(hi_base + (uint32_t)(g.half_bytes + blk * R * 128), &tmB, n0 + blk * 32, 0, q0, raw_full(st));
                    }
                }
                u += s.ob - s.oa;
            }
        }
    } else {
        // =============================== MMA issue ===============================
        // Per K step (8 channels) of a slab there are TWO MMAs instead of the textbook three of 3xTF32:
        //   D[main | cor] += A_hi * [B_hi ; B_lo]      (N = 2*NPAD: the hi*hi product lands in `main`, hi*lo in `cor`)
        //   D[cor]        += A_lo * B_hi
        // and where N allows (3 taps * 2*NPAD <= 256, stride-1 geometry) the three taps of a slab go into the SAME pair: the
        // accumulator slots descend with the output position, so the positions q+1, q, q-1 a slab feeds are adjacent TMEM column
        // blocks and B = the three taps' [hi ; lo] blocks stacked along N (second MMA: the zero-padded image [0 ; hi] per tap).
        // Accumulators are zero-initialised by the epilogue, hence there are no accumulate flags and NO ORDER between contributions:
        // the (slab, tap, K step) items of a group are dealt round-robin to NI issuing warps.  (One thread retires a dependent
        // instruction every ~5 cycles and an MMA costs ~30 of them with its descriptors: a single issuer was the bottleneck, with
        // the workers 60 % of their time waiting for it.)  Each warp runs the loop with warp-uniform values and lane 0 issues.
        const int me = __shfl_sync(0xffffffffu, warp - (NWW + 1), 0);
        const uint32_t lead = lane == 0 ? 1u : 0u;
        const int KS = p.Cin >> 3, ntaps = p.ntaps, NPAD = g.NPAD;
        const uint32_t idesc1 = umma_idesc_tf32(TILE, 2 * NPAD, 1, 0);           // A: MN-major (columns contiguous), B: K-major
        const uint32_t idesc2 = umma_idesc_tf32(TILE, NPAD, 1, 0);
        const uint32_t idescS = umma_idesc_tf32(TILE, ntaps * 2 * NPAD, 1, 0);
        const uint32_t b_sbo = (uint32_t)(p.Cin / 4) * 128u;                     // bytes between 8-row groups of a weight image
        const uint32_t w1_16 = (smem_u32(wsm) & 0x3FFFFu) >> 4, w2_16 = w1_16 + ((uint32_t)g.w_img_bytes >> 4);
        const uint32_t tap16 = ((uint32_t)(2 * NPAD / 8) * b_sbo) >> 4;          // one tap's [hi ; lo] block, 16-byte units
        // descriptor halves that never change: A = MN-major SWIZZLE_128B_BASE32B (layout 1), B = K-major no swizzle (layout 0)
        const uint32_t a_hi32 = (((uint32_t)g.sbo_mn >> 4) & 0x3FFFu) | (1u << 14) | (1u << 29);
        const uint32_t a_lbo = (((uint32_t)g.lbo_mn >> 4) & 0x3FFFu) << 16;
        const uint32_t b_hi32 = ((b_sbo >> 4) & 0x3FFFu) | (1u << 14);
        const uint32_t b_lbo = (128u >> 4) << 16;
        auto mk_desc = [](uint32_t hi32, uint32_t lo32) { return ((uint64_t)hi32 << 32) | (uint64_t)lo32; };
        // weight images hold the taps in order of ascending dp, i.e. descending output position for a given slab
        int dps[3] = {p.dp[0], ntaps > 1 ? p.dp[1] : 0, ntaps > 2 ? p.dp[2] : 0};
        if (ntaps == 3 && dps[0] > dps[2]) { const int t = dps[0]; dps[0] = dps[2]; dps[2] = t; }
        mbar_wait(w_full, 0);
        if (me == 0 && lane == 0) stamp(g, 6);
        int gg = 0, rbase = 0;
        // Position r may be written once the epilogue has drained position r - NACC (and re-zeroed the slot).  The epilogue drains in
        // order, so ONE mbarrier test on the slot of position r - NACC + W - 1 covers W positions (a test costs ~100 cycles even when
        // it succeeds); the host checks that this look-ahead never waits on the group being issued (ring_ok).
        int drained = g.NACC;                                                    // positions < drained may be written
        const int W = g.W;
        auto acquire = [&](int r) {
            if (r >= drained) {
                const int rw_ = (r - g.NACC) | (W - 1);                                  // the signalling position at or after r - NACC
                warp_wait(acc_empty0 + 8u * (NACCm - (rw_ & NACCm)), (uint32_t)((rw_ / g.NACC) & 1), lane);
                tc_fence_after();
                drained = rw_ + g.NACC + 1;
            }
        };
        // work split between the NI issuing warps: whole slabs when a group has at least NI of them (thin layers: the per-slab
        // scalar work is divided too), otherwise the K steps of every slab
        const bool split_slabs = PBI >= NI;
        int slab_no = 0;
        int item = 0;                                                            // running item counter: item % NI == me -> mine
        for (long long u = u0; u < u1;) {
            Seg s; seg_make(p, g, u, u1, s);
            const int ngroups = s.qb >= s.qa ? (s.qb - s.qa + PBI) / PBI : 0;
            for (int gi = 0; gi < ngroups; ++gi, ++gg) {
                const int st = gg % NS;
                const uint32_t ph = (uint32_t)((gg / NS) & 1);
                warp_wait(op_full(st), ph, lane);
                tc_fence_after();
                if (me == 0 && lane == 0 && gg == 6) stamp(g, 7);
                const uint32_t a_hi16 = (((smem0 + (uint32_t)(st * stage_bytes)) & 0x3FFFFu) >> 4), a_lo16 = a_hi16 + ((uint32_t)g.half_bytes >> 4);
                const int q0 = s.qa + gi * PBI;
                const int ql = min(s.qb, q0 + PBI - 1);
                for (int q = q0; q <= ql && !(g.dbg & 2); ++q) {
                    if (split_slabs) { if (slab_no++ % NI != me) continue; item = me; }      // my slab: all of its items are mine
                    const uint32_t row16 = (uint32_t)((q - q0) * p.Cin) * 8u;            // (rows * 128 B) >> 4
                    const int pp0 = out_pos(pmul, pdiv, s.oa, s.ob, q, dps[0]);
                    const int pp1 = ntaps > 1 ? out_pos(pmul, pdiv, s.oa, s.ob, q, dps[1]) : -1;
                    const int pp2 = ntaps > 2 ? out_pos(pmul, pdiv, s.oa, s.ob, q, dps[2]) : -1;
                    const int r0 = rbase + (pp0 - s.oa);
                    if (g.stack && pp2 >= 0 && pp0 == pp2 + 2 && pp1 == pp2 + 1 && (r0 & NACCm) >= 2) {
                        const uint32_t d = tmem_base + (uint32_t)((NACCm - (r0 & NACCm)) * 2 * NPAD);
                        for (int ks = 0; ks < KS; ++ks, ++item) {
                            if (!split_slabs && item % NI != me) continue;
                            acquire(r0);
                            const uint32_t ao = row16 + (uint32_t)ks * 64u, bo = (uint32_t)ks * 16u;            // 1024 B / 256 B per K step
                            umma_tf32_pred(d, mk_desc(a_hi32, (a_hi16 + ao) | a_lbo), mk_desc(b_hi32, (w1_16 + bo) | b_lbo), idescS, lead);
                            umma_tf32_pred(d, mk_desc(a_hi32, (a_lo16 + ao) | a_lbo), mk_desc(b_hi32, (w2_16 + bo) | b_lbo), idescS, lead);
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 3; ++i) {
                            const int pp = i == 0 ? pp0 : i == 1 ? pp1 : pp2;
                            if (pp < 0) continue;
                            const int r = rbase + (pp - s.oa);
                            const uint32_t d = tmem_base + (uint32_t)((NACCm - (r & NACCm)) * 2 * NPAD);
                            for (int ks = 0; ks < KS; ++ks, ++item) {
                                if (!split_slabs && item % NI != me) continue;
                                acquire(r);
                                const uint32_t ao = row16 + (uint32_t)ks * 64u, bo = (uint32_t)i * tap16 + (uint32_t)ks * 16u;
                                umma_tf32_pred(d, mk_desc(a_hi32, (a_hi16 + ao) | a_lbo), mk_desc(b_hi32, (w1_16 + bo) | b_lbo), idesc1, lead);
                                umma_tf32_pred(d + (uint32_t)NPAD, mk_desc(a_hi32, (a_lo16 + ao) | a_lbo), mk_desc(b_hi32, (w1_16 + bo) | b_lbo), idesc2, lead);
                            }
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) {
                    if (g.dbg & 128) { mbar_arrive(slab_empty(st)); mbar_arrive(grp_done(gg % NGD)); }      // measurement only (with dbg 2)
                    else { umma_commit(slab_empty(st)); umma_commit(grp_done(gg % NGD)); }
                    if (me == 0 && gg == 6) stamp(g, 8);
                }
            }
            rbase += s.ob - s.oa;
            u += s.ob - s.oa;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (tid == 0) stamp(g, 13);
    if (warp == NWW) tmem_dealloc(tmem_base, g.tmem_cols);
    wf_bn_tail(p.tail);
}

// =========================================================================================================
// weight packing: reference [Cout][Cin][ntaps] -> per tap K-major core-matrix images [NPAD rows][K], hi images of all taps then lo
//   forward image:  row = cout, k = cin;   backward-data image: row = cin, k = cout
// =========================================================================================================
__global__ void slab_pack_kernel(SlabPackTable tab, const float* params, float* packed)
{
    wf_pdl_enter();
    const SlabPackEntry e = tab.e[blockIdx.y];
    const int f_np = e.cout < 16 ? 16 : (e.cout + 15) / 16 * 16, b_np = e.cin < 16 ? 16 : (e.cin + 15) / 16 * 16;
    // per direction: image 1 rows per tap = [hi (np) ; lo (np)], image 2 rows per tap = [0 (np) ; hi (np)]
    const long long nf = (long long)e.ntaps * 2 * f_np * e.cin, nb = (long long)e.ntaps * 2 * b_np * e.cout;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nf + nb; i += (long long)gridDim.x * blockDim.x) {
        const bool bwd = i >= nf;
        const long long k_ = bwd ? i - nf : i;
        const int np = bwd ? b_np : f_np, K = bwd ? e.cout : e.cin;
        // image order: [row group (rows/8)][k quad (K/4)][row in group (8)][k in quad (4)], rows = (image tap, hi|lo half, channel)
        const int c4 = (int)(k_ & 3), r8 = (int)((k_ >> 2) & 7), kq = (int)((k_ >> 5) % (K / 4)), rg = (int)(k_ / (8 * K));
        const int row = rg * 8 + r8, kk = kq * 4 + c4;
        const int itap = row / (2 * np), within = row % (2 * np), half = within / np, ch = within % np;
        // taps are stored in order of ascending dp: forward dp = (-1, 0, 1) keeps the reference order, backward-data dp = -dp reverses it
        const int tap = bwd ? e.ntaps - 1 - itap : itap;
        float v = 0.f;
        if (!bwd) { if (ch < e.cout) v = params[e.param_off + ((long long)ch * e.cin + kk) * e.ntaps + tap]; }
        else { if (ch < e.cin) v = params[e.param_off + ((long long)kk * e.cin + ch) * e.ntaps + tap]; }
        float hi, lo;
        tf32_split(v, hi, lo);
        float* dst = packed + (bwd ? e.bwd_off : e.fwd_off);
        const long long img = bwd ? nb : nf;
        dst[k_] = half == 0 ? hi : lo;
        dst[img + k_] = half == 0 ? 0.f : hi;
    }
}

// ---- host: tensor maps ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) f = nullptr;
        return reinterpret_cast<EncodeTiledFn>(f);
    }();
    return fn;
}

// [C][P][N] fp32 tensor (n contiguous) as TMA dims {n, c, pos}; box {32, boxC, boxP}; 128-byte swizzle; out-of-range -> zeros
bool make_map(CUtensorMap* tm, const float* base, int C, int P, long long N, long long sc, long long sp, int boxC, int boxP)
{
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)N, (cuuint64_t)C, (cuuint64_t)P};
    const cuuint64_t strides[2] = {(cuuint64_t)sc * 4, (cuuint64_t)sp * 4};
    const cuuint32_t box[3] = {32, (cuuint32_t)boxC, (cuuint32_t)boxP};
    const cuuint32_t es[3] = {1, 1, 1};
    return fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int dev_sms() { return wf_device_sms(); }

constexpr int SMEM_LIMIT = 227 * 1024;

// host-side simulation of the accumulator ring: every output position first touched in group gi must find its slot's previous
// occupant completed by an EARLIER group (its epilogue runs before the workers transform group gi + 1), otherwise the MMA
// thread would wait on an epilogue that waits on the MMA thread
bool ring_ok(const ConvP& p, int dpmin, int dpmax, int PBI, int NACC, int W)
{
    // worst case: a run over all positions of one tile
    const int oa = 0, ob = p.Pout;
    int lo = oa * p.pmul + dpmin, hi = (ob - 1) * p.pmul + dpmax;
    if (p.pdiv == 2) { lo = (lo + 1) >> 1; hi = hi >> 1; }
    const int qa = lo < 0 ? 0 : lo, qb = hi > p.Pin - 1 ? p.Pin - 1 : hi;
    auto qlast = [&](int pp) { int q = pp * p.pmul + dpmax; if (p.pdiv == 2) q >>= 1; return q > qb ? qb : q; };
    auto grp_of = [&](int q) { return (q - qa) / PBI; };
    for (int pp = NACC; pp < ob; ++pp) {
        // first contributing slab of pp
        int qf = -1;
        for (int t = 0; t < p.ntaps; ++t) {
            const int num = pp * p.pmul + p.dp[t];
            if (num < 0 || (p.pdiv == 2 && (num & 1))) continue;
            const int q = p.pdiv == 2 ? num >> 1 : num;
            if (q < qa || q > qb) continue;
            if (qf < 0 || q < qf) qf = q;
        }
        if (qf < 0) continue;
        // the MMA thread may wait for position pp - NACC + W - 1 (look-ahead) before touching pp
        const int pw = pp - NACC + W - 1 < ob ? pp - NACC + W - 1 : ob - 1;
        if (grp_of(qlast(pw)) >= grp_of(qf)) return false;
    }
    return true;
}

bool plan(const ConvP& p, SlabGeom& g)
{
    g = SlabGeom{};
    int dpmin = p.dp[0], dpmax = p.dp[0];
    for (int t = 1; t < p.ntaps; ++t) { dpmin = p.dp[t] < dpmin ? p.dp[t] : dpmin; dpmax = p.dp[t] > dpmax ? p.dp[t] : dpmax; }
    g.dpmin = dpmin; g.dpmax = dpmax;
    g.NPAD = p.Cout <= 16 ? 16 : (p.Cout + 15) / 16 * 16;
    g.NACC = 512 / (2 * g.NPAD);
    if (g.NACC > MAXACC) g.NACC = MAXACC;
    g.tmem_cols = 32;
    while (g.tmem_cols < g.NACC * 2 * g.NPAD) g.tmem_cols *= 2;
    g.w_img_bytes = p.ntaps * 2 * g.NPAD * p.Cin * 4;
    g.stack = (p.ntaps == 3 && p.pmul == 1 && p.ntaps * 2 * g.NPAD <= 256) ? 1 : 0;
    g.w_bytes = g.w_img_bytes * (g.stack ? 2 : 1);
    const int fixed = g.w_bytes + 4 * 64 * 4 + (3 * MAXNS + NGD + MAXACC + 1) * 8 + 16 + 1024;
    int PBI = 64 / p.Cin;
    if (PBI > p.Pin) { PBI = 1; while (PBI * 2 <= p.Pin && PBI * 2 * p.Cin <= 64) PBI *= 2; }
    g.W = g.NACC / 4;
    while (PBI > 1 && !ring_ok(p, dpmin, dpmax, PBI, g.NACC, g.W)) PBI >>= 1;
    while (g.W > 1 && !ring_ok(p, dpmin, dpmax, PBI, g.NACC, g.W)) g.W >>= 1;
    if (!ring_ok(p, dpmin, dpmax, PBI, g.NACC, g.W)) return false;
    g.PBI = PBI;
    g.R = PBI * p.Cin;
    g.half_bytes = 4 * g.R * 128;
    g.NS = (SMEM_LIMIT - fixed) / (2 * g.half_bytes);
    if (g.NS > 3) g.NS = 3;
    if (g.NS < 2) return false;
    g.ntiles = (p.N + TILE - 1) / TILE;
    g.units = (long long)g.ntiles * p.Pout;
    g.lbo_mn = g.R * 128;
    g.sbo_mn = 512;
    return true;
}

size_t smem_bytes(const SlabGeom& g)
{
    return (size_t)g.NS * 2 * g.half_bytes + (size_t)g.w_bytes + 4 * 64 * 4 + (3 * MAXNS + NGD + MAXACC + 1) * 8 + 16;
}

template <int PRO, bool MASK, int CH>
cudaError_t launch_t(const CUtensorMap& a, const CUtensorMap& b, const ConvP& p, const SlabGeom& g, int grid, cudaStream_t st)
{
    static WfSmemOptIn optin;
    if (cudaError_t e = wf_smem_optin(optin, slab_tc_kernel<PRO, MASK, CH>, SMEM_LIMIT)) return e;
    wf_launch_pdl(slab_tc_kernel<PRO, MASK, CH>, dim3(grid), dim3(NTHR), smem_bytes(g), st, a, b, p, g);
    return cudaGetLastError();
}

template <int PRO, bool MASK>
cudaError_t launch_ch(const CUtensorMap& a, const CUtensorMap& b, const ConvP& p, const SlabGeom& g, int grid, cudaStream_t st)
{
    const int ch = p.Cout <= 8 ? 2 : p.Cout <= 16 ? 4 : p.Cout <= 32 ? 8 : 16;        // channels per worker warp (4 warps per TMEM lane quarter)
    switch (ch) {
        case 2: return launch_t<PRO, MASK, 2>(a, b, p, g, grid, st);
        case 4: return launch_t<PRO, MASK, 4>(a, b, p, g, grid, st);
        case 8: return launch_t<PRO, MASK, 8>(a, b, p, g, grid, st);
        default: return launch_t<PRO, MASK, 16>(a, b, p, g, grid, st);
    }
}

int min_columns() { static const int v = [] { const char* e = std::getenv("WF_SLABTC_MIN_N"); return e ? std::atoi(e) : 4096; }(); return v; }
bool wgrad_enabled() { static const bool v = [] { const char* e = std::getenv("WF_DISABLE_SLABTC_WGRAD"); return !(e && e[0] == '1'); }(); return v; }
bool thin_enabled() { static const bool v = [] { const char* e = std::getenv("WF_SLABTC_THIN"); return e && e[0] == '1'; }(); return v; }
// =========================================================================================================
// backward-weights:  dW[o][i][t] = sum over output positions p and columns n of  G'[o][p][n] * X'[i][p*s + dp[t]][n]
//
// Both operands are K-major here (K = the column n, contiguous in the [channel][position][n] layout), so plain SWIZZLE_128B TMA
// boxes {32 n, channels, positions} are the canonical UMMA layout.  One MMA covers a BLOCK of positions at once:
//   A rows = (input position q0 .. q0+PA-1, input channel)        M = PA*Cin  (one or two 128-row tiles)
//   B rows = (output position p0 .. p0+PBk-1, output channel)      N = PBk*Cout <= 128
// so D[(q_rel, i)][(p_rel, o)] holds ALL pairs of the block; the pairs with q_rel = p_rel*s + dp[t] - dpmin are the taps.  The
// relative layout is the same for every position block and every column range, hence a CTA accumulates its whole share of the
// (block, 32-column stage) list into ONE set of TMEM accumulators and extracts the tap diagonals once, at the end (shared-memory
// reduction over the block's positions, one fp32 reduction per weight and CTA into the gradient).  3xTF32 as in the forward
// kernel: D[main | cor] += A_hi * [B_hi ; B_lo],  D[cor] += A_lo * B_hi, accumulators zero-initialised, contributions dealt to
// the NI issuing warps.  Rows of positions outside the tensor and columns beyond N are zeroed by the transform (the activation
// of a zero-filled element is not zero).
// =========================================================================================================
struct WgGeom {
    int MT;                 // 128-row tiles of A
    int PA, PBk;            // input / output positions per block
    int NBr;                // B rows = PBk*Cout rounded up to 8
    int NS, a_half, b_half; // stages; bytes of one of hi/lo of the A / B tile of a stage
    int nblocks, ncol32;    // position blocks, 32-column stages per block
    int s, dpmin;           // effective position stride and smallest tap offset
    int Pin_eff;            // input positions of the (possibly strided) view
    int tmem_cols;
    long long stages;       // nblocks * ncol32
    int dbg;
};

template <int XPRO, bool MASK>
__global__ void __launch_bounds__(NTHR, 1) slab_wgrad_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmG,
                                                             const __grid_constant__ CUtensorMap tmR, const WgradP p, const WgGeom g)
{
    wf_pdl_enter();
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int stage_bytes = 2 * (g.a_half + g.b_half);
    float* tab = reinterpret_cast<float*>(smem + g.NS * stage_bytes);          // X coefficients a, b, d [64] each; G' coefficients a, b, c, d [64] each
    uint64_t* bars = reinterpret_cast<uint64_t*>(tab + 7 * 64);
    const uint32_t bar0 = smem_u32(bars);
    auto raw_full = [&](int s) { return bar0 + 8u * s; };
    auto op_full = [&](int s) { return bar0 + 8u * (MAXNS + s); };
    auto slab_empty = [&](int s) { return bar0 + 8u * (2 * MAXNS + s); };
    const uint32_t all_done = bar0 + 8u * (3 * MAXNS);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * MAXNS + 1);

    if (tid == 0) {
        for (int s = 0; s < g.NS; ++s) { mbar_init(raw_full(s), 1); mbar_init(op_full(s), NWW); mbar_init(slab_empty(s), NI); }
        mbar_init(all_done, NI);
        fence_mbar_init();
    }
    if (warp == NWW) { tmem_alloc(smem_u32(tmem_slot), g.tmem_cols); tmem_relinquish(); }
    for (int i = tid; i < 7 * 64; i += NTHR) {
        const int which = i >> 6, c = i & 63;
        float v = 0.f;
        if (which < 3) { if (XPRO != PRO_NONE && c < p.Cin) v = which == 0 ? p.pro_a[c] : which == 1 ? p.pro_b[c] : p.pro_d[c]; }
        else if (c < p.Cout) v = which == 3 ? p.g_a[c] : which == 4 ? p.g_b[c] : which == 5 ? p.g_c[c] : p.g_d[c];
        tab[i] = v;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (warp < NWW) {                      // zero the accumulators: every MMA accumulates
        const uint32_t t0 = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        const int ncols = g.MT * 2 * g.NBr;
        for (int c = (warp >> 2) * 8; c < ncols; c += 32) tmem_zero<8>(t0 + c);
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    const long long u0 = g.stages * blockIdx.x / gridDim.x, u1 = g.stages * (blockIdx.x + 1) / gridDim.x;
    const int NS = g.NS;
    const uint32_t smem0 = smem_u32(smem);
    const int arows = g.PA * p.Cin, brows = g.PBk * p.Cout;

    if (warp < NWW) {
        // =============================== workers: transform, then the final extraction ===============================
        // chunk id = tid + 512 j over [A rows | B rows] x 8 sixteen-byte chunks; K-major SWIZZLE_128B: chunk c of row r holds the
        // columns 4*(c ^ (r & 7)) .. +3 of the stage.  (Decoding a thread's chunks once into registers was tried: at the 96-register
        // budget of a 20-warp CTA it spills and runs 10-100 % slower than this loop.)
        const int a_chunks = arows * 8, tot_chunks = a_chunks + brows * 8;
        const int cin_sh = 31 - __clz(p.Cin), cout_sh = 31 - __clz(p.Cout);
        int st = 0; uint32_t ph = 0;
        int blk = (int)(u0 / g.ncol32), c32 = (int)(u0 - (long long)blk * g.ncol32);
        for (long long u = u0; u < u1; ++u) {
            const int n0 = c32 * 32, p0 = blk * g.PBk, q0 = p0 * g.s + g.dpmin;
            // Dropout2d values of this thread's X' chunks (one per (window, channel); the A rows are at most 256, i.e. at most four
            // chunks per thread): requested BEFORE the wait for the TMA tile, all at once -- inside the loop each was a dependent
            // global load in front of the chunk's transform (3.4 M two-sector requests per step in the ncu request survey).
            float mkv[4] = {1.f, 1.f, 1.f, 1.f};
            if (MASK) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int id = tid + j * NWORK;
                    if (id < a_chunks) {
                        const int row = id >> 3, quad = (id & 7) ^ (row & 7);
                        const int n = n0 + quad * 4, c = row & (p.Cin - 1);
                        if (n < p.N) mkv[j] = __ldg(p.mask + (long long)(n / WF_T) * p.m_sb + (long long)c * p.m_sc);
                    }
                }
            }
            warp_wait(raw_full(st), ph, lane);
            const uint32_t a_hi = smem0 + (uint32_t)(st * stage_bytes), a_lo = a_hi + (uint32_t)g.a_half;
            const uint32_t b_hi = a_lo + (uint32_t)g.a_half, b_lo = b_hi + (uint32_t)g.b_half;
            // (two chunks per trip with their loads batched was tried: 5-10 % slower at the 96-register budget)
            int jj = 0;
            for (int id = tid; id < ((g.dbg & 1) ? 0 : tot_chunks); id += NWORK, ++jj) {      // (dbg 1: measurement without the transform)
                const bool isA = id < a_chunks;
                const int cid = isA ? id : id - a_chunks;
                const int row = cid >> 3, quad = (cid & 7) ^ (row & 7);
                const int n = n0 + quad * 4;
                const uint32_t off = (uint32_t)cid * 16u;
                float4 y;
                if (isA) {
                    const int pos = row >> cin_sh, c = row & (p.Cin - 1);
                    const float4 x = lds4(a_hi + off);
                    const bool ok = n < p.N && q0 + pos >= 0 && q0 + pos < g.Pin_eff;
                    const float mk = jj == 0 ? mkv[0] : jj == 1 ? mkv[1] : jj == 2 ? mkv[2] : mkv[3];
                    const float4 co = make_float4(tab[c], tab[64 + c], 0.f, tab[128 + c]);
                    y.x = ok ? pro1<XPRO>(x.x, 0.f, mk, co) : 0.f; y.y = ok ? pro1<XPRO>(x.y, 0.f, mk, co) : 0.f;
                    y.z = ok ? pro1<XPRO>(x.z, 0.f, mk, co) : 0.f; y.w = ok ? pro1<XPRO>(x.w, 0.f, mk, co) : 0.f;
                } else {
                    const int pos = row >> cout_sh, o = row & (p.Cout - 1);
                    const float4 x = lds4(b_hi + off), x2 = lds4(b_lo + off);
                    const bool ok = n < p.N && p0 + pos < p.Pout;
                    const float4 co = make_float4(tab[192 + o], tab[256 + o], tab[320 + o], tab[384 + o]);
                    y.x = ok ? pro1<PRO_BNBWD>(x.x, x2.x, 1.f, co) : 0.f; y.y = ok ? pro1<PRO_BNBWD>(x.y, x2.y, 1.f, co) : 0.f;
                    y.z = ok ? pro1<PRO_BNBWD>(x.z, x2.z, 1.f, co) : 0.f; y.w = ok ? pro1<PRO_BNBWD>(x.w, x2.w, 1.f, co) : 0.f;
                }
                float4 h, l;
                tf32_split(y.x, h.x, l.x); tf32_split(y.y, h.y, l.y); tf32_split(y.z, h.z, l.z); tf32_split(y.w, h.w, l.w);
                sts4((isA ? a_hi : b_hi) + off, h);
                sts4((isA ? a_lo : b_lo) + off, l);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(op_full(st));
            if (++st == NS) { st = 0; ph ^= 1u; }
            if (++c32 == g.ncol32) { c32 = 0; ++blk; }
        }
        // ---- extraction: thread = accumulator row (q_rel, i); the taps are the diagonals q_rel = p_rel*s + dp[t] - dpmin ----
        warp_wait(all_done, 0, lane);
        tc_fence_after();
        float* red = reinterpret_cast<float*>(smem);                      // [Cout][Cin][ntaps], reuses the stage memory
        const int nw = p.Cout * p.Cin * p.ntaps;
        for (int i = tid; i < nw; i += NWORK) red[i] = 0.f;
        asm volatile("bar.sync 1, %0;" ::"r"(NWORK) : "memory");
        if (u1 > u0 && !(g.dbg & 4)) {
            const int lq = warp & 3, part = warp >> 2;
            for (int mt = 0; mt < g.MT; ++mt) {
                const int row = mt * 128 + lq * 32 + lane;
                const int q_rel = row / p.Cin, ci = row - q_rel * p.Cin;
                const bool rv = row < arows;
                const uint32_t t_row = tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(mt * 2 * g.NBr);
                for (int pr = part; pr < g.PBk; pr += 4) {                // the four warps of a lane quarter share the output positions
                    const int dq = q_rel - pr * g.s + g.dpmin;           // tap offset this (row, position) pair would be
                    int t = -1;
                    for (int k = 0; k < p.ntaps; ++k) if (p.dp[k] == dq) t = k;
                    for (int o0 = 0; o0 < p.Cout; o0 += 8) {
                        float v[8], cr[8];
                        tmem_ldn<8>(t_row + (uint32_t)(pr * p.Cout + o0), v);
                        tmem_ldn<8>(t_row + (uint32_t)(g.NBr + pr * p.Cout + o0), cr);
                        tmem_ld_wait();
                        if (rv && t >= 0) {
#pragma unroll
                            for (int k = 0; k < 8; ++k) atomicAdd(red + ((o0 + k) * p.Cin + ci) * p.ntaps + t, v[k] + cr[k]);
                        }
                    }
                }
            }
        }
        asm volatile("bar.sync 1, %0;" ::"r"(NWORK) : "memory");
        if (u1 > u0 && !(g.dbg & 8)) {
            // one reduction per weight and CTA into the gradient; 128-bit reductions (sm_90+) where the tensor's offset in the flat
            // gradient allows it: 148 CTAs x 12 288 weights of a 64 -> 64 layer are 1.8 M scalar reductions per launch otherwise
            if (((reinterpret_cast<uintptr_t>(p.dw) & 15) == 0) && (nw & 3) == 0) {
                for (int i = tid * 4; i < nw; i += NWORK * 4)
                    atomicAdd(reinterpret_cast<float4*>(p.dw + i), *reinterpret_cast<const float4*>(red + i));
            } else {
                for (int i = tid; i < nw; i += NWORK) atomicAdd(p.dw + i, red[i]);
            }
        }
    } else if (warp == NWW) {
        // =============================== TMA producer ===============================
        if (lane == 0) {
            tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmG); tma_prefetch_desc(&tmR);
            const uint32_t tx = (uint32_t)(arows + 2 * brows) * 128u;
            int st = 0; uint32_t ph = 0;
            int blk = (int)(u0 / g.ncol32), c32 = (int)(u0 - (long long)blk * g.ncol32);
            for (long long u = u0; u < u1; ++u) {
                const int n0 = c32 * 32, p0 = blk * g.PBk, q0 = p0 * g.s + g.dpmin;
                mbar_wait(slab_empty(st), ph ^ 1u);
                const uint32_t a_hi = smem0 + (uint32_t)(st * stage_bytes);
                const uint32_t b_hi = a_hi + 2u * (uint32_t)g.a_half, b_lo = b_hi + (uint32_t)g.b_half;
                if (g.dbg & 16) mbar_arrive(raw_full(st));                 // (measurement without the TMA loads)
                else {
                    mbar_arrive_expect_tx(raw_full(st), tx);
                    tma_load_3d(a_hi, &tmX, n0, 0, q0, raw_full(st));
                    tma_load_3d(b_hi, &tmG, n0, 0, p0, raw_full(st));
                    tma_load_3d(b_lo, &tmR, n0, 0, p0, raw_full(st));
                }
                if (++st == NS) { st = 0; ph ^= 1u; }
                if (++c32 == g.ncol32) { c32 = 0; ++blk; }
            }
        }
    } else {
        // =============================== MMA issue ===============================
        const int me = __shfl_sync(0xffffffffu, warp - (NWW + 1), 0);
        const uint32_t lead = lane == 0 ? 1u : 0u;
        const uint32_t idesc1 = umma_idesc_tf32(TILE, 2 * g.NBr, 0, 0), idesc2 = umma_idesc_tf32(TILE, g.NBr, 0, 0);     // both operands K-major
        const uint32_t hi32 = ((1024u >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);           // SBO = 1024 B (8 rows), SWIZZLE_128B
        const uint32_t lbo = 1u << 16;
        auto mk_desc = [](uint32_t h, uint32_t l) { return ((uint64_t)h << 32) | (uint64_t)l; };
        int st = 0; uint32_t ph = 0;
        int item = 0;
        for (long long u = u0; u < u1; ++u) {
            warp_wait(op_full(st), ph, lane);
            tc_fence_after();
            const uint32_t a_hi16 = ((smem0 + (uint32_t)(st * stage_bytes)) & 0x3FFFFu) >> 4, a_lo16 = a_hi16 + ((uint32_t)g.a_half >> 4);
            const uint32_t b_hi16 = a_lo16 + ((uint32_t)g.a_half >> 4);
            for (int ks = 0; ks < 4; ++ks)
                for (int mt = 0; mt < g.MT; ++mt, ++item) {
                    if (item % NI != me || (g.dbg & 2)) continue;            // (dbg 2: measurement without the MMAs)
                    const uint32_t ao = (uint32_t)(mt * 128 * 128 + ks * 32) >> 4, bo = (uint32_t)(ks * 32) >> 4;    // 32 B per K step inside the 128 B row
                    const uint32_t d = tmem_base + (uint32_t)(mt * 2 * g.NBr);
                    umma_tf32_pred(d, mk_desc(hi32, (a_hi16 + ao) | lbo), mk_desc(hi32, (b_hi16 + bo) | lbo), idesc1, lead);
                    umma_tf32_pred(d + (uint32_t)g.NBr, mk_desc(hi32, (a_lo16 + ao) | lbo), mk_desc(hi32, (b_hi16 + bo) | lbo), idesc2, lead);
                }
            __syncwarp();
            if (lane == 0) umma_commit(slab_empty(st));
            if (++st == NS) { st = 0; ph ^= 1u; }
        }
        if (lane == 0) umma_commit(all_done);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == NWW) tmem_dealloc(tmem_base, g.tmem_cols);
}

// TMA map with SWIZZLE_128B over a [C][P][N] tensor viewed with a position stride (1-tap strided shortcuts read every s-th position)
bool make_map_k(CUtensorMap* tm, const float* base, int C, int P, long long N, long long sc, long long sp, int boxC, int boxP)
{
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)N, (cuuint64_t)C, (cuuint64_t)P};
    const cuuint64_t strides[2] = {(cuuint64_t)sc * 4, (cuuint64_t)sp * 4};
    const cuuint32_t box[3] = {32, (cuuint32_t)boxC, (cuuint32_t)boxP};
    const cuuint32_t es[3] = {1, 1, 1};
    return fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool plan_wgrad(const WgradP& p, WgGeom& g)
{
    g = WgGeom{};
    int dpmin = p.dp[0], dpmax = p.dp[0];
    for (int t = 1; t < p.ntaps; ++t) { dpmin = p.dp[t] < dpmin ? p.dp[t] : dpmin; dpmax = p.dp[t] > dpmax ? p.dp[t] : dpmax; }
    // a single tap reads every pmul-th input position: the tensor map strides over them and the kernel sees stride 1, offset 0
    g.s = p.ntaps == 1 ? 1 : p.pmul;
    g.dpmin = p.ntaps == 1 ? 0 : dpmin;
    g.Pin_eff = p.ntaps == 1 ? p.Pout : p.Pin;
    const int span = p.ntaps == 1 ? 0 : dpmax - dpmin;
    int best = 0;
    for (int mt = 1; mt <= 2; ++mt) {
        int pbk = 128 / p.Cout;
        while (pbk > 0 && ((pbk - 1) * g.s + span + 1) * p.Cin > 128 * mt) --pbk;
        if (pbk > p.Pout) pbk = p.Pout;
        if (pbk <= 0) continue;
        // cost per output position ~ (MMA work + staging of mt A tiles) / pbk
        if (best == 0 || pbk * g.MT >= g.PBk * mt) { g.MT = mt; g.PBk = pbk; best = 1; }      // twice the rows only for at least twice the positions
    }
    if (!best) return false;
    g.PA = (g.PBk - 1) * g.s + span + 1;
    if (g.PA > 256 || g.PBk > 256) return false;
    g.NBr = (g.PBk * p.Cout + 15) / 16 * 16;          // UMMA N granularity at M = 128
    g.a_half = g.MT * 128 * 128;
    g.b_half = g.NBr * 128;
    const int fixed = 7 * 64 * 4 + (3 * MAXNS + 2) * 8 + 16 + 1024;
    g.NS = (SMEM_LIMIT - fixed) / (2 * (g.a_half + g.b_half));
    if (g.NS > 4) g.NS = 4;
    if (g.NS < 2) return false;
    if ((size_t)p.Cout * p.Cin * p.ntaps * 4 > (size_t)g.NS * 2 * (g.a_half + g.b_half)) return false;
    g.nblocks = (p.Pout + g.PBk - 1) / g.PBk;
    g.ncol32 = (p.N + 31) / 32;
    g.stages = (long long)g.nblocks * g.ncol32;
    g.tmem_cols = 32;
    while (g.tmem_cols < g.MT * 2 * g.NBr) g.tmem_cols *= 2;
    return g.tmem_cols <= 512;
}

template <int XPRO, bool MASK>
cudaError_t launch_wg(const CUtensorMap& tx, const CUtensorMap& tg, const CUtensorMap& tr, const WgradP& p, const WgGeom& g, int grid, cudaStream_t st)
{
    static WfSmemOptIn optin;
    if (cudaError_t e = wf_smem_optin(optin, slab_wgrad_kernel<XPRO, MASK>, SMEM_LIMIT)) return e;
    const size_t smem = (size_t)g.NS * 2 * (g.a_half + g.b_half) + 7 * 64 * 4 + (3 * MAXNS + 2) * 8 + 16;
    wf_launch_pdl(slab_wgrad_kernel<XPRO, MASK>, dim3(grid), dim3(NTHR), smem, st, tx, tg, tr, p, g);
    return cudaGetLastError();
}

const bool g_enabled = [] { const char* e = std::getenv("WF_DISABLE_SLABTC"); return !(e && e[0] == '1'); }();
// debugging aid of the self-test: swap the two stride fields of the MN-major descriptor
const bool g_swap_lbo = [] { const char* e = std::getenv("WF_SLABTC_SWAP_LBO"); return e && e[0] == '1'; }();

}  // namespace

// timeline probe (WF_SLABTC_DBG bit 512): copies and clears the 32 globaltimer stamps of CTA 0
cudaError_t wf_slabtc_debug_cta(unsigned long long* out)
{
    cudaError_t e = cudaMemcpyFromSymbol(out, g_cta, sizeof(g_cta));
    if (e != cudaSuccess) return e;
    void* q = nullptr;
    if ((e = cudaGetSymbolAddress(&q, g_cta)) != cudaSuccess) return e;
    return cudaMemset(q, 0, sizeof(g_cta));
}

cudaError_t wf_slabtc_debug_ts(unsigned long long* out)
{
    cudaError_t e = cudaMemcpyFromSymbol(out, g_ts, sizeof(g_ts));
    if (e != cudaSuccess) return e;
    unsigned long long z[32] = {0};
    return cudaMemcpyToSymbol(g_ts, z, sizeof(z));
}

long long wf_slabtc_pack_floats(int cout, int cin, int ntaps, bool bwd)
{
    const int rows = bwd ? cin : cout, k = bwd ? cout : cin;
    const int np = rows <= 16 ? 16 : (rows + 15) / 16 * 16;
    return 4LL * ntaps * np * k;            // two images of ntaps * 2*np rows
}

cudaError_t wf_launch_slabtc_pack(const SlabPackTable& tab, const float* params, float* packed, cudaStream_t st)
{
    if (tab.n == 0) return cudaSuccess;
    wf_launch_pdl(slab_pack_kernel, dim3(8, tab.n), dim3(256), 0, st, tab, params, packed);
    return cudaGetLastError();
}

bool wf_slabtc_shape_ok(int cin, int cout, int groups, int ntaps, const int* dn)
{
    if (groups != 1 || ntaps > 3) return false;
    for (int t = 0; t < ntaps; ++t) if (dn[t] != 0) return false;
    if (!(cin == 8 || cin == 16 || cin == 32 || cin == 64)) return false;
    if (!(cout == 8 || cout == 16 || cout == 32 || cout == 64)) return false;
    return true;
}

bool wf_slabtc_conv_ok(const ConvP& p)
{
    if (!g_enabled || p.wtc == nullptr) return false;
    // 8 -> 8 channel layers: the per-position cost of the TMEM round trips is not amortised (measured 147 us vs 141 us for the
    // mma.sync kernel on up.block.4); they stay on wf_slide.cu unless WF_SLABTC_THIN=1
    if (p.Cin <= 8 && p.Cout <= 8 && !thin_enabled()) return false;
    // small batches: the kernel's fixed cost (~30 us: 227 KB CTAs, weight images, TMEM set-up, pipeline fill) exceeds the whole
    // layer on the mma.sync path (B = 64: 2.73 ms per step with the slab kernels, 2.50 ms without)
    if (p.N < min_columns()) return false;
    if (!wf_slabtc_shape_ok(p.Cin, p.Cout, p.groups, p.ntaps, p.dn)) return false;
    // taps: {0} or {-1, 0, +1} in ascending (forward image) or descending (backward-data image) order
    if (p.ntaps == 1 ? p.dp[0] != 0 : (p.ntaps != 3 || p.dp[1] != 0 || p.dp[0] * p.dp[2] != -1 || p.dp[0] + p.dp[2] != 0)) return false;
    if (!(p.pmul == 1 || p.pmul == 2) || !(p.pdiv == 1 || p.pdiv == 2) || (p.pmul == 2 && p.pdiv == 2)) return false;
    if (p.in_sb != WF_T || p.out_sb != WF_T || (p.N & 3) || p.N % WF_T) return false;
    if ((reinterpret_cast<uintptr_t>(p.in) & 15) || ((p.in_sc * 4) & 15) || ((p.in_sp * 4) & 15)) return false;
    if (p.pro_mode == PRO_BNBWD && (!p.in2 || (reinterpret_cast<uintptr_t>(p.in2) & 15))) return false;
    if (p.pro_mode == PRO_BNSILU && p.mask && p.m_st != 0) return false;
    if (p.Pin > 4096 || p.Pout > 4096) return false;
    if ((long long)p.Cout * p.out_sc >= (1LL << 31) || (long long)p.Pout * p.out_sp >= (1LL << 31)) return false;      // 32-bit offsets in the epilogue
    if (p.accumulate && p.epi_mode != EPI_STORE) return false;
    SlabGeom g;
    return plan(p, g);
}

cudaError_t wf_launch_slabtc_conv(const ConvP& p, cudaStream_t st)
{
    SlabGeom g;
    if (!plan(p, g)) return cudaErrorInvalidValue;
    if (g_swap_lbo) { const int t = g.lbo_mn; g.lbo_mn = g.sbo_mn; g.sbo_mn = t; }
    { const char* e = std::getenv("WF_SLABTC_DBG"); g.dbg = e ? std::atoi(e) : 0; }
    CUtensorMap ta, tb;
    std::memset(&ta, 0, sizeof(ta)); std::memset(&tb, 0, sizeof(tb));
    if (!make_map(&ta, p.in, p.Cin, p.Pin, p.N, p.in_sc, p.in_sp, p.Cin, g.PBI)) return cudaErrorInvalidValue;
    if (p.pro_mode == PRO_BNBWD) { if (!make_map(&tb, p.in2, p.Cin, p.Pin, p.N, p.in_sc, p.in_sp, p.Cin, g.PBI)) return cudaErrorInvalidValue; }
    else tb = ta;
    int grid = dev_sms();
    if ((long long)grid > g.units) grid = (int)g.units;
    const bool mask = p.pro_mode == PRO_BNSILU && p.mask != nullptr;
    switch (p.pro_mode) {
        case PRO_NONE: return launch_ch<PRO_NONE, false>(ta, tb, p, g, grid, st);
        case PRO_BNSILU: return mask ? launch_ch<PRO_BNSILU, true>(ta, tb, p, g, grid, st) : launch_ch<PRO_BNSILU, false>(ta, tb, p, g, grid, st);
        case PRO_AFFINE: return launch_ch<PRO_AFFINE, false>(ta, tb, p, g, grid, st);
        default: return launch_ch<PRO_BNBWD, false>(ta, tb, p, g, grid, st);
    }
}

bool wf_slabtc_wgrad_ok(const WgradP& p)
{
    if (!g_enabled || !wgrad_enabled()) return false;
    if (p.groups != 1 || !(p.ntaps == 1 || p.ntaps == 3) || p.g_pro != PRO_BNBWD || !p.g2) return false;
    for (int t = 0; t < p.ntaps; ++t) if (p.dn[t] != 0) return false;
    if (p.ntaps == 1 ? p.dp[0] != 0 : (p.dp[1] != 0 || p.dp[0] * p.dp[2] != -1 || p.dp[0] + p.dp[2] != 0)) return false;
    if (!(p.Cin == 8 || p.Cin == 16 || p.Cin == 32 || p.Cin == 64) || !(p.Cout == 8 || p.Cout == 16 || p.Cout == 32 || p.Cout == 64)) return false;
    // (8 -> 8 layers included: 170 us vs 226 us for the mma.sync weight-gradient kernel on up.block.4)
    if (!(p.pmul == 1 || p.pmul == 2) || p.in_sb != WF_T || (p.N & 3) || p.N % WF_T || p.N < min_columns()) return false;
    if (p.pro_mode == PRO_BNBWD || (p.pro_mode == PRO_BNSILU && p.mask && p.m_st != 0)) return false;
    if ((reinterpret_cast<uintptr_t>(p.in) & 15) || (reinterpret_cast<uintptr_t>(p.g) & 15) || (reinterpret_cast<uintptr_t>(p.g2) & 15)) return false;
    if (((p.in_sc * 4) & 15) || ((p.in_sp * 4) & 15)) return false;
    WgGeom g;
    return plan_wgrad(p, g);
}

cudaError_t wf_launch_slabtc_wgrad(const WgradP& p, cudaStream_t st)
{
    WgGeom g;
    if (!plan_wgrad(p, g)) return cudaErrorInvalidValue;
    { const char* e = std::getenv("WF_SLABTC_DBG"); g.dbg = e ? std::atoi(e) : 0; }
    CUtensorMap tx, tg, tr;
    std::memset(&tx, 0, sizeof(tx)); std::memset(&tg, 0, sizeof(tg)); std::memset(&tr, 0, sizeof(tr));
    const bool one = p.ntaps == 1;
    if (!make_map_k(&tx, p.in, p.Cin, one ? p.Pout : p.Pin, p.N, p.in_sc, one ? p.in_sp * p.pmul : p.in_sp, p.Cin, g.PA)) return cudaErrorInvalidValue;
    const long long g_sc = (long long)p.Pout * p.N;
    if (!make_map_k(&tg, p.g, p.Cout, p.Pout, p.N, g_sc, p.N, p.Cout, g.PBk)) return cudaErrorInvalidValue;
    if (!make_map_k(&tr, p.g2, p.Cout, p.Pout, p.N, g_sc, p.N, p.Cout, g.PBk)) return cudaErrorInvalidValue;
    int grid = dev_sms();
    if ((long long)grid > g.stages) grid = (int)g.stages;
    const bool mask = p.pro_mode == PRO_BNSILU && p.mask != nullptr;
    switch (p.pro_mode) {
        case PRO_NONE: return launch_wg<PRO_NONE, false>(tx, tg, tr, p, g, grid, st);
        case PRO_BNSILU: return mask ? launch_wg<PRO_BNSILU, true>(tx, tg, tr, p, g, grid, st) : launch_wg<PRO_BNSILU, false>(tx, tg, tr, p, g, grid, st);
        default: return launch_wg<PRO_AFFINE, false>(tx, tg, tr, p, g, grid, st);
    }
}
