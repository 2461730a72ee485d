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
// Warp roles (320 threads): warp 8 lane 0 issues the TMA loads, warp 9 lane 0 issues tcgen05.mma, warps 0-7 are workers:
//   * transform: the raw slab (TMA) is rewritten IN PLACE as the activated operand -- BatchNorm+SiLU+Dropout2d, BatchNorm only,
//     or BatchNorm-backward of (dy, raw) -- split into tf32 hi / lo images (3xTF32, fp32 parity: a*b ~ a_lo*b_hi + a_hi*b_lo +
//     a_hi*b_hi).  Addresses are linear (the swizzle permutes 32-byte chunks inside a row, a row is one channel), so the
//     128-bit shared-memory accesses are conflict free and no index arithmetic is left in the loop;
//   * epilogue: tcgen05.ld (thread = column n, registers = channels) -> bias / SiLU'*mask / BatchNorm sums -> coalesced stores
//     (a warp writes 128 contiguous bytes per channel).  Per-channel sums stay in registers for the whole walk of the CTA and
//     are reduced across lanes once, at the end.
// Work is the flattened list of (column tile, output position) units cut into one contiguous run per CTA (persistent, 1 CTA/SM).
#include <cuda.h>
#include <cstdlib>
#include <cstring>
#include "wf_tc.cuh"
#include "wf_common.cuh"
#include "wf_elem.h"

namespace {

using namespace tc;

constexpr int TILE = 128;                 // columns per tile = UMMA M
constexpr int NWORK = 256;                // worker threads (warps 0-7)
constexpr int NWW = NWORK / 32;
constexpr int NTHR = NWORK + 64;          // + TMA warp + MMA warp
constexpr int MAXACC = 16;                // accumulator slots (barriers) at most
constexpr int NGD = 4;                    // ring of "group done" barriers
constexpr int MAXNS = 4;
constexpr int MAXJ = 8;                   // float4 chunks per worker thread and stage (R = 64 rows)

struct SlabGeom {
    int R, PBI, NS, NPAD, NACC, tmem_cols, ntiles;
    int half_bytes;                       // one of hi/lo of a stage: 4 * R * 128
    int w_half_bytes;                     // one of hi/lo of all tap weights: ntaps * NPAD * Cin * 4
    int dpmin, dpmax;
    int lbo_mn, sbo_mn;                   // MN-major SW128_BASE32B descriptor strides: 32-column blocks / 4-row K groups of the activation operand (bytes)
    long long units;                      // ntiles * Pout
};

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

// TMEM -> registers, 32 lanes x NV consecutive columns (thread = lane)
template <int NV> __device__ __forceinline__ void tmem_ldn(uint32_t taddr, float* v);
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
__device__ __forceinline__ int fdiv(int x, int d) { return d == 1 ? x : (x >> 1); }                       // floor(x / d), d in {1, 2}
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
__device__ __forceinline__ int q_last(const ConvP& p, const SlabGeom& g, const Seg& s, int pp)
{
    const int q = fdiv(pp * p.pmul + g.dpmax, p.pdiv);
    return q > s.qb ? s.qb : q;
}
// output position that slab q feeds through tap t, or -1
__device__ __forceinline__ int out_pos(const ConvP& p, const Seg& s, int q, int t)
{
    const int num = q * p.pdiv - p.dp[t];
    if (num < 0) return -1;
    if (p.pmul == 2 && (num & 1)) return -1;
    const int pp = p.pmul == 2 ? (num >> 1) : num;
    return (pp >= s.oa && pp < s.ob) ? pp : -1;
}
__device__ __forceinline__ bool has_contrib(const ConvP& p, int pp)
{
    for (int t = 0; t < p.ntaps; ++t) {
        const int num = pp * p.pmul + p.dp[t];
        if (num < 0) continue;
        if (p.pdiv == 2 && (num & 1)) continue;
        const int q = p.pdiv == 2 ? (num >> 1) : num;
        if (q < p.Pin) return true;
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

// =========================================================================================================
// forward / backward-data
// =========================================================================================================
template <int PRO, bool MASK, int CH>
__global__ void __launch_bounds__(NTHR, 1) slab_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                          const ConvP p, const SlabGeom g)
{
    wf_pdl_enter();
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int stage_bytes = 2 * g.half_bytes;
    uint8_t* wsm = smem + g.NS * stage_bytes;                                   // weights: hi images of all taps, then lo
    float* tab = reinterpret_cast<float*>(wsm + 2 * g.w_half_bytes);            // epilogue tables: bias, e_scale, e_shift, e_mean [64] each
    uint64_t* bars = reinterpret_cast<uint64_t*>(tab + 4 * 64);
    const uint32_t bar0 = smem_u32(bars);
    // barrier map: raw_full[NS] | op_full[NS] | slab_empty[NS] | grp_done[NGD] | acc_empty[MAXACC] | w_full
    auto raw_full = [&](int s) { return bar0 + 8u * s; };
    auto op_full = [&](int s) { return bar0 + 8u * (MAXNS + s); };
    auto slab_empty = [&](int s) { return bar0 + 8u * (2 * MAXNS + s); };
    auto grp_done = [&](int s) { return bar0 + 8u * (3 * MAXNS + s); };
    auto acc_empty = [&](int s) { return bar0 + 8u * (3 * MAXNS + NGD + s); };
    const uint32_t w_full = bar0 + 8u * (3 * MAXNS + NGD + MAXACC);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * MAXNS + NGD + MAXACC + 1);

    if (tid == 0) {
        for (int s = 0; s < g.NS; ++s) { mbar_init(raw_full(s), 1); mbar_init(op_full(s), NWW); mbar_init(slab_empty(s), 1); }
        for (int s = 0; s < NGD; ++s) mbar_init(grp_done(s), 1);
        for (int s = 0; s < g.NACC; ++s) mbar_init(acc_empty(s), NWW);
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

    // this CTA's run of (column tile, output position) units
    const long long u0 = g.units * blockIdx.x / gridDim.x, u1 = g.units * (blockIdx.x + 1) / gridDim.x;
    const int R = g.R, PBI = g.PBI, NS = g.NS, NACC = g.NACC;
    const uint32_t smem0 = smem_u32(smem);

    if (warp < NWW) {
        // =============================== workers: transform + epilogue ===============================
        const int NJ = R >> 3;                                   // float4 chunks per thread and stage
        const int ch16 = tid & 7;
        // channels of this thread's rows: row_j = ((tid >> 3) + 32 j) & (R - 1) alternates between two rows at most
        const int rowA = (tid >> 3) & (R - 1), rowB = ((tid >> 3) + 32) & (R - 1);
        const int cA = rowA & (p.Cin - 1), cB = rowB & (p.Cin - 1);
        float4 coA = make_float4(1.f, 0.f, 0.f, 0.f), coB = coA;
        if (PRO != PRO_NONE) {
            coA = make_float4(p.pro_a[cA], p.pro_b[cA], PRO == PRO_BNBWD ? p.pro_c[cA] : 0.f, p.pro_d[cA]);
            coB = make_float4(p.pro_a[cB], p.pro_b[cB], PRO == PRO_BNBWD ? p.pro_c[cB] : 0.f, p.pro_d[cB]);
        }
        // epilogue mapping
        const int lq = warp & 3, half = warp >> 2;
        const int ec0 = half * CH;                               // first channel of this warp half
        const uint32_t t_lane = tmem_base + ((uint32_t)(lq * 32) << 16);
        float s0[CH], s1[CH];
#pragma unroll
        for (int c = 0; c < CH; ++c) { s0[c] = 0.f; s1[c] = 0.f; }
        const float* tb_bias = tab; const float* tb_es = tab + 64; const float* tb_et = tab + 128; const float* tb_em = tab + 192;
        const bool do_stats = p.epi_mode != EPI_STORE && p.stat0 != nullptr;

        long long gg = 0;            // groups processed by this CTA so far (stage / barrier phases)
        long long rbase = 0;         // output positions of earlier segments (accumulator slots)
        // epilogue of the positions completed by a group: deferred by one group so that it overlaps the MMAs of the next one
        struct Pend { bool valid; Seg s; int n0; long long gidx; int p_from, p_to; long long rbase; } pend;
        pend.valid = false;

        auto run_epilogue = [&](const Pend& e) {
            if (e.gidx >= 0) {                                   // gidx < 0: positions no slab contributes to (pure bias / zeros)
                mbar_wait(grp_done((int)(e.gidx % NGD)), (uint32_t)((e.gidx / NGD) & 1));
                tc_fence_after();
            }
            const int n = e.n0 + lq * 32 + lane;
            const bool nv = n < p.N;
            const int b = n / WF_T, t = n - b * WF_T;
            for (int pp = e.p_from; pp < e.p_to; ++pp) {
                const long long r = e.rbase + (pp - e.s.oa);
                const int slot = (int)(r % NACC);
                const bool contrib = has_contrib(p, pp);
                float v[CH];
                if (contrib) {
                    float cr[CH];
                    const uint32_t ta = t_lane + (uint32_t)(slot * 2 * g.NPAD + ec0);
                    constexpr int LC = CH > 16 ? 16 : CH;
#pragma unroll
                    for (int c = 0; c < CH; c += LC) { tmem_ldn<LC>(ta + c, v + c); tmem_ldn<LC>(ta + g.NPAD + c, cr + c); }
                    tmem_ld_wait();
#pragma unroll
                    for (int c = 0; c < CH; ++c) v[c] += cr[c];
                } else {
#pragma unroll
                    for (int c = 0; c < CH; ++c) v[c] = 0.f;
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(acc_empty(slot));
                if (nv) {
                    const long long obase = (long long)pp * p.out_sp + (long long)b * p.out_sb + t;
#pragma unroll
                    for (int c = 0; c < CH; ++c) {
                        const int co = ec0 + c;
                        if (co < p.Cout) {
                            const long long off = (long long)co * p.out_sc + obase;
                            float x = v[c] + tb_bias[co];
                            if (p.accumulate) x += p.out[off];
                            if (p.epi_mode == EPI_STATS) { s0[c] += x; s1[c] = fmaf(x, x, s1[c]); }
                            else if (p.epi_mode == EPI_DSILU || p.epi_mode == EPI_DAFF) {
                                const float rw = p.eraw[off];
                                const float em = tb_em[co];
                                if (p.epi_mode == EPI_DSILU) {
                                    float mk = 1.f;
                                    if (p.emask) mk = p.emask[(long long)b * p.em_sb + (long long)co * p.em_sc + (long long)t * p.em_st];
                                    x = x * mk * wf_dsilu(fmaf(tb_es[co], rw - em, tb_et[co]));
                                }
                                s0[c] += x; s1[c] = fmaf(x, rw - em, s1[c]);
                            }
                            p.out[off] = x;
                        }
                    }
                }
            }
        };

        for (long long u = u0; u < u1;) {
            Seg s; seg_make(p, g, u, u1, s);
            const int n0 = s.ct * TILE;
            // per-segment dropout-mask values of this thread's chunks (Dropout2d: one value per (window, channel))
            float mk[MAXJ];
#pragma unroll
            for (int j = 0; j < MAXJ; ++j) {
                mk[j] = 1.f;
                if (MASK && j < NJ) {
                    const int id = tid + NWORK * j;
                    const int row = (id >> 3) & (R - 1), blk = id / (8 * R);
                    const int n = n0 + blk * 32 + ((ch16 ^ ((row & 3) << 1)) << 2);     // 32-byte chunks are XORed with row mod 4
                    const int c = row & (p.Cin - 1);
                    if (n < p.N) mk[j] = p.mask[(long long)(n / WF_T) * p.m_sb + (long long)c * p.m_sc];
                }
            }
            int p_done = s.oa;
            const int ngroups = s.qb >= s.qa ? (s.qb - s.qa + PBI) / PBI : 0;
            for (int gi = 0; gi < ngroups; ++gi, ++gg) {
                const int st = (int)(gg % NS);
                const uint32_t ph = (uint32_t)((gg / NS) & 1);
                // ---- transform stage st in place ----
                mbar_wait(raw_full(st), ph);
                const uint32_t hi_base = smem0 + (uint32_t)(st * stage_bytes), lo_base = hi_base + (uint32_t)g.half_bytes;
#pragma unroll
                for (int j = 0; j < MAXJ; ++j) {
                    if (j < NJ) {
                        const uint32_t off = (uint32_t)(tid + NWORK * j) * 16u;
                        float4 x = lds4(hi_base + off);
                        float4 x2 = f4zero();
                        if (PRO == PRO_BNBWD) x2 = lds4(lo_base + off);
                        const float4 co = (j & 1) ? coB : coA;
                        x.x = pro1<PRO>(x.x, x2.x, mk[j], co); x.y = pro1<PRO>(x.y, x2.y, mk[j], co);
                        x.z = pro1<PRO>(x.z, x2.z, mk[j], co); x.w = pro1<PRO>(x.w, x2.w, mk[j], co);
                        float4 h, l;
                        tf32_split(x.x, h.x, l.x); tf32_split(x.y, h.y, l.y); tf32_split(x.z, h.z, l.z); tf32_split(x.w, h.w, l.w);
                        sts4(hi_base + off, h);
                        sts4(lo_base + off, l);
                    }
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(op_full(st));
                // ---- epilogue of the previous group's completed positions ----
                if (pend.valid) { run_epilogue(pend); pend.valid = false; }
                // positions completed by this group
                const int ql = min(s.qb, s.qa + (gi + 1) * PBI - 1);
                int p_to = p_done;
                while (p_to < s.ob && q_last(p, g, s, p_to) <= ql) ++p_to;
                pend.valid = true; pend.s = s; pend.n0 = n0; pend.gidx = gg; pend.p_from = p_done; pend.p_to = p_to; pend.rbase = rbase;
                p_done = p_to;
            }
            if (ngroups == 0) {
                // a run of output positions that no input slab feeds (e.g. a single odd position of a stride-2 shortcut's
                // backward-data): nothing to wait for, but the positions are still written and their slots still cycle
                if (pend.valid) { run_epilogue(pend); pend.valid = false; }
                Pend e; e.valid = true; e.s = s; e.n0 = n0; e.gidx = -1; e.p_from = s.oa; e.p_to = s.ob; e.rbase = rbase;
                run_epilogue(e);
            }
            rbase += s.ob - s.oa;
            u += s.ob - s.oa;
        }
        if (pend.valid) run_epilogue(pend);

        // ---- per-channel sums: one cross-lane reduction per CTA, fp64 atomics ----
        if (do_stats) {
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                const float a = warp_sum(s0[c]), bsum = warp_sum(s1[c]);
                const int co = ec0 + c;
                if (lane == 0 && co < p.Cout) { atomicAdd(p.stat0 + co, (double)a); atomicAdd(p.stat1 + co, (double)bsum); }
            }
        }
    } else if (warp == NWW) {
        // =============================== TMA producer ===============================
        if (lane == 0) {
            tma_prefetch_desc(&tmA);
            if (PRO == PRO_BNBWD) tma_prefetch_desc(&tmB);
            // all tap weights (hi + lo images) once
            mbar_arrive_expect_tx(w_full, 2 * g.w_half_bytes);
            bulk_g2s(smem_u32(wsm), p.wtc, 2 * g.w_half_bytes, w_full);
            long long gg = 0;
            const uint32_t tx = (uint32_t)g.half_bytes * (PRO == PRO_BNBWD ? 2u : 1u);
            for (long long u = u0; u < u1;) {
                Seg s; seg_make(p, g, u, u1, s);
                const int n0 = s.ct * TILE;
                const int ngroups = s.qb >= s.qa ? (s.qb - s.qa + PBI) / PBI : 0;
                for (int gi = 0; gi < ngroups; ++gi, ++gg) {
                    const int st = (int)(gg % NS);
                    const uint32_t ph = (uint32_t)((gg / NS) & 1);
                    mbar_wait(slab_empty(st), ph ^ 1u);
                    mbar_arrive_expect_tx(raw_full(st), tx);
                    const uint32_t hi_base = smem0 + (uint32_t)(st * stage_bytes);
                    const int q0 = s.qa + gi * PBI;
#pragma unroll
                    for (int blk = 0; blk < 4; ++blk) {
                        tma_load_3d(hi_base + (uint32_t)(blk * R * 128), &tmA, n0 + blk * 32, 0, q0, raw_full(st));
                        if (PRO == PRO_BNBWD) tma_load_3d(hi_base + (uint32_t)(g.half_bytes + blk * R * 128), &tmB, n0 + blk * 32, 0, q0, raw_full(st));
                    }
                }
                u += s.ob - s.oa;
            }
        }
    } else {
        // =============================== MMA issue ===============================
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_tf32(TILE, g.NPAD, 1, 0);          // A: MN-major (columns contiguous), B: K-major
            const int KS = p.Cin >> 3;
            const uint32_t w_hi = smem_u32(wsm), w_lo = w_hi + (uint32_t)g.w_half_bytes;
            const uint32_t w_tap = (uint32_t)(g.NPAD * p.Cin * 4);                // bytes per tap image
            const uint32_t b_lbo = 128, b_sbo = (uint32_t)(p.Cin / 4) * 128;
            mbar_wait(w_full, 0);
            long long gg = 0, rbase = 0;
            uint32_t started = 0;
            for (long long u = u0; u < u1;) {
                Seg s; seg_make(p, g, u, u1, s);
                int p_done = s.oa;
                const int ngroups = s.qb >= s.qa ? (s.qb - s.qa + PBI) / PBI : 0;
                for (int gi = 0; gi < ngroups; ++gi, ++gg) {
                    const int st = (int)(gg % NS);
                    const uint32_t ph = (uint32_t)((gg / NS) & 1);
                    mbar_wait(op_full(st), ph);
                    tc_fence_after();
                    const uint32_t a_hi = smem0 + (uint32_t)(st * stage_bytes), a_lo = a_hi + (uint32_t)g.half_bytes;
                    const int q0 = s.qa + gi * PBI;
                    const int ql = min(s.qb, q0 + PBI - 1);
                    for (int q = q0; q <= ql; ++q) {
                        const uint32_t row_off = (uint32_t)((q - q0) * p.Cin) * 128u;
                        for (int t = 0; t < p.ntaps; ++t) {
                            const int pp = out_pos(p, s, q, t);
                            if (pp < 0) continue;
                            const long long r = rbase + (pp - s.oa);
                            const int slot = (int)(r % NACC);
                            uint32_t accf = 1u;
                            if (!((started >> slot) & 1u)) {
                                started |= 1u << slot;
                                accf = 0u;
                                const long long use = r / NACC;
                                if (use > 0) { mbar_wait(acc_empty(slot), (uint32_t)((use - 1) & 1)); tc_fence_after(); }
                            }
                            const uint32_t d_main = tmem_base + (uint32_t)(slot * 2 * g.NPAD), d_cor = d_main + (uint32_t)g.NPAD;
                            for (int ks = 0; ks < KS; ++ks) {
                                const uint32_t ao = row_off + (uint32_t)ks * 1024u;
                                const uint64_t dah = umma_desc_l(a_hi + ao, g.lbo_mn, g.sbo_mn, 1), dal = umma_desc_l(a_lo + ao, g.lbo_mn, g.sbo_mn, 1);
                                const uint32_t bo = (uint32_t)t * w_tap + (uint32_t)ks * 256u;
                                const uint64_t dbh = umma_desc_l(w_hi + bo, b_lbo, b_sbo, 0), dbl = umma_desc_l(w_lo + bo, b_lbo, b_sbo, 0);
                                const uint32_t af = (ks == 0) ? accf : 1u;
                                umma_tf32(d_cor, dal, dbh, idesc, af);
                                umma_tf32(d_cor, dah, dbl, idesc, 1u);
                                umma_tf32(d_main, dah, dbh, idesc, af);
                            }
                        }
                    }
                    umma_commit(slab_empty(st));
                    umma_commit(grp_done((int)(gg % NGD)));
                    // positions completed by this group leave the "started" set
                    while (p_done < s.ob && q_last(p, g, s, p_done) <= ql) {
                        const long long r = rbase + (p_done - s.oa);
                        started &= ~(1u << (int)(r % NACC));
                        ++p_done;
                    }
                }
                rbase += s.ob - s.oa;
                u += s.ob - s.oa;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == NWW) tmem_dealloc(tmem_base, g.tmem_cols);
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
    const long long nf = (long long)e.ntaps * f_np * e.cin, nb = (long long)e.ntaps * b_np * e.cout;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nf + nb; i += (long long)gridDim.x * blockDim.x) {
        const bool bwd = i >= nf;
        const long long k_ = bwd ? i - nf : i;
        const int np = bwd ? b_np : f_np, K = bwd ? e.cout : e.cin;
        const int per_tap = np * K;
        const int tap = (int)(k_ / per_tap), within = (int)(k_ % per_tap);
        // image order: [row group (np/8)][k quad (K/4)][row in group (8)][k in quad (4)]
        const int c4 = within & 3, r8 = (within >> 2) & 7, kq = (within >> 5) % (K / 4), rg = within / (8 * K);
        const int row = rg * 8 + r8, kk = kq * 4 + c4;
        float v = 0.f;
        if (!bwd) { if (row < e.cout) v = params[e.param_off + ((long long)row * e.cin + kk) * e.ntaps + tap]; }
        else { if (row < e.cin) v = params[e.param_off + ((long long)kk * e.cin + row) * e.ntaps + tap]; }
        float hi, lo;
        tf32_split(v, hi, lo);
        float* dst = packed + (bwd ? e.bwd_off : e.fwd_off);
        const long long half = (long long)e.ntaps * per_tap;
        dst[k_] = hi;
        dst[half + k_] = lo;
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

int dev_sms()
{
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    return v;
}

constexpr int SMEM_LIMIT = 227 * 1024;

// host-side simulation of the accumulator ring: every output position first touched in group gi must find its slot's previous
// occupant completed by an EARLIER group (its epilogue runs before the workers transform group gi + 1), otherwise the MMA
// thread would wait on an epilogue that waits on the MMA thread
bool ring_ok(const ConvP& p, int dpmin, int dpmax, int PBI, int NACC)
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
        for (int q = qa; q <= qb && qf < 0; ++q)
            for (int t = 0; t < p.ntaps; ++t) {
                const int num = q * p.pdiv - p.dp[t];
                if (num < 0 || (p.pmul == 2 && (num & 1))) continue;
                if ((p.pmul == 2 ? num >> 1 : num) == pp) { qf = q; break; }
            }
        if (qf < 0) continue;
        if (grp_of(qlast(pp - NACC)) >= grp_of(qf)) return false;
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
    g.w_half_bytes = p.ntaps * g.NPAD * p.Cin * 4;
    const int fixed = 2 * g.w_half_bytes + 4 * 64 * 4 + (3 * MAXNS + NGD + MAXACC + 1) * 8 + 16 + 1024;
    int PBI = 64 / p.Cin;
    if (PBI > p.Pin) { PBI = 1; while (PBI * 2 <= p.Pin && PBI * 2 * p.Cin <= 64) PBI *= 2; }
    while (PBI > 1 && !ring_ok(p, dpmin, dpmax, PBI, g.NACC)) PBI >>= 1;
    if (!ring_ok(p, dpmin, dpmax, PBI, g.NACC)) return false;
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
    return (size_t)g.NS * 2 * g.half_bytes + 2 * (size_t)g.w_half_bytes + 4 * 64 * 4 + (3 * MAXNS + NGD + MAXACC + 1) * 8 + 16;
}

template <int PRO, bool MASK, int CH>
cudaError_t launch_t(const CUtensorMap& a, const CUtensorMap& b, const ConvP& p, const SlabGeom& g, int grid, cudaStream_t st)
{
    cudaError_t e = cudaFuncSetAttribute(slab_tc_kernel<PRO, MASK, CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
    if (e != cudaSuccess) return e;
    wf_launch_pdl(slab_tc_kernel<PRO, MASK, CH>, dim3(grid), dim3(NTHR), smem_bytes(g), st, a, b, p, g);
    return cudaGetLastError();
}

template <int PRO, bool MASK>
cudaError_t launch_ch(const CUtensorMap& a, const CUtensorMap& b, const ConvP& p, const SlabGeom& g, int grid, cudaStream_t st)
{
    const int ch = p.Cout <= 8 ? 4 : p.Cout <= 16 ? 8 : p.Cout <= 32 ? 16 : 32;
    switch (ch) {
        case 4: return launch_t<PRO, MASK, 4>(a, b, p, g, grid, st);
        case 8: return launch_t<PRO, MASK, 8>(a, b, p, g, grid, st);
        case 16: return launch_t<PRO, MASK, 16>(a, b, p, g, grid, st);
        default: return launch_t<PRO, MASK, 32>(a, b, p, g, grid, st);
    }
}

const bool g_enabled = [] { const char* e = std::getenv("WF_DISABLE_SLABTC"); return !(e && e[0] == '1'); }();
// debugging aid of the self-test: swap the two stride fields of the MN-major descriptor
const bool g_swap_lbo = [] { const char* e = std::getenv("WF_SLABTC_SWAP_LBO"); return e && e[0] == '1'; }();

}  // namespace

long long wf_slabtc_pack_floats(int cout, int cin, int ntaps, bool bwd)
{
    const int rows = bwd ? cin : cout, k = bwd ? cout : cin;
    const int np = rows <= 16 ? 16 : (rows + 15) / 16 * 16;
    return 2LL * ntaps * np * k;
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
    if (!wf_slabtc_shape_ok(p.Cin, p.Cout, p.groups, p.ntaps, p.dn)) return false;
    if (!(p.pmul == 1 || p.pmul == 2) || !(p.pdiv == 1 || p.pdiv == 2) || (p.pmul == 2 && p.pdiv == 2)) return false;
    if (p.in_sb != WF_T || p.out_sb != WF_T || (p.N & 3) || p.N % WF_T) return false;
    if ((reinterpret_cast<uintptr_t>(p.in) & 15) || ((p.in_sc * 4) & 15) || ((p.in_sp * 4) & 15)) return false;
    if (p.pro_mode == PRO_BNBWD && (!p.in2 || (reinterpret_cast<uintptr_t>(p.in2) & 15))) return false;
    if (p.pro_mode == PRO_BNSILU && p.mask && p.m_st != 0) return false;
    if (p.Pin > 4096 || p.Pout > 4096) return false;
    SlabGeom g;
    return plan(p, g);
}

cudaError_t wf_launch_slabtc_conv(const ConvP& p, cudaStream_t st)
{
    SlabGeom g;
    if (!plan(p, g)) return cudaErrorInvalidValue;
    if (g_swap_lbo) { const int t = g.lbo_mn; g.lbo_mn = g.sbo_mn; g.sbo_mn = t; }
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
