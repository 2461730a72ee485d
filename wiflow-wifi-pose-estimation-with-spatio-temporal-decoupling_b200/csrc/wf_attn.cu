// Axial attention core (models/attention.py:55-68): per (row, group) an LxL attention with head_dim 8,
// no 1/sqrt(d) scaling, BatchNorm2d(8) on the logits, softmax over j, then AV.  L = 20 (width axis, sequences
// along time, rows = (slot h, window b)) or L = 15 (height axis, sequences along slots, rows = n = (b, t)).
//
// One thread owns one query row i of one (row, group): its 8 q values, the L logits, the softmax and the 8 outputs
// live in registers (no shuffles); K and V of the (row, group) are read as float4 broadcasts from a shared-memory
// tile that is staged once per CTA with bn_qkv's affine applied on load.  BatchNorm on the logits needs batch
// statistics, so forward and backward are each two passes: a statistics pass and a main pass that recomputes QK^T
// (0.77 MMAC/sample, cheaper than spilling the 192 KB/sample logits to HBM).
#include <cstdlib>
#include "wf_common.cuh"
#include "wf_elem.h"

namespace {

enum { ATT_FWD_STATS = 0, ATT_FWD = 1, ATT_BWD_STATS = 2, ATT_BWD = 3 };

// A CTA owns RT rows x GH of the 8 groups (blockIdx.y selects which): the groups are independent (group g uses channels
// g*8..g*8+7 of each of q, k, v and of d sv), so narrower CTAs cost no redundant traffic and more of them fit an SM.
constexpr int attn_min_blocks(int nt, int mode)
{
    return mode == ATT_BWD ? (nt <= 160 ? 3 : nt <= 240 ? 2 : 1) : mode == ATT_BWD_STATS ? (nt <= 240 ? 4 : nt <= 320 ? 3 : 2) : 0;
}

template <int L, int LP, int RT, int GH, bool WIDTH, int MODE>
__global__ void __launch_bounds__(RT * GH * L, attn_min_blocks(RT * GH * L, MODE)) attn_kernel(const AttnP p)
{
    wf_pdl_enter();
    constexpr int NT = RT * GH * L;
    constexpr int GC = GH * 8;                          // channels per section (q, k, v, d sv) in this CTA
    constexpr int CS = RT * LP;                         // channel stride inside a tile
    constexpr int TILE = 3 * GC * CS + 3 * GH * 4;      // + 4 floats of padding per group of 8 channels
    constexpr int GTILE = GC * CS + GH * 4;
    extern __shared__ __align__(16) float smem[];
    float* T = smem;                                    // qkv tile (bn_qkv applied)
    float* G = smem + TILE;                             // backward: d sv tile (BN-backward applied)
    float* MX = G + GTILE;                              // backward: two [RT*GH][L][LP] scratch matrices (d logits, probabilities)
    auto tix = [](int c, int r, int s) { return c * CS + (c >> 3) * 4 + r * LP + s; };

    const int tid = threadIdx.x;
    const int N = p.N, B = p.B;
    const long long cstride = 15LL * N;                 // channel stride of [C][15][N] tensors
    const int row0 = blockIdx.x * RT;
    const int g0 = blockIdx.y * GH;                     // first group of this CTA
    auto qkv_chan = [&](int c) { return (c / GC) * 64 + g0 * 8 + (c % GC); };      // tile channel -> channel of the [192] qkv tensor
    const int nrows = WIDTH ? 15 * B : N;

    // global offset of element s of tile row r (excluding the channel term); rows are (h,b) or n
    auto row_base = [&](int r) -> long long {
        const int R = row0 + r;
        if (WIDTH) { const int h = R / B, b = R % B; return (long long)h * N + (long long)b * WF_T; }
        return R;
    };

    // ------------------------------- stage tiles -------------------------------
    // Every global load of the tile is issued before the first transform (the loops are fully unrolled and the guards hold nothing but
    // the load), so a thread has up to NIT + 2*GIT 128-bit requests in flight instead of one load -> use -> store chain per item.
    constexpr int Q = L / 4;                                        // width axis: column quads per row
    constexpr int H = WIDTH ? 1 : RT / 4;                           // height axis: column quads (4 consecutive n) per tile
    constexpr int ITEMS = WIDTH ? 3 * GC * RT * Q : 3 * GC * LP * H;
    constexpr int GITEMS = WIDTH ? GC * RT * Q : GC * LP * H;
    constexpr int NIT = (ITEMS + NT - 1) / NT, GIT = MODE >= ATT_BWD_STATS ? (GITEMS + NT - 1) / NT : 0;
    // height axis: a tile row set is RT = 4 H consecutive n; with H = 2 the two quads of a (channel, slot) are one whole 32-byte sector
    // (with H = 1 every global access of these kernels is a 16-byte piece of its own cache line: 7.5 tag look-ups per request and half
    // of every sector unused, profiles/r2_l1_request_survey.txt)
    static_assert(WIDTH || RT % 4 == 0, "height-axis tiles are whole quads of n wide");
    float4 tv[NIT], gd[GIT > 0 ? GIT : 1], gw[GIT > 0 ? GIT : 1];
    // item -> (tile channel, row / slot, quad) and its global offset without the channel term; valid = inside the tensor
    auto t_item = [&](int idx, int& c, int& r, int& s, long long& off) -> bool {
        if (WIDTH) { const int q = idx % Q; r = (idx / Q) % RT; c = idx / (Q * RT); s = q * 4; off = row_base(r) + s; return row0 + r < nrows; }
        const int h = idx % H; s = (idx / H) % LP; c = idx / (H * LP); r = 4 * h; off = (long long)s * N + row0 + 4 * h;
        return s < L && row0 + 4 * h < nrows;
    };
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
        const int idx = tid + it * NT;
        int c, r, s; long long off;
        tv[it] = f4zero();
        if (idx < ITEMS && t_item(idx, c, r, s, off)) tv[it] = ld4(p.qkv_raw + qkv_chan(c) * cstride + off);
    }
#pragma unroll
    for (int it = 0; it < GIT; ++it) {
        const int idx = tid + it * NT;
        int c, r, s; long long off;
        gd[it] = f4zero(); gw[it] = f4zero();
        if (idx < GITEMS && t_item(idx, c, r, s, off)) {
            gd[it] = ld4(p.dsv + (g0 * 8 + c) * cstride + off);
            gw[it] = ld4(p.sv_raw + (g0 * 8 + c) * cstride + off);
        }
    }
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
        const int idx = tid + it * NT;
        if (idx < ITEMS) {
            int c, r, s; long long off;
            float4 v = tv[it];
            if (t_item(idx, c, r, s, off)) {
                const int cg = qkv_chan(c);
                const float a = p.qkv_scale[cg], b = p.qkv_shift[cg], mu = p.qkv_mean[cg];
                v.x = fmaf(a, v.x - mu, b); v.y = fmaf(a, v.y - mu, b); v.z = fmaf(a, v.z - mu, b); v.w = fmaf(a, v.w - mu, b);
            }
            if (WIDTH) st4(&T[tix(c, r, s)], v);
            else { T[tix(c, r, s)] = v.x; T[tix(c, r + 1, s)] = v.y; T[tix(c, r + 2, s)] = v.z; T[tix(c, r + 3, s)] = v.w; }
        }
    }
#pragma unroll
    for (int it = 0; it < GIT; ++it) {
        const int idx = tid + it * NT;
        if (idx < GITEMS) {
            int c, r, s; long long off;
            float4 v = f4zero();
            if (t_item(idx, c, r, s, off)) {
                const int cg = g0 * 8 + c;
                const float4 d = gd[it], w = gw[it];
                const float a = p.sv_alpha[cg], b = p.sv_beta[cg], e = p.sv_delta[cg], mu = p.sv_mean[cg];
                v.x = fmaf(a, d.x, fmaf(b, w.x - mu, e)); v.y = fmaf(a, d.y, fmaf(b, w.y - mu, e));
                v.z = fmaf(a, d.z, fmaf(b, w.z - mu, e)); v.w = fmaf(a, d.w, fmaf(b, w.w - mu, e));
            }
            if (WIDTH) st4(&G[tix(c, r, s)], v);
            else { G[tix(c, r, s)] = v.x; G[tix(c, r + 1, s)] = v.y; G[tix(c, r + 2, s)] = v.z; G[tix(c, r + 3, s)] = v.w; }
        }
    }
    __syncthreads();

    // ------------------------------- per-thread attention row -------------------------------
    const int i = tid % L, g = (tid / L) % GH, r = tid / (GH * L);
    const int gg0 = g0 + g;                             // group index into the per-group coefficient / statistics arrays
    const bool rvalid = (row0 + r) < nrows;
    float q[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) q[c] = T[tix(g * 8 + c, r, i)];
    float lg[LP];
#pragma unroll
    for (int j = 0; j < LP; ++j) lg[j] = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
#pragma unroll
        for (int j4 = 0; j4 < LP / 4; ++j4) {
            const float4 k4 = ld4(&T[tix(GC + g * 8 + c, r, j4 * 4)]);
            lg[j4 * 4 + 0] = fmaf(q[c], k4.x, lg[j4 * 4 + 0]);
            lg[j4 * 4 + 1] = fmaf(q[c], k4.y, lg[j4 * 4 + 1]);
            lg[j4 * 4 + 2] = fmaf(q[c], k4.z, lg[j4 * 4 + 2]);
            lg[j4 * 4 + 3] = fmaf(q[c], k4.w, lg[j4 * 4 + 3]);
        }
    }

    float st0 = 0.f, st1 = 0.f;          // statistics of this thread (stats passes)
    float pr[LP];                        // softmax probabilities

    if (MODE == ATT_FWD_STATS) {
#pragma unroll
        for (int j = 0; j < L; ++j) { st0 += lg[j]; st1 = fmaf(lg[j], lg[j], st1); }
    } else {
        const float ss = p.sim_scale[gg0], ts = p.sim_shift[gg0], ms = p.sim_mean[gg0];
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < L; ++j) { pr[j] = fmaf(ss, lg[j] - ms, ts); mx = fmaxf(mx, pr[j]); }
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < L; ++j) { pr[j] = __expf(pr[j] - mx); sum += pr[j]; }        // ex2.approx: relative error ~2^-21
        const float inv = 1.f / sum;
#pragma unroll
        for (int j = 0; j < L; ++j) pr[j] *= inv;
#pragma unroll
        for (int j = L; j < LP; ++j) pr[j] = 0.f;
    }

    if (MODE == ATT_FWD) {
        float sv[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            float a = 0.f;
#pragma unroll
            for (int j4 = 0; j4 < LP / 4; ++j4) {
                const float4 v4 = ld4(&T[tix(2 * GC + g * 8 + c, r, j4 * 4)]);
                a = fmaf(pr[j4 * 4 + 0], v4.x, a); a = fmaf(pr[j4 * 4 + 1], v4.y, a);
                a = fmaf(pr[j4 * 4 + 2], v4.z, a); a = fmaf(pr[j4 * 4 + 3], v4.w, a);
            }
            sv[c] = a;
        }
        // the q slot (g*8+c, r, i) is only ever read by this thread: reuse it as the output tile
#pragma unroll
        for (int c = 0; c < 8; ++c) T[tix(g * 8 + c, r, i)] = sv[c];
        __syncthreads();
        if (WIDTH) {
            constexpr int Q = L / 4;
            for (int idx = tid; idx < GC * RT * Q; idx += NT) {
                const int qq = idx % Q, rr = (idx / Q) % RT, c = idx / (Q * RT);
                if (row0 + rr < nrows) st4(p.sv_raw + (g0 * 8 + c) * cstride + row_base(rr) + qq * 4, ld4(&T[tix(c, rr, qq * 4)]));
            }
        } else {
            for (int idx = tid; idx < GC * L * H; idx += NT) {
                const int h = idx % H, s = (idx / H) % L, c = idx / (H * L), r4 = 4 * h;
                if (row0 + r4 < nrows)
                    st4(p.sv_raw + (g0 * 8 + c) * cstride + (long long)s * N + row0 + r4,
                        make_float4(T[tix(c, r4, s)], T[tix(c, r4 + 1, s)], T[tix(c, r4 + 2, s)], T[tix(c, r4 + 3, s)]));
            }
        }
        if (p.sv_s0 && tid < GC) {
            float a = 0.f, b = 0.f;
            for (int rr = 0; rr < RT; ++rr) {
                if (row0 + rr >= nrows) break;
                for (int s = 0; s < L; ++s) { const float v = T[tix(tid, rr, s)]; a += v; b = fmaf(v, v, b); }
            }
            atomicAdd(p.sv_s0 + g0 * 8 + tid, (double)a);
            atomicAdd(p.sv_s1 + g0 * 8 + tid, (double)b);
        }
        return;
    }

    float gs[8];
    if (MODE == ATT_BWD) {
        // Main backward pass.  Register pressure decides the occupancy here, so the logits and the probabilities of this thread's
        // row are parked in two shared-memory matrices as soon as they exist (the dk / dv contractions need them transposed
        // anyway): M holds the row's logits, later overwritten in place by d logits; MP holds the probabilities.
        float* M = MX + (r * GH + g) * (L * LP);
        float* MP = MX + RT * GH * L * LP + (r * GH + g) * (L * LP);
#pragma unroll
        for (int j4 = 0; j4 < LP / 4; ++j4) {
            st4(&M[i * LP + j4 * 4], make_float4(lg[j4 * 4], lg[j4 * 4 + 1], lg[j4 * 4 + 2], lg[j4 * 4 + 3]));
            st4(&MP[i * LP + j4 * 4], make_float4(pr[j4 * 4], pr[j4 * 4 + 1], pr[j4 * 4 + 2], pr[j4 * 4 + 3]));
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) gs[c] = G[tix(g * 8 + c, r, i)];
        // dp_j = sum_c gs[c] v[c][j];  dz_j = p_j (dp_j - sum_k dp_k p_k)
        float dp[LP];
#pragma unroll
        for (int j = 0; j < LP; ++j) dp[j] = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
#pragma unroll
            for (int j4 = 0; j4 < LP / 4; ++j4) {
                const float4 v4 = ld4(&T[tix(2 * GC + g * 8 + c, r, j4 * 4)]);
                dp[j4 * 4 + 0] = fmaf(gs[c], v4.x, dp[j4 * 4 + 0]); dp[j4 * 4 + 1] = fmaf(gs[c], v4.y, dp[j4 * 4 + 1]);
                dp[j4 * 4 + 2] = fmaf(gs[c], v4.z, dp[j4 * 4 + 2]); dp[j4 * 4 + 3] = fmaf(gs[c], v4.w, dp[j4 * 4 + 3]);
            }
        }
        float dot = 0.f;
#pragma unroll
        for (int j = 0; j < L; ++j) dot = fmaf(dp[j], pr[j], dot);
        // d logits through bn_similarity backward, written over the row's logits; dq accumulates on the way
        const float al = p.sim_alpha[gg0], be = p.sim_beta[gg0], de = p.sim_delta[gg0], mu = p.sim_mean[gg0];
        float dq[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) dq[c] = 0.f;
#pragma unroll
        for (int j4 = 0; j4 < LP / 4; ++j4) {
            const float4 l4 = ld4(&M[i * LP + j4 * 4]);
            const float lgv[4] = {l4.x, l4.y, l4.z, l4.w};
            float dl[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int j = j4 * 4 + e;
                dl[e] = (j < L) ? fmaf(al, pr[j] * (dp[j] - dot), fmaf(be, lgv[e] - mu, de)) : 0.f;
            }
            st4(&M[i * LP + j4 * 4], make_float4(dl[0], dl[1], dl[2], dl[3]));
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float4 k4 = ld4(&T[tix(GC + g * 8 + c, r, j4 * 4)]);
                dq[c] = fmaf(dl[0], k4.x, dq[c]); dq[c] = fmaf(dl[1], k4.y, dq[c]);
                dq[c] = fmaf(dl[2], k4.z, dq[c]); dq[c] = fmaf(dl[3], k4.w, dq[c]);
            }
        }
        __syncthreads();
        // dk[c][me] = sum_i' dl[i'][me] q[c][i'];  dv[c][me] = sum_i' p[i'][me] gs[c][i']
        float dk[8], dv[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) { dk[c] = 0.f; dv[c] = 0.f; }
        // four source rows per step: q and d sv arrive as 128-bit broadcasts (one LDS per 4 products instead of one per product;
        // the pass is shared-memory bound), same summation order as the scalar loop
        constexpr int L4 = L / 4 * 4;
#pragma unroll
        for (int i0 = 0; i0 < L4; i0 += 4) {
            float m[4], pp[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) { m[e] = M[(i0 + e) * LP + i]; pp[e] = MP[(i0 + e) * LP + i]; }
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float4 q4 = ld4(&T[tix(g * 8 + c, r, i0)]), g4 = ld4(&G[tix(g * 8 + c, r, i0)]);
                dk[c] = fmaf(m[0], q4.x, dk[c]); dk[c] = fmaf(m[1], q4.y, dk[c]); dk[c] = fmaf(m[2], q4.z, dk[c]); dk[c] = fmaf(m[3], q4.w, dk[c]);
                dv[c] = fmaf(pp[0], g4.x, dv[c]); dv[c] = fmaf(pp[1], g4.y, dv[c]); dv[c] = fmaf(pp[2], g4.z, dv[c]); dv[c] = fmaf(pp[3], g4.w, dv[c]);
            }
        }
#pragma unroll
        for (int ii = L4; ii < L; ++ii) {
            const float m = M[ii * LP + i], pp = MP[ii * LP + i];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                dk[c] = fmaf(m, T[tix(g * 8 + c, r, ii)], dk[c]);
                dv[c] = fmaf(pp, G[tix(g * 8 + c, r, ii)], dv[c]);
            }
        }
        __syncthreads();
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            T[tix(g * 8 + c, r, i)] = dq[c];
            T[tix(GC + g * 8 + c, r, i)] = dk[c];
            T[tix(2 * GC + g * 8 + c, r, i)] = dv[c];
        }
        __syncthreads();
        if (WIDTH) {
            constexpr int Q = L / 4;
            for (int idx = tid; idx < 3 * GC * RT * Q; idx += NT) {
                const int qq = idx % Q, rr = (idx / Q) % RT, c = idx / (Q * RT);
                if (row0 + rr < nrows) st4(p.dqkv + qkv_chan(c) * cstride + row_base(rr) + qq * 4, ld4(&T[tix(c, rr, qq * 4)]));
            }
        } else {
            for (int idx = tid; idx < 3 * GC * L * H; idx += NT) {
                const int h = idx % H, s = (idx / H) % L, c = idx / (H * L), r4 = 4 * h;
                if (row0 + r4 < nrows)
                    st4(p.dqkv + qkv_chan(c) * cstride + (long long)s * N + row0 + r4,
                        make_float4(T[tix(c, r4, s)], T[tix(c, r4 + 1, s)], T[tix(c, r4 + 2, s)], T[tix(c, r4 + 3, s)]));
            }
        }
        return;
    }
    if (MODE == ATT_BWD_STATS) {
#pragma unroll
        for (int c = 0; c < 8; ++c) gs[c] = G[tix(g * 8 + c, r, i)];
        float dp[LP];
#pragma unroll
        for (int j = 0; j < LP; ++j) dp[j] = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
#pragma unroll
            for (int j4 = 0; j4 < LP / 4; ++j4) {
                const float4 v4 = ld4(&T[tix(2 * GC + g * 8 + c, r, j4 * 4)]);
                dp[j4 * 4 + 0] = fmaf(gs[c], v4.x, dp[j4 * 4 + 0]); dp[j4 * 4 + 1] = fmaf(gs[c], v4.y, dp[j4 * 4 + 1]);
                dp[j4 * 4 + 2] = fmaf(gs[c], v4.z, dp[j4 * 4 + 2]); dp[j4 * 4 + 3] = fmaf(gs[c], v4.w, dp[j4 * 4 + 3]);
            }
        }
        float dot = 0.f;
#pragma unroll
        for (int j = 0; j < L; ++j) dot = fmaf(dp[j], pr[j], dot);
        const float mu = p.sim_mean[gg0];
#pragma unroll
        for (int j = 0; j < L; ++j) { const float dz = pr[j] * (dp[j] - dot); st0 += dz; st1 = fmaf(dz, lg[j] - mu, st1); }
    }

    // ------------------------------- statistics reduction (per group) -------------------------------
    if (MODE == ATT_FWD_STATS || MODE == ATT_BWD_STATS) {
        __syncthreads();                                 // everybody is done reading the tile
        float2* part = reinterpret_cast<float2*>(smem);  // reuse the tile
        part[tid] = rvalid ? make_float2(st0, st1) : make_float2(0.f, 0.f);
        __syncthreads();
        if (tid < 2 * GH) {
            const int gg = tid >> 1, which = tid & 1;
            double a = 0;
            for (int rr = 0; rr < RT; ++rr)
                for (int ii = 0; ii < L; ++ii) {
                    const float2 v = part[(rr * GH + gg) * L + ii];
                    a += which ? v.y : v.x;
                }
            double* dst = (MODE == ATT_FWD_STATS) ? (which ? p.sim_s1 : p.sim_s0) : (which ? p.dsim_s1 : p.dsim_s0);
            atomicAdd(dst + g0 + gg, a);
        }
    }
}

template <int L, int LP, int RT, int GH, bool WIDTH, int MODE>
cudaError_t launch_attn(const AttnP& p, cudaStream_t st)
{
    constexpr int CS = RT * LP, GC = GH * 8;
    size_t smem = (3 * GC * CS + 3 * GH * 4) * sizeof(float);
    if (MODE >= ATT_BWD_STATS) smem += (GC * CS + GH * 4) * sizeof(float);
    if (MODE == ATT_BWD) smem += (size_t)2 * RT * GH * L * LP * sizeof(float);
    if (smem < (size_t)RT * GH * L * sizeof(float2)) smem = (size_t)RT * GH * L * sizeof(float2);      // statistics passes reuse the tile
    auto kern = attn_kernel<L, LP, RT, GH, WIDTH, MODE>;
    static WfSmemOptIn optin;
    if (cudaError_t e = wf_smem_optin(optin, kern, smem)) return e;
    const int nrows = WIDTH ? 15 * p.B : p.N;
    wf_launch_pdl(kern, dim3(dim3((nrows + RT - 1) / RT, 8 / GH)), dim3(RT * GH * L), smem, st, p);
    return cudaGetLastError();
}

// CTA shapes (measured, B = 1024): the backward main pass needs ~128 registers per thread, so it runs as narrow CTAs of which two or
// three fit an SM and whose load / compute / store phases interleave; WF_ATTN_CFG=<w><h> (digits) selects other shapes for A/B runs:
// width 0 = 2 rows x 8 groups, 1 = 1 row x 8 groups, 2 = 1 row x 4 groups; height 0 = 8 groups, 1 = 4 groups, 2 = 2 groups.
template <int MODE>
cudaError_t launch_attn_mode(const AttnP& p, cudaStream_t st)
{
    static const int cfg = [] { const char* e = std::getenv("WF_ATTN_CFG"); return e ? std::atoi(e) : -1; }();
    const int wdef = MODE == ATT_BWD ? 1 : 0, hdef = MODE == ATT_BWD ? 1 : 0;
    const int w = cfg >= 0 ? cfg / 10 : wdef, h = cfg >= 0 ? cfg % 10 : hdef;
    if (p.width) {
        if (w == 2) return launch_attn<20, 20, 1, 4, true, MODE>(p, st);
        if (w == 1) return launch_attn<20, 20, 1, 8, true, MODE>(p, st);
        return launch_attn<20, 20, 2, 8, true, MODE>(p, st);
    }
    // default: eight-column tiles with half as many groups per CTA (same thread counts and shared memory as the four-column ones:
    // attention 1.80 -> 1.61 ms per step); WF_ATTN_H8=0 the four-column tiles, 2 sixteen-column tiles
    static const int h8 = [] { const char* e = std::getenv("WF_ATTN_H8"); return e ? std::atoi(e) : 1; }();
    if (h8 == 2) {                 // sixteen-column tiles (64 contiguous bytes per channel and slot)
        if (h >= 1) return launch_attn<15, 16, 16, 1, false, MODE>(p, st);
        return launch_attn<15, 16, 16, 2, false, MODE>(p, st);
    }
    if (h8 == 1) {
        if (h == 2) return launch_attn<15, 16, 8, 1, false, MODE>(p, st);
        if (h == 1) return launch_attn<15, 16, 8, 2, false, MODE>(p, st);
        return launch_attn<15, 16, 8, 4, false, MODE>(p, st);
    }
    if (h == 2) return launch_attn<15, 16, 4, 2, false, MODE>(p, st);
    if (h == 1) return launch_attn<15, 16, 4, 4, false, MODE>(p, st);
    return launch_attn<15, 16, 4, 8, false, MODE>(p, st);
}

}  // namespace

cudaError_t wf_launch_attn_fwd_stats(const AttnP& p, cudaStream_t st) { return launch_attn_mode<ATT_FWD_STATS>(p, st); }
cudaError_t wf_launch_attn_fwd(const AttnP& p, cudaStream_t st) { return launch_attn_mode<ATT_FWD>(p, st); }
cudaError_t wf_launch_attn_bwd_stats(const AttnP& p, cudaStream_t st) { return launch_attn_mode<ATT_BWD_STATS>(p, st); }
cudaError_t wf_launch_attn_bwd(const AttnP& p, cudaStream_t st) { return launch_attn_mode<ATT_BWD>(p, st); }
