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

template <int L, int LP, int RT, bool WIDTH, int MODE>
__global__ void __launch_bounds__(RT * 8 * L, MODE == ATT_BWD_STATS ? (RT * 8 * L <= 320 ? 3 : 2) : (MODE == ATT_BWD ? (RT * 8 * L <= 160 ? 3 : 1) : 0)) attn_kernel(const AttnP p)
{
    constexpr int NT = RT * 8 * L;
    constexpr int CS = RT * LP;                         // channel stride inside a tile
    constexpr int TILE = 192 * CS + 24 * 4;             // + 4 floats of padding per group of 8 channels
    constexpr int GTILE = 64 * CS + 8 * 4;
    extern __shared__ __align__(16) float smem[];
    float* T = smem;                                    // qkv tile (bn_qkv applied)
    float* G = smem + TILE;                             // backward: d sv tile (BN-backward applied)
    float* MX = G + GTILE;                              // backward: two [RT*8][L][LP] scratch matrices (d logits, probabilities)
    auto tix = [](int c, int r, int s) { return c * CS + (c >> 3) * 4 + r * LP + s; };

    const int tid = threadIdx.x;
    const int N = p.N, B = p.B;
    const long long cstride = 15LL * N;                 // channel stride of [C][15][N] tensors
    const int row0 = blockIdx.x * RT;
    const int nrows = WIDTH ? 15 * B : N;

    // global offset of element s of tile row r (excluding the channel term); rows are (h,b) or n
    auto row_base = [&](int r) -> long long {
        const int R = row0 + r;
        if (WIDTH) { const int h = R / B, b = R % B; return (long long)h * N + (long long)b * WF_T; }
        return R;
    };

    // ------------------------------- stage tiles -------------------------------
    if (WIDTH) {
        constexpr int Q = L / 4;
        for (int idx = tid; idx < 192 * RT * Q; idx += NT) {
            const int q = idx % Q, r = (idx / Q) % RT, c = idx / (Q * RT);
            float4 v = f4zero();
            if (row0 + r < nrows) {
                v = ld4(p.qkv_raw + c * cstride + row_base(r) + q * 4);
                const float a = p.qkv_scale[c], b = p.qkv_shift[c], mu = p.qkv_mean[c];
                v.x = fmaf(a, v.x - mu, b); v.y = fmaf(a, v.y - mu, b); v.z = fmaf(a, v.z - mu, b); v.w = fmaf(a, v.w - mu, b);
            }
            st4(&T[tix(c, r, q * 4)], v);
        }
        if (MODE >= ATT_BWD_STATS) {
            for (int idx = tid; idx < 64 * RT * Q; idx += NT) {
                const int q = idx % Q, r = (idx / Q) % RT, c = idx / (Q * RT);
                float4 v = f4zero();
                if (row0 + r < nrows) {
                    const long long off = c * cstride + row_base(r) + q * 4;
                    const float4 d = ld4(p.dsv + off), w = ld4(p.sv_raw + off);
                    const float a = p.sv_alpha[c], b = p.sv_beta[c], e = p.sv_delta[c], mu = p.sv_mean[c];
                    v.x = fmaf(a, d.x, fmaf(b, w.x - mu, e)); v.y = fmaf(a, d.y, fmaf(b, w.y - mu, e));
                    v.z = fmaf(a, d.z, fmaf(b, w.z - mu, e)); v.w = fmaf(a, d.w, fmaf(b, w.w - mu, e));
                }
                st4(&G[tix(c, r, q * 4)], v);
            }
        }
    } else {
        static_assert(WIDTH || RT == 4, "height-axis tiles are 4 consecutive n wide");
        for (int idx = tid; idx < 192 * LP; idx += NT) {
            const int s = idx % LP, c = idx / LP;
            float4 v = f4zero();
            if (s < L && row0 < nrows) {
                v = ld4(p.qkv_raw + c * cstride + (long long)s * N + row0);
                const float a = p.qkv_scale[c], b = p.qkv_shift[c], mu = p.qkv_mean[c];
                v.x = fmaf(a, v.x - mu, b); v.y = fmaf(a, v.y - mu, b); v.z = fmaf(a, v.z - mu, b); v.w = fmaf(a, v.w - mu, b);
            }
            T[tix(c, 0, s)] = v.x; T[tix(c, 1, s)] = v.y; T[tix(c, 2, s)] = v.z; T[tix(c, 3, s)] = v.w;
        }
        if (MODE >= ATT_BWD_STATS) {
            for (int idx = tid; idx < 64 * LP; idx += NT) {
                const int s = idx % LP, c = idx / LP;
                float4 v = f4zero();
                if (s < L && row0 < nrows) {
                    const long long off = c * cstride + (long long)s * N + row0;
                    const float4 d = ld4(p.dsv + off), w = ld4(p.sv_raw + off);
                    const float a = p.sv_alpha[c], b = p.sv_beta[c], e = p.sv_delta[c], mu = p.sv_mean[c];
                    v.x = fmaf(a, d.x, fmaf(b, w.x - mu, e)); v.y = fmaf(a, d.y, fmaf(b, w.y - mu, e));
                    v.z = fmaf(a, d.z, fmaf(b, w.z - mu, e)); v.w = fmaf(a, d.w, fmaf(b, w.w - mu, e));
                }
                G[tix(c, 0, s)] = v.x; G[tix(c, 1, s)] = v.y; G[tix(c, 2, s)] = v.z; G[tix(c, 3, s)] = v.w;
            }
        }
    }
    __syncthreads();

    // ------------------------------- per-thread attention row -------------------------------
    const int i = tid % L, g = (tid / L) % 8, r = tid / (8 * L);
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
            const float4 k4 = ld4(&T[tix(64 + g * 8 + c, r, j4 * 4)]);
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
        const float ss = p.sim_scale[g], ts = p.sim_shift[g], ms = p.sim_mean[g];
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
                const float4 v4 = ld4(&T[tix(128 + g * 8 + c, r, j4 * 4)]);
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
            for (int idx = tid; idx < 64 * RT * Q; idx += NT) {
                const int qq = idx % Q, rr = (idx / Q) % RT, c = idx / (Q * RT);
                if (row0 + rr < nrows) st4(p.sv_raw + c * cstride + row_base(rr) + qq * 4, ld4(&T[tix(c, rr, qq * 4)]));
            }
        } else {
            for (int idx = tid; idx < 64 * L; idx += NT) {
                const int s = idx % L, c = idx / L;
                if (row0 < nrows)
                    st4(p.sv_raw + c * cstride + (long long)s * N + row0,
                        make_float4(T[tix(c, 0, s)], T[tix(c, 1, s)], T[tix(c, 2, s)], T[tix(c, 3, s)]));
            }
        }
        if (p.sv_s0 && tid < 64) {
            float a = 0.f, b = 0.f;
            for (int rr = 0; rr < RT; ++rr) {
                if (row0 + rr >= nrows) break;
                for (int s = 0; s < L; ++s) { const float v = T[tix(tid, rr, s)]; a += v; b = fmaf(v, v, b); }
            }
            atomicAdd(p.sv_s0 + tid, (double)a);
            atomicAdd(p.sv_s1 + tid, (double)b);
        }
        return;
    }

    float gs[8];
    if (MODE == ATT_BWD) {
        // Main backward pass.  Register pressure decides the occupancy here, so the logits and the probabilities of this thread's
        // row are parked in two shared-memory matrices as soon as they exist (the dk / dv contractions need them transposed
        // anyway): M holds the row's logits, later overwritten in place by d logits; MP holds the probabilities.
        float* M = MX + (r * 8 + g) * (L * LP);
        float* MP = MX + RT * 8 * L * LP + (r * 8 + g) * (L * LP);
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
                const float4 v4 = ld4(&T[tix(128 + g * 8 + c, r, j4 * 4)]);
                dp[j4 * 4 + 0] = fmaf(gs[c], v4.x, dp[j4 * 4 + 0]); dp[j4 * 4 + 1] = fmaf(gs[c], v4.y, dp[j4 * 4 + 1]);
                dp[j4 * 4 + 2] = fmaf(gs[c], v4.z, dp[j4 * 4 + 2]); dp[j4 * 4 + 3] = fmaf(gs[c], v4.w, dp[j4 * 4 + 3]);
            }
        }
        float dot = 0.f;
#pragma unroll
        for (int j = 0; j < L; ++j) dot = fmaf(dp[j], pr[j], dot);
        // d logits through bn_similarity backward, written over the row's logits; dq accumulates on the way
        const float al = p.sim_alpha[g], be = p.sim_beta[g], de = p.sim_delta[g], mu = p.sim_mean[g];
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
                const float4 k4 = ld4(&T[tix(64 + g * 8 + c, r, j4 * 4)]);
                dq[c] = fmaf(dl[0], k4.x, dq[c]); dq[c] = fmaf(dl[1], k4.y, dq[c]);
                dq[c] = fmaf(dl[2], k4.z, dq[c]); dq[c] = fmaf(dl[3], k4.w, dq[c]);
            }
        }
        __syncthreads();
        // dk[c][me] = sum_i' dl[i'][me] q[c][i'];  dv[c][me] = sum_i' p[i'][me] gs[c][i']
        float dk[8], dv[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) { dk[c] = 0.f; dv[c] = 0.f; }
        for (int ii = 0; ii < L; ++ii) {
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
            T[tix(64 + g * 8 + c, r, i)] = dk[c];
            T[tix(128 + g * 8 + c, r, i)] = dv[c];
        }
        __syncthreads();
        if (WIDTH) {
            constexpr int Q = L / 4;
            for (int idx = tid; idx < 192 * RT * Q; idx += NT) {
                const int qq = idx % Q, rr = (idx / Q) % RT, c = idx / (Q * RT);
                if (row0 + rr < nrows) st4(p.dqkv + c * cstride + row_base(rr) + qq * 4, ld4(&T[tix(c, rr, qq * 4)]));
            }
        } else {
            for (int idx = tid; idx < 192 * L; idx += NT) {
                const int s = idx % L, c = idx / L;
                if (row0 < nrows)
                    st4(p.dqkv + c * cstride + (long long)s * N + row0,
                        make_float4(T[tix(c, 0, s)], T[tix(c, 1, s)], T[tix(c, 2, s)], T[tix(c, 3, s)]));
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
                const float4 v4 = ld4(&T[tix(128 + g * 8 + c, r, j4 * 4)]);
                dp[j4 * 4 + 0] = fmaf(gs[c], v4.x, dp[j4 * 4 + 0]); dp[j4 * 4 + 1] = fmaf(gs[c], v4.y, dp[j4 * 4 + 1]);
                dp[j4 * 4 + 2] = fmaf(gs[c], v4.z, dp[j4 * 4 + 2]); dp[j4 * 4 + 3] = fmaf(gs[c], v4.w, dp[j4 * 4 + 3]);
            }
        }
        float dot = 0.f;
#pragma unroll
        for (int j = 0; j < L; ++j) dot = fmaf(dp[j], pr[j], dot);
        const float mu = p.sim_mean[g];
#pragma unroll
        for (int j = 0; j < L; ++j) { const float dz = pr[j] * (dp[j] - dot); st0 += dz; st1 = fmaf(dz, lg[j] - mu, st1); }
    }

    // ------------------------------- statistics reduction (per group) -------------------------------
    if (MODE == ATT_FWD_STATS || MODE == ATT_BWD_STATS) {
        __syncthreads();                                 // everybody is done reading the tile
        float2* part = reinterpret_cast<float2*>(smem);  // reuse the tile
        part[tid] = rvalid ? make_float2(st0, st1) : make_float2(0.f, 0.f);
        __syncthreads();
        if (tid < 16) {
            const int gg = tid >> 1, which = tid & 1;
            double a = 0;
            for (int rr = 0; rr < RT; ++rr)
                for (int ii = 0; ii < L; ++ii) {
                    const float2 v = part[(rr * 8 + gg) * L + ii];
                    a += which ? v.y : v.x;
                }
            double* dst = (MODE == ATT_FWD_STATS) ? (which ? p.sim_s1 : p.sim_s0) : (which ? p.dsim_s1 : p.dsim_s0);
            atomicAdd(dst + gg, a);
        }
    }
}

template <int L, int LP, int RT, bool WIDTH, int MODE>
cudaError_t launch_attn(const AttnP& p, cudaStream_t st)
{
    constexpr int CS = RT * LP;
    size_t smem = (192 * CS + 24 * 4) * sizeof(float);
    if (MODE >= ATT_BWD_STATS) smem += (64 * CS + 8 * 4) * sizeof(float);
    if (MODE == ATT_BWD) smem += (size_t)2 * RT * 8 * L * LP * sizeof(float);
    auto kern = attn_kernel<L, LP, RT, WIDTH, MODE>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    const int nrows = WIDTH ? 15 * p.B : p.N;
    kern<<<(nrows + RT - 1) / RT, RT * 8 * L, smem, st>>>(p);
    return cudaGetLastError();
}

template <int MODE>
cudaError_t launch_attn_mode(const AttnP& p, cudaStream_t st)
{
    // backward main pass of the width axis: one row per CTA (160 threads, three CTAs per SM whose staging / compute / store phases
    // interleave) instead of two rows in one 320-thread CTA per SM
    static const bool rt1 = [] { const char* e = std::getenv("WF_ATTN_BWD_RT1"); return !(e && e[0] == '0'); }();
    if (p.width && MODE == ATT_BWD && rt1) return launch_attn<20, 20, 1, true, MODE>(p, st);
    if (p.width) return launch_attn<20, 20, 2, true, MODE>(p, st);
    return launch_attn<15, 16, 4, false, MODE>(p, st);
}

}  // namespace

cudaError_t wf_launch_attn_fwd_stats(const AttnP& p, cudaStream_t st) { return launch_attn_mode<ATT_FWD_STATS>(p, st); }
cudaError_t wf_launch_attn_fwd(const AttnP& p, cudaStream_t st) { return launch_attn_mode<ATT_FWD>(p, st); }
cudaError_t wf_launch_attn_bwd_stats(const AttnP& p, cudaStream_t st) { return launch_attn_mode<ATT_BWD_STATS>(p, st); }
cudaError_t wf_launch_attn_bwd(const AttnP& p, cudaStream_t st) { return launch_attn_mode<ATT_BWD>(p, st); }
