// Element-wise / reduction kernels of the WiFlow path: BatchNorm statistic finalisation (forward and backward),
// residual joins, average pool, pose loss, PCK/MPJPE, weight packing, clip+AdamW and layout permutes.
#include "wf_common.cuh"
#include "wf_elem.h"

namespace {

// ---------------------------------------------------------------------------------------------------------
// BatchNorm finalisation (train mode).  Forward: batch mean / biased var -> (scale, shift, mean, rstd) with
// y = scale*(x - mean) + shift, shift = beta: subtracting the mean BEFORE scaling (as torch does) keeps the normalised value
// free of the per-channel rounding offset a folded `beta - mean*scale` would add to every element of the channel; running
// stats with the unbiased variance and momentum 0.1 (torch BatchNorm semantics used by every BN
// in models/*.py).  Backward: (sum dy, sum dy*raw) -> dgamma, dbeta and the affine (alpha, beta, delta) with
// dx = alpha*dy + beta*raw + delta.
// ---------------------------------------------------------------------------------------------------------
__global__ void bn_finalize_fwd_kernel(BnFwdFin a, BnFwdFin b, int n)
{
    wf_pdl_enter();
    const BnFwdFin& d = (blockIdx.y == 0) ? a : b;
    if ((int)blockIdx.y >= n) return;
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= d.C) return;
    wf_bn_fwd_fin_channel(d, c);
}

__global__ void bn_finalize_bwd_kernel(BnBwdFin a, BnBwdFin b, int n)
{
    wf_pdl_enter();
    const BnBwdFin& d = (blockIdx.y == 0) ? a : b;
    if ((int)blockIdx.y >= n) return;
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= d.C) return;
    wf_bn_bwd_fin_channel(d, c);
}

// eval mode: (scale, shift) of all BatchNorms from the running statistics, one launch
__global__ void bn_eval_coefs_kernel(BnEvalTable tab, const float* params, const float* running, float* coefs)
{
    wf_pdl_enter();
    const BnEvalEntry e = tab.e[blockIdx.y];
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if ((int)blockIdx.y >= tab.n || c >= e.C) return;
    float gam = params[e.gamma_off + c], bet = params[e.gamma_off + e.C + c];
    float rm = running[e.run_off + c], rv = running[e.run_off + e.C + c];
    float rstd = 1.0f / sqrtf(rv + 1e-5f);
    float sc = gam * rstd;
    coefs[e.coef_off + c] = sc;                       // scale
    coefs[e.coef_off + e.Cpad + c] = bet;             // shift (y = scale*(x - mean) + shift)
    coefs[e.coef_off + 2 * e.Cpad + c] = rm;          // mean
    coefs[e.coef_off + 3 * e.Cpad + c] = rstd;        // rstd
}

// ---------------------------------------------------------------------------------------------------------
// Dropout masks of a whole training step in one launch (perf mode of models/tcn.py:30,43 nn.Dropout and models/convnet.py:15,20
// nn.Dropout2d; the parity mode draws them with torch's generator, block.py).  Site i: out[k] = u >= p ? 1/(1-p) : 0 for k < numel,
// u = Philox4x32-10(counter = (k/4, draw), key = seed ^ site)[k%4] * 2^-32.  `draw` lives on the device (state[0]) and is advanced
// by the last CTA to finish, so every replay of a captured graph draws fresh masks; state[1] is that CTA ticket.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const unsigned hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u; k.y += 0xBB67AE85u;
    }
    return c;
}

__global__ void __launch_bounds__(256) dropout_masks_kernel(MaskTable tab, unsigned long long seed, unsigned long long* state)
{
    wf_pdl_enter();
    const unsigned long long draw = __ldcg(state);
    const MaskSite st = tab.s[blockIdx.y];
    const unsigned long long ks = seed ^ (0x9E3779B97F4A7C15ull * (unsigned long long)(blockIdx.y + 1));
    const uint2 key = make_uint2((unsigned)ks, (unsigned)(ks >> 32));
    const float keep = 1.f / (1.f - st.p);
    const long long n4 = (st.numel + 3) / 4;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += (long long)gridDim.x * blockDim.x) {
        const uint4 r = philox4x32_10(make_uint4((unsigned)q, (unsigned)(q >> 32), (unsigned)draw, (unsigned)(draw >> 32)), key);
        float4 m;
        m.x = (float)r.x * 2.3283064365386963e-10f >= st.p ? keep : 0.f;
        m.y = (float)r.y * 2.3283064365386963e-10f >= st.p ? keep : 0.f;
        m.z = (float)r.z * 2.3283064365386963e-10f >= st.p ? keep : 0.f;
        m.w = (float)r.w * 2.3283064365386963e-10f >= st.p ? keep : 0.f;
        if (q * 4 + 3 < st.numel) st4(st.out + q * 4, m);
        else {
            const float v[4] = {m.x, m.y, m.z, m.w};
            for (int j = 0; j < 4 && q * 4 + j < st.numel; ++j) st.out[q * 4 + j] = v[j];
        }
    }
    // every CTA has read `draw` before it takes a ticket, so the last one may advance it
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned long long total = (unsigned long long)gridDim.x * gridDim.y;
        if (atomicAdd(state + 1, 1ull) == total - 1) { state[1] = 0ull; state[0] = draw + 1ull; }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Residual joins.  TCN: silu(mask*silu(bn(pw2)) + res) (models/tcn.py:74), conv blocks: silu(bn(c3) + bn(ds))
// (models/convnet.py:36-37,72-73).  One channel per blockIdx.y, float4 over the [P][N] plane.
// ---------------------------------------------------------------------------------------------------------
// Per-channel state of a join CTA: the coefficients are read once, element offsets are 32-bit (a plane is at most 240 x 20 480
// elements), and the loads of an iteration are separated from its arithmetic so that two float4 groups are in flight per thread.
struct JoinCh {
    const float *a, *r, *mask;
    float as, at, am, rs, rt, rm;
    unsigned N, r_sp, r_sb, m_sb, m_st;
    bool silu, masked;
};
__device__ __forceinline__ JoinCh join_channel(const JoinP& p, int c)
{
    JoinCh k;
    k.a = p.a + (long long)c * p.plane;
    k.r = p.r + (long long)c * p.r_sc;
    k.silu = p.a_mode == PRO_BNSILU;
    k.masked = k.silu && p.mask != nullptr;
    k.mask = k.masked ? p.mask + (long long)c * p.m_sc : nullptr;
    k.as = p.a_scale[c]; k.at = p.a_shift[c]; k.am = p.a_mean[c];
    k.rs = 1.f; k.rt = 0.f; k.rm = 0.f;
    if (p.r_mode == PRO_AFFINE) { k.rs = p.r_scale[c]; k.rt = p.r_shift[c]; k.rm = p.r_mean[c]; }
    k.N = (unsigned)p.N; k.r_sp = (unsigned)p.r_sp; k.r_sb = (unsigned)p.r_sb; k.m_sb = (unsigned)p.m_sb; k.m_st = (unsigned)p.m_st;
    return k;
}
struct JoinLd { float4 a4, r4, m4; };
__device__ __forceinline__ void join_load(const JoinCh& k, unsigned i, JoinLd& l)
{
    const unsigned pos = i / k.N, n = i - pos * k.N;
    const unsigned b = n / WF_T, t = n - b * WF_T;
    l.a4 = ld4(k.a + i);
    l.r4 = ld4(k.r + (pos * k.r_sp + b * k.r_sb + t));
    l.m4 = make_float4(1.f, 1.f, 1.f, 1.f);
    if (k.masked) {
        const float* mp = k.mask + (b * k.m_sb + t * k.m_st);
        if (k.m_st == 1) l.m4 = ld4(mp); else { const float m = *mp; l.m4 = make_float4(m, m, m, m); }
    }
}
__device__ __forceinline__ void join_math(const JoinCh& k, const JoinLd& l, float ya[4], float mk[4], float z[4])
{
    const float a[4] = {l.a4.x, l.a4.y, l.a4.z, l.a4.w};
    const float r[4] = {l.r4.x, l.r4.y, l.r4.z, l.r4.w};
    mk[0] = l.m4.x; mk[1] = l.m4.y; mk[2] = l.m4.z; mk[3] = l.m4.w;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        ya[j] = fmaf(k.as, a[j] - k.am, k.at);
        const float av = k.silu ? mk[j] * wf_silu(ya[j]) : ya[j];
        z[j] = av + fmaf(k.rs, r[j] - k.rm, k.rt);
    }
}

constexpr int JOIN_UNR = 2;

__global__ void __launch_bounds__(256) join_fwd_kernel(JoinP p)
{
    wf_pdl_enter();
    const int c = blockIdx.y;
    const JoinCh k = join_channel(p, c);
    const unsigned total4 = (unsigned)(p.plane / 4), stride = gridDim.x * blockDim.x;
    float* out = p.out + (long long)c * p.plane;
    for (unsigned q0 = blockIdx.x * blockDim.x + threadIdx.x; q0 < total4; q0 += JOIN_UNR * stride) {
        JoinLd l[JOIN_UNR];
#pragma unroll
        for (int u = 0; u < JOIN_UNR; ++u) if (q0 + u * stride < total4) join_load(k, (q0 + u * stride) * 4, l[u]);
#pragma unroll
        for (int u = 0; u < JOIN_UNR; ++u) {
            if (q0 + u * stride < total4) {
                float ya[4], mk[4], z[4];
                join_math(k, l[u], ya, mk, z);
                st4(out + (q0 + u * stride) * 4, make_float4(wf_silu(z[0]), wf_silu(z[1]), wf_silu(z[2]), wf_silu(z[3])));
            }
        }
    }
}

// backward of the join: dz = dout*silu'(z); dya = dz*mask*silu'(ya) (TCN) or dz (conv blocks); the shortcut branch
// receives dz.  Accumulates the BatchNorm-backward sums of both branches.
template <int NT>
__global__ void __launch_bounds__(NT) join_bwd_kernel(JoinP p)
{
    wf_pdl_enter();
    const int c = blockIdx.y;
    const JoinCh k = join_channel(p, c);
    const unsigned total4 = (unsigned)(p.plane / 4), stride = gridDim.x * NT;
    float sa0 = 0.f, sa1 = 0.f, sr0 = 0.f, sr1 = 0.f;
    const float rmean = p.r_mean ? p.r_mean[c] : 0.f;       // centre of the shortcut's BatchNorm-backward sum (also when r enters unscaled)
    const float* dout = p.dout + (long long)c * p.plane;
    float* dzp = p.dz + (long long)c * p.plane;
    float* dap = p.da ? p.da + (long long)c * p.plane : nullptr;
    for (unsigned q0 = blockIdx.x * NT + threadIdx.x; q0 < total4; q0 += JOIN_UNR * stride) {
        JoinLd l[JOIN_UNR];
        float4 g4[JOIN_UNR];
#pragma unroll
        for (int u = 0; u < JOIN_UNR; ++u)
            if (q0 + u * stride < total4) { join_load(k, (q0 + u * stride) * 4, l[u]); g4[u] = ld4(dout + (q0 + u * stride) * 4); }
#pragma unroll
        for (int u = 0; u < JOIN_UNR; ++u) {
            if (q0 + u * stride < total4) {
                float ya[4], mk[4], z[4];
                join_math(k, l[u], ya, mk, z);
                const float g[4] = {g4[u].x, g4[u].y, g4[u].z, g4[u].w};
                const float a[4] = {l[u].a4.x, l[u].a4.y, l[u].a4.z, l[u].a4.w};
                const float r[4] = {l[u].r4.x, l[u].r4.y, l[u].r4.z, l[u].r4.w};
                float dz[4], da[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    dz[j] = g[j] * wf_dsilu(z[j]);
                    da[j] = k.silu ? dz[j] * mk[j] * wf_dsilu(ya[j]) : dz[j];
                    sa0 += da[j]; sa1 = fmaf(da[j], a[j] - k.am, sa1);
                    sr0 += dz[j]; sr1 = fmaf(dz[j], r[j] - rmean, sr1);
                }
                st4(dzp + (q0 + u * stride) * 4, make_float4(dz[0], dz[1], dz[2], dz[3]));
                if (dap) st4(dap + (q0 + u * stride) * 4, make_float4(da[0], da[1], da[2], da[3]));
            }
        }
    }
    block_accum2<NT>(sa0, sa1, p.a_stat0 + c, p.a_stat1 + c);
    if (p.r_stat0) block_accum2<NT>(sr0, sr1, p.r_stat0 + c, p.r_stat1 + c);
    wf_bn_tail(p.tail);
}

// generic per-channel sums (sum dy, sum dy*raw) over [C][plane]
template <int NT>
__global__ void __launch_bounds__(NT) bn_bwd_stats_kernel(const float* dy, const float* raw, const float* mean, long long plane, double* s0, double* s1)
{
    wf_pdl_enter();
    const int c = blockIdx.y;
    const float mu = mean[c];
    float a = 0.f, b = 0.f;
    const long long total4 = plane / 4;
    for (long long q = (long long)blockIdx.x * NT + threadIdx.x; q < total4; q += (long long)gridDim.x * NT) {
        const float4 d = ld4(dy + (long long)c * plane + q * 4), r = ld4(raw + (long long)c * plane + q * 4);
        a += d.x + d.y + d.z + d.w;
        b = fmaf(d.x, r.x - mu, fmaf(d.y, r.y - mu, fmaf(d.z, r.z - mu, fmaf(d.w, r.w - mu, b))));
    }
    block_accum2<NT>(a, b, s0 + c, s1 + c);
}

// ---------------------------------------------------------------------------------------------------------
// decoder tail: BN + SiLU + mean over the 20 time steps (pose_model.py:49-53,94-95).  raw [2][15][N] -> pred [B][15][2]
// ---------------------------------------------------------------------------------------------------------
__global__ void pool_fwd_kernel(const float* raw, const float* scale, const float* shift, const float* mean, float* pred, int B)
{
    wf_pdl_enter();
    int i = blockIdx.x * blockDim.x + threadIdx.x;       // (o, j, b)
    if (i >= 2 * 15 * B) return;
    const int b = i % B, j = (i / B) % 15, o = i / (15 * B);
    const float* src = raw + ((long long)(o * 15 + j) * B + b) * WF_T;
    const float s = scale[o], t = shift[o], mu = mean[o];
    float acc = 0.f;
#pragma unroll
    for (int q = 0; q < WF_T / 4; ++q) {
        float4 v = ld4(src + q * 4);
        acc += wf_silu(fmaf(s, v.x - mu, t)) + wf_silu(fmaf(s, v.y - mu, t)) + wf_silu(fmaf(s, v.z - mu, t)) + wf_silu(fmaf(s, v.w - mu, t));
    }
    pred[((long long)b * 15 + j) * 2 + o] = acc * (1.0f / WF_T);
}

template <int NT>
__global__ void __launch_bounds__(NT) pool_bwd_kernel(const float* raw, const float* scale, const float* shift, const float* mean, const float* dpred,
                                                      float* dy, int B, double* s0, double* s1)
{
    wf_pdl_enter();
    const int o = blockIdx.y;
    const float s = scale[o], t = shift[o], mu = mean[o];
    float a0 = 0.f, a1 = 0.f;
    const int total = 15 * B;
    for (int i = blockIdx.x * NT + threadIdx.x; i < total; i += gridDim.x * NT) {   // (j, b)
        const int b = i % B, j = i / B;
        const float g = dpred[((long long)b * 15 + j) * 2 + o] * (1.0f / WF_T);
        const long long off = ((long long)(o * 15 + j) * B + b) * WF_T;
#pragma unroll
        for (int q = 0; q < WF_T / 4; ++q) {
            const float4 v = ld4(raw + off + q * 4);
            float4 d;
            d.x = g * wf_dsilu(fmaf(s, v.x - mu, t)); d.y = g * wf_dsilu(fmaf(s, v.y - mu, t));
            d.z = g * wf_dsilu(fmaf(s, v.z - mu, t)); d.w = g * wf_dsilu(fmaf(s, v.w - mu, t));
            a0 += d.x + d.y + d.z + d.w;
            a1 = fmaf(d.x, v.x - mu, fmaf(d.y, v.y - mu, fmaf(d.z, v.z - mu, fmaf(d.w, v.w - mu, a1))));
            st4(dy + off + q * 4, d);
        }
    }
    block_accum2<NT>(a0, a1, s0 + o, s1 + o);
}

// ---------------------------------------------------------------------------------------------------------
// Pose loss (losses/pose_loss.py:26-88): position term + bone-length term, three loss types.
// One thread per sample; sums in fp64; optional gradient dpred = gscale * d total / d pred.
// ---------------------------------------------------------------------------------------------------------
__constant__ int c_bone_s[14] = {0, 1, 1, 2, 3, 1, 5, 6, 8, 8, 9, 10, 12, 13};
__constant__ int c_bone_e[14] = {1, 8, 2, 3, 4, 5, 6, 7, 9, 12, 10, 11, 13, 14};

__device__ __forceinline__ void loss_term(int type, float beta, float d, float& val, float& der)
{
    if (type == WF_LOSS_MSE) { val = d * d; der = 2.f * d; }
    else if (type == WF_LOSS_L1) { val = fabsf(d); der = (d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f); }
    else {
        float ad = fabsf(d);
        if (ad < beta) { val = 0.5f * d * d / beta; der = d / beta; }
        else { val = ad - 0.5f * beta; der = (d > 0.f) ? 1.f : -1.f; }
    }
}

template <int NT>
__global__ void __launch_bounds__(NT) pose_loss_kernel(const float* pred, const float* target, int B, int type, float pw, float bw,
                                                       const float* gscale, float* dpred, double* acc)
{
    wf_pdl_enter();
    float ps = 0.f, bs = 0.f;
    const int b = blockIdx.x * NT + threadIdx.x;
    if (b < B) {
        float p[30], t[30], g[30];
#pragma unroll
        for (int i = 0; i < 30; ++i) { p[i] = pred[b * 30 + i]; t[i] = target[b * 30 + i]; }
        const float gs = gscale ? *gscale : 1.f;
        const float wp = pw / (30.f * B), wb = bw / (14.f * B);
#pragma unroll
        for (int i = 0; i < 30; ++i) {
            float v, d;
            loss_term(type, 0.1f, p[i] - t[i], v, d);
            ps += v;
            g[i] = wp * d;
        }
#pragma unroll
        for (int k = 0; k < 14; ++k) {
            const int s = c_bone_s[k], e = c_bone_e[k];
            const float vx = p[2 * e] - p[2 * s], vy = p[2 * e + 1] - p[2 * s + 1];
            const float tx = t[2 * e] - t[2 * s], ty = t[2 * e + 1] - t[2 * s + 1];
            const float lp = sqrtf(vx * vx + vy * vy + 1e-8f), lt = sqrtf(tx * tx + ty * ty + 1e-8f);
            float v, d;
            loss_term(type, 0.05f, lp - lt, v, d);
            bs += v;
            const float gx = wb * d * vx / lp, gy = wb * d * vy / lp;
            g[2 * e] += gx; g[2 * e + 1] += gy; g[2 * s] -= gx; g[2 * s + 1] -= gy;
        }
        if (dpred) {
#pragma unroll
            for (int i = 0; i < 30; ++i) dpred[b * 30 + i] = gs * g[i];
        }
    }
    block_accum2<NT>(ps, bs, acc, acc + 1);
}

__global__ void pose_loss_finish_kernel(double* acc, int B, float pw, float bw, float* out3)
{
    wf_pdl_enter();
    const float pos = (float)(acc[0] / (30.0 * B)), bone = (float)(acc[1] / (14.0 * B));
    out3[0] = pw * pos + bw * bone;
    out3[1] = pos;
    out3[2] = bone;
    acc[0] = 0; acc[1] = 0;
}

// ---------------------------------------------------------------------------------------------------------
// PCK / MPJPE (utils/metrics.py:3-47).  counts are exact integers; distances summed in fp64.
// ---------------------------------------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(NT) metrics_kernel(const float* pred, const float* target, int B, MetricThr thr, int ia, int ib,
                                                     unsigned long long* counts, double* dsum)
{
    wf_pdl_enter();
    __shared__ unsigned int scnt[WF_MAX_THR];
    if (threadIdx.x < WF_MAX_THR) scnt[threadIdx.x] = 0;
    __syncthreads();
    float ds = 0.f;
    const int b = blockIdx.x * NT + threadIdx.x;
    if (b < B) {
        const float* p = pred + b * 30;
        const float* t = target + b * 30;
        const float nx = t[2 * ia] - t[2 * ib], ny = t[2 * ia + 1] - t[2 * ib + 1];
        float norm = sqrtf(nx * nx + ny * ny);
        norm = fmaxf(norm, 0.01f);
        unsigned int cnt[WF_MAX_THR];
#pragma unroll
        for (int k = 0; k < WF_MAX_THR; ++k) cnt[k] = 0;
        for (int j = 0; j < 15; ++j) {
            const float dx = p[2 * j] - t[2 * j], dy = p[2 * j + 1] - t[2 * j + 1];
            const float d = sqrtf(dx * dx + dy * dy);
            ds += d;
            const float nd = d / norm;
#pragma unroll
            for (int k = 0; k < WF_MAX_THR; ++k) if (k < thr.n && nd <= thr.v[k]) cnt[k]++;
        }
#pragma unroll
        for (int k = 0; k < WF_MAX_THR; ++k) if (k < thr.n && cnt[k]) atomicAdd(&scnt[k], cnt[k]);
    }
    __syncthreads();
    if (threadIdx.x < thr.n && scnt[threadIdx.x]) atomicAdd(counts + threadIdx.x, (unsigned long long)scnt[threadIdx.x]);
    double d = warp_sum_d((double)ds);
    if ((threadIdx.x & 31) == 0 && d != 0.0) atomicAdd(dsum, d);
}

__global__ void metrics_finish_kernel(unsigned long long* counts, double* dsum, int B, int nthr, float* out)
{
    wf_pdl_enter();
    const int k = threadIdx.x;
    if (k < nthr) { out[k] = (float)counts[k] / (float)(15 * B); counts[k] = 0; }
    if (k == 0) { out[nthr] = (float)(*dsum / (15.0 * B)); *dsum = 0; }
}

// ---------------------------------------------------------------------------------------------------------
// Weight packing: reference layout [Cout_total][Cin_g][ntaps] -> forward pack [g][tap][ci(Kpad)][co(Mpad)] and
// backward-data pack [g][tap][co(Kpad)][ci(Mpad)] (padding pre-zeroed by the caller).
// ---------------------------------------------------------------------------------------------------------
__global__ void pack_weights_kernel(PackTable tab, const float* params, float* packed)
{
    wf_pdl_enter();
    const PackEntry e = tab.e[blockIdx.y];
    const int cout_g = e.cout / e.groups;
    const int total = e.cout * e.cin * e.ntaps;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int tap = i % e.ntaps, ci = (i / e.ntaps) % e.cin, co = i / (e.ntaps * e.cin);
        const int g = co / cout_g, col = co % cout_g;
        const float v = params[e.param_off + i];
        packed[e.fwd_off + ((size_t)(g * e.ntaps + tap) * e.f_kpad + ci) * e.f_mpad + col] = v;
        if (e.bwd_off >= 0)
            packed[e.bwd_off + ((size_t)(g * e.ntaps + tap) * e.b_kpad + col) * e.b_mpad + ci] = v;
    }
}

// ---------------------------------------------------------------------------------------------------------
// clip_grad_norm_(max_norm) + AdamW over the flat parameter buffer (train.py:105-110,235-236)
// ---------------------------------------------------------------------------------------------------------
// Deterministic: a fixed grid (WF_SUMSQ_BLOCKS, independent of the SM count), a fixed element -> thread assignment, a fixed
// reduction tree per block and one partial per block; adamw_prep_kernel adds the partials in index order.  No atomics, so every
// rank of a data-parallel job computes bit-identical clip coefficients from the all-reduced gradient (replicas stay in lock step).
template <int NT>
__global__ void __launch_bounds__(NT) sumsq_kernel(const float* g, long long n, double* partial)
{
    wf_pdl_enter();
    float a = 0.f, b = 0.f;
    const long long n4 = n / 4;
    for (long long q = (long long)blockIdx.x * NT + threadIdx.x; q < n4; q += (long long)gridDim.x * NT) {
        const float4 v = ld4(g + q * 4);
        a = fmaf(v.x, v.x, fmaf(v.y, v.y, a));
        b = fmaf(v.z, v.z, fmaf(v.w, v.w, b));
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (long long i = n4 * 4; i < n; ++i) a = fmaf(g[i], g[i], a);
    double d = warp_sum_d((double)a + (double)b);
    __shared__ double red[NT / 32];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = d;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0;
        for (int i = 0; i < NT / 32; ++i) s += red[i];
        partial[blockIdx.x] = s;
    }
}

__global__ void adamw_prep_kernel(AdamState* st, float lr, float b1, float b2, float max_norm, float grad_scale)
{
    wf_pdl_enter();
    st->step += 1;
    double ss = 0;
    for (int i = 0; i < WF_SUMSQ_BLOCKS; ++i) ss += st->partial[i];
    st->sumsq = ss;
    const double gn = sqrt(st->sumsq) * grad_scale;
    st->grad_norm = (float)gn;
    float coef = 1.f;
    if (max_norm > 0.f) { coef = (float)(max_norm / (gn + 1e-6)); if (coef > 1.f) coef = 1.f; }
    st->clip_coef = coef * grad_scale;
    const double bc1 = 1.0 - pow((double)b1, (double)st->step), bc2 = 1.0 - pow((double)b2, (double)st->step);
    st->step_size = (float)(lr / bc1);
    st->inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
}

__global__ void adamw_kernel(float* p, const float* g, float* m, float* v, long long n, const AdamState* st,
                             float lr, float b1, float b2, float eps, float wd)
{
    wf_pdl_enter();
    const float coef = st->clip_coef, step_size = st->step_size, ibc2 = st->inv_bc2_sqrt;
    const float decay = 1.f - lr * wd;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float gi = g[i] * coef;
        float pi = p[i] * decay;
        const float mi = m[i] + (gi - m[i]) * (1.f - b1);
        const float vi = v[i] * b2 + gi * gi * (1.f - b2);
        const float denom = sqrtf(vi) * ibc2 + eps;
        pi -= step_size * (mi / denom);
        p[i] = pi; m[i] = mi; v[i] = vi;
    }
}

// ---------------------------------------------------------------------------------------------------------
// layout permutes between the reference's NCHW tensors and the internal [C][P][N] layout (sub-module drop-ins)
//   internal (c, p, b, t)  <->  reference offset  b*r_sb + c*r_sc + p*r_sp + t*r_st
// ---------------------------------------------------------------------------------------------------------
__global__ void permute_kernel(const float* src, float* dst, int C, int P, int B, long long r_sb, long long r_sc, long long r_sp,
                               long long r_st, int to_internal, const float* scale, const float* shift, const float* mean)
{
    wf_pdl_enter();
    const long long total = (long long)C * P * B * WF_T;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int t = (int)(i % WF_T);
        const int b = (int)((i / WF_T) % B);
        const int pp = (int)((i / ((long long)WF_T * B)) % P);
        const int c = (int)(i / ((long long)WF_T * B * P));
        const long long ro = (long long)b * r_sb + (long long)c * r_sc + (long long)pp * r_sp + (long long)t * r_st;
        if (to_internal) dst[i] = src[ro];
        else dst[ro] = scale ? fmaf(scale[c], src[i] - mean[c], shift[c]) : src[i];
    }
}

}  // namespace

// ------------------------------------------- launchers -------------------------------------------
static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

cudaError_t wf_launch_bn_fwd_fin(const BnFwdFin* d, int n, cudaStream_t st)
{
    int cmax = d[0].C;
    if (n > 1 && d[1].C > cmax) cmax = d[1].C;
    dim3 grid(cdiv(cmax, 128), n);
    wf_launch_pdl(bn_finalize_fwd_kernel, dim3(grid), dim3(128), 0, st, d[0], n > 1 ? d[1] : d[0], n);
    return cudaGetLastError();
}
cudaError_t wf_launch_bn_bwd_fin(const BnBwdFin* d, int n, cudaStream_t st)
{
    int cmax = d[0].C;
    if (n > 1 && d[1].C > cmax) cmax = d[1].C;
    dim3 grid(cdiv(cmax, 128), n);
    wf_launch_pdl(bn_finalize_bwd_kernel, dim3(grid), dim3(128), 0, st, d[0], n > 1 ? d[1] : d[0], n);
    return cudaGetLastError();
}
cudaError_t wf_launch_bn_eval_coefs(const BnEvalTable& tab, const float* params, const float* running, float* coefs, cudaStream_t st)
{
    dim3 grid(cdiv(540, 128), tab.n);
    wf_launch_pdl(bn_eval_coefs_kernel, dim3(grid), dim3(128), 0, st, tab, params, running, coefs);
    return cudaGetLastError();
}
static int ew_blocks(long long total4, int C, int num_sms)
{
    long long want = cdiv(total4, 256);
    long long cap = (long long)num_sms * 8 / (C > 0 ? C : 1) + 1;
    if (want > cap) want = cap;
    if (want < 1) want = 1;
    return (int)want;
}
cudaError_t wf_launch_dropout_masks(const MaskTable& tab, unsigned long long seed, unsigned long long* state, int num_sms, cudaStream_t st)
{
    long long nmax = 0;
    for (int i = 0; i < tab.n; ++i) nmax = tab.s[i].numel > nmax ? tab.s[i].numel : nmax;
    long long bx = cdiv((nmax + 3) / 4, 256 * 4);                       // ~4 quads per thread of the largest site
    const long long cap = (long long)num_sms * 8 / tab.n + 1;
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    wf_launch_pdl(dropout_masks_kernel, dim3((unsigned)bx, (unsigned)tab.n), dim3(256), 0, st, tab, seed, state);
    return cudaGetLastError();
}
cudaError_t wf_launch_join_fwd(const JoinP& p, int num_sms, cudaStream_t st)
{
    dim3 grid(ew_blocks(p.plane / 4, p.C, num_sms), p.C);
    wf_launch_pdl(join_fwd_kernel, dim3(grid), dim3(256), 0, st, p);
    return cudaGetLastError();
}
cudaError_t wf_launch_join_bwd(const JoinP& p, int num_sms, cudaStream_t st)
{
    dim3 grid(ew_blocks(p.plane / 4, p.C, num_sms), p.C);
    wf_launch_pdl(join_bwd_kernel<256>, dim3(grid), dim3(256), 0, st, p);
    return cudaGetLastError();
}
cudaError_t wf_launch_bn_bwd_stats(const float* dy, const float* raw, const float* mean, int C, long long plane, double* s0, double* s1, int num_sms, cudaStream_t st)
{
    dim3 grid(ew_blocks(plane / 4, C, num_sms), C);
    wf_launch_pdl(bn_bwd_stats_kernel<256>, dim3(grid), dim3(256), 0, st, dy, raw, mean, plane, s0, s1);
    return cudaGetLastError();
}
cudaError_t wf_launch_pool_fwd(const float* raw, const float* scale, const float* shift, const float* mean, float* pred, int B, cudaStream_t st)
{
    wf_launch_pdl(pool_fwd_kernel, dim3(cdiv(30LL * B, 128)), dim3(128), 0, st, raw, scale, shift, mean, pred, B);
    return cudaGetLastError();
}
cudaError_t wf_launch_pool_bwd(const float* raw, const float* scale, const float* shift, const float* mean, const float* dpred, float* dy, int B,
                               double* s0, double* s1, cudaStream_t st)
{
    dim3 grid(cdiv(15LL * B, 256) > 64 ? 64 : cdiv(15LL * B, 256), 2);
    wf_launch_pdl(pool_bwd_kernel<256>, dim3(grid), dim3(256), 0, st, raw, scale, shift, mean, dpred, dy, B, s0, s1);
    return cudaGetLastError();
}
cudaError_t wf_launch_pose_loss(const float* pred, const float* target, int B, int type, float pw, float bw, const float* gscale,
                                float* dpred, double* acc2, float* out3, cudaStream_t st)
{
    wf_launch_pdl(pose_loss_kernel<128>, dim3(cdiv(B, 128)), dim3(128), 0, st, pred, target, B, type, pw, bw, gscale, dpred, acc2);
    wf_launch_pdl(pose_loss_finish_kernel, dim3(1), dim3(1), 0, st, acc2, B, pw, bw, out3);
    return cudaGetLastError();
}
cudaError_t wf_launch_metrics(const float* pred, const float* target, int B, const MetricThr& thr, int torso, unsigned long long* counts,
                              double* dsum, float* out, cudaStream_t st)
{
    wf_launch_pdl(metrics_kernel<128>, dim3(cdiv(B, 128)), dim3(128), 0, st, pred, target, B, thr, 2, torso ? 12 : 5, counts, dsum);
    wf_launch_pdl(metrics_finish_kernel, dim3(1), dim3(32), 0, st, counts, dsum, B, thr.n, out);
    return cudaGetLastError();
}
cudaError_t wf_launch_pack(const PackTable& tab, const float* params, float* packed, cudaStream_t st)
{
    if (tab.n == 0) return cudaSuccess;
    dim3 grid(64, tab.n);
    wf_launch_pdl(pack_weights_kernel, dim3(grid), dim3(256), 0, st, tab, params, packed);
    return cudaGetLastError();
}
cudaError_t wf_launch_adamw(float* p, const float* g, float* m, float* v, long long n, AdamState* state, float lr, float b1, float b2,
                            float eps, float wd, float max_norm, float grad_scale, int num_sms, cudaStream_t st)
{
    wf_launch_pdl(sumsq_kernel<256>, dim3(WF_SUMSQ_BLOCKS), dim3(256), 0, st, g, n, &state->partial[0]);
    wf_launch_pdl(adamw_prep_kernel, dim3(1), dim3(1), 0, st, state, lr, b1, b2, max_norm, grad_scale);
    wf_launch_pdl(adamw_kernel, dim3(num_sms * 4), dim3(256), 0, st, p, g, m, v, n, state, lr, b1, b2, eps, wd);
    return cudaGetLastError();
}
cudaError_t wf_launch_permute(const float* src, float* dst, int C, int P, int B, long long r_sb, long long r_sc, long long r_sp,
                              long long r_st, int to_internal, int num_sms, cudaStream_t st)
{
    wf_launch_pdl(permute_kernel, dim3(num_sms * 8), dim3(256), 0, st, src, dst, C, P, B, r_sb, r_sc, r_sp, r_st, to_internal, nullptr, nullptr, nullptr);
    return cudaGetLastError();
}
cudaError_t wf_launch_permute_affine(const float* src, float* dst, int C, int P, int B, long long r_sb, long long r_sc, long long r_sp,
                                     long long r_st, const float* scale, const float* shift, const float* mean, int num_sms, cudaStream_t st)
{
    wf_launch_pdl(permute_kernel, dim3(num_sms * 8), dim3(256), 0, st, src, dst, C, P, B, r_sb, r_sc, r_sp, r_st, 0, scale, shift, mean);
    return cudaGetLastError();
}
