// Shared device helpers and kernel-parameter structs for the WiFlow sm_100a kernels.
//
// Internal activation layout (DESIGN.md "Data layout in HBM"): every intermediate is
//   [channel][position][n]   with n = b*20 + t   (b = CSI window, t = time step),
// n contiguous.  TCN tensors have one position, the conv stack uses position = feature-axis
// index w, the attention/decoder tensors use position = keypoint slot h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define WF_T 20            // time steps per CSI window (models/pose_model.py:72)
#define WF_MAX_TAPS 9
#define TC_KC 32          // K elements per pipeline stage of the tcgen05 kernels (wf_tc.cu)

// prologue modes: how an operand element is produced from what is stored in HBM
enum { PRO_NONE = 0,       // x
       PRO_BNSILU = 1,     // mask * silu(a[c]*(x - d[c]) + b[c])  (BatchNorm + SiLU (+Dropout); d = mean, b = beta)
       PRO_AFFINE = 2,     // a[c]*(x - d[c]) + b[c]                (BatchNorm only)
       PRO_BNBWD = 3 };    // a[c]*dy + b[c]*(raw - d[c]) + c[c]    (BatchNorm backward, two tensors; d = batch mean)
// epilogue modes of the conv GEMM
enum { EPI_STORE = 0,      // out = acc + bias
       EPI_STATS = 1,      // ... and accumulate sum / sum-of-squares per output channel
       EPI_DSILU = 2,      // dy = acc * mask * silu'(s*(raw-mean)+t); stats: sum dy, sum dy*(raw-mean)
       EPI_DAFF = 3 };     // dy = acc;                          stats: sum dy, sum dy*(raw-mean)

struct BnFwdFin {
    int C; double count;
    const double *s0, *s1;
    const float *gamma, *beta;
    float *scale, *shift, *mean, *rstd;
    float *run_mean, *run_var;          // nullptr: do not touch running statistics
    long long* nbt;
};
struct BnBwdFin {
    int C; double count;
    const double *s0, *s1;              // sum dy, sum dy*(raw - mean)
    const float *gamma, *mean, *rstd;
    float *dgamma, *dbeta;              // may be nullptr
    float *alpha, *beta_c, *delta;
    int frozen;                         // eval-mode BatchNorm (running statistics): a fixed per-channel affine, dx = gamma*rstd*dy
    float* conv_dbias;                  // frozen only: gradient of the bias of the conv feeding this BatchNorm (or nullptr)
};
// BatchNorm finalize folded into the kernel whose epilogue completes the per-channel sums (training mode): every CTA of that kernel
// calls wf_bn_tail() once after its last statistics atomic, and the CTA that draws the last ticket does the work of the
// bn_finalize_{fwd,bwd} kernels (coefficients for the consumers, running statistics / gamma-beta gradients) -- one launch and one
// dependent-launch gap less per BatchNorm and direction.
struct BnTail {
    unsigned* counter;                  // nullptr: no tail.  Zero before the launch; reset by the last CTA
    int nf, nb;                         // forward finalizes (0/1), backward finalizes (0..2)
    BnFwdFin f;
    BnBwdFin b[2];
};

struct ConvP {
    // B operand ("input" of the conv)
    const float* in;
    const float* in2;                 // PRO_BNBWD: raw tensor, same addressing
    long long in_sc, in_sp, in_sb;    // element (c,p,b,t) at c*in_sc + p*in_sp + b*in_sb + t
    int pro_mode;
    const float *pro_a, *pro_b, *pro_c, *pro_d;
    const float* mask;                // multiplicative dropout mask or nullptr
    long long m_sb, m_sc; int m_st;   // mask[b*m_sb + c*m_sc + t*m_st]
    // A operand (packed weights [groups][ntaps][Kpad][Mpad], m contiguous, zero padded)
    const float* w;
    int Kpad, Mpad;
    const float* wtc;                 // tcgen05 path (wf_tc.cu): hi/lo-split shared-memory images [M tile][K chunk], or nullptr
    int tc_kt;                        //   number of K chunks (TC_KC channels each)
    // geometry
    int Cin, Cout, groups, Pin, Pout, N, ntaps;
    int pmul, pdiv;                   // ipos = (opos*pmul + dp[tap]) / pdiv, valid iff divisible and in range
    int dp[WF_MAX_TAPS], dn[WF_MAX_TAPS];
    // output
    float* out;
    long long out_sc, out_sp, out_sb;
    const float* bias;
    int epi_mode, accumulate;
    const float* eraw;                // EPI_DSILU/DAFF: raw tensor at the output location (addressing of out)
    const float *e_scale, *e_shift, *e_mean;
    const float* emask; long long em_sb, em_sc; int em_st;
    double *stat0, *stat1;            // per output channel
    BnTail tail;                      // finalize of the BatchNorm the statistics belong to (or counter == nullptr)
};

struct WgradP {
    const float* g;  const float* g2;          // dy / raw of the conv output, [groups*Cout][Pout][N]
    int g_pro;  const float *g_a, *g_b, *g_c, *g_d;  // PRO_NONE or PRO_BNBWD
    const float* in; const float* in2;
    long long in_sc, in_sp, in_sb;
    int pro_mode; const float *pro_a, *pro_b, *pro_c, *pro_d;
    const float* mask; long long m_sb, m_sc; int m_st;
    int Cin, Cout, groups, Pin, Pout, N, ntaps, pmul;
    int dp[WF_MAX_TAPS], dn[WF_MAX_TAPS];
    float* dw;                                  // reference layout [groups*Cout][Cin][ntaps], atomically accumulated
    int kchunks;                                // split of the (p, n) reduction across blockIdx.x
};



// sigmoid through the SFU: 2^(-x*log2 e) and an approximate reciprocal (max relative error ~1e-6, far inside the 1e-4 budget;
// the accurate expf/division sequences cost 4x the instructions in kernels that are issue bound)
__device__ __forceinline__ float wf_sigmoid(float x)
{
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
    return r;
}
__device__ __forceinline__ float wf_silu(float x) { return x * wf_sigmoid(x); }
__device__ __forceinline__ float wf_dsilu(float x) { float s = wf_sigmoid(x); return s * (1.f + x * (1.f - s)); }

// Programmatic dependent launch: every kernel of the model path starts with wf_pdl_enter() -- it lets the next kernel of the stream
// begin launching (its CTAs take SM slots as ours retire and park at their own griddepcontrol.wait) and then waits until everything
// the previous kernels wrote is visible.  Nothing may touch global memory before it.  Launched through wf_launch_pdl() below;
// with the attribute off (WF_PDL=0) both instructions are no-ops.
__device__ __forceinline__ void wf_pdl_enter()
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

// one channel of the forward / backward BatchNorm finalize (models/tcn.py:28 ... nn.BatchNorm1d training semantics: biased variance for
// the normalisation, unbiased for the running estimate, momentum 0.1).  The sums come from other CTAs' atomics: read them at L2.
__device__ __forceinline__ void wf_bn_fwd_fin_channel(const BnFwdFin& d, int c)
{
    const double cnt = d.count;
    const double mean = __ldcg(d.s0 + c) / cnt;
    double var = __ldcg(d.s1 + c) / cnt - mean * mean;
    if (var < 0) var = 0;
    const double rstd = 1.0 / sqrt(var + 1e-5);
    const float gam = d.gamma[c], bet = d.beta[c];
    d.scale[c] = (float)(gam * rstd);
    d.shift[c] = bet;
    d.mean[c] = (float)mean;
    d.rstd[c] = (float)rstd;
    if (d.run_mean) {
        const double unb = cnt > 1 ? var * cnt / (cnt - 1) : var;
        d.run_mean[c] = (float)(0.9 * d.run_mean[c] + 0.1 * mean);
        d.run_var[c] = (float)(0.9 * d.run_var[c] + 0.1 * unb);
        if (c == 0 && d.nbt) *d.nbt += 1;
    }
}
__device__ __forceinline__ void wf_bn_bwd_fin_channel(const BnBwdFin& d, int c)
{
    const double cnt = d.count;
    const double rstd = d.rstd[c], gam = d.gamma[c];
    const double s0 = __ldcg(d.s0 + c);
    const double sx = rstd * __ldcg(d.s1 + c);        // sum dy * xhat   (s1 = sum dy * (raw - mean))
    if (d.dgamma) { d.dgamma[c] = (float)sx; d.dbeta[c] = (float)s0; }
    const double alpha = gam * rstd;
    const double c1 = s0 / cnt, c2 = sx / cnt;
    d.alpha[c] = (float)alpha;
    if (d.frozen) {          // statistics are constants: no mean / variance terms; the conv bias in front has a real gradient
        d.beta_c[c] = 0.f;
        d.delta[c] = 0.f;
        if (d.conv_dbias) d.conv_dbias[c] = (float)(alpha * s0);
        return;
    }
    d.beta_c[c] = (float)(-alpha * c2 * rstd);
    d.delta[c] = (float)(-alpha * c1);
}
// called by ALL threads of EVERY CTA of the kernel, after the CTA's last statistics atomic (block-wide barriers inside)
__device__ __forceinline__ void wf_bn_tail(const BnTail& t)
{
    if (t.counter == nullptr) return;
    __syncthreads();                                   // the CTA's statistics atomics are issued ...
    int last = 0;
    if (threadIdx.x == 0) {
        __threadfence();                               // ... and ordered before the ticket (cumulative over the barrier, as in a grid sync)
        const unsigned total = gridDim.x * gridDim.y * gridDim.z;
        last = atomicAdd(t.counter, 1u) == total - 1 ? 1 : 0;
    }
    last = __syncthreads_or(last);
    if (!last) return;
    __threadfence();
    if (t.nf) for (int c = threadIdx.x; c < t.f.C; c += blockDim.x) wf_bn_fwd_fin_channel(t.f, c);
    for (int i = 0; i < t.nb; ++i)
        for (int c = threadIdx.x; c < t.b[i].C; c += blockDim.x) wf_bn_bwd_fin_channel(t.b[i], c);
    if (threadIdx.x == 0) *t.counter = 0u;
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 f4zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }

// Epilogue of the conv kernels for 4 consecutive columns (n multiple of 4) of output channel co at position opos.
// v[] holds acc + bias on entry; stores the result and adds this quad's contribution to the channel statistics.
__device__ __forceinline__ void wf_epilogue_quad(const ConvP& p, int co, int opos, int n, float es, float et, float em, float v[4], float& s0, float& s1)
{
    const int b = n / WF_T, t = n % WF_T;
    const long long off = (long long)co * p.out_sc + (long long)opos * p.out_sp + (long long)b * p.out_sb + t;
    if (p.accumulate) {
        float4 o = ld4(p.out + off);
        v[0] += o.x; v[1] += o.y; v[2] += o.z; v[3] += o.w;
    }
    if (p.epi_mode == EPI_STATS) {
#pragma unroll
        for (int j = 0; j < 4; ++j) { s0 += v[j]; s1 = fmaf(v[j], v[j], s1); }
    } else if (p.epi_mode == EPI_DSILU || p.epi_mode == EPI_DAFF) {
        const float4 r4 = ld4(p.eraw + off);
        const float r[4] = {r4.x, r4.y, r4.z, r4.w};
        if (p.epi_mode == EPI_DSILU) {
            float mk[4] = {1.f, 1.f, 1.f, 1.f};
            if (p.emask) {
                const float* mp = p.emask + (long long)b * p.em_sb + (long long)co * p.em_sc + (long long)t * p.em_st;
                if (p.em_st == 1) { float4 m4 = ld4(mp); mk[0] = m4.x; mk[1] = m4.y; mk[2] = m4.z; mk[3] = m4.w; }
                else { float mm = *mp; mk[0] = mk[1] = mk[2] = mk[3] = mm; }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = v[j] * mk[j] * wf_dsilu(fmaf(es, r[j] - em, et));
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) { s0 += v[j]; s1 = fmaf(v[j], r[j] - em, s1); }
    }
    st4(p.out + off, make_float4(v[0], v[1], v[2], v[3]));
}

// Operand tile staging for the thin-channel kernels: loads the [C][npos][NT] window (positions pos_lo.., columns n0..) of a
// tensor, applies its prologue and writes it to shared memory; out-of-range positions / columns become zeros.
struct TileSrc {
    const float *p, *p2;
    long long sc, sp, sb;
    int mode; const float *a, *b, *c, *d;
    const float* mask; long long m_sb, m_sc; int m_st;
    int C, P;
};
template <int NT, int U = 1>                                // U = items in flight per thread: their loads are issued before any transform
__device__ __forceinline__ void wf_stage_tile(const TileSrc& s, float* sm, int crows, int pos_lo, int npos, int n0, int N, int tid, int nthreads)
{
    constexpr int Q = NT / 4;
    const int items = crows * npos * Q;
    for (int base = tid; base < items; base += U * nthreads) {
        float4 v[U], w[U];
        long long moff[U];
        int cc[U], dst[U];
        bool ok[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int idx = base + u * nthreads;
            v[u] = f4zero(); w[u] = f4zero(); ok[u] = false; cc[u] = 0; dst[u] = -1; moff[u] = -1;
            if (idx < items) {
                const int q = idx % Q, r = (idx / Q) % npos, c = idx / (Q * npos);
                const int pos = pos_lo + r, n = n0 + q * 4;
                cc[u] = c;
                dst[u] = (c * npos + r) * NT + q * 4;
                if (c < s.C && pos >= 0 && pos < s.P && n < N) {
                    ok[u] = true;
                    const int b = n / WF_T, t = n % WF_T;
                    const long long off = (long long)c * s.sc + (long long)pos * s.sp + (long long)b * s.sb + t;
                    v[u] = ld4(s.p + off);
                    if (s.mode == PRO_BNBWD) w[u] = ld4(s.p2 + off);
                    else if (s.mode == PRO_BNSILU && s.mask) {
                        const float* mp = s.mask + (long long)b * s.m_sb + (long long)c * s.m_sc + (long long)t * s.m_st;
                        if (s.m_st == 1) w[u] = ld4(mp); else { const float mm = *mp; w[u] = make_float4(mm, mm, mm, mm); }
                    } else w[u] = make_float4(1.f, 1.f, 1.f, 1.f);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (dst[u] < 0) continue;
            float4 x = v[u];
            if (ok[u]) {
                const int c = cc[u];
                if (s.mode == PRO_BNSILU) {
                    const float ca = s.a[c], cb = s.b[c], cm = s.d[c];
                    const float4 m = w[u];
                    x.x = wf_silu(fmaf(ca, x.x - cm, cb)) * m.x; x.y = wf_silu(fmaf(ca, x.y - cm, cb)) * m.y;
                    x.z = wf_silu(fmaf(ca, x.z - cm, cb)) * m.z; x.w = wf_silu(fmaf(ca, x.w - cm, cb)) * m.w;
                } else if (s.mode == PRO_AFFINE) {
                    const float ca = s.a[c], cb = s.b[c], cm = s.d[c];
                    x.x = fmaf(ca, x.x - cm, cb); x.y = fmaf(ca, x.y - cm, cb); x.z = fmaf(ca, x.z - cm, cb); x.w = fmaf(ca, x.w - cm, cb);
                } else if (s.mode == PRO_BNBWD) {
                    const float4 r = w[u];
                    const float ca = s.a[c], cb = s.b[c], ccf = s.c[c], cd = s.d[c];
                    x.x = fmaf(ca, x.x, fmaf(cb, r.x - cd, ccf)); x.y = fmaf(ca, x.y, fmaf(cb, r.y - cd, ccf));
                    x.z = fmaf(ca, x.z, fmaf(cb, r.z - cd, ccf)); x.w = fmaf(ca, x.w, fmaf(cb, r.w - cd, ccf));
                }
            }
            st4(sm + dst[u], x);
        }
    }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide sum of two floats, result accumulated into two global doubles (one atomic pair per block)
template <int NT>
__device__ __forceinline__ void block_accum2(float a, float b, double* d0, double* d1) {
    __shared__ double red[2][NT / 32];
    double da = warp_sum_d((double)a), db = warp_sum_d((double)b);
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { red[0][w] = da; red[1][w] = db; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s0 = 0, s1 = 0;
#pragma unroll
        for (int i = 0; i < NT / 32; ++i) { s0 += red[0][i]; s1 += red[1][i]; }
        atomicAdd(d0, s0);
        atomicAdd(d1, s1);
    }
    __syncthreads();
}

// ---- host side ----
#include <atomic>
#include <cstdlib>
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) applies to the CURRENT device only, so the opt-in is remembered per device
// (a process may drive several GPUs: model.to('cuda:1'), tests on another ordinal) and is safe to race on (autograd's backward
// runs on its own thread).  `st` is a function-local static of the call site; it holds the largest size opted in per device.
struct WfSmemOptIn { std::atomic<int> bytes[32]; };
template <class K>
inline cudaError_t wf_smem_optin(WfSmemOptIn& st, K kern, size_t bytes)
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
    dev &= 31;
    if ((int)bytes <= st.bytes[dev].load(std::memory_order_relaxed)) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return e;
    st.bytes[dev].store((int)bytes, std::memory_order_relaxed);
    return cudaSuccess;
}
// multiprocessor count of the current device (cached per device ordinal)
inline int wf_device_sms()
{
    static std::atomic<int> cache[32];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
    int v = cache[dev & 31].load(std::memory_order_relaxed);
    if (v <= 0) {
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        cache[dev & 31].store(v, std::memory_order_relaxed);
    }
    return v;
}

// launch with the programmatic-stream-serialisation attribute (CUDA-graph capturable)
// Set per call by the forward / backward schedules (wf_model.cu): on for small batches, where the step is a chain of short launches
// (B = 64: 2.64 -> 2.53 ms), off for large ones, where early-resident dependents only take slots from the running kernel
// (measured: eval forward at 4096 windows per pass 232.9 k samples/s without, 220.6 k with; B = 1024 training: no difference).
// WF_PDL=0 forces it off, WF_PDL=1 on.
extern thread_local int wf_pdl_mode;
// (Selective PDL just around the 78 BatchNorm-finalize launches of a step was tried at B = 1024: 17.37 ms vs 17.04 ms, no gain.)
inline bool wf_pdl_enabled()
{
    static const int env = [] { const char* e = std::getenv("WF_PDL"); return e ? (e[0] == '0' ? 0 : 1) : -1; }();
    return env >= 0 ? env != 0 : wf_pdl_mode != 0;
}
template <typename... KArgs, typename... Args>
inline cudaError_t wf_launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args)
{
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = wf_pdl_enabled() ? 1 : 0;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}
