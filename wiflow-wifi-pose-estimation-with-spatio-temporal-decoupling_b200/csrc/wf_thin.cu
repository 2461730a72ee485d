// Kernels for the thin-channel layers of the conv stack (ConvBlock1 `up`, residual blocks 0 and 1: 1..16 channels on
// 240/120/60-long rows; models/convnet.py:11-28,48-64).  These layers carry 8% of the FLOPs but most of the activation
// bytes (153.6 KB per window per layer), so a GEMM tiling wastes the machine on them.  Here a CTA stages one
// [channels][positions][columns] window of the (BatchNorm+SiLU+Dropout2d-transformed) input in shared memory exactly once
// and every thread computes all output channels of its 4 columns at one output position (direct convolution, weights
// broadcast from shared memory).  The backward-weights kernel stages the same windows of the input and of the
// BatchNorm-backward-transformed output gradient and contracts them over (position, column) with one warp per
// (4 output channels, tap) task, persistent over the tensor so the cross-lane reduction happens once per CTA.
#include "wf_common.cuh"
#include "wf_elem.h"

namespace {

constexpr int THIN_NT = 64;          // columns per CTA tile (forward / backward-data)
constexpr int THIN_MAXPOS = 18;      // input positions staged per tile

__device__ __forceinline__ int floor_div(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

template <int COUTP>
__global__ void __launch_bounds__(256, COUTP == 8 ? 4 : 3) thin_conv_kernel(const ConvP p, int PC)
{
    wf_pdl_enter();
    constexpr int NT = THIN_NT, Q = NT / 4;
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int n0 = blockIdx.x * NT;
    const int p0 = blockIdx.y * PC;
    float* wsm = smem;                                           // [ntaps][Cin][COUTP]
    float* tile = smem + ((p.ntaps * p.Cin * COUTP + 3) & ~3);   // [Cin][npos][NT]
    __shared__ double red[2][COUTP];

    for (int i = tid; i < p.ntaps * p.Cin * COUTP; i += nthreads) {
        const int co = i % COUTP, ci = (i / COUTP) % p.Cin, tap = i / (COUTP * p.Cin);
        wsm[i] = (co < p.Mpad) ? p.w[((size_t)tap * p.Kpad + ci) * p.Mpad + co] : 0.f;
    }
    if (tid < 2 * COUTP) red[tid / COUTP][tid % COUTP] = 0.0;

    // input window needed by output positions [p0, p0+PC)
    int dmin = p.dp[0], dmax = p.dp[0];
    for (int t = 1; t < p.ntaps; ++t) { dmin = min(dmin, p.dp[t]); dmax = max(dmax, p.dp[t]); }
    const int plast = min(p0 + PC, p.Pout) - 1;
    int lo = floor_div(p0 * p.pmul + dmin, p.pdiv), hi = floor_div(plast * p.pmul + dmax, p.pdiv);
    lo = max(lo, 0); hi = min(hi, p.Pin - 1);
    const int npos = max(hi - lo + 1, 0);

    TileSrc src{p.in, p.in2, p.in_sc, p.in_sp, p.in_sb, p.pro_mode, p.pro_a, p.pro_b, p.pro_c, p.pro_d, p.mask, p.m_sb, p.m_sc, p.m_st, p.Cin, p.Pin};
    wf_stage_tile<NT>(src, tile, p.Cin, lo, npos, n0, p.N, tid, nthreads);
    __syncthreads();

    const int q = tid % Q, pl = tid / Q;
    const int opos = p0 + pl, n = n0 + q * 4;
    const bool valid = (pl < PC) && (opos < p.Pout) && (n < p.N);
    float acc[COUTP][4];
#pragma unroll
    for (int i = 0; i < COUTP; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }
    if (valid) {
        for (int tap = 0; tap < p.ntaps; ++tap) {
            int num = opos * p.pmul + p.dp[tap];
            if (num < 0 || (p.pdiv > 1 && (num % p.pdiv))) continue;
            num /= p.pdiv;
            if (num >= p.Pin) continue;
            const int r = num - lo;
            for (int ci = 0; ci < p.Cin; ++ci) {
                const float4 v = ld4(tile + ((ci * npos + r) * NT + q * 4));
                const float* wr = wsm + (tap * p.Cin + ci) * COUTP;
#pragma unroll
                for (int c4 = 0; c4 < COUTP / 4; ++c4) {
                    const float4 w = ld4(wr + c4 * 4);
                    const float ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        acc[c4 * 4 + j][0] = fmaf(ww[j], v.x, acc[c4 * 4 + j][0]);
                        acc[c4 * 4 + j][1] = fmaf(ww[j], v.y, acc[c4 * 4 + j][1]);
                        acc[c4 * 4 + j][2] = fmaf(ww[j], v.z, acc[c4 * 4 + j][2]);
                        acc[c4 * 4 + j][3] = fmaf(ww[j], v.w, acc[c4 * 4 + j][3]);
                    }
                }
            }
        }
    }
    const bool want_stats = (p.epi_mode != EPI_STORE) && (p.stat0 != nullptr);
#pragma unroll
    for (int co = 0; co < COUTP; ++co) {
        float s0 = 0.f, s1 = 0.f;
        if (co < p.Cout) {                         // uniform across the block
            if (valid) {
                const float bias = p.bias ? p.bias[co] : 0.f;
                float es = 0.f, et = 0.f, em = 0.f;
                if (p.epi_mode == EPI_DSILU) { es = p.e_scale[co]; et = p.e_shift[co]; }
                if (p.epi_mode == EPI_DSILU || p.epi_mode == EPI_DAFF) em = p.e_mean[co];
                float v[4] = {acc[co][0] + bias, acc[co][1] + bias, acc[co][2] + bias, acc[co][3] + bias};
                wf_epilogue_quad(p, co, opos, n, es, et, em, v, s0, s1);
            }
            if (want_stats) {
                const double d0 = warp_sum_d((double)s0), d1 = warp_sum_d((double)s1);
                if ((tid & 31) == 0) { atomicAdd(&red[0][co], d0); atomicAdd(&red[1][co], d1); }
            }
        }
    }
    if (want_stats) {
        __syncthreads();
        if (tid < 2 * COUTP) {
            const int which = tid / COUTP, co = tid % COUTP;
            if (co < p.Cout) atomicAdd((which ? p.stat1 : p.stat0) + co, red[which][co]);
        }
    }
    wf_bn_tail(p.tail);
}

// ---------------------------------------------------------------------------------------------------------
// thin backward-weights
// ---------------------------------------------------------------------------------------------------------
constexpr int TW_NT = 128;           // columns per region: one float4 per lane
constexpr int TW_PC = 4;             // output positions per region

template <int CINP>
__global__ void __launch_bounds__(384, 2) thin_wgrad_kernel(const WgradP p, int coutp, int PS, int nregions)
{
    wf_pdl_enter();
    constexpr int NT = TW_NT, PC = TW_PC;
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int cgroups = coutp / 4;
    const int ntasks = cgroups * p.ntaps;
    const int task = warp % ntasks, ps = warp / ntasks;
    const int cog = task % cgroups, tap = task / cgroups;
    int dmin = p.dp[0], dmax = p.dp[0];
    for (int t = 1; t < p.ntaps; ++t) { dmin = min(dmin, p.dp[t]); dmax = max(dmax, p.dp[t]); }
    const int xpos_max = (PC - 1) * p.pmul + (dmax - dmin) + 1;
    float* gt = smem;                                    // [coutp][PC][NT]
    float* xt = smem + coutp * PC * NT;                  // [CINP][xpos_max][NT]
    const int ntile_n = (p.N + NT - 1) / NT;

    float acc[4][CINP];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < CINP; ++j) acc[i][j] = 0.f;

    TileSrc gs{p.g, p.g2, (long long)p.Pout * p.N, p.N, WF_T, p.g_pro, p.g_a, p.g_b, p.g_c, p.g_d, nullptr, 0, 0, 0, p.Cout, p.Pout};
    TileSrc xs{p.in, p.in2, p.in_sc, p.in_sp, p.in_sb, p.pro_mode, p.pro_a, p.pro_b, p.pro_c, p.pro_d, p.mask, p.m_sb, p.m_sc, p.m_st, p.Cin, p.Pin};

    for (int reg = blockIdx.x; reg < nregions; reg += gridDim.x) {
        const int n0 = (reg % ntile_n) * NT;
        const int p0 = (reg / ntile_n) * PC;
        const int xlo = p0 * p.pmul + dmin;
        __syncthreads();                                 // previous region fully consumed
        wf_stage_tile<NT, 4>(gs, gt, coutp, p0, PC, n0, p.N, tid, nthreads);     // 2-3 items per thread: one round trip instead of three
        wf_stage_tile<NT>(xs, xt, CINP, xlo, xpos_max, n0, p.N, tid, nthreads);
        __syncthreads();
        for (int pl = ps; pl < PC; pl += PS) {
            const int r = pl * p.pmul + p.dp[tap] - dmin;             // row of the staged input window
            float4 g[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) g[i] = ld4(gt + (((cog * 4 + i) * PC + pl) * NT + lane * 4));
#pragma unroll
            for (int j = 0; j < CINP; ++j) {
                const float4 x = ld4(xt + ((j * xpos_max + r) * NT + lane * 4));
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    acc[i][j] = fmaf(g[i].x, x.x, fmaf(g[i].y, x.y, fmaf(g[i].z, x.z, fmaf(g[i].w, x.w, acc[i][j]))));
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < CINP; ++j) {
            const float s = warp_sum(acc[i][j]);
            const int co = cog * 4 + i;
            if (lane == 0 && co < p.Cout && j < p.Cin) atomicAdd(p.dw + ((size_t)co * p.Cin + j) * p.ntaps + tap, s);
        }
}

}  // namespace

// ------------------------------------------- launchers -------------------------------------------
bool wf_thin_conv_ok(const ConvP& p)
{
    if (p.groups != 1 || p.Cin > 16 || p.Cout > 16) return false;
    for (int t = 0; t < p.ntaps; ++t) if (p.dn[t] != 0) return false;
    return true;
}

cudaError_t wf_launch_thin_conv(const ConvP& p, cudaStream_t st)
{
    const int coutp = p.Cout <= 8 ? 8 : 16;
    int dmin = p.dp[0], dmax = p.dp[0];
    for (int t = 1; t < p.ntaps; ++t) { dmin = p.dp[t] < dmin ? p.dp[t] : dmin; dmax = p.dp[t] > dmax ? p.dp[t] : dmax; }
    // largest chunk of output positions whose input window fits THIN_MAXPOS rows and 256 threads
    int PC = 16;
    const int extra = p.pdiv > 1 ? 2 : 1;
    while (PC > 1 && ((PC - 1) * p.pmul + (dmax - dmin)) / p.pdiv + extra > THIN_MAXPOS) PC /= 2;
    const int npos_max = ((PC - 1) * p.pmul + (dmax - dmin)) / p.pdiv + extra;
    const size_t smem = (((size_t)p.ntaps * p.Cin * coutp + 3) & ~(size_t)3) * 4 + (size_t)p.Cin * npos_max * THIN_NT * 4;
    dim3 grid((p.N + THIN_NT - 1) / THIN_NT, (p.Pout + PC - 1) / PC);
    const int threads = (THIN_NT / 4) * PC;
    cudaError_t e;
    if (coutp == 8) {
        static WfSmemOptIn optin;
        if ((e = wf_smem_optin(optin, thin_conv_kernel<8>, 100 * 1024))) return e;
        wf_launch_pdl(thin_conv_kernel<8>, dim3(grid), dim3(threads), smem, st, p, PC);
    } else {
        static WfSmemOptIn optin;
        if ((e = wf_smem_optin(optin, thin_conv_kernel<16>, 100 * 1024))) return e;
        wf_launch_pdl(thin_conv_kernel<16>, dim3(grid), dim3(threads), smem, st, p, PC);
    }
    return cudaGetLastError();
}

bool wf_thin_wgrad_ok(const WgradP& p)
{
    if (p.groups != 1 || p.Cin > 16 || p.Cout > 16) return false;
    for (int t = 0; t < p.ntaps; ++t) if (p.dn[t] != 0) return false;
    return true;
}

template <int CINP>
static cudaError_t launch_thin_wgrad_t(const WgradP& p, int num_sms, cudaStream_t st)
{
    const int coutp = (p.Cout + 3) / 4 * 4;
    const int ntasks = (coutp / 4) * p.ntaps;
    int PS = 1;
    while (PS < TW_PC && ntasks * PS * 2 <= 12) PS *= 2;        // aim for 8..12 warps per CTA
    const int warps = ntasks * PS;
    int dmin = p.dp[0], dmax = p.dp[0];
    for (int t = 1; t < p.ntaps; ++t) { dmin = p.dp[t] < dmin ? p.dp[t] : dmin; dmax = p.dp[t] > dmax ? p.dp[t] : dmax; }
    const int xpos_max = (TW_PC - 1) * p.pmul + (dmax - dmin) + 1;
    const size_t smem = ((size_t)coutp * TW_PC + (size_t)CINP * xpos_max) * TW_NT * 4;
    const int nregions = ((p.N + TW_NT - 1) / TW_NT) * ((p.Pout + TW_PC - 1) / TW_PC);
    int grid = num_sms * 2;
    if (grid > nregions) grid = nregions;
    static WfSmemOptIn optin;
    if (cudaError_t e = wf_smem_optin(optin, thin_wgrad_kernel<CINP>, 100 * 1024)) return e;
    wf_launch_pdl(thin_wgrad_kernel<CINP>, dim3(grid), dim3(warps * 32), smem, st, p, coutp, PS, nregions);
    return cudaGetLastError();
}

cudaError_t wf_launch_thin_wgrad(const WgradP& p, int num_sms, cudaStream_t st)
{
    if (p.Cin <= 4) return launch_thin_wgrad_t<4>(p, num_sms, st);
    if (p.Cin <= 8) return launch_thin_wgrad_t<8>(p, num_sms, st);
    return launch_thin_wgrad_t<16>(p, num_sms, st);
}
