// Host side of libwiflow_b200.so: network description, flat parameter / workspace layout, forward and backward
// schedules (one kernel per conv layer, split at every BatchNorm because train-mode BatchNorm is a batch-wide
// reduction -- DESIGN.md "Why split at BatchNorm"), and the extern "C" entry points of include/wiflow_b200.h.
//
// Reference structure being restated (never copied): models/pose_model.py:12-97, models/tcn.py:14-97,
// models/convnet.py:4-74, models/attention.py:7-98.
#include <cuda_runtime.h>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/wiflow_b200.h"
#include "wf_elem.h"

thread_local int wf_pdl_mode = 1;          // see wf_common.cuh: programmatic dependent launch for the launches of this thread
constexpr int WF_PDL_MAX_B = 512;          // windows per call up to which the launch chain, not the kernels, bounds the step

namespace {

thread_local std::string g_err;
int fail(int code, const std::string& msg) { g_err = msg; return code; }

constexpr int T = WF_T;

// opt-in per-launch timing (WF_FLAG_PROFILE): CUDA events around every kernel launch, read back with wf_profile_read
struct ProfRec { std::string name; cudaEvent_t a, b; double flops, bytes; };
thread_local std::vector<ProfRec> g_prof;
std::atomic<long long> g_launches{0};
inline void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
// windows per pass of an eval-mode forward (bounds the workspace); WF_EVAL_CHUNK overrides it for measurements
const int EVAL_CHUNK = [] { const char* e = std::getenv("WF_EVAL_CHUNK"); const int v = e ? std::atoi(e) : 0; return v >= 64 ? v : 4096; }();
// WF_DISABLE_TC=1 routes the pointwise convs through the CUDA-core GEMM instead of tcgen05 (A/B measurements only)
// WF_SERIAL_WGRAD=1 keeps the weight-gradient kernels on the caller's stream (A/B measurements only)
const bool g_overlap_wgrad = [] { const char* e = std::getenv("WF_SERIAL_WGRAD"); return !(e && e[0] == '1'); }();
const bool g_use_tc = [] { const char* e = std::getenv("WF_DISABLE_TC"); return !(e && e[0] == '1'); }();

struct ParamEntry { std::string name; long long off, numel; };

struct BnUnit {
    int C, Cpad;
    long long gamma_off;      // beta at gamma_off + C
    long long run_off;        // running_mean at run_off, running_var at run_off + C
    double count_per_b;       // elements per channel = count_per_b * B
    // workspace
    double *f0, *f1, *b0, *b1;
    unsigned *fctr, *bctr;    // ticket counters of the folded finalizes (BnTail); live in the zeroed statistics regions
    float* coef;              // 8 * Cpad: scale, shift, mean, rstd, alpha, beta, delta, (spare)
    float* scale() const { return coef; }
    float* shift() const { return coef + Cpad; }
    float* mean() const { return coef + 2 * Cpad; }
    float* rstd() const { return coef + 3 * Cpad; }
    float* alpha() const { return coef + 4 * Cpad; }
    float* betac() const { return coef + 5 * Cpad; }
    float* delta() const { return coef + 6 * Cpad; }
};

struct ConvUnit {
    std::string name;
    long long w_off, b_off;   // b_off < 0: no bias
    int cin_g, cout_g, groups, ntaps;
    int pin, pout, stride;
    int dpf[WF_MAX_TAPS], dnf[WF_MAX_TAPS];
    int bn;
    int f_kpad, f_mpad, b_kpad, b_mpad;
    long long fpack, bpack;   // float offsets into the packed-weight region
    bool tc;                  // pointwise conv wide enough for the tcgen05 kernels (wf_tc.cu)
    long long tc_fpack, tc_bpack;
    bool sl;                  // position-tap conv of the conv stack on the TMA + tcgen05 slab kernels (wf_slabtc.cu)
    long long sl_fpack, sl_bpack;
    float *raw, *dy;          // [groups*cout_g][pout][N]
    long long numel_per_n() const { return (long long)groups * cout_g * pout; }
};

struct TcnBlk { std::string name; int g1, pw1, g2, pw2, ds; int cin, cout, dil; float *X, *dX, *dz; int mask0; };
struct CvBlk { std::string name; int c1, c2, c3, ds; int cin, cout, win, wout; float *Y, *dY; int mask0; };
struct AxBlk { std::string name; int qkv; int bn_sim, bn_out; int width; float *sv_raw, *dsv; };
struct DecBlk { int d1, d2; };

struct DebugEntry { std::string name; const float* p; int C, P; };

struct Net {
    std::vector<DebugEntry> dbg;
    std::vector<ParamEntry> params;
    std::vector<BnUnit> bn;
    std::vector<ConvUnit> conv;
    std::vector<TcnBlk> tcn;
    std::vector<CvBlk> cv;
    std::vector<AxBlk> ax;
    bool has_dec = false;
    DecBlk dec{};
    long long nparams = 0, nrunning = 0;
    int nmask = 0;
    // boundary tensors (internal layout) for blocks whose reference layout needs a permute
    int in_C = 0, in_P = 0, out_C = 0, out_P = 0;
    bool in_is_ref_bct = false;   // input is read in place as [B][C][20] (TCN family)
    float *in_buf = nullptr, *din_buf = nullptr, *out_buf = nullptr, *dout_buf = nullptr;
    // workspace regions
    float* packed = nullptr; long long packed_floats = 0;
    float* tcpacked = nullptr; long long tcpacked_floats = 0;
    float* slpacked = nullptr; long long slpacked_floats = 0;
    char* fstats = nullptr; size_t fstats_bytes = 0;
    char* bstats = nullptr; size_t bstats_bytes = 0;
    float* dpred_buf = nullptr;
};

long long add_param(Net& n, const std::string& name, long long numel)
{
    n.params.push_back({name, n.nparams, numel});
    long long o = n.nparams;
    n.nparams += numel;
    return o;
}
int add_bn(Net& n, const std::string& name, int C, double count_per_b)
{
    BnUnit b{};
    b.C = C; b.Cpad = (C + 3) / 4 * 4;
    b.gamma_off = add_param(n, name + ".weight", C);
    add_param(n, name + ".bias", C);
    b.run_off = n.nrunning;
    n.nrunning += 2 * C;
    b.count_per_b = count_per_b;
    n.bn.push_back(b);
    return (int)n.bn.size() - 1;
}
int round_up(int a, int b) { return (a + b - 1) / b * b; }

int add_conv(Net& n, const std::string& name, int cout_total, int cin_g, int groups, int ntaps, bool bias, int pin, int pout, int stride,
             const int* dpf, const int* dnf)
{
    ConvUnit c{};
    c.name = name;
    c.w_off = add_param(n, name + ".weight", (long long)cout_total * cin_g * ntaps);
    c.b_off = bias ? add_param(n, name + ".bias", cout_total) : -1;
    c.cin_g = cin_g; c.cout_g = cout_total / groups; c.groups = groups; c.ntaps = ntaps;
    c.pin = pin; c.pout = pout; c.stride = stride;
    for (int t = 0; t < ntaps; ++t) { c.dpf[t] = dpf ? dpf[t] : 0; c.dnf[t] = dnf ? dnf[t] : 0; }
    c.bn = -1;
    c.f_kpad = round_up(cin_g, wf_conv_bk_for(c.cout_g));
    c.f_mpad = round_up(c.cout_g, wf_conv_bm_for(c.cout_g));
    c.b_kpad = round_up(c.cout_g, wf_conv_bk_for(cin_g));
    c.b_mpad = round_up(cin_g, wf_conv_bm_for(cin_g));
    c.tc = g_use_tc && groups == 1 && ntaps == 1 && stride == 1 && pin == pout && cin_g >= 64 && cout_total >= 64;
    c.sl = !c.tc && pin > 1 && wf_slabtc_shape_ok(cin_g, cout_total, groups, ntaps, c.dnf);
    n.conv.push_back(c);
    return (int)n.conv.size() - 1;
}

// InnerGroupedTemporalBlock (models/tcn.py:14-74); count per channel = 20*B
void add_tcn_block(Net& n, const std::string& pfx, int cin, int cout, int dil)
{
    TcnBlk b{};
    b.name = pfx;
    b.cin = cin; b.cout = cout; b.dil = dil;
    int dn[3] = {-2 * dil, -dil, 0};               // tap k reads t - (2-k)*dil  (causal: left padding + Chomp1d)
    b.g1 = add_conv(n, pfx + "conv1_group", cin, cin / 20, 20, 3, false, 1, 1, 1, nullptr, dn);
    n.conv[b.g1].bn = add_bn(n, pfx + "bn1_group", cin, T);
    b.pw1 = add_conv(n, pfx + "conv1_pw", cout, cin, 1, 1, false, 1, 1, 1, nullptr, nullptr);
    n.conv[b.pw1].bn = add_bn(n, pfx + "bn1_pw", cout, T);
    b.g2 = add_conv(n, pfx + "conv2_group", cout, cout / 20, 20, 3, false, 1, 1, 1, nullptr, dn);
    n.conv[b.g2].bn = add_bn(n, pfx + "bn2_group", cout, T);
    b.pw2 = add_conv(n, pfx + "conv2_pw", cout, cout, 1, 1, false, 1, 1, 1, nullptr, nullptr);
    n.conv[b.pw2].bn = add_bn(n, pfx + "bn2_pw", cout, T);
    b.ds = -1;
    if (cin != cout) {
        b.ds = add_conv(n, pfx + "downsample.0", cout, cin, 1, 1, false, 1, 1, 1, nullptr, nullptr);
        n.conv[b.ds].bn = add_bn(n, pfx + "downsample.1", cout, T);
    }
    b.mask0 = n.nmask;
    n.nmask += 2;
    n.tcn.push_back(b);
}

// ConvBlock1 / AsymmetricConvBlock (models/convnet.py:4-74); count per channel = 20*B*Wout
void add_conv_block(Net& n, const std::string& pfx, int cin, int cout, int win, int stride)
{
    CvBlk b{};
    b.name = pfx;
    b.cin = cin; b.cout = cout; b.win = win; b.wout = (win + 2 - 3) / stride + 1;
    int dp3[3] = {-1, 0, 1};
    const double cnt = (double)T * b.wout;
    b.c1 = add_conv(n, pfx + "block.0", cout, cin, 1, 3, true, win, b.wout, stride, dp3, nullptr);
    n.conv[b.c1].bn = add_bn(n, pfx + "block.1", cout, cnt);
    b.c2 = add_conv(n, pfx + "block.4", cout, cout, 1, 3, true, b.wout, b.wout, 1, dp3, nullptr);
    n.conv[b.c2].bn = add_bn(n, pfx + "block.5", cout, cnt);
    b.c3 = add_conv(n, pfx + "block.8", cout, cout, 1, 3, true, b.wout, b.wout, 1, dp3, nullptr);
    n.conv[b.c3].bn = add_bn(n, pfx + "block.9", cout, cnt);
    b.ds = add_conv(n, pfx + "downsample.0", cout, cin, 1, 1, false, win, b.wout, stride, nullptr, nullptr);
    n.conv[b.ds].bn = add_bn(n, pfx + "downsample.1", cout, cnt);
    b.mask0 = n.nmask;
    n.nmask += 2;
    n.cv.push_back(b);
}

// AxialAttention (models/attention.py:7-80), 64 planes, 8 groups, on the 15x20 grid
void add_axial(Net& n, const std::string& pfx, int width)
{
    AxBlk a{};
    a.name = pfx;
    a.width = width;
    const int L = width ? 20 : 15;
    a.qkv = add_conv(n, pfx + "qkv_transform", 192, 64, 1, 1, false, 15, 15, 1, nullptr, nullptr);
    n.conv[a.qkv].bn = add_bn(n, pfx + "bn_qkv", 192, 15.0 * T);
    a.bn_sim = add_bn(n, pfx + "bn_similarity", 8, 15.0 * T * L);
    a.bn_out = add_bn(n, pfx + "bn_output", 64, 15.0 * T);
    n.ax.push_back(a);
}

// decoder (models/pose_model.py:44-51)
void add_decoder(Net& n, const std::string& pfx)
{
    int dp9[9], dn9[9];
    for (int dh = 0; dh < 3; ++dh)
        for (int dw = 0; dw < 3; ++dw) { dp9[dh * 3 + dw] = dh - 1; dn9[dh * 3 + dw] = dw - 1; }
    n.dec.d1 = add_conv(n, pfx + "0", 32, 64, 1, 9, true, 15, 15, 1, dp9, dn9);
    n.conv[n.dec.d1].bn = add_bn(n, pfx + "1", 32, 15.0 * T);
    n.dec.d2 = add_conv(n, pfx + "3", 2, 32, 1, 1, true, 15, 15, 1, nullptr, nullptr);
    n.conv[n.dec.d2].bn = add_bn(n, pfx + "4", 2, 15.0 * T);
    n.has_dec = true;
}

int build_net(const wf_block_desc* d, Net& n)
{
    static const int tc[5] = {540, 540, 440, 340, 240};
    static const int rc[5] = {8, 8, 16, 32, 64};
    switch (d->block) {
        case WF_BLOCK_MODEL:
            for (int i = 0; i < 4; ++i) add_tcn_block(n, "tcn.network." + std::to_string(i) + ".", tc[i], tc[i + 1], 1 << i);
            add_conv_block(n, "up.", 1, 8, 240, 1);
            for (int i = 0; i < 4; ++i) add_conv_block(n, "residual_blocks." + std::to_string(i) + ".", rc[i], rc[i + 1], 240 >> i, 2);
            add_axial(n, "attention.width_axis.", 1);
            add_axial(n, "attention.height_axis.", 0);
            add_decoder(n, "decoder.");
            n.in_is_ref_bct = true;
            break;
        case WF_BLOCK_TCN:
            for (int i = 0; i < 4; ++i) add_tcn_block(n, "network." + std::to_string(i) + ".", tc[i], tc[i + 1], 1 << i);
            n.in_is_ref_bct = true;
            n.out_C = 240; n.out_P = 1;
            break;
        case WF_BLOCK_INNER_TCN:
            if (d->cin <= 0 || d->cout <= 0 || d->cin % 20 || d->cout % 20 || d->dilation < 1 || d->dilation > 16)
                return fail(WF_E_ARG, "InnerGroupedTemporalBlock: channels must be multiples of 20 (groups=20, tcn.py:18), 1 <= dilation <= 16");
            add_tcn_block(n, "", d->cin, d->cout, d->dilation);
            n.in_is_ref_bct = true;
            n.out_C = d->cout; n.out_P = 1;
            break;
        case WF_BLOCK_CONVBLOCK1:
        case WF_BLOCK_ASYMCONV: {
            const int stride = d->block == WF_BLOCK_ASYMCONV ? 2 : 1;
            if (d->cin <= 0 || d->cout <= 0 || d->width <= 0) return fail(WF_E_ARG, "conv block: cin, cout, width must be positive");
            add_conv_block(n, "", d->cin, d->cout, d->width, stride);
            n.in_C = d->cin; n.in_P = d->width;
            n.out_C = d->cout; n.out_P = n.cv[0].wout;
            break;
        }
        case WF_BLOCK_AXIAL_W:
        case WF_BLOCK_AXIAL_H:
            add_axial(n, "", d->block == WF_BLOCK_AXIAL_W);
            n.in_C = 64; n.in_P = 15; n.out_C = 64; n.out_P = 15;
            break;
        case WF_BLOCK_DUAL_AXIAL:
            add_axial(n, "width_axis.", 1);
            add_axial(n, "height_axis.", 0);
            n.in_C = 64; n.in_P = 15; n.out_C = 64; n.out_P = 15;
            break;
        default:
            return fail(WF_E_ARG, "unknown block id");
    }
    if ((int)n.conv.size() > WF_MAX_CONV || (int)n.bn.size() > WF_MAX_BN) return fail(WF_E_UNSUPPORTED, "too many layers");
    return 0;
}

// ------------------------------- workspace layout -------------------------------
struct Bump {
    char* base; size_t off = 0;
    explicit Bump(char* b) : base(b) {}
    template <class Tp> Tp* take(size_t count)
    {
        off = (off + 255) & ~(size_t)255;
        Tp* p = base ? reinterpret_cast<Tp*>(base + off) : nullptr;
        off += count * sizeof(Tp);
        return p;
    }
};

size_t layout(Net& n, int B, int flags, char* base)
{
    const bool save = (flags & WF_FLAG_SAVE_FOR_BACKWARD) != 0;
    const bool train = (flags & WF_FLAG_TRAIN) != 0;
    const int Bw = train ? B : (B < EVAL_CHUNK ? B : EVAL_CHUNK);
    const long long N = (long long)Bw * T;
    Bump bp(base);
    // packed weights
    long long pf = 0;
    for (auto& c : n.conv) {
        c.fpack = pf; pf += (long long)c.groups * c.ntaps * c.f_kpad * c.f_mpad;
        c.bpack = pf; pf += (long long)c.groups * c.ntaps * c.b_kpad * c.b_mpad;
    }
    n.packed_floats = pf;
    n.packed = bp.take<float>(pf);
    long long tf = 0;
    for (auto& c : n.conv) {
        if (!c.tc) continue;
        c.tc_fpack = tf; tf += wf_tc_pack_floats(c.cout_g, c.cin_g);
        c.tc_bpack = tf; tf += wf_tc_pack_floats(c.cin_g, c.cout_g);
    }
    n.tcpacked_floats = tf;
    n.tcpacked = bp.take<float>(tf);
    long long sf = 0;
    for (auto& c : n.conv) {
        if (!c.sl) continue;
        c.sl_fpack = sf; sf += wf_slabtc_pack_floats(c.cout_g, c.cin_g, c.ntaps, false);
        c.sl_bpack = sf; sf += wf_slabtc_pack_floats(c.cout_g, c.cin_g, c.ntaps, true);
    }
    n.slpacked_floats = sf;
    n.slpacked = bp.take<float>(sf);
    // BN statistics (fp64) and coefficients
    size_t s0 = (bp.off + 255) & ~(size_t)255;
    for (auto& b : n.bn) { b.f0 = bp.take<double>(b.C); b.f1 = bp.take<double>(b.C); b.fctr = reinterpret_cast<unsigned*>(bp.take<double>(1)); }
    n.fstats = base ? base + s0 : nullptr; n.fstats_bytes = bp.off - s0;
    size_t s1 = (bp.off + 255) & ~(size_t)255;
    for (auto& b : n.bn) { b.b0 = bp.take<double>(b.C); b.b1 = bp.take<double>(b.C); b.bctr = reinterpret_cast<unsigned*>(bp.take<double>(1)); }
    n.bstats = base ? base + s1 : nullptr; n.bstats_bytes = bp.off - s1;
    for (auto& b : n.bn) b.coef = bp.take<float>(8 * b.Cpad);
    // activations
    for (auto& c : n.conv) {
        c.raw = bp.take<float>(c.numel_per_n() * N);
        c.dy = save ? bp.take<float>(c.numel_per_n() * N) : nullptr;
    }
    for (auto& t : n.tcn) {
        t.X = bp.take<float>((long long)t.cout * N);
        t.dX = save ? bp.take<float>((long long)t.cout * N) : nullptr;
        t.dz = save ? bp.take<float>((long long)t.cout * N) : nullptr;
    }
    for (auto& v : n.cv) {
        v.Y = bp.take<float>((long long)v.cout * v.wout * N);
        v.dY = save ? bp.take<float>((long long)v.cout * v.wout * N) : nullptr;
    }
    for (auto& a : n.ax) {
        a.sv_raw = bp.take<float>(64LL * 15 * N);
        a.dsv = save ? bp.take<float>(64LL * 15 * N) : nullptr;
    }
    if (n.in_C) {
        n.in_buf = bp.take<float>((long long)n.in_C * n.in_P * N);
        n.din_buf = save ? bp.take<float>((long long)n.in_C * n.in_P * N) : nullptr;
    } else if (n.in_is_ref_bct && save && !n.tcn.empty()) {
        n.din_buf = bp.take<float>((long long)n.tcn[0].cin * N);
    }
    if (n.out_C) {
        n.dout_buf = save ? bp.take<float>((long long)n.out_C * n.out_P * N) : nullptr;
    }
    n.dpred_buf = nullptr;
    n.dbg.clear();
    for (auto& c : n.conv) {
        n.dbg.push_back({c.name + ".raw", c.raw, c.groups * c.cout_g, c.pout});
        if (c.dy) n.dbg.push_back({c.name + ".dy", c.dy, c.groups * c.cout_g, c.pout});
    }
    for (auto& t : n.tcn) {
        n.dbg.push_back({t.name + "out", t.X, t.cout, 1});
        if (t.dX) n.dbg.push_back({t.name + "dout", t.dX, t.cout, 1});
    }
    for (auto& v : n.cv) {
        n.dbg.push_back({v.name + "out", v.Y, v.cout, v.wout});
        if (v.dY) n.dbg.push_back({v.name + "dout", v.dY, v.cout, v.wout});
    }
    for (auto& a : n.ax) {
        n.dbg.push_back({a.name + "sv.raw", a.sv_raw, 64, 15});
        if (a.dsv) n.dbg.push_back({a.name + "sv.dy", a.dsv, 64, 15});
    }
    for (size_t i = 0; i < n.bn.size(); ++i)
        n.dbg.push_back({"bn" + std::to_string(i) + ".coef", n.bn[i].coef, 8, n.bn[i].Cpad});
    return (bp.off + 255) & ~(size_t)255;
}

// ------------------------------- execution context -------------------------------
struct Act { const float* p; long long sc, sp, sb; };     // element (c,pos,b,t) at p + c*sc + pos*sp + b*sb + t
Act internal(const float* p, int P, long long N) { return Act{p, (long long)P * N, N, T}; }

struct Mask { const float* p; long long sb, sc; int st; };
Mask no_mask() { return Mask{nullptr, 0, 0, 0}; }
Mask tcn_mask(const float* p, int C) { return Mask{p, (long long)C * T, T, 1}; }        // [B][C][20]
Mask plane_mask(const float* p, int C) { return Mask{p, C, 1, 0}; }                      // [B][C]

struct Ctx {
    Net& n;
    const float* params; float* grads; float* running; long long* nbt; const float* const* masks;
    int B; long long N; bool train; bool save; cudaStream_t st; int sms;
    bool profile = false;
    cudaError_t err = cudaSuccess;
    cudaStream_t side = nullptr;      // backward only: the weight-gradient kernels run here, concurrently with backward-data on `st`
    cudaEvent_t fork = nullptr;
    std::vector<unsigned char> ftail, btail;      // per BatchNorm: its finalize rides on the kernel that completes the sums (BnTail)
    void ck(cudaError_t e) { if (err == cudaSuccess && e != cudaSuccess) err = e; }
    const float* mask_ptr(int i) const { return (masks && train) ? masks[i] : nullptr; }
};

// One side stream + two events per host thread, created on first use (before any CUDA-graph capture: TrainStep warms up
// eagerly).  A weight gradient depends only on finished tensors (dy, BatchNorm-backward coefficients, forward activations)
// and nothing downstream in the step reads it before the optimizer, so it overlaps the backward-data chain; both families are
// latency bound at low occupancy, which is why sharing the SMs pays.
struct SideStream { cudaStream_t s = nullptr; cudaEvent_t fork = nullptr, join = nullptr; int dev = -1; };
thread_local SideStream g_side;
bool side_stream(SideStream& out)
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return false;
    if (g_side.s == nullptr || g_side.dev != dev) {
        SideStream n{};
        if (cudaStreamCreateWithFlags(&n.s, cudaStreamNonBlocking) != cudaSuccess) return false;
        if (cudaEventCreateWithFlags(&n.fork, cudaEventDisableTiming) != cudaSuccess) return false;
        if (cudaEventCreateWithFlags(&n.join, cudaEventDisableTiming) != cudaSuccess) return false;
        n.dev = dev;
        g_side = n;
    }
    out = g_side;
    return true;
}

struct Scope {          // one kernel launch (or launch pair): counts it and, when profiling, brackets it with events
    Ctx& c; int idx = -1;
    Scope(Ctx& c_, const std::string& name, double flops = 0.0, double bytes = 0.0) : c(c_)
    {
        g_launches.fetch_add(1, std::memory_order_relaxed);
        if (!c.profile) return;
        ProfRec r{name, nullptr, nullptr, flops, bytes};
        cudaEventCreate(&r.a); cudaEventCreate(&r.b);
        cudaEventRecord(r.a, c.st);
        g_prof.push_back(r);
        idx = (int)g_prof.size() - 1;
    }
    ~Scope() { if (idx >= 0) cudaEventRecord(g_prof[idx].b, c.st); }
};

// dense multiply-add count of one pass over conv unit u (forward, backward-data and backward-weights all do this many)
double conv_flops(const ConvUnit& u, long long N) { return 2.0 * u.groups * u.cout_g * u.cin_g * u.ntaps * u.pout * (double)N; }

// algorithmic HBM bytes of one pass over conv unit u: every operand tensor once (kind 0 forward: input + output; 1 backward-data:
// dy + raw of the output side, the input-side raw for the SiLU', the input gradient; 2 backward-weights: dy + raw + input)
double conv_bytes(const ConvUnit& u, long long N, int kind)
{
    const double in = 4.0 * u.groups * u.cin_g * u.pin * (double)N, out = 4.0 * u.groups * u.cout_g * u.pout * (double)N;
    return kind == 0 ? in + out : kind == 1 ? 2 * out + 2 * in : 2 * out + in;
}

struct Pro { int mode; int bn; Mask mask; };
Pro pro_none() { return Pro{PRO_NONE, -1, no_mask()}; }
Pro pro_act(int bn, Mask m = no_mask()) { return Pro{PRO_BNSILU, bn, m}; }
Pro pro_aff(int bn) { return Pro{PRO_AFFINE, bn, no_mask()}; }

void tail_fwd(Ctx& c, BnTail& t, int bi);
void tail_bwd(Ctx& c, BnTail& t, int bn_a, int bn_b);

void fwd_conv(Ctx& c, int ui, Act in, Pro pro)
{
    const ConvUnit& u = c.n.conv[ui];
    const BnUnit& bo = c.n.bn[u.bn];
    ConvP p{};
    p.in = in.p; p.in_sc = in.sc; p.in_sp = in.sp; p.in_sb = in.sb;
    p.pro_mode = pro.mode;
    if (pro.bn >= 0) { p.pro_a = c.n.bn[pro.bn].scale(); p.pro_b = c.n.bn[pro.bn].shift(); p.pro_d = c.n.bn[pro.bn].mean(); }
    p.mask = pro.mask.p; p.m_sb = pro.mask.sb; p.m_sc = pro.mask.sc; p.m_st = pro.mask.st;
    p.w = c.n.packed + u.fpack; p.Kpad = u.f_kpad; p.Mpad = u.f_mpad;
    p.Cin = u.cin_g; p.Cout = u.cout_g; p.groups = u.groups; p.Pin = u.pin; p.Pout = u.pout; p.N = (int)c.N; p.ntaps = u.ntaps;
    p.pmul = u.stride; p.pdiv = 1;
    for (int t = 0; t < u.ntaps; ++t) { p.dp[t] = u.dpf[t]; p.dn[t] = u.dnf[t]; }
    p.out = u.raw; p.out_sc = (long long)u.pout * c.N; p.out_sp = c.N; p.out_sb = T;
    p.bias = u.b_off >= 0 ? c.params + u.b_off : nullptr;
    p.epi_mode = c.train ? EPI_STATS : EPI_STORE;
    p.stat0 = bo.f0; p.stat1 = bo.f1;
    tail_fwd(c, p.tail, u.bn);
    if (u.sl) p.wtc = c.n.slpacked + u.sl_fpack;
    Scope sc(c, std::string(u.tc ? "tc_fwd " : wf_slabtc_conv_ok(p) ? "slab_fwd " : wf_slide_conv_ok(p) ? (wf_slide_conv_is_thin(p) ? "slidethin_fwd " : "slide_fwd ") : wf_thin_conv_ok(p) ? "thin_fwd " : wf_group_conv_ok(p) ? "group_fwd " : "conv_fwd ") + u.name, conv_flops(u, c.N), conv_bytes(u, c.N, 0));
    if (u.tc) { p.wtc = c.n.tcpacked + u.tc_fpack; p.tc_kt = (u.cin_g + TC_KC - 1) / TC_KC; c.ck(wf_launch_tc_conv(p, c.sms, c.st)); }
    else c.ck(wf_launch_conv(p, c.st));
}

// WF_BN_TAIL=1 folds each BatchNorm finalize into the kernel that completes its sums (BnTail, wf_common.cuh).  Off by default:
// measured on a B200 it is SLOWER than the separate one-block launches -- B = 1024: 16.70 vs 16.45 ms / step (175 vs 242 launches),
// B = 64: 2.68 vs 2.58 ms.  With programmatic dependent launch the finalize kernel is already resident when its producer
// drains, so a launch costs ~2 us, while the tail adds a ticket round trip to EVERY CTA of the producer and runs the fp64
// finalize on the critical path of its last CTA.
bool bn_tail_enabled() { static const bool v = [] { const char* e = std::getenv("WF_BN_TAIL"); return e && e[0] == '1'; }(); return v; }

BnFwdFin make_fwd_fin(Ctx& c, int bi)
{
    const BnUnit& b = c.n.bn[bi];
    BnFwdFin f{};
    f.C = b.C; f.count = b.count_per_b * c.B;
    f.s0 = b.f0; f.s1 = b.f1;
    f.gamma = c.params + b.gamma_off; f.beta = c.params + b.gamma_off + b.C;
    f.scale = b.scale(); f.shift = b.shift(); f.mean = b.mean(); f.rstd = b.rstd();
    f.run_mean = c.running ? c.running + b.run_off : nullptr;
    f.run_var = c.running ? c.running + b.run_off + b.C : nullptr;
    f.nbt = c.nbt ? c.nbt + bi : nullptr;
    return f;
}

BnBwdFin make_bwd_fin(Ctx& c, int bi)
{
    const BnUnit& b = c.n.bn[bi];
    BnBwdFin f{};
    f.C = b.C; f.count = b.count_per_b * c.B;
    f.s0 = b.b0; f.s1 = b.b1;
    f.gamma = c.params + b.gamma_off; f.mean = b.mean(); f.rstd = b.rstd();
    f.dgamma = c.grads + b.gamma_off; f.dbeta = c.grads + b.gamma_off + b.C;
    f.alpha = b.alpha(); f.beta_c = b.betac(); f.delta = b.delta();
    f.frozen = c.train ? 0 : 1;
    if (f.frozen)
        for (const ConvUnit& u : c.n.conv)
            if (u.bn == bi && u.b_off >= 0) f.conv_dbias = c.grads + u.b_off;
    return f;
}

// the kernel about to be launched completes the forward sums of BatchNorm bi: let its last CTA finalize (fwd_fin then skips bi)
void tail_fwd(Ctx& c, BnTail& t, int bi)
{
    if (!c.train || !bn_tail_enabled()) return;
    t.counter = c.n.bn[bi].fctr; t.nf = 1; t.f = make_fwd_fin(c, bi);
    if (c.ftail.size() < c.n.bn.size()) c.ftail.resize(c.n.bn.size(), 0);
    c.ftail[bi] = 1;
}
// same for the backward sums of one or two BatchNorms (bwd_fin then skips them)
void tail_bwd(Ctx& c, BnTail& t, int bn_a, int bn_b)
{
    if (!bn_tail_enabled()) return;
    if (c.btail.size() < c.n.bn.size()) c.btail.resize(c.n.bn.size(), 0);
    t.counter = c.n.bn[bn_a].bctr;
    for (int bi : {bn_a, bn_b}) {
        if (bi < 0) continue;
        t.b[t.nb++] = make_bwd_fin(c, bi);
        c.btail[bi] = 1;
    }
}

void fwd_fin(Ctx& c, int bn_a, int bn_b = -1)
{
    if (!c.train) return;
    BnFwdFin d[2];
    int k = 0;
    for (int bi : {bn_a, bn_b}) {
        if (bi < 0) continue;
        if ((size_t)bi < c.ftail.size() && c.ftail[bi]) { c.ftail[bi] = 0; continue; }
        d[k++] = make_fwd_fin(c, bi);
    }
    if (k == 0) return;
    Scope sc(c, "bn_fin_fwd");
    c.ck(wf_launch_bn_fwd_fin(d, k, c.st));
}

void bwd_fin(Ctx& c, int bn_a, int bn_b = -1)
{
    BnBwdFin d[2];
    int k = 0;
    for (int bi : {bn_a, bn_b}) {
        if (bi < 0) continue;
        if ((size_t)bi < c.btail.size() && c.btail[bi]) { c.btail[bi] = 0; continue; }
        d[k++] = make_bwd_fin(c, bi);
    }
    if (k == 0) return;
    Scope sc(c, "bn_fin_bwd");
    c.ck(wf_launch_bn_bwd_fin(d, k, c.st));
}

// backward-data of conv unit ui: consumes (dy, raw) of the unit through its BatchNorm backward, produces the gradient
// w.r.t. the unit's input.  epi: EPI_STORE (input is a materialised tensor), EPI_DSILU / EPI_DAFF (input is the
// BatchNorm(+SiLU) of conv unit `src`'s raw output; also accumulates that BatchNorm's backward sums).
void dgrad_conv(Ctx& c, int ui, float* out, int epi, int src_bn, const float* src_raw, Mask emask, bool accumulate, bool fin_tail = false)
{
    const ConvUnit& u = c.n.conv[ui];
    const BnUnit& bo = c.n.bn[u.bn];
    ConvP p{};
    p.in = u.dy; p.in2 = u.raw;
    p.in_sc = (long long)u.pout * c.N; p.in_sp = c.N; p.in_sb = T;
    p.pro_mode = PRO_BNBWD; p.pro_a = bo.alpha(); p.pro_b = bo.betac(); p.pro_c = bo.delta(); p.pro_d = bo.mean();
    p.w = c.n.packed + u.bpack; p.Kpad = u.b_kpad; p.Mpad = u.b_mpad;
    p.Cin = u.cout_g; p.Cout = u.cin_g; p.groups = u.groups; p.Pin = u.pout; p.Pout = u.pin; p.N = (int)c.N; p.ntaps = u.ntaps;
    p.pmul = 1; p.pdiv = u.stride;
    for (int t = 0; t < u.ntaps; ++t) { p.dp[t] = -u.dpf[t]; p.dn[t] = -u.dnf[t]; }
    p.out = out; p.out_sc = (long long)u.pin * c.N; p.out_sp = c.N; p.out_sb = T;
    p.epi_mode = epi; p.accumulate = accumulate ? 1 : 0;
    if (epi == EPI_DSILU || epi == EPI_DAFF) {
        const BnUnit& bs = c.n.bn[src_bn];
        p.eraw = src_raw; p.e_scale = bs.scale(); p.e_shift = bs.shift(); p.e_mean = bs.mean();
        p.emask = emask.p; p.em_sb = emask.sb; p.em_sc = emask.sc; p.em_st = emask.st;
        p.stat0 = bs.b0; p.stat1 = bs.b1;
        if (fin_tail) tail_bwd(c, p.tail, src_bn, -1);        // this launch alone completes src_bn's backward sums
    }
    if (u.sl) p.wtc = c.n.slpacked + u.sl_bpack;
    Scope sc(c, std::string(u.tc ? "tc_dgrad " : wf_slabtc_conv_ok(p) ? "slab_dgrad " : wf_slide_conv_ok(p) ? (wf_slide_conv_is_thin(p) ? "slidethin_dgrad " : "slide_dgrad ") : wf_thin_conv_ok(p) ? "thin_dgrad " : wf_group_conv_ok(p) ? "group_dgrad " : "conv_dgrad ") + u.name, conv_flops(u, c.N), conv_bytes(u, c.N, 1));
    if (u.tc) { p.wtc = c.n.tcpacked + u.tc_bpack; p.tc_kt = (u.cout_g + TC_KC - 1) / TC_KC; c.ck(wf_launch_tc_conv(p, c.sms, c.st)); }
    else c.ck(wf_launch_conv(p, c.st));
}

void wgrad_conv(Ctx& c, int ui, Act in, Pro pro)
{
    const ConvUnit& u = c.n.conv[ui];
    const BnUnit& bo = c.n.bn[u.bn];
    WgradP p{};
    p.g = u.dy; p.g2 = u.raw; p.g_pro = PRO_BNBWD; p.g_a = bo.alpha(); p.g_b = bo.betac(); p.g_c = bo.delta(); p.g_d = bo.mean();
    p.in = in.p; p.in_sc = in.sc; p.in_sp = in.sp; p.in_sb = in.sb;
    p.pro_mode = pro.mode;
    if (pro.bn >= 0) { p.pro_a = c.n.bn[pro.bn].scale(); p.pro_b = c.n.bn[pro.bn].shift(); p.pro_d = c.n.bn[pro.bn].mean(); }
    p.mask = pro.mask.p; p.m_sb = pro.mask.sb; p.m_sc = pro.mask.sc; p.m_st = pro.mask.st;
    p.Cin = u.cin_g; p.Cout = u.cout_g; p.groups = u.groups; p.Pin = u.pin; p.Pout = u.pout; p.N = (int)c.N; p.ntaps = u.ntaps;
    p.pmul = u.stride;
    for (int t = 0; t < u.ntaps; ++t) { p.dp[t] = u.dpf[t]; p.dn[t] = u.dnf[t]; }
    p.dw = c.grads + u.w_off;
    Scope sc(c, std::string(u.tc ? "tc_wgrad " : wf_slabtc_wgrad_ok(p) ? "slab_wgrad " : wf_slide_wgrad_ok(p) ? "slide_wgrad " : wf_thin_wgrad_ok(p) ? "thin_wgrad " : wf_group_wgrad_ok(p) ? "group_wgrad " : "conv_wgrad ") + u.name, conv_flops(u, c.N), conv_bytes(u, c.N, 2));
    cudaStream_t ws = c.st;
    if (c.side) {                     // fork: everything this kernel reads has been enqueued on the main stream by now
        c.ck(cudaEventRecord(c.fork, c.st));
        c.ck(cudaStreamWaitEvent(c.side, c.fork, 0));
        ws = c.side;
    }
    if (u.tc) c.ck(wf_launch_tc_wgrad(p, c.sms, ws));
    else c.ck(wf_launch_wgrad(p, c.sms, ws));
}

// ------------------------------- forward schedules -------------------------------
void tcn_block_fwd(Ctx& c, TcnBlk& b, Act xin)
{
    Net& n = c.n;
    fwd_conv(c, b.g1, xin, pro_none());
    if (b.ds >= 0) fwd_conv(c, b.ds, xin, pro_none());
    fwd_fin(c, n.conv[b.g1].bn, b.ds >= 0 ? n.conv[b.ds].bn : -1);
    fwd_conv(c, b.pw1, internal(n.conv[b.g1].raw, 1, c.N), pro_act(n.conv[b.g1].bn));
    fwd_fin(c, n.conv[b.pw1].bn);
    fwd_conv(c, b.g2, internal(n.conv[b.pw1].raw, 1, c.N), pro_act(n.conv[b.pw1].bn, tcn_mask(c.mask_ptr(b.mask0), b.cout)));
    fwd_fin(c, n.conv[b.g2].bn);
    fwd_conv(c, b.pw2, internal(n.conv[b.g2].raw, 1, c.N), pro_act(n.conv[b.g2].bn));
    fwd_fin(c, n.conv[b.pw2].bn);
    JoinP j{};
    const BnUnit& ba = n.bn[n.conv[b.pw2].bn];
    j.a = n.conv[b.pw2].raw; j.out = b.X; j.plane = c.N; j.N = (int)c.N; j.C = b.cout;
    j.a_mode = PRO_BNSILU; j.a_scale = ba.scale(); j.a_shift = ba.shift(); j.a_mean = ba.mean();
    Mask m = tcn_mask(c.mask_ptr(b.mask0 + 1), b.cout);
    j.mask = m.p; j.m_sb = m.sb; j.m_sc = m.sc; j.m_st = m.st;
    if (b.ds >= 0) {
        const BnUnit& br = n.bn[n.conv[b.ds].bn];
        j.r = n.conv[b.ds].raw; j.r_mode = PRO_AFFINE; j.r_scale = br.scale(); j.r_shift = br.shift(); j.r_mean = br.mean();
        j.r_sc = c.N; j.r_sp = 0; j.r_sb = T;
    } else {
        j.r = xin.p; j.r_mode = PRO_NONE; j.r_sc = xin.sc; j.r_sp = 0; j.r_sb = xin.sb;
    }
    { Scope sc(c, "join_fwd " + b.name); c.ck(wf_launch_join_fwd(j, c.sms, c.st)); }
}

void conv_block_fwd(Ctx& c, CvBlk& b, Act xin)
{
    Net& n = c.n;
    fwd_conv(c, b.c1, xin, pro_none());
    fwd_conv(c, b.ds, xin, pro_none());
    fwd_fin(c, n.conv[b.c1].bn, n.conv[b.ds].bn);
    fwd_conv(c, b.c2, internal(n.conv[b.c1].raw, b.wout, c.N), pro_act(n.conv[b.c1].bn, plane_mask(c.mask_ptr(b.mask0), b.cout)));
    fwd_fin(c, n.conv[b.c2].bn);
    fwd_conv(c, b.c3, internal(n.conv[b.c2].raw, b.wout, c.N), pro_act(n.conv[b.c2].bn, plane_mask(c.mask_ptr(b.mask0 + 1), b.cout)));
    fwd_fin(c, n.conv[b.c3].bn);
    JoinP j{};
    const BnUnit& ba = n.bn[n.conv[b.c3].bn];
    const BnUnit& br = n.bn[n.conv[b.ds].bn];
    j.a = n.conv[b.c3].raw; j.out = b.Y; j.plane = (long long)b.wout * c.N; j.N = (int)c.N; j.C = b.cout;
    j.a_mode = PRO_AFFINE; j.a_scale = ba.scale(); j.a_shift = ba.shift(); j.a_mean = ba.mean();
    j.r = n.conv[b.ds].raw; j.r_mode = PRO_AFFINE; j.r_scale = br.scale(); j.r_shift = br.shift(); j.r_mean = br.mean();
    j.r_sc = j.plane; j.r_sp = c.N; j.r_sb = T;
    Scope sc(c, "join_fwd " + b.name);
    c.ck(wf_launch_join_fwd(j, c.sms, c.st));
}

AttnP attn_params(Ctx& c, AxBlk& a)
{
    Net& n = c.n;
    const ConvUnit& q = n.conv[a.qkv];
    const BnUnit &bq = n.bn[q.bn], &bs = n.bn[a.bn_sim], &bo = n.bn[a.bn_out];
    AttnP p{};
    p.width = a.width; p.B = c.B; p.N = (int)c.N;
    p.qkv_raw = q.raw; p.qkv_scale = bq.scale(); p.qkv_shift = bq.shift(); p.qkv_mean = bq.mean();
    p.sim_scale = bs.scale(); p.sim_shift = bs.shift(); p.sim_mean = bs.mean();
    p.sim_s0 = bs.f0; p.sim_s1 = bs.f1;
    p.sv_raw = a.sv_raw;
    p.sv_s0 = c.train ? bo.f0 : nullptr; p.sv_s1 = c.train ? bo.f1 : nullptr;
    p.dsv = a.dsv; p.sv_alpha = bo.alpha(); p.sv_beta = bo.betac(); p.sv_delta = bo.delta(); p.sv_mean = bo.mean();
    p.sim_alpha = bs.alpha(); p.sim_beta = bs.betac(); p.sim_delta = bs.delta();
    p.dsim_s0 = bs.b0; p.dsim_s1 = bs.b1;
    p.dqkv = q.dy;
    return p;
}

void axial_fwd(Ctx& c, AxBlk& a, Act xin, Pro pro)
{
    Net& n = c.n;
    fwd_conv(c, a.qkv, xin, pro);
    fwd_fin(c, n.conv[a.qkv].bn);
    AttnP p = attn_params(c, a);
    if (c.train) {
        { Scope sc(c, "attn_fwd_stats " + a.name); c.ck(wf_launch_attn_fwd_stats(p, c.st)); }
        fwd_fin(c, a.bn_sim);
    }
    { Scope sc(c, "attn_fwd " + a.name); c.ck(wf_launch_attn_fwd(p, c.st)); }
    fwd_fin(c, a.bn_out);
}

void decoder_fwd(Ctx& c, Act xin, Pro pro, float* pred)
{
    Net& n = c.n;
    fwd_conv(c, n.dec.d1, xin, pro);
    fwd_fin(c, n.conv[n.dec.d1].bn);
    fwd_conv(c, n.dec.d2, internal(n.conv[n.dec.d1].raw, 15, c.N), pro_act(n.conv[n.dec.d1].bn));
    fwd_fin(c, n.conv[n.dec.d2].bn);
    const BnUnit& b = n.bn[n.conv[n.dec.d2].bn];
    Scope sc(c, "pool_fwd");
    c.ck(wf_launch_pool_fwd(n.conv[n.dec.d2].raw, b.scale(), b.shift(), b.mean(), pred, c.B, c.st));
}

// ------------------------------- backward schedules -------------------------------
// gradient w.r.t. the block input goes to dxin (may be nullptr when nobody needs it)
void decoder_bwd(Ctx& c, Act xin, Pro pro, const float* dpred, float* dxin_dy, int src_bn, const float* src_raw)
{
    Net& n = c.n;
    ConvUnit &d1 = n.conv[n.dec.d1], &d2 = n.conv[n.dec.d2];
    const BnUnit& b2 = n.bn[d2.bn];
    { Scope sc(c, "pool_bwd"); c.ck(wf_launch_pool_bwd(d2.raw, b2.scale(), b2.shift(), b2.mean(), dpred, d2.dy, c.B, b2.b0, b2.b1, c.st)); }
    bwd_fin(c, d2.bn);
    wgrad_conv(c, n.dec.d2, internal(d1.raw, 15, c.N), pro_act(d1.bn));
    dgrad_conv(c, n.dec.d2, d1.dy, EPI_DSILU, d1.bn, d1.raw, no_mask(), false, true);
    bwd_fin(c, d1.bn);
    wgrad_conv(c, n.dec.d1, xin, pro);
    dgrad_conv(c, n.dec.d1, dxin_dy, EPI_DAFF, src_bn, src_raw, no_mask(), false);
}

// on entry a.dsv holds dy of bn_output and its backward sums are complete
void axial_bwd(Ctx& c, AxBlk& a, Act xin, Pro pro, float* dxin, int epi, int src_bn, const float* src_raw)
{
    Net& n = c.n;
    ConvUnit& q = n.conv[a.qkv];
    bwd_fin(c, a.bn_out);
    AttnP p = attn_params(c, a);
    { Scope sc(c, "attn_bwd_stats " + a.name); c.ck(wf_launch_attn_bwd_stats(p, c.st)); }
    bwd_fin(c, a.bn_sim);
    { Scope sc(c, "attn_bwd " + a.name); c.ck(wf_launch_attn_bwd(p, c.st)); }
    const BnUnit& bq = n.bn[q.bn];
    { Scope sc(c, "bn_bwd_stats " + a.name); c.ck(wf_launch_bn_bwd_stats(q.dy, q.raw, bq.mean(), 192, 15LL * c.N, bq.b0, bq.b1, c.sms, c.st)); }
    bwd_fin(c, q.bn);
    wgrad_conv(c, a.qkv, xin, pro);
    if (dxin) dgrad_conv(c, a.qkv, dxin, epi, src_bn, src_raw, no_mask(), false);
}

void conv_block_bwd(Ctx& c, CvBlk& b, Act xin, float* dxin)
{
    Net& n = c.n;
    ConvUnit &c1 = n.conv[b.c1], &c2 = n.conv[b.c2], &c3 = n.conv[b.c3], &ds = n.conv[b.ds];
    JoinP j{};
    const BnUnit& ba = n.bn[c3.bn];
    const BnUnit& br = n.bn[ds.bn];
    j.a = c3.raw; j.plane = (long long)b.wout * c.N; j.N = (int)c.N; j.C = b.cout;
    j.a_mode = PRO_AFFINE; j.a_scale = ba.scale(); j.a_shift = ba.shift();
    j.r = ds.raw; j.r_mode = PRO_AFFINE; j.r_scale = br.scale(); j.r_shift = br.shift();
    j.r_sc = j.plane; j.r_sp = c.N; j.r_sb = T;
    j.a_mean = ba.mean(); j.r_mean = br.mean();
    j.dout = b.dY; j.dz = c3.dy; j.da = nullptr;           // c3.dy doubles as dy of the shortcut BatchNorm
    j.a_stat0 = ba.b0; j.a_stat1 = ba.b1; j.r_stat0 = br.b0; j.r_stat1 = br.b1;
    tail_bwd(c, j.tail, c3.bn, ds.bn);
    { Scope sc(c, "join_bwd " + b.name); c.ck(wf_launch_join_bwd(j, c.sms, c.st)); }
    bwd_fin(c, c3.bn, ds.bn);
    Mask m0 = plane_mask(c.mask_ptr(b.mask0), b.cout), m1 = plane_mask(c.mask_ptr(b.mask0 + 1), b.cout);
    wgrad_conv(c, b.c3, internal(c2.raw, b.wout, c.N), pro_act(c2.bn, m1));
    dgrad_conv(c, b.c3, c2.dy, EPI_DSILU, c2.bn, c2.raw, m1, false, true);
    bwd_fin(c, c2.bn);
    wgrad_conv(c, b.c2, internal(c1.raw, b.wout, c.N), pro_act(c1.bn, m0));
    dgrad_conv(c, b.c2, c1.dy, EPI_DSILU, c1.bn, c1.raw, m0, false, true);
    bwd_fin(c, c1.bn);
    wgrad_conv(c, b.c1, xin, pro_none());
    // the shortcut conv shares dz with c3: temporarily view ds through c3's dy
    float* saved = ds.dy; ds.dy = c3.dy;
    wgrad_conv(c, b.ds, xin, pro_none());
    if (dxin) {
        dgrad_conv(c, b.ds, dxin, EPI_STORE, -1, nullptr, no_mask(), false);
        dgrad_conv(c, b.c1, dxin, EPI_STORE, -1, nullptr, no_mask(), true);
    }
    ds.dy = saved;
}

void tcn_block_bwd(Ctx& c, TcnBlk& b, Act xin, float* dxin)
{
    Net& n = c.n;
    ConvUnit &g1 = n.conv[b.g1], &pw1 = n.conv[b.pw1], &g2 = n.conv[b.g2], &pw2 = n.conv[b.pw2];
    JoinP j{};
    const BnUnit& ba = n.bn[pw2.bn];
    j.a = pw2.raw; j.plane = c.N; j.N = (int)c.N; j.C = b.cout;
    j.a_mode = PRO_BNSILU; j.a_scale = ba.scale(); j.a_shift = ba.shift();
    Mask m1 = tcn_mask(c.mask_ptr(b.mask0 + 1), b.cout), m0 = tcn_mask(c.mask_ptr(b.mask0), b.cout);
    j.mask = m1.p; j.m_sb = m1.sb; j.m_sc = m1.sc; j.m_st = m1.st;
    j.dout = b.dX; j.da = pw2.dy; j.a_mean = ba.mean(); j.r_mean = nullptr;
    j.a_stat0 = ba.b0; j.a_stat1 = ba.b1;
    if (b.ds >= 0) {
        ConvUnit& ds = n.conv[b.ds];
        const BnUnit& br = n.bn[ds.bn];
        j.r = ds.raw; j.r_mode = PRO_AFFINE; j.r_scale = br.scale(); j.r_shift = br.shift();
        j.r_sc = c.N; j.r_sp = 0; j.r_sb = T;
        j.dz = ds.dy; j.r_stat0 = br.b0; j.r_stat1 = br.b1; j.r_mean = br.mean();
    } else {
        j.r = xin.p; j.r_mode = PRO_NONE; j.r_sc = xin.sc; j.r_sp = 0; j.r_sb = xin.sb;
        j.dz = dxin ? dxin : b.dz;          // identity shortcut: dz IS the gradient reaching the block input
    }
    tail_bwd(c, j.tail, pw2.bn, b.ds >= 0 ? n.conv[b.ds].bn : -1);
    { Scope sc(c, "join_bwd " + b.name); c.ck(wf_launch_join_bwd(j, c.sms, c.st)); }
    bwd_fin(c, pw2.bn, b.ds >= 0 ? n.conv[b.ds].bn : -1);
    wgrad_conv(c, b.pw2, internal(g2.raw, 1, c.N), pro_act(g2.bn));
    dgrad_conv(c, b.pw2, g2.dy, EPI_DSILU, g2.bn, g2.raw, no_mask(), false, true);
    bwd_fin(c, g2.bn);
    wgrad_conv(c, b.g2, internal(pw1.raw, 1, c.N), pro_act(pw1.bn, m0));
    dgrad_conv(c, b.g2, pw1.dy, EPI_DSILU, pw1.bn, pw1.raw, m0, false, true);
    bwd_fin(c, pw1.bn);
    wgrad_conv(c, b.pw1, internal(g1.raw, 1, c.N), pro_act(g1.bn));
    dgrad_conv(c, b.pw1, g1.dy, EPI_DSILU, g1.bn, g1.raw, no_mask(), false, true);
    bwd_fin(c, g1.bn);
    wgrad_conv(c, b.g1, xin, pro_none());
    if (b.ds >= 0) {
        wgrad_conv(c, b.ds, xin, pro_none());
        if (dxin) {
            dgrad_conv(c, b.ds, dxin, EPI_STORE, -1, nullptr, no_mask(), false);
            dgrad_conv(c, b.g1, dxin, EPI_STORE, -1, nullptr, no_mask(), true);
        }
    } else if (dxin) {
        dgrad_conv(c, b.g1, dxin, EPI_STORE, -1, nullptr, no_mask(), true);
    }
}

// ------------------------------- shared prologue -------------------------------
void prepare_weights(Ctx& c)
{
    Net& n = c.n;
    Scope sc(c, "pack_weights");
    c.ck(cudaMemsetAsync(n.packed, 0, n.packed_floats * sizeof(float), c.st));
    PackTable tab{};
    for (size_t i = 0; i < n.conv.size(); ++i) {
        const ConvUnit& u = n.conv[i];
        if (u.tc) continue;                         // tcgen05 layers read their own split images (tc_pack below): 86 % of all weights
        PackEntry& e = tab.e[tab.n++];
        e.param_off = (int)u.w_off; e.cout = u.cout_g * u.groups; e.cin = u.cin_g; e.groups = u.groups; e.ntaps = u.ntaps;
        e.f_kpad = u.f_kpad; e.f_mpad = u.f_mpad; e.b_kpad = u.b_kpad; e.b_mpad = u.b_mpad;
        e.fwd_off = u.fpack; e.bwd_off = u.bpack;
    }
    c.ck(wf_launch_pack(tab, c.params, n.packed, c.st));
    TcPackTable tt{};
    for (const ConvUnit& u : n.conv)
        if (u.tc) tt.e[tt.n++] = TcPackEntry{(int)u.w_off, u.cout_g, u.cin_g, u.tc_fpack, u.tc_bpack};
    c.ck(wf_launch_tc_pack(tt, c.params, n.tcpacked, c.st));
    SlabPackTable sp{};
    for (const ConvUnit& u : n.conv)
        if (u.sl) sp.e[sp.n++] = SlabPackEntry{(int)u.w_off, u.cout_g, u.cin_g, u.ntaps, u.sl_fpack, u.sl_bpack};
    c.ck(wf_launch_slabtc_pack(sp, c.params, n.slpacked, c.st));
}

void eval_coefs(Ctx& c)
{
    Net& n = c.n;
    BnEvalTable tab{};
    tab.n = (int)n.bn.size();
    for (int i = 0; i < tab.n; ++i) {
        const BnUnit& b = n.bn[i];
        tab.e[i] = BnEvalEntry{b.C, b.Cpad, (int)b.gamma_off, (int)b.run_off, (int)(b.coef - n.bn[0].coef)};
    }
    Scope sc(c, "bn_eval_coefs");
    c.ck(wf_launch_bn_eval_coefs(tab, c.params, c.running, n.bn[0].coef, c.st));
}

int num_sms() { return wf_device_sms(); }

int check_device()
{
    static std::atomic<unsigned> ok_mask{0};           // bit = device ordinal already checked
    int dev = 0, major = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return fail((int)e, std::string("no CUDA device: ") + cudaGetErrorString(e));
    if (ok_mask.load(std::memory_order_relaxed) & (1u << (dev & 31))) return 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (major != 10) return fail(WF_E_ARCH, "libwiflow_b200 is compiled for sm_100a only; device compute capability major = " + std::to_string(major));
    ok_mask.fetch_or(1u << (dev & 31), std::memory_order_relaxed);
    return 0;
}

// reference layout strides of the block's input / output tensors for the permute kernel: (r_sb, r_sc, r_sp, r_st)
struct RefStrides { long long sb, sc, sp, st; };
RefStrides ref_strides(const wf_block_desc* d, bool output, const Net& n)
{
    const int C = output ? n.out_C : n.in_C, P = output ? n.out_P : n.in_P;
    switch (d->block) {
        case WF_BLOCK_CONVBLOCK1:
        case WF_BLOCK_ASYMCONV:            // [B, C, 20, W]: (b,c,t,w)
            return RefStrides{(long long)C * T * P, (long long)T * P, 1, P};
        case WF_BLOCK_AXIAL_W:
        case WF_BLOCK_AXIAL_H:
        case WF_BLOCK_DUAL_AXIAL:          // [B, 64, 15, 20]: (b,c,h,t)
            return RefStrides{(long long)C * P * T, (long long)P * T, T, 1};
        default:                           // [B, C, 20]
            return RefStrides{(long long)C * T, T, 0, 1};
    }
}

int run_forward(const wf_block_desc* d, const float* x, const float* params, float* running, long long* nbt, const float* const* masks,
                float* y, void* ws, size_t ws_bytes, int B, int flags, cudaStream_t st)
{
    if (int e = check_device()) return e;
    wf_pdl_mode = B <= WF_PDL_MAX_B ? 1 : 0;
    if (!d || !x || !params || !y || !ws || B <= 0) return fail(WF_E_ARG, "null pointer or non-positive batch");
    const bool train = (flags & WF_FLAG_TRAIN) != 0;
    if (!train && !running) return fail(WF_E_ARG, "eval mode needs the running statistics");
    if (!train && (flags & WF_FLAG_SAVE_FOR_BACKWARD) && B > EVAL_CHUNK)
        return fail(WF_E_UNSUPPORTED, "eval-mode forward with saved activations takes at most " + std::to_string(EVAL_CHUNK) + " windows per call");
    Net n;
    if (int e = build_net(d, n)) return e;
    const size_t need = layout(n, B, flags, (char*)ws);
    if (need > ws_bytes) return fail(WF_E_WORKSPACE, "workspace too small: need " + std::to_string(need) + " bytes");
    if (((uintptr_t)ws & 255) || ((uintptr_t)x & 15) || ((uintptr_t)y & 15)) return fail(WF_E_ARG, "workspace must be 256-byte, x/y 16-byte aligned");

    const int chunk = train ? B : (B < EVAL_CHUNK ? B : EVAL_CHUNK);
    Ctx c{n, params, nullptr, running, nbt, masks, chunk, (long long)chunk * T, train, (flags & WF_FLAG_SAVE_FOR_BACKWARD) != 0, st, num_sms()};
    c.profile = (flags & WF_FLAG_PROFILE) != 0;
    prepare_weights(c);
    if (train) c.ck(cudaMemsetAsync(n.fstats, 0, n.fstats_bytes, st));
    else eval_coefs(c);

    const long long in_per_b = n.in_is_ref_bct ? (long long)n.tcn[0].cin * T : (long long)n.in_C * n.in_P * T;
    long long out_per_b = d->block == WF_BLOCK_MODEL ? 30 : (long long)n.out_C * n.out_P * T;
    for (int b0 = 0; b0 < B; b0 += chunk) {
        const int bc = (B - b0 < chunk) ? B - b0 : chunk;
        c.B = bc; c.N = (long long)bc * T;
        if (bc != chunk) layout(n, bc, flags | WF_FLAG_TRAIN, (char*)ws);     // tail chunk: re-lay out for the smaller batch (fits: bc < chunk)
        const float* xc = x + (long long)b0 * in_per_b;
        float* yc = y + (long long)b0 * out_per_b;
        Act cur{};
        Pro cur_pro = pro_none();
        if (n.in_is_ref_bct) {
            cur = Act{xc, T, 0, (long long)n.tcn[0].cin * T};
        } else {
            RefStrides rs = ref_strides(d, false, n);
            { count_launch(); c.ck(wf_launch_permute(xc, n.in_buf, n.in_C, n.in_P, bc, rs.sb, rs.sc, rs.sp, rs.st, 1, c.sms, st)); }
            cur = internal(n.in_buf, n.in_P, c.N);
        }
        for (auto& t : n.tcn) { tcn_block_fwd(c, t, cur); cur = internal(t.X, 1, c.N); }
        if (!n.cv.empty() && !n.tcn.empty()) cur = internal(n.tcn.back().X, 240, c.N);   // [240][N] viewed as [1][240][N]
        for (auto& v : n.cv) { conv_block_fwd(c, v, cur); cur = internal(v.Y, v.wout, c.N); }
        for (auto& a : n.ax) {
            axial_fwd(c, a, cur, cur_pro);
            cur = internal(a.sv_raw, 15, c.N);
            cur_pro = pro_aff(a.bn_out);
        }
        if (n.has_dec) {
            decoder_fwd(c, cur, cur_pro, yc);
        } else {
            // write the block output in the reference layout
            RefStrides rs = ref_strides(d, true, n);
            const float* src = cur.p;
            if (!n.ax.empty()) {
                // output = bn_output(sv): apply the affine while permuting (JoinP-free path: use a conv-free affine permute)
                const BnUnit& bo = n.bn[n.ax.back().bn_out];
                { count_launch(); c.ck(wf_launch_permute_affine(src, yc, n.out_C, n.out_P, bc, rs.sb, rs.sc, rs.sp, rs.st, bo.scale(), bo.shift(), bo.mean(), c.sms, st)); }
            } else {
                { count_launch(); c.ck(wf_launch_permute(src, yc, n.out_C, n.out_P, bc, rs.sb, rs.sc, rs.sp, rs.st, 0, c.sms, st)); }
            }
        }
    }
    if (c.err != cudaSuccess) return fail((int)c.err, std::string("CUDA error in forward: ") + cudaGetErrorString(c.err));
    return 0;
}

int run_backward(const wf_block_desc* d, const float* x, const float* params, const float* const* masks, const float* dy, float* grads,
                 float* dx, void* ws, size_t ws_bytes, int B, int flags, cudaStream_t st)
{
    if (int e = check_device()) return e;
    wf_pdl_mode = B <= WF_PDL_MAX_B ? 1 : 0;
    if (!d || !x || !params || !dy || !grads || !ws || B <= 0) return fail(WF_E_ARG, "null pointer or non-positive batch");
    if (!(flags & WF_FLAG_SAVE_FOR_BACKWARD))
        return fail(WF_E_UNSUPPORTED, "backward needs a forward run with WF_FLAG_SAVE_FOR_BACKWARD");
    const bool train = (flags & WF_FLAG_TRAIN) != 0;          // eval mode: BatchNorm is a fixed affine of the running statistics, dropout is off
    Net n;
    if (int e = build_net(d, n)) return e;
    const size_t need = layout(n, B, flags, (char*)ws);
    if (need > ws_bytes) return fail(WF_E_WORKSPACE, "workspace too small: need " + std::to_string(need) + " bytes");
    Ctx c{n, params, grads, nullptr, nullptr, masks, B, (long long)B * T, train, true, st, num_sms()};
    c.profile = (flags & WF_FLAG_PROFILE) != 0;
    SideStream ss{};
    if (!c.profile && g_overlap_wgrad && side_stream(ss)) { c.side = ss.s; c.fork = ss.fork; }     // profiling times every kernel alone
    c.ck(cudaMemsetAsync(n.bstats, 0, n.bstats_bytes, st));
    c.ck(cudaMemsetAsync(grads, 0, n.nparams * sizeof(float), st));

    // block inputs, as in the forward
    std::vector<Act> tin(n.tcn.size()), vin(n.cv.size());
    Act cur{};
    if (n.in_is_ref_bct) cur = Act{x, T, 0, (long long)n.tcn[0].cin * T};
    else cur = internal(n.in_buf, n.in_P, c.N);
    for (size_t i = 0; i < n.tcn.size(); ++i) { tin[i] = cur; cur = internal(n.tcn[i].X, 1, c.N); }
    if (!n.cv.empty() && !n.tcn.empty()) cur = internal(n.tcn.back().X, 240, c.N);
    for (size_t i = 0; i < n.cv.size(); ++i) { vin[i] = cur; cur = internal(n.cv[i].Y, n.cv[i].wout, c.N); }
    const Act ax_in0 = cur;

    // gradient entering the last block
    float* g_in = nullptr;                       // gradient w.r.t. the current block's output (materialised tensors)
    if (n.has_dec) {
        AxBlk& ah = n.ax.back();
        decoder_bwd(c, internal(ah.sv_raw, 15, c.N), pro_aff(ah.bn_out), dy, ah.dsv, ah.bn_out, ah.sv_raw);
    } else {
        RefStrides rs = ref_strides(d, true, n);
        if (!n.ax.empty()) {
            // dy is the gradient of bn_output's output: permute into dsv and accumulate its BatchNorm-backward sums
            AxBlk& ah = n.ax.back();
            { count_launch(); c.ck(wf_launch_permute(dy, ah.dsv, n.out_C, n.out_P, B, rs.sb, rs.sc, rs.sp, rs.st, 1, c.sms, st)); }
            const BnUnit& bo = n.bn[ah.bn_out];
            c.ck(wf_launch_bn_bwd_stats(ah.dsv, ah.sv_raw, bo.mean(), 64, 15LL * c.N, bo.b0, bo.b1, c.sms, st));
        } else {
            { count_launch(); c.ck(wf_launch_permute(dy, n.dout_buf, n.out_C, n.out_P, B, rs.sb, rs.sc, rs.sp, rs.st, 1, c.sms, st)); }
            g_in = n.dout_buf;
        }
    }
    const bool want_dx = dx != nullptr;
    // attention blocks, last to first
    for (int i = (int)n.ax.size() - 1; i >= 0; --i) {
        AxBlk& a = n.ax[i];
        if (i > 0) {
            AxBlk& prev = n.ax[i - 1];
            axial_bwd(c, a, internal(prev.sv_raw, 15, c.N), pro_aff(prev.bn_out), prev.dsv, EPI_DAFF, prev.bn_out, prev.sv_raw);
        } else {
            float* dst = !n.cv.empty() ? n.cv.back().dY : (want_dx ? n.din_buf : nullptr);
            axial_bwd(c, a, ax_in0, pro_none(), dst, EPI_STORE, -1, nullptr);
        }
    }
    if (n.ax.empty() && !n.cv.empty() && g_in) {
        c.ck(cudaMemcpyAsync(n.cv.back().dY, g_in, sizeof(float) * n.cv.back().cout * n.cv.back().wout * c.N, cudaMemcpyDeviceToDevice, st));
    }
    for (int i = (int)n.cv.size() - 1; i >= 0; --i) {
        float* dst = i > 0 ? n.cv[i - 1].dY : (!n.tcn.empty() ? n.tcn.back().dX : (want_dx ? n.din_buf : nullptr));
        conv_block_bwd(c, n.cv[i], vin[i], dst);
    }
    if (n.cv.empty() && n.ax.empty() && !n.tcn.empty() && g_in) {
        c.ck(cudaMemcpyAsync(n.tcn.back().dX, g_in, sizeof(float) * n.tcn.back().cout * c.N, cudaMemcpyDeviceToDevice, st));
    }
    for (int i = (int)n.tcn.size() - 1; i >= 0; --i) {
        float* dst = i > 0 ? n.tcn[i - 1].dX : (want_dx ? n.din_buf : nullptr);
        tcn_block_bwd(c, n.tcn[i], tin[i], dst);
    }
    if (want_dx) {
        if (n.in_is_ref_bct) {
            const int C = n.tcn[0].cin;
            { count_launch(); c.ck(wf_launch_permute(n.din_buf, dx, C, 1, B, (long long)C * T, T, 0, 1, 0, c.sms, st)); }
        } else {
            RefStrides rs = ref_strides(d, false, n);
            { count_launch(); c.ck(wf_launch_permute(n.din_buf, dx, n.in_C, n.in_P, B, rs.sb, rs.sc, rs.sp, rs.st, 0, c.sms, st)); }
        }
    }
    if (c.side) {                     // join: the caller's stream owns the complete gradient again
        c.ck(cudaEventRecord(ss.join, c.side));
        c.ck(cudaStreamWaitEvent(st, ss.join, 0));
    }
    if (c.err != cudaSuccess) return fail((int)c.err, std::string("CUDA error in backward: ") + cudaGetErrorString(c.err));
    return 0;
}

}  // namespace

// =========================================== C ABI ===========================================
extern "C" {

const char* wf_last_error_string(void) { return g_err.c_str(); }
int wf_version(void) { return 100; }

long long wf_param_count(const wf_block_desc* d) { Net n; return build_net(d, n) ? -1 : n.nparams; }
long long wf_running_count(const wf_block_desc* d) { Net n; return build_net(d, n) ? -1 : n.nrunning; }
int wf_bn_count(const wf_block_desc* d) { Net n; return build_net(d, n) ? -1 : (int)n.bn.size(); }
int wf_dropout_sites(const wf_block_desc* d) { Net n; return build_net(d, n) ? -1 : n.nmask; }

int wf_param_table(const wf_block_desc* d, int i, char* name, int name_cap, long long* offset, long long* numel)
{
    Net n;
    if (int e = build_net(d, n)) return e;
    if (i < 0 || i >= (int)n.params.size()) return WF_E_ARG;
    if (name && name_cap > 0) { std::strncpy(name, n.params[i].name.c_str(), name_cap - 1); name[name_cap - 1] = 0; }
    if (offset) *offset = n.params[i].off;
    if (numel) *numel = n.params[i].numel;
    return 0;
}

size_t wf_workspace_bytes(const wf_block_desc* d, int B, int flags)
{
    Net n;
    if (build_net(d, n) || B <= 0) return 0;
    return layout(n, B, flags, nullptr);
}

// Debug/test introspection: i-th named workspace tensor ([C][P][B*20] fp32) of the block's layout.
WF_API int wf_debug_tensor(const wf_block_desc* d, int B, int flags, int i, char* name, int name_cap, long long* byte_offset, int* C, int* P)
{
    Net n;
    if (int e = build_net(d, n)) return e;
    static char dummy[1];
    layout(n, B, flags, dummy);
    if (i < 0 || i >= (int)n.dbg.size()) return WF_E_ARG;
    if (name && name_cap > 0) { std::strncpy(name, n.dbg[i].name.c_str(), name_cap - 1); name[name_cap - 1] = 0; }
    if (byte_offset) *byte_offset = n.dbg[i].p ? (long long)((const char*)n.dbg[i].p - dummy) : -1;
    if (C) *C = n.dbg[i].C;
    if (P) *P = n.dbg[i].P;
    return 0;
}

// Number of kernel launches issued by this library since it was loaded (bench.py's gpu_launches evidence).
WF_API long long wf_launch_count(void) { return g_launches.load(); }

// Per-launch timings of the calls made on this thread with WF_FLAG_PROFILE since the last wf_profile_reset().
// Synchronises the events; returns the number of records, or fills record i.
WF_API int wf_profile_count(void) { return (int)g_prof.size(); }
WF_API int wf_profile_read(int i, char* name, int name_cap, float* ms, double* flops)
{
    if (i < 0 || i >= (int)g_prof.size()) return WF_E_ARG;
    ProfRec& r = g_prof[i];
    cudaError_t e = cudaEventSynchronize(r.b);
    if (e != cudaSuccess) return (int)e;
    if (name && name_cap > 0) { std::strncpy(name, r.name.c_str(), name_cap - 1); name[name_cap - 1] = 0; }
    if (ms) cudaEventElapsedTime(ms, r.a, r.b);
    if (flops) *flops = r.flops;
    return 0;
}
WF_API int wf_profile_bytes(int i, double* bytes)
{
    if (i < 0 || i >= (int)g_prof.size() || !bytes) return WF_E_ARG;
    *bytes = g_prof[i].bytes;
    return 0;
}
WF_API void wf_profile_reset(void)
{
    for (auto& r : g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    g_prof.clear();
}

int wf_block_forward(const wf_block_desc* d, const float* x, const float* params, float* running, long long* nbt,
                     const float* const* masks, float* y, void* ws, size_t ws_bytes, int B, int flags, wf_stream_t stream)
{
    return run_forward(d, x, params, running, nbt, masks, y, ws, ws_bytes, B, flags, (cudaStream_t)stream);
}

int wf_block_backward(const wf_block_desc* d, const float* x, const float* params, const float* const* masks, const float* dy,
                      float* grads, float* dx, void* ws, size_t ws_bytes, int B, int flags, wf_stream_t stream)
{
    return run_backward(d, x, params, masks, dy, grads, dx, ws, ws_bytes, B, flags, (cudaStream_t)stream);
}

int wf_pose_loss(const float* pred, const float* target, int B, int loss_type, float pw, float bw, const float* gscale, float* dpred,
                 float* out3, double* scratch, wf_stream_t stream)
{
    if (int e = check_device()) return e;
    if (!pred || !target || !out3 || !scratch || B <= 0) return fail(WF_E_ARG, "null pointer or non-positive batch");
    if (loss_type < 0 || loss_type > 2) return fail(WF_E_ARG, "Unknown loss type");
    g_launches.fetch_add(2);
    cudaError_t e = wf_launch_pose_loss(pred, target, B, loss_type, pw, bw, gscale, dpred, scratch, out3, (cudaStream_t)stream);
    return e == cudaSuccess ? 0 : fail((int)e, cudaGetErrorString(e));
}

int wf_pose_metrics(const float* pred, const float* target, int B, const float* thresholds, int nthr, int torso, float* out, void* scratch,
                    wf_stream_t stream)
{
    if (int e = check_device()) return e;
    if (!pred || !target || !out || !scratch || B <= 0 || nthr < 0 || nthr > WF_MAX_THR) return fail(WF_E_ARG, "bad argument (nthr <= 8)");
    MetricThr t{};
    t.n = nthr;
    for (int i = 0; i < nthr; ++i) t.v[i] = thresholds[i];
    unsigned long long* counts = (unsigned long long*)scratch;
    double* dsum = (double*)scratch + WF_MAX_THR;
    g_launches.fetch_add(2);
    cudaError_t e = wf_launch_metrics(pred, target, B, t, torso, counts, dsum, out, (cudaStream_t)stream);
    return e == cudaSuccess ? 0 : fail((int)e, cudaGetErrorString(e));
}

int wf_clip_adamw(float* params, const float* grads, float* m, float* v, long long n, void* state, float lr, float b1, float b2, float eps,
                  float wd, float max_norm, float grad_scale, wf_stream_t stream)
{
    if (int e = check_device()) return e;
    if (!params || !grads || !m || !v || !state || n <= 0) return fail(WF_E_ARG, "null pointer");
    g_launches.fetch_add(3);
    cudaError_t e = wf_launch_adamw(params, grads, m, v, n, (AdamState*)state, lr, b1, b2, eps, wd, max_norm, grad_scale, num_sms(), (cudaStream_t)stream);
    return e == cudaSuccess ? 0 : fail((int)e, cudaGetErrorString(e));
}

int wf_window_load(const float* windows, long long n_windows, const long long* idx, float* x, int B, int C, int T, int t_major,
                   const int* spans, double* stats, wf_stream_t stream)
{
    if (int e = check_device()) return e;
    if (!windows || !x || B < 0 || C <= 0 || T <= 0 || n_windows <= 0) return fail(WF_E_ARG, "null pointer or bad window geometry");
    if (((uintptr_t)windows & 15) || ((uintptr_t)x & 15)) return fail(WF_E_ARG, "window buffers must be 16-byte aligned");
    if (((long long)C * T) % 4 != 0 && (B > 1 || idx)) return fail(WF_E_ARG, "C*T must be a multiple of 4 (16-byte aligned windows) unless a single window is processed in place of the whole buffer");
    if (spans && (size_t)C * T * sizeof(float) > 200 * 1024) return fail(WF_E_UNSUPPORTED, "window larger than 200 KB of shared memory");
    if (B == 0) return 0;
    g_launches.fetch_add(1);
    cudaError_t e = wf_launch_window_load(windows, idx, n_windows, x, B, C, T, t_major ? 1 : T, t_major ? C : 1, spans, stats, (cudaStream_t)stream);
    return e == cudaSuccess ? 0 : fail((int)e, cudaGetErrorString(e));
}

int wf_noise_scale(const float* x, const float* noise, float* y, long long n, float noise_level, float scale, const double* stats,
                   long long n_stat, wf_stream_t stream)
{
    if (int e = check_device()) return e;
    if (!x || !y || n < 0 || (noise && (!stats || n_stat < 2))) return fail(WF_E_ARG, "null pointer, or noise without the statistics of x");
    if (((uintptr_t)x & 15) || ((uintptr_t)y & 15) || ((uintptr_t)noise & 15)) return fail(WF_E_ARG, "buffers must be 16-byte aligned");
    if (n == 0) return 0;
    g_launches.fetch_add(1);
    cudaError_t e = wf_launch_noise_scale(x, noise, y, n, noise_level, scale, stats, n_stat, num_sms(), (cudaStream_t)stream);
    return e == cudaSuccess ? 0 : fail((int)e, cudaGetErrorString(e));
}

int wf_dropout_masks(float* const* out, const long long* numel, const float* p, int n_sites, unsigned long long seed,
                     unsigned long long* state, wf_stream_t stream)
{
    if (int e = check_device()) return e;
    if (n_sites == 0) return 0;
    if (!out || !numel || !p || !state || n_sites < 0 || n_sites > WF_MAX_MASK_SITES) return fail(WF_E_ARG, "null pointer or more than 32 sites");
    MaskTable tab{};
    for (int i = 0; i < n_sites; ++i) {
        if (!out[i] || numel[i] < 0 || !(p[i] >= 0.f && p[i] < 1.f)) return fail(WF_E_ARG, "mask site: null buffer, negative size or p outside [0,1)");
        if ((uintptr_t)out[i] & 15) return fail(WF_E_ARG, "mask buffers must be 16-byte aligned");
        if (numel[i] == 0) continue;
        tab.s[tab.n++] = MaskSite{out[i], numel[i], p[i]};
    }
    if (tab.n == 0) return 0;
    g_launches.fetch_add(1);
    cudaError_t e = wf_launch_dropout_masks(tab, seed, state, num_sms(), (cudaStream_t)stream);
    return e == cudaSuccess ? 0 : fail((int)e, cudaGetErrorString(e));
}

int wf_keypoint_batch(const float* frames, long long n_frames, const long long* idx, float* y, int B, int K, int clean, wf_stream_t stream)
{
    if (int e = check_device()) return e;
    if (!frames || !y || B < 0 || K <= 0 || n_frames < 0) return fail(WF_E_ARG, "null pointer or bad geometry");
    if (((uintptr_t)frames & 7) || ((uintptr_t)y & 7)) return fail(WF_E_ARG, "keypoint buffers must be 8-byte aligned");
    if (B == 0) return 0;
    g_launches.fetch_add(1);
    cudaError_t e = wf_launch_keypoint_repair(frames, n_frames, idx, y, B, K, clean, (cudaStream_t)stream);
    return e == cudaSuccess ? 0 : fail((int)e, cudaGetErrorString(e));
}

int wf_keypoint_sequences(float* frames, const long long* seq_off, int n_seq, int K, wf_stream_t stream)
{
    if (int e = check_device()) return e;
    if (!frames || !seq_off || n_seq < 0 || K <= 0) return fail(WF_E_ARG, "null pointer or bad geometry");
    if ((uintptr_t)frames & 7) return fail(WF_E_ARG, "keypoint buffer must be 8-byte aligned");
    if (n_seq == 0) return 0;
    g_launches.fetch_add(1);
    cudaError_t e = wf_launch_keypoint_seq(frames, seq_off, n_seq, K, (cudaStream_t)stream);
    return e == cudaSuccess ? 0 : fail((int)e, cudaGetErrorString(e));
}

}  // extern "C"
