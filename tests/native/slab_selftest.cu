// Standalone self-test of the TMA + tcgen05 slab kernels (wf_slabtc.cu) against a CPU loop that restates the ConvP contract:
//   out[co][opos][n] = epi( bias[co] + sum_tap sum_ci W[tap][co][ci] * pro(in)[ci][ipos(opos,tap)][n] ),
//   ipos = (opos*pmul + dp[tap]) / pdiv, valid iff divisible and inside [0, Pin).
// First a structured case (identity weights, position-coded activations) that exposes operand-layout / descriptor mistakes,
// then random data over the layer shapes of the conv stack (stride 1 / 2 forward, their backward-data forms, every prologue and
// epilogue mode, ragged column tails).  Built by build.py, run by tests/test_gpu_native.py.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include "../../wiflow-wifi-pose-estimation-with-spatio-temporal-decoupling_b200/csrc/wf_elem.h"

thread_local int wf_pdl_mode = 0;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)

static float frand() { return (float)rand() / RAND_MAX * 2.f - 1.f; }
static double silu(double x) { return x / (1.0 + exp(-x)); }
static double dsilu(double x) { double s = 1.0 / (1.0 + exp(-x)); return s * (1.0 + x * (1.0 - s)); }

struct Case {
    const char* name;
    int cin, cout, ntaps, pin, pout, pmul, pdiv, B;
    int dp[3];
    int pro, epi, mask, bias, accumulate, bwd_image, structured;
};

template <class T> static T* dev(const std::vector<T>& h)
{
    T* d = nullptr;
    if (cudaMalloc(&d, h.size() * sizeof(T) + 16) != cudaSuccess) return nullptr;
    cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
    return d;
}

static int g_bench = 0;          // bench mode: time the launch, skip the CPU reference

static int run_case(const Case& c, int verbose)
{
    const int N = c.B * WF_T;
    // layer weights in the reference layout [cout_l][cin_l][ntaps]; a backward-data case contracts over cout_l
    const int cout_l = c.bwd_image ? c.cin : c.cout, cin_l = c.bwd_image ? c.cout : c.cin;
    std::vector<float> WL((size_t)cout_l * cin_l * c.ntaps);
    for (auto& v : WL) v = frand() * 0.3f;
    auto Wat = [&](int tap, int co, int ci) -> float {      // weight of (tap, output channel co, input channel ci) of THIS conv
        return c.bwd_image ? WL[((size_t)ci * cin_l + co) * c.ntaps + tap] : WL[((size_t)co * cin_l + ci) * c.ntaps + tap];
    };
    if (c.structured) {
        for (auto& v : WL) v = 0.f;
        for (int i = 0; i < c.cin && i < c.cout; ++i) WL[((size_t)i * cin_l + i) * c.ntaps + 0] = 1.f;
    }
    std::vector<float> X((size_t)c.cin * c.pin * N), X2(X.size());
    for (size_t i = 0; i < X.size(); ++i) { X[i] = frand(); X2[i] = frand(); }
    if (c.structured) for (int ci = 0; ci < c.cin; ++ci) for (int q = 0; q < c.pin; ++q) for (int n = 0; n < N; ++n)
        X[((size_t)ci * c.pin + q) * N + n] = (float)(ci * 4096 + q * 512 + (n % 512));
    std::vector<float> pa(c.cin), pb(c.cin), pc(c.cin), pd(c.cin), bias(c.cout), es(c.cout), et(c.cout), em(c.cout);
    for (int i = 0; i < c.cin; ++i) { pa[i] = 0.5f + 0.5f * fabsf(frand()); pb[i] = frand() * 0.2f; pc[i] = frand() * 0.1f; pd[i] = frand() * 0.3f; }
    for (int i = 0; i < c.cout; ++i) { bias[i] = c.bias ? frand() : 0.f; es[i] = 0.5f + fabsf(frand()); et[i] = frand() * 0.2f; em[i] = frand() * 0.3f; }
    std::vector<float> mask((size_t)c.B * c.cin), emask((size_t)c.B * c.cout);
    for (auto& v : mask) v = (rand() % 10 < 3) ? 0.f : 1.f / 0.7f;
    for (auto& v : emask) v = (rand() % 10 < 3) ? 0.f : 1.f / 0.7f;
    std::vector<float> eraw((size_t)c.cout * c.pout * N), out0(eraw.size());
    for (size_t i = 0; i < eraw.size(); ++i) { eraw[i] = frand(); out0[i] = c.accumulate ? frand() : NAN; }

    // ---- CPU reference ----
    std::vector<double> Xp(g_bench ? 0 : X.size());
    if (!g_bench)
    for (int ci = 0; ci < c.cin; ++ci) for (int q = 0; q < c.pin; ++q) for (int n = 0; n < N; ++n) {
        const size_t i = ((size_t)ci * c.pin + q) * N + n;
        double x = X[i];
        if (c.pro == PRO_BNSILU) { x = silu((double)pa[ci] * ((double)X[i] - pd[ci]) + pb[ci]); if (c.mask) x *= mask[(size_t)(n / WF_T) * c.cin + ci]; }
        else if (c.pro == PRO_AFFINE) x = (double)pa[ci] * ((double)X[i] - pd[ci]) + pb[ci];
        else if (c.pro == PRO_BNBWD) x = (double)pa[ci] * X[i] + (double)pb[ci] * ((double)X2[i] - pd[ci]) + pc[ci];
        Xp[i] = x;
    }
    std::vector<double> R(eraw.size()), S0(c.cout, 0.0), S1(c.cout, 0.0);
    if (!g_bench)
    for (int co = 0; co < c.cout; ++co) for (int op = 0; op < c.pout; ++op) for (int n = 0; n < N; ++n) {
        double a = bias[co];
        for (int t = 0; t < c.ntaps; ++t) {
            const int num = op * c.pmul + c.dp[t];
            if (num < 0 || num % c.pdiv) continue;
            const int q = num / c.pdiv;
            if (q >= c.pin) continue;
            for (int ci = 0; ci < c.cin; ++ci) a += (double)Wat(t, co, ci) * Xp[((size_t)ci * c.pin + q) * N + n];
        }
        const size_t o = ((size_t)co * c.pout + op) * N + n;
        if (c.accumulate) a += out0[o];
        if (c.epi == EPI_STATS) { S0[co] += a; S1[co] += a * a; }
        else if (c.epi == EPI_DSILU || c.epi == EPI_DAFF) {
            const double rw = eraw[o];
            if (c.epi == EPI_DSILU) a = a * (c.mask ? emask[(size_t)(n / WF_T) * c.cout + co] : 1.0) * dsilu((double)es[co] * (rw - em[co]) + et[co]);
            S0[co] += a; S1[co] += a * (rw - em[co]);
        }
        R[o] = a;
    }

    // ---- device ----
    float *dWL = dev(WL), *dX = dev(X), *dX2 = dev(X2), *dpa = dev(pa), *dpb = dev(pb), *dpc = dev(pc), *dpd = dev(pd), *dbias = dev(bias);
    float *des = dev(es), *det = dev(et), *dem = dev(em), *dmask = dev(mask), *demask = dev(emask), *deraw = dev(eraw), *dout = dev(out0);
    const long long pf = wf_slabtc_pack_floats(cout_l, cin_l, c.ntaps, false), pbk = wf_slabtc_pack_floats(cout_l, cin_l, c.ntaps, true);
    float* dP; double* dS;
    CK(cudaMalloc(&dP, (pf + pbk) * 4)); CK(cudaMalloc(&dS, 2 * c.cout * sizeof(double))); CK(cudaMemset(dS, 0, 2 * c.cout * sizeof(double)));
    SlabPackTable tab{}; tab.n = 1; tab.e[0] = SlabPackEntry{0, cout_l, cin_l, c.ntaps, 0, pf};
    CK(wf_launch_slabtc_pack(tab, dWL, dP, 0));
    ConvP p{};
    p.in = dX; p.in2 = dX2; p.in_sc = (long long)c.pin * N; p.in_sp = N; p.in_sb = WF_T;
    p.pro_mode = c.pro; p.pro_a = dpa; p.pro_b = dpb; p.pro_c = dpc; p.pro_d = dpd;
    if (c.pro == PRO_BNSILU && c.mask) { p.mask = dmask; p.m_sb = c.cin; p.m_sc = 1; p.m_st = 0; }
    p.wtc = dP + (c.bwd_image ? pf : 0);
    p.Cin = c.cin; p.Cout = c.cout; p.groups = 1; p.Pin = c.pin; p.Pout = c.pout; p.N = N; p.ntaps = c.ntaps; p.pmul = c.pmul; p.pdiv = c.pdiv;
    for (int t = 0; t < c.ntaps; ++t) { p.dp[t] = c.dp[t]; p.dn[t] = 0; }
    p.out = dout; p.out_sc = (long long)c.pout * N; p.out_sp = N; p.out_sb = WF_T;
    p.bias = c.bias ? dbias : nullptr; p.epi_mode = c.epi; p.accumulate = c.accumulate;
    p.eraw = deraw; p.e_scale = des; p.e_shift = det; p.e_mean = dem;
    if (c.epi == EPI_DSILU && c.mask) { p.emask = demask; p.em_sb = c.cout; p.em_sc = 1; p.em_st = 0; }
    p.stat0 = dS; p.stat1 = dS + c.cout;
    if (!wf_slabtc_conv_ok(p)) { printf("%-28s: shape declined by wf_slabtc_conv_ok -> FAIL\n", c.name); return 1; }
    CK(wf_launch_slabtc_conv(p, 0));
    CK(cudaDeviceSynchronize());
    if (g_bench) {
        static const int dbgs[] = {0, 55, 2, 6, 4, 1};
        printf("%-30s", c.name);
        for (int d : dbgs) {
            char buf[16]; snprintf(buf, sizeof buf, "%d", d); setenv("WF_SLABTC_DBG", buf, 1);
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            CK(wf_launch_slabtc_conv(p, 0));
            cudaEventRecord(e0);
            for (int it = 0; it < 20; ++it) CK(wf_launch_slabtc_conv(p, 0));
            cudaEventRecord(e1);
            CK(cudaDeviceSynchronize());
            float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
            printf(" dbg%d %7.1fus", d, ms * 50.f);
        }
        printf("\n");
        {   // timeline of CTA 0 for one full launch
            unsigned long long ts[32];
            setenv("WF_SLABTC_DBG", "512", 1);
            wf_slabtc_debug_ts(ts);
            CK(wf_launch_slabtc_conv(p, 0)); CK(cudaDeviceSynchronize());
            wf_slabtc_debug_ts(ts);
            static const char* nm[] = {"entry", "tmem+bars", "zeroed", "TMA(6) issued", "w: raw_full(6)", "w: op_full(6) arrive", "i: weights in", "i: op_full(6)", "i: commit(6)",
                                       "w: grp_done(6)", "w: epilogue(6) done", "w: loop end", "w: stats done", "final sync"};
            printf("   timeline (us from entry):");
            for (int i = 0; i < 14; ++i) printf(" %s %.1f |", nm[i], ts[i] ? (double)(ts[i] - ts[0]) / 1000.0 : -1.0);
            printf("\n");
            // all CTAs, two back-to-back launches: spread of entry / end times and the gap between the launches
            static unsigned long long ct[6][160];
            wf_slabtc_debug_cta(&ct[0][0]);
            CK(wf_launch_slabtc_conv(p, 0)); CK(wf_launch_slabtc_conv(p, 0)); CK(cudaDeviceSynchronize());
            wf_slabtc_debug_cta(&ct[0][0]);
            unsigned long long t0 = ~0ULL;
            for (int b = 0; b < 160; ++b) if (ct[0][b] && ct[0][b] < t0) t0 = ct[0][b];
            static const char* cn[] = {"L1 launch", "L1 entry", "L1 end", "L2 launch", "L2 entry", "L2 end"};
            printf("   all CTAs (us from the first CTA's start):");
            for (int k = 0; k < 6; ++k) {
                unsigned long long lo = ~0ULL, hi = 0; int n = 0;
                for (int b = 0; b < 160; ++b) if (ct[k][b]) { ++n; lo = ct[k][b] < lo ? ct[k][b] : lo; hi = ct[k][b] > hi ? ct[k][b] : hi; }
                if (n) printf(" %s %.1f..%.1f (%d) |", cn[k], (double)(lo - t0) / 1000.0, (double)(hi - t0) / 1000.0, n);
            }
            printf("\n");
        }
        unsetenv("WF_SLABTC_DBG");
        return 0;
    }
    std::vector<float> D(R.size()); std::vector<double> S(2 * c.cout);
    CK(cudaMemcpy(D.data(), dout, D.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(S.data(), dS, S.size() * sizeof(double), cudaMemcpyDeviceToHost));
    double maxe = 0, maxr = 0; int bad = 0;
    for (size_t i = 0; i < D.size(); ++i) maxr = fmax(maxr, fabs(R[i]));
    for (size_t i = 0; i < D.size(); ++i) {
        double e = fabs((double)D[i] - R[i]); if (!(e <= 1e30)) e = 1e30;
        maxe = fmax(maxe, e);
        if (e > 2e-5 * maxr + 1e-30 && bad < verbose) {
            const int n = (int)(i % N), op = (int)((i / N) % c.pout), co = (int)(i / ((size_t)N * c.pout));
            printf("   mismatch out[%d][%d][%d] = %g expected %g", co, op, n, D[i], R[i]);
            if (c.structured) { const int v = (int)lrintf(D[i]); printf("  (= X[c %d][q %d][n %d])", v / 4096, (v / 512) % 8, v % 512); }
            printf("\n"); ++bad;
        }
    }
    double se = 0;
    if (c.epi != EPI_STORE) for (int co = 0; co < c.cout; ++co) {
        double sa0 = 0, sa1 = 0;
        for (int op = 0; op < c.pout; ++op) for (int n = 0; n < N; ++n) { const size_t o = ((size_t)co * c.pout + op) * N + n; sa0 += fabs(R[o]); sa1 += c.epi == EPI_STATS ? R[o] * R[o] : fabs(R[o] * (eraw[o] - em[co])); }
        se = fmax(se, fabs(S[co] - S0[co]) / (1e-30 + sa0));
        se = fmax(se, fabs(S[c.cout + co] - S1[co]) / (1e-30 + sa1));
    }
    const bool ok = maxe <= 2e-5 * maxr && se < 1e-5;
    printf("%-28s: max abs err %.3g (max |ref| %.3g, rel %.3g)  stats rel err %.3g -> %s\n", c.name, maxe, maxr, maxe / (maxr + 1e-30), se, ok ? "ok" : "FAIL");
    cudaFree(dWL); cudaFree(dX); cudaFree(dX2); cudaFree(dpa); cudaFree(dpb); cudaFree(dpc); cudaFree(dpd); cudaFree(dbias); cudaFree(des); cudaFree(det);
    cudaFree(dem); cudaFree(dmask); cudaFree(demask); cudaFree(deraw); cudaFree(dout); cudaFree(dP); cudaFree(dS);
    return ok ? 0 : 1;
}

struct WCase { const char* name; int cin, cout, ntaps, pin, pout, pmul, B; int dp[3]; int pro, mask; };

static int run_wgrad_case(const WCase& c, int verbose, int bench)
{
    const int N = c.B * WF_T;
    std::vector<float> X((size_t)c.cin * c.pin * N), DY((size_t)c.cout * c.pout * N), RAW(DY.size());
    for (auto& v : X) v = frand();
    for (size_t i = 0; i < DY.size(); ++i) { DY[i] = frand(); RAW[i] = frand(); }
    std::vector<float> pa(c.cin), pb(c.cin), pd(c.cin), ga(c.cout), gb(c.cout), gc(c.cout), gd(c.cout), mask((size_t)c.B * c.cin);
    for (int i = 0; i < c.cin; ++i) { pa[i] = 0.5f + 0.5f * fabsf(frand()); pb[i] = frand() * 0.2f; pd[i] = frand() * 0.3f; }
    for (int i = 0; i < c.cout; ++i) { ga[i] = 0.5f + fabsf(frand()); gb[i] = frand() * 0.2f; gc[i] = frand() * 0.1f; gd[i] = frand() * 0.3f; }
    for (auto& v : mask) v = (rand() % 10 < 3) ? 0.f : 1.f / 0.7f;
    std::vector<double> R((size_t)c.cout * c.cin * c.ntaps, 0.0);
    if (!bench) {
        std::vector<double> Xp(X.size()), Gp(DY.size());
        for (int ci = 0; ci < c.cin; ++ci) for (int q = 0; q < c.pin; ++q) for (int n = 0; n < N; ++n) {
            const size_t i = ((size_t)ci * c.pin + q) * N + n;
            double x = X[i];
            if (c.pro == PRO_BNSILU) { x = silu((double)pa[ci] * ((double)X[i] - pd[ci]) + pb[ci]); if (c.mask) x *= mask[(size_t)(n / WF_T) * c.cin + ci]; }
            else if (c.pro == PRO_AFFINE) x = (double)pa[ci] * ((double)X[i] - pd[ci]) + pb[ci];
            Xp[i] = x;
        }
        for (int o = 0; o < c.cout; ++o) for (size_t j = 0; j < (size_t)c.pout * N; ++j) {
            const size_t i = (size_t)o * c.pout * N + j;
            Gp[i] = (double)ga[o] * DY[i] + (double)gb[o] * ((double)RAW[i] - gd[o]) + gc[o];
        }
        for (int o = 0; o < c.cout; ++o) for (int ci = 0; ci < c.cin; ++ci) for (int t = 0; t < c.ntaps; ++t) {
            double a = 0;
            for (int pp = 0; pp < c.pout; ++pp) {
                const int q = pp * c.pmul + c.dp[t];
                if (q < 0 || q >= c.pin) continue;
                const double* gp = &Gp[((size_t)o * c.pout + pp) * N];
                const double* xp = &Xp[((size_t)ci * c.pin + q) * N];
                for (int n = 0; n < N; ++n) a += gp[n] * xp[n];
            }
            R[((size_t)o * c.cin + ci) * c.ntaps + t] = a;
        }
    }
    float *dX = dev(X), *dDY = dev(DY), *dRAW = dev(RAW), *dpa = dev(pa), *dpb = dev(pb), *dpd = dev(pd), *dga = dev(ga), *dgb = dev(gb), *dgc = dev(gc), *dgd = dev(gd), *dmask = dev(mask);
    float* dW; CK(cudaMalloc(&dW, R.size() * 4)); CK(cudaMemset(dW, 0, R.size() * 4));
    WgradP p{};
    p.g = dDY; p.g2 = dRAW; p.g_pro = PRO_BNBWD; p.g_a = dga; p.g_b = dgb; p.g_c = dgc; p.g_d = dgd;
    p.in = dX; p.in_sc = (long long)c.pin * N; p.in_sp = N; p.in_sb = WF_T;
    p.pro_mode = c.pro; p.pro_a = dpa; p.pro_b = dpb; p.pro_d = dpd;
    if (c.pro == PRO_BNSILU && c.mask) { p.mask = dmask; p.m_sb = c.cin; p.m_sc = 1; p.m_st = 0; }
    p.Cin = c.cin; p.Cout = c.cout; p.groups = 1; p.Pin = c.pin; p.Pout = c.pout; p.N = N; p.ntaps = c.ntaps; p.pmul = c.pmul;
    for (int t = 0; t < c.ntaps; ++t) { p.dp[t] = c.dp[t]; p.dn[t] = 0; }
    p.dw = dW;
    if (!wf_slabtc_wgrad_ok(p)) { printf("%-30s: shape declined by wf_slabtc_wgrad_ok -> FAIL\n", c.name); return 1; }
    CK(wf_launch_slabtc_wgrad(p, 0));
    CK(cudaDeviceSynchronize());
    int rc = 0;
    if (bench) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        for (int it = 0; it < 20; ++it) CK(wf_launch_slabtc_wgrad(p, 0));
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        printf("%-30s wgrad %7.1fus\n", c.name, ms * 50.f);
    } else {
        std::vector<float> D(R.size());
        CK(cudaMemcpy(D.data(), dW, D.size() * 4, cudaMemcpyDeviceToHost));
        double maxe = 0, maxr = 0; int bad = 0;
        for (size_t i = 0; i < D.size(); ++i) maxr = fmax(maxr, fabs(R[i]));
        for (size_t i = 0; i < D.size(); ++i) {
            double e = fabs((double)D[i] - R[i]); if (!(e <= 1e30)) e = 1e30;
            maxe = fmax(maxe, e);
            if (e > 2e-5 * maxr && bad < verbose) {
                const int t = (int)(i % c.ntaps), ci = (int)((i / c.ntaps) % c.cin), o = (int)(i / ((size_t)c.ntaps * c.cin));
                printf("   mismatch dW[%d][%d][%d] = %g expected %g\n", o, ci, t, D[i], R[i]); ++bad;
            }
        }
        const bool ok = maxe <= 2e-5 * maxr;
        printf("%-30s: max abs err %.3g (max |ref| %.3g, rel %.3g) -> %s\n", c.name, maxe, maxr, maxe / (maxr + 1e-30), ok ? "ok" : "FAIL");
        rc = ok ? 0 : 1;
    }
    cudaFree(dX); cudaFree(dDY); cudaFree(dRAW); cudaFree(dpa); cudaFree(dpb); cudaFree(dpd); cudaFree(dga); cudaFree(dgb); cudaFree(dgc); cudaFree(dgd); cudaFree(dmask); cudaFree(dW);
    return rc;
}

static const WCase g_wcases[] = {
    {"wgrad 8->8 s1 P=24",            8,  8, 3, 24, 24, 1, 13, {-1, 0, 1}, PRO_BNSILU, 1},
    {"wgrad 8->8 s2 P=60->30",        8,  8, 3, 60, 30, 2, 7,  {-1, 0, 1}, PRO_NONE, 0},
    {"wgrad 8->16 s2 P=120->60",      8, 16, 3, 120, 60, 2, 5, {-1, 0, 1}, PRO_NONE, 0},
    {"wgrad 8->16 shortcut 1 tap",    8, 16, 1, 120, 60, 2, 5, {0, 0, 0},  PRO_NONE, 0},
    {"wgrad 16->16 s1 P=60",         16, 16, 3, 60, 60, 1, 9,  {-1, 0, 1}, PRO_BNSILU, 1},
    {"wgrad 16->32 s2 P=60->30",     16, 32, 3, 60, 30, 2, 11, {-1, 0, 1}, PRO_NONE, 0},
    {"wgrad 32->32 s1 P=30",         32, 32, 3, 30, 30, 1, 10, {-1, 0, 1}, PRO_BNSILU, 0},
    {"wgrad 32->64 s2 P=30->15",     32, 64, 3, 30, 15, 2, 12, {-1, 0, 1}, PRO_NONE, 0},
    {"wgrad 32->64 shortcut 1 tap",  32, 64, 1, 30, 15, 2, 12, {0, 0, 0},  PRO_NONE, 0},
    {"wgrad 64->64 s1 P=15",         64, 64, 3, 15, 15, 1, 21, {-1, 0, 1}, PRO_BNSILU, 1},
    {"wgrad 64->32 affine 1 tap",    64, 32, 1, 15, 15, 1, 8,  {0, 0, 0},  PRO_AFFINE, 0},
};
static const WCase g_wbench[] = {
    {"wgrad 8->8 s1 P=240 B=1024",    8,  8, 3, 240, 240, 1, 1024, {-1, 0, 1}, PRO_BNSILU, 1},
    {"wgrad 16->16 s1 P=60 B=1024",  16, 16, 3, 60, 60, 1, 1024,  {-1, 0, 1}, PRO_BNSILU, 1},
    {"wgrad 32->32 s1 P=30 B=1024",  32, 32, 3, 30, 30, 1, 1024,  {-1, 0, 1}, PRO_BNSILU, 1},
    {"wgrad 64->64 s1 P=15 B=1024",  64, 64, 3, 15, 15, 1, 1024,  {-1, 0, 1}, PRO_BNSILU, 1},
    {"wgrad 32->64 s2 P=30 B=1024",  32, 64, 3, 30, 15, 2, 1024,  {-1, 0, 1}, PRO_NONE, 0},
};

int main(int argc, char** argv)
{
    setenv("WF_SLABTC_THIN", "1", 1);      // the self-test covers the 8-channel shapes too
    setenv("WF_SLABTC_MIN_N", "0", 1);     // ... and batches below the size the model switches to these kernels at
    const int verbose = argc > 1 ? atoi(argv[1]) : 6;
    const int only = argc > 2 ? atoi(argv[2]) : -1;
    srand(1234);
    const Case cases[] = {
        // name                         cin cout taps pin pout pmul pdiv  B   dp          pro          epi        mask bias acc bwd structured
        {"structured 8->8 1 tap",         8,  8, 1,  1,  1, 1, 1,   7, {0, 0, 0},   PRO_NONE,   EPI_STORE, 0, 0, 0, 0, 1},
        {"structured 64->64 1 tap",      64, 64, 1,  2,  2, 1, 1,   7, {0, 0, 0},   PRO_NONE,   EPI_STORE, 0, 0, 0, 0, 1},
        {"8->8 s1 P=12",                  8,  8, 3, 12, 12, 1, 1,  13, {-1, 0, 1},  PRO_NONE,   EPI_STATS, 0, 1, 0, 0, 0},
        {"8->8 s1 P=240 bnsilu mask",     8,  8, 3, 240, 240, 1, 1, 16, {-1, 0, 1}, PRO_BNSILU, EPI_STATS, 1, 1, 0, 0, 0},
        {"8->16 s2 P=120->60",            8, 16, 3, 120, 60, 2, 1, 11, {-1, 0, 1},  PRO_NONE,   EPI_STATS, 0, 1, 0, 0, 0},
        {"8->16 s2 shortcut 1 tap",       8, 16, 1, 120, 60, 2, 1, 11, {0, 0, 0},   PRO_NONE,   EPI_STATS, 0, 0, 0, 0, 0},
        {"16->16 s1 P=60 bnsilu",        16, 16, 3, 60, 60, 1, 1,  9, {-1, 0, 1},   PRO_BNSILU, EPI_STATS, 1, 1, 0, 0, 0},
        {"32->64 s2 P=30->15",           32, 64, 3, 30, 15, 2, 1, 40, {-1, 0, 1},   PRO_NONE,   EPI_STATS, 0, 1, 0, 0, 0},
        {"64->64 s1 P=15 bnsilu mask",   64, 64, 3, 15, 15, 1, 1, 64, {-1, 0, 1},   PRO_BNSILU, EPI_STATS, 1, 1, 0, 0, 0},
        {"64->64 s1 P=15 eval store",    64, 64, 3, 15, 15, 1, 1,  5, {-1, 0, 1},   PRO_BNSILU, EPI_STORE, 0, 1, 0, 0, 0},
        {"dgrad 64->64 s1 dsilu mask",   64, 64, 3, 15, 15, 1, 1, 33, {1, 0, -1},   PRO_BNBWD,  EPI_DSILU, 1, 0, 0, 1, 0},
        {"dgrad 64->32 s2 15->30 store", 64, 32, 3, 15, 30, 1, 2, 21, {1, 0, -1},   PRO_BNBWD,  EPI_STORE, 0, 0, 1, 1, 0},
        {"dgrad 16->8 s2 1 tap store",   16,  8, 1, 60, 120, 1, 2, 10, {0, 0, 0},   PRO_BNBWD,  EPI_STORE, 0, 0, 0, 1, 0},
        {"dgrad 8->8 s1 P=240 dsilu",     8,  8, 3, 240, 240, 1, 1, 12, {1, 0, -1}, PRO_BNBWD,  EPI_DSILU, 1, 0, 0, 1, 0},
        {"dgrad 32->32 s1 daff",         32, 32, 3, 30, 30, 1, 1, 17, {1, 0, -1},   PRO_BNBWD,  EPI_DAFF,  0, 0, 0, 1, 0},
        {"affine 64->32 1 tap",          64, 32, 1, 15, 15, 1, 1, 19, {0, 0, 0},    PRO_AFFINE, EPI_STATS, 0, 1, 0, 0, 0},
        {"8->8 s1 P=240 B=256",           8,  8, 3, 240, 240, 1, 1, 256, {-1, 0, 1}, PRO_BNSILU, EPI_STATS, 1, 1, 0, 0, 0},
        {"64->64 s1 P=15 B=512",         64, 64, 3, 15, 15, 1, 1, 512, {-1, 0, 1},  PRO_BNSILU, EPI_STATS, 1, 1, 0, 0, 0},
    };
    if (argc > 1 && !strcmp(argv[1], "bench")) {
        g_bench = 1;
        printf("per-launch time; dbg bits: 1 no transform, 2 no MMA, 4 no epilogue memory traffic, 8 hi*hi MMA only\n");
        const Case bc[] = {
            {"64->64 s1 P=15 B=128 fwd",     64, 64, 3, 15, 15, 1, 1, 128, {-1, 0, 1},   PRO_BNSILU, EPI_STATS, 1, 1, 0, 0, 0},
            {"64->64 s1 P=15 B=256 fwd",     64, 64, 3, 15, 15, 1, 1, 256, {-1, 0, 1},   PRO_BNSILU, EPI_STATS, 1, 1, 0, 0, 0},
            {"64->64 s1 P=15 B=512 fwd",     64, 64, 3, 15, 15, 1, 1, 512, {-1, 0, 1},   PRO_BNSILU, EPI_STATS, 1, 1, 0, 0, 0},
            {"64->64 s1 P=15 B=2048 fwd",    64, 64, 3, 15, 15, 1, 1, 2048, {-1, 0, 1},  PRO_BNSILU, EPI_STATS, 1, 1, 0, 0, 0},
            {"8->8 s1 P=240 B=256 fwd",       8,  8, 3, 240, 240, 1, 1, 256, {-1, 0, 1}, PRO_BNSILU, EPI_STATS, 1, 1, 0, 0, 0},
            {"8->8 s1 P=240 B=1024 fwd",      8,  8, 3, 240, 240, 1, 1, 1024, {-1, 0, 1}, PRO_BNSILU, EPI_STATS, 1, 1, 0, 0, 0},
            {"16->16 s1 P=60 B=1024 fwd",    16, 16, 3, 60, 60, 1, 1, 1024, {-1, 0, 1},   PRO_BNSILU, EPI_STATS, 1, 1, 0, 0, 0},
            {"64->64 s1 P=15 B=1024 fwd",    64, 64, 3, 15, 15, 1, 1, 1024, {-1, 0, 1},   PRO_BNSILU, EPI_STATS, 1, 1, 0, 0, 0},
            {"64->64 s1 P=15 B=1024 dgrad",  64, 64, 3, 15, 15, 1, 1, 1024, {1, 0, -1},   PRO_BNBWD,  EPI_DSILU, 1, 0, 0, 1, 0},
            {"8->8 s1 P=240 B=1024 dgrad",    8,  8, 3, 240, 240, 1, 1, 1024, {1, 0, -1}, PRO_BNBWD,  EPI_DSILU, 1, 0, 0, 1, 0},
        };
        const int pick = argc > 2 ? atoi(argv[2]) : -1;          // bench <i>: only case i (conv cases first, then the wgrad cases) -- for ncu
        int k = 0;
        for (const Case& c : bc) if ((pick < 0 || pick == k++) && run_case(c, 0) == 2) return 2;
        k = (int)(sizeof(bc) / sizeof(bc[0]));
        for (const WCase& c : g_wbench) if ((pick < 0 || pick == k++) && run_wgrad_case(c, 0, 1) == 2) return 2;
        return 0;
    }
    if (argc > 1 && !strcmp(argv[1], "wgrad")) {
        int f = 0;
        for (const WCase& c : g_wcases) { const int r = run_wgrad_case(c, 6, 0); if (r == 2) return 2; f += r; }
        for (const WCase& c : g_wbench) if (run_wgrad_case(c, 0, 1) == 2) return 2;
        printf(f ? "WGRAD FAILED (%d)\n" : "WGRAD PASSED\n", f);
        return f ? 1 : 0;
    }
    int fails = 0, idx = 0;
    for (const Case& c : cases) {
        if (only >= 0 && idx++ != only) continue;
        const int r = run_case(c, verbose);
        if (r == 2) { printf("SELFTEST ABORTED (CUDA error)\n"); return 2; }
        fails += r;
    }
    for (const WCase& c : g_wcases) {
        if (only >= 0) break;
        const int r = run_wgrad_case(c, verbose, 0);
        if (r == 2) { printf("SELFTEST ABORTED (CUDA error)\n"); return 2; }
        fails += r;
    }
    printf(fails ? "SLAB SELFTEST FAILED (%d cases)\n" : "SLAB SELFTEST PASSED\n", fails);
    return fails ? 1 : 0;
}
