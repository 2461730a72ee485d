// Standalone self-test of the tcgen05 pointwise-conv kernels (wf_tc.cu) against a CPU loop: structured inputs that expose
// operand-layout mistakes, then random data at the real layer shapes.  Built by build.py, run by tests/test_gpu_native.py.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include "../../wiflow-wifi-pose-estimation-with-spatio-temporal-decoupling_b200/csrc/wf_elem.h"

thread_local int wf_pdl_mode = 0;      // launch switch of wf_common.cuh (defined by wf_model.cu in the library)

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)

static float frand() { return (float)rand() / RAND_MAX * 2.f - 1.f; }

static int run_conv(int M, int K, int N, int structured, int verbose)
{
    std::vector<float> W((size_t)M * K), X((size_t)K * N), D((size_t)M * N), R((size_t)M * N);
    for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) W[(size_t)m * K + k] = structured == 2 ? 1.f : structured ? (m == k ? 1.f : 0.f) : frand();
    for (int k = 0; k < K; ++k) for (int n = 0; n < N; ++n) X[(size_t)k * N + n] = structured == 2 ? 1.f : structured ? (float)(k * 1000 + n) : frand();
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
        double a = 0; for (int k = 0; k < K; ++k) a += (double)W[(size_t)m * K + k] * X[(size_t)k * N + n];
        R[(size_t)m * N + n] = (float)a;
    }
    float *dW, *dX, *dD, *dP; double* dS;
    const long long pf = wf_tc_pack_floats(M, K), pb = wf_tc_pack_floats(K, M);
    CK(cudaMalloc(&dW, W.size() * 4)); CK(cudaMalloc(&dX, X.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
    CK(cudaMalloc(&dP, (pf + pb) * 4)); CK(cudaMalloc(&dS, 2 * M * sizeof(double)));
    CK(cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0xff, D.size() * 4)); CK(cudaMemset(dS, 0, 2 * M * sizeof(double)));
    TcPackTable tab{}; tab.n = 1; tab.e[0] = TcPackEntry{0, M, K, 0, pf};
    CK(wf_launch_tc_pack(tab, dW, dP, 0));
    if (structured == 2) {     // verify the packed image on the host
        std::vector<float> P(pf + pb); CK(cudaMemcpy(P.data(), dP, P.size() * 4, cudaMemcpyDeviceToHost));
        int nz = 0; for (float v : P) nz += (v != 0.f);
        printf("   packed image: %lld floats, %d non-zero, first hi %g %g lo %g\n", (long long)P.size(), nz, P[0], P[1], P[2048]);
    }
    ConvP p{};
    p.in = dX; p.in_sc = N; p.in_sp = 0; p.in_sb = WF_T; p.pro_mode = PRO_NONE;
    p.Cin = K; p.Cout = M; p.groups = 1; p.Pin = 1; p.Pout = 1; p.N = N; p.ntaps = 1; p.pmul = 1; p.pdiv = 1;
    p.out = dD; p.out_sc = N; p.out_sp = 0; p.out_sb = WF_T; p.epi_mode = EPI_STATS; p.stat0 = dS; p.stat1 = dS + M;
    p.wtc = dP; p.tc_kt = (K + TC_KC - 1) / TC_KC;
    CK(wf_launch_tc_conv(p, 148, 0));
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    std::vector<double> S(2 * M); CK(cudaMemcpy(S.data(), dS, 2 * M * sizeof(double), cudaMemcpyDeviceToHost));
    double maxe = 0, maxr = 0; int bad = 0;
    for (size_t i = 0; i < D.size(); ++i) {
        double e = fabs((double)D[i] - R[i]); if (!(e <= 1e30)) e = 1e30;
        if (e > maxe) maxe = e; if (fabs(R[i]) > maxr) maxr = fabs(R[i]);
        if (e > 1e-4 * (1 + fabs(R[i])) && bad < verbose) {
            int m = (int)(i / N), n = (int)(i % N);
            printf("   mismatch D[%d][%d] = %g expected %g", m, n, D[i], R[i]);
            if (structured) { int v = (int)lrintf(D[i]); printf("  (= X[%d][%d])", v / 1000, v % 1000); }
            printf("\n"); ++bad;
        }
    }
    double se = 0;      // per-channel sums, error relative to the sum of magnitudes
    for (int m = 0; m < M; ++m) { double s = 0, sa = 0; for (int n = 0; n < N; ++n) { s += R[(size_t)m * N + n]; sa += fabs(R[(size_t)m * N + n]); } se = fmax(se, fabs(s - S[m]) / (1e-30 + sa)); }
    printf("conv  M=%d K=%d N=%d %s: max abs err %.3g (max |ref| %.3g) rel %.3g  stat-sum rel err %.3g  -> %s\n", M, K, N, structured ? "structured" : "random",
           maxe, maxr, maxe / maxr, se, (maxe <= 5e-6 * maxr && se < 5e-6) ? "OK" : "FAIL");
    cudaFree(dW); cudaFree(dX); cudaFree(dD); cudaFree(dP); cudaFree(dS);
    return (maxe <= 5e-6 * maxr && se < 5e-6) ? 0 : 1;
}

static int run_wgrad(int M, int C, int N, int structured, int verbose)
{
    // dW[m][c] = sum_n G[m][n] * X[c][n]
    std::vector<float> G((size_t)M * N), X((size_t)C * N), D((size_t)M * C), R((size_t)M * C);
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) G[(size_t)m * N + n] = structured ? (n == m % N ? 1.f : 0.f) : frand();
    for (int c = 0; c < C; ++c) for (int n = 0; n < N; ++n) X[(size_t)c * N + n] = structured ? (float)(c * 1000 + n) : frand();
    for (int m = 0; m < M; ++m) for (int c = 0; c < C; ++c) {
        double a = 0; for (int n = 0; n < N; ++n) a += (double)G[(size_t)m * N + n] * X[(size_t)c * N + n];
        R[(size_t)m * C + c] = (float)a;
    }
    float *dG, *dX, *dD;
    CK(cudaMalloc(&dG, G.size() * 4)); CK(cudaMalloc(&dX, X.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
    CK(cudaMemcpy(dG, G.data(), G.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0, D.size() * 4));
    WgradP p{};
    p.g = dG; p.g_pro = PRO_NONE; p.in = dX; p.in_sc = N; p.in_sp = 0; p.in_sb = WF_T; p.pro_mode = PRO_NONE;
    p.Cin = C; p.Cout = M; p.groups = 1; p.Pin = 1; p.Pout = 1; p.N = N; p.ntaps = 1; p.pmul = 1; p.dw = dD;
    CK(wf_launch_tc_wgrad(p, 148, 0));
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    double maxe = 0, maxr = 0; int bad = 0;
    for (size_t i = 0; i < D.size(); ++i) {
        double e = fabs((double)D[i] - R[i]); if (!(e <= 1e30)) e = 1e30;
        if (e > maxe) maxe = e; if (fabs(R[i]) > maxr) maxr = fabs(R[i]);
        if (e > 1e-4 * (1 + fabs(R[i])) && bad < verbose) {
            int m = (int)(i / C), c = (int)(i % C);
            printf("   mismatch dW[%d][%d] = %g expected %g", m, c, D[i], R[i]);
            if (structured) { int v = (int)lrintf(D[i]); printf("  (= X[%d][%d])", v / 1000, v % 1000); }
            printf("\n"); ++bad;
        }
    }
    printf("wgrad M=%d C=%d N=%d %s: max abs err %.3g (max |ref| %.3g) rel %.3g -> %s\n", M, C, N, structured ? "structured" : "random", maxe, maxr,
           maxe / maxr, maxe <= 5e-6 * maxr ? "OK" : "FAIL");
    cudaFree(dG); cudaFree(dX); cudaFree(dD);
    return maxe <= 5e-6 * maxr ? 0 : 1;
}

// timing only (no CPU check): the real layer shape at the benchmark batch
static int run_perf(int M, int K, int N, int pro)
{
    float *dW, *dX, *dD, *dP, *dA; double* dS;
    const long long pf = wf_tc_pack_floats(M, K), pb = wf_tc_pack_floats(K, M);
    CK(cudaMalloc(&dW, (size_t)M * K * 4)); CK(cudaMalloc(&dX, (size_t)K * N * 4)); CK(cudaMalloc(&dD, (size_t)M * N * 4));
    CK(cudaMalloc(&dP, (pf + pb) * 4)); CK(cudaMalloc(&dS, 2 * M * sizeof(double))); CK(cudaMalloc(&dA, 4 * (size_t)(M + K) * 4));
    std::vector<float> W((size_t)M * K), X((size_t)K * N), A(4 * (size_t)(M + K), 0.5f);
    for (auto& v : W) v = frand(); for (auto& v : X) v = frand();
    CK(cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemset(dS, 0, 2 * M * sizeof(double)));
    TcPackTable tab{}; tab.n = 1; tab.e[0] = TcPackEntry{0, M, K, 0, pf};
    CK(wf_launch_tc_pack(tab, dW, dP, 0));
    ConvP p{};
    p.in = dX; p.in2 = dX; p.in_sc = N; p.in_sp = 0; p.in_sb = WF_T; p.pro_mode = pro; p.pro_a = dA; p.pro_b = dA + K; p.pro_c = dA + 2 * K; p.pro_d = dA + 3 * K;
    p.Cin = K; p.Cout = M; p.groups = 1; p.Pin = 1; p.Pout = 1; p.N = N; p.ntaps = 1; p.pmul = 1; p.pdiv = 1;
    p.out = dD; p.out_sc = N; p.out_sp = 0; p.out_sb = WF_T; p.epi_mode = EPI_STATS; p.stat0 = dS; p.stat1 = dS + M;
    p.wtc = dP; p.tc_kt = (K + TC_KC - 1) / TC_KC;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 2; ++i) CK(wf_launch_tc_conv(p, 148, 0));
    cudaEventRecord(a);
    for (int i = 0; i < 5; ++i) CK(wf_launch_tc_conv(p, 148, 0));
    cudaEventRecord(b); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("perf conv  M=%d K=%d N=%d pro=%d: %.1f us  (%.1f TFLOP/s fp32-equivalent)\n", M, K, N, pro, ms * 200, 2.0 * M * K * N / (ms / 5 * 1e-3) / 1e12);
    WgradP w{};
    w.g = dD; w.g2 = dD; w.g_pro = pro == PRO_NONE ? PRO_NONE : PRO_BNBWD; w.g_a = dA + 4 * K; w.g_b = dA + 4 * K + M; w.g_c = dA + 4 * K + 2 * M; w.g_d = dA + 4 * K + 3 * M;
    w.in = dX; w.in_sc = N; w.in_sp = 0; w.in_sb = WF_T; w.pro_mode = pro; w.pro_a = dA; w.pro_b = dA + K; w.pro_d = dA + 3 * K;
    w.Cin = K; w.Cout = M; w.groups = 1; w.Pin = 1; w.Pout = 1; w.N = N; w.ntaps = 1; w.pmul = 1; w.dw = dW;
    for (int i = 0; i < 2; ++i) CK(wf_launch_tc_wgrad(w, 148, 0));
    cudaEventRecord(a);
    for (int i = 0; i < 5; ++i) CK(wf_launch_tc_wgrad(w, 148, 0));
    cudaEventRecord(b); CK(cudaDeviceSynchronize());
    cudaEventElapsedTime(&ms, a, b);
    printf("perf wgrad M=%d C=%d N=%d pro=%d: %.1f us  (%.1f TFLOP/s fp32-equivalent)\n", M, K, N, pro, ms * 200, 2.0 * M * K * N / (ms / 5 * 1e-3) / 1e12);
    cudaFree(dW); cudaFree(dX); cudaFree(dD); cudaFree(dP); cudaFree(dS); cudaFree(dA);
    return 0;
}

int main(int argc, char** argv)
{
    if (argc > 1 && argv[1][0] == 'p') {
        run_perf(540, 540, 20480, PRO_NONE);
        run_perf(540, 540, 20480, PRO_BNSILU);
        run_perf(240, 340, 20480, PRO_BNSILU);
        run_perf(192, 64, 15 * 20480, PRO_AFFINE);
        run_perf(64, 192, 15 * 20480, PRO_BNBWD);      // backward-data of the attention qkv projection
        return 0;
    }
    const int verbose = argc > 1 ? atoi(argv[1]) : 12;
    int fails = 0;
    srand(1);
    fails += run_conv(128, 16, 256, 2, verbose);
    fails += run_conv(128, 32, 64, 1, verbose);
    fails += run_conv(128, 16, 256, 1, verbose);
    fails += run_conv(128, 16, 320, 0, verbose);
    fails += run_conv(128, 32, 256, 1, verbose);
    fails += run_conv(128, 64, 64 * 20, 0, verbose);
    fails += run_conv(540, 540, 80, 0, verbose);
    fails += run_conv(440, 540, 1280, 0, verbose);
    fails += run_conv(240, 340, 1024 * 20, 0, verbose);
    fails += run_wgrad(128, 16, 16, 1, verbose);
    fails += run_wgrad(128, 64, 256, 1, verbose);
    fails += run_wgrad(128, 64, 1280, 0, verbose);
    fails += run_wgrad(540, 540, 80, 0, verbose);
    fails += run_wgrad(340, 440, 4 * 1280, 0, verbose);
    fails += run_wgrad(192, 64, 15 * 1280, 0, verbose);
    printf("%s (%d failing cases)\n", fails ? "SELFTEST FAILED" : "SELFTEST PASSED", fails);
    return fails ? 1 : 0;
}
