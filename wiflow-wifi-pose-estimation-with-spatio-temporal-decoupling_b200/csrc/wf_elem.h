// Parameter structs + launcher prototypes shared between the kernel translation units and the host plan.
#pragma once
#include <cuda_runtime.h>
#include "wf_common.cuh"

#define WF_MAX_THR 8
#define WF_MAX_BN 48
#define WF_MAX_CONV 48
enum { WF_LOSS_SMOOTH_L1 = 0, WF_LOSS_MSE = 1, WF_LOSS_L1 = 2 };

struct BnEvalEntry { int C, Cpad, gamma_off, run_off, coef_off; };
struct BnEvalTable { int n; BnEvalEntry e[WF_MAX_BN]; };

#define WF_MAX_MASK_SITES 32
struct MaskSite { float* out; long long numel; float p; };
struct MaskTable { int n; MaskSite s[WF_MAX_MASK_SITES]; };

struct JoinP {
    const float *a, *r;
    float* out;
    const float* dout; float *dz, *da;
    long long plane; int N, C;
    int a_mode; const float *a_scale, *a_shift, *a_mean;
    const float* mask; long long m_sb, m_sc; int m_st;
    int r_mode; const float *r_scale, *r_shift, *r_mean;
    long long r_sc, r_sp, r_sb;
    double *a_stat0, *a_stat1, *r_stat0, *r_stat1;
    BnTail tail;                        // join_bwd: finalize of the (one or two) BatchNorms whose backward sums it completes
};

struct MetricThr { int n; float v[WF_MAX_THR]; };

struct PackEntry { int param_off, cout, cin, groups, ntaps, f_kpad, f_mpad, b_kpad, b_mpad; long long fwd_off, bwd_off; };
struct PackTable { int n; PackEntry e[WF_MAX_CONV]; };

struct TcPackEntry { int param_off, cout, cin; long long fwd_off, bwd_off; };
struct TcPackTable { int n; TcPackEntry e[WF_MAX_CONV]; };

struct SlabPackEntry { int param_off, cout, cin, ntaps; long long fwd_off, bwd_off; };
struct SlabPackTable { int n; SlabPackEntry e[WF_MAX_CONV]; };

#define WF_SUMSQ_BLOCKS 256
// first 64 bytes: scalars (step counter persists between calls); then one partial sum of squares per block of sumsq_kernel
struct AdamState { double sumsq; long long step; float grad_norm, clip_coef, step_size, inv_bc2_sqrt; double pad_[4]; double partial[WF_SUMSQ_BLOCKS]; };

// attention (wf_attn.cu)
struct AttnP {
    int width;                    // 1: sequences along time (L=20), rows=(h,b); 0: along slots (L=15), rows=n
    int B, N;                     // N = B*20
    const float* qkv_raw;         // [192][15][N]
    const float *qkv_scale, *qkv_shift, *qkv_mean; // bn_qkv: scale*(x - mean) + shift (192)
    const float *sim_scale, *sim_shift, *sim_mean; // bn_similarity affine + batch mean (8)
    double *sim_s0, *sim_s1;                       // stats pass output
    float* sv_raw;                // [64][15][N]
    double *sv_s0, *sv_s1;
    // backward
    const float *dsv, *sv_alpha, *sv_beta, *sv_delta, *sv_mean;   // dy of bn_output + its BN-backward affine (64)
    const float *sim_alpha, *sim_beta, *sim_delta;      // BN-backward affine of bn_similarity (8)
    double *dsim_s0, *dsim_s1;
    float* dqkv;                  // [192][15][N]  dy of bn_qkv
};

cudaError_t wf_launch_conv(const ConvP& p, cudaStream_t st);
cudaError_t wf_launch_wgrad(const WgradP& p, int num_sms, cudaStream_t st);
bool wf_group_conv_ok(const ConvP& p);
cudaError_t wf_launch_group_conv(const ConvP& p, cudaStream_t st);
bool wf_group_wgrad_ok(const WgradP& p);
cudaError_t wf_launch_group_wgrad(const WgradP& p, int num_sms, cudaStream_t st);
bool wf_thin_conv_ok(const ConvP& p);
cudaError_t wf_launch_thin_conv(const ConvP& p, cudaStream_t st);
bool wf_thin_wgrad_ok(const WgradP& p);
cudaError_t wf_launch_thin_wgrad(const WgradP& p, int num_sms, cudaStream_t st);
// sliding-window mma.sync path for position-tap convs (wf_slide.cu)
bool wf_slide_conv_ok(const ConvP& p);
cudaError_t wf_launch_slide_conv(const ConvP& p, cudaStream_t st);
bool wf_slide_conv_is_thin(const ConvP& p);
bool wf_slide_wgrad_ok(const WgradP& p);
cudaError_t wf_launch_slide_wgrad(const WgradP& p, int num_sms, cudaStream_t st);
// TMA + tcgen05 path for position-tap convs (wf_slabtc.cu)
bool wf_slabtc_shape_ok(int cin, int cout, int groups, int ntaps, const int* dn);
long long wf_slabtc_pack_floats(int cout, int cin, int ntaps, bool bwd);
cudaError_t wf_launch_slabtc_pack(const SlabPackTable& tab, const float* params, float* packed, cudaStream_t st);
bool wf_slabtc_conv_ok(const ConvP& p);
cudaError_t wf_slabtc_debug_ts(unsigned long long* out);
cudaError_t wf_slabtc_debug_cta(unsigned long long* out);      // [6][160] per-CTA stamps of two launches (WF_SLABTC_DBG bit 512)
bool wf_slabtc_wgrad_ok(const WgradP& p);
cudaError_t wf_launch_slabtc_wgrad(const WgradP& p, cudaStream_t st);
cudaError_t wf_launch_slabtc_conv(const ConvP& p, cudaStream_t st);
// tcgen05 pointwise-conv path (wf_tc.cu)
long long wf_tc_pack_floats(int m, int k);
cudaError_t wf_launch_tc_pack(const TcPackTable& tab, const float* params, float* packed, cudaStream_t st);
cudaError_t wf_launch_tc_conv(const ConvP& p, int num_sms, cudaStream_t st);
cudaError_t wf_launch_tc_wgrad(const WgradP& p, int num_sms, cudaStream_t st);
int wf_conv_bm_for(int M);
int wf_conv_bk_for(int M);

cudaError_t wf_launch_dropout_masks(const MaskTable& tab, unsigned long long seed, unsigned long long* state, int num_sms, cudaStream_t st);
cudaError_t wf_launch_bn_fwd_fin(const BnFwdFin* d, int n, cudaStream_t st);
cudaError_t wf_launch_bn_bwd_fin(const BnBwdFin* d, int n, cudaStream_t st);
cudaError_t wf_launch_bn_eval_coefs(const BnEvalTable& tab, const float* params, const float* running, float* coefs, cudaStream_t st);
cudaError_t wf_launch_join_fwd(const JoinP& p, int num_sms, cudaStream_t st);
cudaError_t wf_launch_join_bwd(const JoinP& p, int num_sms, cudaStream_t st);
cudaError_t wf_launch_bn_bwd_stats(const float* dy, const float* raw, const float* mean, int C, long long plane, double* s0, double* s1, int num_sms, cudaStream_t st);
cudaError_t wf_launch_pool_fwd(const float* raw, const float* scale, const float* shift, const float* mean, float* pred, int B, cudaStream_t st);
cudaError_t wf_launch_pool_bwd(const float* raw, const float* scale, const float* shift, const float* mean, const float* dpred, float* dy, int B,
                               double* s0, double* s1, cudaStream_t st);
cudaError_t wf_launch_pose_loss(const float* pred, const float* target, int B, int type, float pw, float bw, const float* gscale,
                                float* dpred, double* acc2, float* out3, cudaStream_t st);
cudaError_t wf_launch_metrics(const float* pred, const float* target, int B, const MetricThr& thr, int torso, unsigned long long* counts,
                              double* dsum, float* out, cudaStream_t st);
cudaError_t wf_launch_pack(const PackTable& tab, const float* params, float* packed, cudaStream_t st);
cudaError_t wf_launch_adamw(float* p, const float* g, float* m, float* v, long long n, AdamState* state, float lr, float b1, float b2,
                            float eps, float wd, float max_norm, float grad_scale, int num_sms, cudaStream_t st);
cudaError_t wf_launch_permute(const float* src, float* dst, int C, int P, int B, long long r_sb, long long r_sc, long long r_sp,
                              long long r_st, int to_internal, int num_sms, cudaStream_t st);
cudaError_t wf_launch_permute_affine(const float* src, float* dst, int C, int P, int B, long long r_sb, long long r_sc, long long r_sp,
                                     long long r_st, const float* scale, const float* shift, const float* mean, int num_sms, cudaStream_t st);
cudaError_t wf_launch_attn_fwd_stats(const AttnP& p, cudaStream_t st);
cudaError_t wf_launch_attn_fwd(const AttnP& p, cudaStream_t st);
cudaError_t wf_launch_attn_bwd_stats(const AttnP& p, cudaStream_t st);
cudaError_t wf_launch_attn_bwd(const AttnP& p, cudaStream_t st);
// input side (wf_data.cu)
cudaError_t wf_launch_window_load(const float* src, const long long* idx, long long n_src, float* dst, int B, int C, int T, int sc, int st,
                                  const int* spans, double* stats, cudaStream_t stream);
cudaError_t wf_launch_noise_scale(const float* x, const float* noise, float* y, long long n, float level, float scale, const double* stats,
                                  long long n_stat, int num_sms, cudaStream_t stream);
cudaError_t wf_launch_keypoint_repair(const float* frames, long long n_frames, const long long* idx, float* out, int B, int K, int clean,
                                      cudaStream_t stream);
cudaError_t wf_launch_keypoint_seq(float* frames, const long long* seq_off, int n_seq, int K, cudaStream_t stream);
