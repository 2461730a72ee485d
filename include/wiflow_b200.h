/* wiflow_b200 -- C ABI of the B200-native WiFlow hot path (libwiflow_b200.so).
 *
 * The reference (DY2434/WiFlow-...) is pure Python: its "operator interface" for this path is the nn.Module /
 * function API that train.py consumes.  Each entry point below names the reference interface it replaces; the
 * Python package in this repo binds them with ctypes (see INTEGRATION.md for the stub a reference maintainer adds).
 *
 * Conventions
 *  - every pointer is a DEVICE pointer to fp32 data unless stated otherwise; all buffers are owned by the caller;
 *    the library allocates no device memory and keeps no state that outlives a call except per-THREAD bookkeeping
 *    (last error string, profiler records, the programmatic-launch switch, one side stream + events per host thread)
 *    and a process-wide launch counter, so it is re-entrant across streams and threads;
 *  - work is only enqueued on `stream`; no call synchronises, so every call is CUDA-graph capturable;
 *  - return value: 0 ok, <0 argument error (WF_E_*), >0 a cudaError_t; wf_last_error_string() describes the last
 *    failure on the calling thread.  There is no CPU fallback: without an sm_100a device the calls fail.
 *  - parameters are ONE flat fp32 buffer in the reference's state_dict()/named_parameters() order
 *    (2 225 042 floats for the full model); gradients use the same layout;
 *    BatchNorm running statistics are one flat buffer [running_mean(C), running_var(C)] per BatchNorm in module
 *    order, num_batches_tracked one int64 per BatchNorm.  wf_*_param_count / wf_param_table describe the layout.
 */
#ifndef WIFLOW_B200_H
#define WIFLOW_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* wf_stream_t;              /* cudaStream_t */
#if defined(__GNUC__)
#define WF_API __attribute__((visibility("default")))
#else
#define WF_API
#endif

enum { WF_E_ARG = -1, WF_E_WORKSPACE = -2, WF_E_ARCH = -3, WF_E_UNSUPPORTED = -4 };
enum { WF_FLAG_TRAIN = 1,               /* BatchNorm uses batch statistics and updates the running buffers */
       WF_FLAG_SAVE_FOR_BACKWARD = 2,   /* keep activations in the workspace for wf_*_backward */
       WF_FLAG_PROFILE = 4 };           /* record CUDA events around every kernel launch (wf_profile_*); not graph-capturable */
/* block ids for the per-block entry points (sub-module drop-ins) */
enum { WF_BLOCK_MODEL = 0,              /* models/pose_model.py:9  WiFlowPoseModel            [B,540,20] -> [B,15,2]      */
       WF_BLOCK_TCN = 1,                /* models/tcn.py:76        TemporalBlock              [B,540,20] -> [B,240,20]    */
       WF_BLOCK_CONVBLOCK1 = 2,         /* models/convnet.py:41    ConvBlock1(cin,cout)       [B,cin,20,W] -> [B,cout,20,W]   */
       WF_BLOCK_ASYMCONV = 3,           /* models/convnet.py:4     AsymmetricConvBlock        [B,cin,20,W] -> [B,cout,20,W/2] */
       WF_BLOCK_AXIAL_W = 4,            /* models/attention.py:7   AxialAttention(width=True) [B,64,15,20]                 */
       WF_BLOCK_AXIAL_H = 5,            /* models/attention.py:7   AxialAttention(width=False)                             */
       WF_BLOCK_DUAL_AXIAL = 6,         /* models/attention.py:83  DualAxialAttention                                      */
       WF_BLOCK_INNER_TCN = 7 };        /* models/tcn.py:14        InnerGroupedTemporalBlock(cin,cout,dilation)            */
enum { WF_LOSS_TYPE_SMOOTH_L1 = 0, WF_LOSS_TYPE_MSE = 1, WF_LOSS_TYPE_L1 = 2 };

/* Block geometry for the per-block entry points (ignored fields may be 0). */
typedef struct wf_block_desc {
    int block;          /* WF_BLOCK_* */
    int cin, cout;      /* conv blocks / inner TCN block */
    int width;          /* conv blocks: input feature-axis length W (240, 120, ...) */
    int dilation;       /* inner TCN block */
} wf_block_desc;

WF_API const char* wf_last_error_string(void);
WF_API int wf_version(void);

/* ---- layout queries (host only) ---- */
WF_API long long wf_param_count(const wf_block_desc* d);     /* floats in the flat parameter buffer            */
WF_API long long wf_running_count(const wf_block_desc* d);   /* floats in the flat running-statistics buffer   */
WF_API int wf_bn_count(const wf_block_desc* d);              /* number of BatchNorm layers (num_batches_tracked entries) */
WF_API int wf_dropout_sites(const wf_block_desc* d);         /* number of dropout masks the block consumes in train mode */
/* i-th parameter tensor in state_dict order: name (relative to the block), offset and element count. returns 0, or WF_E_ARG past the end */
WF_API int wf_param_table(const wf_block_desc* d, int i, char* name, int name_cap, long long* offset, long long* numel);
WF_API size_t wf_workspace_bytes(const wf_block_desc* d, int B, int flags);

/* ---- forward / backward of a block (replaces nn.Module.forward + autograd of the reference classes above) ----
 * x, y, dy, dx use the reference's tensor layouts (contiguous).  masks: NULL (no dropout) or wf_dropout_sites()
 * device pointers to multiplicative masks (0 or 1/(1-p)) in the reference's call order: nn.Dropout sites are
 * [B,C,20] tensors (tcn.py:30,43), nn.Dropout2d sites are [B,C] (convnet.py:15,20,51,56).
 * wf_block_backward needs the workspace of the matching forward (WF_FLAG_TRAIN|WF_FLAG_SAVE_FOR_BACKWARD) untouched,
 * writes d loss/d params into `grads` (overwritten, same flat layout) and, if dx != NULL, d loss/d x. */
WF_API int wf_block_forward(const wf_block_desc* d, const float* x, const float* params, float* running, long long* num_batches_tracked,
                     const float* const* masks, float* y, void* workspace, size_t workspace_bytes, int B, int flags, wf_stream_t stream);
WF_API int wf_block_backward(const wf_block_desc* d, const float* x, const float* params, const float* const* masks, const float* dy,
                      float* grads, float* dx, void* workspace, size_t workspace_bytes, int B, int flags, wf_stream_t stream);

/* ---- losses/pose_loss.py:35-88 PoseLoss.forward (+ its autograd) ----
 * out3 = {total, position, bone}; dpred (may be NULL) = gscale * d total / d pred, gscale a device scalar or NULL (=1).
 * scratch: 2 doubles, zero on first use (re-zeroed by the call). */
WF_API int wf_pose_loss(const float* pred, const float* target, int B, int loss_type, float position_weight, float bone_weight,
                 const float* gscale, float* dpred, float* out3, double* scratch, wf_stream_t stream);

/* ---- utils/metrics.py:3-47 calculate_pck + calculate_mpjpe ----
 * out = {pck[0..nthr-1], mpjpe}; thresholds are HOST floats (nthr <= 8).
 * scratch: 16 eight-byte words, zero on first use (re-zeroed by the call). */
WF_API int wf_pose_metrics(const float* pred, const float* target, int B, const float* thresholds, int nthr, int use_torso_norm,
                    float* out, void* scratch, wf_stream_t stream);

/* ---- train.py:105-110,235-236 clip_grad_norm_(max_norm) + torch.optim.AdamW.step over flat buffers ----
 * grads are multiplied by grad_scale first (1/world_size after a sum-allreduce).  state: WF_ADAM_STATE_BYTES bytes, zero before
 * the first step (holds the step counter and the per-block partial sums of the deterministic gradient-norm reduction: ranks that
 * hold the same all-reduced gradient compute bit-identical updates); after the call state[+16] (float) is the pre-clip norm. */
#define WF_ADAM_STATE_BYTES (64 + 8 * 256)
WF_API int wf_clip_adamw(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n, void* state,
                  float lr, float beta1, float beta2, float eps, float weight_decay, float max_norm, float grad_scale,
                  wf_stream_t stream);

/* ---- input side: dataset.py + utils/augmentation.py ----
 * wf_window_load: x[b] = windows[idx[b]] (idx NULL: identity; may run in place with x == windows), the gather that
 *   dataset.py:206 __getitem__ + the DataLoader collate do on the host, fused with utils/augmentation.py:3-19 time_masking
 *   as train.py:189 applies it: a window is C*T dense floats, t_major != 0 means stored [T][C] (the model's [540][20] input,
 *   i.e. time_masking called on x.permute(0,2,1)), else [C][T].  spans: NULL or B x {start0,len0,start1,len1} int32 (len 0 =
 *   unused) drawn by the host; every c-row's mean over T replaces [start,start+len), the second span after the first.
 *   stats: NULL or 2 doubles (zeroed by the caller) that receive sum(x), sum(x*x) of the OUTPUT, for wf_noise_scale.
 * wf_noise_scale: utils/augmentation.py:22-35 add_noise + random_scaling, y = (x + (noise*level)*std(x)) * scale with
 *   std = unbiased standard deviation from `stats` over n_stat elements; noise NULL: scaling only.  In place allowed.
 * wf_keypoint_batch: dataset.py:80-120 _get_keypoint_npy + _clean_single_frame_zeros: y[b] = frames[idx[b]] ([K][2]),
 *   zeros when idx[b] is outside [0,n_frames); clean != 0: all-zero joints take the mean of the frame's other joints.
 * wf_keypoint_sequences: dataset.py:159-206 _clean_zero_keypoints over sequences seq_off[s]..seq_off[s+1] of frames, in place. */
WF_API int wf_window_load(const float* windows, long long n_windows, const long long* idx, float* x, int B, int C, int T, int t_major,
                   const int* spans, double* stats, wf_stream_t stream);
WF_API int wf_noise_scale(const float* x, const float* noise, float* y, long long n, float noise_level, float scale, const double* stats,
                   long long n_stat, wf_stream_t stream);
WF_API int wf_keypoint_batch(const float* frames, long long n_frames, const long long* idx, float* y, int B, int K, int clean,
                      wf_stream_t stream);
/* wf_dropout_masks: the dropout masks of one training step in ONE launch (perf mode of the nn.Dropout / nn.Dropout2d draws of
 *   models/tcn.py:30,43 and models/convnet.py:15,20; the parity mode passes masks drawn by torch's generator instead).
 *   out / numel / p: HOST arrays of n_sites (<= 32) device buffers, their element counts and drop probabilities;
 *   out[i][k] = u >= p[i] ? 1/(1-p[i]) : 0 with u = Philox4x32-10(counter = (k/4, draw), key = seed ^ f(i))[k%4] * 2^-32.
 *   state: 2 x uint64 on the DEVICE, zeroed once by the caller: [0] = draw counter, advanced by one per call ON THE DEVICE (so a
 *   captured CUDA graph draws fresh masks on every replay), [1] = scratch.  Same (seed, draw) -> same masks. */
WF_API int wf_dropout_masks(float* const* out, const long long* numel, const float* p, int n_sites, unsigned long long seed,
                     unsigned long long* state, wf_stream_t stream);
WF_API int wf_keypoint_sequences(float* frames, const long long* seq_off, int n_seq, int K, wf_stream_t stream);

/* ---- measurement hooks (bench.py) ----
 * wf_launch_count: kernels launched by the library since load.  wf_profile_*: per-launch CUDA-event timings of the calls this
 * thread made with WF_FLAG_PROFILE (name = kernel family + layer), read after the work has been enqueued. */
WF_API long long wf_launch_count(void);
WF_API int wf_profile_count(void);
WF_API int wf_profile_read(int i, char* name, int name_cap, float* ms, double* flops);
WF_API int wf_profile_bytes(int i, double* bytes);   /* algorithmic HBM bytes of launch i (every operand tensor once) */
WF_API void wf_profile_reset(void);

/* ---- test / debug introspection (not part of the reference-facing surface) ----
 * i-th named fp32 workspace tensor ([C][P][B*20], n = b*20+t contiguous) of the block's layout: byte offset into the workspace. */
WF_API int wf_debug_tensor(const wf_block_desc* d, int B, int flags, int i, char* name, int name_cap, long long* byte_offset, int* C, int* P);

#ifdef __cplusplus
}
#endif
#endif
