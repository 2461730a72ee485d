// Input side of the hot path: what happens to a batch between the dataset arrays and
// WiFlowPoseModel.forward.  All of it is HBM-bound copy/elementwise work over 43 200-byte CSI windows:
//
//   window_load_kernel      gather of B windows out of the resident `csi_windows` array (dataset.py:206, DataLoader collate) fused with
//                           utils/augmentation.py:3-19 time_masking (as train.py:189 calls it on x.permute(0,2,1)) and with the
//                           sum / sum-of-squares that add_noise's torch.std(x) needs (augmentation.py:25)
//   noise_scale_kernel      utils/augmentation.py:22-35 add_noise + random_scaling: y = (x + noise*level*std(x)) * scale
//   keypoint_repair_kernel  dataset.py:80-120 _get_keypoint_npy + _clean_single_frame_zeros for a batch of window indices
//   keypoint_seq_kernel     dataset.py:159-206 _clean_zero_keypoints (CSV mode: zero keypoints interpolated along a file's frames)
//
// Random decisions (which windows are masked, span start/length, the scale factor) are drawn by the host in the reference's
// call order on the same torch generators; the kernels only consume them, so a seeded run reproduces the reference's batches.
#include <cuda_runtime.h>
#include <stdint.h>

#include "wf_common.cuh"

namespace {

constexpr int DATA_THREADS = 256;

// block-wide sum of two doubles into two global doubles
__device__ __forceinline__ void block_accum2_d(double a, double b, double* out)
{
    __shared__ double red[2][DATA_THREADS / 32];
    a = warp_sum_d(a); b = warp_sum_d(b);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { red[0][w] = a; red[1][w] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s0 = 0, s1 = 0;
#pragma unroll
        for (int i = 0; i < DATA_THREADS / 32; ++i) { s0 += red[0][i]; s1 += red[1][i]; }
        atomicAdd(out, s0);
        atomicAdd(out + 1, s1);
    }
}

// One CTA per window.  A window is W = C*T dense floats; element (c, t) sits at c*sc + t*st (train.py:189 layout: the window
// is [T=540][C=20], sc = 1, st = 20; a contiguous [C][T] window: sc = T, st = 1).  spans[b] = {start0, len0, start1, len1}
// (len 0 = no mask): for every c the mean over t of the CURRENT row is written into [start, start+len) -- the second span
// sees the first one's result, as in the reference's nested loop (augmentation.py:13-17).
__global__ void __launch_bounds__(DATA_THREADS)
window_load_kernel(const float* src, const long long* __restrict__ idx, long long n_src, float* dst,
                   int W, int C, int T, int sc, int st, const int* __restrict__ spans, double* __restrict__ stats)
{
    extern __shared__ float4 win4[];
    float* win = reinterpret_cast<float*>(win4);
    const int b = blockIdx.x, tid = threadIdx.x;
    long long s = idx ? idx[b] : (long long)b;
    const bool have = s >= 0 && s < n_src;
    const float4* sp = reinterpret_cast<const float4*>(src + (have ? s : 0) * (long long)W);
    float4* dp = reinterpret_cast<float4*>(dst + (long long)b * W);
    const int Q = W / 4;
    int m[4] = {0, 0, 0, 0};
    if (spans) { m[0] = spans[b * 4]; m[1] = spans[b * 4 + 1]; m[2] = spans[b * 4 + 2]; m[3] = spans[b * 4 + 3]; }
    double s0 = 0, s1 = 0;
    if (m[1] <= 0 && m[3] <= 0) {
        // plain gather: registers only, four 128-bit requests per thread in flight
        constexpr int U = 4;
        const bool inplace = static_cast<const float4*>(dp) == sp;           // in place and unmasked: nothing to write
        for (int q0 = tid; q0 < Q; q0 += U * DATA_THREADS) {
            float4 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int q = q0 + u * DATA_THREADS;
                v[u] = f4zero();
                if (have && q < Q) v[u] = sp[q];
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int q = q0 + u * DATA_THREADS;
                if (q < Q) {
                    if (!inplace) dp[q] = v[u];
                    if (stats) {
                        float a = (v[u].x + v[u].y) + (v[u].z + v[u].w);
                        float c2 = fmaf(v[u].x, v[u].x, fmaf(v[u].y, v[u].y, fmaf(v[u].z, v[u].z, v[u].w * v[u].w)));
                        s0 += a; s1 += c2;
                    }
                }
            }
        }
        if (4 * Q + tid < W) {           // scalar tail (W not a multiple of 4: single-window calls only)
            const float* s1p = reinterpret_cast<const float*>(sp);
            float* d1p = reinterpret_cast<float*>(dp);
            const float v = have ? s1p[4 * Q + tid] : 0.f;
            if (static_cast<const float*>(d1p) != s1p) d1p[4 * Q + tid] = v;
            if (stats) { s0 += v; s1 += v * v; }
        }
    } else {
        for (int q0 = tid; q0 < Q; q0 += 4 * DATA_THREADS) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { const int q = q0 + u * DATA_THREADS; v[u] = f4zero(); if (have && q < Q) v[u] = sp[q]; }
#pragma unroll
            for (int u = 0; u < 4; ++u) { const int q = q0 + u * DATA_THREADS; if (q < Q) win4[q] = v[u]; }
        }
        if (4 * Q + tid < W) win[4 * Q + tid] = have ? reinterpret_cast<const float*>(sp)[4 * Q + tid] : 0.f;
        __syncthreads();
        const int warp = tid >> 5, lane = tid & 31;
        for (int k = 0; k < 2; ++k) {
            const int start = m[2 * k], len = m[2 * k + 1];
            if (len <= 0) continue;
            for (int c = warp; c < C; c += DATA_THREADS / 32) {
                float* row = win + (long long)c * sc;
                float acc = 0.f;
                for (int t = lane; t < T; t += 32) acc += row[(long long)t * st];
                const float mean = (float)(warp_sum_d((double)acc) / (double)T);
                const int e = min(start + len, T);
                for (int t = max(start, 0) + lane; t < e; t += 32) row[(long long)t * st] = mean;
            }
            __syncthreads();
        }
        for (int q = tid; q < Q; q += DATA_THREADS) {
            const float4 v = win4[q];
            dp[q] = v;
            if (stats) {
                float a = (v.x + v.y) + (v.z + v.w);
                float c2 = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, v.w * v.w)));
                s0 += a; s1 += c2;
            }
        }
        if (4 * Q + tid < W) {
            const float v = win[4 * Q + tid];
            reinterpret_cast<float*>(dp)[4 * Q + tid] = v;
            if (stats) { s0 += v; s1 += v * v; }
        }
    }
    if (stats) block_accum2_d(s0, s1, stats);
}

// y = (x + (noise*level)*std) * scale, each product rounded to fp32 like the reference's tensor expression
// (augmentation.py:25-26, :33); std = unbiased standard deviation of all n elements from the fp64 sums.
__global__ void __launch_bounds__(DATA_THREADS)
noise_scale_kernel(const float* x, const float* __restrict__ noise, float* y, long long n, float level,
                   float scale, const double* __restrict__ stats, long long n_stat)
{
    float sd = 0.f;
    if (noise) {
        const double s0 = stats[0], s1 = stats[1];
        double var = (s1 - s0 * s0 / (double)n_stat) / (double)(n_stat - 1);
        sd = (float)sqrt(var > 0 ? var : 0.0);
    }
    const long long Q = n / 4;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    const float4* n4 = reinterpret_cast<const float4*>(noise);
    float4* y4 = reinterpret_cast<float4*>(y);
    for (long long q = (long long)blockIdx.x * DATA_THREADS + threadIdx.x; q < Q; q += (long long)gridDim.x * DATA_THREADS) {
        float4 v = x4[q];
        if (noise) {
            const float4 z = __ldg(n4 + q);
            v.x = __fadd_rn(v.x, __fmul_rn(__fmul_rn(z.x, level), sd));
            v.y = __fadd_rn(v.y, __fmul_rn(__fmul_rn(z.y, level), sd));
            v.z = __fadd_rn(v.z, __fmul_rn(__fmul_rn(z.z, level), sd));
            v.w = __fadd_rn(v.w, __fmul_rn(__fmul_rn(z.w, level), sd));
        }
        v.x = __fmul_rn(v.x, scale); v.y = __fmul_rn(v.y, scale); v.z = __fmul_rn(v.z, scale); v.w = __fmul_rn(v.w, scale);
        y4[q] = v;
    }
    // tail (n not a multiple of 4)
    for (long long i = Q * 4 + (long long)blockIdx.x * DATA_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * DATA_THREADS) {
        float v = x[i];
        if (noise) v = __fadd_rn(v, __fmul_rn(__fmul_rn(noise[i], level), sd));
        y[i] = __fmul_rn(v, scale);
    }
}

// One thread per window of the batch: frame = frames[idx[b]] ([K][2] floats), zeros if the index is out of range
// (dataset.py:101).  clean: joints with x == 0 and y == 0 take the mean of the others, summed in joint order in fp32 and
// divided by their count (numpy's axis-0 mean of a [k,2] float32 array: sequential adds, one division) -- bit exact.
__global__ void __launch_bounds__(128)
keypoint_repair_kernel(const float* __restrict__ frames, long long n_frames, const long long* __restrict__ idx, float* __restrict__ out,
                       int B, int K, int clean)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const long long f = idx ? idx[b] : (long long)b;
    float2* o = reinterpret_cast<float2*>(out) + (long long)b * K;
    if (f < 0 || f >= n_frames) {
        for (int j = 0; j < K; ++j) o[j] = make_float2(0.f, 0.f);
        return;
    }
    const float2* p = reinterpret_cast<const float2*>(frames) + f * K;
    float sx = 0.f, sy = 0.f;
    int cnt = 0;
    for (int j = 0; j < K; ++j) {
        const float2 v = __ldg(p + j);
        if (v.x != 0.f || v.y != 0.f) { sx = __fadd_rn(sx, v.x); sy = __fadd_rn(sy, v.y); ++cnt; }
    }
    const bool fix = clean && cnt > 0 && cnt < K;
    const float mx = fix ? __fdiv_rn(sx, (float)cnt) : 0.f, my = fix ? __fdiv_rn(sy, (float)cnt) : 0.f;
    for (int j = 0; j < K; ++j) {
        float2 v = __ldg(p + j);
        if (fix && v.x == 0.f && v.y == 0.f) v = make_float2(mx, my);
        o[j] = v;
    }
}

// dataset.py:159-206: one thread per (sequence, joint) walks the frames in order.  A zero joint (x == 0 and y == 0 in the
// ORIGINAL sequence) is rebuilt from the nearest non-zero frame before it -- in the already repaired sequence, as the
// reference's in-place loop sees it -- and the nearest non-zero frame after it: (1-a)*prev + a*next with
// a = (t-prev)/(next-prev) computed in fp64 and rounded to fp32 before the products (numpy's weak python-float promotion),
// or a copy of the only neighbour that exists.  seq_off[s]..seq_off[s+1] delimit sequence s in frames [F][K][2]; in place.
__global__ void __launch_bounds__(128)
keypoint_seq_kernel(float* __restrict__ frames, const long long* __restrict__ seq_off, int n_seq, int K)
{
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n_seq * K) return;
    const int s = id / K, j = id % K;
    const long long f0 = seq_off[s], f1 = seq_off[s + 1];
    float2* p = reinterpret_cast<float2*>(frames) + j;
    long long prev = -1, next = -1;          // absolute frame numbers; next is searched lazily and reused along a run of zeros
    for (long long t = f0; t < f1; ++t) {
        const float2 v = p[t * K];
        if (v.x != 0.f || v.y != 0.f) { prev = t; continue; }
        if (next <= t) {
            next = -1;
            for (long long u = t + 1; u < f1; ++u) {
                const float2 w = p[u * K];
                if (w.x != 0.f || w.y != 0.f) { next = u; break; }
            }
            if (next < 0) next = f1;            // sentinel: nothing valid after t
        }
        const bool hp = prev >= 0, hn = next < f1;
        float2 r = v;
        if (hp && hn) {
            const double a = (double)(t - prev) / (double)(next - prev);
            const float wa = (float)a, wb = (float)(1.0 - a);
            const float2 pv = p[prev * K], nv = p[next * K];
            r.x = __fadd_rn(__fmul_rn(wb, pv.x), __fmul_rn(wa, nv.x));
            r.y = __fadd_rn(__fmul_rn(wb, pv.y), __fmul_rn(wa, nv.y));
        } else if (hp) r = p[prev * K];
        else if (hn) r = p[next * K];
        p[t * K] = r;
        if (r.x != 0.f || r.y != 0.f) prev = t;   // the repaired frame is what the next zero frame finds when it looks back
    }
}

}  // namespace

cudaError_t wf_launch_window_load(const float* src, const long long* idx, long long n_src, float* dst, int B, int C, int T, int sc, int st,
                                  const int* spans, double* stats, cudaStream_t stream)
{
    const int W = C * T;
    const size_t smem = spans ? (size_t)((W + 3) / 4 * 4) * sizeof(float) : 0;     // only masked windows are staged in shared memory
    static WfSmemOptIn optin;
    if (smem > 48 * 1024)
        if (cudaError_t e = wf_smem_optin(optin, window_load_kernel, smem)) return e;
    window_load_kernel<<<B, DATA_THREADS, smem, stream>>>(src, idx, n_src, dst, W, C, T, sc, st, spans, stats);
    return cudaGetLastError();
}

cudaError_t wf_launch_noise_scale(const float* x, const float* noise, float* y, long long n, float level, float scale, const double* stats,
                                  long long n_stat, int num_sms, cudaStream_t stream)
{
    long long blocks = (n / 4 + DATA_THREADS - 1) / DATA_THREADS;
    const long long cap = (long long)num_sms * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    noise_scale_kernel<<<(int)blocks, DATA_THREADS, 0, stream>>>(x, noise, y, n, level, scale, stats, n_stat);
    return cudaGetLastError();
}

cudaError_t wf_launch_keypoint_repair(const float* frames, long long n_frames, const long long* idx, float* out, int B, int K, int clean,
                                      cudaStream_t stream)
{
    keypoint_repair_kernel<<<(B + 127) / 128, 128, 0, stream>>>(frames, n_frames, idx, out, B, K, clean);
    return cudaGetLastError();
}

cudaError_t wf_launch_keypoint_seq(float* frames, const long long* seq_off, int n_seq, int K, cudaStream_t stream)
{
    const int n = n_seq * K;
    keypoint_seq_kernel<<<(n + 127) / 128, 128, 0, stream>>>(frames, seq_off, n_seq, K);
    return cudaGetLastError();
}
