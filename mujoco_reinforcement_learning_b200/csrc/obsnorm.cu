// Per-segment observation normalisation + [N, obs, W] -> [N, W, obs] permute + cast to fp32 (one pass).
//
// replaces: EnvironmentHelper.normalize_state / _normalize / get_state,
//           src/environments/humanoid/running_gym_sequential_vectorized.py:61-92 — for every env and every frame of
//           the window, each index range of the observation vector (positions, velocities, inertias, ...) is centred
//           and divided by its unbiased std (std == 0 -> 1), in the input's precision (gym observations are float64),
//           then the tensor is cast to fp32 and permuted for the policy.
// One warp per (env, segment): the segment's rows obs[n, b:e, 0:W] are one contiguous chunk (coalesced reads), the W
// per-frame sums live in registers, outputs are written frame by frame so consecutive lanes hit consecutive addresses.
#include "common.cuh"

namespace b200ppo {

constexpr int kMaxWindow = 8;

template <typename T>
__global__ void __launch_bounds__(128)
normalize_obs_kernel(const T* __restrict__ obs, int64_t n_envs, int obs_dim, int window, const int32_t* __restrict__ seg_bounds,
                     int n_segments, int normalize, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t wid = int64_t(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (wid >= n_envs * n_segments) return;
  const int64_t n = wid / n_segments;
  const int s = int(wid - n * n_segments);
  const int b = seg_bounds[s], e = seg_bounds[s + 1], L = e - b;
  const T* src = obs + (n * obs_dim + b) * window;  // L x W contiguous
  const int total = L * window;
  T mean[kMaxWindow], inv[kMaxWindow];
#pragma unroll
  for (int w = 0; w < kMaxWindow; ++w) { mean[w] = T(0); inv[w] = T(1); }
  if (normalize) {
    T sum[kMaxWindow];
#pragma unroll
    for (int w = 0; w < kMaxWindow; ++w) sum[w] = T(0);
    for (int i = lane; i < total; i += 32) {
      const int w = i % window;
      const T v = src[i];
#pragma unroll
      for (int k = 0; k < kMaxWindow; ++k)
        if (k == w) sum[k] += v;
    }
#pragma unroll
    for (int w = 0; w < kMaxWindow; ++w) mean[w] = warp_sum(sum[w]) / T(L);
#pragma unroll
    for (int w = 0; w < kMaxWindow; ++w) sum[w] = T(0);
    for (int i = lane; i < total; i += 32) {
      const int w = i % window;
      const T v = src[i];
#pragma unroll
      for (int k = 0; k < kMaxWindow; ++k)
        if (k == w) { const T c = v - mean[k]; sum[k] += c * c; }
    }
#pragma unroll
    for (int w = 0; w < kMaxWindow; ++w) {
      const T sd = sqrt(warp_sum(sum[w]) / T(L - 1));  // unbiased, as torch.std
      inv[w] = (sd == T(0)) ? T(1) : sd;                // std[std == 0] = 1
    }
  }
  for (int w = 0; w < window; ++w) {
    T m = T(0), d = T(1);
#pragma unroll
    for (int k = 0; k < kMaxWindow; ++k)
      if (k == w) { m = mean[k]; d = inv[k]; }
    float* dst = out + (n * window + w) * obs_dim + b;
    for (int i = lane; i < L; i += 32) dst[i] = float((src[i * window + w] - m) / d);
  }
}

}  // namespace b200ppo

using namespace b200ppo;

extern "C" B2_EXPORT int b200ppo_normalize_obs(const void* obs, int obs_is_f64, int64_t n_envs, int32_t obs_dim, int32_t window,
                                               const int32_t* seg_bounds, int32_t n_segments, int normalize, float* out,
                                               b200ppo_stream stream) {
  B2_CHECK_ARG(n_envs >= 0 && obs_dim > 0 && window >= 1 && window <= kMaxWindow && n_segments >= 1,
               "b200ppo_normalize_obs: bad sizes (window <= %d)", kMaxWindow);
  if (n_envs == 0) return B200PPO_OK;
  B2_CHECK_ARG(obs && seg_bounds && out, "b200ppo_normalize_obs: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t warps = n_envs * n_segments;
  const unsigned grid = unsigned((warps + 3) / 4);
  if (obs_is_f64)
    normalize_obs_kernel<double><<<grid, 128, 0, st>>>(static_cast<const double*>(obs), n_envs, obs_dim, window, seg_bounds, n_segments, normalize, out);
  else
    normalize_obs_kernel<float><<<grid, 128, 0, st>>>(static_cast<const float*>(obs), n_envs, obs_dim, window, seg_bounds, n_segments, normalize, out);
  B2_LAUNCH_CHECK();
  return B200PPO_OK;
}
