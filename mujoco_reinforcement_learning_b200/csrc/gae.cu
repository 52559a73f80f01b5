// K1 — reverse-time GAE / TD(lambda) scan with fused per-env normalisation.
//
// replaces: PPO.calculate_advantages (src/entities/algorithms/ppo.py:62-91) and torchrl 0.6.0
//           generalized_advantage_estimate (call site ppo.py:76-80).
//
// Layout: [N_envs, T] row-major, time contiguous (the reference's env-major buffer, ppo.py:60).
// One warp owns one env row and walks it backwards in tiles of 32*VEC steps.  Inside a tile the
// recurrence A_t = delta_t + c_t * A_{t+1} is an affine map per step; each lane composes its VEC
// consecutive maps, a 5-step shuffle suffix-scan composes across lanes, and every lane then replays
// its own VEC steps in the reference's exact sequential form from its carry-in.  All global traffic
// is 128-bit, fully coalesced, streamed past L1; the next tile's loads are issued before the current
// tile's scan so HBM latency overlaps the shuffles.
// Algorithmic bytes: 22 per (env, step) = 3 x fp32 + 2 x bool in, 2 x fp32 out.
#include "common.cuh"

namespace b200ppo {

template <typename RewT, int VEC>
struct GaeTile {
  RewT r[VEC];
  float v[VEC], vn[VEC];
  uint8_t term[VEC], done[VEC];
};

template <typename RewT, int VEC>
__device__ __forceinline__ void gae_load_tile(GaeTile<RewT, VEC>& tl, const RewT* __restrict__ reward,
                                              const float* __restrict__ value, const float* __restrict__ next_value,
                                              const uint8_t* __restrict__ term, const uint8_t* __restrict__ done,
                                              int64_t row, int t, int T) {
  if (t >= T) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      tl.r[i] = RewT(0); tl.v[i] = 0.f; tl.vn[i] = 0.f; tl.term[i] = 0; tl.done[i] = 0;
    }
    return;
  }
  const int64_t off = row + t;
  if constexpr (VEC == 4) {
    if constexpr (sizeof(RewT) == 4) {
      float4 r4 = ldg_stream4(reinterpret_cast<const float*>(reward) + off);
      tl.r[0] = r4.x; tl.r[1] = r4.y; tl.r[2] = r4.z; tl.r[3] = r4.w;
    } else {
      const double2* p = reinterpret_cast<const double2*>(reward + off);
      double2 a = __ldg(p), b = __ldg(p + 1);
      tl.r[0] = a.x; tl.r[1] = a.y; tl.r[2] = b.x; tl.r[3] = b.y;
    }
    float4 v4 = ldg_stream4(value + off), n4 = ldg_stream4(next_value + off);
    tl.v[0] = v4.x; tl.v[1] = v4.y; tl.v[2] = v4.z; tl.v[3] = v4.w;
    tl.vn[0] = n4.x; tl.vn[1] = n4.y; tl.vn[2] = n4.z; tl.vn[3] = n4.w;
    uchar4 t4 = __ldg(reinterpret_cast<const uchar4*>(term + off));
    tl.term[0] = t4.x; tl.term[1] = t4.y; tl.term[2] = t4.z; tl.term[3] = t4.w;
    if (done != nullptr) {
      uchar4 d4 = __ldg(reinterpret_cast<const uchar4*>(done + off));
      tl.done[0] = d4.x; tl.done[1] = d4.y; tl.done[2] = d4.z; tl.done[3] = d4.w;
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) tl.done[i] = tl.term[i] | uint8_t(t + i == T - 1);  // ppo.py:72
    }
  } else {
    tl.r[0] = __ldg(reward + off);
    tl.v[0] = __ldg(value + off);
    tl.vn[0] = __ldg(next_value + off);
    tl.term[0] = __ldg(term + off);
    tl.done[0] = done != nullptr ? __ldg(done + off) : uint8_t(tl.term[0] | uint8_t(t == T - 1));
  }
}

// mean and unbiased std over one row, two passes (the row is L1/L2 resident on the second).
template <typename SrcT, typename AccT, int VEC>
__device__ __forceinline__ void row_mean_std(const SrcT* rowp, int T, int lane, AccT& mean, AccT& stdv) {
  AccT s = 0;
  for (int t = lane * VEC; t < T; t += 32 * VEC) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) s += AccT(rowp[t + i]);
  }
  mean = warp_sum(s) / AccT(T);
  AccT q = 0;
  for (int t = lane * VEC; t < T; t += 32 * VEC) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      AccT c = AccT(rowp[t + i]) - mean;
      q += c * c;
    }
  }
  stdv = sqrt(warp_sum(q) / AccT(T - 1));
}

template <typename RewT, typename AccT, int VEC, bool NORM_REW, bool NORM_ADV>
__global__ void __launch_bounds__(128)
gae_scan_kernel(const RewT* __restrict__ reward, const float* __restrict__ value,
                const float* __restrict__ next_value, const uint8_t* __restrict__ term,
                const uint8_t* __restrict__ done, int64_t n_envs, int T, float gamma, float disc, float scaler,
                float* __restrict__ adv, float* __restrict__ tgt, int row_in_smem) {
  // NORM_ADV with row_in_smem: the warp keeps its row of both outputs (2 x T floats) in shared memory until the
  // row statistics are known, so every output element is written to global memory exactly once (22 B per element
  // like the plain scan) instead of written, re-read twice and re-written.
  extern __shared__ __align__(16) float gae_rows[];
  float* row_a = gae_rows + size_t(threadIdx.x >> 5) * 2 * size_t((T + 3) & ~3);
  float* row_t = row_a + ((T + 3) & ~3);
  const int lane = threadIdx.x & 31;
  const int64_t env = int64_t(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (env >= n_envs) return;
  const int64_t row = env * int64_t(T);
  constexpr int TILE = 32 * VEC;
  const int n_tiles = (T + TILE - 1) / TILE;

  AccT r_mean = 0, r_std = 1;
  if constexpr (NORM_REW) row_mean_std<RewT, AccT, VEC>(reward + row, T, lane, r_mean, r_std);  // ppo.py:66-69

  AccT carry = 0;  // A_{t+1} entering the current tile ("prev = 0", torchrl)
  float sum_a = 0.f, sum_t = 0.f;
  GaeTile<RewT, VEC> cur, nxt;
  gae_load_tile<RewT, VEC>(cur, reward, value, next_value, term, done, row, (n_tiles - 1) * TILE + lane * VEC, T);
  for (int j = n_tiles - 1; j >= 0; --j) {
    const int t0 = j * TILE + lane * VEC;
    if (j > 0) gae_load_tile<RewT, VEC>(nxt, reward, value, next_value, term, done, row, t0 - TILE, T);
    const bool valid = t0 < T;
    AccT delta[VEC];
    AccT c[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      AccT r = AccT(cur.r[i]);
      if constexpr (NORM_REW) r = AccT(((r - r_mean) / r_std) * AccT(scaler));
      // delta = reward + (gamma * not_terminated) * V' - V ; separate roundings like the ATen ops
      const float gv = __fmul_rn(cur.term[i] ? 0.f : gamma, cur.vn[i]);
      if constexpr (sizeof(AccT) == 4) {
        delta[i] = __fsub_rn(__fadd_rn(r, gv), cur.v[i]);
      } else {
        delta[i] = (r + AccT(gv)) - AccT(cur.v[i]);
      }
      c[i] = valid ? AccT(cur.done[i] ? 0.f : disc) : AccT(1);
      if (!valid) delta[i] = 0;
    }
    // lane composite: A_first = D + C * A_in
    AccT C = c[VEC - 1], D = delta[VEC - 1];
#pragma unroll
    for (int i = VEC - 2; i >= 0; --i) {
      D = delta[i] + c[i] * D;
      C = c[i] * C;
    }
    // inclusive suffix scan over lanes (lane l <- lanes l..31)
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      AccT Cn = __shfl_down_sync(0xffffffffu, C, o);
      AccT Dn = __shfl_down_sync(0xffffffffu, D, o);
      if (lane + o < 32) {
        D = D + C * Dn;
        C = C * Cn;
      }
    }
    AccT Cn = __shfl_down_sync(0xffffffffu, C, 1);
    AccT Dn = __shfl_down_sync(0xffffffffu, D, 1);
    AccT x = (lane == 31) ? carry : (Dn + Cn * carry);  // A at the first step after this lane's span
    float a_out[VEC], t_out[VEC];
#pragma unroll
    for (int i = VEC - 1; i >= 0; --i) {
      if constexpr (sizeof(AccT) == 4) {
        x = __fadd_rn(delta[i], __fmul_rn(x, c[i]));  // prev = delta_t + prev * discount_t
      } else {
        x = delta[i] + x * c[i];
      }
      a_out[i] = float(x);
      t_out[i] = __fadd_rn(a_out[i], cur.v[i]);  // value_target = advantage + state_value
    }
    carry = __shfl_sync(0xffffffffu, x, 0);
    if (valid) {
      if (NORM_ADV && row_in_smem) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) { row_a[t0 + i] = a_out[i]; row_t[t0 + i] = t_out[i]; }
      } else if constexpr (VEC == 4) {
        stg_stream4(adv + row + t0, make_float4(a_out[0], a_out[1], a_out[2], a_out[3]));
        stg_stream4(tgt + row + t0, make_float4(t_out[0], t_out[1], t_out[2], t_out[3]));
      } else {
        adv[row + t0] = a_out[0];
        tgt[row + t0] = t_out[0];
      }
      if constexpr (NORM_ADV) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) { sum_a += a_out[i]; sum_t += t_out[i]; }
      }
    }
    if (j > 0) cur = nxt;
  }

  if constexpr (NORM_ADV) {  // ppo.py:81-88: both outputs, per env over time, unbiased std, no epsilon
    const float mean_a = warp_sum(sum_a) / float(T), mean_t = warp_sum(sum_t) / float(T);
    float qa = 0.f, qt = 0.f;
    if (row_in_smem) {  // every lane touches only the shared-memory elements it wrote itself
      for (int t = lane * VEC; t < T; t += TILE) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          float ca = row_a[t + i] - mean_a, ct = row_t[t + i] - mean_t;
          qa += ca * ca; qt += ct * ct;
        }
      }
      const float sa = sqrtf(warp_sum(qa) / float(T - 1)), stt = sqrtf(warp_sum(qt) / float(T - 1));
      for (int t = lane * VEC; t < T; t += TILE) {
        float a4[VEC], t4[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          a4[i] = __fmul_rn(__fdiv_rn(row_a[t + i] - mean_a, sa), scaler);
          t4[i] = __fmul_rn(__fdiv_rn(row_t[t + i] - mean_t, stt), scaler);
        }
        if constexpr (VEC == 4) {
          stg_stream4(adv + row + t, make_float4(a4[0], a4[1], a4[2], a4[3]));
          stg_stream4(tgt + row + t, make_float4(t4[0], t4[1], t4[2], t4[3]));
        } else {
          adv[row + t] = a4[0];
          tgt[row + t] = t4[0];
        }
      }
      return;
    }
    // every lane re-reads exactly the elements it wrote itself
    for (int t = lane * VEC; t < T; t += TILE) {
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        float ca = adv[row + t + i] - mean_a, ct = tgt[row + t + i] - mean_t;
        qa += ca * ca; qt += ct * ct;
      }
    }
    const float std_a = sqrtf(warp_sum(qa) / float(T - 1)), std_t = sqrtf(warp_sum(qt) / float(T - 1));
    for (int t = lane * VEC; t < T; t += TILE) {
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        adv[row + t + i] = __fmul_rn(__fdiv_rn(adv[row + t + i] - mean_a, std_a), scaler);
        tgt[row + t + i] = __fmul_rn(__fdiv_rn(tgt[row + t + i] - mean_t, std_t), scaler);
      }
    }
  }
}

template <typename RewT, typename AccT, int VEC>
static int launch_gae(const void* reward, const float* value, const float* next_value, const uint8_t* term,
                      const uint8_t* done, int64_t N, int T, float gamma, float disc, int nr, int na, float scaler,
                      float* adv, float* tgt, cudaStream_t st) {
  const int warps = 4;
  dim3 grid((unsigned)((N + warps - 1) / warps)), block(warps * 32);
  const RewT* r = static_cast<const RewT*>(reward);
  // rows of both outputs stay in shared memory while their statistics are computed when 4 warps x 2 x T floats fit
  const size_t row_bytes = size_t(warps) * 2 * size_t((T + 3) & ~3) * sizeof(float);
  const int in_smem = (na && row_bytes <= 96 * 1024) ? 1 : 0;
  const size_t smem = in_smem ? row_bytes : 0;
  if (smem > 48 * 1024) {
    if (nr) B2_CUDA(cudaFuncSetAttribute(gae_scan_kernel<RewT, AccT, VEC, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    else B2_CUDA(cudaFuncSetAttribute(gae_scan_kernel<RewT, AccT, VEC, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  }
#define GAE_GO(NR, NA) gae_scan_kernel<RewT, AccT, VEC, NR, NA><<<grid, block, smem, st>>>(r, value, next_value, term, done, N, T, gamma, disc, scaler, adv, tgt, in_smem)
  if (nr && na) GAE_GO(true, true);
  else if (nr) GAE_GO(true, false);
  else if (na) GAE_GO(false, true);
  else GAE_GO(false, false);
#undef GAE_GO
  B2_LAUNCH_CHECK();
  return B200PPO_OK;
}

}  // namespace b200ppo

using namespace b200ppo;

extern "C" B2_EXPORT int b200ppo_gae(const void* reward, int reward_is_f64, const float* value, const float* next_value,
                           const uint8_t* terminated, const uint8_t* done, int64_t n_envs, int64_t n_steps,
                           double gamma, double lmbda, int normalize_rewards, int normalize_advantage,
                           double advantage_scaler, float* advantage, float* value_target, b200ppo_stream stream) {
  B2_CHECK_ARG(n_envs >= 0 && n_steps >= 0 && n_steps < (1ll << 30), "b200ppo_gae: bad shape [%lld,%lld]",
               (long long)n_envs, (long long)n_steps);
  if (n_envs == 0 || n_steps == 0) return B200PPO_OK;
  B2_CHECK_ARG(reward && value && next_value && terminated && advantage && value_target, "b200ppo_gae: null pointer");
  B2_CHECK_ARG((n_envs + 3) / 4 < (1ll << 31), "b200ppo_gae: too many envs");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const float g = float(gamma);          // gamma * int tensor -> float32 tensor
  const float disc = float(lmbda * gamma);  // product in double first, then float32 (torchrl)
  const float sc = float(advantage_scaler);
  const int T = int(n_steps);
  const bool vec = (T % 4 == 0) && aligned16(value) && aligned16(next_value) && aligned16(advantage) &&
                   aligned16(value_target) && aligned16(reward) && ((uintptr_t)terminated % 4 == 0) &&
                   (done == nullptr || (uintptr_t)done % 4 == 0);
  if (reward_is_f64) {
    return vec ? launch_gae<double, double, 4>(reward, value, next_value, terminated, done, n_envs, T, g, disc, normalize_rewards, normalize_advantage, sc, advantage, value_target, st)
               : launch_gae<double, double, 1>(reward, value, next_value, terminated, done, n_envs, T, g, disc, normalize_rewards, normalize_advantage, sc, advantage, value_target, st);
  }
  return vec ? launch_gae<float, float, 4>(reward, value, next_value, terminated, done, n_envs, T, g, disc, normalize_rewards, normalize_advantage, sc, advantage, value_target, st)
             : launch_gae<float, float, 1>(reward, value, next_value, terminated, done, n_envs, T, g, disc, normalize_rewards, normalize_advantage, sc, advantage, value_target, st);
}
