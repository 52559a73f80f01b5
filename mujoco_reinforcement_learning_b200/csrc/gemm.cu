// Grouped fp32 SIMT GEMM (FFMA) — the 1e-5-parity arithmetic of the actor/critic MLP.
//
// replaces: aten::addmm / aten::mm reached from NetworkBlock.forward
//           (src/models/network_block_creator.py:74-86) and from autograd of ppo.py:121,134.
//
// One launch covers a *group* of independent problems (actor and critic layers of the same depth, or every
// weight gradient of a minibatch) so a minibatch needs a handful of launches instead of one per matrix.
// Per CTA: BM x BN output tile, BK = 16 k-slab, double-buffered shared memory with register prefetch of the
// next slab, TM x TN register micro-tile per thread split in 4-wide groups so shared-memory reads are
// conflict-free 128-bit broadcasts.  Operands may be K-contiguous or M/N-contiguous (see gemm.cuh), which
// covers forward, dgrad and wgrad without materialising a transpose.  Split-K writes partials that the
// fused Adam kernel sums in a fixed order (deterministic).
#include "gemm.cuh"

namespace b200ppo {

constexpr int BK = 16;

template <int BMN, int NT>
struct TileLoader {
  // BMN x BK tile, NT threads; LV 128-bit slots per thread.
  static constexpr int LV = (BMN * BK / 4) / NT;
  static_assert(LV >= 1 && (BMN * BK / 4) % NT == 0, "tile/thread mismatch");

  // global -> registers. sm = stride of the M/N index, sk = stride of k (one of them is 1).
  __device__ __forceinline__ static void load(float4 (&r)[LV], const float* __restrict__ base, int64_t sm, int64_t sk,
                                              int mn0, int MN, int k0, int kend, bool vec, int tid) {
    if (sk == 1) {  // K-contiguous: slot -> (row, 4 consecutive k)
#pragma unroll
      for (int i = 0; i < LV; ++i) {
        const int s = tid + i * NT;
        const int row = s / (BK / 4), kv = (s % (BK / 4)) * 4;
        const int m = mn0 + row, k = k0 + kv;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m < MN) {
          const float* p = base + int64_t(m) * sm + k;
          if (vec) {
            if (k < kend) v = __ldg(reinterpret_cast<const float4*>(p));
          } else {
            if (k + 0 < kend) v.x = __ldg(p + 0);
            if (k + 1 < kend) v.y = __ldg(p + 1);
            if (k + 2 < kend) v.z = __ldg(p + 2);
            if (k + 3 < kend) v.w = __ldg(p + 3);
          }
        }
        r[i] = v;
      }
    } else {  // M/N-contiguous: slot -> (k, 4 consecutive m)
#pragma unroll
      for (int i = 0; i < LV; ++i) {
        const int s = tid + i * NT;
        const int kk = s / (BMN / 4), mv = (s % (BMN / 4)) * 4;
        const int m = mn0 + mv, k = k0 + kk;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < kend) {
          const float* p = base + int64_t(k) * sk + m;
          if (vec) {
            if (m < MN) v = __ldg(reinterpret_cast<const float4*>(p));
          } else {
            if (m + 0 < MN) v.x = __ldg(p + 0);
            if (m + 1 < MN) v.y = __ldg(p + 1);
            if (m + 2 < MN) v.z = __ldg(p + 2);
            if (m + 3 < MN) v.w = __ldg(p + 3);
          }
        }
        r[i] = v;
      }
    }
  }

  // registers -> shared tile laid out [BK][BMN + 4]
  __device__ __forceinline__ static void store(const float4 (&r)[LV], float (*tile)[BMN + 4], bool k_contig, int tid) {
    if (k_contig) {
#pragma unroll
      for (int i = 0; i < LV; ++i) {
        const int s = tid + i * NT;
        const int row = s / (BK / 4), kv = (s % (BK / 4)) * 4;
        tile[kv + 0][row] = r[i].x;
        tile[kv + 1][row] = r[i].y;
        tile[kv + 2][row] = r[i].z;
        tile[kv + 3][row] = r[i].w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < LV; ++i) {
        const int s = tid + i * NT;
        const int kk = s / (BMN / 4), mv = (s % (BMN / 4)) * 4;
        *reinterpret_cast<float4*>(&tile[kk][mv]) = r[i];
      }
    }
  }
};

__device__ __forceinline__ float apply_epilogue(float acc, int epi, float bias, float aux, float scale) {
  switch (epi) {
    case EPI_BIAS: return acc + bias;
    case EPI_BIAS_TANH: return tanhf(acc + bias);
    case EPI_BIAS_RELU: return fmaxf(acc + bias, 0.f);
    case EPI_BIAS_TANH_SCALE: return scale * tanhf(acc + bias);
    case EPI_DTANH: return acc * (1.f - aux * aux);
    case EPI_DRELU: return aux > 0.f ? acc : 0.f;
    default: return acc;
  }
}

template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN), 2)  // two CTAs per SM: the second hides the first's load latency
gemm_group_kernel(const __grid_constant__ GemmGroup grp) {
  constexpr int NT = (BM / TM) * (BN / TN);
  constexpr int RG = TM / 4, CG = TN / 4;  // 4-wide row / column groups per thread
  using LoadA = TileLoader<BM, NT>;
  using LoadB = TileLoader<BN, NT>;
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];

  // which problem / tile / split am I?
  int pi = 0;
#pragma unroll 1
  for (int i = 1; i < grp.count; ++i)
    if (int(blockIdx.x) >= grp.p[i].tile_begin) pi = i;
  const GemmProblem& P = grp.p[pi];
  int local = int(blockIdx.x) - P.tile_begin;
  const int tiles_mn = P.tiles_m * P.tiles_n;
  const int split = local / tiles_mn;
  local -= split * tiles_mn;
  const int m0 = (local / P.tiles_n) * BM, n0 = (local % P.tiles_n) * BN;
  const int kbeg = split * P.k_per_split;
  const int kend = min(P.K, kbeg + P.k_per_split);

  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  const bool a_kc = (P.a_sk == 1), b_kc = (P.b_sk == 1);
  const bool a_vec = P.a_vec != 0, b_vec = P.b_vec != 0;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  float bsum = 0.f;  // bias-gradient partial: sum_k A(m0 + tid, k)
  const bool do_bsum = (P.bias_grad != nullptr) && (n0 == 0);

  float4 ra[LoadA::LV], rb[LoadB::LV];
  const int nk = (kend - kbeg + BK - 1) / BK;
  if (nk > 0) {
    LoadA::load(ra, P.A, P.a_sm, P.a_sk, m0, P.M, kbeg, kend, a_vec, tid);
    LoadB::load(rb, P.B, P.b_sn, P.b_sk, n0, P.N, kbeg, kend, b_vec, tid);
    LoadA::store(ra, As[0], a_kc, tid);
    LoadB::store(rb, Bs[0], b_kc, tid);
  }
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int cur = kt & 1;
    if (kt + 1 < nk) {
      LoadA::load(ra, P.A, P.a_sm, P.a_sk, m0, P.M, kbeg + (kt + 1) * BK, kend, a_vec, tid);
      LoadB::load(rb, P.B, P.b_sn, P.b_sk, n0, P.N, kbeg + (kt + 1) * BK, kend, b_vec, tid);
    }
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], b[TN];
#pragma unroll
      for (int g = 0; g < RG; ++g) {
        const float4 v = *reinterpret_cast<const float4*>(&As[cur][k][g * (BM / RG) + ty * 4]);
        a[g * 4 + 0] = v.x; a[g * 4 + 1] = v.y; a[g * 4 + 2] = v.z; a[g * 4 + 3] = v.w;
      }
#pragma unroll
      for (int g = 0; g < CG; ++g) {
        const float4 v = *reinterpret_cast<const float4*>(&Bs[cur][k][g * (BN / CG) + tx * 4]);
        b[g * 4 + 0] = v.x; b[g * 4 + 1] = v.y; b[g * 4 + 2] = v.z; b[g * 4 + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (do_bsum && tid < BM) {
#pragma unroll
      for (int k = 0; k < BK; ++k) bsum += As[cur][k][tid];
    }
    if (kt + 1 < nk) {
      LoadA::store(ra, As[cur ^ 1], a_kc, tid);
      LoadB::store(rb, Bs[cur ^ 1], b_kc, tid);
    }
    __syncthreads();
  }

  // epilogue
  float* Cbase = P.C + int64_t(split) * P.c_split_stride;
  const int epi = P.epilogue;
  const bool need_bias = (epi >= EPI_BIAS && epi <= EPI_BIAS_TANH_SCALE);
  const bool need_aux = (epi == EPI_DTANH || epi == EPI_DRELU);
#pragma unroll
  for (int gi = 0; gi < RG; ++gi) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + gi * (BM / RG) + ty * 4 + i;
      if (m >= P.M) continue;
#pragma unroll
      for (int gj = 0; gj < CG; ++gj) {
        const int n = n0 + gj * (BN / CG) + tx * 4;
        if (n >= P.N) continue;
        float out[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const bool ok = n + j < P.N;
          const float bias = (need_bias && ok) ? __ldg(P.bias + n + j) : 0.f;
          const float aux = (need_aux && ok) ? __ldg(P.aux + int64_t(m) * P.ld_aux + n + j) : 0.f;
          out[j] = apply_epilogue(acc[gi * 4 + i][gj * 4 + j], epi, bias, aux, P.out_scale);
        }
        if (P.C_bf16 != nullptr) {
          __nv_bfloat16* bp = P.C_bf16 + int64_t(m) * P.ldc_bf16 + n;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (n + j < P.N) bp[j] = __float2bfloat16_rn(out[j]);
        }
        float* cp = Cbase + int64_t(m) * P.ldc + n;
        if (P.c_vec && n + 3 < P.N) {
          *reinterpret_cast<float4*>(cp) = make_float4(out[0], out[1], out[2], out[3]);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (n + j < P.N) cp[j] = out[j];
        }
      }
    }
  }
  if (do_bsum && tid < BM && m0 + tid < P.M) P.bias_grad[int64_t(split) * P.c_split_stride + m0 + tid] = bsum;
}

void gemm_group_add(GemmGroup& g, GemmProblem p, int bm, int bn, int split_k) {
  if (p.M <= 0 || p.N <= 0) return;
  if (split_k < 1) split_k = 1;
  int kps = (p.K + split_k - 1) / split_k;
  kps = ((kps + BK - 1) / BK) * BK;
  if (kps < BK) kps = BK;
  p.k_per_split = kps;
  p.split_k = split_k;  // trailing splits may be empty (K small): they store zeros, which keeps the sum right
  p.tiles_m = (p.M + bm - 1) / bm;
  p.tiles_n = (p.N + bn - 1) / bn;
  p.tile_begin = g.total_tiles;
  auto vec_ok = [](const float* base, int64_t s_mn, int64_t s_k, int MN, int K) {
    if (!aligned16(base)) return 0;
    if (s_k == 1) return (s_mn % 4 == 0 && K % 4 == 0) ? 1 : 0;
    return (s_k % 4 == 0 && MN % 4 == 0) ? 1 : 0;
  };
  p.a_vec = vec_ok(p.A, p.a_sm, p.a_sk, p.M, p.K);
  p.b_vec = vec_ok(p.B, p.b_sn, p.b_sk, p.N, p.K);
  p.c_vec = (aligned16(p.C) && p.ldc % 4 == 0 && p.c_split_stride % 4 == 0) ? 1 : 0;
  g.total_tiles += p.tiles_m * p.tiles_n * split_k;
  g.p[g.count++] = p;
}

bool gemm_prefer_large_tile(int M, int N) {
  // 128x128 tiles only when they still give every SM a CTA
  const int64_t tiles = int64_t((M + 127) / 128) * ((N + 127) / 128);
  return tiles >= num_sms() / 2 && N >= 96;
}

int launch_gemm_group(const GemmGroup& g, bool large_tile, cudaStream_t st) {
  if (g.total_tiles == 0) return B200PPO_OK;
  if (large_tile)
    gemm_group_kernel<128, 128, 8, 8><<<g.total_tiles, 256, 0, st>>>(g);
  else
    gemm_group_kernel<64, 64, 4, 4><<<g.total_tiles, 256, 0, st>>>(g);
  B2_LAUNCH_CHECK();
  return B200PPO_OK;
}

}  // namespace b200ppo
