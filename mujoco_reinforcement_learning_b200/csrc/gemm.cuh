// Grouped fp32 GEMM used by the actor/critic MLP forward, dgrad and wgrad (fp32-parity path).
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace b200ppo {

enum GemmEpilogue : int {
  EPI_STORE = 0,       // C = acc
  EPI_BIAS = 1,        // C = acc + bias[n]
  EPI_BIAS_TANH = 2,   // C = tanh(acc + bias[n])
  EPI_BIAS_RELU = 3,   // C = max(acc + bias[n], 0)
  EPI_BIAS_TANH_SCALE = 4,  // C = out_scale * tanh(acc + bias[n])      (linear/actor.py:28)
  EPI_DTANH = 5,       // C = acc * (1 - aux[m,n]^2)                    (dgrad through tanh, aux = layer output)
  EPI_DRELU = 6,       // C = acc * (aux[m,n] > 0)
};

// C[m,n] = sum_k A(m,k) * B(n,k); element (m,k) of A lives at A + m*a_sm + k*a_sk (same for B).
//   forward  Z = X W^T : A = X  (a_sm = ldx, a_sk = 1),  B = W  (b_sn = in,  b_sk = 1)
//   dgrad   dH = dZ W  : A = dZ (a_sm = ldz, a_sk = 1),  B = W  (b_sn = 1,   b_sk = in)
//   wgrad   dW = dZ^T X: A = dZ (a_sm = 1,  a_sk = ldz),  B = X  (b_sn = 1,   b_sk = ldx)   (k = sample)
struct GemmProblem {
  const float* A;
  const float* B;
  float* C;
  const float* bias;
  const float* aux;
  float* bias_grad;  // wgrad only: bias_grad[m] = sum_k A(m,k) (per split), nullable
  __nv_bfloat16* C_bf16;  // optional bf16 mirror of C (operand of the tensor-core path), row pitch ldc_bf16
  int ldc_bf16;
  int64_t a_sm, a_sk, b_sn, b_sk;
  int64_t c_split_stride;  // elements between split-K partials (applies to C and bias_grad)
  int M, N, K;
  int ldc, ld_aux;
  int epilogue;
  int split_k, k_per_split;
  int tiles_m, tiles_n, tile_begin;
  int a_vec, b_vec;  // 128-bit loads legal for this operand
  int c_vec;
  float out_scale;
  // Optional upper bounds of max |A| / max |B| (fp32-tolerance tensor-core path, gemm_split.cu: the operand's fp16 scale comes
  // from its largest magnitude; a known bound — 1 for tanh activations, the rollout's maximum for observations — saves
  // the pass that measures it).  0 / nullptr: unknown.  The device pointer wins over the constant.
  float a_bound, b_bound;
  const float *a_bound_dev, *b_bound_dev;
};

constexpr int kMaxGemmProblems = 2 * B200PPO_MAX_LAYERS;

struct GemmGroup {
  GemmProblem p[kMaxGemmProblems];
  int count;
  int total_tiles;
};

// Fills tiles/vector flags for tile shape (bm, bn) and appends to the group.
void gemm_group_add(GemmGroup& g, GemmProblem p, int bm, int bn, int split_k);

// Launch with the large (128x128) or small (64x64) tile; every problem in a group uses the same tile.
int launch_gemm_group(const GemmGroup& g, bool large_tile, cudaStream_t st);

// Tile choice + split-K heuristic shared by the callers.
bool gemm_prefer_large_tile(int M, int N);

}  // namespace b200ppo
