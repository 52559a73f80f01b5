// Fused forward + PPO loss + dgrad chain of the actor and critic MLPs (tc_chain.cu).
#pragma once
#include "tc_gemm.cuh"

namespace b200ppo {

// One network of the chain: three Linear layers in_dim -> 256 -> 256 -> out (out <= 32).
struct ChainNet {
  CUtensorMap w1;   // W1 [256][in]   K-major, box 64 x 128
  CUtensorMap w2k;  // W2 [256][256]  K-major, box 64 x 128      (forward)
  CUtensorMap w2m;  // W2 [256][256]  as [K = out][N = in], box 64 x 64      (dgrad, MN-major B operand)
  CUtensorMap w3k;  // W3 [out][256]  K-major, box 64 x (16 | 8) (forward; rows >= out zero-filled)
  CUtensorMap w3m;  // W3 [out][256]  as [K = out][N = hidden], box 64 x (32 | 16) (dgrad through the output layer)
  CUtensorMap sH1, sH2, sZ1, sZ2;  // bulk-store maps of the global copies [M][256], box 64 x 32 (rows past M clipped)
  const float *b1, *b2, *b3;
  __nv_bfloat16 *H1, *H2, *dZ1, *dZ2, *dZ3;  // global copies the weight-gradient kernel reads
  int pH1, pH2, pZ1, pZ2, pZ3;               // row pitches (elements)
};

struct ChainArgs {
  CUtensorMap x;     // observations [M][in] bf16, K-major, box 64 x 128
  ChainNet net[2];   // 0 actor, 1 critic
  TcPpo ppo;         // loss operands (dz_out / dz_pitch unused: seeds go to net[n].dZ3)
  float out_scale;
  int M, KB1, act, tiles2;
  int x_early;       // the observations are older than the previous kernel of the stream: X may be requested before the PDL wait
  // forward-only mode (rollout inference): outputs, each nullable; inf_noise [M][act] nullable (action = mean)
  float *inf_mean, *inf_value, *inf_action, *inf_logp;
  const float* inf_noise;
  // row strides (elements) of inf_action / inf_value / inf_logp — a rollout writes slice [:, t] of its [N, T, ...] buffers —
  // and a second destination of the value (the previous step's next_state_value: V(s') of step t - 1 is V(s) of step t)
  long long inf_ld_action, inf_ld_value, inf_ld_logp;
  float* inf_value2;
  long long* trace;  // debug: clock64 timeline of pair 0 (nullptr in production)
};

constexpr int kChainHidden = 256;
constexpr int TC_CHAIN_BK = 64;
extern long long* g_chain_trace;  // set by the debug entry point only
constexpr int kChainMaxIn = 384;

// in_dim <= 384, both hidden layers 256 wide, out <= 24, rows 32-byte aligned
bool tc_chain_shape_ok(int in_dim, int h1, int h2, int out_dim);
// infer: the forward-only instance (mean / action / log-prob / value instead of losses, seeds and dgrads)
int launch_tc_chain(const ChainArgs& a, cudaStream_t st, int* grid_out, bool infer = false);

}  // namespace b200ppo
