// Persistent, weights-stationary tcgen05 GEMM for the forward and dgrad passes of the MLP (large minibatches).
//
// C[M, N] = A[M, K] * W[N, K]^T with M = minibatch rows (tens of thousands), N <= 256, K <= 384: the weight matrix
// is a few hundred KB at most, the activations are what streams.  The tile-per-CTA kernel (tc_gemm.cu) re-reads the
// weight tile for every 128 rows, which makes it L2->SM bandwidth bound (64 FLOP per byte staged).  Here a CTA
// loads its problem's whole W into shared memory ONCE (TMA, 128-byte swizzle, K-major), then walks over its share
// of the 128-row tiles: only A moves (4x the arithmetic intensity), the accumulator is double-buffered in TMEM
// (2 x BN columns) so the epilogue of tile t overlaps the MMAs of tile t+1.
//   warp 0      TMA producer: W once, then the A k-blocks of every tile through an n-stage mbarrier ring
//   warp 1      MMA issuer (one lane): tcgen05.mma M=128, N=BN, K=16, commit -> frees the A stage / publishes the tile
//   warps 2-9   epilogue (shared with tc_gemm.cu): tcgen05.ld, bias + MUFU tanh | act' multiply, bf16/fp32 stores
#include <algorithm>

#include "tc_common.cuh"

namespace b200ppo {

constexpr int kWsMaxSlots = 4;

// A slot = (problem, N tile): its CTAs keep that BN-wide slice of the problem's W in shared memory.
struct WsGroup {
  TcProblem p[2];
  int count;
  int n_slots;
  int slot_prob[kWsMaxSlots], slot_n0[kWsMaxSlots];
  int cta_begin[kWsMaxSlots + 1];  // CTAs [cta_begin[i], cta_begin[i+1]) serve slot i
  long long* trace;                // debug: clock64 timeline of CTA 0, [tile][8] (nullptr in production)
};

long long* g_ws_trace = nullptr;  // set by the debug entry point only

#define WS_TRACE(tile_idx, slot_idx)                                                        \
  do {                                                                                      \
    if (grp.trace != nullptr && blockIdx.x == 0 && (tile_idx) < 64) grp.trace[(tile_idx) * 8 + (slot_idx)] = clock64(); \
  } while (0)

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int BN>
__global__ void __launch_bounds__(TC_THREADS, 1) tc_ws_kernel(const __grid_constant__ WsGroup grp, int a_stages) {
  constexpr uint32_t TMEM_COLS = 2 * BN <= 32 ? 32 : (2 * BN <= 64 ? 64 : (2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512)));
  constexpr int W_KB_BYTES = BN * TC_BK * 2;  // one 64-wide k-block of W
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int slot = 0;
#pragma unroll 1
  for (int i = 1; i < grp.n_slots; ++i)
    if (int(blockIdx.x) >= grp.cta_begin[i]) slot = i;
  const TcProblem& P = grp.p[grp.slot_prob[slot]];
  const int n0 = grp.slot_n0[slot];
  const int cta_local = int(blockIdx.x) - grp.cta_begin[slot];
  const int ctas = grp.cta_begin[slot + 1] - grp.cta_begin[slot];
  const int KB = (P.K + TC_BK - 1) / TC_BK;
  const int tiles_m = P.tiles_m;

  uint8_t* sW = smem;
  uint8_t* sA = smem + size_t(KB) * W_KB_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sA + size_t(a_stages) * TC_A_BYTES);
  uint64_t* w_full = bars;
  uint64_t* a_full = bars + 1;
  uint64_t* a_empty = a_full + a_stages;
  uint64_t* acc_full = a_empty + a_stages;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* bias_s = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 512);  // BN floats (<= 256)
  float* consts_s = bias_s + 256;                                                    // 96 floats (fused PPO epilogue)
  float* red_s = bias_s + 352;                                                       // 8 x 34 floats
  uint8_t* stage_area = reinterpret_cast<uint8_t*>(bias_s) + 3072;                  // TC_EPI_WARPS x 2 x 4 KB

  tc_stage_bias(P, n0, BN, bias_s, threadIdx.x, TC_THREADS);
  tc_ppo_stage_consts(P, consts_s, threadIdx.x);
  if (warp == 1 && lane == 0) {
    mbar_init(w_full, 1);
    for (int s = 0; s < a_stages; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    // fused PPO epilogue: the two warp groups alternate tiles, so only four warps release an accumulator
    const uint32_t releasers = P.epilogue >= TC_EPI_PPO_ACTOR ? TC_EPI_WARPS / 2 : TC_EPI_WARPS;
    for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], releasers); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  } else if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer =====
      mbar_expect_tx(w_full, uint32_t(KB) * W_KB_BYTES);
      if (!P.b_mn_major) {
        for (int kb = 0; kb < KB; ++kb) tma_load_2d(sW + size_t(kb) * W_KB_BYTES, &P.tmB, w_full, kb * TC_BK, n0);
      } else {  // W given as [K][N] (the dgrad reads nn.Linear's [out, in] weight as is): 64-wide N atoms per k-block
        for (int kb = 0; kb < KB; ++kb)
#pragma unroll
          for (int j = 0; j < BN / 64; ++j)
            tma_load_2d(sW + size_t(kb) * W_KB_BYTES + j * 64 * TC_BK * 2, &P.tmB, w_full, n0 + 64 * j, kb * TC_BK);
      }
      int it = 0;
      int tt = 0;
      for (int tile = cta_local; tile < tiles_m; tile += ctas, ++tt) {
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int s = it % a_stages;
          const uint32_t ph = (it / a_stages) & 1;
          mbar_wait(&a_empty[s], ph ^ 1);
          if (kb == 0) WS_TRACE(tt, 0);
          mbar_expect_tx(&a_full[s], TC_A_BYTES);
          tma_load_2d(sA + size_t(s) * TC_A_BYTES, &P.tmA, &a_full[s], kb * TC_BK, tile * TC_BM);
          if (kb == KB - 1) WS_TRACE(tt, 1);
        }
      }
    }
  } else if (warp == 1) {  // ===== MMA issuer: ONE thread runs the whole loop (31 idle lanes would only add polling) =====
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(P.b_mn_major != 0) << 16) | (uint32_t(BN >> 3) << 17) |
                             (uint32_t(TC_BM >> 4) << 24);
      const uint32_t b_lbo = P.b_mn_major ? TC_BK * 128 : 0, b_kstep = P.b_mn_major ? 2048 : 32;
      mbar_wait(w_full, 0);
      int it = 0, t = 0;
      for (int tile = cta_local; tile < tiles_m; tile += ctas, ++t) {
        const int buf = t & 1;
        mbar_wait(&acc_empty[buf], ((t >> 1) & 1) ^ 1);  // epilogue has drained this accumulator
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        WS_TRACE(t, 2);
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int s = it % a_stages;
          const uint32_t ph = (it / a_stages) & 1;
          mbar_wait(&a_full[s], ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (kb == 0) WS_TRACE(t, 3);
          if (kb == KB - 1) WS_TRACE(t, 4);
          const uint32_t a_addr = smem_u32(sA + size_t(s) * TC_A_BYTES), b_addr = smem_u32(sW + size_t(kb) * W_KB_BYTES);
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k)
            umma_bf16(tmem_base + uint32_t(buf * BN), umma_desc(a_addr + k * 32, 0, 1024), umma_desc(b_addr + k * b_kstep, b_lbo, 1024),
                      idesc, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit(&a_empty[s]);
          if (kb == KB - 1) umma_commit(&acc_full[buf]);
        }
      }
    }
  } else {  // ===== epilogue warps =====
    int t = 0;
    if constexpr (BN <= 128) {
      // two staging tiles per warp: while tile t is finished and stored from one, the dgrad's activation slab of tile
      // t + ctas is already streaming into the other (cp.async), i.e. behind the MMAs of the next tile
      uint8_t* st[2] = {stage_area + (warp - 2) * 2 * TC_STAGE_BYTES, stage_area + (warp - 2) * 2 * TC_STAGE_BYTES + TC_STAGE_BYTES};
      if (P.epilogue >= TC_EPI_PPO_ACTOR) {  // output layers with the PPO loss fused in
        // warp group g (warps 2-5 / 6-9) takes the tiles t with t % 2 == g: tile t lives in TMEM buffer t % 2, so each
        // group always drains the same accumulator while the other group works on the next tile.
        PpoAcc acc;
        acc.clear();
        const int grp_id = (warp - 2) >> 2;
        const int step = 2 * ctas;
        int tile = cta_local + grp_id * ctas, u = 0;  // u counts this group's tiles
        tc_ppo_issue(P, tile * TC_BM, warp, lane, st[0], tile < tiles_m);
        for (; tile < tiles_m; tile += step, ++u) {
          const int next = tile + step;
          tc_ppo_issue(P, next * TC_BM, warp, lane, st[(u & 1) ^ 1], next < tiles_m);
          tc_epilogue_ppo<BN>(P, tmem_base + uint32_t(grp_id * BN), tile * TC_BM, warp, lane, &acc_full[grp_id], uint32_t(u & 1),
                              st[u & 1], bias_s, consts_s, 1, acc, true);
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[grp_id]);
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        tc_ppo_finish(P, warp, lane, red_s, acc, true);
      } else {
      if (P.staged) tc_issue_aux<BN>(P, cta_local * TC_BM, n0, warp, lane, st[0], cta_local < tiles_m);
      for (int tile = cta_local; tile < tiles_m; tile += ctas, ++t) {
        const int buf = t & 1;
        if (P.staged) {
          const int next = tile + ctas;
          if (warp == 2 && lane == 0) WS_TRACE(t, 5);
          tc_issue_aux<BN>(P, next * TC_BM, n0, warp, lane, st[buf ^ 1], next < tiles_m);
          tc_epilogue_staged<BN>(P, 0, tmem_base + uint32_t(buf * BN), true, tile * TC_BM, n0, warp, lane, &acc_full[buf],
                                 uint32_t((t >> 1) & 1), st[buf], bias_s, 1);
          if (warp == 2 && lane == 0) WS_TRACE(t, 6);
          if (lane == 0 && grp.trace != nullptr && blockIdx.x == 0 && t < 64)  // slowest epilogue warp of the tile
            atomicMax(reinterpret_cast<unsigned long long*>(grp.trace) + t * 8 + 7, (unsigned long long)clock64());
        } else {
          tc_epilogue<BN>(P, 0, tmem_base + uint32_t(buf * BN), true, tile * TC_BM, n0, warp, lane, &acc_full[buf], uint32_t((t >> 1) & 1), bias_s);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[buf]);
      }
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      }
    } else {
      for (int tile = cta_local; tile < tiles_m; tile += ctas, ++t) {
        const int buf = t & 1;
        tc_epilogue<BN>(P, 0, tmem_base + uint32_t(buf * BN), true, tile * TC_BM, n0, warp, lane, &acc_full[buf], uint32_t((t >> 1) & 1), bias_s);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[buf]);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

constexpr int kWsMaxSmem = 227 * 1024;
constexpr int kWsFixedSmem = 1024 + 512 + 3072 + 2 * TC_EPI_WARPS * TC_STAGE_BYTES;  // staging is double-buffered here  // alignment slack + barriers + epilogue staging

// N tile and A-ring depth for a group, or bn = 0 when the weights-stationary kernel does not apply.
// N tile <= 128: each epilogue warp then owns 64 columns = ONE 128-byte staging row, so the dgrad's whole activation
// tile is requested in a single burst before the accumulator is awaited (and W slices of <= 96 KB leave a deep A ring).
static void ws_plan(int maxN, int maxK, int* bn_out, int* stages_out) {
  *bn_out = 0;
  *stages_out = 0;
  const int kb = (maxK + TC_BK - 1) / TC_BK;
  for (int bn : {128, 64}) {
    if (bn > 64 && maxN <= bn / 2) continue;  // a narrower tile covers N
    const int64_t w_bytes = int64_t(kb) * bn * TC_BK * 2;
    const int64_t avail = kWsMaxSmem - kWsFixedSmem - w_bytes;
    int stages = int(avail / TC_A_BYTES);
    if (stages < 3) continue;
    if ((maxN + bn - 1) / bn * 2 > kWsMaxSlots) continue;
    *bn_out = bn;
    *stages_out = stages > 8 ? 8 : stages;
    return;
  }
}

bool tc_ws_applicable(int64_t total_tiles_m, int maxN, int maxK) {
  int bn, st;
  ws_plan(maxN, maxK, &bn, &st);
  // persistence pays once every CTA owns at least ~two row tiles; below that the one-tile kernel has more CTAs
  return bn != 0 && total_tiles_m >= 2ll * num_sms();
}

int tc_ws_bn(int maxN, int maxK) {
  int bn, st;
  ws_plan(maxN, maxK, &bn, &st);
  return bn;
}

template <int BN>
static int launch_ws_bn(const WsGroup& g, int stages, int kb_max, int grid, cudaStream_t st) {
  const int smem = kb_max * BN * TC_BK * 2 + stages * TC_A_BYTES + kWsFixedSmem;
  static int configured = 0;
  if (configured < smem) {
    B2_CUDA(cudaFuncSetAttribute(tc_ws_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = smem;
  }
  tc_ws_kernel<BN><<<grid, TC_THREADS, smem, st>>>(g, stages);
  B2_LAUNCH_CHECK();
  return B200PPO_OK;
}

// g: one or two forward / dgrad problems built with tc_group_add(..., bn = tc_ws_bn(maxN, maxK), split 1), K-major.
int launch_tc_ws(const TcGroup& g, cudaStream_t st, int* grid_out) {
  B2_CHECK_ARG(g.count >= 1 && g.count <= 2, "weights-stationary launch takes one or two problems");
  int maxN = 0, maxK = 0;
  for (int i = 0; i < g.count; ++i) {
    B2_CHECK_ARG(!g.p[i].a_mn_major && g.p[i].split_k == 1, "weights-stationary kernel: K-major A, no split-K");
    maxN = std::max(maxN, g.p[i].N);
    maxK = std::max(maxK, g.p[i].K);
  }
  int bn, stages;
  ws_plan(maxN, maxK, &bn, &stages);
  B2_CHECK_ARG(bn != 0, "weights do not fit in shared memory");
  WsGroup w{};
  w.count = g.count;
  int64_t work = 0;  // row tiles summed over slots
  for (int i = 0; i < g.count; ++i) {
    w.p[i] = g.p[i];
    for (int n0 = 0; n0 < g.p[i].N; n0 += bn) {
      w.slot_prob[w.n_slots] = i;
      w.slot_n0[w.n_slots] = n0;
      ++w.n_slots;
      work += g.p[i].tiles_m;
    }
  }
  const int grid = int(std::min<int64_t>(num_sms(), work));
  int begin = 0;
  for (int sidx = 0; sidx < w.n_slots; ++sidx) {
    w.cta_begin[sidx] = begin;
    const int64_t tiles = g.p[w.slot_prob[sidx]].tiles_m;
    int share = (sidx == w.n_slots - 1) ? grid - begin : int((int64_t(grid) * tiles + work / 2) / work);
    if (share < 1) share = 1;
    begin += share;
  }
  w.cta_begin[w.n_slots] = std::max(begin, grid);
  w.trace = g_ws_trace;
  const int kb_max = (maxK + TC_BK - 1) / TC_BK;
  const int total = w.cta_begin[w.n_slots];
  if (grid_out) *grid_out = total;
  switch (bn) {
    case 64: return launch_ws_bn<64>(w, stages, kb_max, total, st);
    case 128: return launch_ws_bn<128>(w, stages, kb_max, total, st);
    default: return launch_ws_bn<256>(w, stages, kb_max, total, st);
  }
}

}  // namespace b200ppo
