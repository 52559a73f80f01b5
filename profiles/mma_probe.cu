// Microbenchmark: what paces tcgen05.mma in a persistent, double-buffered-accumulator GEMM main loop on sm_100a?
// Not part of the library.  Build + run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mma_probe profiles/mma_probe.cu && /tmp/mma_probe
// Operand data is whatever is in shared memory (timing of kind::f16 MMAs is data independent); what varies is
//   CG   cta_group 1 (M=128 per CTA) or 2 (M=256 per CTA pair, each CTA holds half of B)
//   BN   MMA N
//   epi  0 none | 1 tcgen05.ld only | 2 ld + bf16 pack + st.shared + ld.shared + coalesced global stores
//        3 as 2 without the global stores | 4 as 2 with every tile stored to the same 64 KB (L2-hot lines)
//   tma  0 no copy traffic | 1 a producer lane streams 16 KB bulk copies into the A ring at the MMA's pace
// Output: cycles per MMA instruction against the floor 128*BN/(256*CG) (B300_MICROARCH.md "tcgen05 floor").
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e = (x);                                                           \
    if (e != cudaSuccess) {                                                        \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(1);                                                                     \
    }                                                                              \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= uint64_t((smem_addr & 0x3FFFF) >> 4);
  d |= uint64_t((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= uint64_t((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= uint64_t(1) << 46;
  d |= uint64_t(2) << 61;
  return d;
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

constexpr int BK = 64, BM = 128, A_BYTES = BM * BK * 2, STAGES = 3, KB = 4, THREADS = 320;

struct Result {
  long long cycles;
  long long mmas;
};

template <int CG, int BN>
__global__ void __launch_bounds__(THREADS, 1) probe(int tiles, int epi, int tma, const uint8_t* gsrc, uint8_t* gdst, Result* res) {
  constexpr int BN_LOCAL = BN / CG;                 // B rows held by this CTA
  constexpr int W_KB_BYTES = BN_LOCAL * BK * 2;
  constexpr uint32_t TMEM_COLS = 2 * BN <= 256 ? 256 : 512;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sW = smem;
  uint8_t* sA = sW + KB * W_KB_BYTES;
  uint8_t* stage = sA + STAGES * A_BYTES;  // 8 warps x 4 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage + 8 * 4096);
  uint64_t* a_full = bars;
  uint64_t* a_empty = bars + STAGES;
  uint64_t* acc_full = a_empty + STAGES;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* done_bar = acc_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = CG == 2 ? cluster_rank() : 0;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 8 * CG); }
    mbar_init(done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  } else if (warp == 2) {
    if (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  // fill the operand area with small finite bf16 values
  for (int i = threadIdx.x; i < (KB * W_KB_BYTES + STAGES * A_BYTES) / 4; i += THREADS) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0 && tma) {  // copy-engine traffic into the A ring, paced by the MMA's stage releases
      int it = 0;
      for (int t = 0; t < tiles; ++t)
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(&a_empty[s], ((it / STAGES) & 1) ^ 1);
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&a_full[s])), "r"(A_BYTES) : "memory");
          const uint8_t* src = gsrc + (size_t(blockIdx.x) * 16 + size_t(it % 16)) * A_BYTES;  // 38 MB in all: L2 resident
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                           smem_u32(sA + s * A_BYTES)),
                       "l"(src), "r"(A_BYTES), "r"(smem_u32(&a_full[s]))
                       : "memory");
        }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank != 0 && tma) {  // the peer's copies must have landed before its CTA may exit
      for (int it = 0; it < tiles * KB; ++it) mbar_wait(&a_full[it % STAGES], (it / STAGES) & 1);
    }
    if (lane == 0 && rank == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(BN >> 3) << 17) | (uint32_t((BM * CG) >> 4) << 24);
      int it = 0;
      const long long t0 = clock64();
      for (int t = 0; t < tiles; ++t) {
        const int buf = t & 1;
        if (epi) {
          mbar_wait(&acc_empty[buf], ((t >> 1) & 1) ^ 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int s = it % STAGES;
          if (tma) {
            mbar_wait(&a_full[s], (it / STAGES) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          }
          const uint32_t a_addr = smem_u32(sA + s * A_BYTES), b_addr = smem_u32(sW + kb * W_KB_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t ad = umma_desc(a_addr + k * 32, 0, 1024), bd = umma_desc(b_addr + k * 32, 0, 1024);
            const uint32_t acc = (kb > 0 || k > 0) ? 1u : 0u;
            if (CG == 1)
              asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(
                               tmem_base + uint32_t(buf * BN)),
                           "l"(ad), "l"(bd), "r"(idesc), "r"(acc)
                           : "memory");
            else
              asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(
                               tmem_base + uint32_t(buf * BN)),
                           "l"(ad), "l"(bd), "r"(idesc), "r"(acc)
                           : "memory");
          }
          if (tma) {
            if (CG == 1)
              asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&a_empty[s])) : "memory");
            else
              asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                               smem_u32(&a_empty[s])),
                           "h"(uint16_t(3))
                           : "memory");
          }
        }
        if (CG == 1)
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&acc_full[buf])) : "memory");
        else
          asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                           smem_u32(&acc_full[buf])),
                       "h"(uint16_t(3))
                       : "memory");
      }
      if (CG == 1)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(done_bar)) : "memory");
      else
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(done_bar)),
                     "h"(uint16_t(1))
                     : "memory");
      mbar_wait(done_bar, 0);
      const long long t1 = clock64();
      res[blockIdx.x].cycles = t1 - t0;
      res[blockIdx.x].mmas = (long long)tiles * KB * (BK / 16);
    }
  } else if (epi) {
    const int q = warp & 3, half = (warp - 2) >> 2;
    uint8_t* st = stage + (warp - 2) * 4096;
    const uint32_t leader_empty[2] = {CG == 2 ? mapa(smem_u32(&acc_empty[0]), 0) : smem_u32(&acc_empty[0]),
                                      CG == 2 ? mapa(smem_u32(&acc_empty[1]), 0) : smem_u32(&acc_empty[1])};
    float sink = 0.f;
    for (int t = 0; t < tiles; ++t) {
      const int buf = t & 1;
      mbar_wait(&acc_full[buf], (t >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      constexpr int COLS = BN / 2;  // per warp
      for (int c0 = 0; c0 < COLS; c0 += 64) {
#pragma unroll
        for (int c = 0; c < 64; c += 16) {
          uint32_t v[16];
          const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(buf * BN + half * COLS + c0 + c);
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
              : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
              : "r"(taddr)
              : "memory");
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (epi == 1) {
#pragma unroll
            for (int j = 0; j < 16; ++j) sink += __uint_as_float(v[j]);
          } else {
            uint32_t pk[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              __nv_bfloat162 h = __floats2bfloat162_rn(__uint_as_float(v[2 * j]) + 1.f, __uint_as_float(v[2 * j + 1]) + 1.f);
              pk[j] = *reinterpret_cast<uint32_t*>(&h);
            }
            // row `lane` of a 32 x 128-byte tile, 16-byte chunks XOR-swizzled by the row
            const int ch = c / 8;
            *reinterpret_cast<uint4*>(st + lane * 128 + (((ch) ^ (lane & 7)) * 16)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            *reinterpret_cast<uint4*>(st + lane * 128 + (((ch + 1) ^ (lane & 7)) * 16)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          }
        }
        if (epi >= 2) {
          __syncwarp();
          // 4 rows per instruction: 8 lanes x 16 bytes per 128-byte row
#pragma unroll
          for (int r = 0; r < 32; r += 4) {
            const int row = r + (lane >> 3), chk = lane & 7;
            const uint4 val = *reinterpret_cast<const uint4*>(st + row * 128 + ((chk ^ (row & 7)) * 16));
            const size_t grow = (size_t(blockIdx.x) * tiles + t) % 512 * 128 + q * 32 + row;
            if (epi == 2) *reinterpret_cast<uint4*>(gdst + (grow * (BN * 2)) + (half * COLS + c0) * 2 + chk * 16) = val;
            else if (epi == 3) sink += __uint_as_float(val.x ^ val.y ^ val.z ^ val.w);
            else *reinterpret_cast<uint4*>(gdst + ((size_t(blockIdx.x) * 128 + q * 32 + row) * (BN * 2)) + (half * COLS + c0) * 2 + chk * 16) = val;
          }
          __syncwarp();
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(leader_empty[buf]) : "memory");
    }
    if (sink == 123.456f) gdst[0] = 1;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  if (warp == 2) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (CG == 1)
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

template <int CG, int BN>
static void run(int tiles, int epi, int tma, const uint8_t* gsrc, uint8_t* gdst, Result* dres) {
  const int smem = 1024 + KB * (BN / CG) * BK * 2 + STAGES * A_BYTES + 8 * 4096 + 256;
  CK(cudaFuncSetAttribute(probe<CG, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(sms / 2 * 2);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CG;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  CK(cudaMemset(dres, 0, sizeof(Result) * 256));
  for (int rep = 0; rep < 2; ++rep) CK(cudaLaunchKernelEx(&cfg, probe<CG, BN>, tiles, epi, tma, gsrc, gdst, dres));
  CK(cudaDeviceSynchronize());
  Result h[256];
  CK(cudaMemcpy(h, dres, sizeof(h), cudaMemcpyDeviceToHost));
  double sum = 0;
  long long mx = 0;
  int n = 0;
  for (int i = 0; i < 256; ++i)
    if (h[i].mmas) {
      sum += double(h[i].cycles) / h[i].mmas;
      mx = h[i].cycles > mx ? h[i].cycles : mx;
      ++n;
    }
  const double floor_c = 128.0 * BN / (256.0 * CG) * CG;  // per instruction: M = 128*CG rows
  printf("cta_group %d  N %3d  epi %d  tma %d : %7.1f cycles/MMA (mean of %d issuers, floor %.0f)  -> %.0f%% of the tensor pipe\n", CG, BN, epi, tma,
         sum / n, n, floor_c, 100.0 * floor_c / (sum / n));
}

int main() {
  uint8_t *gsrc, *gdst;
  Result* dres;
  CK(cudaMalloc(&gsrc, size_t(160) * 64 * A_BYTES));
  CK(cudaMemset(gsrc, 0x3c, size_t(160) * 64 * A_BYTES));
  CK(cudaMalloc(&gdst, size_t(4096) * 128 * 512 + 4096));
  CK(cudaMalloc(&dres, sizeof(Result) * 256));
  const int tiles = 64;
  for (int tma = 0; tma < 2; ++tma)
    for (int epi = 0; epi < 5; ++epi) {
      run<1, 128>(tiles, epi, tma, gsrc, gdst, dres);
      run<1, 256>(tiles, epi, tma, gsrc, gdst, dres);
      run<2, 256>(tiles, epi, tma, gsrc, gdst, dres);
    }
  return 0;
}
