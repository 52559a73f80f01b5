// Microbenchmark: what paces the EPILOGUE of a 128 x 256 fp32 accumulator tile (TMEM -> registers -> bias + tanh -> bf16
// -> shared memory) on sm_100a, with nothing else running on the SM?  Not part of the library.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/epi_probe profiles/epi_probe.cu && /tmp/epi_probe
// One CTA per SM, W epilogue warps (warp % 4 = TMEM lane quarter, 256 / (W / 4) columns each), R tiles back to back;
// reports cycles per tile.  Variants:
//   0 tcgen05.ld.x16 only          1 + bias add + cvt.rn.bf16x2 pack     2 + tanh.approx.bf16x2
//   3 + st.shared.v4 (swizzled)    4 as 3 with tcgen05.ld.x32            5 tanh.approx.f32 then pack (no cvt before the tanh)
//   6 as 3, pack by integer ops (round-half-up + PRMT) instead of cvt    7 as 3 without the tanh (dgrad-like: FMUL only)
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x)                                                                         \
  do {                                                                                \
    cudaError_t e = (x);                                                              \
    if (e != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);  \
      exit(1);                                                                        \
    }                                                                                 \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ uint32_t pack_cvt(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t pack_int(float lo, float hi) {  // round-half-up on the dropped 16 bits, then the two high halves
  const uint32_t a = __float_as_uint(lo) + 0x8000u, b = __float_as_uint(hi) + 0x8000u;
  return __byte_perm(a, b, 0x7632);
}
__device__ __forceinline__ uint32_t tanh2(uint32_t x) {
  uint32_t y;
  asm("tanh.approx.bf16x2 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}
__device__ __forceinline__ float tanh1(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int V>
__device__ __forceinline__ void slab16(const uint32_t (&v)[16], float bias, uint32_t srow, uint32_t sw, int s, uint32_t& sink) {
  uint32_t o[8];
  if (V == 0) {
#pragma unroll
    for (int j = 0; j < 16; ++j) sink ^= v[j];
    return;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float z0 = __uint_as_float(v[2 * j]) + bias, z1 = __uint_as_float(v[2 * j + 1]) + bias;
    if (V == 1) o[j] = pack_cvt(z0, z1);
    else if (V == 2 || V == 3 || V == 4) o[j] = tanh2(pack_cvt(z0, z1));
    else if (V == 5) o[j] = pack_cvt(tanh1(z0), tanh1(z1));
    else if (V == 6) o[j] = tanh2(pack_int(z0, z1));
    else o[j] = pack_cvt(z0 * bias, z1 * bias);
  }
  if (V == 1 || V == 2 || V == 5) {
#pragma unroll
    for (int j = 0; j < 8; ++j) sink ^= o[j];
  } else {
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(srow + (((2 * s) ^ sw) << 4)), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]) : "memory");
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(srow + (((2 * s + 1) ^ sw) << 4)), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]) : "memory");
  }
}

template <int V>
__global__ void __launch_bounds__(512, 1) probe(int warps, int reps, long long* out, uint32_t* sink_out, float bias) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  uint32_t sink = 0;
  const long long t0 = clock64();
  if (warp < warps) {
    const int q = warp & 3, grp = warp >> 2, groups = warps >> 2;
    const int cols = 256 / groups;  // columns of this warp
    const uint32_t lane_base = tmem + (uint32_t(q * 32) << 16);
    const uint32_t sw = lane & 7;
    for (int r = 0; r < reps; ++r) {
      const uint32_t tcol = lane_base + uint32_t((r & 1) * 256 + grp * cols);
      for (int c = 0; c < cols; c += 64) {
        const uint32_t srow = smem_u32(smem) + uint32_t(((grp * cols + c) >> 6) * 16384 + (q * 32 + lane) * 128);
        if (V == 4) {
          uint32_t a[32], b[32];
          tmem_ld32(tcol + c, a);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          tmem_ld32(tcol + c + 32, b);
          slab16<3>(reinterpret_cast<uint32_t(&)[16]>(a[0]), bias, srow, sw, 0, sink);
          slab16<3>(reinterpret_cast<uint32_t(&)[16]>(a[16]), bias, srow, sw, 1, sink);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          slab16<3>(reinterpret_cast<uint32_t(&)[16]>(b[0]), bias, srow, sw, 2, sink);
          slab16<3>(reinterpret_cast<uint32_t(&)[16]>(b[16]), bias, srow, sw, 3, sink);
        } else {
          uint32_t va[16], vb[16];
          tmem_ld16(tcol + c, va);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          tmem_ld16(tcol + c + 16, vb);
          slab16<V>(va, bias, srow, sw, 0, sink);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          tmem_ld16(tcol + c + 32, va);
          slab16<V>(vb, bias, srow, sw, 1, sink);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          tmem_ld16(tcol + c + 48, vb);
          slab16<V>(va, bias, srow, sw, 2, sink);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          slab16<V>(vb, bias, srow, sw, 3, sink);
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (sink == 0x12345678u) sink_out[0] = sink;
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

template <int V>
static void run(int warps, const char* what) {
  const int reps = 64, grid = 148;
  long long* d_out;
  uint32_t* d_sink;
  CK(cudaMalloc(&d_out, grid * sizeof(long long)));
  CK(cudaMalloc(&d_sink, 4));
  CK(cudaFuncSetAttribute(probe<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  for (int it = 0; it < 2; ++it) {
    probe<V><<<grid, 512, 65536>>>(warps, reps, d_out, d_sink, 0.25f);
    CK(cudaDeviceSynchronize());
  }
  long long h[148];
  CK(cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost));
  double mean = 0;
  for (int i = 0; i < grid; ++i) mean += double(h[i]);
  mean /= grid;
  printf("variant %d (%-46s) %2d warps: %7.0f cycles per 128x256 tile\n", V, what, warps, mean / reps);
  cudaFree(d_out);
  cudaFree(d_sink);
}

int main() {
  for (int w : {16, 8, 4}) {
    run<0>(w, "tcgen05.ld only");
    run<1>(w, "ld + bias + cvt pack");
    run<2>(w, "ld + bias + cvt + tanh.bf16x2");
    run<3>(w, "ld + bias + cvt + tanh.bf16x2 + st.shared");
    run<4>(w, "as 3 with tcgen05.ld.x32");
    run<5>(w, "ld + bias + tanh.f32 + cvt pack");
    run<6>(w, "as 3 with integer pack (no cvt)");
    run<7>(w, "ld + fmul + cvt + st.shared (no tanh)");
  }
  return 0;
}
