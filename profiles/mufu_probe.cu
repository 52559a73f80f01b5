// SFU throughput on sm_100a: results per clock per SM for tanh.approx.f32, ex2.approx, rcp.approx, and for an
// FMA-pipe polynomial, with 16 resident warps per SM issuing 8 independent chains each.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o profiles/bin/mufu_probe profiles/mufu_probe.cu
#include <cuda_runtime.h>
#include <cstdio>

template <int OP>
__global__ void probe(float* out, long long* cyc, int iters) {
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = 0.001f * float(threadIdx.x + j * 37 + 1);
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (OP == 0) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(v[j]));
      if (OP == 1) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[j]));
      if (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(v[j]));
      if (OP == 3) {  // 12 dependent FMAs: a degree-11 polynomial's worth of FMA-pipe work
        float x = v[j], a = x;
#pragma unroll
        for (int k = 0; k < 12; ++k) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a) : "f"(x), "f"(0.5f));
        v[j] = a;
      }
      if (OP == 4) {  // bf16x2 packed tanh
        unsigned u = __float_as_uint(v[j]);
        asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(u));
        v[j] = __uint_as_float(u);
      }
      if (OP == 6) {  // fp32 pair -> packed bf16x2 (F2FP)
        unsigned u;
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %1;" : "=r"(u) : "f"(v[j]));
        v[j] = __uint_as_float(u | 0x3f000000u);
      }
      if (OP == 7) {  // integer round-half-up to bf16 + byte-permute pack
        unsigned u = __float_as_uint(v[j]) + 0x8000u, w = __float_as_uint(v[(j + 1) & 7]) + 0x8000u, r;
        asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(r) : "r"(u), "r"(w));
        v[j] = __uint_as_float(r | 0x3f000000u);
      }
      if (OP == 5) {  // f16x2 packed tanh
        unsigned u = __float_as_uint(v[j]);
        asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(u));
        v[j] = __uint_as_float(u);
      }
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += v[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, double per_op) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 2000;
  probe<OP><<<148, 512>>>(out, cyc, iters);
  probe<OP><<<148, 512>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double mean = 0; for (int i = 0; i < 148; ++i) mean += h[i]; mean /= 148;
  printf("%-22s %8.0f cycles -> %6.2f results/clk/SM\n", name, mean, 512.0 * 8 * iters * per_op / mean);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0>("tanh.approx.f32", 1);
  run<1>("ex2.approx.f32", 1);
  run<2>("rcp.approx.f32", 1);
  run<3>("12 x fma.f32", 1);
  run<4>("tanh.approx.bf16x2", 2);
  run<5>("tanh.approx.f16x2", 2);
  run<6>("cvt.rn.bf16x2.f32", 1);
  run<7>("iadd+prmt pack", 1);
  return 0;
}
