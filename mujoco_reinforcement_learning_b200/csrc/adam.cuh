// Internal interface of the fused Adam kernel (adam.cu).
#pragma once
#include "common.cuh"

namespace b200ppo {

// Scalar factors of one Adam step, computed in double on the host like torch's Python code, then rounded
// to fp32 where ATen applies them to fp32 tensors.
struct AdamScalars {
  float one_minus_b1, b2, one_minus_b2, bc2_sqrt, eps, neg_step_size;
};

AdamScalars make_adam_scalars(double lr, double beta1, double beta2, double eps, int64_t step);

// Parameters [0, seg_split) use s0 (actor optimiser), [seg_split, n) use s1 (critic optimiser).
// grad_out (nullable) receives the summed gradient.
int launch_adam(float* params, const float* grads, int n_partials, int64_t partial_stride, float* exp_avg,
                float* exp_avg_sq, int64_t n, int64_t seg_split, const AdamScalars& s0, const AdamScalars& s1,
                float* grad_out, cudaStream_t st);

int launch_reduce_partials(const float* grads, int n_partials, int64_t partial_stride, int64_t n, float* out,
                           cudaStream_t st);

}  // namespace b200ppo
