// fp32 dot products of RB shared-memory rows against one global row (FFMA path).
#pragma once
#include "common.cuh"

namespace vpa {

// xs[RB][D] <- X[r0 .. r0+RB) (zero rows past n_rows); all threads of the CTA cooperate.
template <int RB>
__device__ __forceinline__ void load_rows_to_smem(float* xs, const float* __restrict__ X, int64_t r0,
                                                  int64_t n_rows, int D) {
  const int nvec = D >> 2;
  float4* xs4 = reinterpret_cast<float4*>(xs);
  for (int i = threadIdx.x; i < RB * nvec; i += blockDim.x) {
    const int r = i / nvec, c = i - r * nvec;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r0 + r < n_rows) v = __ldg(reinterpret_cast<const float4*>(X + (r0 + r) * D) + c);
    xs4[i] = v;
  }
}

// out[r] = <xs[r], y>, four independent partial sums per row (pairwise-ish accumulation).
template <int RB>
__device__ __forceinline__ void dot_rows(const float* xs, const float* __restrict__ y, int D, float (&out)[RB]) {
  const int nvec = D >> 2;
  const float4* xs4 = reinterpret_cast<const float4*>(xs);
  const float4* y4p = reinterpret_cast<const float4*>(y);
  float4 acc[RB];
#pragma unroll
  for (int r = 0; r < RB; ++r) acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
  for (int k = 0; k < nvec; ++k) {
    const float4 q = __ldg(y4p + k);
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      const float4 x = xs4[r * nvec + k];
      acc[r].x = fmaf(x.x, q.x, acc[r].x);
      acc[r].y = fmaf(x.y, q.y, acc[r].y);
      acc[r].z = fmaf(x.z, q.z, acc[r].z);
      acc[r].w = fmaf(x.w, q.w, acc[r].w);
    }
  }
#pragma unroll
  for (int r = 0; r < RB; ++r) out[r] = (acc[r].x + acc[r].y) + (acc[r].z + acc[r].w);
}

}  // namespace vpa
