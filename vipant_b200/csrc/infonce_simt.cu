// fp32 "exact mode" InfoNCE sweeps (VPA_PREC_FP32_SIMT): plain FFMA dot products, no tensor cores.
// Serves the fp32-mode parity bars (loss 1e-4 / grads 1e-3), any D % 4 == 0 up to 1024, and is the
// on-device cross-check of the tcgen05 path.  Same decomposition as the tensor-core path: problem 0
// sweeps A_loc against T_all (row statistics / dA), problem 1 sweeps T_loc against A_all (column
// statistics / dT); the B x B logits are never written to memory.
#include "common.cuh"
#include "simt_dot.cuh"

namespace vpa {

constexpr int kRB = 8;          // X rows per CTA
constexpr int kThreads = 256;
constexpr int kJC = 1024;       // backward: columns per G chunk staged in shared memory

__device__ __forceinline__ void ml_merge(float& m, float& l, float m2, float l2) {
  float mn = fmaxf(m, m2);
  float a = (m == -INFINITY) ? 0.f : l * exp2f(m - mn);
  float b = (m2 == -INFINITY) ? 0.f : l2 * exp2f(m2 - mn);
  m = mn;
  l = a + b;
}

__global__ void __launch_bounds__(kThreads)
simt_fwd_kernel(const float* __restrict__ x0, const float* __restrict__ y0,
                const float* __restrict__ x1, const float* __restrict__ y1,
                int64_t rows_local, int64_t rows_global, int D, int n_blk,
                const float* __restrict__ logit_scale, float scale_cap, float2* __restrict__ part) {
  extern __shared__ float4 smem4[];
  float* xs = reinterpret_cast<float*>(smem4);                 // [kRB][D]
  const int p = blockIdx.x >= n_blk;
  const int blk = blockIdx.x - p * n_blk;
  const float* X = p ? x1 : x0;
  const float* Y = p ? y1 : y0;
  const int64_t r0 = (int64_t)blk * kRB;
  load_rows_to_smem<kRB>(xs, X, r0, rows_local, D);
  __syncthreads();
  const float s2 = fminf(expf(*logit_scale), scale_cap) * kLog2e;
  float m[kRB], l[kRB];
#pragma unroll
  for (int r = 0; r < kRB; ++r) { m[r] = -INFINITY; l[r] = 0.f; }
  for (int64_t j = threadIdx.x; j < rows_global; j += kThreads) {
    float dots[kRB];
    dot_rows<kRB>(xs, Y + j * D, D, dots);
#pragma unroll
    for (int r = 0; r < kRB; ++r) {
      float v = dots[r] * s2;
      float mn = fmaxf(m[r], v);
      l[r] = l[r] * exp2f(m[r] - mn) + exp2f(v - mn);   // m = -inf first time: l = 0 * 0 + 1
      m[r] = mn;
    }
  }
  // block-wide merge of the per-thread (m, l) pairs, fixed order
  __shared__ float sm_m[kRB][kThreads / 32], sm_l[kRB][kThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int r = 0; r < kRB; ++r) {
    float mm = m[r], ll = l[r];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float m2 = __shfl_xor_sync(0xffffffffu, mm, o), l2 = __shfl_xor_sync(0xffffffffu, ll, o);
      ml_merge(mm, ll, m2, l2);
    }
    if (lane == 0) { sm_m[r][warp] = mm; sm_l[r][warp] = ll; }
  }
  __syncthreads();
  if (threadIdx.x < kRB && r0 + threadIdx.x < rows_local) {
    const int r = threadIdx.x;
    float mm = sm_m[r][0], ll = sm_l[r][0];
    for (int w = 1; w < kThreads / 32; ++w) ml_merge(mm, ll, sm_m[r][w], sm_l[r][w]);
    part[(int64_t)p * rows_local + r0 + r] = make_float2(mm, ll);
  }
}

__global__ void __launch_bounds__(kThreads)
simt_bwd_kernel(const float* __restrict__ x0, const float* __restrict__ y0,
                const float* __restrict__ x1, const float* __restrict__ y1,
                int64_t rows_local, int64_t rows_global, int64_t row_offset, int D, int n_blk,
                const float* __restrict__ scale,
                const float* __restrict__ lse_x0, const float* __restrict__ lse_y0,
                const float* __restrict__ lse_x1, const float* __restrict__ lse_y1,
                float* __restrict__ part, float* __restrict__ dscale_part) {
  extern __shared__ float4 smem4[];
  float* xs = reinterpret_cast<float*>(smem4);                 // [kRB][D]
  float* gs = xs + kRB * D;                                    // [kRB][kJC]
  const int p = blockIdx.x >= n_blk;
  const int blk = blockIdx.x - p * n_blk;
  const float* X = p ? x1 : x0;
  const float* Y = p ? y1 : y0;
  const float* lse_x = p ? lse_x1 : lse_x0;
  const float* lse_y = p ? lse_y1 : lse_y0;
  const int64_t r0 = (int64_t)blk * kRB;
  load_rows_to_smem<kRB>(xs, X, r0, rows_local, D);
  __syncthreads();
  const float s2 = scale[0] * kLog2e;
  const float lnB = logf((float)rows_global), invB = 1.0f / (float)rows_global;
  float rl2[kRB];
#pragma unroll
  for (int r = 0; r < kRB; ++r)
    rl2[r] = (r0 + r < rows_local) ? (lse_x[row_offset + r0 + r] + lnB) * kLog2e : INFINITY;
  const int nvec = D >> 2;
  float4 acc[kRB];
#pragma unroll
  for (int r = 0; r < kRB; ++r) acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
  float dsc = 0.f;
  for (int64_t j0 = 0; j0 < rows_global; j0 += kJC) {
    const int jn = (int)min((int64_t)kJC, rows_global - j0);
    // phase 1: G for this column chunk (threads over columns)
    for (int jl = threadIdx.x; jl < jn; jl += kThreads) {
      const int64_t j = j0 + jl;
      float dots[kRB];
      dot_rows<kRB>(xs, Y + j * D, D, dots);
      const float cl2 = (lse_y[j] + lnB) * kLog2e;
#pragma unroll
      for (int r = 0; r < kRB; ++r) {
        float v = dots[r] * s2;
        float gv = exp2f(v - rl2[r]) + exp2f(v - cl2);
        if (j == row_offset + r0 + r) gv -= 2.0f * invB;
        if (r0 + r >= rows_local) gv = 0.f;
        gs[r * kJC + jl] = gv;
        dsc += gv * dots[r];
      }
    }
    __syncthreads();
    // phase 2: dX += G . Y (threads over d, coalesced rows of Y)
    if ((int)threadIdx.x < nvec) {
      for (int jl = 0; jl < jn; ++jl) {
        float4 y4 = __ldg(reinterpret_cast<const float4*>(Y + (j0 + jl) * D) + threadIdx.x);
#pragma unroll
        for (int r = 0; r < kRB; ++r) {
          float gv = gs[r * kJC + jl];
          acc[r].x += gv * y4.x; acc[r].y += gv * y4.y; acc[r].z += gv * y4.z; acc[r].w += gv * y4.w;
        }
      }
    }
    __syncthreads();
  }
  if ((int)threadIdx.x < nvec) {
#pragma unroll
    for (int r = 0; r < kRB; ++r)
      if (r0 + r < rows_local)
        reinterpret_cast<float4*>(part + ((int64_t)p * rows_local + r0 + r) * D)[threadIdx.x] = acc[r];
  }
  if (p == 0) {   // sum G*cos of this block, fixed order
    __shared__ float red[kThreads / 32];
    float v = warp_sum(dsc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int w = 0; w < kThreads / 32; ++w) t += red[w];
      dscale_part[blk] = t;
    }
  }
}

int simt_infonce_fwd(const SweepArgs& a, const Workspace& ws, const SweepPlan& plan, cudaStream_t st) {
  const size_t smem = (size_t)kRB * a.D * sizeof(float);
  dim3 grid(2 * plan.n_iblk), block(kThreads);
  simt_fwd_kernel<<<grid, block, smem, st>>>(
      (const float*)a.x[0], (const float*)a.y[0], (const float*)a.x[1], (const float*)a.y[1], a.rows_local,
      a.rows_global, a.D, plan.n_iblk, a.logit_scale, a.scale_cap, reinterpret_cast<float2*>(ws.fwd_part));
  VPA_LAUNCH_CHECK("simt_fwd_kernel");
  return 0;
}

int simt_infonce_bwd(const SweepArgs& a, const Workspace& ws, const SweepPlan& plan, cudaStream_t st) {
  const size_t smem = (size_t)kRB * (a.D + kJC) * sizeof(float);
  static SmemAttrCache attr_cache;
  if (int e = ensure_dynamic_smem(attr_cache, simt_bwd_kernel, 96 * 1024)) return e;
  dim3 grid(2 * plan.n_iblk), block(kThreads);
  simt_bwd_kernel<<<grid, block, smem, st>>>(
      (const float*)a.x[0], (const float*)a.y[0], (const float*)a.x[1], (const float*)a.y[1], a.rows_local,
      a.rows_global, a.row_offset, a.D, plan.n_iblk, a.scale, a.lse_x[0], a.lse_y[0], a.lse_x[1], a.lse_y[1],
      ws.bwd_part, ws.dscale_part);
  VPA_LAUNCH_CHECK("simt_bwd_kernel");
  return 0;
}

}  // namespace vpa
