// Similarity + warp-level rank / top-k for the monitors' scoring (retrieval R@k, zero-shot argmax).
// Replaces `S = Q @ K.t(); ind = S.argsort(descending=True); torch.where(ind == gt)[1]; ind[:, :k]`
// (reference loss_head.py:115-117, 139-142, 156-158, 81-103, 381-385): a full argsort of every row
// is never needed -- the rank of a ground-truth column is a COUNT of larger similarities and the
// top-k (k <= 32) is k warp-argmax passes over a row that stays in L1/L2.
//   sim_store_kernel : fp32 FFMA dot products (bit-stable summation order), S -> workspace
//   rank_topk_kernel : one warp per query row; HBM/L2-bound over S (4*N*M bytes read)
#include "common.cuh"
#include "simt_dot.cuh"

namespace vpa {

constexpr int kSimRB = 8;
constexpr int kSimThreads = 256;
constexpr int kMaxGt = 8;

__global__ void __launch_bounds__(kSimThreads)
sim_store_kernel(const float* __restrict__ Q, const float* __restrict__ K, int64_t N, int64_t M, int D,
                 int64_t ldq, int64_t ldk, float* __restrict__ S) {
  extern __shared__ float4 smem4[];
  float* xs = reinterpret_cast<float*>(smem4);
  const int64_t r0 = (int64_t)blockIdx.x * kSimRB;
  const int nvec = D >> 2;
  float4* xs4 = reinterpret_cast<float4*>(xs);
  for (int i = threadIdx.x; i < kSimRB * nvec; i += blockDim.x) {
    const int r = i / nvec, c = i - r * nvec;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r0 + r < N) v = __ldg(reinterpret_cast<const float4*>(Q + (r0 + r) * ldq) + c);
    xs4[i] = v;
  }
  __syncthreads();
  for (int64_t j = (int64_t)blockIdx.y * kSimThreads + threadIdx.x; j < M; j += (int64_t)gridDim.y * kSimThreads) {
    float dots[kSimRB];
    dot_rows<kSimRB>(xs, K + j * ldk, D, dots);
#pragma unroll
    for (int r = 0; r < kSimRB; ++r)
      if (r0 + r < N) S[(r0 + r) * M + j] = dots[r];
  }
}

// (value, index) ordering of a stable descending sort: larger value first, then smaller index.
__device__ __forceinline__ bool before(float v1, int64_t i1, float v2, int64_t i2) {
  return v1 > v2 || (v1 == v2 && i1 < i2);
}

constexpr int kRankWarps = 4;

__global__ void __launch_bounds__(kRankWarps * 32)
rank_topk_kernel(const float* __restrict__ S, int64_t N, int64_t M, const int32_t* __restrict__ gt, int g,
                 int k, int64_t* __restrict__ topk_idx, float* __restrict__ topk_val,
                 int32_t* __restrict__ ranks) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * kRankWarps + (threadIdx.x >> 5);
  if (row >= N) return;
  const float* s = S + row * M;
  if (g > 0) {
    float ref[kMaxGt];
    int64_t gi[kMaxGt];
    int cnt[kMaxGt];
#pragma unroll
    for (int c = 0; c < kMaxGt; ++c) {
      gi[c] = (c < g) ? (int64_t)gt[row * g + c] : 0;
      ref[c] = (c < g) ? s[gi[c]] : 0.f;
      cnt[c] = 0;
    }
    for (int64_t m = lane; m < M; m += 32) {
      const float v = s[m];
#pragma unroll
      for (int c = 0; c < kMaxGt; ++c) cnt[c] += (c < g) && before(v, m, ref[c], gi[c]);
    }
#pragma unroll
    for (int c = 0; c < kMaxGt; ++c) {
      int t = cnt[c];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      if (lane == 0 && c < g) ranks[row * g + c] = t;
    }
  }
  float pv = INFINITY;
  int64_t pi = -1;
  for (int t = 0; t < k; ++t) {
    float bv = -INFINITY;
    int64_t bi = M;          // sentinel: nothing found
    for (int64_t m = lane; m < M; m += 32) {
      const float v = s[m];
      if (before(pv, pi, v, m) && (bi == M || before(v, m, bv, bi))) { bv = v; bi = m; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      int64_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (oi != M && (bi == M || before(ov, oi, bv, bi))) { bv = ov; bi = oi; }
    }
    if (lane == 0) {
      if (topk_idx) topk_idx[row * k + t] = (bi == M) ? -1 : bi;
      if (topk_val) topk_val[row * k + t] = bv;
    }
    pv = bv;
    pi = bi;
  }
}

int sim_rank_topk_launch(const float* Q, const float* K, int64_t N, int64_t M, int D, int64_t ldq, int64_t ldk,
                         const int32_t* gt_idx, int g, int k, int64_t* topk_idx, float* topk_val,
                         int32_t* ranks, float* S, cudaStream_t st) {
  VPA_CHECK_ARG(Q && K && S, "sim_rank_topk: null pointer");
  VPA_CHECK_ARG(N >= 0 && M > 0 && D > 0 && (D % 4) == 0 && D <= 2048, "sim_rank_topk: bad shape N=%lld M=%lld D=%d",
                (long long)N, (long long)M, D);
  VPA_CHECK_ARG(ldq >= D && ldk >= D && (ldq % 4) == 0 && (ldk % 4) == 0, "sim_rank_topk: bad leading dimension");
  VPA_CHECK_ARG(g >= 0 && g <= kMaxGt && (g == 0 || (gt_idx && ranks)), "sim_rank_topk: need 0 <= g <= %d", kMaxGt);
  VPA_CHECK_ARG(k >= 0 && k <= 32 && k <= M, "sim_rank_topk: need 0 <= k <= min(32, M)");
  if (N == 0) return 0;
  {
    const size_t smem = (size_t)kSimRB * D * sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
      VPA_CUDA(cudaFuncSetAttribute(sim_store_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
      attr_set = true;
    }
    const unsigned gx = (unsigned)((N + kSimRB - 1) / kSimRB);
    // enough column slabs to fill the GPU when there are few query rows
    unsigned gy = (unsigned)((M + kSimThreads - 1) / kSimThreads);
    const unsigned want = (unsigned)((4 * 148 + gx - 1) / gx);
    if (gy > want) gy = want;
    if (gy < 1) gy = 1;
    prof_begin(PROF_SIM, st);
    sim_store_kernel<<<dim3(gx, gy), kSimThreads, smem, st>>>(Q, K, N, M, D, ldq, ldk, S);
    prof_end(PROF_SIM, st);
    VPA_LAUNCH_CHECK("sim_store_kernel");
  }
  if (g > 0 || k > 0) {
    const unsigned gx = (unsigned)((N + kRankWarps - 1) / kRankWarps);
    prof_begin(PROF_RANK, st);
    rank_topk_kernel<<<gx, kRankWarps * 32, 0, st>>>(S, N, M, gt_idx, g, k, topk_idx, topk_val, ranks);
    prof_end(PROF_RANK, st);
    VPA_LAUNCH_CHECK("rank_topk_kernel");
  }
  return 0;
}

}  // namespace vpa
