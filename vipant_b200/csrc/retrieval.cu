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

constexpr int kSimThreads = 256;
constexpr int kMaxGt = 8;

// S = Q . K^T in fp32 FFMA, register-tiled: CTA tile 128 (query rows) x 64 (key rows), k-step 16 staged transposed
// in shared memory, 8 x 4 outputs per thread (32 FFMA per 3 LDS.128), global loads of the next k-step in flight while
// the current one is multiplied.  Every output is one sequential fp32 sum over k = 0..D-1 (fixed order -> bit-stable
// run to run and independent of the grid).
constexpr int kTQ = 128, kTK = 64, kTD = 16;
constexpr int kQS = kTQ + 4, kKS = kTK + 4;      // padded strides (floats), multiples of 4 for 16-byte LDS

__global__ void __launch_bounds__(kSimThreads)
sim_store_kernel(const float* __restrict__ Q, const float* __restrict__ K, int64_t N, int64_t M, int D,
                 int64_t ldq, int64_t ldk, float* __restrict__ S) {
  __shared__ __align__(16) float qs[2][kTD][kQS];
  __shared__ __align__(16) float ks[2][kTD][kKS];
  const int tid = threadIdx.x;
  const int64_t q0 = (int64_t)blockIdx.y * kTQ, k0 = (int64_t)blockIdx.x * kTK;
  const int lr = tid >> 2, lk = (tid & 3) * 4;          // loader: row within the tile, k offset of its float4
  const int ty = tid >> 4, tx = tid & 15;               // compute: rows ty*8.., columns tx*4..
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float4 rq[2], rk;
  auto gload = [&](int d0) {
    const bool kin = d0 + lk < D;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int64_t r = q0 + lr + 64 * h;
      rq[h] = (kin && r < N) ? __ldg(reinterpret_cast<const float4*>(Q + r * ldq + d0 + lk)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const int64_t r = k0 + lr;
    rk = (kin && r < M) ? __ldg(reinterpret_cast<const float4*>(K + r * ldk + d0 + lk)) : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      qs[buf][lk + 0][lr + 64 * h] = rq[h].x; qs[buf][lk + 1][lr + 64 * h] = rq[h].y;
      qs[buf][lk + 2][lr + 64 * h] = rq[h].z; qs[buf][lk + 3][lr + 64 * h] = rq[h].w;
    }
    ks[buf][lk + 0][lr] = rk.x; ks[buf][lk + 1][lr] = rk.y; ks[buf][lk + 2][lr] = rk.z; ks[buf][lk + 3][lr] = rk.w;
  };
  const int nstep = (D + kTD - 1) / kTD;
  gload(0);
  sstore(0);
  __syncthreads();
  for (int st = 0; st < nstep; ++st) {
    const int buf = st & 1;
    if (st + 1 < nstep) gload((st + 1) * kTD);
#pragma unroll
    for (int kk = 0; kk < kTD; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&qs[buf][kk][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&qs[buf][kk][ty * 8 + 4]);
      const float4 b = *reinterpret_cast<const float4*>(&ks[buf][kk][tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    if (st + 1 < nstep) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t r = q0 + ty * 8 + i;
    if (r >= N) continue;
    const int64_t c = k0 + tx * 4;
    float* dst = S + r * M + c;
    if (c + 3 < M && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
      *reinterpret_cast<float4*>(dst) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (c + j < M) dst[j] = acc[i][j];
    }
  }
}

// (value, index) ordering of a stable descending sort: larger value first, then smaller index.
__device__ __forceinline__ bool before(float v1, int i1, float v2, int i2) {
  return v1 > v2 || (v1 == v2 && i1 < i2);
}

constexpr int kRankWarps = 4;

// One warp per query row, ONE pass over the row (16-byte loads, four in flight per lane): rank counts against the
// ground-truth values, and a warp-distributed sorted top-k list (lane j holds the j-th best so far).  An element is
// examined further only if it beats the current k-th best (a register compare); the rare survivors -- about
// k ln(M/k) per row -- are inserted one at a time with a ballot / shuffle-up step.  Ties: lower index first, i.e. the
// position in a stable descending sort.  (v1 re-read the row once per top-k slot: 11 passes for k = 10.)
__global__ void __launch_bounds__(kRankWarps * 32)
rank_topk_kernel(const float* __restrict__ S, int64_t N, int64_t M, const int32_t* __restrict__ gt, int g,
                 int k, int64_t* __restrict__ topk_idx, float* __restrict__ topk_val,
                 int32_t* __restrict__ ranks) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * kRankWarps + (threadIdx.x >> 5);
  if (row >= N) return;
  const float* s = S + row * M;
  float ref[kMaxGt];
  int gi[kMaxGt], cnt[kMaxGt];
#pragma unroll
  for (int c = 0; c < kMaxGt; ++c) {
    gi[c] = (c < g) ? gt[row * g + c] : 0;
    ref[c] = (c < g) ? s[gi[c]] : 0.f;
    cnt[c] = 0;
  }
  float lv = -INFINITY;          // this lane's list entry
  int li = 0x7fffffff;
  float tv = -INFINITY;          // current k-th best (threshold), warp-uniform
  int ti = 0x7fffffff;
  auto visit = [&](float v, int m, bool valid) {
    if (g > 0) {
#pragma unroll
      for (int c = 0; c < kMaxGt; ++c)
        if (c < g) cnt[c] += valid && (v > ref[c] || (v == ref[c] && m < gi[c]));
    }
    if (k > 0) {
      unsigned mask = __ballot_sync(0xffffffffu, valid && before(v, m, tv, ti));
      while (mask) {
        const int src = __ffs(mask) - 1;
        mask &= mask - 1;
        const float cv = __shfl_sync(0xffffffffu, v, src);
        const int ci = __shfl_sync(0xffffffffu, m, src);
        if (!before(cv, ci, tv, ti)) continue;            // the threshold moved since the ballot (warp-uniform)
        const int pos = __popc(__ballot_sync(0xffffffffu, before(lv, li, cv, ci)));   // entries ahead of the candidate
        const float uv = __shfl_up_sync(0xffffffffu, lv, 1);
        const int ui = __shfl_up_sync(0xffffffffu, li, 1);
        if (lane == pos) { lv = cv; li = ci; }
        else if (lane > pos) { lv = uv; li = ui; }
        tv = __shfl_sync(0xffffffffu, lv, k - 1);
        ti = __shfl_sync(0xffffffffu, li, k - 1);
      }
    }
  };
  const bool vec = (M % 4 == 0) && (reinterpret_cast<uintptr_t>(s) % 16 == 0);
  if (vec) {
    const float4* s4 = reinterpret_cast<const float4*>(s);
    const int n4 = (int)(M / 4);
    for (int i0 = 0; i0 < n4; i0 += 128) {
      float4 q[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * 32 + lane;
        q[u] = (i < n4) ? __ldg(s4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * 32 + lane;
        const bool ok = i < n4;
        visit(q[u].x, 4 * i, ok); visit(q[u].y, 4 * i + 1, ok); visit(q[u].z, 4 * i + 2, ok); visit(q[u].w, 4 * i + 3, ok);
      }
    }
  } else {
    for (int m0 = 0; m0 < (int)M; m0 += 32) {
      const int m = m0 + lane;
      const bool ok = m < (int)M;
      visit(ok ? s[m] : 0.f, m, ok);
    }
  }
  if (g > 0) {
#pragma unroll
    for (int c = 0; c < kMaxGt; ++c) {
      int t = cnt[c];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      if (lane == 0 && c < g) ranks[row * g + c] = t;
    }
  }
  if (lane < k) {
    if (topk_idx) topk_idx[row * k + lane] = (li == 0x7fffffff) ? -1 : (int64_t)li;
    if (topk_val) topk_val[row * k + lane] = lv;
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
  VPA_CHECK_ARG(M < (1ll << 31), "sim_rank_topk: M too large");
  VPA_CHECK_ARG(((reinterpret_cast<uintptr_t>(Q) | reinterpret_cast<uintptr_t>(K)) & 15) == 0, "sim_rank_topk: Q / K must be 16-byte aligned");
  VPA_CHECK_ARG(k >= 0 && k <= 32 && k <= M, "sim_rank_topk: need 0 <= k <= min(32, M)");
  if (N == 0) return 0;
  {
    dim3 grid((unsigned)((M + kTK - 1) / kTK), (unsigned)((N + kTQ - 1) / kTQ));
    prof_begin(PROF_SIM, st);
    sim_store_kernel<<<grid, kSimThreads, 0, st>>>(Q, K, N, M, D, ldq, ldk, S);
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
