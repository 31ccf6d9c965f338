// Fused similarity + rank / top-1 for the monitors' scoring: the similarity tile is consumed IN REGISTERS, S is never
// written.  Replaces `S = x1s @ x2s.t(); S.argsort(descending=True); torch.where(ind == gt)` of the reference
// (loss_head.py:115-117, 128-130 N==M; :139-142, 156-158 1-vs-5; :81-103 retrieval_eval; :381-385 zero-shot argmax) for
// BOTH directions from ONE pass over S: the rank of a ground-truth column is a count over a row of S, the rank of a
// ground-truth row (the T->A direction, whose similarity matrix is S^T) a count over a column of the same tile --
// fmaf(a, b, c) == fmaf(b, a, c), so the tile values ARE the other direction's values bit for bit.  report() therefore
// needs 2*N*M*D flops instead of 4*N*M*D, and no 19 MB of S travel to L2 and back.
//
//   sim_gt_ref_kernel   the similarity of every (query, ground truth) pair, summed in exactly the order the tile kernel
//                       uses (one fp32 fmaf chain over k = 0..D-1) so that a ground-truth column compares EQUAL to itself
//                       in the tile; also clears the rank counters / top-1 keys.  One thread per pair.
//   sim_rank_tile_kernel  S tile = 128 x 64 (or 64 x 64 for the tail wave, see the launcher) in registers, 8 x 4 per thread,
//                       fp32 FFMA; epilogue: per-row counts of (v > ref) | (v == ref & col < gt) reduced over the 16 lanes
//                       that share a row -> one atomicAdd per (row, gt); per-column counts through shared-memory atomics;
//                       top-1 as a 64-bit (orderable value, ~index) atomicMax ("larger value, then lower index").
//   sim_top1_decode_kernel  keys -> int64 index + fp32 value.
// Integer atomics: the results do not depend on the order of arrival.  Similarities are one sequential fp32 sum each:
// bit-stable run to run and independent of the tiling.
#include <algorithm>

#include "common.cuh"

namespace vpa {

constexpr int kFT = 256;                        // threads per CTA
constexpr int kFTQ = 128, kFTK = 64, kFTD = 16; // tile: query rows, key rows, k step
constexpr int kFQS = kFTQ + 4, kFKS = kFTK + 4; // padded strides (floats)
constexpr int kFMaxGt = 8;

// monotone map fp32 -> uint32 that agrees with the fp32 comparisons used for the ranks: -0 and +0 get the same key, every
// NaN (either sign) sorts above +inf -- torch.argsort(descending=True) places NaN first
__device__ __forceinline__ uint32_t orderable_u32(float v) {
  if (v != v) return 0xffffffffu;
  const uint32_t u = __float_as_uint(v + 0.0f);      // -0 + 0 = +0
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float orderable_to_float(uint32_t o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}
// (value, index) -> key whose unsigned order is "larger value first, then LOWER index"
__device__ __forceinline__ unsigned long long top1_key(float v, int idx) {
  return ((unsigned long long)orderable_u32(v) << 32) | (unsigned long long)(0xffffffffu - (uint32_t)idx);
}

struct FusedArgs {
  const float *Q, *K;
  int64_t N, M, ldq, ldk;
  int D;
  const int32_t *gt_q, *gt_k;       // (N, g_q) key indices / (M, g_k) query indices; nullptr when g == 0
  int g_q, g_k;
  float *ref_q, *ref_k;             // (N, g_q) / (M, g_k): similarity of each (row, ground truth) pair
  int32_t *ranks_q, *ranks_k;       // counters, cleared by the reference kernel
  unsigned long long *key_q, *key_k;   // top-1 keys (nullptr: not wanted)
  int n_big, tiles_k;               // grid: n_big 128-row tiles first, then pairs of 64-row half tiles (in-order dispatch)
};

// ---- similarity of the ground-truth pairs, clears ----------------------------------------------------------------
// One THREAD per pair: the sum must be the tile kernel's -- one fp32 fmaf chain over k = 0..D-1 -- so a pair cannot be split
// over lanes; the loads do not depend on the chain and are issued ahead of it (unrolled).  (A first version walked the
// accumulator through the 32 lanes of a warp per pair: 31 of 32 lanes idle per step, 18.6 us for 9750 pairs under ncu.)
__global__ void __launch_bounds__(128) sim_gt_ref_kernel(const FusedArgs A) {
  const int64_t nq = A.N * A.g_q, nk = A.M * A.g_k;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = tid; i < nq; i += nthr) A.ranks_q[i] = 0;
  for (int64_t i = tid; i < nk; i += nthr) A.ranks_k[i] = 0;
  if (A.key_q) for (int64_t i = tid; i < A.N; i += nthr) A.key_q[i] = 0ull;
  if (A.key_k) for (int64_t i = tid; i < A.M; i += nthr) A.key_k[i] = 0ull;
  for (int64_t w = tid; w < nq + nk; w += nthr) {
    int64_t qrow, krow;
    float* out;
    if (w < nq) {
      qrow = w / A.g_q;
      krow = A.gt_q[w];
      out = A.ref_q + w;
      if (krow < 0 || krow >= A.M) { *out = __int_as_float(0x7fc00000); continue; }      // invalid index: NaN, never counted
    } else {
      const int64_t v = w - nq;
      krow = v / A.g_k;
      qrow = A.gt_k[v];
      out = A.ref_k + v;
      if (qrow < 0 || qrow >= A.N) { *out = __int_as_float(0x7fc00000); continue; }
    }
    const float4* q = reinterpret_cast<const float4*>(A.Q + qrow * A.ldq);
    const float4* k = reinterpret_cast<const float4*>(A.K + krow * A.ldk);
    float acc = 0.f;
    const int n4 = A.D >> 2;
#pragma unroll 8
    for (int d = 0; d < n4; ++d) {
      const float4 a = __ldg(q + d), b = __ldg(k + d);
      acc = fmaf(a.x, b.x, acc);
      acc = fmaf(a.y, b.y, acc);
      acc = fmaf(a.z, b.z, acc);
      acc = fmaf(a.w, b.w, acc);
    }
    *out = acc;
  }
}

// ---- the tile kernel ---------------------------------------------------------------------------------------------
// RQ = rows per thread (8: 128-row tile, 4: 64-row half tile).  Thread (ty, tx) of 16 x 16 owns rows ty*RQ.. and
// columns tx*4..; every output is ONE sequential fp32 sum over k.
template <int RQ>
__device__ __forceinline__ void sim_rank_tile(const FusedArgs& A, int q0, int k0, float* smem) {
  constexpr int TQ = 16 * RQ;
  constexpr int QS = TQ + 4;
  float (*qs)[kFTD][QS] = reinterpret_cast<float (*)[kFTD][QS]>(smem);                       // [2][kFTD][QS]
  float (*ks)[kFTD][kFKS] = reinterpret_cast<float (*)[kFTD][kFKS]>(smem + 2 * kFTD * QS);   // [2][kFTD][kFKS]
  float* refq_s = smem + 2 * kFTD * QS + 2 * kFTD * kFKS;                                    // [TQ][g_q]
  int* gtq_s = reinterpret_cast<int*>(refq_s + kFTQ * kFMaxGt);                              // [TQ][g_q]
  float* refk_s = reinterpret_cast<float*>(gtq_s + kFTQ * kFMaxGt);                          // [kFTK][g_k]
  int* gtk_s = reinterpret_cast<int*>(refk_s + kFTK * kFMaxGt);                              // [kFTK][g_k]
  int* cntk_s = gtk_s + kFTK * kFMaxGt;                                                      // [kFTK][g_k]
  unsigned long long* keyk_s = reinterpret_cast<unsigned long long*>(cntk_s + kFTK * kFMaxGt);   // [kFTK]
  const int tid = threadIdx.x;
  const int lr = tid >> 2, lk = (tid & 3) * 4;          // loader: row within the tile, k offset of its float4
  const int ty = tid >> 4, tx = tid & 15;
  const int D = A.D, N = (int)A.N, M = (int)A.M;      // (N, M < 2^31: checked by the launcher)
  // ground-truth references of this tile's rows / columns -> shared memory
  for (int i = tid; i < TQ * A.g_q; i += kFT) {
    const int r = q0 + i / A.g_q;
    refq_s[i] = r < N ? A.ref_q[(int64_t)r * A.g_q + i % A.g_q] : 0.f;
    gtq_s[i] = r < N ? A.gt_q[(int64_t)r * A.g_q + i % A.g_q] : 0;
  }
  for (int i = tid; i < kFTK * A.g_k; i += kFT) {
    const int c = k0 + i / A.g_k;
    refk_s[i] = c < M ? A.ref_k[(int64_t)c * A.g_k + i % A.g_k] : 0.f;
    gtk_s[i] = c < M ? A.gt_k[(int64_t)c * A.g_k + i % A.g_k] : 0;
    cntk_s[i] = 0;
  }
  if (A.key_k && tid < kFTK) keyk_s[tid] = 0ull;
  float acc[RQ][4];
#pragma unroll
  for (int i = 0; i < RQ; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  constexpr int QL = TQ / 64;                            // float4 loads of Q per thread and k step (2 or 1)
  float4 rq[QL], rk;
  auto gload = [&](int d0) {
    const bool kin = d0 + lk < D;
#pragma unroll
    for (int h = 0; h < QL; ++h) {
      const int r = q0 + lr + 64 * h;
      rq[h] = (kin && r < N) ? __ldg(reinterpret_cast<const float4*>(A.Q + (int64_t)r * A.ldq + d0 + lk)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const int r = k0 + lr;
    rk = (kin && r < M) ? __ldg(reinterpret_cast<const float4*>(A.K + (int64_t)r * A.ldk + d0 + lk)) : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int h = 0; h < QL; ++h) {
      qs[buf][lk + 0][lr + 64 * h] = rq[h].x; qs[buf][lk + 1][lr + 64 * h] = rq[h].y;
      qs[buf][lk + 2][lr + 64 * h] = rq[h].z; qs[buf][lk + 3][lr + 64 * h] = rq[h].w;
    }
    ks[buf][lk + 0][lr] = rk.x; ks[buf][lk + 1][lr] = rk.y; ks[buf][lk + 2][lr] = rk.z; ks[buf][lk + 3][lr] = rk.w;
  };
  const int nstep = (D + kFTD - 1) / kFTD;
  gload(0);
  sstore(0);
  __syncthreads();
  for (int st = 0; st < nstep; ++st) {
    const int buf = st & 1;
    if (st + 1 < nstep) gload((st + 1) * kFTD);
#pragma unroll
    for (int kk = 0; kk < kFTD; ++kk) {
      float a[RQ];
#pragma unroll
      for (int i4 = 0; i4 < RQ / 4; ++i4) {
        const float4 v = *reinterpret_cast<const float4*>(&qs[buf][kk][ty * RQ + 4 * i4]);
        a[4 * i4] = v.x; a[4 * i4 + 1] = v.y; a[4 * i4 + 2] = v.z; a[4 * i4 + 3] = v.w;
      }
      const float4 b = *reinterpret_cast<const float4*>(&ks[buf][kk][tx * 4]);
      const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < RQ; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    if (st + 1 < nstep) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }
  // ---- epilogue: the tile is consumed in registers
  const int col0 = k0 + tx * 4;
  // rows of S: ranks of the ground-truth columns and the row maximum
#pragma unroll
  for (int i = 0; i < RQ; ++i) {
    const int rl = ty * RQ + i;
    const int row = q0 + rl;
    const bool row_ok = row < N;
    for (int c = 0; c < A.g_q; ++c) {
      const float ref = refq_s[rl * A.g_q + c];
      const int gi = gtq_s[rl * A.g_q + c];
      int cnt = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = col0 + j;
        const float v = acc[i][j];
        cnt += (row_ok && col < M && (v > ref || (v == ref && col < gi))) ? 1 : 0;
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);      // the 16 lanes that share this row
      if (tx == 0 && row_ok && cnt) atomicAdd(A.ranks_q + (int64_t)row * A.g_q + c, cnt);
    }
    if (A.key_q) {
      unsigned long long best = 0ull;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (col0 + j < M) best = max(best, top1_key(acc[i][j], col0 + j));
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
      if (tx == 0 && row_ok) atomicMax(A.key_q + row, best);
    }
  }
  // columns of S (= rows of the transposed similarity): ranks of the ground-truth rows and the column maximum
  if (A.g_k > 0 || A.key_k) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int cl = tx * 4 + j;
      const bool col_ok = col0 + j < M;
      for (int c = 0; c < A.g_k; ++c) {
        const float ref = refk_s[cl * A.g_k + c];
        const int gi = gtk_s[cl * A.g_k + c];
        int cnt = 0;
#pragma unroll
        for (int i = 0; i < RQ; ++i) {
          const int row = q0 + ty * RQ + i;
          const float v = acc[i][j];
          cnt += (col_ok && row < N && (v > ref || (v == ref && row < gi))) ? 1 : 0;
        }
        if (cnt) atomicAdd(cntk_s + cl * A.g_k + c, cnt);
      }
      if (A.key_k && col_ok) {
        unsigned long long best = 0ull;
#pragma unroll
        for (int i = 0; i < RQ; ++i)
          if (q0 + ty * RQ + i < N) best = max(best, top1_key(acc[i][j], q0 + ty * RQ + i));
        atomicMax(keyk_s + cl, best);
      }
    }
    __syncthreads();
    for (int i = tid; i < kFTK * A.g_k; i += kFT) {
      const int c = k0 + i / A.g_k;
      if (c < M && cntk_s[i]) atomicAdd(A.ranks_k + (int64_t)c * A.g_k + i % A.g_k, cntk_s[i]);
    }
    if (A.key_k && tid < kFTK && k0 + tid < M) atomicMax(A.key_k + k0 + tid, keyk_s[tid]);
  }
}

constexpr size_t kFusedSmemFloats = 2 * kFTD * kFQS + 2 * kFTD * kFKS + 2 * kFTQ * kFMaxGt + 3 * kFTK * kFMaxGt + 2 * kFTK;

// (2 CTAs per SM without spills measured the same as 3 with 68 B of spills: 201 vs 198 us at 975 x 4875.)
__global__ void __launch_bounds__(kFT, 3) sim_rank_tile_kernel(const FusedArgs A) {
  extern __shared__ __align__(16) float fused_smem[];
  const int u = blockIdx.x;
  if (u < A.n_big) {                       // 128-row tiles, column-tile major within a row band
    sim_rank_tile<8>(A, (u / A.tiles_k) * kFTQ, (u % A.tiles_k) * kFTK, fused_smem);
  } else {                                 // tail wave: the remaining 128-row tiles as two 64-row halves each
    const int h = u - A.n_big, t = A.n_big + (h >> 1);
    sim_rank_tile<4>(A, (t / A.tiles_k) * kFTQ + (h & 1) * 64, (t % A.tiles_k) * kFTK, fused_smem);
  }
}

__global__ void sim_top1_decode_kernel(const unsigned long long* __restrict__ keys, int64_t n, int64_t* __restrict__ idx,
                                       float* __restrict__ val) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long k = keys[i];
  if (idx) idx[i] = (int64_t)(0xffffffffu - (uint32_t)(k & 0xffffffffull));
  if (val) val[i] = orderable_to_float((uint32_t)(k >> 32));
}

struct FusedWs {
  float *ref_q, *ref_k;
  unsigned long long *key_q, *key_k;
  size_t bytes;
};
static FusedWs carve_fused(void* base, int64_t N, int64_t M, int g_q, int g_k) {
  FusedWs w{};
  size_t o = 0;
  auto take = [&](size_t bytes) {
    void* p = base ? static_cast<char*>(base) + o : nullptr;
    o += align_up(bytes ? bytes : 1, 256);
    return p;
  };
  w.ref_q = static_cast<float*>(take((size_t)N * g_q * 4));
  w.ref_k = static_cast<float*>(take((size_t)M * g_k * 4));
  w.key_q = static_cast<unsigned long long*>(take((size_t)N * 8));
  w.key_k = static_cast<unsigned long long*>(take((size_t)M * 8));
  w.bytes = o;
  return w;
}
size_t sim_fused_workspace_bytes(int64_t N, int64_t M, int g_q, int g_k) {
  if (N <= 0 || M <= 0 || g_q < 0 || g_k < 0) return 0;
  return carve_fused(nullptr, N, M, g_q, g_k).bytes;
}

int sim_rank_fused_launch(const float* Q, const float* K, int64_t N, int64_t M, int D, int64_t ldq, int64_t ldk,
                          const int32_t* gt_q, int g_q, const int32_t* gt_k, int g_k, int32_t* ranks_q, int32_t* ranks_k,
                          int64_t* top1_q, float* top1_val_q, int64_t* top1_k, float* top1_val_k, void* workspace,
                          size_t workspace_bytes, cudaStream_t st) {
  VPA_CHECK_ARG(Q && K && workspace, "sim_rank_fused: null pointer");
  VPA_CHECK_ARG(N >= 0 && M > 0 && D > 0 && (D % 4) == 0 && D <= 4096, "sim_rank_fused: bad shape N=%lld M=%lld D=%d",
                (long long)N, (long long)M, D);
  VPA_CHECK_ARG(ldq >= D && ldk >= D && (ldq % 4) == 0 && (ldk % 4) == 0, "sim_rank_fused: bad leading dimension");
  VPA_CHECK_ARG(g_q >= 0 && g_q <= kFMaxGt && (g_q == 0 || (gt_q && ranks_q)), "sim_rank_fused: need 0 <= g_q <= %d (+ gt_q, ranks_q)", kFMaxGt);
  VPA_CHECK_ARG(g_k >= 0 && g_k <= kFMaxGt && (g_k == 0 || (gt_k && ranks_k)), "sim_rank_fused: need 0 <= g_k <= %d (+ gt_k, ranks_k)", kFMaxGt);
  VPA_CHECK_ARG(M < (1ll << 31) && N < (1ll << 31), "sim_rank_fused: N / M too large");
  VPA_CHECK_ARG(((reinterpret_cast<uintptr_t>(Q) | reinterpret_cast<uintptr_t>(K)) & 15) == 0, "sim_rank_fused: Q / K must be 16-byte aligned");
  const FusedWs w = carve_fused(workspace, N, M, g_q, g_k);
  if (w.bytes > workspace_bytes) return set_error(VPA_E_WORKSPACE, "sim_rank_fused: workspace %zu < %zu", workspace_bytes, w.bytes);
  if (N == 0) return 0;
  const bool want_q = top1_q || top1_val_q, want_k = top1_k || top1_val_k;
  FusedArgs A{};
  A.Q = Q; A.K = K; A.N = N; A.M = M; A.ldq = ldq; A.ldk = ldk; A.D = D;
  A.gt_q = gt_q; A.gt_k = gt_k; A.g_q = g_q; A.g_k = g_k;
  A.ref_q = w.ref_q; A.ref_k = w.ref_k; A.ranks_q = ranks_q; A.ranks_k = ranks_k;
  A.key_q = want_q ? w.key_q : nullptr;
  A.key_k = want_k ? w.key_k : nullptr;
  // Tiling: T tiles of 128 x 64 on S = 3 CTAs per SM.  Whole waves run 128-row tiles; when the last wave would be less
  // than half full its tiles are split into 64-row halves, which shortens it to half a tile time (616 tiles on 444 slots at
  // 975 x 4875: 1.5 instead of 2 tile times).  CTAs are dispatched in blockIdx order: big tiles first.
  int dev = 0, sms = 148;
  VPA_CUDA(cudaGetDevice(&dev));
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) { sms = 148; cudaGetLastError(); }
  const int64_t tiles_q = (N + kFTQ - 1) / kFTQ, tiles_k = (M + kFTK - 1) / kFTK, T = tiles_q * tiles_k;
  VPA_CHECK_ARG(T < (1ll << 30), "sim_rank_fused: too many tiles");
  const int64_t slots = 3ll * sms, rem = T % slots;
  const int64_t split = (rem * 2 <= slots) ? rem : 0;
  A.n_big = (int)(T - split);
  A.tiles_k = (int)tiles_k;
  {
    const int64_t work = std::max<int64_t>(std::max(N * std::max(g_q, 1), M * std::max(g_k, 1)), N * g_q + M * g_k);
    int64_t blocks = std::min<int64_t>((work + 127) / 128, 16ll * sms);
    if (blocks < 1) blocks = 1;
    prof_begin(PROF_SIM, st);
    sim_gt_ref_kernel<<<(unsigned)blocks, 128, 0, st>>>(A);
    VPA_LAUNCH_CHECK("sim_gt_ref_kernel");
  }
  const size_t smem = kFusedSmemFloats * sizeof(float);
  sim_rank_tile_kernel<<<(unsigned)(A.n_big + 2 * split), kFT, smem, st>>>(A);
  prof_end(PROF_SIM, st);
  VPA_LAUNCH_CHECK("sim_rank_tile_kernel");
  if (want_q) {
    sim_top1_decode_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(w.key_q, N, top1_q, top1_val_q);
    VPA_LAUNCH_CHECK("sim_top1_decode_kernel");
  }
  if (want_k) {
    sim_top1_decode_kernel<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(w.key_k, M, top1_k, top1_val_k);
    VPA_LAUNCH_CHECK("sim_top1_decode_kernel");
  }
  return 0;
}

}  // namespace vpa
