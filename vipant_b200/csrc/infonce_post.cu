// Small memory-bound kernels around the two sweeps: merge of per-chunk (max, sum) statistics,
// the scalar loss, and the backward finalisation (chunk sum, scale, normalisation Jacobian, cast).
#include "common.cuh"
#include "p2p.cuh"

namespace vpa {

// ---- single-pass forward: column sums per 32-row group -> [kColSumSplit][B] (fixed order) ----------------
__global__ void __launch_bounds__(256)
colsum_reduce_kernel(const float* __restrict__ colpart, int n_groups, int64_t B, float* __restrict__ colsum,
                     const float* __restrict__ logit_scale, float scale_cap, float s2_limit) {
  if (fminf(expf(*logit_scale), scale_cap) * kLog2e > s2_limit) return;     // not the single-pass regime
  const int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (j >= B) return;
  const int per = (n_groups + kColSumSplit - 1) / kColSumSplit;
  const int g0 = blockIdx.y * per, g1 = min(n_groups, g0 + per);
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  int g = g0;
  for (; g + 4 <= g1; g += 4) {
    a0 += __ldg(colpart + (int64_t)(g + 0) * B + j);
    a1 += __ldg(colpart + (int64_t)(g + 1) * B + j);
    a2 += __ldg(colpart + (int64_t)(g + 2) * B + j);
    a3 += __ldg(colpart + (int64_t)(g + 3) * B + j);
  }
  for (; g < g1; ++g) a0 += __ldg(colpart + (int64_t)g * B + j);
  colsum[(int64_t)blockIdx.y * B + j] = (a0 + a1) + (a2 + a3);
}

// ---- forward: merge chunk partials -> logsumexp (natural log), diag = s * cos, scale_out -----------
// general regime: part = float2 (max, sum) per (problem, chunk, row), base-2 units.
// single-pass regime (fast != 0 and s*log2e <= s2_limit): part[0] = two half row sums per (chunk, row) against the
// fixed reference s2; column sums come from colsum[kColSumSplit][B] (already summed over ranks by the caller).
__global__ void combine_stats_kernel(const float2* __restrict__ part, int n_chunks, int n_chunks_fast, int64_t rows_local,
                                     int64_t rows_global, int64_t row_offset, const float* __restrict__ logit_scale,
                                     float scale_cap, const float* __restrict__ diag_cos, int fast, float s2_limit,
                                     const float* __restrict__ colsum, float* __restrict__ row_lse,
                                     float* __restrict__ col_lse, float* __restrict__ diag,
                                     float* __restrict__ scale_out) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const float e = expf(*logit_scale);
  const float s = fminf(e, scale_cap);
  if (idx == 0 && scale_out) {
    scale_out[0] = s;
    scale_out[1] = (e <= scale_cap) ? 1.0f : 0.0f;   // torch.clamp(max=) passes grad where input <= max
  }
  if (idx >= 2 * rows_local) return;
  const int p = idx >= rows_local;
  const int64_t row = idx - (int64_t)p * rows_local;
  float lse;
  if (fast && s * kLog2e <= s2_limit) {
    const float s2 = s * kLog2e;
    float L = 0.f;
    if (p == 0) {
      for (int c = 0; c < n_chunks_fast; ++c) {
        const float2 h = part[(int64_t)c * rows_local + row];
        L += h.x + h.y;
      }
    } else {
      for (int g = 0; g < kColSumSplit; ++g) L += colsum[(int64_t)g * rows_global + row_offset + row];
    }
    lse = (s2 + log2f(L)) * kLn2;
  } else {
    const float2* base = part + (int64_t)p * n_chunks * rows_local + row;
    float M = -INFINITY;
    for (int c = 0; c < n_chunks; ++c) M = fmaxf(M, base[(int64_t)c * rows_local].x);
    float L = 0.f;
    for (int c = 0; c < n_chunks; ++c) {
      float2 ml = base[(int64_t)c * rows_local];
      L += ml.y * exp2f(ml.x - M);
    }
    lse = (M + log2f(L)) * kLn2;
  }
  (p ? col_lse : row_lse)[row] = lse;
  if (p == 0 && diag) diag[row] = s * diag_cos[row];
}

// ---- row-sharded step: the per-rank message around the ONE statistics exchange -----------------------------------
// msg = [ col_sum (B) | row_lse (b) | col_lse_local (b) | diag (b) ]:  col_sum = this rank's column sums over its own
// rows (single-pass regime; zero otherwise), col_lse_local = column lse of the local rows (exact regime; zero otherwise).
struct PackArgs {
  const float2* part;
  int n_chunks, n_chunks_fast;
  int64_t b, B;
  const float* logit_scale;
  float scale_cap;
  const float* diag_cos;
  int fast;
  float s2_limit;
  const float* colsum8;
  const float* colpart;
  int n_groups;
  float* msg;               // local message buffer (NCCL / single-GPU); peer-memory transport: stores go to every rank's msgs[rank]
  P2PView pv;
  size_t off_msgs, off_msg_flags;
  uint32_t* pack_counter;
};
__device__ __forceinline__ void pack_one(const PackArgs& A, int64_t idx, float s, bool fastr) {
  const int64_t b = A.b, B = A.B;
  const float s2 = s * kLog2e;
  const bool p2p = A.pv.world > 1;          // peer-memory transport: the message goes straight into every rank's msgs[rank]
  const int64_t slot = (int64_t)A.pv.rank * (B + 3 * b);
  auto put = [&](int64_t pos, float v) {
    if (!p2p) { A.msg[pos] = v; return; }
    for (int q = 0; q < A.pv.world; ++q) reinterpret_cast<float*>(A.pv.base[q] + A.off_msgs)[slot + pos] = v;
  };
  if (idx < B) {
    float L = 0.f;
    if (fastr && A.colpart) {                 // few row groups (sharded batch): reduce the sweep's partials here, fixed order
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      int g = 0;
      for (; g + 4 <= A.n_groups; g += 4) {
        a0 += __ldg(A.colpart + (int64_t)(g + 0) * B + idx);
        a1 += __ldg(A.colpart + (int64_t)(g + 1) * B + idx);
        a2 += __ldg(A.colpart + (int64_t)(g + 2) * B + idx);
        a3 += __ldg(A.colpart + (int64_t)(g + 3) * B + idx);
      }
      for (; g < A.n_groups; ++g) a0 += __ldg(A.colpart + (int64_t)g * B + idx);
      L = (a0 + a1) + (a2 + a3);
    } else if (fastr) {
      for (int g = 0; g < kColSumSplit; ++g) L += A.colsum8[(int64_t)g * B + idx];
    }
    put(idx, L);
  }
  if (idx < b) {
    float rl, cl = 0.f;
    if (fastr) {
      float L = 0.f;
      for (int c = 0; c < A.n_chunks_fast; ++c) {
        const float2 h = A.part[(int64_t)c * b + idx];
        L += h.x + h.y;
      }
      rl = (s2 + log2f(L)) * kLn2;
    } else {
      float lse[2];
      for (int p = 0; p < 2; ++p) {
        const float2* base = A.part + (int64_t)p * A.n_chunks * b + idx;
        float M = -INFINITY;
        for (int c = 0; c < A.n_chunks; ++c) M = fmaxf(M, base[(int64_t)c * b].x);
        float L = 0.f;
        for (int c = 0; c < A.n_chunks; ++c) {
          const float2 ml = base[(int64_t)c * b];
          L += ml.y * exp2f(ml.x - M);
        }
        lse[p] = (M + log2f(L)) * kLn2;
      }
      rl = lse[0];
      cl = lse[1];
    }
    put(B + idx, rl);
    put(B + b + idx, cl);
    put(B + 2 * b + idx, s * A.diag_cos[idx]);
  }
}
// peer-memory transport: last block done -> publish the message to every rank (system-scope epoch flag)
__device__ __forceinline__ void pack_publish(const PackArgs& A) {
  __shared__ bool last;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    last = ((atomicAdd(A.pack_counter, 1u) + 1) % gridDim.x) == 0;
    __threadfence();
  }
  __syncthreads();
  if (last && (int)threadIdx.x < A.pv.world) {      // one thread per destination: the release stores travel in parallel
    st_release_sys_u32(reinterpret_cast<uint32_t*>(A.pv.base[threadIdx.x] + A.off_msg_flags) + A.pv.rank, A.pv.epoch);
  }
}
__global__ void pack_stats_kernel(const PackArgs A) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const float s = fminf(expf(*A.logit_scale), A.scale_cap);
  pack_one(A, idx, s, A.fast && s * kLog2e <= A.s2_limit);
}

// msgs[R][B + 3b] (all-gathered) -> stats_all = [row_lse (B) | col_lse (B) | diag (B)], scale_out, loss
__device__ __forceinline__ float ld_cg(const float* p) {      // L2 only: the messages may have been written by a peer GPU
  float v;
  asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
struct MergeArgs {
  const float* msgs;
  int R;
  int64_t b, B;
  const float* logit_scale;
  float scale_cap;
  int fast;
  float s2_limit;
  float *stats_all, *scale_out;
  const uint32_t* msg_flags;      // peer-memory transport: wait for every rank's message of `epoch` (nullptr otherwise)
  uint32_t epoch;
  double* loss_part;
  uint32_t* loss_counter;
  float* loss_out;
};
__device__ __forceinline__ double merge_one(const MergeArgs& A, int64_t j, float s) {
  const int64_t b = A.b, B = A.B, stride = B + 3 * b;
  const int r = (int)(j / b);
  const int64_t i = j - (int64_t)r * b;
  const float* own = A.msgs + (int64_t)r * stride;
  const float rl = ld_cg(own + B + i), dg = ld_cg(own + B + 2 * b + i);
  float cl;
  const float s2 = s * kLog2e;
  if (A.fast && s2 <= A.s2_limit) {
    float L = 0.f;
    for (int q = 0; q < A.R; ++q) L += ld_cg(A.msgs + (int64_t)q * stride + j);      // fixed rank order
    cl = (s2 + log2f(L)) * kLn2;
  } else {
    cl = ld_cg(own + B + b + i);
  }
  A.stats_all[j] = rl;
  A.stats_all[B + j] = cl;
  A.stats_all[2 * B + j] = dg;
  return ((double)rl - (double)dg) + ((double)cl - (double)dg);
}
// loss = mean(row_lse - diag) + mean(col_lse - diag): per-block partials, summed in block order by the last block to
// finish (fixed order -> bitwise deterministic and identical on every rank)
__device__ __forceinline__ void merge_loss(const MergeArgs& A, double term) {
  if (A.loss_out == nullptr) return;
  __shared__ double red[8];
  __shared__ bool last;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) term += __shfl_xor_sync(0xffffffffu, term, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = term;
  __syncthreads();
  if (threadIdx.x == 0) {
    double v = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += red[w];
    A.loss_part[blockIdx.x] = v;
    __threadfence();
    last = ((atomicAdd(A.loss_counter, 1u) + 1) % gridDim.x) == 0;
  }
  __syncthreads();
  if (last && threadIdx.x < 32) {
    __threadfence();
    double v = 0.0;
    for (unsigned i = threadIdx.x; i < gridDim.x; i += 32) v += *reinterpret_cast<volatile double*>(A.loss_part + i);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) *A.loss_out = (float)(v / (double)A.B);
  }
}
__global__ void __launch_bounds__(256) merge_stats_kernel(const MergeArgs A) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const float e = expf(*A.logit_scale);
  const float s = fminf(e, A.scale_cap);
  if (j == 0 && A.scale_out) {
    A.scale_out[0] = s;
    A.scale_out[1] = (e <= A.scale_cap) ? 1.0f : 0.0f;
  }
  merge_loss(A, j < A.B ? merge_one(A, j, s) : 0.0);
}

// Peer-memory transport: the whole statistics exchange in ONE kernel.  Every block first writes its share of this rank's
// message into all ranks' segments (the last block to finish publishes it with one system-scope epoch flag per rank), then
// waits for the R messages of this step and merges its share of the rows.  The grid is bounded (<= 2 blocks per SM, grid-
// stride loops) so that all blocks are resident at once: a block spinning for the messages never keeps a block that
// still has to write this rank's message off the machine.
constexpr int kExchangeMaxBlocks = 2 * 148;
__global__ void __launch_bounds__(256) exchange_stats_kernel(const PackArgs P, const MergeArgs A) {
  const float e = expf(*A.logit_scale);
  const float s = fminf(e, A.scale_cap);
  const bool fastr = P.fast && s * kLog2e <= P.s2_limit;
  const int64_t first = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t n = P.B > P.b ? P.B : P.b;
  for (int64_t idx = first; idx < n; idx += stride) pack_one(P, idx, s, fastr);
  pack_publish(P);
  if ((int)threadIdx.x < A.R) p2p_wait_ge(A.msg_flags + threadIdx.x, A.epoch);
  __syncthreads();
  if (first == 0 && A.scale_out) {
    A.scale_out[0] = s;
    A.scale_out[1] = (e <= A.scale_cap) ? 1.0f : 0.0f;
  }
  double term = 0.0;
  for (int64_t j = first; j < A.B; j += stride) term += merge_one(A, j, s);
  merge_loss(A, term);
}

// ---- several InfoNCE pairs in one launch (composite heads) -- blockIdx.y = pair ----------------------------------
struct PackMulti { PackArgs a[kMaxPairs]; };
struct MergeMulti { MergeArgs a[kMaxPairs]; };
__global__ void __launch_bounds__(256) pack_stats_multi_kernel(const PackMulti M) {
  const PackArgs& A = M.a[blockIdx.y];
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const float s = fminf(expf(*A.logit_scale), A.scale_cap);
  pack_one(A, idx, s, A.fast && s * kLog2e <= A.s2_limit);
}
__global__ void __launch_bounds__(256) merge_stats_multi_kernel(const MergeMulti M) {
  const MergeArgs& A = M.a[blockIdx.y];
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const float e = expf(*A.logit_scale);
  const float s = fminf(e, A.scale_cap);
  if (j == 0 && A.scale_out) {
    A.scale_out[0] = s;
    A.scale_out[1] = (e <= A.scale_cap) ? 1.0f : 0.0f;
  }
  merge_loss(A, j < A.B ? merge_one(A, j, s) : 0.0);
}

// ---- loss = mean(row_lse - diag) + mean(col_lse - diag), fixed-order fp64 reduction ---------------
__global__ void __launch_bounds__(1024)
loss_kernel(const float* __restrict__ row_lse, const float* __restrict__ col_lse,
            const float* __restrict__ diag, int64_t B, float* __restrict__ loss) {
  __shared__ double red[32];
  double acc = 0.0;
  // one block, fixed order; the loads of four consecutive strides are issued together (the kernel is pure latency)
  const bool vec = (B % 4 == 0) && ((reinterpret_cast<uintptr_t>(row_lse) | reinterpret_cast<uintptr_t>(col_lse) |
                                     reinterpret_cast<uintptr_t>(diag)) % 16 == 0);
  if (vec) {
    const int64_t n4 = B / 4;
    const float4* r4 = reinterpret_cast<const float4*>(row_lse);
    const float4* c4 = reinterpret_cast<const float4*>(col_lse);
    const float4* d4 = reinterpret_cast<const float4*>(diag);
    for (int64_t i0 = threadIdx.x; i0 < n4; i0 += 4 * 1024) {
      float4 r[4], c[4], d[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t i = i0 + u * 1024;
        if (i < n4) { r[u] = __ldg(r4 + i); c[u] = __ldg(c4 + i); d[u] = __ldg(d4 + i); }
        else { r[u] = c[u] = d[u] = make_float4(0.f, 0.f, 0.f, 0.f); }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        acc += ((double)r[u].x - (double)d[u].x) + ((double)c[u].x - (double)d[u].x);
        acc += ((double)r[u].y - (double)d[u].y) + ((double)c[u].y - (double)d[u].y);
        acc += ((double)r[u].z - (double)d[u].z) + ((double)c[u].z - (double)d[u].z);
        acc += ((double)r[u].w - (double)d[u].w) + ((double)c[u].w - (double)d[u].w);
      }
    }
  } else {
    for (int64_t i = threadIdx.x; i < B; i += 1024)
      acc += ((double)row_lse[i] - (double)diag[i]) + ((double)col_lse[i] - (double)diag[i]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = red[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) *loss = (float)(v / (double)B);
  }
}

// ---- backward finalisation -------------------------------------------------------------------------
// peer-memory transport (block 0 of the finalize kernel): {epoch, partial} into every rank's slot of this rank, then the sum
// of the R partials in rank order (bitwise identical on every rank).  Every rank's block stores before it waits: no cycle.
static __device__ __noinline__ void dls_exchange(const P2PView& pv, size_t off_dls, float part_dls, float* dlogit_scale) {
  const unsigned long long w = ((unsigned long long)pv.epoch << 32) | (unsigned long long)__float_as_uint(part_dls);
  if ((int)threadIdx.x < pv.world) {           // one thread per destination
    st_release_sys_u64(reinterpret_cast<unsigned long long*>(pv.base[threadIdx.x] + off_dls) + pv.rank, w);
  }
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    const unsigned long long* slots = reinterpret_cast<const unsigned long long*>(pv.base[pv.rank] + off_dls);
    float v = 0.f;
    if (lane < pv.world) {
      const unsigned long long t0 = global_timer_ns();
      uint32_t spins = 0;
      while (true) {
        const unsigned long long got = ld_acquire_sys_u64(slots + lane);
        if ((uint32_t)(got >> 32) == pv.epoch) { v = __uint_as_float((uint32_t)got); break; }
        if ((++spins & 1023u) == 0 && global_timer_ns() - t0 > kP2PTimeoutNs) p2p_timeout(slots + lane, pv.epoch, (uint32_t)(got >> 32));
      }
    }
    double acc = 0.0;
    for (int q = 0; q < pv.world; ++q) acc += (double)__shfl_sync(0xffffffffu, v, q);
    if (lane == 0) *dlogit_scale = (float)acc;
  }
}

constexpr int kFinWarps = 8;
constexpr int kFinMaxVec = 8;   // D <= 1024

template <int DTYPE, int NV>
__global__ void __launch_bounds__(kFinWarps * 32, 3)
finalize_bwd_kernel(const float* __restrict__ part, int n_chunks, int64_t rows_local, int D,
                    const float* __restrict__ scale /* [s, flows] */, const float* __restrict__ grad_out,
                    const void* __restrict__ x1, const void* __restrict__ x2, int64_t ld1, int64_t ld2,
                    const float* __restrict__ inv1, const float* __restrict__ inv2, int already,
                    void* __restrict__ dx1, void* __restrict__ dx2,
                    const float* __restrict__ dscale_part, int n_dscale, float* __restrict__ dlogit_scale,
                    const P2PView pv, size_t off_dls, int dls_sum) {
  const float s = scale[0], g = grad_out[0];
  const int lane = threadIdx.x & 31;
  const int64_t gw = (int64_t)blockIdx.x * kFinWarps + (threadIdx.x >> 5);   // (problem, row)
  if (gw < 2 * rows_local) {
    const int p = gw >= rows_local;
    const int64_t row = gw - (int64_t)p * rows_local;
    const float* pb = part + ((int64_t)p * n_chunks * rows_local + row) * D;
    const void* x = p ? x2 : x1;
    void* dx = p ? dx2 : dx1;
    const int64_t ld = p ? ld2 : ld1;
    const int nvec = D >> 2;
    const float sg = s * g;
    float4 v[NV], a[NV];
    float dot = 0.f;
    const float inv = already ? 1.0f : (p ? inv2 : inv1)[row];
    // all loads of the row (first chunk partial + x) are issued before anything depends on them
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c = lane + 32 * k;
      v[k] = a[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < nvec) {
        v[k] = __ldg(reinterpret_cast<const float4*>(pb) + c);
        if (!already) a[k] = load4<DTYPE>(x, row * ld + 4 * c);
      }
    }
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c = lane + 32 * k;
      if (c < nvec) {
        float4 acc = v[k];
        for (int ch = 1; ch < n_chunks; ++ch) {
          float4 q = __ldg(reinterpret_cast<const float4*>(pb + (int64_t)ch * rows_local * D) + c);
          acc.x += q.x; acc.y += q.y; acc.z += q.z; acc.w += q.w;
        }
        acc.x *= sg; acc.y *= sg; acc.z *= sg; acc.w *= sg;
        v[k] = acc;
        if (!already) {
          float4 q = a[k];
          q.x *= inv; q.y *= inv; q.z *= inv; q.w *= inv;
          a[k] = q;
          dot += q.x * acc.x + q.y * acc.y + q.z * acc.z + q.w * acc.w;
        }
      }
    }
    if (!already) dot = warp_sum(dot);
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      int c = lane + 32 * k;
      if (c < nvec) {
        float4 o = v[k];
        if (!already) {   // (I - a a^T) da / ||x||
          o.x = (o.x - a[k].x * dot) * inv; o.y = (o.y - a[k].y * dot) * inv;
          o.z = (o.z - a[k].z * dot) * inv; o.w = (o.w - a[k].w * dot) * inv;
        }
        store4<DTYPE>(dx, row * ld + 4 * c, o);
      }
    }
  }
  if (blockIdx.x == 0 && dlogit_scale) {   // d/dl: g * flows * s * sum G*cos, fixed-order fp64
    __shared__ double red[kFinWarps * 32];
    double acc = 0.0;
    for (int i = threadIdx.x; i < n_dscale; i += kFinWarps * 32) acc += (double)dscale_part[i];
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int o = kFinWarps * 16; o > 0; o >>= 1) {
      if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
      __syncthreads();
    }
    const float part_dls = (float)(red[0] * (double)s * (double)g * (double)scale[1]);
    if (pv.world > 1 && dls_sum) {
      dls_exchange(pv, off_dls, part_dls, dlogit_scale);
    } else if (threadIdx.x == 0) {
      *dlogit_scale = part_dls;
    }
  }
}

// Several pairs: the gradient of a modality is the sum over every problem that swept ITS rows (a modality shared by two
// pairs receives both contributions), then ONE normalisation Jacobian -- it is linear, so J(sum) = sum(J).
struct FinMultiArgs {
  int n_mod, n_pairs, n_chunks, D, already, n_dscale;
  int64_t rows;
  const void* x[kMaxPairs];
  void* dx[kMaxPairs];
  int64_t ld[kMaxPairs];
  const float* inv[kMaxPairs];
  int n_src[kMaxPairs];                         // problems whose X rows are this modality
  const float* part[kMaxPairs][2 * kMaxPairs];  // their dX partials [n_chunks][rows][D]
  int src_pair[kMaxPairs][2 * kMaxPairs];       // the pair each belongs to (its s and upstream gradient)
  const float* scale[kMaxPairs];                // per pair {s, flows}
  const float* grad_out;                        // [n_pairs] upstream gradients d L / d loss_p
  const float* dscale_part[kMaxPairs];
  float* dlogit_scale;                          // [n_pairs]
};
template <int DTYPE, int NV>
__global__ void __launch_bounds__(kFinWarps * 32, 3) finalize_multi_kernel(const FinMultiArgs A) {
  const int lane = threadIdx.x & 31;
  const int64_t gw = (int64_t)blockIdx.x * kFinWarps + (threadIdx.x >> 5);   // (modality, row)
  if (gw < (int64_t)A.n_mod * A.rows) {
    const int m = (int)(gw / A.rows);
    const int64_t row = gw - (int64_t)m * A.rows;
    const int nvec = A.D >> 2;
    float4 v[NV], a[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c = lane + 32 * k;
      v[k] = a[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < nvec && !A.already) a[k] = load4<DTYPE>(A.x[m], row * A.ld[m] + 4 * c);
    }
    for (int q = 0; q < A.n_src[m]; ++q) {
      const int pr = A.src_pair[m][q];
      const float sg = A.scale[pr][0] * A.grad_out[pr];
      const float* pb = A.part[m][q] + row * A.D;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int c = lane + 32 * k;
        if (c < nvec) {
          float4 acc = __ldg(reinterpret_cast<const float4*>(pb) + c);
          for (int ch = 1; ch < A.n_chunks; ++ch) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(pb + (int64_t)ch * A.rows * A.D) + c);
            acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
          }
          v[k].x = fmaf(acc.x, sg, v[k].x); v[k].y = fmaf(acc.y, sg, v[k].y);
          v[k].z = fmaf(acc.z, sg, v[k].z); v[k].w = fmaf(acc.w, sg, v[k].w);
        }
      }
    }
    const float inv = A.already ? 1.0f : A.inv[m][row];
    float dot = 0.f;
    if (!A.already) {
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        a[k].x *= inv; a[k].y *= inv; a[k].z *= inv; a[k].w *= inv;
        dot += a[k].x * v[k].x + a[k].y * v[k].y + a[k].z * v[k].z + a[k].w * v[k].w;
      }
      dot = warp_sum(dot);
    }
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c = lane + 32 * k;
      if (c < nvec) {
        float4 o = v[k];
        if (!A.already) {   // (I - a a^T) da / ||x||
          o.x = (o.x - a[k].x * dot) * inv; o.y = (o.y - a[k].y * dot) * inv;
          o.z = (o.z - a[k].z * dot) * inv; o.w = (o.w - a[k].w * dot) * inv;
        }
        store4<DTYPE>(A.dx[m], row * A.ld[m] + 4 * c, o);
      }
    }
  }
  if ((int)blockIdx.x < A.n_pairs) {        // d/dl of pair blockIdx.x: g * flows * s * sum G*cos, fixed-order fp64
    const int pr = blockIdx.x;
    __shared__ double red[kFinWarps * 32];
    double acc = 0.0;
    for (int i = threadIdx.x; i < A.n_dscale; i += kFinWarps * 32) acc += (double)A.dscale_part[pr][i];
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int o = kFinWarps * 16; o > 0; o >>= 1) {
      if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0)
      A.dlogit_scale[pr] = (float)(red[0] * (double)A.scale[pr][0] * (double)A.grad_out[pr] * (double)A.scale[pr][1]);
  }
}

int colsum_reduce_launch(const Workspace& ws, const SweepPlan& plan, int64_t rows_global, const float* logit_scale,
                         float scale_cap, float* colsum, cudaStream_t st) {
  dim3 grid((unsigned)((rows_global + 255) / 256), kColSumSplit);
  VPA_CUDA(launch_kernel(colsum_reduce_kernel, grid, dim3(256), 0, st, ws.colpart, plan.n_rowgroups, rows_global, colsum, logit_scale,
                         scale_cap, pair_fast_s2_limit()));
  VPA_LAUNCH_CHECK("colsum_reduce_kernel");
  return 0;
}

int combine_stats_launch(const Workspace& ws, const SweepPlan& plan, int64_t rows_local, int64_t rows_global,
                         int64_t row_offset, const float* logit_scale, float scale_cap, const float* diag_cos,
                         int fast, const float* colsum, float* row_lse, float* col_lse, float* diag, float* scale_out,
                         cudaStream_t st) {
  const int threads = 256;
  const int64_t n = 2 * rows_local;
  dim3 grid((unsigned)((n + threads - 1) / threads));
  VPA_CUDA(launch_kernel(combine_stats_kernel, grid, dim3(threads), 0, st, reinterpret_cast<const float2*>(ws.fwd_part),
                         plan.fwd_chunks, plan.fwd1_chunks, rows_local, rows_global, row_offset, logit_scale, scale_cap, diag_cos,
                         fast, pair_fast_s2_limit(), colsum, row_lse, col_lse, diag, scale_out));
  VPA_LAUNCH_CHECK("combine_stats_kernel");
  return 0;
}

static PackArgs make_pack_args(const Workspace& ws, const SweepPlan& plan, int64_t b, int64_t B, const float* logit_scale,
                               float scale_cap, const float* diag_cos, int fast, const float* colsum8, bool from_colpart,
                               float* msg, const P2PStep* p2p) {
  PackArgs A{};
  A.part = reinterpret_cast<const float2*>(ws.fwd_part);
  A.n_chunks = plan.fwd_chunks; A.n_chunks_fast = plan.fwd1_chunks;
  A.b = b; A.B = B;
  A.logit_scale = logit_scale; A.scale_cap = scale_cap; A.diag_cos = diag_cos;
  A.fast = fast; A.s2_limit = pair_fast_s2_limit();
  A.colsum8 = colsum8; A.colpart = from_colpart ? ws.colpart : nullptr; A.n_groups = plan.n_rowgroups;
  A.msg = msg;
  if (p2p) {
    A.pv = p2p->view; A.off_msgs = p2p->off_msgs; A.off_msg_flags = p2p->off_msg_flags; A.pack_counter = p2p->pack_counter;
  }
  return A;
}
static MergeArgs make_merge_args(const float* msgs, int R, int64_t b, int64_t B, const float* logit_scale, float scale_cap,
                                 int fast, float* stats_all, float* scale_out, const P2PStep* p2p, double* loss_part,
                                 uint32_t* loss_counter, float* loss_out) {
  MergeArgs M{};
  M.msgs = msgs; M.R = R; M.b = b; M.B = B;
  M.logit_scale = logit_scale; M.scale_cap = scale_cap; M.fast = fast; M.s2_limit = pair_fast_s2_limit();
  M.stats_all = stats_all; M.scale_out = scale_out;
  M.msg_flags = p2p ? p2p->msg_flags : nullptr; M.epoch = p2p ? p2p->view.epoch : 0u;
  M.loss_part = loss_part; M.loss_counter = loss_counter; M.loss_out = loss_out;
  return M;
}

int pack_stats_launch(const Workspace& ws, const SweepPlan& plan, int64_t b, int64_t B, const float* logit_scale,
                      float scale_cap, const float* diag_cos, int fast, const float* colsum8, bool from_colpart, float* msg,
                      cudaStream_t st) {
  const int64_t n = B > b ? B : b;
  const PackArgs A = make_pack_args(ws, plan, b, B, logit_scale, scale_cap, diag_cos, fast, colsum8, from_colpart, msg, nullptr);
  VPA_CUDA(launch_kernel(pack_stats_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, A));
  VPA_LAUNCH_CHECK("pack_stats_kernel");
  return 0;
}

int merge_stats_launch(const float* msgs, int R, int64_t b, int64_t B, const float* logit_scale, float scale_cap, int fast,
                       float* stats_all, float* scale_out, double* loss_part, uint32_t* loss_counter, float* loss_out,
                       cudaStream_t st) {
  const MergeArgs M = make_merge_args(msgs, R, b, B, logit_scale, scale_cap, fast, stats_all, scale_out, nullptr, loss_part,
                                      loss_counter, loss_out);
  VPA_CUDA(launch_kernel(merge_stats_kernel, dim3((unsigned)((B + 255) / 256)), dim3(256), 0, st, M));
  VPA_LAUNCH_CHECK("merge_stats_kernel");
  return 0;
}

// peer-memory transport: message out, R messages in, statistics of all rows + the global loss -- one launch
int exchange_stats_launch(const Workspace& ws, const SweepPlan& plan, int64_t b, int64_t B, const float* logit_scale,
                          float scale_cap, const float* diag_cos, int fast, const float* colsum8, bool from_colpart,
                          const P2PStep& p2p, float* loss_out, cudaStream_t st) {
  const PackArgs A = make_pack_args(ws, plan, b, B, logit_scale, scale_cap, diag_cos, fast, colsum8, from_colpart, nullptr, &p2p);
  const MergeArgs M = make_merge_args(p2p.msgs, p2p.view.world, b, B, logit_scale, scale_cap, fast, p2p.stats_all, p2p.scale, &p2p,
                                      p2p.loss_part, p2p.loss_counter, loss_out);
  const int64_t blocks = (B + 255) / 256;
  VPA_CUDA(launch_kernel(exchange_stats_kernel, dim3((unsigned)(blocks < kExchangeMaxBlocks ? blocks : kExchangeMaxBlocks)), dim3(256),
                         0, st, A, M));
  VPA_LAUNCH_CHECK("exchange_stats_kernel");
  return 0;
}

int loss_launch(const float* row_lse, const float* col_lse, const float* diag, int64_t B, float* loss,
                cudaStream_t st) {
  loss_kernel<<<1, 1024, 0, st>>>(row_lse, col_lse, diag, B, loss);
  VPA_LAUNCH_CHECK("loss_kernel");
  return 0;
}

int finalize_bwd_launch(const Workspace& ws, const SweepPlan& plan, int64_t rows_local, int D,
                        const float* scale, const float* grad_out, const void* x1, const void* x2,
                        int in_dtype, int64_t ld1, int64_t ld2, const float* inv1, const float* inv2,
                        int already, void* dx1, void* dx2, float* dlogit_scale, const P2PStep* p2p, int dls_sum,
                        cudaStream_t st) {
  P2PView pv{};
  if (p2p) pv = p2p->view;
  const size_t off_dls = p2p ? p2p->off_dls : 0;
  VPA_CHECK_ARG(D <= 128 * kFinMaxVec, "finalize: D=%d > %d unsupported", D, 128 * kFinMaxVec);
  const int64_t n = 2 * rows_local;
  dim3 grid((unsigned)((n + kFinWarps - 1) / kFinWarps)), block(kFinWarps * 32);
  const int nv = D <= 128 ? 1 : (D <= 256 ? 2 : (D <= 512 ? 4 : 8));
#define VPA_FIN(DT, NV)                                                                                                  \
  VPA_CUDA(launch_kernel(finalize_bwd_kernel<DT, NV>, grid, block, 0, st, ws.bwd_part, plan.bwd_chunks, rows_local, D, scale, \
                         grad_out, x1, x2, ld1, ld2, inv1, inv2, already, dx1, dx2, ws.dscale_part, plan.n_dscale,         \
                         dlogit_scale, pv, off_dls, dls_sum))
#define VPA_FIN_NV(DT)                 \
  switch (nv) {                        \
    case 1: VPA_FIN(DT, 1); break;     \
    case 2: VPA_FIN(DT, 2); break;     \
    case 4: VPA_FIN(DT, 4); break;     \
    default: VPA_FIN(DT, 8); break;    \
  }
  prof_begin(PROF_FINALIZE, st);
  if (in_dtype == VPA_F32) { VPA_FIN_NV(VPA_F32) }
  else if (in_dtype == VPA_BF16) { VPA_FIN_NV(VPA_BF16) }
  else { VPA_FIN_NV(VPA_F16) }
  prof_end(PROF_FINALIZE, st);
#undef VPA_FIN_NV
#undef VPA_FIN
  VPA_LAUNCH_CHECK("finalize_bwd_kernel");
  return 0;
}

// ---- multi-pair launchers (api.cu: vpa_infonce_multi_fwd / _bwd) ---------------------------------------------------
int pack_merge_multi_launch(int n_pairs, const Workspace* ws, const SweepPlan& plan, int64_t rows, const float* const* logit_scale,
                            const float* scale_cap, const float* const* diag_cos, float* const* msg, float* const* stats_all,
                            float* const* scale_out, double* const* loss_part, uint32_t* const* loss_counter, float* loss_out,
                            cudaStream_t st) {
  VPA_CHECK_ARG(n_pairs >= 1 && n_pairs <= kMaxPairs, "multi: 1..%d pairs", kMaxPairs);
  PackMulti P{};
  MergeMulti M{};
  for (int p = 0; p < n_pairs; ++p) {
    P.a[p] = make_pack_args(ws[p], plan, rows, rows, logit_scale[p], scale_cap[p], diag_cos[p], 1, nullptr, true, msg[p], nullptr);
    M.a[p] = make_merge_args(msg[p], 1, rows, rows, logit_scale[p], scale_cap[p], 1, stats_all[p], scale_out[p], nullptr,
                             loss_part[p], loss_counter[p], loss_out + p);
  }
  dim3 grid((unsigned)((rows + 255) / 256), n_pairs);
  pack_stats_multi_kernel<<<grid, 256, 0, st>>>(P);
  VPA_LAUNCH_CHECK("pack_stats_multi_kernel");
  merge_stats_multi_kernel<<<grid, 256, 0, st>>>(M);
  VPA_LAUNCH_CHECK("merge_stats_multi_kernel");
  return 0;
}

int finalize_multi_launch(const FinMultiHost& h, cudaStream_t st) {
  FinMultiArgs A{};
  A.n_mod = h.n_mod; A.n_pairs = h.n_pairs; A.n_chunks = h.n_chunks; A.D = h.D; A.already = h.already; A.n_dscale = h.n_dscale;
  A.rows = h.rows;
  for (int m = 0; m < h.n_mod; ++m) {
    A.x[m] = h.x[m]; A.dx[m] = h.dx[m]; A.ld[m] = h.ld[m]; A.inv[m] = h.inv[m]; A.n_src[m] = h.n_src[m];
    for (int q = 0; q < h.n_src[m]; ++q) { A.part[m][q] = h.part[m][q]; A.src_pair[m][q] = h.src_pair[m][q]; }
  }
  for (int p = 0; p < h.n_pairs; ++p) { A.scale[p] = h.scale[p]; A.dscale_part[p] = h.dscale_part[p]; }
  A.grad_out = h.grad_out; A.dlogit_scale = h.dlogit_scale;
  VPA_CHECK_ARG(h.D <= 128 * kFinMaxVec, "finalize: D=%d > %d unsupported", h.D, 128 * kFinMaxVec);
  const int64_t n = (int64_t)h.n_mod * h.rows;
  int64_t blocks = (n + kFinWarps - 1) / kFinWarps;
  if (blocks < h.n_pairs) blocks = h.n_pairs;
  dim3 grid((unsigned)blocks), block(kFinWarps * 32);
  const int nv = h.D <= 128 ? 1 : (h.D <= 256 ? 2 : (h.D <= 512 ? 4 : 8));
#define VPA_FINM(DT, NV) finalize_multi_kernel<DT, NV><<<grid, block, 0, st>>>(A)
#define VPA_FINM_NV(DT)                 \
  switch (nv) {                         \
    case 1: VPA_FINM(DT, 1); break;     \
    case 2: VPA_FINM(DT, 2); break;     \
    case 4: VPA_FINM(DT, 4); break;     \
    default: VPA_FINM(DT, 8); break;    \
  }
  prof_begin(PROF_FINALIZE, st);
  if (h.in_dtype == VPA_F32) { VPA_FINM_NV(VPA_F32) }
  else if (h.in_dtype == VPA_BF16) { VPA_FINM_NV(VPA_BF16) }
  else { VPA_FINM_NV(VPA_F16) }
  prof_end(PROF_FINALIZE, st);
#undef VPA_FINM_NV
#undef VPA_FINM
  VPA_LAUNCH_CHECK("finalize_multi_kernel");
  return 0;
}

}  // namespace vpa
