// Multi-label ranking metrics of the AudioSet zero-shot / tagging evaluation on the GPU (SURVEY.md section 8f row 3 = Z2).
// Replaces the scikit-learn calls of the reference's BCELossHead.report (cvap/module/decoder/loss_more.py:92-123), which run
// per class on the host after `x1s.cpu().numpy()`:
//     metrics.average_precision_score(y, s, average=None | 'micro')   :92-94, :101
//     metrics.roc_auc_score(y, s, average=None)                       :106
//     metrics.precision_recall_curve(y, s) -> the MIDDLE point        :110-115
// (macro / weighted AP are means of the per-class values: host side).  Algorithm = scikit-learn's `_binary_clf_curve`
// (oracle/map_oracle.py restates it): sort by descending score, cumulative true positives at the LAST index of every
// distinct score ("threshold"), then sums over the thresholds.
//
//   ap_class_kernel   one CTA per class: the class's N scores (as order-preserving uint32 keys) and labels are bitonic-sorted
//                     in shared memory (N <= 32768: 160 KB), then every thread walks a contiguous chunk of the sorted order
//                     carrying (position, tp) of the previous threshold: AP = sum (R_t - R_{t-1}) P_t, ROC-AUC by the
//                     trapezoid rule from the (0, 0) origin, and the middle point of the PR curve (with scikit-learn 1.0.1's
//                     truncation at full recall, or without it: `truncate`).  fp64 sums, fixed order.  Also writes the sorted
//                     keys and cumulative positives of the class for the micro average.
//   micro_ap_kernel   micro AP = AP of the N*C pooled pairs = (1 / P) sum over positives of TP(>= s) / ALL(>= s): for every
//                     positive, the counts over all classes come from binary searches in the sorted per-class arrays (lanes
//                     split the classes) -- no global sort of N*C keys.  Per-warp partial sums, fixed-order final reduction.
#include "common.cuh"

namespace vpa {

constexpr int kApThreads = 1024;
constexpr int kApMaxN = 32768;

// monotone fp32 -> uint32; 0 is below every real key.  -0 and +0 map to ONE key: numpy counts them as the same score, and a
// threshold is a boundary between distinct scores.
__device__ __forceinline__ uint32_t ap_orderable(float v) {
  const uint32_t u = __float_as_uint(v + 0.0f);      // -0 + 0 = +0
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

struct ApArgs {
  const float* S;
  const void* Y;
  int64_t ld_s, ld_y;
  int y_dtype;                // VPA_F32 or VPA_U8
  int N, C, n_pad, truncate;
  double* per_class;          // [C][4]: AP, ROC-AUC, precision and recall at the middle of the PR curve
  int32_t* flags;             // [C]: bit 0 no positive (AP undefined), bit 1 no negative (AUC undefined)
  int32_t* support;           // [C]: positives
  uint32_t* skeys;            // [C][N] sorted keys (descending)
  uint32_t* ctp;              // [C][N] inclusive cumulative positives in that order
};

// block-wide exclusive scan of one int per thread (1024 threads); returns the exclusive prefix, *total = the sum
__device__ __forceinline__ int block_exclusive_scan(int v, int* warp_tot /* [32] */, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int n = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += n;
  }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = warp_tot[lane], wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int n = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += n;
    }
    warp_tot[lane] = wi - w;                     // exclusive prefix of the warp totals
    if (lane == 31) *total = wi;
  }
  __syncthreads();
  const int r = warp_tot[warp] + incl - v;
  __syncthreads();                               // warp_tot may be reused by the caller
  return r;
}
// fixed-order block sum of one double per thread
__device__ __forceinline__ double block_sum_f64(double v, double* red /* [32] */) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x < 32) {
    t = red[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  }
  __syncthreads();
  return t;                                      // valid in thread 0
}

__global__ void __launch_bounds__(kApThreads, 1) ap_class_kernel(const ApArgs A) {
  extern __shared__ __align__(16) uint8_t ap_smem[];
  uint32_t* keys = reinterpret_cast<uint32_t*>(ap_smem);
  uint8_t* lab = ap_smem + (size_t)A.n_pad * 4;
  int* wtot = reinterpret_cast<int*>(lab + A.n_pad);             // [32]
  int* lastpos = wtot + 32;                                      // [1024]
  int* lasttp = lastpos + kApThreads;                            // [1024]
  double* red = reinterpret_cast<double*>(lasttp + kApThreads);  // [32]
  int* misc = reinterpret_cast<int*>(red + 32);                  // [8]
  const int c = blockIdx.x, tid = threadIdx.x, N = A.N, n = A.n_pad;
  for (int i = tid; i < n; i += kApThreads) {
    uint32_t k = 0;
    uint8_t y = 0;
    if (i < N) {
      k = ap_orderable(__ldg(A.S + (int64_t)i * A.ld_s + c));
      if (A.y_dtype == VPA_F32) y = __ldg(reinterpret_cast<const float*>(A.Y) + (int64_t)i * A.ld_y + c) == 1.0f;
      else y = __ldg(reinterpret_cast<const uint8_t*>(A.Y) + (int64_t)i * A.ld_y + c) == 1;
    }
    keys[i] = k;
    lab[i] = y;
  }
  if (tid == 0) misc[0] = 0x7fffffff;
  __syncthreads();
  // ---- bitonic sort, descending by key (ties in any order: every quantity below is taken at distinct-score boundaries)
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < (n >> 1); t += kApThreads) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), l = i | j;
        const uint32_t a = keys[i], b = keys[l];
        const bool desc = (i & k) == 0;
        if (desc ? (a < b) : (a > b)) {
          keys[i] = b; keys[l] = a;
          const uint8_t ya = lab[i];
          lab[i] = lab[l]; lab[l] = ya;
        }
      }
      __syncthreads();
    }
  }
  // ---- pass A: positives and thresholds of this thread's chunk
  const int per = n / kApThreads, lo = tid * per, hi = min(lo + per, N);
  int npos = 0, nthr = 0, lp = -1, ltp = 0;
  for (int i = lo; i < hi; ++i) {
    npos += lab[i];
    if (i == N - 1 || keys[i] != keys[i + 1]) { ++nthr; lp = i; ltp = npos; }
  }
  int P = 0, T = 0;
  const int tp_base = block_exclusive_scan(npos, wtot, &misc[1]);
  const int ord_base = block_exclusive_scan(nthr, wtot, &misc[2]);
  P = misc[1];
  T = misc[2];
  lastpos[tid] = lp;
  lasttp[tid] = tp_base + ltp;
  __syncthreads();
  int prev_pos = -1, prev_tp = 0;                  // the threshold before this chunk (the origin when there is none)
  for (int t = tid - 1; t >= 0; --t)
    if (lastpos[t] >= 0) { prev_pos = lastpos[t]; prev_tp = lasttp[t]; break; }
  const int Nn = N - P;
  const double invP = P > 0 ? 1.0 / (double)P : 0.0, invN = Nn > 0 ? 1.0 / (double)Nn : 0.0;
  // ---- pass B: the sums over the thresholds; sorted keys / cumulative positives out
  double ap = 0.0, auc = 0.0;
  int tp = tp_base, ord = ord_base, first_full = 0x7fffffff;
  uint32_t* gk = A.skeys + (int64_t)c * N;
  uint32_t* gc = A.ctp + (int64_t)c * N;
  for (int i = lo; i < hi; ++i) {
    tp += lab[i];
    gk[i] = keys[i];
    gc[i] = (uint32_t)tp;
    if (i == N - 1 || keys[i] != keys[i + 1]) {
      const double prec = (double)tp / (double)(i + 1);
      ap += (double)(tp - prev_tp) * invP * prec;
      const double fpr = (double)(i + 1 - tp) * invN, tpr = (double)tp * invP;
      const double fpr0 = (double)(prev_pos + 1 - prev_tp) * invN, tpr0 = (double)prev_tp * invP;
      auc += (fpr - fpr0) * (tpr + tpr0) * 0.5;
      if (tp == P && first_full == 0x7fffffff) first_full = ord;
      prev_pos = i; prev_tp = tp;
      ++ord;
    }
  }
  if (first_full != 0x7fffffff) atomicMin(&misc[0], first_full);
  const double ap_tot = block_sum_f64(ap, red);
  const double auc_tot = block_sum_f64(auc, red);   // (its barriers also order the atomicMin above)
  // ---- middle point of the precision-recall curve (reversed thresholds + the (1, 0) end point)
  const int L = A.truncate ? misc[0] : T - 1;       // last curve index before the appended point
  const int mid = (L + 2) / 2;
  if (tid == 0) {
    double* out = A.per_class + (int64_t)c * 4;
    out[0] = P > 0 ? ap_tot : __longlong_as_double(0x7ff8000000000000ll);
    out[1] = (P > 0 && Nn > 0) ? auc_tot : __longlong_as_double(0x7ff8000000000000ll);
    out[2] = 1.0;                                   // mid beyond the curve: the appended (precision 1, recall 0)
    out[3] = 0.0;
    A.flags[c] = (P == 0 ? 1 : 0) | (Nn == 0 ? 2 : 0);
    A.support[c] = P;
  }
  __syncthreads();
  if (mid <= L) {
    const int target = L - mid;
    if (target >= ord_base && target < ord_base + nthr) {      // this thread's chunk holds that threshold
      int tq = tp_base, o = ord_base;
      for (int i = lo; i < hi; ++i) {
        tq += lab[i];
        if (i == N - 1 || keys[i] != keys[i + 1]) {
          if (o == target) {
            double* out = A.per_class + (int64_t)c * 4;
            out[2] = (double)tq / (double)(i + 1);
            out[3] = P > 0 ? (double)tq * invP : 1.0;
            break;
          }
          ++o;
        }
      }
    }
  }
}

// ---- micro average over the pooled (sample, class) pairs ------------------------------------------------------------
__global__ void __launch_bounds__(256) micro_ap_kernel(const uint32_t* __restrict__ skeys, const uint32_t* __restrict__ ctp,
                                                       int N, int C, int warps_per_class, double* __restrict__ part) {
  const int lane = threadIdx.x & 31;
  const int64_t wg = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (wg >= (int64_t)C * warps_per_class) return;
  const int c = (int)(wg / warps_per_class), i = (int)(wg % warps_per_class) * 32 + lane;
  const bool in = i < N;
  const uint32_t mine = in ? ctp[(int64_t)c * N + i] : 0u;
  const uint32_t before = (in && i > 0) ? ctp[(int64_t)c * N + i - 1] : 0u;
  const uint32_t key = in ? skeys[(int64_t)c * N + i] : 0u;
  unsigned mask = __ballot_sync(0xffffffffu, in && mine != before);
  double sum = 0.0;
  while (mask) {
    const int src = __ffs(mask) - 1;
    mask &= mask - 1;
    const uint32_t s = __shfl_sync(0xffffffffu, key, src);
    unsigned all = 0, pos = 0;
    for (int cc = lane; cc < C; cc += 32) {
      const uint32_t* K = skeys + (int64_t)cc * N;
      int lo = 0, hi = N;                       // first index whose key is < s (the array is descending)
      while (lo < hi) {
        const int m = (lo + hi) >> 1;
        if (__ldg(K + m) >= s) lo = m + 1; else hi = m;
      }
      all += (unsigned)lo;
      if (lo) pos += __ldg(ctp + (int64_t)cc * N + lo - 1);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      all += __shfl_xor_sync(0xffffffffu, all, o);
      pos += __shfl_xor_sync(0xffffffffu, pos, o);
    }
    sum += (double)pos / (double)all;           // the precision at this positive's threshold
  }
  if (lane == 0) part[wg] = sum;
}

__global__ void __launch_bounds__(1024) micro_ap_finish_kernel(const double* __restrict__ part, int64_t n,
                                                               const int32_t* __restrict__ support, int C,
                                                               double* __restrict__ micro_out) {
  __shared__ double red[32];
  double v = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += 1024) v += part[i];
  const double tot = block_sum_f64(v, red);
  if (threadIdx.x == 0) {
    long long P = 0;
    for (int c = 0; c < C; ++c) P += support[c];
    *micro_out = P > 0 ? tot / (double)P : __longlong_as_double(0x7ff8000000000000ll);
  }
}

struct ApWs {
  uint32_t *skeys, *ctp;
  double* part;
  int64_t n_part;
  size_t bytes;
};
static ApWs carve_ap(void* base, int64_t N, int C) {
  ApWs w{};
  size_t o = 0;
  auto take = [&](size_t bytes) {
    void* p = base ? static_cast<char*>(base) + o : nullptr;
    o += align_up(bytes ? bytes : 1, 256);
    return p;
  };
  w.skeys = static_cast<uint32_t*>(take((size_t)N * C * 4));
  w.ctp = static_cast<uint32_t*>(take((size_t)N * C * 4));
  w.n_part = (int64_t)C * ((N + 31) / 32);
  w.part = static_cast<double*>(take((size_t)w.n_part * 8));
  w.bytes = o;
  return w;
}
size_t multilabel_workspace_bytes(int64_t N, int C) {
  if (N <= 0 || C <= 0) return 0;
  return carve_ap(nullptr, N, C).bytes;
}

int multilabel_scores_launch(const float* S, int64_t ld_s, const void* Y, int y_dtype, int64_t ld_y, int64_t N, int C,
                             int truncate_pr, double* per_class, int32_t* flags, int32_t* support, double* micro_ap,
                             void* workspace, size_t workspace_bytes, cudaStream_t st) {
  VPA_CHECK_ARG(S && Y && per_class && flags && support && workspace, "multilabel_scores: null pointer");
  VPA_CHECK_ARG(N >= 1 && C >= 1 && ld_s >= C && ld_y >= C, "multilabel_scores: bad shape N=%lld C=%d", (long long)N, C);
  VPA_CHECK_ARG(y_dtype == VPA_F32 || y_dtype == VPA_U8, "multilabel_scores: labels must be fp32 or uint8");
  if (N > kApMaxN) return set_error(VPA_E_UNSUPPORTED, "multilabel_scores: N=%lld > %d samples per call", (long long)N, kApMaxN);
  const ApWs w = carve_ap(workspace, N, C);
  if (w.bytes > workspace_bytes) return set_error(VPA_E_WORKSPACE, "multilabel_scores: workspace %zu < %zu", workspace_bytes, w.bytes);
  ApArgs A{};
  A.S = S; A.Y = Y; A.ld_s = ld_s; A.ld_y = ld_y; A.y_dtype = y_dtype;
  A.N = (int)N; A.C = C; A.truncate = truncate_pr ? 1 : 0;
  int n_pad = kApThreads;
  while (n_pad < N) n_pad <<= 1;
  A.n_pad = n_pad;
  A.per_class = per_class; A.flags = flags; A.support = support; A.skeys = w.skeys; A.ctp = w.ctp;
  const size_t smem = (size_t)n_pad * 5 + (32 + 2 * kApThreads) * sizeof(int) + 32 * sizeof(double) + 8 * sizeof(int);
  static SmemAttrCache attr_cache;
  if (int e = ensure_dynamic_smem(attr_cache, ap_class_kernel, (int)((size_t)kApMaxN * 5 + 16384))) return e;
  ap_class_kernel<<<C, kApThreads, smem, st>>>(A);
  VPA_LAUNCH_CHECK("ap_class_kernel");
  if (micro_ap) {
    const int wpc = (int)((N + 31) / 32);
    micro_ap_kernel<<<(unsigned)((w.n_part + 7) / 8), 256, 0, st>>>(w.skeys, w.ctp, (int)N, C, wpc, w.part);
    VPA_LAUNCH_CHECK("micro_ap_kernel");
    micro_ap_finish_kernel<<<1, 1024, 0, st>>>(w.part, w.n_part, support, C, micro_ap);
    VPA_LAUNCH_CHECK("micro_ap_finish_kernel");
  }
  return 0;
}

}  // namespace vpa
