// Fused L2-normalise + cast (HBM-bound).  Replaces `x / x.norm(dim=-1, keepdim=True)`
// (reference loss_head.py:271-273, :38-40) and produces the bf16 tensor-core operands in the
// same pass.  One warp per row: the row is read ONCE with 16-byte (fp32) / 8-byte (16-bit)
// coalesced vector loads, kept in registers, reduced with shuffles, and written out.
// Algorithmic bytes per row: D*(bytes_in + bytes_out) + 4.
#include "common.cuh"

namespace vpa {

constexpr int kNormWarps = 8;      // warps (rows) per CTA
constexpr int kMaxVec = 8;         // float4 chunks per lane kept in registers -> D <= 1024

template <int DTYPE, bool IN_REGS>
__device__ __forceinline__ void normalize_row(const void* x, int64_t ld, int64_t row, int D,
                                              bool already, __nv_bfloat16* y_bf16, float* y_f32,
                                              float* inv_norm, int lane, float4 (&keep)[kMaxVec],
                                              float& inv_out) {
  const int nvec = D >> 2;
  const int64_t base = row * ld;
  float ss = 0.f;
  if constexpr (IN_REGS) {
#pragma unroll
    for (int v = 0; v < kMaxVec; ++v) {
      int c = lane + 32 * v;
      if (c < nvec) {
        keep[v] = load4<DTYPE>(x, base + 4 * c);
        ss += keep[v].x * keep[v].x + keep[v].y * keep[v].y + keep[v].z * keep[v].z + keep[v].w * keep[v].w;
      }
    }
  } else {
    for (int c = lane; c < nvec; c += 32) {
      float4 q = load4<DTYPE>(x, base + 4 * c);
      ss += q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
    }
  }
  ss = warp_sum(ss);
  const float nrm = already ? 1.0f : sqrtf(ss);
  inv_out = 1.0f / nrm;
  if (lane == 0 && inv_norm) inv_norm[row] = inv_out;
  const int64_t obase = row * (int64_t)D;
  const bool rcp = y_f32 == nullptr;      // bf16 operands only: multiply by 1/||x|| (see normalize_pair_kernel)
  auto emit = [&](int c, float4 q) {
    if (!already) {
      if (rcp) { q.x *= inv_out; q.y *= inv_out; q.z *= inv_out; q.w *= inv_out; }
      else { q.x = q.x / nrm; q.y = q.y / nrm; q.z = q.z / nrm; q.w = q.w / nrm; }      // true division, as the reference does
    }
    if (y_f32) store4<VPA_F32>(y_f32, obase + 4 * c, q);
    if (y_bf16) store4<VPA_BF16>(y_bf16, obase + 4 * c, q);
    return q;
  };
  if constexpr (IN_REGS) {
#pragma unroll
    for (int v = 0; v < kMaxVec; ++v) {
      int c = lane + 32 * v;
      if (c < nvec) keep[v] = emit(c, keep[v]);
    }
  } else {
    for (int c = lane; c < nvec; c += 32) emit(c, load4<DTYPE>(x, base + 4 * c));
  }
}

template <int DTYPE, bool IN_REGS>
__global__ void __launch_bounds__(kNormWarps * 32)
normalize_cast_kernel(const void* __restrict__ x, int64_t rows, int D, int64_t ld, int already,
                      __nv_bfloat16* __restrict__ y_bf16, float* __restrict__ y_f32,
                      float* __restrict__ inv_norm) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * kNormWarps + (threadIdx.x >> 5);
  if (row >= rows) return;
  float4 keep[kMaxVec];
  float inv;
  normalize_row<DTYPE, IN_REGS>(x, ld, row, D, already != 0, y_bf16, y_f32, inv_norm, lane, keep, inv);
}

// Both modalities + the diagonal cosine <a_i, t_i> in one launch (training path).  NV = float4 chunks per lane
// (D <= 128*NV): all 2*NV 16-byte loads of a row pair are issued before anything depends on them, and the small
// register footprint keeps >= 32 warps per SM resident -- the kernel is pure HBM streaming.
// RCP: the outputs are bf16 operands only (tensor-core mode) -> multiply by 1/||x|| instead of the reference's true
// division: the quotient is rounded to 8 mantissa bits anyway, and 2*D IEEE divisions per row pair made the kernel
// instruction-bound (ncu r01: issue slots 70 %, MUFU 67 %, 66 % of the HBM copy peak).  fp32 outputs keep `x / norm`.
template <int DTYPE, int NV, bool RCP>
__global__ void __launch_bounds__(kNormWarps * 32, NV <= 4 ? 4 : 2)
normalize_pair_kernel(const void* __restrict__ x1, const void* __restrict__ x2, int64_t rows, int D,
                      int64_t ld1, int64_t ld2, int already,
                      __nv_bfloat16* __restrict__ a_bf16, __nv_bfloat16* __restrict__ t_bf16,
                      float* __restrict__ a_f32, float* __restrict__ t_f32,
                      float* __restrict__ inv1, float* __restrict__ inv2,
                      float* __restrict__ diag_cos, int diag_from_bf16) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * kNormWarps + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nvec = D >> 2;
  float4 ka[NV], kt[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int c = lane + 32 * v;
    ka[v] = kt[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < nvec) {
      ka[v] = load4<DTYPE>(x1, row * ld1 + 4 * c);
      kt[v] = load4<DTYPE>(x2, row * ld2 + 4 * c);
    }
  }
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    s1 += ka[v].x * ka[v].x + ka[v].y * ka[v].y + ka[v].z * ka[v].z + ka[v].w * ka[v].w;
    s2 += kt[v].x * kt[v].x + kt[v].y * kt[v].y + kt[v].z * kt[v].z + kt[v].w * kt[v].w;
  }
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  const float n1 = already ? 1.0f : sqrtf(s1), n2 = already ? 1.0f : sqrtf(s2);
  const float r1 = 1.0f / n1, r2 = 1.0f / n2;
  if (lane == 0) {
    if (inv1) inv1[row] = r1;
    if (inv2) inv2[row] = r2;
  }
  float dot = 0.f;
  const int64_t obase = row * (int64_t)D;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int c = lane + 32 * v;
    if (c < nvec) {
      float4 p = ka[v], q = kt[v];
      if (!already) {
        if constexpr (RCP) {
          p.x *= r1; p.y *= r1; p.z *= r1; p.w *= r1;
          q.x *= r2; q.y *= r2; q.z *= r2; q.w *= r2;
        } else {        // true division, as the reference does (x / norm)
          p.x = p.x / n1; p.y = p.y / n1; p.z = p.z / n1; p.w = p.w / n1;
          q.x = q.x / n2; q.y = q.y / n2; q.z = q.z / n2; q.w = q.w / n2;
        }
      }
      if (a_f32) store4<VPA_F32>(a_f32, obase + 4 * c, p);
      if (t_f32) store4<VPA_F32>(t_f32, obase + 4 * c, q);
      if (a_bf16) store4<VPA_BF16>(a_bf16, obase + 4 * c, p);
      if (t_bf16) store4<VPA_BF16>(t_bf16, obase + 4 * c, q);
      if (diag_from_bf16) {  // what the tensor cores will multiply: bf16-rounded operands
        p.x = __bfloat162float(__float2bfloat16_rn(p.x)); p.y = __bfloat162float(__float2bfloat16_rn(p.y));
        p.z = __bfloat162float(__float2bfloat16_rn(p.z)); p.w = __bfloat162float(__float2bfloat16_rn(p.w));
        q.x = __bfloat162float(__float2bfloat16_rn(q.x)); q.y = __bfloat162float(__float2bfloat16_rn(q.y));
        q.z = __bfloat162float(__float2bfloat16_rn(q.z)); q.w = __bfloat162float(__float2bfloat16_rn(q.w));
      }
      dot += p.x * q.x + p.y * q.y + p.z * q.z + p.w * q.w;
    }
  }
  if (diag_cos) {
    dot = warp_sum(dot);
    if (lane == 0) diag_cos[row] = dot;
  }
}

// diag_cos[i] = <a_i, t_i> of already normalised bf16 rows (exactly the operands the tensor cores multiply): used when the
// two modalities were normalised by separate launches (pipelined host step).  One warp per row.
__global__ void __launch_bounds__(kNormWarps * 32)
diag_cos_bf16_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ t, int64_t rows, int D,
                     float* __restrict__ diag_cos) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * kNormWarps + (threadIdx.x >> 5);
  if (row >= rows) return;
  float dot = 0.f;
  for (int c = lane; c < (D >> 2); c += 32) {
    const float4 p = load4<VPA_BF16>(a, row * D + 4 * c), q = load4<VPA_BF16>(t, row * D + 4 * c);
    dot += p.x * q.x + p.y * q.y + p.z * q.z + p.w * q.w;
  }
  dot = warp_sum(dot);
  if (lane == 0) diag_cos[row] = dot;
}

// ---- several matrices in one launch (the modalities of a composite head: each is normalised ONCE, however many pairs
// it takes part in) -- blockIdx.y = matrix
struct NormMultiArgs {
  const void* x[kMaxPairs];
  int64_t ld[kMaxPairs];
  __nv_bfloat16* y[kMaxPairs];
  float* inv[kMaxPairs];
  int64_t rows;
  int D, already;
};
template <int DTYPE>
__global__ void __launch_bounds__(kNormWarps * 32) normalize_multi_kernel(const NormMultiArgs A) {
  const int lane = threadIdx.x & 31, m = blockIdx.y;
  const int64_t row = (int64_t)blockIdx.x * kNormWarps + (threadIdx.x >> 5);
  if (row >= A.rows) return;
  float4 keep[kMaxVec];
  float inv;
  normalize_row<DTYPE, true>(A.x[m], A.ld[m], row, A.D, A.already != 0, A.y[m], nullptr, A.inv[m], lane, keep, inv);
}
// diag_cos of several pairs in one launch -- blockIdx.y = pair
struct DiagMultiArgs {
  const __nv_bfloat16* a[kMaxPairs];
  const __nv_bfloat16* t[kMaxPairs];
  float* out[kMaxPairs];
  int64_t rows;
  int D;
};
__global__ void __launch_bounds__(kNormWarps * 32) diag_cos_multi_kernel(const DiagMultiArgs A) {
  const int lane = threadIdx.x & 31, p = blockIdx.y;
  const int64_t row = (int64_t)blockIdx.x * kNormWarps + (threadIdx.x >> 5);
  if (row >= A.rows) return;
  float dot = 0.f;
  for (int c = lane; c < (A.D >> 2); c += 32) {
    const float4 u = load4<VPA_BF16>(A.a[p], row * A.D + 4 * c), v = load4<VPA_BF16>(A.t[p], row * A.D + 4 * c);
    dot += u.x * v.x + u.y * v.y + u.z * v.z + u.w * v.w;
  }
  dot = warp_sum(dot);
  if (lane == 0) A.out[p][row] = dot;
}

static int check_rows(const void* x, int64_t rows, int D, int64_t ld, int in_dtype) {
  VPA_CHECK_ARG(x != nullptr, "normalize: null input");
  VPA_CHECK_ARG(rows >= 0 && D > 0 && (D % 4) == 0, "normalize: need rows >= 0, D %% 4 == 0 (D=%d)", D);
  VPA_CHECK_ARG(ld >= D && (ld % 4) == 0, "normalize: ld must be >= D and a multiple of 4");
  VPA_CHECK_ARG(in_dtype == VPA_F32 || in_dtype == VPA_BF16 || in_dtype == VPA_F16, "normalize: bad dtype %d", in_dtype);
  size_t es = in_dtype == VPA_F32 ? 4 : 2;
  VPA_CHECK_ARG((reinterpret_cast<uintptr_t>(x) % (4 * es)) == 0, "normalize: input not %zu-byte aligned", 4 * es);
  return 0;
}

int normalize_cast_launch(const void* x, int in_dtype, int64_t rows, int D, int64_t ld, int already,
                          void* y_bf16, float* y_f32, float* inv_norm, cudaStream_t st) {
  if (int e = check_rows(x, rows, D, ld, in_dtype)) return e;
  if (rows == 0) return 0;
  dim3 grid((unsigned)((rows + kNormWarps - 1) / kNormWarps)), block(kNormWarps * 32);
  const bool in_regs = D <= 128 * kMaxVec;
  auto* yb = reinterpret_cast<__nv_bfloat16*>(y_bf16);
#define VPA_NORM(DT)                                                                                  \
  if (in_regs) normalize_cast_kernel<DT, true><<<grid, block, 0, st>>>(x, rows, D, ld, already, yb, y_f32, inv_norm); \
  else normalize_cast_kernel<DT, false><<<grid, block, 0, st>>>(x, rows, D, ld, already, yb, y_f32, inv_norm);
  if (in_dtype == VPA_F32) { VPA_NORM(VPA_F32) }
  else if (in_dtype == VPA_BF16) { VPA_NORM(VPA_BF16) }
  else { VPA_NORM(VPA_F16) }
#undef VPA_NORM
  VPA_LAUNCH_CHECK("normalize_cast_kernel");
  return 0;
}

int diag_cos_bf16_launch(const void* a_bf16, const void* t_bf16, int64_t rows, int D, float* diag_cos, cudaStream_t st) {
  if (rows == 0) return 0;
  dim3 grid((unsigned)((rows + kNormWarps - 1) / kNormWarps)), block(kNormWarps * 32);
  diag_cos_bf16_kernel<<<grid, block, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(a_bf16),
                                               reinterpret_cast<const __nv_bfloat16*>(t_bf16), rows, D, diag_cos);
  VPA_LAUNCH_CHECK("diag_cos_bf16_kernel");
  return 0;
}

int normalize_multi_launch(const void* const* x, const int64_t* ld, int in_dtype, int64_t rows, int D, int n, int already,
                           void* const* y_bf16, float* const* inv, cudaStream_t st) {
  VPA_CHECK_ARG(n >= 1 && n <= kMaxPairs && D <= 128 * kMaxVec, "normalize_multi: 1..%d matrices, D <= %d", kMaxPairs, 128 * kMaxVec);
  NormMultiArgs A{};
  for (int m = 0; m < n; ++m) {
    if (int e = check_rows(x[m], rows, D, ld[m], in_dtype)) return e;
    A.x[m] = x[m]; A.ld[m] = ld[m]; A.y[m] = reinterpret_cast<__nv_bfloat16*>(y_bf16[m]); A.inv[m] = inv[m];
  }
  A.rows = rows; A.D = D; A.already = already;
  if (rows == 0) return 0;
  dim3 grid((unsigned)((rows + kNormWarps - 1) / kNormWarps), n), block(kNormWarps * 32);
  prof_begin(PROF_NORMALIZE, st);
  if (in_dtype == VPA_F32) normalize_multi_kernel<VPA_F32><<<grid, block, 0, st>>>(A);
  else if (in_dtype == VPA_BF16) normalize_multi_kernel<VPA_BF16><<<grid, block, 0, st>>>(A);
  else normalize_multi_kernel<VPA_F16><<<grid, block, 0, st>>>(A);
  prof_end(PROF_NORMALIZE, st);
  VPA_LAUNCH_CHECK("normalize_multi_kernel");
  return 0;
}

int diag_cos_multi_launch(const void* const* a_bf16, const void* const* t_bf16, float* const* out, int n, int64_t rows, int D,
                          cudaStream_t st) {
  VPA_CHECK_ARG(n >= 1 && n <= kMaxPairs, "diag_cos_multi: 1..%d pairs", kMaxPairs);
  if (rows == 0) return 0;
  DiagMultiArgs A{};
  for (int p = 0; p < n; ++p) {
    A.a[p] = reinterpret_cast<const __nv_bfloat16*>(a_bf16[p]);
    A.t[p] = reinterpret_cast<const __nv_bfloat16*>(t_bf16[p]);
    A.out[p] = out[p];
  }
  A.rows = rows; A.D = D;
  dim3 grid((unsigned)((rows + kNormWarps - 1) / kNormWarps), n), block(kNormWarps * 32);
  diag_cos_multi_kernel<<<grid, block, 0, st>>>(A);
  VPA_LAUNCH_CHECK("diag_cos_multi_kernel");
  return 0;
}

int normalize_pair_launch(const void* x1, const void* x2, int in_dtype, int64_t rows, int D, int64_t ld1,
                          int64_t ld2, int already, void* a_bf16, void* t_bf16, float* a_f32, float* t_f32,
                          float* inv1, float* inv2, float* diag_cos, int diag_from_bf16, cudaStream_t st) {
  if (int e = check_rows(x1, rows, D, ld1, in_dtype)) return e;
  if (int e = check_rows(x2, rows, D, ld2, in_dtype)) return e;
  if (rows == 0) return 0;
  if (D > 128 * kMaxVec)      // rows too long for the register-resident pair kernel: rejected BEFORE any work is enqueued
    return set_error(VPA_E_UNSUPPORTED, "normalize_pair: D=%d > %d is not supported (use vpa_normalize_cast per matrix)", D,
                     128 * kMaxVec);
  dim3 grid((unsigned)((rows + kNormWarps - 1) / kNormWarps)), block(kNormWarps * 32);
  auto* ab = reinterpret_cast<__nv_bfloat16*>(a_bf16);
  auto* tb = reinterpret_cast<__nv_bfloat16*>(t_bf16);
  const int nv = D <= 128 ? 1 : (D <= 256 ? 2 : (D <= 512 ? 4 : 8));
  prof_begin(PROF_NORMALIZE, st);
  const bool rcp = a_f32 == nullptr && t_f32 == nullptr;      // bf16 operands only
#define VPA_PAIR(DT, NV)                                                                                                  \
  if (rcp) VPA_CUDA(launch_kernel(normalize_pair_kernel<DT, NV, true>, grid, block, 0, st, x1, x2, rows, D, ld1, ld2, already, \
                                  ab, tb, a_f32, t_f32, inv1, inv2, diag_cos, diag_from_bf16));                            \
  else VPA_CUDA(launch_kernel(normalize_pair_kernel<DT, NV, false>, grid, block, 0, st, x1, x2, rows, D, ld1, ld2, already,   \
                              ab, tb, a_f32, t_f32, inv1, inv2, diag_cos, diag_from_bf16))
#define VPA_PAIR_NV(DT)                 \
  switch (nv) {                         \
    case 1: VPA_PAIR(DT, 1); break;     \
    case 2: VPA_PAIR(DT, 2); break;     \
    case 4: VPA_PAIR(DT, 4); break;     \
    default: VPA_PAIR(DT, 8); break;    \
  }
  if (in_dtype == VPA_F32) { VPA_PAIR_NV(VPA_F32) }
  else if (in_dtype == VPA_BF16) { VPA_PAIR_NV(VPA_BF16) }
  else { VPA_PAIR_NV(VPA_F16) }
#undef VPA_PAIR_NV
#undef VPA_PAIR
  prof_end(PROF_NORMALIZE, st);
  VPA_LAUNCH_CHECK("normalize_pair_kernel");
  return 0;
}

}  // namespace vpa
