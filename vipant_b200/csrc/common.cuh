// Shared helpers for the vipant_b200 CUDA sources (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/vipant_b200.h"

namespace vpa {

// ---- error plumbing (api.cu owns the thread-local string) --------------------------------
int set_error(int code, const char* fmt, ...);

#define VPA_CHECK_ARG(cond, ...)                                   \
  do {                                                             \
    if (!(cond)) return ::vpa::set_error(VPA_E_INVALID, __VA_ARGS__); \
  } while (0)

#define VPA_CUDA(expr)                                                                      \
  do {                                                                                      \
    cudaError_t e__ = (expr);                                                               \
    if (e__ != cudaSuccess)                                                                 \
      return ::vpa::set_error((int)e__, "%s failed: %s", #expr, cudaGetErrorString(e__));   \
  } while (0)

void note_launch();      // api.cu: every kernel launch of the library is counted (vpa_launch_count)

#define VPA_LAUNCH_CHECK(name)                                                              \
  do {                                                                                      \
    ::vpa::note_launch();                                                                   \
    cudaError_t e__ = cudaGetLastError();                                                   \
    if (e__ != cudaSuccess)                                                                 \
      return ::vpa::set_error((int)e__, "launch of %s failed: %s", name, cudaGetErrorString(e__)); \
  } while (0)

// ---- opt-in launch timing (vpa_profile_*): CUDA events recorded on the launch stream around the
// dominant kernels; used by bench.py for the roofline numbers, off by default (no cost, no state).
enum { PROF_NORMALIZE = 0, PROF_FWD_SWEEP = 1, PROF_BWD_SWEEP = 2, PROF_SIM = 3, PROF_RANK = 4, PROF_FWD_GENERAL = 5,
       PROF_FINALIZE = 6, PROF_PUSH = 7, PROF_KINDS = 8 };
void prof_begin(int kind, cudaStream_t st);
void prof_end(int kind, cudaStream_t st);

// ---- kernel launch with a typed argument list (cudaLaunchKernelEx; cluster dimensions come from __cluster_dims__) ----------
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- device helpers ----------------------------------------------------------------------
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 4 consecutive elements of a row, widened to fp32.
template <int DTYPE>
__device__ __forceinline__ float4 load4(const void* base, int64_t elem_off) {
  if constexpr (DTYPE == VPA_F32) {
    return __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + elem_off));
  } else if constexpr (DTYPE == VPA_BF16) {
    uint2 u = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(base) + elem_off));
    __nv_bfloat162 lo = *reinterpret_cast<__nv_bfloat162*>(&u.x);
    __nv_bfloat162 hi = *reinterpret_cast<__nv_bfloat162*>(&u.y);
    float2 a = __bfloat1622float2(lo), b = __bfloat1622float2(hi);
    return make_float4(a.x, a.y, b.x, b.y);
  } else {
    uint2 u = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(base) + elem_off));
    __half2 lo = *reinterpret_cast<__half2*>(&u.x);
    __half2 hi = *reinterpret_cast<__half2*>(&u.y);
    float2 a = __half22float2(lo), b = __half22float2(hi);
    return make_float4(a.x, a.y, b.x, b.y);
  }
}

template <int DTYPE>
__device__ __forceinline__ void store4(void* base, int64_t elem_off, float4 v) {
  if constexpr (DTYPE == VPA_F32) {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + elem_off) = v;
  } else if constexpr (DTYPE == VPA_BF16) {
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&lo);
    u.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(base) + elem_off) = u;
  } else {
    __half2 lo = __floats2half2_rn(v.x, v.y), hi = __floats2half2_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&lo);
    u.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(base) + elem_off) = u;
  }
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&p);
}

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a per-DEVICE setting: a process that drives several GPUs (the
// reference's `dp` mode is single-process multi-GPU) must set it on each.  One cache per call site, indexed by device.
struct SmemAttrCache { bool done[64] = {}; };
template <typename F>
inline int ensure_dynamic_smem(SmemAttrCache& cache, F func, int bytes) {
  int dev = 0;
  VPA_CUDA(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && cache.done[dev]) return 0;
  VPA_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  if (dev >= 0 && dev < 64) cache.done[dev] = true;
  return 0;
}

// ---- NCCL binding (comm.cu) ----------------------------------------------------------------
struct ncclUniqueIdBlob { char internal[128]; };
int comm_load(const char* path);
int comm_unique_id(void* out128);
int comm_init(const void* id128, int rank, int world, void** comm_out);
int comm_destroy(void* comm);
int comm_all_gather(void* comm, const void* send, void* recv, size_t count, int nccl_dtype, cudaStream_t st);
int comm_all_reduce_sum_f32(void* comm, const void* send, void* recv, size_t count, cudaStream_t st);

// ---- internal launchers (one per .cu) ----------------------------------------------------
// Work decomposition of the two sweeps (forward statistics, backward dX).  A "unit" is one CTA's
// work: a block of X rows swept against a contiguous chunk of 128-row Y tiles.
struct SweepPlan {
  int n_iblk;               // X row blocks per problem
  int rows_per_blk;         // 128 (tensor-core path) or the SIMT row block
  int n_tiles;              // ceil(rows_global / 128) column tiles (tensor-core path)
  int fwd_chunks, fwd_tiles_per_chunk;
  int bwd_chunks, bwd_tiles_per_chunk;
  int halves;               // backward: D split into 1 or 2 accumulator halves (TMEM capacity)
  int n_dscale;             // number of partial sums of G*cos written by the backward sweep
  int cluster;              // tensor-core path: CTAs per cluster sharing one multicast Y stream (1, 2 or 4)
  int impl;                 // tensor-core path: 0 = single-CTA kernels (infonce_tc.cu), 1 = CTA-pair kernels (infonce_pair.cu)
  int pair_fwd_iblk;        // pair kernels: 256-row blocks (forward) / 128-row blocks (backward); n_tiles counts 256-row tiles
  int pair_bwd_iblk;
  int fast_fwd;             // pair kernels: the single-pass forward (one S sweep, both statistics) is available
  int fwd1_chunks, fwd1_tiles_per_chunk;
  int bwd_small, fwd1_small;   // pair kernels: tiles of the short tail chunk of every row block (0: equal chunks only)
  int n_rowgroups;          // single-pass forward: 32-row groups of column partial sums (8 per 256-row block)
};
constexpr int kColSumSplit = 8;   // column sums are reduced to [kColSumSplit][rows_global] (fixed order) before any all-reduce
// reserved_sms: SMs the sweeps must leave to the relay CTAs of the peer-memory transport (0 on every other path)
SweepPlan plan_sweep(int64_t rows_local, int64_t rows_global, int D, int precision, int reserved_sms = 0);
size_t infonce_workspace_bytes(int64_t rows_local, int64_t rows_global, int D, int precision, int reserved_sms);
int relay_ctas_default();      // relay CTAs of the peer-memory transport in the forward / backward grid (api.cu)
int relay_ctas_bwd_default();

// Scratch layout shared by both precisions.
struct Workspace {
  float* fwd_part;      // [2][fwd_chunks][rows_local] float2 (m, l), base-2 units
  float* bwd_part;      // [2][bwd_chunks][rows_local][D] fp32 partial  sum_j G_ij y_j
  float* dscale_part;   // [n_dscale] partial  sum_ij G_ij cos_ij  (problem 0 only)
  float* colsum;        // [kColSumSplit][rows_global] column sums (internal copy for the unsharded one-call forward)
  float* colpart;       // single-pass forward: [n_rowgroups][rows_global] column sums per 32-row group
  size_t bytes;
};
Workspace carve_workspace(void* base, int64_t rows_local, int64_t rows_global, int D, const SweepPlan& plan);

// Arrival flags of gathered operand rows that are pushed by peer GPUs while the forward already runs (p2p.cuh).
struct P2PRowFlags {
  const uint32_t* flags;      // nullptr: no waiting (single GPU / NCCL transport)
  int rows_per_rank, chunks_per_rank, me;
  uint32_t epoch;
};

// Everything a sweep launch needs besides the plan.  Problem 0: X = A_loc, Y = T_all (rows of S);
// problem 1: X = T_loc, Y = A_all (columns of S, transposed).  Both run in ONE launch.
struct SweepArgs {
  const void* x[2];            // local rows   (bf16 for the tensor-core path, fp32 for SIMT)
  const void* y[2];            // global rows
  int64_t rows_local, rows_global, row_offset;
  int D;
  const float* logit_scale;    // forward: s = min(exp(*logit_scale), scale_cap)
  float scale_cap;
  const float* scale;          // backward: s as written by the forward
  const float* lse_x[2];       // backward: lse of the X rows' direction, GLOBAL vector (rows_global)
  const float* lse_y[2];       // backward: lse of the other direction, GLOBAL vector
  P2PRowFlags yflags;          // single-pass forward over peer memory: arrival flags of y[0]'s rows (zero: none)
  P2PRowFlags aflags;          // backward over peer memory: arrival flags of y[1]'s rows, the peers' x1 operands (zero: none)
  const struct RelayArgs* relay;   // single-pass forward / backward over peer memory: relay CTAs in front of the grid (p2p.cuh)
};

// The CTA-pair kernels sweep up to kMaxPairs InfoNCE pairs (2 * kMaxPairs problems) in ONE launch: the pairs of a composite
// head (VALCELossHead: va / lv / al, VACELossHead: vp / ap / va / vv / aa) share their operand matrices and their launches.
constexpr int kMaxPairs = 5;
struct PairProblem {           // one direction of one pair: X rows (local) swept against Y rows (global), both bf16
  const void* x;
  const void* y;
  const float* lse_x;          // backward: lse of the X rows' direction, GLOBAL vector
  const float* lse_y;          // backward: lse of the other direction
  float* out;                  // forward: float2 partials; backward: fp32 dX partials
  float* dscale;               // backward, first direction of a pair: partial sums of G * cos (nullptr otherwise)
  const float* logit_scale;    // forward
  float scale_cap;
  const float* scale;          // backward: {s, flows} of the pair
  float* colpart;              // single-pass forward: column sums per 32-row group
};
struct PairLaunch {
  PairProblem p[2 * kMaxPairs];
  int n_prob;
  int64_t rows_local, rows_global, row_offset;
  int D;
  P2PRowFlags yflags, aflags;  // peer-memory transport (single pair only)
  const struct RelayArgs* relay;
};
int pair_launch_fwd1(const PairLaunch& L, const SweepPlan& plan, cudaStream_t st);              // single-pass, device-gated
int pair_launch_fwd(const PairLaunch& L, const SweepPlan& plan, int gate, cudaStream_t st);     // exact two-sweep
int pair_launch_bwd(const PairLaunch& L, const SweepPlan& plan, cudaStream_t st);

// host-side description of the multi-pair finalisation (infonce_post.cu)
struct FinMultiHost {
  int n_mod, n_pairs, n_chunks, D, already, n_dscale, in_dtype;
  int64_t rows;
  const void* x[kMaxPairs];
  void* dx[kMaxPairs];
  int64_t ld[kMaxPairs];
  const float* inv[kMaxPairs];
  int n_src[kMaxPairs];
  const float* part[kMaxPairs][2 * kMaxPairs];
  int src_pair[kMaxPairs][2 * kMaxPairs];
  const float* scale[kMaxPairs];
  const float* grad_out;
  const float* dscale_part[kMaxPairs];
  float* dlogit_scale;
};
int finalize_multi_launch(const FinMultiHost& h, cudaStream_t st);
int pack_merge_multi_launch(int n_pairs, const Workspace* ws, const SweepPlan& plan, int64_t rows, const float* const* logit_scale,
                            const float* scale_cap, const float* const* diag_cos, float* const* msg, float* const* stats_all,
                            float* const* scale_out, double* const* loss_part, uint32_t* const* loss_counter, float* loss_out,
                            cudaStream_t st);
int normalize_multi_launch(const void* const* x, const int64_t* ld, int in_dtype, int64_t rows, int D, int n, int already,
                           void* const* y_bf16, float* const* inv, cudaStream_t st);
int diag_cos_multi_launch(const void* const* a_bf16, const void* const* t_bf16, float* const* out, int n, int64_t rows, int D,
                          cudaStream_t st);

int tc_infonce_fwd(const SweepArgs& a, const Workspace& ws, const SweepPlan& plan, cudaStream_t st);
int tc_infonce_bwd(const SweepArgs& a, const Workspace& ws, const SweepPlan& plan, cudaStream_t st);
int pair_infonce_fwd(const SweepArgs& a, const Workspace& ws, const SweepPlan& plan, int which, cudaStream_t st);
float pair_fast_s2_limit();
int pair_infonce_bwd(const SweepArgs& a, const Workspace& ws, const SweepPlan& plan, cudaStream_t st);
int simt_infonce_fwd(const SweepArgs& a, const Workspace& ws, const SweepPlan& plan, cudaStream_t st);
int simt_infonce_bwd(const SweepArgs& a, const Workspace& ws, const SweepPlan& plan, cudaStream_t st);

}  // namespace vpa
