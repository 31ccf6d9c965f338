// extern "C" surface of libvipant_b200.so (see include/vipant_b200.h) + work planning.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <functional>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#include "common.cuh"
#include "p2p.cuh"

namespace vpa {

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

static unsigned long long g_launches = 0;      // kernel launches of this library since it was loaded (not thread-safe: a statistic)
void note_launch() { ++g_launches; }

// ---- opt-in launch timing ---------------------------------------------------------------------------
constexpr int kProfSlots = 512;
struct ProfState {
  bool on = false;
  bool hold = false;      // enabled but paused: launches are not bracketed (vpa_profile_hold)
  cudaEvent_t ev[PROF_KINDS][kProfSlots][2];
  bool made[PROF_KINDS][kProfSlots];
  int n[PROF_KINDS];
};
static ProfState g_prof;

void prof_begin(int kind, cudaStream_t st) {
  if (!g_prof.on || g_prof.hold || g_prof.n[kind] >= kProfSlots) return;
  const int i = g_prof.n[kind];
  if (!g_prof.made[kind][i]) {
    if (cudaEventCreate(&g_prof.ev[kind][i][0]) != cudaSuccess || cudaEventCreate(&g_prof.ev[kind][i][1]) != cudaSuccess) return;
    g_prof.made[kind][i] = true;
  }
  cudaEventRecord(g_prof.ev[kind][i][0], st);
}
void prof_end(int kind, cudaStream_t st) {
  if (!g_prof.on || g_prof.hold || g_prof.n[kind] >= kProfSlots) return;
  const int i = g_prof.n[kind];
  if (!g_prof.made[kind][i]) return;
  cudaEventRecord(g_prof.ev[kind][i][1], st);
  g_prof.n[kind] = i + 1;
}

// launchers defined in the other translation units
int normalize_cast_launch(const void* x, int in_dtype, int64_t rows, int D, int64_t ld, int already,
                          void* y_bf16, float* y_f32, float* inv_norm, cudaStream_t st);
int normalize_pair_launch(const void* x1, const void* x2, int in_dtype, int64_t rows, int D, int64_t ld1,
                          int64_t ld2, int already, void* a_bf16, void* t_bf16, float* a_f32, float* t_f32,
                          float* inv1, float* inv2, float* diag_cos, int diag_from_bf16, cudaStream_t st);
int colsum_reduce_launch(const Workspace& ws, const SweepPlan& plan, int64_t rows_global, const float* logit_scale,
                         float scale_cap, float* colsum, cudaStream_t st);
int combine_stats_launch(const Workspace& ws, const SweepPlan& plan, int64_t rows_local, int64_t rows_global,
                         int64_t row_offset, const float* logit_scale, float scale_cap, const float* diag_cos,
                         int fast, const float* colsum, float* row_lse, float* col_lse, float* diag, float* scale_out,
                         cudaStream_t st);
int pack_stats_launch(const Workspace& ws, const SweepPlan& plan, int64_t b, int64_t B, const float* logit_scale,
                      float scale_cap, const float* diag_cos, int fast, const float* colsum8, bool from_colpart, float* msg,
                      cudaStream_t st);
int merge_stats_launch(const float* msgs, int R, int64_t b, int64_t B, const float* logit_scale, float scale_cap, int fast,
                       float* stats_all, float* scale_out, double* loss_part, uint32_t* loss_counter, float* loss_out,
                       cudaStream_t st);
int exchange_stats_launch(const Workspace& ws, const SweepPlan& plan, int64_t b, int64_t B, const float* logit_scale,
                          float scale_cap, const float* diag_cos, int fast, const float* colsum8, bool from_colpart,
                          const P2PStep& p2p, float* loss_out, cudaStream_t st);
int loss_launch(const float* row_lse, const float* col_lse, const float* diag, int64_t B, float* loss,
                cudaStream_t st);
int finalize_bwd_launch(const Workspace& ws, const SweepPlan& plan, int64_t rows_local, int D,
                        const float* scale, const float* grad_out, const void* x1, const void* x2,
                        int in_dtype, int64_t ld1, int64_t ld2, const float* inv1, const float* inv2,
                        int already, void* dx1, void* dx2, float* dlogit_scale, const P2PStep* p2p, int dls_sum,
                        cudaStream_t st);
int diag_cos_bf16_launch(const void* a_bf16, const void* t_bf16, int64_t rows, int D, float* diag_cos, cudaStream_t st);
int sim_rank_topk_launch(const float* Q, const float* K, int64_t N, int64_t M, int D, int64_t ldq, int64_t ldk,
                         const int32_t* gt_idx, int g, int k, int64_t* topk_idx, float* topk_val,
                         int32_t* ranks, float* S, cudaStream_t st);

size_t sim_fused_workspace_bytes(int64_t N, int64_t M, int g_q, int g_k);
int sim_rank_fused_launch(const float* Q, const float* K, int64_t N, int64_t M, int D, int64_t ldq, int64_t ldk,
                          const int32_t* gt_q, int g_q, const int32_t* gt_k, int g_k, int32_t* ranks_q, int32_t* ranks_k,
                          int64_t* top1_q, float* top1_val_q, int64_t* top1_k, float* top1_val_k, void* workspace,
                          size_t workspace_bytes, cudaStream_t st);

size_t multilabel_workspace_bytes(int64_t N, int C);
int multilabel_scores_launch(const float* S, int64_t ld_s, const void* Y, int y_dtype, int64_t ld_y, int64_t N, int C,
                             int truncate_pr, double* per_class, int32_t* flags, int32_t* support, double* micro_ap,
                             void* workspace, size_t workspace_bytes, cudaStream_t st);

int encoder_tail_launch(const void* x, int in_dtype, int64_t rows, int width, int64_t ld, const float* gamma, const float* beta,
                        float eps, const void* proj_t_bf16, int N, void* ln_bf16, float* mean, float* rstd, void* a_bf16,
                        float* y_f32, float* inv_norm, cudaStream_t st);

static int sm_count() {      // of the CURRENT device (a process may drive several)
  static int cached[64] = {};
  int dev = 0, v = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 148; }      // CPU-only box (plan queries): B200
  if (dev >= 0 && dev < 64 && cached[dev]) return cached[dev];
  if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) { cudaGetLastError(); return 148; }
  if (dev >= 0 && dev < 64) cached[dev] = v;
  return v;
}

// Relay CTAs of the peer-memory transport (whole CTA pairs in front of a sweep kernel's grid, p2p.cuh).  The forward cannot
// finish before its x2 operands have arrived, so it gives the transfer 20 SMs (measured at 8 GPUs: 12 / 20 / 28 relay CTAs
// move 243 / 305 / 335 GB/s into one GPU; the sweep wants the other SMs).  The backward runs 3-4x longer than the transfer
// of its x1 operands and consumes them at <= 75 GB/s: 8 relay CTAs keep ahead of it (measured at 8 GPUs: backward sweep
// 0.623 / 0.405 / 0.389 ms with 2 / 4 / 8 relay CTAs; 0.364 ms without any exchange).
int relay_ctas_default() {
  static int v = 0;
  if (v == 0) {
    v = 20;
    if (const char* e = getenv("VPA_P2P_RELAY_CTAS")) { const int q = atoi(e); if (q >= 2 && q <= 64) v = q & ~1; }   // tuning knob
  }
  return v;
}
int relay_ctas_bwd_default() {
  static int v = 0;
  if (v == 0) {
    v = 8;
    if (const char* e = getenv("VPA_P2P_RELAY_CTAS_BWD")) { const int q = atoi(e); if (q >= 2 && q <= 64) v = q & ~1; }   // tuning knob
  }
  return v;
}

// Number of column chunks that minimises (waves x per-unit length) for `base_units` row-block units
// sweeping `n_tiles` tiles; `fixed` = per-unit fixed cost in tile equivalents (X load, drain).
static void pick_chunks(int base_units, int n_tiles, int fixed, int* chunks, int* tiles_per_chunk, int sms_per_unit = 1,
                        int reserved_sms = 0) {
  int sms = (sm_count() - reserved_sms) / sms_per_unit;
  if (sms < 1) sms = 1;
  long best_cost = -1;
  int best_c = 1;
  const int cmax = n_tiles < 64 ? n_tiles : 64;
  for (int c = 1; c <= cmax; ++c) {
    const int tpc = (n_tiles + c - 1) / c;
    const int real_c = (n_tiles + tpc - 1) / tpc;
    const long waves = ((long)base_units * real_c + sms - 1) / sms;
    const long cost = waves * (tpc + fixed);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_c = real_c; }
  }
  *tiles_per_chunk = (n_tiles + best_c - 1) / best_c;
  *chunks = (n_tiles + *tiles_per_chunk - 1) / *tiles_per_chunk;
}

// Pair kernels: split every row block's n_tiles into k equal chunks (+ optionally one short tail) so that the makespan on the
// SM pairs is minimal.  The grid lists all big units first, then the tails, and the hardware hands the next unit to the
// first free slot: simulated here exactly (in-order list scheduling).  `fixed` = per-unit cost in tile equivalents.
// Equal chunking alone wastes up to a whole wave when units do not divide the slots (64 backward units on 74 slots).
// `relay_slots` of the SM pairs are taken by relay CTAs (peer-memory transport) for the first `relay_tiles` tile times of the
// kernel and join the pool afterwards: the simulation starts them busy instead of leaving them out.
static void pick_split(int base_units, int n_tiles, int fixed, int* chunks, int* tiles_per_chunk, int* small_tiles,
                       int relay_slots = 0, int relay_tiles = 0) {
  int slots = sm_count() / 2;
  if (relay_slots > slots - 1) relay_slots = slots - 1;
  if (slots < 1) slots = 1;
  long best = -1;
  int best_k = 1, best_small = 0;
  const int kmax = n_tiles < 16 ? n_tiles : 16;
  std::vector<long> heap;
  for (int k = 1; k <= kmax; ++k) {
    for (int small = 0; small < n_tiles && small <= 48; ++small) {
      const int big_tiles = n_tiles - small;
      const int tpc = (big_tiles + k - 1) / k;
      if ((long)(k - 1) * tpc >= big_tiles) continue;                 // an empty chunk
      if (small > 0 && small >= tpc) break;                             // the tail must be the short one
      heap.assign(slots, 0);                                            // min-heap of the slots' free times
      for (int r = 0; r < relay_slots; ++r) heap[r] = relay_tiles;
      std::make_heap(heap.begin(), heap.end(), std::greater<long>());
      long makespan = 0;
      auto place = [&](long cost) {
        std::pop_heap(heap.begin(), heap.end(), std::greater<long>());
        heap.back() += cost;
        if (heap.back() > makespan) makespan = heap.back();
        std::push_heap(heap.begin(), heap.end(), std::greater<long>());
      };
      const int last = big_tiles - (k - 1) * tpc;                       // the last big chunk may be shorter
      for (int c = 0; c < k; ++c)
        for (int u = 0; u < base_units; ++u) place((c == k - 1 ? last : tpc) + fixed);
      if (small > 0)
        for (int u = 0; u < base_units; ++u) place(small + fixed);
      if (best < 0 || makespan < best) { best = makespan; best_k = k; best_small = small; }
    }
  }
  const int big_tiles = n_tiles - best_small;
  *tiles_per_chunk = (big_tiles + best_k - 1) / best_k;
  *chunks = best_k + (best_small > 0 ? 1 : 0);
  *small_tiles = best_small;
}

static SweepPlan plan_sweep_uncached(int64_t rows_local, int64_t rows_global, int D, int precision, int reserved_sms);

// The split search costs milliseconds: plans are computed once per shape.
SweepPlan plan_sweep(int64_t rows_local, int64_t rows_global, int D, int precision, int reserved_sms) {
  static std::mutex mu;
  static std::map<std::tuple<int64_t, int64_t, int, int, int, int>, SweepPlan> cache;
  const auto key = std::make_tuple(rows_local, rows_global, D, precision, sm_count(), reserved_sms);
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;
  const SweepPlan p = plan_sweep_uncached(rows_local, rows_global, D, precision, reserved_sms);
  if (cache.size() > 4096) cache.clear();
  cache.emplace(key, p);
  return p;
}

static SweepPlan plan_sweep_uncached(int64_t rows_local, int64_t rows_global, int D, int precision, int reserved_sms) {
  SweepPlan p{};
  if (precision == VPA_PREC_FP32_SIMT) {
    p.rows_per_blk = 8;
    p.n_iblk = (int)((rows_local + 7) / 8);
    p.n_tiles = 0;
    p.fwd_chunks = p.bwd_chunks = p.fwd1_chunks = 1;
    p.fwd_tiles_per_chunk = p.bwd_tiles_per_chunk = 0;
    p.halves = 1;
    p.n_dscale = p.n_iblk;
    return p;
  }
  p.rows_per_blk = 128;
  p.impl = (D == 256 || D == 512) ? 1 : 0;
  if (p.impl == 1) {
    p.cluster = 2;
    p.halves = 1;
    p.n_iblk = (int)((rows_local + 127) / 128);
    p.pair_fwd_iblk = (int)((rows_local + 255) / 256);
    p.pair_bwd_iblk = (int)((rows_local + 127) / 128);
    p.n_tiles = (int)((rows_global + 255) / 256);
    pick_chunks(2 * p.pair_fwd_iblk, p.n_tiles, 1, &p.fwd_chunks, &p.fwd_tiles_per_chunk, 2);
    // per-unit fixed cost (prologue, X load, pipeline fill, dX drain) measured at ~3.5 tiles of 256 columns (b=4096 sweep)
    {
      // backward over peer memory: relay CTA pairs hold their SMs until the peers' x1 operands are in -- in tile times of
      // this kernel: bytes / (per-CTA pull rate, ~25 GB/s with few relays) / (time of one 128 x 256 tile pair of MMAs)
      const int bwd_relay = reserved_sms > 0 ? relay_ctas_bwd_default() : 0;
      const double bytes = (double)(rows_global - rows_local) * D * 2.0;
      const double tile_us = 4.0 * 128.0 * 256.0 * D / (1.45e15 / (sm_count() / 2)) * 1e6;
      const int relay_tiles = bwd_relay ? (int)(bytes / (bwd_relay * 25e9) * 1e6 / tile_us + 1.0) : 0;
      pick_split(2 * p.pair_bwd_iblk, p.n_tiles, 4, &p.bwd_chunks, &p.bwd_tiles_per_chunk, &p.bwd_small, bwd_relay / 2, relay_tiles);
    }
    p.fast_fwd = 1;
    // A row-sharded forward consumes operand rows while they arrive from the peers: it is gated by the transfer, not by
    // the tensor cores, and a second wave of tail units (which the LPT split adds) only starts when the first wave ends --
    // measured at N = 8: forward sweep 0.178 ms with equal chunks in one wave, 0.193 ms with the split.  Equal chunks there,
    // planned on the SM pairs the relay CTAs of the peer-memory transport leave free (reserved_sms).
    const bool sharded = rows_local < rows_global;
    if (!sharded) pick_split(p.pair_fwd_iblk, p.n_tiles, 2, &p.fwd1_chunks, &p.fwd1_tiles_per_chunk, &p.fwd1_small);
    else pick_chunks(p.pair_fwd_iblk, p.n_tiles, 2, &p.fwd1_chunks, &p.fwd1_tiles_per_chunk, 2, reserved_sms);
    p.n_rowgroups = p.pair_fwd_iblk * 8;
    auto force = [&](const char* name, int* chunks, int* tpc, int* small) {     // tuning knobs for measurements
      if (const char* e = getenv(name)) {
        const int c = atoi(e);
        if (c >= 1 && c <= p.n_tiles) { *tpc = (p.n_tiles + c - 1) / c; *chunks = (p.n_tiles + *tpc - 1) / *tpc; *small = 0; }
      }
    };
    force("VPA_FWD1_CHUNKS", &p.fwd1_chunks, &p.fwd1_tiles_per_chunk, &p.fwd1_small);
    force("VPA_BWD_CHUNKS", &p.bwd_chunks, &p.bwd_tiles_per_chunk, &p.bwd_small);
    p.n_dscale = 2 * p.pair_bwd_iblk * p.bwd_chunks;
    return p;
  }
  p.fwd1_chunks = 1;
  p.n_iblk = (int)((rows_local + 127) / 128);
  p.n_tiles = (int)((rows_global + 127) / 128);
  p.halves = D > 256 ? 2 : 1;
  p.cluster = p.n_iblk >= 2 ? 2 : 1;      // measured on B200 at B=32768: 2 is best (4 strands SMs, 1 doubles L2 reads)
  const int padded = (p.n_iblk + p.cluster - 1) / p.cluster * p.cluster;
  pick_chunks(2 * padded, p.n_tiles, 2, &p.fwd_chunks, &p.fwd_tiles_per_chunk);
  pick_chunks(2 * padded * p.halves, p.n_tiles, 3, &p.bwd_chunks, &p.bwd_tiles_per_chunk);
  p.n_dscale = padded * p.halves * p.bwd_chunks;
  return p;
}

Workspace carve_workspace(void* base, int64_t rows_local, int64_t rows_global, int D, const SweepPlan& plan) {
  Workspace w{};
  size_t o = 0;
  auto take = [&](size_t bytes) {
    void* p = base ? static_cast<char*>(base) + o : nullptr;
    o += align_up(bytes, 256);
    return p;
  };
  const int fchunks = plan.fwd_chunks > plan.fwd1_chunks ? plan.fwd_chunks : plan.fwd1_chunks;
  w.fwd_part = static_cast<float*>(take((size_t)2 * fchunks * rows_local * 2 * sizeof(float)));
  w.bwd_part = static_cast<float*>(take((size_t)2 * plan.bwd_chunks * rows_local * D * sizeof(float)));
  w.dscale_part = static_cast<float*>(take((size_t)(plan.n_dscale > 0 ? plan.n_dscale : 1) * sizeof(float)));
  w.colsum = static_cast<float*>(take((size_t)kColSumSplit * rows_global * sizeof(float)));
  w.colpart = plan.fast_fwd ? static_cast<float*>(take((size_t)plan.n_rowgroups * rows_global * sizeof(float))) : nullptr;
  w.bytes = o;
  return w;
}

static int check_infonce_shape(int64_t rows_local, int64_t rows_global, int D, int64_t row_offset, int precision) {
  VPA_CHECK_ARG(precision == VPA_PREC_BF16_TC || precision == VPA_PREC_FP32_SIMT, "bad precision %d", precision);
  VPA_CHECK_ARG(rows_local > 0 && rows_global >= rows_local, "need 0 < rows_local <= rows_global");
  VPA_CHECK_ARG(row_offset >= 0 && row_offset + rows_local <= rows_global, "row_offset out of range");
  if (precision == VPA_PREC_BF16_TC) {
    const int kb = D / 64;
    if (D % 64 != 0 || D < 64 || D > 512 || (kb > 4 && (kb & 1)))
      return set_error(VPA_E_UNSUPPORTED, "tensor-core path supports D in {64,128,192,256,384,512} (D=%d)", D);
  } else {
    if (D % 4 != 0 || D < 4 || D > 1024)
      return set_error(VPA_E_UNSUPPORTED, "fp32 path supports D %% 4 == 0, D <= 1024 (D=%d)", D);
  }
  return 0;
}

size_t infonce_workspace_bytes(int64_t rows_local, int64_t rows_global, int D, int precision, int reserved_sms) {
  if (rows_local <= 0 || rows_global < rows_local || D <= 0) return 0;
  const SweepPlan plan = plan_sweep(rows_local, rows_global, D, precision, reserved_sms);
  return carve_workspace(nullptr, rows_local, rows_global, D, plan).bytes;
}

}  // namespace vpa

using namespace vpa;

extern "C" {

int vpa_version(void) { return VPA_VERSION; }

const char* vpa_last_error_string(void) { return g_err; }

// Work decomposition of the sweeps for a shape (diagnostics / tests): out[0..9] = n_tiles, fwd1 chunks, fwd1 tiles per big
// chunk, fwd1 tail tiles, bwd chunks, bwd tiles per big chunk, bwd tail tiles, fwd1 row blocks, bwd row blocks, impl.
int vpa_plan_query(int64_t rows_local, int64_t rows_global, int D, int precision, int peer_memory, int* out10) {
  VPA_CHECK_ARG(out10 && rows_local > 0 && rows_global >= rows_local && D > 0, "plan_query: bad argument");
  const SweepPlan p = plan_sweep(rows_local, rows_global, D, precision, peer_memory ? relay_ctas_default() : 0);
  const int v[10] = {p.n_tiles, p.fwd1_chunks, p.fwd1_tiles_per_chunk, p.fwd1_small, p.bwd_chunks, p.bwd_tiles_per_chunk,
                     p.bwd_small, p.pair_fwd_iblk, p.pair_bwd_iblk, p.impl};
  for (int i = 0; i < 10; ++i) out10[i] = v[i];
  return 0;
}

unsigned long long vpa_launch_count(void) { return g_launches; }

int vpa_profile_hold(int hold) {
  g_prof.hold = hold != 0;
  return 0;
}

int vpa_profile_enable(int on) {
  g_prof.on = on != 0;
  g_prof.hold = false;
  for (int k = 0; k < PROF_KINDS; ++k) g_prof.n[k] = 0;
  return 0;
}

int vpa_profile_read(int kind, float* total_ms, int* launches) {
  VPA_CHECK_ARG(kind >= 0 && kind < PROF_KINDS && total_ms && launches, "profile_read: bad argument");
  float tot = 0.f;
  for (int i = 0; i < g_prof.n[kind]; ++i) {
    VPA_CUDA(cudaEventSynchronize(g_prof.ev[kind][i][1]));
    float ms = 0.f;
    VPA_CUDA(cudaEventElapsedTime(&ms, g_prof.ev[kind][i][0], g_prof.ev[kind][i][1]));
    tot += ms;
  }
  *total_ms = tot;
  *launches = g_prof.n[kind];
  g_prof.n[kind] = 0;
  return 0;
}

int vpa_normalize_cast(const void* x, int in_dtype, int64_t rows, int D, int64_t ld, int already_normalized,
                       void* y_bf16, float* y_f32, float* inv_norm, void* stream) {
  return normalize_cast_launch(x, in_dtype, rows, D, ld, already_normalized, y_bf16, y_f32, inv_norm,
                               static_cast<cudaStream_t>(stream));
}

int vpa_normalize_pair(const void* x1, const void* x2, int in_dtype, int64_t rows, int D, int64_t ld1, int64_t ld2,
                       int already_normalized, void* a_bf16, void* t_bf16, float* a_f32, float* t_f32,
                       float* inv_norm1, float* inv_norm2, float* diag_cos, int diag_from_bf16, void* stream) {
  return normalize_pair_launch(x1, x2, in_dtype, rows, D, ld1, ld2, already_normalized, a_bf16, t_bf16, a_f32, t_f32,
                               inv_norm1, inv_norm2, diag_cos, diag_from_bf16, static_cast<cudaStream_t>(stream));
}

size_t vpa_infonce_workspace_bytes(int64_t rows_local, int64_t rows_global, int D, int precision) {
  return vpa::infonce_workspace_bytes(rows_local, rows_global, D, precision, 0);
}

size_t vpa_infonce_colsum_floats(int64_t rows_global) { return rows_global > 0 ? (size_t)kColSumSplit * (size_t)rows_global : 0; }

static int fwd_sweep_impl(const void* a_loc, const void* t_loc, const void* a_all, const void* t_all, int precision,
                          int64_t rows_local, int64_t rows_global, int D, int64_t row_offset, const float* logit_scale,
                          float scale_max, void* workspace, size_t workspace_bytes, float* col_sum, bool allow_fast,
                          int parts, cudaStream_t st, const P2PRowFlags* yflags = nullptr, bool reduce_cols = true,
                          const RelayArgs* relay = nullptr, int reserved_sms = 0) {
  // reduce_cols == false: the caller reduces the single-pass column partials itself (pack_stats); col_sum is not touched
  if (int e = check_infonce_shape(rows_local, rows_global, D, row_offset, precision)) return e;
  VPA_CHECK_ARG(a_loc && t_loc && a_all && t_all && logit_scale && workspace, "infonce_fwd_sweep: null pointer");
  const SweepPlan plan = plan_sweep(rows_local, rows_global, D, precision, reserved_sms);
  const Workspace ws = carve_workspace(workspace, rows_local, rows_global, D, plan);
  if (ws.bytes > workspace_bytes) return set_error(VPA_E_WORKSPACE, "infonce_fwd: workspace %zu < %zu", workspace_bytes, ws.bytes);
  SweepArgs a{};
  a.x[0] = a_loc; a.y[0] = t_all; a.x[1] = t_loc; a.y[1] = a_all;
  a.rows_local = rows_local; a.rows_global = rows_global; a.row_offset = row_offset; a.D = D;
  a.logit_scale = logit_scale;
  a.scale_cap = (scale_max > 0.f) ? scale_max : INFINITY;     // `cfg.scale_max or float("inf")`
  if (yflags) a.yflags = *yflags;
  a.relay = relay;
  float* cs = col_sum ? col_sum : ws.colsum;
  const bool fast = precision == VPA_PREC_BF16_TC && plan.impl == 1 && plan.fast_fwd && allow_fast;
  if (parts & 1) {
    if (reduce_cols) VPA_CUDA(cudaMemsetAsync(cs, 0, (size_t)kColSumSplit * rows_global * sizeof(float), st));
    if (fast) {
      if (int e = pair_infonce_fwd(a, ws, plan, 1, st)) return e;
    }
  }
  if (parts & 2) {
    if (precision != VPA_PREC_BF16_TC) return simt_infonce_fwd(a, ws, plan, st);
    if (plan.impl != 1) return tc_infonce_fwd(a, ws, plan, st);
    if (int e = pair_infonce_fwd(a, ws, plan, fast ? 2 : 0, st)) return e;
    if (fast && reduce_cols) return colsum_reduce_launch(ws, plan, rows_global, logit_scale, a.scale_cap, cs, st);
  }
  return 0;
}

static int fwd_finish_impl(int precision, int64_t rows_local, int64_t rows_global, int D, int64_t row_offset,
                           const float* logit_scale, float scale_max, const float* diag_cos, void* workspace,
                           size_t workspace_bytes, const float* col_sum, bool allow_fast, float* row_lse, float* col_lse,
                           float* diag, float* scale_out, cudaStream_t st) {
  if (int e = check_infonce_shape(rows_local, rows_global, D, row_offset, precision)) return e;
  VPA_CHECK_ARG(logit_scale && diag_cos && row_lse && col_lse && diag && workspace, "infonce_fwd_finish: null pointer");
  const SweepPlan plan = plan_sweep(rows_local, rows_global, D, precision);
  const Workspace ws = carve_workspace(workspace, rows_local, rows_global, D, plan);
  if (ws.bytes > workspace_bytes) return set_error(VPA_E_WORKSPACE, "infonce_fwd: workspace %zu < %zu", workspace_bytes, ws.bytes);
  const bool fast = precision == VPA_PREC_BF16_TC && plan.impl == 1 && plan.fast_fwd && allow_fast;
  const float cap = (scale_max > 0.f) ? scale_max : INFINITY;
  return combine_stats_launch(ws, plan, rows_local, rows_global, row_offset, logit_scale, cap, diag_cos, fast ? 1 : 0,
                              col_sum ? col_sum : ws.colsum, row_lse, col_lse, diag, scale_out, st);
}

int vpa_infonce_fwd_sweep(const void* a_loc, const void* t_loc, const void* a_all, const void* t_all, int precision,
                          int64_t rows_local, int64_t rows_global, int D, int64_t row_offset, const float* logit_scale,
                          float scale_max, void* workspace, size_t workspace_bytes, float* col_sum, int parts, void* stream) {
  VPA_CHECK_ARG(col_sum != nullptr, "infonce_fwd_sweep: null col_sum");
  VPA_CHECK_ARG(parts >= 1 && parts <= 3, "infonce_fwd_sweep: parts must be 1, 2 or 3");
  return fwd_sweep_impl(a_loc, t_loc, a_all, t_all, precision, rows_local, rows_global, D, row_offset, logit_scale, scale_max,
                        workspace, workspace_bytes, col_sum, true, parts, static_cast<cudaStream_t>(stream));
}

int vpa_infonce_fwd_finish(int precision, int64_t rows_local, int64_t rows_global, int D, int64_t row_offset,
                           const float* logit_scale, float scale_max, const float* diag_cos, void* workspace,
                           size_t workspace_bytes, const float* col_sum, float* row_lse, float* col_lse, float* diag,
                           float* scale_out, void* stream) {
  VPA_CHECK_ARG(col_sum != nullptr, "infonce_fwd_finish: null col_sum");
  return fwd_finish_impl(precision, rows_local, rows_global, D, row_offset, logit_scale, scale_max, diag_cos, workspace,
                         workspace_bytes, col_sum, true, row_lse, col_lse, diag, scale_out, static_cast<cudaStream_t>(stream));
}

int vpa_infonce_fwd(const void* a_loc, const void* t_loc, const void* a_all, const void* t_all, int precision,
                    int64_t rows_local, int64_t rows_global, int D, int64_t row_offset, const float* logit_scale,
                    float scale_max, const float* diag_cos, void* workspace, size_t workspace_bytes,
                    float* row_lse, float* col_lse, float* diag, float* scale_out, void* stream) {
  // Without a cross-rank reduction in between, the single-pass forward is only complete for an unsharded batch.
  const bool allow_fast = rows_local == rows_global;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (int e = fwd_sweep_impl(a_loc, t_loc, a_all, t_all, precision, rows_local, rows_global, D, row_offset, logit_scale,
                             scale_max, workspace, workspace_bytes, nullptr, allow_fast, 3, st)) return e;
  return fwd_finish_impl(precision, rows_local, rows_global, D, row_offset, logit_scale, scale_max, diag_cos, workspace,
                         workspace_bytes, nullptr, allow_fast, row_lse, col_lse, diag, scale_out, st);
}

int vpa_infonce_loss(const float* row_lse, const float* col_lse, const float* diag, int64_t rows_global,
                     float* loss_out, void* stream) {
  VPA_CHECK_ARG(row_lse && col_lse && diag && loss_out && rows_global > 0, "infonce_loss: bad argument");
  return loss_launch(row_lse, col_lse, diag, rows_global, loss_out, static_cast<cudaStream_t>(stream));
}

static int bwd_impl(const void* a_loc, const void* t_loc, const void* a_all, const void* t_all, int precision,
                    int64_t rows_local, int64_t rows_global, int D, int64_t row_offset, const float* scale,
                    const float* row_lse_all, const float* col_lse_all, const float* grad_out, const void* x1,
                    const void* x2, int in_dtype, int64_t ld1, int64_t ld2, const float* inv_norm1,
                    const float* inv_norm2, int already_normalized, void* workspace, size_t workspace_bytes,
                    void* dx1, void* dx2, float* dlogit_scale, const P2PStep* p2p, int dls_sum, void* stream,
                    const RelayArgs* relay = nullptr, int reserved_sms = 0) {
  if (int e = check_infonce_shape(rows_local, rows_global, D, row_offset, precision)) return e;
  VPA_CHECK_ARG(a_loc && t_loc && a_all && t_all && scale && row_lse_all && col_lse_all && grad_out && dx1 && dx2 && workspace,
                "infonce_bwd: null pointer");
  VPA_CHECK_ARG(already_normalized || (x1 && x2 && inv_norm1 && inv_norm2), "infonce_bwd: x / inv_norm required");
  VPA_CHECK_ARG(in_dtype == VPA_F32 || in_dtype == VPA_BF16 || in_dtype == VPA_F16, "infonce_bwd: bad dtype");
  VPA_CHECK_ARG(ld1 >= D && ld2 >= D && ld1 % 4 == 0 && ld2 % 4 == 0, "infonce_bwd: bad leading dimension");
  const SweepPlan plan = plan_sweep(rows_local, rows_global, D, precision, reserved_sms);
  const Workspace ws = carve_workspace(workspace, rows_local, rows_global, D, plan);
  if (ws.bytes > workspace_bytes) return set_error(VPA_E_WORKSPACE, "infonce_bwd: workspace %zu < %zu", workspace_bytes, ws.bytes);
  SweepArgs a{};
  a.x[0] = a_loc; a.y[0] = t_all; a.x[1] = t_loc; a.y[1] = a_all;
  a.rows_local = rows_local; a.rows_global = rows_global; a.row_offset = row_offset; a.D = D;
  a.scale = scale;
  a.lse_x[0] = row_lse_all; a.lse_y[0] = col_lse_all;   // problem 0: rows of S
  a.lse_x[1] = col_lse_all; a.lse_y[1] = row_lse_all;   // problem 1: columns of S
  if (relay) {                                          // peer-memory transport: the x1 operands arrive while the sweep runs
    a.relay = relay;
    a.aflags = p2p->aflags;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (int e = (precision == VPA_PREC_BF16_TC) ? (plan.impl == 1 ? pair_infonce_bwd(a, ws, plan, st) : tc_infonce_bwd(a, ws, plan, st))
                                              : simt_infonce_bwd(a, ws, plan, st)) return e;
  return finalize_bwd_launch(ws, plan, rows_local, D, scale, grad_out, x1, x2, in_dtype, ld1, ld2, inv_norm1, inv_norm2,
                             already_normalized, dx1, dx2, dlogit_scale, p2p, dls_sum, st);
}

int vpa_infonce_bwd(const void* a_loc, const void* t_loc, const void* a_all, const void* t_all, int precision,
                    int64_t rows_local, int64_t rows_global, int D, int64_t row_offset, const float* scale,
                    const float* row_lse_all, const float* col_lse_all, const float* grad_out, const void* x1,
                    const void* x2, int in_dtype, int64_t ld1, int64_t ld2, const float* inv_norm1,
                    const float* inv_norm2, int already_normalized, void* workspace, size_t workspace_bytes,
                    void* dx1, void* dx2, float* dlogit_scale, void* stream) {
  return bwd_impl(a_loc, t_loc, a_all, t_all, precision, rows_local, rows_global, D, row_offset, scale, row_lse_all,
                  col_lse_all, grad_out, x1, x2, in_dtype, ld1, ld2, inv_norm1, inv_norm2, already_normalized, workspace,
                  workspace_bytes, dx1, dx2, dlogit_scale, nullptr, 0, stream);
}

size_t vpa_sim_workspace_bytes(int64_t N, int64_t M) {
  if (N <= 0 || M <= 0) return 0;
  return align_up((size_t)N * (size_t)M * sizeof(float), 256);
}

int vpa_sim_rank_topk(const float* Q, const float* K, int64_t N, int64_t M, int D, int64_t ldq, int64_t ldk,
                      const int32_t* gt_idx, int g, int k, int64_t* topk_idx, float* topk_val, int32_t* ranks,
                      void* workspace, size_t workspace_bytes, void* stream) {
  VPA_CHECK_ARG(workspace != nullptr, "sim_rank_topk: null workspace");
  if (vpa_sim_workspace_bytes(N, M) > workspace_bytes)
    return set_error(VPA_E_WORKSPACE, "sim_rank_topk: workspace %zu < %zu", workspace_bytes, vpa_sim_workspace_bytes(N, M));
  return sim_rank_topk_launch(Q, K, N, M, D, ldq, ldk, gt_idx, g, k, topk_idx, topk_val, ranks,
                              static_cast<float*>(workspace), static_cast<cudaStream_t>(stream));
}

size_t vpa_sim_fused_workspace_bytes(int64_t N, int64_t M, int g_q, int g_k) { return sim_fused_workspace_bytes(N, M, g_q, g_k); }

int vpa_sim_rank_fused(const float* Q, const float* K, int64_t N, int64_t M, int D, int64_t ldq, int64_t ldk,
                       const int32_t* gt_q, int g_q, const int32_t* gt_k, int g_k, int32_t* ranks_q, int32_t* ranks_k,
                       int64_t* top1_q, float* top1_val_q, int64_t* top1_k, float* top1_val_k, void* workspace,
                       size_t workspace_bytes, void* stream) {
  return sim_rank_fused_launch(Q, K, N, M, D, ldq, ldk, gt_q, g_q, gt_k, g_k, ranks_q, ranks_k, top1_q, top1_val_q, top1_k,
                               top1_val_k, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

int vpa_encoder_tail(const void* x, int in_dtype, int64_t rows, int width, int64_t ld, const float* ln_gamma,
                     const float* ln_beta, float eps, const void* proj_t_bf16, int N, void* ln_scratch_bf16, float* mean,
                     float* rstd, void* a_bf16, float* y_f32, float* inv_norm, void* stream) {
  return encoder_tail_launch(x, in_dtype, rows, width, ld, ln_gamma, ln_beta, eps, proj_t_bf16, N, ln_scratch_bf16, mean, rstd,
                             a_bf16, y_f32, inv_norm, static_cast<cudaStream_t>(stream));
}

size_t vpa_multilabel_workspace_bytes(int64_t N, int C) { return multilabel_workspace_bytes(N, C); }

int vpa_multilabel_scores(const float* S, int64_t ld_s, const void* Y, int y_dtype, int64_t ld_y, int64_t N, int C,
                          int truncate_pr, double* per_class, int32_t* flags, int32_t* support, double* micro_ap,
                          void* workspace, size_t workspace_bytes, void* stream) {
  return multilabel_scores_launch(S, ld_s, Y, y_dtype, ld_y, N, C, truncate_pr, per_class, flags, support, micro_ap, workspace,
                                  workspace_bytes, static_cast<cudaStream_t>(stream));
}

// ---- several InfoNCE pairs over shared modalities in one set of launches (composite heads) ---------------------------
static bool pack_reduces_columns(const SweepPlan& plan, int precision);
struct MultiState {
  void* op[kMaxPairs];            // (rows, D) bf16 operands per modality
  float* inv[kMaxPairs];          // (rows,)
  float* dcos[kMaxPairs];         // per pair
  float* msg[kMaxPairs];          // per pair (4 rows)
  float* stats_all[kMaxPairs];    // per pair (3 rows)
  float* scale[kMaxPairs];        // per pair {s, flows}
  double* loss_part[kMaxPairs];
  uint32_t* loss_counter[kMaxPairs];
  void* ws[kMaxPairs];
  size_t ws_bytes, bytes;
};
static MultiState carve_multi(void* base, int64_t rows, int D, int n_mod, int n_pairs, int precision) {
  MultiState h{};
  size_t o = 0;
  auto take = [&](size_t bytes) {
    void* p = base ? static_cast<char*>(base) + o : nullptr;
    o += align_up(bytes, 256);
    return p;
  };
  for (int p = 0; p < n_pairs; ++p) h.loss_counter[p] = (uint32_t*)take(4);      // (first: one memset clears them all)
  for (int m = 0; m < n_mod; ++m) { h.op[m] = take((size_t)rows * D * 2); h.inv[m] = (float*)take(rows * 4); }
  h.ws_bytes = infonce_workspace_bytes(rows, rows, D, precision, 0);
  for (int p = 0; p < n_pairs; ++p) {
    h.dcos[p] = (float*)take(rows * 4);
    h.msg[p] = (float*)take((size_t)4 * rows * 4);
    h.stats_all[p] = (float*)take((size_t)3 * rows * 4);
    h.scale[p] = (float*)take(16);
    h.loss_part[p] = (double*)take((size_t)((rows + 255) / 256) * 8);
    h.ws[p] = take(h.ws_bytes);
  }
  h.bytes = o;
  return h;
}
static int check_multi(int64_t rows, int D, int n_mod, const int32_t* pair_x, const int32_t* pair_y, int n_pairs, int precision) {
  VPA_CHECK_ARG(n_mod >= 1 && n_mod <= kMaxPairs && n_pairs >= 1 && n_pairs <= kMaxPairs && pair_x && pair_y,
                "infonce_multi: 1..%d modalities and pairs", kMaxPairs);
  for (int p = 0; p < n_pairs; ++p)
    VPA_CHECK_ARG(pair_x[p] >= 0 && pair_x[p] < n_mod && pair_y[p] >= 0 && pair_y[p] < n_mod, "infonce_multi: pair %d names a modality outside [0, %d)", p, n_mod);
  if (int e = check_infonce_shape(rows, rows, D, 0, precision)) return e;
  if (precision != VPA_PREC_BF16_TC || !(D == 256 || D == 512))
    return set_error(VPA_E_UNSUPPORTED, "infonce_multi: the fused multi-pair step covers the tensor-core path with D in {256, 512}");
  const SweepPlan plan = plan_sweep(rows, rows, D, precision);
  if (!pack_reduces_columns(plan, precision))
    return set_error(VPA_E_UNSUPPORTED, "infonce_multi: at most 8192 rows (larger batches gain nothing from sharing launches)");
  return 0;
}

size_t vpa_infonce_multi_state_bytes(int64_t rows, int D, int n_mod, int n_pairs, int precision) {
  if (rows <= 0 || D <= 0 || n_mod < 1 || n_mod > kMaxPairs || n_pairs < 1 || n_pairs > kMaxPairs) return 0;
  return carve_multi(nullptr, rows, D, n_mod, n_pairs, precision).bytes;
}

int vpa_infonce_multi_fwd(const void* const* x, const int64_t* ld, int in_dtype, int64_t rows, int D, int n_mod,
                          int already_normalized, const int32_t* pair_x, const int32_t* pair_y, int n_pairs,
                          const float* const* logit_scale, const float* scale_max, int precision, void* state,
                          size_t state_bytes, float* loss_out, void* stream) {
  if (int e = check_multi(rows, D, n_mod, pair_x, pair_y, n_pairs, precision)) return e;
  VPA_CHECK_ARG(x && ld && logit_scale && scale_max && state && loss_out, "infonce_multi_fwd: null pointer");
  const MultiState h = carve_multi(state, rows, D, n_mod, n_pairs, precision);
  if (h.bytes > state_bytes) return set_error(VPA_E_WORKSPACE, "infonce_multi_fwd: state %zu < %zu", state_bytes, h.bytes);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const SweepPlan plan = plan_sweep(rows, rows, D, precision);
  // every modality is normalised once, whatever number of pairs it takes part in
  if (int e = normalize_multi_launch(x, ld, in_dtype, rows, D, n_mod, already_normalized, h.op, h.inv, st)) return e;
  const void* a_ptr[kMaxPairs];
  const void* t_ptr[kMaxPairs];
  float cap[kMaxPairs];
  Workspace ws[kMaxPairs];
  for (int p = 0; p < n_pairs; ++p) {
    VPA_CHECK_ARG(logit_scale[p] != nullptr, "infonce_multi_fwd: null logit_scale of pair %d", p);
    a_ptr[p] = h.op[pair_x[p]];
    t_ptr[p] = h.op[pair_y[p]];
    cap[p] = (scale_max[p] > 0.f) ? scale_max[p] : INFINITY;
    ws[p] = carve_workspace(h.ws[p], rows, rows, D, plan);
  }
  if (int e = diag_cos_multi_launch(a_ptr, t_ptr, h.dcos, n_pairs, rows, D, st)) return e;
  VPA_CUDA(cudaMemsetAsync(h.loss_counter[0], 0, (size_t)n_pairs * 256, st));
  // single-pass forward of all pairs in one launch (each pair gates itself on the device value of ITS temperature) ...
  PairLaunch L{};
  L.rows_local = L.rows_global = rows; L.row_offset = 0; L.D = D;
  L.n_prob = n_pairs;
  for (int p = 0; p < n_pairs; ++p) {
    L.p[p].x = a_ptr[p]; L.p[p].y = t_ptr[p];
    L.p[p].out = ws[p].fwd_part; L.p[p].colpart = ws[p].colpart;
    L.p[p].logit_scale = logit_scale[p]; L.p[p].scale_cap = cap[p];
  }
  if (int e = pair_launch_fwd1(L, plan, st)) return e;
  // ... and the exact two-sweep kernel for the pairs in the other regime
  PairLaunch X{};
  X.rows_local = X.rows_global = rows; X.row_offset = 0; X.D = D;
  X.n_prob = 2 * n_pairs;
  for (int p = 0; p < n_pairs; ++p)
    for (int d = 0; d < 2; ++d) {
      PairProblem& q = X.p[2 * p + d];
      q.x = d ? t_ptr[p] : a_ptr[p]; q.y = d ? a_ptr[p] : t_ptr[p];
      q.out = ws[p].fwd_part + (int64_t)d * plan.fwd_chunks * rows * 2;
      q.logit_scale = logit_scale[p]; q.scale_cap = cap[p];
    }
  if (int e = pair_launch_fwd(X, plan, 2, st)) return e;
  return pack_merge_multi_launch(n_pairs, ws, plan, rows, logit_scale, cap, h.dcos, h.msg, h.stats_all, h.scale, h.loss_part,
                                 h.loss_counter, loss_out, st);
}

int vpa_infonce_multi_bwd(const void* const* x, const int64_t* ld, int in_dtype, int64_t rows, int D, int n_mod,
                          int already_normalized, const int32_t* pair_x, const int32_t* pair_y, int n_pairs, int precision,
                          const float* grad_out, void* state, size_t state_bytes, void* const* dx, float* dlogit_scale,
                          void* stream) {
  if (int e = check_multi(rows, D, n_mod, pair_x, pair_y, n_pairs, precision)) return e;
  VPA_CHECK_ARG(ld && grad_out && state && dx && dlogit_scale && (already_normalized || x), "infonce_multi_bwd: null pointer");
  VPA_CHECK_ARG(in_dtype == VPA_F32 || in_dtype == VPA_BF16 || in_dtype == VPA_F16, "infonce_multi_bwd: bad dtype");
  const MultiState h = carve_multi(state, rows, D, n_mod, n_pairs, precision);
  if (h.bytes > state_bytes) return set_error(VPA_E_WORKSPACE, "infonce_multi_bwd: state %zu < %zu", state_bytes, h.bytes);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const SweepPlan plan = plan_sweep(rows, rows, D, precision);
  Workspace ws[kMaxPairs];
  PairLaunch L{};
  L.rows_local = L.rows_global = rows; L.row_offset = 0; L.D = D;
  L.n_prob = 2 * n_pairs;
  FinMultiHost f{};
  f.n_mod = n_mod; f.n_pairs = n_pairs; f.n_chunks = plan.bwd_chunks; f.D = D; f.already = already_normalized;
  f.n_dscale = plan.n_dscale; f.in_dtype = in_dtype; f.rows = rows;
  for (int m = 0; m < n_mod; ++m) {
    VPA_CHECK_ARG(dx[m] != nullptr && ld[m] >= D && ld[m] % 4 == 0, "infonce_multi_bwd: bad gradient buffer / leading dimension of modality %d", m);
    f.x[m] = x ? x[m] : nullptr; f.dx[m] = dx[m]; f.ld[m] = ld[m]; f.inv[m] = h.inv[m];
  }
  for (int p = 0; p < n_pairs; ++p) {
    ws[p] = carve_workspace(h.ws[p], rows, rows, D, plan);
    const float* row_lse = h.stats_all[p];
    const float* col_lse = h.stats_all[p] + rows;
    for (int d = 0; d < 2; ++d) {
      PairProblem& q = L.p[2 * p + d];
      const int mx = d ? pair_y[p] : pair_x[p], my = d ? pair_x[p] : pair_y[p];
      q.x = h.op[mx]; q.y = h.op[my];
      q.lse_x = d ? col_lse : row_lse; q.lse_y = d ? row_lse : col_lse;
      q.out = ws[p].bwd_part + (int64_t)d * plan.bwd_chunks * rows * D;
      q.dscale = d == 0 ? ws[p].dscale_part : nullptr;
      q.scale = h.scale[p];
      f.part[mx][f.n_src[mx]] = q.out;
      f.src_pair[mx][f.n_src[mx]] = p;
      ++f.n_src[mx];
    }
    f.scale[p] = h.scale[p];
    f.dscale_part[p] = ws[p].dscale_part;
  }
  f.grad_out = grad_out;
  f.dlogit_scale = dlogit_scale;
  if (int e = pair_launch_bwd(L, plan, st)) return e;
  return finalize_multi_launch(f, st);
}

// ---- row-sharded step orchestrated in the library (two calls per training step) ------------------------------------
struct ShardState {
  void *a_all, *t_all;          // (B, D) operands: bf16 (tensor-core mode) or fp32; this rank's rows are written in place
  float *inv1, *inv2, *dcos;    // (b,)
  float *colsum8, *msg, *msgs;  // (8, B); (B + 3b); (R, B + 3b)
  float *stats_all, *scale;     // (3, B); (2,)
  double* loss_part;            // per-block partial sums of the loss (merge_stats)
  uint32_t* loss_counter;
  void* ws;
  size_t ws_bytes, bytes;
};
// the single-pass column partials are few enough for pack_stats to reduce them itself (saves a launch and a pass)
static bool pack_reduces_columns(const SweepPlan& plan, int precision) {
  return precision == VPA_PREC_BF16_TC && plan.impl == 1 && plan.fast_fwd && plan.n_rowgroups <= 256;
}
static ShardState carve_state(void* base, int64_t b, int world, int D, int precision) {
  ShardState h{};
  size_t o = 0;
  auto take = [&](size_t bytes) {
    void* p = base ? static_cast<char*>(base) + o : nullptr;
    o += align_up(bytes, 256);
    return p;
  };
  const int64_t B = b * world;
  const size_t es = precision == VPA_PREC_BF16_TC ? 2 : 4;
  h.a_all = take((size_t)B * D * es);
  h.t_all = take((size_t)B * D * es);
  h.inv1 = (float*)take(b * 4); h.inv2 = (float*)take(b * 4); h.dcos = (float*)take(b * 4);
  h.colsum8 = (float*)take((size_t)kColSumSplit * B * 4);
  h.msg = (float*)take((size_t)(B + 3 * b) * 4);
  h.msgs = world > 1 ? (float*)take((size_t)world * (B + 3 * b) * 4) : h.msg;
  h.stats_all = (float*)take((size_t)3 * B * 4);
  h.scale = (float*)take(16);
  h.loss_part = (double*)take((size_t)((B + 255) / 256) * 8);
  h.loss_counter = (uint32_t*)take(4);
  h.ws_bytes = vpa_infonce_workspace_bytes(b, B, D, precision);
  h.ws = take(h.ws_bytes);
  h.bytes = o;
  return h;
}

// side stream + events: the all-gather of the x1 operands runs beside the single-pass forward, which does not read them
struct SideStream {
  cudaStream_t s = nullptr;
  cudaEvent_t fork = nullptr, t_done = nullptr, a_done = nullptr;
  int dev = -1;
};
static int side_stream(SideStream** out) {
  static thread_local SideStream ss;
  int dev = 0;
  VPA_CUDA(cudaGetDevice(&dev));
  if (!ss.s || ss.dev != dev) {
    VPA_CUDA(cudaStreamCreateWithFlags(&ss.s, cudaStreamNonBlocking));
    VPA_CUDA(cudaEventCreateWithFlags(&ss.fork, cudaEventDisableTiming));
    VPA_CUDA(cudaEventCreateWithFlags(&ss.t_done, cudaEventDisableTiming));
    VPA_CUDA(cudaEventCreateWithFlags(&ss.a_done, cudaEventDisableTiming));
    ss.dev = dev;
  }
  *out = &ss;
  return 0;
}

int vpa_comm_load(const char* libnccl_path) { return comm_load(libnccl_path); }
int vpa_comm_unique_id(void* out128) { VPA_CHECK_ARG(out128, "comm_unique_id: null"); return comm_unique_id(out128); }
int vpa_comm_init(const void* id128, int rank, int world, void** comm_out) {
  VPA_CHECK_ARG(id128 && comm_out && world >= 1 && rank >= 0 && rank < world, "comm_init: bad argument");
  return comm_init(id128, rank, world, comm_out);
}
int vpa_comm_destroy(void* comm) { return comm_destroy(comm); }

size_t vpa_sharded_state_bytes(int64_t rows_local, int world, int D, int precision) {
  if (rows_local <= 0 || world < 1 || D <= 0) return 0;
  return carve_state(nullptr, rows_local, world, D, precision).bytes;
}

int vpa_infonce_fwd_sharded(void* comm, const void* x1, const void* x2, int in_dtype, int64_t b, int world, int rank,
                            int D, int64_t ld1, int64_t ld2, int already_normalized, const float* logit_scale,
                            float scale_max, int precision, void* state, size_t state_bytes, float* loss_out,
                            void* stream) {
  VPA_CHECK_ARG(world >= 1 && rank >= 0 && rank < world && (world == 1 || comm), "fwd_sharded: bad world / rank / comm");
  const int64_t B = b * world, off = (int64_t)rank * b;
  if (int e = check_infonce_shape(b, B, D, off, precision)) return e;
  VPA_CHECK_ARG(x1 && x2 && logit_scale && state && loss_out, "fwd_sharded: null pointer");
  const ShardState h = carve_state(state, b, world, D, precision);
  if (h.bytes > state_bytes) return set_error(VPA_E_WORKSPACE, "fwd_sharded: state %zu < %zu", state_bytes, h.bytes);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool tcp = precision == VPA_PREC_BF16_TC;
  const size_t es = tcp ? 2 : 4;
  char* a_loc = static_cast<char*>(h.a_all) + (size_t)off * D * es;
  char* t_loc = static_cast<char*>(h.t_all) + (size_t)off * D * es;
  if (int e = normalize_pair_launch(x1, x2, in_dtype, b, D, ld1, ld2, already_normalized, tcp ? a_loc : nullptr,
                                    tcp ? t_loc : nullptr, tcp ? nullptr : (float*)a_loc, tcp ? nullptr : (float*)t_loc,
                                    h.inv1, h.inv2, h.dcos, tcp ? 1 : 0, st)) return e;
  SideStream* ss = nullptr;
  if (world > 1) {
    if (int e = side_stream(&ss)) return e;
    const int dt = tcp ? 9 : 7;     // ncclBfloat16 / ncclFloat32
    VPA_CUDA(cudaEventRecord(ss->fork, st));
    VPA_CUDA(cudaStreamWaitEvent(ss->s, ss->fork, 0));
    if (int e = comm_all_gather(comm, t_loc, h.t_all, (size_t)b * D, dt, ss->s)) return e;     // x2 operands first
    VPA_CUDA(cudaEventRecord(ss->t_done, ss->s));
    if (int e = comm_all_gather(comm, a_loc, h.a_all, (size_t)b * D, dt, ss->s)) return e;
    VPA_CUDA(cudaEventRecord(ss->a_done, ss->s));
    VPA_CUDA(cudaStreamWaitEvent(st, ss->t_done, 0));
  }
  const SweepPlan plan = plan_sweep(b, B, D, precision);
  const Workspace ws = carve_workspace(h.ws, b, B, D, plan);
  const bool from_colpart = pack_reduces_columns(plan, precision);
  VPA_CUDA(cudaMemsetAsync(h.loss_counter, 0, 4, st));
  // single-pass kernel (reads a_loc and t_all) ...
  if (int e = fwd_sweep_impl(a_loc, t_loc, h.a_all, h.t_all, precision, b, B, D, off, logit_scale, scale_max, h.ws, h.ws_bytes,
                             h.colsum8, true, 1, st, nullptr, !from_colpart)) return e;
  if (world > 1) VPA_CUDA(cudaStreamWaitEvent(st, ss->a_done, 0));
  // ... then everything that also reads a_all
  if (int e = fwd_sweep_impl(a_loc, t_loc, h.a_all, h.t_all, precision, b, B, D, off, logit_scale, scale_max, h.ws, h.ws_bytes,
                             h.colsum8, true, 2, st, nullptr, !from_colpart)) return e;
  const int fast = (tcp && plan.impl == 1 && plan.fast_fwd) ? 1 : 0;
  const float cap = (scale_max > 0.f) ? scale_max : INFINITY;
  if (int e = pack_stats_launch(ws, plan, b, B, logit_scale, cap, h.dcos, fast, h.colsum8, from_colpart, h.msg, st)) return e;
  if (world > 1) {
    if (int e = comm_all_gather(comm, h.msg, h.msgs, (size_t)(B + 3 * b), 7, st)) return e;
  }
  // statistics of all rows + the global loss in one kernel
  return merge_stats_launch(h.msgs, world, b, B, logit_scale, cap, fast, h.stats_all, h.scale, h.loss_part, h.loss_counter,
                            loss_out, st);
}

int vpa_infonce_bwd_sharded(void* comm, const void* x1, const void* x2, int in_dtype, int64_t b, int world, int rank,
                            int D, int64_t ld1, int64_t ld2, int already_normalized, int precision,
                            const float* grad_out, void* state, size_t state_bytes, void* dx1, void* dx2,
                            float* dlogit_scale, int dls_reduce, void* stream) {
  VPA_CHECK_ARG(world >= 1 && rank >= 0 && rank < world && (world == 1 || comm), "bwd_sharded: bad world / rank / comm");
  const int64_t B = b * world, off = (int64_t)rank * b;
  VPA_CHECK_ARG(state && grad_out && dx1 && dx2 && dlogit_scale, "bwd_sharded: null pointer");
  const ShardState h = carve_state(state, b, world, D, precision);
  if (h.bytes > state_bytes) return set_error(VPA_E_WORKSPACE, "bwd_sharded: state %zu < %zu", state_bytes, h.bytes);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t es = precision == VPA_PREC_BF16_TC ? 2 : 4;
  const char* a_loc = static_cast<const char*>(h.a_all) + (size_t)off * D * es;
  const char* t_loc = static_cast<const char*>(h.t_all) + (size_t)off * D * es;
  if (int e = vpa_infonce_bwd(a_loc, t_loc, h.a_all, h.t_all, precision, b, B, D, off, h.scale, h.stats_all, h.stats_all + B,
                              grad_out, x1, x2, in_dtype, ld1, ld2, h.inv1, h.inv2, already_normalized, h.ws, h.ws_bytes,
                              dx1, dx2, dlogit_scale, st)) return e;
  if (world > 1 && dls_reduce) return comm_all_reduce_sum_f32(comm, dlogit_scale, dlogit_scale, 1, st);
  return 0;
}

// ---- the same step over peer memory (p2p.cu): no NCCL call on the data path -----------------------------------------
int vpa_p2p_create(int64_t rows_local, int world, int rank, int D, int precision, void** p2p_out, void* ipc_handle_out64) {
  if (int e = check_infonce_shape(rows_local, rows_local * world, D, (int64_t)rank * rows_local, precision)) return e;
  return p2p_create(rows_local, world, rank, D, precision, p2p_out, ipc_handle_out64);
}
int vpa_p2p_connect(void* p2p, const void* all_ipc_handles) { return p2p_connect(p2p, all_ipc_handles); }
int vpa_p2p_destroy(void* p2p) { return p2p_destroy(p2p); }
// the relay CTAs' item -> (matrix, source rank, chunk, first row, rows) map, evaluated on the host (tests)
int vpa_debug_relay_item(int item, int m0, int source_major, int world, int me, int chunks_per_rank, int64_t rows_local, int* out5) {
  VPA_CHECK_ARG(out5 && world >= 2 && me >= 0 && me < world && chunks_per_rank >= 1 && rows_local >= 1 && item >= 0 &&
                (m0 == 0 || m0 == 1) && item < (2 - m0) * chunks_per_rank * (world - 1), "debug_relay_item: bad argument");
  const RelayItem it = relay_item_decode(item, m0, source_major, world, me, chunks_per_rank, rows_local);
  out5[0] = it.m; out5[1] = it.src; out5[2] = it.c; out5[3] = it.row0; out5[4] = it.rows;
  return 0;
}

int vpa_infonce_fwd_p2p(void* p2p, const void* x1, const void* x2, int in_dtype, int64_t b, int world, int rank, int D,
                        int64_t ld1, int64_t ld2, int already_normalized, const float* logit_scale, float scale_max,
                        int precision, float* loss_out, uint32_t* epoch_out, void* stream) {
  const int64_t B = b * world, off = (int64_t)rank * b;
  if (int e = check_infonce_shape(b, B, D, off, precision)) return e;
  VPA_CHECK_ARG(x1 && x2 && logit_scale && loss_out && epoch_out, "fwd_p2p: null pointer");
  if (int e = p2p_check(p2p, b, world, rank, D, precision)) return e;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const uint32_t epoch = p2p_next_epoch(p2p);
  *epoch_out = epoch;
  const P2PStep h = p2p_step(p2p, epoch);
  const bool tcp = precision == VPA_PREC_BF16_TC;
  const size_t es = tcp ? 2 : 4;
  char* a_loc = static_cast<char*>(h.a_all) + (size_t)off * D * es;
  char* t_loc = static_cast<char*>(h.t_all) + (size_t)off * D * es;
  if (int e = normalize_pair_launch(x1, x2, in_dtype, b, D, ld1, ld2, already_normalized, tcp ? a_loc : nullptr,
                                    tcp ? t_loc : nullptr, tcp ? nullptr : (float*)a_loc, tcp ? nullptr : (float*)t_loc,
                                    h.inv1, h.inv2, h.dcos, tcp ? 1 : 0, st)) return e;
  const int reserved = p2p_relay_ctas(p2p);
  const SweepPlan plan = plan_sweep(b, B, D, precision, reserved);
  const Workspace ws = carve_workspace(h.ws, b, B, D, plan);
  const bool from_colpart = pack_reduces_columns(plan, precision);
  const bool single_pass = tcp && plan.impl == 1 && plan.fast_fwd;
  // The x1 operands of the previous step travel in its backward's grid.  If that backward has not been called (two forwards
  // in a row), fetch them now: once this step's message is out, the peers may move on and reuse those buffers.
  if (const uint32_t pend = p2p_a_pending(p2p)) {
    if (int e = p2p_relay_standalone(p2p, pend, 1, false, st)) return e;
    p2p_set_a_pending(p2p, 0);
  }
  if (single_pass) {
    // ONE kernel is the all-gather and the contraction: its first CTAs relay the peers' x2 operand rows, its sweep CTAs
    // start on the local block and consume remote tiles as their flags flip.  (Exact temperature regime, decided on the
    // device: the relays fetch the x1 operands too, which the two-sweep kernel below reads.)
    RelayArgs fr = h.relay;
    fr.m0 = 0; fr.m1 = 1; fr.source_major = 0; fr.signal_ready = 1;
    if (int e = fwd_sweep_impl(a_loc, t_loc, h.a_all, h.t_all, precision, b, B, D, off, logit_scale, scale_max, h.ws, h.ws_bytes,
                               h.colsum8, true, 1, st, &h.yflags, !from_colpart, &fr, reserved)) return e;
    p2p_set_a_pending(p2p, epoch);
  } else {
    if (int e = p2p_relay_standalone(p2p, epoch, 0, true, st)) return e;
  }
  // the exact two-sweep kernel (other temperature regime / other shapes); device-gated when the single-pass kernel exists
  if (int e = fwd_sweep_impl(a_loc, t_loc, h.a_all, h.t_all, precision, b, B, D, off, logit_scale, scale_max, h.ws, h.ws_bytes,
                             h.colsum8, true, single_pass ? 2 : 3, st, nullptr, !from_colpart, nullptr, reserved)) return e;
  const float cap = (scale_max > 0.f) ? scale_max : INFINITY;
  // message out, every rank's message in, statistics of all rows + the global loss
  return exchange_stats_launch(ws, plan, b, B, logit_scale, cap, h.dcos, single_pass ? 1 : 0, h.colsum8, from_colpart, h, loss_out, st);
}

int vpa_infonce_bwd_p2p(void* p2p, uint32_t epoch, const void* x1, const void* x2, int in_dtype, int64_t b, int world,
                        int rank, int D, int64_t ld1, int64_t ld2, int already_normalized, int precision,
                        const float* grad_out, void* dx1, void* dx2, float* dlogit_scale, int dls_reduce, void* stream) {
  const int64_t B = b * world, off = (int64_t)rank * b;
  VPA_CHECK_ARG(grad_out && dx1 && dx2 && dlogit_scale, "bwd_p2p: null pointer");
  if (int e = p2p_check(p2p, b, world, rank, D, precision)) return e;
  const uint32_t cur = p2p_current_epoch(p2p);
  VPA_CHECK_ARG(epoch != 0 && cur - epoch <= 1u,
                "bwd_p2p: the forward of step %u is no longer resident (current step %u; the segment keeps two steps)", epoch, cur);
  const P2PStep h = p2p_step(p2p, epoch);
  const size_t es = precision == VPA_PREC_BF16_TC ? 2 : 4;
  const char* a_loc = static_cast<const char*>(h.a_all) + (size_t)off * D * es;
  const char* t_loc = static_cast<const char*>(h.t_all) + (size_t)off * D * es;
  // The peers' x1 operands (read by the second problem only) arrive through relay CTAs in front of the backward's own grid
  // while its first problem runs; finalize_bwd exchanges the d logit_scale partials.
  RelayArgs br = h.relay;
  br.m0 = 1; br.m1 = 2; br.source_major = 1; br.signal_ready = 0;
  br.n_ctas = relay_ctas_bwd_default();
  const bool fetch = p2p_a_pending(p2p) == epoch;
  if (fetch) p2p_set_a_pending(p2p, 0);
  return bwd_impl(a_loc, t_loc, h.a_all, h.t_all, precision, b, B, D, off, h.scale, h.stats_all, h.stats_all + B, grad_out, x1, x2,
                  in_dtype, ld1, ld2, h.inv1, h.inv2, already_normalized, h.ws, h.ws_bytes, dx1, dx2, dlogit_scale, &h,
                  dls_reduce ? 1 : 0, stream, fetch ? &br : nullptr, p2p_relay_ctas(p2p));
}

// ---- host-buffer end-to-end step ---------------------------------------------------------------------
// Pipelined over V row shards of the batch (the same row-sharded sweeps the multi-GPU path runs, all on one device):
//   copy-in stream:  x2 (whole) | x1 shard 0 | x1 shard 1 | ...
//   compute stream:  normalise x2 | per shard, as it lands: normalise x1 shard, <a_i,t_i>, single-pass forward sweep of the
//                    shard's rows against ALL x2 rows | statistics merge + loss | per shard: backward sweeps + finalize
//   copy-out stream: dx1 / dx2 of shard k while shard k+1 computes
// so the PCIe transfers hide behind the sweeps except the first matrix in and the last shard out.
struct HostScratch {
  float *x1, *x2, *af, *tf, *inv1, *inv2, *dcos, *msgs, *stats_all, *sc, *loss, *dls, *gout, *lsc, *dx1, *dx2, *colsum8;
  double* loss_part;
  uint32_t* loss_counter;
  void *ab, *tb;
  void* ws[16];
  size_t ws_bytes, bytes;
  int shards;
};
static int host_shards(int64_t rows, int D, int precision) {
  if (precision != VPA_PREC_BF16_TC || !(D == 256 || D == 512)) return 1;      // the pipeline is built on the CTA-pair sweeps
  int v = rows >= 32768 ? 8 : (rows >= 8192 ? 4 : 1);      // measured at 32768 x 512: 8.64 ms unpipelined, 6.79 (4 shards), 6.42 (8)
  if (const char* e = getenv("VPA_HOST_SHARDS")) { const int q = atoi(e); if (q >= 1 && q <= 16) v = q; }
  while (v > 1 && (rows % ((int64_t)v * 256) != 0)) --v;      // equal shards on 256-row tile boundaries
  return v;
}
static HostScratch carve_host(void* base, int64_t rows, int D, int precision) {
  HostScratch h{};
  size_t o = 0;
  auto take = [&](size_t bytes) {
    void* p = base ? static_cast<char*>(base) + o : nullptr;
    o += align_up(bytes, 256);
    return p;
  };
  const size_t mat = (size_t)rows * D;
  h.shards = host_shards(rows, D, precision);
  const int64_t b = rows / h.shards;
  h.x1 = (float*)take(mat * 4); h.x2 = (float*)take(mat * 4);
  h.dx1 = (float*)take(mat * 4); h.dx2 = (float*)take(mat * 4);
  h.ab = take(mat * 2); h.tb = take(mat * 2);
  if (precision == VPA_PREC_FP32_SIMT) { h.af = (float*)take(mat * 4); h.tf = (float*)take(mat * 4); }
  h.inv1 = (float*)take(rows * 4); h.inv2 = (float*)take(rows * 4); h.dcos = (float*)take(rows * 4);
  h.msgs = (float*)take((size_t)h.shards * (rows + 3 * b) * 4);
  h.stats_all = (float*)take((size_t)3 * rows * 4);
  h.colsum8 = (float*)take((size_t)kColSumSplit * rows * 4);
  h.loss_part = (double*)take((size_t)((rows + 255) / 256) * 8);
  h.loss_counter = (uint32_t*)take(4);
  h.sc = (float*)take(16); h.loss = (float*)take(16); h.dls = (float*)take(16 * 4);
  h.gout = (float*)take(16); h.lsc = (float*)take(16);
  h.ws_bytes = vpa_infonce_workspace_bytes(b, rows, D, precision);
  for (int k = 0; k < h.shards; ++k) h.ws[k] = take(h.ws_bytes);
  h.bytes = o;
  return h;
}

size_t vpa_infonce_host_scratch_bytes(int64_t rows, int D, int precision) {
  if (rows <= 0 || D <= 0) return 0;
  return carve_host(nullptr, rows, D, precision).bytes;
}

struct HostPipe {       // copy streams + events of the pipelined host step (one set per thread and device)
  cudaStream_t cin = nullptr, cout = nullptr;
  cudaEvent_t start = nullptr, x2_in = nullptr, x1_in[16] = {}, bwd_done[16] = {}, out_done = nullptr;
  int dev = -1;
};
static int host_pipe(HostPipe** out) {
  static thread_local HostPipe hp;
  int dev = 0;
  VPA_CUDA(cudaGetDevice(&dev));
  if (!hp.cin || hp.dev != dev) {
    VPA_CUDA(cudaStreamCreateWithFlags(&hp.cin, cudaStreamNonBlocking));
    VPA_CUDA(cudaStreamCreateWithFlags(&hp.cout, cudaStreamNonBlocking));
    VPA_CUDA(cudaEventCreateWithFlags(&hp.start, cudaEventDisableTiming));
    VPA_CUDA(cudaEventCreateWithFlags(&hp.x2_in, cudaEventDisableTiming));
    VPA_CUDA(cudaEventCreateWithFlags(&hp.out_done, cudaEventDisableTiming));
    for (int k = 0; k < 16; ++k) {
      VPA_CUDA(cudaEventCreateWithFlags(&hp.x1_in[k], cudaEventDisableTiming));
      VPA_CUDA(cudaEventCreateWithFlags(&hp.bwd_done[k], cudaEventDisableTiming));
    }
    hp.dev = dev;
  }
  *out = &hp;
  return 0;
}

int vpa_infonce_step_host(const float* x1_host, const float* x2_host, int64_t rows, int D, float logit_scale,
                          float scale_max, float grad_out, int precision, void* dev_scratch,
                          size_t dev_scratch_bytes, float* loss_host, float* dlogit_scale_host, float* dx1_host,
                          float* dx2_host, void* stream) {
  if (int e = check_infonce_shape(rows, rows, D, 0, precision)) return e;
  VPA_CHECK_ARG(x1_host && x2_host && dev_scratch && loss_host, "infonce_step_host: null pointer");
  const HostScratch h = carve_host(dev_scratch, rows, D, precision);
  if (h.bytes > dev_scratch_bytes) return set_error(VPA_E_WORKSPACE, "infonce_step_host: scratch %zu < %zu", dev_scratch_bytes, h.bytes);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t mat = (size_t)rows * D * sizeof(float);
  const bool tcp = precision == VPA_PREC_BF16_TC;
  if (h.shards == 1) {        // small batches / fp32 mode: one shot
    VPA_CUDA(cudaMemcpyAsync(h.x1, x1_host, mat, cudaMemcpyHostToDevice, st));
    VPA_CUDA(cudaMemcpyAsync(h.x2, x2_host, mat, cudaMemcpyHostToDevice, st));
    VPA_CUDA(cudaMemcpyAsync(h.lsc, &logit_scale, sizeof(float), cudaMemcpyHostToDevice, st));
    VPA_CUDA(cudaMemcpyAsync(h.gout, &grad_out, sizeof(float), cudaMemcpyHostToDevice, st));
    if (int e = vpa_normalize_pair(h.x1, h.x2, VPA_F32, rows, D, D, D, 0, tcp ? h.ab : nullptr, tcp ? h.tb : nullptr,
                                   h.af, h.tf, h.inv1, h.inv2, h.dcos, tcp ? 1 : 0, st)) return e;
    const void* a = tcp ? h.ab : (const void*)h.af;
    const void* t = tcp ? h.tb : (const void*)h.tf;
    float* rl = h.stats_all, *cl = h.stats_all + rows, *dg = h.stats_all + 2 * rows;
    if (int e = vpa_infonce_fwd(a, t, a, t, precision, rows, rows, D, 0, h.lsc, scale_max, h.dcos, h.ws[0], h.ws_bytes,
                                rl, cl, dg, h.sc, st)) return e;
    if (int e = vpa_infonce_loss(rl, cl, dg, rows, h.loss, st)) return e;
    if (int e = vpa_infonce_bwd(a, t, a, t, precision, rows, rows, D, 0, h.sc, rl, cl, h.gout, h.x1, h.x2, VPA_F32,
                                D, D, h.inv1, h.inv2, 0, h.ws[0], h.ws_bytes, h.dx1, h.dx2, h.dls, st)) return e;
    VPA_CUDA(cudaMemcpyAsync(loss_host, h.loss, sizeof(float), cudaMemcpyDeviceToHost, st));
    if (dlogit_scale_host) VPA_CUDA(cudaMemcpyAsync(dlogit_scale_host, h.dls, sizeof(float), cudaMemcpyDeviceToHost, st));
    if (dx1_host) VPA_CUDA(cudaMemcpyAsync(dx1_host, h.dx1, mat, cudaMemcpyDeviceToHost, st));
    if (dx2_host) VPA_CUDA(cudaMemcpyAsync(dx2_host, h.dx2, mat, cudaMemcpyDeviceToHost, st));
    VPA_CUDA(cudaStreamSynchronize(st));
    return 0;
  }

  HostPipe* hp = nullptr;
  if (int e = host_pipe(&hp)) return e;
  const int V = h.shards;
  const int64_t b = rows / V;
  const size_t shard_elems = (size_t)b * D, shard_bytes = shard_elems * sizeof(float);
  const SweepPlan plan = plan_sweep(b, rows, D, precision);
  const bool from_colpart = pack_reduces_columns(plan, precision);
  const float cap = (scale_max > 0.f) ? scale_max : INFINITY;
  char* ab = static_cast<char*>(h.ab);
  char* tb = static_cast<char*>(h.tb);
  // ---- copy-in: x2 first (every forward shard sweeps all of it), then x1 shard by shard
  VPA_CUDA(cudaEventRecord(hp->start, st));
  VPA_CUDA(cudaStreamWaitEvent(hp->cin, hp->start, 0));
  VPA_CUDA(cudaMemcpyAsync(h.x2, x2_host, mat, cudaMemcpyHostToDevice, hp->cin));
  VPA_CUDA(cudaEventRecord(hp->x2_in, hp->cin));
  for (int k = 0; k < V; ++k) {
    VPA_CUDA(cudaMemcpyAsync(h.x1 + k * shard_elems, x1_host + k * shard_elems, shard_bytes, cudaMemcpyHostToDevice, hp->cin));
    VPA_CUDA(cudaEventRecord(hp->x1_in[k], hp->cin));
  }
  VPA_CUDA(cudaMemcpyAsync(h.lsc, &logit_scale, sizeof(float), cudaMemcpyHostToDevice, st));
  VPA_CUDA(cudaMemcpyAsync(h.gout, &grad_out, sizeof(float), cudaMemcpyHostToDevice, st));
  VPA_CUDA(cudaMemsetAsync(h.loss_counter, 0, 4, st));
  // ---- forward
  VPA_CUDA(cudaStreamWaitEvent(st, hp->x2_in, 0));
  if (int e = normalize_cast_launch(h.x2, VPA_F32, rows, D, D, 0, h.tb, nullptr, h.inv2, st)) return e;
  for (int k = 0; k < V; ++k) {
    const int64_t off = (int64_t)k * b;
    VPA_CUDA(cudaStreamWaitEvent(st, hp->x1_in[k], 0));
    if (int e = normalize_cast_launch(h.x1 + k * shard_elems, VPA_F32, b, D, D, 0, ab + off * D * 2, nullptr, h.inv1 + off, st)) return e;
    if (int e = diag_cos_bf16_launch(ab + off * D * 2, tb + off * D * 2, b, D, h.dcos + off, st)) return e;
    // single-pass sweep of this shard's x1 rows against all x2 rows (reads nothing of the x1 shards still in flight)
    if (int e = fwd_sweep_impl(ab + off * D * 2, tb + off * D * 2, h.ab, h.tb, precision, b, rows, D, off, h.lsc, scale_max,
                               h.ws[k], h.ws_bytes, h.colsum8, true, 1, st, nullptr, !from_colpart)) return e;
  }
  for (int k = 0; k < V; ++k) {        // the exact-regime kernel (device-gated) needs all x1 operands; then the shard's message
    const int64_t off = (int64_t)k * b;
    if (int e = fwd_sweep_impl(ab + off * D * 2, tb + off * D * 2, h.ab, h.tb, precision, b, rows, D, off, h.lsc, scale_max,
                               h.ws[k], h.ws_bytes, h.colsum8, true, 2, st, nullptr, !from_colpart)) return e;
    const Workspace ws = carve_workspace(h.ws[k], b, rows, D, plan);
    if (int e = pack_stats_launch(ws, plan, b, rows, h.lsc, cap, h.dcos + off, 1, h.colsum8, from_colpart,
                                  h.msgs + (size_t)k * (rows + 3 * b), st)) return e;
  }
  if (int e = merge_stats_launch(h.msgs, V, b, rows, h.lsc, cap, 1, h.stats_all, h.sc, h.loss_part, h.loss_counter, h.loss,
                                 st)) return e;
  // ---- backward, shard by shard; gradients leave on the copy-out stream while the next shard computes
  for (int k = 0; k < V; ++k) {
    const int64_t off = (int64_t)k * b;
    if (int e = bwd_impl(ab + off * D * 2, tb + off * D * 2, h.ab, h.tb, precision, b, rows, D, off, h.sc, h.stats_all,
                         h.stats_all + rows, h.gout, h.x1 + k * shard_elems, h.x2 + k * shard_elems, VPA_F32, D, D,
                         h.inv1 + off, h.inv2 + off, 0, h.ws[k], h.ws_bytes, h.dx1 + k * shard_elems, h.dx2 + k * shard_elems,
                         h.dls + k, nullptr, 0, st)) return e;
    VPA_CUDA(cudaEventRecord(hp->bwd_done[k], st));
    VPA_CUDA(cudaStreamWaitEvent(hp->cout, hp->bwd_done[k], 0));
    if (dx1_host) VPA_CUDA(cudaMemcpyAsync(dx1_host + k * shard_elems, h.dx1 + k * shard_elems, shard_bytes, cudaMemcpyDeviceToHost, hp->cout));
    if (dx2_host) VPA_CUDA(cudaMemcpyAsync(dx2_host + k * shard_elems, h.dx2 + k * shard_elems, shard_bytes, cudaMemcpyDeviceToHost, hp->cout));
  }
  float dls_parts[16] = {};
  VPA_CUDA(cudaMemcpyAsync(dls_parts, h.dls, sizeof(float) * V, cudaMemcpyDeviceToHost, hp->cout));
  VPA_CUDA(cudaEventRecord(hp->out_done, hp->cout));
  VPA_CUDA(cudaStreamWaitEvent(st, hp->out_done, 0));        // the caller's stream observes the whole step
  VPA_CUDA(cudaMemcpyAsync(loss_host, h.loss, sizeof(float), cudaMemcpyDeviceToHost, st));
  VPA_CUDA(cudaStreamSynchronize(st));
  if (dlogit_scale_host) {
    double acc = 0.0;
    for (int k = 0; k < V; ++k) acc += (double)dls_parts[k];   // shard order: deterministic
    *dlogit_scale_host = (float)acc;
  }
  return 0;
}

}  // extern "C"
