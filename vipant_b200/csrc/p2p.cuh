// Peer-memory transport of the row-sharded step (one process per GPU, NVLink / NVSwitch peer mappings):
// device-side view of the symmetric segment + system-scope flag helpers.  Every rank allocates the SAME layout
// (p2p.cu::seg_layout), so an offset is valid in every peer's segment; base[q] is rank q's segment as mapped here.
#pragma once

#include <cstdio>

#include "common.cuh"

namespace vpa {

constexpr int kMaxPeers = 8;          // one NVSwitch node
constexpr int kPushRows = 256;        // flag granularity of the operand push == Y tile height of the pair kernels

struct P2PView {
  char* base[kMaxPeers];              // base[rank] is the local segment
  char* mc;                           // NVLS transport: multicast mapping of the segment (a store lands in EVERY rank's
                                      // copy at the same offset, replicated by the NVSwitch); nullptr otherwise
  int rank, world;
  uint32_t epoch;                     // step number (monotonic, starts at 1); flags carry the epoch of the data they publish
};

// stores to a multicast address (multimem.st: one store, every replica)
__device__ __forceinline__ void mc_st_f32(float* p, float v) {
  asm volatile("multimem.st.weak.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ void mc_st_v4(uint4* p, uint4 v) {
  asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(__uint_as_float(v.x)),
               "f"(__uint_as_float(v.y)), "f"(__uint_as_float(v.z)), "f"(__uint_as_float(v.w)) : "memory");
}
__device__ __forceinline__ void mc_st_release_sys_u32(uint32_t* p, uint32_t v) {
  asm volatile("multimem.st.release.sys.global.f32 [%0], %1;" ::"l"(p), "f"(__uint_as_float(v)) : "memory");
}
__device__ __forceinline__ void mc_st_release_sys_u64(unsigned long long* p, unsigned long long v) {      // one 8-byte element
  asm volatile("multimem.st.release.sys.global.f64 [%0], %1;" ::"l"(p), "d"(__longlong_as_double((long long)v)) : "memory");
}

__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
constexpr unsigned long long kP2PTimeoutNs = 8000000000ull;      // a peer that never shows up is an error, not a hang

static __device__ __noinline__ void p2p_timeout(const void* flag, uint32_t want, uint32_t have) {
  printf("vipant_b200(p2p): peer flag %p stuck at %u (waiting for %u), block %d thread %d\n", flag, have, want, blockIdx.x,
         threadIdx.x);
  asm volatile("trap;");
}
// spin until *flag >= want (flags are monotonic epochs)
__device__ __forceinline__ void p2p_wait_ge(const uint32_t* flag, uint32_t want) {
  uint32_t v = ld_acquire_sys_u32(flag);
  if ((int32_t)(v - want) >= 0) return;
  const unsigned long long t0 = global_timer_ns();
  uint32_t spins = 0;
  while (true) {
    v = ld_acquire_sys_u32(flag);
    if ((int32_t)(v - want) >= 0) return;
    if ((++spins & 1023u) == 0 && global_timer_ns() - t0 > kP2PTimeoutNs) p2p_timeout(flag, want, v);
  }
}

// What the single-pass forward needs to consume operand rows as they land: flags[src * chunks_per_rank + c] >= epoch
// once rows [c*kPushRows, ...) of rank src's block are complete in the local copy of the gathered matrix.
// (struct P2PRowFlags is declared in common.cuh: it travels inside SweepArgs)
// wait until global rows [r0, r1) of the gathered matrix are complete (one thread)
__device__ __forceinline__ void p2p_wait_rows(const P2PRowFlags& f, int r0, int r1) {
  int row = r0;
  while (row < r1) {
    const int src = row / f.rows_per_rank;
    const int off = row - src * f.rows_per_rank;
    const int c = off / kPushRows;
    if (src != f.me) p2p_wait_ge(f.flags + src * f.chunks_per_rank + c, f.epoch);
    const int next = src * f.rows_per_rank + min(f.rows_per_rank, (c + 1) * kPushRows);
    row = next;
  }
}

// One item of the pull kernel: `rows` rows starting at row `row0` of rank `src`'s block of matrix `m` (chunk `c`).
// Items are ordered matrix-major (x2 operands first), then chunk-major, then slice-major, then over the peers: the first
// (world-1)*slices items are chunk 0 of EVERY peer, so with ~64 CTAs taking items round-robin the early chunks of all peer
// blocks land first and the forward sweep (which visits its tiles chunk-major) finds data a few microseconds after it
// starts, instead of waiting for whole 256-row chunks that one CTA needs ~45 us to pull.  A chunk is complete when its
// `slices` items are (arrival counter).  Shared by the kernel and, host side, vpa_debug_pull_item (CPU test).
struct PullItem { int m, src, c, row0, rows; };
__host__ __device__ inline PullItem pull_item_decode(int item, int m0, int world, int me, int cpr, int slices, int64_t b) {
  const int per_c = (world - 1) * slices;
  const int per_m = cpr * per_c;
  PullItem it;
  it.m = m0 + item / per_m;
  int r = item % per_m;
  it.c = r / per_c;
  r -= it.c * per_c;
  const int s = r / (world - 1);
  const int q = r - s * (world - 1) + 1;
  it.src = (me + q) % world;
  const int64_t crow0 = (int64_t)it.c * kPushRows;
  const int64_t left = b - crow0;
  const int crows = left < kPushRows ? (int)left : kPushRows;
  const int per = (crows + slices - 1) / slices;
  const int lo = s * per < crows ? s * per : crows;
  const int hi = lo + per < crows ? lo + per : crows;
  it.row0 = (int)crow0 + lo;
  it.rows = hi - lo;
  return it;
}

// Host-side description of one step's buffers inside the local segment (p2p.cu::p2p_step) + what kernels need to reach
// the same buffers in the peers (offsets are identical in every segment).
struct P2PStep {
  P2PView view;
  void *a_all, *t_all;
  float *inv1, *inv2, *dcos, *colsum8, *msgs, *stats_all, *scale;
  size_t off_msgs, off_msg_flags, off_dls;
  uint32_t *msg_flags, *pack_counter;
  unsigned long long* dls_slots;
  double* loss_part;
  uint32_t* loss_counter;
  void* ws;
  size_t ws_bytes;
  P2PRowFlags yflags;       // chunk flags of the x2 operands (the Y stream of the single-pass forward)
};

int p2p_create(int64_t b, int world, int rank, int D, int precision, void** out, void* ipc_handle64);
int p2p_connect(void* handle, const void* all_handles);
int p2p_destroy(void* handle);
int p2p_check(void* handle, int64_t b, int world, int rank, int D, int precision);
uint32_t p2p_next_epoch(void* handle);
uint32_t p2p_current_epoch(void* handle);
P2PStep p2p_step(void* handle, uint32_t epoch);
int p2p_push_operands(void* handle, uint32_t epoch, cudaStream_t st);
int p2p_pull_rest(void* handle, uint32_t epoch, cudaStream_t st);
int p2p_join_push(void* handle, cudaStream_t st);
int p2p_wait_operands(void* handle, uint32_t epoch, const float* gate_scale, float scale_cap, cudaStream_t st);
int p2p_dls_sum(const P2PStep& s, float* dlogit_scale, cudaStream_t st);
int p2p_mode(void* handle);
int p2p_nvls_export(void* handle, int* fd_out);
int p2p_nvls_attach(void* handle, int fd);
int p2p_nvls_bind(void* handle);

}  // namespace vpa
