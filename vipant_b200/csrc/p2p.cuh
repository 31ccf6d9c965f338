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
  int rank, world;
  uint32_t epoch;                     // step number (monotonic, starts at 1); flags carry the epoch of the data they publish
};

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

// the same test without waiting: are global rows [r0, r1) of the gathered matrix complete?
__device__ __forceinline__ bool p2p_rows_ready(const P2PRowFlags& f, int r0, int r1) {
  int row = r0;
  while (row < r1) {
    const int src = row / f.rows_per_rank;
    const int off = row - src * f.rows_per_rank;
    const int c = off / kPushRows;
    if (src != f.me && (int32_t)(ld_acquire_sys_u32(f.flags + src * f.chunks_per_rank + c) - f.epoch) < 0) return false;
    row = src * f.rows_per_rank + min(f.rows_per_rank, (c + 1) * kPushRows);
  }
  return true;
}

// ---------------------------------------------------------------- operand relay (the all-gather of the row-sharded step)
// The gathered operand matrices are filled by RELAY CTAs: whole CTAs that do nothing but move rows, driven by one thread
// and the TMA engine.  They run as the first CTAs of the single-pass forward's grid (infonce_pair.cu: one kernel does the
// all-gather and the contraction; CTAs are dispatched in blockIdx order, so the relays are resident before any sweep CTA
// that polls their flags) or as a kernel of their own for the shapes the CTA-pair sweeps do not cover.
// A relay pulls: cp.async.bulk peer-global -> shared-memory ring -> cp.async.bulk local-global, up to kRelaySlots x 32 KB of
// NVLink reads in flight per CTA; one local arrival flag per 256-row chunk once its stores have completed.  (Measured at
// 8 GPUs, all pulling from all: 12 / 20 / 28 relay CTAs move 243 / 305 / 335 GB/s into one GPU.  An NVSwitch-multicast
// variant -- every rank storing its own rows once with multimem.st -- was brought up and measured the same or slower
// (profiles/r02_scaling.md) and was removed, as were the LDG pull / push kernels and the copy-engine variant of round 1.)
// Items are ordered matrix-major (x2 operands first: the forward needs only them), then chunk-major over the peers, so
// the early chunks of every peer block land first -- the order in which the sweep visits its tiles.
struct RelayArgs {
  P2PView v;
  size_t off_mat[2], off_flags[2], off_ready;      // this step's parity; m = 0: x2 operands (t_all), 1: x1 operands (a_all)
  int64_t b;                                       // rows per rank
  int row_bytes, cpr;                              // bytes per operand row; 256-row chunks per rank block
  int n_ctas;                                      // relay CTAs (0: no relay)
  int m0, m1;                                      // matrices [m0, m1) to move
  int source_major;                                // item order: 0 chunk-major over the peers, 1 peer after peer (me+1, me+2, ...)
  int signal_ready;                                // 1: this launch announces "my rows of this step are complete" to the peers
};
constexpr int kRelaySlots = 6;
constexpr int kRelayPieceBytes = 32768;
constexpr uint32_t kRelaySmemBytes = kRelaySlots * kRelayPieceBytes + 64;      // ring + mbarriers (1024-byte aligned base)

struct RelayItem { int m, src, c, row0, rows; };
// item -> (matrix, source rank, chunk): shared by the device code and, host side, vpa_debug_relay_item (CPU test).
// chunk-major: chunk k of every peer before chunk k+1 of any (the forward sweep visits its tiles in that order);
// source-major: all chunks of peer me+1, then me+2, ... (the backward sweep starts on the local block and walks the rank
// blocks in that order).  Either way the ranks start on different peers: nobody's egress is a hot spot.
__host__ __device__ inline RelayItem relay_item_decode(int item, int m0, int source_major, int world, int me, int cpr, int64_t b) {
  const int per_m = cpr * (world - 1);
  RelayItem it;
  it.m = m0 + item / per_m;
  const int r = item % per_m;
  int q;
  if (source_major) {
    q = r / cpr + 1;
    it.c = r - (q - 1) * cpr;
  } else {
    it.c = r / (world - 1);
    q = r - it.c * (world - 1) + 1;
  }
  it.src = (me + q) % world;
  const int64_t row0 = (int64_t)it.c * kPushRows;
  const int64_t left = b - row0;
  it.row0 = (int)row0;
  it.rows = left < kPushRows ? (int)left : kPushRows;
  return it;
}

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t relay_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool relay_mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  return done != 0;
}
static __device__ __noinline__ void relay_timeout(int what, int a, int b) {
  printf("vipant_b200(relay): wait %d timed out (block %d, %d %d)\n", what, blockIdx.x, a, b);
  asm volatile("trap;");
}

// "my rows of this step are complete in my segment" -> every peer (called by ONE relay CTA; the normalise kernel that wrote
// the rows finished before this kernel started: stream order)
__device__ __forceinline__ void relay_signal_ready(const RelayArgs& A) {
  const int q = threadIdx.x;
  if (A.signal_ready && q < A.v.world && q != A.v.rank) {
    __threadfence_system();
    st_release_sys_u32(reinterpret_cast<uint32_t*>(A.v.base[q] + A.off_ready) + A.v.rank, A.v.epoch);
  }
}

// Pull role of relay CTA `cta` of `A.n_ctas`: matrices [A.m0, m1).  `ring` = 1024-byte aligned shared memory of
// kRelaySmemBytes.  One thread drives the TMA engine; the caller lets the other threads of the CTA leave.  Chunks whose
// arrival flag already carries this step's number are skipped (the backward's relay finds the x1 operands in place when
// the forward had to fetch them for the exact-regime kernel).
__device__ __forceinline__ void relay_pull(const RelayArgs& A, int m1, int cta, uint8_t* ring) {
  if (threadIdx.x != 0) return;
  const int me = A.v.rank, world = A.v.world;
  char* mine = A.v.base[me];
  const uint32_t ring_u = relay_smem_u32(ring);
  const uint32_t bars = ring_u + kRelaySlots * kRelayPieceBytes;
  for (int s = 0; s < kRelaySlots; ++s)
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bars + 8u * s) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  const uint32_t* ready = reinterpret_cast<const uint32_t*>(mine + A.off_ready);
  const int total = (m1 - A.m0) * A.cpr * (world - 1);
  const int piece_rows = kRelayPieceBytes / A.row_bytes > 0 ? kRelayPieceBytes / A.row_bytes : 1;
  unsigned seen = 1u << me;                        // sources whose ready flag this thread has already observed
  struct Cursor { int item, piece, npieces; RelayItem it; };
  auto open = [&](Cursor& c) {                     // position on the first piece of the next item that is still missing
    while (c.item < total) {
      c.it = relay_item_decode(c.item, A.m0, A.source_major, world, me, A.cpr, A.b);
      const uint32_t have = ld_acquire_sys_u32(reinterpret_cast<const uint32_t*>(mine + A.off_flags[c.it.m]) + c.it.src * A.cpr + c.it.c);
      if ((int32_t)(have - A.v.epoch) < 0) {       // (both cursors see the same flags: only this thread sets them, after use)
        c.npieces = (c.it.rows + piece_rows - 1) / piece_rows;
        c.piece = 0;
        return;
      }
      c.item += A.n_ctas;
    }
  };
  auto advance = [&](Cursor& c) {
    if (++c.piece == c.npieces) { c.item += A.n_ctas; open(c); }
  };
  auto geometry = [&](const Cursor& c, size_t* off, uint32_t* bytes) {
    const int r0 = c.it.row0 + c.piece * piece_rows;
    const int rows = min(piece_rows, c.it.row0 + c.it.rows - r0);
    *off = A.off_mat[c.it.m] + ((size_t)c.it.src * A.b + r0) * A.row_bytes;
    *bytes = (uint32_t)rows * A.row_bytes;
  };
  Cursor ld{}, st{};
  ld.item = st.item = cta;
  open(ld);
  open(st);
  uint32_t n_ld = 0, n_st = 0;                     // pieces issued / retired (slot = n % kRelaySlots)
  auto issue_load = [&]() {
    if (!(seen >> ld.it.src & 1u)) { p2p_wait_ge(ready + ld.it.src, A.v.epoch); seen |= 1u << ld.it.src; }
    size_t off; uint32_t bytes;
    geometry(ld, &off, &bytes);
    const uint32_t slot = n_ld % kRelaySlots, bar = bars + 8u * slot, dst = ring_u + slot * kRelayPieceBytes;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(A.v.base[ld.it.src] + off), "r"(bytes), "r"(bar) : "memory");
    ++n_ld;
    advance(ld);
  };
  while (ld.item < total && n_ld < (uint32_t)kRelaySlots) issue_load();
  while (st.item < total) {
    const uint32_t slot = n_st % kRelaySlots, bar = bars + 8u * slot, parity = (n_st / kRelaySlots) & 1u;
    if (!relay_mbar_try(bar, parity)) {
      const unsigned long long t0 = global_timer_ns();
      uint32_t spins = 0;
      while (!relay_mbar_try(bar, parity))
        if ((++spins & 255u) == 0 && global_timer_ns() - t0 > kP2PTimeoutNs) relay_timeout(0, st.item, st.piece);
    }
    size_t off; uint32_t bytes;
    geometry(st, &off, &bytes);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(mine + off), "r"(ring_u + slot * kRelayPieceBytes), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    const bool chunk_done = st.piece + 1 == st.npieces;
    const int m = st.it.m, src = st.it.src, c = st.it.c;
    ++n_st;
    advance(st);
    if (chunk_done) {
      // the chunk is complete in local memory once its stores are: publish its flag (the loads of the next pieces are
      // tracked by mbarriers and stay in flight meanwhile)
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      asm volatile("fence.proxy.async;" ::: "memory");            // async-proxy writes -> ordered before the flag store
      __threadfence();
      st_release_sys_u32(reinterpret_cast<uint32_t*>(mine + A.off_flags[m]) + src * A.cpr + c, A.v.epoch);
    } else {
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the slot may be overwritten
    }
    if (ld.item < total) issue_load();
  }
}

#endif  // __CUDACC__

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
  P2PRowFlags aflags;       // chunk flags of the x1 operands (the Y stream of the backward's second problem)
  RelayArgs relay;          // the operand all-gather of this step (n_ctas: relay CTAs in front of the forward / backward grid)
};

int p2p_create(int64_t b, int world, int rank, int D, int precision, void** out, void* ipc_handle64);
int p2p_connect(void* handle, const void* all_handles);
int p2p_destroy(void* handle);
int p2p_check(void* handle, int64_t b, int world, int rank, int D, int precision);
uint32_t p2p_next_epoch(void* handle);
uint32_t p2p_current_epoch(void* handle);
P2PStep p2p_step(void* handle, uint32_t epoch);
int p2p_relay_standalone(void* handle, uint32_t epoch, int m0, bool signal_ready, cudaStream_t st);
int p2p_relay_ctas(void* handle);
uint32_t p2p_a_pending(void* handle);
void p2p_set_a_pending(void* handle, uint32_t epoch);

}  // namespace vpa
