// Peer-memory transport of the row-sharded step: the gathered operand matrices, the per-rank statistics messages and the
// d logit_scale partials live in a SYMMETRIC segment (same layout on every rank, cudaMalloc + CUDA IPC), and every
// exchange is done by kernels reading / writing the peers' segments over NVLink, published with system-scope epoch flags:
//   * operands: RELAY CTAs (p2p.cuh) -- the first CTAs of a sweep kernel's own grid -- pull the peers' normalised rows with
//     TMA bulk copies through a shared-memory ring (256-row chunks) and raise one local arrival flag per chunk.  The sweep CTAs
//     of the same kernel poll the flags of the tiles ahead of them, so the contraction starts on the local block and consumes
//     remote rows as they land: ONE kernel is the all-gather and the GEMM.  The forward kernel fetches the x2 operands (all it
//     reads; chunk-major over the peers), the backward kernel the x1 operands (read by its second problem only; peer after
//     peer) while its first problem runs -- no wait kernel, no side stream.  The only remote store of the pull is one "my rows
//     are complete" flag per peer and step.
//   * statistics: one kernel writes this rank's message into every peer, waits for the R messages and merges them;
//   * d logit_scale: finalize_bwd stores {epoch, partial} into every rank's slot and sums the R partials in rank order
//     (bitwise identical on every rank).
// No NCCL call on the data path.
//
// Buffer reuse.  Flags carry the step number (monotonic); operands, messages and slots are double-buffered by step parity.
// A rank's step k+1 transfer starts after its statistics kernel (k), which waited for every peer's message(k), which a peer
// sends after its forward sweep(k) and after everything it enqueued before that -- in particular its backward(k-1).  Hence
// when parity (k+1)&1 is overwritten (it last held step k-1), no peer can still be reading step k-1, and nobody can be more
// than one step ahead of anybody else.  The segment therefore keeps exactly two steps; vpa_infonce_bwd_p2p refuses older
// ones.  Every spin is bounded (8 s of %globaltimer) and traps: a missing peer is an error on the stream, not a hang.
#include "p2p.cuh"


#include <cstring>
#include <new>

namespace vpa {

struct SegLayout {
  size_t flags[2];          // [m][world][cpr] uint32: m = 0 x2 operands (t_all), 1 x1 operands (a_all)
  size_t msg_flags;         // [world] uint32
  size_t ready;             // [world] uint32: rank q's operands of epoch e are complete in ITS segment (pull mode)
  size_t dls_slots;         // [2][world] uint64 {epoch << 32 | float bits}
  size_t counters;          // local only: [1] pack arrivals, [1] loss arrivals
  size_t loss_part;         // local only: per-block loss partials (doubles)
  size_t mat[2][2];         // [parity][m]: (B, D) operands
  size_t msgs[2];           // [parity]: (world, B + 3b) floats
  size_t stats_all[2], scale[2], inv1[2], inv2[2], dcos[2];
  size_t colsum8, ws;
  size_t ws_bytes, total;
  int cpr;
};

static SegLayout seg_layout(int64_t b, int world, int D, int precision) {
  SegLayout L{};
  size_t o = 0;
  auto take = [&](size_t bytes) {
    const size_t at = o;
    o += align_up(bytes, 256);
    return at;
  };
  const int64_t B = b * world;
  const size_t es = precision == VPA_PREC_BF16_TC ? 2 : 4;
  L.cpr = (int)((b + kPushRows - 1) / kPushRows);
  for (int m = 0; m < 2; ++m) L.flags[m] = take((size_t)world * L.cpr * 4);
  L.msg_flags = take((size_t)world * 4);
  L.ready = take((size_t)world * 4);
  L.dls_slots = take((size_t)2 * world * 8);
  L.counters = take((size_t)2 * 4);
  L.loss_part = take((size_t)((B + 255) / 256) * 8);
  for (int p = 0; p < 2; ++p) {
    for (int m = 0; m < 2; ++m) L.mat[p][m] = take((size_t)B * D * es);
    L.msgs[p] = take((size_t)world * (B + 3 * b) * 4);
    L.stats_all[p] = take((size_t)3 * B * 4);
    L.scale[p] = take(16);
    L.inv1[p] = take(b * 4);
    L.inv2[p] = take(b * 4);
    L.dcos[p] = take(b * 4);
  }
  L.colsum8 = take((size_t)kColSumSplit * B * 4);
  L.ws_bytes = infonce_workspace_bytes(b, B, D, precision, relay_ctas_default());
  L.ws = take(L.ws_bytes);
  L.total = o;
  return L;
}

struct P2PHandle {
  int rank = 0, world = 1, D = 0, precision = 0, dev = 0;
  int64_t b = 0;
  char* base[kMaxPeers] = {};
  bool opened[kMaxPeers] = {};
  bool connected = false;
  SegLayout L{};
  uint32_t epoch = 0;
  int relay_ctas = 20;      // relay CTAs in front of the forward / backward grids (whole CTA pairs; VPA_P2P_RELAY_CTAS)
  uint32_t a_pending = 0;   // step whose x1 operands are still to be gathered (by its backward, or by the next forward)
};

// ---------------------------------------------------------------- kernels
// The relay as a kernel of its own: shapes the CTA-pair sweeps do not cover (fp32 mode, D not in {256, 512}) gather their
// operands before the forward starts.  Same device code as the relay CTAs of the fused forward kernel.
__global__ void __launch_bounds__(256) p2p_relay_kernel(const RelayArgs A) {
  extern __shared__ uint8_t relay_smem[];
  if (blockIdx.x == 0) relay_signal_ready(A);
  const uint32_t base = relay_smem_u32(relay_smem);
  relay_pull(A, A.m1, blockIdx.x, relay_smem + (((base + 1023u) & ~1023u) - base));
}

// ---------------------------------------------------------------- host side
static P2PView make_view(const P2PHandle* h, uint32_t epoch) {
  P2PView v{};
  for (int q = 0; q < kMaxPeers; ++q) v.base[q] = h->base[q];
  v.rank = h->rank;
  v.world = h->world;
  v.epoch = epoch;
  return v;
}

int p2p_create(int64_t b, int world, int rank, int D, int precision, void** out, void* ipc_handle64) {
  VPA_CHECK_ARG(world >= 2 && world <= kMaxPeers && rank >= 0 && rank < world, "p2p: world must be 2..%d", kMaxPeers);
  VPA_CHECK_ARG(b > 0 && D > 0 && out && ipc_handle64, "p2p_create: bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
  P2PHandle* h = new (std::nothrow) P2PHandle();
  if (!h) return set_error(VPA_E_INVALID, "p2p_create: out of host memory");
  h->rank = rank; h->world = world; h->D = D; h->precision = precision; h->b = b;
  h->L = seg_layout(b, world, D, precision);
  auto fail = [&](cudaError_t e, const char* what) {
    const int rc = set_error((int)e, "p2p_create: %s failed: %s", what, cudaGetErrorString(e));
    if (h->base[rank]) cudaFree(h->base[rank]);
    delete h;
    return rc;
  };
  cudaError_t e;
  if ((e = cudaGetDevice(&h->dev)) != cudaSuccess) return fail(e, "cudaGetDevice");
  void* p = nullptr;
  if ((e = cudaMalloc(&p, h->L.total)) != cudaSuccess) return fail(e, "cudaMalloc");
  h->base[rank] = static_cast<char*>(p);
  cudaIpcMemHandle_t ih;
  if ((e = cudaIpcGetMemHandle(&ih, p)) != cudaSuccess) return fail(e, "cudaIpcGetMemHandle");
  memcpy(ipc_handle64, &ih, 64);
  if ((e = cudaMemset(p, 0, h->L.mat[0][0])) != cudaSuccess) return fail(e, "cudaMemset");      // flags, slots, counters
  h->relay_ctas = relay_ctas_default();      // whole CTA pairs: the forward kernel is launched in clusters of two
  if ((e = cudaDeviceSynchronize()) != cudaSuccess) return fail(e, "cudaDeviceSynchronize");
  *out = h;
  return 0;
}

int p2p_connect(void* handle, const void* all_handles) {
  P2PHandle* h = static_cast<P2PHandle*>(handle);
  VPA_CHECK_ARG(h && all_handles, "p2p_connect: bad argument");
  if (h->connected) return 0;
  for (int q = 0; q < h->world; ++q) {
    if (q == h->rank) continue;
    cudaIpcMemHandle_t ih;
    memcpy(&ih, static_cast<const char*>(all_handles) + (size_t)q * 64, 64);
    void* p = nullptr;
    VPA_CUDA(cudaIpcOpenMemHandle(&p, ih, cudaIpcMemLazyEnablePeerAccess));
    h->base[q] = static_cast<char*>(p);
    h->opened[q] = true;
  }
  h->connected = true;
  return 0;
}

int p2p_destroy(void* handle) {
  P2PHandle* h = static_cast<P2PHandle*>(handle);
  if (!h) return 0;
  cudaDeviceSynchronize();
  for (int q = 0; q < h->world; ++q)
    if (h->opened[q]) cudaIpcCloseMemHandle(h->base[q]);
  if (h->base[h->rank]) cudaFree(h->base[h->rank]);
  delete h;
  return 0;
}

// ---- accessors used by api.cu ------------------------------------------------------------------------------------
int p2p_check(void* handle, int64_t b, int world, int rank, int D, int precision) {
  P2PHandle* h = static_cast<P2PHandle*>(handle);
  VPA_CHECK_ARG(h && h->connected, "p2p: handle not connected");
  VPA_CHECK_ARG(h->b == b && h->world == world && h->rank == rank && h->D == D && h->precision == precision,
                "p2p: handle was created for another shape (b=%lld world=%d rank=%d D=%d precision=%d)", (long long)h->b,
                h->world, h->rank, h->D, h->precision);
  int dev = -1;
  VPA_CUDA(cudaGetDevice(&dev));
  VPA_CHECK_ARG(dev == h->dev, "p2p: handle belongs to device %d, current device is %d", h->dev, dev);
  return 0;
}

uint32_t p2p_next_epoch(void* handle) { return ++static_cast<P2PHandle*>(handle)->epoch; }
uint32_t p2p_current_epoch(void* handle) { return static_cast<P2PHandle*>(handle)->epoch; }

static RelayArgs relay_args(const P2PHandle* h, uint32_t epoch) {
  const SegLayout& L = h->L;
  const int p = (int)(epoch & 1u);
  RelayArgs A{};
  A.v = make_view(h, epoch);
  A.off_mat[0] = L.mat[p][0]; A.off_mat[1] = L.mat[p][1];
  A.off_flags[0] = L.flags[0]; A.off_flags[1] = L.flags[1];
  A.off_ready = L.ready;
  A.b = h->b;
  A.row_bytes = h->D * (h->precision == VPA_PREC_BF16_TC ? 2 : 4);
  A.cpr = L.cpr;
  A.n_ctas = h->relay_ctas;
  A.m0 = 0; A.m1 = 2; A.source_major = 0; A.signal_ready = 1;      // (callers narrow this down)
  return A;
}

P2PStep p2p_step(void* handle, uint32_t epoch) {
  P2PHandle* h = static_cast<P2PHandle*>(handle);
  const SegLayout& L = h->L;
  const int p = (int)(epoch & 1u);
  char* base = h->base[h->rank];
  P2PStep s{};
  s.view = make_view(h, epoch);
  s.t_all = base + L.mat[p][0];
  s.a_all = base + L.mat[p][1];
  s.inv1 = reinterpret_cast<float*>(base + L.inv1[p]);
  s.inv2 = reinterpret_cast<float*>(base + L.inv2[p]);
  s.dcos = reinterpret_cast<float*>(base + L.dcos[p]);
  s.colsum8 = reinterpret_cast<float*>(base + L.colsum8);
  s.msgs = reinterpret_cast<float*>(base + L.msgs[p]);
  s.off_msgs = L.msgs[p];
  s.off_msg_flags = L.msg_flags;
  s.msg_flags = reinterpret_cast<uint32_t*>(base + L.msg_flags);
  s.pack_counter = reinterpret_cast<uint32_t*>(base + L.counters);
  s.loss_counter = s.pack_counter + 1;
  s.loss_part = reinterpret_cast<double*>(base + L.loss_part);
  s.stats_all = reinterpret_cast<float*>(base + L.stats_all[p]);
  s.scale = reinterpret_cast<float*>(base + L.scale[p]);
  s.ws = base + L.ws;
  s.ws_bytes = L.ws_bytes;
  s.off_dls = L.dls_slots + (size_t)p * h->world * 8;
  s.dls_slots = reinterpret_cast<unsigned long long*>(base + s.off_dls);
  s.yflags.flags = reinterpret_cast<const uint32_t*>(base + L.flags[0]);
  s.yflags.rows_per_rank = (int)h->b;
  s.yflags.chunks_per_rank = L.cpr;
  s.yflags.me = h->rank;
  s.yflags.epoch = epoch;
  s.aflags = s.yflags;
  s.aflags.flags = reinterpret_cast<const uint32_t*>(base + L.flags[1]);
  s.relay = relay_args(h, epoch);
  return s;
}

int p2p_relay_ctas(void* handle) { return static_cast<P2PHandle*>(handle)->relay_ctas; }
// step whose x1 operands (read by the backward only) have not been gathered yet; 0: none
uint32_t p2p_a_pending(void* handle) { return static_cast<P2PHandle*>(handle)->a_pending; }
void p2p_set_a_pending(void* handle, uint32_t epoch) { static_cast<P2PHandle*>(handle)->a_pending = epoch; }

// the operand all-gather as a kernel of its own on `st` (shapes without the fused kernels; the x1 operands of a step whose
// backward has not been called when the next forward starts): matrices [m0, 2), everything has landed when it ends
int p2p_relay_standalone(void* handle, uint32_t epoch, int m0, bool signal_ready, cudaStream_t st) {
  P2PHandle* h = static_cast<P2PHandle*>(handle);
  RelayArgs A = relay_args(h, epoch);
  A.m0 = m0;
  A.signal_ready = signal_ready ? 1 : 0;
  const int items = (2 - m0) * A.cpr * (h->world - 1);
  A.n_ctas = items < 32 ? items : 32;
  static SmemAttrCache attr_cache;
  if (int e = ensure_dynamic_smem(attr_cache, p2p_relay_kernel, (int)kRelaySmemBytes + 1024)) return e;
  prof_begin(PROF_PUSH, st);
  VPA_CUDA(launch_kernel(p2p_relay_kernel, dim3(A.n_ctas), dim3(256), kRelaySmemBytes + 1024, st, A));
  prof_end(PROF_PUSH, st);
  VPA_LAUNCH_CHECK("p2p_relay_kernel");
  return 0;
}

}  // namespace vpa
