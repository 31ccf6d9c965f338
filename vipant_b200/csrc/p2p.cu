// Peer-memory transport of the row-sharded step: the gathered operand matrices, the per-rank statistics messages and the
// d logit_scale partials live in a SYMMETRIC segment (same layout on every rank, cudaMalloc + CUDA IPC), and every
// exchange is done by kernels reading / writing the peers' segments over NVLink, published with system-scope epoch flags:
//   * operands (default: pull): a kernel on a high-priority side stream loads the peers' normalised rows (x2 operands
//     first, 256-row chunks, chunk-major over the peers) and stores them locally; one local arrival flag per chunk.  The
//     single-pass forward's TMA producer polls the flag of every tile before loading it, so the sweep starts on the local
//     block and consumes remote rows as they land -- the all-gather overlaps the contraction tile by tile.  The only remote
//     store of the pull is one "my rows are complete" flag per peer and step.  Alternatives kept for A/B (VPA_P2P_MODE):
//     push (stores to all peers + fence.sys per batch), stream (one chunk to one peer per CTA), ce (copy engines);
//     VPA_P2P_PLAN=serial moves the x2 operands before and the x1 operands after the forward instead of beside it.
//   * statistics: pack_stats writes its message into every peer, merge_stats waits for the R flags;
//   * d logit_scale: finalize_bwd stores {epoch, partial} into every peer's slot, a one-warp kernel sums them in rank order
//     (bitwise identical on every rank).
// No NCCL call on the data path.
//
// Buffer reuse.  Flags carry the step number (monotonic); operands, messages and slots are double-buffered by step parity.
// A rank's step k+1 transfer starts after its merge_stats(k), which waited for every peer's message(k), which a peer sends
// after its forward sweep(k) and after everything it enqueued before that -- in particular its backward(k-1).  Hence when
// parity (k+1)&1 is overwritten (it last held step k-1), no peer can still be reading step k-1, and nobody can be more
// than one step ahead of anybody else.  The segment therefore keeps exactly two steps; vpa_infonce_bwd_p2p refuses older
// ones.  Every spin is bounded (8 s of %globaltimer) and traps: a missing peer is an error on the stream, not a hang.
#include "p2p.cuh"

#include <cuda.h>
#include <unistd.h>

#include <cstring>
#include <new>

namespace vpa {

struct SegLayout {
  size_t flags[2];          // [m][world][cpr] uint32: m = 0 x2 operands (t_all), 1 x1 operands (a_all)
  size_t msg_flags;         // [world] uint32
  size_t ready;             // [world] uint32: rank q's operands of epoch e are complete in ITS segment (pull mode)
  size_t dls_slots;         // [2][world] uint64 {epoch << 32 | float bits}
  size_t counters;          // local only: [2][cpr] push arrivals, [1] pack arrivals, [1] loss arrivals
  size_t pull_counters;     // local only: [2][world][cpr] slice arrivals of the pull kernel
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
  L.counters = take((size_t)(2 * L.cpr + 2) * 4);
  L.pull_counters = take((size_t)2 * world * L.cpr * 4);
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
  L.ws_bytes = vpa_infonce_workspace_bytes(b, B, D, precision);
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
  cudaStream_t side = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
  bool join_pending = false;
  int push_groups = 8, push_ctas = 4, push_rounds = 8;
  int pull = 1, pull_ctas = 64;      // pull: 0 = push kernel, 1 = pull kernel, 2 = copy engines, 3 = streaming push kernel,
                                     //       4 = NVLS multicast stores (segment allocated with the VMM API, see below)
  CUmemGenericAllocationHandle vmm_mem = 0, vmm_mc = 0;
  CUdeviceptr uc_va = 0, mc_va = 0;
  size_t vmm_size = 0;
  bool mc_added = false, mc_bound = false;
  int mc_ctas = 24;
  int stream_ctas = 96;
  cudaStream_t side2 = nullptr;
  cudaEvent_t ready_ev = nullptr, join2 = nullptr;
  int ce_streams = 2;
  int pull_slices = 4;      // items per 256-row chunk of the pull (VPA_P2P_PULL_SLICES); 1 = one CTA pulls a whole chunk
  int pull_threads = 256;   // threads per pull CTA (VPA_P2P_PULL_THREADS: lighter CTAs when many SMs pull)
  int strong_ld = 1;        // system-scope relaxed loads of peer rows (VPA_P2P_PULL_LD=weak: L1::no_allocate weak loads, same speed at N=2)
  int serial = 0, pull_ctas_alone = 148;    // serial plan: x2 operands alone before the forward, x1 operands after it
  cudaEvent_t t_done = nullptr, fwd_done = nullptr;
  size_t ce_bytes = 4u << 20;
};

// ---------------------------------------------------------------- kernels
struct PushArgs {
  P2PView v;
  size_t off_mat[2];        // m = 0: t_all, 1: a_all (this step's parity)
  size_t off_flags[2];
  size_t off_counters;
  int64_t b;
  int row_bytes, cpr, groups, ctas_per_group, batch;
};

constexpr int kPushUnroll = 8;      // 16-byte loads in flight per thread: the copy is bound by L2 load latency otherwise

__global__ void __launch_bounds__(256) p2p_push_kernel(const PushArgs A) {
  const int g = blockIdx.x / A.ctas_per_group, cg = blockIdx.x - g * A.ctas_per_group;
  const int nthreads = A.ctas_per_group * blockDim.x;
  const int tid = cg * blockDim.x + threadIdx.x;
  char* mine = A.v.base[A.v.rank];
  uint32_t* counters = reinterpret_cast<uint32_t*>(mine + A.off_counters);
  const int nbatch = (A.cpr + A.batch - 1) / A.batch;         // batches of `batch` chunks per matrix: one fence each
  for (int item = g; item < 2 * nbatch; item += A.groups) {
    const int m = item / nbatch, k = item - m * nbatch;        // all x2-operand batches first: the forward needs them first
    const int c0 = k * A.batch, c1 = min(A.cpr, c0 + A.batch);
    const int64_t row0 = (int64_t)c0 * kPushRows;
    const int rows = (int)min((int64_t)(c1 - c0) * kPushRows, A.b - row0);
    const int n16 = rows * (A.row_bytes / 16);
    const size_t off = A.off_mat[m] + ((size_t)A.v.rank * A.b + row0) * A.row_bytes;
    const uint4* src = reinterpret_cast<const uint4*>(mine + off);
    for (int i = tid; i < n16; i += nthreads * kPushUnroll) {
      uint4 val[kPushUnroll];
#pragma unroll
      for (int u = 0; u < kPushUnroll; ++u) {
        const int idx = i + u * nthreads;
        if (idx < n16) val[u] = __ldg(src + idx);
      }
      for (int q = 1; q < A.v.world; ++q) {
        const int peer = (A.v.rank + q) % A.v.world;           // rotate so that the ranks do not all hit the same target
        uint4* dst = reinterpret_cast<uint4*>(A.v.base[peer] + off);
#pragma unroll
        for (int u = 0; u < kPushUnroll; ++u) {
          const int idx = i + u * nthreads;
          if (idx < n16) dst[idx] = val[u];
        }
      }
    }
    __threadfence_system();                                     // this thread's peer stores are performed
    __syncthreads();
    if (threadIdx.x == 0) {
      const uint32_t old = atomicAdd(&counters[m * A.cpr + c0], 1u);
      if ((old + 1) % (uint32_t)A.ctas_per_group == 0) {        // last CTA of the group for this batch: publish its chunks
        __threadfence();
        for (int q = 1; q < A.v.world; ++q) {
          const int peer = (A.v.rank + q) % A.v.world;
          uint32_t* fl = reinterpret_cast<uint32_t*>(A.v.base[peer] + A.off_flags[m]) + A.v.rank * A.cpr;
          for (int c = c0; c < c1; ++c) st_release_sys_u32(fl + c, A.v.epoch);
        }
      }
    }
    __syncthreads();
  }
}

// ---- operands, consumer driven: every rank PULLS its peers' rows (loads over NVLink, stores to local memory) ------------
// No system-scope fence per chunk (the puller knows when its own loads have returned) and the only remote store is one
// "my rows are complete" flag per peer and step; a chunk's arrival flag is a local store.  Items are ordered x2 operands
// first, chunk-major over the peers, so the forward sweep finds the early tiles of every peer block first.
struct PullArgs {
  P2PView v;
  size_t off_mat[2], off_flags[2], off_ready;
  int64_t b;
  int row_bytes, cpr;
  int m0, m1;               // matrices [m0, m1) of {0: x2 operands, 1: x1 operands}; the ready signal goes out with m0 == 0
  int slices;               // pull kernel: items per chunk (see pull_item_decode)
  size_t off_pull_counters;
};
constexpr int kPullUnroll = 8;

// Loads of peer rows.  Peer addresses bypass the local L2 and are cached by the local L1 only (B300_MICROARCH.md).
// STRONG = 1 (default): system-scope relaxed loads, never served by L1 -- a line cached two steps ago (same buffer parity)
// cannot come back.  STRONG = 0: weak loads that do not allocate in L1 (every address is read once per kernel and L1 is
// invalidated at kernel boundaries); measured equally fast at N = 2 (72 vs 76 us for 32 MiB), kept for A/B.
template <int STRONG>
__device__ __forceinline__ uint4 ld_peer_v4(const uint4* p) {
  uint4 v;
  if (STRONG)
    asm volatile("ld.relaxed.sys.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  else
    asm volatile("ld.global.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}

// Register budget: a pull CTA must fit beside a resident single-pass forward CTA (320 threads x 104 registers = 33280 of
// the SM's 65536), or ranks polling for rows could starve the kernels that deliver them: <= 126 registers at 256 threads.
// __maxnreg__(96) pins that (ptxas -v: 80 registers, no spills); p2p_create re-checks the budget of the actual build.
template <int STRONG, int THREADS>
__global__ void __maxnreg__(96) p2p_pull_kernel(const PullArgs A) {
  const int me = A.v.rank, world = A.v.world;
  char* mine = A.v.base[me];
  if (A.m0 == 0 && blockIdx.x == 0 && (int)threadIdx.x < world && (int)threadIdx.x != me) {
    // the normalise kernel that wrote this rank's rows finished before this kernel started (stream order)
    __threadfence_system();
    st_release_sys_u32(reinterpret_cast<uint32_t*>(A.v.base[threadIdx.x] + A.off_ready) + me, A.v.epoch);
  }
  const uint32_t* ready = reinterpret_cast<const uint32_t*>(mine + A.off_ready);
  uint32_t* arrivals = reinterpret_cast<uint32_t*>(mine + A.off_pull_counters);
  const int total = (A.m1 - A.m0) * A.cpr * (world - 1) * A.slices;
  for (int item = blockIdx.x; item < total; item += gridDim.x) {
    const PullItem it = pull_item_decode(item, A.m0, world, me, A.cpr, A.slices, A.b);
    const int m = it.m, c = it.c, src = it.src;
    if (threadIdx.x == 0) p2p_wait_ge(ready + src, A.v.epoch);
    __syncthreads();
    const int n16 = it.rows * (A.row_bytes / 16);
    const size_t off = A.off_mat[m] + ((size_t)src * A.b + it.row0) * A.row_bytes;
    const uint4* from = reinterpret_cast<const uint4*>(A.v.base[src] + off);
    uint4* to = reinterpret_cast<uint4*>(mine + off);
    // 8 x 16 B per thread in flight; a deeper software pipeline measured no faster (the rate is set by the requests an SM
    // can keep outstanding over NVLink, not by the per-thread dependency chain) and costs registers the forward needs
    constexpr int nthr = THREADS, stride = nthr * kPullUnroll;
    for (int i = threadIdx.x; i < n16; i += stride) {
      uint4 val[kPullUnroll];
#pragma unroll
      for (int u = 0; u < kPullUnroll; ++u) {
        const int idx = i + u * nthr;
        if (idx < n16) val[u] = ld_peer_v4<STRONG>(from + idx);
      }
#pragma unroll
      for (int u = 0; u < kPullUnroll; ++u) {
        const int idx = i + u * nthr;
        if (idx < n16) to[idx] = val[u];
      }
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
      // the chunk is complete when all its slices are: monotonic arrival counter (every step adds `slices`), last one publishes
      const uint32_t old = atomicAdd(&arrivals[(m * world + src) * A.cpr + c], 1u);
      if ((old + 1) % (uint32_t)A.slices == 0) {
        __threadfence();
        st_release_sys_u32(reinterpret_cast<uint32_t*>(mine + A.off_flags[m]) + src * A.cpr + c, A.v.epoch);
      }
    }
  }
}

// ---- operands, producer driven, streaming variant (VPA_P2P_MODE=stream): posted stores are not bound by the number of
// outstanding read requests an SM can hold.  One item = one 256-row chunk to ONE peer, handled by one CTA start to finish:
// no cross-CTA counters, one system fence per item, then the chunk's arrival flag in that peer.  Items are ordered
// x2 operands first, chunk-major, peers rotating; CTAs take them round-robin.
__global__ void __launch_bounds__(128) p2p_stream_push_kernel(const PullArgs A) {
  const int me = A.v.rank, world = A.v.world;
  char* mine = A.v.base[me];
  const int per_m = (world - 1) * A.cpr, total = (A.m1 - A.m0) * per_m;
  for (int item = blockIdx.x; item < total; item += gridDim.x) {
    const int m = A.m0 + item / per_m, r = item % per_m;
    const int c = r / (world - 1), q = r - c * (world - 1) + 1;
    const int dst = (me + q) % world;
    const int64_t row0 = (int64_t)c * kPushRows;
    const int rows = (int)min((int64_t)kPushRows, A.b - row0);
    const int n16 = rows * (A.row_bytes / 16);
    const size_t off = A.off_mat[m] + ((size_t)me * A.b + row0) * A.row_bytes;
    const uint4* from = reinterpret_cast<const uint4*>(mine + off);
    uint4* to = reinterpret_cast<uint4*>(A.v.base[dst] + off);
    constexpr int kStride = 128 * kPullUnroll;
    for (int i = threadIdx.x; i < n16; i += kStride) {
      uint4 val[kPullUnroll];
#pragma unroll
      for (int u = 0; u < kPullUnroll; ++u) {
        const int idx = i + u * 128;
        if (idx < n16) val[u] = __ldg(from + idx);
      }
#pragma unroll
      for (int u = 0; u < kPullUnroll; ++u) {
        const int idx = i + u * 128;
        if (idx < n16) to[idx] = val[u];
      }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0)
      st_release_sys_u32(reinterpret_cast<uint32_t*>(A.v.base[dst] + A.off_flags[m]) + me * A.cpr + c, A.v.epoch);
  }
}

// ---- operands by the COPY ENGINES: no SM, no issue slots and no L2->SM bandwidth taken from the sweep that runs meanwhile --
// signal "my rows are complete" to every peer / wait for every peer's signal / publish the chunks a finished copy delivered
__global__ void p2p_signal_ready_kernel(const P2PView v, size_t off_ready) {
  const int q = threadIdx.x;
  if (q < v.world && q != v.rank) {
    __threadfence_system();
    st_release_sys_u32(reinterpret_cast<uint32_t*>(v.base[q] + off_ready) + v.rank, v.epoch);
  }
}
__global__ void p2p_wait_ready_kernel(const uint32_t* __restrict__ ready, int world, int me, uint32_t epoch) {
  const int q = threadIdx.x;
  if (q < world && q != me) p2p_wait_ge(ready + q, epoch);
}
__global__ void p2p_set_flags_kernel(uint32_t* flags, int n, uint32_t epoch) {
  if ((int)threadIdx.x < n) st_release_sys_u32(flags + threadIdx.x, epoch);
}

// every operand chunk of every peer has landed (both matrices): what all later kernels of the step rely on
// gate_scale != nullptr: only in the exact two-sweep regime (s * log2e > limit, decided on the device like the sweeps do);
// the single-pass forward has consumed every x2-operand chunk itself and reads no x1 operands of the peers.
__global__ void p2p_wait_all_kernel(const uint32_t* __restrict__ flags_t, const uint32_t* __restrict__ flags_a, int world,
                                    int cpr, int me, uint32_t epoch, const float* __restrict__ gate_scale, float scale_cap,
                                    float s2_limit) {
  if (gate_scale != nullptr && fminf(expf(*gate_scale), scale_cap) * kLog2e <= s2_limit) return;
  const int n = world * cpr;
  for (int i = threadIdx.x; i < 2 * n; i += blockDim.x) {
    const int m = i >= n, k = i - m * n;
    if (k / cpr == me) continue;
    p2p_wait_ge((m ? flags_a : flags_t) + k, epoch);
  }
}

// d logit_scale = sum over ranks (rank order) of the partials the finalize kernels stored into this rank's slots
__global__ void p2p_dls_sum_kernel(const unsigned long long* __restrict__ slots, int world, uint32_t epoch,
                                   float* __restrict__ out) {
  const int lane = threadIdx.x;
  float v = 0.f;
  if (lane < world) {
    const unsigned long long t0 = global_timer_ns();
    uint32_t spins = 0;
    while (true) {
      const unsigned long long w = ld_acquire_sys_u64(slots + lane);
      if ((uint32_t)(w >> 32) == epoch) { v = __uint_as_float((uint32_t)w); break; }
      if ((++spins & 1023u) == 0 && global_timer_ns() - t0 > kP2PTimeoutNs) p2p_timeout(slots + lane, epoch, (uint32_t)(w >> 32));
    }
  }
  double acc = 0.0;
  for (int q = 0; q < world; ++q) acc += (double)__shfl_sync(0xffffffffu, v, q);
  if (lane == 0) *out = (float)acc;
}

// ---------------------------------------------------------------- host side
static P2PView make_view(const P2PHandle* h, uint32_t epoch) {
  P2PView v{};
  for (int q = 0; q < kMaxPeers; ++q) v.base[q] = h->base[q];
  v.mc = reinterpret_cast<char*>(h->mc_va);
  v.rank = h->rank;
  v.world = h->world;
  v.epoch = epoch;
  return v;
}

static void launch_pull(const P2PHandle* h, const PullArgs& G, int ctas) {
  if (h->pull_threads == 128) {
    if (h->strong_ld) p2p_pull_kernel<1, 128><<<ctas, 128, 0, h->side>>>(G);
    else p2p_pull_kernel<0, 128><<<ctas, 128, 0, h->side>>>(G);
  } else {
    if (h->strong_ld) p2p_pull_kernel<1, 256><<<ctas, 256, 0, h->side>>>(G);
    else p2p_pull_kernel<0, 256><<<ctas, 256, 0, h->side>>>(G);
  }
}

// ---------------------------------------------------------------- NVLS (NVSwitch multicast) variant of the segment
// VPA_P2P_MODE=nvls.  EXPERIMENTAL: compiles, written ahead of the hardware time to debug it.
// Every rank backs its segment with VMM physical memory (cuMemCreate) and binds it, at offset 0, into ONE multicast object
// created by rank 0 and shared as a POSIX file descriptor (the host passes it between the processes, SCM_RIGHTS).  The
// multicast mapping `mc_va` then aliases all R segments: a multimem.st to mc_va + off lands at `off` in every rank's copy,
// replicated inside the NVSwitch -- one store and 1/(R-1) of the egress of the unicast transports.  Reads and flag polls
// use the local unicast mapping; no rank maps another rank's memory.
struct DriverApi {
  bool ok = false;
  CUresult (*MemCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
  CUresult (*MemRelease)(CUmemGenericAllocationHandle) = nullptr;
  CUresult (*MemAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
  CUresult (*MemAddressFree)(CUdeviceptr, size_t) = nullptr;
  CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
  CUresult (*MemUnmap)(CUdeviceptr, size_t) = nullptr;
  CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
  CUresult (*MemExportToShareableHandle)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long) = nullptr;
  CUresult (*MemImportFromShareableHandle)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType) = nullptr;
  CUresult (*MulticastCreate)(CUmemGenericAllocationHandle*, const CUmulticastObjectProp*) = nullptr;
  CUresult (*MulticastAddDevice)(CUmemGenericAllocationHandle, CUdevice) = nullptr;
  CUresult (*MulticastBindMem)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t, size_t, unsigned long long) = nullptr;
  CUresult (*MulticastUnbind)(CUmemGenericAllocationHandle, CUdevice, size_t, size_t) = nullptr;
  CUresult (*MulticastGetGranularity)(size_t*, const CUmulticastObjectProp*, CUmulticastGranularity_flags) = nullptr;
  CUresult (*DeviceGet)(CUdevice*, int) = nullptr;
  CUresult (*DeviceGetAttribute)(int*, CUdevice_attribute, CUdevice) = nullptr;
};
static DriverApi& driver_api() {
  static DriverApi d;
  static bool tried = false;
  if (tried) return d;
  tried = true;
  bool all = true;
  auto get = [&](const char* name, void** fn) {
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !*fn) {
      all = false;
      cudaGetLastError();
    }
  };
#define VPA_DRV(field, name) get(name, reinterpret_cast<void**>(&d.field))
  VPA_DRV(MemCreate, "cuMemCreate"); VPA_DRV(MemRelease, "cuMemRelease");
  VPA_DRV(MemAddressReserve, "cuMemAddressReserve"); VPA_DRV(MemAddressFree, "cuMemAddressFree");
  VPA_DRV(MemMap, "cuMemMap"); VPA_DRV(MemUnmap, "cuMemUnmap"); VPA_DRV(MemSetAccess, "cuMemSetAccess");
  VPA_DRV(MemExportToShareableHandle, "cuMemExportToShareableHandle");
  VPA_DRV(MemImportFromShareableHandle, "cuMemImportFromShareableHandle");
  VPA_DRV(MulticastCreate, "cuMulticastCreate"); VPA_DRV(MulticastAddDevice, "cuMulticastAddDevice");
  VPA_DRV(MulticastBindMem, "cuMulticastBindMem"); VPA_DRV(MulticastUnbind, "cuMulticastUnbind");
  VPA_DRV(MulticastGetGranularity, "cuMulticastGetGranularity");
  VPA_DRV(DeviceGet, "cuDeviceGet"); VPA_DRV(DeviceGetAttribute, "cuDeviceGetAttribute");
#undef VPA_DRV
  d.ok = all;
  return d;
}
#define VPA_DRV_CALL(expr)                                                                     \
  do {                                                                                         \
    CUresult r__ = (expr);                                                                     \
    if (r__ != CUDA_SUCCESS) return ::vpa::set_error(VPA_E_COMM, "%s failed (CUresult %d)", #expr, (int)r__); \
  } while (0)

static CUmulticastObjectProp nvls_mc_prop(const P2PHandle* h) {
  CUmulticastObjectProp mp{};
  mp.numDevices = (unsigned)h->world;
  mp.size = h->vmm_size;
  mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  mp.flags = 0;
  return mp;
}

// physical memory + local unicast mapping of the segment; size rounded up to the multicast granularity
static int nvls_alloc(P2PHandle* h) {
  DriverApi& d = driver_api();
  if (!d.ok) return set_error(VPA_E_NO_DEVICE, "nvls: CUDA driver entry points (cuMem* / cuMulticast*) unavailable");
  CUdevice dev;
  VPA_DRV_CALL(d.DeviceGet(&dev, h->dev));
  int mc_ok = 0;
  VPA_DRV_CALL(d.DeviceGetAttribute(&mc_ok, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev));
  if (!mc_ok) return set_error(VPA_E_UNSUPPORTED, "nvls: device %d does not support multicast objects", h->dev);
  h->vmm_size = h->L.total;
  CUmulticastObjectProp mp = nvls_mc_prop(h);
  size_t gran = 0;
  VPA_DRV_CALL(d.MulticastGetGranularity(&gran, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED));
  if (gran == 0) gran = (size_t)2 << 20;
  h->vmm_size = align_up(h->L.total, gran);
  CUmemAllocationProp ap{};
  ap.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  ap.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  ap.location.id = h->dev;
  ap.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  VPA_DRV_CALL(d.MemCreate(&h->vmm_mem, h->vmm_size, &ap, 0));
  VPA_DRV_CALL(d.MemAddressReserve(&h->uc_va, h->vmm_size, gran, 0, 0));
  VPA_DRV_CALL(d.MemMap(h->uc_va, h->vmm_size, 0, h->vmm_mem, 0));
  CUmemAccessDesc ad{};
  ad.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  ad.location.id = h->dev;
  ad.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  VPA_DRV_CALL(d.MemSetAccess(h->uc_va, h->vmm_size, &ad, 1));
  h->base[h->rank] = reinterpret_cast<char*>(h->uc_va);
  return 0;
}

static void nvls_free(P2PHandle* h) {
  DriverApi& d = driver_api();
  if (!d.ok) return;
  if (h->mc_va) { d.MemUnmap(h->mc_va, h->vmm_size); d.MemAddressFree(h->mc_va, h->vmm_size); }
  if (h->mc_bound) { CUdevice dev; if (d.DeviceGet(&dev, h->dev) == CUDA_SUCCESS) d.MulticastUnbind(h->vmm_mc, dev, 0, h->vmm_size); }
  if (h->vmm_mc) d.MemRelease(h->vmm_mc);
  if (h->uc_va) { d.MemUnmap(h->uc_va, h->vmm_size); d.MemAddressFree(h->uc_va, h->vmm_size); }
  if (h->vmm_mem) d.MemRelease(h->vmm_mem);
  h->mc_va = h->uc_va = 0;
  h->vmm_mc = h->vmm_mem = 0;
  h->base[h->rank] = nullptr;
}

int p2p_mode(void* handle) { return handle ? static_cast<P2PHandle*>(handle)->pull : -1; }

// rank 0: create the multicast object and export it; the host hands the descriptor to the other ranks' processes
int p2p_nvls_export(void* handle, int* fd_out) {
  P2PHandle* h = static_cast<P2PHandle*>(handle);
  VPA_CHECK_ARG(h && fd_out && h->pull == 4 && h->vmm_mem, "nvls_export: not an NVLS handle");
  DriverApi& d = driver_api();
  CUmulticastObjectProp mp = nvls_mc_prop(h);
  VPA_DRV_CALL(d.MulticastCreate(&h->vmm_mc, &mp));
  int fd = -1;
  VPA_DRV_CALL(d.MemExportToShareableHandle(&fd, h->vmm_mc, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
  *fd_out = fd;
  return 0;
}

// every rank (fd < 0 on the rank that created the object): import, then add this device to the multicast team
int p2p_nvls_attach(void* handle, int fd) {
  P2PHandle* h = static_cast<P2PHandle*>(handle);
  VPA_CHECK_ARG(h && h->pull == 4 && h->vmm_mem, "nvls_attach: not an NVLS handle");
  DriverApi& d = driver_api();
  if (fd >= 0) {
    VPA_CHECK_ARG(h->vmm_mc == 0, "nvls_attach: this rank already holds the multicast object");
    VPA_DRV_CALL(d.MemImportFromShareableHandle(&h->vmm_mc, reinterpret_cast<void*>(static_cast<intptr_t>(fd)),
                                                CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
    close(fd);
  }
  VPA_CHECK_ARG(h->vmm_mc != 0, "nvls_attach: no multicast object (rank 0 exports it, the others pass its descriptor)");
  CUdevice dev;
  VPA_DRV_CALL(d.DeviceGet(&dev, h->dev));
  VPA_DRV_CALL(d.MulticastAddDevice(h->vmm_mc, dev));
  h->mc_added = true;
  return 0;
}

// after EVERY rank has attached (host barrier): bind the local memory and map the multicast view
int p2p_nvls_bind(void* handle) {
  P2PHandle* h = static_cast<P2PHandle*>(handle);
  VPA_CHECK_ARG(h && h->pull == 4 && h->mc_added, "nvls_bind: attach first");
  DriverApi& d = driver_api();
  VPA_DRV_CALL(d.MulticastBindMem(h->vmm_mc, 0, h->vmm_mem, 0, h->vmm_size, 0));
  h->mc_bound = true;
  CUmulticastObjectProp mp = nvls_mc_prop(h);
  size_t gran = 0;
  VPA_DRV_CALL(d.MulticastGetGranularity(&gran, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED));
  VPA_DRV_CALL(d.MemAddressReserve(&h->mc_va, h->vmm_size, gran, 0, 0));
  VPA_DRV_CALL(d.MemMap(h->mc_va, h->vmm_size, 0, h->vmm_mc, 0));
  CUmemAccessDesc ad{};
  ad.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  ad.location.id = h->dev;
  ad.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  VPA_DRV_CALL(d.MemSetAccess(h->mc_va, h->vmm_size, &ad, 1));
  return 0;
}

// operands through the multicast mapping: every CTA takes whole 256-row chunks of THIS rank's block (x2 operands first),
// one multimem.st per 16 bytes reaches all ranks, then one system fence and the chunk's flag -- again one store for all.
__global__ void __launch_bounds__(256) p2p_mc_push_kernel(const PullArgs A) {
  const int me = A.v.rank;
  char* mine = A.v.base[me];
  const int total = 2 * A.cpr;
  for (int item = blockIdx.x; item < total; item += gridDim.x) {
    const int m = item / A.cpr, c = item - m * A.cpr;
    const int64_t row0 = (int64_t)c * kPushRows;
    const int rows = (int)min((int64_t)kPushRows, A.b - row0);
    const int n16 = rows * (A.row_bytes / 16);
    const size_t off = A.off_mat[m] + ((size_t)me * A.b + row0) * A.row_bytes;
    const uint4* from = reinterpret_cast<const uint4*>(mine + off);
    uint4* to = reinterpret_cast<uint4*>(A.v.mc + off);
    for (int i = threadIdx.x; i < n16; i += 256 * kPullUnroll) {
      uint4 val[kPullUnroll];
#pragma unroll
      for (int u = 0; u < kPullUnroll; ++u) {
        const int idx = i + u * 256;
        if (idx < n16) val[u] = __ldg(from + idx);
      }
#pragma unroll
      for (int u = 0; u < kPullUnroll; ++u) {
        const int idx = i + u * 256;
        if (idx < n16) mc_st_v4(to + idx, val[u]);
      }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0)
      mc_st_release_sys_u32(reinterpret_cast<uint32_t*>(A.v.mc + A.off_flags[m]) + me * A.cpr + c, A.v.epoch);
  }
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
    if (h->pull == 4) nvls_free(h);
    else if (h->base[rank]) cudaFree(h->base[rank]);
    delete h;
    return rc;
  };
  cudaError_t e;
  if ((e = cudaGetDevice(&h->dev)) != cudaSuccess) return fail(e, "cudaGetDevice");
  const char* mode_env = getenv("VPA_P2P_MODE");
  void* p = nullptr;
  if (mode_env && strcmp(mode_env, "nvls") == 0) {      // VMM memory bound into a multicast object; no IPC handle
    h->pull = 4;
    cudaFree(nullptr);                                  // make sure the primary context is current for the driver calls
    if (int rc = nvls_alloc(h)) {
      nvls_free(h);
      delete h;
      return rc;
    }
    p = h->base[rank];
    memset(ipc_handle64, 0, 64);
  } else {
    if ((e = cudaMalloc(&p, h->L.total)) != cudaSuccess) return fail(e, "cudaMalloc");
    h->base[rank] = static_cast<char*>(p);
    cudaIpcMemHandle_t ih;
    if ((e = cudaIpcGetMemHandle(&ih, p)) != cudaSuccess) return fail(e, "cudaIpcGetMemHandle");
    memcpy(ipc_handle64, &ih, 64);
  }
  if ((e = cudaMemset(p, 0, h->L.mat[0][0])) != cudaSuccess) return fail(e, "cudaMemset");      // flags, slots, counters
  int lo = 0, hi = 0;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);
  if ((e = cudaStreamCreateWithPriority(&h->side, cudaStreamNonBlocking, hi)) != cudaSuccess) return fail(e, "cudaStreamCreate");
  if ((e = cudaEventCreateWithFlags(&h->fork, cudaEventDisableTiming)) != cudaSuccess) return fail(e, "cudaEventCreate");
  if ((e = cudaEventCreateWithFlags(&h->join, cudaEventDisableTiming)) != cudaSuccess) return fail(e, "cudaEventCreate");
  if (const char* s = getenv("VPA_P2P_PUSH_GROUPS")) { const int v = atoi(s); if (v >= 1 && v <= 32) h->push_groups = v; }
  if (const char* s = getenv("VPA_P2P_PUSH_CTAS")) { const int v = atoi(s); if (v >= 1 && v <= 32) h->push_ctas = v; }
  if (const char* s = getenv("VPA_P2P_MODE")) h->pull = strcmp(s, "push") == 0 ? 0 : (strcmp(s, "ce") == 0 ? 2 : (strcmp(s, "stream") == 0 ? 3 : (strcmp(s, "nvls") == 0 ? 4 : 1)));
  if (const char* s = getenv("VPA_P2P_MC_CTAS")) { const int v = atoi(s); if (v >= 1 && v <= 148) h->mc_ctas = v; }
  if (const char* s = getenv("VPA_P2P_STREAM_CTAS")) { const int v = atoi(s); if (v >= 1 && v <= 1024) h->stream_ctas = v; }
  if (const char* s = getenv("VPA_P2P_PULL_SLICES")) { const int v = atoi(s); if (v >= 1 && v <= 16) h->pull_slices = v; }
  if (const char* s = getenv("VPA_P2P_PULL_THREADS")) { const int v = atoi(s); if (v == 128 || v == 256) h->pull_threads = v; }
  if (const char* s = getenv("VPA_P2P_PULL_LD")) h->strong_ld = strcmp(s, "weak") != 0;
  if (const char* s = getenv("VPA_P2P_PLAN")) h->serial = strcmp(s, "serial") == 0;
  if (const char* s = getenv("VPA_P2P_PULL_CTAS_ALONE")) { const int v = atoi(s); if (v >= 1 && v <= 1024) h->pull_ctas_alone = v; }
  if ((e = cudaEventCreateWithFlags(&h->t_done, cudaEventDisableTiming)) != cudaSuccess) return fail(e, "cudaEventCreate");
  if ((e = cudaEventCreateWithFlags(&h->fwd_done, cudaEventDisableTiming)) != cudaSuccess) return fail(e, "cudaEventCreate");
  if (const char* s = getenv("VPA_P2P_CE_STREAMS")) { const int v = atoi(s); if (v == 1 || v == 2) h->ce_streams = v; }
  if (const char* s = getenv("VPA_P2P_CE_KB")) { const int v = atoi(s); if (v >= 256) h->ce_bytes = (size_t)v << 10; }
  if ((e = cudaStreamCreateWithPriority(&h->side2, cudaStreamNonBlocking, hi)) != cudaSuccess) return fail(e, "cudaStreamCreate");
  if ((e = cudaEventCreateWithFlags(&h->ready_ev, cudaEventDisableTiming)) != cudaSuccess) return fail(e, "cudaEventCreate");
  if ((e = cudaEventCreateWithFlags(&h->join2, cudaEventDisableTiming)) != cudaSuccess) return fail(e, "cudaEventCreate");
  if (const char* s = getenv("VPA_P2P_PULL_CTAS")) { const int v = atoi(s); if (v >= 1 && v <= 1024) h->pull_ctas = v; }
  if (const char* s = getenv("VPA_P2P_PUSH_ROUNDS")) { const int v = atoi(s); if (v >= 1 && v <= 1024) h->push_rounds = v; }
  if ((e = cudaDeviceSynchronize()) != cudaSuccess) return fail(e, "cudaDeviceSynchronize");
  {
    // The pull CTAs must be able to become resident beside the single-pass forward (see p2p_pull_kernel): check the
    // register budget of THIS build instead of trusting a comment.
    cudaFuncAttributes fa;
    const int fwd_regs = pair_fwd1_regs_per_cta();
    if (fwd_regs > 0 && cudaFuncGetAttributes(&fa, p2p_pull_kernel<1, 256>) == cudaSuccess) {
      const int pull_regs = (fa.numRegs + 7) / 8 * 8 * 256;
      if (fwd_regs + pull_regs > 65536) {
        const int rc = set_error(VPA_E_UNSUPPORTED, "p2p: pull kernel (%d regs/CTA) cannot co-reside with the forward sweep (%d regs/CTA)",
                                 pull_regs, fwd_regs);
        if (h->pull == 4) nvls_free(h);
        else cudaFree(h->base[rank]);
        delete h;
        return rc;
      }
    } else {
      cudaGetLastError();
    }
  }
  *out = h;
  return 0;
}

int p2p_connect(void* handle, const void* all_handles) {
  P2PHandle* h = static_cast<P2PHandle*>(handle);
  VPA_CHECK_ARG(h && (all_handles || h->pull == 4), "p2p_connect: bad argument");
  if (h->connected) return 0;
  if (h->pull == 4) {      // NVLS: nothing to map -- every exchange goes through the multicast view
    VPA_CHECK_ARG(h->mc_va != 0, "p2p_connect: NVLS segment is not bound yet (export / attach / bind first)");
    h->connected = true;
    return 0;
  }
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
  if (h->pull == 4) nvls_free(h);
  else if (h->base[h->rank]) cudaFree(h->base[h->rank]);
  if (h->side) cudaStreamDestroy(h->side);
  if (h->side2) cudaStreamDestroy(h->side2);
  if (h->ready_ev) cudaEventDestroy(h->ready_ev);
  if (h->t_done) cudaEventDestroy(h->t_done);
  if (h->fwd_done) cudaEventDestroy(h->fwd_done);
  if (h->join2) cudaEventDestroy(h->join2);
  if (h->fork) cudaEventDestroy(h->fork);
  if (h->join) cudaEventDestroy(h->join);
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
  s.pack_counter = reinterpret_cast<uint32_t*>(base + L.counters) + 2 * L.cpr;
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
  return s;
}

// operands of this step -> all peers, on the side stream (forked after the normalise kernel on `st`)
int p2p_push_operands(void* handle, uint32_t epoch, cudaStream_t st) {
  P2PHandle* h = static_cast<P2PHandle*>(handle);
  const SegLayout& L = h->L;
  const int p = (int)(epoch & 1u);
  PushArgs A{};
  A.v = make_view(h, epoch);
  A.off_mat[0] = L.mat[p][0]; A.off_mat[1] = L.mat[p][1];
  A.off_flags[0] = L.flags[0]; A.off_flags[1] = L.flags[1];
  A.off_counters = L.counters;
  A.b = h->b;
  A.row_bytes = h->D * (h->precision == VPA_PREC_BF16_TC ? 2 : 4);
  A.cpr = L.cpr;
  A.ctas_per_group = h->push_ctas;
  // fences (one NVLink round trip each) per group and matrix are bounded by push_rounds: large blocks go in multi-chunk batches
  A.batch = (L.cpr + h->push_groups * h->push_rounds - 1) / (h->push_groups * h->push_rounds);
  if (A.batch < 1) A.batch = 1;
  const int nbatch = (L.cpr + A.batch - 1) / A.batch;
  A.groups = h->push_groups < 2 * nbatch ? h->push_groups : 2 * nbatch;
  VPA_CUDA(cudaEventRecord(h->fork, st));
  VPA_CUDA(cudaStreamWaitEvent(h->side, h->fork, 0));
  if (h->pull == 2) {
    // copy engines: signal / wait readiness once, then one peer-to-local copy per (matrix, peer, <= ce_bytes piece), each
    // followed by a one-warp kernel that flips the arrival flags of the chunks it delivered.  x2 operands first.
    char* mine = h->base[h->rank];
    p2p_signal_ready_kernel<<<1, 32, 0, h->side>>>(A.v, L.ready);
    p2p_wait_ready_kernel<<<1, 32, 0, h->side>>>(reinterpret_cast<const uint32_t*>(mine + L.ready), h->world, h->rank, epoch);
    VPA_LAUNCH_CHECK("p2p ready kernels");
    cudaStream_t ss[2] = {h->side, h->ce_streams == 2 ? h->side2 : h->side};
    if (h->ce_streams == 2) {
      VPA_CUDA(cudaEventRecord(h->ready_ev, h->side));
      VPA_CUDA(cudaStreamWaitEvent(h->side2, h->ready_ev, 0));
    }
    prof_begin(PROF_PUSH, h->side);
    const size_t chunk_bytes = (size_t)kPushRows * A.row_bytes;
    int per = (int)(h->ce_bytes / chunk_bytes);
    if (per < 1) per = 1;
    if (per > 32) per = 32;
    int n = 0;
    for (int m = 0; m < 2; ++m)
      for (int c0 = 0; c0 < L.cpr; c0 += per)
        for (int q = 1; q < h->world; ++q, ++n) {
          const int src = (h->rank + q) % h->world;
          const int c1 = c0 + per < L.cpr ? c0 + per : L.cpr;
          const int64_t row0 = (int64_t)c0 * kPushRows;
          const int64_t rows = ((int64_t)c1 * kPushRows < h->b ? (int64_t)c1 * kPushRows : h->b) - row0;
          const size_t off = A.off_mat[m] + ((size_t)src * h->b + row0) * A.row_bytes;
          cudaStream_t cs = ss[n & 1];
          VPA_CUDA(cudaMemcpyAsync(mine + off, h->base[src] + off, (size_t)rows * A.row_bytes, cudaMemcpyDeviceToDevice, cs));
          p2p_set_flags_kernel<<<1, 32, 0, cs>>>(reinterpret_cast<uint32_t*>(mine + L.flags[m]) + src * L.cpr + c0, c1 - c0, epoch);
        }
    VPA_LAUNCH_CHECK("p2p_set_flags_kernel");
    if (h->ce_streams == 2) {
      VPA_CUDA(cudaEventRecord(h->join2, h->side2));
      VPA_CUDA(cudaStreamWaitEvent(h->side, h->join2, 0));
    }
    prof_end(PROF_PUSH, h->side);
  } else {
  prof_begin(PROF_PUSH, h->side);
  if (h->pull == 4) {
    PullArgs G{};
    G.v = A.v;
    G.off_mat[0] = A.off_mat[0]; G.off_mat[1] = A.off_mat[1];
    G.off_flags[0] = A.off_flags[0]; G.off_flags[1] = A.off_flags[1];
    G.off_ready = L.ready;
    G.b = h->b; G.row_bytes = A.row_bytes; G.cpr = L.cpr;
    G.m0 = 0; G.m1 = 2;
    const int items = 2 * L.cpr;
    p2p_mc_push_kernel<<<items < h->mc_ctas ? items : h->mc_ctas, 256, 0, h->side>>>(G);
  } else if (h->pull == 3) {
    PullArgs G{};
    G.v = A.v;
    G.off_mat[0] = A.off_mat[0]; G.off_mat[1] = A.off_mat[1];
    G.off_flags[0] = A.off_flags[0]; G.off_flags[1] = A.off_flags[1];
    G.off_ready = L.ready;
    G.b = h->b; G.row_bytes = A.row_bytes; G.cpr = L.cpr;
    G.m0 = 0; G.m1 = 2;
    const int items = 2 * (h->world - 1) * L.cpr;
    p2p_stream_push_kernel<<<items < h->stream_ctas ? items : h->stream_ctas, 128, 0, h->side>>>(G);
  } else if (h->pull) {
    PullArgs G{};
    G.v = A.v;
    G.off_mat[0] = A.off_mat[0]; G.off_mat[1] = A.off_mat[1];
    G.off_flags[0] = A.off_flags[0]; G.off_flags[1] = A.off_flags[1];
    G.off_ready = L.ready;
    G.b = h->b; G.row_bytes = A.row_bytes; G.cpr = L.cpr;
    const bool serial = h->serial && h->pull == 1;
    G.m0 = 0; G.m1 = serial ? 1 : 2;
    G.slices = h->pull_slices; G.off_pull_counters = L.pull_counters;
    const int items = (G.m1 - G.m0) * (h->world - 1) * L.cpr * G.slices;
    const int ctas = serial ? h->pull_ctas_alone : h->pull_ctas;
    launch_pull(h, G, items < ctas ? items : ctas);
    if (serial) {      // the forward starts when all x2 operands are here: the transfer has the fabric and the L2s to itself
      VPA_CUDA(cudaEventRecord(h->t_done, h->side));
      VPA_CUDA(cudaStreamWaitEvent(st, h->t_done, 0));
    }
  } else {
    p2p_push_kernel<<<A.groups * A.ctas_per_group, 256, 0, h->side>>>(A);
  }
  prof_end(PROF_PUSH, h->side);
  VPA_LAUNCH_CHECK("p2p_push / p2p_pull kernel");
  }
  VPA_CUDA(cudaEventRecord(h->join, h->side));
  h->join_pending = true;
  return 0;
}

// serial plan: the x1 operands (read by the backward only) move once the forward sweep on `st` has finished
int p2p_pull_rest(void* handle, uint32_t epoch, cudaStream_t st) {
  P2PHandle* h = static_cast<P2PHandle*>(handle);
  if (!(h->serial && h->pull == 1)) return 0;
  const SegLayout& L = h->L;
  const int p = (int)(epoch & 1u);
  PullArgs G{};
  G.v = make_view(h, epoch);
  G.off_mat[0] = L.mat[p][0]; G.off_mat[1] = L.mat[p][1];
  G.off_flags[0] = L.flags[0]; G.off_flags[1] = L.flags[1];
  G.off_ready = L.ready;
  G.b = h->b; G.row_bytes = h->D * (h->precision == VPA_PREC_BF16_TC ? 2 : 4); G.cpr = L.cpr;
  G.m0 = 1; G.m1 = 2;
  G.slices = h->pull_slices; G.off_pull_counters = L.pull_counters;
  VPA_CUDA(cudaEventRecord(h->fwd_done, st));
  VPA_CUDA(cudaStreamWaitEvent(h->side, h->fwd_done, 0));
  const int items = (h->world - 1) * L.cpr * G.slices;
  launch_pull(h, G, items < h->pull_ctas_alone ? items : h->pull_ctas_alone);
  VPA_LAUNCH_CHECK("p2p_pull_kernel");
  VPA_CUDA(cudaEventRecord(h->join, h->side));
  h->join_pending = true;
  return 0;
}

// the main stream must not overwrite this rank's operand block while an earlier push still reads it
int p2p_join_push(void* handle, cudaStream_t st) {
  P2PHandle* h = static_cast<P2PHandle*>(handle);
  if (h->join_pending) VPA_CUDA(cudaStreamWaitEvent(st, h->join, 0));
  h->join_pending = false;
  return 0;
}

int p2p_wait_operands(void* handle, uint32_t epoch, const float* gate_scale, float scale_cap, cudaStream_t st) {
  P2PHandle* h = static_cast<P2PHandle*>(handle);
  const SegLayout& L = h->L;
  char* base = h->base[h->rank];
  p2p_wait_all_kernel<<<1, 256, 0, st>>>(reinterpret_cast<const uint32_t*>(base + L.flags[0]),
                                         reinterpret_cast<const uint32_t*>(base + L.flags[1]), h->world, L.cpr, h->rank, epoch,
                                         gate_scale, scale_cap, pair_fast_s2_limit());
  VPA_LAUNCH_CHECK("p2p_wait_all_kernel");
  return 0;
}

int p2p_dls_sum(const P2PStep& s, float* dlogit_scale, cudaStream_t st) {
  p2p_dls_sum_kernel<<<1, 32, 0, st>>>(s.dls_slots, s.view.world, s.view.epoch, dlogit_scale);
  VPA_LAUNCH_CHECK("p2p_dls_sum_kernel");
  return 0;
}

}  // namespace vpa
