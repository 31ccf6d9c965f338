// Peer-memory transport of the row-sharded step: the gathered operand matrices, the per-rank statistics messages and the
// d logit_scale partials live in a SYMMETRIC segment (same layout on every rank, cudaMalloc + CUDA IPC), and every
// exchange is done by kernels reading / writing the peers' segments over NVLink, published with system-scope epoch flags:
//   * operands: RELAY CTAs (p2p.cuh) -- the first CTAs of the single-pass forward's own grid -- pull the peers' normalised
//     rows with TMA bulk copies through a shared-memory ring (x2 operands first, 256-row chunks, chunk-major over the
//     peers) and raise one local arrival flag per chunk.  The sweep CTAs of the same kernel poll the flag of every tile
//     before loading it, so the contraction starts on the local block and consumes remote rows as they land: ONE kernel
//     is the all-gather and the GEMM.  The x1 operands (read by the backward only) follow in the same relay CTAs; when the
//     kernel has finished everything has landed -- no wait kernel, no side stream.  The only remote store of the pull is
//     one "my rows are complete" flag per peer and step.  VPA_P2P_MODE=nvls: the segment is NVSwitch multicast memory and
//     every rank stores its own rows once (multimem.st) instead.
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
  int nvls = 0;             // 1: segment allocated with the VMM API and bound into an NVSwitch multicast object (see below)
  int relay_ctas = 20;      // relay CTAs in front of the forward / backward grids (whole CTA pairs; VPA_P2P_RELAY_CTAS)
  uint32_t a_pending = 0;   // step whose x1 operands are still to be gathered (by its backward, or by the next forward)
  CUmemGenericAllocationHandle vmm_mem = 0, vmm_mc = 0;
  CUdeviceptr uc_va = 0, mc_va = 0;
  size_t vmm_size = 0;
  bool mc_added = false, mc_bound = false;
};

// ---------------------------------------------------------------- kernels
// The relay as a kernel of its own: shapes the CTA-pair sweeps do not cover (fp32 mode, D not in {256, 512}) gather their
// operands before the forward starts.  Same device code as the relay CTAs of the fused forward kernel.
__global__ void __launch_bounds__(256) p2p_relay_kernel(const RelayArgs A) {
  extern __shared__ uint8_t relay_smem[];
  pdl_trigger();
  pdl_wait();
  if (blockIdx.x == 0) relay_signal_ready(A);
  if (A.multicast) {
    relay_multicast(A, A.m1, blockIdx.x);
    return;
  }
  const uint32_t base = relay_smem_u32(relay_smem);
  relay_pull(A, A.m1, blockIdx.x, relay_smem + (((base + 1023u) & ~1023u) - base));
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

// ---------------------------------------------------------------- NVLS (NVSwitch multicast) variant of the segment
// VPA_P2P_MODE=nvls at vpa_p2p_create time.
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

int p2p_mode(void* handle) { return handle ? (static_cast<P2PHandle*>(handle)->nvls ? 4 : 1) : -1; }

// rank 0: create the multicast object and export it; the host hands the descriptor to the other ranks' processes
int p2p_nvls_export(void* handle, int* fd_out) {
  P2PHandle* h = static_cast<P2PHandle*>(handle);
  VPA_CHECK_ARG(h && fd_out && h->nvls && h->vmm_mem, "nvls_export: not an NVLS handle");
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
  VPA_CHECK_ARG(h && h->nvls && h->vmm_mem, "nvls_attach: not an NVLS handle");
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
  VPA_CHECK_ARG(h && h->nvls && h->mc_added, "nvls_bind: attach first");
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
    if (h->nvls) nvls_free(h);
    else if (h->base[rank]) cudaFree(h->base[rank]);
    delete h;
    return rc;
  };
  cudaError_t e;
  if ((e = cudaGetDevice(&h->dev)) != cudaSuccess) return fail(e, "cudaGetDevice");
  const char* mode_env = getenv("VPA_P2P_MODE");
  void* p = nullptr;
  if (mode_env && strcmp(mode_env, "nvls") == 0) {      // VMM memory bound into a multicast object; no IPC handle
    h->nvls = 1;
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
  h->relay_ctas = relay_ctas_default();      // whole CTA pairs: the forward kernel is launched in clusters of two
  if ((e = cudaDeviceSynchronize()) != cudaSuccess) return fail(e, "cudaDeviceSynchronize");
  *out = h;
  return 0;
}

int p2p_connect(void* handle, const void* all_handles) {
  P2PHandle* h = static_cast<P2PHandle*>(handle);
  VPA_CHECK_ARG(h && (all_handles || h->nvls), "p2p_connect: bad argument");
  if (h->connected) return 0;
  if (h->nvls) {      // NVLS: nothing to map -- every exchange goes through the multicast view
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
  if (h->nvls) nvls_free(h);
  else if (h->base[h->rank]) cudaFree(h->base[h->rank]);
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
  A.multicast = h->nvls;
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
  const int items = (2 - m0) * (A.multicast ? A.cpr : A.cpr * (h->world - 1));
  A.n_ctas = items < 32 ? items : 32;
  static SmemAttrCache attr_cache;
  if (int e = ensure_dynamic_smem(attr_cache, p2p_relay_kernel, (int)kRelaySmemBytes + 1024)) return e;
  prof_begin(PROF_PUSH, st);
  VPA_CUDA(launch_kernel(p2p_relay_kernel, dim3(A.n_ctas), dim3(256), A.multicast ? 0 : kRelaySmemBytes + 1024, st, A));
  prof_end(PROF_PUSH, st);
  VPA_LAUNCH_CHECK("p2p_relay_kernel");
  return 0;
}

}  // namespace vpa
