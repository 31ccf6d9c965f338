// InfoNCE sweeps on the 5th-generation tensor cores (sm_100a): tcgen05.mma + TMEM accumulators,
// operands staged by TMA (cp.async.bulk.tensor, 128B swizzle), warp-specialised roles, mbarrier
// pipelines.  bf16 operands, fp32 accumulate.  The B x B logits exist only as 128 x 128 fp32 tiles
// in tensor memory.
//
// One CTA ("unit") owns a 128-row block X_i of one modality, keeps it RESIDENT in shared memory
// (D/64 boxes of 128 rows x 64 bf16, 16 KB each) and sweeps a chunk of 128-row tiles Y_j of the
// other modality through a TMA ring:
//   S_ij = X_i . Y_j^T                (32 x tcgen05.mma M128 N128 K16 per tile at D = 512)
// forward  (MODE_FWD): epilogue warps read S from TMEM (tcgen05.ld) and keep an online
//          (max, sum-of-exp2) per row -> per-chunk partial statistics;
// backward (MODE_BWD): epilogue warps recompute G_ij = (exp(S-lse_x) + exp(S-lse_y) - 2 delta)/B in
//          registers, write it as bf16 into shared memory in the canonical K-major SW128 layout and
//          the MMA warp contracts it with the SAME Y_j boxes (now read MN-major):
//   dX_i[:, half] += G_ij . Y_j[:, half]     (TMEM-resident fp32 accumulator, D/2 <= 256 columns)
//          TMEM holds 512 fp32 columns per lane: 2 x 128 (double-buffered S) + 256 (dX half), which
//          is why D = 512 is processed as two independent halves (plan.halves).
// Two problems (A against T, T against A) run in one launch; see common.cuh::SweepArgs.
//
// Reference semantics: loss_head.py:277-283 (logits + 2 x cross entropy) and its autograd.
#include <cuda.h>

#include <cstdio>

#include "common.cuh"

namespace vpa {
namespace tc {

constexpr int kBM = 128;                 // X rows per CTA  (UMMA M)
constexpr int kBN = 128;                 // Y rows per tile (UMMA N of the S product)
constexpr int kBoxK = 64;                // bf16 elements per 128-byte swizzle row
constexpr int kBoxBytes = kBM * kBoxK * 2;   // 16 KB: one [128 rows][64 elems] TMA box
constexpr int kMaxStages = 12;
constexpr int kThreads = 192;            // warp 0: TMA, warp 1: MMA + TMEM alloc, warps 2..5: epilogue
constexpr int kEpiThreads = 128;
constexpr uint32_t kSmemLimit = 232448;  // 227 KB opt-in limit per CTA

enum { MODE_FWD = 0, MODE_BWD = 1 };

struct Problem {
  int n_x, n_y;             // valid rows of X (local) and Y (global)
  int diag_offset;          // global Y index of the diagonal partner of local X row 0
  const float* lse_x;       // BWD: + diag_offset indexes the local rows' lse
  const float* lse_y;       // BWD: indexed by Y row
  float* out;               // FWD: float2[n_chunks][n_x]; BWD: float[n_chunks][n_x][D]
  float* dscale;            // BWD problem 0: one partial per unit; nullptr otherwise
};

struct Params {
  Problem p[2];
  int units_per_problem;    // CTAs per problem = n_igroups * cluster * halves * n_chunks
  int n_iblk, halves, n_chunks, tiles_per_chunk, n_tiles;
  int cluster, n_igroups;   // CTAs per cluster (1, 2 or 4) sweeping the SAME Y tiles; row-block groups
  int D, kboxes;            // kboxes = D / 64
  int stages;
  const float* logit_scale; // FWD
  float scale_cap;          // FWD
  const float* scale;       // BWD
  float inv_B, ln_B;
};

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded wait: a protocol bug must abort the kernel (trap) instead of hanging the GPU.  try_wait suspends
// the thread in hardware for a bounded time, so the loop body runs rarely; the bound is an iteration count
// (no clock reads or integer division on the critical path of the single-thread issue loops).
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done != 0;
}
__device__ __noinline__ void mbar_timeout(uint32_t bar, uint32_t parity) {
  printf("vipant_b200: mbarrier wait timeout (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
  asm volatile("trap;");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try(bar, parity)) return;
  uint32_t spins = 0;
  while (!mbar_try(bar, parity))
    if (++spins > (1u << 24)) mbar_timeout(bar, parity);
}
// true for exactly one lane of a converged warp (the lane that issues the asynchronous instructions)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, %1;\n\t"
      "@px mov.s32 %0, 1;\n\t}"
      : "+r"(pred)
      : "r"(0xffffffffu));
  return pred != 0;
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
// One slice of a box, written to the same shared-memory offset of every CTA in `mask`; each destination
// CTA's mbarrier (same offset) receives the complete_tx for the bytes that landed in ITS shared memory.
__device__ __forceinline__ void tma_load_2d_mcast(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar,
                                                  uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%2, %3}], [%4], %5;"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// D[tmem] (+)= A[smem] . B[smem], issued by ONE thread for the whole CTA.
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrive when every previously issued MMA of this thread has completed.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// Same, arriving on the barrier at this offset in every CTA of `mask` (frees a multicast ring stage).
__device__ __forceinline__ void umma_commit_mcast(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread t of the warp gets lane (base + t).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): 128B swizzle, 8-row groups 1024 B apart.
//   bits [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout=2 (SW128)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor), kind::f16: bf16 x bf16 -> fp32.
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) /*D=f32*/ | (1u << 7) /*A=bf16*/ | (1u << 10) /*B=bf16*/ | ((uint32_t)a_mn_major << 15) |
         ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---------------------------------------------------------------- shared-memory carve-up
struct SmemLayout {
  uint32_t x;        // kboxes resident X boxes
  uint32_t g;        // BWD: 2 boxes [128 i][64 j] bf16 (K-major A operand of the dX product)
  uint32_t ring;     // stages x 16 KB
  uint32_t cl;       // BWD: float[2][128] column lse of the current tile (base-2, + log2 B)
  uint32_t red;      // float[8]
  uint32_t bars;     // mbarriers
  uint32_t tmem_slot;
  uint32_t total;
};
__host__ __device__ inline SmemLayout smem_layout(int mode, int kboxes, int stages) {
  SmemLayout L;
  uint32_t o = 0;
  L.x = o; o += kboxes * kBoxBytes;
  L.g = o; o += (mode == MODE_BWD) ? 2 * kBoxBytes : 0;
  L.ring = o; o += stages * kBoxBytes;
  L.cl = o; o += 2 * 128 * 4;
  L.red = o; o += 64;
  L.bars = o; o += (2 * kMaxStages + 16) * 8;
  L.tmem_slot = o; o += 16;
  L.total = o + 1024;   // slack for the 1024-byte alignment of the dynamic smem base
  return L;
}
// barrier indices
enum { B_XFULL = 0, B_TFULL0, B_TFULL1, B_TEMPTY0, B_TEMPTY1, B_GFULL, B_GEMPTY, B_DXFULL, B_RING };

template <int MODE>
__global__ void __launch_bounds__(kThreads, 1)
sweep_kernel(const __grid_constant__ CUtensorMap mx0, const __grid_constant__ CUtensorMap my0,
             const __grid_constant__ CUtensorMap mx1, const __grid_constant__ CUtensorMap my1,
             const Params P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sptr = smem_raw + (sbase - smem_u32(smem_raw));
  const SmemLayout L = smem_layout(MODE, P.kboxes, P.stages);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- unit decode.  The `cluster` CTAs of a cluster own consecutive row blocks and sweep the SAME Y tiles
  // (same problem / chunk / half): every Y box is fetched from L2 once per cluster (each CTA loads 1/cluster
  // of it and multicasts), which divides the L2 -> SM traffic -- the measured bound of the v1 kernel -- by `cluster`.
  int u = blockIdx.x;
  const int prob = u >= P.units_per_problem;
  u -= prob * P.units_per_problem;           // CTA index within the problem (also the dscale slot)
  const int crank = P.cluster > 1 ? (int)cluster_ctarank() : 0;
  const uint16_t cmask = (uint16_t)((1u << P.cluster) - 1u);
  const int cl = u / P.cluster;              // cluster index within the problem
  const int per_chunk = P.n_igroups * P.halves;
  const int chunk = cl / per_chunk;
  const int rem = cl - chunk * per_chunk;
  const int igroup = rem / P.halves;
  const int half = rem - igroup * P.halves;
  const int iblk = igroup * P.cluster + crank;   // may be >= n_iblk in the last group: rows masked, TMA zero-fills
  const Problem& pb = P.p[prob];
  const CUtensorMap* mapx = prob ? &mx1 : &mx0;
  const CUtensorMap* mapy = prob ? &my1 : &my0;
  const int tile0 = chunk * P.tiles_per_chunk;
  const int nt = min(P.tiles_per_chunk, P.n_tiles - tile0);
  const int hboxes = P.kboxes / P.halves;          // boxes (64 columns each) of this CTA's dX half
  const uint32_t bar0 = sbase + L.bars;
  auto bar = [&](int i) { return bar0 + 8u * i; };
  auto ring_full = [&](int s) { return bar(B_RING + s); };
  auto ring_empty = [&](int s) { return bar(B_RING + kMaxStages + s); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sptr + L.tmem_slot);

  if (threadIdx.x == 0) {
    mbar_init(bar(B_XFULL), 1);
    mbar_init(bar(B_TFULL0), 1);
    mbar_init(bar(B_TFULL1), 1);
    mbar_init(bar(B_TEMPTY0), kEpiThreads);
    mbar_init(bar(B_TEMPTY1), kEpiThreads);
    mbar_init(bar(B_GFULL), kEpiThreads);
    mbar_init(bar(B_GEMPTY), 1);
    mbar_init(bar(B_DXFULL), 1);
    for (int s = 0; s < P.stages; ++s) { mbar_init(ring_full(s), 1); mbar_init(ring_empty(s), P.cluster); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    prefetch_tmap(mapx);
    prefetch_tmap(mapy);
  }
  constexpr uint32_t kTmemCols = (MODE == MODE_BWD) ? 512u : 256u;
  if (warp == 1) tmem_alloc(sbase + L.tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  if (P.cluster > 1) cluster_sync_all();     // peers' barriers are initialised before any remote arrive / multicast
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =========================== TMA producer: converged warp, one elected lane issues ===========================
    const bool leader = elect_one();
    if (leader) {
      mbar_arrive_expect_tx(bar(B_XFULL), P.kboxes * kBoxBytes);
      for (int kb = 0; kb < P.kboxes; ++kb)
        tma_load_2d(sbase + L.x + kb * kBoxBytes, mapx, kb * kBoxK, iblk * kBM, bar(B_XFULL));
    }
    const int slice_rows = kBN / P.cluster;            // this CTA's share of every Y box
    const int row_shift = crank * slice_rows;
    const uint32_t ring0 = sbase + L.ring + (uint32_t)row_shift * 128u;
    int stage = 0;
    uint32_t phase = 0;
    auto push = [&](int c0, int c1) {
      mbar_wait(ring_empty(stage), phase ^ 1);       // every CTA of the cluster has drained this stage
      if (leader) {
        mbar_arrive_expect_tx(ring_full(stage), kBoxBytes);
        if (P.cluster > 1)
          tma_load_2d_mcast(ring0 + stage * kBoxBytes, mapy, c0, c1 + row_shift, ring_full(stage), cmask);
        else
          tma_load_2d(ring0 + stage * kBoxBytes, mapy, c0, c1, ring_full(stage));
      }
      if (++stage == P.stages) { stage = 0; phase ^= 1; }
    };
    const int d0 = half * hboxes * kBoxK;
    for (int j = 0; j < nt; ++j) {
      const int r = (tile0 + j) * kBN;
      for (int kb = 0; kb < P.kboxes; ++kb) push(kb * kBoxK, r);                               // S(j)
      if (MODE == MODE_BWD && j >= 1)
        for (int hb = 0; hb < hboxes; ++hb) push(d0 + hb * kBoxK, r - kBN);                     // dX(j-1)
    }
    if (MODE == MODE_BWD)
      for (int hb = 0; hb < hboxes; ++hb) push(d0 + hb * kBoxK, (tile0 + nt - 1) * kBN);
  } else if (warp == 1) {
    // =========================== MMA issuer: converged warp, one elected lane issues ===========================
    const bool leader = elect_one();
    constexpr uint32_t idesc_s = make_idesc(kBM, kBN, 0, 0);      // S: both operands K-major
    constexpr uint32_t idesc_dx = make_idesc(kBM, 64, 0, 1);      // dX: A = G K-major, B = Y MN-major, N = 64
    // descriptor templates; only the 14-bit start-address field (units of 16 B) changes per instruction
    const uint64_t dk = make_desc(sbase, 16, 1024);               // K-major SW128 operand at smem base
    const uint64_t dmn = make_desc(sbase, kBoxBytes, 1024);       // MN-major SW128 operand at smem base
    const uint32_t ring_u = (L.ring) >> 4, x_u = (L.x) >> 4, g_u = (L.g) >> 4;
    constexpr uint32_t box_u = kBoxBytes >> 4;
    int stage = 0;
    uint32_t phase = 0;
    auto release = [&](int st) {        // the stage may be refilled once EVERY CTA of the cluster has consumed it
      if (P.cluster > 1) umma_commit_mcast(ring_empty(st), cmask);
      else umma_commit(ring_empty(st));
    };
    mbar_wait(bar(B_XFULL), 0);
    tc_fence_after();
    auto issue_dx = [&](int t) {
      mbar_wait(bar(B_GFULL), t & 1);
      tc_fence_after();
      for (int hb = 0; hb < hboxes; ++hb) {
        mbar_wait(ring_full(stage), phase);
        tc_fence_after();
        if (leader) {
          const uint64_t db0 = dmn + (uint64_t)(ring_u + stage * box_u);
#pragma unroll
          for (int kk = 0; kk < kBN / 16; ++kk) {   // K = 16 rows of Y (j) per instruction
            const uint64_t da = dk + (uint64_t)(g_u + (kk >> 2) * box_u + (kk & 3) * 2);
            umma_f16(tmem_base + 256 + hb * 64, da, db0 + (uint64_t)(kk * 128), idesc_dx, (t > 0 || kk > 0) ? 1u : 0u);
          }
          release(stage);
          if (hb == hboxes - 1) umma_commit(bar(B_GEMPTY));
        }
        __syncwarp();
        if (++stage == P.stages) { stage = 0; phase ^= 1; }
      }
    };
    for (int j = 0; j < nt; ++j) {
      const int b = j & 1;
      mbar_wait(bar(B_TEMPTY0 + b), ((j >> 1) & 1) ^ 1);
      tc_fence_after();
      for (int kb = 0; kb < P.kboxes; ++kb) {
        mbar_wait(ring_full(stage), phase);
        tc_fence_after();
        if (leader) {
          const uint64_t da0 = dk + (uint64_t)(x_u + kb * box_u);
          const uint64_t db0 = dk + (uint64_t)(ring_u + stage * box_u);
#pragma unroll
          for (int k = 0; k < kBoxK / 16; ++k)
            umma_f16(tmem_base + b * kBN, da0 + (uint64_t)(k * 2), db0 + (uint64_t)(k * 2), idesc_s, (kb > 0 || k > 0) ? 1u : 0u);
          release(stage);
          if (kb == P.kboxes - 1) umma_commit(bar(B_TFULL0 + b));
        }
        __syncwarp();
        if (++stage == P.stages) { stage = 0; phase ^= 1; }
      }
      if (MODE == MODE_BWD && j >= 1) issue_dx(j - 1);
    }
    if (MODE == MODE_BWD) {
      issue_dx(nt - 1);
      if (leader) umma_commit(bar(B_DXFULL));
      __syncwarp();
    }
  } else {
    // =========================== epilogue warps (128 threads, thread = row) ===========================
    const int sub = warp & 3;                         // TMEM sub-partition this warp may access
    const int row_in_blk = sub * 32 + lane;
    const int row = iblk * kBM + row_in_blk;          // local X row
    const bool row_ok = row < pb.n_x;
    const uint32_t t_lane = tmem_base + ((uint32_t)(sub * 32) << 16);
    const int et = threadIdx.x - 64;                  // 0..127
    if (MODE == MODE_FWD) {
      const float s2 = fminf(expf(*P.logit_scale), P.scale_cap) * kLog2e;
      float m = -INFINITY, l = 0.f;
      for (int j = 0; j < nt; ++j) {
        const int b = j & 1;
        mbar_wait(bar(B_TFULL0 + b), (j >> 1) & 1);
        tc_fence_after();
        const int col0 = (tile0 + j) * kBN;
        const bool edge = col0 + kBN > pb.n_y;
#pragma unroll 1
        for (int c = 0; c < kBN / 32; ++c) {
          uint32_t r[32];
          tmem_ld32(t_lane + b * kBN + c * 32, r);
          tmem_ld_wait();
          float cmax = -INFINITY;
          if (edge) {
#pragma unroll
            for (int e = 0; e < 32; ++e)
              if (col0 + c * 32 + e >= pb.n_y) r[e] = 0xff800000u;   // -inf
          }
#pragma unroll
          for (int e = 0; e < 32; ++e) cmax = fmaxf(cmax, __uint_as_float(r[e]));
          const float mn = fmaxf(m, cmax * s2);
          if (mn != -INFINITY) {
            l *= ex2_approx(m - mn);
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
            for (int e = 0; e < 32; e += 4) {
              a0 += ex2_approx(fmaf(__uint_as_float(r[e + 0]), s2, -mn));
              a1 += ex2_approx(fmaf(__uint_as_float(r[e + 1]), s2, -mn));
              a2 += ex2_approx(fmaf(__uint_as_float(r[e + 2]), s2, -mn));
              a3 += ex2_approx(fmaf(__uint_as_float(r[e + 3]), s2, -mn));
            }
            l += (a0 + a1) + (a2 + a3);
            m = mn;
          }
        }
        tc_fence_before();
        mbar_arrive(bar(B_TEMPTY0 + b));
      }
      if (row_ok) reinterpret_cast<float2*>(pb.out)[(int64_t)chunk * pb.n_x + row] = make_float2(m, l);
    } else {
      const float s2 = P.scale[0] * kLog2e;
      const float lb = P.ln_B;
      const float rl2 = row_ok ? (pb.lse_x[pb.diag_offset + row] + lb) * kLog2e : INFINITY;
      const int dcol = row + pb.diag_offset;          // global Y index of this row's positive
      float* cl_s = reinterpret_cast<float*>(sptr + L.cl);
      float dsc = 0.f;
      for (int j = 0; j < nt; ++j) {
        const int b = j & 1;
        const int col0 = (tile0 + j) * kBN;
        {   // stage this tile's column lse (base-2, + log2 B); +inf masks columns past n_y
          const int cj = col0 + et;
          cl_s[b * 128 + et] = (cj < pb.n_y) ? (pb.lse_y[cj] + lb) * kLog2e : INFINITY;
        }
        epi_bar_sync();
        mbar_wait(bar(B_TFULL0 + b), (j >> 1) & 1);
        tc_fence_after();
        const bool has_diag = row_ok && dcol >= col0 && dcol < col0 + kBN;
#pragma unroll 1
        for (int c = 0; c < kBN / 32; ++c) {
          uint32_t r[32];
          tmem_ld32(t_lane + b * kBN + c * 32, r);
          tmem_ld_wait();
          uint32_t packed[16];
          const float4* cl4 = reinterpret_cast<const float4*>(cl_s + b * 128 + c * 32);
#pragma unroll
          for (int e4 = 0; e4 < 8; ++e4) {
            const float4 cl = cl4[e4];
            const float clv[4] = {cl.x, cl.y, cl.z, cl.w};
            float gv[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float acc = __uint_as_float(r[e4 * 4 + q]);
              float g = ex2_approx(fmaf(acc, s2, -rl2)) + ex2_approx(fmaf(acc, s2, -clv[q]));
              if (clv[q] == INFINITY) g = 0.f;                       // column past n_y
              gv[q] = g;
            }
            if (has_diag) {
#pragma unroll
              for (int q = 0; q < 4; ++q)
                if (col0 + c * 32 + e4 * 4 + q == dcol) gv[q] -= 2.0f * P.inv_B;
            }
            if (!row_ok) { gv[0] = gv[1] = gv[2] = gv[3] = 0.f; }
#pragma unroll
            for (int q = 0; q < 4; ++q) dsc = fmaf(gv[q], __uint_as_float(r[e4 * 4 + q]), dsc);
            packed[e4 * 2 + 0] = pack_bf16x2(gv[0], gv[1]);
            packed[e4 * 2 + 1] = pack_bf16x2(gv[2], gv[3]);
          }
          if (c == 0) mbar_wait(bar(B_GEMPTY), (j & 1) ^ 1);   // dX(j-1) has consumed the previous G
          // K-major SW128 store: row = i, 16-byte chunk index XOR (i & 7); 64 columns per box
          uint8_t* gbox = sptr + L.g + (c >> 1) * kBoxBytes + row_in_blk * 128;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int chunk16 = ((c & 1) * 4 + q) ^ (row_in_blk & 7);
            *reinterpret_cast<uint4*>(gbox + chunk16 * 16) =
                make_uint4(packed[q * 4 + 0], packed[q * 4 + 1], packed[q * 4 + 2], packed[q * 4 + 3]);
          }
        }
        tc_fence_before();
        mbar_arrive(bar(B_TEMPTY0 + b));
        fence_async_smem();               // generic-proxy writes of G -> visible to the tensor core
        mbar_arrive(bar(B_GFULL));
      }
      // ---- drain the dX accumulator of this half: TMEM -> registers -> fp32 partial in global memory
      mbar_wait(bar(B_DXFULL), 0);
      tc_fence_after();
      const int dh = hboxes * 64;
      float* orow = pb.out + ((int64_t)chunk * pb.n_x + row) * P.D + half * dh;
#pragma unroll 1
      for (int c = 0; c < dh / 32; ++c) {
        uint32_t r[32];
        tmem_ld32(t_lane + 256 + c * 32, r);
        tmem_ld_wait();
        if (row_ok) {
#pragma unroll
          for (int q = 0; q < 8; ++q)
            reinterpret_cast<uint4*>(orow + c * 32)[q] = make_uint4(r[q * 4], r[q * 4 + 1], r[q * 4 + 2], r[q * 4 + 3]);
        }
      }
      if (pb.dscale) {                      // fixed-order block reduction of sum G*cos
        float* red = reinterpret_cast<float*>(sptr + L.red);
        const float v = warp_sum(dsc);
        if (lane == 0) red[sub] = v;
        epi_bar_sync();
        if (et == 0) pb.dscale[u] = (half == 0) ? ((red[0] + red[1]) + (red[2] + red[3])) : 0.f;   // u: CTA slot
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
  if (P.cluster > 1) cluster_sync_all();     // no CTA exits while a peer may still multicast into it / arrive on it
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// [rows][D] bf16 row-major, box = 128 rows x 64 elements, 128-byte swizzle, zero fill out of bounds.
static int make_map(CUtensorMap* m, const void* base, int64_t rows, int D, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return set_error(VPA_E_NO_DEVICE, "cuTensorMapEncodeTiled entry point unavailable");
  // The encode is a DRIVER call and needs a current context in THIS thread; a thread that has not touched the
  // runtime yet (e.g. an autograd worker on device 0) has none -> bind the primary context once per thread.
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) {
    cudaFree(nullptr);
    ctx_bound = true;
  }
  cuuint64_t dims[2] = {(cuuint64_t)D, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)D * 2};
  cuuint32_t box[2] = {(cuuint32_t)kBoxK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(VPA_E_INVALID, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return 0;
}

static int pick_stages(int mode, int kboxes) {
  int st = kMaxStages;
  while (st > 2 && smem_layout(mode, kboxes, st).total > kSmemLimit) --st;
  return st;
}

template <int MODE>
static int launch(const SweepArgs& a, const Workspace& ws, const SweepPlan& plan, cudaStream_t st) {
  VPA_CHECK_ARG(a.D % 64 == 0 && a.D >= 64 && a.D <= 512, "tensor-core path needs D %% 64 == 0, 64 <= D <= 512 (D=%d)", a.D);
  VPA_CHECK_ARG(a.rows_global < (1ll << 30), "rows_global too large");
  for (int i = 0; i < 2; ++i)
    VPA_CHECK_ARG((reinterpret_cast<uintptr_t>(a.x[i]) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.y[i]) & 15) == 0,
                  "operands must be 16-byte aligned");
  CUtensorMap maps[4];
  for (int p = 0; p < 2; ++p) {
    if (int e = make_map(&maps[2 * p + 0], a.x[p], a.rows_local, a.D, kBM)) return e;
    if (int e = make_map(&maps[2 * p + 1], a.y[p], a.rows_global, a.D, kBN / plan.cluster)) return e;
  }
  Params P{};
  const bool bwd = MODE == MODE_BWD;
  P.n_iblk = plan.n_iblk;
  P.halves = bwd ? plan.halves : 1;
  P.n_chunks = bwd ? plan.bwd_chunks : plan.fwd_chunks;
  P.tiles_per_chunk = bwd ? plan.bwd_tiles_per_chunk : plan.fwd_tiles_per_chunk;
  P.n_tiles = plan.n_tiles;
  P.cluster = plan.cluster;
  P.n_igroups = (P.n_iblk + P.cluster - 1) / P.cluster;
  P.units_per_problem = P.n_igroups * P.cluster * P.halves * P.n_chunks;
  P.D = a.D;
  P.kboxes = a.D / 64;
  P.stages = pick_stages(MODE, P.kboxes);
  P.logit_scale = a.logit_scale;
  P.scale_cap = a.scale_cap;
  P.scale = a.scale;
  P.inv_B = 1.0f / (float)a.rows_global;
  P.ln_B = logf((float)a.rows_global);
  for (int p = 0; p < 2; ++p) {
    P.p[p].n_x = (int)a.rows_local;
    P.p[p].n_y = (int)a.rows_global;
    P.p[p].diag_offset = (int)a.row_offset;
    P.p[p].lse_x = a.lse_x[p];
    P.p[p].lse_y = a.lse_y[p];
    if (bwd) {
      P.p[p].out = ws.bwd_part + (int64_t)p * P.n_chunks * a.rows_local * a.D;
      P.p[p].dscale = p == 0 ? ws.dscale_part : nullptr;
    } else {
      P.p[p].out = ws.fwd_part + (int64_t)p * P.n_chunks * a.rows_local * 2;
      P.p[p].dscale = nullptr;
    }
  }
  const SmemLayout L = smem_layout(MODE, P.kboxes, P.stages);
  static SmemAttrCache attr_cache[2];
  if (int e = ensure_dynamic_smem(attr_cache[MODE], sweep_kernel<MODE>, (int)kSmemLimit)) return e;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * P.units_per_problem);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = L.total;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = P.cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  prof_begin(bwd ? PROF_BWD_SWEEP : PROF_FWD_SWEEP, st);
  VPA_CUDA(cudaLaunchKernelEx(&cfg, sweep_kernel<MODE>, maps[0], maps[1], maps[2], maps[3], P));
  prof_end(bwd ? PROF_BWD_SWEEP : PROF_FWD_SWEEP, st);
  VPA_LAUNCH_CHECK(bwd ? "sweep_kernel<BWD>" : "sweep_kernel<FWD>");
  return 0;
}

}  // namespace tc

int tc_infonce_fwd(const SweepArgs& a, const Workspace& ws, const SweepPlan& plan, cudaStream_t st) {
  return tc::launch<tc::MODE_FWD>(a, ws, plan, st);
}
int tc_infonce_bwd(const SweepArgs& a, const Workspace& ws, const SweepPlan& plan, cudaStream_t st) {
  return tc::launch<tc::MODE_BWD>(a, ws, plan, st);
}

}  // namespace vpa
