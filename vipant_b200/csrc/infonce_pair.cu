// InfoNCE sweeps on CTA PAIRS (tcgen05 cta_group::2): two SMs of one TPC execute every MMA together, each
// holding half of the B operand, so the shared-memory operand traffic per SM drops below the 128 B/clk the
// tensor core can read -- the bound of the single-CTA kernel (infonce_tc.cu, tensor pipe 51 % active, ncu).
//
// forward (MODE_FWD)   pair = 256 X rows (128 per CTA, resident in shared memory), tile = 256 Y rows:
//     S = X . Y^T      UMMA M=256 N=256 K=16, fp32 in TMEM (128 lanes x 256 columns per CTA, double-buffered);
//     epilogue warps keep an online (max, sum exp2) per row -> per-chunk partial statistics.
// backward (MODE_BWD)  pair = 128 X rows (64 per CTA).  UMMA M=128 in cta_group::2 puts 64 rows in each CTA
//     with the N columns folded over the two lane halves (lanes 0-63: columns [0,N/2), lanes 64-127: [N/2,N)),
//     so a 64 x 512 fp32 dX accumulator needs only 256 TMEM columns per CTA and leaves 2 x 128 columns for a
//     double-buffered 64 x 256 logit tile: S recompute and the dX contraction run in ONE sweep with NO
//     duplicated work (the single-CTA kernel recomputes S once per D-half):
//       S_ij = X_i . Y_j^T                          (32 x UMMA M128 N256 K16 per 256-row tile)
//       G_ij = (exp(S-lse_x) + exp(S-lse_y) - 2 delta)/B   in registers -> bf16 -> shared memory (K-major SW128)
//       dX_i += G_ij . Y_j                          (2 x 16 x UMMA M128 N256 K16, B operand = Y MN-major)
// Both "problems" (A against T, T against A) run in one launch.  Work per step at D = 512:
// forward 4 B^2 D flops (S and S^T), backward 8 B^2 D; the B x B logits only ever exist as TMEM tiles.
//
// Reference semantics: loss_head.py:277-283 (logits + 2 x cross entropy) and its autograd.
#include <cuda.h>

#include <cstdio>

#include "common.cuh"
#include "p2p.cuh"

namespace vpa {
namespace pr {

constexpr int kBN = 256;                      // Y rows per tile (UMMA N of the S product); 128 loaded by each CTA
constexpr int kBoxK = 64;                     // bf16 elements per 128-byte swizzle row
constexpr int kYBox = 128 * 128;              // one TMA box of Y: [128 rows][64 elems] = 16 KB
constexpr int kStage = 2 * kYBox;             // ring stage: two boxes (32 KB)
constexpr int kStages = 3;
constexpr int kThreadsFwd = 192;              // warp 0: TMA, warp 1: MMA issue + TMEM alloc, warps 2..5: epilogue
constexpr int kThreadsBwd = 320;              // backward: 8 epilogue warps (two per TMEM sub-partition / scheduler)
constexpr uint32_t kSmemLimit = 232448;

enum { MODE_FWD = 0, MODE_BWD = 1, MODE_FWD1 = 2 };
constexpr float kFastS2Limit = 62.0f;     // single-pass forward valid while 2*s*log2(e) stays inside the fp32 exponent range

constexpr int kMaxProblems = 2 * kMaxPairs;   // every InfoNCE pair is swept in both directions
constexpr int kMaxMaps = 2 * kMaxPairs;       // distinct (matrix, box height) tensor maps of one launch

struct Problem {
  int n_x, n_y;
  int diag_offset;
  int mapx, mapy;           // tensor maps of the X rows (box = rows per CTA) and the Y rows (box = 128)
  const float* lse_x;
  const float* lse_y;
  float* out;               // FWD: float2[n_chunks][n_x]; BWD: float[n_chunks][n_x][D]
  float* dscale;            // BWD, first direction of a pair: one partial per CTA; nullptr otherwise
  const float* logit_scale; // FWD: s = min(exp(*logit_scale), scale_cap) -- every pair has its own temperature
  float scale_cap;
  const float* scale;       // BWD: {s, flows} as written by the forward
  float* colpart;           // FWD1: float[n_iblk*8][n_y] column sums of exp2(S*s2 - s2) per 32-row group
};

struct TensorMaps {
  CUtensorMap m[kMaxMaps];
};

struct Params {
  Problem p[kMaxProblems];
  int pairs_per_problem;    // n_iblk * n_chunks
  int n_iblk, n_chunks, tiles_per_chunk, n_tiles;
  // Every X row block is swept by n_big equal chunks of tiles_per_chunk tiles plus (small_tiles > 0) one short tail chunk.
  // All big units come first in the grid (all problems), the tails last: CTAs are dispatched in blockIdx order to the
  // first free SM pair, so the tails fill the slots a single wave of big units leaves idle (LPT scheduling; api.cu).
  int n_big, small_tiles, n_prob;
  int D, kboxes, nblk;      // kboxes = D / 64; nblk = D / 256 accumulator blocks (BWD)
  float inv_B, ln_B;
  int gate;                 // 0: always run; 1: run only if s*log2e <= kFastS2Limit; 2: run only if it is larger (per problem)
  P2PRowFlags yflags;       // FWD1 over peer memory: arrival flags of the Y rows (nullptr: everything is already there)
  P2PRowFlags aflags;       // BWD over peer memory: arrival flags of problem 1's Y rows (the x1 operands of the peers)
  int rot;                  // BWD: every unit visits tile (t + rot) mod n_tiles -- the local rank block first, then the peers'
                            // blocks in the order the relay CTAs fetch them
  RelayArgs relay;          // FWD1 / BWD over peer memory: the first relay.n_ctas CTAs of the grid are the operand all-gather
};

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// shared::cluster address of the same shared-memory offset in CTA `rank` of this cluster
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// arrive on a barrier that may live in the peer CTA (cluster address).  Default (.release.cta) semantics, as
// CUTLASS' umma_arrive_2x1SM_sm0: what is handed over is TMEM (ordered by tcgen05.fence) and shared memory
// already published with fence.proxy.async; cluster-scope release/acquire costs an ERRBAR / CCTL.IVALL per
// arrive / wait (15 % of the forward kernel's samples in ncu).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
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
  printf("vipant_b200(pair): mbarrier wait timeout (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
  asm volatile("trap;");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try(bar, parity)) return;
  uint32_t spins = 0;
  while (!mbar_try(bar, parity))
    if (++spins > (1u << 21)) mbar_timeout(bar, parity);
}
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
// TMA load executed by either CTA of the pair into ITS shared memory; the bytes are accounted on the
// LEADER CTA's mbarrier (`leader_bar` is a shared::cluster address).
__device__ __forceinline__ void tma_load_2sm(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t leader_bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(leader_bar)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// D[tmem] (+)= A[smem] . B[smem] over the CTA pair, issued by ONE thread of the leader CTA.
__device__ __forceinline__ void umma2_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once all previously issued MMAs completed) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma2_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}
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
template <int N>
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(N) : "memory"); }

// Shared-memory matrix descriptor, 128B swizzle (see infonce_tc.cu::make_desc)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---------------------------------------------------------------- shared-memory carve-up (identical in both CTAs)
template <int MODE>
struct Cfg {
  static constexpr int RC = (MODE == MODE_BWD) ? 64 : 128;     // X rows per CTA
  static constexpr int RP = 2 * RC;                            // X rows per pair
  static constexpr int XBox = RC * 128;                        // one [RC rows][64 elems] box
};
struct SmemLayout {
  uint32_t x, g, ring, cl, red, bars, tmem_slot, total;
};
template <int MODE>
__host__ __device__ inline SmemLayout smem_layout(int kboxes) {
  SmemLayout L;
  uint32_t o = 0;
  L.x = o; o += kboxes * Cfg<MODE>::XBox;
  L.g = o; o += (MODE == MODE_BWD) ? 2 * 32768 : 0;            // two G buffers: [64 rows][256 j] bf16 each
  L.ring = o; o += kStages * kStage;
  L.cl = o; o += (MODE == MODE_BWD) ? 2 * 256 * 4 : 0;         // column lse of the current / next tile
  L.red = o; o += 64;
  L.bars = o; o += 32 * 8;
  L.tmem_slot = o; o += 16;
  L.total = o;                                                 // + alignment pad of the dynamic base (checked in the kernel)
  return L;
}
// barrier indices
enum { B_XFULL = 0, B_TFULL0, B_TFULL1, B_TEMPTY0, B_TEMPTY1, B_GFULL0, B_GFULL1, B_GEMPTY0, B_GEMPTY1, B_DXFULL,
       B_RFULL, B_REMPTY = B_RFULL + kStages, B_COUNT = B_REMPTY + kStages };

template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(MODE == MODE_FWD ? kThreadsFwd : kThreadsBwd, 1)
pair_kernel(const __grid_constant__ TensorMaps M, const Params P) {
  using C = Cfg<MODE>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sptr = smem_raw + (sbase - smem_u32(smem_raw));
  const SmemLayout L = smem_layout<MODE>(P.kboxes);
  if (sbase - smem_u32(smem_raw) + L.total > kSmemLimit) asm volatile("trap;");     // alignment pad does not fit
  if ((MODE == MODE_FWD1 || MODE == MODE_BWD) && (int)blockIdx.x < P.relay.n_ctas) {
    // Relay CTAs: the all-gather of the operand rows over NVLink, in the same grid as the sweep that consumes them.  They
    // come first in blockIdx order, i.e. they are resident before any sweep CTA that polls their flags.  The forward
    // fetches the x2 operands (all it reads); the x1 operands, which only the backward's second problem reads, travel in
    // the backward's grid while its first problem already runs -- except in the exact temperature regime, whose two-sweep
    // forward kernel reads them: then the forward's relays fetch them as well and the backward's find them in place.
    if (blockIdx.x == 0) relay_signal_ready(P.relay);
    int m1 = P.relay.m1;
    if (MODE == MODE_FWD1 && fminf(expf(*P.p[0].logit_scale), P.p[0].scale_cap) * kLog2e > kFastS2Limit) m1 = 2;
    relay_pull(P.relay, m1, blockIdx.x, sptr);
    return;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();            // 0 = leader (issues the MMAs), 1 = peer
  const bool is_leader = crank == 0;

  int u = ((int)blockIdx.x - ((MODE == MODE_FWD1 || MODE == MODE_BWD) ? P.relay.n_ctas : 0)) >> 1;   // pair index (after the relay CTAs)
  int prob_, chunk_, iblk_, tile0_, nt_;
  {
    const int per = P.n_iblk * P.n_big, big_units = P.n_prob * per;
    if (u < big_units) {
      prob_ = u / per;
      u -= prob_ * per;
      chunk_ = u / P.n_iblk;
      iblk_ = u - chunk_ * P.n_iblk;
      tile0_ = chunk_ * P.tiles_per_chunk;
      nt_ = min(P.tiles_per_chunk, (P.n_tiles - P.small_tiles) - tile0_);
    } else {
      u -= big_units;
      prob_ = u / P.n_iblk;
      iblk_ = u - prob_ * P.n_iblk;
      chunk_ = P.n_big;
      tile0_ = P.n_tiles - P.small_tiles;
      nt_ = P.small_tiles;
    }
  }
  const int prob = prob_, chunk = chunk_, iblk = iblk_, tile0 = tile0_, nt = nt_;
  const Problem& pb = P.p[prob];
  if (P.gate != 0) {       // regime gate on the DEVICE value of this pair's temperature (no host sync): uniform over its CTAs
    const float gs2 = fminf(expf(*pb.logit_scale), pb.scale_cap) * kLog2e;
    if ((P.gate == 1) != (gs2 <= kFastS2Limit)) return;
  }
  const CUtensorMap* mapx = &M.m[pb.mapx];
  const CUtensorMap* mapy = &M.m[pb.mapy];
  const int row0 = iblk * C::RP + (int)crank * C::RC;   // first local X row of this CTA
  // Order in which this CTA visits its tiles.  Over peer memory (single-pass forward) the rows of every peer block
  // arrive chunk by chunk, all blocks at the same pace: a CTA whose range touches several rank blocks visits it
  // chunk-major (chunk 0 of each block, chunk 1 of each block, ...) instead of one block after the other, so it never
  // needs the LAST chunk of one peer before the first of the next.  Producer and epilogue run the same iterator.
  struct TileOrder {
    int cpr, blk0, blk1, lo, hi, c, blk, lin;      // cpr == 0: ascending
    __device__ int first() {
      if (cpr == 0) return lin;
      c = 0; blk = blk0 - 1;
      return next();
    }
    __device__ int next() {
      if (cpr == 0) return ++lin;
      while (true) {
        if (++blk > blk1) { blk = blk0; ++c; }
        const int t = blk * cpr + c;
        if (t >= lo && t < hi) return t;
      }
    }
  };
  auto make_order = [&]() {
    TileOrder o{};
    o.cpr = 0; o.lin = tile0; o.lo = tile0; o.hi = tile0 + nt;
    if (MODE == MODE_FWD1 && P.yflags.flags != nullptr && P.yflags.rows_per_rank % kBN == 0) {
      const int cpr = P.yflags.chunks_per_rank;
      const int b0 = tile0 / cpr, b1 = (tile0 + nt - 1) / cpr;
      if (b1 > b0) { o.cpr = cpr; o.blk0 = b0; o.blk1 = b1; }
    }
    return o;
  };
  // BWD: the j-th tile of this unit (rotated start, see Params::rot)
  auto tile_at = [&](int j) {
    const int t = tile0 + j + P.rot;
    return t >= P.n_tiles ? t - P.n_tiles : t;
  };
  const uint32_t bar0 = sbase + L.bars;
  auto bar = [&](int i) { return bar0 + 8u * i; };
  auto lbar = [&](int i) { return mapa(bar0 + 8u * i, 0); };      // the leader's copy (cluster address)
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sptr + L.tmem_slot);

  if (threadIdx.x == 0) {
    mbar_init(bar(B_XFULL), 1);
    mbar_init(bar(B_TFULL0), 1);
    mbar_init(bar(B_TFULL1), 1);
    constexpr int kEpiWarps2 = (MODE == MODE_FWD) ? 8 : 16;    // epilogue warps of BOTH CTAs (the leader's copy is used)
    mbar_init(bar(B_TEMPTY0), kEpiWarps2);
    mbar_init(bar(B_TEMPTY1), kEpiWarps2);
    mbar_init(bar(B_GFULL0), kEpiWarps2);
    mbar_init(bar(B_GFULL1), kEpiWarps2);
    mbar_init(bar(B_GEMPTY0), 1);
    mbar_init(bar(B_GEMPTY1), 1);
    mbar_init(bar(B_DXFULL), 1);
    for (int s = 0; s < kStages; ++s) { mbar_init(bar(B_RFULL + s), 1); mbar_init(bar(B_REMPTY + s), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    prefetch_tmap(mapx);
    prefetch_tmap(mapy);
  }
  if (warp == 1) tmem_alloc2(sbase + L.tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =========================== TMA producer (both CTAs; each fills ITS half of every operand) ===========================
    const bool elected = elect_one();
    if (elected) {
      if (is_leader) mbar_arrive_expect_tx(bar(B_XFULL), 2 * P.kboxes * C::XBox);
      const uint32_t xf = lbar(B_XFULL);
      for (int kb = 0; kb < P.kboxes; ++kb) tma_load_2sm(sbase + L.x + kb * C::XBox, mapx, kb * kBoxK, row0, xf);
    }
    int stage = 0;
    uint32_t phase = 0;
    // one ring stage = two boxes at (c0a, r), (c0b, r)
    auto push = [&](int c0a, int c0b, int r) {
      mbar_wait(bar(B_REMPTY + stage), phase ^ 1);
      if (elected) {
        if (is_leader) mbar_arrive_expect_tx(bar(B_RFULL + stage), 2 * kStage);
        const uint32_t fb = lbar(B_RFULL + stage);
        const uint32_t dst = sbase + L.ring + stage * kStage;
        tma_load_2sm(dst, mapy, c0a, r, fb);
        tma_load_2sm(dst + kYBox, mapy, c0b, r, fb);
      }
      if (++stage == kStages) { stage = 0; phase ^= 1; }
    };
    auto push_dx = [&](int t) {       // Y rows of tile t as the MN-major B operand: this CTA's 128 of every 256 d columns
      for (int jh = 0; jh < 2; ++jh)
        for (int nb = 0; nb < P.nblk; ++nb) {
          const int d0 = nb * 256 + (int)crank * 128;
          push(d0, d0 + 64, tile_at(t) * kBN + jh * 128);
        }
    };
    TileOrder order = make_order();
    int tcur = order.first();
    const bool polling = (MODE == MODE_FWD1 && P.yflags.flags != nullptr) || (MODE == MODE_BWD && prob == 1 && P.aflags.flags != nullptr);
    int verified = 0;               // tiles, from the current one on, whose rows are known to have arrived
    for (int j = 0; j < nt; ++j, tcur = (j < nt) ? order.next() : tcur) {
      if (MODE == MODE_BWD) tcur = tile_at(j);
      const int r = tcur * kBN + (int)crank * 128;             // this CTA's half of the tile's Y rows
      if (polling) {
        // Peer-memory all-gather in flight: these rows may still be on their way from another GPU (p2p.cuh).  One poll
        // (a system-scope acquire load: an L2 round trip) per tile would add ~0.5 us to every tile of the sweep, so the
        // whole warp looks ahead: lane L tests the L-th upcoming tile, lane 0 waits for the current one, and the sweep does
        // not poll again until it has used up the tiles found complete (the transfer normally runs well ahead of it).
        if (verified == 0) {
          const P2PRowFlags& fl = MODE == MODE_BWD ? P.aflags : P.yflags;
          bool ok = false;
          if (j + lane < nt) {
            int t = tcur;
            if (MODE == MODE_BWD) {
              t = tile_at(j + lane);
            } else {
              TileOrder ahead = order;
              for (int k = 0; k < lane; ++k) t = ahead.next();
            }
            const int rl = t * kBN + (int)crank * 128;
            if (lane == 0) { p2p_wait_rows(fl, rl, min(rl + 128, pb.n_y)); ok = true; }
            else ok = p2p_rows_ready(fl, rl, min(rl + 128, pb.n_y));
          }
          const unsigned mask = __ballot_sync(0xffffffffu, ok);
          verified = __ffs(~mask) - 1;                          // consecutive complete tiles from the current one (>= 1)
          if (verified < 0) verified = 32;
          __syncwarp();                                         // the lanes' observations are ordered before ...
          if (elected) asm volatile("fence.acq_rel.sys;\n\tfence.proxy.async;" ::: "memory");   // ... the TMA (async proxy) reads
          __syncwarp();
        }
        --verified;
      }
      for (int kb = 0; kb < P.kboxes; kb += 2) push(kb * kBoxK, (kb + 1) * kBoxK, r);
      if (MODE == MODE_BWD && j >= 1) push_dx(j - 1);
    }
    if (MODE == MODE_BWD) push_dx(nt - 1);
  } else if (warp == 1) {
    // =========================== MMA issuer (leader CTA only) ===========================
    if (is_leader) {
      const bool elected = elect_one();
      constexpr uint32_t idesc_s = make_idesc(C::RP, kBN, 0, 0);
      constexpr uint32_t idesc_dx = make_idesc(128, 256, 0, 1);
      const uint64_t dk = make_desc(sbase, 16, 1024);            // K-major SW128 operand template
      const uint64_t dmn = make_desc(sbase, kYBox, 1024);        // MN-major: two 64-column atoms, 16 KB apart
      const uint32_t ring_u = L.ring >> 4, x_u = L.x >> 4, g_u = L.g >> 4;
      constexpr uint32_t xbox_u = C::XBox >> 4, ybox_u = kYBox >> 4, stage_u = kStage >> 4;
      const uint32_t s_cols = (MODE == MODE_BWD) ? 128u : 256u;  // TMEM columns of one S buffer
      int stage = 0;
      uint32_t phase = 0;
      mbar_wait(bar(B_XFULL), 0);
      tc_fence_after();
      auto issue_dx = [&](int t) {
        const int g = t & 1;
        mbar_wait(bar(B_GFULL0 + g), (t >> 1) & 1);
        tc_fence_after();
        for (int jh = 0; jh < 2; ++jh)
          for (int nb = 0; nb < P.nblk; ++nb) {
            mbar_wait(bar(B_RFULL + stage), phase);
            tc_fence_after();
            if (elected) {
              const uint64_t db0 = dmn + (uint64_t)(ring_u + stage * stage_u);
              const uint64_t da0 = dk + (uint64_t)(g_u + g * 2048 + jh * 1024);       // G buffer g, boxes 2*jh, 2*jh+1 (8 KB each)
#pragma unroll
              for (int kk = 0; kk < 8; ++kk)       // 16 Y rows (j) per instruction
                umma2_f16(tmem_base + 256 + nb * 128, da0 + (uint64_t)((kk >> 2) * 512 + (kk & 3) * 2), db0 + (uint64_t)(kk * 128),
                          idesc_dx, (t > 0 || jh > 0 || kk > 0) ? 1u : 0u);
              umma2_commit(bar(B_REMPTY + stage));
              if (jh == 1 && nb == P.nblk - 1) umma2_commit(bar(B_GEMPTY0 + g));
            }
            __syncwarp();
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
      };
      for (int j = 0; j < nt; ++j) {
        const int b = j & 1;
        mbar_wait(bar(B_TEMPTY0 + b), ((j >> 1) & 1) ^ 1);
        tc_fence_after();
        for (int kb = 0; kb < P.kboxes; kb += 2) {
          mbar_wait(bar(B_RFULL + stage), phase);
          tc_fence_after();
          if (elected) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const uint64_t da0 = dk + (uint64_t)(x_u + (kb + h) * xbox_u);
              const uint64_t db0 = dk + (uint64_t)(ring_u + stage * stage_u + h * ybox_u);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma2_f16(tmem_base + b * s_cols, da0 + (uint64_t)(k * 2), db0 + (uint64_t)(k * 2), idesc_s,
                          (kb > 0 || h > 0 || k > 0) ? 1u : 0u);
            }
            umma2_commit(bar(B_REMPTY + stage));
            if (kb + 2 >= P.kboxes) umma2_commit(bar(B_TFULL0 + b));
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if (MODE == MODE_BWD && j >= 1) issue_dx(j - 1);
      }
      if (MODE == MODE_BWD) {
        issue_dx(nt - 1);
        if (elected) umma2_commit(bar(B_DXFULL));
        __syncwarp();
      }
    }
  } else {
    // =========================== epilogue warps (128 threads per CTA) ===========================
    const int sub = warp & 3;                         // TMEM sub-partition (lanes 32*sub .. +32) this warp may access
    const uint32_t t_lane = tmem_base + ((uint32_t)(sub * 32) << 16);
    if (MODE == MODE_FWD) {
      const int row = row0 + sub * 32 + lane;         // lane = row (128 rows per CTA)
      const bool row_ok = row < pb.n_x;
      const float s2 = fminf(expf(*pb.logit_scale), pb.scale_cap) * kLog2e;
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
          tmem_ld32(t_lane + b * 256 + c * 32, r);
          tmem_ld_wait();
          if (edge) {
#pragma unroll
            for (int e = 0; e < 32; ++e)
              if (col0 + c * 32 + e >= pb.n_y) r[e] = 0xff800000u;   // -inf
          }
          float cmax = -INFINITY;
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
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(lbar(B_TEMPTY0 + b));
      }
      if (row_ok) reinterpret_cast<float2*>(pb.out)[(int64_t)chunk * pb.n_x + row] = make_float2(m, l);
    } else if (MODE == MODE_FWD1) {
      // Single-pass forward: ONE sweep of S yields both statistics.  With s*log2e <= 62 every term exp2(S*s2 - s2)
      // (cos <= 1) is a normal fp32 number, so a fixed reference replaces the online maximum and the SAME exponential
      // serves the row sums (thread-local) and the column sums (32x32 butterfly over the lanes of a warp, one value per
      // lane, written as per-32-row-group partials).  Eight epilogue warps: two per lane quadrant, 128 columns each.
      const int ch = (warp - 2) >> 2;                 // which 128 of the tile's 256 columns
      const int row = row0 + sub * 32 + lane;         // lane = row (128 rows per CTA)
      const bool row_ok = row < pb.n_x;
      const bool rows_full = row0 + C::RC <= pb.n_x;
      const float s2 = fminf(expf(*pb.logit_scale), pb.scale_cap) * kLog2e;
      float* cp = pb.colpart + (int64_t)(iblk * 8 + (int)crank * 4 + sub) * pb.n_y;
      float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
      TileOrder order = make_order();
      int tcur = order.first();
      for (int j = 0; j < nt; ++j, tcur = (j < nt) ? order.next() : tcur) {
        const int b = j & 1;
        mbar_wait(bar(B_TFULL0 + b), (j >> 1) & 1);
        tc_fence_after();
        const int col0 = tcur * kBN;
        const bool full = rows_full && (col0 + kBN <= pb.n_y);
#pragma unroll 1
        for (int cc = 0; cc < 4; ++cc) {
          const int c = ch * 4 + cc;
          const int cstart = col0 + c * 32;
          uint32_t r[32];
          tmem_ld32(t_lane + b * 256 + c * 32, r);
          tmem_ld_wait();
          float e[32];
#pragma unroll
          for (int k = 0; k < 32; ++k) e[k] = ex2_approx(fmaf(__uint_as_float(r[k]), s2, -s2));
          if (!full) {
#pragma unroll
            for (int k = 0; k < 32; ++k)
              if (!row_ok || cstart + k >= pb.n_y) e[k] = 0.f;
          }
#pragma unroll
          for (int k = 0; k < 32; k += 4) { l0 += e[k]; l1 += e[k + 1]; l2 += e[k + 2]; l3 += e[k + 3]; }
          // column sums over the warp's 32 rows: after the butterfly lane L holds column cstart + L
#pragma unroll
          for (int n = 16; n >= 1; n >>= 1) {
            const bool up = (lane & n) != 0;
#pragma unroll
            for (int i = 0; i < n; ++i) {
              const float keep = up ? e[i + n] : e[i];
              const float send = up ? e[i] : e[i + n];
              e[i] = keep + __shfl_xor_sync(0xffffffffu, send, n);
            }
          }
          if (cstart + lane < pb.n_y) cp[cstart + lane] = e[0];
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(lbar(B_TEMPTY0 + b));
      }
      // two warps (ch = 0, 1) hold halves of each row's sum: slot [chunk][ch][row] of a float2-per-row array
      if (row_ok) reinterpret_cast<float*>(pb.out)[((int64_t)chunk * pb.n_x + row) * 2 + ch] = (l0 + l1) + (l2 + l3);
    } else {
      // 2x2 TMEM layout of M=128 cta_group::2: lanes 0-63 hold this CTA's 64 rows x columns [0,128) of the tile,
      // lanes 64-127 the same rows x columns [128,256).  Eight epilogue warps: two per sub-partition (and per
      // scheduler), each owning 64 of the thread's 128 columns -- a single warp per scheduler is latency-bound
      // (ncu: 4.3 cycles per issued instruction).
      const int r_in = (sub & 1) * 32 + lane;         // row within the CTA's 64
      const int jh = sub >> 1;                        // which 128-column half of the tile this lane quadrant holds
      const int ch = (warp - 2) >> 2;                 // which 64 of those 128 columns this warp processes
      const int row = row0 + r_in;
      const bool row_ok = row < pb.n_x;
      const bool rows_full = row0 + C::RC <= pb.n_x;  // CTA-uniform
      const float s2 = pb.scale[0] * kLog2e;
      const float lb = P.ln_B;
      const float rl2 = row_ok ? (pb.lse_x[pb.diag_offset + row] + lb) * kLog2e : INFINITY;
      const int dcol = row + pb.diag_offset;
      const int dwarp0 = row0 + (sub & 1) * 32 + pb.diag_offset;     // diagonal columns of this warp: [dwarp0, dwarp0 + 32)
      const bool want_dsc = pb.dscale != nullptr;
      float* cl_s = reinterpret_cast<float*>(sptr + L.cl);
      float dsc = 0.f;
      const int etb = threadIdx.x - 64;               // 0..255
      auto load_cl = [&](int t) {                     // raw column lse of tile t for shared-memory slot etb; +inf masks columns past n_y
        const int cj = tile_at(t) * kBN + etb;
        return (cj < pb.n_y) ? __ldg(pb.lse_y + cj) : INFINITY;
      };
      float cl_raw = load_cl(0);
      for (int j = 0; j < nt; ++j) {
        const int b = j & 1;
        const int col0 = tile_at(j) * kBN;
        float* clb = cl_s + b * 256;                  // double-buffered: one barrier per tile
        clb[etb] = (cl_raw + lb) * kLog2e;            // base-2, + log2 B
        if (j + 1 < nt) cl_raw = load_cl(j + 1);      // prefetch (raw: nothing depends on it until the next tile)
        epi_bar_sync<256>();
        mbar_wait(bar(B_TFULL0 + b), (j >> 1) & 1);
        tc_fence_after();
        const bool full = rows_full && (col0 + kBN <= pb.n_y);      // no masking needed anywhere in this tile
        uint8_t* gbuf = sptr + L.g + b * 32768;       // G buffer = tile parity
        if (j >= 2) mbar_wait(bar(B_GEMPTY0 + b), ((j >> 1) - 1) & 1);   // dX(j-2) has consumed this G buffer
#pragma unroll 1
        for (int cc = 0; cc < 2; ++cc) {
          const int c = ch * 2 + cc;                  // 32-column chunk within the 128
          const int cstart = col0 + jh * 128 + c * 32;
          uint32_t r[32];
          tmem_ld32(t_lane + b * 128 + c * 32, r);
          tmem_ld_wait();
          uint32_t packed[16];
          const float4* cl4 = reinterpret_cast<const float4*>(clb + jh * 128 + c * 32);
          const bool diag_here = dwarp0 < cstart + 32 && dwarp0 + 32 > cstart;    // warp-uniform
          if (full && !diag_here) {
#pragma unroll
            for (int e4 = 0; e4 < 8; ++e4) {
              const float4 cl = cl4[e4];
              const float a0 = __uint_as_float(r[e4 * 4 + 0]), a1 = __uint_as_float(r[e4 * 4 + 1]);
              const float a2 = __uint_as_float(r[e4 * 4 + 2]), a3 = __uint_as_float(r[e4 * 4 + 3]);
              const float g0 = ex2_approx(fmaf(a0, s2, -rl2)) + ex2_approx(fmaf(a0, s2, -cl.x));
              const float g1 = ex2_approx(fmaf(a1, s2, -rl2)) + ex2_approx(fmaf(a1, s2, -cl.y));
              const float g2 = ex2_approx(fmaf(a2, s2, -rl2)) + ex2_approx(fmaf(a2, s2, -cl.z));
              const float g3 = ex2_approx(fmaf(a3, s2, -rl2)) + ex2_approx(fmaf(a3, s2, -cl.w));
              if (want_dsc) dsc = fmaf(g0, a0, fmaf(g1, a1, fmaf(g2, a2, fmaf(g3, a3, dsc))));
              packed[e4 * 2 + 0] = pack_bf16x2(g0, g1);
              packed[e4 * 2 + 1] = pack_bf16x2(g2, g3);
            }
          } else {
#pragma unroll
            for (int e4 = 0; e4 < 8; ++e4) {
              const float4 cl = cl4[e4];
              const float clv[4] = {cl.x, cl.y, cl.z, cl.w};
              float gv[4];
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float acc = __uint_as_float(r[e4 * 4 + q]);
                float g = ex2_approx(fmaf(acc, s2, -rl2)) + ex2_approx(fmaf(acc, s2, -clv[q]));
                if (clv[q] == INFINITY || !row_ok) g = 0.f;
                if (cstart + e4 * 4 + q == dcol && row_ok) g -= 2.0f * P.inv_B;
                gv[q] = g;
                dsc = fmaf(g, acc, dsc);
              }
              packed[e4 * 2 + 0] = pack_bf16x2(gv[0], gv[1]);
              packed[e4 * 2 + 1] = pack_bf16x2(gv[2], gv[3]);
            }
          }
          // K-major SW128: box (64 j columns, 8 KB) = jh*2 + ch; row = r_in; 16-byte chunk index XOR (row & 7)
          uint8_t* gbox = gbuf + (jh * 2 + ch) * 8192 + r_in * 128;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int chunk16 = (cc * 4 + q) ^ (r_in & 7);
            *reinterpret_cast<uint4*>(gbox + chunk16 * 16) =
                make_uint4(packed[q * 4 + 0], packed[q * 4 + 1], packed[q * 4 + 2], packed[q * 4 + 3]);
          }
        }
        tc_fence_before();
        fence_async_smem();               // generic-proxy writes of G -> visible to the tensor cores
        __syncwarp();
        if (lane == 0) {
          mbar_arrive_cluster(lbar(B_TEMPTY0 + b));
          mbar_arrive_cluster(lbar(B_GFULL0 + b));
        }
      }
      // ---- drain dX: TMEM -> registers -> fp32 partial.  Block nb: lanes 0-63 = d [nb*256, +128), 64-127 = d [nb*256+128, +128)
      mbar_wait(bar(B_DXFULL), 0);
      tc_fence_after();
      float* orow = pb.out + ((int64_t)chunk * pb.n_x + row) * P.D;
      for (int nb = 0; nb < P.nblk; ++nb) {
#pragma unroll 1
        for (int cc = 0; cc < 2; ++cc) {
          const int c = ch * 2 + cc;
          uint32_t r[32];
          tmem_ld32(t_lane + 256 + nb * 128 + c * 32, r);
          tmem_ld_wait();
          if (row_ok) {
            float* dst = orow + nb * 256 + jh * 128 + c * 32;
#pragma unroll
            for (int q = 0; q < 8; ++q)
              reinterpret_cast<uint4*>(dst)[q] = make_uint4(r[q * 4], r[q * 4 + 1], r[q * 4 + 2], r[q * 4 + 3]);
          }
        }
      }
      if (want_dsc) {                       // fixed-order block reduction of sum G*cos
        float* red = reinterpret_cast<float*>(sptr + L.red);
        const float v = warp_sum(dsc);
        if (lane == 0) red[warp - 2] = v;
        epi_bar_sync<256>();
        if (etb == 0)
          pb.dscale[(chunk * P.n_iblk + iblk) * 2 + (int)crank] =
              ((red[0] + red[1]) + (red[2] + red[3])) + ((red[4] + red[5]) + (red[6] + red[7]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                    // neither CTA frees TMEM / exits while the pair may still touch it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
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

template <int MODE>
static int launch(const PairLaunch& L, const SweepPlan& plan, int gate, cudaStream_t st) {
  VPA_CHECK_ARG(L.D == 256 || L.D == 512, "pair kernels need D in {256, 512} (D=%d)", L.D);
  VPA_CHECK_ARG(L.rows_global < (1ll << 30), "rows_global too large");
  VPA_CHECK_ARG(L.n_prob >= 1 && L.n_prob <= kMaxProblems, "pair kernels: 1..%d problems per launch (got %d)", kMaxProblems, L.n_prob);
  const bool bwd = MODE == MODE_BWD;
  const bool fwd1 = MODE == MODE_FWD1;
  Params P{};
  TensorMaps M{};
  // one tensor map per distinct (matrix, rows, box height): the pairs of a composite head share their operands
  struct Key { const void* base; int64_t rows; int box; };
  Key keys[kMaxMaps];
  int n_maps = 0;
  auto map_id = [&](const void* base, int64_t rows, int box, int* id) -> int {
    for (int i = 0; i < n_maps; ++i)
      if (keys[i].base == base && keys[i].rows == rows && keys[i].box == box) { *id = i; return 0; }
    if (n_maps == kMaxMaps) return set_error(VPA_E_UNSUPPORTED, "pair kernels: more than %d distinct operand maps in one launch", kMaxMaps);
    if (int e = make_map(&M.m[n_maps], base, rows, L.D, box)) return e;
    keys[n_maps] = Key{base, rows, box};
    *id = n_maps++;
    return 0;
  };
  P.n_iblk = bwd ? plan.pair_bwd_iblk : plan.pair_fwd_iblk;
  P.n_chunks = bwd ? plan.bwd_chunks : (fwd1 ? plan.fwd1_chunks : plan.fwd_chunks);
  P.tiles_per_chunk = bwd ? plan.bwd_tiles_per_chunk : (fwd1 ? plan.fwd1_tiles_per_chunk : plan.fwd_tiles_per_chunk);
  P.n_tiles = plan.n_tiles;
  P.pairs_per_problem = P.n_iblk * P.n_chunks;
  P.small_tiles = bwd ? plan.bwd_small : (fwd1 ? plan.fwd1_small : 0);
  P.n_big = P.n_chunks - (P.small_tiles > 0 ? 1 : 0);
  P.n_prob = L.n_prob;
  P.D = L.D;
  P.kboxes = L.D / 64;
  P.nblk = L.D / 256;
  P.inv_B = 1.0f / (float)L.rows_global;
  P.ln_B = logf((float)L.rows_global);
  P.gate = gate;
  if (fwd1) P.yflags = L.yflags;
  if (bwd) {
    P.aflags = L.aflags;
    // (any rotation covers every tile once; this one starts every unit on the local rank block)
    if (L.relay || L.aflags.flags) P.rot = (int)((L.row_offset / kBN) % (P.n_tiles > 0 ? P.n_tiles : 1));
  }
  if ((fwd1 || bwd) && L.relay) {
    VPA_CHECK_ARG(L.n_prob <= 2, "pair kernels: the peer-memory relay serves one InfoNCE pair per launch");
    P.relay = *L.relay;
  }
  for (int q = 0; q < L.n_prob; ++q) {
    const PairProblem& s = L.p[q];
    Problem& d = P.p[q];
    if (int e = map_id(s.x, L.rows_local, Cfg<MODE>::RC, &d.mapx)) return e;
    if (int e = map_id(s.y, L.rows_global, 128, &d.mapy)) return e;
    d.n_x = (int)L.rows_local;
    d.n_y = (int)L.rows_global;
    d.diag_offset = (int)L.row_offset;
    d.lse_x = s.lse_x; d.lse_y = s.lse_y;
    d.out = s.out; d.dscale = s.dscale;
    d.logit_scale = s.logit_scale; d.scale_cap = s.scale_cap; d.scale = s.scale;
    d.colpart = s.colpart;
  }
  const SmemLayout SL = smem_layout<MODE>(P.kboxes);
  if (SL.total > kSmemLimit) return set_error(VPA_E_UNSUPPORTED, "pair kernel needs %u bytes of shared memory", SL.total);
  static_assert(kRelaySmemBytes + 1024 <= kSmemLimit, "the relay ring must fit into the forward kernel's shared memory");
  const uint32_t dyn_smem = kSmemLimit;      // the layout plus whatever pad aligns the dynamic base to 1024 B
  static SmemAttrCache attr_cache[3];
  if (int e = ensure_dynamic_smem(attr_cache[MODE], pair_kernel<MODE>, (int)kSmemLimit)) return e;
  dim3 grid(L.n_prob * 2 * P.pairs_per_problem + P.relay.n_ctas), block(MODE == MODE_FWD ? kThreadsFwd : kThreadsBwd);
  const int kind = bwd ? PROF_BWD_SWEEP : (fwd1 ? PROF_FWD_SWEEP : (gate == 2 ? PROF_FWD_GENERAL : PROF_FWD_SWEEP));
  prof_begin(kind, st);
  VPA_CUDA(launch_kernel(pair_kernel<MODE>, grid, block, dyn_smem, st, M, P));
  prof_end(kind, st);
  VPA_LAUNCH_CHECK(bwd ? "pair_kernel<BWD>" : (fwd1 ? "pair_kernel<FWD1>" : "pair_kernel<FWD>"));
  return 0;
}

}  // namespace pr

int pair_launch_fwd1(const PairLaunch& L, const SweepPlan& plan, cudaStream_t st) { return pr::launch<pr::MODE_FWD1>(L, plan, 1, st); }
int pair_launch_fwd(const PairLaunch& L, const SweepPlan& plan, int gate, cudaStream_t st) { return pr::launch<pr::MODE_FWD>(L, plan, gate, st); }
int pair_launch_bwd(const PairLaunch& L, const SweepPlan& plan, cudaStream_t st) { return pr::launch<pr::MODE_BWD>(L, plan, 0, st); }

// One pair (the plain CELossHead step): problem 0 sweeps the local x1 rows against all x2 rows, problem 1 the reverse.
static PairLaunch single_pair(const SweepArgs& a) {
  PairLaunch L{};
  L.rows_local = a.rows_local; L.rows_global = a.rows_global; L.row_offset = a.row_offset; L.D = a.D;
  L.yflags = a.yflags; L.aflags = a.aflags; L.relay = a.relay;
  for (int p = 0; p < 2; ++p) {
    L.p[p].x = a.x[p]; L.p[p].y = a.y[p];
    L.p[p].lse_x = a.lse_x[p]; L.p[p].lse_y = a.lse_y[p];
    L.p[p].logit_scale = a.logit_scale; L.p[p].scale_cap = a.scale_cap; L.p[p].scale = a.scale;
  }
  return L;
}

// Forward.  In the fast configuration the single-pass kernel (valid while s*log2e <= 62) and the exact two-problem
// kernel are both enqueued; each checks the DEVICE value of the temperature and returns at once when it is not its regime.
// which: 0 = exact kernel unconditionally; 1 = single-pass kernel (reads x[0] = local x1 rows and y[0] = all x2 rows
// only), gated on s*log2e <= 62; 2 = exact kernel gated on the complementary regime.
int pair_infonce_fwd(const SweepArgs& a, const Workspace& ws, const SweepPlan& plan, int which, cudaStream_t st) {
  PairLaunch L = single_pair(a);
  if (which == 1) {      // one problem: the local x1 rows against all x2 rows yield both statistics
    L.n_prob = 1;
    L.p[0].out = ws.fwd_part;
    L.p[0].colpart = ws.colpart;
    return pair_launch_fwd1(L, plan, st);
  }
  L.n_prob = 2;
  for (int p = 0; p < 2; ++p) L.p[p].out = ws.fwd_part + (int64_t)p * plan.fwd_chunks * a.rows_local * 2;
  return pair_launch_fwd(L, plan, which == 2 ? 2 : 0, st);
}
int pair_infonce_bwd(const SweepArgs& a, const Workspace& ws, const SweepPlan& plan, cudaStream_t st) {
  PairLaunch L = single_pair(a);
  L.n_prob = 2;
  for (int p = 0; p < 2; ++p) {
    L.p[p].out = ws.bwd_part + (int64_t)p * plan.bwd_chunks * a.rows_local * a.D;
    L.p[p].dscale = p == 0 ? ws.dscale_part : nullptr;
  }
  return pair_launch_bwd(L, plan, st);
}
float pair_fast_s2_limit() { return pr::kFastS2Limit; }

}  // namespace vpa
