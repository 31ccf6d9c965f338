// Encoder tail fused up to the InfoNCE operand: the last three ops of the reference's towers
//     x = ln_post(x[:, 0, :]);  x = x @ proj;  x = x / x.norm(dim=-1, keepdim=True)
// (cvap/module/val.py:288-290 ViTPostEncoder, :143-146 GPTPostEncoder; the normalisation in clip_head.py:117-118 /
// audio_head.py:209-210) produce the rows the loss head contracts.  Here:
//   ln_cast_kernel     CLS rows (any row stride: `hidden[:, 0, :]` is read in place) -> LayerNorm in fp32 (CLIP's LayerNorm
//                      computes in fp32) -> bf16 rows, + mean / rstd for the backward.  One warp per row, HBM-bound.
//   proj_norm_pair_kernel   Y = LN(x) . proj on the tensor cores.  A CTA pair (tcgen05 cta_group::2) owns 256 rows and ALL
//                      N <= 512 output columns: each CTA's accumulator (128 lanes x N fp32) fills its tensor memory, so the L2
//                      norm of a row is a reduction over the thread's own TMEM lane and the epilogue normalises in place: it
//                      writes the bf16 operand rows (what the sweeps' TMA maps read), 1/||y||, and -- for training -- the
//                      un-normalised fp32 features the backward's normalisation Jacobian needs.  LN(x) / proj k-slices stream
//                      through a 4-stage TMA ring (48 KB per stage and CTA), MMAs M256 N256 K16, fp32 accumulate.
// The projected features never make a round trip through HBM between the GEMM and the normalise + cast.
// Measured (B200, 32768 x 768 -> 512, profiles/r02_encoder_tail.md): LayerNorm 28.5 us (HBM-bound, 5.3 TB/s), projection 40 us,
// 77 us for the pair against 654 us (fp32) / 261 us (bf16 autocast) of the same span in PyTorch eager.  A/B history: single-CTA
// kernel (cta_group::1, 2 stages of 80 KB) 46 us; direct per-thread row stores instead of the transposed epilogue 61-63 us.
#include <cuda.h>

#include <cstdio>

#include "common.cuh"

namespace vpa {
namespace tail {

constexpr int kLnWarps = 8;
constexpr int kLnMaxVec = 8;                  // float4 chunks per lane -> width <= 1024
constexpr int kBM = 128, kBoxK = 64;
constexpr int kBoxBytes = kBM * kBoxK * 2;    // 16 KB: one [128 rows][64 elems] bf16 TMA box
constexpr int kThreads = 192;                 // warp 0: TMA, warp 1: MMA issue + TMEM alloc, warps 2..5: epilogue
constexpr uint32_t kSmemLimit = 232448;

// ---------------------------------------------------------------- LayerNorm + cast
template <int DTYPE>
__global__ void __launch_bounds__(kLnWarps * 32)
ln_cast_kernel(const void* __restrict__ x, int64_t rows, int W, int64_t ld, const float* __restrict__ gamma,
               const float* __restrict__ beta, float eps, __nv_bfloat16* __restrict__ y, float* __restrict__ mean_out,
               float* __restrict__ rstd_out) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * kLnWarps + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nvec = W >> 2;
  float4 v[kLnMaxVec];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < kLnMaxVec; ++k) {
    const int c = lane + 32 * k;
    v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < nvec) {
      v[k] = load4<DTYPE>(x, row * ld + 4 * c);
      s += (v[k].x + v[k].y) + (v[k].z + v[k].w);
    }
  }
  const float mean = warp_sum(s) / (float)W;
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < kLnMaxVec; ++k) {
    const int c = lane + 32 * k;
    if (c < nvec) {
      const float a = v[k].x - mean, b = v[k].y - mean, cc = v[k].z - mean, d = v[k].w - mean;
      q += a * a + b * b + cc * cc + d * d;
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)W + eps);      // biased variance, as torch.layer_norm
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
#pragma unroll
  for (int k = 0; k < kLnMaxVec; ++k) {
    const int c = lane + 32 * k;
    if (c < nvec) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c), b = __ldg(reinterpret_cast<const float4*>(beta) + c);
      float4 o;
      o.x = (v[k].x - mean) * rstd * g.x + b.x;
      o.y = (v[k].y - mean) * rstd * g.y + b.y;
      o.z = (v[k].z - mean) * rstd * g.z + b.z;
      o.w = (v[k].w - mean) * rstd * g.w + b.w;
      store4<VPA_BF16>(y, row * (int64_t)W + 4 * c, o);
    }
  }
}

// ---------------------------------------------------------------- PTX wrappers (the idioms of infonce_tc.cu)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  return done != 0;
}
__device__ __noinline__ void mbar_timeout(uint32_t bar, uint32_t parity) {
  printf("vipant_b200(tail): mbarrier wait timeout (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
  asm volatile("trap;");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try(bar, parity)) return;
  uint32_t spins = 0;
  while (!mbar_try(bar, parity))
    if (++spins > (1u << 22)) mbar_timeout(bar, parity);
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, %1;\n\t"
      "@px mov.s32 %0, 1;\n\t}"
      : "+r"(pred) : "r"(0xffffffffu));
  return pred != 0;
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// K-major SW128 shared-memory matrix descriptor and the kind::f16 instruction descriptor: see infonce_tc.cu
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) /*D=f32*/ | (1u << 7) /*A=bf16*/ | (1u << 10) /*B=bf16*/ | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct ProjParams {
  int64_t rows;
  int kboxes, nblk;            // width / 64; N / 128
  __nv_bfloat16* a_out;        // (rows, N) normalised bf16 operand rows
  float* y_out;                // (rows, N) un-normalised fp32 features (nullptr: not wanted)
  float* inv_out;              // (rows,) 1 / ||y||
};

// Epilogue of one warp: thread = row (its TMEM lane holds all N columns of that row).  Pass 1: sum of squares (+ the raw fp32
// features for the backward); pass 2: normalise, cast, store the bf16 operand row.  A thread's 32 columns are 128 contiguous
// bytes of ITS row, so storing them directly makes every warp store touch 32 lines with half a sector each (measured: twice the
// sectors into L2, the epilogue 2/3 of the kernel).  They are transposed through a 4 KB shared-memory tile per warp (the operand
// ring is idle by then; 16-byte chunks XOR-swizzled by row: conflict-free both ways) and leave as 4 full lines (fp32) / 8 rows x
// 64 bytes (bf16) per warp store.
__device__ __forceinline__ void epilogue_rows(uint32_t tmem_base, int row0, int sub, int lane, int N, const ProjParams& P,
                                              uint8_t* stage_smem) {
  const int64_t rbase = row0 + sub * 32;
  const bool row_ok = rbase + lane < P.rows;
  const uint32_t t_lane = tmem_base + ((uint32_t)(sub * 32) << 16);
  uint4* s4 = reinterpret_cast<uint4*>(stage_smem + sub * 4096);
  float ss = 0.f;
#pragma unroll 1
  for (int c = 0; c < N / 32; ++c) {
    uint32_t r[32];
    tmem_ld32(t_lane + c * 32, r);
    tmem_ld_wait();
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
    for (int e = 0; e < 32; e += 4) {
      const float v0 = __uint_as_float(r[e]), v1 = __uint_as_float(r[e + 1]), v2 = __uint_as_float(r[e + 2]), v3 = __uint_as_float(r[e + 3]);
      a0 = fmaf(v0, v0, a0); a1 = fmaf(v1, v1, a1); a2 = fmaf(v2, v2, a2); a3 = fmaf(v3, v3, a3);
    }
    ss += (a0 + a1) + (a2 + a3);
    if (P.y_out) {
#pragma unroll
      for (int q = 0; q < 8; ++q) s4[lane * 8 + (q ^ (lane & 7))] = make_uint4(r[q * 4], r[q * 4 + 1], r[q * 4 + 2], r[q * 4 + 3]);
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int rr = (lane >> 3) + 4 * i, ch = lane & 7;
        const uint4 v = s4[rr * 8 + (ch ^ (rr & 7))];
        if (rbase + rr < P.rows) *reinterpret_cast<uint4*>(P.y_out + (rbase + rr) * N + c * 32 + ch * 4) = v;
      }
      __syncwarp();
    }
  }
  const float inv = 1.0f / sqrtf(ss);
  if (row_ok && P.inv_out) P.inv_out[rbase + lane] = inv;
#pragma unroll 1
  for (int c = 0; c < N / 32; ++c) {
    uint32_t r[32];
    tmem_ld32(t_lane + c * 32, r);
    tmem_ld_wait();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint32_t w[4];
#pragma unroll
      for (int h = 0; h < 4; ++h)
        w[h] = pack_bf16x2(__uint_as_float(r[q * 8 + 2 * h]) * inv, __uint_as_float(r[q * 8 + 2 * h + 1]) * inv);
      s4[lane * 4 + (q ^ ((lane >> 1) & 3))] = make_uint4(w[0], w[1], w[2], w[3]);
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int rr = (lane >> 2) + 8 * i, ch = lane & 3;
      const uint4 v = s4[rr * 4 + (ch ^ ((rr >> 1) & 3))];
      if (rbase + rr < P.rows) *reinterpret_cast<uint4*>(P.a_out + (rbase + rr) * N + c * 32 + ch * 8) = v;
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------- projection + normalise on a CTA pair (tcgen05 cta_group::2)
// Two CTAs (one TPC) own 256 rows.  Every MMA is M = 256 (128 rows per CTA) x N = 256, and each CTA stages only HALF of the
// projection's rows of a 256-column block, so a stage is 48 KB and the ring is 4 deep.  Protocol as in infonce_pair.cu: both
// producers account their bytes on the LEADER's full barrier, the leader issues, commits are multicast to both CTAs.
constexpr int kPairStages = 4;
enum { Q_DONE = 0, Q_FULL = 1, Q_EMPTY = 1 + kPairStages, Q_COUNT = 1 + 2 * kPairStages };

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2sm(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t leader_bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(leader_bar) : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma2_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
proj_norm_pair_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, const ProjParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sptr = smem_raw + (sbase - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();            // 0 = leader (issues the MMAs), 1 = peer
  const bool is_leader = crank == 0;
  const int N = P.nblk * 128;
  const int nb2 = N / 256;                             // 256-column blocks
  const uint32_t stage_bytes = (uint32_t)(1 + nb2) * kBoxBytes;
  const uint32_t bar0 = sbase + kPairStages * stage_bytes;
  auto bar = [&](int i) { return bar0 + 8u * i; };
  auto lbar = [&](int i) { return mapa(bar0 + 8u * i, 0); };     // the leader's copy (cluster address)
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sptr + kPairStages * stage_bytes + 8 * Q_COUNT);
  const uint32_t tmem_cols = N <= 256 ? 256u : 512u;
  if (threadIdx.x == 0) {
    mbar_init(bar(Q_DONE), 1);
    for (int s = 0; s < kPairStages; ++s) { mbar_init(bar(Q_FULL + s), 1); mbar_init(bar(Q_EMPTY + s), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    prefetch_tmap(&map_a);
    prefetch_tmap(&map_w);
  }
  if (warp == 1) tmem_alloc2(smem_u32(const_cast<uint32_t*>(tmem_slot)), tmem_cols);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int row0 = (int)(blockIdx.x >> 1) * 2 * kBM + (int)crank * kBM;

  if (warp == 0) {
    // ---- TMA producer (both CTAs): own 128 rows of LN(x), and this CTA's 128 of every 256 projection rows
    const bool elected = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = 0; kb < P.kboxes; ++kb) {
      mbar_wait(bar(Q_EMPTY + stage), phase ^ 1);
      if (elected) {
        if (is_leader) mbar_arrive_expect_tx(bar(Q_FULL + stage), 2 * stage_bytes);
        const uint32_t fb = lbar(Q_FULL + stage);
        const uint32_t dst = sbase + stage * stage_bytes;
        tma_load_2sm(dst, &map_a, kb * kBoxK, row0, fb);
        for (int nb = 0; nb < nb2; ++nb) tma_load_2sm(dst + (1 + nb) * kBoxBytes, &map_w, kb * kBoxK, nb * 256 + (int)crank * 128, fb);
      }
      if (++stage == kPairStages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    // ---- MMA issuer (leader CTA only)
    if (is_leader) {
      const bool elected = elect_one();
      constexpr uint32_t idesc = make_idesc(256, 256);
      const uint64_t dk = make_desc(sbase, 16, 1024);
      constexpr uint32_t box_u = kBoxBytes >> 4;
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < P.kboxes; ++kb) {
        mbar_wait(bar(Q_FULL + stage), phase);
        tc_fence_after();
        if (elected) {
          const uint32_t st_u = (stage * stage_bytes) >> 4;
          const uint64_t da0 = dk + (uint64_t)st_u;
          for (int nb = 0; nb < nb2; ++nb) {
            const uint64_t db0 = dk + (uint64_t)(st_u + (1 + nb) * box_u);
#pragma unroll
            for (int k = 0; k < kBoxK / 16; ++k)
              umma2_f16(tmem_base + nb * 256, da0 + (uint64_t)(k * 2), db0 + (uint64_t)(k * 2), idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma2_commit(bar(Q_EMPTY + stage));
          if (kb == P.kboxes - 1) umma2_commit(bar(Q_DONE));
        }
        __syncwarp();
        if (++stage == kPairStages) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    mbar_wait(bar(Q_DONE), 0);
    tc_fence_after();
    epilogue_rows(tmem_base, row0, warp & 3, lane, N, P, sptr);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                    // neither CTA frees tensor memory / exits while the pair may still touch it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, tmem_cols);
  }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int make_map(CUtensorMap* m, const void* base, int64_t rows, int cols) {
  static EncodeTiledFn enc = nullptr;
  if (!enc) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      enc = reinterpret_cast<EncodeTiledFn>(p);
  }
  if (!enc) return set_error(VPA_E_NO_DEVICE, "cuTensorMapEncodeTiled entry point unavailable");
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) { cudaFree(nullptr); ctx_bound = true; }      // the encode is a driver call: needs a current context
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)kBoxK, (cuuint32_t)kBM};
  cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(VPA_E_INVALID, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return 0;
}

}  // namespace tail

int encoder_tail_launch(const void* x, int in_dtype, int64_t rows, int width, int64_t ld, const float* gamma, const float* beta,
                        float eps, const void* proj_t_bf16, int N, void* ln_bf16, float* mean, float* rstd, void* a_bf16,
                        float* y_f32, float* inv_norm, cudaStream_t st) {
  using namespace tail;
  VPA_CHECK_ARG(x && gamma && beta && proj_t_bf16 && ln_bf16 && a_bf16, "encoder_tail: null pointer");
  VPA_CHECK_ARG(rows >= 0 && ld >= width && ld % 4 == 0, "encoder_tail: bad rows / leading dimension");
  VPA_CHECK_ARG(in_dtype == VPA_F32 || in_dtype == VPA_BF16 || in_dtype == VPA_F16, "encoder_tail: bad dtype %d", in_dtype);
  if (width % 64 != 0 || width < 64 || width > 128 * kLnMaxVec || (N != 256 && N != 512))
    return set_error(VPA_E_UNSUPPORTED, "encoder_tail: width %% 64 == 0, 64 <= width <= %d, N in {256, 512} (width=%d N=%d)",
                     128 * kLnMaxVec, width, N);
  VPA_CHECK_ARG((reinterpret_cast<uintptr_t>(x) % (in_dtype == VPA_F32 ? 16 : 8)) == 0 &&
                ((reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta)) % 16) == 0, "encoder_tail: misaligned pointer");
  if (rows == 0) return 0;
  {
    dim3 grid((unsigned)((rows + kLnWarps - 1) / kLnWarps)), block(kLnWarps * 32);
    auto* y = reinterpret_cast<__nv_bfloat16*>(ln_bf16);
    if (in_dtype == VPA_F32) ln_cast_kernel<VPA_F32><<<grid, block, 0, st>>>(x, rows, width, ld, gamma, beta, eps, y, mean, rstd);
    else if (in_dtype == VPA_BF16) ln_cast_kernel<VPA_BF16><<<grid, block, 0, st>>>(x, rows, width, ld, gamma, beta, eps, y, mean, rstd);
    else ln_cast_kernel<VPA_F16><<<grid, block, 0, st>>>(x, rows, width, ld, gamma, beta, eps, y, mean, rstd);
    VPA_LAUNCH_CHECK("ln_cast_kernel");
  }
  CUtensorMap map_a, map_w;
  if (int e = make_map(&map_a, ln_bf16, rows, width)) return e;
  if (int e = make_map(&map_w, proj_t_bf16, N, width)) return e;
  ProjParams P{};
  P.rows = rows; P.kboxes = width / 64; P.nblk = N / 128;
  P.a_out = reinterpret_cast<__nv_bfloat16*>(a_bf16); P.y_out = y_f32; P.inv_out = inv_norm;
  const uint32_t smem = kPairStages * (1 + N / 256) * kBoxBytes + 8 * Q_COUNT + 16 + 1024;
  if (smem > kSmemLimit) return set_error(VPA_E_UNSUPPORTED, "encoder_tail: %u bytes of shared memory", smem);
  static SmemAttrCache attr_cache;
  if (int e = ensure_dynamic_smem(attr_cache, proj_norm_pair_kernel, (int)kSmemLimit)) return e;
  proj_norm_pair_kernel<<<2 * (unsigned)((rows + 2 * kBM - 1) / (2 * kBM)), kThreads, smem, st>>>(map_a, map_w, P);
  VPA_LAUNCH_CHECK("proj_norm_pair_kernel");
  return 0;
}

}  // namespace vpa
