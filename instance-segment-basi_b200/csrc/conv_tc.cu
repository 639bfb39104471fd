// tcgen05 / TMEM / TMA implicit-GEMM convolutions for sm_100a (bf16 operands, fp32 accumulate in TMEM).
//
//   fprop / dgrad :  D[128 pixels][BN ch] = sum_{tap, 64-ch chunk} A_tap[128 px][64] * B_tap[BN][64]^T
//       A_tap is a 4-D TMA box (64 ch, TW, TH, TN) of the NHWC activation shifted by the filter tap;
//       out-of-image coordinates are zero-filled by TMA, which *is* the convolution padding.
//       B_tap is a 3-D TMA box of the bf16 weights laid out [tap][N][K] (K contiguous).
//       Both land in shared memory in the canonical K-major SWIZZLE_128B layout that tcgen05.mma reads.
//   wgrad        :  dW_tap[128 ci][BN co] += sum_{pixel tiles} X_tap[px][ci]^T * dY[px][co]
//       the same boxes, consumed as MN-major operands (pixels are the GEMM K dimension), split over
//       pixel tiles across CTAs, fp32 red.global.add epilogue into the HWIO gradient.
//
// Persistent CTAs, warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane),
// warp 2 = TMEM allocator, warps 4..7 = epilogue (TMEM -> registers -> global).  Two TMEM accumulators
// so the epilogue of tile i overlaps the main loop of tile i+1.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "ptx.cuh"

namespace basi {
namespace tc {

constexpr int BM = 128;            // pixels per tile (UMMA M)
constexpr int KC = 64;             // channels per k-chunk: 64 bf16 = 128 B = one swizzle row
constexpr int A_BYTES = BM * 128;  // 16 KB
constexpr int NTHREADS = 256;        // wgrad: 4 control warps + 4 epilogue warps
constexpr int NTHREADS_CONV = 384;   // fprop/dgrad: 4 control warps + 8 epilogue warps (2 per TMEM lane quarter)

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// CTA-pair (cta_group::2) variants: both CTAs of a cluster load into their own shared memory and signal the LEADER's
// mbarrier (an address in the shared::cluster window); the leader issues one M = 256 MMA that reads both halves.
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void tma_load_4d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                                 int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, "
      "%6}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                                 int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, "
      "%5}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
      "h"(mask)
      : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
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
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(map),
               "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2,
                                                  int c3) {
  asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::
                   "l"(map),
               "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// UMMA shared-memory descriptor, SWIZZLE_128B (layout_type 2), descriptor version 1 (sm_100).
//   K-major : LBO unused (1), SBO = 1024 B (8 rows of 128 B)
//   MN-major: LBO = distance between 64-element MN blocks, SBO = 1024 B (8 k-rows of 128 B)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// instruction descriptor: D=f32, A=B=bf16, M=128 (256 for a CTA pair)
__host__ __device__ constexpr uint32_t make_idesc(int n, int a_mn_major, int b_mn_major, int m = BM) {
  // (bits 7-9 / 10-12: A / B format of kind::f16 -- 0 = fp16, 1 = bf16)
  constexpr uint32_t fmt = BASI_H16_FP16 ? 0u : 1u;
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

struct ConvParams {
  // tiling of the destination pixels: tile = TN images x TH rows x TW cols = 128 pixels
  int TW, TH, TN;
  int tiles_w, tiles_h, tiles_n;   // number of tiles per axis
  int m_tiles, n_tiles;
  int N, H, W;                     // destination tensor extents
  int ldd;                         // destination pixel stride (elements)
  int Cdst;                        // destination channels
  int taps, kw;
  int k_chunks;                    // source channels / 64 (split mode: chunks of all passes, see a_wrap0)
  // split-operand (fp32-grade) mode: the source tensor holds [hi | mid | lo] bf16 parts of an fp32 tensor (3C
  // channels) and the main loop runs three passes over it against [w_hi | w_mid | w_lo]-style weight columns: the
  // activation chunk of k-chunk kc is kc, kc - a_wrap0 or kc - a_wrap1 (the weight chunk is always kc).  Plain bf16
  // mode: both = INT_MAX.
  int a_wrap0, a_wrap1;
  // split mode, accumulation groups: tcgen05 adds into the fp32 accumulator with truncation (measured: ~0.4 ulp of
  // bias per MMA, 2e-5 relative after the 864 MMAs of a 3x3x256 split convolution), so the k-steps whose products
  // have full magnitude (hi x hi: the first big_chunks chunks of every tap) run LAST, in groups of group_steps
  // k-steps that each start a fresh TMEM accumulator; the epilogue warps add every finished group into fp32
  // registers (round to nearest).  All small k-steps (cross terms, 2^-8 .. 2^-16 of the result) form the first group.
  int big_chunks, group_steps;
  // fused BN apply (bf16 mode): every accumulator of this CTA stays in TMEM (fresh columns per work item), the
  // statistics are flushed, the grid meets at a barrier on bn_counter (cooperative launch), every CTA derives
  // scale / shift of its channel tile from the final sums and a second epilogue pass writes
  // relu(scale * acc + shift) -- normalised from the fp32 accumulators -- through mapD2.  Removes the separate
  // bn_apply launch, one read pass over the conv output and one rounding per layer.
  int fuse_apply, apply_relu;
  const float* bias;     // fprop: optional per-channel bias added in the epilogue (slim vgg_16 convolutions, variant B)
  int bias_relu;         // ... followed by ReLU
  int keep_acc;      // fuse_apply || fuse_bwd: accumulators stay in TMEM (one column block and one barrier per work item)
  // fused BN backward (dgrad plans): the accumulator is dA, the gradient wrt the activated output of the layer that
  // produced this plan's destination.  Pass 1 loads the matching tile of that layer's raw conv output x (mapD2),
  // masks (ReLU recomputed from x) and accumulates sum(g), sum(g * (x - mean)) per channel; grid barrier; pass 2
  // writes dx = A g - A S1/N - (x - mean) A istd S2/N (A = gamma * istd) to the destination and publishes dgamma /
  // dbeta.  dA itself never goes to memory, and the cooperative bn_bwd launch of that layer disappears.
  int fuse_bwd, bwd_relu;
  int xbuf_off, coef_off, xbar_off;   // byte offsets from the output staging area: x tile buffer, coefficient table, barrier
  const float* bwd_bnp;      // [mean | istd | gamma*istd | beta]
  double* bwd_sums;          // [BASI_BN_REPLICAS][2C], zero at the start of a step
  float* bwd_dgamma;
  float* bwd_dbeta;
  double bwd_count;
  unsigned int* bwd_counter;
  // source coordinate of tap (r,s) for destination pixel p: p*1 + off0 + r*step (fprop: off0=-pad, step=dil;
  // dgrad: off0=+pad, step=-dil)
  int off_h, off_w, step;
  int accumulate;
  int stages;
  int ring_bytes; // bytes of the operand rings in front of the barriers
  // halo mode (3x3, stride 1): per pixel tile (8 cols x 16 rows x 1 image) and 64-channel chunk ONE activation box
  // of 16 x (16 + 2*dil) pixels is loaded and all nine taps read it through row-shifted descriptors
  // (start = halo + (dr * 16 + dc) * 128 B, SBO = 2048 B); stages = nh halo slots + nb weight slots
  int halo, dil, halo_bytes, nh, nb;
  int bres;       // halo mode, one 64-channel chunk, one channel tile: the nine weight boxes stay resident in shared
                  // memory for the whole CTA (loaded once), a tile only pulls its activation halo
  int out_bufs;   // 1 or 2 output staging buffers (2: the TMA store of tile i overlaps the epilogue of tile i+1)
  int debug;      // timing experiments only: 1 = skip the statistics atomics, 2 = skip the column sums too
  // optional fused batch-norm statistics of the produced tensor (fprop): per-channel sum / sum of squares of the
  // bf16-rounded outputs, double atomics per tile, last CTA finalizes bnp = [mean | istd | gamma*istd | beta]
  double* bn_sums;
  const float* bn_gamma;
  const float* bn_beta;
  float* bn_bnp;
  unsigned int* bn_counter;
  double bn_count;
  float bn_eps;
};

// BASI_TC_DEBUG_STATS=30: CTA 0 accumulates the cycles its roles spend waiting (see basi_tc_conv_run)
__device__ unsigned long long g_tc_dbg[16];
__device__ __forceinline__ long long dbg_clock() { return clock64(); }

// transpose-reduce across the 32 lanes of a warp: on return v[0] of lane l is the sum over all lanes of their v[l]
__device__ __forceinline__ void warp_column_sums(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int j = 0; j < s; ++j) {
      const float send = up ? v[j] : v[j + s];
      const float keep = up ? v[j + s] : v[j];
      v[j] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
}

template <int BN, int MT = 1, int CL = 1>
struct SmemLayout {
  static constexpr int B_BYTES = BN * 128 / CL;   // a CTA pair holds half of the weight box in each CTA
  static constexpr int STAGE = MT * A_BYTES + B_BYTES;
};

// ------------------------------------------------------------------------------------------------------
// fprop / dgrad
// ------------------------------------------------------------------------------------------------------
// VAR: 0 = the production kernel, 1 = halo mode (+ role timers), 2 = production kernel with the role timers,
// 3 = dgrad + fused batch-norm backward (opt-in, see ConvParams::fuse_bwd), 4 = fprop + fused batch-norm apply
// (compile-time so that the experimental paths cost the production kernel nothing: as run-time branches they
// measured 1.8 % of the training step)
// k-step ks of a tile -> (filter tap, k-chunk).  Plain mode: tap-major.  Split mode: small k-steps first, then the
// full-magnitude ones (see ConvParams::big_chunks).
__device__ __forceinline__ void kstep_coords(const ConvParams& p, int ks, int& tap, int& kc) {
  if (p.big_chunks == 0) {
    tap = ks / p.k_chunks;
    kc = ks - tap * p.k_chunks;
    return;
  }
  const int small_per_tap = p.k_chunks - p.big_chunks, n_small = p.taps * small_per_tap;
  if (ks < n_small) {
    tap = ks / small_per_tap;
    kc = p.big_chunks + (ks - tap * small_per_tap);
  } else {
    const int j = ks - n_small;
    tap = j / p.big_chunks;
    kc = j - tap * p.big_chunks;
  }
}
// end (exclusive) of the accumulation group that starts at k-step gs
__device__ __forceinline__ int group_end(const ConvParams& p, int gs, int ksteps) {
  if (p.big_chunks == 0) return ksteps;
  const int n_small = p.taps * (p.k_chunks - p.big_chunks);
  if (gs < n_small) return n_small;
  return min(ksteps, gs + p.group_steps);
}

// OUT32: the destination tensor is float32 (split-operand mode): fp32 staging boxes of [128 px][32 ch], TMA store /
// reduce-add of float32, statistics from the fp32 values.
template <int BN, int CL, int MT, int VAR, bool OUT32 = false>
__global__ void __launch_bounds__(NTHREADS_CONV, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
               const __grid_constant__ CUtensorMap mapD, const __grid_constant__ CUtensorMap mapD2, const ConvParams p) {
  // CL == 2: a CTA pair (tcgen05 cta_group::2).  The two CTAs own adjacent pixel tiles of the same channel tile;
  // each loads its own activation box and HALF of the weight box into its own shared memory, the leader (rank 0)
  // issues ONE M = 256 MMA per K slice that reads both CTAs' shared memory and writes both CTAs' TMEM.  The main
  // loop is bound by shared-memory bandwidth (every operand byte is written once by TMA and read once by the MMA:
  // 64 KB per 128x128x64 k-step at 128 B/clk = 0.27 us, measured 0.35 us); a pair moves 48 KB per CTA instead.
  // MT == 2: one work item is TWO adjacent pixel tiles against the same weight tile: per k-step the CTA loads
  // 2 x 16 KB of activations + one weight box and issues 2 x 4 MMAs into two TMEM accumulators.  The main loop is
  // bound by L2 -> SM bytes (148 SMs x 32 KB per 0.37 us k-step = the ~12 TB/s L2 cap); sharing the weight box cuts
  // the bytes per flop by 25 %.
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr bool HALO = VAR == 1, TIMERS = (VAR == 1 || VAR == 2);
  static_assert(!HALO || (CL == 1 && MT == 1), "halo mode: one tile per CTA, no pairs");
  constexpr int STAGE = SmemLayout<BN, MT, CL>::STAGE;
  static_assert(CL == 1 || MT == 1, "pairs take one pixel tile per CTA");
  constexpr uint32_t TMEM_COLS2 = (2 * MT * BN < 32) ? 32 : 2 * MT * BN;
  static_assert(2 * MT * BN <= 512, "TMEM has 512 columns");
  const uint32_t TMEM_COLS = p.keep_acc ? 512u : TMEM_COLS2;   // fused apply: one accumulator per work item
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.ring_bytes);
  // bars: full[stages], empty[stages], tmem_full[NTF], tmem_empty[2], then the TMEM base slot.  (tmem_full: two are
  // used in turn by the double-buffered accumulators; the fused-apply mode keeps every accumulator of the CTA and
  // uses one barrier per work item -- the MMA warp may run several items ahead of the epilogue, and an mbarrier
  // that completes two phases before it is waited on can no longer be distinguished.)
  constexpr int NTF = 16;
  const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * p.stages;
  const uint32_t tfull0 = empty0 + 8 * p.stages, tempty0 = tfull0 + 8 * NTF;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * p.stages + NTF + 2);
  float* stat_s = reinterpret_cast<float*>(bars + 2 * p.stages + NTF + 4);   // [row groups][2*BN] = 1024 floats
  // output staging for the TMA store: NBOX boxes of [128 px][64 ch] bf16, SWIZZLE_128B, 1024-B aligned
  constexpr int NBOX = OUT32 ? BN / 32 : (BN + 63) / 64;
  constexpr int BOXC = OUT32 ? 32 : 64;      // channels per output box (128-byte rows)
  static_assert(!OUT32 || (CL == 1 && VAR == 0 && BN <= 128), "fp32 output: single CTA, production variant, BN <= 128");
  uint8_t* stage_out = smem + (size_t)p.ring_bytes + 1024 + 1024 * sizeof(float);
  stage_out = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(stage_out) + 1023) & ~(uintptr_t)1023);
  volatile uint32_t* s_is_last = tmem_slot + 1;   // (no static __shared__: the dynamic window is the full 227 KB)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (p.debug == 10) return;   // timing experiment: launch cost only

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapD) : "memory");
    if (p.keep_acc) asm volatile("prefetch.tensormap [%0];" ::"l"(&mapD2) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(full0 + 8 * i, 1);    // pair: only the leader's is used (it expects the bytes of both CTAs)
      mbar_init(empty0 + 8 * i, 1);   // pair: the leader's commit is multicast to both CTAs
    }
    for (int i = 0; i < (p.keep_acc ? NTF : 2); ++i) mbar_init(tfull0 + 8 * i, 1);
    if (p.fuse_bwd) mbar_init(smem_u32(stage_out) + (uint32_t)p.xbar_off, 1);
    for (int i = 0; i < 2; ++i)
      mbar_init(tempty0 + 8 * i, 8 * CL);   // one arrive per epilogue warp (pair: of both CTAs, on the leader's)
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2 && p.debug != 11) {
    if (CL == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"(TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"(TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CL == 2) cluster_sync_all();   // the peer's barriers are initialised before anything is multicast into them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // PDL: everything above (barrier init, TMEM allocation, descriptor prefetch) overlapped the previous kernel's tail
  pdl_launch_dependents();
  pdl_wait();

  const int crank = CL == 2 ? (int)cluster_ctarank() : 0;
  const int cid = CL == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int ncl = CL == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int m_groups = (p.m_tiles + CL * MT - 1) / (CL * MT);
  // loop space; pixel tile = (group * CL + crank) * MT + sub (may be a phantom past the end)
  const int total_tiles = m_groups * p.n_tiles;
  const int ksteps = p.taps * p.k_chunks;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0 && HALO) {
      // halo mode: per (tile, 64-channel chunk) one activation halo box + nine weight boxes
      constexpr int B_BYTES = BN * 128;
      int hs = 0, bs = 0;
      uint32_t hph = 0, bph = 0;
      if (p.bres && cid < total_tiles) {
        // resident weights: all taps of the (only) channel tile and chunk, once
        mbar_expect_tx(full0 + 8 * p.nh, p.taps * B_BYTES);
        for (int tap = 0; tap < p.taps; ++tap)
          tma_load_3d(smem_u32(smem + (size_t)p.nh * p.halo_bytes + (size_t)tap * B_BYTES), &mapB, full0 + 8 * p.nh, 0, 0,
                      tap);
      }
      for (int t = cid; t < total_tiles; t += ncl) {
        const int nt = t % p.n_tiles, mt = t / p.n_tiles;
        const int tw = mt % p.tiles_w, th = (mt / p.tiles_w) % p.tiles_h, tn = mt / (p.tiles_w * p.tiles_h);
        const int w0 = tw * p.TW, h0 = th * p.TH;
        for (int kc = 0; kc < p.k_chunks; ++kc) {
          mbar_wait(empty0 + 8 * hs, hph ^ 1);
          mbar_expect_tx(full0 + 8 * hs, p.halo_bytes);
          tma_load_4d(smem_u32(smem + (size_t)hs * p.halo_bytes), &mapA, full0 + 8 * hs, kc * KC, w0 - p.dil, h0 - p.dil,
                      tn);
          if (++hs == p.nh) {
            hs = 0;
            hph ^= 1;
          }
          if (p.bres) continue;
          for (int tap = 0; tap < p.taps; ++tap) {
            const int bi = p.nh + bs;
            mbar_wait(empty0 + 8 * bi, bph ^ 1);
            mbar_expect_tx(full0 + 8 * bi, B_BYTES);
            tma_load_3d(smem_u32(smem + (size_t)p.nh * p.halo_bytes + (size_t)bs * B_BYTES), &mapB, full0 + 8 * bi, kc * KC,
                        nt * BN, tap);
            if (++bs == p.nb) {
              bs = 0;
              bph ^= 1;
            }
          }
        }
      }
    } else if (!HALO) {
      // warp-uniform loops, the elected lane issues the copies (see elect_one)
      int stage = 0;
      uint32_t phase = 0;
      for (int t = cid; t < total_tiles; t += ncl) {
        const int nt = t % p.n_tiles, mt0 = ((t / p.n_tiles) * CL + crank) * MT;
        int w0[MT], h0[MT], n0[MT];
#pragma unroll
        for (int sub = 0; sub < MT; ++sub) {
          const int mt = mt0 + sub;
          const int tw = mt % p.tiles_w, th = (mt / p.tiles_w) % p.tiles_h, tn = mt / (p.tiles_w * p.tiles_h);
          w0[sub] = tw * p.TW; h0[sub] = th * p.TH; n0[sub] = tn * p.TN;
        }
        {
          int tap = 0, kc = 0, tr = 0, ts = 0;          // plain mode: tap-major counters, no divisions in the loop
          for (int ks = 0; ks < ksteps; ++ks) {
            if constexpr (OUT32) {
              kstep_coords(p, ks, tap, kc);
              tr = tap / p.kw;
              ts = tap - tr * p.kw;
            }
            const int dh = p.off_h + tr * p.step, dw = p.off_w + ts * p.step;
            mbar_wait(empty0 + 8 * stage, phase ^ 1);
            const uint32_t sa = smem_u32(smem + (size_t)stage * STAGE);
            const int ac = kc >= p.a_wrap1 ? kc - p.a_wrap1 : (kc >= p.a_wrap0 ? kc - p.a_wrap0 : kc);
            if (elect_one()) {
              if (CL == 2) {
                // both CTAs' bytes are counted on the leader's barrier
                const uint32_t fb = mapa_u32(full0 + 8 * stage, 0);
                if (crank == 0) mbar_expect_tx(full0 + 8 * stage, 2 * STAGE);
                tma_load_4d_pair(sa, &mapA, fb, ac * KC, w0[0] + dw, h0[0] + dh, n0[0]);
                tma_load_3d_pair(sa + A_BYTES, &mapB, fb, kc * KC, nt * BN + crank * (BN / 2), tap);
              } else {
                const uint32_t fb = full0 + 8 * stage;
                mbar_expect_tx(fb, STAGE);
#pragma unroll
                for (int sub = 0; sub < MT; ++sub)
                  tma_load_4d(sa + sub * A_BYTES, &mapA, fb, ac * KC, w0[sub] + dw, h0[sub] + dh, n0[sub]);
                tma_load_3d(sa + MT * A_BYTES, &mapB, fb, kc * KC, nt * BN, tap);
              }
            }
            __syncwarp();
            if (++stage == p.stages) {
              stage = 0;
              phase ^= 1;
            }
            if constexpr (!OUT32) {
              if (++kc == p.k_chunks) {
                kc = 0;
                ++tap;
                if (++ts == p.kw) { ts = 0; ++tr; }
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0 && HALO) {
      constexpr uint32_t idesc = make_idesc(BN, 0, 0, BM);
      constexpr int B_BYTES = BN * 128;
      int hs = 0, bs = 0;
      uint32_t hph = 0, bph = 0;
      int it = 0;
      if (p.bres && cid < total_tiles) {
        mbar_wait(full0 + 8 * p.nh, 0);          // the resident weight boxes have landed
        tc_fence_after();
      }
      for (int t = cid; t < total_tiles; t += ncl, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        const bool dbg = TIMERS && p.debug == 30 && blockIdx.x == 0;
        long long m0 = dbg ? dbg_clock() : 0;
        mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1);
        if (dbg) { const long long m1 = dbg_clock(); g_tc_dbg[4] += m1 - m0; m0 = m1; }
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kc = 0; kc < p.k_chunks; ++kc) {
          mbar_wait(full0 + 8 * hs, hph);
          if (dbg) { const long long m1 = dbg_clock(); g_tc_dbg[5] += m1 - m0; m0 = m1; }
          tc_fence_after();
          const uint32_t ha = smem_u32(smem + (size_t)hs * p.halo_bytes);
          if (p.bres) {
            for (int tap = 0; tap < p.taps; ++tap) {
              const int r = tap / p.kw, sx = tap - r * p.kw;
              const int dr = p.off_h + r * p.step + p.dil, dc = p.off_w + sx * p.step + p.dil;
              const uint64_t adesc = make_desc(ha + (uint32_t)(dr * 16 + dc) * 128, 16, 2048);
              const uint64_t bdesc =
                  make_desc(smem_u32(smem + (size_t)p.nh * p.halo_bytes + (size_t)tap * B_BYTES), 16, 1024);
#pragma unroll
              for (int k = 0; k < KC / 16; ++k)
                umma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (tap | k) != 0);
            }
            umma_commit(empty0 + 8 * hs);
            if (++hs == p.nh) {
              hs = 0;
              hph ^= 1;
            }
            continue;
          }
          for (int tap = 0; tap < p.taps; ++tap) {
            const int r = tap / p.kw, sx = tap - r * p.kw;
            int dr = p.off_h + r * p.step + p.dil, dc = p.off_w + sx * p.step + p.dil;
            if (p.debug == 20) dr = dc = 0;          // timing experiment: aligned descriptor starts (wrong results)
            if (p.debug == 21) dc = 0;               // timing experiment: 1024-B aligned starts only
            const int bi = p.nh + bs;
            mbar_wait(full0 + 8 * bi, bph);
            tc_fence_after();
            // rows of the tile's 8-pixel groups: 16 halo pixels (2048 B) apart; the tap only moves the start row
            const uint64_t adesc = make_desc(ha + (uint32_t)(dr * 16 + dc) * 128, 16, 2048);
            const uint64_t bdesc = make_desc(smem_u32(smem + (size_t)p.nh * p.halo_bytes + (size_t)bs * B_BYTES), 16, 1024);
#pragma unroll
            for (int k = 0; k < KC / 16; ++k)
              umma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kc | tap | k) != 0);
            umma_commit(empty0 + 8 * bi);
            if (++bs == p.nb) {
              bs = 0;
              bph ^= 1;
            }
          }
          umma_commit(empty0 + 8 * hs);      // every tap has read the halo
          if (++hs == p.nh) {
            hs = 0;
            hph ^= 1;
          }
        }
        umma_commit(tfull0 + 8 * acc);
      }
    } else if (!HALO && crank == 0) {
      // all 32 lanes run the loops (warp-uniform control flow); the elected lane issues the tcgen05 instructions
      constexpr uint32_t idesc = make_idesc(BN, 0, 0, CL * BM);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;      // accumulation groups issued so far (plain mode: one group per tile)
      for (int t = cid; t < total_tiles; t += ncl)
      for (int gs = 0; gs < ksteps; ++it) {
        const int ge = group_end(p, gs, ksteps);
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        const bool dbg = TIMERS && p.debug == 30 && blockIdx.x == 0 && lane == 0;
        long long m0 = dbg ? dbg_clock() : 0;
        if (!p.keep_acc) mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1);   // epilogue has drained this accumulator
        if (dbg) { const long long m1 = dbg_clock(); g_tc_dbg[4] += m1 - m0; m0 = m1; }
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (p.keep_acc ? it : acc) * (MT * BN);
        const int ks0 = gs;
        for (int ks = ks0; ks < ge; ++ks) {
          if (dbg) m0 = dbg_clock();
          mbar_wait(full0 + 8 * stage, phase);
          if (dbg) g_tc_dbg[5] += dbg_clock() - m0;
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)stage * STAGE);
          const uint64_t bdesc = make_desc(sa + MT * A_BYTES, 16, 1024);
          if (elect_one()) {
#pragma unroll
            for (int sub = 0; sub < MT; ++sub) {
              const uint64_t adesc = make_desc(sa + sub * A_BYTES, 16, 1024);
#pragma unroll
              for (int k = 0; k < KC / 16; ++k) {
                // advance 16 elements (32 B) along K inside the 128-B swizzle row: +2 in 16-B units
                if (CL == 2) umma_f16_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, ((ks - ks0) | k) != 0);
                else umma_f16(d_tmem + sub * BN, adesc + 2 * k, bdesc + 2 * k, idesc, ((ks - ks0) | k) != 0);
              }
            }
            // frees the smem slot (pair: in both CTAs)
            if (CL == 2) umma_commit_pair(empty0 + 8 * stage, (uint16_t)3);
            else umma_commit(empty0 + 8 * stage);
          }
          __syncwarp();
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        // accumulator complete -> epilogue (pair: of both CTAs)
        if (elect_one()) {
          if (CL == 2) umma_commit_pair(tfull0 + 8 * acc, (uint16_t)3);
          else umma_commit(tfull0 + 8 * (p.keep_acc ? it : acc));
        }
        __syncwarp();
        gs = ge;
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: TMEM -> registers -> swizzled smem -> TMA store =====================
    const int q = warp & 3;                  // TMEM lane quarter (hardware: warp w may touch lanes 32*(w%4)..+31)
    const int half = (warp - 4) >> 2;        // the two warps of a quarter split the 32-column chunks
    const int row = q * 32 + lane;           // pixel row inside the tile
    const int wl = row % p.TW, hl = (row / p.TW) % p.TH, nl = row / (p.TW * p.TH);
    const bool issuer = (warp == 4 && lane == 0);
    const uint32_t so_base = smem_u32(stage_out);
    const bool do_stats = p.bn_sums != nullptr;
    int it = 0, sidx = 0;
    int run_nt = -1;              // fused statistics: channel tile of the running sums below
    float run1 = 0.f, run2 = 0.f;
    double rund1 = 0.0, rund2 = 0.0;   // (fp32-output mode keeps the running sums in double)
    bool bwd_done = false;
    // (a separate template variant, VAR == 3: merely compiling this path into the production kernel cost the
    // ordinary fprop / dgrad launches 0.41 ms per step -- 5.81 vs 5.19 ms of conv time, A/B on the same GPU)
    if constexpr (!OUT32 && CL == 1 && VAR == 3) {
      if (p.fuse_bwd) {
        // =============== dgrad + batch-norm backward of the produced gradient's layer (see ConvParams::fuse_bwd) ====
        bwd_done = true;
        const int te = (int)threadIdx.x - 128;
        const uint32_t so = so_base;                                     // one staging buffer [NBOX][128 px][64 ch]
        const uint32_t xb = so_base + (uint32_t)p.xbuf_off;              // the x tile, same layout (TMA, SWIZZLE_128B)
        float* cf = reinterpret_cast<float*>(stage_out + p.coef_off);    // [5][BN] per-channel coefficients
        const uint32_t xbar = so_base + (uint32_t)p.xbar_off;
        uint32_t xphase = 0;
        const int C = p.Cdst;
        constexpr int CP = BN / 2, RG = 256 / CP, RPG = BM / RG;
        const int cp = te % CP, rg = te / CP;
        // ---- per (work item, sub tile): x tile -> mask -> g = dA * mask (bf16, staged) -> column sums
        for (int pass = 0; pass < 2; ++pass) {
          int itb = 0, cur_nt = -1;
          for (int t = cid; t < total_tiles; t += ncl, ++itb) {
            const int nt = t % p.n_tiles;
            if (nt != cur_nt) {
              asm volatile("bar.sync 1, 256;" ::: "memory");             // readers of the previous table are done
              if (te < BN) {
                const int c = nt * BN + te;
                const float mean = p.bwd_bnp[c], istd = p.bwd_bnp[C + c], A = p.bwd_bnp[2 * C + c];
                cf[te] = mean;
                cf[BN + te] = A;
                cf[2 * BN + te] = p.bwd_bnp[3 * C + c];
                if (pass == 1) {
                  double s1 = 0, s2 = 0;
#pragma unroll
                  for (int r = 0; r < BASI_BN_REPLICAS; ++r) {
                    s1 += __ldcg(p.bwd_sums + (size_t)r * 2 * C + c);
                    s2 += __ldcg(p.bwd_sums + (size_t)r * 2 * C + C + c);
                  }
                  cf[3 * BN + te] = A * (float)(s1 / p.bwd_count);                    // Bc
                  cf[4 * BN + te] = A * istd * (float)(s2 / p.bwd_count);             // Cc
                  if (t / p.n_tiles == 0) {                                           // one CTA per channel tile
                    p.bwd_dbeta[c] += (float)s1;
                    p.bwd_dgamma[c] += (float)s2;
                  }
                }
              }
              asm volatile("bar.sync 1, 256;" ::: "memory");
              cur_nt = nt;
            }
#pragma unroll 1
            for (int sub = 0; sub < MT; ++sub) {
              const int mt = (t / p.n_tiles) * MT + sub;
              const int tw = mt % p.tiles_w, th = (mt / p.tiles_w) % p.tiles_h, tn = mt / (p.tiles_w * p.tiles_h);
              const int w0 = tw * p.TW, h0 = th * p.TH, n0 = tn * p.TN;
              const bool valid = (w0 + wl < p.W) && (h0 + hl < p.H) && (n0 + nl < p.N);
              if (issuer) {
                // (pass 2) the TMA store of the previous tile must have read the staging buffer; then fetch the x tile
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                mbar_expect_tx(xbar, BN >= 64 ? NBOX * A_BYTES : BM * 2 * BN);   // (32-channel tiles: 64-byte rows)
#pragma unroll
                for (int j = 0; j < NBOX; ++j)
                  tma_load_4d(xb + j * A_BYTES, &mapD2, xbar, nt * BN + j * BOXC, w0, h0, n0);
              }
              if (pass == 0 && sub == 0) {
                mbar_wait(tfull0 + 8 * itb, 0);
                tc_fence_after();
              }
              asm volatile("bar.sync 1, 256;" ::: "memory");             // staging buffer free (issuer waited)
              mbar_wait(xbar, xphase);
              xphase ^= 1;
              const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + itb * (MT * BN) + sub * BN;
#pragma unroll
              for (int c = 0; c < BN / 32; ++c) {
                if ((c & 1) != half && BN >= 64) continue;
                if (BN < 64 && half != 0) continue;
                uint32_t r[32];
                tmem_ld32(taddr + c * 32, r);
                const uint32_t xline = xb + (uint32_t)(c >> 1) * A_BYTES + (uint32_t)row * (BN >= 64 ? 128 : 2 * BN);
                const uint32_t line = so + (uint32_t)(c >> 1) * A_BYTES + (uint32_t)row * (BN >= 64 ? 128 : 2 * BN);
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                  const uint32_t chunk = BN >= 64 ? (uint32_t)(((c & 1) * 4 + v) ^ (row & 7)) : (uint32_t)v;
                  uint32_t xw[4], pk[4];
                  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(xw[0]), "=r"(xw[1]), "=r"(xw[2]), "=r"(xw[3])
                               : "r"(xline + chunk * 16));
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const int col = c * 32 + v * 8 + 2 * e;
                    float res[2];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                      const float xv = h ? h16_hi(xw[e]) : h16_lo(xw[e]);
                      const float xc = xv - cf[col + h];
                      float g = __uint_as_float(r[v * 8 + 2 * e + h]);
                      if (!valid) g = 0.f;
                      if (p.bwd_relu) g = fmaf(xc, cf[BN + col + h], cf[2 * BN + col + h]) > 0.f ? g : 0.f;
                      res[h] = pass == 0 ? g : fmaf(cf[BN + col + h], g, -cf[3 * BN + col + h]) - xc * cf[4 * BN + col + h];
                    }
                    pk[e] = h16_pack(res[0], res[1]);
                  }
                  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(line + chunk * 16), "r"(pk[0]), "r"(pk[1]),
                               "r"(pk[2]), "r"(pk[3])
                               : "memory");
                }
              }
              if (pass == 1) {
                tc_fence_before();
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (issuer) {
#pragma unroll
                  for (int j = 0; j < NBOX; ++j) {
                    if (p.accumulate) tma_reduce_add_4d(&mapD, so + j * A_BYTES, nt * BN + j * BOXC, w0, h0, n0);
                    else tma_store_4d(&mapD, so + j * A_BYTES, nt * BN + j * BOXC, w0, h0, n0);
                  }
                  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                continue;
              }
              asm volatile("bar.sync 1, 256;" ::: "memory");             // g staged
              // column sums of g and g * (x - mean): thread = (column pair, group of rows), conflict-free word loads
              float s0 = 0.f, s1v = 0.f, q0 = 0.f, q1 = 0.f;
              {
                const uint32_t colb = (BN >= 64 ? (uint32_t)(cp >> 5) * A_BYTES : 0u);
                const float m0 = cf[2 * cp], m1 = cf[2 * cp + 1];
#pragma unroll 8
                for (int r = rg * RPG; r < (rg + 1) * RPG; ++r) {
                  const uint32_t off = BN >= 64 ? colb + (uint32_t)r * 128 + (uint32_t)((((cp & 31) >> 2) ^ (r & 7)) << 4) +
                                                      (uint32_t)(cp & 3) * 4
                                                : (uint32_t)r * (2 * BN) + (uint32_t)cp * 4;
                  uint32_t wg, wx;
                  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(wg) : "r"(so + off));
                  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(wx) : "r"(xb + off));
                  const float ga = h16_lo(wg), gb = h16_hi(wg);
                  const float xa = h16_lo(wx) - m0, xbv = h16_hi(wx) - m1;
                  s0 += ga; q0 = fmaf(ga, xa, q0);
                  s1v += gb; q1 = fmaf(gb, xbv, q1);
                }
              }
              stat_s[rg * (2 * BN) + 2 * cp] = s0;
              stat_s[rg * (2 * BN) + 2 * cp + 1] = s1v;
              stat_s[rg * (2 * BN) + BN + 2 * cp] = q0;
              stat_s[rg * (2 * BN) + BN + 2 * cp + 1] = q1;
              asm volatile("bar.sync 2, 256;" ::: "memory");
              if (te < BN) {
                float t1 = 0.f, t2 = 0.f;
#pragma unroll
                for (int g2 = 0; g2 < RG; ++g2) {
                  t1 += stat_s[g2 * (2 * BN) + te];
                  t2 += stat_s[g2 * (2 * BN) + BN + te];
                }
                if (nt != run_nt) {
                  if (run_nt >= 0) {
                    double* rep = p.bwd_sums + (size_t)(cid % BASI_BN_REPLICAS) * 2 * C;
                    atomicAdd(rep + run_nt * BN + te, (double)run1);
                    atomicAdd(rep + C + run_nt * BN + te, (double)run2 * (double)p.bwd_bnp[C + run_nt * BN + te]);
                  }
                  run_nt = nt; run1 = 0.f; run2 = 0.f;
                }
                run1 += t1;
                run2 += t2;
              }
            }
          }
          if (pass == 0) {
            if (run_nt >= 0 && te < BN) {
              double* rep = p.bwd_sums + (size_t)(cid % BASI_BN_REPLICAS) * 2 * C;
              atomicAdd(rep + run_nt * BN + te, (double)run1);
              atomicAdd(rep + C + run_nt * BN + te, (double)run2 * (double)p.bwd_bnp[C + run_nt * BN + te]);
            }
            __threadfence();
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (threadIdx.x == 128) grid_barrier_thread0(p.bwd_counter);
            asm volatile("bar.sync 1, 256;" ::: "memory");
          }
        }
        tc_fence_before();
      }
    }
    for (int t = cid; t < total_tiles && !bwd_done; t += ncl, ++it) {
      int acc = it & 1;
      uint32_t acc_phase = (it >> 1) & 1;
      const int nt = t % p.n_tiles;
      const bool dbg = TIMERS && p.debug == 30 && blockIdx.x == 0 && issuer;
      long long c0 = dbg ? dbg_clock() : 0;
      // fp32 output (split mode): the tile arrives as a sequence of accumulation groups, each in a fresh TMEM
      // accumulator; they are summed here in fp32 registers (round to nearest) -- see ConvParams::big_chunks
      constexpr int NCH = OUT32 ? (BN >= 64 ? BN / 64 : 1) : 1;
      float accv[NCH][32];
      if constexpr (OUT32) {
        static_assert(!OUT32 || MT == 1, "split mode: one pixel tile per work item");
        bool first = true;
        for (int gs = 0; gs < ksteps; ++it) {
          const int ge = group_end(p, gs, ksteps);
          acc = it & 1;
          acc_phase = (it >> 1) & 1;
          mbar_wait(tfull0 + 8 * acc, acc_phase);
          tc_fence_after();
          const uint32_t ga = tmem_base + ((uint32_t)(q * 32) << 16) + acc * (MT * BN);
#pragma unroll
          for (int c = 0; c < BN / 32; ++c) {
            if ((c & 1) != half && BN >= 64) continue;
            if (BN < 64 && half != 0) continue;
            uint32_t r[32];
            tmem_ld32(ga + c * 32, r);
#pragma unroll
            for (int j = 0; j < 32; ++j)
              accv[c >> 1][j] = first ? __uint_as_float(r[j]) : accv[c >> 1][j] + __uint_as_float(r[j]);
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty0 + 8 * acc);      // this accumulator may be overwritten
          first = false;
          gs = ge;
        }
        --it;      // (the tile loop increments once more)
      } else {
        if (p.fuse_apply) mbar_wait(tfull0 + 8 * it, 0);     // one barrier per work item, each completes once
        else mbar_wait(tfull0 + 8 * acc, acc_phase);
        if (dbg) { const long long c1 = dbg_clock(); g_tc_dbg[0] += c1 - c0; g_tc_dbg[1] += 1; c0 = c1; }
        tc_fence_after();
      }
#pragma unroll 1
      for (int sub = 0; sub < MT; ++sub, ++sidx) {
      const int mt = ((t / p.n_tiles) * CL + crank) * MT + sub;
      const int tw = mt % p.tiles_w, th = (mt / p.tiles_w) % p.tiles_h, tn = mt / (p.tiles_w * p.tiles_h);
      const int w0 = tw * p.TW, h0 = th * p.TH, n0 = tn * p.TN;
      const bool valid = (w0 + wl < p.W) && (h0 + hl < p.H) && (n0 + nl < p.N);
      // the TMA store that last used this staging buffer must have finished reading it
      const uint32_t so = so_base + (p.out_bufs == 2 ? (uint32_t)(sidx & 1) * (NBOX * A_BYTES) : 0u);
      if (issuer) {
        const long long w0c = dbg ? dbg_clock() : 0;
        if (p.out_bufs == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        if (dbg) g_tc_dbg[2] += dbg_clock() - w0c;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (p.fuse_apply ? it : acc) * (MT * BN) + sub * BN;
#pragma unroll
      for (int c = 0; c < BN / 32; ++c) {
        if ((c & 1) != half && BN >= 64) continue;
        if (BN < 64 && half != 0) continue;
        uint32_t r[32];
        if constexpr (OUT32) {
          // fp32 box c: row `row` is one 128-byte line of 32 floats, 16-byte chunks XOR-swizzled with (row & 7)
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = (do_stats && !valid) ? 0u : __float_as_uint(accv[c >> 1][j]);
          const uint32_t line32 = so + (uint32_t)c * A_BYTES + (uint32_t)row * 128;
#pragma unroll
          for (int v = 0; v < 8; ++v) {
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(line32 + (uint32_t)((v ^ (row & 7)) << 4)),
                         "r"(r[4 * v]), "r"(r[4 * v + 1]), "r"(r[4 * v + 2]), "r"(r[4 * v + 3])
                         : "memory");
          }
          continue;
        }
        tmem_ld32(taddr + c * 32, r);
        if (p.bias != nullptr) {
          const float* bp = p.bias + nt * BN + c * 32;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float v = __uint_as_float(r[j]) + __ldg(bp + j);
            if (p.bias_relu) v = fmaxf(v, 0.f);
            r[j] = __float_as_uint(v);
          }
        }
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          pk[j] = h16_pack(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
          // rows outside the tensor are never stored (TMA clips them) but the statistics below are summed from the
          // staging tile: for k > 1 such rows hold phantom outputs of the padded convolution, so they are zeroed
          if (do_stats && !valid) pk[j] = 0u;
        }
        // row `row` of box (c/2): 128-byte line, 16-byte chunks XOR-swizzled with (row & 7) like TMA SWIZZLE_128B
        // (BN == 32: one 32-channel box, 64-byte rows, no swizzle)
        const uint32_t line = so + (uint32_t)(c >> 1) * A_BYTES + (uint32_t)row * (BN >= 64 ? 128 : 2 * BN);
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const uint32_t chunk = BN >= 64 ? (uint32_t)(((c & 1) * 4 + v) ^ (row & 7)) : (uint32_t)v;
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(line + chunk * 16), "r"(pk[4 * v]),
                       "r"(pk[4 * v + 1]), "r"(pk[4 * v + 2]), "r"(pk[4 * v + 3])
                       : "memory");
        }
      }
      tc_fence_before();
      __syncwarp();
      if (!OUT32 && !p.fuse_apply && lane == 0 && sub == MT - 1) {      // TMEM accumulator is free again
        if (CL == 2) mbar_arrive_cluster(mapa_u32(tempty0 + 8 * acc, 0));
        else mbar_arrive(tempty0 + 8 * acc);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // staging writes -> visible to the TMA engine
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (issuer) {
#pragma unroll
        for (int j = 0; j < NBOX; ++j) {
          if (p.accumulate) tma_reduce_add_4d(&mapD, so + j * A_BYTES, nt * BN + j * BOXC, w0, h0, n0);
          else tma_store_4d(&mapD, so + j * A_BYTES, nt * BN + j * BOXC, w0, h0, n0);
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      if (do_stats && p.debug != 2) {
        // statistics of the values as stored (bf16): column sums straight from the staging tile.  Thread = (column
        // pair, group of rows): one conflict-free 32-bit shared-memory load per row (a warp reads 32 consecutive
        // words of one 128-byte line) -- the 4 x 31-shuffle transpose-reduce this replaces cost 6 of the 8 us the
        // statistics added to a conv4 1x1-increase fprop.
        const int te = (int)threadIdx.x - 128;
        if constexpr (OUT32) {
          // thread = (column, group of rows); a warp reads 32 consecutive floats of one 128-byte line per row
          constexpr int RG32 = 256 / BN, RPG32 = BM / RG32;
          const int col = te % BN, rg32 = te / BN;
          const uint32_t cb = so + (uint32_t)(col >> 5) * A_BYTES + (uint32_t)(col & 3) * 4;
          const int cw = (col & 31) >> 2;
          // float32 mode: the sums (and the variance E[x^2] - mean^2 derived from them) are accumulated in double
          double s0 = 0.0, q0 = 0.0;
#pragma unroll 8
          for (int r = rg32 * RPG32; r < (rg32 + 1) * RPG32; ++r) {
            float a;
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(a) : "r"(cb + (uint32_t)r * 128 + (uint32_t)((cw ^ (r & 7)) << 4)));
            s0 += (double)a; q0 = fma((double)a, (double)a, q0);
          }
          double* stat_d = reinterpret_cast<double*>(stat_s);          // [row groups][2*BN] = 512 doubles
          stat_d[rg32 * (2 * BN) + col] = s0;
          stat_d[rg32 * (2 * BN) + BN + col] = q0;
          asm volatile("bar.sync 2, 256;" ::: "memory");
          if (te < BN) {
            double t1 = 0.0, t2 = 0.0;
#pragma unroll
            for (int g2 = 0; g2 < RG32; ++g2) {
              t1 += stat_d[g2 * (2 * BN) + te];
              t2 += stat_d[g2 * (2 * BN) + BN + te];
            }
            if (nt != run_nt) {
              if (run_nt >= 0 && p.debug == 0) {
                double* rep = p.bn_sums + (size_t)(cid % BASI_BN_REPLICAS) * 2 * p.Cdst;
                atomicAdd(rep + run_nt * BN + te, rund1);
                atomicAdd(rep + p.Cdst + run_nt * BN + te, rund2);
              }
              run_nt = nt; rund1 = 0.0; rund2 = 0.0;
            }
            rund1 += t1;
            rund2 += t2;
          }
          continue;
        }
        constexpr int CP = BN / 2, RG = 256 / CP, RPG = BM / RG;
        const int cp = te % CP, rg = te / CP;
        float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
        const uint32_t colbase = BN >= 64 ? so + (uint32_t)(cp >> 5) * A_BYTES : so;
#pragma unroll 8
        for (int r = rg * RPG; r < (rg + 1) * RPG; ++r) {
          const uint32_t addr = BN >= 64 ? colbase + (uint32_t)r * 128 + (uint32_t)((((cp & 31) >> 2) ^ (r & 7)) << 4) +
                                               (uint32_t)(cp & 3) * 4
                                         : colbase + (uint32_t)r * (2 * BN) + (uint32_t)cp * 4;
          uint32_t w;
          asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w) : "r"(addr));
          const float a = h16_lo(w), b = h16_hi(w);
          s0 += a; q0 = fmaf(a, a, q0);
          s1 += b; q1 = fmaf(b, b, q1);
        }
        stat_s[rg * (2 * BN) + 2 * cp] = s0;
        stat_s[rg * (2 * BN) + 2 * cp + 1] = s1;
        stat_s[rg * (2 * BN) + BN + 2 * cp] = q0;
        stat_s[rg * (2 * BN) + BN + 2 * cp + 1] = q1;
        asm volatile("bar.sync 2, 256;" ::: "memory");
        if (te < BN) {
          float t1 = 0.f, t2 = 0.f;
#pragma unroll
          for (int g2 = 0; g2 < RG; ++g2) {
            t1 += stat_s[g2 * (2 * BN) + te];
            t2 += stat_s[g2 * (2 * BN) + BN + te];
          }
          // running sums of this CTA's tiles of one channel tile; flushed (2 double atomics per column) only when
          // the channel tile changes and at the end
          if (nt != run_nt) {
            if (run_nt >= 0 && p.debug == 0) {
              double* rep = p.bn_sums + (size_t)(cid % BASI_BN_REPLICAS) * 2 * p.Cdst;
              atomicAdd(rep + run_nt * BN + te, (double)run1);
              atomicAdd(rep + p.Cdst + run_nt * BN + te, (double)run2);
            }
            run_nt = nt; run1 = 0.f; run2 = 0.f;
          }
          run1 += t1;
          run2 += t2;
        }
      }
      }   // sub
      if (dbg) g_tc_dbg[3] += dbg_clock() - c0;     // epilogue work of this tile (after the accumulator arrived)
    }
    if (do_stats && run_nt >= 0 && (int)threadIdx.x - 128 < BN && p.debug == 0) {
      double* rep = p.bn_sums + (size_t)(cid % BASI_BN_REPLICAS) * 2 * p.Cdst;
      atomicAdd(rep + run_nt * BN + ((int)threadIdx.x - 128), OUT32 ? rund1 : (double)run1);
      atomicAdd(rep + p.Cdst + run_nt * BN + ((int)threadIdx.x - 128), OUT32 ? rund2 : (double)run2);
    }
    if constexpr (!OUT32 && CL == 1 && VAR == 4) {      // (own template variant, like VAR == 3)
      if (p.fuse_apply) {
        // ---- grid barrier: every CTA's statistics are in the global sums
        __threadfence();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (threadIdx.x == 128) grid_barrier_thread0(p.bn_counter);
        asm volatile("bar.sync 1, 256;" ::: "memory");
        // ---- second pass: y = relu(scale * acc + shift) straight from the fp32 accumulators
        const int te = (int)threadIdx.x - 128;
        float* sc_s = stat_s;              // [BN] scale, [BN] shift of the current channel tile
        int it2 = 0, cur_nt = -1;
        for (int t = cid; t < total_tiles; t += ncl, ++it2) {
          const int nt = t % p.n_tiles;
          if (nt != cur_nt) {
            asm volatile("bar.sync 1, 256;" ::: "memory");       // the previous tile's readers of sc_s are done
            if (te < BN) {
              const int C = p.Cdst, c = nt * BN + te;
              double s1 = 0, s2 = 0;
#pragma unroll
              for (int r = 0; r < BASI_BN_REPLICAS; ++r) {
                s1 += __ldcg(p.bn_sums + (size_t)r * 2 * C + c);
                s2 += __ldcg(p.bn_sums + (size_t)r * 2 * C + C + c);
              }
              const double mean = s1 / p.bn_count;
              double var = s2 / p.bn_count - mean * mean;
              if (var < 0) var = 0;
              const double istd = 1.0 / sqrt(var + (double)p.bn_eps);
              const float g = p.bn_gamma[c], b = p.bn_beta[c];
              const float scl = (float)((double)g * istd);
              sc_s[te] = scl;
              sc_s[BN + te] = (float)((double)b - mean * (double)g * istd);
              if (t / p.n_tiles == 0 && p.bn_bnp != nullptr) {     // one CTA per channel tile publishes the parameters
                p.bn_bnp[c] = (float)mean;
                p.bn_bnp[C + c] = (float)istd;
                p.bn_bnp[2 * C + c] = scl;
                p.bn_bnp[3 * C + c] = b;
              }
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            cur_nt = nt;
          }
#pragma unroll 1
          for (int sub = 0; sub < MT; ++sub, ++sidx) {
            const int mt = (t / p.n_tiles) * MT + sub;
            const int tw = mt % p.tiles_w, th = (mt / p.tiles_w) % p.tiles_h, tn = mt / (p.tiles_w * p.tiles_h);
            const int w0 = tw * p.TW, h0 = th * p.TH, n0 = tn * p.TN;
            const uint32_t so = so_base + (p.out_bufs == 2 ? (uint32_t)(sidx & 1) * (NBOX * A_BYTES) : 0u);
            if (issuer) {
              if (p.out_bufs == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
              else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + it2 * (MT * BN) + sub * BN;
#pragma unroll
            for (int c = 0; c < BN / 32; ++c) {
              if ((c & 1) != half && BN >= 64) continue;
              if (BN < 64 && half != 0) continue;
              uint32_t r[32];
              tmem_ld32(taddr + c * 32, r);
              uint32_t pk[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const int col = c * 32 + 2 * j;
                float v0 = fmaf(__uint_as_float(r[2 * j]), sc_s[col], sc_s[BN + col]);
                float v1 = fmaf(__uint_as_float(r[2 * j + 1]), sc_s[col + 1], sc_s[BN + col + 1]);
                if (p.apply_relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
                pk[j] = h16_pack(v0, v1);
              }
              const uint32_t line = so + (uint32_t)(c >> 1) * A_BYTES + (uint32_t)row * (BN >= 64 ? 128 : 2 * BN);
#pragma unroll
              for (int v = 0; v < 4; ++v) {
                const uint32_t chunk = BN >= 64 ? (uint32_t)(((c & 1) * 4 + v) ^ (row & 7)) : (uint32_t)v;
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(line + chunk * 16), "r"(pk[4 * v]),
                             "r"(pk[4 * v + 1]), "r"(pk[4 * v + 2]), "r"(pk[4 * v + 3])
                             : "memory");
              }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (issuer) {
#pragma unroll
              for (int j = 0; j < NBOX; ++j) tma_store_4d(&mapD2, so + j * A_BYTES, nt * BN + j * BOXC, w0, h0, n0);
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
          }
        }
        tc_fence_before();
      }
    }
    // the staging buffers must have been READ before the CTA exits; the writes themselves complete with the grid
    // (waiting for them here put ~1.5 us of HBM write latency on every launch's tail; BASI_TC_DEBUG_STATS=40: old wait)
    if (issuer) {
      if (p.debug == 40) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    if (p.bn_sums != nullptr) __threadfence();
  }
  tc_fence_before();
  __syncthreads();
  if (CL == 2) cluster_sync_all();   // no CTA exits while its peer may still multicast into it / arrive on its barriers
  if (warp == 2 && p.debug != 11) {
    tc_fence_after();
    if (CL == 2)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
  if (p.bn_sums != nullptr && p.bn_bnp != nullptr && !p.fuse_apply) {
    // last CTA to finish turns the sums into the per-channel parameters (fused finalize)
    if (threadIdx.x == 0) {
      const unsigned int t = atomicAdd(p.bn_counter, 1u);
      *s_is_last = (t == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (*s_is_last) {
      __threadfence();
      const int C = p.Cdst;
      for (int c = threadIdx.x; c < C; c += NTHREADS_CONV) {
        double s1 = 0, s2 = 0;
#pragma unroll
        for (int r = 0; r < BASI_BN_REPLICAS; ++r) {
          s1 += __ldcg(p.bn_sums + (size_t)r * 2 * C + c);
          s2 += __ldcg(p.bn_sums + (size_t)r * 2 * C + C + c);
        }
        const double mean = s1 / p.bn_count;
        double var = s2 / p.bn_count - mean * mean;
        if (var < 0) var = 0;
        const double istd = 1.0 / sqrt(var + (double)p.bn_eps);
        p.bn_bnp[c] = (float)mean;
        p.bn_bnp[C + c] = (float)istd;
        p.bn_bnp[2 * C + c] = (float)((double)p.bn_gamma[c] * istd);
        p.bn_bnp[3 * C + c] = p.bn_beta[c];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------
// wgrad: D[co 128][ci BN] = sum_px dY[px][co]^T * X_tap[px][ci]   (both operands MN-major: pixels are K)
// one work item = (tap, co tile of 128, ci tile of BN, pixel-tile range); the accumulator row is an output
// channel, so for a fixed ci the 32 lanes of a warp add to 32 consecutive floats of the HWIO gradient
// (coalesced red.global.add).  Channel counts that are not multiples of the tile are zero-filled by TMA.
// ------------------------------------------------------------------------------------------------------
struct WgradParams {
  int TW, TH, TN;
  int tiles_w, tiles_h, tiles_n, m_tiles;
  int taps, kw;
  int ci_tiles, co_tiles;   // ceil(Cin/BN), ceil(Cout/128)
  int splits, tiles_per_split;
  int off_h, off_w, step;   // x coordinate = p + off + r*step (fprop mapping)
  int Cin, Cout;
  int stages;
  int nterms;               // 1, or 6 in split-operand mode: (dy part, x part) = (0,0) (0,1) (1,0) (0,2) (2,0) (1,1)
};

template <int BN, int GBOX>
__global__ void __launch_bounds__(NTHREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapG,
                float* __restrict__ dw, const WgradParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 128 co = two 64-channel boxes of [128 px][128 B]; layers with at most 64 output channels (GBOX == 1) load one box
  // and alias the second 64-row block of the MMA onto it (LBO = 0): its accumulator rows are never read
  constexpr int GB = GBOX * A_BYTES;
  constexpr int XB = (BN / 64) * A_BYTES;    // BN ci = BN/64 boxes
  constexpr int STAGE = GB + XB;
  constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * STAGE);
  const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * p.stages, tfull = empty0 + 8 * p.stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * p.stages + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapX) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapG) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(full0 + 8 * i, 1);
      mbar_init(empty0 + 8 * i, 1);
    }
    mbar_init(tfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();

  // work item of this CTA
  int wi = blockIdx.x;
  const int split = wi % p.splits; wi /= p.splits;
  const int cit = wi % p.ci_tiles; wi /= p.ci_tiles;
  const int cot = wi % p.co_tiles; wi /= p.co_tiles;
  const int tap = wi;
  const int r = tap / p.kw, s = tap - r * p.kw;
  const int mt_beg = split * p.tiles_per_split;
  const int mt_end = min(p.m_tiles, mt_beg + p.tiles_per_split);
  const int nsteps = (mt_end - mt_beg) * p.nterms;

  if (warp == 0) {
    {
      int stage = 0;
      uint32_t phase = 0;
      // split-operand mode: the six significant products of (dy_hi + dy_mid + dy_lo) x (x_hi + x_mid + x_lo), the
      // small ones first and hi x hi last (tcgen05 accumulates with truncation: small terms added to a full-size
      // accumulator would lose their low bits); plain mode runs only the last term
      for (int term = 6 - p.nterms; term < 6; ++term) {
        const int gpart = term == 0 ? 1 : (term == 1 ? 2 : (term == 3 ? 1 : 0));     // (1,1) (2,0) (0,2) (1,0) (0,1) (0,0)
        const int xpart = term == 0 ? 1 : (term == 2 ? 2 : (term == 4 ? 1 : 0));
        const int gc = gpart * p.Cout + cot * 128, xc = xpart * p.Cin + cit * BN;
        for (int mt = mt_beg; mt < mt_end; ++mt) {
          const int tw = mt % p.tiles_w, th = (mt / p.tiles_w) % p.tiles_h, tn = mt / (p.tiles_w * p.tiles_h);
          const int w0 = tw * p.TW, h0 = th * p.TH, n0 = tn * p.TN;
          mbar_wait(empty0 + 8 * stage, phase ^ 1);
          const uint32_t sa = smem_u32(smem + (size_t)stage * STAGE);
          const uint32_t fb = full0 + 8 * stage;
          if (elect_one()) {
            mbar_expect_tx(fb, STAGE);
            tma_load_4d(sa, &mapG, fb, gc, w0, h0, n0);
            if (GBOX == 2) tma_load_4d(sa + A_BYTES, &mapG, fb, gc + 64, w0, h0, n0);
            const int xw = w0 + p.off_w + s * p.step, xh = h0 + p.off_h + r * p.step;
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) tma_load_4d(sa + GB + j * A_BYTES, &mapX, fb, xc + j * 64, xw, xh, n0);
          }
          __syncwarp();
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // warp-uniform loops, the elected lane issues (see elect_one)
    {
      constexpr uint32_t idesc = make_idesc(BN, 1, 1);
      int stage = 0;
      uint32_t phase = 0;
      for (int st = 0; st < nsteps; ++st) {
        mbar_wait(full0 + 8 * stage, phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + (size_t)stage * STAGE);
        // MN-major: LBO = 16 KB between 64-channel blocks, SBO = 1 KB between groups of 8 pixel rows
        const uint64_t adesc = make_desc(sa, GBOX == 2 ? A_BYTES : 0, 1024);
        const uint64_t bdesc = make_desc(sa + GB, A_BYTES, 1024);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < BM / 16; ++k) {
            // 16 pixels (K) further = 16 rows of 128 B = 2 KB -> +128 in 16-B units
            umma_f16(tmem_base, adesc + 128 * k, bdesc + 128 * k, idesc, (st | k) != 0);
          }
          umma_commit(empty0 + 8 * stage);
        }
        __syncwarp();
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (elect_one()) umma_commit(tfull);
      __syncwarp();
    }
  } else if (warp >= 4) {
    const int q = warp - 4;
    const int co = cot * 128 + q * 32 + lane;
    if (nsteps > 0) {
      mbar_wait(tfull, 0);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
      const bool covalid = co < p.Cout;
#pragma unroll
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t rr[32];
        tmem_ld32(taddr + c * 32, rr);
        const int ci0 = cit * BN + c * 32;
        if (covalid && ci0 < p.Cin) {
          float* out = dw + ((int64_t)tap * p.Cin + ci0) * p.Cout + co;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (ci0 + j < p.Cin) atomicAdd(out + (int64_t)j * p.Cout, __uint_as_float(rr[j]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// w f32 [taps][cin][cout] -> bf16 [taps][cin][cout] and bf16 [taps][cout][cin]
__global__ void pack_weights_kernel(const float* __restrict__ w, bf16* __restrict__ w_io, bf16* __restrict__ w_oi,
                                    int taps, int cin, int cout) {
  __shared__ float tile[32][33];
  pdl_prologue();
  const int tap = blockIdx.z;
  const int ci0 = blockIdx.y * 32, co0 = blockIdx.x * 32;
  const float* src = w + (size_t)tap * cin * cout;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int ci = ci0 + i, co = co0 + threadIdx.x;
    float v = (ci < cin && co < cout) ? src[(size_t)ci * cout + co] : 0.f;
    tile[i][threadIdx.x] = v;
    if (ci < cin && co < cout) w_io[(size_t)tap * cin * cout + (size_t)ci * cout + co] = h16_from(v);
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int co = co0 + i, ci = ci0 + threadIdx.x;
    if (ci < cin && co < cout)
      w_oi[(size_t)tap * cin * cout + (size_t)co * cin + ci] = h16_from(tile[threadIdx.x][i]);
  }
}

// all layers in one launch: blockIdx.x -> (layer, tap, ci tile, co tile) through a small table in device memory
struct PackEntry {
  const float* w;
  bf16* w_io;
  bf16* w_oi;
  int taps, cin, cout;
  int block_start;   // first block of this layer
  int tiles_co, tiles_ci;
  int pad0, pad1;
};
// One 32 x 32 tile of one layer (thread block = 32 x 8).  Ends with all threads past their reads of `tile`.
__device__ __forceinline__ void pack_tile(const PackEntry& e, int blk, float (*tile)[33]);

// Persistent grid: every block copies the layer table into shared memory ONCE and then walks tiles blk = blockIdx.x,
// + gridDim.x, ...  (One block per tile with a binary search of the table in GLOBAL memory by thread 0 put seven
// dependent L2 round trips in front of every 4 KB tile: 78 us per step for 130 MB of traffic.)
__global__ void __launch_bounds__(256) pack_weights_multi_kernel(const PackEntry* __restrict__ table, int n_layers,
                                                                 int total_blocks) {
  __shared__ float tile[32][33];
  extern __shared__ __align__(16) uint8_t pack_tab_raw[];
  PackEntry* tab = reinterpret_cast<PackEntry*>(pack_tab_raw);
  pdl_prologue();
  {
    const int tid = threadIdx.y * 32 + threadIdx.x;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(table);
    uint32_t* dst = reinterpret_cast<uint32_t*>(pack_tab_raw);
    const int words = n_layers * (int)(sizeof(PackEntry) / 4);
    for (int i = tid; i < words; i += 256) dst[i] = src[i];
  }
  __syncthreads();
  for (int blk = blockIdx.x; blk < total_blocks; blk += gridDim.x) {
    int lo = 0, hi = n_layers - 1;
    while (lo < hi) {   // last entry with block_start <= blk (every thread searches the shared copy)
      const int mid = (lo + hi + 1) >> 1;
      if (tab[mid].block_start <= blk) lo = mid; else hi = mid - 1;
    }
    pack_tile(tab[lo], blk, tile);
    __syncthreads();    // the tile buffer is reused by the next iteration
  }
}

__device__ __forceinline__ void pack_tile(const PackEntry& e, int blk, float (*tile)[33]) {
  int b = blk - e.block_start;
  const int cot = b % e.tiles_co; b /= e.tiles_co;
  const int cit = b % e.tiles_ci; b /= e.tiles_ci;
  const int tap = b;
  const int ci0 = cit * 32, co0 = cot * 32;
  const size_t base = (size_t)tap * e.cin * e.cout;
  if (e.pad0 != 0) {
    // split-operand layouts (fp32-grade mode): w = hi + mid (+ lo), bf16 each.  The operand tensor of the convolution
    // holds `parts` channel blocks [hi|mid(|lo)] and the main loop runs `parts` passes of 64-column chunks over its
    // first parts, parts-1, .. blocks; the weight column of (pass p, operand part q, channel c) is
    // (passoff[p] * 64 + q * C + c) and holds weight part p for q = 0 .. parts-1-p: the significant products
    // (parts = 3: six, float32-exact; parts = 2: three).  Every other column stays zero (zero-initialised buffers).
    //   w_oi: fprop layout [tap][co][Kf], reduced channel C = cin, fparts;  w_io: dgrad layout [tap][ci][Kd], C = cout,
    //   dparts.  pad0 = 1: (3, 3), 2: (2, 2), 3: (3, 2) -- exact forward, three-product backward.
    const int fparts = e.pad0 == 2 ? 2 : 3, dparts = e.pad0 == 1 ? 3 : 2;
    int foff[3], doff[3];
    int accf = 0, accd = 0;
    for (int p = 0; p < 3; ++p) {
      foff[p] = accf; doff[p] = accd;
      if (p < fparts) accf += ((fparts - p) * e.cin + 63) / 64;
      if (p < dparts) accd += ((dparts - p) * e.cout + 63) / 64;
    }
    const size_t Kf = (size_t)accf * 64, Kd = (size_t)accd * 64;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
      const int ci = ci0 + i, co = co0 + threadIdx.x;
      const float v = (ci < e.cin && co < e.cout) ? e.w[base + (size_t)ci * e.cout + co] : 0.f;
      tile[i][threadIdx.x] = v;
      if (ci < e.cin && co < e.cout) {
        __nv_bfloat16 part[3];
        part[0] = __float2bfloat16_rn(v);
        const float r1 = v - __bfloat162float(part[0]);
        part[1] = __float2bfloat16_rn(r1);
        part[2] = __float2bfloat16_rn(r1 - __bfloat162float(part[1]));
        __nv_bfloat16* row = reinterpret_cast<__nv_bfloat16*>(e.w_io) + ((size_t)tap * e.cin + ci) * Kd;
        for (int p = 0; p < dparts; ++p)
          for (int q = 0; q < dparts - p; ++q) row[(size_t)doff[p] * 64 + (size_t)q * e.cout + co] = part[p];
      }
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
      const int co = co0 + i, ci = ci0 + threadIdx.x;
      if (ci < e.cin && co < e.cout) {
        const float v = tile[threadIdx.x][i];
        __nv_bfloat16 part[3];
        part[0] = __float2bfloat16_rn(v);
        const float r1 = v - __bfloat162float(part[0]);
        part[1] = __float2bfloat16_rn(r1);
        part[2] = __float2bfloat16_rn(r1 - __bfloat162float(part[1]));
        __nv_bfloat16* row = reinterpret_cast<__nv_bfloat16*>(e.w_oi) + ((size_t)tap * e.cout + co) * Kf;
        for (int p = 0; p < fparts; ++p)
          for (int q = 0; q < fparts - p; ++q) row[(size_t)foff[p] * 64 + (size_t)q * e.cin + ci] = part[p];
      }
    }
    return;
  }
  if (((e.cin | e.cout) & 3) == 0 && ((reinterpret_cast<uintptr_t>(e.w) & 15) == 0) &&
      (((reinterpret_cast<uintptr_t>(e.w_io) | reinterpret_cast<uintptr_t>(e.w_oi)) & 7) == 0)) {
    // one 128-bit load per thread covers the whole 4 KB tile at once (the row loop below is not unrolled -- its bound
    // is blockDim.y -- so a thread had ONE 4-byte load in flight: 1 KB per block, 1.8 TB/s for the whole refresh)
    const int tid = threadIdx.y * 32 + threadIdx.x;
    const int r = tid >> 3, c4 = (tid & 7) * 4;
    {
      const int ci = ci0 + r, co = co0 + c4;
      const bool ok = ci < e.cin && co < e.cout;          // (cout % 4 == 0: the four columns are valid together)
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ok) v = __ldg(reinterpret_cast<const float4*>(e.w + base + (size_t)ci * e.cout + co));
      tile[r][c4] = v.x; tile[r][c4 + 1] = v.y; tile[r][c4 + 2] = v.z; tile[r][c4 + 3] = v.w;
      if (ok)
        *reinterpret_cast<uint2*>(e.w_io + base + (size_t)ci * e.cout + co) = make_uint2(h16_pack(v.x, v.y), h16_pack(v.z, v.w));
    }
    __syncthreads();
    {
      const int co = co0 + r, ci = ci0 + c4;              // transposed: row = output channel, four input channels
      if (co < e.cout && ci < e.cin)
        *reinterpret_cast<uint2*>(e.w_oi + base + (size_t)co * e.cin + ci) =
            make_uint2(h16_pack(tile[c4][r], tile[c4 + 1][r]), h16_pack(tile[c4 + 2][r], tile[c4 + 3][r]));
    }
    return;
  }
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int ci = ci0 + i, co = co0 + threadIdx.x;
    const float v = (ci < e.cin && co < e.cout) ? e.w[base + (size_t)ci * e.cout + co] : 0.f;
    tile[i][threadIdx.x] = v;
    if (ci < e.cin && co < e.cout) e.w_io[base + (size_t)ci * e.cout + co] = h16_from(v);
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int co = co0 + i, ci = ci0 + threadIdx.x;
    if (ci < e.cin && co < e.cout) e.w_oi[base + (size_t)co * e.cin + ci] = h16_from(tile[threadIdx.x][i]);
  }
}

// ---------------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess || !p) {
    cudaGetLastError();
    return nullptr;
  }
  fn = (EncodeTiledFn)p;
  return fn;
}

// NHWC activation: dims (C, W, H, N), box (64, TW, TH, TN)
static int make_act_map(CUtensorMap* m, const basi_tensor* t, int TW, int TH, int TN, int box_c = 64) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("tc: cuTensorMapEncodeTiled not available");
    return BASI_E_CUDA;
  }
  // bf16: 64-channel boxes (128-byte rows); float32 (split-operand mode outputs): 32-channel boxes (128-byte rows)
  const bool f32 = t->dtype == BASI_F32;
  const cuuint64_t es_b = f32 ? 4 : 2;
  cuuint64_t dims[4] = {(cuuint64_t)t->c, (cuuint64_t)t->w, (cuuint64_t)t->h, (cuuint64_t)t->n};
  cuuint64_t strides[3] = {(cuuint64_t)t->ld * es_b, (cuuint64_t)t->ld * es_b * t->w,
                           (cuuint64_t)t->ld * es_b * t->w * t->h};
  cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)TW, (cuuint32_t)TH, (cuuint32_t)TN};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : (BASI_H16_FP16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16),
                   4, t->ptr, dims, strides,
                   box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   (uint64_t)box_c * es_b == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("tc: cuTensorMapEncodeTiled(activation) failed with %d", (int)r);
    return BASI_E_CUDA;
  }
  return BASI_OK;
}
// weights [taps][Ndim][Kdim] bf16, box (64, BN, 1)
static int make_w_map(CUtensorMap* m, const void* w, int taps, int ndim, int kdim, int bn) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("tc: cuTensorMapEncodeTiled not available");
    return BASI_E_CUDA;
  }
  cuuint64_t dims[3] = {(cuuint64_t)kdim, (cuuint64_t)ndim, (cuuint64_t)taps};
  cuuint64_t strides[2] = {(cuuint64_t)kdim * 2, (cuuint64_t)kdim * 2 * ndim};
  cuuint32_t box[3] = {64, (cuuint32_t)bn, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(m, BASI_H16_FP16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3,
                   const_cast<void*>(w), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("tc: cuTensorMapEncodeTiled(weights) failed with %d", (int)r);
    return BASI_E_CUDA;
  }
  return BASI_OK;
}

static void pick_tile(int N, int H, int W, int* TW, int* TH, int* TN) {
  long best = -1;
  for (int tw = 1; tw <= 128; tw *= 2)
    for (int th = 1; tw * th <= 128; th *= 2) {
      int tn = 128 / (tw * th);
      if (tw > 256 || th > 256 || tn > 256) continue;
      long tiles = (long)((W + tw - 1) / tw) * ((H + th - 1) / th) * ((N + tn - 1) / tn);
      // fewer tiles first (less padding waste); then prefer wide rows (longer contiguous runs)
      long score = tiles * 1024 - tw;
      if (best < 0 || score < best) {
        best = score;
        *TW = tw; *TH = th; *TN = tn;
      }
    }
}

}  // namespace tc
}  // namespace basi

using namespace basi;
using namespace basi::tc;

struct basi_tc_conv {
  int kind;
  int bn;
  int cluster;
  int mt;        // pixel tiles per work item (1 or 2)
  int gbox;      // wgrad: 64-channel boxes of the output-gradient operand (1 when Cout <= 64)
  int split;     // split-operand (fp32-grade) mode: bf16 [hi|mid|lo] operands, float32 destination
  CUtensorMap mapA, mapB, mapD, mapD2;
  ConvParams cp;
  int TW, TH, TN;   // pixel tile (for the second output map of the fused BN apply)
  WgradParams wp;
  bf16* dst;
  float* dw;
  int grid;
  size_t smem;
};

static bool tc_geometry_ok(int kind, const basi_conv_desc* d, const basi_tensor* x, const basi_tensor* y) {
  if (!d || !x || !y) return false;
  if (x->dtype != BASI_BF16 || y->dtype != BASI_BF16) return false;
  if (d->stride != 1 || d->kh != d->kw || d->relu) return false;
  if (x->n != y->n) return false;
  // 'same-size' convolutions only: output extent == input extent, symmetric padding
  const int span = (d->kh - 1) * d->dil;
  if (y->h != x->h || y->w != x->w || 2 * d->pad_t != span || 2 * d->pad_l != span) return false;
  if (x->ld % 8 || y->ld % 8) return false;
  if (((uintptr_t)x->ptr & 15) || ((uintptr_t)y->ptr & 15)) return false;
  const int cin = x->c, cout = y->c;
  // K (the reduced channel axis) only has to keep the TMA strides 16-byte aligned: a 64-channel box over a
  // narrower tensor is zero-filled.  N (the produced channel axis) is tiled by 32/64/128.
  if (kind == BASI_TC_FPROP) return cin % 8 == 0 && cin >= 16 && cout % 32 == 0;
  if (kind == BASI_TC_DGRAD) return cout % 8 == 0 && cout >= 16 && cin % 32 == 0;
  if (kind == BASI_TC_WGRAD) return cin % 8 == 0 && cin >= 16 && cout % 8 == 0 && cout >= 16;
  return false;
}

// split-operand mode: x, y are the logical float32 tensors (the bf16 [hi|mid|lo] operand tensors have 3x the channels)
static bool tc_geometry_ok_split(int kind, const basi_conv_desc* d, const basi_tensor* x, const basi_tensor* y) {
  if (!d || !x || !y) return false;
  if (x->dtype != BASI_F32 || y->dtype != BASI_F32) return false;
  if (d->stride != 1 || d->kh != d->kw || d->relu) return false;
  if (x->n != y->n) return false;
  const int span = (d->kh - 1) * d->dil;
  if (y->h != x->h || y->w != x->w || 2 * d->pad_t != span || 2 * d->pad_l != span) return false;
  if (x->ld % 4 || y->ld % 4) return false;
  if (((uintptr_t)x->ptr & 15) || ((uintptr_t)y->ptr & 15)) return false;
  const int cin = x->c, cout = y->c;
  if (kind == BASI_TC_FPROP) return cin % 8 == 0 && cin >= 16 && cout % 32 == 0;
  if (kind == BASI_TC_DGRAD) return cout % 8 == 0 && cout >= 16 && cin % 32 == 0;
  if (kind == BASI_TC_WGRAD) return cin % 8 == 0 && cin >= 16 && cout % 8 == 0 && cout >= 16;
  return false;
}

template <int BN>
static int launch_conv(basi_tc_conv* pl, cudaStream_t st) {
  static bool attr_set = false;
  constexpr int BNS = BN <= 128 ? BN : 128;      // channel-tile widths the two-tile / halo variants exist for
  if (!attr_set) {
    cudaFuncSetAttribute(conv_tc_kernel<BN, 1, 1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(conv_tc_kernel<BN, 2, 1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(conv_tc_kernel<BN, 1, 1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(conv_tc_kernel<BNS, 1, 2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(conv_tc_kernel<BNS, 1, 2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(conv_tc_kernel<BNS, 1, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr_set = true;
  }
  const dim3 grid(pl->grid), block(NTHREADS_CONV);
  if (pl->split) {
    if constexpr (BN <= 128) {
      static bool attr32 = false;
      if (!attr32) {
        cudaFuncSetAttribute(conv_tc_kernel<BN, 1, 1, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        attr32 = true;
      }
      basi::launch(conv_tc_kernel<BN, 1, 1, 0, true>, grid, block, pl->smem, st, pl->mapA, pl->mapB, pl->mapD, pl->mapD2, pl->cp);
    }
    return BASI_OK;
  }
  const bool timers = pl->cp.debug == 30;
  if (pl->cp.fuse_apply || pl->cp.fuse_bwd) {
    // grid barrier inside: cooperative launch (all CTAs co-resident; grid <= #SMs at one CTA per SM)
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = pl->smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int na = 0;
    // BASI_TC_FUSED_LAUNCH (experiment): "coop" cooperative only, "pdl" programmatic dependent launch only (co-residency
    // then rests on grid <= #SMs at one CTA per SM), "both" (default)
    const char* mode = exp_env("BASI_TC_FUSED_LAUNCH");
    const bool want_coop = !mode || strcmp(mode, "pdl") != 0;
    const bool want_pdl = (!mode || strcmp(mode, "coop") != 0) && pdl_enabled();
    if (want_coop) {
      attr[na].id = cudaLaunchAttributeCooperative;
      attr[na].val.cooperative = 1;
      ++na;
    }
    if (want_pdl) {
      attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[na].val.programmaticStreamSerializationAllowed = 1;
      ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    if (pl->cp.fuse_bwd) {
      if constexpr (BN <= 128) {
        static bool attr3 = false;
        if (!attr3) {
          cudaFuncSetAttribute(conv_tc_kernel<BN, 1, 1, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
          cudaFuncSetAttribute(conv_tc_kernel<BN, 1, 2, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
          attr3 = true;
        }
        if (pl->mt == 2)
          cudaLaunchKernelEx(&cfg, conv_tc_kernel<BN, 1, 2, 3>, pl->mapA, pl->mapB, pl->mapD, pl->mapD2, pl->cp);
        else
          cudaLaunchKernelEx(&cfg, conv_tc_kernel<BN, 1, 1, 3>, pl->mapA, pl->mapB, pl->mapD, pl->mapD2, pl->cp);
      }
      return BASI_OK;
    }
    static bool attr4 = false;
    if (!attr4) {
      cudaFuncSetAttribute(conv_tc_kernel<BN, 1, 1, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      cudaFuncSetAttribute(conv_tc_kernel<BNS, 1, 2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      attr4 = true;
    }
    if (pl->mt == 2 && BN <= 128)
      cudaLaunchKernelEx(&cfg, conv_tc_kernel<BNS, 1, 2, 4>, pl->mapA, pl->mapB, pl->mapD, pl->mapD2, pl->cp);
    else
      cudaLaunchKernelEx(&cfg, conv_tc_kernel<BN, 1, 1, 4>, pl->mapA, pl->mapB, pl->mapD, pl->mapD2, pl->cp);
    return BASI_OK;
  }
  if (pl->cp.halo && BN <= 128)
    basi::launch(conv_tc_kernel<BNS, 1, 1, 1>, grid, block, pl->smem, st, pl->mapA, pl->mapB, pl->mapD, pl->mapD2, pl->cp);
  else if (pl->mt == 2 && BN <= 128) {
    if (timers) basi::launch(conv_tc_kernel<BNS, 1, 2, 2>, grid, block, pl->smem, st, pl->mapA, pl->mapB, pl->mapD, pl->mapD2, pl->cp);
    else basi::launch(conv_tc_kernel<BNS, 1, 2, 0>, grid, block, pl->smem, st, pl->mapA, pl->mapB, pl->mapD, pl->mapD2, pl->cp);
  } else if (pl->cluster == 2)
    basi::launch_ex(conv_tc_kernel<BN, 2, 1, 0>, grid, block, pl->smem, st, 2, pl->mapA, pl->mapB, pl->mapD, pl->mapD2, pl->cp);
  else if (timers)
    basi::launch(conv_tc_kernel<BN, 1, 1, 2>, grid, block, pl->smem, st, pl->mapA, pl->mapB, pl->mapD, pl->mapD2, pl->cp);
  else
    basi::launch(conv_tc_kernel<BN, 1, 1, 0>, grid, block, pl->smem, st, pl->mapA, pl->mapB, pl->mapD, pl->mapD2, pl->cp);
  return BASI_OK;
}
template <int BN>
static int launch_wgrad(basi_tc_conv* pl, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(wgrad_tc_kernel<BN, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(wgrad_tc_kernel<BN, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr_set = true;
  }
  if (pl->gbox == 1)
    basi::launch(wgrad_tc_kernel<BN, 1>, dim3(pl->grid), dim3(NTHREADS), pl->smem, st, pl->mapA, pl->mapB, pl->dw, pl->wp);
  else
    basi::launch(wgrad_tc_kernel<BN, 2>, dim3(pl->grid), dim3(NTHREADS), pl->smem, st, pl->mapA, pl->mapB, pl->dw, pl->wp);
  return BASI_OK;
}

extern "C" {

int basi_tc_conv_supported(int kind, const basi_conv_desc* d, const basi_tensor* x, const basi_tensor* y) {
  return tc_geometry_ok(kind, d, x, y) ? 1 : 0;
}

int basi_tc_pack_weights(const float* w, void* w_io_bf16, void* w_oi_bf16, int taps, int cin, int cout, void* stream) {
  BASI_CHECK_ARG(w && w_io_bf16 && w_oi_bf16 && taps > 0 && cin > 0 && cout > 0, "tc_pack_weights: bad argument");
  dim3 grid((cout + 31) / 32, (cin + 31) / 32, taps), block(32, 8);
  basi::launch(pack_weights_kernel, grid, block, 0, (cudaStream_t)stream, w, (bf16*)w_io_bf16, (bf16*)w_oi_bf16, taps, cin,
               cout);
  BASI_CHECK_LAUNCH("tc_pack_weights");
  return BASI_OK;
}

int basi_tc_pack_weights_multi(const void* table_dev, int n_layers, int total_blocks, void* stream) {
  BASI_CHECK_ARG(table_dev && n_layers > 0 && total_blocks > 0, "tc_pack_weights_multi: bad argument");
  static_assert(sizeof(PackEntry) == 56, "PackEntry layout is part of the C ABI (see include/basi_b200.h)");
  const size_t tab_bytes = (size_t)n_layers * sizeof(PackEntry);
  BASI_CHECK_ARG(tab_bytes <= 40 * 1024, "tc_pack_weights_multi: more than 730 layers in one table");
  int grid = basi::sm_count() * 8;
  if (grid > total_blocks) grid = total_blocks;
  basi::launch(pack_weights_multi_kernel, dim3(grid), dim3(32, 8), tab_bytes, (cudaStream_t)stream,
               (const PackEntry*)table_dev, n_layers, total_blocks);
  BASI_CHECK_LAUNCH("tc_pack_weights_multi");
  return BASI_OK;
}

/* kind FPROP: a = x, b = y (written), w_bf16 = [tap][Cout][Cin]
 * kind DGRAD: a = dy, b = dx (written / accumulated), w_bf16 = [tap][Cin][Cout]
 * kind WGRAD: a = x, b = dy, dw = HWIO float32 gradient (added into) */
}  // extern "C"

static int create_plan(int kind, const basi_conv_desc* d, const basi_tensor* a, const basi_tensor* b,
                       const void* w_bf16, float* dw, int accumulate, int split, basi_tc_conv** out) {
  BASI_CHECK_ARG(d && a && b && out, "tc_conv_create: null argument");
  const basi_tensor* x = (kind == BASI_TC_DGRAD) ? b : a;   // forward input geometry
  const basi_tensor* y = (kind == BASI_TC_DGRAD) ? a : b;   // forward output geometry
#if BASI_H16_FP16
  BASI_CHECK_ARG(!split, "tc_conv_create_split: the split-operand (f32) mode lives in the bfloat16 build of the library");
#endif
  if (split) {
    // operands are bf16 [hi|mid|lo] tensors with 3x the logical channels; destinations are float32
    basi_tensor xl = *x, yl = *y;
    const int px = (kind == BASI_TC_WGRAD && split >= 10) ? split / 10 : split;
    const int pg = (kind == BASI_TC_WGRAD && split >= 10) ? split % 10 : split;
    BASI_CHECK_ARG((px == 2 || px == 3) && (pg == 2 || pg == 3),
                   "tc_conv_create_split: parts must be 2 ([hi|mid]) or 3 ([hi|mid|lo])");
    if (kind == BASI_TC_FPROP || kind == BASI_TC_WGRAD) {
      BASI_CHECK_ARG(a->dtype == BASI_BF16 && a->c % px == 0, "tc_conv_create_split: a must be a bf16 parts tensor");
      xl.c = a->c / px; xl.dtype = BASI_F32; xl.ld = (xl.c + 3) / 4 * 4;
    }
    if (kind == BASI_TC_DGRAD || kind == BASI_TC_WGRAD) {
      const basi_tensor* g = kind == BASI_TC_DGRAD ? a : b;
      BASI_CHECK_ARG(g->dtype == BASI_BF16 && g->c % pg == 0, "tc_conv_create_split: dy must be a bf16 parts tensor");
      yl.c = g->c / pg; yl.dtype = BASI_F32; yl.ld = (yl.c + 3) / 4 * 4;
    }
    BASI_CHECK_ARG((a->ld % 8 == 0) && (((uintptr_t)a->ptr & 15) == 0), "tc_conv_create_split: operand alignment");
    BASI_CHECK_ARG(tc_geometry_ok_split(kind, d, &xl, &yl), "tc_conv_create_split: geometry not supported");
  } else {
    BASI_CHECK_ARG(tc_geometry_ok(kind, d, x, y), "tc_conv_create: geometry not supported by the tcgen05 path");
  }
  basi_tc_conv* pl = new basi_tc_conv();
  memset(pl, 0, sizeof(*pl));
  pl->kind = kind;
  pl->split = split;
  int TW, TH, TN;
  pick_tile(x->n, x->h, x->w, &TW, &TH, &TN);
  int tiles_w = (x->w + TW - 1) / TW, tiles_h = (x->h + TH - 1) / TH, tiles_n = (x->n + TN - 1) / TN;
  int m_tiles = tiles_w * tiles_h * tiles_n;
  const int sms = basi::sm_count();
  int rc = BASI_OK;
  if (kind == BASI_TC_FPROP || kind == BASI_TC_DGRAD) {
    BASI_CHECK_ARG(w_bf16, "tc_conv_create: null weights");
    const basi_tensor* src = a;
    const basi_tensor* dstt = b;
    const int ndim = dstt->c;
    // split mode: three passes over the [hi|mid|lo] operand (all parts x w_hi, hi+mid x w_mid, hi x w_lo)
    // (parts = 3: passes over [hi|mid|lo], [hi|mid], [hi] = six products; parts = 2: passes over [hi|mid], [hi] = the
    // three products hi*hi, mid*hi, hi*mid -- 16 mantissa bits per operand, unbiased 2^-17 representation error)
    const int csrc = split ? src->c / split : src->c;
    const int n0 = split ? (split * csrc + 63) / 64 : (csrc + 63) / 64;
    const int n1 = split ? ((split - 1) * csrc + 63) / 64 : 0, n2 = split == 3 ? (csrc + 63) / 64 : 0;
    const int nbig = split ? (csrc + 63) / 64 : 0;
    const int kchunks_all = n0 + n1 + n2;
    const int kdim = split ? kchunks_all * 64 : src->c;     // weight columns
    int bn = ndim % 128 == 0 ? 128 : (ndim % 64 == 0 ? 64 : 32);
    // (narrower channel tiles for the tiny pyramid-branch maps, to spread the weight traffic over more SMs, measured
    // no change: those launches are 6 us of launch + 6 us of pipeline latency; BASI_TC_SMALL_M=<tiles> re-enables it)
    {
      const int few = exp_env("BASI_TC_SMALL_M") ? atoi(exp_env("BASI_TC_SMALL_M")) : 0;
      while (bn > 32 && ndim % (bn / 2) == 0 && (long)m_tiles * (ndim / bn) < few) bn /= 2;
    }
    const int bn256_mink = exp_env("BASI_TC_BN256_MINK") ? atoi(exp_env("BASI_TC_BN256_MINK")) : 8;
    if (!split && ndim % 256 == 0 && d->kh * d->kw * ((kdim + 63) / 64) >= bn256_mink && !exp_env("BASI_TC_NO_BN256")) {
      const long t128 = ((long)m_tiles * (ndim / 128) + sms - 1) / sms * 10;
      const long t256 = ((long)m_tiles * (ndim / 256) + sms - 1) / sms * 14;
      if (t256 < t128) bn = 256;
    }
    pl->bn = bn;
    // CTA pairs (cta_group::2, M = 256): per CTA 25-50 % fewer shared-memory bytes per flop.  Measured (B=16, 40x40):
    // conv5_4 fprop 250 -> 213 us, dgrad 170 -> 152 us (1593 TFLOP/s), conv5 3x3 41 -> 39 us; the short-K layers
    // (conv1..conv4, <= 18 k-steps per tile) are paced by launch + epilogue and lose 10-60 % to the cluster launch
    // and the lock-step of the two epilogues, so pairs are used for long main loops on 256-wide tiles only.
    {
      const int ksteps = d->kh * d->kw * ((kdim + 63) / 64);
      const char* env_cl = exp_env("BASI_TC_CLUSTER");
      pl->cluster = (m_tiles >= 2 && bn == 256 && ksteps >= 32) ? 2 : 1;
      if (env_cl) pl->cluster = (atoi(env_cl) >= 1 && m_tiles >= 2) ? 2 : 1;   // 1: force pairs, 0: never
    }
    // two pixel tiles per work item (shared weight box) when that still leaves every SM at least ~2/3 of a wave and
    // the main loop is long enough to matter
    pl->mt = 1;
    {
      const int ksteps = d->kh * d->kw * ((kdim + 63) / 64);
      const long items2 = (long)((m_tiles + 1) / 2) * (ndim / bn);
      const char* env_mt = exp_env("BASI_TC_MT");
      int min_k = exp_env("BASI_TC_MT_MINK") ? atoi(exp_env("BASI_TC_MT_MINK")) : 4;
      if (bn <= 128 && pl->cluster == 1 && m_tiles >= 2 && items2 * 3 >= (long)sms * 2 && ksteps >= min_k) pl->mt = 2;
      if (split) pl->mt = 1;     // the epilogue keeps the running fp32 sums of ONE tile in registers
      if (env_mt && atoi(env_mt) == 1) pl->mt = 1;
      if (env_mt && atoi(env_mt) == 2 && bn <= 128 && pl->cluster == 1 && m_tiles >= 2 && !split) pl->mt = 2;   // tests
    }
    // halo mode for 3x3 convolutions (see ConvParams): tile = 8 x 16 pixels of one image, one activation box of
    // 16 x (16 + 2 dil) pixels per 64-channel chunk serves all nine taps
    bool halo = false;
    {
      const char* env_h = exp_env("BASI_TC_HALO");
      const bool can = d->kh == 3 && d->kw == 3 && d->dil >= 1 && d->dil <= 4 && d->pad_t == d->dil &&
                       d->pad_l == d->dil && pl->cluster == 1 && bn <= 128;
      halo = can && env_h && atoi(env_h) == 1 && !split;
    }
    if (halo) {
      TW = 8; TH = 16; TN = 1;
      tiles_w = (x->w + TW - 1) / TW; tiles_h = (x->h + TH - 1) / TH; tiles_n = x->n;
      m_tiles = tiles_w * tiles_h * tiles_n;
      pl->mt = 1;
      rc = make_act_map(&pl->mapA, src, 16, 16 + 2 * d->dil, 1);
    } else {
      rc = make_act_map(&pl->mapA, src, TW, TH, TN);
    }
    if (rc == BASI_OK) rc = make_w_map(&pl->mapB, w_bf16, d->kh * d->kw, ndim, kdim, pl->cluster == 2 ? bn / 2 : bn);
    if (rc == BASI_OK) rc = make_act_map(&pl->mapD, dstt, TW, TH, TN, split ? 32 : (bn >= 64 ? 64 : bn));
    if (rc != BASI_OK) {
      delete pl;
      return rc;
    }
    pl->mapD2 = pl->mapD;
    pl->TW = TW; pl->TH = TH; pl->TN = TN;
    ConvParams& cp = pl->cp;
    cp.TW = TW; cp.TH = TH; cp.TN = TN;
    cp.tiles_w = tiles_w; cp.tiles_h = tiles_h; cp.tiles_n = tiles_n;
    cp.m_tiles = m_tiles; cp.n_tiles = ndim / bn;
    cp.N = dstt->n; cp.H = dstt->h; cp.W = dstt->w; cp.ldd = dstt->ld; cp.Cdst = dstt->c;
    cp.taps = d->kh * d->kw; cp.kw = d->kw; cp.k_chunks = (kdim + 63) / 64;
    cp.a_wrap0 = split ? n0 : 0x7fffffff;
    cp.a_wrap1 = split ? n0 + n1 : 0x7fffffff;
    cp.big_chunks = nbig;
    cp.group_steps = exp_env("BASI_TC_SPLIT_GROUP") ? atoi(exp_env("BASI_TC_SPLIT_GROUP")) : 4;
    if (cp.group_steps < 1) cp.group_steps = 1;
    if (kind == BASI_TC_FPROP) {
      cp.off_h = -d->pad_t; cp.off_w = -d->pad_l; cp.step = d->dil;
    } else {
      cp.off_h = d->pad_t; cp.off_w = d->pad_l; cp.step = -d->dil;
    }
    cp.accumulate = accumulate;
    const int stage_bytes = pl->mt * A_BYTES + bn * 128 / pl->cluster;
    // small-K layers are paced by the epilogue: give them two output staging buffers; large-K layers keep the
    // shared memory for pipeline stages
    cp.out_bufs = (cp.taps * cp.k_chunks <= 8 || pl->mt == 2) ? 2 : 1;
    if (split && bn == 128) cp.out_bufs = 1;             // fp32 staging: 64 KB per 128-channel tile
    const int out_stage = cp.out_bufs * (split ? bn / 32 : (bn + 63) / 64) * A_BYTES;
    const int fixed = 1024 /*align*/ + 1024 /*barriers*/ + 1024 * (int)sizeof(float) + 1024 /*align*/ + out_stage;
    int stages = (227 * 1024 - fixed) / stage_bytes;
    if (stages > 8) stages = 8;
    if (exp_env("BASI_TC_STAGES") && atoi(exp_env("BASI_TC_STAGES")) >= 2 && atoi(exp_env("BASI_TC_STAGES")) < stages)
      stages = atoi(exp_env("BASI_TC_STAGES"));   // experiment: bytes in flight vs main-loop time
    cp.stages = stages;
    cp.ring_bytes = stages * stage_bytes;
    pl->smem = (size_t)stages * stage_bytes + fixed;
    if (halo) {
      cp.halo = 1; cp.dil = d->dil;
      cp.halo_bytes = 16 * (16 + 2 * d->dil) * 128;
      cp.nh = 2;
      const int b_bytes = bn * 128;
      if (cp.k_chunks == 1 && cp.n_tiles == 1 && !exp_env("BASI_TC_HALO_NO_BRES")) {
        // resident weights: the nine boxes once per CTA, the rest of the shared memory for activation halos
        int nh = (227 * 1024 - fixed - cp.taps * b_bytes) / cp.halo_bytes;
        if (nh > 4) nh = 4;
        if (nh >= 2) {
          cp.bres = 1; cp.nh = nh; cp.nb = 1;
          cp.stages = cp.nh + 1;
          cp.ring_bytes = cp.nh * cp.halo_bytes + cp.taps * b_bytes;
          pl->smem = (size_t)cp.ring_bytes + fixed;
        }
      }
      int nb = (227 * 1024 - fixed - cp.nh * cp.halo_bytes) / b_bytes;
      if (nb > 9) nb = 9;
      if (nb < 2 && !cp.bres) {
        delete pl;
        set_error("tc_conv_create: halo mode does not fit in shared memory");
        return BASI_E_INVALID;
      }
      if (!cp.bres) {
        cp.nb = nb;
        cp.stages = cp.nh + cp.nb;
        cp.ring_bytes = cp.nh * cp.halo_bytes + cp.nb * b_bytes;
        pl->smem = (size_t)cp.ring_bytes + fixed;
      }
    }
    pl->dst = (bf16*)dstt->ptr;
    if (pl->mt == 2) {
      const int total = ((cp.m_tiles + 1) / 2) * cp.n_tiles;     // double tiles
      pl->grid = total < sms ? total : sms;
    } else if (pl->cluster == 2) {
      const int total = ((cp.m_tiles + 1) / 2) * cp.n_tiles;     // pairs
      const int pairs = total < sms / 2 ? total : sms / 2;
      pl->grid = 2 * pairs;
    } else {
      const int total = cp.m_tiles * cp.n_tiles;
      pl->grid = total < sms ? total : sms;
    }
    if (exp_env("BASI_TC_DEBUG_EMPTY")) cp.m_tiles = 0;   // timing experiment: prologue + teardown only
    cp.debug = exp_env("BASI_TC_DEBUG_STATS") ? atoi(exp_env("BASI_TC_DEBUG_STATS")) : 0;
  } else {
    BASI_CHECK_ARG(dw, "tc_conv_create: null dw");
    // (split mode, weight gradient: parts may be 10 * x_parts + dy_parts when the two operands were split differently)
    const int pa = split >= 10 ? split / 10 : split, pb = split >= 10 ? split % 10 : split;
    const int cin = split ? a->c / pa : a->c, cout = split ? b->c / pb : b->c;
    int bn = cin > 64 ? 128 : 64;            // ci tile (UMMA N); co is the UMMA M = 128
    pl->bn = bn;
    rc = make_act_map(&pl->mapA, a, TW, TH, TN);
    if (rc == BASI_OK) rc = make_act_map(&pl->mapB, b, TW, TH, TN);
    if (rc != BASI_OK) {
      delete pl;
      return rc;
    }
    WgradParams& wp = pl->wp;
    wp.TW = TW; wp.TH = TH; wp.TN = TN;
    wp.tiles_w = tiles_w; wp.tiles_h = tiles_h; wp.tiles_n = tiles_n; wp.m_tiles = m_tiles;
    wp.taps = d->kh * d->kw; wp.kw = d->kw;
    wp.ci_tiles = (cin + bn - 1) / bn; wp.co_tiles = (cout + 127) / 128;
    wp.off_h = -d->pad_t; wp.off_w = -d->pad_l; wp.step = d->dil;
    wp.Cin = cin; wp.Cout = cout;
    wp.nterms = split == 0 ? 1 : ((pa == 3 && pb == 3) ? 6 : 3);
    const int out_tiles = wp.taps * wp.ci_tiles * wp.co_tiles;
    // split-K over pixel tiles.  Every split adds |dW| fp32 atomics and every CTA pays its prologue / pipeline fill /
    // epilogue, so the 16-bit plans aim at ~64 CTAs per launch: weight gradients of neighbouring layers run on two
    // side streams and share the GPU (measured on cfg3: 9.13 -> 8.81 ms/step against one full wave per launch;
    // 48 / 74 / 96 CTAs: 8.95 / 8.88 / 8.81, tools/wgrad_sched.py).  The split-operand (float32-grade) plans keep
    // the round-1 rule: a single wave for tiny gradients, two waves for long pixel loops (the new rule is neutral there).
    int splits;
    if (exp_env("BASI_TC_WGRAD_SPLITS")) {
      splits = atoi(exp_env("BASI_TC_WGRAD_SPLITS"));                                      // experiment
    } else if (split == 0 && !exp_env("BASI_TC_WGRAD_WAVES")) {
      const int target = exp_env("BASI_TC_WGRAD_TARGET") ? atoi(exp_env("BASI_TC_WGRAD_TARGET"))
                         : (basi::wgrad_cta_target() > 0 ? basi::wgrad_cta_target() : 64);
      splits = (target + out_tiles - 1) / out_tiles;
      if (exp_env("BASI_TC_WGRAD_MAXTILES")) {  // experiment: no CTA loops over more than this many pixel tiles
        const int mx = atoi(exp_env("BASI_TC_WGRAD_MAXTILES"));
        if (mx > 0 && (m_tiles + mx - 1) / mx > splits) splits = (m_tiles + mx - 1) / mx;
      }
    } else {
      int waves = 1;
      const int s1 = (sms + out_tiles - 1) / out_tiles;
      if (m_tiles / (s1 > 0 ? s1 : 1) > 10) waves = 2;
      if (exp_env("BASI_TC_WGRAD_WAVES")) waves = atoi(exp_env("BASI_TC_WGRAD_WAVES"));
      splits = (waves * sms + out_tiles - 1) / out_tiles;
    }
    if (splits > m_tiles) splits = m_tiles;
    if (splits < 1) splits = 1;
    wp.tiles_per_split = (m_tiles + splits - 1) / splits;
    wp.splits = (m_tiles + wp.tiles_per_split - 1) / wp.tiles_per_split;
    pl->gbox = (cout <= 64 && !exp_env("BASI_TC_WGRAD_GBOX2")) ? 1 : 2;
    const int stage_bytes = pl->gbox * A_BYTES + (bn / 64) * A_BYTES;
    int stages = (int)((227 * 1024 - 2048) / stage_bytes);
    if (stages > 6) stages = 6;
    if (exp_env("BASI_TC_WGRAD_STAGES") && atoi(exp_env("BASI_TC_WGRAD_STAGES")) >= 2 &&
        atoi(exp_env("BASI_TC_WGRAD_STAGES")) < stages)
      stages = atoi(exp_env("BASI_TC_WGRAD_STAGES"));   // experiment: smaller footprint -> co-residency with BN kernels
    wp.stages = stages;
    pl->smem = (size_t)stages * stage_bytes + 1024 + 256;
    pl->dw = dw;
    pl->grid = out_tiles * wp.splits;
  }
  *out = pl;
  return BASI_OK;
}

extern "C" {

int basi_tc_conv_create(int kind, const basi_conv_desc* d, const basi_tensor* a, const basi_tensor* b,
                        const void* w_bf16, float* dw, int accumulate, basi_tc_conv** out) {
  return create_plan(kind, d, a, b, w_bf16, dw, accumulate, 0, out);
}

/* Split-operand (fp32-grade) plans: operands are bf16 [hi|mid|lo] tensors produced by basi_split3_bf16 (3C channels),
 * weights are the packed split layout of basi_tc_pack_weights_multi (mode 1), destinations are float32.
 *   FPROP: a = x3, b = y (f32, written)      DGRAD: a = dy3, b = dx (f32, written / accumulated)
 *   WGRAD: a = x3, b = dy3, dw = HWIO float32 gradient (added into) */
int basi_tc_conv_create_split(int kind, const basi_conv_desc* d, const basi_tensor* a, const basi_tensor* b,
                              const void* w_split, float* dw, int accumulate, int parts, basi_tc_conv** out) {
  return create_plan(kind, d, a, b, w_split, dw, accumulate, parts, out);
}

int basi_tc_conv_supported_split(int kind, const basi_conv_desc* d, const basi_tensor* x, const basi_tensor* y) {
  return tc_geometry_ok_split(kind, d, x, y) ? 1 : 0;
}

/* number of bf16 weight columns (K') of the split layout for a reduced channel count c: three passes of 64-channel
 * chunks over [hi|mid|lo] (3c), [hi|mid] (2c) and [hi] (c) */
int basi_tc_split_kcols(int c, int parts) {
  if (parts == 2) return (((2 * c + 63) / 64) + ((c + 63) / 64)) * 64;
  return (((3 * c + 63) / 64) + ((2 * c + 63) / 64) + ((c + 63) / 64)) * 64;
}

int basi_tc_conv_set_bn_stats(basi_tc_conv* pl, double* sums, const float* gamma, const float* beta, double count,
                              float eps, float* bnp, uint32_t* counter) {
  BASI_CHECK_ARG(pl && pl->kind == BASI_TC_FPROP, "tc_conv_set_bn_stats: needs an fprop plan");
  BASI_CHECK_ARG(sums && counter && (!bnp || (gamma && beta && count > 0)), "tc_conv_set_bn_stats: bad argument");
  pl->cp.bn_sums = sums;
  pl->cp.bn_gamma = gamma;
  pl->cp.bn_beta = beta;
  pl->cp.bn_bnp = bnp;
  pl->cp.bn_counter = counter;
  pl->cp.bn_count = count;
  pl->cp.bn_eps = eps;
  return BASI_OK;
}

/* Fused BN apply for an fprop plan that already has basi_tc_conv_set_bn_stats: the kernel additionally writes
 * out = [relu](gamma * (y - mean) * istd + beta), normalised from the fp32 accumulators, after a grid barrier.
 * Returns 1 when the plan was switched to the fused mode, 0 when this layer cannot be fused (accumulators of a CTA's
 * work items exceed the 512 TMEM columns, CTA pairs, halo mode, fp32 output) -- the caller then runs basi_bn_apply. */
int basi_tc_conv_set_bn_apply(basi_tc_conv* pl, const basi_tensor* out, int relu) {
  if (!pl || !out || pl->kind != BASI_TC_FPROP || !pl->cp.bn_sums || !pl->cp.bn_bnp) return 0;
  if (pl->split || pl->cluster != 1 || pl->cp.halo || pl->cp.accumulate) return 0;
  if (exp_env("BASI_TC_NO_FUSED_APPLY")) return 0;
  if (out->dtype != BASI_BF16 || out->c != pl->cp.Cdst || out->n != pl->cp.N || out->h != pl->cp.H || out->w != pl->cp.W)
    return 0;
  if (out->ld % 8 || ((uintptr_t)out->ptr & 15)) return 0;
  const ConvParams& cp = pl->cp;
  const int items = ((cp.m_tiles + pl->mt - 1) / pl->mt) * cp.n_tiles;
  if (items <= 0 || pl->grid <= 0) return 0;
  const int per_cta = (items + pl->grid - 1) / pl->grid;
  if (per_cta * pl->mt * pl->bn > 512 || per_cta > 16) return 0;     // TMEM columns; one tmem_full barrier per item
  if (pl->grid > basi::sm_count()) return 0;
  if (make_act_map(&pl->mapD2, out, pl->TW, pl->TH, pl->TN, pl->bn >= 64 ? 64 : pl->bn) != BASI_OK) return 0;
  pl->cp.fuse_apply = 1;
  pl->cp.keep_acc = 1;
  pl->cp.apply_relu = relu ? 1 : 0;
  return 1;
}

/* Fused BN backward for a dgrad plan whose destination `dx` is the gradient wrt the RAW conv output x of the layer
 * that feeds this convolution through BN(+ReLU): the kernel computes dA in TMEM, reduces sum(g), sum(g * xhat) with
 * the ReLU mask recomputed from x, meets the grid at a barrier and writes dx = BN'(dA) directly; dgamma / dbeta are
 * added into their gradient slots.  Returns 1 if the plan was switched, 0 if it cannot be fused. */
int basi_tc_conv_set_bn_bwd(basi_tc_conv* pl, const basi_tensor* x, const float* bnp, int relu, double* dsums,
                            double count, float* dgamma, float* dbeta, uint32_t* counter) {
  if (!pl || !x || !bnp || !dsums || !dgamma || !dbeta || !counter || count <= 0) return 0;
  if (pl->kind != BASI_TC_DGRAD || pl->split || pl->cluster != 1 || pl->cp.halo || pl->cp.accumulate) return 0;
  if (pl->bn > 128 || exp_env("BASI_TC_NO_FUSED_BWD")) return 0;
  ConvParams& cp = pl->cp;
  if (x->dtype != BASI_BF16 || x->c != cp.Cdst || x->n != cp.N || x->h != cp.H || x->w != cp.W) return 0;
  if (x->ld % 8 || ((uintptr_t)x->ptr & 15)) return 0;
  const int items = ((cp.m_tiles + pl->mt - 1) / pl->mt) * cp.n_tiles;
  if (items <= 0 || pl->grid <= 0 || pl->grid > basi::sm_count()) return 0;
  const int per_cta = (items + pl->grid - 1) / pl->grid;
  if (per_cta * pl->mt * pl->bn > 512 || per_cta > 16) return 0;
  // shared memory: one output staging buffer + the x tile buffer + the coefficient table + one mbarrier
  const int nbox = (pl->bn + 63) / 64;
  const int staging = nbox * A_BYTES;
  const int out_area = 2 * staging + 5 * pl->bn * 4 + 64;
  const int fixed = 1024 + 1024 + 1024 * (int)sizeof(float) + 1024 + out_area;
  const int stage_bytes = pl->mt * A_BYTES + pl->bn * 128;
  int stages = (227 * 1024 - fixed) / stage_bytes;
  if (stages > 8) stages = 8;
  if (stages < 2) return 0;
  if (make_act_map(&pl->mapD2, x, pl->TW, pl->TH, pl->TN, pl->bn >= 64 ? 64 : pl->bn) != BASI_OK) return 0;
  cp.stages = stages;
  cp.ring_bytes = stages * stage_bytes;
  cp.out_bufs = 1;
  pl->smem = (size_t)cp.ring_bytes + fixed;
  cp.xbuf_off = staging;
  cp.coef_off = 2 * staging;
  cp.xbar_off = 2 * staging + 5 * pl->bn * 4;
  cp.fuse_bwd = 1;
  cp.keep_acc = 1;
  cp.bwd_relu = relu ? 1 : 0;
  cp.bwd_bnp = bnp;
  cp.bwd_sums = dsums;
  cp.bwd_dgamma = dgamma;
  cp.bwd_dbeta = dbeta;
  cp.bwd_count = count;
  cp.bwd_counter = counter;
  return 1;
}

/* fprop plans: y = [relu](conv(x, w) + bias) -- the slim conv2d of vgg_16 (bias + ReLU, no batch norm) */
int basi_tc_conv_set_bias(basi_tc_conv* pl, const float* bias, int relu) {
  BASI_CHECK_ARG(pl && pl->kind == BASI_TC_FPROP && bias && !pl->split, "tc_conv_set_bias: needs a bf16 fprop plan");
  pl->cp.bias = bias;
  pl->cp.bias_relu = relu ? 1 : 0;
  return BASI_OK;
}

int basi_tc_conv_run(basi_tc_conv* pl, void* stream) {
  BASI_CHECK_ARG(pl, "tc_conv_run: null plan");
  cudaStream_t st = (cudaStream_t)stream;
  if (pl->kind == BASI_TC_WGRAD) {
    if (pl->bn == 128) launch_wgrad<128>(pl, st);
    else launch_wgrad<64>(pl, st);
  } else {
    if (pl->bn == 256) launch_conv<256>(pl, st);
    else if (pl->bn == 128) launch_conv<128>(pl, st);
    else if (pl->bn == 64) launch_conv<64>(pl, st);
    else launch_conv<32>(pl, st);
  }
  BASI_CHECK_LAUNCH("tc_conv_run");
  if (pl->kind != BASI_TC_WGRAD && pl->cp.debug == 30) {
    unsigned long long h[16];
    cudaStreamSynchronize(st);
    cudaMemcpyFromSymbol(h, g_tc_dbg, sizeof(h));
    const double n = h[1] ? (double)h[1] : 1.0;
    fprintf(stderr, "tc debug (CTA 0, %llu tiles, cycles per tile): epilogue waits accumulator %.0f | waits store-read %.0f | "
            "epilogue work %.0f || MMA waits epilogue %.0f | MMA waits operands %.0f\n",
            h[1], h[0] / n, h[2] / n, h[3] / n, h[4] / n, h[5] / n);
    unsigned long long z[16] = {0};
    cudaMemcpyToSymbol(g_tc_dbg, z, sizeof(z));
  }
  return BASI_OK;
}

void basi_tc_conv_destroy(basi_tc_conv* pl) { delete pl; }

}  // extern "C"
