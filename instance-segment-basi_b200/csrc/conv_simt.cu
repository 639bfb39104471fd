// CUDA-core implicit-GEMM convolutions (fp32 accumulate) for every geometry on the BAIS PSPNet
// path: fprop, dgrad (adjoint w.r.t. input) and wgrad (adjoint w.r.t. HWIO weights), plus the
// skinny GEMMs of the attention-class head.  This is the exact-fp32 path (parity mode) and the
// fallback for layers the tcgen05 path does not take (Cin=4 stem, strided 1x1, tiny heads).
#include <stdlib.h>

#include "common.cuh"

namespace basi {

struct GemmConv {
  // gather-source tensor S and destination tensor D of the implicit GEMM
  const void* S;
  void* D;
  const float* W;     // HWIO
  const float* bias;  // fprop only
  int N, SH, SW, SC, lds;
  int DH, DW, DC, ldd;
  int Cin, Cout;  // of the convolution (weights)
  int kh, kw, stride, dil, pad_t, pad_l;
  int relu, accumulate;
  int M, K;       // GEMM sizes: M = N*DH*DW, K = taps*SC
  int fastA;      // SC % 16 == 0 and aligned -> vector loads within one tap
  int fastB;
  int fastD;      // vector stores in the epilogue
};

enum { MODE_FPROP = 0, MODE_DGRAD = 1 };

constexpr int BM = 128, BN = 64, BK = 16, BNP = BN + 4;

template <typename T>
__device__ __forceinline__ void load8(const T* p, float (&o)[8]);
template <>
__device__ __forceinline__ void load8<float>(const float* p, float (&o)[8]) {
  float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
}
template <>
__device__ __forceinline__ void load8<bf16>(const bf16* p, float (&o)[8]) {
  Vec<bf16> v = Vec<bf16>::load(p);
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i] = v.v[i];
}
template <typename T>
__device__ __forceinline__ void load4(const T* p, float (&o)[4]);
template <>
__device__ __forceinline__ void load4<float>(const float* p, float (&o)[4]) {
  float4 a = *reinterpret_cast<const float4*>(p);
  o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w;
}
template <>
__device__ __forceinline__ void load4<bf16>(const bf16* p, float (&o)[4]) {
  uint2 t = *reinterpret_cast<const uint2*>(p);
  o[0] = h16_lo(t.x); o[1] = h16_hi(t.x);
  o[2] = h16_lo(t.y); o[3] = h16_hi(t.y);
}

template <typename T>
__device__ __forceinline__ void store4(T* p, const float (&v)[4]);
template <>
__device__ __forceinline__ void store4<float>(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <>
__device__ __forceinline__ void store4<bf16>(bf16* p, const float (&v)[4]) {
  *reinterpret_cast<uint2*>(p) = make_uint2(h16_pack(v[0], v[1]), h16_pack(v[2], v[3]));
}

// source coordinate of (dest coordinate d, tap offset r) along one axis; returns false if invalid
template <int MODE>
__device__ __forceinline__ bool src_coord(int d, int r, int stride, int dil, int pad, int S, int& s) {
  if (MODE == MODE_FPROP) {
    s = d * stride + r * dil - pad;
    return s >= 0 && s < S;
  } else {
    int t = d + pad - r * dil;
    if (t < 0) return false;
    if (stride == 1) {
      s = t;
    } else {
      if (t % stride) return false;
      s = t / stride;
    }
    return s < S;
  }
}

template <int MODE, typename TS, typename TD>
__global__ void __launch_bounds__(256) conv_gemm_kernel(const GemmConv g) {
  pdl_prologue();
  __shared__ __align__(16) float As[2][BK][BM];
  __shared__ __align__(16) float Bs[2][BK][BNP];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const TS* __restrict__ S = (const TS*)g.S;
  const float* __restrict__ W = g.W;

  // A-load role: one destination pixel per thread, 8 consecutive k
  const int am = tid & (BM - 1), akh = (tid >> 7) * 8;
  const int gm = m0 + am;
  int pn = 0, pdh = 0, pdw = 0;
  const bool mvalid = gm < g.M;
  if (mvalid) {
    pdw = gm % g.DW;
    int t = gm / g.DW;
    pdh = t % g.DH;
    pn = t / g.DH;
  }
  // B-load role
  const int bk = MODE == MODE_FPROP ? (tid >> 4) : (tid & 3) * 4;
  const int bn = MODE == MODE_FPROP ? (tid & 15) * 4 : (tid >> 2);

  float ra[8], rb[4];
  auto load_tile = [&](int k0) {
    // ---- A ----
    if (g.fastA) {
      const int tap = k0 / g.SC, c0 = k0 - tap * g.SC + akh;
      const int r = tap / g.kw, s = tap - r * g.kw;
      int sh, sw;
      bool ok = mvalid && src_coord<MODE>(pdh, r, g.stride, g.dil, g.pad_t, g.SH, sh) &&
                src_coord<MODE>(pdw, s, g.stride, g.dil, g.pad_l, g.SW, sw);
      if (ok) {
        load8<TS>(S + (((int64_t)pn * g.SH + sh) * g.SW + sw) * g.lds + c0, ra);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) ra[j] = 0.f;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int k = k0 + akh + j;
        float v = 0.f;
        if (mvalid && k < g.K) {
          const int tap = k / g.SC, c = k - tap * g.SC;
          const int r = tap / g.kw, s = tap - r * g.kw;
          int sh, sw;
          if (src_coord<MODE>(pdh, r, g.stride, g.dil, g.pad_t, g.SH, sh) &&
              src_coord<MODE>(pdw, s, g.stride, g.dil, g.pad_l, g.SW, sw))
            v = to_f32(S[(((int64_t)pn * g.SH + sh) * g.SW + sw) * g.lds + c]);
        }
        ra[j] = v;
      }
    }
    // ---- B ----
    if (MODE == MODE_FPROP) {
      const int k = k0 + bk, n = n0 + bn;
      if (g.fastB && k < g.K && n + 3 < g.DC) {
        float4 t = *reinterpret_cast<const float4*>(W + (int64_t)k * g.Cout + n);
        rb[0] = t.x; rb[1] = t.y; rb[2] = t.z; rb[3] = t.w;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) rb[j] = (k < g.K && n + j < g.DC) ? W[(int64_t)k * g.Cout + n + j] : 0.f;
      }
    } else {
      // B[k=(tap,co)][n=ci] = W[(tap*Cin + ci)*Cout + co]; 4 consecutive co for one ci
      const int k = k0 + bk, n = n0 + bn;
      if (g.fastB && n < g.DC && k + 3 < g.K) {
        const int tap = k / g.SC, co = k - tap * g.SC;
        float4 t = *reinterpret_cast<const float4*>(W + ((int64_t)tap * g.Cin + n) * g.Cout + co);
        rb[0] = t.x; rb[1] = t.y; rb[2] = t.z; rb[3] = t.w;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int kk = k + j;
          float v = 0.f;
          if (n < g.DC && kk < g.K) {
            const int tap = kk / g.SC, co = kk - tap * g.SC;
            v = W[((int64_t)tap * g.Cin + n) * g.Cout + co];
          }
          rb[j] = v;
        }
      }
    }
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int j = 0; j < 8; ++j) As[buf][akh + j][am] = ra[j];
    if (MODE == MODE_FPROP) {
      *reinterpret_cast<float4*>(&Bs[buf][bk][bn]) = make_float4(rb[0], rb[1], rb[2], rb[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) Bs[buf][bk + j][bn] = rb[j];
    }
  };

  const int tm = tid >> 4, tn = tid & 15;
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int ktiles = (g.K + BK - 1) / BK;
  load_tile(0);
  store_tile(0);
  __syncthreads();
  for (int kt = 0; kt < ktiles; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < ktiles) load_tile((kt + 1) * BK);
    // two-level summation: a fresh partial per k-tile keeps the fp32 error growth at sqrt(K/16)
    float part[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) part[i][j] = 0.f;
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][tm * 8]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][tm * 8 + 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[buf][k][tn * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) part[i][j] = fmaf(av[i], bv[j], part[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] += part[i][j];
    if (kt + 1 < ktiles) {
      store_tile(buf ^ 1);
      __syncthreads();
    }
  }

  // ---- epilogue ----
  TD* __restrict__ D = (TD*)g.D;
  const int nc = n0 + tn * 4;
  float bvals[4] = {0.f, 0.f, 0.f, 0.f};
  if (g.bias) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (nc + j < g.DC) bvals[j] = g.bias[nc + j];
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + tm * 8 + i;
    if (m >= g.M) continue;
    TD* dst = D + (int64_t)m * g.ldd + nc;
    if (g.fastD && nc + 3 < g.DC) {
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = acc[i][j] + bvals[j];
      if (g.accumulate) {
        float o[4];
        load4<TD>(dst, o);
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] += o[j];
      }
      if (g.relu) {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = fmaxf(v[j], 0.f);
      }
      store4<TD>(dst, v);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (nc + j >= g.DC) continue;
        float v = acc[i][j] + bvals[j];
        if (g.accumulate) v += to_f32(dst[j]);
        if (g.relu) v = fmaxf(v, 0.f);
        dst[j] = from_f32<TD>(v);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// wgrad: dW[kk=(tap,ci)][co] += sum_m A[m][kk] * dY[m][co], split over m across blockIdx.z
// ------------------------------------------------------------------------------------------
struct WgradConv {
  const void* X;
  const void* G;
  float* dW;
  int N, IH, IW, Cin, ldx;
  int OH, OW, Cout, ldg;
  int kh, kw, stride, dil, pad_t, pad_l;
  int M, K;
  int m_per_split;
  int fastA, fastG;
};

constexpr int WK = 64, WN = 64, WP = 16, WS = 68;

template <typename TX, typename TG>
__global__ void __launch_bounds__(256) conv_wgrad_kernel(const WgradConv g) {
  pdl_prologue();
  __shared__ __align__(16) float As[2][WP][WS];
  __shared__ __align__(16) float Gs[2][WP][WS];
  const int tid = threadIdx.x;
  const int k0 = blockIdx.x * WK, n0 = blockIdx.y * WN;
  const int mbeg = blockIdx.z * g.m_per_split;
  const int mend = min(g.M, mbeg + g.m_per_split);
  const TX* __restrict__ X = (const TX*)g.X;
  const TG* __restrict__ G = (const TG*)g.G;

  const int lp = tid >> 4, lq = (tid & 15) * 4;
  // fixed tap / channel of this thread's 4 A columns (fast path)
  const int kk = k0 + lq;
  int tr = 0, ts = 0, tci = 0;
  if (kk < g.K) {
    int tap = kk / g.Cin;
    tci = kk - tap * g.Cin;
    tr = tap / g.kw;
    ts = tap - tr * g.kw;
  }
  float ra[4], rg[4];
  auto load_chunk = [&](int mb) {
    const int m = mb + lp;
    const bool mv = m < mend;
    int ow = 0, oh = 0, n = 0;
    if (mv) {
      ow = m % g.OW;
      int t = m / g.OW;
      oh = t % g.OH;
      n = t / g.OH;
    }
    if (g.fastA) {
      const int ih = oh * g.stride + tr * g.dil - g.pad_t, iw = ow * g.stride + ts * g.dil - g.pad_l;
      if (mv && kk < g.K && ih >= 0 && ih < g.IH && iw >= 0 && iw < g.IW) {
        load4<TX>(X + (((int64_t)n * g.IH + ih) * g.IW + iw) * g.ldx + tci, ra);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) ra[j] = 0.f;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float v = 0.f;
        const int k = kk + j;
        if (mv && k < g.K) {
          const int tap = k / g.Cin, ci = k - tap * g.Cin;
          const int r = tap / g.kw, s = tap - r * g.kw;
          const int ih = oh * g.stride + r * g.dil - g.pad_t, iw = ow * g.stride + s * g.dil - g.pad_l;
          if (ih >= 0 && ih < g.IH && iw >= 0 && iw < g.IW)
            v = to_f32(X[(((int64_t)n * g.IH + ih) * g.IW + iw) * g.ldx + ci]);
        }
        ra[j] = v;
      }
    }
    const int nn = n0 + lq;
    if (g.fastG && mv && nn + 3 < g.Cout) {
      load4<TG>(G + (int64_t)m * g.ldg + nn, rg);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) rg[j] = (mv && nn + j < g.Cout) ? to_f32(G[(int64_t)m * g.ldg + nn + j]) : 0.f;
    }
  };
  auto store_chunk = [&](int buf) {
    *reinterpret_cast<float4*>(&As[buf][lp][lq]) = make_float4(ra[0], ra[1], ra[2], ra[3]);
    *reinterpret_cast<float4*>(&Gs[buf][lp][lq]) = make_float4(rg[0], rg[1], rg[2], rg[3]);
  };

  const int tk = tid >> 4, tn = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int chunks = (mend - mbeg + WP - 1) / WP;
  if (chunks > 0) {
    load_chunk(mbeg);
    store_chunk(0);
    __syncthreads();
    for (int c = 0; c < chunks; ++c) {
      const int buf = c & 1;
      if (c + 1 < chunks) load_chunk(mbeg + (c + 1) * WP);
      float part[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) part[i][j] = 0.f;
#pragma unroll
      for (int p = 0; p < WP; ++p) {
        float4 a = *reinterpret_cast<const float4*>(&As[buf][p][tk * 4]);
        float4 b = *reinterpret_cast<const float4*>(&Gs[buf][p][tn * 4]);
        const float av[4] = {a.x, a.y, a.z, a.w};
        const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) part[i][j] = fmaf(av[i], bv[j], part[i][j]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += part[i][j];
      if (c + 1 < chunks) {
        store_chunk(buf ^ 1);
        __syncthreads();
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int k = k0 + tk * 4 + i;
    if (k >= g.K) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tn * 4 + j;
      if (n < g.Cout) atomicAdd(g.dW + (int64_t)k * g.Cout + n, acc[i][j]);
    }
  }
}

// column sums: out[n] += sum_m G[m][n]
template <typename TG>
__global__ void colsum_kernel(const TG* __restrict__ G, int64_t M, int C, int ld, float* __restrict__ out) {
  pdl_prologue();
  __shared__ float sm[8][33];
  const int n = blockIdx.x * 32 + threadIdx.x;
  float a = 0.f;
  if (n < C)
    for (int64_t m = (int64_t)blockIdx.y * 8 + threadIdx.y; m < M; m += (int64_t)gridDim.y * 8)
      a += to_f32(G[m * ld + n]);
  sm[threadIdx.y][threadIdx.x] = a;
  __syncthreads();
  if (threadIdx.y == 0 && n < C) {
    for (int y = 1; y < 8; ++y) a += sm[y][threadIdx.x];
    atomicAdd(out + n, a);
  }
}

// ------------------------------------------------------------------------------------------
// Skinny GEMMs (M <= 64 rows): attention-class head
// ------------------------------------------------------------------------------------------
constexpr int SK_KT = 64;  // k-slab per block (fwd / wgrad)

template <typename TA>
__global__ void __launch_bounds__(128) skinny_fwd_kernel(const TA* __restrict__ a, int64_t lda,
                                                         const float* __restrict__ w, float* __restrict__ y, int M,
                                                         int K, int N, int slabs_per_block,
                                                         float* __restrict__ part) {
  // block = (128 output columns, a run of k-slabs); the partial sums of the run stay in registers and are added to y
  // once (with one slab per block the 52 MB class_attention_conv GEMV was bound by its 3.3 M atomics, not by HBM)
  pdl_prologue();
  __shared__ __align__(16) float sa[16][SK_KT + 4];      // (row stride 68 floats: 16-byte aligned rows, shifted banks)
  const int n = blockIdx.x * 128 + threadIdx.x;
  for (int mb = 0; mb < M; mb += 16) {
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0.f;
    for (int sl = 0; sl < slabs_per_block; ++sl) {
      const int k0 = (blockIdx.y * slabs_per_block + sl) * SK_KT;
      if (k0 >= K) break;
      const int kt = min(SK_KT, K - k0);
      __syncthreads();
      for (int i = threadIdx.x; i < 16 * SK_KT; i += 128) {   // rows >= M are zero-filled (they are multiplied)
        const int m = i / SK_KT, k = i - m * SK_KT;
        sa[m][k] = (mb + m < M && k < kt) ? to_f32(a[(int64_t)(mb + m) * lda + k0 + k]) : 0.f;
      }
      __syncthreads();
      if (n < N) {
        // four weight rows per step: 4 coalesced loads in flight, one 128-bit (broadcast) shared-memory read per batch
        // row feeds 4 FMAs (the scalar version issued one shared-memory load per FMA)
        int k = 0;
        for (; k + 8 <= kt; k += 8) {
          const float* wp = w + (int64_t)(k0 + k) * N + n;
          float wv[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) wv[j] = __ldg(wp + (int64_t)j * N);       // eight rows in flight
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float4 a0 = *reinterpret_cast<const float4*>(&sa[i][k]);
            const float4 a1 = *reinterpret_cast<const float4*>(&sa[i][k + 4]);
            float t = fmaf(a0.x, wv[0], fmaf(a0.y, wv[1], fmaf(a0.z, wv[2], fmaf(a0.w, wv[3], acc[i]))));
            acc[i] = fmaf(a1.x, wv[4], fmaf(a1.y, wv[5], fmaf(a1.z, wv[6], fmaf(a1.w, wv[7], t))));
          }
        }
        for (; k < kt; ++k) {
          const float wv = __ldg(w + (int64_t)(k0 + k) * N + n);
#pragma unroll
          for (int i = 0; i < 16; ++i) acc[i] = fmaf(sa[i][k], wv, acc[i]);
        }
      }
    }
    if (n < N) {
      // part != nullptr: this block's partial sums go to their own slice [blockIdx.y][M][N] and are added up in a fixed
      // order by skinny_reduce_bias_act_kernel (deterministic); else fp32 atomics into y (order-dependent rounding)
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (mb + i < M) {
          if (part) part[((int64_t)blockIdx.y * M + mb + i) * N + n] = acc[i];
          else atomicAdd(y + (int64_t)(mb + i) * N + n, acc[i]);
        }
    }
  }
}

// Vectorised forward GEMV for N % 4 == 0 (class_attention_conv: a 25600 x 512 float32 matrix, 52 MB, streamed once).
// A thread owns FOUR consecutive columns and keeps eight 128-bit weight loads in flight; a block (128 threads = 512
// columns) owns one contiguous run of at most SK4_KB k rows whose 16 x run activations are staged in shared memory ONCE,
// so nothing but the weight stream is on the critical path (the scalar kernel above stalled at a barrier + a dependent
// activation load every 64 rows and had 4 bytes per load in flight: 1.2 TB/s).
constexpr int SK4_KB = 192;

template <typename TA>
__global__ void __launch_bounds__(128) skinny_fwd4_kernel(const TA* __restrict__ a, int64_t lda,
                                                          const float* __restrict__ w, float* __restrict__ y, int M,
                                                          int K, int N, int k_per_block, float* __restrict__ part) {
  pdl_prologue();
  __shared__ __align__(16) float sa[16][SK4_KB + 4];
  const int n4 = (blockIdx.x * 128 + threadIdx.x) * 4;
  const int kbeg = blockIdx.y * k_per_block;
  const int kend = min(K, kbeg + k_per_block);
  const int kn = kend - kbeg;               // 1 .. SK4_KB (the launch geometry leaves no empty block)
  const int kpad = (kn + 7) & ~7;           // rows kn .. kpad-1 are multiplied by zero-filled activations
  for (int mb = 0; mb < M; mb += 16) {
    __syncthreads();
    for (int i = threadIdx.x; i < 16 * kpad; i += 128) {
      const int m = i / kpad, k = i - m * kpad;
      sa[m][k] = (mb + m < M && k < kn) ? to_f32(a[(int64_t)(mb + m) * lda + kbeg + k]) : 0.f;
    }
    __syncthreads();
    if (n4 >= N) continue;
    float4 acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 wv[8], wn[8];
    auto load8 = [&](int k, float4 (&x)[8]) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int kk = min(kbeg + k + j, kend - 1);
        x[j] = __ldg(reinterpret_cast<const float4*>(w + (int64_t)kk * N + n4));
      }
    };
    load8(0, wv);
    for (int k = 0; k < kpad; k += 8) {
      if (k + 8 < kpad) load8(k + 8, wn);       // the next eight rows are in flight while these are multiplied
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float4 a0 = *reinterpret_cast<const float4*>(&sa[i][k]);
        const float4 a1 = *reinterpret_cast<const float4*>(&sa[i][k + 4]);
        float4 t = acc[i];
        t.x = fmaf(a0.x, wv[0].x, fmaf(a0.y, wv[1].x, fmaf(a0.z, wv[2].x, fmaf(a0.w, wv[3].x, t.x))));
        t.y = fmaf(a0.x, wv[0].y, fmaf(a0.y, wv[1].y, fmaf(a0.z, wv[2].y, fmaf(a0.w, wv[3].y, t.y))));
        t.z = fmaf(a0.x, wv[0].z, fmaf(a0.y, wv[1].z, fmaf(a0.z, wv[2].z, fmaf(a0.w, wv[3].z, t.z))));
        t.w = fmaf(a0.x, wv[0].w, fmaf(a0.y, wv[1].w, fmaf(a0.z, wv[2].w, fmaf(a0.w, wv[3].w, t.w))));
        t.x = fmaf(a1.x, wv[4].x, fmaf(a1.y, wv[5].x, fmaf(a1.z, wv[6].x, fmaf(a1.w, wv[7].x, t.x))));
        t.y = fmaf(a1.x, wv[4].y, fmaf(a1.y, wv[5].y, fmaf(a1.z, wv[6].y, fmaf(a1.w, wv[7].y, t.y))));
        t.z = fmaf(a1.x, wv[4].z, fmaf(a1.y, wv[5].z, fmaf(a1.z, wv[6].z, fmaf(a1.w, wv[7].z, t.z))));
        t.w = fmaf(a1.x, wv[4].w, fmaf(a1.y, wv[5].w, fmaf(a1.z, wv[6].w, fmaf(a1.w, wv[7].w, t.w))));
        acc[i] = t;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) wv[j] = wn[j];
    }
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (mb + i < M) {
        if (part) {
          *reinterpret_cast<float4*>(part + ((int64_t)blockIdx.y * M + mb + i) * N + n4) = acc[i];
        } else {
          float* yp = y + (int64_t)(mb + i) * N + n4;
          atomicAdd(yp, acc[i].x); atomicAdd(yp + 1, acc[i].y); atomicAdd(yp + 2, acc[i].z); atomicAdd(yp + 3, acc[i].w);
        }
      }
  }
}

__global__ void __launch_bounds__(256) skinny_reduce_bias_act_kernel(const float* __restrict__ part, int parts,
                                                                     float* __restrict__ y,
                                                                     const float* __restrict__ bias, int M, int N,
                                                                     int relu) {
  // block = 32 consecutive outputs x 8 groups of partial sums: group g adds the slices g, g + 8, ... in order, the eight
  // group sums are added in order -- a fixed summation tree (bit-reproducible) with 256 blocks of loads in flight
  // (one thread per output walking all ~300 slices serially made this 8 KB reduction a 35 us kernel)
  pdl_prologue();
  __shared__ float sm[8][32];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + tx;
  float v = 0.f;
  if (i < M * N) {
#pragma unroll 4
    for (int p = ty; p < parts; p += 8) v += part[(int64_t)p * M * N + i];
  }
  sm[ty][tx] = v;
  __syncthreads();
  if (ty == 0 && i < M * N) {
#pragma unroll
    for (int g = 1; g < 8; ++g) v += sm[g][tx];
    v += bias ? bias[i % N] : 0.f;
    y[i] = relu ? fmaxf(v, 0.f) : v;
  }
}

__global__ void bias_act_kernel(float* __restrict__ y, const float* __restrict__ bias, int M, int N, int relu) {
  pdl_prologue();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * N) return;
  float v = y[i] + (bias ? bias[i % N] : 0.f);
  y[i] = relu ? fmaxf(v, 0.f) : v;
}

// da[m][k] (+)= sum_n dy[m][n] * w[k][n]; one warp per k row
// transpose-reduce across the 32 lanes of a warp: on return v[0] of lane l is the sum over all lanes of their v[l]
__device__ __forceinline__ void warp_transpose_sums(float (&v)[32], int lane) {
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

template <typename TA>
__global__ void __launch_bounds__(256) skinny_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ w,
                                                           TA* __restrict__ da, int64_t lda, int M, int K, int N,
                                                           int acc_flag) {
  // a warp takes TWO weight rows k per iteration against 16 batch rows: every dy value read from shared memory feeds
  // two FMAs, and the 32 partial sums (2 k x 16 m) are reduced across the lanes with one 31-shuffle transpose-reduce
  pdl_prologue();
  extern __shared__ __align__(16) float sdy[];  // [M][N]
  if (((M * N) & 3) == 0 && (reinterpret_cast<uintptr_t>(dy) & 15) == 0) {
    for (int i = threadIdx.x * 4; i < M * N; i += blockDim.x * 4)
      *reinterpret_cast<float4*>(sdy + i) = __ldg(reinterpret_cast<const float4*>(dy + i));
  } else {
    for (int i = threadIdx.x; i < M * N; i += blockDim.x) sdy[i] = dy[i];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if ((N & 127) == 0 && N <= 512) {
    // Software-pipelined path (class_attention_conv, N = 512): the 2 x N weights of the NEXT pair of k rows are loaded
    // (up to eight 128-bit loads per lane) before the current pair is multiplied, reduced and stored, so the weight
    // stream never waits for the shuffle reduction.  The grid is two blocks per SM (one wave): every warp walks
    // several pairs and the shared-memory copy of dy is made 2 x SMs times, not K / 16 times.
    const int ns = N >> 7;
    const int kstride = gridDim.x * 16;
    float4 wa[4], wb[4];
    int k = (blockIdx.x * 8 + wid) * 2;
    auto load_pair = [&](int kk, float4 (&xa)[4], float4 (&xb)[4]) {
      const float* w0 = w + (int64_t)kk * N + lane * 4;
      const float* w1 = w + (int64_t)(kk + 1 < K ? kk + 1 : kk) * N + lane * 4;
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (q < ns) {
          xa[q] = __ldg(reinterpret_cast<const float4*>(w0 + q * 128));
          xb[q] = __ldg(reinterpret_cast<const float4*>(w1 + q * 128));
        }
    };
    if (k < K) load_pair(k, wa, wb);
    for (; k < K; k += kstride) {
      float4 na[4], nb[4];
      if (k + kstride < K) load_pair(k + kstride, na, nb);
      const bool two = k + 1 < K;
      for (int mb = 0; mb < M; mb += 16) {
        float acc[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (q < ns) {
            const int n = q * 128 + lane * 4;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
              if (mb + i < M) g = *reinterpret_cast<const float4*>(sdy + (mb + i) * N + n);
              acc[i] = fmaf(g.x, wa[q].x, fmaf(g.y, wa[q].y, fmaf(g.z, wa[q].z, fmaf(g.w, wa[q].w, acc[i]))));
              acc[16 + i] = fmaf(g.x, wb[q].x, fmaf(g.y, wb[q].y, fmaf(g.z, wb[q].z, fmaf(g.w, wb[q].w, acc[16 + i]))));
            }
          }
        warp_transpose_sums(acc, lane);
        const int kk = k + (lane >> 4), m = mb + (lane & 15);
        if (m < M && kk < K && (two || lane < 16)) {
          TA* dst = da + (int64_t)m * lda + kk;
          float v = acc[0];
          if (acc_flag) v += to_f32(*dst);
          *dst = from_f32<TA>(v);
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        wa[q] = na[q];
        wb[q] = nb[q];
      }
    }
    return;
  }
  for (int k = (blockIdx.x * 8 + wid) * 2; k < K; k += gridDim.x * 16) {
    const bool two = k + 1 < K;
    const float* w0 = w + (int64_t)k * N;
    const float* w1 = w + (int64_t)(two ? k + 1 : k) * N;
    for (int mb = 0; mb < M; mb += 16) {
      float acc[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[i] = 0.f;
      if ((N & 127) == 0) {
        // 128-bit loads: a lane takes 4 consecutive columns per step (the 52 MB class_attention_conv weight matrix is
        // read once; with scalar loads the kernel was bound by load / shared-memory instructions, not by HBM)
        for (int n = lane * 4; n < N; n += 128) {
          const float4 wa = __ldg(reinterpret_cast<const float4*>(w0 + n));
          const float4 wb = __ldg(reinterpret_cast<const float4*>(w1 + n));
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
            if (mb + i < M) g = *reinterpret_cast<const float4*>(sdy + (mb + i) * N + n);
            acc[i] = fmaf(g.x, wa.x, fmaf(g.y, wa.y, fmaf(g.z, wa.z, fmaf(g.w, wa.w, acc[i]))));
            acc[16 + i] = fmaf(g.x, wb.x, fmaf(g.y, wb.y, fmaf(g.z, wb.z, fmaf(g.w, wb.w, acc[16 + i]))));
          }
        }
      } else {
        for (int n = lane; n < N; n += 32) {
          const float wa = __ldg(w0 + n), wb = __ldg(w1 + n);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float g = mb + i < M ? sdy[(mb + i) * N + n] : 0.f;
            acc[i] = fmaf(g, wa, acc[i]);
            acc[16 + i] = fmaf(g, wb, acc[16 + i]);
          }
        }
      }
      warp_transpose_sums(acc, lane);          // lane l now holds the total of entry l: (k + l / 16, mb + l % 16)
      const int kk = k + (lane >> 4), m = mb + (lane & 15);
      if (m < M && kk < K && (two || lane < 16)) {
        TA* dst = da + (int64_t)m * lda + kk;
        float v = acc[0];
        if (acc_flag) v += to_f32(*dst);
        *dst = from_f32<TA>(v);
      }
    }
  }
}

// Register-resident variant for N in {128, 256, 512} and at most 16 batch rows (class_attention_conv).  In the kernel
// above every weight is multiplied with 16 dy values read from SHARED memory (64 bytes of shared-memory traffic per
// 4-byte weight: the kernel is bound by the shared-memory pipe at ~1.3 TB/s of weight stream).  Here a warp owns ONE
// group of 128 columns for the whole kernel, so its 16 x 4 dy values live in registers; the inner loop is a 128-bit
// weight load + 64 FMAs per row, the next iteration's eight rows are loaded before the current eight are multiplied,
// and the partial dot products of the NQ column groups are combined through a 4 KB shared-memory exchange once per
// block iteration (8 * 8 / NQ rows).  The combine step writes runs of consecutive k per batch row.
template <typename TA, int NQ>
__global__ void __launch_bounds__(256, 1) skinny_dgrad_reg_kernel(const float* __restrict__ dy,
                                                                  const float* __restrict__ w, TA* __restrict__ da,
                                                                  int64_t lda, int M, int K, int N, int acc_flag) {
  pdl_prologue();
  constexpr int WG = 8 / NQ;            // warps per column group = row sub-blocks per iteration
  constexpr int RB = WG * 8;            // k rows per block iteration
  __shared__ float part_s[2][8][4][32]; // [buffer][warp][row pair][lane]
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int q = wid % NQ, rsub = wid / NQ;
  const int col = q * 128 + lane * 4;
  float4 g[16];
#pragma unroll
  for (int i = 0; i < 16; ++i)
    g[i] = i < M ? __ldg(reinterpret_cast<const float4*>(dy + (int64_t)i * N + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
  const int nblk = (K + RB - 1) / RB;
  auto load_rows = [&](int rb, float4 (&x)[8]) {
    const int r0 = rb * RB + rsub * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int r = min(r0 + j, K - 1);                 // (rows past the end are computed and dropped at the store)
      x[j] = __ldg(reinterpret_cast<const float4*>(w + (int64_t)r * N + col));
    }
  };
  float4 wv[8], wn[8];
  int rb = blockIdx.x, it = 0;
  if (rb < nblk) load_rows(rb, wv);
  for (; rb < nblk; rb += gridDim.x, ++it) {
    if (rb + (int)gridDim.x < nblk) load_rows(rb + gridDim.x, wn);
    float (*ps)[4][32] = part_s[it & 1];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      float acc[32];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float4 a = wv[2 * p], b = wv[2 * p + 1];
        acc[i] = fmaf(g[i].x, a.x, fmaf(g[i].y, a.y, fmaf(g[i].z, a.z, g[i].w * a.w)));
        acc[16 + i] = fmaf(g[i].x, b.x, fmaf(g[i].y, b.y, fmaf(g[i].z, b.z, g[i].w * b.w)));
      }
      warp_transpose_sums(acc, lane);     // lane l: row 2p + l / 16, batch row l % 16, over this warp's 128 columns
      ps[wid][p][lane] = acc[0];
    }
    __syncthreads();                      // (double-buffered exchange: one barrier per iteration)
    // combine the NQ column groups; thread -> (batch row m, row r of this block iteration): consecutive threads write
    // consecutive k of one batch row
    for (int t = threadIdx.x; t < 16 * RB; t += 256) {
      const int m = t / RB, r = t - m * RB;
      const int rs = r >> 3, j = r & 7;              // row sub-block, row inside it
      const int kk = rb * RB + r;
      if (m < M && kk < K) {
        float v = 0.f;
#pragma unroll
        for (int qq = 0; qq < NQ; ++qq) v += ps[rs * NQ + qq][j >> 1][(j & 1) * 16 + m];
        TA* dst = da + (int64_t)m * lda + kk;
        if (acc_flag) v += to_f32(*dst);
        *dst = from_f32<TA>(v);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) wv[j] = wn[j];
  }
}

// dw[k][n] += sum_m a[m][k] * dy[m][n]
template <typename TA, int MB>
__global__ void __launch_bounds__(128) skinny_wgrad_kernel(const TA* __restrict__ a, int64_t lda,
                                                           const float* __restrict__ dy, float* __restrict__ dw,
                                                           float* __restrict__ dbias, int M, int K, int N) {
  pdl_prologue();
  __shared__ __align__(16) float sa[SK_KT][MB];           // [k][m]: the MB batch values of one k are contiguous
  const int n = blockIdx.x * 128 + threadIdx.x;
  const int k0 = blockIdx.y * SK_KT;
  const int kt = min(SK_KT, K - k0);
  for (int i = threadIdx.x; i < MB * SK_KT; i += 128) {   // rows >= M are zero-filled (they are multiplied)
    int m = i / SK_KT, k = i - m * SK_KT;
    sa[k][m] = (m < M && k < kt) ? to_f32(a[(int64_t)m * lda + k0 + k]) : 0.f;
  }
  __syncthreads();
  if (n >= N) return;
  float g[MB];
  float bs = 0.f;
#pragma unroll
  for (int m = 0; m < MB; ++m) {
    g[m] = m < M ? dy[(int64_t)m * N + n] : 0.f;
    bs += g[m];
  }
  if (dbias && blockIdx.y == 0) dbias[n] += bs;
  // eight k rows per step: the eight read-modify-writes of dw are independent and in flight together (one dependent
  // load -> add -> store per k made the 512 x 21 fc gradient a 40 us chain of 64 global round trips)
  for (int kb = 0; kb < kt; kb += 8) {
    float acc[8], old[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      acc[j] = 0.f;
      old[j] = kb + j < kt ? dw[(int64_t)(k0 + kb + j) * N + n] : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (kb + j >= kt) break;
#pragma unroll
      for (int m4 = 0; m4 < MB; m4 += 4) {
        const float4 av = *reinterpret_cast<const float4*>(&sa[kb + j][m4]);
        acc[j] = fmaf(av.x, g[m4], fmaf(av.y, g[m4 + 1], fmaf(av.z, g[m4 + 2], fmaf(av.w, g[m4 + 3], acc[j]))));
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (kb + j < kt) dw[(int64_t)(k0 + kb + j) * N + n] = old[j] + acc[j];
  }
}

// Vectorised variant for N % 4 == 0 and at most 16 batch rows: a thread owns four consecutive columns, eight 128-bit
// read-modify-writes of dw in flight (the 52 MB class_attention_conv gradient: 105 MB of traffic per step).
template <typename TA>
__global__ void __launch_bounds__(128) skinny_wgrad4_kernel(const TA* __restrict__ a, int64_t lda,
                                                            const float* __restrict__ dy, float* __restrict__ dw,
                                                            float* __restrict__ dbias, int M, int K, int N) {
  pdl_prologue();
  constexpr int MB = 16;
  __shared__ __align__(16) float sa[SK_KT][MB];
  const int n4 = (blockIdx.x * 128 + threadIdx.x) * 4;
  const int k0 = blockIdx.y * SK_KT;
  const int kt = min(SK_KT, K - k0);
  for (int i = threadIdx.x; i < MB * SK_KT; i += 128) {
    int m = i / SK_KT, k = i - m * SK_KT;
    sa[k][m] = (m < M && k < kt) ? to_f32(a[(int64_t)m * lda + k0 + k]) : 0.f;
  }
  __syncthreads();
  if (n4 >= N) return;
  float4 g[MB];
  float4 bs = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int m = 0; m < MB; ++m) {
    g[m] = m < M ? __ldg(reinterpret_cast<const float4*>(dy + (int64_t)m * N + n4)) : make_float4(0.f, 0.f, 0.f, 0.f);
    bs.x += g[m].x; bs.y += g[m].y; bs.z += g[m].z; bs.w += g[m].w;
  }
  if (dbias && blockIdx.y == 0) {
    dbias[n4] += bs.x; dbias[n4 + 1] += bs.y; dbias[n4 + 2] += bs.z; dbias[n4 + 3] += bs.w;
  }
  for (int kb = 0; kb < kt; kb += 8) {
    float4 old[8], acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      old[j] = kb + j < kt ? *reinterpret_cast<const float4*>(dw + (int64_t)(k0 + kb + j) * N + n4)
                           : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (kb + j >= kt) break;
#pragma unroll
      for (int m4 = 0; m4 < MB; m4 += 4) {
        const float4 av = *reinterpret_cast<const float4*>(&sa[kb + j][m4]);
        float4 t = acc[j];
        t.x = fmaf(av.x, g[m4].x, fmaf(av.y, g[m4 + 1].x, fmaf(av.z, g[m4 + 2].x, fmaf(av.w, g[m4 + 3].x, t.x))));
        t.y = fmaf(av.x, g[m4].y, fmaf(av.y, g[m4 + 1].y, fmaf(av.z, g[m4 + 2].y, fmaf(av.w, g[m4 + 3].y, t.y))));
        t.z = fmaf(av.x, g[m4].z, fmaf(av.y, g[m4 + 1].z, fmaf(av.z, g[m4 + 2].z, fmaf(av.w, g[m4 + 3].z, t.z))));
        t.w = fmaf(av.x, g[m4].w, fmaf(av.y, g[m4 + 1].w, fmaf(av.z, g[m4 + 2].w, fmaf(av.w, g[m4 + 3].w, t.w))));
        acc[j] = t;
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (kb + j < kt)
        *reinterpret_cast<float4*>(dw + (int64_t)(k0 + kb + j) * N + n4) =
            make_float4(old[j].x + acc[j].x, old[j].y + acc[j].y, old[j].z + acc[j].z, old[j].w + acc[j].w);
  }
}

// ------------------------------------------------------------------------------------------
// Stem convolution conv1_1_3x3_s2 (BAISPSPNet.py:274): Cin = 4 (RGB + click map) is too narrow for the GEMM tiles
// above (K = 36) and for a TMA box, but the whole filter bank fits in shared memory.  fprop: one thread per
// (output pixel, 8 output channels); the 9 float4 input loads are shared by the 4..8 lanes of a pixel and the
// weights are shared-memory broadcasts.  wgrad: lane = output channel, a warp walks pixels and keeps the 36 x NC
// partial sums in registers; block-level reduction in shared memory, then one atomic per weight per block.
// ------------------------------------------------------------------------------------------
struct StemConv {
  const float* X; const void* Y; const float* W; float* dW;
  int N, IH, IW, OH, OW, ldy;
  int stride, pad_t, pad_l, relu;
  int64_t M;     // N*OH*OW
  // optional fused batch-norm statistics of the produced tensor (same protocol as the tcgen05 fprop epilogue):
  // per-channel sum / sum of squares of the values as stored, double atomics into BASI_BN_REPLICAS replicas, the
  // last block finalizes bnp = [mean | istd | gamma*istd | beta]
  double* bn_sums; const float* bn_gamma; const float* bn_beta; float* bn_bnp; unsigned int* bn_counter;
  double bn_count; float bn_eps;
};

template <typename TD, int COUT>
__global__ void __launch_bounds__(256) stem_fprop_kernel(const StemConv g) {
  pdl_prologue();
  // thread = (4 consecutive output pixels of one row, 8 output channels): every weight fetched from shared memory
  // feeds 4 FMAs, and neighbouring pixels share their input columns
  __shared__ __align__(16) float ws[36 * COUT];
  for (int i = threadIdx.x; i < 36 * COUT; i += blockDim.x) ws[i] = g.W[i];
  __syncthreads();
  constexpr int CG = COUT / 8, PX = 4;
  float ssum[8], ssq[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) ssum[j] = ssq[j] = 0.f;
  const int owg = (g.OW + PX - 1) / PX;
  const int total = g.N * g.OH * owg * CG;
  TD* Y = (TD*)g.Y;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int cg = i % CG;
    int t = i / CG;
    const int ow0 = (t % owg) * PX;
    t /= owg;
    const int oh = t % g.OH;
    const int n = t / g.OH;
    float acc[PX][8];
#pragma unroll
    for (int px = 0; px < PX; ++px)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[px][j] = 0.f;
    const int ih0 = oh * g.stride - g.pad_t;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int ih = ih0 + r;
      const bool rok = (unsigned)ih < (unsigned)g.IH;
      const float* xrow = g.X + ((int64_t)n * g.IH + ih) * g.IW * 4;
      float4 xv[PX][3];
#pragma unroll
      for (int px = 0; px < PX; ++px)
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          const int iw = (ow0 + px) * g.stride - g.pad_l + q;
          xv[px][q] = (rok && (unsigned)iw < (unsigned)g.IW) ? __ldg(reinterpret_cast<const float4*>(xrow + iw * 4))
                                                             : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
      for (int q = 0; q < 3; ++q)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float* wp = ws + ((r * 3 + q) * 4 + c) * COUT + cg * 8;
          const float4 w0 = *reinterpret_cast<const float4*>(wp), w1 = *reinterpret_cast<const float4*>(wp + 4);
#pragma unroll
          for (int px = 0; px < PX; ++px) {
            const float xs = c == 0 ? xv[px][q].x : (c == 1 ? xv[px][q].y : (c == 2 ? xv[px][q].z : xv[px][q].w));
            acc[px][0] = fmaf(xs, w0.x, acc[px][0]); acc[px][1] = fmaf(xs, w0.y, acc[px][1]);
            acc[px][2] = fmaf(xs, w0.z, acc[px][2]); acc[px][3] = fmaf(xs, w0.w, acc[px][3]);
            acc[px][4] = fmaf(xs, w1.x, acc[px][4]); acc[px][5] = fmaf(xs, w1.y, acc[px][5]);
            acc[px][6] = fmaf(xs, w1.z, acc[px][6]); acc[px][7] = fmaf(xs, w1.w, acc[px][7]);
          }
        }
    }
#pragma unroll
    for (int px = 0; px < PX; ++px) {
      if (ow0 + px >= g.OW) break;
      if (g.relu) {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[px][j] = fmaxf(acc[px][j], 0.f);
      }
      TD* yp = Y + (((int64_t)n * g.OH + oh) * g.OW + ow0 + px) * g.ldy + cg * 8;
      const float lo[4] = {acc[px][0], acc[px][1], acc[px][2], acc[px][3]};
      const float hi[4] = {acc[px][4], acc[px][5], acc[px][6], acc[px][7]};
      store4<TD>(yp, lo);
      store4<TD>(yp + 4, hi);
      if (g.bn_sums) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float v = to_f32(from_f32<TD>(acc[px][j]));     // the value as stored
          ssum[j] += v;
          ssq[j] = fmaf(v, v, ssq[j]);
        }
      }
    }
  }
  if (g.bn_sums) {
    // every thread keeps the same channel group for the whole grid-stride loop (grid * block is a multiple of CG)
    __shared__ float sred[256][16];
    __shared__ unsigned int s_last;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sred[threadIdx.x][j] = ssum[j];
      sred[threadIdx.x][8 + j] = ssq[j];
    }
    __syncthreads();
    if (threadIdx.x < COUT) {
      const int c = threadIdx.x, cgc = c / 8, j = c % 8;
      double s1 = 0, s2 = 0;
      for (int t = cgc; t < 256; t += CG) {
        s1 += (double)sred[t][j];
        s2 += (double)sred[t][8 + j];
      }
      double* rep = g.bn_sums + (size_t)(blockIdx.x % BASI_BN_REPLICAS) * 2 * COUT;
      atomicAdd(rep + c, s1);
      atomicAdd(rep + COUT + c, s2);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(g.bn_counter, 1u) == gridDim.x - 1) ? 1u : 0u;
    __syncthreads();
    if (s_last && g.bn_bnp && threadIdx.x < COUT) {
      __threadfence();
      const int c = threadIdx.x;
      double s1 = 0, s2 = 0;
#pragma unroll
      for (int r = 0; r < BASI_BN_REPLICAS; ++r) {
        s1 += __ldcg(g.bn_sums + (size_t)r * 2 * COUT + c);
        s2 += __ldcg(g.bn_sums + (size_t)r * 2 * COUT + COUT + c);
      }
      const double mean = s1 / g.bn_count;
      double var = s2 / g.bn_count - mean * mean;
      if (var < 0) var = 0;
      const double istd = 1.0 / sqrt(var + (double)g.bn_eps);
      g.bn_bnp[c] = (float)mean;
      g.bn_bnp[COUT + c] = (float)istd;
      g.bn_bnp[2 * COUT + c] = (float)((double)g.bn_gamma[c] * istd);
      g.bn_bnp[3 * COUT + c] = g.bn_beta[c];
    }
  }
}

template <typename TG, int COUT>
__global__ void __launch_bounds__(128) stem_wgrad_kernel(const StemConv g) {
  pdl_prologue();
  // lane = output channel.  A warp takes 32 consecutive output pixels of one row per iteration: it stages their input
  // patch (3 rows x (31*stride + 3) columns of float4) and their 32 x COUT gradients in shared memory with coalesced
  // loads, then walks the pixels reading the taps as shared-memory broadcasts; the 36 x NC partial sums stay in
  // registers for the whole kernel.
  constexpr int NC = COUT / 32, XC = 66, WARPS = (sizeof(TG) * COUT > 128) ? 2 : 4;   // <= 48 KB static smem
  __shared__ __align__(16) float4 sx[WARPS][3][XC];
  __shared__ TG sdy[WARPS][32][COUT];
  __shared__ float red[36 * COUT];
  for (int i = threadIdx.x; i < 36 * COUT; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int warp = blockIdx.x * WARPS + wib, nwarps = gridDim.x * WARPS;
  const TG* G = (const TG*)g.Y;
  const int chunks_per_row = (g.OW + 31) / 32;
  const int nchunks = g.N * g.OH * chunks_per_row;
  const int ncols = 31 * g.stride + 3;
  float acc[36 * NC];
#pragma unroll
  for (int j = 0; j < 36 * NC; ++j) acc[j] = 0.f;
  for (int ch = warp; ch < nchunks; ch += nwarps) {
    const int ow0 = (ch % chunks_per_row) * 32;
    const int t = ch / chunks_per_row;
    const int oh = t % g.OH, n = t / g.OH;
    const int npx = min(32, g.OW - ow0);
    const int iw_base = ow0 * g.stride - g.pad_l;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int ih = oh * g.stride - g.pad_t + r;
      const bool rok = (unsigned)ih < (unsigned)g.IH;
      const float* xrow = g.X + ((int64_t)n * g.IH + ih) * g.IW * 4;
      for (int c = lane; c < ncols; c += 32) {
        const int iw = iw_base + c;
        sx[wib][r][c] = (rok && (unsigned)iw < (unsigned)g.IW) ? __ldg(reinterpret_cast<const float4*>(xrow + iw * 4))
                                                              : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    const int64_t p0 = ((int64_t)n * g.OH + oh) * g.OW + ow0;
    for (int px = 0; px < npx; ++px)
#pragma unroll
      for (int j = 0; j < NC; ++j) sdy[wib][px][lane + 32 * j] = G[(p0 + px) * g.ldy + lane + 32 * j];
    __syncwarp();
    for (int px = 0; px < npx; ++px) {
      float gv[NC];
#pragma unroll
      for (int j = 0; j < NC; ++j) gv[j] = to_f32(sdy[wib][px][lane + 32 * j]);
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          const float4 xv = sx[wib][r][px * g.stride + q];
          const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
          for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int j = 0; j < NC; ++j)
              acc[((r * 3 + q) * 4 + c) * NC + j] = fmaf(xs[c], gv[j], acc[((r * 3 + q) * 4 + c) * NC + j]);
        }
    }
    __syncwarp();
  }
#pragma unroll
  for (int k = 0; k < 36; ++k)
#pragma unroll
    for (int j = 0; j < NC; ++j) atomicAdd(&red[k * COUT + lane + 32 * j], acc[k * NC + j]);
  __syncthreads();
  for (int i = threadIdx.x; i < 36 * COUT; i += blockDim.x) atomicAdd(g.dW + i, red[i]);
}

// ------------------------------------------------------------------------------------------
// Segment heads conv6_n* (BAISPSPNet.py:722-724): a 1x1 convolution to 1..4 fp32 logit channels is a per-pixel dot
// product, not a GEMM tile.  fprop: warp per pixel; dgrad: thread per 128-bit fragment of dx; wgrad: thread owns a
// channel group and walks pixels, block reduction, one atomic per weight per block (+ the bias gradient).
// ------------------------------------------------------------------------------------------
template <typename TX, int CO>
__global__ void __launch_bounds__(256) head_fprop_kernel(const TX* __restrict__ x, int ldx, int C, const float* __restrict__ w,
                                                         const float* __restrict__ bias, float* __restrict__ y, int ldy,
                                                         int relu, int64_t M) {
  pdl_prologue();
  constexpr int VN = Vec<TX>::N;
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int cgs = C / VN;
  for (int64_t p = warp; p < M; p += nwarps) {
    float acc[CO];
#pragma unroll
    for (int o = 0; o < CO; ++o) acc[o] = 0.f;
    for (int cg = lane; cg < cgs; cg += 32) {
      const Vec<TX> v = Vec<TX>::load(x + p * ldx + cg * VN);
#pragma unroll
      for (int j = 0; j < VN; ++j)
#pragma unroll
        for (int o = 0; o < CO; ++o) acc[o] = fmaf(v.v[j], __ldg(w + (cg * VN + j) * CO + o), acc[o]);
    }
#pragma unroll
    for (int o = 0; o < CO; ++o) acc[o] = warp_sum(acc[o]);
    if (lane == 0) {
#pragma unroll
      for (int o = 0; o < CO; ++o) {
        float r = acc[o] + (bias ? bias[o] : 0.f);
        y[p * ldy + o] = relu ? fmaxf(r, 0.f) : r;
      }
    }
  }
}

template <typename TD, int CO>
__global__ void __launch_bounds__(256) head_dgrad_kernel(const float* __restrict__ dy, int ldy, const float* __restrict__ w,
                                                         TD* __restrict__ dx, int ldx, int C, int acc, int64_t total) {
  pdl_prologue();
  constexpr int VN = Vec<TD>::N;
  const int cgs = C / VN;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cg = (int)(i % cgs);
    const int64_t p = i / cgs;
    float g[CO];
#pragma unroll
    for (int o = 0; o < CO; ++o) g[o] = dy[p * ldy + o];
    Vec<TD> r = acc ? Vec<TD>::load(dx + p * ldx + cg * VN) : Vec<TD>::zero();
#pragma unroll
    for (int j = 0; j < VN; ++j)
#pragma unroll
      for (int o = 0; o < CO; ++o) r.v[j] = fmaf(g[o], __ldg(w + (cg * VN + j) * CO + o), r.v[j]);
    r.store(dx + p * ldx + cg * VN);
  }
}

template <typename TX, int CO>
__global__ void __launch_bounds__(256) head_wgrad_kernel(const TX* __restrict__ x, int ldx, int C, const float* __restrict__ dy,
                                                         int ldy, float* __restrict__ dw, float* __restrict__ dbias,
                                                         int64_t M) {
  pdl_prologue();
  constexpr int VN = Vec<TX>::N;
  extern __shared__ float hred[];   // [blockDim.y][blockDim.x][VN*CO + CO]
  const int cg = threadIdx.x;       // blockDim.x == C / VN
  float acc[VN][CO], bsum[CO];
#pragma unroll
  for (int o = 0; o < CO; ++o) {
    bsum[o] = 0.f;
#pragma unroll
    for (int j = 0; j < VN; ++j) acc[j][o] = 0.f;
  }
  for (int64_t p = (int64_t)blockIdx.x * blockDim.y + threadIdx.y; p < M; p += (int64_t)gridDim.x * blockDim.y) {
    const Vec<TX> v = Vec<TX>::load(x + p * ldx + cg * VN);
    float g[CO];
#pragma unroll
    for (int o = 0; o < CO; ++o) {
      g[o] = dy[p * ldy + o];
      bsum[o] += g[o];
    }
#pragma unroll
    for (int j = 0; j < VN; ++j)
#pragma unroll
      for (int o = 0; o < CO; ++o) acc[j][o] = fmaf(v.v[j], g[o], acc[j][o]);
  }
  constexpr int PER = VN * CO + CO;
  float* mine = hred + ((size_t)threadIdx.y * blockDim.x + threadIdx.x) * PER;
#pragma unroll
  for (int j = 0; j < VN; ++j)
#pragma unroll
    for (int o = 0; o < CO; ++o) mine[j * CO + o] = acc[j][o];
#pragma unroll
  for (int o = 0; o < CO; ++o) mine[VN * CO + o] = bsum[o];
  __syncthreads();
  if (threadIdx.y == 0) {
    for (int k = 0; k < PER; ++k) {
      float t = 0.f;
      for (int yy = 0; yy < (int)blockDim.y; ++yy) t += hred[((size_t)yy * blockDim.x + threadIdx.x) * PER + k];
      if (k < VN * CO) atomicAdd(dw + (cg * VN + k / CO) * CO + (k % CO), t);
      else if (dbias && cg == 0) atomicAdd(dbias + (k - VN * CO), t);
    }
  }
}

static bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

}  // namespace basi

using namespace basi;

static int check_conv(const basi_conv_desc* d, const basi_tensor* x, const basi_tensor* y, const char* who) {
  BASI_CHECK_ARG(d && x && y && x->ptr && y->ptr, "%s: null argument", who);
  BASI_CHECK_ARG(d->kh > 0 && d->kw > 0 && d->stride > 0 && d->dil > 0 && d->pad_t >= 0 && d->pad_l >= 0,
                 "%s: bad geometry", who);
  BASI_CHECK_ARG(x->n == y->n, "%s: batch mismatch", who);
  // output extent must be reachable: (OH-1)*stride + (kh-1)*dil - pad_t <= IH-1 + pad_bottom (pad_bottom >= 0 implied)
  BASI_CHECK_ARG((y->h - 1) * d->stride - d->pad_t < x->h && (y->w - 1) * d->stride - d->pad_l < x->w,
                 "%s: output larger than the padded input allows", who);
  return BASI_OK;
}

// conv1_1-shaped problem: fp32 NHWC4 input, 3x3, dilation 1, 32 or 64 output channels
static bool stem_shape(const basi_conv_desc* d, const basi_tensor* x, const basi_tensor* y) {
  return x->dtype == BASI_F32 && x->c == 4 && x->ld == 4 && aligned16(x->ptr) && d->kh == 3 && d->kw == 3 &&
         d->dil == 1 && d->stride <= 2 && (y->c == 32 || y->c == 64) && y->ld % 4 == 0 && aligned16(y->ptr) &&
         exp_env("BASI_NO_STEM") == nullptr;
}
static StemConv stem_args(const basi_conv_desc* d, const basi_tensor* x, const basi_tensor* y) {
  StemConv s{};
  s.X = (const float*)x->ptr; s.Y = y->ptr;
  s.N = x->n; s.IH = x->h; s.IW = x->w; s.OH = y->h; s.OW = y->w; s.ldy = y->ld;
  s.stride = d->stride; s.pad_t = d->pad_t; s.pad_l = d->pad_l;
  s.M = pixels(y);
  return s;
}

static void launch_stem_fprop(const StemConv& s, const basi_tensor* y, cudaStream_t st) {
  int64_t total = (int64_t)s.N * s.OH * ((s.OW + 3) / 4) * (y->c / 8);
  // with fused statistics every block ends with 2 * Cout double atomics: fewer, longer blocks
  int grid = grid_for(total, 256, s.bn_sums ? 3 : 16);
  if (y->dtype == BASI_BF16) {
    if (y->c == 32) basi::launch(stem_fprop_kernel<bf16, 32>, grid, 256, 0, st, s);
    else basi::launch(stem_fprop_kernel<bf16, 64>, grid, 256, 0, st, s);
  } else {
    if (y->c == 32) basi::launch(stem_fprop_kernel<float, 32>, grid, 256, 0, st, s);
    else basi::launch(stem_fprop_kernel<float, 64>, grid, 256, 0, st, s);
  }
}

// conv6_n-shaped problem: 1x1, stride 1, no padding, <= 4 fp32 output channels
static bool head_shape(const basi_conv_desc* d, const basi_tensor* x, const basi_tensor* y) {
  const int vn = x->dtype == BASI_F32 ? 4 : 8;
  return d->kh == 1 && d->kw == 1 && d->stride == 1 && d->pad_t == 0 && d->pad_l == 0 && y->c >= 1 && y->c <= 4 &&
         y->dtype == BASI_F32 && x->h == y->h && x->w == y->w && basi::vec_ok(x) && x->c / vn <= 256 &&
         exp_env("BASI_NO_HEAD") == nullptr;
}
#define BASI_HEAD_CO(co, ...)                    \
  switch (co) {                                  \
    case 1: { constexpr int CO = 1; __VA_ARGS__ } break; \
    case 2: { constexpr int CO = 2; __VA_ARGS__ } break; \
    case 3: { constexpr int CO = 3; __VA_ARGS__ } break; \
    default: { constexpr int CO = 4; __VA_ARGS__ } break; \
  }

template <int MODE>
static int launch_gemm_conv(const GemmConv& g, int ts, int td, cudaStream_t st) {
  dim3 grid((g.M + BM - 1) / BM, (g.DC + BN - 1) / BN);
  if (ts == BASI_F32 && td == BASI_F32) basi::launch(conv_gemm_kernel<MODE, float, float>, grid, 256, 0, st, g);
  else if (ts == BASI_BF16 && td == BASI_BF16) basi::launch(conv_gemm_kernel<MODE, bf16, bf16>, grid, 256, 0, st, g);
  else if (ts == BASI_BF16 && td == BASI_F32) basi::launch(conv_gemm_kernel<MODE, bf16, float>, grid, 256, 0, st, g);
  else basi::launch(conv_gemm_kernel<MODE, float, bf16>, grid, 256, 0, st, g);
  return BASI_OK;
}

extern "C" {

int basi_conv_fprop(const basi_conv_desc* d, const basi_tensor* x, const float* w, const float* bias,
                    const basi_tensor* y, void* stream) {
  int rc = check_conv(d, x, y, "conv_fprop");
  if (rc) return rc;
  BASI_CHECK_ARG(w, "conv_fprop: null weights");
  if (head_shape(d, x, y)) {
    const int64_t M = pixels(y);
    const int grid = basi::grid_for(M * 32, 256, 16);
    cudaStream_t st = (cudaStream_t)stream;
    BASI_HEAD_CO(y->c, {
      if (x->dtype == BASI_BF16)
        basi::launch(head_fprop_kernel<bf16, CO>, grid, 256, 0, st, (const bf16*)x->ptr, x->ld, x->c, w, bias, (float*)y->ptr,
                     y->ld, d->relu, M);
      else
        basi::launch(head_fprop_kernel<float, CO>, grid, 256, 0, st, (const float*)x->ptr, x->ld, x->c, w, bias,
                     (float*)y->ptr, y->ld, d->relu, M);
    })
    BASI_CHECK_LAUNCH("conv_fprop(head)");
    return BASI_OK;
  }
  if (stem_shape(d, x, y) && !bias) {
    StemConv s = stem_args(d, x, y);
    s.W = w; s.relu = d->relu;
    launch_stem_fprop(s, y, (cudaStream_t)stream);
    BASI_CHECK_LAUNCH("conv_fprop(stem)");
    return BASI_OK;
  }
  GemmConv g{};
  g.S = x->ptr; g.D = y->ptr; g.W = w; g.bias = bias;
  g.N = x->n; g.SH = x->h; g.SW = x->w; g.SC = x->c; g.lds = x->ld;
  g.DH = y->h; g.DW = y->w; g.DC = y->c; g.ldd = y->ld;
  g.Cin = x->c; g.Cout = y->c;
  g.kh = d->kh; g.kw = d->kw; g.stride = d->stride; g.dil = d->dil; g.pad_t = d->pad_t; g.pad_l = d->pad_l;
  g.relu = d->relu; g.accumulate = 0;
  g.M = (int)pixels(y); g.K = d->kh * d->kw * x->c;
  g.fastA = (x->c % 16 == 0) && (x->ld % 8 == 0) && aligned16(x->ptr);
  g.fastB = (y->c % 4 == 0) && aligned16(w);
  g.fastD = (y->c % 4 == 0) && (y->ld % 4 == 0) && aligned16(y->ptr);
  launch_gemm_conv<MODE_FPROP>(g, x->dtype, y->dtype, (cudaStream_t)stream);
  BASI_CHECK_LAUNCH("conv_fprop");
  return BASI_OK;
}

int basi_stem_fprop_stats_supported(const basi_conv_desc* d, const basi_tensor* x, const basi_tensor* y) {
  return (d && x && y && check_conv(d, x, y, "stem_fprop_stats") == BASI_OK && stem_shape(d, x, y)) ? 1 : 0;
}

int basi_stem_fprop_stats(const basi_conv_desc* d, const basi_tensor* x, const float* w, const basi_tensor* y,
                          double* sums, const float* gamma, const float* beta, double count, float eps, float* bnp,
                          uint32_t* counter, void* stream) {
  int rc = check_conv(d, x, y, "stem_fprop_stats");
  if (rc) return rc;
  BASI_CHECK_ARG(w && sums && counter && (!bnp || (gamma && beta && count > 0)), "stem_fprop_stats: bad argument");
  BASI_CHECK_ARG(stem_shape(d, x, y), "stem_fprop_stats: not a stem-shaped convolution (fp32 NHWC4 input, 3x3, 32/64 out)");
  StemConv s = stem_args(d, x, y);
  s.W = w; s.relu = d->relu;
  s.bn_sums = sums; s.bn_gamma = gamma; s.bn_beta = beta; s.bn_bnp = bnp; s.bn_counter = counter;
  s.bn_count = count; s.bn_eps = eps;
  launch_stem_fprop(s, y, (cudaStream_t)stream);
  BASI_CHECK_LAUNCH("stem_fprop_stats");
  return BASI_OK;
}

int basi_conv_dgrad(const basi_conv_desc* d, const basi_tensor* dy, const float* w, const basi_tensor* dx,
                    int accumulate, void* stream) {
  int rc = check_conv(d, dx, dy, "conv_dgrad");
  if (rc) return rc;
  BASI_CHECK_ARG(w, "conv_dgrad: null weights");
  if (head_shape(d, dx, dy)) {
    cudaStream_t st = (cudaStream_t)stream;
    BASI_HEAD_CO(dy->c, {
      if (dx->dtype == BASI_BF16) {
        const int64_t total = pixels(dx) * (dx->c / 8);
        basi::launch(head_dgrad_kernel<bf16, CO>, basi::grid_for(total, 256), 256, 0, st, (const float*)dy->ptr, dy->ld, w,
                     (bf16*)dx->ptr, dx->ld, dx->c, accumulate, total);
      } else {
        const int64_t total = pixels(dx) * (dx->c / 4);
        basi::launch(head_dgrad_kernel<float, CO>, basi::grid_for(total, 256), 256, 0, st, (const float*)dy->ptr, dy->ld, w,
                     (float*)dx->ptr, dx->ld, dx->c, accumulate, total);
      }
    })
    BASI_CHECK_LAUNCH("conv_dgrad(head)");
    return BASI_OK;
  }
  GemmConv g{};
  g.S = dy->ptr; g.D = dx->ptr; g.W = w; g.bias = nullptr;
  g.N = dx->n; g.SH = dy->h; g.SW = dy->w; g.SC = dy->c; g.lds = dy->ld;
  g.DH = dx->h; g.DW = dx->w; g.DC = dx->c; g.ldd = dx->ld;
  g.Cin = dx->c; g.Cout = dy->c;
  g.kh = d->kh; g.kw = d->kw; g.stride = d->stride; g.dil = d->dil; g.pad_t = d->pad_t; g.pad_l = d->pad_l;
  g.relu = 0; g.accumulate = accumulate;
  g.M = (int)pixels(dx); g.K = d->kh * d->kw * dy->c;
  g.fastA = (dy->c % 16 == 0) && (dy->ld % 8 == 0) && aligned16(dy->ptr);
  g.fastB = (dy->c % 16 == 0) && aligned16(w);
  g.fastD = (dx->c % 4 == 0) && (dx->ld % 4 == 0) && aligned16(dx->ptr);
  launch_gemm_conv<MODE_DGRAD>(g, dy->dtype, dx->dtype, (cudaStream_t)stream);
  BASI_CHECK_LAUNCH("conv_dgrad");
  return BASI_OK;
}

int basi_conv_wgrad(const basi_conv_desc* d, const basi_tensor* x, const basi_tensor* dy, float* dw, float* dbias,
                    void* stream) {
  int rc = check_conv(d, x, dy, "conv_wgrad");
  if (rc) return rc;
  BASI_CHECK_ARG(dw, "conv_wgrad: null dw");
  if (head_shape(d, x, dy)) {
    const int vn = x->dtype == BASI_F32 ? 4 : 8;
    const int bx = x->c / vn;
    int by = 256 / bx;
    if (by < 1) by = 1;
    const int64_t M = pixels(dy);
    int64_t want = (M + by * 8 - 1) / (by * 8);
    const int64_t cap = (int64_t)basi::sm_count() * 4;
    const int grid = (int)(want < cap ? (want < 1 ? 1 : want) : cap);
    cudaStream_t st = (cudaStream_t)stream;
    BASI_HEAD_CO(dy->c, {
      const size_t smem = (size_t)bx * by * (vn * CO + CO) * sizeof(float);
      if (x->dtype == BASI_BF16)
        basi::launch(head_wgrad_kernel<bf16, CO>, dim3(grid), dim3(bx, by), smem, st, (const bf16*)x->ptr, x->ld, x->c,
                     (const float*)dy->ptr, dy->ld, dw, dbias, M);
      else
        basi::launch(head_wgrad_kernel<float, CO>, dim3(grid), dim3(bx, by), smem, st, (const float*)x->ptr, x->ld, x->c,
                     (const float*)dy->ptr, dy->ld, dw, dbias, M);
    })
    BASI_CHECK_LAUNCH("conv_wgrad(head)");
    return BASI_OK;
  }
  if (stem_shape(d, x, dy) && !dbias) {
    StemConv s = stem_args(d, x, dy);
    s.dW = dw;
    int grid = basi::sm_count() * 6;
    cudaStream_t st = (cudaStream_t)stream;
    if (dy->dtype == BASI_BF16) {
      if (dy->c == 32) basi::launch(stem_wgrad_kernel<bf16, 32>, grid, 128, 0, st, s);
      else basi::launch(stem_wgrad_kernel<bf16, 64>, grid, 128, 0, st, s);
    } else {
      if (dy->c == 32) basi::launch(stem_wgrad_kernel<float, 32>, grid, 128, 0, st, s);
      else basi::launch(stem_wgrad_kernel<float, 64>, grid * 2, 64, 0, st, s);     // 2 warps per block (shared memory)
    }
    BASI_CHECK_LAUNCH("conv_wgrad(stem)");
    return BASI_OK;
  }
  WgradConv g{};
  g.X = x->ptr; g.G = dy->ptr; g.dW = dw;
  g.N = x->n; g.IH = x->h; g.IW = x->w; g.Cin = x->c; g.ldx = x->ld;
  g.OH = dy->h; g.OW = dy->w; g.Cout = dy->c; g.ldg = dy->ld;
  g.kh = d->kh; g.kw = d->kw; g.stride = d->stride; g.dil = d->dil; g.pad_t = d->pad_t; g.pad_l = d->pad_l;
  g.M = (int)pixels(dy); g.K = d->kh * d->kw * x->c;
  g.fastA = (x->c % 4 == 0) && (x->ld % 4 == 0) && aligned16(x->ptr);
  g.fastG = (dy->c % 4 == 0) && (dy->ld % 4 == 0) && aligned16(dy->ptr);
  int tiles = ((g.K + WK - 1) / WK) * ((g.Cout + WN - 1) / WN);
  int want = (basi::sm_count() * 4 + tiles - 1) / tiles;
  int maxsplit = (g.M + 255) / 256;  // >= 256 pixels per split
  int splits = want < maxsplit ? want : maxsplit;
  if (splits < 1) splits = 1;
  g.m_per_split = ((g.M + splits - 1) / splits + WP - 1) / WP * WP;
  splits = (g.M + g.m_per_split - 1) / g.m_per_split;
  dim3 grid((g.K + WK - 1) / WK, (g.Cout + WN - 1) / WN, splits);
  cudaStream_t st = (cudaStream_t)stream;
  if (x->dtype == BASI_F32 && dy->dtype == BASI_F32) basi::launch(conv_wgrad_kernel<float, float>, grid, 256, 0, st, g);
  else if (x->dtype == BASI_BF16 && dy->dtype == BASI_BF16) basi::launch(conv_wgrad_kernel<bf16, bf16>, grid, 256, 0, st, g);
  else if (x->dtype == BASI_BF16 && dy->dtype == BASI_F32) basi::launch(conv_wgrad_kernel<bf16, float>, grid, 256, 0, st, g);
  else basi::launch(conv_wgrad_kernel<float, bf16>, grid, 256, 0, st, g);
  BASI_CHECK_LAUNCH("conv_wgrad");
  if (dbias) {
    int64_t M = g.M;
    int gy = (int)((M + 255) / 256);
    if (gy > 256) gy = 256;
    dim3 cg((g.Cout + 31) / 32, gy), cb(32, 8);
    if (dy->dtype == BASI_F32) basi::launch(colsum_kernel<float>, cg, cb, 0, st, (const float*)dy->ptr, M, g.Cout, g.ldg, dbias);
    else basi::launch(colsum_kernel<bf16>, cg, cb, 0, st, (const bf16*)dy->ptr, M, g.Cout, g.ldg, dbias);
    BASI_CHECK_LAUNCH("conv_wgrad(dbias)");
  }
  return BASI_OK;
}

// Rows (M = batch) are independent in fwd / dgrad and additive in wgrad, so any M is served in chunks of rows that fit
// the kernels' register / shared-memory budgets (<= 64 rows; dgrad: M_chunk * N * 4 bytes <= 96 KB).
static int skinny_rows_per_chunk(int N) {
  int mc = (int)((96 * 1024) / ((size_t)N * 4));
  if (mc > 64) mc = 64;
  return mc;
}

int basi_skinny_supported(int M, int K, int N) {
  return (M > 0 && K > 0 && N > 0 && skinny_rows_per_chunk(N) >= 1) ? 1 : 0;
}

// k-slab partition of the forward GEMV: about 4 blocks per SM in total; every block walks `spb` consecutive k-slabs
static bool skinny_vec4(int N, const void* w) { return (N & 3) == 0 && (((uintptr_t)w) & 15) == 0; }

// the vectorised kernel (N % 4 == 0): a block = 512 columns x one run of `kpb` (<= SK4_KB, multiple of 8) k rows,
// about two blocks per SM
static void skinny_fwd4_geometry(int K, int N, int* gx, int* gy, int* kpb) {
  *gx = (N + 511) / 512;
  int y = (2 * basi::sm_count() + *gx - 1) / *gx;
  int k = ((K + y - 1) / y + 7) & ~7;
  if (k > SK4_KB) k = SK4_KB;
  if (k < 8) k = 8;
  *kpb = k;
  *gy = (K + k - 1) / k;
}

static void skinny_fwd_geometry(int K, int N, int* gx, int* gy, int* spb) {
  if ((N & 3) == 0) {        // (the workspace size must not depend on pointer alignment: see basi_skinny_fwd_ws)
    skinny_fwd4_geometry(K, N, gx, gy, spb);
    return;
  }
  const int nslabs = (K + SK_KT - 1) / SK_KT;
  *gx = (N + 127) / 128;
  int y = (4 * basi::sm_count() + *gx - 1) / *gx;
  if (y > nslabs) y = nslabs;
  *spb = (nslabs + y - 1) / y;
  *gy = (nslabs + *spb - 1) / *spb;
}

int64_t basi_skinny_fwd_workspace_floats(int M, int K, int N) {
  if (!basi_skinny_supported(M, K, N)) return -1;
  int gx, gy, spb;
  skinny_fwd_geometry(K, N, &gx, &gy, &spb);
  return (int64_t)gy * M * N;
}

int basi_skinny_fwd_ws(const void* a, int dtype_a, int64_t lda, const float* w, const float* bias, float* y, int M,
                       int K, int N, int relu, float* workspace, void* stream) {
  BASI_CHECK_ARG(a && w && y && basi_skinny_supported(M, K, N), "skinny_fwd: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  int gx, gy, spb;
  skinny_fwd_geometry(K, N, &gx, &gy, &spb);
  const bool vec4 = (N & 3) == 0;
  BASI_CHECK_ARG(!vec4 || (skinny_vec4(N, w) && (!workspace || (((uintptr_t)workspace) & 15) == 0)),
                 "skinny_fwd: weights / workspace must be 16-byte aligned when N is a multiple of 4");
  if (!workspace) cudaMemsetAsync(y, 0, sizeof(float) * (size_t)M * N, st);
  dim3 grid(gx, gy);
  const size_t es = dtype_a == BASI_F32 ? 4 : 2;
  for (int m0 = 0; m0 < M; m0 += 64) {
    const int mc = M - m0 < 64 ? M - m0 : 64;
    const char* ap = (const char*)a + (size_t)m0 * lda * es;
    float* yp = y + (size_t)m0 * N;
    // every row chunk owns a [gy][mc][N] block of the workspace (the blocks add up to gy * M * N floats)
    float* pp = workspace ? workspace + (size_t)gy * m0 * N : nullptr;
    if (vec4) {
      if (dtype_a == BASI_F32) basi::launch(skinny_fwd4_kernel<float>, grid, 128, 0, st, (const float*)ap, lda, w, yp, mc, K, N, spb, pp);
      else basi::launch(skinny_fwd4_kernel<bf16>, grid, 128, 0, st, (const bf16*)ap, lda, w, yp, mc, K, N, spb, pp);
    } else if (dtype_a == BASI_F32) basi::launch(skinny_fwd_kernel<float>, grid, 128, 0, st, (const float*)ap, lda, w, yp, mc, K, N, spb, pp);
    else basi::launch(skinny_fwd_kernel<bf16>, grid, 128, 0, st, (const bf16*)ap, lda, w, yp, mc, K, N, spb, pp);
    if (workspace) {
      basi::launch(skinny_reduce_bias_act_kernel, (mc * N + 31) / 32, 256, 0, st, (const float*)pp, gy, yp, bias, mc, N, relu);
    }
  }
  BASI_CHECK_LAUNCH("skinny_fwd");
  if (!workspace) {
    basi::launch(bias_act_kernel, (M * N + 255) / 256, 256, 0, st, y, bias, M, N, relu);
    BASI_CHECK_LAUNCH("skinny_fwd(bias)");
  }
  return BASI_OK;
}

int basi_skinny_fwd(const void* a, int dtype_a, int64_t lda, const float* w, const float* bias, float* y, int M, int K,
                    int N, int relu, void* stream) {
  return basi_skinny_fwd_ws(a, dtype_a, lda, w, bias, y, M, K, N, relu, nullptr, stream);
}

int basi_skinny_dgrad(const float* dy, const float* w, void* da, int dtype_a, int64_t lda, int M, int K, int N,
                      int accumulate, void* stream) {
  BASI_CHECK_ARG(dy && w && da && basi_skinny_supported(M, K, N), "skinny_dgrad: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int rows = skinny_rows_per_chunk(N);
  int blocks = (K + 15) / 16;
  // (the pipelined N <= 512 path keeps ~110 registers per thread: two blocks per SM are one full wave)
  int cap = basi::sm_count() * (((N & 127) == 0 && N <= 512) ? 2 : 8);
  if (blocks > cap) blocks = cap;
  const size_t es = dtype_a == BASI_F32 ? 4 : 2;
  if (M <= 16 && (N == 128 || N == 256 || N == 512) && (((uintptr_t)dy | (uintptr_t)w) & 15) == 0) {
    // register-resident kernel: one block per SM, every block walks K / (RB * SMs) row blocks
    const int nq = N / 128, rb = (8 / nq) * 8;
    int grid = (K + rb - 1) / rb;
    if (grid > basi::sm_count()) grid = basi::sm_count();
#define BASI_DGRAD_REG(TT, NQ) \
  basi::launch(skinny_dgrad_reg_kernel<TT, NQ>, grid, 256, 0, st, dy, w, (TT*)da, lda, M, K, N, accumulate)
    if (dtype_a == BASI_F32) {
      if (nq == 1) BASI_DGRAD_REG(float, 1); else if (nq == 2) BASI_DGRAD_REG(float, 2); else BASI_DGRAD_REG(float, 4);
    } else {
      if (nq == 1) BASI_DGRAD_REG(bf16, 1); else if (nq == 2) BASI_DGRAD_REG(bf16, 2); else BASI_DGRAD_REG(bf16, 4);
    }
#undef BASI_DGRAD_REG
    BASI_CHECK_LAUNCH("skinny_dgrad");
    return BASI_OK;
  }
  for (int m0 = 0; m0 < M; m0 += rows) {
    const int mc = M - m0 < rows ? M - m0 : rows;
    const size_t smem = sizeof(float) * (size_t)mc * N;
    const float* dyp = dy + (size_t)m0 * N;
    char* dap = (char*)da + (size_t)m0 * lda * es;
    if (dtype_a == BASI_F32) {
      if (smem > 48 * 1024)
        cudaFuncSetAttribute(skinny_dgrad_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
      basi::launch(skinny_dgrad_kernel<float>, blocks, 256, smem, st, dyp, w, (float*)dap, lda, mc, K, N, accumulate);
    } else {
      if (smem > 48 * 1024)
        cudaFuncSetAttribute(skinny_dgrad_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
      basi::launch(skinny_dgrad_kernel<bf16>, blocks, 256, smem, st, dyp, w, (bf16*)dap, lda, mc, K, N, accumulate);
    }
  }
  BASI_CHECK_LAUNCH("skinny_dgrad");
  return BASI_OK;
}

int basi_skinny_wgrad(const void* a, int dtype_a, int64_t lda, const float* dy, float* dw, float* dbias, int M, int K,
                      int N, void* stream) {
  BASI_CHECK_ARG(a && dy && dw && basi_skinny_supported(M, K, N), "skinny_wgrad: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((N + 127) / 128, (K + SK_KT - 1) / SK_KT);
  const size_t es = dtype_a == BASI_F32 ? 4 : 2;
  for (int m0 = 0; m0 < M; m0 += 64) {       // dw += ... : the row chunks simply add up (stream order)
    const int mc = M - m0 < 64 ? M - m0 : 64;
    const char* ap = (const char*)a + (size_t)m0 * lda * es;
    const float* dyp = dy + (size_t)m0 * N;
    if (mc <= 16 && skinny_vec4(N, dw) && (((uintptr_t)dyp) & 15) == 0) {
      dim3 grid4((N + 511) / 512, (K + SK_KT - 1) / SK_KT);
      if (dtype_a == BASI_F32) basi::launch(skinny_wgrad4_kernel<float>, grid4, 128, 0, st, (const float*)ap, lda, dyp, dw, dbias, mc, K, N);
      else basi::launch(skinny_wgrad4_kernel<bf16>, grid4, 128, 0, st, (const bf16*)ap, lda, dyp, dw, dbias, mc, K, N);
    } else if (mc <= 16) {
      if (dtype_a == BASI_F32) basi::launch(skinny_wgrad_kernel<float, 16>, grid, 128, 0, st, (const float*)ap, lda, dyp, dw, dbias, mc, K, N);
      else basi::launch(skinny_wgrad_kernel<bf16, 16>, grid, 128, 0, st, (const bf16*)ap, lda, dyp, dw, dbias, mc, K, N);
    } else {
      if (dtype_a == BASI_F32) basi::launch(skinny_wgrad_kernel<float, 64>, grid, 128, 0, st, (const float*)ap, lda, dyp, dw, dbias, mc, K, N);
      else basi::launch(skinny_wgrad_kernel<bf16, 64>, grid, 128, 0, st, (const bf16*)ap, lda, dyp, dw, dbias, mc, K, N);
    }
  }
  BASI_CHECK_LAUNCH("skinny_wgrad");
  return BASI_OK;
}

}  // extern "C"
