// Batch normalisation with batch statistics (tf.layers.batch_normalization(training=True)) for the BAIS
// PSPNet path: statistics (+ fused finalize), apply with fused ReLU / residual add / second BN, and the
// backward pair (per-channel reductions + fused finalize, then the dx pass with the fused residual gradient).
//
// HBM-bound streaming kernels.  Thread (tx, ty): tx owns one 128-bit channel group for the whole kernel (its
// per-channel parameters live in registers), ty strides over pixel rows with UNR independent 128-bit loads in
// flight per input.  Reductions: per-thread partials -> shared memory -> one double atomicAdd per channel per
// block; the last block to finish (ticket counter) turns the sums into the per-channel parameters, so no
// separate finalize launch is needed.
#include "common.cuh"
#include "ptx.cuh"
#include <stdlib.h>

namespace basi {

constexpr int UNR = 4;
constexpr int NREP = BASI_BN_REPLICAS;   // replicated accumulators: same-address atomic chains are NREP x shorter

struct RowGeom {
  dim3 grid, block;
  size_t smem;
};
// bx threads over channel groups, by over rows; grid.y covers channel groups beyond 256
static RowGeom row_geom(int64_t R, int C, int VN, int rows_per_thread, int waves, size_t smem_per_thread) {
  int cgs = C / VN;
  int bx = cgs < 256 ? cgs : 256;
  int by = 256 / bx;
  if (by < 1) by = 1;
  int gy = (cgs + bx - 1) / bx;
  int64_t want = (R + (int64_t)by * rows_per_thread - 1) / ((int64_t)by * rows_per_thread);
  int64_t cap = (int64_t)sm_count() * waves / gy;
  if (cap < 1) cap = 1;
  int gx = (int)(want < cap ? want : cap);
  if (gx < 1) gx = 1;
  RowGeom g;
  g.grid = dim3(gx, gy);
  g.block = dim3(bx, by);
  g.smem = (size_t)bx * by * smem_per_thread;
  return g;
}

// block reduction of per-thread (a, b) over threadIdx.y, then one atomicAdd per channel; returns true in the
// threads of the LAST block of the whole grid to finish (sums are then final and visible).
template <int VN>
__device__ __forceinline__ bool reduce_and_ticket(double (&a)[VN], double (&b)[VN], double* ga, double* gb, int c0,
                                                  bool valid, unsigned int* counter) {
  extern __shared__ double sred[];  // [blockDim.y][blockDim.x][2*VN]
  __shared__ bool is_last;
  const int tx = threadIdx.x, ty = threadIdx.y;
  double* mine = sred + ((size_t)ty * blockDim.x + tx) * (2 * VN);
#pragma unroll
  for (int i = 0; i < VN; ++i) {
    mine[i] = a[i];
    mine[VN + i] = b[i];
  }
  __syncthreads();
  if (ty == 0 && valid) {
    for (int y = 1; y < blockDim.y; ++y) {
      const double* o = sred + ((size_t)y * blockDim.x + tx) * (2 * VN);
#pragma unroll
      for (int i = 0; i < VN; ++i) {
        a[i] += o[i];
        b[i] += o[VN + i];
      }
    }
#pragma unroll
    for (int i = 0; i < VN; ++i) {
      atomicAdd(ga + c0 + i, a[i]);
      atomicAdd(gb + c0 + i, b[i]);
    }
  }
  __threadfence();
  __syncthreads();
  if (tx == 0 && ty == 0) {
    unsigned int total = gridDim.x * gridDim.y;
    unsigned int t = atomicAdd(counter, 1u);
    is_last = (t == total - 1);
  }
  __syncthreads();
  if (is_last) __threadfence();
  return is_last;
}

__device__ __forceinline__ void finalize_channel(const double* sums, const float* gamma, const float* beta,
                                                 double count, float eps, float* bnp, int C, int c) {
  double s1 = 0, s2 = 0;
#pragma unroll
  for (int r = 0; r < NREP; ++r) {
    s1 += __ldcg(sums + (size_t)r * 2 * C + c);
    s2 += __ldcg(sums + (size_t)r * 2 * C + C + c);
  }
  double mean = s1 / count;
  double var = s2 / count - mean * mean;
  if (var < 0) var = 0;
  double istd = 1.0 / sqrt(var + (double)eps);
  bnp[c] = (float)mean;
  bnp[C + c] = (float)istd;
  bnp[2 * C + c] = (float)((double)gamma[c] * istd);
  bnp[3 * C + c] = beta[c];
}

// ---- statistics: sums += (sum x, sum x^2) in double; last block writes bnp when gamma != NULL
template <typename T>
__global__ void __launch_bounds__(256) bn_stats_kernel(const T* __restrict__ x, int64_t R, int C, int ld,
                                                       double* __restrict__ sums, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, double count, float eps,
                                                       float* __restrict__ bnp, unsigned int* counter) {
  pdl_prologue();
  constexpr int VN = Vec<T>::N;
  const int cg = blockIdx.y * blockDim.x + threadIdx.x;
  const bool valid = cg * VN < C;
  const int c0 = cg * VN;
  double s[VN], q[VN];
#pragma unroll
  for (int i = 0; i < VN; ++i) s[i] = q[i] = 0.0;
  if (valid) {
    const int64_t rstep = (int64_t)gridDim.x * blockDim.y;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.y + threadIdx.y; r < R; r += rstep * UNR) {
      Vec<T> v[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int64_t rr = r + u * rstep;
        v[u] = rr < R ? Vec<T>::load(x + rr * ld + c0) : Vec<T>::zero();
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u)
#pragma unroll
        for (int i = 0; i < VN; ++i) {
          const double d = (double)v[u].v[i];
          s[i] += d;
          q[i] = fma(d, d, q[i]);
        }
    }
  }
  double* rep = sums + (size_t)(blockIdx.x % NREP) * 2 * C;
  const bool last = reduce_and_ticket<VN>(s, q, rep, rep + C, c0, valid, counter);
  if (last && gamma) {
    for (int c = threadIdx.y * blockDim.x + threadIdx.x; c < C; c += blockDim.x * blockDim.y)
      finalize_channel(sums, gamma, beta, count, eps, bnp, C, c);
  }
}

__global__ void bn_finalize_kernel(const double* __restrict__ sums, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, double count, float eps, float* __restrict__ bnp,
                                   int C) {
  pdl_prologue();
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) finalize_channel(sums, gamma, beta, count, eps, bnp, C, c);
}

// 128-bit packed row fragment: elements are unpacked at the point of use so that the UNR rows in flight cost
// 4 registers each (not 8 floats) -- these kernels live or die by occupancy (ncu: 210 regs -> 1 block/SM before).
template <typename T>
struct Pack;
template <>
struct Pack<float> {
  static constexpr int N = 4;
  float4 r;
  __device__ __forceinline__ void load(const float* p) { r = *reinterpret_cast<const float4*>(p); }
  __device__ __forceinline__ void store(float* p) const { *reinterpret_cast<float4*>(p) = r; }
  __device__ __forceinline__ float get(int i) const { return i == 0 ? r.x : (i == 1 ? r.y : (i == 2 ? r.z : r.w)); }
  __device__ __forceinline__ void set2(int i2, float a, float b) {
    if (i2 == 0) { r.x = a; r.y = b; } else { r.z = a; r.w = b; }
  }
};
template <>
struct Pack<bf16> {
  static constexpr int N = 8;
  uint4 r;
  __device__ __forceinline__ void load(const bf16* p) { r = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void store(bf16* p) const { *reinterpret_cast<uint4*>(p) = r; }
  __device__ __forceinline__ uint32_t word(int w) const { return w == 0 ? r.x : (w == 1 ? r.y : (w == 2 ? r.z : r.w)); }
  __device__ __forceinline__ float get(int i) const {
    const uint32_t w = word(i >> 1);
    return (i & 1) ? h16_hi(w) : h16_lo(w);
  }
  __device__ __forceinline__ void set2(int i2, float a, float b) {
    const uint32_t w = h16_pack(a, b);
    if (i2 == 0) r.x = w; else if (i2 == 1) r.y = w; else if (i2 == 2) r.z = w; else r.w = w;
  }
};

// ---- apply: out = act((x-mean)*scale+beta [+ res | + (res-rmean)*rscale+rbeta])
template <typename T, bool HAS_RES, bool RES_BN>
__global__ void __launch_bounds__(256, 2) bn_apply_kernel(const T* __restrict__ x, int ldx,
                                                          const float* __restrict__ bnp, const T* __restrict__ res,
                                                          int ldr, const float* __restrict__ rbnp, int relu,
                                                          T* __restrict__ out, int ldo, int64_t R, int C) {
  pdl_prologue();
  constexpr int VN = Pack<T>::N;
  constexpr int AUNR = 4;
  const int cg = blockIdx.y * blockDim.x + threadIdx.x;
  if (cg * VN >= C) return;
  const int c0 = cg * VN;
  // y = (x - mean)*scale + beta: the centred form keeps the B=1 / tiny-variance case exact (istd up to 316 would
  // amplify the rounding of x*scale - mean*scale)
  float mean[VN], scale[VN], beta[VN], rmean[VN], rscale[VN];
#pragma unroll
  for (int i = 0; i < VN; ++i) {
    mean[i] = bnp[c0 + i];
    scale[i] = bnp[2 * C + c0 + i];
    beta[i] = bnp[3 * C + c0 + i];
    if (RES_BN) {
      rmean[i] = rbnp[c0 + i];
      rscale[i] = rbnp[2 * C + c0 + i];
      beta[i] += rbnp[3 * C + c0 + i];
    }
  }
  const int64_t rstep = (int64_t)gridDim.x * blockDim.y;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.y + threadIdx.y; r < R; r += rstep * AUNR) {
    Pack<T> v[AUNR], rv[AUNR];
#pragma unroll
    for (int u = 0; u < AUNR; ++u) {
      const int64_t rr = r + u * rstep;
      if (rr < R) {
        v[u].load(x + rr * ldx + c0);
        if (HAS_RES) rv[u].load(res + rr * ldr + c0);
      }
    }
#pragma unroll
    for (int u = 0; u < AUNR; ++u) {
      const int64_t rr = r + u * rstep;
      if (rr >= R) break;
#pragma unroll
      for (int i2 = 0; i2 < VN / 2; ++i2) {
        float o[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int i = 2 * i2 + h;
          float t = fmaf(v[u].get(i) - mean[i], scale[i], beta[i]);
          if (HAS_RES) t = RES_BN ? fmaf(rv[u].get(i) - rmean[i], rscale[i], t) : t + rv[u].get(i);
          o[h] = relu ? fmaxf(t, 0.f) : t;
        }
        v[u].set2(i2, o[0], o[1]);
      }
      v[u].store(out + rr * ldo + c0);
    }
  }
}

constexpr int BUNR = 2;   // rows in flight per thread in the register-staged backward kernels

// ---- backward reduce: dsums += (sum dy, sum dy*xhat); dy = dout * mask.
// mask: out > 0 when out != NULL; else, when relu_from_x, (x-mean)*scale+beta > 0 (plain BN+ReLU: no need to read out)
template <typename T>
__global__ void __launch_bounds__(256, 2)
bn_bwd_reduce_kernel(const T* __restrict__ dout, int ldd, const T* __restrict__ out, int ldo, const T* __restrict__ x,
                     int ldx, const float* __restrict__ bnp, int relu_from_x, int64_t R, int C,
                     double* __restrict__ dsums, double count, float* __restrict__ dgamma, float* __restrict__ dbeta,
                     float* __restrict__ coef, unsigned int* counter) {
  pdl_prologue();
  constexpr int VN = Pack<T>::N;
  const int cg = blockIdx.y * blockDim.x + threadIdx.x;
  const bool valid = cg * VN < C;
  const int c0 = cg * VN;
  float fs[VN], fq[VN];
#pragma unroll
  for (int i = 0; i < VN; ++i) fs[i] = fq[i] = 0.f;
  if (valid) {
    float mean[VN], thr[VN];   // thr: dy passes iff (x - mean) * sgn > thr  <=>  (x-mean)*scale+beta > 0
    float sgn[VN];
#pragma unroll
    for (int i = 0; i < VN; ++i) {
      mean[i] = bnp[c0 + i];
      const float sc = bnp[2 * C + c0 + i], be = bnp[3 * C + c0 + i];
      sgn[i] = sc;
      thr[i] = be;                        // same (x-mean)*scale+beta expression as bn_apply -> identical ReLU mask
    }
    const int64_t rstep = (int64_t)gridDim.x * blockDim.y;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.y + threadIdx.y; r < R; r += rstep * BUNR) {
      Pack<T> d[BUNR], xv[BUNR], o[BUNR];
      bool ok[BUNR];
#pragma unroll
      for (int u = 0; u < BUNR; ++u) {
        const int64_t rr = r + u * rstep;
        ok[u] = rr < R;
        if (ok[u]) {
          d[u].load(dout + rr * ldd + c0);
          xv[u].load(x + rr * ldx + c0);
          if (out) o[u].load(out + rr * ldo + c0);
        }
      }
#pragma unroll
      for (int u = 0; u < BUNR; ++u) {
        if (!ok[u]) continue;
#pragma unroll
        for (int i = 0; i < VN; ++i) {
          const float xr = xv[u].get(i);
          const float xc = xr - mean[i];
          float dy = d[u].get(i);
          if (out) dy = o[u].get(i) > 0.f ? dy : 0.f;
          else if (relu_from_x) dy = fmaf(xc, sgn[i], thr[i]) > 0.f ? dy : 0.f;
          fs[i] += dy;
          fq[i] = fmaf(dy, xc, fq[i]);      // istd is applied once at the end
        }
      }
    }
  }
  double s[VN], q[VN];
#pragma unroll
  for (int i = 0; i < VN; ++i) {
    s[i] = (double)fs[i];
    q[i] = valid ? (double)fq[i] * (double)bnp[C + c0 + i] : 0.0;
  }
  double* rep = dsums + (size_t)(blockIdx.x % NREP) * 2 * C;
  const bool last = reduce_and_ticket<VN>(s, q, rep, rep + C, c0, valid, counter);
  if (last && coef) {
    for (int c = threadIdx.y * blockDim.x + threadIdx.x; c < C; c += blockDim.x * blockDim.y) {
      double s1 = 0, s2 = 0;
#pragma unroll
      for (int r = 0; r < NREP; ++r) {
        s1 += __ldcg(dsums + (size_t)r * 2 * C + c);
        s2 += __ldcg(dsums + (size_t)r * 2 * C + C + c);
      }
      dbeta[c] += (float)s1;
      dgamma[c] += (float)s2;
      coef[c] = (float)(s1 / count);
      coef[C + c] = (float)(s2 / count);
    }
  }
}

__global__ void bn_bwd_finalize_kernel(const double* __restrict__ dsums, double count, float* __restrict__ dgamma,
                                       float* __restrict__ dbeta, float* __restrict__ coef, int C) {
  pdl_prologue();
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s1 = 0, s2 = 0;
  for (int r = 0; r < NREP; ++r) {
    s1 += dsums[(size_t)r * 2 * C + c];
    s2 += dsums[(size_t)r * 2 * C + C + c];
  }
  dbeta[c] += (float)s1;
  dgamma[c] += (float)s2;
  coef[c] = (float)(s1 / count);
  coef[C + c] = (float)(s2 / count);
}

// ---- backward apply: dx = scale*(dy - c1 - xhat*c2) = A*dy - B - (x-mean)*Cc ; dres (+)= dy
template <typename T>
__global__ void __launch_bounds__(256, 2)
bn_bwd_apply_kernel(const T* __restrict__ dout, int ldd, const T* __restrict__ out, int ldo, const T* __restrict__ x,
                    int ldx, const float* __restrict__ bnp, const float* __restrict__ coef, int relu_from_x,
                    T* __restrict__ dx, int lddx, T* __restrict__ dres, int lddr, int dres_acc, int64_t R, int C) {
  pdl_prologue();
  constexpr int VN = Pack<T>::N;
  const int cg = blockIdx.y * blockDim.x + threadIdx.x;
  if (cg * VN >= C) return;
  const int c0 = cg * VN;
  float mean[VN], A[VN], Bc[VN], Cc[VN], beta[VN];
#pragma unroll
  for (int i = 0; i < VN; ++i) {
    mean[i] = bnp[c0 + i];
    const float istd = bnp[C + c0 + i];
    A[i] = bnp[2 * C + c0 + i];
    beta[i] = bnp[3 * C + c0 + i];
    Bc[i] = A[i] * coef[c0 + i];
    Cc[i] = A[i] * istd * coef[C + c0 + i];
  }
  const int64_t rstep = (int64_t)gridDim.x * blockDim.y;
  // rows are visited from the end: the reduce pass that ran just before touched the tail last, so it is the part
  // of dout / x / out most likely still in L2
  for (int64_t rb = (int64_t)blockIdx.x * blockDim.y + threadIdx.y; rb < R; rb += rstep * BUNR) {
    Pack<T> d[BUNR], xv[BUNR], o[BUNR], dr[BUNR];
    int64_t rows[BUNR];
#pragma unroll
    for (int u = 0; u < BUNR; ++u) {
      const int64_t rr = rb + u * rstep;
      rows[u] = rr < R ? (R - 1 - rr) : -1;
      if (rows[u] >= 0) {
        d[u].load(dout + rows[u] * ldd + c0);
        xv[u].load(x + rows[u] * ldx + c0);
        if (out) o[u].load(out + rows[u] * ldo + c0);
        if (dres && dres_acc) dr[u].load(dres + rows[u] * lddr + c0);
      }
    }
#pragma unroll
    for (int u = 0; u < BUNR; ++u) {
      if (rows[u] < 0) continue;
#pragma unroll
      for (int i2 = 0; i2 < VN / 2; ++i2) {
        float res[2], drs[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int i = 2 * i2 + h;
          const float xr = xv[u].get(i);
          const float xc = xr - mean[i];
          float dy = d[u].get(i);
          if (out) dy = o[u].get(i) > 0.f ? dy : 0.f;
          else if (relu_from_x) dy = fmaf(xc, A[i], beta[i]) > 0.f ? dy : 0.f;
          drs[h] = (dres && dres_acc) ? dr[u].get(i) + dy : dy;
          res[h] = fmaf(A[i], dy, -Bc[i]) - xc * Cc[i];
        }
        d[u].set2(i2, res[0], res[1]);
        if (dres) dr[u].set2(i2, drs[0], drs[1]);
      }
      if (dres) dr[u].store(dres + rows[u] * lddr + c0);
      d[u].store(dx + rows[u] * lddx + c0);
    }
  }
}


// ==========================================================================================================
// Shared-memory-staged streaming (the B200 way to keep an HBM-bound kernel fed): one elected thread issues 1-D bulk
// copies (cp.async.bulk, completion on an mbarrier) of whole row slabs of every input tensor into an NST-deep ring,
// all threads consume 128-bit fragments from shared memory.  Bytes in flight per SM = the ring (about 100 KB per
// block, 2 blocks/SM), independent of the register budget -- the register-staged kernels above topped out at
// 1-2.6 TB/s with 125 registers and 2 rows in flight (ncu, profiles/).
// ==========================================================================================================
constexpr int SNT_MAX = 4;
struct StreamArgs {
  const char* src[SNT_MAX];
  long long ldb[SNT_MAX];   // row pitch in bytes
  int row_bytes;            // C * sizeof(T) == blockDim.x * 16
  long long R;
  int rows_per_stage;       // multiple of blockDim.y
  int nst;
  int reverse;              // visit slabs from the end (L2 reuse after a forward-order pass)
  const unsigned char* bits;   // optional packed ReLU mask, [R][row_bytes / 16] bytes (bit i of byte tx = channel 8*tx+i)
};

template <typename T, int NT, bool BITS, typename Body>
__device__ __forceinline__ void stream_rows(const StreamArgs& a, Body&& body) {
  extern __shared__ __align__(128) uint8_t stream_smem[];
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  const int SR = a.rows_per_stage;
  const int stage_bytes = SR * a.row_bytes;   // per tensor
  const uint32_t bar0 = smem_u32(stream_smem + (size_t)a.nst * NT * stage_bytes);
  // mask bytes of a stage: one byte per 16-byte fragment, staged behind the barriers
  const int bx = a.row_bytes / 16;
  unsigned char* bits_s = stream_smem + (size_t)a.nst * NT * stage_bytes + (((size_t)8 * a.nst + 15) & ~(size_t)15);
  if (tid == 0) {
    for (int i = 0; i < a.nst; ++i) mbar_init(bar0 + 8 * i, 1);
    fence_mbar_init();
  }
  __syncthreads();
  const long long n_slabs = (a.R + SR - 1) / SR;
  const long long nk = n_slabs > blockIdx.x ? (n_slabs - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  auto slab_of = [&](long long k) {
    const long long sidx = blockIdx.x + k * gridDim.x;
    return a.reverse ? n_slabs - 1 - sidx : sidx;
  };
  auto issue = [&](long long k) {
    const long long row0 = slab_of(k) * SR;
    const int rows = (int)min((long long)SR, a.R - row0);
    const int st = (int)(k % a.nst);
    const uint32_t bar = bar0 + 8 * st;
    mbar_expect_tx(bar, (uint32_t)(rows * a.row_bytes * NT + (BITS ? rows * bx : 0)));
    if (BITS) bulk_load_1d(smem_u32(bits_s + (size_t)st * SR * bx), a.bits + row0 * bx, (uint32_t)(rows * bx), bar);
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      const uint32_t dst = smem_u32(stream_smem + (size_t)(st * NT + t) * stage_bytes);
      const char* g = a.src[t] + row0 * a.ldb[t];
      if (a.ldb[t] == a.row_bytes) {
        bulk_load_1d(dst, g, (uint32_t)(rows * a.row_bytes), bar);
      } else {
        for (int r = 0; r < rows; ++r) bulk_load_1d(dst + r * a.row_bytes, g + r * a.ldb[t], (uint32_t)a.row_bytes, bar);
      }
    }
  };
  const bool warp0 = tid < 32;           // warp 0 issues the copies warp-uniformly through its elected lane
  if (warp0) {
    for (long long k = 0; k < nk && k < a.nst - 1; ++k) {
      if (elect_one()) issue(k);
      __syncwarp();
    }
  }
  for (long long k = 0; k < nk; ++k) {
    const int st = (int)(k % a.nst);
    if (warp0 && k + a.nst - 1 < nk) {   // refills the stage consumed in iteration k-1
      if (elect_one()) issue(k + a.nst - 1);
      __syncwarp();
    }
    mbar_wait(bar0 + 8 * st, (uint32_t)((k / a.nst) & 1));
    const long long row0 = slab_of(k) * SR;
    const int rows = (int)min((long long)SR, a.R - row0);
    constexpr int M = NT >= 4 ? 3 : 4;   // rows per thread in a full stage of the default geometry (see stream_geom)
    if (rows == SR && SR == M * (int)blockDim.y) {
      // full stage: all shared-memory loads of the thread's M rows are issued before any arithmetic
      Pack<T> f[M][NT];
      unsigned mb[M];
#pragma unroll
      for (int u = 0; u < M; ++u) {
        const int rr = threadIdx.y + u * blockDim.y;
#pragma unroll
        for (int t = 0; t < NT; ++t)
          f[u][t].load(reinterpret_cast<const T*>(stream_smem + (size_t)(st * NT + t) * stage_bytes +
                                                  (size_t)rr * a.row_bytes + threadIdx.x * 16));
        mb[u] = BITS ? (unsigned)bits_s[(size_t)st * SR * bx + (size_t)rr * bx + threadIdx.x] : 0u;
      }
#pragma unroll
      for (int u = 0; u < M; ++u) body(row0 + threadIdx.y + u * blockDim.y, f[u], mb[u]);
    } else {
      for (int rr = threadIdx.y; rr < rows; rr += blockDim.y) {
        Pack<T> f[NT];
#pragma unroll
        for (int t = 0; t < NT; ++t)
          f[t].load(reinterpret_cast<const T*>(stream_smem + (size_t)(st * NT + t) * stage_bytes +
                                               (size_t)rr * a.row_bytes + threadIdx.x * 16));
        body(row0 + rr, f, BITS ? (unsigned)bits_s[(size_t)st * SR * bx + (size_t)rr * bx + threadIdx.x] : 0u);
      }
    }
    __syncthreads();
  }
  // the ring memory may be reused by the caller (block reductions, a second pass with another geometry)
  if (tid == 0)
    for (int i = 0; i < a.nst; ++i) asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(bar0 + 8 * i) : "memory");
  __syncthreads();
}

template <typename T, bool HAS_RES, bool RES_BN>
__global__ void __launch_bounds__(256, 2) bn_apply_stream_kernel(const StreamArgs a, const float* __restrict__ bnp,
                                                                 const float* __restrict__ rbnp, int relu,
                                                                 T* __restrict__ out, int ldo, int C,
                                                                 unsigned char* __restrict__ maskbits) {
  pdl_prologue();
  constexpr int VN = Pack<T>::N;
  const int c0 = threadIdx.x * VN;
  float mean[VN], scale[VN], beta[VN], rmean[VN], rscale[VN];
#pragma unroll
  for (int i = 0; i < VN; ++i) {
    mean[i] = bnp[c0 + i];
    scale[i] = bnp[2 * C + c0 + i];
    beta[i] = bnp[3 * C + c0 + i];
    if (RES_BN) {
      rmean[i] = rbnp[c0 + i];
      rscale[i] = rbnp[2 * C + c0 + i];
      beta[i] += rbnp[3 * C + c0 + i];
    }
  }
  stream_rows<T, HAS_RES ? 2 : 1, false>(a, [&](long long row, Pack<T>(&f)[HAS_RES ? 2 : 1], unsigned) {
    Pack<T> o;
    unsigned mb = 0;
#pragma unroll
    for (int i2 = 0; i2 < VN / 2; ++i2) {
      float v[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int i = 2 * i2 + h;
        float t = fmaf(f[0].get(i) - mean[i], scale[i], beta[i]);
        if (HAS_RES) t = RES_BN ? fmaf(f[HAS_RES ? 1 : 0].get(i) - rmean[i], rscale[i], t) : t + f[HAS_RES ? 1 : 0].get(i);
        v[h] = relu ? fmaxf(t, 0.f) : t;
      }
      o.set2(i2, v[0], v[1]);
      // v >= 0 after the ReLU and bf16 keeps the fp32 exponent range, so (v > 0) == (stored bf16 value > 0) for every
      // normal number: the same mask as the backward kernels that read `out`
      if (v[0] > 0.f) mb |= 1u << (2 * i2);
      if (v[1] > 0.f) mb |= 1u << (2 * i2 + 1);
    }
    o.store(out + row * ldo + c0);
    if (maskbits) {
      // four neighbouring lanes (same row: blockDim.x is a multiple of 16) merge their bytes into one 32-bit store
      const unsigned am = __activemask();
      unsigned w = mb | (__shfl_down_sync(am, mb, 1) << 8);
      w |= __shfl_down_sync(am, w, 2) << 16;
      if ((threadIdx.x & 3) == 0)
        *reinterpret_cast<unsigned*>(maskbits + row * (a.row_bytes / 16) + threadIdx.x) = w;
    }
  });
}

// inputs: 0 = dout, 1 = x, 2 = out (MASK == 1); MASK == 2: packed mask bits instead of `out`
template <typename T, int MASK>
__global__ void __launch_bounds__(256, 2)
bn_bwd_reduce_stream_kernel(const StreamArgs a, const float* __restrict__ bnp, int relu_from_x, int C,
                            double* __restrict__ dsums, double count, float* __restrict__ dgamma,
                            float* __restrict__ dbeta, float* __restrict__ coef, unsigned int* counter) {
  pdl_prologue();
  constexpr int VN = Pack<T>::N;
  constexpr bool HAS_OUT = MASK == 1;
  constexpr int NT = HAS_OUT ? 3 : 2;
  const int c0 = threadIdx.x * VN;
  float fs[VN], fq[VN], mean[VN], sgn[VN], thr[VN];
#pragma unroll
  for (int i = 0; i < VN; ++i) {
    fs[i] = fq[i] = 0.f;
    mean[i] = bnp[c0 + i];
    sgn[i] = bnp[2 * C + c0 + i];
    thr[i] = bnp[3 * C + c0 + i];
  }
  stream_rows<T, NT, MASK == 2>(a, [&](long long, Pack<T>(&f)[NT], unsigned mb) {
#pragma unroll
    for (int i = 0; i < VN; ++i) {
      const float xc = f[1].get(i) - mean[i];
      float dy = f[0].get(i);
      if (HAS_OUT) dy = f[NT - 1].get(i) > 0.f ? dy : 0.f;
      else if (MASK == 2) dy = ((mb >> i) & 1u) ? dy : 0.f;
      else if (relu_from_x) dy = fmaf(xc, sgn[i], thr[i]) > 0.f ? dy : 0.f;
      fs[i] += dy;
      fq[i] = fmaf(dy, xc, fq[i]);
    }
  });
  // the ring is idle now (stream_rows ends with __syncthreads): reduce_and_ticket reuses the dynamic shared memory
  double s[VN], q[VN];
#pragma unroll
  for (int i = 0; i < VN; ++i) {
    s[i] = (double)fs[i];
    q[i] = (double)fq[i] * (double)bnp[C + c0 + i];
  }
  double* rep = dsums + (size_t)(blockIdx.x % NREP) * 2 * C;
  const bool last = reduce_and_ticket<VN>(s, q, rep, rep + C, c0, true, counter);
  if (last && coef) {
    for (int c = threadIdx.y * blockDim.x + threadIdx.x; c < C; c += blockDim.x * blockDim.y) {
      double s1 = 0, s2 = 0;
#pragma unroll
      for (int r = 0; r < NREP; ++r) {
        s1 += __ldcg(dsums + (size_t)r * 2 * C + c);
        s2 += __ldcg(dsums + (size_t)r * 2 * C + C + c);
      }
      dbeta[c] += (float)s1;
      dgamma[c] += (float)s2;
      coef[c] = (float)(s1 / count);
      coef[C + c] = (float)(s2 / count);
    }
  }
}

// inputs: 0 = dout, 1 = x, then out (MASK == 1), then the old dres (DRES_ACC); MASK == 2: packed mask bits
template <typename T, int MASK, bool DRES_ACC>
__global__ void __launch_bounds__(256, 2)
bn_bwd_apply_stream_kernel(const StreamArgs a, const float* __restrict__ bnp, const float* __restrict__ coef,
                           int relu_from_x, int C, T* __restrict__ dx, int lddx, T* __restrict__ dres, int lddr) {
  pdl_prologue();
  constexpr int VN = Pack<T>::N;
  constexpr bool HAS_OUT = MASK == 1;
  constexpr int NT = 2 + (HAS_OUT ? 1 : 0) + (DRES_ACC ? 1 : 0);
  const int c0 = threadIdx.x * VN;
  float mean[VN], A[VN], Bc[VN], Cc[VN], beta[VN];
#pragma unroll
  for (int i = 0; i < VN; ++i) {
    mean[i] = bnp[c0 + i];
    const float istd = bnp[C + c0 + i];
    A[i] = bnp[2 * C + c0 + i];
    beta[i] = bnp[3 * C + c0 + i];
    Bc[i] = A[i] * coef[c0 + i];
    Cc[i] = A[i] * istd * coef[C + c0 + i];
  }
  stream_rows<T, NT, MASK == 2>(a, [&](long long row, Pack<T>(&f)[NT], unsigned mb) {
    Pack<T> o, dr;
#pragma unroll
    for (int i2 = 0; i2 < VN / 2; ++i2) {
      float res[2], drs[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int i = 2 * i2 + h;
        const float xc = f[1].get(i) - mean[i];
        float dy = f[0].get(i);
        if (HAS_OUT) dy = f[HAS_OUT ? 2 : 0].get(i) > 0.f ? dy : 0.f;
        else if (MASK == 2) dy = ((mb >> i) & 1u) ? dy : 0.f;
        else if (relu_from_x) dy = fmaf(xc, A[i], beta[i]) > 0.f ? dy : 0.f;
        drs[h] = DRES_ACC ? f[NT - 1].get(i) + dy : dy;
        res[h] = fmaf(A[i], dy, -Bc[i]) - xc * Cc[i];
      }
      o.set2(i2, res[0], res[1]);
      dr.set2(i2, drs[0], drs[1]);
    }
    if (dres) dr.store(dres + row * lddr + c0);
    o.store(dx + row * lddx + c0);
  });
}

// ---- host side of the streamed kernels
struct StreamGeom {
  bool ok;
  dim3 grid, block;
  size_t smem;
  int rows_per_stage, nst;
};
static StreamGeom stream_geom(int64_t R, int C, int es, int nt, size_t min_smem, bool bits = false) {
  StreamGeom g{};
  const int row_bytes = C * es;
  const int bx = row_bytes / 16;
  g.ok = false;
  if (row_bytes % 16 || bx < 1 || bx > 256 || (bx & (bx - 1))) return g;   // one 16-byte fragment per thread per row
  const int by = 256 / bx;
  // Stage size: 16 KB per tensor (4 x 256 fragments).  The per-stage cost (mbarrier wake-up, issue of the next
  // copies, block barrier) is what paces these kernels, not the bytes in flight: measured on the training step,
  // 8 KB stages x 4-6 deep = 10.8 ms, 16 KB stages x 2-3 deep = 10.25 ms, 32 KB stages (ring too shallow) = 12.0 ms.
  static int sr_env = -1;
  if (sr_env < 0) {
    const char* e = exp_env("BASI_BN_STAGE_ROWS");
    sr_env = (e && atoi(e) >= 1 && atoi(e) <= 16) ? atoi(e) : 0;
  }
  int sr_mult = sr_env ? sr_env : (12 / nt < 4 ? 12 / nt : 4);
  if (sr_mult < 1) sr_mult = 1;
  const int SR = sr_mult * by;
  const int64_t n_slabs = (R + SR - 1) / SR;
  if (n_slabs < 64) return g;                                              // tiny tensors: the direct kernels
  {
    // a small single-tensor pass (forward BN+ReLU of the conv2..conv5 bottleneck layers: ~22 KB per CTA) gains
    // nothing from the ring; the register-staged kernel measured 0.07 ms/step faster there
    static int min1 = -1;
    if (min1 < 0) min1 = exp_env("BASI_BN_STREAM_MIN1") ? atoi(exp_env("BASI_BN_STREAM_MIN1")) : 1000;
    if (nt == 1 && n_slabs < min1) return g;
  }
  static int ring_kb = 0;
  if (!ring_kb) {
    const char* e = exp_env("BASI_BN_RING_KB");
    ring_kb = (e && atoi(e) >= 32 && atoi(e) <= 220) ? atoi(e) : 96;
  }
  int nst = (int)(((size_t)ring_kb * 1024) / ((size_t)nt * SR * row_bytes));
  if (nst > 8) nst = 8;
  if (nst < 2) return g;
  g.rows_per_stage = SR;
  g.nst = nst;
  g.block = dim3(bx, by);
  const int64_t cap = 2 * (int64_t)sm_count();
  g.grid = dim3((unsigned)(n_slabs < cap ? n_slabs : cap));
  g.smem = (size_t)nst * nt * SR * row_bytes + 8 * nst + 16;
  if (bits) {
    if (bx % 16) return g;                               // mask rows are bulk-copied: 16-byte granularity
    g.smem += (size_t)nst * SR * bx + 16;
  }
  if (g.smem < min_smem) g.smem = min_smem;
  g.ok = true;
  return g;
}
static void fill_stream_args(StreamArgs* a, const StreamGeom& g, int64_t R, int C, int es, int reverse) {
  a->row_bytes = C * es;
  a->R = R;
  a->rows_per_stage = g.rows_per_stage;
  a->nst = g.nst;
  a->reverse = reverse;
}
template <typename K>
static void allow_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}
static bool stream_disabled() {
  static int v = -1;
  if (v < 0) v = exp_env("BASI_BN_DIRECT") ? 1 : 0;
  return v == 1;
}

template <typename T>
static bool launch_apply_stream(const basi_tensor* x, const float* bnp, const basi_tensor* res, const float* rbnp,
                                int relu, const basi_tensor* out, cudaStream_t st, unsigned char* maskbits = nullptr) {
  if (stream_disabled()) return false;
  const int es = sizeof(T);
  const int nt = res ? 2 : 1;
  const int64_t R = pixels(x);
  StreamGeom g = stream_geom(R, x->c, es, nt, 0);
  if (!g.ok) return false;
  StreamArgs a{};
  fill_stream_args(&a, g, R, x->c, es, 0);
  a.src[0] = (const char*)x->ptr; a.ldb[0] = (long long)x->ld * es;
  if (res) { a.src[1] = (const char*)res->ptr; a.ldb[1] = (long long)res->ld * es; }
  if (!res) {
    allow_smem(bn_apply_stream_kernel<T, false, false>, g.smem);
    basi::launch(bn_apply_stream_kernel<T, false, false>, g.grid, g.block, g.smem, st, a, bnp, nullptr, relu, (T*)out->ptr, out->ld, x->c, maskbits);
  } else if (!rbnp) {
    allow_smem(bn_apply_stream_kernel<T, true, false>, g.smem);
    basi::launch(bn_apply_stream_kernel<T, true, false>, g.grid, g.block, g.smem, st, a, bnp, nullptr, relu, (T*)out->ptr, out->ld, x->c, maskbits);
  } else {
    allow_smem(bn_apply_stream_kernel<T, true, true>, g.smem);
    basi::launch(bn_apply_stream_kernel<T, true, true>, g.grid, g.block, g.smem, st, a, bnp, rbnp, relu, (T*)out->ptr, out->ld, x->c, maskbits);
  }
  return true;
}

template <typename T>
static bool launch_bwd_reduce_stream(const basi_tensor* dout, const basi_tensor* out, const basi_tensor* x,
                                     const float* bnp, int relu_from_x, double* dsums, double count, float* dgamma,
                                     float* dbeta, float* coef, uint32_t* counter, cudaStream_t st,
                                     const unsigned char* bits = nullptr) {
  if (stream_disabled()) return false;
  const int es = sizeof(T);
  const int nt = out ? 3 : 2;
  const int64_t R = pixels(x);
  StreamGeom g = stream_geom(R, x->c, es, nt, (size_t)256 * 2 * Pack<T>::N * sizeof(double), bits != nullptr);
  if (!g.ok) return false;
  StreamArgs a{};
  fill_stream_args(&a, g, R, x->c, es, 0);
  a.bits = bits;
  a.src[0] = (const char*)dout->ptr; a.ldb[0] = (long long)dout->ld * es;
  a.src[1] = (const char*)x->ptr; a.ldb[1] = (long long)x->ld * es;
  if (out) { a.src[2] = (const char*)out->ptr; a.ldb[2] = (long long)out->ld * es; }
  if (bits) {
    allow_smem(bn_bwd_reduce_stream_kernel<T, 2>, g.smem);
    basi::launch(bn_bwd_reduce_stream_kernel<T, 2>, g.grid, g.block, g.smem, st, a, bnp, relu_from_x, x->c, dsums, count, dgamma, dbeta, coef, counter);
  } else if (out) {
    allow_smem(bn_bwd_reduce_stream_kernel<T, 1>, g.smem);
    basi::launch(bn_bwd_reduce_stream_kernel<T, 1>, g.grid, g.block, g.smem, st, a, bnp, relu_from_x, x->c, dsums, count, dgamma, dbeta, coef, counter);
  } else {
    allow_smem(bn_bwd_reduce_stream_kernel<T, 0>, g.smem);
    basi::launch(bn_bwd_reduce_stream_kernel<T, 0>, g.grid, g.block, g.smem, st, a, bnp, relu_from_x, x->c, dsums, count, dgamma, dbeta, coef, counter);
  }
  return true;
}

template <typename T>
static bool launch_bwd_apply_stream(const basi_tensor* dout, const basi_tensor* out, const basi_tensor* x,
                                    const float* bnp, const float* coef, int relu_from_x, const basi_tensor* dx,
                                    const basi_tensor* dres, int dres_acc, cudaStream_t st,
                                    const unsigned char* bits = nullptr) {
  if (stream_disabled()) return false;
  const int es = sizeof(T);
  const bool acc = dres && dres_acc;
  const int nt = 2 + (out ? 1 : 0) + (acc ? 1 : 0);
  const int64_t R = pixels(x);
  StreamGeom g = stream_geom(R, x->c, es, nt, 0, bits != nullptr);
  if (!g.ok) return false;
  StreamArgs a{};
  fill_stream_args(&a, g, R, x->c, es, 1);
  a.bits = bits;
  int k = 0;
  a.src[k] = (const char*)dout->ptr; a.ldb[k++] = (long long)dout->ld * es;
  a.src[k] = (const char*)x->ptr; a.ldb[k++] = (long long)x->ld * es;
  if (out) { a.src[k] = (const char*)out->ptr; a.ldb[k++] = (long long)out->ld * es; }
  if (acc) { a.src[k] = (const char*)dres->ptr; a.ldb[k++] = (long long)dres->ld * es; }
  T* dxp = (T*)dx->ptr;
  T* drp = dres ? (T*)dres->ptr : nullptr;
  const int lddr = dres ? dres->ld : 0;
#define BASI_LAUNCH_BWD_APPLY(HO, DA)                                                                               \
  do {                                                                                                              \
    allow_smem(bn_bwd_apply_stream_kernel<T, HO, DA>, g.smem);                                                      \
    basi::launch(bn_bwd_apply_stream_kernel<T, HO, DA>, g.grid, g.block, g.smem, st, a, bnp, coef, relu_from_x, x->c, dxp,    \
                                                                            dx->ld, drp, lddr);                     \
  } while (0)
  if (bits && acc) BASI_LAUNCH_BWD_APPLY(2, true);
  else if (bits) BASI_LAUNCH_BWD_APPLY(2, false);
  else if (out && acc) BASI_LAUNCH_BWD_APPLY(1, true);
  else if (out) BASI_LAUNCH_BWD_APPLY(1, false);
  else if (acc) BASI_LAUNCH_BWD_APPLY(0, true);
  else BASI_LAUNCH_BWD_APPLY(0, false);
#undef BASI_LAUNCH_BWD_APPLY
  return true;
}

// ==========================================================================================================
// Resident backward: reduce + apply in ONE cooperative launch.  The pixel rows are split evenly over one CTA per SM;
// each CTA bulk-copies its slab of dout and x into shared memory ONCE (the 148 x ~200 KB of shared memory hold both
// tensors of every plain BN+ReLU layer from conv2 on), reduces its slab, meets the other CTAs at a grid barrier,
// then computes dx from the resident copies.  HBM traffic: read 2 tensors + write 1, instead of read 4 + write 1
// for the reduce/apply pair, and one launch + one tail instead of two.
// ==========================================================================================================
struct ResidentArgs {
  const char* dout;
  const char* x;
  char* dx;
  long long R;
  int row_bytes;       // C * sizeof(T) == blockDim.x * 16
  int rows_per_cta;
  int chunk_rows;      // rows per bulk copy (per tensor)
  int n_chunks;
  int debug;           // BASI_BN_RESIDENT_DEBUG: CTA 0 records globaltimer stamps (start, sums done, barrier passed,
                       // coefficients ready, end) into g_resident_dbg
};

__device__ unsigned long long g_resident_dbg[8];
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

template <typename T, int NTHR>
__global__ void __launch_bounds__(NTHR, 1)
bn_bwd_resident_kernel(const ResidentArgs a, const float* __restrict__ bnp, int relu_from_x, int C,
                       double* __restrict__ dsums, double count, float* __restrict__ dgamma,
                       float* __restrict__ dbeta, float* __restrict__ coef, unsigned int* gbar) {
  extern __shared__ __align__(128) uint8_t rs[];
  constexpr int VN = Pack<T>::N;
  const int tx = threadIdx.x, ty = threadIdx.y, by = blockDim.y;
  const int tid = ty * blockDim.x + tx;
  const bool dbg = a.debug && blockIdx.x == 0 && tid == 0;
  if (dbg) g_resident_dbg[0] = gtimer();
  const size_t slab = (size_t)a.rows_per_cta * a.row_bytes;
  uint8_t* s_dout = rs;
  uint8_t* s_x = rs + slab;
  float* red = reinterpret_cast<float*>(rs + 2 * slab);        // [by][C][2]
  float* coef_s = red + (size_t)by * C * 2;                    // [2][C]
  const uint32_t bar0 = smem_u32(coef_s + 2 * C);
  const long long row0 = (long long)blockIdx.x * a.rows_per_cta;
  long long left = a.R - row0;
  const int rows = left <= 0 ? 0 : (left < a.rows_per_cta ? (int)left : a.rows_per_cta);
  if (tid == 0) {
    for (int k = 0; k < a.n_chunks; ++k) mbar_init(bar0 + 8 * k, 1);
    fence_mbar_init();
  }
  __syncthreads();
  pdl_prologue();     // launched with the cooperative AND the programmatic-serialization attribute (see launch site)
  if (tid == 0) {
    for (int k = 0; k < a.n_chunks; ++k) {
      const int r0 = k * a.chunk_rows;
      const int n = min(a.chunk_rows, rows - r0);
      if (n <= 0) break;
      const uint32_t bytes = (uint32_t)n * a.row_bytes;
      mbar_expect_tx(bar0 + 8 * k, 2 * bytes);
      bulk_load_1d(smem_u32(s_dout + (size_t)r0 * a.row_bytes), a.dout + (row0 + r0) * a.row_bytes, bytes, bar0 + 8 * k);
      bulk_load_1d(smem_u32(s_x + (size_t)r0 * a.row_bytes), a.x + (row0 + r0) * a.row_bytes, bytes, bar0 + 8 * k);
    }
  }
  const int c0 = tx * VN;
  float fs[VN], fq[VN], mean[VN], A[VN], beta[VN];
#pragma unroll
  for (int i = 0; i < VN; ++i) {
    fs[i] = fq[i] = 0.f;
    mean[i] = bnp[c0 + i];
    A[i] = bnp[2 * C + c0 + i];
    beta[i] = bnp[3 * C + c0 + i];
  }
  // ---- phase 1: per-channel sums of dy and dy * (x - mean) over the resident slab
  for (int k = 0; k < a.n_chunks; ++k) {
    const int r0 = k * a.chunk_rows;
    const int n = min(a.chunk_rows, rows - r0);
    if (n <= 0) break;
    mbar_wait(bar0 + 8 * k, 0);
    for (int rr = r0 + ty; rr < r0 + n; rr += by) {
      Pack<T> d, xv;
      d.load(reinterpret_cast<const T*>(s_dout + (size_t)rr * a.row_bytes + tx * 16));
      xv.load(reinterpret_cast<const T*>(s_x + (size_t)rr * a.row_bytes + tx * 16));
#pragma unroll
      for (int i = 0; i < VN; ++i) {
        const float xc = xv.get(i) - mean[i];
        float dy = d.get(i);
        if (relu_from_x) dy = fmaf(xc, A[i], beta[i]) > 0.f ? dy : 0.f;
        fs[i] += dy;
        fq[i] = fmaf(dy, xc, fq[i]);
      }
    }
  }
  if (dbg) g_resident_dbg[5] = gtimer();
#pragma unroll
  for (int i = 0; i < VN; ++i) {
    red[((size_t)ty * C + c0 + i) * 2] = fs[i];
    red[((size_t)ty * C + c0 + i) * 2 + 1] = fq[i];
  }
  __syncthreads();
  double* rep = dsums + (size_t)(blockIdx.x % NREP) * 2 * C;
  for (int c = tid; c < C; c += NTHR) {
    double s1 = 0, s2 = 0;
    for (int y = 0; y < by; ++y) {
      s1 += (double)red[((size_t)y * C + c) * 2];
      s2 += (double)red[((size_t)y * C + c) * 2 + 1];
    }
    atomicAdd(rep + c, s1);
    atomicAdd(rep + C + c, s2 * (double)bnp[C + c]);
  }
  __syncthreads();   // (thread 0's fence inside the barrier is cumulative over the atomics ordered before this sync)
  if (dbg) g_resident_dbg[1] = gtimer();
  if (tid == 0) grid_barrier_thread0(gbar);
  __syncthreads();
  if (dbg) g_resident_dbg[2] = gtimer();
  // ---- every CTA turns the (now final) sums into the two per-channel coefficients; CTA 0 publishes the gradients
  for (int c = tid; c < C; c += NTHR) {
    double s1 = 0, s2 = 0;
#pragma unroll
    for (int r = 0; r < NREP; ++r) {
      s1 += __ldcg(dsums + (size_t)r * 2 * C + c);
      s2 += __ldcg(dsums + (size_t)r * 2 * C + C + c);
    }
    const float k1 = (float)(s1 / count), k2 = (float)(s2 / count);
    coef_s[c] = k1;
    coef_s[C + c] = k2;
    if (blockIdx.x == 0) {
      dbeta[c] += (float)s1;
      dgamma[c] += (float)s2;
      if (coef) {
        coef[c] = k1;
        coef[C + c] = k2;
      }
    }
  }
  __syncthreads();
  // ---- phase 2: dx = scale * (dy - mean(dy)) - (x - mean) * scale * istd * mean(dy * xhat), from shared memory
  float Bc[VN], Cc[VN];
#pragma unroll
  for (int i = 0; i < VN; ++i) {
    Bc[i] = A[i] * coef_s[c0 + i];
    Cc[i] = A[i] * bnp[C + c0 + i] * coef_s[C + c0 + i];
  }
  if (dbg) g_resident_dbg[3] = gtimer();
  T* dxp = reinterpret_cast<T*>(a.dx + row0 * a.row_bytes);
  for (int rr = ty; rr < rows; rr += by) {
    Pack<T> d, xv, o;
    d.load(reinterpret_cast<const T*>(s_dout + (size_t)rr * a.row_bytes + tx * 16));
    xv.load(reinterpret_cast<const T*>(s_x + (size_t)rr * a.row_bytes + tx * 16));
#pragma unroll
    for (int i2 = 0; i2 < VN / 2; ++i2) {
      float res[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int i = 2 * i2 + h;
        const float xc = xv.get(i) - mean[i];
        float dy = d.get(i);
        if (relu_from_x) dy = fmaf(xc, A[i], beta[i]) > 0.f ? dy : 0.f;
        res[h] = fmaf(A[i], dy, -Bc[i]) - xc * Cc[i];
      }
      o.set2(i2, res[0], res[1]);
    }
    o.store(dxp + (size_t)rr * (a.row_bytes / sizeof(T)) + c0);
  }
  if (dbg) g_resident_dbg[4] = gtimer();
}

// ==========================================================================================================
// Cooperative streamed backward for the tensors that do not fit the resident kernel (residual junctions, the
// 160x160 layers): the reduce pass and the dx pass of the pair above in ONE launch with a grid barrier in between.
// The second pass visits the rows in reverse, so most of what it re-reads is still in the 126 MB L2; one launch and
// one tail instead of two.
// ==========================================================================================================
template <typename T, int MASK, bool DRES_ACC>
__global__ void __launch_bounds__(256, 2)
bn_bwd_coop_kernel(const StreamArgs a1, const StreamArgs a2, const float* __restrict__ bnp, int relu_from_x, int C,
                   double* __restrict__ dsums, double count, float* __restrict__ dgamma, float* __restrict__ dbeta,
                   float* __restrict__ coef, unsigned int* gbar, T* __restrict__ dx, int lddx, T* __restrict__ dres,
                   int lddr) {
  extern __shared__ __align__(128) uint8_t coop_smem[];
  constexpr int VN = Pack<T>::N;
  constexpr bool HAS_OUT = MASK == 1;
  constexpr int NT1 = HAS_OUT ? 3 : 2;
  constexpr int NT2 = NT1 + (DRES_ACC ? 1 : 0);
  const int c0 = threadIdx.x * VN;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  pdl_prologue();     // cooperative + programmatic-serialization launch: the launch latency overlaps the predecessor
  float mean[VN], A[VN], beta[VN];
#pragma unroll
  for (int i = 0; i < VN; ++i) {
    mean[i] = bnp[c0 + i];
    A[i] = bnp[2 * C + c0 + i];
    beta[i] = bnp[3 * C + c0 + i];
  }
  // ---- pass 1: per-channel sums
  {
    float fs[VN], fq[VN];
#pragma unroll
    for (int i = 0; i < VN; ++i) fs[i] = fq[i] = 0.f;
    stream_rows<T, NT1, MASK == 2>(a1, [&](long long, Pack<T>(&f)[NT1], unsigned mb) {
#pragma unroll
      for (int i = 0; i < VN; ++i) {
        const float xc = f[1].get(i) - mean[i];
        float dy = f[0].get(i);
        if (HAS_OUT) dy = f[NT1 - 1].get(i) > 0.f ? dy : 0.f;
        else if (MASK == 2) dy = ((mb >> i) & 1u) ? dy : 0.f;
        else if (relu_from_x) dy = fmaf(xc, A[i], beta[i]) > 0.f ? dy : 0.f;
        fs[i] += dy;
        fq[i] = fmaf(dy, xc, fq[i]);
      }
    });
    // block reduction over threadIdx.y in the (idle) ring memory, one double atomic per channel and block
    float* red = reinterpret_cast<float*>(coop_smem);      // [blockDim.y][C][2]
#pragma unroll
    for (int i = 0; i < VN; ++i) {
      red[((size_t)threadIdx.y * C + c0 + i) * 2] = fs[i];
      red[((size_t)threadIdx.y * C + c0 + i) * 2 + 1] = fq[i];
    }
    __syncthreads();
    double* rep = dsums + (size_t)(blockIdx.x % NREP) * 2 * C;
    for (int c = tid; c < C; c += 256) {
      double s1 = 0, s2 = 0;
      for (int y = 0; y < (int)blockDim.y; ++y) {
        s1 += (double)red[((size_t)y * C + c) * 2];
        s2 += (double)red[((size_t)y * C + c) * 2 + 1];
      }
      atomicAdd(rep + c, s1);
      atomicAdd(rep + C + c, s2 * (double)bnp[C + c]);
    }
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) grid_barrier_thread0(gbar);
  __syncthreads();
  // ---- final sums -> coefficients (every CTA), gradients of gamma / beta (CTA 0)
  float* coef_s = reinterpret_cast<float*>(coop_smem);     // [2][C]
  for (int c = tid; c < C; c += 256) {
    double s1 = 0, s2 = 0;
#pragma unroll
    for (int r = 0; r < NREP; ++r) {
      s1 += __ldcg(dsums + (size_t)r * 2 * C + c);
      s2 += __ldcg(dsums + (size_t)r * 2 * C + C + c);
    }
    const float k1 = (float)(s1 / count), k2 = (float)(s2 / count);
    coef_s[c] = k1;
    coef_s[C + c] = k2;
    if (blockIdx.x == 0) {
      dbeta[c] += (float)s1;
      dgamma[c] += (float)s2;
      if (coef) {
        coef[c] = k1;
        coef[C + c] = k2;
      }
    }
  }
  __syncthreads();
  float Bc[VN], Cc[VN];
#pragma unroll
  for (int i = 0; i < VN; ++i) {
    Bc[i] = A[i] * coef_s[c0 + i];
    Cc[i] = A[i] * bnp[C + c0 + i] * coef_s[C + c0 + i];
  }
  __syncthreads();   // coef_s lives in the ring memory pass 2 is about to overwrite
  // ---- pass 2: dx (and the residual-path gradient)
  stream_rows<T, NT2, MASK == 2>(a2, [&](long long row, Pack<T>(&f)[NT2], unsigned mb) {
    Pack<T> o, dr;
#pragma unroll
    for (int i2 = 0; i2 < VN / 2; ++i2) {
      float res[2], drs[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int i = 2 * i2 + h;
        const float xc = f[1].get(i) - mean[i];
        float dy = f[0].get(i);
        if (HAS_OUT) dy = f[HAS_OUT ? 2 : 0].get(i) > 0.f ? dy : 0.f;
        else if (MASK == 2) dy = ((mb >> i) & 1u) ? dy : 0.f;
        else if (relu_from_x) dy = fmaf(xc, A[i], beta[i]) > 0.f ? dy : 0.f;
        drs[h] = DRES_ACC ? f[NT2 - 1].get(i) + dy : dy;
        res[h] = fmaf(A[i], dy, -Bc[i]) - xc * Cc[i];
      }
      o.set2(i2, res[0], res[1]);
      dr.set2(i2, drs[0], drs[1]);
    }
    if (dres) dr.store(dres + row * lddr + c0);
    o.store(dx + row * lddx + c0);
  });
}

static bool coop_disabled() {
  static int v = -1;
  if (v < 0) v = exp_env("BASI_BN_NO_COOP") ? 1 : 0;
  return v == 1;
}

// returns false when the tensor does not take the cooperative streamed path
template <typename T>
static bool launch_bwd_coop(bool dry, const basi_tensor* dout, const basi_tensor* out, const unsigned char* bits,
                            const basi_tensor* x, const float* bnp, int relu_from_x, double* dsums, double count,
                            float* dgamma, float* dbeta, float* coef, uint32_t* gbar, const basi_tensor* dx,
                            const basi_tensor* dres, int dres_acc, cudaStream_t st) {
  if (stream_disabled() || coop_disabled()) return false;
  const int es = sizeof(T);
  const bool acc = dres && dres_acc;
  const int nt1 = out ? 3 : 2, nt2 = nt1 + (acc ? 1 : 0);
  const int64_t R = pixels(x);
  const size_t red_bytes = (size_t)(256 / (x->c * es / 16 > 0 ? x->c * es / 16 : 1)) * x->c * 2 * sizeof(float);
  StreamGeom g1 = stream_geom(R, x->c, es, nt1, red_bytes, bits != nullptr);
  StreamGeom g2 = stream_geom(R, x->c, es, nt2, 0, bits != nullptr);
  if (!g1.ok || !g2.ok) return false;
  const size_t smem = g1.smem > g2.smem ? g1.smem : g2.smem;
  if (smem > 110 * 1024) return false;                     // two CTAs per SM must stay co-resident
  if (dry) {
    // the grid barrier needs all 2 x #SMs CTAs co-resident: ask the occupancy calculator (another context sharing the
    // GPU, or a smaller shared-memory carve-out, makes this fail -> the caller uses the two-launch pair)
    int nb = 0;
#define BASI_COOP_OCC(MK, DA)                                                                                      \
  do {                                                                                                             \
    allow_smem(bn_bwd_coop_kernel<T, MK, DA>, smem);                                                               \
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, bn_bwd_coop_kernel<T, MK, DA>, 256, smem);                  \
  } while (0)
    if (bits && acc) BASI_COOP_OCC(2, true);
    else if (bits) BASI_COOP_OCC(2, false);
    else if (out && acc) BASI_COOP_OCC(1, true);
    else if (out) BASI_COOP_OCC(1, false);
    else if (acc) BASI_COOP_OCC(0, true);
    else BASI_COOP_OCC(0, false);
#undef BASI_COOP_OCC
    if (cudaGetLastError() != cudaSuccess) return false;
    return nb >= 2;
  }
  StreamArgs a1{}, a2{};
  fill_stream_args(&a1, g1, R, x->c, es, 0);
  fill_stream_args(&a2, g2, R, x->c, es, 1);
  a1.bits = a2.bits = bits;
  int k = 0;
  a1.src[k] = a2.src[k] = (const char*)dout->ptr; a1.ldb[k] = a2.ldb[k] = (long long)dout->ld * es; ++k;
  a1.src[k] = a2.src[k] = (const char*)x->ptr; a1.ldb[k] = a2.ldb[k] = (long long)x->ld * es; ++k;
  if (out) { a1.src[k] = a2.src[k] = (const char*)out->ptr; a1.ldb[k] = a2.ldb[k] = (long long)out->ld * es; ++k; }
  if (acc) { a2.src[k] = (const char*)dres->ptr; a2.ldb[k] = (long long)dres->ld * es; ++k; }
  T* dxp = (T*)dx->ptr;
  T* drp = dres ? (T*)dres->ptr : nullptr;
  const int lddr = dres ? dres->ld : 0;
  const int C = x->c;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * sm_count());
  cfg.blockDim = g1.block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (pdl_enabled() && !(exp_env("BASI_BN_COOP_PDL") && atoi(exp_env("BASI_BN_COOP_PDL")) == 0)) ? 2 : 1;
#define BASI_LAUNCH_COOP(MK, DA)                                                                                      \
  do {                                                                                                                \
    allow_smem(bn_bwd_coop_kernel<T, MK, DA>, smem);                                                                  \
    cudaLaunchKernelEx(&cfg, bn_bwd_coop_kernel<T, MK, DA>, a1, a2, bnp, relu_from_x, C, dsums, count, dgamma, dbeta, \
                       coef, (unsigned int*)gbar, dxp, (int)dx->ld, drp, lddr);                                       \
  } while (0)
  if (bits && acc) BASI_LAUNCH_COOP(2, true);
  else if (bits) BASI_LAUNCH_COOP(2, false);
  else if (out && acc) BASI_LAUNCH_COOP(1, true);
  else if (out) BASI_LAUNCH_COOP(1, false);
  else if (acc) BASI_LAUNCH_COOP(0, true);
  else BASI_LAUNCH_COOP(0, false);
#undef BASI_LAUNCH_COOP
  return true;
}

struct ResidentGeom {
  bool ok;
  int threads;
  int grid, bx, by, rows_per_cta, chunk_rows, n_chunks;
  size_t smem;
};
static bool resident_disabled() {
  static int v = -1;
  if (v < 0) v = exp_env("BASI_BN_NO_RESIDENT") ? 1 : 0;
  return v == 1;
}
static ResidentGeom resident_geom(int64_t R, int C, int es) {
  ResidentGeom g{};
  const int row_bytes = C * es;
  const int bx = row_bytes / 16;
  if (resident_disabled() || row_bytes % 16 || bx < 1 || bx > 256 || (bx & (bx - 1))) return g;
  const int G = sm_count();
  if (R < (int64_t)G * 8) return g;
  const int64_t rpc = (R + G - 1) / G;
  // Threads per CTA (one CTA per SM).  Measured phases at C = 128, 256 threads, cold / L2-warm: rows 5.2 / 2.6 us,
  // block reduction + atomics 2.2, grid barrier 1.6, coefficients 1.5 / 1.1, dx 1.6.  512 / 1024 threads make the
  // block reduction longer (2.2 -> 6.3 us) without speeding the row pass up, so 256 it is.
  static int thr_env = -1;
  if (thr_env < 0) thr_env = exp_env("BASI_BN_RESIDENT_THREADS") ? atoi(exp_env("BASI_BN_RESIDENT_THREADS")) : 0;
  int threads = (thr_env == 256 || thr_env == 512 || thr_env == 1024) ? thr_env : 256;
  while (threads > 256 && threads / bx > rpc) threads /= 2;
  if (threads < bx) threads = bx;
  int chunk_rows = (16 * 1024) / row_bytes;
  if (chunk_rows < 1) chunk_rows = 1;
  const int64_t n_chunks = (rpc + chunk_rows - 1) / chunk_rows;
  if (n_chunks > 64) return g;
  size_t smem = 0;
  int by = 1;
  for (;; threads /= 2) {          // fewer threads -> smaller block-reduction scratch, until the slab fits
    by = threads / bx;
    smem = 2 * (size_t)rpc * row_bytes + (size_t)by * C * 2 * sizeof(float) + 2 * (size_t)C * sizeof(float) +
           8 * (size_t)n_chunks + 128;
    if (smem <= 220 * 1024) break;
    if (threads <= 256 || threads / 2 < bx) return g;
  }
  g.threads = threads;
  g.ok = true;
  g.grid = (int)((R + rpc - 1) / rpc);
  g.bx = bx; g.by = by;
  g.rows_per_cta = (int)rpc;
  g.chunk_rows = chunk_rows;
  g.n_chunks = (int)n_chunks;
  g.smem = smem;
  return g;
}
static bool dense_rows(const basi_tensor* t) { return t->ld == t->c; }

template <typename T>
static void launch_bwd_resident(const ResidentGeom& g, const basi_tensor* dout, const basi_tensor* x, const float* bnp,
                                int relu_from_x, double* dsums, double count, float* dgamma, float* dbeta, float* coef,
                                uint32_t* gbar, const basi_tensor* dx, cudaStream_t st) {
  ResidentArgs a{};
  a.dout = (const char*)dout->ptr; a.x = (const char*)x->ptr; a.dx = (char*)dx->ptr;
  a.R = pixels(x);
  a.row_bytes = x->c * (int)sizeof(T);
  a.rows_per_cta = g.rows_per_cta; a.chunk_rows = g.chunk_rows; a.n_chunks = g.n_chunks;
  a.debug = exp_env("BASI_BN_RESIDENT_DEBUG") ? 1 : 0;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(g.grid);
  cfg.blockDim = dim3(g.bx, g.by);
  cfg.dynamicSmemBytes = g.smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeCooperative;     // all CTAs co-resident: the grid barrier cannot deadlock
  attr[0].val.cooperative = 1;
  // ... and programmatic dependent launch on top: the prologue (barrier init, launch latency) overlaps the tail of
  // the producing dgrad kernel; the kernel touches nothing before its griddepcontrol.wait (BASI_BN_COOP_PDL=0: off)
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  const bool coop_pdl = pdl_enabled() && !(exp_env("BASI_BN_COOP_PDL") && atoi(exp_env("BASI_BN_COOP_PDL")) == 0);
  cfg.attrs = attr;
  cfg.numAttrs = exp_env("BASI_BN_RESIDENT_NOCOOP") ? 0 : (coop_pdl ? 2 : 1);   // (NOCOOP: experiment only)
#define BASI_LAUNCH_RESIDENT(NT_)                                                                                  \
  do {                                                                                                             \
    allow_smem(bn_bwd_resident_kernel<T, NT_>, g.smem);                                                            \
    cudaLaunchKernelEx(&cfg, bn_bwd_resident_kernel<T, NT_>, a, bnp, relu_from_x, (int)x->c, dsums, count, dgamma,  \
                       dbeta, coef, (unsigned int*)gbar);                                                          \
  } while (0)
  if (g.threads == 1024) BASI_LAUNCH_RESIDENT(1024);
  else if (g.threads == 512) BASI_LAUNCH_RESIDENT(512);
  else BASI_LAUNCH_RESIDENT(256);
#undef BASI_LAUNCH_RESIDENT
  if (a.debug) {
    unsigned long long h[8];
    cudaStreamSynchronize(st);
    cudaMemcpyFromSymbol(h, g_resident_dbg, sizeof(h));
    fprintf(stderr, "resident R=%lld C=%d: rows %.2f us | block reduce + atomics %.2f | barrier %.2f | coef %.2f | dx %.2f | total %.2f\n",
            (long long)a.R, (int)x->c, (h[5] - h[0]) * 1e-3, (h[1] - h[5]) * 1e-3, (h[2] - h[1]) * 1e-3,
            (h[3] - h[2]) * 1e-3, (h[4] - h[3]) * 1e-3, (h[4] - h[0]) * 1e-3);
  }
}

}  // namespace basi

using namespace basi;

#define DISPATCH_T(dtype, ...) \
  if ((dtype) == BASI_F32) {   \
    typedef float T;           \
    __VA_ARGS__                \
  } else {                     \
    typedef bf16 T;            \
    __VA_ARGS__                \
  }

extern "C" {

int basi_bn_stats(const basi_tensor* x, double* sums, const float* gamma, const float* beta, double count, float eps,
                  float* bnp, uint32_t* counter, void* stream) {
  BASI_CHECK_ARG(x && sums && counter && vec_ok(x), "bn_stats: tensor must have c, ld multiple of the vector width");
  BASI_CHECK_ARG(!gamma || (beta && bnp && count > 0), "bn_stats: fused finalize needs gamma, beta, bnp, count");
  int64_t R = pixels(x);
  DISPATCH_T(x->dtype, {
    RowGeom g = row_geom(R, x->c, Vec<T>::N, 2 * UNR, 8, 2 * Vec<T>::N * sizeof(double));
    basi::launch(bn_stats_kernel<T>, g.grid, g.block, g.smem, (cudaStream_t)stream, (const T*)x->ptr, R, x->c, x->ld, sums, gamma,
                                                                          beta, count, eps, bnp, counter);
  })
  BASI_CHECK_LAUNCH("bn_stats");
  return BASI_OK;
}

int basi_bn_finalize(const double* sums, const float* gamma, const float* beta, double count, float eps, float* bnp,
                     int C, void* stream) {
  BASI_CHECK_ARG(sums && gamma && beta && bnp && C > 0 && count > 0, "bn_finalize: bad argument");
  basi::launch(bn_finalize_kernel, (C + 127) / 128, 128, 0, (cudaStream_t)stream, sums, gamma, beta, count, eps, bnp, C);
  BASI_CHECK_LAUNCH("bn_finalize");
  return BASI_OK;
}

int basi_bn_apply(const basi_tensor* x, const float* bnp, const basi_tensor* res, const float* res_bnp, int relu,
                  const basi_tensor* out, void* stream) {
  BASI_CHECK_ARG(x && bnp && out && vec_ok(x) && vec_ok(out) && same_shape(x, out) && x->dtype == out->dtype,
                 "bn_apply: bad x/out");
  BASI_CHECK_ARG(!res || (vec_ok(res) && same_shape(x, res) && res->dtype == x->dtype), "bn_apply: bad residual");
  BASI_CHECK_ARG(res || !res_bnp, "bn_apply: res_bnp without res");
  int64_t R = pixels(x);
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH_T(x->dtype, {
    if (launch_apply_stream<T>(x, bnp, res, res_bnp, relu, out, st)) {
      BASI_CHECK_LAUNCH("bn_apply(stream)");
      return BASI_OK;
    }
    RowGeom g = row_geom(R, x->c, Vec<T>::N, UNR, 16, 0);
    const T* rp = res ? (const T*)res->ptr : nullptr;
    const int ldr = res ? res->ld : 0;
    if (!res)
      basi::launch(bn_apply_kernel<T, false, false>, g.grid, g.block, 0, st, (const T*)x->ptr, x->ld, bnp, rp, ldr, nullptr, relu,
                                                                    (T*)out->ptr, out->ld, R, x->c);
    else if (!res_bnp)
      basi::launch(bn_apply_kernel<T, true, false>, g.grid, g.block, 0, st, (const T*)x->ptr, x->ld, bnp, rp, ldr, nullptr, relu,
                                                                   (T*)out->ptr, out->ld, R, x->c);
    else
      basi::launch(bn_apply_kernel<T, true, true>, g.grid, g.block, 0, st, (const T*)x->ptr, x->ld, bnp, rp, ldr, res_bnp, relu,
                                                                  (T*)out->ptr, out->ld, R, x->c);
  })
  BASI_CHECK_LAUNCH("bn_apply");
  return BASI_OK;
}

int basi_bn_bwd_reduce(const basi_tensor* dout, const basi_tensor* out, const basi_tensor* x, const float* bnp,
                       int relu_from_x, double* dsums, double count, float* dgamma, float* dbeta, float* coef,
                       uint32_t* counter, void* stream) {
  BASI_CHECK_ARG(dout && x && bnp && dsums && counter && vec_ok(dout) && vec_ok(x) && same_shape(dout, x) &&
                     dout->dtype == x->dtype,
                 "bn_bwd_reduce: bad dout/x");
  BASI_CHECK_ARG(!out || (vec_ok(out) && same_shape(out, x) && out->dtype == x->dtype), "bn_bwd_reduce: bad out");
  BASI_CHECK_ARG(!coef || (dgamma && dbeta && count > 0), "bn_bwd_reduce: fused finalize needs dgamma, dbeta, count");
  int64_t R = pixels(x);
  DISPATCH_T(x->dtype, {
    if (launch_bwd_reduce_stream<T>(dout, out, x, bnp, relu_from_x, dsums, count, dgamma, dbeta, coef, counter,
                                    (cudaStream_t)stream)) {
      BASI_CHECK_LAUNCH("bn_bwd_reduce(stream)");
      return BASI_OK;
    }
    RowGeom g = row_geom(R, x->c, Vec<T>::N, 4 * BUNR, 12, 2 * Vec<T>::N * sizeof(double));
    basi::launch(bn_bwd_reduce_kernel<T>, g.grid, g.block, g.smem, (cudaStream_t)stream, (const T*)dout->ptr, dout->ld, out ? (const T*)out->ptr : nullptr, out ? out->ld : 0, (const T*)x->ptr, x->ld,
        bnp, relu_from_x, R, x->c, dsums, count, dgamma, dbeta, coef, counter);
  })
  BASI_CHECK_LAUNCH("bn_bwd_reduce");
  return BASI_OK;
}

int basi_bn_bwd_finalize(const double* dsums, double count, float* dgamma, float* dbeta, float* coef, int C,
                         void* stream) {
  BASI_CHECK_ARG(dsums && dgamma && dbeta && coef && C > 0, "bn_bwd_finalize: bad argument");
  basi::launch(bn_bwd_finalize_kernel, (C + 127) / 128, 128, 0, (cudaStream_t)stream, dsums, count, dgamma, dbeta, coef, C);
  BASI_CHECK_LAUNCH("bn_bwd_finalize");
  return BASI_OK;
}

int basi_bn_bwd_apply(const basi_tensor* dout, const basi_tensor* out, const basi_tensor* x, const float* bnp,
                      const float* coef, int relu_from_x, const basi_tensor* dx, const basi_tensor* dres,
                      int dres_accumulate, void* stream) {
  BASI_CHECK_ARG(dout && x && dx && bnp && coef && vec_ok(dout) && vec_ok(x) && vec_ok(dx) && same_shape(dout, x) &&
                     same_shape(dx, x) && dout->dtype == x->dtype && dx->dtype == x->dtype,
                 "bn_bwd_apply: bad dout/x/dx");
  BASI_CHECK_ARG(!out || (vec_ok(out) && same_shape(out, x) && out->dtype == x->dtype), "bn_bwd_apply: bad out");
  BASI_CHECK_ARG(!dres || (vec_ok(dres) && same_shape(dres, x) && dres->dtype == x->dtype), "bn_bwd_apply: bad dres");
  int64_t R = pixels(x);
  DISPATCH_T(x->dtype, {
    if (launch_bwd_apply_stream<T>(dout, out, x, bnp, coef, relu_from_x, dx, dres, dres_accumulate,
                                   (cudaStream_t)stream)) {
      BASI_CHECK_LAUNCH("bn_bwd_apply(stream)");
      return BASI_OK;
    }
    RowGeom g = row_geom(R, x->c, Vec<T>::N, 2 * BUNR, 12, 0);
    basi::launch(bn_bwd_apply_kernel<T>, g.grid, g.block, 0, (cudaStream_t)stream, (const T*)dout->ptr, dout->ld, out ? (const T*)out->ptr : nullptr, out ? out->ld : 0, (const T*)x->ptr, x->ld,
        bnp, coef, relu_from_x, (T*)dx->ptr, dx->ld, dres ? (T*)dres->ptr : nullptr, dres ? dres->ld : 0,
        dres_accumulate, R, x->c);
  })
  BASI_CHECK_LAUNCH("bn_bwd_apply");
  return BASI_OK;
}

/* ---- packed ReLU mask (one bit per element) for the residual junctions: the backward pair then reads 1/16 of the
 * bytes of the stored output.  bf16 tensors with c % 128 == 0 that take the streamed kernels. */
int basi_bn_maskbits_supported(const basi_tensor* x) {
  if (!x || x->dtype != BASI_BF16 || !vec_ok(x) || x->c % 128 != 0 || stream_disabled()) return 0;
  return stream_geom(pixels(x), x->c, 2, 4, 0, true).ok ? 1 : 0;
}

int basi_bn_apply_bits(const basi_tensor* x, const float* bnp, const basi_tensor* res, const float* res_bnp, int relu,
                       const basi_tensor* out, unsigned char* maskbits, void* stream) {
  BASI_CHECK_ARG(x && bnp && out && maskbits && vec_ok(x) && vec_ok(out) && same_shape(x, out) &&
                     x->dtype == BASI_BF16 && out->dtype == BASI_BF16,
                 "bn_apply_bits: bad x/out (bf16 only)");
  BASI_CHECK_ARG(!res || (vec_ok(res) && same_shape(x, res) && res->dtype == x->dtype), "bn_apply_bits: bad residual");
  BASI_CHECK_ARG(res || !res_bnp, "bn_apply_bits: res_bnp without res");
  BASI_CHECK_ARG(basi_bn_maskbits_supported(x) == 1, "bn_apply_bits: tensor not supported (see basi_bn_maskbits_supported)");
  bool ok = launch_apply_stream<bf16>(x, bnp, res, res_bnp, relu, out, (cudaStream_t)stream, maskbits);
  BASI_CHECK_ARG(ok, "bn_apply_bits: streamed kernel unavailable");
  BASI_CHECK_LAUNCH("bn_apply_bits");
  return BASI_OK;
}

int basi_bn_bwd_reduce_bits(const basi_tensor* dout, const unsigned char* maskbits, const basi_tensor* x,
                            const float* bnp, double* dsums, double count, float* dgamma, float* dbeta, float* coef,
                            uint32_t* counter, void* stream) {
  BASI_CHECK_ARG(dout && x && maskbits && bnp && dsums && counter && vec_ok(dout) && vec_ok(x) && same_shape(dout, x) &&
                     dout->dtype == BASI_BF16 && x->dtype == BASI_BF16,
                 "bn_bwd_reduce_bits: bad dout/x (bf16 only)");
  BASI_CHECK_ARG(!coef || (dgamma && dbeta && count > 0), "bn_bwd_reduce_bits: fused finalize needs dgamma, dbeta, count");
  BASI_CHECK_ARG(basi_bn_maskbits_supported(x) == 1, "bn_bwd_reduce_bits: tensor not supported");
  bool ok = launch_bwd_reduce_stream<bf16>(dout, nullptr, x, bnp, 0, dsums, count, dgamma, dbeta, coef, counter,
                                           (cudaStream_t)stream, maskbits);
  BASI_CHECK_ARG(ok, "bn_bwd_reduce_bits: streamed kernel unavailable");
  BASI_CHECK_LAUNCH("bn_bwd_reduce_bits");
  return BASI_OK;
}

int basi_bn_bwd_apply_bits(const basi_tensor* dout, const unsigned char* maskbits, const basi_tensor* x,
                           const float* bnp, const float* coef, const basi_tensor* dx, const basi_tensor* dres,
                           int dres_accumulate, void* stream) {
  BASI_CHECK_ARG(dout && x && dx && maskbits && bnp && coef && vec_ok(dout) && vec_ok(x) && vec_ok(dx) &&
                     same_shape(dout, x) && same_shape(dx, x) && dout->dtype == BASI_BF16 && x->dtype == BASI_BF16 &&
                     dx->dtype == BASI_BF16,
                 "bn_bwd_apply_bits: bad dout/x/dx (bf16 only)");
  BASI_CHECK_ARG(!dres || (vec_ok(dres) && same_shape(dres, x) && dres->dtype == x->dtype), "bn_bwd_apply_bits: bad dres");
  BASI_CHECK_ARG(basi_bn_maskbits_supported(x) == 1, "bn_bwd_apply_bits: tensor not supported");
  bool ok = launch_bwd_apply_stream<bf16>(dout, nullptr, x, bnp, coef, 0, dx, dres, dres_accumulate,
                                          (cudaStream_t)stream, maskbits);
  BASI_CHECK_ARG(ok, "bn_bwd_apply_bits: streamed kernel unavailable");
  BASI_CHECK_LAUNCH("bn_bwd_apply_bits");
  return BASI_OK;
}

int basi_bn_bwd_coop_supported(const basi_tensor* x, int has_out, int has_bits, int dres_accumulate) {
  if (!x || !vec_ok(x) || (has_out && has_bits)) return 0;
  if (has_bits && basi_bn_maskbits_supported(x) != 1) return 0;
  basi_tensor dres = *x;
  const basi_tensor* outp = has_out ? x : nullptr;
  const unsigned char* bitsp = has_bits ? (const unsigned char*)x->ptr : nullptr;
  bool ok = false;
  DISPATCH_T(x->dtype, {
    ok = launch_bwd_coop<T>(true, x, outp, bitsp, x, nullptr, 0, nullptr, 1.0, nullptr, nullptr, nullptr, nullptr, x,
                            dres_accumulate ? &dres : nullptr, dres_accumulate, nullptr);
  })
  return ok ? 1 : 0;
}

int basi_bn_bwd_coop(const basi_tensor* dout, const basi_tensor* out, const unsigned char* maskbits,
                     const basi_tensor* x, const float* bnp, int relu_from_x, double* dsums, double count,
                     float* dgamma, float* dbeta, float* coef, uint32_t* barrier, const basi_tensor* dx,
                     const basi_tensor* dres, int dres_accumulate, void* stream) {
  BASI_CHECK_ARG(dout && x && dx && bnp && dsums && dgamma && dbeta && barrier && count > 0 && vec_ok(dout) &&
                     vec_ok(x) && vec_ok(dx) && same_shape(dout, x) && same_shape(dx, x) && dout->dtype == x->dtype &&
                     dx->dtype == x->dtype && !(out && maskbits),
                 "bn_bwd_coop: bad dout/x/dx");
  BASI_CHECK_ARG(!out || (vec_ok(out) && same_shape(out, x) && out->dtype == x->dtype), "bn_bwd_coop: bad out");
  BASI_CHECK_ARG(!dres || (vec_ok(dres) && same_shape(dres, x) && dres->dtype == x->dtype), "bn_bwd_coop: bad dres");
  BASI_CHECK_ARG(!maskbits || basi_bn_maskbits_supported(x) == 1, "bn_bwd_coop: mask bits not supported for this tensor");
  bool ok = false;
  DISPATCH_T(x->dtype, {
    ok = launch_bwd_coop<T>(false, dout, out, maskbits, x, bnp, relu_from_x, dsums, count, dgamma, dbeta, coef, barrier,
                            dx, dres, dres_accumulate, (cudaStream_t)stream);
  })
  BASI_CHECK_ARG(ok, "bn_bwd_coop: tensor does not take the cooperative path (see basi_bn_bwd_coop_supported)");
  BASI_CHECK_LAUNCH("bn_bwd_coop");
  return BASI_OK;
}

int basi_bn_bwd_fused_supported(const basi_tensor* x) {
  if (!x || !vec_ok(x) || !dense_rows(x)) return 0;
  ResidentGeom g = resident_geom(pixels(x), x->c, x->dtype == BASI_F32 ? 4 : 2);
  if (!g.ok) return 0;
  // one CTA per SM must fit (the grid barrier needs the whole grid co-resident)
  int nb = 0;
  DISPATCH_T(x->dtype, {
    if (g.threads == 1024) {
      allow_smem(bn_bwd_resident_kernel<T, 1024>, g.smem);
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, bn_bwd_resident_kernel<T, 1024>, g.threads, g.smem);
    } else if (g.threads == 512) {
      allow_smem(bn_bwd_resident_kernel<T, 512>, g.smem);
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, bn_bwd_resident_kernel<T, 512>, g.threads, g.smem);
    } else {
      allow_smem(bn_bwd_resident_kernel<T, 256>, g.smem);
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, bn_bwd_resident_kernel<T, 256>, g.threads, g.smem);
    }
  })
  if (cudaGetLastError() != cudaSuccess) return 0;
  return (nb >= 1 && g.grid <= nb * sm_count()) ? 1 : 0;
}

int basi_bn_bwd_fused(const basi_tensor* dout, const basi_tensor* x, const float* bnp, int relu_from_x, double* dsums,
                      double count, float* dgamma, float* dbeta, float* coef, uint32_t* barrier,
                      const basi_tensor* dx, void* stream) {
  BASI_CHECK_ARG(dout && x && dx && bnp && dsums && dgamma && dbeta && barrier && count > 0 && vec_ok(dout) &&
                     vec_ok(x) && vec_ok(dx) && same_shape(dout, x) && same_shape(dx, x) && dout->dtype == x->dtype &&
                     dx->dtype == x->dtype && dense_rows(dout) && dense_rows(x) && dense_rows(dx),
                 "bn_bwd_fused: bad dout/x/dx (dense rows required)");
  ResidentGeom g = resident_geom(pixels(x), x->c, x->dtype == BASI_F32 ? 4 : 2);
  BASI_CHECK_ARG(g.ok, "bn_bwd_fused: tensor does not fit the resident kernel (see basi_bn_bwd_fused_supported)");
  DISPATCH_T(x->dtype, {
    launch_bwd_resident<T>(g, dout, x, bnp, relu_from_x, dsums, count, dgamma, dbeta, coef, barrier, dx,
                           (cudaStream_t)stream);
  })
  BASI_CHECK_LAUNCH("bn_bwd_fused");
  return BASI_OK;
}

}  // extern "C"
