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
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) finalize_channel(sums, gamma, beta, count, eps, bnp, C, c);
}

// ---- apply: out = act((x-mean)*scale+beta [+ res | + (res-rmean)*rscale+rbeta])
template <typename T, bool HAS_RES, bool RES_BN>
__global__ void __launch_bounds__(256) bn_apply_kernel(const T* __restrict__ x, int ldx, const float* __restrict__ bnp,
                                                       const T* __restrict__ res, int ldr,
                                                       const float* __restrict__ rbnp, int relu, T* __restrict__ out,
                                                       int ldo, int64_t R, int C) {
  constexpr int VN = Vec<T>::N;
  const int cg = blockIdx.y * blockDim.x + threadIdx.x;
  if (cg * VN >= C) return;
  const int c0 = cg * VN;
  float mean[VN], scale[VN], beta[VN];
  float rmean[VN], rscale[VN], rbeta[VN];
#pragma unroll
  for (int i = 0; i < VN; ++i) {
    mean[i] = bnp[c0 + i];
    scale[i] = bnp[2 * C + c0 + i];
    beta[i] = bnp[3 * C + c0 + i];
    if (RES_BN) {
      rmean[i] = rbnp[c0 + i];
      rscale[i] = rbnp[2 * C + c0 + i];
      rbeta[i] = rbnp[3 * C + c0 + i];
    }
  }
  const int64_t rstep = (int64_t)gridDim.x * blockDim.y;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.y + threadIdx.y; r < R; r += rstep * UNR) {
    Vec<T> v[UNR], rv[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int64_t rr = r + u * rstep;
      if (rr < R) {
        v[u] = Vec<T>::load(x + rr * ldx + c0);
        if (HAS_RES) rv[u] = Vec<T>::load(res + rr * ldr + c0);
      }
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int64_t rr = r + u * rstep;
      if (rr >= R) break;
#pragma unroll
      for (int i = 0; i < VN; ++i) {
        float o = fmaf(v[u].v[i] - mean[i], scale[i], beta[i]);
        if (HAS_RES) {
          float t = rv[u].v[i];
          if (RES_BN) t = fmaf(t - rmean[i], rscale[i], rbeta[i]);
          o += t;
        }
        if (relu) o = fmaxf(o, 0.f);
        v[u].v[i] = o;
      }
      v[u].store(out + rr * ldo + c0);
    }
  }
}

// ---- backward reduce: dsums += (sum dy, sum dy*xhat); dy = dout * mask.
// mask: out > 0 when out != NULL; else, when relu_from_x, (x-mean)*scale+beta > 0 (plain BN+ReLU: no need to read out)
template <typename T>
__global__ void __launch_bounds__(256)
bn_bwd_reduce_kernel(const T* __restrict__ dout, int ldd, const T* __restrict__ out, int ldo, const T* __restrict__ x,
                     int ldx, const float* __restrict__ bnp, int relu_from_x, int64_t R, int C,
                     double* __restrict__ dsums, double count, float* __restrict__ dgamma, float* __restrict__ dbeta,
                     float* __restrict__ coef, unsigned int* counter) {
  constexpr int VN = Vec<T>::N;
  const int cg = blockIdx.y * blockDim.x + threadIdx.x;
  const bool valid = cg * VN < C;
  const int c0 = cg * VN;
  double s[VN], q[VN];
#pragma unroll
  for (int i = 0; i < VN; ++i) s[i] = q[i] = 0.0;
  if (valid) {
    float mean[VN], istd[VN], scale[VN], beta[VN];
#pragma unroll
    for (int i = 0; i < VN; ++i) {
      mean[i] = bnp[c0 + i];
      istd[i] = bnp[C + c0 + i];
      scale[i] = bnp[2 * C + c0 + i];
      beta[i] = bnp[3 * C + c0 + i];
    }
    const int64_t rstep = (int64_t)gridDim.x * blockDim.y;
    float fs[VN], fq[VN];
#pragma unroll
    for (int i = 0; i < VN; ++i) fs[i] = fq[i] = 0.f;
    int flush = 0;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.y + threadIdx.y; r < R; r += rstep * UNR) {
      Vec<T> d[UNR], xv[UNR], o[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int64_t rr = r + u * rstep;
        if (rr < R) {
          d[u] = Vec<T>::load(dout + rr * ldd + c0);
          xv[u] = Vec<T>::load(x + rr * ldx + c0);
          if (out) o[u] = Vec<T>::load(out + rr * ldo + c0);
        } else {
          d[u] = Vec<T>::zero();
          xv[u] = Vec<T>::zero();
          o[u] = Vec<T>::zero();
        }
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u)
#pragma unroll
        for (int i = 0; i < VN; ++i) {
          const float xc = xv[u].v[i] - mean[i];
          float dy = d[u].v[i];
          if (out) dy = o[u].v[i] > 0.f ? dy : 0.f;
          else if (relu_from_x) dy = fmaf(xc, scale[i], beta[i]) > 0.f ? dy : 0.f;
          fs[i] += dy;
          fq[i] = fmaf(dy, xc * istd[i], fq[i]);
        }
      if (++flush == 4) {   // fp32 partials over 16 rows, then double
        flush = 0;
#pragma unroll
        for (int i = 0; i < VN; ++i) {
          s[i] += (double)fs[i];
          q[i] += (double)fq[i];
          fs[i] = fq[i] = 0.f;
        }
      }
    }
#pragma unroll
    for (int i = 0; i < VN; ++i) {
      s[i] += (double)fs[i];
      q[i] += (double)fq[i];
    }
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

// ---- backward apply: dx = scale*(dy - c1 - xhat*c2); dres (+)= dy
template <typename T>
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const T* __restrict__ dout, int ldd, const T* __restrict__ out, int ldo, const T* __restrict__ x,
                    int ldx, const float* __restrict__ bnp, const float* __restrict__ coef, int relu_from_x,
                    T* __restrict__ dx, int lddx, T* __restrict__ dres, int lddr, int dres_acc, int64_t R, int C) {
  constexpr int VN = Vec<T>::N;
  const int cg = blockIdx.y * blockDim.x + threadIdx.x;
  if (cg * VN >= C) return;
  const int c0 = cg * VN;
  float mean[VN], istd[VN], scale[VN], beta[VN], k1[VN], k2[VN];
#pragma unroll
  for (int i = 0; i < VN; ++i) {
    mean[i] = bnp[c0 + i];
    istd[i] = bnp[C + c0 + i];
    scale[i] = bnp[2 * C + c0 + i];
    beta[i] = bnp[3 * C + c0 + i];
    k1[i] = coef[c0 + i];
    k2[i] = coef[C + c0 + i];
  }
  const int64_t rstep = (int64_t)gridDim.x * blockDim.y;
  // rows are visited from the end: the reduce pass that ran just before touched the tail last, so it is the part
  // of dout / x / out most likely still in L2
  const int64_t first = (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
  for (int64_t rb = first; rb < R; rb += rstep * UNR) {
    Vec<T> d[UNR], xv[UNR], o[UNR], dr[UNR];
    int64_t rows[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int64_t rr = rb + u * rstep;
      rows[u] = rr < R ? (R - 1 - rr) : -1;
      if (rows[u] >= 0) {
        d[u] = Vec<T>::load(dout + rows[u] * ldd + c0);
        xv[u] = Vec<T>::load(x + rows[u] * ldx + c0);
        if (out) o[u] = Vec<T>::load(out + rows[u] * ldo + c0);
        if (dres && dres_acc) dr[u] = Vec<T>::load(dres + rows[u] * lddr + c0);
      }
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      if (rows[u] < 0) continue;
#pragma unroll
      for (int i = 0; i < VN; ++i) {
        const float xc = xv[u].v[i] - mean[i];
        float dy = d[u].v[i];
        if (out) dy = o[u].v[i] > 0.f ? dy : 0.f;
        else if (relu_from_x) dy = fmaf(xc, scale[i], beta[i]) > 0.f ? dy : 0.f;
        if (dres) dr[u].v[i] = dres_acc ? dr[u].v[i] + dy : dy;
        d[u].v[i] = scale[i] * (dy - k1[i] - xc * istd[i] * k2[i]);
      }
      if (dres) dr[u].store(dres + rows[u] * lddr + c0);
      d[u].store(dx + rows[u] * lddx + c0);
    }
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
    bn_stats_kernel<T><<<g.grid, g.block, g.smem, (cudaStream_t)stream>>>((const T*)x->ptr, R, x->c, x->ld, sums, gamma,
                                                                          beta, count, eps, bnp, counter);
  })
  BASI_CHECK_LAUNCH("bn_stats");
  return BASI_OK;
}

int basi_bn_finalize(const double* sums, const float* gamma, const float* beta, double count, float eps, float* bnp,
                     int C, void* stream) {
  BASI_CHECK_ARG(sums && gamma && beta && bnp && C > 0 && count > 0, "bn_finalize: bad argument");
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(sums, gamma, beta, count, eps, bnp, C);
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
    RowGeom g = row_geom(R, x->c, Vec<T>::N, UNR, 16, 0);
    const T* rp = res ? (const T*)res->ptr : nullptr;
    const int ldr = res ? res->ld : 0;
    if (!res)
      bn_apply_kernel<T, false, false><<<g.grid, g.block, 0, st>>>((const T*)x->ptr, x->ld, bnp, rp, ldr, nullptr, relu,
                                                                    (T*)out->ptr, out->ld, R, x->c);
    else if (!res_bnp)
      bn_apply_kernel<T, true, false><<<g.grid, g.block, 0, st>>>((const T*)x->ptr, x->ld, bnp, rp, ldr, nullptr, relu,
                                                                   (T*)out->ptr, out->ld, R, x->c);
    else
      bn_apply_kernel<T, true, true><<<g.grid, g.block, 0, st>>>((const T*)x->ptr, x->ld, bnp, rp, ldr, res_bnp, relu,
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
    RowGeom g = row_geom(R, x->c, Vec<T>::N, 2 * UNR, 8, 2 * Vec<T>::N * sizeof(double));
    bn_bwd_reduce_kernel<T><<<g.grid, g.block, g.smem, (cudaStream_t)stream>>>(
        (const T*)dout->ptr, dout->ld, out ? (const T*)out->ptr : nullptr, out ? out->ld : 0, (const T*)x->ptr, x->ld,
        bnp, relu_from_x, R, x->c, dsums, count, dgamma, dbeta, coef, counter);
  })
  BASI_CHECK_LAUNCH("bn_bwd_reduce");
  return BASI_OK;
}

int basi_bn_bwd_finalize(const double* dsums, double count, float* dgamma, float* dbeta, float* coef, int C,
                         void* stream) {
  BASI_CHECK_ARG(dsums && dgamma && dbeta && coef && C > 0, "bn_bwd_finalize: bad argument");
  bn_bwd_finalize_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(dsums, count, dgamma, dbeta, coef, C);
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
    RowGeom g = row_geom(R, x->c, Vec<T>::N, UNR, 16, 0);
    bn_bwd_apply_kernel<T><<<g.grid, g.block, 0, (cudaStream_t)stream>>>(
        (const T*)dout->ptr, dout->ld, out ? (const T*)out->ptr : nullptr, out ? out->ld : 0, (const T*)x->ptr, x->ld,
        bnp, coef, relu_from_x, (T*)dx->ptr, dx->ld, dres ? (T*)dres->ptr : nullptr, dres ? dres->ld : 0,
        dres_accumulate, R, x->c);
  })
  BASI_CHECK_LAUNCH("bn_bwd_apply");
  return BASI_OK;
}

}  // extern "C"
