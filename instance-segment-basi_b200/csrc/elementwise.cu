// HBM-bound kernels of the BAIS PSPNet hot path: batch-norm (batch statistics) forward /
// backward with fused ReLU and residual add, pooling, align-corners bilinear and its adjoint,
// attention gating, click-map rasterisation, fused losses, SGD and prediction helpers.
// All tensors NHWC; 128-bit vector access over channels (4 x f32 / 8 x bf16).
#include <stdarg.h>
#include <stdlib.h>

#include <string.h>

#include "common.cuh"

namespace basi {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* exp_env(const char* name) {
  const char* v = getenv(name);
  if (!v) return nullptr;
  const char* on = getenv("BASI_EXPERIMENTS");
  if (on && atoi(on) == 1) return v;
  static thread_local char warned[1024] = "";
  if (!strstr(warned, name) && strlen(warned) + strlen(name) + 2 < sizeof(warned)) {
    strcat(warned, name);
    strcat(warned, ",");
    fprintf(stderr, "basi_b200: %s is set but ignored (experiment switches need BASI_EXPERIMENTS=1)\n", name);
  }
  return nullptr;
}
bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    // Programmatic dependent launch: on unless BASI_PDL=0.  Every kernel launched through basi::launch starts with
    // griddepcontrol.launch_dependents + griddepcontrol.wait, so the next kernel's CTAs become resident during this
    // kernel's tail wherever the SM still has room.  While every kernel used the whole shared memory this measured
    // nothing (12.31 vs 12.29 ms); with the register-staged BN kernels and the other small-footprint kernels between
    // the tcgen05 launches it is worth 0.37 ms of a 9.2 ms step.
    const char* e = exp_env("BASI_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}
// SM partition (basi_set_sm_budget): while a budget is set, every grid-size rule of the library sees that many SMs, so
// the main chain of the backward pass leaves the rest of the machine to the weight gradients on the side stream.
static int g_sm_budget = 0, g_wgrad_ctas = 0;
static int hw_sm_count() {
  static int cached = 0;
  if (cached > 0) return cached;
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
    cudaGetLastError();
    return 148;
  }
  cached = n;
  return n;
}
int sm_count() {
  const int hw = hw_sm_count();
  return (g_sm_budget > 0 && g_sm_budget < hw) ? g_sm_budget : hw;
}
int wgrad_cta_target() { return g_wgrad_ctas; }

// ------------------------------------------------------------------------------------------
// Pooling
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void maxpool3s2_fwd_kernel(const T* __restrict__ x, int ldx, int H, int W, int C, T* __restrict__ y,
                                      int ldy, int OH, int OW, int pad_t, int pad_l, uint8_t* __restrict__ amax,
                                      int64_t total) {
  pdl_prologue();
  constexpr int VN = Vec<T>::N;
  const int cgs = C / VN;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    // 32-bit index arithmetic (the host checks total < 2^31): 64-bit divisions cost tens of instructions each
    const unsigned iu = (unsigned)i;
    int cg = (int)(iu % (unsigned)cgs);
    const unsigned p = iu / (unsigned)cgs;
    int ow = (int)(p % (unsigned)OW);
    const unsigned t = p / (unsigned)OW;
    int oh = (int)(t % (unsigned)OH);
    int n = (int)(t / (unsigned)OH);
    float best[VN];
    uint8_t bi[VN];
#pragma unroll
    for (int j = 0; j < VN; ++j) {
      best[j] = -INFINITY;
      bi[j] = 0;
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      int ih = oh * 2 - pad_t + r;
      if (ih < 0 || ih >= H) continue;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        int iw = ow * 2 - pad_l + s;
        if (iw < 0 || iw >= W) continue;
        Vec<T> v = Vec<T>::load(x + (((int64_t)n * H + ih) * W + iw) * ldx + cg * VN);
#pragma unroll
        for (int j = 0; j < VN; ++j)
          if (v.v[j] > best[j]) {
            best[j] = v.v[j];
            bi[j] = (uint8_t)(r * 3 + s);
          }
      }
    }
    Vec<T> o;
#pragma unroll
    for (int j = 0; j < VN; ++j) o.v[j] = best[j];
    o.store(y + p * ldy + cg * VN);
    uint32_t ab[2] = {0u, 0u};
#pragma unroll
    for (int j = 0; j < VN; ++j) ab[j >> 2] |= (uint32_t)bi[j] << (8 * (j & 3));
    if (VN == 8) *reinterpret_cast<uint2*>(amax + p * C + cg * VN) = make_uint2(ab[0], ab[1]);
    else *reinterpret_cast<uint32_t*>(amax + p * C + cg * VN) = ab[0];
  }
}

template <typename T>
__global__ void maxpool3s2_bwd_kernel(const T* __restrict__ dy, int ldy, int OH, int OW, const uint8_t* __restrict__ amax,
                                      T* __restrict__ dx, int ldx, int H, int W, int C, int pad_t, int pad_l, int acc,
                                      int64_t total) {
  pdl_prologue();
  constexpr int VN = Vec<T>::N;
  const int cgs = C / VN;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const unsigned iu = (unsigned)i;      // 32-bit index arithmetic, see the forward kernel
    int cg = (int)(iu % (unsigned)cgs);
    const int64_t p = iu / (unsigned)cgs;
    int iw = (int)((unsigned)p % (unsigned)W);
    const unsigned t = (unsigned)p / (unsigned)W;
    int ih = (int)(t % (unsigned)H);
    int n = (int)(t / (unsigned)H);
    Vec<T> g = Vec<T>::zero();
    if (acc) g = Vec<T>::load(dx + p * ldx + cg * VN);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      int th = ih + pad_t - r;
      if (th < 0 || (th & 1)) continue;
      int oh = th >> 1;
      if (oh >= OH) continue;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        int tw = iw + pad_l - s;
        if (tw < 0 || (tw & 1)) continue;
        int ow = tw >> 1;
        if (ow >= OW) continue;
        int64_t op = ((int64_t)n * OH + oh) * OW + ow;
        Vec<T> d = Vec<T>::load(dy + op * ldy + cg * VN);
        // the VN argmax bytes of this fragment in one load (C % VN == 0 keeps them VN-byte aligned)
        uint32_t ab[2];
        if (VN == 8) {
          const uint2 t2 = *reinterpret_cast<const uint2*>(amax + op * C + cg * VN);
          ab[0] = t2.x; ab[1] = t2.y;
        } else {
          ab[0] = *reinterpret_cast<const uint32_t*>(amax + op * C + cg * VN);
          ab[1] = 0;
        }
#pragma unroll
        for (int j = 0; j < VN; ++j)
          if (((ab[j >> 2] >> (8 * (j & 3))) & 0xffu) == (uint32_t)(r * 3 + s)) g.v[j] += d.v[j];
      }
    }
    g.store(dx + p * ldx + cg * VN);
  }
}

// one block per (output pixel, chunk of blockDim.x channel groups); ty strides over the window pixels
template <typename T>
__global__ void avgpool_fwd_kernel(const T* __restrict__ x, int ldx, int H, int W, int C, int k, T* __restrict__ y,
                                   int ldy, int OH, int OW) {
  pdl_prologue();
  constexpr int VN = Vec<T>::N;
  extern __shared__ float sacc[];  // [blockDim.y][blockDim.x*VN]
  const int ow = blockIdx.x % OW;
  const int oh = (blockIdx.x / OW) % OH;
  const int n = blockIdx.x / (OW * OH);
  const int cgs = C / VN;
  const float inv = 1.0f / (float)(k * k);
  const int cg = blockIdx.y * blockDim.x + threadIdx.x;
  float a[VN];
#pragma unroll
  for (int j = 0; j < VN; ++j) a[j] = 0.f;
  if (cg < cgs) {
    const T* base = x + (((int64_t)n * H + oh * k) * W + ow * k) * ldx + cg * VN;
    for (int p = threadIdx.y; p < k * k; p += blockDim.y) {
      const int dy = p / k, dx = p - dy * k;
      Vec<T> v = Vec<T>::load(base + ((int64_t)dy * W + dx) * ldx);
#pragma unroll
      for (int j = 0; j < VN; ++j) a[j] += v.v[j];
    }
  }
  float* mine = sacc + ((size_t)threadIdx.y * blockDim.x + threadIdx.x) * VN;
#pragma unroll
  for (int j = 0; j < VN; ++j) mine[j] = a[j];
  __syncthreads();
  // tree over ty
  for (int s2 = blockDim.y >> 1; s2 >= 1; s2 >>= 1) {
    if ((int)threadIdx.y < s2) {
      const float* o = sacc + ((size_t)(threadIdx.y + s2) * blockDim.x + threadIdx.x) * VN;
#pragma unroll
      for (int j = 0; j < VN; ++j) mine[j] += o[j];
    }
    __syncthreads();
  }
  if (threadIdx.y == 0 && cg < cgs) {
    Vec<T> r;
#pragma unroll
    for (int j = 0; j < VN; ++j) r.v[j] = mine[j] * inv;
    r.store(y + (((int64_t)n * OH + oh) * OW + ow) * ldy + cg * VN);
  }
}

template <typename T>
__global__ void avgpool_bwd_kernel(const T* __restrict__ dy, int ldy, int OH, int OW, int k, T* __restrict__ dx,
                                   int ldx, int H, int W, int C, int acc, int64_t total) {
  pdl_prologue();
  constexpr int VN = Vec<T>::N;
  const int cgs = C / VN;
  const float inv = 1.0f / (float)(k * k);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int cg = (int)(i % cgs);
    int64_t p = i / cgs;
    int iw = (int)(p % W);
    int64_t t = p / W;
    int ih = (int)(t % H);
    int n = (int)(t / H);
    int oh = ih / k, ow = iw / k;
    Vec<T> g = Vec<T>::zero();
    if (acc) g = Vec<T>::load(dx + p * ldx + cg * VN);
    if (oh < OH && ow < OW) {
      Vec<T> d = Vec<T>::load(dy + (((int64_t)n * OH + oh) * OW + ow) * ldy + cg * VN);
#pragma unroll
      for (int j = 0; j < VN; ++j) g.v[j] += d.v[j] * inv;
    } else if (acc) {
      continue;
    }
    g.store(dx + p * ldx + cg * VN);
  }
}


// ------------------------------------------------------------------------------------------
// Pyramid pooling (BAISPSPNet.py:683-710): the four average pools of conv5_3 read the SAME 52 MB tensor with windows
// 40/20/13/6.  One pass: block = (image, band of rows, 64 channels); thread = (8-channel group, column) keeps the
// running window sums of all pools in registers, flushes them to per-cell shared-memory accumulators when a row
// window ends, the block adds its cells to an fp32 scratch; a tiny second kernel scales, converts and re-zeroes the
// scratch.  The adjoint is one read-modify-write pass over dx instead of four.
// ------------------------------------------------------------------------------------------
constexpr int MP_MAX = 4;
struct MultiPool {
  int np, cells;
  int k[MP_MAX], OH[MP_MAX], OW[MP_MAX], cell0[MP_MAX];
  void* y[MP_MAX];   // pooled tensors (fwd: outputs, bwd: their gradients)
  int ldy[MP_MAX];
};

// Forward, two passes without atomics (bit-reproducible: an fp32 atomic version was 1e-7-noisy in the pooled sums,
// which a 16-bit rounding flip plus the 4-sample batch norm of the 1x1 branch amplified to 3e-3 on the logits):
//   rows pass   thread = (image, row, 8-channel group) walks along its row keeping the running window sum of every
//               pool and writes it when a column window ends: scratch[n][h][row cell][c], row cells = sum of OW_p
//   cells pass  thread = (image, cell, 8-channel group) adds the k rows of its window in order, scales, converts
template <typename T>
__global__ void __launch_bounds__(256) avgpool_multi_rows_kernel(const T* __restrict__ x, int ldx, int H, int W, int C,
                                                                 const MultiPool mp, int rcells,
                                                                 float* __restrict__ scratch, int64_t total) {
  pdl_prologue();
  constexpr int VN = Vec<T>::N;
  const int cgs = C / VN;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cg = (int)(i % cgs);
    const int64_t row = i / cgs;                 // n * H + h
    const T* xr = x + row * W * ldx + cg * VN;
    float* sr = scratch + (row * rcells) * C + cg * VN;
    float acc[MP_MAX][VN];
    int ow[MP_MAX], nextb[MP_MAX], rc0[MP_MAX];
    int off = 0;
#pragma unroll
    for (int p = 0; p < MP_MAX; ++p) {
      ow[p] = 0;
      nextb[p] = p < mp.np ? mp.k[p] : (1 << 30);
      rc0[p] = off;
      off += p < mp.np ? mp.OW[p] : 0;
#pragma unroll
      for (int j = 0; j < VN; ++j) acc[p][j] = 0.f;
    }
#pragma unroll 8
    for (int w = 0; w < W; ++w) {
      const Vec<T> v = Vec<T>::load(xr + (int64_t)w * ldx);
#pragma unroll
      for (int p = 0; p < MP_MAX; ++p) {
        if (p >= mp.np) continue;
#pragma unroll
        for (int j = 0; j < VN; ++j) acc[p][j] += v.v[j];
        if (w + 1 == nextb[p]) {                 // the column window of pool p ends with this column
          if (ow[p] < mp.OW[p]) {
            float* d = sr + (size_t)(rc0[p] + ow[p]) * C;
#pragma unroll
            for (int j = 0; j < VN; ++j) d[j] = acc[p][j];
          }
          ++ow[p];
          nextb[p] += mp.k[p];
#pragma unroll
          for (int j = 0; j < VN; ++j) acc[p][j] = 0.f;
        }
      }
    }
  }
}

template <typename T>
__global__ void avgpool_multi_cells_kernel(const MultiPool mp, int H, int C, int rcells,
                                           const float* __restrict__ scratch, int64_t total) {
  pdl_prologue();
  constexpr int VN = Vec<T>::N;
  const int cgs = C / VN;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cg = (int)(i % cgs);
    const int64_t t = i / cgs;
    const int cell = (int)(t % mp.cells), n = (int)(t / mp.cells);
    int p = 0, rc0 = 0;
#pragma unroll
    for (int q = 1; q < MP_MAX; ++q)
      if (q < mp.np && cell >= mp.cell0[q]) p = q;
#pragma unroll
    for (int q = 0; q < MP_MAX; ++q)
      if (q < p) rc0 += mp.OW[q];
    const int local = cell - mp.cell0[p];
    const int oh = local / mp.OW[p], ow = local % mp.OW[p];
    const float inv = 1.0f / (float)(mp.k[p] * mp.k[p]);
    float sum[VN];
#pragma unroll
    for (int j = 0; j < VN; ++j) sum[j] = 0.f;
#pragma unroll 4
    for (int h = oh * mp.k[p]; h < (oh + 1) * mp.k[p]; ++h) {
      const float* sp = scratch + (((size_t)n * H + h) * rcells + rc0 + ow) * C + cg * VN;
#pragma unroll
      for (int j = 0; j < VN; ++j) sum[j] += sp[j];
    }
    Vec<T> r;
#pragma unroll
    for (int j = 0; j < VN; ++j) r.v[j] = sum[j] * inv;
    r.store((T*)mp.y[p] + ((size_t)n * mp.OH[p] * mp.OW[p] + local) * mp.ldy[p] + cg * VN);
  }
}

// block = one pixel row (n, ih); thread = one 8-channel group, walks the columns: the row cell of every pool is
// computed once per block and the column cell advances with a counter (no per-element integer division)
template <typename T>
__global__ void __launch_bounds__(256) avgpool_multi_bwd_kernel(const MultiPool mp, T* __restrict__ dx, int ldx, int H,
                                                                int W, int C, int acc) {
  pdl_prologue();
  constexpr int VN = Vec<T>::N;
  const int cgs = C / VN;
  const int n = blockIdx.x / H, ih = blockIdx.x % H;
  const T* yrow[MP_MAX];
  float inv[MP_MAX];
  bool rowok[MP_MAX];
#pragma unroll
  for (int p = 0; p < MP_MAX; ++p) {
    rowok[p] = false;
    inv[p] = 0.f;
    yrow[p] = nullptr;
    if (p < mp.np) {
      const int oh = ih / mp.k[p];
      rowok[p] = oh < mp.OH[p];
      inv[p] = 1.0f / (float)(mp.k[p] * mp.k[p]);
      yrow[p] = (const T*)mp.y[p] + ((int64_t)n * mp.OH[p] + (rowok[p] ? oh : 0)) * mp.OW[p] * mp.ldy[p];
    }
  }
  T* xrow = dx + ((int64_t)n * H + ih) * W * ldx;
  for (int cg = threadIdx.x; cg < cgs; cg += blockDim.x) {
    int ow[MP_MAX], nextb[MP_MAX];
    Vec<T> cur[MP_MAX];           // the pooled gradient of the current column cell, already scaled
#pragma unroll
    for (int p = 0; p < MP_MAX; ++p) {
      ow[p] = 0;
      nextb[p] = p < mp.np ? mp.k[p] : (1 << 30);
      cur[p] = Vec<T>::zero();
      if (p < mp.np && rowok[p]) {
        cur[p] = Vec<T>::load(yrow[p] + cg * VN);
#pragma unroll
        for (int j = 0; j < VN; ++j) cur[p].v[j] *= inv[p];
      }
    }
    for (int iw = 0; iw < W; ++iw) {
      bool any = false;
      Vec<T> g = Vec<T>::zero();
#pragma unroll
      for (int p = 0; p < MP_MAX; ++p) {
        if (p >= mp.np) continue;
        if (iw == nextb[p]) {
          ++ow[p];
          nextb[p] += mp.k[p];
          cur[p] = Vec<T>::zero();
          if (rowok[p] && ow[p] < mp.OW[p]) {
            cur[p] = Vec<T>::load(yrow[p] + (int64_t)ow[p] * mp.ldy[p] + cg * VN);
#pragma unroll
            for (int j = 0; j < VN; ++j) cur[p].v[j] *= inv[p];
          }
        }
        if (rowok[p] && ow[p] < mp.OW[p]) {
          any = true;
#pragma unroll
          for (int j = 0; j < VN; ++j) g.v[j] += cur[p].v[j];
        }
      }
      T* dst = xrow + (int64_t)iw * ldx + cg * VN;
      if (acc) {
        if (!any) continue;
        const Vec<T> o = Vec<T>::load(dst);
#pragma unroll
        for (int j = 0; j < VN; ++j) g.v[j] += o.v[j];
      }
      g.store(dst);
    }
  }
}

// ------------------------------------------------------------------------------------------
// Bilinear, align_corners=True
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void ac_coord(int o, float scale, int in_size, int& lo, int& hi, float& f) {
  float src = (float)o * scale;
  lo = (int)floorf(src);
  if (lo > in_size - 1) lo = in_size - 1;
  hi = min(lo + 1, in_size - 1);
  f = src - (float)lo;
}

template <typename T>
__global__ void bilinear_ac_fwd_kernel(const T* __restrict__ x, int ldx, int IH, int IW, int C, T* __restrict__ y,
                                       int ldy, int OH, int OW, float sh, float sw, int64_t total) {
  pdl_prologue();
  constexpr int VN = Vec<T>::N;
  const int cgs = C / VN;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int cg = (int)(i % cgs);
    int64_t p = i / cgs;
    int ow = (int)(p % OW);
    int64_t t = p / OW;
    int oh = (int)(t % OH);
    int n = (int)(t / OH);
    int y0, y1, x0, x1;
    float fy, fx;
    ac_coord(oh, sh, IH, y0, y1, fy);
    ac_coord(ow, sw, IW, x0, x1, fx);
    const T* base = x + (int64_t)n * IH * IW * ldx + cg * VN;
    Vec<T> tl = Vec<T>::load(base + ((int64_t)y0 * IW + x0) * ldx);
    Vec<T> tr = Vec<T>::load(base + ((int64_t)y0 * IW + x1) * ldx);
    Vec<T> bl = Vec<T>::load(base + ((int64_t)y1 * IW + x0) * ldx);
    Vec<T> br = Vec<T>::load(base + ((int64_t)y1 * IW + x1) * ldx);
    Vec<T> o;
#pragma unroll
    for (int j = 0; j < VN; ++j) {
      float top = tl.v[j] + (tr.v[j] - tl.v[j]) * fx;
      float bot = bl.v[j] + (br.v[j] - bl.v[j]) * fx;
      o.v[j] = top + (bot - top) * fy;
    }
    o.store(y + p * ldy + cg * VN);
  }
}

// adjoint: one block per (input pixel, channel chunk); gathers only over the output window whose bilinear support
// contains that input pixel
__device__ __forceinline__ void ac_window(int s, float scale, int in_size, int out_size, int& lo, int& hi) {
  if (scale <= 0.f || in_size <= 1) {
    lo = 0;
    hi = out_size - 1;
    return;
  }
  lo = max(0, (int)floorf((float)(s - 1) / scale) - 1);
  hi = min(out_size - 1, (int)ceilf((float)(s + 1) / scale) + 1);
}

template <typename T>
__global__ void bilinear_ac_bwd_kernel(const T* __restrict__ dy, int ldy, int OH, int OW, T* __restrict__ dx, int ldx,
                                       int IH, int IW, int C, float sh, float sw, int acc) {
  pdl_prologue();
  constexpr int VN = Vec<T>::N;
  extern __shared__ float sacc[];
  const int sx = blockIdx.x % IW;
  const int sy = (blockIdx.x / IW) % IH;
  const int n = blockIdx.x / (IW * IH);
  const int cgs = C / VN;
  const int cg = blockIdx.y * blockDim.x + threadIdx.x;
  int oh0, oh1, ow0, ow1;
  ac_window(sy, sh, IH, OH, oh0, oh1);
  ac_window(sx, sw, IW, OW, ow0, ow1);
  const int nh = oh1 - oh0 + 1, nw = ow1 - ow0 + 1;
  float a[VN];
#pragma unroll
  for (int j = 0; j < VN; ++j) a[j] = 0.f;
  if (cg < cgs) {
    for (int p = threadIdx.y; p < nh * nw; p += blockDim.y) {
      const int oh = oh0 + p / nw, ow = ow0 + p % nw;
      int y0, y1, x0, x1;
      float fy, fx;
      ac_coord(oh, sh, IH, y0, y1, fy);
      ac_coord(ow, sw, IW, x0, x1, fx);
      const float wy = (y0 == sy ? 1.f - fy : 0.f) + (y1 == sy ? fy : 0.f);
      const float wx = (x0 == sx ? 1.f - fx : 0.f) + (x1 == sx ? fx : 0.f);
      const float wgt = wy * wx;
      if (wgt == 0.f) continue;
      Vec<T> d = Vec<T>::load(dy + (((int64_t)n * OH + oh) * OW + ow) * ldy + cg * VN);
#pragma unroll
      for (int j = 0; j < VN; ++j) a[j] = fmaf(d.v[j], wgt, a[j]);
    }
  }
  float* mine = sacc + ((size_t)threadIdx.y * blockDim.x + threadIdx.x) * VN;
#pragma unroll
  for (int j = 0; j < VN; ++j) mine[j] = a[j];
  __syncthreads();
  for (int s2 = blockDim.y >> 1; s2 >= 1; s2 >>= 1) {
    if ((int)threadIdx.y < s2) {
      const float* o = sacc + ((size_t)(threadIdx.y + s2) * blockDim.x + threadIdx.x) * VN;
#pragma unroll
      for (int j = 0; j < VN; ++j) mine[j] += o[j];
    }
    __syncthreads();
  }
  if (threadIdx.y == 0 && cg < cgs) {
    T* dst = dx + (((int64_t)n * IH + sy) * IW + sx) * ldx + cg * VN;
    Vec<T> r;
    if (acc) {
      r = Vec<T>::load(dst);
#pragma unroll
      for (int j = 0; j < VN; ++j) r.v[j] += mine[j];
    } else {
#pragma unroll
      for (int j = 0; j < VN; ++j) r.v[j] = mine[j];
    }
    r.store(dst);
  }
}

// ------------------------------------------------------------------------------------------
// Attention gating
// ------------------------------------------------------------------------------------------
template <typename T, bool RELU>
__global__ void gate_mul_fwd_kernel(const T* __restrict__ feat, int ldf, const float* __restrict__ logits, int nseg,
                                    int att, T* __restrict__ out, int ldo, int64_t R, int C) {
  pdl_prologue();
  constexpr int VN = Vec<T>::N;
  const int cgs = C / VN;
  const int64_t total = R * cgs;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cgs;
    const int c0 = (int)(i - r * cgs) * VN;
    const float g = __ldg(logits + r * nseg + att);
    Vec<T> v = Vec<T>::load(feat + r * ldf + c0);
#pragma unroll
    for (int j = 0; j < VN; ++j) v.v[j] = (RELU ? fmaxf(v.v[j], 0.f) : v.v[j]) * g;
    v.store(out + r * ldo + c0);
  }
}

// one warp per pixel
template <typename T, bool RELU>
__global__ void gate_mul_bwd_kernel(const T* __restrict__ dout, int ldd, const T* __restrict__ feat, int ldf,
                                    const float* __restrict__ logits, int nseg, int att, T* __restrict__ dfeat,
                                    int lddf, int acc, float* __restrict__ dlogits, int64_t R, int C) {
  pdl_prologue();
  constexpr int VN = Vec<T>::N;
  const int cgs = C / VN;
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < R; r += nwarps) {
    const float g = __ldg(logits + r * nseg + att);
    float s = 0.f;
    for (int cg = lane; cg < cgs; cg += 32) {
      Vec<T> d = Vec<T>::load(dout + r * ldd + cg * VN);
      Vec<T> f = Vec<T>::load(feat + r * ldf + cg * VN);
      Vec<T> o;
      if (acc) o = Vec<T>::load(dfeat + r * lddf + cg * VN);
#pragma unroll
      for (int j = 0; j < VN; ++j) {
        float fr = RELU ? fmaxf(f.v[j], 0.f) : f.v[j];
        s = fmaf(d.v[j], fr, s);
        float gf = (!RELU || f.v[j] > 0.f) ? d.v[j] * g : 0.f;
        o.v[j] = acc ? o.v[j] + gf : gf;
      }
      o.store(dfeat + r * lddf + cg * VN);
    }
    s = warp_sum(s);
    if (lane == 0 && dlogits) dlogits[r * nseg + att] += s;
  }
}


// ------------------------------------------------------------------------------------------
// Variant-B gating: softmax attention gate with a hard threshold, nearest-neighbour resize
// ------------------------------------------------------------------------------------------
// gate[r] = p > thr ? p : 0 with p = softmax(logits[r, :])[sel]        (90AttentionSingle2/BAISNet.py:743-746)
__global__ void softmax_gate_fwd_kernel(const float* __restrict__ logits, int64_t rows, int C, int sel, float thr,
                                        float* __restrict__ gate) {
  pdl_prologue();
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
    const float* x = logits + r * C;
    float mx = x[0];
    for (int c = 1; c < C; ++c) mx = fmaxf(mx, x[c]);
    float se = 0.f;
    for (int c = 0; c < C; ++c) se += expf(x[c] - mx);
    const float p = expf(x[sel] - mx) / se;
    gate[r] = p > thr ? p : 0.f;
  }
}
// dlogits[r, j] += dgate[r] * p_sel * (delta(j, sel) - p_j) where the gate passed, 0 elsewhere (tf.where adjoint)
__global__ void softmax_gate_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ dgate, int64_t rows,
                                        int C, int sel, float thr, float* __restrict__ dlogits, int accumulate) {
  pdl_prologue();
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
    const float* x = logits + r * C;
    float mx = x[0];
    for (int c = 1; c < C; ++c) mx = fmaxf(mx, x[c]);
    float se = 0.f;
    for (int c = 0; c < C; ++c) se += expf(x[c] - mx);
    const float inv = 1.f / se;
    const float ps = expf(x[sel] - mx) * inv;
    const float g = ps > thr ? dgate[r] * ps : 0.f;
    for (int c = 0; c < C; ++c) {
      const float pj = expf(x[c] - mx) * inv;
      const float v = g * ((c == sel ? 1.f : 0.f) - pj);
      dlogits[r * C + c] = accumulate ? dlogits[r * C + c] + v : v;
    }
  }
}

__device__ __forceinline__ int nn_src(int o, float scale, int in_size) { return min((int)floorf((float)o * scale), in_size - 1); }

// TF1 resize_nearest_neighbor (align_corners=False): src = min(floor(dst * in/out), in - 1)
template <typename T>
__global__ void resize_nearest_fwd_kernel(const T* __restrict__ x, int ldx, int IH, int IW, int C, T* __restrict__ y,
                                          int ldy, int OH, int OW, float sh, float sw, int64_t total) {
  pdl_prologue();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    int64_t p = i / C;
    const int ow = (int)(p % OW);
    int64_t t = p / OW;
    const int oh = (int)(t % OH);
    const int n = (int)(t / OH);
    const int ih = nn_src(oh, sh, IH), iw = nn_src(ow, sw, IW);
    y[p * ldy + c] = x[(((int64_t)n * IH + ih) * IW + iw) * ldx + c];
  }
}
// adjoint in gather form: every input pixel sums the output pixels that map onto it
template <typename T>
__global__ void resize_nearest_bwd_kernel(const T* __restrict__ dy, int ldy, int OH, int OW, T* __restrict__ dx, int ldx,
                                          int IH, int IW, int C, float sh, float sw, int acc, int64_t total) {
  pdl_prologue();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    int64_t p = i / C;
    const int iw = (int)(p % IW);
    int64_t t = p / IW;
    const int ih = (int)(t % IH);
    const int n = (int)(t / IH);
    const int oh_lo = max(0, (int)floorf((float)ih / sh) - 1), oh_hi = min(OH - 1, (int)ceilf((float)(ih + 1) / sh) + 1);
    const int ow_lo = max(0, (int)floorf((float)iw / sw) - 1), ow_hi = min(OW - 1, (int)ceilf((float)(iw + 1) / sw) + 1);
    float a = acc ? to_f32(dx[p * ldx + c]) : 0.f;
    for (int oh = oh_lo; oh <= oh_hi; ++oh) {
      if (nn_src(oh, sh, IH) != ih) continue;
      for (int ow = ow_lo; ow <= ow_hi; ++ow)
        if (nn_src(ow, sw, IW) == iw) a += to_f32(dy[(((int64_t)n * OH + oh) * OW + ow) * ldy + c]);
    }
    dx[p * ldx + c] = from_f32<T>(a);
  }
}


// ------------------------------------------------------------------------------------------
// Spatial subsampling x[:, ::s, ::s, :] and its adjoint: a strided 1x1 VALID convolution (conv3_1_1x1_proj / _reduce,
// BAISPSPNet.py:310,314) is the stride-1 1x1 convolution of the subsampled tensor, which the tcgen05 path takes.
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void subsample_fwd_kernel(const T* __restrict__ x, int ldx, int IH, int IW, int C, int s, T* __restrict__ y,
                                     int ldy, int OH, int OW, int64_t total) {
  pdl_prologue();
  constexpr int VN = Vec<T>::N;
  const int cgs = C / VN;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cg = (int)(i % cgs);
    int64_t p = i / cgs;
    const int ow = (int)(p % OW);
    int64_t t = p / OW;
    const int oh = (int)(t % OH);
    const int n = (int)(t / OH);
    Vec<T> v = Vec<T>::load(x + (((int64_t)n * IH + oh * s) * IW + ow * s) * ldx + cg * VN);
    v.store(y + p * ldy + cg * VN);
  }
}
// dx[n,h,w,:] (+)= (h % s == 0 && w % s == 0) ? dy[n,h/s,w/s,:] : 0
template <typename T>
__global__ void subsample_bwd_kernel(const T* __restrict__ dy, int ldy, int OH, int OW, int s, T* __restrict__ dx, int ldx,
                                     int IH, int IW, int C, int acc, int64_t total) {
  pdl_prologue();
  constexpr int VN = Vec<T>::N;
  const int cgs = C / VN;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cg = (int)(i % cgs);
    int64_t p = i / cgs;
    const int iw = (int)(p % IW);
    int64_t t = p / IW;
    const int ih = (int)(t % IH);
    const int n = (int)(t / IH);
    const bool hit = (ih % s == 0) && (iw % s == 0) && (ih / s < OH) && (iw / s < OW);
    if (!hit && acc) continue;
    Vec<T> g = Vec<T>::zero();
    if (hit) {
      g = Vec<T>::load(dy + (((int64_t)n * OH + ih / s) * OW + iw / s) * ldy + cg * VN);
      if (acc) {
        Vec<T> o = Vec<T>::load(dx + p * ldx + cg * VN);
#pragma unroll
        for (int j = 0; j < VN; ++j) g.v[j] += o.v[j];
      }
    }
    g.store(dx + p * ldx + cg * VN);
  }
}

// ------------------------------------------------------------------------------------------
// Click map + NHWC4 packing.  No fast-math: the table holds float32 subnormals.
// ------------------------------------------------------------------------------------------
__global__ void clickmap_pack_kernel(const uint8_t* __restrict__ img_u8, const float* __restrict__ img_f32,
                                     const int32_t* __restrict__ clicks, const float* __restrict__ lut,
                                     int64_t lut_len, float4* __restrict__ out, int B, int H, int W) {
  pdl_prologue();
  const int64_t total = (int64_t)B * H * W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int x = (int)(i % W);
    int64_t t = i / W;
    int y = (int)(t % H);
    int b = (int)(t / H);
    int64_t dy = y - clicks[2 * b], dx = x - clicks[2 * b + 1];
    int64_t d2 = dx * dx + dy * dy;
    float m = d2 < lut_len ? __ldg(lut + d2) : 0.f;
    float4 o;
    if (img_u8) {
      const uint8_t* p = img_u8 + i * 3;
      o.x = (float)p[0] / 255.0f;
      o.y = (float)p[1] / 255.0f;
      o.z = (float)p[2] / 255.0f;
    } else {
      const float* p = img_f32 + i * 3;
      o.x = p[0];
      o.y = p[1];
      o.z = p[2];
    }
    o.w = m;
    out[i] = o;
  }
}

// ------------------------------------------------------------------------------------------
// Losses (fused forward + gradient), SGD, predictions
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void block_sum_atomic(double v, double* dst) {
  __shared__ double wsum[32];
  v = warp_sum(v);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) wsum[w] = v;
  __syncthreads();
  if (w == 0) {
    v = lane < (int)((blockDim.x + 31) >> 5) ? wsum[lane] : 0.0;
    v = warp_sum(v);
    if (lane == 0) atomicAdd(dst, v);
  }
}

// one logit: loss term and d loss / d logit.  softplus(-x) = log1p(e) + max(-x, 0) with e = exp(-|x|) in (0, 1]:
// log1p through the two-term series below 2^-10 (error e^3/3 < 3e-10) and log(1 + e) above it; sigmoid(-x) from the
// same e.  (expf / log1pf / the IEEE division made this kernel instruction-bound: 0.35 ms for 64 M logits against
// 0.12 ms of HBM time.)
__device__ __forceinline__ float wbce_term(float x, float z, float q, float gscale, float& grad) {
  const float w = fmaf(q - 1.f, z, 1.f);
  const float e = __expf(-fabsf(x));
  const float l1p = e < 9.765625e-4f ? e * fmaf(-0.5f, e, 1.f) : __logf(1.f + e);
  const float sp = l1p + fmaxf(-x, 0.f);
  const float r = __fdividef(1.f, 1.f + e);
  const float sneg = x >= 0.f ? e * r : r;          // sigmoid(-x), stable on both sides
  grad = gscale * ((1.f - z) - w * sneg);
  return fmaf(1.f - z, x, w * sp);
}

__global__ void __launch_bounds__(256)
wbce_kernel(const float* __restrict__ logits, const float* __restrict__ labels, float q, double scale,
            float gscale, int64_t n, double* __restrict__ loss_acc, float* __restrict__ dlogits) {
  pdl_prologue();
  double acc = 0.0;
  const bool vec = ((reinterpret_cast<uintptr_t>(logits) | reinterpret_cast<uintptr_t>(labels) |
                     reinterpret_cast<uintptr_t>(dlogits)) & 15) == 0;
  const int64_t n4 = vec ? n / 4 : 0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 x = __ldcs(reinterpret_cast<const float4*>(logits) + i);
    const float4 z = __ldcs(reinterpret_cast<const float4*>(labels) + i);
    float4 g;
    float part = wbce_term(x.x, z.x, q, gscale, g.x);
    part += wbce_term(x.y, z.y, q, gscale, g.y);
    part += wbce_term(x.z, z.z, q, gscale, g.z);
    part += wbce_term(x.w, z.w, q, gscale, g.w);
    acc += (double)part;
    if (dlogits) __stcs(reinterpret_cast<float4*>(dlogits) + i, g);
  }
  for (int64_t i = n4 * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float g;
    acc += (double)wbce_term(logits[i], labels[i], q, gscale, g);
    if (dlogits) dlogits[i] = g;
  }
  block_sum_atomic(acc * scale, loss_acc);
}

__global__ void softmax_ce_kernel(const float* __restrict__ logits, const int32_t* __restrict__ labels, int64_t rows,
                                  int C, double scale, float gscale, double* __restrict__ loss_acc,
                                  float* __restrict__ dlogits) {
  pdl_prologue();
  double acc = 0.0;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
    const float* x = logits + r * C;
    float mx = x[0];
    for (int c = 1; c < C; ++c) mx = fmaxf(mx, x[c]);
    float se = 0.f;
    for (int c = 0; c < C; ++c) se += expf(x[c] - mx);
    int lab = labels[r];
    float lse = logf(se) + mx;
    acc += (double)(lse - x[lab]);
    if (dlogits) {
      float inv = 1.f / se;
      for (int c = 0; c < C; ++c) {
        float p = expf(x[c] - mx) * inv;
        dlogits[r * C + c] = gscale * (p - (c == lab ? 1.f : 0.f));
      }
    }
  }
  block_sum_atomic(acc * scale, loss_acc);
}

// F4 (back/8AttentionU): Net.sigmoid on the C-channel decoder logits and its adjoint; weighted BCE that reads channel
// `sel` of C-channel "logits" (cal_loss takes tf.split(segment, 2, axis=3)[1]) and writes a full-width gradient.
__global__ void sigmoid_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n) {
  pdl_prologue();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = x[i];
    const float e = expf(-fabsf(v));
    y[i] = v >= 0.f ? 1.f / (1.f + e) : e / (1.f + e);
  }
}

__global__ void sigmoid_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ dx,
                                   int64_t n, int accumulate) {
  pdl_prologue();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float s = y[i];
    const float g = dy[i] * s * (1.f - s);
    dx[i] = accumulate ? dx[i] + g : g;
  }
}

__global__ void __launch_bounds__(256)
wbce_sel_kernel(const float* __restrict__ logits, int C, int sel, const float* __restrict__ labels, float q,
                double scale, float gscale, int64_t rows, double* __restrict__ loss_acc, float* __restrict__ dlogits) {
  pdl_prologue();
  double acc = 0.0;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
    float g;
    acc += (double)wbce_term(logits[r * C + sel], labels[r], q, gscale, g);
    if (dlogits)
      for (int c = 0; c < C; ++c) dlogits[r * C + c] = c == sel ? g : 0.f;
  }
  block_sum_atomic(acc * scale, loss_acc);
}

__global__ void sgd_kernel(float* __restrict__ w, const float* __restrict__ g, const float* __restrict__ lr_dev,
                           int64_t n, bf16* __restrict__ wb) {
  pdl_prologue();
  const float lr = __ldg(lr_dev);
  const int64_t n4 = n >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 a = reinterpret_cast<float4*>(w)[i];
    float4 b = reinterpret_cast<const float4*>(g)[i];
    a.x -= lr * b.x;
    a.y -= lr * b.y;
    a.z -= lr * b.z;
    a.w -= lr * b.w;
    reinterpret_cast<float4*>(w)[i] = a;
    if (wb) {
      reinterpret_cast<uint2*>(wb)[i] = make_uint2(h16_pack(a.x, a.y), h16_pack(a.z, a.w));
    }
  }
  if (blockIdx.x == 0) {
    for (int64_t i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) {
      float a = w[i] - lr * g[i];
      w[i] = a;
      if (wb) wb[i] = h16_from(a);
    }
  }
}

__global__ void threshold_kernel(const float* __restrict__ x, float thr, int32_t* __restrict__ out, int64_t n) {
  pdl_prologue();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = x[i] > thr ? 1 : 0;
}

__global__ void argmax_kernel(const float* __restrict__ x, int64_t rows, int C, int32_t* __restrict__ out) {
  pdl_prologue();
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
    const float* p = x + r * C;
    float b = p[0];
    int bi = 0;
    for (int c = 1; c < C; ++c)
      if (p[c] > b) {
        b = p[c];
        bi = c;
      }
    out[r] = bi;
  }
}

// TF1 legacy resize (src = dst*in/out) + argmax(sigmoid(.)) == argmax of the interpolated logits only if sigmoid is
// monotone per channel -- it is, but the reference takes sigmoid first in float32, so do the same to keep ties equal.
__global__ void upsample_legacy_argmax_kernel(const float* __restrict__ x, int B, int PH, int PW, int C, int SH, int SW,
                                              float sh, float sw, int32_t* __restrict__ out) {
  pdl_prologue();
  const int64_t total = (int64_t)B * SH * SW;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int ox = (int)(i % SW);
    int64_t t = i / SW;
    int oy = (int)(t % SH);
    int b = (int)(t / SH);
    float sy = (float)oy * sh, sx = (float)ox * sw;
    int y0 = min((int)floorf(sy), PH - 1), x0 = min((int)floorf(sx), PW - 1);
    int y1 = min(y0 + 1, PH - 1), x1 = min(x0 + 1, PW - 1);
    float fy = sy - (float)y0, fx = sx - (float)x0;
    const float* base = x + (int64_t)b * PH * PW * C;
    float best = -1.f;
    int bi = 0;
    for (int c = 0; c < C; ++c) {
      float tl = base[((int64_t)y0 * PW + x0) * C + c], tr = base[((int64_t)y0 * PW + x1) * C + c];
      float bl = base[((int64_t)y1 * PW + x0) * C + c], br = base[((int64_t)y1 * PW + x1) * C + c];
      float top = tl + (tr - tl) * fx, bot = bl + (br - bl) * fx;
      float v = top + (bot - top) * fy;
      float s = 1.f / (1.f + expf(-v));
      if (s > best) {
        best = s;
        bi = c;
      }
    }
    out[i] = bi;
  }
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ s, bf16* __restrict__ d, int64_t n) {
  pdl_prologue();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    d[i] = h16_from(s[i]);
}
__global__ void cast_bf16_f32_kernel(const bf16* __restrict__ s, float* __restrict__ d, int64_t n) {
  pdl_prologue();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    d[i] = h16_to(s[i]);
}
__global__ void relu_bwd_f32_kernel(float* __restrict__ dy, const float* __restrict__ y, int64_t n) {
  pdl_prologue();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    if (!(y[i] > 0.f)) dy[i] = 0.f;
}

}  // namespace basi

using namespace basi;

#define DISPATCH_T(dtype, ...)             \
  if ((dtype) == BASI_F32) {               \
    typedef float T;                       \
    __VA_ARGS__                            \
  } else {                                 \
    typedef bf16 T;                        \
    __VA_ARGS__                            \
  }

// x (float32) -> [hi | mid | lo] bf16 parts with x == hi + mid + lo up to 2^-24 |x|: hi = bf16(x), mid = bf16(x - hi),
// lo = bf16(x - hi - mid) (both differences are exact in float32).  The operands of the split-operand (fp32-grade)
// tcgen05 convolutions: six bf16 MMAs (hi*hi, hi*mid, mid*hi, hi*lo, lo*hi, mid*mid) reproduce the float32 product.
__global__ void __launch_bounds__(256) split3_kernel(const float* __restrict__ x, int ldx, int C,
                                                     __nv_bfloat16* __restrict__ y, int ldy, int64_t total, int parts) {
  pdl_prologue();
  const int q = C / 4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t px = i / q;
    const int c = (int)(i - px * q) * 4;
    const float4 v = *reinterpret_cast<const float4*>(x + px * ldx + c);
    const float in[4] = {v.x, v.y, v.z, v.w};
    __nv_bfloat16 h[4], m[4], l[4];     // (always bfloat16: the parts rely on its float32 exponent range)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      h[j] = __float2bfloat16_rn(in[j]);
      const float r1 = in[j] - __bfloat162float(h[j]);
      m[j] = __float2bfloat16_rn(r1);
      l[j] = __float2bfloat16_rn(r1 - __bfloat162float(m[j]));
    }
    __nv_bfloat16* o = y + px * ldy + c;
    *reinterpret_cast<uint2*>(o) = *reinterpret_cast<const uint2*>(h);
    *reinterpret_cast<uint2*>(o + C) = *reinterpret_cast<const uint2*>(m);
    if (parts == 3) *reinterpret_cast<uint2*>(o + 2 * C) = *reinterpret_cast<const uint2*>(l);
  }
}

// ------------------------------------------------------------------------------------------
// A3 / F3: label encodings and click sampling on the device (integer work, bit-exact with the numpy expressions).
//   mode 0  binary            back/2AddClass/BAISData.py:150-151     (ann == num) ? 1 : 0
//   mode 1  4-class border    back/4BorderClass/BAISData.py:143-160  has_255=True: 255 -> 171, num -> 85, (v - 1) // 84
//                             in uint8 arithmetic: 0 wraps to 255 -> 3 (background), 85 -> 1, 171 -> 2, others (k-1)//84
//   mode 2  3-class           same file, has_255=False: 255 -> 0, num -> 128, (v - 1) // 127 (0 wraps to 255 -> 2)
//   mode 3  COCO              back/5COCO/BAISData.py:361-369         attention > 0 ? 2 : (sum > 0 ? 1 : 0)
// ------------------------------------------------------------------------------------------
__global__ void label_encode_kernel(const uint8_t* __restrict__ ann, const uint8_t* __restrict__ att,
                                    const int32_t* __restrict__ nums, int mode, int32_t* __restrict__ out_i32,
                                    float* __restrict__ out_f32, int64_t per_image, int64_t total) {
  pdl_prologue();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / per_image);
    uint8_t v = ann[i];
    int lab;
    if (mode == 0) {
      lab = (int)v == nums[b] ? 1 : 0;
    } else if (mode == 1) {
      if (v == 255) v = 171;
      if ((int)v == nums[b]) v = 85;
      lab = (uint8_t)(v - 1) / 84;
    } else if (mode == 2) {
      if (v == 255) v = 0;
      if ((int)v == nums[b]) v = 128;
      lab = (uint8_t)(v - 1) / 127;
    } else {
      lab = att[i] > 0 ? 2 : (v > 0 ? 1 : 0);
    }
    if (out_i32) out_i32[i] = lab;
    if (out_f32) out_f32[i] = (float)lab;
  }
}

// counts[b] = number of pixels of image b whose label equals `target` (np.argwhere(ann == target) has that many rows)
template <typename TL>
__global__ void __launch_bounds__(256) click_count_kernel(const TL* __restrict__ lab, TL target, int per_image,
                                                          int32_t* __restrict__ counts) {
  pdl_prologue();
  __shared__ int wsum[8];
  const TL* p = lab + (int64_t)blockIdx.x * per_image;
  int c = 0;
  for (int i = threadIdx.x; i < per_image; i += 256) c += p[i] == target ? 1 : 0;
  c = (int)warp_sum((float)c);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < 8; ++w) t += wsum[w];
    counts[blockIdx.x] = t;
  }
}

// clicks[b] = ratio * (row, col) of the k[b]-th pixel (row-major order) of image b whose label equals `target` --
// np.argwhere(ann == target)[k] * ratio (back/2AddClass/BAISData.py:63-66); k out of range leaves (-1, -1)
template <typename TL>
__global__ void __launch_bounds__(256) click_select_kernel(const TL* __restrict__ lab, TL target, int H, int W,
                                                           const int32_t* __restrict__ k, int ratio,
                                                           int32_t* __restrict__ clicks) {
  pdl_prologue();
  __shared__ int wcnt[8];
  __shared__ int base_s;
  const int per_image = H * W;
  const TL* p = lab + (int64_t)blockIdx.x * per_image;
  const int want = k[blockIdx.x];
  if (threadIdx.x == 0) {
    base_s = 0;
    clicks[2 * blockIdx.x] = -1;
    clicks[2 * blockIdx.x + 1] = -1;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i0 = 0; i0 < per_image; i0 += 256) {       // 256 consecutive pixels per round, ranked with ballots
    const int i = i0 + threadIdx.x;
    const bool hit = i < per_image && p[i] == target;
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if (lane == 0) wcnt[warp] = __popc(m);
    __syncthreads();
    int before = base_s;
    for (int w = 0; w < warp; ++w) before += wcnt[w];
    const int rank = before + __popc(m & ((1u << lane) - 1u));
    if (hit && rank == want) {
      clicks[2 * blockIdx.x] = (i / W) * ratio;
      clicks[2 * blockIdx.x + 1] = (i % W) * ratio;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
      for (int w = 0; w < 8; ++w) t += wcnt[w];
      base_s += t;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// 2x2 / stride 2 VALID max-pool (slim.max_pool2d(net, [2, 2]) in vgg_16, slim/nets/vgg.py:188-196; variant B trunk).
// Non-overlapping windows: the window index of the maximum (first maximum wins, row-major) is kept in one byte per
// element; the adjoint routes dy to that element and writes zeros elsewhere (also into a dropped odd row / column).
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void maxpool2s2_fwd_kernel(const T* __restrict__ x, int ldx, int H, int W, int C, T* __restrict__ y, int ldy,
                                      int OH, int OW, uint8_t* __restrict__ amax, int64_t total) {
  pdl_prologue();
  constexpr int VN = Vec<T>::N;
  const int cgs = C / VN;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cg = (int)(i % cgs);
    const int64_t p = i / cgs;
    const int ow = (int)(p % OW);
    const int64_t t = p / OW;
    const int oh = (int)(t % OH);
    const int n = (int)(t / OH);
    float best[VN];
    uint8_t bi[VN];
#pragma unroll
    for (int j = 0; j < VN; ++j) { best[j] = -INFINITY; bi[j] = 0; }
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        Vec<T> v = Vec<T>::load(x + (((int64_t)n * H + oh * 2 + r) * W + ow * 2 + q) * ldx + cg * VN);
#pragma unroll
        for (int j = 0; j < VN; ++j)
          if (v.v[j] > best[j]) { best[j] = v.v[j]; bi[j] = (uint8_t)(r * 2 + q); }
      }
    Vec<T> o;
#pragma unroll
    for (int j = 0; j < VN; ++j) o.v[j] = best[j];
    o.store(y + p * ldy + cg * VN);
#pragma unroll
    for (int j = 0; j < VN; ++j) amax[p * C + cg * VN + j] = bi[j];
  }
}

template <typename T>
__global__ void maxpool2s2_bwd_kernel(const T* __restrict__ dy, int lddy, int OH, int OW, const uint8_t* __restrict__ amax,
                                      T* __restrict__ dx, int lddx, int H, int W, int C, int accumulate, int64_t total) {
  pdl_prologue();
  constexpr int VN = Vec<T>::N;
  const int cgs = C / VN;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cg = (int)(i % cgs);
    const int64_t p = i / cgs;
    const int iw = (int)(p % W);
    const int64_t t = p / W;
    const int ih = (int)(t % H);
    const int n = (int)(t / H);
    const int oh = ih >> 1, ow = iw >> 1;
    Vec<T> o = Vec<T>::zero();
    if (oh < OH && ow < OW) {
      const int64_t op = ((int64_t)n * OH + oh) * OW + ow;
      Vec<T> g = Vec<T>::load(dy + op * lddy + cg * VN);
      const int pos = (ih & 1) * 2 + (iw & 1);
#pragma unroll
      for (int j = 0; j < VN; ++j) o.v[j] = amax[op * C + cg * VN + j] == pos ? g.v[j] : 0.f;
    }
    T* d = dx + p * lddx + cg * VN;
    if (accumulate) {
      Vec<T> old = Vec<T>::load(d);
#pragma unroll
      for (int j = 0; j < VN; ++j) o.v[j] += old.v[j];
    }
    o.store(d);
  }
}

// t[i] = one_hot(label[i], 2) as float32 pairs [1 - z, z] (tf.one_hot(labels, depth=2), variant B cal_loss)
__global__ void onehot2_kernel(const float* __restrict__ lab, float2* __restrict__ out, int64_t n) {
  pdl_prologue();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float z = lab[i] > 0.5f ? 1.f : 0.f;
    out[i] = make_float2(1.f - z, z);
  }
}

// Adjoint of y = relu(conv + bias) for 16-bit / float32 activations: dy *= (y > 0) in place and
// dbias[c] += sum over pixels of the masked dy (block partial sums in shared memory, one atomic per channel and block)
template <typename T>
__global__ void __launch_bounds__(256) bias_relu_bwd_kernel(T* __restrict__ dy, int lddy, const T* __restrict__ y, int ldy,
                                                            int C, int64_t R, int relu, float* __restrict__ dbias) {
  pdl_prologue();
  extern __shared__ float bsum[];          // [C]
  constexpr int VN = Vec<T>::N;
  const int cgs = C / VN;
  for (int c = threadIdx.x; c < C; c += blockDim.x) bsum[c] = 0.f;
  __syncthreads();
  const int64_t total = R * cgs;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cg = (int)(i % cgs);
    const int64_t r = i / cgs;
    Vec<T> g = Vec<T>::load(dy + r * lddy + cg * VN);
    if (relu) {
      Vec<T> v = Vec<T>::load(y + r * ldy + cg * VN);
#pragma unroll
      for (int j = 0; j < VN; ++j) g.v[j] = v.v[j] > 0.f ? g.v[j] : 0.f;
      g.store(dy + r * lddy + cg * VN);
    }
    if (dbias) {
#pragma unroll
      for (int j = 0; j < VN; ++j) atomicAdd(&bsum[cg * VN + j], g.v[j]);
    }
  }
  __syncthreads();
  if (dbias)
    for (int c = threadIdx.x; c < C; c += blockDim.x) atomicAdd(dbias + c, bsum[c]);
}

// Stand-alone ReLU of a MATERIALISED residual sum and its adjoint dX (+)= dY * (Y > 0): the hand-unrolled trunk of
// back/8AttentionU/BAISNet.py:163-165 reads both the junction sum (next 1x1_reduce / 1x1_proj) and its ReLU (shortcut)
template <typename T>
__global__ void relu_fwd_kernel(const T* __restrict__ x, int ldx, T* __restrict__ y, int ldy, int C, int64_t total) {
  pdl_prologue();
  constexpr int VN = Vec<T>::N;
  const int cgs = C / VN;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cg = (int)(i % cgs);
    const int64_t r = i / cgs;
    Vec<T> v = Vec<T>::load(x + r * ldx + cg * VN);
#pragma unroll
    for (int j = 0; j < VN; ++j) v.v[j] = v.v[j] > 0.f ? v.v[j] : 0.f;
    v.store(y + r * ldy + cg * VN);
  }
}
template <typename T>
__global__ void relu_bwd_kernel(const T* __restrict__ g, int ldg, const T* __restrict__ y, int ldy, T* __restrict__ dx,
                                int ldx, int acc, int C, int64_t total) {
  pdl_prologue();
  constexpr int VN = Vec<T>::N;
  const int cgs = C / VN;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cg = (int)(i % cgs);
    const int64_t r = i / cgs;
    Vec<T> o = Vec<T>::load(g + r * ldg + cg * VN);
    const Vec<T> m = Vec<T>::load(y + r * ldy + cg * VN);
#pragma unroll
    for (int j = 0; j < VN; ++j) o.v[j] = m.v[j] > 0.f ? o.v[j] : 0.f;
    if (acc) {
      const Vec<T> old = Vec<T>::load(dx + r * ldx + cg * VN);
#pragma unroll
      for (int j = 0; j < VN; ++j) o.v[j] += old.v[j];
    }
    o.store(dx + r * ldx + cg * VN);
  }
}

// Plain tensor add (Net.add of the top-level LinkNet, BAISNet.py:244) and its adjoint (dA (+)= dOut, dB (+)= dOut)
template <typename T>
__global__ void add_fwd_kernel(const T* __restrict__ a, int lda, const T* __restrict__ b, int ldb, T* __restrict__ o,
                               int ldo, int C, int64_t total) {
  pdl_prologue();
  constexpr int VN = Vec<T>::N;
  const int cgs = C / VN;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cg = (int)(i % cgs);
    const int64_t r = i / cgs;
    Vec<T> x = Vec<T>::load(a + r * lda + cg * VN), y = Vec<T>::load(b + r * ldb + cg * VN);
#pragma unroll
    for (int j = 0; j < VN; ++j) x.v[j] += y.v[j];
    x.store(o + r * ldo + cg * VN);
  }
}
template <typename T>
__global__ void add_bwd_kernel(const T* __restrict__ g, int ldg, T* __restrict__ da, int lda, int acc_a, T* __restrict__ db,
                               int ldb, int acc_b, int C, int64_t total) {
  pdl_prologue();
  constexpr int VN = Vec<T>::N;
  const int cgs = C / VN;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cg = (int)(i % cgs);
    const int64_t r = i / cgs;
    const Vec<T> x = Vec<T>::load(g + r * ldg + cg * VN);
    if (da) {
      Vec<T> o = x;
      if (acc_a) {
        Vec<T> old = Vec<T>::load(da + r * lda + cg * VN);
#pragma unroll
        for (int j = 0; j < VN; ++j) o.v[j] += old.v[j];
      }
      o.store(da + r * lda + cg * VN);
    }
    if (db) {
      Vec<T> o = x;
      if (acc_b) {
        Vec<T> old = Vec<T>::load(db + r * ldb + cg * VN);
#pragma unroll
        for (int j = 0; j < VN; ++j) o.v[j] += old.v[j];
      }
      o.store(db + r * ldb + cg * VN);
    }
  }
}

extern "C" {

const char* basi_last_error(void) { return g_err; }
int basi_version(void) { return 100; }
int basi_sm_count(void) {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
    set_error("basi_sm_count: %s", cudaGetErrorString(cudaGetLastError()));
    return BASI_E_NOGPU;
  }
  return n;
}
int basi_set_sm_budget(int main_sms, int wgrad_ctas) {
  if (main_sms < 0 || wgrad_ctas < 0) {
    set_error("basi_set_sm_budget: negative argument");
    return BASI_E_INVALID;
  }
  basi::g_sm_budget = main_sms;
  basi::g_wgrad_ctas = wgrad_ctas;
  return BASI_OK;
}
int basi_half_format(void) { return BASI_H16_FP16; }   /* 0: the 16-bit storage type of this build is bfloat16, 1: IEEE fp16 */
int basi_memset(void* ptr, int value, int64_t bytes, void* stream) {
  cudaError_t e = cudaMemsetAsync(ptr, value, (size_t)bytes, (cudaStream_t)stream);
  if (e != cudaSuccess) {
    set_error("basi_memset: %s", cudaGetErrorString(e));
    return BASI_E_CUDA;
  }
  return BASI_OK;
}

int basi_label_encode(const uint8_t* ann, const uint8_t* attention, const int32_t* nums, int mode, int32_t* out_i32,
                      float* out_f32, int B, int64_t per_image, void* stream) {
  BASI_CHECK_ARG(ann && (out_i32 || out_f32) && B > 0 && per_image > 0 && mode >= 0 && mode <= 3,
                 "label_encode: bad argument");
  BASI_CHECK_ARG(mode == 3 ? attention != nullptr : nums != nullptr, "label_encode: mode 3 needs the attention map, "
                 "modes 0-2 the instance numbers");
  const int64_t total = (int64_t)B * per_image;
  basi::launch(label_encode_kernel, grid_for(total, 256), 256, 0, (cudaStream_t)stream, ann, attention, nums, mode,
               out_i32, out_f32, per_image, total);
  BASI_CHECK_LAUNCH("label_encode");
  return BASI_OK;
}

int basi_click_count(const void* labels, int labels_are_f32, int target, int B, int per_image, int32_t* counts,
                     void* stream) {
  BASI_CHECK_ARG(labels && counts && B > 0 && per_image > 0, "click_count: bad argument");
  if (labels_are_f32)
    basi::launch(click_count_kernel<float>, B, 256, 0, (cudaStream_t)stream, (const float*)labels, (float)target,
                 per_image, counts);
  else
    basi::launch(click_count_kernel<int32_t>, B, 256, 0, (cudaStream_t)stream, (const int32_t*)labels, (int32_t)target,
                 per_image, counts);
  BASI_CHECK_LAUNCH("click_count");
  return BASI_OK;
}

int basi_click_select(const void* labels, int labels_are_f32, int target, int B, int H, int W, const int32_t* k,
                      int ratio, int32_t* clicks, void* stream) {
  BASI_CHECK_ARG(labels && k && clicks && B > 0 && H > 0 && W > 0 && ratio > 0, "click_select: bad argument");
  if (labels_are_f32)
    basi::launch(click_select_kernel<float>, B, 256, 0, (cudaStream_t)stream, (const float*)labels, (float)target, H, W,
                 k, ratio, clicks);
  else
    basi::launch(click_select_kernel<int32_t>, B, 256, 0, (cudaStream_t)stream, (const int32_t*)labels, (int32_t)target,
                 H, W, k, ratio, clicks);
  BASI_CHECK_LAUNCH("click_select");
  return BASI_OK;
}

int basi_clickmap_pack(const void* img, int img_is_f32, const int32_t* clicks, const float* lut, int64_t lut_len,
                       float* out, int B, int H, int W, void* stream) {
  BASI_CHECK_ARG(img && clicks && lut && out && B > 0 && H > 0 && W > 0, "clickmap_pack: bad argument");
  int64_t total = (int64_t)B * H * W;
  basi::launch(clickmap_pack_kernel, grid_for(total, 256), 256, 0, (cudaStream_t)stream, img_is_f32 ? nullptr : (const uint8_t*)img, img_is_f32 ? (const float*)img : nullptr, clicks, lut, lut_len,
      (float4*)out, B, H, W);
  BASI_CHECK_LAUNCH("clickmap_pack");
  return BASI_OK;
}

static void same_pad_3s2(int in, int out, int* before) {
  int total = (out - 1) * 2 + 3 - in;
  if (total < 0) total = 0;
  *before = total / 2;
}

int basi_bias_relu_bwd(const basi_tensor* dy, const basi_tensor* y, int relu, float* dbias, void* stream) {
  BASI_CHECK_ARG(dy && y && vec_ok(dy) && vec_ok(y) && same_shape(dy, y) && dy->dtype == y->dtype && dy->c <= 8192,
                 "bias_relu_bwd: bad tensors");
  const int64_t R = pixels(dy);
  DISPATCH_T(dy->dtype, {
    const int64_t total = R * (dy->c / Vec<T>::N);
    basi::launch(bias_relu_bwd_kernel<T>, grid_for(total, 256, 4), 256, (size_t)dy->c * sizeof(float), (cudaStream_t)stream,
                 (T*)dy->ptr, dy->ld, (const T*)y->ptr, y->ld, dy->c, R, relu, dbias);
  })
  BASI_CHECK_LAUNCH("bias_relu_bwd");
  return BASI_OK;
}

int basi_add_fwd(const basi_tensor* a, const basi_tensor* b, const basi_tensor* out, void* stream) {
  BASI_CHECK_ARG(a && b && out && vec_ok(a) && vec_ok(b) && vec_ok(out) && same_shape(a, b) && same_shape(a, out) &&
                     a->dtype == b->dtype && a->dtype == out->dtype, "add fwd: bad tensors");
  DISPATCH_T(a->dtype, {
    const int64_t total = pixels(a) * (a->c / Vec<T>::N);
    basi::launch(add_fwd_kernel<T>, grid_for(total, 256), 256, 0, (cudaStream_t)stream, (const T*)a->ptr, a->ld,
                 (const T*)b->ptr, b->ld, (T*)out->ptr, out->ld, a->c, total);
  })
  BASI_CHECK_LAUNCH("add_fwd");
  return BASI_OK;
}

int basi_add_bwd(const basi_tensor* dout, const basi_tensor* da, int acc_a, const basi_tensor* db, int acc_b,
                 void* stream) {
  BASI_CHECK_ARG(dout && vec_ok(dout) && (da || db), "add bwd: bad tensors");
  BASI_CHECK_ARG((!da || (vec_ok(da) && same_shape(da, dout) && da->dtype == dout->dtype)) &&
                     (!db || (vec_ok(db) && same_shape(db, dout) && db->dtype == dout->dtype)), "add bwd: bad tensors");
  DISPATCH_T(dout->dtype, {
    const int64_t total = pixels(dout) * (dout->c / Vec<T>::N);
    basi::launch(add_bwd_kernel<T>, grid_for(total, 256), 256, 0, (cudaStream_t)stream, (const T*)dout->ptr, dout->ld,
                 da ? (T*)da->ptr : (T*)nullptr, da ? da->ld : 0, acc_a, db ? (T*)db->ptr : (T*)nullptr, db ? db->ld : 0,
                 acc_b, dout->c, total);
  })
  BASI_CHECK_LAUNCH("add_bwd");
  return BASI_OK;
}

int basi_relu_fwd(const basi_tensor* x, const basi_tensor* y, void* stream) {
  BASI_CHECK_ARG(x && y && vec_ok(x) && vec_ok(y) && same_shape(x, y) && x->dtype == y->dtype, "relu fwd: bad tensors");
  DISPATCH_T(x->dtype, {
    const int64_t total = pixels(x) * (x->c / Vec<T>::N);
    basi::launch(relu_fwd_kernel<T>, grid_for(total, 256), 256, 0, (cudaStream_t)stream, (const T*)x->ptr, x->ld,
                 (T*)y->ptr, y->ld, x->c, total);
  })
  BASI_CHECK_LAUNCH("relu_fwd");
  return BASI_OK;
}

int basi_relu_bwd(const basi_tensor* dy, const basi_tensor* y, const basi_tensor* dx, int accumulate, void* stream) {
  BASI_CHECK_ARG(dy && y && dx && vec_ok(dy) && vec_ok(y) && vec_ok(dx) && same_shape(dy, y) && same_shape(dy, dx) &&
                     dy->dtype == y->dtype && dy->dtype == dx->dtype, "relu bwd: bad tensors");
  DISPATCH_T(dy->dtype, {
    const int64_t total = pixels(dy) * (dy->c / Vec<T>::N);
    basi::launch(relu_bwd_kernel<T>, grid_for(total, 256), 256, 0, (cudaStream_t)stream, (const T*)dy->ptr, dy->ld,
                 (const T*)y->ptr, y->ld, (T*)dx->ptr, dx->ld, accumulate, dy->c, total);
  })
  BASI_CHECK_LAUNCH("relu_bwd");
  return BASI_OK;
}

int basi_onehot2_f32(const float* labels, float* out, int64_t n, void* stream) {
  BASI_CHECK_ARG(labels && out && n > 0 && (((uintptr_t)out) & 7) == 0, "onehot2: bad argument");
  basi::launch(onehot2_kernel, grid_for(n, 256), 256, 0, (cudaStream_t)stream, labels, (float2*)out, n);
  BASI_CHECK_LAUNCH("onehot2");
  return BASI_OK;
}

int basi_maxpool2s2_fwd(const basi_tensor* x, const basi_tensor* y, uint8_t* argmax, void* stream) {
  BASI_CHECK_ARG(x && y && argmax && vec_ok(x) && vec_ok(y) && x->dtype == y->dtype && x->c == y->c && x->n == y->n &&
                     y->h == x->h / 2 && y->w == x->w / 2 && y->h > 0 && y->w > 0, "maxpool2s2 fwd: bad tensors");
  DISPATCH_T(x->dtype, {
    int64_t total = pixels(y) * (y->c / Vec<T>::N);
    basi::launch(maxpool2s2_fwd_kernel<T>, grid_for(total, 256), 256, 0, (cudaStream_t)stream, (const T*)x->ptr, x->ld,
                 x->h, x->w, x->c, (T*)y->ptr, y->ld, y->h, y->w, argmax, total);
  })
  BASI_CHECK_LAUNCH("maxpool2s2_fwd");
  return BASI_OK;
}

int basi_maxpool2s2_bwd(const basi_tensor* dy, const uint8_t* argmax, const basi_tensor* dx, int accumulate,
                        void* stream) {
  BASI_CHECK_ARG(dy && dx && argmax && vec_ok(dy) && vec_ok(dx) && dx->dtype == dy->dtype && dx->c == dy->c &&
                     dx->n == dy->n && dy->h == dx->h / 2 && dy->w == dx->w / 2, "maxpool2s2 bwd: bad tensors");
  DISPATCH_T(dx->dtype, {
    int64_t total = pixels(dx) * (dx->c / Vec<T>::N);
    basi::launch(maxpool2s2_bwd_kernel<T>, grid_for(total, 256), 256, 0, (cudaStream_t)stream, (const T*)dy->ptr, dy->ld,
                 dy->h, dy->w, argmax, (T*)dx->ptr, dx->ld, dx->h, dx->w, dx->c, accumulate, total);
  })
  BASI_CHECK_LAUNCH("maxpool2s2_bwd");
  return BASI_OK;
}

int basi_maxpool3s2_fwd(const basi_tensor* x, const basi_tensor* y, uint8_t* argmax, void* stream) {
  BASI_CHECK_ARG(x && y && argmax && vec_ok(x) && vec_ok(y) && x->dtype == y->dtype && x->c == y->c && x->n == y->n,
                 "maxpool fwd: bad tensors");
  BASI_CHECK_ARG(y->h == (x->h + 1) / 2 && y->w == (x->w + 1) / 2, "maxpool fwd: output must be ceil(in/2)");
  int pt, pl;
  same_pad_3s2(x->h, y->h, &pt);
  same_pad_3s2(x->w, y->w, &pl);
  BASI_CHECK_ARG(pixels(x) * (int64_t)x->c < ((int64_t)1 << 31), "maxpool fwd: tensor too large for 32-bit indexing");
  DISPATCH_T(x->dtype, {
    int64_t total = pixels(y) * (y->c / Vec<T>::N);
    basi::launch(maxpool3s2_fwd_kernel<T>, grid_for(total, 256), 256, 0, (cudaStream_t)stream, (const T*)x->ptr, x->ld, x->h, x->w, x->c, (T*)y->ptr, y->ld, y->h, y->w, pt, pl, argmax, total);
  })
  BASI_CHECK_LAUNCH("maxpool3s2_fwd");
  return BASI_OK;
}

int basi_maxpool3s2_bwd(const basi_tensor* dy, const uint8_t* argmax, const basi_tensor* dx, int accumulate,
                        void* stream) {
  BASI_CHECK_ARG(dy && dx && argmax && vec_ok(dy) && vec_ok(dx) && dx->dtype == dy->dtype && dx->c == dy->c &&
                     dx->n == dy->n && dy->h == (dx->h + 1) / 2 && dy->w == (dx->w + 1) / 2,
                 "maxpool bwd: bad tensors");
  int pt, pl;
  same_pad_3s2(dx->h, dy->h, &pt);
  same_pad_3s2(dx->w, dy->w, &pl);
  BASI_CHECK_ARG(pixels(dx) * (int64_t)dx->c < ((int64_t)1 << 31), "maxpool bwd: tensor too large for 32-bit indexing");
  DISPATCH_T(dx->dtype, {
    int64_t total = pixels(dx) * (dx->c / Vec<T>::N);
    basi::launch(maxpool3s2_bwd_kernel<T>, grid_for(total, 256), 256, 0, (cudaStream_t)stream, (const T*)dy->ptr, dy->ld, dy->h, dy->w, argmax, (T*)dx->ptr, dx->ld, dx->h, dx->w, dx->c, pt, pl, accumulate,
        total);
  })
  BASI_CHECK_LAUNCH("maxpool3s2_bwd");
  return BASI_OK;
}

// (bx channel groups) x (by = 256/bx window threads, a power of two for the tree reduction); grid.y chunks
static void pool_block(int cgs, int VN, dim3* block, size_t* smem, unsigned* chunks, int64_t out_pixels = 1 << 30) {
  int bx = 1;
  while (bx * 2 <= 8 && bx * 2 <= cgs) bx *= 2;
  // few output pixels (the 1x1 .. 6x6 pyramid maps): narrower channel slices per block so the grid still fills the GPU
  while (bx > 1 && out_pixels * ((cgs + bx - 1) / bx) < 4 * (int64_t)sm_count()) bx /= 2;
  int by = 256 / bx;
  *block = dim3(bx, by);
  *smem = (size_t)bx * by * VN * sizeof(float);
  *chunks = (unsigned)((cgs + bx - 1) / bx);
}

int basi_avgpool_fwd(const basi_tensor* x, int k, const basi_tensor* y, void* stream) {
  BASI_CHECK_ARG(x && y && k > 0 && vec_ok(x) && vec_ok(y) && x->dtype == y->dtype && x->c == y->c && x->n == y->n &&
                     y->h == x->h / k && y->w == x->w / k && y->h > 0 && y->w > 0,
                 "avgpool fwd: bad tensors");
  DISPATCH_T(x->dtype, {
    dim3 block;
    size_t smem;
    unsigned chunks;
    pool_block(x->c / Vec<T>::N, Vec<T>::N, &block, &smem, &chunks);
    basi::launch(avgpool_fwd_kernel<T>, dim3((unsigned)pixels(y), chunks), block, smem, (cudaStream_t)stream, (const T*)x->ptr, x->ld, x->h, x->w, x->c, k, (T*)y->ptr, y->ld, y->h, y->w);
  })
  BASI_CHECK_LAUNCH("avgpool_fwd");
  return BASI_OK;
}

int basi_avgpool_bwd(const basi_tensor* dy, int k, const basi_tensor* dx, int accumulate, void* stream) {
  BASI_CHECK_ARG(dy && dx && k > 0 && vec_ok(dy) && vec_ok(dx) && dx->dtype == dy->dtype && dx->c == dy->c &&
                     dx->n == dy->n && dy->h == dx->h / k && dy->w == dx->w / k,
                 "avgpool bwd: bad tensors");
  DISPATCH_T(dx->dtype, {
    int64_t total = pixels(dx) * (dx->c / Vec<T>::N);
    basi::launch(avgpool_bwd_kernel<T>, grid_for(total, 256), 256, 0, (cudaStream_t)stream, (const T*)dy->ptr, dy->ld, dy->h, dy->w, k, (T*)dx->ptr, dx->ld, dx->h, dx->w, dx->c, accumulate, total);
  })
  BASI_CHECK_LAUNCH("avgpool_bwd");
  return BASI_OK;
}

static int fill_multipool(MultiPool* mp, const basi_tensor* x, int n_pools, const int* ks, const basi_tensor* const* ys,
                          const char* who) {
  BASI_CHECK_ARG(x && ks && ys && n_pools >= 1 && n_pools <= MP_MAX && vec_ok(x), "%s: bad argument", who);
  memset(mp, 0, sizeof(*mp));
  mp->np = n_pools;
  int cells = 0;
  for (int p = 0; p < n_pools; ++p) {
    const basi_tensor* y = ys[p];
    BASI_CHECK_ARG(y && ks[p] > 0 && vec_ok(y) && y->dtype == x->dtype && y->c == x->c && y->n == x->n &&
                       y->h == x->h / ks[p] && y->w == x->w / ks[p] && y->h > 0 && y->w > 0,
                   "%s: pool %d does not match its input", who, p);
    mp->k[p] = ks[p]; mp->OH[p] = y->h; mp->OW[p] = y->w; mp->cell0[p] = cells;
    mp->y[p] = y->ptr; mp->ldy[p] = y->ld;
    cells += y->h * y->w;
  }
  mp->cells = cells;
  return BASI_OK;
}

int64_t basi_avgpool_multi_scratch_floats(const basi_tensor* x, int n_pools, const int* ks) {
  if (!x || !ks || n_pools < 1 || n_pools > MP_MAX) return -1;
  int64_t rcells = 0;
  for (int p = 0; p < n_pools; ++p) {
    if (ks[p] < 1 || x->h / ks[p] < 1 || x->w / ks[p] < 1) return -1;
    rcells += x->w / ks[p];
  }
  return (int64_t)x->n * x->h * rcells * x->c;     // [n][h][row cells][c]
}

int basi_avgpool_multi_fwd(const basi_tensor* x, int n_pools, const int* ks, const basi_tensor* const* ys,
                           float* scratch, void* stream) {
  MultiPool mp;
  int rc = fill_multipool(&mp, x, n_pools, ks, ys, "avgpool_multi fwd");
  if (rc) return rc;
  BASI_CHECK_ARG(scratch, "avgpool_multi fwd: null scratch");
  cudaStream_t st = (cudaStream_t)stream;
  int rcells = 0;
  for (int p = 0; p < n_pools; ++p) rcells += mp.OW[p];
  DISPATCH_T(x->dtype, {
    const int cgs = x->c / Vec<T>::N;
    int64_t total = (int64_t)x->n * x->h * cgs;
    basi::launch(avgpool_multi_rows_kernel<T>, grid_for(total, 256), 256, 0, st, (const T*)x->ptr, x->ld, x->h, x->w,
                 x->c, mp, rcells, scratch, total);
    total = (int64_t)x->n * mp.cells * cgs;
    basi::launch(avgpool_multi_cells_kernel<T>, grid_for(total, 256), 256, 0, st, mp, x->h, x->c, rcells,
                 (const float*)scratch, total);
  })
  BASI_CHECK_LAUNCH("avgpool_multi_fwd");
  return BASI_OK;
}

int basi_avgpool_multi_bwd(const basi_tensor* const* dys, int n_pools, const int* ks, const basi_tensor* dx,
                           int accumulate, void* stream) {
  MultiPool mp;
  int rc = fill_multipool(&mp, dx, n_pools, ks, dys, "avgpool_multi bwd");
  if (rc) return rc;
  DISPATCH_T(dx->dtype, {
    int threads = dx->c / Vec<T>::N;
    threads = threads >= 256 ? 256 : (threads < 32 ? 32 : (threads + 31) / 32 * 32);
    basi::launch(avgpool_multi_bwd_kernel<T>, dim3((unsigned)(dx->n * dx->h)), dim3(threads), 0, (cudaStream_t)stream, mp,
                 (T*)dx->ptr, dx->ld, dx->h, dx->w, dx->c, accumulate);
  })
  BASI_CHECK_LAUNCH("avgpool_multi_bwd");
  return BASI_OK;
}

static float ac_scale(int in, int out) { return out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f; }

int basi_bilinear_ac_fwd(const basi_tensor* x, const basi_tensor* y, void* stream) {
  BASI_CHECK_ARG(x && y && vec_ok(x) && vec_ok(y) && x->dtype == y->dtype && x->c == y->c && x->n == y->n,
                 "bilinear fwd: bad tensors");
  DISPATCH_T(x->dtype, {
    int64_t total = pixels(y) * (y->c / Vec<T>::N);
    basi::launch(bilinear_ac_fwd_kernel<T>, grid_for(total, 256), 256, 0, (cudaStream_t)stream, (const T*)x->ptr, x->ld, x->h, x->w, x->c, (T*)y->ptr, y->ld, y->h, y->w, ac_scale(x->h, y->h),
        ac_scale(x->w, y->w), total);
  })
  BASI_CHECK_LAUNCH("bilinear_ac_fwd");
  return BASI_OK;
}

int basi_bilinear_ac_bwd(const basi_tensor* dy, const basi_tensor* dx, int accumulate, void* stream) {
  BASI_CHECK_ARG(dy && dx && vec_ok(dy) && vec_ok(dx) && dx->dtype == dy->dtype && dx->c == dy->c && dx->n == dy->n,
                 "bilinear bwd: bad tensors");
  DISPATCH_T(dx->dtype, {
    dim3 block;
    size_t smem;
    unsigned chunks;
    pool_block(dx->c / Vec<T>::N, Vec<T>::N, &block, &smem, &chunks, pixels(dx));
    basi::launch(bilinear_ac_bwd_kernel<T>, dim3((unsigned)pixels(dx), chunks), block, smem, (cudaStream_t)stream, (const T*)dy->ptr, dy->ld, dy->h, dy->w, (T*)dx->ptr, dx->ld, dx->h, dx->w, dx->c, ac_scale(dx->h, dy->h),
        ac_scale(dx->w, dy->w), accumulate);
  })
  BASI_CHECK_LAUNCH("bilinear_ac_bwd");
  return BASI_OK;
}

int basi_gate_mul_fwd(const basi_tensor* feat, const float* logits, int nseg, int att, const basi_tensor* out,
                      void* stream) {
  BASI_CHECK_ARG(feat && logits && out && vec_ok(feat) && vec_ok(out) && same_shape(feat, out) &&
                     feat->dtype == out->dtype && att >= 0 && att < nseg,
                 "gate_mul fwd: bad argument");
  int64_t R = pixels(feat);
  DISPATCH_T(feat->dtype, {
    int64_t total = R * (feat->c / Vec<T>::N);
    basi::launch(gate_mul_fwd_kernel<T, true>, grid_for(total, 256), 256, 0, (cudaStream_t)stream, (const T*)feat->ptr, feat->ld, logits, nseg, att, (T*)out->ptr, out->ld, R, feat->c);
  })
  BASI_CHECK_LAUNCH("gate_mul_fwd");
  return BASI_OK;
}

int basi_gate_mul_bwd(const basi_tensor* dout, const basi_tensor* feat, const float* logits, int nseg, int att,
                      const basi_tensor* dfeat, int dfeat_accumulate, float* dlogits, void* stream) {
  BASI_CHECK_ARG(dout && feat && logits && dfeat && dlogits && vec_ok(dout) && vec_ok(feat) && vec_ok(dfeat) &&
                     same_shape(dout, feat) && same_shape(dfeat, feat) && dout->dtype == feat->dtype &&
                     dfeat->dtype == feat->dtype && att >= 0 && att < nseg,
                 "gate_mul bwd: bad argument");
  int64_t R = pixels(feat);
  DISPATCH_T(feat->dtype, {
    basi::launch(gate_mul_bwd_kernel<T, true>, grid_for(R * 32, 256), 256, 0, (cudaStream_t)stream, (const T*)dout->ptr, dout->ld, (const T*)feat->ptr, feat->ld, logits, nseg, att, (T*)dfeat->ptr, dfeat->ld,
        dfeat_accumulate, dlogits, R, feat->c);
  })
  BASI_CHECK_LAUNCH("gate_mul_bwd");
  return BASI_OK;
}

/* ---- A12: click / attention map times features, broadcast over channels (no ReLU) ---- */
int basi_mask_mul_fwd(const basi_tensor* feat, const float* mask, int nch, int ch, const basi_tensor* out,
                      void* stream) {
  BASI_CHECK_ARG(feat && mask && out && vec_ok(feat) && vec_ok(out) && same_shape(feat, out) &&
                     feat->dtype == out->dtype && ch >= 0 && ch < nch,
                 "mask_mul fwd: bad argument");
  int64_t R = pixels(feat);
  DISPATCH_T(feat->dtype, {
    int64_t total = R * (feat->c / Vec<T>::N);
    basi::launch(gate_mul_fwd_kernel<T, false>, grid_for(total, 256), 256, 0, (cudaStream_t)stream, (const T*)feat->ptr, feat->ld, mask, nch, ch, (T*)out->ptr, out->ld, R, feat->c);
  })
  BASI_CHECK_LAUNCH("mask_mul_fwd");
  return BASI_OK;
}

int basi_mask_mul_bwd(const basi_tensor* dout, const basi_tensor* feat, const float* mask, int nch, int ch,
                      const basi_tensor* dfeat, int dfeat_accumulate, float* dmask, void* stream) {
  BASI_CHECK_ARG(dout && feat && mask && dfeat && vec_ok(dout) && vec_ok(feat) && vec_ok(dfeat) &&
                     same_shape(dout, feat) && same_shape(dfeat, feat) && dout->dtype == feat->dtype &&
                     dfeat->dtype == feat->dtype && ch >= 0 && ch < nch,
                 "mask_mul bwd: bad argument");
  int64_t R = pixels(feat);
  DISPATCH_T(feat->dtype, {
    basi::launch(gate_mul_bwd_kernel<T, false>, grid_for(R * 32, 256), 256, 0, (cudaStream_t)stream, (const T*)dout->ptr, dout->ld, (const T*)feat->ptr, feat->ld, mask, nch, ch, (T*)dfeat->ptr, dfeat->ld,
        dfeat_accumulate, dmask, R, feat->c);
  })
  BASI_CHECK_LAUNCH("mask_mul_bwd");
  return BASI_OK;
}

int basi_softmax_gate_fwd(const float* logits, int64_t rows, int C, int sel, float thr, float* gate, void* stream) {
  BASI_CHECK_ARG(logits && gate && rows > 0 && C > 0 && sel >= 0 && sel < C, "softmax_gate fwd: bad argument");
  basi::launch(softmax_gate_fwd_kernel, grid_for(rows, 128), 128, 0, (cudaStream_t)stream, logits, rows, C, sel, thr, gate);
  BASI_CHECK_LAUNCH("softmax_gate_fwd");
  return BASI_OK;
}

int basi_softmax_gate_bwd(const float* logits, const float* dgate, int64_t rows, int C, int sel, float thr,
                          float* dlogits, int accumulate, void* stream) {
  BASI_CHECK_ARG(logits && dgate && dlogits && rows > 0 && C > 0 && sel >= 0 && sel < C,
                 "softmax_gate bwd: bad argument");
  basi::launch(softmax_gate_bwd_kernel, grid_for(rows, 128), 128, 0, (cudaStream_t)stream, logits, dgate, rows, C, sel, thr,
                                                                               dlogits, accumulate);
  BASI_CHECK_LAUNCH("softmax_gate_bwd");
  return BASI_OK;
}

int basi_resize_nearest_fwd(const basi_tensor* x, const basi_tensor* y, void* stream) {
  BASI_CHECK_ARG(x && y && x->ptr && y->ptr && x->dtype == y->dtype && x->c == y->c && x->n == y->n,
                 "resize_nearest fwd: bad tensors");
  DISPATCH_T(x->dtype, {
    int64_t total = pixels(y) * y->c;
    basi::launch(resize_nearest_fwd_kernel<T>, grid_for(total, 256), 256, 0, (cudaStream_t)stream, (const T*)x->ptr, x->ld, x->h, x->w, x->c, (T*)y->ptr, y->ld, y->h, y->w, (float)x->h / (float)y->h,
        (float)x->w / (float)y->w, total);
  })
  BASI_CHECK_LAUNCH("resize_nearest_fwd");
  return BASI_OK;
}

int basi_resize_nearest_bwd(const basi_tensor* dy, const basi_tensor* dx, int accumulate, void* stream) {
  BASI_CHECK_ARG(dy && dx && dy->ptr && dx->ptr && dx->dtype == dy->dtype && dx->c == dy->c && dx->n == dy->n,
                 "resize_nearest bwd: bad tensors");
  DISPATCH_T(dx->dtype, {
    int64_t total = pixels(dx) * dx->c;
    basi::launch(resize_nearest_bwd_kernel<T>, grid_for(total, 256), 256, 0, (cudaStream_t)stream, (const T*)dy->ptr, dy->ld, dy->h, dy->w, (T*)dx->ptr, dx->ld, dx->h, dx->w, dx->c, (float)dx->h / (float)dy->h,
        (float)dx->w / (float)dy->w, accumulate, total);
  })
  BASI_CHECK_LAUNCH("resize_nearest_bwd");
  return BASI_OK;
}

int basi_subsample_fwd(const basi_tensor* x, int stride, const basi_tensor* y, void* stream) {
  BASI_CHECK_ARG(x && y && stride >= 1 && vec_ok(x) && vec_ok(y) && x->dtype == y->dtype && x->c == y->c &&
                     x->n == y->n && y->h == (x->h + stride - 1) / stride && y->w == (x->w + stride - 1) / stride,
                 "subsample fwd: bad tensors");
  DISPATCH_T(x->dtype, {
    int64_t total = pixels(y) * (y->c / Vec<T>::N);
    basi::launch(subsample_fwd_kernel<T>, grid_for(total, 256), 256, 0, (cudaStream_t)stream, (const T*)x->ptr, x->ld,
                 x->h, x->w, x->c, stride, (T*)y->ptr, y->ld, y->h, y->w, total);
  })
  BASI_CHECK_LAUNCH("subsample_fwd");
  return BASI_OK;
}

int basi_subsample_bwd(const basi_tensor* dy, int stride, const basi_tensor* dx, int accumulate, void* stream) {
  BASI_CHECK_ARG(dy && dx && stride >= 1 && vec_ok(dy) && vec_ok(dx) && dx->dtype == dy->dtype && dx->c == dy->c &&
                     dx->n == dy->n && dy->h == (dx->h + stride - 1) / stride && dy->w == (dx->w + stride - 1) / stride,
                 "subsample bwd: bad tensors");
  DISPATCH_T(dx->dtype, {
    int64_t total = pixels(dx) * (dx->c / Vec<T>::N);
    basi::launch(subsample_bwd_kernel<T>, grid_for(total, 256), 256, 0, (cudaStream_t)stream, (const T*)dy->ptr, dy->ld,
                 dy->h, dy->w, stride, (T*)dx->ptr, dx->ld, dx->h, dx->w, dx->c, accumulate, total);
  })
  BASI_CHECK_LAUNCH("subsample_bwd");
  return BASI_OK;
}

int basi_wbce_fwd_bwd(const float* logits, const float* labels, float pos_weight, double scale, float grad_scale,
                      int64_t n, double* loss_acc, float* dlogits, void* stream) {
  BASI_CHECK_ARG(logits && labels && loss_acc && n > 0, "wbce: bad argument");
  basi::launch(wbce_kernel, grid_for((n + 3) / 4, 256, 8), 256, 0, (cudaStream_t)stream, logits, labels, pos_weight, scale, grad_scale, n,
                                                                    loss_acc, dlogits);
  BASI_CHECK_LAUNCH("wbce_fwd_bwd");
  return BASI_OK;
}

int basi_sigmoid_fwd(const float* x, float* y, int64_t n, void* stream) {
  BASI_CHECK_ARG(x && y && n > 0, "sigmoid_fwd: bad argument");
  basi::launch(sigmoid_fwd_kernel, grid_for(n, 256), 256, 0, (cudaStream_t)stream, x, y, n);
  BASI_CHECK_LAUNCH("sigmoid_fwd");
  return BASI_OK;
}

int basi_sigmoid_bwd(const float* dy, const float* y, float* dx, int64_t n, int accumulate, void* stream) {
  BASI_CHECK_ARG(dy && y && dx && n > 0, "sigmoid_bwd: bad argument");
  basi::launch(sigmoid_bwd_kernel, grid_for(n, 256), 256, 0, (cudaStream_t)stream, dy, y, dx, n, accumulate);
  BASI_CHECK_LAUNCH("sigmoid_bwd");
  return BASI_OK;
}

int basi_wbce_sel_fwd_bwd(const float* logits, int C, int sel, const float* labels, float pos_weight, double scale,
                          float grad_scale, int64_t rows, double* loss_acc, float* dlogits, void* stream) {
  BASI_CHECK_ARG(logits && labels && loss_acc && rows > 0 && C > 0 && sel >= 0 && sel < C, "wbce_sel: bad argument");
  basi::launch(wbce_sel_kernel, grid_for(rows, 256, 4), 256, 0, (cudaStream_t)stream, logits, C, sel, labels, pos_weight,
               scale, grad_scale, rows, loss_acc, dlogits);
  BASI_CHECK_LAUNCH("wbce_sel_fwd_bwd");
  return BASI_OK;
}

int basi_softmax_ce_fwd_bwd(const float* logits, const int32_t* labels, int64_t rows, int C, double scale,
                            float grad_scale, double* loss_acc, float* dlogits, void* stream) {
  BASI_CHECK_ARG(logits && labels && loss_acc && rows > 0 && C > 0, "softmax_ce: bad argument");
  basi::launch(softmax_ce_kernel, grid_for(rows, 128, 4), 128, 0, (cudaStream_t)stream, logits, labels, rows, C, scale,
                                                                             grad_scale, loss_acc, dlogits);
  BASI_CHECK_LAUNCH("softmax_ce_fwd_bwd");
  return BASI_OK;
}

int basi_sgd_step(float* w, const float* g, const float* lr_dev, int64_t n, void* w_bf16, void* stream) {
  BASI_CHECK_ARG(w && g && lr_dev && n > 0 && ((uintptr_t)w & 15) == 0 && ((uintptr_t)g & 15) == 0 &&
                     ((uintptr_t)w_bf16 & 7) == 0,
                 "sgd_step: bad argument / alignment");
  basi::launch(sgd_kernel, grid_for(n / 4 + 1, 256), 256, 0, (cudaStream_t)stream, w, g, lr_dev, n, (bf16*)w_bf16);
  BASI_CHECK_LAUNCH("sgd_step");
  return BASI_OK;
}

int basi_threshold(const float* logits, float thr, int32_t* out, int64_t n, void* stream) {
  BASI_CHECK_ARG(logits && out && n > 0, "threshold: bad argument");
  basi::launch(threshold_kernel, grid_for(n, 256), 256, 0, (cudaStream_t)stream, logits, thr, out, n);
  BASI_CHECK_LAUNCH("threshold");
  return BASI_OK;
}

int basi_argmax(const float* logits, int64_t rows, int C, int32_t* out, void* stream) {
  BASI_CHECK_ARG(logits && out && rows > 0 && C > 0, "argmax: bad argument");
  basi::launch(argmax_kernel, grid_for(rows, 128), 128, 0, (cudaStream_t)stream, logits, rows, C, out);
  BASI_CHECK_LAUNCH("argmax");
  return BASI_OK;
}

int basi_upsample_legacy_argmax(const float* logits, int B, int P_h, int P_w, int C, int S_h, int S_w, int32_t* out,
                                void* stream) {
  BASI_CHECK_ARG(logits && out && B > 0 && P_h > 0 && P_w > 0 && C > 0 && S_h > 0 && S_w > 0,
                 "upsample_legacy_argmax: bad argument");
  int64_t total = (int64_t)B * S_h * S_w;
  basi::launch(upsample_legacy_argmax_kernel, grid_for(total, 256), 256, 0, (cudaStream_t)stream, logits, B, P_h, P_w, C, S_h, S_w, (float)P_h / (float)S_h, (float)P_w / (float)S_w, out);
  BASI_CHECK_LAUNCH("upsample_legacy_argmax");
  return BASI_OK;
}

int basi_split3_bf16(const basi_tensor* x, const basi_tensor* y, void* stream) {
  BASI_CHECK_ARG(x && y && x->ptr && y->ptr, "split3_bf16: null argument");
  BASI_CHECK_ARG(x->dtype == BASI_F32 && y->dtype == BASI_BF16 && (y->c == 3 * x->c || y->c == 2 * x->c) &&
                     x->n == y->n && x->h == y->h && x->w == y->w,
                 "split3_bf16: y must be a bf16 tensor with 2x or 3x the channels of the float32 x");
  BASI_CHECK_ARG(x->c % 4 == 0 && x->ld % 4 == 0 && y->ld % 4 == 0 && (((uintptr_t)x->ptr | (uintptr_t)y->ptr) & 15) == 0,
                 "split3_bf16: channels / strides must be multiples of 4 and pointers 16-byte aligned");
  const int64_t total = pixels(x) * (x->c / 4);
#if BASI_H16_FP16
  set_error("split3_bf16: the split-operand (f32) mode lives in the bfloat16 build of the library");
  return BASI_E_INVALID;
#endif
  basi::launch(split3_kernel, grid_for(total, 256), 256, 0, (cudaStream_t)stream, (const float*)x->ptr, x->ld, x->c,
               (__nv_bfloat16*)y->ptr, y->ld, total, y->c / x->c);
  BASI_CHECK_LAUNCH("split3_bf16");
  return BASI_OK;
}
int basi_cast_f32_to_bf16(const float* src, void* dst, int64_t n, void* stream) {
  BASI_CHECK_ARG(src && dst && n > 0, "cast: bad argument");
  basi::launch(cast_f32_bf16_kernel, grid_for(n, 256), 256, 0, (cudaStream_t)stream, src, (bf16*)dst, n);
  BASI_CHECK_LAUNCH("cast_f32_to_bf16");
  return BASI_OK;
}
int basi_cast_bf16_to_f32(const void* src, float* dst, int64_t n, void* stream) {
  BASI_CHECK_ARG(src && dst && n > 0, "cast: bad argument");
  basi::launch(cast_bf16_f32_kernel, grid_for(n, 256), 256, 0, (cudaStream_t)stream, (const bf16*)src, dst, n);
  BASI_CHECK_LAUNCH("cast_bf16_to_f32");
  return BASI_OK;
}
int basi_relu_bwd_f32(float* dy, const float* y, int64_t n, void* stream) {
  BASI_CHECK_ARG(dy && y && n > 0, "relu_bwd: bad argument");
  basi::launch(relu_bwd_f32_kernel, grid_for(n, 256), 256, 0, (cudaStream_t)stream, dy, y, n);
  BASI_CHECK_LAUNCH("relu_bwd_f32");
  return BASI_OK;
}

}  // extern "C"
