// Experiment (round-1 hand-off): can a tcgen05 K-major SWIZZLE_128B operand START AT AN ARBITRARY ROW of a TMA-written
// shared-memory tile?  If yes, a 3x3 convolution can load one halo tile per 64-channel chunk and feed all nine taps
// from it (start = base + (dr * pitch + dc) * 128 B, SBO = pitch * 128 B), instead of re-fetching the activation box
// once per tap through L2.
//
// Test: A_full = [256 rows][64 bf16] (row i, col c) = small integers, TMA-loaded with SWIZZLE_128B into 32 KB of
// shared memory.  B = [64 n][64 k] identity.  D = A_full[shift .. shift+127] * B^T must equal rows shift..shift+127.
// The UMMA A descriptor start address is base + shift * 128 B, LBO unused, SBO = 1024 B, base_offset field = `bo`.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o umma_row_shift umma_row_shift.cu -lcuda
//   ./umma_row_shift            (prints PASS/FAIL per shift for base_offset = 0 and base_offset = (shift & 7))
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  long long t0 = clock64();
  while (!mbar_try(bar, parity))
    if (clock64() - t0 > 4000000000LL) { printf("timeout\n"); __trap(); }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t base_off) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_off & 7) << 49;
  d |= (uint64_t)2 << 61;
  return d;
}

__global__ void __launch_bounds__(128, 1)
shift_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, float* out, int shift,
             int bo, int sbo) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                 // 256 rows x 128 B = 32 KB
  uint8_t* sB = smem + 32768;         // 64 rows x 128 B = 8 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 32768 + 8192);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 2);
  const uint32_t full = smem_u32(bars), done = full + 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(full, 1);
    mbar_init(done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(64) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(full, 32768 + 8192);
    tma_load_2d(smem_u32(sA), &mapA, full, 0, 0);
    tma_load_2d(smem_u32(sB), &mapB, full, 0, 0);
    mbar_wait(full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // idesc: D f32, A/B bf16, K-major both, N = 64, M = 128
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t adesc = make_desc(smem_u32(sA) + shift * 128, 16, sbo, bo);
    const uint64_t bdesc = make_desc(smem_u32(sB), 16, 1024, 0);
    for (int k = 0; k < 4; ++k) {
      uint32_t acc = k != 0;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                   ::"r"(tmem), "l"(adesc + 2 * k), "l"(bdesc + 2 * k), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(done) : "memory");
  }
  mbar_wait(done, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // each warp reads its 32 TMEM lanes (rows), 64 columns
  for (int c = 0; c < 2; ++c) {
    uint32_t r[32];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c * 32;
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
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 64 + c * 32 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64) : "memory");
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static void make_map(EncodeFn enc, CUtensorMap* m, void* ptr, int rows) {
  cuuint64_t dims[2] = {64, (cuuint64_t)rows};
  cuuint64_t strides[1] = {128};
  cuuint32_t box[2] = {64, (cuuint32_t)rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
}

int main() {
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  EncodeFn enc = (EncodeFn)fp;
  std::vector<bf16> hA(256 * 64), hB(64 * 64);
  for (int i = 0; i < 256; ++i)
    for (int c = 0; c < 64; ++c) hA[i * 64 + c] = __float2bfloat16((float)((i * 3 + c * 5) % 251));   // exact in bf16
  for (int n = 0; n < 64; ++n)
    for (int k = 0; k < 64; ++k) hB[n * 64 + k] = __float2bfloat16(n == k ? 1.f : 0.f);
  bf16 *dA, *dB;
  float* dO;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dO, 128 * 64 * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap mA, mB;
  make_map(enc, &mA, dA, 256);
  make_map(enc, &mB, dB, 64);
  cudaFuncSetAttribute(shift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  std::vector<float> hO(128 * 64);
  // (shift rows, SBO): SBO 1024 = rows of a group and groups back to back; SBO 2048 = groups 16 rows apart (halo pitch 16)
  const int cases[][2] = {{0, 1024}, {1, 1024}, {3, 1024}, {7, 1024}, {8, 1024}, {9, 1024}, {17, 1024},
                          {0, 2048}, {1, 2048}, {5, 2048}, {8, 2048}, {18, 2048}};
  for (auto& cs : cases) {
    for (int mode = 0; mode < 2; ++mode) {
      const int shift = cs[0], sbo = cs[1], bo = mode ? (shift & 7) : 0;
      cudaMemset(dO, 0, 128 * 64 * 4);
      shift_kernel<<<1, 128, 64 * 1024>>>(mA, mB, dO, shift, bo, sbo);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("shift %d sbo %d bo %d: CUDA error %s\n", shift, sbo, bo, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(hO.data(), dO, hO.size() * 4, cudaMemcpyDeviceToHost);
      int bad = 0, first = -1;
      for (int j = 0; j < 128; ++j) {
        // MMA row j = group j/8, row j%8: smem row = shift + (j/8) * (sbo/128) + j%8
        const int src = shift + (j / 8) * (sbo / 128) + (j % 8);
        if (src >= 256) continue;
        for (int c = 0; c < 64; ++c) {
          const float want = (float)((src * 3 + c * 5) % 251);
          if (hO[j * 64 + c] != want) { ++bad; if (first < 0) first = j * 64 + c; }
        }
      }
      printf("shift %2d  SBO %4d  base_offset %d : %s", shift, sbo, bo, bad ? "FAIL" : "PASS");
      if (bad) printf("  (%d wrong, first at row %d col %d: got %.0f)", bad, first / 64, first % 64, hO[first]);
      printf("\n");
    }
  }
  return 0;
}
