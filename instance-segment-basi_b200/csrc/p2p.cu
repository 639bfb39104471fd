// Gradient all-reduce over NVLink peer memory for the data-parallel step.
//
// The persistent tcgen05 / cooperative BN kernels of the backward pass occupy every SM, so NCCL's CTAs only run in
// the gaps and the bucketed all-reduce ends up serialised behind the backward pass (measured: 0.77 ms exposed of a
// 9.1 ms step at 2 GPUs, 0.93 ms at 8).  This is the full-machine alternative: every rank keeps
// its flat gradient in a symmetric buffer that all peers have mapped (torch symmetric memory = cuMem handles exchanged
// at start-up), and ONE kernel per rank does a two-shot all-reduce with plain peer loads and stores:
//
//   barrier (flags in peer memory)  ->  rank r averages slice r of all ranks and writes it into slice r of every
//   rank's buffer  ->  barrier
//
// Each rank moves (W-1)/W of the buffer in and out over NVLink: 2 x 103 MB at W = 8 for the 118 MB of cfg3, i.e.
// about 0.3 ms at the measured 770 GB/s per direction, with all SMs pulling.
#include "common.cuh"

namespace basi {

constexpr int P2P_MAX_WORLD = 8;
struct P2PArgs {
  float* bufs[P2P_MAX_WORLD];          // bufs[q]: rank q's symmetric gradient buffer as mapped into THIS process
  unsigned int* flags[P2P_MAX_WORLD];  // flags[q]: rank q's flag array [world] (same mapping)
  int rank, world;
};

// One block of 32 threads.  seq is a device counter that every rank advances identically (graph-replay safe).
// Thread q signals peer q (flags[q][rank] = seq) and waits for peer q's signal in its own array.
__global__ void p2p_barrier_kernel(const P2PArgs a, unsigned int* __restrict__ seq_dev) {
  __shared__ unsigned int s_seq;
  if (threadIdx.x == 0) {
    s_seq = *seq_dev + 1u;
    *seq_dev = s_seq;
  }
  __syncthreads();
  __threadfence_system();                      // everything this GPU wrote before is visible to the peers first
  const int q = threadIdx.x;
  if (q < a.world) {
    const unsigned int seq = s_seq;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(a.flags[q] + a.rank), "r"(seq) : "memory");
    const unsigned int* mine = a.flags[a.rank] + q;
    const long long t0 = clock64();
    for (;;) {
      unsigned int v;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
      if ((int)(v - seq) >= 0) break;
      if (clock64() - t0 > 20000000000LL) {     // ~10 s: a missing peer must surface as an error, not as a hang
        printf("basi: p2p barrier timed out (rank %d waits for rank %d, flag %u, seq %u)\n", a.rank, q, v, seq);
        __trap();
      }
    }
  }
  __syncthreads();
  __threadfence_system();
}

template <int W>
__global__ void __launch_bounds__(256) p2p_allreduce_mean_kernel(const P2PArgs a, int64_t n4) {
  const int64_t per = (n4 + W - 1) / W;
  const int64_t lo = (int64_t)a.rank * per;
  const int64_t hi = lo + per < n4 ? lo + per : n4;
  const float inv = 1.0f / (float)W;
  // U independent 16-byte elements per thread and iteration: W x U peer loads in flight per thread (an NVLink round
  // trip is a few microseconds; at W = 2 one load per thread does not cover it)
  constexpr int U = W <= 2 ? 4 : (W <= 4 ? 2 : 1);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i0 = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < hi; i0 += stride * U) {
    float4 v[U][W];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * stride;
      if (i < hi) {
#pragma unroll
        for (int q = 0; q < W; ++q) v[u][q] = reinterpret_cast<const float4*>(a.bufs[q])[i];   // peer loads
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * stride;
      if (i < hi) {
        float4 s = v[u][0];
#pragma unroll
        for (int q = 1; q < W; ++q) { s.x += v[u][q].x; s.y += v[u][q].y; s.z += v[u][q].z; s.w += v[u][q].w; }
        s.x *= inv; s.y *= inv; s.z *= inv; s.w *= inv;
#pragma unroll
        for (int q = 0; q < W; ++q) reinterpret_cast<float4*>(a.bufs[q])[i] = s;               // peer stores
      }
    }
  }
}

}  // namespace basi

using namespace basi;

extern "C" {

/* bufs / flags: `world` pointers each (this process's mappings of every rank's symmetric buffer and flag array, the
 * flags zero at start-up); seq_dev: a device uint32 that starts at 0 on every rank.  n floats (multiple of 4) are
 * averaged in place in every rank's buffer.  Two flag barriers bracket the exchange; all ranks must call it. */
int basi_p2p_allreduce_mean(void* const* bufs, void* const* flags, int rank, int world, int64_t n, uint32_t* seq_dev,
                            void* stream) {
  BASI_CHECK_ARG(bufs && flags && seq_dev && world >= 2 && world <= P2P_MAX_WORLD && rank >= 0 && rank < world &&
                     n > 0 && n % 4 == 0, "p2p_allreduce_mean: bad argument (2..8 ranks, n a multiple of 4)");
  P2PArgs a{};
  for (int q = 0; q < world; ++q) {
    BASI_CHECK_ARG(bufs[q] && flags[q] && ((uintptr_t)bufs[q] & 15) == 0, "p2p_allreduce_mean: null / unaligned peer buffer");
    a.bufs[q] = (float*)bufs[q];
    a.flags[q] = (unsigned int*)flags[q];
  }
  a.rank = rank;
  a.world = world;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n4 = n / 4;
  const int grid = sm_count() * 4;
  p2p_barrier_kernel<<<1, 32, 0, st>>>(a, seq_dev);
  switch (world) {
    case 2: p2p_allreduce_mean_kernel<2><<<grid, 256, 0, st>>>(a, n4); break;
    case 3: p2p_allreduce_mean_kernel<3><<<grid, 256, 0, st>>>(a, n4); break;
    case 4: p2p_allreduce_mean_kernel<4><<<grid, 256, 0, st>>>(a, n4); break;
    case 5: p2p_allreduce_mean_kernel<5><<<grid, 256, 0, st>>>(a, n4); break;
    case 6: p2p_allreduce_mean_kernel<6><<<grid, 256, 0, st>>>(a, n4); break;
    case 7: p2p_allreduce_mean_kernel<7><<<grid, 256, 0, st>>>(a, n4); break;
    default: p2p_allreduce_mean_kernel<8><<<grid, 256, 0, st>>>(a, n4); break;
  }
  p2p_barrier_kernel<<<1, 32, 0, st>>>(a, seq_dev);
  BASI_CHECK_LAUNCH("p2p_allreduce_mean");
  return BASI_OK;
}

}  // extern "C"
