// tcgen05.mma issue-rate probe (sm_100a): how fast can ONE SM (cta_group::1) or one SM pair (cta_group::2) retire
// SS-mode bf16 MMAs of a given shape when nothing else runs?  Operands are whatever is in shared memory; only the
// timing matters. Used to decide tile shapes of the convolution kernels (profiles/r02_mma_rate.md).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/mma_rate tools/mma_rate.cu && build/mma_rate
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

template <int CG>
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (CG == 1)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a),
                 "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a),
                 "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}

__host__ __device__ constexpr uint32_t idesc_bf16(uint32_t M, uint32_t N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24); }

// NACC accumulators of N columns are cycled (like the double-buffered sub-tiles of the conv kernels); STAGES operand
// stages are cycled so descriptor addresses move like in a real ring. EXTRA: 0 none, 1 = the other 4 warps stream
// shared-memory stores (stand-in for producer traffic) while the MMAs run.
template <int N, int CG, int NACC, int EXTRA>
__global__ void __launch_bounds__(256, 1) probe(int iters, long long* cycles, int arow) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  constexpr int STAGES = 4;
  constexpr int ABYTES = 128 * 128 + 1024, BBYTES = (N / CG) * 128;
  uint8_t* a_ring = smem;
  uint8_t* b_ring = smem + STAGES * ABYTES;
  uint8_t* junk = b_ring + STAGES * BBYTES;  // 16 KB scratch for EXTRA traffic
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t rank = 0;
  if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  for (int i = threadIdx.x; i < (STAGES * (ABYTES + BBYTES) + 16384) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    stop = 0;
  }
  constexpr uint32_t TCOLS = (NACC * N) <= 32 ? 32 : (NACC * N) <= 64 ? 64 : (NACC * N) <= 128 ? 128 : (NACC * N) <= 256 ? 256 : 512;
  if (warp == 0) {
    if (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&slot)), "n"(TCOLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&slot)), "n"(TCOLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
    }
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (CG == 2) {
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tbase = slot;
  if (warp == 0 && rank == 0) {
    if (lane == 0) {
      constexpr uint32_t idesc = idesc_bf16(128 * CG, N);
      constexpr uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
      const long long t0 = clock64();
      for (int it = 0; it < iters; ++it) {
        const int s = it % STAGES;
        const uint32_t a_lo = (((smem_u32(a_ring + s * ABYTES) >> 4) + arow * 8) & 0x3FFFu) | (1u << 16);  // arow: start the A window at an unaligned 128-byte row, like a conv tap
        const uint32_t b_lo = ((smem_u32(b_ring + s * BBYTES) >> 4) & 0x3FFFu) | (1u << 16);
        const uint32_t d = tbase + (it % NACC) * N;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          mma<CG>(d, (static_cast<uint64_t>(hi) << 32) | (a_lo + 2 * k), (static_cast<uint64_t>(hi) << 32) | (b_lo + 2 * k), idesc, (it >= NACC || k) ? 1u : 0u);
      }
      if (CG == 1)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&bar)) : "memory");
      else
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(smem_u32(&bar)),
                     "h"((uint16_t)3)
                     : "memory");
      uint32_t ok = 0;
      while (!ok)
        asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], 0;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
      const long long t1 = clock64();
      if (blockIdx.x == 0) cycles[0] = t1 - t0;
      stop = 1;
    }
    __syncwarp();
  } else if (EXTRA && warp >= 4) {
    // 128 threads x 16 B stores, back to back, until the issuer is done
    const uint32_t dst = smem_u32(junk) + (threadIdx.x - 128) * 16;
    while (!stop) {
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};\n" ::"r"(dst + i * 2048), "r"(i) : "memory");
    }
  } else if (CG == 2 && rank == 1 && threadIdx.x == 0) {
    // peer CTA: the leader's commit is multicast to this CTA's barrier too
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], 0;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    stop = 1;
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (CG == 2) {
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
  }
  if (warp == 0) {
    if (CG == 1)
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tbase), "n"(TCOLS) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(tbase), "n"(TCOLS) : "memory");
  }
}

template <int N, int CG, int NACC, int EXTRA>
void run(const char* name, int sms, int arow = 0) {
  const int iters = 8192;
  const size_t smem = 4 * (128 * 128 + 1024 + (N / CG) * 128) + 16384 + 2048;
  auto k = probe<N, CG, NACC, EXTRA>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  long long* cyc;
  cudaMalloc(&cyc, 8);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(sms);
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CG;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    cudaError_t err = cudaLaunchKernelEx(&cfg, k, iters, cyc, arow);
    cudaEventRecord(e1);
    if (err != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
      printf("%-34s FAILED: %s\n", name, cudaGetErrorString(cudaGetLastError()));
      return;
    }
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep && ms < best) best = ms;
  }
  long long c;
  cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  const double flops = 2.0 * 128 * CG * N * 16 * 4.0 * iters * (sms / CG);
  printf("%-34s %8.3f ms  %7.1f TFLOP/s  %6.1f cycles/MMA (floor %d)\n", name, best, flops / (best * 1e-3) / 1e12, (double)c / (4.0 * iters), 128 * N / 256 / 1);
  cudaFree(cyc);
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount & ~1;
  printf("%s, %d SMs used\n", prop.name, sms);
  run<64, 1, 2, 0>("cg1 M128 N64", sms);
  run<128, 1, 2, 0>("cg1 M128 N128", sms);
  run<256, 1, 2, 0>("cg1 M128 N256", sms);
  run<64, 2, 2, 0>("cg2 M256 N64", sms);
  run<128, 2, 2, 0>("cg2 M256 N128", sms);
  run<256, 2, 2, 0>("cg2 M256 N256", sms);
  run<128, 1, 2, 0>("cg1 M128 N128 A start row+1", sms, 1);
  run<128, 1, 2, 0>("cg1 M128 N128 A start row+3", sms, 3);
  run<128, 1, 2, 0>("cg1 M128 N128 A start row+4", sms, 4);
  run<64, 1, 2, 0>("cg1 M128 N64 A start row+1", sms, 1);
  run<256, 1, 2, 0>("cg1 M128 N256 A start row+1", sms, 1);
  run<128, 2, 2, 0>("cg2 M256 N128 A start row+1", sms, 1);
  run<256, 2, 2, 0>("cg2 M256 N256 A start row+1", sms, 1);
  run<128, 1, 2, 1>("cg1 M128 N128 + st.shared traffic", sms);
  run<256, 1, 2, 1>("cg1 M128 N256 + st.shared traffic", sms);
  run<128, 2, 2, 1>("cg2 M256 N128 + st.shared traffic", sms);
  run<256, 2, 2, 1>("cg2 M256 N256 + st.shared traffic", sms);
  return 0;
}
