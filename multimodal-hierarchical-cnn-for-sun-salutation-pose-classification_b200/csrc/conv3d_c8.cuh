// First Conv3d of Quadtree3DCNN (3dcnn/models.py:107-112: Conv3d(3 -> 32, 3x3x3, pad 1, bias) on 16 x 112 x 112 clips) as a
// persistent tcgen05 kernel. The layer is HBM bound (75 FLOP per byte of output): what matters is that every input byte
// is read once and nothing is expanded.
//
// Input: NDHWC bf16 with the 3 channels padded to 8 (one 16-byte chunk per pixel, functional.PackClip). In the virtual
// zero-padded plane [H+1][W+2] of the slab kernels (conv3x3.cuh) the three kw taps of a pixel are the three NEXT 16-byte
// chunks of the staged slab, so — as in the 2-D stem (stem.cuh) — consecutive output pixels are operand rows 16 bytes
// apart: the no-swizzle K-major canonical layout (8 rows = 128 contiguous bytes, SBO = 128 B) with LBO = 16 B between the
// two K chunks of a K = 16 MMA. Per (depth tap, kh) two MMAs cover kw = -1,0 and kw = +1,(+2: zero weights). A tile of
// 256 virtual pixels needs 3 slabs (one per depth plane, ~8 KB each) and 2 x 18 MMAs (M = 128, N = 32, K = 16); the
// 18 weight tiles (18 KB) stay resident in shared memory.
// Roles (288 threads): warps 0-3 epilogue (TMEM -> +bias -> bf16 -> global, BatchNorm partial sums), warps 4-7 slab
// producers (cp.async, one 16-byte row each), warp 8 MMA issuer. Accumulators double-buffered in TMEM (2 x 2 x 32 columns).
#pragma once
#include "igemm.cuh"

namespace qt {

constexpr int kC8Threads = 288;
constexpr int kC8Slabs = 6;      // ring of staged slabs (a tile uses 3)
constexpr int kC8N = 32;         // output channels

struct Conv3dC8Params {
  const __nv_bfloat16* x;    // [NP][H][W][8]
  const __nv_bfloat16* wb;   // [18][2][32][8]: tile (kd, kh, half), K chunk, cout, channel (qt_wpack_conv3d_c8)
  __nv_bfloat16* y;          // [NP][H][W][32]
  const float* bias;         // [32] or NULL
  float* stats;              // [gridDim.x][2][32] or NULL
  int NP, D, H, W;           // NP = clips * D planes
  int V, num_tiles, R;       // virtual pixels, 256-pixel tiles, slab rows
};

__global__ void __launch_bounds__(kC8Threads, 1) conv3d_c8_kernel(const __grid_constant__ Conv3dC8Params p) {
  constexpr int BM = 2 * kBM;
  constexpr uint32_t TCOLS = 128;  // 2 buffers x 2 sub-tiles x 32 columns
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* wsm = smem;                                              // 18 tiles x 1 KB
  float* scratch = reinterpret_cast<float*>(smem + 18 * 1024);      // [4 warps][2][32]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 18 * 1024 + 1024);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + kC8Slabs;
  uint64_t* acc_full = a_empty + kC8Slabs;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  uint8_t* slabs = smem + 20 * 1024;
  const int slab_bytes = (p.R * 16 + 1023) / 1024 * 1024;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Wp = p.W + 2, Hp = p.H + 1;

  // resident weights: global [tile][kchunk][cout][16 B] is already the shared-memory layout (no-swizzle K-major planes)
  for (int i = threadIdx.x; i < 18 * 64; i += kC8Threads)
    reinterpret_cast<uint4*>(wsm)[i] = reinterpret_cast<const uint4*>(p.wb)[i];
  fence_proxy_async_smem();
  if (warp == 8) {
    if (lane == 0) {
      for (int s = 0; s < kC8Slabs; ++s) { mbar_init(&a_full[s], kProducerThreads); mbar_init(&a_empty[s], 1); }
      for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], kProducerThreads); }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc<TCOLS>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= 4 && warp < 8) {
    // ================================================================= producers: slab row j <-> virtual pixel q0 - (W+3) + j
    const int t = threadIdx.x - 128;
    const int adv_w = 128 % Wp, adv_h = 128 / Wp;
    const long long pix_bytes = 16, row_bytes = static_cast<long long>(p.W) * 16, plane_bytes = row_bytes * p.H;
    uint32_t cnt = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int q0 = tile * BM;
      for (int kd = 0; kd < 3; ++kd, ++cnt) {
        const int s = cnt % kC8Slabs;
        if (cnt >= kC8Slabs) mbar_wait(&a_empty[s], ((cnt / kC8Slabs) - 1) & 1);
        const int dd = kd - 1;
        int n, hp, wp, dz;
        {
          const int vv = q0 - (p.W + 3) + t + Wp * Hp;
          wp = vv % Wp;
          const int rest = vv / Wp;
          hp = rest % Hp;
          n = rest / Hp - 1;
          dz = (n + p.D) % p.D;
        }
        uint32_t dst = smem_u32(slabs + s * slab_bytes) + t * 16;
        for (int j = t; j < p.R; j += 128) {
          const bool ok = (static_cast<unsigned>(n) < static_cast<unsigned>(p.NP)) && (static_cast<unsigned>(wp - 1) < static_cast<unsigned>(p.W)) &&
                          (hp >= 1) && (static_cast<unsigned>(dz + dd) < static_cast<unsigned>(p.D));
          const char* src = reinterpret_cast<const char*>(p.x) + dd * plane_bytes +
                            (static_cast<long long>(n) * p.H + (hp - 1)) * row_bytes + static_cast<long long>(wp - 1) * pix_bytes;
          cp_async16(dst, ok ? static_cast<const void*>(src) : static_cast<const void*>(p.x), ok ? 16u : 0u);
          dst += 128 * 16;
          wp += adv_w; hp += adv_h;
          while (wp >= Wp) { wp -= Wp; ++hp; }
          while (hp >= Hp) { hp -= Hp; ++n; dz = (dz + 1 == p.D) ? 0 : dz + 1; }
        }
        cp_async_mbar_arrive_noinc(&a_full[s]);
      }
    }
    cp_async_wait<0>();
  } else if (warp == 8) {
    // ================================================================= MMA issuer
    constexpr uint32_t idesc = make_idesc_bf16(kBM, kC8N, 0, 0);
    constexpr uint32_t hi = (128u >> 4) | (1u << 14) | (kLayoutNone << 29);  // SBO = 128 B (next 8 rows)
    constexpr uint32_t a_lbo = (16u >> 4) << 16;    // next K chunk = the next pixel (16 B further)
    constexpr uint32_t b_lbo = (512u >> 4) << 16;   // next K chunk = next [32 cout][16 B] plane
    const uint32_t tbase = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t w_lo = ((smem_u32(wsm) >> 4) & 0x3FFFu) | b_lbo;
    uint32_t cnt = 0, it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const uint32_t ab = it & 1;
      if (it >= 2) mbar_wait(&acc_empty[ab], ((it >> 1) - 1) & 1);
      tc_fence_after();
      for (int kd = 0; kd < 3; ++kd, ++cnt) {
        const int s = cnt % kC8Slabs;
        mbar_wait(&a_full[s], (cnt / kC8Slabs) & 1);
        fence_proxy_async_smem();
        tc_fence_after();
        const uint32_t slab_lo = ((smem_u32(slabs + s * slab_bytes) >> 4) & 0x3FFFu) | a_lbo;
        if (elect_one()) {
#pragma unroll
          for (int u = 0; u < 2; ++u) {
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
              for (int half = 0; half < 2; ++half) {
                // output pixel m, tap (kh-1, dw): slab row m + kh*Wp + (dw+1); half 0 starts at dw = -1, half 1 at dw = +1
                const uint32_t a_lo = slab_lo + static_cast<uint32_t>(u * kBM + kh * Wp + 2 * half);
                const uint32_t b_lo = w_lo + static_cast<uint32_t>(((kd * 3 + kh) * 2 + half) * (1024 >> 4));
                umma_bf16(tbase + ab * (2 * kC8N) + u * kC8N, (static_cast<uint64_t>(hi) << 32) | a_lo, (static_cast<uint64_t>(hi) << 32) | b_lo,
                          idesc, (kd | kh | half) ? 1u : 0u);
              }
            }
          }
          umma_commit(&a_empty[s]);
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(&acc_full[ab]);
      __syncwarp();
    }
  } else {
    // ================================================================= epilogue (warps 0-3): one accumulator row per lane
    float run1 = 0.f, run2 = 0.f;  // per-lane (= column) BatchNorm partial sums of this warp
    const bool want_stats = p.stats != nullptr;
    float bias_r[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) bias_r[j] = p.bias ? __ldg(p.bias + j) : 0.f;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const uint32_t ab = it & 1;
      mbar_wait(&acc_full[ab], (it >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int v = tile * BM + u * kBM + warp * 32 + lane;
        bool ok = v < p.V;
        long long orow = 0;
        if (ok) {
          const int wp = v % Wp;
          const int rest = v / Wp;
          const int hp = rest % Hp;
          const int n = rest / Hp;
          ok = (wp >= 1) && (wp <= p.W) && (hp >= 1);
          orow = ((static_cast<long long>(n) * p.H + (hp - 1)) * p.W + (wp - 1)) * kC8N;
        }
        uint32_t r[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + ab * (2 * kC8N) + u * kC8N, r);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2(__uint_as_float(r[2 * j]) + bias_r[2 * j], __uint_as_float(r[2 * j + 1]) + bias_r[2 * j + 1]);
        if (ok) {
          uint4* dst = reinterpret_cast<uint4*>(p.y + orow);
#pragma unroll
          for (int j = 0; j < 4; ++j) dst[j] = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
        }
        if (want_stats) {
          float s1v[32], s2v[32];
          const uint32_t keep = ok ? 0xffffffffu : 0u;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float lo = __uint_as_float((pk[j] << 16) & keep), hi2 = __uint_as_float(pk[j] & 0xffff0000u & keep);
            s1v[2 * j] = lo; s1v[2 * j + 1] = hi2;
            s2v[2 * j] = lo * lo; s2v[2 * j + 1] = hi2 * hi2;
          }
          run1 += warp_transpose_reduce(s1v);
          run2 += warp_transpose_reduce(s2v);
        }
      }
      tc_fence_before();
      mbar_arrive(&acc_empty[ab]);
    }
    if (want_stats) {
      scratch[(warp * 2 + 0) * 32 + lane] = run1;
      scratch[(warp * 2 + 1) * 32 + lane] = run2;
      asm volatile("bar.sync 1, 128;\n" ::: "memory");
      if (threadIdx.x < 64) {
        const int which = threadIdx.x >> 5, col = threadIdx.x & 31;
        const float tot = (scratch[(0 * 2 + which) * 32 + col] + scratch[(1 * 2 + which) * 32 + col]) +
                          (scratch[(2 * 2 + which) * 32 + col] + scratch[(3 * 2 + which) * 32 + col]);
        p.stats[(static_cast<long long>(blockIdx.x) * 2 + which) * kC8N + col] = tot;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc<TCOLS>(tmem_base);
}

}  // namespace qt
