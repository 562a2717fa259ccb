// ResNet stem convolution (7x7, stride 2, pad 3, 3 input channels -> 64) as a tcgen05 kernel that needs NO
// im2col at all.
//
// The input is packed as zero-padded NHWC4 bf16 rows ([N][H+7][W+8][4], qt_stem_pack_input). For output row ho
// and filter row r, the 32 K-values of output pixel wo (7 taps x 4 channels + one zero-weight tap) are the 64
// contiguous bytes of input row 2*ho+r starting at byte 16*wo. Neighbouring output pixels therefore OVERLAP: row
// i of the A operand starts 16 bytes after row i-1 — which is exactly the no-swizzle K-major canonical layout
// (rows of a core matrix 16 B apart, SBO = 128 B per 8 rows) with LBO = 16 B between consecutive K chunks.
// So the raw input rows, copied once into shared memory with a 1-D bulk async copy (TMA, `cp.async.bulk`), ARE
// the A operand of all seven row-taps; nothing is expanded, gathered or re-read.
//
// Tile = 4 output rows of one image (13 contiguous input rows = one bulk copy of 24 KB). Per tile: 4 accumulators
// [128 x 64] (112 valid pixels each at 224x224), 4 x 7 x 2 MMAs (M=128, N=64, K=16). Filter (7 x [64][32]) is
// resident in shared memory. Persistent CTAs, 4-stage input ring, double-buffered TMEM (2 x 256 columns).
// Roles: warps 0-3 epilogue, warp 4 TMA producer, warp 5 MMA issuer.
#pragma once
#include "igemm.cuh"

namespace qt {

constexpr int kStemThreads = 192;
constexpr int kStemRows = 4;     // output rows per tile
constexpr int kStemStages = 4;

struct StemParams {
  const __nv_bfloat16* xp;  // [N][Hp][Wp][4]
  const __nv_bfloat16* w8;  // [64][8][32]
  __nv_bfloat16* y;         // [N][Ho][Wo][64]
  float* stats;             // [gridDim.x][2][64] (optional)
  int N, H, W, Ho, Wo, Hp, Wp;
  int strips;               // ceil(Ho / 4)
  int num_tiles;            // N * strips
  int row_bytes;            // Wp * 8
  int stage_bytes;          // 13 * row_bytes + 2304 slack (rows 112..127 of the M=128 operand read past the row), multiple of 128
};

__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst_smem),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__global__ void __launch_bounds__(kStemThreads, 1) stem_fprop_kernel(const __grid_constant__ StemParams p) {
  constexpr int BN = 64;
  constexpr uint32_t TCOLS = 2 * kStemRows * BN;  // 512
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // layout: [weights 7 x 4 planes x 1 KB = 28 KB][scratch 2*4*64 floats + running 2*64][barriers][input stages]
  uint8_t* wsm = smem;
  float* scratch = reinterpret_cast<float*>(smem + 28 * 1024);
  float* running = scratch + 2 * 4 * BN;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 28 * 1024 + 4096);
  uint64_t* in_full = bars;
  uint64_t* in_empty = in_full + kStemStages;
  uint64_t* acc_full = in_empty + kStemStages;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  uint8_t* stages = smem + 28 * 1024 + 4096 + 256;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // filter -> shared memory, no-swizzle K-major planes per row-tap r: plane j = [64 couts][16 B]
  for (int i = threadIdx.x; i < 7 * 64 * 4; i += kStemThreads) {
    const int j = i & 3, n = (i >> 2) & 63, r = i >> 8;
    const uint4 v = *reinterpret_cast<const uint4*>(p.w8 + (n * 8 + r) * 32 + j * 8);
    *reinterpret_cast<uint4*>(wsm + r * 4096 + j * 1024 + n * 16) = v;
  }
  fence_proxy_async_smem();
  if (warp == 5) {
    if (lane == 0) {
      for (int s = 0; s < kStemStages; ++s) { mbar_init(&in_full[s], 1); mbar_init(&in_empty[s], 1); }
      for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 128); }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc<TCOLS>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    // ------------------------------------------------------------- TMA producer: one bulk copy per tile
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const int s = it % kStemStages;
        if (it >= kStemStages) mbar_wait(&in_empty[s], ((it / kStemStages) - 1) & 1);
        const int n = tile / p.strips, strip = tile - n * p.strips;
        const int ho0 = strip * kStemRows;
        const int rows = min(2 * kStemRows + 5, p.Hp - 2 * ho0);
        const uint32_t bytes = static_cast<uint32_t>(rows) * p.row_bytes;
        const __nv_bfloat16* src = p.xp + (static_cast<long long>(n) * p.Hp + 2 * ho0) * p.Wp * 4;
        mbar_arrive_expect_tx(&in_full[s], bytes);
        bulk_copy_g2s(smem_u32(stages + static_cast<size_t>(s) * p.stage_bytes), src, bytes, &in_full[s]);
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    // ------------------------------------------------------------- MMA issuer
    constexpr uint32_t idesc = make_idesc_bf16(kBM, BN, 0, 0);
    // A: no swizzle, LBO = 16 B (next K chunk = next 16 bytes of the same input row), SBO = 128 B
    constexpr uint32_t a_hi = (128u >> 4) | (1u << 14) | (kLayoutNone << 29);
    constexpr uint32_t a_lbo = (16u >> 4) << 16;
    // B: no swizzle planes, LBO = 1024 B between K chunks, SBO = 128 B between 8-cout groups
    constexpr uint32_t b_hi = (128u >> 4) | (1u << 14) | (kLayoutNone << 29);
    constexpr uint32_t b_lbo = (1024u >> 4) << 16;
    const uint32_t tbase = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t w_lo = ((smem_u32(wsm) >> 4) & 0x3FFFu) | b_lbo;
    const uint32_t row16 = static_cast<uint32_t>(p.row_bytes) >> 4;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int s = it % kStemStages;
      const uint32_t ab = it & 1;
      if (it >= 2) mbar_wait(&acc_empty[ab], ((it >> 1) - 1) & 1);
      mbar_wait(&in_full[s], (it / kStemStages) & 1);
      tc_fence_after();
      const uint32_t st_lo = ((smem_u32(stages + static_cast<size_t>(s) * p.stage_bytes) >> 4) & 0x3FFFu) | a_lbo;
      if (elect_one()) {
#pragma unroll 1
        for (int q = 0; q < kStemRows; ++q) {
          const uint32_t d = tbase + ab * (kStemRows * BN) + q * BN;
#pragma unroll
          for (int r = 0; r < 7; ++r) {
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
              const uint64_t ad = (static_cast<uint64_t>(a_hi) << 32) | (st_lo + (2 * q + r) * row16 + kk * 2);
              const uint64_t bd = (static_cast<uint64_t>(b_hi) << 32) | (w_lo + r * (4096 >> 4) + kk * (2048 >> 4));
              umma_bf16(d, ad, bd, idesc, (r | kk) ? 1u : 0u);
            }
          }
        }
        umma_commit(&in_empty[s]);
        umma_commit(&acc_full[ab]);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------- epilogue (warps 0-3)
    const bool want_stats = p.stats != nullptr;
    if (want_stats) {
      for (int i = threadIdx.x; i < 2 * BN; i += 128) running[i] = 0.f;
      asm volatile("bar.sync 1, 128;\n" ::: "memory");
    }
    uint32_t it = 0;
    const int wo = warp * 32 + lane;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const uint32_t ab = it & 1;
      const int n = tile / p.strips, strip = tile - n * p.strips;
      mbar_wait(&acc_full[ab], (it >> 1) & 1);
      tc_fence_after();
      // chunk-outer / row-inner: the four output rows of a tile share their columns, so the per-thread partial
      // sums are combined across rows first and transposed/reduced across lanes once per 32-column chunk.
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        float a1[32], a2[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) { a1[j] = 0.f; a2[j] = 0.f; }
#pragma unroll 1
        for (int q = 0; q < kStemRows; ++q) {
          const int ho = strip * kStemRows + q;
          const bool row_ok = (ho < p.Ho) && (wo < p.Wo);
          __nv_bfloat16* dst = p.y + ((static_cast<long long>(n) * p.Ho + ho) * p.Wo + wo) * BN;
          uint32_t r[32];
          tmem_ld32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + ab * (kStemRows * BN) + q * BN + c0, r);
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = bf16_round(__uint_as_float(r[j]));
          if (row_ok) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              uint4 qv;
              qv.x = pack_bf16x2(v[j], v[j + 1]);
              qv.y = pack_bf16x2(v[j + 2], v[j + 3]);
              qv.z = pack_bf16x2(v[j + 4], v[j + 5]);
              qv.w = pack_bf16x2(v[j + 6], v[j + 7]);
              *reinterpret_cast<uint4*>(dst + c0 + j) = qv;
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) { a1[j] += v[j]; a2[j] = fmaf(v[j], v[j], a2[j]); }
          }
        }
        if (want_stats) {
          const float s1 = warp_transpose_reduce(a1);
          const float s2 = warp_transpose_reduce(a2);
          scratch[(0 * 4 + warp) * BN + c0 + lane] = s1;
          scratch[(1 * 4 + warp) * BN + c0 + lane] = s2;
        }
      }
      if (want_stats) {
        asm volatile("bar.sync 1, 128;\n" ::: "memory");
        if (threadIdx.x < 2 * BN) {
          const int which = threadIdx.x / BN, col = threadIdx.x - which * BN;
          const float* sc = scratch + which * 4 * BN + col;
          running[which * BN + col] += (sc[0] + sc[BN]) + (sc[2 * BN] + sc[3 * BN]);
        }
        asm volatile("bar.sync 1, 128;\n" ::: "memory");
      }
      tc_fence_before();
      mbar_arrive(&acc_empty[ab]);
    }
    if (want_stats && threadIdx.x < 2 * BN) p.stats[static_cast<long long>(blockIdx.x) * 2 * BN + threadIdx.x] = running[threadIdx.x];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc<TCOLS>(tmem_base);
}

}  // namespace qt

namespace qt {

// ---------------------------------------------------------------------------------------------------------
// Stem weight gradient: dW8[cout][r][k] = sum_{n,ho,wo} dy[n,ho,wo,cout] * xp[n][2ho+r][8*wo + k], k = 0..31.
// Per output row and row-tap r one GEMM  D[cout (64, M padded to 128 with a zero block)][32] += dy^T * X_r with
// the pixels of the row as the K dimension (16 per MMA; Wo % 16 == 0):
//   A = dy row [Wo][64] gathered with cp.async into the 128B-swizzled MN-major layout (pixel rows of 128 B),
//   B = the raw input row, again used in place: MN-major no-swizzle with SBO = 16 B (next 8 features), 16 B between
//       consecutive pixels of a core matrix and LBO = 128 B between 8-pixel groups.
// Accumulators (7 taps x 32 columns) stay in TMEM for the whole kernel; every CTA writes one fp32 partial
// [64][224] at the end and a small kernel reduces over CTAs into the PyTorch layout [64][3][7][7].
// Tile = 2 output rows (9 contiguous input rows by one bulk copy + 2 dy rows by cp.async), 3-stage ring.
// ---------------------------------------------------------------------------------------------------------
constexpr int kSWRows = 2;
constexpr int kSWStages = 3;
constexpr int kSWThreads = 160;

struct StemWgradParams {
  const __nv_bfloat16* xp;
  const __nv_bfloat16* dy;   // [N][Ho][Wo][64]
  float* partial;            // [gridDim.x][64][224]
  int N, Ho, Wo, Hp, Wp;
  int strips, num_tiles;
  int row_bytes;             // Wp * 8
  int x_bytes;               // stage bytes for the input rows (9 rows + slack), multiple of 1024
  int dy_bytes;              // Wo * 128 per dy row, multiple of 1024 required
};

__global__ void __launch_bounds__(kSWThreads, 1) stem_wgrad_kernel(const __grid_constant__ StemWgradParams p) {
  constexpr uint32_t TCOLS = 256;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int stage_bytes = kSWRows * p.dy_bytes + p.x_bytes;
  uint8_t* zero_blk = smem + kSWStages * stage_bytes;          // p.dy_bytes of zeros (M rows 64..127)
  uint64_t* bars = reinterpret_cast<uint64_t*>(zero_blk + p.dy_bytes);
  uint64_t* full = bars;
  uint64_t* empty = full + kSWStages;
  uint64_t* done = empty + kSWStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // zero block (M rows 64..127) and, so that never-loaded halo rows cannot hold NaN patterns, all stages
  for (int i = threadIdx.x * 16; i < kSWStages * stage_bytes + p.dy_bytes; i += kSWThreads * 16)
    *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  if (warp == 4) {
    if (lane == 0) {
      for (int s = 0; s < kSWStages; ++s) { mbar_init(&full[s], 129); mbar_init(&empty[s], 1); }
      mbar_init(done, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc<TCOLS>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  int my_tiles = 0;
  for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) ++my_tiles;

  if (warp < 4) {
    // ------------------------------------------------------------- producers
    const int t = threadIdx.x;
    const int chunk = t & 7;
    const int rbase = t >> 3;  // dy pixels rbase + 16*i
    const uint32_t sw = static_cast<uint32_t>((chunk ^ (rbase & 7)) << 4);
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int s = it % kSWStages;
      if (it >= kSWStages) mbar_wait(&empty[s], ((it / kSWStages) - 1) & 1);
      const int n = tile / p.strips, strip = tile - n * p.strips;
      const int ho0 = strip * kSWRows;
      uint8_t* st = smem + s * stage_bytes;
      if (t == 0) {
        const int rows = min(2 * kSWRows + 5, p.Hp - 2 * ho0);
        const uint32_t bytes = static_cast<uint32_t>(rows) * p.row_bytes;
        mbar_arrive_expect_tx(&full[s], bytes);
        bulk_copy_g2s(smem_u32(st + kSWRows * p.dy_bytes), p.xp + (static_cast<long long>(n) * p.Hp + 2 * ho0) * p.Wp * 4, bytes,
                      &full[s]);
      }
      for (int q = 0; q < kSWRows; ++q) {
        const int ho = ho0 + q;
        const bool row_ok = ho < p.Ho;
        const __nv_bfloat16* src_row = p.dy + ((static_cast<long long>(n) * p.Ho + (row_ok ? ho : 0)) * p.Wo) * 64 + chunk * 8;
        const uint32_t dst = smem_u32(st + q * p.dy_bytes) + sw;
        for (int px = rbase; px < p.Wo; px += 16)
          cp_async16(dst + px * 128, src_row + static_cast<long long>(px) * 64, row_ok ? 16u : 0u);
      }
      cp_async_commit();
      // publish the previous tile (lag 1)
      if (it >= 1) {
        cp_async_wait<1>();
        fence_proxy_async_smem();
        mbar_arrive(&full[(it - 1) % kSWStages]);
      }
    }
    cp_async_wait<0>();
    fence_proxy_async_smem();
    if (it >= 1) mbar_arrive(&full[(it - 1) % kSWStages]);
  } else {
    // ------------------------------------------------------------- MMA issuer
    constexpr uint32_t idesc = make_idesc_bf16(kBM, 32, 1, 1);
    constexpr uint32_t a_hi = (1024u >> 4) | (1u << 14) | (kLayoutSW128 << 29);  // SBO 1024: next 8 pixels
    constexpr uint32_t b_hi = (16u >> 4) | (1u << 14) | (kLayoutNone << 29);     // SBO 16: next 8 features
    constexpr uint32_t b_lbo = (128u >> 4) << 16;                                // LBO 128: next 8 pixels
    const uint32_t tbase = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t row16 = static_cast<uint32_t>(p.row_bytes) >> 4;
    const int ksteps = p.Wo / 16;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int s = it % kSWStages;
      mbar_wait(&full[s], (it / kSWStages) & 1);
      tc_fence_after();
      uint8_t* st = smem + s * stage_bytes;
      const uint32_t x_lo = ((smem_u32(st + kSWRows * p.dy_bytes) >> 4) & 0x3FFFu) | b_lbo;
      if (elect_one()) {
#pragma unroll 1
        for (int q = 0; q < kSWRows; ++q) {
          const uint32_t a_addr = smem_u32(st + q * p.dy_bytes);
          // LBO of A: distance from this dy row to the zero block (cout rows 64..127 of the M=128 operand)
          const uint32_t a_lbo = (((smem_u32(zero_blk) - a_addr) >> 4) & 0x3FFFu) << 16;
          const uint32_t a_lo = ((a_addr >> 4) & 0x3FFFu) | a_lbo;
#pragma unroll 1
          for (int ks = 0; ks < ksteps; ++ks) {
#pragma unroll
            for (int r = 0; r < 7; ++r) {
              const uint64_t ad = (static_cast<uint64_t>(a_hi) << 32) | (a_lo + ks * (2048 >> 4));
              const uint64_t bd = (static_cast<uint64_t>(b_hi) << 32) | (x_lo + (2 * q + r) * row16 + ks * (256 >> 4));
              umma_bf16(tbase + r * 32, ad, bd, idesc, (it | q | ks) ? 1u : 0u);
            }
          }
        }
        umma_commit(&empty[s]);
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(done);
    __syncwarp();
  }

  if (warp < 4) {
    // ------------------------------------------------------------- final epilogue: accumulators -> partial
    if (my_tiles > 0) {
      mbar_wait(done, 0);
      tc_fence_after();
    }
    const int cout = warp * 32 + lane;
#pragma unroll 1
    for (int c0 = 0; c0 < 224; c0 += 32) {
      uint32_t r[32];
      tmem_ld32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + c0, r);
      tmem_ld_wait();
      if (cout < 64) {
        float* dst = p.partial + (static_cast<long long>(blockIdx.x) * 64 + cout) * 224 + c0;
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(dst + j) = my_tiles > 0 ? make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                                           __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]))
                                                             : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc<TCOLS>(tmem_base);
}

// partial[T][64][7*32 (r, s*4+c)] -> dw[64][cin][7][7] (+= when accumulate)
__global__ void stem_wgrad_reduce_kernel(const float* __restrict__ partial, int T, float* __restrict__ dw, int cout, int cin,
                                         int accumulate) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cout * cin * 49) return;
  const int s = i % 7, r = (i / 7) % 7, c = (i / 49) % cin, co = i / (49 * cin);
  float acc = 0.f;
  for (int t = 0; t < T; ++t) acc += partial[(static_cast<long long>(t) * 64 + co) * 224 + r * 32 + s * 4 + c];
  dw[i] = accumulate ? dw[i] + acc : acc;
}

}  // namespace qt
