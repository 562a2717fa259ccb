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
    const long long row_bytes = static_cast<long long>(p.W) * 16, plane_bytes = row_bytes * p.H;
    uint32_t cnt = 0;
    // Coordinates are carried, never divided, inside the tile loop, and the rows of a tile are decomposed once for its three
    // depth slabs (same producer structure as conv3d_c8_wgrad_kernel below; a division set per slab made these four warps the
    // bottleneck: 416 us per launch with a 154 us MMA floor — profiles/r02_conv3d.md).
    const int tile_adv = gridDim.x * BM;
    const int taw = tile_adv % Wp, tah = (tile_adv / Wp) % Hp, tan = tile_adv / (Wp * Hp);
    const int raw = 128 % Wp, rah = (128 / Wp) % Hp, ran = 128 / (Wp * Hp);  // +128 rows inside a slab
    int swp, shp, sn;  // virtual coordinates of this thread's first slab row: q0 - (W+3) + t
    {
      const int vv = blockIdx.x * BM - (p.W + 3) + t + Wp * Hp;  // shifted by one plane: non-negative
      swp = vv % Wp; shp = (vv / Wp) % Hp; sn = vv / (Wp * Hp) - 1;
    }
    constexpr int kMaxRows = 6;  // ceil(R / 128) for W <= 248 (c8_ok)
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      long long roff[kMaxRows];
      int rdz[kMaxRows];
      bool rok[kMaxRows];
      {
        int wp = swp, hp = shp, n = sn;
        int dz = n % p.D;  // n >= -1
        if (dz < 0) dz += p.D;
#pragma unroll
        for (int i = 0; i < kMaxRows; ++i) {
          rok[i] = (static_cast<unsigned>(n) < static_cast<unsigned>(p.NP)) && (static_cast<unsigned>(wp - 1) < static_cast<unsigned>(p.W)) && (hp >= 1);
          roff[i] = (static_cast<long long>(n) * p.H + (hp - 1)) * row_bytes + static_cast<long long>(wp - 1) * 16;
          rdz[i] = dz;
          wp += raw; hp += rah; n += ran;
          dz += ran % p.D;
          if (wp >= Wp) { wp -= Wp; ++hp; }
          if (hp >= Hp) { hp -= Hp; ++n; ++dz; }
          if (dz >= p.D) dz -= p.D;
        }
        swp += taw; shp += tah; sn += tan;
        if (swp >= Wp) { swp -= Wp; ++shp; }
        if (shp >= Hp) { shp -= Hp; ++sn; }
      }
      for (int kd = 0; kd < 3; ++kd, ++cnt) {
        const int s = cnt % kC8Slabs;
        if (cnt >= kC8Slabs) mbar_wait(&a_empty[s], ((cnt / kC8Slabs) - 1) & 1);
        const int dd = kd - 1;
        const uint32_t dst0 = smem_u32(slabs + s * slab_bytes) + t * 16;
        const char* base = reinterpret_cast<const char*>(p.x) + dd * plane_bytes;
#pragma unroll
        for (int i = 0; i < kMaxRows; ++i) {
          if (t + 128 * i < p.R) {
            const bool ok = rok[i] && (static_cast<unsigned>(rdz[i] + dd) < static_cast<unsigned>(p.D));
            cp_async16(dst0 + i * 2048, ok ? static_cast<const void*>(base + roff[i]) : static_cast<const void*>(p.x), ok ? 16u : 0u);
          }
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
    float s1v[32] = {}, s2v[32] = {};
    const bool want_stats = p.stats != nullptr;
    float bias_r[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) bias_r[j] = p.bias ? __ldg(p.bias + j) : 0.f;
    uint32_t it = 0;
    // output coordinates of this lane's row, carried across tiles like the producers' (no division in the loop)
    const int tile_adv = gridDim.x * BM;
    const int taw = tile_adv % Wp, tah = (tile_adv / Wp) % Hp, tan = tile_adv / (Wp * Hp);
    const int raw = kBM % Wp, rah = (kBM / Wp) % Hp, ran = kBM / (Wp * Hp);
    int ewp, ehp, en;
    {
      const int v0 = blockIdx.x * BM + warp * 32 + lane;
      ewp = v0 % Wp; ehp = (v0 / Wp) % Hp; en = v0 / (Wp * Hp);
    }
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const uint32_t ab = it & 1;
      mbar_wait(&acc_full[ab], (it >> 1) & 1);
      tc_fence_after();
      int wp = ewp, hp = ehp, n = en;
      ewp += taw; ehp += tah; en += tan;
      if (ewp >= Wp) { ewp -= Wp; ++ehp; }
      if (ehp >= Hp) { ehp -= Hp; ++en; }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const bool ok = (n < p.NP) && (wp >= 1) && (wp <= p.W) && (hp >= 1);
        const long long orow = ((static_cast<long long>(n) * p.H + (hp - 1)) * p.W + (wp - 1)) * kC8N;
        wp += raw; hp += rah; n += ran;
        if (wp >= Wp) { wp -= Wp; ++hp; }
        if (hp >= Hp) { hp -= Hp; ++n; }
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
          // per-lane (= per-row) running sums of every column; the cross-lane transpose-reduce happens once after the tile loop
          // (two 32-value shuffle reductions per 128 pixels made these four warps the slowest role of the kernel)
          const uint32_t keep = ok ? 0xffffffffu : 0u;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float lo = __uint_as_float((pk[j] << 16) & keep), hi2 = __uint_as_float(pk[j] & 0xffff0000u & keep);
            s1v[2 * j] += lo; s1v[2 * j + 1] += hi2;
            s2v[2 * j] = fmaf(lo, lo, s2v[2 * j]); s2v[2 * j + 1] = fmaf(hi2, hi2, s2v[2 * j + 1]);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&acc_empty[ab]);
    }
    if (want_stats) {
      run1 = warp_transpose_reduce(s1v);
      run2 = warp_transpose_reduce(s2v);
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


// ---------------------------------------------------------------------------------------------------------------------
// Weight gradient of the same layer: dW[cout][c][kd][kh][kw] = sum over pixels dy[p][cout] * x[p + tap][c].
// GEMM view per (kd, kh): D[cout (M = 64, 32 real)][n = kw*8 + c (N = 32, 24 real)] += dy^T[cout][pixels] * X[pixels][n],
// K = pixels. Both operands are MN-major:
//   A = dy tile re-laid as four 8-cout planes [plane][pixel][16 B] (+4 zero planes for the unused rows of M = 64; an M = 128 operand
//       with 12 zero planes measured 67 cycles per MMA: the A read from shared memory is what bounds these small-N MMAs);
//   B = the staged x slab itself: pixel rows 16 bytes apart, and the n-chunk (kw) stride is ALSO 16 bytes (LBO = 16 B), i.e.
//       overlapping descriptors again — the three kw taps (and one meaningless fourth) of a pixel are its next chunks.
// A persistent CTA keeps its 9 accumulators [64 x 32] in TMEM (M = 64: row r lives in TMEM lane 32*(r/16) + r%16) over its whole pixel range (no per-tile epilogue) and writes
// one fp32 partial [9][32 n][32 cout] at the end; conv3d_c8_wgrad_reduce_kernel sums the CTAs in fixed order and scatters
// to the parameter layout. Roles (160 threads): warps 0-3 producers (then the final epilogue), warp 4 MMA issuer.
constexpr int kC8wThreads = 160;
constexpr int kC8wKP = 128;     // virtual pixels per k-tile
constexpr int kC8wSlabs = 12;   // x slab ring (a tile uses 3): four tiles of lookahead — the kernel is latency bound on the HBM stream
constexpr int kC8wDy = 5;       // dy tile ring

struct Conv3dC8WgradParams {
  const __nv_bfloat16* x;    // [NP][H][W][8]
  const __nv_bfloat16* dy;   // [NP][H][W][32]
  float* partial;            // [gridDim.x][9][32][32]
  int NP, D, H, W;
  int V, num_tiles, R;       // R: slab rows
  int debug;                 // developer experiments (g_tune[9]): 1 = skip the MMAs, 2 = skip the slab copies
};

__global__ void __launch_bounds__(kC8wThreads, 1) conv3d_c8_wgrad_kernel(const __grid_constant__ Conv3dC8WgradParams p) {
  constexpr uint32_t TCOLS = 512;  // 9 accumulators x 32 columns
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int slab_bytes = (p.R * 16 + 1023) / 1024 * 1024;
  // plane pitch 2048 + 16 B: the MMA fetches one 16-byte chunk from each of the 8 planes for a given pixel; with a pitch
  // of 2048 B all eight fall into the same shared-memory banks (measured 67 cycles per MMA instead of 16)
  constexpr int kPlane = kC8wKP * 16 + 16;
  constexpr int kDyBytes = (8 * kPlane + 127) / 128 * 128;  // 4 data planes + 4 zero planes of [128 pixels][16 B]
  uint8_t* dys = smem;                                      // [kC8wDy][kDyBytes]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kC8wDy * kDyBytes);
  uint64_t* a_full = bars;                                  // x slabs
  uint64_t* a_empty = a_full + kC8wSlabs;
  uint64_t* d_full = a_empty + kC8wSlabs;                   // dy tiles
  uint64_t* d_empty = d_full + kC8wDy;
  uint64_t* done = d_empty + kC8wDy;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  uint8_t* slabs = smem + kC8wDy * kDyBytes + 1024;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Wp = p.W + 2, Hp = p.H + 1;
  // the zero planes (cout rows 32..63 of the M = 64 operand) are written once
  for (int i = threadIdx.x; i < kC8wDy * 4 * kC8wKP; i += kC8wThreads) {
    const int st = i / (4 * kC8wKP), r = i - st * 4 * kC8wKP;
    *reinterpret_cast<uint4*>(dys + st * kDyBytes + (4 + r / kC8wKP) * kPlane + (r % kC8wKP) * 16) = make_uint4(0, 0, 0, 0);
  }
  fence_proxy_async_smem();
  if (warp == 4) {
    if (lane == 0) {
      for (int s = 0; s < kC8wSlabs; ++s) { mbar_init(&a_full[s], kProducerThreads); mbar_init(&a_empty[s], 1); }
      for (int s = 0; s < kC8wDy; ++s) { mbar_init(&d_full[s], kProducerThreads); mbar_init(&d_empty[s], 1); }
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
  int ntiles_mine = 0;
  for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) ++ntiles_mine;

  if (warp < 4) {
    // ================================================================= producers
    const int t = threadIdx.x;
    const long long row_bytes_x = static_cast<long long>(p.W) * 16, plane_bytes_x = row_bytes_x * p.H;
    uint32_t scnt = 0, dcnt = 0;
    // Coordinates are carried, never divided, inside the tile loop: a single warp runs dependent integer code at ~4 cycles
    // per instruction, and a division set per row made the four producer warps — not the copies or the MMAs — the
    // bottleneck (1.02 ms -> 0.50 ms with one set per tile -> this version; ablation in profiles/r02_conv3d.md).
    // State: virtual coordinates of the thread's dy row (q0 + t) and of its first slab row (q0 - (W+3) + t, shifted by one
    // plane so the decomposition is non-negative); both advance by gridDim.x * 128 virtual pixels per tile.
    const int tile_adv = gridDim.x * kC8wKP;
    const int taw = tile_adv % Wp, tah = (tile_adv / Wp) % Hp, tan = tile_adv / (Wp * Hp);
    const int raw = 128 % Wp, rah = (128 / Wp) % Hp, ran = 128 / (Wp * Hp);  // +128 rows inside a slab
    int dwp, dhp, dn, swp, shp, sn;
    {
      const int v = blockIdx.x * kC8wKP + t;
      dwp = v % Wp; dhp = (v / Wp) % Hp; dn = v / (Wp * Hp);
      const int vv = v - (p.W + 3) + Wp * Hp;
      swp = vv % Wp; shp = (vv / Wp) % Hp; sn = vv / (Wp * Hp) - 1;
    }
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      {  // dy tile: row t <-> virtual pixel q0 + t, four 16-byte cout chunks -> four planes
        const int s = dcnt % kC8wDy;
        if (dcnt >= kC8wDy) mbar_wait(&d_empty[s], ((dcnt / kC8wDy) - 1) & 1);
        const bool ok = (dn < p.NP) && (dwp >= 1) && (dwp <= p.W) && (dhp >= 1);
        const char* src = reinterpret_cast<const char*>(p.dy) + ((static_cast<long long>(dn) * p.H + (dhp - 1)) * p.W + (dwp - 1)) * 64;
        const uint32_t dst = smem_u32(dys + s * kDyBytes) + t * 16;
#pragma unroll
        for (int c = 0; c < 4; ++c)
          cp_async16(dst + c * kPlane, ok ? static_cast<const void*>(src + c * 16) : static_cast<const void*>(p.dy), ok ? 16u : 0u);
        cp_async_mbar_arrive_noinc(&d_full[s]);
        ++dcnt;
        dwp += taw; dhp += tah; dn += tan;
        if (dwp >= Wp) { dwp -= Wp; ++dhp; }
        if (dhp >= Hp) { dhp -= Hp; ++dn; }
      }
      // x slabs: row j <-> virtual pixel q0 - (W+3) + j of the plane kd - 1 away; the three slabs of a tile cover the same
      // rows, so their coordinates are derived once per tile and reused for the three depth taps
      constexpr int kMaxRows = 5;  // ceil(R / 128) for W <= 250
      long long roff[kMaxRows];
      int rdz[kMaxRows];
      bool rok[kMaxRows];
      {
        int wp = swp, hp = shp, n = sn;
        int dz = n % p.D;  // (n >= -1; the one division left per tile)
        if (dz < 0) dz += p.D;
#pragma unroll
        for (int i = 0; i < kMaxRows; ++i) {
          rok[i] = (static_cast<unsigned>(n) < static_cast<unsigned>(p.NP)) && (static_cast<unsigned>(wp - 1) < static_cast<unsigned>(p.W)) && (hp >= 1);
          roff[i] = (static_cast<long long>(n) * p.H + (hp - 1)) * row_bytes_x + static_cast<long long>(wp - 1) * 16;
          rdz[i] = dz;
          wp += raw; hp += rah; n += ran;
          dz += ran % p.D;
          if (wp >= Wp) { wp -= Wp; ++hp; }
          if (hp >= Hp) { hp -= Hp; ++n; ++dz; }
          if (dz >= p.D) dz -= p.D;
        }
        swp += taw; shp += tah; sn += tan;
        if (swp >= Wp) { swp -= Wp; ++shp; }
        if (shp >= Hp) { shp -= Hp; ++sn; }
      }
      for (int kd = 0; kd < 3; ++kd, ++scnt) {
        const int s = scnt % kC8wSlabs;
        if (scnt >= kC8wSlabs) mbar_wait(&a_empty[s], ((scnt / kC8wSlabs) - 1) & 1);
        const int dd = kd - 1;
        const uint32_t dst0 = smem_u32(slabs + s * slab_bytes) + t * 16;
        const char* base = reinterpret_cast<const char*>(p.x) + dd * plane_bytes_x;
#pragma unroll
        for (int i = 0; i < kMaxRows; ++i) {
          if (t + 128 * i < p.R) {
            const bool ok = rok[i] && (static_cast<unsigned>(rdz[i] + dd) < static_cast<unsigned>(p.D));
            if (p.debug != 2) cp_async16(dst0 + i * 2048, ok ? static_cast<const void*>(base + roff[i]) : static_cast<const void*>(p.x), ok ? 16u : 0u);
          }
        }
        cp_async_mbar_arrive_noinc(&a_full[s]);
      }
    }
    cp_async_wait<0>();
  } else {
    // ================================================================= MMA issuer
    constexpr uint32_t idesc = make_idesc_bf16(64, 32, 1, 1);  // both operands MN-major
    // MN-major, no swizzle (cute mma_sm100_desc.hpp: ((1,n),(8,k)):((X,SBO),(1,LBO)) in 16-byte units): SBO = stride between
    // 8-element MN chunks, LBO = stride between groups of 8 K rows (the opposite roles of the swizzled MN-major layouts)
    constexpr uint32_t a_hi = ((kC8wKP * 16u + 16u) >> 4) | (1u << 14) | (kLayoutNone << 29);  // next 8-cout chunk = next plane (padded pitch)
    constexpr uint32_t b_hi = (16u >> 4) | (1u << 14) | (kLayoutNone << 29);             // next 8-"channel" chunk = the next pixel
    constexpr uint32_t a_lbo = (128u >> 4) << 16;                                        // next 8 pixels (K)
    constexpr uint32_t b_lbo = (128u >> 4) << 16;
    const uint32_t tbase = __shfl_sync(0xffffffffu, tmem_base, 0);
    uint32_t scnt = 0, dcnt = 0;
    bool first = true;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int sd = dcnt % kC8wDy;
      mbar_wait(&d_full[sd], (dcnt / kC8wDy) & 1);
      const uint32_t a_lo0 = ((smem_u32(dys + sd * kDyBytes) >> 4) & 0x3FFFu) | a_lbo;
      // all three depth slabs of the tile first: the MMAs are then issued K-step outermost, so consecutive instructions
      // accumulate into nine DIFFERENT TMEM accumulators. Back-to-back accumulation into the same small (N = 32) accumulator
      // serialises on the MMA pipeline latency: 67 cycles per MMA instead of the 16-cycle floor (measured).
      uint32_t b_lo0[3];
      int sl[3];
#pragma unroll
      for (int kd = 0; kd < 3; ++kd, ++scnt) {
        sl[kd] = scnt % kC8wSlabs;
        mbar_wait(&a_full[sl[kd]], (scnt / kC8wSlabs) & 1);
        b_lo0[kd] = ((smem_u32(slabs + sl[kd] * slab_bytes) >> 4) & 0x3FFFu) | b_lbo;
      }
      fence_proxy_async_smem();
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < (p.debug == 1 ? 0 : kC8wKP / 16); ++ks) {
#pragma unroll
          for (int kd = 0; kd < 3; ++kd) {
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
              // K step ks = pixels 16*ks .. 16*ks+15 of the tile; tap (kd, kh), kw = 0 starts at slab row m + kh*Wp
              const uint32_t a_lo = a_lo0 + static_cast<uint32_t>(ks * 16);
              const uint32_t b_lo = b_lo0[kd] + static_cast<uint32_t>(ks * 16 + kh * Wp);
              umma_bf16(tbase + (kd * 3 + kh) * 32, (static_cast<uint64_t>(a_hi) << 32) | a_lo, (static_cast<uint64_t>(b_hi) << 32) | b_lo, idesc,
                        (first && ks == 0) ? 0u : 1u);
            }
          }
        }
#pragma unroll
        for (int kd = 0; kd < 3; ++kd) umma_commit(&a_empty[sl[kd]]);
        umma_commit(&d_empty[sd]);
      }
      __syncwarp();
      first = false;
      ++dcnt;
    }
    if (elect_one()) umma_commit(done);
    __syncwarp();
  }

  if (warp < 4) {
    // ================================================================= final epilogue: rows 0..31 of each accumulator = cout
    if (ntiles_mine > 0) {
      mbar_wait(done, 0);
      tc_fence_after();
    }
    if (warp < 2) {  // M = 64 accumulator: cout rows 0..15 in TMEM lanes 0..15 (warp 0), rows 16..31 in lanes 32..47 (warp 1)
      for (int acc = 0; acc < 9; ++acc) {
        uint32_t r[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + acc * 32, r);
        tmem_ld_wait();
        if (lane < 16) {
          float* dst = p.partial + (static_cast<long long>(blockIdx.x) * 9 + acc) * 32 * 32 + warp * 16 + lane;  // [n][cout]
#pragma unroll
          for (int j = 0; j < 32; ++j) dst[j * 32] = ntiles_mine > 0 ? __uint_as_float(r[j]) : 0.f;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc<TCOLS>(tmem_base);
}

// dw[cout][cin][kd][kh][kw] (+)= sum over CTAs (fixed order) of partial[cta][kd*3+kh][kw*8 + c][cout]
__global__ void conv3d_c8_wgrad_reduce_kernel(const float* __restrict__ partial, int nparts, float* __restrict__ dw, int cin, int accumulate) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 32 * cin * 27) return;
  const int kw = i % 3, kh = (i / 3) % 3, kd = (i / 9) % 3, c = (i / 27) % cin, co = i / (27 * cin);
  const float* src = partial + ((kd * 3 + kh) * 32 + kw * 8 + c) * 32 + co;
  double acc = 0.0;
  for (int q = 0; q < nparts; ++q) acc += static_cast<double>(src[static_cast<long long>(q) * 9 * 32 * 32]);
  dw[i] = accumulate ? dw[i] + static_cast<float>(acc) : static_cast<float>(acc);
}

}  // namespace qt
