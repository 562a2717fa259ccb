// Persistent 3x3 / stride-1 / pad-1 convolution on tcgen05 with input reuse across the nine taps.
//
// The M index runs over the pixels of a *virtual* zero-padded map [N][H+1][W+2] (row 0 of every image is a zero row that doubles as the bottom padding of the image before it); in that space a filter tap is
// a constant row shift (dh*(W+2) + dw), so one "slab" — the BM + 2(W+3) consecutive virtual pixels around an
// M tile, 64 channels wide — staged ONCE in shared memory serves all nine taps: the A descriptor of tap t just
// starts (dh+1)*(W+2) + (dw+1) rows further down. This cuts the L2->SM traffic of the activation operand ~6x
// against re-gathering every tap (the v1 kernel in igemm.cuh), which is what bounds 3x3 convs on B200.
//
// Shared-memory layouts
//   A slab : no-swizzle K-major "planes": plane c (16-byte channel chunk) holds [rows][16 B]; an 8-row core
//            matrix is 128 contiguous bytes, SBO = 128 B between 8-row groups, LBO = plane stride between the
//            two 16-byte K chunks of one K=16 MMA. Any 16-byte-aligned start row is legal, so shifted windows
//            need no swizzle-phase bookkeeping. Border / out-of-image rows are zero-filled by cp.async.
//   B tile : [BN rows][128 B] with the 128-byte XOR swizzle (same as igemm.cuh), one tile per (tap, slab).
//
// Roles (288 threads, one CTA per SM, persistent over tiles): warps 0-3 epilogue (TMEM -> registers -> global,
// BatchNorm partial sums), warps 4-7 producers (cp.async gathers, lag-published through mbarriers), warp 8
// MMA issuer. Accumulators are double-buffered in TMEM so the epilogue of tile i overlaps the main loop of
// tile i+1; when the whole filter fits in the B ring (64->64 layers) it is loaded once and stays resident.
#pragma once
#include <cuda.h>  // CUtensorMap (type only; the encoder is fetched through cudaGetDriverEntryPoint)

#include "igemm.cuh"

namespace qt {

constexpr int kC3Threads = 352;
constexpr int kNoPlane = 1 << 20;  // depth offset of a slab half that has no source plane (zero filled)  // 4 epilogue + 4 A-slab producer + 2 MMA-issuer warps + 1 TMA (weights) warp

struct Conv3x3Params {
  const __nv_bfloat16* a;   // dense NHWC input  [N][H][W][cin]
  const __nv_bfloat16* b;   // weights [nout][wtaps][cin]
  void* out;                // dense NHWC output [N][H][W][nout] (bf16)
  const __nv_bfloat16* addend;
  float* stats;             // [gridDim.x][2][nout]: one running partial per persistent CTA
  int N, H, W, cin, nout, wtaps;
  int flags;
  int slabs;                // cin / 64
  int V;                    // N * (H+1) * (W+2) virtual pixels
  int num_m_tiles, num_n_tiles;
  int R;                    // slab rows (multiple of 16)
  int b_resident;           // 1: all 9*slabs weight tiles stay in the B ring
  int b_tma;                // 1: weight tiles arrive by TMA (cp.async.bulk.tensor.2d through `wmap`), else cp.async
  signed char off_h[9], off_w[9];
  short wtap[9];
  // 3-D convolutions (Conv3d 3x3x3 / stride 1 / pad 1, 3dcnn/models.py:107-139): N counts depth PLANES (clips * D); a
  // depth tap is a whole-plane shift of the source, zero when it leaves the clip. A tile walks `slabs` = kdn * (cin/64)
  // staged slabs (depth tap major). pair = 1 (cin == 32): one 128-byte slab row holds TWO depth planes of the same pixel
  // (32 channels each), so slab 0 covers depth taps 0 and 1, slab 1 depth tap 2 (+ a zero half) against weights packed
  // [nout][2][9][64] (qt_wpack_conv3d_pair).
  int D;                    // planes per clip (1: plain 2-D convolution)
  int kdn;                  // depth taps: 1 or 3
  int pair;
  int bias_on;              // EPI_BIAS: bias[nout] added before the store / statistics (Conv3d has bias=True)
  const float* bias;
  signed char dplane[4];    // source plane offset of depth tap kd (fprop: kd-1, dgrad: 1-kd)
  signed char wkd[4];       // depth index of that tap inside the weight tensor
};

template <int BN, int MT, int NSLAB, int NB, bool STAGED>
struct C3Smem {
  static constexpr int kBTile = BN * 128;
  static constexpr int kBBytes = NB * kBTile;
  static constexpr int kScratch = 2 * 4 * BN * 4 + 8 * 2 * BN * 4;  // cross-warp combine + running sums (<= 8 n-tiles)
  static constexpr int kStage = STAGED ? 4 * 32 * 64 : 0;  // per epilogue warp: 32 rows x 64 B output staging (coalesced write-out)
  static constexpr int kBarBytes = 1024;  // keeps the slabs 1024-byte aligned
  // slab bytes depend on W (runtime): computed on the host; layout = [B ring][scratch][staging][barriers][slabs...]
};

template <int BN, int MT, int NSLAB, int NB, bool STAGED>
__global__ void __launch_bounds__(kC3Threads, MT == 1 ? 2 : 1) conv3x3_kernel(const __grid_constant__ Conv3x3Params p,
                                                                                const __grid_constant__ CUtensorMap wmap) {
  using L = C3Smem<BN, MT, NSLAB, NB, STAGED>;
  constexpr uint32_t TCOLS = 2 * MT * BN;  // double-buffered accumulators
  static_assert(TCOLS <= 512 && (TCOLS & (TCOLS - 1)) == 0, "TMEM columns");
  static_assert(MT == 1 || MT == 2, "one MMA-issuer warp per sub-tile: warps 8 and 9 (warp 9 idles when MT == 1)");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* b_ring = smem;
  float* scratch = reinterpret_cast<float*>(smem + L::kBBytes);
  float* running = scratch + 2 * 4 * BN;  // [n_tile][2][BN] per-CTA BatchNorm partial sums
  uint8_t* stage = smem + L::kBBytes + L::kScratch;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kBBytes + L::kScratch + L::kStage);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + NSLAB;
  uint64_t* b_full = a_empty + NSLAB;
  uint64_t* b_empty = b_full + NB;
  uint64_t* acc_full = b_empty + NB;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  uint8_t* slab_base = smem + L::kBBytes + L::kScratch + L::kStage + L::kBarBytes;
  const int slab_bytes = (p.R * 128 + 1023) / 1024 * 1024;  // [R rows][64 channels], 16-byte chunks xor-swizzled by row & 7

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (warp == 0) QT_TRACE_GT(8);
  const int Wp = p.W + 2, Hp = p.H + 1;  // one shared zero row between consecutive images
  const int num_tiles = p.num_m_tiles * p.num_n_tiles;

  if (warp == 8) {
    if (lane == 0) {
      // two MMA-issuer warps (one per 128-row sub-tile) release every operand / accumulator stage together
      for (int s = 0; s < NSLAB; ++s) { mbar_init(&a_full[s], kProducerThreads); mbar_init(&a_empty[s], MT); }
      for (int s = 0; s < NB; ++s) { mbar_init(&b_full[s], p.b_tma ? 1 : kProducerThreads); mbar_init(&b_empty[s], MT); }
      for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], MT); mbar_init(&acc_empty[s], kProducerThreads); }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc<TCOLS>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 0) QT_TRACE_GT(9);

  if (warp >= 4 && warp < 8) {
    // ================================================================= producers
    const int t = threadIdx.x - 128;
    const int chunk = t & 7;
    const int rbase = t >> 3;
    const uint32_t b_sw = static_cast<uint32_t>((chunk ^ (rbase & 7)) << 4);
    uint32_t a_cnt = 0, b_cnt = 0;   // items issued so far (ring positions)
    // cp.async completion is tracked by the mbarriers themselves (cp.async.mbarrier.arrive.noinc): producers run
    // ahead by the full depth of the rings and never block in wait_group; the issuers fence after their waits.
    bool first_tile = true;
    QT_TRACE_DECL(tr_a_empty);
    QT_TRACE_T0(tr_p0);
    const int adv_w = 16 % Wp, adv_h = 16 / Wp;
    // kernel parameters used in the gather loop live in registers (the cp.async asm has a memory clobber, which would
    // otherwise make the compiler re-read them from the constant bank every iteration)
    const int pW = p.W, pH = p.H, pN = p.N, pR = p.R, pD = p.D;
    const long long pix_bytes = static_cast<long long>(p.cin) * 2;
    const long long row_bytes = pW * pix_bytes;
    const long long plane_bytes = static_cast<long long>(pH) * row_bytes;
    const long long adv_off = (static_cast<long long>(adv_h) * pW + adv_w) * pix_bytes;
    const void* dummy = p.a;
    const int cs = p.pair ? 1 : p.cin / 64;  // 64-channel chunks per depth tap
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_tile = tile / p.num_n_tiles;
      const int n0 = (tile - m_tile * p.num_n_tiles) * BN;
      const int q0 = m_tile * (kBM * MT);
      for (int c = 0; c < p.slabs; ++c) {
        {  // ---- A slab: rows j <-> virtual pixel q0 - (W+3) + j
          const int s = a_cnt % NSLAB;
          if (a_cnt >= NSLAB) QT_TRACE_WAIT(tr_a_empty, mbar_wait(&a_empty[s], ((a_cnt / NSLAB) - 1) & 1));
          // depth tap of this slab and the channel chunk inside it (2-D: kd = 0, dd = 0). Pair mode: this thread's 16-byte
          // chunk column decides which of the slab's two planes it copies (chunks 0-3: first plane, 4-7: second plane).
          int dd, ch_off;
          if (p.pair) {
            const int kd = 2 * c + (chunk >> 2);
            dd = kd < 3 ? p.dplane[kd] : kNoPlane;  // the zero half of the second pair slab
            ch_off = (chunk & 3) * 8;
          } else {
            const int kd = c / cs;
            dd = p.dplane[kd];
            ch_off = (c - kd * cs) * 64 + chunk * 8;
          }
          const __nv_bfloat16* src_c = p.a + ch_off;
          // first row of this thread: virtual pixel v0 (shifted by one image so the decomposition is non-negative),
          // then advance 16 virtual pixels per iteration with carries instead of dividing per row
          int n, hp, wp, dz;  // n: plane index (image for 2-D), dz: its depth inside the clip
          {
            const int vv = q0 - (pW + 3) + rbase + Wp * Hp;
            wp = vv % Wp;
            const int rest = vv / Wp;
            hp = rest % Hp;
            n = rest / Hp - 1;
            dz = (n + pD) % pD;
          }
          // byte address of real pixel (n, hp-1, wp-1), kept incrementally (only dereferenced when the row is real):
          // +16 virtual pixels = adv_off; a w-carry skips the 2 pad columns, an h-carry the shared zero row.
          // Addresses are produced in batches of kBatch into distinct registers and only then handed to cp.async:
          // the LSU releases a cp.async's address registers late, so reusing one register pair per copy would
          // serialise the copies on that release.
          const char* srcb = reinterpret_cast<const char*>(src_c) + (dd == kNoPlane ? 0 : dd) * plane_bytes +
                             (static_cast<long long>(n * pH + (hp - 1)) * pW + (wp - 1)) * pix_bytes;
          // slab row j holds virtual pixel q0 - (W+3) + j; a thread's rows are 16 apart, so its swizzle phase is fixed
          uint32_t dst = smem_u32(slab_base + s * slab_bytes) + rbase * 128 + ((chunk ^ (rbase & 7)) << 4);
          constexpr uint32_t dstep = 2048u;  // 16 rows further
          constexpr int kBatch = 4;
          for (int j = rbase; j < pR; j += 16 * kBatch) {
            const void* sp[kBatch];
            uint32_t sz[kBatch];
#pragma unroll
            for (int b = 0; b < kBatch; ++b) {
              const bool ok = (static_cast<unsigned>(n) < static_cast<unsigned>(pN)) &&
                              (static_cast<unsigned>(wp - 1) < static_cast<unsigned>(pW)) && (hp >= 1) &&
                              (static_cast<unsigned>(dz + dd) < static_cast<unsigned>(pD));
              sp[b] = ok ? static_cast<const void*>(srcb) : dummy;
              sz[b] = ok ? 16u : 0u;
              srcb += adv_off;
              wp += adv_w; hp += adv_h;
              if (wp >= Wp) { wp -= Wp; ++hp; srcb -= 2 * pix_bytes; }
              if (hp >= Hp) { hp -= Hp; ++n; srcb -= row_bytes; dz = (dz + 1 == pD) ? 0 : dz + 1; }
            }
#pragma unroll
            for (int b = 0; b < kBatch; ++b)
              if (j + 16 * b < pR) cp_async16(dst + b * dstep, sp[b], sz[b]);
            dst += dstep * kBatch;
          }
          cp_async_mbar_arrive_noinc(&a_full[s]);
          ++a_cnt;
        }
        if (!p.b_tma && (!p.b_resident || first_tile)) {
          for (int tp = 0; tp < 9; ++tp) {
            const int s = b_cnt % NB;
            if (!p.b_resident && b_cnt >= NB) mbar_wait(&b_empty[s], ((b_cnt / NB) - 1) & 1);
            const uint32_t dst0 = smem_u32(b_ring + s * L::kBTile) + rbase * 128 + b_sw;
            const int kd_w = p.pair ? 0 : c / cs;
            const long long woff = p.pair ? (static_cast<long long>(c) * 9 + p.wtap[tp]) * 64 + chunk * 8
                                          : (static_cast<long long>(p.wkd[kd_w]) * 9 + p.wtap[tp]) * p.cin + (c - kd_w * cs) * 64 + chunk * 8;
#pragma unroll
            for (int i = 0; i < BN / 16; ++i) {
              const int n = n0 + rbase + 16 * i;
              const bool ok = n < p.nout;
              const __nv_bfloat16* src = ok ? (p.b + static_cast<long long>(n) * (p.pair ? 2 * 9 * 64 : p.wtaps * p.cin) + woff) : p.b;
              cp_async16(dst0 + i * 16 * 128, src, ok ? 16u : 0u);
            }
            cp_async_mbar_arrive_noinc(&b_full[s]);
            ++b_cnt;
          }
        }
      }
      first_tile = false;
    }
    cp_async_wait<0>();  // nothing may be in flight when the CTA retires
    if (warp == 4) { QT_TRACE_PUT(6, tr_a_empty); QT_TRACE_PUT(7, QT_TRACE_NOW() - tr_p0); }
  } else if (warp == 10) {
    // ================================================================= TMA producer for the weight tiles
    // box = [BN rows][64 K-elements] of the 2-D weight matrix [nout][9*cin], 128B-swizzled by the TMA unit —
    // byte-identical to what the cp.async path writes. One instruction per (tap, slab) tile.
    if (p.b_tma && lane == 0) {
      uint32_t b_cnt = 0;
      bool first_tile = true;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_tile = tile / p.num_n_tiles;
        const int n0 = (tile - m_tile * p.num_n_tiles) * BN;
        if (!p.b_resident || first_tile) {
          for (int c = 0; c < p.slabs; ++c) {
            for (int tp = 0; tp < 9; ++tp) {
              const int s = b_cnt % NB;
              if (!p.b_resident && b_cnt >= NB) mbar_wait(&b_empty[s], ((b_cnt / NB) - 1) & 1);
              mbar_arrive_expect_tx(&b_full[s], L::kBTile);
              const int cs_w = p.pair ? 1 : p.cin / 64;
              const int kd_w = p.pair ? 0 : c / cs_w;
              const int kx = p.pair ? (c * 9 + p.wtap[tp]) * 64 : (p.wkd[kd_w] * 9 + p.wtap[tp]) * p.cin + (c - kd_w * cs_w) * 64;
              asm volatile(
                  "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(
                      smem_u32(b_ring + s * L::kBTile)),
                  "l"(reinterpret_cast<uint64_t>(&wmap)), "r"(kx), "r"(n0), "r"(smem_u32(&b_full[s]))
                  : "memory");
              ++b_cnt;
            }
          }
        }
        first_tile = false;
      }
    }
    __syncwarp();
  } else if (warp >= 8 + MT) {
    // idle issuer warp (MT == 1)
  } else if (warp >= 8) {
    // ================================================================= MMA issuers (warp 8 + u owns sub-tile u)
    // The whole warp runs the (warp-uniform) control flow so descriptor arithmetic stays on the uniform
    // datapath; one elected lane issues tcgen05.mma / tcgen05.commit.
    {
      constexpr uint32_t idesc = make_idesc_bf16(kBM, BN, 0, 0);
      // descriptor high words are constant: SBO>>4 [0,14) | version 1 [14,16) | layout [29,32)
      // Both operands are K-major with the 128-byte swizzle (SBO = 1024: next 8 rows). The hardware applies the
      // swizzle to the generated ADDRESS bits, so an A descriptor may start at any slab row (start = slab + row*128,
      // base-offset field 0): that is how one slab serves all nine taps.
      constexpr uint32_t b_hi = (1024u >> 4) | (1u << 14) | (kLayoutSW128 << 29);
      constexpr uint32_t a_hi = b_hi;
      constexpr uint32_t a_lbo = 1u << 16;
      constexpr uint32_t b_lbo = 1u << 16;
      constexpr uint32_t kstep = 2;  // 32 bytes (K = 16) inside the 128-byte row
      const uint32_t tbase = __shfl_sync(0xffffffffu, tmem_base, 0);
      uint32_t a_cnt = 0, b_cnt = 0, tile_it = 0;
      bool first_tile = true;
      QT_TRACE_DECL(tr_acc_empty); QT_TRACE_DECL(tr_a_full); QT_TRACE_DECL(tr_b_full);
      QT_TRACE_T0(tr_m0);
      uint32_t tapoff[9];  // slab row of each filter tap relative to the tile's first pixel, in 16-byte descriptor units
#pragma unroll
      for (int tp = 0; tp < 9; ++tp) tapoff[tp] = static_cast<uint32_t>((p.off_h[tp] + 1) * Wp + (p.off_w[tp] + 1)) * 8;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tile_it) {
        const uint32_t ab = tile_it & 1;
        if (tile_it >= 2) QT_TRACE_WAIT(tr_acc_empty, mbar_wait(&acc_empty[ab], ((tile_it >> 1) - 1) & 1));
        tc_fence_after();
        const uint32_t d_base = tbase + ab * (MT * BN);
        for (int c = 0; c < p.slabs; ++c) {
          const int sa = a_cnt % NSLAB;
          QT_TRACE_WAIT(tr_a_full, mbar_wait(&a_full[sa], (a_cnt / NSLAB) & 1));
          fence_proxy_async_smem();
          tc_fence_after();
          const uint32_t slab_lo = ((smem_u32(slab_base + sa * slab_bytes) >> 4) & 0x3FFFu) | a_lbo;
          if (p.b_resident && !first_tile) {
            // resident filter, already waited for and fenced on the first tile: nine taps back to back
            const uint32_t b0 = (((smem_u32(b_ring) >> 4) + c * 9 * (L::kBTile >> 4)) & 0x3FFFu) | b_lbo;
            if (elect_one()) {
              const int u = warp - 8;
#pragma unroll
              for (int tp = 0; tp < 9; ++tp) {
                const uint32_t a_lo = slab_lo + tapoff[tp] + u * (kBM * 8);
                const uint32_t b_lo = b0 + tp * (L::kBTile >> 4);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const uint64_t ad = (static_cast<uint64_t>(a_hi) << 32) | (a_lo + k * kstep);
                  const uint64_t bd = (static_cast<uint64_t>(b_hi) << 32) | (b_lo + k * 2);
                  umma_bf16(d_base + u * BN, ad, bd, idesc, (c | tp | k) ? 1u : 0u);
                }
              }
            }
            __syncwarp();
          } else {
#pragma unroll 1
            for (int tp = 0; tp < 9; ++tp) {
              int sb;
              if (p.b_resident) {
                sb = c * 9 + tp;
                mbar_wait(&b_full[sb], 0);
              } else {
                sb = b_cnt % NB;
                QT_TRACE_WAIT(tr_b_full, mbar_wait(&b_full[sb], (b_cnt / NB) & 1));
              }
              if (!p.b_tma) fence_proxy_async_smem();  // cp.async-written filter tile (TMA writes are async-proxy already)
              tc_fence_after();
              const uint32_t b_lo = ((smem_u32(b_ring + sb * L::kBTile) >> 4) & 0x3FFFu) | b_lbo;
              const uint32_t a_lo = slab_lo + (static_cast<uint32_t>((p.off_h[tp] + 1) * Wp + (p.off_w[tp] + 1)) + (warp - 8) * kBM) * 8;
              if (elect_one()) {
                const int u = warp - 8;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const uint64_t ad = (static_cast<uint64_t>(a_hi) << 32) | (a_lo + k * kstep);
                  const uint64_t bd = (static_cast<uint64_t>(b_hi) << 32) | (b_lo + k * 2);
                  umma_bf16(d_base + u * BN, ad, bd, idesc, (c | tp | k) ? 1u : 0u);
                }
                if (!p.b_resident) umma_commit(&b_empty[sb]);
              }
              __syncwarp();
              if (!p.b_resident) ++b_cnt;
            }
          }
          if (elect_one()) umma_commit(&a_empty[sa]);
          __syncwarp();
          ++a_cnt;
        }
        if (elect_one()) umma_commit(&acc_full[ab]);
        __syncwarp();
        first_tile = false;
      }
      if (warp == 8) { QT_TRACE_GT(10); QT_TRACE_PUT(0, tr_acc_empty); QT_TRACE_PUT(1, tr_a_full); QT_TRACE_PUT(2, tr_b_full); QT_TRACE_PUT(3, QT_TRACE_NOW() - tr_m0); }
    }
  } else {
    // ================================================================= epilogue (warps 0-3)
    const int flags = p.flags;
    uint32_t tile_it = 0;
    QT_TRACE_DECL(tr_acc_full);
    QT_TRACE_T0(tr_e0);
    if (flags & EPI_STATS) {
      for (int i = threadIdx.x; i < p.num_n_tiles * 2 * BN; i += kProducerThreads) running[i] = 0.f;
      asm volatile("bar.sync 1, 128;\n" ::: "memory");
    }
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tile_it) {
      const int m_tile = tile / p.num_n_tiles;
      const int n_tile = tile - m_tile * p.num_n_tiles;
      const int n0 = n_tile * BN;
      const uint32_t ab = tile_it & 1;
      QT_TRACE_WAIT(tr_acc_full, mbar_wait(&acc_full[ab], (tile_it >> 1) & 1));
      tc_fence_after();
      bool row_ok[MT];
      long long orow[MT];
#pragma unroll
      for (int u = 0; u < MT; ++u) {
        const int v = (m_tile * MT + u) * kBM + warp * 32 + lane;
        row_ok[u] = v < p.V;
        orow[u] = 0;
        if (row_ok[u]) {
          const int wp = v % Wp;
          const int rest = v / Wp;
          const int hp = rest % Hp;
          const int n = rest / Hp;
          row_ok[u] = (wp >= 1) && (wp <= p.W) && (hp >= 1) && (hp <= p.H);
          orow[u] = ((static_cast<long long>(n) * p.H + (hp - 1)) * p.W + (wp - 1)) * p.nout;
        }
      }
      // Coalesced write-out: every 32x32 chunk goes through a per-warp staging tile so that one store instruction
      // covers 8 rows x 64 contiguous bytes (consecutive virtual pixels are consecutive real pixels inside an image
      // row) instead of 32 rows x 16 bytes - the scattered form keeps the LSU busy for 32 tag cycles per instruction
      // and starves the producers' cp.async (measured: epilogue + gather together cost 2x either alone).
      // srow[u][i]: output offset of row 8*i + lane/4 (the rows this lane writes out), -1 outside the image.
      long long srow[MT][4];
      if (STAGED) {
#pragma unroll
        for (int u = 0; u < MT; ++u) {
          const long long mine = row_ok[u] ? orow[u] : -1;
#pragma unroll
          for (int i = 0; i < 4; ++i) srow[u][i] = __shfl_sync(0xffffffffu, mine, 8 * i + (lane >> 2));
        }
      }
      const uint32_t stage_w = smem_u32(stage) + warp * 2048;
      const uint32_t st_own = stage_w + lane * 64, st_sw = (lane >> 1) & 3;
      const uint32_t ld_row = stage_w + (lane >> 2) * 64 + (((lane & 3) ^ ((lane >> 3) & 3)) << 4);
      // column-sum reads: row 2k + (lane >> 4), 4-byte word (lane & 15); that row's 16-byte units are xor-swizzled by
      // (row >> 1) & 3 = k & 3
      const uint32_t cs_addr = stage_w + (lane >> 4) * 64 + (lane & 3) * 4;
      // chunk-outer / sub-tile-inner: the per-column sums of the MT sub-tiles are added in registers first, so the
      // 32x32 transpose-reduce (the expensive part: 62 shuffles per thread) runs once per chunk, not once per sub-tile
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        const int ncol = n0 + c0;
        if (ncol >= p.nout) break;
        float s1v[32], s2v[32];
        float cs1[2] = {0.f, 0.f}, cs2[2] = {0.f, 0.f};
#pragma unroll
        for (int u = 0; u < MT; ++u) {
          uint32_t r[32];
          tmem_ld32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + ab * (MT * BN) + u * BN + c0, r);
          tmem_ld_wait();
          float vv[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) vv[j] = __uint_as_float(r[j]);
          if ((flags & EPI_ADDEND) && STAGED) {
            // coalesced read of the addend chunk (8 rows x 64 B per instruction) through the staging tile
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint4 q = make_uint4(0, 0, 0, 0);
              if (srow[u][i] >= 0) q = *reinterpret_cast<const uint4*>(p.addend + srow[u][i] + ncol + (lane & 3) * 8);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(ld_row + i * 512), "r"(q.x), "r"(q.y), "r"(q.z), "r"(q.w) : "memory");
            }
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint32_t w4[4];
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];\n" : "=r"(w4[0]), "=r"(w4[1]), "=r"(w4[2]), "=r"(w4[3]) : "r"(st_own + ((j ^ st_sw) << 4)) : "memory");
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                vv[8 * j + 2 * e] += __uint_as_float(w4[e] << 16);
                vv[8 * j + 2 * e + 1] += __uint_as_float(w4[e] & 0xffff0000u);
              }
            }
            __syncwarp();
          } else if ((flags & EPI_ADDEND) && row_ok[u]) {
            const __nv_bfloat16* ad = p.addend + orow[u] + ncol;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              const uint4 q = *reinterpret_cast<const uint4*>(ad + j);
              const uint32_t w4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                vv[j + 2 * e] += __uint_as_float(w4[e] << 16);
                vv[j + 2 * e + 1] += __uint_as_float(w4[e] & 0xffff0000u);
              }
            }
          }
          if (p.bias_on) {  // Conv3d(bias=True): same value for every row of the tile (broadcast loads)
#pragma unroll
            for (int j = 0; j < 32; ++j) vv[j] += __ldg(p.bias + ncol + j);
          }
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2(vv[2 * j], vv[2 * j + 1]);
          if (STAGED) {
            if ((flags & EPI_STATS) && !row_ok[u]) {
#pragma unroll
              for (int j = 0; j < 16; ++j) pk[j] = 0u;  // rows outside the image contribute zero to the statistics
            }
  #pragma unroll
            for (int j = 0; j < 4; ++j)
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(st_own + ((j ^ st_sw) << 4)), "r"(pk[4 * j]),
                           "r"(pk[4 * j + 1]), "r"(pk[4 * j + 2]), "r"(pk[4 * j + 3])
                           : "memory");
            __syncwarp();
  #pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint4 q;
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];\n" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "r"(ld_row + i * 512) : "memory");
              if (srow[u][i] >= 0)
                *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + srow[u][i] + ncol + (lane & 3) * 8) = q;
            }
            if (flags & EPI_STATS) {
              // column sums straight from the staged bf16 tile: lane l owns the column pair (l & 15) and every other
              // row (parity l >> 4): 16 conflict-free 4-byte reads instead of two 32x32 shuffle transposes
#pragma unroll
              for (int k = 0; k < 16; ++k) {
                uint32_t w;
                asm volatile("ld.shared.b32 %0, [%1];\n" : "=r"(w) : "r"(cs_addr + k * 128 + ((((lane >> 2) & 3) ^ (k & 3)) << 4)) : "memory");
                const float lo = __uint_as_float(w << 16), hi = __uint_as_float(w & 0xffff0000u);
                cs1[0] += lo; cs1[1] += hi;
                cs2[0] = fmaf(lo, lo, cs2[0]); cs2[1] = fmaf(hi, hi, cs2[1]);
              }
            }
            __syncwarp();
          } else if (row_ok[u]) {
            __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.out) + orow[u] + ncol;
#pragma unroll
            for (int j = 0; j < 4; ++j)
              *reinterpret_cast<uint4*>(dst + 8 * j) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
          }
          if (!STAGED && (flags & EPI_STATS)) {
            // statistics of the stored (bf16-rounded) values; rows outside the image contribute zero
            const uint32_t keep = row_ok[u] ? 0xffffffffu : 0u;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float lo = __uint_as_float((pk[j] << 16) & keep), hi = __uint_as_float(pk[j] & 0xffff0000u & keep);
              if (u == 0) {
                s1v[2 * j] = lo; s1v[2 * j + 1] = hi;
                s2v[2 * j] = lo * lo; s2v[2 * j + 1] = hi * hi;
              } else {
                s1v[2 * j] += lo; s1v[2 * j + 1] += hi;
                s2v[2 * j] = fmaf(lo, lo, s2v[2 * j]); s2v[2 * j + 1] = fmaf(hi, hi, s2v[2 * j + 1]);
              }
            }
          }
        }
        if (STAGED && (flags & EPI_STATS)) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            cs1[e] += __shfl_xor_sync(0xffffffffu, cs1[e], 16);
            cs2[e] += __shfl_xor_sync(0xffffffffu, cs2[e], 16);
          }
          if (lane < 16) {
            *reinterpret_cast<float2*>(&scratch[(0 * 4 + warp) * BN + c0 + 2 * lane]) = make_float2(cs1[0], cs1[1]);
            *reinterpret_cast<float2*>(&scratch[(1 * 4 + warp) * BN + c0 + 2 * lane]) = make_float2(cs2[0], cs2[1]);
          }
        } else if (flags & EPI_STATS) {
          const float s1 = warp_transpose_reduce(s1v);
          const float s2 = warp_transpose_reduce(s2v);
          scratch[(0 * 4 + warp) * BN + c0 + lane] = s1;
          scratch[(1 * 4 + warp) * BN + c0 + lane] = s2;
        }
      }
      // every tcgen05.ld of this accumulator has completed: hand it back before the cross-warp combine
      tc_fence_before();
      mbar_arrive(&acc_empty[ab]);
      if (flags & EPI_STATS) {
        asm volatile("bar.sync 1, 128;\n" ::: "memory");
        for (int i = threadIdx.x; i < 2 * BN; i += kProducerThreads) {
          const int which = i / BN, col = i - which * BN;
          if (n0 + col < p.nout) {
            const float* sc = scratch + which * 4 * BN + col;
            running[(n_tile * 2 + which) * BN + col] += (sc[0] + sc[BN]) + (sc[2 * BN] + sc[3 * BN]);
          }
        }
        asm volatile("bar.sync 1, 128;\n" ::: "memory");
      }
    }
    if (warp == 0) { QT_TRACE_GT(11); QT_TRACE_PUT(4, tr_acc_full); QT_TRACE_PUT(5, QT_TRACE_NOW() - tr_e0); }
    if (flags & EPI_STATS) {
      // one deterministic partial row per CTA: [2][nout]
      for (int i = threadIdx.x; i < p.num_n_tiles * 2 * BN; i += kProducerThreads) {
        const int nt = i / (2 * BN), rem = i - nt * 2 * BN;
        const int which = rem / BN, col = rem - which * BN;
        if (nt * BN + col < p.nout)
          p.stats[(static_cast<long long>(blockIdx.x) * 2 + which) * p.nout + nt * BN + col] = running[i];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc<TCOLS>(tmem_base);
  if (warp == 8) QT_TRACE_GT(12);
}

}  // namespace qt
