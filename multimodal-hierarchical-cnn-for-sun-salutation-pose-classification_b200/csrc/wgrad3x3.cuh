// Weight gradient of 3x3 / stride-1 / pad-1 convolutions with input reuse across the taps (the backward twin of
// conv3x3.cuh):   dW[tap][cin][cout] = sum over pixels  x[pixel + shift(tap)][cin] * dy[pixel][cout].
//
// The reduction (GEMM K) runs over the virtual zero-padded pixel space [N][H+1][W+2] in tiles of 128 pixels. Per
// tile the producers stage
//   * the x "slab" (128 + 2(W+3) virtual pixels x 64*NSLAB channels) as no-swizzle planes [8-channel chunk][pixel][16 B]
//     — used as the MN-major B operand (N = channels): chunk stride = SBO, 16 B between the pixels of a core
//     matrix, LBO = 128 B between 8-pixel groups; a tap is just a different start pixel, so ONE slab feeds all taps;
//   * the dy tile (128 virtual pixels x 64*CB couts, zero rows for border pixels) in the 128B-swizzled MN-major
//     layout — the A operand (M = cout; for 64-cout layers the upper 64 rows point at a zero block).
// Each CTA owns TAPS accumulators [128 cout x 64*NSLAB cin] in TMEM for its whole pixel range (split-K over
// pixel ranges), then writes fp32 partials in the workspace layout of splitk_reduce_wgrad_kernel.
// Roles (192 threads): warps 0-3 producers (and final epilogue), warps 4 and 5 MMA issuers (even / odd taps of the
// CTA's tap group: distinct accumulators, so two threads feed the tensor core's queue and the single-thread
// issue rate stops being the limiter).
#pragma once
#include "igemm.cuh"

namespace qt {

constexpr int kW3KP = 128;  // virtual pixels per k-tile

struct Wgrad3x3Params {
  const __nv_bfloat16* x;    // dense NHWC [N][H][W][cin]
  const __nv_bfloat16* dy;   // dense NHWC [N][H][W][cout]
  float* ws;                 // [splits][Mpad = 9*cin][Npad = cout]
  int N, H, W, cin, cout;
  int V, num_kt, kt_per_split;
  int R;                     // slab rows (multiple of 8)
  int cout_tiles, cin_groups, tap_groups;
  signed char off_h[9], off_w[9];
  // Conv3d (3x3x3 / s1 / p1): N counts depth planes (clips * D); blockIdx.z = depth tap kd, whose x slab is the plane
  // kd - 1 away (zero outside the clip); partials land at tap index kd * 9 + tap of the [27 * cin][cout] workspace.
  int D, kdn;
  // pair3d (Conv3d with 32 input channels, 64 outputs): the 64 slab "channels" are two depth planes of 32 channels each
  // (blockIdx.z = pair item: planes (-1, 0) and (+1, none)), as in conv3x3's pair mode; accumulator columns 0..31 / 32..63
  // belong to depth taps 2*item / 2*item + 1.
  int pair3d;
};

template <int NSLAB, int TAPS, int CB, int STAGES, int NMMA>
__global__ void __launch_bounds__(192, 1) wgrad3x3_kernel(const __grid_constant__ Wgrad3x3Params p) {
  constexpr int NCH = 64 * NSLAB;          // channels (GEMM N) per CTA
  constexpr uint32_t TCOLS = 512;
  static_assert(TAPS * NCH <= 512, "TMEM budget");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // PAIR (64 -> 64 layers, CB == 1): the dy tile carries 8 extra pixel rows and the A descriptor's second 64-row atom
  // starts ONE pixel row (128 B) after the first, so accumulator rows 64..127 see dy[p+1]: with the x slab started at
  // tap (dh,+1) one MMA yields the gradients of taps (dh,+1) [rows 0..63] and (dh,0) [rows 64..127]. Six MMA groups
  // (3 pairs + the 3 single dw=-1 taps) replace nine half-empty ones and fit one CTA, so nothing is gathered twice.
  constexpr bool PAIR = (CB == 1);
  constexpr int kDyRows = PAIR ? kW3KP + 8 : kW3KP;
  const int dy_bytes = CB * kDyRows * 128;                      // CB blocks of [pixels][128 B]
  const int xblk_bytes = (p.R * 128 + 1023) / 1024 * 1024;     // one 64-channel block of the slab: [R rows][128 B], 128-byte swizzle
  const int slab_bytes = NSLAB * xblk_bytes;
  const int stage_bytes = dy_bytes + slab_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = full + STAGES;
  uint64_t* done = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int Wp = p.W + 2, Hp = p.H + 1;  // one shared zero row between consecutive images
  // CTA coordinates
  int type = blockIdx.x;
  const int tap_group = type % p.tap_groups; type /= p.tap_groups;
  const int cin_group = type % p.cin_groups;
  const int cout_tile = type / p.cin_groups;
  const int tap0 = tap_group * TAPS;
  const int ntap = PAIR ? TAPS : min(TAPS, 9 - tap0);  // PAIR: TAPS == 6 MMA groups, one tap group
  const int cin0 = cin_group * NCH;
  const int cout0 = cout_tile * (64 * CB);
  const int kt0 = blockIdx.y * p.kt_per_split;
  const int kt1 = min(p.num_kt, kt0 + p.kt_per_split);
  const int nit = max(0, kt1 - kt0);

  if (warp == 4) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], kProducerThreads); mbar_init(&empty[s], NMMA); }
      mbar_init(done, NMMA);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc<TCOLS>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= 4 + NMMA) {
    // spare warp (single-issuer configuration)
  } else if (warp < 4) {
    // ------------------------------------------------------------------ producers
    const int t = threadIdx.x;
    const int chunk = t & 7;
    const int rbase = t >> 3;
    const uint32_t dsw = static_cast<uint32_t>((chunk ^ (rbase & 7)) << 4);
    const int adv_w = 16 % Wp, adv_h = 16 / Wp;
    // loop-invariant kernel parameters in registers (the cp.async asm carries a memory clobber); real-pixel byte
    // offsets are kept incrementally: +16 virtual pixels = adv, a w-carry skips the 2 pad columns, an h-carry the
    // shared zero row
    const int pW = p.W, pN = p.N, pR = p.R, pD = p.D;
    // source plane offset of this thread's slab chunk: the CTA's depth tap, or (pair3d) one of the item's two planes
    int dd = p.kdn > 1 ? static_cast<int>(blockIdx.z) - 1 : 0;
    if (p.pair3d) {
      const int kd = 2 * static_cast<int>(blockIdx.z) + (chunk >> 2);
      dd = kd < 3 ? kd - 1 : (1 << 20);
    }
    const long long dy_pix = static_cast<long long>(p.cout) * 2, x_pix = p.pair3d ? 64 : static_cast<long long>(p.cin) * 2;
    const long long dy_adv = (static_cast<long long>(adv_h) * pW + adv_w) * dy_pix, x_adv = (static_cast<long long>(adv_h) * pW + adv_w) * x_pix;
    const long long dy_row = pW * dy_pix, x_row = pW * x_pix;
    const char* dy_c = reinterpret_cast<const char*>(p.dy + cout0 + chunk * 8);
    const char* x_c = reinterpret_cast<const char*>(p.x + (p.pair3d ? (chunk & 3) * 8 : cin0 + chunk * 8));
    const void* dummy_dy = p.dy;
    const void* dummy_x = p.x;
    for (int it = 0; it < nit; ++it) {
      const int s = it % STAGES;
      if (it >= STAGES) mbar_wait(&empty[s], ((it / STAGES) - 1) & 1);
      uint8_t* st = smem + s * stage_bytes;
      const int q0 = (kt0 + it) * kW3KP;
      {  // ---- dy tile: row i <-> virtual pixel q0 + i (zero for border / out-of-range pixels)
        int n, hp, wp;
        {
          const int vv = q0 + rbase;
          wp = vv % Wp;
          const int rest = vv / Wp;
          hp = rest % Hp;
          n = rest / Hp;
        }
        const char* src = dy_c + (static_cast<long long>(n * p.H + (hp - 1)) * pW + (wp - 1)) * dy_pix;
        const uint32_t dst0 = smem_u32(st) + dsw + rbase * 128;
#pragma unroll
        for (int i = 0; i < (kDyRows + 15) / 16; ++i) {
          const bool ok = (n < pN) && (static_cast<unsigned>(wp - 1) < static_cast<unsigned>(pW)) && (hp >= 1);
          if (rbase + 16 * i < kDyRows) {
#pragma unroll
            for (int b = 0; b < CB; ++b)
              cp_async16(dst0 + b * (kDyRows * 128) + i * 2048, ok ? static_cast<const void*>(src + b * 128) : dummy_dy, ok ? 16u : 0u);
          }
          src += dy_adv;
          wp += adv_w; hp += adv_h;
          if (wp >= Wp) { wp -= Wp; ++hp; src -= 2 * dy_pix; }
          if (hp >= Hp) { hp -= Hp; ++n; src -= dy_row; }
        }
      }
      {  // ---- x slab: row j <-> virtual pixel q0 - (W+3) + j
        int n, hp, wp, dz;
        {
          const int vv = q0 - (pW + 3) + rbase + Wp * Hp;
          wp = vv % Wp;
          const int rest = vv / Wp;
          hp = rest % Hp;
          n = rest / Hp - 1;
          dz = (n + pD) % pD;
        }
        const char* src = x_c + (dd == (1 << 20) ? 0 : dd) * (static_cast<long long>(p.H) * x_row) +
                          (static_cast<long long>(n * p.H + (hp - 1)) * pW + (wp - 1)) * x_pix;
        uint32_t dst = smem_u32(st + dy_bytes) + rbase * 128 + dsw;  // rows 16 apart keep the swizzle phase (row & 7)
#pragma unroll 2
        for (int j = rbase; j < pR; j += 16) {
          const bool ok = (static_cast<unsigned>(n) < static_cast<unsigned>(pN)) &&
                          (static_cast<unsigned>(wp - 1) < static_cast<unsigned>(pW)) && (hp >= 1) &&
                          (static_cast<unsigned>(dz + dd) < static_cast<unsigned>(pD));
#pragma unroll
          for (int sl = 0; sl < NSLAB; ++sl)
            cp_async16(dst + sl * xblk_bytes, ok ? static_cast<const void*>(src + sl * 128) : dummy_x, ok ? 16u : 0u);
          dst += 2048;
          src += x_adv;
          wp += adv_w; hp += adv_h;
          if (wp >= Wp) { wp -= Wp; ++hp; src -= 2 * x_pix; }
          if (hp >= Hp) { hp -= Hp; ++n; src -= x_row; dz = (dz + 1 == pD) ? 0 : dz + 1; }
        }
      }
      // completion is tracked by the mbarrier itself (no wait_group): the producers run ahead by up to STAGES
      // tiles; the consumer issues the generic->async proxy fence after its barrier wait.
      cp_async_mbar_arrive_noinc(&full[s]);
    }
    cp_async_wait<0>();  // nothing may be in flight when the CTA retires
  } else {
    // ------------------------------------------------------------------ MMA issuers (warp 4: even taps, warp 5: odd taps)
    const int tpar = warp - 4;
    constexpr uint32_t idesc = make_idesc_bf16(kBM, NCH, 1, 1);
    constexpr uint32_t a_hi = (1024u >> 4) | (1u << 14) | (kLayoutSW128 << 29);   // SBO: next 8 pixels of dy
    // x slab: MN-major SW128 like the dy tile (SBO: next 8 pixels, LBO: next 64-channel block). The swizzle acts on
    // address bits, so the descriptor may start at any slab row: start = slab + (tap row) * 128.
    constexpr uint32_t b_hi = a_hi;
    const uint32_t b_lbo = ((static_cast<uint32_t>(xblk_bytes) >> 4) & 0x3FFFu) << 16;
    const uint32_t tbase = __shfl_sync(0xffffffffu, tmem_base, 0);
    // per-tap start rows (in 16-byte units) kept in registers; the 8 x TAPS MMAs of a k-tile are fully unrolled
    uint32_t row0[TAPS];
#pragma unroll
    for (int tp = 0; tp < TAPS; ++tp) {
      // PAIR: group 2*kh starts the slab at tap (kh, kw = 2) [pair with (kh, 1)], group 2*kh + 1 at tap (kh, kw = 0)
      const int tsel = PAIR ? (tp >> 1) * 3 + ((tp & 1) ? 0 : 2) : min(tap0 + tp, 8);
      row0[tp] = static_cast<uint32_t>((p.off_h[tsel] + 1) * Wp + (p.off_w[tsel] + 1));
    }
    for (int it = 0; it < nit; ++it) {
      const int s = it % STAGES;
      mbar_wait(&full[s], (it / STAGES) & 1);
      fence_proxy_async_smem();
      tc_fence_after();
      uint8_t* st = smem + s * stage_bytes;
      const uint32_t a_addr = smem_u32(st);
      const uint32_t a_lbo_bytes = (CB == 2) ? static_cast<uint32_t>(kW3KP * 128) : 128u;  // PAIR: rows 64..127 <- next pixel
      const uint32_t a_lo = ((a_addr >> 4) & 0x3FFFu) | (((a_lbo_bytes >> 4) & 0x3FFFu) << 16);
      const uint32_t x_lo = ((smem_u32(st + dy_bytes) >> 4) & 0x3FFFu) | b_lbo;
      const uint32_t acc = it ? 1u : 0u;
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < kW3KP / 16; ++ks) {
#pragma unroll
          for (int tp = 0; tp < TAPS; ++tp) {
            if ((NMMA == 1 || (tp & 1) == tpar) && tp < ntap) {
              const uint64_t ad = (static_cast<uint64_t>(a_hi) << 32) | (a_lo + ks * (2048 >> 4));
              const uint64_t bd = (static_cast<uint64_t>(b_hi) << 32) | (x_lo + (row0[tp] + ks * 16) * 8);
              umma_bf16(tbase + tp * NCH, ad, bd, idesc, ks ? 1u : acc);
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
    // ------------------------------------------------------------------ final epilogue: TMEM -> workspace partials
    if (nit > 0) {
      mbar_wait(done, 0);
      tc_fence_after();
    }
    const int crow = warp * 32 + lane;            // row of the 128-row accumulator
    const long long Mpad = p.pair3d ? 27LL * 32 : 9LL * p.kdn * p.cin;
    const int tap_base = (p.kdn > 1 && !p.pair3d) ? static_cast<int>(blockIdx.z) * 9 : 0;
    for (int tp = 0; tp < ntap; ++tp) {
      // PAIR: rows 0..63 of group tp belong to the tap the slab was started at, rows 64..127 (pair groups only) to its
      // left neighbour (kh, 1); otherwise row == cout and the accumulator is tap0 + tp
      const int tap = PAIR ? (tp >> 1) * 3 + ((tp & 1) ? 0 : (crow < 64 ? 2 : 1)) : tap0 + tp;
      const int co = PAIR ? (crow & 63) : crow;
      const bool row_ok = PAIR ? (crow < 64 || !(tp & 1)) : ((cout0 + crow) < p.cout);
#pragma unroll 1
      for (int c0 = 0; c0 < NCH; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + tp * NCH + c0, r);
        tmem_ld_wait();
        const int kd_pair = 2 * static_cast<int>(blockIdx.z) + (c0 >> 5);  // pair3d: depth tap of this 32-column half
        if (row_ok && !(p.pair3d && kd_pair > 2)) {
          const long long wrow = p.pair3d ? static_cast<long long>(kd_pair * 9 + tap) * 32
                                          : static_cast<long long>(tap_base + tap) * p.cin + cin0 + c0;
          float* dst = p.ws + (static_cast<long long>(blockIdx.y) * Mpad + wrow) * p.cout + cout0 + co;
#pragma unroll
          for (int j = 0; j < 32; ++j) dst[static_cast<long long>(j) * p.cout] = nit > 0 ? __uint_as_float(r[j]) : 0.f;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc<TCOLS>(tmem_base);
}

}  // namespace qt
