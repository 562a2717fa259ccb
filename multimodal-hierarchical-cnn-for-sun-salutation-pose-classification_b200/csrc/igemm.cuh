// Implicit-GEMM convolution / linear kernels on tcgen05 tensor cores (sm_100a).
//
//   igemm_kmajor_kernel : D[M, Nout] = gather(A)[M, K] * B[Nout, K]^T      (fprop, dgrad, linear)
//   igemm_wgrad_kernel  : D[F, Nout] = gather(A)[P, F]^T * Dy[P, Nout]     (weight gradient, P = pixels)
//
// Both are warp-specialised: warps 0-3 gather operand tiles into 128B-swizzled shared memory
// with 16-byte cp.async (zero-fill gives conv padding, ragged tiles and strided-dgrad holes for
// free), warp 4 owns TMEM and issues tcgen05.mma from one thread, and when the main loop ends
// warps 0-3 turn into the epilogue (tcgen05.ld -> bias / residual / ReLU / BN partial sums ->
// global). Accumulators live in TMEM (fp32). Two CTAs are resident per SM so one CTA's epilogue
// overlaps the other's main loop.
//
// The M index space is a logical output grid (n, od, oh, ow); the input coordinate of tap t is
// o*mult + off[t], which expresses forward convs (mult = stride, off = r*dil - pad), stride-1
// dgrad (off = pad - r, flipped taps), parity-decomposed stride-2 dgrad, the four quadrant
// views of the quadtree split (group offsets) and plain linear layers (one tap, 1x1 grid).
#pragma once
#include <cuda.h>  // CUtensorMap (type only)

#include "ptx.cuh"

namespace qt {

constexpr int kMaxTaps = 32;
constexpr int kBM = 128;   // UMMA M (TMEM lanes)
constexpr int kBK = 64;    // bf16 per k-block = one 128-byte swizzle row
constexpr int kProducerThreads = 128;   // epilogue threads (warps 0-3); they also produce
constexpr int kGemmThreads = 160;
constexpr int kGemmThreadsWide = 288;   // + warps 5-8 as extra producers
// kLag (template parameter of the kernels) = cp.async groups in flight before a stage is published; the ring
// needs STAGES - kLag - 1 >= 1 stages of slack or producer and MMA issuer serialise.

enum : int {
  EPI_BIAS = 1,
  EPI_RELU = 2,
  EPI_STATS = 4,    // per-CTA column sum / sum-of-squares of the stored (rounded) values
  EPI_ADDEND = 8,   // out = acc + addend (bf16, same view as out)
  EPI_OUT_F32 = 16, // store fp32 instead of bf16
  EPI_SPLITK = 32,  // store raw fp32 partials into the split-K workspace
  EPI_WGRAD_DIRECT = 64,  // weight-gradient kernel, one split, one tap (linear layers): write dw[nout][F] straight from the
                          // accumulator (+= with EPI_ADDEND) instead of a workspace round trip through the reduce kernel
};

struct View4 {  // element strides of an (n, d, h, w, c) view; c is contiguous
  long long sn, sd, sh, sw;
};

struct IgemmParams {
  // A operand: gathered activations
  const __nv_bfloat16* a;
  View4 av;
  int id, ih, iw;            // input bounds for the zero-fill test
  int nb, od, oh, ow;        // logical output grid, M = nb*od*oh*ow
  int mult_d, mult_h, mult_w;
  int ntaps, cin, cin_log2;  // K = ntaps*cin ; cin is a power of two when ntaps > 1
  signed char off_d[kMaxTaps], off_h[kMaxTaps], off_w[kMaxTaps];
  short wtap[kMaxTaps];      // tap index inside the weight tensor
  // B operand: weights [nout][wtaps][cin] (K-major kernel) or dy [P][nout] (wgrad kernel)
  const __nv_bfloat16* b;
  int nout, wtaps;
  // Output
  void* out;
  View4 ov;
  const float* bias;
  const __nv_bfloat16* addend;
  float* stats;              // [groups*gridDim.x][2][nout]
  float* splitk_ws;          // [groups*ksplit][Mpad][Npad]
  int flags, ksplit, groups;
  long long a_goff[4], o_goff[4], b_goff[4];
  int M, num_kb, kb_per_split;
  int Mpad, Npad;
  int adv_n, adv_d, adv_h, adv_w;  // wgrad: 64 pixels decomposed over (n, od, oh, ow)
  int b_tma;  // K-major kernel: 0 = weights by cp.async, 1 = by TMA (k-block inside one tap), 2 = by TMA (natural tap order)
};

// ---------------------------------------------------------------------------------------------
struct RowCoord {
  long long base;  // element offset of (n, od*mult, oh*mult, ow*mult, 0), or -1 when the row is past M
  int cd, ch, cw;
};

__device__ __forceinline__ RowCoord decompose_row(const IgemmParams& p, int m, long long goff) {
  RowCoord rc;
  if (m >= p.M) {
    rc.base = -1; rc.cd = rc.ch = rc.cw = 0;
    return rc;
  }
  int idx = m;
  const int ow = idx % p.ow; idx /= p.ow;
  const int oh = idx % p.oh; idx /= p.oh;
  const int od = idx % p.od;
  const int n = idx / p.od;
  rc.cd = od * p.mult_d; rc.ch = oh * p.mult_h; rc.cw = ow * p.mult_w;
  rc.base = goff + n * p.av.sn + rc.cd * p.av.sd + rc.ch * p.av.sh + rc.cw * p.av.sw;
  return rc;
}

__device__ __forceinline__ long long out_row_offset(const IgemmParams& p, int m, long long goff) {
  int idx = m;
  const int ow = idx % p.ow; idx /= p.ow;
  const int oh = idx % p.oh; idx /= p.oh;
  const int od = idx % p.od;
  const int n = idx / p.od;
  return goff + n * p.ov.sn + od * p.ov.sd + oh * p.ov.sh + ow * p.ov.sw;
}

// Sum a[32] across the 32 lanes of a warp, leaving column `lane` in lane `lane` (31 shuffles).
__device__ __forceinline__ float warp_transpose_reduce(float (&v)[32]) {
  const uint32_t lane = lane_id();
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const bool hi = lane & 16;
    const float send = hi ? v[i] : v[i + 16];
    const float keep = hi ? v[i + 16] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const bool hi = lane & 8;
    const float send = hi ? v[i] : v[i + 8];
    const float keep = hi ? v[i + 8] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const bool hi = lane & 4;
    const float send = hi ? v[i] : v[i + 4];
    const float keep = hi ? v[i + 4] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const bool hi = lane & 2;
    const float send = hi ? v[i] : v[i + 2];
    const float keep = hi ? v[i + 2] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  {
    const bool hi = lane & 1;
    const float send = hi ? v[0] : v[1];
    const float keep = hi ? v[1] : v[0];
    v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
  }
  return v[0];
}

template <int BN, int STAGES>
struct KMajorSmem {
  static constexpr int kABytes = kBM * 128;
  static constexpr int kBBytes = BN * 128;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarOffset = STAGES * kStageBytes;
  static constexpr int kTotal = kBarOffset + 256 + 1024;  // barriers + alignment slack
};

// =============================================================================================
// K-major kernel: fprop / dgrad / linear
// =============================================================================================
template <int BN, int STAGES, int kLag, int NPW>
__global__ void __launch_bounds__(NPW == 8 ? kGemmThreadsWide : kGemmThreads, 2) igemm_kmajor_kernel(const __grid_constant__ IgemmParams p,
                                                                                       const __grid_constant__ CUtensorMap bmap) {
  using L = KMajorSmem<BN, STAGES>;
  constexpr uint32_t TCOLS = BN < 32 ? 32 : BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* accum_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * kBM;
  const int n0 = blockIdx.y * BN;
  const int g = blockIdx.z / p.ksplit;
  const int split = blockIdx.z - g * p.ksplit;
  const int kb_begin = split * p.kb_per_split;
  const int kb_end = min(p.num_kb, kb_begin + p.kb_per_split);
  const int nit = kb_end - kb_begin;

  constexpr int kNP = NPW * 32;          // producer threads
  constexpr int kRowStep = NPW * 4;      // rows covered per pass (8 chunks per row)
  constexpr int kARows = kBM / kRowStep; // A rows per producer thread
  if (warp == 4) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(&full_bar[s], kNP + (p.b_tma ? 1 : 0));  // + the expect_tx arrive of the TMA weight tile
        mbar_init(&empty_bar[s], 1);
      }
      mbar_init(accum_bar, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc<TCOLS>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp != 4) {
    // ------------------------------------------------------------------ producer
    const int t = warp < 4 ? threadIdx.x : threadIdx.x - 32;
    const int chunk = t & 7;
    const int rbase = t >> 3;
    const uint32_t sw_off = static_cast<uint32_t>((chunk ^ (rbase & 7)) << 4);
    RowCoord rc[kARows];
#pragma unroll
    for (int i = 0; i < kARows; ++i) rc[i] = decompose_row(p, m0 + rbase + kRowStep * i, p.a_goff[g]);
    const __nv_bfloat16* bptr = p.b + p.b_goff[g];
    const long long brow_stride = static_cast<long long>(p.wtaps) * p.cin;

    for (int it = 0; it < nit; ++it) {
      const int s = it % STAGES;
      if (it >= STAGES) mbar_wait(&empty_bar[s], ((it / STAGES) - 1) & 1);
      const int k = (kb_begin + it) * kBK + chunk * 8;
      int tap, c;
      if (p.ntaps == 1) { tap = 0; c = k; } else { tap = k >> p.cin_log2; c = k & (p.cin - 1); }
      const bool tap_ok = (p.ntaps == 1) ? (k < p.cin) : (tap < p.ntaps);
      const int tsel = tap_ok ? tap : 0;
      const int td = p.off_d[tsel], th = p.off_h[tsel], tw = p.off_w[tsel];
      const long long toff = td * p.av.sd + th * p.av.sh + tw * p.av.sw + c;
      const long long woff = static_cast<long long>(p.wtap[tsel]) * p.cin + c;
      uint8_t* stage = smem + s * L::kStageBytes;
      const uint32_t a_dst = smem_u32(stage) + rbase * 128 + sw_off;
      const uint32_t b_dst = smem_u32(stage + L::kABytes) + rbase * 128 + sw_off;
      if (p.b_tma && t == 0) {
        // weight tile [BN rows][64 K] by one TMA instruction (128B-swizzled by the copy engine, zero OOB fill)
        const int kbase = (kb_begin + it) * kBK;
        int kx = kbase;
        if (p.b_tma == 1) {
          const int tp1 = p.ntaps > 1 ? (kbase >> p.cin_log2) : 0;
          const int c1 = p.ntaps > 1 ? (kbase & (p.cin - 1)) : kbase;
          kx = p.wtap[tp1] * p.cin + c1;
        }
        mbar_arrive_expect_tx(&full_bar[s], BN * 128);
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(
                smem_u32(stage + L::kABytes)),
            "l"(reinterpret_cast<uint64_t>(&bmap)), "r"(kx), "r"(n0), "r"(smem_u32(&full_bar[s]))
            : "memory");
      }
#pragma unroll
      for (int i = 0; i < kARows; ++i) {
        const bool ok = tap_ok && rc[i].base >= 0 && static_cast<unsigned>(rc[i].cd + td) < static_cast<unsigned>(p.id) &&
                        static_cast<unsigned>(rc[i].ch + th) < static_cast<unsigned>(p.ih) &&
                        static_cast<unsigned>(rc[i].cw + tw) < static_cast<unsigned>(p.iw);
        const __nv_bfloat16* src = ok ? (p.a + rc[i].base + toff) : p.a;
        cp_async16(a_dst + i * kRowStep * 128, src, ok ? 16u : 0u);
      }
      if (!p.b_tma) {
#pragma unroll
        for (int i = 0; i < BN / kRowStep; ++i) {
          const int n = n0 + rbase + kRowStep * i;
          const bool ok = tap_ok && n < p.nout;
          const __nv_bfloat16* src = ok ? (bptr + n * brow_stride + woff) : bptr;
          cp_async16(b_dst + i * kRowStep * 128, src, ok ? 16u : 0u);
        }
      }
      // completion tracked by the mbarrier (no wait_group); the issuer fences after its wait
      cp_async_mbar_arrive_noinc(&full_bar[s]);
    }
    cp_async_wait<0>();  // nothing may be in flight when the CTA retires
  } else {
    // ------------------------------------------------------------------ MMA issuer (warp 4)
    // warp-uniform control flow (descriptor math on the uniform datapath); lane 0 issues
    {
      constexpr uint32_t idesc = make_idesc_bf16(kBM, BN, 0, 0);
      constexpr uint32_t d_hi = (1024u >> 4) | (1u << 14) | (kLayoutSW128 << 29);
      constexpr uint32_t d_lbo = 1u << 16;
      const uint32_t tbase = __shfl_sync(0xffffffffu, tmem_base, 0);
      for (int it = 0; it < nit; ++it) {
        const int s = it % STAGES;
        mbar_wait(&full_bar[s], (it / STAGES) & 1);
        fence_proxy_async_smem();
        tc_fence_after();
        const uint32_t a_lo = ((smem_u32(smem + s * L::kStageBytes) >> 4) & 0x3FFFu) | d_lbo;
        const uint32_t b_lo = a_lo + (L::kABytes >> 4);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            const uint64_t ad = (static_cast<uint64_t>(d_hi) << 32) | (a_lo + k * 2);
            const uint64_t bd = (static_cast<uint64_t>(d_hi) << 32) | (b_lo + k * 2);
            umma_bf16(tbase, ad, bd, idesc, (it | k) ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(accum_bar);
      __syncwarp();
    }
  }

  if (warp < 4) {
    // ------------------------------------------------------------------ epilogue
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const int row = warp * 32 + lane;
    const int m = m0 + row;
    const bool row_ok = m < p.M;
    const long long orow = row_ok ? out_row_offset(p, m, p.o_goff[g]) : 0;
    float* stat_scratch = reinterpret_cast<float*>(smem);  // [2][4][BN], aliases stage 0 (all MMAs retired)
    const int flags = p.flags;
    // rows this lane writes out in the staged (coalesced) path: 8*i + lane/4; -1 = outside the problem
    long long srow[4];
    {
      const long long mine = row_ok ? orow : -1;
#pragma unroll
      for (int i = 0; i < 4; ++i) srow[i] = __shfl_sync(0xffffffffu, mine, 8 * i + (lane >> 2));
    }
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t r[32];
      tmem_ld32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + c0, r);
      tmem_ld_wait();
      const int ncol = n0 + c0;
      if (ncol >= p.nout) break;  // uniform across the CTA
      const bool full_chunk = (ncol + 32 <= p.nout);
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);

      if (flags & EPI_SPLITK) {
        if (row_ok) {
          float* dst = p.splitk_ws + (static_cast<long long>(blockIdx.z) * p.Mpad + m) * p.Npad + ncol;
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
        continue;
      }
      if (flags & EPI_BIAS) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += (ncol + j < p.nout) ? __ldg(p.bias + ncol + j) : 0.f;
      }
      if ((flags & EPI_ADDEND) && row_ok) {
        const __nv_bfloat16* ad = p.addend + orow + ncol;
        if (full_chunk) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            const uint4 q = *reinterpret_cast<const uint4*>(ad + j);
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[e]);
              v[j + 2 * e] += __low2float(h);
              v[j + 2 * e + 1] += __high2float(h);
            }
          }
        } else {
          for (int j = 0; j < 32; ++j)
            if (ncol + j < p.nout) v[j] += __bfloat162float(ad[j]);
        }
      }
      if (flags & EPI_RELU) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
      }
      if (flags & EPI_OUT_F32) {
        if (row_ok) {
          float* dst = reinterpret_cast<float*>(p.out) + orow + ncol;
          if (full_chunk && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          } else {
            for (int j = 0; j < 32; ++j)
              if (ncol + j < p.nout) dst[j] = v[j];
          }
        }
      } else if (__all_sync(0xffffffffu, full_chunk && (!row_ok || ((orow | ncol) & 7) == 0) &&
                                         ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0))) {
        // Coalesced write-out through a per-warp staging tile in the (retired) first pipeline stage: one store
        // instruction covers 8 rows x 64 contiguous bytes instead of 32 rows x 16 bytes, which keeps the LSU free for
        // the co-resident CTA's cp.async gathers (same scheme as conv3x3.cuh). BatchNorm column sums are read back
        // from the staged bf16 values.
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) pk[j] = row_ok ? pack_bf16x2(v[2 * j], v[2 * j + 1]) : 0u;
        const uint32_t stage_w = smem_u32(smem) + 4096 + warp * 2048;
        const uint32_t st_own = stage_w + lane * 64, st_sw = (lane >> 1) & 3;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(st_own + ((j ^ st_sw) << 4)), "r"(pk[4 * j]),
                       "r"(pk[4 * j + 1]), "r"(pk[4 * j + 2]), "r"(pk[4 * j + 3])
                       : "memory");
        __syncwarp();
        const uint32_t ld_row = stage_w + (lane >> 2) * 64 + (((lane & 3) ^ ((lane >> 3) & 3)) << 4);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 q;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];\n" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "r"(ld_row + i * 512) : "memory");
          if (srow[i] >= 0) *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + srow[i] + ncol + (lane & 3) * 8) = q;
        }
        if (flags & EPI_STATS) {
          // lane l: column pair (l & 15), rows of parity l >> 4 (see conv3x3.cuh)
          float cs1[2] = {0.f, 0.f}, cs2[2] = {0.f, 0.f};
          const uint32_t cs_addr = stage_w + (lane >> 4) * 64 + (lane & 3) * 4;
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            uint32_t w;
            asm volatile("ld.shared.b32 %0, [%1];\n" : "=r"(w) : "r"(cs_addr + k * 128 + ((((lane >> 2) & 3) ^ (k & 3)) << 4)) : "memory");
            const float lo = __uint_as_float(w << 16), hi = __uint_as_float(w & 0xffff0000u);
            cs1[0] += lo; cs1[1] += hi;
            cs2[0] = fmaf(lo, lo, cs2[0]); cs2[1] = fmaf(hi, hi, cs2[1]);
          }
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            cs1[e] += __shfl_xor_sync(0xffffffffu, cs1[e], 16);
            cs2[e] += __shfl_xor_sync(0xffffffffu, cs2[e], 16);
          }
          if (lane < 16) {
            *reinterpret_cast<float2*>(&stat_scratch[(0 * 4 + warp) * BN + c0 + 2 * lane]) = make_float2(cs1[0], cs1[1]);
            *reinterpret_cast<float2*>(&stat_scratch[(1 * 4 + warp) * BN + c0 + 2 * lane]) = make_float2(cs2[0], cs2[1]);
          }
        }
        __syncwarp();
        continue;
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = bf16_round(v[j]);
        if (row_ok) {
          __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.out) + orow + ncol;
          if (full_chunk && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              uint4 q;
              q.x = pack_bf16x2(v[j], v[j + 1]);
              q.y = pack_bf16x2(v[j + 2], v[j + 3]);
              q.z = pack_bf16x2(v[j + 4], v[j + 5]);
              q.w = pack_bf16x2(v[j + 6], v[j + 7]);
              *reinterpret_cast<uint4*>(dst + j) = q;
            }
          } else {
            for (int j = 0; j < 32; ++j)
              if (ncol + j < p.nout) dst[j] = __float2bfloat16_rn(v[j]);
          }
        }
      }
      if (flags & EPI_STATS) {
        float sq[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          v[j] = row_ok ? v[j] : 0.f;
          sq[j] = v[j] * v[j];
        }
        const float s1 = warp_transpose_reduce(v);
        const float s2 = warp_transpose_reduce(sq);
        stat_scratch[(0 * 4 + warp) * BN + c0 + lane] = s1;
        stat_scratch[(1 * 4 + warp) * BN + c0 + lane] = s2;
      }
    }
    if (flags & EPI_STATS) {
      asm volatile("bar.sync 1, 128;\n" ::: "memory");
      const int tile = g * gridDim.x + blockIdx.x;
      for (int i = threadIdx.x; i < 2 * BN; i += kProducerThreads) {
        const int which = i / BN, col = i - which * BN;
        if (n0 + col < p.nout) {
          const float* sc = stat_scratch + which * 4 * BN + col;
          const float tot = (sc[0] + sc[BN]) + (sc[2 * BN] + sc[3 * BN]);
          p.stats[(static_cast<long long>(tile) * 2 + which) * p.nout + n0 + col] = tot;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc<TCOLS>(tmem_base);
}

// =============================================================================================
// Weight-gradient kernel: both operands MN-major (rows of the smem tiles are pixels = GEMM K).
//   A tile : 2 blocks of [64 pixels][64 features]  (features = (tap, cin) of the forward input)
//   B tile : BN/64 blocks of [64 pixels][64 couts] (dy)
//   D      : [128 features][BN couts] fp32 partial, one per pixel split, reduced afterwards.
// =============================================================================================
template <int BN, int STAGES>
struct WgradSmem {
  static constexpr int kABytes = 2 * 64 * 128;
  static constexpr int kBBytes = (BN / 64) * 64 * 128;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarOffset = STAGES * kStageBytes;
  static constexpr int kTotal = kBarOffset + 256 + 1024;
};

template <int BN, int STAGES, int kLag, int NPW>
__global__ void __launch_bounds__(NPW == 8 ? kGemmThreadsWide : kGemmThreads, 2) igemm_wgrad_kernel(const __grid_constant__ IgemmParams p) {
  using L = WgradSmem<BN, STAGES>;
  constexpr uint32_t TCOLS = BN;
  static_assert(BN % 64 == 0, "wgrad BN must be a multiple of 64");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* accum_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int f0 = blockIdx.x * kBM;   // feature tile
  const int n0 = blockIdx.y * BN;    // cout tile
  const int g = blockIdx.z / p.ksplit;
  const int split = blockIdx.z - g * p.ksplit;
  const int kb_begin = split * p.kb_per_split;
  const int kb_end = min(p.num_kb, kb_begin + p.kb_per_split);
  const int nit = kb_end - kb_begin;
  const int F = p.ntaps * p.cin;

  constexpr int kNP = NPW * 32;
  constexpr int kRowStep = NPW * 4;
  constexpr int kPRows = 64 / kRowStep;  // pixel rows per producer thread and k-block
  if (warp == 4) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(&full_bar[s], kNP);
        mbar_init(&empty_bar[s], 1);
      }
      mbar_init(accum_bar, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc<TCOLS>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp != 4) {
    const int t = warp < 4 ? threadIdx.x : threadIdx.x - 32;
    const int chunk = t & 7;
    const int rbase = t >> 3;  // pixel rows rbase + kRowStep*i
    const uint32_t sw_off = static_cast<uint32_t>((chunk ^ (rbase & 7)) << 4);
    // Feature decode for the two A blocks handled by this thread's chunk (fixed per CTA).
    bool f_ok[2];
    int f_td[2], f_th[2], f_tw[2];
    long long f_off[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int f = f0 + j * 64 + chunk * 8;
      f_ok[j] = f < F;
      int tap = 0, c = f;
      if (p.ntaps > 1) { tap = f >> p.cin_log2; c = f & (p.cin - 1); }
      if (!f_ok[j]) { tap = 0; c = 0; }
      f_td[j] = p.off_d[tap]; f_th[j] = p.off_h[tap]; f_tw[j] = p.off_w[tap];
      f_off[j] = f_td[j] * p.av.sd + f_th[j] * p.av.sh + f_tw[j] * p.av.sw + c;
    }
    const __nv_bfloat16* bptr = p.b + p.b_goff[g];
    // Pixel coordinates of this thread's four rows, advanced incrementally by 64 pixels per k-block.
    int rn[kPRows], rd[kPRows], rh[kPRows], rw[kPRows];
#pragma unroll
    for (int i = 0; i < kPRows; ++i) {
      int idx = kb_begin * 64 + rbase + kRowStep * i;
      rw[i] = idx % p.ow; idx /= p.ow;
      rh[i] = idx % p.oh; idx /= p.oh;
      rd[i] = idx % p.od;
      rn[i] = idx / p.od;
    }
    for (int it = 0; it < nit; ++it) {
      const int s = it % STAGES;
      if (it >= STAGES) mbar_wait(&empty_bar[s], ((it / STAGES) - 1) & 1);
      uint8_t* stage = smem + s * L::kStageBytes;
      const uint32_t a_dst = smem_u32(stage) + rbase * 128 + sw_off;
      const uint32_t b_dst = smem_u32(stage + L::kABytes) + rbase * 128 + sw_off;
#pragma unroll
      for (int i = 0; i < kPRows; ++i) {
        const bool row_ok = rn[i] < p.nb;
        const int cd = rd[i] * p.mult_d, ch = rh[i] * p.mult_h, cw = rw[i] * p.mult_w;
        const long long abase = p.a_goff[g] + rn[i] * p.av.sn + cd * p.av.sd + ch * p.av.sh + cw * p.av.sw;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const bool ok = f_ok[j] && row_ok &&
                          static_cast<unsigned>(cd + f_td[j]) < static_cast<unsigned>(p.id) &&
                          static_cast<unsigned>(ch + f_th[j]) < static_cast<unsigned>(p.ih) &&
                          static_cast<unsigned>(cw + f_tw[j]) < static_cast<unsigned>(p.iw);
          const __nv_bfloat16* src = ok ? (p.a + abase + f_off[j]) : p.a;
          cp_async16(a_dst + j * 8192 + i * kRowStep * 128, src, ok ? 16u : 0u);
        }
        // dy rows through the output view
        const long long drow = rn[i] * p.ov.sn + rd[i] * p.ov.sd + rh[i] * p.ov.sh + rw[i] * p.ov.sw;
#pragma unroll
        for (int j = 0; j < BN / 64; ++j) {
          const int n = n0 + j * 64 + chunk * 8;
          const bool ok = row_ok && (n < p.nout);
          const __nv_bfloat16* src = ok ? (bptr + drow + n) : bptr;
          cp_async16(b_dst + j * 8192 + i * kRowStep * 128, src, ok ? 16u : 0u);
        }
        // advance by 64 pixels (mixed-radix add with single carries)
        rw[i] += p.adv_w; if (rw[i] >= p.ow) { rw[i] -= p.ow; rh[i] += 1; }
        rh[i] += p.adv_h; if (rh[i] >= p.oh) { rh[i] -= p.oh; rd[i] += 1; }
        rd[i] += p.adv_d; if (rd[i] >= p.od) { rd[i] -= p.od; rn[i] += 1; }
        rn[i] += p.adv_n;
      }
      // completion tracked by the mbarrier (no wait_group); the issuer fences after its wait
      cp_async_mbar_arrive_noinc(&full_bar[s]);
    }
    cp_async_wait<0>();  // nothing may be in flight when the CTA retires
  } else {
    {
      constexpr uint32_t idesc = make_idesc_bf16(kBM, BN, 1, 1);
      // MN-major SW128: LBO = 8192 B between 64-wide MN blocks, SBO = 1024 B between 8-pixel groups
      constexpr uint32_t d_hi = (1024u >> 4) | (1u << 14) | (kLayoutSW128 << 29);
      constexpr uint32_t d_lbo = (8192u >> 4) << 16;
      const uint32_t tbase = __shfl_sync(0xffffffffu, tmem_base, 0);
      for (int it = 0; it < nit; ++it) {
        const int s = it % STAGES;
        mbar_wait(&full_bar[s], (it / STAGES) & 1);
        fence_proxy_async_smem();
        tc_fence_after();
        const uint32_t a_lo = ((smem_u32(smem + s * L::kStageBytes) >> 4) & 0x3FFFu) | d_lbo;
        const uint32_t b_lo = a_lo + (L::kABytes >> 4);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {  // 16 pixels (two 8-row swizzle atoms) per MMA
            const uint64_t ad = (static_cast<uint64_t>(d_hi) << 32) | (a_lo + k * (2048 >> 4));
            const uint64_t bd = (static_cast<uint64_t>(d_hi) << 32) | (b_lo + k * (2048 >> 4));
            umma_bf16(tbase, ad, bd, idesc, (it | k) ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(accum_bar);
      __syncwarp();
    }
  }

  if (warp < 4) {
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const int f = f0 + warp * 32 + lane;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t r[32];
      tmem_ld32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + c0, r);
      tmem_ld_wait();
      const int ncol = n0 + c0;
      if (ncol >= p.Npad) break;
      if (p.flags & EPI_WGRAD_DIRECT) {
        // dw[n][f]: for a fixed column the 32 lanes (consecutive f) write 128 contiguous bytes
        const int F = p.ntaps * p.cin;
        if (f < F) {
          float* dw = reinterpret_cast<float*>(p.out) + f;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (ncol + j < p.nout) {
              float* dst = dw + static_cast<long long>(ncol + j) * F;
              const float v = __uint_as_float(r[j]);
              *dst = (p.flags & EPI_ADDEND) ? (*dst + v) : v;
            }
          }
        }
      } else if (f < p.Mpad) {
        float* dst = p.splitk_ws + (static_cast<long long>(blockIdx.z) * p.Mpad + f) * p.Npad + ncol;
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(dst + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                            __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc<TCOLS>(tmem_base);
}

// ---------------------------------------------------------------------------------------------
// Split-K reductions (deterministic: fixed summation order over the splits).
// ---------------------------------------------------------------------------------------------
// Row-major result: out[m][n] = act(sum_z ws[z][m][n] + bias[n]); out is bf16 or fp32 with row stride ldo.
__global__ void splitk_reduce_rows_kernel(const float* __restrict__ ws, int nsplit, int M, int N, int Mpad, int Npad,
                                          const float* __restrict__ bias, int relu, int out_f32, void* out,
                                          long long ldo) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(M) * N) return;
  const int m = idx / N, n = idx - static_cast<long long>(m) * N;
  float acc = 0.f;
  for (int z = 0; z < nsplit; ++z) acc += ws[(static_cast<long long>(z) * Mpad + m) * Npad + n];
  if (bias) acc += bias[n];
  if (relu) acc = fmaxf(acc, 0.f);
  if (out_f32) reinterpret_cast<float*>(out)[m * ldo + n] = acc;
  else reinterpret_cast<__nv_bfloat16*>(out)[m * ldo + n] = __float2bfloat16_rn(acc);
}

// Weight-gradient result: ws[z][f = tap*cin + c][n] -> grad[n][c][tap] (PyTorch parameter layout
// [nout][cin][ntaps]). grid (ceil(cin/CT), ceil(N/32), ntaps), block 256 = 8 warps x 32 lanes. Reads: a warp owns
// CT/8 channel rows, lanes run along n (128-byte rows of the workspace), eight splits in flight per thread. The
// CT x 32 tile is transposed through shared memory so the gradient is written with lanes along c (contiguous for
// linear layers, stride ntaps for convs). CT = 8 is for the small layers (64x64x9 has only 36 tiles of 32x32, and
// up to 148 splits to sum: the narrow tile spreads that over 144 blocks). accumulate != 0 adds.
template <int CT, int ZS>
__global__ void __launch_bounds__(256 * ZS) splitk_reduce_wgrad_kernel(const float* __restrict__ ws, int nsplit, int F, int N, int Mpad,
                                                                       int Npad, int cin, int ntaps, float* __restrict__ grad,
                                                                       int accumulate) {
  // ZS > 1: the splits are divided over ZS groups of 8 warps (more loads in flight when few blocks must sum many
  // partials, e.g. 148 partials of a 64x64x9 filter); the groups' sums are combined in fixed order
  __shared__ float tile[ZS][CT][33];  // [z slice][c][n]
  const int c0 = blockIdx.x * CT, n0 = blockIdx.y * 32, tap = blockIdx.z;
  const int lane = threadIdx.x & 31, wrp = (threadIdx.x >> 5) & 7, zs = threadIdx.x >> 8;
  const long long zstride = static_cast<long long>(Mpad) * Npad;
  const int zper = (nsplit + ZS - 1) / ZS;
  const int z_begin = zs * zper, z_end = min(nsplit, z_begin + zper);
#pragma unroll
  for (int j = 0; j < CT / 8; ++j) {
    const int cl = wrp * (CT / 8) + j;
    const int c = c0 + cl, n = n0 + lane;
    float acc = 0.f;
    if (c < cin && n < N) {
      const float* src = ws + (static_cast<long long>(tap) * cin + c) * Npad + n;
      float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      int z = z_begin;
      for (; z + 8 <= z_end; z += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u) a[u] += src[(z + u) * zstride];
      }
      for (; z < z_end; ++z) a[0] += src[z * zstride];
      acc = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
    }
    tile[zs][cl][lane] = acc;
  }
  __syncthreads();
  if (threadIdx.x >= 256) return;
  constexpr int kCLanes = CT < 32 ? CT : 32;  // lanes along c in the write phase
  const int wc = threadIdx.x % kCLanes;
  for (int nl = threadIdx.x / kCLanes; nl < 32; nl += 256 / kCLanes) {
    for (int cl = wc; cl < CT; cl += kCLanes) {
      const int n = n0 + nl, c = c0 + cl;
      if (c < cin && n < N) {
        float* dst = grad + (static_cast<long long>(n) * cin + c) * ntaps + tap;
        float v = tile[0][cl][nl];
#pragma unroll
        for (int q = 1; q < ZS; ++q) v += tile[q][cl][nl];
        *dst = accumulate ? (*dst + v) : v;
      }
    }
  }
}
inline void launch_splitk_reduce_wgrad(const float* ws, int nsplit, int F, int N, int Mpad, int Npad, int cin, int ntaps,
                                       float* grad, int accumulate, cudaStream_t st) {
  const long long blocks32 = static_cast<long long>((cin + 31) / 32) * ((N + 31) / 32) * ntaps;
  if (blocks32 >= 296) {
    dim3 g((cin + 31) / 32, (N + 31) / 32, ntaps);
    splitk_reduce_wgrad_kernel<32, 1><<<g, 256, 0, st>>>(ws, nsplit, F, N, Mpad, Npad, cin, ntaps, grad, accumulate);
  } else {
    dim3 g((cin + 7) / 8, (N + 31) / 32, ntaps);
    if (nsplit >= 32) splitk_reduce_wgrad_kernel<8, 4><<<g, 1024, 0, st>>>(ws, nsplit, F, N, Mpad, Npad, cin, ntaps, grad, accumulate);
    else splitk_reduce_wgrad_kernel<8, 1><<<g, 256, 0, st>>>(ws, nsplit, F, N, Mpad, Npad, cin, ntaps, grad, accumulate);
  }
}

}  // namespace qt
