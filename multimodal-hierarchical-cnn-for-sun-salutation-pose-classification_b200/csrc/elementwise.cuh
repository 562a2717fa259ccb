// Bandwidth-bound kernels of the QuadtreeCNN hot path (NHWC / NDHWC bf16 activations, fp32 statistics):
// layout transforms, train-mode BatchNorm (finalize / apply+ReLU+residual / backward), max pools,
// the quadtree pooling stage (quadrant 2x2 max-pool + flatten + global average pool written straight
// into the fused feature buffer), weight packing and the small fp32 linear layers of the fusion head.
// All of them use 128-bit accesses along the contiguous channel dimension.
#pragma once
#include "ptx.cuh"

namespace qt {

struct bf16x8 {
  uint4 q;
};
__device__ __forceinline__ void unpack8(const uint4& q, float (&f)[8]) {
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[e]);
    f[2 * e] = __low2float(h);
    f[2 * e + 1] = __high2float(h);
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 q;
  q.x = pack_bf16x2(f[0], f[1]);
  q.y = pack_bf16x2(f[2], f[3]);
  q.z = pack_bf16x2(f[4], f[5]);
  q.w = pack_bf16x2(f[6], f[7]);
  return q;
}
__device__ __forceinline__ uint4 ld_nc16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];\n"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

// Counter-based Bernoulli keep-mask for dropout: identical in forward and backward, nothing stored.
__device__ __forceinline__ uint32_t hash_u32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}
__device__ __forceinline__ float dropout_scale(unsigned long long seed, uint32_t idx, float p) {
  if (p <= 0.f) return 1.f;
  const uint32_t h = hash_u32(idx ^ hash_u32(static_cast<uint32_t>(seed) + 0x9e3779b9U * static_cast<uint32_t>(seed >> 32)));
  const float u = (h >> 8) * (1.0f / 16777216.0f);
  return (u >= p) ? 1.f / (1.f - p) : 0.f;
}

// ---------------------------------------------------------------------------------------------
// Stem input packing: NCHW fp32 [N,3,H,W] -> zero-padded NHWC4 bf16 [N, H+7, W+8, 4]
// (3 rows/cols of padding before the image so that every 7x7/s2 window starts 16-byte aligned).
// ---------------------------------------------------------------------------------------------
// One block per padded row (grid.x = N * (H+7)): only 32-bit index arithmetic per pixel; reads are 128-byte coalesced
// per channel plane, writes 8 bytes per pixel.
// TIn = float (normalised fp32, what `images.to(device)` delivers), __nv_bfloat16, or uint8_t (decoded pixels: the
// ToTensor + Normalize of dataloader.py:35-36 is applied here as v * scale[c] + shift[c], scale = 1/(255 std), shift = -mean/std,
// so only one byte per value crosses PCIe).
template <typename TIn>
__device__ __forceinline__ float stem_in(const TIn* p, float sc, float sh);
template <>
__device__ __forceinline__ float stem_in<float>(const float* p, float, float) { return __ldg(p); }
template <>
__device__ __forceinline__ float stem_in<__nv_bfloat16>(const __nv_bfloat16* p, float, float) { return __bfloat162float(*p); }
template <>
__device__ __forceinline__ float stem_in<uint8_t>(const uint8_t* p, float sc, float sh) { return fmaf(static_cast<float>(__ldg(p)), sc, sh); }

template <typename TIn>
__global__ void stem_pack_input_kernel(const TIn* __restrict__ x, __nv_bfloat16* __restrict__ out, int N, int C,
                                       int H, int W, const float* __restrict__ scale, const float* __restrict__ shift) {
  const int Hp = H + 7, Wp = W + 8;
  const int row = blockIdx.x;            // n * Hp + hp
  const int n = row / Hp, hp = row - n * Hp;
  const int h = hp - 3;
  const bool row_in = h >= 0 && h < H;
  const TIn* src = x + (static_cast<long long>(n) * C * H + (row_in ? h : 0)) * W;
  const long long plane = static_cast<long long>(H) * W;
  __nv_bfloat16* dst = out + static_cast<long long>(row) * Wp * 4;
  float sc[4] = {1.f, 1.f, 1.f, 1.f}, sh[4] = {0.f, 0.f, 0.f, 0.f};
  if (scale) {
#pragma unroll
    for (int c = 0; c < 4; ++c)
      if (c < C) { sc[c] = scale[c]; sh[c] = shift[c]; }
  }
  for (int wp = threadIdx.x; wp < Wp; wp += blockDim.x) {
    const int w = wp - 3;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (row_in && w >= 0 && w < W) {
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (c < C) v[c] = stem_in<TIn>(src + c * plane + w, sc[c], sh[c]);
    }
    uint2 q;
    q.x = pack_bf16x2(v[0], v[1]);
    q.y = pack_bf16x2(v[2], v[3]);
    *reinterpret_cast<uint2*>(dst + wp * 4) = q;
  }
}

// NCHW fp32 -> NHWC bf16 with channel padding to Cp (generic; used by the 3-D stack and tests).
// Cp == 8 (3-channel clips padded to one 16-byte chunk per pixel): one thread per pixel, each channel plane read coalesced,
// one 16-byte store per pixel.
template <typename TIn>
__global__ void nchw_to_nhwc8_bf16_kernel(const TIn* __restrict__ x, __nv_bfloat16* __restrict__ out, int N, int C, long long HW,
                                          const float* __restrict__ scale, const float* __restrict__ shift) {
  const long long total = static_cast<long long>(N) * HW;
  float sc[8], sh[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) { sc[c] = (scale && c < C) ? scale[c] : 1.f; sh[c] = (scale && c < C) ? shift[c] : 0.f; }
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long n = i / HW, p = i - n * HW;
    float v[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) v[c] = c < C ? stem_in<TIn>(x + (n * C + c) * HW + p, sc[c], sh[c]) : 0.f;
    *reinterpret_cast<uint4*>(out + i * 8) = pack8(v);
  }
}
template <typename TIn>
__global__ void nchw_to_nhwc_bf16_kernel(const TIn* __restrict__ x, __nv_bfloat16* __restrict__ out, int N, int C, long long HW,
                                         int Cp, const float* __restrict__ scale, const float* __restrict__ shift) {
  const long long total = static_cast<long long>(N) * HW * Cp;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = i % Cp;
    const long long p = (i / Cp) % HW;
    const long long n = i / (static_cast<long long>(Cp) * HW);
    float v = 0.f;
    if (c < C) v = stem_in<TIn>(x + (n * C + c) * HW + p, scale ? scale[c] : 1.f, scale ? shift[c] : 0.f);
    out[i] = __float2bfloat16_rn(v);
  }
}
__global__ void nhwc_bf16_to_nchw_f32_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out, int N,
                                             int C, long long HW, int Cp) {
  const long long total = static_cast<long long>(N) * C * HW;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long p = i % HW;
    const int c = (i / HW) % C;
    const long long n = i / (static_cast<long long>(C) * HW);
    out[i] = __bfloat162float(x[(n * HW + p) * Cp + c]);
  }
}

// ---------------------------------------------------------------------------------------------
// Weight packing. w: fp32 [Cout][Cin][T] (PyTorch layout, T = taps).
//   fprop layout wf[Cout][T][Cin]  (K-major B operand of the forward GEMM)
//   dgrad layout wd[Cin][T][Cout]  (K-major B operand of the data-gradient GEMM)
// ---------------------------------------------------------------------------------------------
__global__ void wpack_fprop_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wf, int Cout, int Cin,
                                   int T) {
  const long long total = static_cast<long long>(Cout) * T * Cin;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ci = i % Cin;
    const int t = (i / Cin) % T;
    const long long co = i / (static_cast<long long>(Cin) * T);
    wf[i] = __float2bfloat16_rn(w[(co * Cin + ci) * T + t]);
  }
}
// Tiled transpose: block (32, 8), grid (ceil(Cin/32), ceil(Cout/32), T).
__global__ void wpack_dgrad_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wd, int Cout, int Cin,
                                   int T) {
  __shared__ float tile[32][33];
  const int t = blockIdx.z;
  const int ci0 = blockIdx.x * 32, co0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int co = co0 + j, ci = ci0 + threadIdx.x;
    tile[j][threadIdx.x] = (co < Cout && ci < Cin) ? w[(static_cast<long long>(co) * Cin + ci) * T + t] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int ci = ci0 + j, co = co0 + threadIdx.x;
    if (ci < Cin && co < Cout)
      wd[(static_cast<long long>(ci) * T + t) * Cout + co] = __float2bfloat16_rn(tile[threadIdx.x][j]);
  }
}
// Both layouts from one read of w: one block packs a tile of COT output channels x 32 input channels x T taps.
// Dynamic smem COT*(32*(T|1)+1) floats, odd pitches in both directions (bank-conflict free). COT = 8 for filters
// with taps (8 x 32 x 9 tiles keep even the 64x64 layers at 16+ blocks; the 16-byte runs of wd[ci][t][co0..co0+7]
// are one sector), COT = 32 for T == 1 (1x1 convs and linear layers: a plain 32 x 32 transpose tile).
template <int TC>  // TC > 0: taps known at compile time (index divisions become multiplies); 0: runtime T
__device__ __forceinline__ void wpack_tile_t(const float* __restrict__ w, __nv_bfloat16* __restrict__ wf,
                                             __nv_bfloat16* __restrict__ wd, int Cout, int Cin, int Trt, int COT, int bx, int by,
                                             float* tile) {
  const int T = TC > 0 ? TC : Trt;
  const int ci0 = bx * 32, co0 = by * COT;
  const int TP = T | 1;
  const int CP = 32 * TP + 1;
  const int run = 32 * T;
  const int cot_shift = COT == 32 ? 5 : 3;  // COT is 8 or 32
  const bool full = (co0 + COT <= Cout) && (ci0 + 32 <= Cin);
  for (int i = threadIdx.x; i < COT * run; i += blockDim.x) {
    const int co = i / run, r = i - co * run;  // r = ci_local*T + t, contiguous in w for a fixed co
    const int cil = r / T, t = r - cil * T;
    float v = 0.f;
    if (full || (co0 + co < Cout && ci0 + cil < Cin)) v = w[(static_cast<long long>(co0 + co) * Cin + ci0) * T + r];
    tile[co * CP + cil * TP + t] = v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < COT * run; i += blockDim.x) {
    {  // wf[co][t][ci]: ci fastest
      const int cil = i & 31, q = i >> 5, co = q / T, t = q - co * T;
      if (full || (co0 + co < Cout && ci0 + cil < Cin))
        wf[(static_cast<long long>(co0 + co) * T + t) * Cin + ci0 + cil] = __float2bfloat16_rn(tile[co * CP + cil * TP + t]);
    }
    if (wd) {  // wd[ci][t][co]: co fastest (COT per block)
      const int col = i & (COT - 1), q = i >> cot_shift, cil = q / T, t = q - cil * T;
      if (full || (co0 + col < Cout && ci0 + cil < Cin))
        wd[(static_cast<long long>(ci0 + cil) * T + t) * Cout + co0 + col] = __float2bfloat16_rn(tile[col * CP + cil * TP + t]);
    }
  }
}
__device__ __forceinline__ void wpack_tile(const float* __restrict__ w, __nv_bfloat16* __restrict__ wf,
                                           __nv_bfloat16* __restrict__ wd, int Cout, int Cin, int T, int COT, int bx, int by,
                                           float* tile) {
  if (T == 9) wpack_tile_t<9>(w, wf, wd, Cout, Cin, T, COT, bx, by, tile);
  else if (T == 1) wpack_tile_t<1>(w, wf, wd, Cout, Cin, T, COT, bx, by, tile);
  else if (T == 27) wpack_tile_t<27>(w, wf, wd, Cout, Cin, T, COT, bx, by, tile);
  else wpack_tile_t<0>(w, wf, wd, Cout, Cin, T, COT, bx, by, tile);
}
__global__ void wpack_both_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wf,
                                  __nv_bfloat16* __restrict__ wd, int Cout, int Cin, int T, int COT) {
  extern __shared__ float tile[];
  wpack_tile(w, wf, wd, Cout, Cin, T, COT, blockIdx.x, blockIdx.y, tile);
}
// Every weight of a model in ONE launch (the per-step repack after an optimizer update): block b belongs to the
// item with first_block <= b < next first_block (binary search over the device table).
struct WpackItem {
  const float* w;
  __nv_bfloat16* wf;
  __nv_bfloat16* wd;
  int cout, cin, taps;
  int co_tile, ci_tiles, first_block;
};
__global__ void wpack_multi_kernel(const WpackItem* __restrict__ items, int nitems) {
  extern __shared__ float tile[];
  int lo = 0, hi = nitems - 1;
  const int b = blockIdx.x;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (items[mid].first_block <= b) lo = mid; else hi = mid - 1;
  }
  const WpackItem it = items[lo];
  const int local = b - it.first_block;
  wpack_tile(it.w, it.wf, it.wd, it.cout, it.cin, it.taps, it.co_tile, local % it.ci_tiles, local / it.ci_tiles, tile);
}
// First Conv3d layer: [32][Cin <= 8][3][3][3] fp32 -> [18 tiles (kd, kh, half)][2 K chunks][32 cout][8 channels] bf16 for
// conv3d_c8_kernel: K chunk kc of half h is the pixel at dw = 2h + kc - 1 (dw = +2 does not exist: zero weights).
__global__ void wpack_conv3d_c8_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int Cin) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 18 * 2 * 32 * 8) return;
  const int c = i & 7, co = (i >> 3) & 31, kc = (i >> 8) & 1, ti = i >> 9;
  const int half = ti & 1, kh = (ti >> 1) % 3, kd = ti / 6;
  const int kw = 2 * half + kc;  // 0..3
  float v = 0.f;
  if (kw < 3 && c < Cin) v = w[(co * Cin + c) * 27 + kd * 9 + kh * 3 + kw];
  out[i] = __float2bfloat16_rn(v);
}
// Conv3d weights with 32 input channels [Cout][32][3][3][3] -> pair layout [Cout][2][9][64] for the slab kernel: K chunk
// (item, tap) = [depth tap 2*item, 32 channels | depth tap 2*item + 1, 32 channels] (the missing fourth depth tap is zero).
__global__ void wpack_conv3d_pair_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int Cout) {
  const long long total = static_cast<long long>(Cout) * 2 * 9 * 64;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int k = i & 63;
    const int tp = (i >> 6) % 9;
    const int item = (i / (64 * 9)) & 1;
    const long long co = i / (2 * 9 * 64);
    const int kd = 2 * item + (k >> 5), ci = k & 31;
    out[i] = __float2bfloat16_rn(kd < 3 ? w[(co * 32 + ci) * 27 + kd * 9 + tp] : 0.f);
  }
}
// Stem 7x7x3 filter -> [Cout][8 row-taps][32 = (s, c4)] with zero padding (s == 7, c == 3, r == 7).
__global__ void wpack_stem_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wf, int Cout, int Cin,
                                  int R, int S) {
  const int total = Cout * 8 * 32;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = i & 3, s = (i >> 2) & 7, r = (i >> 5) & 7, co = i >> 8;
    float v = 0.f;
    if (c < Cin && s < S && r < R) v = w[((co * Cin + c) * R + r) * S + s];
    wf[i] = __float2bfloat16_rn(v);
  }
}
// Inverse map for the stem weight gradient: g8[Cout][Cin=32 -> (s,c4)][8 taps r] (layout produced by
// splitk_reduce_wgrad_kernel with cin=32, ntaps=8) -> grad[Cout][3][7][7].
__global__ void stem_wgrad_unpack_kernel(const float* __restrict__ g8, float* __restrict__ grad, int Cout, int Cin,
                                         int R, int S, int accumulate) {
  const int total = Cout * Cin * R * S;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int s = i % S, r = (i / S) % R, c = (i / (S * R)) % Cin, co = i / (S * R * Cin);
    const float v = g8[(co * 32 + (s * 4 + c)) * 8 + r];
    grad[i] = accumulate ? grad[i] + v : v;
  }
}

// ---------------------------------------------------------------------------------------------
// Column reductions over tiles of partial sums: in[T][K] (K = 2*C, or C) -> out[K] (double
// accumulation, fixed order => deterministic). Stage 1 reduces T -> S slices, stage 2 finishes.
// ---------------------------------------------------------------------------------------------
__global__ void colreduce_stage1_kernel(const float* __restrict__ in, int T, int K, int S, double* __restrict__ out) {
  // grid (ceil(K/32), S), block (32, 8): slice s of the rows, 8 row lanes, fp64 combine in fixed order
  __shared__ double sh[8][32];
  const int k = blockIdx.x * 32 + threadIdx.x;
  const int s = blockIdx.y;
  const int per = (T + S - 1) / S;
  const int t0 = s * per, t1 = min(T, t0 + per);
  double acc = 0.0;
  if (k < K) {
    float f = 0.f;
    int cnt = 0;
    for (int t = t0 + threadIdx.y; t < t1; t += 8) {
      f += in[static_cast<long long>(t) * K + k];
      if (++cnt == 16) { acc += f; f = 0.f; cnt = 0; }
    }
    acc += f;
  }
  sh[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && k < K) {
    double tot = 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) tot += sh[j][threadIdx.x];
    out[static_cast<long long>(s) * K + k] = tot;
  }
}

// BatchNorm finalize (train mode): sums[S][2][C] (double) over `count` rows per channel.
//   mean, biased var -> invstd ; scale = gamma*invstd ; shift = beta - mean*scale
//   running_mean/var updated with momentum (unbiased var), as torch.nn.BatchNorm*d does.
__global__ void bn_finalize_kernel(const double* __restrict__ sums, int S, int C, double count,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                   float momentum, float* __restrict__ running_mean, float* __restrict__ running_var,
                                   float* __restrict__ mean_out, float* __restrict__ invstd_out,
                                   float* __restrict__ scale_out, float* __restrict__ shift_out,
                                   long long* __restrict__ num_batches_tracked) {
  if (num_batches_tracked && blockIdx.x == 0 && threadIdx.x == 0) *num_batches_tracked += 1;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s1 = 0.0, s2 = 0.0;
  for (int s = 0; s < S; ++s) {
    s1 += sums[(static_cast<long long>(s) * 2 + 0) * C + c];
    s2 += sums[(static_cast<long long>(s) * 2 + 1) * C + c];
  }
  const double mean = s1 / count;
  double var = s2 / count - mean * mean;
  if (var < 0.0) var = 0.0;
  const float invstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  mean_out[c] = static_cast<float>(mean);
  invstd_out[c] = invstd;
  scale_out[c] = g * invstd;
  shift_out[c] = b - static_cast<float>(mean) * g * invstd;
  if (running_mean) {
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * static_cast<float>(mean);
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * static_cast<float>(unbiased);
  }
}
// Block (32 channels x 32 row lanes): per-channel sums of partial[T][2][C] (fp32 rows, fp64 combine, fixed order).
// Each thread sums every 32nd row with four independent loads in flight, so T ~ 600 rows take ~5 iterations.
__device__ __forceinline__ void reduce_rows_2(const float* __restrict__ partial, int T, int C, int c, int ry,
                                              double& s1, double& s2) {
  __shared__ double sh[2][32][33];
  double a1 = 0.0, a2 = 0.0;
  if (c < C) {
    const long long stride = 2LL * C;
    const float* base = partial + c;
    int t = ry;
    for (; t + 96 < T; t += 128) {
      const float x0 = base[t * stride], y0 = base[t * stride + C];
      const float x1 = base[(t + 32) * stride], y1 = base[(t + 32) * stride + C];
      const float x2 = base[(t + 64) * stride], y2 = base[(t + 64) * stride + C];
      const float x3 = base[(t + 96) * stride], y3 = base[(t + 96) * stride + C];
      a1 += static_cast<double>((x0 + x1) + (x2 + x3));
      a2 += static_cast<double>((y0 + y1) + (y2 + y3));
    }
    for (; t < T; t += 32) {
      a1 += static_cast<double>(base[t * stride]);
      a2 += static_cast<double>(base[t * stride + C]);
    }
  }
  sh[0][ry][threadIdx.x] = a1;
  sh[1][ry][threadIdx.x] = a2;
  __syncthreads();
  s1 = 0.0; s2 = 0.0;
  if (ry == 0) {
#pragma unroll 8
    for (int j = 0; j < 32; ++j) { s1 += sh[0][j][threadIdx.x]; s2 += sh[1][j][threadIdx.x]; }
  }
}
// BatchNorm finalize straight from fp32 partial rows [T][2][C]. grid = ceil(C/32), block = (32, 32).
__global__ void bn_finalize_rows_kernel(const float* __restrict__ partial, int T, int C, double count,
                                        const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                        float momentum, float* __restrict__ running_mean, float* __restrict__ running_var,
                                        float* __restrict__ mean_out, float* __restrict__ invstd_out,
                                        float* __restrict__ scale_out, float* __restrict__ shift_out,
                                        long long* __restrict__ num_batches_tracked) {
  const int c = blockIdx.x * 32 + threadIdx.x;
  if (num_batches_tracked && blockIdx.x == 0 && threadIdx.x == 0 && threadIdx.y == 0) *num_batches_tracked += 1;
  double s1, s2;
  reduce_rows_2(partial, T, C, c, threadIdx.y, s1, s2);
  if (threadIdx.y != 0 || c >= C) return;
  const double mean = s1 / count;
  double var = s2 / count - mean * mean;
  if (var < 0.0) var = 0.0;
  const float invstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  mean_out[c] = static_cast<float>(mean);
  invstd_out[c] = invstd;
  scale_out[c] = g * invstd;
  shift_out[c] = b - static_cast<float>(mean) * g * invstd;
  if (running_mean) {
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * static_cast<float>(mean);
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * static_cast<float>(unbiased);
  }
}
// Also folds everything pass 2 needs into three per-channel coefficients:
//   dy = g*is*(dz - c1 - (y - mu)*is*c2) = A*dz + B*y + K,  A = g*is, B = -g*is^2*c2, K = g*is*(mu*is*c2 - c1).
__global__ void bn_bwd_finalize_rows_kernel(const float* __restrict__ partial, int T, int C, double count,
                                            const float* __restrict__ mean, const float* __restrict__ invstd,
                                            const float* __restrict__ gamma, float* __restrict__ dgamma,
                                            float* __restrict__ dbeta, int accumulate, int eval_mode,
                                            float* __restrict__ coef, float* __restrict__ dbias = nullptr) {
  const int c = blockIdx.x * 32 + threadIdx.x;
  double s1, s2;
  reduce_rows_2(partial, T, C, c, threadIdx.y, s1, s2);
  if (threadIdx.y != 0 || c >= C) return;
  // gradient of a bias added in front of this BatchNorm = sum over positions of dy = g*is*(s1 - count*c1 - c2*sum xhat):
  // exactly zero under batch statistics (sum xhat = 0), g*is*s1 under running statistics
  if (dbias) dbias[c] = eval_mode ? static_cast<float>((gamma ? gamma[c] : 1.0) * invstd[c] * s1) : 0.f;
  if (dbeta) dbeta[c] = accumulate ? dbeta[c] + static_cast<float>(s1) : static_cast<float>(s1);
  if (dgamma) dgamma[c] = accumulate ? dgamma[c] + static_cast<float>(s2) : static_cast<float>(s2);
  const double c1 = eval_mode ? 0.0 : s1 / count;
  const double c2 = eval_mode ? 0.0 : s2 / count;
  const double g = gamma ? gamma[c] : 1.0, is = invstd[c], mu = mean[c];
  coef[c] = static_cast<float>(g * is);
  coef[C + c] = static_cast<float>(-g * is * is * c2);
  coef[2 * C + c] = static_cast<float>(g * is * (mu * is * c2 - c1));
}
// Eval mode: scale/shift from running statistics.
__global__ void bn_eval_coeffs_kernel(int C, const float* __restrict__ gamma, const float* __restrict__ beta,
                                      const float* __restrict__ running_mean, const float* __restrict__ running_var,
                                      float eps, float* __restrict__ mean_out, float* __restrict__ invstd_out,
                                      float* __restrict__ scale_out, float* __restrict__ shift_out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float invstd = 1.0f / sqrtf(running_var[c] + eps);
  if (mean_out) mean_out[c] = running_mean[c];
  if (invstd_out) invstd_out[c] = invstd;
  const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  scale_out[c] = g * invstd;
  shift_out[c] = b - running_mean[c] * g * invstd;
}

// Per-channel sum / sum of squares of a dense bf16 [M][C] tensor -> partial[blocks][2][C] (used when the
// statistics are not produced by a GEMM epilogue). Block = 256 threads: (C/8) channel groups x row lanes.
__global__ void bn_stats_kernel(const __nv_bfloat16* __restrict__ y, long long M, int C, float* __restrict__ partial) {
  extern __shared__ float sm[];  // [lanes][2][C]
  const int groups = C / 8;
  const int lanes = blockDim.x / groups;
  const int cg = threadIdx.x % groups, rl = threadIdx.x / groups;
  float s1[8] = {0}, s2[8] = {0};
  if (rl < lanes) {
    for (long long r = static_cast<long long>(blockIdx.x) * lanes + rl; r < M; r += static_cast<long long>(gridDim.x) * lanes) {
      float f[8];
      unpack8(ld_nc16(y + r * C + cg * 8), f);
#pragma unroll
      for (int e = 0; e < 8; ++e) { s1[e] += f[e]; s2[e] += f[e] * f[e]; }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      sm[(rl * 2 + 0) * C + cg * 8 + e] = s1[e];
      sm[(rl * 2 + 1) * C + cg * 8 + e] = s2[e];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    float acc = 0.f;
    for (int l = 0; l < lanes; ++l) acc += sm[l * 2 * C + i];
    partial[static_cast<long long>(blockIdx.x) * 2 * C + i] = acc;
  }
}

// out = act(y*scale[c] + shift[c] (+ residual)); dense [M][C] bf16, C % 8 == 0.
__global__ void bn_apply_kernel(const __nv_bfloat16* __restrict__ y, const float* __restrict__ scale,
                                const float* __restrict__ shift, const __nv_bfloat16* __restrict__ residual,
                                __nv_bfloat16* __restrict__ out, long long total8, int C, int relu) {
  // When the grid stride is a multiple of C/8 vectors, a thread sees the same 8 channels in every iteration: the
  // coefficients are loaded once (otherwise 4 extra L1 loads per 16-byte element vector).
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long i0 = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const bool fixed = ((stride * 8) % C) == 0;
  float sc[8], sh[8];
  auto load_coef = [&](long long i) {
    const int c0 = static_cast<int>((i * 8) % C);
    *reinterpret_cast<float4*>(sc) = __ldg(reinterpret_cast<const float4*>(scale + c0));
    *reinterpret_cast<float4*>(sc + 4) = __ldg(reinterpret_cast<const float4*>(scale + c0 + 4));
    *reinterpret_cast<float4*>(sh) = __ldg(reinterpret_cast<const float4*>(shift + c0));
    *reinterpret_cast<float4*>(sh + 4) = __ldg(reinterpret_cast<const float4*>(shift + c0 + 4));
  };
  if (fixed && i0 < total8) load_coef(i0);
  auto one = [&](long long i, const uint4& qy, const uint4& qr) {
    float f[8], r[8];
    unpack8(qy, f);
    if (residual) unpack8(qr, r);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float v = fmaf(f[e], sc[e], sh[e]);
      if (residual) v += r[e];
      if (relu) v = fmaxf(v, 0.f);
      f[e] = v;
    }
    *reinterpret_cast<uint4*>(out + i * 8) = pack8(f);
  };
  long long i = i0;
  if (fixed) {
    // two vectors per iteration: all loads of both are issued before either is consumed
    for (; i + stride < total8; i += 2 * stride) {
      const long long j = i + stride;
      const uint4 y0 = ld_nc16(y + i * 8), y1 = ld_nc16(y + j * 8);
      uint4 r0 = make_uint4(0, 0, 0, 0), r1 = r0;
      if (residual) { r0 = ld_nc16(residual + i * 8); r1 = ld_nc16(residual + j * 8); }
      one(i, y0, r0);
      one(j, y1, r1);
    }
  }
  for (; i < total8; i += stride) {
    if (!fixed) load_coef(i);
    const uint4 y0 = ld_nc16(y + i * 8);
    uint4 r0 = make_uint4(0, 0, 0, 0);
    if (residual) r0 = ld_nc16(residual + i * 8);
    one(i, y0, r0);
  }
}

// BatchNorm backward, pass 1: dz = dout * (act > 0) (mask only when act != nullptr);
// partial[blocks][2][C] = { sum dz, sum dz * xhat }, xhat = (y - mean) * invstd.
// ReLU mask: from the stored activation (act > 0) or, when act == nullptr and msc != nullptr, recomputed from
// the raw conv output (y*msc + msh > 0 — the forward's own expression, so the mask is identical and the
// activation tensor need not be read).
template <int kRows>  // rows in flight per thread and iteration (up to 3 * kRows independent 16-byte loads)
__global__ void bn_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ dout, const __nv_bfloat16* __restrict__ act,
                                     const __nv_bfloat16* __restrict__ y, const float* __restrict__ mean,
                                     const float* __restrict__ invstd, const float* __restrict__ msc,
                                     const float* __restrict__ msh, long long M, int C,
                                     float* __restrict__ partial) {
  extern __shared__ float sm[];
  const int groups = C / 8;
  const int lanes = blockDim.x / groups;
  const int cg = threadIdx.x % groups, rl = threadIdx.x / groups;
  float s1[8] = {0}, s2[8] = {0};
  if (rl < lanes) {
    float mu[8], is[8], ms[8], mh[8];
    const bool ymask = (act == nullptr) && (msc != nullptr);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      mu[e] = mean[cg * 8 + e]; is[e] = invstd[cg * 8 + e];
      ms[e] = ymask ? msc[cg * 8 + e] : 0.f; mh[e] = ymask ? msh[cg * 8 + e] : 1.f;
    }
    const long long step = static_cast<long long>(gridDim.x) * lanes;
    for (long long r = static_cast<long long>(blockIdx.x) * lanes + rl; r < M; r += kRows * step) {
      uint4 qd[kRows], qv[kRows], qa[kRows];
      bool has[kRows];
#pragma unroll
      for (int k = 0; k < kRows; ++k) {
        const long long rk = r + k * step;
        has[k] = rk < M;
        qd[k] = qv[k] = qa[k] = make_uint4(0, 0, 0, 0);
        if (has[k]) {
          qd[k] = ld_nc16(dout + rk * C + cg * 8);
          qv[k] = ld_nc16(y + rk * C + cg * 8);
          if (act) qa[k] = ld_nc16(act + rk * C + cg * 8);
        }
      }
#pragma unroll
      for (int k = 0; k < kRows; ++k) {
        if (!has[k]) continue;
        float d[8], a[8], v[8];
        unpack8(qd[k], d); unpack8(qv[k], v); unpack8(qa[k], a);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const bool off = act ? !(a[e] > 0.f) : !(fmaf(v[e], ms[e], mh[e]) > 0.f);
          const float dz = off ? 0.f : d[e];
          s1[e] += dz;
          s2[e] = fmaf(dz, v[e] - mu[e], s2[e]);  // invstd is applied once per thread below, not per element
        }
      }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      sm[(rl * 2 + 0) * C + cg * 8 + e] = s1[e];
      sm[(rl * 2 + 1) * C + cg * 8 + e] = s2[e] * is[e];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    float acc = 0.f;
    for (int l = 0; l < lanes; ++l) acc += sm[l * 2 * C + i];
    partial[static_cast<long long>(blockIdx.x) * 2 * C + i] = acc;
  }
}
// Finish the reduction: sums[S][2][C] -> dgamma, dbeta (fp32, optional accumulate) and the two
// per-channel coefficients used by pass 2: c1 = sum_dz / M, c2 = sum_dz_xhat / M.
__global__ void bn_bwd_finalize_kernel(const double* __restrict__ sums, int S, int C, double count,
                                       float* __restrict__ dgamma, float* __restrict__ dbeta, int accumulate,
                                       int eval_mode, float* __restrict__ c1, float* __restrict__ c2) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s1 = 0.0, s2 = 0.0;
  for (int s = 0; s < S; ++s) {
    s1 += sums[(static_cast<long long>(s) * 2 + 0) * C + c];
    s2 += sums[(static_cast<long long>(s) * 2 + 1) * C + c];
  }
  if (dbeta) dbeta[c] = accumulate ? dbeta[c] + static_cast<float>(s1) : static_cast<float>(s1);
  if (dgamma) dgamma[c] = accumulate ? dgamma[c] + static_cast<float>(s2) : static_cast<float>(s2);
  // eval mode: statistics are constants, so the two batch-coupling terms vanish
  c1[c] = eval_mode ? 0.f : static_cast<float>(s1 / count);
  c2[c] = eval_mode ? 0.f : static_cast<float>(s2 / count);
}
// Pass 2: dy = A[c]*dz + B[c]*y + K[c] (coef = [A | B | K], see bn_bwd_finalize_rows_kernel); optionally also
// writes dz (gradient of the residual/identity branch).
__global__ void bn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dout, const __nv_bfloat16* __restrict__ act,
                                    const __nv_bfloat16* __restrict__ y, const float* __restrict__ coef,
                                    const float* __restrict__ msc, const float* __restrict__ msh,
                                    __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ dz_out,
                                    long long total8, int C) {
  const bool ymask = (act == nullptr) && (msc != nullptr);
  // per-thread channel group is loop-invariant when the grid stride is a multiple of C/8 vectors (see bn_apply_kernel)
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long i0 = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const bool fixed = ((stride * 8) % C) == 0;
  float ca[8], cb[8], ck[8], ms[8], mh[8];
  auto load_coef = [&](long long i) {
    const int c0 = static_cast<int>((i * 8) % C);
#pragma unroll
    for (int e = 0; e < 8; e += 4) {
      *reinterpret_cast<float4*>(ca + e) = __ldg(reinterpret_cast<const float4*>(coef + c0 + e));
      *reinterpret_cast<float4*>(cb + e) = __ldg(reinterpret_cast<const float4*>(coef + C + c0 + e));
      *reinterpret_cast<float4*>(ck + e) = __ldg(reinterpret_cast<const float4*>(coef + 2 * C + c0 + e));
      if (ymask) {
        *reinterpret_cast<float4*>(ms + e) = __ldg(reinterpret_cast<const float4*>(msc + c0 + e));
        *reinterpret_cast<float4*>(mh + e) = __ldg(reinterpret_cast<const float4*>(msh + c0 + e));
      }
    }
  };
  if (fixed && i0 < total8) load_coef(i0);
  auto one = [&](long long i, const uint4& qd, const uint4& qv, const uint4& qa) {
    float d[8], a[8], v[8], o[8];
    unpack8(qd, d);
    unpack8(qv, v);
    if (act) unpack8(qa, a);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const bool off = act ? !(a[e] > 0.f) : (ymask && !(fmaf(v[e], ms[e], mh[e]) > 0.f));
      const float dzv = off ? 0.f : d[e];
      o[e] = fmaf(ca[e], dzv, fmaf(cb[e], v[e], ck[e]));
      d[e] = dzv;
    }
    *reinterpret_cast<uint4*>(dy + i * 8) = pack8(o);
    if (dz_out) *reinterpret_cast<uint4*>(dz_out + i * 8) = pack8(d);
  };
  // (two vectors per iteration, as in bn_apply_kernel, measured 4 % slower here: five streams per thread already)
  long long i = i0;
  for (; i < total8; i += stride) {
    if (!fixed) load_coef(i);
    const uint4 d0 = ld_nc16(dout + i * 8), v0 = ld_nc16(y + i * 8);
    uint4 a0 = make_uint4(0, 0, 0, 0);
    if (act) a0 = ld_nc16(act + i * 8);
    one(i, d0, v0, a0);
  }
}

// dz = dout * (act > 0): ReLU backward for layers without BatchNorm.
__global__ void relu_bwd_kernel(const __nv_bfloat16* __restrict__ dout, const __nv_bfloat16* __restrict__ act,
                                __nv_bfloat16* __restrict__ dz, long long total8) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total8;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float d[8], a[8];
    unpack8(ld_nc16(dout + i * 8), d);
    unpack8(ld_nc16(act + i * 8), a);
#pragma unroll
    for (int e = 0; e < 8; ++e) d[e] = (a[e] > 0.f) ? d[e] : 0.f;
    *reinterpret_cast<uint4*>(dz + i * 8) = pack8(d);
  }
}

// Column sums of a dense bf16 [M][C] tensor -> partial[blocks][C] (bias gradients). blockIdx.y walks
// channel chunks of Cc <= 2048 channels so any C % 8 == 0 is supported.
__global__ void colsum_bf16_kernel(const __nv_bfloat16* __restrict__ x, long long M, int C, int Cc,
                                   float* __restrict__ partial) {
  extern __shared__ float sm[];
  const int cbeg = blockIdx.y * Cc;
  const int cw = min(Cc, C - cbeg);
  const int groups = cw / 8;
  const int lanes = blockDim.x / groups;
  const int cg = threadIdx.x % groups, rl = threadIdx.x / groups;
  float s1[8] = {0};
  if (rl < lanes) {
    for (long long r = static_cast<long long>(blockIdx.x) * lanes + rl; r < M; r += static_cast<long long>(gridDim.x) * lanes) {
      float f[8];
      unpack8(ld_nc16(x + r * C + cbeg + cg * 8), f);
#pragma unroll
      for (int e = 0; e < 8; ++e) s1[e] += f[e];
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) sm[rl * cw + cg * 8 + e] = s1[e];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < cw; i += blockDim.x) {
    float acc = 0.f;
    for (int l = 0; l < lanes; ++l) acc += sm[l * cw + i];
    partial[static_cast<long long>(blockIdx.x) * C + cbeg + i] = acc;
  }
}
__global__ void colreduce_final_f32_kernel(const double* __restrict__ sums, int S, int K, float* __restrict__ out,
                                           int accumulate) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  double acc = 0.0;
  for (int s = 0; s < S; ++s) acc += sums[static_cast<long long>(s) * K + k];
  out[k] = accumulate ? out[k] + static_cast<float>(acc) : static_cast<float>(acc);
}

// ---------------------------------------------------------------------------------------------
// MaxPool2d(kernel, stride, pad) on NHWC bf16 with an int8 argmax plane (first maximum in
// row-major window order, like ATen's max_pool2d_with_indices), C % 8 == 0.
// ---------------------------------------------------------------------------------------------
__global__ void maxpool2d_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                     signed char* __restrict__ argmax, int N, int H, int W, int C, int Ho, int Wo,
                                     int ksize, int stride, int pad) {
  const int groups = C / 8;
  const long long total = static_cast<long long>(N) * Ho * Wo * groups;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cg = i % groups;
    const int wo = (i / groups) % Wo;
    const int ho = (i / (static_cast<long long>(groups) * Wo)) % Ho;
    const int n = i / (static_cast<long long>(groups) * Wo * Ho);
    float best[8];
    int bi[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { best[e] = -INFINITY; bi[e] = -1; }
    for (int r = 0; r < ksize; ++r) {
      const int h = ho * stride - pad + r;
      if (h < 0 || h >= H) continue;
      for (int s = 0; s < ksize; ++s) {
        const int w = wo * stride - pad + s;
        if (w < 0 || w >= W) continue;
        float f[8];
        unpack8(ld_nc16(x + ((static_cast<long long>(n) * H + h) * W + w) * C + cg * 8), f);
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (f[e] > best[e] || bi[e] < 0) { best[e] = f[e]; bi[e] = r * ksize + s; }
      }
    }
    const long long o = ((static_cast<long long>(n) * Ho + ho) * Wo + wo) * C + cg * 8;
    *reinterpret_cast<uint4*>(out + o) = pack8(best);
    if (argmax) {
      uint2 pk;
      pk.x = (bi[0] & 0xff) | ((bi[1] & 0xff) << 8) | ((bi[2] & 0xff) << 16) | ((bi[3] & 0xff) << 24);
      pk.y = (bi[4] & 0xff) | ((bi[5] & 0xff) << 8) | ((bi[6] & 0xff) << 16) | ((bi[7] & 0xff) << 24);
      *reinterpret_cast<uint2*>(argmax + o) = pk;
    }
  }
}
// Scatter-free backward: each input pixel gathers from the output windows that cover it.
__global__ void maxpool2d_bwd_kernel(const __nv_bfloat16* __restrict__ dout, const signed char* __restrict__ argmax,
                                     __nv_bfloat16* __restrict__ dx, int N, int H, int W, int C, int Ho, int Wo,
                                     int ksize, int stride, int pad) {
  const int groups = C / 8;
  const long long total = static_cast<long long>(N) * H * W * groups;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cg = i % groups;
    const int w = (i / groups) % W;
    const int h = (i / (static_cast<long long>(groups) * W)) % H;
    const int n = i / (static_cast<long long>(groups) * W * H);
    float acc[8] = {0};
    // output rows ho with ho*stride - pad <= h <= ho*stride - pad + ksize - 1
    const int ho_lo = max(0, (h + pad - ksize + stride) / stride);
    const int ho_hi = min(Ho - 1, (h + pad) / stride);
    const int wo_lo = max(0, (w + pad - ksize + stride) / stride);
    const int wo_hi = min(Wo - 1, (w + pad) / stride);
    for (int ho = ho_lo; ho <= ho_hi; ++ho) {
      const int r = h - (ho * stride - pad);
      for (int wo = wo_lo; wo <= wo_hi; ++wo) {
        const int s = w - (wo * stride - pad);
        const int code = r * ksize + s;
        const long long o = ((static_cast<long long>(n) * Ho + ho) * Wo + wo) * C + cg * 8;
        const uint2 pk = *reinterpret_cast<const uint2*>(argmax + o);
        float d[8];
        unpack8(ld_nc16(dout + o), d);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int a = (e < 4 ? (pk.x >> (8 * e)) : (pk.y >> (8 * (e - 4)))) & 0xff;
          if (a == code) acc[e] += d[e];
        }
      }
    }
    *reinterpret_cast<uint4*>(dx + i * 8) = pack8(acc);
  }
}

// ---------------------------------------------------------------------------------------------
// Stem tail fused: BatchNorm-apply + ReLU + MaxPool2d(3,2,1) in one pass over the raw conv output, and its
// backward (max-pool gather + ReLU mask recomputed from y + BatchNorm backward) without ever materialising the
// 112x112 activated map or its gradient (torchvision resnet.py:198-200).
// ---------------------------------------------------------------------------------------------
__global__ void bn_relu_maxpool_fwd_kernel(const __nv_bfloat16* __restrict__ y, const float* __restrict__ scale,
                                           const float* __restrict__ shift, __nv_bfloat16* __restrict__ out,
                                           signed char* __restrict__ argmax, __nv_bfloat16* __restrict__ yarg, int N, int H, int W,
                                           int C, int Ho, int Wo) {
  const int groups = C / 8;
  const long long total = static_cast<long long>(N) * Ho * Wo * groups;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cg = i % groups;
    const int wo = (i / groups) % Wo;
    const int ho = (i / (static_cast<long long>(groups) * Wo)) % Ho;
    const int n = i / (static_cast<long long>(groups) * Wo * Ho);
    float sc[8], sh[8];
    *reinterpret_cast<float4*>(sc) = __ldg(reinterpret_cast<const float4*>(scale + cg * 8));
    *reinterpret_cast<float4*>(sc + 4) = __ldg(reinterpret_cast<const float4*>(scale + cg * 8 + 4));
    *reinterpret_cast<float4*>(sh) = __ldg(reinterpret_cast<const float4*>(shift + cg * 8));
    *reinterpret_cast<float4*>(sh + 4) = __ldg(reinterpret_cast<const float4*>(shift + cg * 8 + 4));
    // all nine window loads are issued before any arithmetic (clamped coordinates + validity mask)
    uint4 v[9];
    unsigned okmask = 0;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int h = ho * 2 - 1 + r;
      const int hc = min(max(h, 0), H - 1);
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const int w = wo * 2 - 1 + q;
        const int wc = min(max(w, 0), W - 1);
        if (h == hc && w == wc) okmask |= 1u << (r * 3 + q);
        v[r * 3 + q] = ld_nc16(y + ((static_cast<long long>(n) * H + hc) * W + wc) * C + cg * 8);
      }
    }
    // running max / arg-max on packed bf16 pairs: the activated value is a bf16 (same rounding as the unfused
    // path), a > best is strict so the first maximum wins, and best starts at -inf so the first valid tap always
    // takes (activations are >= 0)
    unsigned best2[4] = {0xff80ff80u, 0xff80ff80u, 0xff80ff80u, 0xff80ff80u}, bi2[4] = {0u, 0u, 0u, 0u}, ya2[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      if (!((okmask >> t) & 1u)) continue;
      const unsigned raw[4] = {v[t].x, v[t].y, v[t].z, v[t].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float lo = __uint_as_float(raw[j] << 16), hi = __uint_as_float(raw[j] & 0xffff0000u);
        const unsigned a2 = pack_bf16x2(fmaxf(fmaf(lo, sc[2 * j], sh[2 * j]), 0.f),
                                        fmaxf(fmaf(hi, sc[2 * j + 1], sh[2 * j + 1]), 0.f));
        const unsigned m = __hgt2_mask(*reinterpret_cast<const __nv_bfloat162*>(&a2),
                                       *reinterpret_cast<const __nv_bfloat162*>(&best2[j]));
        best2[j] = (a2 & m) | (best2[j] & ~m);
        bi2[j] = ((t * 0x00010001u) & m) | (bi2[j] & ~m);
        ya2[j] = (raw[j] & m) | (ya2[j] & ~m);
      }
    }
    const long long o = ((static_cast<long long>(n) * Ho + ho) * Wo + wo) * C + cg * 8;
    *reinterpret_cast<uint4*>(out + o) = make_uint4(best2[0], best2[1], best2[2], best2[3]);
    // raw conv output at the arg-max: lets the backward form its BatchNorm statistics from pooled-size tensors
    if (yarg) *reinterpret_cast<uint4*>(yarg + o) = make_uint4(ya2[0], ya2[1], ya2[2], ya2[3]);
    if (argmax) {
      uint2 pk;
      pk.x = __byte_perm(bi2[0], bi2[1], 0x6420);
      pk.y = __byte_perm(bi2[2], bi2[3], 0x6420);
      *reinterpret_cast<uint2*>(argmax + o) = pk;
    }
  }
}
// Backward works on 2x2 input blocks (rows 2a,2a+1 x cols 2b,2b+1; H and W even): exactly four pooling windows
// (a,b) (a,b+1) (a+1,b) (a+1,b+1) cover such a block, and each of its pixels sits at a fixed window position
// (arg-max code r*3+q) in each of them, so one thread issues all 12 loads (4 y, 4 dpool, 4 arg-max words) up
// front and then only compares codes. dz = pooled gradient gathered through the arg-max plane, times the ReLU
// mask recomputed from y.
struct StemQuad {
  uint4 y[4];   // pixels (2a,2b) (2a,2b+1) (2a+1,2b) (2a+1,2b+1)
  uint4 d[4];   // windows (a,b) (a,b+1) (a+1,b) (a+1,b+1)
  uint2 am[4];
};
__device__ __forceinline__ void stem_quad_load(const __nv_bfloat16* __restrict__ dpool, const signed char* __restrict__ argmax,
                                               const __nv_bfloat16* __restrict__ y, long long quad, int cg, int W, int C, int Ho,
                                               int Wo, StemQuad& q) {
  const int b = static_cast<int>(quad % Wo);
  const long long na = quad / Wo;  // n*Ho + a
  const int a = static_cast<int>(na % Ho);
  const bool hb = a + 1 < Ho, wb = b + 1 < Wo;
  const long long y0 = ((na * 2) * W + 2 * b) * C + cg * 8;  // rows of image n are contiguous: n*H + 2a == 2*(n*Ho + a)
  const long long o0 = (na * Wo + b) * C + cg * 8;
  q.y[0] = ld_nc16(y + y0);
  q.y[1] = ld_nc16(y + y0 + C);
  q.y[2] = ld_nc16(y + y0 + static_cast<long long>(W) * C);
  q.y[3] = ld_nc16(y + y0 + static_cast<long long>(W) * C + C);
  const long long o[4] = {o0, o0 + C, o0 + static_cast<long long>(Wo) * C, o0 + static_cast<long long>(Wo) * C + C};
  const bool ok[4] = {true, wb, hb, hb && wb};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    q.d[k] = ok[k] ? ld_nc16(dpool + o[k]) : make_uint4(0, 0, 0, 0);
    q.am[k] = ok[k] ? __ldg(reinterpret_cast<const uint2*>(argmax + o[k])) : make_uint2(0xffffffffu, 0xffffffffu);
  }
}
// dz += d where the window's arg-max code equals `code`
__device__ __forceinline__ void stem_take(const uint2 am, const float (&d)[8], unsigned code, float (&dz)[8]) {
  const unsigned rep = code * 0x01010101u;
  const unsigned m0 = __vcmpeq4(am.x, rep), m1 = __vcmpeq4(am.y, rep);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    if (m0 & (1u << (8 * e))) dz[e] += d[e];
    if (m1 & (1u << (8 * e))) dz[4 + e] += d[4 + e];
  }
}
__device__ __forceinline__ void stem_quad_dz(const StemQuad& q, const float (&sc)[8], const float (&sh)[8], float (&yv)[4][8],
                                             float (&dz)[4][8]) {
  float d[4][8];
#pragma unroll
  for (int k = 0; k < 4; ++k) { unpack8(q.d[k], d[k]); unpack8(q.y[k], yv[k]); }
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int e = 0; e < 8; ++e) dz[k][e] = 0.f;
  // pixel (2a,2b): window (a,b) position (1,1)
  stem_take(q.am[0], d[0], 4, dz[0]);
  // pixel (2a,2b+1): (a,b) at (1,2); (a,b+1) at (1,0)
  stem_take(q.am[0], d[0], 5, dz[1]);
  stem_take(q.am[1], d[1], 3, dz[1]);
  // pixel (2a+1,2b): (a,b) at (2,1); (a+1,b) at (0,1)
  stem_take(q.am[0], d[0], 7, dz[2]);
  stem_take(q.am[2], d[2], 1, dz[2]);
  // pixel (2a+1,2b+1): (a,b) at (2,2); (a,b+1) at (2,0); (a+1,b) at (0,2); (a+1,b+1) at (0,0)
  stem_take(q.am[0], d[0], 8, dz[3]);
  stem_take(q.am[1], d[1], 6, dz[3]);
  stem_take(q.am[2], d[2], 2, dz[3]);
  stem_take(q.am[3], d[3], 0, dz[3]);
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int e = 0; e < 8; ++e)
      if (!(fmaf(yv[k][e], sc[e], sh[e]) > 0.f)) dz[k][e] = 0.f;
}
// pass 1: partial[blocks][2][C] = { sum dz, sum dz*xhat }
__global__ void __launch_bounds__(256, 2)
stem_bn_pool_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ dpool, const signed char* __restrict__ argmax,
                               const __nv_bfloat16* __restrict__ y, const float* __restrict__ scale,
                               const float* __restrict__ shift, const float* __restrict__ mean,
                               const float* __restrict__ invstd, int N, int H, int W, int C, int Ho, int Wo,
                               float* __restrict__ partial) {
  extern __shared__ float sm[];
  const int groups = C / 8;
  const int lanes = blockDim.x / groups;
  const int cg = threadIdx.x % groups, rl = threadIdx.x / groups;
  float s1[8] = {0}, s2[8] = {0};
  if (rl < lanes) {
    float sc[8], sh[8], mu[8], is[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { sc[e] = scale[cg * 8 + e]; sh[e] = shift[cg * 8 + e]; mu[e] = mean[cg * 8 + e]; is[e] = invstd[cg * 8 + e]; }
    const long long Q = static_cast<long long>(N) * Ho * Wo;
    for (long long r = static_cast<long long>(blockIdx.x) * lanes + rl; r < Q; r += static_cast<long long>(gridDim.x) * lanes) {
      StemQuad q;
      stem_quad_load(dpool, argmax, y, r, cg, W, C, Ho, Wo, q);
      float yv[4][8], dz[4][8];
      stem_quad_dz(q, sc, sh, yv, dz);
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int e = 0; e < 8; ++e) { s1[e] += dz[k][e]; s2[e] += dz[k][e] * (yv[k][e] - mu[e]) * is[e]; }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      sm[(rl * 2 + 0) * C + cg * 8 + e] = s1[e];
      sm[(rl * 2 + 1) * C + cg * 8 + e] = s2[e];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    float acc = 0.f;
    for (int l = 0; l < lanes; ++l) acc += sm[l * 2 * C + i];
    partial[static_cast<long long>(blockIdx.x) * 2 * C + i] = acc;
  }
}
// pass 2: dy = A*dz + B*y + K (coef = [A | B | K])
__global__ void __launch_bounds__(256, 2)
stem_bn_pool_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dpool, const signed char* __restrict__ argmax,
                              const __nv_bfloat16* __restrict__ y, const float* __restrict__ scale,
                              const float* __restrict__ shift, const float* __restrict__ coef,
                              __nv_bfloat16* __restrict__ dy, int N, int H, int W, int C, int Ho, int Wo) {
  const int groups = C / 8;
  const long long total = static_cast<long long>(N) * Ho * Wo * groups;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cg = static_cast<int>(i % groups);
    const long long quad = i / groups;
    float sc[8], sh[8], ca[8], cb[8], ck[8];
#pragma unroll
    for (int e = 0; e < 8; e += 4) {
      *reinterpret_cast<float4*>(sc + e) = __ldg(reinterpret_cast<const float4*>(scale + cg * 8 + e));
      *reinterpret_cast<float4*>(sh + e) = __ldg(reinterpret_cast<const float4*>(shift + cg * 8 + e));
      *reinterpret_cast<float4*>(ca + e) = __ldg(reinterpret_cast<const float4*>(coef + cg * 8 + e));
      *reinterpret_cast<float4*>(cb + e) = __ldg(reinterpret_cast<const float4*>(coef + C + cg * 8 + e));
      *reinterpret_cast<float4*>(ck + e) = __ldg(reinterpret_cast<const float4*>(coef + 2 * C + cg * 8 + e));
    }
    StemQuad q;
    stem_quad_load(dpool, argmax, y, quad, cg, W, C, Ho, Wo, q);
    float yv[4][8], dz[4][8];
    stem_quad_dz(q, sc, sh, yv, dz);
    const int b = static_cast<int>(quad % Wo);
    const long long y0 = (((quad / Wo) * 2) * W + 2 * b) * C + cg * 8;
    const long long off[4] = {y0, y0 + C, y0 + static_cast<long long>(W) * C, y0 + static_cast<long long>(W) * C + C};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float o[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = fmaf(ca[e], dz[k][e], fmaf(cb[e], yv[k][e], ck[e]));
      *reinterpret_cast<uint4*>(dy + off[k]) = pack8(o);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Quadtree pooling stage of QuadtreeCNN (reference: Quadtree_from scratch/models.py:277-294).
//   q    : [4][B][QH][QW][Cq] bf16, quadrant conv output after bias+ReLU (quadrant-major)
//   l4   : [B][GH*GW][Cg] bf16, layer4 output
//   feat : [B][ldf] bf16, columns [0,Cg) = mean over GH*GW, then per quadrant (TL,TR,BL,BR)
//          Cq*PH*PW values in NCHW flatten order c*(PH*PW) + ph*PW + pw, MaxPool2d(2,2) floor mode.
// Generic-shape fallback: one thread per (b, quadrant, channel) / (b, global channel). The shapes of the reference
// (7x7x128 quadrants, 7x7x512 global map, 16-byte aligned feature rows) take the vectorised kernels below.
// ---------------------------------------------------------------------------------------------
__global__ void quadtree_pool_fwd_generic_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ l4,
                                         __nv_bfloat16* __restrict__ feat, int B, int QH, int QW, int Cq, int GHW,
                                         int Cg, int ldf) {
  const int PH = QH / 2, PW = QW / 2;
  const long long nq = static_cast<long long>(B) * 4 * Cq;
  const long long ng = static_cast<long long>(B) * Cg;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nq + ng;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    if (i < nq) {
      const int c = i % Cq;
      const int qi = (i / Cq) % 4;
      const int b = i / (static_cast<long long>(Cq) * 4);
      const __nv_bfloat16* src = q + ((static_cast<long long>(qi) * B + b) * QH * QW) * Cq + c;
      __nv_bfloat16* dst = feat + static_cast<long long>(b) * ldf + Cg + static_cast<long long>(qi) * Cq * PH * PW +
                           static_cast<long long>(c) * PH * PW;
      for (int ph = 0; ph < PH; ++ph)
        for (int pw = 0; pw < PW; ++pw) {
          float m = -INFINITY;
          for (int r = 0; r < 2; ++r)
            for (int s = 0; s < 2; ++s)
              m = fmaxf(m, __bfloat162float(src[((2 * ph + r) * QW + (2 * pw + s)) * static_cast<long long>(Cq)]));
          dst[ph * PW + pw] = __float2bfloat16_rn(m);
        }
    } else {
      const long long j = i - nq;
      const int c = j % Cg;
      const int b = j / Cg;
      const __nv_bfloat16* src = l4 + static_cast<long long>(b) * GHW * Cg + c;
      float acc = 0.f;
      for (int p = 0; p < GHW; ++p) acc += __bfloat162float(src[static_cast<long long>(p) * Cg]);
      feat[static_cast<long long>(b) * ldf + c] = __float2bfloat16_rn(acc / GHW);
    }
  }
}
// Backward (scatter-free broadcast): every quadrant-conv output element looks up whether it is the
// (first) arg-max of its pooling window and whether its ReLU was active; every layer4 element
// receives dfeat/GHW.  dfeat: [B][ldf] bf16.
__global__ void quadtree_pool_bwd_generic_kernel(const __nv_bfloat16* __restrict__ dfeat, const __nv_bfloat16* __restrict__ q,
                                         __nv_bfloat16* __restrict__ dq, __nv_bfloat16* __restrict__ dl4, int B,
                                         int QH, int QW, int Cq, int GHW, int Cg, int ldf) {
  const int PH = QH / 2, PW = QW / 2;
  const long long nq = static_cast<long long>(B) * 4 * QH * QW * Cq;
  const long long ng = static_cast<long long>(B) * GHW * Cg;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nq + ng;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    if (i < nq) {
      const int c = i % Cq;
      const int w = (i / Cq) % QW;
      const int h = (i / (static_cast<long long>(Cq) * QW)) % QH;
      const int b = (i / (static_cast<long long>(Cq) * QW * QH)) % B;
      const int qi = i / (static_cast<long long>(Cq) * QW * QH * B);
      const int ph = h >> 1, pw = w >> 1;
      float g = 0.f;
      if (ph < PH && pw < PW) {
        const __nv_bfloat16* win = q + (((static_cast<long long>(qi) * B + b) * QH + 2 * ph) * QW + 2 * pw) * Cq + c;
        const float v00 = __bfloat162float(win[0]);
        const float v01 = __bfloat162float(win[Cq]);
        const float v10 = __bfloat162float(win[static_cast<long long>(QW) * Cq]);
        const float v11 = __bfloat162float(win[static_cast<long long>(QW) * Cq + Cq]);
        int am = 0; float m = v00;
        if (v01 > m) { m = v01; am = 1; }
        if (v10 > m) { m = v10; am = 2; }
        if (v11 > m) { m = v11; am = 3; }
        const int me = (h & 1) * 2 + (w & 1);
        if (am == me && m > 0.f)
          g = __bfloat162float(dfeat[static_cast<long long>(b) * ldf + Cg + static_cast<long long>(qi) * Cq * PH * PW +
                                     static_cast<long long>(c) * PH * PW + ph * PW + pw]);
      }
      dq[i] = __float2bfloat16_rn(g);
    } else {
      const long long j = i - nq;
      const int c = j % Cg;
      const int b = j / (static_cast<long long>(Cg) * GHW);
      dl4[j] = __float2bfloat16_rn(__bfloat162float(dfeat[static_cast<long long>(b) * ldf + c]) / GHW);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Quadtree stage, vectorised (the kernel `north_star` describes): one CTA per (image, part), part = quadrant 0..3 or
// the global branch. Every global access is a 128-bit load/store on a 16-byte chunk of 8 channels, consecutive lanes
// on consecutive chunks (256 / 1024 contiguous bytes per pixel).
//   quadrant part : the QHxQWxCq tile (12.5 KB at 7x7x128) is staged in shared memory with coalesced 16-byte loads;
//                   thread t then produces the 8 consecutive outputs 8t..8t+7 of the NCHW-flatten order
//                   c*(PH*PW) + ph*PW + pw (the shared-memory read IS the transpose) and writes them with one 16-byte
//                   store, so the 2304-byte quadrant slice of the feature row leaves as full lines.
//   global part   : lane = 8 chunks x 4 pixel phases; each lane adds its pixels (p = phase, phase+4, ...) in fp32 and
//                   the four phases are combined with two warp shuffles (fixed order: deterministic).
// Requirements (checked by the host wrapper, else the generic kernels run): Cq % 8 == 0, Cg % 64 == 0, ldf % 8 == 0,
// (Cq*PH*PW) % 8 == 0, the quadrant tile fits in shared memory.
// ---------------------------------------------------------------------------------------------
constexpr int kQtThreads = 256;

__global__ void __launch_bounds__(kQtThreads, 6) quadtree_pool_fwd_kernel(const __nv_bfloat16* __restrict__ q,
                                                                       const __nv_bfloat16* __restrict__ l4,
                                                                       __nv_bfloat16* __restrict__ feat, int B, int QH, int QW,
                                                                       int Cq, int GHW, int Cg, int ldf) {
  extern __shared__ __align__(16) uint8_t qt_smem[];
  const int b = blockIdx.x, part = blockIdx.y;
  const int PH = QH / 2, PW = QW / 2;
  if (part < 4) {
    __nv_bfloat16* tile = reinterpret_cast<__nv_bfloat16*>(qt_smem);  // [QH*QW][Cq]
    const int n16 = QH * QW * Cq / 8;
    const uint4* src = reinterpret_cast<const uint4*>(q + (static_cast<long long>(part) * B + b) * QH * QW * Cq);
    // all of a thread's 16-byte loads are issued before the first one is consumed (four per batch: 784 chunks / 256 threads)
    for (int i0 = threadIdx.x; i0 < n16; i0 += 4 * kQtThreads) {
      uint4 v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (i0 + k * kQtThreads < n16) v[k] = ld_nc16(src + i0 + k * kQtThreads);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (i0 + k * kQtThreads < n16) reinterpret_cast<uint4*>(tile)[i0 + k * kQtThreads] = v[k];
    }
    __syncthreads();
    const int PP = PH * PW;
    const int nout8 = Cq * PP / 8;
    __nv_bfloat16* dst = feat + static_cast<long long>(b) * ldf + Cg + static_cast<long long>(part) * Cq * PP;
    for (int t = threadIdx.x; t < nout8; t += kQtThreads) {
      uint32_t pk[4];
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        float m2[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int o = 8 * t + j + e;
          const int c = o / PP, r = o - c * PP;
          const int ph = r / PW, pw = r - ph * PW;
          const __nv_bfloat16* win = tile + ((2 * ph) * QW + 2 * pw) * Cq + c;
          // max of bf16 values is exact in any order; same operands as MaxPool2d(2,2) floor mode
          m2[e] = fmaxf(fmaxf(__bfloat162float(win[0]), __bfloat162float(win[Cq])),
                        fmaxf(__bfloat162float(win[QW * Cq]), __bfloat162float(win[QW * Cq + Cq])));
        }
        pk[j >> 1] = pack_bf16x2(m2[0], m2[1]);
      }
      reinterpret_cast<uint4*>(dst)[t] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
  } else {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int phase = lane >> 3;
    const float inv = 1.0f / static_cast<float>(GHW);
    const __nv_bfloat16* src = l4 + static_cast<long long>(b) * GHW * Cg;
    for (int c0 = warp * 64; c0 < Cg; c0 += (kQtThreads / 32) * 64) {
      const int c = c0 + (lane & 7) * 8;
      float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      for (int p0 = phase; p0 < GHW; p0 += 16) {  // four pixels in flight per lane
        uint4 v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (p0 + 4 * k < GHW) v[k] = ld_nc16(src + static_cast<long long>(p0 + 4 * k) * Cg + c);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (p0 + 4 * k < GHW) {
            float f[8];
            unpack8(v[k], f);
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] += f[e];
          }
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 8);
        acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 16);
        acc[e] *= inv;
      }
      if (phase == 0) *reinterpret_cast<uint4*>(feat + static_cast<long long>(b) * ldf + c) = pack8(acc);
    }
  }
}

// Backward, scatter-free: the quadrant part re-derives the (first) arg-max of every 2x2 window from q, masks by the
// fused ReLU (max > 0) and routes the gradient of output c*(PH*PW)+ph*PW+pw to that pixel, writing all four pixels of
// the window (and the row / column the floor-mode pool dropped) as 16-byte chunks; the global part broadcasts dfeat/GHW.
__global__ void __launch_bounds__(kQtThreads, 8) quadtree_pool_bwd_kernel(const __nv_bfloat16* __restrict__ dfeat,
                                                                       const __nv_bfloat16* __restrict__ q,
                                                                       __nv_bfloat16* __restrict__ dq,
                                                                       __nv_bfloat16* __restrict__ dl4, int B, int QH, int QW,
                                                                       int Cq, int GHW, int Cg, int ldf) {
  extern __shared__ __align__(16) uint8_t qt_smem[];
  const int b = blockIdx.x, part = blockIdx.y;
  const int PH = QH / 2, PW = QW / 2, PP = PH * PW;
  if (part < 4) {
    __nv_bfloat16* gs = reinterpret_cast<__nv_bfloat16*>(qt_smem);  // this quadrant's slice of the feature-row gradient
    const uint4* gsrc = reinterpret_cast<const uint4*>(dfeat + static_cast<long long>(b) * ldf + Cg + static_cast<long long>(part) * Cq * PP);
    for (int i = threadIdx.x; i < Cq * PP / 8; i += kQtThreads) reinterpret_cast<uint4*>(gs)[i] = ld_nc16(gsrc + i);
    __syncthreads();
    const long long base = (static_cast<long long>(part) * B + b) * QH * QW * Cq;
    const int nch = Cq / 8;
    for (int i = threadIdx.x; i < PP * nch; i += kQtThreads) {
      const int ch = i % nch, win = i / nch;
      const int ph = win / PW, pw = win - ph * PW;
      const long long o00 = base + (static_cast<long long>(2 * ph) * QW + 2 * pw) * Cq + ch * 8;
      const long long offs[4] = {o00, o00 + Cq, o00 + static_cast<long long>(QW) * Cq, o00 + static_cast<long long>(QW) * Cq + Cq};
      float v[4][8];
#pragma unroll
      for (int k = 0; k < 4; ++k) unpack8(ld_nc16(q + offs[k]), v[k]);
      float out[4][8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        int am = 0;
        float m = v[0][e];
        if (v[1][e] > m) { m = v[1][e]; am = 1; }
        if (v[2][e] > m) { m = v[2][e]; am = 2; }
        if (v[3][e] > m) { m = v[3][e]; am = 3; }
        const float g = (m > 0.f) ? __bfloat162float(gs[(ch * 8 + e) * PP + win]) : 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) out[k][e] = (k == am) ? g : 0.f;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) *reinterpret_cast<uint4*>(dq + offs[k]) = pack8(out[k]);
    }
    // pixels outside every pooling window (odd QH / QW: last row / column) receive no gradient
    if ((QH & 1) || (QW & 1)) {
      const uint4 z = make_uint4(0, 0, 0, 0);
      for (int i = threadIdx.x; i < QH * QW * nch; i += kQtThreads) {
        const int ch = i % nch, pix = i / nch;
        const int h = pix / QW, w = pix - h * QW;
        if (h >= 2 * PH || w >= 2 * PW) *reinterpret_cast<uint4*>(dq + base + static_cast<long long>(pix) * Cq + ch * 8) = z;
      }
    }
  } else {
    const int nch = Cg / 8;
    const float inv_n = static_cast<float>(GHW);
    if (kQtThreads % nch == 0) {
      // the thread's channel chunk never changes: one load, one divide, GHW * nch / 256 stores
      const int ch = threadIdx.x % nch;
      float f[8];
      unpack8(ld_nc16(dfeat + static_cast<long long>(b) * ldf + ch * 8), f);
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = f[e] / inv_n;
      const uint4 v = pack8(f);
      for (int p = threadIdx.x / nch; p < GHW; p += kQtThreads / nch)
        *reinterpret_cast<uint4*>(dl4 + (static_cast<long long>(b) * GHW + p) * Cg + ch * 8) = v;
    } else {
      for (int i = threadIdx.x; i < GHW * nch; i += kQtThreads) {
        const int ch = i % nch, p = i / nch;
        float f[8];
        unpack8(ld_nc16(dfeat + static_cast<long long>(b) * ldf + ch * 8), f);
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = f[e] / inv_n;
        *reinterpret_cast<uint4*>(dl4 + (static_cast<long long>(b) * GHW + p) * Cg + ch * 8) = pack8(f);
      }
    }
  }
}

// Region average pooling for the level-2 models (AttentionHierarchicalCNN): x [R][P][C] bf16 (R regions of
// P pixels, already bias+ReLU) -> out[R][C] written at feat + r_off(r): caller passes a dense [R][C]
// destination with row stride ldo.
__global__ void region_avgpool_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                          long long R, int P, int C, long long ldo) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < R * C;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = i % C;
    const long long r = i / C;
    const __nv_bfloat16* src = x + r * P * C + c;
    float acc = 0.f;
    for (int p = 0; p < P; ++p) acc += __bfloat162float(src[static_cast<long long>(p) * C]);
    out[r * ldo + c] = __float2bfloat16_rn(acc / P);
  }
}
// Backward of (ReLU -> mean over P): dx[r][p][c] = dout[r][c]/P * (x[r][p][c] > 0).
__global__ void region_avgpool_bwd_kernel(const __nv_bfloat16* __restrict__ dout, const __nv_bfloat16* __restrict__ x,
                                          __nv_bfloat16* __restrict__ dx, long long R, int P, int C, long long ldo,
                                          int relu_mask) {
  const long long total = R * P * C;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = i % C;
    const long long r = i / (static_cast<long long>(C) * P);
    float g = __bfloat162float(dout[r * ldo + c]) / P;
    if (relu_mask && !(__bfloat162float(x[i]) > 0.f)) g = 0.f;
    dx[i] = __float2bfloat16_rn(g);
  }
}

// ---------------------------------------------------------------------------------------------
// MaxPool3d with kernel == stride == (kd, kh, kw), no padding, floor mode (3dcnn/models.py:111-135) on NDHWC
// bf16, int8 arg-max code = (dz*kh + dy)*kw + dx (first maximum in that order). C % 8 == 0.
// ---------------------------------------------------------------------------------------------
__global__ void maxpool3d_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                     signed char* __restrict__ argmax, int N, int D, int H, int W, int C, int kd, int kh,
                                     int kw) {
  const int Do = D / kd, Ho = H / kh, Wo = W / kw, groups = C / 8;
  const long long total = static_cast<long long>(N) * Do * Ho * Wo * groups;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    long long r = i;
    const int cg = r % groups; r /= groups;
    const int wo = r % Wo; r /= Wo;
    const int ho = r % Ho; r /= Ho;
    const int dd = r % Do;
    const int n = r / Do;
    float best[8];
    int bi[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { best[e] = -INFINITY; bi[e] = 0; }
    int code = 0;
    for (int a = 0; a < kd; ++a)
      for (int b = 0; b < kh; ++b)
        for (int c = 0; c < kw; ++c, ++code) {
          float f[8];
          unpack8(ld_nc16(x + ((((static_cast<long long>(n) * D + dd * kd + a) * H + ho * kh + b) * W + wo * kw + c) * C + cg * 8)), f);
#pragma unroll
          for (int e = 0; e < 8; ++e)
            if (f[e] > best[e] || code == 0) { best[e] = f[e]; bi[e] = code; }
        }
    *reinterpret_cast<uint4*>(out + i * 8) = pack8(best);
    if (argmax) {
      uint2 pk;
      pk.x = (bi[0] & 0xff) | ((bi[1] & 0xff) << 8) | ((bi[2] & 0xff) << 16) | ((bi[3] & 0xff) << 24);
      pk.y = (bi[4] & 0xff) | ((bi[5] & 0xff) << 8) | ((bi[6] & 0xff) << 16) | ((bi[7] & 0xff) << 24);
      *reinterpret_cast<uint2*>(argmax + i * 8) = pk;
    }
  }
}
__global__ void maxpool3d_bwd_kernel(const __nv_bfloat16* __restrict__ dout, const signed char* __restrict__ argmax,
                                     __nv_bfloat16* __restrict__ dx, int N, int D, int H, int W, int C, int kd, int kh,
                                     int kw) {
  const int Do = D / kd, Ho = H / kh, Wo = W / kw, groups = C / 8;
  const long long total = static_cast<long long>(N) * D * H * W * groups;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    long long r = i;
    const int cg = r % groups; r /= groups;
    const int w = r % W; r /= W;
    const int h = r % H; r /= H;
    const int d = r % D;
    const int n = r / D;
    float acc[8] = {0};
    const int dd = d / kd, ho = h / kh, wo = w / kw;
    if (dd < Do && ho < Ho && wo < Wo) {
      const int code = ((d - dd * kd) * kh + (h - ho * kh)) * kw + (w - wo * kw);
      const long long o = ((((static_cast<long long>(n) * Do + dd) * Ho + ho) * Wo + wo) * groups + cg) * 8;
      const uint2 pk = *reinterpret_cast<const uint2*>(argmax + o);
      float g[8];
      unpack8(ld_nc16(dout + o), g);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int a = (e < 4 ? (pk.x >> (8 * e)) : (pk.y >> (8 * (e - 4)))) & 0xff;
        if (a == code) acc[e] = g[e];
      }
    }
    *reinterpret_cast<uint4*>(dx + i * 8) = pack8(acc);
  }
}

// ---------------------------------------------------------------------------------------------
// Conv3d block tail fused (3dcnn/models.py:108-135): BatchNorm3d-apply + ReLU + MaxPool3d((KD,2,2), stride = kernel) in one
// pass over the raw conv output. The full-resolution activation is never written: the pass stores the pooled activation, the
// int8 arg-max code ((dz*2 + dy)*2 + dx, first maximum in that order, on the bf16-rounded activation: bit-identical to
// bn_apply + maxpool3d_fwd) and `yarg`, the raw conv output at the arg-max. With yarg the backward statistics
// (sum dz, sum dz*xhat) come from pooled-size tensors only (bn_bwd_reduce_kernel over the pooled rows, mask recomputed from
// yarg), and the apply pass reads y once and writes dy once — the full-size `a` and `da` round trips are gone.
// D % KD == 0, H and W even, C % 8 == 0, vector count < 2^31.
// ---------------------------------------------------------------------------------------------
template <int KD>
__global__ void __launch_bounds__(256) bn_relu_maxpool3d_fwd_kernel(
    const __nv_bfloat16* __restrict__ y, const float* __restrict__ scale, const float* __restrict__ shift,
    __nv_bfloat16* __restrict__ out, __nv_bfloat16* __restrict__ yarg, signed char* __restrict__ argmax, int N, int D, int H,
    int W, int C) {
  constexpr int kTaps = KD * 4;
  const unsigned groups = C / 8, Wo = W / 2, Ho = H / 2, Do = D / KD;
  const unsigned total = static_cast<unsigned>(N) * Do * Ho * Wo * groups;
  const long long row = static_cast<long long>(W) * C, plane = row * H;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    unsigned r = i;
    const unsigned cg = r % groups; r /= groups;
    const unsigned wo = r % Wo; r /= Wo;
    const unsigned ho = r % Ho; r /= Ho;  // r = n * Do + do
    const __nv_bfloat16* base = y + (static_cast<long long>(r) * KD * H + ho * 2) * row + static_cast<long long>(wo) * 2 * C + cg * 8;
    uint4 v[kTaps];
#pragma unroll
    for (int t = 0; t < kTaps; ++t) v[t] = ld_nc16(base + (t >> 2) * plane + ((t >> 1) & 1) * row + (t & 1) * C);
    float sc[8], sh[8];
    *reinterpret_cast<float4*>(sc) = __ldg(reinterpret_cast<const float4*>(scale + cg * 8));
    *reinterpret_cast<float4*>(sc + 4) = __ldg(reinterpret_cast<const float4*>(scale + cg * 8 + 4));
    *reinterpret_cast<float4*>(sh) = __ldg(reinterpret_cast<const float4*>(shift + cg * 8));
    *reinterpret_cast<float4*>(sh + 4) = __ldg(reinterpret_cast<const float4*>(shift + cg * 8 + 4));
    unsigned best2[4] = {0xff80ff80u, 0xff80ff80u, 0xff80ff80u, 0xff80ff80u}, bi2[4] = {0u, 0u, 0u, 0u}, ya2[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int t = 0; t < kTaps; ++t) {
      const unsigned raw[4] = {v[t].x, v[t].y, v[t].z, v[t].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float lo = __uint_as_float(raw[j] << 16), hi = __uint_as_float(raw[j] & 0xffff0000u);
        const unsigned a2 = pack_bf16x2(fmaxf(fmaf(lo, sc[2 * j], sh[2 * j]), 0.f),
                                        fmaxf(fmaf(hi, sc[2 * j + 1], sh[2 * j + 1]), 0.f));
        const unsigned m = __hgt2_mask(*reinterpret_cast<const __nv_bfloat162*>(&a2),
                                       *reinterpret_cast<const __nv_bfloat162*>(&best2[j]));
        best2[j] = (a2 & m) | (best2[j] & ~m);
        ya2[j] = (raw[j] & m) | (ya2[j] & ~m);
        bi2[j] = ((t * 0x00010001u) & m) | (bi2[j] & ~m);
      }
    }
    const long long o = static_cast<long long>(i) * 8;
    *reinterpret_cast<uint4*>(out + o) = make_uint4(best2[0], best2[1], best2[2], best2[3]);
    if (yarg) {
      *reinterpret_cast<uint4*>(yarg + o) = make_uint4(ya2[0], ya2[1], ya2[2], ya2[3]);
      uint2 pk;
      pk.x = __byte_perm(bi2[0], bi2[1], 0x6420);
      pk.y = __byte_perm(bi2[2], bi2[3], 0x6420);
      *reinterpret_cast<uint2*>(argmax + o) = pk;
    }
  }
}
// Backward apply: dy = A*dz + B*y + K at every full-resolution position (coef = [A | B | K] from bn_bwd_finalize_rows_kernel),
// dz = dpool routed to the arg-max position and masked by ReLU (recomputed from y with the forward's own expression). One
// thread per pooled vector: it reads its dpool / arg-max once and streams its KD*4 window positions of y -> dy.
template <int KD>
__global__ void __launch_bounds__(256) bn_pool3d_bwd_apply_kernel(
    const __nv_bfloat16* __restrict__ dpool, const signed char* __restrict__ argmax, const __nv_bfloat16* __restrict__ y,
    const float* __restrict__ coef, const float* __restrict__ msc, const float* __restrict__ msh, __nv_bfloat16* __restrict__ dy,
    int N, int D, int H, int W, int C) {
  constexpr int kTaps = KD * 4;
  const unsigned groups = C / 8, Wo = W / 2, Ho = H / 2, Do = D / KD;
  const unsigned total = static_cast<unsigned>(N) * Do * Ho * Wo * groups;
  const long long row = static_cast<long long>(W) * C, plane = row * H;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    unsigned r = i;
    const unsigned cg = r % groups; r /= groups;
    const unsigned wo = r % Wo; r /= Wo;
    const unsigned ho = r % Ho; r /= Ho;
    const long long off = (static_cast<long long>(r) * KD * H + ho * 2) * row + static_cast<long long>(wo) * 2 * C + cg * 8;
    uint4 v[kTaps];
#pragma unroll
    for (int t = 0; t < kTaps; ++t) v[t] = ld_nc16(y + off + (t >> 2) * plane + ((t >> 1) & 1) * row + (t & 1) * C);
    const uint4 qg = ld_nc16(dpool + static_cast<long long>(i) * 8);
    const uint2 pk = __ldg(reinterpret_cast<const uint2*>(argmax + static_cast<long long>(i) * 8));
    float ca[8], cb[8], ck[8], ms[8], mh[8], g[8];
#pragma unroll
    for (int e = 0; e < 8; e += 4) {
      *reinterpret_cast<float4*>(ca + e) = __ldg(reinterpret_cast<const float4*>(coef + cg * 8 + e));
      *reinterpret_cast<float4*>(cb + e) = __ldg(reinterpret_cast<const float4*>(coef + C + cg * 8 + e));
      *reinterpret_cast<float4*>(ck + e) = __ldg(reinterpret_cast<const float4*>(coef + 2 * C + cg * 8 + e));
      *reinterpret_cast<float4*>(ms + e) = __ldg(reinterpret_cast<const float4*>(msc + cg * 8 + e));
      *reinterpret_cast<float4*>(mh + e) = __ldg(reinterpret_cast<const float4*>(msh + cg * 8 + e));
    }
    unpack8(qg, g);
    int code[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) code[e] = ((e < 4 ? (pk.x >> (8 * e)) : (pk.y >> (8 * (e - 4)))) & 0xff);
#pragma unroll
    for (int t = 0; t < kTaps; ++t) {
      float f[8], o[8];
      unpack8(v[t], f);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const bool on = (code[e] == t) && (fmaf(f[e], ms[e], mh[e]) > 0.f);
        o[e] = fmaf(ca[e], on ? g[e] : 0.f, fmaf(cb[e], f[e], ck[e]));
      }
      *reinterpret_cast<uint4*>(dy + off + (t >> 2) * plane + ((t >> 1) & 1) * row + (t & 1) * C) = pack8(o);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Attention pooling over R region vectors (AttentionHierarchicalCNN, QS/models.py:86-90):
//   w = softmax(scores[b, :]); out[b, :] = sum_i w_i * x[b, i, :].  x [B][R][C] fp32, scores [B][R] fp32.
// One block per sample, C threads.
// ---------------------------------------------------------------------------------------------
__global__ void attn_pool_fwd_kernel(const float* __restrict__ x, const float* __restrict__ scores, float* __restrict__ wts,
                                     float* __restrict__ out, int R, int C) {
  const int b = blockIdx.x;
  float m = -INFINITY;
  for (int i = 0; i < R; ++i) m = fmaxf(m, scores[b * R + i]);
  float den = 0.f;
  for (int i = 0; i < R; ++i) den += expf(scores[b * R + i] - m);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc = 0.f;
    for (int i = 0; i < R; ++i) acc += (expf(scores[b * R + i] - m) / den) * x[(static_cast<long long>(b) * R + i) * C + c];
    out[static_cast<long long>(b) * C + c] = acc;
  }
  for (int i = threadIdx.x; i < R; i += blockDim.x) wts[b * R + i] = expf(scores[b * R + i] - m) / den;
}
// dx[b,i,c] = w_i*dout[c];  dscore_i = w_i * (t_i - sum_j w_j t_j), t_i = x_i . dout.  Block per sample, 32*k threads.
__global__ void attn_pool_bwd_kernel(const float* __restrict__ x, const float* __restrict__ wts, const float* __restrict__ dout,
                                     float* __restrict__ dx, float* __restrict__ dscores, int R, int C) {
  extern __shared__ float t[];  // [R]
  const int b = blockIdx.x;
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int i = wrp; i < R; i += nw) {
    float acc = 0.f;
    for (int c = lane; c < C; c += 32) acc += x[(static_cast<long long>(b) * R + i) * C + c] * dout[static_cast<long long>(b) * C + c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) t[i] = acc;
  }
  __syncthreads();
  float mean = 0.f;
  for (int j = 0; j < R; ++j) mean += wts[b * R + j] * t[j];
  for (int i = threadIdx.x; i < R; i += blockDim.x) dscores[b * R + i] = wts[b * R + i] * (t[i] - mean);
  for (int k = threadIdx.x; k < R * C; k += blockDim.x) {
    const int i = k / C, c = k - i * C;
    dx[static_cast<long long>(b) * R * C + k] = wts[b * R + i] * dout[static_cast<long long>(b) * C + c];
  }
}

// ---------------------------------------------------------------------------------------------
// Small fp32 linear layers of the fusion head (numerical MLP 47->94->256, classifier.3 2688->nc).
// One warp per output element; inputs fp32 or bf16, fp32 weights and accumulation.
// ---------------------------------------------------------------------------------------------
template <typename TIn>
__device__ __forceinline__ float ld_as_float(const TIn* p);
template <>
__device__ __forceinline__ float ld_as_float<float>(const float* p) { return *p; }
template <>
__device__ __forceinline__ float ld_as_float<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

// out[b][n] = drop(act(x[b][:] . w[n][:] + bias[n])); out fp32 (ldo) and/or bf16 (ldo16).
template <typename TIn>
__global__ void small_linear_fwd_kernel(const TIn* __restrict__ x, long long ldx, const float* __restrict__ w,
                                        const float* __restrict__ bias, int B, int N, int K, int relu, float drop_p,
                                        unsigned long long seed, float* __restrict__ out, long long ldo,
                                        __nv_bfloat16* __restrict__ out16, long long ldo16) {
  const long long warp = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= static_cast<long long>(B) * N) return;
  const int n = warp % N;
  const int b = warp / N;
  float acc = 0.f;
  for (int k = lane; k < K; k += 32) acc = fmaf(ld_as_float<TIn>(x + b * ldx + k), w[static_cast<long long>(n) * K + k], acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) {
    if (bias) acc += bias[n];
    if (relu) acc = fmaxf(acc, 0.f);
    acc *= dropout_scale(seed, static_cast<uint32_t>(b) * N + n, drop_p);
    if (out) out[b * ldo + n] = acc;
    if (out16) out16[b * ldo16 + n] = __float2bfloat16_rn(acc);
  }
}
// dx[b][k] = mask(b,k) * sum_n dy[b][n] * w[n][k]; mask = (act[b][k] > 0) (post-ReLU/dropout output, optional).
template <typename TDy>
__global__ void small_linear_bwd_dx_kernel(const TDy* __restrict__ dy, long long ldy, const float* __restrict__ w,
                                           int B, int N, int K, const float* __restrict__ act, long long lda,
                                           float drop_p, unsigned long long seed, float* __restrict__ dx,
                                           long long ldx, __nv_bfloat16* __restrict__ dx16, long long ldx16) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(B) * K) return;
  const int k = i % K;
  const int b = i / K;
  float acc = 0.f;
  for (int n = 0; n < N; ++n) acc = fmaf(ld_as_float<TDy>(dy + b * ldy + n), w[static_cast<long long>(n) * K + k], acc);
  if (act) {
    // act holds the layer's stored output relu(z)*dropout_scale: zero wherever either gate was closed.
    const float a = act[b * lda + k];
    acc = (a > 0.f) ? acc * dropout_scale(seed, static_cast<uint32_t>(b) * K + k, drop_p) : 0.f;
  }
  if (dx) dx[b * ldx + k] = acc;
  if (dx16) dx16[b * ldx16 + k] = __float2bfloat16_rn(acc);
}
// dw[n][k] (+)= sum_b dy[b][n] * x[b][k]; db[n] (+)= sum_b dy[b][n]. Block (32 k, 8 batch lanes) per n:
// loads of x are coalesced along k, the batch is split over 8 lanes and combined through shared memory.
template <typename TDy, typename TX>
__global__ void small_linear_bwd_dw_kernel(const TDy* __restrict__ dy, long long ldy, const TX* __restrict__ x,
                                           long long ldx, int B, int N, int K, float* __restrict__ dw,
                                           float* __restrict__ db, int accumulate) {
  __shared__ float sh[8][33];
  __shared__ float shb[8];
  const int n = blockIdx.y;
  const int k = blockIdx.x * 32 + threadIdx.x;
  float acc = 0.f, accb = 0.f;
  for (int b = threadIdx.y; b < B; b += 8) {
    const float d = ld_as_float<TDy>(dy + b * ldy + n);
    if (k < K) acc = fmaf(d, ld_as_float<TX>(x + b * ldx + k), acc);
    accb += d;
  }
  sh[threadIdx.y][threadIdx.x] = acc;
  if (threadIdx.x == 0) shb[threadIdx.y] = accb;
  __syncthreads();
  if (threadIdx.y == 0) {
    float tot = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) tot += sh[j][threadIdx.x];
    if (k < K) {
      const long long i = static_cast<long long>(n) * K + k;
      dw[i] = accumulate ? dw[i] + tot : tot;
    }
    if (db && blockIdx.x == 0 && threadIdx.x == 0) {
      float tb = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) tb += shb[j];
      db[n] = accumulate ? db[n] + tb : tb;
    }
  }
}

// Elementwise helpers on dense tensors.
__global__ void f32_to_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, long long n) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    out[i] = __float2bfloat16_rn(x[i]);
}
__global__ void add_bf16_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                                __nv_bfloat16* __restrict__ out, long long total8) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total8;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float x[8], y[8];
    unpack8(ld_nc16(a + i * 8), x);
    unpack8(ld_nc16(b + i * 8), y);
#pragma unroll
    for (int e = 0; e < 8; ++e) x[e] += y[e];
    *reinterpret_cast<uint4*>(out + i * 8) = pack8(x);
  }
}
// Post-GEMM head epilogue on an fp32 [B][N] tensor: h = relu(h) * dropout -> fp32 in place + bf16 copy.
__global__ void relu_dropout_kernel(float* __restrict__ h, __nv_bfloat16* __restrict__ h16, long long total, float drop_p,
                                    unsigned long long seed, int relu) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float v = h[i];
    if (relu) v = fmaxf(v, 0.f);
    v *= dropout_scale(seed, static_cast<uint32_t>(i), drop_p);
    h[i] = v;
    if (h16) h16[i] = __float2bfloat16_rn(v);
  }
}

// dz = dout * 1[act > 0] * dropout_scale (act = stored relu/dropout output): backward of relu_dropout_kernel.
__global__ void relu_dropout_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ act, float* __restrict__ dz,
                                        __nv_bfloat16* __restrict__ dz16, long long total, float drop_p,
                                        unsigned long long seed, int relu) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float v = dout[i];
    if (relu && !(act[i] > 0.f)) v = 0.f;
    else v *= dropout_scale(seed, static_cast<uint32_t>(i), drop_p);
    if (dz) dz[i] = v;
    if (dz16) dz16[i] = __float2bfloat16_rn(v);
  }
}

}  // namespace qt
