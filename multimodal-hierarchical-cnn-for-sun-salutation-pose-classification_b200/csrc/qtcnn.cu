// libqtcnn.so — C ABI (include/qtcnn.h) over the sm_100a kernels in igemm.cuh / elementwise.cuh.
// Single translation unit; build: see __graft_entry__.build().
#include "../../include/qtcnn.h"

#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "elementwise.cuh"
#include "igemm.cuh"
#include "conv3x3.cuh"
#include "stem.cuh"
#include "wgrad3x3.cuh"
#include "conv3d_c8.cuh"
#include "head.cuh"
#include "optim.cuh"
#include "lstm.cuh"
#include "sgemm.cuh"

using namespace qt;

namespace {

thread_local char g_err[512] = "";

int fail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return -1;
}
int cuda_status(const char* what) {
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
    return static_cast<int>(e);
  }
  return 0;
}
inline cudaStream_t S(qt_stream_t s) { return static_cast<cudaStream_t>(s); }
inline int grid_for(long long total, int block, int cap = 148 * 16) {
  long long g = (total + block - 1) / block;
  if (g < 1) g = 1;
  if (g > cap) g = cap;
  return static_cast<int>(g);
}
inline int ilog2_exact(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return ((1 << l) == v) ? l : -1;
}
inline int out_dim(int in, int k, int s, int p) { return (in + 2 * p - k) / s + 1; }

constexpr int kNumSMs = 148;

// ------------------------------------------------------------------------------------------------
// GEMM launch helpers
// ------------------------------------------------------------------------------------------------
int g_tune[16] = {0};  // [0] wgrad pipeline variant, [1] kmajor pipeline variant
bool make_weight_tmap(CUtensorMap* map, const void* base, long long rows, long long cols, int box_rows);

template <int BN, int STAGES, int LAG, int NPW>
int launch_kmajor(IgemmParams& p, dim3 grid, cudaStream_t st) {
  using L = KMajorSmem<BN, STAGES>;
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(igemm_kmajor_kernel<BN, STAGES, LAG, NPW>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal);
    configured = true;
  }
  // Weight tiles by TMA when a 64-wide k-block maps onto 64 contiguous columns of the [nout][wtaps*cin] matrix.
  alignas(64) CUtensorMap bmap;
  memset(&bmap, 0, sizeof(bmap));
  p.b_tma = 0;
  const long long cols = static_cast<long long>(p.wtaps) * p.cin;
  bool natural = true;
  for (int t = 0; t < p.ntaps; ++t) natural = natural && (p.wtap[t] == t);
  int mode = 0;
  if (p.ntaps == 1 || p.cin % 64 == 0) mode = 1;
  else if (natural) mode = 2;
  if (g_tune[6] == 0 && mode && p.groups >= 1 && p.b_goff[0] == 0 && p.b_goff[1] == 0 && p.b_goff[2] == 0 && p.b_goff[3] == 0 &&
      (reinterpret_cast<uintptr_t>(p.b) & 15) == 0 && (cols % 8) == 0 && make_weight_tmap(&bmap, p.b, p.nout, cols, BN))
    p.b_tma = mode;
  igemm_kmajor_kernel<BN, STAGES, LAG, NPW><<<grid, NPW == 8 ? kGemmThreadsWide : kGemmThreads, L::kTotal, st>>>(p, bmap);
  return cuda_status("igemm_kmajor_kernel");
}
template <int BN, int STAGES, int LAG, int NPW>
int launch_wgrad(const IgemmParams& p, dim3 grid, cudaStream_t st) {
  using L = WgradSmem<BN, STAGES>;
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(igemm_wgrad_kernel<BN, STAGES, LAG, NPW>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal);
    configured = true;
  }
  igemm_wgrad_kernel<BN, STAGES, LAG, NPW><<<grid, NPW == 8 ? kGemmThreadsWide : kGemmThreads, L::kTotal, st>>>(p);
  return cuda_status("igemm_wgrad_kernel");
}

inline int pick_bn(int nout) { return nout <= 64 ? 64 : 128; }

// Chooses split-K and launches the K-major kernel (+ the split-K reduction when used).
// `dense_out_ld` > 0 means the output is a dense row-major [M][nout] matrix with that row stride
// (required for split-K).
int run_kmajor(IgemmParams p, cudaStream_t st, void* ws, size_t ws_bytes, long long dense_out_ld) {  // p by value: launch_kmajor edits it
  const int BN = pick_bn(p.nout);
  const int gx = (p.M + kBM - 1) / kBM, gy = (p.nout + BN - 1) / BN;
  p.num_kb = (p.ntaps * p.cin + kBK - 1) / kBK;
  if (p.num_kb < 1) return fail("igemm: empty K");
  int ksplit = 1;
  const long long tiles = static_cast<long long>(gx) * gy * p.groups;
  if (dense_out_ld > 0 && p.groups == 1 && !(p.flags & (EPI_STATS | EPI_ADDEND)) && tiles < kNumSMs && p.num_kb >= 8) {
    ksplit = static_cast<int>((2 * kNumSMs + tiles - 1) / tiles);
    if (ksplit > p.num_kb / 4) ksplit = p.num_kb / 4;
    if (ksplit > 32) ksplit = 32;
    if (ksplit < 1) ksplit = 1;
  }
  p.kb_per_split = (p.num_kb + ksplit - 1) / ksplit;
  ksplit = (p.num_kb + p.kb_per_split - 1) / p.kb_per_split;  // no empty splits
  p.ksplit = ksplit;
  p.Mpad = gx * kBM;
  p.Npad = gy * BN;
  const int user_flags = p.flags;
  if (ksplit > 1) {
    const size_t need = static_cast<size_t>(ksplit) * p.Mpad * p.Npad * sizeof(float);
    if (ws == nullptr || ws_bytes < need) return fail("igemm: split-K workspace too small (%zu < %zu)", ws_bytes, need);
    p.splitk_ws = static_cast<float*>(ws);
    p.flags = EPI_SPLITK;
  }
  dim3 grid(gx, gy, p.groups * ksplit);
  int rc;
  // default: 8 producer warps, two CTAs per SM (measured best on B200, profiles/r01_conv_tuning.md)
  if (g_tune[1] == 1) rc = (BN == 64) ? launch_kmajor<64, 8, 3, 8>(p, grid, st) : launch_kmajor<128, 6, 3, 8>(p, grid, st);
  else if (g_tune[1] == 2) rc = (BN == 64) ? launch_kmajor<64, 4, 1, 4>(p, grid, st) : launch_kmajor<128, 3, 1, 4>(p, grid, st);
  else rc = (BN == 64) ? launch_kmajor<64, 4, 1, 8>(p, grid, st) : launch_kmajor<128, 3, 1, 8>(p, grid, st);
  if (rc) return rc;
  if (ksplit > 1) {
    const long long total = static_cast<long long>(p.M) * p.nout;
    splitk_reduce_rows_kernel<<<grid_for(total, 256, 1 << 20), 256, 0, st>>>(
        p.splitk_ws, ksplit, p.M, p.nout, p.Mpad, p.Npad, (user_flags & EPI_BIAS) ? p.bias : nullptr,
        (user_flags & EPI_RELU) ? 1 : 0, (user_flags & EPI_OUT_F32) ? 1 : 0, p.out, dense_out_ld);
    rc = cuda_status("splitk_reduce_rows_kernel");
  }
  return rc;
}

void fill_groups(IgemmParams& p, const qt_conv_desc* d, bool a_is_x) {
  p.groups = d->groups < 1 ? 1 : d->groups;
  for (int g = 0; g < 4; ++g) {
    p.a_goff[g] = a_is_x ? d->x_group_off[g] : d->y_group_off[g];
    p.o_goff[g] = a_is_x ? d->y_group_off[g] : d->x_group_off[g];
    p.b_goff[g] = 0;
  }
}

int check_desc(const qt_conv_desc* d) {
  if (!d) return fail("null conv descriptor");
  if (d->n < 1 || d->in_c < 1 || d->out_c < 1) return fail("conv: bad sizes");
  if (d->groups < 1 || d->groups > 4) return fail("conv: groups must be 1..4");
  if (d->k_d * d->k_h * d->k_w > kMaxTaps) return fail("conv: more than %d taps", kMaxTaps);
  if (d->in_c % 8 || d->out_c % 8) return fail("conv: channel counts must be multiples of 8 (got %d -> %d)", d->in_c, d->out_c);
  return 0;
}

struct OutDims { int d, h, w; };
OutDims conv_out_dims(const qt_conv_desc* d) {
  return {out_dim(d->in_d, d->k_d, d->stride_d, d->pad_d), out_dim(d->in_h, d->k_h, d->stride_h, d->pad_h),
          out_dim(d->in_w, d->k_w, d->stride_w, d->pad_w)};
}

size_t wgrad_ws_bytes(int F, int nout, int groups, long long pixels, int* ksplit_out) {
  const int BN = pick_bn(nout);
  const int gx = (F + kBM - 1) / kBM, gy = (nout + BN - 1) / BN;
  const long long num_kb = (pixels + 63) / 64;
  const long long tiles = static_cast<long long>(gx) * gy * groups;
  long long ksplit = (2 * kNumSMs + tiles - 1) / tiles;
  if (ksplit > num_kb / 8) ksplit = num_kb / 8;
  if (ksplit < 1) ksplit = 1;
  if (ksplit > 256) ksplit = 256;
  long long per = (num_kb + ksplit - 1) / ksplit;
  ksplit = (num_kb + per - 1) / per;
  if (ksplit_out) *ksplit_out = static_cast<int>(ksplit);
  return static_cast<size_t>(groups) * ksplit * gx * kBM * gy * BN * sizeof(float);
}

// Shared wgrad driver. p must describe the forward gather (A = x) and the dy view (ov / b).
int run_wgrad(IgemmParams p, cudaStream_t st, void* ws, size_t ws_bytes, float* dw, int accumulate, int cin_real,
              int taps_real) {
  const int F = p.ntaps * p.cin;
  const int BN = pick_bn(p.nout);
  const int gx = (F + kBM - 1) / kBM, gy = (p.nout + BN - 1) / BN;
  int ksplit = 1;
  const size_t need = wgrad_ws_bytes(F, p.nout, p.groups, p.M, &ksplit);
  if (ws == nullptr || ws_bytes < need) return fail("wgrad: workspace too small (%zu < %zu)", ws_bytes, need);
  p.num_kb = (p.M + 63) / 64;
  p.kb_per_split = (p.num_kb + ksplit - 1) / ksplit;
  p.ksplit = ksplit;
  p.Mpad = gx * kBM;
  p.Npad = gy * BN;
  p.splitk_ws = static_cast<float*>(ws);
  {  // 64 pixels decomposed over the logical output grid
    int r = 64;
    p.adv_w = r % p.ow; r /= p.ow;
    p.adv_h = r % p.oh; r /= p.oh;
    p.adv_d = r % p.od; r /= p.od;
    p.adv_n = r;
  }
  dim3 grid(gx, gy, p.groups * ksplit);
  // one split, one tap, one group (linear layers with enough tiles): the epilogue writes dw itself
  const bool direct = ksplit == 1 && p.groups == 1 && taps_real == 1 && cin_real == F && g_tune[8] == 0;
  if (direct) {
    p.out = dw;
    p.flags = EPI_WGRAD_DIRECT | (accumulate ? EPI_ADDEND : 0);
  }
  int rc;
  if (g_tune[0] == 1) rc = (BN == 64) ? launch_wgrad<64, 8, 3, 8>(p, grid, st) : launch_wgrad<128, 6, 3, 8>(p, grid, st);
  else if (g_tune[0] == 2) rc = (BN == 64) ? launch_wgrad<64, 4, 1, 4>(p, grid, st) : launch_wgrad<128, 3, 1, 4>(p, grid, st);
  else rc = (BN == 64) ? launch_wgrad<64, 4, 1, 8>(p, grid, st) : launch_wgrad<128, 3, 1, 8>(p, grid, st);
  if (rc || direct) return rc;
  launch_splitk_reduce_wgrad(p.splitk_ws, p.groups * ksplit, F, p.nout, p.Mpad, p.Npad, cin_real, taps_real, dw, accumulate, st);
  return cuda_status("splitk_reduce_wgrad_kernel");
}

// Forward-gather description shared by fprop and wgrad.
int fill_forward(IgemmParams& p, const qt_conv_desc* d) {
  const OutDims o = conv_out_dims(d);
  if (o.d < 1 || o.h < 1 || o.w < 1) return fail("conv: empty output");
  memset(&p, 0, sizeof(p));
  p.av = {d->x_stride[0], d->x_stride[1], d->x_stride[2], d->x_stride[3]};
  p.ov = {d->y_stride[0], d->y_stride[1], d->y_stride[2], d->y_stride[3]};
  p.id = d->in_d; p.ih = d->in_h; p.iw = d->in_w;
  p.nb = d->n; p.od = o.d; p.oh = o.h; p.ow = o.w;
  p.mult_d = d->stride_d; p.mult_h = d->stride_h; p.mult_w = d->stride_w;
  p.ntaps = d->k_d * d->k_h * d->k_w;
  p.wtaps = p.ntaps;
  p.cin = d->in_c;
  p.cin_log2 = ilog2_exact(d->in_c);
  if (p.ntaps > 1 && p.cin_log2 < 0) return fail("conv: in_c must be a power of two for multi-tap kernels (got %d)", d->in_c);
  int t = 0;
  for (int kd = 0; kd < d->k_d; ++kd)
    for (int kh = 0; kh < d->k_h; ++kh)
      for (int kw = 0; kw < d->k_w; ++kw, ++t) {
        p.off_d[t] = static_cast<signed char>(kd - d->pad_d);
        p.off_h[t] = static_cast<signed char>(kh - d->pad_h);
        p.off_w[t] = static_cast<signed char>(kw - d->pad_w);
        p.wtap[t] = static_cast<short>(t);
      }
  p.nout = d->out_c;
  p.M = d->n * o.d * o.h * o.w;
  fill_groups(p, d, true);
  return 0;
}


// ------------------------------------------------------------------------------------------------
// Persistent 3x3/s1/p1 kernel dispatch (conv3x3.cuh)
// ------------------------------------------------------------------------------------------------
bool g_use_conv3x3 = true;

bool dense_nhwc(const long long* st, int d, int h, int w, int c) {
  if (d > 1 && st[1] != static_cast<long long>(h) * w * c) return false;
  return st[3] == c && st[2] == static_cast<long long>(w) * c && st[0] == static_cast<long long>(d) * h * w * c;
}
// 3x3 (k_d == 1, one plane) or 3x3x3 (Conv3d) stride-1 pad-1 geometry of the slab kernels
bool slab_geometry(const qt_conv_desc* d) {
  if (d->k_h != 3 || d->k_w != 3 || d->stride_h != 1 || d->stride_w != 1 || d->pad_h != 1 || d->pad_w != 1) return false;
  if (d->k_d == 1) return d->in_d == 1 && d->pad_d == 0 && d->stride_d == 1;
  return d->k_d == 3 && d->pad_d == 1 && d->stride_d == 1 && d->in_d >= 1;
}

struct C3Plan {
  bool ok;
  int pair;  // cin == 32 Conv3d forward: two depth planes per 128-byte slab row (weights from qt_wpack_conv3d_pair)
  int bn, mt, nslab, nb, R, resident, staged;
  int num_m_tiles, num_n_tiles, V;
  size_t smem;
};

// cin/nout are the GEMM-side channel counts (swapped for dgrad).
C3Plan plan_conv3x3(const qt_conv_desc* d, int cin, int nout, int flags, bool dgrad = false) {
  C3Plan pl{};
  pl.ok = false;
  if (!g_use_conv3x3) return pl;
  if (!slab_geometry(d)) return pl;
  if (d->groups != 1) return pl;
  const bool is3d = d->k_d == 3;
  if (is3d && g_tune[7] == 1) return pl;  // knob 7: Conv3d through the generic gather kernel
  if (d->in_w < 10 && g_tune[4] == 1) return pl;  // knob 4: send small maps (7x7) to the gather kernel instead
  // the producers advance 16 virtual pixels with ONE row carry and ONE image carry: needs 16/(W+2) < H+1 (fails for 2x2 maps)
  if (16 / (d->in_w + 2) >= d->in_h + 1) return pl;
  pl.pair = (is3d && cin == 32 && !dgrad) ? 1 : 0;  // forward only: the pair-packed operand exists for the forward weights
  if ((cin % 64 && !pl.pair) || nout % 32) return pl;
  if (flags & (EPI_RELU | EPI_OUT_F32)) return pl;
  if ((flags & EPI_BIAS) && !is3d) return pl;
  if (!dense_nhwc(d->x_stride, d->in_d, d->in_h, d->in_w, d->in_c) || !dense_nhwc(d->y_stride, d->in_d, d->in_h, d->in_w, d->out_c)) return pl;
  const long long V = static_cast<long long>(d->n) * d->in_d * (d->in_h + 1) * (d->in_w + 2);
  if (V > (1ll << 30)) return pl;
  pl.V = static_cast<int>(V);
  pl.bn = nout <= 64 ? 64 : 128;
  pl.mt = 2;  // two 128-pixel sub-tiles per tile, one MMA-issuer warp each (single-sub-tile variants measured 5-40 % slower)
  const int bm = kBM * pl.mt;
  pl.R = ((bm + 2 * (d->in_w + 3)) + 7) / 8 * 8;
  const int slab_bytes = (pl.R * 128 + 1023) / 1024 * 1024;
  const int slabs = pl.pair ? 2 : d->k_d * (cin / 64);
  const int btile = pl.bn * 128;
  // staged (coalesced) write-out pays off on long image rows; on 14x14 / 7x7 maps the extra epilogue work costs more
  // than the scattered 16-byte stores (profiles/r01_conv_tuning.md). Knob 5: 1 never, 2 always.
  pl.staged = (g_tune[5] == 1) ? 0 : (g_tune[5] == 2 ? 1 : (d->in_w >= 20 ? 1 : 0));
  const int fixed = 1024 + (2 * 4 * pl.bn * 4 + 8 * 2 * pl.bn * 4) + (pl.staged ? 4 * 32 * 64 : 0) + 1024;  // align slack + scratch + staging + barriers
  if ((nout + pl.bn - 1) / pl.bn > 8) return pl;
  const int budget = 227 * 1024;
  pl.num_m_tiles = static_cast<int>((V + bm - 1) / bm);
  pl.num_n_tiles = (nout + pl.bn - 1) / pl.bn;
  if (pl.bn == 64) {
    // try the resident-filter configuration (NB = 9, 64->64 layers)
    pl.nb = 9;
    pl.resident = (slabs == 1 && pl.num_n_tiles == 1 && !is3d) ? 1 : 0;
    pl.nslab = 3;
    if (fixed + pl.nb * btile + pl.nslab * slab_bytes > budget) pl.nslab = 2;
  } else {
    pl.nb = 6;
    pl.resident = 0;
    pl.nslab = 3;
    if (fixed + pl.nb * btile + pl.nslab * slab_bytes > budget) pl.nb = 5;
    if (fixed + pl.nb * btile + pl.nslab * slab_bytes > budget) { pl.nb = 6; pl.nslab = 2; }
  }
  pl.smem = static_cast<size_t>(fixed) + static_cast<size_t>(pl.nb) * btile + static_cast<size_t>(pl.nslab) * slab_bytes;
  if (pl.smem > static_cast<size_t>(budget)) return pl;
  pl.ok = true;
  return pl;
}

// ------------------------------------------------------------------------------------------------
// First Conv3d layer (8 padded input channels -> 32): conv3d_c8.cuh
// ------------------------------------------------------------------------------------------------
bool c8_ok(const qt_conv_desc* d, int flags) {
  if (g_tune[7] == 1 || !g_use_conv3x3) return false;
  if (d->k_d != 3 || !slab_geometry(d) || d->groups != 1 || d->in_c != 8 || d->out_c != kC8N) return false;
  if (flags & (EPI_RELU | EPI_OUT_F32 | EPI_ADDEND)) return false;
  if (!dense_nhwc(d->x_stride, d->in_d, d->in_h, d->in_w, d->in_c) || !dense_nhwc(d->y_stride, d->in_d, d->in_h, d->in_w, d->out_c)) return false;
  const long long V = static_cast<long long>(d->n) * d->in_d * (d->in_h + 1) * (d->in_w + 2);
  return V <= (1ll << 30) && d->in_w <= 248;  // (the producers keep at most 6 / 5 slab rows per thread in registers)
}
int c8_tiles(const qt_conv_desc* d) {
  const long long V = static_cast<long long>(d->n) * d->in_d * (d->in_h + 1) * (d->in_w + 2);
  return static_cast<int>((V + 2 * kBM - 1) / (2 * kBM));
}
int run_conv3d_c8(const qt_conv_desc* d, const void* x, const void* wb, void* y, const float* bias, float* stats, cudaStream_t st) {
  Conv3dC8Params p;
  memset(&p, 0, sizeof(p));
  p.x = static_cast<const __nv_bfloat16*>(x);
  p.wb = static_cast<const __nv_bfloat16*>(wb);
  p.y = static_cast<__nv_bfloat16*>(y);
  p.bias = bias;
  p.stats = stats;
  p.NP = d->n * d->in_d; p.D = d->in_d; p.H = d->in_h; p.W = d->in_w;
  p.V = p.NP * (p.H + 1) * (p.W + 2);
  p.num_tiles = c8_tiles(d);
  p.R = 2 * kBM + 2 * (p.W + 3) + 8;
  const size_t slab = (static_cast<size_t>(p.R) * 16 + 1023) / 1024 * 1024;
  const size_t smem = 1024 + 20 * 1024 + kC8Slabs * slab;
  static size_t configured = 0;
  if (configured < smem) {
    cudaFuncSetAttribute(conv3d_c8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    configured = smem;
  }
  const int grid = p.num_tiles < kNumSMs ? p.num_tiles : kNumSMs;
  conv3d_c8_kernel<<<grid, kC8Threads, smem, st>>>(p);
  return cuda_status("conv3d_c8_kernel");
}

size_t c8_wgrad_ws_bytes() { return static_cast<size_t>(kNumSMs) * 9 * 32 * 32 * sizeof(float); }
int run_conv3d_c8_wgrad(const qt_conv_desc* d, const void* x, const void* dy, float* dw, int accumulate, void* ws, size_t ws_bytes,
                        cudaStream_t st) {
  if (ws == nullptr || ws_bytes < c8_wgrad_ws_bytes()) return fail("conv3d_c8 wgrad: workspace too small");
  Conv3dC8WgradParams p;
  memset(&p, 0, sizeof(p));
  p.x = static_cast<const __nv_bfloat16*>(x);
  p.dy = static_cast<const __nv_bfloat16*>(dy);
  p.partial = static_cast<float*>(ws);
  p.NP = d->n * d->in_d; p.D = d->in_d; p.H = d->in_h; p.W = d->in_w;
  p.V = p.NP * (p.H + 1) * (p.W + 2);
  p.num_tiles = (p.V + kC8wKP - 1) / kC8wKP;
  p.R = kC8wKP + 2 * (p.W + 3) + 8;
  p.debug = g_tune[9];
  const size_t slab = (static_cast<size_t>(p.R) * 16 + 1023) / 1024 * 1024;
  const size_t smem = 1024 + kC8wDy * ((8 * (kC8wKP * 16 + 16) + 127) / 128 * 128) + 1024 + kC8wSlabs * slab;
  static size_t configured = 0;
  if (configured < smem) {
    cudaFuncSetAttribute(conv3d_c8_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    configured = smem;
  }
  const int grid = p.num_tiles < kNumSMs ? p.num_tiles : kNumSMs;
  conv3d_c8_wgrad_kernel<<<grid, kC8wThreads, smem, st>>>(p);
  if (int rc = cuda_status("conv3d_c8_wgrad_kernel")) return rc;
  conv3d_c8_wgrad_reduce_kernel<<<(32 * 8 * 27 + 255) / 256, 256, 0, st>>>(p.partial, grid, dw, 8, accumulate);
  return cuda_status("conv3d_c8_wgrad_reduce");
}

// cuTensorMapEncodeTiled through the runtime's driver-entry-point query (no link-time libcuda dependency).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}
// 2-D bf16 matrix [rows][cols] (cols contiguous), box [box_rows][64 cols], 128B swizzle, zero OOB fill.
bool make_weight_tmap(CUtensorMap* map, const void* base, long long rows, long long cols, int box_rows) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return false;
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(cols) * 2};
  const cuuint32_t box[2] = {64, static_cast<cuuint32_t>(box_rows)};
  const cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int BN, int MT, int NSLAB, int NB, bool STAGED>
int launch_conv3x3(Conv3x3Params& p, size_t smem, int grid, cudaStream_t st) {
  static size_t configured = 0;
  if (configured < smem) {
    cudaFuncSetAttribute(conv3x3_kernel<BN, MT, NSLAB, NB, STAGED>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    configured = smem;
  }
  alignas(64) CUtensorMap wmap;
  memset(&wmap, 0, sizeof(wmap));
  p.b_tma = (g_tune[6] == 0 && (reinterpret_cast<uintptr_t>(p.b) & 15) == 0 &&
             make_weight_tmap(&wmap, p.b, p.nout, p.pair ? 2LL * 9 * 64 : static_cast<long long>(p.wtaps) * p.cin, BN))
                ? 1 : 0;
  conv3x3_kernel<BN, MT, NSLAB, NB, STAGED><<<grid, kC3Threads, smem, st>>>(p, wmap);
  return cuda_status("conv3x3_kernel");
}

// taps: per filter tap the input offset (dh, dw) and its index in the weight tensor.
int run_conv3x3(const C3Plan& pl, const qt_conv_desc* d, int cin, int nout, const void* a, const void* b, void* out,
                const void* addend, float* stats, int flags, bool dgrad, cudaStream_t st, const float* bias = nullptr) {
  Conv3x3Params p;
  memset(&p, 0, sizeof(p));
  p.a = static_cast<const __nv_bfloat16*>(a);
  p.b = static_cast<const __nv_bfloat16*>(b);
  p.out = out;
  p.addend = static_cast<const __nv_bfloat16*>(addend);
  p.stats = stats;
  p.N = d->n * d->in_d; p.H = d->in_h; p.W = d->in_w;
  p.D = d->in_d; p.kdn = d->k_d; p.pair = pl.pair;
  p.cin = cin; p.nout = nout; p.wtaps = 9 * d->k_d;
  p.flags = flags;
  p.bias_on = (flags & EPI_BIAS) ? 1 : 0;
  p.bias = bias;
  p.slabs = pl.pair ? 2 : d->k_d * (cin / 64);
  for (int kd = 0; kd < 3; ++kd) {
    p.dplane[kd] = static_cast<signed char>(d->k_d == 3 ? (dgrad ? 1 - kd : kd - 1) : 0);
    p.wkd[kd] = static_cast<signed char>(d->k_d == 3 ? kd : 0);
  }
  p.V = pl.V;
  p.num_m_tiles = pl.num_m_tiles; p.num_n_tiles = pl.num_n_tiles;
  p.R = pl.R; p.b_resident = pl.resident;
  int t = 0;
  for (int kh = 0; kh < 3; ++kh)
    for (int kw = 0; kw < 3; ++kw, ++t) {
      p.off_h[t] = static_cast<signed char>(dgrad ? 1 - kh : kh - 1);
      p.off_w[t] = static_cast<signed char>(dgrad ? 1 - kw : kw - 1);
      p.wtap[t] = static_cast<short>(t);
    }
  const int tiles = pl.num_m_tiles * pl.num_n_tiles;
  const int grid = tiles < kNumSMs ? tiles : kNumSMs;
  if (pl.bn == 64) {
    if (pl.staged) {
      if (pl.nslab == 3) return launch_conv3x3<64, 2, 3, 9, true>(p, pl.smem, grid, st);
      return launch_conv3x3<64, 2, 2, 9, true>(p, pl.smem, grid, st);
    }
    if (pl.nslab == 3) return launch_conv3x3<64, 2, 3, 9, false>(p, pl.smem, grid, st);
    return launch_conv3x3<64, 2, 2, 9, false>(p, pl.smem, grid, st);
  }
  if (pl.staged) {
    if (pl.nslab == 3 && pl.nb == 6) return launch_conv3x3<128, 2, 3, 6, true>(p, pl.smem, grid, st);
    if (pl.nslab == 3) return launch_conv3x3<128, 2, 3, 5, true>(p, pl.smem, grid, st);
    return launch_conv3x3<128, 2, 2, 6, true>(p, pl.smem, grid, st);
  }
  if (pl.nslab == 3 && pl.nb == 6) return launch_conv3x3<128, 2, 3, 6, false>(p, pl.smem, grid, st);
  if (pl.nslab == 3) return launch_conv3x3<128, 2, 3, 5, false>(p, pl.smem, grid, st);
  return launch_conv3x3<128, 2, 2, 6, false>(p, pl.smem, grid, st);
}


// ------------------------------------------------------------------------------------------------
// Slab weight-gradient kernel dispatch (wgrad3x3.cuh); g_tune[3] = 1 forces the generic gather kernel.
// ------------------------------------------------------------------------------------------------
struct W3Plan {
  bool ok;
  int cfg;  // 0: 64->64 (NSLAB 1, 5 taps, 3 stages), 1: 128-multiples (NSLAB 2, 3 taps, 2 stages)
  int V, num_kt, splits, kt_per_split, R, cout_tiles, cin_groups, tap_groups;
  size_t smem, ws_bytes;
};
W3Plan plan_wgrad3x3(const qt_conv_desc* d) {
  W3Plan pl{};
  pl.ok = false;
  if (g_tune[3] != 0) return pl;
  if (!slab_geometry(d) || d->groups != 1) return pl;
  if (d->k_d == 3 && g_tune[7] == 1) return pl;
  if (!dense_nhwc(d->x_stride, d->in_d, d->in_h, d->in_w, d->in_c) || !dense_nhwc(d->y_stride, d->in_d, d->in_h, d->in_w, d->out_c)) return pl;
  if (d->in_w < 10 && g_tune[4] == 1) return pl;  // knob 4: send small maps (7x7) to the gather kernel instead
  if (16 / (d->in_w + 2) >= d->in_h + 1) return pl;  // single-carry coordinate advance (see plan_conv3x3)
  if (d->in_c == 64 && d->out_c == 64 && d->k_d == 1) pl.cfg = 0;
  else if (d->in_c % 128 == 0 && d->out_c % 128 == 0) pl.cfg = 1;
  else if (d->in_c == 64 && d->out_c % 128 == 0) pl.cfg = 2;  // 64 -> 128k (Conv3d block3): one 64-channel slab, 128-cout dy tiles
  else if (d->k_d == 3 && d->in_c == 32 && d->out_c == 64) pl.cfg = 3;  // Conv3d block2: pair3d on the 64 -> 64 configuration
  else return pl;
  const long long V = static_cast<long long>(d->n) * d->in_d * (d->in_h + 1) * (d->in_w + 2);
  if (V > (1ll << 30)) return pl;
  pl.V = static_cast<int>(V);
  pl.num_kt = static_cast<int>((V + kW3KP - 1) / kW3KP);
  pl.R = (kW3KP + 2 * (d->in_w + 3) + 7) / 8 * 8;
  // cfg 0 pairs horizontally adjacent taps in one MMA (wgrad3x3.cuh): 6 MMA groups cover all 9 taps in a single CTA
  const bool c0like = pl.cfg == 0 || pl.cfg == 3;
  const int nslab = pl.cfg == 1 ? 2 : 1, cb = c0like ? 1 : 2, stages = c0like ? 4 : (pl.cfg == 1 ? 2 : 3), taps = c0like ? 9 : 3;
  const int dy_rows = c0like ? kW3KP + 8 : kW3KP;
  pl.cout_tiles = d->out_c / (64 * cb);
  pl.cin_groups = pl.cfg == 3 ? 1 : d->in_c / (64 * nslab);
  pl.tap_groups = (9 + taps - 1) / taps;
  const int types = pl.cout_tiles * pl.cin_groups * pl.tap_groups * (pl.cfg == 3 ? 2 : d->k_d);
  int splits = kNumSMs / types;
  if (splits < 1) splits = 1;
  if (splits > pl.num_kt / 4) splits = pl.num_kt / 4;
  if (splits < 1) splits = 1;
  pl.kt_per_split = (pl.num_kt + splits - 1) / splits;
  pl.splits = (pl.num_kt + pl.kt_per_split - 1) / pl.kt_per_split;
  const size_t slab_bytes = static_cast<size_t>(nslab) * ((static_cast<size_t>(pl.R) * 128 + 1023) / 1024 * 1024);
  pl.smem = 1024 + stages * (static_cast<size_t>(cb) * dy_rows * 128 + slab_bytes) + 256;
  if (pl.smem > 227 * 1024) return pl;
  pl.ws_bytes = static_cast<size_t>(pl.splits) * 9 * d->k_d * d->in_c * d->out_c * sizeof(float);
  pl.ok = true;
  return pl;
}

template <int NSLAB, int TAPS, int CB, int STAGES, int NMMA>
int launch_wgrad3x3(const Wgrad3x3Params& p, const W3Plan& pl, cudaStream_t st) {
  static size_t configured = 0;
  if (configured < pl.smem) {
    cudaFuncSetAttribute(wgrad3x3_kernel<NSLAB, TAPS, CB, STAGES, NMMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(pl.smem));
    configured = pl.smem;
  }
  dim3 grid(pl.cout_tiles * pl.cin_groups * pl.tap_groups, pl.splits, p.pair3d ? 2 : p.kdn);
  wgrad3x3_kernel<NSLAB, TAPS, CB, STAGES, NMMA><<<grid, 192, pl.smem, st>>>(p);
  return cuda_status("wgrad3x3_kernel");
}

int run_wgrad3x3(const W3Plan& pl, const qt_conv_desc* d, const void* x, const void* dy, float* dw, int accumulate, void* ws,
                 size_t ws_bytes, cudaStream_t st) {
  if (ws == nullptr || ws_bytes < pl.ws_bytes) return fail("wgrad3x3: workspace too small (%zu < %zu)", ws_bytes, pl.ws_bytes);
  Wgrad3x3Params p;
  memset(&p, 0, sizeof(p));
  p.x = static_cast<const __nv_bfloat16*>(x);
  p.dy = static_cast<const __nv_bfloat16*>(dy);
  p.ws = static_cast<float*>(ws);
  p.N = d->n * d->in_d; p.H = d->in_h; p.W = d->in_w; p.cin = d->in_c; p.cout = d->out_c;
  p.D = d->in_d; p.kdn = d->k_d; p.pair3d = pl.cfg == 3 ? 1 : 0;
  p.V = pl.V; p.num_kt = pl.num_kt; p.kt_per_split = pl.kt_per_split;
  p.R = pl.R;
  p.cout_tiles = pl.cout_tiles; p.cin_groups = pl.cin_groups; p.tap_groups = pl.tap_groups;
  int t = 0;
  for (int kh = 0; kh < 3; ++kh)
    for (int kw = 0; kw < 3; ++kw, ++t) { p.off_h[t] = static_cast<signed char>(kh - 1); p.off_w[t] = static_cast<signed char>(kw - 1); }
  int rc = (pl.cfg == 0 || pl.cfg == 3) ? launch_wgrad3x3<1, 6, 1, 4, 2>(p, pl, st)
                       : (pl.cfg == 1 ? launch_wgrad3x3<2, 3, 2, 2, 2>(p, pl, st) : launch_wgrad3x3<1, 3, 2, 3, 2>(p, pl, st));
  if (rc) return rc;
  const int taps = 9 * d->k_d;
  launch_splitk_reduce_wgrad(p.ws, pl.splits, taps * d->in_c, d->out_c, taps * d->in_c, d->out_c, d->in_c, taps, dw, accumulate, st);
  return cuda_status("splitk_reduce_wgrad_kernel");
}

template <typename K, typename... Args>
int launch_cluster8(K kernel, int clusters, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(clusters * kLstmCluster);
  cfg.blockDim = dim3(kLstmThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = kLstmCluster;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  const cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, args...);
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof(g_err), "lstm cluster launch: %s", cudaGetErrorString(e));
    return static_cast<int>(e);
  }
  return 0;
}

}  // namespace

// ================================================================================================
extern "C" {

int qt_version(void) { return 200; }  // 200: round 2 (loss / head tail / Adam / uint8 input / vectorised quadtree stage)
void qt_set_conv3x3_enabled(int on) { g_use_conv3x3 = on != 0; }
void qt_set_tuning(int key, int value) { if (key >= 0 && key < 16) g_tune[key] = value; }
const char* qt_last_error(void) { return g_err; }
int qt_take_timeout_flag(void) {
  unsigned int v = 0, z = 0;
  cudaMemcpyFromSymbol(&v, g_timeout_flag, sizeof(v));
  if (v) cudaMemcpyToSymbol(g_timeout_flag, &z, sizeof(z));
  return static_cast<int>(v);
}

#ifdef QT_TRACE
/* developer build only (not declared in include/qtcnn.h): copies the per-CTA wait counters to the host and clears them */
int qt_debug_read_trace(long long* host, int count) {
  if (count > 1024 * 16) count = 1024 * 16;
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(host, g_trace, sizeof(long long) * count);
  static long long zeros[1024 * 16];
  cudaMemcpyToSymbol(g_trace, zeros, sizeof(zeros));
  return 0;
}
#endif

// ---- layout / packing -----------------------------------------------------------------------------
int qt_stem_pack_input_ex(const void* x, int dtype, const float* scale, const float* shift, void* xp, int n, int c, int h, int w,
                          qt_stream_t stream) {
  if (c > 4) return fail("stem_pack_input: at most 4 channels");
  const long long rows = static_cast<long long>(n) * (h + 7);
  if (rows > 0x7fffffffLL) return fail("stem_pack_input: too many rows");
  if ((scale == nullptr) != (shift == nullptr)) return fail("stem_pack_input: scale and shift come together");
  if (dtype == QT_DTYPE_U8 && !scale) return fail("stem_pack_input: uint8 input needs per-channel scale / shift");
  const int block = (w + 8) > 128 ? 256 : ((w + 8) > 64 ? 128 : 64);
  __nv_bfloat16* o = static_cast<__nv_bfloat16*>(xp);
  const unsigned grid = static_cast<unsigned>(rows);
  if (dtype == QT_DTYPE_F32) stem_pack_input_kernel<float><<<grid, block, 0, S(stream)>>>(static_cast<const float*>(x), o, n, c, h, w, scale, shift);
  else if (dtype == QT_DTYPE_BF16)
    stem_pack_input_kernel<__nv_bfloat16><<<grid, block, 0, S(stream)>>>(static_cast<const __nv_bfloat16*>(x), o, n, c, h, w, scale, shift);
  else if (dtype == QT_DTYPE_U8) stem_pack_input_kernel<uint8_t><<<grid, block, 0, S(stream)>>>(static_cast<const uint8_t*>(x), o, n, c, h, w, scale, shift);
  else return fail("stem_pack_input: unknown dtype %d", dtype);
  return cuda_status("stem_pack_input");
}
int qt_stem_pack_input(const float* x, void* xp, int n, int c, int h, int w, qt_stream_t stream) {
  return qt_stem_pack_input_ex(x, QT_DTYPE_F32, nullptr, nullptr, xp, n, c, h, w, stream);
}
int qt_nchw_to_nhwc_bf16_ex(const void* x, int dtype, const float* scale, const float* shift, void* out, int n, int c, long long hw,
                            int c_pad, qt_stream_t stream) {
  const long long total = static_cast<long long>(n) * hw * c_pad;
  if ((scale == nullptr) != (shift == nullptr)) return fail("nchw_to_nhwc: scale and shift come together");
  if (dtype == QT_DTYPE_U8 && !scale) return fail("nchw_to_nhwc: uint8 input needs per-channel scale / shift");
  __nv_bfloat16* o = static_cast<__nv_bfloat16*>(out);
  if (c_pad == 8 && c <= 8 && (dtype == QT_DTYPE_F32 || dtype == QT_DTYPE_BF16 || dtype == QT_DTYPE_U8)) {
    const int g8 = grid_for(static_cast<long long>(n) * hw, 256, 148 * 32);
    if (dtype == QT_DTYPE_F32) nchw_to_nhwc8_bf16_kernel<float><<<g8, 256, 0, S(stream)>>>(static_cast<const float*>(x), o, n, c, hw, scale, shift);
    else if (dtype == QT_DTYPE_BF16)
      nchw_to_nhwc8_bf16_kernel<__nv_bfloat16><<<g8, 256, 0, S(stream)>>>(static_cast<const __nv_bfloat16*>(x), o, n, c, hw, scale, shift);
    else nchw_to_nhwc8_bf16_kernel<uint8_t><<<g8, 256, 0, S(stream)>>>(static_cast<const uint8_t*>(x), o, n, c, hw, scale, shift);
    return cuda_status("nchw_to_nhwc8_bf16");
  }
  const int grid = grid_for(total, 256, 1 << 20);
  if (dtype == QT_DTYPE_F32) nchw_to_nhwc_bf16_kernel<float><<<grid, 256, 0, S(stream)>>>(static_cast<const float*>(x), o, n, c, hw, c_pad, scale, shift);
  else if (dtype == QT_DTYPE_BF16)
    nchw_to_nhwc_bf16_kernel<__nv_bfloat16><<<grid, 256, 0, S(stream)>>>(static_cast<const __nv_bfloat16*>(x), o, n, c, hw, c_pad, scale, shift);
  else if (dtype == QT_DTYPE_U8) nchw_to_nhwc_bf16_kernel<uint8_t><<<grid, 256, 0, S(stream)>>>(static_cast<const uint8_t*>(x), o, n, c, hw, c_pad, scale, shift);
  else return fail("nchw_to_nhwc: unknown dtype %d", dtype);
  return cuda_status("nchw_to_nhwc_bf16");
}
int qt_nchw_f32_to_nhwc_bf16(const float* x, void* out, int n, int c, long long hw, int c_pad, qt_stream_t stream) {
  return qt_nchw_to_nhwc_bf16_ex(x, QT_DTYPE_F32, nullptr, nullptr, out, n, c, hw, c_pad, stream);
}
int qt_nhwc_bf16_to_nchw_f32(const void* x, float* out, int n, int c, long long hw, int c_pad, qt_stream_t stream) {
  const long long total = static_cast<long long>(n) * hw * c;
  nhwc_bf16_to_nchw_f32_kernel<<<grid_for(total, 256, 1 << 20), 256, 0, S(stream)>>>(static_cast<const __nv_bfloat16*>(x), out, n, c, hw, c_pad);
  return cuda_status("nhwc_bf16_to_nchw_f32");
}
int qt_wpack_fprop(const float* w, void* wf, int cout, int cin, int taps, qt_stream_t stream) {
  const long long total = static_cast<long long>(cout) * cin * taps;
  wpack_fprop_kernel<<<grid_for(total, 256, 1 << 20), 256, 0, S(stream)>>>(w, static_cast<__nv_bfloat16*>(wf), cout, cin, taps);
  return cuda_status("wpack_fprop");
}
int qt_wpack_dgrad(const float* w, void* wd, int cout, int cin, int taps, qt_stream_t stream) {
  dim3 grid((cin + 31) / 32, (cout + 31) / 32, taps), block(32, 8);
  wpack_dgrad_kernel<<<grid, block, 0, S(stream)>>>(w, static_cast<__nv_bfloat16*>(wd), cout, cin, taps);
  return cuda_status("wpack_dgrad");
}
namespace {
inline int wpack_co_tile(int taps) { return taps == 1 ? 32 : 8; }
inline size_t wpack_smem(int taps) { return static_cast<size_t>(wpack_co_tile(taps)) * (32 * (taps | 1) + 1) * sizeof(float); }
}  // namespace
int qt_wpack_both(const float* w, void* wf, void* wd, int cout, int cin, int taps, qt_stream_t stream) {
  if (taps < 1 || taps > 32) return fail("wpack_both: taps must be 1..32");
  const int cot = wpack_co_tile(taps);
  dim3 grid((cin + 31) / 32, (cout + cot - 1) / cot);
  wpack_both_kernel<<<grid, 256, wpack_smem(taps), S(stream)>>>(w, static_cast<__nv_bfloat16*>(wf), static_cast<__nv_bfloat16*>(wd),
                                                                 cout, cin, taps, cot);
  return cuda_status("wpack_both");
}
int qt_wpack_item_plan(qt_wpack_item* item) {
  if (!item || item->cout < 1 || item->cin < 1 || item->taps < 1 || item->taps > 32) return fail("wpack_item_plan: bad item");
  item->co_tile = wpack_co_tile(item->taps);
  item->ci_tiles = (item->cin + 31) / 32;
  return item->ci_tiles * ((item->cout + item->co_tile - 1) / item->co_tile);
}
int qt_wpack_multi(const void* items_dev, int nitems, int total_blocks, int max_taps, qt_stream_t stream) {
  static_assert(sizeof(qt_wpack_item) == sizeof(WpackItem), "qt_wpack_item layout");
  if (nitems < 1 || total_blocks < 1) return 0;
  if (max_taps < 1 || max_taps > 32) return fail("wpack_multi: taps must be 1..32");
  size_t smem = wpack_smem(1);
  for (int t = 2; t <= max_taps; ++t) smem = wpack_smem(t) > smem ? wpack_smem(t) : smem;
  wpack_multi_kernel<<<total_blocks, 256, smem, S(stream)>>>(static_cast<const WpackItem*>(items_dev), nitems);
  return cuda_status("wpack_multi");
}
int qt_wpack_conv3d_c8(const float* w, void* wb, int cout, int cin, qt_stream_t stream) {
  if (cout != kC8N || cin < 1 || cin > 8) return fail("wpack_conv3d_c8: expects a [32][cin <= 8][3][3][3] weight");
  wpack_conv3d_c8_kernel<<<(18 * 2 * 32 * 8 + 255) / 256, 256, 0, S(stream)>>>(w, static_cast<__nv_bfloat16*>(wb), cin);
  return cuda_status("wpack_conv3d_c8");
}
int qt_wpack_conv3d_pair(const float* w, void* wp, int cout, qt_stream_t stream) {
  const long long total = static_cast<long long>(cout) * 2 * 9 * 64;
  wpack_conv3d_pair_kernel<<<grid_for(total, 256, 1 << 20), 256, 0, S(stream)>>>(w, static_cast<__nv_bfloat16*>(wp), cout);
  return cuda_status("wpack_conv3d_pair");
}
int qt_wpack_stem(const float* w, void* w8, int cout, int cin, int r, int s, qt_stream_t stream) {
  if (cin > 4 || r > 8 || s > 8) return fail("wpack_stem: filter does not fit the 8x(8x4) packing");
  wpack_stem_kernel<<<grid_for(cout * 256, 256), 256, 0, S(stream)>>>(w, static_cast<__nv_bfloat16*>(w8), cout, cin, r, s);
  return cuda_status("wpack_stem");
}
int qt_f32_to_bf16(const float* x, void* out, long long n, qt_stream_t stream) {
  f32_to_bf16_kernel<<<grid_for(n, 256, 1 << 20), 256, 0, S(stream)>>>(x, static_cast<__nv_bfloat16*>(out), n);
  return cuda_status("f32_to_bf16");
}

// ---- convolutions -----------------------------------------------------------------------------------
/* which kernel a pass will use: 0 generic gather GEMM, 1 persistent slab kernel (conv3x3 / wgrad3x3) */
int qt_conv_plan(const qt_conv_desc* d, int pass) {
  if (check_desc(d)) return -1;
  if (pass == 0) {
    if (c8_ok(d, EPI_STATS | EPI_BIAS)) return 3;  // first Conv3d layer: conv3d_c8 kernel on qt_wpack_conv3d_c8 weights
    const C3Plan pl = plan_conv3x3(d, d->in_c, d->out_c, EPI_STATS | (d->k_d == 3 ? EPI_BIAS : 0));
    return pl.ok ? (pl.pair ? 2 : 1) : 0;  // 2: slab kernel on pair-packed weights (qt_wpack_conv3d_pair)
  }
  if (pass == 1) return plan_conv3x3(d, d->out_c, d->in_c, 0, true).ok ? 1 : 0;
  if (c8_ok(d, 0)) return 3;
  return plan_wgrad3x3(d).ok ? 1 : 0;
}
int qt_conv_stat_rows(const qt_conv_desc* d) {
  if (check_desc(d)) return -1;
  if (c8_ok(d, EPI_STATS | EPI_BIAS)) { const int t = c8_tiles(d); return t < kNumSMs ? t : kNumSMs; }
  const C3Plan pl = plan_conv3x3(d, d->in_c, d->out_c, EPI_STATS | (d->k_d == 3 ? EPI_BIAS : 0));
  if (pl.ok) { const int tiles = pl.num_m_tiles * pl.num_n_tiles; return tiles < kNumSMs ? tiles : kNumSMs; }
  const OutDims o = conv_out_dims(d);
  const long long M = static_cast<long long>(d->n) * o.d * o.h * o.w;
  return static_cast<int>(d->groups * ((M + kBM - 1) / kBM));
}
size_t qt_conv_fprop_workspace_bytes(const qt_conv_desc* d) { (void)d; return 0; }

int qt_conv_fprop(const qt_conv_desc* d, const void* x, const void* wf, void* y, const float* bias, float* stats,
                  int flags, void* ws, size_t ws_bytes, qt_stream_t stream) {
  if (int rc = check_desc(d)) return rc;
  IgemmParams p;
  if (int rc = fill_forward(p, d)) return rc;
  p.a = static_cast<const __nv_bfloat16*>(x);
  p.b = static_cast<const __nv_bfloat16*>(wf);
  p.out = y;
  p.bias = bias;
  p.stats = stats;
  p.flags = flags & (EPI_BIAS | EPI_RELU | EPI_STATS | EPI_OUT_F32);
  if ((p.flags & EPI_BIAS) && !bias) return fail("conv_fprop: QT_EPI_BIAS without bias");
  if ((p.flags & EPI_STATS) && !stats) return fail("conv_fprop: QT_EPI_STATS without stats buffer");
  if (c8_ok(d, p.flags)) return run_conv3d_c8(d, x, wf, y, (p.flags & EPI_BIAS) ? bias : nullptr, (p.flags & EPI_STATS) ? stats : nullptr, S(stream));
  const C3Plan pl = plan_conv3x3(d, d->in_c, d->out_c, p.flags);
  if (pl.ok) return run_conv3x3(pl, d, d->in_c, d->out_c, x, wf, y, nullptr, stats, p.flags, false, S(stream), bias);
  return run_kmajor(p, S(stream), ws, ws_bytes, 0);
}

int qt_conv_dgrad(const qt_conv_desc* d, const void* dy, const void* wd, void* dx, int accumulate,
                  qt_stream_t stream) {
  if (int rc = check_desc(d)) return rc;
  const OutDims o = conv_out_dims(d);
  const int sd = d->stride_d, sh = d->stride_h, sw = d->stride_w;
  if (ilog2_exact(d->out_c) < 0 && d->k_d * d->k_h * d->k_w > 1) return fail("conv_dgrad: out_c must be a power of two");
  {
    const C3Plan pl = plan_conv3x3(d, d->out_c, d->in_c, accumulate ? EPI_ADDEND : 0, true);
    if (pl.ok)
      return run_conv3x3(pl, d, d->out_c, d->in_c, dy, wd, dx, accumulate ? dx : nullptr, nullptr, accumulate ? EPI_ADDEND : 0,
                         true, S(stream));
  }
  bool any_empty = false;
  // One launch per stride-parity class; for stride 1 there is exactly one.
  for (int pd = 0; pd < sd; ++pd)
    for (int ph = 0; ph < sh; ++ph)
      for (int pw = 0; pw < sw; ++pw) {
        IgemmParams p;
        memset(&p, 0, sizeof(p));
        int t = 0;
        for (int kd = 0; kd < d->k_d; ++kd) {
          if ((pd + d->pad_d - kd) % sd) continue;
          for (int kh = 0; kh < d->k_h; ++kh) {
            if ((ph + d->pad_h - kh) % sh) continue;
            for (int kw = 0; kw < d->k_w; ++kw) {
              if ((pw + d->pad_w - kw) % sw) continue;
              p.off_d[t] = static_cast<signed char>((pd + d->pad_d - kd) / sd);
              p.off_h[t] = static_cast<signed char>((ph + d->pad_h - kh) / sh);
              p.off_w[t] = static_cast<signed char>((pw + d->pad_w - kw) / sw);
              p.wtap[t] = static_cast<short>((kd * d->k_h + kh) * d->k_w + kw);
              ++t;
            }
          }
        }
        const int gd = (d->in_d - pd + sd - 1) / sd, gh = (d->in_h - ph + sh - 1) / sh, gw = (d->in_w - pw + sw - 1) / sw;
        if (gd < 1 || gh < 1 || gw < 1) continue;
        if (t == 0) { any_empty = true; continue; }
        p.ntaps = t;
        p.wtaps = d->k_d * d->k_h * d->k_w;
        p.cin = d->out_c;
        p.cin_log2 = ilog2_exact(d->out_c);
        p.a = static_cast<const __nv_bfloat16*>(dy);
        p.av = {d->y_stride[0], d->y_stride[1], d->y_stride[2], d->y_stride[3]};
        p.id = o.d; p.ih = o.h; p.iw = o.w;
        p.nb = d->n; p.od = gd; p.oh = gh; p.ow = gw;
        p.mult_d = p.mult_h = p.mult_w = 1;
        p.b = static_cast<const __nv_bfloat16*>(wd);
        p.nout = d->in_c;
        p.out = dx;
        p.ov = {d->x_stride[0], d->x_stride[1] * sd, d->x_stride[2] * sh, d->x_stride[3] * sw};
        fill_groups(p, d, false);
        const long long poff = pd * d->x_stride[1] + ph * d->x_stride[2] + pw * d->x_stride[3];
        for (int g = 0; g < 4; ++g) p.o_goff[g] += poff;
        p.M = d->n * gd * gh * gw;
        p.flags = accumulate ? EPI_ADDEND : 0;
        p.addend = static_cast<const __nv_bfloat16*>(dx);
        if (int rc = run_kmajor(p, S(stream), nullptr, 0, 0)) return rc;
      }
  if (any_empty && !accumulate)
    return fail("conv_dgrad: stride leaves input positions without taps; zero dx and call with accumulate=1");
  return 0;
}

size_t qt_conv_wgrad_workspace_bytes(const qt_conv_desc* d) {
  if (check_desc(d)) return 0;
  const OutDims o = conv_out_dims(d);
  const long long M = static_cast<long long>(d->n) * o.d * o.h * o.w;
  const size_t generic = wgrad_ws_bytes(d->k_d * d->k_h * d->k_w * d->in_c, d->out_c, d->groups, M, nullptr);
  const W3Plan pl = plan_wgrad3x3(d);
  size_t need = (pl.ok && pl.ws_bytes > generic) ? pl.ws_bytes : generic;
  if (c8_ok(d, 0) && c8_wgrad_ws_bytes() > need) need = c8_wgrad_ws_bytes();
  return need;
}
int qt_conv_wgrad(const qt_conv_desc* d, const void* x, const void* dy, float* dw, int accumulate, void* ws,
                  size_t ws_bytes, qt_stream_t stream) {
  if (int rc = check_desc(d)) return rc;
  if (c8_ok(d, 0)) return run_conv3d_c8_wgrad(d, x, dy, dw, accumulate, ws, ws_bytes, S(stream));
  {
    const W3Plan pl = plan_wgrad3x3(d);
    if (pl.ok) return run_wgrad3x3(pl, d, x, dy, dw, accumulate, ws, ws_bytes, S(stream));
  }
  IgemmParams p;
  if (int rc = fill_forward(p, d)) return rc;
  p.a = static_cast<const __nv_bfloat16*>(x);
  p.b = static_cast<const __nv_bfloat16*>(dy);
  for (int g = 0; g < 4; ++g) p.b_goff[g] = d->y_group_off[g];
  return run_wgrad(p, S(stream), ws, ws_bytes, dw, accumulate, d->in_c, p.ntaps);
}

// ---- stem (7x7 s2 p3 on 3 channels) as an 8-tap x 32-"channel" GEMM over the packed input ---------------
static int fill_stem(IgemmParams& p, int n, int h, int w, int cout) {
  if (h % 2 || w % 2) return fail("stem: even image sizes only");
  memset(&p, 0, sizeof(p));
  const int Hp = h + 7, Wp = w + 8;
  p.av = {static_cast<long long>(Hp) * Wp * 4, 0, static_cast<long long>(Wp) * 4, 4};
  p.id = 1; p.ih = Hp; p.iw = Wp;  // packed buffer already holds the zero padding
  p.nb = n; p.od = 1; p.oh = h / 2; p.ow = w / 2;
  p.mult_d = 1; p.mult_h = 2; p.mult_w = 2;
  p.ntaps = 8; p.wtaps = 8; p.cin = 32; p.cin_log2 = 5;
  for (int r = 0; r < 8; ++r) { p.off_d[r] = 0; p.off_h[r] = static_cast<signed char>(r); p.off_w[r] = 0; p.wtap[r] = static_cast<short>(r); }
  p.nout = cout;
  const long long ohw = static_cast<long long>(p.oh) * p.ow;
  p.ov = {ohw * cout, 0, static_cast<long long>(p.ow) * cout, cout};
  p.M = n * p.oh * p.ow;
  p.groups = 1;
  return 0;
}
// Dedicated overlapping-row stem kernels (stem.cuh) when the shape allows; g_tune[2] = 1 forces the generic path.
static bool stem_fast_ok(int h, int w, int cout) { return g_tune[2] == 0 && cout == 64 && h % 2 == 0 && w % 2 == 0 && w <= 248; }
static bool stem_wgrad_fast_ok(int h, int w, int cout, int cin) {
  return stem_fast_ok(h, w, cout) && cin <= 4 && ((w / 2) % 16 == 0);
}
static int stem_tiles(int n, int h, int rows) { return n * (((h / 2) + rows - 1) / rows); }

int qt_stem_stat_rows(int n, int h, int w) {
  if (stem_fast_ok(h, w, 64)) { const int t = stem_tiles(n, h, kStemRows); return t < kNumSMs ? t : kNumSMs; }
  return (n * (h / 2) * (w / 2) + kBM - 1) / kBM;
}
int qt_stem_fprop(const void* xp, const void* w8, void* y, float* stats, int n, int h, int w, int cout,
                  qt_stream_t stream) {
  if (stem_fast_ok(h, w, cout)) {
    StemParams sp;
    memset(&sp, 0, sizeof(sp));
    sp.xp = static_cast<const __nv_bfloat16*>(xp);
    sp.w8 = static_cast<const __nv_bfloat16*>(w8);
    sp.y = static_cast<__nv_bfloat16*>(y);
    sp.stats = stats;
    sp.N = n; sp.H = h; sp.W = w; sp.Ho = h / 2; sp.Wo = w / 2; sp.Hp = h + 7; sp.Wp = w + 8;
    sp.strips = (sp.Ho + kStemRows - 1) / kStemRows;
    sp.num_tiles = n * sp.strips;
    sp.row_bytes = sp.Wp * 8;
    sp.stage_bytes = ((2 * kStemRows + 5) * sp.row_bytes + 2304 + 127) / 128 * 128;  // slack: MMA rows up to 127 read 2096 B past the last row start
    const size_t smem = 1024 + 28 * 1024 + 4096 + 256 + static_cast<size_t>(kStemStages) * sp.stage_bytes;
    static size_t configured = 0;
    if (configured < smem) {
      cudaFuncSetAttribute(stem_fprop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
      configured = smem;
    }
    const int grid = sp.num_tiles < kNumSMs ? sp.num_tiles : kNumSMs;
    stem_fprop_kernel<<<grid, kStemThreads, smem, S(stream)>>>(sp);
    return cuda_status("stem_fprop_kernel");
  }
  IgemmParams p;
  if (int rc = fill_stem(p, n, h, w, cout)) return rc;
  p.a = static_cast<const __nv_bfloat16*>(xp);
  p.b = static_cast<const __nv_bfloat16*>(w8);
  p.out = y;
  p.stats = stats;
  p.flags = stats ? EPI_STATS : 0;
  return run_kmajor(p, S(stream), nullptr, 0, 0);
}
/* r3d_18 stem (torchvision video/resnet.py BasicStem: Conv3d(3, 64, (3,7,7), stride (1,2,2), pad (1,3,3), no bias)) on the packed
 * frames [n*t][h+7][w+8][4]: the 7x7 part is the 8-row x (8 pixels x 4 channels) form of the 2-D stem, the three depth taps
 * are whole-frame shifts with zero fill outside the clip -> a 24-tap x 32-"channel" gather GEMM. */
int qt_stem3d_stat_rows(int n, int t, int h, int w) {
  return static_cast<int>((static_cast<long long>(n) * t * (h / 2) * (w / 2) + kBM - 1) / kBM);
}
int qt_stem3d_fprop(const void* xp, const void* w24, void* y, float* stats, int n, int t, int h, int w, int cout,
                    qt_stream_t stream) {
  if (h % 2 || w % 2) return fail("stem3d: even frame sizes only");
  IgemmParams p;
  memset(&p, 0, sizeof(p));
  const int Hp = h + 7, Wp = w + 8;
  const long long frame = static_cast<long long>(Hp) * Wp * 4;
  p.av = {frame * t, frame, static_cast<long long>(Wp) * 4, 4};
  p.id = t; p.ih = Hp; p.iw = Wp;  // the packed frames already hold the spatial zero padding; depth padding = zero fill
  p.nb = n; p.od = t; p.oh = h / 2; p.ow = w / 2;
  p.mult_d = 1; p.mult_h = 2; p.mult_w = 2;
  p.ntaps = 24; p.wtaps = 24; p.cin = 32; p.cin_log2 = 5;
  for (int kd = 0; kd < 3; ++kd)
    for (int r = 0; r < 8; ++r) {
      const int tp = kd * 8 + r;
      p.off_d[tp] = static_cast<signed char>(kd - 1);
      p.off_h[tp] = static_cast<signed char>(r);
      p.off_w[tp] = 0;
      p.wtap[tp] = static_cast<short>(tp);
    }
  p.nout = cout;
  const long long ohw = static_cast<long long>(p.oh) * p.ow;
  p.ov = {ohw * t * cout, ohw * cout, static_cast<long long>(p.ow) * cout, cout};
  p.M = n * t * p.oh * p.ow;
  p.groups = 1;
  p.a = static_cast<const __nv_bfloat16*>(xp);
  p.b = static_cast<const __nv_bfloat16*>(w24);
  p.out = y;
  p.stats = stats;
  p.flags = stats ? EPI_STATS : 0;
  return run_kmajor(p, S(stream), nullptr, 0, 0);
}
size_t qt_stem_wgrad_workspace_bytes(int n, int h, int w, int cout) {
  const size_t generic = wgrad_ws_bytes(256, cout, 1, static_cast<long long>(n) * (h / 2) * (w / 2), nullptr) +
                         static_cast<size_t>(cout) * 256 * sizeof(float);
  const size_t fast = static_cast<size_t>(kNumSMs) * 64 * 224 * sizeof(float);
  return generic > fast ? generic : fast;
}
int qt_stem_wgrad(const void* xp, const void* dy, float* dw, int accumulate, int n, int h, int w, int cout, int cin,
                  void* ws, size_t ws_bytes, qt_stream_t stream) {
  if (stem_wgrad_fast_ok(h, w, cout, cin)) {
    StemWgradParams sp;
    memset(&sp, 0, sizeof(sp));
    sp.xp = static_cast<const __nv_bfloat16*>(xp);
    sp.dy = static_cast<const __nv_bfloat16*>(dy);
    sp.partial = static_cast<float*>(ws);
    sp.N = n; sp.Ho = h / 2; sp.Wo = w / 2; sp.Hp = h + 7; sp.Wp = w + 8;
    sp.strips = (sp.Ho + kSWRows - 1) / kSWRows;
    sp.num_tiles = n * sp.strips;
    sp.row_bytes = sp.Wp * 8;
    sp.x_bytes = ((2 * kSWRows + 5) * sp.row_bytes + 512 + 1023) / 1024 * 1024;
    sp.dy_bytes = sp.Wo * 128;
    if (sp.dy_bytes % 1024) return fail("stem_wgrad: Wo must be a multiple of 8");
    const int grid = sp.num_tiles < kNumSMs ? sp.num_tiles : kNumSMs;
    if (ws_bytes < static_cast<size_t>(grid) * 64 * 224 * sizeof(float)) return fail("stem_wgrad: workspace too small");
    const size_t smem = 1024 + static_cast<size_t>(kSWStages) * (kSWRows * sp.dy_bytes + sp.x_bytes) + sp.dy_bytes + 256;
    if (smem > 227 * 1024) return fail("stem_wgrad: image too wide for the dedicated kernel");
    static size_t configured = 0;
    if (configured < smem) {
      cudaFuncSetAttribute(stem_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
      configured = smem;
    }
    stem_wgrad_kernel<<<grid, kSWThreads, smem, S(stream)>>>(sp);
    if (int rc = cuda_status("stem_wgrad_kernel")) return rc;
    stem_wgrad_reduce_kernel<<<(cout * cin * 49 + 255) / 256, 256, 0, S(stream)>>>(sp.partial, grid, dw, cout, cin, accumulate);
    return cuda_status("stem_wgrad_reduce");
  }
  IgemmParams p;
  if (int rc = fill_stem(p, n, h, w, cout)) return rc;
  const size_t g8_bytes = static_cast<size_t>(cout) * 256 * sizeof(float);
  if (ws_bytes < g8_bytes) return fail("stem_wgrad: workspace too small");
  float* g8 = static_cast<float*>(ws);
  p.a = static_cast<const __nv_bfloat16*>(xp);
  p.b = static_cast<const __nv_bfloat16*>(dy);
  if (int rc = run_wgrad(p, S(stream), static_cast<char*>(ws) + g8_bytes, ws_bytes - g8_bytes, g8, 0, 32, 8)) return rc;
  stem_wgrad_unpack_kernel<<<grid_for(cout * cin * 49, 256), 256, 0, S(stream)>>>(g8, dw, cout, cin, 7, 7, accumulate);
  return cuda_status("stem_wgrad_unpack");
}

// ---- tensor-core linear layers ------------------------------------------------------------------------
static void fill_linear(IgemmParams& p, const void* a, long long lda, const void* b, int rows, int nout, int k) {
  memset(&p, 0, sizeof(p));
  p.a = static_cast<const __nv_bfloat16*>(a);
  p.av = {lda, 0, 0, 0};
  p.id = p.ih = p.iw = 1;
  p.nb = rows; p.od = p.oh = p.ow = 1;
  p.mult_d = p.mult_h = p.mult_w = 1;
  p.ntaps = 1; p.wtaps = 1; p.cin = k; p.cin_log2 = 0;
  p.b = static_cast<const __nv_bfloat16*>(b);
  p.nout = nout;
  p.M = rows;
  p.groups = 1;
}
size_t qt_linear_workspace_bytes(int b, int n, int k) {
  // upper bound over fprop / dgrad split-K and wgrad partials
  const size_t mp = static_cast<size_t>((b + kBM - 1) / kBM) * kBM;
  size_t need = 32 * mp * (static_cast<size_t>((n + 127) / 128) * 128) * sizeof(float);
  const size_t need2 = 32 * mp * (static_cast<size_t>((k + 127) / 128) * 128) * sizeof(float);
  if (need2 > need) need = need2;
  const size_t need3 = wgrad_ws_bytes(k, n, 1, b, nullptr);
  return need3 > need ? need3 : need;
}
int qt_linear_fprop(const void* x, long long ldx, const void* w, const float* bias, void* out, long long ldo,
                    int flags, int b, int n, int k, void* ws, size_t ws_bytes, qt_stream_t stream) {
  if (k % 8 || ldx % 8) return fail("linear_fprop: k and ldx must be multiples of 8");
  IgemmParams p;
  fill_linear(p, x, ldx, w, b, n, k);
  p.out = out;
  p.ov = {ldo, 0, 0, 0};
  p.bias = bias;
  p.flags = flags & (EPI_BIAS | EPI_RELU | EPI_OUT_F32);
  return run_kmajor(p, S(stream), ws, ws_bytes, ldo);
}
int qt_linear_dgrad(const void* dy, long long ldy, const void* wt, void* dx, long long ldx, int b, int n, int k,
                    void* ws, size_t ws_bytes, qt_stream_t stream) {
  if (n % 8 || ldy % 8) return fail("linear_dgrad: n and ldy must be multiples of 8");
  IgemmParams p;
  fill_linear(p, dy, ldy, wt, b, k, n);  // dx[b][k] = dy[b][:] . wt[k][:]
  p.out = dx;
  p.ov = {ldx, 0, 0, 0};
  p.flags = 0;
  return run_kmajor(p, S(stream), ws, ws_bytes, ldx);
}
int qt_linear_wgrad(const void* x, long long ldx, const void* dy, long long ldy, float* dw, int accumulate, int b,
                    int n, int k, void* ws, size_t ws_bytes, qt_stream_t stream) {
  if (k % 8 || n % 8) return fail("linear_wgrad: n and k must be multiples of 8");
  IgemmParams p;
  fill_linear(p, x, ldx, dy, b, n, k);
  p.ov = {ldy, 0, 0, 0};
  return run_wgrad(p, S(stream), ws, ws_bytes, dw, accumulate, k, 1);
}

// ---- BatchNorm ----------------------------------------------------------------------------------------------
namespace {
constexpr int kRedSlices = 64;
int reduce_partials(const float* partial, int rows, int K, double* sums, int* slices, cudaStream_t st) {
  int s = rows < kRedSlices ? rows : kRedSlices;
  if (s < 1) s = 1;
  dim3 grid((K + 31) / 32, s);
  colreduce_stage1_kernel<<<grid, dim3(32, 8), 0, st>>>(partial, rows, K, s, sums);
  *slices = s;
  return cuda_status("colreduce_stage1");
}
int rowlane_block(int c) {  // threads per block for the (C/8 groups) x lanes kernels
  const int groups = c / 8;
  int lanes = 256 / groups;
  if (lanes < 1) lanes = 1;
  return groups * lanes;
}
constexpr int kBwdBlocks = 148 * 2;  // two resident 256-thread blocks per SM: one wave (4 per SM measured 2-3 us slower per call,
                                     // in the reduce tail and in the finalize kernel that sums the per-block rows)
constexpr int kDirectRows = 640;  // partial-row counts a single finalize kernel reduces by itself
}  // namespace

size_t qt_bn_workspace_bytes(int c) {
  // stage-1 sums (double) + per-block partials of the backward reduction + c1/c2
  return static_cast<size_t>(kRedSlices) * 2 * c * sizeof(double) + static_cast<size_t>(kBwdBlocks) * 2 * c * sizeof(float) +
         3 * static_cast<size_t>(c) * sizeof(float) + 256;
}
int qt_bn_stats(const void* y, long long m, int c, float* partial, int partial_rows, qt_stream_t stream) {
  if (c % 8 || c > 2048) return fail("bn_stats: c must be a multiple of 8 and <= 2048");
  const int block = rowlane_block(c);
  const int lanes = block / (c / 8);
  bn_stats_kernel<<<partial_rows, block, static_cast<size_t>(lanes) * 2 * c * sizeof(float), S(stream)>>>(
      static_cast<const __nv_bfloat16*>(y), m, c, partial);
  return cuda_status("bn_stats");
}
static int bn_finalize_impl(const float* partial, int partial_rows, int c, double count, const float* gamma, const float* beta,
                            float eps, float momentum, float* running_mean, float* running_var, long long* nbt, float* mean,
                            float* invstd, float* scale, float* shift, void* ws, size_t ws_bytes, qt_stream_t stream) {
  if (ws_bytes < static_cast<size_t>(kRedSlices) * 2 * c * sizeof(double)) return fail("bn_finalize: workspace too small");
  if (partial_rows <= kDirectRows) {
    bn_finalize_rows_kernel<<<(c + 31) / 32, dim3(32, 32), 0, S(stream)>>>(partial, partial_rows, c, count, gamma, beta, eps,
                                                                          momentum, running_mean, running_var, mean, invstd,
                                                                          scale, shift, nbt);
    return cuda_status("bn_finalize_rows");
  }
  double* sums = static_cast<double*>(ws);
  int slices = 0;
  if (int rc = reduce_partials(partial, partial_rows, 2 * c, sums, &slices, S(stream))) return rc;
  bn_finalize_kernel<<<(c + 127) / 128, 128, 0, S(stream)>>>(sums, slices, c, count, gamma, beta, eps, momentum,
                                                              running_mean, running_var, mean, invstd, scale, shift, nbt);
  return cuda_status("bn_finalize");
}
int qt_bn_finalize(const float* partial, int partial_rows, int c, double count, const float* gamma,
                   const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                   float* mean, float* invstd, float* scale, float* shift, void* ws, size_t ws_bytes,
                   qt_stream_t stream) {
  return bn_finalize_impl(partial, partial_rows, c, count, gamma, beta, eps, momentum, running_mean, running_var, nullptr, mean,
                          invstd, scale, shift, ws, ws_bytes, stream);
}
/* The same, also counting the batch in nn.BatchNorm's `num_batches_tracked` (an int64 device scalar, may be NULL): the
 * increment rides on the finalize launch instead of one ATen kernel per BatchNorm layer per step. */
int qt_bn_finalize_tracked(const float* partial, int partial_rows, int c, double count, const float* gamma, const float* beta,
                           float eps, float momentum, float* running_mean, float* running_var, long long* num_batches_tracked,
                           float* mean, float* invstd, float* scale, float* shift, void* ws, size_t ws_bytes, qt_stream_t stream) {
  return bn_finalize_impl(partial, partial_rows, c, count, gamma, beta, eps, momentum, running_mean, running_var,
                          num_batches_tracked, mean, invstd, scale, shift, ws, ws_bytes, stream);
}
int qt_bn_eval_coeffs(int c, const float* gamma, const float* beta, const float* running_mean,
                      const float* running_var, float eps, float* mean, float* invstd, float* scale, float* shift,
                      qt_stream_t stream) {
  bn_eval_coeffs_kernel<<<(c + 127) / 128, 128, 0, S(stream)>>>(c, gamma, beta, running_mean, running_var, eps, mean, invstd,
                                                                 scale, shift);
  return cuda_status("bn_eval_coeffs");
}
int qt_bn_apply(const void* y, const float* scale, const float* shift, const void* residual, void* out, long long m,
                int c, int relu, qt_stream_t stream) {
  if (c % 8) return fail("bn_apply: c must be a multiple of 8");
  const long long total8 = m * c / 8;
  bn_apply_kernel<<<grid_for(total8, 256), 256, 0, S(stream)>>>(static_cast<const __nv_bfloat16*>(y), scale, shift,
                                                                static_cast<const __nv_bfloat16*>(residual),
                                                                static_cast<__nv_bfloat16*>(out), total8, c, relu);
  return cuda_status("bn_apply");
}
int qt_bn_backward(const void* dout, const void* act, const void* y, const float* mean, const float* invstd,
                   const float* gamma, const float* mask_scale, const float* mask_shift, long long m, int c, float* dgamma,
                   float* dbeta, int accumulate, int eval_mode, void* dy, void* dz_out, void* ws, size_t ws_bytes,
                   qt_stream_t stream) {
  if (c % 8 || c > 2048) return fail("bn_backward: c must be a multiple of 8 and <= 2048");
  if (ws_bytes < qt_bn_workspace_bytes(c)) return fail("bn_backward: workspace too small");
  float* partial = reinterpret_cast<float*>(static_cast<char*>(ws) + static_cast<size_t>(kRedSlices) * 2 * c * sizeof(double));
  float* coef = partial + static_cast<size_t>(kBwdBlocks) * 2 * c;  // [A | B | K]
  const int block = rowlane_block(c);
  const int lanes = block / (c / 8);
  long long want = (m + lanes - 1) / lanes;
  const int blocks = static_cast<int>(want < kBwdBlocks ? (want < 1 ? 1 : want) : kBwdBlocks);
  bn_bwd_reduce_kernel<4><<<blocks, block, static_cast<size_t>(lanes) * 2 * c * sizeof(float), S(stream)>>>(
      static_cast<const __nv_bfloat16*>(dout), static_cast<const __nv_bfloat16*>(act),
      static_cast<const __nv_bfloat16*>(y), mean, invstd, mask_scale, mask_shift, m, c, partial);
  if (int rc = cuda_status("bn_bwd_reduce")) return rc;
  bn_bwd_finalize_rows_kernel<<<(c + 31) / 32, dim3(32, 32), 0, S(stream)>>>(partial, blocks, c, static_cast<double>(m), mean,
                                                                            invstd, gamma, dgamma, dbeta, accumulate, eval_mode,
                                                                            coef);
  if (int rc = cuda_status("bn_bwd_finalize")) return rc;
  const long long total8 = m * c / 8;
  bn_bwd_apply_kernel<<<grid_for(total8, 256), 256, 0, S(stream)>>>(
      static_cast<const __nv_bfloat16*>(dout), static_cast<const __nv_bfloat16*>(act),
      static_cast<const __nv_bfloat16*>(y), coef, mask_scale, mask_shift, static_cast<__nv_bfloat16*>(dy),
      static_cast<__nv_bfloat16*>(dz_out), total8, c);
  return cuda_status("bn_bwd_apply");
}
int qt_bn_relu_maxpool_fwd(const void* y, const float* scale, const float* shift, void* out, void* argmax, void* yarg, int n,
                           int h, int w, int c, qt_stream_t stream) {
  if (c % 8) return fail("bn_relu_maxpool: c must be a multiple of 8");
  const int ho = out_dim(h, 3, 2, 1), wo = out_dim(w, 3, 2, 1);
  const long long total = static_cast<long long>(n) * ho * wo * (c / 8);
  bn_relu_maxpool_fwd_kernel<<<grid_for(total, 256), 256, 0, S(stream)>>>(static_cast<const __nv_bfloat16*>(y), scale, shift,
                                                                          static_cast<__nv_bfloat16*>(out),
                                                                          static_cast<signed char*>(argmax),
                                                                          static_cast<__nv_bfloat16*>(yarg), n, h, w, c, ho, wo);
  return cuda_status("bn_relu_maxpool_fwd");
}
int qt_bn_relu_maxpool_bwd(const void* dpool, const void* argmax, const void* y, const void* yarg, const float* scale,
                           const float* shift, const float* mean, const float* invstd, const float* gamma, int n, int h, int w, int c,
                           float* dgamma, float* dbeta, int eval_mode, void* dy, void* ws, size_t ws_bytes,
                           qt_stream_t stream) {
  if (c % 8 || c > 2048) return fail("bn_relu_maxpool_bwd: c must be a multiple of 8 and <= 2048");
  if (ws_bytes < qt_bn_workspace_bytes(c)) return fail("bn_relu_maxpool_bwd: workspace too small");
  if ((h | w) & 1) return fail("bn_relu_maxpool_bwd: h and w must be even (2x2 block decomposition)");
  const int ho = out_dim(h, 3, 2, 1), wo = out_dim(w, 3, 2, 1);
  float* partial = reinterpret_cast<float*>(static_cast<char*>(ws) + static_cast<size_t>(kRedSlices) * 2 * c * sizeof(double));
  float* coef = partial + static_cast<size_t>(kBwdBlocks) * 2 * c;
  const int block = rowlane_block(c);
  const int lanes = block / (c / 8);
  const long long m = static_cast<long long>(n) * h * w;
  const long long quads = static_cast<long long>(n) * ho * wo;
  long long want = (quads + lanes - 1) / lanes;
  const int blocks = static_cast<int>(want < kBwdBlocks ? (want < 1 ? 1 : want) : kBwdBlocks);
  if (yarg)
    // dz is non-zero only at arg-max positions and the sums are linear in the windows: statistics from the pooled-size tensors
    // (pooled gradient + conv output at each window's arg-max) instead of a pass over the full 112 x 112 map
    bn_bwd_reduce_kernel<4><<<blocks, block, static_cast<size_t>(lanes) * 2 * c * sizeof(float), S(stream)>>>(
        static_cast<const __nv_bfloat16*>(dpool), nullptr, static_cast<const __nv_bfloat16*>(yarg), mean, invstd, scale, shift, quads, c,
        partial);
  else
    stem_bn_pool_bwd_reduce_kernel<<<blocks, block, static_cast<size_t>(lanes) * 2 * c * sizeof(float), S(stream)>>>(
        static_cast<const __nv_bfloat16*>(dpool), static_cast<const signed char*>(argmax), static_cast<const __nv_bfloat16*>(y), scale,
        shift, mean, invstd, n, h, w, c, ho, wo, partial);
  if (int rc = cuda_status("stem_bn_pool_bwd_reduce")) return rc;
  bn_bwd_finalize_rows_kernel<<<(c + 31) / 32, dim3(32, 32), 0, S(stream)>>>(partial, blocks, c, static_cast<double>(m), mean,
                                                                             invstd, gamma, dgamma, dbeta, 0, eval_mode, coef);
  if (int rc = cuda_status("bn_bwd_finalize")) return rc;
  const long long total = quads * (c / 8);
  stem_bn_pool_bwd_apply_kernel<<<grid_for(total, 256), 256, 0, S(stream)>>>(
      static_cast<const __nv_bfloat16*>(dpool), static_cast<const signed char*>(argmax), static_cast<const __nv_bfloat16*>(y), scale,
      shift, coef, static_cast<__nv_bfloat16*>(dy), n, h, w, c, ho, wo);
  return cuda_status("stem_bn_pool_bwd_apply");
}
int qt_relu_backward(const void* dout, const void* act, void* dz, long long n, qt_stream_t stream) {
  if (n % 8) return fail("relu_backward: n must be a multiple of 8");
  relu_bwd_kernel<<<grid_for(n / 8, 256), 256, 0, S(stream)>>>(static_cast<const __nv_bfloat16*>(dout),
                                                               static_cast<const __nv_bfloat16*>(act),
                                                               static_cast<__nv_bfloat16*>(dz), n / 8);
  return cuda_status("relu_backward");
}
int qt_colsum(const void* x, long long m, int c, float* out, int accumulate, void* ws, size_t ws_bytes,
              qt_stream_t stream) {
  if (c % 8) return fail("colsum: c must be a multiple of 8");
  if (ws_bytes < qt_bn_workspace_bytes(c)) return fail("colsum: workspace too small");
  double* sums = static_cast<double*>(ws);
  float* partial = reinterpret_cast<float*>(static_cast<char*>(ws) + static_cast<size_t>(kRedSlices) * 2 * c * sizeof(double));
  // channel chunks of at most 1024 channels (128 groups) -> at least 2 row lanes per block
  int chunks = (c + 1023) / 1024;
  int cc = ((c / 8 + chunks - 1) / chunks) * 8;
  chunks = (c + cc - 1) / cc;
  const int groups_max = cc / 8;
  int lanes = 256 / groups_max;
  if (lanes < 1) lanes = 1;
  // block size must be a multiple of every chunk's group count: use the widest chunk and let the (narrower)
  // last chunk recompute its own lanes from blockDim
  const int block = groups_max * lanes;
  long long want = (m + lanes - 1) / lanes;
  const int blocks = static_cast<int>(want < kBwdBlocks ? (want < 1 ? 1 : want) : kBwdBlocks);
  dim3 grid(blocks, chunks);
  colsum_bf16_kernel<<<grid, block, static_cast<size_t>(block / 1) * 8 * sizeof(float) + 64, S(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), m, c, cc, partial);
  if (int rc = cuda_status("colsum")) return rc;
  int slices = 0;
  if (int rc = reduce_partials(partial, blocks, c, sums, &slices, S(stream))) return rc;
  colreduce_final_f32_kernel<<<(c + 127) / 128, 128, 0, S(stream)>>>(sums, slices, c, out, accumulate);
  return cuda_status("colreduce_final");
}
int qt_add_bf16(const void* a, const void* b, void* out, long long n, qt_stream_t stream) {
  if (n % 8) return fail("add_bf16: n must be a multiple of 8");
  add_bf16_kernel<<<grid_for(n / 8, 256), 256, 0, S(stream)>>>(static_cast<const __nv_bfloat16*>(a),
                                                               static_cast<const __nv_bfloat16*>(b),
                                                               static_cast<__nv_bfloat16*>(out), n / 8);
  return cuda_status("add_bf16");
}

// ---- pooling ------------------------------------------------------------------------------------------------
int qt_maxpool2d_fwd(const void* x, void* out, void* argmax, int n, int h, int w, int c, int ksize, int stride,
                     int pad, qt_stream_t stream) {
  if (c % 8) return fail("maxpool2d: c must be a multiple of 8");
  const int ho = out_dim(h, ksize, stride, pad), wo = out_dim(w, ksize, stride, pad);
  const long long total = static_cast<long long>(n) * ho * wo * (c / 8);
  maxpool2d_fwd_kernel<<<grid_for(total, 256), 256, 0, S(stream)>>>(static_cast<const __nv_bfloat16*>(x),
                                                                    static_cast<__nv_bfloat16*>(out),
                                                                    static_cast<signed char*>(argmax), n, h, w, c, ho, wo,
                                                                    ksize, stride, pad);
  return cuda_status("maxpool2d_fwd");
}
int qt_maxpool2d_bwd(const void* dout, const void* argmax, void* dx, int n, int h, int w, int c, int ksize,
                     int stride, int pad, qt_stream_t stream) {
  if (c % 8) return fail("maxpool2d: c must be a multiple of 8");
  const int ho = out_dim(h, ksize, stride, pad), wo = out_dim(w, ksize, stride, pad);
  const long long total = static_cast<long long>(n) * h * w * (c / 8);
  maxpool2d_bwd_kernel<<<grid_for(total, 256), 256, 0, S(stream)>>>(static_cast<const __nv_bfloat16*>(dout),
                                                                    static_cast<const signed char*>(argmax),
                                                                    static_cast<__nv_bfloat16*>(dx), n, h, w, c, ho, wo,
                                                                    ksize, stride, pad);
  return cuda_status("maxpool2d_bwd");
}
namespace {
bool quadtree_fast_ok(int qh, int qw, int cq, int cg, int ldf) {
  const int pp = (qh / 2) * (qw / 2);
  return cq % 8 == 0 && cg % 64 == 0 && ldf % 8 == 0 && (cq * pp) % 8 == 0 && pp > 0 &&
         static_cast<size_t>(qh) * qw * cq * 2 <= 48 * 1024;
}
}  // namespace
int qt_quadtree_pool_fwd(const void* q, const void* l4, void* feat, int b, int qh, int qw, int cq, int ghw, int cg,
                         int ldf, qt_stream_t stream) {
  if (b < 1) return 0;
  if (quadtree_fast_ok(qh, qw, cq, cg, ldf)) {
    quadtree_pool_fwd_kernel<<<dim3(b, 5), kQtThreads, static_cast<size_t>(qh) * qw * cq * 2, S(stream)>>>(
        static_cast<const __nv_bfloat16*>(q), static_cast<const __nv_bfloat16*>(l4), static_cast<__nv_bfloat16*>(feat), b, qh, qw,
        cq, ghw, cg, ldf);
    return cuda_status("quadtree_pool_fwd");
  }
  const long long total = static_cast<long long>(b) * 4 * cq + static_cast<long long>(b) * cg;
  quadtree_pool_fwd_generic_kernel<<<grid_for(total, 128), 128, 0, S(stream)>>>(
      static_cast<const __nv_bfloat16*>(q), static_cast<const __nv_bfloat16*>(l4), static_cast<__nv_bfloat16*>(feat), b, qh, qw, cq,
      ghw, cg, ldf);
  return cuda_status("quadtree_pool_fwd");
}
int qt_quadtree_pool_bwd(const void* dfeat, const void* q, void* dq, void* dl4, int b, int qh, int qw, int cq,
                         int ghw, int cg, int ldf, qt_stream_t stream) {
  if (b < 1) return 0;
  if (quadtree_fast_ok(qh, qw, cq, cg, ldf)) {
    quadtree_pool_bwd_kernel<<<dim3(b, 5), kQtThreads, static_cast<size_t>(cq) * (qh / 2) * (qw / 2) * 2, S(stream)>>>(
        static_cast<const __nv_bfloat16*>(dfeat), static_cast<const __nv_bfloat16*>(q), static_cast<__nv_bfloat16*>(dq),
        static_cast<__nv_bfloat16*>(dl4), b, qh, qw, cq, ghw, cg, ldf);
    return cuda_status("quadtree_pool_bwd");
  }
  const long long total = static_cast<long long>(b) * 4 * qh * qw * cq + static_cast<long long>(b) * ghw * cg;
  quadtree_pool_bwd_generic_kernel<<<grid_for(total, 256), 256, 0, S(stream)>>>(
      static_cast<const __nv_bfloat16*>(dfeat), static_cast<const __nv_bfloat16*>(q), static_cast<__nv_bfloat16*>(dq),
      static_cast<__nv_bfloat16*>(dl4), b, qh, qw, cq, ghw, cg, ldf);
  return cuda_status("quadtree_pool_bwd");
}
int qt_region_avgpool_fwd(const void* x, void* out, long long regions, int p, int c, long long ldo,
                          qt_stream_t stream) {
  region_avgpool_fwd_kernel<<<grid_for(regions * c, 128), 128, 0, S(stream)>>>(static_cast<const __nv_bfloat16*>(x),
                                                                               static_cast<__nv_bfloat16*>(out), regions, p,
                                                                               c, ldo);
  return cuda_status("region_avgpool_fwd");
}
int qt_region_avgpool_bwd(const void* dout, const void* x, void* dx, long long regions, int p, int c, long long ldo,
                          int relu_mask, qt_stream_t stream) {
  region_avgpool_bwd_kernel<<<grid_for(regions * p * c, 256), 256, 0, S(stream)>>>(
      static_cast<const __nv_bfloat16*>(dout), static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(dx), regions,
      p, c, ldo, relu_mask);
  return cuda_status("region_avgpool_bwd");
}

int qt_maxpool3d_fwd(const void* x, void* out, void* argmax, int n, int d, int h, int w, int c, int kd, int kh, int kw,
                     qt_stream_t stream) {
  if (c % 8) return fail("maxpool3d: c must be a multiple of 8");
  const long long total = static_cast<long long>(n) * (d / kd) * (h / kh) * (w / kw) * (c / 8);
  maxpool3d_fwd_kernel<<<grid_for(total, 256), 256, 0, S(stream)>>>(static_cast<const __nv_bfloat16*>(x),
                                                                    static_cast<__nv_bfloat16*>(out),
                                                                    static_cast<signed char*>(argmax), n, d, h, w, c, kd, kh, kw);
  return cuda_status("maxpool3d_fwd");
}
int qt_maxpool3d_bwd(const void* dout, const void* argmax, void* dx, int n, int d, int h, int w, int c, int kd, int kh,
                     int kw, qt_stream_t stream) {
  if (c % 8) return fail("maxpool3d: c must be a multiple of 8");
  const long long total = static_cast<long long>(n) * d * h * w * (c / 8);
  maxpool3d_bwd_kernel<<<grid_for(total, 256), 256, 0, S(stream)>>>(static_cast<const __nv_bfloat16*>(dout),
                                                                    static_cast<const signed char*>(argmax),
                                                                    static_cast<__nv_bfloat16*>(dx), n, d, h, w, c, kd, kh, kw);
  return cuda_status("maxpool3d_bwd");
}
static int pool3d_fused_check(int n, int d, int h, int w, int c, int kd, int kh, int kw) {
  if (c % 8 || c > 2048) return fail("bn_relu_maxpool3d: c must be a multiple of 8 and <= 2048");
  if ((kd != 1 && kd != 2) || kh != 2 || kw != 2) return fail("bn_relu_maxpool3d: pool must be (1|2, 2, 2)");
  if (d % kd || (h | w) & 1) return fail("bn_relu_maxpool3d: d must divide by kd, h and w must be even");
  if (static_cast<long long>(n) * d * h * w * (c / 8) >= (1ll << 31)) return fail("bn_relu_maxpool3d: tensor too large for 32-bit indexing");
  return 0;
}
int qt_bn_relu_maxpool3d_fwd(const void* y, const float* scale, const float* shift, void* out, void* yarg, void* argmax, int n,
                             int d, int h, int w, int c, int kd, int kh, int kw, qt_stream_t stream) {
  if (int rc = pool3d_fused_check(n, d, h, w, c, kd, kh, kw)) return rc;
  if ((yarg == nullptr) != (argmax == nullptr)) return fail("bn_relu_maxpool3d_fwd: yarg and argmax go together");
  const long long total = static_cast<long long>(n) * (d / kd) * (h / 2) * (w / 2) * (c / 8);
  const int grid = grid_for(total, 256);
  auto* yy = static_cast<const __nv_bfloat16*>(y);
  auto* oo = static_cast<__nv_bfloat16*>(out);
  auto* ya = static_cast<__nv_bfloat16*>(yarg);
  auto* am = static_cast<signed char*>(argmax);
  if (kd == 1) bn_relu_maxpool3d_fwd_kernel<1><<<grid, 256, 0, S(stream)>>>(yy, scale, shift, oo, ya, am, n, d, h, w, c);
  else bn_relu_maxpool3d_fwd_kernel<2><<<grid, 256, 0, S(stream)>>>(yy, scale, shift, oo, ya, am, n, d, h, w, c);
  return cuda_status("bn_relu_maxpool3d_fwd");
}
int qt_bn_relu_maxpool3d_bwd(const void* dpool, const void* argmax, const void* y, const void* yarg, const float* scale,
                             const float* shift, const float* mean, const float* invstd, const float* gamma, int n, int d, int h,
                             int w, int c, int kd, int kh, int kw, float* dgamma, float* dbeta, float* dbias, int eval_mode,
                             void* dy, void* ws, size_t ws_bytes, qt_stream_t stream) {
  if (int rc = pool3d_fused_check(n, d, h, w, c, kd, kh, kw)) return rc;
  if (ws_bytes < qt_bn_workspace_bytes(c)) return fail("bn_relu_maxpool3d_bwd: workspace too small");
  float* partial = reinterpret_cast<float*>(static_cast<char*>(ws) + static_cast<size_t>(kRedSlices) * 2 * c * sizeof(double));
  float* coef = partial + static_cast<size_t>(kBwdBlocks) * 2 * c;
  const int block = rowlane_block(c);
  const int lanes = block / (c / 8);
  const long long m = static_cast<long long>(n) * d * h * w;
  const long long mp = static_cast<long long>(n) * (d / kd) * (h / 2) * (w / 2);
  long long want = (mp + lanes - 1) / lanes;
  const int blocks = static_cast<int>(want < kBwdBlocks ? (want < 1 ? 1 : want) : kBwdBlocks);
  // statistics from pooled-size tensors: every non-arg-max position has dz = 0
  bn_bwd_reduce_kernel<4><<<blocks, block, static_cast<size_t>(lanes) * 2 * c * sizeof(float), S(stream)>>>(
      static_cast<const __nv_bfloat16*>(dpool), nullptr, static_cast<const __nv_bfloat16*>(yarg), mean, invstd, scale, shift, mp, c,
      partial);
  if (int rc = cuda_status("bn_pool3d_bwd_reduce")) return rc;
  bn_bwd_finalize_rows_kernel<<<(c + 31) / 32, dim3(32, 32), 0, S(stream)>>>(partial, blocks, c, static_cast<double>(m), mean,
                                                                             invstd, gamma, dgamma, dbeta, 0, eval_mode, coef, dbias);
  if (int rc = cuda_status("bn_bwd_finalize")) return rc;
  const int grid = grid_for(mp * (c / 8), 256);
  auto* dp = static_cast<const __nv_bfloat16*>(dpool);
  auto* am = static_cast<const signed char*>(argmax);
  auto* yy = static_cast<const __nv_bfloat16*>(y);
  auto* dd = static_cast<__nv_bfloat16*>(dy);
  if (kd == 1) bn_pool3d_bwd_apply_kernel<1><<<grid, 256, 0, S(stream)>>>(dp, am, yy, coef, scale, shift, dd, n, d, h, w, c);
  else bn_pool3d_bwd_apply_kernel<2><<<grid, 256, 0, S(stream)>>>(dp, am, yy, coef, scale, shift, dd, n, d, h, w, c);
  return cuda_status("bn_pool3d_bwd_apply");
}
int qt_attn_pool_fwd(const float* x, const float* scores, float* wts, float* out, int b, int r, int c, qt_stream_t stream) {
  attn_pool_fwd_kernel<<<b, 64, 0, S(stream)>>>(x, scores, wts, out, r, c);
  return cuda_status("attn_pool_fwd");
}
int qt_attn_pool_bwd(const float* x, const float* wts, const float* dout, float* dx, float* dscores, int b, int r, int c,
                     qt_stream_t stream) {
  attn_pool_bwd_kernel<<<b, 128, r * sizeof(float), S(stream)>>>(x, wts, dout, dx, dscores, r, c);
  return cuda_status("attn_pool_bwd");
}

// ---- small linears ------------------------------------------------------------------------------------------
int qt_small_linear_fwd(const void* x, int x_is_bf16, long long ldx, const float* w, const float* bias, int b, int n,
                        int k, int relu, float drop_p, unsigned long long seed, float* out, long long ldo,
                        void* out16, long long ldo16, qt_stream_t stream) {
  if (sgemm_worthwhile(b, n, k)) {
    SgemmParams p{};
    p.a = x; p.lda = ldx; p.b = w; p.ldb = k; p.M = b; p.N = n; p.K = k;
    p.bias = bias; p.relu = relu; p.drop_p = drop_p; p.seed = seed;
    p.out = out; p.ldo = ldo; p.out16 = static_cast<__nv_bfloat16*>(out16); p.ldo16 = ldo16;
    sgemm_launch<1, 1, 0>(p, x_is_bf16 != 0, false, S(stream));
    return cuda_status("small_linear_fwd(tiled)");
  }
  const long long warps = static_cast<long long>(b) * n;
  const int grid = static_cast<int>((warps * 32 + 255) / 256);
  if (x_is_bf16)
    small_linear_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, S(stream)>>>(static_cast<const __nv_bfloat16*>(x), ldx, w, bias, b, n,
                                                                       k, relu, drop_p, seed, out, ldo,
                                                                       static_cast<__nv_bfloat16*>(out16), ldo16);
  else
    small_linear_fwd_kernel<float><<<grid, 256, 0, S(stream)>>>(static_cast<const float*>(x), ldx, w, bias, b, n, k, relu,
                                                               drop_p, seed, out, ldo, static_cast<__nv_bfloat16*>(out16),
                                                               ldo16);
  return cuda_status("small_linear_fwd");
}
int qt_small_linear_bwd_dx(const void* dy, int dy_is_bf16, long long ldy, const float* w, int b, int n, int k,
                           const float* act, long long lda, float drop_p, unsigned long long seed, float* dx,
                           long long ldx, void* dx16, long long ldx16, qt_stream_t stream) {
  if (sgemm_worthwhile(b, k, n)) {
    SgemmParams p{};
    p.a = dy; p.lda = ldy; p.b = w; p.ldb = k; p.M = b; p.N = k; p.K = n;
    p.drop_p = drop_p; p.seed = seed; p.act = act; p.ldact = lda;
    p.out = dx; p.ldo = ldx; p.out16 = static_cast<__nv_bfloat16*>(dx16); p.ldo16 = ldx16;
    sgemm_launch<1, 0, 1>(p, dy_is_bf16 != 0, false, S(stream));
    return cuda_status("small_linear_bwd_dx(tiled)");
  }
  const long long total = static_cast<long long>(b) * k;
  const int grid = static_cast<int>((total + 255) / 256);
  if (dy_is_bf16)
    small_linear_bwd_dx_kernel<__nv_bfloat16><<<grid, 256, 0, S(stream)>>>(static_cast<const __nv_bfloat16*>(dy), ldy, w, b, n, k,
                                                                          act, lda, drop_p, seed, dx, ldx,
                                                                          static_cast<__nv_bfloat16*>(dx16), ldx16);
  else
    small_linear_bwd_dx_kernel<float><<<grid, 256, 0, S(stream)>>>(static_cast<const float*>(dy), ldy, w, b, n, k, act, lda,
                                                                  drop_p, seed, dx, ldx, static_cast<__nv_bfloat16*>(dx16),
                                                                  ldx16);
  return cuda_status("small_linear_bwd_dx");
}
int qt_small_linear_bwd_dw(const void* dy, int dy_is_bf16, long long ldy, const void* x, int x_is_bf16,
                           long long ldx, int b, int n, int k, float* dw, float* db, int accumulate,
                           qt_stream_t stream) {
  if (sgemm_worthwhile(n, k, b)) {
    SgemmParams p{};
    p.a = dy; p.lda = ldy; p.b = x; p.ldb = ldx; p.M = n; p.N = k; p.K = b;
    p.out = dw; p.ldo = k; p.db = db; p.accumulate = accumulate;
    sgemm_launch<0, 0, 2>(p, dy_is_bf16 != 0, x_is_bf16 != 0, S(stream));
    return cuda_status("small_linear_bwd_dw(tiled)");
  }
  const dim3 grid((k + 31) / 32, n);
  const dim3 blk(32, 8);
  cudaStream_t st = S(stream);
  if (dy_is_bf16 && x_is_bf16)
    small_linear_bwd_dw_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, blk, 0, st>>>(
        static_cast<const __nv_bfloat16*>(dy), ldy, static_cast<const __nv_bfloat16*>(x), ldx, b, n, k, dw, db, accumulate);
  else if (dy_is_bf16)
    small_linear_bwd_dw_kernel<__nv_bfloat16, float><<<grid, blk, 0, st>>>(static_cast<const __nv_bfloat16*>(dy), ldy,
                                                                          static_cast<const float*>(x), ldx, b, n, k, dw, db,
                                                                          accumulate);
  else if (x_is_bf16)
    small_linear_bwd_dw_kernel<float, __nv_bfloat16><<<grid, blk, 0, st>>>(static_cast<const float*>(dy), ldy,
                                                                          static_cast<const __nv_bfloat16*>(x), ldx, b, n, k,
                                                                          dw, db, accumulate);
  else
    small_linear_bwd_dw_kernel<float, float><<<grid, blk, 0, st>>>(static_cast<const float*>(dy), ldy,
                                                                  static_cast<const float*>(x), ldx, b, n, k, dw, db,
                                                                  accumulate);
  return cuda_status("small_linear_bwd_dw");
}
int qt_relu_dropout(float* h, void* h16, long long n, float drop_p, unsigned long long seed, int relu,
                    qt_stream_t stream) {
  relu_dropout_kernel<<<grid_for(n, 256), 256, 0, S(stream)>>>(h, static_cast<__nv_bfloat16*>(h16), n, drop_p, seed, relu);
  return cuda_status("relu_dropout");
}

int qt_relu_dropout_bwd(const float* dout, const float* act, float* dz, void* dz16, long long n, float drop_p,
                        unsigned long long seed, int relu, qt_stream_t stream) {
  relu_dropout_bwd_kernel<<<grid_for(n, 256), 256, 0, S(stream)>>>(dout, act, dz, static_cast<__nv_bfloat16*>(dz16), n, drop_p, seed,
                                                                  relu);
  return cuda_status("relu_dropout_bwd");
}


// ---- loss + fused classifier tail -------------------------------------------------------------------------------
int qt_cross_entropy(const float* logits, long long ld, const long long* labels, int b, int nc, float grad_scale,
                     const float* upstream, float* loss_rows, float* loss_mean, float* dlogits, unsigned int* counter,
                     qt_stream_t stream) {
  if (b < 1 || nc < 1 || nc > kMaxClasses) return fail("cross_entropy: 1 <= classes <= %d (got %d)", kMaxClasses, nc);
  if (!loss_rows || !loss_mean || !counter) return fail("cross_entropy: loss_rows / loss_mean / counter are required");
  cross_entropy_kernel<<<(b + 7) / 8, 256, 0, S(stream)>>>(logits, ld, labels, b, nc, grad_scale, upstream, loss_rows, loss_mean, dlogits, counter);
  return cuda_status("cross_entropy");
}
int qt_head_tail_fwd(float* h, void* h16, int nhid, const float* w3, const float* b3, int nc, const long long* labels, int b,
                     float drop_p, unsigned long long seed, float* logits, float* loss_rows, float* loss_mean,
                     unsigned int* counter, qt_stream_t stream) {
  if (b < 1) return 0;
  if (nc < 1 || nc > kMaxClasses) return fail("head_tail: 1 <= classes <= %d (got %d)", kMaxClasses, nc);
  if (nhid < 1 || nhid > kHeadThreads * kHeadCols) return fail("head_tail: hidden width must be <= %d (got %d)", kHeadThreads * kHeadCols, nhid);
  if (labels && (!loss_rows || !loss_mean || !counter)) return fail("head_tail: labels need loss_rows / loss_mean / counter");
  head_tail_fwd_kernel<<<b, kHeadThreads, 0, S(stream)>>>(h, static_cast<__nv_bfloat16*>(h16), nhid, w3, b3, nc, labels, b, drop_p,
                                                          seed, logits, loss_rows, loss_mean, counter);
  return cuda_status("head_tail_fwd");
}
int qt_head_tail_bwd(const float* act, int nhid, const float* w3, int nc, const float* logits, const long long* labels,
                     float grad_scale, const float* upstream, float drop_p, unsigned long long seed, float* dlogits, void* dh16,
                     int b, qt_stream_t stream) {
  if (b < 1) return 0;
  if (nc < 1 || nc > kMaxClasses) return fail("head_tail: 1 <= classes <= %d (got %d)", kMaxClasses, nc);
  if (!dlogits || !dh16) return fail("head_tail_bwd: dlogits and dh16 are required");
  if (labels && !logits) return fail("head_tail_bwd: labels need the stored logits");
  head_tail_bwd_kernel<<<b, kHeadThreads, 0, S(stream)>>>(act, nhid, w3, nc, logits, labels, grad_scale, upstream, drop_p, seed,
                                                          dlogits, static_cast<__nv_bfloat16*>(dh16));
  return cuda_status("head_tail_bwd");
}

// ---- optimizer ----------------------------------------------------------------------------------------------------
int qt_adam_item_plan(qt_adam_item* item) {
  static_assert(sizeof(qt_adam_item) == sizeof(AdamItem), "qt_adam_item layout");
  static_assert(sizeof(qt_adam_group) == sizeof(AdamGroup), "qt_adam_group layout");
  if (!item || item->n < 1) return fail("adam_item_plan: bad item");
  if (item->wf) {
    if (item->cout < 1 || item->cin < 1 || item->taps < 1 || item->taps > 32 ||
        static_cast<long long>(item->cout) * item->cin * item->taps != item->n)
      return fail("adam_item_plan: packed item needs cout*cin*taps == n and taps <= 32");
    item->co_tile = wpack_co_tile(item->taps);
    item->ci_tiles = (item->cin + 31) / 32;
    return item->ci_tiles * ((item->cout + item->co_tile - 1) / item->co_tile);
  }
  item->co_tile = item->ci_tiles = 0;
  const long long blocks = (item->n + kAdamElemsPerBlock - 1) / kAdamElemsPerBlock;
  if (blocks > 0x7fffffff) return fail("adam_item_plan: tensor too large");
  return static_cast<int>(blocks);
}
int qt_adam_multi(const void* items_dev, int nitems, int total_blocks, int max_taps, const qt_adam_group* groups, int ngroups,
                  const float* clip_coef, float grad_scale, qt_stream_t stream) {
  if (nitems < 1 || total_blocks < 1) return 0;
  if (ngroups < 1 || ngroups > kAdamMaxGroups) return fail("adam_multi: 1..%d parameter groups", kAdamMaxGroups);
  if (max_taps < 1 || max_taps > 32) return fail("adam_multi: taps must be 1..32");
  AdamGroups g;
  memset(&g, 0, sizeof(g));
  memcpy(g.g, groups, sizeof(AdamGroup) * ngroups);
  g.grad_scale = grad_scale;
  size_t smem = wpack_smem(1);
  for (int t = 2; t <= max_taps; ++t) smem = wpack_smem(t) > smem ? wpack_smem(t) : smem;
  adam_multi_kernel<<<total_blocks, 256, smem, S(stream)>>>(static_cast<const AdamItem*>(items_dev), nitems, g, clip_coef);
  return cuda_status("adam_multi");
}
int qt_grad_norm_blocks(long long n) { return static_cast<int>((n + kAdamElemsPerBlock - 1) / kAdamElemsPerBlock); }
int qt_grad_clip_coef(const void* items_dev, int nitems, int total_blocks, float max_norm, float grad_scale, float* partial,
                      float* total_norm, float* coef, qt_stream_t stream) {
  static_assert(sizeof(qt_norm_item) == sizeof(NormItem), "qt_norm_item layout");
  if (nitems < 1 || total_blocks < 1) return fail("grad_clip_coef: no gradients");
  grad_sqnorm_multi_kernel<<<total_blocks, 256, 0, S(stream)>>>(static_cast<const NormItem*>(items_dev), nitems, partial);
  if (int rc = cuda_status("grad_sqnorm_multi")) return rc;
  grad_clip_coef_kernel<<<1, 256, 0, S(stream)>>>(partial, total_blocks, max_norm, grad_scale, total_norm, coef);
  return cuda_status("grad_clip_coef");
}


// ---- LSTM (numeric-sequence branches) -------------------------------------------------------------------------------
int qt_transpose_f32(const float* in, float* out, int rows, int cols, qt_stream_t stream) {
  if (rows < 1 || cols < 1) return fail("transpose: empty matrix");
  transpose_f32_kernel<<<dim3((cols + 31) / 32, (rows + 31) / 32), dim3(32, 8), 0, S(stream)>>>(in, out, rows, cols);
  return cuda_status("transpose_f32");
}
int qt_lstm_layer_fwd(const float* xproj, const float* whh_t, const float* bhh, int b, int t, int h, float* hseq, float* hprev,
                      float* cseq, float* gates, qt_stream_t stream) {
  if (b < 1 || t < 1) return 0;
  if (h < 1 || h > 256) return fail("lstm: hidden size must be <= 256 (got %d)", h);
  const int hs = 2 * ((h + 2 * kLstmCluster - 1) / (2 * kLstmCluster));  // units per CTA, even (row / unit pairs per thread)
  const size_t smem = sizeof(float) * lstm_fwd_smem_floats(h, hs);
  static size_t configured = 0;
  if (configured < smem) {
    cudaFuncSetAttribute(lstm_layer_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    configured = smem;
  }
  if (int rc = launch_cluster8(lstm_layer_fwd_kernel, (b + kLstmBT - 1) / kLstmBT, smem, S(stream), xproj, whh_t, bhh, b, t, h, hs, hseq,
                               hprev, cseq, gates))
    return rc;
  return cuda_status("lstm_layer_fwd");
}
int qt_lstm_layer_bwd(const float* dhseq, float out_drop_p, unsigned long long seed, const float* whh, const float* gates,
                      const float* cseq, int b, int t, int h, float* dgates, qt_stream_t stream) {
  if (b < 1 || t < 1) return 0;
  if (h < 1 || h > 256) return fail("lstm: hidden size must be <= 256 (got %d)", h);
  const int hs = 2 * ((h + 2 * kLstmCluster - 1) / (2 * kLstmCluster));
  const size_t smem = sizeof(float) * lstm_bwd_smem_floats(h, hs);
  static size_t configured = 0;
  if (configured < smem) {
    cudaFuncSetAttribute(lstm_layer_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    configured = smem;
  }
  if (int rc = launch_cluster8(lstm_layer_bwd_kernel, (b + kLstmBT - 1) / kLstmBT, smem, S(stream), dhseq, out_drop_p, seed, whh, gates,
                               cseq, b, t, h, hs, dgates))
    return rc;
  return cuda_status("lstm_layer_bwd");
}

}  // extern "C"
