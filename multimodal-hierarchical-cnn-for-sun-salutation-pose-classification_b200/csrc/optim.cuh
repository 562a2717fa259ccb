// Multi-tensor Adam for the training loops of the reference scripts (`optim.Adam(model.parameters(), lr, weight_decay)`,
// Quadtree_from scratch/Quadtree_train.py:45; 3dcnn/train_3D_Quadtree_cnn_model.py:88 with
// `clip_grad_norm_(model.parameters(), 1.0)` at :123): torch.optim.Adam semantics (L2 weight decay folded into the
// gradient, bias-corrected moments, no amsgrad), every parameter of the model in ONE launch, and — for the weights
// that feed the tensor-core GEMMs — the refreshed bf16 operand copies wf[cout][taps][cin] / wd[cin][taps][cout]
// written by the same pass (what a separate qt_wpack_multi launch did after torch's optimizer).
#pragma once
#include "elementwise.cuh"

namespace qt {

struct AdamGroup {  // one torch param_group; step-dependent factors are folded on the host
  float step_size;      // lr / (1 - beta1^t)
  float beta1, beta2;
  float eps;
  float weight_decay;
  float inv_bc2_sqrt;   // 1 / sqrt(1 - beta2^t)
  float omb1, omb2;     // 1 - beta1, 1 - beta2 rounded from double (torch passes them as double scalars; 1.f - 0.999f is 5e-5 off)
};
constexpr int kAdamMaxGroups = 8;
struct AdamGroups {
  AdamGroup g[kAdamMaxGroups];
  float grad_scale;  // every gradient is multiplied by this first (1/world after a SUM all-reduce; 1 otherwise)
};
struct AdamItem {
  float* p;
  const float* g;
  float* m;
  float* v;
  __nv_bfloat16* wf;  // NULL: plain elementwise item
  __nv_bfloat16* wd;  // may be NULL
  long long n;
  int cout, cin, taps;
  int co_tile, ci_tiles, first_block;
  int group, pad;
};
constexpr int kAdamElemsPerBlock = 256 * 8;

__device__ __forceinline__ float adam_one(float p, float g, float& m, float& v, const AdamGroup& G, float clip) {
  g = g * clip;
  g = fmaf(G.weight_decay, p, g);               // grad + wd * param (torch: grad.add(param, alpha=weight_decay))
  m = m + G.omb1 * (g - m);                     // exp_avg.lerp_(grad, 1 - beta1)
  v = v * G.beta2 + G.omb2 * (g * g);           // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
  const float denom = sqrtf(v) * G.inv_bc2_sqrt + G.eps;
  return p - G.step_size * (m / denom);
}

// Tile variant of wpack_tile_t: the load of w applies the Adam update first. TC > 0: taps known at compile time (the index
// divisions become multiplies). Full, 16-byte aligned tiles stream p / g / m / v as float4 and write the bf16 copies as pairs:
// with scalar accesses and runtime divisions the kernel was instruction bound (0.31 ms for 772 MB at B = 256).
template <int TC>
__device__ __forceinline__ void adam_pack_tile_t(const AdamItem& it, const AdamGroup& G, float clip, int bx, int by, float* tile) {
  const int T = TC > 0 ? TC : it.taps;
  const int Cout = it.cout, Cin = it.cin, COT = it.co_tile;
  const int ci0 = bx * 32, co0 = by * COT;
  const int TP = T | 1;
  const int CP = 32 * TP + 1;
  const int run = 32 * T;
  const int cot_shift = COT == 32 ? 5 : 3;
  const bool full = (co0 + COT <= Cout) && (ci0 + 32 <= Cin);
  const bool vec = full && (Cin & 3) == 0 &&
                   ((reinterpret_cast<uintptr_t>(it.p) | reinterpret_cast<uintptr_t>(it.g) | reinterpret_cast<uintptr_t>(it.m) |
                     reinterpret_cast<uintptr_t>(it.v)) & 15) == 0;
  if (vec) {
    for (int i4 = threadIdx.x; i4 < COT * run / 4; i4 += blockDim.x) {
      const int i = i4 * 4;
      const int co = i / run, r = i - co * run;  // r % 4 == 0 and run % 4 == 0: the four elements share the cout row
      const long long idx = (static_cast<long long>(co0 + co) * Cin + ci0) * T + r;
      float4 p = *reinterpret_cast<const float4*>(it.p + idx), m = *reinterpret_cast<const float4*>(it.m + idx),
             v = *reinterpret_cast<const float4*>(it.v + idx);
      const float4 g = *reinterpret_cast<const float4*>(it.g + idx);
      p.x = adam_one(p.x, g.x, m.x, v.x, G, clip);
      p.y = adam_one(p.y, g.y, m.y, v.y, G, clip);
      p.z = adam_one(p.z, g.z, m.z, v.z, G, clip);
      p.w = adam_one(p.w, g.w, m.w, v.w, G, clip);
      *reinterpret_cast<float4*>(it.p + idx) = p;
      *reinterpret_cast<float4*>(it.m + idx) = m;
      *reinterpret_cast<float4*>(it.v + idx) = v;
      const float pv[4] = {p.x, p.y, p.z, p.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int re = r + e;
        const int cil = re / T, t = re - cil * T;
        tile[co * CP + cil * TP + t] = pv[e];
      }
    }
  } else {
    for (int i = threadIdx.x; i < COT * run; i += blockDim.x) {
      const int co = i / run, r = i - co * run;
      const int cil = r / T, t = r - cil * T;
      float val = 0.f;
      if (full || (co0 + co < Cout && ci0 + cil < Cin)) {
        const long long idx = (static_cast<long long>(co0 + co) * Cin + ci0) * T + r;
        float m = it.m[idx], v = it.v[idx];
        val = adam_one(it.p[idx], it.g[idx], m, v, G, clip);
        it.p[idx] = val;
        it.m[idx] = m;
        it.v[idx] = v;
      }
      tile[co * CP + cil * TP + t] = val;
    }
  }
  __syncthreads();
  if (full && ((Cin | Cout) & 1) == 0) {
    // pairs: wf[co][t][ci, ci+1] and wd[ci][t][co, co+1] as 4-byte stores
    for (int i = threadIdx.x; i < COT * run / 2; i += blockDim.x) {
      {
        const int cil = (i & 15) * 2, q = i >> 4, co = q / T, t = q - co * T;
        const uint32_t pk = pack_bf16x2(tile[co * CP + cil * TP + t], tile[co * CP + (cil + 1) * TP + t]);
        *reinterpret_cast<uint32_t*>(it.wf + (static_cast<long long>(co0 + co) * T + t) * Cin + ci0 + cil) = pk;
      }
      if (it.wd) {
        const int col = (i & (COT / 2 - 1)) * 2, q = i >> (cot_shift - 1), cil = q / T, t = q - cil * T;
        const uint32_t pk = pack_bf16x2(tile[col * CP + cil * TP + t], tile[(col + 1) * CP + cil * TP + t]);
        *reinterpret_cast<uint32_t*>(it.wd + (static_cast<long long>(ci0 + cil) * T + t) * Cout + co0 + col) = pk;
      }
    }
    return;
  }
  for (int i = threadIdx.x; i < COT * run; i += blockDim.x) {
    {
      const int cil = i & 31, q = i >> 5, co = q / T, t = q - co * T;
      if (full || (co0 + co < Cout && ci0 + cil < Cin))
        it.wf[(static_cast<long long>(co0 + co) * T + t) * Cin + ci0 + cil] = __float2bfloat16_rn(tile[co * CP + cil * TP + t]);
    }
    if (it.wd) {
      const int col = i & (COT - 1), q = i >> cot_shift, cil = q / T, t = q - cil * T;
      if (full || (co0 + col < Cout && ci0 + cil < Cin))
        it.wd[(static_cast<long long>(ci0 + cil) * T + t) * Cout + co0 + col] = __float2bfloat16_rn(tile[col * CP + cil * TP + t]);
    }
  }
}
__device__ __forceinline__ void adam_pack_tile(const AdamItem& it, const AdamGroup& G, float clip, int bx, int by, float* tile) {
  if (it.taps == 1) adam_pack_tile_t<1>(it, G, clip, bx, by, tile);
  else if (it.taps == 9) adam_pack_tile_t<9>(it, G, clip, bx, by, tile);
  else if (it.taps == 27) adam_pack_tile_t<27>(it, G, clip, bx, by, tile);
  else adam_pack_tile_t<0>(it, G, clip, bx, by, tile);
}

__device__ __forceinline__ int find_item(const AdamItem* __restrict__ items, int nitems, int b) {
  int lo = 0, hi = nitems - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (items[mid].first_block <= b) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// clip: device scalar (global-norm clip coefficient from grad_clip_coef_kernel) or NULL.
__global__ void __launch_bounds__(256) adam_multi_kernel(const AdamItem* __restrict__ items, int nitems, const AdamGroups groups,
                                                         const float* __restrict__ clip) {
  extern __shared__ float tile[];
  const AdamItem it = items[find_item(items, nitems, blockIdx.x)];
  const AdamGroup G = groups.g[it.group];
  const float cc = (clip ? *clip : 1.f) * groups.grad_scale;
  const int local = blockIdx.x - it.first_block;
  if (it.wf) {
    adam_pack_tile(it, G, cc, local % it.ci_tiles, local / it.ci_tiles, tile);
    return;
  }
  const long long base = static_cast<long long>(local) * kAdamElemsPerBlock;
  if (base + kAdamElemsPerBlock <= it.n && ((reinterpret_cast<uintptr_t>(it.p) | reinterpret_cast<uintptr_t>(it.g) |
                                             reinterpret_cast<uintptr_t>(it.m) | reinterpret_cast<uintptr_t>(it.v)) & 15) == 0) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const long long i = base + (h * 256 + threadIdx.x) * 4;
      float4 p = *reinterpret_cast<const float4*>(it.p + i), m = *reinterpret_cast<const float4*>(it.m + i),
             v = *reinterpret_cast<const float4*>(it.v + i);
      const float4 g = *reinterpret_cast<const float4*>(it.g + i);
      p.x = adam_one(p.x, g.x, m.x, v.x, G, cc);
      p.y = adam_one(p.y, g.y, m.y, v.y, G, cc);
      p.z = adam_one(p.z, g.z, m.z, v.z, G, cc);
      p.w = adam_one(p.w, g.w, m.w, v.w, G, cc);
      *reinterpret_cast<float4*>(it.p + i) = p;
      *reinterpret_cast<float4*>(it.m + i) = m;
      *reinterpret_cast<float4*>(it.v + i) = v;
    }
  } else {
    for (long long i = base + threadIdx.x; i < base + kAdamElemsPerBlock && i < it.n; i += 256) {
      float m = it.m[i], v = it.v[i];
      it.p[i] = adam_one(it.p[i], it.g[i], m, v, G, cc);
      it.m[i] = m;
      it.v[i] = v;
    }
  }
}

// Global gradient norm (torch.nn.utils.clip_grad_norm_, norm_type 2): per-block partial sums of squares over the same
// item table (each block covers kAdamElemsPerBlock elements of one gradient), reduced in block order by one block.
struct NormItem {
  const float* g;
  long long n;
  int first_block, pad;
};
__global__ void __launch_bounds__(256) grad_sqnorm_multi_kernel(const NormItem* __restrict__ items, int nitems,
                                                                float* __restrict__ partial) {
  __shared__ float sh[8];
  int lo = 0, hi = nitems - 1;
  const int b = blockIdx.x;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (items[mid].first_block <= b) lo = mid; else hi = mid - 1;
  }
  const NormItem it = items[lo];
  const long long base = static_cast<long long>(b - it.first_block) * kAdamElemsPerBlock;
  float acc = 0.f;
  for (long long i = base + threadIdx.x; i < base + kAdamElemsPerBlock && i < it.n; i += 256) {
    const float g = it.g[i];
    acc = fmaf(g, g, acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sh[w];
    partial[b] = t;
  }
}
// total_norm = sqrt(sum partial); coef = min(1, max_norm / (total_norm + 1e-6)) (clip_grad_norm_'s clamp).
__global__ void __launch_bounds__(256) grad_clip_coef_kernel(const float* __restrict__ partial, int n, float max_norm,
                                                             float grad_scale, float* __restrict__ total_norm,
                                                             float* __restrict__ coef) {
  __shared__ double sh[256];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) acc += static_cast<double>(partial[i]);
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float tn = static_cast<float>(sqrt(sh[0])) * grad_scale;  // norm of the gradients as the optimizer will see them
    *total_norm = tn;
    const float c = max_norm / (tn + 1e-6f);
    *coef = c < 1.f ? c : 1.f;
  }
}

}  // namespace qt
