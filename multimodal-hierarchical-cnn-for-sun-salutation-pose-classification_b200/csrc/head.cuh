// Loss and fused classifier tail of the fusion head.
//
//   cross_entropy_kernel : nn.CrossEntropyLoss() of the training scripts (Quadtree_from scratch/Quadtree_train.py:44,64):
//                          mean over the batch of logsumexp(logits) - logits[label], and its gradient
//                          dlogits = (softmax - onehot) * grad_scale / B in the same pass. One warp per row; the mean is
//                          formed by the last block in row order (fixed order -> bitwise reproducible).
//   head_tail_kernel     : everything behind the classifier.0 GEMM in ONE launch (QS/models.py:268-271,303 + the script's
//                          criterion): ReLU + Dropout of the hidden row, classifier.3 (nhid -> nc), log-softmax + NLL,
//                          dlogits, and the gradient w.r.t. the hidden row (dlogits . W3 through the ReLU / dropout mask),
//                          written as the bf16 operand of the classifier.0 data/weight-gradient GEMMs. One CTA per sample.
#pragma once
#include "elementwise.cuh"

namespace qt {

constexpr int kMaxClasses = 32;

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Row-wise softmax cross entropy for one warp: lane c holds logit c (lanes >= nc: -inf). Returns the row loss and leaves
// the softmax probability of class `lane` in `prob`.
__device__ __forceinline__ float ce_row(float logit, int label, int lane, float& prob) {
  const float mx = warp_max(logit);
  const float e = expf(logit - mx);  // exp(-inf) = 0 for the padding lanes
  const float se = warp_sum(e);
  prob = e / se;
  const float picked = __shfl_sync(0xffffffffu, logit, label & 31);
  return (mx + logf(se)) - picked;
}

// Last-block-done mean: every block publishes its rows' losses, the last one to finish adds them in row order.
__device__ __forceinline__ void loss_mean_last_block(const float* loss_rows, int B, float* loss_mean, unsigned int* counter,
                                                     unsigned int nblocks) {
  __shared__ bool is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(counter, 1u) == nblocks - 1);
  __syncthreads();
  if (is_last && threadIdx.x < 32) {
    // fixed order: lane l adds rows l, l+32, ... then a shuffle tree
    float acc = 0.f;
    for (int r = threadIdx.x; r < B; r += 32) acc += __ldcg(loss_rows + r);
    acc = warp_sum(acc);
    if (threadIdx.x == 0) {
      *loss_mean = acc / static_cast<float>(B);
      *counter = 0;  // ready for the next launch
    }
  }
}

// logits fp32 [B][nc] (row stride ld), labels int64. dlogits may be NULL (evaluation).
__global__ void cross_entropy_kernel(const float* __restrict__ logits, long long ld, const long long* __restrict__ labels, int B,
                                     int nc, float grad_scale, const float* __restrict__ upstream, float* __restrict__ loss_rows,
                                     float* __restrict__ loss_mean, float* __restrict__ dlogits,
                                     unsigned int* __restrict__ counter) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + warp;
  if (row < B) {
    const int label = static_cast<int>(labels[row]);
    const float lg = lane < nc ? logits[row * ld + lane] : -INFINITY;
    float prob;
    const float loss = ce_row(lg, label, lane, prob);
    if (lane == 0) loss_rows[row] = loss;
    if (dlogits && lane < nc)
      dlogits[static_cast<long long>(row) * nc + lane] = (prob - (lane == label ? 1.f : 0.f)) * grad_scale * (upstream ? *upstream : 1.f);
  }
  loss_mean_last_block(loss_rows, B, loss_mean, counter, gridDim.x);
}

// One CTA (256 threads) per sample. h: fp32 [B][nhid] pre-activation of classifier.0 (bias included); on exit it holds
// the activated row relu(h)*dropout (operand of the classifier.3 weight gradient and the mask of the backward).
// labels == NULL: logits only (the scripts own the criterion). nhid <= 256 * kHeadCols, nc <= kMaxClasses.
constexpr int kHeadThreads = 256;
constexpr int kHeadCols = 16;
__global__ void __launch_bounds__(kHeadThreads) head_tail_fwd_kernel(float* __restrict__ h, __nv_bfloat16* __restrict__ h16, int nhid,
                                                                     const float* __restrict__ w3, const float* __restrict__ b3,
                                                                     int nc, const long long* __restrict__ labels, int B,
                                                                     float drop_p, unsigned long long seed,
                                                                     float* __restrict__ logits, float* __restrict__ loss_rows,
                                                                     float* __restrict__ loss_mean,
                                                                     unsigned int* __restrict__ counter) {
  __shared__ float part[kHeadThreads / 32][kMaxClasses];
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* hrow = h + static_cast<long long>(b) * nhid;
  float a[kHeadCols];
#pragma unroll
  for (int i = 0; i < kHeadCols; ++i) {
    const int j = threadIdx.x + i * kHeadThreads;
    a[i] = 0.f;
    if (j < nhid) {
      const float s = dropout_scale(seed, static_cast<uint32_t>(b) * nhid + j, drop_p);  // same index as relu_dropout_kernel
      a[i] = fmaxf(hrow[j], 0.f) * s;
      hrow[j] = a[i];
      if (h16) h16[static_cast<long long>(b) * nhid + j] = __float2bfloat16_rn(a[i]);
    }
  }
  // classifier.3: per class a block-wide dot product (coalesced reads of w3, fixed reduction order)
  for (int c = 0; c < nc; ++c) {
    const float* wr = w3 + static_cast<long long>(c) * nhid;
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < kHeadCols; ++i) {
      const int j = threadIdx.x + i * kHeadThreads;
      if (j < nhid) acc = fmaf(a[i], __ldg(wr + j), acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) part[warp][c] = acc;
  }
  __syncthreads();
  if (warp == 0) {
    float lg = -INFINITY;
    if (lane < nc) {
      float t = b3 ? b3[lane] : 0.f;
#pragma unroll
      for (int w = 0; w < kHeadThreads / 32; ++w) t += part[w][lane];
      lg = t;
      logits[static_cast<long long>(b) * nc + lane] = t;
    }
    if (labels) {
      float prob;
      const float loss = ce_row(lg, static_cast<int>(labels[b]), lane, prob);
      if (lane == 0) loss_rows[b] = loss;
    }
  }
  if (labels) loss_mean_last_block(loss_rows, B, loss_mean, counter, gridDim.x);
}

// Backward of the tail, one CTA per sample. With labels: dlogits = (softmax(logits) - onehot) * grad_scale * upstream
// (upstream: device scalar, the gradient arriving at the mean loss; NULL = 1) is formed here and also written out (the
// classifier.3 weight gradient needs it); without labels the caller supplies dlogits (autograd's gradient of the logits).
// dh16 = (dlogits . W3) * dropout_scale * 1[act > 0] as bf16.
__global__ void __launch_bounds__(kHeadThreads) head_tail_bwd_kernel(const float* __restrict__ act, int nhid,
                                                                     const float* __restrict__ w3, int nc,
                                                                     const float* __restrict__ logits,
                                                                     const long long* __restrict__ labels, float grad_scale,
                                                                     const float* __restrict__ upstream, float drop_p,
                                                                     unsigned long long seed, float* __restrict__ dlogits,
                                                                     __nv_bfloat16* __restrict__ dh16) {
  __shared__ float dl_s[kMaxClasses];
  const int b = blockIdx.x;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x < 32) {
    if (labels) {
      const int label = static_cast<int>(labels[b]);
      const float lg = lane < nc ? logits[static_cast<long long>(b) * nc + lane] : -INFINITY;
      float prob;
      (void)ce_row(lg, label, lane, prob);
      const float d = (prob - (lane == label ? 1.f : 0.f)) * grad_scale * (upstream ? *upstream : 1.f);
      if (lane < nc) {
        dl_s[lane] = d;
        dlogits[static_cast<long long>(b) * nc + lane] = d;
      }
    } else if (lane < nc) {
      dl_s[lane] = dlogits[static_cast<long long>(b) * nc + lane];
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < nhid; j += kHeadThreads) {
    float acc = 0.f;
    for (int c = 0; c < nc; ++c) acc = fmaf(dl_s[c], __ldg(w3 + static_cast<long long>(c) * nhid + j), acc);
    const float a = act[static_cast<long long>(b) * nhid + j];
    acc = (a > 0.f) ? acc * dropout_scale(seed, static_cast<uint32_t>(b) * nhid + j, drop_p) : 0.f;
    dh16[static_cast<long long>(b) * nhid + j] = __float2bfloat16_rn(acc);
  }
}

}  // namespace qt
