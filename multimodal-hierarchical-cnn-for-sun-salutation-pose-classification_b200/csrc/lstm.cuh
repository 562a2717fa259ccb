// nn.LSTM(batch_first=True) layers of the numeric-sequence branches (3dcnn/models.py:144-158,200-203: LSTM(47 -> 188, 2 layers);
// cnn+lstm/models.py:43-49,82-85: LSTM(640 -> 256, 2 layers)); torch gate order i, f, g, o; zero initial state.
//
// Sequences of different samples are independent, so a CTA owns kLstmBT samples for the whole time loop (persistent over
// t, the recurrent state never leaves the SM): thread j produces gate row j for its samples from the TRANSPOSED weights
// (row k of WihT / WhhT is read by consecutive threads -> coalesced; x_t and h_{t-1} are broadcast from shared memory),
// threads u < H then update c and h. Everything the backward needs (activated gates, c, h_{t-1}) is stored as it is made.
// The backward kernel runs the same loop in reverse (BPTT): it forms the gate gradients and carries dh / dc across time; the
// batched products that do not depend on the recurrence (dX = dG . Wih, dWih = dG^T X, dWhh = dG^T Hprev, db) run afterwards
// on the small-linear kernels. 4H <= 1024 threads per CTA (H <= 256: both reference sizes).
#pragma once
#include "elementwise.cuh"

namespace qt {

constexpr int kLstmBT = 2;  // samples per CTA: more CTAs matter more than weight reuse at the reference batch sizes (8-32)

__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }

__global__ void transpose_f32_kernel(const float* __restrict__ in, float* __restrict__ out, int rows, int cols) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int r = r0 + j, c = c0 + threadIdx.x;
    tile[j][threadIdx.x] = (r < rows && c < cols) ? in[static_cast<long long>(r) * cols + c] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int c = c0 + j, r = r0 + threadIdx.x;
    if (c < cols && r < rows) out[static_cast<long long>(c) * rows + r] = tile[threadIdx.x][j];
  }
}

// x [B][T][I] fp32 (in_drop_p > 0: inverted dropout with the counter-hash mask on element (b*T+t)*I + k, i.e. nn.LSTM's
// inter-layer dropout applied to the previous layer's output). Outputs: hseq, hprev (h_{t-1}, zeros at t = 0), cseq
// [B][T][H]; gates [B][T][4H] (activated i, f, g, o); x_used (optional) [B][T][I] the input after dropout. dynamic smem: kLstmBT * (I + H + 4H) floats.
__global__ void __launch_bounds__(1024) lstm_layer_fwd_kernel(const float* __restrict__ x, int I, const float* __restrict__ WihT,
                                                              const float* __restrict__ WhhT, const float* __restrict__ bih,
                                                              const float* __restrict__ bhh, int B, int T, int H, float in_drop_p,
                                                              unsigned long long seed, float* __restrict__ hseq,
                                                              float* __restrict__ hprev, float* __restrict__ cseq,
                                                              float* __restrict__ gates, float* __restrict__ x_used) {
  extern __shared__ float lsm[];
  float* xs = lsm;                      // [BT][I]
  float* hs = xs + kLstmBT * I;         // [BT][H]
  float* gs = hs + kLstmBT * H;         // [BT][4H]
  const int G = 4 * H;
  const int b0 = blockIdx.x * kLstmBT;
  const int j = threadIdx.x;
  for (int i = j; i < kLstmBT * H; i += blockDim.x) hs[i] = 0.f;
  float c[kLstmBT];
#pragma unroll
  for (int s = 0; s < kLstmBT; ++s) c[s] = 0.f;
  const float bias = j < G ? (bih ? bih[j] : 0.f) + (bhh ? bhh[j] : 0.f) : 0.f;
  for (int t = 0; t < T; ++t) {
    for (int i = j; i < kLstmBT * I; i += blockDim.x) {
      const int s = i / I, k = i - s * I;
      float v = 0.f;
      if (b0 + s < B) {
        const long long idx = (static_cast<long long>(b0 + s) * T + t) * I + k;
        v = x[idx] * dropout_scale(seed, static_cast<uint32_t>(idx), in_drop_p);
        if (x_used) x_used[idx] = v;  // the dropped-out input is the operand of this layer's weight gradient
      }
      xs[i] = v;
    }
    __syncthreads();  // xs of this step and hs of the previous step are complete
    if (j < G) {
      float acc[kLstmBT];
#pragma unroll
      for (int s = 0; s < kLstmBT; ++s) acc[s] = bias;
#pragma unroll 8
      for (int k = 0; k < I; ++k) {  // unrolled: eight independent L2 loads in flight per thread (the loop is latency bound)
        const float w = __ldg(WihT + static_cast<long long>(k) * G + j);
#pragma unroll
        for (int s = 0; s < kLstmBT; ++s) acc[s] = fmaf(w, xs[s * I + k], acc[s]);
      }
#pragma unroll 8
      for (int k = 0; k < H; ++k) {
        const float w = __ldg(WhhT + static_cast<long long>(k) * G + j);
#pragma unroll
        for (int s = 0; s < kLstmBT; ++s) acc[s] = fmaf(w, hs[s * H + k], acc[s]);
      }
      const bool is_g = (j >= 2 * H) && (j < 3 * H);
#pragma unroll
      for (int s = 0; s < kLstmBT; ++s) {
        const float a = is_g ? tanhf(acc[s]) : sigmoid_f(acc[s]);
        gs[s * G + j] = a;
        if (b0 + s < B) gates[(static_cast<long long>(b0 + s) * T + t) * G + j] = a;
      }
    }
    __syncthreads();  // gates complete; every thread is done reading hs
    if (j < H) {
#pragma unroll
      for (int s = 0; s < kLstmBT; ++s) {
        const float ig = gs[s * G + j], fg = gs[s * G + H + j], gg = gs[s * G + 2 * H + j], og = gs[s * G + 3 * H + j];
        const float hp = hs[s * H + j];
        c[s] = fmaf(fg, c[s], ig * gg);
        const float h = og * tanhf(c[s]);
        hs[s * H + j] = h;
        if (b0 + s < B) {
          const long long o = (static_cast<long long>(b0 + s) * T + t) * H + j;
          hseq[o] = h;
          cseq[o] = c[s];
          hprev[o] = hp;
        }
      }
    }
    // the next iteration's first __syncthreads orders these hs writes before the next reads
  }
}

// BPTT of one layer. dhseq [B][T][H]: gradient arriving at every h_t from above (out_drop_p > 0: it is the gradient of the
// dropped-out copy that fed the next layer, so the same mask (element (b*T+t)*H + u) is applied here). dgates [B][T][4H]
// receives the pre-activation gate gradients. dynamic smem: kLstmBT * (4H + H) + 4 * kLstmBT * H floats.
__global__ void __launch_bounds__(1024) lstm_layer_bwd_kernel(const float* __restrict__ dhseq, float out_drop_p,
                                                              unsigned long long seed, const float* __restrict__ Whh,
                                                              const float* __restrict__ gates, const float* __restrict__ cseq, int B,
                                                              int T, int H, float* __restrict__ dgates) {
  extern __shared__ float lsm[];
  const int G = 4 * H;
  float* dgs = lsm;                       // [BT][4H]
  float* dhn = dgs + kLstmBT * G;         // [BT][H]   dh carried from step t+1
  float* part = dhn + kLstmBT * H;        // [4][BT][H] partial products of the recurrent back-projection
  const int b0 = blockIdx.x * kLstmBT;
  const int j = threadIdx.x;
  for (int i = j; i < kLstmBT * H; i += blockDim.x) dhn[i] = 0.f;
  float dcn[kLstmBT];
#pragma unroll
  for (int s = 0; s < kLstmBT; ++s) dcn[s] = 0.f;
  __syncthreads();
  for (int t = T - 1; t >= 0; --t) {
    if (j < H) {
#pragma unroll
      for (int s = 0; s < kLstmBT; ++s) {
        float di = 0.f, df = 0.f, dg = 0.f, dob = 0.f;
        if (b0 + s < B) {
          const long long o = (static_cast<long long>(b0 + s) * T + t) * H + j;
          const long long go = (static_cast<long long>(b0 + s) * T + t) * G;
          float dh = dhn[s * H + j];
          if (dhseq) dh += dhseq[o] * dropout_scale(seed, static_cast<uint32_t>(o), out_drop_p);
          const float ig = gates[go + j], fg = gates[go + H + j], gg = gates[go + 2 * H + j], og = gates[go + 3 * H + j];
          const float ct = cseq[o];
          const float cp = t > 0 ? cseq[o - H] : 0.f;
          const float tc = tanhf(ct);
          dob = dh * tc * og * (1.f - og);
          const float dc = fmaf(dh * og, 1.f - tc * tc, dcn[s]);
          di = dc * gg * ig * (1.f - ig);
          df = dc * cp * fg * (1.f - fg);
          dg = dc * ig * (1.f - gg * gg);
          dcn[s] = dc * fg;
          dgates[go + j] = di;
          dgates[go + H + j] = df;
          dgates[go + 2 * H + j] = dg;
          dgates[go + 3 * H + j] = dob;
        }
        dgs[s * G + j] = di;
        dgs[s * G + H + j] = df;
        dgs[s * G + 2 * H + j] = dg;
        dgs[s * G + 3 * H + j] = dob;
      }
    }
    __syncthreads();
    if (j < G) {  // dh_{t-1}[k] = sum_j dG[j] * Whh[j][k]: quarter q of the rows per thread group, coalesced over k
      const int q = j / H, k = j - q * H;
      float acc[kLstmBT];
#pragma unroll
      for (int s = 0; s < kLstmBT; ++s) acc[s] = 0.f;
#pragma unroll 8
      for (int r = q * H; r < (q + 1) * H; ++r) {
        const float w = __ldg(Whh + static_cast<long long>(r) * H + k);
#pragma unroll
        for (int s = 0; s < kLstmBT; ++s) acc[s] = fmaf(w, dgs[s * G + r], acc[s]);
      }
#pragma unroll
      for (int s = 0; s < kLstmBT; ++s) part[(q * kLstmBT + s) * H + k] = acc[s];
    }
    __syncthreads();
    if (j < H) {
#pragma unroll
      for (int s = 0; s < kLstmBT; ++s)
        dhn[s * H + j] = (part[(0 * kLstmBT + s) * H + j] + part[(1 * kLstmBT + s) * H + j]) +
                         (part[(2 * kLstmBT + s) * H + j] + part[(3 * kLstmBT + s) * H + j]);
    }
    __syncthreads();
  }
}

}  // namespace qt
