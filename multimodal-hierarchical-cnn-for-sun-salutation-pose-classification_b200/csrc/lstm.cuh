// nn.LSTM(batch_first=True) layers of the numeric-sequence branches (3dcnn/models.py:144-158,200-203: LSTM(47 -> 188, 2 layers);
// cnn+lstm/models.py:43-49,82-85: LSTM(640 -> 256, 2 layers)); torch gate order i, f, g, o; zero initial state.
//
// The input projection x_t Wih^T + b_ih of all time steps is one batched product (qt_small_linear_fwd) outside the time
// loop. The recurrence runs on thread-block CLUSTERS of 8 CTAs that keep the recurrent weights in shared memory for the
// whole sequence: CTA r of a cluster owns the hidden units [r*Hs, (r+1)*Hs) (Hs = ceil(H/8)) — their four gate rows of
// Whh (H x 4Hs floats: 72 KB at H = 188, 128 KB at H = 256) are loaded once, and each step only exchanges the new h slice
// with the seven peers through distributed shared memory (st.shared::cluster) followed by one cluster barrier. A first
// version that re-read Whh from L2 every step was latency bound at ~30 GB/s per SM (0.33 ms per layer forward at B = 32);
// with resident weights a step is a few hundred shared-memory FMAs plus the barrier.
// The backward kernel (BPTT) has the same shape: the CTA owns the same units, forms their gate gradients, broadcasts that
// dG slice, and back-projects dh_{t-1}[k] = sum_j dG[j] Whh[j][k] for its own k from the column slice of Whh it keeps in
// shared memory. The batched products that do not depend on the recurrence (dX = dG . Wih, dWih = dG^T X,
// dWhh = dG^T Hprev, db) run afterwards on the small-linear kernels. H <= 256.
#pragma once
#include "elementwise.cuh"

namespace qt {

constexpr int kLstmCluster = 8;   // CTAs per cluster (portable maximum)
constexpr int kLstmBT = 8;        // samples per cluster
constexpr int kLstmThreads = 512;

__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
// store one float into the same shared-memory offset of CTA `rank` of this cluster
__device__ __forceinline__ void dsmem_store(float* local_ptr, uint32_t rank, float v) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(local_ptr)), "r"(rank));
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(remote), "f"(v) : "memory");
}

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

// xproj [B][T][4H] = x Wih^T + b_ih (fp32); whh_t [H][4H] (transposed torch weight); bhh [4H] or NULL.
// Outputs hseq, hprev (h_{t-1}, zeros at t = 0), cseq [B][T][H]; gates [B][T][4H] (activated i, f, g, o).
// grid = ceil(B / kLstmBT) clusters of 8 CTAs. dynamic smem: H*4Hs (weights) + 2*BT*H (h ping-pong) + BT*4Hs (gates) floats.
__global__ void __launch_bounds__(kLstmThreads, 1) lstm_layer_fwd_kernel(const float* __restrict__ xproj, const float* __restrict__ whh_t,
                                                                         const float* __restrict__ bhh, int B, int T, int H, int Hs,
                                                                         float* __restrict__ hseq, float* __restrict__ hprev,
                                                                         float* __restrict__ cseq, float* __restrict__ gates) {
  extern __shared__ float lsm[];
  const int G = 4 * H, R = 4 * Hs;       // global / local gate rows
  float* wsl = lsm;                      // [H][R]: wsl[k][g*Hs + ul] = Whh[g*H + u0 + ul][k]
  float* hs = wsl + H * R;               // [2][BT][H] full hidden state, ping-pong
  float* gs = hs + 2 * kLstmBT * H;      // [BT][R] activated gates of the own units
  const uint32_t rank = cluster_rank();
  const int b0 = (blockIdx.x / kLstmCluster) * kLstmBT;
  const int u0 = rank * Hs;
  const int nu = max(0, min(Hs, H - u0));  // units this CTA really owns
  for (int i = threadIdx.x; i < H * R; i += kLstmThreads) {
    const int k = i / R, r = i - k * R;
    const int g = r / Hs, ul = r - g * Hs;
    wsl[i] = ul < nu ? whh_t[static_cast<long long>(k) * G + g * H + u0 + ul] : 0.f;
  }
  for (int i = threadIdx.x; i < 2 * kLstmBT * H; i += kLstmThreads) hs[i] = 0.f;
  cluster_sync_all();  // every CTA's h buffers are zeroed before any peer writes into them
  // gate phase: thread (r, sg) handles local gate row r for the samples s = sg, sg + SG, ...
  const int SG = kLstmThreads / R;       // >= 4 for Hs <= 32
  const int r_g = threadIdx.x % R, sg = threadIdx.x / R;
  const int g_g = r_g / Hs, ul_g = r_g - g_g * Hs;
  const bool gate_thread = sg < SG && ul_g < nu;
  const int jrow = g_g * H + u0 + ul_g;  // global gate row
  const float bj = (gate_thread && bhh) ? bhh[jrow] : 0.f;
  // cell phase: thread (ul, s)
  const int ul_c = threadIdx.x % Hs, s_c = threadIdx.x / Hs;
  const bool cell_thread = s_c < kLstmBT && ul_c < nu && (b0 + s_c) < B;
  float c = 0.f;
  for (int t = 0; t < T; ++t) {
    const float* hcur = hs + (t & 1) * kLstmBT * H;
    float* hnext = hs + ((t + 1) & 1) * kLstmBT * H;
    if (gate_thread) {
      for (int s = sg; s < kLstmBT; s += SG) {
        if (b0 + s >= B) break;
        float acc = xproj[(static_cast<long long>(b0 + s) * T + t) * G + jrow] + bj;
        const float* hrow = hcur + s * H;
#pragma unroll 4
        for (int k = 0; k < H; ++k) acc = fmaf(wsl[k * R + r_g], hrow[k], acc);
        const float a = (g_g == 2) ? tanhf(acc) : sigmoid_f(acc);
        gs[s * R + r_g] = a;
        gates[(static_cast<long long>(b0 + s) * T + t) * G + jrow] = a;
      }
    }
    __syncthreads();
    if (cell_thread) {
      const float ig = gs[s_c * R + ul_c], fg = gs[s_c * R + Hs + ul_c], gg = gs[s_c * R + 2 * Hs + ul_c], og = gs[s_c * R + 3 * Hs + ul_c];
      const int u = u0 + ul_c;
      c = fmaf(fg, c, ig * gg);
      const float h = og * tanhf(c);
      const long long o = (static_cast<long long>(b0 + s_c) * T + t) * H + u;
      hseq[o] = h;
      cseq[o] = c;
      hprev[o] = hcur[s_c * H + u];
      float* dst = hnext + s_c * H + u;
#pragma unroll
      for (uint32_t pr = 0; pr < kLstmCluster; ++pr) dsmem_store(dst, pr, h);  // the new h slice goes to every CTA of the cluster
    }
    cluster_sync_all();  // all slices of h_t have landed everywhere (release / acquire), gs may be overwritten
  }
}

// BPTT of one layer. dhseq [B][T][H]: gradient arriving at every h_t from above (out_drop_p > 0: it is the gradient of the
// dropped-out copy that fed the next layer, so the same counter-hash mask (element (b*T+t)*H + u) is applied here).
// whh [4H][H] (torch layout). dgates [B][T][4H] receives the pre-activation gate gradients.
// dynamic smem: 4H*Hs (weight column slice) + 2*BT*4H (dG ping-pong) + BT*Hs (dh carried) + 512 (partials) floats.
__global__ void __launch_bounds__(kLstmThreads, 1) lstm_layer_bwd_kernel(const float* __restrict__ dhseq, float out_drop_p,
                                                                         unsigned long long seed, const float* __restrict__ whh,
                                                                         const float* __restrict__ gates, const float* __restrict__ cseq,
                                                                         int B, int T, int H, int Hs, float* __restrict__ dgates) {
  extern __shared__ float lsm[];
  const int G = 4 * H;
  float* wk = lsm;                        // [4H][Hs]: wk[j][kl] = Whh[j][u0 + kl]
  float* dgf = wk + G * Hs;               // [2][BT][4H] full gate gradients, ping-pong
  float* dhn = dgf + 2 * kLstmBT * G;     // [BT][Hs] dh carried to step t-1 (own units)
  float* part = dhn + kLstmBT * Hs;       // [P][BT][Hs] partial back-projections (<= 512 floats)
  const uint32_t rank = cluster_rank();
  const int b0 = (blockIdx.x / kLstmCluster) * kLstmBT;
  const int u0 = rank * Hs;
  const int nu = max(0, min(Hs, H - u0));
  for (int i = threadIdx.x; i < G * Hs; i += kLstmThreads) {
    const int j = i / Hs, kl = i - j * Hs;
    wk[i] = kl < nu ? whh[static_cast<long long>(j) * H + u0 + kl] : 0.f;
  }
  for (int i = threadIdx.x; i < 2 * kLstmBT * G; i += kLstmThreads) dgf[i] = 0.f;
  for (int i = threadIdx.x; i < kLstmBT * Hs; i += kLstmThreads) dhn[i] = 0.f;
  cluster_sync_all();
  const int ul_c = threadIdx.x % Hs, s_c = threadIdx.x / Hs;
  const bool cell_thread = s_c < kLstmBT && ul_c < nu && (b0 + s_c) < B;
  const int P = kLstmThreads / (Hs * kLstmBT);  // >= 2 for Hs <= 32
  const int o_idx = threadIdx.x % (Hs * kLstmBT), part_id = threadIdx.x / (Hs * kLstmBT);
  const int kl_p = o_idx % Hs, s_p = o_idx / Hs;
  float dcn = 0.f;
  for (int t = T - 1; t >= 0; --t) {
    float* dcur = dgf + (t & 1) * kLstmBT * G;
    if (cell_thread) {
      const int u = u0 + ul_c;
      const long long o = (static_cast<long long>(b0 + s_c) * T + t) * H + u;
      const long long go = (static_cast<long long>(b0 + s_c) * T + t) * G;
      float dh = dhn[s_c * Hs + ul_c];
      if (dhseq) dh += dhseq[o] * dropout_scale(seed, static_cast<uint32_t>(o), out_drop_p);
      const float ig = gates[go + u], fg = gates[go + H + u], gg = gates[go + 2 * H + u], og = gates[go + 3 * H + u];
      const float ct = cseq[o];
      const float cp = t > 0 ? cseq[o - H] : 0.f;
      const float tc = tanhf(ct);
      const float dob = dh * tc * og * (1.f - og);
      const float dc = fmaf(dh * og, 1.f - tc * tc, dcn);
      const float di = dc * gg * ig * (1.f - ig);
      const float df = dc * cp * fg * (1.f - fg);
      const float dg = dc * ig * (1.f - gg * gg);
      dcn = dc * fg;
      dgates[go + u] = di;
      dgates[go + H + u] = df;
      dgates[go + 2 * H + u] = dg;
      dgates[go + 3 * H + u] = dob;
      float* base = dcur + s_c * G + u;
#pragma unroll
      for (uint32_t pr = 0; pr < kLstmCluster; ++pr) {
        dsmem_store(base, pr, di);
        dsmem_store(base + H, pr, df);
        dsmem_store(base + 2 * H, pr, dg);
        dsmem_store(base + 3 * H, pr, dob);
      }
    }
    cluster_sync_all();  // the full dG_t of every sample is in every CTA
    if (t > 0) {
      // dh_{t-1}[u0 + kl] = sum_j dG[j] * Whh[j][u0 + kl]: P thread groups take interleaved j, combined in fixed order
      if (part_id < P) {
        float acc = 0.f;
        if ((b0 + s_p) < B) {
          const float* dgrow = dcur + s_p * G;
#pragma unroll 4
          for (int j = part_id; j < G; j += P) acc = fmaf(dgrow[j], wk[j * Hs + kl_p], acc);
        }
        part[(part_id * kLstmBT + s_p) * Hs + kl_p] = acc;
      }
      __syncthreads();
      if (threadIdx.x < Hs * kLstmBT) {
        float tot = 0.f;
        for (int q = 0; q < P; ++q) tot += part[(q * kLstmBT + s_p) * Hs + kl_p];
        dhn[s_p * Hs + kl_p] = tot;
      }
      __syncthreads();
    }
  }
  cluster_sync_all();  // no CTA exits while peers may still write into its shared memory
}

}  // namespace qt
