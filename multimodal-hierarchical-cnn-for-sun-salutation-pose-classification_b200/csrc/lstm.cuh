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

// Shared-memory budget helpers (floats) — the host sizes the launch with the same expressions.
__host__ __device__ constexpr int lstm_fwd_kgroups(int R) { return (kLstmThreads / (R / 2)) < 16 ? (kLstmThreads / (R / 2)) : 16; }
__host__ __device__ constexpr size_t lstm_fwd_smem_floats(int H, int Hs) {
  return static_cast<size_t>(H) * 4 * Hs + 2 * kLstmBT * H + static_cast<size_t>(lstm_fwd_kgroups(4 * Hs)) * kLstmBT * 4 * Hs + kLstmBT * 4 * Hs;
}
__host__ __device__ constexpr size_t lstm_bwd_smem_floats(int H, int Hs) {
  return static_cast<size_t>(4 * H) * Hs + 2 * kLstmBT * 4 * H + kLstmBT * Hs + (kLstmThreads / 32) * kLstmBT * Hs;
}

// xproj [B][T][4H] = x Wih^T + b_ih (fp32); whh_t [H][4H] (transposed torch weight); bhh [4H] or NULL.
// Outputs hseq, hprev (h_{t-1}, zeros at t = 0), cseq [B][T][H]; gates [B][T][4H] (activated i, f, g, o).
// grid = ceil(B / kLstmBT) clusters of 8 CTAs; Hs even. One time step:
//   (1) partial mat-vec: thread (row pair, k group) accumulates 2 gate rows x 8 samples over its k = kg, kg + KG, ...: one
//       8-byte weight load and two 16-byte broadcast loads of h (stored [k][sample]) per 16 FMAs. The first version did two
//       shared loads per FMA and was bound by shared-memory instruction issue at 11 us per step;
//   (2) the KG partials of each (sample, row) are summed in fixed order, the input projection (prefetched one step ahead: a
//       global load on the recurrence's critical path costs more than the mat-vec) and bias are added, gates activated;
//   (3) cell update for the own units, new h slice broadcast to the eight CTAs (st.shared::cluster), one cluster barrier.
__global__ void __launch_bounds__(kLstmThreads, 1) lstm_layer_fwd_kernel(const float* __restrict__ xproj, const float* __restrict__ whh_t,
                                                                         const float* __restrict__ bhh, int B, int T, int H, int Hs,
                                                                         float* __restrict__ hseq, float* __restrict__ hprev,
                                                                         float* __restrict__ cseq, float* __restrict__ gates) {
  extern __shared__ __align__(16) float lsm[];
  const int G = 4 * H, R = 4 * Hs, RP = R / 2;  // global / local gate rows, local row pairs
  const int KG = lstm_fwd_kgroups(R);
  float* wsl = lsm;                       // [H][R]: wsl[k][g*Hs + ul] = Whh[g*H + u0 + ul][k]
  float* hs = wsl + H * R;                // [2][H][BT] full hidden state, ping-pong, sample fastest
  float* part = hs + 2 * kLstmBT * H;     // [KG][BT][R] partial pre-activations
  float* gs = part + KG * kLstmBT * R;    // [BT][R] activated gates of the own units
  const uint32_t rank = cluster_rank();
  const int b0 = (blockIdx.x / kLstmCluster) * kLstmBT;
  const int u0 = rank * Hs;
  const int nu = max(0, min(Hs, H - u0));  // units this CTA really owns
  const int tid = threadIdx.x;
  for (int i = tid; i < H * R; i += kLstmThreads) {
    const int k = i / R, r = i - k * R;
    const int g = r / Hs, ul = r - g * Hs;
    wsl[i] = ul < nu ? whh_t[static_cast<long long>(k) * G + g * H + u0 + ul] : 0.f;
  }
  for (int i = tid; i < 2 * kLstmBT * H; i += kLstmThreads) hs[i] = 0.f;
  cluster_sync_all();  // every CTA's h buffers are zeroed before any peer writes into them
  // (1) mat-vec role
  const int rp = tid % RP, kg = tid / RP;
  const bool mv_thread = kg < KG;
  // (2) reduce role: outputs o = s*R + r for o = tid and tid + 512
  int o_s[2], o_r[2], o_g[2];
  long long o_x[2];   // offset of (b0+s, t = 0, jrow) in xproj / gates
  float o_bias[2];
  bool o_ok[2];
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int o = tid + q * kLstmThreads;
    const int sidx = o / R, r = o - sidx * R;
    const int g = r / Hs, ul = r - g * Hs;
    o_s[q] = sidx; o_r[q] = r; o_g[q] = g;
    o_ok[q] = (o < kLstmBT * R) && (ul < nu) && (b0 + sidx < B);
    const int jrow = g * H + u0 + ul;
    o_x[q] = o_ok[q] ? static_cast<long long>(b0 + sidx) * T * G + jrow : 0;
    o_bias[q] = (o_ok[q] && bhh) ? bhh[jrow] : 0.f;
  }
  // (3) cell role: thread (s, ul), sample fastest so the broadcast h slice is contiguous
  const int s_c = tid % kLstmBT, ul_c = tid / kLstmBT;
  const bool cell_thread = ul_c < nu && (b0 + s_c) < B;
  float c = 0.f;
  float xp[2];
#pragma unroll
  for (int q = 0; q < 2; ++q) xp[q] = o_ok[q] ? xproj[o_x[q]] : 0.f;
  for (int t = 0; t < T; ++t) {
    const float* hcur = hs + (t & 1) * kLstmBT * H;
    float* hnext = hs + ((t + 1) & 1) * kLstmBT * H;
    if (mv_thread) {
      float acc0[kLstmBT] = {}, acc1[kLstmBT] = {};
#pragma unroll 4
      for (int k = kg; k < H; k += KG) {
        const float2 w = *reinterpret_cast<const float2*>(wsl + k * R + 2 * rp);
        const float4 ha = *reinterpret_cast<const float4*>(hcur + k * kLstmBT);
        const float4 hb = *reinterpret_cast<const float4*>(hcur + k * kLstmBT + 4);
        const float hv[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
#pragma unroll
        for (int sidx = 0; sidx < kLstmBT; ++sidx) {
          acc0[sidx] = fmaf(w.x, hv[sidx], acc0[sidx]);
          acc1[sidx] = fmaf(w.y, hv[sidx], acc1[sidx]);
        }
      }
#pragma unroll
      for (int sidx = 0; sidx < kLstmBT; ++sidx)
        *reinterpret_cast<float2*>(part + (kg * kLstmBT + sidx) * R + 2 * rp) = make_float2(acc0[sidx], acc1[sidx]);
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      if (!o_ok[q]) continue;
      float acc = xp[q] + o_bias[q];
      for (int g2 = 0; g2 < KG; ++g2) acc += part[(g2 * kLstmBT + o_s[q]) * R + o_r[q]];
      const float a = (o_g[q] == 2) ? tanhf(acc) : sigmoid_f(acc);
      gs[o_s[q] * R + o_r[q]] = a;
      gates[o_x[q] + static_cast<long long>(t) * G] = a;
      if (t + 1 < T) xp[q] = xproj[o_x[q] + static_cast<long long>(t + 1) * G];
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
      hprev[o] = hcur[u * kLstmBT + s_c];
      float* dst = hnext + u * kLstmBT + s_c;
#pragma unroll
      for (uint32_t pr = 0; pr < kLstmCluster; ++pr) dsmem_store(dst, pr, h);  // the new h slice goes to every CTA of the cluster
    }
    cluster_sync_all();  // all slices of h_t have landed everywhere (release / acquire), part / gs may be overwritten
  }
}

// BPTT of one layer. dhseq [B][T][H]: gradient arriving at every h_t from above (out_drop_p > 0: it is the gradient of the
// dropped-out copy that fed the next layer, so the same counter-hash mask (element (b*T+t)*H + u) is applied here).
// whh [4H][H] (torch layout). dgates [B][T][4H] receives the pre-activation gate gradients. Hs even.
// One step: gate gradients of the own units (inputs prefetched one step ahead) -> broadcast of that dG slice (stored
// [j][sample]) -> cluster barrier -> back-projection dh_{t-1}[own k] = sum_j dG[j] Whh[j][k]: thread (unit pair, j group)
// accumulates 2 units x 8 samples with one 8-byte weight load and two 16-byte dG loads per 16 FMAs; the 32 j groups are
// combined by one shuffle and a fixed-order sum over the 16 warps.
__global__ void __launch_bounds__(kLstmThreads, 1) lstm_layer_bwd_kernel(const float* __restrict__ dhseq, float out_drop_p,
                                                                         unsigned long long seed, const float* __restrict__ whh,
                                                                         const float* __restrict__ gates, const float* __restrict__ cseq,
                                                                         int B, int T, int H, int Hs, float* __restrict__ dgates) {
  extern __shared__ __align__(16) float lsm[];
  const int G = 4 * H;
  constexpr int kWarps = kLstmThreads / 32;
  float* wk = lsm;                        // [4H][Hs]: wk[j][kl] = Whh[j][u0 + kl]
  float* dgf = wk + G * Hs;               // [2][4H][BT] full gate gradients, ping-pong, sample fastest
  float* dhn = dgf + 2 * kLstmBT * G;     // [BT][Hs] dh carried to step t-1 (own units)
  float* part = dhn + kLstmBT * Hs;       // [kWarps][BT][Hs] partial back-projections
  const uint32_t rank = cluster_rank();
  const int b0 = (blockIdx.x / kLstmCluster) * kLstmBT;
  const int u0 = rank * Hs;
  const int nu = max(0, min(Hs, H - u0));
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < G * Hs; i += kLstmThreads) {
    const int j = i / Hs, kl = i - j * Hs;
    wk[i] = kl < nu ? whh[static_cast<long long>(j) * H + u0 + kl] : 0.f;
  }
  for (int i = tid; i < 2 * kLstmBT * G; i += kLstmThreads) dgf[i] = 0.f;
  for (int i = tid; i < kLstmBT * Hs; i += kLstmThreads) dhn[i] = 0.f;
  cluster_sync_all();
  // cell role: thread (s, ul), sample fastest (contiguous broadcast)
  const int s_c = tid % kLstmBT, ul_c = tid / kLstmBT;
  const bool cell_thread = ul_c < nu && (b0 + s_c) < B;
  // back-projection role: lane = (unit pair, j parity), 32 j groups over the CTA
  const int kp = lane & 15, jl = (lane >> 4) + 2 * warp;
  const bool bp_thread = 2 * kp < Hs;
  // final-sum role: thread (s, kl)
  const int s_f = tid / Hs, kl_f = tid - s_f * Hs;
  const bool fin_thread = tid < kLstmBT * Hs;
  float dcn = 0.f;
  // step inputs, prefetched one step ahead
  float p_ig = 0.f, p_fg = 0.f, p_gg = 0.f, p_og = 0.f, p_ct = 0.f, p_cp = 0.f, p_dh = 0.f;
  auto prefetch = [&](int t) {
    if (!cell_thread) return;
    const int u = u0 + ul_c;
    const long long o = (static_cast<long long>(b0 + s_c) * T + t) * H + u;
    const long long go = (static_cast<long long>(b0 + s_c) * T + t) * G;
    p_ig = gates[go + u]; p_fg = gates[go + H + u]; p_gg = gates[go + 2 * H + u]; p_og = gates[go + 3 * H + u];
    p_ct = cseq[o];
    p_cp = t > 0 ? cseq[o - H] : 0.f;
    p_dh = dhseq ? dhseq[o] * dropout_scale(seed, static_cast<uint32_t>(o), out_drop_p) : 0.f;
  };
  prefetch(T - 1);
  for (int t = T - 1; t >= 0; --t) {
    float* dcur = dgf + (t & 1) * kLstmBT * G;
    if (cell_thread) {
      const int u = u0 + ul_c;
      const long long go = (static_cast<long long>(b0 + s_c) * T + t) * G;
      const float dh = dhn[s_c * Hs + ul_c] + p_dh;
      const float ig = p_ig, fg = p_fg, gg = p_gg, og = p_og;
      const float tc = tanhf(p_ct);
      const float dob = dh * tc * og * (1.f - og);
      const float dc = fmaf(dh * og, 1.f - tc * tc, dcn);
      const float di = dc * gg * ig * (1.f - ig);
      const float df = dc * p_cp * fg * (1.f - fg);
      const float dg = dc * ig * (1.f - gg * gg);
      dcn = dc * fg;
      dgates[go + u] = di;
      dgates[go + H + u] = df;
      dgates[go + 2 * H + u] = dg;
      dgates[go + 3 * H + u] = dob;
      float* base = dcur + u * kLstmBT + s_c;
#pragma unroll
      for (uint32_t pr = 0; pr < kLstmCluster; ++pr) {
        dsmem_store(base, pr, di);
        dsmem_store(base + H * kLstmBT, pr, df);
        dsmem_store(base + 2 * H * kLstmBT, pr, dg);
        dsmem_store(base + 3 * H * kLstmBT, pr, dob);
      }
    }
    if (t > 0) prefetch(t - 1);  // in flight during the barrier and the back-projection
    cluster_sync_all();  // the full dG_t of every sample is in every CTA
    if (t > 0) {
      float acc0[kLstmBT] = {}, acc1[kLstmBT] = {};
      if (bp_thread) {
#pragma unroll 4
        for (int j = jl; j < G; j += 2 * kWarps) {
          const float2 w = *reinterpret_cast<const float2*>(wk + j * Hs + 2 * kp);
          const float4 da = *reinterpret_cast<const float4*>(dcur + j * kLstmBT);
          const float4 db = *reinterpret_cast<const float4*>(dcur + j * kLstmBT + 4);
          const float dv[8] = {da.x, da.y, da.z, da.w, db.x, db.y, db.z, db.w};
#pragma unroll
          for (int sidx = 0; sidx < kLstmBT; ++sidx) {
            acc0[sidx] = fmaf(w.x, dv[sidx], acc0[sidx]);
            acc1[sidx] = fmaf(w.y, dv[sidx], acc1[sidx]);
          }
        }
      }
#pragma unroll
      for (int sidx = 0; sidx < kLstmBT; ++sidx) {
        acc0[sidx] += __shfl_xor_sync(0xffffffffu, acc0[sidx], 16);
        acc1[sidx] += __shfl_xor_sync(0xffffffffu, acc1[sidx], 16);
      }
      if (bp_thread && lane < 16) {
#pragma unroll
        for (int sidx = 0; sidx < kLstmBT; ++sidx)
          *reinterpret_cast<float2*>(part + (warp * kLstmBT + sidx) * Hs + 2 * kp) = make_float2(acc0[sidx], acc1[sidx]);
      }
      __syncthreads();
      if (fin_thread) {
        float tot = 0.f;
#pragma unroll
        for (int q = 0; q < kWarps; ++q) tot += part[(q * kLstmBT + s_f) * Hs + kl_f];
        dhn[s_f * Hs + kl_f] = tot;
      }
      __syncthreads();
    }
  }
  cluster_sync_all();  // no CTA exits while peers may still write into its shared memory
}

}  // namespace qt
