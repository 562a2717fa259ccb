// Tiled fp32 GEMM for the mid-sized linear layers of the sequence branches: the LSTM input projections over all time steps
// ([B*T, 640] x [640, 1024] in CnnLstm, cnn+lstm/models.py:43-49; [B*T, 47|188] x [., 752] in Quadtree3DCNN,
// 3dcnn/models.py:144-150), their data / weight gradients and the recurrent weight gradient dWhh = dG^T Hprev. The
// warp-per-output kernels of elementwise.cuh (written for the 47 -> 94 -> 256 head MLP at batch <= 256) re-read a weight row per
// output and ran these shapes at 3-5 TFLOP/s; here a 64 x 64 (or 32 x 32) output tile is shared by 256 threads (4 x 4 or 2 x 2 outputs each) and
// operands are staged through shared memory in K steps of 16 with the next step prefetched into registers.
// fp32 accumulation and fp32 (or bf16-stored) operands exactly as before: the reference runs these layers in fp32.
//
//   C[i][j] = sum_l A(i, l) * B(l, j)
//   A(i, l) = AK ? a[i*lda + l] : a[l*lda + i]         B(l, j) = BK ? b[j*ldb + l] : b[l*ldb + j]
//   forward   (EPI 0): i = sample, j = out feature, l = in feature: A = x (AK), B = w[n][k] (BK)
//   data grad (EPI 1): i = sample, j = in feature, l = out feature: A = dy (AK), B = w[n][k] (l-major rows: !BK)
//   weight grad (EPI 2): i = out feature, j = in feature, l = sample: A = dy (!AK), B = x (!BK); db[i] = sum_l A(i, l)
#pragma once
#include "elementwise.cuh"

namespace qt {

constexpr int kSgBK = 16, kSgThreads = 256;  // output tile = (16*TM) x (16*TM), TM x TM outputs per thread (TM = 4 or 2)

struct SgemmParams {
  const void* a;
  long long lda;
  const void* b;
  long long ldb;
  int M, N, K;
  const float* bias;        // EPI 0: [N]
  int relu;                 // EPI 0
  float drop_p;             // EPI 0 / 1: counter-hash dropout on element i*N + j
  unsigned long long seed;
  const float* act;         // EPI 1: stored post-ReLU/dropout output, gate = act > 0
  long long ldact;
  float* out;               // fp32 result (ldo), optional for EPI 0 / 1
  long long ldo;
  __nv_bfloat16* out16;     // bf16 copy (ldo16), optional
  long long ldo16;
  float* db;                // EPI 2: row sums of A, optional
  int accumulate;           // EPI 2: out += , db +=
};

template <int TM, int AK, int BK, typename TA, typename TB, int EPI>
__global__ void __launch_bounds__(kSgThreads) sgemm_tile_kernel(SgemmParams p) {
  constexpr int kT = 16 * TM;  // tile edge
  __shared__ __align__(16) float As[kSgBK][kT + 4];
  __shared__ __align__(16) float Bs[kSgBK][kT + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int i0 = blockIdx.y * kT, j0 = blockIdx.x * kT;
  const TA* a = static_cast<const TA*>(p.a);
  const TB* b = static_cast<const TB*>(p.b);
  // loader coordinates inside a tile: TM elements per thread and operand, contiguous direction fastest across threads
  int ai[TM], al[TM], bj[TM], bl[TM];
#pragma unroll
  for (int r = 0; r < TM; ++r) {
    if (AK) { al[r] = tid & 15; ai[r] = (tid >> 4) + 16 * r; } else { ai[r] = tid % kT; al[r] = tid / kT + (kSgThreads / kT) * r; }
    if (BK) { bl[r] = tid & 15; bj[r] = (tid >> 4) + 16 * r; } else { bj[r] = tid % kT; bl[r] = tid / kT + (kSgThreads / kT) * r; }
  }
  float ra[TM], rb[TM];
  auto fetch = [&](int l0) {
#pragma unroll
    for (int r = 0; r < TM; ++r) {
      const int i = i0 + ai[r], l = l0 + al[r];
      ra[r] = (i < p.M && l < p.K) ? ld_as_float<TA>(a + (AK ? i * p.lda + l : l * p.lda + i)) : 0.f;
      const int j = j0 + bj[r], lb = l0 + bl[r];
      rb[r] = (j < p.N && lb < p.K) ? ld_as_float<TB>(b + (BK ? j * p.ldb + lb : lb * p.ldb + j)) : 0.f;
    }
  };
  float acc[TM][TM] = {};
  float rowsum[TM] = {};
  fetch(0);
  for (int l0 = 0; l0 < p.K; l0 += kSgBK) {
#pragma unroll
    for (int r = 0; r < TM; ++r) {
      As[al[r]][ai[r]] = ra[r];
      Bs[bl[r]][bj[r]] = rb[r];
    }
    __syncthreads();
    if (l0 + kSgBK < p.K) fetch(l0 + kSgBK);
#pragma unroll
    for (int l = 0; l < kSgBK; ++l) {
      __align__(16) float aa[TM], bb[TM];
      if (TM == 4) {
        *reinterpret_cast<float4*>(aa) = *reinterpret_cast<const float4*>(&As[l][ty * 4]);
        *reinterpret_cast<float4*>(bb) = *reinterpret_cast<const float4*>(&Bs[l][tx * 4]);
      } else {
        *reinterpret_cast<float2*>(aa) = *reinterpret_cast<const float2*>(&As[l][ty * 2]);
        *reinterpret_cast<float2*>(bb) = *reinterpret_cast<const float2*>(&Bs[l][tx * 2]);
      }
#pragma unroll
      for (int e = 0; e < TM; ++e) {
        if (EPI == 2) rowsum[e] += aa[e];
#pragma unroll
        for (int f = 0; f < TM; ++f) acc[e][f] = fmaf(aa[e], bb[f], acc[e][f]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int e = 0; e < TM; ++e) {
    const int i = i0 + ty * TM + e;
    if (i >= p.M) continue;
    if (EPI == 2 && p.db && blockIdx.x == 0 && tx == 0) p.db[i] = p.accumulate ? p.db[i] + rowsum[e] : rowsum[e];
#pragma unroll
    for (int f = 0; f < TM; ++f) {
      const int j = j0 + tx * TM + f;
      if (j >= p.N) continue;
      float v = acc[e][f];
      if (EPI == 0) {
        if (p.bias) v += p.bias[j];
        if (p.relu) v = fmaxf(v, 0.f);
        v *= dropout_scale(p.seed, static_cast<uint32_t>(i) * p.N + j, p.drop_p);
      } else if (EPI == 1) {
        if (p.act) {
          const float g = p.act[i * p.ldact + j];
          v = (g > 0.f) ? v * dropout_scale(p.seed, static_cast<uint32_t>(i) * p.N + j, p.drop_p) : 0.f;
        }
      } else {
        if (p.accumulate) v += p.out[i * p.ldo + j];
      }
      if (p.out) p.out[i * p.ldo + j] = v;
      if (EPI != 2 && p.out16) p.out16[i * p.ldo16 + j] = __float2bfloat16_rn(v);
    }
  }
}

// The tiled kernel pays off once a 64 x 64 tile is reasonably full and there is at least a million multiply-adds.
inline bool sgemm_worthwhile(int M, int N, int K) {
  return M >= 32 && N >= 32 && K >= 16 && static_cast<long long>(M) * N * K >= (1ll << 20);
}

template <int TM, int AK, int BK, int EPI>
inline void sgemm_launch_t(const SgemmParams& p, bool a_bf16, bool b_bf16, cudaStream_t st) {
  constexpr int kT = 16 * TM;
  const dim3 grid((p.N + kT - 1) / kT, (p.M + kT - 1) / kT);
  if (a_bf16 && b_bf16) sgemm_tile_kernel<TM, AK, BK, __nv_bfloat16, __nv_bfloat16, EPI><<<grid, kSgThreads, 0, st>>>(p);
  else if (a_bf16) sgemm_tile_kernel<TM, AK, BK, __nv_bfloat16, float, EPI><<<grid, kSgThreads, 0, st>>>(p);
  else if (b_bf16) sgemm_tile_kernel<TM, AK, BK, float, __nv_bfloat16, EPI><<<grid, kSgThreads, 0, st>>>(p);
  else sgemm_tile_kernel<TM, AK, BK, float, float, EPI><<<grid, kSgThreads, 0, st>>>(p);
}
// 64 x 64 tiles when they fill the 148 SMs at least twice over, else 32 x 32 tiles (four times the CTAs: these GEMMs have only
// 256-512 rows, and a quarter-filled GPU costs more than the lower register reuse)
template <int AK, int BK, int EPI>
inline void sgemm_launch(const SgemmParams& p, bool a_bf16, bool b_bf16, cudaStream_t st) {
  const long long big = static_cast<long long>((p.N + 63) / 64) * ((p.M + 63) / 64);
  if (big >= 296) sgemm_launch_t<4, AK, BK, EPI>(p, a_bf16, b_bf16, st);
  else sgemm_launch_t<2, AK, BK, EPI>(p, a_bf16, b_bf16, st);
}

}  // namespace qt
