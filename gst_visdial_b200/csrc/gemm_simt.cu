// SIMT GEMM  C[M,N] = act(A[M,K] * W[N,K]^T + bias)  with fp32 accumulation.
//
// This is the arithmetic of the fp32 parity path (BASELINE.json config 1: "fp32 ... logits within 1e-3, greedy ids
// identical") where the tensor cores cannot be used without changing the numerics, and a debugging aid for the
// bf16 path.  It replaces the nn.Linear call sites of models/vilbert_dialog.py (:366-368, :412, :437, :454, ...)
// when the context's compute dtype is fp32.  Classic 64x64x16 tiling, 256 threads, 4x4 register micro-tile.
#include <stdexcept>

#include "common.cuh"
#include "kernels.h"

namespace gstvd {

constexpr int TM = 64, TN = 64, TK = 16;

template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256) gemm_simt_kernel(GemmArgs p) {
  __shared__ float As[TK][TM + 4];
  __shared__ float Ws[TK][TN + 4];
  const TIn* __restrict__ A = reinterpret_cast<const TIn*>(p.A);
  const TIn* __restrict__ W = reinterpret_cast<const TIn*>(p.W);
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;          // micro-tile position: rows ty*4.., cols tx*4..
  const int lr = tid / 4, lk = (tid % 4) * 4;      // loader: row lr of the tile, 4 consecutive k
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < p.K; k0 += TK) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int k = k0 + lk + j;
      int ra = m0 + lr, rw = n0 + lr;
      As[lk + j][lr] = (ra < p.M && k < p.K) ? to_f32(A[(int64_t)ra * p.lda + k]) : 0.f;
      Ws[lk + j][lr] = (rw < p.N && k < p.K) ? to_f32(W[(int64_t)rw * p.ldw + k]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      float a[4], w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) w[j] = Ws[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int row = m0 + ty * 4 + i;
    if (row >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int col = n0 + tx * 4 + j;
      if (col >= p.N) continue;
      float v = acc[i][j];
      if (p.bias) v += p.bias[col];
      if (p.act == 1) v = gelu_erf(v);
      int64_t idx;
      if (p.hm_D > 0) {
        int b = row / p.hm_L, pos = row % p.hm_L;
        int g = col / p.hm_D, d = col % p.hm_D;
        int layer = g / p.hm_G, r = g % p.hm_G;
        idx = ((((int64_t)layer * p.hm_B + b) * p.hm_G + r) * p.hm_L + pos) * p.hm_D + d;
      } else {
        idx = (int64_t)row * p.ldc + col;
      }
      reinterpret_cast<TOut*>(p.C)[idx] = from_f32<TOut>(v);
    }
  }
}

int launch_gemm_simt(const GemmArgs& a, int dtype, cudaStream_t stream) {
  if (a.M <= 0 || a.N <= 0) return 0;
  dim3 grid((a.N + TN - 1) / TN, (a.M + TM - 1) / TM);
  if (dtype == kF32) {
    gemm_simt_kernel<float, float><<<grid, 256, 0, stream>>>(a);
  } else if (a.out_f32) {
    gemm_simt_kernel<bf16, float><<<grid, 256, 0, stream>>>(a);
  } else {
    gemm_simt_kernel<bf16, bf16><<<grid, 256, 0, stream>>>(a);
  }
  return 1;
}

}  // namespace gstvd
