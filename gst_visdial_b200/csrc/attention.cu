// Exact-softmax attention  softmax(q k^T / sqrt(D) + (1 - m) * neg) v  for arbitrary (Lq, Lk <= 512, D <= 128).
//
// Generic SIMT kernel (fp32 math, any activation dtype).  It is the attention of the fp32 parity path and the
// reference implementation the tensor-core kernel in attention_mma.cu is tested against.  Call sites replaced:
// text / image self-attention (models/vilbert_dialog.py:385-407, :512-534), both co-attention directions
// (:671-710), decoder causal self-attention and cross-attention (HF BertSelfAttention, call site
// models/visual_dialog_decoder.py:300-311).  Masks are ADDITIVE exactly like the reference ((1-m)*-10000 for
// self/co-attention, :1352-1370, and (1-m)*-1e9 for decoder cross-attention), so a fully masked row degrades to
// the same uniform distribution the reference produces.  Scores are divided by sqrt(D) (not multiplied by a
// reciprocal) as at :391,:518,:672,:692.
#include <stdexcept>

#include "common.cuh"
#include "kernels.h"

namespace gstvd {

namespace {

constexpr int kWarps = 4;
constexpr int kRowsPerCta = 16;
constexpr int kMaxLk = 512;
constexpr int kMaxD = 128;

template <typename T>
__global__ void __launch_bounds__(kWarps * 32) attention_generic_kernel(AttnArgs p) {
  __shared__ float ss[kWarps][kMaxLk];
  __shared__ float qs[kWarps][kMaxD];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, h = blockIdx.y;
  const int bkv = b / p.kv_batch_div;
  const T* __restrict__ Q = reinterpret_cast<const T*>(p.q) + (int64_t)b * p.q_bs + (int64_t)h * p.q_hs;
  const T* __restrict__ K = reinterpret_cast<const T*>(p.k) + (int64_t)bkv * p.k_bs + (int64_t)h * p.k_hs;
  const T* __restrict__ V = reinterpret_cast<const T*>(p.v) + (int64_t)bkv * p.v_bs + (int64_t)h * p.v_hs;
  T* __restrict__ O = reinterpret_cast<T*>(p.o) + (int64_t)b * p.o_bs + (int64_t)h * p.o_hs;
  const float* __restrict__ km = p.kmask ? p.kmask + (int64_t)bkv * p.kmask_bs : nullptr;
  const float scale_div = sqrtf((float)p.D);

  for (int r = warp; r < kRowsPerCta; r += kWarps) {
    const int i = blockIdx.x * kRowsPerCta + r;
    if (i >= p.Lq) break;
    for (int d = lane; d < p.D; d += 32) qs[warp][d] = to_f32(Q[(int64_t)i * p.q_rs + d]);
    __syncwarp();
    float mx = -INFINITY;
    for (int j = lane; j < p.Lk; j += 32) {
      const T* kr = K + (int64_t)j * p.k_rs;
      float dot = 0.f;
      for (int d0 = 0; d0 < p.D; d0 += 8) {
        float kv[8];
        Vec8<T>::load(kr + d0, kv);
#pragma unroll
        for (int u = 0; u < 8; ++u) dot = fmaf(qs[warp][d0 + u], kv[u], dot);
      }
      float m = km ? km[j] : 1.f;
      if (p.causal && j > i) m = 0.f;
      const float s = dot / scale_div + (1.0f - m) * p.neg;
      ss[warp][j] = s;
      mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < p.Lk; j += 32) {
      const float e = expf(ss[warp][j] - mx);
      ss[warp][j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    __syncwarp();
    for (int d = lane; d < p.D; d += 32) {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      int j = 0;
      for (; j + 4 <= p.Lk; j += 4) {
        a0 = fmaf(ss[warp][j] / sum, to_f32(V[(int64_t)j * p.v_rs + d]), a0);
        a1 = fmaf(ss[warp][j + 1] / sum, to_f32(V[(int64_t)(j + 1) * p.v_rs + d]), a1);
        a2 = fmaf(ss[warp][j + 2] / sum, to_f32(V[(int64_t)(j + 2) * p.v_rs + d]), a2);
        a3 = fmaf(ss[warp][j + 3] / sum, to_f32(V[(int64_t)(j + 3) * p.v_rs + d]), a3);
      }
      for (; j < p.Lk; ++j) a0 = fmaf(ss[warp][j] / sum, to_f32(V[(int64_t)j * p.v_rs + d]), a0);
      O[(int64_t)i * p.o_rs + d] = from_f32<T>((a0 + a1) + (a2 + a3));
    }
    __syncwarp();
  }
}

}  // namespace

int launch_attention_generic(const AttnArgs& a, int dtype, cudaStream_t stream) {
  if (a.B <= 0 || a.Lq <= 0) return 0;
  if (a.Lk > kMaxLk || a.D > kMaxD || a.D % 8 != 0) throw std::runtime_error("attention_generic: needs Lk <= 512, D <= 128, D % 8 == 0");
  dim3 grid((a.Lq + kRowsPerCta - 1) / kRowsPerCta, a.H, a.B);
  if (dtype == kF32) attention_generic_kernel<float><<<grid, kWarps * 32, 0, stream>>>(a);
  else attention_generic_kernel<bf16><<<grid, kWarps * 32, 0, stream>>>(a);
  return 1;
}

}  // namespace gstvd
