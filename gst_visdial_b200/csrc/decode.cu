// Decode-step kernels over the persistent KV cache (HBM-bound).
//
// The reference has no KV cache: it re-runs the whole decoder over the whole prefix and re-projects the
// cross-attention K/V on every step (use_cache=False hard-wired, models/visual_dialog_decoder.py:64; loop at
// models/visual_dialog_model.py:86-110).  Here each step processes ONE new position per beam row:
//   dec_self_attn  : appends the new K/V at position *d_step and attends over positions 0..*d_step
//   dec_cross_attn : every beam row of an image attends over that image's cross K/V (stored once per image)
//   reorder_cache  : in-place beam gather, the semantic of _reorder_cache / index_select(0, beam_idx)
//                    (models/visual_dialog_decoder.py:29-31,177-181)
// Layouts:  self cache  [layer][k|v][image][position][beam][hidden]   (a beam gather touches contiguous rows)
//           cross cache [layer][image][k|v, head][position][head_dim]  (one contiguous block per (image, head))
#include <stdexcept>

#include "common.cuh"
#include "kernels.h"

namespace gstvd {

namespace {

constexpr int kMaxSteps = 32;     // one score per lane
constexpr int kMaxBeams = 8;

template <typename T> struct Raw16 {};
template <> struct Raw16<bf16> {
  static __device__ __forceinline__ void unpack(const uint4& r, float (&o)[8]) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); o[2 * i] = f.x; o[2 * i + 1] = f.y; }
  }
};
template <> struct Raw16<float> {
  static __device__ __forceinline__ void unpack(const uint4& r, float (&o)[4]) {
    o[0] = __uint_as_float(r.x); o[1] = __uint_as_float(r.y); o[2] = __uint_as_float(r.z); o[3] = __uint_as_float(r.w);
  }
};

// Self-attention of one new position per (beam row, head) over the persistent cache (one warp per (row, head)).
// Scores: lane t owns cached position t and reads its whole key row (16-byte loads, all issued back to back) against the
// query held in shared memory - no serial chain of load + warp-reduction per position; the new position's score is one
// warp reduction.  P*V: lanes split the head dimension, the value rows of 6 positions are in flight at a time.
template <typename T>
__global__ void __launch_bounds__(128)
dec_self_attn_kernel(DecodeGeom g, int layer, const T* __restrict__ qkv, T* __restrict__ cache, const int* __restrict__ d_step,
                     T* __restrict__ out) {
  constexpr int EPL = 16 / sizeof(T);
  __shared__ __align__(16) float qs[4][128];
  pdl_wait();
  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x;
  const int h = blockIdx.y * 4 + warp;
  if (h >= g.heads) return;
  const int b = row / g.K, kb = row - b * g.K;
  const int step = *d_step;
  const int H = g.H, D = g.D;
  const int nd = D / 32;                       // dims per lane (2 for D=64, 4 for D=128)
  const T* qrow = qkv + (int64_t)row * 3 * H + h * D;
  float q[4], kn[4], vn[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    if (u < nd) {
      const int d = lane + 32 * u;
      q[u] = to_f32(qrow[d]); kn[u] = to_f32(qrow[H + d]); vn[u] = to_f32(qrow[2 * H + d]);
      qs[warp][d] = q[u];
    } else { q[u] = kn[u] = vn[u] = 0.f; }
  }
  // cache row of (kv, position t) for this beam/head
  auto cache_ptr = [&](int kv, int t) -> T* {
    return cache + (((((int64_t)layer * 2 + kv) * g.B + b) * g.T + t) * g.K + kb) * H + h * D;
  };
  {
    T* kc = cache_ptr(0, step); T* vc = cache_ptr(1, step);
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (u < nd) { kc[lane + 32 * u] = from_f32<T>(kn[u]); vc[lane + 32 * u] = from_f32<T>(vn[u]); }
  }
  __syncwarp();
  const float scale_div = sqrtf((float)D);
  // score of the new position: one reduction over the lanes' dims
  float part = 0.f;
#pragma unroll
  for (int u = 0; u < 4; ++u) part = fmaf(q[u], kn[u], part);
  const float s_new = warp_sum(part) / scale_div;
  // scores of the cached positions: lane t reads key row t
  float my_score = -INFINITY;
  if (lane < step) {
    const T* kc = cache_ptr(0, lane);
    float dot = 0.f;
    for (int c = 0; c < D / EPL; ++c) {
      float kr[EPL];
      const uint4 raw = *reinterpret_cast<const uint4*>(kc + c * EPL);
      Raw16<T>::unpack(raw, kr);
#pragma unroll
      for (int e4 = 0; e4 < EPL; e4 += 4) {
        const float4 qv = *reinterpret_cast<const float4*>(&qs[warp][c * EPL + e4]);
        dot = fmaf(qv.x, kr[e4], dot); dot = fmaf(qv.y, kr[e4 + 1], dot);
        dot = fmaf(qv.z, kr[e4 + 2], dot); dot = fmaf(qv.w, kr[e4 + 3], dot);
      }
    }
    my_score = dot / scale_div;
  } else if (lane == step) {
    my_score = s_new;
  }
  const float mx = warp_max(my_score);
  const float e = (lane <= step) ? expf(my_score - mx) : 0.f;
  const float sum = warp_sum(e);
  const float pr = e / sum;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int t0 = 0; t0 < step; t0 += 6) {
    float vv[6][4];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const int t = t0 + i;
      const T* vc = cache_ptr(1, t < step ? t : 0);
#pragma unroll
      for (int u = 0; u < 4; ++u) vv[i][u] = (u < nd && t < step) ? to_f32(vc[lane + 32 * u]) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const float pt = __shfl_sync(0xffffffffu, pr, (t0 + i) & 31);
      if (t0 + i < step) {
#pragma unroll
        for (int u = 0; u < 4; ++u) acc[u] = fmaf(pt, vv[i][u], acc[u]);
      }
    }
  }
  {
    const float pt = __shfl_sync(0xffffffffu, pr, step);
    // in bf16 mode later steps read the ROUNDED new value from the cache; use the same value now
#pragma unroll
    for (int u = 0; u < 4; ++u) acc[u] = fmaf(pt, to_f32(from_f32<T>(vn[u])), acc[u]);
  }
  T* o = out + (int64_t)row * H + h * D;
#pragma unroll
  for (int u = 0; u < 4; ++u)
    if (u < nd) o[lane + 32 * u] = from_f32<T>(acc[u]);
}

constexpr int kCrossThreads = 320;            // 10 warps: one thread per key in the score phase (Le = 293 in the real model)

template <typename T, int KB, int D>
__global__ void __launch_bounds__(kCrossThreads, 2)
dec_cross_attn_kernel(DecodeGeom g, const T* __restrict__ q, const T* __restrict__ kv_layer, const float* __restrict__ enc_mask,
                      T* __restrict__ out) {
  constexpr int EPL = 16 / sizeof(T);          // elements per 16-byte load
  constexpr int CPR = D / EPL;                 // 16-byte chunks per key row (8 for bf16, 16 for fp32)
  constexpr int LPR = CPR;                     // phase 2: lanes per key row
  constexpr int RPI = 32 / LPR;                // phase 2: key rows per warp-wide load
  constexpr int kWarps = kCrossThreads / 32;
  constexpr int NIT = 8;                       // phase 2: V loads in flight per lane (10 warps x RPI x 8 keys per pass)
  constexpr int kPass = kWarps * RPI * NIT;
  extern __shared__ float smem[];
  const int Le = g.Le;
  float* S = smem;                             // [KB][Le] scores, then probabilities
  float* part = smem + ((KB * Le + 3) & ~3);   // [kWarps][KB][D] partial outputs
  float* qs = part + kWarps * KB * D;          // [KB][D] queries (fp32), 16-byte aligned
  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x, b = blockIdx.y;
  const T* Kp = kv_layer + ((int64_t)b * 2 * g.heads + h) * Le * D;
  const T* Vp = kv_layer + ((int64_t)b * 2 * g.heads + g.heads + h) * Le * D;
  const float* mrow = enc_mask ? enc_mask + (int64_t)b * Le : nullptr;
  const uint4 zero4 = make_uint4(0, 0, 0, 0);

  // ---- phase 1: scores, one thread per key.  The whole key row (D elements) is loaded with CPR independent 16-byte
  // loads issued back to back (in flight together), the queries are broadcast from shared memory; no shuffles. ----
  const int j1 = threadIdx.x;
  uint4 kreg[CPR];
  float madd = 0.f;
  if (j1 < Le) {
#pragma unroll
    for (int c = 0; c < CPR; ++c) kreg[c] = *reinterpret_cast<const uint4*>(Kp + (int64_t)j1 * D + c * EPL);
    madd = (1.0f - (mrow ? mrow[j1] : 1.f)) * -1e9f;
  } else {
#pragma unroll
    for (int c = 0; c < CPR; ++c) kreg[c] = zero4;
  }
  pdl_wait();                                 // the K rows above were written at prefill time; q comes from the previous kernel
  for (int i = threadIdx.x; i < KB * D; i += blockDim.x) {
    const int k = i / D, d = i - k * D;
    qs[i] = (k < g.K) ? to_f32(q[((int64_t)(b * g.K + k)) * g.H + h * D + d]) : 0.f;
  }
  __syncthreads();
  const float scale_div = sqrtf((float)D);
  for (int j = j1; j < Le; j += blockDim.x) {
    if (j != j1) {                             // only when Le > blockDim.x
#pragma unroll
      for (int c = 0; c < CPR; ++c) kreg[c] = *reinterpret_cast<const uint4*>(Kp + (int64_t)j * D + c * EPL);
      madd = (1.0f - (mrow ? mrow[j] : 1.f)) * -1e9f;
    }
    float dot[KB];
#pragma unroll
    for (int k = 0; k < KB; ++k) dot[k] = 0.f;
#pragma unroll
    for (int c = 0; c < CPR; ++c) {
      float kr[EPL];
      Raw16<T>::unpack(kreg[c], kr);
#pragma unroll
      for (int k = 0; k < KB; ++k) {
#pragma unroll
        for (int e4 = 0; e4 < EPL; e4 += 4) {
          const float4 qv = *reinterpret_cast<const float4*>(qs + k * D + c * EPL + e4);   // warp-wide broadcast
          dot[k] = fmaf(qv.x, kr[e4], dot[k]); dot[k] = fmaf(qv.y, kr[e4 + 1], dot[k]);
          dot[k] = fmaf(qv.z, kr[e4 + 2], dot[k]); dot[k] = fmaf(qv.w, kr[e4 + 3], dot[k]);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < KB; ++k)
      if (k < g.K) S[k * Le + j] = dot[k] / scale_div + madd;
  }
  // V loads of the first pass: in flight while the softmax runs
  const int sub = lane / LPR, chunk = lane % LPR;
  uint4 vreg[NIT];
#pragma unroll
  for (int it = 0; it < NIT; ++it) {
    const int j = (it * kWarps + warp) * RPI + sub;
    vreg[it] = (j < Le) ? *reinterpret_cast<const uint4*>(Vp + (int64_t)j * D + chunk * EPL) : zero4;
  }
  __syncthreads();
  // ---- softmax: one warp per beam ----
  for (int k = warp; k < g.K; k += kWarps) {
    float mx = -INFINITY;
    for (int j = lane; j < Le; j += 32) mx = fmaxf(mx, S[k * Le + j]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < Le; j += 32) { const float e = expf(S[k * Le + j] - mx); S[k * Le + j] = e; sum += e; }
    sum = warp_sum(sum);
    for (int j = lane; j < Le; j += 32) S[k * Le + j] = S[k * Le + j] / sum;
  }
  __syncthreads();
  // ---- phase 2: P * V (lanes split the head dimension; every V byte is read once, coalesced) ----
  float acc[KB][EPL];
#pragma unroll
  for (int k = 0; k < KB; ++k)
#pragma unroll
    for (int e = 0; e < EPL; ++e) acc[k][e] = 0.f;
  for (int base = 0; base < Le; base += kPass) {
    if (base > 0) {
#pragma unroll
      for (int it = 0; it < NIT; ++it) {
        const int j = base + (it * kWarps + warp) * RPI + sub;
        vreg[it] = (j < Le) ? *reinterpret_cast<const uint4*>(Vp + (int64_t)j * D + chunk * EPL) : zero4;
      }
    }
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int j = base + (it * kWarps + warp) * RPI + sub;
      if (j < Le) {
        float vr[EPL];
        Raw16<T>::unpack(vreg[it], vr);
#pragma unroll
        for (int k = 0; k < KB; ++k) {
          const float p = (k < g.K) ? S[k * Le + j] : 0.f;
#pragma unroll
          for (int e = 0; e < EPL; ++e) acc[k][e] = fmaf(p, vr[e], acc[k][e]);
        }
      }
    }
  }
  // reduce over the RPI key sub-rows held by different lane groups, then over warps through shared memory
#pragma unroll
  for (int k = 0; k < KB; ++k)
#pragma unroll
    for (int e = 0; e < EPL; ++e) {
      float v = acc[k][e];
#pragma unroll
      for (int o = LPR; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      acc[k][e] = v;
    }
  if (sub == 0) {
#pragma unroll
    for (int k = 0; k < KB; ++k)
#pragma unroll
      for (int e = 0; e < EPL; ++e) part[(warp * KB + k) * D + chunk * EPL + e] = acc[k][e];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < g.K * D; i += blockDim.x) {
    const int k = i / D, d = i - k * D;
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) v += part[(w * KB + k) * D + d];
    out[((int64_t)(b * g.K + k)) * g.H + h * D + d] = from_f32<T>(v);
  }
}

template <typename T, int KB>
void launch_cross_kb(const DecodeGeom& g, const T* q, const T* kv_layer, const float* enc_mask, T* out, cudaStream_t stream) {
  const size_t s_elems = ((size_t)KB * g.Le + 3) & ~size_t(3);          // keeps the query tile 16-byte aligned
  const size_t smem = (s_elems + (size_t)(kCrossThreads / 32) * KB * 64 + (size_t)KB * 64) * sizeof(float);
  dim3 grid(g.heads, g.B);
  launch_k(dec_cross_attn_kernel<T, KB, 64>, grid, dim3(kCrossThreads), smem, stream, g, q, kv_layer, enc_mask, out);
}
template <typename T>
void launch_cross_t(const DecodeGeom& g, const void* q, const void* kv_layer, const float* enc_mask, void* out, cudaStream_t stream) {
  const T* qq = (const T*)q; const T* kv = (const T*)kv_layer; T* o = (T*)out;
  if (g.K == 1) launch_cross_kb<T, 1>(g, qq, kv, enc_mask, o, stream);
  else if (g.K == 2) launch_cross_kb<T, 2>(g, qq, kv, enc_mask, o, stream);
  else if (g.K <= 4) launch_cross_kb<T, 4>(g, qq, kv, enc_mask, o, stream);
  else if (g.K == 5) launch_cross_kb<T, 5>(g, qq, kv, enc_mask, o, stream);
  else launch_cross_kb<T, 8>(g, qq, kv, enc_mask, o, stream);
}

// One CTA per (image, layer*2+kv).  Thread-local dependency only: every thread loads the K source values of its
// column chunk before it stores any of them, so duplicated parents (beam_idx is not a permutation) are safe in place.
template <typename T>
__global__ void __launch_bounds__(256)
reorder_cache_kernel(DecodeGeom g, T* __restrict__ cache, const int32_t* __restrict__ beam_idx, const int* __restrict__ d_len,
                     int len_host, const uint8_t* __restrict__ d_skip) {
  pdl_wait();
  pdl_launch_dependents();
  const int b = blockIdx.x, lk = blockIdx.y;
  if (d_skip && d_skip[b]) return;
  const int len = d_len ? (*d_len + 1) : len_host;
  int parent[kMaxBeams];
  bool moved = false;
#pragma unroll
  for (int k = 0; k < kMaxBeams; ++k) {
    parent[k] = (k < g.K) ? beam_idx[b * g.K + k] : k;
    moved |= (parent[k] != k);
  }
  if (!moved) return;
  const int chunks = g.H / 8;
  T* base = cache + ((int64_t)lk * g.B + b) * g.T * g.K * g.H;
  for (int i = threadIdx.x; i < len * chunks; i += blockDim.x) {
    const int t = i / chunks, c = (i - t * chunks) * 8;
    T* p = base + (int64_t)t * g.K * g.H + c;
    float vals[kMaxBeams][8];
#pragma unroll
    for (int k = 0; k < kMaxBeams; ++k)
      if (k < g.K && parent[k] != k) Vec8<T>::load(p + (int64_t)parent[k] * g.H, vals[k]);
#pragma unroll
    for (int k = 0; k < kMaxBeams; ++k)
      if (k < g.K && parent[k] != k) Vec8<T>::store(p + (int64_t)k * g.H, vals[k]);
  }
}

}  // namespace

int launch_dec_self_attn(int dtype, const DecodeGeom& g, int layer, const void* qkv, void* self_cache, const int* d_step,
                         void* out, cudaStream_t stream) {
  if (g.D != 64 && g.D != 128) throw std::runtime_error("dec_self_attn: head_dim must be 64 or 128");
  if (g.T > kMaxSteps) throw std::runtime_error("dec_self_attn: at most 32 cached positions");
  dim3 grid(g.B * g.K, (g.heads + 3) / 4);
  if (dtype == kF32) launch_k(dec_self_attn_kernel<float>, grid, dim3(128), 0, stream, g, layer, (const float*)qkv, (float*)self_cache, d_step, (float*)out);
  else launch_k(dec_self_attn_kernel<bf16>, grid, dim3(128), 0, stream, g, layer, (const bf16*)qkv, (bf16*)self_cache, d_step, (bf16*)out);
  return 1;
}

int launch_dec_cross_attn(int dtype, const DecodeGeom& g, int layer, const void* q, const void* cross_cache,
                          const float* enc_mask, void* out, cudaStream_t stream) {
  const int64_t esz = dtype == kF32 ? 4 : 2;
  const int64_t per_image = (int64_t)2 * g.heads * g.Le * g.D;
  const char* kbase = (const char*)cross_cache + ((int64_t)layer * g.B) * per_image * esz;
  if (g.D == 64 && g.K <= 8 && g.Le <= 1024) {
    if (dtype == kF32) launch_cross_t<float>(g, q, kbase, enc_mask, out, stream);
    else launch_cross_t<bf16>(g, q, kbase, enc_mask, out, stream);
    return 1;
  }
  AttnArgs a;
  a.q = q; a.q_bs = g.H; a.q_hs = g.D; a.q_rs = 0;
  a.k = kbase; a.k_bs = per_image; a.k_hs = (int64_t)g.Le * g.D; a.k_rs = g.D;
  a.v = kbase + (int64_t)g.heads * g.Le * g.D * esz; a.v_bs = per_image; a.v_hs = a.k_hs; a.v_rs = g.D;
  a.o = out; a.o_bs = g.H; a.o_hs = g.D; a.o_rs = 0;
  a.kmask = enc_mask; a.kmask_bs = g.Le; a.neg = -1e9f; a.causal = 0;
  a.B = g.B * g.K; a.H = g.heads; a.Lq = 1; a.Lk = g.Le; a.D = g.D; a.kv_batch_div = g.K;
  return launch_attention_generic(a, dtype, stream);
}

int launch_reorder_cache(int dtype, const DecodeGeom& g, void* self_cache, const int32_t* beam_idx, const int* d_len,
                         int len_host, const uint8_t* d_skip, cudaStream_t stream) {
  if (g.K > kMaxBeams) throw std::runtime_error("reorder_cache: at most 8 beams");
  if (g.H % 8) throw std::runtime_error("reorder_cache: hidden % 8 != 0");
  dim3 grid(g.B, g.layers * 2);
  if (dtype == kF32) launch_k(reorder_cache_kernel<float>, grid, dim3(256), 0, stream, g, (float*)self_cache, beam_idx, d_len, len_host, d_skip);
  else launch_k(reorder_cache_kernel<bf16>, grid, dim3(256), 0, stream, g, (bf16*)self_cache, beam_idx, d_len, len_host, d_skip);
  return 1;
}

}  // namespace gstvd
