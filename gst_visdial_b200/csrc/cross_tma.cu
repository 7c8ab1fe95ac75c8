// Decode-step cross-attention, TMA-streamed (bf16, head_dim 64, <= 8 beams, Le <= 304).
//
// The K beams of an image attend over that image's encoder states (HF BertSelfAttention in cross mode, call site
// models/visual_dialog_decoder.py:300-311; mask (1-m)*-1e9, :285).  Per layer the step reads every cached cross K / V
// byte once - 10.8 MB per image, the dominant HBM stream of decoding - so the kernel is organised around
// keeping bulk copies in flight, not around the (tiny) math:
//   * two 8-warp CTAs per SM, each looping over (image, head) work items; K and V of an item arrive through TMA
//     (cp.async.bulk.tensor, 64-row boxes, 128-byte swizzle) into one 40 KB buffer each and complete on an mbarrier;
//   * the K buffer is re-armed with the NEXT item's K as soon as the scores are done, the V buffer after the context
//     product, so ~75 KB per SM are always in flight while the other CTA of the SM computes;
//   * the first item's K / V are requested before the programmatic-dependent-launch wait: the cross cache is written at
//     prefill, so the stream overlaps the tail of the preceding query-projection GEMM;
//   * only the boxes up to the image's last unmasked key are fetched (cross_len, computed at prefill): the fused mask is
//     [37 image regions | 256 history tokens] and the padded tail of the history has weight exp(-1e9) = 0 exactly.
// Math: mma.sync m16n8k16 with the beams as the (zero padded) M rows, ldmatrix from the swizzled tiles, exact two-pass
// softmax over all fetched keys (scores in shared memory) in the log2 domain.
#include <cuda.h>

#include <cstdlib>
#include <stdexcept>

#include "kernels.h"
#include "tma.cuh"

namespace gstvd {

namespace {

using namespace dev;

constexpr int kXtRows = 64;                                // key rows per TMA box
constexpr int kXtMaxBoxes = 5;
constexpr int kXtLeP = 304;                                // row stride of the score / probability tiles (19 groups of 16 keys)
constexpr int kXtBoxBytes = kXtRows * 128;                 // 64 rows x 64 bf16
constexpr int kXtBufBytes = kXtMaxBoxes * kXtBoxBytes;     // 40 KB
constexpr int kXtWarps = 8;
constexpr int kXtThreads = kXtWarps * 32;
constexpr int kXtSmem = 2 * kXtBufBytes + 8 * kXtLeP * 4 + 8 * kXtLeP * 2 + 2 * kXtLeP * 4 + 4 * 8 * 64 * 4 + 64 + 1024;
constexpr float kLog2e = 1.44269504088896340736f;

__global__ void __launch_bounds__(kXtThreads, 2)
dec_cross_tma_kernel(const __grid_constant__ CUtensorMap tm, DecodeGeom g, int layer, const bf16* __restrict__ q,
                     const float* __restrict__ enc_mask, const int* __restrict__ cross_len, bf16* __restrict__ out) {
  extern __shared__ uint8_t xt_raw[];
  const uint32_t raw = smem_u32(xt_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;             // swizzled boxes need 1024-byte alignment
  uint8_t* gen = xt_raw + (base - raw);
  const uint32_t k_buf = base, v_buf = base + kXtBufBytes;
  float* S = reinterpret_cast<float*>(gen + 2 * kXtBufBytes);            // [8][304]
  bf16* P = reinterpret_cast<bf16*>(S + 8 * kXtLeP);                     // [8][304]
  float* madd = reinterpret_cast<float*>(P + 8 * kXtLeP);                // [2][304]: this item's and the next item's mask row
  float* part = madd + 2 * kXtLeP;                                       // [4][8][64]
  const uint32_t bar_k = smem_u32(part + 4 * 8 * 64), bar_v = bar_k + 8;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int beam = lane >> 2, p4 = lane & 3, mi = lane >> 3;
  const int items = g.B * g.heads;
  if (tid == 0) {
    tma_prefetch_desc(&tm);
    mbar_init(bar_k, 1);
    mbar_init(bar_v, 1);
    fence_barrier_init();
  }
  __syncthreads();
  pdl_launch_dependents();

  auto keys_of = [&](int b) { const int n = cross_len ? cross_len[b] : g.Le; return n < 1 ? 1 : (n > g.Le ? g.Le : n); };
  // warp 0: request K (kv = 0) or V (kv = 1) of `item`; lane i fetches box i
  auto issue = [&](int item, int kv, uint32_t buf, uint32_t bar) {
    const int b = item / g.heads, h = item - b * g.heads;
    const int nb = (keys_of(b) + kXtRows - 1) / kXtRows;
    if (lane == 0) mbar_arrive_expect_tx(bar, (uint32_t)(nb * kXtBoxBytes));
    __syncwarp();
    const int z = ((layer * g.B + b) * 2 + kv) * g.heads + h;
    if (lane < nb) tma_load_3d(buf + lane * kXtBoxBytes, &tm, 0, lane * kXtRows, z, bar);
  };
  // additive mask row of image b in the log2 domain; -inf past the encoder length
  // (two entries per thread: 2 x 256 >= 304; loaded early into registers, stored once the previous reader is done)
  float mk[2];
  auto load_mask = [&](int b) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int j = tid + i * kXtThreads;
      mk[i] = (j < g.Le) ? (1.0f - (enc_mask ? enc_mask[(int64_t)b * g.Le + j] : 1.f)) * (-1e9f * kLog2e) : -INFINITY;
    }
  };
  auto store_mask = [&](float* dst) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int j = tid + i * kXtThreads;
      if (j < kXtLeP) dst[j] = mk[i];
    }
  };
  // query fragments (A operand: row = beam, natural head-dim order) straight from global
  uint32_t qa0[4], qa2[4];
  auto load_q = [&](int item) {
    const int b = item / g.heads, h = item - b * g.heads;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) { qa0[kk] = 0u; qa2[kk] = 0u; }
    if (beam < g.K) {
      const bf16* qp = q + ((int64_t)(b * g.K + beam)) * g.H + h * 64 + p4 * 2;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        qa0[kk] = *reinterpret_cast<const uint32_t*>(qp + kk * 16);
        qa2[kk] = *reinterpret_cast<const uint32_t*>(qp + kk * 16 + 8);
      }
    }
  };
  const int first = blockIdx.x;
  if (first >= items) return;
  if (warp == 0) {
    issue(first, 0, k_buf, bar_k);                          // cross K/V, cross_len and the mask date from the prefill: safe before the wait
    issue(first, 1, v_buf, bar_v);
  }
  load_mask(first / g.heads);
  store_mask(madd);
  pdl_wait();                                               // the query projection of this step is complete
  load_q(first);

  const float sc = kLog2e / 8.0f;                           // scores / sqrt(64), log2 domain
  uint32_t phase = 0;
  int cur = 0;
  for (int item = first; item < items; item += gridDim.x, phase ^= 1u, cur ^= 1) {
    const int b = item / g.heads, h = item - b * g.heads;
    const int nk = keys_of(b);
    const int ngroups = (nk + 15) >> 4, nk16 = ngroups << 4;
    const int next = item + gridDim.x;
    const float* md = madd + cur * kXtLeP;
    uint32_t a_q0[4], a_q2[4];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) { a_q0[kk] = qa0[kk]; a_q2[kk] = qa2[kk]; }
    if (next < items) load_mask(next / g.heads);            // consumed after the scores: the loads have landed by then
    __syncthreads();                                        // mask row visible; the previous item's partial sums have been read
    mbar_wait(bar_k, phase);
    // ---- scores: 16-key groups strided over the warps ----
    for (int G = warp; G < ngroups; G += kXtWarps) {
      float c0[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f};
      const int row = G * 16 + (mi >> 1) * 8 + (lane & 7);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        uint32_t kb[4];
        const int chunk = kk * 2 + (mi & 1);
        ldmatrix_x4(kb, k_buf + row * 128 + ((chunk ^ (row & 7)) << 4));
        mma_bf16_m8(c0, a_q0[kk], a_q2[kk], kb[0], kb[1]);
        mma_bf16_m8(c1, a_q0[kk], a_q2[kk], kb[2], kb[3]);
      }
      if (beam < g.K) {
        const int key = G * 16 + p4 * 2;                    // C fragment: row = beam, columns p4*2 + {0,1}
        float* sr = S + beam * kXtLeP;
        sr[key] = fmaf(c0[0], sc, md[key]);
        sr[key + 1] = fmaf(c0[1], sc, md[key + 1]);
        sr[key + 8] = fmaf(c1[0], sc, md[key + 8]);
        sr[key + 9] = fmaf(c1[1], sc, md[key + 9]);
      }
    }
    __syncthreads();                                        // scores complete, K buffer free
    if (next < items) {
      if (warp == 0) issue(next, 0, k_buf, bar_k);
      load_q(next);                                         // in flight during the softmax / context phases
      store_mask(madd + (cur ^ 1) * kXtLeP);
    }
    // ---- softmax over the fetched keys, one warp per beam; probabilities as bf16 ----
    for (int k = warp; k < g.K; k += kXtWarps) {
      float* sr = S + k * kXtLeP;
      bf16* pr = P + k * kXtLeP;
      float mx = -INFINITY;
      for (int j = lane; j < nk16; j += 32) mx = fmaxf(mx, sr[j]);
      mx = warp_max(mx);
      float sum = 0.f;
      for (int j = lane; j < nk16; j += 32) { const float e = ex2_approx(sr[j] - mx); sr[j] = e; sum += e; }
      sum = warp_sum(sum);
      const float inv = 1.0f / sum;
      for (int j = lane; j < nk16; j += 32) pr[j] = __float2bfloat16_rn(sr[j] * inv);
    }
    __syncthreads();                                        // probabilities complete
    mbar_wait(bar_v, phase);
    // ---- context: O[beam][d] += P[beam][keys of the group] V[keys][d] ----
    float o[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) { o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f; }
    for (int G = warp; G < ngroups; G += kXtWarps) {
      uint32_t a0 = 0u, a2 = 0u;
      if (beam < g.K) {
        const uint32_t* prow = reinterpret_cast<const uint32_t*>(P + beam * kXtLeP + G * 16);
        a0 = prow[p4]; a2 = prow[4 + p4];
      }
      const int row = G * 16 + (mi & 1) * 8 + (lane & 7);
#pragma unroll
      for (int dt = 0; dt < 8; dt += 2) {
        uint32_t vb[4];
        const int chunk = dt + (mi >> 1);
        ldmatrix_x4_trans(vb, v_buf + row * 128 + ((chunk ^ (row & 7)) << 4));
        mma_bf16_m8(o[dt], a0, a2, vb[0], vb[1]);
        mma_bf16_m8(o[dt + 1], a0, a2, vb[2], vb[3]);
      }
    }
    // cross-warp reduction in two rounds over a [4][8][64] tile: warps 4..7 publish, warps 0..3 add their own and publish
    float* pp = part + ((warp & 3) * 8 + beam) * 64 + p4 * 2;
    if (warp >= 4) {
#pragma unroll
      for (int dt = 0; dt < 8; ++dt) { pp[dt * 8] = o[dt][0]; pp[dt * 8 + 1] = o[dt][1]; }
    }
    __syncthreads();                                        // V buffer free (every warp is past its context product)
    if (warp == 0 && next < items) issue(next, 1, v_buf, bar_v);
    if (warp < 4) {
#pragma unroll
      for (int dt = 0; dt < 8; ++dt) { pp[dt * 8] += o[dt][0]; pp[dt * 8 + 1] += o[dt][1]; }
    }
    __syncthreads();
    for (int i = tid; i < g.K * 64; i += kXtThreads) {
      const int k = i >> 6, d = i & 63;
      const float v = (part[(0 * 8 + k) * 64 + d] + part[(1 * 8 + k) * 64 + d]) + (part[(2 * 8 + k) * 64 + d] + part[(3 * 8 + k) * 64 + d]);
      out[((int64_t)(b * g.K + k)) * g.H + h * 64 + d] = __float2bfloat16_rn(v);
    }
  }
}

// cross_len[b] = 1 + index of the last key whose mask is non-zero (Le when the whole row is masked: the reference then
// spreads uniform weight over every key, which needs them all).
__global__ void cross_len_kernel(int B, int Le, const float* __restrict__ mask, int* __restrict__ out) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (b >= B) return;
  int last = -1;
  for (int j = lane; j < Le; j += 32)
    if (mask[(int64_t)b * Le + j] != 0.f) last = j;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
  if (lane == 0) out[b] = last < 0 ? Le : last + 1;
}

}  // namespace

bool dec_cross_tma_supported(int dtype, const DecodeGeom& g) {
  static const bool off = getenv("GSTVD_CROSS_NO_TMA") != nullptr;
  return !off && dtype == kBF16 && g.D == 64 && g.K <= 8 && g.Le <= kXtLeP && g.H == g.heads * 64;
}

int launch_dec_cross_tma(const DecodeGeom& g, int layer, const void* q, const void* cross_cache, const float* enc_mask,
                         const int* cross_len, void* out, int num_sms, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(dec_cross_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kXtSmem);
    if (e != cudaSuccess) throw std::runtime_error(std::string("dec_cross_tma: ") + cudaGetErrorString(e));
    configured = true;
  }
  const CUtensorMap* tm = reinterpret_cast<const CUtensorMap*>(
      tma_map_rows3(cross_cache, 64, g.Le, (int64_t)g.layers * g.B * 2 * g.heads, kXtRows));
  const int items = g.B * g.heads;
  const int grid = items < 2 * num_sms ? items : 2 * num_sms;
  launch_k(dec_cross_tma_kernel, dim3(grid), dim3(kXtThreads), (size_t)kXtSmem, stream, *tm, g, layer, (const bf16*)q, enc_mask, cross_len,
           (bf16*)out);
  return 1;
}

int launch_cross_len(int B, int Le, const float* mask, int* out, cudaStream_t stream) {
  cross_len_kernel<<<(B + 3) / 4, 128, 0, stream>>>(B, Le, mask, out);
  return 1;
}

}  // namespace gstvd
