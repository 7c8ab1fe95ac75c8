// Tensor-core attention for the encoder / teacher-forced shapes (bf16 operands, fp32 softmax and accumulation).
//
// Shapes served (SURVEY.md 2.2 "A"): text self-attention 256x256 d64 (models/vilbert_dialog.py:385-407), image
// self-attention 37x37 d128 (:512-534), co-attention 256q x 37k and 37q x 256k d128 (:671-710), and the teacher-forced
// decoder's causal self-attention (L x L d64) and cross-attention (L x 293 d64).  These tiles are far too small to
// amortise a TMEM allocation + tcgen05 pipeline per (sample, head), so the kernel runs mma.sync m16n8k16 with an online
// softmax: one CTA per (sample, head, 16*QW query rows); K / V stream through a two-stage cp.async ring of 64-key tiles
// (the next tile lands while the current one is multiplied), Q is staged once.
//   * Key tiles past the last unmasked key are skipped.  This is exact: a masked key carries the additive -10000 (or
//     -1e9), whose exp underflows to exactly 0 in fp32 next to any unmasked key; a row with no
//     unmasked key at all takes the full path, which reproduces the reference's uniform weights.
//   * The softmax runs in the log2 domain (scores * log2(e)/sqrt(d), ex2.approx) - about 6 instructions per score
//     instead of ~20 with an IEEE divide and expf; the ALU work of the softmax, not the MMAs, was the issue-slot bound.
// Attention is ~5 % of the encoder FLOPs; the dense projections run on tcgen05 (gemm_tc.cu).
#include <stdexcept>

#include "common.cuh"
#include "kernels.h"

namespace gstvd {

namespace {

constexpr int kKTile = 64;      // keys per online-softmax step
constexpr float kLog2e = 1.44269504088896340736f;

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 16-byte global -> shared copy without a register round trip; bytes == 0 zero-fills the destination
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int D, int QW>
__global__ void __launch_bounds__(QW * 32) attention_mma_kernel(AttnArgs p, int lk_pad) {
  constexpr int LD = D + 8;                      // padded row (16 bytes) -> conflict-free ldmatrix
  constexpr int QT = QW * 16, NT = QW * 32, CH = D / 8;
  constexpr int kStageElems = kKTile * LD;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  bf16* Qs = reinterpret_cast<bf16*>(smem_raw);
  bf16* Ks = Qs + QT * LD;                       // [2][64][LD]
  bf16* Vs = Ks + 2 * kStageElems;               // [2][64][LD]
  float* madd = reinterpret_cast<float*>(Vs + 2 * kStageElems);
  __shared__ int s_last;

  pdl_wait();
  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * QT;
  const int bkv = b / p.kv_batch_div;
  const bf16* Q = reinterpret_cast<const bf16*>(p.q) + (int64_t)b * p.q_bs + (int64_t)h * p.q_hs;
  const bf16* K = reinterpret_cast<const bf16*>(p.k) + (int64_t)bkv * p.k_bs + (int64_t)h * p.k_hs;
  const bf16* V = reinterpret_cast<const bf16*>(p.v) + (int64_t)bkv * p.v_bs + (int64_t)h * p.v_hs;
  bf16* O = reinterpret_cast<bf16*>(p.o) + (int64_t)b * p.o_bs + (int64_t)h * p.o_hs;
  const float* km = p.kmask ? p.kmask + (int64_t)bkv * p.kmask_bs : nullptr;
  const uint32_t qs_base = static_cast<uint32_t>(__cvta_generic_to_shared(Qs));
  const uint32_t ks_base = static_cast<uint32_t>(__cvta_generic_to_shared(Ks));
  const uint32_t vs_base = static_cast<uint32_t>(__cvta_generic_to_shared(Vs));

  auto load_kv = [&](int t, int st) {
    for (int i = threadIdx.x; i < kKTile * CH; i += NT) {
      const int r = i / CH, c = (i % CH) * 8;
      const int key = t * kKTile + r;
      const int bytes = key < p.Lk ? 16 : 0;
      const int ksafe = key < p.Lk ? key : 0;
      const uint32_t off = (uint32_t)(st * kStageElems + r * LD + c) * 2u;
      cp_async16(ks_base + off, K + (int64_t)ksafe * p.k_rs + c, bytes);
      cp_async16(vs_base + off, V + (int64_t)ksafe * p.v_rs + c, bytes);
    }
  };
  // ---- Q tile + first K/V tile in flight while the mask row is scanned ----
  for (int i = threadIdx.x; i < QT * CH; i += NT) {
    const int r = i / CH, c = (i % CH) * 8;
    const bool ok = q0 + r < p.Lq;
    cp_async16(qs_base + (uint32_t)(r * LD + c) * 2u, Q + (int64_t)(ok ? q0 + r : 0) * p.q_rs + c, ok ? 16 : 0);
  }
  load_kv(0, 0);
  cp_async_commit();
  if (threadIdx.x == 0) s_last = -1;
  __syncthreads();
  int last = -1;
  for (int j = threadIdx.x; j < lk_pad; j += NT) {
    float a = -INFINITY;
    if (j < p.Lk) {
      const float m = km ? km[j] : 1.0f;
      a = (1.0f - m) * p.neg * kLog2e;
      if (m != 0.f) last = j;
    }
    madd[j] = a;
  }
  if (last >= 0) atomicMax(&s_last, last);
  __syncthreads();
  int key_end = s_last < 0 ? p.Lk : s_last + 1;                    // keys [key_end, Lk) are all masked: their weight is exactly 0
  if (p.causal) key_end = min(key_end, min(p.Lk, q0 + QT));        // keys past the last query of the tile are masked for every row
  const int n_tiles = (key_end + kKTile - 1) / kKTile;
  if (n_tiles > 1) load_kv(1, 1);
  cp_async_commit();

  const int qrow0 = warp * 16;                    // this warp's 16 query rows inside the tile
  const bool active = q0 + qrow0 < p.Lq;          // idle warps still take part in the copies and barriers
  uint32_t qa[D / 16][4];
  float o[D / 8][4];
#pragma unroll
  for (int i = 0; i < D / 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  const float sc = kLog2e / sqrtf((float)D);
  const float neg2 = p.neg * kLog2e;
  const int row_a = q0 + qrow0 + (lane >> 2), row_b = row_a + 8;   // global query indices of this thread's two rows

  for (int t = 0; t < n_tiles; ++t) {
    cp_async_wait<1>();                           // everything but the most recent group has landed: tile t (and Q)
    __syncthreads();
    if (active) {
      const int st = t & 1, kt = t * kKTile;
      const int nkeys = min(kKTile, key_end - kt);                 // warp-uniform; 8-key blocks past it are skipped
      const uint32_t ks_t = ks_base + (uint32_t)(st * kStageElems) * 2u, vs_t = vs_base + (uint32_t)(st * kStageElems) * 2u;
      if (t == 0) {
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
          const int r = qrow0 + (lane & 15), c = kk * 16 + (lane >> 4) * 8;
          ldmatrix_x4(qa[kk], qs_base + (uint32_t)(r * LD + c) * 2u);
        }
      }
      float s[kKTile / 8][4];
#pragma unroll
      for (int i = 0; i < kKTile / 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
#pragma unroll
      for (int nt = 0; nt < kKTile / 8; nt += 2) {
        if (nt * 8 < nkeys) {
#pragma unroll
          for (int kk = 0; kk < D / 16; ++kk) {
            uint32_t kb[4];
            const int mi = lane >> 3;
            const int r = nt * 8 + (mi >> 1) * 8 + (lane & 7), c = kk * 16 + (mi & 1) * 8;
            ldmatrix_x4(kb, ks_t + (uint32_t)(r * LD + c) * 2u);
            mma_bf16(s[nt], qa[kk], kb[0], kb[1]);
            mma_bf16(s[nt + 1], qa[kk], kb[2], kb[3]);
          }
        }
      }
      // scale (log2 domain), additive mask, running max
      float tmax[2] = {-INFINITY, -INFINITY};
#pragma unroll
      for (int nt = 0; nt < kKTile / 8; ++nt) {
        const int j0 = kt + nt * 8 + (lane & 3) * 2;
        const float2 ad = *reinterpret_cast<const float2*>(madd + j0);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float add = (e & 1) ? ad.y : ad.x;
          if (p.causal) {
            const int j = j0 + (e & 1), qi = (e < 2) ? row_a : row_b;
            if (j > qi && j < p.Lk) add = neg2;
          }
          const float v = fmaf(s[nt][e], sc, add);
          s[nt][e] = v;
          tmax[e >> 1] = fmaxf(tmax[e >> 1], v);
        }
      }
      float corr[2];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        tmax[r] = fmaxf(tmax[r], __shfl_xor_sync(0xffffffffu, tmax[r], 1));
        tmax[r] = fmaxf(tmax[r], __shfl_xor_sync(0xffffffffu, tmax[r], 2));
        const float m_new = fmaxf(m_run[r], tmax[r]);             // finite: every tile holds at least one key < Lk
        corr[r] = ex2_approx(m_run[r] - m_new);                   // 2^(-inf) = 0 on the first tile
        m_run[r] = m_new;
      }
      float tsum[2] = {0.f, 0.f};
#pragma unroll
      for (int nt = 0; nt < kKTile / 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float pv = ex2_approx(s[nt][e] - m_run[e >> 1]);   // 2^(-inf) = 0 for keys past Lk
          s[nt][e] = pv;
          tsum[e >> 1] += pv;
        }
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        tsum[r] += __shfl_xor_sync(0xffffffffu, tsum[r], 1);
        tsum[r] += __shfl_xor_sync(0xffffffffu, tsum[r], 2);
        l_run[r] = l_run[r] * corr[r] + tsum[r];
      }
#pragma unroll
      for (int i = 0; i < D / 8; ++i) { o[i][0] *= corr[0]; o[i][1] *= corr[0]; o[i][2] *= corr[1]; o[i][3] *= corr[1]; }
      // O += P * V
#pragma unroll
      for (int jb = 0; jb < kKTile / 16; ++jb) {
        if (jb * 16 < nkeys) {
          uint32_t pa[4];
          pa[0] = pack_bf16(s[2 * jb][0], s[2 * jb][1]);
          pa[1] = pack_bf16(s[2 * jb][2], s[2 * jb][3]);
          pa[2] = pack_bf16(s[2 * jb + 1][0], s[2 * jb + 1][1]);
          pa[3] = pack_bf16(s[2 * jb + 1][2], s[2 * jb + 1][3]);
#pragma unroll
          for (int dt = 0; dt < D / 8; dt += 2) {
            uint32_t vb[4];
            const int mi = lane >> 3;
            const int r = jb * 16 + (mi & 1) * 8 + (lane & 7), c = dt * 8 + (mi >> 1) * 8;
            ldmatrix_x4_trans(vb, vs_t + (uint32_t)(r * LD + c) * 2u);
            mma_bf16(o[dt], pa, vb[0], vb[1]);
            mma_bf16(o[dt + 1], pa, vb[2], vb[3]);
          }
        }
      }
    }
    __syncthreads();                              // every warp is done with stage t & 1
    if (t + 2 < n_tiles) load_kv(t + 2, t & 1);
    cp_async_commit();                            // (possibly empty) keeps the group count uniform
  }
  cp_async_wait<0>();
  if (!active) return;
  // ---- normalise and store ----
  const float inv_a = 1.0f / l_run[0], inv_b = 1.0f / l_run[1];
#pragma unroll
  for (int dt = 0; dt < D / 8; ++dt) {
    const int c = dt * 8 + (lane & 3) * 2;
    if (row_a < p.Lq) *reinterpret_cast<uint32_t*>(O + (int64_t)row_a * p.o_rs + c) = pack_bf16(o[dt][0] * inv_a, o[dt][1] * inv_a);
    if (row_b < p.Lq) *reinterpret_cast<uint32_t*>(O + (int64_t)row_b * p.o_rs + c) = pack_bf16(o[dt][2] * inv_b, o[dt][3] * inv_b);
  }
}

template <int D, int QW>
size_t attn_smem(int lk_pad) { return ((size_t)QW * 16 + 4 * (size_t)kKTile) * (D + 8) * 2 + (size_t)lk_pad * 4; }

template <int D, int QW>
void launch_d(const AttnArgs& a, cudaStream_t stream) {
  const int lk_pad = (a.Lk + kKTile - 1) / kKTile * kKTile;
  const size_t smem = attn_smem<D, QW>(lk_pad);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attention_mma_kernel<D, QW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    if (e != cudaSuccess) throw std::runtime_error(std::string("attention_mma: ") + cudaGetErrorString(e));
    configured = true;
  }
  dim3 grid((a.Lq + QW * 16 - 1) / (QW * 16), a.H, a.B);
  launch_k(attention_mma_kernel<D, QW>, grid, dim3(QW * 32), smem, stream, a, lk_pad);
}

}  // namespace

bool attention_mma_supported(const AttnArgs& a) {
  if (a.D != 64 && a.D != 128) return false;
  if (a.Lk < 1 || a.Lk > 16384) return false;                     // the mask row (4 bytes per key) lives in shared memory
  auto al = [](int64_t v) { return v % 8 == 0; };
  return al(a.q_bs) && al(a.q_hs) && al(a.q_rs) && al(a.k_bs) && al(a.k_hs) && al(a.k_rs) && al(a.v_bs) && al(a.v_hs) && al(a.v_rs) &&
         (a.o_bs % 2 == 0) && (a.o_hs % 2 == 0) && (a.o_rs % 2 == 0) &&
         (reinterpret_cast<uintptr_t>(a.q) % 16 == 0) && (reinterpret_cast<uintptr_t>(a.k) % 16 == 0) &&
         (reinterpret_cast<uintptr_t>(a.v) % 16 == 0) && (reinterpret_cast<uintptr_t>(a.o) % 4 == 0);
}

int launch_attention_mma(const AttnArgs& a, cudaStream_t stream) {
  if (a.B <= 0 || a.Lq <= 0) return 0;
  if (!attention_mma_supported(a)) throw std::runtime_error("attention_mma: unsupported shape / alignment");
  if (a.D == 64) {
    // Query tile = 16 * QW rows.  Big tiles stream K / V through shared memory less often, but a tile that sticks out past Lq is
    // wasted math: the dialog loop trims the text to multiples of 32 tokens, so e.g. Lq = 160 is two exact 80-row tiles (QW = 5)
    // instead of 128 + 32 of 128.  Pick the QW in 4..8 with the fewest padded rows, the larger one on ties.
    int best = 4;
    if (a.Lq > 64) {
      int best_pad = 1 << 30;
      for (int qw = 8; qw >= 4; --qw) {
        const int t = qw * 16, pad = (a.Lq + t - 1) / t * t - a.Lq;
        if (pad < best_pad) { best_pad = pad; best = qw; }
      }
    }
    switch (best) {
      case 8: launch_d<64, 8>(a, stream); break;
      case 7: launch_d<64, 7>(a, stream); break;
      case 6: launch_d<64, 6>(a, stream); break;
      case 5: launch_d<64, 5>(a, stream); break;
      default: launch_d<64, 4>(a, stream); break;
    }
  } else {
    launch_d<128, 4>(a, stream);
  }
  return 1;
}

}  // namespace gstvd
