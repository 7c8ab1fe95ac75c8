// Tensor-core attention for the encoder / teacher-forced shapes (bf16 operands, fp32 softmax and accumulation).
//
// Shapes served (SURVEY.md 2.2 "A"): text self-attention 256x256 d64 (models/vilbert_dialog.py:385-407), image
// self-attention 37x37 d128 (:512-534), co-attention 256q x 37k and 37q x 256k d128 (:671-710), and the teacher-forced
// decoder's causal self-attention (L x L d64) and cross-attention (L x 293 d64).  These tiles are far too small to
// amortise a TMEM allocation + tcgen05 pipeline per (sample, head) - the whole K/V of a head fits in shared memory - so
// the kernel stages Q (64 rows), K and V of one (sample, head) in shared memory once and runs mma.sync m16n8k16 with an
// online softmax over 64-key tiles.  Attention is ~5 % of the encoder FLOPs; the dense projections run on tcgen05
// (gemm_tc.cu).  Masks are additive like the reference ((1-m)*neg); keys past Lk are excluded.
#include <stdexcept>

#include "common.cuh"
#include "kernels.h"

namespace gstvd {

namespace {

constexpr int kQTile = 64;      // query rows per CTA (16 per warp)
constexpr int kKTile = 64;      // keys per online-softmax step

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

template <int D>
__global__ void __launch_bounds__(128) attention_mma_kernel(AttnArgs p, int lk_pad) {
  constexpr int LD = D + 8;                      // padded row (16 bytes) -> conflict-free ldmatrix
  extern __shared__ __align__(16) uint8_t smem_raw[];
  bf16* Qs = reinterpret_cast<bf16*>(smem_raw);
  bf16* Ks = Qs + kQTile * LD;
  bf16* Vs = Ks + (size_t)lk_pad * LD;
  float* madd = reinterpret_cast<float*>(Vs + (size_t)lk_pad * LD);

  pdl_wait();
  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * kQTile;
  const int bkv = b / p.kv_batch_div;
  const bf16* Q = reinterpret_cast<const bf16*>(p.q) + (int64_t)b * p.q_bs + (int64_t)h * p.q_hs;
  const bf16* K = reinterpret_cast<const bf16*>(p.k) + (int64_t)bkv * p.k_bs + (int64_t)h * p.k_hs;
  const bf16* V = reinterpret_cast<const bf16*>(p.v) + (int64_t)bkv * p.v_bs + (int64_t)h * p.v_hs;
  bf16* O = reinterpret_cast<bf16*>(p.o) + (int64_t)b * p.o_bs + (int64_t)h * p.o_hs;
  const float* km = p.kmask ? p.kmask + (int64_t)bkv * p.kmask_bs : nullptr;

  // ---- stage Q / K / V (16-byte chunks, zero fill past the valid rows) ----
  constexpr int CH = D / 8;
  const uint4 zero = make_uint4(0, 0, 0, 0);
  for (int i = threadIdx.x; i < kQTile * CH; i += blockDim.x) {
    const int r = i / CH, c = (i % CH) * 8;
    uint4 v = zero;
    if (q0 + r < p.Lq) v = *reinterpret_cast<const uint4*>(Q + (int64_t)(q0 + r) * p.q_rs + c);
    *reinterpret_cast<uint4*>(Qs + r * LD + c) = v;
  }
  for (int i = threadIdx.x; i < lk_pad * CH; i += blockDim.x) {
    const int r = i / CH, c = (i % CH) * 8;
    uint4 kv = zero, vv = zero;
    if (r < p.Lk) {
      kv = *reinterpret_cast<const uint4*>(K + (int64_t)r * p.k_rs + c);
      vv = *reinterpret_cast<const uint4*>(V + (int64_t)r * p.v_rs + c);
    }
    *reinterpret_cast<uint4*>(Ks + (size_t)r * LD + c) = kv;
    *reinterpret_cast<uint4*>(Vs + (size_t)r * LD + c) = vv;
  }
  for (int j = threadIdx.x; j < lk_pad; j += blockDim.x)
    madd[j] = (j < p.Lk) ? (1.0f - (km ? km[j] : 1.0f)) * p.neg : -INFINITY;
  __syncthreads();

  const int qrow0 = warp * 16;                    // this warp's 16 query rows inside the tile
  if (q0 + qrow0 >= p.Lq) return;                 // whole warp idle (no further block-level sync below)
  const uint32_t qs_base = static_cast<uint32_t>(__cvta_generic_to_shared(Qs));
  const uint32_t ks_base = static_cast<uint32_t>(__cvta_generic_to_shared(Ks));
  const uint32_t vs_base = static_cast<uint32_t>(__cvta_generic_to_shared(Vs));

  uint32_t qa[D / 16][4];
#pragma unroll
  for (int kk = 0; kk < D / 16; ++kk) {
    const int r = qrow0 + (lane & 15), c = kk * 16 + (lane >> 4) * 8;
    ldmatrix_x4(qa[kk], qs_base + (uint32_t)(r * LD + c) * 2u);
  }
  float o[D / 8][4];
#pragma unroll
  for (int i = 0; i < D / 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  const float scale_div = sqrtf((float)D);
  const int row_a = q0 + qrow0 + (lane >> 2), row_b = row_a + 8;   // global query indices of this thread's two rows

  for (int kt = 0; kt < lk_pad; kt += kKTile) {
    float s[kKTile / 8][4];
#pragma unroll
    for (int i = 0; i < kKTile / 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
#pragma unroll
    for (int kk = 0; kk < D / 16; ++kk) {
#pragma unroll
      for (int nt = 0; nt < kKTile / 8; nt += 2) {
        uint32_t kb[4];
        const int mi = lane >> 3;
        const int r = kt + nt * 8 + (mi >> 1) * 8 + (lane & 7), c = kk * 16 + (mi & 1) * 8;
        ldmatrix_x4(kb, ks_base + (uint32_t)(r * LD + c) * 2u);
        mma_bf16(s[nt], qa[kk], kb[0], kb[1]);
        mma_bf16(s[nt + 1], qa[kk], kb[2], kb[3]);
      }
    }
    // scale, additive mask, running max
    float tmax[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < kKTile / 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int j = kt + nt * 8 + (lane & 3) * 2 + (e & 1);
        const int qi = (e < 2) ? row_a : row_b;
        float add = madd[j];
        if (p.causal && j > qi && j < p.Lk) add = p.neg;
        const float v = s[nt][e] / scale_div + add;
        s[nt][e] = v;
        tmax[e >> 1] = fmaxf(tmax[e >> 1], v);
      }
    }
    float corr[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      tmax[r] = fmaxf(tmax[r], __shfl_xor_sync(0xffffffffu, tmax[r], 1));
      tmax[r] = fmaxf(tmax[r], __shfl_xor_sync(0xffffffffu, tmax[r], 2));
      const float m_new = fmaxf(m_run[r], tmax[r]);
      corr[r] = (m_run[r] == -INFINITY) ? 0.f : expf(m_run[r] - m_new);
      m_run[r] = m_new;
    }
    float tsum[2] = {0.f, 0.f};
#pragma unroll
    for (int nt = 0; nt < kKTile / 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float pv = expf(s[nt][e] - m_run[e >> 1]);     // exp(-inf) = 0 for keys past Lk
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
      uint32_t pa[4];
      pa[0] = pack_bf16(s[2 * jb][0], s[2 * jb][1]);
      pa[1] = pack_bf16(s[2 * jb][2], s[2 * jb][3]);
      pa[2] = pack_bf16(s[2 * jb + 1][0], s[2 * jb + 1][1]);
      pa[3] = pack_bf16(s[2 * jb + 1][2], s[2 * jb + 1][3]);
#pragma unroll
      for (int dt = 0; dt < D / 8; dt += 2) {
        uint32_t vb[4];
        const int mi = lane >> 3;
        const int r = kt + jb * 16 + (mi & 1) * 8 + (lane & 7), c = dt * 8 + (mi >> 1) * 8;
        ldmatrix_x4_trans(vb, vs_base + (uint32_t)(r * LD + c) * 2u);
        mma_bf16(o[dt], pa, vb[0], vb[1]);
        mma_bf16(o[dt + 1], pa, vb[2], vb[3]);
      }
    }
  }
  // ---- normalise and store ----
  const float inv_a = 1.0f / l_run[0], inv_b = 1.0f / l_run[1];
#pragma unroll
  for (int dt = 0; dt < D / 8; ++dt) {
    const int c = dt * 8 + (lane & 3) * 2;
    if (row_a < p.Lq) *reinterpret_cast<uint32_t*>(O + (int64_t)row_a * p.o_rs + c) = pack_bf16(o[dt][0] * inv_a, o[dt][1] * inv_a);
    if (row_b < p.Lq) *reinterpret_cast<uint32_t*>(O + (int64_t)row_b * p.o_rs + c) = pack_bf16(o[dt][2] * inv_b, o[dt][3] * inv_b);
  }
}

template <int D>
void launch_d(const AttnArgs& a, cudaStream_t stream) {
  const int lk_pad = (a.Lk + kKTile - 1) / kKTile * kKTile;
  const size_t smem = ((size_t)kQTile + 2 * (size_t)lk_pad) * (D + 8) * 2 + (size_t)lk_pad * 4;
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(attention_mma_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) throw std::runtime_error(std::string("attention_mma: ") + cudaGetErrorString(e));
    configured = 200 * 1024;
  }
  dim3 grid((a.Lq + kQTile - 1) / kQTile, a.H, a.B);
  launch_k(attention_mma_kernel<D>, grid, dim3(128), smem, stream, a, lk_pad);
}

}  // namespace

bool attention_mma_supported(const AttnArgs& a) {
  if (a.D != 64 && a.D != 128) return false;
  const int lk_pad = (a.Lk + kKTile - 1) / kKTile * kKTile;
  const size_t smem = ((size_t)kQTile + 2 * (size_t)lk_pad) * (a.D + 8) * 2 + (size_t)lk_pad * 4;
  if (smem > 200 * 1024) return false;
  auto al = [](int64_t v) { return v % 8 == 0; };
  return al(a.q_bs) && al(a.q_hs) && al(a.q_rs) && al(a.k_bs) && al(a.k_hs) && al(a.k_rs) && al(a.v_bs) && al(a.v_hs) && al(a.v_rs) &&
         (a.o_bs % 2 == 0) && (a.o_hs % 2 == 0) && (a.o_rs % 2 == 0) &&
         (reinterpret_cast<uintptr_t>(a.q) % 16 == 0) && (reinterpret_cast<uintptr_t>(a.k) % 16 == 0) &&
         (reinterpret_cast<uintptr_t>(a.v) % 16 == 0) && (reinterpret_cast<uintptr_t>(a.o) % 4 == 0);
}

int launch_attention_mma(const AttnArgs& a, cudaStream_t stream) {
  if (a.B <= 0 || a.Lq <= 0) return 0;
  if (!attention_mma_supported(a)) throw std::runtime_error("attention_mma: unsupported shape / alignment");
  if (a.D == 64) launch_d<64>(a, stream);
  else launch_d<128>(a, stream);
  return 1;
}

}  // namespace gstvd
