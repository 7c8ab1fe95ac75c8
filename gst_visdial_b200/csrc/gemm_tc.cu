// tcgen05 / TMEM / TMA GEMM for sm_100a:  C[M,N] = act(A[M,K] * W[N,K]^T + bias), bf16 operands, fp32 accumulate.
//
// Replaces every nn.Linear on the hot path when the compute dtype is bf16: the q/k/v, output and FFN projections of
// the text / image / connection layers (models/vilbert_dialog.py:366-368,412,437,454,493-495,539,564,581,624-633,
// 718,725,753-757), the image embedding (:1415), VLFusion (models/visual_dialog_model.py:127-128), the decoder
// layers and the LM head (models/visual_dialog_decoder.py:300-311,329-339).
//
// Design (persistent over output tiles, warp specialised; one CTA per SM, two in the shared-SM decode configuration):
//   warp 0      : TMA producer  - cp.async.bulk.tensor loads of the A (128 rows) and W (BN rows) k-blocks into a shared-
//                 memory ring (128-byte swizzle), completion on mbarriers.  Wide tiles move one 64-element k-chunk per
//                 2-D box; the narrow decode tiles move several chunks per 3-D box (one thread issues a TMA operation
//                 every ~0.13 us).  The weights of the first ring are requested before griddepcontrol.wait (PDL).
//   warp 1      : MMA issuer    - one lane issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN, K=16) x4 per k-chunk
//                 into one of two TMEM accumulator stages; tcgen05.commit releases ring slots / signals the epilogue
//   warps 2..   : epilogue      - 16 warps (4 per TMEM lane quadrant; 4 in the shared-SM configuration): tcgen05.ld
//                 (32 lanes x 32 columns), + bias, GELU, convert, then either a swizzled shared-memory tile + TMA store
//                 (throughput problems), direct 16-byte row stores (skinny decode problems) or the generic
//                 transpose-through-shared-memory path (outputs a tensor map cannot describe).
//                 Two accumulator stages let the epilogue of tile i overlap the main loop of tile i+1.
// Both operands are K-major ("TN"), which is the layout of activations [rows, features] and nn.Linear weights
// [out, in], so no transposes are ever materialised.  TMA zero-fills out-of-range rows / k, so M, N, K need not be
// multiples of the tile (K % 8 == 0 for the 16-byte stride rule).
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <stdexcept>
#include <string>
#include <unordered_map>

#include "common.cuh"
#include "kernels.h"

namespace gstvd {

namespace {

constexpr int BM = 128;          // rows per tile  (UMMA M)
constexpr int BK = 64;           // k per stage: 64 bf16 = 128 bytes = one swizzle row
constexpr int UMMA_K = 16;
// Epilogue warps: 16 (4 per TMEM lane quadrant) for the throughput configurations - the epilogue is ALU/latency bound and
// needs the warps; 4 (one per quadrant) for the "skinny" decode configuration (64-row tiles, see TileCfg), whose 6-warp CTA
// with ~100 KB of shared memory lets TWO CTAs share an SM: decode GEMMs (M = 320) are latency bound, so a co-resident CTA -
// the next GEMM of the same stream under PDL, or another stream's - fills the SM time this one spends waiting.
constexpr int kEpiWarpsWide = 16;
constexpr int kEpiWarpsSkinny = 4;
constexpr unsigned long long kWaitTimeoutNs = 4000000000ull;   // 4 s: far beyond any legitimate wait

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0, spins = 0;
  unsigned long long t0 = 0;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!ok && (++spins & 1023u) == 0) {         // a protocol bug must fail the launch, never hang the GPU
      unsigned long long now = global_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > kWaitTimeoutNs) __trap();
    }
  } while (!ok);
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "n"(COLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}

// K-major operand tile in shared memory, 128-byte swizzle: rows of 64 bf16 (128 B); 8-row groups are 1024 B apart.
// Descriptor fields (PTX ISA "tcgen05 shared memory descriptor"): start address >> 4 [0,14), leading byte offset >> 4
// [16,30) (unused for swizzled K-major, set to 1), stride byte offset >> 4 [32,46) = 1024 >> 4, version 1 at [46,48),
// layout type SWIZZLE_128B = 2 at [61,64).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// Instruction descriptor for kind::f16: D fp32 (bits 4-5 = 1), A/B bf16 (bits 7-9 / 10-12 = 1), both K-major
// (bits 15, 16 = 0), N >> 3 at [17,23), M >> 4 at [24,29).
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 32 consecutive 32-bit columns: thread i of the warp receives lane (base_lane + i), columns c..c+31
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

template <int BN, int EW> struct TileCfg {
  static constexpr bool kSkinny = EW == kEpiWarpsSkinny;
  static constexpr int kThreads = (2 + EW) * 32;
  // Rows per tile = UMMA M.  The skinny (decode, M = 320) configuration uses M = 64: five exact row tiles instead of three
  // 128-row tiles with 17 % padding, 40 % fewer bytes through the L2->SM port per CTA ((64+BN) instead of (128+BN) rows per
  // k-step) on more SMs, and a CTA small enough for two per SM.  The M = 64 accumulator occupies lanes 0-15 of every
  // 32-lane TMEM quadrant: row r lives in lane (r / 16) * 32 + r % 16 (tools/probes/umma_m64_layout.cu).
  static constexpr int kBM = kSkinny ? 64 : BM;
  // k-chunks (64 elements each) per pipeline stage.  The narrow tiles serve the M <= 512 decode GEMMs, whose main loop is
  // bound by the RATE of TMA operations issued by one thread (~0.13 us each), not by bytes: one 3-D box
  // {64, rows, chunks} moves several k-blocks per operation.
  static constexpr int kCK = kSkinny ? (BN <= 32 ? 4 : 2) : (BN <= 64 ? 4 : 1);
  static constexpr int kStages = kSkinny ? (BN <= 32 ? 2 : 3) : (BN == 256 ? 4 : (BN == 128 ? 6 : 2));
  static constexpr int kAChunk = kBM * BK * 2;                 // one kBM-row x 64-element SW128 tile
  static constexpr int kBChunk = BN * BK * 2;
  static constexpr int kABytes = kAChunk * kCK;
  static constexpr int kBBytes = kBChunk * kCK;
  static constexpr int kTmemCols = 2 * BN < 32 ? 32 : 2 * BN;   // power of two for BN in {16,32,64,128,256}
  static constexpr int kBarBytes = 256;
  static constexpr int kStageWords = 32 * 33;                   // per epilogue warp: 32 rows x 32 words, padded rows
  // wide: 33 KB = 4 x 8 KB (bf16) / 2 x 16 KB (fp32) TMA-store tiles, or 8 generic tiles; skinny: none (direct row stores only)
  static constexpr int kStagingBytes = kSkinny ? 0 : 8 * kStageWords * 4;
  static constexpr int kSmemBytes = kStages * (kABytes + kBBytes) + kBarBytes + kStagingBytes + 1024;   // +1024: alignment slack
  static_assert(!kSkinny || 2 * (kSmemBytes + 1024) <= 233472, "skinny configuration must fit two CTAs per SM");
};

// GELU(x) = 0.5 x (1 + erf(x / sqrt 2)) with erf from Abramowitz-Stegun 7.1.25 (|error| <= 2.5e-5, two orders of magnitude below
// bf16 resolution) and approximate MUFU reciprocal / exp2: ~15 instructions per element instead of erff's ~35, so the GELU
// epilogue of the FFN1 GEMM hides behind the MMA main loop.  (The fp32 parity path uses the exact erff in gemm_simt.cu.)
__device__ __forceinline__ float gelu_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.47047f, z, 1.0f)));
  float poly = fmaf(t, 0.7478556f, -0.0958798f);
  poly = fmaf(t, poly, 0.3480242f);
  poly *= t;
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.44269504088896340736f * z * z));
  const float erf_abs = fmaf(-poly, e, 1.0f);
  const float erf = copysignf(erf_abs, x);
  return 0.5f * x * (1.0f + erf);
}

__device__ __forceinline__ int64_t out_index(const GemmArgs& p, int row, int col) {
  if (p.hm_D > 0) {
    const int b = row / p.hm_L, pos = row - b * p.hm_L;
    const int g = col / p.hm_D, d = col - g * p.hm_D;
    const int layer = g / p.hm_G, r = g - layer * p.hm_G;
    return ((((int64_t)layer * p.hm_B + b) * p.hm_G + r) * p.hm_L + pos) * p.hm_D + d;
  }
  return (int64_t)row * p.ldc + col;
}

// Epilogue of one W-column block of this warp's 32 accumulator rows: TMEM -> registers (thread = row) -> + bias,
// activation, convert -> transpose through a warp-private padded shared-memory tile -> global stores in which the 32
// lanes of one instruction cover CONTIGUOUS bytes of one (or two) output rows (128 B segments), instead of 32 different
// rows.  The uncoalesced variant (one 16-byte store per row per instruction) was L2-transaction bound: 3x slower.
template <typename OutT, int W>
__device__ __forceinline__ void epilogue_block(const GemmArgs& p, uint32_t taddr, uint32_t* stage, int row0, int col0, int lane) {
  constexpr bool kBf16 = sizeof(OutT) == 2;
  constexpr int kWords = kBf16 ? W / 2 : W;       // 32-bit words per output row in this block
  static_assert(kWords <= 32, "block too wide for the staging tile");
  const bool full = (col0 + W <= p.N);
  const bool bias_vec = p.bias != nullptr && full && ((reinterpret_cast<uintptr_t>(p.bias + col0) & 15) == 0);
#pragma unroll
  for (int part = 0; part < W / 32; ++part) {
    uint32_t acc[32];
    tmem_ld32(taddr + part * 32, acc);
#pragma unroll
    for (int g8 = 0; g8 < 4; ++g8) {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(acc[g8 * 8 + j]);
      const int cb = col0 + part * 32 + g8 * 8;
      if (p.bias) {
        if (bias_vec) {
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + cb));
          const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + cb + 4));
          v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w; v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (cb + j < p.N) v[j] += __ldg(p.bias + cb + j);
        }
      }
      if (p.act == 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = gelu_fast(v[j]);
      }
      if constexpr (kBf16) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
          stage[lane * 33 + part * 16 + g8 * 4 + j] = *reinterpret_cast<uint32_t*>(&h);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) stage[lane * 33 + part * 32 + g8 * 8 + j] = __float_as_uint(v[j]);
      }
    }
  }
  __syncwarp();
  constexpr int kRowsPerInstr = 32 / kWords;
  constexpr int kColsPerWord = kBf16 ? 2 : 1;
  const int w = lane % kWords, rsub = lane / kWords;
  const int col = col0 + w * kColsPerWord;
  const bool ld_even = kBf16 ? ((p.hm_D > 0) || ((p.ldc & 1) == 0)) : true;
#pragma unroll 4
  for (int i0 = 0; i0 < 32; i0 += kRowsPerInstr) {
    const int r = i0 + rsub;
    const int row = row0 + r;
    const uint32_t word = stage[r * 33 + w];
    if (row < p.M && col < p.N && p.dbg != 1) {
      OutT* dst = reinterpret_cast<OutT*>(p.C) + out_index(p, row, col);
      if constexpr (kBf16) {
        if (col + 1 < p.N && ld_even) {
          *reinterpret_cast<uint32_t*>(dst) = word;
        } else {
          const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&word);
          dst[0] = h.x;
          if (col + 1 < p.N) dst[1] = h.y;
        }
      } else {
        *reinterpret_cast<uint32_t*>(dst) = word;
      }
    }
  }
  __syncwarp();
}

// Epilogue of one CW-column block for the 128 rows of a tile, through a TMA store.  The 4 warps of a half-group (128
// threads, thread = accumulator row) convert their row to the output type, write it into a 128-row staging tile in the
// TMA swizzle pattern (16-byte chunk index XOR row bits -> conflict-free st.shared.v4), and one thread issues
// cp.async.bulk.tensor (global <- shared), which clips at the M / N edges.  ~2 instructions per output element instead of
// ~28 for the register/shared transpose with per-row address arithmetic, so 8 epilogue warps keep up with the MMA.
template <typename OutT, int CW>
__device__ __forceinline__ void epilogue_tma_block(const GemmArgs& p, const CUtensorMap* tm_c, uint32_t taddr, uint32_t stage_addr,
                                                   int tile_row0, int col0, int r, int grp, bool issuer) {
  constexpr int kRowBytes = CW * (int)sizeof(OutT);     // 64 or 128
  constexpr int kChunks = kRowBytes / 16;
  constexpr int kWordsRow = kRowBytes / 4;
  if (issuer) bulk_wait_read0();                        // the previous store out of this staging tile has been read
  named_bar_sync(1 + grp, 128);
  uint32_t packed[kWordsRow];
  const bool full = (col0 + CW <= p.N);
  const bool bias_vec = p.bias != nullptr && full && ((reinterpret_cast<uintptr_t>(p.bias + col0) & 15) == 0);
#pragma unroll
  for (int part = 0; part < CW / 32; ++part) {
    uint32_t acc[32];
    tmem_ld32(taddr + part * 32, acc);
#pragma unroll
    for (int g8 = 0; g8 < 4; ++g8) {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(acc[g8 * 8 + j]);
      const int cb = col0 + part * 32 + g8 * 8;
      if (p.bias) {
        if (bias_vec) {
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + cb));
          const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + cb + 4));
          v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w; v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (cb + j < p.N) v[j] += __ldg(p.bias + cb + j);
        }
      }
      if (p.act == 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = gelu_fast(v[j]);
      }
      if constexpr (sizeof(OutT) == 2) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
          packed[part * 16 + g8 * 4 + j] = *reinterpret_cast<uint32_t*>(&h);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) packed[part * 32 + g8 * 8 + j] = __float_as_uint(v[j]);
      }
    }
  }
  const int sw = (kRowBytes == 128) ? (r & 7) : ((r >> 1) & 3);
  const uint32_t row_addr = stage_addr + r * kRowBytes;
  if (p.dbg == 5) {                                     // measurement: keep the math alive without touching shared memory
    uint32_t x = 0;
#pragma unroll
    for (int c = 0; c < kWordsRow; ++c) x ^= packed[c];
    if (x == 0x12345678u) st_shared_v4(row_addr, x, x, x, x);
  } else {
#pragma unroll
    for (int c = 0; c < kChunks; ++c)
      st_shared_v4(row_addr + ((c ^ sw) << 4), packed[4 * c], packed[4 * c + 1], packed[4 * c + 2], packed[4 * c + 3]);
  }
  if (p.dbg != 5) fence_proxy_async();                  // generic-proxy writes -> visible to the async (TMA) proxy
  named_bar_sync(1 + grp, 128);
  if (issuer && p.dbg != 1 && p.dbg != 5) {
    if (p.hm_D > 0) {
      // head-major scatter: column block = one (layer, k|v, head); the M tiles of this mode never straddle two images
      // (hm_tpi tiles per image, rows past hm_L are clipped by the store - TMA stores reject negative coordinates)
      const int g = col0 / p.hm_D, layer = g / p.hm_G, rr = g - layer * p.hm_G;
      const int m_blk = tile_row0 / BM;
      const int b = m_blk / p.hm_tpi, pos0 = (m_blk - b * p.hm_tpi) * BM;
      tma_store_3d(tm_c, stage_addr, col0 - g * p.hm_D, pos0, (layer * p.hm_B + b) * p.hm_G + rr);
    } else {
      tma_store_2d(tm_c, stage_addr, col0, tile_row0);
    }
    bulk_commit();
  }
}

// Epilogue of one 32-column block straight from registers: thread = accumulator row, 64 (bf16) or 128 (fp32) contiguous
// bytes per row as 16-byte stores.  Uncoalesced across the warp, which costs L2 transactions on big outputs (that is why
// the throughput configurations stage through shared memory + TMA), but for the skinny decode GEMMs the whole tile is 8 KB
// and what matters is latency: no staging tile, no proxy fence, no group barrier, no wait for the TMA engine at the end.
template <typename OutT>
__device__ __forceinline__ void epilogue_direct_block(const GemmArgs& p, uint32_t taddr, int row, int col0, const float* bpre) {
  constexpr int kWords = 32 * (int)sizeof(OutT) / 4;
  uint32_t acc[32];
  tmem_ld32(taddr, acc);
  uint32_t packed[kWords];
  const bool full = (col0 + 32 <= p.N);
  const bool bias_vec = p.bias != nullptr && full && ((reinterpret_cast<uintptr_t>(p.bias + col0) & 15) == 0);
#pragma unroll
  for (int g8 = 0; g8 < 4; ++g8) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(acc[g8 * 8 + j]);
    const int cb = col0 + g8 * 8;
    if (bpre != nullptr) {                               // bias values fetched while the main loop was running
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += bpre[g8 * 8 + j];
    } else if (p.bias) {
      if (bias_vec) {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + cb));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + cb + 4));
        v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w; v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (cb + j < p.N) v[j] += __ldg(p.bias + cb + j);
      }
    }
    if (p.act == 1) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = gelu_fast(v[j]);
    }
    if constexpr (sizeof(OutT) == 2) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
        packed[g8 * 4 + j] = *reinterpret_cast<uint32_t*>(&h);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) packed[g8 * 8 + j] = __float_as_uint(v[j]);
    }
  }
  if (row >= p.M || p.dbg == 1) return;
  OutT* dst = reinterpret_cast<OutT*>(p.C) + (int64_t)row * p.ldc + col0;
  if (full) {                                            // launch_cfg guarantees 16-byte aligned rows in this mode
#pragma unroll
    for (int c = 0; c < kWords / 4; ++c)
      *reinterpret_cast<uint4*>(reinterpret_cast<uint32_t*>(dst) + 4 * c) = make_uint4(packed[4 * c], packed[4 * c + 1], packed[4 * c + 2], packed[4 * c + 3]);
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      if (col0 + j < p.N) {
        if constexpr (sizeof(OutT) == 2) {
          const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&packed[j >> 1]);
          dst[j] = (j & 1) ? h.y : h.x;
        } else {
          dst[j] = __uint_as_float(packed[j]);
        }
      }
    }
  }
}

template <int BN, int EW>
__global__ void __launch_bounds__((2 + EW) * 32, EW == kEpiWarpsSkinny ? 2 : 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
               const __grid_constant__ CUtensorMap tm_c, const GemmArgs p) {
  using Cfg = TileCfg<BN, EW>;
  constexpr int kStages = Cfg::kStages;
  constexpr int kGroups = EW / 4;                               // epilogue groups of 4 warps (one warp per TMEM lane quadrant)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;                 // 128B-swizzle atoms need 1024-byte alignment
  const uint32_t a_base = base;
  const uint32_t b_base = base + kStages * Cfg::kABytes;
  const uint32_t stage_base = b_base + kStages * Cfg::kBBytes;   // epilogue staging (1024-byte aligned: TMA-store source)
  const uint32_t bar_base = stage_base + Cfg::kStagingBytes;
  // barrier layout: full[kStages] | empty[kStages] | tmem_full[2] | tmem_empty[2] | tmem base address (u32)
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * kStages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * kStages + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 4);
  uint8_t* smem_gen = smem_raw + (base - raw);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned long long* stamps = p.dbg_times ? p.dbg_times + (size_t)blockIdx.x * 8 : nullptr;
  if (stamps && threadIdx.x == 0) stamps[0] = global_ns();
  const int tiles_n = (p.N + BN - 1) / BN;
  constexpr int kBM = Cfg::kBM;
  const int tiles_m = p.hm_tpi > 0 ? p.hm_B * p.hm_tpi : (p.M + kBM - 1) / kBM;
  const int num_tiles = tiles_m * tiles_n;
  constexpr int kCK = Cfg::kCK;
  const int num_kb = (p.K + BK * kCK - 1) / (BK * kCK);      // pipeline stages per tile

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    if (p.tma_store == 1) tma_prefetch_desc(&tm_c);
    for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), EW * 32); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  if (stamps && threadIdx.x == 0) stamps[1] = global_ns();

  if (warp == 0) {
    // ---------------- TMA producer ----------------
    if (lane == 0) {
      pdl_launch_dependents();                       // the next kernel may start its prologue / weight prefetch
      int stage = 0; uint32_t phase = 0;
      bool first = true;
      auto load_b = [&](int st, int kb, int n_blk) {
        if constexpr (kCK > 1) tma_load_3d(b_base + st * Cfg::kBBytes, &tm_b, 0, n_blk * BN, kb * kCK, full_bar(st));
        else tma_load_2d(b_base + st * Cfg::kBBytes, &tm_b, kb * BK, n_blk * BN, full_bar(st));
      };
      auto load_a = [&](int st, int kb, int m_blk) {
        if constexpr (kCK > 1) {
          tma_load_3d(a_base + st * Cfg::kABytes, &tm_a, 0, m_blk * kBM, kb * kCK, full_bar(st));
        } else if (p.hm_tpi > 0) {
          const int img = m_blk / p.hm_tpi;
          tma_load_3d(a_base + st * Cfg::kABytes, &tm_a, kb * BK, (m_blk - img * p.hm_tpi) * BM, img, full_bar(st));
        } else {
          tma_load_2d(a_base + st * Cfg::kABytes, &tm_a, kb * BK, m_blk * BM, full_bar(st));
        }
      };
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / tiles_n, n_blk = tile - m_blk * tiles_n;
        int kb0 = 0;
        if (first) {
          // Programmatic dependent launch: the weights (B) never depend on the previous kernel, so the first ring of
          // B tiles is requested before waiting for it; the activations (A) are requested after the wait.
          first = false;
          const int pre = num_kb < kStages ? num_kb : kStages;
          for (int st = 0; st < pre; ++st) {
            mbar_arrive_expect_tx(full_bar(st), Cfg::kABytes + Cfg::kBBytes);
            load_b(st, st, n_blk);
          }
          pdl_wait();
          if (stamps) stamps[2] = global_ns();
          for (int st = 0; st < pre; ++st) load_a(st, st, m_blk);
          kb0 = pre;
          if (pre == kStages) { stage = 0; phase = 1u; } else { stage = pre; }
        }
        for (int kb = kb0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          mbar_arrive_expect_tx(full_bar(stage), Cfg::kABytes + Cfg::kBBytes);
          load_a(stage, kb, m_blk);
          load_b(stage, kb, n_blk);
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(kBM, BN);
      int stage = 0; uint32_t phase = 0; int iter = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++iter) {
        const int as = iter & 1; const uint32_t aphase = (iter >> 1) & 1;
        mbar_wait(tempty_bar(as), aphase ^ 1u);      // epilogue has drained this accumulator stage
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);         // TMA bytes have landed
          tc_fence_after();
          if (stamps && iter == 0 && kb == 0) stamps[3] = global_ns();
#pragma unroll
          for (int ck = 0; ck < kCK; ++ck) {
            const uint64_t a_desc = make_smem_desc(a_base + stage * Cfg::kABytes + ck * Cfg::kAChunk);
            const uint64_t b_desc = make_smem_desc(b_base + stage * Cfg::kBBytes + ck * Cfg::kBChunk);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              // advance 16 elements (32 bytes) along K inside the 128-byte swizzle row: +2 in the (addr >> 4) field
              if (p.dbg != 3) umma_f16(d_tmem, a_desc + 2u * k, b_desc + 2u * k, idesc, (kb | ck | k) != 0 ? 1u : 0u);
            }
          }
          umma_commit(empty_bar(stage));             // frees the smem slot when these MMAs retire
          if (kb == num_kb - 1) { umma_commit(tfull_bar(as)); if (stamps) stamps[4] = global_ns(); }
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else {
    // ---------------- epilogue ----------------
    const int e = warp - 2;
    const int quad = warp & 3;                        // TMEM lane quadrant this warp may access
    const int grp = e >> 2;                           // 4 groups of 4 warps (one warp per quadrant = 128 accumulator rows)
    const int half = grp;                             // the generic (non-TMA) path only uses groups 0 and 1
    constexpr bool kSplit = BN >= 64 && kGroups >= 2;       // generic path: two groups share the columns of a wide tile
    constexpr int kColsPerHalf = kSplit ? BN / 2 : BN;
    const int c_begin = kSplit ? half * kColsPerHalf : 0;
    const int c_end = (grp >= 2) ? 0 : (kSplit ? c_begin + kColsPerHalf : (half == 0 ? BN : 0));
    uint32_t* stage = reinterpret_cast<uint32_t*>(smem_gen + (stage_base - base)) + (e & 7) * Cfg::kStageWords;
    int iter = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++iter) {
      const int m_blk = tile / tiles_n, n_blk = tile - m_blk * tiles_n;
      const int as = iter & 1; const uint32_t aphase = (iter >> 1) & 1;
      if (p.bias != nullptr && lane == 0) {              // pull this tile's bias segment into L1 while the main loop runs
        const int c0 = n_blk * BN + (e & 7) * (BN / 8);
        if (c0 < p.N) asm volatile("prefetch.global.L1 [%0];" ::"l"(p.bias + c0));
      }
      // direct-store (decode) mode: the bias of this warp's first column block is fetched now, under the main loop
      float bpre[32];
      bool have_bpre = false;
      if (p.tma_store == 2 && p.bias != nullptr && grp < BN / 32) {
        const int c0 = n_blk * BN + grp * 32;
        if (c0 + 32 <= p.N && (reinterpret_cast<uintptr_t>(p.bias + c0) & 15) == 0) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + c0) + j);
            bpre[4 * j] = b4.x; bpre[4 * j + 1] = b4.y; bpre[4 * j + 2] = b4.z; bpre[4 * j + 3] = b4.w;
          }
          have_bpre = true;
        }
      }
      mbar_wait(tfull_bar(as), aphase);
      tc_fence_after();
      if (stamps && e == 0 && lane == 0) stamps[5] = global_ns();
      const int row0 = m_blk * BM + quad * 32;
      if (p.dbg == 2) { tc_fence_before(); mbar_arrive(tempty_bar(as)); continue; }
      constexpr int kWb = (BN >= 128) ? 64 : 32;           // block width of the generic (non-TMA) bf16 path
      if (p.tma_store == 2) {
        const uint32_t tq = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * BN;
        for (int j = grp; j < BN / 32; j += kGroups) {
          const int col0 = n_blk * BN + j * 32;
          if (col0 >= p.N) break;
          const float* bp = (have_bpre && j == grp) ? bpre : nullptr;
          // M = 128: TMEM lane == tile row.  M = 64: lanes 0-15 of each quadrant hold rows quad * 16 + lane, the rest nothing.
          const int row = kBM == 128 ? row0 + lane : (lane < 16 ? m_blk * kBM + quad * 16 + lane : p.M);
          if (p.out_f32) epilogue_direct_block<float>(p, tq + j * 32, row, col0, bp);
          else epilogue_direct_block<bf16>(p, tq + j * 32, row, col0, bp);
        }
      } else if (p.tma_store) {
        const int r = quad * 32 + lane;                     // accumulator row == TMEM lane == staging row
        const bool issuer = (e & 3) == 0 && lane == 0;
        const uint32_t tq = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * BN;
        if (p.out_f32) {
          // fp32 rows of 32 columns are 128 bytes: two 16 KB staging tiles, groups 0 and 1 only
          constexpr int kFGroups = kGroups < 2 ? kGroups : 2;
          if (grp < kFGroups) {
            const uint32_t stg = stage_base + grp * 16384;
            for (int j = grp; j < BN / 32; j += kFGroups) {
              const int col0 = n_blk * BN + j * 32;
              if (col0 >= p.N) break;                       // uniform across the group
              epilogue_tma_block<float, 32>(p, &tm_c, tq + j * 32, stg, m_blk * BM, col0, r, grp, issuer);
            }
          }
        } else {
          // bf16 rows of 32 columns are 64 bytes: four 8 KB staging tiles, one per group
          const uint32_t stg = stage_base + grp * 8192;
          for (int j = grp; j < BN / 32; j += kGroups) {
            const int col0 = n_blk * BN + j * 32;
            if (col0 >= p.N) break;
            epilogue_tma_block<bf16, 32>(p, &tm_c, tq + j * 32, stg, m_blk * BM, col0, r, grp, issuer);
          }
        }
      } else if (p.out_f32) {
        for (int c = c_begin; c < c_end; c += 32) {
          const int col0 = n_blk * BN + c;
          if (col0 >= p.N) break;                           // warp-uniform
          epilogue_block<float, 32>(p, tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * BN + c, stage, row0, col0, lane);
        }
      } else {
        for (int c = c_begin; c < c_end; c += kWb) {
          const int col0 = n_blk * BN + c;
          if (col0 >= p.N) break;
          epilogue_block<bf16, kWb>(p, tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * BN + c, stage, row0, col0, lane);
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar(as));
    }
    // the staging tiles must have been read before the CTA's shared memory goes away; the global writes themselves are
    // complete (and visible to the next kernel) at grid end like any other store
    if (p.tma_store == 1 && (e & 3) == 0 && lane == 0) bulk_wait_read0();
    if (stamps && e == 0 && lane == 0) stamps[6] = global_ns();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  if (stamps && threadIdx.x == 32) stamps[7] = global_ns();
}

// ------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
std::once_flag g_once;

struct MapKey {
  const void* ptr; int64_t rows, cols, ld; int box_rows, box_cols, esz, kind; int64_t d2, d3;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows && box_cols == o.box_cols &&
           esz == o.esz && kind == o.kind && d2 == o.d2 && d3 == o.d3;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    for (int64_t v : {k.rows, k.cols, k.ld, (int64_t)k.box_rows, (int64_t)k.box_cols, (int64_t)k.esz, (int64_t)k.kind, k.d2, k.d3})
      h = h * 1000003u ^ static_cast<size_t>(v);
    return h;
  }
};
std::mutex g_map_mu;
std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_maps;

const CUtensorMap& cached_map(const MapKey& key, CUtensorMapDataType dt, int rank, const cuuint64_t* gdim, const cuuint64_t* gstride,
                              const cuuint32_t* box, CUtensorMapSwizzle swz) {
  std::lock_guard<std::mutex> lk(g_map_mu);
  auto it = g_maps.find(key);
  if (it != g_maps.end()) return it->second;
  CUtensorMap m;
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = g_encode(&m, dt, rank, const_cast<void*>(key.ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
  if (g_maps.size() > 4096) g_maps.clear();
  return g_maps.emplace(key, m).first->second;
}

// 2-D bf16 operand [rows, cols] with row stride ld (elements); box = 64 columns x box_rows rows, 128-byte swizzle.
const CUtensorMap& get_map(const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || (ld * 2) % 16 != 0)
    throw std::runtime_error("gemm_tc: operand must be 16-byte aligned with a row stride that is a multiple of 8 elements");
  MapKey key{ptr, rows, cols, ld, box_rows, BK, 2, 0, 0, 0};
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld * 2)};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(BK), static_cast<cuuint32_t>(box_rows)};
  return cached_map(key, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, gdim, gstride, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

// Operand viewed as [K/64 chunks][rows][64]: one box = 64 elements x box_rows rows x `chunks` consecutive k-chunks, each chunk
// landing as its own 128-byte-swizzled tile (row stride 128 B) - the layout the UMMA descriptors expect.
const CUtensorMap& get_map_k3(const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows, int chunks) {
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || (ld * 2) % 16 != 0 || cols % BK != 0)
    throw std::runtime_error("gemm_tc: chunked operand needs 16-byte alignment, ld % 8 == 0 and K % 64 == 0");
  MapKey key{ptr, rows, cols, ld, box_rows, BK, 2, 5, chunks, 0};
  cuuint64_t gdim[3] = {(cuuint64_t)BK, (cuuint64_t)rows, (cuuint64_t)(cols / BK)};
  cuuint64_t gstride[2] = {(cuuint64_t)ld * 2, (cuuint64_t)BK * 2};
  cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)box_rows, (cuuint32_t)chunks};
  return cached_map(key, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, gdim, gstride, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

// Output map for the TMA-store epilogue: [M, N] row-major with box = box_cols x 128 rows.
const CUtensorMap& get_map_c(const void* ptr, int64_t rows, int64_t cols, int64_t ld, int esz, int box_cols) {
  MapKey key{ptr, rows, cols, ld, BM, box_cols, esz, 1, 0, 0};
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld * esz)};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(BM)};
  return cached_map(key, esz == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, gdim, gstride, box,
                    box_cols * esz == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
}

// Head-major cross-KV output [layers*B][G][L][D] bf16 (GemmArgs::hm_*): box = D x 128 positions of one (layer-image, group).
// Activations of the head-major mode viewed as [B][L][K]: box = 64 k x 128 positions of one image (rows past L are zero-filled).
const CUtensorMap& get_map_a3(const void* ptr, int64_t K, int L, int B, int64_t ld) {
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || (ld * 2) % 16 != 0)
    throw std::runtime_error("gemm_tc: operand must be 16-byte aligned with a row stride that is a multiple of 8 elements");
  MapKey key{ptr, L, K, ld, BM, BK, 2, 4, B, 0};
  cuuint64_t gdim[3] = {(cuuint64_t)K, (cuuint64_t)L, (cuuint64_t)B};
  cuuint64_t gstride[2] = {(cuuint64_t)ld * 2, (cuuint64_t)L * ld * 2};
  cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)BM, 1};
  return cached_map(key, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, gdim, gstride, box, CU_TENSOR_MAP_SWIZZLE_128B);
}
const CUtensorMap& get_map_hm3(const void* ptr, int D, int L, int64_t LBG, int box_cols) {
  MapKey key{ptr, L, D, D, BM, box_cols, 2, 3, LBG, 0};
  cuuint64_t gdim[3] = {(cuuint64_t)D, (cuuint64_t)L, (cuuint64_t)LBG};
  cuuint64_t gstride[2] = {(cuuint64_t)D * 2, (cuuint64_t)L * D * 2};
  cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)BM, 1};
  return cached_map(key, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, gdim, gstride, box,
                    box_cols * 2 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
}
const CUtensorMap& get_map_hm(const void* ptr, int D, int L, int G, int64_t LB) {
  MapKey key{ptr, L, D, D, BM, D, 2, 2, G, LB};
  cuuint64_t gdim[4] = {(cuuint64_t)D, (cuuint64_t)L, (cuuint64_t)G, (cuuint64_t)LB};
  cuuint64_t gstride[3] = {(cuuint64_t)D * 2, (cuuint64_t)L * D * 2, (cuuint64_t)G * L * D * 2};
  cuuint32_t box[4] = {(cuuint32_t)D, (cuuint32_t)BM, 1, 1};
  return cached_map(key, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, gdim, gstride, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

template <int BN, int EW = kEpiWarpsWide>
void launch_cfg(const GemmArgs& a_in, int num_sms, cudaStream_t stream) {
  using Cfg = TileCfg<BN, EW>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN, EW>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) throw std::runtime_error(std::string("gemm_tc: cudaFuncSetAttribute: ") + cudaGetErrorString(e));
    attr_set = true;
  }
  GemmArgs a = a_in;
  const CUtensorMap* ma_ptr = Cfg::kCK > 1 ? &get_map_k3(a.A, a.M, a.K, a.lda, Cfg::kBM, Cfg::kCK) : &get_map(a.A, a.M, a.K, a.lda, Cfg::kBM);
  const CUtensorMap& mb = Cfg::kCK > 1 ? get_map_k3(a.W, a.N, a.K, a.ldw, BN, Cfg::kCK) : get_map(a.W, a.N, a.K, a.ldw, BN);
  // TMA-store epilogue when the output is expressible as a tensor map; otherwise the generic register/shared path
  const int esz = a.out_f32 ? 4 : 2;
  const int cw = 32;                                  // columns per TMA-store box (bf16: 64-byte rows, fp32: 128-byte rows)
  const CUtensorMap* mc = ma_ptr;
  a.tma_store = 0;
  static const bool no_tma_store = getenv("GSTVD_GEMM_NO_TMA_STORE") != nullptr;
  static const bool no_tma_hm = getenv("GSTVD_GEMM_NO_TMA_HM") != nullptr;
  if (!no_tma_store && (reinterpret_cast<uintptr_t>(a.C) & 15) == 0) {
    if (a.hm_D > 0) {
      if (!no_tma_hm && Cfg::kCK == 1 && !a.out_f32 && a.hm_D % cw == 0 && a.M == a.hm_B * a.hm_L) {
        const int64_t LB = (int64_t)(a.N / (a.hm_D * a.hm_G)) * a.hm_B;
        mc = &get_map_hm3(a.C, a.hm_D, a.hm_L, LB * a.hm_G, cw);
        a.hm_tpi = (a.hm_L + BM - 1) / BM;
        ma_ptr = &get_map_a3(a.A, a.K, a.hm_L, a.hm_B, a.lda);
        a.tma_store = 1;
      }
    } else if ((a.ldc * esz) % 16 == 0) {
      static const bool no_direct = getenv("GSTVD_GEMM_NO_DIRECT") != nullptr;     // A/B aid
      if (a.M <= 4 * BM && (int64_t)a.M * a.N * esz <= (4 << 20) && !no_direct) {
        a.tma_store = 2;                                 // skinny (decode) problems: latency matters, the tile is tiny
      } else {
        mc = &get_map_c(a.C, a.M, a.N, a.ldc, esz, cw);
        a.tma_store = 1;
      }
    }
  }
  const int tiles_m = a.hm_tpi > 0 ? a.hm_B * a.hm_tpi : (a.M + Cfg::kBM - 1) / Cfg::kBM;
  const int tiles = tiles_m * ((a.N + BN - 1) / BN);
  if (Cfg::kSkinny && a.tma_store != 2) throw std::runtime_error("gemm_tc: the skinny configuration only has the direct-store epilogue");
  const int slots = Cfg::kSkinny ? 2 * num_sms : num_sms;      // resident CTAs: the kernel is persistent over the remaining tiles
  const int grid = tiles < slots ? tiles : slots;
  launch_k(gemm_tc_kernel<BN, EW>, dim3(grid), dim3(Cfg::kThreads), (size_t)Cfg::kSmemBytes, stream, *ma_ptr, mb, *mc, a);
}

// ------------------------------------------------------------------------------------------------------------
// Y = LayerNorm(A * W^T + bias + residual)   for the decode step's three N = H projections (attention output, cross-attention
// output, FFN2; models/visual_dialog_decoder.py:300-311 -> HF BertSelfOutput / BertOutput: dense -> dropout -> LN(x + input)).
//
// The separate add_layernorm launch after each of those GEMMs costs a whole kernel boundary (~3 us of a latency-bound step, 36
// per decode step) for 0.5 MB of traffic.  Here one thread-block CLUSTER owns a 64-row block: CTA j of the cluster computes the
// 64 x BN accumulator tile of columns [j*BN, (j+1)*BN) (CL * BN = N), keeps its rows in registers, pushes one (mean, M2) pair per
// row into every peer's shared memory (st.shared::cluster), and after ONE cluster barrier each CTA combines the CL pairs of its
// rows (Chan's parallel variance, fixed order -> deterministic and identical in every CTA) and normalises its own columns.
// An earlier attempt synchronised the column tiles through a global counter per row block and lost: the tiles
// of a row block were not co-scheduled.  A cluster is co-scheduled by construction.
//
// Pipeline per CTA (one tile, not persistent): warp 0 = TMA producer (A: 64 rows x 256 k, W: BN rows x 256 k per stage, 3-D
// boxes of four 128-byte-swizzled k-chunks), warp 1 = tcgen05.mma issuer (M = 64, N = BN), warps 2-5 = epilogue (one per TMEM
// lane quadrant; the M = 64 accumulator keeps rows in lanes 0-15 of each quadrant).
template <int CL, int BN, int CK = 4> struct LnTileCfg {
  static constexpr int kBM = 64;
  static constexpr int kCK = CK;                               // k-chunks (64 elements) per ring stage
  static constexpr int kStages = BN <= 48 ? 3 : 2;
  // CK = 2: 95 KB per CTA instead of 182 KB, so that two CTAs (this kernel's, or another stream's GEMM) share an SM - the 182 KB
  // configuration lost 6 % dialogs/s with three streams in flight although it won single-stream
  static constexpr int kCtasPerSm = CK <= 2 && BN <= 48 ? 2 : 1;
  static constexpr int kAChunk = kBM * BK * 2;
  static constexpr int kBChunk = BN * BK * 2;
  static constexpr int kABytes = kAChunk * kCK;
  static constexpr int kBBytes = kBChunk * kCK;
  static constexpr int kTmemCols = BN <= 32 ? 32 : (BN <= 64 ? 64 : (BN <= 128 ? 128 : 256));
  static constexpr int kStatsBytes = CL * kBM * 8;             // (mean, M2) per source CTA per row
  static constexpr int kParamBytes = 3 * BN * 4;               // bias | gamma | beta of this CTA's columns
  static constexpr int kBarBytes = 128;
  static constexpr int kThreads = 6 * 32;
  static constexpr int kSmemBytes = kStages * (kABytes + kBBytes) + kStatsBytes + kParamBytes + kBarBytes + 1024;
  static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "BN: multiple of 16 (tcgen05.ld x16 blocks, UMMA N % 8)");
  static_assert(kBChunk % 1024 == 0, "each k-chunk of the W tile must start on a swizzle-atom boundary");
  static_assert(kSmemBytes <= 232448, "shared memory budget");
  static_assert(kCtasPerSm == 1 || 2 * (kSmemBytes + 1024) <= 233472, "two CTAs per SM must fit");
};

struct GemmLnArgs {
  const float* bias;        // [N] or null
  const bf16* res;          // [M, N] residual rows (stride ldr) or null
  int64_t ldr;
  const float* gamma;       // [N]
  const float* beta;        // [N]
  bf16* Y;                  // [M, N] (stride ldy)
  int64_t ldy;
  int M, N, K;
  float eps;
  int round_bf16;           // 1: round (A W^T + bias) to bf16 before adding the residual, like the unfused GEMM -> add_layernorm pair
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_arrive_release() { asm volatile("barrier.cluster.arrive.release;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait_acquire() { asm volatile("barrier.cluster.wait.acquire;" ::: "memory"); }
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_f32x2(uint32_t addr, float a, float b) {
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(a), "f"(b) : "memory");
}
// 32 lanes x 16 consecutive 32-bit columns, no wait (several loads are issued back to back, then one tcgen05.wait::ld)
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

template <int CL, int BN, int CK>
__global__ void __launch_bounds__(6 * 32, LnTileCfg<CL, BN, CK>::kCtasPerSm)
gemm_ln_cluster_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, const GemmLnArgs p) {
  using Cfg = LnTileCfg<CL, BN, CK>;
  constexpr int kStages = Cfg::kStages;
  constexpr int kCK = Cfg::kCK;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;       // the same offset in every CTA of the cluster (same kernel, same layout)
  const uint32_t a_base = base;
  const uint32_t b_base = base + kStages * Cfg::kABytes;
  const uint32_t stats_base = b_base + kStages * Cfg::kBBytes;
  const uint32_t param_base = stats_base + Cfg::kStatsBytes;
  const uint32_t bar_base = param_base + Cfg::kParamBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  const uint32_t tfull_bar = bar_base + 8u * (2 * kStages);
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 1);
  uint8_t* smem_gen = smem_raw + (base - raw);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - base));
  const float2* stats = reinterpret_cast<const float2*>(smem_gen + (stats_base - base));
  float* params = reinterpret_cast<float*>(smem_gen + (param_base - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();             // column tile of this CTA (cluster spans gridDim.x == CL)
  const int m_blk = blockIdx.y;
  const int n0 = (int)rank * BN;
  const int num_kb = p.K / (BK * kCK);                 // host guarantees K % 256 == 0 (a multiple of every BK * kCK in use)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(tfull_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  if (warp >= 2) {
    // bias | gamma | beta of this CTA's columns: weights, independent of the previous kernel -> fetched before the PDL wait
    for (int i = threadIdx.x - 64; i < 3 * BN; i += 128) {
      const int which = i / BN, c = i - which * BN;
      const float* src = which == 0 ? p.bias : (which == 1 ? p.gamma : p.beta);
      params[i] = src ? __ldg(src + n0 + c) : 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  // every CTA of the cluster is running before anybody writes into a peer's shared memory
  cluster_arrive_release();
  cluster_wait_acquire();

  float z[BN];                                         // epilogue threads: this row's BN pre-LN values (live across the barrier)
  const int quad = warp & 3;
  const bool epi_active = warp >= 2 && lane < 16;      // M = 64: rows in lanes 0-15 of each TMEM lane quadrant
  const int rt = quad * 16 + (lane & 15);              // row within the tile
  const int row = m_blk * Cfg::kBM + rt;

  if (warp == 0) {
    // ---------------- TMA producer ----------------
    if (lane == 0) {
      pdl_launch_dependents();
      const int pre = num_kb < kStages ? num_kb : kStages;
      for (int st = 0; st < pre; ++st) {               // weights first: they never depend on the previous kernel
        mbar_arrive_expect_tx(full_bar(st), Cfg::kABytes + Cfg::kBBytes);
        tma_load_3d(b_base + st * Cfg::kBBytes, &tm_b, 0, n0, st * kCK, full_bar(st));
      }
      pdl_wait();
      for (int st = 0; st < pre; ++st) tma_load_3d(a_base + st * Cfg::kABytes, &tm_a, 0, m_blk * Cfg::kBM, st * kCK, full_bar(st));
      int stage = pre == kStages ? 0 : pre;
      uint32_t phase = pre == kStages ? 1u : 0u;
      for (int kb = pre; kb < num_kb; ++kb) {
        mbar_wait(empty_bar(stage), phase ^ 1u);
        mbar_arrive_expect_tx(full_bar(stage), Cfg::kABytes + Cfg::kBBytes);
        tma_load_3d(a_base + stage * Cfg::kABytes, &tm_a, 0, m_blk * Cfg::kBM, kb * kCK, full_bar(stage));
        tma_load_3d(b_base + stage * Cfg::kBBytes, &tm_b, 0, n0, kb * kCK, full_bar(stage));
        if (++stage == kStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(Cfg::kBM, BN);
      int stage = 0; uint32_t phase = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
#pragma unroll
        for (int ck = 0; ck < kCK; ++ck) {
          const uint64_t a_desc = make_smem_desc(a_base + stage * Cfg::kABytes + ck * Cfg::kAChunk);
          const uint64_t b_desc = make_smem_desc(b_base + stage * Cfg::kBBytes + ck * Cfg::kBChunk);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) umma_f16(tmem_base, a_desc + 2u * k, b_desc + 2u * k, idesc, (kb | ck | k) != 0 ? 1u : 0u);
        }
        umma_commit(empty_bar(stage));
        if (kb == num_kb - 1) umma_commit(tfull_bar);
        if (++stage == kStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else {
    // ---------------- epilogue, part 1: z = acc + bias + residual, local statistics, push to the cluster ----------------
    uint4 rres[BN / 8];
    pdl_wait();                                        // the residual rows were written by an earlier kernel of the chain
    if (epi_active && row < p.M && p.res != nullptr) {
      const uint4* rp = reinterpret_cast<const uint4*>(p.res + (int64_t)row * p.ldr + n0);
#pragma unroll
      for (int i = 0; i < BN / 8; ++i) rres[i] = __ldg(rp + i);
    } else {
#pragma unroll
      for (int i = 0; i < BN / 8; ++i) rres[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    mbar_wait(tfull_bar, 0u);
    __syncwarp();
    tc_fence_after();
    const uint32_t tq = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    uint32_t acc[BN];
#pragma unroll
    for (int i = 0; i < BN / 16; ++i) tmem_ld16_nowait(tq + i * 16, acc + i * 16);
    tmem_ld_wait();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < BN / 8; ++i) {
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&rres[i]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 r2 = __bfloat1622float2(h[j]);
        float v0 = __uint_as_float(acc[i * 8 + 2 * j]) + params[i * 8 + 2 * j];
        float v1 = __uint_as_float(acc[i * 8 + 2 * j + 1]) + params[i * 8 + 2 * j + 1];
        if (p.round_bf16) { v0 = __bfloat162float(__float2bfloat16_rn(v0)); v1 = __bfloat162float(__float2bfloat16_rn(v1)); }
        v0 += r2.x; v1 += r2.y;
        z[i * 8 + 2 * j] = v0; z[i * 8 + 2 * j + 1] = v1;
        s += v0 + v1;
      }
    }
    const float mean_l = s * (1.0f / BN);
    float m2 = 0.f;
#pragma unroll
    for (int i = 0; i < BN; ++i) { const float d = z[i] - mean_l; m2 = fmaf(d, d, m2); }
    if (epi_active) {
      const uint32_t slot = stats_base + (rank * Cfg::kBM + rt) * 8u;    // stats[source = this CTA][row]
#pragma unroll
      for (int dst = 0; dst < CL; ++dst) st_cluster_f32x2(map_to_cta(slot, (uint32_t)dst), mean_l, m2);
    }
    tc_fence_before();
  }
  // one barrier for the whole cluster: the pushed statistics are visible after it, and nobody touches a peer's shared
  // memory past this point (so any CTA may exit as soon as it has finished its own rows)
  cluster_arrive_release();
  cluster_wait_acquire();

  if (warp >= 2) {
    // ---------------- epilogue, part 2: combine, normalise, store ----------------
    if (epi_active && row < p.M) {
      float msum = 0.f;
#pragma unroll
      for (int j = 0; j < CL; ++j) msum += stats[j * Cfg::kBM + rt].x;
      const float mean = msum * (1.0f / CL);
      float m2 = 0.f;
#pragma unroll
      for (int j = 0; j < CL; ++j) {
        const float2 st = stats[j * Cfg::kBM + rt];
        const float d = st.x - mean;
        m2 += st.y + (float)BN * d * d;
      }
      const float var = m2 * (1.0f / (CL * BN));
      const float denom = sqrtf(var + p.eps);
      uint4* yp = reinterpret_cast<uint4*>(p.Y + (int64_t)row * p.ldy + n0);
#pragma unroll
      for (int i = 0; i < BN / 8; ++i) {
        uint4 o;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = i * 8 + 2 * j;
          const float y0 = params[BN + c] * ((z[c] - mean) / denom) + params[2 * BN + c];
          const float y1 = params[BN + c + 1] * ((z[c + 1] - mean) / denom) + params[2 * BN + c + 1];
          h[j] = __floats2bfloat162_rn(y0, y1);
        }
        yp[i] = o;
      }
    }
  } else if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

template <int CL, int BN, int CK>
void launch_ln_cfg(const GemmArgs& a, const GemmLnArgs& p, cudaStream_t stream) {
  using Cfg = LnTileCfg<CL, BN, CK>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_ln_cluster_kernel<CL, BN, CK>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e == cudaSuccess && CL > 8) e = cudaFuncSetAttribute(gemm_ln_cluster_kernel<CL, BN, CK>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) throw std::runtime_error(std::string("gemm_ln: cudaFuncSetAttribute: ") + cudaGetErrorString(e));
    attr_set = true;
  }
  const CUtensorMap& ma = get_map_k3(a.A, a.M, a.K, a.lda, Cfg::kBM, Cfg::kCK);
  const CUtensorMap& mb = get_map_k3(a.W, a.N, a.K, a.ldw, BN, Cfg::kCK);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(CL, (a.M + Cfg::kBM - 1) / Cfg::kBM, 1);
  cfg.blockDim = dim3(Cfg::kThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_flag() ? 2 : 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_ln_cluster_kernel<CL, BN, CK>, ma, mb, p);
  if (e != cudaSuccess) throw std::runtime_error(std::string("gemm_ln: launch failed: ") + cudaGetErrorString(e));
}

// ------------------------------------------------------------------------------------------------------------
// CTA-pair GEMM (tcgen05 cta_group::2) for the throughput problems (encoder / prefill / teacher-forced passes, M >= 1024).
//
// Why: with one CTA per 128 x 256 tile every SM pulls 48 KB of operands through L2 per 64-wide k-block, i.e. 106 GB/s per SM at
// the full MMA rate - 15.7 TB/s over 148 SMs, above what L2 delivers (~12 TB/s): the single-CTA kernel is L2-bound at 57 % of the
// measured bf16 peak.  A CTA pair (two SMs of one TPC, cluster of 2) computes a 256 x BN tile with ONE tcgen05.mma.cta_group::2
// per k-step issued by the leader: each CTA stages its own 128 rows of A and only HALF of the W tile (BN/2 rows) - the tensor
// cores read the other half from the peer's shared memory - so the per-SM operand traffic drops to 32 KB per k-block (-33 %)
// for the same math.
//
// Protocol (follows the published CUTLASS sm100 2-SM pipeline):
//   * both CTAs run a TMA producer (warp 0): cp.async.bulk.tensor ... .cta_group::2 with the mbarrier address' peer bit cleared,
//     so the bytes of BOTH CTAs complete on the LEADER's full barrier; only the leader's producer arms it (expect_tx = 2 x stage);
//   * the leader's warp 1 issues the MMAs; tcgen05.commit ... multicast::cluster (mask 0b11) releases the ring slot in both CTAs
//     and, after the last k-block, signals both CTAs' accumulator-full barriers;
//   * each CTA's 16 epilogue warps drain their own 128 accumulator rows (bias / GELU / convert / TMA store, shared with the
//     single-CTA kernel) and arrive on the leader's accumulator-empty barrier (remote mbarrier.arrive for the follower);
//   * TMEM is allocated / freed with the cta_group::2 forms by warp 1 of both CTAs.
// EXPERIMENTAL: selected only with env GSTVD_GEMM_2CTA=1 until it has been validated on the GPU.
template <int BN> struct Tile2Cfg {
  static constexpr int kBMc = BM;                               // rows per CTA; the pair tile has 2 * BM rows
  static constexpr int kBH = BN / 2;                            // W rows staged by each CTA
  static constexpr int kStages = BN == 256 ? 6 : 8;
  static constexpr int kABytes = kBMc * BK * 2;
  static constexpr int kBBytes = kBH * BK * 2;
  static constexpr int kTmemCols = 2 * BN;                      // two accumulator stages
  static constexpr int kBarBytes = 256;
  static constexpr int kStageWords = 32 * 33;
  static constexpr int kStagingBytes = 8 * kStageWords * 4;
  static constexpr int kSmemBytes = kStages * (kABytes + kBBytes) + kBarBytes + kStagingBytes + 1024;
  static constexpr int kThreads = (2 + kEpiWarpsWide) * 32;
  static_assert(BN == 128 || BN == 256, "pair tile width");
  static_assert(kSmemBytes <= 232448, "shared memory budget");
  static_assert(8 * (2 * kStages + 4) + 4 <= kBarBytes, "barrier block");
};

constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;                  // shared::cluster address of the same offset in the even CTA of a pair

__device__ __forceinline__ void tma2_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t leader_bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(leader_bar)
      : "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_alloc2(uint32_t smem_dst) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "n"(COLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  const uint32_t z = 0u;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc), "r"(z)
      : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint32_t bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cta(uint32_t bar, uint32_t cta_rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(bar), "r"(cta_rank) : "memory");
}

template <int BN>
__global__ void __launch_bounds__((2 + kEpiWarpsWide) * 32, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                const __grid_constant__ CUtensorMap tm_c, const GemmArgs p) {
  using Cfg = Tile2Cfg<BN>;
  constexpr int kStages = Cfg::kStages;
  constexpr int EW = kEpiWarpsWide;
  constexpr int kGroups = EW / 4;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t a_base = base;
  const uint32_t b_base = base + kStages * Cfg::kABytes;
  const uint32_t stage_base = b_base + kStages * Cfg::kBBytes;
  const uint32_t bar_base = stage_base + Cfg::kStagingBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * kStages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * kStages + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 4);
  uint8_t* smem_gen = smem_raw + (base - raw);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();              // 0 = leader (issues the MMAs), 1 = follower
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int tiles_n = (p.N + BN - 1) / BN;
  const int tiles_m = (p.M + 2 * BM - 1) / (2 * BM);
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = (p.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    tma_prefetch_desc(&tm_c);
    for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 2 * EW * 32); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc2<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  // the barriers of both CTAs are initialised (and both halves of the TMEM allocation exist) before either CTA signals the other
  cluster_arrive_release();
  cluster_wait_acquire();

  if (warp == 0) {
    // ---------------- TMA producer (both CTAs) ----------------
    if (lane == 0) {
      pdl_launch_dependents();
      int stage = 0; uint32_t phase = 0;
      bool first = true;
      constexpr uint32_t kPairStageBytes = 2u * (Cfg::kABytes + Cfg::kBBytes);
      auto load_b = [&](int st, int kb, int n_blk) {
        tma2_load_2d(b_base + st * Cfg::kBBytes, &tm_b, kb * BK, n_blk * BN + (int)rank * Cfg::kBH, full_bar(st) & kPeerBitMask);
      };
      auto load_a = [&](int st, int kb, int m_blk) {
        tma2_load_2d(a_base + st * Cfg::kABytes, &tm_a, kb * BK, m_blk * 2 * BM + (int)rank * BM, full_bar(st) & kPeerBitMask);
      };
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const int m_blk = tile / tiles_n, n_blk = tile - m_blk * tiles_n;
        int kb0 = 0;
        if (first) {
          first = false;
          const int pre = num_kb < kStages ? num_kb : kStages;
          for (int st = 0; st < pre; ++st) {
            if (leader) mbar_arrive_expect_tx(full_bar(st), kPairStageBytes);
            load_b(st, st, n_blk);
          }
          pdl_wait();
          for (int st = 0; st < pre; ++st) load_a(st, st, m_blk);
          kb0 = pre;
          if (pre == kStages) { stage = 0; phase = 1u; } else { stage = pre; }
        }
        for (int kb = kb0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);      // own ring slot released (multicast commit of the leader's MMAs)
          if (leader) mbar_arrive_expect_tx(full_bar(stage), kPairStageBytes);
          load_a(stage, kb, m_blk);
          load_b(stage, kb, n_blk);
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer (leader CTA only) ----------------
    if (lane == 0 && leader) {
      constexpr uint32_t idesc = make_idesc(2 * BM, BN);
      int stage = 0; uint32_t phase = 0; int iter = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs, ++iter) {
        const int as = iter & 1; const uint32_t aphase = (iter >> 1) & 1;
        mbar_wait(tempty_bar(as), aphase ^ 1u);         // both CTAs' epilogues have drained this accumulator stage
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);            // the bytes of BOTH CTAs have landed
          tc_fence_after();
          const uint64_t a_desc = make_smem_desc(a_base + stage * Cfg::kABytes);
          const uint64_t b_desc = make_smem_desc(b_base + stage * Cfg::kBBytes);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) umma2_f16(d_tmem, a_desc + 2u * k, b_desc + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
          umma2_commit_mc(empty_bar(stage), 3);         // frees the slot in both CTAs when these MMAs retire
          if (kb == num_kb - 1) umma2_commit_mc(tfull_bar(as), 3);
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else {
    // ---------------- epilogue (both CTAs: own 128 rows of the pair tile) ----------------
    const int e = warp - 2;
    const int quad = warp & 3;
    const int grp = e >> 2;
    int iter = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs, ++iter) {
      const int m_blk = tile / tiles_n, n_blk = tile - m_blk * tiles_n;
      const int as = iter & 1; const uint32_t aphase = (iter >> 1) & 1;
      if (p.bias != nullptr && lane == 0) {
        const int c0 = n_blk * BN + (e & 7) * (BN / 8);
        if (c0 < p.N) asm volatile("prefetch.global.L1 [%0];" ::"l"(p.bias + c0));
      }
      mbar_wait(tfull_bar(as), aphase);
      tc_fence_after();
      const int tile_row0 = m_blk * 2 * BM + (int)rank * BM;
      const int r = quad * 32 + lane;
      const bool issuer = (e & 3) == 0 && lane == 0;
      const uint32_t tq = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * BN;
      if (tile_row0 < p.M) {                            // block-uniform: a pair tile whose lower half is past M has nothing to store
        if (p.out_f32) {
          if (grp < 2) {
            const uint32_t stg = stage_base + grp * 16384;
            for (int j = grp; j < BN / 32; j += 2) {
              const int col0 = n_blk * BN + j * 32;
              if (col0 >= p.N) break;
              epilogue_tma_block<float, 32>(p, &tm_c, tq + j * 32, stg, tile_row0, col0, r, grp, issuer);
            }
          }
        } else {
          const uint32_t stg = stage_base + grp * 8192;
          for (int j = grp; j < BN / 32; j += kGroups) {
            const int col0 = n_blk * BN + j * 32;
            if (col0 >= p.N) break;
            epilogue_tma_block<bf16, 32>(p, &tm_c, tq + j * 32, stg, tile_row0, col0, r, grp, issuer);
          }
        }
      }
      tc_fence_before();
      mbar_arrive_cta(tempty_bar(as), 0u);              // on the leader's barrier (remote arrive for the follower)
    }
    if ((e & 3) == 0 && lane == 0) bulk_wait_read0();
  }
  tc_fence_before();
  __syncthreads();
  // neither CTA may leave while the other can still signal its barriers or read its shared memory
  cluster_arrive_release();
  cluster_wait_acquire();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2<Cfg::kTmemCols>(tmem_base);
  }
}

template <int BN>
void launch_cfg2(const GemmArgs& a_in, int num_sms, cudaStream_t stream) {
  using Cfg = Tile2Cfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc2_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) throw std::runtime_error(std::string("gemm_tc2: cudaFuncSetAttribute: ") + cudaGetErrorString(e));
    attr_set = true;
  }
  GemmArgs a = a_in;
  const int esz = a.out_f32 ? 4 : 2;
  const CUtensorMap& ma = get_map(a.A, a.M, a.K, a.lda, BM);
  const CUtensorMap& mb = get_map(a.W, a.N, a.K, a.ldw, Cfg::kBH);
  const CUtensorMap& mc = get_map_c(a.C, a.M, a.N, a.ldc, esz, 32);
  a.tma_store = 1; a.hm_tpi = 0;
  const int tiles = ((a.M + 2 * BM - 1) / (2 * BM)) * ((a.N + BN - 1) / BN);
  const int pairs = std::min(tiles, num_sms / 2);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * pairs); cfg.blockDim = dim3(Cfg::kThreads); cfg.dynamicSmemBytes = Cfg::kSmemBytes; cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_flag() ? 2 : 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_tc2_kernel<BN>, ma, mb, mc, a);
  if (e != cudaSuccess) throw std::runtime_error(std::string("gemm_tc2: launch failed: ") + cudaGetErrorString(e));
}

// CTA-pair configuration for this problem, or 0: needs the plain [M, N] TMA-store epilogue and enough rows to fill pair tiles.
int pick_pair_bn(const GemmArgs& a, int num_sms) {
  const char* env = getenv("GSTVD_GEMM_2CTA");
  if (env == nullptr || atoi(env) == 0) return 0;
  const int esz = a.out_f32 ? 4 : 2;
  if (a.hm_D != 0 || a.M < 8 * BM || a.N < 128 || a.K % 8 != 0 || (reinterpret_cast<uintptr_t>(a.C) & 15) != 0 || (a.ldc * esz) % 16 != 0 ||
      getenv("GSTVD_GEMM_NO_TMA_STORE") != nullptr)
    return 0;
  const int forced = atoi(env);
  if (forced == 128 || forced == 256) return forced;
  // the width that fills the 74 pairs best (a 128-column pair tile runs the tensor pipe at ~0.85 of the 256-column one)
  const int cand[2] = {256, 128};
  const double tile_eff[2] = {1.0, 0.85};
  const int pairs = num_sms / 2;
  const int64_t tm = (a.M + 2 * BM - 1) / (2 * BM);
  double best = -1.0; int bn = 256;
  for (int i = 0; i < 2; ++i) {
    const int64_t tiles = tm * ((a.N + cand[i] - 1) / cand[i]);
    const int64_t waves = (tiles + pairs - 1) / pairs;
    const double used = (double)a.N / ((double)((a.N + cand[i] - 1) / cand[i]) * cand[i]);
    const double eff = (double)tiles / (double)(waves * pairs) * tile_eff[i] * used;
    if (eff > best + 1e-9) { best = eff; bn = cand[i]; }
  }
  return bn;
}

}  // namespace

// 0 = off (default until measured on the GPU), else the cluster size to use (8 or 16); env GSTVD_FUSE_LN.
int gemm_ln_mode() {
  static const int mode = [] {
    const char* e = getenv("GSTVD_FUSE_LN");
    if (!e) return 0;
    const int v = atoi(e);
    return v == 8 ? 8 : (v != 0 ? 16 : 0);
  }();
  return mode;
}

bool gemm_ln_tc_supported(int M, int N, int K, const void* A, int64_t lda, const void* W, int64_t ldw, const void* res, int64_t ldr,
                          const void* Y, int64_t ldy) {
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  return M > 0 && N == 768 && K > 0 && K % 256 == 0 && al16(A) && al16(W) && al16(Y) && (res == nullptr || al16(res)) &&
         lda % 8 == 0 && ldw % 8 == 0 && ldy % 8 == 0 && (res == nullptr || ldr % 8 == 0);
}

// Y = LN(A W^T + bias + res) * gamma + beta, bf16 operands / output, fp32 accumulate and statistics.  a.C / a.ldc are unused.
int launch_gemm_ln_tc(const GemmArgs& a, const void* res, int64_t ldr, const float* gamma, const float* beta, float eps, void* Y,
                      int64_t ldy, int cluster, cudaStream_t stream) {
  if (!gemm_ln_tc_supported(a.M, a.N, a.K, a.A, a.lda, a.W, a.ldw, res, ldr, Y, ldy))
    throw std::runtime_error("gemm_ln: unsupported shape or alignment (N must be 768, K a multiple of 256)");
  if (a.act != 0 || a.hm_D != 0) throw std::runtime_error("gemm_ln: no activation / head-major output in the LayerNorm epilogue");
  gemm_tc_init();
  GemmLnArgs p;
  p.bias = a.bias; p.res = reinterpret_cast<const bf16*>(res); p.ldr = ldr; p.gamma = gamma; p.beta = beta;
  p.Y = reinterpret_cast<bf16*>(Y); p.ldy = ldy; p.M = a.M; p.N = a.N; p.K = a.K; p.eps = eps;
  static const bool exact_sum = getenv("GSTVD_FUSE_LN_NO_ROUND") != nullptr;   // keep the fp32 GEMM result instead of mirroring the bf16 round trip
  p.round_bf16 = exact_sum ? 0 : 1;
  const char* small_env = getenv("GSTVD_FUSE_LN_SMALL");      // read per launch: 95 KB configuration (two CTAs per SM), not yet run on a GPU
  const bool small = small_env != nullptr && atoi(small_env) != 0;
  if (cluster == 8) launch_ln_cfg<8, 96, 4>(a, p, stream);
  else if (small) launch_ln_cfg<16, 48, 2>(a, p, stream);
  else launch_ln_cfg<16, 48, 4>(a, p, stream);
  return 1;
}

void gemm_tc_init() {
  std::call_once(g_once, [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || fn == nullptr)
      throw std::runtime_error("gemm_tc: cannot resolve cuTensorMapEncodeTiled");
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  });
}

// [groups][L][D] bf16 rows (the decoder's cross K/V cache) as a 3-D map: box = D x box_rows rows of one group, 128-byte swizzle.
const void* tma_map_rows3(const void* ptr, int D, int L, int64_t groups, int box_rows) {
  gemm_tc_init();
  if (D * 2 != 128 || (reinterpret_cast<uintptr_t>(ptr) & 15) != 0) throw std::runtime_error("tma_map_rows3: rows must be 128 bytes, base 16-byte aligned");
  MapKey key{ptr, L, D, D, box_rows, D, 2, 6, groups, 0};
  cuuint64_t gdim[3] = {(cuuint64_t)D, (cuuint64_t)L, (cuuint64_t)groups};
  cuuint64_t gstride[2] = {(cuuint64_t)D * 2, (cuuint64_t)L * D * 2};
  cuuint32_t box[3] = {(cuuint32_t)D, (cuuint32_t)box_rows, 1};
  return &cached_map(key, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, gdim, gstride, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

int launch_gemm_tc(const GemmArgs& a_in, int num_sms, cudaStream_t stream, bool shared_sm) {
  if (a_in.M <= 0 || a_in.N <= 0) return 0;
  static const int dbg_env = [] { const char* e = getenv("GSTVD_GEMM_DBG"); return e ? atoi(e) : 0; }();
  static const int bn_env = [] { const char* e = getenv("GSTVD_GEMM_BN"); return e ? atoi(e) : 0; }();
  GemmArgs a = a_in;
  a.dbg = dbg_env;
  if (a.K % 8 != 0) throw std::runtime_error("gemm_tc: K must be a multiple of 8");
  if (a.hm_D > 0 && (a.hm_D % 32 != 0)) throw std::runtime_error("gemm_tc: head-major scatter needs head_dim % 32 == 0");
  gemm_tc_init();
  if (const int bn2 = pick_pair_bn(a, num_sms)) {          // EXPERIMENTAL CTA-pair kernel (env GSTVD_GEMM_2CTA)
    if (bn2 == 256) launch_cfg2<256>(a, num_sms, stream); else launch_cfg2<128>(a, num_sms, stream);
    return 1;
  }
  const int tiles_m = (a.M + BM - 1) / BM;
  int bn = 32;
  const int cand[4] = {256, 128, 64, 32};
  if (tiles_m <= 4) {
    // skinny (decode) problems are latency-bound: prefer ONE wave - the narrowest tile whose tile count still fits the SMs
    bn = 256;
    for (int i = 3; i >= 0; --i) {
      if (tiles_m * ((a.N + cand[i] - 1) / cand[i]) <= num_sms) { bn = cand[i]; break; }
    }
  } else {
    // Throughput problems: the tile width that maximises (wave efficiency) x (per-tile efficiency of that width).  Wide tiles
    // run the tensor pipe best, but a persistent grid of 148 CTAs wastes the last wave: e.g. M = 2048, N = 3072 is 192 tiles
    // of 256 columns (two waves, 65 % full) or 384 tiles of 128 columns (three waves, 86 % full).
    const double tile_eff[4] = {1.0, 0.88, 0.6, 0.35};
    double best = -1.0;
    for (int i = 0; i < 4; ++i) {
      const int64_t tiles = (int64_t)tiles_m * ((a.N + cand[i] - 1) / cand[i]);
      const int64_t waves = (tiles + num_sms - 1) / num_sms;
      const double used = (double)a.N / ((double)((a.N + cand[i] - 1) / cand[i]) * cand[i]);     // columns of the last tile that exist
      const double eff = (double)tiles / (double)(waves * num_sms) * tile_eff[i] * used;
      if (eff > best + 1e-9) { best = eff; bn = cand[i]; }
    }
  }
  if (bn_env) bn = bn_env;
  if (bn <= 64 && a.K % BK != 0) bn = 128;               // the chunked (3-D box) operand view needs whole 64-element chunks
  // Skinny (decode) problems whose output qualifies for direct row stores run the M = 64 configuration (two CTAs per SM):
  // 32-column tiles when they all fit the 2 x SMs resident slots, else 64-column tiles.
  static const int skinny_env = [] { const char* e = getenv("GSTVD_GEMM_SKINNY"); return e ? atoi(e) : -1; }();   // A/B aid: force 0 / 1
  (void)shared_sm;                                         // kept in the signature: the M = 64 configuration is shared-SM by construction
  const bool skinny = skinny_env >= 0 ? skinny_env != 0 : true;
  {
    const int esz = a.out_f32 ? 4 : 2;
    const bool direct_ok = a.hm_D == 0 && (reinterpret_cast<uintptr_t>(a.C) & 15) == 0 && (a.ldc * esz) % 16 == 0 && a.M <= 4 * BM &&
                           (int64_t)a.M * a.N * esz <= (4 << 20) && getenv("GSTVD_GEMM_NO_DIRECT") == nullptr &&
                           getenv("GSTVD_GEMM_NO_TMA_STORE") == nullptr;
    if (skinny && direct_ok && a.K % BK == 0 && !bn_env) {
      const int tm64 = (a.M + 63) / 64;
      static const int multi = [] { const char* e = getenv("GSTVD_GEMM_SKINNY_WAVES"); return e ? atoi(e) : 1; }();   // A/B aid
      if (tm64 * ((a.N + 31) / 32) <= 2 * num_sms * multi) { launch_cfg<32, kEpiWarpsSkinny>(a, num_sms, stream); return 1; }
      // wider outputs (N = 2304 / 3072 at M = 320) measured faster in the 128-row configuration with 16 epilogue warps
      // (5.5 / 5.7 us vs 5.9 / 6.2 us): 64-column tiles make each of the four epilogue warps handle two column blocks
      static const bool wide64 = getenv("GSTVD_GEMM_SKINNY64") != nullptr;
      if (wide64 && tm64 * ((a.N + 63) / 64) <= 2 * num_sms) { launch_cfg<64, kEpiWarpsSkinny>(a, num_sms, stream); return 1; }
    }
  }
  switch (bn) {
    case 256: launch_cfg<256>(a, num_sms, stream); break;
    case 128: launch_cfg<128>(a, num_sms, stream); break;
    case 64: launch_cfg<64>(a, num_sms, stream); break;
    default: launch_cfg<32>(a, num_sms, stream); break;
  }
  return 1;
}

}  // namespace gstvd
