// Cluster split-K tcgen05 GEMM for the decode step:  C[M,N] = act(A[M,K] * W[N,K]^T + bias), M <= 512 rows (B * K beams).
// EXPERIMENTAL: selected only with env GSTVD_GEMM_SPLITK=1 - written after the GPU budget of round 1 was spent, not yet run.
//
// Why: the operand bytes a tiled GEMM pulls through L2 are M*N*K*2*(1/tn + 1/tm) whatever else it does, and
// the L2 -> SM fabric is what several streams in flight share.  The decode GEMMs use 64 x 32 tiles to get enough CTAs out of
// M = 320 rows (120 CTAs for N = 768): 17.7 MB of operand traffic per 768 x 768 GEMM for 1.7 MB of unique operands, 215 MB per
// decoder layer and step.  Splitting K instead of shrinking the tile keeps the CTA count AND the big tile: a cluster of SPLIT = 4
// CTAs owns one 128 x 128 output tile, CTA r multiplies the k-range [r K/4, (r+1) K/4) into its own TMEM accumulator, and the four
// partial tiles are reduced through distributed shared memory:
//     operand traffic  7.1 MB per 768 x 768 GEMM (-60 %),  28 MB for FFN2 (K = 3072; 71 MB today),  106 MB per layer (-50 %)
//     CTAs             3 x 6 x 4 = 72 for N = 768, 288 for N = 3072 (one wave at two CTAs per SM)
//     per-CTA bytes    98 KB at K = 768 (147 KB today), 393 KB at K = 3072 (590 KB today): the main loop gets shorter too
// Reduction (deterministic - fixed summation order, no atomics): after every CTA's MMAs have retired (cluster barrier A, which also
// frees the operand rings the receive buffers alias), each epilogue thread (= accumulator row) keeps the 32-column quarter its CTA
// will finish and pushes the other three quarters into the owners' shared memory (st.shared::cluster.v4, rows padded to 144 bytes:
// conflict-free); after cluster barrier B each CTA sums its quarter in rank order (own partial at its own rank), adds the bias,
// applies the activation and stores 128 rows x 32 columns with 16-byte row stores.
#include <cstdlib>
#include <stdexcept>
#include <string>

#include "gemm_tc.cuh"

namespace gstvd {

using namespace tc;

namespace {

constexpr int kSplit = 4;
constexpr int kSkBN = 128;                       // tile columns; each CTA of the cluster finishes kSkBN / kSplit = 32 of them
constexpr int kSkQ = kSkBN / kSplit;
constexpr int kSkStages = 3;
constexpr int kSkABytes = BM * BK * 2;           // one 128-row x 64-element SW128 tile
constexpr int kSkBBytes = kSkBN * BK * 2;
constexpr int kSkRing = kSkStages * (kSkABytes + kSkBBytes);
constexpr int kSkRowBytes = kSkQ * 4 + 16;       // received quarter rows, padded: 144-byte stride spreads a quarter-warp over all banks
constexpr int kSkRecvBytes = kSplit * BM * kSkRowBytes;          // indexed by source rank (the own slot stays unused)
constexpr int kSkBarBytes = 128;
constexpr int kSkSmem = kSkRing + kSkBarBytes + 1024;
constexpr int kSkThreads = 6 * 32;
static_assert(kSkRecvBytes <= kSkRing, "the receive buffers alias the operand ring");
static_assert(2 * (kSkSmem + 1024) <= 233472, "two CTAs per SM");
static_assert(kSkQ == 32, "one tcgen05.ld x32 block per quarter");

__device__ __forceinline__ void st_cluster_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

__global__ void __launch_bounds__(kSkThreads, 2)
gemm_splitk_cluster_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, const GemmArgs p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;              // the same offset in every CTA of the cluster
  const uint32_t a_base = base;
  const uint32_t b_base = base + kSkStages * kSkABytes;
  const uint32_t recv_base = base;                           // aliases the ring: written by peers only after cluster barrier A
  const uint32_t bar_base = base + kSkRing;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kSkStages + s); };
  const uint32_t tfull_bar = bar_base + 8u * (2 * kSkStages);
  const uint32_t tmem_slot = bar_base + 8u * (2 * kSkStages + 1);
  uint8_t* smem_gen = smem_raw + (base - raw);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();                   // k-split index (the cluster spans gridDim.z == kSplit)
  const int n_blk = blockIdx.x, m_blk = blockIdx.y;
  const int k_len = p.K / kSplit;                            // host guarantees K % (kSplit * 64) == 0
  const int k0 = (int)rank * k_len;
  const int num_kb = k_len / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    for (int s = 0; s < kSkStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(tfull_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<kSkBN>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  uint32_t acc[kSkBN];                                       // epilogue threads: this row's partial sums (live across both barriers)
  const int quad = warp & 3;
  const int rt = quad * 32 + lane;                           // accumulator row == TMEM lane
  const int row = m_blk * BM + rt;

  if (warp == 0) {
    // ---------------- TMA producer ----------------
    if (lane == 0) {
      pdl_launch_dependents();
      const int pre = num_kb < kSkStages ? num_kb : kSkStages;
      for (int st = 0; st < pre; ++st) {                     // weights first: they never depend on the previous kernel
        mbar_arrive_expect_tx(full_bar(st), kSkABytes + kSkBBytes);
        tma_load_2d(b_base + st * kSkBBytes, &tm_b, k0 + st * BK, n_blk * kSkBN, full_bar(st));
      }
      pdl_wait();
      for (int st = 0; st < pre; ++st) tma_load_2d(a_base + st * kSkABytes, &tm_a, k0 + st * BK, m_blk * BM, full_bar(st));
      int stage = pre == kSkStages ? 0 : pre;
      uint32_t phase = pre == kSkStages ? 1u : 0u;
      for (int kb = pre; kb < num_kb; ++kb) {
        mbar_wait(empty_bar(stage), phase ^ 1u);
        mbar_arrive_expect_tx(full_bar(stage), kSkABytes + kSkBBytes);
        tma_load_2d(a_base + stage * kSkABytes, &tm_a, k0 + kb * BK, m_blk * BM, full_bar(stage));
        tma_load_2d(b_base + stage * kSkBBytes, &tm_b, k0 + kb * BK, n_blk * kSkBN, full_bar(stage));
        if (++stage == kSkStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BM, kSkBN);
      int stage = 0; uint32_t phase = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint64_t a_desc = make_smem_desc(a_base + stage * kSkABytes);
        const uint64_t b_desc = make_smem_desc(b_base + stage * kSkBBytes);
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) umma_f16(tmem_base, a_desc + 2u * k, b_desc + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
        umma_commit(empty_bar(stage));
        if (kb == num_kb - 1) umma_commit(tfull_bar);
        if (++stage == kSkStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else {
    // ---------------- epilogue, part 1: the partial tile leaves TMEM (all MMAs of this CTA have retired) ----------------
    mbar_wait(tfull_bar, 0u);
    __syncwarp();
    tc_fence_after();
    const uint32_t tq = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
#pragma unroll
    for (int i = 0; i < kSkBN / 16; ++i) tmem_ld16_nowait(tq + i * 16, acc + i * 16);
    tmem_ld_wait();
    tc_fence_before();
  }
  // barrier A: every CTA of the cluster has drained its accumulator, hence finished reading its operand ring - the rings may
  // now be overwritten with partial sums (and every CTA is known to be running)
  cluster_arrive_release();
  cluster_wait_acquire();

  if (warp >= 2) {
    // ---------------- epilogue, part 2: the three foreign quarters go to their owners ----------------
    const uint32_t my_slot = recv_base + ((uint32_t)rank * BM + (uint32_t)rt) * kSkRowBytes;   // recv[source = this CTA][row]
#pragma unroll
    for (int q = 0; q < kSplit; ++q) {
      if (q == (int)rank) continue;
      const uint32_t dst = map_to_cta(my_slot, (uint32_t)q);
#pragma unroll
      for (int j = 0; j < kSkQ / 4; ++j)
        st_cluster_v4(dst + 16u * j, acc[q * kSkQ + 4 * j], acc[q * kSkQ + 4 * j + 1], acc[q * kSkQ + 4 * j + 2], acc[q * kSkQ + 4 * j + 3]);
    }
  }
  // barrier B: the pushed partials are visible; nobody touches a peer's shared memory past this point
  cluster_arrive_release();
  cluster_wait_acquire();

  if (warp >= 2) {
    // ---------------- epilogue, part 3: sum this CTA's quarter in rank order, bias, activation, store ----------------
    const int col0 = n_blk * kSkBN + (int)rank * kSkQ;
    if (row < p.M && col0 < p.N) {
      float v[kSkQ];
#pragma unroll
      for (int j = 0; j < kSkQ; ++j) v[j] = 0.f;
#pragma unroll
      for (int s = 0; s < kSplit; ++s) {
        if (s == (int)rank) {
          // own partial: acc[rank * 32 + j]; the index is a run-time value, so select it without dynamic register indexing
#pragma unroll
          for (int q = 0; q < kSplit; ++q)
            if (q == (int)rank) {
#pragma unroll
              for (int j = 0; j < kSkQ; ++j) v[j] += __uint_as_float(acc[q * kSkQ + j]);
            }
        } else {
          const float4* src = reinterpret_cast<const float4*>(smem_gen + (recv_base - base) + ((size_t)s * BM + rt) * kSkRowBytes);
#pragma unroll
          for (int j = 0; j < kSkQ / 4; ++j) {
            const float4 t = src[j];
            v[4 * j] += t.x; v[4 * j + 1] += t.y; v[4 * j + 2] += t.z; v[4 * j + 3] += t.w;
          }
        }
      }
      if (p.bias != nullptr) {
#pragma unroll
        for (int j = 0; j < kSkQ / 4; ++j) {
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0) + j);
          v[4 * j] += b4.x; v[4 * j + 1] += b4.y; v[4 * j + 2] += b4.z; v[4 * j + 3] += b4.w;
        }
      }
      if (p.act == 1) {
#pragma unroll
        for (int j = 0; j < kSkQ; ++j) v[j] = gelu_fast(v[j]);
      }
      if (p.out_f32) {
        float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.C) + (int64_t)row * p.ldc + col0);
#pragma unroll
        for (int j = 0; j < kSkQ / 4; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      } else {
        uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(p.C) + (int64_t)row * p.ldc + col0);
#pragma unroll
        for (int j = 0; j < kSkQ / 8; ++j) {
          uint4 o;
          __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
          for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[8 * j + 2 * i], v[8 * j + 2 * i + 1]);
          dst[j] = o;
        }
      }
    }
  } else if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<kSkBN>(tmem_base);
  }
}

// The problems this kernel takes: decode-step projections (few rows, whole 128-column tiles, whole k-ranges per split, plain
// row-major output with 16-byte aligned rows and bias).
bool splitk_eligible(const GemmArgs& a, int max_rows) {
  const int esz = a.out_f32 ? 4 : 2;
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  return a.hm_D == 0 && a.M >= 1 && a.M <= max_rows && a.N >= kSkBN && a.N % kSkBN == 0 && a.N <= 4096 && a.K % (kSplit * BK) == 0 &&
         al16(a.A) && al16(a.W) && al16(a.C) && (a.bias == nullptr || al16(a.bias)) && a.lda % 8 == 0 && a.ldw % 8 == 0 &&
         (a.ldc * esz) % 16 == 0;
}

}  // namespace

namespace tc {

int launch_gemm_splitk_if_selected(const GemmArgs& a, cudaStream_t stream) {
  const char* env = getenv("GSTVD_GEMM_SPLITK");             // read per launch
  // 1: the decode problems (M <= 512).  2: also the half-wave problems of the encoder's image stream (M <= 4096: 2 368 rows at batch
  // 64 are 76 single-CTA tiles on 148 SMs; split over K they are 600 CTAs)
  const int mode = env ? atoi(env) : 0;
  if (mode == 0 || !splitk_eligible(a, mode >= 2 ? 4096 : 4 * BM)) return 0;
  gemm_tc_init();
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_splitk_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSkSmem);
    if (e != cudaSuccess) throw std::runtime_error(std::string("gemm_splitk: cudaFuncSetAttribute: ") + cudaGetErrorString(e));
    attr_set = true;
  }
  const CUtensorMap& ma = get_map(a.A, a.M, a.K, a.lda, BM);
  const CUtensorMap& mb = get_map(a.W, a.N, a.K, a.ldw, kSkBN);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(a.N / kSkBN, (a.M + BM - 1) / BM, kSplit);
  cfg.blockDim = dim3(kSkThreads);
  cfg.dynamicSmemBytes = kSkSmem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = kSplit;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_flag() ? 2 : 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_splitk_cluster_kernel, ma, mb, a);
  if (e != cudaSuccess) throw std::runtime_error(std::string("gemm_splitk: launch failed: ") + cudaGetErrorString(e));
  return 1;
}

}  // namespace tc

}  // namespace gstvd
