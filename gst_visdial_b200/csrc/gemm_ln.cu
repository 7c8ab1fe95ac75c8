// Fused GEMM + residual + LayerNorm for the decode step (tcgen05 / TMEM / TMA, one thread-block cluster per 64-row block).
#include <cstdlib>
#include <stdexcept>
#include <string>

#include "gemm_tc.cuh"

namespace gstvd {

using namespace tc;

namespace {

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

}  // namespace gstvd
