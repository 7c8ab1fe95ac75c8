// Probe: is the ~85 GB/s per-SM operand ingest limit of the decode GEMMs (profiles/r2_gemm_phases.txt) a limit of the SM's receive
// port, or of the L2 request path - i.e. does TMA MULTICAST (one L2 read delivered to every CTA of a cluster) let an SM receive
// faster than unicast?  Every CTA (one per SM, 200 KB of shared memory) receives `total` bytes as `chunk`-byte bulk copies into a ring:
//   unicast  : every CTA issues all of its own copies (all CTAs of a cluster read the SAME source chunk - the shared A operand)
//   multicast: chunk i is issued once, by CTA (i mod C) of the cluster, with the cluster-wide destination mask
// One cluster barrier per ring round keeps the rounds apart.  Prints GB/s received per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o multicast_ingest tools/probes/multicast_ingest.cu && ./multicast_ingest
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ unsigned long long gns() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t phase) {
  uint32_t ok;
  asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(bar), "r"(phase) : "memory");
  return ok != 0;
}

constexpr int kStages = 8;

__global__ void __launch_bounds__(128) ingest(const uint8_t* __restrict__ src, size_t src_bytes, int chunk, int rounds, int csize, int mode,
                                              unsigned long long* out_ns, int* err) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[kStages];
  const uint32_t rank = csize > 1 ? cluster_ctarank() : 0;
  const int cluster_id = blockIdx.x / csize;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[s])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (csize > 1) cluster_sync_all();
  unsigned long long t0 = 0;
  if (threadIdx.x == 0) t0 = gns();
  // every cluster reads its own region of the source (L2 resident after the warm-up launch); all CTAs of a cluster read the same bytes
  const size_t region = (size_t)kStages * chunk;
  const uint8_t* base = src + ((size_t)cluster_id * region) % (src_bytes / 4);
  for (int r = 0; r < rounds; ++r) {
    if (threadIdx.x == 0) {
      for (int s = 0; s < kStages; ++s)
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[s])), "r"(chunk) : "memory");
    }
    if (csize > 1) cluster_sync_all(); else __syncthreads();       // every receiver is armed before anything is sent
    if (threadIdx.x == 0) {
      const uint8_t* rsrc = base + ((size_t)r * region) % (src_bytes / 4);
      for (int s = 0; s < kStages; ++s) {
        const uint32_t dst = smem_u32(smem + (size_t)s * chunk), bar = smem_u32(&bars[s]);
        if (mode == 0 || csize == 1) {
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(dst), "l"(rsrc + (size_t)s * chunk), "r"(chunk), "r"(bar) : "memory");
        } else if ((uint32_t)(s % csize) == rank) {
          const uint16_t mask = (uint16_t)((1u << csize) - 1u);
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                       ::"r"(dst), "l"(rsrc + (size_t)s * chunk), "r"(chunk), "r"(bar), "h"(mask) : "memory");
        }
      }
      const unsigned long long tw = gns();
      for (int s = 0; s < kStages; ++s) {
        while (!mbar_try(smem_u32(&bars[s]), (uint32_t)(r & 1))) {
          if (gns() - tw > 50000000ull) { atomicExch(err, 1); break; }
        }
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) out_ns[blockIdx.x] = gns() - t0;
  if (csize > 1) cluster_sync_all();
}

int main() {
  const size_t src_bytes = 64u << 20;
  uint8_t* src; CK(cudaMalloc(&src, src_bytes)); CK(cudaMemset(src, 1, src_bytes));
  unsigned long long* out; CK(cudaMalloc(&out, 256 * 8));
  int* err; CK(cudaMalloc(&err, 4)); CK(cudaMemset(err, 0, 4));
  const int chunk = 24 * 1024, rounds = 12;          // 8 x 24 KB = 192 KB ring, 2.3 MB per CTA
  const size_t smem = (size_t)kStages * chunk;
  CK(cudaFuncSetAttribute(ingest, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CK(cudaFuncSetAttribute(ingest, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  for (int csize : {1, 2, 4, 8}) {
    for (int mode = 0; mode < 2; ++mode) {
      if (csize == 1 && mode == 1) continue;
      for (int ctas : {24, 144}) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = csize; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        double best = 1e30, mean_best = 0;
        for (int rep = 0; rep < 4; ++rep) {
          CK(cudaLaunchKernelEx(&cfg, ingest, (const uint8_t*)src, src_bytes, chunk, rounds, csize, mode, out, err));
          CK(cudaDeviceSynchronize());
          std::vector<unsigned long long> h(ctas);
          CK(cudaMemcpy(h.data(), out, ctas * 8, cudaMemcpyDeviceToHost));
          double mx = 0, mean = 0;
          for (auto v : h) { mx = v > mx ? v : mx; mean += v; }
          mean /= ctas;
          if (mx < best) { best = mx; mean_best = mean; }
        }
        const double bytes = (double)kStages * chunk * rounds;
        int herr; CK(cudaMemcpy(&herr, err, 4, cudaMemcpyDeviceToHost));
        printf("cluster %d %-9s ctas %3d: %.0f KB per CTA in %.2f us (slowest CTA; mean %.2f) = %.1f GB/s per SM received  (timeout %d)\n", csize,
               mode ? "multicast" : "unicast", ctas, bytes / 1024, best / 1000, mean_best / 1000, bytes / best, herr);
      }
    }
  }
  return 0;
}
