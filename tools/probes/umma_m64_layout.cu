// Probe: where do the rows of a tcgen05.mma (cta_group::1, kind::f16) accumulator with M = 64 live in TMEM?
// A[r][0] = r + 1 (other k zero), B[n][0] = 1  =>  D[r][n] = r + 1.  Every warp then dumps its 32-lane quadrant
// (tcgen05.ld.32x32b.x32) and the host prints the value found in column 0 of each of the 128 lanes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o umma_m64_layout tools/probes/umma_m64_layout.cu && ./umma_m64_layout [M]
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__global__ void __launch_bounds__(128) probe(int M, float* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar;
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  uint8_t* gen = smem + (base - smem_u32(smem));
  const uint32_t a_addr = base, b_addr = base + 16384;
  // K-major SWIZZLE_128B tiles: element (r, k) at r*128 + (((k/8) ^ (r&7)) * 16) + (k%8)*2
  for (int i = threadIdx.x; i < 2 * 16384 / 2; i += blockDim.x) reinterpret_cast<__nv_bfloat16*>(gen)[i] = __float2bfloat16(0.f);
  __syncthreads();
  for (int r = threadIdx.x; r < 128; r += blockDim.x) {
    const int off = r * 128 + ((0 ^ (r & 7)) * 16);
    if (r < M) *reinterpret_cast<__nv_bfloat16*>(gen + off) = __float2bfloat16((float)(r + 1));
    if (r < 32) *reinterpret_cast<__nv_bfloat16*>(gen + 16384 + off) = __float2bfloat16(1.f);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  // poison the accumulator columns first so that untouched lanes are recognisable
  {
    const uint32_t taddr = tmem + ((uint32_t)((threadIdx.x >> 5) * 32) << 16);
    const uint32_t neg = __float_as_uint(-1.0f);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(neg) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint64_t ad = make_desc(a_addr), bd = make_desc(b_addr);
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(0u) : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  {
    uint32_t ok = 0, spins = 0;
    while (!ok) {
      if (++spins > (1u << 24)) __trap();                 // never hang the GPU
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t v;
  const uint32_t taddr = tmem + ((uint32_t)((threadIdx.x >> 5) * 32) << 16);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  out[threadIdx.x] = __uint_as_float(v);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem) : "memory");
}

int main(int argc, char** argv) {
  const int M = argc > 1 ? atoi(argv[1]) : 64;
  float* d; cudaMalloc(&d, 128 * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40960);
  probe<<<1, 128, 40960>>>(M, d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  float h[128]; cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
  printf("M = %d: accumulator row found in column 0 of each TMEM lane (-1 = lane not written)\n", M);
  for (int l = 0; l < 128; ++l) printf("%s%3d:%4.0f", (l % 16 == 0) ? "\n  lane " : "  ", l, h[l] - (h[l] > 0 ? 1.f : 0.f));
  printf("\n");
  return 0;
}
