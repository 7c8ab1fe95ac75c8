// Probe: what does one stage boundary of a dependent kernel chain cost on B200 -
//   mode 0: programmatic dependent launch, consumer calls griddepcontrol.wait (what the decode step does today)
//   mode 1: programmatic dependent launch, NO griddepcontrol.wait: the producer CTAs publish a counter (st data, fence, red.release),
//           the consumer CTAs spin on it (ld.acquire) - "dataflow flags over PDL"
//   mode 2: plain stream order (no PDL attribute)
// Every stage: 120 CTAs x 192 threads, `smem_kb` of dynamic shared memory (co-residency of two stages per SM like the decode GEMMs),
// reads four values other CTAs of the previous stage wrote, spins `body_ns`, writes its own.  The chain of `n` stages is captured
// into a CUDA graph and replayed; per-stage time = replay time / n; boundary cost = per-stage time - body.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o flag_vs_pdl tools/probes/flag_vs_pdl.cu && ./flag_vs_pdl
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ unsigned long long gns() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) { unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void red_release(unsigned* p, unsigned v) { asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

__global__ void __launch_bounds__(192) stage(const float* __restrict__ in, float* __restrict__ out, const unsigned* flag_prev, unsigned* flag_me,
                                             int mode, int body_ns, unsigned expected, int* err) {
  extern __shared__ unsigned char smem[];
  if (threadIdx.x == 0) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (mode == 0) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
  } else if (mode == 1) {
    if (threadIdx.x == 0 && flag_prev != nullptr) {
      const unsigned long long t0 = gns();
      while (ld_acquire(flag_prev) < expected) {
        if (gns() - t0 > 20000000ull) { atomicExch(err, 1); break; }       // 20 ms: never hang the box
      }
    }
    __syncthreads();
  }
  const int n = gridDim.x * blockDim.x;
  const int me = blockIdx.x * blockDim.x + threadIdx.x;
  float v = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) v += __ldcg(in + (me + (j * 31 + 7) * blockDim.x + j) % n);
  if (body_ns > 0) {
    const unsigned long long t0 = gns();
    while (gns() - t0 < (unsigned long long)body_ns) {}
  }
  smem[threadIdx.x] = (unsigned char)v;
  out[me] = v * 0.25f + 1.0f;
  if (mode == 1) {
    __syncthreads();
    if (threadIdx.x == 0) { __threadfence(); red_release(flag_me, 1u); }
  }
}

int main(int argc, char** argv) {
  const int n_stages = 400, ctas = argc > 1 ? atoi(argv[1]) : 120, threads = 192;
  const int n = ctas * threads;
  float *a, *b; unsigned* flags; int* err;
  CK(cudaMalloc(&a, n * 4)); CK(cudaMalloc(&b, n * 4)); CK(cudaMalloc(&flags, (n_stages + 1) * 4)); CK(cudaMalloc(&err, 4));
  CK(cudaMemset(err, 0, 4));
  cudaStream_t s; CK(cudaStreamCreate(&s));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int smem_kb : {1, 100}) {
    CK(cudaFuncSetAttribute(stage, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    for (int body_ns : {0, 2000}) {
      double per_stage[3]; float checks[3];
      for (int mode = 0; mode < 3; ++mode) {
        cudaGraph_t g; cudaGraphExec_t ge;
        CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
        CK(cudaMemsetAsync(flags, 0, (n_stages + 1) * 4, s));
        CK(cudaMemsetAsync(a, 0, n * 4, s));
        for (int i = 0; i < n_stages; ++i) {
          cudaLaunchConfig_t cfg = {};
          cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = (size_t)smem_kb * 1024; cfg.stream = s;
          cudaLaunchAttribute at[1];
          at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
          cfg.attrs = at; cfg.numAttrs = (mode == 2 || i == 0) ? 0 : 1;
          const float* in = (i & 1) ? b : a; float* out = (i & 1) ? a : b;
          const unsigned* fp = i == 0 ? nullptr : flags + i - 1;
          CK(cudaLaunchKernelEx(&cfg, stage, in, out, fp, flags + i, mode, body_ns, (unsigned)ctas, err));
        }
        CK(cudaStreamEndCapture(s, &g));
        CK(cudaGraphInstantiate(&ge, g, 0));
        for (int w = 0; w < 3; ++w) CK(cudaGraphLaunch(ge, s));
        CK(cudaStreamSynchronize(s));
        const int reps = 10;
        CK(cudaEventRecord(e0, s));
        for (int r = 0; r < reps; ++r) CK(cudaGraphLaunch(ge, s));
        CK(cudaEventRecord(e1, s));
        CK(cudaStreamSynchronize(s));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        per_stage[mode] = ms * 1000.0 / reps / n_stages;
        std::vector<float> h(n);
        CK(cudaMemcpy(h.data(), (n_stages & 1) ? b : a, n * 4, cudaMemcpyDeviceToHost));
        checks[mode] = h[0] + h[n / 2] + h[n - 1];
        CK(cudaGraphExecDestroy(ge)); CK(cudaGraphDestroy(g));
      }
      int herr; CK(cudaMemcpy(&herr, err, 4, cudaMemcpyDeviceToHost));
      printf("ctas %d smem %3d KB body %4d ns: us per stage  pdl+wait %.3f  pdl+flags %.3f  stream order %.3f   (checks %.4f %.4f %.4f, flag timeout %d)\n",
             ctas, smem_kb, body_ns, per_stage[0], per_stage[1], per_stage[2], checks[0], checks[1], checks[2], herr);
    }
  }
  return 0;
}
