// Kernel-boundary cost inside a CUDA graph, with and without programmatic dependent launch (PDL).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pdl_probe pdl_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void k_chain(float* buf, int n, int spin, int early) {
  if (early) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float v = i < n ? buf[i] : 0.f;
  for (int k = 0; k < spin; ++k) v = v * 1.0001f + 0.5f;
  if (i < n) buf[i] = v;
  if (!early) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

int run(int grid, int block, int spin, int pdl, int early, int chain, float* buf, int n, float* ms_out) {
  cudaStream_t s;
  CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
  for (int i = 0; i < chain; ++i) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.stream = s;
    cudaLaunchAttribute at;
    at.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &at; cfg.numAttrs = pdl ? 1 : 0;
    CK(cudaLaunchKernelEx(&cfg, k_chain, buf, n, spin, early));
  }
  cudaGraph_t g; CK(cudaStreamEndCapture(s, &g));
  cudaGraphExec_t ex; CK(cudaGraphInstantiate(&ex, g, 0));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) CK(cudaGraphLaunch(ex, s));
  CK(cudaEventRecord(e0, s));
  for (int i = 0; i < 20; ++i) CK(cudaGraphLaunch(ex, s));
  CK(cudaEventRecord(e1, s));
  CK(cudaStreamSynchronize(s));
  CK(cudaEventElapsedTime(ms_out, e0, e1));
  *ms_out /= 20.f * chain;
  cudaGraphExecDestroy(ex); cudaGraphDestroy(g); cudaStreamDestroy(s);
  return 0;
}

int main() {
  const int n = 148 * 8 * 256;
  float* buf; CK(cudaMalloc(&buf, n * 4)); CK(cudaMemset(buf, 0, n * 4));
  const int chain = 64;
  for (int spin : {0, 2000, 20000})
    for (int grid : {1, 148, 148 * 8}) {
      float a, b, c;
      if (run(grid, 256, spin, 0, 0, chain, buf, n, &a)) return 1;
      if (run(grid, 256, spin, 1, 0, chain, buf, n, &b)) return 1;
      if (run(grid, 256, spin, 1, 1, chain, buf, n, &c)) return 1;
      printf("grid %5d spin %6d : plain %.2f us/kernel | PDL trigger-at-end %.2f | PDL trigger-at-start %.2f\n", grid, spin, a * 1e3, b * 1e3, c * 1e3);
    }
  return 0;
}
