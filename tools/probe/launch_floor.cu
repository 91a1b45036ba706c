// Launch-to-launch floor of a persistent grid on this GPU: graph replay of N dependent launches of a kernel that does
// (almost) nothing, as a function of dynamic shared memory, block size and programmatic dependent launch.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/probe/launch_floor tools/probe/launch_floor.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t err_ = (x); if (err_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(err_)); exit(1); } } while (0)

template <bool PDL, bool WORK>
__global__ void k(float *out, const float *in, int spin) {
  extern __shared__ float sm[];
  if (PDL) { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); asm volatile("griddepcontrol.wait;" ::: "memory"); }
  if (WORK) {       // a dependent chain of `spin` shared-memory round trips per thread
    sm[threadIdx.x] = in[threadIdx.x & 31];
    __syncthreads();
    float a = 0.f;
    int j = threadIdx.x;
    for (int i = 0; i < spin; ++i) { a += sm[j]; j = (j * 5 + 1) & (blockDim.x - 1); }
    if (a == 12345.f) out[0] = a;
  }
  if (threadIdx.x == 0 && blockIdx.x == 0) out[1] = 1.f;
}

template <bool PDL, bool WORK>
float run(int grid, int block, size_t smem, int spin, float *d, cudaStream_t st) {
  CK(cudaFuncSetAttribute(k<PDL, WORK>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = PDL ? 1 : 0;
  cfg.attrs = at; cfg.numAttrs = 1;
  const int N = 50;
  cudaGraph_t g; cudaGraphExec_t ge;
  CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
  for (int i = 0; i < N; ++i) CK(cudaLaunchKernelEx(&cfg, k<PDL, WORK>, d, (const float *)d + 64, spin));
  CK(cudaStreamEndCapture(st, &g));
  CK(cudaGraphInstantiate(&ge, g, 0));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 5; ++i) CK(cudaGraphLaunch(ge, st));
  CK(cudaStreamSynchronize(st));
  CK(cudaEventRecord(e0, st));
  for (int i = 0; i < 40; ++i) CK(cudaGraphLaunch(ge, st));
  CK(cudaEventRecord(e1, st));
  CK(cudaStreamSynchronize(st));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  cudaGraphExecDestroy(ge); cudaGraphDestroy(g);
  return ms * 1e3f / (40 * N);
}

int main() {
  float *d; CK(cudaMalloc(&d, 4096)); CK(cudaMemset(d, 0, 4096));
  cudaStream_t st; CK(cudaStreamCreate(&st));
  printf("us per launch (graph of 50 dependent launches)\n  grid block smem_KB :  empty/noPDL  empty/PDL | 3us-of-work/noPDL  work/PDL\n");
  const int grids[] = {148, 296}, blocks[] = {512, 256};
  const size_t smems[] = {0, 48 << 10, 100 << 10, 200 << 10};
  for (int gi = 0; gi < 2; ++gi)
    for (size_t s : smems) {
      const int grid = grids[gi], block = blocks[gi];
      if (grid == 296 && s > (100 << 10)) continue;
      const float a = run<false, false>(grid, block, s, 0, d, st), b = run<true, false>(grid, block, s, 0, d, st);
      const float c = run<false, true>(grid, block, s > 2048 ? s : 2048, 200, d, st), e = run<true, true>(grid, block, s > 2048 ? s : 2048, 200, d, st);
      printf("  %4d %5d %7zu : %11.2f %10.2f | %17.2f %9.2f\n", grid, block, s >> 10, a, b, c, e);
    }
  return 0;
}
