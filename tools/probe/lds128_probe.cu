// How many shared-memory wavefronts does one LDS.128 take as a function of the 32 lanes' addresses?
// One warp (and, second column, 16 warps) issues a long stream of independent LDS.128 with a fixed per-lane address
// pattern; cycles per instruction ~ wavefronts.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/probe/lds128_probe tools/probe/lds128_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t err_ = (x); if (err_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(err_)); exit(1); } } while (0)

__global__ void probe(const int *pattern, int iters, long long *cycles, float *sink) {
  extern __shared__ float4 sm[];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = make_float4(float(i), 1.f, 2.f, 3.f);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int a = pattern[lane];
  float4 acc0 = make_float4(0, 0, 0, 0), acc1 = acc0, acc2 = acc0, acc3 = acc0;
  __syncthreads();
  const long long t0 = clock64();
  const unsigned base = unsigned(__cvta_generic_to_shared(sm));
  for (int i = 0; i < iters; ++i) {
    // 8 independent loads per iteration; addresses move by multiples of 64 float4 (1 KB: every lane keeps its bank)
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const unsigned addr = base + 16u * unsigned((a + k * 64 + (i & 3) * 512) & 2047);
      float4 v;
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
      if (k & 1) acc1.x += v.x + v.w; else acc0.x += v.y + v.z;
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc0.x + acc1.x + acc2.x + acc3.x == 12345.f) sink[0] = 1.f;
}

int main() {
  struct Pat { const char *name; int a[32]; };
  Pat pats[12];
  int np = 0;
  auto add = [&](const char *name, auto f) { pats[np].name = name; for (int l = 0; l < 32; ++l) pats[np].a[l] = f(l); ++np; };
  add("32 distinct, consecutive (conflict-free)", [](int l) { return l; });
  add("all 32 lanes one address", [](int l) { (void)l; return 5; });
  add("4 quarter-warps read the same 8 consecutive", [](int l) { return l & 7; });
  add("16 distinct: lanes l and l+16 share", [](int l) { return l & 15; });
  add("16 distinct: lanes 2k, 2k+1 share (same quarter)", [](int l) { return l >> 1; });
  add("8 distinct, each bank group once, shared by 4 lanes in a quarter", [](int l) { return (l >> 2); });
  add("32 distinct, all in ONE bank group (stride 8 float4)", [](int l) { return 8 * l; });
  add("8 distinct in one bank group, every quarter-warp reads all 8", [](int l) { return 8 * (l & 7); });
  add("32 distinct, 2-way conflict inside each quarter-warp", [](int l) { return (l & ~7) + (l & 3) + 8 * 32 * ((l >> 2) & 1); });
  add("24 distinct: quarters 0 and 1 identical, 2 and 3 distinct", [](int l) { return l < 16 ? (l & 7) : l; });
  add("random-like 32 distinct", [](int l) { return (l * 37 + 11) & 255; });
  add("4 distinct addresses (quarter-warp q reads address q)", [](int l) { return l >> 3; });
  int *dp; long long *dc; float *ds;
  CK(cudaMalloc(&dp, 128)); CK(cudaMalloc(&dc, 8 * 256)); CK(cudaMalloc(&ds, 16));
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 2048 * 16));
  const int iters = 4096;
  printf("cycles per LDS.128:  1 warp | 16 warps (per warp-instruction, SM-wide throughput)\n");
  for (int p = 0; p < np; ++p) {
    CK(cudaMemcpy(dp, pats[p].a, 128, cudaMemcpyHostToDevice));
    long long c1 = 0, c16 = 0;
    for (int rep = 0; rep < 2; ++rep) {
      probe<<<1, 32, 2048 * 16>>>(dp, iters, dc, ds); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&c1, dc, 8, cudaMemcpyDeviceToHost));
      probe<<<1, 512, 2048 * 16>>>(dp, iters, dc, ds); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&c16, dc, 8, cudaMemcpyDeviceToHost));
    }
    printf("  %-68s %6.2f | %6.2f\n", pats[p].name, double(c1) / (iters * 8.0), double(c16) / (iters * 8.0 * 16));
  }
  return 0;
}
