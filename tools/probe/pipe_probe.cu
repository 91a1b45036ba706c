// Timeline probe for the host-buffer pipeline: which of upload / kernel / download really overlap on this box.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/pipe_probe tools/probe/pipe_probe.cu ; run: /tmp/pipe_probe [bytes] [kernel_us]
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

__global__ void spin(long long cycles, float *out) {
  const long long t0 = clock64();
  while (clock64() - t0 < cycles) {}
  if (out && threadIdx.x == 0 && blockIdx.x == 0) out[0] = 1.f;
}
#define CK(x) do { cudaError_t err_ = (x); if (err_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(err_)); exit(1); } } while (0)

int main(int argc, char **argv) {
  const size_t nb = argc > 1 ? atol(argv[1]) : 642048;
  const double kus = argc > 2 ? atof(argv[2]) : 10.0;
  int clk = 0; CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
  int ae = 0; CK(cudaDeviceGetAttribute(&ae, cudaDevAttrAsyncEngineCount, 0));
  printf("async engines %d, clock %d kHz, %zu bytes, kernel %.1f us\n", ae, clk, nb, kus);
  const long long cycles = (long long)(kus * 1e-6 * clk * 1e3);
  float *hx, *hg, *he; CK(cudaMallocHost(&hx, nb)); CK(cudaMallocHost(&hg, nb)); CK(cudaMallocHost(&he, 16));
  float *dx[2], *dg[2], *de[2];
  cudaStream_t s[3]; for (auto &q : s) CK(cudaStreamCreateWithFlags(&q, cudaStreamNonBlocking));
  for (int k = 0; k < 2; ++k) { CK(cudaMalloc(&dx[k], nb)); CK(cudaMalloc(&dg[k], nb)); CK(cudaMalloc(&de[k], 16)); }
  const int N = 12;
  for (int mode = 0; mode < 4; ++mode) {
    // mode 0: one stream, serial.  1: two alternating streams, kernels ordered by an event.  2: three streams (up | run | down).
    // mode 3: like 1 but download on a third stream
    std::vector<cudaEvent_t> ev(N * 4 + 1);
    for (auto &e : ev) CK(cudaEventCreate(&e));
    cudaEvent_t run[2], up[2], down[2];
    for (int k = 0; k < 2; ++k) { CK(cudaEventCreateWithFlags(&run[k], cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&up[k], cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&down[k], cudaEventDisableTiming)); }
    for (int rep = 0; rep < 2; ++rep) {       // rep 0 warms up
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(ev[N * 4], s[0]));
      for (int i = 0; i < N; ++i) {
        const int k = i & 1;
        if (mode == 0) {
          CK(cudaEventRecord(ev[4 * i], s[0]));
          CK(cudaMemcpyAsync(dx[k], hx, nb, cudaMemcpyHostToDevice, s[0]));
          CK(cudaEventRecord(ev[4 * i + 1], s[0]));
          spin<<<148, 512, 0, s[0]>>>(cycles, de[k]);
          CK(cudaEventRecord(ev[4 * i + 2], s[0]));
          CK(cudaMemcpyAsync(he, de[k], 12, cudaMemcpyDeviceToHost, s[0]));
          CK(cudaMemcpyAsync(hg, dg[k], nb, cudaMemcpyDeviceToHost, s[0]));
          CK(cudaEventRecord(ev[4 * i + 3], s[0]));
        } else if (mode == 1 || mode == 3) {
          cudaStream_t sk = s[k];
          CK(cudaEventRecord(ev[4 * i], sk));
          CK(cudaMemcpyAsync(dx[k], hx, nb, cudaMemcpyHostToDevice, sk));
          CK(cudaEventRecord(ev[4 * i + 1], sk));
          CK(cudaStreamWaitEvent(sk, run[k ^ 1], 0));
          if (mode == 3) CK(cudaStreamWaitEvent(sk, down[k], 0));
          spin<<<148, 512, 0, sk>>>(cycles, de[k]);
          CK(cudaEventRecord(run[k], sk));
          CK(cudaEventRecord(ev[4 * i + 2], sk));
          cudaStream_t sd = mode == 3 ? s[2] : sk;
          if (mode == 3) CK(cudaStreamWaitEvent(sd, run[k], 0));
          CK(cudaMemcpyAsync(he, de[k], 12, cudaMemcpyDeviceToHost, sd));
          CK(cudaMemcpyAsync(hg, dg[k], nb, cudaMemcpyDeviceToHost, sd));
          CK(cudaEventRecord(ev[4 * i + 3], sd));
          if (mode == 3) CK(cudaEventRecord(down[k], sd));
        } else {
          CK(cudaStreamWaitEvent(s[0], run[k], 0));
          CK(cudaEventRecord(ev[4 * i], s[0]));
          CK(cudaMemcpyAsync(dx[k], hx, nb, cudaMemcpyHostToDevice, s[0]));
          CK(cudaEventRecord(ev[4 * i + 1], s[0]));
          CK(cudaEventRecord(up[k], s[0]));
          CK(cudaStreamWaitEvent(s[1], up[k], 0));
          CK(cudaStreamWaitEvent(s[1], down[k], 0));
          spin<<<148, 512, 0, s[1]>>>(cycles, de[k]);
          CK(cudaEventRecord(run[k], s[1]));
          CK(cudaEventRecord(ev[4 * i + 2], s[1]));
          CK(cudaStreamWaitEvent(s[2], run[k], 0));
          CK(cudaMemcpyAsync(he, de[k], 12, cudaMemcpyDeviceToHost, s[2]));
          CK(cudaMemcpyAsync(hg, dg[k], nb, cudaMemcpyDeviceToHost, s[2]));
          CK(cudaEventRecord(ev[4 * i + 3], s[2]));
          CK(cudaEventRecord(down[k], s[2]));
        }
      }
      CK(cudaDeviceSynchronize());
    }
    printf("mode %d: call: upload [start,end]  kernel end  download end   (us from the first record)\n", mode);
    float last = 0;
    for (int i = 0; i < N; ++i) {
      float t[4];
      for (int j = 0; j < 4; ++j) CK(cudaEventElapsedTime(&t[j], ev[N * 4], ev[4 * i + j]));
      printf("  %2d: up [%7.1f, %7.1f]  k_end %7.1f  down_end %7.1f\n", i, t[0] * 1e3, t[1] * 1e3, t[2] * 1e3, t[3] * 1e3);
      last = t[3];
    }
    printf("  -> %.1f us per call\n", last * 1e3 / N);
  }
  return 0;
}
