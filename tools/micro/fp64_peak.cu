// fp64 issue rate of one B200 SM: DFMA / DADD / DMUL lanes per clock per SM, as a function of resident warps and of
// the number of independent chains per thread.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a fp64_peak.cu -o fp64_peak
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS, int OP>
__global__ void k(double* out, double a, double b, int iters) {
  double v[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) v[c] = a + c + threadIdx.x;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) {
      if (OP == 0) v[c] = fma(v[c], a, b);
      if (OP == 1) v[c] = v[c] + b;
      if (OP == 2) v[c] = v[c] * a;
    }
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) s += v[c];
  if (s == 123.456) out[threadIdx.x] = s;
}

template <int CHAINS, int OP>
void run(int threads_per_sm, const char* name) {
  int dev = 0, sms = 0, khz = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  double* out;
  cudaMalloc(&out, 1 << 20);
  const int iters = 20000, block = 256, blocks = sms * threads_per_sm / block;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<CHAINS, OP><<<blocks, block>>>(out, 1.0000001, 1e-9, 1000);
  cudaEventRecord(e0);
  k<CHAINS, OP><<<blocks, block>>>(out, 1.0000001, 1e-9, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double lane_ops = (double)blocks * block * iters * CHAINS;
  printf("%s chains %d threads/SM %4d: %.3f ms, %.1f lane-ops / clk / SM at the nominal %d MHz, %.2f T lane-ops/s\n", name,
         CHAINS, threads_per_sm, ms, lane_ops / (ms * 1e-3) / sms / (khz * 1e3), khz / 1000, lane_ops / (ms * 1e-3) / 1e12);
  cudaFree(out);
}

int main() {
  run<8, 0>(2048, "DFMA");
  run<8, 0>(1024, "DFMA");
  run<8, 0>(512, "DFMA");
  run<4, 0>(512, "DFMA");
  run<2, 0>(512, "DFMA");
  run<1, 0>(512, "DFMA");
  run<8, 1>(1024, "DADD");
  run<8, 2>(1024, "DMUL");
  run<8, 0>(256, "DFMA");
  return 0;
}
