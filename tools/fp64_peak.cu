// fp64_peak.cu — measures the fp64 FMA peak of the GPU (the driver's MEASURED_PEAKS.json has no
// fp64 figure).  The CRNN kernels are fp64-ALU bound, so this is the denominator of the
// "fp64 fraction" reported beside the HBM roofline.  Prints one JSON line.
#include <cuda_runtime.h>
#include <cstdio>

template <int ILP>
__global__ void __launch_bounds__(256) k_dfma(double* out, double a, double b, int iters) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 16; ++r)
#pragma unroll
      for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  if (s == 12345.678) out[0] = s;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  double* d; cudaMalloc(&d, 8);
  const int iters = 4096, ILP = 8;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 0;
  for (int blocks_per_sm = 1; blocks_per_sm <= 8; blocks_per_sm *= 2) {
    int grid = p.multiProcessorCount * blocks_per_sm;
    k_dfma<ILP><<<grid, 256>>>(d, 0.999999, 1e-7, 64);
    cudaDeviceSynchronize();
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      k_dfma<ILP><<<grid, 256>>>(d, 0.999999, 1e-7, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double fma_count = (double)grid * 256 * iters * 16 * ILP;
      double tf = 2.0 * fma_count / (ms * 1e-3) / 1e12;
      if (tf > best) best = tf;
    }
  }
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("{\"fp64_fma_tflops\": %.3f, \"sms\": %d, \"dfma_per_clk_per_sm_at_max_clock\": %.2f, \"max_clock_mhz\": %.0f, \"gpu\": \"%s\"}\n",
         best, p.multiProcessorCount, best * 1e12 / 2.0 / p.multiProcessorCount / (clk * 1e3), clk / 1e3, p.name);
  return 0;
}
