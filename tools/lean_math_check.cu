// Accuracy of crnn_dev.cuh's lean_log / lean_exp / lean_pow on the GPU against long double on the host.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I crnn_b200/csrc -I include tools/lean_math_check.cu -o tools/lean_math_check
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "crnn_dev.cuh"
using namespace crnn;
__global__ void k(const double* x, const double* y, double* lg, double* ex, double* cl, double* ce, long n) {
  long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  lg[i] = lean_log(x[i]); ex[i] = lean_exp(y[i]); cl[i] = log(x[i]); ce[i] = exp(y[i]);
}
static double ulp_err(double got, long double ref) {
  if (std::isnan(got) && std::isnan((double)ref)) return 0;
  if (std::isinf(got) || std::isinf((double)ref) || ref == 0) return got == (double)ref ? 0 : 1e9;
  return (double)fabsl((long double)got - ref) / ldexp(1.0, ilogb((double)ref) - 52);
}
int main() {
  const long n = 1 << 24;
  std::vector<double> x(n), y(n);
  srand48(7);
  for (long i = 0; i < n; ++i) {
    double u = drand48();
    switch (i % 5) {
      case 0: x[i] = exp((u - 0.5) * 1400); y[i] = (u - 0.5) * 1398; break;
      case 1: x[i] = 0.5 + 1.5 * u; y[i] = (u - 0.5) * 2; break;
      case 2: x[i] = 1e-8 + 3 * u; y[i] = (u - 0.5) * 100; break;
      case 3: x[i] = exp((u - 0.5) * 40); y[i] = (u - 0.5) * 20; break;
      default: x[i] = 1.0 + (u - 0.5) * 1e-6; y[i] = (u - 0.5) * 1e-6; break;
    }
  }
  const double sp[] = {0.0, -1.0, 1.0, INFINITY, NAN, 4.9e-324, 2.2250738585072014e-308, 1.7976931348623157e308, -0.0, 1e-310};
  const double se[] = {0.0, -800.0, 800.0, INFINITY, -INFINITY, NAN, 709.78, -745.2, 700.0, -700.0};
  for (int i = 0; i < 10; ++i) { x[i] = sp[i]; y[i] = se[i]; }
  double *dx, *dy, *d[4];
  cudaMalloc(&dx, n * 8); cudaMalloc(&dy, n * 8);
  for (auto& p : d) cudaMalloc(&p, n * 8);
  cudaMemcpy(dx, x.data(), n * 8, cudaMemcpyHostToDevice); cudaMemcpy(dy, y.data(), n * 8, cudaMemcpyHostToDevice);
  k<<<(n + 255) / 256, 256>>>(dx, dy, d[0], d[1], d[2], d[3], n);
  std::vector<double> r[4];
  for (int j = 0; j < 4; ++j) { r[j].resize(n); cudaMemcpy(r[j].data(), d[j], n * 8, cudaMemcpyDeviceToHost); }
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("cuda error\n"); return 1; }
  double m[4] = {0, 0, 0, 0}; long special_bad = 0;
  for (long i = 0; i < n; ++i) {
    long double rl = logl((long double)x[i]), re = expl((long double)y[i]);
    if (i < 10) {  // special values must agree with the library bit for bit (they take the library path)
      if (!((r[0][i] == r[2][i]) || (std::isnan(r[0][i]) && std::isnan(r[2][i])))) ++special_bad;
      if (!((r[1][i] == r[3][i]) || (std::isnan(r[1][i]) && std::isnan(r[3][i])))) ++special_bad;
      continue;
    }
    double e;
    e = ulp_err(r[0][i], rl); if (e > m[0]) m[0] = e;
    e = ulp_err(r[2][i], rl); if (e > m[2]) m[2] = e;
    if (fabs(y[i]) < 708) { e = ulp_err(r[1][i], re); if (e > m[1]) m[1] = e; e = ulp_err(r[3][i], re); if (e > m[3]) m[3] = e; }
    else if (r[1][i] != r[3][i]) ++special_bad;
  }
  printf("{\"samples\": %ld, \"lean_log_max_ulp\": %.3f, \"lean_exp_max_ulp\": %.3f, \"cuda_log_max_ulp\": %.3f, \"cuda_exp_max_ulp\": %.3f, \"special_mismatches\": %ld}\n",
         n, m[0], m[1], m[2], m[3], special_bad);
  return (m[0] < 1.0 && m[1] < 1.0 && special_bad == 0) ? 0 : 2;
}
