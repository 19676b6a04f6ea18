// Accuracy of the tanh-form GELU of the GEGLU epilogue (dfb_common.cuh: gelu_tanh_f) on the hardware's tanh.approx.f32,
// against the exact erf form, next to the sigmoid form (gelu_sigmoid_f).  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I difashion_b200/csrc -I include -o tools/ubench/gelu_tanh_accuracy tools/ubench/gelu_tanh_accuracy.cu
#include "dfb_common.cuh"
#include <math.h>
#include <vector>
__global__ void k(const float* x, float* yt, float* ys, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { yt[i] = dfb::gelu_tanh_f(x[i]); ys[i] = dfb::gelu_sigmoid_f(x[i]); }
}
int main() {
  const int n = 4000001;
  std::vector<float> x(n), yt(n), ys(n);
  for (int i = 0; i < n; ++i) x[i] = -10.0f + 20.0f * (float)i / (float)(n - 1);
  float *dx, *dt, *ds;
  cudaMalloc(&dx, n * 4); cudaMalloc(&dt, n * 4); cudaMalloc(&ds, n * 4);
  cudaMemcpy(dx, x.data(), n * 4, cudaMemcpyHostToDevice);
  k<<<(n + 255) / 256, 256>>>(dx, dt, ds, n);
  cudaMemcpy(yt.data(), dt, n * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(ys.data(), ds, n * 4, cudaMemcpyDeviceToHost);
  const double edges[] = {-10, -6, -4, -3, -2, -1, 0, 1, 2, 4, 10};
  for (int b = 0; b < 10; ++b) {
    double et = 0, es = 0, rt = 0, rs = 0;
    for (int i = 0; i < n; ++i) {
      if (x[i] < edges[b] || x[i] >= edges[b + 1]) continue;
      const double ref = 0.5 * (double)x[i] * (1.0 + erf((double)x[i] / sqrt(2.0)));
      et = fmax(et, fabs(yt[i] - ref)); es = fmax(es, fabs(ys[i] - ref));
      if (fabs(ref) > 1e-3) { rt = fmax(rt, fabs(yt[i] - ref) / fabs(ref)); rs = fmax(rs, fabs(ys[i] - ref) / fabs(ref)); }
    }
    printf("g in [%5.1f, %5.1f): max abs err tanh form %.3e (rel %.2e)   sigmoid form %.3e (rel %.2e)\n", edges[b], edges[b + 1], et, rt, es, rs);
  }
  return 0;
}
