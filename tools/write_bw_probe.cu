// write_bw_probe.cu -- pure HBM write bandwidth (cudaMemset and a plain store kernel) and pure read bandwidth
#include <cstdio>
#include <cuda_runtime.h>
__global__ void store_k(double2 *p, size_t n, double v) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) p[i] = make_double2(v, v);
}
__global__ void load_k(const double2 *p, size_t n, double *out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  double s = 0;
  for (; i < n; i += stride) { double2 v = p[i]; s += v.x + v.y; }
  if (s == 1.2345) *out = s;
}
int main() {
  size_t bytes = (size_t)8 << 30;
  void *p; cudaMalloc(&p, bytes); double *o; cudaMalloc(&o, 8);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  float ms;
  for (int r = 0; r < 3; ++r) {
    cudaEventRecord(a); cudaMemsetAsync(p, 0, bytes); cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
    printf("memset      %.2f ms  %.0f GB/s\n", ms, bytes / ms / 1e6);
    cudaEventRecord(a); store_k<<<148 * 16, 256>>>((double2 *)p, bytes / 16, 1.0); cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
    printf("store kernel %.2f ms  %.0f GB/s\n", ms, bytes / ms / 1e6);
    cudaEventRecord(a); load_k<<<148 * 16, 256>>>((const double2 *)p, bytes / 16, o); cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
    printf("load kernel  %.2f ms  %.0f GB/s\n", ms, bytes / ms / 1e6);
  }
  return 0;
}
