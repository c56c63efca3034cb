// los_peaks.cu -- FP64 roofline probe: 16 independent DFMA chains per thread, no memory traffic.
#include <cuda_runtime.h>
namespace songlos {
__global__ void __launch_bounds__(256) dfma_probe_kernel(double *out, int iters)
{
  double a[16];
  double x = 1.0 + 1e-9 * threadIdx.x, y = 1e-12 * (blockIdx.x + 1);
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = i * 0.125;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fma(a[i], x, y);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) dmma_probe_kernel(double *out, int iters)
{
  double c0[8], c1[8];
  double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-3 * (threadIdx.x % 7);
#pragma unroll
  for (int i = 0; i < 8; ++i) { c0[i] = i; c1[i] = -i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c0[i] + c1[i];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}
cudaError_t launch_dmma_probe(double *out, int iters, int blocks, int threads, cudaStream_t st)
{
  dmma_probe_kernel<<<blocks, threads, 0, st>>>(out, iters);
  return cudaGetLastError();
}
cudaError_t launch_dfma_probe(double *out, int iters, int blocks, int threads, cudaStream_t st)
{
  dfma_probe_kernel<<<blocks, threads, 0, st>>>(out, iters);
  return cudaGetLastError();
}
}  // namespace songlos
