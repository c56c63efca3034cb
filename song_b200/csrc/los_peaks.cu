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
cudaError_t launch_dfma_probe(double *out, int iters, int blocks, int threads, cudaStream_t st)
{
  dfma_probe_kernel<<<blocks, threads, 0, st>>>(out, iters);
  return cudaGetLastError();
}
}  // namespace songlos
