// dmma_occupancy_probe.cu -- how many warps per SM sub-partition does the FP64 tensor pipe need?
// One CTA per SM with W warps; every warp loops over NACC independent m8n8k4 accumulators, operands
// either in registers or re-read from shared memory (LDS.64) each step like the LOS kernel does.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("err %s line %d\n",cudaGetErrorString(e),__LINE__);return 1;}}while(0)
__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int NACC, bool LDS>
__global__ void k(double *out, int iters) {
  __shared__ double sm[32 * 64];
  for (int i = threadIdx.x; i < 32 * 64; i += blockDim.x) sm[i] = 1e-3 * (i % 13);
  __syncthreads();
  double c0[NACC], c1[NACC];
  int lane = threadIdx.x & 31;
  double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-3 * (threadIdx.x % 7);
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c0[i] = i; c1[i] = -i; }
  for (int it = 0; it < iters; ++it) {
    if (LDS) {
      double bf[NACC / 2];
      double a0 = sm[((it & 7) * 2) * 32 + lane], a1 = sm[((it & 7) * 2 + 1) * 32 + lane];
#pragma unroll
      for (int i = 0; i < NACC / 2; ++i) bf[i] = sm[(16 + ((it + i) & 31)) * 32 + lane];
#pragma unroll
      for (int i = 0; i < NACC / 2; ++i) { dmma(c0[i], c1[i], a0, bf[i]); dmma(c0[i + NACC / 2], c1[i + NACC / 2], a1, bf[i]); }
    } else {
#pragma unroll
      for (int i = 0; i < NACC; ++i) dmma(c0[i], c1[i], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c0[i] + c1[i];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC, bool LDS>
int run(int warps, int sms, double *out) {
  int iters = 20000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<NACC, LDS><<<sms, warps * 32>>>(out, 100);
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    cudaEventRecord(e0); k<NACC, LDS><<<sms, warps * 32>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  double flops = 512.0 * NACC * iters * (double)warps * sms;
  printf("NACC=%2d %s warps/SM=%2d : %6.2f TFLOP/s  (%.1f clk per DMMA per SMSP at 1.965 GHz)\n", NACC, LDS ? "LDS" : "reg", warps,
         flops / best / 1e9, best * 1e-3 * 1.965e9 / (NACC * (double)iters * warps / 4.0));
  return 0;
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  double *out; CK(cudaMalloc(&out, (size_t)p.multiProcessorCount * 1024 * 8));
  for (int w : {4, 8, 12, 16, 32}) { run<16, false>(w, p.multiProcessorCount, out); }
  for (int w : {4, 8, 12, 16, 32}) { run<16, true>(w, p.multiProcessorCount, out); }
  for (int w : {4, 8, 16}) { run<8, false>(w, p.multiProcessorCount, out); }
  for (int w : {4, 8, 16}) { run<32, true>(w, p.multiProcessorCount, out); }
  return 0;
}
