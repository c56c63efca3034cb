// fp64_probe.cu -- B200 micro-probes that decide the FAST kernel design:
//   (1) DFMA peak (16 independent chains/thread)
//   (2) DMMA m8n8k4 f64 peak (8 independent accumulators/warp)
//   (3) DFMA with one operand from shared memory per FMA (LDS.64 broadcast / per-lane)
//   (4) DMMA with A from LDS.64 and B from LDG.64 per MMA (operand-delivery-bound variant)
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("err %s line %d\n",cudaGetErrorString(e),__LINE__);return 1;}}while(0)

__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256) k_dfma(double *out, int iters) {
  double a[16]; double x = 1.0 + 1e-9 * threadIdx.x, y = 1e-12 * (blockIdx.x + 1);
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

template <int NACC>
__global__ void __launch_bounds__(256) k_dmma(double *out, int iters) {
  double c0[NACC], c1[NACC];
  double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-3 * (threadIdx.x % 7);
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c0[i] = i; c1[i] = -i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) dmma(c0[i], c1[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c0[i] + c1[i];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// DFMA, register tile RN x RL, operands from smem each k-step: J per lane (LDS.128), S broadcast
template <int RN, int RL>
__global__ void __launch_bounds__(256) k_dfma_tile(double *out, int iters) {
  __shared__ double2 Js[64 * 32 * RL / RL > 0 ? 2048 : 1];
  __shared__ double Ss[1024];
  for (int i = threadIdx.x; i < 2048; i += 256) Js[i] = make_double2(1e-3 * i, 1e-4 * i);
  for (int i = threadIdx.x; i < 1024; i += 256) Ss[i] = 1.0 + 1e-6 * i;
  __syncthreads();
  double a0[RN][RL], a1[RN][RL];
#pragma unroll
  for (int n = 0; n < RN; ++n)
#pragma unroll
    for (int l = 0; l < RL; ++l) { a0[n][l] = 0; a1[n][l] = 0; }
  int lane = threadIdx.x & 31;
  for (int it = 0; it < iters; ++it) {
    int kk = it & 15;
    double2 j[RL];
#pragma unroll
    for (int l = 0; l < RL; ++l) j[l] = Js[(kk * RL + l) * 32 + lane];
    double s[RN];
#pragma unroll
    for (int n = 0; n < RN; ++n) s[n] = Ss[kk * RN + n + (threadIdx.x >> 5) * 64 % 512];
#pragma unroll
    for (int n = 0; n < RN; ++n)
#pragma unroll
      for (int l = 0; l < RL; ++l) { a0[n][l] = fma(j[l].x, s[n], a0[n][l]); a1[n][l] = fma(j[l].y, s[n], a1[n][l]); }
  }
  double sum = 0;
#pragma unroll
  for (int n = 0; n < RN; ++n)
#pragma unroll
    for (int l = 0; l < RL; ++l) sum += a0[n][l] + a1[n][l];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = sum;
}

// DMMA with operands fetched per MMA group: NA A-frags from smem, NB B-frags from global (L1/L2 resident)
template <int NA, int NB>
__global__ void __launch_bounds__(256) k_dmma_mem(double *out, const double *Bg, int iters) {
  __shared__ double As[4096];
  for (int i = threadIdx.x; i < 4096; i += 256) As[i] = 1.0 + 1e-6 * i;
  __syncthreads();
  double c0[NA][NB], c1[NA][NB];
#pragma unroll
  for (int i = 0; i < NA; ++i)
#pragma unroll
    for (int j = 0; j < NB; ++j) { c0[i][j] = 0; c1[i][j] = 0; }
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const double *Bw = Bg + (size_t)(blockIdx.x % 64) * 8192;
  for (int it = 0; it < iters; ++it) {
    int kk = it & 7;
    double a[NA], b[NB];
#pragma unroll
    for (int i = 0; i < NA; ++i) a[i] = As[((kk * NA + i) * 32 + lane + warp * 64) & 4095];
#pragma unroll
    for (int j = 0; j < NB; ++j) b[j] = __ldg(&Bw[((kk * NB + j) * 32 + lane) & 8191]);
#pragma unroll
    for (int i = 0; i < NA; ++i)
#pragma unroll
      for (int j = 0; j < NB; ++j) dmma(c0[i][j], c1[i][j], a[i], b[j]);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NA; ++i)
#pragma unroll
    for (int j = 0; j < NB; ++j) s += c0[i][j] + c1[i][j];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F> float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) { cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount; printf("device %s SMs %d\n", p.name, sms);
  double *out, *Bg; CK(cudaMalloc(&out, (size_t)sms * 16 * 256 * 8)); CK(cudaMalloc(&Bg, 64 * 8192 * 8)); CK(cudaMemset(Bg, 0, 64 * 8192 * 8));
  const int iters = 8192;
  for (int bps : {1, 2, 4, 8}) {
    int blocks = sms * bps;
    float ms = timeit([&] { k_dfma<<<blocks, 256>>>(out, iters); });
    printf("DFMA  blocks/SM %d : %.2f TFLOP/s\n", bps, 2.0 * 16 * iters * (double)blocks * 256 / (ms * 1e-3) / 1e12);
    ms = timeit([&] { k_dmma<8><<<blocks, 256>>>(out, iters); });
    printf("DMMA8 blocks/SM %d : %.2f TFLOP/s\n", bps, 2.0 * 256 * 8 * iters * (double)blocks * 8 / (ms * 1e-3) / 1e12);
    ms = timeit([&] { k_dmma<16><<<blocks, 256>>>(out, iters); });
    printf("DMMA16 blocks/SM %d : %.2f TFLOP/s\n", bps, 2.0 * 256 * 16 * iters * (double)blocks * 8 / (ms * 1e-3) / 1e12);
  }
  for (int bps : {1, 2}) {
    int blocks = sms * bps;
    float ms;
    ms = timeit([&] { k_dfma_tile<8, 2><<<blocks, 256>>>(out, iters); });
    printf("DFMA tile 8x2 blocks/SM %d : %.2f TFLOP/s\n", bps, 2.0 * 2 * 16 * iters * (double)blocks * 256 / (ms * 1e-3) / 1e12);
    ms = timeit([&] { k_dfma_tile<8, 4><<<blocks, 256>>>(out, iters); });
    printf("DFMA tile 8x4 blocks/SM %d : %.2f TFLOP/s\n", bps, 2.0 * 2 * 32 * iters * (double)blocks * 256 / (ms * 1e-3) / 1e12);
    ms = timeit([&] { k_dfma_tile<4, 4><<<blocks, 256>>>(out, iters); });
    printf("DFMA tile 4x4 blocks/SM %d : %.2f TFLOP/s\n", bps, 2.0 * 2 * 16 * iters * (double)blocks * 256 / (ms * 1e-3) / 1e12);
    ms = timeit([&] { k_dmma_mem<1, 4><<<blocks, 256>>>(out, Bg, iters); });
    printf("DMMA mem A1xB4 blocks/SM %d : %.2f TFLOP/s\n", bps, 2.0 * 256 * 4 * iters * (double)blocks * 8 / (ms * 1e-3) / 1e12);
    ms = timeit([&] { k_dmma_mem<2, 4><<<blocks, 256>>>(out, Bg, iters); });
    printf("DMMA mem A2xB4 blocks/SM %d : %.2f TFLOP/s\n", bps, 2.0 * 256 * 8 * iters * (double)blocks * 8 / (ms * 1e-3) / 1e12);
    ms = timeit([&] { k_dmma_mem<2, 8><<<blocks, 256>>>(out, Bg, iters); });
    printf("DMMA mem A2xB8 blocks/SM %d : %.2f TFLOP/s\n", bps, 2.0 * 256 * 16 * iters * (double)blocks * 8 / (ms * 1e-3) / 1e12);
    ms = timeit([&] { k_dmma_mem<1, 8><<<blocks, 256>>>(out, Bg, iters); });
    printf("DMMA mem A1xB8 blocks/SM %d : %.2f TFLOP/s\n", bps, 2.0 * 256 * 8 * iters * (double)blocks * 8 / (ms * 1e-3) / 1e12);
  }
  return 0;
}
