// Microbenchmark for the north_star clause "tensor cores only if the FP64 DMMA path beats the shared-memory kernel":
// peak issue rate of the FP64 FMA pipe (DFMA) against the FP64 tensor path (mma.sync m8n8k4 f64 -> SASS DMMA) on one B200.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_pipes fp64_pipes.cu && ./fp64_pipes
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_kernel(double *out, int iters, double a, double b) {
  double acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dmma_kernel(double *out, int iters, double a, double b) {
  // m8n8k4: per thread A 1 double, B 1 double, C/D 2 doubles; 4 independent accumulator tiles
  double c[4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i) c[i][0] = c[i][1] = threadIdx.x * 1e-3 + i;
  double fa = a + threadIdx.x * 1e-6, fb = b;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1])
                   : "d"(fa), "d"(fb));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int threads = 256, blocks = sms * 8, iters = 20000;
  double *out;
  cudaMalloc(&out, sizeof(double) * threads * blocks);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float ms;
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0);
    dfma_kernel<<<blocks, threads>>>(out, iters, 0.999999, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    const double fl = 2.0 * 8 * iters * (double)threads * blocks;
    if (rep) printf("{\"pipe\": \"DFMA\", \"tflops\": %.2f, \"ms\": %.3f}\n", fl / ms * 1e-9, ms);
    cudaEventRecord(e0);
    dmma_kernel<<<blocks, threads>>>(out, iters, 0.999999, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    const double fl2 = 2.0 * 8 * 8 * 4 * 4 * iters * (double)(threads / 32) * blocks;  // 512 flop per warp-level mma, 4 per iteration
    if (rep) printf("{\"pipe\": \"DMMA m8n8k4\", \"tflops\": %.2f, \"ms\": %.3f}\n", fl2 / ms * 1e-9, ms);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
