// fp64_peak.cu — measures the FP64 FMA throughput of the GPU (the denominator for the FP64-bound kernels; it is not
// in MEASURED_PEAKS.json).  nvcc -cudart shared -gencode arch=compute_100a,code=sm_100a -O3 tools/fp64_peak.cu -o tools/fp64_peak
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) dfma_kernel(double *out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double b = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; i++) {
    a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
    a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int blocks = sms * 8, threads = 256, iters = 1 << 16;
  double *out;
  cudaMalloc(&out, sizeof(double) * blocks * threads);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 0;
  for (int rep = 0; rep < 6; rep++) {
    cudaEventRecord(e0);
    dfma_kernel<<<blocks, threads>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fma_per_s = double(blocks) * threads * iters * 8 / (ms * 1e-3);
    if (rep >= 1 && fma_per_s > best) best = fma_per_s;
  }
  printf("{\"fp64_fma_per_s\": %.4e, \"fp64_tflops\": %.2f, \"sms\": %d}\n", best, 2 * best / 1e12, sms);
  return cudaGetLastError() != cudaSuccess;
}
