// Dependent-issue latency of DADD / DFMA on one warp (clock64 around a chain of N dependent operations).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dadd_latency_probe dadd_latency_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void probe(double *out, long long *cycles, double a, double b, int n) {
  double x = a, y = b;
  long long t0 = clock64();
  for (int i = 0; i < n; i++) {
#pragma unroll
    for (int u = 0; u < 16; u++) x = __dadd_rn(x, y);
  }
  long long t1 = clock64();
  for (int i = 0; i < n; i++) {
#pragma unroll
    for (int u = 0; u < 16; u++) y = __fma_rn(y, a, x);
  }
  long long t2 = clock64();
  float f = (float)a, g = (float)b;
  for (int i = 0; i < n; i++) {
#pragma unroll
    for (int u = 0; u < 16; u++) f = __fadd_rn(f, g);
  }
  long long t3 = clock64();
  out[threadIdx.x] = x + y + f;
  if (threadIdx.x == 0) { cycles[0] = t1 - t0; cycles[1] = t2 - t1; cycles[2] = t3 - t2; }
}
int main() {
  double *d; long long *c, h[3];
  cudaMalloc(&d, 1024 * 8); cudaMalloc(&c, 24);
  for (int threads : {32, 128, 1024}) {
    probe<<<1, threads>>>(d, c, 1.0000001, 1e-9, 4096);
    cudaMemcpy(h, c, 24, cudaMemcpyDeviceToHost);
    printf("threads %4d: DADD %.1f  DFMA %.1f  FADD %.1f cycles per dependent op\n", threads, h[0] / 65536.0, h[1] / 65536.0, h[2] / 65536.0);
  }
  return 0;
}
