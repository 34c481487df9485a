// Micro-benchmark: sustained issue rate of FP32 FMA forms on sm_100a (one 512-thread CTA per SM).
// Prints warp-instructions per cycle per SM sub-partition and the equivalent FMA lanes/clk/SM.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float lo, float hi) { u64 d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi)); return d; }
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float lo32(u64 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a + b; }

#define ITERS 2048
template <int V>
__global__ void __launch_bounds__(512, 1) probe(const float *in, float *out, long long *cyc, float m, float c) {
  float a[8], b[8];
  for (int i = 0; i < 8; i++) { a[i] = in[threadIdx.x + 32 * i]; b[i] = in[threadIdx.x + 512 + 32 * i]; }
  float acc[8][4];
  u64 accp[4][4];
  for (int i = 0; i < 8; i++) for (int j = 0; j < 4; j++) acc[i][j] = 0.f;
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) accp[i][j] = 0ull;
  u64 ap[4], bp[4];
  for (int i = 0; i < 4; i++) { ap[i] = pack2(a[2 * i], a[2 * i + 1]); bp[i] = pack2(b[2 * i], b[2 * i + 1]); }
  float mn[8];
  for (int i = 0; i < 8; i++) mn[i] = a[i];
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; it++) {
    if (V == 1) {  // scalar FFMA, 8x4 register tile, all register operands
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    } else if (V == 2) {  // FFMA2: pair(a) x broadcast(b)
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) accp[i][j] = ffma2(ap[i], pack2(b[j], b[j]), accp[i][j]);
    } else if (V == 3) {  // FFMA2: pair x pair
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) accp[i][j] = ffma2(ap[i], bp[j], accp[i][j]);
    } else if (V == 4) {  // scalar FFMA with constant-bank multiplier and addend (the peak probe form)
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = fmaf(acc[i][j], m, c);
    } else if (V == 5) {  // scalar FFMA: register x constant + register
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], m, acc[i][j]);
    } else if (V == 6) {  // FFMA2 pair x broadcast + min/max mix (16 FFMA2 : 10 FMNMX)
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) accp[i][j] = ffma2(ap[i], pack2(b[j], b[j]), accp[i][j]);
#pragma unroll
      for (int i = 0; i < 8; i++) mn[i] = fminf(mn[i], fmaxf(a[i], b[(i + it) & 7]));
      mn[0] = fminf(mn[0], mn[1]); mn[2] = fmaxf(mn[2], mn[3]);
    } else if (V == 7) {  // min/max only (alu pipe rate): 32 FMNMX
#pragma unroll
      for (int r = 0; r < 4; r++)
#pragma unroll
        for (int i = 0; i < 8; i++) mn[i] = fminf(fmaxf(mn[i], b[(i + r) & 7]), a[(i + r + 1) & 7]);
    } else if (V == 8) {  // FFMA2 with the accumulator chain of the assign kernel: 4 chains of dependent ops
#pragma unroll
      for (int e = 0; e < 4; e++)
#pragma unroll
        for (int ch = 0; ch < 4; ch++) accp[0][ch] = ffma2(ap[e], pack2(b[ch + (e & 1) * 4], b[ch + (e & 1) * 4]), accp[0][ch]);
    } else if (V == 9) {  // scalar FFMA 8x4 tile with FADD-free "two-source" form: acc = a*b + acc, b from constant
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = fmaf(acc[i][j], b[j], a[i]);
    }
  }
  long long t1 = clock64();
  float s = 0.f;
  for (int i = 0; i < 8; i++) { s += mn[i]; for (int j = 0; j < 4; j++) s += acc[i][j]; }
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) s += lo32(accp[i][j]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int V>
void run(const char *name, double warp_instr_per_iter, double fma_per_instr, const float *in, float *out, long long *cyc, int sms) {
  probe<V><<<sms, 512>>>(in, out, cyc, 0.999f, 1e-4f);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  probe<V><<<sms, 512>>>(in, out, cyc, 0.999f, 1e-4f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long h[256]; cudaMemcpy(h, cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < sms; i++) avg += h[i]; avg /= sms;
  // per SMSP: 4 warps, each issues warp_instr_per_iter * ITERS instructions
  double ipc = 4.0 * warp_instr_per_iter * ITERS / avg;
  printf("%-52s cycles/CTA %10.0f  instr/clk/SMSP %.3f  FMA lanes/clk/SM %.1f  (%.3f ms, err=%s)\n", name, avg, ipc,
         ipc * 4 * 32 * fma_per_instr, ms, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  printf("%s, %d SMs, clock %d kHz\n", p.name, sms, p.clockRate);
  float *in, *out; long long *cyc;
  cudaMalloc(&in, 4096 * 4); cudaMalloc(&out, sms * 512 * 4); cudaMalloc(&cyc, 256 * 8);
  float h[4096]; for (int i = 0; i < 4096; i++) h[i] = 1.0f + (i % 17) * 1e-3f;
  cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice);
  run<1>("V1 FFMA  reg*reg+reg (8x4 tile)", 32, 1, in, out, cyc, sms);
  run<2>("V2 FFMA2 pair*bcast+pair (4x4 tile)", 16, 2, in, out, cyc, sms);
  run<3>("V3 FFMA2 pair*pair+pair (4x4 tile)", 16, 2, in, out, cyc, sms);
  run<4>("V4 FFMA  reg*const+const", 32, 1, in, out, cyc, sms);
  run<5>("V5 FFMA  reg*const+reg", 32, 1, in, out, cyc, sms);
  run<6>("V6 16 FFMA2 + 18 FMNMX mix (instr counted: 34)", 34, 16.0 * 2 / 34, in, out, cyc, sms);
  run<7>("V7 FMNMX only (64 per iter)", 64, 0, in, out, cyc, sms);
  run<8>("V8 FFMA2 4 dependent chains (16 per iter)", 16, 2, in, out, cyc, sms);
  run<9>("V9 FFMA  acc*reg+reg", 32, 1, in, out, cyc, sms);
  return 0;
}
