// Stand-alone hardware check of fast_exact_draft2.cu (second draft) (K = 1 and K = 4 cells over random bytes): results against the
// floating-point loop on the host, kernel times.  nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o fast_exact_test.bin fast_exact_test.cu
#include "fast_exact_draft2.cu"
namespace fx = fx2;

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 2; } } while (0)

int main(int argc, char **argv) {
  const long n = argc > 1 ? atol(argv[1]) : 4194304;
  const int dim = 12, stride = 12, C = 512;
  fx::Tables tab;
  memset(&tab, 0, sizeof tab);
  for (int t = 1; t < 256; t++) {
    double x = (double)t / 255.0;
    int e;
    frexp(x, &e);
    tab.X[t] = (unsigned long long)ldexp(x, 60);
    tab.U[t] = 1ull << (e - 1 - 52 + 60);
  }
  CK(cudaMemcpyToSymbol(fx::c_tab, &tab, sizeof tab));
  int total_bad = 0;
  for (int mode = 0; mode < 3; mode++) {  // 0: noise K=1; 1: 6% saturated K=1; 2: noise, K=4 interleaved cells
    const int K = mode == 2 ? 4 : 1;
    std::vector<uint8_t> bytes((size_t)n * stride);
    srand(100 + mode);
    for (auto &b : bytes) {
      int r = rand();
      int t = mode == 1 && (r % 16 == 0) ? 255 : (r >> 8) & 255;
      b = (uint8_t)(t ^ 0x80);
    }
    // members: cell k holds vectors v with v % K == k, in increasing v (what a stable sort by cell gives)
    std::vector<uint32_t> order(n), beg(K + 1), seg_off(K + 1);
    {
      long p = 0;
      for (int k = 0; k < K; k++) {
        beg[k] = (uint32_t)p;
        for (long v = k; v < n; v += K) order[p++] = (uint32_t)v;
      }
      beg[K] = (uint32_t)p;
      seg_off[0] = 0;
      for (int k = 0; k < K; k++) seg_off[k + 1] = seg_off[k] + (beg[k + 1] - beg[k]) / C + 2;
    }
    // host reference: the floating-point loop per chain
    std::vector<double> want((size_t)K * dim);
    auto t0 = std::chrono::steady_clock::now();
    for (int k = 0; k < K; k++)
      for (int e = 0; e < dim; e++) {
        volatile double s = 0.0, c = 0.0;
        for (uint32_t p = beg[k]; p < beg[k + 1]; p++) {
          double x = (double)(bytes[(size_t)order[p] * stride + e] ^ 0x80) / 255.0, y = x - c, t2 = s + y;
          c = (t2 - s) - y;
          s = t2;
        }
        want[(size_t)k * dim + e] = s;
      }
    const double host_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    uint8_t *d_bytes; uint32_t *d_order, *d_beg, *d_seg_off; fx::Segs sg; fx::u128 *d_A; unsigned int *d_head, *d_reruns; double *d_sum;
    const long long nseg = seg_off[K];
    CK(cudaMalloc(&d_bytes, bytes.size())); CK(cudaMalloc(&d_order, n * 4)); CK(cudaMalloc(&d_beg, (K + 1) * 4));
    CK(cudaMalloc(&d_seg_off, (K + 1) * 4));
    const long long nrec = nseg * dim;
    CK(cudaMalloc(&sg.begin, nrec * 4)); CK(cudaMalloc(&sg.end, nrec * 4)); CK(cudaMalloc(&sg.has255, nrec)); CK(cudaMalloc(&sg.invalid, nrec));
    CK(cudaMalloc(&sg.sumX, nrec * 16)); CK(cudaMalloc(&sg.est, nrec * 16)); CK(cudaMalloc(&sg.start, nrec * 64));
    CK(cudaMalloc(&sg.margin, nrec * 64)); CK(cudaMalloc(&sg.delta, nrec * 64));
    CK(cudaMalloc(&d_A, sizeof(fx::u128) * K * dim)); CK(cudaMalloc(&d_head, 4 * K * dim)); CK(cudaMalloc(&d_reruns, 4));
    CK(cudaMalloc(&d_sum, 8 * K * dim));
    CK(cudaMemcpy(d_bytes, bytes.data(), bytes.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_order, order.data(), n * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_beg, beg.data(), (K + 1) * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_seg_off, seg_off.data(), (K + 1) * 4, cudaMemcpyHostToDevice));
    fx::Chains ch{d_bytes, (unsigned)stride, d_order, d_beg, K, dim, C, d_seg_off};
    cudaEvent_t ev[8];
    for (auto &e : ev) cudaEventCreate(&e);
    for (int rep = 0; rep < 2; rep++) {
      CK(cudaMemset(d_reruns, 0, 4));
      cudaEventRecord(ev[0]);
      fx::fx_head_kernel<<<(K * dim + 63) / 64, 64>>>(ch, d_A, d_head, d_sum);
      cudaEventRecord(ev[1]);
      fx::fx_bounds_kernel<<<(unsigned)((nrec + 127) / 128), 128>>>(ch, d_head, sg);
      cudaEventRecord(ev[2]);
      fx::fx_prefix_kernel<<<(K * dim + 63) / 64, 64>>>(ch, d_A, sg);
      cudaEventRecord(ev[3]);
      fx::fx_runs_kernel<<<(unsigned)((nrec * 4 + 127) / 128), 128>>>(ch, sg);
      cudaEventRecord(ev[4]);
      fx::fx_chain_kernel<<<(K * dim + 63) / 64, 64>>>(ch, d_A, d_head, sg, 0, d_sum, d_reruns);
      cudaEventRecord(ev[5]);
      fx::fx_runs_kernel<<<(unsigned)((nrec * 4 + 127) / 128), 128>>>(ch, sg);   // second round: the marked segments only
      cudaEventRecord(ev[6]);
      fx::fx_chain_kernel<<<(K * dim + 63) / 64, 64>>>(ch, d_A, d_head, sg, 1, d_sum, d_reruns);
      cudaEventRecord(ev[7]);
      CK(cudaDeviceSynchronize());
    }
    std::vector<double> got((size_t)K * dim);
    unsigned int reruns = 0;
    CK(cudaMemcpy(got.data(), d_sum, 8 * K * dim, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&reruns, d_reruns, 4, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (size_t i = 0; i < got.size(); i++) bad += memcmp(&got[i], &want[i], 8) != 0;
    float ms[7], total_ms = 0;
    for (int i = 0; i < 7; i++) { cudaEventElapsedTime(&ms[i], ev[i], ev[i + 1]); total_ms += ms[i]; }
    printf("mode %d: K=%d n=%ld: %d of %zu sums differ from the floating-point loop; %u of %lld segment runs re-run; "
           "kernels head %.3f bounds %.3f prefix %.3f runs %.3f chain %.3f runs2 %.3f chain2 %.3f = %.3f ms (host loop %.0f ms)\n",
           mode, K, n, bad, got.size(), reruns, nseg * dim, ms[0], ms[1], ms[2], ms[3], ms[4], ms[5], ms[6], total_ms, host_ms);
    total_bad += bad;
    cudaFree(d_bytes); cudaFree(d_order); cudaFree(d_beg); cudaFree(d_seg_off); cudaFree(sg.begin); cudaFree(sg.end); cudaFree(sg.has255); cudaFree(sg.invalid); cudaFree(sg.sumX); cudaFree(sg.est); cudaFree(sg.start); cudaFree(sg.margin); cudaFree(sg.delta); cudaFree(d_A);
    cudaFree(d_head); cudaFree(d_reruns); cudaFree(d_sum);
  }
  return total_bad != 0;
}
