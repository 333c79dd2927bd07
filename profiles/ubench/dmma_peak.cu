// micro-benchmark: FP64 tensor-core peak of this GPU as seen by mma.sync.m8n8k4.f64 (DMMA.8x8x4), the
// instruction K6 / K7 are built on (there is no tcgen05 kind for float64).  Every warp keeps ACC
// independent accumulator tiles in flight; prints TFLOP/s for several occupancies plus a plain DFMA line.
// The best figure is the denominator of bench.py's `ddqn.*.frac_of_f64_peak`.
#include <cstdio>
#include <cuda_runtime.h>

template <int ACC>
__global__ void __launch_bounds__(1024, 1) dmma(double* out, int iters, double seed) {
  double c[ACC][2];
#pragma unroll
  for (int i = 0; i < ACC; ++i) c[i][0] = c[i][1] = seed * (i + 1);
  double a = seed + threadIdx.x * 1e-9, b = 1.0 - seed;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ACC; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ACC; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(1024, 1) dfma(double* out, int iters, double seed) {
  double c[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i] = seed * (i + 1);
  const double a = 1.0 + seed * 1e-9, b = seed;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ACC>
double run(double* out, int sms, int threads) {
  const int iters = 20000;
  dmma<ACC><<<sms, threads>>>(out, 100, 0.5);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9f;
  for (int r = 0; r < 3; ++r) {
    cudaEventRecord(e0); dmma<ACC><<<sms, threads>>>(out, iters, 0.5); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  const double flop = (double)sms * (threads / 32) * iters * ACC * 512.0;
  const double tf = flop / (best * 1e-3) / 1e12;
  printf("DMMA.8x8x4  %4d threads/SM x %2d accumulators: %8.3f ms  %6.2f TFLOP/s\n", threads, ACC, best, tf);
  return tf;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  double* out; cudaMalloc(&out, sizeof(double) * sms * 1024);
  double best = 0;
  for (int threads : {128, 256, 512, 1024}) {
    double t = run<4>(out, sms, threads); if (t > best) best = t;
    t = run<8>(out, sms, threads); if (t > best) best = t;
    t = run<16>(out, sms, threads); if (t > best) best = t;
  }
  {
    const int iters = 20000;
    dfma<<<sms, 1024>>>(out, 100, 0.5);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); dfma<<<sms, 1024>>>(out, iters, 0.5); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("DFMA (CUDA cores) 1024 threads/SM x 8 chains: %8.3f ms  %6.2f TFLOP/s\n", ms, (double)sms * 1024 * iters * 8 * 2 / (ms * 1e-3) / 1e12);
  }
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("{\"f64_dmma_peak_tflops\": %.3f, \"sms\": %d, \"sm_clock_khz_attr\": %d, \"gpu\": \"%s\"}\n", best, sms, clk, p.name);
  return 0;
}
