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

// DMMA fed like a GEMM inner loop: MT x 2 accumulator tiles per warp, the A fragment of every tile row loaded from
// shared memory (LDS.64) right in the instruction stream, B fragments once per k-step.
template <int MT>
__global__ void __launch_bounds__(256, 1) dmma_lds(double* out, int iters, double seed) {
  __shared__ double as[MT * 8 * 20], bs[128 * 20];
  for (int i = threadIdx.x; i < MT * 8 * 20; i += 256) as[i] = seed + i * 1e-6;
  for (int i = threadIdx.x; i < 128 * 20; i += 256) bs[i] = seed - i * 1e-6;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, fr = lane >> 2, fk = lane & 3;
  double c[MT][2][2];
#pragma unroll
  for (int i = 0; i < MT; ++i) c[i][0][0] = c[i][0][1] = c[i][1][0] = c[i][1][1] = 0.0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const double b0 = bs[(warp * 16 + fr) * 20 + ks * 4 + fk], b1 = bs[(warp * 16 + 8 + fr) * 20 + ks * 4 + fk];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const double a = as[(mt * 8 + fr) * 20 + ks * 4 + fk];
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[mt][0][0]), "+d"(c[mt][0][1]) : "d"(a), "d"(b0));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[mt][1][0]), "+d"(c[mt][1][1]) : "d"(a), "d"(b1));
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < MT; ++i) s += c[i][0][0] + c[i][0][1] + c[i][1][0] + c[i][1][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MT>
void run_lds(double* out, int sms) {
  const int iters = 2000;
  dmma_lds<MT><<<sms, 256>>>(out, 10, 0.5);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); dmma_lds<MT><<<sms, 256>>>(out, iters, 0.5); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double flop = (double)sms * 8 * iters * 4 * MT * 2 * 512.0;
  printf("DMMA + LDS.64 A fragments, 256 threads/SM, %2d x 2 tiles per warp: %8.3f ms  %6.2f TFLOP/s\n", MT, ms, flop / (ms * 1e-3) / 1e12);
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
  run_lds<4>(out, sms); run_lds<8>(out, sms); run_lds<17>(out, sms);
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("{\"f64_dmma_peak_tflops\": %.3f, \"sms\": %d, \"sm_clock_khz_attr\": %d, \"gpu\": \"%s\"}\n", best, sms, clk, p.name);
  return 0;
}
