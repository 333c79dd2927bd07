// micro-benchmark: per-SMSP issue rate of candidate instructions (B200)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template<int OP> __global__ void k(uint32_t* out, uint32_t a, uint32_t b, int iters) {
  uint32_t x0 = threadIdx.x + a, x1 = x0 ^ 0x1234567, x2 = x0 * 3 + 1, x3 = x0 + 77, x4 = x0 ^ b, x5 = x1 + b, x6 = x2 ^ a, x7 = x3 + a;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (OP == 0) { x0 = (x0 & a) ^ b; x1 = (x1 & a) ^ b; x2 = (x2 & a) ^ b; x3 = (x3 & a) ^ b; x4 = (x4 & a) ^ b; x5 = (x5 & a) ^ b; x6 = (x6 & a) ^ b; x7 = (x7 & a) ^ b; }           // LOP3
      if (OP == 1) { x0 = x0 * a + b; x1 = x1 * a + b; x2 = x2 * a + b; x3 = x3 * a + b; x4 = x4 * a + b; x5 = x5 * a + b; x6 = x6 * a + b; x7 = x7 * a + b; }                         // IMAD
      if (OP == 2) { x0 = __umulhi(x0, a); x1 = __umulhi(x1, a); x2 = __umulhi(x2, a); x3 = __umulhi(x3, a); x4 = __umulhi(x4, a); x5 = __umulhi(x5, a); x6 = __umulhi(x6, a); x7 = __umulhi(x7, a); }  // IMAD.HI
      if (OP == 3) { x0 = __dp2a_lo(x0, a, x0); x1 = __dp2a_lo(x1, a, x1); x2 = __dp2a_lo(x2, a, x2); x3 = __dp2a_lo(x3, a, x3); x4 = __dp2a_lo(x4, a, x4); x5 = __dp2a_lo(x5, a, x5); x6 = __dp2a_lo(x6, a, x6); x7 = __dp2a_lo(x7, a, x7); } // IDP.2A
      if (OP == 4) { x0 = __byte_perm(x0, a, x1); x1 = __byte_perm(x1, a, x2); x2 = __byte_perm(x2, a, x3); x3 = __byte_perm(x3, a, x4); x4 = __byte_perm(x4, a, x5); x5 = __byte_perm(x5, a, x6); x6 = __byte_perm(x6, a, x7); x7 = __byte_perm(x7, a, x0); } // PRMT
      if (OP == 5) { x0 = __funnelshift_r(x0, x1, b); x1 = __funnelshift_r(x1, x2, b); x2 = __funnelshift_r(x2, x3, b); x3 = __funnelshift_r(x3, x4, b); x4 = __funnelshift_r(x4, x5, b); x5 = __funnelshift_r(x5, x6, b); x6 = __funnelshift_r(x6, x7, b); x7 = __funnelshift_r(x7, x0, b); } // SHF
      if (OP == 6) { x0 = __popc(x0) + x1; x1 = __popc(x1) + x2; x2 = __popc(x2) + x3; x3 = __popc(x3) + x4; x4 = __popc(x4)+x5; x5 = __popc(x5)+x6; x6 = __popc(x6)+x7; x7 = __popc(x7)+x0; } // POPC + IADD
      if (OP == 7) { uint64_t p0 = (uint64_t)x0 * a, p1 = (uint64_t)x1 * a, p2 = (uint64_t)x2 * a, p3 = (uint64_t)x3 * a; x0 = (uint32_t)(p0>>32) ^ x4; x4 = (uint32_t)p0; x1 = (uint32_t)(p1>>32)^x5; x5 = (uint32_t)p1; x2 = (uint32_t)(p2>>32)^x6; x6=(uint32_t)p2; x3 = (uint32_t)(p3>>32)^x7; x7=(uint32_t)p3; } // IMAD.WIDE + LOP
      if (OP == 8) { x0 = (x0 & a) + b; x1 = (x1 * a) + b; x2 = (x2 & a) + b; x3 = (x3 * a) + b; x4 = (x4 & a) + b; x5 = (x5 * a) + b; x6 = (x6 & a) + b; x7 = (x7 * a) + b; } // mix LOP3+IADD vs IMAD
      if (OP == 9) { x0 = min(x0, a) ^ b; x1 = min(x1, a) ^ b; x2 = min(x2,a)^b; x3 = min(x3,a)^b; x4 = min(x4,a)^b; x5=min(x5,a)^b; x6=min(x6,a)^b; x7=min(x7,a)^b; }
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7;
}
template<int OP> void run(const char* name, int opsPerIter) {
  uint32_t* out; cudaMalloc(&out, 148 * 1024 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int iters = 20000;
  k<OP><<<148, 1024>>>(out, 0xF0F0F0F1u, 7, 100);
  cudaEventRecord(e0); k<OP><<<148, 1024>>>(out, 0xF0F0F0F1u, 7, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double warp_instr_per_smsp = (double)iters * 8 * opsPerIter * (1024 / 32) / 4;   // per SMSP
  double cycles = ms * 1e-3 * 1.965e9;
  printf("%-28s %.3f ms  -> %.2f cycles per warp-instr per SMSP (assuming 1965 MHz)\n", name, ms, cycles / warp_instr_per_smsp);
}
int main() {
  run<0>("LOP3 (x&a)^b", 8); run<1>("IMAD", 8); run<2>("IMAD.HI", 8); run<3>("IDP.2A", 8); run<4>("PRMT", 8); run<5>("SHF funnel", 8);
  run<6>("POPC+IADD (2 ops)", 16); run<7>("IMAD.WIDE+LOP3 (4 pairs)", 8); run<8>("mix 4x(LOP3+IADD)+4xIMAD", 12); run<9>("VIMNMX+LOP3", 16);
  return 0;
}
