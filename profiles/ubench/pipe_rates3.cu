// micro-benchmark 2: pin individual SASS ops with inline PTX; 8 independent chains per thread.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define REP8(M) M(0,1,2) M(1,2,3) M(2,3,4) M(3,4,5) M(4,5,6) M(5,6,7) M(6,7,0) M(7,0,1)
template<int OP> __global__ void k(uint32_t* out, uint32_t a, uint32_t b, int iters) {
  uint32_t x[8];
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * (i + 3) + a;
  float f[8]; for (int i = 0; i < 8; ++i) f[i] = (float)x[i];
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#define LOP(i,j,k_) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(x[j]), "r"(x[k_]));
#define PRI(i,j,k_) asm volatile("prmt.b32 %0, %0, %1, 0x5410;" : "+r"(x[i]) : "r"(x[j]));
#define PRR(i,j,k_) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(x[j]), "r"(a));
#define SHR(i,j,k_) asm volatile("shr.u32 %0, %0, 1;" : "+r"(x[i]));
#define SHFV(i,j,k_) asm volatile("shr.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(b));
#define IAD(i,j,k_) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(x[j]));
#define MAD(i,j,k_) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(x[j]));
#define MHI(i,j,k_) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(a));
#define POP(i,j,k_) asm volatile("popc.b32 %0, %0;" : "+r"(x[i]));
#define FLO_(i,j,k_) asm volatile("bfind.u32 %0, %0;" : "+r"(x[i]));
#define FFM(i,j,k_) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(f[j]), "f"(f[k_]));
#define SELP(i,j,k_) asm volatile("{.reg .pred p; setp.lt.u32 p, %0, %1; selp.u32 %0, %1, %2, p;}" : "+r"(x[i]) : "r"(x[j]), "r"(x[k_]));
#define DP2(i,j,k_) asm volatile("dp2a.lo.u32.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(x[j]));
#define DP4(i,j,k_) asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(x[j]));
#define DPH(i,j,k_) asm volatile("dp2a.hi.u32.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(x[j]));
#define MX3(i,j,k_) asm volatile("{.reg .b32 t; max.u16x2 t, %0, %1; max.u16x2 %0, t, %2;}" : "+r"(x[i]) : "r"(x[j]), "r"(x[k_]));
#define MN2(i,j,k_) asm volatile("min.u16x2 %0, %0, %1;" : "+r"(x[i]) : "r"(x[j]));
#define PMAD(i,j,k_) asm volatile("{.reg .pred p; setp.ne.u32 p, %1, 0; @p mad.lo.u32 %0, %2, %2, %0;}" : "+r"(x[i]) : "r"(x[j]), "r"(a));
#define LEA_(i,j,k_) asm volatile("{.reg .b32 t; shl.b32 t, %1, 2; add.u32 %0, %0, t;}" : "+r"(x[i]) : "r"(x[j]));
#define MNX(i,j,k_) asm volatile("min.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(x[j]));
      if (OP == 0) { REP8(LOP) }
      if (OP == 1) { REP8(PRI) }
      if (OP == 2) { REP8(PRR) }
      if (OP == 3) { REP8(SHR) }
      if (OP == 4) { REP8(SHFV) }
      if (OP == 5) { REP8(IAD) }
      if (OP == 6) { REP8(MAD) }
      if (OP == 7) { REP8(MHI) }
      if (OP == 8) { REP8(POP) }
      if (OP == 9) { REP8(FLO_) }
      if (OP == 10) { REP8(FFM) }
      if (OP == 11) { REP8(SELP) }
      if (OP == 12) { REP8(DP2) }
      if (OP == 13) { LOP(0,1,2) PRI(1,2,3) LOP(2,3,4) PRI(3,4,5) LOP(4,5,6) PRI(5,6,7) LOP(6,7,0) PRI(7,0,1) }
      if (OP == 14) { MAD(0,1,2) PRI(1,2,3) MAD(2,3,4) PRI(3,4,5) MAD(4,5,6) PRI(5,6,7) MAD(6,7,0) PRI(7,0,1) }
      if (OP == 15) { LOP(0,1,2) MAD(1,2,3) LOP(2,3,4) MAD(3,4,5) LOP(4,5,6) MAD(5,6,7) LOP(6,7,0) MAD(7,0,1) }
      if (OP == 16) { FFM(0,1,2) MAD(1,2,3) FFM(2,3,4) MAD(3,4,5) FFM(4,5,6) MAD(5,6,7) FFM(6,7,0) MAD(7,0,1) }
      if (OP == 17) { LOP(0,1,2) MHI(1,2,3) LOP(2,3,4) MHI(3,4,5) LOP(4,5,6) MHI(5,6,7) LOP(6,7,0) MHI(7,0,1) }
      if (OP == 18) { LOP(0,1,2) POP(1,2,3) LOP(2,3,4) POP(3,4,5) LOP(4,5,6) POP(5,6,7) LOP(6,7,0) POP(7,0,1) }
      if (OP == 19) { LOP(0,1,2) IAD(1,2,3) LOP(2,3,4) IAD(3,4,5) LOP(4,5,6) IAD(5,6,7) LOP(6,7,0) IAD(7,0,1) }
      if (OP == 20) { REP8(MNX) }
      if (OP == 21) { MAD(0,1,2) MHI(1,2,3) MAD(2,3,4) MHI(3,4,5) MAD(4,5,6) MHI(5,6,7) MAD(6,7,0) MHI(7,0,1) }
      if (OP == 23) { REP8(DP4) }
      if (OP == 24) { REP8(DPH) }
      if (OP == 25) { REP8(MX3) }
      if (OP == 26) { REP8(MN2) }
      if (OP == 27) { LOP(0,1,2) DP2(1,2,3) LOP(2,3,4) DP2(3,4,5) LOP(4,5,6) DP2(5,6,7) LOP(6,7,0) DP2(7,0,1) }
      if (OP == 28) { MAD(0,1,2) DP2(1,2,3) MAD(2,3,4) DP2(3,4,5) MAD(4,5,6) DP2(5,6,7) MAD(6,7,0) DP2(7,0,1) }
      if (OP == 29) { LOP(0,1,2) MX3(1,2,3) LOP(2,3,4) MX3(3,4,5) LOP(4,5,6) MX3(5,6,7) LOP(6,7,0) MX3(7,0,1) }
      if (OP == 30) { MAD(0,1,2) MX3(1,2,3) MAD(2,3,4) MX3(3,4,5) MAD(4,5,6) MX3(5,6,7) MAD(6,7,0) MX3(7,0,1) }
      if (OP == 31) { REP8(PMAD) }
      if (OP == 32) { REP8(LEA_) }
      if (OP == 33) { LOP(0,1,2) MAD(1,2,3) PRI(2,3,4) MAD(3,4,5) LOP(4,5,6) DP2(5,6,7) SHR(6,7,0) MAD(7,0,1) }
      if (OP == 34) { LOP(0,1,2) LOP(1,2,3) PRI(2,3,4) MAD(3,4,5) LOP(4,5,6) DP2(5,6,7) SHR(6,7,0) LOP(7,0,1) }
      if (OP == 22) { LOP(0,1,2) SHR(1,2,3) LOP(2,3,4) SHR(3,4,5) LOP(4,5,6) SHR(5,6,7) LOP(6,7,0) SHR(7,0,1) }
    }
  }
  uint32_t r = 0; for (int i = 0; i < 8; ++i) r ^= x[i] ^ __float_as_uint(f[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template<int OP> void run(const char* name) {
  uint32_t* out; cudaMalloc(&out, 148 * 1024 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int iters = 10000;
  k<OP><<<148, 1024>>>(out, 0x00003210u, 7, 100);
  cudaEventRecord(e0); k<OP><<<148, 1024>>>(out, 0x00003210u, 7, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double wi = (double)iters * 4 * 8 * (1024 / 32) / 4;
  printf("%-22s %8.3f ms -> %.2f cyc/warp-instr/SMSP\n", name, ms, ms * 1e-3 * 1.965e9 / wi);
  cudaFree(out);
}
int main() {
  run<0>("LOP3"); run<1>("PRMT imm"); run<2>("PRMT reg"); run<3>("SHR imm"); run<4>("SHR reg"); run<5>("IADD"); run<6>("IMAD"); run<7>("IMAD.HI");
  run<8>("POPC"); run<9>("FLO"); run<10>("FFMA"); run<11>("ISETP+SEL(2)"); run<12>("IDP.2A"); run<20>("VIMNMX");
  run<13>("LOP3+PRMT mix"); run<14>("IMAD+PRMT mix"); run<15>("LOP3+IMAD mix"); run<16>("FFMA+IMAD mix"); run<17>("LOP3+IMAD.HI mix");
  run<18>("LOP3+POPC mix"); run<19>("LOP3+IADD mix"); run<21>("IMAD+IMAD.HI mix"); run<22>("LOP3+SHR mix");
  run<23>("IDP.4A"); run<24>("IDP.2A.HI"); run<25>("VIMNMX3.U16x2"); run<26>("VIMNMX.U16x2"); run<27>("LOP3+IDP mix"); run<28>("IMAD+IDP mix");
  run<29>("LOP3+VIMNMX3 mix"); run<30>("IMAD+VIMNMX3 mix"); run<31>("ISETP+@P IMAD (2)"); run<32>("LEA"); run<33>("4 ALU + 4 FMAh mix"); run<34>("6 ALU + 2 FMAh mix");
  return 0;
}
