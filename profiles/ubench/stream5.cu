// micro-benchmark: the I/O pattern of the env-step kernel (K1) without its arithmetic.
// Five streams per quad of boards: read 2 x 16 B boards + 4 B actions, write 2 x 16 B next + 16 B reward
// + 4 B flags (22 B per board), persistent grid-stride CTAs, software prefetch PF iterations ahead, and a
// knob for dummy integer work per iteration (WORK x 8 dependent LOP3/IMAD pairs over 4 chains).
// Answers: what does this access pattern reach with no compute, and how does it degrade with
// CTA shape / prefetch depth / store flavour, independent of the slide+merge arithmetic?
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint4 ldv4(const void* p, int mode) {
  uint4 v;
  if (mode == 0) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  else if (mode == 1) asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  else asm volatile("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ uint32_t ld32(const void* p) {
  uint32_t v;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void stv4(void* p, uint4 v, int mode) {
  if (mode == 0) asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
  else if (mode == 1) asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
  else if (mode == 2) asm volatile("st.global.wt.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
  else asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

struct U8 { uint4 a, b; };
__device__ __forceinline__ U8 ldv8(const void* p, int mode) {
  U8 v;
  if (mode == 0) asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(v.a.x), "=r"(v.a.y), "=r"(v.a.z), "=r"(v.a.w), "=r"(v.b.x), "=r"(v.b.y), "=r"(v.b.z), "=r"(v.b.w) : "l"(p));
  else if (mode == 1) asm volatile("ld.global.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(v.a.x), "=r"(v.a.y), "=r"(v.a.z), "=r"(v.a.w), "=r"(v.b.x), "=r"(v.b.y), "=r"(v.b.z), "=r"(v.b.w) : "l"(p));
  else asm volatile("ld.global.cs.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(v.a.x), "=r"(v.a.y), "=r"(v.a.z), "=r"(v.a.w), "=r"(v.b.x), "=r"(v.b.y), "=r"(v.b.z), "=r"(v.b.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void stv8(void* p, uint4 a, uint4 b, int mode) {
  if (mode == 0) asm volatile("st.global.cs.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
  else if (mode == 1) asm volatile("st.global.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
  else if (mode == 2) asm volatile("st.global.wt.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
  else asm volatile("st.global.L1::no_allocate.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}

template <int THREADS, int PF, int LMODE, int SMODE, bool W256 = false>
__global__ void __launch_bounds__(THREADS, 1)
k(const uint4* __restrict__ boards2, const uint32_t* __restrict__ actions4, uint4* __restrict__ next2,
  uint4* __restrict__ reward4, uint32_t* __restrict__ flags4, uint32_t nq, int work, uint32_t salt) {
  extern __shared__ unsigned char smem[];
  if (salt == 0xFFFFFFFFu) smem[threadIdx.x] = 1;   // keep the dynamic allocation alive
  const uint32_t stride = gridDim.x * THREADS;
  uint32_t quad = blockIdx.x * THREADS + threadIdx.x;
  uint4 ba[PF], bb[PF];
  uint32_t a4[PF];
#pragma unroll
  for (int p = 0; p < PF; ++p) {
    const uint32_t q = quad + p * stride;
    ba[p] = bb[p] = make_uint4(0, 0, 0, 0); a4[p] = 0;
    if (q < nq) {
      if (W256) { U8 v = ldv8(boards2 + 2u * q, LMODE); ba[p] = v.a; bb[p] = v.b; }
      else { ba[p] = ldv4(boards2 + 2u * q, LMODE); bb[p] = ldv4(boards2 + 2u * q + 1, LMODE); }
      a4[p] = ld32(actions4 + q);
    }
  }
  while (quad < nq) {
    uint4 xa = ba[0], xb = bb[0];
    uint32_t xact = a4[0];
#pragma unroll
    for (int p = 0; p + 1 < PF; ++p) { ba[p] = ba[p + 1]; bb[p] = bb[p + 1]; a4[p] = a4[p + 1]; }
    const uint32_t nxt = quad + PF * stride;
    ba[PF - 1] = bb[PF - 1] = make_uint4(0, 0, 0, 0); a4[PF - 1] = 0;
    if (nxt < nq && nxt >= quad) {
      if (W256) { U8 v = ldv8(boards2 + 2u * nxt, LMODE); ba[PF - 1] = v.a; bb[PF - 1] = v.b; }
      else { ba[PF - 1] = ldv4(boards2 + 2u * nxt, LMODE); bb[PF - 1] = ldv4(boards2 + 2u * nxt + 1, LMODE); }
      a4[PF - 1] = ld32(actions4 + nxt);
    }
    uint32_t c0 = xa.x ^ salt, c1 = xa.z, c2 = xb.x, c3 = xb.z;
    for (int w = 0; w < work; ++w) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(c0) : "r"(c1), "r"(xact));
        asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(c1) : "r"(salt), "r"(c2));
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(c2) : "r"(c3), "r"(xact));
        asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(c3) : "r"(salt), "r"(c0));
      }
    }
    if (W256) stv8(next2 + 2u * quad, make_uint4(c0, xa.y, c1, xa.w), make_uint4(c2, xb.y, c3, xb.w), SMODE);
    else {
      stv4(next2 + 2u * quad, make_uint4(c0, xa.y, c1, xa.w), SMODE);
      stv4(next2 + 2u * quad + 1, make_uint4(c2, xb.y, c3, xb.w), SMODE);
    }
    stv4(reward4 + quad, make_uint4(c0, c1, c2, c3), SMODE);
    flags4[quad] = xact ^ c0;
    quad += stride;
  }
}

template <int THREADS, int MAP, int SMODE>
__global__ void __launch_bounds__(THREADS, 1)
k8(const uint4* __restrict__ boards2, const uint32_t* __restrict__ actions4, uint4* __restrict__ next2,
   uint4* __restrict__ reward4, uint32_t* __restrict__ flags4, uint32_t noct, int work, uint32_t salt) {
  extern __shared__ unsigned char smem[];
  if (salt == 0xFFFFFFFFu) smem[threadIdx.x] = 1;
  const uint32_t stride = gridDim.x * THREADS;
  uint32_t oct = blockIdx.x * THREADS + threadIdx.x;
  auto qa_of = [](uint32_t o) { return MAP == 0 ? 2u * o : ((o >> 5) << 6) + (o & 31u); };
  auto qb_of = [](uint32_t o) { return MAP == 0 ? 2u * o + 1u : ((o >> 5) << 6) + (o & 31u) + 32u; };
  uint4 xa = make_uint4(0, 0, 0, 0), xb = xa, ya = xa, yb = xa;
  uint32_t ax = 0, ay = 0;
  if (oct < noct) { U8 v = ldv8(boards2 + 2u * qa_of(oct), 0); xa = v.a; xb = v.b; ax = ld32(actions4 + qa_of(oct)); }
  while (oct < noct) {
    { U8 v = ldv8(boards2 + 2u * qb_of(oct), 0); ya = v.a; yb = v.b; ay = ld32(actions4 + qb_of(oct)); }
    for (int half = 0; half < 2; ++half) {
      const uint4 ua = half ? ya : xa, ub = half ? yb : xb;
      const uint32_t xact = half ? ay : ax, quad = half ? qb_of(oct) : qa_of(oct);
      uint32_t c0 = ua.x ^ salt, c1 = ua.z, c2 = ub.x, c3 = ub.z;
      for (int w = 0; w < work; ++w) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(c0) : "r"(c1), "r"(xact));
          asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(c1) : "r"(salt), "r"(c2));
          asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(c2) : "r"(c3), "r"(xact));
          asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(c3) : "r"(salt), "r"(c0));
        }
      }
      stv8(next2 + 2u * quad, make_uint4(c0, ua.y, c1, ua.w), make_uint4(c2, ub.y, c3, ub.w), SMODE);
      stv4(reward4 + quad, make_uint4(c0, c1, c2, c3), SMODE);
      flags4[quad] = xact ^ c0;
      if (half == 0) {
        const uint32_t nxt = oct + stride;
        if (nxt < noct) { U8 v = ldv8(boards2 + 2u * qa_of(nxt), 0); xa = v.a; xb = v.b; ax = ld32(actions4 + qa_of(nxt)); }
      }
    }
    oct += stride;
  }
}

struct Bufs { uint4 *boards, *next, *reward; uint32_t *actions, *flags; };

template <int THREADS, int PF, int LMODE, int SMODE, bool W256 = false>
void run(const char* name, const Bufs& b, uint32_t nq, int grid, int smem, int work) {
  auto kern = k<THREADS, PF, LMODE, SMODE, W256>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int i = 0; i < 3; ++i) kern<<<grid, THREADS, smem>>>(b.boards, b.actions, b.next, b.reward, b.flags, nq, work, 17);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    for (int i = 0; i < 10; ++i) kern<<<grid, THREADS, smem>>>(b.boards, b.actions, b.next, b.reward, b.flags, nq, work, 17 + i);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 10;
    if (ms < best) best = ms;
  }
  cudaError_t e = cudaGetLastError();
  printf("%-44s grid %4d smem %6d work %3d : %.4f ms  %6.0f GB/s%s\n", name, grid, smem, work, best,
         (double)nq * 88.0 / best / 1e6, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

template <int THREADS, int MAP, int SMODE>
void run8(const char* name, const Bufs& b, uint32_t nq, int grid, int smem, int work) {
  auto kern = k8<THREADS, MAP, SMODE>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int i = 0; i < 3; ++i) kern<<<grid, THREADS, smem>>>(b.boards, b.actions, b.next, b.reward, b.flags, nq / 2, work, 17);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    for (int i = 0; i < 10; ++i) kern<<<grid, THREADS, smem>>>(b.boards, b.actions, b.next, b.reward, b.flags, nq / 2, work, 17 + i);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 10;
    if (ms < best) best = ms;
  }
  cudaError_t e = cudaGetLastError();
  printf("%-44s grid %4d smem %6d work %3d : %.4f ms  %6.0f GB/s%s\n", name, grid, smem, work, best,
         (double)nq * 88.0 / best / 1e6, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  const uint32_t nq = (1u << 26) / 4;
  Bufs b;
  cudaMalloc(&b.boards, (size_t)nq * 32); cudaMalloc(&b.next, (size_t)nq * 32); cudaMalloc(&b.reward, (size_t)nq * 16);
  cudaMalloc(&b.actions, (size_t)nq * 4); cudaMalloc(&b.flags, (size_t)nq * 4);
  cudaMemset(b.boards, 1, (size_t)nq * 32); cudaMemset(b.actions, 2, (size_t)nq * 4);
  const int BIG = 229664;   // the dynamic shared memory K1 asks for (forces one CTA per SM)
  printf("# 64 Mi boards, 22 B/board = 1476 MB per launch; work = dummy 32-instruction blocks per quad\n");
  for (int work : {0, 8, 12, 14, 16, 18}) run<1024, 1, 0, 0>("1024thr pf1 ld.nc st.cs (K1 today)", b, nq, 148, BIG, work);
  printf("# eight boards per thread and iteration (K1 v2): adjacent quads vs warp-tile mapping\n");
  for (int work : {0, 4, 6}) run8<1024, 0, 0>("OCT adjacent quads (64 B lane stride) st.cs", b, nq, 148, 229664, work);
  for (int work : {0, 4, 6}) run8<1024, 1, 0>("OCT warp tile (contiguous 1 KB) st.cs", b, nq, 148, 229664, work);
  for (int work : {0, 4, 6}) run8<1024, 0, 3>("OCT adjacent quads st.noalloc", b, nq, 148, 229664, work);
  for (int work : {0, 4, 6}) run8<1024, 1, 3>("OCT warp tile st.noalloc", b, nq, 148, 229664, work);
  printf("# 256-bit board loads / next stores (one full 32 B sector per lane and instruction)\n");
  for (int work : {0, 8, 12}) run<1024, 1, 0, 0, true>("W256 1024thr pf1 ld.nc st.cs", b, nq, 148, BIG, work);
  for (int work : {0, 8, 12}) run<1024, 1, 1, 0, true>("W256 1024thr pf1 ld.default st.cs", b, nq, 148, BIG, work);
  for (int work : {0, 8, 12}) run<1024, 1, 2, 0, true>("W256 1024thr pf1 ld.cs st.cs", b, nq, 148, BIG, work);
  for (int work : {0, 8, 12}) run<1024, 1, 0, 1, true>("W256 1024thr pf1 ld.nc st.default", b, nq, 148, BIG, work);
  for (int work : {0, 8, 12}) run<1024, 1, 0, 3, true>("W256 1024thr pf1 ld.nc st.noalloc", b, nq, 148, BIG, work);
  for (int work : {0, 8, 12}) run<1024, 1, 1, 1, true>("W256 1024thr pf1 ld.default st.default", b, nq, 148, BIG, work);
  for (int work : {0, 8, 12}) run<1024, 2, 0, 0, true>("W256 1024thr pf2 ld.nc st.cs", b, nq, 148, BIG, work);
  for (int work : {0, 8, 12}) run<1024, 2, 1, 1, true>("W256 1024thr pf2 ld.default st.default", b, nq, 148, BIG, work);
  for (int work : {0, 8, 12}) run<512, 2, 0, 0, true>("W256 512thr pf2 ld.nc st.cs", b, nq, 148, BIG, work);
  for (int work : {0}) run<256, 1, 0, 0, true>("W256 256thr grid 148*32 no smem", b, nq, 148 * 32, 0, work);
  for (int work : {0}) run<256, 1, 1, 1, true>("W256 256thr grid 148*32 no smem default", b, nq, 148 * 32, 0, work);
  printf("# 128-bit pairs (K1 round 1)\n");
  for (int work : {0, 12, 16}) run<1024, 2, 0, 0>("1024thr pf2 ld.nc st.cs", b, nq, 148, BIG, work);
  for (int work : {0, 12, 16}) run<1024, 1, 0, 1>("1024thr pf1 ld.nc st.default", b, nq, 148, BIG, work);
  for (int work : {0, 12, 16}) run<1024, 1, 0, 2>("1024thr pf1 ld.nc st.wt", b, nq, 148, BIG, work);
  for (int work : {0, 12, 16}) run<1024, 1, 0, 3>("1024thr pf1 ld.nc st.noalloc", b, nq, 148, BIG, work);
  for (int work : {0, 12, 16}) run<1024, 1, 1, 0>("1024thr pf1 ld.default st.cs", b, nq, 148, BIG, work);
  for (int work : {0, 12, 16}) run<1024, 1, 2, 0>("1024thr pf1 ld.cs st.cs", b, nq, 148, BIG, work);
  for (int work : {0, 12, 16}) run<512, 1, 0, 0>("512thr pf1 (1 CTA/SM)", b, nq, 148, BIG, work);
  for (int work : {0, 12, 16}) run<512, 2, 0, 0>("512thr pf2 (1 CTA/SM)", b, nq, 148, BIG, work);
  for (int work : {0, 12, 16}) run<256, 1, 0, 0>("256thr x many CTAs, no smem", b, nq, 148 * 8, 0, work);
  for (int work : {0, 12, 16}) run<1024, 1, 0, 0>("1024thr x 2/SM, no smem", b, nq, 148 * 2, 0, work);
  for (int work : {0}) run<256, 1, 0, 0>("256thr non-persistent-ish (grid 148*32)", b, nq, 148 * 32, 0, work);
  return 0;
}
