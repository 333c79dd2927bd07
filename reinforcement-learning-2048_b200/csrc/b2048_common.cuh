// b2048_common.cuh — shared device helpers: Philox4x32-10, nibble SWAR, row-table access.
// sm_100a only (see build.py); no other architecture is compiled.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/b2048.h"

namespace b2048 {

// ---- row table geometry ---------------------------------------------------------------------
// Full table: 65536 x u32 in global memory (256 KB, L2-resident).  The streaming step kernel
// stages the first LUT_SMEM_ROWS entries (rows whose top nibble is < 14) into shared memory:
// 57344 * 4 B = 224 KB of the 227 KB a CTA may own.  Rows holding a 16384/32768 tile in their
// last position are looked up in the global table instead (never on the benchmark
// distribution, vanishingly rare in real games, still exact).
constexpr int LUT_ROWS = 65536;
constexpr int LUT_SMEM_ROWS = 57344;
constexpr int LUT_SMEM_BYTES = LUT_SMEM_ROWS * 4;

// Philox key domains (xored into the high key word) so that streams never collide.
enum : uint32_t {
  DOM_SPAWN = 0x00000000u,
  DOM_RESET = 0x5BD1E995u,
  DOM_BOARDS = 0x1B873593u,
  DOM_ACTIONS = 0xCC9E2D51u,
  DOM_SAMPLE = 0x85EBCA6Bu,
  DOM_EGREEDY = 0xC2B2AE35u,
};

struct DeviceCtx {
  uint32_t* lut = nullptr;       // [65536] row table
  double* partials = nullptr;    // [MAX_PARTIALS] loss partial sums
  unsigned int* ticket = nullptr;// last-block-done counter
  int sm_count = 0;
  int max_smem_optin = 0;
  bool ready = false;
  // host-API workspace (lazy)
  void* ws = nullptr;
  size_t ws_bytes = 0;
  cudaStream_t ws_streams[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t ws_events[3] = {nullptr, nullptr, nullptr};
};
constexpr int MAX_DEVICES = 16;
constexpr int MAX_PARTIALS = 1024;

// defined in host_api.cu
DeviceCtx* current_ctx(int* err);
DeviceCtx* ctx_for(int device);

// ---- Philox4x32-10 (Salmon et al., SC'11; same constants as Random123 / cuRAND) -----------------
__host__ __device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c.x;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c.z;
    const uint32_t nx = (uint32_t)(p1 >> 32) ^ c.y ^ k0;
    const uint32_t nz = (uint32_t)(p0 >> 32) ^ c.w ^ k1;
    c.y = (uint32_t)p1;
    c.w = (uint32_t)p0;
    c.x = nx;
    c.z = nz;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return c;
}

__host__ __device__ __forceinline__ uint4 philox_at(uint64_t seed, uint32_t domain, uint64_t idx,
                                                    uint64_t step) {
  return philox4x32_10(make_uint4((uint32_t)idx, (uint32_t)(idx >> 32), (uint32_t)step,
                                  (uint32_t)(step >> 32)),
                       (uint32_t)seed, (uint32_t)(seed >> 32) ^ domain);
}

// ---- nibble SWAR ------------------------------------------------------------------------------
// bit 3 of every nibble set iff the nibble is non-zero (carry-free: 7+7 < 16).
__device__ __forceinline__ uint32_t nz3(uint32_t v) {
  return (((v & 0x77777777u) + 0x77777777u) | v) & 0x88888888u;
}

// Legal-move mask (bits: up, down, left, right) of a board, without the row table: a move toward
// a side is legal iff some adjacent pair along that axis has (near cell empty, far cell filled)
// or (both filled and equal).  Equivalent to "the move changes the board" (src/board.py:128-135).
__device__ __forceinline__ uint32_t legal_mask(uint32_t lo, uint32_t hi) {
  const uint32_t n_lo = nz3(lo), n_hi = nz3(hi);
  // horizontal pairs (c, c+1) sit at nibble c, c < 3
  const uint32_t ne_lo = nz3(lo ^ (lo >> 4)), ne_hi = nz3(hi ^ (hi >> 4));
  const uint32_t ns_lo = n_lo >> 4, ns_hi = n_hi >> 4;
  const uint32_t L = ((ns_lo & ~(n_lo & ne_lo)) | (ns_hi & ~(n_hi & ne_hi))) & 0x08880888u;
  const uint32_t R = ((n_lo & ~(ns_lo & ne_lo)) | (n_hi & ~(ns_hi & ne_hi))) & 0x08880888u;
  // vertical pairs (r, r+1) sit at row r, r < 3
  const uint32_t v_lo = __funnelshift_r(lo, hi, 16), v_hi = hi >> 16;
  const uint32_t nev_lo = nz3(lo ^ v_lo), nev_hi = nz3(hi ^ v_hi);
  const uint32_t nv_lo = __funnelshift_r(n_lo, n_hi, 16), nv_hi = n_hi >> 16;
  const uint32_t U = (nv_lo & ~(n_lo & nev_lo)) | (nv_hi & ~(n_hi & nev_hi));
  const uint32_t D = (n_lo & ~(nv_lo & nev_lo)) | ((n_hi & ~(nv_hi & nev_hi)) & 0x00008888u);
  return (U ? 1u : 0u) | (D ? 2u : 0u) | (L ? 4u : 0u) | (R ? 8u : 0u);
}

// ---- direction handling ----------------------------------------------------------------------
// Every move is reduced to "slide rows left" by a reversible transform P_a:
//   up    : transpose                  down : transpose o vertical-flip
//   left  : identity                   right: horizontal-flip
// P_a = (masked nibble delta-swap) o (byte permutation).  The byte permutation is one PRMT per
// 32-bit half with an action-indexed selector; the delta-swap
//     t = (v ^ (v >> s)) & m;  v ^= t ^ (t << s)
// is the nibble part of the 4x4 transpose (s = 12, m = 0x0000F0F0) or the nibble swap inside each
// byte of a horizontal flip (s = 4, m = 0x0F0F0F0F); for `left` m = 0.  Both halves are
// involutions, so the inverse is delta-swap first, then the inverse byte permutation.
struct ActXform {
  uint32_t sel_fwd;  // lo selector | hi selector << 16
  uint32_t sel_inv;
  uint32_t shift;
  uint32_t mask;
};

__host__ __device__ constexpr ActXform act_xform(int a) {
  return a == 0   ? ActXform{0x6240u | (0x7351u << 16), 0x6240u | (0x7351u << 16), 12u, 0x0000F0F0u}
         : a == 1 ? ActXform{0x0426u | (0x1537u << 16), 0x5173u | (0x4062u << 16), 12u, 0x0000F0F0u}
         : a == 2 ? ActXform{0x3210u | (0x7654u << 16), 0x3210u | (0x7654u << 16), 0u, 0u}
                  : ActXform{0x2301u | (0x6745u << 16), 0x2301u | (0x6745u << 16), 4u, 0x0F0F0F0Fu};
}

__device__ __forceinline__ uint32_t delta_swap(uint32_t v, uint32_t s, uint32_t m) {
  const uint32_t t = (v ^ (v >> s)) & m;
  return v ^ t ^ (t << s);
}

// Look one transformed row up.  `slut` may be the shared-memory copy (first LUT_SMEM_ROWS rows)
// or NULL, in which case every lookup goes to the global table.
template <bool SMEM>
__device__ __forceinline__ uint32_t row_lookup(uint32_t idx, const uint32_t* slut,
                                               const uint32_t* __restrict__ glut) {
  if (SMEM) {
    if (idx < (uint32_t)LUT_SMEM_ROWS) return slut[idx];
    return __ldg(glut + idx);
  } else {
    return __ldg(glut + idx);
  }
}

// Slide + merge one board by one action.  Outputs the slid board (no spawn), merge reward and
// overflow flag (0 or non-zero).
template <bool SMEM>
__device__ __forceinline__ void slide_board(uint32_t lo, uint32_t hi, const ActXform x,
                                            const uint32_t* slut, const uint32_t* __restrict__ glut,
                                            uint32_t& olo, uint32_t& ohi, uint32_t& reward,
                                            uint32_t& overflow) {
  uint32_t zl = __byte_perm(lo, hi, x.sel_fwd);
  uint32_t zh = __byte_perm(lo, hi, x.sel_fwd >> 16);
  zl = delta_swap(zl, x.shift, x.mask);
  zh = delta_swap(zh, x.shift, x.mask);
  const uint32_t e0 = row_lookup<SMEM>(zl & 0xFFFFu, slut, glut);
  const uint32_t e1 = row_lookup<SMEM>(zl >> 16, slut, glut);
  const uint32_t e2 = row_lookup<SMEM>(zh & 0xFFFFu, slut, glut);
  const uint32_t e3 = row_lookup<SMEM>(zh >> 16, slut, glut);
  uint32_t wl = __byte_perm(e0, e1, 0x5410);
  uint32_t wh = __byte_perm(e2, e3, 0x5410);
  const uint32_t h01 = __byte_perm(e0, e1, 0x7632);
  const uint32_t h23 = __byte_perm(e2, e3, 0x7632);
  overflow = (h01 | h23) & 0x80008000u;
  const uint32_t s = (h01 & 0x7FFF7FFFu) + (h23 & 0x7FFF7FFFu);  // two 16-bit lanes, no carry
  reward = ((s & 0xFFFFu) + (s >> 16)) << 2;
  wl = delta_swap(wl, x.shift, x.mask);
  wh = delta_swap(wh, x.shift, x.mask);
  olo = __byte_perm(wl, wh, x.sel_inv);
  ohi = __byte_perm(wl, wh, x.sel_inv >> 16);
}

// ---- spawn ---------------------------------------------------------------------------------------
// Put exponent `e` into the k-th empty cell (row-major) where k = floor(w_pos * n_empty / 2^32).
// Requires at least one empty cell and at most 15 (true after any board-changing move).
__device__ __forceinline__ void spawn_kth_empty(uint32_t& lo, uint32_t& hi, uint32_t w_pos,
                                                uint32_t e) {
  const uint32_t e_lo = (~nz3(lo) & 0x88888888u) >> 3;  // bit 0 of every empty nibble
  const uint32_t e_hi = (~nz3(hi) & 0x88888888u) >> 3;
  // inclusive prefix counts per nibble: multiply by 0x11111111 (counts <= 15 never carry)
  const uint32_t p_lo = e_lo * 0x11111111u;
  const uint32_t c_lo = p_lo >> 28;
  const uint32_t p_hi = (e_hi + c_lo) * 0x11111111u;
  const uint32_t cnt = p_hi >> 28;
  const uint32_t k1 = __umulhi(w_pos, cnt) + 1u;  // 1-based rank of the chosen empty cell
  const uint32_t tgt = k1 * 0x11111111u;
  // the chosen nibble is the empty one whose prefix count equals k1
  const uint32_t h_lo = ~nz3(p_lo ^ tgt) & (e_lo << 3);
  const uint32_t h_hi = ~nz3(p_hi ^ tgt) & (e_hi << 3);
  // exactly one bit (bit 3 of the chosen nibble) is set across h_lo/h_hi
  lo |= (h_lo >> 3) * e;
  hi |= (h_hi >> 3) * e;
}

// Insert exponent e at cell (0..15); returns false if the cell is occupied.
__device__ __forceinline__ bool spawn_at(uint32_t& lo, uint32_t& hi, uint32_t cell, uint32_t e) {
  const uint32_t sh = (cell & 7u) * 4u;
  uint32_t& w = (cell & 8u) ? hi : lo;
  if ((w >> sh) & 0xFu) return false;
  w |= e << sh;
  return true;
}

__device__ __forceinline__ uint32_t ld_stream_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ uint4 ld_stream_v4(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ uint2 ld_stream_v2(const void* p) {
  uint2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_stream_v4(void* p, uint4 v) {
  asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}
__device__ __forceinline__ void st_stream_v2(void* p, uint2 v) {
  asm volatile("st.global.cs.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}

}  // namespace b2048
