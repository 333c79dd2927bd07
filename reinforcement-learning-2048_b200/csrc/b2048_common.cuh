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
// The shared-memory copy is stored bank-swizzled: entry i sits at i ^ ((i >> 6) & 31).  The bank of
// an entry is its low 5 index bits = cell 0 and the low bit of cell 1, and real boards make those
// very non-uniform (30 % of the synthetic benchmark's cells are empty): 6.3 wavefronts per warp
// lookup unswizzled, 3.7 swizzled, 3.5 for uniformly random banks (simulated, profiles/bank_sim.py).
// The XOR only permutes entries inside aligned groups of 32, so the staged range stays contiguous.
// DeviceCtx::lut holds the plain table [0, 65536) followed by this swizzled copy [65536, +57344).
#ifndef B2048_V_SWZ
#define B2048_V_SWZ 1
#endif
constexpr int LUT_SWZ_SHIFT = 6;
constexpr uint32_t LUT_SWZ_MASK = B2048_V_SWZ ? 31u : 0u;
__host__ __device__ constexpr uint32_t lut_swizzle(uint32_t i) { return i ^ ((i >> LUT_SWZ_SHIFT) & LUT_SWZ_MASK); }

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
  uint32_t* lut = nullptr;       // [65536] row table + [LUT_SMEM_ROWS] swizzled copy for shared memory
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

// defined in host_api.cu
DeviceCtx* current_ctx(int* err);
DeviceCtx* ctx_for(int device);

// ---- Philox4x32-10 (Salmon et al., SC'11; same constants as Random123 / cuRAND) -----------------
__host__ __device__ __forceinline__ void mulhilo32(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
#ifdef __CUDA_ARCH__
  // one IMAD.WIDE; the plain C++ 64-bit product makes nvcc add a spurious zero to the high word
  asm("{\n\t.reg .b64 p;\n\tmul.wide.u32 p, %2, %3;\n\tmov.b64 {%0, %1}, p;\n\t}" : "=r"(lo), "=r"(hi) : "r"(a), "r"(b));
#else
  const uint64_t p = (uint64_t)a * b;
  hi = (uint32_t)(p >> 32);
  lo = (uint32_t)p;
#endif
}

// Rounds: every stream of the library is Philox4x32-10 except the spawn stream of the env-step kernels, which
// is Philox4x32-7 -- the smallest round count Salmon et al. report as passing BigCrush ("Crush-resistant"),
// offered as philox4x32_7 by Random123.  The generator sits inside the ALU/issue-bound step loop, where three
// rounds less are worth 2 % of the kernel (0.317 -> 0.310 ms per 64 Mi boards).
constexpr int PHILOX_ROUNDS = 10;
constexpr int SPAWN_PHILOX_ROUNDS = 7;
template <int ROUNDS = PHILOX_ROUNDS>
__host__ __device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int i = 0; i < ROUNDS; ++i) {
    uint32_t h0, l0, h1, l1;
    mulhilo32(0xD2511F53u, c.x, h0, l0);
    mulhilo32(0xCD9E8D57u, c.z, h1, l1);
    c = make_uint4(h1 ^ c.y ^ k0, l1, h0 ^ c.w ^ k1, l0);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return c;
}

// The ten round keys of one seed, computed once on the host and passed by value as a kernel
// argument: they then sit in the constant bank and feed the round XORs directly, instead of
// being re-derived with 20 integer adds per call inside the ALU-bound loop.
struct PhiloxKeys {
  uint32_t k[20];
};
__host__ __device__ inline PhiloxKeys philox_keys(uint64_t seed, uint32_t domain) {
  PhiloxKeys pk;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32) ^ domain;
  for (int i = 0; i < 10; ++i) {
    pk.k[2 * i] = k0;
    pk.k[2 * i + 1] = k1;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return pk;
}
template <int ROUNDS = PHILOX_ROUNDS>
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, const PhiloxKeys& pk) {
#pragma unroll
  for (int i = 0; i < ROUNDS; ++i) {
    uint32_t h0, l0, h1, l1;
    mulhilo32(0xD2511F53u, c.x, h0, l0);
    mulhilo32(0xCD9E8D57u, c.z, h1, l1);
    c = make_uint4(h1 ^ c.y ^ pk.k[2 * i], l1, h0 ^ c.w ^ pk.k[2 * i + 1], l0);
  }
  return c;
}

template <int ROUNDS = PHILOX_ROUNDS>
__host__ __device__ __forceinline__ uint4 philox_at(uint64_t seed, uint32_t domain, uint64_t idx,
                                                    uint64_t step) {
  return philox4x32_10<ROUNDS>(make_uint4((uint32_t)idx, (uint32_t)(idx >> 32), (uint32_t)step,
                                  (uint32_t)(step >> 32)),
                       (uint32_t)seed, (uint32_t)(seed >> 32) ^ domain);
}

// ---- nibble SWAR ------------------------------------------------------------------------------
// The step kernel is bound by the SM's integer ALU pipe (LOP3/SHF/PRMT/ISETP/SEL: one warp
// instruction per 2 cycles per sub-partition, measured in profiles/ubench), not by HBM, so the code
// below is written to minimise ALU-pipe instructions.  Measured on B200 and NOT adopted: right
// shifts as IMAD.HI (4-cycle issue on the FMA pipe), the +0x7777.. adds as IMAD (no gain: ptxas
// already balances plain adds across both pipes), booleans as IMAD.WIDE carries.
// The "+ 0x7777.." of the carry-free nibble tests, as a functor: plain add (ptxas picks VIADD on the
// ALU pipe) or, in the streaming kernel, a multiply-add through a run-time 1 (IMAD on the FMA pipe).
struct Add7 {
  __device__ __forceinline__ uint32_t operator()(uint32_t x) const { return x + 0x77777777u; }
};
struct Add7Fma {
  uint32_t one;
  __device__ __forceinline__ uint32_t operator()(uint32_t x) const {
    uint32_t r;
    asm("mad.lo.u32 %0, %1, 0x77777777, %2;" : "=r"(r) : "r"(one), "r"(x));
    return r;
  }
};
// bit 3 of every nibble set iff the nibble is non-zero (carry-free: 7+7 < 16).
template <class A = Add7>
__device__ __forceinline__ uint32_t nz3(uint32_t v, A add = A()) {
  return (add(v & 0x77777777u) | v) & 0x88888888u;
}
// same for a ^ b, WITHOUT the final mask (bits other than bit 3 of each nibble are garbage).
// Two 3-input LOP3s, written as lop3 so that ptxas does not split the xor out (one ALU op more).
template <class A = Add7>
__device__ __forceinline__ uint32_t ne3_dirty(uint32_t a, uint32_t b, A add = A()) {
  uint32_t t, r;
  asm("lop3.b32 %0, %1, %2, 0x77777777, 0x28;" : "=r"(t) : "r"(a), "r"(b));   // (a ^ b) & 0x7777...
  t = add(t);
  asm("lop3.b32 %0, %1, %2, %3, 0xF6;" : "=r"(r) : "r"(t), "r"(a), "r"(b));   // t | (a ^ b)
  return r;
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
// The left shift is stored as a multiplier (mul_l = 2^s) so that it issues as IMAD on the FMA pipe.
// PRMT with a run-time selector whose nibbles are all < 8.  `__byte_perm` masks the selector with
// 0x7777 first (an extra LOP3 on the saturated ALU pipe); the raw instruction does not need it.
__device__ __forceinline__ uint32_t prmt_raw(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}

struct ActXform {
  uint32_t sel_fwd;  // lo selector | hi selector << 16
  uint32_t sel_inv;
  uint32_t mul_l;    // 1 << s
  uint32_t unused;   // keeps the row 32 bytes
  uint32_t mask;
  uint32_t shift;    // s
  uint32_t sel_fwd_hi, sel_inv_hi;
};

__host__ __device__ constexpr ActXform act_xform(int a) {
  return a == 0   ? ActXform{0x6240u | (0x7351u << 16), 0x6240u | (0x7351u << 16), 1u << 12, 0u, 0x0000F0F0u, 12u, 0x7351u, 0x7351u}
         : a == 1 ? ActXform{0x0426u | (0x1537u << 16), 0x5173u | (0x4062u << 16), 1u << 12, 0u, 0x0000F0F0u, 12u, 0x1537u, 0x4062u}
         : a == 2 ? ActXform{0x3210u | (0x7654u << 16), 0x3210u | (0x7654u << 16), 1u << 4, 0u, 0u, 4u, 0x7654u, 0x7654u}
                  : ActXform{0x2301u | (0x6745u << 16), 0x2301u | (0x6745u << 16), 1u << 4, 0u, 0x0F0F0F0Fu, 4u, 0x6745u, 0x6745u};
}

__device__ __forceinline__ uint32_t delta_swap(uint32_t v, const ActXform& x) {
  const uint32_t t = (v ^ (v >> x.shift)) & x.mask;
  return v ^ t ^ (t * x.mul_l);
}

// Legal-mask byte for (action a, transformed-frame mask m): m bit0 = rows can slide left (= the
// move changes the board), bit1 = right, bit2 = toward row 0, bit3 = toward row 3 of the
// TRANSFORMED board; the table maps them back to [up, down, left, right] of the real board and
// adds B2048_FLAG_DONE when nothing is legal.
__host__ __device__ constexpr uint32_t zframe_to_legal(int a, uint32_t m) {
  const uint32_t b0 = m & 1u, b1 = (m >> 1) & 1u, b2 = (m >> 2) & 1u, b3 = (m >> 3) & 1u;
  const uint32_t legal = a == 0   ? (b0 | b1 << 1 | b2 << 2 | b3 << 3)    // z = T(x): left=up, right=down, rows<->cols
                         : a == 1 ? (b1 | b0 << 1 | b2 << 2 | b3 << 3)    // z = T(flipV(x)): left=down, right=up
                         : a == 2 ? (b2 | b3 << 1 | b0 << 2 | b1 << 3)    // z = x
                                  : (b2 | b3 << 1 | b1 << 2 | b0 << 3);   // z = flipH(x): left=right
  return legal | (legal ? 0u : (uint32_t)B2048_FLAG_DONE);
}

// per-CTA constant tables in shared memory
struct SmemTabs {
  ActXform act[4];        // 128 B
  uint8_t legal[4][16];   //  64 B
};

__device__ __forceinline__ void fill_tabs(SmemTabs* t) {
  if (threadIdx.x < 4) t->act[threadIdx.x] = act_xform((int)threadIdx.x);
  if (threadIdx.x < 64) t->legal[threadIdx.x >> 4][threadIdx.x & 15] =
      (uint8_t)zframe_to_legal((int)(threadIdx.x >> 4), threadIdx.x & 15u);
}

// Row-table entry (host_api.cu builds it):
//   bits  0-15 result row            bits 16-29 merge reward / 4
//   bit  30    the row can move RIGHT (toward nibble 3)      bit 31 overflow (32768+32768)
// Row 0xEEEE (reward/4 = 0x4000) does not fit 14 bits; it is never in the shared-memory part
// (top nibble 14) and the global path adds its 65536 separately.  The shared-memory copy additionally marks every
// row with reward/4 >= 2^12 as OVERFLOW (host_api.cu: staged_entry), which sends its quad to the global path.
constexpr uint32_t ENTRY_RIGHT = 0x40000000u, ENTRY_OVF = 0x80000000u;

// Look the four transformed rows up in the full table in global memory (L2-resident, 256 KB).  Used by
// the small-batch, all-four-actions and cold paths; the streaming kernel has its own shared-memory
// lookups (env_kernels.cu: stream_board).  `extra`: 65536 for every row equal to 0xEEEE, whose reward
// does not fit the table's 14-bit field.
__device__ __forceinline__ void lookup4_global(uint32_t zl, uint32_t zh, const uint32_t* __restrict__ glut,
                                               uint32_t& e0, uint32_t& e1, uint32_t& e2, uint32_t& e3,
                                               uint32_t& extra) {
  const uint32_t i0 = zl & 0xFFFFu, i1 = zl >> 16, i2 = zh & 0xFFFFu, i3 = zh >> 16;
  e0 = __ldg(glut + i0);
  e1 = __ldg(glut + i1);
  e2 = __ldg(glut + i2);
  e3 = __ldg(glut + i3);
  extra = ((i0 == 0xEEEEu) + (i1 == 0xEEEEu) + (i2 == 0xEEEEu) + (i3 == 0xEEEEu)) * 65536u;
}

// Slide + merge one board by one action, and derive the legal mask of the INPUT board in the
// transformed frame.  Outputs: slid board (no spawn), reward, flags (legal | done | changed |
// overflow).
// tabs == nullptr: the per-action constants are computed instead of read from shared memory (cold
// paths that have no SmemTabs at hand).
template <bool LEGAL = true>
__device__ __forceinline__ void slide_board(uint32_t lo, uint32_t hi, uint32_t a, const SmemTabs* tabs,
                                            const uint32_t* __restrict__ glut,
                                            uint32_t& olo, uint32_t& ohi, uint32_t& reward,
                                            uint32_t& flags, uint32_t& changed) {
  const ActXform x = tabs ? tabs->act[a] : act_xform((int)a);
  uint32_t zl = prmt_raw(lo, hi, x.sel_fwd & 0xFFFFu);
  uint32_t zh = prmt_raw(lo, hi, x.sel_fwd_hi);
  zl = delta_swap(zl, x);
  zh = delta_swap(zh, x);
  uint32_t e0, e1, e2, e3, extra;
  lookup4_global(zl, zh, glut, e0, e1, e2, e3, extra);
  uint32_t wl = (e0 & 0xFFFFu) + (e1 << 16);
  uint32_t wh = (e2 & 0xFFFFu) + (e3 << 16);
  const uint32_t h01 = __byte_perm(e0, e1, 0x7632);
  const uint32_t h23 = __byte_perm(e2, e3, 0x7632);
  const uint32_t fl = h01 | h23;
  // two 16-bit lanes of reward/4 (each <= 2 * 0x3000), summed and scaled by 4 with one dp2a
  const uint32_t s = (h01 & 0x3FFF3FFFu) + (h23 & 0x3FFF3FFFu);
  reward = __dp2a_lo(s, 0x0404u, extra);

  // transformed-frame legality: left = changed, right = any row's RIGHT bit, and the
  // perpendicular axis by SWAR on z: pair (row r, row r+1) sits at row r.
  changed = (wl ^ zl) | (wh ^ zh);
  if (LEGAL) {
    const uint32_t n_l = nz3(zl), n_h = nz3(zh);
    const uint32_t v_l = __byte_perm(zl, zh, 0x5432), v_h = (zh >> 16);
    const uint32_t ne_l = ne3_dirty(zl, v_l), ne_h = ne3_dirty(zh, v_h);
    const uint32_t nv_l = __byte_perm(n_l, n_h, 0x5432), nv_h = (n_h >> 16);
    const uint32_t up = (nv_l & ~(n_l & ne_l)) | (nv_h & ~(n_h & ne_h));
    const uint32_t dn_l = n_l & ~(nv_l & ne_l), dn_h = n_h & ~(nv_h & ne_h);
    const uint32_t m = (changed ? 1u : 0u) | ((fl & 0x40004000u) ? 2u : 0u) | (up ? 4u : 0u) |
                       ((dn_l | (dn_h & 0x0000FFFFu)) ? 8u : 0u);
    flags = (tabs ? (uint32_t)tabs->legal[a][m] : zframe_to_legal((int)a, m)) | (changed ? (uint32_t)B2048_FLAG_CHANGED : 0u) |
            ((fl & 0x80008000u) ? (uint32_t)B2048_FLAG_OVERFLOW : 0u);
  } else {
    flags = (fl & 0x80008000u) ? (uint32_t)B2048_FLAG_OVERFLOW : 0u;
  }

  wl = delta_swap(wl, x);
  wh = delta_swap(wh, x);
  olo = prmt_raw(wl, wh, x.sel_inv & 0xFFFFu);
  ohi = prmt_raw(wl, wh, x.sel_inv_hi);
}

// ---- spawn ---------------------------------------------------------------------------------------
// Put exponent e (0 = nothing) into the k-th empty cell (row-major) where
// k = floor(w_pos * n_empty / 2^32).  Requires at most 15 empty cells.  `e29` = e << 29.
// Bit 3 of the chosen empty nibble (exactly one bit across both words; none if the board is full).
template <class A = Add7>
__device__ __forceinline__ void kth_empty_mask(uint32_t lo, uint32_t hi, uint32_t w_pos, uint32_t& h_lo,
                                               uint32_t& h_hi, A add = A()) {
  const uint32_t e3_lo = ~(add(lo & 0x77777777u) | lo) & 0x88888888u;  // bit 3 of empty nibbles
  const uint32_t e3_hi = ~(add(hi & 0x77777777u) | hi) & 0x88888888u;
  const uint32_t e_lo = (e3_lo >> 3), e_hi = (e3_hi >> 3);
  // inclusive prefix counts per nibble: multiply by 0x11111111 (counts <= 15 never carry)
  const uint32_t p_lo = e_lo * 0x11111111u;
  const uint32_t c_lo = (p_lo >> 28);
  const uint32_t p_hi = e_hi * 0x11111111u + c_lo * 0x11111111u;
  const uint32_t cnt = (p_hi >> 28);
  const uint32_t tgt = __umulhi(w_pos, cnt) * 0x11111111u + 0x11111111u;  // (k+1) in every nibble
  // the chosen nibble is the empty one whose prefix count equals k+1
  h_lo = ~ne3_dirty(p_lo, tgt, add) & e3_lo;
  h_hi = ~ne3_dirty(p_hi, tgt, add) & e3_hi;
}
template <class A = Add7>
__device__ __forceinline__ void spawn_kth_empty(uint32_t& lo, uint32_t& hi, uint32_t w_pos,
                                                uint32_t e29, A add = A()) {
  uint32_t h_lo, h_hi;
  kth_empty_mask(lo, hi, w_pos, h_lo, h_hi, add);
  // the cell is empty so + == |;  hi32(h * (e << 29)) = (h >> 3) * e: shift, scale and insert in one
  // IMAD.HI each
  lo = __umulhi(h_lo, e29) + lo;
  hi = __umulhi(h_hi, e29) + hi;
}

// ---- spawn stream v2 (B2048_ABI_VERSION 2) -----------------------------------------------------------
// Board g owns ONE 16-bit lane of a Philox4x32-7 call: call counter (g >> 3, step), lane g & 7 (lane j =
// half j & 1 of word j >> 1, low half first).  With the lane value d (0..65535) in the UPPER half of a
// 32-bit word D = d << 16 and n = number of empty cells of the slid board, one 32x32 -> 64 bit product
// D * n gives both draws: the high word k = floor(d * n / 65536) is the row-major rank of the cell, the
// low word frac = ((d * n) mod 65536) << 16 is a uniform fraction that decides the value ("4" iff
// frac < p4_threshold).  For odd n the map d -> d*n mod 65536 is a bijection, for even n it is uniform
// over multiples of n's power of two, so P(4) is exact to 2^-13 and independent of k; each cell's
// probability differs from 1/n by at most 2^-16.  One Philox call now serves eight boards: the step
// kernel's per-board cost of the generator drops from 10 to 5 instructions.
// The cell: one IMAD.WIDE per half of the board turns the bit-3 mask e3 of its empty nibbles into exclusive prefix
// counts (low word of e3 * 0x22222222: nibble i = #{empty j < i}) AND inclusive suffix counts (high word); their
// sum is the half's total in every nibble.  The cell of rank k is the empty cell whose exclusive prefix count
// equals k: x = k * 0x11111111 - prefix is zero exactly in the nibbles with that count (the counts rise
// monotonically along the board, so the borrows run through the nibbles ABOVE the target only and the nibbles
// below stay positive), and "zero nibble AND empty" is the cell -- two LOP3 per word.  (Round 1 / early round 2:
// shifted masks, inclusive prefix counts by two multiplies and a three-LOP3 equality test per word;
// tests/test_swar_model.py checks both forms against each other for every occupancy pattern.)
// DLOW: D holds the lane value d itself (low half) instead of d << 16; the product is then taken with
// cnt << 16, which the replicated count word provides for free (same 64-bit product, one LOP3 less).
template <class A = Add7, bool DLOW = false>
__device__ __forceinline__ void spawn_draw16(uint32_t& lo, uint32_t& hi, uint32_t D, uint32_t p4, uint32_t changed,
                                             A add = A()) {
  const uint32_t e3_lo = ~(add(lo & 0x77777777u) | lo) & 0x88888888u;  // bit 3 of empty nibbles
  const uint32_t e3_hi = ~(add(hi & 0x77777777u) | hi) & 0x88888888u;
  uint32_t q_lo, q_hi, r_lo, r_hi;
  mulhilo32(e3_lo, 0x22222222u, q_hi, q_lo);
  mulhilo32(e3_hi, 0x22222222u, r_hi, r_lo);
  const uint32_t ep_hi = r_lo + q_lo + q_hi;           // prefix counts of the high half incl. the low half's total
  const uint32_t cnt = (ep_hi + r_hi) & (DLOW ? 0x000F0000u : 15u);   // every nibble of the sum holds the total
  uint32_t k, frac;
  mulhilo32(D, cnt, k, frac);
  const uint32_t km = k * 0x11111111u;
  const uint32_t x_lo = km - q_lo, x_hi = km - ep_hi;
  uint32_t h_lo, h_hi;                                                 // bit 3 of the chosen nibble
  {
    const uint32_t t_lo = add(x_lo & 0x77777777u), t_hi = add(x_hi & 0x77777777u);
    asm("lop3.b32 %0, %1, %2, %3, 0x02;" : "=r"(h_lo) : "r"(t_lo), "r"(x_lo), "r"(e3_lo));   // ~(t | x) & e3
    asm("lop3.b32 %0, %1, %2, %3, 0x02;" : "=r"(h_hi) : "r"(t_hi), "r"(x_hi), "r"(e3_hi));
  }
  // insert with a variable shift: bit 3 -> bit 1 ("4") or bit 0 ("2").  (IMAD.HI with e << 29: 0.324 vs 0.318 ms)
  const uint32_t sh = (frac < p4) ? 2u : 3u;
  if (changed) {
    lo += h_lo >> sh;
    hi += h_hi >> sh;
  }
}

// lane l (0..7) of a Philox result, as D = d << 16
__device__ __forceinline__ uint32_t draw_lane(const uint4& r, uint32_t l) {
  const uint32_t j = l >> 1;
  const uint32_t w = j == 0 ? r.x : j == 1 ? r.y : j == 2 ? r.z : r.w;
  return (l & 1u) ? (w & 0xFFFF0000u) : (w << 16);
}
// the draw of global board g (one Philox call per board: small / cold paths only)
__device__ __forceinline__ uint32_t spawn_draw_of(uint64_t seed, uint64_t step, uint64_t g) {
  return draw_lane(philox_at<SPAWN_PHILOX_ROUNDS>(seed, DOM_SPAWN, g >> 3, step), (uint32_t)g & 7u);
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
// 256-bit accesses (sm_100: LDG.E.256 / STG.E.256): one full 32-byte sector per lane and instruction.
// Two 128-bit accesses per lane to the halves of one sector make every warp instruction touch 32 sectors
// for 512 B of payload, and with L1::no_allocate the second one re-fetches all of them from L2
// (profiles/ubench/stream5.cu: 0.309 -> 0.27 ms for K1's five streams with no arithmetic at all).
#ifndef B2048_V_LDMODE
#define B2048_V_LDMODE 0
#endif
#ifndef B2048_V_STMODE
#define B2048_V_STMODE 2   // st.global.L1::no_allocate (0: .cs, 1: default); A/B: 0.324 vs 0.333 ms
#endif
__device__ __forceinline__ void ld_stream_v8(const void* p, uint4& a, uint4& b) {
#if B2048_V_LDMODE == 0
  asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
#elif B2048_V_LDMODE == 1
  asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
#else
  asm volatile("ld.global.cs.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
#endif
               : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
               : "l"(p));
}
__device__ __forceinline__ void st_stream_v8(void* p, uint4 a, uint4 b) {
#if B2048_V_STMODE == 0
  asm volatile("st.global.cs.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
#elif B2048_V_STMODE == 1
  asm volatile("st.global.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
#else
  asm volatile("st.global.L1::no_allocate.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
#endif
               ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
               : "memory");
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
