// wgrad_kernels.cu — weight + bias gradient of a Linear / conv-as-GEMM layer whose weight matrix is tiny.
//
// In train_step's backward (src/dqn_lib.py:159-161) the first convolution (Conv2d(1,64,2): weight 64x4)
// and the output layer (Linear(64,4)) of src/configs/double_dqn_conv.py:19-28 have weight gradients
//     dW[c][k] = sum_r g[r][c] * x[r][k]        db[c] = sum_r g[r][c]
// with 256 outputs and a reduction over 45 000 / 5 000 rows.  cuBLAS runs these tall-skinny products as
// a 32x32-tile DGEMM on one or two CTAs (57 us and 8 us inside the update graph) and ATen adds a
// generic column reduction for the bias (21 us and 7 us); the work is a single pass over 23 MB.  Here
// the rows are split over the whole GPU, every block accumulates all outputs for its rows from
// shared-memory tiles, and a second small kernel adds the per-block partial sums in block order, so the
// result is bit-reproducible (no atomics).
#include "b2048_common.cuh"
#include <cstdlib>

namespace b2048 {
namespace {

constexpr int WG_THREADS = 256;
constexpr int WG_TILE = 32;        // rows per shared-memory tile
constexpr int WG_MAX_DIM = 64;     // C, K <= 64
constexpr int WG_MAX_OUT = 1024;   // C * K <= 1024 (4 outputs per thread)

// Pass 1: block b accumulates every output over its rows and writes them to partials[b][O + C].
template <int OPT>   // outputs per thread = ceil(C*K / 256)
__global__ void __launch_bounds__(WG_THREADS)
    wgrad_small_kernel(const double* __restrict__ g, const double* __restrict__ x, double* __restrict__ partials,
                       int64_t rows, int C, int K, int64_t rows_per_block) {
  __shared__ double gs[WG_TILE * WG_MAX_DIM];
  __shared__ double xs[WG_TILE * WG_MAX_DIM];
  const int tid = threadIdx.x, O = C * K;
  int oc[OPT], ok[OPT];
  double acc[OPT], accb = 0.0;
#pragma unroll
  for (int j = 0; j < OPT; ++j) {
    const int o = tid + j * WG_THREADS;
    oc[j] = o < O ? o / K : 0;
    ok[j] = o < O ? o % K : 0;
    acc[j] = 0.0;
  }
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  // software pipeline: the next tile travels from global memory into registers while the current one
  // (already in shared memory) is multiplied
  constexpr int PF = WG_TILE * WG_MAX_DIM / WG_THREADS;   // 8 elements per thread and operand
  double gr[PF], xr[PF];
  auto fetch = [&](int64_t t0) {
    const int64_t left = r1 - t0;
    const int nr = (int)(left < WG_TILE ? (left > 0 ? left : 0) : WG_TILE);
#pragma unroll
    for (int i = 0; i < PF; ++i) {
      const int e = tid + i * WG_THREADS;
      gr[i] = e < nr * C ? g[t0 * C + e] : 0.0;
      xr[i] = e < nr * K ? x[t0 * K + e] : 0.0;
    }
  };
  fetch(r0);
  for (int64_t t0 = r0; t0 < r1; t0 += WG_TILE) {
    const int nr = (int)(r1 - t0 < WG_TILE ? r1 - t0 : WG_TILE);
#pragma unroll
    for (int i = 0; i < PF; ++i) {
      const int e = tid + i * WG_THREADS;
      if (e < nr * C) gs[e] = gr[i];
      if (e < nr * K) xs[e] = xr[i];
    }
    __syncthreads();
    fetch(t0 + WG_TILE);
    for (int r = 0; r < nr; ++r) {
#pragma unroll
      for (int j = 0; j < OPT; ++j) acc[j] = fma(gs[r * C + oc[j]], xs[r * K + ok[j]], acc[j]);
      if (tid < C) accb += gs[r * C + tid];
    }
    __syncthreads();
  }
  double* mine = partials + (int64_t)blockIdx.x * (O + C);
#pragma unroll
  for (int j = 0; j < OPT; ++j)
    if (tid + j * WG_THREADS < O) mine[tid + j * WG_THREADS] = acc[j];
  if (tid < C) mine[O + tid] = accb;
}

// Pass 2: output o = sum over blocks.  Eight lanes per output each add a contiguous run of blocks in
// order, then the eight run sums are added in lane order: a fixed association, independent of timing.
__global__ void __launch_bounds__(256)
    wgrad_reduce_kernel(const double* __restrict__ partials, double* __restrict__ dw, double* __restrict__ db, int O,
                        int C, int blocks) {
  __shared__ double sm[8][32];
  const int lane_o = threadIdx.x & 31, run = threadIdx.x >> 5;
  const int o = blockIdx.x * 32 + lane_o;
  const int per = (blocks + 7) / 8;
  const int b0 = run * per, b1 = b0 + per < blocks ? b0 + per : blocks;
  double s = 0.0;
  if (o < O + C) {
#pragma unroll 8
    for (int b = b0; b < b1; ++b) s += partials[(int64_t)b * (O + C) + o];
  }
  sm[run][lane_o] = s;
  __syncthreads();
  if (run == 0 && o < O + C) {
    double t = sm[0][lane_o];
#pragma unroll
    for (int j = 1; j < 8; ++j) t += sm[j][lane_o];
    if (o < O) dw[o] = t;
    else db[o - O] = t;
  }
}

int64_t wg_blocks(int64_t rows, int sms) {
  int64_t b = (rows + 2 * WG_TILE - 1) / (2 * WG_TILE);   // at least 2 tiles per block
  if (b > sms) b = sms;
  return b < 1 ? 1 : b;
}

// ---- 64 x K weight gradient on the FP64 tensor cores (conv2 and fc1 of the conv Q-network) -----------
// dW[64][K] = g^T x over `rows` rows with K = 32 * (warps in use) <= 256: a split over rows.  Every CTA
// multiplies its row range with DMMA.8x8x4 (M = the 64 output channels, N = K, reduction = rows): warp w
// owns columns [32w, 32w+32) = 8 x 4 accumulator tiles, 12 shared-memory fragment loads per 32 DMMAs.
// Row tiles of 32 arrive by cp.async into a double buffer (row strides 68 / 260 doubles: 64-bit fragment
// loads of a half-warp then hit 16 different banks).  The bias gradient (column sums of g) rides along
// on 64 threads.  Per-CTA results go to `partials`; wgrad_reduce_kernel adds them in a fixed order.
constexpr int WT_ROWS = 32;
constexpr int WT_GS = 68;                               // doubles per g row in shared memory
constexpr int WT_XS = 260;                              // doubles per x row in shared memory
constexpr int WT_BUF = WT_ROWS * (WT_GS + WT_XS);       // doubles per buffer
constexpr int WT_SMEM_BYTES = 2 * WT_BUF * 8;

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src)
               : "memory");
}

__global__ void __launch_bounds__(256, 1)
    wgrad_dmma_kernel(const double* __restrict__ g, const double* __restrict__ x, double* __restrict__ partials,
                      int64_t rows, int K, int64_t rows_per_cta) {
  extern __shared__ __align__(16) double wsm[];
  __shared__ double bsum[3][64];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, fr = lane >> 2, fk = lane & 3;
  const int O = 64 * K;
  // CTA b owns the contiguous rows [b * rows_per_cta, ..) (a multiple of 4, the same for every CTA: the partial
  // sums do not depend on scheduling) in tiles of 32; the last tile runs only the k-steps it has rows for, so a
  // batch that is not a multiple of 32 * grid rows costs no CTA a whole extra tile (20 000 rows on 148 SMs:
  // 136 rows each instead of 4 or 5 tiles of 32)
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t r1 = r0 + rows_per_cta < rows ? r0 + rows_per_cta : rows;
  const bool active = warp * 32 < K;

  auto stage = [&](int buf, int64_t t0) {     // rows [t0, t0+32) -> buffer `buf`; rows past r1 become zeros
    double* gs = wsm + buf * WT_BUF;
    double* xs = gs + WT_ROWS * WT_GS;
    for (int i = tid; i < WT_ROWS * 32; i += 256) {          // g: 32 chunks of 2 doubles per row
      const int r = i >> 5, c2 = (i & 31) * 2;
      if (t0 + r < r1) cp_async16(gs + r * WT_GS + c2, g + (t0 + r) * 64 + c2);
      else gs[r * WT_GS + c2] = gs[r * WT_GS + c2 + 1] = 0.0;
    }
    const int xc = K / 2;                                      // chunks per x row
    for (int i = tid; i < WT_ROWS * xc; i += 256) {
      const int r = i / xc, c2 = (i - r * xc) * 2;
      if (t0 + r < r1) cp_async16(xs + r * WT_XS + c2, x + (t0 + r) * K + c2);
      else xs[r * WT_XS + c2] = xs[r * WT_XS + c2 + 1] = 0.0;
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  double acc[8][4][2];
#pragma unroll
  for (int mt = 0; mt < 8; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
  double accb = 0.0;               // bias gradient: thread = (channel tid % 64, rows 8 * (tid / 64) .. + 7 of each tile)

  int buf = 0;
  if (r0 < r1) stage(0, r0);
  for (int64_t t0 = r0; t0 < r1; t0 += WT_ROWS) {
    if (t0 + WT_ROWS < r1) {
      stage(buf ^ 1, t0 + WT_ROWS);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const double* gs = wsm + buf * WT_BUF;
    const double* xs = gs + WT_ROWS * WT_GS;
    const int ksteps = (int)(r1 - t0 < WT_ROWS ? (r1 - t0 + 3) / 4 : WT_ROWS / 4);
    if (active) {
#pragma unroll 2
      for (int ks = 0; ks < ksteps; ++ks) {
        // A[m = channel][k = row] = g[row][channel];  B[k = row][n = column] = x[row][column]
        const double* ga = gs + (ks * 4 + fk) * WT_GS + fr;
        const double* xb = xs + (ks * 4 + fk) * WT_XS + warp * 32 + fr;
        double b[4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) b[nt] = xb[nt * 8];
#pragma unroll
        for (int mt = 0; mt < 8; ++mt) {
          const double a = ga[mt * 8];
#pragma unroll
          for (int nt = 0; nt < 4; ++nt)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(acc[mt][nt][0]), "+d"(acc[mt][nt][1])
                         : "d"(a), "d"(b[nt]));
        }
      }
    }
    {
      const double* gb = gs + (tid >> 6) * 8 * WT_GS + (tid & 63);
#pragma unroll
      for (int r = 0; r < 8; ++r) accb += gb[r * WT_GS];
    }
    __syncthreads();          // everyone is done with `buf` before the next iteration refills it
    buf ^= 1;
  }
  double* mine = partials + (int64_t)blockIdx.x * (O + 64);
  if (active) {
#pragma unroll
    for (int mt = 0; mt < 8; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)   // C fragment: row = fr, columns 2*fk, 2*fk+1 of the tile
        *reinterpret_cast<double2*>(mine + (mt * 8 + fr) * K + warp * 32 + nt * 8 + 2 * fk) =
            make_double2(acc[mt][nt][0], acc[mt][nt][1]);
  }
  if (tid >= 64) bsum[(tid >> 6) - 1][tid & 63] = accb;
  __syncthreads();
  if (tid < 64) mine[O + tid] = ((accb + bsum[0][tid]) + bsum[1][tid]) + bsum[2][tid];
}

// rows per CTA (a multiple of 4) and number of CTAs of wgrad_dmma_kernel
void wt_plan(int64_t rows, int sms, int64_t* rows_per_cta, int64_t* blocks) {
  int64_t b = (rows + WT_ROWS - 1) / WT_ROWS;
  if (b > sms) b = sms;
  if (b < 1) b = 1;
  const int64_t per = ((rows + b - 1) / b + 3) / 4 * 4;
  *rows_per_cta = per;
  *blocks = (rows + per - 1) / per;
}
int64_t wt_blocks(int64_t rows, int sms) {
  int64_t per, b;
  wt_plan(rows, sms, &per, &b);
  return b;
}

// ---- the same product for K = 256, one QUARTER of the 64 x 256 result per CTA -------------------------------------
// With one full 64 x 256 partial result per CTA the 148 CTAs write 19 MB of partial sums for a 131 KB result, and the
// reduce pass reads them back (12 us for the second convolution's gradient, as long as a third of the product itself).
// Here blockIdx.y picks a quadrant (32 channels x 128 columns) and blockIdx.x a row range four times as long: the same
// work per CTA, a quarter of the partial sums (37 instead of 148 per output), four times fewer, longer main loops.
// Warp w owns columns [n0 + 16 w, + 16) x the quadrant's 32 channels = 4 x 2 accumulator tiles (6 fragment loads per
// 8 DMMAs).  Same shared-memory layout and strides as wgrad_dmma_kernel; only the quadrant's columns are staged.
__global__ void __launch_bounds__(256, 1)
    wgrad_dmma_quad_kernel(const double* __restrict__ g, const double* __restrict__ x, double* __restrict__ partials,
                           int64_t rows, int64_t rows_per_cta) {
  constexpr int K = 256, O = 64 * K;
  extern __shared__ __align__(16) double wsm[];
  __shared__ double bsum[7][32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, fr = lane >> 2, fk = lane & 3;
  const int m0 = 32 * (blockIdx.y & 1), n0 = 128 * (blockIdx.y >> 1);
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t r1 = r0 + rows_per_cta < rows ? r0 + rows_per_cta : rows;

  auto stage = [&](int buf, int64_t t0) {     // rows [t0, t0+32) -> buffer `buf`; rows past r1 become zeros
    double* gs = wsm + buf * WT_BUF;
    double* xs = gs + WT_ROWS * WT_GS;
    for (int i = tid; i < WT_ROWS * 16; i += 256) {          // g: 16 chunks of 2 doubles per row (32 channels)
      const int r = i >> 4, c2 = m0 + (i & 15) * 2;
      if (t0 + r < r1) cp_async16(gs + r * WT_GS + c2, g + (t0 + r) * 64 + c2);
      else gs[r * WT_GS + c2] = gs[r * WT_GS + c2 + 1] = 0.0;
    }
    for (int i = tid; i < WT_ROWS * 64; i += 256) {          // x: 64 chunks per row (128 columns)
      const int r = i >> 6, c2 = n0 + (i & 63) * 2;
      if (t0 + r < r1) cp_async16(xs + r * WT_XS + c2, x + (t0 + r) * K + c2);
      else xs[r * WT_XS + c2] = xs[r * WT_XS + c2 + 1] = 0.0;
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  double acc[4][2][2];
#pragma unroll
  for (int mt = 0; mt < 4; ++mt)
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
  double accb = 0.0;               // bias gradient (quadrants with n0 == 0): thread = (channel m0 + tid % 32, rows 4 * (tid / 32) .. + 3)

  int buf = 0;
  if (r0 < r1) stage(0, r0);
  for (int64_t t0 = r0; t0 < r1; t0 += WT_ROWS) {
    if (t0 + WT_ROWS < r1) {
      stage(buf ^ 1, t0 + WT_ROWS);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const double* gs = wsm + buf * WT_BUF;
    const double* xs = gs + WT_ROWS * WT_GS;
    const int ksteps = (int)(r1 - t0 < WT_ROWS ? (r1 - t0 + 3) / 4 : WT_ROWS / 4);
#pragma unroll 2
    for (int ks = 0; ks < ksteps; ++ks) {
      // A[m = channel][k = row] = g[row][channel];  B[k = row][n = column] = x[row][column]
      const double* ga = gs + (ks * 4 + fk) * WT_GS + m0 + fr;
      const double* xb = xs + (ks * 4 + fk) * WT_XS + n0 + warp * 16 + fr;
      const double b0 = xb[0], b1 = xb[8];
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) {
        const double a = ga[mt * 8];
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(acc[mt][0][0]), "+d"(acc[mt][0][1]) : "d"(a), "d"(b0));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(acc[mt][1][0]), "+d"(acc[mt][1][1]) : "d"(a), "d"(b1));
      }
    }
    if (n0 == 0) {
      const double* gb = gs + (tid >> 5) * 4 * WT_GS + m0 + (tid & 31);
#pragma unroll
      for (int r = 0; r < 4; ++r) accb += gb[r * WT_GS];
    }
    __syncthreads();          // everyone is done with `buf` before the next iteration refills it
    buf ^= 1;
  }
  double* mine = partials + (int64_t)blockIdx.x * (O + 64);
#pragma unroll
  for (int mt = 0; mt < 4; ++mt)
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)   // C fragment: row = fr, columns 2*fk, 2*fk+1 of the tile
      *reinterpret_cast<double2*>(mine + (m0 + mt * 8 + fr) * K + n0 + warp * 16 + nt * 8 + 2 * fk) =
          make_double2(acc[mt][nt][0], acc[mt][nt][1]);
  if (n0 == 0) {
    if (tid >= 32) bsum[(tid >> 5) - 1][tid & 31] = accb;
    __syncthreads();
    if (tid < 32) {
      double t = accb;
#pragma unroll
      for (int j = 0; j < 7; ++j) t += bsum[j][tid];
      mine[O + m0 + tid] = t;
    }
  }
}

// row ranges of the quartered kernel: a quarter of the SMs' worth of ranges, each a multiple of 4 rows
void wtq_plan(int64_t rows, int sms, int64_t* rows_per_cta, int64_t* chunks) {
  int64_t c = sms / 4;
  const int64_t tiles = (rows + WT_ROWS - 1) / WT_ROWS;
  if (c > tiles) c = tiles;
  if (c < 1) c = 1;
  const int64_t per = ((rows + c - 1) / c + 3) / 4 * 4;
  *rows_per_cta = per;
  *chunks = (rows + per - 1) / per;
}
// K = 256 with enough rows to give every quadrant CTA at least two tiles: the quartered kernel
bool wt_quartered(int64_t rows, int K, int sms) { return K == 256 && rows >= (int64_t)(sms / 4) * 2 * WT_ROWS; }

// ---- first convolution of the conv Q-network: col2im + ReLU mask + weight / bias gradient in one pass ----
// Backward of "conv1 -> ReLU -> im2col" given gp2 = d loss / d patches2 [4n, 256] (row = board*4 + conv2
// position, column = channel*4 + tap) and the forward patches2 (the ReLU mask: every entry that reads a
// conv1 output holds its value).  Unfused this is a col2im kernel that materialises g1 [9n, 64], a patch
// gather of the boards and the small weight-gradient kernel: 23 + 4 + 17 us and 2 x 23 MB of traffic for g1.
// Here a thread owns (board, channel): the 16 gp2 / patches2 entries of that pair are four 32-byte rows,
// the nine conv1 outputs' gradients are sums of up to four of them, and dW1[c][tap] / db1[c] accumulate
// against the board's cells in registers.  64 channels x 4 board lanes per block; per-block partial sums,
// then wgrad_reduce_kernel (fixed order).
__global__ void __launch_bounds__(256)
    conv1_wgrad_fused_kernel(const double* __restrict__ gp2, const double* __restrict__ p2,
                             const double* __restrict__ x, double* __restrict__ partials, int64_t n,
                             int64_t boards_per_block) {
  __shared__ double sm[3][5][64];
  const int ci = threadIdx.x & 63, bl = threadIdx.x >> 6;
  const int64_t b0 = (int64_t)blockIdx.x * boards_per_block;
  const int64_t b1 = b0 + boards_per_block < n ? b0 + boards_per_block : n;
  double dw[4] = {0.0, 0.0, 0.0, 0.0}, db = 0.0;
  for (int64_t b = b0 + bl; b < b1; b += 4) {
    double g[4][4], a[4][4], xb[16];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const double2* gr = reinterpret_cast<const double2*>(gp2 + (b * 4 + q) * 256 + ci * 4);
      const double2* ar = reinterpret_cast<const double2*>(p2 + (b * 4 + q) * 256 + ci * 4);
      const double2 g0 = gr[0], g1 = gr[1], a0 = ar[0], a1 = ar[1];
      g[q][0] = g0.x; g[q][1] = g0.y; g[q][2] = g1.x; g[q][3] = g1.y;
      a[q][0] = a0.x; a[q][1] = a0.y; a[q][2] = a1.x; a[q][3] = a1.y;
    }
#pragma unroll
    for (int c = 0; c < 16; ++c) xb[c] = x[b * 16 + c];
#pragma unroll
    for (int y = 0; y < 3; ++y)
#pragma unroll
      for (int xx = 0; xx < 3; ++xx) {
        // conv1 output (y, xx) is read by conv2 position (y-ky, xx-kx) through tap (ky, kx)
        double s = 0.0, act = 0.0;
        bool seen = false;
#pragma unroll
        for (int ky = 0; ky < 2; ++ky)
#pragma unroll
          for (int kx = 0; kx < 2; ++kx) {
            const int oy = y - ky, ox = xx - kx;
            if (oy < 0 || oy > 1 || ox < 0 || ox > 1) continue;
            s += g[oy * 2 + ox][ky * 2 + kx];
            if (!seen) { act = a[oy * 2 + ox][ky * 2 + kx]; seen = true; }
          }
        if (!(act > 0.0)) s = 0.0;
        db += s;
#pragma unroll
        for (int k = 0; k < 4; ++k) dw[k] = fma(s, xb[(y + (k >> 1)) * 4 + xx + (k & 1)], dw[k]);
      }
  }
  // combine the four board lanes in lane order, then one partial record per block: [64*4 dW | 64 db]
  if (bl > 0) {
#pragma unroll
    for (int k = 0; k < 4; ++k) sm[bl - 1][k][ci] = dw[k];
    sm[bl - 1][4][ci] = db;
  }
  __syncthreads();
  if (bl == 0) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
#pragma unroll
      for (int k = 0; k < 4; ++k) dw[k] += sm[j][k][ci];
      db += sm[j][4][ci];
    }
    double* mine = partials + (int64_t)blockIdx.x * 320;
#pragma unroll
    for (int k = 0; k < 4; ++k) mine[ci * 4 + k] = dw[k];
    mine[256 + ci] = db;
  }
}

int64_t c1_blocks(int64_t n, int sms) {
  int64_t b = (n + 15) / 16;          // at least 16 boards (4 per lane) per block
  if (b > 2 * (int64_t)sms) b = 2 * sms;
  return b < 1 ? 1 : b;
}

// ---- second convolution's input gradient fused with the first convolution's whole backward ------------------
// gp2 = g2 W2 ([4n,64] x [64,256], g2 = d loss / d conv2 pre-activation in (board, position) x channel rows)
// is the gradient of the patch matrix; all that is ever done with it is the pass of conv1_wgrad_fused_kernel
// above.  Materialised it is 2 x 41 MB of traffic at batch 5000 and two kernels on the critical path of the
// update (28 + 48 us).  Here the product stays in the DMMA accumulators:
//   * W2 sits in shared memory in B-fragment order for the whole launch (128 KB); every CTA stages ALL the g2
//     rows it will ever need (<= 18 row tiles of 8 = 77 KB) in the prologue, so the eight warps run without a
//     single CTA-wide barrier afterwards and drift apart: one warp's epilogue hides behind the others' DMMAs;
//   * warp w owns patch columns [32w, 32w+32) = conv1 channels 8w .. 8w+7 (x 4 taps) for every row: a lane's two
//     accumulator columns are the taps (ky, 0) and (ky, 1) of ONE channel, its row is one (board, position), so
//     the ReLU mask is one 16-byte load of patches2 and dW1[c][.] / db1[c] accumulate against six board cells:
//         dW1[c][t] += sum_tap mask * gp2[(b,pos), (c,tap)] * x[b, pos + tap + t]
//     (col2im never happens: summing over (pos, tap) pairs that hit the same conv1 output commutes with the mask,
//     which depends on that output only);
//   * per-CTA partial sums in the layout conv1_wgrad_fused_kernel writes; wgrad_reduce_kernel adds them in a
//     fixed order (bit-reproducible).
constexpr int CB_THREADS = 256;
constexpr int CB_SA = 68;                                  // doubles per staged g2 row (= 4 mod 16: conflict-free A fragments)
constexpr int CB_MAX_MT = 18;                              // row tiles (2 boards each) a CTA stages at once
constexpr int CB_W2_ELEMS = 64 * 256;
constexpr int CB_SMEM_BYTES = (CB_W2_ELEMS + CB_MAX_MT * 8 * CB_SA) * 8;

// MODE 0: the fused conv backward above.  MODE 1: the same [rows,64] x [64,256] product with a masked store instead —
// out = (g w) * (h > 0), optionally regrouped (column c * T + t of row i -> row i * T + t, column c): fc1's input
// gradient of the conv Q-network, whose reduction (64) is too short for K8's pipelined tiles to pay off
// (dgemm_dmma_kernel: 23 us, a fifth of it DMMA).
struct CbArgs {
  const double* g;        // [rows, 64]
  const double* w;        // [64, 256]
  const double* h;        // [rows, 256]: patches2 (MODE 0) / the layer below's output (MODE 1)
  const double* x;        // MODE 0: states [rows / 4, 16]
  double* out;            // MODE 0: per-CTA partial records [grid][320];  MODE 1: the product
  int64_t rows;
  int chunk_mt;
  int64_t nchunks;
  int group;              // MODE 1: 0 = plain [rows, 256], T = regrouped [rows * T, 256 / T]
  int stagger;            // cycles by which warps 4..7 start behind warps 0..3 (0 = together)
};

template <int MODE>
__global__ void __launch_bounds__(CB_THREADS, 1) cb_kernel(const CbArgs p) {
  extern __shared__ __align__(16) double cbsm[];
  double* w2f = cbsm;                      // [16 k-steps][32 n-tiles][32 lanes]
  double* g2s = cbsm + CB_W2_ELEMS;        // [chunk_mt * 8 rows][CB_SA]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, fr = lane >> 2, fk = lane & 3;
  const int64_t rows = p.rows;
  const int chunk_mt = p.chunk_mt;

  // B fragment of mma.m8n8k4 (col): lane holds B[k = lane % 4][n = lane / 4]; k-step ks covers reduction indices
  // 4 ks .. 4 ks + 3, n-tile nt output columns 8 nt .. 8 nt + 7:  w2f[(ks*32 + nt)*32 + lane] = W[4 ks + lane%4][8 nt + lane/4]
  for (int u = warp; u < 16 * 32; u += CB_THREADS / 32) {
    const int ks = u >> 5, nt = u & 31;
    const double* src = p.w + (ks * 4 + fk) * 256 + nt * 8 + fr;
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(w2f + u * 32 + lane);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
  }

  // MODE 0 — this lane's place in the fragments: row = (board, conv2 position pos), columns = taps (ky, 0), (ky, 1) of a channel
  const int pos = fr & 3, ky = fk & 1;
  const int cell0 = ((pos >> 1) + ky) * 4 + (pos & 1);      // board cell under tap (ky, 0) of position pos
  double dw[4][4], db[4];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    db[nt] = 0.0;
#pragma unroll
    for (int t = 0; t < 4; ++t) dw[nt][t] = 0.0;
  }

  for (int64_t chunk = blockIdx.x; chunk < p.nchunks; chunk += gridDim.x) {
    const int64_t row0 = chunk * chunk_mt * 8;
    if (chunk != (int64_t)blockIdx.x) __syncthreads();      // everyone is done with the previous chunk's rows
    for (int i = tid; i < chunk_mt * 8 * 16; i += CB_THREADS) {   // 32-byte units, 16 per row; rows past the batch are zero-filled
      const int r = i >> 4, c = (i & 15) * 4;
      const bool ok = row0 + r < rows;
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(g2s + r * CB_SA + c);
      const double* src = ok ? p.g + (row0 + r) * 64 + c : p.g;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(ok ? 16 : 0) : "memory");
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + 16), "l"(src + 2), "r"(ok ? 16 : 0) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    // Each scheduler hosts warps w and w + 4.  Left alone they run load issue, products and epilogue in lockstep;
    // starting the upper four half a product loop later lets one warp's DMMA stream cover the other's epilogue.
    if (p.stagger > 0 && warp >= 4 && chunk == (int64_t)blockIdx.x) {
      const long long t0 = clock64();
      while (clock64() - t0 < p.stagger) {}
    }

    for (int mt0 = 0; mt0 < chunk_mt; mt0 += 2) {
      // ---- what the epilogue needs, requested before the products: ReLU masks (the forward values) and board cells
      double2 act[2][4];
      double xc[2][6];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const int64_t row = row0 + (mt0 + mt) * 8 + fr;
        const bool ok = mt0 + mt < chunk_mt && row < rows;
        const double* pr = p.h + row * 256 + warp * 32 + 2 * fk;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) act[mt][nt] = ok ? __ldg(reinterpret_cast<const double2*>(pr + nt * 8)) : make_double2(0.0, 0.0);
        if (MODE == 0) {
          const double* xr = p.x + (row >> 2) * 16 + cell0;
#pragma unroll
          for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) xc[mt][dy * 3 + dx] = ok ? __ldg(xr + dy * 4 + dx) : 0.0;
        }
      }
      // ---- product tile: 16 rows x 32 columns per warp, reduction over 64 -------------------------------------
      double acc[2][4][2];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
      const double* ap = g2s + (mt0 * 8 + fr) * CB_SA + fk;
      const double* bp = w2f + (warp * 4) * 32 + lane;
      // fragments of k-step ks + 1 are requested before the products of k-step ks are issued: a warp sits on its DMMAs
      // (16 cycles of the pipe each), so loads issued after them would start a full k-step late
      double a[2][2], b[2][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) b[0][nt] = bp[nt * 32];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) a[0][mt] = ap[mt * 8 * CB_SA];
#pragma unroll
      for (int ks = 0; ks < 16; ++ks) {
        const int cur = ks & 1, nxt = cur ^ 1;
        if (ks + 1 < 16) {
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) b[nxt][nt] = bp[((ks + 1) * 32 + nt) * 32];
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) a[nxt][mt] = ap[mt * 8 * CB_SA + (ks + 1) * 4];
        }
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int nt = 0; nt < 4; ++nt)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(acc[mt][nt][0]), "+d"(acc[mt][nt][1])
                         : "d"(a[cur][mt]), "d"(b[cur][nt]));
      }
      // (the empty asm pins the compares behind the products: ptxas otherwise hoists them — and with them the wait
      // for the mask loads — in front of the DMMA loop to save registers)
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) asm volatile("" : "+d"(act[mt][nt].x), "+d"(act[mt][nt].y));
      if (MODE == 0) {
        // ---- mask, then accumulate against the cells: tap (ky, e) of weight tap t = (ty, tx) reads cell (ty, e + tx)
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {
            const double s0 = act[mt][nt].x > 0.0 ? acc[mt][nt][0] : 0.0;
            const double s1 = act[mt][nt].y > 0.0 ? acc[mt][nt][1] : 0.0;
            db[nt] += s0 + s1;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const int o = (t >> 1) * 3 + (t & 1);
              dw[nt][t] = fma(s1, xc[mt][o + 1], fma(s0, xc[mt][o], dw[nt][t]));
            }
          }
      } else {
        // ---- masked store; regrouped: columns j, j + 1 (j even, T even) share c = j / T and land in rows i*T + t, + 1
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          const int64_t row = row0 + (mt0 + mt) * 8 + fr;
          if (mt0 + mt >= chunk_mt || row >= rows) continue;
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {
            const int j = warp * 32 + nt * 8 + 2 * fk;
            const double v0 = act[mt][nt].x > 0.0 ? acc[mt][nt][0] : 0.0;
            const double v1 = act[mt][nt].y > 0.0 ? acc[mt][nt][1] : 0.0;
            if (p.group) {
              const int wd = 256 / p.group;
              double* o = p.out + (row * p.group + j % p.group) * wd + j / p.group;
              o[0] = v0;
              o[wd] = v1;
            } else {
              *reinterpret_cast<double2*>(p.out + row * 256 + j) = make_double2(v0, v1);
            }
          }
        }
      }
    }
  }
  if (MODE != 0) return;

  // ---- per-CTA record [64 channels x 4 taps | 64 bias]: sum over the 8 rows x 2 tap rows that share a channel ----
  double* mine = p.out + (int64_t)blockIdx.x * 320;
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    double v[5] = {dw[nt][0], dw[nt][1], dw[nt][2], dw[nt][3], db[nt]};
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      v[j] += __shfl_xor_sync(0xFFFFFFFFu, v[j], 1);
      v[j] += __shfl_xor_sync(0xFFFFFFFFu, v[j], 4);
      v[j] += __shfl_xor_sync(0xFFFFFFFFu, v[j], 8);
      v[j] += __shfl_xor_sync(0xFFFFFFFFu, v[j], 16);
    }
    if (fr == 0 && (fk & 1) == 0) {
      const int c = warp * 8 + nt * 2 + (fk >> 1);
#pragma unroll
      for (int t = 0; t < 4; ++t) mine[c * 4 + t] = v[t];
      mine[256 + c] = v[4];
    }
  }
}

int cb_stagger() {
  static const int v = [] {
    const char* e = getenv("B2048_CB_STAGGER");
    return e ? atoi(e) : 2000;
  }();
  return v;
}

// chunking of cb_kernel: row tiles per chunk (even, <= CB_MAX_MT) and number of chunks, so that one round
// of chunks covers the batch when it fits (5 000 boards: 2 500 row tiles = 139 chunks of 18)
void cb_plan(int64_t rows, int sms, int* chunk_mt, int64_t* nchunks) {
  const int64_t mtiles = (rows + 7) / 8;
  const int64_t rounds = (mtiles + (int64_t)sms * CB_MAX_MT - 1) / ((int64_t)sms * CB_MAX_MT);
  int64_t per = (mtiles + sms * rounds - 1) / (sms * rounds);
  per = (per + 1) / 2 * 2;
  if (per > CB_MAX_MT) per = CB_MAX_MT;
  if (per < 2) per = 2;
  *chunk_mt = (int)per;
  *nchunks = (mtiles + per - 1) / per;
}

}  // namespace
}  // namespace b2048

using namespace b2048;

extern "C" int64_t layer_wgrad_small_scratch_elems(int64_t rows, int C, int K) {
  int err = 0;
  DeviceCtx* ctx = current_ctx(&err);          // one partial result per block of the current device's launch
  if (!ctx || rows <= 0 || C <= 0 || K <= 0) return 0;
  return wg_blocks(rows, ctx->sm_count) * ((int64_t)C * K + C);
}

extern "C" int layer_wgrad_small_f64(const double* g, const double* x, double* dw, double* db, double* scratch,
                                     int64_t rows, int C, int K, void* stream) {
  if (rows <= 0 || C <= 0 || K <= 0 || C > WG_MAX_DIM || K > WG_MAX_DIM || C * K > WG_MAX_OUT || !g || !x || !dw ||
      !db || !scratch)
    return B2048_EINVAL;
  int err = 0;
  DeviceCtx* ctx = current_ctx(&err);
  if (!ctx) return err;
  const int64_t blocks = wg_blocks(rows, ctx->sm_count);
  const int64_t rpb = (rows + blocks - 1) / blocks;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int O = C * K, opt = (O + WG_THREADS - 1) / WG_THREADS;
  if (opt == 1) wgrad_small_kernel<1><<<(unsigned)blocks, WG_THREADS, 0, st>>>(g, x, scratch, rows, C, K, rpb);
  else if (opt == 2) wgrad_small_kernel<2><<<(unsigned)blocks, WG_THREADS, 0, st>>>(g, x, scratch, rows, C, K, rpb);
  else wgrad_small_kernel<4><<<(unsigned)blocks, WG_THREADS, 0, st>>>(g, x, scratch, rows, C, K, rpb);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  wgrad_reduce_kernel<<<(O + C + 31) / 32, 256, 0, st>>>(scratch, dw, db, O, C, (int)blocks);
  return (int)cudaGetLastError();
}

extern "C" int64_t layer_wgrad64_scratch_elems(int64_t rows, int K) {
  int err = 0;
  DeviceCtx* ctx = current_ctx(&err);
  if (!ctx || rows <= 0 || K <= 0) return 0;
  if (wt_quartered(rows, K, ctx->sm_count)) {
    int64_t per, chunks;
    wtq_plan(rows, ctx->sm_count, &per, &chunks);
    return chunks * (64 * (int64_t)K + 64);
  }
  return wt_blocks(rows, ctx->sm_count) * (64 * (int64_t)K + 64);
}

extern "C" int layer_wgrad64_f64(const double* g, const double* x, double* dw, double* db, double* scratch,
                                 int64_t rows, int K, void* stream) {
  if (rows <= 0 || K <= 0 || K > 256 || (K & 31) || !g || !x || !dw || !db || !scratch ||
      ((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(scratch)) & 15u))
    return B2048_EINVAL;
  int err = 0;
  DeviceCtx* ctx = current_ctx(&err);
  if (!ctx) return err;
  int64_t blocks, rows_per_cta;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (wt_quartered(rows, K, ctx->sm_count)) {
    wtq_plan(rows, ctx->sm_count, &rows_per_cta, &blocks);
    wgrad_dmma_quad_kernel<<<dim3((unsigned)blocks, 4), 256, WT_SMEM_BYTES, st>>>(g, x, scratch, rows, rows_per_cta);
  } else {
    wt_plan(rows, ctx->sm_count, &rows_per_cta, &blocks);
    wgrad_dmma_kernel<<<(unsigned)blocks, 256, WT_SMEM_BYTES, st>>>(g, x, scratch, rows, K, rows_per_cta);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  const int O = 64 * K;
  wgrad_reduce_kernel<<<(O + 64 + 31) / 32, 256, 0, st>>>(scratch, dw, db, O, 64, (int)blocks);
  return (int)cudaGetLastError();
}

namespace b2048 {
cudaError_t wgrad_kernels_configure() {
  cudaError_t e = cudaFuncSetAttribute(wgrad_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WT_SMEM_BYTES);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(wgrad_dmma_quad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WT_SMEM_BYTES);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(cb_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, CB_SMEM_BYTES);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(cb_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, CB_SMEM_BYTES);
}
}  // namespace b2048

extern "C" int64_t conv1_wgrad_fused_scratch_elems(int64_t n) {
  int err = 0;
  DeviceCtx* ctx = current_ctx(&err);
  if (!ctx || n <= 0) return 0;
  return c1_blocks(n, ctx->sm_count) * 320;
}

extern "C" int conv1_wgrad_fused_f64(const double* gpatches2, const double* patches2, const double* states, double* dw1,
                                     double* db1, double* scratch, int64_t n, void* stream) {
  if (n <= 0 || !gpatches2 || !patches2 || !states || !dw1 || !db1 || !scratch ||
      ((reinterpret_cast<uintptr_t>(gpatches2) | reinterpret_cast<uintptr_t>(patches2)) & 15u))
    return B2048_EINVAL;
  int err = 0;
  DeviceCtx* ctx = current_ctx(&err);
  if (!ctx) return err;
  const int64_t blocks = c1_blocks(n, ctx->sm_count);
  const int64_t bpb = ((n + blocks - 1) / blocks + 3) / 4 * 4;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  conv1_wgrad_fused_kernel<<<(unsigned)blocks, 256, 0, st>>>(gpatches2, patches2, states, scratch, n, bpb);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  wgrad_reduce_kernel<<<(320 + 31) / 32, 256, 0, st>>>(scratch, dw1, db1, 256, 64, (int)blocks);
  return (int)cudaGetLastError();
}

extern "C" int64_t conv2_dgrad_conv1_wgrad_scratch_elems(int64_t n) {
  int err = 0;
  DeviceCtx* ctx = current_ctx(&err);
  if (!ctx || n <= 0) return 0;
  int chunk_mt;
  int64_t nchunks;
  cb_plan(4 * n, ctx->sm_count, &chunk_mt, &nchunks);
  return (nchunks < ctx->sm_count ? nchunks : (int64_t)ctx->sm_count) * 320;
}

extern "C" int conv2_dgrad_conv1_wgrad_f64(const double* g2, const double* w2, const double* patches2, const double* states,
                                           double* dw1, double* db1, double* scratch, int64_t n, void* stream) {
  if (n <= 0 || !g2 || !w2 || !patches2 || !states || !dw1 || !db1 || !scratch ||
      ((reinterpret_cast<uintptr_t>(g2) | reinterpret_cast<uintptr_t>(patches2)) & 15u))
    return B2048_EINVAL;
  int err = 0;
  DeviceCtx* ctx = current_ctx(&err);
  if (!ctx) return err;
  int chunk_mt;
  int64_t nchunks;
  cb_plan(4 * n, ctx->sm_count, &chunk_mt, &nchunks);
  const int64_t blocks = nchunks < ctx->sm_count ? nchunks : (int64_t)ctx->sm_count;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const CbArgs p{g2, w2, patches2, states, scratch, 4 * n, chunk_mt, nchunks, 0, cb_stagger()};
  cb_kernel<0><<<(unsigned)blocks, CB_THREADS, CB_SMEM_BYTES, st>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  wgrad_reduce_kernel<<<(320 + 31) / 32, 256, 0, st>>>(scratch, dw1, db1, 256, 64, (int)blocks);
  return (int)cudaGetLastError();
}

// out = (g[rows,64] w[64,256]) * (h[rows,256] > 0), regrouped when group > 0 (see dense_linear_dgrad_regroup_f64):
// the short-reduction input gradient on the weights-in-shared-memory kernel (MODE 1 of cb_kernel).
namespace b2048 {
int dgrad64x256_masked(const double* g, const double* w, const double* h, double* out, int64_t rows, int group,
                       cudaStream_t st) {
  int err = 0;
  DeviceCtx* ctx = current_ctx(&err);
  if (!ctx) return err;
  int chunk_mt;
  int64_t nchunks;
  cb_plan(rows, ctx->sm_count, &chunk_mt, &nchunks);
  const int64_t blocks = nchunks < ctx->sm_count ? nchunks : (int64_t)ctx->sm_count;
  const CbArgs p{g, w, h, nullptr, out, rows, chunk_mt, nchunks, group, cb_stagger()};
  cb_kernel<1><<<(unsigned)blocks, CB_THREADS, CB_SMEM_BYTES, st>>>(p);
  return (int)cudaGetLastError();
}
}  // namespace b2048
