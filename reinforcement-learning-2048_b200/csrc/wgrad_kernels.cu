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

}  // namespace
}  // namespace b2048

using namespace b2048;

extern "C" int64_t layer_wgrad_small_scratch_elems(int64_t rows, int C, int K) {
  (void)rows;
  return 160 * ((int64_t)C * K + C);   // upper bound of wg_blocks() on any supported device (<= 160 SMs)
}

extern "C" int layer_wgrad_small_f64(const double* g, const double* x, double* dw, double* db, double* scratch,
                                     int64_t rows, int C, int K, void* stream) {
  if (rows <= 0 || C <= 0 || K <= 0 || C > WG_MAX_DIM || K > WG_MAX_DIM || C * K > WG_MAX_OUT || !g || !x || !dw ||
      !db || !scratch)
    return B2048_EINVAL;
  int err = 0;
  DeviceCtx* ctx = current_ctx(&err);
  if (!ctx) return err;
  if (ctx->sm_count > 160) return B2048_EINVAL;
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
