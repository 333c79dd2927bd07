// qnet_kernels.cu — K6: the forward pass of the reference's convolutional Q-network as ONE kernel.
//
// Replaces `model(state)` for the conv config (src/configs/double_dqn_conv.py:19-28:
// Conv2d(1,64,2) ReLU Conv2d(64,64,2) ReLU Flatten Linear(256,64) ReLU Linear(64,4), float64): action
// selection in epsilon_greedy_policy (src/dqn_lib.py:24-25), greedy play (src/player.py:47), the two
// target-side forwards of train_step (src/dqn_lib.py:126-128) and — in the SAVE instantiation, which
// also stores the activations a backward pass needs — Q(s) of train_step itself (:146).
// Input is either packed boards (exponents as in board.log_scale(), or tiles / max tile as in
// board.normalized()) or the float64 [n,16] states the replay ring emits; output is Q[n,4].
//
// Per sample the network is 84 480 multiply-adds, 97 % of them in two GEMM-shaped layers:
//   conv2: [4 positions x 256] x [256 x 64]      fc1: [1 x 256] x [256 x 64]
// Both run on the FP64 tensor cores (DMMA.8x8x4 via mma.sync.m8n8k4.f64; sm_100a has no tcgen05 path
// for float64).  One CTA per SM keeps the conv2 weights (128 KB) in shared memory in B-fragment order
// for the whole launch; a warp owns 4 samples = 16 conv2 rows x 64 columns = 16 accumulator tiles.
// conv1 (4 multiply-adds per element) is recomputed on the fly as the A fragment of conv2: a thread's
// fragment element always belongs to the same (sample, conv2 position, kernel tap), so its four board
// cells stay in registers for the whole K loop and no im2col buffer exists anywhere.  Two warps then
// pool their 8 samples in shared memory for fc1 (full 8-row tiles, fc1 weights streamed from L2 through
// double-buffered register blocks); the 64 x 4 output layer is 8 FMAs per lane and action plus shuffles.
// DFMA/DADD/DSETP share the FP64 pipe with DMMA, so biases are folded into the accumulator
// initialisation and ReLU is done with integer instructions.
#include "b2048_common.cuh"
#include <cstdlib>

namespace b2048 {
namespace {

#ifndef QC_STAGGER
#define QC_STAGGER 20000                  // cycles; 0 disables
#endif
constexpr int FC1_BLK = 8;                // k-steps per register block of fc1 weights
constexpr int IN2_STRIDE = 260;           // doubles per pooled sample row: 256 + 4 (bank spread for 64-bit loads)

// shared-memory map (bytes) for a CTA of PAIRS warp pairs (4 pairs = 256 threads, or 5 = 320 threads)
template <int PAIRS>
struct QS {
  static constexpr int W2 = 0;                                   // [64 c1][8 n-tiles][32 lanes] doubles
  static constexpr int IN2 = W2 + 64 * 8 * 32 * 8;               // [PAIRS][8 samples][IN2_STRIDE]
  static constexpr int QPART = IN2 + PAIRS * 8 * IN2_STRIDE * 8; // [PAIRS][8 samples][4 actions]: second warp's partial Q
  static constexpr int W1B = QPART + PAIRS * 8 * 4 * 8;          // [64 c1][8]: w0 w1 w2 w3 bias - - -
  static constexpr int B2 = W1B + 64 * 8 * 8;                    // [64]
  static constexpr int B3 = B2 + 64 * 8;                         // [64]
  static constexpr int W4 = B3 + 64 * 8;                         // [4][64]
  static constexpr int B4 = W4 + 4 * 64 * 8;                     // [4]
  static constexpr int BYTES = B4 + 4 * 8;
};

struct QConvWeights {
  const double *w1, *b1, *w2, *b2, *w3, *b3, *w4, *b4;
};

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
// max(v, 0) with integer instructions (the FP64 pipe is shared with DMMA and is the bottleneck):
// clear everything when the sign bit is set.  -0.0 -> +0.0, NaN stays NaN.
__device__ __forceinline__ double relu(double v) {
  const int hi = __double2hiint(v), lo = __double2loint(v);
  const int keep = ~(hi >> 31);
  return __hiloint2double(hi & keep, lo & keep);
}
__device__ __forceinline__ void pair_barrier(int pair) {
  asm volatile("bar.sync %0, 64;" ::"r"(pair + 1) : "memory");
}

// network input of one cell: exponent (log_scale) or tile / max tile (normalized)
__device__ __forceinline__ double cell_value(uint64_t bd, int cell, int scaling, int emax) {
  const int e = (int)((bd >> (4 * cell)) & 15u);
  if (scaling == 0) return (double)e;
  if (e == 0) return 0.0;
  return __longlong_as_double((long long)(1023 + e - emax) << 52);   // 2^(e - emax), exact
}

// MT = conv2 row tiles per warp: 2 (4 boards per warp, 32 per CTA iteration) for long launches, 1 (2 per
// warp, 16 per iteration; fc1 then runs on half-filled row tiles) when the launch is so short that the
// finer granularity fills more SMs (n <= 16 * #SMs: one short round instead of one long one).
// PAIRS = warp pairs per CTA: 4, or 5 (320 threads, 200 registers each) for launches that fit one round
// of 40-board iterations but not one of 32 (the 5 000-board batches of the update: 125 CTAs, one round).
// SAVE = also write what the backward pass needs (train_step, src/dqn_lib.py:146-161): the conv2 input in
// im2col form (exactly the A operand this kernel builds on the fly), and the post-ReLU outputs of conv2
// (nn.Flatten order) and fc1.
struct QConvSaved {
  double* patches2;   // [4n, 256]: row (board, conv2 position), column c1*4 + tap = relu(conv1) at that tap
  double* act2;       // [n, 256]:  relu(conv2) in nn.Flatten order (channel*4 + position)
  double* act3;       // [n, 64]:   relu(fc1)
};
// One launch = up to two networks (weight sets), each with its own CTAs, each evaluating up to two batches
// ("jobs").  The unit of work is a PAIR TILE: the 4 * MT boards one warp pair carries through the network.  A
// network's jobs are laid end to end in pair tiles and dealt out to its (CTA, pair) slots round-robin with the CTA
// index running fastest — tile t goes to pair (t / ctas) % PAIRS of CTA t % ctas — so every CTA gets the same number
// of tiles (+-1) and the same mix of jobs.  A plain forward is one network with one job; the
// Double-DQN update (src/dqn_lib.py:126-128, :146) is {online: Q(s) saving, Q(s')} + {target: Q(s')} in one
// launch: 15 000 boards spread evenly over all SMs instead of three 125-CTA launches that run one after the other.
struct QJob {
  const uint64_t* boards;   // packed boards, or
  const double* states;     // float64 [n,16]
  double* q;
  QConvSaved sv;            // SAVE instantiation: patches2 == nullptr -> this job stores nothing
  int64_t n;
  int64_t pt0;              // first pair tile of the job in its network's range
};
struct QNet {
  QConvWeights w;
  int cta0, ctas;           // this network's CTAs: [cta0, cta0 + ctas)
  int job0, njobs;          // its jobs: job[job0 .. job0 + njobs)
  int64_t pts;              // pair tiles in total
};
struct QLaunch {
  QNet net[2];
  QJob job[4];
  int nnets, scaling;
  int stagger;              // cycles by which pairs 2, 3 start behind pairs 0, 1 (0 = together)
};
template <int MT, int PAIRS, bool SAVE>
__global__ void __launch_bounds__(64 * PAIRS, 1) qconv_forward_kernel(const QLaunch m) {
  extern __shared__ __align__(16) unsigned char qsm[];
  using S = QS<PAIRS>;
  constexpr int QC_THREADS = 64 * PAIRS;
  double* w2f = reinterpret_cast<double*>(qsm + S::W2);
  double* in2 = reinterpret_cast<double*>(qsm + S::IN2);
  double* qpart = reinterpret_cast<double*>(qsm + S::QPART);
  double* w1b = reinterpret_cast<double*>(qsm + S::W1B);
  double* b2s = reinterpret_cast<double*>(qsm + S::B2);
  double* b3s = reinterpret_cast<double*>(qsm + S::B3);
  double* w4s = reinterpret_cast<double*>(qsm + S::W4);
  double* b4s = reinterpret_cast<double*>(qsm + S::B4);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, pair = warp >> 1, wip = warp & 1;
  const int fr = lane >> 2, fk = lane & 3;     // fragment row (A, C) / column (B) and k index
  const int ni = (m.nnets > 1 && (int)blockIdx.x >= m.net[1].cta0) ? 1 : 0;
  const QNet& net = m.net[ni];
  const QConvWeights wts = net.w;
  const int scaling = m.scaling;

  // ---- stage the weights: conv2 in B-fragment order, the small ones as they are --------------------
  // B fragment of mma.m8n8k4 (col): lane holds B[k = lane % 4][n = lane / 4];  k-step = c1, k = tap,
  // n-tile nt covers output channels nt*8 .. nt*8+7:  w2f[(c1*8 + nt)*32 + lane] = W2[nt*8 + lane/4][c1*4 + lane%4]
  // The four taps of one (c1, channel) are 32 contiguous bytes on both sides: 16-byte cp.async copies,
  // all in flight at once (one L2 round trip instead of 64 dependent ones).
  for (int i = tid; i < 64 * 8 * 32 / 2; i += QC_THREADS) {
    const int half = i & 1, ch8 = (i >> 1) & 7, nt = (i >> 4) & 7, c1 = i >> 7;   // unit = 2 doubles
    const double* src = wts.w2 + (nt * 8 + ch8) * 256 + c1 * 4 + half * 2;
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(w2f + (c1 * 8 + nt) * 32 + ch8 * 4 + half * 2);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  for (int i = tid; i < 64 * 8; i += QC_THREADS) {
    const int c1 = i >> 3, j = i & 7;
    w1b[i] = j < 4 ? __ldg(wts.w1 + c1 * 4 + j) : (j == 4 ? __ldg(wts.b1 + c1) : 0.0);
  }
  if (tid < 64) {
    b2s[tid] = __ldg(wts.b2 + tid);
    b3s[tid] = __ldg(wts.b3 + tid);
  }
  for (int i = tid; i < 4 * 64; i += QC_THREADS) w4s[i] = __ldg(wts.w4 + i);
  if (tid < 4) b4s[tid] = __ldg(wts.b4 + tid);
  for (int i = tid; i < PAIRS * 8 * IN2_STRIDE; i += QC_THREADS) in2[i] = 0.0;   // rows a short tile never writes
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  // this thread's conv2 row is (sample, position qp) with tap fk: the conv1 output it needs sits at
  // (qy + ky, qx + kx) of the 3x3 map, and covers board cells (py + jy, px + jx)
  const int qp = fr & 3, py = (qp >> 1) + (fk >> 1), px = (qp & 1) + (fk & 1);
  const int cell0 = py * 4 + px;                       // cells cell0, +1, +4, +5
  double* in2p = in2 + pair * 8 * IN2_STRIDE;
  double* qpartp = qpart + pair * 32;

  constexpr int PT = 4 * MT;                           // boards per pair tile, 2 * MT per warp
  const int64_t pt_step = (int64_t)net.ctas * PAIRS;
#if QC_STAGGER
  // Each scheduler hosts one warp of pairs 0/1 and one of pairs 2/3.  Left alone they run conv2 (DMMA
  // bound) and epilogue + fc1 (latency bound) in lockstep; starting pairs 2/3 a third of a tile later
  // lets one warp's DMMA stream cover the other's epilogue.  Only worth it for long launches.
  if (PAIRS == 4 && pair >= 2 && m.stagger > 0) {
    const long long t0 = clock64();
    while (clock64() - t0 < m.stagger) {}
  }
#endif
  for (int64_t pt = (int64_t)pair * net.ctas + ((int)blockIdx.x - net.cta0); pt < net.pts; pt += pt_step) {
    // the job this pair tile belongs to (at most two per network) and the tile's first board in it
    const int ji = net.job0 + ((net.njobs > 1 && pt >= m.job[net.job0 + 1].pt0) ? 1 : 0);
    const QJob& job = m.job[ji];
    const uint64_t* __restrict__ boards = job.boards;
    const double* __restrict__ states = job.states;
    double* __restrict__ q = job.q;
    const QConvSaved sv = job.sv;
    const bool save = SAVE && sv.patches2 != nullptr;
    const int64_t n = job.n;
    const int64_t s_pair = (pt - job.pt0) * PT;             // first of the pair's boards
    const int64_t s_warp = s_pair + wip * (2 * MT);         // first of this warp's boards
    // ---- inputs: 4 cells for each row tile (2 boards x 4 conv2 positions per tile) -------------------
    double x[MT][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const int64_t s = s_warp + mt * 2 + (fr >> 2);
      if (s < n) {
        if (boards) {
          const uint64_t bd = boards[s];
          int emax = 0;
          if (scaling != 0) {
#pragma unroll
            for (int c = 0; c < 16; ++c) emax = max(emax, (int)((bd >> (4 * c)) & 15u));
          }
          x[mt][0] = cell_value(bd, cell0, scaling, emax);
          x[mt][1] = cell_value(bd, cell0 + 1, scaling, emax);
          x[mt][2] = cell_value(bd, cell0 + 4, scaling, emax);
          x[mt][3] = cell_value(bd, cell0 + 5, scaling, emax);
        } else {
          const double* st = states + 16 * s + cell0;
          x[mt][0] = __ldg(st); x[mt][1] = __ldg(st + 1); x[mt][2] = __ldg(st + 4); x[mt][3] = __ldg(st + 5);
        }
      } else {
        x[mt][0] = x[mt][1] = x[mt][2] = x[mt][3] = 0.0;
      }
    }

    // ---- conv1 (on the fly) + conv2: 8 * MT rows x 64 columns per warp, K = 64 channels x 4 taps ----
    // accumulators start at the bias: C fragment, lane holds C[row = lane / 4][col = 2 * (lane % 4) + {0,1}]
    double acc[MT][8][2];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        acc[mt][nt][0] = b2s[nt * 8 + 2 * fk];
        acc[mt][nt][1] = b2s[nt * 8 + 2 * fk + 1];
      }
#pragma unroll 2
    for (int c1 = 0; c1 < 64; ++c1) {
      const double4 w = *reinterpret_cast<const double4*>(w1b + c1 * 8);
      const double bias = w1b[c1 * 8 + 4];
      double a[MT];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        a[mt] = relu(fma(x[mt][3], w.w, fma(x[mt][2], w.z, fma(x[mt][1], w.y, fma(x[mt][0], w.x, bias)))));
        if (save && s_warp + mt * 2 + (fr >> 2) < n)       // patch row 4*board + position = 4*s_warp + 8*mt + fr
          sv.patches2[(4 * s_warp + 8 * mt + fr) * 256 + c1 * 4 + fk] = a[mt];
      }
      const double* bf = w2f + c1 * 256 + lane;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const double b = bf[nt * 32];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) dmma(acc[mt][nt][0], acc[mt][nt][1], a[mt], b);
      }
    }
    // first block of fc1 weights: issued now so that the L2 latency hides behind the epilogue + barrier
    const double* brow = wts.w3 + (int64_t)(wip * 32 + fr) * 256 + fk;   // B[k = fk][n = fr] of n-tile 0, k-step 0
    double wb[2][FC1_BLK][4];
#pragma unroll
    for (int ks = 0; ks < FC1_BLK; ++ks)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) wb[0][ks][nt] = __ldg(brow + nt * 8 * 256 + ks * 4);
    // ---- ReLU, pooled per warp pair in nn.Flatten order (channel * 4 + position) ---------------------
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      double* row = in2p + (wip * 2 * MT + mt * 2 + (fr >> 2)) * IN2_STRIDE + qp;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int c2 = nt * 8 + 2 * fk;
        row[c2 * 4] = relu(acc[mt][nt][0]);
        row[c2 * 4 + 4] = relu(acc[mt][nt][1]);
      }
    }
    pair_barrier(pair);
    if (save) {   // the pair's pooled rows are contiguous boards: 64 threads copy them out, coalesced
      const int t64 = wip * 32 + lane;
      for (int i = t64; i < 4 * MT * 256; i += 64) {
        const int r = i >> 8, c = i & 255;
        if (s_pair + r < n) sv.act2[(s_pair + r) * 256 + c] = in2p[r * IN2_STRIDE + c];
      }
    }

    // ---- fc1: the pair's pooled boards (8 row slots) x 32 of the 64 hidden units per warp, K = 256; the weights stream from
    // L2 through two register blocks of FC1_BLK k-steps (load block j+1 while block j multiplies) ------
    // two accumulator sets (even / odd k-steps): 8 independent DMMA chains per warp instead of 4
    double h[4][2], g[4][2];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      h[nt][0] = b3s[wip * 32 + nt * 8 + 2 * fk];
      h[nt][1] = b3s[wip * 32 + nt * 8 + 2 * fk + 1];
      g[nt][0] = g[nt][1] = 0.0;
    }
    const double* arow = in2p + fr * IN2_STRIDE + fk;
#pragma unroll
    for (int blk = 0; blk < 64 / FC1_BLK; ++blk) {
      if (blk + 1 < 64 / FC1_BLK) {
#pragma unroll
        for (int ks = 0; ks < FC1_BLK; ++ks)
#pragma unroll
          for (int nt = 0; nt < 4; ++nt)
            wb[(blk + 1) & 1][ks][nt] = __ldg(brow + nt * 8 * 256 + ((blk + 1) * FC1_BLK + ks) * 4);
      }
#pragma unroll
      for (int ks = 0; ks < FC1_BLK; ++ks) {
        const double a = arow[(blk * FC1_BLK + ks) * 4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          if (ks & 1) dmma(g[nt][0], g[nt][1], a, wb[blk & 1][ks][nt]);
          else dmma(h[nt][0], h[nt][1], a, wb[blk & 1][ks][nt]);
        }
      }
    }
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      h[nt][0] += g[nt][0];
      h[nt][1] += g[nt][1];
    }

    // ---- output layer: each lane owns 8 hidden units of one sample; 4 partial dot products, summed
    // over the 4 lanes of the sample by shuffles and over the two warps of the pair through 256 B ------
    double part[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int hh = wip * 32 + nt * 8 + 2 * fk;
      const double v0 = relu(h[nt][0]), v1 = relu(h[nt][1]);
      if (save && fr < 4 * MT && s_pair + fr < n)
        *reinterpret_cast<double2*>(sv.act3 + (s_pair + fr) * 64 + hh) = make_double2(v0, v1);
#pragma unroll
      for (int a = 0; a < 4; ++a) part[a] = fma(v1, w4s[a * 64 + hh + 1], fma(v0, w4s[a * 64 + hh], part[a]));
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      part[a] += __shfl_xor_sync(0xFFFFFFFFu, part[a], 1);
      part[a] += __shfl_xor_sync(0xFFFFFFFFu, part[a], 2);
    }
    const double mine = fk == 0 ? part[0] : fk == 1 ? part[1] : fk == 2 ? part[2] : part[3];   // action fk of sample fr
    if (wip == 1) qpartp[lane] = mine;
    pair_barrier(pair);
    if (wip == 0) {
      const int64_t s = s_pair + fr;   // pooled row fr of the pair (4 * MT valid rows)
      if (fr < 4 * MT && s < n) q[4 * s + fk] = mine + qpartp[lane] + b4s[fk];
    }
    // Reuse across tiles is ordered by the barriers themselves: a warp stores into in2 for tile t+1 only
    // after barrier 2 of tile t, which its partner reaches after its last in2 read; qpart of tile t+1 is
    // written after barrier 1 of tile t+1, which the first warp reaches after reading qpart of tile t.
  }
}

}  // namespace

template <int MT, int PAIRS, bool SAVE>
cudaError_t configure_one() {
  return cudaFuncSetAttribute(qconv_forward_kernel<MT, PAIRS, SAVE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              QS<PAIRS>::BYTES);
}

cudaError_t qnet_kernels_configure() {
  cudaError_t e;
  if ((e = configure_one<1, 4, false>()) != cudaSuccess) return e;
  if ((e = configure_one<2, 4, false>()) != cudaSuccess) return e;
  if ((e = configure_one<2, 5, false>()) != cudaSuccess) return e;
  if ((e = configure_one<2, 4, true>()) != cudaSuccess) return e;
  return configure_one<2, 5, true>();
}

}  // namespace b2048

using namespace b2048;

namespace {

// stagger (cycles) for a launch whose pairs run `rounds` tiles each
int stagger_for(int64_t rounds) {
  static const int min_rounds = [] { const char* e = getenv("B2048_QC_STAGGER_MIN"); return e ? atoi(e) : 3; }();
  static const int cycles = [] { const char* e = getenv("B2048_QC_STAGGER"); return e ? atoi(e) : QC_STAGGER; }();
  return rounds >= min_rounds ? cycles : 0;
}

// one network, one job, `ctas` CTAs
template <int MT, int PAIRS, bool SAVE>
cudaError_t launch_single(const QJob& job, const QConvWeights& w, int scaling, int64_t sms, cudaStream_t st) {
  constexpr int PT = 4 * MT;
  QLaunch m{};
  m.nnets = 1;
  m.scaling = scaling;
  m.job[0] = job;
  m.job[0].pt0 = 0;
  QNet& net = m.net[0];
  net.w = w;
  net.job0 = 0;
  net.njobs = 1;
  net.pts = (job.n + PT - 1) / PT;
  const int64_t iters = (net.pts + PAIRS - 1) / PAIRS;          // CTA iterations if one CTA did everything
  net.cta0 = 0;
  net.ctas = (int)(iters < sms ? iters : sms);
  m.stagger = stagger_for(net.pts / ((int64_t)net.ctas * PAIRS));
  qconv_forward_kernel<MT, PAIRS, SAVE><<<net.ctas, 64 * PAIRS, QS<PAIRS>::BYTES, st>>>(m);
  return cudaGetLastError();
}

}  // namespace

extern "C" int qnet_conv_forward_f64(const uint64_t* boards, const double* states, int scaling, const double* w1,
                                     const double* b1, const double* w2, const double* b2, const double* w3,
                                     const double* b3, const double* w4, const double* b4, double* q, int64_t n,
                                     void* stream) {
  int err = 0;
  DeviceCtx* ctx = current_ctx(&err);
  if (!ctx) return err;
  if (n < 0 || (n > 0 && (!q || (!boards && !states) || (boards && states))) || scaling < 0 || scaling > 1 ||
      !w1 || !b1 || !w2 || !b2 || !w3 || !b3 || !w4 || !b4)
    return B2048_EINVAL;
  if (reinterpret_cast<uintptr_t>(w2) & 15u) return B2048_EINVAL;   // staged with 16-byte cp.async
  if (n == 0) return B2048_OK;
  // Three variants: 16, 32 or 40 boards per CTA iteration.  Pick the one with the least estimated time =
  // rounds of iterations on this device x measured cost of one round on B200 (24 / 36 / ~45 us: a short
  // round still pays the full fc1 / epilogue latency, the 40-board CTA has three warps on two schedulers).
  const int64_t sms = ctx->sm_count;
  const int64_t t16 = (n + 15) / 16, t32 = (n + 31) / 32, t40 = (n + 39) / 40;
  const int64_t c16 = 24 * ((t16 + sms - 1) / sms), c32 = 36 * ((t32 + sms - 1) / sms), c40 = 45 * ((t40 + sms - 1) / sms);
  const QConvWeights w{w1, b1, w2, b2, w3, b3, w4, b4};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const QJob job{boards, states, q, QConvSaved{nullptr, nullptr, nullptr}, n, 0};
  if (c16 < c32 && c16 <= c40) return (int)launch_single<1, 4, false>(job, w, scaling, sms, st);
  if (c40 < c32) return (int)launch_single<2, 5, false>(job, w, scaling, sms, st);
  return (int)launch_single<2, 4, false>(job, w, scaling, sms, st);
}

extern "C" int qnet_conv_forward_train_f64(const double* states, const double* w1, const double* b1, const double* w2,
                                           const double* b2, const double* w3, const double* b3, const double* w4,
                                           const double* b4, double* q, double* patches2, double* act2, double* act3,
                                           int64_t n, void* stream) {
  int err = 0;
  DeviceCtx* ctx = current_ctx(&err);
  if (!ctx) return err;
  if (n <= 0 || !states || !q || !patches2 || !act2 || !act3 || !w1 || !b1 || !w2 || !b2 || !w3 || !b3 || !w4 || !b4 ||
      ((reinterpret_cast<uintptr_t>(w2) | reinterpret_cast<uintptr_t>(act3)) & 15u))
    return B2048_EINVAL;
  const int64_t sms = ctx->sm_count;
  const int64_t t32 = (n + 31) / 32, t40 = (n + 39) / 40;
  const int64_t c32 = 36 * ((t32 + sms - 1) / sms), c40 = 45 * ((t40 + sms - 1) / sms);
  const QConvWeights w{w1, b1, w2, b2, w3, b3, w4, b4};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const QJob job{nullptr, states, q, QConvSaved{patches2, act2, act3}, n, 0};
  if (c40 < c32) return (int)launch_single<2, 5, true>(job, w, 0, sms, st);
  return (int)launch_single<2, 4, true>(job, w, 0, sms, st);
}

// The forwards of one Double-DQN update in ONE launch (src/dqn_lib.py:126-128 and :146): Q(s) of the online network with
// the activations its backward pass needs, Q(s') of the online network (NULL: plain DQN) and Q(s') of the target
// network.  The SMs are divided between the two weight sets in proportion to their boards and every CTA gets the
// same number of boards (+-8): at batch 5000 that is 101 boards on each of 148 SMs instead of three launches of
// 125 CTAs x 40 boards that queue behind each other.
extern "C" int qnet_conv_forward_update_f64(const double* states, const double* next_states, const double* const* online,
                                            const double* const* target, double* q, double* patches2, double* act2,
                                            double* act3, double* q_next_online, double* q_next_target, int64_t n,
                                            void* stream) {
  int err = 0;
  DeviceCtx* ctx = current_ctx(&err);
  if (!ctx) return err;
  if (n <= 0 || !states || !next_states || !online || !target || !q || !patches2 || !act2 || !act3 || !q_next_target)
    return B2048_EINVAL;
  for (int i = 0; i < 8; ++i)
    if (!online[i] || !target[i]) return B2048_EINVAL;
  if ((reinterpret_cast<uintptr_t>(online[2]) | reinterpret_cast<uintptr_t>(target[2]) | reinterpret_cast<uintptr_t>(act3)) & 15u)
    return B2048_EINVAL;
  // four warp pairs (two warps per scheduler): with five, two schedulers carry three warps and bound the launch
  constexpr int MT = 2, PAIRS = 4, PT = 4 * MT;
  const int64_t sms = ctx->sm_count;
  const int64_t pts = (n + PT - 1) / PT;                       // pair tiles per batch
  QLaunch m{};
  m.nnets = 2;
  m.scaling = 0;
  const QConvSaved none{nullptr, nullptr, nullptr};
  int nj = 0;
  m.job[nj++] = QJob{nullptr, states, q, QConvSaved{patches2, act2, act3}, n, 0};
  if (q_next_online) m.job[nj++] = QJob{nullptr, next_states, q_next_online, none, n, pts};
  m.net[0].w = QConvWeights{online[0], online[1], online[2], online[3], online[4], online[5], online[6], online[7]};
  m.net[0].job0 = 0;
  m.net[0].njobs = nj;
  m.net[0].pts = pts * nj;
  m.job[nj] = QJob{nullptr, next_states, q_next_target, none, n, 0};
  m.net[1].w = QConvWeights{target[0], target[1], target[2], target[3], target[4], target[5], target[6], target[7]};
  m.net[1].job0 = nj;
  m.net[1].njobs = 1;
  m.net[1].pts = pts;
  // CTAs per network in proportion to the pair tiles, at least one each, never more than one per CTA iteration
  const int64_t total = m.net[0].pts + m.net[1].pts;
  int64_t c1 = (sms * m.net[1].pts + total / 2) / total;
  if (c1 < 1) c1 = 1;
  if (c1 > sms - 1) c1 = sms - 1;
  int64_t c0 = sms - c1;
  int cta = 0;
  int64_t want[2] = {c0, c1};
  for (int i = 0; i < 2; ++i) {
    QNet& net = m.net[i];
    const int64_t slots = (net.pts + PAIRS - 1) / PAIRS;
    const int64_t c = want[i] < slots ? want[i] : slots;
    net.cta0 = cta;
    net.ctas = (int)c;
    cta += (int)c;
  }
  m.stagger = stagger_for(total / ((int64_t)cta * PAIRS));
  qconv_forward_kernel<MT, PAIRS, true><<<cta, 64 * PAIRS, QS<PAIRS>::BYTES, static_cast<cudaStream_t>(stream)>>>(m);
  return (int)cudaGetLastError();
}
