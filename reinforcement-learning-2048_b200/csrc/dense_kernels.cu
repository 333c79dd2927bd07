// dense_kernels.cu — K8: float64 tensor-core GEMMs for the dense Q-network of the reference
// (src/configs/double_dqn_dense.py:7-15: Linear 16 -> 512 -> 512 -> 256 -> 4, ReLU between, float64).
//
// Replaces what torch runs as cuBLAS `cutlass_80_tensorop_d884gemm_*` (an sm_80-era kernel) + ATen bias /
// ReLU / threshold_backward / column-sum kernels inside model(states), model(next_states),
// target_model(next_states) and loss.backward() of dqn_lib.train_step (src/dqn_lib.py:125-161).
// Every matrix product of one Double-DQN update at batch 5000 is one launch of ONE kernel template:
//
//   forward   C[m][n]  = relu?( sum_k A[m][k] W[n][k] + bias[n] )          A: activations, W: nn.Linear weight
//   input grad C[m][k'] = ( sum_n G[m][n] W[n][k'] ) * (H[m][k'] > 0)        G: d loss / d (pre-activation)
//   weight grad P[s][n][k'] = sum_{m in split s} G[m][n] X[m][k']            then a fixed-order sum over s
//
// with mma.sync.m8n8k4.f64 (DMMA.8x8x4; there is no tcgen05 kind for float64).  A CTA owns an (8 MT) x
// (64 NTW) tile of the output: eight warps side by side, each holding MT x NTW accumulator tiles, every A
// fragment loaded once per warp and k-step and reused for NTW DMMAs.  Tile shapes are picked on the host so
// that one launch is one full wave of the 148 SMs (5000 rows = 37 tiles of 136; 37 x 4 column tiles = 148).
// Operand tiles arrive by cp.async, issued by a ninth (producer) warp into a four-stage ring guarded by
// mbarriers, in a layout whose row strides are = 4 (mod 16) doubles, which
// makes every 64-bit fragment load of a half-warp hit 16 different bank pairs in both orientations
// (reduction index contiguous, as the forward pass reads A and W, or strided, as the gradients do).
#include "b2048_common.cuh"

namespace b2048 {
namespace {

constexpr int DG_CONSUMERS = 256;      // eight DMMA warps ...
constexpr int DG_THREADS = 320;        // ... plus two producer warps that only issue cp.async (one per operand)
constexpr int DG_RC = 16;              // reduction elements per staged chunk
constexpr int DG_STAGES = 4;           // ring of staged chunks (42 KB each for the 136 x 128 tile)

__device__ __forceinline__ uint32_t dg_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// 16-byte async copy; `bytes` = 0 writes zeros instead (rows / reduction elements past the end of the operand)
__device__ __forceinline__ void dg_cp16(void* dst, const void* src, int bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dg_smem(dst)), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void dg_bar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(dg_smem(bar)), "r"(count));
}
__device__ __forceinline__ void dg_bar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(dg_smem(bar)), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void dg_bar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(dg_smem(bar)) : "memory");
}
// arrive on `bar` once all cp.async issued so far by this thread have landed (does not add to the expected count)
__device__ __forceinline__ void dg_bar_arrive_after_copies(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(dg_smem(bar)) : "memory");
}

__host__ __device__ constexpr int dg_pad(int w) {      // smallest stride >= w that is = 4 (mod 16)
  return w + ((4 - (w % 16)) + 16) % 16;
}

enum : int { EPI_BIAS = 0, EPI_BIAS_RELU = 1, EPI_MASK = 2, EPI_PARTIAL = 3, EPI_PLAIN = 4 };

struct DgArgs {
  const double* a;      // A operand
  const double* b;      // B operand
  double* c;            // output (or partials)
  const double* aux;    // bias [cols] (EPI_BIAS*), H [rows x ldc] (EPI_MASK), unused otherwise
  double* colsum;       // EPI_PARTIAL only, nullable: per-split sums over the reduction index of one operand (the
                        // bias gradient): of A -> [splits][rows of C] (colsum_of == 1), of B -> [splits][cols of C] (== 2)
  int colsum_of;
  int64_t lda, ldb, ldc;
  int rows, cols;       // output extent
  int red;              // reduction extent
  int red_per_split;    // reduction elements per split (multiple of DG_RC); = red rounded up when not split
  int tgroup;           // EPI_MASK only: 0 = C as it is; T > 0 = column j = c * T + t of row i goes to row i * T + t, column c
                        // of a [rows * T, cols / T] matrix (nn.Flatten order -> one row per (board, position))
};

// A_RC / B_RC: the operand is "reduction-contiguous" in global memory:
//   A_RC:  A(i, r) = a[i * lda + r]      else  A(i, r) = a[r * lda + i]
//   B_RC:  B(r, j) = b[j * ldb + r]      else  B(r, j) = b[r * ldb + j]
// Warp-specialised: warps 8 and 9 stream the two operand tiles into a DG_STAGES-deep ring with cp.async and signal each
// chunk through an mbarrier ("full"); the eight DMMA warps never touch global memory inside the loop and hand a
// chunk back through a second mbarrier ("empty").  No __syncthreads in the loop: the FP64 pipe does not drain
// while a tile is being staged.
template <int MT, int NTW, bool A_RC, bool B_RC, int EPI>
__global__ void __launch_bounds__(DG_THREADS, 1) dgemm_dmma_kernel(const DgArgs p) {
  constexpr int BM = 8 * MT, BN = 64 * NTW;
  constexpr int SA = A_RC ? dg_pad(DG_RC) : dg_pad(BM);       // doubles per smem row of the A tile
  constexpr int SB = B_RC ? dg_pad(DG_RC) : dg_pad(BN);
  constexpr int A_ELEMS = A_RC ? BM * SA : DG_RC * SA;
  constexpr int B_ELEMS = B_RC ? BN * SB : DG_RC * SB;
  constexpr int STAGE_ELEMS = A_ELEMS + B_ELEMS;
  extern __shared__ __align__(16) double dsm[];
  uint64_t* full = reinterpret_cast<uint64_t*>(dsm + DG_STAGES * STAGE_ELEMS);
  uint64_t* empty = full + DG_STAGES;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, fr = lane >> 2, fk = lane & 3;
  const int i0 = blockIdx.x * BM, j0 = blockIdx.y * BN;
  const int r_begin = blockIdx.z * p.red_per_split;
  const int r_end = min(p.red, r_begin + p.red_per_split);
  const int nchunks = r_end > r_begin ? (r_end - r_begin + DG_RC - 1) / DG_RC : 0;

  if (tid == 0) {
    for (int s = 0; s < DG_STAGES; ++s) {
      dg_bar_init(full + s, 64);            // one deferred arrival per producer lane (two producer warps)
      dg_bar_init(empty + s, 8);            // one arrival per DMMA warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp >= 8) {
    // ---- producers: warp 8 streams the A tiles, warp 9 the B tiles ---------------------------------------------------------------------------------------------
    // Every lane owns a fixed set of 16-byte slots of the tile; its source pointer and shared-memory address advance
    // by constant strides, so one slot costs a cp.async and two adds (the first version recomputed a 64-bit address
    // per slot: ~25 instructions, which made this one warp the bottleneck of the whole CTA).
    constexpr int WA = A_RC ? DG_RC : BM, WB = B_RC ? DG_RC : BN;       // doubles per staged row
    constexpr int RA = A_RC ? BM : DG_RC, RB = B_RC ? BN : DG_RC;       // staged rows
    // reduction-contiguous tiles: 8 lanes cover one row of 16 doubles, 4 rows per pass of the warp
    // reduction-strided tiles:    the warp covers 64 doubles of a row per pass, row after row
    const int a_row0 = A_RC ? (lane >> 3) : 0, a_col = A_RC ? (lane & 7) * 2 : lane * 2;
    const int b_row0 = B_RC ? (lane >> 3) : 0, b_col = B_RC ? (lane & 7) * 2 : lane * 2;
    for (int c = 0; c < nchunks; ++c) {
      const int s = c % DG_STAGES, r0 = r_begin + c * DG_RC;
      if (c >= DG_STAGES) dg_bar_wait(empty + s, ((c / DG_STAGES) - 1) & 1);
      double* as = dsm + s * STAGE_ELEMS;
      double* bs = as + A_ELEMS;
      if (warp == 9) goto stage_b;
      if (A_RC) {
        const bool col_ok = r0 + a_col < r_end;
        const double* src = p.a + (int64_t)(i0 + a_row0) * p.lda + r0 + a_col;
        double* dst = as + a_row0 * SA + a_col;
        const int64_t sstep = 4 * p.lda;
#pragma unroll 2
        for (int i = a_row0; i < RA; i += 4, src += sstep, dst += 4 * SA) {
          const bool ok = col_ok && i0 + i < p.rows;
          dg_cp16(dst, ok ? src : p.a, ok ? 16 : 0);
        }
      } else {
#pragma unroll
        for (int u = 0; u < (WA + 63) / 64; ++u) {
          const int col = a_col + 64 * u;
          if (col < WA) {
            const bool col_ok = i0 + col < p.rows;
            const double* src = p.a + (int64_t)r0 * p.lda + i0 + col;
            double* dst = as + col;
#pragma unroll 2
            for (int r = 0; r < RA; ++r, src += p.lda, dst += SA) {
              const bool ok = col_ok && r0 + r < r_end;
              dg_cp16(dst, ok ? src : p.a, ok ? 16 : 0);
            }
          }
        }
      }
      goto staged;
    stage_b:
      if (B_RC) {
        const bool col_ok = r0 + b_col < r_end;
        const double* src = p.b + (int64_t)(j0 + b_row0) * p.ldb + r0 + b_col;
        double* dst = bs + b_row0 * SB + b_col;
        const int64_t sstep = 4 * p.ldb;
#pragma unroll 2
        for (int j = b_row0; j < RB; j += 4, src += sstep, dst += 4 * SB) {
          const bool ok = col_ok && j0 + j < p.cols;
          dg_cp16(dst, ok ? src : p.b, ok ? 16 : 0);
        }
      } else {
#pragma unroll
        for (int u = 0; u < (WB + 63) / 64; ++u) {
          const int col = b_col + 64 * u;
          if (col < WB) {
            const bool col_ok = j0 + col < p.cols;
            const double* src = p.b + (int64_t)r0 * p.ldb + j0 + col;
            double* dst = bs + col;
#pragma unroll 2
            for (int r = 0; r < RB; ++r, src += p.ldb, dst += SB) {
              const bool ok = col_ok && r0 + r < r_end;
              dg_cp16(dst, ok ? src : p.b, ok ? 16 : 0);
            }
          }
        }
      }
    staged:
      dg_bar_arrive_after_copies(full + s);
    }
    return;
  }

  // ---- DMMA warps ---------------------------------------------------------------------------------------------
  double acc[MT][NTW][2];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < NTW; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
  double csum = 0.0;                 // EPI_PARTIAL: sums of one operand over this split's reduction range

  for (int c = 0; c < nchunks; ++c) {
    const int s = c % DG_STAGES;
    dg_bar_wait(full + s, (c / DG_STAGES) & 1);
    const double* as = dsm + s * STAGE_ELEMS;
    const double* bs = as + A_ELEMS;
#pragma unroll
    for (int ks = 0; ks < DG_RC / 4; ++ks) {
      double b[NTW];
#pragma unroll
      for (int nt = 0; nt < NTW; ++nt) {
        const int j = warp * (8 * NTW) + nt * 8 + fr;
        b[nt] = B_RC ? bs[j * SB + ks * 4 + fk] : bs[(ks * 4 + fk) * SB + j];
      }
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const double a = A_RC ? as[(mt * 8 + fr) * SA + ks * 4 + fk] : as[(ks * 4 + fk) * SA + mt * 8 + fr];
#pragma unroll
        for (int nt = 0; nt < NTW; ++nt)
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                       : "+d"(acc[mt][nt][0]), "+d"(acc[mt][nt][1])
                       : "d"(a), "d"(b[nt]));
      }
    }
    if (EPI == EPI_PARTIAL && !A_RC && p.colsum_of == 1 && blockIdx.y == 0 && tid < BM) {
#pragma unroll
      for (int r = 0; r < DG_RC; ++r) csum += as[r * SA + tid];
    }
    if (EPI == EPI_PARTIAL && !B_RC && p.colsum_of == 2 && blockIdx.x == 0 && tid < BN) {
#pragma unroll
      for (int r = 0; r < DG_RC; ++r) csum += bs[r * SB + tid];
    }
    __syncwarp();
    if (lane == 0) dg_bar_arrive(empty + s);      // this warp is done with the chunk
  }

  // ---- epilogue: C fragment = row (8 mt + fr), columns (.. + 2 fk, + 1) ------------------------------------
  double* cbase = p.c;
  if (EPI == EPI_PARTIAL) cbase += (int64_t)blockIdx.z * p.rows * p.ldc;
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    const int i = i0 + mt * 8 + fr;
    if (i >= p.rows) continue;
#pragma unroll
    for (int nt = 0; nt < NTW; ++nt) {
      const int j = j0 + warp * (8 * NTW) + nt * 8 + 2 * fk;
      if (j >= p.cols) continue;
      double v0 = acc[mt][nt][0], v1 = acc[mt][nt][1];
      if (EPI == EPI_BIAS || EPI == EPI_BIAS_RELU) {
        const double2 bb = *reinterpret_cast<const double2*>(p.aux + j);
        v0 += bb.x;
        v1 += bb.y;
        if (EPI == EPI_BIAS_RELU) {
          v0 = v0 > 0.0 ? v0 : 0.0;
          v1 = v1 > 0.0 ? v1 : 0.0;
        }
      } else if (EPI == EPI_MASK) {
        const double2 h = *reinterpret_cast<const double2*>(p.aux + (int64_t)i * p.ldc + j);
        v0 = h.x > 0.0 ? v0 : 0.0;
        v1 = h.y > 0.0 ? v1 : 0.0;
        if (p.tgroup) {                            // T is even and j is even: j, j + 1 share c
          const int t = j % p.tgroup, w = p.cols / p.tgroup;
          double* o = cbase + ((int64_t)i * p.tgroup + t) * w + j / p.tgroup;
          o[0] = v0;
          o[w] = v1;
          continue;
        }
      }
      *reinterpret_cast<double2*>(cbase + (int64_t)i * p.ldc + j) = make_double2(v0, v1);
    }
  }
  if (EPI == EPI_PARTIAL && !A_RC && p.colsum_of == 1 && blockIdx.y == 0 && tid < BM && i0 + tid < p.rows)
    p.colsum[(int64_t)blockIdx.z * p.rows + i0 + tid] = csum;
  if (EPI == EPI_PARTIAL && !B_RC && p.colsum_of == 2 && blockIdx.x == 0 && tid < BN && j0 + tid < p.cols)
    p.colsum[(int64_t)blockIdx.z * p.cols + j0 + tid] = csum;
}

// dw[e] = sum over splits (fixed order) of partials[s][e'] with e' = e, or the transposed index when the product
// was computed as dW^T (t_rows x t_cols = shape of the partial matrices); the per-split column sums
// (`n_sum` per split) are added into `db` the same way.
__global__ void dgemm_reduce_kernel(const double* __restrict__ partials, const double* __restrict__ colsum,
                                    double* __restrict__ dw, double* __restrict__ db, int64_t n_main, int n_sum,
                                    int splits, int transposed, int t_rows, int t_cols) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n_main) {
    int64_t src = e;
    if (transposed) {                       // dw is [t_cols][t_rows], the partials are [t_rows][t_cols]
      const int64_t j = e / t_rows, i = e - j * t_rows;
      src = i * t_cols + j;
    }
    double s = 0.0;
    for (int k = 0; k < splits; ++k) s += partials[(int64_t)k * n_main + src];
    dw[e] = s;
  } else if (e < n_main + n_sum && db) {
    const int64_t c = e - n_main;
    double s = 0.0;
    for (int k = 0; k < splits; ++k) s += colsum[(int64_t)k * n_sum + c];
    db[c] = s;
  }
}

// Last layer, 4 outputs: q[m][j] = b[j] + sum_k h[m][k] w[j][k].  One warp per row (the matrix is 4 x K).
__global__ void __launch_bounds__(256) dense_out4_kernel(const double* __restrict__ h, const double* __restrict__ w,
                                                         const double* __restrict__ bias, double* __restrict__ q,
                                                         int rows, int K) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  const double* hr = h + (int64_t)row * K;
  for (int k = lane; k < K; k += 32) {
    const double x = hr[k];
    s0 = fma(x, w[k], s0);
    s1 = fma(x, w[K + k], s1);
    s2 = fma(x, w[2 * K + k], s2);
    s3 = fma(x, w[3 * K + k], s3);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s0 += __shfl_xor_sync(0xFFFFFFFFu, s0, o);
    s1 += __shfl_xor_sync(0xFFFFFFFFu, s1, o);
    s2 += __shfl_xor_sync(0xFFFFFFFFu, s2, o);
    s3 += __shfl_xor_sync(0xFFFFFFFFu, s3, o);
  }
  if (lane == 0) {
    double2* o = reinterpret_cast<double2*>(q + 4 * (int64_t)row);
    o[0] = make_double2(s0 + bias[0], s1 + bias[1]);
    o[1] = make_double2(s2 + bias[2], s3 + bias[3]);
  }
}

// Its input gradient: dz[m][k] = (sum_j g[m][j] w[j][k]) * (h[m][k] > 0)   (g = d loss / d q, 4 columns)
__global__ void __launch_bounds__(256) dense_out4_dgrad_kernel(const double* __restrict__ g, const double* __restrict__ w,
                                                               const double* __restrict__ h, double* __restrict__ dz,
                                                               int64_t total, int K) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int64_t m = e / K;
  const int k = (int)(e - m * K);
  const double2 ga = *reinterpret_cast<const double2*>(g + 4 * m), gb = *reinterpret_cast<const double2*>(g + 4 * m + 2);
  const double v = ga.x * w[k] + ga.y * w[K + k] + gb.x * w[2 * K + k] + gb.y * w[3 * K + k];
  dz[e] = h[e] > 0.0 ? v : 0.0;
}

template <int MT, int NTW, bool A_RC, bool B_RC, int EPI>
constexpr int dg_smem_bytes() {
  constexpr int BM = 8 * MT, BN = 64 * NTW;
  constexpr int SA = A_RC ? dg_pad(DG_RC) : dg_pad(BM);
  constexpr int SB = B_RC ? dg_pad(DG_RC) : dg_pad(BN);
  return DG_STAGES * ((A_RC ? BM * SA : DG_RC * SA) + (B_RC ? BN * SB : DG_RC * SB)) * 8 + 2 * DG_STAGES * 8;
}

template <int MT, int NTW, bool A_RC, bool B_RC, int EPI>
cudaError_t dg_launch(const DgArgs& p, int splits, cudaStream_t st) {
  constexpr int BM = 8 * MT, BN = 64 * NTW;
  auto kern = dgemm_dmma_kernel<MT, NTW, A_RC, B_RC, EPI>;
  constexpr int smem = dg_smem_bytes<MT, NTW, A_RC, B_RC, EPI>();
  static bool configured = false;          // per instantiation; cudaFuncSetAttribute is idempotent
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  dim3 grid((p.rows + BM - 1) / BM, (p.cols + BN - 1) / BN, splits);
  kern<<<grid, DG_THREADS, smem, st>>>(p);
  return cudaGetLastError();
}

// tile choice: among heights 8 * {17, 9, 5} and widths 64 * {2, 1} the (MT, NTW) that wastes the fewest SM slots
// in the last wave (5000 x 512: 136 x 128 tiles = 148 CTAs; 5000 x 256: 136 x 64 tiles = 148 CTAs)
inline double dg_waste(int rows, int cols, int mt, int ntw, int sms) {
  const int64_t ctas = (int64_t)((rows + 8 * mt - 1) / (8 * mt)) * ((cols + 64 * ntw - 1) / (64 * ntw));
  const int64_t waves = (ctas + sms - 1) / sms;
  // SM-time of the launch relative to the useful work; the narrow tile pays ~6 % more per tile (one A fragment
  // load per DMMA instead of one per two)
  return (double)(waves * sms) * mt * ntw * (ntw == 1 ? 1.02 : 1.0) / ((double)rows / 8.0 * ((double)cols / 64.0));
}

template <bool A_RC, bool B_RC, int EPI>
cudaError_t dg_dispatch(const DgArgs& p, int sms, cudaStream_t st) {
  int best_mt = 17, best_ntw = 2;
  double best = 1e30;
  for (int ntw = 2; ntw >= 1; --ntw) {
    if (ntw == 2 && p.cols % 128 != 0 && p.cols > 64) continue;       // keep full column tiles where possible
    for (int mt : {17, 9, 5}) {
      const double w = dg_waste(p.rows, p.cols, mt, ntw, sms);
      if (w < best) { best = w; best_mt = mt; best_ntw = ntw; }
    }
  }
  if (best_ntw == 2) {
    if (best_mt == 17) return dg_launch<17, 2, A_RC, B_RC, EPI>(p, 1, st);
    if (best_mt == 9) return dg_launch<9, 2, A_RC, B_RC, EPI>(p, 1, st);
    return dg_launch<5, 2, A_RC, B_RC, EPI>(p, 1, st);
  }
  if (best_mt == 17) return dg_launch<17, 1, A_RC, B_RC, EPI>(p, 1, st);
  if (best_mt == 9) return dg_launch<9, 1, A_RC, B_RC, EPI>(p, 1, st);
  return dg_launch<5, 1, A_RC, B_RC, EPI>(p, 1, st);
}

}  // namespace
}  // namespace b2048

namespace b2048 {
int dgrad64x256_masked(const double* g, const double* w, const double* h, double* out, int64_t rows, int group,
                       cudaStream_t st);                                  // wgrad_kernels.cu
}
using namespace b2048;

#define DG_CTX()                         \
  int err__ = 0;                         \
  DeviceCtx* ctx = current_ctx(&err__);  \
  if (!ctx) return err__;

// C[rows x n_out] = act(A[rows x n_in] W[n_out x n_in]^T + bias)   (nn.Linear forward, src/dqn_lib.py:126-147)
extern "C" int dense_linear_forward_f64(const double* a, const double* w, const double* bias, double* c, int64_t rows,
                                        int n_in, int n_out, int relu, void* stream) {
  if (!a || !w || !bias || !c || rows <= 0 || n_in <= 0 || n_out <= 0 || rows > (1 << 30)) return B2048_EINVAL;
  if ((n_in & 1) || ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(c) |
                      reinterpret_cast<uintptr_t>(bias)) & 15u))
    return B2048_EINVAL;
  DG_CTX();
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n_out == 4) {
    if (relu) return B2048_EINVAL;
    dense_out4_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(a, w, bias, c, (int)rows, n_in);
    return (int)cudaGetLastError();
  }
  if (n_out & 1) return B2048_EINVAL;
  DgArgs p{a, w, c, bias, nullptr, 0, n_in, n_in, n_out, (int)rows, n_out, n_in, ((n_in + DG_RC - 1) / DG_RC) * DG_RC, 0};
  const cudaError_t e = relu ? dg_dispatch<true, true, EPI_BIAS_RELU>(p, ctx->sm_count, st)
                             : dg_dispatch<true, true, EPI_BIAS>(p, ctx->sm_count, st);
  return (int)e;
}

// dz_in[rows x n_in] = (G[rows x n_out] W[n_out x n_in]) * (H[rows x n_in] > 0): input gradient of a Linear layer
// fused with the ReLU mask of the layer below (H = that layer's output); h == NULL: no mask (plain G W).
static int dgrad_impl(const double* g, const double* w, const double* h, double* dz, int64_t rows, int n_in, int n_out,
                      int tgroup, void* stream) {
  if (!g || !w || !dz || rows <= 0 || n_in <= 0 || n_out <= 0 || rows > (1 << 30)) return B2048_EINVAL;
  if (tgroup && (!h || n_out == 4 || tgroup < 2 || (tgroup & 1) || n_in % tgroup)) return B2048_EINVAL;
  if ((n_in & 1) || (n_out & 1) || ((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(w) |
                                     reinterpret_cast<uintptr_t>(h) | reinterpret_cast<uintptr_t>(dz)) & 15u))
    return B2048_EINVAL;
  DG_CTX();
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n_out == 4) {
    if (!h) return B2048_EINVAL;
    const int64_t total = rows * n_in;
    dense_out4_dgrad_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(g, w, h, dz, total, n_in);
    return (int)cudaGetLastError();
  }
  // a masked 64 -> 256 layer (fc1 of the conv Q-network): reduction too short for the pipelined tiles
  if (h && n_out == 64 && n_in == 256 && (tgroup == 0 || 256 % tgroup == 0)) return dgrad64x256_masked(g, w, h, dz, rows, tgroup, st);
  // C(i = row, j = input unit) = sum_r G(i, r) W(r, j): A reduction-contiguous, B reduction-strided
  DgArgs p{g, w, dz, h, nullptr, 0, n_out, n_in, n_in, (int)rows, n_in, n_out, ((n_out + DG_RC - 1) / DG_RC) * DG_RC, tgroup};
  const cudaError_t e = h ? dg_dispatch<true, false, EPI_MASK>(p, ctx->sm_count, st)
                          : dg_dispatch<true, false, EPI_PLAIN>(p, ctx->sm_count, st);
  return (int)e;
}

extern "C" int dense_linear_dgrad_f64(const double* g, const double* w, const double* h, double* dz, int64_t rows,
                                      int n_in, int n_out, void* stream) {
  return dgrad_impl(g, w, h, dz, rows, n_in, n_out, 0, stream);
}

// The same product with the result regrouped: input unit j = c * group + t of row i is stored at row i * group + t,
// column c of dz [rows * group, n_in / group].  For a layer that follows nn.Flatten of a [C, positions] map
// (src/configs/double_dqn_conv.py:24-25) with group = positions this writes the gradient as the row matrix
// (board, position) x channel the convolution's GEMM-form backward consumes: no transpose pass.
extern "C" int dense_linear_dgrad_regroup_f64(const double* g, const double* w, const double* h, double* dz, int64_t rows,
                                              int n_in, int n_out, int group, void* stream) {
  if (group <= 0) return B2048_EINVAL;
  return dgrad_impl(g, w, h, dz, rows, n_in, n_out, group, stream);
}

// splits and scratch of the weight gradient for (rows, n_in, n_out)
static void wgrad_plan(int64_t rows, int n_in, int n_out, int sms, bool* swap, int* splits, int* per_split) {
  *swap = n_in < 64;                         // dW^T = X^T G when the layer has few inputs (16): output [n_in x n_out]
  const int out_r = *swap ? n_in : n_out, out_c = *swap ? n_out : n_in;
  const int mt = *swap ? (n_in > 32 ? 8 : n_in > 16 ? 4 : 2) : (n_out >= 128 ? 16 : n_out > 32 ? 8 : 4);
  const int tiles = ((out_r + 8 * mt - 1) / (8 * mt)) * ((out_c + 127) / 128);
  int s = sms / tiles;
  if (s < 1) s = 1;
  int per = (int)(((rows + s - 1) / s + DG_RC - 1) / DG_RC) * DG_RC;
  s = (int)((rows + per - 1) / per);
  *splits = s;
  *per_split = per;
}

extern "C" int64_t dense_linear_wgrad_scratch_elems(int64_t rows, int n_in, int n_out) {
  int err = 0;
  DeviceCtx* ctx = current_ctx(&err);
  const int sms = ctx ? ctx->sm_count : 148;
  bool swap;
  int splits, per;
  wgrad_plan(rows, n_in, n_out, sms, &swap, &splits, &per);
  return (int64_t)splits * ((int64_t)n_in * n_out + n_out);
}

// dW[n_out x n_in] = G^T X, db[n_out] = column sums of G, over `rows` rows: split over the rows, per-split partial
// products in `scratch` (dense_linear_wgrad_scratch_elems doubles), then a fixed-order sum (bit-reproducible).
extern "C" int dense_linear_wgrad_f64(const double* g, const double* x, double* dw, double* db, double* scratch,
                                      int64_t rows, int n_in, int n_out, void* stream) {
  if (!g || !x || !dw || !db || !scratch || rows <= 0 || n_in <= 0 || n_out <= 0 || rows > (1 << 30)) return B2048_EINVAL;
  if ((n_in & 1) || (n_out & 1) || ((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(x) |
                                     reinterpret_cast<uintptr_t>(scratch)) & 15u))
    return B2048_EINVAL;
  DG_CTX();
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  bool swap;
  int splits, per;
  wgrad_plan(rows, n_in, n_out, ctx->sm_count, &swap, &splits, &per);
  const int64_t n_main = (int64_t)n_in * n_out;
  double* colsum = scratch + (int64_t)splits * n_main;
  cudaError_t e;
  if (!swap) {
    // P(i = output unit, j = input unit) = sum_m G(m, i) X(m, j): both operands reduction-strided;
    // db = sums of G over the rows = "column sums of A"
    DgArgs p{g, x, scratch, nullptr, colsum, 1, n_out, n_in, n_in, n_out, n_in, (int)rows, per};
    if (n_out >= 128) e = dg_launch<16, 2, false, false, EPI_PARTIAL>(p, splits, st);
    else if (n_out > 32) e = dg_launch<8, 2, false, false, EPI_PARTIAL>(p, splits, st);
    else e = dg_launch<4, 2, false, false, EPI_PARTIAL>(p, splits, st);
    if (e != cudaSuccess) return (int)e;
    const int64_t total = n_main + n_out;
    dgemm_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(scratch, colsum, dw, db, n_main, n_out, splits,
                                                                         0, n_out, n_in);
  } else {
    // few inputs (the first layer, 16): P^T(i = input unit, j = output unit) = sum_m X(m, i) G(m, j), so that the
    // wide dimension fills the eight warps; db = sums of G = "column sums of B"; the reduce pass transposes
    DgArgs p{x, g, scratch, nullptr, colsum, 2, n_in, n_out, n_out, n_in, n_out, (int)rows, per};
    e = (n_in > 32) ? dg_launch<8, 2, false, false, EPI_PARTIAL>(p, splits, st)
        : (n_in > 16) ? dg_launch<4, 2, false, false, EPI_PARTIAL>(p, splits, st)
                      : dg_launch<2, 2, false, false, EPI_PARTIAL>(p, splits, st);
    if (e != cudaSuccess) return (int)e;
    const int64_t total = n_main + n_out;
    dgemm_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(scratch, colsum, dw, db, n_main, n_out, splits,
                                                                         1, n_in, n_out);
  }
  return (int)cudaGetLastError();
}
