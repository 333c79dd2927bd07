// ddqn_kernels.cu — K3: fused Double-DQN target + summed-MSE residual, K0: batched epsilon-greedy.
//
// K3 replaces src/dqn_lib.py:125-158 (argmax over online Q(s'), one-hot gather of target Q(s',a*),
// r + gamma32*(1-done)*Q, one-hot gather of Q(s,a), MSELoss(reduction='sum')) with one kernel that
// also emits d loss / d Q(s,.) for the autograd backward.  K0 replaces
// epsilon_greedy_policy (src/dqn_lib.py:16-30) for n boards.  All arithmetic is float64 with the
// reference's rounding points kept (explicit _rn intrinsics: no FMA contraction), and gamma is
// applied as a float32 value exactly like the reference does.
#include "b2048_common.cuh"

namespace b2048 {
namespace {

// One thread-block cluster (8 CTAs x 512 threads, one per SM) handles the whole batch: per-CTA partial
// sums meet in CTA 0 through distributed shared memory, so the launch needs NO scratch in global memory
// and is re-entrant — any number of calls may be in flight on different streams of one device
// (include/b2048.h).  The order of every addition is fixed: bit-reproducible losses.
constexpr int K3_THREADS = 512;
constexpr int K3_CLUSTER = 8;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ double ld_dsmem_f64(const double* local_smem_ptr, uint32_t cta) {
  uint32_t a = (uint32_t)__cvta_generic_to_shared(local_smem_ptr), ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(cta));
  double v;
  asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(ra) : "memory");
  return v;
}

__device__ __forceinline__ double block_sum(double v, double* sh) {
  // fixed-order tree: warp shuffle, then the warp partials added in order by thread 0
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = __dadd_rn(v, __shfl_down_sync(0xFFFFFFFFu, v, o));
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 0; w < K3_THREADS / 32; ++w) s = __dadd_rn(s, sh[w]);
  }
  return s;
}

__global__ void __launch_bounds__(K3_THREADS)
    ddqn_target_loss_kernel(const double* __restrict__ qno, const double* __restrict__ qnt,
                            const double* __restrict__ qc, const int64_t* __restrict__ actions,
                            const int64_t* __restrict__ rewards, const int64_t* __restrict__ dones,
                            float gamma, int use_double, double* __restrict__ target,
                            double* __restrict__ q_sa, double* __restrict__ loss,
                            double* __restrict__ grad, int64_t B) {
  __shared__ double sh[K3_THREADS / 32];
  __shared__ double cta_sum;
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * K3_THREADS + threadIdx.x; i < B;
       i += (int64_t)gridDim.x * K3_THREADS) {
    const double2* t2 = reinterpret_cast<const double2*>(qnt + 4 * i);
    const double2 ta = t2[0], tb = t2[1];
    const double tq[4] = {ta.x, ta.y, tb.x, tb.y};
    double nb;
    if (use_double) {
      const double2* o2 = reinterpret_cast<const double2*>(qno + 4 * i);
      const double2 oa = o2[0], ob = o2[1];
      const double oq[4] = {oa.x, oa.y, ob.x, ob.y};
      int best = 0;
      double bv = oq[0];
#pragma unroll
      for (int j = 1; j < 4; ++j)
        if (oq[j] > bv) { bv = oq[j]; best = j; }  // strict '>' keeps the first maximum
      nb = tq[best];
    } else {
      nb = fmax(fmax(tq[0], tq[1]), fmax(tq[2], tq[3]));
    }
    const float g32 = __fmul_rn((float)(1 - dones[i]), gamma);  // int64 * python float -> float32
    const double tgt = __dadd_rn((double)rewards[i], __dmul_rn((double)g32, nb));
    const int a = (int)(actions[i] & 3);
    const double2* c2 = reinterpret_cast<const double2*>(qc + 4 * i);
    const double2 ca = c2[0], cb = c2[1];
    const double cq[4] = {ca.x, ca.y, cb.x, cb.y};
    const double q = cq[a];
    const double diff = __dsub_rn(q, tgt);
    target[i] = tgt;
    q_sa[i] = q;
    if (grad) {
      const double g = __dmul_rn(2.0, diff);
      double2* g2 = reinterpret_cast<double2*>(grad + 4 * i);
      g2[0] = make_double2(a == 0 ? g : 0.0, a == 1 ? g : 0.0);
      g2[1] = make_double2(a == 2 ? g : 0.0, a == 3 ? g : 0.0);
    }
    acc = __dadd_rn(acc, __dmul_rn(diff, diff));
  }
  const double bs = block_sum(acc, sh);
  if (threadIdx.x == 0) cta_sum = bs;
  cluster_sync_all();                       // every CTA's partial is visible cluster-wide
  if (cluster_ctarank() == 0 && threadIdx.x == 0) {
    double s = 0.0;
    for (uint32_t c = 0; c < (uint32_t)K3_CLUSTER; ++c) s = __dadd_rn(s, ld_dsmem_f64(&cta_sum, c));
    loss[0] = s;
  }
  cluster_sync_all();                       // peers keep their shared memory alive until CTA 0 has read it
}

// Activations are row matrices [b*h*w, c] (what a GEMM over patches produces; for c = 1 the same bytes as
// NCHW), so consecutive convolutions need no layout round trip.
// cols[(b*oh + oy)*ow + ox][(ci*kh + ky)*kw + kx] = x[(b*h + oy+ky)*w + ox+kx][ci] — the (c, kh, kw)
// column order of conv.weight.reshape(out, -1).
// One thread per (patch, channel): it writes kh*kw consecutive doubles; consecutive threads take
// consecutive channels, so a warp reads and writes contiguous runs.  Index type is 32-bit when it fits.
template <typename I>
__global__ void patches_kernel(const double* __restrict__ x, double* __restrict__ cols, I total, int c, int h,
                               int w, int kh, int kw, int oh, int ow) {
  const I i = (I)blockIdx.x * blockDim.x + threadIdx.x;   // = patch * c + ci
  if (i >= total) return;
  const I patch = i / (I)c;
  const int ci = (int)(i - patch * (I)c);
  const int ox = (int)(patch % (I)ow);
  const I t = patch / (I)ow;
  const int oy = (int)(t % (I)oh);
  const I b = t / (I)oh;
  const double* src = x + (((b * h + oy) * (I)w + ox) * (I)c + ci);
  double* dst = cols + i * (I)(kh * kw);
  for (int ky = 0; ky < kh; ++ky)
    for (int kx = 0; kx < kw; ++kx) dst[ky * kw + kx] = src[(I)(ky * w + kx) * (I)c];
}

// dx[(b*h + y)*w + x][ci] = sum over (ky,kx) with 0 <= y-ky < oh, 0 <= x-kx < ow of
// dcols[patch (y-ky, x-kx)][(ci*kh + ky)*kw + kx]: a gather per input element (deterministic).
template <typename I>
__global__ void patches_grad_kernel(const double* __restrict__ dcols, double* __restrict__ dx, I total, int c,
                                    int h, int w, int kh, int kw, int oh, int ow) {
  const I i = (I)blockIdx.x * blockDim.x + threadIdx.x;   // = ((b*h + y)*w + x)*c + ci
  if (i >= total) return;
  const I pix = i / (I)c;
  const int ci = (int)(i - pix * (I)c);
  const int xx = (int)(pix % (I)w);
  const I t = pix / (I)w;
  const int yy = (int)(t % (I)h);
  const I b = t / (I)h;
  const I K = (I)c * kh * kw;
  double acc = 0.0;
  for (int ky = 0; ky < kh; ++ky) {
    const int oy = yy - ky;
    if (oy < 0 || oy >= oh) continue;
    for (int kx = 0; kx < kw; ++kx) {
      const int ox = xx - kx;
      if (ox < 0 || ox >= ow) continue;
      acc += dcols[((b * oh + oy) * ow + ox) * K + (I)((ci * kh + ky) * kw + kx)];
    }
  }
  dx[i] = acc;
}

__global__ void adam_kernel(double* __restrict__ p, const double* __restrict__ g, double* __restrict__ m,
                            double* __restrict__ v, const int64_t* __restrict__ step, int64_t n, double lr,
                            double b1, double b2, double eps) {
  __shared__ double s_step_size, s_inv_bc2_sqrt;
  if (threadIdx.x == 0) {
    const double t = (double)(step[0] + 1);
    s_step_size = lr / (1.0 - pow(b1, t));
    s_inv_bc2_sqrt = 1.0 / sqrt(1.0 - pow(b2, t));
  }
  __syncthreads();
  const double step_size = s_step_size, inv_bc2_sqrt = s_inv_bc2_sqrt;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double gi = g[i];
    const double mi = b1 * m[i] + (1.0 - b1) * gi;
    const double vi = b2 * v[i] + (1.0 - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= step_size * mi / (sqrt(vi) * inv_bc2_sqrt + eps);
  }
}

__global__ void adam_bump_kernel(int64_t* step) {
  if (threadIdx.x == 0 && blockIdx.x == 0) step[0] += 1;
}

__global__ void egreedy_kernel(const double* __restrict__ q, const uint8_t* __restrict__ flags,
                               double eps, uint64_t seed, uint64_t ctr, uint64_t index_base,
                               const uint8_t* __restrict__ override_bytes,
                               uint8_t* __restrict__ actions, double* __restrict__ max_q, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t ov = override_bytes ? (uint32_t)override_bytes[i] : 0xFFu;
  bool explore;
  uint32_t ra;
  if (ov == 0xFFu) {
    const uint64_t g = index_base + (uint64_t)i;
    const uint4 r = philox_at(seed, DOM_EGREEDY, g >> 1, ctr);
    const uint32_t wu = (g & 1ull) ? r.z : r.x, wa = (g & 1ull) ? r.w : r.y;
    explore = ((double)wu * (1.0 / 4294967296.0)) < eps;  // np.random.rand() < epsilon
    ra = wa >> 30;                                        // np.random.randint(4)
  } else {
    explore = (ov & 0x80u) == 0;
    ra = ov & 3u;
  }
  if (explore) {
    actions[i] = (uint8_t)ra;
    max_q[i] = 0.0;
    return;
  }
  const double2* q2 = reinterpret_cast<const double2*>(q + 4 * i);
  const double2 qa = q2[0], qb = q2[1];
  const double qq[4] = {qa.x, qa.y, qb.x, qb.y};
  const double mn = fmin(fmin(qq[0], qq[1]), fmin(qq[2], qq[3]));
  const double mx = fmax(fmax(qq[0], qq[1]), fmax(qq[2], qq[3]));
  const double mm = __dmul_rn(mn, mx);
  const uint32_t legal = flags[i] & 0xFu;
  int best = 0;
  double bv = 0.0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const double qn = __dsub_rn(__dsub_rn(qq[j], mm), mn);  // Q - min*max - min, left to right
    const double v = __dmul_rn(((legal >> j) & 1u) ? 1.0 : 0.0, qn);
    if (j == 0 || v > bv) { bv = v; best = j; }
  }
  actions[i] = (uint8_t)best;
  max_q[i] = mx;
}

}  // namespace
}  // namespace b2048

using namespace b2048;

extern "C" int ddqn_target_loss(const double* q_next_online, const double* q_next_target,
                                const double* q_cur, const int64_t* actions, const int64_t* rewards,
                                const int64_t* dones, float gamma_f32, int use_double,
                                double* target, double* q_sa, double* loss, double* grad_q_cur,
                                int64_t B, void* stream) {
  if (B <= 0) return B2048_EINVAL;
  if (!q_next_target || !q_cur || !actions || !rewards || !dones || !target || !q_sa || !loss)
    return B2048_EINVAL;
  if (use_double && !q_next_online) return B2048_EINVAL;
  int err = 0;
  if (!current_ctx(&err)) return err;
  // one cluster of K3_CLUSTER CTAs, whatever B is (grid-stride inside): no global scratch, re-entrant
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(K3_CLUSTER, 1, 1);
  cfg.blockDim = dim3(K3_THREADS, 1, 1);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = static_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = K3_CLUSTER;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return (int)cudaLaunchKernelEx(&cfg, ddqn_target_loss_kernel, q_next_online, q_next_target, q_cur, actions,
                                 rewards, dones, gamma_f32, use_double, target, q_sa, loss, grad_q_cur, B);
}

static int patches_args_ok(int64_t n, int c, int h, int w, int kh, int kw) {
  return n > 0 && c > 0 && kh > 0 && kw > 0 && h >= kh && w >= kw;
}

extern "C" int conv_patches_f64(const double* x, double* cols, int64_t n, int c, int h, int w, int kh, int kw,
                                void* stream) {
  if (!x || !cols || !patches_args_ok(n, c, h, w, kh, kw)) return B2048_EINVAL;
  int err = 0;
  if (!current_ctx(&err)) return err;
  const int oh = h - kh + 1, ow = w - kw + 1;
  const int64_t total = n * oh * ow * c;
  const unsigned grid = (unsigned)((total + 255) / 256);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (total * kh * kw < (1ll << 31) && n * c * h * w < (1ll << 31))
    patches_kernel<uint32_t><<<grid, 256, 0, st>>>(x, cols, (uint32_t)total, c, h, w, kh, kw, oh, ow);
  else
    patches_kernel<int64_t><<<grid, 256, 0, st>>>(x, cols, total, c, h, w, kh, kw, oh, ow);
  return (int)cudaGetLastError();
}

extern "C" int conv_patches_grad_f64(const double* dcols, double* dx, int64_t n, int c, int h, int w, int kh,
                                     int kw, void* stream) {
  if (!dcols || !dx || !patches_args_ok(n, c, h, w, kh, kw)) return B2048_EINVAL;
  int err = 0;
  if (!current_ctx(&err)) return err;
  const int oh = h - kh + 1, ow = w - kw + 1;
  const int64_t total = n * c * h * w;
  const unsigned grid = (unsigned)((total + 255) / 256);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n * oh * ow * c * kh * kw < (1ll << 31) && total < (1ll << 31))
    patches_grad_kernel<uint32_t><<<grid, 256, 0, st>>>(dcols, dx, (uint32_t)total, c, h, w, kh, kw, oh, ow);
  else
    patches_grad_kernel<int64_t><<<grid, 256, 0, st>>>(dcols, dx, total, c, h, w, kh, kw, oh, ow);
  return (int)cudaGetLastError();
}

extern "C" int ddqn_adam_step(double* params, const double* grads, double* exp_avg, double* exp_avg_sq,
                              int64_t* step_counter, int64_t n, double lr, double beta1, double beta2,
                              double eps, void* stream) {
  if (n <= 0 || !params || !grads || !exp_avg || !exp_avg_sq || !step_counter) return B2048_EINVAL;
  int err = 0;
  if (!current_ctx(&err)) return err;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int64_t blocks = (n + 255) / 256;
  if (blocks > 1184) blocks = 1184;   // 8 CTAs of 256 threads per SM
  adam_kernel<<<(unsigned)blocks, 256, 0, st>>>(params, grads, exp_avg, exp_avg_sq, step_counter, n, lr, beta1,
                                                beta2, eps);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  adam_bump_kernel<<<1, 32, 0, st>>>(step_counter);
  return (int)cudaGetLastError();
}

extern "C" int egreedy_select(const double* q, const uint8_t* flags, double eps, uint64_t seed,
                              uint64_t ctr, uint64_t index_base, const uint8_t* override_bytes,
                              uint8_t* actions, double* max_q, int64_t n, void* stream) {
  if (n < 0) return B2048_EINVAL;
  if (n == 0) return B2048_OK;
  if (!q || !flags || !actions || !max_q) return B2048_EINVAL;
  int err = 0;
  if (!current_ctx(&err)) return err;
  egreedy_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      q, flags, eps, seed, ctr, index_base, override_bytes, actions, max_q, n);
  return (int)cudaGetLastError();
}
