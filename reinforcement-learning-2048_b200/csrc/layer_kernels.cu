// layer_kernels.cu — the element-wise halves of the Q-networks' float64 layers, fused.
//
// The reference's networks (src/configs/double_dqn_conv.py:19-28, double_dqn_dense.py:7-15) are
// Conv2d / Linear layers followed by ReLU, trained through autograd in train_step
// (src/dqn_lib.py:146-163).  The GEMMs stay with cuBLAS (plain library DGEMMs); what ATen spends
// around them — broadcasting the bias into the output, a separate clamp kernel, threshold_backward,
// a generic reduce_kernel for every bias gradient (17 us each at batch 5000), NCHW <-> row-matrix
// copies — is two kernels here:
//   layer_bias_act_f64       y = act(y + bias), in place on the GEMM output
//   layer_act_grad_bias_f64  g = gy * (y > 0) and dbias = column sums of g, one pass, deterministic
// plus im2col / col2im for activations kept as [N*H*W, C] row matrices (the layout the GEMMs produce).
#include "b2048_common.cuh"

namespace b2048 {
namespace {

__global__ void __launch_bounds__(256)
    bias_act_kernel(double2* __restrict__ y, const double2* __restrict__ bias, int64_t total2, int cols2, int relu) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total2; i += (int64_t)gridDim.x * 256) {
    double2 v = y[i];
    const double2 b = bias[(int)(i % cols2)];
    v.x += b.x;
    v.y += b.y;
    if (relu) {
      v.x = fmax(v.x, 0.0);
      v.y = fmax(v.y, 0.0);
    }
    y[i] = v;
  }
}

constexpr int AG_ROWS = 64;   // rows per block of the backward kernel

// Block b owns rows [b*AG_ROWS, +AG_ROWS).  Threads are (tx = column pair, ty = row lane); each adds its
// rows in order, the row lanes are combined in order through shared memory, the per-block column sums go
// to `partials`, and the last block to finish adds the blocks in order: the result does not depend on
// scheduling (no atomics on the data).
__global__ void __launch_bounds__(256)
    act_grad_bias_kernel(const double2* gy, const double2* __restrict__ y, double2* g,   // g may alias gy
                         double* __restrict__ dbias, int64_t rows, int cols2, int relu,
                         double2* __restrict__ partials, unsigned int* __restrict__ ticket) {
  __shared__ double2 sm[256];
  __shared__ bool is_last;
  const int tpr = cols2 < 256 ? cols2 : 256;        // threads per row pass
  const int lanes = 256 / tpr;                      // row lanes (threads beyond lanes * tpr idle)
  const int tx = threadIdx.x % tpr, ty = threadIdx.x / tpr;
  const int64_t r0 = (int64_t)blockIdx.x * AG_ROWS;
  const int64_t r1 = r0 + AG_ROWS < rows ? r0 + AG_ROWS : rows;
  for (int cb = 0; cb < cols2; cb += tpr) {
    const int c2 = cb + tx;
    double2 acc = make_double2(0.0, 0.0);
    if (ty < lanes && c2 < cols2) {
      for (int64_t r = r0 + ty; r < r1; r += lanes) {
        double2 v = gy[r * cols2 + c2];
        if (relu) {
          const double2 a = y[r * cols2 + c2];
          if (!(a.x > 0.0)) v.x = 0.0;
          if (!(a.y > 0.0)) v.y = 0.0;
        }
        g[r * cols2 + c2] = v;
        acc.x += v.x;
        acc.y += v.y;
      }
    }
    sm[threadIdx.x] = acc;
    __syncthreads();
    if (ty == 0 && c2 < cols2) {
      double2 s = sm[tx];
      for (int j = 1; j < lanes; ++j) {
        s.x += sm[j * tpr + tx].x;
        s.y += sm[j * tpr + tx].y;
      }
      partials[(int64_t)blockIdx.x * cols2 + c2] = s;
    }
    __syncthreads();
  }
  __threadfence();
  if (threadIdx.x == 0) is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  for (int c2 = threadIdx.x; c2 < cols2; c2 += 256) {
    double2 s = make_double2(0.0, 0.0);
    for (unsigned b = 0; b < gridDim.x; ++b) {
      const double2 p = __ldcg(&partials[(int64_t)b * cols2 + c2]);
      s.x += p.x;
      s.y += p.y;
    }
    dbias[2 * c2] = s.x;
    dbias[2 * c2 + 1] = s.y;
  }
  if (threadIdx.x == 0) *ticket = 0u;
}

// im2col of an activation stored as rows (b, y, x) x c:  cols[(b, oy, ox)][(ci*kh + ky)*kw + kx] =
// x[(b, oy+ky, ox+kx)][ci] — the (c, kh, kw) column order of conv.weight.reshape(out, -1).
__global__ void patches_rows_kernel(const double* __restrict__ x, double* __restrict__ cols, uint32_t total, int c,
                                    int h, int w, int kh, int kw, int oh, int ow) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;   // = patch * c + ci
  if (i >= total) return;
  const uint32_t patch = i / (uint32_t)c;
  const int ci = (int)(i - patch * (uint32_t)c);
  const int ox = (int)(patch % (uint32_t)ow);
  const uint32_t t = patch / (uint32_t)ow;
  const int oy = (int)(t % (uint32_t)oh);
  const uint32_t b = t / (uint32_t)oh;
  const double* src = x + ((size_t)(b * h + oy) * w + ox) * c + ci;
  double* dst = cols + (size_t)i * (kh * kw);
  for (int ky = 0; ky < kh; ++ky)
    for (int kx = 0; kx < kw; ++kx) dst[ky * kw + kx] = src[(size_t)(ky * w + kx) * c];
}

// col2im into the same row layout: dx[(b, y, x)][ci] = sum of the dcols entries that read it.
__global__ void patches_rows_grad_kernel(const double* __restrict__ dcols, double* __restrict__ dx, uint32_t total,
                                         int c, int h, int w, int kh, int kw, int oh, int ow) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;   // = ((b*h + y)*w + x)*c + ci
  if (i >= total) return;
  const uint32_t pix = i / (uint32_t)c;
  const int ci = (int)(i - pix * (uint32_t)c);
  const int xx = (int)(pix % (uint32_t)w);
  const uint32_t t = pix / (uint32_t)w;
  const int yy = (int)(t % (uint32_t)h);
  const uint32_t b = t / (uint32_t)h;
  const size_t K = (size_t)c * kh * kw;
  double acc = 0.0;
  for (int ky = 0; ky < kh; ++ky) {
    const int oy = yy - ky;
    if (oy < 0 || oy >= oh) continue;
    for (int kx = 0; kx < kw; ++kx) {
      const int ox = xx - kx;
      if (ox < 0 || ox >= ow) continue;
      acc += dcols[((size_t)(b * oh + oy) * ow + ox) * K + (size_t)((ci * kh + ky) * kw + kx)];
    }
  }
  dx[i] = acc;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace
}  // namespace b2048

using namespace b2048;

extern "C" int layer_bias_act_f64(double* y, const double* bias, int64_t rows, int cols, int relu, void* stream) {
  if (rows < 0 || cols <= 0 || (cols & 1) || !bias || (rows > 0 && !y) || !aligned16(y) || !aligned16(bias))
    return B2048_EINVAL;
  if (rows == 0) return B2048_OK;
  int err = 0;
  if (!current_ctx(&err)) return err;
  const int64_t total2 = rows * (cols / 2);
  int64_t blocks = (total2 + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  bias_act_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<double2*>(y), reinterpret_cast<const double2*>(bias), total2, cols / 2, relu);
  return (int)cudaGetLastError();
}

extern "C" int64_t layer_act_grad_scratch_elems(int64_t rows, int cols) {
  return ((rows + AG_ROWS - 1) / AG_ROWS) * (int64_t)cols;
}

extern "C" int layer_act_grad_bias_f64(const double* gy, const double* y, double* g, double* dbias, double* scratch,
                                       int64_t rows, int cols, int relu, void* stream) {
  if (rows <= 0 || cols <= 0 || (cols & 1) || !gy || !g || !dbias || !scratch || (relu && !y) || !aligned16(gy) ||
      !aligned16(g) || !aligned16(scratch) || (relu && !aligned16(y)))
    return B2048_EINVAL;
  int err = 0;
  DeviceCtx* ctx = current_ctx(&err);
  if (!ctx) return err;
  const int64_t blocks = (rows + AG_ROWS - 1) / AG_ROWS;
  if (blocks > 0x7FFFFFFF) return B2048_EINVAL;
  act_grad_bias_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const double2*>(gy), reinterpret_cast<const double2*>(y), reinterpret_cast<double2*>(g), dbias,
      rows, cols / 2, relu, reinterpret_cast<double2*>(scratch), ctx->ticket2);
  return (int)cudaGetLastError();
}

static bool rows_args_ok(int64_t n, int c, int h, int w, int kh, int kw) {
  return n > 0 && c > 0 && h > 0 && w > 0 && kh > 0 && kw > 0 && kh <= h && kw <= w &&
         n * (int64_t)(h - kh + 1) * (w - kw + 1) * c * kh * kw < (1ll << 31) && n * (int64_t)c * h * w < (1ll << 31);
}

extern "C" int conv_patches_rows_f64(const double* x, double* cols, int64_t n, int c, int h, int w, int kh, int kw,
                                     void* stream) {
  if (!x || !cols || !rows_args_ok(n, c, h, w, kh, kw)) return B2048_EINVAL;
  int err = 0;
  if (!current_ctx(&err)) return err;
  const int oh = h - kh + 1, ow = w - kw + 1;
  const uint32_t total = (uint32_t)(n * oh * ow * c);
  patches_rows_kernel<<<(total + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, cols, total, c, h, w, kh,
                                                                                        kw, oh, ow);
  return (int)cudaGetLastError();
}

extern "C" int conv_patches_rows_grad_f64(const double* dcols, double* dx, int64_t n, int c, int h, int w, int kh,
                                          int kw, void* stream) {
  if (!dcols || !dx || !rows_args_ok(n, c, h, w, kh, kw)) return B2048_EINVAL;
  int err = 0;
  if (!current_ctx(&err)) return err;
  const int oh = h - kh + 1, ow = w - kw + 1;
  const uint32_t total = (uint32_t)(n * c * h * w);
  patches_rows_grad_kernel<<<(total + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(dcols, dx, total, c, h,
                                                                                             w, kh, kw, oh, ow);
  return (int)cudaGetLastError();
}
