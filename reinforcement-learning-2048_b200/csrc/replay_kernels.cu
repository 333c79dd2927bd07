// replay_kernels.cu — K2: GPU-resident ring replay buffer (sm_100a).
//
// Replaces the reference's deque(maxlen) of (board, action, reward, next_board, done) tuples
// (src/dqn_lib.py:106,172) and sample_experiences + extract_samples_conv/dense
// (src/dqn_lib.py:33-84): transitions are stored as packed u64 boards (22 B per transition) and a
// single kernel draws the batch indices, gathers the five fields and unpacks both boards straight
// into the Q-network's float64 input layout.
#include "b2048_common.cuh"

namespace b2048 {
namespace {

__global__ void ring_append_kernel(b2048_ring ring, const uint64_t* __restrict__ s,
                                   const uint8_t* __restrict__ a, const int32_t* __restrict__ r,
                                   const uint64_t* __restrict__ s2,
                                   const uint8_t* __restrict__ done_flags, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (i < n - ring.capacity) return;  // deque(maxlen): only the newest `capacity` items survive
  const int64_t head = ring.head_size[0];
  const int64_t slot = (head + i) % ring.capacity;
  ring.s[slot] = s[i];
  ring.s2[slot] = s2[i];
  ring.r[slot] = r[i];
  ring.a[slot] = a[i] & 3u;
  ring.d[slot] = (done_flags[i] & B2048_FLAG_DONE) ? 1 : 0;
}

__global__ void ring_advance_kernel(b2048_ring ring, int64_t n) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const int64_t head = ring.head_size[0], size = ring.head_size[1];
    ring.head_size[0] = (head + n) % ring.capacity;
    ring.head_size[1] = (size + n > ring.capacity) ? ring.capacity : size + n;
  }
}

// 16 threads per sample: thread (j, c) writes cell c of states[j] and next_states[j]; lane c == 0
// also writes the scalars.  Writes are fully coalesced (16 consecutive doubles per sample).
__global__ void ring_sample_kernel(b2048_ring ring, int64_t B, uint64_t seed, uint64_t ctr,
                                   const int64_t* __restrict__ idx_override,
                                   double* __restrict__ states, double* __restrict__ next_states,
                                   int64_t* __restrict__ actions, int64_t* __restrict__ rewards,
                                   int64_t* __restrict__ dones, int64_t* __restrict__ idx_out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t j = t >> 4;
  const int c = (int)(t & 15);
  if (j >= B) return;
  const int64_t head = ring.head_size[0], size = ring.head_size[1];
  if (ctr == B2048_CTR_AUTO) ctr = (uint64_t)ring.head_size[2];
  int64_t idx;
  if (idx_override) {
    idx = idx_override[j];
  } else {
    const uint4 w4 = philox_at(seed, DOM_SAMPLE, (uint64_t)(j >> 2), ctr);
    const uint32_t w = (j & 2) ? ((j & 1) ? w4.w : w4.z) : ((j & 1) ? w4.y : w4.x);
    idx = (int64_t)(((uint64_t)w * (uint64_t)size) >> 32);
  }
  if (size > 0) {
    idx %= size;
    if (idx < 0) idx += size;  // python-style negative index
  } else {
    idx = 0;
  }
  const int64_t slot = (head + ring.capacity - size + idx) % ring.capacity;
  const uint64_t sb = ring.s[slot], nb = ring.s2[slot];
  states[j * 16 + c] = (double)((uint32_t)(sb >> (4 * c)) & 0xFu);
  next_states[j * 16 + c] = (double)((uint32_t)(nb >> (4 * c)) & 0xFu);
  if (c == 0) {
    actions[j] = ring.a[slot];
    rewards[j] = ring.r[slot];
    dones[j] = ring.d[slot];
    if (idx_out) idx_out[j] = idx;
  }
}

__global__ void ring_bump_kernel(b2048_ring ring) {
  if (threadIdx.x == 0 && blockIdx.x == 0) ring.head_size[2] += 1;
}

}  // namespace
}  // namespace b2048

using namespace b2048;

static int ring_ok(const b2048_ring* ring) {
  return ring && ring->s && ring->s2 && ring->r && ring->a && ring->d && ring->head_size &&
         ring->capacity > 0;
}

extern "C" int replay_append(const b2048_ring* ring, const uint64_t* s, const uint8_t* a,
                             const int32_t* r, const uint64_t* s2, const uint8_t* done_flags,
                             int64_t n, void* stream) {
  if (!ring_ok(ring) || n < 0) return B2048_EINVAL;
  if (n == 0) return B2048_OK;
  if (!s || !a || !r || !s2 || !done_flags) return B2048_EINVAL;
  int err = 0;
  if (!current_ctx(&err)) return err;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ring_append_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(*ring, s, a, r, s2, done_flags, n);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  ring_advance_kernel<<<1, 32, 0, st>>>(*ring, n);
  return (int)cudaGetLastError();
}

extern "C" int replay_sample(const b2048_ring* ring, int64_t B, uint64_t seed, uint64_t ctr,
                             const int64_t* idx_override, double* states, double* next_states,
                             int64_t* actions, int64_t* rewards, int64_t* dones, int64_t* idx_out,
                             void* stream) {
  if (!ring_ok(ring) || B < 0) return B2048_EINVAL;
  if (B == 0) return B2048_OK;
  if (!states || !next_states || !actions || !rewards || !dones) return B2048_EINVAL;
  int err = 0;
  if (!current_ctx(&err)) return err;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t threads = B * 16;
  ring_sample_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(
      *ring, B, seed, ctr, idx_override, states, next_states, actions, rewards, dones, idx_out);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess || ctr != B2048_CTR_AUTO) return (int)e;
  ring_bump_kernel<<<1, 32, 0, st>>>(*ring);
  return (int)cudaGetLastError();
}
