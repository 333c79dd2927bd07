// p2p_kernels.cu — gradient allreduce + Adam in ONE kernel over NVLink peer memory (sm_100a).
//
// The reference is single-GPU; the data-parallel build sums the online Q-network's gradient over
// the ranks once per update (268 KB conv / 3.2 MB dense in float64: latency-bound).
// Instead of NCCL allreduce + a separate optimizer kernel, each rank runs one kernel that reads the
// peers' gradient buffers directly over NVLink (ld.global on IPC-mapped peer pointers), sums them in
// a fixed rank order and applies the Adam update to its own replica.  Two flag barriers (release /
// acquire at system scope) order the peers' reads against the local writes before and after.
#include <string.h>

#include "b2048_common.cuh"

namespace b2048 {
namespace {

constexpr int P2P_THREADS = 256;
// Bounded wait: a peer that has not arrived after this long (wall clock, %globaltimer) is treated as lost.
// The kernel then raises the error flag (sync_state[2], sticky) and leaves WITHOUT touching the replica:
// summing a stale or half-written peer buffer would make the replicas diverge silently.  The host side
// (PeerGradExchange.check, called wherever the trainers read a scalar back) turns the flag into an
// exception.
constexpr unsigned long long WAIT_LIMIT_NS = 30ull * 1000000000ull;
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ void st_release_sys(uint64_t* p, uint64_t v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t ld_acquire_sys(const uint64_t* p) {
  uint64_t v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double ld_peer(const double* p) {   // peer data is written once per epoch: bypass L1
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

// wait until flags[i] >= epoch for all i < world (threads 0..world-1 poll one flag each);
// returns false (for every thread of the block) if the wait expired or the error flag is already up
__device__ __forceinline__ bool wait_flags(const uint64_t* flags, int world, uint64_t epoch, uint64_t* err) {
  if ((int)threadIdx.x < world) {
    const unsigned long long t0 = global_ns();
    unsigned int spins = 0;
    while (ld_acquire_sys(flags + threadIdx.x) < epoch) {
      if ((++spins & 1023u) == 0 && global_ns() - t0 > WAIT_LIMIT_NS) {
        st_release_sys(err, 1ull);
        break;
      }
      __nanosleep(64);
    }
  }
  __syncthreads();
  return ld_acquire_sys(err) == 0;
}

__global__ void __launch_bounds__(P2P_THREADS)
    p2p_allreduce_adam_kernel(const double* const* __restrict__ peer_grads, uint64_t* const* __restrict__ peer_flags,
                              uint64_t* __restrict__ sync_state, int rank, int world, double* __restrict__ p,
                              double* __restrict__ m, double* __restrict__ v, int64_t* __restrict__ step, int64_t n,
                              double lr, double b1, double b2, double eps) {
  __shared__ double s_step_size, s_inv_bc2_sqrt;
  __shared__ bool s_last;
  const uint64_t epoch = sync_state[0] + 1;
  uint64_t* my_flags = peer_flags[rank];

  // ---- barrier 1: every rank's gradient buffer is complete ----------------------------------------
  if (blockIdx.x == 0 && (int)threadIdx.x < world) {
    __threadfence_system();                                   // this rank's backward wrote the buffer in earlier kernels
    st_release_sys(peer_flags[threadIdx.x] + rank, epoch);    // flag block 0 of peer `threadIdx.x`, slot `rank`
  }
  if (threadIdx.x == 0) {
    const double t = (double)(step[0] + 1);
    s_step_size = lr / (1.0 - pow(b1, t));
    s_inv_bc2_sqrt = 1.0 / sqrt(1.0 - pow(b2, t));
  }
  if (!wait_flags(my_flags, world, epoch, sync_state + 2)) return;   // a peer is missing: no update, error stays up

  // ---- sum over ranks in rank order + Adam ----------------------------------------------------------
  const double step_size = s_step_size, inv_bc2_sqrt = s_inv_bc2_sqrt;
  for (int64_t i = (int64_t)blockIdx.x * P2P_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * P2P_THREADS) {
    double g = 0.0;
    for (int r = 0; r < world; ++r) g += ld_peer(peer_grads[r] + i);
    const double mi = b1 * m[i] + (1.0 - b1) * g;
    const double vi = b2 * v[i] + (1.0 - b2) * g * g;
    m[i] = mi;
    v[i] = vi;
    p[i] -= step_size * mi / (sqrt(vi) * inv_bc2_sqrt + eps);
  }

  // ---- barrier 2: every rank has finished reading; the last block of this grid does the signalling ----
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned long long done = atomicAdd(reinterpret_cast<unsigned long long*>(sync_state + 1), 1ull);
    s_last = (done == (unsigned long long)gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  if ((int)threadIdx.x < world) st_release_sys(peer_flags[threadIdx.x] + world + rank, epoch);   // flag block 1
  if (!wait_flags(my_flags + world, world, epoch, sync_state + 2)) return;   // epoch not published: the next call fails too
  if (threadIdx.x == 0) {
    sync_state[1] = 0;
    sync_state[0] = epoch;
    step[0] += 1;
  }
}

}  // namespace
}  // namespace b2048

using namespace b2048;

extern "C" int p2p_open_ipc_handle(const unsigned char* handle64, void** out) {
  if (!handle64 || !out) return B2048_EINVAL;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  return (int)cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess);
}

extern "C" int p2p_get_ipc_handle(const void* ptr, unsigned char* handle64) {
  if (!ptr || !handle64) return B2048_EINVAL;
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, const_cast<void*>(ptr));
  if (e != cudaSuccess) return (int)e;
  memcpy(handle64, &h, sizeof(h));
  return B2048_OK;
}

extern "C" int p2p_allreduce_adam_f64(const double* const* peer_grads, uint64_t* const* peer_flags,
                                      uint64_t* sync_state, int rank, int world, double* params,
                                      double* exp_avg, double* exp_avg_sq, int64_t* step_counter, int64_t n,
                                      double lr, double beta1, double beta2, double eps, void* stream) {
  if (!peer_grads || !peer_flags || !sync_state || !params || !exp_avg || !exp_avg_sq || !step_counter)
    return B2048_EINVAL;
  if (n <= 0 || world < 1 || world > P2P_THREADS || rank < 0 || rank >= world) return B2048_EINVAL;
  int err = 0;
  DeviceCtx* ctx = current_ctx(&err);
  if (!ctx) return err;
  int64_t blocks = (n + P2P_THREADS - 1) / P2P_THREADS;
  if (blocks > ctx->sm_count) blocks = ctx->sm_count;   // all blocks co-resident: the last-block barrier cannot starve
  p2p_allreduce_adam_kernel<<<(unsigned)blocks, P2P_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      peer_grads, peer_flags, sync_state, rank, world, params, exp_avg, exp_avg_sq, step_counter, n, lr, beta1,
      beta2, eps);
  return (int)cudaGetLastError();
}
