// host_api.cu — library lifetime, row-table construction and the host-buffer step entry point.
#include <ctype.h>
#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <vector>

#include "b2048_common.cuh"

namespace b2048 {

cudaError_t env_kernels_configure();
cudaError_t qnet_kernels_configure();
cudaError_t wgrad_kernels_configure();
int step_device(DeviceCtx* ctx, const uint64_t* boards, const uint8_t* actions, uint64_t* next,
                int32_t* reward, uint8_t* flags, int64_t n, uint64_t seed, uint64_t step,
                uint64_t index_base, uint32_t p4, const uint8_t* ovr, cudaStream_t st);

static DeviceCtx g_ctx[MAX_DEVICES];
static std::mutex g_mu;
static std::mutex g_host_mu[MAX_DEVICES];   // b2048_step_host owns the device's staging slots for the whole call
static std::vector<uint32_t> g_host_lut;

// Canonical 2048 row move toward index 0 on four 4-bit exponents: compress, merge each tile at
// most once (leftmost pair first), compress.  Equals the reference's while-loop
// (src/board.py:92-126) on every one of the 65536 rows (tests/test_oracle_golden.py proves it
// against the reference's own outputs).  reward = sum of merged tile values (src/board.py:114).
static void move_row(const int in[4], int out[4], uint32_t* reward, bool* overflow) {
  int t[4], nt = 0;
  for (int c = 0; c < 4; ++c)
    if (in[c]) t[nt++] = in[c];
  int no = 0;
  out[0] = out[1] = out[2] = out[3] = 0;
  for (int i = 0; i < nt; ++i) {
    if (i + 1 < nt && t[i] == t[i + 1]) {
      const int m = t[i] + 1;
      *reward += 1u << m;
      if (m > 15) *overflow = true;
      out[no++] = m & 0xF;
      ++i;
    } else {
      out[no++] = t[i];
    }
  }
}

static uint32_t build_row_entry(uint32_t row) {
  int in[4], rev[4], out[4], tmp[4];
  for (int c = 0; c < 4; ++c) {
    in[c] = (row >> (4 * c)) & 0xF;
    rev[3 - c] = in[c];
  }
  uint32_t reward = 0, r2 = 0;
  bool overflow = false, o2 = false;
  move_row(in, out, &reward, &overflow);
  move_row(rev, tmp, &r2, &o2);                       // the same row moved RIGHT
  const bool can_right = tmp[0] != rev[0] || tmp[1] != rev[1] || tmp[2] != rev[2] || tmp[3] != rev[3];
  uint32_t res = 0;
  for (int c = 0; c < 4; ++c) res |= (uint32_t)out[c] << (4 * c);
  // 14-bit reward/4: exact for every non-overflow row except 0xEEEE (65536), see b2048_common.cuh
  const uint32_t r4 = (reward >> 2) & 0x3FFFu;
  return res | (r4 << 16) | (can_right ? 0x40000000u : 0u) | (overflow ? 0x80000000u : 0u);
}

// The staged (shared-memory) copy keeps reward/4 below 2^12 (bits 28-29 zero), so that the streaming kernel can
// sum the four rows' upper halves UNMASKED -- 4 * sum(reward/4) stays below 2^16 and the RIGHT / OVERFLOW bits land
// at bit 16 and above -- and mask once after the sum.  A row whose merges are worth 16384 or more (two 8192 tiles, or
// two pairs of 4096 tiles: never seen in play) gets the OVERFLOW bit instead: its quad is redone from the plain table.
static uint32_t staged_entry(uint32_t e) {
  const uint32_t r4 = (e >> 16) & 0x3FFFu;
  return r4 < 4096u ? e : ((e & 0xCFFFFFFFu) | ENTRY_OVF);
}

static const std::vector<uint32_t>& host_lut() {
  if (g_host_lut.empty()) {
    g_host_lut.resize(LUT_ROWS);
    for (uint32_t r = 0; r < (uint32_t)LUT_ROWS; ++r) g_host_lut[r] = build_row_entry(r);
  }
  return g_host_lut;
}

DeviceCtx* ctx_for(int device) {
  if (device < 0 || device >= MAX_DEVICES) return nullptr;
  return &g_ctx[device];
}

DeviceCtx* current_ctx(int* err) {
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    cudaGetLastError();
    if (err) *err = B2048_ENODEV;
    return nullptr;
  }
  DeviceCtx* c = ctx_for(dev);
  if (!c || !c->ready) {
    if (err) *err = B2048_ENOTINIT;
    return nullptr;
  }
  return c;
}

}  // namespace b2048

using namespace b2048;

extern "C" int b2048_abi_version(void) { return B2048_ABI_VERSION; }

extern "C" const char* b2048_error_string(int code) {
  switch (code) {
    case B2048_OK: return "ok";
    case B2048_ENOTINIT: return "b2048_init(device) has not been called for the current device";
    case B2048_EINVAL: return "invalid argument";
    case B2048_ENODEV: return "no usable CUDA device (this library has no CPU fallback)";
    default: break;
  }
  if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
  return "unknown b2048 error";
}

extern "C" int b2048_init(int device) {
  std::lock_guard<std::mutex> lock(g_mu);
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    cudaGetLastError();
    return B2048_ENODEV;
  }
  if (device < 0 || device >= count || device >= MAX_DEVICES) return B2048_EINVAL;
  DeviceCtx* c = &g_ctx[device];
  if (c->ready) return B2048_OK;
  int prev = 0;
  cudaGetDevice(&prev);
  if ((e = cudaSetDevice(device)) != cudaSuccess) return (int)e;
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return (int)e;
  c->sm_count = prop.multiProcessorCount;
  c->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
  const std::vector<uint32_t>& lut = host_lut();
  std::vector<uint32_t> both(lut.begin(), lut.end());      // plain table, then the swizzled smem image
  both.resize(LUT_ROWS + LUT_SMEM_ROWS);
  for (uint32_t i = 0; i < (uint32_t)LUT_SMEM_ROWS; ++i) both[LUT_ROWS + lut_swizzle(i)] = staged_entry(lut[i]);
  if ((e = cudaMalloc(&c->lut, both.size() * sizeof(uint32_t))) != cudaSuccess) return (int)e;
  if ((e = cudaMemcpy(c->lut, both.data(), both.size() * sizeof(uint32_t), cudaMemcpyHostToDevice)) !=
      cudaSuccess)
    return (int)e;
  if ((e = env_kernels_configure()) != cudaSuccess) return (int)e;
  if ((e = qnet_kernels_configure()) != cudaSuccess) return (int)e;
  if ((e = wgrad_kernels_configure()) != cudaSuccess) return (int)e;
  if ((e = cudaDeviceSynchronize()) != cudaSuccess) return (int)e;
  c->ready = true;
  cudaSetDevice(prev);
  return B2048_OK;
}

extern "C" int b2048_shutdown(int device) {
  std::lock_guard<std::mutex> lock(g_mu);
  DeviceCtx* c = ctx_for(device);
  if (!c) return B2048_EINVAL;
  if (!c->ready) return B2048_OK;
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(device);
  cudaDeviceSynchronize();
  cudaFree(c->lut);
  if (c->ws) cudaFree(c->ws);
  for (int i = 0; i < 3; ++i) {
    if (c->ws_streams[i]) cudaStreamDestroy(c->ws_streams[i]);
    if (c->ws_events[i]) cudaEventDestroy(c->ws_events[i]);
  }
  *c = DeviceCtx();
  cudaSetDevice(prev);
  return B2048_OK;
}

extern "C" int b2048_copy_row_lut_host(uint32_t* out65536) {
  if (!out65536) return B2048_EINVAL;
  std::lock_guard<std::mutex> lock(g_mu);
  const std::vector<uint32_t>& lut = host_lut();
  memcpy(out65536, lut.data(), LUT_ROWS * sizeof(uint32_t));
  return B2048_OK;
}

// ---- NUMA-local pinned host memory ------------------------------------------------------------------
// cudaHostAlloc places pages on the NUMA node of the calling thread (first touch under the default
// policy).  A rank that happens to run on the far socket therefore stages its e2e buffers across the
// inter-socket link: with eight ranks that halves the PCIe rate of half the GPUs.  b2048_host_alloc binds
// the calling thread to the CPUs next to `device` (sysfs: /sys/bus/pci/devices/<id>/local_cpulist) for the
// duration of the allocation and the first touch, then restores the previous affinity.  No libnuma needed.

namespace {
// parse "0-15,32-47" into a cpu_set_t; returns the number of CPUs found
int parse_cpulist(const char* txt, cpu_set_t* set) {
  CPU_ZERO(set);
  int count = 0;
  const char* p = txt;
  while (*p) {
    while (*p && !isdigit((unsigned char)*p)) ++p;
    if (!*p) break;
    char* end = nullptr;
    long a = strtol(p, &end, 10), b = a;
    p = end;
    if (*p == '-') {
      b = strtol(p + 1, &end, 10);
      p = end;
    }
    for (long c = a; c <= b && c < CPU_SETSIZE; ++c) {
      CPU_SET((int)c, set);
      ++count;
    }
  }
  return count;
}

// CPUs local to `device` intersected with the CPUs this process may use; false if unknown
bool local_cpus(int device, cpu_set_t* out, int* numa_node) {
  char bus[32] = {0};
  if (cudaDeviceGetPCIBusId(bus, sizeof(bus), device) != cudaSuccess) return false;
  for (char* c = bus; *c; ++c) *c = (char)tolower((unsigned char)*c);
  char path[128], buf[4096] = {0};
  snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/local_cpulist", bus);
  FILE* f = fopen(path, "r");
  if (!f) return false;
  const size_t got = fread(buf, 1, sizeof(buf) - 1, f);
  fclose(f);
  if (got == 0) return false;
  cpu_set_t local, allowed;
  if (parse_cpulist(buf, &local) == 0) return false;
  if (sched_getaffinity(0, sizeof(allowed), &allowed) != 0) return false;
  CPU_AND(out, &local, &allowed);
  if (numa_node) {
    *numa_node = -1;
    snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", bus);
    if ((f = fopen(path, "r"))) {
      if (fscanf(f, "%d", numa_node) != 1) *numa_node = -1;
      fclose(f);
    }
  }
  return CPU_COUNT(out) > 0;
}
}  // namespace

extern "C" int b2048_host_alloc(void** out, size_t bytes, int device, int* numa_node_out, int* bound_out) {
  if (!out || bytes == 0) return B2048_EINVAL;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
    cudaGetLastError();
    return B2048_ENODEV;
  }
  if (device < 0 || device >= count) return B2048_EINVAL;
  cpu_set_t prev, want;
  int node = -1;
  const bool have_prev = sched_getaffinity(0, sizeof(prev), &prev) == 0;
  const char* off = getenv("B2048_NO_NUMA_BIND");          // A/B switch for profiles/pinned_ab.py
  const bool bind = !(off && off[0] == '1') && have_prev && local_cpus(device, &want, &node) &&
                    sched_setaffinity(0, sizeof(want), &want) == 0;
  int prev_dev = 0;
  cudaGetDevice(&prev_dev);
  cudaSetDevice(device);
  cudaError_t e = cudaHostAlloc(out, bytes, cudaHostAllocPortable);
  if (e == cudaSuccess) memset(*out, 0, bytes);      // first touch while bound
  cudaSetDevice(prev_dev);
  if (bind) sched_setaffinity(0, sizeof(prev), &prev);
  if (numa_node_out) *numa_node_out = node;
  if (bound_out) *bound_out = bind ? 1 : 0;
  return (int)e;
}

extern "C" int b2048_host_free(void* p) {
  if (!p) return B2048_OK;
  return (int)cudaFreeHost(p);
}

// Bind the CALLING thread to the CPUs next to `device` (the thread that drives b2048_step_host's copies).
// Returns 0 if bound, B2048_EINVAL if the topology is unknown (nothing changed).
extern "C" int b2048_bind_thread_near(int device) {
  cpu_set_t want;
  int node = -1;
  if (!local_cpus(device, &want, &node)) return B2048_EINVAL;
  return sched_setaffinity(0, sizeof(want), &want) == 0 ? B2048_OK : B2048_EINVAL;
}

// ---- host-buffer step: chunked 3-stage pipeline (H2D | kernel | D2H) -------------------------------
namespace {
constexpr int64_t HOST_CHUNK = 4 << 20;  // boards per chunk
constexpr int HOST_SLOTS = 3;
// per-board device bytes: board 8 + next 8 + reward 4 + action 1 + flags 1 + override 1
constexpr size_t SLOT_BYTES = (size_t)HOST_CHUNK * (8 + 8 + 4 + 1 + 1 + 1);
}  // namespace

extern "C" int b2048_step_host(const uint64_t* h_boards, const uint8_t* h_actions, uint64_t* h_next,
                               int32_t* h_reward, uint8_t* h_flags, int64_t n, uint64_t seed,
                               uint64_t step, uint64_t index_base, uint32_t p4_threshold,
                               const uint8_t* h_spawn_override, int device) {
  if (n < 0) return B2048_EINVAL;
  if (n == 0) return B2048_OK;
  if (!h_boards || !h_actions || !h_next || !h_reward || !h_flags) return B2048_EINVAL;
  DeviceCtx* c = ctx_for(device);
  if (!c || !c->ready) return B2048_ENOTINIT;
  cudaError_t e;
  int prev = 0;
  cudaGetDevice(&prev);
  if ((e = cudaSetDevice(device)) != cudaSuccess) return (int)e;
  std::lock_guard<std::mutex> host_lock(g_host_mu[device]);   // concurrent host calls on one device take turns
  {
    std::lock_guard<std::mutex> lock(g_mu);
    if (!c->ws) {
      if ((e = cudaMalloc(&c->ws, SLOT_BYTES * HOST_SLOTS)) != cudaSuccess) return (int)e;
      c->ws_bytes = SLOT_BYTES * HOST_SLOTS;
      for (int i = 0; i < HOST_SLOTS; ++i) {
        if ((e = cudaStreamCreateWithFlags(&c->ws_streams[i], cudaStreamNonBlocking)) != cudaSuccess)
          return (int)e;
        if ((e = cudaEventCreateWithFlags(&c->ws_events[i], cudaEventDisableTiming)) != cudaSuccess)
          return (int)e;
      }
    }
  }
  int rc = B2048_OK;
  int64_t off = 0;
  int slot = 0;
  while (off < n && rc == B2048_OK) {
    const int64_t m = (n - off < HOST_CHUNK) ? (n - off) : HOST_CHUNK;
    cudaStream_t st = c->ws_streams[slot];
    unsigned char* base = static_cast<unsigned char*>(c->ws) + SLOT_BYTES * slot;
    uint64_t* d_b = reinterpret_cast<uint64_t*>(base);
    uint64_t* d_n = d_b + HOST_CHUNK;
    int32_t* d_r = reinterpret_cast<int32_t*>(d_n + HOST_CHUNK);
    uint8_t* d_a = reinterpret_cast<uint8_t*>(d_r + HOST_CHUNK);
    uint8_t* d_f = d_a + HOST_CHUNK;
    uint8_t* d_o = d_f + HOST_CHUNK;
    // the slot's previous D2H copies are ordered before these H2D copies by the stream itself
    if ((e = cudaMemcpyAsync(d_b, h_boards + off, m * 8, cudaMemcpyHostToDevice, st)) != cudaSuccess) { rc = (int)e; break; }
    if ((e = cudaMemcpyAsync(d_a, h_actions + off, m, cudaMemcpyHostToDevice, st)) != cudaSuccess) { rc = (int)e; break; }
    if (h_spawn_override &&
        (e = cudaMemcpyAsync(d_o, h_spawn_override + off, m, cudaMemcpyHostToDevice, st)) != cudaSuccess) { rc = (int)e; break; }
    rc = step_device(c, d_b, d_a, d_n, d_r, d_f, m, seed, step, index_base + (uint64_t)off,
                     p4_threshold, h_spawn_override ? d_o : nullptr, st);
    if (rc != B2048_OK) break;
    if ((e = cudaMemcpyAsync(h_next + off, d_n, m * 8, cudaMemcpyDeviceToHost, st)) != cudaSuccess) { rc = (int)e; break; }
    if ((e = cudaMemcpyAsync(h_reward + off, d_r, m * 4, cudaMemcpyDeviceToHost, st)) != cudaSuccess) { rc = (int)e; break; }
    if ((e = cudaMemcpyAsync(h_flags + off, d_f, m, cudaMemcpyDeviceToHost, st)) != cudaSuccess) { rc = (int)e; break; }
    off += m;
    slot = (slot + 1) % HOST_SLOTS;
  }
  for (int i = 0; i < HOST_SLOTS; ++i) {
    e = cudaStreamSynchronize(c->ws_streams[i]);
    if (e != cudaSuccess && rc == B2048_OK) rc = (int)e;
  }
  cudaSetDevice(prev);
  return rc;
}
